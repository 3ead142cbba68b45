// Device-side continuation of a NumPy `RandomState` (MT19937) stream.
//
// The reference draws every initial factor / auxiliary / dual variable from ONE host `np.random.RandomState`
// (decomposition.py:31-39, 78-89; penalties.py:125-147, 239-261), and bit-parity of the whole trajectory hangs on that
// stream.  For the B-mode variables that is sum_i J_i * R doubles per array — seconds of single-threaded host time and
// a host->device copy at BASELINE sizes.  MT19937 is a linear recurrence over a 624-word state: here one CTA advances
// the state exactly like the host generator would (three dependent phases of ~227 independent words per block) and
// writes the tempered 32-bit words straight into the output buffer; a second, fully parallel kernel turns each pair of
// words into the double `random_sample()` would have returned ((a >> 5) * 2^26 + (b >> 6)) / 2^53 — identical bits.
// The advanced state goes back to the host generator, so later host draws continue the same stream.
#include "common.cuh"

namespace {

constexpr int kN = 624, kM = 397;
constexpr unsigned kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrixA = 0x9908b0dfu;

__device__ __forceinline__ unsigned temper(unsigned y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
__device__ __forceinline__ unsigned twist(unsigned cur, unsigned nxt, unsigned far) {
    const unsigned y = (cur & kUpper) | (nxt & kLower);
    return far ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
}

// state_io: 624 key words + position (0..624).  words: n_words tempered outputs, in stream order.
__global__ void __launch_bounds__(256) mt19937_words_kernel(unsigned* __restrict__ state_io, unsigned* __restrict__ words,
                                                            long long n_words) {
    __shared__ unsigned bufs[2][kN];
    const int tid = threadIdx.x;
    unsigned* mt = bufs[0];
    unsigned* nx = bufs[1];
    for (int k = tid; k < kN; k += blockDim.x) mt[k] = state_io[k];
    int pos = (int)state_io[kN];
    __syncthreads();
    long long done = 0;
    while (done < n_words) {
        if (pos >= kN) {  // regenerate the block (mt19937_gen of the host generator), three dependent phases
            for (int k = tid; k < kN - kM; k += blockDim.x) nx[k] = twist(mt[k], mt[k + 1], mt[k + kM]);
            __syncthreads();
            for (int k = kN - kM + tid; k < 2 * (kN - kM); k += blockDim.x) nx[k] = twist(mt[k], mt[k + 1], nx[k - (kN - kM)]);
            __syncthreads();
            for (int k = 2 * (kN - kM) + tid; k < kN; k += blockDim.x)
                nx[k] = twist(mt[k], k + 1 < kN ? mt[k + 1] : nx[0], nx[k - (kN - kM)]);
            __syncthreads();
            unsigned* t = mt;
            mt = nx;
            nx = t;
            pos = 0;
        }
        const long long left = n_words - done;
        const int take = (int)((long long)(kN - pos) < left ? (kN - pos) : left);
        for (int k = tid; k < take; k += blockDim.x) words[done + k] = temper(mt[pos + k]);
        done += take;
        pos += take;
        // the block is regenerated from `mt` only after every thread passed the barriers above; reads here are safe
    }
    __syncthreads();
    for (int k = tid; k < kN; k += blockDim.x) state_io[k] = mt[k];
    if (tid == 0) state_io[kN] = (unsigned)pos;
}

// in place: 8 bytes (a, b) -> double (a >> 5, b >> 6)   (legacy mt19937_next_double / random_sample)
__global__ void words_to_double_kernel(double* __restrict__ io, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint2 w = *(const uint2*)(io + i);
        const double a = (double)(w.x >> 5), b = (double)(w.y >> 6);
        io[i] = (a * 67108864.0 + b) / 9007199254740992.0;
    }
}

}  // namespace

extern "C" int b2_mt19937_uniform(void* state_io, double* out, long long n, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B2_REQUIRE(state_io != nullptr && (out != nullptr || n == 0) && n >= 0, "b2_mt19937_uniform: bad arguments");
    if (n == 0) return B2_OK;
    mt19937_words_kernel<<<1, 256, 0, st>>>((unsigned*)state_io, (unsigned*)out, 2 * n);
    B2_LAUNCH_CHECK();
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)b2_num_sms() * 16) blocks = (long long)b2_num_sms() * 16;
    words_to_double_kernel<<<(int)blocks, 256, 0, st>>>(out, n);
    B2_LAUNCH_CHECK();
    return B2_OK;
}
