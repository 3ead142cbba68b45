// Library-level plumbing of the C ABI: error reporting, version, device queries, micro-benchmarks used to
// measure the roofline denominators (fp64 FMA / DMMA / fp32 FMA peak) on the box the benchmark runs on.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_b2_launches = 0;

void b2_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int b2_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

static int g_options[B2_OPT_COUNT] = {2, 2, 1, 0, 14};  // XSTREAM_HYBRID measured slower than padded DMMA (profiles/): off
int b2_option_value(int option) { return (option >= 0 && option < B2_OPT_COUNT) ? g_options[option] : 0; }

namespace {

template <int KIND>
__global__ void __launch_bounds__(256) flops_kernel(int iters, double* sink) {
    if (KIND == 0) {
        double acc[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[u] = threadIdx.x * 1e-3 + u;
        const double a = 1.0000001, b = 1e-9;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 16; ++u) acc[u] = fma(acc[u], a, b);
        }
        double s = 0;
#pragma unroll
        for (int u = 0; u < 16; ++u) s += acc[u];
        if (s == 123.456) sink[0] = s;
    } else if (KIND == 1) {
        double c[8][2];
#pragma unroll
        for (int u = 0; u < 8; ++u) c[u][0] = c[u][1] = threadIdx.x * 1e-3 + u;
        const double a = 1.0000001 + threadIdx.x * 1e-9, b = 1e-9 * (threadIdx.x + 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) dmma884(c[u][0], c[u][1], a, b);
        }
        double s = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
        if (s == 123.456) sink[0] = s;
    } else if (KIND == 3) {
        // DMMA and DFMA interleaved (8 DMMA + 16 DFMA per iteration): do the two pipes overlap?
        double c[8][2], acc[16];
#pragma unroll
        for (int u = 0; u < 8; ++u) c[u][0] = c[u][1] = threadIdx.x * 1e-3 + u;
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[u] = threadIdx.x * 1e-3 + u;
        const double a = 1.0000001 + threadIdx.x * 1e-9, b = 1e-9 * (threadIdx.x + 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                dmma884(c[u][0], c[u][1], a, b);
                acc[2 * u] = fma(acc[2 * u], a, b);
                acc[2 * u + 1] = fma(acc[2 * u + 1], a, b);
            }
        }
        double s = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
#pragma unroll
        for (int u = 0; u < 16; ++u) s += acc[u];
        if (s == 123.456) sink[0] = s;
    } else {
        float acc[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[u] = threadIdx.x * 1e-3f + u;
        const float a = 1.0000001f, b = 1e-9f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 16; ++u) acc[u] = fmaf(acc[u], a, b);
        }
        float s = 0;
#pragma unroll
        for (int u = 0; u < 16; ++u) s += acc[u];
        if (s == 123.456f) sink[0] = s;
    }
}

}  // namespace

extern "C" {

const char* b2_last_error(void) { return g_err; }
int b2_version(void) { return 100; }
int b2_device_sm_count(void) { return b2_num_sms(); }
unsigned long long b2_launch_count(void) { return g_b2_launches; }

int b2_set_option(int option, int value) {
    B2_REQUIRE(option >= 0 && option < B2_OPT_COUNT, "unknown option %d", option);
    g_options[option] = value;
    return B2_OK;
}
int b2_get_option(int option) { return b2_option_value(option); }

int b2_microbench_flops(int kind, int iters, double* flops_host, void* sink, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = b2_num_sms() * 4, threads = 256;
    B2_REQUIRE(kind >= 0 && kind <= 3, "kind must be 0 (fp64 FMA), 1 (DMMA), 2 (fp32 FMA) or 3 (DMMA + DFMA mixed)");
    if (kind == 0) {
        flops_kernel<0><<<blocks, threads, 0, st>>>(iters, (double*)sink);
        *flops_host = (double)blocks * threads * (double)iters * 16.0 * 2.0;
    } else if (kind == 1) {
        flops_kernel<1><<<blocks, threads, 0, st>>>(iters, (double*)sink);
        *flops_host = (double)blocks * (threads / 32) * (double)iters * 8.0 * 512.0;
    } else if (kind == 3) {
        flops_kernel<3><<<blocks, threads, 0, st>>>(iters, (double*)sink);
        *flops_host = (double)blocks * (threads / 32) * (double)iters * 8.0 * 512.0 +
                      (double)blocks * threads * (double)iters * 16.0 * 2.0;
    } else {
        flops_kernel<2><<<blocks, threads, 0, st>>>(iters, (double*)sink);
        *flops_host = (double)blocks * threads * (double)iters * 16.0 * 2.0;
    }
    B2_LAUNCH_CHECK();
    return B2_OK;
}

}  // extern "C"
