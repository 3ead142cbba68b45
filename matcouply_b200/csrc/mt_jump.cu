// Host-side jump-ahead of a NumPy `RandomState` (MT19937) stream: advance (key, pos) by n output words in O(1) of n.
//
// Why: the reference draws the whole initial state from ONE RandomState (decomposition.py:31-39, 78-89), and a rank of a
// sharded run owns only a contiguous range of the rows of every B-mode array.  Walking the global stream on every rank
// (rng.cu: one CTA, sequential) makes the initialisation cost grow with the number of ranks; with this jump a rank skips
// the rows before and after its own share and generates only those (bit-identical to the full walk).
//
// How (Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer 2008): the generator is a linear map F over GF(2) on the
// sliding window (x_k .. x_{k+623}); its minimal polynomial is t * phi(t), phi the degree-19937 characteristic polynomial
// of MT19937 (135 non-zero terms, table below: computed with Berlekamp-Massey from the generator's own output and
// re-checked against NumPy in tests/test_host.py).  g(t) = t^J mod phi(t) by square-and-shift with sparse folding,
// then g(F) applied to the window by Horner's rule.  g(F) s and F^J s differ at most in the low 31 bits of the first
// word (the part of the window the recurrence never reads), so the jump targets the block BEFORE the wanted one and the
// last block is regenerated with the ordinary recurrence: the resulting 624-word key is exact.
//
// Pure host code (no CUDA call): state_io is a HOST pointer.
#include <cstdint>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int kN = 624, kM = 397, kDeg = 19937;
constexpr uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrixA = 0x9908b0dfu;
constexpr int kPolyWords = 624;  // 64-bit words: room for a square (degree < 2 * 19937)

// exponents of phi(t), ascending; the last one is the leading term t^19937
const int kPhiExps[135] = {
    0, 1189, 1416, 1585, 1643, 1870, 2493, 2773, 3000, 3227, 3454, 3681,
    3908, 4135, 4362, 4753, 5661, 6337, 6569, 7129, 7477, 7525, 7583, 7752,
    7979, 8206, 9505, 9901, 9969, 10128, 10693, 10761, 10920, 11089, 11147, 11157,
    11215, 11321, 11374, 11384, 11485, 11611, 11712, 11717, 11838, 11881, 11944, 11997,
    12277, 12335, 12393, 12504, 12509, 12620, 12673, 12731, 12736, 12789, 12905, 12958,
    12963, 13137, 13185, 13190, 13243, 13301, 13412, 13528, 13533, 13639, 13697, 13760,
    13813, 13866, 14093, 14151, 14209, 14320, 14325, 14436, 14547, 14552, 14605, 14721,
    14774, 14779, 14953, 15001, 15006, 15059, 15117, 15228, 15344, 15349, 15455, 15513,
    15576, 15629, 15682, 15909, 15967, 16025, 16136, 16141, 16252, 16363, 16368, 16421,
    16537, 16590, 16595, 16817, 16822, 16875, 16933, 17044, 17160, 17271, 17329, 17445,
    17498, 17725, 17783, 17841, 17952, 18068, 18179, 18237, 18406, 18633, 18691, 18860,
    19087, 19314, 19937,
};

inline uint32_t twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & kUpper) | (nxt & kLower);
    return far ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
}

// one block of the ordinary generator (mt19937_gen): key <- the next 624 words
void next_block(uint32_t* mt) {
    int k = 0;
    for (; k < kN - kM; ++k) mt[k] = twist(mt[k], mt[k + 1], mt[k + kM]);
    for (; k < kN - 1; ++k) mt[k] = twist(mt[k], mt[k + 1], mt[k + (kM - kN)]);
    mt[kN - 1] = twist(mt[kN - 1], mt[0], mt[kM - 1]);
}

// p (degree < 2 * 19937) <- p mod phi: fold 64-bit words from the top; every image lands >= 623 bits lower
void reduce(uint64_t* p) {
    const int top_word = kDeg / 64, top_bit = kDeg % 64;  // bit 19937 = word 311, bit 33
    for (int wi = kPolyWords - 1; wi >= top_word; --wi) {
        uint64_t v;
        long long base;  // exponent offset d of bit 0 of v: t^(19937 + d + b) for bit b
        if (wi > top_word) {
            v = p[wi];
            p[wi] = 0;
            base = (long long)wi * 64 - kDeg;
        } else {
            v = p[wi] >> top_bit;
            p[wi] &= (~0ull) >> (64 - top_bit);
            base = 0;
        }
        if (!v) continue;
        for (int i = 0; i < 134; ++i) {
            const long long o = base + kPhiExps[i];
            const int w = (int)(o >> 6), s = (int)(o & 63);
            p[w] ^= v << s;
            if (s) p[w + 1] ^= v >> (64 - s);
        }
    }
}

struct SpreadTable {  // byte -> its 8 bits interleaved with zeros (squaring over GF(2))
    uint16_t v[256];
    SpreadTable() {
        for (int b = 0; b < 256; ++b) {
            uint16_t s = 0;
            for (int i = 0; i < 8; ++i) s |= (uint16_t)(((b >> i) & 1) << (2 * i));
            v[b] = s;
        }
    }
};

void square(uint64_t* p) {  // p (degree < 19937, words 0..311) <- p^2 (bits interleaved with zeros)
    static const SpreadTable table;  // thread-safe one-time initialisation
    const uint16_t* g_spread = table.v;
    for (int wi = kPolyWords / 2 - 1; wi >= 0; --wi) {
        const uint64_t v = p[wi];
        uint64_t lo = 0, hi = 0;
        for (int b = 0; b < 4; ++b) {
            lo |= (uint64_t)g_spread[(v >> (8 * b)) & 0xff] << (16 * b);
            hi |= (uint64_t)g_spread[(v >> (32 + 8 * b)) & 0xff] << (16 * b);
        }
        p[2 * wi] = lo;
        p[2 * wi + 1] = hi;
    }
}

// g <- t^J mod phi
void jump_polynomial(unsigned long long J, uint64_t* g) {
    std::memset(g, 0, sizeof(uint64_t) * kPolyWords);
    g[0] = 1;
    int top = 63;
    while (top > 0 && !((J >> top) & 1ull)) --top;
    for (int b = top; b >= 0; --b) {
        square(g);
        reduce(g);
        if ((J >> b) & 1ull) {  // times t
            for (int wi = kDeg / 64 + 1; wi > 0; --wi) g[wi] = (g[wi] << 1) | (g[wi - 1] >> 63);
            g[0] <<= 1;
            reduce(g);
        }
    }
}

// window <- g(F) window (Horner; exact except for the low 31 bits of word 0)
void apply_polynomial(const uint64_t* g, uint32_t* window) {
    uint32_t h[kN];
    std::memset(h, 0, sizeof(h));
    int head = 0;  // logical word j lives at h[(head + j) % 624]
    int deg = kDeg - 1;
    while (deg >= 0 && !((g[deg >> 6] >> (deg & 63)) & 1ull)) --deg;
    for (int i = deg; i >= 0; --i) {
        // h <- F h: one step of the recurrence on the sliding window
        const uint32_t x0 = h[head], x1 = h[head + 1 < kN ? head + 1 : head + 1 - kN];
        const uint32_t xm = h[head + kM < kN ? head + kM : head + kM - kN];
        h[head] = twist(x0, x1, xm);  // the new last word takes the slot of the dropped first word
        head = head + 1 < kN ? head + 1 : 0;
        if ((g[i >> 6] >> (i & 63)) & 1ull) {
            const int first = kN - head;  // logical words 0..first-1 are h[head..623]
            for (int j = 0; j < first; ++j) h[head + j] ^= window[j];
            for (int j = first; j < kN; ++j) h[j - first] ^= window[j];
        }
    }
    for (int j = 0; j < kN; ++j) window[j] = h[head + j < kN ? head + j : head + j - kN];
}

}  // namespace

extern "C" int b2_mt19937_jump_host(unsigned* state_io, unsigned long long n_words) {
    B2_REQUIRE(state_io != nullptr, "b2_mt19937_jump_host: null state");
    uint32_t* key = state_io;
    const unsigned pos = state_io[kN];
    B2_REQUIRE(pos <= (unsigned)kN, "b2_mt19937_jump_host: position %u out of range", pos);
    if (n_words == 0) return B2_OK;
    const unsigned long long target = (unsigned long long)pos + n_words;  // words from the start of the current block
    unsigned long long q = target / kN;
    unsigned r = (unsigned)(target % kN);
    if (r == 0) {  // the generator regenerates lazily: a stream position at a block boundary is (previous block, 624)
        q -= 1;
        r = kN;
    }
    if (q > 0) {
        unsigned long long blocks_by_recurrence = q;
        if (q > 64) {  // jump to block q - 1, then one ordinary block makes every word of the key exact
            static thread_local uint64_t g[kPolyWords];
            jump_polynomial((q - 1) * (unsigned long long)kN, g);
            apply_polynomial(g, key);
            blocks_by_recurrence = 1;
        }
        for (unsigned long long b = 0; b < blocks_by_recurrence; ++b) next_block(key);
    }
    state_io[kN] = r;
    return B2_OK;
}
