// HOST-only (g++): vectorised inner loops of the np.random.permutation stream, see hostperm_simd.h.
//
// numpy's legacy shuffle (numpy/random/mtrand.pyx, RandomState.shuffle -> _shuffle_raw; randomkit's rk_interval) walks
// i = n-1 .. 1 and draws j = random_interval(i): mask = smallest 2^k - 1 >= i, redraw (next 32-bit word & mask) until it is
// <= i.  Everything here reproduces that stream bit for bit; only the evaluation order inside a group of draws differs:
// for a group of G consecutive words taken while i stays in one mask range, a word v <= i - G is accepted whatever the
// earlier ones did (i drops by at most one per draw), a word v > i is rejected whatever they did, and if no word of the
// group falls in (i - G, i] the whole group is decided by two vector compares.  The accepted ones are compacted in draw
// order; i drops by their count.  A group with a word in the uncertain band (probability ~ G^2 / mask) is walked scalar.
#include "hostperm_simd.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

namespace tsb_hp {

int detect_isa() {
    __builtin_cpu_init();
    int best = kScalar;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt")) best = kAvx2;
    if (best == kAvx2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512bw"))
        best = kAvx512;
    const char* e = getenv("TS_B200_PERM_ISA");
    if (e != nullptr) {
        int want = best;
        if (strcmp(e, "scalar") == 0) want = kScalar;
        else if (strcmp(e, "avx2") == 0) want = kAvx2;
        else if (strcmp(e, "avx512") == 0) want = kAvx512;
        if (want < best) best = want;
    }
    return best;
}

__attribute__((target_clones("avx512f", "avx2", "default")))
void mt_next(const uint32_t* __restrict__ in, uint32_t* __restrict__ out) {
    constexpr uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
    constexpr int M = 397;
    for (int i = 0; i < kMtN - M; ++i) {             // partner word i + 397: still the old state
        const uint32_t y = (in[i] & UP) | (in[i + 1] & LO);
        out[i] = in[i + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
    }
    for (int i = kMtN - M; i < kMtN - 1; ++i) {      // partner word i - 227: already the new state (written 227 iterations ago)
        const uint32_t y = (in[i] & UP) | (in[i + 1] & LO);
        out[i] = out[i - (kMtN - M)] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
    }
    const uint32_t y = (in[kMtN - 1] & UP) | (out[0] & LO);
    out[kMtN - 1] = out[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
}

__attribute__((target_clones("avx512f", "avx2", "default")))
void mt_temper(const uint32_t* __restrict__ key, uint32_t* __restrict__ out) {
    for (int d = 0; d < kMtN; ++d) {
        uint32_t y = key[d];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        out[d] = y;
    }
}

namespace {

inline uint32_t mask_of(int64_t i) { return 0xffffffffu >> __builtin_clz((uint32_t)i); }

// One draw per iteration, no data-dependent branch (a do/while rejection loop mispredicts ~30 % of the time).
inline int walk_scalar_span(const uint32_t* t, int count, uint32_t mask, int64_t& i, uint32_t*& cur) {
    for (int d = 0; d < count; ++d) {
        const uint32_t v = t[d] & mask;
        *cur = v;
        const int64_t acc = (int64_t)(v <= (uint32_t)i);
        cur += acc; i -= acc;
    }
    return count;
}

int walk_scalar(const uint32_t* t, int avail, Walk* w) {
    int64_t i = w->i; uint32_t* cur = w->cur;
    int d = 0;
    while (d < avail && i >= 1) {
        const uint32_t mask = mask_of(i);
        const int64_t lower = (int64_t)(mask >> 1);
        for (; d + 4 <= avail && i - 4 > lower; d += 4) walk_scalar_span(t + d, 4, mask, i, cur);
        for (; d < avail && i > lower; ++d) walk_scalar_span(t + d, 1, mask, i, cur);
    }
    w->i = i; w->cur = cur;
    return d;
}

__attribute__((target("avx512f,avx512vl,avx512bw,popcnt")))
int walk_avx512(const uint32_t* t, int avail, Walk* w) {
    constexpr int G = 32;
    int64_t i = w->i; uint32_t* cur = w->cur;
    int d = 0;
    while (d < avail && i >= 1) {
        const uint32_t mask = mask_of(i);
        const int64_t lower = (int64_t)(mask >> 1);
        const __m512i vmask = _mm512_set1_epi32((int)mask);
        for (; d + G <= avail && i - G > lower; d += G) {
            const __m512i v0 = _mm512_and_si512(_mm512_loadu_si512(t + d), vmask);
            const __m512i v1 = _mm512_and_si512(_mm512_loadu_si512(t + d + 16), vmask);
            const __m512i hi = _mm512_set1_epi32((int)i), lo = _mm512_set1_epi32((int)(i - G));
            const __mmask16 a0 = _mm512_cmple_epu32_mask(v0, lo), a1 = _mm512_cmple_epu32_mask(v1, lo);
            const __mmask16 r0 = _mm512_cmpgt_epu32_mask(v0, hi), r1 = _mm512_cmpgt_epu32_mask(v1, hi);
            if (__builtin_expect((uint32_t)((a0 | r0) & (a1 | r1)) == 0xffffu, 1)) {
                const int c0 = __builtin_popcount(a0), c1 = __builtin_popcount(a1);
                _mm512_storeu_si512(cur, _mm512_maskz_compress_epi32(a0, v0));        // whole register: the tail is overwritten later
                _mm512_storeu_si512(cur + c0, _mm512_maskz_compress_epi32(a1, v1));
                cur += c0 + c1; i -= c0 + c1;
            } else {
                walk_scalar_span(t + d, G, mask, i, cur);      // i - G > lower: still one mask for all G draws
            }
        }
        for (; d < avail && i > lower; ++d) walk_scalar_span(t + d, 1, mask, i, cur);
    }
    w->i = i; w->cur = cur;
    return d;
}

// AVX2 has no compress: permute the accepted lanes to the front with a 256-entry table of lane orders.
struct Lut8 {
    alignas(32) uint32_t idx[256][8];
    Lut8() {
        for (int m = 0; m < 256; ++m) {
            int k = 0;
            for (int b = 0; b < 8; ++b) if (m & (1 << b)) idx[m][k++] = (uint32_t)b;
            for (; k < 8; ++k) idx[m][k] = 0;
        }
    }
};
const Lut8& lut8() { static const Lut8 l; return l; }

__attribute__((target("avx2,popcnt")))
int walk_avx2(const uint32_t* t, int avail, Walk* w) {
    constexpr int G = 16;
    const Lut8& L = lut8();
    int64_t i = w->i; uint32_t* cur = w->cur;
    int d = 0;
    while (d < avail && i >= 1) {
        const uint32_t mask = mask_of(i);
        const int64_t lower = (int64_t)(mask >> 1);
        const __m256i vmask = _mm256_set1_epi32((int)mask);
        // n <= 2^31 - 1, so masked words and i are non-negative as int32: signed compares are exact
        for (; d + G <= avail && i - G > lower; d += G) {
            const __m256i v0 = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(t + d)), vmask);
            const __m256i v1 = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(t + d + 8)), vmask);
            const __m256i hi = _mm256_set1_epi32((int)i), lo = _mm256_set1_epi32((int)(i - G));
            const int n0 = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpgt_epi32(v0, lo)));    // NOT surely accepted
            const int n1 = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpgt_epi32(v1, lo)));
            const int r0 = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpgt_epi32(v0, hi)));    // surely rejected
            const int r1 = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpgt_epi32(v1, hi)));
            if (__builtin_expect(n0 == r0 && n1 == r1, 1)) {
                const int a0 = (~n0) & 0xff, a1 = (~n1) & 0xff;
                const int c0 = __builtin_popcount((unsigned)a0), c1 = __builtin_popcount((unsigned)a1);
                _mm256_storeu_si256((__m256i*)cur, _mm256_permutevar8x32_epi32(v0, _mm256_load_si256((const __m256i*)L.idx[a0])));
                _mm256_storeu_si256((__m256i*)(cur + c0), _mm256_permutevar8x32_epi32(v1, _mm256_load_si256((const __m256i*)L.idx[a1])));
                cur += c0 + c1; i -= c0 + c1;
            } else {
                walk_scalar_span(t + d, G, mask, i, cur);
            }
        }
        for (; d < avail && i > lower; ++d) walk_scalar_span(t + d, 1, mask, i, cur);
    }
    w->i = i; w->cur = cur;
    return d;
}

}  // namespace

int walk(int isa, const uint32_t* words, int avail, Walk* w) {
    if (isa == kAvx512) return walk_avx512(words, avail, w);
    if (isa == kAvx2) return walk_avx2(words, avail, w);
    return walk_scalar(words, avail, w);
}

}  // namespace tsb_hp
