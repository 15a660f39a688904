// hostrng.cu -- HOST helper: numpy's legacy normal stream, bit for bit, at about twice numpy's pace.
//
// The samplers must consume the reference's random numbers in the reference's order: legacy global
// numpy RNG = MT19937 seeded with seed + rank, per proposal randint(Lmin, Lmax + 1), randn(M),
// rand() (inversion/hmc.py:260, 297, 95, 165).  At one million voxels and 64 chains on 8 GPUs that is
// ~3e8 normals per second of host work, which the background draw threads (inversion/batched.py:
// _DrawAhead) share with everything else on the box.  This routine continues a RandomState's stream
// exactly where numpy left it (state in, state out: key[624], pos, has_gauss, cached gaussian) and
// writes randn(n) * scale:
//   MT19937 next32 with numpy's tempering; a double = (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53;
//   Marsaglia's polar method with the second deviate cached, exactly numpy's legacy_gauss
//   (x2 * f is returned first, x1 * f is kept), using libm's log like numpy does.
// No device code here; it is part of the library because the host side of the C ABI needs it.
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

// The generator state is numpy's (raw key[624] + position); outputs are tempered a whole block at a
// time into a side buffer (two simple loops the host compiler vectorises), which is about twice as
// fast as tempering word by word behind a position check.
struct Mt {
    uint32_t *key;
    int pos;              // next raw word of the current block (numpy's `pos`)
    uint32_t buf[kN];     // tempered outputs of the current block, valid for [pos0, kN)
};

inline void mt_twist(uint32_t *key) {
    int i;
    uint32_t y;
    for (i = 0; i < kN - kM; ++i) {
        y = (key[i] & kUpper) | (key[i + 1] & kLower);
        key[i] = key[i + kM] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    }
    for (; i < kN - 1; ++i) {
        y = (key[i] & kUpper) | (key[i + 1] & kLower);
        key[i] = key[i + (kM - kN)] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    }
    y = (key[kN - 1] & kUpper) | (key[0] & kLower);
    key[kN - 1] = key[kM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
}

inline void mt_temper(Mt &s, int from) {
    for (int i = from; i < kN; ++i) {
        uint32_t y = s.key[i];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        s.buf[i] = y;
    }
}

inline uint32_t mt_next(Mt &s) {
    if (s.pos == kN) {
        mt_twist(s.key);
        mt_temper(s, 0);
        s.pos = 0;
    }
    return s.buf[s.pos++];
}

inline double mt_double(Mt &s) {
    const int32_t a = (int32_t)(mt_next(s) >> 5), b = (int32_t)(mt_next(s) >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

}  // namespace

extern "C" int gi_legacy_randn_scaled(uint32_t *key624, int32_t *pos, int32_t *has_gauss, double *cached_gauss,
                                      int64_t n, double scale, double *out_host) {
    GI_REQUIRE(key624 && pos && has_gauss && cached_gauss && (out_host || n == 0) && n >= 0,
               "gi_legacy_randn_scaled: bad argument");
    GI_REQUIRE(*pos >= 0 && *pos <= kN, "gi_legacy_randn_scaled: bad generator position");
    Mt s;
    s.key = key624;
    s.pos = *pos;
    mt_temper(s, s.pos);
    int have = *has_gauss;
    double cached = *cached_gauss;
    for (int64_t i = 0; i < n; ++i) {
        double v;
        if (have) {
            v = cached;
            have = 0;
            cached = 0.0;
        } else {
            double x1, x2, r2;
            do {
                x1 = 2.0 * mt_double(s) - 1.0;
                x2 = 2.0 * mt_double(s) - 1.0;
                r2 = x1 * x1 + x2 * x2;
            } while (r2 >= 1.0 || r2 == 0.0);
            const double f = sqrt(-2.0 * log(r2) / r2);
            cached = f * x1;
            have = 1;
            v = f * x2;
        }
        out_host[i] = v * scale;
    }
    *pos = s.pos;
    *has_gauss = have;
    *cached_gauss = cached;
    return GI_OK;
}

// ---- release / acquire on the shared draw ring's control words (inversion/batched.py: _DrawRing) ----
// The ring lives in memory mapped by several processes; the payload written before a "ready" /
// "done" word must be visible to whoever reads that word.  Plain numpy stores only guarantee that on
// x86 (TSO); these two make it hold on aarch64 hosts (Grace) as well.
extern "C" void gi_ring_store_release(int64_t *word, int64_t value) {
    __atomic_store_n(word, value, __ATOMIC_RELEASE);
}
extern "C" int64_t gi_ring_load_acquire(const int64_t *word) { return __atomic_load_n(word, __ATOMIC_ACQUIRE); }
