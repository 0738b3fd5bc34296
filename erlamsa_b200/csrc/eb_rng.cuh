// erlamsa_b200 -- erlamsa_rnd on the device.
//
// Two generators behind one interface (eb200_opts.rng_mode):
//  * AS183 (OTP stdlib `random`, which reference src/erlamsa_rnd.erl:73-105 wraps): three 16-bit
//    multiplicative LCGs, uniform() = frac(B1/30269 + B2/30307 + B3/30323) in IEEE double -- the
//    draw-for-draw "exact" mode. Each component is x_k = x_0 * a^k mod p, so the stream can be
//    entered at any draw index with three modular exponentiations (used to derive per-case seeds
//    without replaying the parent stream, reference src/erlamsa_main.erl:179).
//  * Philox4x32-10 keyed by the option seed; counter = (index inside the slot, SLOT, global case id), where a slot is
//    (mutation round, who draws: scheduler / mutator id / pattern): every decision has its own counter range, so a mutator
//    that fails or draws more does not shift anybody else's draws, and per-byte draws (sp / snand / srnd) are computed by
//    32 lanes at once (uniform_philox_at). Distribution-equivalent, not bit-equal, to the reference stream.
//
// All helpers mirror erlamsa_rnd's API one to one (names and N = 0 "no draw" rules included),
// reference src/erlamsa_rnd.erl:65-242.
#pragma once
#include <stdint.h>
#include <math.h>
#include "eb_common.cuh"
#include "eb_bigint.cuh"

namespace eb {

#define EB_HD __host__ __device__ __forceinline__

constexpr int32_t AS_P1 = 30269, AS_P2 = 30307, AS_P3 = 30323;
constexpr int32_t AS_M1 = 171, AS_M2 = 172, AS_M3 = 170;

// a^k mod P for the three 15-bit AS183 moduli: products stay below 2^30, and the modulus is a
// compile-time constant so `%` lowers to a multiply-high instead of a software division
template <uint32_t P>
EB_HD uint32_t modpow_u32(uint32_t a, uint64_t k) {
    uint32_t r = 1, b = a % P;
    while (k) { if (k & 1) r = (r * b) % P; b = (b * b) % P; k >>= 1; }
    return r;
}

EB_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
    int32_t a1, a2, a3;      // AS183 state
    uint64_t draws;          // draw index (also the Philox counter)
    int32_t mode;            // 0 AS183, 1 Philox
    uint64_t key, ctr_hi;    // Philox: key = option seed hash, ctr_hi = global case id
    uint32_t slot, local;    // Philox: current slot and the next index inside it
    // Philox only: start drawing from slot s (no-op for AS183, whose stream is one sequence by definition)
    // (re-entering the slot that is already current continues its sequence: a loop that comes back to the same slot -- `bu` over
    //  an emptied block list -- must not see the same draw again and again)
    EB_HD void set_slot(uint32_t s) { if (mode != 0 && slot != s) { slot = s; local = 0; } }
    // Philox only: draw number `idx` of the current slot, without advancing anything (lane-parallel loops)
    EB_HD double uniform_philox_at(uint32_t idx) const {
        uint32_t o[4];
        philox4x32_10(idx, slot, (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32), (uint32_t)key, (uint32_t)(key >> 32), o);
        uint64_t v = ((uint64_t)o[0] << 32) | o[1];
        return (double)(v >> 11) * (1.0 / 9007199254740992.0);
    }
    EB_HD void philox_skip(uint32_t n) { local += n; draws += n; }

    // erlamsa_rnd:seed/1 -> random:seed/3
    EB_HD void seed(int64_t s1, int64_t s2, int64_t s3) {
        a1 = (int32_t)((s1 < 0 ? -s1 : s1) % 30268) + 1;
        a2 = (int32_t)((s2 < 0 ? -s2 : s2) % 30306) + 1;
        a3 = (int32_t)((s3 < 0 ? -s3 : s3) % 30322) + 1;
        draws = 0; slot = 0; local = 0;
    }
    // skip k draws of the AS183 stream in O(log k)
    EB_HD void jump(uint64_t k) {
        a1 = (int32_t)(((uint32_t)a1 * modpow_u32<30269>(AS_M1, k)) % 30269u);
        a2 = (int32_t)(((uint32_t)a2 * modpow_u32<30307>(AS_M2, k)) % 30307u);
        a3 = (int32_t)(((uint32_t)a3 * modpow_u32<30323>(AS_M3, k)) % 30323u);
        draws += k;
    }
    // random:uniform/0
    // IEEE-correct a/c for the integers a < c that AS183 produces, without the ~30-instruction
    // software division: with y = RN(1/c), q0 = a*y, r = fma(-q0, c, a), q = fma(r, y, q0) is the
    // correctly rounded quotient for EVERY a in [0, c) and c in {30269, 30307, 30323}
    // (checked exhaustively, tests/test_as183_division.py), so results stay bit-identical to `a / c`.
    static EB_HD double div_exact(double a, double c, double y) {
#ifdef __CUDA_ARCH__
        double q0 = __dmul_rn(a, y);
        double r = __fma_rn(-q0, c, a);
        return __fma_rn(r, y, q0);
#else
        (void)y; return a / c;
#endif
    }
    EB_HD double uniform_philox() {
        uint32_t o[4];
        philox4x32_10(local, slot, (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32), (uint32_t)key, (uint32_t)(key >> 32), o);
        draws++; local++;
        uint64_t v = ((uint64_t)o[0] << 32) | o[1];
        return (double)(v >> 11) * (1.0 / 9007199254740992.0);
    }
    EB_HD double uniform_as183() {
        a1 = (a1 * AS_M1) % AS_P1; a2 = (a2 * AS_M2) % AS_P2; a3 = (a3 * AS_M3) % AS_P3;
        draws++;
        double r = div_exact((double)a1, 30269.0, 1.0 / 30269.0) + div_exact((double)a2, 30307.0, 1.0 / 30307.0) + div_exact((double)a3, 30323.0, 1.0 / 30323.0);
        return r - trunc(r);
    }
#ifdef __CUDA_ARCH__
    // out of line: one copy of each generator in the kernel (the pipeline has ~100 call sites)
    __device__ __noinline__ double uniform_philox_ool() { return uniform_philox(); }
    __device__ __noinline__ double uniform() { return mode == 0 ? uniform_as183() : uniform_philox_ool(); }
#else
    double uniform() { return mode == 0 ? uniform_as183() : uniform_philox(); }
#endif
    // random:uniform/1 (N < 2^63)
    EB_HD uint64_t uniform_n(uint64_t n) { return (uint64_t)trunc(uniform() * (double)n) + 1; }
    EB_HD uint64_t rand(uint64_t n) { return n == 0 ? 0 : uniform_n(n) - 1; }     // :77-78
    EB_HD uint64_t erand(uint64_t n) { return n == 0 ? 0 : uniform_n(n); }        // :82-83
    EB_HD int64_t rand_range(int64_t l, int64_t r) {                              // :87-92
        if (r > l) return (int64_t)rand((uint64_t)(r - l)) + l;
        if (l == r) return l;
        return 0;
    }
    EB_HD int rand_bit() { return uniform() >= 0.5 ? 1 : 0; }                     // :105
    EB_HD int rand_delta() { return rand_bit() == 0 ? 1 : -1; }                   // :224-231
    EB_HD bool rand_occurs_fixed(uint64_t nom, uint64_t denom) {                  // :121-130
        uint64_t n = rand(denom);
        return nom == 1 ? n != 0 : n < nom;
    }
    // rand_log/1 for N <= 63 :134-143
    EB_HD uint64_t rand_log_small(uint64_t n) {
        if (n == 0) return 0;
        uint64_t k = rand(n);
        if (k == 0) return 0;
        uint64_t hi = 1ull << (k - 1);
        return hi | rand(hi);
    }
    EB_HD int64_t rand_elem_idx(uint64_t len) { return len == 0 ? -1 : (int64_t)uniform_n(len) - 1; }   // :148-151
    // rand/1 with a wide bound (mutate_num), BEAM integer->float conversion included
    EB_HD Big256 rand_big(const Big256& n) {
        Big256 r; r.zero();
        if (n.is_zero()) return r;
        double p = uniform() * n.to_double_erl_abs();
        r.from_double_trunc(trunc(p));
        return r;
    }
    // rand_log/1 for N <= 129
    EB_HD Big256 rand_log_big(uint64_t n) {
        Big256 r; r.zero();
        if (n == 0) return r;
        uint64_t k = rand(n);
        if (k == 0) return r;
        Big256 hi; hi.set_pow2((unsigned)(k - 1));
        Big256 lo = rand_big(hi);
        hi.add_abs(lo);
        return hi;
    }
};

}  // namespace eb
