// erlamsa_b200 -- fixed-width (256-bit magnitude + sign) integers for the textual-number mutator.
// The reference works on unbounded Erlang integers (sed_num, reference
// src/erlamsa_mutations.erl:92-169); the device supports numbers of up to 77 decimal digits
// (|x| < 2^255) and flags anything wider (EB200_CASE_UNSUPPORTED) instead of truncating silently.
#pragma once
#include <stdint.h>
#include <math.h>

namespace eb {

struct Big256 {
    uint32_t m[8];
    uint32_t neg;       // 0/1; zero is never negative
    uint32_t ovf;       // sticky: a result did not fit

    __host__ __device__ void zero() { for (int i = 0; i < 8; i++) m[i] = 0; neg = 0; ovf = 0; }
    __host__ __device__ bool is_zero() const { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= m[i]; return o == 0; }
    __host__ __device__ void set_u64(uint64_t v) { zero(); m[0] = (uint32_t)v; m[1] = (uint32_t)(v >> 32); }
    __host__ __device__ void set_pow2(unsigned k) { zero(); if (k >= 256) { ovf = 1; return; } m[k >> 5] = 1u << (k & 31); }
    __host__ __device__ int top_limb() const { for (int i = 7; i >= 0; i--) if (m[i]) return i; return -1; }
    __host__ __device__ static int cmp_abs(const Big256& a, const Big256& b) {
        for (int i = 7; i >= 0; i--) if (a.m[i] != b.m[i]) return a.m[i] < b.m[i] ? -1 : 1;
        return 0;
    }
    __host__ __device__ void add_abs(const Big256& o) {   // |this| += |o|
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) { uint64_t s = (uint64_t)m[i] + o.m[i] + c; m[i] = (uint32_t)s; c = s >> 32; }
        if (c) ovf = 1;
        ovf |= o.ovf;
    }
    __host__ __device__ void sub_abs(const Big256& o) {   // |this| -= |o|, requires |this| >= |o|
        int64_t br = 0;
        for (int i = 0; i < 8; i++) { int64_t d = (int64_t)m[i] - o.m[i] - br; if (d < 0) { d += 4294967296ll; br = 1; } else br = 0; m[i] = (uint32_t)d; }
        ovf |= o.ovf;
    }
    // this = this + o (signed)
    __host__ __device__ void add(const Big256& o) {
        if (neg == o.neg) { add_abs(o); }
        else {
            int c = cmp_abs(*this, o);
            if (c == 0) { uint32_t f = ovf | o.ovf; zero(); ovf = f; }
            else if (c > 0) sub_abs(o);
            else { Big256 t = o; t.sub_abs(*this); t.ovf |= ovf; *this = t; }
        }
        if (is_zero()) neg = 0;
    }
    __host__ __device__ void negate() { if (!is_zero()) neg ^= 1u; }
    __host__ __device__ void sub(const Big256& o) { Big256 t = o; t.negate(); add(t); }
    __host__ __device__ void mul_small(uint32_t k) {
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) { uint64_t p = (uint64_t)m[i] * k + c; m[i] = (uint32_t)p; c = p >> 32; }
        if (c) ovf = 1;
        if (is_zero()) neg = 0;
    }
    __host__ __device__ void add_small(uint32_t k) {   // magnitude += k
        uint64_t c = k;
        for (int i = 0; i < 8 && c; i++) { uint64_t s = (uint64_t)m[i] + c; m[i] = (uint32_t)s; c = s >> 32; }
        if (c) ovf = 1;
    }
    // divide magnitude by 10^9, return remainder
    __host__ __device__ uint32_t divmod_1e9() {
        uint64_t rem = 0;
        for (int i = 7; i >= 0; i--) { uint64_t cur = (rem << 32) | m[i]; m[i] = (uint32_t)(cur / 1000000000ull); rem = cur % 1000000000ull; }
        return (uint32_t)rem;
    }
    // integer_to_list/1 into buf (at most 80 bytes); returns length
    __host__ __device__ int to_decimal(uint8_t* buf) const {
        if (is_zero()) { buf[0] = '0'; return 1; }
        Big256 t = *this; uint32_t groups[9]; int ng = 0;
        while (!t.is_zero()) groups[ng++] = t.divmod_1e9();
        int n = 0;
        if (neg) buf[n++] = '-';
        // most significant group without padding
        { uint32_t g = groups[ng - 1]; uint8_t tmp[10]; int k = 0; do { tmp[k++] = (uint8_t)('0' + g % 10); g /= 10; } while (g); while (k) buf[n++] = tmp[--k]; }
        for (int i = ng - 2; i >= 0; i--) { uint32_t g = groups[i]; for (int k = 8; k >= 0; k--) { buf[n + k] = (uint8_t)('0' + g % 10); g /= 10; } n += 9; }
        return n;
    }
    // BEAM integer->float of the magnitude: one 64-bit digit converts like a C cast; wider values
    // accumulate d = d * 2^64 + digit from the most significant 64-bit digit (erts big_to_double).
    __host__ __device__ double to_double_erl_abs() const {
        int tl = top_limb();
        if (tl < 2) { uint64_t v = ((uint64_t)m[1] << 32) | m[0]; return (double)v; }
        int n64 = tl / 2 + 1; double d = 0.0;
        for (int i = n64 - 1; i >= 0; i--) { uint64_t digit = ((uint64_t)m[2 * i + 1] << 32) | m[2 * i]; d = d * 18446744073709551616.0 + (double)digit; }
        return d;
    }
    // trunc/1 of a finite non-negative double (already truncated)
    __host__ __device__ void from_double_trunc(double x) {
        zero();
        if (!(x >= 1.0)) return;
        uint64_t bits;
#ifdef __CUDA_ARCH__
        bits = (uint64_t)__double_as_longlong(x);
#else
        union { double d; uint64_t u; } cv; cv.d = x; bits = cv.u;
#endif
        int e = (int)((bits >> 52) & 0x7ff) - 1075;           // x = mant * 2^e
        uint64_t mant = (bits & 0xfffffffffffffull) | (1ull << 52);
        if (e <= 0) { uint64_t v = (-e >= 64) ? 0 : (mant >> (-e)); m[0] = (uint32_t)v; m[1] = (uint32_t)(v >> 32); return; }
        if (e + 53 > 256) { ovf = 1; return; }
        int limb = e >> 5, sh = e & 31;
        // place the 53-bit mantissa shifted left by e bits
        uint64_t lo = mant << sh;                      // low 64 bits of (mant << sh)
        uint64_t hi = sh ? (mant >> (64 - sh)) : 0;    // overflow bits
        if (limb < 8) m[limb] = (uint32_t)lo;
        if (limb + 1 < 8) m[limb + 1] = (uint32_t)(lo >> 32);
        if (limb + 2 < 8) m[limb + 2] = (uint32_t)hi;
        if (limb + 3 < 8) m[limb + 3] = (uint32_t)(hi >> 32);
    }
};

}  // namespace eb
