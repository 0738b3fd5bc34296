// ORACLE (test infrastructure only -- never linked into the product path).
// Arbitrary-precision signed integers with the handful of operations Erlang
// integers need on erlamsa's hot path (sed_num / mutate_num,
// src/erlamsa_mutations.erl:92-169; rand_nbit/rand_log, src/erlamsa_rnd.erl:134-143).
//
// Integer -> float conversion follows the BEAM: small integers (60-bit) convert
// exactly like a C cast; bignums are accumulated most-significant 64-bit digit
// first as d = d * 2^64 + digit (erts big_to_double), which is NOT always the
// correctly rounded value -- restated here because random:uniform/1 multiplies
// a float by the integer bound.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>

namespace eo {

struct BigInt {
    bool neg = false;
    std::vector<uint32_t> mag;  // little-endian limbs, no leading zero limbs; zero == empty

    BigInt() {}
    BigInt(int64_t v) {
        uint64_t u;
        if (v < 0) { neg = true; u = (uint64_t)(-(v + 1)) + 1; } else u = (uint64_t)v;
        set_u64(u);
    }
    static BigInt from_u64(uint64_t u) { BigInt b; b.set_u64(u); return b; }
    void set_u64(uint64_t u) {
        mag.clear();
        while (u) { mag.push_back((uint32_t)u); u >>= 32; }
    }
    bool is_zero() const { return mag.empty(); }
    void norm() { while (!mag.empty() && mag.back() == 0) mag.pop_back(); if (mag.empty()) neg = false; }

    static int cmp_abs(const BigInt& a, const BigInt& b) {
        if (a.mag.size() != b.mag.size()) return a.mag.size() < b.mag.size() ? -1 : 1;
        for (size_t i = a.mag.size(); i-- > 0;) if (a.mag[i] != b.mag[i]) return a.mag[i] < b.mag[i] ? -1 : 1;
        return 0;
    }
    static std::vector<uint32_t> add_abs(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
        std::vector<uint32_t> r; uint64_t c = 0; size_t n = std::max(a.size(), b.size());
        for (size_t i = 0; i < n; i++) {
            uint64_t s = c + (i < a.size() ? a[i] : 0) + (i < b.size() ? b[i] : 0);
            r.push_back((uint32_t)s); c = s >> 32;
        }
        if (c) r.push_back((uint32_t)c);
        return r;
    }
    // |a| >= |b|
    static std::vector<uint32_t> sub_abs(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
        std::vector<uint32_t> r; int64_t br = 0;
        for (size_t i = 0; i < a.size(); i++) {
            int64_t d = (int64_t)a[i] - br - (i < b.size() ? b[i] : 0);
            if (d < 0) { d += ((int64_t)1 << 32); br = 1; } else br = 0;
            r.push_back((uint32_t)d);
        }
        return r;
    }
    BigInt operator-() const { BigInt r = *this; if (!r.is_zero()) r.neg = !r.neg; return r; }
    BigInt abs() const { BigInt r = *this; r.neg = false; return r; }
    BigInt operator+(const BigInt& o) const {
        BigInt r;
        if (neg == o.neg) { r.mag = add_abs(mag, o.mag); r.neg = neg; }
        else {
            int c = cmp_abs(*this, o);
            if (c == 0) return BigInt();
            if (c > 0) { r.mag = sub_abs(mag, o.mag); r.neg = neg; }
            else { r.mag = sub_abs(o.mag, mag); r.neg = o.neg; }
        }
        r.norm(); return r;
    }
    BigInt operator-(const BigInt& o) const { return *this + (-o); }
    BigInt mul_small(uint32_t m) const {
        BigInt r; r.neg = neg; uint64_t c = 0;
        for (size_t i = 0; i < mag.size(); i++) { uint64_t p = (uint64_t)mag[i] * m + c; r.mag.push_back((uint32_t)p); c = p >> 32; }
        if (c) r.mag.push_back((uint32_t)c);
        r.norm(); return r;
    }
    // this * s where s is +1/-1
    BigInt mul_sign(int s) const { return s < 0 ? -*this : *this; }
    static BigInt pow2(unsigned k) { BigInt r; r.mag.assign(k / 32 + 1, 0); r.mag[k / 32] = 1u << (k % 32); return r; }
    bool operator==(const BigInt& o) const { return neg == o.neg && mag == o.mag; }
    bool fits_u64() const { return !neg && mag.size() <= 2; }
    uint64_t to_u64() const { uint64_t v = 0; for (size_t i = mag.size(); i-- > 0;) v = (v << 32) | mag[i]; return v; }

    static BigInt from_decimal(const std::string& digits, bool negative) {
        BigInt r;
        for (char ch : digits) { r = r.mul_small(10); r = r + BigInt((int64_t)(ch - '0')); }
        if (negative && !r.is_zero()) r.neg = true;
        return r;
    }
    // integer_to_list/1
    std::string to_string() const {
        if (is_zero()) return "0";
        std::vector<uint32_t> t = mag; std::vector<uint32_t> groups;
        while (!t.empty()) {
            uint64_t rem = 0;
            for (size_t i = t.size(); i-- > 0;) { uint64_t cur = (rem << 32) | t[i]; t[i] = (uint32_t)(cur / 1000000000u); rem = cur % 1000000000u; }
            while (!t.empty() && t.back() == 0) t.pop_back();
            groups.push_back((uint32_t)rem);
        }
        std::string s = neg ? "-" : "";
        s += std::to_string(groups.back());
        for (size_t i = groups.size() - 1; i-- > 0;) { std::string g = std::to_string(groups[i]); s += std::string(9 - g.size(), '0') + g; }
        return s;
    }
    // BEAM integer -> float (see header comment). Magnitude only; caller applies sign.
    double to_double_erl_abs() const {
        size_t n64 = (mag.size() + 1) / 2;
        if (n64 <= 1) {
            uint64_t v = to_u64();
            // smalls (< 2^59) and one-digit bignums both reduce to a correctly rounded u64->double
            return (double)v;
        }
        double d = 0.0; const double dbase = 18446744073709551616.0;  // 2^64
        for (size_t i = n64; i-- > 0;) {
            uint64_t lo = mag[2 * i]; uint64_t hi = (2 * i + 1 < mag.size()) ? mag[2 * i + 1] : 0;
            uint64_t digit = (hi << 32) | lo;
            d = d * dbase + (double)digit;
        }
        return d;
    }
    // trunc/1 of a non-negative finite double
    static BigInt from_double_trunc(double x) {
        BigInt r; if (!(x >= 1.0)) return r;
        int e; double m = std::frexp(x, &e);           // x = m * 2^e, m in [0.5,1)
        uint64_t mant = (uint64_t)std::ldexp(m, 53);   // 53-bit integer mantissa
        int sh = e - 53;
        if (sh <= 0) { r.set_u64(mant >> (-sh)); return r; }
        r.set_u64(mant);
        // shift left by sh bits
        unsigned limbs = sh / 32, bits = sh % 32;
        std::vector<uint32_t> m2(limbs, 0); uint64_t c = 0;
        for (uint32_t w : r.mag) { uint64_t v = ((uint64_t)w << bits) | c; m2.push_back((uint32_t)v); c = v >> 32; }
        if (c) m2.push_back((uint32_t)c);
        r.mag = m2; r.norm(); return r;
    }
};

}  // namespace eo
