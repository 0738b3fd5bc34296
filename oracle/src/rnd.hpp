// ORACLE (test infrastructure only -- never linked into the product path).
// Restatement of erlamsa_rnd (src/erlamsa_rnd.erl:65-242) on top of OTP stdlib
// `random` (AS183 Wichmann-Hill; not in the reference tree -- restated from the
// published OTP algorithm: seed/3 maps {A1,A2,A3} to
// {abs(A1) rem 30268 + 1, abs(A2) rem 30306 + 1, abs(A3) rem 30322 + 1};
// uniform/0 steps B1=A1*171 rem 30269, B2=A2*172 rem 30307, B3=A3*170 rem 30323
// and returns frac(B1/30269 + B2/30307 + B3/30323); uniform/1 = trunc(uniform()*N)+1).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <string>
#include <cmath>
#include <vector>
#include <algorithm>
#include <stdexcept>
#include "bigint.hpp"

namespace eo {

struct Rng {
    int64_t a1 = 3172, a2 = 9814, a3 = 20125;  // random:seed0()
    uint64_t draws = 0;

    // erlamsa_rnd:seed/1 -> random:seed/1   (src/erlamsa_rnd.erl:72-73)
    void seed(int64_t s1, int64_t s2, int64_t s3) {
        a1 = (std::llabs(s1) % 30268) + 1;
        a2 = (std::llabs(s2) % 30306) + 1;
        a3 = (std::llabs(s3) % 30322) + 1;
    }
    // random:uniform/0
    double uniform() {
        a1 = (a1 * 171) % 30269; a2 = (a2 * 172) % 30307; a3 = (a3 * 170) % 30323;
        draws++;
        double r = (double)a1 / 30269.0 + (double)a2 / 30307.0 + (double)a3 / 30323.0;
        return r - std::trunc(r);
    }
    // random:uniform/1 for bounds below 2^63
    uint64_t uniform_n(uint64_t n) { return (uint64_t)std::trunc(uniform() * (double)n) + 1; }
    // rand/1 :77-78 -- [0,N); N = 0 consumes no draw
    uint64_t rand(uint64_t n) { return n == 0 ? 0 : uniform_n(n) - 1; }
    // erand/1 :82-83 -- [1,N]; N = 0 consumes no draw
    uint64_t erand(uint64_t n) { return n == 0 ? 0 : uniform_n(n); }
    // rand/1 on arbitrary-size bounds (mutate_num case 9, rand_nbit for wide N)
    BigInt rand_big(const BigInt& n) {
        if (n.is_zero()) return BigInt();
        // random:uniform/1 multiplies a float by the bound: a bignum beyond the double range makes BEAM raise badarith
        // and the worker process of the case dies (found by running the reference's source: oracle/erlref, input "007"
        // grown to 1700 digits by line repeats, mutate_num case 9)
        double u = uniform();                      // the draw is consumed before the multiplication fails
        double nd = n.to_double_erl_abs();
        if (std::isinf(nd)) throw std::overflow_error("badarith: bignum bound does not fit a float");
        double p = u * nd;
        return BigInt::from_double_trunc(std::trunc(p));
    }
    // rand_range/2 :87-92
    int64_t rand_range(int64_t l, int64_t r) {
        if (r > l) return (int64_t)rand((uint64_t)(r - l)) + l;
        if (l == r) return l;
        return 0;
    }
    double rand_float() { return uniform(); }                          // :101
    int rand_bit() { return uniform() >= 0.5 ? 1 : 0; }                // :105 round/1: half rounds up
    // rand_occurs_fixed/2 :121-130 (note the Nom == 1 quirk)
    bool rand_occurs_fixed(uint64_t nom, uint64_t denom) {
        uint64_t n = rand(denom);
        return nom == 1 ? n != 0 : n < nom;
    }
    // rand_nbit/1 :134-137
    BigInt rand_nbit(uint64_t n) {
        if (n == 0) return BigInt();
        BigInt hi = BigInt::pow2((unsigned)(n - 1));
        return hi + rand_big(hi);  // Hi bor rand(Hi): rand(Hi) < Hi = 2^(n-1), so bor == +
    }
    // rand_log/1 :141-143
    BigInt rand_log(uint64_t n) { return n == 0 ? BigInt() : rand_nbit(rand(n)); }
    uint64_t rand_log_u64(uint64_t n) { return rand_log(n).to_u64(); }  // n <= 64
    // rand_elem/1 :148-151 -> 0-based index, or -1 for the empty list (no draw)
    int64_t rand_elem_idx(uint64_t len) { return len == 0 ? -1 : (int64_t)uniform_n(len) - 1; }
    // rand_delta/0 :224-231
    int rand_delta() { return rand_bit() == 0 ? +1 : -1; }
    // random_block/1 :165,173-174 -- N draws, each prepended
    std::string random_block(uint64_t n) {
        std::string s(n, '\0');
        for (uint64_t i = 0; i < n; i++) s[n - 1 - i] = (char)rand(256);
        return s;
    }
    // random_numbers/2 :178-183
    std::string random_numbers_256(uint64_t cnt) { return random_block(cnt); }
    // random_permutation/1 :190-196
    template <class T, class Less>
    std::vector<T> random_permutation(const std::vector<T>& l, Less less) {
        if (l.size() == 2) {
            if (rand(2) == 1) return std::vector<T>{l[1], l[0]};
            return l;
        }
        std::vector<std::pair<double, size_t>> keyed;
        for (size_t i = 0; i < l.size(); i++) { double k = uniform(); keyed.push_back({k, i}); }
        // lists:sort/1 on {Float, Elem}: ascending term order, merge sort (stable)
        std::stable_sort(keyed.begin(), keyed.end(), [&](const std::pair<double, size_t>& x, const std::pair<double, size_t>& y) {
            if (x.first != y.first) return x.first < y.first;
            return less(l[x.second], l[y.second]);
        });
        std::vector<T> out;
        for (auto& k : keyed) out.push_back(l[k.second]);
        return out;
    }
    // reservoir_sample/2 :201-214 -> indices into the list
    std::vector<size_t> reservoir_sample_idx(size_t n, size_t k) {
        std::vector<size_t> r;
        if (k >= n) { for (size_t i = 0; i < n; i++) r.push_back(i); return r; }
        for (size_t i = 0; i < k; i++) r.push_back(i);
        for (size_t i = k + 1; i <= n; i++) {
            uint64_t j = erand(i);
            if (j <= k) r[j - 1] = i - 1;
        }
        return r;
    }
    // gen_predictable_seed/0 :65
    void gen_predictable_seed(int64_t out[3]) { out[0] = erand(99999); out[1] = erand(99999); out[2] = erand(99999); }
};

}  // namespace eo
