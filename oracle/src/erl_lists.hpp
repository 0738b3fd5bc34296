// ORACLE (test infrastructure only -- never linked into the product path).
// Restatement of OTP stdlib lists:sort/2 (not in the reference tree; part of the
// OTP runtime the reference runs on). erlamsa calls it with a comparator that
// is NOT a total "less-or-equal" (`A > B`, src/erlamsa_utils.erl:114-117), so
// the order of equal-priority entries is decided by the exact merge-sort
// OTP implements; with a proper `=<`-style comparator (src/erlamsa_mutations.erl:1249,
// `A >= B`) it is a stable sort, which tests/test_oracle_units.py checks.
//
// The algorithm: split the input into alternating ascending/descending runs
// (fsplit_1 / fsplit_2 with one look-aside element), then merge runs pairwise,
// alternating between "merge into reversed" (rfmergel) and "merge back"
// (fmergel) passes until one run is left.
#pragma once
#include <functional>
#include <memory>
#include <vector>

namespace eo {

template <class T>
struct ErlSort {
    using Fun = std::function<bool(const T&, const T&)>;
    struct Cell; using L = std::shared_ptr<Cell>;
    struct Cell { T h; L t; };
    struct RCell; using LL = std::shared_ptr<RCell>;   // list of lists
    struct RCell { L h; LL t; };

    static L cons(const T& h, L t) { return std::make_shared<Cell>(Cell{h, t}); }
    static LL lcons(L h, LL t) { return std::make_shared<RCell>(RCell{h, t}); }
    static L rev(L l, L acc) { while (l) { acc = cons(l->h, acc); l = l->t; } return acc; }

    Fun f;
    explicit ErlSort(Fun fn) : f(fn) {}

    std::vector<T> sort(const std::vector<T>& in) {
        if (in.size() < 2) return in;
        L l = nullptr; for (size_t i = in.size(); i-- > 0;) l = cons(in[i], l);
        const T& x = l->h; const T& y = l->t->h; L t = l->t->t;
        L r = f(x, y) ? fsplit_1(y, x, t, nullptr, nullptr) : fsplit_2(y, x, t, nullptr, nullptr);
        std::vector<T> out; while (r) { out.push_back(r->h); r = r->t; } return out;
    }

    // ---- run detection, ascending flavour
    L fsplit_1(T y, T x, L l, L r, LL rs) {
        for (;;) {
            if (!l) return rfmergel(lcons(cons(y, cons(x, r)), rs), nullptr, true);
            T z = l->h; l = l->t;
            if (f(y, z)) { r = cons(x, r); x = y; y = z; }
            else if (f(x, z)) { r = cons(x, r); x = z; }
            else if (!r) { r = cons(z, nullptr); }
            else return fsplit_1_1(y, x, l, r, rs, z);
        }
    }
    L fsplit_1_1(T y, T x, L l, L r, LL rs, T s) {
        for (;;) {
            if (!l) return rfmergel(lcons(cons(s, nullptr), lcons(cons(y, cons(x, r)), rs)), nullptr, true);
            T z = l->h; l = l->t;
            if (f(y, z)) { r = cons(x, r); x = y; y = z; }
            else if (f(x, z)) { r = cons(x, r); x = z; }
            else if (f(s, z)) return fsplit_1(z, s, l, nullptr, lcons(cons(y, cons(x, r)), rs));
            else return fsplit_1(s, z, l, nullptr, lcons(cons(y, cons(x, r)), rs));
        }
    }
    // ---- run detection, descending flavour
    L fsplit_2(T y, T x, L l, L r, LL rs) {
        for (;;) {
            if (!l) return fmergel(lcons(cons(y, cons(x, r)), rs), nullptr, false);
            T z = l->h; l = l->t;
            if (!f(y, z)) { r = cons(x, r); x = y; y = z; }
            else if (!f(x, z)) { r = cons(x, r); x = z; }
            else if (!r) { r = cons(z, nullptr); }
            else return fsplit_2_1(y, x, l, r, rs, z);
        }
    }
    L fsplit_2_1(T y, T x, L l, L r, LL rs, T s) {
        for (;;) {
            if (!l) return fmergel(lcons(cons(s, nullptr), lcons(cons(y, cons(x, r)), rs)), nullptr, false);
            T z = l->h; l = l->t;
            if (!f(y, z)) { r = cons(x, r); x = y; y = z; }
            else if (!f(x, z)) { r = cons(x, r); x = z; }
            else if (!f(s, z)) return fsplit_2(z, s, l, nullptr, lcons(cons(y, cons(x, r)), rs));
            else return fsplit_2(s, z, l, nullptr, lcons(cons(y, cons(x, r)), rs));
        }
    }
    // ---- merge passes. asc == true is the `asc` tag of the OTP code.
    L fmergel(LL ls, LL acc, bool asc) {
        for (;;) {
            if (ls && ls->t) {
                L a = ls->h, b = ls->t->h; LL rest = ls->t->t;
                L m = asc ? fmerge2_1(a, b->h, b->t, nullptr)      // [T1, [H2|T2] | L]
                          : fmerge2_1(b, a->h, a->t, nullptr);     // [[H2|T2], T1 | L]
                acc = lcons(m, acc); ls = rest; continue;
            }
            if (ls && !ls->t) {
                if (!acc) return ls->h;
                return rfmergel(lcons(rev(ls->h, nullptr), acc), nullptr, asc);
            }
            return rfmergel(acc, nullptr, asc);
        }
    }
    L rfmergel(LL ls, LL acc, bool asc) {
        for (;;) {
            if (ls && ls->t) {
                L a = ls->h, b = ls->t->h; LL rest = ls->t->t;
                L m = asc ? rfmerge2_1(b, a->h, a->t, nullptr)     // [[H2|T2], T1 | L]
                          : rfmerge2_1(a, b->h, b->t, nullptr);    // [T1, [H2|T2] | L]
                acc = lcons(m, acc); ls = rest; continue;
            }
            if (ls && !ls->t) return fmergel(lcons(rev(ls->h, nullptr), acc), nullptr, asc);
            return fmergel(acc, nullptr, asc);
        }
    }
    // Elements from the first list are prioritized.
    L fmerge2_1(L t1, T h2, L t2, L m) {
        for (;;) {
            // state A: have list t1 (head not yet taken) and pending h2
            if (!t1) return rev(t2, cons(h2, m));
            T h1 = t1->h; L tt1 = t1->t;
            if (f(h1, h2)) { m = cons(h1, m); t1 = tt1; continue; }
            m = cons(h2, m);
            // state B: pending h1 (with tt1), walk t2
            for (;;) {
                if (!t2) return rev(tt1, cons(h1, m));
                T nh2 = t2->h; L nt2 = t2->t;
                if (f(h1, nh2)) { m = cons(h1, m); t1 = tt1; h2 = nh2; t2 = nt2; break; }
                m = cons(nh2, m); t2 = nt2;
            }
        }
    }
    L rfmerge2_1(L t1, T h2, L t2, L m) {
        for (;;) {
            if (!t1) return rev(t2, cons(h2, m));
            T h1 = t1->h; L tt1 = t1->t;
            if (!f(h1, h2)) { m = cons(h1, m); t1 = tt1; continue; }
            m = cons(h2, m);
            for (;;) {
                if (!t2) return rev(tt1, cons(h1, m));
                T nh2 = t2->h; L nt2 = t2->t;
                if (!f(h1, nh2)) { m = cons(h1, m); t1 = tt1; h2 = nh2; t2 = nt2; break; }
                m = cons(nh2, m); t2 = nt2;
            }
        }
    }
};

}  // namespace eo
