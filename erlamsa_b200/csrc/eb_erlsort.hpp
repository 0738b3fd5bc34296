// erlamsa_b200 (host side) -- OTP lists:sort/2 as the reference observes it.
//
// erlamsa orders its pattern and generator tables with lists:sort(fun({A,_},{B,_}) -> A > B end, L)
// (reference src/erlamsa_utils.erl:114-117). `A > B` is not a valid "less or equal" ordering
// function, so where equal priorities land is decided by OTP's concrete merge sort; the pattern a
// case gets for a given random draw depends on it. The routine below
// follows the OTP stdlib algorithm (alternating ascending/descending run split with one spare
// element, then pairwise merges that flip direction every pass) on plain std::vector "lists"
// whose element 0 is the head.
#pragma once
#include <functional>
#include <vector>

namespace eb {

template <class T>
class ErlangListSort {
  public:
    using Le = std::function<bool(const T&, const T&)>;
    explicit ErlangListSort(Le f) : fun_(std::move(f)) {}

    std::vector<T> operator()(std::vector<T> l) {
        if (l.size() < 2) return l;
        T x = l[0], y = l[1];
        std::vector<T> t(l.begin() + 2, l.end());
        return fun_(x, y) ? split_asc(y, x, t, 0, {}, {}) : split_desc(y, x, t, 0, {}, {});
    }

  private:
    using List = std::vector<T>;
    using Runs = std::vector<List>;   // element 0 is the head of the list of runs
    Le fun_;

    static List cons(const T& h, const List& t) { List r; r.reserve(t.size() + 1); r.push_back(h); r.insert(r.end(), t.begin(), t.end()); return r; }
    static Runs rcons(const List& h, const Runs& t) { Runs r; r.push_back(h); r.insert(r.end(), t.begin(), t.end()); return r; }
    static List rev_onto(const List& l, List acc) { for (const T& e : l) acc.insert(acc.begin(), e); return acc; }
    static List run3(const T& y, const T& x, const List& r) { return cons(y, cons(x, r)); }

    // fsplit_1 / fsplit_1_1 (have_s tells which)
    List split_asc(T y, T x, const List& l, size_t i, List r, Runs rs) {
        bool have_s = false; T s = y;
        for (; i < l.size(); i++) {
            const T& z = l[i];
            if (fun_(y, z)) { r = cons(x, r); x = y; y = z; continue; }
            if (fun_(x, z)) { r = cons(x, r); x = z; continue; }
            if (!have_s) {
                if (r.empty()) { r = List{z}; continue; }
                have_s = true; s = z; continue;
            }
            Runs nrs = rcons(run3(y, x, r), rs);
            if (fun_(s, z)) return split_asc(z, s, l, i + 1, {}, nrs);
            return split_asc(s, z, l, i + 1, {}, nrs);
        }
        Runs all = have_s ? rcons(List{s}, rcons(run3(y, x, r), rs)) : rcons(run3(y, x, r), rs);
        return rmergel(all, {}, true);
    }
    // fsplit_2 / fsplit_2_1
    List split_desc(T y, T x, const List& l, size_t i, List r, Runs rs) {
        bool have_s = false; T s = y;
        for (; i < l.size(); i++) {
            const T& z = l[i];
            if (!fun_(y, z)) { r = cons(x, r); x = y; y = z; continue; }
            if (!fun_(x, z)) { r = cons(x, r); x = z; continue; }
            if (!have_s) {
                if (r.empty()) { r = List{z}; continue; }
                have_s = true; s = z; continue;
            }
            Runs nrs = rcons(run3(y, x, r), rs);
            if (!fun_(s, z)) return split_desc(z, s, l, i + 1, {}, nrs);
            return split_desc(s, z, l, i + 1, {}, nrs);
        }
        Runs all = have_s ? rcons(List{s}, rcons(run3(y, x, r), rs)) : rcons(run3(y, x, r), rs);
        return mergel(all, {}, false);
    }
    // fmergel
    List mergel(Runs ls, Runs acc, bool asc) {
        while (ls.size() >= 2) {
            List m = asc ? merge_fwd(ls[0], ls[1]) : merge_fwd(ls[1], ls[0]);
            acc = rcons(m, acc); ls.erase(ls.begin(), ls.begin() + 2);
        }
        if (ls.size() == 1) {
            if (acc.empty()) return ls[0];
            return rmergel(rcons(rev_onto(ls[0], {}), acc), {}, asc);
        }
        return rmergel(acc, {}, asc);
    }
    // rfmergel
    List rmergel(Runs ls, Runs acc, bool asc) {
        while (ls.size() >= 2) {
            List m = asc ? merge_rev(ls[1], ls[0]) : merge_rev(ls[0], ls[1]);
            acc = rcons(m, acc); ls.erase(ls.begin(), ls.begin() + 2);
        }
        if (ls.size() == 1) return mergel(rcons(rev_onto(ls[0], {}), acc), {}, asc);
        return mergel(acc, {}, asc);
    }
    // fmerge2_1/fmerge2_2: t1 is the prioritised list, second = [H2|T2]; result is built reversed
    List merge_fwd(const List& t1, const List& second) {
        size_t i = 0, j = 1; T h2 = second[0]; List m;
        for (;;) {
            if (i == t1.size()) { m.insert(m.begin(), h2); for (; j < second.size(); j++) m.insert(m.begin(), second[j]); return m; }
            const T& h1 = t1[i];
            if (fun_(h1, h2)) { m.insert(m.begin(), h1); i++; continue; }
            m.insert(m.begin(), h2);
            for (;;) {
                if (j == second.size()) { m.insert(m.begin(), h1); for (size_t k = i + 1; k < t1.size(); k++) m.insert(m.begin(), t1[k]); return m; }
                const T& n2 = second[j];
                if (fun_(h1, n2)) { m.insert(m.begin(), h1); i++; h2 = n2; j++; break; }
                m.insert(m.begin(), n2); j++;
            }
        }
    }
    // rfmerge2_1/rfmerge2_2
    List merge_rev(const List& t1, const List& second) {
        size_t i = 0, j = 1; T h2 = second[0]; List m;
        for (;;) {
            if (i == t1.size()) { m.insert(m.begin(), h2); for (; j < second.size(); j++) m.insert(m.begin(), second[j]); return m; }
            const T& h1 = t1[i];
            if (!fun_(h1, h2)) { m.insert(m.begin(), h1); i++; continue; }
            m.insert(m.begin(), h2);
            for (;;) {
                if (j == second.size()) { m.insert(m.begin(), h1); for (size_t k = i + 1; k < t1.size(); k++) m.insert(m.begin(), t1[k]); return m; }
                const T& n2 = second[j];
                if (!fun_(h1, n2)) { m.insert(m.begin(), h1); i++; h2 = n2; j++; break; }
                m.insert(m.begin(), n2); j++;
            }
        }
    }
};

}  // namespace eb
