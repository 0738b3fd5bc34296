// ORACLE (test infrastructure only -- never linked into the product path).
// Restatement of erlamsa_mutations (src/erlamsa_mutations.erl) + erlamsa_generic
// (src/erlamsa_generic.erl) + erlamsa_field_predict (src/erlamsa_field_predict.erl):
// every mutator as a scalar function with the reference's exact RNG draw order,
// and the weighted scheduler (mux_fuzzers :1258-1280).
#pragma once
#include <zlib.h>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include "common.hpp"
#include "erl_lists.hpp"
#include "strlex.hpp"
#include "tree.hpp"
#include "fuse.hpp"

namespace eo {

// One row of the scheduler's list: {Score, Pri, Fn, Name} plus the state the
// reference keeps inside the closure Fn.
struct MutNode {
    double score = 0; int pri = 0;
    int name = 0;     // MutId reported in meta
    int fn = 0;       // MutId of the function currently bound (uri rebinds itself to b64, :784)
    // construct_st_line_muta state [Count, L1..L10] (src/erlamsa_generic.erl:122-162):
    // each stored line is {first element, rest} because step_state replaces only the head cons.
    std::vector<std::pair<Bin, Bin>> st;
    // remember/1 closure of sed_fuse_old :404-427
    bool has_block = false; Bin block;
};

struct MutRes { Blocks ll; double delta = 0; };

struct Mutations {
    Rng& rng; const Opts& opts;
    int snand_kind = 0;   // 0 nand, 1 or, 2 xor -- fixed when the table is built (:1313)
    int64_t thread_seed[3] = {0, 0, 0};   // of the case being run (donor choice, see fuse_old)
    Mutations(Rng& r, const Opts& o) : rng(r), opts(o) {}

    // mutations/1 :1290-1332 -- building the table costs two draws
    void build_table_draws() {
        snand_kind = (int)rng.rand_elem_idx(3);
        (void)rng.rand_elem_idx(1);
    }

    // ------------------------------------------------------------ byte level
    // edit_byte_vector/3 :56-61 callers (:176-223)
    MutRes byte_muta(int id, const Blocks& ll) {
        const Bin& h = ll[0]; MutRes r; r.ll = ll;
        uint64_t p = rng.rand(h.size());
        int d = rng.rand_delta(); r.delta = d;
        if (h.empty()) return r;
        uint8_t b = (uint8_t)h[p]; Bin c;
        switch (id) {
        case M_BD: break;
        case M_BEI: c.push_back((char)((b + 1) & 255)); break;
        case M_BED: c.push_back((char)((b - 1) & 255)); break;
        case M_BR: c.push_back((char)b); c.push_back((char)b); break;
        case M_BF: c.push_back((char)(b ^ (1u << rng.rand(8)))); break;
        case M_BI: c.push_back((char)rng.rand(256)); c.push_back((char)b); break;
        case M_BER: c.push_back((char)rng.rand(256)); break;
        }
        r.ll[0] = h.substr(0, p) + c + h.substr(p + 1);
        return r;
    }
    // sed_utf8_widen :1080-1089, sed_utf8_insert :1091-1099
    static const std::vector<Bin>& funny_unicode() {
        static std::vector<Bin> v;
        if (!v.empty()) return v;
        auto B = [](std::initializer_list<int> l) { Bin s; for (int x : l) s.push_back((char)x); return s; };
        v = {B({239, 191, 191}), B({240, 144, 128, 128}), B({0xef, 0xbb, 0xbf}), B({0xfe, 0xff}), B({0xff, 0xfe}),
             B({0, 0, 0xff, 0xff}), B({0xff, 0xff, 0, 0}), B({43, 47, 118, 56}), B({43, 47, 118, 57}), B({43, 47, 118, 43}),
             B({43, 47, 118, 47}), B({247, 100, 76}), B({221, 115, 102, 115}), B({14, 254, 255}), B({251, 238, 40}),
             B({251, 238, 40, 255}), B({132, 49, 149, 51})};
        // Codes :1065-1072; foldl prepends each group (ranges expanded ascending)
        std::vector<std::pair<int, int>> codes = {
            {0x0009, 0x000d}, {0x008D, -1}, {0x00a0, -1}, {0x1680, -1}, {0x180e, -1}, {0x2000, 0x200a}, {0x2028, -1}, {0x2029, -1},
            {0x202f, -1}, {0x205f, -1}, {0x3000, -1}, {0x200e, 0x200f}, {0x202a, 0x202e}, {0x200c, 0x200d}, {0x0345, -1}, {0x00b7, -1},
            {0x02d0, 0x02d1}, {0xff70, -1}, {0x02b0, 0x02b8}, {0xfdd0, -1}, {0x034f, -1}, {0x115f, 0x1160}, {0x2065, 0x2069},
            {0x3164, -1}, {0xffa0, -1}, {0xe0001, -1}, {0xe0020, 0xe007f}, {0x0e40, 0x0e44}, {0x1f4a9, -1}};
        std::vector<int> numbers;
        for (auto& c : codes) {
            std::vector<int> grp; if (c.second < 0) grp.push_back(c.first); else for (int x = c.first; x <= c.second; x++) grp.push_back(x);
            numbers.insert(numbers.begin(), grp.begin(), grp.end());
        }
        auto ext = [](int n) { return (n & 0x3f) | 0x80; };
        for (int p : numbers) {   // encode_point/1 :1036-1049
            Bin s;
            if (p < 0x80) s.push_back((char)p);
            else if (p < 0x800) { s.push_back((char)(0xc0 | (0x1f & (p >> 6)))); s.push_back((char)ext(p)); }
            else if (p < 0x10000) { s.push_back((char)(0xe0 | (0x0f & (p >> 12)))); s.push_back((char)ext(p >> 6)); s.push_back((char)ext(p)); }
            else { s.push_back((char)(0xf0 | (0x7 & (p >> 18)))); s.push_back((char)ext(p >> 12)); s.push_back((char)ext(p >> 6)); s.push_back((char)ext(p)); }
            v.push_back(s);
        }
        return v;
    }
    MutRes utf8_widen(const Blocks& ll) {
        const Bin& h = ll[0]; MutRes r; r.ll = ll;
        uint64_t p = rng.rand(h.size()); r.delta = rng.rand_delta();
        if (h.empty()) return r;
        uint8_t b = (uint8_t)h[p];
        if (b == (b & 0x3f)) { Bin c; c.push_back((char)0xc0); c.push_back((char)(b | 0x80)); r.ll[0] = h.substr(0, p) + c + h.substr(p + 1); }
        return r;
    }
    MutRes utf8_insert(const Blocks& ll) {
        const Bin& h = ll[0]; MutRes r; r.ll = ll;
        uint64_t p = rng.rand(h.size()); r.delta = rng.rand_delta();
        const Bin& u = funny_unicode()[rng.rand_elem_idx(funny_unicode().size())];
        if (h.empty()) return r;
        r.ll[0] = h.substr(0, p + 1) + u + h.substr(p + 1);
        return r;
    }

    // ------------------------------------------------------------ byte sequences :232-318
    Bin randmask(int kind, const Bin& bs) {   // kind: 0 nand 1 or 2 xor 3 replace; :279-307
        uint64_t prob = rng.erand(100);
        bool occ = rng.rand_occurs_fixed(prob, 100);
        Bin out;
        for (size_t i = 0; i < bs.size(); i++) {
            bool next = rng.rand_occurs_fixed(prob, 100);   // argument 3 is evaluated before argument 5
            uint8_t h = (uint8_t)bs[i];
            if (occ) {
                switch (kind) {
                case 0: h = h & (uint8_t)~(1u << rng.rand(8)); break;
                case 1: h = h | (uint8_t)(1u << rng.rand(8)); break;
                case 2: h = h ^ (uint8_t)(1u << rng.rand(8)); break;
                default: h = (uint8_t)rng.rand(256);
                }
            }
            out.push_back((char)h); occ = next;
        }
        return out;
    }
    MutRes bytes_muta(int id, const Blocks& ll) {
        const Bin& v = ll[0]; MutRes r; r.ll = ll;
        if (v.empty()) { r.delta = -1; return r; }
        uint64_t n = v.size();
        uint64_t s = rng.rand(n);
        uint64_t l = (uint64_t)rng.rand_range(1, (int64_t)(n - s + 1));
        Bin h = v.substr(0, s), p = v.substr(s, l), t = v.substr(s + l), c;
        switch (id) {
        case M_SP: {
            std::vector<uint8_t> bl(p.begin(), p.end());
            auto pm = rng.random_permutation(bl, [](uint8_t a, uint8_t b) { return a < b; });
            c.assign(pm.begin(), pm.end()); break;
        }
        case M_SR: {
            uint64_t k = std::max<uint64_t>(2, rng.rand_log_u64(10));
            for (uint64_t i = 0; i < k; i++) c += p; break;
        }
        case M_SD: break;
        case M_SNAND: c = randmask(snand_kind, p); break;
        case M_SRND: c = randmask(3, p); break;
        }
        r.ll[0] = h + c + t;
        r.delta = rng.rand_delta();
        return r;
    }

    // ------------------------------------------------------------ sed_num :63-169
    static const std::vector<BigInt>& interesting_numbers() {
        static std::vector<BigInt> v;
        if (v.empty()) {
            const int is[] = {1, 7, 8, 15, 16, 31, 32, 63, 64, 127, 128};
            for (int i : is) { BigInt x = BigInt::pow2(i); std::vector<BigInt> g = {x - BigInt(1), x, x + BigInt(1)}; v.insert(v.begin(), g.begin(), g.end()); }
        }
        return v;
    }
    BigInt mutate_num(const BigInt& num) {
        uint64_t c = rng.rand(12);
        const auto& in = interesting_numbers();
        switch (c) {
        case 0: return num + BigInt(1);
        case 1: return num - BigInt(1);
        case 2: return BigInt(0);
        case 3: return BigInt(1);
        case 4: case 5: return in[rng.rand_elem_idx(in.size())];
        case 7: return num + in[rng.rand_elem_idx(in.size())];
        case 8: return num - in[rng.rand_elem_idx(in.size())];
        case 9: { BigInt r = rng.rand_big(num.abs().mul_small(2)); return num - (num.neg ? -r : r); }
        case 10: return -num;
        default: {   // 6 and 11
            uint64_t n = (uint64_t)rng.rand_range(1, 129);
            BigInt l = rng.rand_log(n);
            uint64_t s = rng.rand(3);
            return s == 0 ? num - l : num + l;
        }
        }
    }
    MutRes sed_num(const Blocks& ll) {
        const Bin& h = ll[0]; size_t n = h.size();
        // tokenise exactly like get_num/4: '-'* then digits; a lone '-' run is not a number
        struct Tok { size_t a, b; bool neg; size_t d0; };
        std::vector<Tok> toks; size_t p = 0;
        while (p < n) {
            size_t q = p; bool neg = false;
            while (q < n && h[q] == '-') { neg = true; q++; }
            size_t d0 = q;
            while (q < n && h[q] >= '0' && h[q] <= '9') q++;
            if (q > d0) { toks.push_back({p, q, neg, d0}); p = q; } else p++;
        }
        uint64_t which = rng.rand(toks.size());
        Bin lst = h; int nres = 0;
        if (!toks.empty()) {
            const Tok& t = toks[toks.size() - 1 - which];   // counted from the last number backwards
            BigInt val = BigInt::from_decimal(h.substr(t.d0, t.b - t.d0), t.neg);
            BigInt nv = mutate_num(val);
            lst = h.substr(0, t.a) + nv.to_string() + h.substr(t.b);
            nres = -1;
        }
        bool isbin = binarish(lst);
        MutRes r; flush_bvecs(lst, r.ll); r.ll.insert(r.ll.end(), ll.begin() + 1, ll.end());
        if (nres == 0) { uint64_t x = rng.rand(10); r.delta = x == 0 ? -1 : 0; }
        else if (isbin) r.delta = -1;
        else r.delta = +2;
        return r;
    }

    // ------------------------------------------------------------ lines :320-378 + erlamsa_generic
    static std::vector<Bin> lines(const Bin& b) {
        std::vector<Bin> out; size_t s = 0;
        for (size_t i = 0; i < b.size(); i++) if (b[i] == 10) { out.push_back(b.substr(s, i + 1 - s)); s = i + 1; }
        if (s < b.size()) out.push_back(b.substr(s));
        return out;
    }
    static Bin unlines(const std::vector<Bin>& ls) { Bin o; for (auto& l : ls) o += l; return o; }
    void line_op(int id, std::vector<Bin>& l) {
        size_t len = l.size();   // never 0 here (try_lines refuses the empty block)
        switch (id) {
        case M_LD: { uint64_t p = rng.erand(len); l.erase(l.begin() + (p - 1)); break; }
        case M_LDS: { uint64_t st = rng.erand(len); uint64_t n = rng.erand(len - st + 1); l.erase(l.begin() + (st - 1), l.begin() + (st - 1 + n)); break; }
        case M_LR2: { uint64_t p = rng.erand(len); l.insert(l.begin() + (p - 1), l[p - 1]); break; }
        case M_LR: { uint64_t p = rng.erand(len); uint64_t n = std::max<uint64_t>(2, rng.rand_log_u64(10)); Bin e = l[p - 1]; l.insert(l.begin() + (p - 1), n - 1, e); break; }
        case M_LRI: { uint64_t from = rng.erand(len); uint64_t to = rng.erand(len); Bin e = l[from - 1]; l[to - 1] = e; break; }
        case M_LS: { if (len < 2) break; uint64_t p = rng.erand(len - 1); std::swap(l[p - 1], l[p]); break; }
        case M_LP: {
            if (len < 3) break;
            uint64_t from = rng.erand(len - 1);
            uint64_t a = (uint64_t)rng.rand_range(2, (int64_t)(len - from));
            uint64_t b = rng.rand_log_u64(10);
            uint64_t n = std::max<uint64_t>(2, std::min(a, b));
            std::vector<Bin> seg(l.begin() + (from - 1), l.begin() + (from - 1 + n));
            auto pm = rng.random_permutation(seg, [](const Bin& x, const Bin& y) { return x < y; });
            std::copy(pm.begin(), pm.end(), l.begin() + (from - 1)); break;
        }
        }
    }
    MutRes line_muta(int id, const Blocks& ll) {
        MutRes r; r.ll = ll; const Bin& h = ll[0];
        std::vector<Bin> ls = lines(h);
        if (ls.empty() || binarish(h)) { r.delta = -1; return r; }
        line_op(id, ls);
        r.ll[0] = unlines(ls); r.delta = 1; return r;
    }
    // st_list_ins / st_list_replace, src/erlamsa_generic.erl:122-162
    MutRes st_line_muta(MutNode& node, const Blocks& ll) {
        MutRes r; r.ll = ll; const Bin& h = ll[0];
        std::vector<Bin> ls = lines(h);
        if (ls.empty() || binarish(h)) { r.delta = -1; return r; }
        size_t n = ls.size();
        auto store = [](const Bin& line) { return std::make_pair(line.substr(0, 1), line.substr(1)); };
        if (node.st.size() < 10) {
            // step_state recurses while Count < 10; each new element is consed right after the count
            while (node.st.size() < 10) { uint64_t p = rng.erand(n); node.st.insert(node.st.begin(), store(ls[p - 1])); }
        }
        uint64_t up = rng.erand(20);
        if (up < 10) { uint64_t ep = rng.erand(n); node.st[up - 1].first = ls[ep - 1]; }   // [New | tl(Old)]
        uint64_t pk = rng.erand(node.st.size());
        Bin x = node.st[pk - 1].first + node.st[pk - 1].second;
        uint64_t p = rng.erand(n);
        if (node.fn == M_LIS) ls.insert(ls.begin() + (p - 1), x); else ls[p - 1] = x;
        r.ll[0] = unlines(ls); r.delta = 1; return r;
    }

    // ------------------------------------------------------------ fuse :386-427
    static void halve(const Bin& l, Bin& a, Bin& b) { size_t k = l.size() / 2; a = l.substr(0, k); b = l.substr(k); }
    MutRes fuse_this(const Blocks& ll) {
        MutRes r; r.ll = ll; r.ll[0] = fuse(rng, ll[0], ll[0]); r.delta = rng.rand_delta(); return r;
    }
    MutRes fuse_next(const Blocks& ll) {
        Bin a1, a2; halve(ll[0], a1, a2);
        const Bin& b = ll.size() > 1 ? ll[1] : ll[0];
        Bin abl = fuse(rng, a1, b);
        Bin abal = fuse(rng, abl, a2);
        MutRes r; r.delta = rng.rand_delta();
        flush_bvecs(abal, r.ll);
        if (ll.size() > 2) r.ll.insert(r.ll.end(), ll.begin() + 2, ll.end());
        return r;
    }
    MutRes fuse_old(MutNode& node, const Blocks& ll) {
        const Bin& h = ll[0];
        // remember/1: the closure starts out remembering H itself (:424-427). With a donor pool (config C5: windows of
        // other seeds, all-gathered across GPUs) the FIRST remembered block is a donor instead -- chosen from the case's
        // thread seed, so that no draw of the case's own stream is consumed and everything else stays the reference's.
        if (!node.has_block) {
            node.has_block = true; node.block = h;
            if (opts.donor_pool && opts.n_donors) {
                uint64_t d = (uint64_t)(thread_seed[0] * 31 + thread_seed[1] * 17 + thread_seed[2]) % opts.n_donors;
                node.block = Bin((const char*)opts.donor_pool + d * opts.donor_stride, opts.donor_len[d]);
            }
        }
        Bin al1, al2, ol1, ol2; halve(h, al1, al2); halve(node.block, ol1, ol2);
        Bin a = fuse(rng, al1, ol1);
        Bin b = fuse(rng, ol2, al2);
        uint64_t swap = rng.rand(3);
        MutRes r; r.delta = rng.rand_delta();
        if (swap == 0) node.block = h;
        flush_bvecs(a, r.ll); flush_bvecs(b, r.ll);
        r.ll.insert(r.ll.end(), ll.begin() + 1, ll.end());
        return r;
    }

    // ------------------------------------------------------------ ascii :586-651
    MutRes ascii_muta(int id, const Blocks& ll) {
        MutRes r; r.ll = ll;
        Chunks cs = lex(ll[0]);
        if (!stringy(cs)) { r.delta = -1; return r; }
        TextCtx tc{&rng, opts.ssrf_host, opts.ssrf_port};
        if (id == M_AB) string_generic_mutate(tc, cs, {INSERT_BADNESS, REPLACE_BADNESS, INSERT_TRAVERSAL, INSERT_AAAS, INSERT_NULL});
        else string_delimeter_mutate(tc, cs);
        r.delta = rng.rand_delta();
        r.ll[0] = unlex(cs); return r;
    }

    // ------------------------------------------------------------ tree :907-1023
    MutRes tree_muta(int id, const Blocks& ll) {
        MutRes r; r.ll = ll; const Bin& h = ll[0];
        if (binarish(h)) { r.delta = -1; return r; }
        TreeParser tp{(const uint8_t*)h.data(), h.size()};
        const bool trace = getenv("EO_TRACE") != nullptr;
        std::vector<Term> lst = tp.parse();
        if (trace) fprintf(stderr, "[eo]   tree: parsed, clock %ld\n", (long)clock());
        std::vector<const Term*> subs = sublists_erl(lst);
        if (trace) fprintf(stderr, "[eo]   tree: %zu sublists, clock %ld\n", subs.size(), (long)clock());
        Bin out;
        if (id == M_TR2 || id == M_TD) {
            const Term* sub = subs.empty() ? nullptr : subs[rng.rand_elem_idx(subs.size())];
            auto dup = [](const std::vector<Term>& l, size_t i, Bin& o) { flatten_into(l[i], o); flatten_tail(l, i, o); };
            auto del = [](const std::vector<Term>& l, size_t i, Bin& o) { flatten_tail(l, i + 1, o); };
            if (id == M_TR2) edit_sublist_emit(lst, sub, dup, out); else edit_sublist_emit(lst, sub, del, out);
            r.ll[0] = out; r.delta = 1; return r;
        }
        if (id == M_TS1 || id == M_TS2) {
            if (subs.size() < 2) { r.delta = -1; return r; }
            std::vector<size_t> idx = rng.reservoir_sample_idx(subs.size(), 2);
            if (id == M_TS1) {
                if (rng.rand(2) == 1) std::swap(idx[0], idx[1]);   // random_permutation of a 2-list
                const Term* a = subs[idx[0]]; const Term* b = subs[idx[1]];
                auto op = [&](const std::vector<Term>& l, size_t i, Bin& o) { flatten_into(*b, o); flatten_tail(l, i + 1, o); };
                edit_sublist_emit(lst, a, op, out);
            } else {
                edit_sublists_emit(lst, *subs[idx[0]], *subs[idx[1]], out);
            }
            r.ll[0] = out; r.delta = 1; return r;
        }
        // M_TR: sed_tree_stutter
        std::vector<const Term*> rs = rng.random_permutation(subs, [](const Term* a, const Term* b) { return a < b; });
        const Term* parent = nullptr; const Term* child = nullptr;
        for (const Term* cand : rs) {
            std::vector<const Term*> cs = sublists_erl(cand->k);
            if (cs.empty()) continue;
            child = cs[rng.rand_elem_idx(cs.size())]; parent = cand; break;
        }
        uint64_t reps = rng.rand_log_u64(10);
        if (trace) fprintf(stderr, "[eo]   tree: stutter reps %llu, clock %ld\n", (unsigned long long)reps, (long)clock());
        if (!parent) { r.delta = -1; return r; }
        auto op = [&](const std::vector<Term>& l, size_t i, Bin& o) { repeat_path_emit(*parent, *child, reps, o); flatten_tail(l, i + 1, o); };
        edit_sublist_emit(lst, child, op, out);
        r.ll[0] = out; r.delta = 1; return r;
    }

    // ------------------------------------------------------------ length field :1107-1143 + field_predict
    struct Sizer { int size; bool big; uint64_t len, a, b; };
    static uint64_t rd(const Bin& s, uint64_t a, int bytes, bool big) {
        uint64_t v = 0;
        for (int i = 0; i < bytes; i++) { uint8_t c = (uint8_t)s[a + (big ? i : bytes - 1 - i)]; v = (v << 8) | c; }
        return v;
    }
    static void basic_len(int64_t a, int64_t b, const Bin& bin, std::vector<Sizer>& out) {   // :66-78
        int64_t sz = (int64_t)bin.size();
        if (!(a < b && b > 0 && a < sz)) return;
        static const int widths[3] = {2, 4, 8};
        for (int e = 0; e < 2; e++) for (int w : widths) {
            if (a + w > sz) continue;
            uint64_t len = rd(bin, a, w, e == 0);
            int64_t want = b - a - w;
            if (want >= 0 && len == (uint64_t)want && len > 2) { out.push_back({w * 8, e == 0, len, (uint64_t)a, (uint64_t)b}); return; }
        }
    }
    static void simple_len(int64_t a, int64_t b, const Bin& bin, std::vector<Sizer>& out) {   // :80-88
        basic_len(a, b, bin, out); basic_len(a, b - 1, bin, out); basic_len(a, b - 2, bin, out); basic_len(a, b - 4, bin, out); basic_len(a, b - 8, bin, out);
    }
    static void simple_u8len(int64_t a, const Bin& bin, std::vector<Sizer>& out) {   // :51-64
        int64_t sz = (int64_t)bin.size();
        for (int x = 0; x <= 8; x++) {
            int64_t b = sz - x;
            if (!(a < b && b > 0 && a < sz)) continue;
            uint64_t len = (uint8_t)bin[a]; int64_t want = b - a - 1;
            if (want >= 0 && len == (uint64_t)want && len > 2) out.push_back({8, true, len, (uint64_t)a, (uint64_t)b});
        }
    }
    std::vector<Sizer> get_possible_simple_lens(const Bin& bin) {   // :90-105
        std::vector<Sizer> out; int64_t len = (int64_t)bin.size();
        if (len > 10) {
            int64_t sub = std::min<int64_t>(len / 5, SIZER_MAX_FIRST_BYTES);
            std::vector<int64_t> varb;
            for (int64_t i = 0; i <= sub; i++) varb.push_back(rng.rand_range(sub, len));
            for (int64_t a = 0; a <= sub; a++) simple_u8len(a, bin, out);           // SmallLens first
            // BigLens = foldl-reversed AllRanges; AllRanges = [{A,Len}..] ++ [{X,Y} || X, Y]
            for (int64_t x = sub; x >= 0; x--) for (size_t yi = varb.size(); yi-- > 0;) simple_len(x, varb[yi], bin, out);
            for (int64_t a = sub; a >= 0; a--) simple_len(a, len, bin, out);
        } else {
            for (int64_t x = 0; x <= 3; x++) { simple_len(x, len, bin, out); simple_u8len(x, bin, out); }
        }
        return out;
    }
    static Bin enc(uint64_t v, int bits, bool big) {
        Bin s; int bytes = bits / 8;
        for (int i = 0; i < bytes; i++) { int sh = big ? (bytes - 1 - i) * 8 : i * 8; s.push_back((char)(sh >= 64 ? 0 : (v >> sh) & 255)); }
        return s;
    }
    Bin fast_pseudorandom_block(uint64_t n) {   // src/erlamsa_rnd.erl:155-160
        if (n < ABSMAXHALF_BINARY_BLOCK) return rng.random_block(n);
        Bin rnd = rng.random_block(ABSMAXHALF_BINARY_BLOCK);
        uint64_t zbits = n - ABSMAXHALF_BINARY_BLOCK;
        if (zbits % 8) throw CaseDied("fast_pseudorandom_block: non byte-aligned bitstring");
        Bin z(zbits / 8, '\0'); if (!z.empty()) z[z.size() - 1] = 42;
        return z + rnd;
    }
    MutRes length_predict(const Blocks& ll) {
        MutRes r; r.ll = ll; const Bin& bin = ll[0];
        std::vector<Sizer> els = get_possible_simple_lens(bin);
        int64_t ei = rng.rand_elem_idx(els.size());
        if (ei < 0) { r.delta = -2; return r; }
        const Sizer& e = els[ei];
        Bin h = bin.substr(0, e.a); int fb = e.size / 8;
        Bin blob = bin.substr(e.a + fb, e.len), rest = bin.substr(e.a + fb + e.len);
        Bin tmp = rng.random_block(fb);
        // <<TmpNewLen:Size>> big-endian; NewLen = min(1e6, 2*Tmp) -- 2*Tmp may exceed 64 bits
        uint64_t tv = 0; bool huge = false;
        for (int i = 0; i < fb; i++) { if (tv >> 56) huge = true; tv = (tv << 8) | (uint8_t)tmp[i]; }
        uint64_t newlen = (huge || tv > ABSMAX_BINARY_BLOCK) ? ABSMAX_BINARY_BLOCK : std::min<uint64_t>(ABSMAX_BINARY_BLOCK, tv * 2);
        uint64_t c = rng.rand(7);
        Bin res;
        switch (c) {
        case 0: res = h + Bin(fb, '\0') + blob + rest; break;
        case 1: res = h + Bin(fb, '\xff') + blob + rest; break;
        case 2: { Bin rb = fast_pseudorandom_block(newlen); res = h + enc(e.len, e.size, e.big) + blob + rb + rest; break; }
        case 3: res = h + enc(newlen, e.size, e.big) + rest; break;
        default: res = h + enc(newlen, e.size, e.big) + blob + rest;
        }
        r.ll[0] = res; r.delta = 1; return r;
    }

    // ------------------------------------------------------------ uri :734-784
    static Bin change_scheme_rev(const Bin& acc_rev) {   // Acc is the reversed prefix
        Bin fwd(acc_rev.rbegin(), acc_rev.rend());
        if (acc_rev.size() >= 4 && acc_rev.compare(0, 4, "elif") == 0) return fwd.substr(0, fwd.size() - 4) + "http";
        return fwd;
    }
    static std::vector<Bin> tokens(const Bin& s, char sep) {   // string:tokens/2
        std::vector<Bin> o; Bin cur;
        for (char c : s) { if (c == sep) { if (!cur.empty()) o.push_back(cur); cur.clear(); } else cur.push_back(c); }
        if (!cur.empty()) o.push_back(cur);
        return o;
    }
    static Bin join(const std::vector<Bin>& v, size_t from, const Bin& sep) { Bin o; for (size_t i = from; i < v.size(); i++) { if (i > from) o += sep; o += v[i]; } return o; }
    bool try_uri_mutate(const Bin& l, Bin& out) {
        size_t p = l.find("://");
        if (p == Bin::npos) { out = l; return false; }
        Bin acc_fwd = l.substr(0, p), t = l.substr(p + 3);
        Bin acc_rev(acc_fwd.rbegin(), acc_fwd.rend());
        uint64_t k = rng.erand(3);
        Bin hostport = opts.ssrf_host + ":" + std::to_string(opts.ssrf_port);
        if (k == 1) { out = change_scheme_rev(acc_rev) + "://" + hostport + "/" + t; return true; }
        if (k == 2) {
            static const std::vector<Bin> f = {" @", "@"};
            Bin at = f[rng.rand_elem_idx(2)] + hostport;
            std::vector<Bin> tk = tokens(t, '/');
            if (tk.empty()) throw CaseDied("uri: badmatch on empty domain");
            out = change_scheme_rev(acc_rev) + "://" + tk[0] + at + "/" + join(tk, 1, "/"); return true;
        }
        std::vector<Bin> tk = tokens(t, '/');
        if (tk.empty()) throw CaseDied("uri: badmatch on empty domain");
        uint64_t nt = rng.erand(10); Bin trav = "/"; for (uint64_t i = 0; i < nt; i++) trav += "../";
        uint64_t w = rng.erand(4);
        Bin q = w == 1 ? join(tk, 1, "/") : w == 2 ? "Windows/win.ini" : w == 3 ? "etc/shadow" : "etc/passwd";
        out = acc_fwd + "://" + tk[0] + trav + q; return true;
    }
    MutRes uri_mutator(const Blocks& ll) {
        MutRes r; r.ll = ll; Chunks cs = lex(ll[0]); double d = -1;
        for (Chunk& c : cs) if (c.type == Chunk::TEXT && c.bytes.size() > 5) { Bin nb; if (try_uri_mutate(c.bytes, nb)) d += 1; c.bytes = nb; }
        r.ll[0] = unlex(cs); r.delta = d; return r;
    }

    // ------------------------------------------------------------ base64 :657-690
    static bool b64_decode(const Bin& in, Bin& out) {   // base64:decode/1 (strict alphabet, skips whitespace, needs full quanta)
        auto val = [](uint8_t c) -> int {
            if (c >= 'A' && c <= 'Z') return c - 'A'; if (c >= 'a' && c <= 'z') return c - 'a' + 26;
            if (c >= '0' && c <= '9') return c - '0' + 52; if (c == '+') return 62; if (c == '/') return 63;
            if (c == '=') return -2; if (c == ' ' || c == '\t' || c == '\n' || c == '\r') return -3; return -1;
        };
        std::vector<int> v; for (uint8_t c : in) { int x = val(c); if (x == -1) return false; if (x == -3) continue; v.push_back(x); }
        if (v.size() % 4) return false;
        out.clear();
        for (size_t i = 0; i < v.size(); i += 4) {
            int a = v[i], b = v[i + 1], c = v[i + 2], d = v[i + 3];
            bool last = i + 4 == v.size();
            if (a < 0 || b < 0) return false;
            if (c == -2) { if (d != -2 || !last) return false; out.push_back((char)((a << 2) | (b >> 4))); continue; }
            if (d == -2) { if (!last) return false; out.push_back((char)((a << 2) | (b >> 4))); out.push_back((char)(((b & 15) << 4) | (c >> 2))); continue; }
            out.push_back((char)((a << 2) | (b >> 4))); out.push_back((char)(((b & 15) << 4) | (c >> 2))); out.push_back((char)(((c & 3) << 6) | d));
        }
        return true;
    }
    static Bin b64_encode(const Bin& in) {
        static const char* tb = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"; Bin o; size_t i = 0;
        for (; i + 3 <= in.size(); i += 3) { uint32_t x = ((uint8_t)in[i] << 16) | ((uint8_t)in[i + 1] << 8) | (uint8_t)in[i + 2]; o.push_back(tb[x >> 18]); o.push_back(tb[(x >> 12) & 63]); o.push_back(tb[(x >> 6) & 63]); o.push_back(tb[x & 63]); }
        if (in.size() - i == 1) { uint32_t x = (uint8_t)in[i] << 16; o.push_back(tb[x >> 18]); o.push_back(tb[(x >> 12) & 63]); o += "=="; }
        else if (in.size() - i == 2) { uint32_t x = ((uint8_t)in[i] << 16) | ((uint8_t)in[i + 1] << 8); o.push_back(tb[x >> 18]); o.push_back(tb[(x >> 12) & 63]); o.push_back(tb[(x >> 6) & 63]); o.push_back('='); }
        return o;
    }
    MutRes base64_mutator(const Blocks& ll);   // defined after the scheduler (it runs a nested one)
    // sgml.hpp / json.hpp (they run nested schedulers on inner text)
    int depth = 0;
    std::string ssrf_uri() const;
    void inner_muta(const std::vector<int>& ids, std::unique_ptr<Mutations>& m, Opts& o2, std::vector<MutNode>& nodes);
    static Bin inner_round(Mutations& m, const std::vector<MutNode>& nodes, const Bin& b);
    MutRes sgml_mutate(const Blocks& ll);
    MutRes json_mutate(const Blocks& ll);

    // ------------------------------------------------------------ scheduler :1234-1280, 1385-1395
    std::vector<MutNode> make_mutator_nodes() {   // make_mutator/2 + mutators_mutator/1
        std::vector<MutNode> sel;
        for (int i = 0; i < M_COUNT; i++) if (opts.muta_pri[i] >= 0) { MutNode n; n.pri = opts.muta_pri[i]; n.name = n.fn = i; sel.push_back(n); }
        // scores are drawn walking the REVERSED table (nil first ... sgm last)
        for (size_t i = sel.size(); i-- > 0;) { uint64_t s = rng.rand(10); sel[i].score = (double)std::max<uint64_t>(2, s); }
        return sel;
    }
    MutRes apply(MutNode& node, const Blocks& ll) {
        if (getenv("EO_TRACE")) fprintf(stderr, "[eo] %*sapply %s on %zu bytes (%zu blocks) draws=%llu\n", depth * 2, "", MUT_CODES[node.fn], ll[0].size(), ll.size(), (unsigned long long)rng.draws);
        switch (node.fn) {
        case M_BD: case M_BEI: case M_BED: case M_BF: case M_BI: case M_BER: case M_BR: return byte_muta(node.fn, ll);
        case M_UW: return utf8_widen(ll);
        case M_UI: return utf8_insert(ll);
        case M_SP: case M_SR: case M_SD: case M_SNAND: case M_SRND: return bytes_muta(node.fn, ll);
        case M_NUM: return sed_num(ll);
        case M_LD: case M_LDS: case M_LR2: case M_LRI: case M_LR: case M_LS: case M_LP: return line_muta(node.fn, ll);
        case M_LIS: case M_LRS: return st_line_muta(node, ll);
        case M_FT: return fuse_this(ll);
        case M_FN: return fuse_next(ll);
        case M_FO: return fuse_old(node, ll);
        case M_AB: case M_AD: return ascii_muta(node.fn, ll);
        case M_TR2: case M_TD: case M_TS1: case M_TS2: case M_TR: return tree_muta(node.fn, ll);
        case M_LEN: return length_predict(ll);
        case M_SGM: return sgml_mutate(ll);
        case M_JS: return json_mutate(ll);
        case M_URI: { MutRes r = uri_mutator(ll); node.fn = M_B64; return r; }
        case M_B64: return base64_mutator(ll);
        case M_NIL: { MutRes r; r.ll = ll; r.delta = -1; return r; }
        case M_ZIP: {   // zip_path_traversal :1149-1163: zip:foldl fails on non-archives -> unchanged, delta -1, no draws
            if (ll[0].find(std::string("PK\x05\x06", 4)) != Bin::npos) throw Unsupported("zip mutator on ZIP-looking data");
            MutRes r; r.ll = ll; r.delta = -1; return r;
        }
        default: throw Unsupported(std::string("mutator not restated in the oracle: ") + MUT_CODES[node.fn]);
        }
    }
    static double adjust_priority(double pri, double delta) {   // :1238-1242
        if (delta == 0) return pri;
        return std::max(2.0, std::min(10.0, pri + delta));
    }
    // mux_fuzzers closure: one round. Returns the new block list; fs is replaced by the reordered list.
    Blocks mux_fuzzers(std::vector<MutNode>& fs, const Blocks& ll, Meta* meta) {
        if (ll.size() == 1 && ll[0].empty()) return ll;
        if (ll.empty()) return ll;
        // weighted_permutations/1
        std::vector<std::pair<uint64_t, size_t>> keyed;
        for (size_t i = 0; i < fs.size(); i++) keyed.push_back({rng.rand((uint64_t)std::trunc(fs[i].score * fs[i].pri)), i});
        ErlSort<std::pair<uint64_t, size_t>> sorter([](const std::pair<uint64_t, size_t>& a, const std::pair<uint64_t, size_t>& b) { return a.first >= b.first; });
        keyed = sorter.sort(keyed);
        std::vector<MutNode> perm; for (auto& k : keyed) perm.push_back(fs[k.second]);
        std::vector<MutNode> out;   // holds the Erlang `Out` list REVERSED
        for (size_t i = 0; i < perm.size(); i++) {
            if (ll[0].size() > ABSMAX_BINARY_BLOCK) {
                // {mux_fuzzers(Out ++ Tail), Ll, [{skipped_big,..}]} -- the current node is dropped
                std::vector<MutNode> nf(out.rbegin(), out.rend()); nf.insert(nf.end(), perm.begin() + i + 1, perm.end());
                fs = nf; return ll;
            }
            MutNode node = perm[i];
            MutRes res = apply(node, ll);
            { uint64_t tot = 0; for (auto& b : res.ll) tot += b.size(); if (tot > opts.max_case_out) throw CaseOverflow("case output cap"); }
            node.score = adjust_priority(node.score, res.delta);
            out.push_back(node);
            if (!res.ll.empty() && res.ll[0] == ll[0]) { if (meta) meta->n_failed++; continue; }
            std::vector<MutNode> nf(out.rbegin(), out.rend()); nf.insert(nf.end(), perm.begin() + i + 1, perm.end());
            fs = nf;
            if (meta) { if (meta->n_used < 16) meta->used[meta->n_used] = node.name; meta->n_used++; }
            return res.ll;
        }
        fs.assign(out.rbegin(), out.rend());
        return ll;
    }
};

inline MutRes Mutations::base64_mutator(const Blocks& ll) {
    MutRes r; r.ll = ll; Chunks cs = lex(ll[0]);
    // MutasList = mutas_list(mutations([])): a fresh table (2 draws), default priorities, all 41 rows
    Opts o2 = opts;
    for (int i = 0; i < M_COUNT; i++) o2.muta_pri[i] = MUT_DEFAULT_PRI[i];
    Mutations inner2(rng, o2); inner2.build_table_draws();
    double d = -1;
    for (Chunk& c : cs) {
        if (c.type != Chunk::TEXT || c.bytes.size() <= 6) continue;
        Bin dec; if (!b64_decode(c.bytes, dec)) continue;
        int dd = rng.rand_delta();
        // mutators_mutator(MutasList, []) walks the table in FORWARD order here (no prior reversal)
        std::vector<MutNode> fs;
        for (int i = 0; i < M_COUNT; i++) { MutNode n; n.pri = MUT_DEFAULT_PRI[i]; n.name = n.fn = i; uint64_t s = rng.rand(10); n.score = (double)std::max<uint64_t>(2, s); fs.insert(fs.begin(), n); }
        Blocks in1{dec};
        Blocks nl = inner2.mux_fuzzers(fs, in1, nullptr);
        Bin nb; for (auto& b : nl) nb += b;
        c.bytes = b64_encode(nb); d += dd;
    }
    r.ll[0] = unlex(cs); r.delta = d; return r;
}

}  // namespace eo

#include "sgml.hpp"
#include "json.hpp"
