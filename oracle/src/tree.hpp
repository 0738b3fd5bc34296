// ORACLE (test infrastructure only -- never linked into the product path).
// Restatement of the "guessed parse-tree" mutators, src/erlamsa_mutations.erl:786-1023:
// usual_delims :793-799, grow :804-823, sublists :838-845, edit_sublist :858-869,
// edit_sublists :874-884, partial_parse :887-905, tree dup/del :930-936,
// swap one/two :940-971, stutter :974-1023.
// Results are only ever flattened (iolist_to_binary), so the edit functions emit
// bytes directly instead of building the edited nested list.
#pragma once
#include <memory>
#include "common.hpp"

namespace eo {

struct Term {
    bool is_list = false; uint8_t b = 0;
    std::vector<Term> k;
    bool operator==(const Term& o) const {
        if (is_list != o.is_list) return false;
        if (!is_list) return b == o.b;
        return k == o.k;
    }
    size_t flat_size() const { if (!is_list) return 1; size_t s = 0; for (auto& c : k) s += c.flat_size(); return s; }
};
inline Term tbyte(uint8_t b) { Term t; t.b = b; return t; }

inline int usual_delims(uint8_t c) {
    switch (c) { case 40: return 41; case 91: return 93; case 60: return 62; case 123: return 125; case 34: return 34; case 39: return 39; }
    return -1;
}

struct TreeParser {
    const uint8_t* d; size_t n;
    // grow/3: returns true when the close was found (rest position in `pos`), false when data ran out
    bool grow(size_t& pos, uint8_t close, std::vector<Term>& out) {
        while (pos < n) {
            uint8_t h = d[pos];
            if (h == close) { out.push_back(tbyte(close)); pos++; return true; }
            int nc = usual_delims(h);
            if (nc < 0) { out.push_back(tbyte(h)); pos++; continue; }
            std::vector<Term> sub; size_t p2 = pos + 1;
            bool ok = grow(p2, (uint8_t)nc, sub);
            if (!ok) {   // ran out of data: partial parse is spliced flat after H
                out.push_back(tbyte(h)); for (auto& t : sub) out.push_back(std::move(t)); pos = n; return false;
            }
            Term node; node.is_list = true; node.k.push_back(tbyte(h)); for (auto& t : sub) node.k.push_back(std::move(t));
            out.push_back(std::move(node)); pos = p2;
        }
        return false;
    }
    // partial_parse/1 :887-905 over grow/3. Written out without recursion: grow fails only by running out of data, and
    // then every enclosing grow fails too, so the opens that stay literal bytes are exactly the ones still on the
    // bracket stack at the end of the block; everything else closes. (The recursive form above splices the partial
    // lists level by level, which is quadratic on blocks with thousands of unclosed brackets.)
    std::vector<Term> parse() {
        std::vector<uint8_t> literal(n, 0);
        { std::vector<size_t> st;
          for (size_t pos = 0; pos < n; pos++) {
              uint8_t h = d[pos];
              if (!st.empty() && h == (uint8_t)usual_delims(d[st.back()])) { st.pop_back(); continue; }
              if (usual_delims(h) >= 0) st.push_back(pos);
          }
          for (size_t p : st) literal[p] = 1; }
        struct Frame { Term node; uint8_t close; };
        std::vector<Frame> st; std::vector<Term> out;
        for (size_t pos = 0; pos < n; pos++) {
            uint8_t h = d[pos];
            if (!st.empty() && h == st.back().close) {
                st.back().node.k.push_back(tbyte(h));
                Term done = std::move(st.back().node); st.pop_back();
                (st.empty() ? out : st.back().node.k).push_back(std::move(done));
                continue;
            }
            int nc = usual_delims(h);
            if (nc >= 0 && !literal[pos]) { Frame f; f.node.is_list = true; f.node.k.push_back(tbyte(h)); f.close = (uint8_t)nc; st.push_back(std::move(f)); continue; }
            (st.empty() ? out : st.back().node.k).push_back(tbyte(h));
        }
        return out;
    }
    // the clause-by-clause form, kept for the self check (eo_tree_selfcheck)
    std::vector<Term> parse_ref() {
        std::vector<Term> out; size_t pos = 0;
        while (pos < n) {
            uint8_t h = d[pos]; int nc = usual_delims(h);
            if (nc < 0) { out.push_back(tbyte(h)); pos++; continue; }
            std::vector<Term> sub; size_t p2 = pos + 1;
            bool ok = grow(p2, (uint8_t)nc, sub);
            if (!ok) { out.push_back(tbyte(h)); for (auto& t : sub) out.push_back(std::move(t)); return out; }
            Term node; node.is_list = true; node.k.push_back(tbyte(h)); for (auto& t : sub) node.k.push_back(std::move(t));
            out.push_back(std::move(node)); pos = p2;
        }
        return out;
    }
};

// sublists/1 in DISCOVERY order; the Erlang list is the reverse of this vector.
inline void sublists_disc(const std::vector<Term>& l, std::vector<const Term*>& found) {
    for (const Term& h : l) if (h.is_list) { found.push_back(&h); sublists_disc(h.k, found); }
}
inline std::vector<const Term*> sublists_erl(const std::vector<Term>& l) {
    std::vector<const Term*> f; sublists_disc(l, f); std::reverse(f.begin(), f.end()); return f;
}

inline void flatten_into(const Term& t, Bin& out) { if (!t.is_list) out.push_back((char)t.b); else for (auto& c : t.k) flatten_into(c, out); }
inline void flatten_tail(const std::vector<Term>& l, size_t from, Bin& out) { for (size_t i = from; i < l.size(); i++) flatten_into(l[i], out); }

// edit_sublist/3: op(list, index_of_match, out) emits Op([H|T]) flattened.
template <class Op>
inline void edit_sublist_emit(const std::vector<Term>& l, const Term* sub, Op& op, Bin& out) {
    for (size_t i = 0; i < l.size(); i++) {
        const Term& h = l[i];
        if (sub && h.is_list && h == *sub) { op(l, i, out); return; }
        if (h.is_list) edit_sublist_emit(h.k, sub, op, out); else out.push_back((char)h.b);
    }
}
// edit_sublists/2 with the two-entry mapping of sed_tree_swap_two
inline void edit_sublists_emit(const std::vector<Term>& l, const Term& a, const Term& b, Bin& out) {
    for (const Term& h : l) {
        if (!h.is_list) { out.push_back((char)h.b); continue; }
        // gb_trees:enter(B, ->A, enter(A, ->B, empty)): when A == B the single key maps to A
        if (h == b) { flatten_into(a, out); continue; }
        if (h == a) { flatten_into(b, out); continue; }
        edit_sublists_emit(h.k, a, b, out);
    }
}

constexpr size_t STUTTER_MEM_CAP = 64u << 20;   // stands in for the 256 MB process-memory guard, :979-981

// repeat_path/3 :974-985, emitted flattened
inline void repeat_path_emit(const Term& parent, const Term& child, uint64_t n, Bin& out) {
    // the reference stops nesting when the BEAM process passes 256 MB (non-deterministic); the restatement
    // refuses such blow-ups instead of guessing where the guard would have fired
    if (out.size() > STUTTER_MEM_CAP) throw Unsupported("tree stutter blow-up beyond the memory guard");
    if (n < 2) { flatten_into(parent, out); return; }
    auto op = [&](const std::vector<Term>& l, size_t i, Bin& o) { repeat_path_emit(parent, child, n - 1, o); flatten_tail(l, i + 1, o); };
    edit_sublist_emit(parent.k, &child, op, out);
}

}  // namespace eo
