// ORACLE (test infrastructure only -- never linked into the product path).
// Restatement of erlamsa_fuse:fuse/2 (src/erlamsa_fuse.erl:47-135).
// A suffix is represented by its start position (the empty suffix [] is
// position len). A search node pairs source suffixes ("froms", positions in A)
// with target suffixes ("tos", positions in B) that share the k bytes BEFORE
// them; list orders are kept exactly as the reference builds them because
// rand_elem indexes into them.
#pragma once
#include <map>
#include "common.hpp"

namespace eo {

struct FuseNode { std::vector<uint32_t> froms, tos; };

struct Fuser {
    Rng& rng; const Bin& a; const Bin& b;
    static constexpr int64_t SEARCH_FUEL = 100000;
    static constexpr uint64_t SEARCH_STOP_IP = 8;

    // char_suffixes/1 :62-70. Each suffix [H|T] files T under H, newest first.
    // fix_empty_list: a class whose list is exactly [[]] collapses to [].
    static std::map<uint8_t, std::vector<uint32_t>> char_suffixes(const std::vector<uint32_t>& sufs, const Bin& s) {
        std::map<uint8_t, std::vector<uint32_t>> m;   // vectors hold the Erlang list REVERSED (push_back == cons)
        uint32_t len = (uint32_t)s.size();
        for (uint32_t p : sufs) {
            if (p >= len) continue;                   // the [] suffix is skipped
            uint8_t h = (uint8_t)s[p]; uint32_t t = p + 1;
            auto it = m.find(h);
            if (it == m.end()) {
                std::vector<uint32_t> v; if (t != len) v.push_back(t);   // [[]] -> []
                m[h] = v;
            } else {
                // [T | Existing]; if Existing == [] and T == [] this is [[]] again -> []
                if (it->second.empty() && t == len) continue;
                it->second.push_back(t);
            }
        }
        return m;
    }
    static std::vector<uint32_t> erl_order(const std::vector<uint32_t>& rev) { return std::vector<uint32_t>(rev.rbegin(), rev.rend()); }

    // split/2 :85-100 -- prepends the children of `node` to acc (acc holds the Erlang list reversed)
    void split(const FuseNode& node, std::vector<FuseNode>& acc_rev) {
        auto sas = char_suffixes(node.froms, a);
        auto sbs = char_suffixes(node.tos, b);
        for (auto& kv : sas) {                        // gb_trees:to_list -> ascending byte order
            if (kv.second.empty()) {                  // {_Char, []} -> [[[]], []]: froms = [[]], tos = [[]]
                FuseNode n; n.froms.push_back((uint32_t)a.size()); n.tos.push_back((uint32_t)b.size());
                acc_rev.push_back(n); continue;
            }
            auto it = sbs.find(kv.first);
            if (it == sbs.end()) continue;
            FuseNode n; n.froms = erl_order(kv.second); n.tos = erl_order(it->second);
            acc_rev.push_back(n);
        }
    }
    // any_position_pair/1 :73-77
    void any_position_pair(const std::vector<FuseNode>& nodes, uint32_t& from, uint32_t& to) {
        const FuseNode& n = nodes[rng.rand_elem_idx(nodes.size())];
        int64_t fi = rng.rand_elem_idx(n.froms.size());
        from = fi < 0 ? (uint32_t)a.size() : n.froms[fi];
        int64_t ti = rng.rand_elem_idx(n.tos.size());
        to = ti < 0 ? (uint32_t)b.size() : n.tos[ti];
    }
    // find_jump_points/2 :103-128
    void find(uint32_t& from, uint32_t& to) {
        std::vector<FuseNode> nodes(1);
        for (uint32_t i = 0; i < a.size(); i++) nodes[0].froms.push_back(i);
        for (uint32_t i = 0; i < b.size(); i++) nodes[0].tos.push_back(i);
        int64_t fuel = SEARCH_FUEL;
        for (;;) {
            if (fuel < 0) { any_position_pair(nodes, from, to); return; }
            if (rng.rand(SEARCH_STOP_IP) == 0) { any_position_pair(nodes, from, to); return; }
            std::vector<FuseNode> rev;
            for (const FuseNode& n : nodes) split(n, rev);
            if (rev.empty()) { any_position_pair(nodes, from, to); return; }
            std::vector<FuseNode> next(rev.rbegin(), rev.rend());
            fuel -= (int64_t)next.size();
            nodes.swap(next);
        }
    }
};

// fuse/2 :130-135
inline Bin fuse(Rng& rng, const Bin& a, const Bin& b) {
    if (a.empty()) return b;
    if (b.empty()) return a;
    Fuser f{rng, a, b};
    uint32_t from, to; f.find(from, to);
    return a.substr(0, from) + b.substr(to);
}

}  // namespace eo
