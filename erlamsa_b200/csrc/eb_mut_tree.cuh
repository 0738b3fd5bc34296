// erlamsa_b200 -- "guessed parse tree" mutators on the device: tr2 td ts1 ts2 tr
// (reference src/erlamsa_mutations.erl:786-1023).
//
// The reference builds a nested list with a greedy, non-validating matcher over ()[]<>{}"" ''
// (grow/3 :804-823, partial_parse/1 :887-905) and then edits list cells. Three observations turn
// that into flat arrays:
//  * a node is exactly a matched (open .. close) byte range found by a stack matcher in which the
//    expected close of the innermost open delimiter is tested before "is this an opener"; openers
//    still unmatched at the end of the data create no node and their contents splice into the
//    enclosing level -- so the node set is the set of matched ranges, properly nested;
//  * sublists/1 (:838-845) enumerates nodes in DFS preorder = ascending start offset (the Erlang
//    list is that order reversed), a node's sublists are the nodes starting inside it;
//  * two nodes are structurally equal iff their byte ranges are equal (a closed node's interior
//    parse depends only on its own bytes), so `H =:= Sub` is a memcmp.
// edit_sublist/3 (:858-869) = walk levels left to right; the FIRST node of a level equal to Sub is
// edited and the rest of that level is emitted raw; earlier siblings are searched recursively.
#pragma once
#include "eb_state.cuh"

namespace eb {

struct TNode { uint32_t s, e; };   // byte range [s, e) of a node: open delimiter .. close delimiter

EB_DEV int usual_close(uint32_t c) {   // usual_delims/1 :793-799
    switch (c) { case 40: return 41; case 91: return 93; case 60: return 62; case 123: return 125; case 34: return 34; case 39: return 39; }
    return -1;
}
EB_DEV bool range_equal(const uint8_t* a, const uint8_t* b, uint32_t len) {
    uint32_t diff = 0;
    for (uint32_t i = lane_id(); i < len; i += 32) diff |= (uint32_t)(a[i] ^ b[i]);
    return !__any_sync(0xffffffffu, diff != 0);
}

// matched ranges in ascending start order; tables live in the warp's temp region.
// All lanes run the automaton and all lanes store (same value, same address): each lane only ever
// reads back its own stores, so no intra-warp synchronisation is needed inside the loop.
EB_DEV uint32_t tree_parse(CaseCtx& c, const uint8_t* p, uint32_t n, TNode** out) {
    uint32_t nopen = 0;
    for (uint32_t i = lane_id(); i < n; i += 32) nopen += usual_close(p[i]) >= 0 ? 1u : 0u;
    nopen = warp_sum(nopen);
    TNode* nodes = (TNode*)temp_alloc(c, (uint64_t)(nopen + 1) * sizeof(TNode));
    uint32_t* stack = (uint32_t*)temp_alloc(c, (uint64_t)(nopen + 1) * 4);
    *out = nodes;
    if (!nodes || !stack) return 0;
    uint32_t sp = 0, nn = 0; int topclose = -1;
    // Only the ten delimiter bytes ( ) [ ] < > { } " ' can open or close a node; everything else leaves the automaton where
    // it is. Each 512-byte window is classified 32 lanes wide (one 16-bit mask per lane), the automaton then visits the
    // candidates only -- a dependent memory access per byte was 80-110 ms per 256 KiB call (gpurun_out/tc_c4.log).
    {
        const uint32_t lead = (uint32_t)((uintptr_t)p & 15u); const uint8_t* base = p - lead; const uint32_t span = lead + n;
        for (uint32_t w0 = 0; w0 < span; w0 += 512) {
            uint32_t wofs = w0 + (uint32_t)lane_id() * 16u;
            uint4 r = make_uint4(0, 0, 0, 0);
            if (wofs < span && wofs + 16 > lead) r = *reinterpret_cast<const uint4*>(base + wofs);
            uint32_t words[4] = {r.x, r.y, r.z, r.w};
            uint32_t m = 0;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                uint32_t ch = (words[q >> 2] >> ((q & 3) * 8)) & 255u;
                bool d = ch == '(' || ch == ')' || ch == '[' || ch == ']' || ch == '<' || ch == '>' || ch == '{' || ch == '}' || ch == '"' || ch == 39;
                if (d && wofs + q >= lead && wofs + q < span) m |= 1u << q;
            }
            uint32_t lanes = __ballot_sync(0xffffffffu, m != 0);
            while (lanes) {
                int L = __ffs(lanes) - 1; lanes &= lanes - 1;
                uint32_t mm = __shfl_sync(0xffffffffu, m, L);
                uint32_t w4[4] = {__shfl_sync(0xffffffffu, r.x, L), __shfl_sync(0xffffffffu, r.y, L), __shfl_sync(0xffffffffu, r.z, L), __shfl_sync(0xffffffffu, r.w, L)};
                while (mm) {
                    int q = __ffs(mm) - 1; mm &= mm - 1;
                    uint32_t i = w0 + (uint32_t)L * 16u + (uint32_t)q - lead;
                    uint32_t h = (w4[q >> 2] >> ((q & 3) * 8)) & 255u;
                    if ((int)h == topclose) {   // the innermost open node closes (checked before "is it an opener", :807-808)
                        nodes[stack[sp - 1] & 0xffffffu].e = i + 1; sp--;
                        topclose = sp ? (int)(stack[sp - 1] >> 24) : -1;
                        continue;
                    }
                    int cl = usual_close(h);
                    if (cl >= 0) { TNode t; t.s = i; t.e = 0; nodes[nn] = t; stack[sp++] = nn | ((uint32_t)cl << 24); nn++; topclose = cl; }
                }
            }
        }
    }
    uint32_t k = 0;   // drop the openers that never closed
    for (uint32_t j = 0; j < nn; j++) { TNode t = nodes[j]; if (t.e) { nodes[k] = t; k++; } }
    __syncwarp();
    return k;
}

// edit_sublist/3 traversal: indices of the nodes (inside [lo, hi) of the node table, whose enclosing
// level ends at `outer_end`) that get edited when looking for ranges equal to target [ts, te).
// Results go to `hits` (temp). With two targets (ts2) a hit is any node equal to A or B and `which`
// records 0 for A / 1 for B; edited nodes are never descended into.
EB_DEV uint32_t tree_find_edits(CaseCtx& c, const uint8_t* p, const TNode* nodes, uint32_t lo, uint32_t hi, uint32_t outer_end,
                                const TNode* ta, const TNode* tb, bool all_occurrences, uint32_t* hits, uint8_t* which, uint32_t* ends) {
    uint32_t nh = 0, depth = 0, skip_until = 0;
    for (uint32_t i = lo; i < hi; i++) {
        TNode nd = nodes[i];
        while (depth > 0 && ends[depth - 1] <= nd.s) depth--;
        uint32_t parent_end = depth > 0 ? ends[depth - 1] : outer_end;
        if (nd.s < skip_until) continue;
        uint32_t len = nd.e - nd.s;
        int m = -1;
        if (tb && len == tb->e - tb->s && range_equal(p + nd.s, p + tb->s, len)) m = 1;   // gb_trees:enter(B, ..) wins when A == B
        else if (len == ta->e - ta->s && range_equal(p + nd.s, p + ta->s, len)) m = 0;
        if (m >= 0) {
            hits[nh] = i; if (which) which[nh] = (uint8_t)m; nh++;
            skip_until = all_occurrences ? nd.e : parent_end;   // ts2 keeps scanning the level; edit_sublist stops it
            continue;
        }
        ends[depth++] = nd.e;
    }
    __syncwarp();
    return nh;
}

// output builder: segments while they fit, otherwise bytes in scratch
struct TreeOut { CaseCtx* c; bool overflow; };
EB_DEV void tree_emit(CaseCtx& c, Seg s) { t_push(c.ws, s); }

// collapse the candidate script into one scratch buffer when it has grown too long
EB_DEV void tree_compact(CaseCtx& c) {
    WarpState* ws = c.ws;
    if (ws->ntseg < MAX_VSEG - 4) return;
    if (ws->tlen > c.bp->max_case_out) { ws->status = CASE_OVERFLOW; ws->reason = 5; return; }
    uint8_t* buf = scratch_alloc(c, ws->tlen);
    if (!buf) return;
    segs_write(ws->tseg, ws->ntseg, buf);
    uint32_t len = ws->tlen;
    t_reset(ws); t_push(ws, seg_copy(buf, len));
}

// warp bitonic sort of (key, payload) pairs, m a power of two
EB_DEV void warp_bitonic_kv(uint64_t* K, uint32_t* V, uint32_t m) {
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane_id(); i < m; i += 32) {
                uint32_t x = i ^ j;
                if (x > i) {
                    uint64_t a = K[i], b = K[x]; uint32_t va = V[i], vb = V[x];
                    bool up = (i & k) == 0;
                    bool gt = (a > b) || (a == b && va > vb);
                    if (gt == up) { K[i] = b; K[x] = a; V[i] = vb; V[x] = va; }
                }
            }
            __syncwarp();
        }
}

EB_DEV void mut_tree(CaseCtx& c, int id, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0;
    if (mem_binarish(p, n)) { r.kind = RES_SAME; r.delta = -1; return; }
    TNode* nodes; uint32_t N = tree_parse(c, p, n, &nodes);
    if (ws->status != CASE_OK) { r.kind = RES_SAME; r.delta = 0; return; }
    uint32_t* hits = (uint32_t*)temp_alloc(c, (uint64_t)(N + 1) * 4);
    uint8_t* which = (uint8_t*)temp_alloc(c, (uint64_t)N + 1);
    uint32_t* ends = (uint32_t*)temp_alloc(c, (uint64_t)(N + 1) * 4);
    if (!hits || !which || !ends) { r.kind = RES_SAME; r.delta = 0; return; }
    t_reset(ws);
    r.delta = 1; r.kind = RES_SEGS;

    if (id == M_TR2 || id == M_TD) {   // sed_tree_op :917-936 with pick_sublist :848-855
        if (N == 0) { t_push(ws, seg_copy(p, n)); return; }          // Sub = false: nothing matches
        uint32_t ri = (uint32_t)g.rand_elem_idx(N);
        TNode sub = nodes[N - 1 - ri];
        uint32_t nh = tree_find_edits(c, p, nodes, 0, N, n, &sub, nullptr, false, hits, nullptr, ends);
        uint32_t pos = 0;
        for (uint32_t k = 0; k < nh && ws->status == CASE_OK; k++) {
            TNode m = nodes[hits[k]];
            if (id == M_TR2) { t_push(ws, seg_copy(p + pos, m.e - pos)); pos = m.s; }     // [H, H | T]
            else { t_push(ws, seg_copy(p + pos, m.s - pos)); pos = m.e; }                 // T
            tree_compact(c);
        }
        t_push(ws, seg_copy(p + pos, n - pos));
        return;
    }
    if (id == M_TS1 || id == M_TS2) {   // construct_sed_tree_swap :956-971
        if (N < 2) { r.kind = RES_SAME; r.delta = -1; return; }
        // reservoir_sample(Subs, 2) over the Erlang list (descending start): element I is node N-I
        uint32_t ra = N - 1, rb = N - 2;
        for (uint32_t I = 3; I <= N; I++) { uint64_t J = g.erand(I); if (J == 1) ra = N - I; else if (J == 2) rb = N - I; }
        if (id == M_TS1) {
            if (g.rand(2) == 1) { uint32_t t = ra; ra = rb; rb = t; }   // random_permutation of the pair
            TNode a = nodes[ra], b = nodes[rb];
            uint32_t nh = tree_find_edits(c, p, nodes, 0, N, n, &a, nullptr, false, hits, nullptr, ends);
            uint32_t pos = 0;
            for (uint32_t k = 0; k < nh && ws->status == CASE_OK; k++) {
                TNode m = nodes[hits[k]];
                t_push(ws, seg_copy(p + pos, m.s - pos)); t_push(ws, seg_copy(p + b.s, b.e - b.s)); pos = m.e;
                tree_compact(c);
            }
            t_push(ws, seg_copy(p + pos, n - pos));
        } else {
            TNode a = nodes[ra], b = nodes[rb];
            uint32_t nh = tree_find_edits(c, p, nodes, 0, N, n, &a, &b, true, hits, which, ends);
            uint32_t pos = 0;
            for (uint32_t k = 0; k < nh && ws->status == CASE_OK; k++) {
                TNode m = nodes[hits[k]]; TNode rep = which[k] ? a : b;   // B -> A, A -> B
                t_push(ws, seg_copy(p + pos, m.s - pos)); t_push(ws, seg_copy(p + rep.s, rep.e - rep.s)); pos = m.e;
                tree_compact(c);
            }
            t_push(ws, seg_copy(p + pos, n - pos));
        }
        return;
    }
    // ---- M_TR sed_tree_stutter :1004-1023
    uint32_t pi = 0xffffffffu, ci = 0;
    if (N == 2) {
        bool sw = g.rand(2) == 1;
        // list order is descending start: [node1, node0]; swapped -> [node0, node1]
        uint32_t ord[2]; ord[0] = sw ? 0u : 1u; ord[1] = sw ? 1u : 0u;
        for (int q = 0; q < 2 && pi == 0xffffffffu; q++) {
            uint32_t i = ord[q]; uint32_t cnt = 0;
            for (uint32_t j = i + 1; j < N && nodes[j].s < nodes[i].e; j++) cnt++;
            if (cnt) { uint32_t ri = (uint32_t)g.rand_elem_idx(cnt); pi = i; ci = i + cnt - ri; }
        }
    } else if (N > 0) {
        uint32_t m = 1; while (m < N) m <<= 1;
        uint64_t* K = (uint64_t*)temp_alloc(c, (uint64_t)m * 8); uint32_t* V = (uint32_t*)temp_alloc(c, (uint64_t)m * 4);
        if (!K || !V) { r.kind = RES_SAME; r.delta = 0; return; }
        // keys are drawn in list order (descending start)
        for (uint32_t q = 0; q < N; q++) { double u = g.uniform(); K[q] = (uint64_t)__double_as_longlong(u); V[q] = N - 1 - q; }
        for (uint32_t q = N + lane_id(); q < m; q += 32) { K[q] = ~0ull; V[q] = 0xffffffffu; }
        __syncwarp();
        warp_bitonic_kv(K, V, m);
        for (uint32_t q = 0; q < N && pi == 0xffffffffu; q++) {   // choose_stutr_nodes :995-1002
            uint32_t i = V[q]; uint32_t cnt = 0;
            for (uint32_t j = i + 1; j < N && nodes[j].s < nodes[i].e; j++) cnt++;
            if (cnt) { uint32_t ri = (uint32_t)g.rand_elem_idx(cnt); pi = i; ci = i + cnt - ri; }
        }
    }
    uint64_t reps = g.rand_log_small(10);
    if (pi == 0xffffffffu) { r.kind = RES_SAME; r.delta = -1; return; }
    TNode P = nodes[pi], C = nodes[ci];
    // matches of Child inside Parent (Parent's own level ends at P.e) and in the whole block
    uint32_t pcnt = 0; for (uint32_t j = pi + 1; j < N && nodes[j].s < P.e; j++) pcnt++;
    uint32_t* ihits = (uint32_t*)temp_alloc(c, (uint64_t)(pcnt + 1) * 4);
    if (!ihits) { r.kind = RES_SAME; r.delta = 0; return; }
    uint32_t mi = tree_find_edits(c, p, nodes, pi + 1, pi + 1 + pcnt, P.e, &C, nullptr, false, ihits, nullptr, ends);
    uint32_t nt = tree_find_edits(c, p, nodes, 0, N, n, &C, nullptr, false, hits, nullptr, ends);
    uint32_t plen = P.e - P.s, clen = C.e - C.s;
    // repeat_path/3 :974-985: R(k) = Parent for k < 2, else Parent with every matched Child replaced by R(k-1)
    uint64_t rlen = plen;
    for (uint64_t k = 2; k <= reps; k++) { rlen = plen + (uint64_t)mi * (rlen - clen); if (rlen > c.bp->max_case_out) break; }
    if (rlen * nt > c.bp->max_case_out) { ws->status = CASE_OVERFLOW; ws->reason = 5; r.kind = RES_SAME; return; }
    const uint8_t* Rp = p + P.s; uint32_t Rl = plen;          // materialised R(reps) unless mi == 1
    bool closed_form = (mi == 1) || reps < 2;
    TNode M0 = nodes[ihits[0]];
    if (!closed_form) {
        for (uint64_t k = 2; k <= reps; k++) {
            uint64_t nl = plen + (uint64_t)mi * ((uint64_t)Rl - clen);
            uint8_t* buf = scratch_alloc(c, nl);
            if (!buf) { r.kind = RES_SAME; return; }
            uint32_t pos = P.s; uint8_t* w = buf;
            for (uint32_t q = 0; q < mi; q++) { TNode m = nodes[ihits[q]]; warp_copy(w, p + pos, m.s - pos); w += m.s - pos; warp_copy(w, Rp, Rl); w += Rl; pos = m.e; }
            warp_copy(w, p + pos, P.e - pos);
            __syncwarp();
            Rp = buf; Rl = (uint32_t)nl;
        }
    }
    uint32_t pos = 0;
    for (uint32_t k = 0; k < nt && ws->status == CASE_OK; k++) {
        TNode m = nodes[hits[k]];
        t_push(ws, seg_copy(p + pos, m.s - pos));
        if (closed_form && reps >= 2) {   // one match per level: R(k) = pre^(k-1) . Parent . post^(k-1)
            uint32_t pre = M0.s - P.s, post = P.e - M0.e;
            if (pre) t_push(ws, seg_repeat(p + P.s, pre, (uint32_t)(pre * (reps - 1))));
            t_push(ws, seg_copy(p + P.s, plen));
            if (post) t_push(ws, seg_repeat(p + M0.e, post, (uint32_t)(post * (reps - 1))));
        } else t_push(ws, seg_copy(Rp, Rl));
        pos = m.e;
        tree_compact(c);
    }
    t_push(ws, seg_copy(p + pos, n - pos));
}

}  // namespace eb
