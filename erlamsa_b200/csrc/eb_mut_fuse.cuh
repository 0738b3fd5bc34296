// erlamsa_b200 -- the fuse mutators ft fn fo on the device (reference src/erlamsa_mutations.erl:386-427
// over src/erlamsa_fuse.erl:47-135).
//
// fuse(A, B) = prefix of A up to i ++ suffix of B from j, where the k bytes before i and before j are
// equal; k grows level by level (each level stops with probability 1/8) and the pair (i, j) is drawn
// by index from ORDERED lists, so exact parity needs the reference's list orders, not just its sets:
//   * a node = (source suffixes, target suffixes) sharing the bytes consumed so far; suffix = start
//     position, the empty suffix = len;
//   * char_suffixes/1 (:62-70) files each suffix's tail under its first byte, newest first (so every
//     level reverses the order inside a class) -- a stable counting sort written back to front;
//     the quirk fix_empty_list ([[]] -> []) drops the empty tail when it is the first of its class;
//   * split/2 (:85-100) walks the classes in ascending byte order and PREPENDS the children, so a
//     level's node list is the emission order reversed.
// Positions are kept in ping-pong arrays per side in the warp's temp region; the automaton is run
// warp-uniformly (all lanes compute and store the same values, each lane reads back its own stores).
#pragma once
#include "eb_state.cuh"

namespace eb {

struct FNode { uint32_t fo, fc, to, tc; };

struct FuseSide {
    const uint8_t* s; uint32_t len;
    uint32_t* cnt;     // [256] per-class element count; all zero between nodes
    uint32_t* start;   // [256] class start inside the next-level array
    uint32_t bits[8];  // classes that exist for the node being split (possibly with an empty list)
    uint32_t size[8];  // scratch: unused
};

// One node, one side: classify the suffix list src[0,k) by first byte into dst[base ...), classes in ascending
// byte order, each class newest-first. After the call: bits = existing classes, start[c]/csize[c] describe class c
// (cnt is consumed back to zero by the placement pass). Thirty-two suffixes per step: lanes holding the same byte
// find each other with match.any, the lowest of them owns the class counter for that step.
EB_DEV uint32_t fuse_classify(FuseSide& sd, const uint32_t* src, uint32_t k, uint32_t* dst, uint32_t base, uint32_t* csize) {
    const int l = lane_id(); const uint32_t lt = (1u << l) - 1u;
    for (int i = 0; i < 8; i++) sd.bits[i] = 0;
    // the suffix holding only the block's last byte is dropped when it is the FIRST of its class ([[]] -> [], :68-70)
    uint32_t sq = 0xffffffffu, sch = 0;
    for (uint32_t q0 = 0; q0 < k; q0 += 32) {
        uint32_t q = q0 + l; uint32_t p = q < k ? src[q] : 0xffffffffu;
        uint32_t hit = __ballot_sync(0xffffffffu, q < k && p + 1 == sd.len);
        if (hit) { sq = q0 + (uint32_t)__ffs(hit) - 1; sch = sd.s[sd.len - 1]; break; }
    }
    bool special_dropped = false;
    if (sq != 0xffffffffu) {
        uint32_t before = 0;
        for (uint32_t q0 = 0; q0 < sq; q0 += 32) {
            uint32_t q = q0 + l; uint32_t p = q < sq ? src[q] : 0xffffffffu;
            before |= __ballot_sync(0xffffffffu, q < sq && p < sd.len && sd.s[p] == sch);
            if (before) break;
        }
        special_dropped = before == 0;
    }
    for (uint32_t q0 = 0; q0 < k; q0 += 32) {                          // count
        uint32_t q = q0 + l; uint32_t p = q < k ? src[q] : 0xffffffffu;
        bool live = q < k && p < sd.len;                                // the [] suffix contributes nothing (:68)
        uint32_t ch = live ? sd.s[p] : 0x100u + (uint32_t)l;
        for (int w = 0; w < 8; w++) sd.bits[w] |= __reduce_or_sync(0xffffffffu, (live && (int)(ch >> 5) == w) ? 1u << (ch & 31) : 0u);
        bool counted = live && !(special_dropped && q == sq);
        uint32_t key = counted ? ch : 0x100u + (uint32_t)l;
        uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (counted && (peers & lt) == 0) sd.cnt[ch] = sd.cnt[ch] + (uint32_t)__popc(peers);
        __syncwarp();
    }
    uint32_t run = base;
    for (int w = 0; w < 8; w++) {                                       // class starts, ascending byte
        uint32_t m = sd.bits[w];
        while (m) {
            uint32_t b = (uint32_t)__ffs(m) - 1; m &= m - 1; uint32_t ch = (uint32_t)w * 32 + b;
            uint32_t cc = sd.cnt[ch];
            if (l == 0) { sd.start[ch] = run; csize[ch] = cc; }
            run += cc;
        }
    }
    __syncwarp();
    for (uint32_t q0 = 0; q0 < k; q0 += 32) {                          // place: first arrival lands last in its class
        uint32_t q = q0 + l; uint32_t p = q < k ? src[q] : 0xffffffffu;
        bool placed = q < k && p < sd.len && !(special_dropped && q == sq);
        uint32_t ch = placed ? sd.s[p] : 0;
        uint32_t key = placed ? ch : 0x100u + (uint32_t)l;
        uint32_t peers = __match_any_sync(0xffffffffu, key);
        uint32_t left = placed ? sd.cnt[ch] : 0;
        __syncwarp();
        if (placed) {
            dst[sd.start[ch] + left - 1 - (uint32_t)__popc(peers & lt)] = p + 1;
            if ((peers & lt) == 0) sd.cnt[ch] = left - (uint32_t)__popc(peers);
        }
        __syncwarp();
    }
    return run - base;
}

// ---- one small node per lane (see fuse_device, path 1)
constexpr uint32_t FUSE_K = 24;
constexpr uint32_t FUSE_NONE = 0xffffffffu;
struct SmallSide {
    uint32_t pos[FUSE_K];      // suffix positions in list order
    uint8_t ch[FUSE_K];        // first byte of each live suffix
    uint8_t live[FUSE_K];      // 0: the [] suffix (contributes nothing), 1: placed, 2: the dropped special
    uint32_t n;
    uint8_t cbyte[FUSE_K];     // classes in ascending byte order (a class whose only member was dropped has size 0)
    uint8_t csize[FUSE_K];
    uint8_t cstart[FUSE_K];    // start of the class inside this node's output
    uint32_t ncls, nplaced;
};
__device__ __forceinline__ void small_side_load(SmallSide& s, const uint32_t* src, uint32_t k, const uint8_t* data, uint32_t len) {
    s.n = k;
    int special = -1;
    for (uint32_t i = 0; i < k; i++) {
        uint32_t p = src[i]; s.pos[i] = p;
        bool lv = p < len;
        s.live[i] = lv ? 1 : 0; s.ch[i] = lv ? data[p] : 0;
        if (lv && p + 1 == len && special < 0) special = (int)i;
    }
    // the suffix holding only the block's last byte is dropped when it is the FIRST of its class ([[]] -> [], :68-70)
    if (special >= 0) {
        bool before = false;
        for (int i = 0; i < special; i++) before |= s.live[i] && s.ch[i] == s.ch[special];
        if (!before) s.live[special] = 2;
    }
}
__device__ __forceinline__ void small_side_classes(SmallSide& s) {
    s.ncls = 0; s.nplaced = 0;
    int last = -1;
    for (;;) {
        int best = 256;
        for (uint32_t i = 0; i < s.n; i++) if (s.live[i] && (int)s.ch[i] > last && (int)s.ch[i] < best) best = s.ch[i];
        if (best == 256) break;
        uint32_t cnt = 0;
        for (uint32_t i = 0; i < s.n; i++) cnt += (s.live[i] == 1 && s.ch[i] == best) ? 1u : 0u;
        s.cbyte[s.ncls] = (uint8_t)best; s.csize[s.ncls] = (uint8_t)cnt; s.cstart[s.ncls] = (uint8_t)s.nplaced;
        s.ncls++; s.nplaced += cnt; last = best;
    }
}
__device__ __forceinline__ int small_side_find(const SmallSide& s, uint32_t byte) {
    for (uint32_t i = 0; i < s.ncls; i++) if (s.cbyte[i] == byte) return (int)i;
    return -1;
}
// classes ascending, inside a class the LAST arrival first (char_suffixes prepends)
__device__ __forceinline__ void small_side_write(const SmallSide& s, uint32_t* dst) {
    uint32_t w = 0;
    for (uint32_t ci = 0; ci < s.ncls; ci++)
        for (int i = (int)s.n - 1; i >= 0; i--) if (s.live[i] == 1 && s.ch[i] == s.cbyte[ci]) dst[w++] = s.pos[i] + 1;
}

// find_jump_points/2 :103-128 + any_position_pair/1 :73-77
EB_DEV bool fuse_device(CaseCtx& c, const uint8_t* a, uint32_t na, const uint8_t* b, uint32_t nb, uint32_t& from, uint32_t& to) {
    Rng& g = c.rng;
    uint64_t fcap = (uint64_t)na + 80, tcap = (uint64_t)nb + 80, ncap = (uint64_t)(na < nb ? na : nb) + 80;
    uint32_t* F[2]; uint32_t* T[2]; FNode* ND[2];
    for (int i = 0; i < 2; i++) { F[i] = (uint32_t*)temp_alloc(c, fcap * 4); T[i] = (uint32_t*)temp_alloc(c, tcap * 4); ND[i] = (FNode*)temp_alloc(c, ncap * sizeof(FNode)); }
    uint32_t* tabs = (uint32_t*)temp_alloc(c, 6 * 256 * 4);
    if (!F[0] || !F[1] || !T[0] || !T[1] || !ND[0] || !ND[1] || !tabs) return false;
    FuseSide sa, sb; sa.s = a; sa.len = na; sa.cnt = tabs; sa.start = tabs + 256; sb.s = b; sb.len = nb; sb.cnt = tabs + 512; sb.start = tabs + 768;
    uint32_t* sizeA = tabs + 1024; uint32_t* sizeB = tabs + 1280;
    for (uint32_t i = lane_id(); i < 256; i += 32) { sa.cnt[i] = 0; sb.cnt[i] = 0; }
    for (uint32_t i = lane_id(); i < na; i += 32) F[0][i] = i;
    for (uint32_t i = lane_id(); i < nb; i += 32) T[0][i] = i;
    __syncwarp();
    int cur = 0; uint32_t ncur = 1;
    { FNode r0; r0.fo = 0; r0.fc = na; r0.to = 0; r0.tc = nb; ND[0][0] = r0; }
    int64_t fuel = 100000;
    bool compact = false;
    for (;;) {
        bool stop = fuel < 0;
        if (!stop) stop = g.rand(8) == 0;
        uint32_t nnext = 0;
        if (!stop && compact) {
            // ---- every node of this level holds at most one suffix per side (and so will all their descendants): the level is
            // two flat arrays, PA[k] / PB[k] = the node's source / target position (NONE = no suffix), kept in the F / T ping-pong
            // buffers; a step reads 32 nodes with two coalesced loads and two byte gathers, nothing depends on the previous step
            // but the output offset. Same rules as the general path written out for one-element lists.
            const int nx = cur ^ 1;
            const uint32_t* PA = F[cur]; const uint32_t* PB = T[cur]; uint32_t* QA = F[nx]; uint32_t* QB = T[nx];
            const int l = lane_id(); const uint32_t ltm = (1u << l) - 1u;
            for (uint32_t e = ncur; e > 0;) {
                uint32_t take = e < 32 ? e : 32;
                bool act = (uint32_t)l < take;
                uint32_t pa = FUSE_NONE, pb = FUSE_NONE;
                if (act) { pa = PA[e - 1 - (uint32_t)l]; pb = PB[e - 1 - (uint32_t)l]; }
                bool ha = pa != FUSE_NONE && pa < na, hb = pb != FUSE_NONE && pb < nb;
                uint32_t cha = ha ? a[pa] : 0x100u, chb = hb ? b[pb] : 0x200u;
                bool a_drop = ha && pa + 1 == na, b_drop = hb && pb + 1 == nb;
                bool child = false; uint32_t ca = FUSE_NONE, cb = FUSE_NONE;
                if (a_drop) { child = true; ca = na; cb = nb; }                                   // {_Char, []} -> [[[]], []] (:91-93)
                else if (ha && cha == chb) { child = true; ca = pa + 1; cb = b_drop ? FUSE_NONE : pb + 1; }
                uint32_t cm = __ballot_sync(0xffffffffu, child);
                if (child) { uint32_t at = nnext + (uint32_t)__popc(cm & ltm); QA[at] = ca; QB[at] = cb; }
                nnext += (uint32_t)__popc(cm);
                e -= take;
            }
            __syncwarp();
            if (nnext == 0) stop = true;
            else { fuel -= (int64_t)nnext; cur = nx; ncur = nnext; continue; }
        }
        if (!stop) {
            const int nx = cur ^ 1; uint32_t fo = 0, to = 0;
            uint32_t maxsz = 0;                                              // largest suffix list among the children of this level
            // nodes are stored in emission order; the reference's list is that order reversed
            uint32_t e = ncur;
            const int l = lane_id(); const uint32_t ltm = (1u << l) - 1u;
            while (e > 0) {
                // (1) a run of SMALL nodes (at most FUSE_K suffixes on each side -- what a level degenerates to after the
                //     first split or two): one node per LANE, the same classification rules written out serially per lane
                //     (classes ascending by byte, newest first inside a class, the [[]] -> [] drop, the {Char, []} quirk).
                //     Measured before this path existed: ft on a 4 KiB random block 24 ms, nearly all of it in the ~256
                //     sixteen-by-sixteen nodes of the second level taken one node per warp step (profiles/tc_c2_r2a.txt).
                bool valid = (uint32_t)l < e;
                FNode nd; nd.fo = nd.to = 0; nd.fc = nd.tc = FUSE_K + 1;
                if (valid) nd = ND[cur][e - 1 - (uint32_t)l];
                uint32_t sm = __ballot_sync(0xffffffffu, valid && nd.fc <= FUSE_K && nd.tc <= FUSE_K);
                uint32_t runlen = sm == 0xffffffffu ? 32u : (uint32_t)__ffs(~sm) - 1u;
                if (runlen > 0) {
                    bool act = (uint32_t)l < runlen;
                    SmallSide A, B;
                    uint32_t nquirk = 0, nchild = 0;
                    if (act) {
                        small_side_load(A, F[cur] + nd.fo, nd.fc, a, na);
                        small_side_load(B, T[cur] + nd.to, nd.tc, b, nb);
                        small_side_classes(A); small_side_classes(B);
                        for (uint32_t ci = 0; ci < A.ncls; ci++) {
                            if (A.csize[ci] == 0) { nquirk++; nchild++; }
                            else if (small_side_find(B, A.cbyte[ci]) >= 0) nchild++;
                        }
                    }
                    uint32_t fadd = act ? A.nplaced + nquirk : 0u, tadd = act ? B.nplaced + nquirk : 0u, cadd = act ? nchild : 0u;
                    uint32_t fpre = fadd, tpre = tadd, cpre = cadd;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        uint32_t x = __shfl_up_sync(0xffffffffu, fpre, o), y = __shfl_up_sync(0xffffffffu, tpre, o), z = __shfl_up_sync(0xffffffffu, cpre, o);
                        if (l >= o) { fpre += x; tpre += y; cpre += z; }
                    }
                    uint32_t ftot = __shfl_sync(0xffffffffu, fpre, 31), ttot = __shfl_sync(0xffffffffu, tpre, 31), ctot = __shfl_sync(0xffffffffu, cpre, 31);
                    if ((uint64_t)fo + ftot > fcap || (uint64_t)to + ttot > tcap || (uint64_t)nnext + ctot > ncap) { c.ws->status = CASE_OVERFLOW; c.ws->reason = 8; return false; }
                    uint32_t lane_max = 0;
                    if (act) {
                        uint32_t f0 = fo + fpre - fadd, t0 = to + tpre - tadd, c0 = nnext + cpre - cadd;
                        small_side_write(A, F[nx] + f0);
                        small_side_write(B, T[nx] + t0);
                        uint32_t fq = f0 + A.nplaced, tq = t0 + B.nplaced;
                        for (uint32_t ci = 0; ci < A.ncls; ci++) {
                            FNode ch_n;
                            if (A.csize[ci] == 0) {        // {_Char, []} -> [[[]], []]: the two empty suffixes, whatever B holds (:91-93)
                                F[nx][fq] = na; T[nx][tq] = nb;
                                ch_n.fo = fq; ch_n.fc = 1; ch_n.to = tq; ch_n.tc = 1; fq++; tq++;
                                ND[nx][c0++] = ch_n; continue;
                            }
                            int bi = small_side_find(B, A.cbyte[ci]);
                            if (bi < 0) continue;                                   // notfound
                            ch_n.fo = f0 + A.cstart[ci]; ch_n.fc = A.csize[ci]; ch_n.to = t0 + B.cstart[bi]; ch_n.tc = B.csize[bi];
                            ND[nx][c0++] = ch_n;
                            uint32_t mx = ch_n.fc > ch_n.tc ? ch_n.fc : ch_n.tc; if (mx > lane_max) lane_max = mx;
                        }
                    }
                    maxsz = max(maxsz, __reduce_max_sync(0xffffffffu, lane_max));
                    fo += ftot; to += ttot; nnext += ctot; e -= runlen;
                    __syncwarp();
                    continue;
                }
                // (2) a node with longer lists: warp-parallel classification, children walked class by class
                e--;
                nd = ND[cur][e];
                uint32_t fa = fuse_classify(sa, F[cur] + nd.fo, nd.fc, F[nx], fo, sizeA);
                uint32_t tb = fuse_classify(sb, T[cur] + nd.to, nd.tc, T[nx], to, sizeB);
                for (int w = 0; w < 8; w++) {
                    uint32_t m = sa.bits[w];
                    while (m) {
                        uint32_t bit = (uint32_t)__ffs(m) - 1; m &= m - 1; uint32_t ch = (uint32_t)w * 32 + bit;
                        FNode ch_n;
                        if (sizeA[ch] == 0) {       // {_Char, []} -> [[[]], []]: the two empty suffixes, whatever B holds (:91-93)
                            if (fo + fa + 1 > fcap || to + tb + 1 > tcap || nnext >= ncap) { c.ws->status = CASE_OVERFLOW; c.ws->reason = 8; return false; }
                            if (l == 0) { F[nx][fo + fa] = na; T[nx][to + tb] = nb; }
                            ch_n.fo = fo + fa; ch_n.fc = 1; ch_n.to = to + tb; ch_n.tc = 1; fa++; tb++;
                            if (l == 0) ND[nx][nnext] = ch_n;
                            nnext++; continue;
                        }
                        if (!((sb.bits[w] >> bit) & 1u)) continue;                     // notfound
                        if (nnext >= ncap) { c.ws->status = CASE_OVERFLOW; c.ws->reason = 8; return false; }
                        ch_n.fo = sa.start[ch]; ch_n.fc = sizeA[ch]; ch_n.to = sb.start[ch]; ch_n.tc = sizeB[ch];
                        if (l == 0) ND[nx][nnext] = ch_n;
                        nnext++;
                        maxsz = max(maxsz, max(ch_n.fc, ch_n.tc));
                    }
                }
                fo += fa; to += tb;
                __syncwarp();
            }
            if (nnext == 0) stop = true;
            else {
                fuel -= (int64_t)nnext; cur = nx; ncur = nnext;
                if (maxsz <= 1) {      // from here on every node is at most 1 x 1: flatten the level into position arrays
                    const int l2 = lane_id(); const int ot = cur ^ 1;
                    for (uint32_t k = (uint32_t)l2; k < ncur; k += 32) {
                        FNode nd = ND[cur][k];
                        F[ot][k] = nd.fc ? F[cur][nd.fo] : FUSE_NONE;
                        T[ot][k] = nd.tc ? T[cur][nd.to] : FUSE_NONE;
                    }
                    __syncwarp();
                    cur = ot; compact = true;
                }
                continue;
            }
        }
        // any_position_pair(Nodes)
        if (compact) {
            uint32_t r = (uint32_t)g.rand_elem_idx(ncur);
            uint32_t pa = F[cur][ncur - 1 - r], pb = T[cur][ncur - 1 - r];
            int64_t fi = g.rand_elem_idx(pa != FUSE_NONE ? 1 : 0);
            from = fi < 0 ? na : pa;
            int64_t ti = g.rand_elem_idx(pb != FUSE_NONE ? 1 : 0);
            to = ti < 0 ? nb : pb;
            return true;
        }
        uint32_t r = (uint32_t)g.rand_elem_idx(ncur);
        FNode nd = ND[cur][ncur - 1 - r];
        int64_t fi = g.rand_elem_idx(nd.fc);
        from = fi < 0 ? na : F[cur][nd.fo + (uint32_t)fi];
        int64_t ti = g.rand_elem_idx(nd.tc);
        to = ti < 0 ? nb : T[cur][nd.to + (uint32_t)ti];
        return true;
    }
}

// fuse/2 :130-135 as an edit script: pushes A[0,i) and B[j,nb) to tseg
EB_DEV bool fuse_push(CaseCtx& c, const uint8_t* a, uint32_t na, const uint8_t* b, uint32_t nb) {
    WarpState* ws = c.ws;
    if (na == 0) { t_push(ws, seg_copy(b, nb)); return true; }
    if (nb == 0) { t_push(ws, seg_copy(a, na)); return true; }
    uint32_t from, to;
    uint64_t mark = c.temp_used;
    bool ok = fuse_device(c, a, na, b, nb, from, to);
    c.temp_used = mark;                       // the level tables die here
    if (!ok) return false;
    t_push(ws, seg_copy(a, from)); t_push(ws, seg_copy(b + to, nb - to));
    return true;
}
// materialise the candidate script into scratch (the second fuse of fn/fo reads the first one's bytes)
EB_DEV const uint8_t* tseg_to_scratch(CaseCtx& c, uint32_t& len) {
    WarpState* ws = c.ws;
    len = ws->tlen;
    uint8_t* buf = scratch_alloc(c, len);
    if (!buf) return nullptr;
    segs_write(ws->tseg, ws->ntseg, buf);
    return buf;
}

EB_DEV void mut_fuse(CaseCtx& c, int id, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0; r.kind = RES_SEGS;
    t_reset(ws);
    if (id == M_FT) {   // sed_fuse_this :386-390
        if (!fuse_push(c, p, n, p, n)) { r.kind = RES_SAME; r.delta = 0; return; }
        r.delta = g.rand_delta(); return;
    }
    uint32_t h1 = n / 2;   // erlamsa_utils:halve/1: the first half is floor(n/2)
    if (id == M_FN) {   // sed_fuse_next :393-402
        const uint8_t* bp = c.has_next ? c.next_p : p; uint32_t bn = c.has_next ? c.next_n : n;
        if (!fuse_push(c, p, h1, bp, bn)) { r.kind = RES_SAME; r.delta = 0; return; }
        uint32_t abl_n; const uint8_t* abl = tseg_to_scratch(c, abl_n);
        if (!abl) { r.kind = RES_SAME; r.delta = 0; return; }
        t_reset(ws);
        if (!fuse_push(c, abl, abl_n, p + h1, n - h1)) { r.kind = RES_SAME; r.delta = 0; return; }
        r.delta = g.rand_delta(); r.rechunk = 1; r.consumed_next = c.has_next ? 1 : 0;
        return;
    }
    // sed_fuse_old :404-427 -- remember/1 closure: the first call remembers H itself
    // (with a donor pool -- config C5 -- the first remembered block is a window of another seed, possibly of another GPU's
    // shard; chosen from the thread seed so that the case's own draws stay the reference's)
    if (!ws->fo_has) {
        ws->fo_has = 1; ws->fo_p = p; ws->fo_n = n;
        if (c.bp->donor_pool && c.bp->n_donors) { uint64_t d = ws->donor % c.bp->n_donors; ws->fo_p = c.bp->donor_pool + d * c.bp->donor_stride; ws->fo_n = c.bp->donor_len[d]; }
    }
    const uint8_t* op = ws->fo_p; uint32_t on = ws->fo_n; uint32_t o1 = on / 2;
    if (!fuse_push(c, p, h1, op, o1)) { r.kind = RES_SAME; r.delta = 0; return; }          // a -> o
    uint32_t an; const uint8_t* ab = tseg_to_scratch(c, an);
    if (!ab) { r.kind = RES_SAME; r.delta = 0; return; }
    t_reset(ws);
    if (!fuse_push(c, op + o1, on - o1, p + h1, n - h1)) { r.kind = RES_SAME; r.delta = 0; return; }   // o -> a
    uint32_t bn2; const uint8_t* bb = tseg_to_scratch(c, bn2);
    if (!bb) { r.kind = RES_SAME; r.delta = 0; return; }
    uint64_t swap = g.rand(3);
    r.delta = g.rand_delta();
    if (swap == 0) { ws->fo_p = p; ws->fo_n = n; }
    // flush_bvecs(A, flush_bvecs(B, T)): two re-chunked regions -> handed back as block runs
    ws->rrun_n = 0;
    { uint32_t k = an / AVG_BLOCK_SIZE; Blk f; f.p = ab; f.len = AVG_BLOCK_SIZE; f.cnt = k; if (k) ws->rrun[ws->rrun_n++] = f;
      Blk l; l.p = ab + (uint64_t)k * AVG_BLOCK_SIZE; l.len = an - k * AVG_BLOCK_SIZE; l.cnt = 1; ws->rrun[ws->rrun_n++] = l; }
    { uint32_t k = bn2 / AVG_BLOCK_SIZE; Blk f; f.p = bb; f.len = AVG_BLOCK_SIZE; f.cnt = k; if (k) ws->rrun[ws->rrun_n++] = f;
      Blk l; l.p = bb + (uint64_t)k * AVG_BLOCK_SIZE; l.len = bn2 - k * AVG_BLOCK_SIZE; l.cnt = 1; ws->rrun[ws->rrun_n++] = l; }
    r.kind = RES_RUNS;
}

}  // namespace eb
