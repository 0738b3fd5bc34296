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

constexpr uint32_t FUSE_NONE = 0xffffffffu;

// The same classification with the hot table in SHARED memory: for lists shorter than 65 535 suffixes the per-class counter is a
// 16-bit word in the warp's scan scratch (WarpState::sc, idle while a fuse runs), so the read-modify-write that every 32-suffix
// step of the count and placement passes depends on is a shared-memory access instead of an L2 round trip. Lane l owns the
// classes 8l .. 8l+7: it turns its eight counters into class starts (one warp scan), publishes start and size to the global
// tables the children walk reads, and keeps which of its classes exist in `mybits`.
EB_DEV uint32_t fuse_classify_sm(const uint8_t* s, uint32_t len, const uint32_t* src, uint32_t k, uint32_t* dst, uint32_t base,
                                 uint16_t* tbl, uint32_t* gstart, uint32_t* gsize, uint32_t& mybits) {
    const int l = lane_id(); const uint32_t lt = (1u << l) - 1u;
#pragma unroll
    for (int i = 0; i < 8; i++) tbl[8 * l + i] = 0;
    __syncwarp();
    // Both passes take the list 128 suffixes at a time: four position loads in flight, then four byte gathers -- two memory round
    // trips per 128 suffixes instead of two per 32.
    // The suffix holding only the block's last byte is dropped when it is the FIRST of its class ([[]] -> [], :68-70): decided when
    // the count pass meets it -- "first" = the class counter is still zero and no lower lane of the step holds the same byte.
    uint32_t sq = 0xffffffffu, sch = 0x100u; bool special_dropped = false;
    for (uint32_t q0 = 0; q0 < k; q0 += 128) {                         // count
        uint32_t p[4], ch[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { uint32_t q = q0 + 32u * (uint32_t)j + (uint32_t)l; p[j] = q < k ? src[q] : 0xffffffffu; }
#pragma unroll
        for (int j = 0; j < 4; j++) ch[j] = p[j] < len ? (uint32_t)s[p[j]] : 0x100u;      // the [] suffix contributes nothing (:68)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (q0 + 32u * (uint32_t)j < k) {
                bool counted = ch[j] < 0x100u;
                const uint32_t sp = __ballot_sync(0xffffffffu, counted && p[j] + 1 == len);
                if (sp) {
                    const int ls = __ffs(sp) - 1;
                    const uint32_t c0 = __shfl_sync(0xffffffffu, ch[j], ls);
                    const uint32_t same = __ballot_sync(0xffffffffu, counted && ch[j] == c0);
                    if (tbl[c0] == 0 && (same & ((1u << ls) - 1u)) == 0) {
                        special_dropped = true; sq = q0 + 32u * (uint32_t)j + (uint32_t)ls; sch = c0;
                        if (l == ls) counted = false;
                    }
                }
                const uint32_t peers = __match_any_sync(0xffffffffu, counted ? ch[j] : 0x100u + (uint32_t)l);
                if (counted && (peers & lt) == 0) tbl[ch[j]] = (uint16_t)(tbl[ch[j]] + (uint32_t)__popc(peers));
                __syncwarp();
            }
        }
    }
    uint32_t cn[8]; uint32_t mine = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { cn[i] = tbl[8 * l + i]; mine += cn[i]; }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += x; }
    uint32_t rel = incl - mine; mybits = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {                                       // class starts, ascending byte
        uint32_t ch = 8u * (uint32_t)l + (uint32_t)i;
        bool exists = cn[i] != 0 || (special_dropped && sch == ch);
        if (exists) { mybits |= 1u << i; gstart[ch] = base + rel; gsize[ch] = cn[i]; }
        rel += cn[i];
        tbl[ch] = (uint16_t)rel;                                        // one past the class's last slot
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    for (uint32_t q0 = 0; q0 < k; q0 += 128) {                         // place: first arrival lands last in its class
        uint32_t p[4], ch[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { uint32_t q = q0 + 32u * (uint32_t)j + (uint32_t)l; p[j] = q < k ? src[q] : 0xffffffffu; }
#pragma unroll
        for (int j = 0; j < 4; j++) ch[j] = p[j] < len ? (uint32_t)s[p[j]] : 0x100u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (q0 + 32u * (uint32_t)j < k) {
                const bool placed = ch[j] < 0x100u && !(special_dropped && q0 + 32u * (uint32_t)j + (uint32_t)l == sq);
                const uint32_t peers = __match_any_sync(0xffffffffu, placed ? ch[j] : 0x100u + (uint32_t)l);
                const uint32_t end = placed ? tbl[ch[j]] : 0u;
                __syncwarp();
                if (placed) {
                    dst[base + end - 1 - (uint32_t)__popc(peers & lt)] = p[j] + 1;
                    if ((peers & lt) == 0) tbl[ch[j]] = (uint16_t)(end - (uint32_t)__popc(peers));
                }
                __syncwarp();
            }
        }
    }
    return total;
}

// ---- one node of at most FUSE_MID suffixes per side, whole warp, first bytes held in registers (fuse_mid, path 2a of fuse_step).
// Slot r of lane l holds list element r * 32 + l (arrival order). Encoding of a first byte: 0..255 live; 0x1ff the [] suffix
// (contributes nothing); byte | 0x400 the suffix holding only the block's last byte when it is the FIRST of its class -- the
// class exists, the element is not placed ([[]] -> [], :68-70). Classes of A are taken in ascending byte order by a min
// reduction; inside a class the LAST arrival comes first (char_suffixes prepends). No tables, no memory traffic but the
// two position lists and their first bytes.
constexpr uint32_t FUSE_MID = 128;
constexpr int FUSE_MID_SLOTS = 4;
// positions of one list and where its special suffix (position len - 1; positions of a list are distinct, so at most one) sits
__device__ __forceinline__ void fuse_mid_pos(const uint32_t* src, uint32_t k, uint32_t len, uint32_t (&pos)[FUSE_MID_SLOTS], int& rs, uint32_t& ls) {
    const int l = lane_id();
#pragma unroll
    for (int r = 0; r < FUSE_MID_SLOTS; r++) { uint32_t q = (uint32_t)r * 32 + (uint32_t)l; pos[r] = q < k ? src[q] : FUSE_NONE; }
    rs = -1; ls = 0;
#pragma unroll
    for (int r = 0; r < FUSE_MID_SLOTS; r++) {
        uint32_t sp = __ballot_sync(0xffffffffu, pos[r] + 1 == len);
        if (sp && rs < 0) { rs = r; ls = (uint32_t)__ffs(sp) - 1; }
    }
}
__device__ __forceinline__ void fuse_mid_drop(uint32_t (&cls)[FUSE_MID_SLOTS], int rs, uint32_t ls) {
    if (rs < 0) return;
    const int l = lane_id();
    uint32_t mine = 0x1ffu;
#pragma unroll
    for (int r = 0; r < FUSE_MID_SLOTS; r++) if (r == rs) mine = cls[r];
    uint32_t cs = __shfl_sync(0xffffffffu, mine, (int)ls);
    uint32_t before = 0;
#pragma unroll
    for (int r = 0; r < FUSE_MID_SLOTS; r++) {
        uint32_t m = __ballot_sync(0xffffffffu, cls[r] == cs);
        if (r < rs) before |= m; else if (r == rs) before |= m & ((1u << ls) - 1u);
    }
    if (!before && (uint32_t)l == ls) {
#pragma unroll
        for (int r = 0; r < FUSE_MID_SLOTS; r++) if (r == rs) cls[r] = cs | 0x400u;
    }
}

// Every node path below is its own function working on the search state in SHARED memory (FuseSh in the warp's WarpState):
// the per-case program's stack does not fit the L1 next to 24 other warps', so a spilled register or a by-reference local
// is an L2 round trip; shared memory is not. Each path reads what it needs, writes the next level's counters back.

// (2a) one node of up to FUSE_MID suffixes per side
EB_DEV bool fuse_mid(FuseSh* s, uint32_t sfo, uint32_t kA, uint32_t sto, uint32_t kB) {
    const int l = lane_id(); const uint32_t ltm = (1u << l) - 1u;
    const int cur = s->cur, nx = cur ^ 1;
    const uint8_t* a = s->a; const uint8_t* b = s->b; const uint32_t na = s->na, nb = s->nb;
    const uint32_t* srcA = s->F[cur] + sfo; const uint32_t* srcB = s->T[cur] + sto;
    uint32_t* Fn = s->F[nx]; uint32_t* Tn = s->T[nx]; FNode* NDn = s->ND[nx];
    uint32_t fo = s->fo, to = s->to, nnext = s->nnext, maxsz = s->maxsz;
    uint32_t ca[FUSE_MID_SLOTS], cb[FUSE_MID_SLOTS];
    {
        uint32_t pa[FUSE_MID_SLOTS], pb[FUSE_MID_SLOTS]; int rsa, rsb; uint32_t lsa, lsb;
        fuse_mid_pos(srcA, kA, na, pa, rsa, lsa);                          // both position lists first, then both byte gathers:
        fuse_mid_pos(srcB, kB, nb, pb, rsb, lsb);                          // two memory round trips for the node, not four
#pragma unroll
        for (int r = 0; r < FUSE_MID_SLOTS; r++) ca[r] = pa[r] < na ? (uint32_t)a[pa[r]] : 0x1ffu;
#pragma unroll
        for (int r = 0; r < FUSE_MID_SLOTS; r++) cb[r] = pb[r] < nb ? (uint32_t)b[pb[r]] : 0x1ffu;
        fuse_mid_drop(ca, rsa, lsa);
        fuse_mid_drop(cb, rsb, lsb);
    }
    const int nsa = (int)((kA + 31) >> 5), nsb = (int)((kB + 31) >> 5);
    int last = -1;
    for (;;) {
        uint32_t m = 0x1ffu;
#pragma unroll
        for (int r = 0; r < FUSE_MID_SLOTS; r++) { uint32_t v = ca[r] & 0x3ffu; if (r < nsa && (int)v > last && v < m) m = v; }
        const uint32_t ch = __reduce_min_sync(0xffffffffu, m);
        if (ch == 0x1ffu) break;
        last = (int)ch;
        uint32_t na_l = 0, nb_l = 0, eb_l = 0;
#pragma unroll
        for (int r = 0; r < FUSE_MID_SLOTS; r++) { if (r < nsa) na_l += ca[r] == ch ? 1u : 0u; }
#pragma unroll
        for (int r = 0; r < FUSE_MID_SLOTS; r++) { if (r < nsb) { nb_l += cb[r] == ch ? 1u : 0u; eb_l |= (cb[r] & 0x3ffu) == ch ? 1u : 0u; } }
        const uint32_t totA = __reduce_add_sync(0xffffffffu, na_l);
        if (totA == 0) {                 // {_Char, []} -> [[[]], []]: the two empty suffixes, whatever B holds (:91-93)
            if (fo + 1 > s->fcap || to + 1 > s->tcap || nnext >= s->ncap) return false;
            if (l == 0) { Fn[fo] = na; Tn[to] = nb; FNode q; q.fo = fo; q.fc = 1; q.to = to; q.tc = 1; NDn[nnext] = q; }
            fo++; to++; nnext++;
            maxsz = max(maxsz, 1u);
            continue;
        }
        if (!__any_sync(0xffffffffu, eb_l != 0)) continue;                  // notfound
        const uint32_t totB = __reduce_add_sync(0xffffffffu, nb_l);
        if ((uint64_t)fo + totA > s->fcap || (uint64_t)to + totB > s->tcap || nnext >= s->ncap) return false;
        uint32_t run = 0;
#pragma unroll
        for (int r = 0; r < FUSE_MID_SLOTS; r++) {
            if (r < nsa) {
                bool e = ca[r] == ch; uint32_t mb = __ballot_sync(0xffffffffu, e);
                if (e) Fn[fo + totA - 1 - (run + (uint32_t)__popc(mb & ltm))] = srcA[(uint32_t)r * 32 + (uint32_t)l] + 1;
                run += (uint32_t)__popc(mb);
            }
        }
        run = 0;
#pragma unroll
        for (int r = 0; r < FUSE_MID_SLOTS; r++) {
            if (r < nsb) {
                bool e = cb[r] == ch; uint32_t mb = __ballot_sync(0xffffffffu, e);
                if (e) Tn[to + totB - 1 - (run + (uint32_t)__popc(mb & ltm))] = srcB[(uint32_t)r * 32 + (uint32_t)l] + 1;
                run += (uint32_t)__popc(mb);
            }
        }
        if (l == 0) { FNode q; q.fo = fo; q.fc = totA; q.to = to; q.tc = totB; NDn[nnext] = q; }
        nnext++; fo += totA; to += totB;
        maxsz = max(maxsz, max(totA, totB));
    }
    __syncwarp();
    if (l == 0) { s->fo = fo; s->to = to; s->nnext = nnext; s->maxsz = maxsz; }
    __syncwarp();
    return true;
}

// (1) several consecutive nodes whose lists together hold at most 32 suffixes per side: lane j holds the j-th source suffix and the
// j-th target suffix of the run, lists back to back in the order the nodes are processed. A child is a (node, first byte) group:
// key = node << 9 | byte. match.any gives every lane its group, its size and its arrival rank; one all-pairs sweep over the 32
// lanes (shuffles, no memory) gives how many placed suffixes / target suffixes sort before the group. Same rules as everywhere:
// classes ascending inside a node, the last arrival first inside a class, the suffix holding only the block's last byte dropped
// when it is the first of its class, a class left empty by that turned into [[[]], []].
// nfo/nfc/nto/ntc = this lane's node (lane i = i-th node from the top), ia / ib = inclusive prefix sums of the list lengths,
// G = number of nodes taken (>= 1).
EB_DEV bool fuse_packed(FuseSh* s, uint32_t nfo, uint32_t nfc, uint32_t nto, uint32_t ntc, uint32_t ia, uint32_t ib, uint32_t G) {
    const int l = lane_id(); const uint32_t ltm = (1u << l) - 1u; const uint32_t j = (uint32_t)l;
    const int cur = s->cur, nx = cur ^ 1;
    const uint32_t na = s->na, nb = s->nb;
    const uint32_t totA = __shfl_sync(0xffffffffu, ia, (int)G - 1), totB = __shfl_sync(0xffffffffu, ib, (int)G - 1);
    // which node does suffix j belong to: the first node whose inclusive sum exceeds j
    int loa = 0, hia = (int)G - 1, lob = 0, hib = (int)G - 1;
#pragma unroll
    for (int it = 0; it < 5; it++) {
        int ma = (loa + hia) >> 1, mb = (lob + hib) >> 1;
        uint32_t va = __shfl_sync(0xffffffffu, ia, ma), vb = __shfl_sync(0xffffffffu, ib, mb);
        if (loa < hia) { if (va > j) hia = ma; else loa = ma + 1; }
        if (lob < hib) { if (vb > j) hib = mb; else lob = mb + 1; }
    }
    const uint32_t exa = __shfl_sync(0xffffffffu, ia - nfc, loa), foa = __shfl_sync(0xffffffffu, nfo, loa);
    const uint32_t exb = __shfl_sync(0xffffffffu, ib - ntc, lob), tob = __shfl_sync(0xffffffffu, nto, lob);
    uint32_t pa = FUSE_NONE, pb = FUSE_NONE;
    if (j < totA) pa = s->F[cur][foa + (j - exa)];
    if (j < totB) pb = s->T[cur][tob + (j - exb)];
    const bool la = pa < na, lb = pb < nb;                                   // NONE and the [] suffix contribute nothing
    const uint32_t ka = la ? ((uint32_t)loa << 9) | (uint32_t)s->a[pa] : 0x80000000u | j;
    const uint32_t kb = lb ? ((uint32_t)lob << 9) | (uint32_t)s->b[pb] : 0xc0000000u | j;
    const uint32_t ga = __match_any_sync(0xffffffffu, ka), gb = __match_any_sync(0xffffffffu, kb);
    const bool firsta = la && (ga & ltm) == 0, firstb = lb && (gb & ltm) == 0;
    const bool placeda = la && !(firsta && pa + 1 == na), placedb = lb && !(firstb && pb + 1 == nb);
    const uint32_t pma = __ballot_sync(0xffffffffu, placeda), pmb = __ballot_sync(0xffffffffu, placedb), lvb = __ballot_sync(0xffffffffu, lb);
    const uint32_t sza = (uint32_t)__popc(ga & pma), rka = (uint32_t)__popc(ga & pma & ltm);
    const uint32_t szb = (uint32_t)__popc(gb & pmb), rkb = (uint32_t)__popc(gb & pmb & ltm);
    uint32_t lessAA = 0, lessBA = 0, eqBA = 0, existsBA = 0, lessBB = 0;
    const int span = (int)max(totA, totB);
#pragma unroll 4
    for (int t = 0; t < span; t++) {
        uint32_t xa = __shfl_sync(0xffffffffu, ka, t), xb = __shfl_sync(0xffffffffu, kb, t);
        uint32_t pla = (pma >> t) & 1u, plb = (pmb >> t) & 1u;
        lessAA += (xa < ka) ? pla : 0u;
        lessBA += (xb < ka) ? plb : 0u;
        eqBA += (xb == ka) ? plb : 0u;
        existsBA |= (xb == ka) ? ((lvb >> t) & 1u) : 0u;
        lessBB += (xb < kb) ? plb : 0u;
    }
    const uint32_t nplA = (uint32_t)__popc(pma), nplB = (uint32_t)__popc(pmb);
    const bool quirk = firsta && sza == 0;
    const bool emit = firsta && (sza == 0 || existsBA);
    const uint32_t em = __ballot_sync(0xffffffffu, emit), qm = __ballot_sync(0xffffffffu, quirk);
    const uint32_t nq = (uint32_t)__popc(qm), nch = (uint32_t)__popc(em);
    const uint32_t fo = s->fo, to = s->to, nnext = s->nnext;
    if ((uint64_t)fo + nplA + nq > s->fcap || (uint64_t)to + nplB + nq > s->tcap || (uint64_t)nnext + nch > s->ncap) return false;
    uint32_t cidx = 0;                                                       // children with a smaller key come first
    for (uint32_t mm = em; mm;) { int t = __ffs(mm) - 1; mm &= mm - 1; uint32_t xk = __shfl_sync(0xffffffffu, ka, t); cidx += (xk < ka) ? 1u : 0u; }
    uint32_t* Fn = s->F[nx]; uint32_t* Tn = s->T[nx];
    if (placeda) Fn[fo + lessAA + sza - 1 - rka] = pa + 1;
    if (placedb) Tn[to + lessBB + szb - 1 - rkb] = pb + 1;
    uint32_t big = 0;
    if (emit) {
        FNode q;
        if (quirk) {
            uint32_t qi = (uint32_t)__popc(qm & ltm);
            Fn[fo + nplA + qi] = na; Tn[to + nplB + qi] = nb;
            q.fo = fo + nplA + qi; q.fc = 1; q.to = to + nplB + qi; q.tc = 1; big = 1;
        } else { q.fo = fo + lessAA; q.fc = sza; q.to = to + lessBA; q.tc = eqBA; big = max(sza, eqBA); }
        s->ND[nx][nnext + cidx] = q;
    }
    big = __reduce_max_sync(0xffffffffu, big);
    __syncwarp();
    if (l == 0) { s->fo = fo + nplA + nq; s->to = to + nplB + nq; s->nnext = nnext + nch; s->maxsz = max(s->maxsz, big); }
    __syncwarp();
    return true;
}

// (0) a run of TINY nodes (at most one suffix per side), one node per lane: the rules written out for one-element lists
EB_DEV bool fuse_tiny(FuseSh* s, uint32_t nfo, uint32_t nfc, uint32_t nto, uint32_t ntc, uint32_t tinyrun) {
    const int l = lane_id(); const uint32_t ltm = (1u << l) - 1u;
    const int cur = s->cur, nx = cur ^ 1;
    const uint32_t na = s->na, nb = s->nb;
    const bool act = (uint32_t)l < tinyrun;
    uint32_t pa = FUSE_NONE, pb = FUSE_NONE;
    if (act && nfc) pa = s->F[cur][nfo];
    if (act && ntc) pb = s->T[cur][nto];
    const bool ha = pa < na, hb = pb < nb;                       // NONE and the [] suffix contribute nothing
    const uint32_t cha = ha ? s->a[pa] : 0x100u, chb = hb ? s->b[pb] : 0x200u;
    bool child = false; uint32_t ca = 0, cb = 0, tcn = 1;
    if (ha && pa + 1 == na) { child = true; ca = na; cb = nb; }                    // {_Char, []} -> [[[]], []] (:91-93)
    else if (cha == chb) { child = true; ca = pa + 1; cb = pb + 1; if (pb + 1 == nb) tcn = 0; }
    const uint32_t cm = __ballot_sync(0xffffffffu, child);
    const uint32_t cn = (uint32_t)__popc(cm);
    const uint32_t fo = s->fo, to = s->to, nnext = s->nnext;
    if ((uint64_t)fo + cn > s->fcap || (uint64_t)to + cn > s->tcap || (uint64_t)nnext + cn > s->ncap) return false;
    if (child) {
        uint32_t at = (uint32_t)__popc(cm & ltm);
        s->F[nx][fo + at] = ca; s->T[nx][to + at] = cb;
        FNode q; q.fo = fo + at; q.fc = 1; q.to = to + at; q.tc = tcn; s->ND[nx][nnext + at] = q;
    }
    __syncwarp();
    if (l == 0) { s->fo = fo + cn; s->to = to + cn; s->nnext = nnext + cn; if (cn) s->maxsz = max(s->maxsz, 1u); }
    __syncwarp();
    return true;
}

// (2b) a node with longer lists, counters in shared memory (fuse_classify_sm), then the children taken eight classes per lane;
// (2c) lists of 65 535 suffixes and more: the same through counters in global memory (fuse_classify)
EB_DEV bool fuse_tables(FuseSh* s, uint16_t* sc, uint32_t sfo, uint32_t kA, uint32_t sto, uint32_t kB) {
    const int l = lane_id();
    const int cur = s->cur, nx = cur ^ 1;
    const uint32_t na = s->na, nb = s->nb;
    uint32_t* tabs = s->tabs;
    uint32_t* startA = tabs + 256; uint32_t* startB = tabs + 768; uint32_t* sizeA = tabs + 1024; uint32_t* sizeB = tabs + 1280;
    uint32_t fo = s->fo, to = s->to, nnext = s->nnext, maxsz = s->maxsz;
    if (kA < 65535u && kB < 65535u) {
        uint32_t ma, mb;
        const uint32_t fa2 = fuse_classify_sm(s->a, na, s->F[cur] + sfo, kA, s->F[nx], fo, sc, startA, sizeA, ma);
        const uint32_t tb2 = fuse_classify_sm(s->b, nb, s->T[cur] + sto, kB, s->T[nx], to, sc, startB, sizeB, mb);
        uint32_t szs[8]; uint32_t nch_l = 0, nq_l = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            szs[i] = 0xffffffffu;                                                     // no child
            if ((ma >> i) & 1u) {
                uint32_t sz = sizeA[8 * l + i];
                if (sz == 0) { szs[i] = 0; nch_l++; nq_l++; }                         // {_Char, []} -> [[[]], []] (:91-93)
                else if ((mb >> i) & 1u) { szs[i] = sz; nch_l++; }                    // else notfound
            }
        }
        uint32_t cpre = nch_l, qpre = nq_l;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            uint32_t x = __shfl_up_sync(0xffffffffu, cpre, o2), y = __shfl_up_sync(0xffffffffu, qpre, o2);
            if (l >= o2) { cpre += x; qpre += y; }
        }
        const uint32_t ctot = __shfl_sync(0xffffffffu, cpre, 31), qtot = __shfl_sync(0xffffffffu, qpre, 31);
        if ((uint64_t)fo + fa2 + qtot > s->fcap || (uint64_t)to + tb2 + qtot > s->tcap || (uint64_t)nnext + ctot > s->ncap) return false;
        uint32_t ci = nnext + cpre - nch_l, qi = qpre - nq_l, big = 0;
        uint32_t* Fn = s->F[nx]; uint32_t* Tn = s->T[nx]; FNode* NDn = s->ND[nx];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (szs[i] == 0xffffffffu) continue;
            FNode q; uint32_t ch = 8u * (uint32_t)l + (uint32_t)i;
            if (szs[i] == 0) {
                Fn[fo + fa2 + qi] = na; Tn[to + tb2 + qi] = nb;
                q.fo = fo + fa2 + qi; q.fc = 1; q.to = to + tb2 + qi; q.tc = 1; qi++; big = max(big, 1u);
            } else { q.fo = startA[ch]; q.fc = szs[i]; q.to = startB[ch]; q.tc = sizeB[ch]; big = max(big, max(q.fc, q.tc)); }
            NDn[ci++] = q;
        }
        big = __reduce_max_sync(0xffffffffu, big);
        __syncwarp();
        if (l == 0) { s->fo = fo + fa2 + qtot; s->to = to + tb2 + qtot; s->nnext = nnext + ctot; s->maxsz = max(maxsz, big); }
        __syncwarp();
        return true;
    }
    FuseSide sa, sb; sa.s = s->a; sa.len = na; sa.cnt = tabs; sa.start = startA; sb.s = s->b; sb.len = nb; sb.cnt = tabs + 512; sb.start = startB;
    uint32_t* Fn = s->F[nx]; uint32_t* Tn = s->T[nx]; FNode* NDn = s->ND[nx];
    uint32_t fa = fuse_classify(sa, s->F[cur] + sfo, kA, Fn, fo, sizeA);
    uint32_t tb = fuse_classify(sb, s->T[cur] + sto, kB, Tn, to, sizeB);
    for (int w = 0; w < 8; w++) {
        uint32_t m = sa.bits[w];
        while (m) {
            uint32_t bit = (uint32_t)__ffs(m) - 1; m &= m - 1; uint32_t ch = (uint32_t)w * 32 + bit;
            FNode ch_n;
            if (sizeA[ch] == 0) {       // {_Char, []} -> [[[]], []]: the two empty suffixes, whatever B holds (:91-93)
                if (fo + fa + 1 > s->fcap || to + tb + 1 > s->tcap || nnext >= s->ncap) return false;
                if (l == 0) { Fn[fo + fa] = na; Tn[to + tb] = nb; }
                ch_n.fo = fo + fa; ch_n.fc = 1; ch_n.to = to + tb; ch_n.tc = 1; fa++; tb++;
                if (l == 0) NDn[nnext] = ch_n;
                nnext++; continue;
            }
            if (!((sb.bits[w] >> bit) & 1u)) continue;                     // notfound
            if (nnext >= s->ncap) return false;
            ch_n.fo = sa.start[ch]; ch_n.fc = sizeA[ch]; ch_n.to = sb.start[ch]; ch_n.tc = sizeB[ch];
            if (l == 0) NDn[nnext] = ch_n;
            nnext++;
            maxsz = max(maxsz, max(ch_n.fc, ch_n.tc));
        }
    }
    __syncwarp();
    if (l == 0) { s->fo = fo + fa; s->to = to + tb; s->nnext = nnext; s->maxsz = maxsz; }
    __syncwarp();
    return true;
}

// one flat level (every node at most 1 x 1, and so all their descendants): two position arrays PA[k] / PB[k] = the node's source /
// target position (NONE = no suffix), kept in the F / T ping-pong buffers; a step reads 32 nodes with two coalesced loads and
// two byte gathers, nothing depends on the previous step but the output offset. Returns the number of children.
EB_DEV uint32_t fuse_flat_level(FuseSh* s) {
    const int cur = s->cur, nx = cur ^ 1;
    const uint32_t* PA = s->F[cur]; const uint32_t* PB = s->T[cur]; uint32_t* QA = s->F[nx]; uint32_t* QB = s->T[nx];
    const uint8_t* a = s->a; const uint8_t* b = s->b; const uint32_t na = s->na, nb = s->nb;
    const int l = lane_id(); const uint32_t ltm = (1u << l) - 1u;
    uint32_t nnext = 0;
    for (uint32_t e = s->ncur; e > 0;) {
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
    return nnext;
}

// one step of a general level: lane i looks at the i-th node from the top of the list and the run goes to the path that fits.
// What a level looks like after the first split or two (profiles/tc_c2_r2d.txt, tc_c5_r2d.txt): nine nodes in ten hold one suffix
// per side, most of the rest a handful, a few dozen hold 25-250 (a phrase or a period repeated in the block), and the root and
// its first children thousands.
EB_DEV bool fuse_step(FuseSh* s, uint16_t* sc) {
    const int l = lane_id();
    const uint32_t e = s->e;
    const bool valid = (uint32_t)l < e;
    FNode nd; nd.fo = nd.to = 0; nd.fc = nd.tc = 64;
    if (valid) nd = s->ND[s->cur][e - 1 - (uint32_t)l];
    FuseClock clk; clk.tab = s->clk; clk.t0 = 0;
    const uint32_t tm = __ballot_sync(0xffffffffu, valid && nd.fc <= 1 && nd.tc <= 1);
    const uint32_t tinyrun = tm == 0xffffffffu ? 32u : (uint32_t)__ffs(~tm) - 1u;
    if (tinyrun >= 16 || (tinyrun > 0 && tinyrun == e)) {
        clk.start();
        if (!fuse_tiny(s, nd.fo, nd.fc, nd.to, nd.tc, tinyrun)) return false;
        if (l == 0) s->e = e - tinyrun;
        __syncwarp();
        clk.stop(FPH_TINY);
        return true;
    }
    uint32_t ia = min(nd.fc, 64u), ib = min(nd.tc, 64u);
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
        uint32_t x = __shfl_up_sync(0xffffffffu, ia, o2), y = __shfl_up_sync(0xffffffffu, ib, o2);
        if (l >= o2) { ia += x; ib += y; }
    }
    const uint32_t okm = __ballot_sync(0xffffffffu, valid && ia <= 32 && ib <= 32);
    const uint32_t G = okm == 0xffffffffu ? 32u : (uint32_t)__ffs(~okm) - 1u;
    if (G > 0) {
        clk.start();
        if (!fuse_packed(s, nd.fo, nd.fc, nd.to, nd.tc, ia, ib, G)) return false;
        if (l == 0) s->e = e - G;
        __syncwarp();
        clk.stop(FPH_SMALL);
        return true;
    }
    const uint32_t tfo = __shfl_sync(0xffffffffu, nd.fo, 0), tfc = __shfl_sync(0xffffffffu, nd.fc, 0);
    const uint32_t tto = __shfl_sync(0xffffffffu, nd.to, 0), ttc = __shfl_sync(0xffffffffu, nd.tc, 0);
    clk.start();
    bool ok;
    const bool midp = tfc <= FUSE_MID && ttc <= FUSE_MID;
    if (midp) ok = fuse_mid(s, tfo, tfc, tto, ttc);
    else ok = fuse_tables(s, sc, tfo, tfc, tto, ttc);
    if (l == 0) s->e = e - 1;
    __syncwarp();
    clk.stop(midp ? FPH_MID : FPH_BIG);
    return ok;
}

// find_jump_points/2 :103-128 + any_position_pair/1 :73-77
EB_DEV bool fuse_device(CaseCtx& c, const uint8_t* a, uint32_t na, const uint8_t* b, uint32_t nb, uint32_t& from, uint32_t& to) {
    Rng& g = c.rng;
    FuseSh* s = &c.ws->fsh;
    const uint64_t fcap = (uint64_t)na + 80, tcap = (uint64_t)nb + 80, ncap = (uint64_t)(na < nb ? na : nb) + 80;
    {
        uint32_t* F[2]; uint32_t* T[2]; FNode* ND[2];
        for (int i = 0; i < 2; i++) { F[i] = (uint32_t*)temp_alloc(c, fcap * 4); T[i] = (uint32_t*)temp_alloc(c, tcap * 4); ND[i] = (FNode*)temp_alloc(c, ncap * sizeof(FNode)); }
        uint32_t* tabs = (uint32_t*)temp_alloc(c, 6 * 256 * 4);
        if (!F[0] || !F[1] || !T[0] || !T[1] || !ND[0] || !ND[1] || !tabs) return false;
        for (uint32_t i = lane_id(); i < 256; i += 32) { tabs[i] = 0; tabs[512 + i] = 0; }      // the global class counters rest at zero
        for (uint32_t i = lane_id(); i < na; i += 32) F[0][i] = i;
        for (uint32_t i = lane_id(); i < nb; i += 32) T[0][i] = i;
        __syncwarp();
        if (lane_id() == 0) {
            s->a = a; s->b = b; s->na = na; s->nb = nb;
            for (int i = 0; i < 2; i++) { s->F[i] = F[i]; s->T[i] = T[i]; s->ND[i] = ND[i]; }
            s->fcap = (uint32_t)fcap; s->tcap = (uint32_t)tcap; s->ncap = (uint32_t)ncap; s->tabs = tabs;
            s->cur = 0; s->ncur = 1;
            s->clk = c.ar.mut_ns ? c.ar.mut_ns + 2 * M_COUNT : nullptr;
            FNode r0; r0.fo = 0; r0.fc = na; r0.to = 0; r0.tc = nb; ND[0][0] = r0;
        }
        __syncwarp();
    }
    int64_t fuel = 100000;
    bool compact = false;
    for (;;) {
        bool stop = fuel < 0;
        if (!stop) stop = g.rand(8) == 0;
        if (!stop && compact) {
            FuseClock clk; clk.tab = s->clk; clk.t0 = 0; clk.start();
            const uint32_t nnext = fuse_flat_level(s);
            clk.stop(FPH_COMPACT);
            if (nnext == 0) stop = true;
            else { fuel -= (int64_t)nnext; if (lane_id() == 0) { s->cur ^= 1; s->ncur = nnext; } __syncwarp(); continue; }
        }
        if (!stop) {
            // nodes are stored in emission order; the reference's list is that order reversed: a level is split from the top
            if (lane_id() == 0) { s->fo = 0; s->to = 0; s->nnext = 0; s->maxsz = 0; s->e = s->ncur; }
            __syncwarp();
            while (s->e > 0) {
                if (!fuse_step(s, c.ws->sc)) { c.ws->status = CASE_OVERFLOW; c.ws->reason = 8; return false; }
            }
            const uint32_t nnext = s->nnext;
            if (nnext == 0) stop = true;
            else {
                fuel -= (int64_t)nnext;
                const uint32_t maxsz = s->maxsz;
                __syncwarp();
                if (lane_id() == 0) { s->cur ^= 1; s->ncur = nnext; }
                __syncwarp();
                if (maxsz <= 1) {      // from here on every node is at most 1 x 1: flatten the level into position arrays
                    const int cur = s->cur, ot = cur ^ 1;
                    const FNode* NDc = s->ND[cur]; const uint32_t* Fc = s->F[cur]; const uint32_t* Tc = s->T[cur]; uint32_t* Fo = s->F[ot]; uint32_t* To = s->T[ot];
                    for (uint32_t k = (uint32_t)lane_id(); k < nnext; k += 32) {
                        FNode nd = NDc[k];
                        Fo[k] = nd.fc ? Fc[nd.fo] : FUSE_NONE;
                        To[k] = nd.tc ? Tc[nd.to] : FUSE_NONE;
                    }
                    __syncwarp();
                    if (lane_id() == 0) s->cur = ot;
                    __syncwarp();
                    compact = true;
                }
                continue;
            }
        }
        // any_position_pair(Nodes)
        const int cur = s->cur; const uint32_t ncur = s->ncur;
        if (compact) {
            uint32_t r = (uint32_t)g.rand_elem_idx(ncur);
            uint32_t pa = s->F[cur][ncur - 1 - r], pb = s->T[cur][ncur - 1 - r];
            int64_t fi = g.rand_elem_idx(pa != FUSE_NONE ? 1 : 0);
            from = fi < 0 ? na : pa;
            int64_t ti = g.rand_elem_idx(pb != FUSE_NONE ? 1 : 0);
            to = ti < 0 ? nb : pb;
            return true;
        }
        uint32_t r = (uint32_t)g.rand_elem_idx(ncur);
        FNode nd = s->ND[cur][ncur - 1 - r];
        int64_t fi = g.rand_elem_idx(nd.fc);
        from = fi < 0 ? na : s->F[cur][nd.fo + (uint32_t)fi];
        int64_t ti = g.rand_elem_idx(nd.tc);
        to = ti < 0 ? nb : s->T[cur][nd.to + (uint32_t)ti];
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
    EB_RECONVERGE();
    ws->rrun_n = 0;
    { uint32_t k = an / AVG_BLOCK_SIZE; Blk f; f.p = ab; f.len = AVG_BLOCK_SIZE; f.cnt = k; if (k) ws->rrun[ws->rrun_n++] = f;
      Blk l; l.p = ab + (uint64_t)k * AVG_BLOCK_SIZE; l.len = an - k * AVG_BLOCK_SIZE; l.cnt = 1; ws->rrun[ws->rrun_n++] = l; }
    { uint32_t k = bn2 / AVG_BLOCK_SIZE; Blk f; f.p = bb; f.len = AVG_BLOCK_SIZE; f.cnt = k; if (k) ws->rrun[ws->rrun_n++] = f;
      Blk l; l.p = bb + (uint64_t)k * AVG_BLOCK_SIZE; l.len = bn2 - k * AVG_BLOCK_SIZE; l.cnt = 1; ws->rrun[ws->rrun_n++] = l; }
    r.kind = RES_RUNS;
}

}  // namespace eb
