// erlamsa_b200 -- the mutators (device side).
//
// Every function here mirrors one mutator of reference src/erlamsa_mutations.erl (file:line
// cited per function) with the reference's exact order of RNG draws, but instead of building
// a new binary it emits an EDIT SCRIPT (tseg) over the existing block memory; the bytes move
// once, later, in the apply kernel. `This` = (p, n) is the head block hd(Ll).
//
// Result protocol: RES_SAME  -> the mutator returned Ll unchanged (counts as "failed"),
//                  RES_SEGS  -> ws->tseg/ntseg/tlen hold the new head; `rechunk` = flush_bvecs.
#pragma once
#include "eb_state.cuh"

namespace eb {

enum { RES_SAME = 0, RES_SEGS = 1, RES_UNSUPPORTED = 2, RES_RUNS = 3 };
struct MutResult { int kind; double delta; int rechunk; int consumed_next; };

// funny_unicode/0 table (reference :1052-1078), built on the host at init
struct FunnyEntry { uint32_t len; uint32_t bytes; };
__constant__ FunnyEntry c_funny[192];
__constant__ int c_funny_n;

// ------------------------------------------------------------------ single byte :171-223, utf8 :1080-1099
EB_DEV void mut_byte(CaseCtx& c, int id, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    uint32_t pos = (uint32_t)g.rand(n);
    r.delta = g.rand_delta(); r.rechunk = 0; r.consumed_next = 0;
    if (id == M_UI) {   // the list element is picked before edit_byte_vector runs
        int64_t e = g.rand_elem_idx((uint64_t)c_funny_n);
        if (n == 0) { r.kind = RES_SAME; return; }
        t_reset(ws);
        t_push(ws, seg_copy(p, pos + 1)); t_push(ws, seg_inline(c_funny[e].bytes, c_funny[e].len)); t_push(ws, seg_copy(p + pos + 1, n - pos - 1));
        r.kind = RES_SEGS; return;
    }
    if (n == 0) { r.kind = RES_SAME; return; }
    uint32_t b = p[pos];
    uint64_t lit = 0; uint32_t ll = 0;
    switch (id) {
    case M_BD: break;
    case M_BEI: lit = (b + 1) & 255; ll = 1; break;
    case M_BED: lit = (b - 1) & 255; ll = 1; break;
    case M_BR: lit = b | (b << 8); ll = 2; break;
    case M_BF: lit = b ^ (1u << g.rand(8)); ll = 1; break;
    case M_BI: lit = (uint32_t)g.rand(256) | (b << 8); ll = 2; break;
    case M_BER: lit = g.rand(256); ll = 1; break;
    case M_UW: if (b == (b & 0x3f)) { lit = 0xc0u | ((b | 0x80u) << 8); ll = 2; } else { r.kind = RES_SAME; return; } break;
    }
    t_reset(ws);
    t_push(ws, seg_copy(p, pos)); t_push(ws, seg_inline(lit, ll)); t_push(ws, seg_copy(p + pos + 1, n - pos - 1));
    r.kind = RES_SEGS;
}

// ------------------------------------------------------------------ byte sequences :232-318
// warp bitonic sort of (key, byte) pairs living in scratch; m is a power of two
EB_DEV void warp_bitonic(uint64_t* K, uint8_t* V, uint32_t m) {
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane_id(); i < m; i += 32) {
                uint32_t x = i ^ j;
                if (x > i) {
                    uint64_t a = K[i], b = K[x]; uint8_t va = V[i], vb = V[x];
                    bool up = (i & k) == 0;
                    bool gt = (a > b) || (a == b && va > vb);
                    if (gt == up) { K[i] = b; K[x] = a; V[i] = vb; V[x] = va; }
                }
            }
            __syncwarp();
        }
}
EB_DEV void mut_bytes(CaseCtx& c, int id, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0;
    if (n == 0) { r.kind = RES_SAME; r.delta = -1; return; }
    uint32_t s = (uint32_t)g.rand(n);
    uint32_t l = (uint32_t)g.rand_range(1, (int64_t)(n - s + 1));
    t_reset(ws);
    t_push(ws, seg_copy(p, s));
    if (id == M_SD) {
    } else if (id == M_SR) {
        uint64_t k = g.rand_log_small(10); if (k < 2) k = 2;
        uint64_t tot = k * (uint64_t)l;
        if (tot > c.bp->max_case_out) { ws->status = CASE_OVERFLOW; ws->reason = 5; r.kind = RES_SAME; r.delta = 0; return; }
        t_push(ws, seg_repeat(p + s, l, (uint32_t)tot));
    } else if (id == M_SP) {
        // random_permutation/1, reference src/erlamsa_rnd.erl:190-196
        if (l == 2) {
            if (g.rand(2) == 1) t_push(ws, seg_inline((uint64_t)p[s + 1] | ((uint64_t)p[s] << 8), 2));
            else t_push(ws, seg_copy(p + s, 2));
        } else {
            uint32_t m = 1; while (m < l) m <<= 1;
            uint8_t* buf = scratch_alloc(c, (uint64_t)m * 9);
            if (!buf) { r.kind = RES_SAME; r.delta = 0; return; }
            uint64_t* K = (uint64_t*)buf; uint8_t* V = buf + (uint64_t)m * 8;
            if (g.mode != 0) {   // Philox: key i is draw i of this mutator's slot -- 32 keys per step
                for (uint32_t i = lane_id(); i < l; i += 32) { K[i] = (uint64_t)__double_as_longlong(g.uniform_philox_at(g.local + i)); V[i] = p[s + i]; }
                g.philox_skip(l);
            } else
            for (uint32_t i = 0; i < l; i++) { double u = g.uniform(); if (lane_id() == 0) { K[i] = (uint64_t)__double_as_longlong(u); V[i] = p[s + i]; } }
            for (uint32_t i = l + lane_id(); i < m; i += 32) { K[i] = ~0ull; V[i] = 0xff; }
            __syncwarp();
            warp_bitonic(K, V, m);
            t_push(ws, seg_copy(V, l));
        }
    } else {   // randmask :279-307 -- one occurs draw per byte (drawn one byte ahead), plus a mask draw when it hit
        int kind = (id == M_SNAND) ? c.snand_kind : 3;
        uint8_t* buf = scratch_alloc(c, l);
        if (!buf) { r.kind = RES_SAME; r.delta = 0; return; }
        uint64_t prob = g.erand(100);
        if (g.mode != 0) {   // Philox: byte i owns draws 2i (does the mask apply) and 2i + 1 (which bit / byte) of this slot
            for (uint32_t i = lane_id(); i < l; i += 32) {
                uint64_t nn = (uint64_t)trunc(g.uniform_philox_at(g.local + 2 * i) * 100.0);
                bool oc = prob == 1 ? nn != 0 : nn < prob;                      // rand_occurs_fixed/2, quirk included
                uint32_t h = p[s + i];
                if (oc) {
                    double u = g.uniform_philox_at(g.local + 2 * i + 1);
                    uint32_t bit = 1u << (uint32_t)trunc(u * 8.0);
                    if (kind == 0) h &= ~bit; else if (kind == 1) h |= bit; else if (kind == 2) h ^= bit; else h = (uint32_t)trunc(u * 256.0);
                }
                buf[i] = (uint8_t)h;
            }
            g.philox_skip(2 * l);
        } else {
        bool occ = g.rand_occurs_fixed(prob, 100);
        for (uint32_t i = 0; i < l; i++) {
            bool nxt = g.rand_occurs_fixed(prob, 100);
            uint32_t h = p[s + i];
            if (occ) {
                if (kind == 0) h &= ~(1u << g.rand(8));
                else if (kind == 1) h |= (1u << g.rand(8));
                else if (kind == 2) h ^= (1u << g.rand(8));
                else h = (uint32_t)g.rand(256);
            }
            if (lane_id() == 0) buf[i] = (uint8_t)h;
            occ = nxt;
        }
        }
        __syncwarp();
        t_push(ws, seg_copy(buf, l));
    }
    t_push(ws, seg_copy(p + s + l, n - s - l));
    r.delta = g.rand_delta();
    r.kind = RES_SEGS;
}

// ------------------------------------------------------------------ sed_num :63-169
EB_DEV void interesting_number(int idx, Big256& o) {   // :68-75 (list order after the foldl)
    const int is[11] = {128, 127, 64, 63, 32, 31, 16, 15, 8, 7, 1};
    int g = idx / 3, w = idx % 3;
    o.set_pow2((unsigned)is[g]);
    Big256 one; one.set_u64(1);
    if (w == 0) o.sub_abs(one); else if (w == 2) o.add_abs(one);
}
EB_DEV void mutate_num(Rng& g, Big256& num) {   // :92-112
    uint64_t cs = g.rand(12);
    Big256 t;
    switch (cs) {
    case 0: t.set_u64(1); num.add(t); return;
    case 1: t.set_u64(1); num.sub(t); return;
    case 2: num.zero(); return;
    case 3: num.set_u64(1); return;
    case 4: case 5: interesting_number((int)g.rand_elem_idx(33), num); return;
    case 7: interesting_number((int)g.rand_elem_idx(33), t); num.add(t); return;
    case 8: interesting_number((int)g.rand_elem_idx(33), t); num.sub(t); return;
    case 9: {
        Big256 a = num; a.neg = 0; a.mul_small(2);
        Big256 rr = g.rand_big(a); rr.ovf |= a.ovf;
        if (num.neg) num.add(rr); else num.sub(rr);
        return;
    }
    case 10: num.negate(); return;
    default: {
        uint64_t n = (uint64_t)g.rand_range(1, 129);
        Big256 l = g.rand_log_big(n);
        uint64_t s = g.rand(3);
        if (s == 0) num.sub(l); else num.add(l);
        return;
    }
    }
}
EB_DEV void mut_num(CaseCtx& c, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 1; r.consumed_next = 0;
    // numbers = maximal digit runs (get_num/4 :114-124 swallows the '-' run right before one)
    uint32_t nfound = scan_count_jobs<PRED_DIGIT, true>(c.q, p, n, ws->sc, &ws->qpend);
    uint64_t which = g.rand(nfound);
    t_reset(ws);
    if (nfound == 0) {
        t_push(ws, seg_copy(p, n));
    } else {
        uint32_t d0 = scan_select<PRED_DIGIT, true>(p, n, ws->sc, nfound - 1 - (uint32_t)which);
        uint32_t a = d0; while (a > 0 && p[a - 1] == '-') a--;
        uint32_t b = d0; while (b < n && (uint32_t)(p[b] - '0') < 10u) b++;
        Big256 v; v.zero();
        if (b - d0 > 77) v.ovf = 1;
        else for (uint32_t i = d0; i < b; i++) { v.mul_small(10); v.add_small((uint32_t)(p[i] - '0')); }
        if (a < d0 && !v.is_zero()) v.neg = 1;
        mutate_num(g, v);
        if (v.ovf) { ws->status = CASE_UNSUPPORTED; r.kind = RES_SAME; r.delta = 0; return; }
        uint8_t dec[88]; int dl = v.to_decimal(dec);
        uint8_t* lit = scratch_alloc(c, (uint64_t)dl);
        if (!lit) { r.kind = RES_SAME; r.delta = 0; return; }
        if (lane_id() < dl) lit[lane_id()] = dec[lane_id()];
        if (lane_id() + 32 < dl) lit[lane_id() + 32] = dec[lane_id() + 32];
        if (lane_id() + 64 < dl) lit[lane_id() + 64] = dec[lane_id() + 64];
        __syncwarp();
        t_push(ws, seg_copy(p, a)); t_push(ws, seg_copy(lit, (uint32_t)dl)); t_push(ws, seg_copy(p + b, n - b));
    }
    bool isbin = segs_binarish(ws->tseg, ws->ntseg, ws->tlen);
    if (nfound == 0) { uint64_t x = g.rand(10); r.delta = x == 0 ? -1 : 0; }
    else if (isbin) r.delta = -1;
    else r.delta = 2;
    r.kind = RES_SEGS;
}

// ------------------------------------------------------------------ lines :320-378 + reference src/erlamsa_generic.erl
struct LineTab { const uint8_t* p; uint32_t n; uint32_t nl; uint32_t nlines; const uint16_t* sc; };
EB_DEV uint32_t line_start(const LineTab& t, uint32_t i) { return i == 0 ? 0 : scan_select<PRED_NEWLINE, false>(t.p, t.n, t.sc, i - 1) + 1; }
EB_DEV uint32_t line_end(const LineTab& t, uint32_t i) { return i < t.nl ? scan_select<PRED_NEWLINE, false>(t.p, t.n, t.sc, i) + 1 : t.n; }

EB_DEV bool try_lines(CaseCtx& c, const uint8_t* p, uint32_t n, LineTab& t) {   // :341-348
    if (n == 0) return false;
    t.p = p; t.n = n; t.sc = c.ws->sc;
    t.nl = scan_count_jobs<PRED_NEWLINE, false>(c.q, p, n, c.ws->sc, &c.ws->qpend);
    t.nlines = t.nl + (p[n - 1] != 10 ? 1u : 0u);
    if (mem_binarish(p, n)) return false;
    return true;
}
EB_DEV void mut_line(CaseCtx& c, int id, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0;
    LineTab t;
    if (!try_lines(c, p, n, t)) { r.kind = RES_SAME; r.delta = -1; return; }
    uint32_t len = t.nlines;
    r.delta = 1; r.kind = RES_SEGS;
    t_reset(ws);
    switch (id) {
    case M_LD: {
        uint32_t q = (uint32_t)g.erand(len) - 1; uint32_t a = line_start(t, q), b = line_end(t, q);
        t_push(ws, seg_copy(p, a)); t_push(ws, seg_copy(p + b, n - b)); break;
    }
    case M_LDS: {
        uint32_t st = (uint32_t)g.erand(len); uint32_t cnt = (uint32_t)g.erand(len - st + 1);
        uint32_t a = line_start(t, st - 1), b = line_end(t, st - 1 + cnt - 1);
        t_push(ws, seg_copy(p, a)); t_push(ws, seg_copy(p + b, n - b)); break;
    }
    case M_LR2: {
        uint32_t q = (uint32_t)g.erand(len) - 1; uint32_t a = line_start(t, q), b = line_end(t, q);
        t_push(ws, seg_copy(p, b)); t_push(ws, seg_copy(p + a, n - a)); break;
    }
    case M_LR: {
        uint32_t q = (uint32_t)g.erand(len) - 1; uint64_t k = g.rand_log_small(10); if (k < 2) k = 2;
        uint32_t a = line_start(t, q), b = line_end(t, q);
        uint64_t tot = k * (uint64_t)(b - a);
        if (tot > c.bp->max_case_out) { ws->status = CASE_OVERFLOW; ws->reason = 5; r.kind = RES_SAME; return; }
        t_push(ws, seg_copy(p, a)); t_push(ws, seg_repeat(p + a, b - a, (uint32_t)tot)); t_push(ws, seg_copy(p + b, n - b)); break;
    }
    case M_LRI: {
        uint32_t from = (uint32_t)g.erand(len) - 1, to = (uint32_t)g.erand(len) - 1;
        uint32_t fa = line_start(t, from), fb = line_end(t, from), ta = line_start(t, to), tb = line_end(t, to);
        t_push(ws, seg_copy(p, ta)); t_push(ws, seg_copy(p + fa, fb - fa)); t_push(ws, seg_copy(p + tb, n - tb)); break;
    }
    case M_LS: {
        if (len < 2) { r.kind = RES_SAME; return; }
        uint32_t q = (uint32_t)g.erand(len - 1) - 1;
        uint32_t a = line_start(t, q), m = line_end(t, q), b = line_end(t, q + 1);
        t_push(ws, seg_copy(p, a)); t_push(ws, seg_copy(p + m, b - m)); t_push(ws, seg_copy(p + a, m - a)); t_push(ws, seg_copy(p + b, n - b)); break;
    }
    default: {   // M_LP list_perm, reference src/erlamsa_generic.erl:105-116
        if (len < 3) { r.kind = RES_SAME; return; }
        uint32_t from = (uint32_t)g.erand(len - 1);
        uint64_t a = (uint64_t)g.rand_range(2, (int64_t)(len - from));
        uint64_t b = g.rand_log_small(10);
        uint32_t k = (uint32_t)(a < b ? a : b); if (k < 2) k = 2;
        // the k lines from `from` (1-based) are permuted: keys drawn in list order, sorted ascending
        uint32_t first = from - 1;
        uint32_t ra = line_start(t, first), rb = line_end(t, first + k - 1);
        uint8_t* out = scratch_alloc(c, (uint64_t)(rb - ra));
        uint8_t* tab = scratch_alloc(c, (uint64_t)k * 16);
        if (!out || !tab) { r.kind = RES_SAME; return; }
        uint64_t* K = (uint64_t*)tab; uint32_t* S = (uint32_t*)(tab + (uint64_t)k * 8); uint32_t* E = S + k;
        if (k == 2) {
            bool sw = g.rand(2) == 1;
            uint32_t m = line_end(t, first);
            if (lane_id() == 0) { S[0] = sw ? m : ra; E[0] = sw ? rb : m; S[1] = sw ? ra : m; E[1] = sw ? m : rb; }
        } else {
            for (uint32_t i = 0; i < k; i++) {
                double u = g.uniform(); uint32_t s0 = line_start(t, first + i), e0 = line_end(t, first + i);
                if (lane_id() == 0) { K[i] = (uint64_t)__double_as_longlong(u); S[i] = s0; E[i] = e0; }
            }
            __syncwarp();
            // insertion sort by key (k <= 511; ties are not expected and keep list order)
            if (lane_id() == 0) for (uint32_t i = 1; i < k; i++) { uint64_t kk = K[i]; uint32_t ss = S[i], ee = E[i]; int j = (int)i - 1; while (j >= 0 && K[j] > kk) { K[j + 1] = K[j]; S[j + 1] = S[j]; E[j + 1] = E[j]; j--; } K[j + 1] = kk; S[j + 1] = ss; E[j + 1] = ee; }
        }
        __syncwarp();
        uint32_t w = 0;
        for (uint32_t i = 0; i < k; i++) { uint32_t s0 = S[i], e0 = E[i]; warp_copy(out + w, p + s0, e0 - s0); w += e0 - s0; }
        __syncwarp();
        t_push(ws, seg_copy(p, ra)); t_push(ws, seg_copy(out, rb - ra)); t_push(ws, seg_copy(p + rb, n - rb)); break;
    }
    }
}
// construct_st_line_muta :364-378 with st_list_ins / st_list_replace, reference src/erlamsa_generic.erl:122-162
EB_DEV void mut_st_line(CaseCtx& c, int id, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0;
    LineTab t;
    if (!try_lines(c, p, n, t)) { r.kind = RES_SAME; r.delta = -1; return; }
    int w = id == M_LIS ? 0 : 1;
    uint32_t len = t.nlines;
    // step_state: fill up to 10 stored lines (each new one goes to the front), then maybe replace one
    while (ws->st_n[w] < 10) {
        uint32_t q = (uint32_t)g.erand(len) - 1; uint32_t a = line_start(t, q), b = line_end(t, q);
        for (int i = ws->st_n[w]; i > 0; i--) ws->st[w][i] = ws->st[w][i - 1];
        StSlot s; s.hp = p + a; s.hl = 1; s.tp = p + a + 1; s.tl = b - a - 1;
        ws->st[w][0] = s; ws->st_n[w]++;
    }
    uint32_t up = (uint32_t)g.erand(20);
    if (up < 10) {
        uint32_t q = (uint32_t)g.erand(len) - 1; uint32_t a = line_start(t, q), b = line_end(t, q);
        ws->st[w][up - 1].hp = p + a; ws->st[w][up - 1].hl = b - a;   // [New | tl(Old)]
    }
    uint32_t pk = (uint32_t)g.erand((uint64_t)ws->st_n[w]) - 1;
    StSlot x = ws->st[w][pk];
    uint32_t q = (uint32_t)g.erand(len) - 1;
    uint32_t a = line_start(t, q), b = line_end(t, q);
    t_reset(ws);
    t_push(ws, seg_copy(p, a)); t_push(ws, seg_copy(x.hp, x.hl)); t_push(ws, seg_copy(x.tp, x.tl));
    if (id == M_LIS) t_push(ws, seg_copy(p + a, n - a)); else t_push(ws, seg_copy(p + b, n - b));
    r.delta = 1; r.kind = RES_SEGS;
}

}  // namespace eb
#include "eb_mut_text.cuh"
#include "eb_mut_tree.cuh"
#include "eb_mut_fuse.cuh"
#include "eb_field.cuh"
#include "eb_mut_nested.cuh"
namespace eb {

// mutators whose working tables live in the per-warp temp arena
__host__ __device__ inline bool mut_needs_temp(int id) {
    switch (id) {
    case M_AB: case M_AD: case M_TR2: case M_TD: case M_TS1: case M_TS2: case M_TR: case M_FT: case M_FN: case M_FO: case M_LEN:
    case M_URI: case M_B64: case M_SGM: case M_JS: return true;
    default: return false;
    }
}

// which mutators have a device implementation
__host__ EB_DEV bool mut_supported(int id) {
    switch (id) {
    case M_AB: case M_AD: case M_TR2: case M_TD: case M_TS1: case M_TS2: case M_TR: case M_FT: case M_FN: case M_FO:
    case M_LEN: case M_ZIP: case M_URI: case M_B64: case M_SGM: case M_JS:   // sgm / js: refusal and single-token paths; the rest flags the case
    case M_UW: case M_UI: case M_NUM:
    case M_BD: case M_BEI: case M_BED: case M_BF: case M_BI: case M_BER: case M_BR:
    case M_SP: case M_SR: case M_SD: case M_SNAND: case M_SRND:
    case M_LD: case M_LDS: case M_LR2: case M_LRI: case M_LR: case M_LS: case M_LP: case M_LIS: case M_LRS:
    case M_NIL: return true;
    default: return false;
    }
}

// the mutators that never run a nested scheduler; false when `row` is not one of them
__device__ __forceinline__ bool mut_apply_base(CaseCtx& c, MutRow& row, const uint8_t* p, uint32_t n, MutResult& r) {
    int id = row.fn;
    switch (id) {
    case M_BD: case M_BEI: case M_BED: case M_BF: case M_BI: case M_BER: case M_BR: case M_UW: case M_UI: mut_byte(c, id, p, n, r); return true;
    case M_SP: case M_SR: case M_SD: case M_SNAND: case M_SRND: mut_bytes(c, id, p, n, r); return true;
    case M_NUM: mut_num(c, p, n, r); return true;
    case M_AB: case M_AD: mut_ascii(c, id, p, n, r); return true;
    case M_TR2: case M_TD: case M_TS1: case M_TS2: case M_TR: mut_tree(c, id, p, n, r); return true;
    case M_FT: case M_FN: case M_FO: mut_fuse(c, id, p, n, r); return true;
    case M_LEN: mut_len(c, p, n, r); return true;
    case M_URI: mut_uri(c, row, p, n, r); return true;
    case M_ZIP: {   // zip_path_traversal :1149-1163 on data that is not a ZIP archive: zip:foldl fails, delta -1, no draws
        bool z = false; { uint32_t hit = 0; for (uint32_t i = lane_id(); i + 4 <= n; i += 32) hit |= (p[i] == 'P' && p[i + 1] == 'K' && p[i + 2] == 5 && p[i + 3] == 6) ? 1u : 0u; z = __any_sync(0xffffffffu, hit != 0); }
        r.kind = z ? RES_UNSUPPORTED : RES_SAME; r.delta = -1; r.rechunk = 0; r.consumed_next = 0; return true;
    }
    case M_LD: case M_LDS: case M_LR2: case M_LRI: case M_LR: case M_LS: case M_LP: mut_line(c, id, p, n, r); return true;
    case M_LIS: case M_LRS: mut_st_line(c, id, p, n, r); return true;
    case M_NIL: r.kind = RES_SAME; r.delta = -1; r.rechunk = 0; r.consumed_next = 0; return true;
    default: return false;
    }
}
// dispatch at nesting level LVL (0 = the case's own scheduler; eb_mut_nested.cuh runs rounds at 1 and 2). The three
// mutators that can open a nested round are instantiated per level so that the call graph stays acyclic.
template <int LVL>
EB_DEV void mut_apply_level(CaseCtx& c, MutRow& row, const uint8_t* p, uint32_t n, MutResult& r) {
    if (mut_apply_base(c, row, p, n, r)) return;
    if (row.fn == M_B64) { mut_b64<LVL>(c, p, n, r); return; }
    if (row.fn == M_JS) { mut_js<LVL>(c, p, n, r); return; }
    if (row.fn == M_SGM) { mut_sgm<LVL>(c, p, n, r); return; }
    r.kind = RES_UNSUPPORTED; r.delta = 0; r.rechunk = 0; r.consumed_next = 0;
}
EB_DEV void mut_apply(CaseCtx& c, MutRow& row, const uint8_t* p, uint32_t n, MutResult& r) { mut_apply_level<0>(c, row, p, n, r); }

// The LIGHT kernel flavour: byte / sequence / number / line / utf-8 mutators only -- no lexer, trees, fuse, field
// search or nested rounds in its call graph, so its static stack and its instruction footprint stay small. The host
// picks it when the selected mutators and patterns allow (eb_engine.cu); results are identical by construction.
__host__ __device__ inline bool mut_is_light(int id) {
    switch (id) {
    case M_BD: case M_BEI: case M_BED: case M_BF: case M_BI: case M_BER: case M_BR: case M_UW: case M_UI:
    case M_SP: case M_SR: case M_SD: case M_SNAND: case M_SRND: case M_NUM:
    case M_LD: case M_LDS: case M_LR2: case M_LRI: case M_LR: case M_LS: case M_LP: case M_LIS: case M_LRS: case M_NIL: return true;
    default: return false;
    }
}
EB_DEV void mut_apply_light(CaseCtx& c, MutRow& row, const uint8_t* p, uint32_t n, MutResult& r) {
    int id = row.fn;
    switch (id) {
    case M_BD: case M_BEI: case M_BED: case M_BF: case M_BI: case M_BER: case M_BR: case M_UW: case M_UI: mut_byte(c, id, p, n, r); return;
    case M_SP: case M_SR: case M_SD: case M_SNAND: case M_SRND: mut_bytes(c, id, p, n, r); return;
    case M_NUM: mut_num(c, p, n, r); return;
    case M_LD: case M_LDS: case M_LR2: case M_LRI: case M_LR: case M_LS: case M_LP: mut_line(c, id, p, n, r); return;
    case M_LIS: case M_LRS: mut_st_line(c, id, p, n, r); return;
    case M_NIL: r.kind = RES_SAME; r.delta = -1; r.rechunk = 0; r.consumed_next = 0; return;
    default: r.kind = RES_UNSUPPORTED; r.delta = 0; r.rechunk = 0; r.consumed_next = 0; return;
    }
}

}  // namespace eb
