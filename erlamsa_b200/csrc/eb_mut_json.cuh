// erlamsa_b200 -- `js` on documents with arrays / objects (reference src/erlamsa_json.erl:297-731).
//
// Same idea as eb_mut_sgml.cuh: fold_ast/1 re-serialises the AST as the concatenation of its atoms in source order
// (scalars, brackets, commas, colons -- white space is gone), an AST element is an atom range, and the structural
// mutations are re-orderings / repetitions of ranges. What the reference's list juggling adds on top:
//   * two numberings: walk(all, ..) counts a pair, its key and its value; walk(pairs, ..) / select(values, ..) skip the
//     key subtree entirely (:297-345, :347-399). count/1's first result NV is the size of the second numbering;
//   * an element that turns into several (dup / repeat / insert) stays comma-joined inside an array or object, but as
//     the value of a pair or as the whole document it folds as a bracketed list (walk_uncons1 :293-295, fold_ast :277-279);
//   * json_pump always pumps twice (:560): prefix^4 ++ element ++ suffix^4.
#pragma once

namespace eb {

enum { JE_CONT = 0, JE_PAIR, JE_SCALAR };
enum { JS_TOP = 0, JS_LIST, JS_PVAL, JS_PKEY };
struct JEl { uint32_t lo, hi; uint32_t val_no; uint32_t ct_no; uint32_t kind; uint32_t slot; };   // 'all' number = index + 1
struct JOver { const uint8_t* ptr; uint32_t len; uint32_t on; };

struct JDoc {
    const uint8_t* S; uint32_t n;
    JAtom* at; uint32_t nat;
    JEl* el; uint32_t nel;
    JOver* ov;
    uint32_t N, NT, NV;
};

// elements in pre-order from the atom sequence. false: tables full
EB_DEV bool js_build(CaseCtx& c, JDoc& d) {
    struct Fr { uint32_t el; uint32_t is_obj; uint32_t in_key; uint32_t pair; uint32_t phase; uint32_t as_key; };   // phase (objects): 0 key next, 1 value next
    uint32_t fcap = d.nat / 2 + 2; if (fcap > 8192) fcap = 8192;
    Fr* st = (Fr*)temp_alloc(c, (uint64_t)fcap * sizeof(Fr));
    d.el = (JEl*)temp_alloc(c, (uint64_t)(d.nat + 2) * sizeof(JEl));
    if (!st || !d.el) return false;
    uint32_t sp = 0, ne = 0, V = 0, CT = 0;
    auto new_el = [&](uint32_t lo, uint32_t kind, uint32_t slot, bool in_key) -> uint32_t {
        JEl e; e.lo = lo; e.hi = lo; e.kind = kind; e.slot = slot; e.val_no = in_key ? 0u : ++V; e.ct_no = kind == JE_CONT ? ++CT : 0u;
        d.el[ne] = e; return ne++;
    };
    // a value (scalar or container) is complete at atom index `hi`
    auto value_done = [&](uint32_t hi) {
        if (sp == 0) return;
        Fr& f = st[sp - 1];
        if (!f.is_obj) return;
        if (f.phase == 0) f.phase = 1;                                   // that was the key
        else { d.el[f.pair].hi = hi; f.phase = 0; }                      // that was the value: the pair is complete
    };
    for (uint32_t i = 0; i < d.nat; i++) {
        uint32_t k = d.at[i].kind;
        if (k == JA_COMMA || k == JA_COLON) continue;
        if (k == JA_CARR || k == JA_COBJ) {
            if (sp == 0) return false;
            Fr f = st[--sp];
            d.el[f.el].hi = i; d.at[d.el[f.el].lo].match = i; d.at[i].match = d.el[f.el].lo;
            value_done(i);
            continue;
        }
        // a value starts here: which slot?
        uint32_t slot = JS_TOP; bool in_key = false;
        if (sp) {
            Fr& f = st[sp - 1]; in_key = f.in_key != 0;
            if (!f.is_obj) slot = JS_LIST;
            else if (f.phase == 0) { f.pair = new_el(i, JE_PAIR, JS_LIST, in_key); slot = JS_PKEY; }
            else slot = JS_PVAL;
        }
        bool key_ctx = in_key || slot == JS_PKEY;
        if (k == JA_OARR || k == JA_OOBJ) {
            if (sp >= fcap) return false;
            Fr f; f.el = new_el(i, JE_CONT, slot, key_ctx); f.is_obj = k == JA_OOBJ; f.in_key = key_ctx; f.pair = 0; f.phase = 0; f.as_key = slot == JS_PKEY;
            st[sp++] = f;
            continue;
        }
        (void)new_el(i, JE_SCALAR, slot, key_ctx);
        value_done(i);
    }
    d.nel = ne; d.N = ne; d.NT = CT; d.NV = V;
    __syncwarp();
    return sp == 0;
}

__device__ const char* const c_js_null_alt[7] = {"-1", "1000000000", "true", "[]", "\"%n%s\"", "0", "\"AAAAAAAAAAAA\""};   // mutate_null/2 :643

struct JEmit { Bld b; const BatchParams* bp; };
EB_DEV void js_emit_atom(JEmit& e, const JDoc& d, uint32_t i) {
    const JAtom& a = d.at[i];
    if (d.ov && d.ov[i].on) {
        if (a.kind == JA_STR) { bld_put(e.b, '"'); bld_copy(e.b, d.ov[i].ptr, d.ov[i].len); bld_put(e.b, '"'); }
        else bld_copy(e.b, d.ov[i].ptr, d.ov[i].len);
        return;
    }
    switch (a.kind) {
    case JA_STR: bld_copy(e.b, d.S + a.a - 1, a.b - a.a + 2); break;
    case JA_JUNK: bld_copy(e.b, d.S + a.a - 1, d.n - (a.a - 1)); bld_put(e.b, '"'); bld_put(e.b, '"'); break;
    default: bld_copy(e.b, d.S + a.a, a.b - a.a); break;                  // scalars verbatim; brackets, commas and colons are their own text
    }
}
// a run of atoms, 32 per step: every lane sizes its own atom (override or source text, quotes included), one warp scan gives the
// output offsets, short atoms are copied by their lane and long ones by the whole warp -- the serial walk paid a dependent
// table load and a warp copy of a few bytes per atom, twice (sizing pass, writing pass).
EB_DEV void js_emit_range(JEmit& e, const JDoc& d, uint32_t lo, uint32_t hi) {
    if (d.nat == 0 || lo >= d.nat || lo > hi) return;
    if (hi >= d.nat) hi = d.nat - 1;
    const int l = lane_id();
    Bld& b = e.b;
    for (uint32_t i0 = lo; i0 <= hi; i0 += 32) {
        const uint32_t i = i0 + (uint32_t)l; const bool act = i <= hi && i >= i0;
        const uint8_t* src = nullptr; uint32_t len = 0, pre = 0, post = 0;
        if (act) {
            const JAtom a = d.at[i];
            if (d.ov && d.ov[i].on) { src = d.ov[i].ptr; len = d.ov[i].len; if (a.kind == JA_STR) { pre = 1; post = 1; } }
            else if (a.kind == JA_STR) { src = d.S + a.a - 1; len = a.b - a.a + 2; }
            else if (a.kind == JA_JUNK) { src = d.S + a.a - 1; len = d.n - (a.a - 1); post = 2; }
            else { src = d.S + a.a; len = a.b - a.a; }                  // scalars verbatim; brackets, commas and colons are their own text
        }
        unsigned long long incl = (unsigned long long)pre + len + post;
        const unsigned long long mine = incl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long x = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += x; }
        const unsigned long long step = __shfl_sync(0xffffffffu, incl, 31);
        if ((unsigned long long)b.n + step > 0x7fffffffull) { b.ovf = 1; return; }
        if (b.p) {
            const uint32_t at = b.n + (uint32_t)(incl - mine);
            constexpr uint32_t SHORT = 48;
            if (act) {
                if (pre && at < b.cap) b.p[at] = '"';
                if (len <= SHORT) for (uint32_t k = 0; k < len; k++) if (at + pre + k < b.cap) b.p[at + pre + k] = src[k];
                for (uint32_t k = 0; k < post; k++) if (at + pre + len + k < b.cap) b.p[at + pre + len + k] = '"';
            }
            uint32_t longs = __ballot_sync(0xffffffffu, act && len > SHORT);
            while (longs) {
                const int t = __ffs(longs) - 1; longs &= longs - 1;
                const uint8_t* sp = (const uint8_t*)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)src, t);
                const uint32_t sl = __shfl_sync(0xffffffffu, len, t), sa = __shfl_sync(0xffffffffu, at + pre, t);
                for (uint32_t k = (uint32_t)l; k < sl; k += 32) if (sa + k < b.cap) b.p[sa + k] = sp[k];
            }
        }
        b.n += (uint32_t)step;
        if (hi - i0 < 32) break;                                          // (also keeps i0 += 32 from wrapping)
    }
    __syncwarp();
}

struct JPlan { int which; uint32_t a, b; uint32_t times; uint32_t x; };   // a, b, x: element indices
EB_DEV void js_emit_plan(JEmit& e, const JDoc& d, const JPlan& pl) {
    const uint32_t last = d.nat - 1;
    auto all = [&](uint32_t lo, uint32_t hi) { if (lo <= hi && hi != 0xffffffffu) js_emit_range(e, d, lo, hi); };
    switch (pl.which) {
    case 0: {                                                            // json_swap :573-586
        uint32_t alo = d.el[pl.a].lo, ahi = d.el[pl.a].hi, blo = d.el[pl.b].lo, bhi = d.el[pl.b].hi;
        if (pl.a == pl.b) { all(0, last); return; }
        bool b_in_a = blo >= alo && bhi <= ahi, a_in_b = alo >= blo && ahi <= bhi;
        if (b_in_a) { if (alo) all(0, alo - 1); all(blo, bhi); all(ahi + 1, last); return; }
        if (a_in_b) { if (blo) all(0, blo - 1); all(alo, ahi); all(bhi + 1, last); return; }
        if (alo < blo) { if (alo) all(0, alo - 1); all(blo, bhi); all(ahi + 1, blo - 1); all(alo, ahi); all(bhi + 1, last); }
        else { if (blo) all(0, blo - 1); all(alo, ahi); all(bhi + 1, alo - 1); all(blo, bhi); all(ahi + 1, last); }
        return;
    }
    case 1: case 3: case 4: {                                            // dup / repeat / insert: Elem, then the copies (or NewElem)
        const JEl& t = d.el[pl.which == 4 ? pl.b : pl.a];                // the element at whose place the list grows
        const JEl& src = d.el[pl.a];
        bool wrap = t.slot == JS_PVAL || t.slot == JS_TOP;
        if (t.lo) all(0, t.lo - 1);
        if (wrap) bld_put(e.b, '[');
        all(t.lo, t.hi);
        for (uint32_t k = 0; k < pl.times; k++) { bld_put(e.b, ','); all(src.lo, src.hi); }
        if (wrap) bld_put(e.b, ']');
        all(t.hi + 1, last); return;
    }
    case 2: {                                                            // json_pump :554-563
        if (pl.times == 0) { all(0, last); return; }
        const JEl& s0 = d.el[pl.a]; const JEl& x = d.el[pl.x];
        if (s0.lo) all(0, s0.lo - 1);
        for (uint32_t k = 0; k < pl.times; k++) if (x.lo > s0.lo) all(s0.lo, x.lo - 1);
        all(x.lo, x.hi);
        for (uint32_t k = 0; k < pl.times; k++) all(x.hi + 1, s0.hi);
        all(s0.hi + 1, last); return;
    }
    default: all(0, last); return;
    }
}

template <int LVL>
EB_DEV void mut_js_document(CaseCtx& c, const uint8_t* p, uint32_t n, JAtom* atoms, uint32_t natoms, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    JDoc d; d.S = p; d.n = n; d.at = atoms; d.nat = natoms; d.el = nullptr; d.nel = 0; d.ov = nullptr;
    if (!js_build(c, d)) { r.kind = RES_UNSUPPORTED; return; }
    const uint32_t N = d.N, NT = d.NT, NV = d.NV;
    JPlan pl; pl.which = (int)g.rand(21); pl.a = pl.b = pl.x = 0; pl.times = 0;
    double D = 1;
    auto by_val = [&](uint32_t want) -> uint32_t { for (uint32_t i = 0; i < d.nel; i++) if (d.el[i].val_no == want) return i; return 0xffffffffu; };
    bool payload = false; uint8_t* pbuf = nullptr; uint32_t plen = 0;
    switch (pl.which) {
    case 0: case 4: {
        uint32_t r1 = (uint32_t)g.erand(NV), r2 = (uint32_t)g.erand(NV);
        pl.a = by_val(r1); pl.b = by_val(r2);
        if (pl.a == 0xffffffffu || pl.b == 0xffffffffu) { ws->status = CASE_DIED; return; }
        pl.times = 1;
        break;
    }
    case 1: case 3: {
        uint32_t rr = (uint32_t)g.erand(NV);
        pl.times = pl.which == 1 ? 1u : (uint32_t)g.erand(100);
        pl.a = by_val(rr);
        if (pl.a == 0xffffffffu) { ws->status = CASE_DIED; return; }
        break;
    }
    case 2: {
        D = -2;
        uint32_t rr = (uint32_t)g.erand(NT);
        uint32_t s0 = 0xffffffffu;
        for (uint32_t i = 0; i < d.nel; i++) if (d.el[i].ct_no == rr) { s0 = i; break; }
        if (s0 == 0xffffffffu) { ws->status = CASE_DIED; return; }
        uint32_t sub = 0; for (uint32_t i = s0; i < d.nel && d.el[i].lo <= d.el[s0].hi; i++) sub++;
        uint32_t e = (uint32_t)g.erand(sub - 1) + 1;
        pl.a = s0; pl.x = s0 + e - 1; pl.times = e == 1 ? 0u : 4u;       // PumpCnt = 2
        break;
    }
    case 5: {                                                            // make_json_unserialize :615-618
        const char* f = c_js_payload[g.rand_elem_idx(6)];
        pbuf = scratch_alloc(c, 1024);
        if (!pbuf) { r.delta = 0; return; }
        Bld b; b.p = pbuf; b.n = 0; b.cap = 1024; b.ovf = 0;
        for (int q = 0; f[q]; q++) {
            if (f[q] == '~' && f[q + 1] == 's') { bld_puts(b, "://"); bld_hostport(b, c.bp); bld_put(b, '/'); q++; }
            else bld_put(b, (uint8_t)f[q]);
        }
        __syncwarp();
        if (b.n > b.cap) { r.kind = RES_UNSUPPORTED; return; }
        payload = true; plen = b.n; D = -2;
        break;
    }
    default: {                                                           // inner values :670-718
        if (LVL >= MAX_NEST) { r.kind = RES_UNSUPPORTED; return; }
        const int kind0 = c.snand_kind;
        const int inner_kind = (int)g.rand_elem_idx(3); (void)g.rand_elem_idx(1);
        const uint8_t ids[9] = {M_SGM, M_AB, M_AD, M_NUM, M_SP, M_SR, M_SD, M_B64, M_URI};
        MutRow* rows = (MutRow*)temp_alloc(c, sizeof(MutRow) * 9);
        d.ov = (JOver*)temp_alloc(c, (uint64_t)d.nat * sizeof(JOver));
        if (!rows || !d.ov) { r.delta = 0; return; }
        for (uint32_t i = lane_id(); i < d.nat; i += 32) d.ov[i].on = 0;
        __syncwarp();
        int nr = inner_table(c, ids, 9, false, rows);
        const uint64_t floor0 = c.temp_floor, keep = c.temp_used;
        const double pn = 3.0 / (double)N, pk = 0.6 / (double)N;
        // walk2acc visits everything in source order; the element table tells which scalars are string keys
        for (uint32_t ei = 0; ei < d.nel && ws->status == CASE_OK; ei++) {
            const JEl& el = d.el[ei];
            if (el.kind != JE_SCALAR) continue;
            uint32_t ai = el.lo; const JAtom& a = d.at[ai];
            c.temp_used = keep;
            if (a.kind == JA_STR) {
                double rnd = g.uniform();
                if (rnd > (el.slot == JS_PKEY ? pk : pn)) continue;
                c.temp_floor = c.temp_used; c.snand_kind = inner_kind;
                InnerRes res; res.kind = 0; res.len = a.b - a.a;
                if constexpr (LVL < MAX_NEST) res = inner_round<LVL>(c, rows, nr, p + a.a, a.b - a.a, true);
                c.snand_kind = kind0; c.temp_floor = floor0;
                if (ws->status != CASE_OK) break;
                if (res.kind == 0) continue;
                uint8_t* lit = scratch_alloc(c, res.len);
                if (!lit) { r.delta = 0; return; }
                inner_result_write(c, res, p + a.a, lit);
                d.ov[ai].ptr = lit; d.ov[ai].len = res.len; d.ov[ai].on = 1;
            } else if (a.kind == JA_NULL) {                              // mutate_null/2 :640-643
                double rnd = g.uniform();
                if (rnd >= pn) continue;
                const char* t = c_js_null_alt[g.rand_elem_idx(7)];
                uint32_t tl = 0; while (t[tl]) tl++;
                uint8_t* lit = scratch_alloc(c, tl);
                if (!lit) { r.delta = 0; return; }
                if (lane_id() < (int)tl) lit[lane_id()] = (uint8_t)t[lane_id()];
                __syncwarp();
                d.ov[ai].ptr = lit; d.ov[ai].len = tl; d.ov[ai].on = 1;
            } else if (a.kind == JA_TRUE || a.kind == JA_FALSE) {         // basic_type_mutation(Boolean, Prob)
                double rnd = g.uniform();
                if (rnd >= pn) continue;
                const char* t = a.kind == JA_TRUE ? "false" : "true"; uint32_t tl = a.kind == JA_TRUE ? 5 : 4;
                uint8_t* lit = scratch_alloc(c, tl);
                if (!lit) { r.delta = 0; return; }
                if (lane_id() < (int)tl) lit[lane_id()] = (uint8_t)t[lane_id()];
                __syncwarp();
                d.ov[ai].ptr = lit; d.ov[ai].len = tl; d.ov[ai].on = 1;
            } else if (a.kind == JA_NUM) {                               // list_to_integer/1 or no draw at all :700-709
                uint32_t q = a.a; bool neg = false;
                if (q < a.b && (p[q] == '+' || p[q] == '-')) { neg = p[q] == '-'; q++; }
                bool ok = q < a.b;
                for (uint32_t z = q; z < a.b; z++) if ((uint32_t)(p[z] - '0') >= 10u) ok = false;
                if (!ok) continue;
                double rnd = g.uniform();
                if (rnd >= pn) continue;
                if (a.b - q > 77) { r.kind = RES_UNSUPPORTED; return; }
                Big256 v; v.zero();
                for (uint32_t z = q; z < a.b; z++) { v.mul_small(10); v.add_small((uint32_t)(p[z] - '0')); }
                if (neg && !v.is_zero()) v.neg = 1;
                Big256 o = v;
                mutate_num(g, v);
                if (v.ovf) { r.kind = RES_UNSUPPORTED; return; }
                bool same = v.neg == o.neg; for (int z = 0; z < 8; z++) same = same && v.m[z] == o.m[z];
                if (same) continue;
                uint8_t dec[88]; int dl = v.to_decimal(dec);
                uint8_t* lit = scratch_alloc(c, (uint64_t)dl);
                if (!lit) { r.delta = 0; return; }
                for (int z = lane_id(); z < dl; z += 32) lit[z] = dec[z];
                __syncwarp();
                d.ov[ai].ptr = lit; d.ov[ai].len = (uint32_t)dl; d.ov[ai].on = 1;
            }
        }
        c.temp_used = keep; c.temp_floor = floor0;
        if (ws->status != CASE_OK || r.kind == RES_UNSUPPORTED) return;
        pl.which = 21;
        break;
    }
    }
    uint8_t* out = pbuf; uint32_t total = plen;
    if (!payload) {
        JEmit em; em.bp = c.bp; em.b.p = nullptr; em.b.n = 0; em.b.cap = 0; em.b.ovf = 0;
        js_emit_plan(em, d, pl);
        total = em.b.n;
        if (em.b.ovf || total > c.bp->max_case_out) { ws->status = CASE_OVERFLOW; ws->reason = 5; return; }
        out = scratch_alloc(c, (uint64_t)total + 1);
        if (!out) { r.delta = 0; return; }
        em.b.p = out; em.b.n = 0; em.b.cap = total;
        js_emit_plan(em, d, pl);
        __syncwarp();
    }
    bool same = total == n && range_equal(out, p, n);
    if (same) { r.kind = RES_SAME; r.delta = -1; return; }               // NewBinStr =:= H :723-724
    t_reset(ws); t_push(ws, seg_copy(out, total));
    r.kind = RES_SEGS; r.delta = D + trunc((double)total / (double)(AVG_BLOCK_SIZE * 10));
}

}  // namespace eb
