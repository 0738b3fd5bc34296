// erlamsa_b200 -- the SGML/XML mutator `sgm` on the device (reference src/erlamsa_sgml.erl).
//
// The reference tokenizes (:65-176), builds an AST with recovery rules for unpaired / early-closed tags (:187-279),
// numbers the elements in pre-order (:344-404), applies one of twelve mutations (:478-737) and folds the AST back to
// text (:290-331). On the device the AST is never built, because of three facts about that pipeline:
//   * AST pre-order IS document order, and what build_ast2/4 decides is only which open tags are PAIRED with which
//     close tag: a stack discipline over the token stream (a close tag pairs with the nearest open tag of the same
//     lower-cased name, every open tag in between stays unpaired, as does everything still open at the end);
//   * the fold of an element is the concatenation of the folds of its tokens (a paired tag's subtree = the token
//     range [open, close]), so element n of walk/3 is a token range and the structural mutations (swap, dup, repeat,
//     insert, pump, break) are re-orderings / repetitions of token ranges;
//   * pump_path/3 (:496-508), which re-inserts the growing tree at index 2E-1, 4E-3, ..., yields
//     prefix^(2^k) ++ element ++ suffix^(2^k) with prefix/suffix the parts of the start tag around element E.
// The CPU restatement used by the tests keeps the AST form; the tests compare the two byte for byte.
#pragma once

namespace eb {

enum { ST_OPEN = 0, ST_SC, ST_CLOSE, ST_TEXT, ST_BANG, ST_COMMENT, ST_QUE, ST_EOFTEXT };
constexpr uint32_t SG_NOMATCH = 0xffffffffu;
struct STok { uint32_t kind; uint32_t len; const uint8_t* ptr; uint32_t par0, npar; uint32_t match; uint32_t flags; };   // ptr/len: tag name or payload
struct SPar { const uint8_t* np; const uint8_t* vp; uint32_t nl, vl; uint32_t q; uint32_t choice; };                  // q: 0, '\'' or '"'
constexpr uint32_t SF_MATCHED_CLOSE = 1u;   // a close tag that paired: not an element of its own
constexpr uint32_t SF_PAIRED = 2u;          // an open tag with a partner (match = index of its close token)
constexpr uint32_t SF_XMLNS = 4u;           // xmlns_modify/2 fired for this tag

struct SgDoc {
    const uint8_t* S; uint32_t n;
    STok* tok; uint32_t ntok, tok_cap;
    SPar* par; uint32_t npar, par_cap;
    uint32_t N, NT;
};

__device__ __forceinline__ bool sg_ws(uint32_t ch) { return ch == ' ' || ch == '\r' || ch == '\n' || ch == '\t'; }
__device__ __forceinline__ bool sg_ev(uint32_t ch) { return sg_ws(ch) || ch == '>' || ch == '='; }
__device__ __forceinline__ uint32_t sg_lower(uint32_t c) { return ((c >= 'A' && c <= 'Z') || (c >= 0xC0 && c <= 0xD6) || (c >= 0xD8 && c <= 0xDE)) ? c + 32 : c; }   // string:to_lower/1, Latin-1

// tz/2 (:82-148) from state {tag,""} at S[i..]: 0 = token written to *t (params appended to d.par), 1 = throw(incorrect_sgml),
// 2 = function_clause (unterminated comment), 3 = device table full
EB_DEV int sg_scan_tag(SgDoc& d, uint32_t i, STok* t, uint32_t* next) {
    const uint8_t* S = d.S; const uint32_t n = d.n;
    enum { TAGN, BANG, COMMENT, QUE, ETAG, ENDTAG, ENDTAG_GT, ATTR, EATT, VAL, SQVAL, DQVAL, UQVAL };
    int st = TAGN;
    uint32_t tag0 = i, tag1 = i, a0 = 0, a1 = 0, v0 = 0;
    const uint32_t par_base = d.npar; uint32_t np = 0;
    auto ws = [&](uint32_t q) { while (q < n && sg_ws(S[q])) q++; return q; };
    auto st2 = [&](uint32_t q, uint32_t c0, uint32_t c1) { return q + 2 <= n && S[q] == c0 && S[q + 1] == c1; };
    auto add_par = [&](uint32_t vs, uint32_t ve, uint32_t quote) -> bool {
        if (par_base + np >= d.par_cap) return false;
        SPar pr; pr.np = S + a0; pr.nl = a1 - a0; pr.vp = S + vs; pr.vl = ve - vs; pr.q = quote; pr.choice = 0;
        d.par[par_base + np] = pr; np++; return true;
    };
    auto done = [&](uint32_t kind, const uint8_t* ptr, uint32_t len, uint32_t nx) {
        t->kind = kind; t->ptr = ptr; t->len = len; t->par0 = par_base; t->npar = np; t->match = SG_NOMATCH; t->flags = 0; *next = nx;
        d.npar = par_base + np; return 0;
    };
    for (;;) {
        switch (st) {
        case TAGN:
            if (tag1 == tag0 && i < n) {
                if (i + 3 <= n && S[i] == '!' && S[i + 1] == '-' && S[i + 2] == '-') { st = COMMENT; i += 3; tag0 = i; continue; }
                if (S[i] == '!') { st = BANG; i = ws(i + 1); tag0 = i; continue; }
                if (S[i] == '?') { st = QUE; i = ws(i + 1); tag0 = i; continue; }
                if (S[i] == '/') { st = ENDTAG; i = ws(i + 1); tag0 = tag1 = i; continue; }
            }
            if (st2(i, '/', '>')) return done(ST_SC, S + tag0, tag1 - tag0, i + 2);
            if (i < n && sg_ev(S[i])) { st = ATTR; a0 = a1 = 0; i = ws(i); continue; }
            if (i < n) { i++; tag1 = i; continue; }
            return 1;
        case BANG: { uint32_t q = find_byte(S, i, n, '>'); if (q >= n) return 1; return done(ST_BANG, S + tag0, q - tag0, q + 1); }
        case COMMENT: { uint32_t q = i; for (;;) { q = find_byte(S, q, n, '-'); if (q >= n) return 2; if (q + 3 <= n && S[q + 1] == '-' && S[q + 2] == '>') return done(ST_COMMENT, S + tag0, q - tag0, q + 3); q++; } }
        case QUE: { uint32_t q = i; for (;;) { q = find_byte(S, q, n, '?'); if (q >= n) return 1; if (q + 2 <= n && S[q + 1] == '>') return done(ST_QUE, S + tag0, q - tag0, q + 2); q++; } }
        case ETAG:
            if (st2(i, '/', '>')) return done(ST_SC, S + tag0, tag1 - tag0, i + 2);
            if (i < n && S[i] == '>') return done(ST_OPEN, S + tag0, tag1 - tag0, i + 1);
            return 1;
        case ENDTAG:
            if (i < n && sg_ev(S[i])) { st = ENDTAG_GT; i = ws(i); continue; }
            if (i < n) { i++; tag1 = i; continue; }
            return 1;
        case ENDTAG_GT:
            if (i < n && S[i] == '>') return done(ST_CLOSE, S + tag0, tag1 - tag0, i + 1);
            return 1;
        case ATTR:
            if (a1 == a0 && ((i < n && sg_ev(S[i])) || st2(i, '/', '>'))) { st = ETAG; continue; }
            if ((i < n && sg_ev(S[i])) || st2(i, '/', '>')) { st = EATT; i = ws(i); continue; }
            if (i < n) { if (a1 == a0) a0 = i; i++; a1 = i; continue; }
            return 1;
        case EATT:
            if (i < n && S[i] == '=') { st = VAL; i = ws(i + 1); continue; }
            if (!add_par(0, 0, 0)) return 3;
            a0 = a1 = 0; st = ATTR; i = ws(i); continue;
        case VAL:
            if (i < n && S[i] == '\'') { st = SQVAL; i++; v0 = i; continue; }
            if (i < n && S[i] == '"') { st = DQVAL; i++; v0 = i; continue; }
            st = UQVAL; v0 = i; continue;
        case SQVAL: { uint32_t q = find_byte(S, i, n, '\''); if (q >= n) return 1; if (!add_par(v0, q, '\'')) return 3; a0 = a1 = 0; st = ATTR; i = ws(q + 1); continue; }
        case DQVAL: { uint32_t q = find_byte(S, i, n, '"'); if (q >= n) return 1; if (!add_par(v0, q, '"')) return 3; a0 = a1 = 0; st = ATTR; i = ws(q + 1); continue; }
        default:   // UQVAL
            if ((i < n && sg_ev(S[i])) || st2(i, '/', '>')) { if (!add_par(v0, i, 0)) return 3; a0 = a1 = 0; st = ATTR; i = ws(i); continue; }
            if (i < n) { i++; continue; }
            return 1;
        }
    }
}

// tokenize/1 (:65-96). 0 ok, 1 throw, 2 error, 3 tables full / text with more skipped gaps than the device keeps
EB_DEV int sg_tokenize(CaseCtx& c, SgDoc& d) {
    const uint8_t* S = d.S; const uint32_t n = d.n;
    uint32_t lt = find_byte(S, 0, n, '<');
    if (lt >= n) return 1;
    auto ws = [&](uint32_t q) { while (q < n && sg_ws(S[q])) q++; return q; };
    STok cur; uint32_t p = 0;
    int k = sg_scan_tag(d, ws(lt + 1), &cur, &p);
    if (k) { d.npar = 0; return k; }
    for (;;) {
        // text up to the next tag that scans; a tag that does not is kept as text, minus the white space behind its '<' (:74-91)
        uint32_t piece0[8], piece1[8]; int np = 0; uint32_t pstart = p;
        STok t2; uint32_t nx = 0; bool eof = false;
        for (;;) {
            uint32_t q = find_byte(S, p, n, '<');
            if (q >= n) { eof = true; break; }
            uint32_t e = ws(q + 1);
            const uint32_t par_mark = d.npar;
            int kk = sg_scan_tag(d, e, &t2, &nx);
            if (kk == 3) return 3;
            if (kk == 0) { p = q; break; }                               // text = [pstart, q)
            d.npar = par_mark;
            if (e > q + 1) {                                             // "<" stays, the white space after it is lost
                if (np >= 8) return 3;
                piece0[np] = pstart; piece1[np] = q + 1; np++; pstart = e;
            }
            p = e;
        }
        uint32_t tend = eof ? n : p;
        if (d.ntok + 2 >= d.tok_cap) return 3;
        d.tok[d.ntok++] = cur;
        STok tx; tx.kind = eof ? ST_EOFTEXT : ST_TEXT; tx.par0 = 0; tx.npar = 0; tx.match = SG_NOMATCH; tx.flags = 0;
        if (np == 0) { tx.ptr = S + pstart; tx.len = tend - pstart; }
        else {
            uint32_t tot = tend - pstart; for (int j = 0; j < np; j++) tot += piece1[j] - piece0[j];
            uint8_t* buf = scratch_alloc(c, tot);
            if (!buf) return 3;
            uint32_t o = 0;
            for (int j = 0; j < np; j++) { warp_copy(buf + o, S + piece0[j], piece1[j] - piece0[j]); o += piece1[j] - piece0[j]; }
            warp_copy(buf + o, S + pstart, tend - pstart);
            __syncwarp();
            tx.ptr = buf; tx.len = tot;
        }
        d.tok[d.ntok++] = tx;
        if (eof) return 0;
        cur = t2; p = nx;
    }
}

// is token i an element of the AST? (an empty text is dropped :238-240, a paired close is part of its tag)
__device__ __forceinline__ bool sg_is_elem(const STok& t) {
    if (t.kind == ST_CLOSE) return !(t.flags & SF_MATCHED_CLOSE);
    if (t.kind == ST_TEXT || t.kind == ST_EOFTEXT) return t.len > 0;
    return true;
}
EB_DEV bool sg_name_eq_lower(const STok& a, const STok& b) {
    if (a.len != b.len) return false;
    for (uint32_t i = 0; i < a.len; i++) if (sg_lower(a.ptr[i]) != sg_lower(b.ptr[i])) return false;
    return true;
}
// build_ast2/4 (:187-279) as pairing decisions; fills N and NT. false: nesting deeper than the device stack
EB_DEV bool sg_pair(CaseCtx& c, SgDoc& d) {
    uint32_t cap = d.ntok < 4096u ? d.ntok + 1 : 4096u;
    uint32_t* stk = (uint32_t*)temp_alloc(c, (uint64_t)cap * 4);
    if (!stk) return false;
    uint32_t sp = 0; d.N = 0; d.NT = 0;
    for (uint32_t i = 0; i < d.ntok; i++) {
        STok t = d.tok[i];
        if (t.kind == ST_OPEN) { if (sp >= cap) return false; stk[sp++] = i; d.N++; continue; }
        if (t.kind == ST_CLOSE) {
            uint32_t j = sp;
            while (j > 0 && !sg_name_eq_lower(d.tok[stk[j - 1]], t)) j--;
            if (j == 0) { d.N++; continue; }                             // {close, Tag} element
            uint32_t o = stk[j - 1];
            d.tok[o].match = i; d.tok[o].flags |= SF_PAIRED; d.tok[i].flags |= SF_MATCHED_CLOSE; d.tok[i].match = o;
            sp = j - 1; d.NT++;
            continue;
        }
        if (sg_is_elem(t)) d.N++;
    }
    __syncwarp();
    return true;
}

// element number (1-based, pre-order) -> token index; tag number -> token index. 32 tokens per step: ballot of the
// predicate, the wanted one is the (want - seen)-th set bit of the step that crosses it.
template <typename Pred>
__device__ __forceinline__ uint32_t sg_nth_tok(const SgDoc& d, uint32_t want, Pred pred) {
    if (want == 0) return SG_NOMATCH;
    const int l = lane_id(); uint32_t seen = 0;
    for (uint32_t i0 = 0; i0 < d.ntok; i0 += 32) {
        uint32_t i = i0 + (uint32_t)l;
        uint32_t m = __ballot_sync(0xffffffffu, i < d.ntok && pred(d.tok[i]));
        uint32_t k = (uint32_t)__popc(m);
        if (seen + k >= want) {
            for (uint32_t t = seen + 1; t < want; t++) m &= m - 1;
            return i0 + (uint32_t)__ffs(m) - 1;
        }
        seen += k;
    }
    return SG_NOMATCH;
}
EB_DEV uint32_t sg_elem_tok(const SgDoc& d, uint32_t want) { return sg_nth_tok(d, want, [](const STok& t) { return sg_is_elem(t); }); }
EB_DEV uint32_t sg_tag_tok(const SgDoc& d, uint32_t want) {
    return sg_nth_tok(d, want, [](const STok& t) { return t.kind == ST_OPEN && (t.flags & SF_PAIRED) != 0; });
}

__device__ __forceinline__ uint32_t sg_range_hi(const SgDoc& d, uint32_t i) { return (d.tok[i].kind == ST_OPEN && (d.tok[i].flags & SF_PAIRED)) ? d.tok[i].match : i; }

// ---------------------------------------------------------------- folder (:290-331), per token
struct SgEmit { Bld b; const BatchParams* bp; };
// does [v, v+len) spell "http://Host:Port/" ?
EB_DEV bool sg_is_ssrf_uri(const BatchParams* bp, const uint8_t* v, uint32_t len) {
    const char* pre = "http://"; uint32_t k = 0;
    for (int i = 0; pre[i]; i++, k++) if (k >= len || v[k] != (uint8_t)pre[i]) return false;
    for (int i = 0; i < 64 && bp->ssrf_host[i]; i++, k++) if (k >= len || v[k] != (uint8_t)bp->ssrf_host[i]) return false;
    if (k >= len || v[k] != ':') return false;
    k++;
    char tmp[12]; int nd = 0; int pv = bp->ssrf_port; if (pv < 0) return false;
    do { tmp[nd++] = (char)('0' + pv % 10); pv /= 10; } while (pv);
    while (nd) { if (k >= len || v[k] != (uint8_t)tmp[--nd]) return false; k++; }
    return k + 1 == len && v[k] == '/';
}
EB_DEV void sg_emit_uri(SgEmit& e) { bld_puts(e.b, "http://"); bld_hostport(e.b, e.bp); bld_put(e.b, '/'); }   // "http" ++ get_ssrf_uri()
EB_DEV void sg_emit_params(SgEmit& e, const SgDoc& d, const STok& t, const uint8_t* order) {
    if (t.flags & SF_XMLNS) {                                            // xmlns_modify_params/1 :603-627
        bool changed = false;
        for (uint32_t k = 0; k < t.npar; k++) if (d.par[t.par0 + k].choice) changed = true;
        if (!changed) {
            bld_puts(e.b, " xmlns=\""); sg_emit_uri(e); bld_puts(e.b, "\" xmlns:xsi=\""); sg_emit_uri(e);
            bld_puts(e.b, "\" xsi:schemaLocation=\""); sg_emit_uri(e); bld_put(e.b, '"');
        }
    }
    for (uint32_t k = 0; k < t.npar; k++) {
        const SPar& pr = d.par[t.par0 + (order ? order[k] : k)];
        bld_put(e.b, ' '); bld_copy(e.b, pr.np, pr.nl);
        bool has_val = pr.vl > 0 || pr.choice != 0;
        if (!has_val) continue;                                          // {Name, [], _} folds to " Name" (:297-298)
        bld_put(e.b, '=');
        if (pr.q) bld_put(e.b, pr.q);
        if (pr.choice == 1) { bld_copy(e.b, pr.vp, pr.vl); bld_put(e.b, ' '); sg_emit_uri(e); }
        else if (pr.choice == 2) sg_emit_uri(e);
        else bld_copy(e.b, pr.vp, pr.vl);
        if (pr.q) bld_put(e.b, pr.q);
    }
}
EB_DEV void sg_emit_open(SgEmit& e, const SgDoc& d, uint32_t i, const uint8_t* order) {
    const STok& t = d.tok[i];
    bld_put(e.b, '<'); bld_copy(e.b, t.ptr, t.len); sg_emit_params(e, d, t, order); bld_put(e.b, '>');
}
EB_DEV void sg_emit_tok(SgEmit& e, const SgDoc& d, uint32_t i, uint32_t perm_tok, const uint8_t* perm) {
    const STok& t = d.tok[i];
    switch (t.kind) {
    case ST_OPEN: sg_emit_open(e, d, i, i == perm_tok ? perm : nullptr); break;
    case ST_SC: bld_put(e.b, '<'); bld_copy(e.b, t.ptr, t.len); sg_emit_params(e, d, t, nullptr); bld_puts(e.b, " />"); break;
    case ST_CLOSE: bld_puts(e.b, "</"); bld_copy(e.b, t.ptr, t.len); bld_put(e.b, '>'); break;
    case ST_BANG: bld_puts(e.b, "<!"); bld_copy(e.b, t.ptr, t.len); bld_put(e.b, '>'); break;
    case ST_COMMENT: bld_puts(e.b, "<!--"); bld_copy(e.b, t.ptr, t.len); bld_puts(e.b, "-->"); break;
    case ST_QUE: bld_puts(e.b, "<?"); bld_copy(e.b, t.ptr, t.len); bld_puts(e.b, "?>"); break;
    default: bld_copy(e.b, t.ptr, t.len); break;
    }
}
EB_DEV void sg_emit_range(SgEmit& e, const SgDoc& d, uint32_t lo, uint32_t hi_incl, uint32_t perm_tok, const uint8_t* perm) {
    for (uint32_t i = lo; i <= hi_incl && i < d.ntok; i++) sg_emit_tok(e, d, i, perm_tok, perm);
}

// the mutation as an emission plan; run twice (size, then bytes)
struct SgPlan {
    int which;
    uint32_t a_lo, a_hi, b_lo, b_hi;     // token ranges of the selected elements
    uint32_t times;                      // repeat count / 2^PumpCnt
    uint32_t x_lo, x_hi;                 // pump: the re-inserted element inside [a_lo, a_hi]
    uint32_t perm_tok; const uint8_t* perm;
    bool wrap;                           // sgml_insert with a tag as the new element
};
EB_DEV void sg_emit_plan(SgEmit& e, const SgDoc& d, const SgPlan& pl) {
    const uint32_t last = d.ntok - 1;
    auto all = [&](uint32_t lo, uint32_t hi) { if (lo <= hi && hi != SG_NOMATCH) sg_emit_range(e, d, lo, hi, pl.perm_tok, pl.perm); };
    switch (pl.which) {
    case 0: {                                                            // sgml_swap :540-553
        uint32_t alo = pl.a_lo, ahi = pl.a_hi, blo = pl.b_lo, bhi = pl.b_hi;
        if (alo == blo) { all(0, last); return; }
        bool b_in_a = blo >= alo && bhi <= ahi, a_in_b = alo >= blo && ahi <= bhi;
        if (b_in_a) { if (alo) all(0, alo - 1); all(blo, bhi); all(ahi + 1, last); return; }
        if (a_in_b) { if (blo) all(0, blo - 1); all(alo, ahi); all(bhi + 1, last); return; }
        if (alo < blo) { if (alo) all(0, alo - 1); all(blo, bhi); all(ahi + 1, blo - 1); all(alo, ahi); all(bhi + 1, last); }
        else { if (blo) all(0, blo - 1); all(alo, ahi); all(bhi + 1, alo - 1); all(blo, bhi); all(ahi + 1, last); }
        return;
    }
    case 1: case 3:                                                      // sgml_dup / sgml_repeat :532-538
        all(0, pl.a_hi);
        for (uint32_t k = 0; k < pl.times; k++) all(pl.a_lo, pl.a_hi);
        all(pl.a_hi + 1, last); return;
    case 2: {                                                            // sgml_pump :511-530
        if (pl.a_lo == SG_NOMATCH) { all(0, last); return; }
        if (pl.a_lo) all(0, pl.a_lo - 1);
        for (uint32_t k = 0; k < pl.times; k++) if (pl.x_lo > pl.a_lo) all(pl.a_lo, pl.x_lo - 1);
        all(pl.x_lo, pl.x_hi);
        for (uint32_t k = 0; k < pl.times; k++) all(pl.x_hi + 1, pl.a_hi);
        all(pl.a_hi + 1, last); return;
    }
    case 4: all(0, pl.b_hi); all(pl.a_lo, pl.a_hi); all(pl.b_hi + 1, last); return;   // sgml_insert2 :574-578
    case 7:                                                              // sgml_insert :557-571
        if (!pl.wrap) { all(0, pl.b_hi); all(pl.a_lo, pl.a_hi); all(pl.b_hi + 1, last); return; }
        if (pl.b_lo) all(0, pl.b_lo - 1);
        sg_emit_open(e, d, pl.a_lo, nullptr); all(pl.b_lo, pl.b_hi); sg_emit_tok(e, d, pl.a_hi, SG_NOMATCH, nullptr);
        all(pl.b_hi + 1, last); return;
    case 6: {                                                            // sgml_breaktag :590-601
        if (pl.a_lo == SG_NOMATCH) { all(0, last); return; }
        if (pl.a_lo) all(0, pl.a_lo - 1);
        sg_emit_open(e, d, pl.a_lo, nullptr);
        // the children land behind the broken tag in REVERSE order: walk the child list backwards
        uint32_t hi = pl.a_hi;                                            // the close token
        while (hi > pl.a_lo + 1) {
            uint32_t cend = hi - 1, cbeg = cend;
            if (d.tok[cend].kind == ST_CLOSE && (d.tok[cend].flags & SF_MATCHED_CLOSE)) cbeg = d.tok[cend].match;
            all(cbeg, cend); hi = cbeg;
        }
        all(pl.a_hi + 1, last); return;
    }
    default: all(0, last); return;                                       // 5, 8, 9..11: per-token changes only
    }
}

template <int LVL>
EB_DEV void mut_sgm(CaseCtx& c, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0; r.kind = RES_SAME; r.delta = -1;
    if (mem_binarish(p, n)) return;                                      // parse/2 :185-186
    SgDoc d; d.S = p; d.n = n; d.ntok = 0; d.npar = 0; d.N = d.NT = 0;
    // table sizes for ordinary markup (a tag every four bytes or denser is not)
    d.tok_cap = n / 4 + 64; d.par_cap = n / 4 + 16;                      // beyond that ("<><><>...") the case is flagged
    d.tok = (STok*)temp_alloc(c, (uint64_t)d.tok_cap * sizeof(STok));
    d.par = (SPar*)temp_alloc(c, (uint64_t)d.par_cap * sizeof(SPar));
    if (!d.tok || !d.par) { r.delta = 0; return; }
    FuseClock clk = phase_clock(c); clk.start();
    int k = sg_tokenize(c, d);
    clk.stop(PH_SGM_TOK);
    if (ws->status != CASE_OK) return;
    if (k == 1) return;                                                  // throw(incorrect_sgml) -> {_, Ll, Meta, -1} :754-756
    if (k == 2) { ws->status = CASE_DIED; return; }                      // function_clause: not what sgml_mutate/2 catches
    if (k == 3) { r.kind = RES_UNSUPPORTED; return; }
    clk.start();
    const bool paired = sg_pair(c, d);
    clk.stop(PH_SGM_PAIR);
    if (!paired) { r.kind = RES_UNSUPPORTED; return; }
    const uint32_t N = d.N, NT = d.NT;
    SgPlan pl; pl.which = (int)g.rand(12); pl.a_lo = pl.a_hi = pl.b_lo = pl.b_hi = SG_NOMATCH; pl.times = 0; pl.x_lo = pl.x_hi = 0;
    pl.perm_tok = SG_NOMATCH; pl.perm = nullptr; pl.wrap = false;
    double D = 1;
    switch (pl.which) {
    case 0: case 4: case 7: {
        uint32_t r1 = (uint32_t)g.erand(N), r2 = (uint32_t)g.erand(N);
        pl.a_lo = sg_elem_tok(d, r1); pl.b_lo = sg_elem_tok(d, r2);
        if (pl.a_lo == SG_NOMATCH || pl.b_lo == SG_NOMATCH) { ws->status = CASE_DIED; return; }   // select_elem badmatch (N = 0)
        pl.a_hi = sg_range_hi(d, pl.a_lo); pl.b_hi = sg_range_hi(d, pl.b_lo);
        pl.wrap = pl.which == 7 && pl.a_hi != pl.a_lo;
        break;
    }
    case 1: case 3: {
        uint32_t rr = (uint32_t)g.erand(N);
        pl.times = pl.which == 1 ? 1u : (uint32_t)g.erand(100);
        pl.a_lo = sg_elem_tok(d, rr);
        if (pl.a_lo == SG_NOMATCH) { pl.which = 8; break; }              // nothing carries number 0: the walk changes nothing
        pl.a_hi = sg_range_hi(d, pl.a_lo);
        break;
    }
    case 2: {
        D = -2;
        if (NT == 0) break;
        uint32_t rr = (uint32_t)g.erand(NT);
        pl.a_lo = sg_tag_tok(d, rr); pl.a_hi = d.tok[pl.a_lo].match;
        uint32_t sub = 0; for (uint32_t i = pl.a_lo; i <= pl.a_hi; i++) if (sg_is_elem(d.tok[i])) sub++;
        uint32_t e = (uint32_t)g.erand(sub - 1) + 1;
        uint32_t cnt = (uint32_t)g.erand((uint64_t)trunc(1000.0 / (100.0 + (double)sub)));
        uint32_t cc = 0; pl.x_lo = pl.a_lo;
        for (uint32_t i = pl.a_lo; i <= pl.a_hi; i++) if (sg_is_elem(d.tok[i]) && ++cc == e) { pl.x_lo = i; break; }
        pl.x_hi = sg_range_hi(d, pl.x_lo);
        pl.times = 1u << cnt;
        if (pl.x_lo == pl.a_lo) pl.times = 0;                            // E = 1: the tag replaces itself
        break;
    }
    case 5: {                                                            // sgml_permparams :580-588
        uint32_t rr = (uint32_t)g.erand(NT);
        if (rr == 0) break;
        uint32_t ti = sg_tag_tok(d, rr); uint32_t np = d.tok[ti].npar;
        if (np > 64) { r.kind = RES_UNSUPPORTED; return; }
        uint8_t* ord = temp_alloc(c, 64); double* key = (double*)temp_alloc(c, 64 * sizeof(double));
        if (!ord || !key) { r.delta = 0; return; }
        for (uint32_t q = 0; q < np; q++) ord[q] = (uint8_t)q;
        if (np == 2) { if (g.rand(2) == 1) { ord[0] = 1; ord[1] = 0; } }  // random_permutation/1, src/erlamsa_rnd.erl:190-196
        else {
            for (uint32_t q = 0; q < np; q++) key[q] = g.uniform();
            for (uint32_t q = 1; q < np; q++) { uint8_t o = ord[q]; double kq = key[o]; int j = (int)q - 1; while (j >= 0 && key[ord[j]] > kq) { ord[j + 1] = ord[j]; j--; } ord[j + 1] = o; }
        }
        __syncwarp();
        pl.perm_tok = ti; pl.perm = ord;
        break;
    }
    case 6: {
        uint32_t rr = (uint32_t)g.erand(NT);
        if (rr == 0) break;
        pl.a_lo = sg_tag_tok(d, rr); pl.a_hi = d.tok[pl.a_lo].match;
        (void)g.rand(1);
        break;
    }
    case 8: {                                                            // sgml_xmlfeatures(_, NT, 1) :664-677
        D = -1;
        if (NT == 0) break;
        // walk/3 hands a tag to the fun after its children: visit the tags in the order of their close tokens, with the
        // pre-order tag numbers assigned on the way in
        uint32_t* tnum = (uint32_t*)temp_alloc(c, (uint64_t)d.ntok * 4);
        if (!tnum) { r.delta = 0; return; }
        uint32_t tagno = 0;
        for (uint32_t i = 0; i < d.ntok; i++) if (d.tok[i].kind == ST_OPEN && (d.tok[i].flags & SF_PAIRED)) tnum[i] = ++tagno;
        bool any = false;
        for (uint32_t i = 0; i < d.ntok; i++) {
            if (!(d.tok[i].kind == ST_CLOSE && (d.tok[i].flags & SF_MATCHED_CLOSE))) continue;
            uint32_t o = d.tok[i].match; uint32_t T = tnum[o];
            if (g.erand((uint64_t)trunc((double)T * 1.5)) != 1) continue;     // xmlns_modify/2 :629-636
            d.tok[o].flags |= SF_XMLNS; any = true;
            for (uint32_t q = 0; q < d.tok[o].npar; q++) {
                SPar& pr = d.par[d.tok[o].par0 + q];
                if (pr.nl >= 5 && pr.np[0] == 'x' && pr.np[1] == 'm' && pr.np[2] == 'l' && pr.np[3] == 'n' && pr.np[4] == 's') {
                    uint32_t ch = g.erand(2) == 1 ? 1u : 2u;
                    if (ch == 2 && sg_is_ssrf_uri(c.bp, pr.vp, pr.vl)) ch = 0;   // NewUri =:= Uri: this attribute did not change (:604-608)
                    pr.choice = ch;
                }
            }
        }
        __syncwarp();
        D = any ? 1 : -1;
        break;
    }
    default: {                                                           // inner text :721-733
        if (LVL >= MAX_NEST) { r.kind = RES_UNSUPPORTED; return; }
        const int kind0 = c.snand_kind;
        const int inner_kind = (int)g.rand_elem_idx(3); (void)g.rand_elem_idx(1);
        const uint8_t ids[11] = {M_AB, M_AD, M_NUM, M_BD, M_SD, M_LD, M_LRI, M_LR, M_LP, M_B64, M_URI};
        MutRow* rows = (MutRow*)temp_alloc(c, sizeof(MutRow) * 11);
        if (!rows) { r.delta = 0; return; }
        int nr = inner_table(c, ids, 11, false, rows);
        const uint64_t floor0 = c.temp_floor;
        auto innertext = [&](const uint8_t*& tp, uint32_t& tl, uint32_t nt) {   // mutate_innertext/3 :683-690
            uint32_t nw = 0;
            for (uint32_t i = lane_id(); i < tl; i += 32) { uint32_t ch = tp[i]; nw |= (ch != 0 && ch != 10 && ch != 13 && ch != 32) ? 1u : 0u; }
            if (!__any_sync(0xffffffffu, nw != 0) || nt == 0) return;
            double rnd = g.uniform();
            if (rnd > 3.0 / (double)nt) return;
            c.temp_floor = c.temp_used; c.snand_kind = inner_kind;
            InnerRes res; res.kind = 0; res.len = tl;
            if constexpr (LVL < MAX_NEST) res = inner_round<LVL>(c, rows, nr, tp, tl, true);
            c.snand_kind = kind0;
            if (ws->status == CASE_OK && res.kind != 0) {
                uint8_t* lit = scratch_alloc(c, res.len);
                if (lit) { inner_result_write(c, res, tp, lit); tp = lit; tl = res.len; }
            }
            c.temp_floor = floor0;
        };
        // walk2acc :361-380: text nodes as they come, a tag's attribute values once its children are done
        const uint64_t keep = c.temp_used;
        for (uint32_t i = 0; i < d.ntok && ws->status == CASE_OK; i++) {
            STok& t = d.tok[i];
            c.temp_used = keep;
            if ((t.kind == ST_TEXT || t.kind == ST_EOFTEXT) && t.len > 0) { const uint8_t* tp = t.ptr; uint32_t tl = t.len; innertext(tp, tl, NT); t.ptr = tp; t.len = tl; }
            else if (t.kind == ST_CLOSE && (t.flags & SF_MATCHED_CLOSE)) {
                STok& o = d.tok[t.match];
                for (uint32_t q = 0; q < o.npar && ws->status == CASE_OK; q++) {
                    SPar& pr = d.par[o.par0 + q];
                    c.temp_used = keep;
                    const uint8_t* tp = pr.vp; uint32_t tl = pr.vl; innertext(tp, tl, NT + o.npar); pr.vp = tp; pr.vl = tl;
                }
            }
        }
        c.temp_used = keep;
        if (ws->status != CASE_OK) return;
        break;
    }
    }
    // fold_ast/2: size, then bytes
    SgEmit em; em.bp = c.bp; em.b.p = nullptr; em.b.n = 0; em.b.cap = 0; em.b.ovf = 0;
    sg_emit_plan(em, d, pl);                                             // cap 0: nothing is stored, the lengths add up
    uint32_t total = em.b.n;
    if (em.b.ovf || total > c.bp->max_case_out) { ws->status = CASE_OVERFLOW; ws->reason = 5; return; }
    uint8_t* out = scratch_alloc(c, (uint64_t)total + 1);
    if (!out) { r.delta = 0; return; }
    em.b.p = out; em.b.n = 0; em.b.cap = total;
    sg_emit_plan(em, d, pl);
    __syncwarp();
    bool same = total == n && range_equal(out, p, n);
    if (same) { r.kind = RES_SAME; r.delta = -1; return; }               // NewBinStr =:= H :745-747
    t_reset(ws); t_push(ws, seg_copy(out, total));
    r.kind = RES_SEGS; r.delta = D + trunc((double)total / (double)(AVG_BLOCK_SIZE * 10));
}

}  // namespace eb
