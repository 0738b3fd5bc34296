// erlamsa_b200 -- erlamsa_field_predict on the device (reference src/erlamsa_field_predict.erl:51-166):
// the brute-force length-field search (u8/u16/u32/u64, both endians; up to 513 x 514 x 5 candidate
// (offset, end) pairs per call) and the xor8 / crc32 trailer search (up to 961 preamble lengths),
// plus the `len` mutator (reference src/erlamsa_mutations.erl:1107-1143).
//
// The reference returns LISTS and then picks an element by index (rand_elem), so the device first
// counts the matches row by row in the reference's list order (lanes work on candidates of a row in
// parallel), draws the index, and then re-walks only the row that holds the chosen match.
#pragma once
#include "eb_state.cuh"

namespace eb {

struct Sizer { uint32_t size_bits; uint32_t big; uint64_t len; uint32_t a; uint32_t b; };

__device__ __forceinline__ uint64_t rd_field(const uint8_t* p, uint32_t a, int bytes, bool big) {
    uint64_t v = 0;
    for (int i = 0; i < bytes; i++) { uint32_t c = p[a + (big ? i : bytes - 1 - i)]; v = (v << 8) | c; }
    return v;
}
// basic_len/2 :66-78: first matching field encoding for the range {A, B}; returns 0 or 1
__device__ __forceinline__ int basic_len_dev(const uint8_t* p, uint32_t size, int64_t a, int64_t b, Sizer* out) {
    if (!(a < b && b > 0 && a < (int64_t)size)) return 0;
#pragma unroll
    for (int e = 0; e < 2; e++) {
#pragma unroll
        for (int wi = 0; wi < 3; wi++) {
            int w = 2 << wi;
            if (a + w > (int64_t)size) continue;
            int64_t want = b - a - w;
            if (want <= 2) continue;
            uint64_t len = rd_field(p, (uint32_t)a, w, e == 0);
            if (len == (uint64_t)want) { if (out) { out->size_bits = (uint32_t)w * 8; out->big = e == 0; out->len = len; out->a = (uint32_t)a; out->b = (uint32_t)b; } return 1; }
        }
    }
    return 0;
}
__device__ __forceinline__ int basic_u8len_dev(const uint8_t* p, uint32_t size, int64_t a, int64_t b, Sizer* out) {   // :51-58
    if (!(a < b && b > 0 && a < (int64_t)size)) return 0;
    int64_t want = b - a - 1;
    if (want <= 2 || (uint64_t)p[a] != (uint64_t)want) return 0;
    if (out) { out->size_bits = 8; out->big = 1; out->len = (uint64_t)want; out->a = (uint32_t)a; out->b = (uint32_t)b; }
    return 1;
}
// the five end offsets of simple_len/2 :80-88
__device__ __forceinline__ int64_t simple_len_end(int64_t b, int v) { return v == 0 ? b : v == 1 ? b - 1 : v == 2 ? b - 2 : v == 3 ? b - 4 : b - 8; }

// get_possible_simple_lens/1 :90-105 + rand_elem: returns false when the list is empty (no draw then).
// List order: SmallLens (A ascending, X = 0..8), then simple_len for (X descending, Y descending), then
// simple_len({A, Len}) for A descending.
EB_DEV bool lens_pick(CaseCtx& c, const uint8_t* p, uint32_t size, Sizer& out) {
    Rng& g = c.rng;
    int l = lane_id();
    if (size <= 10) {   // :101-105
        Sizer found[40]; int nf = 0;
        for (int64_t x = 0; x <= 3; x++) {
            for (int v = 0; v < 5; v++) { Sizer s; if (basic_len_dev(p, size, x, simple_len_end((int64_t)size, v), &s)) found[nf++] = s; }
            for (int k = 0; k <= 8; k++) { Sizer s; if (basic_u8len_dev(p, size, x, (int64_t)size - k, &s)) found[nf++] = s; }
        }
        int64_t ei = g.rand_elem_idx((uint64_t)nf);
        if (ei < 0) return false;
        out = found[ei]; return true;
    }
    uint32_t sub = size / 5; if (sub > 512) sub = 512;
    uint32_t nrow = sub + 1;
    uint32_t* varb = (uint32_t*)temp_alloc(c, (uint64_t)nrow * 4);
    uint32_t* rowcnt = (uint32_t*)temp_alloc(c, (uint64_t)nrow * 4);
    if (!varb || !rowcnt) return false;
    for (uint32_t i = 0; i < nrow; i++) { uint32_t v = (uint32_t)g.rand_range((int64_t)sub, (int64_t)size); varb[i] = v; }   // all lanes store
    // ---- count
    uint32_t small = 0;
    for (uint32_t a = l; a < nrow; a += 32) for (int k = 0; k <= 8; k++) small += basic_u8len_dev(p, size, a, (int64_t)size - k, nullptr);
    small = warp_sum(small);
    uint32_t big_total = 0;
    for (uint32_t xi = 0; xi < nrow; xi++) {           // rows in list order: X = sub - xi
        int64_t x = (int64_t)sub - xi; uint32_t cnt = 0;
        // the six field readings depend on X only: read them once per row, compare per candidate end
        uint64_t fv[6]; bool fok[6];
#pragma unroll
        for (int e = 0; e < 2; e++)
#pragma unroll
            for (int wi = 0; wi < 3; wi++) { int w = 2 << wi; fok[e * 3 + wi] = x + w <= (int64_t)size; fv[e * 3 + wi] = fok[e * 3 + wi] ? rd_field(p, (uint32_t)x, w, e == 0) : 0; }
        for (uint32_t yi = l; yi < nrow; yi += 32) {
            int64_t y = varb[yi];
            for (int v = 0; v < 5; v++) {
                int64_t b = simple_len_end(y, v);
                if (!(x < b && b > 0)) continue;
                int hit = 0;
#pragma unroll
                for (int k = 0; k < 6; k++) { int w = 2 << (k % 3); int64_t want = b - x - w; if (!hit && fok[k] && want > 2 && fv[k] == (uint64_t)want) hit = 1; }
                cnt += hit;
            }
        }
        cnt = warp_sum(cnt);
        rowcnt[xi] = cnt; big_total += cnt;
    }
    uint32_t tail = 0;
    for (uint32_t a = l; a < nrow; a += 32) for (int v = 0; v < 5; v++) tail += basic_len_dev(p, size, a, simple_len_end((int64_t)size, v), nullptr);
    tail = warp_sum(tail);
    uint64_t total = (uint64_t)small + big_total + tail;
    int64_t ei = g.rand_elem_idx(total);
    if (ei < 0) return false;
    uint64_t r = (uint64_t)ei;
    // ---- locate (serial re-walk of one region; warp-uniform)
    if (r < small) {
        for (uint32_t a = 0; a < nrow; a++) for (int k = 0; k <= 8; k++) { Sizer s; if (basic_u8len_dev(p, size, a, (int64_t)size - k, &s)) { if (r == 0) { out = s; return true; } r--; } }
        return false;
    }
    r -= small;
    if (r < big_total) {
        uint32_t xi = 0; while (r >= rowcnt[xi]) { r -= rowcnt[xi]; xi++; }
        int64_t x = (int64_t)sub - xi;
        for (uint32_t yk = nrow; yk-- > 0;) { int64_t y = varb[yk]; for (int v = 0; v < 5; v++) { Sizer s; if (basic_len_dev(p, size, x, simple_len_end(y, v), &s)) { if (r == 0) { out = s; return true; } r--; } } }
        return false;
    }
    r -= big_total;
    for (int64_t a = sub; a >= 0; a--) for (int v = 0; v < 5; v++) { Sizer s; if (basic_len_dev(p, size, a, simple_len_end((int64_t)size, v), &s)) { if (r == 0) { out = s; return true; } r--; } }
    return false;
}

// field value written the way <<Len:Size/endian>> does (truncating to Size bits)
__device__ __forceinline__ uint64_t enc_field(uint64_t v, uint32_t bits, bool big) {
    uint32_t bytes = bits / 8; uint64_t o = 0;
    for (uint32_t i = 0; i < bytes; i++) { uint32_t sh = big ? (bytes - 1 - i) * 8 : i * 8; uint64_t byte = sh >= 64 ? 0 : (v >> sh) & 255; o |= byte << (8 * i); }
    return o;   // little-endian packing of the byte sequence, ready for a SEG_INLINE
}

// random_block/1 with the draws spread over the lanes: each AS183 component is x_k = x_0 * a^k, so lane l
// takes draws l, l+32, ... by stepping with a^32; byte N-1-i holds draw i (reference src/erlamsa_rnd.erl:165,173-174)
EB_DEV void random_block_fill(CaseCtx& c, uint8_t* buf, uint32_t nbytes) {
    Rng& g = c.rng;
    if (g.mode != 0 || nbytes < 64) {
        for (uint32_t i = 0; i < nbytes; i++) { uint32_t b = (uint32_t)g.rand(256); if (lane_id() == 0) buf[nbytes - 1 - i] = (uint8_t)b; }
        __syncwarp(); return;
    }
    Rng mine = g; mine.jump((uint64_t)lane_id());
    uint32_t s1 = modpow_u32<30269>(AS_M1, 31), s2 = modpow_u32<30307>(AS_M2, 31), s3 = modpow_u32<30323>(AS_M3, 31);
    for (uint32_t i = lane_id(); i < nbytes; i += 32) {
        uint32_t b = (uint32_t)mine.rand(256);      // advances one draw
        buf[nbytes - 1 - i] = (uint8_t)b;
        mine.a1 = (int32_t)(((uint32_t)mine.a1 * s1) % 30269u); mine.a2 = (int32_t)(((uint32_t)mine.a2 * s2) % 30307u); mine.a3 = (int32_t)(((uint32_t)mine.a3 * s3) % 30323u);
    }
    g.jump((uint64_t)nbytes);
    __syncwarp();
}

// length_predict/2 :1137-1143 with mutate_length/2 :1113-1135
EB_DEV void mut_len(CaseCtx& c, const uint8_t* p, uint32_t n, MutResult& r) {
    WarpState* ws = c.ws; Rng& g = c.rng;
    r.rechunk = 0; r.consumed_next = 0;
    Sizer e;
    if (!lens_pick(c, p, n, e)) { r.kind = RES_SAME; r.delta = -2; return; }
    if (ws->status != CASE_OK) { r.kind = RES_SAME; r.delta = 0; return; }
    uint32_t fb = e.size_bits / 8;
    uint32_t blob0 = e.a + fb, blob1 = blob0 + (uint32_t)e.len;
    // <<TmpNewLen:Size>> = random_block(Size/8): draw i lands at byte fb-1-i, read big-endian
    uint64_t tv = 0; uint8_t tb[8];
    for (uint32_t i = 0; i < fb; i++) tb[fb - 1 - i] = (uint8_t)g.rand(256);
    bool over = false;
    for (uint32_t i = 0; i < fb; i++) { if (tv >> 56) over = true; tv = (tv << 8) | tb[i]; }
    uint64_t newlen = (over || tv > ABSMAX_BINARY_BLOCK) ? ABSMAX_BINARY_BLOCK : (tv * 2 < ABSMAX_BINARY_BLOCK ? tv * 2 : ABSMAX_BINARY_BLOCK);
    uint64_t cs = g.rand(7);
    t_reset(ws);
    r.delta = 1; r.kind = RES_SEGS;
    if (cs == 0 || cs == 1) {
        t_push(ws, seg_copy(p, e.a)); t_push(ws, seg_fill(cs == 0 ? 0 : 0xff, fb)); t_push(ws, seg_copy(p + blob0, n - blob0)); return;
    }
    if (cs == 2) {   // the blob grows by a random block; the length field keeps its value
        uint32_t nl = (uint32_t)newlen;
        uint8_t* buf = scratch_alloc(c, nl);
        if (!buf) { r.kind = RES_SAME; r.delta = 0; return; }
        if (nl < ABSMAXHALF_BINARY_BLOCK) random_block_fill(c, buf, nl);
        else {   // fast_pseudorandom_block/1, reference src/erlamsa_rnd.erl:155-160: <<42:Z8L, Rnd/binary>> with Z8L = N - 500000 BITS
            uint32_t zbits = nl - ABSMAXHALF_BINARY_BLOCK;
            if (zbits % 8) { ws->status = CASE_DIED; r.kind = RES_SAME; return; }      // not a binary: the reference's worker crashes
            uint32_t zb = zbits / 8;
            random_block_fill(c, buf + zb, ABSMAXHALF_BINARY_BLOCK);
            for (uint32_t i = lane_id(); i < zb; i += 32) buf[i] = (i + 1 == zb) ? 42 : 0;
            __syncwarp();
            nl = zb + ABSMAXHALF_BINARY_BLOCK;
        }
        t_push(ws, seg_copy(p, blob1)); t_push(ws, seg_copy(buf, nl)); t_push(ws, seg_copy(p + blob1, n - blob1)); return;
    }
    uint64_t enc = enc_field(newlen, e.size_bits, e.big != 0);
    t_push(ws, seg_copy(p, e.a)); t_push(ws, seg_inline(enc, fb));
    if (cs == 3) t_push(ws, seg_copy(p + blob1, n - blob1));        // blob dropped
    else t_push(ws, seg_copy(p + blob0, n - blob0));
}

// ------------------------------------------------------------------ crc32 (zlib polynomial, reflected)
__device__ __forceinline__ uint32_t crc32_byte(uint32_t crc, uint32_t b) {
    crc ^= b;
#pragma unroll
    for (int k = 0; k < 8; k++) crc = (crc >> 1) ^ (0xEDB88320u & (0u - (crc & 1u)));
    return crc;
}
__device__ __forceinline__ uint32_t crc_multmodp(uint32_t a, uint32_t b) {   // a(x) * b(x) mod p(x), reflected
    uint32_t m = 1u << 31, pr = 0;
    for (;;) {
        if (a & m) { pr ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return pr;
}
__device__ __forceinline__ uint32_t crc_x2nmodp(uint64_t n, uint32_t k) {   // x^(n * 2^k) mod p(x)
    uint32_t p = 1u << 31;
    // x^(2^k) by repeated squaring from x^1
    uint32_t sq = 1u << 30;                   // x^1
    for (uint32_t i = 0; i < k; i++) sq = crc_multmodp(sq, sq);
    while (n) { if (n & 1) p = crc_multmodp(sq, p); sq = crc_multmodp(sq, sq); n >>= 1; }
    return p;
}
// CRC(A || B) from CRC(A), CRC(B), |B|   (zlib crc32_combine)
__device__ __forceinline__ uint32_t crc_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) { return crc_multmodp(crc_x2nmodp(len2, 3), crc1) ^ crc2; }

// erlang:crc32 of [p, p+n): each lane takes one contiguous slice, then the slices are combined in order
EB_DEV uint32_t warp_crc32(const uint8_t* p, uint64_t n) {
    uint64_t per = (n + 31) / 32; uint64_t s = per * lane_id(); uint64_t e = s + per < n ? s + per : n; if (s > n) s = n;
    uint32_t crc = 0xffffffffu;
    for (uint64_t i = s; i < e; i++) crc = crc32_byte(crc, p[i]);
    crc = ~crc; uint64_t mylen = e - s;
    uint32_t acc = 0;   // crc of the empty string
    for (int l = 0; l < 32; l++) {
        uint32_t cl = __shfl_sync(0xffffffffu, crc, l); uint64_t ll = __shfl_sync(0xffffffffu, mylen, l);
        if (ll) acc = crc_combine(acc, cl, ll);
    }
    return acc;
}

// get_possible_csum_locations/1 :154-161 + rand_elem. type: 0 xor8, 1 crc32
struct Csum { uint32_t crc; uint32_t plen; uint32_t blen; };
EB_DEV bool csum_pick(CaseCtx& c, const uint8_t* p, uint32_t len, Csum& out) {
    Rng& g = c.rng; int l = lane_id();
    if (len == 0) return false;
    uint32_t maxa = (uint32_t)trunc(2.0 * (double)len / 3.0); if (maxa > 960) maxa = 960;
    uint32_t na = maxa + 1;
    uint8_t* hit = (uint8_t*)temp_alloc(c, (uint64_t)2 * na);
    if (!hit) return false;
    // xor8 of [A, len-1) = total ^ prefix(A)
    uint32_t x = 0;
    for (uint32_t i = l; i + 1 < len; i += 32) x ^= p[i];
    for (int o = 16; o; o >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, o);
    uint32_t last = p[len - 1];
    uint32_t pre = 0, nx = 0;
    for (uint32_t a = 0; a < na; a++) { bool h = ((x ^ pre) & 0xff) == last; hit[a] = h ? 1 : 0; nx += h ? 1u : 0u; if (a + 1 < len) pre ^= p[a]; }
    // crc32 of [A, len-4) against the big-endian trailer
    uint32_t nc = 0;
    if (len >= 4) {
        uint32_t body = len - 4;
        uint32_t trailer = ((uint32_t)p[len - 4] << 24) | ((uint32_t)p[len - 3] << 16) | ((uint32_t)p[len - 2] << 8) | p[len - 1];
        uint32_t whole = warp_crc32(p, body);
        // CRC(prefix_A) for A <= maxa, serial (<= 961 bytes); CRC(suffix_A) = whole ^ shift(CRC(prefix_A), body - A)
        uint32_t* pc = (uint32_t*)temp_alloc(c, (uint64_t)na * 4);
        if (!pc) return false;
        uint32_t run = 0xffffffffu;
        for (uint32_t a = 0; a < na; a++) { pc[a] = ~run; if (a < body) run = crc32_byte(run, p[a]); }
        for (uint32_t a = l; a < na; a += 32) {
            uint8_t h = 0;
            if (len - a >= 4 && a <= body) { uint32_t suf = whole ^ crc_multmodp(crc_x2nmodp(body - a, 3), pc[a]); h = suf == trailer ? 1 : 0; }
            hit[na + a] = h;
        }
        __syncwarp();
        for (uint32_t a = 0; a < na; a++) nc += hit[na + a];
    } else for (uint32_t a = 0; a < na; a++) hit[na + a] = 0;
    int64_t ei = g.rand_elem_idx((uint64_t)nx + nc);
    if (ei < 0) return false;
    uint32_t r = (uint32_t)ei;
    for (uint32_t a = 0; a < na; a++) if (hit[a]) { if (r == 0) { out.crc = 0; out.plen = a; out.blen = len - a - 1; return true; } r--; }
    for (uint32_t a = 0; a < na; a++) if (hit[na + a]) { if (r == 0) { out.crc = 1; out.plen = a; out.blen = len - a - 4; return true; } r--; }
    return false;
}

}  // namespace eb
