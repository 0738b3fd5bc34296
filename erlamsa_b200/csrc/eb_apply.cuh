// erlamsa_b200 -- output placement (prefix sum) and the apply kernel.
//
// The apply kernel is the only place payload bytes move in the common case: it executes every
// case's edit script, reading each input byte once and writing each output byte once
// (algorithmic traffic = len_in + len_out + descriptors; HBM roofline).
// Work is laid out on PACKED OUTPUT coordinates: CTA b produces output bytes
// [b*TILE, (b+1)*TILE), whatever cases they belong to, so load balance does not depend on the
// size mix of the corpus. Every case starts on a 16-byte boundary, every thread emits aligned
// 16-byte stores; sources are read with aligned 16-byte loads and realigned in registers.
//
// Per tile: warp 0 turns the (case, segment) records that intersect the tile into a short list of
// SPANS in shared memory (dst range inside the tile + source), one lane per case, then the whole
// CTA streams the spans' full 16-byte words with several independent loads in flight per thread;
// the few words that straddle span boundaries are assembled byte-wise afterwards.
#pragma once
#include "eb_state.cuh"

namespace eb {

// ------------------------------------------------------------------ exclusive prefix sum of 16-aligned sizes
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                       // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ inline uint64_t block_excl_scan_u64(uint64_t v, uint64_t* sh, uint64_t& total) {
    int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, x, o); if (l >= o) x += y; }
    if (l == 31) sh[w] = x;
    __syncthreads();
    uint64_t woff = 0, tot = 0;
    for (int i = 0; i < SCAN_THREADS / 32; i++) { uint64_t s = sh[i]; if (i < w) woff += s; tot += s; }
    __syncthreads();
    total = tot;
    return woff + x - v;
}
// pass 1: per-tile totals
__global__ void __launch_bounds__(SCAN_THREADS) eb_scan_tiles(const uint64_t* __restrict__ sz, uint64_t n, uint64_t* __restrict__ tile_sum) {
    __shared__ uint64_t sh[SCAN_THREADS / 32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t s = 0;
    for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) s += sz[base + i];
    uint64_t tot; block_excl_scan_u64(s, sh, tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}
// pass 2: one CTA turns tile totals into tile offsets (serial over chunks of SCAN_THREADS tiles)
__global__ void __launch_bounds__(SCAN_THREADS) eb_scan_tile_offsets(uint64_t* __restrict__ tile_sum, uint64_t ntiles, uint64_t* __restrict__ grand_total) {
    __shared__ uint64_t sh[SCAN_THREADS / 32];
    uint64_t carry = 0;
    for (uint64_t b = 0; b < ntiles; b += SCAN_THREADS) {
        uint64_t i = b + threadIdx.x;
        uint64_t v = i < ntiles ? tile_sum[i] : 0, tot;
        uint64_t ex = block_excl_scan_u64(v, sh, tot);
        if (i < ntiles) tile_sum[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}
// pass 3: per-element exclusive offsets; out_off has n+1 entries
__global__ void __launch_bounds__(SCAN_THREADS) eb_scan_finish(const uint64_t* __restrict__ sz, uint64_t n, const uint64_t* __restrict__ tile_off,
                                                               const uint64_t* __restrict__ grand_total, uint64_t* __restrict__ out_off) {
    __shared__ uint64_t sh[SCAN_THREADS / 32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS], s = 0;
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = base + i < n ? sz[base + i] : 0; s += v[i]; }
    uint64_t tot; uint64_t ex = block_excl_scan_u64(s, sh, tot) + tile_off[blockIdx.x];
    for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out_off[base + i] = ex; ex += v[i]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out_off[n] = *grand_total;
}

// ------------------------------------------------------------------ apply
constexpr int APPLY_THREADS = 256;
constexpr int MAX_SPANS = 96;
constexpr int SPAN_LANE_SEGS = 24;     // a lane walks at most this many segments of its case before giving up

// tile -> first case table: tile_case[t] = the case that owns packed output byte t*tile.
// One thread per case marks the tiles that START inside its 16-byte-aligned extent.
__global__ void __launch_bounds__(256) eb_tile_cases(const uint64_t* __restrict__ out_off, uint64_t n_cases, uint32_t* __restrict__ tile_case, uint32_t tile) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_cases) return;
    uint64_t beg = out_off[k], end = out_off[k + 1];
    for (uint64_t t = (beg + tile - 1) / tile; t * tile < end; t++) tile_case[t] = (uint32_t)k;
}

// byte `i` of one segment (slow path)
__device__ __forceinline__ uint32_t seg_byte(const Seg& s, uint32_t i) {
    switch (s.kind()) {
    case SEG_COPY: return ((const uint8_t*)(uintptr_t)s.src)[i];
    case SEG_INLINE: return (uint32_t)((s.src >> (8 * i)) & 0xff);
    case SEG_REPEAT: return ((const uint8_t*)(uintptr_t)s.src)[i % s.arg()];
    default: return s.arg() & 0xff;
    }
}

// a segment clipped to the tile: output bytes [d0, d1) (tile relative) come from `seg` starting at its byte `so`
struct __align__(16) Span { uint32_t d0, d1; uint32_t so; uint32_t pad; Seg seg; };

__device__ __forceinline__ uint32_t span_byte(const Span& sp, uint32_t d) { return seg_byte(sp.seg, sp.so + (d - sp.d0)); }

// generic per-word walk over (case, segment) records straight from global memory: used for tiles whose
// span list does not fit (very many tiny cases or very long scripts inside one tile)
template <int WORDS>
__device__ __noinline__ void apply_tile_generic(const CaseOut* __restrict__ cases, const Seg* __restrict__ segs, const uint64_t* __restrict__ out_off,
                                                uint64_t ci, uint64_t tile0, uint64_t total, uint8_t* __restrict__ out, uint64_t out_capacity) {
    uint64_t c_beg = out_off[ci], c_end = out_off[ci + 1];
    CaseOut co = cases[ci];
    uint32_t sk = 0; uint64_t seg_dst = 0; Seg sg; sg.src = 0; sg.len = 0; sg.meta = 0; bool have = false;
#pragma unroll 1
    for (int w = 0; w < WORDS; w++) {
        uint64_t g = tile0 + ((uint64_t)w * APPLY_THREADS + threadIdx.x) * 16;
        if (g >= total || g + 16 > out_capacity) break;
        while (g >= c_end) { ci++; c_beg = c_end; c_end = out_off[ci + 1]; co = cases[ci]; have = false; }
        uint64_t l = g - c_beg;
        if (l >= co.out_len) continue;
        uint32_t want = (uint32_t)((co.out_len - l) < 16 ? (co.out_len - l) : 16);
        if (!have) { sg = segs[co.seg_begin]; have = true; sk = 0; seg_dst = 0; }
        while (l >= seg_dst + sg.len) { seg_dst += sg.len; sk++; sg = segs[co.seg_begin + sk]; }
        uint32_t bytes[4] = {0, 0, 0, 0};
        uint32_t k2 = sk; uint64_t d2 = seg_dst; Seg s2 = sg;
        for (uint32_t j = 0; j < want; j++) {
            uint64_t lj = l + j;
            while (lj >= d2 + s2.len) { d2 += s2.len; k2++; s2 = segs[co.seg_begin + k2]; }
            bytes[j >> 2] |= seg_byte(s2, (uint32_t)(lj - d2)) << (8 * (j & 3));
        }
        stg16_stream(out + g, make_uint4(bytes[0], bytes[1], bytes[2], bytes[3]));
    }
}

// full 16-byte word at tile offset d (d0 <= d, d + 16 <= d1) of a span
__device__ __forceinline__ uint4 span_word(const Span& sp, uint32_t d) {
    uint32_t o = sp.so + (d - sp.d0);
    uint32_t kind = sp.seg.kind();
    if (kind == SEG_COPY) return load16_unaligned<true>((const uint8_t*)(uintptr_t)sp.seg.src + o, 16);
    if (kind == SEG_FILL) { uint32_t b = sp.seg.arg() & 0xff; b |= b << 8; b |= b << 16; return make_uint4(b, b, b, b); }
    // SEG_REPEAT (inline literals are at most 8 bytes long and never hold a full word)
    uint32_t unit = sp.seg.arg(), r = o % unit;
    const uint8_t* src = (const uint8_t*)(uintptr_t)sp.seg.src;
    if (r + 16 <= unit) return load16_unaligned<true>(src + r, 16);
    uint32_t bytes[4] = {0, 0, 0, 0};
    for (uint32_t j = 0; j < 16; j++) { bytes[j >> 2] |= (uint32_t)src[r] << (8 * (j & 3)); r++; if (r == unit) r = 0; }
    return make_uint4(bytes[0], bytes[1], bytes[2], bytes[3]);
}

template <int WORDS, int BATCH, int MIN_CTAS>
__global__ void __launch_bounds__(APPLY_THREADS, MIN_CTAS)
eb_apply_kernel(const CaseOut* __restrict__ cases, const Seg* __restrict__ segs, const uint64_t* __restrict__ out_off,
                const uint32_t* __restrict__ tile_case, uint64_t n_cases, uint8_t* __restrict__ out, uint64_t out_capacity) {
    constexpr uint32_t TILE = APPLY_THREADS * WORDS * 16;
    __shared__ Span spans[MAX_SPANS];
    __shared__ uint32_t wpre[MAX_SPANS + 1];      // prefix sum of full words per span
    __shared__ int s_nspan, s_fallback;
    uint64_t total = out_off[n_cases];
    uint64_t tile0 = (uint64_t)blockIdx.x * TILE;
    if (tile0 >= total) return;
    uint64_t tile1 = tile0 + TILE < total ? tile0 + TILE : total;
    if (tile1 > out_capacity) tile1 = out_capacity & ~15ull;
    uint64_t c0 = tile_case[blockIdx.x];
    if (threadIdx.x == 0) { s_nspan = 0; s_fallback = 0; }
    __syncthreads();
    // ---- span construction: warp 0, one lane per case
    if (threadIdx.x < 32) {
        for (uint64_t base = c0;; base += 32) {
            uint64_t ci = base + threadIdx.x;
            bool valid = ci < n_cases;
            uint64_t beg = valid ? out_off[ci] : ~0ull;
            bool starts_before_end = valid && beg < tile1;
            if (starts_before_end) {
                CaseOut co = cases[ci];
                if (beg + co.out_len > tile0 && co.out_len > 0) {
                    uint64_t pos = beg;                       // packed output position of the current segment
                    for (uint32_t s = 0; s < co.nseg && pos < tile1; s++) {
                        if (s >= SPAN_LANE_SEGS) { s_fallback = 1; break; }
                        Seg sg = segs[co.seg_begin + s];
                        uint64_t e = pos + sg.len;
                        if (e > tile0) {
                            uint64_t a = pos > tile0 ? pos : tile0, z = e < tile1 ? e : tile1;
                            int slot = atomicAdd(&s_nspan, 1);
                            if (slot < MAX_SPANS) {
                                Span sp; sp.d0 = (uint32_t)(a - tile0); sp.d1 = (uint32_t)(z - tile0); sp.so = (uint32_t)(a - pos); sp.pad = 0; sp.seg = sg;
                                spans[slot] = sp;
                            } else s_fallback = 1;
                        }
                        pos = e;
                    }
                }
            }
            // continue while the last lane's case still starts inside the tile
            bool more = __shfl_sync(0xffffffffu, starts_before_end ? 1 : 0, 31) != 0;
            if (!more) break;
        }
    }
    __syncthreads();
    if (s_fallback) { apply_tile_generic<WORDS>(cases, segs, out_off, c0, tile0, total, out, out_capacity); return; }
    int ns = s_nspan;
    // ---- prefix of full-word counts (spans are few: one thread)
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int i = 0; i < ns; i++) {
            wpre[i] = acc;
            uint32_t f0 = (spans[i].d0 + 15u) & ~15u, f1 = spans[i].d1 & ~15u;
            acc += f1 > f0 ? (f1 - f0) >> 4 : 0;
        }
        wpre[ns] = acc;
    }
    __syncthreads();
    uint32_t nwords = wpre[ns];
    uint8_t* tout = out + tile0;
    // ---- bulk: every full word of every span; BATCH independent loads in flight per thread
    int cur = 0;
    for (uint32_t it0 = threadIdx.x; it0 < nwords; it0 += APPLY_THREADS * BATCH) {
        uint4 v[BATCH]; uint32_t dpos[BATCH]; bool ok[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; j++) {
            uint32_t it = it0 + j * APPLY_THREADS;
            ok[j] = it < nwords;
            if (ok[j]) {
                while (it >= wpre[cur + 1]) cur++;
                const Span& sp = spans[cur];
                uint32_t d = ((sp.d0 + 15u) & ~15u) + ((it - wpre[cur]) << 4);
                dpos[j] = d;
                v[j] = span_word(sp, d);
            }
        }
#pragma unroll
        for (int j = 0; j < BATCH; j++) if (ok[j]) stg16_stream(tout + dpos[j], v[j]);
    }
    // ---- seam words: the word holding a span's last byte when that byte does not end the word
    if ((int)threadIdx.x < ns) {
        const Span& me = spans[threadIdx.x];
        if ((me.d1 & 15u) != 0) {
            uint32_t w = (me.d1 - 1) & ~15u;
            if (w >= me.d0) {   // this span owns the word (it holds the word's first byte)
                uint32_t bytes[4] = {0, 0, 0, 0};
                for (uint32_t j = 0; j < 16; j++) {
                    uint32_t d = w + j, b = 0;
                    if (d < me.d1) b = span_byte(me, d);
                    else for (int q = 0; q < ns; q++) if (d >= spans[q].d0 && d < spans[q].d1) { b = span_byte(spans[q], d); break; }
                    bytes[j >> 2] |= b << (8 * (j & 3));
                }
                if (tile0 + w + 16 <= out_capacity) stg16_stream(tout + w, make_uint4(bytes[0], bytes[1], bytes[2], bytes[3]));
            }
        }
    }
}

}  // namespace eb
