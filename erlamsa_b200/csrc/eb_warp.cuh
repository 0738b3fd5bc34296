// erlamsa_b200 -- warp-level building blocks: 16-byte realigning loads, cooperative copies and
// the ballot/popcount byte scans (newline / digit-run boundaries) the line and number mutators use.
//
// Execution model of the decide kernel: ONE WARP PER TEST CASE. Scalar decision logic (RNG
// draws, scheduler) is executed redundantly and identically by all 32 lanes ("warp-uniform"),
// so there is no divergence and no broadcast traffic; data-parallel work (scans, compares,
// copies) is split across lanes with 16-byte vector accesses.
#pragma once
#include <stdint.h>
#include "eb_common.cuh"

namespace eb {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t& total) {
    uint32_t x = v; int l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (l >= o) x += y; }
    total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// streaming 16-byte accesses: read-only path, no L1 allocation for data touched once
__device__ __forceinline__ uint4 ldg16_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void stg16_stream(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// the same with an L2 evict-first policy: the single-pass kernel streams 13 GB through L2 next to its warps' stacks,
// which should be what stays resident
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <bool NC>
__device__ __forceinline__ uint4 ldg16_ef(const void* p, uint64_t pol) {
    uint4 r;
    if (NC) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    else asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ void stg16_ef(void* p, uint4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

// 16 bytes starting at an arbitrary address: two aligned 16-byte loads + a 128-bit funnel shift.
// `need` = how many of the 16 bytes the caller will use; the second aligned word is touched only
// if it holds needed bytes, so no aligned word lying wholly outside [s, s+need) is ever read.
template <bool STREAM>
__device__ __forceinline__ uint4 load16_unaligned(const uint8_t* s, uint32_t need) {
    uint32_t off = (uint32_t)((uintptr_t)s & 15u);
    const uint8_t* base = s - off;
    uint4 lo = STREAM ? ldg16_stream(base) : ldg16(base);
    if (off == 0) return lo;
    uint4 hi = make_uint4(0, 0, 0, 0);
    if (off + need > 16) hi = STREAM ? ldg16_stream(base + 16) : ldg16(base + 16);
    uint64_t x0 = ((uint64_t)lo.y << 32) | lo.x, x1 = ((uint64_t)lo.w << 32) | lo.z;
    uint64_t x2 = ((uint64_t)hi.y << 32) | hi.x, x3 = ((uint64_t)hi.w << 32) | hi.z;
    if (off & 8) { x0 = x1; x1 = x2; x2 = x3; }
    uint32_t sh = (off & 7) * 8;
    uint64_t o0 = sh ? (x0 >> sh) | (x1 << (64 - sh)) : x0;
    uint64_t o1 = sh ? (x1 >> sh) | (x2 << (64 - sh)) : x1;
    return make_uint4((uint32_t)o0, (uint32_t)(o0 >> 32), (uint32_t)o1, (uint32_t)(o1 >> 32));
}

// Cooperative copy of n bytes, any alignment of dst/src. dst gets 16-byte aligned vector stores.
__device__ __noinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint64_t n) {
    int l = lane_id();
    if (n == 0) return;
    uint64_t head = (16 - ((uintptr_t)dst & 15u)) & 15u;
    if (head > n) head = n;
    if ((uint64_t)l < head) dst[l] = src[l];
    dst += head; src += head; n -= head;
    uint64_t nvec = n >> 4;
    for (uint64_t i = l; i < nvec; i += 32) {
        uint4 v = load16_unaligned<false>(src + (i << 4), 16);
        *reinterpret_cast<uint4*>(dst + (i << 4)) = v;
    }
    uint64_t done = nvec << 4;
    uint64_t tail = n - done;
    if ((uint64_t)l < tail) dst[done + l] = src[done + l];
}
// Streaming variant for the single-pass mode: four independent 16-byte loads in flight per lane, no-allocate
// stores; NC selects the read-only path (only legal when the source is not written by this kernel, i.e. the corpus).
template <bool NC>
__device__ __noinline__ void warp_copy_stream(uint8_t* dst, const uint8_t* src, uint64_t n) {
    int l = lane_id();
    if (n == 0) return;
    uint64_t head = (16 - ((uintptr_t)dst & 15u)) & 15u;
    if (head > n) head = n;
    if ((uint64_t)l < head) dst[l] = src[l];
    dst += head; src += head; n -= head;
    uint64_t nvec = n >> 4;
    uint64_t i = l;
    // 4 KiB tiles, eight aligned 16-byte loads in flight per lane. A misaligned source is realigned in registers:
    // each lane takes the following aligned word from its neighbour by shuffle (lane 31 from lane 0's next word,
    // and one extra load for the last), so every source byte crosses the memory pipe once.
    {
        const uint32_t off = (uint32_t)((uintptr_t)src & 15u);
        const uint8_t* base = src - off;
        const uint32_t sh = (off & 7u) * 8u;
        const int nl = (l + 1) & 31;
        const uint64_t pol = l2_evict_first_policy();
        for (; i - l + 256 <= nvec; i += 256) {          // warp-uniform trip count: the body shuffles
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = ldg16_ef<NC>(base + ((i + 32 * j) << 4), pol);
            if (off == 0) {
#pragma unroll
                for (int j = 0; j < 8; j++) stg16_ef(dst + ((i + 32 * j) << 4), v[j], pol);
                continue;
            }
            uint4 ex = make_uint4(0, 0, 0, 0);
            if (l == 31) ex = ldg16_ef<NC>(base + ((i + 225) << 4), pol);
            uint4 r = make_uint4(__shfl_sync(0xffffffffu, v[0].x, nl), __shfl_sync(0xffffffffu, v[0].y, nl),
                                 __shfl_sync(0xffffffffu, v[0].z, nl), __shfl_sync(0xffffffffu, v[0].w, nl));
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint4 rn = ex;                                   // rotation of the following tile row
                if (j < 7) rn = make_uint4(__shfl_sync(0xffffffffu, v[j + 1].x, nl), __shfl_sync(0xffffffffu, v[j + 1].y, nl),
                                           __shfl_sync(0xffffffffu, v[j + 1].z, nl), __shfl_sync(0xffffffffu, v[j + 1].w, nl));
                uint4 hi = (l == 31) ? rn : r;
                uint64_t x0 = ((uint64_t)v[j].y << 32) | v[j].x, x1 = ((uint64_t)v[j].w << 32) | v[j].z;
                uint64_t x2 = ((uint64_t)hi.y << 32) | hi.x, x3 = ((uint64_t)hi.w << 32) | hi.z;
                if (off & 8) { x0 = x1; x1 = x2; x2 = x3; }
                uint64_t o0 = sh ? (x0 >> sh) | (x1 << (64 - sh)) : x0;
                uint64_t o1 = sh ? (x1 >> sh) | (x2 << (64 - sh)) : x1;
                stg16_ef(dst + ((i + 32 * j) << 4), make_uint4((uint32_t)o0, (uint32_t)(o0 >> 32), (uint32_t)o1, (uint32_t)(o1 >> 32)), pol);
                r = rn;
            }
        }
    }
    for (; i + 96 < nvec; i += 128) {
        uint4 v0 = load16_unaligned<NC>(src + (i << 4), 16);
        uint4 v1 = load16_unaligned<NC>(src + ((i + 32) << 4), 16);
        uint4 v2 = load16_unaligned<NC>(src + ((i + 64) << 4), 16);
        uint4 v3 = load16_unaligned<NC>(src + ((i + 96) << 4), 16);
        stg16_stream(dst + (i << 4), v0); stg16_stream(dst + ((i + 32) << 4), v1);
        stg16_stream(dst + ((i + 64) << 4), v2); stg16_stream(dst + ((i + 96) << 4), v3);
    }
    for (; i < nvec; i += 32) stg16_stream(dst + (i << 4), load16_unaligned<NC>(src + (i << 4), 16));
    uint64_t done = nvec << 4;
    uint64_t tail = n - done;
    if ((uint64_t)l < tail) dst[done + l] = src[done + l];
}
__device__ __noinline__ void warp_fill(uint8_t* dst, uint8_t b, uint64_t n) {
    for (uint64_t i = lane_id(); i < n; i += 32) dst[i] = b;
}

// ------------------------------------------------------------------ byte scans
// A scan walks [p, p+n) in 512-byte warp chunks laid out on 16-byte ALIGNED coordinates
// (chunk 0 starts at p rounded down to 16). Each lane owns one aligned 16-byte word and reduces
// it to a 16-bit predicate mask with SIMD-in-a-word byte compares.
enum ScanPred { PRED_NEWLINE = 0, PRED_DIGIT = 1 };

__device__ __forceinline__ uint32_t bytemask4(uint32_t cmp) {   // 0xFF/0x00 per byte -> 4 bits
    return ((cmp & 0x08040201u) * 0x01010101u) >> 24;
}
// SWAR predicate on four packed bytes -> 4 result bits (bit i = byte i matches), ~7 integer ops:
//   digit   : t = b ^ 0x30 is < 10   <=>  bit 7 of ((t & 0x7f) + 0x76) | t is clear
//   newline : t = b ^ 0x0a is zero   <=>  bit 7 of ((t & 0x7f) + 0x7f) | t is clear
// (no carries cross byte lanes: 0x7f + 0x7f < 0x100); the four bit-7 flags are then gathered with one multiply.
template <int PRED>
__device__ __forceinline__ uint32_t pred_flags(uint32_t w) {   // bit 7 of each byte = predicate
    uint32_t t = w ^ (PRED == PRED_DIGIT ? 0x30303030u : 0x0a0a0a0au);
    uint32_t u = ((t & 0x7f7f7f7fu) + (PRED == PRED_DIGIT ? 0x76767676u : 0x7f7f7f7fu)) | t;
    return ~u & 0x80808080u;
}
template <int PRED>
__device__ __forceinline__ uint32_t pred4(uint32_t w) {
    uint32_t f = pred_flags<PRED>(w) >> 7;         // bits 0, 8, 16, 24
    return (f * 0x01020408u) >> 24;                // -> bits 0..3
}
template <int PRED>
__device__ __forceinline__ uint32_t pred_mask16(uint4 w) {
    return pred4<PRED>(w.x) | (pred4<PRED>(w.y) << 4) | (pred4<PRED>(w.z) << 8) | (pred4<PRED>(w.w) << 12);
}

constexpr uint32_t SC_SHIFT = 12;                 // superchunk = 4 KiB of aligned coordinates
constexpr uint32_t SC_MAX = 264;                  // covers a 1 000 000-byte block (+ alignment slack)

struct ScanCursor {
    const uint8_t* base;   // p rounded down to 16
    uint32_t lead;         // p - base
    uint32_t n;            // logical length
    uint32_t span;         // lead + n
};
__device__ __forceinline__ ScanCursor scan_cursor(const uint8_t* p, uint32_t n) {
    ScanCursor c; c.lead = (uint32_t)((uintptr_t)p & 15u); c.base = p - c.lead; c.n = n; c.span = c.lead + n; return c;
}
// predicate mask of this lane's word in warp chunk `it`; bytes outside [p,p+n) read as "no match".
// RUNSTART: a match only counts if the previous byte (in logical order) is not a match.
// this lane's aligned word of warp chunk `it` (zeros when the word lies outside [p, p+n))
__device__ __forceinline__ uint4 scan_chunk_load(const ScanCursor& c, uint32_t it) {
    uint32_t wofs = (it * 32 + lane_id()) * 16;     // aligned coordinate of this lane's word
    if (wofs < c.span && wofs + 16 > c.lead) return ldg16(c.base + wofs);
    return make_uint4(0, 0, 0, 0);
}
template <int PRED, bool RUNSTART>
__device__ __forceinline__ uint32_t scan_chunk_reduce(const ScanCursor& c, uint32_t it, uint4 w, uint32_t& carry) {
    uint32_t wofs = (it * 32 + lane_id()) * 16;
    uint32_t m = 0;
    if (wofs < c.span && wofs + 16 > c.lead) {
        m = pred_mask16<PRED>(w);
        if (wofs < c.lead) m &= 0xffffu << (c.lead - wofs);
        if (wofs + 16 > c.span) m &= 0xffffu >> (wofs + 16 - c.span);
    }
    if (RUNSTART) {
        uint32_t prev = __shfl_up_sync(0xffffffffu, m >> 15, 1);
        if (lane_id() == 0) prev = carry;
        carry = __shfl_sync(0xffffffffu, m >> 15, 31);
        m = m & ~((m << 1) | prev) & 0xffffu;
    }
    return m;
}
template <int PRED, bool RUNSTART>
__device__ __forceinline__ uint32_t scan_chunk_mask(const ScanCursor& c, uint32_t it, uint32_t& carry) {
    return scan_chunk_reduce<PRED, RUNSTART>(c, it, scan_chunk_load(c, it), carry);
}
// count matches; sc[k] receives the count inside superchunk k (4 KiB of aligned coordinates).
// One superchunk = 8 warp chunks: all 8 loads are issued before the first is consumed, so a warp
// keeps 4 KiB in flight instead of serialising on DRAM latency (profiles/decide_r1a: 27% of the
// decide kernel's stall samples sat on the first use of a single in-flight load).
// matches inside ONE superchunk (warp chunks it0 .. it0+7); `carry` = "the byte before this superchunk matches"
// on entry, the same for the next superchunk on return. Returns the warp-wide count (same value in every lane).
template <int PRED, bool RUNSTART>
__device__ __forceinline__ uint32_t scan_count_sc_w(const ScanCursor& c, uint32_t it0, uint32_t& carry, const uint4* w) {
    uint32_t acc = 0;
    if (it0 * 512u >= c.lead && (it0 + 8) * 512u <= c.span) {
        // interior superchunk: every byte is valid, so count straight on the bit-7 byte flags
        // (no validity masks, no packing): ~33 integer ops per 16-byte word instead of ~70
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t f0 = pred_flags<PRED>(w[j].x), f1 = pred_flags<PRED>(w[j].y), f2 = pred_flags<PRED>(w[j].z), f3 = pred_flags<PRED>(w[j].w);
            if (RUNSTART) {
                uint32_t top = f3 >> 24;                                  // last byte's flag, moved to bit 7
                uint32_t prev = __shfl_up_sync(0xffffffffu, top, 1);
                if (lane_id() == 0) prev = carry << 7;
                carry = __shfl_sync(0xffffffffu, top, 31) >> 7;
                uint32_t s0 = f0 & ~((f0 << 8) | prev), s1 = f1 & ~((f1 << 8) | (f0 >> 24)), s2 = f2 & ~((f2 << 8) | (f1 >> 24)), s3 = f3 & ~((f3 << 8) | (f2 >> 24));
                acc += __popc((s0 >> 7) | (s1 >> 6) | (s2 >> 5) | (s3 >> 4));
            } else {
                acc += __popc((f0 >> 7) | (f1 >> 6) | (f2 >> 5) | (f3 >> 4));
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) acc += __popc(scan_chunk_reduce<PRED, RUNSTART>(c, it0 + j, w[j], carry));
    }
    return warp_sum(acc);
}
template <int PRED, bool RUNSTART>
__device__ __forceinline__ uint32_t scan_count_sc(const ScanCursor& c, uint32_t it0, uint32_t& carry) {
    uint4 w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) w[j] = scan_chunk_load(c, it0 + j);
    return scan_count_sc_w<PRED, RUNSTART>(c, it0, carry, w);
}
// carry-in of superchunk s > 0 computed from the data alone: is the last byte before it a (valid) match?
template <int PRED>
__device__ __forceinline__ uint32_t scan_carry_in(const ScanCursor& c, uint32_t s) {
    if (s == 0) return 0;
    uint32_t wofs = s * 4096u - 16u;
    uint32_t m = 0;
    if (wofs + 16 > c.lead && wofs < c.span) { uint4 w = ldg16(c.base + wofs); m = pred_mask16<PRED>(w); if (wofs < c.lead) m &= 0xffffu << (c.lead - wofs); if (wofs + 16 > c.span) m &= 0xffffu >> (wofs + 16 - c.span); }
    return m >> 15;
}
// count matches; sc[k] receives the count inside superchunk k (4 KiB of aligned coordinates).
// One superchunk = 8 warp chunks: all 8 loads are issued before the first is consumed, so a warp
// keeps 4 KiB in flight instead of serialising on DRAM latency (profiles/decide_r1a: 27% of the
// decide kernel's stall samples sat on the first use of a single in-flight load).
template <int PRED, bool RUNSTART>
__device__ __noinline__ uint32_t scan_count(const uint8_t* p, uint32_t n, uint16_t* sc) {
    ScanCursor c = scan_cursor(p, n);
    uint32_t iters = (c.span + 511) >> 9, carry = 0, total = 0;
    for (uint32_t it0 = 0; it0 < iters; it0 += 8) {
        uint32_t s = scan_count_sc<PRED, RUNSTART>(c, it0, carry); total += s;
        if (lane_id() == 0) sc[it0 >> 3] = (uint16_t)s;
    }
    __syncwarp();
    return total;
}
// logical index of the k-th (0-based) match; requires sc[] from scan_count of the same predicate
template <int PRED, bool RUNSTART>
__device__ __noinline__ uint32_t scan_select(const uint8_t* p, uint32_t n, const uint16_t* sc, uint32_t k) {
    ScanCursor c = scan_cursor(p, n);
    uint32_t iters = (c.span + 511) >> 9;
    uint32_t s = 0, before = 0;
    while (before + sc[s] <= k) { before += sc[s]; s++; }    // warp-uniform walk over <= 264 entries
    uint32_t carry = 0;
    uint32_t it0 = s << 3;
    if (RUNSTART && it0 > 0) {   // carry-in: is the last byte of the previous superchunk a match?
        uint32_t wofs = it0 * 512 - 16;
        uint32_t m = 0;
        if (wofs + 16 > c.lead) { uint4 w = ldg16(c.base + wofs); m = pred_mask16<PRED>(w); if (wofs < c.lead) m &= 0xffffu << (c.lead - wofs); }
        carry = m >> 15;
    }
    for (uint32_t it = it0; it < iters; it++) {
        uint32_t m = scan_chunk_mask<PRED, RUNSTART>(c, it, carry);
        uint32_t cnt = __popc(m), tot;
        uint32_t ex = warp_excl_scan(cnt, tot);
        if (before + tot > k) {
            uint32_t want = k - before;                       // rank inside this warp chunk
            uint32_t hit = (want >= ex && want < ex + cnt) ? 1u : 0u;
            uint32_t pos = 0;
            if (hit) { uint32_t r = want - ex, mm = m; for (uint32_t j = 0; j < r; j++) mm &= mm - 1; pos = (it * 32 + lane_id()) * 16 + (__ffs(mm) - 1) - c.lead; }
            uint32_t who = __ffs(__ballot_sync(0xffffffffu, hit)) - 1;
            return __shfl_sync(0xffffffffu, pos, who);
        }
        before += tot;
    }
    return n;   // not found (callers never ask beyond the count)
}

}  // namespace eb
