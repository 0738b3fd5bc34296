// erlamsa_b200 -- the CTA's job queue: DECIDER warps run the scalar per-case program and hand the bulk byte work
// (streaming copies of >= 16 KiB, digit-run / newline counts over whole blocks) to WORKER warps of the same CTA.
//
// Why (profiles/fused_r1c.txt): with one warp doing everything for its case, (a) `num` read its block
// twice from DRAM -- 300 MB of blocks in flight chip-wide never survive in the 126 MB L2 between the count and the copy;
// (b) the copy of a case started only after its own decision, so a CTA-wide barrier per case was needed to keep the
// big scalar program in the instruction cache, and the slowest case of each round set the pace (25 % barrier stalls).
// With the split, a block's count and copy are executed by many warps within microseconds of each other (the second
// read hits L2), the copy loop is a tiny program that never leaves the instruction cache, and cases are taken from a
// global counter, so nobody waits for anybody else's case.
//
// The queue is a bounded multi-producer / multi-consumer ring in shared memory (one sequence word per cell: free for
// position p when seq == p, ready when seq == p + 1). Producers and consumers claim positions with one shared-memory
// atomic; workers never wait for anything but the publication of a claimed cell, so the ring cannot deadlock.
#pragma once
#include "eb_warp.cuh"

namespace eb {

constexpr uint32_t QCAP = 256;          // cells in the ring
constexpr uint32_t JOB_TILE = 8192;     // bytes per copy job (two 4 KiB register tiles)
constexpr uint32_t JOB_MIN_COPY = 16384; // shorter copies are executed inline by the deciding warp
constexpr uint32_t JOB_MIN_SCAN = 16384; // shorter scans likewise

enum JobKind : uint32_t { JOB_COPY_NC = 0, JOB_COPY = 1, JOB_COUNT_DIGIT = 2, JOB_COUNT_NL = 3, JOB_EDIT = 4, JOB_SELECT_DIGIT = 5, JOB_COUNTCOPY_DIGIT = 6 };

struct __align__(16) Job {
    uint64_t a;       // copy: destination | count: aligned base of the scanned block
    uint64_t b;       // copy: source      | count: (lead << 32) | span of the scan cursor
    uint32_t len;     // copy: bytes       | count: superchunk index
    uint32_t kind;
    uint32_t res;     // count: shared-window address of the uint16 result slot
    uint32_t pend;    // count: shared-window address of the poster's pending counter
};

struct JobQ {
    Job jobs[QCAP];
    uint32_t seq[QCAP];
    unsigned int head, tail;
    uint32_t closed;
    unsigned int deciders_left;
};

__device__ __forceinline__ uint32_t ld_shared_volatile(const uint32_t* p) { return *(const volatile uint32_t*)p; }

__device__ __forceinline__ void jobq_init(JobQ* q, int deciders) {
    for (uint32_t i = threadIdx.x; i < QCAP; i += blockDim.x) q->seq[i] = i;
    if (threadIdx.x == 0) { q->head = 0; q->tail = 0; q->closed = 0; q->deciders_left = (unsigned)deciders; }
}
// claim `cnt` consecutive positions (warp-collective); every lane gets the first one
__device__ __forceinline__ uint32_t jobq_claim(JobQ* q, uint32_t cnt) {
    uint32_t pos = 0;
    if (lane_id() == 0) pos = atomicAdd(&q->tail, cnt);
    return __shfl_sync(0xffffffffu, pos, 0);
}
// publish one job at a claimed position (any single lane)
__device__ __forceinline__ void jobq_put(JobQ* q, uint32_t pos, const Job& j) {
    uint32_t cell = pos % QCAP;
    while (ld_shared_volatile(&q->seq[cell]) != pos) __nanosleep(100);
    q->jobs[cell] = j;
    __threadfence_block();
    *(volatile uint32_t*)&q->seq[cell] = pos + 1;
}
// take the next job (warp-collective); false when the queue is closed and drained
__device__ __forceinline__ bool jobq_get(JobQ* q, Job& j) {
    uint32_t pos = 0;
    if (lane_id() == 0) pos = atomicAdd(&q->head, 1u);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    uint32_t cell = pos % QCAP;
    int ok = 1;
    if (lane_id() == 0) {
        while (ld_shared_volatile(&q->seq[cell]) != pos + 1) {
            if (ld_shared_volatile(&q->closed) && (int32_t)(pos - *(volatile unsigned int*)&q->tail) >= 0) { ok = 0; break; }
            __nanosleep(200);
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) return false;
    __threadfence_block();
    j = q->jobs[cell];
    __syncwarp();
    if (lane_id() == 0) *(volatile uint32_t*)&q->seq[cell] = pos + QCAP;
    return true;
}
// a decider is done with its last case
__device__ __forceinline__ void jobq_decider_done(JobQ* q) {
    __syncwarp();
    if (lane_id() == 0) {
        __threadfence_block();
        if (atomicSub(&q->deciders_left, 1u) == 1u) { __threadfence_block(); *(volatile uint32_t*)&q->closed = 1u; }
    }
}

template <int PRED, bool RUNSTART>
__device__ __forceinline__ void job_count(const Job& j) {
    ScanCursor c; c.base = (const uint8_t*)(uintptr_t)j.a; c.lead = (uint32_t)(j.b >> 32); c.span = (uint32_t)j.b; c.n = c.span - c.lead;
    uint32_t carry = RUNSTART ? scan_carry_in<PRED>(c, j.len) : 0;
    uint32_t s = scan_count_sc<PRED, RUNSTART>(c, j.len * 8u, carry);
    if (lane_id() == 0) {
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(j.res), "h"((unsigned short)s) : "memory");
        __threadfence_block();
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(j.pend), "r"(0xffffffffu) : "memory");
    }
}

// EDIT job (posted by the front warps, eb_fast.cuh): out = src[0,pos) ++ <ll literal bytes> ++ src[pos+skip, n); src is a
// corpus blob; the literal bytes are written by the poster itself.
// a = dst, b = src, len = n, kind = JOB_EDIT | ll << 8 | skip << 16, res = pos
__device__ __forceinline__ void job_edit(const Job& j) {
    uint8_t* dst = (uint8_t*)(uintptr_t)j.a; const uint8_t* src = (const uint8_t*)(uintptr_t)j.b;
    uint32_t n = j.len, pos = j.res, ll = (j.kind >> 8) & 255u, skip = j.kind >> 16;
    warp_copy_stream<true>(dst, src, pos);
    warp_copy_stream<true>(dst + pos + ll, src + pos + skip, n - pos - skip);
}
// per-lane scratch of a front warp (shared memory): superchunk counts, countdown, selected position
constexpr uint32_t FRONT_SC = 32;
struct FrontState { uint16_t sc[32][FRONT_SC]; uint32_t pend[32]; uint32_t sel[32]; };
// SELECT job: logical index of the k-th digit-run start of a block whose per-superchunk counts are in the poster's
// FrontState. a = aligned base, b = lead << 32 | span, len = k, kind = JOB_SELECT_DIGIT | lane << 8, res = shared-window
// address of the FrontState
__device__ __forceinline__ void job_select(const Job& j) {
    uint32_t lead = (uint32_t)(j.b >> 32), span = (uint32_t)j.b, lane = (j.kind >> 8) & 31u;
    const uint8_t* p = (const uint8_t*)(uintptr_t)j.a + lead;
    FrontState* fs = (FrontState*)__cvta_shared_to_generic((size_t)j.res);
    uint32_t pos = scan_select<PRED_DIGIT, true>(p, span - lead, fs->sc[lane], j.len);
    if (lane_id() == 0) {
        *(volatile uint32_t*)&fs->sel[lane] = pos;
        __threadfence_block();
        atomicSub(&fs->pend[lane], 1u);
    }
}

// COUNT + COPY job (front warps, sed_num): count the digit runs of one 4 KiB superchunk of a block AND copy it, byte for
// byte at the same 16-byte phase, into the case's output slot. The prefix of a sed_num result is the prefix of its input
// at the same offsets, so this speculative copy is already final for everything before the edited number; after the
// decision only the SHORTER side is copied again at its shifted position (eb_fast.cuh) -- the block crosses HBM once.
// a = aligned source base, b = destination of aligned coordinate 0, len = superchunk, pend = span,
// kind = JOB_COUNTCOPY_DIGIT | lane << 8 | lead << 16, res = shared-window address of the poster's FrontState
__device__ __forceinline__ void job_countcopy(const Job& j) {
    uint32_t lane = (j.kind >> 8) & 31u;
    ScanCursor c; c.base = (const uint8_t*)(uintptr_t)j.a; c.lead = (j.kind >> 16) & 15u; c.span = j.pend; c.n = c.span - c.lead;
    uint8_t* dbase = (uint8_t*)(uintptr_t)j.b;
    FrontState* fs = (FrontState*)__cvta_shared_to_generic((size_t)j.res);
    uint32_t it0 = j.len * 8u;
    uint4 w[8];
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = scan_chunk_load(c, it0 + k);
    uint32_t carry = scan_carry_in<PRED_DIGIT>(c, j.len);
    uint32_t s = scan_count_sc_w<PRED_DIGIT, true>(c, it0, carry, w);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t wofs = ((it0 + k) * 32u + (uint32_t)lane_id()) * 16u;
        if (wofs >= c.span || wofs + 16 <= c.lead) continue;
        if (wofs >= c.lead && wofs + 16 <= c.span) *reinterpret_cast<uint4*>(dbase + wofs) = w[k];
        else {
            uint32_t lo = wofs < c.lead ? c.lead - wofs : 0, hi = wofs + 16 > c.span ? c.span - wofs : 16;
            uint32_t ww[4] = {w[k].x, w[k].y, w[k].z, w[k].w};
            for (uint32_t i = lo; i < hi; i++) dbase[wofs + i] = (uint8_t)(ww[i >> 2] >> (8 * (i & 3)));
        }
    }
    // every lane's stores must be ordered before the countdown the poster waits on; poster and the workers that later
    // overwrite part of this copy are warps of the same CTA, so CTA scope is enough (a gpu-scope fence here was 18 % of all
    // stall samples, profiles/fused_r2b.txt)
    __threadfence_block();
    __syncwarp();
    if (lane_id() == 0) {
        *(volatile uint16_t*)&fs->sc[lane][j.len] = (uint16_t)s;
        __threadfence_block();
        atomicSub(&fs->pend[lane], 1u);
    }
}

// ---- bulk-async (TMA) variant of the streaming copy, for A/B runs (EB200_TMA_WORKERS; profiles/variants_r2.txt):
// a range whose source and destination share their 16-byte phase goes global -> shared -> global with cp.async.bulk
// (SASS: UBLKCP), 8 KiB tiles, two staging buffers per worker, completion through an mbarrier (loads) and bulk groups
// (stores); one elected lane issues everything, no payload byte touches a register. Other ranges take the register path.
constexpr uint32_t TMA_TILE = 8192;
struct TmaStage { uint8_t buf[2][TMA_TILE]; uint64_t mbar[2]; uint32_t phase[2]; uint32_t inited; uint32_t pad; };
__device__ __forceinline__ void tma_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_store(void* gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
}
__device__ __noinline__ void tma_copy(TmaStage* st, uint8_t* dst, const uint8_t* src, uint64_t n) {
    if ((((uintptr_t)dst ^ (uintptr_t)src) & 15u) != 0 || n < 2 * TMA_TILE) { warp_copy_stream<true>(dst, src, n); return; }
    const int l = lane_id();
    uint64_t head = (16 - ((uintptr_t)dst & 15u)) & 15u;
    if ((uint64_t)l < head) dst[l] = src[l];
    dst += head; src += head; n -= head;
    uint64_t body = n & ~15ull, ntiles = (body + TMA_TILE - 1) / TMA_TILE;
    if (l == 0) {
        uint32_t b0 = (uint32_t)__cvta_generic_to_shared(&st->buf[0][0]), b1 = (uint32_t)__cvta_generic_to_shared(&st->buf[1][0]);
        uint32_t m0 = (uint32_t)__cvta_generic_to_shared(&st->mbar[0]), m1 = (uint32_t)__cvta_generic_to_shared(&st->mbar[1]);
        auto tile_bytes = [&](uint64_t t) { uint64_t o = t * TMA_TILE; return (uint32_t)(body - o < TMA_TILE ? body - o : TMA_TILE); };
        tma_load(b0, src, tile_bytes(0), m0);
        for (uint64_t t = 0; t < ntiles; t++) {
            uint32_t cur = (uint32_t)(t & 1);
            if (t + 1 < ntiles) {
                // the other buffer was the source of store t-1: its bytes must have left shared memory before it is refilled
                if (t >= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                tma_load(cur ? b0 : b1, src + (t + 1) * TMA_TILE, tile_bytes(t + 1), cur ? m0 : m1);
            }
            mbar_wait_parity(cur ? m1 : m0, st->phase[cur]); st->phase[cur] ^= 1u;
            tma_store(dst + t * TMA_TILE, cur ? b1 : b0, tile_bytes(t));
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
    uint64_t tail = n - body;
    if ((uint64_t)l < tail) dst[body + l] = src[body + l];
}
__device__ __forceinline__ void job_edit_tma(TmaStage* st, const Job& j) {
    uint8_t* dst = (uint8_t*)(uintptr_t)j.a; const uint8_t* src = (const uint8_t*)(uintptr_t)j.b;
    uint32_t n = j.len, pos = j.res, ll = (j.kind >> 8) & 255u, skip = j.kind >> 16;
    tma_copy(st, dst, src, pos);
    tma_copy(st, dst + pos + ll, src + pos + skip, n - pos - skip);
}

// the worker warps' whole program (st != nullptr: this worker owns a bulk-async staging area)
__device__ __noinline__ void worker_loop(JobQ* q, TmaStage* st) {
    Job j;
    if (st && lane_id() == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&st->mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&st->mbar[1])) : "memory");
        st->phase[0] = st->phase[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    while (jobq_get(q, j)) {
        switch (j.kind & 255u) {
        case JOB_EDIT: if (st) job_edit_tma(st, j); else job_edit(j); break;
        case JOB_SELECT_DIGIT: job_select(j); break;
        case JOB_COUNTCOPY_DIGIT: job_countcopy(j); break;
        case JOB_COPY_NC: if (st) tma_copy(st, (uint8_t*)(uintptr_t)j.a, (const uint8_t*)(uintptr_t)j.b, j.len); else warp_copy_stream<true>((uint8_t*)(uintptr_t)j.a, (const uint8_t*)(uintptr_t)j.b, j.len); break;
        case JOB_COPY: warp_copy_stream<false>((uint8_t*)(uintptr_t)j.a, (const uint8_t*)(uintptr_t)j.b, j.len); break;
        case JOB_COUNT_DIGIT: job_count<PRED_DIGIT, true>(j); break;
        default: job_count<PRED_NEWLINE, false>(j); break;
        }
        __syncwarp();
    }
}

// ---- decider side
// stream n bytes src -> dst: inline when short or when the kernel runs without workers, else as tile jobs whose
// destinations (after the first) start on 16-byte boundaries
__device__ __noinline__ void post_copy(JobQ* q, uint8_t* dst, const uint8_t* src, uint64_t n, bool nc) {
    if (!q || n < JOB_MIN_COPY) {
        if (nc) warp_copy_stream<true>(dst, src, n); else warp_copy_stream<false>(dst, src, n);
        return;
    }
    uint64_t b1 = JOB_TILE + ((16 - ((uintptr_t)dst & 15u)) & 15u);
    uint64_t ntiles = 1 + (n - b1 + JOB_TILE - 1) / JOB_TILE;
    for (uint64_t base = 0; base < ntiles; base += 32) {
        uint32_t cnt = (uint32_t)(ntiles - base < 32 ? ntiles - base : 32);
        uint32_t pos0 = jobq_claim(q, cnt);
        uint32_t l = (uint32_t)lane_id();
        if (l < cnt) {
            uint64_t t = base + l;
            uint64_t s = t == 0 ? 0 : b1 + (t - 1) * JOB_TILE;
            uint64_t e = b1 + t * JOB_TILE; if (e > n) e = n;
            Job j; j.a = (uint64_t)(uintptr_t)(dst + s); j.b = (uint64_t)(uintptr_t)(src + s); j.len = (uint32_t)(e - s); j.kind = nc ? JOB_COPY_NC : JOB_COPY; j.res = 0; j.pend = 0;
            jobq_put(q, pos0 + l, j);
        }
        __syncwarp();
    }
}

// scan_count through the workers: one job per 4 KiB superchunk, results land in sc[] (shared memory), the poster
// waits on a countdown word next to them
template <int PRED, bool RUNSTART>
__device__ __noinline__ uint32_t scan_count_jobs(JobQ* q, const uint8_t* p, uint32_t n, uint16_t* sc, uint32_t* pend) {
    if (!q || n < JOB_MIN_SCAN) return scan_count<PRED, RUNSTART>(p, n, sc);
    ScanCursor c = scan_cursor(p, n);
    uint32_t nsc = (c.span + 4095u) >> 12;
    if (lane_id() == 0) *(volatile uint32_t*)pend = nsc;
    __syncwarp();
    __threadfence_block();
    uint32_t pend_sa = (uint32_t)__cvta_generic_to_shared(pend), sc_sa = (uint32_t)__cvta_generic_to_shared(sc);
    for (uint32_t base = 0; base < nsc; base += 32) {
        uint32_t cnt = nsc - base < 32 ? nsc - base : 32;
        uint32_t pos0 = jobq_claim(q, cnt);
        uint32_t l = (uint32_t)lane_id();
        if (l < cnt) {
            Job j; j.a = (uint64_t)(uintptr_t)c.base; j.b = ((uint64_t)c.lead << 32) | c.span; j.len = base + l;
            j.kind = PRED == PRED_DIGIT ? JOB_COUNT_DIGIT : JOB_COUNT_NL; j.res = sc_sa + 2u * (base + l); j.pend = pend_sa;
            jobq_put(q, pos0 + l, j);
        }
        __syncwarp();
    }
    if (lane_id() == 0) while (ld_shared_volatile(pend) != 0) __nanosleep(200);
    __syncwarp();
    __threadfence_block();
    uint32_t acc = 0;
    for (uint32_t i = lane_id(); i < nsc; i += 32) acc += ((volatile uint16_t*)sc)[i];
    return warp_sum(acc);
}

}  // namespace eb
