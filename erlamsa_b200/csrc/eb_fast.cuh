// erlamsa_b200 -- the FRONT warps: one LANE per test case for the part of the per-case program that is pure RNG
// arithmetic, and for the single-byte mutators whose whole decision is "position + literal".
//
// Why: the general per-case program (eb_decide.cuh, one warp per case) is ~190 KB of SASS walked once per case; measured
// free-running it costs ~100 us per case and warp, almost all of it instruction fetch (round 2: 8 deciders per SM gave
// 9.6 ms per C3 step, 16 gave 7.2 ms). A case whose pattern is `od`, whose block list is the one corpus blob and whose
// first scheduled mutator is one of bd bei bed bf bi ber br uw never needs that program: thread seed, generator,
// pattern choice, Ip, the weighted permutation and the mutator's own draws are ~25 AS183 steps and ONE byte of the
// blob. Here a warp decides 32 such cases at once (each lane its own case, identical draw order as the general path,
// reference file:line cited there), posts one EDIT job per case to the CTA's workers, and hands every other case to
// the general deciders through the slow ring -- which therefore see (on C3) only `num` cases and walk one code path.
#pragma once
#include "eb_decide.cuh"

namespace eb {

constexpr uint32_t SLOWCAP = 128;
struct SlowRing {
    uint64_t ids[SLOWCAP];
    uint32_t seq[SLOWCAP];
    unsigned int head, tail;
    uint32_t closed;
    unsigned int fronts_left;
};
__device__ __forceinline__ void slow_init(SlowRing* r, int fronts) {
    for (uint32_t i = threadIdx.x; i < SLOWCAP; i += blockDim.x) r->seq[i] = i;
    if (threadIdx.x == 0) { r->head = 0; r->tail = 0; r->closed = 0; r->fronts_left = (unsigned)fronts; }
}
// any single lane
__device__ __forceinline__ void slow_push(SlowRing* r, uint64_t k) {
    uint32_t pos = atomicAdd(&r->tail, 1u), cell = pos % SLOWCAP;
    while (ld_shared_volatile(&r->seq[cell]) != pos) __nanosleep(100);
    r->ids[cell] = k;
    __threadfence_block();
    *(volatile uint32_t*)&r->seq[cell] = pos + 1;
}
// warp-collective; false when closed and drained
__device__ __forceinline__ bool slow_pop(SlowRing* r, uint64_t& k) {
    uint32_t pos = 0;
    if (lane_id() == 0) pos = atomicAdd(&r->head, 1u);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    uint32_t cell = pos % SLOWCAP;
    int ok = 1;
    if (lane_id() == 0) {
        while (ld_shared_volatile(&r->seq[cell]) != pos + 1) {
            if (ld_shared_volatile(&r->closed) && (int32_t)(pos - *(volatile unsigned int*)&r->tail) >= 0) { ok = 0; break; }
            __nanosleep(200);
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) return false;
    __threadfence_block();
    k = r->ids[cell];
    __syncwarp();
    if (lane_id() == 0) *(volatile uint32_t*)&r->seq[cell] = pos + SLOWCAP;
    return true;
}
__device__ __forceinline__ void slow_front_done(SlowRing* r) {
    __syncwarp();
    if (lane_id() == 0) {
        __threadfence_block();
        if (atomicSub(&r->fronts_left, 1u) == 1u) { __threadfence_block(); *(volatile uint32_t*)&r->closed = 1u; }
    }
}

// can the front decide cases of this batch at all? (direct generator, `od` selectable, a byte mutator selected)
__host__ __device__ inline bool fast_byte_mut(int id) {
    switch (id) { case M_BD: case M_BEI: case M_BED: case M_BF: case M_BI: case M_BER: case M_BR: case M_UW: return true; default: return false; }
}

struct FastOut { uint32_t pos; uint32_t skip; uint32_t ll; uint32_t name; uint8_t lit[88]; };

// lane-level post of a streaming copy (tile jobs; destinations after the first start on 16-byte boundaries)
__device__ __forceinline__ void post_copy_lane(JobQ* q, uint8_t* dst, const uint8_t* src, uint32_t n) {
    uint32_t s = 0;
    while (s < n) {
        uint32_t e = s + JOB_TILE + ((16u - ((uint32_t)(uintptr_t)(dst + s) & 15u)) & 15u);
        if (e > n || n - e < 1024) e = n;
        Job j; j.a = (uint64_t)(uintptr_t)(dst + s); j.b = (uint64_t)(uintptr_t)(src + s); j.len = e - s; j.kind = JOB_COPY_NC; j.res = 0; j.pend = 0;
        uint32_t pos = atomicAdd(&q->tail, 1u);
        jobq_put(q, pos, j);
        s = e;
    }
}

constexpr uint32_t NUM_PAD = 96;    // room in front of the speculative copy for a result that grows at the front

// sed_num for one lane (reference src/erlamsa_mutations.erl:114-169; same draws as mut_num). The workers count the digit
// runs of the block while copying it into the slot (COUNTCOPY jobs), select the number (SELECT job); the arithmetic is
// here; then the shorter of prefix / suffix is copied again at its shifted place. dst0 = 16-byte aligned address inside the
// slot that stands for the block's aligned coordinate 0. On success *out_start = where the result begins.
__device__ __noinline__ bool fast_num_lane(Rng& g, JobQ* q, FrontState* fs, const uint8_t* blob, uint32_t blen, uint8_t* dst0, FastOut& fo, uint8_t** out_start) {
    const uint32_t l = (uint32_t)lane_id();
    if (blen < 2304) return false;                      // shorter blocks: "did the head block change" needs the general path
    ScanCursor c = scan_cursor(blob, blen);
    uint32_t nsc = (c.span + 4095u) >> 12;
    if (nsc > FRONT_SC) return false;
    uint32_t fs_sa = (uint32_t)__cvta_generic_to_shared(fs);
    *(volatile uint32_t*)&fs->pend[l] = nsc;
    __threadfence_block();
    uint32_t pos0 = atomicAdd(&q->tail, nsc);
    for (uint32_t i = 0; i < nsc; i++) {
        Job j; j.a = (uint64_t)(uintptr_t)c.base; j.b = (uint64_t)(uintptr_t)dst0; j.len = i; j.kind = JOB_COUNTCOPY_DIGIT | (l << 8) | (c.lead << 16); j.res = fs_sa; j.pend = c.span;
        jobq_put(q, pos0 + i, j);
    }
    while (ld_shared_volatile(&fs->pend[l]) != 0) __nanosleep(200);
    __threadfence_block();
    uint32_t nfound = 0;
    for (uint32_t i = 0; i < nsc; i++) nfound += ((volatile uint16_t*)fs->sc[l])[i];
    uint64_t which = g.rand(nfound);
    uint8_t* x = dst0 + c.lead;                        // the speculative copy of the block starts here
    if (nfound == 0) { (void)g.rand(10); fo.pos = 0; fo.skip = 0; fo.ll = 0; fo.name = M_NUM; *out_start = x; return true; }
    *(volatile uint32_t*)&fs->pend[l] = 1;
    __threadfence_block();
    { Job j; j.a = (uint64_t)(uintptr_t)c.base; j.b = ((uint64_t)c.lead << 32) | c.span; j.len = nfound - 1 - (uint32_t)which; j.kind = JOB_SELECT_DIGIT | (l << 8); j.res = fs_sa; j.pend = 0;
      uint32_t pos = atomicAdd(&q->tail, 1u); jobq_put(q, pos, j); }
    while (ld_shared_volatile(&fs->pend[l]) != 0) __nanosleep(200);
    __threadfence_block();
    uint32_t d0 = *(volatile uint32_t*)&fs->sel[l];
    uint32_t a = d0; while (a > 0 && blob[a - 1] == '-') a--;
    uint32_t b = d0; while (b < blen && (uint32_t)(blob[b] - '0') < 10u) b++;
    if (b - d0 > 77 || b - a > 64) return false;        // wider numbers / long dash runs: general path (flags what it cannot hold)
    Big256 v; v.zero();
    for (uint32_t i = d0; i < b; i++) { v.mul_small(10); v.add_small((uint32_t)(blob[i] - '0')); }
    if (a < d0 && !v.is_zero()) v.neg = 1;
    mutate_num(g, v);
    if (v.ovf) return false;
    uint32_t dl = (uint32_t)v.to_decimal(fo.lit);
    fo.ll = dl; fo.pos = a; fo.skip = b - a; fo.name = M_NUM;
    // result = blob[0,a) ++ lit ++ blob[b,n): keep the longer side of the speculative copy, redo the shorter one
    int32_t d = (int32_t)dl - (int32_t)(b - a);
    if (a <= blen - b) {   // prefix moves by -d, the literal ends where the old number ended
        uint8_t* start = x - d;
        if (d != 0) post_copy_lane(q, start, blob, a);
        for (uint32_t i = 0; i < dl; i++) x[b - dl + i] = fo.lit[i];
        *out_start = start;
    } else {               // suffix moves by +d
        if (d != 0) post_copy_lane(q, x + a + dl, blob + b, blen - b);
        for (uint32_t i = 0; i < dl; i++) x[a + i] = fo.lit[i];
        *out_start = x;
    }
    return true;
}

// One lane, one case. Returns true when the case is fully decided here: out = blob[0,pos) ++ lit[0,ll) ++ blob[pos+skip, n).
// Draw order = decide_one_case -> generate -> run_case_machine(P_OD) -> mux_fuzzers -> mut_byte / mut_num, single-block case.
__device__ __forceinline__ bool fast_decide_lane(const BatchParams& bp, Rng& g, JobQ* q, FrontState* fs, uint32_t blen, const uint8_t* blob, uint8_t* slot, uint64_t cap, FastOut& fo, uint8_t** num_start) {
    if (bp.generator != 0 || blen == 0 || blen > ABSMAX_BINARY_BLOCK) return false;     // random / file / stdin generators: general path
    g.set_slot(SLOT_PATTERN);
    (void)g.rand((uint64_t)bp.rbs_bound);                                   // direct_generator: unused rand_block_size
    if (g.rand((uint64_t)blen + 1) == blen) return false;                    // finish/1 appends a random tail: general path
    int pat = -1;
    { int64_t x = (int64_t)g.rand((uint64_t)bp.pat_sum);                     // mux_patterns + choose_pri
      for (int i = 0; i < bp.n_pats; i++) { if (x == 0 || x < bp.pat_pri[i]) { pat = bp.pat_id[i]; break; } x -= bp.pat_pri[i]; } }
    if (pat != P_OD) return false;
    uint64_t ip = g.rand(INITIAL_IP);                                        // mutate_once/4
    (void)g.rand(ip);                                                        // mutate_once_loop: one draw, then the only block is it
    // weighted_permutations: key_i = rand(trunc(Score_i * Pri_i)) in table order; the stable descending sort puts the
    // first row with the largest key in front
    g.set_slot((1u << 8) | SLOT_SCHED);
    int best = -1; uint64_t bestk = 0;
    for (int i = 0; i < bp.n_rows; i++) {
        uint64_t k = g.rand((uint64_t)trunc((double)bp.row_score[i] * (double)bp.row_pri[i]));
        if (best < 0 || k > bestk) { best = i; bestk = k; }
    }
    if (best < 0) return false;
    int id = bp.row_id[best];
    g.set_slot((1u << 8) | (uint32_t)id);
    if (id == M_NUM) { if (cap < (uint64_t)blen + NUM_PAD + 112) return false; return fast_num_lane(g, q, fs, blob, blen, slot + NUM_PAD, fo, num_start); }
    if (!fast_byte_mut(id)) return false;
    // mut_byte (sed_byte_* / sed_utf8_widen)
    uint32_t pos = (uint32_t)g.rand(blen);
    (void)g.rand_delta();
    uint32_t b = blob[pos];
    uint32_t lit = 0, ll = 0;
    switch (id) {
    case M_BD: break;
    case M_BEI: lit = (b + 1) & 255; ll = 1; break;
    case M_BED: lit = (b - 1) & 255; ll = 1; break;
    case M_BR: lit = b | (b << 8); ll = 2; break;
    case M_BF: lit = b ^ (1u << g.rand(8)); ll = 1; break;
    case M_BI: lit = (uint32_t)g.rand(256) | (b << 8); ll = 2; break;
    case M_BER: lit = (uint32_t)g.rand(256); ll = 1; if (lit == b) return false; break;   // unchanged block = "failed": next mutator runs
    default: if (b == (b & 0x3f)) { lit = 0xc0u | ((b | 0x80u) << 8); ll = 2; } else return false; break;   // M_UW
    }
    fo.pos = pos; fo.skip = 1; fo.ll = ll; fo.name = (uint32_t)id; fo.lit[0] = (uint8_t)lit; fo.lit[1] = (uint8_t)(lit >> 8);
    return true;
}

// the front warps' whole program: 32 cases per grab
__device__ __noinline__ void front_loop(const BatchParams& bp, const DecideArgs& a, const FusedArgs& fa, JobQ* q, SlowRing* slow, const uint32_t* pw, FrontState* fs) {
    const uint64_t total = fa.case_list ? fa.n_list : bp.n_cases;
    const int l = lane_id();
    for (;;) {
        unsigned long long idx0 = 0;
        if (l == 0) idx0 = atomicAdd(fa.case_counter, 32ull);
        idx0 = __shfl_sync(0xffffffffu, idx0, 0);
        if (idx0 >= total) break;
        unsigned long long idx = idx0 + (unsigned)l;
        if (idx < total) {
            uint64_t k = fa.case_list ? (uint64_t)fa.case_list[idx] : (uint64_t)idx;
            uint64_t I = bp.first_case + k, e = I - 1;
            uint32_t a1 = (uint32_t)bp.parent_a1, a2 = (uint32_t)bp.parent_a2, a3 = (uint32_t)bp.parent_a3;
            for (int j = 0; e && j < PW_BITS; j++, e >>= 1) if (e & 1) { a1 = a1 * pw[j] % 30269u; a2 = a2 * pw[PW_BITS + j] % 30307u; a3 = a3 * pw[2 * PW_BITS + j] % 30323u; }
            bool fast = e == 0 && a.fused;
            FastOut fo; fo.pos = 0; fo.skip = 0; fo.ll = 0; fo.name = 0;
            Rng g; int64_t ts0 = 0, ts1 = 0, ts2 = 0;
            uint64_t bi = (I - 1) % bp.n_blobs;
            const uint8_t* blob = a.data + a.off[bi]; uint32_t blen = (uint32_t)(a.off[bi + 1] - a.off[bi]);
            uint64_t s0 = 0, cap = 0;
            if (fast) { s0 = a.slot_off[k]; cap = a.slot_off[k + 1] - s0; }
            uint8_t* num_start = nullptr;
            if (fast) {
                Rng par; par.mode = 0; par.a1 = (int32_t)a1; par.a2 = (int32_t)a2; par.a3 = (int32_t)a3; par.draws = 0; par.key = 0; par.ctr_hi = 0;
                ts0 = (int64_t)par.erand(99999); ts1 = (int64_t)par.erand(99999); ts2 = (int64_t)par.erand(99999);
                g.mode = bp.rng_mode; g.key = bp.philox_key; g.ctr_hi = I; g.seed(ts0, ts1, ts2);
                fast = fast_decide_lane(bp, g, q, fs, blen, blob, a.out + s0, cap, fo, &num_start);
            }
            uint64_t olen = (uint64_t)blen - fo.skip + fo.ll;
            if (fast && (olen > cap || olen > bp.max_case_out)) fast = false;
            if (!fast) slow_push(slow, k);
            else {
                uint64_t oo = s0;
                if (num_start) oo = (uint64_t)(num_start - a.out);      // sed_num: the workers already moved the bytes
                else {
                    Job j; j.a = (uint64_t)(uintptr_t)(a.out + s0); j.b = (uint64_t)(uintptr_t)blob; j.len = blen; j.kind = JOB_EDIT | (fo.ll << 8) | (fo.skip << 16); j.res = fo.pos; j.pend = 0;
                    for (uint32_t i = 0; i < fo.ll; i++) a.out[s0 + fo.pos + i] = fo.lit[i];
                    // the fronts run far ahead of the workers: keep the ring shallow, so that the scan and copy jobs of the
                    // general deciders (whose cases wait for them) are never queued behind hundreds of whole-case edits
                    while ((int32_t)(*(volatile unsigned int*)&q->tail - *(volatile unsigned int*)&q->head) > fa.front_depth) __nanosleep(500);
                    uint32_t pos = atomicAdd(&q->tail, 1u);
                    jobq_put(q, pos, j);
                }
                a.out_off[k] = oo; a.out_len[k] = olen;
                if (a.ar.case_status) a.ar.case_status[k] = 0;
                if (a.meta) {
                    MetaDev m; m.pattern = P_OD; m.generator = 0; m.n_used = 1; m.n_failed = 0;
                    for (int i = 0; i < 16; i++) m.used[i] = -1;
                    m.used[0] = (int32_t)fo.name;
                    m.draws = g.draws; m.status = 0; m.pad = 0; m.thread_seed[0] = ts0; m.thread_seed[1] = ts1; m.thread_seed[2] = ts2;
                    a.meta[k] = m;
                }
            }
        }
        __syncwarp();
    }
    slow_front_done(slow);
}

// The engine's one kernel: a persistent CTA per SM with three kinds of warps --
//   FRONT warps (0 .. fronts)      : take 32 cases at a time from a global counter, one lane per case; byte-mutator cases
//                                    under `od` are decided right there (an EDIT job each), the rest go to the slow ring;
//   DECIDER warps (next `deciders`): the general per-case program (decide_one_case), one warp per case from the slow ring
//                                    (or straight from the global counter when there are no fronts);
//   WORKER warps (the rest)        : execute the job queue -- streaming copies, block scans, EDIT jobs (eb_jobs.cuh).
// fronts = 0 and deciders = all warps gives the round-1 arrangement (every warp does its own byte work inline) for A/B runs.
// FULL / LIGHT: see mut_is_light().
constexpr int MAX_FRONTS = 4;
static inline __host__ __device__ size_t case_smem_layout(int deciders, size_t* pw_off, size_t* slow_off, size_t* ws_off, size_t* fs_off = nullptr, int tma_workers = 0, size_t* tma_off = nullptr) {
    size_t o = sizeof(JobQ);
    if (pw_off) *pw_off = o;
    o += 3 * PW_BITS * 4; o = (o + 15) & ~(size_t)15;
    if (slow_off) *slow_off = o;
    o += sizeof(SlowRing); o = (o + 15) & ~(size_t)15;
    if (fs_off) *fs_off = o;
    o += sizeof(FrontState) * MAX_FRONTS; o = (o + 15) & ~(size_t)15;
    if (ws_off) *ws_off = o;
    o += sizeof(WarpState) * (size_t)deciders; o = (o + 127) & ~(size_t)127;
    if (tma_off) *tma_off = o;
    return o + sizeof(TmaStage) * (size_t)tma_workers;
}
template <bool FULL>
__global__ void __launch_bounds__(CASE_THREADS, 1)
eb_case_kernel(const uint8_t* __restrict__ data, const uint64_t* __restrict__ off, BatchParams bp, Arenas ar,
               CaseOut* __restrict__ cases, uint64_t* __restrict__ out_len, uint64_t* __restrict__ out_sz16, MetaDev* __restrict__ meta, FusedArgs fa) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int nwarps = (int)(blockDim.x >> 5), warp = (int)(threadIdx.x >> 5);
    int fronts = fa.fronts, deciders = fa.deciders;
    if (deciders > nwarps) deciders = nwarps;
    if (fronts + deciders >= nwarps) fronts = 0;       // fronts only make sense with workers behind them
    if (fronts > MAX_FRONTS) fronts = MAX_FRONTS;
    size_t pw_off, slow_off, ws_off, fs_off, tma_off;
    case_smem_layout(deciders, &pw_off, &slow_off, &ws_off, &fs_off, fa.tma_workers, &tma_off);
    JobQ* q = reinterpret_cast<JobQ*>(smem_raw);
    uint32_t* pw = reinterpret_cast<uint32_t*>(smem_raw + pw_off);                 // [3][PW_BITS]: a^(3 * 2^j) mod p
    SlowRing* slow = reinterpret_cast<SlowRing*>(smem_raw + slow_off);
    WarpState* wsbase = reinterpret_cast<WarpState*>(smem_raw + ws_off);
    jobq_init(q, deciders);
    slow_init(slow, fronts);
    if (threadIdx.x < 3 * PW_BITS) {
        int comp = threadIdx.x / PW_BITS, j = threadIdx.x % PW_BITS;
        uint32_t m = comp == 0 ? 30269u : comp == 1 ? 30307u : 30323u, a0 = comp == 0 ? AS_M1 : comp == 1 ? AS_M2 : AS_M3;
        uint32_t v = (a0 * a0 % m) * a0 % m;
        for (int i = 0; i < j; i++) v = v * v % m;
        pw[threadIdx.x] = v;
    }
    __syncthreads();
    DecideArgs a; a.data = data; a.off = off; a.ar = ar; a.cases = cases; a.out_len = out_len; a.out_sz16 = out_sz16; a.meta = meta;
    a.fused = fa.fused; a.out = fa.out; a.out_capacity = fa.out_capacity; a.slot_off = fa.slot_off; a.out_off = fa.out_off;
    a.ovf_base = fa.ovf_base; a.ovf_used = fa.ovf_used; a.data_bytes = fa.data_bytes;
    if (warp < fronts) { front_loop(bp, a, fa, q, slow, pw, reinterpret_cast<FrontState*>(smem_raw + fs_off) + warp); return; }
    if (warp >= fronts + deciders) {
        int wi = warp - fronts - deciders;       // the first tma_workers workers stage through shared memory with bulk-async copies
        worker_loop(q, wi < fa.tma_workers ? reinterpret_cast<TmaStage*>(smem_raw + tma_off) + wi : nullptr);
        return;
    }
    const bool have_workers = fronts + deciders < nwarps;
    JobQ* qq = have_workers ? q : nullptr;
    WarpState* ws = wsbase + (warp - fronts);
    const uint64_t total = fa.case_list ? fa.n_list : bp.n_cases;
    const uint32_t temp_slot = blockIdx.x * (uint32_t)deciders + (uint32_t)(warp - fronts);
    for (;;) {
        uint64_t k;
        if (fronts) { if (!slow_pop(slow, k)) break; }
        else {
            unsigned long long idx = 0;
            if (lane_id() == 0) idx = atomicAdd(fa.case_counter, 1ull);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            if (idx >= total) break;
            k = fa.case_list ? (uint64_t)fa.case_list[idx] : (uint64_t)idx;
        }
        // parent stream at case k: x0 * (a^3)^(first_case - 1 + k) mod p per AS183 component, from the power table
        uint64_t e = bp.first_case - 1 + k;
        uint32_t a1 = (uint32_t)bp.parent_a1, a2 = (uint32_t)bp.parent_a2, a3 = (uint32_t)bp.parent_a3;
        for (int j = 0; e && j < PW_BITS; j++, e >>= 1) if (e & 1) { a1 = a1 * pw[j] % 30269u; a2 = a2 * pw[PW_BITS + j] % 30307u; a3 = a3 * pw[2 * PW_BITS + j] % 30323u; }
        if (e) {   // beyond 2^48 cases: finish with the generic jump
            Rng t; t.mode = 0; t.a1 = (int32_t)a1; t.a2 = (int32_t)a2; t.a3 = (int32_t)a3; t.draws = 0; t.key = 0; t.ctr_hi = 0; t.jump(3 * (e << PW_BITS));
            a1 = (uint32_t)t.a1; a2 = (uint32_t)t.a2; a3 = (uint32_t)t.a3;
        }
        decide_one_case<FULL>(ws, bp, a, k, (int32_t)a1, (int32_t)a2, (int32_t)a3, qq, temp_slot);
    }
    if (qq) jobq_decider_done(q);
}

}  // namespace eb
