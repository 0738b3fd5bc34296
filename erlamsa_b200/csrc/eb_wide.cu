// erlamsa_b200 -- the general per-case program compiled a second time for 512 threads per CTA (128 registers per thread).
//
// The general program (eb_case_kernel<FULL>) is ~190 KB of scalar code whose working set does not fit 64 registers: at 1024
// threads per CTA it spills, and with 24 warps' stacks next to each other a spill is an L2 round trip, not an L1 hit. When a
// batch has only a few cases per warp the kernel's time is the slowest cases' latency, not throughput -- then half the warps
// with twice the registers finish sooner. Same sources, same results; only the register allocation differs. The engine picks
// per launch (eb_engine.cu plan_launch; EB200_WIDE=0/1 forces it).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#define EB_CASE_THREADS 512
#define EB_WIDE_RECONVERGE 1        // EB_RECONVERGE() = __syncwarp() in this build only (eb_common.cuh)
#define eb ebw                      // every device symbol of this translation unit lives in its own namespace
#include "../../include/erlamsa_b200.h"
#include "eb_fast.cuh"

extern "C" __attribute__((visibility("hidden"))) int eb200_wide_init(void) {
    return cudaFuncSetAttribute(ebw::eb_case_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448) == cudaSuccess ? 0 : 1;
}
// bp / ar / fa point to the engine's BatchParams / Arenas / FusedArgs (identical layouts: same headers)
extern "C" __attribute__((visibility("hidden"))) void eb200_wide_launch(int grid, int threads, size_t smem, cudaStream_t st, const uint8_t* d_data, const uint64_t* d_off,
                                                                         const void* bp, const void* ar, void* cases, uint64_t* out_len, uint64_t* sz16, void* meta, const void* fa) {
    ebw::eb_case_kernel<true><<<grid, threads, smem, st>>>(d_data, d_off, *static_cast<const ebw::BatchParams*>(bp), *static_cast<const ebw::Arenas*>(ar),
                                                            static_cast<ebw::CaseOut*>(cases), out_len, sz16, static_cast<ebw::MetaDev*>(meta), *static_cast<const ebw::FusedArgs*>(fa));
}
