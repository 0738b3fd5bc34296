// A stand-in liberlamsa_b200.so for the CPU test of the PYTHON binding of the asynchronous pair (tests/test_async_harness.py): the real
// erlamsa_b200/csrc/eb_async.cpp over a mock engine whose "device" memory is host memory. Loaded through EB200_LIB in a subprocess;
// only what erlamsa_b200/_native.py binds and erlamsa_b200/options.py calls exists here. Test infrastructure, never shipped.
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>
#include <string>
#include "erlamsa_b200.h"

struct eb200_ctx { int device; void* async_state; std::string err; };

static const char* const kMut[EB200_N_MUTATORS] = {"sgm", "js", "uw", "ui", "ab", "ad", "tr2", "td", "num", "ts1", "tr", "ts2", "bd", "bei", "bed", "bf", "bi", "ber", "br", "sp", "sr",
    "sd", "snand", "srnd", "ld", "lds", "lr2", "lri", "lr", "ls", "lp", "lis", "lrs", "ft", "fn", "fo", "len", "b64", "uri", "zip", "nil"};
static const char* const kPat[EB200_N_PATTERNS] = {"od", "nd", "bu", "sk", "sz", "cs", "ar", "cp", "co", "nu"};

extern "C" {
void eb200_async_teardown(void* state);
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned) { *st = (cudaStream_t)(intptr_t)1; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }

void eb200_default_opts(eb200_opts* o) { memset(o, 0, sizeof(*o)); o->seed[0] = 1; o->seed[1] = 2; o->seed[2] = 3; o->blockscale = 1.0; o->first_case = 1;
    for (int i = 0; i < EB200_N_MUTATORS; i++) o->muta_pri[i] = 1;
    for (int i = 0; i < EB200_N_PATTERNS; i++) o->pat_pri[i] = 1;
    o->gen_direct_pri = 500; o->gen_random_pri = 1; o->gen_file_pri = o->gen_stdin_pri = o->gen_jump_pri = -1; strcpy(o->ssrf_host, "localhost"); o->ssrf_port = 51234; }
int eb200_init(int device, eb200_ctx** out) { eb200_ctx* c = new eb200_ctx(); c->device = device; c->async_state = nullptr; *out = c; return EB200_OK; }
void eb200_shutdown(eb200_ctx* c) { if (!c) return; if (c->async_state) { eb200_async_teardown(c->async_state); c->async_state = nullptr; } delete c; }
const char* eb200_last_cuda_error(eb200_ctx* c) { return c ? c->err.c_str() : ""; }
void** eb200_ctx_async_slot(eb200_ctx* c) { return c ? &c->async_state : nullptr; }
int eb200_ctx_device(eb200_ctx* c) { return c ? c->device : -1; }
void eb200_ctx_set_error(eb200_ctx* c, const char* m) { if (c && m) c->err = m; }

// the mock batch: case k's output = blob k with every byte XOR seed[0] (low byte), placed at 16 * k + (sum of earlier lengths rounded up to 16)
int eb200_fuzz_batch_device(eb200_ctx*, const eb200_opts* o, const uint8_t* d_data, const uint64_t* d_off, uint64_t n_blobs, uint64_t, uint64_t n_cases,
                            uint8_t* d_out, uint64_t out_capacity, uint64_t* d_out_off, uint64_t* d_out_len, eb200_meta* d_meta, void*, eb200_stats* st) {
    uint64_t pos = 0, first = o->first_case ? o->first_case : 1;
    for (uint64_t k = 0; k < n_cases; k++) {
        uint64_t b = (first - 1 + k) % n_blobs, len = d_off[b + 1] - d_off[b];
        if (pos + len > out_capacity) return EB200_ERR_NOMEM;
        for (uint64_t i = 0; i < len; i++) d_out[pos + i] = d_data[d_off[b] + i] ^ (uint8_t)o->seed[0];
        d_out_off[k] = pos; d_out_len[k] = len;
        if (d_meta) { memset(&d_meta[k], 0, sizeof(eb200_meta)); d_meta[k].draws = first + k; d_meta[k].pattern = (int32_t)(k % 10); }
        pos += (len + 15) & ~15ull;
    }
    d_out_off[n_cases] = pos;
    if (st) { memset(st, 0, sizeof(*st)); st->n_cases = n_cases; st->kernels_launched = 5; st->bytes_out = pos; }
    return EB200_OK;
}
// host-buffer batch of the mock: case I = first_case + k gives blob (I-1) mod n_blobs with every byte XOR (I + case_stream_seed[0] + 7 * case_stream_first)
int eb200_fuzz_batch(eb200_ctx*, const eb200_opts* o, const uint8_t* data, const uint64_t* off, uint64_t n_blobs, uint64_t n_cases, uint8_t** out_data, uint64_t* out_off,
                     uint64_t* out_len, eb200_meta* meta, eb200_stats* st) {
    uint64_t first = o->first_case ? o->first_case : 1, total = 0;
    for (uint64_t k = 0; k < n_cases; k++) { uint64_t b = (first - 1 + k) % n_blobs; total += off[b + 1] - off[b]; }
    uint8_t* out = (uint8_t*)malloc(total ? total : 1); uint64_t pos = 0;
    for (uint64_t k = 0; k < n_cases; k++) {
        uint64_t i = first + k, b = (i - 1) % n_blobs, len = off[b + 1] - off[b];
        uint8_t x = (uint8_t)(i + (uint64_t)o->case_stream_seed[0] + 7 * o->case_stream_first);
        for (uint64_t j = 0; j < len; j++) out[pos + j] = data[off[b] + j] ^ x;
        out_off[k] = pos; out_len[k] = len; pos += len;
        if (meta) memset(&meta[k], 0, sizeof(eb200_meta));
    }
    out_off[n_cases] = pos; *out_data = out;
    if (st) { memset(st, 0, sizeof(*st)); st->n_cases = n_cases; }
    return EB200_OK;
}
int eb200_fuzz_batch_into(eb200_ctx*, const eb200_opts*, const uint8_t*, const uint64_t*, uint64_t, uint64_t, uint8_t*, uint64_t, uint64_t*, uint64_t*, eb200_meta*, eb200_stats*) { return EB200_ERR_NO_DEVICE; }
int eb200_sample_donors(eb200_ctx*, const uint8_t*, const uint64_t*, uint64_t, uint64_t, uint32_t, uint8_t*, uint32_t*, void*) { return EB200_ERR_NO_DEVICE; }
void* eb200_host_alloc(eb200_ctx*, uint64_t) { return nullptr; }
void eb200_host_free(eb200_ctx*, void*) {}
int eb200_numa_node(eb200_ctx*) { return -1; }
void eb200_free(void* p) { free(p); }
uint64_t eb200_debug_case_times(eb200_ctx*, uint32_t*, uint64_t) { return 0; }
int eb200_debug_mutator_times(eb200_ctx*, uint64_t*) { return 0; }
int eb200_debug_parent_draws(const eb200_opts*, uint64_t, uint64_t, int64_t*) { return EB200_ERR_NO_DEVICE; }
const char* eb200_mutator_code(int i) { return (i >= 0 && i < EB200_N_MUTATORS) ? kMut[i] : nullptr; }
int eb200_mutator_default_pri(int) { return 1; }
int eb200_mutator_supported(int) { return 1; }
const char* eb200_pattern_code(int i) { return (i >= 0 && i < EB200_N_PATTERNS) ? kPat[i] : nullptr; }
int eb200_pattern_default_pri(int) { return 1; }
int eb200_pattern_supported(int) { return 1; }
const char* eb200_strerror(int code) { return code == EB200_OK ? "ok" : code == EB200_ERR_ARG ? "bad argument" : "mock error"; }
const char* eb200_version(void) { return "erlamsa_b200 MOCK (tests/mock_cuda/mock_lib.cpp)"; }
}
