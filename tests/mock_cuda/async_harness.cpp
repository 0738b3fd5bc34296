// Drives erlamsa_b200/csrc/eb_async.cpp against a MOCK engine (no GPU): checks the ticket protocol, that lanes overlap,
// that results and stats come back for the right ticket in any collect order, error propagation, and teardown with batches
// still queued. Built and run by tests/test_async_harness.py (also under -fsanitize=thread).
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "erlamsa_b200.h"

struct eb200_ctx { int device; void* async_state; std::string err; int id; };
static std::atomic<int> g_ctx_made{0}, g_ctx_freed{0}, g_running{0}, g_max_running{0}, g_streams{0}, g_calls{0};
static std::mutex g_ids_m; static std::vector<int> g_lane_ids;

extern "C" {
void eb200_async_teardown(void* state);
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned) { *st = (cudaStream_t)(intptr_t)(++g_streams); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { --g_streams; return cudaSuccess; }

int eb200_init(int device, eb200_ctx** out) { eb200_ctx* c = new eb200_ctx(); c->device = device; c->async_state = nullptr; c->id = ++g_ctx_made; *out = c; return EB200_OK; }
void eb200_shutdown(eb200_ctx* c) { if (!c) return; if (c->async_state) { eb200_async_teardown(c->async_state); c->async_state = nullptr; } ++g_ctx_freed; delete c; }
const char* eb200_last_cuda_error(eb200_ctx* c) { return c ? c->err.c_str() : ""; }
void** eb200_ctx_async_slot(eb200_ctx* c) { return c ? &c->async_state : nullptr; }
int eb200_ctx_device(eb200_ctx* c) { return c ? c->device : -1; }
void eb200_ctx_set_error(eb200_ctx* c, const char* m) { if (c && m) c->err = m; }

// the mock "kernel": out[k] = data[k] ^ seed[0] for n_cases bytes, takes ~n_cases/1000 ms; first_case == 666 fails
int eb200_fuzz_batch_device(eb200_ctx* ctx, const eb200_opts* o, const uint8_t* d_data, const uint64_t*, uint64_t, uint64_t, uint64_t n_cases,
                            uint8_t* d_out, uint64_t out_capacity, uint64_t* d_out_off, uint64_t* d_out_len, eb200_meta*, void* stream, eb200_stats* st) {
    ++g_calls;
    int r = ++g_running; int m = g_max_running.load(); while (r > m && !g_max_running.compare_exchange_weak(m, r)) {}
    { std::lock_guard<std::mutex> lk(g_ids_m); g_lane_ids.push_back(ctx->id); }
    if (!stream) { --g_running; return EB200_ERR_ARG; }      // lanes must bring their own stream
    std::this_thread::sleep_for(std::chrono::microseconds(n_cases));
    int rc = EB200_OK;
    if (o->first_case == 666) { ctx->err = "mock launch failure"; rc = EB200_ERR_CUDA; }
    else if (n_cases > out_capacity) rc = EB200_ERR_NOMEM;
    else {
        for (uint64_t k = 0; k < n_cases; k++) { d_out[k] = d_data[k] ^ (uint8_t)o->seed[0]; d_out_off[k] = k; d_out_len[k] = 1; }
        memset(st, 0, sizeof(*st)); st->n_cases = n_cases; st->kernels_launched = 5; st->bytes_out = n_cases;
    }
    --g_running;
    return rc;
}
}

#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main() {
    eb200_ctx* ctx = nullptr;
    CHECK(eb200_init(0, &ctx) == EB200_OK);
    CHECK(eb200_async_lanes(ctx) == 0);
    const int NB = 12; const uint64_t N = 20000;
    std::vector<std::vector<uint8_t>> in(NB, std::vector<uint8_t>(N)), out(NB, std::vector<uint8_t>(N));
    std::vector<std::vector<uint64_t>> off(NB, std::vector<uint64_t>(N + 1)), len(NB, std::vector<uint64_t>(N));
    std::vector<uint64_t> doff(2, 0);
    // 16-byte aligned "device" buffers
    std::vector<uint8_t*> pin(NB), pout(NB);
    for (int b = 0; b < NB; b++) { pin[b] = (uint8_t*)aligned_alloc(16, N); pout[b] = (uint8_t*)aligned_alloc(16, N); for (uint64_t k = 0; k < N; k++) pin[b][k] = (uint8_t)(k * 7 + b); }
    eb200_opts o; memset(&o, 0, sizeof(o));
    std::vector<eb200_ticket*> t(NB, nullptr);
    // argument checks happen at submit time
    eb200_ticket* bad = nullptr;
    CHECK(eb200_submit_device(ctx, &o, nullptr, doff.data(), 1, N, N, pout[0], N, off[0].data(), len[0].data(), nullptr, &bad) == EB200_ERR_ARG);
    CHECK(eb200_submit_device(ctx, &o, pin[0] + 1, doff.data(), 1, N, N, pout[0], N, off[0].data(), len[0].data(), nullptr, &bad) == EB200_ERR_ARG);
    CHECK(eb200_async_lanes(ctx) == 0);
    auto t0 = std::chrono::steady_clock::now();
    for (int b = 0; b < NB; b++) {
        o.seed[0] = b + 1; o.first_case = 1 + b;
        CHECK(eb200_submit_device(ctx, &o, pin[b], doff.data(), 1, N, N, pout[b], N, off[b].data(), len[b].data(), nullptr, &t[b]) == EB200_OK);
        o.seed[0] = 99;            // options were copied at submit: changing the struct afterwards must not matter
    }
    double submit_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    CHECK(eb200_async_lanes(ctx) == 2);
    CHECK(submit_ms < 0.5 * NB * N / 1000.0);          // submit returns at once (the 12 batches take ~240 ms of mock work)
    // collect in a scrambled order
    int order[NB] = {5, 0, 11, 3, 1, 2, 10, 4, 9, 6, 8, 7};
    for (int i = 0; i < NB; i++) {
        int b = order[i]; eb200_stats st; memset(&st, 0xff, sizeof(st));
        CHECK(eb200_collect(ctx, t[b], &st) == EB200_OK);
        CHECK(st.n_cases == N && st.kernels_launched == 5 && st.bytes_out == N);
        for (uint64_t k = 0; k < N; k++) CHECK(pout[b][k] == (uint8_t)(pin[b][k] ^ (uint8_t)(b + 1)));
        CHECK(eb200_collect(ctx, t[b], &st) == EB200_ERR_ARG);       // a ticket is collected once
    }
    double total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    CHECK(g_max_running.load() == 2);                                 // the two lanes really overlapped ...
    CHECK(total_ms < 0.8 * NB * N / 1000.0);                          // ... so 12 batches took about half of 12 x 20 ms
    { std::lock_guard<std::mutex> lk(g_ids_m); int a = 0, c = 0; for (int id : g_lane_ids) { if (id == 2) a++; else if (id == 3) c++; } CHECK(a == NB / 2 && c == NB / 2); }   // round robin over the lanes' own contexts
    // a foreign ticket
    CHECK(eb200_collect(ctx, (eb200_ticket*)&o, nullptr) == EB200_ERR_ARG);
    // errors come back at collect time, with the lane's error text on the caller's context
    o.first_case = 666; eb200_ticket* te = nullptr;
    CHECK(eb200_submit_device(ctx, &o, pin[0], doff.data(), 1, N, 100, pout[0], N, off[0].data(), len[0].data(), nullptr, &te) == EB200_OK);
    CHECK(eb200_collect(ctx, te, nullptr) == EB200_ERR_CUDA);
    CHECK(std::string(eb200_last_cuda_error(ctx)) == "mock launch failure");
    o.first_case = 1;
    CHECK(eb200_submit_device(ctx, &o, pin[0], doff.data(), 1, N, N, pout[0], 10, off[0].data(), len[0].data(), nullptr, &te) == EB200_OK);
    CHECK(eb200_collect(ctx, te, nullptr) == EB200_ERR_NOMEM);
    // concurrent submitters / collectors on one context
    {
        std::atomic<int> fails{0};
        auto worker = [&](int b) {
            eb200_opts oo; memset(&oo, 0, sizeof(oo)); oo.seed[0] = 40 + b; oo.first_case = 1;
            for (int rep = 0; rep < 4; rep++) {
                eb200_ticket* tk = nullptr; eb200_stats st;
                if (eb200_submit_device(ctx, &oo, pin[b], doff.data(), 1, N, 2000, pout[b], N, off[b].data(), len[b].data(), nullptr, &tk) != EB200_OK) { fails++; continue; }
                if (eb200_collect(ctx, tk, &st) != EB200_OK || st.n_cases != 2000) fails++;
                for (uint64_t k = 0; k < 2000; k++) if (pout[b][k] != (uint8_t)(pin[b][k] ^ (uint8_t)(40 + b))) { fails++; break; }
            }
        };
        std::vector<std::thread> th; for (int b = 0; b < 6; b++) th.emplace_back(worker, b);
        for (auto& x : th) x.join();
        CHECK(fails.load() == 0);
    }
    // shutdown with batches still queued and never collected: they are run to the end, nothing leaks, nothing hangs
    int calls_before = g_calls.load();
    for (int b = 0; b < 5; b++) { o.seed[0] = 7; CHECK(eb200_submit_device(ctx, &o, pin[b], doff.data(), 1, N, 3000, pout[b], N, off[b].data(), len[b].data(), nullptr, &t[b]) == EB200_OK); }
    eb200_shutdown(ctx);
    CHECK(g_calls.load() == calls_before + 5);
    CHECK(g_ctx_made.load() == 3 && g_ctx_freed.load() == 3 && g_streams.load() == 0);
    for (int b = 0; b < 5; b++) for (uint64_t k = 0; k < 3000; k++) CHECK(pout[b][k] == (uint8_t)(pin[b][k] ^ 7));
    for (int b = 0; b < NB; b++) { free(pin[b]); free(pout[b]); }
    printf("OK lanes=2 max_running=%d submit_ms=%.2f total_ms=%.1f\n", g_max_running.load(), submit_ms, total_ms);
    return 0;
}
