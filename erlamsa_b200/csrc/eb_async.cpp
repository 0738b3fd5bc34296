// erlamsa_b200 -- asynchronous submit / collect pair of the C ABI (include/erlamsa_b200.h).
//
// eb200_fuzz_batch_device is synchronous: it has to read the batch's arena counters back (flagged cases are re-run in a
// follow-up launch), so consecutive device batches are separated by stream synchronisations and the GPU idles for the
// host's share of every step. The pair below removes that gap without a second code path: a context owns a few LANES, each
// lane = its own engine context (arenas, counters, events) + its own non-blocking stream + one host thread that runs the
// synchronous entry point. While one lane's thread sits in a synchronise, the other lane's kernel is already running.
// The persistent case kernel fills every SM, so lanes do not share the GPU -- they only hide each other's host gaps.
//
// Host code only: nothing here launches a kernel of its own.
#include <cuda_runtime.h>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include "../../include/erlamsa_b200.h"

extern "C" {
void** eb200_ctx_async_slot(eb200_ctx* ctx);
int eb200_ctx_device(eb200_ctx* ctx);
void eb200_ctx_set_error(eb200_ctx* ctx, const char* msg);
void eb200_async_teardown(void* state);
}

namespace {

enum { MAX_LANES = 4 };

struct Job {
    eb200_opts opts;
    const uint8_t* d_data; const uint64_t* d_off; uint64_t n_blobs, data_bytes, n_cases;
    uint8_t* d_out; uint64_t out_capacity; uint64_t* d_out_off; uint64_t* d_out_len; eb200_meta* d_meta;
    eb200_stats stats;
    int rc = EB200_OK;
    bool done = false;
    std::string err;
    struct Lane* lane = nullptr;
};

struct Lane {
    eb200_ctx* sub = nullptr;
    cudaStream_t st = nullptr;
    std::thread th;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::deque<Job*> q;           // submitted, not yet run
    bool stop = false;

    void run() {
        for (;;) {
            Job* j = nullptr;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_work.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;          // stop requested and nothing left to run
                j = q.front(); q.pop_front();
            }
            int rc = eb200_fuzz_batch_device(sub, &j->opts, j->d_data, j->d_off, j->n_blobs, j->data_bytes, j->n_cases, j->d_out, j->out_capacity,
                                             j->d_out_off, j->d_out_len, j->d_meta, (void*)st, &j->stats);
            std::string err = rc == EB200_ERR_CUDA ? std::string(eb200_last_cuda_error(sub)) : std::string();
            {
                std::lock_guard<std::mutex> lk(m);
                j->rc = rc; j->err.swap(err); j->done = true;
            }
            cv_done.notify_all();
        }
    }
};

struct AsyncState {
    int device = 0;
    int n_lanes = 0;
    Lane lanes[MAX_LANES];
    std::mutex m;                  // guards `outstanding` and `next`
    std::deque<Job*> outstanding;  // every ticket handed out and not yet collected (freed at teardown if never collected)
    uint64_t next = 0;
};

int make_state(eb200_ctx* ctx, AsyncState** out) {
    int lanes = 2;
    if (const char* v = getenv("EB200_ASYNC_LANES")) { int k = atoi(v); if (k >= 1 && k <= MAX_LANES) lanes = k; }
    AsyncState* s = new AsyncState();
    s->device = eb200_ctx_device(ctx);
    if (cudaSetDevice(s->device) != cudaSuccess) { delete s; return EB200_ERR_CUDA; }
    for (int i = 0; i < lanes; i++) {
        Lane& l = s->lanes[i];
        int rc = eb200_init(s->device, &l.sub);
        if (rc == EB200_OK && cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking) != cudaSuccess) { eb200_shutdown(l.sub); l.sub = nullptr; rc = EB200_ERR_CUDA; }
        if (rc != EB200_OK) { s->n_lanes = i; eb200_async_teardown(s); return rc; }
        s->n_lanes = i + 1;
        l.th = std::thread([&l] { l.run(); });
    }
    *out = s;
    return EB200_OK;
}

}  // namespace

extern "C" {

void eb200_async_teardown(void* state) {
    AsyncState* s = (AsyncState*)state;
    if (!s) return;
    for (int i = 0; i < s->n_lanes; i++) {
        Lane& l = s->lanes[i];
        { std::lock_guard<std::mutex> lk(l.m); l.stop = true; }
        l.cv_work.notify_all();
        if (l.th.joinable()) l.th.join();          // the thread drains its queue first: no batch is dropped half-way
        cudaSetDevice(s->device);
        if (l.st) { cudaStreamSynchronize(l.st); cudaStreamDestroy(l.st); }
        if (l.sub) eb200_shutdown(l.sub);
    }
    for (Job* j : s->outstanding) delete j;
    delete s;
}

int eb200_submit_device(eb200_ctx* ctx, const eb200_opts* opts, const uint8_t* d_data, const uint64_t* d_off, uint64_t n_blobs, uint64_t data_bytes,
                        uint64_t n_cases, uint8_t* d_out, uint64_t out_capacity, uint64_t* d_out_off, uint64_t* d_out_len, eb200_meta* d_meta,
                        eb200_ticket** ticket) {
    if (!ctx || !opts || !ticket) return EB200_ERR_ARG;
    // the argument checks of the synchronous call, so that a bad batch is refused here and not at collect time
    if (!d_data || !d_off || !d_out || !d_out_off || !d_out_len || n_blobs == 0) return EB200_ERR_ARG;
    if (((uintptr_t)d_data & 15u) || ((uintptr_t)d_out & 15u)) return EB200_ERR_ARG;
    *ticket = nullptr;
    void** slot = eb200_ctx_async_slot(ctx);
    static std::mutex create_m;
    std::unique_lock<std::mutex> create_lk(create_m);
    if (!*slot) {
        AsyncState* s = nullptr;
        int rc = make_state(ctx, &s);
        if (rc != EB200_OK) { eb200_ctx_set_error(ctx, "eb200_submit_device: could not create the lanes"); return rc; }
        *slot = s;
    }
    AsyncState* s = (AsyncState*)*slot;
    create_lk.unlock();
    Job* j = new Job();
    j->opts = *opts;                                   // options are copied: the caller's struct may change after submit
    j->d_data = d_data; j->d_off = d_off; j->n_blobs = n_blobs; j->data_bytes = data_bytes; j->n_cases = n_cases;
    j->d_out = d_out; j->out_capacity = out_capacity; j->d_out_off = d_out_off; j->d_out_len = d_out_len; j->d_meta = d_meta;
    memset(&j->stats, 0, sizeof(j->stats));
    Lane* l;
    {
        std::lock_guard<std::mutex> lk(s->m);
        l = &s->lanes[s->next++ % (uint64_t)s->n_lanes];   // round robin: batch k and k+1 never wait for the same host thread
        s->outstanding.push_back(j);
    }
    j->lane = l;
    { std::lock_guard<std::mutex> lk(l->m); l->q.push_back(j); }
    l->cv_work.notify_one();
    *ticket = (eb200_ticket*)j;
    return EB200_OK;
}

int eb200_collect(eb200_ctx* ctx, eb200_ticket* ticket, eb200_stats* stats) {
    if (!ctx || !ticket) return EB200_ERR_ARG;
    void** slot = eb200_ctx_async_slot(ctx);
    AsyncState* s = slot ? (AsyncState*)*slot : nullptr;
    if (!s) return EB200_ERR_ARG;
    Job* j = (Job*)ticket;
    {   // only tickets of this context that were not collected before
        std::lock_guard<std::mutex> lk(s->m);
        bool found = false;
        for (auto it = s->outstanding.begin(); it != s->outstanding.end(); ++it) if (*it == j) { s->outstanding.erase(it); found = true; break; }
        if (!found) return EB200_ERR_ARG;
    }
    Lane* l = j->lane;
    {
        std::unique_lock<std::mutex> lk(l->m);
        l->cv_done.wait(lk, [&] { return j->done; });
    }
    int rc = j->rc;
    if (stats) *stats = j->stats;
    if (rc == EB200_ERR_CUDA) eb200_ctx_set_error(ctx, j->err.c_str());
    delete j;
    return rc;
}

int eb200_async_lanes(eb200_ctx* ctx) {
    void** slot = eb200_ctx_async_slot(ctx);
    AsyncState* s = slot ? (AsyncState*)*slot : nullptr;
    return s ? s->n_lanes : 0;
}

}  // extern "C"
