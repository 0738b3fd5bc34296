/*
 * erlamsa_b200_nif.c -- the Erlang NIF shim over include/erlamsa_b200.h.
 *
 * NOT BUILT IN THIS REPOSITORY'S CI: the build image has no Erlang/OTP (no erl_nif.h). It is the binding a
 * maintainer adds to erlamsa so that erlamsa_main:fuzzer/1 (reference src/erlamsa_main.erl:124) can hand a whole
 * corpus to the GPU engine; the same C ABI is exercised in CI through Python ctypes (tests/).
 *
 *   erlamsa_b200:fuzz_batch_nif(Blobs :: [binary()], NCases, Seed :: {A,B,C}, MutaPri :: [integer()] (41),
 *                               PatPri :: [integer()] (10), FirstCase, BlockScale :: float())
 *       -> {ok, [binary()]} | {error, Reason :: atom()}
 *
 * Build (where OTP and the engine are installed):
 *   gcc -O2 -fPIC -shared -I$ERL_INCLUDE -I../include erlamsa_b200_nif.c -L../erlamsa_b200 -lerlamsa_b200 \
 *       -o priv/erlamsa_b200_nif.so
 */
#include <string.h>
#include <stdlib.h>
#include "erl_nif.h"
#include "erlamsa_b200.h"

static eb200_ctx* g_ctx = NULL;   /* one context per VM; the engine serialises batches on its stream */

static int load(ErlNifEnv* env, void** priv, ERL_NIF_TERM info) {
    (void)env; (void)priv; (void)info;
    return eb200_init(0, &g_ctx) == EB200_OK ? 0 : 1;   /* no GPU => the NIF refuses to load; erlamsa keeps its Erlang path */
}
static void unload(ErlNifEnv* env, void* priv) { (void)env; (void)priv; if (g_ctx) eb200_shutdown(g_ctx); }

static ERL_NIF_TERM err(ErlNifEnv* env, const char* why) {
    return enif_make_tuple2(env, enif_make_atom(env, "error"), enif_make_atom(env, why));
}
static int get_int_list(ErlNifEnv* env, ERL_NIF_TERM l, int32_t* out, unsigned n) {
    unsigned len; ERL_NIF_TERM h, t = l;
    if (!enif_get_list_length(env, l, &len) || len != n) return 0;
    for (unsigned i = 0; i < n; i++) { int v; if (!enif_get_list_cell(env, t, &h, &t) || !enif_get_int(env, h, &v)) return 0; out[i] = v; }
    return 1;
}

/* runs on a dirty IO scheduler: a batch takes milliseconds to seconds */
static ERL_NIF_TERM fuzz_batch_nif(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    (void)argc;
    unsigned n_blobs; ErlNifUInt64 n_cases, first_case; double blockscale;
    const ERL_NIF_TERM* seed; int seed_arity;
    eb200_opts o; eb200_default_opts(&o);
    if (!enif_get_list_length(env, argv[0], &n_blobs) || n_blobs == 0) return err(env, "badarg");
    if (!enif_get_uint64(env, argv[1], &n_cases)) return err(env, "badarg");
    if (!enif_get_tuple(env, argv[2], &seed_arity, &seed) || seed_arity != 3) return err(env, "badarg");
    for (int i = 0; i < 3; i++) { ErlNifSInt64 v; if (!enif_get_int64(env, seed[i], &v)) return err(env, "badarg"); o.seed[i] = v; }
    if (!get_int_list(env, argv[3], o.muta_pri, EB200_N_MUTATORS) || !get_int_list(env, argv[4], o.pat_pri, EB200_N_PATTERNS)) return err(env, "badarg");
    if (!enif_get_uint64(env, argv[5], &first_case) || !enif_get_double(env, argv[6], &blockscale)) return err(env, "badarg");
    o.first_case = first_case; o.blockscale = blockscale;

    /* pack the corpus: one contiguous buffer + offsets (the engine's input layout) */
    ErlNifBinary* bins = (ErlNifBinary*)enif_alloc(sizeof(ErlNifBinary) * n_blobs);
    uint64_t* off = (uint64_t*)enif_alloc(sizeof(uint64_t) * (n_blobs + 1));
    ERL_NIF_TERM h, t = argv[0]; uint64_t total = 0;
    for (unsigned i = 0; i < n_blobs; i++) {
        if (!enif_get_list_cell(env, t, &h, &t) || !enif_inspect_binary(env, h, &bins[i])) { enif_free(bins); enif_free(off); return err(env, "badarg"); }
        off[i] = total; total += bins[i].size;
    }
    off[n_blobs] = total;
    uint8_t* data = (uint8_t*)enif_alloc(total ? total : 1);
    for (unsigned i = 0; i < n_blobs; i++) memcpy(data + off[i], bins[i].data, bins[i].size);

    uint8_t* out = NULL;
    uint64_t* out_off = (uint64_t*)enif_alloc(sizeof(uint64_t) * (n_cases + 1));
    uint64_t* out_len = (uint64_t*)enif_alloc(sizeof(uint64_t) * (n_cases ? n_cases : 1));
    eb200_meta* meta = (eb200_meta*)enif_alloc(sizeof(eb200_meta) * (n_cases ? n_cases : 1));
    int rc = eb200_fuzz_batch(g_ctx, &o, data, off, n_blobs, n_cases, &out, out_off, out_len, meta, NULL);
    ERL_NIF_TERM res;
    if (rc != EB200_OK) {
        res = err(env, rc == EB200_ERR_UNSUPPORTED ? "unsupported" : rc == EB200_ERR_SCRATCH ? "scratch" : "engine");
    } else {
        /* {ok, [Binary | {flagged, CaseNo}]}: a case the engine flagged (unsupported path / capacity)
         * is handed back by number so that the Erlang side re-runs exactly that case on the reference path;
         * a case whose worker died (status 2) is an empty binary, as in the reference */
        ERL_NIF_TERM list = enif_make_list(env, 0);
        for (uint64_t k = n_cases; k-- > 0;) {   /* build back to front; empty outputs are kept, the caller filters (record_result/2) */
            ERL_NIF_TERM b;
            if (meta[k].status == EB200_CASE_UNSUPPORTED || meta[k].status == EB200_CASE_OVERFLOW) {
                b = enif_make_tuple2(env, enif_make_atom(env, "flagged"), enif_make_uint64(env, first_case + k));
            } else {
                unsigned char* p = enif_make_new_binary(env, out_len[k], &b);
                memcpy(p, out + out_off[k], out_len[k]);
            }
            list = enif_make_list_cell(env, b, list);
        }
        res = enif_make_tuple2(env, enif_make_atom(env, "ok"), list);
        eb200_free(out);
    }
    enif_free(bins); enif_free(off); enif_free(data); enif_free(out_off); enif_free(out_len); enif_free(meta);
    return res;
}

static ErlNifFunc nif_funcs[] = {{"fuzz_batch_nif", 7, fuzz_batch_nif, ERL_NIF_DIRTY_JOB_IO_BOUND}};
ERL_NIF_INIT(erlamsa_b200, nif_funcs, load, NULL, NULL, unload)
