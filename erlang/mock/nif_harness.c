/* Drives erlang/erlamsa_b200_nif.c (compiled against the mock erl_nif.h) the way the VM would: load/2, fuzz_batch_nif/10 with a
 * list of binaries, then walks the result. Prints one line per case: "<case> <len> <fnv1a64 of the bytes>" or "<case> flagged s r",
 * which tests/test_nif_harness.py compares with what the Python binding of the same C ABI returns.
 * usage: nif_harness <n_cases> <seed a> <b> <c> [bad]      (corpus: a fixed set of small blobs; `bad` exercises argument checks) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "erl_nif.h"
#include "erlamsa_b200.h"

ErlNifEntry* nif_init(void);

static uint64_t fnv(const unsigned char* p, size_t n) { uint64_t h = 1469598103934665603ull; for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; } return h; }

int main(int argc, char** argv) {
    int n_cases = argc > 1 ? atoi(argv[1]) : 16;
    long sa = argc > 4 ? atol(argv[2]) : 1, sb = argc > 4 ? atol(argv[3]) : 2, sc = argc > 4 ? atol(argv[4]) : 3;
    int bad = argc > 5 && !strcmp(argv[5], "bad");
    ErlNifEntry* entry = nif_init();
    ErlNifEnv* env = enif_mock_new_env();
    void* priv = NULL;
    if (entry->load(env, &priv, 0) != 0) { printf("load_failed\n"); return 3; }     /* no GPU: the NIF refuses to load */
    static const char* blobs[] = {"hello 100 world\n", "line one\nline two 42\nline three\n", "<a href=\"http://x/y\">t</a>", "{\"k\":[1,2,3],\"s\":\"v\"}", "AAAABBBBCCCCDDDD 7 8 9",
                                  "kittenslartibartfasterthaneelslartibartfastenyourseatbelts", "(x (Y x))", "A\n B\n C\n D\n"};
    int nb = (int)(sizeof(blobs) / sizeof(blobs[0]));
    ERL_NIF_TERM list = enif_make_list(env, 0);
    for (int i = nb - 1; i >= 0; i--) list = enif_make_list_cell(env, enif_mock_make_binary(env, blobs[i], strlen(blobs[i])), list);
    ERL_NIF_TERM mp = enif_make_list(env, 0), pp = enif_make_list(env, 0);
    for (int i = EB200_N_MUTATORS - 1; i >= 0; i--) mp = enif_make_list_cell(env, enif_make_int(env, eb200_mutator_default_pri(i)), mp);
    for (int i = EB200_N_PATTERNS - (bad ? 2 : 1); i >= 0; i--) pp = enif_make_list_cell(env, enif_make_int(env, eb200_pattern_default_pri(i)), pp);
    ERL_NIF_TERM args[10] = {list, enif_make_uint64(env, (uint64_t)n_cases), enif_make_tuple3(env, enif_make_int64(env, sa), enif_make_int64(env, sb), enif_make_int64(env, sc)),
                             mp, pp, enif_make_uint64(env, 1), enif_make_double(env, 1.0),
                             enif_make_tuple2(env, enif_mock_make_binary(env, "localhost", 9), enif_make_int(env, 51234)),
                             enif_make_tuple2(env, enif_make_int(env, 500), enif_make_int(env, 1)), enif_make_int(env, 0)};
    ERL_NIF_TERM res = entry->funcs[0].fptr(env, 10, args);
    int arity; const ERL_NIF_TERM* el; const char* tag = "?";
    if (!enif_get_tuple(env, res, &arity, &el) || !enif_get_atom_name(env, el[0], &tag)) { printf("bad_result\n"); return 2; }
    if (strcmp(tag, "ok")) { const char* why = "?"; enif_get_atom_name(env, el[1], &why); printf("error %s\n", why); entry->unload(env, priv); return bad ? 0 : 1; }
    ERL_NIF_TERM h, t = el[1]; int k = 0;
    while (enif_get_list_cell(env, t, &h, &t)) {
        ErlNifBinary b; int ar; const ERL_NIF_TERM* fe;
        if (enif_inspect_binary(env, h, &b)) printf("%d %zu %016llx\n", k, b.size, (unsigned long long)fnv(b.data, b.size));
        else if (enif_get_tuple(env, h, &ar, &fe) && ar == 4) { int s = 0, r = 0; enif_get_int(env, fe[2], &s); enif_get_int(env, fe[3], &r); printf("%d flagged %d %d\n", k, s, r); }
        k++;
    }
    enif_mock_free_env(env);          /* the binaries go: the resource destructor hands the pinned buffer back to the pool */
    entry->unload(NULL, priv);
    return k == n_cases ? 0 : 1;
}
