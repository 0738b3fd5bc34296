/* Implementation of erlang/mock/erl_nif.h: a tiny term heap. TEST INFRASTRUCTURE for the NIF shim, not product code. */
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "erl_nif.h"

enum { T_ATOM = 1, T_INT, T_DOUBLE, T_TUPLE, T_NIL, T_CONS, T_BIN };
typedef struct Term {
    int tag;
    union {
        const char* atom; int64_t i; double d;
        struct { int n; ERL_NIF_TERM* el; } tup;
        struct { ERL_NIF_TERM h, t; } cons;
        struct { unsigned char* data; size_t size; void* resource; int owned; } bin;
    } u;
    struct Term* next;   /* env's allocation list */
} Term;
struct enif_environment_t { Term* terms; };
struct enif_resource_type_t { ErlNifResourceDtor* dtor; };
typedef struct { ErlNifResourceType* type; int refs; } ResHdr;
struct ErlNifMutex_ { pthread_mutex_t m; };

static Term* mk(ErlNifEnv* env, int tag) { Term* t = (Term*)calloc(1, sizeof(Term)); t->tag = tag; t->next = env->terms; env->terms = t; return t; }
#define T(x) ((Term*)(x))

void* enif_alloc(size_t n) { return malloc(n); }
void enif_free(void* p) { free(p); }
ERL_NIF_TERM enif_make_atom(ErlNifEnv* e, const char* s) { Term* t = mk(e, T_ATOM); t->u.atom = s; return (ERL_NIF_TERM)t; }
ERL_NIF_TERM enif_make_int(ErlNifEnv* e, int v) { Term* t = mk(e, T_INT); t->u.i = v; return (ERL_NIF_TERM)t; }
ERL_NIF_TERM enif_make_uint64(ErlNifEnv* e, ErlNifUInt64 v) { Term* t = mk(e, T_INT); t->u.i = (int64_t)v; return (ERL_NIF_TERM)t; }
ERL_NIF_TERM enif_make_int64(ErlNifEnv* e, ErlNifSInt64 v) { Term* t = mk(e, T_INT); t->u.i = v; return (ERL_NIF_TERM)t; }
ERL_NIF_TERM enif_make_double(ErlNifEnv* e, double v) { Term* t = mk(e, T_DOUBLE); t->u.d = v; return (ERL_NIF_TERM)t; }
static ERL_NIF_TERM tuple(ErlNifEnv* e, int n, ...) {
    Term* t = mk(e, T_TUPLE); t->u.tup.n = n; t->u.tup.el = (ERL_NIF_TERM*)malloc(sizeof(ERL_NIF_TERM) * n);
    va_list ap; va_start(ap, n); for (int i = 0; i < n; i++) t->u.tup.el[i] = va_arg(ap, ERL_NIF_TERM); va_end(ap);
    return (ERL_NIF_TERM)t;
}
ERL_NIF_TERM enif_make_tuple2(ErlNifEnv* e, ERL_NIF_TERM a, ERL_NIF_TERM b) { return tuple(e, 2, a, b); }
ERL_NIF_TERM enif_make_tuple3(ErlNifEnv* e, ERL_NIF_TERM a, ERL_NIF_TERM b, ERL_NIF_TERM c) { return tuple(e, 3, a, b, c); }
ERL_NIF_TERM enif_make_tuple4(ErlNifEnv* e, ERL_NIF_TERM a, ERL_NIF_TERM b, ERL_NIF_TERM c, ERL_NIF_TERM d) { return tuple(e, 4, a, b, c, d); }
ERL_NIF_TERM enif_make_list(ErlNifEnv* e, unsigned cnt, ...) { (void)cnt; return (ERL_NIF_TERM)mk(e, T_NIL); }
ERL_NIF_TERM enif_make_list_cell(ErlNifEnv* e, ERL_NIF_TERM h, ERL_NIF_TERM tl) { Term* t = mk(e, T_CONS); t->u.cons.h = h; t->u.cons.t = tl; return (ERL_NIF_TERM)t; }
unsigned char* enif_make_new_binary(ErlNifEnv* e, size_t n, ERL_NIF_TERM* out) {
    Term* t = mk(e, T_BIN); t->u.bin.data = (unsigned char*)malloc(n ? n : 1); t->u.bin.size = n; t->u.bin.owned = 1; *out = (ERL_NIF_TERM)t; return t->u.bin.data;
}
ERL_NIF_TERM enif_mock_make_binary(ErlNifEnv* e, const void* data, size_t n) { ERL_NIF_TERM t; unsigned char* p = enif_make_new_binary(e, n, &t); memcpy(p, data, n); return t; }
ERL_NIF_TERM enif_make_resource_binary(ErlNifEnv* e, void* obj, const void* data, size_t n) {
    Term* t = mk(e, T_BIN); t->u.bin.data = (unsigned char*)data; t->u.bin.size = n; t->u.bin.resource = obj; enif_keep_resource(obj); return (ERL_NIF_TERM)t;
}
int enif_get_list_length(ErlNifEnv* e, ERL_NIF_TERM l, unsigned* n) {
    (void)e; unsigned c = 0; Term* t = T(l);
    while (t->tag == T_CONS) { c++; t = T(t->u.cons.t); }
    if (t->tag != T_NIL) return 0;
    *n = c; return 1;
}
int enif_get_list_cell(ErlNifEnv* e, ERL_NIF_TERM l, ERL_NIF_TERM* h, ERL_NIF_TERM* tl) { (void)e; if (T(l)->tag != T_CONS) return 0; *h = T(l)->u.cons.h; *tl = T(l)->u.cons.t; return 1; }
int enif_get_int(ErlNifEnv* e, ERL_NIF_TERM t, int* v) { (void)e; if (T(t)->tag != T_INT || T(t)->u.i < -2147483648LL || T(t)->u.i > 2147483647LL) return 0; *v = (int)T(t)->u.i; return 1; }
int enif_get_uint(ErlNifEnv* e, ERL_NIF_TERM t, unsigned* v) { (void)e; if (T(t)->tag != T_INT || T(t)->u.i < 0 || T(t)->u.i > 4294967295LL) return 0; *v = (unsigned)T(t)->u.i; return 1; }
int enif_get_uint64(ErlNifEnv* e, ERL_NIF_TERM t, ErlNifUInt64* v) { (void)e; if (T(t)->tag != T_INT || T(t)->u.i < 0) return 0; *v = (uint64_t)T(t)->u.i; return 1; }
int enif_get_int64(ErlNifEnv* e, ERL_NIF_TERM t, ErlNifSInt64* v) { (void)e; if (T(t)->tag != T_INT) return 0; *v = T(t)->u.i; return 1; }
int enif_get_double(ErlNifEnv* e, ERL_NIF_TERM t, double* v) { (void)e; if (T(t)->tag != T_DOUBLE) return 0; *v = T(t)->u.d; return 1; }
int enif_get_tuple(ErlNifEnv* e, ERL_NIF_TERM t, int* n, const ERL_NIF_TERM** a) { (void)e; if (T(t)->tag != T_TUPLE) return 0; *n = T(t)->u.tup.n; *a = T(t)->u.tup.el; return 1; }
int enif_get_atom_name(ErlNifEnv* e, ERL_NIF_TERM t, const char** s) { (void)e; if (T(t)->tag != T_ATOM) return 0; *s = T(t)->u.atom; return 1; }
int enif_inspect_binary(ErlNifEnv* e, ERL_NIF_TERM t, ErlNifBinary* b) { (void)e; if (T(t)->tag != T_BIN) return 0; b->data = T(t)->u.bin.data; b->size = T(t)->u.bin.size; return 1; }
int enif_inspect_iolist_as_binary(ErlNifEnv* e, ERL_NIF_TERM t, ErlNifBinary* b) { return enif_inspect_binary(e, t, b); }
ErlNifResourceType* enif_open_resource_type(ErlNifEnv* e, const char* m, const char* n, ErlNifResourceDtor* d, ErlNifResourceFlags f, ErlNifResourceFlags* tried) {
    (void)e; (void)m; (void)n; (void)f; ErlNifResourceType* t = (ErlNifResourceType*)calloc(1, sizeof(*t)); t->dtor = d; if (tried) *tried = ERL_NIF_RT_CREATE; return t;
}
void* enif_alloc_resource(ErlNifResourceType* type, size_t n) { ResHdr* h = (ResHdr*)calloc(1, sizeof(ResHdr) + n); h->type = type; h->refs = 1; return h + 1; }
int enif_keep_resource(void* obj) { ((ResHdr*)obj - 1)->refs++; return 1; }
void enif_release_resource(void* obj) { ResHdr* h = (ResHdr*)obj - 1; if (--h->refs == 0) { if (h->type->dtor) h->type->dtor(NULL, obj); free(h); } }
ErlNifMutex* enif_mutex_create(char* name) { (void)name; ErlNifMutex* m = (ErlNifMutex*)malloc(sizeof(*m)); pthread_mutex_init(&m->m, NULL); return m; }
void enif_mutex_destroy(ErlNifMutex* m) { pthread_mutex_destroy(&m->m); free(m); }
void enif_mutex_lock(ErlNifMutex* m) { pthread_mutex_lock(&m->m); }
void enif_mutex_unlock(ErlNifMutex* m) { pthread_mutex_unlock(&m->m); }
ErlNifEnv* enif_mock_new_env(void) { return (ErlNifEnv*)calloc(1, sizeof(ErlNifEnv)); }
void enif_mock_free_env(ErlNifEnv* e) {
    Term* t = e->terms;
    while (t) {
        Term* n = t->next;
        if (t->tag == T_TUPLE) free(t->u.tup.el);
        if (t->tag == T_BIN) { if (t->u.bin.owned) free(t->u.bin.data); if (t->u.bin.resource) enif_release_resource(t->u.bin.resource); }
        free(t); t = n;
    }
    free(e);
}
