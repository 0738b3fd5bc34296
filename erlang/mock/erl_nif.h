/* A stand-in for the part of OTP's erl_nif.h that erlang/erlamsa_b200_nif.c uses, so that the NIF compiles and can be driven
 * from C in an image without Erlang/OTP (tests/test_nif_harness.py). NOT the real header: terms are heap objects of the mock
 * (erl_nif_mock.c), there is no VM, no scheduler and no garbage collector -- enif_mock_free_env releases what an env made. */
#ifndef ERLAMSA_B200_MOCK_ERL_NIF_H
#define ERLAMSA_B200_MOCK_ERL_NIF_H
#include <stddef.h>
#include <stdint.h>

typedef uintptr_t ERL_NIF_TERM;
typedef struct enif_environment_t ErlNifEnv;
typedef uint64_t ErlNifUInt64;
typedef int64_t ErlNifSInt64;
typedef struct { size_t size; unsigned char* data; void* ref_bin; void* spare[2]; } ErlNifBinary;
typedef struct enif_resource_type_t ErlNifResourceType;
typedef void ErlNifResourceDtor(ErlNifEnv*, void*);
typedef enum { ERL_NIF_RT_CREATE = 1, ERL_NIF_RT_TAKEOVER = 2 } ErlNifResourceFlags;
typedef struct ErlNifMutex_ ErlNifMutex;
typedef struct { const char* name; unsigned arity; ERL_NIF_TERM (*fptr)(ErlNifEnv*, int, const ERL_NIF_TERM[]); unsigned flags; } ErlNifFunc;
typedef struct { const char* name; int num_of_funcs; ErlNifFunc* funcs; int (*load)(ErlNifEnv*, void**, ERL_NIF_TERM);
                 void* reload; void* upgrade; void (*unload)(ErlNifEnv*, void*); } ErlNifEntry;
#define ERL_NIF_DIRTY_JOB_IO_BOUND 2
#define ERL_NIF_DIRTY_JOB_CPU_BOUND 1

#define ERL_NIF_INIT(MOD, FUNCS, LOAD, RELOAD, UPGRADE, UNLOAD) \
    static ErlNifEntry mock_entry = {#MOD, (int)(sizeof(FUNCS) / sizeof(FUNCS[0])), FUNCS, LOAD, (void*)RELOAD, (void*)UPGRADE, UNLOAD}; \
    ErlNifEntry* nif_init(void) { return &mock_entry; }

void* enif_alloc(size_t);
void enif_free(void*);
ERL_NIF_TERM enif_make_atom(ErlNifEnv*, const char*);
ERL_NIF_TERM enif_make_int(ErlNifEnv*, int);
ERL_NIF_TERM enif_make_uint64(ErlNifEnv*, ErlNifUInt64);
ERL_NIF_TERM enif_make_int64(ErlNifEnv*, ErlNifSInt64);
ERL_NIF_TERM enif_make_double(ErlNifEnv*, double);
ERL_NIF_TERM enif_make_tuple2(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_tuple3(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_tuple4(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_list(ErlNifEnv*, unsigned cnt, ...);     /* only cnt == 0 is used */
ERL_NIF_TERM enif_make_list_cell(ErlNifEnv*, ERL_NIF_TERM head, ERL_NIF_TERM tail);
unsigned char* enif_make_new_binary(ErlNifEnv*, size_t, ERL_NIF_TERM*);
ERL_NIF_TERM enif_make_resource_binary(ErlNifEnv*, void* obj, const void* data, size_t size);
int enif_get_list_length(ErlNifEnv*, ERL_NIF_TERM, unsigned*);
int enif_get_list_cell(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM* head, ERL_NIF_TERM* tail);
int enif_get_int(ErlNifEnv*, ERL_NIF_TERM, int*);
int enif_get_uint(ErlNifEnv*, ERL_NIF_TERM, unsigned*);
int enif_get_uint64(ErlNifEnv*, ERL_NIF_TERM, ErlNifUInt64*);
int enif_get_int64(ErlNifEnv*, ERL_NIF_TERM, ErlNifSInt64*);
int enif_get_double(ErlNifEnv*, ERL_NIF_TERM, double*);
int enif_get_tuple(ErlNifEnv*, ERL_NIF_TERM, int* arity, const ERL_NIF_TERM** array);
int enif_get_atom_name(ErlNifEnv*, ERL_NIF_TERM, const char** name);      /* mock-only convenience */
int enif_inspect_binary(ErlNifEnv*, ERL_NIF_TERM, ErlNifBinary*);
int enif_inspect_iolist_as_binary(ErlNifEnv*, ERL_NIF_TERM, ErlNifBinary*);
ErlNifResourceType* enif_open_resource_type(ErlNifEnv*, const char* module, const char* name, ErlNifResourceDtor*, ErlNifResourceFlags, ErlNifResourceFlags* tried);
void* enif_alloc_resource(ErlNifResourceType*, size_t);
void enif_release_resource(void*);
int enif_keep_resource(void*);
ErlNifMutex* enif_mutex_create(char* name);
void enif_mutex_destroy(ErlNifMutex*);
void enif_mutex_lock(ErlNifMutex*);
void enif_mutex_unlock(ErlNifMutex*);

/* mock-only: environments for the harness */
ErlNifEnv* enif_mock_new_env(void);
void enif_mock_free_env(ErlNifEnv*);      /* frees the env's terms; resource binaries drop their reference (destructors run) */
ERL_NIF_TERM enif_mock_make_binary(ErlNifEnv*, const void* data, size_t size);
#endif
