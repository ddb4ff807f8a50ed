/* A plain C99 caller of the C ABI (test infrastructure): proves that include/arrow_b200.h is valid C, that the
 * library can be bound with nothing but dlopen + plain pointers and sizes, and that without a GPU the product fails
 * loudly instead of computing on the CPU.  Usage: c_abi_caller /path/to/libarrow_b200.so
 * Exit code 0: behaved as expected (prints "gpu" or "no-gpu" and, with a GPU, the checksum of a tiny product). */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "../include/arrow_b200.h"

typedef int (*version_fn)(void);
typedef int (*create_fn)(int, void *, arrow_ctx **);
typedef void (*destroy_fn)(arrow_ctx *);
typedef const char *(*error_fn)(const arrow_ctx *);
typedef int (*upload_fn)(arrow_ctx *, int64_t, int64_t, int64_t, const void *, int, const void *, int, const float *, int *);
typedef int (*alloc_fn)(arrow_ctx *, int64_t, int, int *);
typedef int (*copy_fn)(arrow_ctx *, int, int64_t, int64_t, float *);
typedef int (*put_fn)(arrow_ctx *, int, int64_t, int64_t, const float *);
typedef int (*spmm_fn)(arrow_ctx *, int, int, int, int, int, int);
typedef int (*sync_fn)(arrow_ctx *);

/* POSIX idiom: dlsym returns an object pointer, ISO C has no cast from it to a function pointer */
#define SYM(type, name) type p_##name; *(void **)(&p_##name) = dlsym(lib, #name); \
    if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

int main(int argc, char **argv) {
    void *lib;
    arrow_ctx *ctx = NULL;
    int rc;
    if (argc < 2) return 2;
    lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    {
        SYM(version_fn, arrow_b200_abi_version)
        SYM(create_fn, arrow_ctx_create)
        SYM(destroy_fn, arrow_ctx_destroy)
        SYM(error_fn, arrow_last_error)
        SYM(upload_fn, arrow_csr_upload)
        SYM(alloc_fn, arrow_dense_alloc)
        SYM(put_fn, arrow_dense_h2d)
        SYM(copy_fn, arrow_dense_d2h)
        SYM(spmm_fn, arrow_spmm)
        SYM(sync_fn, arrow_sync)
        if (p_arrow_b200_abi_version() != ARROW_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 3; }
        rc = p_arrow_ctx_create(0, NULL, &ctx);
        if (rc != ARROW_OK) {
            const char *msg = p_arrow_last_error(NULL);
            if (rc == ARROW_ERR_CUDA && ctx == NULL && msg && strstr(msg, "no CPU fallback")) { printf("no-gpu\n"); return 0; }
            fprintf(stderr, "unexpected failure %d: %s\n", rc, msg ? msg : "(null)");
            return 4;
        }
        {   /* 2x2 identity times a 2x4 tile */
            const int32_t indptr[3] = {0, 1, 2}, indices[2] = {0, 1};
            const float vals[2] = {1.0f, 1.0f}, x[8] = {1, 2, 3, 4, 5, 6, 7, 8};
            float c[8] = {0};
            int A = -1, X = -1, C = -1, i;
            float sum = 0.0f;
            if (p_arrow_csr_upload(ctx, 2, 2, 2, indptr, 4, indices, 4, vals, &A) || p_arrow_dense_alloc(ctx, 2, 4, &X) ||
                p_arrow_dense_alloc(ctx, 2, 4, &C) || p_arrow_dense_h2d(ctx, X, 0, 2, x) ||
                p_arrow_spmm(ctx, A, X, C, -1, 0, ARROW_VARIANT_AUTO) || p_arrow_dense_d2h(ctx, C, 0, 2, c) || p_arrow_sync(ctx)) {
                fprintf(stderr, "call failed: %s\n", p_arrow_last_error(ctx));
                return 5;
            }
            for (i = 0; i < 8; ++i) sum += c[i];
            printf("gpu %.1f\n", sum);
            p_arrow_ctx_destroy(ctx);
            return sum == 36.0f ? 0 : 6;
        }
    }
}
