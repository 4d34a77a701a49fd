/* c_abi_host_demo.c — a non-Python caller of the C-ABI (TEST PROGRAM): dlopen libcvcs_b200.so, run
 * cvcs_host_ce_fused on host buffers, check the results against an in-file scalar restatement of
 * nn.CrossEntropyLoss + torch.max + MulticlassConfusionMatrix (the same arithmetic as oracle/cvcs_oracle.c).
 *   gcc -O2 -I include tests/c_abi_host_demo.c -o demo -ldl -lm && ./demo cvcs_b200/libcvcs_b200.so
 * Prints "C-ABI DEMO OK ..." and exits 0 on success. */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cvcs_b200.h"

typedef int (*create_fn)(cvcs_host_ctx**, int, long long, int, int);
typedef int (*destroy_fn)(cvcs_host_ctx*);
typedef int (*ce_fn)(cvcs_host_ctx*, const void*, int, int, const void*, int, const float*, long long, int, int, int, int,
                     int, void*, void*, int, unsigned long long*, float*, double*);
typedef const char* (*err_fn)(void);

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "cvcs_b200/libcvcs_b200.so";
    void* h = dlopen(path, RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    create_fn create = (create_fn)dlsym(h, "cvcs_host_ctx_create");
    destroy_fn destroy = (destroy_fn)dlsym(h, "cvcs_host_ctx_destroy");
    ce_fn ce = (ce_fn)dlsym(h, "cvcs_host_ce_fused");
    err_fn last_error = (err_fn)dlsym(h, "cvcs_last_error");
    if (!create || !destroy || !ce || !last_error) { fprintf(stderr, "missing symbol\n"); return 2; }

    enum { B = 2, C = 7, H = 64, W = 64 };
    const long long hw = (long long)H * W, n = B * hw;
    float* x = malloc(sizeof(float) * n * C);
    float* d = malloc(sizeof(float) * n * C);
    unsigned char* t = malloc(n);
    unsigned char* am = malloc(n);
    float w[C];
    unsigned int s = 12345u;
    for (long long i = 0; i < n * C; ++i) { s = s * 1664525u + 1013904223u; x[i] = ((int)(s >> 8) % 2001 - 1000) * 0.006f; }
    for (long long i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; t[i] = (unsigned char)((s >> 10) % C); if ((s >> 20) % 10 == 0) t[i] = 255; }
    for (int c = 0; c < C; ++c) w[c] = 0.5f + 0.25f * c;

    cvcs_host_ctx* ctx = NULL;
    if (create(&ctx, 0, n, C, CVCS_F32)) { fprintf(stderr, "create: %s\n", last_error()); return 1; }
    unsigned long long cm[C * C]; memset(cm, 0, sizeof cm);
    float loss = 0.f; double sums[3];
    if (ce(ctx, x, CVCS_F32, CVCS_NCHW, t, CVCS_U8, w, 255, B, C, H, W, 1, d, am, CVCS_U8, cm, &loss, sums)) {
        fprintf(stderr, "cvcs_host_ce_fused: %s\n", last_error()); return 1;
    }
    /* scalar restatement */
    double lsum = 0, wsum = 0; unsigned long long rcm[C * C]; memset(rcm, 0, sizeof rcm);
    double max_gerr = 0, max_g = 0; long long am_bad = 0;
    for (int b = 0; b < B; ++b) for (long long p = 0; p < hw; ++p) if (t[b * hw + p] != 255) wsum += w[t[b * hw + p]];
    for (int b = 0; b < B; ++b) for (long long p = 0; p < hw; ++p) {
        const float* xp = x + (long long)b * C * hw + p;
        int arg = 0; float m = xp[0];
        for (int c = 1; c < C; ++c) if (xp[c * hw] > m) { m = xp[c * hw]; arg = c; }
        double se = 0; for (int c = 0; c < C; ++c) se += exp((double)xp[c * hw] - m);
        const int tv = t[b * hw + p];
        if (am[b * hw + p] != arg) ++am_bad;
        for (int c = 0; c < C; ++c) {
            double g = 0;
            if (tv != 255) { g = w[tv] * (exp((double)xp[c * hw] - m) / se - (c == tv)) / wsum; }
            const double e = fabs(g - d[(long long)b * C * hw + c * hw + p]);
            if (e > max_gerr) max_gerr = e;
            if (fabs(g) > max_g) max_g = fabs(g);
        }
        if (tv != 255) { lsum += w[tv] * (m + log(se) - xp[tv * hw]); rcm[tv * C + arg]++; }
    }
    const double lref = lsum / wsum;
    int ok = fabs(loss - lref) <= 1e-5 * fabs(lref) && max_gerr <= 1e-5 * max_g && am_bad == 0 && memcmp(cm, rcm, sizeof cm) == 0;
    destroy(ctx);
    printf("C-ABI DEMO %s loss %.7f (ref %.7f) grad err %.2e of %.2e argmax mismatches %lld\n", ok ? "OK" : "FAILED", loss, lref,
           max_gerr, max_g, am_bad);
    return ok ? 0 : 1;
}
