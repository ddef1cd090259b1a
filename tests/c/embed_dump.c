/* embed_dump.c -- plain C consumer of the C ABI: one ResNet-50 embedding forward (ssg_embed_*) with deterministic
 * pseudo-random weights and images, features written to a file.  The kernel variants are selected by environment
 * variables that are read once per process, so a variant is checked by running this program twice and comparing the
 * dumps byte for byte (tools/gpu_next.sh):
 *     tests/c/_build/embed_dump /tmp/a.bin;  SSG_CONV_EPI2=1 tests/c/_build/embed_dump /tmp/b.bin;  cmp /tmp/a.bin /tmp/b.bin
 * The forward runs twice in the process (the second run replays a recorded CUDA graph where one is used) and the two
 * results must agree.  Usage: embed_dump OUT [n_images=21] [batch_max=32] */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ssg_b200.h"

#define CK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "FAIL %s:%d %s -> %d (%s)\n", __FILE__, __LINE__, #x, rc_, ssg_last_error()); return 2; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 2; } } while (0)

static unsigned long long g_seed = 20261017ull;
static float urand(void) { g_seed = g_seed * 6364136223846793005ull + 1442695040888963407ull; return (float)((double)(g_seed >> 11) / 9007199254740992.0); }

static int upload(float** d, const float* h, size_t n) {
    CU(cudaMalloc((void**)d, sizeof(float) * n));
    CU(cudaMemcpy(*d, h, sizeof(float) * n, cudaMemcpyHostToDevice));
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: embed_dump OUT [n_images] [batch_max]\n"); return 2; }
    const int n = argc > 2 ? atoi(argv[2]) : 21, batch = argc > 3 ? atoi(argv[3]) : 32, banks = 3;
    ssg_embed_plan* plan = NULL;
    CK(ssg_embed_plan_create(&plan, 0, batch, 256, 128));
    const int nl = ssg_embed_num_layers();
    for (int idx = 0; idx < nl; ++idx) {
        int cin, cout, k, stride;
        char ck[64], bk[64];
        CK(ssg_embed_layer_info(idx, &cin, &cout, &k, &stride, ck, bk, sizeof(ck)));
        const size_t wn = (size_t)cout * cin * k * k;
        float* w = (float*)malloc(sizeof(float) * wn);
        float* bn = (float*)malloc(sizeof(float) * 4 * cout);
        const float a = sqrtf(6.0f / (float)(cin * k * k));              /* keeps the activations O(1) through ReLUs */
        for (size_t i = 0; i < wn; ++i) w[i] = (2.f * urand() - 1.f) * a;
        for (int c = 0; c < cout; ++c) {
            bn[c] = 0.5f + urand();                                      /* gamma */
            bn[cout + c] = 0.2f * (urand() - 0.5f);                      /* beta  */
            bn[2 * cout + c] = 0.2f * (urand() - 0.5f);                  /* running mean */
            bn[3 * cout + c] = 0.5f + urand();                           /* running var  */
        }
        float *d_w, *d_bn;
        if (upload(&d_w, w, wn) || upload(&d_bn, bn, (size_t)4 * cout)) return 2;
        CK(ssg_embed_load_layer(plan, idx, d_w, d_bn, d_bn + cout, d_bn + 2 * cout, d_bn + 3 * cout, 1e-5f, NULL));
        CU(cudaDeviceSynchronize());
        cudaFree(d_w); cudaFree(d_bn); free(w); free(bn);
    }
    const size_t in = (size_t)n * 3 * 256 * 128, fn = (size_t)banks * n * 2048;
    float* img = (float*)malloc(sizeof(float) * in);
    for (size_t i = 0; i < in; ++i) img[i] = 4.f * urand() - 2.f;
    float *d_img, *d_f;
    if (upload(&d_img, img, in)) return 2;
    CU(cudaMalloc((void**)&d_f, sizeof(float) * fn));
    float* f1 = (float*)malloc(sizeof(float) * fn);
    float* f2 = (float*)malloc(sizeof(float) * fn);
    for (int rep = 0; rep < 2; ++rep) {
        CU(cudaMemset(d_f, 0xff, sizeof(float) * fn));
        CK(ssg_embed_forward(plan, d_img, n, 2, 0, 1, d_f, (size_t)n * 2048, 0, NULL));
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(rep ? f2 : f1, d_f, sizeof(float) * fn, cudaMemcpyDeviceToHost));
    }
    size_t bad = 0;
    double sum = 0;
    for (size_t i = 0; i < fn; ++i) { bad += !(f1[i] == f1[i]) || isinf(f1[i]); sum += f1[i]; }
    const int same = memcmp(f1, f2, sizeof(float) * fn) == 0;
    FILE* fo = fopen(argv[1], "wb");
    if (!fo || fwrite(f1, sizeof(float), fn, fo) != fn) { fprintf(stderr, "FAIL: cannot write %s\n", argv[1]); return 2; }
    fclose(fo);
    printf("embed_dump: %d images, %zu features, sum %.9g, non-finite %zu, second forward identical: %s -> %s\n", n, fn, sum, bad,
           same ? "yes" : "NO", (bad || !same) ? "EMBED_DUMP FAILED" : "EMBED_DUMP OK");
    ssg_embed_plan_destroy(plan);
    return (bad || !same) ? 1 : 0;
}
