/* C harness for the drop-in boundary: include/g16_cuda.h must compile as plain C (C99, -pedantic), every declared
 * function must link, and a caller written in C gets the same bytes as the Python binding.
 *
 *   abi_harness symbols          link-time check only: prints the number of entry points it holds the address of
 *   abi_harness nodevice         g16_ctx_create must fail with G16_ERR_NO_DEVICE (CPU box: no CPU fallback)
 *   abi_harness msm <file>       file = u64 n | n x 12 u64 points | n bytes flags | n x 4 u64 scalars | 12 u64 expected |
 *                                1 byte expected flag; runs g16_g1_msm_oneshot and the resident path
 *                                (g16_g1_bases_upload + g16_bases_precompute + g16_g1_msm) and compares both
 * TEST INFRASTRUCTURE: built and run by tests/test_abi_exports.py / tests/test_gpu_parity.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "g16_cuda.h"

typedef void (*fn_t)(void);
#define F(x) (fn_t)(x)
static fn_t table[] = {
    F(g16_ctx_create), F(g16_ctx_destroy), F(g16_last_error), F(g16_ctx_set_stream), F(g16_ctx_synchronize),
    F(g16_ctx_set_window_bits), F(g16_ctx_set_h2d_pipeline_min), F(g16_ctx_set_item_max), F(g16_device_count), F(g16_version),
    F(g16_g1_bases_upload), F(g16_g2_bases_upload), F(g16_g1_bases_from_device), F(g16_g2_bases_from_device),
    F(g16_bases_precompute), F(g16_bases_free), F(g16_bases_len), F(g16_g1_msm), F(g16_g2_msm), F(g16_g1_msm_oneshot),
    F(g16_g2_msm_oneshot), F(g16_g1_msm_device), F(g16_g2_msm_device), F(g16_g1_msm_async), F(g16_g2_msm_async),
    F(g16_g1_combine_partials_device), F(g16_g2_combine_partials_device), F(g16_g1_fixed_base_mul), F(g16_g2_fixed_base_mul),
    F(g16_g1_fixed_base_mul_device), F(g16_g2_fixed_base_mul_device), F(g16_pk_upload), F(g16_pk_precompute), F(g16_pk_precompute_bits), F(g16_pk_free),
    F(g16_prove), F(g16_quotient_h), F(g16_quotient_h_device), F(g16_r1cs_upload), F(g16_r1cs_free), F(g16_r1cs_domain_size), F(g16_r1cs_domain_evals),
    F(g16_r1cs_eval_at), F(g16_setup_crs), F(g16_prove_r1cs), F(g16_g1_serialize), F(g16_g2_serialize), F(g16_g1_deserialize),
    F(g16_g2_deserialize), F(g16_proof_serialize), F(g16_proof_deserialize), F(g16_launch_count), F(g16_ctx_enable_stage_timing),
    F(g16_ctx_last_stage_ms), F(g16_ctx_prove_timeline), F(g16_debug_fq_op), F(g16_debug_fr_from_mont), F(g16_debug_g1_add), F(g16_debug_g2_add),
};

static int run_msm(const char *path) {
    FILE *f = fopen(path, "rb");
    uint64_t n = 0, *pts, *sc, expect[12], got[12];
    uint8_t *inf, expect_inf = 0, got_inf = 0;
    g16_ctx *ctx = NULL;
    g16_bases *bases = NULL;
    unsigned used = 0;
    int rc;
    if (!f || fread(&n, 8, 1, f) != 1) { fprintf(stderr, "cannot read %s\n", path); return 2; }
    pts = (uint64_t *)malloc((size_t)n * 96 + 8);
    inf = (uint8_t *)malloc((size_t)n + 8);
    sc = (uint64_t *)malloc((size_t)n * 32 + 8);
    if (fread(pts, 96, n, f) != n || fread(inf, 1, n, f) != n || fread(sc, 32, n, f) != n || fread(expect, 8, 12, f) != 12 ||
        fread(&expect_inf, 1, 1, f) != 1) { fprintf(stderr, "short file\n"); return 2; }
    fclose(f);
    rc = g16_ctx_create(NULL, 0, &ctx);
    if (rc != G16_OK) { fprintf(stderr, "ctx: %s (%d)\n", g16_last_error(NULL), rc); return 3; }
    rc = g16_g1_msm_oneshot(ctx, pts, inf, sc, (size_t)n, got, &got_inf);
    if (rc != G16_OK) { fprintf(stderr, "oneshot: %s (%d)\n", g16_last_error(ctx), rc); return 4; }
    if (got_inf != expect_inf || memcmp(got, expect, 96) != 0) { fprintf(stderr, "oneshot result differs\n"); return 5; }
    rc = g16_g1_bases_upload(ctx, pts, inf, (size_t)n, &bases);
    if (rc == G16_OK) rc = g16_bases_precompute(ctx, bases, 0, 0, &used);
    memset(got, 0, sizeof got);
    if (rc == G16_OK) rc = g16_g1_msm(ctx, bases, sc, (size_t)n, got, &got_inf);
    if (rc != G16_OK) { fprintf(stderr, "resident: %s (%d)\n", g16_last_error(ctx), rc); return 6; }
    if (got_inf != expect_inf || memcmp(got, expect, 96) != 0) { fprintf(stderr, "resident result differs\n"); return 7; }
    /* length mismatch: more scalars than bases -> G16_ERR_LENGTH (ark: Err(min_len)) */
    if (g16_g1_msm(ctx, bases, sc, (size_t)n + 1, got, &got_inf) != G16_ERR_LENGTH) { fprintf(stderr, "length check\n"); return 8; }
    printf("msm ok n=%llu precompute_c=%u launches=%llu %s\n", (unsigned long long)n, used, g16_launch_count(), g16_version());
    g16_bases_free(bases);
    g16_ctx_destroy(ctx);
    free(pts); free(inf); free(sc);
    return 0;
}

int main(int argc, char **argv) {
    size_t k, cnt = sizeof table / sizeof table[0];
    for (k = 0; k < cnt; ++k)
        if (!table[k]) return 1;
    if (argc >= 2 && strcmp(argv[1], "symbols") == 0) {
        printf("entry points: %u  %s\n", (unsigned)cnt, g16_version());
        return 0;
    }
    if (argc >= 2 && strcmp(argv[1], "nodevice") == 0) {
        g16_ctx *ctx = NULL;
        int rc = g16_ctx_create(NULL, 0, &ctx);
        printf("g16_ctx_create rc=%d msg=%s\n", rc, g16_last_error(NULL));
        if (rc == G16_OK) { g16_ctx_destroy(ctx); return 10; }
        return rc == G16_ERR_NO_DEVICE ? 0 : 11;
    }
    if (argc >= 3 && strcmp(argv[1], "msm") == 0) return run_msm(argv[2]);
    fprintf(stderr, "usage: abi_harness symbols | nodevice | msm <file>\n");
    return 64;
}
