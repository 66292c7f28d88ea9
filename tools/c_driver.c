/* Plain-C caller of libsezkp_cuda.so — the same calls a Rust `sezkp-cuda` FFI crate makes, with no
 * Python in the process.  Known answers are the SURVEY.md §8c values (NTT([1..8]), evaluate_on_coset_pow2([1,2,3,4],3,3),
 * and the length / BLAKE3-pinned bytes of prove_v1 on the tau=1 demo block of sezkp-stark/tests/{air_ok,stream_fri_equiv}.rs).
 *
 *   c_driver [n_gpus] [proof_out_path]
 * exit 0 = all checks passed, 3 = no usable CUDA device (ENODEV), 1 = a check failed.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sezkp_cuda.h"

#define T 64

static int fail(const char* what, sezkp_ctx* ctx, int rc) {
    fprintf(stderr, "c_driver: %s failed: rc=%d (%s)\n", what, rc, sezkp_cuda_last_error(ctx));
    return 1;
}

int main(int argc, char** argv) {
    const int n_gpus = argc > 1 ? atoi(argv[1]) : 1;
    const char* out_path = argc > 2 ? argv[2] : NULL;
    if (sezkp_cuda_abi_version() != SEZKP_CUDA_ABI_VERSION) {
        fprintf(stderr, "c_driver: ABI mismatch\n");
        return 1;
    }
    sezkp_ctx* ctx = NULL;
    int rc;
    if (n_gpus > 1) {
        int ids[64];
        for (int i = 0; i < n_gpus && i < 64; i++) ids[i] = (argc > 3 && strcmp(argv[3], "same") == 0) ? 0 : i;
        rc = sezkp_cuda_create_multi(ids, n_gpus, &ctx);
    } else {
        rc = sezkp_cuda_create(0, &ctx);
    }
    if (rc == SEZKP_CUDA_ENODEV) {
        printf("ENODEV: %s\n", sezkp_cuda_last_error(NULL));
        return 3;
    }
    if (rc != 0) return fail("create", NULL, rc);
    printf("gpus=%d\n", (int)sezkp_cuda_group_size(ctx));

    /* forward_ntt_in_place([1..8]) (sezkp-ffts/src/ntt.rs:79-111) */
    uint64_t v[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    static const uint64_t want_ntt[8] = {36ULL, 18445622567621360637ULL, 18445618169507741693ULL, 1130298020461564ULL,
                                         18446744069414584317ULL, 18445613771394122749ULL, 1125899906842620ULL, 1121501793223676ULL};
    if ((rc = sezkp_ntt_batch(ctx, v, 3, 1, 0)) != 0) return fail("sezkp_ntt_batch", ctx, rc);
    if (memcmp(v, want_ntt, sizeof v) != 0) {
        fprintf(stderr, "c_driver: NTT([1..8]) mismatch\n");
        return 1;
    }
    /* evaluate_on_coset_pow2([1,2,3,4], k=3, shift=3) (coset.rs:85-102) */
    uint64_t co[4] = {1, 2, 3, 4}, ev[8];
    static const uint64_t want_coset[8] = {142ULL, 7481077014752257ULL, 18418033621790097383ULL, 18439137646161692162ULL,
                                           18446744069414584235ULL, 7718571727623169ULL, 28710447624486886ULL, 18439150843925101058ULL};
    if ((rc = sezkp_coset_lde_batch(ctx, co, 2, 1, 3, 1, ev)) != 0) return fail("sezkp_coset_lde_batch", ctx, rc);
    if (memcmp(ev, want_coset, sizeof ev) != 0) {
        fprintf(stderr, "c_driver: coset evaluation mismatch\n");
        return 1;
    }
    /* non-canonical input is rejected, the ctx stays usable */
    uint64_t bad[2] = {0xffffffff00000001ULL, 0};
    if (sezkp_ntt_batch(ctx, bad, 1, 1, 0) != SEZKP_CUDA_EINVAL) {
        fprintf(stderr, "c_driver: non-canonical input was not rejected\n");
        return 1;
    }

    /* prove_v1 on demo_block(64): mv = 1,0,1,0,...; write 5 when i%3==0; window [0,T-1]; manifest_root = [7;32] */
    uint64_t block_len[1] = {T};
    int64_t wl[1] = {0}, wr[1] = {T - 1};
    uint32_t in_off[1] = {0}, out_off[1] = {0};
    int8_t input_mv[T], mv[T];
    uint8_t wf[T];
    uint16_t ws[T];
    for (int i = 0; i < T; i++) {
        input_mv[i] = 0;
        mv[i] = (i % 2 == 0);
        wf[i] = (i % 3 == 0);
        ws[i] = (i % 3 == 0) ? 5 : 0;
        out_off[0] += (uint32_t)mv[i];
    }
    sezkp_trace_desc d;
    memset(&d, 0, sizeof d);
    d.tau = 1;
    d.n_blocks = 1;
    d.n_rows = T;
    d.block_len = block_len;
    d.win_left = wl;
    d.win_right = wr;
    d.head_in_off = in_off;
    d.head_out_off = out_off;
    d.input_mv = input_mv;
    d.mv = mv;
    d.write_flag = wf;
    d.write_sym = ws;
    uint8_t root[32];
    memset(root, 7, 32);
    size_t len = 0;
    if ((rc = sezkp_stark_v1_prove(ctx, &d, root, NULL, 0, &len)) != 0) return fail("sezkp_stark_v1_prove (size query)", ctx, rc);
    if (len != 197199) {  /* bincode length of prove_v1(demo_block(64), [7;32]) */
        fprintf(stderr, "c_driver: proof length %zu != 197199\n", len);
        return 1;
    }
    uint8_t* proof = (uint8_t*)malloc(len);
    size_t len2 = 0;
    if (sezkp_stark_v1_prove(ctx, &d, root, proof, len - 1, &len2) != SEZKP_CUDA_ERANGE || len2 != len) {
        fprintf(stderr, "c_driver: short buffer was not reported as ERANGE\n");
        return 1;
    }
    if ((rc = sezkp_stark_v1_prove(ctx, &d, root, proof, len, &len2)) != 0) return fail("sezkp_stark_v1_prove", ctx, rc);
    if (out_path) {
        FILE* f = fopen(out_path, "wb");
        if (!f || fwrite(proof, 1, len2, f) != len2) {
            fprintf(stderr, "c_driver: cannot write %s\n", out_path);
            return 1;
        }
        fclose(f);
    }
    printf("proof_len=%zu launches=%llu\n", len2, (unsigned long long)sezkp_cuda_launch_count(ctx, 0));
    free(proof);
    sezkp_cuda_destroy(ctx);
    printf("ok\n");
    return 0;
}
