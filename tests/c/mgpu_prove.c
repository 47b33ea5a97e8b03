/* A C host driving the multi-GPU prover through include/ligero_b200.h only -- what a Rust `LigeroCircuit::prove` does after
 * binding lg_mgpu_* (INTEGRATION.md section 5).  Builds a seeded synthetic circuit, proves it on GPU 0 (lg_prove) and on G
 * GPUs from ONE process (lg_mgpu_prove), and compares the serialized proofs byte for byte; also commits a small matrix
 * over all GPUs.
 *
 *   gcc -std=c11 -O2 -I include tests/c/mgpu_prove.c -o /tmp/mgpu_prove -L ligero_b200 -lligero_b200 -Wl,-rpath,$PWD/ligero_b200
 *   /tmp/mgpu_prove 2 [log2_gates]        -> prints C_MGPU_OK
 * tests/test_abi.py compiles it without a GPU (the header must be valid C); tests/test_gpu_multi.py runs it on >= 2 GPUs. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ligero_b200.h"

#define CHECK(expr, who)                                                              \
  do {                                                                                \
    int st_ = (expr);                                                                 \
    if (st_ != LG_OK) {                                                               \
      fprintf(stderr, "%s failed: %d (%s)\n", #expr, st_, (who) ? (who) : "");        \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

static int serialize(const lg_proof* p, uint8_t** buf, size_t* len) {
  if (lg_proof_serialize(p, NULL, 0, len) != LG_OK) return 1;
  *buf = (uint8_t*)malloc(*len);
  return *buf == NULL || lg_proof_serialize(p, *buf, *len, len) != LG_OK;
}

int main(int argc, char** argv) {
  const int n_gpus = argc > 1 ? atoi(argv[1]) : 2;
  const int log_gates = argc > 2 ? atoi(argv[2]) : 14;
  lg_circuit* circuit = NULL;
  size_t output = 0, var_idx[2];
  uint64_t var_vals[8];
  CHECK(lg_circuit_synthetic((size_t)1 << log_gates, 2024, &circuit, &output, var_idx, var_vals), "");

  /* single GPU */
  lg_ctx* ctx = NULL;
  CHECK(lg_ctx_create(0, &ctx), "no usable GPU");
  lg_ligero* one = NULL;
  CHECK(lg_ligero_new(ctx, circuit, &output, 1, 128, &one), lg_last_error(ctx));
  lg_sponge* sponge = NULL;
  lg_proof* proof1 = NULL;
  CHECK(lg_sponge_test(&sponge), "");
  CHECK(lg_prove(one, var_idx, var_vals, 2, 1, sponge, &proof1), lg_last_error(ctx));
  uint8_t sq1[32], sq2[32];
  CHECK(lg_sponge_squeeze_bytes(sponge, sq1, 32), "");
  lg_sponge_free(sponge);

  /* the same call over n_gpus GPUs from this process */
  lg_mgpu* box = NULL;
  CHECK(lg_mgpu_create(NULL, n_gpus, &box), "lg_mgpu_create");
  lg_mligero* many = NULL;
  CHECK(lg_mgpu_ligero_new(box, circuit, &output, 1, 128, &many), lg_mgpu_last_error(box));
  lg_proof* proof2 = NULL;
  CHECK(lg_sponge_test(&sponge), "");
  CHECK(lg_mgpu_prove(many, var_idx, var_vals, 2, 1, sponge, &proof2), lg_mgpu_last_error(box));
  CHECK(lg_sponge_squeeze_bytes(sponge, sq2, 32), "");

  uint8_t *b1 = NULL, *b2 = NULL;
  size_t l1 = 0, l2 = 0;
  if (serialize(proof1, &b1, &l1) || serialize(proof2, &b2, &l2)) return 1;
  const int same = l1 == l2 && memcmp(b1, b2, l1) == 0 && memcmp(sq1, sq2, 32) == 0;
  printf("2^%d gates on %d GPUs: proof %zu bytes, %s the single-GPU proof; sponge advanced identically: %s\n", log_gates, n_gpus, l2,
         same ? "==" : "!=", memcmp(sq1, sq2, 32) == 0 ? "yes" : "no");

  /* verify the multi-GPU proof with the single-GPU verifier */
  int accepted = 0;
  lg_sponge* vs = NULL;
  CHECK(lg_sponge_test(&vs), "");
  CHECK(lg_verify(one, proof2, vs, &accepted), lg_last_error(ctx));
  printf("verify: %s\n", accepted ? "accepted" : "REJECTED");

  lg_sponge_free(vs);
  lg_sponge_free(sponge);
  lg_proof_free(proof1);
  lg_proof_free(proof2);
  free(b1);
  free(b2);
  lg_mgpu_ligero_free(many);
  lg_mgpu_destroy(box);
  lg_ligero_free(one);
  lg_ctx_destroy(ctx);
  lg_circuit_free(circuit);
  puts(same && accepted ? "C_MGPU_OK" : "C_MGPU_FAIL");
  return same && accepted ? 0 : 1;
}
