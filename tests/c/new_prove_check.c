/* A C host timing LigeroCircuit::new and checking what it builds, through include/ligero_b200.h only: a seeded synthetic
 * circuit large enough for the device constraint-matrix builder and the device trace, proved once with the trace evaluated
 * on the device (level schedule of lg_ligero_new) and once with the host evaluator (formatted node copy + slot map), the
 * two proofs compared byte for byte, then verified.
 *
 *   gcc -std=c11 -O2 -I include tests/c/new_prove_check.c -o /tmp/new_prove_check -L ligero_b200 -lligero_b200 -Wl,-rpath,$PWD/ligero_b200
 *   /tmp/new_prove_check [log2_gates]        -> prints C_NEW_OK
 * tests/test_abi.py compiles it without a GPU; tests/test_gpu_host_driver.py runs it. */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ligero_b200.h"

#define CHECK(expr, who)                                                              \
  do {                                                                                \
    int st_ = (expr);                                                                 \
    if (st_ != LG_OK) {                                                               \
      fprintf(stderr, "%s failed: %d (%s)\n", #expr, st_, (who) ? (who) : "");        \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int serialize(const lg_proof* p, uint8_t** buf, size_t* len) {
  if (lg_proof_serialize(p, NULL, 0, len) != LG_OK) return 1;
  *buf = (uint8_t*)malloc(*len);
  return *buf == NULL || lg_proof_serialize(p, *buf, *len, len) != LG_OK;
}

int main(int argc, char** argv) {
  const int log_gates = argc > 1 ? atoi(argv[1]) : 17;
  lg_circuit* circuit = NULL;
  size_t output = 0, var_idx[2];
  uint64_t var_vals[8];
  double t0 = now_ms();
  CHECK(lg_circuit_synthetic((size_t)1 << log_gates, 2024, &circuit, &output, var_idx, var_vals), "");
  const double t_circuit = now_ms() - t0;
  lg_ctx* ctx = NULL;
  CHECK(lg_ctx_create(0, &ctx), "no usable GPU");
  lg_ligero* L = NULL;
  t0 = now_ms();
  CHECK(lg_ligero_new(ctx, circuit, &output, 1, 128, &L), lg_last_error(ctx));
  CHECK(lg_ctx_sync(ctx), lg_last_error(ctx));
  const double t_new = now_ms() - t0;
  size_t gates = 0, levels = 0, launches = 0;
  int on_device = 0;
  CHECK(lg_ligero_trace_info(L, &gates, &levels, &launches, &on_device), "");
  printf("2^%d gates: circuit %.1f ms, LigeroCircuit::new %.1f ms (%zu gates in %zu levels, device trace by default: %d)\n", log_gates,
         t_circuit, t_new, gates, levels, on_device);

  lg_proof* proof[2] = {NULL, NULL};
  uint8_t* blob[2] = {NULL, NULL};
  size_t len[2] = {0, 0};
  for (int mode = 1; mode >= 0; mode--) { /* 1: trace on the device, 0: host evaluator */
    lg_sponge* sponge = NULL;
    CHECK(lg_sponge_test(&sponge), "");
    CHECK(lg_ligero_set_trace_mode(L, mode), "");
    t0 = now_ms();
    CHECK(lg_prove(L, var_idx, var_vals, 2, 1, sponge, &proof[mode]), lg_last_error(ctx));
    printf("prove (trace on the %s): %.1f ms\n", mode ? "device" : "host", now_ms() - t0);
    lg_sponge_free(sponge);
    if (serialize(proof[mode], &blob[mode], &len[mode])) return 1;
  }
  const int same = len[0] == len[1] && memcmp(blob[0], blob[1], len[0]) == 0;
  printf("proof %zu bytes; device-trace proof %s host-trace proof\n", len[1], same ? "==" : "!=");
  int accepted = 0;
  lg_sponge* vs = NULL;
  CHECK(lg_sponge_test(&vs), "");
  t0 = now_ms();
  CHECK(lg_verify(L, proof[1], vs, &accepted), lg_last_error(ctx));
  printf("verify: %s in %.1f ms\n", accepted ? "accepted" : "REJECTED", now_ms() - t0);
  lg_sponge_free(vs);
  for (int i = 0; i < 2; i++) {
    lg_proof_free(proof[i]);
    free(blob[i]);
  }
  lg_ligero_free(L);
  lg_ctx_destroy(ctx);
  lg_circuit_free(circuit);
  puts(same && accepted ? "C_NEW_OK" : "C_NEW_FAIL");
  return same && accepted ? 0 : 1;
}
