// Pipe-rate probe for sm_100a: which integer / FP64 instructions run at what rate, and which co-issue.
// Used to choose the Fr multiplier (IMAD.WIDE CIOS vs fixed-multiplicand Barrett vs DFMA limbs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define UNR 8

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed, double dseed) {
  uint32_t a[UNR], b = seed | 1u, c = seed * 3u + 7u;
  uint64_t w[UNR];
  double d[UNR], e = dseed, f = dseed * 0.5;
#pragma unroll
  for (int i = 0; i < UNR; i++) { a[i] = threadIdx.x + i; w[i] = threadIdx.x * 7 + i; d[i] = dseed + i; }
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < UNR; i++) {
      if (MODE == 0 || MODE == 5 || MODE == 7) {  // IMAD.WIDE.U32
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
      }
      if (MODE == 1) {  // IMAD (lo)
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      }
      if (MODE == 2) {  // IMAD.HI
        asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      }
      if (MODE == 3 || MODE == 5 || MODE == 6) {  // DFMA
        asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e), "d"(f));
      }
      if (MODE == 4) {  // DADD
        asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e));
      }
      if (MODE == 6 || MODE == 7 || MODE == 8) {  // IADD3 / LOP3 (alu pipe)
        asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
      }
      if (MODE == 9) {  // wide carry chain pair as used by the CIOS multiplier
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(a[(i + 1) % UNR]) : "r"(b), "r"(c));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < UNR; i++) s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32) + (uint32_t)__double_as_longlong(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double ops_per_iter) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  uint32_t* out;
  const int ctas = sms * 8;
  cudaMalloc(&out, (size_t)ctas * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<ctas, 256>>>(out, 12345u, 1.0000001);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    probe<MODE><<<ctas, 256>>>(out, 12345u, 1.0000001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double lane_ops = (double)ctas * 256 * ITER * UNR * ops_per_iter;
  const double per_clk_sm = lane_ops / (best * 1e-3) / ((double)khz * 1e3) / sms;
  printf("{\"probe\": \"%s\", \"ms\": %.3f, \"lane_ops_per_clk_per_sm\": %.2f, \"ops_counted_per_slot\": %.0f, \"clock_khz\": %d, \"sms\": %d}\n", name, best, per_clk_sm, ops_per_iter, khz, sms);
  cudaFree(out);
}

int main() {
  run<0>("imad_wide_u32", 1);
  run<1>("imad_lo_u32", 1);
  run<2>("imad_hi_u32", 1);
  run<9>("mad_lo_cc+madc_hi pair", 2);
  run<3>("dfma_rz", 1);
  run<4>("dadd_rz", 1);
  run<8>("iadd+lop (alu)", 2);
  run<5>("imad_wide + dfma (count 2)", 2);
  run<6>("dfma + 2 alu (count 3)", 3);
  run<7>("imad_wide + 2 alu (count 3)", 3);
  return 0;
}
