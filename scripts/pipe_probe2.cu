// Issue-model probe for sm_100a: does a half/quarter-rate integer multiply block the dispatch port, i.e. do
// IMAD.WIDE and plain ALU instructions add up (exclusive) or overlap (separate pipes)?  And does DFMA overlap both?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe2 pipe_probe2.cu && ./pipe_probe2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define UNR 8

// NW wide multiplies, NA alu ops, ND dfma per slot
template <int NW, int NA, int ND, int NL>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed, double dseed) {
  uint64_t w[UNR];
  uint32_t a[UNR], l[UNR];
  double d[UNR], e = dseed, f = dseed * 0.5;
  const uint32_t m = (threadIdx.x | 1u) * seed;
#pragma unroll
  for (int i = 0; i < UNR; i++) { w[i] = threadIdx.x * 7 + i + seed; a[i] = threadIdx.x + i; d[i] = dseed + i; l[i] = seed + i * 3; }
  uint32_t xv = seed ^ threadIdx.x;
  for (int it = 0; it < ITER; it++) {
    xv += 0x9e3779b9u;
#pragma unroll
    for (int i = 0; i < UNR; i++) {
#pragma unroll
      for (int r = 0; r < NW; r++) {  // the carry-chain pair of the multipliers: ptxas fuses it into IMAD.WIDE.U32(.X)
        uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32);
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"((uint32_t)(w[(i + 1) % UNR] >> 32) ^ xv), "r"(m));
        w[i] = ((uint64_t)hi << 32) | lo;
      }
#pragma unroll
      for (int r = 0; r < NL; r++)
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(l[i]) : "r"(l[(i + 1) % UNR]), "r"(m));
#pragma unroll
      for (int r = 0; r < NA; r++) {
        if (r & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 3) % UNR]));
        else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 5) % UNR]));
      }
#pragma unroll
      for (int r = 0; r < ND; r++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e), "d"(f));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < UNR; i++) s += a[i] + l[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32) + (uint32_t)__double_as_longlong(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NW, int NA, int ND, int NL>
static void run(const char* name) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  uint32_t* out;
  const int ctas = sms * 8;
  cudaMalloc(&out, (size_t)ctas * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<NW, NA, ND, NL><<<ctas, 256>>>(out, 12345u, 1.0000001);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    probe<NW, NA, ND, NL><<<ctas, 256>>>(out, 12345u, 1.0000001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  // SMSP cycles per slot (one slot = NW wide + NL lo + NA alu + ND dfma warp instructions)
  const double slots_per_smsp = (double)ctas * 256 / 32 * ITER * UNR / (sms * 4.0);
  const double cyc = best * 1e-3 * khz * 1e3 / slots_per_smsp;
  printf("{\"probe\": \"%s\", \"wide\": %d, \"lo\": %d, \"alu\": %d, \"dfma\": %d, \"ms\": %.3f, \"smsp_cycles_per_slot\": %.2f}\n", name, NW, NL, NA, ND, best, cyc);
  cudaFree(out);
}

int main() {
  run<1, 0, 0, 0>("wide");
  run<0, 0, 0, 1>("lo");
  run<0, 1, 0, 0>("alu1");
  run<0, 4, 0, 0>("alu4");
  run<0, 0, 1, 0>("dfma");
  run<1, 1, 0, 0>("wide+1alu");
  run<1, 2, 0, 0>("wide+2alu");
  run<1, 3, 0, 0>("wide+3alu");
  run<1, 4, 0, 0>("wide+4alu");
  run<0, 1, 0, 1>("lo+1alu");
  run<0, 2, 0, 1>("lo+2alu");
  run<1, 0, 1, 0>("wide+dfma");
  run<1, 0, 2, 0>("wide+2dfma");
  run<1, 2, 1, 0>("wide+2alu+dfma");
  run<1, 2, 2, 0>("wide+2alu+2dfma");
  run<0, 2, 1, 0>("2alu+dfma");
  run<0, 4, 1, 0>("4alu+dfma");
  run<0, 4, 2, 0>("4alu+2dfma");
  run<1, 0, 0, 1>("wide+lo");
  return 0;
}
