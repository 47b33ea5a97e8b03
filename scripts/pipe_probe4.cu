// Round-2 co-issue probe: can another pipe make progress while an SM sub-partition streams carry-chained IMAD.WIDE.U32?
// (a second multiplier port on the FP64 pipe only helps if both ports run at once; section 4: a
// column hash -- ALU-pipe work -- sharing an SM with the encoder.)  Per iteration and thread:
//   W    64 wide multiplies in 32 carry chains (the multiplier's own instruction mix; probe C of pipe_probe3)
//   F    64 DFMA on 8 dependent accumulators (fma.rz.f64)
//   A    128 ALU-pipe instructions (xor / add on 8 registers)
//   W+F  both, interleaved 1:1;   W+F/2  64 wide + 32 DFMA;   W+A  64 wide + 128 ALU
// If W+F takes max(W, F) the FP64 pipe is a free second port; if it takes W + F the issue port is the bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/pipe_probe4 scripts/pipe_probe4.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048

template <int NW, int NF, int NA>  // wide groups of 8, DFMA groups of 8, ALU groups of 16 per 8-step iteration
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed, double dseed) {
  uint32_t a[8], b[8], lo[8], hi[8], x[8];
  double d[8];
  const double e = dseed, f = dseed * 0.5;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = seed * (threadIdx.x + i + 1) | 1u;
    b[i] = seed * (2 * threadIdx.x + i + 3) | 1u;
    lo[i] = threadIdx.x + i;
    hi[i] = threadIdx.x * 3 + i;
    x[i] = seed + i * threadIdx.x;
    d[i] = dseed + i;
  }
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (j < NW) {
#pragma unroll
        for (int i = 0; i < 8; i += 2)
          asm volatile("mad.lo.cc.u32 %0, %4, %5, %0; madc.hi.cc.u32 %1, %4, %5, %1; madc.lo.cc.u32 %2, %4, %6, %2; madc.hi.u32 %3, %4, %6, %3;"
                       : "+r"(lo[i]), "+r"(hi[i]), "+r"(lo[i + 1]), "+r"(hi[i + 1])
                       : "r"(a[j]), "r"(b[i]), "r"(b[i + 1]));
      }
      if (j < NF) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e), "d"(f));
      }
      if (j < NA) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(seed));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]));
        }
      }
    }
    if (NW) {
#pragma unroll
      for (int i = 0; i < 8; i++) {  // keep the multiplicands loop-variant (ptxas hoists invariant products)
        a[i] ^= hi[i];
        b[i] += lo[(i + 3) & 7];
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += lo[i] + hi[i] + a[i] + b[i] + x[i] + (uint32_t)__double_as_longlong(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NW, int NF, int NA>
static double run(const char* name, uint32_t* out, int ctas, int khz) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0);
    probe<NW, NF, NA><<<ctas, 256>>>(out, 12345u, 1.0000001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  const double cyc = (double)khz * 1e3 * (best * 1e-3) / ((double)ITER * 16.0);  // 16 warps per sub-partition
  printf("{\"probe\": \"%s\", \"wide\": %d, \"dfma\": %d, \"alu\": %d, \"ms\": %.3f, \"cycles_per_warp_iteration\": %.1f}\n", name, NW * 8, NF * 8,
         NA * 16, best, cyc);
  return cyc;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int ctas = sms * 8;
  uint32_t* out;
  cudaMalloc(&out, (size_t)ctas * 256 * 4);
  run<8, 0, 0>("W: 64 carry-chained wide multiplies", out, ctas, khz);
  run<0, 8, 0>("F: 64 DFMA", out, ctas, khz);
  run<0, 0, 8>("A: 128 ALU", out, ctas, khz);
  run<8, 8, 0>("W+F: 64 wide + 64 DFMA", out, ctas, khz);
  run<8, 4, 0>("W+F/2: 64 wide + 32 DFMA", out, ctas, khz);
  run<8, 0, 8>("W+A: 64 wide + 128 ALU", out, ctas, khz);
  run<8, 0, 4>("W+A/2: 64 wide + 64 ALU", out, ctas, khz);
  run<8, 4, 4>("W+F/2+A/2: 64 wide + 32 DFMA + 64 ALU", out, ctas, khz);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("cuda error: %s\n", cudaGetErrorString(e));
  return 0;
}
