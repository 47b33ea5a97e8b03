// Round-2 pipe probe: is IMAD.WIDE.U32 slow (4 cycles per warp) only in its carry-chained form (.X)?
//   A  mad.wide.u32 into independent 64-bit accumulators, 8 x 8 distinct loop-invariant operand registers
//   B  the same with the multiplicands changing every iteration (ALU xor), so no operand is reusable
//   C  carry-chained wide products as the Shoup body issues them: mad.lo.cc / madc.hi.cc / madc.lo.cc / madc.hi
//   D  9 x 9 schoolbook product of 29-bit limbs, column sums in 64-bit accumulators with NO carries (81 mad.wide) + carry sweep
//   E  8 x 8 schoolbook product of 32-bit limbs with even/odd carry chains (64 wide multiplies, the CIOS-style rows)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/pipe_probe3 scripts/pipe_probe3.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed) {
  uint32_t a[8], b[8];
  uint64_t w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + i + 1) | 1u; b[i] = seed * (2 * threadIdx.x + i + 3) | 1u; w[i] = threadIdx.x * 7 + i; }
  for (int it = 0; it < ITER; it++) {
    if (MODE == 0 || MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[(i + j) & 7]));
        if (MODE == 1) {
#pragma unroll
          for (int i = 0; i < 8; i++) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"((uint32_t)w[(i + 1) & 7]));
        }
      }
    }
    if (MODE == 2) {  // 64 wide multiplies in 32 four-instruction carry chains
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          uint32_t lo0 = (uint32_t)w[i], hi0 = (uint32_t)(w[i] >> 32), lo1 = (uint32_t)w[i + 1], hi1 = (uint32_t)(w[i + 1] >> 32);
          asm volatile("mad.lo.cc.u32 %0, %4, %5, %0; madc.hi.cc.u32 %1, %4, %5, %1; madc.lo.cc.u32 %2, %4, %6, %2; madc.hi.u32 %3, %4, %6, %3;"
                       : "+r"(lo0), "+r"(hi0), "+r"(lo1), "+r"(hi1)
                       : "r"(a[j]), "r"(b[i]), "r"(b[i + 1]));
          w[i] = ((uint64_t)hi0 << 32) | lo0;
          w[i + 1] = ((uint64_t)hi1 << 32) | lo1;
        }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += (uint32_t)w[i] + (uint32_t)(w[i] >> 32) + a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// D: x, y in 9 limbs of 29 bits -> 18 limbs of 29 bits (full product), repeated as a chain z = x*y, x = low 9 limbs of z
__global__ void __launch_bounds__(256) probe_mul29(uint32_t* out, uint32_t seed, int iters) {
  uint32_t x[9], y[9];
#pragma unroll
  for (int i = 0; i < 9; i++) { x[i] = (seed * (threadIdx.x + i + 1)) & 0x1fffffffu; y[i] = (seed * (3 * threadIdx.x + i + 5)) & 0x1fffffffu; }
  for (int it = 0; it < iters; it++) {
    uint64_t c[17];
#pragma unroll
    for (int k = 0; k < 17; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
      for (int j = 0; j < 9; j++) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[i + j]) : "r"(x[i]), "r"(y[j]));
    // carry sweep: limb k = c[k] mod 2^29, carry into c[k+1]
    uint64_t carry = 0;
    uint32_t z[18];
#pragma unroll
    for (int k = 0; k < 17; k++) {
      const uint64_t v = c[k] + carry;
      z[k] = (uint32_t)v & 0x1fffffffu;
      carry = v >> 29;
    }
    z[17] = (uint32_t)carry;
#pragma unroll
    for (int i = 0; i < 9; i++) x[i] = z[i] ^ z[i + 9];
#pragma unroll
    for (int i = 0; i < 9; i++) x[i] &= 0x1fffffffu;
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// E: 8 x 8 limbs of 32 bits, row-wise with even/odd accumulators and carry chains (64 wide multiplies)
__device__ __forceinline__ void mul4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
  asm("mul.lo.u32 %0, %8, %12; mul.hi.u32 %1, %8, %12; mul.lo.u32 %2, %9, %12; mul.hi.u32 %3, %9, %12;"
      "mul.lo.u32 %4, %10, %12; mul.hi.u32 %5, %10, %12; mul.lo.u32 %6, %11, %12; mul.hi.u32 %7, %11, %12;"
      : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
__device__ __forceinline__ void mad4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1; madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;"
      "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5; madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
__global__ void __launch_bounds__(256) probe_mul32(uint32_t* out, uint32_t seed, int iters) {
  uint32_t x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = seed * (threadIdx.x + i + 1); y[i] = seed * (3 * threadIdx.x + i + 5); }
  for (int it = 0; it < iters; it++) {
    // even[k] holds columns (2k, 2k+1) of products x_even * y_j shifted ..., a simplified operand-scanning product:
    uint32_t ev[16], od[16];
#pragma unroll
    for (int k = 0; k < 16; k++) ev[k] = od[k] = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      // x_even * y_j accumulates at column j (even alignment when j even), x_odd * y_j at column j+1
      uint32_t* e = (j & 1) ? od : ev;
      uint32_t* o = (j & 1) ? ev : od;
      mad4(e + (j & ~1), x[0], x[2], x[4], x[6], y[j]);
      mad4(o + ((j + 1) & ~1), x[1], x[3], x[5], x[7], y[j]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = ev[i] ^ od[i] ^ ev[i + 8] ^ od[i + 8];
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int ctas = sms * 8;
  uint32_t* out;
  cudaMalloc(&out, (size_t)ctas * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[5] = {"A mad.wide.u32, 64 independent products / iteration, invariant distinct operands",
                          "B mad.wide.u32, multiplicands rewritten every 8 products (+8 xor)",
                          "C carry-chained wide products (mad.lo.cc / madc.hi.cc / madc.lo.cc / madc.hi), 64 / iteration",
                          "D 9x9 product of 29-bit limbs, carry-free column sums (81 mad.wide + sweep)",
                          "E 8x8 product of 32-bit limbs, even/odd carry chains (64 wide multiplies)"};
  for (int mode = 0; mode < 5; mode++) {
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
      cudaEventRecord(e0);
      if (mode == 0) probe<0><<<ctas, 256>>>(out, 12345u);
      if (mode == 1) probe<1><<<ctas, 256>>>(out, 12345u);
      if (mode == 2) probe<2><<<ctas, 256>>>(out, 12345u);
      if (mode == 3) probe_mul29<<<ctas, 256>>>(out, 12345u, ITER);
      if (mode == 4) probe_mul32<<<ctas, 256>>>(out, 12345u, ITER);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (r && ms < best) best = ms;
    }
    const double threads = (double)ctas * 256;
    const double per_iter = mode == 3 ? 81 : 64;
    const double wide_per_clk_sm = threads * ITER * per_iter / (best * 1e-3) / ((double)khz * 1e3) / sms;
    const double cycles_per_warp_iter = (double)khz * 1e3 * (best * 1e-3) / ((double)ITER * 16.0);  // 16 warps / sub-partition
    printf("{\"probe\": \"%s\", \"ms\": %.3f, \"wide_lanes_per_clk_per_sm\": %.2f, \"cycles_per_warp_iteration\": %.1f, \"clock_khz\": %d}\n",
           names[mode], best, wide_per_clk_sm, cycles_per_warp_iter, khz);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("cuda error: %s\n", cudaGetErrorString(e));
  return 0;
}
