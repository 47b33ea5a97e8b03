// Lazy-reduction arithmetic for the NTT butterflies on sm_100a.
//
// Every Fr multiplication inside the row transforms has one operand from a table (twiddle, coset scale,
// 1/k).  For such a constant w the table also carries p = floor(w * 2^256 / r) and the product is
//     T = y*w - qhat*r,   qhat = floor( trunc(y*p) / 2^256 )  in {Q-1, Q},  Q = floor(y*w / r)
// (Shoup / Barrett with a per-constant quotient multiplier).  T is in [0, 2r) for every y < 2^256(1-2^-29):
// 43 + 28 + 28 wide and 16 low 32x32 MACs instead of the 136 wide MACs of a Montgomery product -- the
// IMAD.WIDE pipe (32 lanes/clk/SM) is the unit that bounds the encoder.
// w is the plain integer of the twiddle, so  y*w mod r  keeps whatever form y is in (the data stay in the
// reference's Montgomery form R = 2^256; nothing is converted).
//
// Butterflies keep values only partially reduced (Harvey's lazy butterflies; 4r < 0.76 * 2^256):
//     forward (DIT):  X, Y < 4r + d :  X~ = X - [X7 > top(2r)] 2r  (< 2r + d, d <= 2^224: ONE compare + 8 predicated subs)
//                     T = Y*w in [0,2r);   X' = X~ + T;   Y' = X~ + (2r - T)        (both < 4r + d)
//     inverse (DIF):  X, Y < 2r + d :  X' = csub2r(X + Y);   Y' = (X + (3r - Y)) * w     (d at most doubles per stage)
// and fr_normalize brings a finished value back to the canonical representative in [0, r), so the stored
// codewords are bit-identical to fully reduced arithmetic.
#pragma once
#include "fr.cuh"

namespace lg {

// table entry: the constant and its quotient multiplier (64 bytes)
struct alignas(16) FrTw {
  Fr w;   // plain integer < r (NOT Montgomery form)
  Fr p;   // floor(w * 2^256 / r)
};

// 2r, 3r little-endian limbs
#define LG_2R0 0xe0000002u
#define LG_2R1 0x87c3eb27u
#define LG_2R2 0xf372e122u
#define LG_2R3 0x5067d090u
#define LG_2R4 0x0302b0bau
#define LG_2R5 0x70a08b6du
#define LG_2R6 0xc2634053u
#define LG_2R7 0x60c89ce5u
#define LG_3R0 0xd0000003u
#define LG_3R1 0xcba5e0bbu
#define LG_3R2 0x6d2c51b3u
#define LG_3R3 0x789bb8d9u
#define LG_3R4 0x84840917u
#define LG_3R5 0x28f0d123u
#define LG_3R6 0xa394e07du
#define LG_3R7 0x912ceb58u

// p = floor(w * 2^256 / r) for w < r: 256 steps of restoring division (table generation only)
LG_HD Fr fr_shoup_quotient(const Fr& w) {
  uint32_t rem[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { rem[i] = w.v[i]; q[i] = 0; }
  for (int bit = 0; bit < 256; bit++) {
    // rem <- 2 rem (rem < r < 2^254: no overflow), q <- 2 q
#pragma unroll
    for (int i = 7; i > 0; i--) {
      rem[i] = (rem[i] << 1) | (rem[i - 1] >> 31);
      q[i] = (q[i] << 1) | (q[i - 1] >> 31);
    }
    rem[0] <<= 1;
    q[0] <<= 1;
    uint32_t t[8], br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint64_t d = (uint64_t)rem[i] - fr_p(i) - br;
      t[i] = (uint32_t)d;
      br = (uint32_t)(d >> 63);
    }
    if (!br) {
#pragma unroll
      for (int i = 0; i < 8; i++) rem[i] = t[i];
      q[0] |= 1u;
    }
  }
  Fr o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = q[i];
  return o;
}

// y*w - qhat*r in [0, 2r)
LG_HD Fr fr_mul_shoup(const Fr& y, const FrTw& tw) {
  const uint32_t y0 = y.v[0], y1 = y.v[1], y2 = y.v[2], y3 = y.v[3], y4 = y.v[4], y5 = y.v[5], y6 = y.v[6], y7 = y.v[7];
  const uint32_t w0 = tw.w.v[0], w1 = tw.w.v[1], w2 = tw.w.v[2], w3 = tw.w.v[3], w4 = tw.w.v[4], w5 = tw.w.v[5],
                 w6 = tw.w.v[6], w7 = tw.w.v[7];
  const uint32_t p0 = tw.p.v[0], p1 = tw.p.v[1], p2 = tw.p.v[2], p3 = tw.p.v[3], p4 = tw.p.v[4], p5 = tw.p.v[5],
                 p6 = tw.p.v[6], p7 = tw.p.v[7];
#include "fr_shoup_body_q.inc"
#include "fr_shoup_body_t.inc"
  (void)hx; (void)he6;
  Fr t;
  t.v[0] = e0; t.v[1] = e1; t.v[2] = e2; t.v[3] = e3; t.v[4] = e4; t.v[5] = e5; t.v[6] = e6; t.v[7] = e7;
  return t;
}

#ifdef __CUDACC__
// The same product with the table entry still in memory: the quotient multiplier p is loaded first, the
// constant w only once the quotient is done, so a thread never holds both halves (8 registers less at the
// peak of the radix-4 shared-memory passes, which run at 80 registers per thread).
__device__ __forceinline__ Fr fr_mul_shoup_ld(const Fr& y, const FrTw* __restrict__ tw) {
  const uint32_t y0 = y.v[0], y1 = y.v[1], y2 = y.v[2], y3 = y.v[3], y4 = y.v[4], y5 = y.v[5], y6 = y.v[6], y7 = y.v[7];
  const uint4* tp = reinterpret_cast<const uint4*>(tw);
  uint32_t p0, p1, p2, p3, p4, p5, p6, p7;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p0), "=r"(p1), "=r"(p2), "=r"(p3) : "l"(tp + 2));
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p4), "=r"(p5), "=r"(p6), "=r"(p7) : "l"(tp + 3));
#include "fr_shoup_body_q.inc"
  uint32_t w0, w1, w2, w3, w4, w5, w6, w7;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "l"(tp));
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7) : "l"(tp + 1));
#include "fr_shoup_body_t.inc"
  (void)hx; (void)he6;
  Fr t;
  t.v[0] = e0; t.v[1] = e1; t.v[2] = e2; t.v[3] = e3; t.v[4] = e4; t.v[5] = e5; t.v[6] = e6; t.v[7] = e7;
  return t;
}
#endif

#ifdef __CUDACC__
// The same product with the table entry split into four 16-byte pieces (structure of arrays): consecutive lanes
// reading consecutive entries touch consecutive 16-byte slots, so a warp's read of one piece is four full wavefronts
// of shared memory (or four full sectors of global memory) instead of 32 scattered 64-byte entries.
//   shared-memory twiddles (persistent encoder): 32-bit shared addresses of {w.lo, w.hi, p.lo, p.hi}
__device__ __forceinline__ Fr fr_mul_shoup_sm(const Fr& y, uint32_t a_wlo, uint32_t a_whi, uint32_t a_plo, uint32_t a_phi) {
  const uint32_t y0 = y.v[0], y1 = y.v[1], y2 = y.v[2], y3 = y.v[3], y4 = y.v[4], y5 = y.v[5], y6 = y.v[6], y7 = y.v[7];
  uint32_t p0, p1, p2, p3, p4, p5, p6, p7;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p0), "=r"(p1), "=r"(p2), "=r"(p3) : "r"(a_plo));
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p4), "=r"(p5), "=r"(p6), "=r"(p7) : "r"(a_phi));
#include "fr_shoup_body_q.inc"
  uint32_t w0, w1, w2, w3, w4, w5, w6, w7;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(a_wlo));
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7) : "r"(a_whi));
#include "fr_shoup_body_t.inc"
  (void)hx; (void)he6;
  Fr t;
  t.v[0] = e0; t.v[1] = e1; t.v[2] = e2; t.v[3] = e3; t.v[4] = e4; t.v[5] = e5; t.v[6] = e6; t.v[7] = e7;
  return t;
}
//   global SoA table (coset scale factors): piece pointers `stride` uint4 apart in the order w.lo, w.hi, p.lo, p.hi
__device__ __forceinline__ Fr fr_mul_shoup_g4(const Fr& y, const uint4* __restrict__ tp, uint32_t stride) {
  const uint32_t y0 = y.v[0], y1 = y.v[1], y2 = y.v[2], y3 = y.v[3], y4 = y.v[4], y5 = y.v[5], y6 = y.v[6], y7 = y.v[7];
  uint32_t p0, p1, p2, p3, p4, p5, p6, p7;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p0), "=r"(p1), "=r"(p2), "=r"(p3) : "l"(tp + 2 * stride));
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p4), "=r"(p5), "=r"(p6), "=r"(p7) : "l"(tp + 3 * stride));
#include "fr_shoup_body_q.inc"
  uint32_t w0, w1, w2, w3, w4, w5, w6, w7;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "l"(tp));
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7) : "l"(tp + stride));
#include "fr_shoup_body_t.inc"
  (void)hx; (void)he6;
  Fr t;
  t.v[0] = e0; t.v[1] = e1; t.v[2] = e2; t.v[3] = e3; t.v[4] = e4; t.v[5] = e5; t.v[6] = e6; t.v[7] = e7;
  return t;
}
#endif

// plain 256-bit add (callers guarantee no overflow)
LG_HD Fr lz_add(const Fr& a, const Fr& b) {
  Fr r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, %17; addc.cc.u32 %2, %10, %18; addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20; addc.cc.u32 %5, %13, %21; addc.cc.u32 %6, %14, %22; addc.u32 %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
#else
  uint32_t c = 0;
  for (int i = 0; i < 8; i++) r.v[i] = emu::addc(a.v[i], b.v[i], c);
#endif
  return r;
}

// c - b for a constant c given as 8 limbs (callers guarantee b <= c)
LG_HD Fr lz_const_minus(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4, uint32_t c5, uint32_t c6,
                        uint32_t c7, const Fr& b) {
  Fr r;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %8, %16; subc.cc.u32 %1, %9, %17; subc.cc.u32 %2, %10, %18; subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20; subc.cc.u32 %5, %13, %21; subc.cc.u32 %6, %14, %22; subc.u32 %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7])
      : "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(c5), "r"(c6), "r"(c7), "r"(b.v[0]), "r"(b.v[1]),
        "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
#else
  const uint32_t c[8] = {c0, c1, c2, c3, c4, c5, c6, c7};
  uint32_t br = 0;
  for (int i = 0; i < 8; i++) r.v[i] = emu::subb(c[i], b.v[i], br);
#endif
  return r;
}
LG_HD Fr lz_2r_minus(const Fr& b) {  // b <= 2r
  return lz_const_minus(LG_2R0, LG_2R1, LG_2R2, LG_2R3, LG_2R4, LG_2R5, LG_2R6, LG_2R7, b);
}
LG_HD Fr lz_3r_minus(const Fr& b) {  // b <= 3r
  return lz_const_minus(LG_3R0, LG_3R1, LG_3R2, LG_3R3, LG_3R4, LG_3R5, LG_3R6, LG_3R7, b);
}

// x -= 2r when the top limb alone proves x > 2r.  x < 4r + d  ->  x < 2r + max(d, 2^224)
LG_HD void lz_csub2r(Fr& x) {
#ifdef __CUDA_ARCH__
  asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %7, %8;\n\t"
      "@p sub.cc.u32 %0, %0, %9;\n\t@p subc.cc.u32 %1, %1, %10;\n\t@p subc.cc.u32 %2, %2, %11;\n\t"
      "@p subc.cc.u32 %3, %3, %12;\n\t@p subc.cc.u32 %4, %4, %13;\n\t@p subc.cc.u32 %5, %5, %14;\n\t"
      "@p subc.cc.u32 %6, %6, %15;\n\t@p subc.u32 %7, %7, %8;\n\t}"
      : "+r"(x.v[0]), "+r"(x.v[1]), "+r"(x.v[2]), "+r"(x.v[3]), "+r"(x.v[4]), "+r"(x.v[5]), "+r"(x.v[6]),
        "+r"(x.v[7])
      : "r"(LG_2R7), "r"(LG_2R0), "r"(LG_2R1), "r"(LG_2R2), "r"(LG_2R3), "r"(LG_2R4), "r"(LG_2R5), "r"(LG_2R6));
#else
  if (x.v[7] > LG_2R7) {
    const uint32_t c[8] = {LG_2R0, LG_2R1, LG_2R2, LG_2R3, LG_2R4, LG_2R5, LG_2R6, LG_2R7};
    uint32_t br = 0;
    for (int i = 0; i < 8; i++) x.v[i] = emu::subb(x.v[i], c[i], br);
  }
#endif
}

// exact conditional subtraction of the constant c: x >= c ? x - c : x
LG_HD void lz_csub_exact(uint32_t x[8], uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4, uint32_t c5,
                         uint32_t c6, uint32_t c7) {
  uint32_t t[8], borrow;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22; subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
        "=r"(borrow)
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "r"(c4), "r"(c5), "r"(c6), "r"(c7));
#else
  const uint32_t c[8] = {c0, c1, c2, c3, c4, c5, c6, c7};
  uint32_t br = 0;
  for (int i = 0; i < 8; i++) t[i] = emu::subb(x[i], c[i], br);
  borrow = br ? 0xffffffffu : 0u;
#endif
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = borrow ? x[i] : t[i];
}

// canonical representative in [0, r) of a lazily reduced value x < 4r + r
LG_HD Fr fr_normalize(Fr x) {
  lz_csub_exact(x.v, LG_2R0, LG_2R1, LG_2R2, LG_2R3, LG_2R4, LG_2R5, LG_2R6, LG_2R7);   // < 3r
  lz_csub_exact(x.v, LG_2R0, LG_2R1, LG_2R2, LG_2R3, LG_2R4, LG_2R5, LG_2R6, LG_2R7);   // < 2r
  fr_final_sub(x.v);                                                                     // < r
  return x;
}
// the same for x < 2r (a Shoup product)
LG_HD Fr fr_normalize_2r(Fr x) {
  fr_final_sub(x.v);
  return x;
}

// forward (decimation in time) butterfly:  (X, Y) -> (X + w Y, X - w Y), values < 4r + 2^224 in and out
LG_HD void lz_bfly_dit(Fr& X, Fr& Y, const FrTw& tw) {
  lz_csub2r(X);
  const Fr T = fr_mul_shoup(Y, tw);
  Y = lz_add(X, lz_2r_minus(T));
  X = lz_add(X, T);
}
#ifdef __CUDACC__
__device__ __forceinline__ void lz_bfly_dit_ld(Fr& X, Fr& Y, const FrTw* __restrict__ tw) {
  lz_csub2r(X);
  const Fr T = fr_mul_shoup_ld(Y, tw);
  Y = lz_add(X, lz_2r_minus(T));
  X = lz_add(X, T);
}
__device__ __forceinline__ void lz_bfly_dif_ld(Fr& X, Fr& Y, const FrTw* __restrict__ tw) {
  const Fr D = lz_add(X, lz_3r_minus(Y));
  X = lz_add(X, Y);
  lz_csub2r(X);
  Y = fr_mul_shoup_ld(D, tw);
}
#endif
// forward butterfly with w = 1
LG_HD void lz_bfly_dit1(Fr& X, Fr& Y) {
  lz_csub2r(X);
  lz_csub2r(Y);
  const Fr T = Y;                      // < 2r + 2^224
  Y = lz_add(X, lz_3r_minus(T));       // X + 3r - T: stays >= 0 although T may exceed 2r by < 2^224
  lz_csub2r(Y);                        // < 5r + d  ->  < 3r + d
  X = lz_add(X, T);                    // < 4r + 2d
}
// inverse (decimation in frequency) butterfly: (X, Y) -> (X + Y, (X - Y) w), values < 2r + 2^240 in and out
LG_HD void lz_bfly_dif(Fr& X, Fr& Y, const FrTw& tw) {
  const Fr D = lz_add(X, lz_3r_minus(Y));
  X = lz_add(X, Y);
  lz_csub2r(X);
  Y = fr_mul_shoup(D, tw);
}
LG_HD void lz_bfly_dif1(Fr& X, Fr& Y) {
  Fr D = lz_add(X, lz_3r_minus(Y));    // < 5r + d
  X = lz_add(X, Y);
  lz_csub2r(X);
  lz_csub2r(D);                        // < 3r + d
  lz_csub2r(D);                        // < 2r + 2^224
  Y = D;
}

}  // namespace lg
