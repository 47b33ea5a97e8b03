// BN254 scalar field Fr for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256), little endian.
//
// Bit layout is identical to arkworks' Fp256<MontBackend<FrConfig,4>> ([u64;4] LE limbs of a*R mod r),
// so a `&[Fr]` from the reference (src/ligero/mod.rs: every Vec<F>) is this struct array, zero-copy.
//
// The multiplier is a CIOS Montgomery product written as interleaved even/odd carry chains
// (mad.lo.cc / madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.U32 with predicate carries).  Every
// chain lives inside ONE asm statement, so the PTX carry flag never crosses a statement boundary and
// the compiler is free to interleave independent multiplications.  The same statements have a plain
// C emulation (no __CUDA_ARCH__) so the algorithm is unit-tested on the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LG_HD __host__ __device__ __forceinline__
#define LG_D __device__ __forceinline__
#else
#define LG_HD inline
#define LG_D inline
#endif

namespace lg {

struct alignas(16) Fr {
  uint32_t v[8];
};

// modulus r, little-endian 32-bit limbs
#define LG_P0 0xf0000001u
#define LG_P1 0x43e1f593u
#define LG_P2 0x79b97091u
#define LG_P3 0x2833e848u
#define LG_P4 0x8181585du
#define LG_P5 0xb85045b6u
#define LG_P6 0xe131a029u
#define LG_P7 0x30644e72u
#define LG_INV 0xefffffffu  // -r^{-1} mod 2^32

LG_HD uint32_t fr_p(int i) {
  switch (i) {
    case 0: return LG_P0; case 1: return LG_P1; case 2: return LG_P2; case 3: return LG_P3;
    case 4: return LG_P4; case 5: return LG_P5; case 6: return LG_P6; default: return LG_P7;
  }
}

// R mod r  (Montgomery form of 1)
LG_HD Fr fr_one() {
  Fr o;
  o.v[0] = 0x4ffffffbu; o.v[1] = 0xac96341cu; o.v[2] = 0x9f60cd29u; o.v[3] = 0x36fc7695u;
  o.v[4] = 0x7879462eu; o.v[5] = 0x666ea36fu; o.v[6] = 0x9a07df2fu; o.v[7] = 0x0e0a77c1u;
  return o;
}
// R^2 mod r: multiplying by it takes a plain integer (< r) to its Montgomery form
LG_HD Fr fr_r2() {
  Fr o;
  o.v[0] = 0xae216da7u; o.v[1] = 0x1bb8e645u; o.v[2] = 0xe35c59e3u; o.v[3] = 0x53fe3ab1u;
  o.v[4] = 0x53bb8085u; o.v[5] = 0x8c49833du; o.v[6] = 0x7f4e44a5u; o.v[7] = 0x0216d0b1u;
  return o;
}
LG_HD Fr fr_zero() {
  Fr o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = 0;
  return o;
}
LG_HD bool fr_is_zero(const Fr& a) {
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x |= a.v[i];
  return x == 0;
}
LG_HD bool fr_eq(const Fr& a, const Fr& b) {
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x |= a.v[i] ^ b.v[i];
  return x == 0;
}

// ------------------------------------------------------------------------------------------
// carry-chain statements
// ------------------------------------------------------------------------------------------
#ifndef __CUDA_ARCH__
namespace emu {
inline uint32_t addc(uint32_t a, uint32_t b, uint32_t& c) {
  uint64_t t = (uint64_t)a + b + c;
  c = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t subb(uint32_t a, uint32_t b, uint32_t& br) {
  uint64_t t = (uint64_t)a - b - br;
  br = (uint32_t)(t >> 63);
  return (uint32_t)t;
}
inline uint32_t lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
}  // namespace emu
#endif

// acc[2j],acc[2j+1] = x_j * b   (no carries: each product owns its pair)
LG_HD void fr_mul4(uint32_t acc[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mul.lo.u32 %0, %8, %12; mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12; mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12; mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12; mul.hi.u32 %7, %11, %12;"
      : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
        "=r"(acc[7])
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
  const uint32_t x[4] = {x0, x1, x2, x3};
  for (int j = 0; j < 4; j++) {
    acc[2 * j] = emu::lo(x[j], b);
    acc[2 * j + 1] = emu::hi(x[j], b);
  }
#endif
}

// acc += {x0,x1,x2,x3} * b laid out pairwise, one carry chain through acc[0..7]; top += carry-out
LG_HD void fr_mad4_carry(uint32_t acc[8], uint32_t& top, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                         uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %9, %13, %0; madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4; madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6; madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "+r"(top)
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
  const uint32_t x[4] = {x0, x1, x2, x3};
  uint32_t c = 0;
  for (int j = 0; j < 4; j++) {
    acc[2 * j] = emu::addc(acc[2 * j], emu::lo(x[j], b), c);
    acc[2 * j + 1] = emu::addc(acc[2 * j + 1], emu::hi(x[j], b), c);
  }
  top += c;
#endif
}

// same, carry-out of the chain provably zero for our operand ranges (dropped)
LG_HD void fr_mad4(uint32_t acc[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7])
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
  uint32_t dummy = 0;
  fr_mad4_carry(acc, dummy, x0, x1, x2, x3, b);
#endif
}

// lo0 += sh[1] (carry c);  sh[j-1],sh[j] = x_j*b + sh[j+1],sh[j+2] + c...  : the "shift right by two limbs
// while accumulating" step of the even/odd CIOS (operand x = odd limbs of a).
LG_HD void fr_mad4_shift(uint32_t sh[8], uint32_t& lo0, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                         uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %8, %8, %1;\n\t"
      "madc.lo.cc.u32 %0, %9, %13, %2; madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %4; madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %6; madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, 0; madc.hi.u32 %7, %12, %13, 0;"
      : "+r"(sh[0]), "+r"(sh[1]), "+r"(sh[2]), "+r"(sh[3]), "+r"(sh[4]), "+r"(sh[5]), "+r"(sh[6]), "+r"(sh[7]),
        "+r"(lo0)
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
  const uint32_t x[4] = {x0, x1, x2, x3};
  uint32_t c = 0;
  lo0 = emu::addc(lo0, sh[1], c);
  for (int j = 0; j < 4; j++) {
    uint32_t s0 = (2 * j + 2 < 8) ? sh[2 * j + 2] : 0, s1 = (2 * j + 3 < 8) ? sh[2 * j + 3] : 0;
    sh[2 * j] = emu::addc(s0, emu::lo(x[j], b), c);
    sh[2 * j + 1] = emu::addc(s1, emu::hi(x[j], b), c);
  }
#endif
}

// conditional subtract of r: x in [0, 2r) -> [0, r)
LG_HD void fr_final_sub(uint32_t x[8]) {
  uint32_t t[8], borrow;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22; subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
        "=r"(borrow)
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(LG_P0),
        "r"(LG_P1), "r"(LG_P2), "r"(LG_P3), "r"(LG_P4), "r"(LG_P5), "r"(LG_P6), "r"(LG_P7));
#else
  uint32_t br = 0;
  for (int i = 0; i < 8; i++) t[i] = emu::subb(x[i], fr_p(i), br);
  borrow = br ? 0xffffffffu : 0u;
#endif
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = borrow ? x[i] : t[i];
}

// one CIOS row: accumulate a*bi, then one Montgomery reduction step.  On exit ev[0] == 0 and the running
// value is  sum_{j>=1} ev[j] W^(j-1) + sum_j od[j] W^j ; callers alternate (ev, od) <-> (od, ev).
template <bool FIRST>
LG_HD void fr_cios_row(uint32_t ev[8], uint32_t od[8], const uint32_t a[8], uint32_t bi) {
  if (FIRST) {
    fr_mul4(od, a[1], a[3], a[5], a[7], bi);
    fr_mul4(ev, a[0], a[2], a[4], a[6], bi);
  } else {
    fr_mad4_shift(od, ev[0], a[1], a[3], a[5], a[7], bi);
    fr_mad4_carry(ev, od[7], a[0], a[2], a[4], a[6], bi);
  }
  uint32_t m = ev[0] * LG_INV;
  fr_mad4(od, LG_P1, LG_P3, LG_P5, LG_P7, m);
  fr_mad4_carry(ev, od[7], LG_P0, LG_P2, LG_P4, LG_P6, m);
}

LG_HD Fr fr_mul(const Fr& a, const Fr& b) {
  uint32_t ev[8], od[8];
  fr_cios_row<true>(ev, od, a.v, b.v[0]);
  fr_cios_row<false>(od, ev, a.v, b.v[1]);
  fr_cios_row<false>(ev, od, a.v, b.v[2]);
  fr_cios_row<false>(od, ev, a.v, b.v[3]);
  fr_cios_row<false>(ev, od, a.v, b.v[4]);
  fr_cios_row<false>(od, ev, a.v, b.v[5]);
  fr_cios_row<false>(ev, od, a.v, b.v[6]);
  fr_cios_row<false>(od, ev, a.v, b.v[7]);
  // last call had (ev_param, od_param) = (od, ev): od[0] == 0; value = sum_{j>=1} od[j] W^(j-1) + sum ev[j] W^j
  Fr r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, %17; addc.cc.u32 %2, %10, %18; addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20; addc.cc.u32 %5, %13, %21; addc.cc.u32 %6, %14, %22; addc.u32 %7, %15, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7])
      : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
#else
  uint32_t c = 0;
  for (int i = 0; i < 7; i++) r.v[i] = emu::addc(ev[i], od[i + 1], c);
  r.v[7] = ev[7] + c;
#endif
  fr_final_sub(r.v);
  return r;
}

LG_HD Fr fr_sqr(const Fr& a) { return fr_mul(a, a); }

// Montgomery reduction of a single element: a * R^{-1} mod r, i.e. Montgomery form -> canonical integer.
LG_HD Fr fr_from_mont(const Fr& a) {
  uint32_t ev[8], od[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { ev[i] = a.v[i]; od[i] = 0; }
  // 8 reduction rows; same even/odd alternation as fr_mul with the a*b terms absent
#pragma unroll
  for (int row = 0; row < 8; row++) {
    uint32_t* e = (row & 1) ? od : ev;
    uint32_t* o = (row & 1) ? ev : od;
    if (row) fr_mad4_shift(o, e[0], 0, 0, 0, 0, 0);
    uint32_t m = e[0] * LG_INV;
    fr_mad4(o, LG_P1, LG_P3, LG_P5, LG_P7, m);
    fr_mad4_carry(e, o[7], LG_P0, LG_P2, LG_P4, LG_P6, m);
  }
  Fr r;
  uint32_t c = 0;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, %17; addc.cc.u32 %2, %10, %18; addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20; addc.cc.u32 %5, %13, %21; addc.cc.u32 %6, %14, %22; addc.u32 %7, %15, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7])
      : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
  (void)c;
#else
  for (int i = 0; i < 7; i++) r.v[i] = emu::addc(ev[i], od[i + 1], c);
  r.v[7] = ev[7] + c;
#endif
  fr_final_sub(r.v);
  return r;
}

LG_HD Fr fr_add(const Fr& a, const Fr& b) {
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
  fr_final_sub(r.v);  // a+b < 2r < 2^255: no carry out of limb 7
  return r;
}

LG_HD Fr fr_sub(const Fr& a, const Fr& b) {
  Fr r;
  uint32_t borrow;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22; subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7]), "=r"(borrow)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // add back r masked by the borrow
  asm("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12; addc.cc.u32 %5, %5, %13; addc.cc.u32 %6, %6, %14; addc.u32 %7, %7, %15;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
        "+r"(r.v[7])
      : "r"(LG_P0 & borrow), "r"(LG_P1 & borrow), "r"(LG_P2 & borrow), "r"(LG_P3 & borrow), "r"(LG_P4 & borrow),
        "r"(LG_P5 & borrow), "r"(LG_P6 & borrow), "r"(LG_P7 & borrow));
#else
  uint32_t br = 0;
  for (int i = 0; i < 8; i++) r.v[i] = emu::subb(a.v[i], b.v[i], br);
  borrow = br ? 0xffffffffu : 0u;
  uint32_t c = 0;
  for (int i = 0; i < 8; i++) r.v[i] = emu::addc(r.v[i], fr_p(i) & borrow, c);
#endif
  return r;
}

LG_HD Fr fr_neg(const Fr& a) { return fr_sub(fr_zero(), a); }

}  // namespace lg
