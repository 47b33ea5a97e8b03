// Host-side check of the lazy-reduction NTT arithmetic (plain-C emulation of the PTX carry chains in
// ligero_b200/csrc/fr_lazy.cuh).  Prints hex vectors; tests/test_fr_lazy_host.py verifies them with Python ints.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "../../ligero_b200/csrc/fr_lazy.cuh"
using namespace lg;

static uint64_t s = 0x9e3779b97f4a7c15ull;
static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); }
static void pr(const char* tag, const Fr& x) {
  printf("%s ", tag);
  for (int i = 7; i >= 0; i--) printf("%08x", x.v[i]);
  printf("\n");
}
static Fr rand_below_r() {
  Fr x;
  for (;;) {
    for (int i = 0; i < 8; i++) x.v[i] = rnd();
    x.v[7] &= 0x3fffffffu;
    bool lt = false;
    for (int i = 7; i >= 0; i--) { if (x.v[i] < fr_p(i)) { lt = true; break; } if (x.v[i] > fr_p(i)) break; }
    if (lt) return x;
  }
}
int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 100;
  if (argc > 2) s ^= strtoull(argv[2], 0, 10) * 0x2545F4914F6CDD1Dull;
  for (int it = 0; it < n; it++) {
    FrTw tw;
    tw.w = rand_below_r();
    if (it % 7 == 0) { tw.w = fr_zero(); tw.w.v[0] = it / 7; }          // small constants incl. 0 and 1
    if (it % 11 == 0) { for (int i = 0; i < 8; i++) tw.w.v[i] = fr_p(i); tw.w.v[0] -= 1 + it / 11; }  // r-1-...
    tw.p = fr_shoup_quotient(tw.w);
    Fr y;
    for (int i = 0; i < 8; i++) y.v[i] = rnd();
    if (it % 3 == 0) y.v[7] = 0xfffffff0u - (rnd() & 7);   // close to the bound 2^256(1-2^-29)
    if (y.v[7] > 0xfffffff0u) y.v[7] = 0xfffffff0u;
    if (it % 5 == 0) for (int i = 0; i < 6; i++) y.v[i] = 0xffffffffu;
    if (it % 13 == 0) y = fr_zero();
    pr("w", tw.w); pr("p", tw.p); pr("y", y);
    pr("t", fr_mul_shoup(y, tw));
    // butterflies on lazily reduced inputs
    Fr X = rand_below_r(), Y = rand_below_r();
    // push them into the lazy range: add r a few times
    Fr rr; for (int i = 0; i < 8; i++) rr.v[i] = fr_p(i);
    for (int a = it % 4; a > 0; a--) X = lz_add(X, rr);
    for (int a = (it / 4) % 4; a > 0; a--) Y = lz_add(Y, rr);
    pr("X", X); pr("Y", Y);
    { Fr a = X, b = Y; lz_bfly_dit(a, b, tw); pr("ditX", a); pr("ditY", b); pr("nX", fr_normalize(a)); pr("nY", fr_normalize(b)); }
    { Fr a = X, b = Y; lz_bfly_dit1(a, b); pr("dit1X", a); pr("dit1Y", b); }
    // DIF inputs must be < 2r + d: reduce first
    { Fr a = X, b = Y; lz_csub2r(a); lz_csub2r(b); pr("fX", a); pr("fY", b);
      Fr c = a, d = b; lz_bfly_dif(c, d, tw); pr("difX", c); pr("difY", d);
      c = a; d = b; lz_bfly_dif1(c, d); pr("dif1X", c); pr("dif1Y", d); }
  }
  return 0;
}
