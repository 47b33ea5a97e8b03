// Host check of ligero_b200/csrc/host_field.h: prints a, b, both Montgomery products (portable and, when the CPU has
// BMI2, the MULX variant) and a + b, a - b as hex; tests/test_host_field.py checks them against Python integers.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../ligero_b200/csrc/host_field.h"

using namespace lgh;

static void put(const char* k, const Fq& x) { printf("%s %016llx%016llx%016llx%016llx\n", k, (unsigned long long)x.l[3],
                                                     (unsigned long long)x.l[2], (unsigned long long)x.l[1], (unsigned long long)x.l[0]); }

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000;
  std::mt19937_64 g(argc > 2 ? atoi(argv[2]) : 1);
#ifdef LGH_HAVE_MULX
  printf("mulx %d\n", cpu_has_mulx() ? 1 : 0);
#else
  printf("mulx 0\n");
#endif
  for (int it = 0; it < n; it++) {
    Fq a, b;
    for (int i = 0; i < 4; i++) {
      a.l[i] = g();
      b.l[i] = g();
    }
    a.l[3] &= (1ULL << 62) - 1;
    b.l[3] &= (1ULL << 62) - 1;
    if (it % 7 == 0) {  // r - 1 - small
      for (int i = 0; i < 4; i++) a.l[i] = kP[i];
      a.l[0] -= 1 + (it % 5);
    }
    if (it % 11 == 0) b = (it % 22 == 0) ? kZero : kOne;
    while (geq_p(a.l)) sub_p(a.l);
    while (geq_p(b.l)) sub_p(b.l);
    put("a", a);
    put("b", b);
    put("mp", mul_portable(a, b));
    put("md", mul(a, b));
#ifdef LGH_HAVE_MULX
    if (cpu_has_mulx()) put("mx", mul_mulx(a, b));
#endif
    put("add", add(a, b));
    put("sub", sub(a, b));
    put("p17", pow_u64(a, 17));
    put("p5", pow_u64(a, 5));
    put("p6", pow_u64(a, 6));
  }
  return 0;
}
