// Host harness for csrc/fastmath.cuh: max relative error of exp_neg / sqrt_pos against long double libm over the ranges the
// K* kernels use.   g++ -O2 -x c++ -o build/fastmath_check tools/fastmath_check.cu && build/fastmath_check
#include <cstdio>
#include <cmath>
#include <random>
#include "../trieste_b200/csrc/fastmath.cuh"
#ifndef FM_ITERS
#define FM_ITERS 20000000
#endif
int main() {
  const tb::fm::Consts C;
  std::mt19937_64 rng(1);
  double worst_e = 0, worst_s = 0, at_e = 0, at_s = 0;
  std::uniform_real_distribution<double> u(0.0, 1.0);
  for (int i = 0; i < FM_ITERS; ++i) {
    double s;
    const double r = u(rng);
    if (i % 3 == 0) s = r * 50.0; else if (i % 3 == 1) s = std::exp(-40.0 * r); else s = 700.0 * r;
    const double got = tb::fm::exp_neg(s, tb::fm::EXP2_TABLE_HOST, C);
    const long double ref = expl(-(long double)s);
    const double err = (double)fabsl(((long double)got - ref) / ref);
    if (err > worst_e) { worst_e = err; at_e = s; }
    const double x = std::exp(-82.0 + 164.0 * r);  // 2.4e-36 .. 1.6e35
    const double gs = tb::fm::sqrt_pos(x);
    const long double rs = sqrtl((long double)x);
    const double es = (double)fabsl(((long double)gs - rs) / rs);
    if (es > worst_s) { worst_s = es; at_s = x; }
  }
  const tb::fm::TrigConsts TC;
  double worst_c = 0, at_c = 0;
  for (int i = 0; i < FM_ITERS; ++i) {
    const double r = u(rng);
    const double a = (i % 2 ? 60.0 : 3000.0) * (2.0 * r - 1.0);
    const double got = tb::fm::cos_fast(a, TC);
    const double err = (double)fabsl((long double)got - cosl((long double)a));
    if (err > worst_c) { worst_c = err; at_c = a; }
  }
  printf("cos_fast: max ABS err %.3e at a = %.17g; cos_fast(0) = %.17g, cos_fast(pi) = %.17g\n", worst_c, at_c, tb::fm::cos_fast(0.0, TC),
         tb::fm::cos_fast(3.141592653589793, TC));
  if (worst_c > 1e-13) return 1;
  printf("exp_neg: max rel err %.3e at s = %.17g\nsqrt_pos: max rel err %.3e at x = %.17g\n", worst_e, at_e, worst_s, at_s);
  printf("exp_neg(0) = %.17g, exp_neg(1e-17) = %.17g, exp_neg(800) = %.3e, exp_neg(1e4) = %.3e\n", tb::fm::exp_neg(0.0, tb::fm::EXP2_TABLE_HOST, C),
         tb::fm::exp_neg(1e-17, tb::fm::EXP2_TABLE_HOST, C), tb::fm::exp_neg(800.0, tb::fm::EXP2_TABLE_HOST, C), tb::fm::exp_neg(1e4, tb::fm::EXP2_TABLE_HOST, C));
  return (worst_e < 4e-16 && worst_s < 2.3e-16) ? 0 : 1;
}
