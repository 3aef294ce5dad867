// oracle-build stand-in for dune/common/math.hh (test infrastructure)
#pragma once
namespace Dune {
template <class B, class E>
constexpr B power(B b, E e) {
  B r = 1;
  for (E i = 0; i < e; ++i) r *= b;
  return r;
}
template <class T>
constexpr T binomial(const T& n, const T& k) {
  if (k < 0 || k > n) return 0;
  T kk = (k > n - k) ? n - k : k;
  T r  = 1;
  for (T i = 0; i < kk; ++i) r = r * (n - i) / (i + 1);  // exact: product of i+1 consecutive ints divisible by (i+1)!
  return r;
}
}  // namespace Dune
