// oracle-build stand-in for Dune::FloatCmp (test infrastructure).
// dune-common's default comparison is "relativeWeak" with epsilon = 8 * machine epsilon:
//   eq(a,b)  <=>  |a-b| <= eps * max(|a|,|b|)
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
namespace Dune::FloatCmp {
template <class T>
constexpr T defaultEps() { return T(8) * std::numeric_limits<T>::epsilon(); }
template <class T>
bool eq(const T& a, const T& b, T eps = defaultEps<T>()) {
  return std::abs(a - b) <= eps * std::max(std::abs(a), std::abs(b));
}
template <class T>
bool ne(const T& a, const T& b, T eps = defaultEps<T>()) { return !eq(a, b, eps); }
template <class T>
bool gt(const T& a, const T& b, T eps = defaultEps<T>()) { return a > b && !eq(a, b, eps); }
template <class T>
bool lt(const T& a, const T& b, T eps = defaultEps<T>()) { return a < b && !eq(a, b, eps); }
template <class T>
bool ge(const T& a, const T& b, T eps = defaultEps<T>()) { return a > b || eq(a, b, eps); }
template <class T>
bool le(const T& a, const T& b, T eps = defaultEps<T>()) { return a < b || eq(a, b, eps); }
}  // namespace Dune::FloatCmp
