// boost/math pulls <complex> and friends into every translation unit of the reference (include/estimator.h:805 names
// std::complex without including it).
#pragma once
#include <complex>
#include <utility>
namespace boost { namespace math { namespace tools {
template <class F, class T> std::pair<T, T> brent_find_minima(F, T a, T, int) { return {a, a}; }
}}}
