// Stand-in for <boost/math/distributions/normal.hpp> (Boost is not in this image): the one class and the one free function the reference's
// DepthFilter::updateSeed uses (src/depth_filter.cpp: boost::math::normal_distribution<double> nd(mean, sd); boost::math::pdf(nd, x)).
// pdf follows Boost's own formula: exp(-(x - mean)^2 / (2 sd^2)) / (sd * sqrt(2 pi)).
#pragma once
#include <cmath>
namespace boost { namespace math {
template <class T = double>
class normal_distribution {
 public:
  normal_distribution(T mean = 0, T sd = 1) : m_(mean), s_(sd) {}
  T mean() const { return m_; }
  T standard_deviation() const { return s_; }
 private:
  T m_, s_;
};
typedef normal_distribution<double> normal;
template <class T>
inline T pdf(const normal_distribution<T>& d, const T& x) {
  const T sd = d.standard_deviation(), diff = x - d.mean();
  T e = -(diff * diff) / (2 * sd * sd);
  e = std::exp(e);
  return e / (sd * std::sqrt(2 * static_cast<T>(3.141592653589793238462643383279502884L)));
}
}}  // namespace boost::math
