// stand-in: the chi-squared quantile of Boost.Math (Blake-Zisserman's epsilon is the only caller).  NOT reference code: regularised lower
// incomplete gamma by series / continued fraction, inverted by bisection to machine precision; the test that uses it compares the
// epsilon against scipy.stats.chi2.ppf as well.
#ifndef KB_SHIM_BOOST_CHI_SQUARED
#define KB_SHIM_BOOST_CHI_SQUARED
#include <cmath>
#include <cstddef>
namespace boost { namespace math {
template <typename T = double>
class chi_squared_distribution {
 public:
  explicit chi_squared_distribution(T df) : df_(df) {}
  T degrees_of_freedom() const { return df_; }
 private:
  T df_;
};
namespace kb_shim_detail {
inline double gamma_p(double a, double x) {  // regularised lower incomplete gamma P(a, x)
  if (x <= 0.0) return 0.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {
    double sum = 1.0 / a, term = sum;
    for (int n = 1; n < 10000; ++n) {
      term *= x / (a + n);
      sum += term;
      if (std::fabs(term) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;
  for (int i = 1; i < 10000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < 1e-300) d = 1e-300;
    c = b + an / c;
    if (std::fabs(c) < 1e-300) c = 1e-300;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}
}  // namespace kb_shim_detail
template <typename T>
inline double quantile(const chi_squared_distribution<T>& d, double p) {
  const double a = 0.5 * (double)d.degrees_of_freedom();
  double lo = 0.0, hi = 1.0;
  while (kb_shim_detail::gamma_p(a, 0.5 * hi) < p) hi *= 2.0;
  for (int i = 0; i < 200 && hi - lo > 0.0; ++i) {
    const double mid = 0.5 * (lo + hi);
    if (mid == lo || mid == hi) break;
    (kb_shim_detail::gamma_p(a, 0.5 * mid) < p ? lo : hi) = mid;
  }
  return 0.5 * (lo + hi);
}
} }
#endif
