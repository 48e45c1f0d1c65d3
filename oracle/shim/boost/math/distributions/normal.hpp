// TEST INFRASTRUCTURE (oracle/): stand-in for <boost/math/distributions/normal.hpp>.
//
// The reference's lib/cbs/smooth.cpp:8,19-20,28 uses exactly three things from
// Boost.Math (Boost is NOT vendored in /root/reference and not installed in
// this image; reference CMakeLists.txt:25 floats the version, README says
// >= 1.56):
//     boost::math::normal_distribution<double> nd;   // standard normal
//     boost::math::quantile(nd, p)
//     boost::math::pdf(nd, x)
// Both only feed inflfact(trim) (smooth.cpp:13-31), one scalar per run.
// This shim provides them so that smooth.cpp compiles UNMODIFIED.  The quantile
// is Acklam's rational approximation polished with two Halley steps on
// erfc (full double precision, a few ulp from Boost's erfc_inv based value).
// The same routine (restated in C) is used by oracle/cbs_oracle.c and by the
// product's host code, so all three agree bit-for-bit on inflfact.
#ifndef ORACLE_SHIM_BOOST_NORMAL_HPP
#define ORACLE_SHIM_BOOST_NORMAL_HPP

#include <cmath>
#include <limits>
#include <stdexcept>

namespace boost {
namespace math {

template <typename Real = double>
class normal_distribution {
public:
    normal_distribution(Real mean = 0, Real sd = 1) : m_(mean), s_(sd) {}
    Real mean() const { return m_; }
    Real standard_deviation() const { return s_; }
private:
    Real m_, s_;
};

namespace shim_detail {

inline double std_normal_cdf(double x) { return 0.5 * std::erfc(-x * 0.70710678118654752440); }

inline double std_normal_pdf(double x) {
    return 0.39894228040143267794 * std::exp(-0.5 * x * x);
}

inline double std_normal_quantile(double p) {
    if (!(p > 0.0 && p < 1.0)) {
        // Boost's default error policy: overflow_error at the end points, domain_error outside
        if (p == 0.0 || p == 1.0) throw std::overflow_error("Error in function boost::math::quantile: Overflow Error");
        throw std::domain_error("Error in function boost::math::quantile: probability out of range");
    }
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01,  -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
    static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                3.754408661907416e+00};
    const double plow = 0.02425, phigh = 1.0 - plow;
    double x;
    if (p < plow) {
        const double q = std::sqrt(-2.0 * std::log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    } else if (p <= phigh) {
        const double q = p - 0.5, r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
    } else {
        const double q = std::sqrt(-2.0 * std::log(1.0 - p));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    for (int it = 0; it < 2; ++it) {
        const double e = std_normal_cdf(x) - p;
        const double u = e / std_normal_pdf(x);
        x = x - u / (1.0 + 0.5 * x * u);
    }
    return x;
}

} // namespace shim_detail

template <typename Real>
inline Real pdf(const normal_distribution<Real>& nd, const Real& x) {
    const double z = (static_cast<double>(x) - nd.mean()) / nd.standard_deviation();
    return static_cast<Real>(shim_detail::std_normal_pdf(z) / nd.standard_deviation());
}

template <typename Real>
inline Real quantile(const normal_distribution<Real>& nd, const Real& p) {
    return static_cast<Real>(nd.mean() + nd.standard_deviation() * shim_detail::std_normal_quantile(static_cast<double>(p)));
}

} // namespace math
} // namespace boost

#endif
