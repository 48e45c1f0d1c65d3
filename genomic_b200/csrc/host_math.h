// host_math.h -- the one scalar of the smoothing pass that is computed on the host:
// inflfact(trim) (lib/cbs/smooth.cpp:13-31), the variance inflation factor of the trimmed
// normal.  The reference takes the normal quantile and density from Boost.Math (not vendored,
// version floating); this restates them (Acklam start + two Halley steps on erfc) and is
// bit-identical to the stand-in the oracle compiles the reference against.
#pragma once
#include <cmath>

namespace cbsg {

inline double std_normal_cdf(double x) { return 0.5 * std::erfc(-x * 0.70710678118654752440); }
inline double std_normal_pdf(double x) { return 0.39894228040143267794 * std::exp(-0.5 * x * x); }

inline double std_normal_quantile(double p) {
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

// precondition: 0 < trim < 0.5
inline double inflfact(double trim) {
    const double a = std_normal_quantile(1.0 - trim);
    const int ngrid = 10000;
    const double step = (2.0 * a) / ngrid;
    double sum = 0.0;
    for (int i = 0; i < ngrid; ++i) {
        const double left = -a + i * step;
        const double right = left + step;
        const double x = 0.5 * (left + right);
        sum += x * x * std_normal_pdf(x) / (1.0 - 2.0 * trim);
    }
    return 1.0 / (sum * step);
}

}  // namespace cbsg
