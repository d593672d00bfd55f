// sc_special.h -- the two special functions of the LSC impedance models, fp64, usable from
// device code and (for the CPU unit test that checks them against scipy) from host code.
//
//   exp_e1(x)    = exp(x) * E1(x)   the reference calls np.exp(a2) * scipy.special.exp1(a2)
//                                   (ocelot/cpbd/sc.py:331), 1e-16 <= x <= 40
//   bessel_k1(x) = K1(x)            scipy.special.k1 (ocelot/cpbd/sc.py:366), x > 0
//
// Neither is in the CUDA math library.  Both follow published formulas (Abramowitz & Stegun
// 5.1.11, 5.1.22, 9.6.24), not the cephes/specfun sources scipy wraps.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define OCL_HD __host__ __device__ __forceinline__
#else
#define OCL_HD inline
#endif

namespace ocl {

// exp(x) E1(x).
//   x <= 1: E1(x) = -gamma - ln x - sum_{k>=1} (-x)^k / (k k!)                    (A&S 5.1.11)
//   x  > 1: exp(x) E1(x) = 1/(x+ 1/(1+ 1/(x+ 2/(1+ 2/(x+ ...)))))                 (A&S 5.1.22),
//           evaluated bottom-up with a term count that covers x -> 1+.
OCL_HD double exp_e1(double x) {
    if (x <= 1.0) {
        double term = 1.0, sum = 0.0;
        for (int k = 1; k <= 30; ++k) {
            term *= -x / (double)k;                 // (-x)^k / k!
            const double add = term / (double)k;
            sum += add;
            if (fabs(add) < 1e-18 * fabs(sum)) break;
        }
        return exp(x) * (-0.57721566490153286061 - log(x) - sum);
    }
    const int m = 24 + (int)(96.0 / x);
    double t = 0.0;
    for (int k = m; k >= 1; --k) t = (double)k / (1.0 + (double)k / (x + t));
    return 1.0 / (x + t);
}

// K1(x) = int_0^inf exp(-x cosh t) cosh t dt                                       (A&S 9.6.24)
// The integrand is entire and decays doubly exponentially, so the trapezoid rule converges
// geometrically: error ~ exp(-pi^2 / h) from the strip of analyticity and ~ exp(-2 pi^2 / (x h^2))
// from the Gaussian-like peak of width x^(-1/2) at large x; h = min(1/8, 0.6 / sqrt(x)) keeps both
// below 1e-20 relative.
// Terms first, first+stride, ... of the trapezoid sum, unscaled: K1(x) = exp(-x) * h * (sum of all
// partials), with the t = 0 end point (weight 1/2) counted by the caller that owns first == 0.
// Lets a warp split one evaluation across its lanes.
OCL_HD double bessel_k1_step(double x) { return fmin(0.125, 0.6 / sqrt(x)); }
OCL_HD double bessel_k1_partial(double x, int first, int stride) {
    const double h = bessel_k1_step(x);
    // terms beyond x cosh t > 750 are below the smallest normal double relative to the sum
    const double tmax = log(2.0 * 750.0 / x) + 1e-9;
    const int nterm = (int)(tmax / h) + 1;
    double sum = (first == 0) ? 0.5 : 0.0;          // t = 0 end point, weight 1/2
    // exp(-x cosh t) = exp(-x) exp(-2 x sinh^2(t/2)): keeps the exponent's rounding error relative
    // to the small quantity x (cosh t - 1) instead of to x cosh t
    for (int k = (first == 0) ? stride : first; k <= nterm; k += stride) {
        const double t = h * (double)k;
        const double sh = sinh(0.5 * t);
        const double d = 2.0 * sh * sh;             // cosh t - 1
        sum += exp(-x * d) * (1.0 + d);
    }
    return sum;
}

OCL_HD double bessel_k1(double x) {
    if (!(x > 0.0)) return INFINITY;
    if (x > 745.0) return 0.0;
    return bessel_k1_step(x) * exp(-x) * bessel_k1_partial(x, 0, 1);
}

}  // namespace ocl
