// gp_bessel.cuh -- modified Bessel function of the second kind K_nu(x), real order nu > 0, x > 0, FP64, for the
// general-nu Matern kernel (reference src/gpr/kernels.py:131-135: dists**nu * scipy.special.kv(nu, sqrt(2 nu) dists)
// * 2**(1-nu) / gamma(nu)).  scipy's kv is AMOS zbesk, which is not in the reference checkout; this is the textbook
// pair of algorithms with the same accuracy class (a few ulp):
//   * x <= 2: Temme's series for K_mu, K_mu+1 with |mu| <= 1/2 (N. M. Temme, J. Comput. Phys. 19 (1975) 324),
//     the reciprocal-gamma combinations taken from the Taylor series of 1/Gamma(1+x) split into even and odd parts so
//     that gamma1 = (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu) has no cancellation;
//   * x > 2: Steed's algorithm on the second continued fraction (CF2) with the Thompson-Barnett sum;
//   then the forward recurrence K_{v+1} = K_{v-1} + (2 v / x) K_v (stable for K) from mu up to nu.
// Compiles for host and device (the host build serves the CPU selftests against scipy).
#pragma once
#include "common.cuh"

namespace b200id {

// Taylor coefficients of 1/Gamma(1+x) at 0 (mpmath, 60 digits, rounded): truncation < 1e-17 for |x| <= 1/2
#define B200ID_RGAMMA_COEF                                                                               \
    {1.0, 0.57721566490153286061, -0.65587807152025388108, -0.042002635034095235529,                     \
     0.1665386113822914895, -0.042197734555544336748, -0.0096219715278769735621,                         \
     0.0072189432466630995424, -0.0011651675918590651121, -0.00021524167411495097282,                    \
     0.00012805028238811618615, -0.000020134854780788238656, -1.2504934821426706573e-6,                  \
     1.1330272319816958824e-6, -2.0563384169776071035e-7, 6.1160951044814158179e-9,                      \
     5.0020076444692229301e-9, -1.1812745704870201446e-9, 1.0434267116911005105e-10,                     \
     7.782263439905071254e-12, -3.6968056186422057082e-12, 5.100370287454475979e-13}

// E(mu^2) = sum_{k even} c_k mu^k, O(mu^2) = sum_{k odd} c_k mu^(k-1):  1/Gamma(1 +- mu) = E +- mu O
B200ID_HD void rgamma_even_odd(double mu, double& E, double& O) {
    const double c[22] = B200ID_RGAMMA_COEF;
    const double m2 = mu * mu;
    double e = c[20], o = c[21];
#pragma unroll
    for (int k = 18; k >= 0; k -= 2) {
        e = fma(e, m2, c[k]);
        o = fma(o, m2, c[k + 1]);
    }
    E = e; O = o;
}

// K_mu(x), K_{mu+1}(x) for |mu| <= 1/2
B200ID_HD void bessel_k_pair(double mu, double x, double& k0, double& k1) {
    const double PI = 3.141592653589793238462643383279502884;
    if (x <= 2.0) {
        double E, O;
        rgamma_even_odd(mu, E, O);
        const double gampl = fma(mu, O, E), gammi = fma(-mu, O, E);       // 1/Gamma(1+mu), 1/Gamma(1-mu)
        const double gam1 = -O, gam2 = E;
        const double x2 = 0.5 * x;
        const double pimu = PI * mu;
        const double fact = fabs(pimu) < 1e-15 ? 1.0 : pimu / sin(pimu);
        const double d = -log(x2);
        const double e = mu * d;
        const double fact2 = fabs(e) < 1e-15 ? 1.0 : sinh(e) / e;
        double ff = fact * (gam1 * cosh(e) + gam2 * fact2 * d);
        double sum = ff;
        const double ee = exp(e);
        double p = 0.5 * ee / gampl;
        double q = 0.5 / (ee * gammi);
        double c = 1.0;
        const double dd = x2 * x2;
        double sum1 = p;
        const double mu2 = mu * mu;
        for (int i = 1; i < 500; ++i) {
            const double di = (double)i;
            ff = (di * ff + p + q) / (di * di - mu2);
            c *= dd / di;
            p /= (di - mu);
            q /= (di + mu);
            const double del = c * ff;
            sum += del;
            sum1 += c * (p - di * ff);
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        k0 = sum;
        k1 = sum1 * (2.0 / x);
        return;
    }
    // Steed / Thompson-Barnett on CF2
    double a = mu * mu - 0.25;
    double b = 2.0 * (x + 1.0);
    double D = 1.0 / b;
    double f = D, delta = D;
    double prev = 0.0, current = 1.0;
    double C = -a, Q = -a;
    double S = 1.0 + Q * delta;
    for (int k = 2; k < 10000; ++k) {
        a -= 2.0 * (double)(k - 1);
        b += 2.0;
        D = 1.0 / (b + a * D);
        delta *= b * D - 1.0;
        f += delta;
        const double qn = (prev - (b - 2.0) * current) / a;
        prev = current;
        current = qn;
        C *= -a / (double)k;
        Q += C * qn;
        S += Q * delta;
        if (fabs(Q * delta) < fabs(S) * 1e-17) break;
    }
    k0 = sqrt(PI / (2.0 * x)) * exp(-x) / S;
    k1 = k0 * (0.5 + mu + x + (mu * mu - 0.25) * f) / x;
}

// K_nu(x), nu >= 0, x > 0
B200ID_HD double bessel_k(double nu, double x) {
    const int nl = (int)(nu + 0.5);
    const double mu = nu - (double)nl;
    double km, k1;
    bessel_k_pair(mu, x, km, k1);
    const double xi2 = 2.0 / x;
    for (int i = 1; i <= nl; ++i) {
        const double kt = fma((mu + (double)i) * xi2, k1, km);
        km = k1;
        k1 = kt;
    }
    return km;
}

// 2^(1-nu) / Gamma(nu): Gamma(nu) = Gamma(1+mu) prod_{i=1}^{nl-1} (mu + i)  (nl = 0: Gamma(mu) = Gamma(1+mu) / mu)
B200ID_HD double matern_scale(double nu) {
    const int nl = (int)(nu + 0.5);
    const double mu = nu - (double)nl;
    double E, O;
    rgamma_even_odd(mu, E, O);
    double g = 1.0 / fma(mu, O, E);
    if (nl == 0) g /= mu;
    for (int i = 1; i < nl; ++i) g *= (mu + (double)i);
    return pow(2.0, 1.0 - nu) / g;
}

}  // namespace b200id
