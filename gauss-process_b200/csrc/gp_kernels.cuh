// gp_kernels.cuh -- covariance functions of the 13 GP kernels for 1-D inputs.
//
// Formulas and operation order follow reference src/gpr/kernels.py:90-382 (NumPy evaluates each
// ufunc separately, so products and sums are NOT contracted into FMAs here: mul_/add_ below pin the
// rounding).  Parameter order per kernel is the reference's get_params() order (include/b200id.h).
#pragma once
#include "common.cuh"
#include "gp_bessel.cuh"

namespace b200id {

#ifdef __CUDA_ARCH__
#define B200ID_MUL(a, b) __dmul_rn((a), (b))
#define B200ID_ADD(a, b) __dadd_rn((a), (b))
#define B200ID_SUB(a, b) __dsub_rn((a), (b))
#define B200ID_DIV(a, b) __ddiv_rn((a), (b))
#else
// host build (selftest only): compiled with -ffp-contract=off semantics via volatile-free helpers
static inline double b200id_nofuse_mul(double a, double b) { volatile double r = a * b; return r; }
static inline double b200id_nofuse_add(double a, double b) { volatile double r = a + b; return r; }
#define B200ID_MUL(a, b) b200id_nofuse_mul((a), (b))
#define B200ID_ADD(a, b) b200id_nofuse_add((a), (b))
#define B200ID_SUB(a, b) b200id_nofuse_add((a), -(b))
#define B200ID_DIV(a, b) ((a) / (b))
#endif

struct KernelParams {
    int id;
    double p0, p1, p2;
    // derived once per candidate
    double c0, c1;
};

B200ID_HD KernelParams make_kernel_params(int id, const double* theta) {
    KernelParams k;
    k.id = id;
    k.p0 = theta[0];
    k.p1 = theta[1];
    k.p2 = (id == B200ID_K_DC || id == B200ID_K_RQ || id == B200ID_K_MATERN_NU) ? theta[2] : 0.0;
    k.c0 = 0.0;
    k.c1 = 0.0;
    switch (id) {
        case B200ID_K_RBF: k.c0 = B200ID_MUL(k.p0, k.p0); break;                       // length_scale ** 2
        case B200ID_K_RQ:  k.c0 = B200ID_MUL(B200ID_MUL(2.0, k.p2), B200ID_MUL(k.p0, k.p0)); break;  // 2*alpha*ls**2
        case B200ID_K_SS2: k.c0 = B200ID_MUL(3.0, k.p0); break;                        // 3.0 * beta
        case B200ID_K_MATERN_NU:                                                        // np.sqrt(2.0 * nu), 2**(1-nu) / gamma(nu)
            k.c0 = sqrt(B200ID_MUL(2.0, k.p2)); k.c1 = matern_scale(k.p2); break;
        default: break;
    }
    return k;
}

B200ID_HD double heaviside(double x) { return x >= 0.0 ? 1.0 : 0.0; }

// k(xa, xb) with the kernel id as a compile-time constant (the switch folds away); xa indexes the row
// (X1), xb the column (X2).
template <int ID>
B200ID_HD double kernel_entry_t(const KernelParams& k, double xa, double xb) {
    switch (ID) {
        case B200ID_K_RBF: {            // variance * exp(-0.5 * d2 / ls**2)            kernels.py:98-100
            const double d = B200ID_SUB(xa, xb);
            const double d2 = B200ID_MUL(d, d);
            return B200ID_MUL(k.p1, exp(B200ID_DIV(B200ID_MUL(-0.5, d2), k.c0)));
        }
        case B200ID_K_MATERN12:
        case B200ID_K_MATERN32:
        case B200ID_K_MATERN52: {       // kernels.py:121-136
            const double d = B200ID_DIV(fabs(B200ID_SUB(xa, xb)), k.p0);
            double v;
            if (ID == B200ID_K_MATERN12) {
                v = exp(-d);
            } else if (ID == B200ID_K_MATERN32) {
                const double s3 = 1.7320508075688772;       // np.sqrt(3.0)
                const double a = B200ID_MUL(s3, d);
                v = B200ID_MUL(B200ID_ADD(1.0, a), exp(-a));
            } else {
                const double s5 = 2.23606797749979;         // np.sqrt(5.0)
                const double a = B200ID_MUL(s5, d);
                const double q = B200ID_MUL(5.0 / 3.0, B200ID_MUL(d, d));
                v = B200ID_MUL(B200ID_ADD(B200ID_ADD(1.0, a), q), exp(-a));
            }
            return B200ID_MUL(v, k.p1);
        }
        case B200ID_K_MATERN_NU: {      // general nu (theta[2], fixed): kernels.py:131-135
            const double d = B200ID_DIV(fabs(B200ID_SUB(xa, xb)), k.p0);
            double v = 1.0;                                  // K[dists == 0] = 1.0
            if (d != 0.0) {
                v = pow(d, k.p2);                            // K = dists ** nu
                v = B200ID_MUL(v, bessel_k(k.p2, B200ID_MUL(k.c0, d)));   // K *= kv(nu, sqrt(2 nu) * dists)
                v = B200ID_MUL(v, k.c1);                     // K *= 2 ** (1 - nu) / gamma(nu)
            }
            return B200ID_MUL(v, k.p1);
        }
        case B200ID_K_RQ: {             // variance * (1 + d2 / (2 alpha ls^2)) ** (-alpha)   kernels.py:158-161
            const double d = B200ID_SUB(xa, xb);
            const double base = B200ID_ADD(1.0, B200ID_DIV(B200ID_MUL(d, d), k.c0));
            return B200ID_MUL(k.p1, pow(base, -k.p2));
        }
        case B200ID_K_EXP: {            // variance * (exp(-omega (x+x')) * (H H'))     kernels.py:183-188
            const double e = exp(B200ID_MUL(-k.p0, B200ID_ADD(xa, xb)));
            return B200ID_MUL(k.p1, B200ID_MUL(e, B200ID_MUL(heaviside(xa), heaviside(xb))));
        }
        case B200ID_K_TC: {             // variance * (exp(-omega max(x,x')) * (H H'))  kernels.py:209-214
            const double e = exp(B200ID_MUL(-k.p0, fmax(xa, xb)));
            return B200ID_MUL(k.p1, B200ID_MUL(e, B200ID_MUL(heaviside(xa), heaviside(xb))));
        }
        case B200ID_K_DC: {             // beta * (alpha ** ((i+j)/2) * rho ** |i-j|)   kernels.py:235-240
            const double i = rint(xa), j = rint(xb);
            const double hs = (i + j) / 2.0;
            const double gap = fabs(i - j);
            return B200ID_MUL(k.p1, B200ID_MUL(pow(k.p0, hs), pow(k.p2, gap)));
        }
        case B200ID_K_DI: {             // beta * alpha ** i on the rounded diagonal      kernels.py:261-268
            const double i = rint(xa), j = rint(xb);
            return (i == j) ? B200ID_MUL(k.p0, pow(k.p1, i)) : 0.0;
        }
        case B200ID_K_SS1:              // variance * exp(-beta min(s,t))               kernels.py:290-291
            return B200ID_MUL(k.p1, exp(B200ID_MUL(-k.p0, fmin(xa, xb))));
        case B200ID_K_SS2: {            // kernels.py:312-319
            const double mx = fmax(xa, xb);
            const double t1 = B200ID_MUL(0.5, exp(B200ID_MUL(-k.p0, B200ID_ADD(B200ID_ADD(xa, xb), mx))));
            const double t2 = B200ID_MUL(1.0 / 6.0, exp(B200ID_MUL(B200ID_MUL(-3.0, k.p0), mx)));
            return B200ID_MUL(k.p1, B200ID_SUB(t1, t2));
        }
        case B200ID_K_SSHF: {           // variance * (-1)^(i+j) * max(e^-bs, e^-bt)     kernels.py:340-345
            const double i = rint(xa), j = rint(xb);
            const double par = fabs(fmod(i + j, 2.0));
            const double sign = par == 0.0 ? 1.0 : -1.0;
            const double m = fmax(exp(B200ID_MUL(-k.p0, xa)), exp(B200ID_MUL(-k.p0, xb)));
            return B200ID_MUL(B200ID_MUL(k.p1, sign), m);
        }
        case B200ID_K_STABLE_SPLINE: {  // variance * (0.5 r^2 (R - r/3))               kernels.py:366-372
            const double ea = exp(B200ID_MUL(-k.p0, xa)), eb = exp(B200ID_MUL(-k.p0, xb));
            const double r = fmin(ea, eb), R = fmax(ea, eb);
            const double v = B200ID_MUL(B200ID_MUL(0.5, B200ID_MUL(r, r)), B200ID_SUB(R, B200ID_DIV(r, 3.0)));
            return B200ID_MUL(k.p1, v);
        }
        default:
            return nan("");
    }
}

// run FN<ID>(args...) for the runtime kernel id
#define B200ID_DISPATCH_KERNEL(id, FN, ...)                                   \
    switch (id) {                                                             \
        case B200ID_K_RBF: FN<B200ID_K_RBF>(__VA_ARGS__); break;              \
        case B200ID_K_MATERN12: FN<B200ID_K_MATERN12>(__VA_ARGS__); break;    \
        case B200ID_K_MATERN32: FN<B200ID_K_MATERN32>(__VA_ARGS__); break;    \
        case B200ID_K_MATERN52: FN<B200ID_K_MATERN52>(__VA_ARGS__); break;    \
        case B200ID_K_EXP: FN<B200ID_K_EXP>(__VA_ARGS__); break;              \
        case B200ID_K_DC: FN<B200ID_K_DC>(__VA_ARGS__); break;                \
        case B200ID_K_DI: FN<B200ID_K_DI>(__VA_ARGS__); break;                \
        case B200ID_K_SS1: FN<B200ID_K_SS1>(__VA_ARGS__); break;              \
        case B200ID_K_SS2: FN<B200ID_K_SS2>(__VA_ARGS__); break;              \
        case B200ID_K_SSHF: FN<B200ID_K_SSHF>(__VA_ARGS__); break;            \
        case B200ID_K_STABLE_SPLINE: FN<B200ID_K_STABLE_SPLINE>(__VA_ARGS__); break; \
        case B200ID_K_RQ: FN<B200ID_K_RQ>(__VA_ARGS__); break;                \
        case B200ID_K_TC: FN<B200ID_K_TC>(__VA_ARGS__); break;                \
        case B200ID_K_MATERN_NU: FN<B200ID_K_MATERN_NU>(__VA_ARGS__); break;  \
        default: break;                                                       \
    }

template <int ID>
B200ID_HD void kernel_entry_into(const KernelParams& k, double xa, double xb, double* out) {
    *out = kernel_entry_t<ID>(k, xa, xb);
}

// runtime-id form (single evaluations: Gram export, predict, selftests)
B200ID_HD double kernel_entry(const KernelParams& k, double xa, double xb) {
    double v = nan("");
    B200ID_DISPATCH_KERNEL(k.id, kernel_entry_into, k, xa, xb, &v)
    return v;
}

B200ID_HD int kernel_n_params(int id) {
    return (id == B200ID_K_DC || id == B200ID_K_RQ) ? 3 : 2;
}

}  // namespace b200id
