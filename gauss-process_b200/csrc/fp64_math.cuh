// fp64_math.cuh -- branch-free FP64 sincos for the FRF projection.
//
// The reference forms the phase as the FP64 product fl(w_k * t_i) and then takes np.cos / np.sin of
// that rounded value (frf_estimator.py:227-228).  Phases reach 1e7..1e9 rad on long records, where
// CUDA's sincos() leaves its 3-FMA path for Payne-Hanek (|x| > ~1e5).  The FRF only needs an
// ABSOLUTE error of ~1e-16 in sin/cos (the values are multiplied by samples and summed), so a
// three-constant Cody-Waite reduction with FMA is enough for any |x| < 2^40: the first FMA forms
// x - q*C1 with a single rounding of a result that is O(1), the next two fold in the tail of pi/2.
#pragma once
#include "common.cuh"

namespace b200id {

// pi/2 split into three FP64 terms (sum accurate to ~2^-160) and 2/pi.
#define B200ID_PIO2_HI  1.5707963267948966e+00   /* 0x3FF921FB54442D18 */
#define B200ID_PIO2_MID 6.123233995736766e-17    /* 0x3C91A62633145C07 */
#define B200ID_PIO2_LO  -1.4973849048591698e-33  /* 0xB91F1976B7ED8FBC */
#define B200ID_2_OVER_PI 6.3661977236758138e-01
#define B200ID_RND_MAGIC 6755399441055744.0      /* 1.5 * 2^52 */

B200ID_HD int lo_word(double x) {
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    int64_t b; memcpy(&b, &x, 8); return (int)(uint32_t)(b & 0xffffffff);
#endif
}

B200ID_HD double flip_sign_if(double x, int cond_bit /*0 or 1*/) {
#ifdef __CUDA_ARCH__
    int hi = __double2hiint(x) ^ (cond_bit << 31);
    return __hiloint2double(hi, __double2loint(x));
#else
    int64_t b; memcpy(&b, &x, 8); b ^= ((int64_t)cond_bit) << 63; memcpy(&x, &b, 8); return x;
#endif
}

// Polynomial and reduction constants live in constant memory on the device so that every DFMA of
// the Horner chains reads two registers + one constant-bank operand.  B200's FP64 pipe issues one
// warp DFMA per 2 cycles only when at most two distinct 64-bit REGISTER operands are read (three
// cost 3 cycles: measured 60% FP64-pipe utilisation with the constants held in registers).
struct SinCosCoef {
    double two_over_pi, magic, pio2_hi, pio2_mid;
    double s6, s5, s4, s3, s2, s1;
    double c6, c5, c4, c3, c2, c1;
};
#define B200ID_SINCOS_COEF_INIT                                                                      \
    {B200ID_2_OVER_PI, B200ID_RND_MAGIC, B200ID_PIO2_HI, B200ID_PIO2_MID,                             \
     1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,             \
     -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,            \
     -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,            \
     2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02}
#ifdef __CUDACC__
__constant__ SinCosCoef g_sincos_coef = B200ID_SINCOS_COEF_INIT;
#endif
static const SinCosCoef h_sincos_coef = B200ID_SINCOS_COEF_INIT;

// sin and cos of x, |x| < 2^36, absolute error < ~2e-16.  Two-constant Cody-Waite: the third term of
// pi/2 (1.5e-33) times q < 2^36 is below 1e-22 and is dropped.
B200ID_HD void sincos_cw(double x, double& s_out, double& c_out) {
#ifdef __CUDA_ARCH__
    const SinCosCoef& K = g_sincos_coef;
#else
    const SinCosCoef& K = h_sincos_coef;
#endif
    double qd = fma(x, K.two_over_pi, K.magic);
    const int q = lo_word(qd);
    qd -= K.magic;
    double r = fma(-qd, K.pio2_hi, x);
    r = fma(-qd, K.pio2_mid, r);
    const double z = r * r;
    // minimax polynomials on [-pi/4, pi/4] (fdlibm __kernel_sin / __kernel_cos coefficients)
    double ps = fma(K.s6, z, K.s5);
    ps = fma(ps, z, K.s4);
    ps = fma(ps, z, K.s3);
    ps = fma(ps, z, K.s2);
    ps = fma(ps, z, K.s1);
    const double sn = fma(z * r, ps, r);
    double pc = fma(K.c6, z, K.c5);
    pc = fma(pc, z, K.c4);
    pc = fma(pc, z, K.c3);
    pc = fma(pc, z, K.c2);
    pc = fma(pc, z, K.c1);
    const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
    // quadrant: q&1 swaps, q&2 negates sin, (q+1)&2 negates cos
    const bool swap = q & 1;
    const double s0 = swap ? cs : sn;
    const double c0 = swap ? sn : cs;
    s_out = flip_sign_if(s0, (q >> 1) & 1);
    c_out = flip_sign_if(c0, ((q + 1) >> 1) & 1);
}

}  // namespace b200id
