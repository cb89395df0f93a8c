// fir_fft.cu -- K6 fast path: causal FIR validation by in-shared-memory overlap-save FFT (FP64).
//
// Reference call sites: np.convolve(u, g, 'full')[:N] at src/fir_model/fir_fitting.py:285 and
// fir_validation.py:190, metrics at fir_fitting.py:288-306 / fir_validation.py:115-128.
//
// Direct form costs 2*taps FP64 flops per output (2048 at 1024 taps) and is FP64-issue bound at ~81 % of the
// FP64 peak (fir_conv.cu).  Overlap-save turns that into ~90 flops per output, which moves the kernel from
// the FP64 pipe to shared-memory bandwidth:
//   * one work unit = TWO consecutive output blocks A, B carried as the real and imaginary part of one
//     complex sequence z = uA + j uB of length N = 4096.  The taps are real, so IFFT(FFT(z) G) =
//     (uA * g) + j (uB * g): no real-FFT untangling is needed at all;
//   * radix-4 decimation-in-frequency forward stages leave the spectrum in base-4 digit-reversed order, the
//     tap spectrum G is produced by the SAME forward routine (so it is in the same order), and the inverse
//     runs the decimation-in-time stages in the opposite order: no reordering pass either.  The last forward
//     stage, the product with G and the first inverse stage touch the same four elements per thread and are
//     fused in registers;
//   * the circular convolution is valid for m in [taps-1, N): each block yields V = N - taps + 1 outputs.
// Data (N complex, padded one element per 8 against bank conflicts) + N/4 twiddles = 90 KB per CTA -> 2 CTAs
// per SM, persistent grid, error sums fused into the epilogue (y_pred is optional).
#include <stdint.h>
#include "common.cuh"

namespace b200id {

constexpr int FF_N = 4096;
constexpr int FF_THREADS = 256;
constexpr int FF_MAX_TAPS = FF_N / 2;       // keeps >= 50 % of every block valid
constexpr int FF_CTAS_PER_SM = 2;

__device__ __forceinline__ int ff_pad(int p) { return p + (p >> 3); }
constexpr int FF_DATA_ELEMS = FF_N + FF_N / 8;

struct cplx { double x, y; };
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)}; }
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return {fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y))}; }  // a * conj(b)
__device__ __forceinline__ cplx mul_mj(cplx a) { return {a.y, -a.x}; }     // -j a
__device__ __forceinline__ cplx mul_pj(cplx a) { return {-a.y, a.x}; }     // +j a
__device__ __forceinline__ cplx lds(const double2* s, int p) { const double2 v = s[ff_pad(p)]; return {v.x, v.y}; }
__device__ __forceinline__ void sts(double2* s, int p, cplx v) { s[ff_pad(p)] = make_double2(v.x, v.y); }

// tw[k] = exp(-2 pi i k / N), k < N/4
__device__ __forceinline__ void ff_make_twiddles(double2* tw) {
    for (int k = threadIdx.x; k < FF_N / 4; k += FF_THREADS) {
        double s, c;
        sincospi(2.0 * (double)k / (double)FF_N, &s, &c);
        tw[k] = make_double2(c, -s);
    }
}

// forward DIF stages L = N .. 16 (the L = 4 stage is fused by the callers)
__device__ __forceinline__ void ff_forward_stages(double2* x, const double2* tw) {
    for (int lg = 12; lg >= 4; lg -= 2) {               // L = 2^lg = 4096, 1024, 256, 64, 16
        const int L = 1 << lg, q4 = L >> 2, tstep = FF_N >> lg;
        for (int bidx = threadIdx.x; bidx < FF_N / 4; bidx += FF_THREADS) {
            const int g = bidx >> (lg - 2), j = bidx & (q4 - 1);
            const int base = (g << lg) + j;
            const cplx a0 = lds(x, base), a1 = lds(x, base + q4), a2 = lds(x, base + 2 * q4), a3 = lds(x, base + 3 * q4);
            const double2 t1 = tw[j * tstep];
            const cplx w1 = {t1.x, t1.y}, w2 = cmul(w1, w1), w3 = cmul(w2, w1);
            const cplx b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), b3 = mul_mj(csub(a1, a3));
            sts(x, base, cadd(b0, b2));
            sts(x, base + q4, cmul(cadd(b1, b3), w1));
            sts(x, base + 2 * q4, cmul(csub(b0, b2), w2));
            sts(x, base + 3 * q4, cmul(csub(b1, b3), w3));
        }
        __syncthreads();
    }
}

// inverse DIT stages L = 16 .. N (the L = 4 stage is fused by the caller); no 1/N scale here
__device__ __forceinline__ void ff_inverse_stages(double2* x, const double2* tw) {
    for (int lg = 4; lg <= 12; lg += 2) {
        const int L = 1 << lg, q4 = L >> 2, tstep = FF_N >> lg;
        for (int bidx = threadIdx.x; bidx < FF_N / 4; bidx += FF_THREADS) {
            const int g = bidx >> (lg - 2), j = bidx & (q4 - 1);
            const int base = (g << lg) + j;
            const double2 t1 = tw[j * tstep];
            const cplx w1 = {t1.x, t1.y}, w2 = cmul(w1, w1), w3 = cmul(w2, w1);
            const cplx c0 = lds(x, base), c1 = cmulc(lds(x, base + q4), w1), c2 = cmulc(lds(x, base + 2 * q4), w2),
                       c3 = cmulc(lds(x, base + 3 * q4), w3);
            const cplx e0 = cadd(c0, c2), e1 = csub(c0, c2), e2 = cadd(c1, c3), e3 = mul_pj(csub(c1, c3));
            sts(x, base, cadd(e0, e2));
            sts(x, base + q4, cadd(e1, e3));
            sts(x, base + 2 * q4, csub(e0, e2));
            sts(x, base + 3 * q4, csub(e1, e3));
        }
        __syncthreads();
    }
}

// spectrum of the zero-padded taps in the forward routine's (digit-reversed) order, scaled by 1/N
__global__ void __launch_bounds__(FF_THREADS)
fir_fft_taps_kernel(const double* __restrict__ g, int n_taps, double2* __restrict__ Grev) {
    extern __shared__ __align__(16) double2 ff_sm[];
    double2* x = ff_sm;
    double2* tw = ff_sm + FF_DATA_ELEMS;
    ff_make_twiddles(tw);
    for (int m = threadIdx.x; m < FF_N; m += FF_THREADS) sts(x, m, {m < n_taps ? g[m] : 0.0, 0.0});
    __syncthreads();
    ff_forward_stages(x, tw);
    const double inv_n = 1.0 / (double)FF_N;
    for (int bidx = threadIdx.x; bidx < FF_N / 4; bidx += FF_THREADS) {
        const int base = 4 * bidx;
        const cplx a0 = lds(x, base), a1 = lds(x, base + 1), a2 = lds(x, base + 2), a3 = lds(x, base + 3);
        const cplx b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), b3 = mul_mj(csub(a1, a3));
        const cplx y0 = cadd(b0, b2), y1 = cadd(b1, b3), y2 = csub(b0, b2), y3 = csub(b1, b3);
        Grev[base] = make_double2(y0.x * inv_n, y0.y * inv_n);
        Grev[base + 1] = make_double2(y1.x * inv_n, y1.y * inv_n);
        Grev[base + 2] = make_double2(y2.x * inv_n, y2.y * inv_n);
        Grev[base + 3] = make_double2(y3.x * inv_n, y3.y * inv_n);
    }
}

__global__ void __launch_bounds__(FF_THREADS, FF_CTAS_PER_SM)
fir_fft_kernel(const double* __restrict__ u, const double* __restrict__ y, int64_t n, int64_t lead, int n_taps,
               const double2* __restrict__ Grev, int64_t skip, double y0, int64_t n_units,
               double* __restrict__ y_pred, double* __restrict__ cta_sums) {
    extern __shared__ __align__(16) double2 ff_sm[];
    double2* x = ff_sm;
    double2* tw = ff_sm + FF_DATA_ELEMS;
    __shared__ double scratch[FF_THREADS / 32];
    const int tid = threadIdx.x;
    const int hist = n_taps - 1;
    const int V = FF_N - hist;                       // valid outputs per block
    ff_make_twiddles(tw);
    double se = 0.0, sy = 0.0, st = 0.0, cnt = 0.0;

    for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int64_t oA = unit * 2 * (int64_t)V, oB = oA + V;       // first output of blocks A and B
        __syncthreads();                                             // previous unit fully consumed / twiddles ready
        {
            // all 32 global loads of this thread are issued before the first shared-memory store, so the DRAM
            // latency is paid once per unit and not once per element
            constexpr int PER = FF_N / FF_THREADS;           // 16
            double ra[PER], rb[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int m = tid + k * FF_THREADS;
                const int64_t ia = oA - hist + m, ib = oB - hist + m;
                ra[k] = (ia >= -lead && ia < n) ? __ldg(u + ia) : 0.0;
                rb[k] = (ib >= -lead && ib < n) ? __ldg(u + ib) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < PER; ++k) sts(x, tid + k * FF_THREADS, {ra[k], rb[k]});
        }
        __syncthreads();
        ff_forward_stages(x, tw);
        // fused: last forward stage (L = 4), product with the tap spectrum, first inverse stage (L = 4)
        for (int bidx = tid; bidx < FF_N / 4; bidx += FF_THREADS) {
            const int base = 4 * bidx;
            const cplx a0 = lds(x, base), a1 = lds(x, base + 1), a2 = lds(x, base + 2), a3 = lds(x, base + 3);
            const cplx b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), b3 = mul_mj(csub(a1, a3));
            const double2 g0 = __ldg(Grev + base), g1 = __ldg(Grev + base + 1), g2 = __ldg(Grev + base + 2), g3 = __ldg(Grev + base + 3);
            const cplx c0 = cmul(cadd(b0, b2), {g0.x, g0.y}), c1 = cmul(cadd(b1, b3), {g1.x, g1.y}),
                       c2 = cmul(csub(b0, b2), {g2.x, g2.y}), c3 = cmul(csub(b1, b3), {g3.x, g3.y});
            const cplx e0 = cadd(c0, c2), e1 = csub(c0, c2), e2 = cadd(c1, c3), e3 = mul_pj(csub(c1, c3));
            sts(x, base, cadd(e0, e2));
            sts(x, base + 1, cadd(e1, e3));
            sts(x, base + 2, csub(e0, e2));
            sts(x, base + 3, csub(e1, e3));
        }
        __syncthreads();
        ff_inverse_stages(x, tw);
        // epilogue: valid part m in [hist, N) -> outputs oA + (m - hist) (real part) and oB + (m - hist) (imaginary)
        {
            constexpr int PER = FF_N / FF_THREADS;           // at most 16 valid rows per thread
            double ya[PER], yb[PER];
            if (y) {
#pragma unroll
                for (int k = 0; k < PER; ++k) {              // prefetch the measured outputs (latency paid once)
                    const int m = hist + tid + k * FF_THREADS;
                    const int64_t ia = oA + (m - hist), ib = oB + (m - hist);
                    ya[k] = (m < FF_N && ia < n) ? __ldg(y + ia) : 0.0;
                    yb[k] = (m < FF_N && ib < n) ? __ldg(y + ib) : 0.0;
                }
            }
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int m = hist + tid + k * FF_THREADS;
                if (m < FF_N) {
                    const cplx v = lds(x, m);
                    const int64_t ia = oA + (m - hist), ib = oB + (m - hist);
                    if (ia < n) {
                        if (y_pred) y_pred[ia] = v.x;
                        if (y && ia >= skip) {
                            const double yi = ya[k], e = yi - v.x, d0 = yi - y0;
                            se = fma(e, e, se); sy = fma(yi, yi, sy); st = fma(d0, d0, st); cnt += 1.0;
                        }
                    }
                    if (ib < n) {
                        if (y_pred) y_pred[ib] = v.y;
                        if (y && ib >= skip) {
                            const double yi = yb[k], e = yi - v.y, d0 = yi - y0;
                            se = fma(e, e, se); sy = fma(yi, yi, sy); st = fma(d0, d0, st); cnt += 1.0;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    se = block_sum(se, scratch);
    sy = block_sum(sy, scratch);
    st = block_sum(st, scratch);
    cnt = block_sum(cnt, scratch);
    if (tid == 0) {
        double* o = cta_sums + 4 * (size_t)blockIdx.x;
        o[0] = se; o[1] = sy; o[2] = st; o[3] = cnt;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Radix-16 organisation of the same transform: every pass keeps 16 elements of a thread in registers and runs TWO
// of the radix-4 stages above on them (identical butterflies in identical order, so the values are the ones the
// radix-4 kernel produces), which cuts the shared-memory round trips per unit from 11 to 4 and the barriers from
// 13 to 5:
//   A  : global -> registers, stages L = 4096, 1024            -> shared (natural order, padded)
//   B  : stages L = 256, 64                                     -> shared, TRANSPOSED (element 16 g + e at e * 257 + g)
//   C  : stages L = 16, 4, product with the tap spectrum, inverse stages L = 4, 16 (in place, transposed layout:
//        a thread's 16 consecutive elements are 16 rows of pitch 257, so a warp reads consecutive addresses)
//   B' : inverse stages L = 64, 256                             -> shared (natural order)
//   A' : inverse stages L = 1024, 4096 -> registers -> error sums / y_pred straight from registers
constexpr int FF_TPITCH = 257;

__device__ __forceinline__ void bf4_dif(cplx& v0, cplx& v1, cplx& v2, cplx& v3, const cplx w1) {
    const cplx w2 = cmul(w1, w1), w3 = cmul(w2, w1);
    const cplx b0 = cadd(v0, v2), b1 = csub(v0, v2), b2 = cadd(v1, v3), b3 = mul_mj(csub(v1, v3));
    v0 = cadd(b0, b2); v1 = cmul(cadd(b1, b3), w1); v2 = cmul(csub(b0, b2), w2); v3 = cmul(csub(b1, b3), w3);
}
__device__ __forceinline__ void bf4_dif_last(cplx& v0, cplx& v1, cplx& v2, cplx& v3) {
    const cplx b0 = cadd(v0, v2), b1 = csub(v0, v2), b2 = cadd(v1, v3), b3 = mul_mj(csub(v1, v3));
    v0 = cadd(b0, b2); v1 = cadd(b1, b3); v2 = csub(b0, b2); v3 = csub(b1, b3);
}
__device__ __forceinline__ void bf4_dit(cplx& v0, cplx& v1, cplx& v2, cplx& v3, const cplx w1) {
    const cplx w2 = cmul(w1, w1), w3 = cmul(w2, w1);
    const cplx c0 = v0, c1 = cmulc(v1, w1), c2 = cmulc(v2, w2), c3 = cmulc(v3, w3);
    const cplx e0 = cadd(c0, c2), e1 = csub(c0, c2), e2 = cadd(c1, c3), e3 = mul_pj(csub(c1, c3));
    v0 = cadd(e0, e2); v1 = cadd(e1, e3); v2 = csub(e0, e2); v3 = csub(e1, e3);
}
__device__ __forceinline__ void bf4_dit_first(cplx& v0, cplx& v1, cplx& v2, cplx& v3) {
    const cplx e0 = cadd(v0, v2), e1 = csub(v0, v2), e2 = cadd(v1, v3), e3 = mul_pj(csub(v1, v3));
    v0 = cadd(e0, e2); v1 = cadd(e1, e3); v2 = csub(e0, e2); v3 = csub(e1, e3);
}
__device__ __forceinline__ cplx ldtw(const double2* tw, int k) { const double2 t = tw[k]; return {t.x, t.y}; }

// tap spectrum by the same radix-16 passes (forward half only): identical values to fir_fft_taps_kernel, a third of
// its barriers -- this single-CTA launch sits serially in front of the main kernel
__global__ void __launch_bounds__(FF_THREADS)
fir_fft16_taps_kernel(const double* __restrict__ g, int n_taps, double2* __restrict__ Grev) {
    extern __shared__ __align__(16) double2 ff_sm[];
    double2* x = ff_sm;
    double2* tw = ff_sm + FF_DATA_ELEMS;
    const int tid = threadIdx.x;
    const int gB = tid >> 4, jB = tid & 15;
    cplx v[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int m = a * 1024 + b * 256 + tid;
            v[a][b] = {m < n_taps ? __ldg(g + m) : 0.0, 0.0};
        }
    ff_make_twiddles(tw);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 4; ++b) bf4_dif(v[0][b], v[1][b], v[2][b], v[3][b], ldtw(tw, b * 256 + tid));
    {
        const cplx w = ldtw(tw, 4 * tid);
#pragma unroll
        for (int a = 0; a < 4; ++a) bf4_dif(v[a][0], v[a][1], v[a][2], v[a][3], w);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) sts(x, a * 1024 + b * 256 + tid, v[a][b]);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) v[a][b] = lds(x, gB * 256 + a * 64 + b * 16 + jB);
#pragma unroll
    for (int b = 0; b < 4; ++b) bf4_dif(v[0][b], v[1][b], v[2][b], v[3][b], ldtw(tw, (b * 16 + jB) * 16));
    {
        const cplx w = ldtw(tw, jB * 64);
#pragma unroll
        for (int a = 0; a < 4; ++a) bf4_dif(v[a][0], v[a][1], v[a][2], v[a][3], w);
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) x[jB * FF_TPITCH + 16 * gB + 4 * a + b] = make_double2(v[a][b].x, v[a][b].y);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { const double2 t = x[(4 * a + b) * FF_TPITCH + tid]; v[a][b] = {t.x, t.y}; }
#pragma unroll
    for (int b = 0; b < 4; ++b) bf4_dif(v[0][b], v[1][b], v[2][b], v[3][b], ldtw(tw, b * 256));
    const double inv_n = 1.0 / (double)FF_N;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        bf4_dif_last(v[a][0], v[a][1], v[a][2], v[a][3]);
#pragma unroll
        for (int b = 0; b < 4; ++b) Grev[16 * tid + 4 * a + b] = make_double2(v[a][b].x * inv_n, v[a][b].y * inv_n);
    }
}

__global__ void __launch_bounds__(FF_THREADS, FF_CTAS_PER_SM)
fir_fft16_kernel(const double* __restrict__ u, const double* __restrict__ y, int64_t n, int64_t lead, int n_taps,
                 const double2* __restrict__ Grev, int64_t skip, double y0, int64_t n_units,
                 double* __restrict__ y_pred, double* __restrict__ cta_sums) {
    extern __shared__ __align__(16) double2 ff_sm[];
    double2* x = ff_sm;
    double2* tw = ff_sm + FF_DATA_ELEMS;
    __shared__ double scratch[FF_THREADS / 32];
    const int tid = threadIdx.x;
    const int hist = n_taps - 1;
    const int V = FF_N - hist;                       // valid outputs per block
    ff_make_twiddles(tw);
    double se = 0.0, sy = 0.0, st = 0.0, cnt = 0.0;
    const int gB = tid >> 4, jB = tid & 15;          // pass B / B': group of 256, offset inside the group's 16-blocks
    cplx v[4][4];

    // ---- TMA staging of the input window.  A unit reads 2V + hist CONSECUTIVE samples of u (block B's window is block
    // A's shifted by V).  The FFT buffer x is idle from the moment pass A' has read it until the next unit's pass A
    // writes it, i.e. for the whole epilogue: one thread sends the NEXT unit's window into it with a bulk copy
    // (cp.async.bulk, completion on an mbarrier), and pass A then takes its 32 samples per thread from shared memory
    // instead of issuing 32 global loads whose latency nothing covers at 16 warps per SM.  Units whose window leaves
    // the record (first / last) keep the predicated global loads.
    __shared__ unsigned long long ff_bar;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&ff_bar);
    double* rawd = reinterpret_cast<double*>(x);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // window of `unit` as [g0, g0 + len): g0 = sA - sh with sh in {0, 1} making u + g0 16-byte aligned, len even
    auto stage_next = [&](int64_t unit_n, int& sh_out) -> bool {
        if (unit_n >= n_units) return false;
        const int64_t sA_n = unit_n * 2 * (int64_t)V - hist;
        const int sh = (int)((reinterpret_cast<uintptr_t>(u + sA_n) >> 3) & 1);
        const int64_t g0 = sA_n - sh;
        const int len = (2 * V + hist + sh + 1) & ~1;
        if (g0 < -lead || g0 + len > n) return false;
        sh_out = sh;
        if (tid == 0) {
            const unsigned bytes = (unsigned)len * (unsigned)sizeof(double);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((unsigned)__cvta_generic_to_shared(rawd)), "l"(u + g0), "r"(bytes), "r"(bar) : "memory");
        }
        return true;
    };
    int sh_cur = 0;
    unsigned bar_parity = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bool staged = stage_next(blockIdx.x, sh_cur);

    for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int64_t oA = unit * 2 * (int64_t)V, oB = oA + V;       // first output of blocks A and B
        // ---- pass A: element a*1024 + b*256 + tid of z = uA + j uB straight from global memory.  Index windows as
        // ints relative to the block start, so that every load is base pointer + immediate offset
        const int64_t sA = oA - hist, sB = oB - hist;
        const int loA = (int)max((int64_t)0, min((int64_t)FF_N, -lead - sA)), hiA = (int)max((int64_t)0, min((int64_t)FF_N, n - sA));
        const int loB = (int)max((int64_t)0, min((int64_t)FF_N, -lead - sB)), hiB = (int)max((int64_t)0, min((int64_t)FF_N, n - sB));
        const int skA = (int)max((int64_t)hist, min((int64_t)FF_N, skip - sA)), skB = (int)max((int64_t)hist, min((int64_t)FF_N, skip - sB));
        {
            // L2 prefetch, one 128-byte line per thread and step: this unit's measured outputs (read in the epilogue)
            // and the NEXT unit's input window [sA', sA' + 2V + hist), so that neither is a DRAM round trip when the
            // 16 warps of an SM get to it
            const int64_t nA = sA + (int64_t)gridDim.x * 2 * V;
            for (int ln = tid; ln * 16 < 2 * V + hist + 16; ln += FF_THREADS) {
                const int64_t iu = nA + (int64_t)ln * 16;
                if (iu >= -lead && iu < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(u + iu));
                const int64_t iy = oA + (int64_t)ln * 16;
                if (y && iy < n && ln * 16 < 2 * V + 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + iy));
            }
            if (staged) {
                // the window landed in x during the previous unit's epilogue: block A at rawd[sh + m], block B V further
                unsigned ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(bar), "r"(bar_parity) : "memory");
                bar_parity ^= 1u;
                const double* pa = rawd + sh_cur + tid;
                const double* pb = pa + V;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        v[a][b].x = pa[a * 1024 + b * 256];
                        v[a][b].y = pb[a * 1024 + b * 256];
                    }
            } else {
                const double* pa = u + sA + tid;
                const double* pb = u + sB + tid;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int m = a * 1024 + b * 256 + tid;
                        v[a][b].x = (m >= loA && m < hiA) ? __ldg(pa + (a * 1024 + b * 256)) : 0.0;
                        v[a][b].y = (m >= loB && m < hiB) ? __ldg(pb + (a * 1024 + b * 256)) : 0.0;
                    }
            }
        }
        __syncthreads();                                             // twiddles ready / previous unit's pass A' has read x
#pragma unroll
        for (int b = 0; b < 4; ++b) bf4_dif(v[0][b], v[1][b], v[2][b], v[3][b], ldtw(tw, b * 256 + tid));
        {
            const cplx w = ldtw(tw, 4 * tid);
#pragma unroll
            for (int a = 0; a < 4; ++a) bf4_dif(v[a][0], v[a][1], v[a][2], v[a][3], w);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) sts(x, a * 1024 + b * 256 + tid, v[a][b]);
        __syncthreads();
        // ---- pass B: element g*256 + a*64 + b*16 + j
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) v[a][b] = lds(x, gB * 256 + a * 64 + b * 16 + jB);
#pragma unroll
        for (int b = 0; b < 4; ++b) bf4_dif(v[0][b], v[1][b], v[2][b], v[3][b], ldtw(tw, (b * 16 + jB) * 16));
        {
            const cplx w = ldtw(tw, jB * 64);
#pragma unroll
            for (int a = 0; a < 4; ++a) bf4_dif(v[a][0], v[a][1], v[a][2], v[a][3], w);
        }
        __syncthreads();                                             // every thread has read its pass-B inputs
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) x[jB * FF_TPITCH + 16 * gB + 4 * a + b] = make_double2(v[a][b].x, v[a][b].y);
        __syncthreads();
        // ---- pass C: elements 16 tid + 4a + b (transposed layout), spectrum product in the middle
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) { const double2 t = x[(4 * a + b) * FF_TPITCH + tid]; v[a][b] = {t.x, t.y}; }
#pragma unroll
        for (int b = 0; b < 4; ++b) bf4_dif(v[0][b], v[1][b], v[2][b], v[3][b], ldtw(tw, b * 256));
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            bf4_dif_last(v[a][0], v[a][1], v[a][2], v[a][3]);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double2 g = __ldg(Grev + 16 * tid + 4 * a + b);
                v[a][b] = cmul(v[a][b], {g.x, g.y});
            }
            bf4_dit_first(v[a][0], v[a][1], v[a][2], v[a][3]);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) bf4_dit(v[0][b], v[1][b], v[2][b], v[3][b], ldtw(tw, b * 256));
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) x[(4 * a + b) * FF_TPITCH + tid] = make_double2(v[a][b].x, v[a][b].y);
        __syncthreads();
        // ---- pass B': inverse stages L = 64 (over b), 256 (over a)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) { const double2 t = x[jB * FF_TPITCH + 16 * gB + 4 * a + b]; v[a][b] = {t.x, t.y}; }
        {
            const cplx w = ldtw(tw, jB * 64);
#pragma unroll
            for (int a = 0; a < 4; ++a) bf4_dit(v[a][0], v[a][1], v[a][2], v[a][3], w);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) bf4_dit(v[0][b], v[1][b], v[2][b], v[3][b], ldtw(tw, (b * 16 + jB) * 16));
        __syncthreads();                                             // every thread has read its pass-B' inputs
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) sts(x, gB * 256 + a * 64 + b * 16 + jB, v[a][b]);
        __syncthreads();
        // ---- pass A': inverse stages L = 1024 (over b), 4096 (over a); results stay in registers
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) v[a][b] = lds(x, a * 1024 + b * 256 + tid);
        // x is free now: once every thread has its elements, the next unit's window goes into it (the plain stores of
        // the passes precede the bulk copy's writes: cross-proxy fence, then the barrier)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        staged = stage_next(unit + gridDim.x, sh_cur);
        {
            const cplx w = ldtw(tw, 4 * tid);
#pragma unroll
            for (int a = 0; a < 4; ++a) bf4_dit(v[a][0], v[a][1], v[a][2], v[a][3], w);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) bf4_dit(v[0][b], v[1][b], v[2][b], v[3][b], ldtw(tw, b * 256 + tid));
        // ---- epilogue: m = a*1024 + b*256 + tid; valid part m in [hist, N) -> outputs sA + m (real part), sB + m
        // (imaginary part); the sums take m >= skA / skB (>= hist, and past the transient skip)
        {
            const double* qa = y + sA + tid;
            const double* qb = y + sB + tid;
            double* ra = y_pred + sA + tid;
            double* rb = y_pred + sB + tid;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double ya[4], yb[4];
                if (y) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int m = a * 1024 + b * 256 + tid;
                        ya[b] = (m >= skA && m < hiA) ? __ldg(qa + (a * 1024 + b * 256)) : 0.0;
                        yb[b] = (m >= skB && m < hiB) ? __ldg(qb + (a * 1024 + b * 256)) : 0.0;
                    }
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int m = a * 1024 + b * 256 + tid;
                    if (y_pred) {
                        if (m >= hist && m < hiA) ra[a * 1024 + b * 256] = v[a][b].x;
                        if (m >= hist && m < hiB) rb[a * 1024 + b * 256] = v[a][b].y;
                    }
                    if (y) {
                        if (m >= skA && m < hiA) {
                            const double yi = ya[b], e = yi - v[a][b].x, d0 = yi - y0;
                            se = fma(e, e, se); sy = fma(yi, yi, sy); st = fma(d0, d0, st); cnt += 1.0;
                        }
                        if (m >= skB && m < hiB) {
                            const double yi = yb[b], e = yi - v[a][b].y, d0 = yi - y0;
                            se = fma(e, e, se); sy = fma(yi, yi, sy); st = fma(d0, d0, st); cnt += 1.0;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    se = block_sum(se, scratch);
    sy = block_sum(sy, scratch);
    st = block_sum(st, scratch);
    cnt = block_sum(cnt, scratch);
    if (tid == 0) {
        double* o = cta_sums + 4 * (size_t)blockIdx.x;
        o[0] = se; o[1] = sy; o[2] = st; o[3] = cnt;
    }
}

constexpr size_t FF_SMEM_BYTES = (size_t)(FF_DATA_ELEMS + FF_N / 4) * sizeof(double2);

}  // namespace b200id

using namespace b200id;

size_t b200id_fir_fft_workspace_bytes() {
    return align_up((size_t)FF_N * sizeof(double2), 256) + align_up((size_t)(sm_count() * FF_CTAS_PER_SM + 8) * 4 * sizeof(double), 256);
}
bool b200id_fir_fft_applicable(int64_t n, int n_taps) {
    return n_taps >= 128 && n_taps <= FF_MAX_TAPS && n >= 32768;
}

// organisation of the overlap-save transform: 0 = radix-16 passes (default), 1 = radix-4 passes; -1 = not yet read
// from the environment (B200ID_FIR_RADIX4=1)
static int g_fir_radix4 = -1;
extern "C" int b200id_fir_select_fft(int radix4) {
    const int prev = g_fir_radix4 == 1 ? 1 : 0;
    g_fir_radix4 = radix4 < 0 ? -1 : (radix4 ? 1 : 0);
    return prev;
}

// fir_fold_kernel lives in fir_conv.cu
namespace b200id { __global__ void fir_fold_kernel(const double* __restrict__ part, int64_t n_cta, double* __restrict__ out); }

int b200id_fir_eval_fft(const double* u, const double* y, int64_t n, int64_t lead, const double* g, int n_taps,
                        int64_t skip, double y0, double* y_pred_out, double* sums_out, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
    B200ID_REQUIRE(ws_bytes >= b200id_fir_fft_workspace_bytes(), B200ID_E_WORKSPACE, "fir_eval(fft): workspace too small");
    double2* Grev = (double2*)ws;
    double* cta_sums = (double*)((char*)ws + align_up((size_t)FF_N * sizeof(double2), 256));
    int rc = check_cuda(cudaFuncSetAttribute(fir_fft_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM_BYTES), "cudaFuncSetAttribute(fir fft taps)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(fir_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM_BYTES), "cudaFuncSetAttribute(fir fft)");
    if (rc) return rc;
    if (g_fir_radix4 < 0) { const char* e = getenv("B200ID_FIR_RADIX4"); g_fir_radix4 = (e && e[0] == '1') ? 1 : 0; }
    if (g_fir_radix4) {
        fir_fft_taps_kernel<<<1, FF_THREADS, FF_SMEM_BYTES, st>>>(g, n_taps, Grev);
        B200ID_LAUNCH_CHECK("fir_fft_taps_kernel");
    } else {
        rc = check_cuda(cudaFuncSetAttribute(fir_fft16_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM_BYTES), "cudaFuncSetAttribute(fir fft16 taps)");
        if (rc) return rc;
        fir_fft16_taps_kernel<<<1, FF_THREADS, FF_SMEM_BYTES, st>>>(g, n_taps, Grev);
        B200ID_LAUNCH_CHECK("fir_fft16_taps_kernel");
    }
    const int V = FF_N - (n_taps - 1);
    const int64_t n_units = (n + 2 * (int64_t)V - 1) / (2 * (int64_t)V);
    int grid = sm_count() * FF_CTAS_PER_SM;
    if (n_units < grid) grid = (int)n_units;
    if (g_fir_radix4) {
        fir_fft_kernel<<<grid, FF_THREADS, FF_SMEM_BYTES, st>>>(u, y, n, lead, n_taps, Grev, skip, y0, n_units, y_pred_out, cta_sums);
        B200ID_LAUNCH_CHECK("fir_fft_kernel");
    } else {
        rc = check_cuda(cudaFuncSetAttribute(fir_fft16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM_BYTES), "cudaFuncSetAttribute(fir fft16)");
        if (rc) return rc;
        fir_fft16_kernel<<<grid, FF_THREADS, FF_SMEM_BYTES, st>>>(u, y, n, lead, n_taps, Grev, skip, y0, n_units, y_pred_out, cta_sums);
        B200ID_LAUNCH_CHECK("fir_fft16_kernel");
    }
    fir_fold_kernel<<<1, 256, 0, st>>>(cta_sums, grid, sums_out);
    B200ID_LAUNCH_CHECK("fir_fold_kernel");
    return B200ID_OK;
}
