// frf_adaptive.cuh -- K1 uniform-timebase projection, bin-adaptive (included by frf_project.cu; reference
// src/frequency_transform/frf_estimator.py:226-231 for u and y of one record).
//
// The reference forms the phase as the rounded product fl(w_k t_i).  With theta_ki = w_k (t0 + i dt) exactly,
//     fl(w_k t_i) = theta_ki + eps_ki,   eps_ki = w_k delta_i - e_ki,
//     delta_i = t_i - (t0 + i dt),       e_ki = w_k t_i - fl(w_k t_i)      (the residual of the product, exact by FMA)
//     sum_i a_i exp(-j fl(w_k t_i)) = sum_i a_i exp(-j theta_ki)  -  j sum_i a_i eps_ki exp(-j theta_ki)  + O(eps^2)
//            ^ what the reference computes      ^ MAIN term                  ^ ROUNDING term
// The main term is a matrix product (below) and costs one FP64 FMA per sample, bin and output; the rounding term needs
// the residual of every (sample, bin) product -- three more FP64-pipe instructions plus an FP32 recurrence -- and it
// was what bounded the all-recurrence kernels (frf_project_uniform_g_kernel: 7 FP64 instructions per sample-bin).
// It only MATTERS where a bin is weakly excited: e_ki is rounding noise of size ulp(w_k t), so the rounding term is a
// random walk of standard deviation  sigma_k ~ ulp(w_k t_max) ||a||_2 / sqrt(12),  to be compared with |U_k|.
//
//   0. frf_timebase_fit_kernel   least-squares line through (every 64th of) the time stamps: the caller's (t0, dt) are FP64 numbers, and
//                                dt = span / (n - 1) is off by up to half an ulp of t_end over the record -- a RAMP in
//                                delta_i that is coherent with every excited bin (w_k ulp(t_end) / 4 ~ 1e-9 rad at the top
//                                bin of a 1e7-sample record).  The fit's (alpha, beta) extend (t0, dt) to double-double,
//                                so the ramp is part of the exact phase theta_ki and delta_i is what remains.
//   1. frf_main_ws_kernel        main term of ALL bins on the FP64 tensor pipe (DMMA): warp-specialised, raw samples by
//      (frf_main_dmma_kernel      TMA bulk copies (the plain build where the arrays are not 16-byte aligned);
//       = plain build)           also ||a||^2 and the moments of delta that the decision needs.
//   2. frf_fold_kernel           per-CTA partials -> sums (fixed order)
//   3. frf_decide_kernel         per bin, on the device: the rounding term is SKIPPED when
//                                    DA_SIGMAS sigma_k + rigorous_k  <=  DA_BUDGET |S_k|     for u and for y,
//                                sigma_k^2 = ulp(w_k t_max)^2/12 ||a||^2 + w_k^2 sum a_i^2 delta_i^2 (delta_i at rounding level)
//                                rigorous_k = w_k sum |a_i delta_i|                    (delta_i beyond rounding level: a real
//                                                                                       timebase deviation, never treated as noise)
//                                and the other bins are compacted into a list (ascending, deterministic).
//   4. frf_round_goertzel_kernel rounding term of the listed bins only (packed-FP32 Goertzel recurrence, u and y
//                                together), geometry from the device-side count: no host read in between.
//   5. frf_fold_add_kernel       adds the listed bins' rounding terms to the sums (fixed order).
//
// On a multisine record every INPUT bin is excited; what gets listed there are the output's bins above the plant's corner
// (|Y_k| rolls off as 1/f^4), about a third of the grid; on the square-wave record about half.  If the list is empty
// step 4 returns at once.  DA_BUDGET = 4e-10 leaves the
// 1e-9 per-bin parity of U and Y (and of G = Y/U) intact: a skipped rounding term above the budget needs a
// DA_SIGMAS = 4 sigma event, one above 1e-9 a 10 sigma event (measured |rounding term| / sigma_k <= 2 over the test
// records, sigma_k itself an over-estimate: it takes ulp(w_k t) at the END of the record for every sample).  What the noise model cannot
// exclude is a sawtooth of the rounding errors that lines up with a difference of two excitation lines to within
// 1/n; its size is then <= w_k ulp(t) / pi ~ 1e-9, i.e. still of the order of the tolerance.
//
// Main term as matrix products: samples in blocks of 32, i = i0 + 32 b + r,
//     sum_i a_i (cos, sin)(theta_ki) = sum_b Rot(Theta_kb) [ sum_r a_{b,r} (cos, sin)(w_k r dt) ],
// the bracket a dense [blocks x 32] x [32 x 2 bins] product on mma.sync.m8n8k4.f64: one DMMA = 256 FMAs from ONE issue
// slot.  A warp owns (4-bin column group, signal) units; per k-step one B fragment (table Qt) and the A fragments of
// ALL block groups of the tile are loaded and the block groups' DMMAs issue back to back, so a DMMA never waits for
// its accumulator.  The block phasors (tile seed x row table x block-group phasor, all exact) are applied by Horner over
// the block groups (one phasor per bin), the row phasor once, the tile seed after the three-shuffle row sum.
//
// Deterministic for a given grid: every accumulator entry has exactly one owner and a fixed order.
#pragma once

namespace b200id {

constexpr int DA_THREADS = 512;
constexpr int DA_WARPS = DA_THREADS / 32;
constexpr int DA_LD = 36;                        // padded row length (doubles) of the A and Qt rows: conflict-free fragments
#ifndef DA_TILE
#define DA_TILE 1536
#endif
constexpr int DA_TS = DA_TILE;                   // samples per tile (multiple of 256 = one block group of 8 blocks of 32)
constexpr int DA_NB = DA_TS / 32;                // blocks per tile
constexpr int DA_NBG = DA_NB / 8;                // block groups (8 blocks = one m-tile) per tile
constexpr int DA_RESEED = 64;                    // tiles between exact re-seeds of the tile phasor
constexpr int DA_NORMS = 6;                      // {sum au^2, sum ay^2, sum au^2 ds^2, sum ay^2 ds^2, sum |au dr|, sum |ay dr|}
constexpr size_t DA_SMEM_LIMIT = 115712;         // (228 KB - 2 x 1 KB) / 2: two CTAs per SM
constexpr int DA_BPB = 128;                      // bins per CTA pass of the rounding-term kernel (stripes R >= 8)
constexpr double DA_MAGIC = 4503599627370496.0 + 1262485504.0;   // 2^52 + 0x4B400000: low word of (DA_MAGIC + n) = bits of the float 2^23 + 2^22 + n
constexpr float DA_MAGIC_F = 12582912.0f;                        // 2^23 + 2^22
constexpr int DA_CM = DA_TS + 2 * DA_THREADS;    // entries of the chain-major arrays (R (ceil(TS / R) | 1) < TS + 2 R)
#ifndef DA_ROUND_PREFETCH
#define DA_ROUND_PREFETCH 1
#endif
#ifndef DA_SIGMAS
#define DA_SIGMAS 4.0
#endif
#ifndef DA_BUDGET
#define DA_BUDGET 4e-10
#endif
static_assert(DA_TS % 256 == 0, "tile = whole block groups");
static_assert(DA_TS % DA_THREADS == 0, "every thread stages the same number of samples");

__host__ __device__ inline int da_pad4(int nb) { return (nb + 3) & ~3; }
static inline size_t da_main_smem_bytes(int bins_per_group) {
    const size_t bp = (size_t)da_pad4(bins_per_group);
    return (size_t)2 * bp * DA_LD * sizeof(double)              // Qt
           + (size_t)2 * DA_NB * DA_LD * sizeof(double)         // A rows (u blocks, y blocks)
           + (size_t)8 * bp * sizeof(double2)                   // P1
           + (size_t)bp * sizeof(double2)                       // R8
           + (size_t)2 * bp * sizeof(double2)                   // Pt0, RT
           + 2 * sizeof(double2)                                // (t, delta) of the first sample of this tile and the next
           + (size_t)4 * bp * sizeof(double)                    // main-term sums
           + bp * sizeof(double);                               // w
}
// smallest number of bin groups whose CTA fits two to an SM
static inline int da_main_groups(int nd) {
    int g = 1;
    while (da_main_smem_bytes((nd + g - 1) / g) > DA_SMEM_LIMIT) ++g;
    return g;
}
static inline size_t da_round_smem_bytes() {
    return (size_t)DA_CM * (sizeof(double) + sizeof(float2) + sizeof(float)) + (size_t)DA_BPB * (sizeof(double) + sizeof(float2))
           + 2 * sizeof(double2);
}
// geometry of the rounding-term pass for nf listed bins on a grid of n_cta CTAs (host and device agree by construction)
struct DaRoundGeom { int groups, bpg, gx; };
__host__ __device__ inline DaRoundGeom da_round_geom(int nf, int n_cta) {
    DaRoundGeom g;
    g.groups = (nf + DA_BPB - 1) / DA_BPB;
    if (g.groups < 1) g.groups = 1;
    g.bpg = (nf + g.groups - 1) / g.groups;
    g.gx = n_cta / g.groups;
    return g;
}

__device__ __forceinline__ void da_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// Time base t0 + i dt in double-double: (t0, dt) from the caller, the low words from the least-squares line through
// delta_i = t_i - (t0 + i dt) over the SAMPLED stamps i = begin, begin + DA_FIT_STRIDE, ...:
// fit = {sum delta, sum (i - c) delta, m, sum (i - c), sum (i - c)^2}, c = (begin + end - 1) / 2.
constexpr int DA_FIT_STRIDE = 64;
constexpr int DA_FIT_SUMS = 5;
struct DaBase { double t0h, t0l, dth, dtl; };
__device__ __forceinline__ DaBase da_base(const double* __restrict__ fit, double t0, double dt, int64_t begin, int64_t end) {
    DaBase b; b.t0h = t0; b.t0l = 0.0; b.dth = dt; b.dtl = 0.0;
    if (fit) {
        const double c = 0.5 * (double)(begin + end - 1);
        const double sd = fit[0], sxd = fit[1], m = fit[2], sx = fit[3], sxx = fit[4];
        const double det = m * sxx - sx * sx;
        double alpha = m > 0.0 ? sd / m : 0.0, beta = 0.0;
        if (m > 1.5 && det > 0.0) {
            beta = (m * sxd - sx * sd) / det;
            alpha = (sd - beta * sx) / m;
        }
        b.t0l = alpha - beta * c; b.dtl = beta;
    }
    return b;
}
// t0 + k dt as an unevaluated sum (k integer-valued)
__device__ __forceinline__ DD da_time(const DaBase& b, double k) {
    const double ph = k * b.dth, pl = fma(k, b.dth, -ph) + k * b.dtl;
    const double sh = b.t0h + ph, bb = sh - b.t0h;
    const double sl = (b.t0h - (sh - bb)) + (ph - bb);
    DD r; r.hi = sh; r.lo = (sl + pl) + b.t0l; return r;
}
// (cos, sin)(w * m * dt) for an integer m, through the double-double product m dt
__device__ __forceinline__ void da_sincos_steps(double w, double m, const DaBase& b, double& s, double& c) {
    DD st; st.hi = m * b.dth; st.lo = fma(m, b.dth, -st.hi) + m * b.dtl;
    sincos_dd(w, st, s, c);
}
// 2^(exponent(x) - 52): the spacing of doubles at |x| (clamped at the bottom of the normal range)
__device__ __forceinline__ double da_ulp(double x) {
    int e = __double2hiint(x) & 0x7ff00000;
    e = max(e, 53 << 20) - (52 << 20);
    return __hiloint2double(e, 0);
}
// (t, delta) of sample i: its time stamp and the deviation from the double-double grid (t_i - hi is exact: the two
// are within a few ulp of each other)
__device__ __forceinline__ double2 da_tile_seed(const double* __restrict__ t, int64_t i, const DaBase& b) {
    const double ti = __ldg(t + i);
    const DD x = da_time(b, (double)i);
    return make_double2(ti, (ti - x.hi) - x.lo);
}
// one sample of a tile, j samples after the tile's first (seed = da_tile_seed of that one): trapezoid weight x
// (signal - mean) for u and y, and delta_i = delta_i0 + (t_i - t_i0) - j dt: the difference of two nearby stamps is
// exact, the product j dt enters through one FMA whose result is O(delta) -- three FP64 instructions instead of the
// double-double sum of da_time
struct DaSample { double ti, au, ay, dl; };
__device__ __forceinline__ DaSample da_sample(const double* __restrict__ t, const double* __restrict__ u,
                                              const double* __restrict__ y, int64_t i, int64_t n, double jd,
                                              double2 seed, const DaBase& b, double mean_u, double mean_y) {
    DaSample s;
    s.ti = __ldg(t + i);
    const double tm = __ldg(t + (i > 0 ? i - 1 : 0)), tp = __ldg(t + (i < n - 1 ? i + 1 : n - 1));
    const double wt = 0.5 * (tp - tm);
    s.dl = fma(-jd, b.dtl, fma(-jd, b.dth, s.ti - seed.x)) + seed.y;
    s.au = wt * (__ldg(u + i) - mean_u);
    s.ay = wt * (__ldg(y + i) - mean_y);
    return s;
}

// ------------------------------------------------------------------------------------------------------------------
// 0. sums for the least-squares line through the stamps against the caller's (t0, dt).  Every DA_FIT_STRIDE-th stamp is
// enough (one 32-byte sector out of four 128-byte lines): the line is determined to a small fraction of an ulp by 1e5+ points, and whatever line comes out, delta is
// taken against it consistently afterwards (a quarter of the DRAM sectors of t instead of all of them)
constexpr int DA_FIT_THREADS = 256;
__global__ void __launch_bounds__(DA_FIT_THREADS)
frf_timebase_fit_kernel(const double* __restrict__ t, int64_t begin, int64_t end, double t0, double dt, int64_t chunk,
                        double* __restrict__ cta_partial) {
    __shared__ double scratch[DA_FIT_THREADS / 32];
    const int64_t lo = begin + (int64_t)blockIdx.x * chunk, hi = min(lo + chunk, end);        // chunk is a multiple of the stride
    const double c = 0.5 * (double)(begin + end - 1);
    DaBase b; b.t0h = t0; b.t0l = 0.0; b.dth = dt; b.dtl = 0.0;
    double s[DA_FIT_SUMS] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t i = lo + (int64_t)threadIdx.x * DA_FIT_STRIDE; i < hi; i += (int64_t)DA_FIT_THREADS * DA_FIT_STRIDE) {
        const DD x = da_time(b, (double)i);
        const double d = (__ldg(t + i) - x.hi) - x.lo;
        const double xi = (double)i - c;
        s[0] += d;
        s[1] = fma(xi, d, s[1]);
        s[2] += 1.0;
        s[3] += xi;
        s[4] = fma(xi, xi, s[4]);
    }
#pragma unroll
    for (int q = 0; q < DA_FIT_SUMS; ++q) {
        const double v = block_sum(s[q], scratch);
        if (threadIdx.x == 0) cta_partial[DA_FIT_SUMS * blockIdx.x + q] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 1. main term of all bins (grid.y = bin groups, grid.x CTAs stride over the tiles)
__global__ void __launch_bounds__(DA_THREADS, 2)
frf_main_dmma_kernel(const double* __restrict__ t, const double* __restrict__ u, const double* __restrict__ y,
                     int64_t n, int64_t own_begin, int64_t own_end, const double* __restrict__ w, int nd,
                     const double* __restrict__ sums, double inv_count, double t0, double dt,
                     const double* __restrict__ fit, int n_tiles, int bins_per_group, double* __restrict__ cta_partial,
                     double* __restrict__ norm_partial) {
    extern __shared__ __align__(16) unsigned char da_smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = (int)__reduce_max_sync(0xffffffffu, (unsigned)(tid >> 5));      // uniform register
    const int bin0 = blockIdx.y * bins_per_group;
    const int nb_all = min(nd - bin0, bins_per_group);
    const int bp = da_pad4(nb_all);
    const int ncg = bp >> 2;

    double* Qt = reinterpret_cast<double*>(da_smem);                 // [2 bp][DA_LD]  row 2k: cos(w_k r dt), row 2k+1: sin
    double* As = Qt + (size_t)2 * bp * DA_LD;                        // [2 DA_NB][DA_LD]  u blocks then y blocks
    double2* P1 = reinterpret_cast<double2*>(As + (size_t)2 * DA_NB * DA_LD);   // [8][bp]   (cos, sin)(w_k 32 g dt)
    double2* R8 = P1 + (size_t)8 * bp;                               // [bp]      (cos, sin)(w_k 256 dt): block group to block group
    double2* Pt0 = R8 + bp;                                          // [bp]      (cos, sin)(w_k (t0 + i0 dt)) of the current tile
    double2* RT = Pt0 + bp;                                          // [bp]      tile-to-tile rotation
    double2* seeds = RT + bp;                                        // [2]       (t, delta) of the first sample of tile it, it + 1
    double* accM = reinterpret_cast<double*>(seeds + 2);             // [4][bp]   main-term sums {uc, us, yc, ys}
    double* ws = accM + (size_t)4 * bp;                              // [bp]      w_k (0 for the padding bins)

    const double mean_u = sums ? sums[0] * inv_count : 0.0;
    const double mean_y = sums ? sums[1] * inv_count : 0.0;
    const bool do_norms = blockIdx.y == 0 && norm_partial != nullptr;
    const DaBase tb = da_base(fit, t0, dt, own_begin, own_end);
    if (nb_all <= 0) return;

    for (int k = tid; k < bp; k += DA_THREADS) ws[k] = k < nb_all ? w[bin0 + k] : 0.0;
    __syncthreads();
    // ---- exact tables (once per CTA)
    for (int e = tid; e < bp * 32; e += DA_THREADS) {
        const int k = e >> 5, rr = e & 31;
        double s, c;
        da_sincos_steps(ws[k], (double)rr, tb, s, c);
        Qt[(size_t)(2 * k) * DA_LD + rr] = c;
        Qt[(size_t)(2 * k + 1) * DA_LD + rr] = s;
    }
    for (int e = tid; e < bp * 9; e += DA_THREADS) {
        const int g = e / bp, k = e - g * bp;
        double s, c;
        da_sincos_steps(ws[k], (double)(32 * g), tb, s, c);
        if (g < 8) P1[g * bp + k] = make_double2(c, s);
        else R8[k] = make_double2(c, s);
    }
    if (tid == DA_THREADS - 1 && (int)blockIdx.x < n_tiles)
        seeds[0] = da_tile_seed(t, own_begin + (int64_t)blockIdx.x * DA_TS, tb);
    for (int k = tid; k < bp; k += DA_THREADS) {
        double s, c;
        da_sincos_steps(ws[k], (double)gridDim.x * (double)DA_TS, tb, s, c);
        RT[k] = make_double2(c, s);
        accM[k] = 0.0; accM[bp + k] = 0.0; accM[2 * bp + k] = 0.0; accM[3 * bp + k] = 0.0;
    }

    double nrm[DA_NORMS];
#pragma unroll
    for (int c = 0; c < DA_NORMS; ++c) nrm[c] = 0.0;

    int it = 0;
    for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++it) {
        const int64_t i0 = own_begin + (int64_t)tl * DA_TS;
        const int cnt = (int)min((int64_t)DA_TS, own_end - i0);
        __syncthreads();                                                 // previous tile consumed (first pass: tables written)
        const double2 seed = seeds[it & 1];
        // ---- stage the tile as block rows
#pragma unroll
        for (int j = tid; j < DA_TS; j += DA_THREADS) {
            double au = 0.0, ay = 0.0;
            if (j < cnt) {
                const DaSample s = da_sample(t, u, y, i0 + j, n, (double)j, seed, tb, mean_u, mean_y);
                au = s.au; ay = s.ay;
                if (do_norms) {
                    const double ad = fabs(s.dl);
                    const bool noise = ad <= da_ulp(s.ti);               // a rounded grid point: |delta| <= 1 ulp of t
                    const double ds = noise ? ad : 0.0, dr = noise ? 0.0 : ad;
                    const double u2 = au * au, y2 = ay * ay, ds2 = ds * ds;
                    nrm[0] += u2; nrm[1] += y2;
                    nrm[2] = fma(u2, ds2, nrm[2]); nrm[3] = fma(y2, ds2, nrm[3]);
                    nrm[4] = fma(fabs(au), dr, nrm[4]); nrm[5] = fma(fabs(ay), dr, nrm[5]);
                }
            }
            const int b = j >> 5, rr = j & 31;
            As[(size_t)b * DA_LD + rr] = au;
            As[(size_t)(DA_NB + b) * DA_LD + rr] = ay;
        }
        // ---- phasor of the tile start: exact double-double seed, advanced by the exact tile-to-tile rotation in between
        if (tid < bp) {
            if (it % DA_RESEED == 0) {
                const DD x = da_time(tb, (double)i0);
                double s, c;
                sincos_dd(ws[tid], x, s, c);
                Pt0[tid] = make_double2(c, s);
            } else {
                const double2 p = Pt0[tid], q = RT[tid];
                Pt0[tid] = make_double2(fma(p.x, q.x, -(p.y * q.y)), fma(p.y, q.x, p.x * q.y));
            }
        } else if (tid == DA_THREADS - 1 && tl + (int)gridDim.x < n_tiles) {
            seeds[(it + 1) & 1] = da_tile_seed(t, i0 + (int64_t)gridDim.x * DA_TS, tb);      // read after the next loop-top barrier
        }
        __syncthreads();

        // ---- block-phasor matrix products: unit = (4-bin column group, signal)
        const int g4 = lane >> 2, l4 = lane & 3;
        for (int unit = warp; unit < 2 * ncg; unit += DA_WARPS) {
            const int sig = unit & 1, cg = unit >> 1;
            const int k = cg * 4 + l4;
            // all DA_NBG block groups of the tile advance together: inside one k-step the DA_NBG DMMAs are independent
            // (one accumulator fragment each), so the pipe never waits for an accumulator.  The k loop stays rolled:
            // unrolled, ptxas chains the eight steps of one accumulator back to back.
            double c[DA_NBG][2];
#pragma unroll
            for (int bg = 0; bg < DA_NBG; ++bg) { c[bg][0] = 0.0; c[bg][1] = 0.0; }
            const double* bq = Qt + (size_t)(cg * 8 + g4) * DA_LD + l4;
            const double* ap = As + (size_t)(sig * DA_NB + g4) * DA_LD + l4;
#pragma unroll 1
            for (int ks = 0; ks < 8; ++ks) {
                const double b = bq[ks * 4];
                double a[DA_NBG];
#pragma unroll
                for (int bg = 0; bg < DA_NBG; ++bg) a[bg] = ap[(size_t)(bg * 8) * DA_LD + ks * 4];
#pragma unroll
                for (int bg = 0; bg < DA_NBG; ++bg) da_dmma(c[bg][0], c[bg][1], a[bg], b);
            }
            // block phasors: Theta(row g4 of group bg) = tile seed + w 32 g4 dt + bg (w 256 dt).  With z = c0 + j c1 the
            // unit's sum is  e^{j seed} sum_rows e^{j w 32 g4 dt} sum_bg e^{j bg w 256 dt} z_bg : Horner over the block
            // groups with ONE phasor per bin (4 FMAs a step), the row phasor once, the tile seed after the row sum.
            const double2 r8 = R8[k];
            double hr = c[DA_NBG - 1][0], hi = c[DA_NBG - 1][1];
#pragma unroll
            for (int bg = DA_NBG - 2; bg >= 0; --bg) {
                const double nr = fma(hr, r8.x, fma(-hi, r8.y, c[bg][0]));
                hi = fma(hr, r8.y, fma(hi, r8.x, c[bg][1]));
                hr = nr;
            }
            const double2 p1 = P1[g4 * bp + k];
            double sre = fma(hr, p1.x, -(hi * p1.y)), sim = fma(hr, p1.y, hi * p1.x);
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                sre += __shfl_xor_sync(0xffffffffu, sre, o);
                sim += __shfl_xor_sync(0xffffffffu, sim, o);
            }
            if (g4 == 0) {
                const double2 pt = Pt0[k];
                const double fr = fma(sre, pt.x, -(sim * pt.y)), fi = fma(sre, pt.y, sim * pt.x);
                sre = fr; sim = fi;
            }
            if (g4 == 0) {
                accM[(2 * sig) * bp + k] += sre;
                accM[(2 * sig + 1) * bp + k] += sim;
            }
        }
    }
    __syncthreads();
    if (tid < nb_all) {
        double* out = cta_partial + (size_t)blockIdx.x * 4 * nd + bin0 + tid;
#pragma unroll
        for (int c = 0; c < 4; ++c) out[(size_t)c * nd] = accM[c * bp + tid];
    }
    if (do_norms) {
        // fixed-order CTA sums of the six moments: lanes by shuffle, warps in order
        double* red = As;
#pragma unroll
        for (int c = 0; c < DA_NORMS; ++c) {
            double v = nrm[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[c * DA_WARPS + warp] = v;
        }
        __syncthreads();
        if (tid < DA_NORMS) {
            double v = 0.0;
            for (int q = 0; q < DA_WARPS; ++q) v += red[tid * DA_WARPS + q];
            norm_partial[(size_t)blockIdx.x * DA_NORMS + tid] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 1'. main term, warp-specialised (the default): ONE CTA of 1024 threads per SM, the tables once per SM.
//   * producer warps 0-11: the raw t / u / y segments of a tile arrive by TMA bulk copies (cp.async.bulk, completion on
//     an mbarrier) two tiles ahead; the producers turn them into the weighted block rows As[tile & 1], the tile's seed
//     phasors and the moments of delta, then hand the buffer to the consumers and re-arm the copies of tile + 2;
//   * consumer warps 12-31: the block-phasor matrix products of frf_main_dmma_kernel on the buffer the producers
//     filled while the previous tile was multiplied.
// Named barriers full[b] / empty[b] (b = tile parity) carry the hand-over; there is no CTA-wide barrier in the tile
// loop, so the DMMA pipe never waits for a global load or for the staging arithmetic.
constexpr int DW_THREADS = 1024;
// producer threads: measured 128 / 256 / 384 / 512 -> 0.689 / 0.677 / 0.666 / 0.682 ms (N_d = 100) and 1.039 / 0.996 /
// 0.982 / 0.991 ms (150 bins) per adaptive call: a producer's FP64 instructions queue behind the consumers' DMMAs, so
// the per-tile staging chain is short only when it is spread over many warps
#ifndef DW_PRODUCER_THREADS
#define DW_PRODUCER_THREADS 384
#endif
constexpr int DW_PRODUCERS = DW_PRODUCER_THREADS;       // threads of the producer warps (warps 0 .. DW_PRODUCERS/32 - 1)
constexpr int DW_CWARPS = (DW_THREADS - DW_PRODUCERS) / 32;
constexpr int DW_RAW_T = DA_TS + 4;                     // t[i0 - 2 .. i0 + TS + 2): 16-byte aligned start, halo for the weights
constexpr int DW_RAW_DOUBLES = DW_RAW_T + 2 * DA_TS;    // t, u, y of one tile
constexpr size_t DA_SMEM_LIMIT_1CTA = 227 * 1024;
static inline size_t dw_smem_bytes(int bins_per_group, int raw_bufs = 2) {
    const size_t bp = (size_t)da_pad4(bins_per_group);
    return (size_t)2 * bp * DA_LD * sizeof(double)              // Qt
           + (size_t)2 * 2 * DA_NB * DA_LD * sizeof(double)     // A rows, two tiles
           + (size_t)raw_bufs * DW_RAW_DOUBLES * sizeof(double) // raw t, u, y, one or two tiles
           + (size_t)8 * bp * sizeof(double2)                   // P1
           + (size_t)bp * sizeof(double2)                       // R8
           + (size_t)3 * bp * sizeof(double2)                   // Pt0 (two tiles), RT
           + (size_t)4 * bp * sizeof(double)                    // main-term sums
           + bp * sizeof(double)                                // w
           + 64;                                                // mbarriers
}
static inline int dw_groups(int nd, int raw_bufs = 2) {
    int g = 1;
    while (dw_smem_bytes((nd + g - 1) / g, raw_bufs) > DA_SMEM_LIMIT_1CTA) ++g;
    return g;
}
// one raw buffer instead of two when that saves a bin group: every group stages the whole record again, and the
// producers -- not the copies -- are what the consumers wait for (150 bins: one group of 150 instead of two of 75)
static inline int dw_raw_bufs(int nd) { return dw_groups(nd, 1) < dw_groups(nd, 2) ? 1 : 2; }
__device__ __forceinline__ unsigned dw_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dw_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void dw_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void dw_mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}

__global__ void __launch_bounds__(DW_THREADS, 1)
frf_main_ws_kernel(const double* __restrict__ t, const double* __restrict__ u, const double* __restrict__ y,
                   int64_t n, int64_t own_begin, int64_t own_end, const double* __restrict__ w, int nd,
                   const double* __restrict__ sums, double inv_count, double t0, double dt,
                   const double* __restrict__ fit, int n_tiles, int bins_per_group, int raw_bufs,
                   double* __restrict__ cta_partial, double* __restrict__ norm_partial) {
    extern __shared__ __align__(16) unsigned char da_smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = (int)__reduce_max_sync(0xffffffffu, (unsigned)(tid >> 5));      // uniform register
    const int bin0 = blockIdx.y * bins_per_group;
    const int nb_all = min(nd - bin0, bins_per_group);
    const int bp = da_pad4(nb_all);
    const int ncg = bp >> 2;
    if (nb_all <= 0) return;

    double* raw = reinterpret_cast<double*>(da_smem);                // [raw_bufs][DW_RAW_DOUBLES]  t (halo 2 + 2), u, y
    double* Qt = raw + (size_t)raw_bufs * DW_RAW_DOUBLES;            // [2 bp][DA_LD]
    double* As = Qt + (size_t)2 * bp * DA_LD;                        // [2][2 DA_NB][DA_LD]
    double2* P1 = reinterpret_cast<double2*>(As + (size_t)2 * 2 * DA_NB * DA_LD);   // [8][bp]
    double2* R8 = P1 + (size_t)8 * bp;                               // [bp]
    double2* Pt0 = R8 + bp;                                          // [2][bp]  tile seed phasors of the tile in As[b]
    double2* RT = Pt0 + 2 * bp;                                      // [bp]     tile-to-tile rotation
    double* accM = reinterpret_cast<double*>(RT + bp);               // [4][bp]
    double* ws = accM + (size_t)4 * bp;                              // [bp]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(ws + bp);      // [2] raw[b] landed

    const double mean_u = sums ? sums[0] * inv_count : 0.0;
    const double mean_y = sums ? sums[1] * inv_count : 0.0;
    const bool do_norms = blockIdx.y == 0 && norm_partial != nullptr;
    const DaBase tb = da_base(fit, t0, dt, own_begin, own_end);

    for (int k = tid; k < bp; k += DW_THREADS) ws[k] = k < nb_all ? w[bin0 + k] : 0.0;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dw_smem_addr(bars + b)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int e = tid; e < bp * 32; e += DW_THREADS) {
        const int k = e >> 5, rr = e & 31;
        double s, c;
        da_sincos_steps(ws[k], (double)rr, tb, s, c);
        Qt[(size_t)(2 * k) * DA_LD + rr] = c;
        Qt[(size_t)(2 * k + 1) * DA_LD + rr] = s;
    }
    for (int e = tid; e < bp * 9; e += DW_THREADS) {
        const int g = e / bp, k = e - g * bp;
        double s, c;
        da_sincos_steps(ws[k], (double)(32 * g), tb, s, c);
        if (g < 8) P1[g * bp + k] = make_double2(c, s);
        else R8[k] = make_double2(c, s);
    }
    for (int k = tid; k < bp; k += DW_THREADS) {
        double s, c;
        da_sincos_steps(ws[k], (double)gridDim.x * (double)DA_TS, tb, s, c);
        RT[k] = make_double2(c, s);
        accM[k] = 0.0; accM[bp + k] = 0.0; accM[2 * bp + k] = 0.0; accM[3 * bp + k] = 0.0;
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    // a tile whose raw segments lie inside the record goes through the TMA; the first and the last tile of the record
    // (halo or length not a whole number of 16-byte pieces) are fetched by the producers themselves
    auto tile_i0 = [&](int tl) { return own_begin + (int64_t)tl * DA_TS; };
    auto uses_tma = [&](int tl) { const int64_t i0 = tile_i0(tl); return i0 >= 2 && i0 + DA_TS + 2 <= n && i0 + DA_TS <= own_end; };

    if (warp < DW_PRODUCERS / 32) {
        // =========================================== producers
        auto issue = [&](int tl, int b) {                      // one thread: arm the barrier, three bulk copies
            const int64_t i0 = tile_i0(tl);
            double* dst = raw + (size_t)b * DW_RAW_DOUBLES;
            const unsigned bar = dw_smem_addr(bars + b);
            const unsigned bytes_t = DW_RAW_T * sizeof(double), bytes_s = DA_TS * sizeof(double);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_t + 2 * bytes_s) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dw_smem_addr(dst)), "l"(t + i0 - 2), "r"(bytes_t), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dw_smem_addr(dst + DW_RAW_T)), "l"(u + i0), "r"(bytes_s), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dw_smem_addr(dst + DW_RAW_T + DA_TS)), "l"(y + i0), "r"(bytes_s), "r"(bar) : "memory");
        };
        if (tid == 0) {
            const int tl0 = blockIdx.x, tl1 = blockIdx.x + gridDim.x;
            if (tl0 < n_tiles && uses_tma(tl0)) issue(tl0, 0);
            if (raw_bufs == 2 && tl1 < n_tiles && uses_tma(tl1)) issue(tl1, 1);
        }
        // moments for the decision, in FP32 (they only need a digit or two, and every FP64 instruction of a producer
        // queues behind the consumers' DMMAs on the same pipe); delta is scaled by 2^40 so that its squares stay normal
        float nrm[DA_NORMS];
#pragma unroll
        for (int c = 0; c < DA_NORMS; ++c) nrm[c] = 0.f;
        unsigned phase0 = 0u, phase1 = 0u;                          // parity of the next landing in raw[0], raw[1]
        int it = 0;
        for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++it) {
            const int b = it & 1;
            const int rb = raw_bufs == 2 ? b : 0;                    // raw buffer of this tile
            const int64_t i0 = tile_i0(tl);
            const int cnt = (int)min((int64_t)DA_TS, own_end - i0);
            double* rw = raw + (size_t)rb * DW_RAW_DOUBLES;
            if (uses_tma(tl)) {
                dw_mbar_wait(dw_smem_addr(bars + rb), rb ? phase1 : phase0);
                if (rb) phase1 ^= 1u; else phase0 ^= 1u;
            } else {
                // edge tile: the producers fetch it (zeros past the owned range, clamped halo)
                for (int j = tid; j < DW_RAW_T; j += DW_PRODUCERS) {
                    const int64_t i = i0 - 2 + j;
                    rw[j] = __ldg(t + (i < 0 ? 0 : (i > n - 1 ? n - 1 : i)));
                }
                for (int j = tid; j < DA_TS; j += DW_PRODUCERS) {
                    const bool ok = j < cnt;
                    rw[DW_RAW_T + j] = ok ? __ldg(u + i0 + j) : 0.0;
                    rw[DW_RAW_T + DA_TS + j] = ok ? __ldg(y + i0 + j) : 0.0;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // these plain stores precede later bulk copies into raw[b]
                dw_bar_sync(5, DW_PRODUCERS);                        // producers only: the tile is complete
            }
            if (it >= 2) dw_bar_sync(3 + b, DW_THREADS);             // consumers are done with As[b] / Pt0[b] (tile it - 2)
            // ---- raw -> weighted block rows, moments of delta
            const double* rt = rw + 2;                               // rt[j] = t[i0 + j], j = -2 .. TS + 1
            const double ti0 = rt[0];
            const DD x0 = da_time(tb, (double)i0);
            const double d0 = (ti0 - x0.hi) - x0.lo;
            double* Ab = As + (size_t)b * 2 * DA_NB * DA_LD;
#pragma unroll 4
            for (int j = tid; j < DA_TS; j += DW_PRODUCERS) {
                double au = 0.0, ay = 0.0;
                if (j < cnt) {
                    const int64_t i = i0 + j;
                    const double ti = rt[j];
                    const double tm = i > 0 ? rt[j - 1] : ti, tp = i < n - 1 ? rt[j + 1] : ti;
                    const double wt = 0.5 * (tp - tm);
                    au = wt * (rw[DW_RAW_T + j] - mean_u);
                    ay = wt * (rw[DW_RAW_T + DA_TS + j] - mean_y);
                    if (do_norms) {
                        const double jd = (double)j;
                        const double dl = fma(-jd, tb.dtl, fma(-jd, tb.dth, ti - ti0)) + d0;
                        const double ad = fabs(dl);
                        const bool noise = ad <= da_ulp(ti);
                        const float adf = (float)(ad * 0x1p40), fu = fabsf((float)au), fy = fabsf((float)ay);
                        const float ds = noise ? adf : 0.f, dr = noise ? 0.f : adf;
                        const float u2 = fu * fu, y2 = fy * fy, ds2 = ds * ds;
                        nrm[0] += u2; nrm[1] += y2;
                        nrm[2] = fmaf(u2, ds2, nrm[2]); nrm[3] = fmaf(y2, ds2, nrm[3]);
                        nrm[4] = fmaf(fu, dr, nrm[4]); nrm[5] = fmaf(fy, dr, nrm[5]);
                    }
                }
                const int blk = j >> 5, rr = j & 31;
                Ab[(size_t)blk * DA_LD + rr] = au;
                Ab[(size_t)(DA_NB + blk) * DA_LD + rr] = ay;
            }
            // ---- phasor of the tile start: exact seed every DA_RESEED tiles, the exact tile-to-tile rotation in between
            if (tid < bp) {
                if (it % DA_RESEED == 0) {
                    double s, c;
                    sincos_dd(ws[tid], x0, s, c);
                    Pt0[b * bp + tid] = make_double2(c, s);
                } else {
                    const double2 p = Pt0[(b ^ 1) * bp + tid], q = RT[tid];
                    Pt0[b * bp + tid] = make_double2(fma(p.x, q.x, -(p.y * q.y)), fma(p.y, q.x, p.x * q.y));
                }
            }
            dw_bar_sync(5, DW_PRODUCERS);                            // every producer is done with raw[b] (and wrote As[b])
            dw_bar_arrive(1 + b, DW_THREADS);                        // full[b]
            if (tid == 0) {
                const int tl2 = tl + raw_bufs * (int)gridDim.x;      // the tile that will use this raw buffer next
                if (tl2 < n_tiles && uses_tma(tl2)) issue(tl2, rb);
            }
        }
        if (do_norms) {
            double* red = raw;                                       // producers own raw[] once their loop is over
            dw_bar_sync(5, DW_PRODUCERS);
#pragma unroll
            for (int c = 0; c < DA_NORMS; ++c) {
                double v = (double)nrm[c] * (c < 2 ? 1.0 : (c < 4 ? 0x1p-80 : 0x1p-40));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[c * (DW_PRODUCERS / 32) + warp] = v;
            }
            dw_bar_sync(5, DW_PRODUCERS);
            if (tid < DA_NORMS) {
                double v = 0.0;
                for (int q = 0; q < DW_PRODUCERS / 32; ++q) v += red[tid * (DW_PRODUCERS / 32) + q];
                norm_partial[(size_t)blockIdx.x * DA_NORMS + tid] = v;
            }
        }
    } else {
        // =========================================== consumers
        const int cw = warp - DW_PRODUCERS / 32;
        const int g4 = lane >> 2, l4 = lane & 3;
        int it = 0;
        for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++it) {
            const int b = it & 1;
            dw_bar_sync(1 + b, DW_THREADS);                          // full[b]
            const double* Ab = As + (size_t)b * 2 * DA_NB * DA_LD;
            const double2* Ptb = Pt0 + b * bp;
            for (int unit = cw; unit < 2 * ncg; unit += DW_CWARPS) {
                const int sig = unit & 1, cg = unit >> 1;
                const int k = cg * 4 + l4;
                double c[DA_NBG][2];
#pragma unroll
                for (int bg = 0; bg < DA_NBG; ++bg) { c[bg][0] = 0.0; c[bg][1] = 0.0; }
                const double* bq = Qt + (size_t)(cg * 8 + g4) * DA_LD + l4;
                const double* ap = Ab + (size_t)(sig * DA_NB + g4) * DA_LD + l4;
#pragma unroll 1
                for (int ks = 0; ks < 8; ++ks) {
                    const double bv = bq[ks * 4];
                    double a[DA_NBG];
#pragma unroll
                    for (int bg = 0; bg < DA_NBG; ++bg) a[bg] = ap[(size_t)(bg * 8) * DA_LD + ks * 4];
#pragma unroll
                    for (int bg = 0; bg < DA_NBG; ++bg) da_dmma(c[bg][0], c[bg][1], a[bg], bv);
                }
                const double2 r8 = R8[k];
                double hr = c[DA_NBG - 1][0], hi = c[DA_NBG - 1][1];
#pragma unroll
                for (int bg = DA_NBG - 2; bg >= 0; --bg) {
                    const double nr = fma(hr, r8.x, fma(-hi, r8.y, c[bg][0]));
                    hi = fma(hr, r8.y, fma(hi, r8.x, c[bg][1]));
                    hr = nr;
                }
                const double2 p1 = P1[g4 * bp + k];
                double sre = fma(hr, p1.x, -(hi * p1.y)), sim = fma(hr, p1.y, hi * p1.x);
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    sre += __shfl_xor_sync(0xffffffffu, sre, o);
                    sim += __shfl_xor_sync(0xffffffffu, sim, o);
                }
                if (g4 == 0) {
                    const double2 pt = Ptb[k];
                    accM[(2 * sig) * bp + k] += fma(sre, pt.x, -(sim * pt.y));
                    accM[(2 * sig + 1) * bp + k] += fma(sre, pt.y, sim * pt.x);
                }
            }
            dw_bar_arrive(3 + b, DW_THREADS);                        // empty[b]
        }
    }
    // the producers' last two empty[] waits never happen (they only wait from their third tile on): every barrier
    // phase that was armed by an arrive has been consumed or is abandoned at exit, which PTX allows.
    __syncthreads();
    if (tid < nb_all) {
        double* out = cta_partial + (size_t)blockIdx.x * 4 * nd + bin0 + tid;
#pragma unroll
        for (int c = 0; c < 4; ++c) out[(size_t)c * nd] = accM[c * bp + tid];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 3. which bins need the rounding term (one CTA).  main = folded main-term sums [4][nd] = {uc, us, yc, ys}.
//    diag (optional) receives {count, max over skipped bins of (DA_SIGMAS sigma + rigorous) / |S|}.
constexpr int DA_DECIDE_THREADS = 256;
__global__ void __launch_bounds__(DA_DECIDE_THREADS)
frf_decide_kernel(const double* __restrict__ main, const double* __restrict__ norm_partial, int norm_rows,
                  const double* __restrict__ w, int nd, double tmax_abs, int force_all, int* __restrict__ flag_count,
                  int* __restrict__ flag_idx, double* __restrict__ diag) {
    __shared__ double nrm[DA_NORMS];
    __shared__ double worst[DA_DECIDE_THREADS];
    extern __shared__ unsigned char flags[];                           // [nd]
    const int tid = threadIdx.x;
    if (tid < DA_NORMS) {
        double v = 0.0;
        for (int r = 0; r < norm_rows; ++r) v += norm_partial[(size_t)r * DA_NORMS + tid];
        nrm[tid] = v;
    }
    __syncthreads();
    double wr = 0.0;
    for (int k = tid; k < nd; k += DA_DECIDE_THREADS) {
        const double wk = fabs(w[k]);
        const double e = da_ulp(wk * tmax_abs);
        const double var_e = e * e * (1.0 / 12.0);
        bool need = force_all != 0;
        double ratio = 0.0;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const double sigma = sqrt(var_e * nrm[s] + wk * wk * nrm[2 + s]);
            const double bound = DA_SIGMAS * sigma + wk * nrm[4 + s];
            const double mag = hypot(main[(size_t)(2 * s) * nd + k], main[(size_t)(2 * s + 1) * nd + k]);
            if (!(bound <= DA_BUDGET * mag)) need = true;              // also true for NaN / Inf anywhere
            ratio = fmax(ratio, bound / mag);
        }
        if (force_all == 2) need = false;                               // timing experiments only: main term alone
        flags[k] = need ? 1 : 0;
        if (!need) wr = fmax(wr, ratio);
    }
    worst[tid] = wr;
    __syncthreads();
    if (tid == 0) {
        int cnt = 0;
        for (int k = 0; k < nd; ++k)
            if (flags[k]) flag_idx[cnt++] = k;
        flag_count[0] = cnt;
        if (diag) {
            double v = 0.0;
            for (int q = 0; q < DA_DECIDE_THREADS; ++q) v = fmax(v, worst[q]);
            diag[0] = (double)cnt;
            diag[1] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 4. rounding term of the listed bins.  1-D grid; CTA -> (bin group, tile stripe) from the device-side count.
//    thread = (bin pair b2, sample stripe r): sample j of a tile belongs to chain j % R, step j / R; the tile is staged
//    chain-major (a thread's samples are consecutive, so every load has an immediate offset), the 8 residuals of a
//    group of 4 steps x 2 bins are 8 independent DMUL-DFMA-F2F-FFMA chains, and only the packed-FP32 Goertzel updates
//    (2 FFMA2 per step and bin, u and y together) are sequential.  Bins whose chain step is ill-conditioned for
//    Goertzel (|sin(w R dt)| < FRFU_SIN_MIN) take a plain FP32 rotation.
__global__ void __launch_bounds__(DA_THREADS, 2)
frf_round_goertzel_kernel(const double* __restrict__ t, const double* __restrict__ u, const double* __restrict__ y,
                          int64_t n, int64_t own_begin, int64_t own_end, const double* __restrict__ w, int nd,
                          const double* __restrict__ sums, double inv_count, double t0, double dt,
                          const double* __restrict__ fit, double tmax_abs, int n_tiles,
                          const int* __restrict__ flag_count, const int* __restrict__ flag_idx,
                          double* __restrict__ cta_partial) {
    const int nf = flag_count[0];
    if (nf <= 0) return;
    const DaRoundGeom geo = da_round_geom(nf, gridDim.x);
    const int grp = blockIdx.x % geo.groups, bx = blockIdx.x / geo.groups;
    if (bx >= geo.gx) return;
    extern __shared__ __align__(16) unsigned char da_smem[];
    const int tid = threadIdx.x;
    const int kc0 = grp * geo.bpg;                                   // first compact column of this CTA
    const int nb_all = min(nf - kc0, geo.bpg);
    if (nb_all <= 0) return;
    const DaBase tb = da_base(fit, t0, dt, own_begin, own_end);

    double* tt = reinterpret_cast<double*>(da_smem);                 // [DA_CM]   t_i, chain-major
    float2* fa = reinterpret_cast<float2*>(tt + DA_CM);              // [DA_CM]   (a_u, a_y) in FP32, chain-major
    double* ws = reinterpret_cast<double*>(fa + DA_CM);              // [DA_BPB]  w_k
    float2* SD = reinterpret_cast<float2*>(ws + DA_BPB);             // [DA_BPB]  (cos, sin)(w_k R dt)
    float* fd = reinterpret_cast<float*>(SD + DA_BPB);               // [DA_CM]   delta_i in FP32, chain-major

    const double mean_u = sums ? sums[0] * inv_count : 0.0;
    const double mean_y = sums ? sums[1] * inv_count : 0.0;

    const int nb2 = (nb_all + 1) >> 1;
    for (int k = tid; k < nb_all; k += DA_THREADS) ws[k] = w[flag_idx[kc0 + k]];
    __syncthreads();
    // sample stripes R: as many as the threads allow, minus a few if that keeps every bin's chain step w R dt away from
    // a multiple of pi -- ONE ill-conditioned bin sends a thread of every warp down the rotation path below, and the
    // divergent warps then pay for both paths (measured: one such bin in 50 tripled the kernel's time)
    int R = DA_THREADS / nb2;
    for (int tries = 0; tries < 6; ++tries) {
        int any_bad = 0;
        for (int k = tid; k < nb_all; k += DA_THREADS) {
            double s, c;
            da_sincos_steps(ws[k], (double)R, tb, s, c);
            SD[k] = make_float2((float)c, (float)s);
            any_bad |= fabsf((float)s) < (float)FRFU_SIN_MIN;
        }
        if (!__syncthreads_or(any_bad) || R <= 8 || tries == 5) break;
        --R;
    }
    __syncthreads();
    // steps of the longest chain, made ODD: consecutive samples of a tile go to consecutive chains, i.e. Mmax entries
    // apart, and an even pitch puts the 16 lanes of a 64-bit store on a handful of banks (measured: 33.5 M bank
    // conflicts per launch with Mmax = 48)
    const int Mmax = ((DA_TS + R - 1) / R) | 1;
    const int b2 = tid % nb2, r = tid / nb2;
    const bool active = r < R;
    const int kA = b2, kB = (b2 + nb2 < nb_all) ? b2 + nb2 : b2;     // odd count: the last pair's second bin is a dummy

    // The residual e = w t - fl(w t) reaches FP32 WITHOUT a conversion instruction (F2F.F32.F64 runs at a quarter of the
    // FP64 rate and was the busiest pipe of this loop): with S_k = 2^21 / ulp(w_k t_max) a power of two, w' = w S gives
    // p' = fl(w' t) = S fl(w t) and e' = fma(w', t, -p') = S e exactly, |e'| <= 2^20; adding 2^52 + 0x4B400000 rounds e'
    // to an integer whose low word IS the FP32 number 2^23 + 2^22 + e'.  The recurrence then runs on S eps and the
    // chain ends are scaled back by 1 / S.
    const double SA = fmin(2097152.0 / da_ulp(ws[kA] * tmax_abs), 0x1p100), SB = fmin(2097152.0 / da_ulp(ws[kB] * tmax_abs), 0x1p100);
    const double wSA = ws[kA] * SA, wSB = ws[kB] * SB;
    const float wSAf = (float)wSA, wSBf = (float)wSB;
    const float invSA = (float)(1.0 / SA), invSB = (float)(1.0 / SB);
    const float KA = 2.0f * SD[kA].x, KB = 2.0f * SD[kB].x;
    const bool bad = fabsf(SD[kA].y) < (float)FRFU_SIN_MIN || fabsf(SD[kB].y) < (float)FRFU_SIN_MIN;
    // float accumulators of the rounding term: (d S_c, d S_s) of u and y for bins A and B
    float2 gAc = make_float2(0.f, 0.f), gAs = gAc, gBc = gAc, gBs = gAc;      // .x: u, .y: y

    // chain-major slot of this thread's samples j = tid + q DA_THREADS: chain j % R, step j / R (the same every tile)
    int pos_q[DA_TS / DA_THREADS];
#pragma unroll
    for (int q = 0; q < DA_TS / DA_THREADS; ++q) {
        const int j = tid + q * DA_THREADS, m = j / R;
        pos_q[q] = (j - m * R) * Mmax + m;
    }
    double2* seeds = reinterpret_cast<double2*>(fd + DA_CM);             // [2] (t, delta) of the first sample of tile it, it + 1
    if (tid == DA_THREADS - 1 && bx < n_tiles) seeds[0] = da_tile_seed(t, own_begin + (int64_t)bx * DA_TS, tb);

    int it = 0;
    for (int tl = bx; tl < n_tiles; tl += geo.gx, ++it) {
        const int64_t i0 = own_begin + (int64_t)tl * DA_TS;
        const int cnt = (int)min((int64_t)DA_TS, own_end - i0);
        __syncthreads();                                                 // previous tile consumed
        const double2 seed = seeds[it & 1];
#pragma unroll
        for (int q = 0; q < DA_TS / DA_THREADS; ++q) {
            const int j = tid + q * DA_THREADS;
            DaSample s; s.ti = 0.0; s.au = 0.0; s.ay = 0.0; s.dl = 0.0;
            if (j < cnt) s = da_sample(t, u, y, i0 + j, n, (double)j, seed, tb, mean_u, mean_y);
            tt[pos_q[q]] = s.ti;
            fa[pos_q[q]] = make_float2((float)s.au, (float)s.ay);
            fd[pos_q[q]] = (float)s.dl;
        }
        if (tid == DA_THREADS - 1 && tl + geo.gx < n_tiles)
            seeds[(it + 1) & 1] = da_tile_seed(t, i0 + (int64_t)geo.gx * DA_TS, tb);       // read after the next loop-top barrier
#if DA_ROUND_PREFETCH
        // L2 prefetch of this CTA's next tile (one 128-byte line per thread: t, u, y), so that its staging loads -- the
        // only global reads of the kernel, issued right after the chains of this tile -- find their lines in the L2
        if (tl + geo.gx < n_tiles && tid < 3 * (DA_TS / 16)) {
            const int arr = tid / (DA_TS / 16), ln = tid - arr * (DA_TS / 16);
            const int64_t ip = i0 + (int64_t)geo.gx * DA_TS + 16 * ln;
            if (ip < n) {
                const double* src = arr == 0 ? t : (arr == 1 ? u : y);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ip));
            }
        }
#endif
        __syncthreads();
        if (!active || r >= cnt) continue;
        const int M = (cnt - r + R - 1) / R;                             // steps of this chain in this tile
        const double* pt_ = tt + r * Mmax;
        const float2* pa_ = fa + r * Mmax;
        const float* pd_ = fd + r * Mmax;
        if (!bad) {
            float2 qA1 = make_float2(0.f, 0.f), qA2 = qA1, qB1 = qA1, qB2 = qA1;
            // eps of one step for both bins: e = w t - fl(w t) exactly, narrowed, then w delta - e
#define DA_EPS(tv, dv, eA, eB)                                                                       \
    {                                                                                                \
        const double pA_ = wSA * (tv), pB_ = wSB * (tv);                                             \
        const double rA_ = fma(wSA, (tv), -pA_) + DA_MAGIC, rB_ = fma(wSB, (tv), -pB_) + DA_MAGIC;   \
        eA = fmaf(wSAf, (dv), DA_MAGIC_F) - __int_as_float(__double2loint(rA_));                     \
        eB = fmaf(wSBf, (dv), DA_MAGIC_F) - __int_as_float(__double2loint(rB_));                     \
    }
            // one Goertzel step of one bin: Q1 holds s[m-1], Q2 holds s[m-2] and receives s[m] = a eps + K s[m-1] - s[m-2]
#define DA_STEP(av, e, K, Q1, Q2)                                                                    \
    Q2 = __ffma2_rn(make_float2((K), (K)), Q1, __ffma2_rn((av), make_float2((e), (e)), make_float2(-Q2.x, -Q2.y)));
            int m = 0;
#pragma unroll 1
            for (; m + 4 <= M; m += 4) {
                const double t0_ = pt_[m], t1_ = pt_[m + 1], t2_ = pt_[m + 2], t3_ = pt_[m + 3];
                const float d0_ = pd_[m], d1_ = pd_[m + 1], d2_ = pd_[m + 2], d3_ = pd_[m + 3];
                const float2 a0_ = pa_[m], a1_ = pa_[m + 1], a2_ = pa_[m + 2], a3_ = pa_[m + 3];
                float eA0, eB0, eA1, eB1, eA2, eB2, eA3, eB3;
                DA_EPS(t0_, d0_, eA0, eB0)
                DA_EPS(t1_, d1_, eA1, eB1)
                DA_EPS(t2_, d2_, eA2, eB2)
                DA_EPS(t3_, d3_, eA3, eB3)
                DA_STEP(a0_, eA0, KA, qA1, qA2) DA_STEP(a0_, eB0, KB, qB1, qB2)
                DA_STEP(a1_, eA1, KA, qA2, qA1) DA_STEP(a1_, eB1, KB, qB2, qB1)
                DA_STEP(a2_, eA2, KA, qA1, qA2) DA_STEP(a2_, eB2, KB, qB1, qB2)
                DA_STEP(a3_, eA3, KA, qA2, qA1) DA_STEP(a3_, eB3, KB, qB2, qB1)
            }
            for (; m < M; ++m) {                                     // up to three single steps, roles swapped back each time
                const double tv = pt_[m]; const float dv = pd_[m]; const float2 av = pa_[m];
                float eA, eB;
                DA_EPS(tv, dv, eA, eB)
                DA_STEP(av, eA, KA, qA1, qA2) DA_STEP(av, eB, KB, qB1, qB2)
                float2 sq;
                sq = qA1; qA1 = qA2; qA2 = sq; sq = qB1; qB1 = qB2; qB2 = sq;
            }
#undef DA_EPS
#undef DA_STEP
            // (.1) = s[M-1], (.2) = s[M-2].  Phasor of the chain's last sample jl = r + R (M - 1): the phase in turns
            // w/(2 pi) (t0 + (i0 + jl) dt) reduced in FP64 (error ~1e-9 turns at 1e7 rad), then FP32 -- the rounding term
            // is <= 1e-6 of the sum and needs three digits.
            const DD xl = da_time(tb, (double)(i0 + r + (int64_t)R * (M - 1)));
            // E = (s1 - cd s2) + j (sd s2);  d S_c = cl Ei - sl Er,  d S_s = sl Ei + cl Er
#define DA_FOLD(k_, inv_, Q1, Q2, GC, GS)                                                            \
    {                                                                                                \
        const double tr_ = ws[k_] * 0.15915494309189535 * xl.hi;                                     \
        const float fr_ = (float)(tr_ - rint(tr_));                                                  \
        float sl_, cl_;                                                                              \
        sincospif(2.0f * fr_, &sl_, &cl_);                                                           \
        const float2 sd_ = SD[k_];                                                                   \
        const float eru = (inv_) * fmaf(-sd_.x, Q2.x, Q1.x), eiu = (inv_) * (sd_.y * Q2.x);          \
        const float ery = (inv_) * fmaf(-sd_.x, Q2.y, Q1.y), eiy = (inv_) * (sd_.y * Q2.y);          \
        GC.x += cl_ * eiu - sl_ * eru; GS.x += sl_ * eiu + cl_ * eru;                                \
        GC.y += cl_ * eiy - sl_ * ery; GS.y += sl_ * eiy + cl_ * ery;                                \
    }
            DA_FOLD(kA, invSA, qA1, qA2, gAc, gAs)
            DA_FOLD(kB, invSB, qB1, qB2, gBc, gBs)
#undef DA_FOLD
        } else {
            // ill-conditioned chain step for one of this thread's bins: plain FP32 rotation along the chain
            // (d S_c = -sum x sin(theta), d S_s = +sum x cos(theta), x = a eps); seeds from the exact phase
            const DD x0 = da_time(tb, (double)(i0 + r));
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
                const double wk = ws[which ? kB : kA];
                const float wkf = (float)wk;
                const float2 sd = SD[which ? kB : kA];
                double s0, c0;
                sincos_dd(wk, x0, s0, c0);
                float c = (float)c0, s = (float)s0;
                float2 ac = make_float2(0.f, 0.f), as = ac;
                for (int m = 0; m < M; ++m) {
                    const double tv = pt_[m]; const float2 av = pa_[m]; const float dv = pd_[m];
                    const double p = wk * tv;
                    const float ef = (float)fma(wk, tv, -p);
                    const float epsf = fmaf(wkf, dv, -ef);
                    const float xu = av.x * epsf, xy = av.y * epsf;
                    ac.x = fmaf(-xu, s, ac.x); as.x = fmaf(xu, c, as.x);
                    ac.y = fmaf(-xy, s, ac.y); as.y = fmaf(xy, c, as.y);
                    const float cn = c * sd.x - s * sd.y;
                    s = s * sd.x + c * sd.y; c = cn;
                }
                if (which) { gBc.x += ac.x; gBc.y += ac.y; gBs.x += as.x; gBs.y += as.y; }
                else       { gAc.x += ac.x; gAc.y += ac.y; gAs.x += as.x; gAs.y += as.y; }
            }
        }
    }
    __syncthreads();
    // ---- per-bin result = the stripes' rounding terms (fixed order)
    float* red = reinterpret_cast<float*>(tt);                          // [8][DA_THREADS], reuses the t rows
    red[0 * DA_THREADS + tid] = gAc.x; red[1 * DA_THREADS + tid] = gAs.x;
    red[2 * DA_THREADS + tid] = gAc.y; red[3 * DA_THREADS + tid] = gAs.y;
    red[4 * DA_THREADS + tid] = gBc.x; red[5 * DA_THREADS + tid] = gBs.x;
    red[6 * DA_THREADS + tid] = gBc.y; red[7 * DA_THREADS + tid] = gBs.y;
    __syncthreads();
    if (tid < nb_all) {
        const int slot = tid < nb2 ? 0 : 4, base = tid < nb2 ? tid : tid - nb2;
        double* out = cta_partial + (size_t)bx * 4 * nd + kc0 + tid;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double v = 0.0;
            for (int rr = 0; rr < R; ++rr) v += (double)red[(slot + c) * DA_THREADS + rr * nb2 + base];
            out[(size_t)c * nd] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 5. sums[c][flag_idx[kc]] += sum over the tile stripes of the rounding-term partials (fixed order; same shape as
//    frf_fold_kernel, geometry from the device-side count).  One CTA per 32 compact columns of one component.
__global__ void __launch_bounds__(FOLD_WARPS * 32)
frf_fold_add_kernel(const double* __restrict__ part, int n_cta, int nd, const int* __restrict__ flag_count,
                    const int* __restrict__ flag_idx, double* __restrict__ out) {
    __shared__ double sh[FOLD_WARPS][33];
    const int nf = flag_count[0];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.y, kc = blockIdx.x * 32 + lane;
    if (blockIdx.x * 32 >= nf) return;
    const DaRoundGeom geo = da_round_geom(nf, n_cta);
    double v = 0.0;
    if (kc < nf) {
#pragma unroll 4
        for (int x = warp; x < geo.gx; x += FOLD_WARPS) v += part[(size_t)x * 4 * nd + (size_t)c * nd + kc];
    }
    sh[warp][lane] = v;
    __syncthreads();
    if (warp == 0 && kc < nf) {
        double acc = sh[0][lane];
#pragma unroll
        for (int k = 1; k < FOLD_WARPS; ++k) acc += sh[k][lane];
        out[(size_t)c * nd + flag_idx[kc]] += acc;
    }
}

}  // namespace b200id
