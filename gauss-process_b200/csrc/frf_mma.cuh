// frf_mma.cuh -- K1 uniform-timebase projection as a block-phasor matrix product (included by frf_project.cu).
//
// Same sums as frf_project_uniform*_kernel (reference src/frequency_transform/frf_estimator.py:226-231 for u and
// y in one pass), reorganised so that the FP64 pipe only executes the 4 algorithmic FMAs per sample-bin:
//
//   fl(w_k t_i) = theta_ki + eps_ki,  theta_ki = w_k (t0 + i dt)  (exact arithmetic progression),
//   cos/sin(fl(w_k t_i)) = cos/sin(theta_ki) -/+ eps_ki sin/cos(theta_ki) + O(eps^2).
//
// MAIN TERM.  Samples are cut into blocks of 32: i = i0 + 32 b + r.  With Q[r][k] = exp(j w_k r dt),
//   sum_i a_i exp(j theta_ki) = sum_b exp(j w_k (t0 + (i0 + 32 b) dt)) * ( sum_r a_{b,r} Q[r][k] ),
// and the inner sum over r is a dense [blocks x 32] x [32 x 2 N_d] FP64 matrix product, evaluated with
// mma.sync.m8n8k4.f64 (one warp instruction = 256 FMAs, so the FP64 pipe is kept busy from one issue slot in
// sixteen and the other slots are free for the correction below).  Rows are the 32-sample blocks of u and of y,
// columns are (cos, sin) pairs per bin.  The block phasor is applied to the accumulator fragment in two levels
// (block inside a superblock of 8 blocks: table P1; superblock inside the tile: rotation from the tile's exact
// double-double seed), 8 FMAs per 256 sample-bins.
//
// CORRECTION.  eps_ki = fl(w_k t_i) - theta_ki = w_k delta_i - e_ki with e_ki = w_k t_i - fl(w_k t_i) the rounding
// residual of the reference's product.  Its contribution sum_i (a_i eps_ki) exp(j theta_ki) is at most ~1e-7 of a
// bin, so 1e-3 relative accuracy is enough and it runs on the integer / FP32 pipes while the FP64 pipe does the
// matrix product.  Inside one binade of |w_k t| the rounding unit u = ulp(w_k t) is constant and
//   e_ki / u = centred fractional part of  (theta_ki + w_k delta_i) / u ,
// which is evaluated in 32-bit fixed point: X = X0 + r d + m_i g (mod 2^32, read as a signed number = the centred
// fraction), where X0 and d come from the 64-bit fixed-point phase of the block start and of one sample step
// (exact integers prepared per bin by the setup kernel from the double-double products w_k t0 and w_k dt), m_i is
// delta_i in units of 2^-15 ulp(t), and g = w_k in the matching unit.  Lanes are blocks, bins are warp-uniform,
// Q as float2 comes from the constant bank, the two signals share packed f32x2 FMAs, and the block phasor is
// applied once per (block, bin) from the FP64 tables.  Wherever the integer evaluation is not safe -- the
// fraction within 2^-13 of a rounding tie, a binade boundary of w_k t inside the block, a phase below the
// fixed-point window, or |delta_i| above 8 ulp(t) -- eps is formed in FP64 exactly as the recurrence kernels do
// (p = w t; e = fma(w, t, -p); eps = fma(w, delta, -e)); that costs a few per mille of the sample-bins.
//
// Warp roles: warps 0..7 run the matrix product, warps 8..15 the correction; both read the same staged tile.
// Per-(slot, component, bin) FP64 accumulators live in shared memory, every entry is updated by exactly one warp
// per tile in a fixed order, so the result is deterministic for a given grid.
#pragma once

namespace b200id {

#ifndef FM_VARIANT
#define FM_VARIANT 0          // timing experiments only: bit 0 skips the matrix product, bit 1 the correction
#endif
constexpr int FM_THREADS = 512;
constexpr int FM_MMA_WARPS = 8;
constexpr int FM_COR_WARPS = 8;
constexpr int FM_BLK = 32;                 // samples per block = K extent of the product
constexpr int FM_LD = 36;                  // padded row length (doubles) of the A and Q rows: conflict-free fragments
constexpr int FM_SB = 8;                   // blocks per superblock = rows of one m-tile
constexpr int FM_NDMAX = 160;              // bins per launch (table sizes; shared memory decides what actually fits)
constexpr int FM_NGRP = 5;                 // n-tiles (of 4 bins) per matrix-product task
constexpr int FM_KC = 4;                   // bins per correction task (even: Q is fetched two bins at a time)
constexpr int FM_MBITS = 15;               // delta_i is quantised to 2^-15 ulp(t)
constexpr int FM_MLIM = 1 << 18;           // |m_i| limit (8 ulp(t)); larger jitter takes the FP64 path
// tie window of the centred fraction: 2^-13 on either side of +-1/2 (the fixed-point fraction is good to ~2^-14:
// 2^-15 from the quantised jitter, 2^-15 from the rounded g)
constexpr unsigned FM_TIE_OFF = 0x80000000u + (1u << 19), FM_TIE_WIN = 1u << 20;


struct FmTables {                                        // global-memory tables written by frf_mma_setup_kernel
    double* qt;                  // [2 ndp][32]  row 2k: cos(w_k r dt), row 2k+1: sin(w_k r dt)
    double2* p1;                 // [8][FM_NDMAX] (cos, sin)(w_k 32 b dt), b < 8
    double2* rot;                // [FM_NDMAX]    (cos, sin)(w_k 256 dt)
    double* wpad;                // [FM_NDMAX]
    float2* q32;                 // [32][FM_NDMAX]
    unsigned long long* a64;     // [FM_NDMAX] floor(w_k t0 2^S_k) mod 2^64
    unsigned long long* d64;     // [FM_NDMAX] floor(w_k dt 2^S_k) mod 2^64
    int* ek;                     // [FM_NDMAX] E_k = largest binade exponent of |w_k t| over the launch; S_k = 116 - E_k
};
static inline size_t fm_tables_bytes() {
    const size_t nd = FM_NDMAX;
    return align_up(2 * nd * FM_BLK * sizeof(double), 256) + align_up(nd * FM_SB * sizeof(double2), 256) +
           align_up(nd * sizeof(double2), 256) + align_up(nd * sizeof(double), 256) +
           align_up((size_t)FM_BLK * nd * sizeof(float2), 256) + 2 * align_up(nd * sizeof(unsigned long long), 256) +
           align_up(nd * sizeof(int), 256);
}
static inline FmTables fm_tables_at(void* base) {
    const size_t nd = FM_NDMAX;
    char* p = (char*)base;
    FmTables T;
    T.qt = (double*)p;    p += align_up(2 * nd * FM_BLK * sizeof(double), 256);
    T.p1 = (double2*)p;   p += align_up(nd * FM_SB * sizeof(double2), 256);
    T.rot = (double2*)p;  p += align_up(nd * sizeof(double2), 256);
    T.wpad = (double*)p;  p += align_up(nd * sizeof(double), 256);
    T.q32 = (float2*)p;   p += align_up((size_t)FM_BLK * nd * sizeof(float2), 256);
    T.a64 = (unsigned long long*)p; p += align_up(nd * sizeof(unsigned long long), 256);
    T.d64 = (unsigned long long*)p; p += align_up(nd * sizeof(unsigned long long), 256);
    T.ek = (int*)p;
    return T;
}

// row stride (in double2) of the P1 table in shared memory: = 4 mod 8 so that the fragment-shaped reads
// (4 consecutive bins x 2 rows per quarter warp) touch distinct banks
__host__ __device__ inline int fm_p1_stride(int ndp) { return ndp + ((4 - ndp % 8) + 8) % 8; }

static inline size_t fm_smem_bytes(int ndp, int nblk) {
    const size_t nsb = nblk / FM_SB, nacc = (size_t)(2 * nblk / 16) * 2 + (size_t)(nblk / 32) * 4;
    return (size_t)2 * ndp * FM_LD * sizeof(double)            // Qt
           + (size_t)2 * nblk * FM_LD * sizeof(double)         // A (u rows, y rows)
           + (size_t)FM_BLK * (nblk + 1) * sizeof(float4)      // (m, delta 2^M, a_u, a_y) by [r][block]
           + (size_t)FM_SB * fm_p1_stride(ndp) * sizeof(double2)   // P1
           + (size_t)ndp * nsb * sizeof(double2)               // tile phasors per superblock
           + nacc * ndp * sizeof(double)                       // accumulators
           + (size_t)ndp * 2 * sizeof(unsigned long long)      // fixed-point phase of the tile start, step
           + (size_t)ndp * sizeof(int)                         // E_k
           + (size_t)nblk * sizeof(int)                        // per-block jitter flags
           + (size_t)FM_BLK * ndp * sizeof(float2)             // Q as float2 [r][k]
           + (size_t)ndp * sizeof(double);                     // w_k
}

__device__ __forceinline__ int fm_expo(double x) {           // E with 2^E <= |x| < 2^(E+1) (x normal, non-zero)
    return ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
}
__device__ __forceinline__ double fm_pow2(int e) {           // 2^e, -1022 <= e <= 1023
    return __hiloint2double((e + 1023) << 20, 0);
}
// floor((hi + lo) 2^S) mod 2^64 for an unevaluated sum of two doubles
__device__ __forceinline__ unsigned long long fm_fixed_part(double x, int S) {
    if (x == 0.0) return 0ull;
    const long long bits = __double_as_longlong(x);
    const int ex = (int)((bits >> 52) & 0x7ff);
    long long m = (bits & 0x000fffffffffffffll) | (ex ? 0x0010000000000000ll : 0ll);
    if (bits < 0) m = -m;
    const int sh = (ex ? ex : 1) - 1075 + S;                  // x = m 2^(ex - 1075)
    if (sh >= 64) return 0ull;
    if (sh >= 0) return (unsigned long long)m << sh;
    if (sh <= -64) return m < 0 ? ~0ull : 0ull;
    return (unsigned long long)(m >> (-sh));                  // arithmetic shift = floor
}
__device__ __forceinline__ unsigned long long fm_fixed(double a, double b, int S) {   // of the exact product a b
    const double ph = a * b, pl = fma(a, b, -ph);
    return fm_fixed_part(ph, S) + fm_fixed_part(pl, S);
}
// exponent of the delta quantisation unit of the block that starts at sample index ib (global): 2^-M with
// M = 52 - E(t) + FM_MBITS, E(t) from the larger of the block's end-point times; the staging threads and the
// correction lanes both call this, so the two sides agree bit for bit
__device__ __forceinline__ int fm_block_m(int64_t ib, double t0, double dt) {
    const double ta = fabs(fma((double)ib, dt, t0)), tb = fabs(fma((double)(ib + FM_BLK - 1), dt, t0));
    const double tm = fmax(fmax(ta, tb), 1e-300);
    return min(max(52 - fm_expo(tm) + FM_MBITS, -1000), 1000);
}

// The stored time stamp t_i = t0 + i dt + delta_i, rebuilt without touching memory: the sum is exactly representable
// (it IS t_i), so evaluating it as TwoSum(t0, i dt) plus the small terms rounds to t_i even with delta known to 24 bits
// (|delta| <= 8 ulp(t): the error of the small terms is < 2^-20 ulp).
__device__ __forceinline__ double fm_time_of(int64_t i, double t0, double dt, double delta) {
    const double kd = (double)i;
    const double ph = kd * dt, pl = fma(kd, dt, -ph);
    const double sh = t0 + ph, bb = sh - t0;
    const double err = (t0 - (sh - bb)) + (ph - bb);
    return sh + ((err + pl) + delta);
}

// delta_i = t_i - (t0 + i dt), exact: TwoDiff(t_i, t0), then the exact product i dt is taken off
__device__ __forceinline__ double fm_delta_of(double ti, int64_t i, double t0, double dt) {
    const double kd = (double)i;
    const double ph = kd * dt, pl = fma(kd, dt, -ph);
    const double d = ti - t0, bb = d - ti;
    const double derr = (ti - (d - bb)) + (-t0 - bb);
    return ((d - ph) - pl) + derr;
}

// one thread per (bin, r): the small exact tables.  Padded bins (k >= nd) use w = 0.
__global__ void frf_mma_setup_kernel(const double* __restrict__ w, int nd, int ndp, double t0, double dt,
                                     double t_first, double t_last, FmTables T) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = e / FM_BLK, r = e % FM_BLK;
    if (k >= FM_NDMAX) return;
    const double wk = k < nd ? w[k] : 0.0;
    double s, c;
    {
        DD st; const double kk = (double)r;
        st.hi = kk * dt; st.lo = fma(kk, dt, -st.hi);
        sincos_dd(wk, st, s, c);
    }
    T.q32[r * FM_NDMAX + k] = make_float2((float)c, (float)s);
    if (r == 0) T.wpad[k] = wk;
    if (r == 9) {
        const double pm = fmax(fmax(fabs(wk * t_first), fabs(wk * t_last)), 1e-300);
        const int ek = fm_expo(pm);
        const int S = 116 - ek;
        T.ek[k] = ek;
        T.a64[k] = fm_fixed(wk, t0, S);
        T.d64[k] = fm_fixed(wk, dt, S);
    }
    if (r < FM_SB) {
        DD st; const double kk = (double)(r * FM_BLK);
        st.hi = kk * dt; st.lo = fma(kk, dt, -st.hi);
        double s1, c1;
        sincos_dd(wk, st, s1, c1);
        T.p1[r * FM_NDMAX + k] = make_double2(c1, s1);
    } else if (r == FM_SB) {
        DD st; const double kk = (double)(FM_SB * FM_BLK);
        st.hi = kk * dt; st.lo = fma(kk, dt, -st.hi);
        double s1, c1;
        sincos_dd(wk, st, s1, c1);
        T.rot[k] = make_double2(c1, s1);
    }
    if (k >= ndp) return;
    T.qt[(size_t)(2 * k) * FM_BLK + r] = c;
    T.qt[(size_t)(2 * k + 1) * FM_BLK + r] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NBLK>
__global__ void __launch_bounds__(FM_THREADS, 1)
frf_project_mma_kernel(const double* __restrict__ t, const double* __restrict__ u, const double* __restrict__ y,
                       int64_t n, int64_t own_begin, int64_t own_end, int nd, int ndp,
                       const double* __restrict__ sums, double inv_count, double t0, double dt, int n_tiles,
                       FmTables T, double* __restrict__ cta_partial) {
    constexpr int TS = NBLK * FM_BLK;                  // samples per tile
    constexpr int NSB = NBLK / FM_SB;                  // superblocks per tile
    constexpr int NMG = 2 * NBLK / 16;                 // m-groups of 16 rows (u rows then y rows)
    constexpr int NCG = NBLK / 32;                     // correction lane groups
    constexpr int NACC = NMG * 2 + NCG * 4;            // accumulator rows of ndp doubles
    constexpr int CPAD = NBLK + 1;
    const int S1 = fm_p1_stride(ndp);
    extern __shared__ __align__(16) unsigned char fm_smem[];
    double* Qt = reinterpret_cast<double*>(fm_smem);                   // [2 ndp][FM_LD]
    double* As = Qt + (size_t)2 * ndp * FM_LD;                         // [2 NBLK][FM_LD]
    float4* cin = reinterpret_cast<float4*>(As + (size_t)2 * NBLK * FM_LD);     // [32][CPAD] (m as int bits, delta 2^M, a_u, a_y)
    double2* P1 = reinterpret_cast<double2*>(cin + FM_BLK * CPAD);     // [8][S1]
    double2* Pt = P1 + (size_t)FM_SB * S1;                             // [NSB][ndp]
    double* acc = reinterpret_cast<double*>(Pt + (size_t)ndp * NSB);   // [NMG][2][ndp] product, then [NCG][4][ndp] correction
    unsigned long long* Th = reinterpret_cast<unsigned long long*>(acc + (size_t)NACC * ndp);   // [ndp] fixed-point phase at the tile start
    unsigned long long* D64 = Th + ndp;                                // [ndp] fixed-point phase step of one sample
    int* Ek = reinterpret_cast<int*>(D64 + ndp);                       // [ndp]
    int* blkflag = Ek + ndp;                                           // [NBLK] 1: a sample of the block has |delta| > 8 ulp(t)
    double* Ws = reinterpret_cast<double*>(blkflag + NBLK);           // [ndp] w_k (0 beyond nd)
    float2* Q32 = reinterpret_cast<float2*>(Ws + ndp);                 // [32][ndp] (cos, sin)(w_k r dt); 16-byte aligned rows (ndp % 4 == 0)

    __shared__ int fm_next;                                            // next task of the current tile
    const int tid = threadIdx.x, lane = tid & 31;
    // warp index through a warp reduction: REDUX writes a uniform register, so ptxas knows the value (and every
    // branch derived from it) is warp-uniform
    const int warp = (int)__reduce_max_sync(0xffffffffu, (unsigned)(tid >> 5));
    const double mean_u = sums ? sums[0] * inv_count : 0.0;
    const double mean_y = sums ? sums[1] * inv_count : 0.0;

    for (int e = tid; e < 2 * ndp * FM_BLK; e += FM_THREADS) Qt[(e >> 5) * FM_LD + (e & 31)] = T.qt[e];
    for (int e = tid; e < ndp * FM_SB; e += FM_THREADS) P1[(e / ndp) * S1 + e % ndp] = T.p1[(e / ndp) * FM_NDMAX + e % ndp];
    for (int e = tid; e < NACC * ndp; e += FM_THREADS) acc[e] = 0.0;
    for (int e = tid; e < ndp; e += FM_THREADS) { D64[e] = T.d64[e]; Ek[e] = T.ek[e]; Ws[e] = T.wpad[e]; }
    for (int e = tid; e < FM_BLK * ndp; e += FM_THREADS) Q32[e] = T.q32[(e / ndp) * FM_NDMAX + e % ndp];

    for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const int64_t i0 = own_begin + (int64_t)tl * TS;
        const int cnt = (int)min((int64_t)TS, own_end - i0);
        __syncthreads();                                               // previous tile consumed, tables in place
        if (tid < NBLK) blkflag[tid] = 0;
        if (tid == 0) fm_next = 0;
        // next tile of this CTA: pull its lines into L2 while this one is processed
        {
            const int64_t nx = i0 + (int64_t)gridDim.x * TS;
            if (nx < own_end) {
                const int64_t nl = min((int64_t)TS, own_end - nx);
                for (int64_t off = (int64_t)tid * 16; off < nl; off += (int64_t)FM_THREADS * 16) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(t + nx + off));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(u + nx + off));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(y + nx + off));
                }
            }
        }
        __syncthreads();
        // ---- stage the tile: weighted, mean-free samples as FP64 rows (matrix product) and, per sample, the
        //      correction inputs, transposed so that lanes = blocks read rows
#pragma unroll 4
        for (int j = tid; j < TS; j += FM_THREADS) {
            double dl = 0.0, au = 0.0, ay = 0.0;
            if (j < cnt) {
                const int64_t i = i0 + j;
                const double ti = __ldg(t + i);
                const double tm = __ldg(t + (i > 0 ? i - 1 : 0)), tp = __ldg(t + (i < n - 1 ? i + 1 : n - 1));
                const double wt = 0.5 * (tp - tm);
                const double k = (double)i;
                const double ph = k * dt, pl = fma(k, dt, -ph);
                const double d = ti - t0, bb = d - ti;
                const double derr = (ti - (d - bb)) + (-t0 - bb);
                dl = ((d - ph) - pl) + derr;
                au = wt * (__ldg(u + i) - mean_u);
                ay = wt * (__ldg(y + i) - mean_y);
            }
            const int b = j >> 5, r = j & 31;
            As[b * FM_LD + r] = au;
            As[(NBLK + b) * FM_LD + r] = ay;
            const double dm = dl * fm_pow2(fm_block_m(i0 + (j & ~31), t0, dt));
            int mi = 0;
            if (fabs(dm) < (double)FM_MLIM) mi = __double2int_rn(dm);
            else atomicOr(&blkflag[b], 1);                             // also NaN
            // the float copy is the QUANTISED jitter: the correction uses m g both inside the fraction and as the w delta
            // term, so the quantisation error cancels (eps = w delta_q - e(delta_q) = w delta - e(delta) unless the
            // rounding direction flips, which the tie test below excludes)
            cin[r * CPAD + b] = make_float4(__int_as_float(mi), (float)mi, (float)au, (float)ay);
        }
        // ---- exact phasor of every superblock start: double-double seed at the tile start, then rotations;
        //      and the tile start's phase in 64-bit fixed point
        for (int k = tid; k < ndp; k += FM_THREADS) {
            const double wk = Ws[k];
            const DD x = dd_time(t0, (double)i0, dt);
            double s, c;
            sincos_dd(wk, x, s, c);
            const double2 rt = T.rot[k];
            Pt[k] = make_double2(c, s);
#pragma unroll
            for (int sb = 1; sb < NSB; ++sb) {
                const double cn = fma(c, rt.x, -(s * rt.y));
                s = fma(s, rt.x, c * rt.y);
                c = cn;
                Pt[sb * ndp + k] = make_double2(c, s);
            }
            Th[k] = T.a64[k] + (unsigned long long)i0 * D64[k];
        }
        __syncthreads();

        // ================= matrix product: task = (m-group of 16 rows, n-group of <= 5 n-tiles of 4 bins)
        const int n_ntiles = ndp >> 2;
        const int n_ngroups = (n_ntiles + FM_NGRP - 1) / FM_NGRP;
        const int n_mma_tasks = NMG * n_ngroups;
        auto mma_task = [&](int task) {
            {
                const int g4 = lane >> 2, l4 = lane & 3;
                const int mg = task % NMG, ng = task / NMG;
                const int nt0 = ng * FM_NGRP;
                const int ntn = min(FM_NGRP, n_ntiles - nt0);
                double c0[FM_NGRP][2], c1[FM_NGRP][2];
#pragma unroll
                for (int q = 0; q < FM_NGRP; ++q) { c0[q][0] = c0[q][1] = c1[q][0] = c1[q][1] = 0.0; }
                const double* arow0 = As + (mg * 16 + g4) * FM_LD + l4;
                const double* arow1 = arow0 + 8 * FM_LD;
                const double* brow[FM_NGRP];
#pragma unroll
                for (int q = 0; q < FM_NGRP; ++q) brow[q] = Qt + ((nt0 + (q < ntn ? q : 0)) * 8 + g4) * FM_LD + l4;
                // k loop kept rolled with the next step's fragments loaded ahead: inside one step the 10 DMMAs are
                // independent (10 accumulator tiles), so the FP64 pipe never waits on an accumulator; fully unrolled,
                // ptxas chains the steps of one tile back to back and the pipe runs at half rate
                double a0n = arow0[0], a1n = arow1[0], bn[FM_NGRP];
#pragma unroll
                for (int q = 0; q < FM_NGRP; ++q) bn[q] = brow[q][0];
#pragma unroll 1
                for (int ks = 0; ks < FM_BLK / 4; ++ks) {
                    const double a0 = a0n, a1 = a1n;
                    double bq[FM_NGRP];
#pragma unroll
                    for (int q = 0; q < FM_NGRP; ++q) bq[q] = bn[q];
                    const int kn = (ks + 1 < FM_BLK / 4 ? ks + 1 : ks) * 4;
                    a0n = arow0[kn]; a1n = arow1[kn];
#pragma unroll
                    for (int q = 0; q < FM_NGRP; ++q) bn[q] = brow[q][kn];
#pragma unroll
                    for (int q = 0; q < FM_NGRP; ++q) {
                        dmma884(c0[q][0], c0[q][1], a0, bq[q]);
                        dmma884(c1[q][0], c1[q][1], a1, bq[q]);
                    }
                }
                // ---- block phasors on the fragment: lane holds (sum a cos, sum a sin) of block row g4, bin 4 nt + l4
                const int mt0 = mg * 2;
                const int sb0 = mt0 % NSB;                             // both m-tiles belong to the same signal (NSB is even)
                double* aslot = acc + (size_t)mg * 2 * ndp;            // rows (cos sum, sin sum) of this m-group
#pragma unroll
                for (int q = 0; q < FM_NGRP; ++q) {
                    if (q < ntn) {
                        const int k = (nt0 + q) * 4 + l4;
                        const double2 p1 = P1[g4 * S1 + k];
                        const double2 pa = Pt[sb0 * ndp + k], pb = Pt[(sb0 + 1) * ndp + k];
                        // v = C * p1 (complex), z = v * pt
                        const double v0r = fma(c0[q][0], p1.x, -(c0[q][1] * p1.y)), v0i = fma(c0[q][0], p1.y, c0[q][1] * p1.x);
                        const double v1r = fma(c1[q][0], p1.x, -(c1[q][1] * p1.y)), v1i = fma(c1[q][0], p1.y, c1[q][1] * p1.x);
                        double zr = fma(v0r, pa.x, -(v0i * pa.y)), zi = fma(v0r, pa.y, v0i * pa.x);
                        zr += fma(v1r, pb.x, -(v1i * pb.y));
                        zi += fma(v1r, pb.y, v1i * pb.x);
#pragma unroll
                        for (int o = 4; o < 32; o <<= 1) {
                            zr += __shfl_xor_sync(0xffffffffu, zr, o);
                            zi += __shfl_xor_sync(0xffffffffu, zi, o);
                        }
                        if (g4 == 0) {
                            aslot[k] += zr;
                            aslot[ndp + k] += zi;
                        }
                    }
                }
            }
        };
        // ================= correction: task = (group of 32 blocks, chunk of FM_KC bins); lane = block
        const int n_chunks = (nd + FM_KC - 1) / FM_KC;
        const int n_cor_tasks = NCG * n_chunks;
        auto cor_task = [&](int task) {
            {
                const int g = task % NCG, k0 = (task / NCG) * FM_KC;
                const int b = g * 32 + lane;
                const int64_t ib = i0 + 32 * b;                       // global index of the block's first sample
                const int Mb = fm_block_m(ib, t0, dt);
                const int bflag = blkflag[b];
                // first and last owned time stamp of the block (they bound |w t| inside it): rebuilt from the staged
                // jitter, read back only when the jitter was out of the staged range
                const int r_last = (int)min((int64_t)(FM_BLK - 1), max((int64_t)0, own_end - 1 - ib));
                double tf, tl_;
                if (bflag) {
                    tf = __ldg(t + min(ib, n - 1));
                    tl_ = __ldg(t + min(ib + r_last, n - 1));
                } else {
                    tf = fm_time_of(ib, t0, dt, (double)cin[b].y * fm_pow2(-Mb));
                    tl_ = fm_time_of(ib + r_last, t0, dt, (double)cin[r_last * CPAD + b].y * fm_pow2(-Mb));
                }
                unsigned X[FM_KC], dstep[FM_KC];
                int gq[FM_KC], e0[FM_KC];
                float gf[FM_KC];
                float2 au2[FM_KC], ay2[FM_KC];
                unsigned slow = 0;                                     // bit kk: this (block, bin) takes the FP64 path
#pragma unroll
                for (int kk = 0; kk < FM_KC; ++kk) {
                    const int k = k0 + kk;
                    const double wk = Ws[k < ndp ? k : 0];
                    const double p0 = wk * tf, p1 = wk * tl_;
                    const int E0 = fm_expo(p0), E1 = fm_expo(p1);
                    const int ekk = Ek[k < ndp ? k : 0];
                    const int ps = E0 - ekk + 32;
                    const bool pad = k >= nd;                          // bins of the last chunk beyond nd: inert
                    const bool bad = !pad && ((E0 != E1) || ps < 0 || ps > 32 || bflag != 0 || wk == 0.0);
                    const int psc = min(max(ps, 0), 32);
                    const unsigned long long dk = D64[k < ndp ? k : 0];
                    const unsigned long long thb = Th[k < ndp ? k : 0] + (unsigned long long)(32 * b) * dk;
                    X[kk] = pad ? 0u : (unsigned)(thb >> psc);
                    dstep[kk] = pad ? 0u : (unsigned)(dk >> psc);
                    const int Es = pad ? 0 : (bad ? max(max(E0, E1), -900) : E0);   // scale exponent of this (block, bin)
                    e0[kk] = Es;
                    const double gd = wk * fm_pow2(min(max(84 - Es - Mb, -1000), 1000));
                    gq[kk] = (bad || pad) ? 0 : __double2int_rn(gd);
                    gf[kk] = (float)gq[kk];
                    if (bad) slow |= 1u << kk;
                    au2[kk] = make_float2(0.0f, 0.0f);
                    ay2[kk] = make_float2(0.0f, 0.0f);
                }
#pragma unroll 4
                for (int r = 0; r < FM_BLK; ++r) {
                    const float4 ci = cin[r * CPAD + b];
                    const int mi = __float_as_int(ci.x);
                    const float2 a_u = make_float2(ci.z, ci.z), a_y = make_float2(ci.w, ci.w);
                    unsigned tk[FM_KC], tmin = 0xffffffffu;
                    float2 qq[FM_KC];
#pragma unroll
                    for (int j = 0; j < FM_KC / 2; ++j) {
                        const float4 q4 = reinterpret_cast<const float4*>(Q32)[(r * ndp + k0) / 2 + j];
                        qq[2 * j] = make_float2(q4.x, q4.y);
                        qq[2 * j + 1] = make_float2(q4.z, q4.w);
                    }
#pragma unroll
                    for (int kk = 0; kk < FM_KC; ++kk) {
                        const int Xj = (int)((unsigned)mi * (unsigned)gq[kk] + X[kk]);
                        X[kk] += dstep[kk];
                        // distance of the centred fraction from the rounding ties +-1/2, as an unsigned number that is
                        // small near either of them
                        tk[kk] = (unsigned)Xj + FM_TIE_OFF;
                        tmin = min(tmin, tk[kk]);
                        const float epsn = fmaf(ci.y, gf[kk], -(float)Xj);
                        const float2 z = __fmul2_rn(qq[kk], make_float2(epsn, epsn));
                        au2[kk] = __ffma2_rn(a_u, z, au2[kk]);
                        ay2[kk] = __ffma2_rn(a_y, z, ay2[kk]);
                    }
                    if (tmin < FM_TIE_WIN) {
                        // some fraction is within 2^-13 of a tie: the rounding direction is not safe in fixed point;
                        // replace that sample-bin's eps by the FP64 one (rare: 2^-12 of the sample-bins)
                        const double ti = fm_time_of(ib + r, t0, dt, (double)ci.y * fm_pow2(-Mb));
                        const double dl = fm_delta_of(ti, ib + r, t0, dt);
#pragma unroll
                        for (int kk = 0; kk < FM_KC; ++kk) {
                            if (tk[kk] < FM_TIE_WIN && !((slow >> kk) & 1u)) {
                                const double wk = Ws[k0 + kk];
                                const double p = wk * ti;
                                const double e = fma(wk, ti, -p);
                                const float exact = (float)(fma(wk, dl, -e) * fm_pow2(84 - e0[kk]));
                                const float fast = fmaf(ci.y, gf[kk], -(float)(int)(tk[kk] - FM_TIE_OFF));
                                const float de = exact - fast;
                                const float2 q = Q32[r * ndp + k0 + kk];
                                const float2 z = make_float2(q.x * de, q.y * de);
                                au2[kk].x = fmaf(ci.z, z.x, au2[kk].x); au2[kk].y = fmaf(ci.z, z.y, au2[kk].y);
                                ay2[kk].x = fmaf(ci.w, z.x, ay2[kk].x); ay2[kk].y = fmaf(ci.w, z.y, ay2[kk].y);
                            }
                        }
                    }
                }
                if (slow) {
                    // FP64 evaluation of eps for the (block, bin) pairs the fixed-point path cannot serve
#pragma unroll
                    for (int kk = 0; kk < FM_KC; ++kk) {
                        if ((slow >> kk) & 1u) {
                            const double wk = Ws[k0 + kk];
                            const double sc = fm_pow2(84 - e0[kk]);
                            float2 su = make_float2(0.0f, 0.0f), sy = make_float2(0.0f, 0.0f);
                            for (int r = 0; r < FM_BLK; ++r) {
                                const float4 ci = cin[r * CPAD + b];
                                const int64_t i = ib + r;
                                double ti, dl;
                                if (bflag) {
                                    // jitter beyond the staged range: the time stamp itself, and delta_i again in FP64
                                    ti = __ldg(t + min(i, n - 1));
                                    dl = fm_delta_of(ti, i, t0, dt);
                                } else {
                                    ti = fm_time_of(i, t0, dt, (double)ci.y * fm_pow2(-Mb));
                                    dl = fm_delta_of(ti, i, t0, dt);
                                }
                                const double p = wk * ti;
                                const double e = fma(wk, ti, -p);
                                const float epsn = (float)(fma(wk, dl, -e) * sc);
                                const float2 q = Q32[r * ndp + k0 + kk];
                                const float2 z = make_float2(q.x * epsn, q.y * epsn);
                                su.x = fmaf(ci.z, z.x, su.x); su.y = fmaf(ci.z, z.y, su.y);
                                sy.x = fmaf(ci.w, z.x, sy.x); sy.y = fmaf(ci.w, z.y, sy.y);
                            }
                            au2[kk] = su;
                            ay2[kk] = sy;
                        }
                    }
                }
                // ---- block phasor and scale, then the sum over the 32 blocks of this lane group
                float vals[4 * FM_KC];
#pragma unroll
                for (int kk = 0; kk < FM_KC; ++kk) {
                    const int k = min(k0 + kk, ndp - 1);
                    const double2 p1 = P1[(b & 7) * S1 + k];
                    const double2 pt = Pt[(b >> 3) * ndp + k];
                    const float sc = (float)fm_pow2(e0[kk] - 84);
                    const float pr = sc * __double2float_rn(fma(p1.x, pt.x, -(p1.y * pt.y)));
                    const float pi = sc * __double2float_rn(fma(p1.x, pt.y, p1.y * pt.x));
                    // (sum a' cos, sum a' sin) of the full phase, a' = a eps
                    const float rcu = au2[kk].x * pr - au2[kk].y * pi, rsu = au2[kk].x * pi + au2[kk].y * pr;
                    const float rcy = ay2[kk].x * pr - ay2[kk].y * pi, rsy = ay2[kk].x * pi + ay2[kk].y * pr;
                    // cos(theta + eps) = cos - eps sin ; sin(theta + eps) = sin + eps cos
                    vals[4 * kk + 0] = -rsu; vals[4 * kk + 1] = rcu; vals[4 * kk + 2] = -rsy; vals[4 * kk + 3] = rcy;
                }
                // sum over the 32 blocks with a halving butterfly: after the step with lane mask o a lane keeps half of the
                // values (chosen by its bit o), so 8 + 4 + 2 + 1 + 1 shuffles instead of 16 x 5
                static_assert(FM_KC == 4, "the butterfly below is written for 16 values");
#pragma unroll
                for (int o = 16, cntv = 16; o > 1; o >>= 1, cntv >>= 1) {
                    const bool up = (lane & o) != 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i < cntv / 2) {
                            const float lo = vals[i], hi = vals[i + cntv / 2];
                            vals[i] = (up ? hi : lo) + __shfl_xor_sync(0xffffffffu, up ? lo : hi, o);
                        }
                    }
                }
                vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], 1);
                {
                    // value index held by this lane: bit 16 chose between halves of 16, bit 8 of 8, bit 4 of 4, bit 2 of 2
                    const int vi = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    const int k = k0 + (vi >> 2), c = vi & 3;
                    double* aslot = acc + (size_t)(NMG * 2 + g * 4) * ndp;
                    if ((lane & 1) == 0 && k < nd) aslot[c * ndp + k] += (double)vals[0];
                }
            }
        };
        // ---- schedule: one list with the two kinds of task interleaved (product, correction, product, ...), handed out
        //      through a shared counter, so every scheduler holds warps of both kinds most of the time (the FP64 pipe and
        //      the integer / FP32 pipes overlap) and no warp waits for a slower one until the list is empty.  Every
        //      accumulator entry belongs to exactly one task, so the result does not depend on who runs what.
        {
            const int n_pair = min(n_mma_tasks, n_cor_tasks), n_all = n_mma_tasks + n_cor_tasks;
            for (;;) {
                int id = 0;
                if (lane == 0) id = atomicAdd(&fm_next, 1);
                id = __shfl_sync(0xffffffffu, id, 0);
                if (id >= n_all) break;
                bool is_mma; int idx;
                if (id < 2 * n_pair) { is_mma = (id & 1) == 0; idx = id >> 1; }
                else { is_mma = n_mma_tasks > n_cor_tasks; idx = id - n_pair; }
                if (is_mma) {
#if !(FM_VARIANT & 1)
                    mma_task(idx);
#endif
                } else {
#if !(FM_VARIANT & 2)
                    cor_task(idx);
#endif
                }
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < 4 * nd; e += FM_THREADS) {
        const int comp = e / nd, k = e % nd;
        const int sig = comp >> 1;                                     // components 0,1: u; 2,3: y
        double v = 0.0;
#pragma unroll
        for (int m = 0; m < NMG / 2; ++m) v += acc[((size_t)(sig * (NMG / 2) + m) * 2 + (comp & 1)) * ndp + k];
#pragma unroll
        for (int g = 0; g < NCG; ++g) v += acc[((size_t)NMG * 2 + g * 4 + comp) * ndp + k];
        cta_partial[(size_t)blockIdx.x * 4 * nd + e] = v;
    }
}

}  // namespace b200id
