// gp_batch_tiled.cu -- K2 fast path (n <= 127): one CTA of 4 warps per candidate, left-looking BLOCKED
// Cholesky in shared memory with the trailing updates on the FP64 tensor pipe.
//
// Reference call sites: src/gpr/grid_search.py:91-133 (candidate metric), src/gpr/gpr_fitting.py:104-123.
//
// Layout in shared memory.  The (n+1) x n array [K + noise*I ; y^T] is stored by BLOCK COLUMNS of 16:
// block column J keeps rows [16J, R) (R = n+1 rounded up to 4) column-major with leading dimension
// ld_J = R - 16J, so a thread's 4 consecutive rows of one column are one 32-byte vector and the upper
// triangle above each block is never stored (50 KB at n = 100 -> 4 CTAs per SM, which the launch bound keeps).
//
// Gram entries: evaluated per candidate (gt_fill), or -- b200id_gp_batch_eval_shared, product grids -- taken from a
// table of covariance SHAPES (gp_shape_table_kernel, one table per distinct theta-without-amplitude) and scaled by the
// candidate's amplitude with the reference's own final multiplication; same bits either way.
//
// Per block column J:
//   update : C[i, c] -= sum_{k < 16J} L[i, k] L[16J + c, k].  Each warp owns 32 rows x 16 columns as 4 x 2
//            accumulator fragments of mma.sync.m8n8k4.f64 (GT_DMMA; the register-tiled DFMA loop it replaced, 4 x 4
//            per lane, is kept under GT_DMMA=0); both add the products in k order: the factor is bit-identical.
//   diag   : one warp factorises the 16 x 16 diagonal block in registers (shuffles, reciprocal scaling; straight-line
//            code for a full block) while the other warps already apply block columns < J to block column J+1 (lookahead).
//   panel  : one thread per row below solves x L_JJ^T = c by forward substitution in registers, multiplying
//            by the stored reciprocal diagonal (as OpenBLAS TRSM does).
// Row n carries y (row n + 1 a second target), so after the last block column it holds z = L^-1 y and y^T K^-1 y = z.z.
// Validation RMSE: blocked back substitution (16 x 16 solves in registers + shuffles, columns prefetched a block ahead),
// then K(Xv, X) alpha with the entries evaluated on the fly or read from the shape table.
// Failure conventions as in gp_linalg.cuh: pivot <= 0 or +inf -> 1e10, NaN pivot -> NaN metric.
// Measured structure (profiles/r02_ncu_gp_tiled.txt): barrier bound -- three warps wait while one walks a serial chain.
#include "common.cuh"
#include "gp_kernels.cuh"

namespace b200id {

// Optional phase timing (build with -DGT_PROFILE): per-warp clock64 deltas accumulated per phase.
#ifdef GT_PROFILE
__device__ unsigned long long g_gt_prof[16];
#define GT_TICK(phase)                                                                    \
    do {                                                                                  \
        const long long _now = clock64();                                                 \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_gt_prof[phase], (unsigned long long)(_now - _t0)); \
        _t0 = _now;                                                                       \
    } while (0)
#define GT_TICK_INIT long long _t0 = clock64();
#else
#define GT_TICK(phase)
#define GT_TICK_INIT
#endif

// independent covariance entries in flight per thread in the Gram fill and the validation prediction (each entry is a
// ~45-instruction FP64 chain through exp; with 16 warps per SM nothing else covers its latency)
#ifndef GT_FILL_UNROLL
#define GT_FILL_UNROLL 1
#endif
#ifndef GT_PRED_UNROLL
#define GT_PRED_UNROLL 4
#endif
constexpr int GT_FILL_UNROLL_N = GT_FILL_UNROLL;
constexpr int GT_PRED_UNROLL_N = GT_PRED_UNROLL;
#ifndef GT_DMMA
#define GT_DMMA 1
#endif
#ifndef GT_MIN_CTAS
#define GT_MIN_CTAS 4
#endif
constexpr int GT_THREADS = 128;
constexpr int GT_NB = 16;
constexpr int GT_NMAX = 127;            // n + 1 rows <= 128 = 4 warps x 32 rows
constexpr int GT_METRIC_FIT = 2;        // internal mode: write alpha and L instead of a metric

struct TiledLayout {
    int R;                              // rows incl. y row, rounded up to a multiple of 4
    int nblk;                           // number of block columns
};

// rows = n matrix rows + one row per right-hand side (y; in pair mode also the second target), rounded up to 4
__host__ __device__ inline int gt_rows(int n, int nrhs = 1) { return (n + nrhs + 3) & ~3; }
__host__ __device__ inline int gt_nblk(int n) { return (n + GT_NB - 1) / GT_NB; }
// offset (in doubles) of block column J: sum_{q<J} 16 * (R - 16 q)
__host__ __device__ inline int gt_off(int R, int J) { return GT_NB * (J * R - GT_NB * J * (J - 1) / 2); }
static inline size_t gt_smem_bytes(int n, int nrhs = 1) {
    const int R = gt_rows(n, nrhs), nb = gt_nblk(n);
    return ((size_t)gt_off(R, nb) + (size_t)(2 + nrhs) * (size_t)(n + 4) + 16) * sizeof(double);
}

// Second regression target sharing the candidate's factorisation (pair mode): the Re and Im searches of the
// pipeline use the same inputs, candidates and noise, so K + noise I and its Cholesky factor are common and only
// the right-hand side differs.  All-null when unused.
struct GtSecond {
    const double* y;            // [n] second target (NULL -> single-target mode)
    const double* yv;           // [nv] its validation values (VAL_RMSE)
    double y_mean, y_scale;
    double* metric_out;         // [C]
};

// Gram fill of the block-column storage, kernel id fixed at compile time.  Thread (c = tid / 8, r8 = tid % 8)
// walks rows r8, r8 + 8, ... of column c of each block column: no integer division, 8 consecutive rows per
// column are contiguous in shared memory.
template <int ID>
__device__ __forceinline__ void gt_fill(const KernelParams& kp, const double* __restrict__ xs,
                                        const double* __restrict__ y, int n, double noise, int R, int nblk,
                                        double* __restrict__ Lb, const double* __restrict__ y2) {
    const int c = threadIdx.x >> 3, r8 = threadIdx.x & 7;
    for (int J = 0; J < nblk; ++J) {
        const int ld = R - GT_NB * J;
        double* col = Lb + gt_off(R, J) + c * ld;
        const int j = GT_NB * J + c;
        const double xj = j < n ? xs[j] : 0.0;
        const double yj = j < n ? y[j] : 0.0;
        const double y2j = (y2 && j < n) ? y2[j] : 0.0;
#pragma unroll GT_FILL_UNROLL_N
        for (int r = r8; r < ld; r += 8) {
            const int i = GT_NB * J + r;
            double v = 0.0;
            if (j < n) {
                if (i < n) {
                    // evaluate with the larger index first, as the lower triangle of kernel(X) is laid out
                    v = (i >= j) ? kernel_entry_t<ID>(kp, xs[i], xj) : kernel_entry_t<ID>(kp, xj, xs[i]);
                    if (i == j) v = B200ID_ADD(v, noise);
                } else if (i == n) {
                    v = yj;
                } else if (i == n + 1) {
                    v = y2j;                                  // zero in single-target mode (padding row)
                }
            }
            col[r] = v;
        }
    }
}

// sum_j k(xv, x_j) alpha_j for one validation point, kernel id fixed at compile time
template <int ID>
__device__ __forceinline__ void gt_predict_point(const KernelParams& kp, double xv, const double* __restrict__ xs,
                                                 const double* __restrict__ alpha, int n, double* out) {
    double pred = 0.0;
#pragma unroll GT_PRED_UNROLL_N
    for (int j = 0; j < n; ++j) pred = fma(kernel_entry_t<ID>(kp, xv, xs[j]), alpha[j], pred);
    *out = pred;
}

// same for two coefficient vectors at once: the covariance entries are evaluated once
template <int ID>
__device__ __forceinline__ void gt_predict_point2(const KernelParams& kp, double xv, const double* __restrict__ xs,
                                                  const double* __restrict__ alpha, const double* __restrict__ alpha2,
                                                  int n, double* out, double* out2) {
    double pred = 0.0, pred2 = 0.0;
#pragma unroll GT_PRED_UNROLL_N
    for (int j = 0; j < n; ++j) {
        const double kv = kernel_entry_t<ID>(kp, xv, xs[j]);
        pred = fma(kv, alpha[j], pred);
        pred2 = fma(kv, alpha2[j], pred2);
    }
    *out = pred;
    *out2 = pred2;
}

// Shared covariance shapes (b200id_gp_batch_eval_shared).  In every kernel but DI the amplitude theta[1] is the LAST
// factor of an entry, k = fl(theta[1] * s(x, x'; other parameters)), and a grid search is a product of 1-D grids
// (grid_search.py:173): 900 Matern candidates are 30 shapes x 30 amplitudes.  The shape values -- the division and
// the exp, ~50 FP64 instructions per entry -- are evaluated ONCE per distinct shape into a table (amplitude 1: 1 * s
// is exact) and a candidate's CTA forms fl(theta[1] * s) from it: the same operations on the same operands as the
// direct evaluation, so the metrics are bit-identical.
//   table of shape q at tab + q * stride:  [ Gram in the CTA's block-column layout, zeros above the diagonal and in the
//   target / padding rows : gt_off(R, nblk) ][ K(Xv, X) shape values, entry (v, j) at j * nv + v : n * nv ]
//   (coalesced over the validation points; a row-per-point layout with vector loads measured slower)
struct GtShapes {
    const double* tab;          // NULL -> direct evaluation
    const int* shape_of;        // [C] table index of each candidate
    int64_t stride;             // doubles per table
};

__host__ __device__ inline int64_t gt_shape_stride(int n, int nv, int nrhs) {
    return ((int64_t)gt_off(gt_rows(n, nrhs), gt_nblk(n)) + (int64_t)n * nv + 1) & ~(int64_t)1;
}

template <int ID>
__device__ __forceinline__ void gt_shape_entry(const KernelParams& kp, double xa, double xb, double* out) {
    *out = kernel_entry_t<ID>(kp, xa, xb);
}

// grid (nblk + ceil(n * nv / GT_TAB_CHUNK), n_shapes): blockIdx.x < nblk fills block column blockIdx.x of the Gram part,
// the other CTAs one chunk of the validation part each
#ifndef GT_TAB_UNROLL_N
#define GT_TAB_UNROLL_N 10
#endif
constexpr int GT_TAB_UNROLL = GT_TAB_UNROLL_N;     // table loads in flight per thread in the validation prediction
constexpr int GT_TAB_THREADS = 256;
constexpr int GT_TAB_CHUNK = 2048;
__global__ void __launch_bounds__(GT_TAB_THREADS)
gp_shape_table_kernel(int kernel_id, const double* __restrict__ shape_theta, const double* __restrict__ X, int n,
                      const double* __restrict__ Xv, int nv, int nrhs, double* __restrict__ tab, int64_t stride) {
    const int q = blockIdx.y;
    double th[B200ID_MAX_PARAMS];
    for (int p = 0; p < B200ID_MAX_PARAMS; ++p) th[p] = shape_theta[(int64_t)q * B200ID_MAX_PARAMS + p];
    th[1] = 1.0;                                         // the amplitude is applied per candidate
    const KernelParams kp = make_kernel_params(kernel_id, th);
    const int R = gt_rows(n, nrhs), nblk = gt_nblk(n);
    double* out = tab + (int64_t)q * stride;
    if ((int)blockIdx.x < nblk) {
        const int J = blockIdx.x, ld = R - GT_NB * J;
        double* col = out + gt_off(R, J);
        for (int e = threadIdx.x; e < GT_NB * ld; e += GT_TAB_THREADS) {
            const int c = e / ld, r = e - c * ld;
            const int i = GT_NB * J + r, j = GT_NB * J + c;
            double v = 0.0;
            if (i < n && j < n && i >= j) { B200ID_DISPATCH_KERNEL(kernel_id, gt_shape_entry, kp, X[i], X[j], &v) }
            col[e] = v;
        }
    } else {
        double* val = out + gt_off(R, nblk);
        const int e0 = ((int)blockIdx.x - nblk) * GT_TAB_CHUNK;
        const int e1 = min(e0 + GT_TAB_CHUNK, n * nv);
        for (int e = e0 + threadIdx.x; e < e1; e += GT_TAB_THREADS) {
            const int j = e / nv, v = e - j * nv;
            double k = 0.0;
            B200ID_DISPATCH_KERNEL(kernel_id, gt_shape_entry, kp, Xv[v], X[j], &k)
            val[e] = k;
        }
    }
}

// a[k] for a run-time k out of a register array (a chain of selects: no local-memory copy of the array)
__device__ __forceinline__ double lc_at(const double (&a)[GT_NB], int k) {
    double v = a[0];
#pragma unroll
    for (int q = 1; q < GT_NB; ++q) v = (k == q) ? a[q] : v;
    return v;
}

// 4 CTAs per SM: 128 registers per thread (without the bound ptxas takes 154, i.e. 3 CTAs per SM, and the search
// loses a quarter of its throughput: 900 x 2 validation-RMSE candidates 0.42 -> 0.26 ms, 4096 NLL 0.84 -> 0.60 ms)
__global__ void __launch_bounds__(GT_THREADS, GT_MIN_CTAS)
gp_batch_tiled_kernel(int kernel_id, const int* __restrict__ kernel_ids, const double* __restrict__ theta,
                      const double* __restrict__ X, const double* __restrict__ y, int n, double noise,
                      int metric, const double* __restrict__ Xv, const double* __restrict__ yv, int nv,
                      double y_mean, double y_scale, double* __restrict__ metric_out,
                      double* __restrict__ fit_alpha, double* __restrict__ fit_L, int* __restrict__ fit_info,
                      const double* __restrict__ noise_v /*per-candidate diagonal term, or NULL*/,
                      const GtSecond sec, const GtShapes shp) {
    extern __shared__ __align__(16) double sm[];
    const int nrhs = sec.y ? 2 : 1;
    const int R = gt_rows(n, nrhs), nblk = gt_nblk(n);
    double* Lb = sm;                                   // block-column storage
    double* xs = Lb + gt_off(R, nblk);                 // [n]  inputs
    double* idg = xs + (n + 4);                        // [n]  reciprocal diagonal of L
    double* zz = idg + (n + 4);                        // [n]  z, then alpha
    double* scratch = zz + (n + 4);                    // [16]
    double* zz2 = scratch + 16;                        // [n]  second target's z / alpha (pair mode only)
    __shared__ int status;

    const int64_t cand = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    // rotate the warp -> row-range map with the CTA index so that the 4 co-resident CTAs of an SM load
    // its 4 sub-partitions evenly (the last rows do the most work)
    const int wrow = ((tid >> 5) + (int)(blockIdx.x & 3)) & 3;
    const int kid = kernel_ids ? kernel_ids[cand] : kernel_id;
    const KernelParams kp = make_kernel_params(kid, theta + cand * B200ID_MAX_PARAMS);

    GT_TICK_INIT
    if (tid == 0) status = 0;
    for (int i = tid; i < n; i += GT_THREADS) xs[i] = X[i];
    __syncthreads();

    // ---- fill: K + noise*I (rows < n), y (row n), zeros (padding rows), block column by block column
    if (noise_v) noise = noise_v[cand];
    if (metric == GT_METRIC_FIT) {
        // fit mode: CTA b factorises problem b of a small batch (theta[b], y[b]) -> alpha[b], L[b], info[b]
        y += cand * n;
        fit_alpha += cand * n;
        fit_L += cand * (int64_t)n * n;
        fit_info += cand;
    }
    const double* shape = shp.tab ? shp.tab + (int64_t)shp.shape_of[cand] * shp.stride : nullptr;
    if (shape) {
        // shared shapes: fl(amplitude * shape value) for the whole storage (zeros stay zeros), then the diagonal term
        // and the target rows
        const int G = gt_off(R, nblk);
        const double2* src = reinterpret_cast<const double2*>(shape);
        double2* dst = reinterpret_cast<double2*>(Lb);
#pragma unroll 4
        for (int e = tid; e < G / 2; e += GT_THREADS) {
            const double2 v = src[e];
            dst[e] = make_double2(B200ID_MUL(kp.p1, v.x), B200ID_MUL(kp.p1, v.y));
        }
        __syncthreads();
        for (int j = tid; j < n; j += GT_THREADS) {
            const int J = j >> 4;
            double* col = Lb + gt_off(R, J) + (j & 15) * (R - GT_NB * J) - GT_NB * J;      // col[i] = entry (i, j)
            col[j] = B200ID_ADD(col[j], noise);
            col[n] = y[j];
            if (sec.y) col[n + 1] = sec.y[j];
        }
    } else {
        B200ID_DISPATCH_KERNEL(kid, gt_fill, kp, xs, y, n, noise, R, nblk, Lb, sec.y)
    }
    GT_TICK(0);
    __syncthreads();
    GT_TICK(1);

#if GT_DMMA
    // Updates on the FP64 tensor pipe (mma.sync.m8n8k4.f64): a warp owns 32 rows x 16 columns of a block column as
    // 4 x 2 accumulator fragments of 8 x 8; lane l holds C[l / 4][2 (l % 4) .. +1] of each.  Per 4 steps of k a lane
    // loads 4 A values (its row of each 8-row block, column k + l % 4, negated) and 2 B values and the warp issues 8 DMMAs
    // = 2048 FMAs from 8 issue slots, where the register-tiled loop needed 64 DFMAs and 16 vector loads; a DMMA adds its
    // four products in k order, so the factor is bit-identical to that loop's.
    const int fr = lane >> 2, fk = lane & 3;
    const int wrow0 = 32 * wrow;                       // first row of this warp
    double acc[4][2][2];
    bool acc_valid = false;                            // warp-uniform
    const bool diag_warp = (wrow == 0);
    // rows of this warp that take part in block column J: fragments rbk with [wrow0 + 8 rbk, +8) inside [16 J, R)
    auto frag_on = [&](int J, int rbk) { return wrow0 + 8 * rbk + 7 >= GT_NB * J && wrow0 + 8 * rbk < R; };
    auto warp_on = [&](int J) { return wrow0 + 31 >= GT_NB * J && wrow0 < R; };
    auto apply_columns = [&](int J, int Jk0, int Jk1) {
        for (int Jk = Jk0; Jk < Jk1; ++Jk) {
            const int ldk = R - GT_NB * Jk;
            const double* base = Lb + gt_off(R, Jk) + fk * ldk - GT_NB * Jk;        // base[kk * ldk + row] = L[row][16 Jk + kk + fk]
            int ra[4], rbx[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) ra[i] = min(wrow0 + 8 * i + fr, R - 1);     // rows past the storage: any finite address
#pragma unroll
            for (int j = 0; j < 2; ++j) rbx[j] = min(GT_NB * J + 8 * j + fr, R - 1);
#pragma unroll
            for (int kk = 0; kk < GT_NB; kk += 4) {
                double av[4], bv[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = -base[kk * ldk + max(ra[i], GT_NB * Jk)];
#pragma unroll
                for (int j = 0; j < 2; ++j) bv[j] = base[kk * ldk + rbx[j]];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (frag_on(J, i)) {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(av[i]), "d"(bv[j]));
                    }
            }
        }
    };
    auto load_tile = [&](int J) {
        const int ldJ = R - GT_NB * J;
        const double* colJ = Lb + gt_off(R, J) - GT_NB * J;                        // colJ[c * ldJ + row]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = wrow0 + 8 * i + fr;
            const bool ok = row >= GT_NB * J && row < R;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) acc[i][j][e] = ok ? colJ[(8 * j + 2 * fk + e) * ldJ + row] : 0.0;
        }
    };
    auto store_tile = [&](int J) {
        const int ldJ = R - GT_NB * J;
        double* colJ = Lb + gt_off(R, J) - GT_NB * J;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = wrow0 + 8 * i + fr;
            if (row >= GT_NB * J && row < R) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) colJ[(8 * j + 2 * fk + e) * ldJ + row] = acc[i][j][e];
            }
        }
    };
#else
    const int rb = lane >> 2, cb = lane & 3;           // 8 row blocks x 4 column blocks of 4 x 4
    const int i0 = 32 * wrow + 4 * rb;                 // first of this lane's 4 rows

    // acc carries this lane's 4x4 tile of the NEXT block column across the diagonal/panel phases ("lookahead"):
    // while one warp factorises diagonal block J, the others already subtract block columns < J from block
    // column J+1, so only the rank-16 contribution of block column J itself remains on the critical path.
    double acc[4][4];
    bool acc_valid = false;
    // the diagonal block is factorised by the warp that owns the lowest rows (it runs out of update work first)
    const bool diag_warp = (wrow == 0);

    // subtract block columns [Jk0, Jk1) from acc (tile rows i0.., columns of block column J)
    auto apply_columns = [&](int J, int Jk0, int Jk1) {
        for (int Jk = Jk0; Jk < Jk1; ++Jk) {
            const int ldk = R - GT_NB * Jk;
            const double* pa = Lb + gt_off(R, Jk) + (i0 - GT_NB * Jk);
            const double* pb = Lb + gt_off(R, Jk) + (GT_NB * (J - Jk) + 4 * cb);
#pragma unroll 4
            for (int kk = 0; kk < GT_NB; ++kk) {
                const double2 a01 = *reinterpret_cast<const double2*>(pa + kk * ldk);
                const double2 a23 = *reinterpret_cast<const double2*>(pa + kk * ldk + 2);
                const double2 b01 = *reinterpret_cast<const double2*>(pb + kk * ldk);
                const double2 b23 = *reinterpret_cast<const double2*>(pb + kk * ldk + 2);
                const double a[4] = {a01.x, a01.y, a23.x, a23.y};
                const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = fma(-a[p], b[q], acc[p][q]);
            }
        }
    };
    auto load_tile = [&](int J) {
        const int ldJ = R - GT_NB * J;
        const double* colJ = Lb + gt_off(R, J);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double2 v01 = *reinterpret_cast<const double2*>(colJ + (4 * cb + b) * ldJ + (i0 - GT_NB * J));
            const double2 v23 = *reinterpret_cast<const double2*>(colJ + (4 * cb + b) * ldJ + (i0 - GT_NB * J) + 2);
            acc[0][b] = v01.x; acc[1][b] = v01.y; acc[2][b] = v23.x; acc[3][b] = v23.y;
        }
    };

#endif
    for (int J = 0; J < nblk; ++J) {
        const int ldJ = R - GT_NB * J;
        double* colJ = Lb + gt_off(R, J);
        const int ncols = min(GT_NB, n - GT_NB * J);
        // ---------------- finish the update of block column J: the lookahead covered block columns < J-1
#if GT_DMMA
        if (J > 0 && warp_on(J)) {
            if (!acc_valid) load_tile(J);
            apply_columns(J, acc_valid ? J - 1 : 0, J);
            store_tile(J);
        }
#else
        if (J > 0 && i0 >= GT_NB * J && i0 < R) {
            if (!acc_valid) load_tile(J);
            apply_columns(J, acc_valid ? J - 1 : 0, J);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double* dst = colJ + (4 * cb + b) * ldJ + (i0 - GT_NB * J);
                *reinterpret_cast<double2*>(dst) = make_double2(acc[0][b], acc[1][b]);
                *reinterpret_cast<double2*>(dst + 2) = make_double2(acc[2][b], acc[3][b]);
            }
        }
#endif
        acc_valid = false;
        GT_TICK(2);
        __syncthreads();
        GT_TICK(3);
        // ---------------- diagonal block (warp 0), in registers: lane r owns row r (d[c] = D[r][c], c <= r).
        // Right-looking; column k of L is broadcast lane by lane with shuffles, so the only serial latency
        // per column is one sqrt/rsqrt pair.  Scaling multiplies by the reciprocal like OpenBLAS potf2.
        if (diag_warp) {
            double d[GT_NB];
#pragma unroll
            for (int c = 0; c < GT_NB; ++c)
                d[c] = (lane < ncols && c <= lane) ? colJ[c * ldJ + lane] : 0.0;
            int fail = 0;
            if (ncols == GT_NB) {
                // full block: straight-line code, no branch per pivot (the 16 pivots are the critical path of the
                // block column; a failed pivot only raises the flag, what follows it is discarded).  The next pivot is
                // broadcast from lane k + 1's own product l * l, ahead of the 15 shuffles of the trailing update
                // (same value as its update below: its partner in that update is its own l).
                double piv = __shfl_sync(0xffffffffu, d[0], 0);
#pragma unroll
                for (int k = 0; k < GT_NB; ++k) {
                    const bool bad = (piv <= 0.0) || (piv == INFINITY);     // NaN compares false: the factorisation goes on
                    fail = (fail == 0 && bad) ? k + 1 : fail;
                    double inv = rsqrt(piv);
                    double dk = piv * inv;
                    dk = fma(0.5 * inv, fma(-dk, dk, piv), dk);
                    inv = fma(inv, fma(-dk, inv, 1.0), inv);
                    const double l = (lane == k) ? dk : d[k] * inv;
                    d[k] = l;
                    if (lane == k) idg[GT_NB * J + k] = inv;
                    if (k + 1 < GT_NB) piv = __shfl_sync(0xffffffffu, fma(-l, l, d[k + 1]), k + 1);
#pragma unroll
                    for (int c = k + 1; c < GT_NB; ++c) {
                        const double lc = __shfl_sync(0xffffffffu, l, c);
                        d[c] = fma(-l, lc, d[c]);
                    }
                }
            } else {
#pragma unroll
            for (int k = 0; k < GT_NB; ++k) {
                if (k < ncols && fail == 0) {
                    const double piv = __shfl_sync(0xffffffffu, d[k], k);
                    if (piv <= 0.0 || piv == INFINITY) {        // NaN compares false: the factorisation goes on
                        fail = k + 1;
                    } else {
                        // sqrt(piv) and its reciprocal from ONE rsqrt + Newton steps: the IEEE sqrt followed by an
                        // IEEE division is ~300 cycles of serial latency per pivot, and the n pivots are the
                        // critical path of the whole factorisation.  dk and inv end up within 1 ulp.
                        double inv = rsqrt(piv);
                        double dk = piv * inv;
                        dk = fma(0.5 * inv, fma(-dk, dk, piv), dk);
                        inv = fma(inv, fma(-dk, inv, 1.0), inv);
                        const double l = (lane == k) ? dk : d[k] * inv;
                        d[k] = l;
                        if (lane == k) idg[GT_NB * J + k] = inv;
#pragma unroll
                        for (int c = k + 1; c < GT_NB; ++c) {
                            const double lc = __shfl_sync(0xffffffffu, l, c);
                            d[c] = fma(-l, lc, d[c]);
                        }
                    }
                }
            }
            }
            if (fail != 0) {
                if (lane == 0) status = GT_NB * J + fail;
            } else {
#pragma unroll
                for (int c = 0; c < GT_NB; ++c)
                    if (lane < ncols && c <= lane) colJ[c * ldJ + lane] = d[c];
            }
        }
        GT_TICK(4);
        // ---------------- lookahead: block column J+1 minus block columns < J (all final), kept in registers
#if GT_DMMA
        if (J + 1 < nblk && J >= 1 && !diag_warp && warp_on(J + 1)) {
            load_tile(J + 1);
            apply_columns(J + 1, 0, J);
            acc_valid = true;
        }
#else
        if (J + 1 < nblk && J >= 1 && i0 >= GT_NB * (J + 1) && i0 < R) {
            load_tile(J + 1);
            apply_columns(J + 1, 0, J);
            acc_valid = true;
        }
#endif
        GT_TICK(5);
        __syncthreads();
        GT_TICK(6);
        if (status != 0) {
            if (tid == 0) {
                if (metric == GT_METRIC_FIT) fit_info[0] = status;
                else {
                    metric_out[cand] = B200ID_FAIL_METRIC;
                    if (sec.y) sec.metric_out[cand] = B200ID_FAIL_METRIC;
                }
            }
            return;
        }
        // ---------------- panel: rows below the diagonal block (incl. the y row), one thread per row
        {
            // rows start right after the ncols diagonal rows: in a partial last block the y row (offset n - 16J)
            // sits inside the 16-row range of the block
            const int r = ncols + tid;                  // row offset inside the block column
            if (r < ldJ && GT_NB * J + r < n + nrhs) {
                double x[GT_NB];
                if (ncols == GT_NB) {
                    // full block: no per-column tests
#pragma unroll
                    for (int c = 0; c < GT_NB; ++c) x[c] = colJ[c * ldJ + r];
#pragma unroll
                    for (int c = 0; c < GT_NB; ++c) {
                        x[c] *= idg[GT_NB * J + c];
#pragma unroll
                        for (int q = c + 1; q < GT_NB; ++q) x[q] = fma(-x[c], colJ[c * ldJ + q], x[q]);
                    }
#pragma unroll
                    for (int c = 0; c < GT_NB; ++c) colJ[c * ldJ + r] = x[c];
                } else {
#pragma unroll
                for (int c = 0; c < GT_NB; ++c) x[c] = (c < ncols) ? colJ[c * ldJ + r] : 0.0;
#pragma unroll
                for (int c = 0; c < GT_NB; ++c) {
                    if (c < ncols) {
                        // right-looking inside the row: x[c] is final, push it into the later columns at once
                        // (independent FMAs, so the dependent chain is 16 steps, not 136)
                        x[c] *= idg[GT_NB * J + c];
#pragma unroll
                        for (int q = c + 1; q < GT_NB; ++q)
                            if (q < ncols) x[q] = fma(-x[c], colJ[c * ldJ + q], x[q]);
                    }
                }
#pragma unroll
                for (int c = 0; c < GT_NB; ++c)
                    if (c < ncols) colJ[c * ldJ + r] = x[c];
                }
            }
        }
        GT_TICK(7);
        __syncthreads();
        GT_TICK(8);
    }

    // L(i, j) for i >= j lives at Lb[gt_off(R, j/16) + (j%16) * (R - 16 (j/16)) + (i - 16 (j/16))]
    auto Lat = [&](int i, int j) -> double& {
        const int J = j >> 4;
        return Lb[gt_off(R, J) + (j & 15) * (R - GT_NB * J) + (i - GT_NB * J)];
    };

    if (metric == B200ID_METRIC_NLL) {
        // 0.5 y.alpha + sum log L_ii + 0.5 n log(2 pi), with y.alpha = z.z        grid_search.py:128-130
        double part = 0.0, part2 = 0.0;
        for (int j = tid; j < n; j += GT_THREADS) {
            const double zj = Lat(n, j);
            const double lg = log(Lat(j, j));
            part += 0.5 * zj * zj + lg;
            if (sec.y) {
                const double z2 = Lat(n + 1, j);
                part2 += 0.5 * z2 * z2 + lg;
            }
        }
        const double tot = block_sum(part, scratch);
        if (tid == 0) metric_out[cand] = tot + 0.5 * (double)n * 1.8378770664093453;
        if (sec.y) {
            __syncthreads();
            const double tot2 = block_sum(part2, scratch);
            if (tid == 0) sec.metric_out[cand] = tot2 + 0.5 * (double)n * 1.8378770664093453;
        }
        return;
    }

    // ---- validation RMSE: alpha = L^-T z by blocked back substitution, then K(Xv, X) alpha
    for (int j = tid; j < n; j += GT_THREADS) {
        zz[j] = Lat(n, j);
        if (sec.y) zz2[j] = Lat(n + 1, j);
    }
    __syncthreads();
    // Column `lane` of diagonal block J below its diagonal, lc[k] = L[16 J + k][16 J + lane] for lane < k < ncols (zero
    // elsewhere): 16 consecutive doubles per lane, fetched as 8 vector loads.  Warp 0 fetches the NEXT block's column
    // while all warps run the update of the current one.
    auto load_lc = [&](int J, int ncols, double (&lc)[GT_NB]) {
        const double2* colp = reinterpret_cast<const double2*>(Lb + gt_off(R, J) + min(lane, GT_NB - 1) * (R - GT_NB * J));
#pragma unroll
        for (int q = 0; q < GT_NB / 2; ++q) {
            const double2 v = colp[q];
            lc[2 * q] = (lane < 2 * q && 2 * q < ncols) ? v.x : 0.0;
            lc[2 * q + 1] = (lane < 2 * q + 1 && 2 * q + 1 < ncols) ? v.y : 0.0;
        }
    };
    double lc[GT_NB];
    if (tid < 32) load_lc(nblk - 1, n - GT_NB * (nblk - 1), lc);
    for (int J = nblk - 1; J >= 0; --J) {
        const int ncols = min(GT_NB, n - GT_NB * J);
        GT_TICK(12);
        if (tid < 32) {
            // 16 x 16 upper-triangular solve inside warp 0: alpha_k = (z_k - sum_{r>k} L[r][k] alpha_r) / L[k][k].
            // Lane j keeps z_j and column j of the diagonal block in registers and alpha_k travels by one shuffle per
            // step: the same operations in the same order as a sweep through shared memory, without its two round
            // trips per step; straight-line code for a full block (a branch per step costs more than the arithmetic).
            const bool own = lane < ncols;
            double z = own ? zz[GT_NB * J + lane] : 0.0;
            double z2 = (own && sec.y) ? zz2[GT_NB * J + lane] : 0.0;
            const double dinv = own ? idg[GT_NB * J + lane] : 0.0;
            GT_TICK(13);
            if (ncols == GT_NB && sec.y) {
#pragma unroll
                for (int k = GT_NB - 1; k >= 0; --k) {
                    const double ak = __shfl_sync(0xffffffffu, z * dinv, k);
                    const double ak2 = __shfl_sync(0xffffffffu, z2 * dinv, k);
                    const double t = fma(-lc[k], ak, z), t2 = fma(-lc[k], ak2, z2);
                    z = (lane == k) ? ak : ((lane < k) ? t : z);
                    z2 = (lane == k) ? ak2 : ((lane < k) ? t2 : z2);
                }
            } else if (ncols == GT_NB) {
#pragma unroll
                for (int k = GT_NB - 1; k >= 0; --k) {
                    const double ak = __shfl_sync(0xffffffffu, z * dinv, k);
                    const double t = fma(-lc[k], ak, z);
                    z = (lane == k) ? ak : ((lane < k) ? t : z);
                }
            } else {
                for (int k = ncols - 1; k >= 0; --k) {
                    const double ak = __shfl_sync(0xffffffffu, z * dinv, k);
                    const double ak2 = sec.y ? __shfl_sync(0xffffffffu, z2 * dinv, k) : 0.0;
                    const double lk = lc_at(lc, k);
                    if (lane == k) { z = ak; z2 = ak2; }
                    else if (lane < k) { z = fma(-lk, ak, z); z2 = fma(-lk, ak2, z2); }
                }
            }
            GT_TICK(14);
            if (own) {
                zz[GT_NB * J + lane] = z;
                if (sec.y) zz2[GT_NB * J + lane] = z2;
            }
        }
        __syncthreads();
        GT_TICK(15);
        if (tid < 32 && J > 0) load_lc(J - 1, GT_NB, lc);
        // z[j] -= sum_{k in block J} L[k][j] alpha_k for j < 16 J (row 16 J + k of column j: consecutive k are
        // consecutive doubles, so a full block is 8 vector loads issued ahead of the 16 dependent FMAs)
        for (int j = tid; j < GT_NB * J; j += GT_THREADS) {
            double s = zz[j], s2 = sec.y ? zz2[j] : 0.0;
            const int Jc = j >> 4;
            const double* lrow = Lb + gt_off(R, Jc) + (j & 15) * (R - GT_NB * Jc) + GT_NB * (J - Jc);
            if (ncols == GT_NB) {
                double2 lv[GT_NB / 2];
#pragma unroll
                for (int q = 0; q < GT_NB / 2; ++q) lv[q] = reinterpret_cast<const double2*>(lrow)[q];
#pragma unroll
                for (int q = 0; q < GT_NB / 2; ++q) {
                    s = fma(-lv[q].x, zz[GT_NB * J + 2 * q], s);
                    s = fma(-lv[q].y, zz[GT_NB * J + 2 * q + 1], s);
                }
                if (sec.y) {
#pragma unroll
                    for (int q = 0; q < GT_NB / 2; ++q) {
                        s2 = fma(-lv[q].x, zz2[GT_NB * J + 2 * q], s2);
                        s2 = fma(-lv[q].y, zz2[GT_NB * J + 2 * q + 1], s2);
                    }
                }
            } else {
                for (int k = 0; k < ncols; ++k) {
                    const double lkj = lrow[k];
                    s = fma(-lkj, zz[GT_NB * J + k], s);
                    if (sec.y) s2 = fma(-lkj, zz2[GT_NB * J + k], s2);
                }
            }
            zz[j] = s;
            if (sec.y) zz2[j] = s2;
        }
        __syncthreads();
    }
    GT_TICK(9);
    if (metric == GT_METRIC_FIT) {
        // final fit (gpr_fitting.py:80-82): alpha and the dense lower factor (row-major) for the K_inv kernel
        for (int j = tid; j < n; j += GT_THREADS) fit_alpha[j] = zz[j];
        for (int e = tid; e < n * n; e += GT_THREADS) {
            const int i = e / n, j = e - i * n;
            fit_L[e] = (j <= i) ? Lat(i, j) : 0.0;
        }
        if (tid == 0) fit_info[0] = 0;
        return;
    }
    double se = 0.0, se2 = 0.0;
    const double* shape_val = shape ? shape + gt_off(R, nblk) : nullptr;
    for (int v = tid; v < nv; v += GT_THREADS) {
        const double xv = Xv[v];
        double pred = 0.0, pred2 = 0.0;
        if (shape_val) {
            // covariance entries from the shape table (coalesced over the validation points), same dot product
            const double* kcol = shape_val + v;
            if (sec.y) {
#pragma unroll GT_TAB_UNROLL
                for (int j = 0; j < n; ++j) {
                    const double kv = B200ID_MUL(kp.p1, kcol[(int64_t)j * nv]);
                    pred = fma(kv, zz[j], pred);
                    pred2 = fma(kv, zz2[j], pred2);
                }
            } else {
#pragma unroll GT_TAB_UNROLL
                for (int j = 0; j < n; ++j) pred = fma(B200ID_MUL(kp.p1, kcol[(int64_t)j * nv]), zz[j], pred);
            }
        } else if (sec.y) {
            B200ID_DISPATCH_KERNEL(kid, gt_predict_point2, kp, xv, xs, zz, zz2, n, &pred, &pred2)
        } else {
            B200ID_DISPATCH_KERNEL(kid, gt_predict_point, kp, xv, xs, zz, n, &pred)
        }
        if (sec.y) {
            pred2 = pred2 * sec.y_scale + sec.y_mean;
            const double e2 = sec.yv[v] - pred2;
            se2 = fma(e2, e2, se2);
        }
        pred = pred * y_scale + y_mean;
        const double e = yv[v] - pred;
        se = fma(e, e, se);
    }
    GT_TICK(10);
    const double tot = block_sum(se, scratch);
    if (tid == 0) metric_out[cand] = sqrt(tot / (double)nv);
    if (sec.y) {
        __syncthreads();
        const double tot2 = block_sum(se2, scratch);
        if (tid == 0) sec.metric_out[cand] = sqrt(tot2 / (double)nv);
    }
}

}  // namespace b200id

using namespace b200id;

// Called by b200id_gp_batch_eval (gp_batch.cu) when n <= GT_NMAX.
int b200id_gp_batch_tiled_launch(int kernel_id, const int* kernel_ids, const double* theta, int64_t n_candidates,
                                 const double* X, const double* y, int n, double noise, int metric,
                                 const double* Xv, const double* yv, int nv, double y_mean, double y_scale,
                                 double* metric_out, const double* noise_v, cudaStream_t st,
                                 const double* y_b, const double* yv_b, double y_mean_b, double y_scale_b,
                                 double* metric_out_b, const double* shape_tab, const int* shape_of) {
    const GtSecond sec{y_b, yv_b, y_mean_b, y_scale_b, metric_out_b};
    const GtShapes shp{shape_tab, shape_of, gt_shape_stride(n, nv, y_b ? 2 : 1)};
    const size_t smem = gt_smem_bytes(n, y_b ? 2 : 1);
    int rc = check_cuda(cudaFuncSetAttribute(gp_batch_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(gp tiled)");
    if (rc) return rc;
    gp_batch_tiled_kernel<<<(unsigned)n_candidates, GT_THREADS, smem, st>>>(kernel_id, kernel_ids, theta, X, y, n, noise, metric,
                                                                           Xv, yv, nv, y_mean, y_scale, metric_out,
                                                                           nullptr, nullptr, nullptr, noise_v, sec, shp);
    B200ID_LAUNCH_CHECK("gp_batch_tiled_kernel");
    return B200ID_OK;
}

// Shape tables for b200id_gp_batch_eval_shared: n_shapes tables of gt_shape_stride(n, nv, nrhs) doubles each.
size_t b200id_gp_shape_table_doubles(int n_shapes, int n, int nv, int nrhs) {
    return (size_t)n_shapes * (size_t)gt_shape_stride(n, nv, nrhs);
}
int b200id_gp_shape_table_launch(int kernel_id, const double* shape_theta, int n_shapes, const double* X, int n,
                                 const double* Xv, int nv, int nrhs, double* tab, cudaStream_t st) {
    const dim3 grid((unsigned)(gt_nblk(n) + ((int64_t)n * nv + GT_TAB_CHUNK - 1) / GT_TAB_CHUNK), (unsigned)n_shapes);
    gp_shape_table_kernel<<<grid, GT_TAB_THREADS, 0, st>>>(kernel_id, shape_theta, X, n, Xv, nv, nrhs, tab,
                                                           gt_shape_stride(n, nv, nrhs));
    B200ID_LAUNCH_CHECK("gp_shape_table_kernel");
    return B200ID_OK;
}

#ifdef GT_PROFILE
extern "C" int b200id_debug_gt_profile(unsigned long long* out16_host, int reset) {
    if (reset) {
        unsigned long long z[16] = {0};
        return check_cuda(cudaMemcpyToSymbol(g_gt_prof, z, sizeof(z)), "reset prof");
    }
    return check_cuda(cudaMemcpyFromSymbol(out16_host, g_gt_prof, 16 * sizeof(unsigned long long)), "read prof");
}
#endif

// Final fits through the same blocked kernel, one CTA per problem: theta [B][3], y [B][n] -> alpha [B][n],
// dense L [B][n*n] (row-major lower) and info [B].
int b200id_gp_fit_tiled_launch(int kernel_id, const double* theta, const double* X, const double* y, int n,
                               double diag_add, double* alpha_out, double* L_out, int* info_out, cudaStream_t st,
                               int n_problems) {
    const size_t smem = gt_smem_bytes(n);
    int rc = check_cuda(cudaFuncSetAttribute(gp_batch_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(gp tiled)");
    if (rc) return rc;
    gp_batch_tiled_kernel<<<n_problems, GT_THREADS, smem, st>>>(kernel_id, nullptr, theta, X, y, n, diag_add, GT_METRIC_FIT,
                                                               nullptr, nullptr, 0, 0.0, 1.0, nullptr, alpha_out, L_out, info_out, nullptr, GtSecond{nullptr, nullptr, 0.0, 1.0, nullptr},
                                                               GtShapes{nullptr, nullptr, 0});
    B200ID_LAUNCH_CHECK("gp_batch_tiled_kernel(fit)");
    return B200ID_OK;
}
