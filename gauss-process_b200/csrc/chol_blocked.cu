// chol_blocked.cu -- K4: blocked FP64 Cholesky for large N_d (BASELINE config 4, N_d = 2048).
//
// Reference: np.linalg.cholesky at src/gpr/gpr_fitting.py:80 / grid_search.py:103 on a Gram matrix
// too large for one CTA's shared memory.  Right-looking blocked algorithm on the row-major lower
// triangle in global memory (32 MB at n = 2048, i.e. L2 resident on B200):
//   for each 64-wide block column k:
//     (1) diag : one CTA factorises the 64x64 diagonal block in shared memory (fused into step (3) of the block
//                column before, except for the first)
//     (2) panel: P L_kk^T = A[k+64:, k:k+64] by substitution, the row held in registers
//     (3) syrk : A[i,j] -= P_i P_j^T for tiles i >= j  -- 64x64x64 FP64 tile products, 256 threads, each a 4x4 block
//                of the output tile, operands staged k-major in shared memory (one 32-byte vector load per 4 rows /
//                4 columns, 16 DFMA per 8 loaded doubles)
// Extra rows (n_rows > n, the right-hand sides of the GP solve stored as rows n.. of the same array) ride along: the
// panel step turns them into z^T = (L^-1 y)^T block column by block column and the trailing update keeps them current,
// so the forward substitution of the solve costs no launch of its own (cholb_factor_rows; gp_large.cu).
// One factorisation is a chain of 2 n / 64 dependent launches, i.e. latency bound at n = 2048 (1.9 ms where the flops
// alone would take 0.1 ms); throughput comes from running several candidates side by side (gp_large.cu).
#include <stdint.h>
#include "common.cuh"
#include "gp_linalg.cuh"

namespace b200id {

constexpr int CB_NB = 64;            // block size
constexpr int CB_THREADS = 256;
constexpr int CB_LDT = CB_NB + 2;    // k-major tile stride (even: keeps double2 loads aligned)

// acc[a][b] += sum_k As[k][4*ty + a] * Bs[k][4*tx + b]
__device__ __forceinline__ void tile_mm(const double* __restrict__ As, const double* __restrict__ Bs, int kdim,
                                        int ty, int tx, double acc[4][4]) {
#pragma unroll 8
    for (int k = 0; k < kdim; ++k) {
        const double2 a01 = *reinterpret_cast<const double2*>(As + k * CB_LDT + 4 * ty);
        const double2 a23 = *reinterpret_cast<const double2*>(As + k * CB_LDT + 4 * ty + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(Bs + k * CB_LDT + 4 * tx);
        const double2 b23 = *reinterpret_cast<const double2*>(Bs + k * CB_LDT + 4 * tx + 2);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
}

// Stage rows [r0, r0+64) x cols [c0, c0+64) of row-major M (ld) into k-major T[k][r] (k = column).
// Rows >= r_end are zero filled.
__device__ __forceinline__ void stage_tile_kmajor(const double* __restrict__ M, int ld, int r0, int r_end, int c0,
                                                  double* __restrict__ T) {
    // all 16 global loads of a thread are issued before the first shared-memory store, so their latencies overlap
    constexpr int PER = CB_NB * CB_NB / CB_THREADS;
    double v[PER];
#pragma unroll
    for (int it = 0; it < PER; ++it) {
        const int e = threadIdx.x + it * CB_THREADS;
        const int r = e >> 6, k = e & 63;
        v[it] = (r0 + r < r_end) ? __ldg(M + (size_t)(r0 + r) * ld + c0 + k) : 0.0;
    }
#pragma unroll
    for (int it = 0; it < PER; ++it) {
        const int e = threadIdx.x + it * CB_THREADS;
        const int r = e >> 6, k = e & 63;
        T[k * CB_LDT + r] = v[it];
    }
}

// (1) diagonal block: factorise in shared memory, write L back.
// This routine is the serial link of the factorisation (one CTA, n / 64 times), so it is written for latency.
constexpr int CB_LDD = CB_NB + 1;
// Factorises the nb x nb block held in D (lower triangle, column-major D[j * CB_LDD + i]) in place and writes L to
// A[k0 + i][k0 + j].  All CB_THREADS threads call; D must be complete (caller synchronised); dg needs 17 doubles.
// Returns false on a failed pivot.
__device__ __forceinline__ bool cholb_factor_block(double* __restrict__ D, double* __restrict__ dg, int nb,
                                                   double* __restrict__ A, int lda, int k0, int* __restrict__ info) {
    // Two-level: four 16-wide block columns.  Per block column
    //   (i)   warp 0 factorises the 16 x 16 diagonal sub-block in REGISTERS (lane = row, columns broadcast by shuffle:
    //         the only serial latency per pivot is one rsqrt + Newton pair, no barrier),
    //   (ii)  one thread per row below solves x L^T = a by forward substitution in registers (reciprocal diagonal),
    //   (iii) all threads apply the rank-16 update to what is left.
    // 12 block barriers per 64 x 64 block instead of 64 (this routine is the serial link of the whole factorisation:
    // 39 us -> measured in profiles/r02_chol_blocked.txt).  dg[0..15] holds the reciprocal pivots of the current
    // sub-block, dg[16] the failure flag.
    constexpr int ld = CB_LDD;
    constexpr int SB = 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) dg[SB] = 0.0;
    __syncthreads();
    for (int c0 = 0; c0 < nb; c0 += SB) {
        const int w = min(SB, nb - c0);
        if (warp == 0) {
            double d[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) d[c] = (lane < w && c <= lane) ? D[(c0 + c) * ld + c0 + lane] : 0.0;
            int fail = 0;
#pragma unroll
            for (int k = 0; k < SB; ++k) {
                if (k < w && fail == 0) {
                    const double piv = __shfl_sync(0xffffffffu, d[k], k);
                    if (piv <= 0.0 || piv == INFINITY) {           // NaN compares false: the factorisation goes on
                        fail = k + 1;
                    } else {
                        double inv = rsqrt(piv);
                        double dk = piv * inv;
                        dk = fma(0.5 * inv, fma(-dk, dk, piv), dk);
                        inv = fma(inv, fma(-dk, inv, 1.0), inv);
                        const double l = (lane == k) ? dk : d[k] * inv;
                        d[k] = l;
                        if (lane == k) dg[k] = inv;
#pragma unroll
                        for (int c = k + 1; c < SB; ++c) {
                            const double lc = __shfl_sync(0xffffffffu, l, c);
                            d[c] = fma(-l, lc, d[c]);
                        }
                    }
                }
            }
            if (fail != 0) {
                if (lane == 0) { dg[SB] = 1.0; info[0] = k0 + c0 + fail; }
            } else {
#pragma unroll
                for (int c = 0; c < SB; ++c)
                    if (lane < w && c <= lane) D[(c0 + c) * ld + c0 + lane] = d[c];
            }
        }
        __syncthreads();
        if (dg[SB] != 0.0) return false;
        const int below = nb - (c0 + w);
        if (below > 0) {
            if (tid < below) {
                const int i = c0 + w + tid;
                double x[SB];
#pragma unroll
                for (int c = 0; c < SB; ++c) {
                    if (c < w) {
                        double v = D[(c0 + c) * ld + i];
#pragma unroll
                        for (int k = 0; k < c; ++k) v = fma(-x[k], D[(c0 + k) * ld + c0 + c], v);
                        x[c] = v * dg[c];
                    } else {
                        x[c] = 0.0;
                    }
                }
#pragma unroll
                for (int c = 0; c < SB; ++c)
                    if (c < w) D[(c0 + c) * ld + i] = x[c];
            }
            __syncthreads();
            // rank-w update of the lower triangle of rows / columns [c0 + w, nb): thread (ty, tx) of a 16 x 16 grid
            const int ty = tid >> 4, tx = tid & 15, base = c0 + w;
            for (int i = base + ty; i < nb; i += 16)
                for (int j = base + tx; j <= i; j += 16) {
                    double v = D[j * ld + i];
#pragma unroll
                    for (int k = 0; k < SB; ++k)
                        if (k < w) v = fma(-D[(c0 + k) * ld + i], D[(c0 + k) * ld + j], v);
                    D[j * ld + i] = v;
                }
            __syncthreads();
        }
    }
    for (int e = tid; e < CB_NB * CB_NB; e += CB_THREADS) {
        const int i = e >> 6, j = e & 63;
        if (i < nb && j <= i) A[(size_t)(k0 + i) * lda + k0 + j] = D[j * ld + i];
    }
    return true;
}

__global__ void __launch_bounds__(CB_THREADS)
cholb_diag_kernel(double* __restrict__ A, int lda, int n, int k0, int* __restrict__ info, size_t a_stride, size_t info_stride) {
    __shared__ double D[CB_NB * CB_LDD];
    __shared__ double dg[CB_NB];
    A += blockIdx.z * a_stride;                      // blockIdx.z = matrix of a batch advancing in lockstep
    info += blockIdx.z * info_stride;
    if (info[0] != 0) return;                       // an earlier block already failed
    const int nb = min(CB_NB, n - k0);
    const int tid = threadIdx.x;
    {
        constexpr int PER = CB_NB * CB_NB / CB_THREADS;
        double v[PER];
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            const int e = tid + it * CB_THREADS;
            const int i = e >> 6, j = e & 63;
            v[it] = (i < nb && j <= i) ? A[(size_t)(k0 + i) * lda + k0 + j] : 0.0;
        }
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            const int e = tid + it * CB_THREADS;
            const int i = e >> 6, j = e & 63;
            if (i < nb && j <= i) D[j * CB_LDD + i] = v[it];
        }
    }
    __syncthreads();
    cholb_factor_block(D, dg, nb, A, lda, k0, info);
}

// (2) panel: rows [k0+64, n) of block column k0:  P L_kk^T = A, by forward substitution over the 64 columns (the
// backward-stable TRSM; a product with an explicit L_kk^-1 loses cond(L_kk) eps).  A CTA owns 64 rows; four threads
// share a row and split each dot product, so a column step is <= 16 FMAs, two shuffles and a __syncwarp -- the rows
// of a warp never touch each other's data, no block barrier inside the loop.
constexpr int CB_LDP = CB_NB + 1;
__global__ void __launch_bounds__(CB_THREADS)
cholb_panel_kernel(double* __restrict__ A, int lda, int n, int n_rows, int k0, const int* __restrict__ info,
                   size_t a_stride, size_t info_stride) {
    extern __shared__ __align__(16) double cb_sm[];
    double* Ps = cb_sm;                              // [64][65] rows of the panel tile
    double* Ls = cb_sm + CB_NB * CB_LDP;             // [64][65] L_kk, Ls[c][p] = L[c][p], p <= c
    __shared__ double idg[CB_NB];
    A += blockIdx.z * a_stride;
    info += blockIdx.z * info_stride;
    if (info[0] != 0) return;
    const int tid = threadIdx.x;
    // nb < 64 only for the last block column, which has a panel only when extra rows ride along: L_kk is then
    // completed by an identity block and the columns past n are neither read nor written
    const int nb = min(CB_NB, n - k0);
    const int r0 = k0 + nb + blockIdx.x * CB_NB;
    {
        constexpr int PER = CB_NB * CB_NB / CB_THREADS;
        double vp[PER], vl[PER];
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            const int e = tid + it * CB_THREADS;
            const int r = e >> 6, c = e & 63;
            vp[it] = (r0 + r < n_rows && c < nb) ? A[(size_t)(r0 + r) * lda + k0 + c] : 0.0;
            vl[it] = (r < nb) ? ((c <= r) ? A[(size_t)(k0 + r) * lda + k0 + c] : 0.0) : (c == r ? 1.0 : 0.0);
        }
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            const int e = tid + it * CB_THREADS;
            const int r = e >> 6, c = e & 63;
            Ps[r * CB_LDP + c] = vp[it];
            Ls[r * CB_LDP + c] = vl[it];
        }
    }
    __syncthreads();
    if (tid < CB_NB) idg[tid] = 1.0 / Ls[tid * CB_LDP + tid];
    __syncthreads();
    // thread (r, q) keeps P[r][q], P[r][q + 4], ... in registers (column loop fully unrolled, so the indices are
    // static); L_kk rows are read-only broadcast loads that never wait for a store
    const int r = tid >> 2, q = tid & 3;
    double pr[CB_NB / 4];
#pragma unroll
    for (int m = 0; m < CB_NB / 4; ++m) pr[m] = Ps[r * CB_LDP + 4 * m + q];
#pragma unroll
    for (int c = 0; c < CB_NB; ++c) {
        const double* lrow = Ls + c * CB_LDP + q;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int m = 0; m < CB_NB / 4; ++m) {
            if (4 * m < c) {                         // compile-time bound on m; the lane's own column p = 4 m + q must be < c
                const double lv = (4 * m + q < c) ? lrow[4 * m] : 0.0;
                if (m & 1) s1 = fma(pr[m], lv, s1);
                else s0 = fma(pr[m], lv, s0);
            }
        }
        double sacc = s0 + s1;
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        if (q == (c & 3)) pr[c >> 2] = (pr[c >> 2] - sacc) * idg[c];
    }
#pragma unroll
    for (int m = 0; m < CB_NB / 4; ++m) Ps[r * CB_LDP + 4 * m + q] = pr[m];
    __syncthreads();
    for (int e = tid; e < CB_NB * CB_NB; e += CB_THREADS) {
        const int rr = e >> 6, c = e & 63;
        if (r0 + rr < n_rows && c < nb) A[(size_t)(r0 + rr) * lda + k0 + c] = Ps[rr * CB_LDP + c];
    }
}

// (3) trailing update over lower-triangle tiles (ti >= tj) of the matrix starting at s = k0 + 64
__global__ void __launch_bounds__(CB_THREADS)
cholb_syrk_kernel(double* __restrict__ A, int lda, int n, int n_rows, int k0, int* __restrict__ info,
                  size_t a_stride, size_t info_stride) {
    extern __shared__ __align__(16) double cb_sm[];
    double* As = cb_sm;
    double* Bs = cb_sm + CB_NB * CB_LDT;
    A += blockIdx.z * a_stride;
    info += blockIdx.z * info_stride;
    if (info[0] != 0) return;
    // linear tile index -> (ti, tj), ti >= tj
    const int lin = blockIdx.x;
    int ti = (int)((sqrt(8.0 * lin + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= lin) ++ti;
    while (ti * (ti + 1) / 2 > lin) --ti;
    const int tj = lin - ti * (ti + 1) / 2;
    const int s = k0 + CB_NB;
    const int r0 = s + ti * CB_NB, c0 = s + tj * CB_NB;
    stage_tile_kmajor(A, lda, r0, n_rows, k0, As);
    stage_tile_kmajor(A, lda, c0, n, k0, Bs);
    __syncthreads();
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    double acc[4][4] = {};
    tile_mm(As, Bs, CB_NB, ty, tx, acc);
    double cv[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + 4 * ty + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * tx + j;
            cv[i][j] = (r < n_rows && c <= r && c < n) ? A[(size_t)r * lda + c] : 0.0;
        }
    }
    if (lin == 0) {
        // tile (0, 0) of the trailing matrix is the NEXT diagonal block: factorise it here instead of in a launch of
        // its own (the other CTAs of this launch do not depend on it, the next launch sees it finished) -- one link
        // less in the chain of dependent launches per block column
        __syncthreads();                             // everybody is done with the staged operands
        double* D = cb_sm;                           // reuses the operand tiles: 64 x 65 doubles + 64
        double* dg = cb_sm + CB_NB * CB_LDD;
        const int nb = min(CB_NB, n - s);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ii = 4 * ty + i, jj = 4 * tx + j;
                if (ii < nb && jj <= ii) D[jj * CB_LDD + ii] = cv[i][j] - acc[i][j];
            }
        }
        // rows of this tile below the diagonal block exist only as extra rows (nb < 64): plain update
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + 4 * ty + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + 4 * tx + j;
                if (4 * ty + i >= nb && r < n_rows && c < n) A[(size_t)r * lda + c] = cv[i][j] - acc[i][j];
            }
        }
        __syncthreads();
        cholb_factor_block(D, dg, nb, A, lda, s, info);
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + 4 * ty + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * tx + j;
            if (r < n_rows && c <= r && c < n) A[(size_t)r * lda + c] = cv[i][j] - acc[i][j];
        }
    }
}

// (3') the same trailing update on the FP64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA.8x8x4 in SASS), one warp
// instruction = 256 FMAs.  FP64 MMA and FP64 FMA share one pipe on B200 (same 37 TFLOP/s peak), so
// the gain is not a higher ceiling but reaching it: a DFMA reading three distinct registers holds the pipe for 3 cycles
// instead of 2 and needs one issue slot per 32 FMAs, the register-tiled loop above runs at ~36 % of the peak; here the
// 64 x 64 x 64 tile product is 128 DMMAs per warp fed by 6 conflict-free LDS.64 per 8 of them.
// Operands are staged ROW-major (Ts[r][k], pitch 68 doubles): no transposition on the way in (coalesced global rows ->
// consecutive shared-memory words), and the fragment of lane l -- row l / 4, k-column l % 4, for the A operand and for
// the B operand alike, since the update is P_r P_c^T -- touches every bank exactly twice per 64-bit warp load.
// Warp w owns rows [16 (w / 2), +16) x columns [32 (w % 2), +32) of the tile = 2 x 4 accumulator fragments; lane l
// holds C[l / 4][2 (l % 4) .. +1] of each 8 x 8 fragment.  The C tile is fetched BEFORE the product so that its latency
// hides under the DMMAs; that needs 128 registers, i.e. 2 CTAs per SM -- measured (16 candidates at n = 2048): late C,
// 3 CTAs/SM 5.22 ms; early C, 3 CTAs/SM (spills) 5.77 ms; late C, 2 CTAs/SM 5.0 ms; early C, 2 CTAs/SM 4.7-4.9 ms.
constexpr int CB_LDR = CB_NB + 4;
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void stage_tile_rowmajor(const double* __restrict__ M, int ld, int r0, int r_end, int c0,
                                                    double* __restrict__ T) {
    constexpr int PER = CB_NB * CB_NB / CB_THREADS;
    double v[PER];
#pragma unroll
    for (int it = 0; it < PER; ++it) {
        const int e = threadIdx.x + it * CB_THREADS;
        const int r = e >> 6, k = e & 63;
        v[it] = (r0 + r < r_end) ? __ldg(M + (size_t)(r0 + r) * ld + c0 + k) : 0.0;
    }
#pragma unroll
    for (int it = 0; it < PER; ++it) {
        const int e = threadIdx.x + it * CB_THREADS;
        T[(e >> 6) * CB_LDR + (e & 63)] = v[it];
    }
}

#ifndef CB_SYRK_CTAS
#define CB_SYRK_CTAS 2
#endif
#ifndef CB_CPASYNC
#define CB_CPASYNC 1            // 1: operands go global -> shared with cp.async in two k-halves (no staging registers,
#endif                          //    the second half lands while the first is multiplied)
// k-columns [kh, kh + 32) of rows [r0, r0 + 64) x cols [c0 + kh, ...) of row-major M into Ts[r][k]; rows >= r_end are
// zero filled (cp.async with source size 0)
__device__ __forceinline__ void stage_half_async(const double* __restrict__ M, int ld, int r0, int r_end, int c0, int kh,
                                                 double* __restrict__ T) {
#pragma unroll
    for (int it = 0; it < CB_NB * 32 / CB_THREADS; ++it) {
        const int e = threadIdx.x + it * CB_THREADS;
        const int r = e >> 5, k = kh + (e & 31);
        const bool ok = r0 + r < r_end;
        const double* src = M + (size_t)(ok ? r0 + r : r0) * ld + c0 + k;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(T + r * CB_LDR + k);
        const int sz = ok ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    }
}
#ifndef CB_EARLY_C
#define CB_EARLY_C 1            // 1: the C tile is loaded before the product instead of after it (more live registers)
#endif
__global__ void __launch_bounds__(CB_THREADS, CB_SYRK_CTAS)
cholb_syrk_mma_kernel(double* __restrict__ A, int lda, int n, int n_rows, int k0, int* __restrict__ info,
                      size_t a_stride, size_t info_stride, int first_column_only) {
    extern __shared__ __align__(16) double cb_sm[];
    double* As = cb_sm;
    double* Bs = cb_sm + CB_NB * CB_LDR;
    A += blockIdx.z * a_stride;
    info += blockIdx.z * info_stride;
    if (info[0] != 0) return;
    const int lin = blockIdx.x;
    // first_column_only: the grid is one CTA per tile ROW and only the first block column of the trailing matrix is
    // updated (the second half of a 128-wide super-panel; the rest waits for the rank-128 update below)
    int ti = first_column_only ? lin : (int)((sqrt(8.0 * lin + 1.0) - 1.0) * 0.5);
    if (!first_column_only) {
        while ((ti + 1) * (ti + 2) / 2 <= lin) ++ti;
        while (ti * (ti + 1) / 2 > lin) --ti;
    }
    const int tj = first_column_only ? 0 : lin - ti * (ti + 1) / 2;
    const int s = k0 + CB_NB;
    const int r0 = s + ti * CB_NB, c0 = s + tj * CB_NB;
#if CB_CPASYNC
    stage_half_async(A, lda, r0, n_rows, k0, 0, As);
    stage_half_async(A, lda, c0, n, k0, 0, Bs);
    asm volatile("cp.async.commit_group;" ::: "memory");
    stage_half_async(A, lda, r0, n_rows, k0, 32, As);
    stage_half_async(A, lda, c0, n, k0, 32, Bs);
    asm volatile("cp.async.commit_group;" ::: "memory");
#else
    stage_tile_rowmajor(A, lda, r0, n_rows, k0, As);
    stage_tile_rowmajor(A, lda, c0, n, k0, Bs);
    __syncthreads();
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = warp >> 1, wx = warp & 1, fr = lane >> 2, fk = lane & 3;
    // this lane's elements: tile row 16 wy + 8 i + fr, tile columns 32 wx + 8 j + 2 fk + {0, 1}
    double cv[2][4][2];
    auto load_c = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = r0 + 16 * wy + 8 * i + fr;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = c0 + 32 * wx + 8 * j + 2 * fk + q;
                    cv[i][j][q] = (r < n_rows && c <= r && c < n) ? A[(size_t)r * lda + c] : 0.0;
                }
        }
    };
    if (CB_EARLY_C) load_c();
    double acc[2][4][2] = {};
    const double* pa = As + (16 * wy + fr) * CB_LDR + fk;
    const double* pb = Bs + (32 * wx + fr) * CB_LDR + fk;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#if CB_CPASYNC
        if (half == 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#endif
#pragma unroll 4
        for (int kk = 32 * half; kk < 32 * half + 32; kk += 4) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = pa[8 * i * CB_LDR + kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = pb[8 * j * CB_LDR + kk];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    if (!CB_EARLY_C) load_c();
    if (lin == 0) {
        // tile (0, 0) of the trailing matrix is the NEXT diagonal block: factorise it here (see cholb_syrk_kernel)
        __syncthreads();                             // everybody is done with the staged operands
        double* D = cb_sm;
        double* dg = cb_sm + CB_NB * CB_LDD;
        const int nb = min(CB_NB, n - s);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int ii = 16 * wy + 8 * i + fr;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int jj = 32 * wx + 8 * j + 2 * fk + q;
                    const double v = cv[i][j][q] - acc[i][j][q];
                    if (ii < nb && jj <= ii) D[jj * CB_LDD + ii] = v;
                    // rows of this tile below the diagonal block exist only as extra rows (nb < 64): plain update
                    else if (ii >= nb && r0 + ii < n_rows && c0 + jj < n) A[(size_t)(r0 + ii) * lda + c0 + jj] = v;
                }
        }
        __syncthreads();
        cholb_factor_block(D, dg, nb, A, lda, s, info);
        return;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = r0 + 16 * wy + 8 * i + fr;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int c = c0 + 32 * wx + 8 * j + 2 * fk + q;
                if (r < n_rows && c <= r && c < n) A[(size_t)r * lda + c] = cv[i][j][q] - acc[i][j][q];
            }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// (3') rank-128 trailing update over 128 x 64 tiles: A[i, j] -= P_i P_j^T with P = the 128-wide super-panel
// [k0, k0 + 128), tiles of the matrix starting at s = k0 + 128 that touch its lower triangle.  Twice the panel width
// and twice the tile area of cholb_syrk_mma_kernel: 6.4 flop per byte moved instead of 4, half as many dependent
// update launches.  Operands arrive by TMA bulk copies (cp.async.bulk, one 256-byte row segment each, completion
// counted on an mbarrier) in four k-chunks of 32 through a two-stage ring: the copy of chunk c + 2 is issued as soon as
// chunk c has been multiplied and lands under the DMMAs of chunk c + 1; no thread spends registers or issue slots on
// staging.  8 warps, warp (wy, wx) owns rows [32 wy, +32) x columns [32 wx, +32) of the tile = 4 x 4 accumulator
// fragments: 8 conflict-free LDS.64 feed 16 DMMAs.  Two CTAs per SM, so one CTA's prologue and read-modify-write of C
// run under the other's products.
constexpr int CW_TM = 128;                // tile rows and panel width
constexpr int CW_TN = 64;                 // tile columns
constexpr int CW_KC = 32;                 // k-chunk
constexpr int CW_NCHUNK = CW_TM / CW_KC;  // 4
constexpr int CW_STAGES = 2;
constexpr int CW_THREADS = 256;
constexpr int CW_LD = CW_KC + 4;          // row pitch of a staged chunk (doubles): 288 bytes, fragments conflict-free
constexpr size_t CW_STAGE_DOUBLES = (size_t)(CW_TM + CW_TN) * CW_LD;
constexpr size_t CW_SMEM_BYTES = CW_STAGES * CW_STAGE_DOUBLES * sizeof(double) + 64;

__device__ __forceinline__ unsigned cw_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cw_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cw_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cw_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cw_mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
// tiles of a trailing matrix of t128 row tiles: row tile ti has column tiles tj = 0 .. 2 ti + 1
__host__ __device__ inline int cw_tile_count(int t128) { return t128 * (t128 + 1); }

__global__ void __launch_bounds__(CW_THREADS, 2)
cholb_syrk_wide_kernel(double* __restrict__ A, int lda, int n, int n_rows, int k0, int* __restrict__ info,
                       size_t a_stride, size_t info_stride) {
    extern __shared__ __align__(16) double cb_sm[];
    A += blockIdx.z * a_stride;
    info += blockIdx.z * info_stride;
    if (info[0] != 0) return;
    const int lin = blockIdx.x;
    int ti = (int)((sqrt(4.0 * lin + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) <= lin) ++ti;
    while (ti * (ti + 1) > lin) --ti;
    const int tj = lin - ti * (ti + 1);
    const int s = k0 + CW_TM;
    const int r0 = s + ti * CW_TM, c0 = s + tj * CW_TN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int validA = max(0, min(CW_TM, n_rows - r0)), validB = max(0, min(CW_TN, n - c0));
    if (validB == 0) return;                                      // tile entirely right of the matrix
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(cb_sm + CW_STAGES * CW_STAGE_DOUBLES);

    // rows past the matrix are never copied: they stay zero in every stage
    for (int e = tid; e < CW_STAGES * (CW_TM + CW_TN); e += CW_THREADS) {
        const int st = e / (CW_TM + CW_TN), row = e - st * (CW_TM + CW_TN);
        const bool is_b = row >= CW_TM;
        const int rr = is_b ? row - CW_TM : row;
        if (rr >= (is_b ? validB : validA)) {
            double* dst = cb_sm + st * CW_STAGE_DOUBLES + (size_t)row * CW_LD;
            for (int k = 0; k < CW_LD; ++k) dst[k] = 0.0;
        }
    }
    if (tid == 0) {
        for (int st = 0; st < CW_STAGES; ++st) cw_mbar_init(cw_smem_addr(bars + st), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    // chunk c -> stage c % 2: A rows r0.. and B rows c0.., columns [k0 + 32 c, +32).  Every warp issues the copies of
    // its own 24 rows (a bulk copy takes its addresses from uniform registers, so a warp issues its lanes' copies one
    // after the other: spread over the 8 warps that is 24 short issues each instead of 192 on one warp that also has
    // products to do); thread 0 arms the barrier with the byte count -- the transaction count may run negative until
    // then, the phase cannot complete before thread 0's arrival.
    auto issue = [&](int c) {
        const int st = c % CW_STAGES;
        const unsigned bar = cw_smem_addr(bars + st);
        double* base = cb_sm + st * CW_STAGE_DOUBLES;
        if (tid == 0) cw_mbar_expect_tx(bar, (unsigned)((validA + validB) * CW_KC * sizeof(double)));
        constexpr int ROWS_PER_WARP = (CW_TM + CW_TN) / (CW_THREADS / 32);
        if (lane < ROWS_PER_WARP) {
            const int row = warp * ROWS_PER_WARP + lane;
            const bool is_b = row >= CW_TM;
            const int rr = is_b ? row - CW_TM : row;
            if (rr < (is_b ? validB : validA)) {
                const double* src = A + (size_t)((is_b ? c0 : r0) + rr) * lda + k0 + CW_KC * c;
                cw_bulk_g2s(cw_smem_addr(base + (size_t)row * CW_LD), src, CW_KC * sizeof(double), bar);
            }
        }
    };
    issue(0); issue(1);
    const int wy = warp >> 1, wx = warp & 1, fr = lane >> 2, fk = lane & 3;
    double acc[4][4][2] = {};
#pragma unroll 1
    for (int c = 0; c < CW_NCHUNK; ++c) {
        const int st = c % CW_STAGES;
        cw_mbar_wait(cw_smem_addr(bars + st), (unsigned)((c / CW_STAGES) & 1));
        const double* pa = cb_sm + st * CW_STAGE_DOUBLES + (size_t)(32 * wy + fr) * CW_LD + fk;
        const double* pb = cb_sm + st * CW_STAGE_DOUBLES + (size_t)(CW_TM + 32 * wx + fr) * CW_LD + fk;
#pragma unroll 2
        for (int kk = 0; kk < CW_KC; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = pa[8 * i * CW_LD + kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = pb[8 * j * CW_LD + kk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (c + CW_STAGES < CW_NCHUNK) {
            // this stage is free once every warp has finished the chunk: refill it with chunk c + 2
            __syncthreads();
            issue(c + CW_STAGES);
        }
    }
    if (lin == 0) {
        // tile (0, 0) holds the NEXT diagonal block (its top 64 rows): update it with plain loads and factorise it
        // here, so the serial chain of the factorisation has no launch of its own for it
        __syncthreads();                                         // every warp is done with the staged operands
        double* D = cb_sm;
        double* dg = cb_sm + CB_NB * CB_LDD;
        const int nb = min(CB_NB, n - s);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ii = 32 * wy + 8 * i + fr, r = r0 + ii;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int jj = 32 * wx + 8 * j + 2 * fk + q, c = c0 + jj;
                    if (r < n_rows && c <= r && c < n) {
                        const double v = A[(size_t)r * lda + c] - acc[i][j][q];
                        if (ii < nb) D[jj * CB_LDD + ii] = v;
                        else A[(size_t)r * lda + c] = v;
                    }
                }
        }
        __syncthreads();
        cholb_factor_block(D, dg, nb, A, lda, s, const_cast<int*>(info));
        return;
    }
    // C -= acc on the lower triangle (c <= r), inside the matrix.  Every element has exactly one contribution per
    // launch, so the subtraction is handed to the L2 as a reduction (RED.ADD.F64: same IEEE addition, same result):
    // nothing waits for a load of C, and C never occupies a register.
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + 32 * wy + 8 * i + fr;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int c = c0 + 32 * wx + 8 * j + 2 * fk + q;
                if (r < n_rows && c <= r && c < n) atomicAdd(A + (size_t)r * lda + c, -acc[i][j][q]);
            }
    }
}

}  // namespace b200id

using namespace b200id;

extern "C" size_t b200id_chol_blocked_workspace_bytes(int n) {
    (void)n;
    return align_up((size_t)CB_NB * CB_NB * sizeof(double), 256);
}

// Factorise the n x n lower triangle of row-major A (ld = lda) and carry rows [n, n_rows) along (see the header).
// batch > 1: `batch` matrices a_stride doubles apart (status words info_stride ints apart) advance in LOCKSTEP, one
// launch per step for all of them (blockIdx.z): the chain of dependent launches is paid once per batch, not once per
// matrix, and every launch brings batch times the CTAs.
int cholb_factor_rows(double* A, int n, int n_rows, int lda, int* info_out, cudaStream_t st, int batch, size_t a_stride,
                      size_t info_stride) {
    int rc = B200ID_OK;
    for (int z = 0; z < batch; ++z) {
        rc = check_cuda(cudaMemsetAsync(info_out + (size_t)z * info_stride, 0, sizeof(int), st), "cudaMemsetAsync(info)");
        if (rc) return rc;
    }
    const size_t smem_panel = 2 * (size_t)CB_NB * CB_LDP * sizeof(double);
    static int use_mma = -1;                     // B200ID_CHOL_MMA=0: the register-tiled DFMA update (timing comparisons)
    if (use_mma < 0) { const char* e = getenv("B200ID_CHOL_MMA"); use_mma = (e && e[0] == '0') ? 0 : 1; }
    const size_t smem = 2 * (size_t)CB_NB * (use_mma ? CB_LDR : CB_LDT) * sizeof(double);
    rc = check_cuda(cudaFuncSetAttribute(cholb_syrk_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(syrk mma)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(cholb_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panel), "cudaFuncSetAttribute(panel)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(cholb_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(syrk)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(cholb_syrk_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM_BYTES), "cudaFuncSetAttribute(syrk wide)");
    if (rc) return rc;
    static int use_wide = -1;                    // B200ID_CHOL_WIDE=0: 64-wide updates only (round-1 schedule, timing comparisons)
    if (use_wide < 0) { const char* e = getenv("B200ID_CHOL_WIDE"); use_wide = (e && e[0] == '0') ? 0 : 1; }
    // the bulk copies and the paired stores of the wide update need 16-byte aligned row segments
    const bool aligned = (lda % 2 == 0) && (a_stride % 2 == 0) && (((uintptr_t)A & 15) == 0);
    const bool wide = use_wide && use_mma && aligned;
    // the first diagonal block has a launch of its own; every later one is factorised by the update kernel of the
    // block column before it (CTA of tile (0, 0))
    cholb_diag_kernel<<<dim3(1, 1, batch), CB_THREADS, 0, st>>>(A, lda, n, 0, info_out, a_stride, info_stride);
    if (wide) {
        // super-panels of 128 columns: panel | update of the second block column (its tile (0, 0) factorises the next
        // diagonal block) | panel | rank-128 update of everything to the right (tile (0, 0) factorises the next one)
        for (int k0 = 0; k0 < n; k0 += CW_TM) {
            for (int h = 0; h < 2; ++h) {
                const int kk = k0 + h * CB_NB;
                if (kk >= n) break;
                const int nb = n - kk < CB_NB ? n - kk : CB_NB;
                const int below = n_rows - (kk + nb);
                if (below <= 0) break;
                cholb_panel_kernel<<<dim3((below + CB_NB - 1) / CB_NB, 1, batch), CB_THREADS, smem_panel, st>>>(A, lda, n, n_rows, kk, info_out, a_stride, info_stride);
                if (h == 0 && n - (kk + CB_NB) > 0) {
                    const int tiles = (n_rows - (kk + CB_NB) + CB_NB - 1) / CB_NB;
                    cholb_syrk_mma_kernel<<<dim3(tiles, 1, batch), CB_THREADS, smem, st>>>(A, lda, n, n_rows, kk, info_out, a_stride, info_stride, 1);
                }
            }
            if (n - (k0 + CW_TM) > 0) {
                const int t128 = (n_rows - (k0 + CW_TM) + CW_TM - 1) / CW_TM;
                cholb_syrk_wide_kernel<<<dim3(cw_tile_count(t128), 1, batch), CW_THREADS, CW_SMEM_BYTES, st>>>(A, lda, n, n_rows, k0, info_out, a_stride, info_stride);
            }
        }
        B200ID_LAUNCH_CHECK("chol_blocked kernels");
        return B200ID_OK;
    }
    for (int k0 = 0; k0 < n; k0 += CB_NB) {
        const int nb = n - k0 < CB_NB ? n - k0 : CB_NB;
        const int below = n_rows - (k0 + nb);                // rows under this diagonal block (incl. extra rows)
        if (below <= 0) break;
        cholb_panel_kernel<<<dim3((below + CB_NB - 1) / CB_NB, 1, batch), CB_THREADS, smem_panel, st>>>(A, lda, n, n_rows, k0, info_out, a_stride, info_stride);
        if (n - (k0 + CB_NB) > 0) {                           // trailing columns left
            const int tiles = (n_rows - (k0 + CB_NB) + CB_NB - 1) / CB_NB;
            if (use_mma)
                cholb_syrk_mma_kernel<<<dim3(tiles * (tiles + 1) / 2, 1, batch), CB_THREADS, smem, st>>>(A, lda, n, n_rows, k0, info_out, a_stride, info_stride, 0);
            else
                cholb_syrk_kernel<<<dim3(tiles * (tiles + 1) / 2, 1, batch), CB_THREADS, smem, st>>>(A, lda, n, n_rows, k0, info_out, a_stride, info_stride);
        }
    }
    B200ID_LAUNCH_CHECK("chol_blocked kernels");
    return B200ID_OK;
}

extern "C" int b200id_chol_blocked(double* A, int n, int lda, int* info_out, void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(A && info_out && ws, B200ID_E_ARG, "chol_blocked: null pointer");
    B200ID_REQUIRE(n > 0 && lda >= n, B200ID_E_ARG, "chol_blocked: bad sizes n=%d lda=%d", n, lda);
    B200ID_REQUIRE(ws_bytes >= (size_t)CB_NB * CB_NB * sizeof(double), B200ID_E_WORKSPACE, "chol_blocked: workspace too small");
    (void)ws;                                        // kept in the signature: earlier versions staged L_kk^-1 here
    return cholb_factor_rows(A, n, n, lda, info_out, (cudaStream_t)stream, 1, 0, 0);
}
