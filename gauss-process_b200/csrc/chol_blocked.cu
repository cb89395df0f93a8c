// chol_blocked.cu -- K4: blocked FP64 Cholesky for large N_d (BASELINE config 4, N_d = 2048).
//
// Reference: np.linalg.cholesky at src/gpr/gpr_fitting.py:80 / grid_search.py:103 on a Gram matrix
// too large for one CTA's shared memory.  Right-looking blocked algorithm on the row-major lower
// triangle in global memory (32 MB at n = 2048, i.e. L2 resident on B200):
//   for each 64-wide block column k:
//     (1) diag : one CTA factorises the 64x64 diagonal block in shared memory and inverts L_kk
//     (2) panel: P = A[k+64:, k:k+64] L_kk^-T          -- 64x64x64 FP64 tile products
//     (3) syrk : A[i,j] -= P_i P_j^T for tiles i >= j  -- 64x64x64 FP64 tile products
// Steps (2) and (3) share one register-tiled micro-kernel: 256 threads, each a 4x4 block of the
// 64x64 output tile, operands staged k-major in shared memory so that a thread's 4 rows / 4 columns
// are one 32-byte vector load and the FP64 pipe (16 DFMA per 8 loaded doubles) is the bound.
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
    for (int e = threadIdx.x; e < CB_NB * CB_NB; e += CB_THREADS) {
        const int r = e / CB_NB, k = e - r * CB_NB;
        T[k * CB_LDT + r] = (r0 + r < r_end) ? M[(size_t)(r0 + r) * ld + c0 + k] : 0.0;
    }
}

// (1) diagonal block: factorise in shared memory, write L back, write L^-1 (row-major [c][p], lower).
__global__ void __launch_bounds__(CB_THREADS)
cholb_diag_kernel(double* __restrict__ A, int lda, int n, int k0, double* __restrict__ Linv, int* __restrict__ info) {
    __shared__ double D[CB_NB * (CB_NB + 1)];      // column-major, ld = 65
    __shared__ int status;
    if (info[0] != 0) return;                       // an earlier block already failed
    const int nb = min(CB_NB, n - k0);
    const int ld = CB_NB + 1;
    const int tid = threadIdx.x;
    if (tid == 0) status = 0;
    for (int e = tid; e < nb * nb; e += CB_THREADS) {
        const int i = e / nb, j = e - i * nb;
        if (i >= j) D[j * ld + i] = A[(size_t)(k0 + i) * lda + k0 + j];
    }
    __syncthreads();
    chol_smem(D, nb, ld, 0, &status);
    if (status != 0) {
        if (tid == 0) info[0] = k0 + status;
        return;
    }
    for (int e = tid; e < nb * nb; e += CB_THREADS) {
        const int i = e / nb, j = e - i * nb;
        if (i >= j) A[(size_t)(k0 + i) * lda + k0 + j] = D[j * ld + i];
    }
    // L^-1 column c by forward substitution (thread c), stored Linv[i*CB_NB + c] = (L^-1)[i][c]
    for (int c = tid; c < CB_NB; c += CB_THREADS) {
        for (int i = 0; i < CB_NB; ++i) {
            double v = 0.0;
            if (c < nb && i < nb && i >= c) {
                double s = (i == c) ? 1.0 : 0.0;
                for (int p = c; p < i; ++p) s = fma(-D[p * ld + i], Linv[p * CB_NB + c], s);
                v = s / D[i * ld + i];
            }
            Linv[i * CB_NB + c] = v;
        }
    }
}

// (2) panel: rows [k0+64, n) of block column k0:  P[r][c] = sum_{p<=c} A[r][k0+p] * Linv[c][p]
__global__ void __launch_bounds__(CB_THREADS)
cholb_panel_kernel(double* __restrict__ A, int lda, int n, int k0, const double* __restrict__ Linv,
                   const int* __restrict__ info) {
    extern __shared__ __align__(16) double cb_sm[];
    double* As = cb_sm;
    double* Bs = cb_sm + CB_NB * CB_LDT;
    if (info[0] != 0) return;
    const int r0 = k0 + CB_NB + blockIdx.x * CB_NB;
    stage_tile_kmajor(A, lda, r0, n, k0, As);
    // Bs[p][c] = Linv[c][p]
    for (int e = threadIdx.x; e < CB_NB * CB_NB; e += CB_THREADS) {
        const int c = e / CB_NB, p = e - c * CB_NB;
        Bs[p * CB_LDT + c] = Linv[c * CB_NB + p];
    }
    __syncthreads();
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    double acc[4][4] = {};
    tile_mm(As, Bs, CB_NB, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + 4 * ty + i;
        if (r < n) {
#pragma unroll
            for (int j = 0; j < 4; ++j) A[(size_t)r * lda + k0 + 4 * tx + j] = acc[i][j];
        }
    }
}

// (3) trailing update over lower-triangle tiles (ti >= tj) of the matrix starting at s = k0 + 64
__global__ void __launch_bounds__(CB_THREADS)
cholb_syrk_kernel(double* __restrict__ A, int lda, int n, int k0, const int* __restrict__ info) {
    extern __shared__ __align__(16) double cb_sm[];
    double* As = cb_sm;
    double* Bs = cb_sm + CB_NB * CB_LDT;
    if (info[0] != 0) return;
    // linear tile index -> (ti, tj), ti >= tj
    const int lin = blockIdx.x;
    int ti = (int)((sqrt(8.0 * lin + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= lin) ++ti;
    while (ti * (ti + 1) / 2 > lin) --ti;
    const int tj = lin - ti * (ti + 1) / 2;
    const int s = k0 + CB_NB;
    const int r0 = s + ti * CB_NB, c0 = s + tj * CB_NB;
    stage_tile_kmajor(A, lda, r0, n, k0, As);
    stage_tile_kmajor(A, lda, c0, n, k0, Bs);
    __syncthreads();
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    double acc[4][4] = {};
    tile_mm(As, Bs, CB_NB, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + 4 * ty + i;
        if (r >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * tx + j;
            if (c <= r) A[(size_t)r * lda + c] -= acc[i][j];
        }
    }
}

}  // namespace b200id

using namespace b200id;

extern "C" size_t b200id_chol_blocked_workspace_bytes(int n) {
    (void)n;
    return align_up((size_t)CB_NB * CB_NB * sizeof(double), 256);
}

extern "C" int b200id_chol_blocked(double* A, int n, int lda, int* info_out, void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(A && info_out && ws, B200ID_E_ARG, "chol_blocked: null pointer");
    B200ID_REQUIRE(n > 0 && lda >= n, B200ID_E_ARG, "chol_blocked: bad sizes n=%d lda=%d", n, lda);
    B200ID_REQUIRE(ws_bytes >= (size_t)CB_NB * CB_NB * sizeof(double), B200ID_E_WORKSPACE, "chol_blocked: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemsetAsync(info_out, 0, sizeof(int), st), "cudaMemsetAsync(info)");
    if (rc) return rc;
    double* Linv = (double*)ws;
    const size_t smem = 2 * (size_t)CB_NB * CB_LDT * sizeof(double);
    rc = check_cuda(cudaFuncSetAttribute(cholb_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(panel)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(cholb_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(syrk)");
    if (rc) return rc;
    for (int k0 = 0; k0 < n; k0 += CB_NB) {
        cholb_diag_kernel<<<1, CB_THREADS, 0, st>>>(A, lda, n, k0, Linv, info_out);
        const int m = n - (k0 + CB_NB);
        if (m <= 0) break;
        const int tiles = (m + CB_NB - 1) / CB_NB;
        cholb_panel_kernel<<<tiles, CB_THREADS, smem, st>>>(A, lda, n, k0, Linv, info_out);
        cholb_syrk_kernel<<<tiles * (tiles + 1) / 2, CB_THREADS, smem, st>>>(A, lda, n, k0, info_out);
    }
    B200ID_LAUNCH_CHECK("chol_blocked kernels");
    return B200ID_OK;
}
