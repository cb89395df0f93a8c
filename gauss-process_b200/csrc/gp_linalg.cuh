// gp_linalg.cuh -- in-shared-memory Cholesky and triangular solves shared by the GP kernels.
#pragma once
#include "common.cuh"

namespace b200id {

// ------------------------------------------------------------------------------------------
// In-shared-memory Cholesky of the n x n lower triangle stored column-major at A[j*ld + i],
// with `extra` additional rows (0 or 1) carried through the updates.  Returns (via *status in
// shared memory) 0 = ok, k+1 = non-positive pivot at column k.  All threads must call.
__device__ inline void chol_smem(double* A, int n, int ld, int extra, int* status) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int rows = n + extra;
    for (int k = 0; k < n; ++k) {
        double* colk = A + (size_t)k * ld;
        const double piv = colk[k];
        // pivot <= 0 -> failure (LinAlgError branch).  NaN compares false and the factorisation goes on
        // (OpenBLAS potf2 has no isnan test).  A +inf pivot (overflowed Gram) also counts as failure: in
        // the reference the next inf*0 / inf-inf inside potrf raises the FP-invalid flag, which NumPy
        // turns into LinAlgError (numpy.linalg.cholesky runs under errstate(invalid='call')).
        if (piv <= 0.0 || piv == INFINITY) {
            if (tid == 0) *status = k + 1;
            __syncthreads();
            return;
        }
        const double d = sqrt(piv);
        const double inv = 1.0 / d;     // potf2 scales the column by the reciprocal (SCAL), not by division
        __syncthreads();                // everyone has read the pivot before it is overwritten
        for (int i = k + tid; i < rows; i += nt) colk[i] = (i == k) ? d : colk[i] * inv;
        __syncthreads();
        // trailing update: A[i][j] -= L[i][k] * L[j][k] for k < j <= i < rows, j < n.
        // Threads sweep rows (contiguous in shared memory) of one column after another.
        const int m = rows - (k + 1);
        const int ncols = n - (k + 1);
        // flatten (j, i) over the rectangle ncols x m and skip i < j: for n <= 160 the wasted
        // half is cheaper than the index arithmetic of a packed triangle.
        for (int e = tid; e < ncols * m; e += nt) {
            const int jj = e / m, ii = e - jj * m;
            if (ii >= jj) {
                const int j = k + 1 + jj, i = k + 1 + ii;
                A[(size_t)j * ld + i] = fma(-colk[i], colk[j], A[(size_t)j * ld + i]);
            }
        }
        __syncthreads();
    }
}

// alpha = L^-T z by column-oriented back substitution; z in/out (length n).  All threads call.
__device__ inline void backsolve_smem(const double* A, int n, int ld, double* z) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = n - 1; k >= 0; --k) {
        __syncthreads();
        const double ak = z[k] / A[(size_t)k * ld + k];
        __syncthreads();
        if (tid == 0) z[k] = ak;
        // z[j] -= L[k][j] * alpha[k] for j < k   (L[k][j] sits at A[j*ld + k])
        for (int j = tid; j < k; j += nt) z[j] = fma(-A[(size_t)j * ld + k], ak, z[j]);
    }
    __syncthreads();
}

}  // namespace b200id
