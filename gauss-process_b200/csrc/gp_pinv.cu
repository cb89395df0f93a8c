// gp_pinv.cu -- pseudo-inverse branch of the final GP fit.
//
// Reference: src/gpr/gpr_fitting.py:84-87 -- when np.linalg.cholesky raises, K_inv = np.linalg.pinv(K)
// and alpha = K_inv @ y.  NumPy's pinv is an SVD with cutoff rcond * s_max, rcond = 1e-15; for the
// symmetric K + diag_add*I the singular values are |lambda_i|, so the same operator is obtained from
// the symmetric eigendecomposition  pinv = sum_{|lambda_i| > cutoff} v_i v_i^T / lambda_i.
// The eigendecomposition is a cyclic two-sided Jacobi method run by ONE CTA: A lives in shared
// memory, V in the caller's workspace (L2 resident); each round applies n/2 disjoint rotations
// chosen by the round-robin (chess tournament) ordering.
#include "common.cuh"
#include "gp_kernels.cuh"

namespace b200id {

constexpr int PINV_THREADS = 256;
constexpr int PINV_NMAX = 160;
constexpr int PINV_MAX_SWEEPS = 40;

__host__ __device__ inline int pinv_ld(int ne) { return ne | 1; }

__global__ void __launch_bounds__(PINV_THREADS)
gp_pinv_kernel(int kernel_id, const double* __restrict__ theta, const double* __restrict__ X,
               const double* __restrict__ y, int n, double diag_add, double rcond,
               double* __restrict__ alpha_out, double* __restrict__ Kinv_out, double* __restrict__ V /*[ne*ne]*/) {
    extern __shared__ double sm[];
    const int ne = (n + 1) & ~1;                 // even size; the padding row/col is zero
    const int ld = pinv_ld(ne);
    double* A = sm;                              // [ne][ld] row-major, symmetric
    double* rc = A + (size_t)ne * ld;            // [ne/2] cos
    double* rs = rc + ne / 2;                    // [ne/2] sin
    double* lam = rs + ne / 2;                   // [ne] eigenvalues, then reciprocal-or-zero
    double* scratch = lam + ne;                  // [8]
    __shared__ int rp[PINV_NMAX / 2 + 1], rq[PINV_NMAX / 2 + 1];
    __shared__ int done;

    const int tid = threadIdx.x;
    const KernelParams kp = make_kernel_params(kernel_id, theta);
    for (int e = tid; e < ne * ne; e += PINV_THREADS) {
        const int i = e / ne, j = e - i * ne;
        double v = 0.0;
        if (i < n && j < n) {
            // build from the lower triangle so that A is exactly symmetric
            const int a = i > j ? i : j, b = i > j ? j : i;
            v = kernel_entry(kp, X[a], X[b]);
            if (i == j) v = B200ID_ADD(v, diag_add);
        }
        A[(size_t)i * ld + j] = v;
        V[(size_t)i * ne + j] = (i == j) ? 1.0 : 0.0;
    }
    if (tid == 0) done = 0;
    __syncthreads();

    const int half = ne / 2;
    for (int sweep = 0; sweep < PINV_MAX_SWEEPS; ++sweep) {
        // convergence: off-diagonal Frobenius mass against the total
        double off = 0.0, tot = 0.0;
        for (int e = tid; e < ne * ne; e += PINV_THREADS) {
            const int i = e / ne, j = e - i * ne;
            const double v = A[(size_t)i * ld + j];
            tot = fma(v, v, tot);
            if (i != j) off = fma(v, v, off);
        }
        off = block_sum(off, scratch);
        tot = block_sum(tot, scratch);
        if (tid == 0 && (off <= 1e-32 * tot || !(tot > 0.0))) done = 1;
        __syncthreads();
        if (done) break;
        for (int r = 0; r < ne - 1; ++r) {
            // pairs of this round + their rotations
            for (int k = tid; k < half; k += PINV_THREADS) {
                int p, q;
                if (k == 0) { p = r % (ne - 1); q = ne - 1; }
                else { p = (r + k) % (ne - 1); q = (r - k + (ne - 1)) % (ne - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                rp[k] = p; rq[k] = q;
                const double apq = A[(size_t)p * ld + q];
                double c = 1.0, s = 0.0;
                if (apq != 0.0) {
                    const double tau = (A[(size_t)q * ld + q] - A[(size_t)p * ld + p]) / (2.0 * apq);
                    const double t = (tau >= 0.0) ? 1.0 / (tau + sqrt(1.0 + tau * tau)) : 1.0 / (tau - sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                }
                rc[k] = c; rs[k] = s;
            }
            __syncthreads();
            // A <- A J, V <- V J : columns p, q of every row
            for (int e = tid; e < half * ne; e += PINV_THREADS) {
                const int k = e / ne, i = e - k * ne;
                const int p = rp[k], q = rq[k];
                const double c = rc[k], s = rs[k];
                const double aip = A[(size_t)i * ld + p], aiq = A[(size_t)i * ld + q];
                A[(size_t)i * ld + p] = c * aip - s * aiq;
                A[(size_t)i * ld + q] = s * aip + c * aiq;
                const double vip = V[(size_t)i * ne + p], viq = V[(size_t)i * ne + q];
                V[(size_t)i * ne + p] = c * vip - s * viq;
                V[(size_t)i * ne + q] = s * vip + c * viq;
            }
            __syncthreads();
            // A <- J^T A : rows p, q of every column
            for (int e = tid; e < half * ne; e += PINV_THREADS) {
                const int k = e / ne, j = e - k * ne;
                const int p = rp[k], q = rq[k];
                const double c = rc[k], s = rs[k];
                const double apj = A[(size_t)p * ld + j], aqj = A[(size_t)q * ld + j];
                A[(size_t)p * ld + j] = c * apj - s * aqj;
                A[(size_t)q * ld + j] = s * apj + c * aqj;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // cutoff on |lambda| relative to the largest, as np.linalg.pinv does on singular values
    double mx = 0.0;
    for (int i = tid; i < n; i += PINV_THREADS) mx = fmax(mx, fabs(A[(size_t)i * ld + i]));
    // block max via shared scratch
    __shared__ double smax[PINV_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) smax[tid >> 5] = mx;
    __syncthreads();
    mx = 0.0;
    for (int w = 0; w < PINV_THREADS / 32; ++w) mx = fmax(mx, smax[w]);
    const double cutoff = rcond * mx;
    // the zero padding eigenpair (index set containing row n when n is odd) is dropped by the cutoff:
    // its eigenvalue is exactly 0.
    for (int i = tid; i < ne; i += PINV_THREADS) {
        const double l = A[(size_t)i * ld + i];
        lam[i] = (fabs(l) > cutoff) ? 1.0 / l : 0.0;
    }
    __threadfence_block();
    __syncthreads();
    for (int e = tid; e < n * n; e += PINV_THREADS) {
        const int i = e / n, j = e - i * n;
        if (j > i) continue;
        double s = 0.0;
        for (int k = 0; k < ne; ++k) s = fma(V[(size_t)i * ne + k] * lam[k], V[(size_t)j * ne + k], s);
        Kinv_out[(size_t)i * n + j] = s;
        Kinv_out[(size_t)j * n + i] = s;
    }
    __threadfence_block();
    __syncthreads();
    for (int i = tid; i < n; i += PINV_THREADS) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = fma(Kinv_out[(size_t)i * n + j], y[j], s);
        alpha_out[i] = s;
    }
}

}  // namespace b200id

using namespace b200id;

extern "C" int b200id_gp_fit_pinv(int kernel_id, const double* theta, const double* X, const double* y, int n,
                                  double diag_add, double rcond, double* alpha_out, double* Kinv_out,
                                  void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(theta && X && y && alpha_out && Kinv_out && ws, B200ID_E_ARG, "gp_fit_pinv: null pointer");
    B200ID_REQUIRE(kernel_id >= 0 && kernel_id < B200ID_K_COUNT, B200ID_E_ARG, "gp_fit_pinv: unknown kernel id %d", kernel_id);
    B200ID_REQUIRE(n > 0 && n <= PINV_NMAX, B200ID_E_UNSUPPORTED, "gp_fit_pinv: n=%d outside (0, %d]", n, PINV_NMAX);
    const int ne = (n + 1) & ~1;
    B200ID_REQUIRE(ws_bytes >= (size_t)ne * ne * sizeof(double), B200ID_E_WORKSPACE, "gp_fit_pinv: workspace too small");
    const size_t smem = ((size_t)ne * pinv_ld(ne) + 2 * (size_t)ne + 16) * sizeof(double);
    int rc = check_cuda(cudaFuncSetAttribute(gp_pinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(pinv)");
    if (rc) return rc;
    gp_pinv_kernel<<<1, PINV_THREADS, smem, (cudaStream_t)stream>>>(kernel_id, theta, X, y, n, diag_add, rcond, alpha_out,
                                                                   Kinv_out, (double*)ws);
    B200ID_LAUNCH_CHECK("gp_pinv_kernel");
    return B200ID_OK;
}
