// gp_batch.cu -- K2/K3: Gram matrices, batched candidate evaluation, final fit and predict.
//
// Reference call sites: src/gpr/grid_search.py:91-133 (candidate metric), :186-196 (argmin scan),
// src/gpr/gpr_fitting.py:73-100 (fit / predict) and :104-123 (NLL).
//
// Batched evaluation: ONE CTA PER CANDIDATE.  The CTA builds the lower triangle of
// K + noise*I in shared memory (column-major, odd leading dimension), appends y as an extra ROW
// (row n), and runs a right-looking Cholesky on the (n+1) x n array: after the factorisation row n
// holds z = L^-1 y, so the forward solve is free and y^T K^-1 y = z.z.  NLL needs nothing else;
// the validation-RMSE metric back-substitutes alpha = L^-T z and forms K(Xv, X) alpha on the fly.
// Failure conventions (hard part 4): pivot <= 0 -> metric 1e10 (LinAlgError
// branch, grid_search.py:132-133); NaN pivot -> factorisation continues and the metric is NaN.
#include <stdlib.h>
#include "common.cuh"
#include "gp_kernels.cuh"
#include "gp_linalg.cuh"

namespace b200id {

constexpr int GP_THREADS = 128;
constexpr int GP_NMAX_SMEM = 160;   // (n+1) * ld doubles must fit in 227 KB

__host__ __device__ inline int gp_ld(int n) { return (n + 1) | 1; }   // rows n+1 (y row), odd stride
static inline size_t gp_smem_bytes(int n) {
    return ((size_t)n * gp_ld(n) + 3 * (size_t)n + 64) * sizeof(double);
}

__global__ void __launch_bounds__(GP_THREADS)
gp_batch_kernel(int kernel_id, const int* __restrict__ kernel_ids, const double* __restrict__ theta,
                const double* __restrict__ X, const double* __restrict__ y, int n, double noise,
                int metric, const double* __restrict__ Xv, const double* __restrict__ yv, int nv,
                double y_mean, double y_scale, double* __restrict__ metric_out,
                const double* __restrict__ noise_v /*per-candidate diagonal term, or NULL*/) {
    extern __shared__ double sm[];
    const int ld = gp_ld(n);
    double* A = sm;                              // [n][ld]
    double* xs = A + (size_t)n * ld;             // [n]
    double* z = xs + n;                          // [n]
    double* scratch = z + n;                     // [GP_THREADS/32 + ...]
    __shared__ int status;

    const int64_t cand = blockIdx.x;
    const int tid = threadIdx.x;
    const int kid = kernel_ids ? kernel_ids[cand] : kernel_id;
    const KernelParams kp = make_kernel_params(kid, theta + cand * B200ID_MAX_PARAMS);
    if (noise_v) noise = noise_v[cand];

    if (tid == 0) status = 0;
    for (int i = tid; i < n; i += GP_THREADS) xs[i] = X[i];
    __syncthreads();
    // lower triangle of K + noise*I, plus y as row n
    for (int e = tid; e < n * (n + 1); e += GP_THREADS) {
        const int j = e / (n + 1), i = e - j * (n + 1);
        if (i == n) {
            A[(size_t)j * ld + n] = y[j];
        } else if (i >= j) {
            double v = kernel_entry(kp, xs[i], xs[j]);
            if (i == j) v = B200ID_ADD(v, noise);
            A[(size_t)j * ld + i] = v;
        }
    }
    __syncthreads();
    chol_smem(A, n, ld, 1, &status);
    if (status != 0) {
        if (tid == 0) metric_out[cand] = B200ID_FAIL_METRIC;
        return;
    }
    if (metric == B200ID_METRIC_NLL) {
        // 0.5 y.alpha + sum log L_ii + 0.5 n log(2 pi)        grid_search.py:128-130
        double part = 0.0;
        for (int j = tid; j < n; j += GP_THREADS) {
            const double zj = A[(size_t)j * ld + n];
            part += 0.5 * zj * zj + log(A[(size_t)j * ld + j]);
        }
        const double tot = block_sum(part, scratch);
        if (tid == 0) metric_out[cand] = tot + 0.5 * (double)n * 1.8378770664093453;  // log(2 pi)
        return;
    }
    // validation RMSE on the de-standardised prediction                  grid_search.py:107-126
    for (int j = tid; j < n; j += GP_THREADS) z[j] = A[(size_t)j * ld + n];
    backsolve_smem(A, n, ld, z);
    double se = 0.0;
    for (int v = tid; v < nv; v += GP_THREADS) {
        const double xv = Xv[v];
        double pred = 0.0;
        for (int j = 0; j < n; ++j) pred = fma(kernel_entry(kp, xv, xs[j]), z[j], pred);
        pred = pred * y_scale + y_mean;
        const double e = yv[v] - pred;
        se = fma(e, e, se);
    }
    const double tot = block_sum(se, scratch);
    if (tid == 0) metric_out[cand] = sqrt(tot / (double)nv);
}

// ------------------------------------------------------------------------------------------
// first strict minimum in index order, NaNs skipped  (grid_search.py:186-196)
constexpr int ARGMIN_THREADS = 1024;
__global__ void __launch_bounds__(ARGMIN_THREADS)
first_strict_min_kernel(const double* __restrict__ metric, int64_t n, double* best_out, int64_t* idx_out) {
    __shared__ double sv[ARGMIN_THREADS];
    __shared__ long long si[ARGMIN_THREADS];
    double bv = INFINITY;
    long long bi = -1;
    for (int64_t i = threadIdx.x; i < n; i += ARGMIN_THREADS) {
        const double v = metric[i];
        if (v < bv) { bv = v; bi = i; }          // increasing i per thread: keeps the first of equals
    }
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = ARGMIN_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double ov = sv[threadIdx.x + s];
            const long long oi = si[threadIdx.x + s];
            const double mv = sv[threadIdx.x];
            const long long mi = si[threadIdx.x];
            if (oi >= 0 && (mi < 0 || ov < mv || (ov == mv && oi < mi))) {
                sv[threadIdx.x] = ov;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        best_out[0] = sv[0];
        idx_out[0] = si[0];
    }
}

// Same rule per segment [offsets[s], offsets[s+1]) of one metric array, one CTA per segment: the per-kernel minima of a
// mixed-kernel candidate population (BASELINE config 3: 11 kernels x 4096 start points in one launch).
// pend_out[s] = (best metric or +inf, index into metric[] as a double or -1) -- the layout b200id_argmin_pack takes.
__global__ void __launch_bounds__(ARGMIN_THREADS)
first_strict_min_segments_kernel(const double* __restrict__ metric, const int64_t* __restrict__ offsets,
                                 double* __restrict__ pend_out) {
    __shared__ double sv[ARGMIN_THREADS];
    __shared__ long long si[ARGMIN_THREADS];
    const int64_t lo = offsets[blockIdx.x], hi = offsets[blockIdx.x + 1];
    double bv = INFINITY;
    long long bi = -1;
    for (int64_t i = lo + threadIdx.x; i < hi; i += ARGMIN_THREADS) {
        const double v = metric[i];
        if (v < bv) { bv = v; bi = i; }
    }
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = ARGMIN_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double ov = sv[threadIdx.x + s];
            const long long oi = si[threadIdx.x + s];
            const double mv = sv[threadIdx.x];
            const long long mi = si[threadIdx.x];
            if (oi >= 0 && (mi < 0 || ov < mv || (ov == mv && oi < mi))) {
                sv[threadIdx.x] = ov;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pend_out[2 * blockIdx.x] = sv[0];
        pend_out[2 * blockIdx.x + 1] = (double)si[0];
    }
}

// ------------------------------------------------------------------------------------------
__global__ void gram_kernel(int kernel_id, const double* __restrict__ theta, const double* __restrict__ X1,
                            int n1, const double* __restrict__ X2, int n2, double* __restrict__ K) {
    const KernelParams kp = make_kernel_params(kernel_id, theta);
    const int64_t total = (int64_t)n1 * n2;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / n2), j = (int)(e - (int64_t)i * n2);
        K[e] = kernel_entry(kp, X1[i], X2[j]);
    }
}

// ------------------------------------------------------------------------------------------
// Final fit (single problem): Cholesky in shared memory, alpha, then K_inv = L^-T L^-1 through
// global scratch (Kinv_out first holds L^-1, column j solved by thread j).
__global__ void __launch_bounds__(GP_THREADS)
gp_fit_kernel(int kernel_id, const double* __restrict__ theta, const double* __restrict__ X,
              const double* __restrict__ y, int n, double diag_add, double* __restrict__ alpha_out,
              double* __restrict__ Linv /*[n*n] scratch*/, double* __restrict__ Kinv_out, int* __restrict__ info_out) {
    extern __shared__ double sm[];
    const int ld = gp_ld(n);
    double* A = sm;
    double* xs = A + (size_t)n * ld;
    double* z = xs + n;
    __shared__ int status;
    const int tid = threadIdx.x;
    const KernelParams kp = make_kernel_params(kernel_id, theta);
    if (tid == 0) status = 0;
    for (int i = tid; i < n; i += GP_THREADS) xs[i] = X[i];
    __syncthreads();
    for (int e = tid; e < n * (n + 1); e += GP_THREADS) {
        const int j = e / (n + 1), i = e - j * (n + 1);
        if (i == n) {
            A[(size_t)j * ld + n] = y[j];
        } else if (i >= j) {
            double v = kernel_entry(kp, xs[i], xs[j]);
            if (i == j) v = B200ID_ADD(v, diag_add);
            A[(size_t)j * ld + i] = v;
        }
    }
    __syncthreads();
    chol_smem(A, n, ld, 1, &status);
    if (tid == 0) info_out[0] = status;
    if (status != 0) return;
    for (int j = tid; j < n; j += GP_THREADS) z[j] = A[(size_t)j * ld + n];
    backsolve_smem(A, n, ld, z);
    for (int j = tid; j < n; j += GP_THREADS) alpha_out[j] = z[j];
    // L^-1: column c solves L v = e_c (forward substitution), v stored at Linv[c*n + i], i >= c
    for (int c = tid; c < n; c += GP_THREADS) {
        double* v = Linv + (size_t)c * n;
        for (int i = 0; i < c; ++i) v[i] = 0.0;
        for (int i = c; i < n; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int p = c; p < i; ++p) s = fma(-A[(size_t)p * ld + i], v[p], s);
            v[i] = s / A[(size_t)i * ld + i];
        }
    }
    __threadfence_block();
    __syncthreads();
    // K_inv[i][j] = sum_{k >= max(i,j)} Linv[k][i] Linv[k][j]   (Linv[k][i] at Linv[i*n + k])
    for (int e = tid; e < n * n; e += GP_THREADS) {
        const int i = e / n, j = e - i * n;
        if (j > i) continue;
        double s = 0.0;
        const double* vi = Linv + (size_t)i * n;
        const double* vj = Linv + (size_t)j * n;
        for (int k = i; k < n; ++k) s = fma(vi[k], vj[k], s);
        Kinv_out[(size_t)i * n + j] = s;
        Kinv_out[(size_t)j * n + i] = s;
    }
}

// K_inv = (L L^T)^-1 column by column: one warp per column c solves L v = e_c then L^T x = v with the
// dense lower factor staged in shared memory (row-major, odd stride).
constexpr int KINV_WARPS = 8;
__global__ void __launch_bounds__(KINV_WARPS * 32)
gp_kinv_kernel(const double* __restrict__ L, int n, const int* __restrict__ info, double* __restrict__ Kinv) {
    extern __shared__ double sm[];
    if (info[0] != 0) return;
    const int ld = n | 1;
    double* Ls = sm;                               // [n][ld]
    double* vbuf = Ls + (size_t)n * ld;            // [KINV_WARPS][n]
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e - i * n;
        Ls[(size_t)i * ld + j] = L[e];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * KINV_WARPS + w;
    if (c >= n) return;
    double* v = vbuf + (size_t)w * n;
    for (int i = lane; i < n; i += 32) v[i] = 0.0;
    __syncwarp();
    for (int i = c; i < n; ++i) {                  // forward: v[i] = (e_c[i] - sum_{c<=p<i} L[i][p] v[p]) / L[i][i]
        double s = 0.0;
        for (int p = c + lane; p < i; p += 32) s = fma(Ls[(size_t)i * ld + p], v[p], s);
        s = warp_sum(s);
        if (lane == 0) v[i] = (((i == c) ? 1.0 : 0.0) - s) / Ls[(size_t)i * ld + i];
        __syncwarp();
    }
    for (int i = n - 1; i >= 0; --i) {             // backward: x[i] = (v[i] - sum_{p>i} L[p][i] x[p]) / L[i][i]
        double s = 0.0;
        for (int p = i + 1 + lane; p < n; p += 32) s = fma(Ls[(size_t)p * ld + i], v[p], s);
        s = warp_sum(s);
        if (lane == 0) v[i] = (v[i] - s) / Ls[(size_t)i * ld + i];
        __syncwarp();
    }
    for (int i = lane; i < n; i += 32) Kinv[(size_t)i * n + c] = v[i];
}

// One CTA per test point: mean = K* alpha; var = k(xs,xs) - sum_j (K* K_inv)_j K*_j   (gpr_fitting.py:91-97)
__global__ void __launch_bounds__(GP_THREADS)
gp_predict_kernel(int kernel_id, const double* __restrict__ theta, const double* __restrict__ X, int n,
                  const double* __restrict__ alpha, const double* __restrict__ Kinv,
                  const double* __restrict__ Xs, int m, double* __restrict__ mean_out, double* __restrict__ std_out) {
    extern __shared__ double sm[];
    double* ks = sm;                     // [n] row of K*
    double* scratch = ks + n;
    const int tid = threadIdx.x;
    const KernelParams kp = make_kernel_params(kernel_id, theta);
    for (int q = blockIdx.x; q < m; q += gridDim.x) {
        const double xq = Xs[q];
        __syncthreads();
        double part = 0.0;
        for (int j = tid; j < n; j += GP_THREADS) {
            const double kv = kernel_entry(kp, xq, X[j]);
            ks[j] = kv;
            part = fma(kv, alpha[j], part);
        }
        const double mean = block_sum(part, scratch);   // contains the __syncthreads that publishes ks
        if (tid == 0) mean_out[q] = mean;
        if (std_out) {
            double acc = 0.0;
            for (int j = tid; j < n; j += GP_THREADS) {
                double v = 0.0;
                for (int i = 0; i < n; ++i) v = fma(ks[i], Kinv[(size_t)i * n + j], v);
                acc = fma(v, ks[j], acc);
            }
            const double quad = block_sum(acc, scratch);
            if (tid == 0) {
                const double var = kernel_entry(kp, xq, xq) - quad;
                std_out[q] = sqrt(fmax(var, 0.0));
            }
        }
    }
}

}  // namespace b200id

using namespace b200id;

// gp_batch_tiled.cu: blocked register-tiled path for n <= 127
int b200id_gp_fit_tiled_launch(int kernel_id, const double* theta, const double* X, const double* y, int n,
                               double diag_add, double* alpha_out, double* L_out, int* info_out, cudaStream_t st,
                               int n_problems = 1);
int b200id_gp_batch_tiled_launch(int kernel_id, const int* kernel_ids, const double* theta, int64_t n_candidates,
                                 const double* X, const double* y, int n, double noise, int metric,
                                 const double* Xv, const double* yv, int nv, double y_mean, double y_scale,
                                 double* metric_out, const double* noise_v, cudaStream_t st,
                                 const double* y_b = nullptr, const double* yv_b = nullptr, double y_mean_b = 0.0,
                                 double y_scale_b = 1.0, double* metric_out_b = nullptr,
                                 const double* shape_tab = nullptr, const int* shape_of = nullptr);
size_t b200id_gp_shape_table_doubles(int n_shapes, int n, int nv, int nrhs);
int b200id_gp_shape_table_launch(int kernel_id, const double* shape_theta, int n_shapes, const double* X, int n,
                                 const double* Xv, int nv, int nrhs, double* tab, cudaStream_t st);
constexpr int GT_PAIR_NMAX = 126;       // n + 2 rows must fit the 128-row warp map of the tiled kernel

// gp_large.cu: global-memory path for n > 160 (blocked Cholesky)
size_t b200id_gp_large_workspace_bytes(int n, int nv);
int b200id_gp_fit_large(int kernel_id, const double* theta, const double* X, const double* y, int n, double diag_add,
                        double* alpha_out, double* Kinv_out, int* info_out, void* ws, size_t ws_bytes, cudaStream_t st);
int b200id_gp_batch_eval_large(int kernel_id, const int* kernel_ids_host_or_null, const double* theta, int64_t n_candidates,
                               const double* X, const double* y, int n, double noise, int metric, const double* Xv,
                               const double* yv, int nv, double y_mean, double y_scale, double* metric_out,
                               const double* noise_v, void* ws, size_t ws_bytes, cudaStream_t st);
constexpr int GP_LARGE_NMAX = 8192;
constexpr int GP_LARGE_NV_CAP = 8192;

static bool force_generic_gp_path() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("B200ID_GP_PATH");
        cached = (e && strcmp(e, "generic") == 0) ? 1 : 0;
    }
    return cached == 1;
}

static int set_smem_attr(const void* fn, size_t bytes) {
    return check_cuda(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "cudaFuncSetAttribute(smem)");
}

extern "C" int b200id_gp_gram(int kernel_id, const double* theta, const double* X1, int n1, const double* X2,
                              int n2, double* K_out, void* stream) {
    B200ID_REQUIRE(theta && X1 && K_out && n1 > 0, B200ID_E_ARG, "gp_gram: bad argument");
    B200ID_REQUIRE(kernel_id >= 0 && kernel_id < B200ID_K_COUNT, B200ID_E_ARG, "gp_gram: unknown kernel id %d", kernel_id);
    if (!X2) { X2 = X1; n2 = n1; }
    B200ID_REQUIRE(n2 > 0, B200ID_E_ARG, "gp_gram: n2 <= 0");
    const int64_t total = (int64_t)n1 * n2;
    int grid = (int)((total + 255) / 256);
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    gram_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kernel_id, theta, X1, n1, X2, n2, K_out);
    B200ID_LAUNCH_CHECK("gram_kernel");
    return B200ID_OK;
}

extern "C" size_t b200id_gp_batch_workspace_bytes(int64_t n_candidates, int n) {
    // large n: one slice per concurrently factorised candidate (up to 16 lanes, see b200id_gp_batch_eval_large)
    if (n > GP_NMAX_SMEM) {
        const size_t lanes = (size_t)(n_candidates < 1 ? 1 : (n_candidates > 16 ? 16 : n_candidates));
        return lanes * (b200id_gp_large_workspace_bytes(n, GP_LARGE_NV_CAP) + 512);
    }
    return 256;     // the shared-memory paths need no global scratch
}

static int gp_batch_eval_impl(int kernel_id, const int* kernel_ids, const double* theta, int64_t n_candidates,
                              const double* X, const double* y, int n, double noise, const double* noise_v, int metric,
                              const double* Xv, const double* yv, int nv, double y_mean, double y_scale,
                              double* metric_out, void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(theta && X && y && metric_out, B200ID_E_ARG, "gp_batch_eval: null pointer");
    B200ID_REQUIRE(n_candidates > 0 && n > 0, B200ID_E_ARG, "gp_batch_eval: empty problem");
    B200ID_REQUIRE(kernel_ids || (kernel_id >= 0 && kernel_id < B200ID_K_COUNT), B200ID_E_ARG, "gp_batch_eval: unknown kernel id %d", kernel_id);
    B200ID_REQUIRE(metric == B200ID_METRIC_NLL || metric == B200ID_METRIC_VAL_RMSE, B200ID_E_ARG, "gp_batch_eval: unknown metric %d", metric);
    B200ID_REQUIRE(metric == B200ID_METRIC_NLL || (Xv && yv && nv > 0), B200ID_E_ARG, "gp_batch_eval: validation metric needs Xv, yv");
    if (n > GP_NMAX_SMEM) {
        B200ID_REQUIRE(n <= GP_LARGE_NMAX && nv <= GP_LARGE_NV_CAP, B200ID_E_UNSUPPORTED, "gp_batch_eval: n=%d / nv=%d beyond the large-n limits", n, nv);
        B200ID_REQUIRE(ws, B200ID_E_ARG, "gp_batch_eval: the large-n path needs a workspace");
        return b200id_gp_batch_eval_large(kernel_id, kernel_ids, theta, n_candidates, X, y, n, noise, metric, Xv, yv, nv,
                                          y_mean, y_scale, metric_out, noise_v, ws, ws_bytes, (cudaStream_t)stream);
    }
    B200ID_REQUIRE(n_candidates < (1LL << 31), B200ID_E_UNSUPPORTED, "gp_batch_eval: too many candidates");
    if (n <= 127 && !force_generic_gp_path())
        return b200id_gp_batch_tiled_launch(kernel_id, kernel_ids, theta, n_candidates, X, y, n, noise, metric, Xv, yv, nv,
                                            y_mean, y_scale, metric_out, noise_v, (cudaStream_t)stream);
    const size_t smem = gp_smem_bytes(n);
    int rc = set_smem_attr((const void*)gp_batch_kernel, smem);
    if (rc) return rc;
    gp_batch_kernel<<<(unsigned)n_candidates, GP_THREADS, smem, (cudaStream_t)stream>>>(
        kernel_id, kernel_ids, theta, X, y, n, noise, metric, Xv, yv, nv, y_mean, y_scale, metric_out, noise_v);
    B200ID_LAUNCH_CHECK("gp_batch_kernel");
    return B200ID_OK;
}

extern "C" int b200id_gp_batch_eval(int kernel_id, const int* kernel_ids, const double* theta, int64_t n_candidates,
                                    const double* X, const double* y, int n, double noise, int metric,
                                    const double* Xv, const double* yv, int nv, double y_mean, double y_scale,
                                    double* metric_out, void* ws, size_t ws_bytes, void* stream) {
    return gp_batch_eval_impl(kernel_id, kernel_ids, theta, n_candidates, X, y, n, noise, nullptr, metric, Xv, yv, nv,
                              y_mean, y_scale, metric_out, ws, ws_bytes, stream);
}

extern "C" int b200id_gp_batch_eval_noise(int kernel_id, const int* kernel_ids, const double* theta, const double* noise,
                                          int64_t n_candidates, const double* X, const double* y, int n, int metric,
                                          const double* Xv, const double* yv, int nv, double y_mean, double y_scale,
                                          double* metric_out, void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(noise, B200ID_E_ARG, "gp_batch_eval_noise: null noise vector");
    return gp_batch_eval_impl(kernel_id, kernel_ids, theta, n_candidates, X, y, n, 0.0, noise, metric, Xv, yv, nv,
                              y_mean, y_scale, metric_out, ws, ws_bytes, stream);
}

extern "C" int b200id_gp_batch_eval_pair(int kernel_id, const int* kernel_ids, const double* theta, int64_t n_candidates,
                                         const double* X, const double* y_a, const double* y_b, int n, double noise, int metric,
                                         const double* Xv, const double* yv_a, const double* yv_b, int nv,
                                         double y_mean_a, double y_scale_a, double y_mean_b, double y_scale_b,
                                         double* metric_out_a, double* metric_out_b, void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(y_b && metric_out_b, B200ID_E_ARG, "gp_batch_eval_pair: null second target");
    const bool val = metric == B200ID_METRIC_VAL_RMSE;
    B200ID_REQUIRE(!val || yv_b, B200ID_E_ARG, "gp_batch_eval_pair: validation metric needs yv_b");
    if (n > GT_PAIR_NMAX || force_generic_gp_path()) {
        // no shared factorisation on the other paths: two plain evaluations
        int rc = gp_batch_eval_impl(kernel_id, kernel_ids, theta, n_candidates, X, y_a, n, noise, nullptr, metric, Xv, yv_a, nv,
                                    y_mean_a, y_scale_a, metric_out_a, ws, ws_bytes, stream);
        if (rc) return rc;
        return gp_batch_eval_impl(kernel_id, kernel_ids, theta, n_candidates, X, y_b, n, noise, nullptr, metric, Xv, yv_b, nv,
                                  y_mean_b, y_scale_b, metric_out_b, ws, ws_bytes, stream);
    }
    B200ID_REQUIRE(theta && X && y_a && metric_out_a, B200ID_E_ARG, "gp_batch_eval_pair: null pointer");
    B200ID_REQUIRE(n_candidates > 0 && n_candidates < (1LL << 31) && n > 0, B200ID_E_ARG, "gp_batch_eval_pair: bad problem size");
    B200ID_REQUIRE(kernel_ids || (kernel_id >= 0 && kernel_id < B200ID_K_COUNT), B200ID_E_ARG, "gp_batch_eval_pair: unknown kernel id %d", kernel_id);
    B200ID_REQUIRE(metric == B200ID_METRIC_NLL || metric == B200ID_METRIC_VAL_RMSE, B200ID_E_ARG, "gp_batch_eval_pair: unknown metric %d", metric);
    B200ID_REQUIRE(!val || (Xv && yv_a && nv > 0), B200ID_E_ARG, "gp_batch_eval_pair: validation metric needs Xv, yv");
    return b200id_gp_batch_tiled_launch(kernel_id, kernel_ids, theta, n_candidates, X, y_a, n, noise, metric, Xv, yv_a, nv,
                                        y_mean_a, y_scale_a, metric_out_a, nullptr, (cudaStream_t)stream,
                                        y_b, yv_b, y_mean_b, y_scale_b, metric_out_b);
}

extern "C" size_t b200id_gp_shared_workspace_bytes(int n_shapes, int n, int nv, int two_targets) {
    if (n_shapes < 1 || n < 1 || nv < 0) return 0;
    return b200id_gp_shape_table_doubles(n_shapes, n, nv, two_targets ? 2 : 1) * sizeof(double) + 256;
}

extern "C" int b200id_gp_batch_eval_shared(int kernel_id, const double* theta, int64_t n_candidates, const int* shape_of,
                                           const double* shape_theta, int n_shapes,
                                           const double* X, const double* y_a, const double* y_b, int n, double noise, int metric,
                                           const double* Xv, const double* yv_a, const double* yv_b, int nv,
                                           double y_mean_a, double y_scale_a, double y_mean_b, double y_scale_b,
                                           double* metric_out_a, double* metric_out_b, void* ws, size_t ws_bytes, void* stream) {
    const bool val = metric == B200ID_METRIC_VAL_RMSE;
    B200ID_REQUIRE(theta && shape_of && shape_theta && X && y_a && metric_out_a, B200ID_E_ARG, "gp_batch_eval_shared: null pointer");
    B200ID_REQUIRE((y_b == nullptr) == (metric_out_b == nullptr), B200ID_E_ARG, "gp_batch_eval_shared: second target and its output go together");
    B200ID_REQUIRE(n_candidates > 0 && n_candidates < (1LL << 31) && n > 0 && n_shapes > 0, B200ID_E_ARG, "gp_batch_eval_shared: bad problem size");
    B200ID_REQUIRE(kernel_id >= 0 && kernel_id < B200ID_K_COUNT, B200ID_E_ARG, "gp_batch_eval_shared: unknown kernel id %d", kernel_id);
    B200ID_REQUIRE(kernel_id != B200ID_K_DI, B200ID_E_UNSUPPORTED, "gp_batch_eval_shared: the DI kernel's amplitude is not theta[1]");
    B200ID_REQUIRE(metric == B200ID_METRIC_NLL || val, B200ID_E_ARG, "gp_batch_eval_shared: unknown metric %d", metric);
    B200ID_REQUIRE(!val || (Xv && yv_a && nv > 0 && (!y_b || yv_b)), B200ID_E_ARG, "gp_batch_eval_shared: validation metric needs Xv, yv");
    B200ID_REQUIRE(n <= (y_b ? GT_PAIR_NMAX : 127) && !force_generic_gp_path(), B200ID_E_UNSUPPORTED,
                   "gp_batch_eval_shared: n = %d is beyond the shared-memory kernel", n);
    if (!val) nv = 0;
    const size_t need = b200id_gp_shared_workspace_bytes(n_shapes, n, nv, y_b != nullptr);
    B200ID_REQUIRE(ws && ws_bytes >= need, B200ID_E_WORKSPACE, "gp_batch_eval_shared: workspace %zu < %zu bytes", ws_bytes, need);
    double* tab = (double*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    int rc = b200id_gp_shape_table_launch(kernel_id, shape_theta, n_shapes, X, n, Xv, nv, y_b ? 2 : 1, tab, (cudaStream_t)stream);
    if (rc) return rc;
    return b200id_gp_batch_tiled_launch(kernel_id, nullptr, theta, n_candidates, X, y_a, n, noise, metric, Xv, yv_a, nv,
                                        y_mean_a, y_scale_a, metric_out_a, nullptr, (cudaStream_t)stream,
                                        y_b, yv_b, y_mean_b, y_scale_b, metric_out_b, tab, shape_of);
}

namespace b200id {
__global__ void argmin_pack_kernel(const double* __restrict__ pend, int k, const int64_t* __restrict__ gmap,
                                   double* __restrict__ mine) {
    const int j = threadIdx.x;
    if (j >= k) return;
    const int64_t l = (int64_t)pend[2 * j + 1];
    mine[2 * j] = pend[2 * j];
    mine[2 * j + 1] = l >= 0 ? (gmap ? (double)gmap[l] : (double)l) : -1.0;
}
// same rule as b200id.dist.pick_first_strict_min: m < best, or m == best with a smaller global index
__global__ void argmin_pick_kernel(const double* __restrict__ tab, int ws, int k, const double* __restrict__ full_theta,
                                   double* __restrict__ out) {
    const int j = threadIdx.x;
    if (j >= k) return;
    double best = INFINITY, gsel = -1.0;
    for (int r = 0; r < ws; ++r) {
        const double m = tab[((size_t)r * k + j) * 2], g = tab[((size_t)r * k + j) * 2 + 1];
        if (g < 0.0 || !(m == m)) continue;
        if (m < best || (m == best && g < gsel)) { best = m; gsel = g; }
    }
    const int64_t row = gsel >= 0.0 ? (int64_t)gsel : 0;
    for (int p = 0; p < B200ID_MAX_PARAMS; ++p) out[5 * j + p] = full_theta[row * B200ID_MAX_PARAMS + p];
    out[5 * j + 3] = gsel;
    out[5 * j + 4] = best;
}
}  // namespace b200id

extern "C" int b200id_argmin_pack(const double* pend, int n_searches, const int64_t* global_index_map, double* mine_out,
                                  void* stream) {
    B200ID_REQUIRE(pend && mine_out && n_searches > 0 && n_searches <= 256, B200ID_E_ARG, "argmin_pack: bad argument");
    b200id::argmin_pack_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pend, n_searches, global_index_map, mine_out);
    B200ID_LAUNCH_CHECK("argmin_pack_kernel");
    return B200ID_OK;
}
extern "C" int b200id_argmin_pick(const double* gathered, int n_ranks, int n_searches, const double* full_theta, double* out,
                                  void* stream) {
    B200ID_REQUIRE(gathered && full_theta && out && n_ranks > 0 && n_searches > 0 && n_searches <= 256, B200ID_E_ARG,
                   "argmin_pick: bad argument");
    b200id::argmin_pick_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(gathered, n_ranks, n_searches, full_theta, out);
    B200ID_LAUNCH_CHECK("argmin_pick_kernel");
    return B200ID_OK;
}

extern "C" int b200id_first_strict_min(const double* metric, int64_t n_candidates, double* best_out, int64_t* idx_out,
                                       void* ws, size_t ws_bytes, void* stream) {
    (void)ws; (void)ws_bytes;
    B200ID_REQUIRE(metric && best_out && idx_out && n_candidates > 0, B200ID_E_ARG, "first_strict_min: bad argument");
    first_strict_min_kernel<<<1, ARGMIN_THREADS, 0, (cudaStream_t)stream>>>(metric, n_candidates, best_out, idx_out);
    B200ID_LAUNCH_CHECK("first_strict_min_kernel");
    return B200ID_OK;
}

extern "C" int b200id_first_strict_min_segments(const double* metric, const int64_t* offsets, int n_segments,
                                                double* pend_out, void* stream) {
    B200ID_REQUIRE(metric && offsets && pend_out && n_segments > 0 && n_segments <= 65535, B200ID_E_ARG,
                   "first_strict_min_segments: bad argument");
    first_strict_min_segments_kernel<<<n_segments, ARGMIN_THREADS, 0, (cudaStream_t)stream>>>(metric, offsets, pend_out);
    B200ID_LAUNCH_CHECK("first_strict_min_segments_kernel");
    return B200ID_OK;
}

extern "C" size_t b200id_gp_fit_workspace_bytes(int n) {
    if (n > GP_NMAX_SMEM) return b200id_gp_large_workspace_bytes(n, 0) + align_up((size_t)n * n * sizeof(double), 256) + 256;
    return n > 0 ? align_up((size_t)(n + 2) * (n + 2) * sizeof(double), 256) : 0;   // covers gp_fit and gp_fit_pinv
}

extern "C" int b200id_gp_fit(int kernel_id, const double* theta, const double* X, const double* y, int n,
                             double diag_add, double* alpha_out, double* Kinv_out, int* info_out,
                             void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(theta && X && y && alpha_out && info_out && ws, B200ID_E_ARG, "gp_fit: null pointer");
    B200ID_REQUIRE(kernel_id >= 0 && kernel_id < B200ID_K_COUNT, B200ID_E_ARG, "gp_fit: unknown kernel id %d", kernel_id);
    B200ID_REQUIRE(n > 0 && n <= GP_LARGE_NMAX, B200ID_E_UNSUPPORTED, "gp_fit: n=%d outside (0, %d]", n, GP_LARGE_NMAX);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (n > GP_NMAX_SMEM)
        return b200id_gp_fit_large(kernel_id, theta, X, y, n, diag_add, alpha_out, Kinv_out, info_out, ws, ws_bytes, st);
    B200ID_REQUIRE(ws_bytes >= (size_t)n * n * sizeof(double), B200ID_E_WORKSPACE, "gp_fit: workspace too small");
    if (n <= 127 && !force_generic_gp_path()) {
        // blocked kernel writes alpha + dense L (workspace); K_inv (optional) by one warp per column
        rc = b200id_gp_fit_tiled_launch(kernel_id, theta, X, y, n, diag_add, alpha_out, (double*)ws, info_out, st);
        if (rc) return rc;
        if (Kinv_out) {
            const size_t smem = ((size_t)n * (n | 1) + (size_t)KINV_WARPS * n) * sizeof(double);
            rc = set_smem_attr((const void*)gp_kinv_kernel, smem);
            if (rc) return rc;
            gp_kinv_kernel<<<(n + KINV_WARPS - 1) / KINV_WARPS, KINV_WARPS * 32, smem, st>>>((const double*)ws, n, info_out, Kinv_out);
            B200ID_LAUNCH_CHECK("gp_kinv_kernel");
        }
        return B200ID_OK;
    }
    B200ID_REQUIRE(Kinv_out, B200ID_E_ARG, "gp_fit: the generic path needs Kinv_out");
    const size_t smem = gp_smem_bytes(n);
    rc = set_smem_attr((const void*)gp_fit_kernel, smem);
    if (rc) return rc;
    gp_fit_kernel<<<1, GP_THREADS, smem, st>>>(kernel_id, theta, X, y, n, diag_add, alpha_out, (double*)ws, Kinv_out, info_out);
    B200ID_LAUNCH_CHECK("gp_fit_kernel");
    return B200ID_OK;
}

// Several final fits that share X and the kernel family (the Re and Im models of one identification) as ONE
// launch: theta [B][B200ID_MAX_PARAMS], y [B][n] -> alpha [B][n], info [B]; K_inv is not produced (the
// posterior mean needs only alpha).  n <= 127 runs one CTA per problem; other sizes fit one after the other.
extern "C" int b200id_gp_fit_batch(int kernel_id, const double* theta, const double* X, const double* y, int n,
                                   int n_problems, double diag_add, double* alpha_out, int* info_out,
                                   void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(theta && X && y && alpha_out && info_out && ws, B200ID_E_ARG, "gp_fit_batch: null pointer");
    B200ID_REQUIRE(kernel_id >= 0 && kernel_id < B200ID_K_COUNT, B200ID_E_ARG, "gp_fit_batch: unknown kernel id %d", kernel_id);
    B200ID_REQUIRE(n > 0 && n_problems > 0 && n_problems <= 64, B200ID_E_ARG, "gp_fit_batch: bad sizes n=%d B=%d", n, n_problems);
    if (n <= 127 && !force_generic_gp_path()) {
        B200ID_REQUIRE(ws_bytes >= (size_t)n_problems * n * n * sizeof(double), B200ID_E_WORKSPACE,
                       "gp_fit_batch: workspace %zu < %zu", ws_bytes, (size_t)n_problems * n * n * sizeof(double));
        return b200id_gp_fit_tiled_launch(kernel_id, theta, X, y, n, diag_add, alpha_out, (double*)ws, info_out,
                                          (cudaStream_t)stream, n_problems);
    }
    B200ID_REQUIRE(n > GP_NMAX_SMEM, B200ID_E_UNSUPPORTED, "gp_fit_batch: 128 <= n <= 160 needs K_inv storage; use b200id_gp_fit");
    for (int b = 0; b < n_problems; ++b) {
        int rc = b200id_gp_fit(kernel_id, theta + (size_t)b * B200ID_MAX_PARAMS, X, y + (size_t)b * n, n, diag_add,
                               alpha_out + (size_t)b * n, nullptr, info_out + b, ws, ws_bytes, stream);
        if (rc) return rc;
    }
    return B200ID_OK;
}

extern "C" int b200id_gp_predict(int kernel_id, const double* theta, const double* X, int n, const double* alpha,
                                 const double* Kinv, const double* Xs, int m, double* mean_out, double* std_out,
                                 void* stream) {
    B200ID_REQUIRE(theta && X && alpha && Xs && mean_out, B200ID_E_ARG, "gp_predict: null pointer");
    B200ID_REQUIRE(!std_out || Kinv, B200ID_E_ARG, "gp_predict: std requested without K_inv");
    B200ID_REQUIRE(kernel_id >= 0 && kernel_id < B200ID_K_COUNT, B200ID_E_ARG, "gp_predict: unknown kernel id %d", kernel_id);
    B200ID_REQUIRE(n > 0 && m > 0 && n <= 16384, B200ID_E_ARG, "gp_predict: bad sizes n=%d m=%d", n, m);
    const size_t smem = ((size_t)n + 64) * sizeof(double);
    int rc = set_smem_attr((const void*)gp_predict_kernel, smem);
    if (rc) return rc;
    int grid = m < sm_count() * 4 ? m : sm_count() * 4;
    gp_predict_kernel<<<grid, GP_THREADS, smem, (cudaStream_t)stream>>>(kernel_id, theta, X, n, alpha, Kinv, Xs, m, mean_out, std_out);
    B200ID_LAUNCH_CHECK("gp_predict_kernel");
    return B200ID_OK;
}

// Host evaluation of the covariance functions, for CPU-side unit tests of the numerics only.
extern "C" void b200id_selftest_gram_host(int kernel_id, const double* theta, const double* X1, int n1,
                                          const double* X2, int n2, double* K_out) {
    const KernelParams kp = make_kernel_params(kernel_id, theta);
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) K_out[(size_t)i * n2 + j] = kernel_entry(kp, X1[i], X2[j]);
}
