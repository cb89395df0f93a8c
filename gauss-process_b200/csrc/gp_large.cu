// gp_large.cu -- GP fit / candidate metric for N_d beyond one CTA's shared memory (BASELINE config 4,
// N_d = 2048): Gram in global memory (L2 resident), blocked Cholesky (chol_blocked.cu), blocked
// triangular solves, NLL / validation RMSE, and K_inv = L^-T L^-1 as 64x64x64 FP64 tile products.
//
// Reference call sites: src/gpr/gpr_fitting.py:73-100 (fit / predict), :104-123 (NLL),
// src/gpr/grid_search.py:91-133 (candidate metric).  Same failure conventions as the small-n kernels.
#include "common.cuh"
#include "gp_kernels.cuh"

extern "C" size_t b200id_chol_blocked_workspace_bytes(int n);
extern "C" int b200id_chol_blocked(double* A, int n, int lda, int* info_out, void* ws, size_t ws_bytes, void* stream);
// chol_blocked.cu: same factorisation with rows [n, n_rows) of A carried along (right-hand sides as extra rows)
int cholb_factor_rows(double* A, int n, int n_rows, int lda, int* info_out, cudaStream_t st, int batch, size_t a_stride,
                      size_t info_stride);

namespace b200id {

constexpr int GL_THREADS = 1024;
constexpr int GL_BLK = 64;

// lower triangle (incl. diagonal + diag_add) of K(X, X), row-major
template <int ID>
__device__ __forceinline__ void gl_gram_rows(const KernelParams& kp, const double* __restrict__ X, int n,
                                             double diag_add, double* __restrict__ A) {
    const int64_t total = (int64_t)n * n;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / n), j = (int)(e - (int64_t)i * n);
        if (j > i) continue;
        double v = kernel_entry_t<ID>(kp, X[i], X[j]);
        if (i == j) v = B200ID_ADD(v, diag_add);
        A[e] = v;
    }
}
// blockIdx.z = candidate of a lockstep batch: theta [z][B200ID_MAX_PARAMS], diag_add_dev [z], arrays a_stride apart
__global__ void gl_gram_kernel(int kernel_id, const double* __restrict__ theta, const double* __restrict__ X, int n,
                               double diag_add, const double* __restrict__ diag_add_dev, double* __restrict__ A,
                               const double* __restrict__ y, size_t a_stride) {
    const KernelParams kp = make_kernel_params(kernel_id, theta + (size_t)blockIdx.z * B200ID_MAX_PARAMS);
    if (diag_add_dev) diag_add = diag_add_dev[blockIdx.z];
    A += blockIdx.z * a_stride;
    // row n of the array: the right-hand side, which the factorisation turns into z = L^-1 y on the way
    if (y)
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) A[(size_t)n * n + j] = y[j];
    B200ID_DISPATCH_KERNEL(kernel_id, gl_gram_rows, kp, X, n, diag_add, A)
}

// What is left of the solve after the factorisation: z = L^-1 y rode along as row n of the array (chol_blocked.cu), so
// this kernel forms the NLL pieces nll_terms_out[0] = sum log L_ii, [1] = z.z and, when alpha_out is given, the backward
// sweep alpha = L^-T z.  Single CTA per candidate (blockIdx.x, arrays lane_stride doubles apart), blocks of 64; L
// row-major lower, ld = n.
__global__ void __launch_bounds__(GL_THREADS)
gl_solve_kernel(const double* __restrict__ L, int n, const double* __restrict__ z, const int* __restrict__ info,
                double* __restrict__ alpha_out, double* __restrict__ nll_terms_out, double* __restrict__ work /*[n]*/,
                size_t lane_stride /*doubles between the candidates of a batch; 0 = single*/) {
    __shared__ double xb[GL_BLK];
    __shared__ double red[GL_THREADS / 32];
    {
        const size_t off = blockIdx.x * lane_stride;
        L += off; z += off; work += off;
        info += 2 * off;                                            // ints: same byte offset
        if (alpha_out) alpha_out += off;
        if (nll_terms_out) nll_terms_out += off;
    }
    if (info[0] != 0) return;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < n; i += GL_THREADS) work[i] = z[i];
    __syncthreads();
    // ---- NLL pieces: sum log L_ii and z.z
    double lg = 0.0, zz = 0.0;
    for (int i = tid; i < n; i += GL_THREADS) {
        lg += log(L[(size_t)i * n + i]);
        zz = fma(work[i], work[i], zz);
    }
    lg = block_sum(lg, red);
    zz = block_sum(zz, red);
    if (tid == 0 && nll_terms_out) { nll_terms_out[0] = lg; nll_terms_out[1] = zz; }
    if (!alpha_out) return;
    // ---- backward: L^T alpha = z
    for (int kb = ((n - 1) / GL_BLK) * GL_BLK; kb >= 0; kb -= GL_BLK) {
        const int nb = min(GL_BLK, n - kb);
        if (tid < 32) {
            for (int c = nb - 1; c >= 0; --c) {
                double s = 0.0;
                for (int p = c + 1 + lane; p < nb; p += 32) s = fma(L[(size_t)(kb + p) * n + kb + c], xb[p], s);
                s = warp_sum(s);
                if (lane == 0) xb[c] = (work[kb + c] - s) / L[(size_t)(kb + c) * n + kb + c];
                __syncwarp();
            }
        }
        __syncthreads();
        for (int j = tid; j < kb; j += GL_THREADS) {
            double s = 0.0;
            for (int c = 0; c < nb; ++c) s = fma(L[(size_t)(kb + c) * n + j], xb[c], s);
            work[j] -= s;
        }
        if (tid < nb) work[kb + tid] = xb[tid];
        __syncthreads();
    }
    for (int i = tid; i < n; i += GL_THREADS) alpha_out[i] = work[i];
}

// metric from the solve: NLL, or (with pred = K* alpha already formed) validation RMSE
__global__ void gl_metric_kernel(int metric, int n, const int* __restrict__ info, const double* __restrict__ nll_terms,
                                 const double* __restrict__ pred, const double* __restrict__ yv, int nv,
                                 double y_mean, double y_scale, double* __restrict__ out, size_t lane_stride) {
    __shared__ double red[8];
    {
        const size_t off = blockIdx.x * lane_stride;               // candidate of a batch; out is dense
        info += 2 * off; nll_terms += off; pred += off; out += blockIdx.x;
    }
    if (info[0] != 0) {
        if (threadIdx.x == 0) out[0] = B200ID_FAIL_METRIC;
        return;
    }
    if (metric == B200ID_METRIC_NLL) {
        if (threadIdx.x == 0) out[0] = 0.5 * nll_terms[1] + nll_terms[0] + 0.5 * (double)n * 1.8378770664093453;
        return;
    }
    double se = 0.0;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        const double e = yv[v] - (pred[v] * y_scale + y_mean);
        se = fma(e, e, se);
    }
    se = block_sum(se, red);
    if (threadIdx.x == 0) out[0] = sqrt(se / (double)nv);
}

// L^-1 stored TRANSPOSED: T[c][i] = (L^-1)[i][c], i >= c (row c of T contiguous).  One warp per column c.
__global__ void __launch_bounds__(256)
gl_linv_t_kernel(const double* __restrict__ L, int n, const int* __restrict__ info, double* __restrict__ T) {
    if (info[0] != 0) return;
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    double* v = T + (size_t)c * n;
    for (int i = lane; i < c; i += 32) v[i] = 0.0;
    for (int i = c; i < n; ++i) {
        const double* row = L + (size_t)i * n;
        double s = 0.0;
        for (int p = c + lane; p < i; p += 32) s = fma(row[p], v[p], s);
        s = warp_sum(s);
        if (lane == 0) v[i] = (((i == c) ? 1.0 : 0.0) - s) / row[i];
        __syncwarp();
    }
}

// K_inv = T T^T (T = (L^-1)^T, upper-trapezoidal rows): 64x64 output tiles, k from max(i0, j0) to n
constexpr int GL_LDT = GL_BLK + 2;
__global__ void __launch_bounds__(256)
gl_kinv_tiles_kernel(const double* __restrict__ T, int n, const int* __restrict__ info, double* __restrict__ Kinv) {
    extern __shared__ __align__(16) double gl_sm[];
    double* As = gl_sm;
    double* Bs = gl_sm + GL_BLK * GL_LDT;
    if (info[0] != 0) return;
    const int lin = blockIdx.x;
    int ti = (int)((sqrt(8.0 * lin + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= lin) ++ti;
    while (ti * (ti + 1) / 2 > lin) --ti;
    const int tj = lin - ti * (ti + 1) / 2;
    const int i0 = ti * GL_BLK, j0 = tj * GL_BLK;
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    double acc[4][4] = {};
    for (int k0 = i0; k0 < n; k0 += GL_BLK) {           // T[i][k] = 0 for k < i, and i0 >= j0
        __syncthreads();
        for (int e = threadIdx.x; e < GL_BLK * GL_BLK; e += 256) {
            const int r = e / GL_BLK, k = e - r * GL_BLK;
            const bool kin = k0 + k < n;
            As[k * GL_LDT + r] = (kin && i0 + r < n) ? T[(size_t)(i0 + r) * n + k0 + k] : 0.0;
            Bs[k * GL_LDT + r] = (kin && j0 + r < n) ? T[(size_t)(j0 + r) * n + k0 + k] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < GL_BLK; ++k) {
            const double2 a01 = *reinterpret_cast<const double2*>(As + k * GL_LDT + 4 * ty);
            const double2 a23 = *reinterpret_cast<const double2*>(As + k * GL_LDT + 4 * ty + 2);
            const double2 b01 = *reinterpret_cast<const double2*>(Bs + k * GL_LDT + 4 * tx);
            const double2 b23 = *reinterpret_cast<const double2*>(Bs + k * GL_LDT + 4 * tx + 2);
            const double a[4] = {a01.x, a01.y, a23.x, a23.y};
            const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = i0 + 4 * ty + p;
        if (i >= n) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + 4 * tx + q;
            if (j < n && j <= i) {
                Kinv[(size_t)i * n + j] = acc[p][q];
                Kinv[(size_t)j * n + i] = acc[p][q];
            }
        }
    }
}

}  // namespace b200id

using namespace b200id;

// workspace layout: A [(n+1)*n] (row n = right-hand side) | work [n] | nll_terms [2] | pred [nv] | chol ws
size_t b200id_gp_large_workspace_bytes(int n, int nv) {
    return align_up((size_t)(n + 1) * n * sizeof(double), 256) + align_up((size_t)(n + 2 + (nv > 0 ? nv : 0) + 8) * sizeof(double), 256)
           + b200id_chol_blocked_workspace_bytes(n) + 256;
}

struct LargeWs {
    double *A, *work, *terms, *pred;
    void* chol_ws;
    size_t chol_bytes;
};
static LargeWs carve(void* ws, int n, int nv) {
    LargeWs w;
    char* p = (char*)ws;
    w.A = (double*)p; p += align_up((size_t)(n + 1) * n * sizeof(double), 256);
    w.work = (double*)p;
    w.terms = w.work + n;
    w.pred = w.terms + 2;
    p += align_up((size_t)(n + 2 + (nv > 0 ? nv : 0) + 8) * sizeof(double), 256);
    w.chol_ws = p;
    w.chol_bytes = b200id_chol_blocked_workspace_bytes(n);
    return w;
}

// Gram + factorisation of `batch` candidates in lockstep (theta [batch][3], arrays lane_bytes apart); with y the
// right-hand side rides along as row n and comes back as z = L^-1 y (A + n*n)
static int large_factor(int kernel_id, const double* theta, const double* X, const double* y, int n, double diag_add,
                        const LargeWs& w, int* info, cudaStream_t st, const double* diag_add_dev = nullptr, int batch = 1,
                        size_t lane_bytes = 0) {
    int gx = sm_count() * 8 / batch;
    if (gx < sm_count()) gx = sm_count();
    gl_gram_kernel<<<dim3(gx, 1, batch), 256, 0, st>>>(kernel_id, theta, X, n, diag_add, diag_add_dev, w.A, y, lane_bytes / sizeof(double));
    B200ID_LAUNCH_CHECK("gl_gram_kernel");
    return cholb_factor_rows(w.A, n, y ? n + 1 : n, n, info, st, batch, lane_bytes / sizeof(double), lane_bytes / sizeof(int));
}

// final fit for n > 160: alpha, optional K_inv; ws from b200id_gp_large_workspace_bytes(n, 0) (+ n*n for K_inv scratch)
int b200id_gp_fit_large(int kernel_id, const double* theta, const double* X, const double* y, int n, double diag_add,
                        double* alpha_out, double* Kinv_out, int* info_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t base = b200id_gp_large_workspace_bytes(n, 0);
    const size_t need = base + (Kinv_out ? align_up((size_t)n * n * sizeof(double), 256) : 0);
    B200ID_REQUIRE(ws_bytes >= need, B200ID_E_WORKSPACE, "gp_fit(large): workspace %zu < %zu", ws_bytes, need);
    const LargeWs w = carve(ws, n, 0);
    int rc = large_factor(kernel_id, theta, X, y, n, diag_add, w, info_out, st);
    if (rc) return rc;
    gl_solve_kernel<<<1, GL_THREADS, 0, st>>>(w.A, n, w.A + (size_t)n * n, info_out, alpha_out, w.terms, w.work, 0);
    B200ID_LAUNCH_CHECK("gl_solve_kernel");
    if (Kinv_out) {
        double* T = (double*)((char*)ws + base);
        gl_linv_t_kernel<<<(n + 7) / 8, 256, 0, st>>>(w.A, n, info_out, T);
        B200ID_LAUNCH_CHECK("gl_linv_t_kernel");
        const size_t smem = 2 * (size_t)GL_BLK * GL_LDT * sizeof(double);
        rc = check_cuda(cudaFuncSetAttribute(gl_kinv_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(kinv tiles)");
        if (rc) return rc;
        const int tiles = (n + GL_BLK - 1) / GL_BLK;
        gl_kinv_tiles_kernel<<<tiles * (tiles + 1) / 2, 256, smem, st>>>(T, n, info_out, Kinv_out);
        B200ID_LAUNCH_CHECK("gl_kinv_tiles_kernel");
    }
    return B200ID_OK;
}

// declared in gp_batch.cu
extern "C" int b200id_gp_predict(int kernel_id, const double* theta, const double* X, int n, const double* alpha,
                                 const double* Kinv, const double* Xs, int m, double* mean_out, double* std_out, void* stream);

// candidate metrics for n > 160.  One factorisation is a chain of ~2 n / 64 short, mutually dependent launches, i.e.
// latency bound; candidates are independent, so up to GL_BATCH of them advance in LOCKSTEP: every launch of the chain
// carries all of them (blockIdx.z), each with its own workspace slice.  The chain's latency is paid once per batch,
// every launch brings GL_BATCH times the CTAs, and the host issues the ~70 launches once per batch (no internal
// streams, no host synchronisation; capturable).
constexpr int GL_BATCH = 16;

int b200id_gp_batch_eval_large(int kernel_id, const int* kernel_ids_host_or_null, const double* theta, int64_t n_candidates,
                               const double* X, const double* y, int n, double noise, int metric, const double* Xv,
                               const double* yv, int nv, double y_mean, double y_scale, double* metric_out,
                               const double* noise_v, void* ws, size_t ws_bytes, cudaStream_t st) {
    B200ID_REQUIRE(kernel_ids_host_or_null == nullptr, B200ID_E_UNSUPPORTED, "gp_batch_eval(large n): per-candidate kernel ids are not supported");
    const size_t slice = align_up(b200id_gp_large_workspace_bytes(n, nv) + 256, 256);
    B200ID_REQUIRE(ws_bytes >= slice, B200ID_E_WORKSPACE, "gp_batch_eval(large n): workspace %zu < %zu", ws_bytes, slice);
    int lanes = (int)(ws_bytes / slice);
    if (lanes > GL_BATCH) lanes = GL_BATCH;
    const bool val = metric == B200ID_METRIC_VAL_RMSE;
    const LargeWs w = carve(ws, n, nv);                     // slice 0; slice z sits z * slice bytes further
    int* info = (int*)((char*)w.chol_ws + w.chol_bytes);
    const size_t lane_d = slice / sizeof(double);
    for (int64_t c0 = 0; c0 < n_candidates; c0 += lanes) {
        const int b = (int)((n_candidates - c0 < lanes) ? (n_candidates - c0) : lanes);
        const double* th = theta + c0 * B200ID_MAX_PARAMS;
        int rc = large_factor(kernel_id, th, X, y, n, noise, w, info, st, noise_v ? noise_v + c0 : nullptr, b, slice);
        if (rc) return rc;
        // z = L^-1 y came back as row n of each array: only the NLL pieces (and, for the validation metric, the
        // backward sweep) are left; alpha lives in `work` (in place) when the validation metric needs it
        gl_solve_kernel<<<b, GL_THREADS, 0, st>>>(w.A, n, w.A + (size_t)n * n, info, val ? w.work : nullptr, w.terms, w.work, lane_d);
        B200ID_LAUNCH_CHECK("gl_solve_kernel");
        if (val) {
            for (int z = 0; z < b; ++z) {
                rc = b200id_gp_predict(kernel_id, th + (size_t)z * B200ID_MAX_PARAMS, X, n, w.work + z * lane_d, nullptr, Xv, nv,
                                       w.pred + z * lane_d, nullptr, (void*)st);
                if (rc) return rc;
            }
        }
        gl_metric_kernel<<<b, 256, 0, st>>>(metric, n, info, w.terms, w.pred, yv, nv, y_mean, y_scale, metric_out + c0, lane_d);
        B200ID_LAUNCH_CHECK("gl_metric_kernel");
    }
    return B200ID_OK;
}
