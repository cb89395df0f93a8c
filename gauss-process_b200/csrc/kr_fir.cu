// kr_fir.cu -- kernel-regularised FIR identification in the time domain (reference
// src/fir_model/kernel_regularized.py): regressor summaries G = U^T U, b = U^T y, yy = y^T y of the Toeplitz
// regressor (:84-92, :337-340), the three covariance builders (DC :99-112, second-order stable spline :115-127,
// simulation-induced :130-181) behind the reference's unconstrained parametrisation (:192-217), the marginal
// likelihood through the L x L identities (:224-261: two Choleskys, S = I + Lk^T G Lk / sigma^2, fallback jitter) for a
// BATCH of hyper-parameter points -- the forward-difference points of one L-BFGS-B gradient -- and the posterior mean
// (:264-291).
//
// One CTA of 1024 threads per hyper-parameter point; its four L x L matrices live in the workspace (L2-resident for
// the reference's L = 128 ... 1024), every product and factorisation is blocked through 32 x 32 shared-memory tiles.
// This is latency-bound work on small matrices: the batch is what fills the GPU.
#include "common.cuh"

#include <math.h>

namespace b200id {

constexpr int KR_THREADS = 1024;
constexpr int KR_T = 32;                  // tile edge
constexpr int KR_LDT = KR_T + 1;
constexpr int KR_SUM_CHUNK = 2048;        // samples per CTA of the summaries kernel
enum { KR_DC = 0, KR_SS = 1, KR_SI = 2 };

__host__ __device__ inline int kr_kdim(int kernel) { return kernel == KR_DC ? 3 : kernel == KR_SS ? 2 : 4; }

// ---------------------------------------------------------------------------------------------------------------
// Regressor summaries.  Row n of U is [u[n], u[n-1], ..., u[n-L+1]] with zeros before the record (:84-92), so
//   G[i][j] = sum_n u[n-i] u[n-j],  b[i] = sum_n u[n-i] y[n].
// grid = (lower-triangular 32 x 32 tiles of (i, j) plus one "b" tile per block row, sample chunks); each CTA writes its
// partial tile, kr_sum_fold_kernel adds the chunks in order and mirrors the triangle.
__global__ void __launch_bounds__(KR_THREADS)
kr_summaries_kernel(const double* __restrict__ u, const double* __restrict__ y, int64_t n, int L, int n_tiles_1d,
                    int64_t chunk, double* __restrict__ part /*[chunks][L][L+1]*/) {
    extern __shared__ double kr_sm[];
    // tile index -> (ti, tj): tj <= ti are the G tiles, tj == ti + 1 is the b column of block row ti
    int ti = 0, rem = blockIdx.x;
    while (rem >= ti + 2) { rem -= ti + 2; ++ti; }
    const int tj = rem;
    const bool is_b = tj == ti + 1;
    const int i0 = ti * KR_T, j0 = is_b ? 0 : tj * KR_T;
    const int64_t n0 = (int64_t)blockIdx.y * chunk, n1 = min(n0 + chunk, n);
    const int halo = i0 + KR_T - 1;                                   // oldest sample index needed: n0 - halo
    const int len = (int)(n1 - n0) + halo;
    double* su = kr_sm;                                               // su[k] = u[n0 - halo + k]
    double* sy = kr_sm + (chunk + L + KR_T);
    for (int k = threadIdx.x; k < len; k += KR_THREADS) {
        const int64_t idx = n0 - halo + k;
        su[k] = idx >= 0 ? __ldg(u + idx) : 0.0;
    }
    if (is_b)
        for (int k = threadIdx.x; k < (int)(n1 - n0); k += KR_THREADS) sy[k] = __ldg(y + n0 + k);
    __syncthreads();
    const int li = threadIdx.x >> 5, lj = threadIdx.x & 31;
    const int i = i0 + li, j = j0 + lj;
    double acc = 0.0;
    if (i < L) {
        const double* pi = su + (halo - i);                           // pi[m] = u[n0 + m - i]
        if (is_b) {
            // 32 lanes split the chunk, then a shuffle tree (fixed order)
            for (int m = lj; m < (int)(n1 - n0); m += 32) acc = fma(pi[m], sy[m], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        } else if (j < L) {
            const double* pj = su + (halo - j);
            for (int m = 0; m < (int)(n1 - n0); ++m) acc = fma(pi[m], pj[m], acc);
        }
    }
    double* out = part + (size_t)blockIdx.y * L * (L + 1);
    if (i < L) {
        if (is_b) { if (lj == 0) out[(size_t)i * (L + 1) + L] = acc; }
        else if (j < L) out[(size_t)i * (L + 1) + j] = acc;
    }
}
__global__ void kr_sum_fold_kernel(const double* __restrict__ part, int n_chunks, int L, double* __restrict__ G,
                                   double* __restrict__ b) {
    const int i = blockIdx.x, L1 = L + 1;
    for (int j = threadIdx.x; j <= L; j += blockDim.x) {
        if (j < L && j > i) continue;
        double v = 0.0;
        for (int c = 0; c < n_chunks; ++c) v += part[(size_t)c * L * L1 + (size_t)i * L1 + j];
        if (j == L) b[i] = v;
        else { G[(size_t)i * L + j] = v; G[(size_t)j * L + i] = v; }
    }
}
__global__ void __launch_bounds__(256) kr_yy_kernel(const double* __restrict__ y, int64_t n, int64_t chunk, double* __restrict__ part) {
    __shared__ double scratch[8];
    const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(lo + chunk, n);
    double s = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) { const double v = __ldg(y + i); s = fma(v, v, s); }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void kr_yy_fold_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { double v = 0.0; for (int c = 0; c < n; ++c) v += part[c]; out[0] = v; }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-wide building blocks (all threads of the CTA call them; every one ends with a barrier)

// element (r, c) of a row-major lower-triangular factor stored in a full square buffer
__device__ __forceinline__ double kr_low(const double* A, int ld, int r, int c) { return r >= c ? A[(size_t)r * ld + c] : 0.0; }

enum { KR_NN_LOW = 0, KR_TN_LOW = 1, KR_NT = 2 };
// C(i, j) = epilogue( sum_k A(i, k) B(k, j) ) over i in [i_lo, M), j in [j_lo, j_hi), k in [0, K):
//   KR_NN_LOW : A(i, k) = A[i][k],            B(k, j) = lower(B)[k][j]        (T = G Lk)
//   KR_TN_LOW : A(i, k) = lower(A)[k][i],     B(k, j) = B[k][j]               (Lk^T T)
//   KR_NT     : A(i, k) = A[i][k],            B(k, j) = B[j][k]               (panel update of the factorisation)
// Every global load of a tile is a row segment (coalesced); transposes happen in shared memory.  The operands are
// buffers this same kernel wrote a moment ago: no __restrict__ / read-only loads on them.
template <int FORM, class Epi>
__device__ void kr_gemm(int i_lo, int M, int j_lo, int j_hi, int K, const double* A, int lda, const double* B, int ldb,
                        double* sm, Epi epi) {
    double (*As)[KR_LDT] = reinterpret_cast<double (*)[KR_LDT]>(sm);
    double (*Bs)[KR_LDT] = reinterpret_cast<double (*)[KR_LDT]>(sm + KR_T * KR_LDT);
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    for (int i0 = i_lo; i0 < M; i0 += KR_T)
        for (int j0 = j_lo; j0 < j_hi; j0 += KR_T) {
            double acc = 0.0;
            for (int k0 = 0; k0 < K; k0 += KR_T) {
                if (FORM == KR_NN_LOW) {
                    const int ai = i0 + ty, ak = k0 + tx, bk = k0 + ty, bj = j0 + tx;
                    As[ty][tx] = (ai < M && ak < K) ? A[(size_t)ai * lda + ak] : 0.0;
                    Bs[ty][tx] = (bk < K && bj < j_hi) ? kr_low(B, ldb, bk, bj) : 0.0;
                } else if (FORM == KR_TN_LOW) {
                    const int ak = k0 + ty, ai = i0 + tx, bk = k0 + ty, bj = j0 + tx;      // As[k][i]
                    As[ty][tx] = (ak < K && ai < M) ? kr_low(A, lda, ak, ai) : 0.0;
                    Bs[ty][tx] = (bk < K && bj < j_hi) ? B[(size_t)bk * ldb + bj] : 0.0;
                } else {
                    const int ai = i0 + ty, ak = k0 + tx, bj = j0 + ty, bk = k0 + tx;      // Bs[j][k]
                    As[ty][tx] = (ai < M && ak < K) ? A[(size_t)ai * lda + ak] : 0.0;
                    Bs[ty][tx] = (bj < j_hi && bk < K) ? B[(size_t)bj * ldb + bk] : 0.0;
                }
                __syncthreads();
#pragma unroll 8
                for (int kk = 0; kk < KR_T; ++kk) {
                    const double a = FORM == KR_TN_LOW ? As[kk][ty] : As[ty][kk];
                    const double b = FORM == KR_NT ? Bs[tx][kk] : Bs[kk][tx];
                    acc = fma(a, b, acc);
                }
                __syncthreads();
            }
            const int i = i0 + ty, j = j0 + tx;
            if (i < M && j < j_hi) epi(i, j, acc);
        }
}

// In-place lower Cholesky of the row-major n x n matrix A (upper triangle left as it was), blocked left-looking:
// panel update (kr_gemm), 32 x 32 diagonal block factorised by one warp in shared memory, rows below solved one thread
// a row.  Returns false (for every thread) at the first non-positive or non-finite pivot, as LAPACK's potrf would
// raise through scipy.linalg.cholesky.
__device__ bool kr_cholesky(double* A, int n, int ld, double* sm, int* sm_flag) {
    double (*D)[KR_LDT] = reinterpret_cast<double (*)[KR_LDT]>(sm + 2 * KR_T * KR_LDT);
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    if (threadIdx.x == 0) *sm_flag = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += KR_T) {
        const int nb = min(KR_T, n - c0);
        if (c0 > 0)
            kr_gemm<KR_NT>(c0, n, c0, c0 + nb, c0, A, ld, A, ld, sm,
                           [&](int i, int j, double acc) { if (i >= j) A[(size_t)i * ld + j] -= acc; });
        __syncthreads();
        D[ty][tx] = (ty < nb && tx < nb && ty >= tx) ? A[(size_t)(c0 + ty) * ld + c0 + tx] : 0.0;
        __syncthreads();
        if (ty == 0) {                                                 // warp 0: unblocked factorisation, lane = row
            for (int k = 0; k < nb; ++k) {
                const double p = D[k][k];
                if (!(p > 0.0) || !isfinite(p)) { if (tx == 0) *sm_flag = 1; break; }
                const double d = sqrt(p);
                __syncwarp();
                if (tx == k) D[k][k] = d;
                else if (tx > k && tx < nb) D[tx][k] /= d;
                __syncwarp();
                if (tx > k && tx < nb)
                    for (int c = k + 1; c <= tx; ++c) D[tx][c] -= D[tx][k] * D[c][k];
                __syncwarp();
            }
        }
        __syncthreads();
        if (*sm_flag) return false;
        if (ty < nb && tx < nb && ty >= tx) A[(size_t)(c0 + ty) * ld + c0 + tx] = D[ty][tx];
        // rows below the block: x L_D^T = a  ->  forward substitution along the row
        for (int i = c0 + nb + threadIdx.x; i < n; i += KR_THREADS) {
            double* row = A + (size_t)i * ld + c0;
            for (int c = 0; c < nb; ++c) {
                double v = row[c];
                for (int k = 0; k < c; ++k) v -= row[k] * D[c][k];
                row[c] = v / D[c][c];
            }
        }
        __syncthreads();
    }
    return true;
}

// x = (Ls Ls^T)^{-1} w by one warp (row-oriented forward sweep, row-oriented backward sweep); x in shared memory
__device__ void kr_cho_solve(const double* Ls, int n, int ld, double* x /*in: w, out: x (shared)*/) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int i = 0; i < n; ++i) {                                  // Ls z = w
            const double* row = Ls + (size_t)i * ld;
            double s = 0.0;
            for (int k = lane; k < i; k += 32) s = fma(row[k], x[k], s);
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            __syncwarp();
            if (lane == 0) x[i] = (x[i] - s) / row[i];
            __syncwarp();
        }
        for (int i = n - 1; i >= 0; --i) {                             // Ls^T x = z
            const double* row = Ls + (size_t)i * ld;
            const double xi = x[i] / row[i];
            __syncwarp();
            if (lane == 0) x[i] = xi;
            for (int k = lane; k < i; k += 32) x[k] = fma(-row[k], xi, x[k]);
            __syncwarp();
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Covariance builders behind the reference's unconstrained parameters (:192-217)
__device__ __forceinline__ double kr_sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double kr_clamp(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// K (full, symmetric) into the row-major buffer; tab = 3 L doubles of shared memory
__device__ void kr_build_kernel(int kernel, const double* __restrict__ theta, int L, double* K, double* tab) {
    const int tid = threadIdx.x;
    if (kernel == KR_DC) {
        const double c = exp(theta[0]), lam = kr_sigmoid(theta[1]), rho = tanh(theta[2]);
        const double arho = fabs(rho);
        // lam^(m/2) for m = i + j <= 2L - 2 and |rho|^d with the sign of rho^d (:106-111)
        for (int m = tid; m < 2 * L - 1; m += KR_THREADS) tab[m] = pow(lam, 0.5 * (double)m);
        for (int d = tid; d < L; d += KR_THREADS) tab[2 * L + d] = pow(arho, (double)d) * ((rho < 0.0 && (d & 1)) ? -1.0 : 1.0);
        __syncthreads();
        for (int e = tid; e < L * L; e += KR_THREADS) {
            const int i = e / L, j = e - i * L, d = i > j ? i - j : j - i;
            K[e] = c * (tab[i + j] * tab[2 * L + d]);
        }
    } else if (kernel == KR_SS) {
        const double c = exp(theta[0]);
        const double lam = kr_clamp(kr_sigmoid(theta[1]), 1e-6, 1.0 - 1e-6);
        // (C H)[i][j] = sum_{m=j..i} lam^(m-j) = s[i-j]: the cumulative sums of the powers (:120-126)
        if (tid == 0) { double s = 0.0; for (int d = 0; d < L; ++d) { s += pow(lam, (double)d); tab[d] = s; } }
        __syncthreads();
        for (int e = tid; e < L * L; e += KR_THREADS) {
            const int i = e / L, j = e - i * L;
            if (j > i) continue;
            double acc = 0.0;
            for (int m = 0; m <= j; ++m) acc = fma(tab[i - m], tab[j - m], acc);
            K[e] = c * acc;
            K[(size_t)j * L + i] = c * acc;
        }
    } else {
        const double r = kr_clamp(kr_sigmoid(theta[0]), 1e-4, 1.0 - 1e-6);
        const double omega = kr_clamp(3.14159265358979323846 * kr_sigmoid(theta[1]), 1e-3, 3.14159265358979323846 - 1e-3);
        const double q = exp(theta[2]);
        const double barlam = kr_clamp(kr_sigmoid(theta[3]), 1e-6, 1.0 - 1e-6);
        // first row of A^n by the reference's own recurrence A^n = A^(n-1) A (:146-148): (a, b) = r^n (cos n w, -sin n w)
        if (tid == 0) {
            const double a11 = r * cos(omega), a12 = -(r * sin(omega)), a21 = r * sin(omega), a22 = r * cos(omega);
            double p11 = 1.0, p12 = 0.0, p21 = 0.0, p22 = 1.0;
            for (int n = 0; n < L; ++n) {
                tab[n] = p11; tab[L + n] = p12;                        // v_n = C A^n; g_n = C A^n B = p11
                const double q11 = p11 * a11 + p12 * a21, q12 = p11 * a12 + p12 * a22;
                const double q21 = p21 * a11 + p22 * a21, q22 = p21 * a12 + p22 * a22;
                p11 = q11; p12 = q12; p21 = q21; p22 = q22;
            }
            for (int k = 0; k < L; ++k) tab[2 * L + k] = pow(barlam, (double)k);
        }
        __syncthreads();
        for (int e = tid; e < L * L; e += KR_THREADS) {
            const int t = e / L, s = e - t * L;
            if (s > t) continue;
            double v = q * (tab[t] * tab[s] + tab[L + t] * tab[L + s]);
            double acc = 0.0;
            for (int k = 0; k < s; ++k) acc += tab[2 * L + k] * tab[t - 1 - k] * tab[s - 1 - k];   // m = min(t, s) = s
            v += acc;
            K[e] = v;
            K[(size_t)s * L + t] = v;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// One hyper-parameter point per CTA: nll (:224-261) and, when g_out is given, the posterior mean (:264-291).
// info: 0 ok, 2 = S took the reference's 1e-6 fallback jitter, 1 = K + jitter I not positive definite,
// 3 = S not positive definite even with the fallback (the reference raises LinAlgError in both cases).
struct KrWs { double *K, *T, *S, *Ls; };

__global__ void __launch_bounds__(KR_THREADS)
kr_nll_kernel(int kernel, const double* __restrict__ theta_all, int stride, const double* __restrict__ G,
              const double* __restrict__ bvec, double yy, double n_samples, int L, double jitter, int posterior,
              double* __restrict__ nll_out, int* __restrict__ info_out, double* __restrict__ g_out,
              double* __restrict__ K_out, double* __restrict__ sig2_out, double* __restrict__ ws_all) {
    extern __shared__ double kr_sm[];
    double* tiles = kr_sm;                                            // 3 tiles of KR_T x KR_LDT
    double* vec = kr_sm + 3 * KR_T * KR_LDT;                          // [L]   w, then x
    double* tab = vec + L;                                            // [3 L] kernel tables / scratch
    __shared__ int flag;
    __shared__ double red[KR_THREADS / 32];
    const int tid = threadIdx.x, cand = blockIdx.x;
    const double* theta = theta_all + (size_t)cand * stride;
    const int kdim = kr_kdim(kernel);
    const double sig2 = exp(theta[kdim]);
    const double s2 = fmax(sig2, 1e-30);
    const size_t LL = (size_t)L * L;
    double* Kb = ws_all + (size_t)cand * 4 * LL;
    double* Tb = Kb + LL;
    double* Sb = Tb + LL;
    double* Lsb = Sb + LL;

    kr_build_kernel(kernel, theta, L, Kb, tab);
    if (posterior && K_out)
        for (size_t e = tid; e < LL; e += KR_THREADS) K_out[e] = Kb[e];
    if (posterior == 2) {                                              // covariance only
        if (tid == 0) info_out[cand] = 0;
        return;
    }
    __syncthreads();
    for (int i = tid; i < L; i += KR_THREADS) Kb[(size_t)i * L + i] += jitter;
    __syncthreads();
    if (!kr_cholesky(Kb, L, L, tiles, &flag)) {
        if (tid == 0) { nll_out[cand] = nan(""); info_out[cand] = 1; }
        return;
    }
    // T = G Lk ; S = I + Lk^T T / sigma^2
    kr_gemm<KR_NN_LOW>(0, L, 0, L, L, G, L, Kb, L, tiles, [&](int i, int j, double acc) { Tb[(size_t)i * L + j] = acc; });
    __syncthreads();
    kr_gemm<KR_TN_LOW>(0, L, 0, L, L, Kb, L, Tb, L, tiles,
                       [&](int i, int j, double acc) { Sb[(size_t)i * L + j] = (i == j ? 1.0 : 0.0) + acc / s2; });
    __syncthreads();
    int info = 0;
    for (size_t e = tid; e < LL; e += KR_THREADS) { const int i = (int)(e / L), j = (int)(e - (size_t)i * L); Lsb[e] = Sb[e] + (i == j ? jitter : 0.0); }
    __syncthreads();
    if (!kr_cholesky(Lsb, L, L, tiles, &flag)) {
        if (posterior) {                                               // _posterior_mean_g has no fallback
            if (tid == 0) { nll_out[cand] = nan(""); info_out[cand] = 3; }
            return;
        }
        info = 2;
        for (size_t e = tid; e < LL; e += KR_THREADS) { const int i = (int)(e / L), j = (int)(e - (size_t)i * L); Lsb[e] = Sb[e] + (i == j ? 1e-6 : 0.0); }
        __syncthreads();
        if (!kr_cholesky(Lsb, L, L, tiles, &flag)) {
            if (tid == 0) { nll_out[cand] = nan(""); info_out[cand] = 3; }
            return;
        }
    }
    // w = Lk^T b (thread = column: consecutive threads read consecutive addresses of each row)
    for (int j = tid; j < L; j += KR_THREADS) {
        double acc = 0.0;
        for (int i = j; i < L; ++i) acc = fma(Kb[(size_t)i * L + j], bvec[i], acc);
        vec[j] = acc;
        tab[j] = acc;                                                  // keep w for the quadratic form
    }
    __syncthreads();
    kr_cho_solve(Lsb, L, L, vec);
    // quad = (yy - w.x / s2) / s2 ; logdet S = 2 sum log diag(Ls)
    double wx = 0.0, ld = 0.0;
    for (int i = tid; i < L; i += KR_THREADS) { wx = fma(tab[i], vec[i], wx); ld += log(Lsb[(size_t)i * L + i]); }
    wx = block_sum(wx, red);
    ld = block_sum(ld, red);
    if (tid == 0) {
        const double quad = (yy - wx / s2) / s2;
        nll_out[cand] = 0.5 * (quad + n_samples * log(s2) + 2.0 * ld);
        info_out[cand] = info;
        if (posterior && sig2_out) sig2_out[0] = sig2;
    }
    if (posterior && g_out) {
        // g_hat = Lk x / sigma^2 (warp per row)
        const int warp = tid >> 5, lane = tid & 31;
        for (int i = warp; i < L; i += KR_THREADS / 32) {
            double acc = 0.0;
            for (int k = lane; k <= i; k += 32) acc = fma(Kb[(size_t)i * L + k], vec[k], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) g_out[i] = acc / s2;
        }
    }
}

static size_t kr_nll_smem_bytes(int L) { return (size_t)(3 * KR_T * KR_LDT + 4 * L) * sizeof(double); }

}  // namespace b200id

using namespace b200id;

// sample chunks of the summaries pass: at most 64, fewer when L is large (the partial area is capped at 128 MB)
static int kr_max_chunks(int L) {
    const size_t per = (size_t)L * (L + 1) * sizeof(double);
    size_t m = ((size_t)128 << 20) / per;
    return m < 1 ? 1 : m > 64 ? 64 : (int)m;
}
extern "C" size_t b200id_kr_workspace_bytes(int L, int batch) {
    if (L <= 0 || batch <= 0) return 0;
    const size_t nll = (size_t)batch * 4 * L * L * sizeof(double);
    const size_t sums = (size_t)kr_max_chunks(L) * L * (L + 1) * sizeof(double) + 4096 * sizeof(double);
    return align_up(nll > sums ? nll : sums, 256) + 256;
}

extern "C" int b200id_kr_summaries(const double* u, const double* y, int64_t n, int L, double* G_out, double* b_out,
                                   double* yy_out, void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(u && y && G_out && b_out && yy_out && ws, B200ID_E_ARG, "kr_summaries: null pointer");
    B200ID_REQUIRE(n > 0 && L > 0 && L <= 4096, B200ID_E_ARG, "kr_summaries: need n > 0 and 0 < L <= 4096");
    B200ID_REQUIRE(ws_bytes >= b200id_kr_workspace_bytes(L, 1), B200ID_E_WORKSPACE, "kr_summaries: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    // chunks: KR_SUM_CHUNK samples each, at most kr_max_chunks(L) of them
    const int max_chunks = kr_max_chunks(L);
    int64_t chunk = KR_SUM_CHUNK;
    int n_chunks = (int)((n + chunk - 1) / chunk);
    if (n_chunks > max_chunks) { chunk = (n + max_chunks - 1) / max_chunks; n_chunks = (int)((n + chunk - 1) / chunk); }
    const int nt = (L + KR_T - 1) / KR_T;
    const int tiles = nt * (nt + 1) / 2 + nt;
    const size_t smem = (size_t)(2 * chunk + L + KR_T + 8) * sizeof(double);
    B200ID_REQUIRE(smem <= 220 * 1024, B200ID_E_UNSUPPORTED, "kr_summaries: record too long for one pass at this L (n <= ~8e5 at L = 128; the reference truncates to 1e5 samples)");
    int rc = check_cuda(cudaFuncSetAttribute(kr_summaries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
    if (rc) return rc;
    double* part = (double*)ws;
    kr_summaries_kernel<<<dim3(tiles, n_chunks), KR_THREADS, smem, st>>>(u, y, n, L, nt, chunk, part);
    B200ID_LAUNCH_CHECK("kr_summaries_kernel");
    kr_sum_fold_kernel<<<L, 128, 0, st>>>(part, n_chunks, L, G_out, b_out);
    B200ID_LAUNCH_CHECK("kr_sum_fold_kernel");
    double* ypart = part + (size_t)max_chunks * L * (L + 1);
    int yg = 1024;
    int64_t ychunk = (n + yg - 1) / yg;
    ychunk = (ychunk + 255) / 256 * 256;
    yg = (int)((n + ychunk - 1) / ychunk);
    kr_yy_kernel<<<yg, 256, 0, st>>>(y, n, ychunk, ypart);
    B200ID_LAUNCH_CHECK("kr_yy_kernel");
    kr_yy_fold_kernel<<<1, 32, 0, st>>>(ypart, yg, yy_out);
    B200ID_LAUNCH_CHECK("kr_yy_fold_kernel");
    return B200ID_OK;
}

static int kr_launch(int kernel, const double* theta, int batch, const double* G, const double* b, double yy, double n_samples,
                     int L, double jitter, int posterior, double* nll_out, int* info_out, double* g_out, double* K_out,
                     double* sig2_out, void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(kernel >= 0 && kernel <= 2, B200ID_E_ARG, "kr: kernel must be 0 (dc), 1 (ss) or 2 (si)");
    B200ID_REQUIRE(theta && G && b && nll_out && info_out && ws, B200ID_E_ARG, "kr: null pointer");
    B200ID_REQUIRE(L > 0 && L <= 4096 && batch > 0, B200ID_E_ARG, "kr: need 0 < L <= 4096 and batch > 0");
    B200ID_REQUIRE(ws_bytes >= b200id_kr_workspace_bytes(L, batch), B200ID_E_WORKSPACE, "kr: workspace too small");
    const size_t smem = kr_nll_smem_bytes(L);
    int rc = check_cuda(cudaFuncSetAttribute(kr_nll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
    if (rc) return rc;
    kr_nll_kernel<<<batch, KR_THREADS, smem, (cudaStream_t)stream>>>(kernel, theta, kr_kdim(kernel) + 1, G, b, yy, n_samples, L, jitter,
                                                                      posterior, nll_out, info_out, g_out, K_out, sig2_out, (double*)ws);
    B200ID_LAUNCH_CHECK("kr_nll_kernel");
    return B200ID_OK;
}

extern "C" int b200id_kr_nll_batch(int kernel, const double* theta, int batch, const double* G, const double* b, double yy,
                                   double n_samples, int L, double jitter, double* nll_out, int* info_out, void* ws,
                                   size_t ws_bytes, void* stream) {
    return kr_launch(kernel, theta, batch, G, b, yy, n_samples, L, jitter, 0, nll_out, info_out, nullptr, nullptr, nullptr, ws, ws_bytes, stream);
}

extern "C" int b200id_kr_posterior(int kernel, const double* theta, const double* G, const double* b, int L, double* g_out,
                                   double* K_out, double* sig2_out, int* info_out, void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE((g_out && sig2_out) || (!g_out && K_out), B200ID_E_ARG, "kr_posterior: need g_out and sig2_out, or K_out alone");
    // nll_out is not part of this call: the first double of the workspace tail receives it
    double* scratch = (double*)((char*)ws + align_up((size_t)4 * L * L * sizeof(double), 256));
    return kr_launch(kernel, theta, 1, G, b, 0.0, 0.0, L, 1e-10, g_out ? 1 : 2, scratch, info_out, g_out, K_out, sig2_out, ws, ws_bytes, stream);
}
