// fir_conv.cu -- K5 Hermitian IDFT -> FIR taps, K6 causal FIR convolution fused with error sums.
//
// Reference call sites: src/fir_model/fir_helpers.py:75-137 (two-sided Hermitian extension + ifft),
// src/fir_model/fir_fitting.py:285-306 and src/fir_model/fir_validation.py:93-136,190
// (np.convolve(u, g, 'full')[:N], then RMSE / FIT / R2 over i >= skip).
//
// K6 layout.  A CTA of 128 threads produces a tile of 2048 consecutive outputs; each thread owns
// 16 CONSECUTIVE outputs and keeps a circular 24-entry window of u in registers, so one block of 8 taps
// costs 8 shared-memory loads of u (+ 8 uniform constant-bank loads of the taps) for 128 DFMA
// (FP64-issue bound, not shared-memory bound).  The u window [i0 - (taps-1), i0 + 2048) is staged in shared memory with one
// padding double per 16 (lane stride 17 doubles, odd) so that the 16-output-strided lane addresses fall in distinct
// banks.  y_pred is optional: the validation metric only needs the three error sums, which are
// folded per CTA and then by a fixed-order tree (deterministic).
#include <stdlib.h>
#include "common.cuh"
#include <mutex>

namespace b200id {

// ------------------------------------------------------------------------------------------ K5
// h[m] = (1/M) [ Re G0 + 2 sum_{k=1}^{nd-1} ( Re G_k cos(2 pi k m / M) - Im G_k sin(2 pi k m / M) ) ]
// which is Re ifft(X)[m] for the Hermitian X of fir_helpers.py:97-107.  One warp per tap; the
// twiddle index (k*m mod M) is exact in integers and sincospi keeps the angle exact too.
__global__ void __launch_bounds__(256)
idft_hermitian_kernel(const double* __restrict__ G, int nd, double* __restrict__ h, int n_taps) {
    const int warps_per_cta = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int M = 2 * nd - 1;
    const double invM = 1.0 / (double)M;
    for (int m = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); m < n_taps; m += gridDim.x * warps_per_cta) {
        double acc = 0.0;
        for (int k = 1 + lane; k < nd; k += 32) {
            const long long idx = ((long long)k * m) % M;
            double s, c;
            sincospi(2.0 * (double)idx * invM, &s, &c);
            acc += G[2 * k] * c - G[2 * k + 1] * s;
        }
        acc = warp_sum(acc);
        if (lane == 0) h[m] = (G[0] + 2.0 * acc) / (double)M;
    }
}

// ------------------------------------------------------------------------------------------ K6
constexpr int FIR_THREADS = 128;
constexpr int FIR_OPT = 16;                         // outputs per thread
constexpr int FIR_TILE = FIR_THREADS * FIR_OPT;     // 2048 outputs per CTA
constexpr int FIR_TAP_BLOCK = 8;
constexpr int FIR_TAP_GROUP = 24;                   // taps per unrolled group: 3 blocks, the window slots repeat with period 24
constexpr int FIR_WIN = 24;                         // circular register window (23 live entries)
constexpr int FIR_MAX_TAPS = 8192;
constexpr int FIR_CONST_TAPS = 4104;                // taps up to this count are served from constant memory

// FIR taps in constant memory: the tap index is warp-uniform, so every DFMA of the inner loop reads
// acc (register), u window (register) and the tap through a uniform register -- two distinct
// register operands, which is what lets the FP64 pipe issue at its full rate (see fp64_math.cuh).
__constant__ double c_taps[FIR_CONST_TAPS];

// The constant bank is one per process: a call stages its taps there only while it OWNS the bank -- the stream that
// used it last keeps it, another stream takes it over once the previous owner's kernel has finished (event query, no
// wait), and otherwise the call runs the shared-memory tap path (same arithmetic, same order of operations).  Staging
// and launch happen under one host mutex, so two host threads cannot interleave them either.
struct ConstTapsOwner {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool init = false;
};
static ConstTapsOwner g_ctaps;

__host__ __device__ inline int fir_pad(int x) { return x + (x >> 4); }
static inline size_t fir_smem_bytes(int taps_padded) {
    return ((size_t)fir_pad(FIR_TILE + taps_padded) + 8 + taps_padded) * sizeof(double);
}

template <bool G_CONST>
__global__ void __launch_bounds__(FIR_THREADS)
fir_eval_kernel(const double* __restrict__ u, const double* __restrict__ y, int64_t n, int64_t lead,
                const double* __restrict__ g, int n_taps, int taps_padded, int64_t skip, double y0,
                double* __restrict__ y_pred, double* __restrict__ cta_sums) {
    extern __shared__ double sm[];
    double* gs = sm;                         // [taps_padded], zero padded
    double* us = gs + taps_padded;           // padded window, logical index x -> fir_pad(x)
    __shared__ double scratch[FIR_THREADS / 32];

    const int tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * FIR_TILE;
    const int hist = taps_padded - 1;        // samples of history in front of the tile
    if (!G_CONST)
        for (int j = tid; j < taps_padded; j += FIR_THREADS) gs[j] = j < n_taps ? g[j] : 0.0;
    // logical window index x in [0, hist + FIR_TILE) maps to sample i0 - hist + x
    for (int x = tid; x < hist + FIR_TILE; x += FIR_THREADS) {
        const int64_t i = i0 - hist + x;
        us[fir_pad(x)] = (i >= -lead && i < n) ? __ldg(u + i) : 0.0;
    }
    __syncthreads();

    // thread owns outputs o = 0..15 at tile positions p + o, p = 16 * tid; window position of
    // u[i - j] for output position q and tap j is  x = hist + q - j.
    const int p = tid * FIR_OPT;
    double acc[FIR_OPT];
#pragma unroll
    for (int o = 0; o < FIR_OPT; ++o) acc[o] = 0.0;

    // Sliding register window.  Relative position r = x - (hist + p) of u[i - j] for output o and tap j is
    // r = o - j.  A block of 8 taps j0..j0+7 touches r in [-j0-7, -j0+15] (23 values); moving to the next
    // block shifts that range down by 8, so only 8 NEW values are loaded per block and the other 15 stay in
    // registers.  Slot of r is (r mod 24): inside a group of 3 blocks (24 taps) every slot index is a
    // compile-time constant, so the window never moves between registers.
    double win[FIR_WIN];
    const int origin = hist + p;             // window coordinate of r = 0
#pragma unroll
    for (int r = 1; r <= 15; ++r) win[r] = us[fir_pad(origin + r)];      // r = 1..15 of block 0
    for (int g0 = 0; g0 < taps_padded; g0 += FIR_TAP_GROUP) {
#pragma unroll
        for (int b = 0; b < FIR_TAP_GROUP / FIR_TAP_BLOCK; ++b) {
            const int j0 = g0 + b * FIR_TAP_BLOCK;
            // new entries r = -j0-7 .. -j0  ->  slots (-8b - 7 + c) mod 24, c = 0..7
#pragma unroll
            for (int c = 0; c < FIR_TAP_BLOCK; ++c)
                win[((-8 * b - 7 + c) % FIR_WIN + FIR_WIN) % FIR_WIN] = us[fir_pad(origin - j0 - 7 + c)];
#pragma unroll
            for (int jj = 0; jj < FIR_TAP_BLOCK; ++jj) {
                const double gj = G_CONST ? c_taps[j0 + jj] : gs[j0 + jj];
#pragma unroll
                for (int o = 0; o < FIR_OPT; ++o)
                    acc[o] = fma(gj, win[((o - 8 * b - jj) % FIR_WIN + FIR_WIN) % FIR_WIN], acc[o]);
            }
        }
    }

    double se = 0.0, sy = 0.0, st = 0.0, cnt = 0.0;
#pragma unroll
    for (int o = 0; o < FIR_OPT; ++o) {
        const int64_t i = i0 + p + o;
        if (i < n) {
            if (y_pred) y_pred[i] = acc[o];
            if (y && i >= skip) {
                const double yi = __ldg(y + i);
                const double e = yi - acc[o];
                const double d0 = yi - y0;
                se = fma(e, e, se);
                sy = fma(yi, yi, sy);
                st = fma(d0, d0, st);
                cnt += 1.0;
            }
        }
    }
    se = block_sum(se, scratch);
    sy = block_sum(sy, scratch);
    st = block_sum(st, scratch);
    cnt = block_sum(cnt, scratch);
    if (tid == 0) {
        double* o = cta_sums + 4 * (size_t)blockIdx.x;
        o[0] = se; o[1] = sy; o[2] = st; o[3] = cnt;
    }
}

// out[c] = sum over CTAs of part[cta][c], c < 4; pairwise tree in a single CTA (fixed order).
__global__ void __launch_bounds__(256) fir_fold_kernel(const double* __restrict__ part, int64_t n_cta, double* __restrict__ out) {
    __shared__ double scratch[8];
    for (int c = 0; c < 4; ++c) {
        double v = 0.0;
        // contiguous chunk per thread, then the block tree: order depends only on n_cta
        const int64_t per = (n_cta + blockDim.x - 1) / blockDim.x;
        const int64_t lo = (int64_t)threadIdx.x * per, hi = min(lo + per, n_cta);
        for (int64_t k = lo; k < hi; ++k) v += part[4 * k + c];
        v = block_sum(v, scratch);
        if (threadIdx.x == 0) out[c] = v;
    }
}

}  // namespace b200id

using namespace b200id;

// fir_fft.cu: overlap-save FFT path
size_t b200id_fir_fft_workspace_bytes();
bool b200id_fir_fft_applicable(int64_t n, int n_taps);
int b200id_fir_eval_fft(const double* u, const double* y, int64_t n, int64_t lead, const double* g, int n_taps,
                        int64_t skip, double y0, double* y_pred_out, double* sums_out, void* ws, size_t ws_bytes,
                        cudaStream_t st);

static int fir_path_override() {          // B200ID_FIR_PATH = auto | direct | fft
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("B200ID_FIR_PATH");
        cached = (e && strcmp(e, "direct") == 0) ? 1 : (e && strcmp(e, "fft") == 0) ? 2 : 0;
    }
    return cached;
}

extern "C" int b200id_idft_hermitian(const double* G_pos, int nd, double* h_out, int n_taps, void* stream) {
    B200ID_REQUIRE(G_pos && h_out, B200ID_E_ARG, "idft_hermitian: null pointer");
    B200ID_REQUIRE(nd >= 2, B200ID_E_ARG, "idft_hermitian: need nd >= 2");
    B200ID_REQUIRE(n_taps > 0 && n_taps <= 2 * nd - 1, B200ID_E_ARG, "idft_hermitian: N=%d exceeds available length M=%d", n_taps, 2 * nd - 1);
    int grid = (n_taps + 7) / 8;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    idft_hermitian_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(G_pos, nd, h_out, n_taps);
    B200ID_LAUNCH_CHECK("idft_hermitian_kernel");
    return B200ID_OK;
}

extern "C" size_t b200id_fir_workspace_bytes(int64_t n) {
    if (n <= 0) return 0;
    const int64_t ctas = (n + FIR_TILE - 1) / FIR_TILE;
    const size_t direct = align_up((size_t)ctas * 4 * sizeof(double), 256);
    const size_t fft = b200id_fir_fft_workspace_bytes();
    return direct > fft ? direct : fft;
}

extern "C" int b200id_fir_eval(const double* u, const double* y, int64_t n, int64_t lead, const double* g,
                               int n_taps, int64_t skip, double y0, double* y_pred_out, double* sums_out,
                               void* ws, size_t ws_bytes, void* stream) {
    B200ID_REQUIRE(u && g && sums_out && ws, B200ID_E_ARG, "fir_eval: null pointer");
    B200ID_REQUIRE(n > 0 && lead >= 0 && skip >= 0, B200ID_E_ARG, "fir_eval: bad sizes");
    B200ID_REQUIRE(n_taps > 0 && n_taps <= FIR_MAX_TAPS, B200ID_E_UNSUPPORTED, "fir_eval: n_taps=%d outside (0, %d]", n_taps, FIR_MAX_TAPS);
    const int ov = fir_path_override();
    if (ov != 1 && ((ov == 2 && n_taps <= 2048) || b200id_fir_fft_applicable(n, n_taps)))
        return b200id_fir_eval_fft(u, y, n, lead, g, n_taps, skip, y0, y_pred_out, sums_out, ws, ws_bytes, (cudaStream_t)stream);
    const int64_t ctas = (n + FIR_TILE - 1) / FIR_TILE;
    B200ID_REQUIRE(ctas < (1LL << 31), B200ID_E_UNSUPPORTED, "fir_eval: record too long for one launch");
    B200ID_REQUIRE(ws_bytes >= (size_t)ctas * 4 * sizeof(double), B200ID_E_WORKSPACE, "fir_eval: workspace too small");
    const int taps_padded = (n_taps + FIR_TAP_GROUP - 1) / FIR_TAP_GROUP * FIR_TAP_GROUP;
    const size_t smem = fir_smem_bytes(taps_padded);
    cudaStream_t st = (cudaStream_t)stream;
    bool use_const = taps_padded <= FIR_CONST_TAPS;
    int rc;
    std::unique_lock<std::mutex> bank(g_ctaps.mu, std::defer_lock);
    if (use_const) {
        bank.lock();
        if (!g_ctaps.init) {
            rc = check_cuda(cudaEventCreateWithFlags(&g_ctaps.done, cudaEventDisableTiming), "cudaEventCreate(c_taps)");
            if (rc) return rc;
            g_ctaps.init = true;
            g_ctaps.stream = st;
        } else if (g_ctaps.stream != st) {
            if (cudaEventQuery(g_ctaps.done) == cudaSuccess) g_ctaps.stream = st;     // previous owner is done: take over
            else { use_const = false; bank.unlock(); (void)cudaGetLastError(); }      // busy on another stream: shared-memory taps
        }
    }
    if (use_const) {
        // stage the (zero padded) taps into the constant bank on the caller's stream
        void* sym = nullptr;
        rc = check_cuda(cudaGetSymbolAddress(&sym, c_taps), "cudaGetSymbolAddress(c_taps)");
        if (rc) return rc;
        if (taps_padded > n_taps) {
            rc = check_cuda(cudaMemsetAsync((double*)sym + n_taps, 0, (size_t)(taps_padded - n_taps) * sizeof(double), st), "cudaMemsetAsync(c_taps)");
            if (rc) return rc;
        }
        rc = check_cuda(cudaMemcpyAsync(sym, g, (size_t)n_taps * sizeof(double), cudaMemcpyDeviceToDevice, st), "cudaMemcpyAsync(c_taps)");
        if (rc) return rc;
        rc = check_cuda(cudaFuncSetAttribute(fir_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(fir)");
        if (rc) return rc;
        fir_eval_kernel<true><<<(unsigned)ctas, FIR_THREADS, smem, st>>>(u, y, n, lead, g, n_taps, taps_padded, skip, y0,
                                                                        y_pred_out, (double*)ws);
        rc = check_cuda(cudaEventRecord(g_ctaps.done, st), "cudaEventRecord(c_taps)");
        if (rc) return rc;
        bank.unlock();
    } else {
        rc = check_cuda(cudaFuncSetAttribute(fir_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(fir)");
        if (rc) return rc;
        fir_eval_kernel<false><<<(unsigned)ctas, FIR_THREADS, smem, st>>>(u, y, n, lead, g, n_taps, taps_padded, skip, y0,
                                                                         y_pred_out, (double*)ws);
    }
    B200ID_LAUNCH_CHECK("fir_eval_kernel");
    fir_fold_kernel<<<1, 256, 0, st>>>((const double*)ws, ctas, sums_out);
    B200ID_LAUNCH_CHECK("fir_fold_kernel");
    return B200ID_OK;
}
