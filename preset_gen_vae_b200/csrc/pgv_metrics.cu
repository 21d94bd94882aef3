// Monitoring metrics of the training step, the inference tail, the inverse-flow controls loss and the data-set statistics pass.
//   QuantizedNumericalParamsLoss    model/loss.py:187-261   (train.py:232: every step, under no_grad)
//   CategoricalParamsAccuracy       model/loss.py:265-315   (train.py:233)
//   learnable -> full presets       data/preset.py:341-369  (PresetsParams.get_full; eval.py inference tail)
//   FlowParamsLoss                  model/loss.py:318-346   (+ utils/probability.py:21-29)
//   check_nan_values                utils/exception.py:13-22 (train.py:245) as a device-side flag
//   spectrogram statistics          data/abstractbasedataset.py:348-391
// All of them are per-row / per-column reductions over [B, 610]-sized tensors or streaming reductions over spectrograms: one block per
// output, fixed-order tree reductions (no atomics), results stay in device memory.
#include <float.h>

#include <algorithm>

#include "pgv_common.cuh"

namespace pgv {

enum { PM_NUM_AS_NUM = 0, PM_NUM_AS_CAT = 1, PM_CAT_AS_NUM = 2, PM_CAT_AS_CAT = 3, PM_NONE = 4 };

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T t = 0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += red[i];
    __syncthreads();
    return t;                      // valid in thread 0
}

// first index of the maximum of v[0..n) (torch.argmax semantics for ties)
__device__ __forceinline__ int argmax_first(const float* __restrict__ v, int n) {
    int best = 0;
    float bv = v[0];
    for (int i = 1; i < n; ++i)
        if (v[i] > bv) { bv = v[i]; best = i; }
    return best;
}

// One block per monitored VST parameter p.  partial[p] = sum over rows of the error term (numerical parameters: |d| or d^2 of the
// quantised values) or the number of rows whose class matches (categorical parameters).
__global__ void __launch_bounds__(128) preset_metrics_kernel(const float* __restrict__ v_out, const float* __restrict__ v_in, int B, int L,
                                                             const int* __restrict__ kind, const int* __restrict__ col,
                                                             const int* __restrict__ len, const int* __restrict__ card, int l1,
                                                             float* __restrict__ partial) {
    __shared__ float red[4];
    const int p = blockIdx.x, k = kind[p], c = col[p], n = len[p];
    const float cm1 = static_cast<float>(card[p]) - 1.0f;
    float acc = 0.0f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float* ro = v_out + static_cast<size_t>(b) * L + c;
        const float* ri = v_in + static_cast<size_t>(b) * L + c;
        if (k == PM_NUM_AS_NUM) {
            float o = ro[0];
            if (card[p] > 0) o = rintf(o * cm1) / cm1;                        // torch.round: half to even
            const float d = o - ri[0];
            acc += l1 ? fabsf(d) : d * d;
        } else if (k == PM_NUM_AS_CAT) {
            const float nm1 = static_cast<float>(n) - 1.0f;
            const float d = static_cast<float>(argmax_first(ro, n)) / nm1 - static_cast<float>(argmax_first(ri, n)) / nm1;
            acc += l1 ? fabsf(d) : d * d;
        } else if (k == PM_CAT_AS_NUM) {
            acc += (static_cast<int>(rintf(ri[0] * cm1)) == static_cast<int>(rintf(ro[0] * cm1))) ? 1.0f : 0.0f;
        } else if (k == PM_CAT_AS_CAT) {
            acc += (argmax_first(ro, n) == argmax_first(ri, n)) ? 1.0f : 0.0f;
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partial[p] = acc;
}

// out[0] = numerical loss (mean over B x n_numerical), out[1] = mean of the per-parameter accuracies (x acc_scale), out[2] = number
// of numerical, out[3] = number of categorical parameters; acc[p] = accuracy of categorical parameter p (x acc_scale), -1 for others.
__global__ void __launch_bounds__(128) preset_metrics_finish_kernel(const float* __restrict__ partial, const int* __restrict__ kind, int P, int B,
                                                                    float acc_scale, float* __restrict__ out, float* __restrict__ acc) {
    __shared__ double red[4];
    double num = 0.0, cat = 0.0;
    int n_num = 0, n_cat = 0;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const int k = kind[p];
        if (k == PM_NUM_AS_NUM || k == PM_NUM_AS_CAT) { num += partial[p]; ++n_num; if (acc) acc[p] = -1.0f; }
        else if (k == PM_CAT_AS_NUM || k == PM_CAT_AS_CAT) {
            const double a = static_cast<double>(partial[p]) / B * acc_scale;
            cat += a; ++n_cat;
            if (acc) acc[p] = static_cast<float>(a);
        } else if (acc) acc[p] = -1.0f;
    }
    num = block_sum(num, red);
    cat = block_sum(cat, red);
    const double nn = block_sum(static_cast<double>(n_num), red), nc = block_sum(static_cast<double>(n_cat), red);
    if (threadIdx.x == 0) {
        out[0] = nn > 0 ? static_cast<float>(num / (nn * B)) : 0.0f;
        out[1] = nc > 0 ? static_cast<float>(cat / nc) : 0.0f;
        out[2] = static_cast<float>(nn);
        out[3] = static_cast<float>(nc);
    }
}

// full[b, p]: kind NONE -> fill[p] (default value or -0.1), numerical -> copy, categorical -> argmax / (cardinal - 1)
__global__ void __launch_bounds__(256) learnable_to_full_kernel(const float* __restrict__ v, int B, int L, const int* __restrict__ kind,
                                                                const int* __restrict__ col, const int* __restrict__ len,
                                                                const int* __restrict__ card, const float* __restrict__ fill, int P,
                                                                float* __restrict__ full) {
    const size_t total = static_cast<size_t>(B) * P;
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < total; i += 256ull * gridDim.x) {
        const int b = static_cast<int>(i / P), p = static_cast<int>(i % P), k = kind[p];
        float r;
        if (k == PM_NONE) r = fill[p];
        else if (k == PM_NUM_AS_NUM || k == PM_CAT_AS_NUM) r = v[static_cast<size_t>(b) * L + col[p]];
        else r = static_cast<float>(argmax_first(v + static_cast<size_t>(b) * L + col[p], len[p])) / (static_cast<float>(card[p]) - 1.0f);
        full[i] = r;
    }
}

// FlowParamsLoss: loss = -mean_b( log N(z0; mu, exp(lv)) + ld_t[b] + ld_u[b] ) / divisor.  One block per row, then a one-block finish.
constexpr float LOG_2PI = 1.8378770664093453f;
__global__ void __launch_bounds__(128) flow_params_rows_kernel(const float* __restrict__ ml, const float* __restrict__ z0,
                                                               const float* __restrict__ ld_t, const float* __restrict__ ld_u, int B, int D,
                                                               float* __restrict__ rows) {
    __shared__ float red[4];
    const int b = blockIdx.x;
    const float* mu = ml + static_cast<size_t>(b) * 2 * D;
    const float* lv = mu + D;
    const float* z = z0 + static_cast<size_t>(b) * D;
    float acc = 0.0f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float df = z[d] - mu[d];
        acc += lv[d] + df * df / expf(lv[d]);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) rows[b] = -0.5f * (D * LOG_2PI + acc) + ld_t[b] + ld_u[b];
}
__global__ void __launch_bounds__(256) mean_scale_kernel(const float* __restrict__ rows, int B, float scale, float* __restrict__ out) {
    __shared__ double red[8];
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) acc += rows[b];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = static_cast<float>(acc / B * scale);
}
__global__ void __launch_bounds__(256) flow_params_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ ml,
                                                              const float* __restrict__ z0, int B, int D, float scale, float* __restrict__ dml,
                                                              float* __restrict__ dz0, float* __restrict__ dld) {
    const float g = gout[0] * scale / B;                 // d loss / d row value
    const size_t n = static_cast<size_t>(B) * D;
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += 256ull * gridDim.x) {
        const size_t b = i / D, d = i % D, im = (b * 2) * D + d, il = im + D;
        const float iv = expf(-ml[il]), df = z0[i] - ml[im];
        dz0[i] = g * (-df * iv);
        dml[im] = g * (df * iv);
        dml[il] = g * (-0.5f * (1.0f - df * df * iv));
        if (d == 0) dld[b] = g;
    }
}

// flags[0] |= 1 << i when vals[i][0] is NaN (vals: device pointers to scalars, passed by value)
struct NanArgs { const float* v[8]; int n; };
__global__ void nan_flag_kernel(NanArgs a, int* __restrict__ flags) {
    int f = 0;
    for (int i = 0; i < a.n; ++i)
        if (a.v[i] != nullptr && isnan(a.v[i][0])) f |= 1 << i;
    if (f) atomicOr(flags, f);
}

// Per-spectrogram statistics: stats[i] = (min, max, mean, unbiased variance) of x[i, 0..n).  One block per spectrogram, two passes.
__global__ void __launch_bounds__(256) spec_stats_kernel(const float* __restrict__ x, size_t n, float* __restrict__ stats) {
    __shared__ double red[8];
    __shared__ float redf[8];
    __shared__ double s_mean;
    const float* xr = x + static_cast<size_t>(blockIdx.x) * n;
    float mn = FLT_MAX, mx = -FLT_MAX;
    double sum = 0.0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = xr[i];
        mn = fminf(mn, v); mx = fmaxf(mx, v); sum += v;
    }
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) { for (int i = 1; i < 8; ++i) mn = fminf(mn, redf[i]); stats[4 * blockIdx.x] = mn; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) { for (int i = 1; i < 8; ++i) mx = fmaxf(mx, redf[i]); stats[4 * blockIdx.x + 1] = mx; }
    __syncthreads();
    sum = block_sum(sum, red);
    if (threadIdx.x == 0) s_mean = sum / static_cast<double>(n);
    __syncthreads();
    const double mean = s_mean;
    double sq = 0.0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) { const double d = xr[i] - mean; sq += d * d; }
    sq = block_sum(sq, red);
    if (threadIdx.x == 0) {
        stats[4 * blockIdx.x + 2] = static_cast<float>(mean);
        stats[4 * blockIdx.x + 3] = static_cast<float>(n > 1 ? sq / static_cast<double>(n - 1) : 0.0);
    }
}
// dataset[0..3] = (min of mins, max of maxes, mean of means, sqrt(mean of variances))   (abstractbasedataset.py:357-360)
__global__ void __launch_bounds__(256) spec_stats_finish_kernel(const float* __restrict__ stats, int N, float* __restrict__ dataset) {
    __shared__ double red[8];
    __shared__ float redf[8];
    float mn = FLT_MAX, mx = -FLT_MAX;
    double sm = 0.0, sv = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        mn = fminf(mn, stats[4 * i]); mx = fmaxf(mx, stats[4 * i + 1]); sm += stats[4 * i + 2]; sv += stats[4 * i + 3];
    }
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) { for (int i = 1; i < 8; ++i) mn = fminf(mn, redf[i]); dataset[0] = mn; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) { for (int i = 1; i < 8; ++i) mx = fmaxf(mx, redf[i]); dataset[1] = mx; }
    __syncthreads();
    sm = block_sum(sm, red);
    sv = block_sum(sv, red);
    if (threadIdx.x == 0) { dataset[2] = static_cast<float>(sm / N); dataset[3] = static_cast<float>(sqrt(sv / N)); }
}

// Backward of the INVERSE direction of the affine coupling (x_t = (y_t - t) / s, logdet -= sum log s), given the result x:
//   dy_t = dx_t / s,  dt = -dx_t / s,  du = (-dx_t * x_t / s - dld / s) * sg (1 - sg);  identity columns pass through.
__global__ void __launch_bounds__(128) coupling_inv_bwd_kernel(const float* __restrict__ dx_out, const float* __restrict__ dld,
                                                               const float* __restrict__ x_out, const float* __restrict__ params,
                                                               const int* __restrict__ id_idx, const int* __restrict__ tr_idx,
                                                               float* __restrict__ dy, float* __restrict__ dparams, int B, int D, int n_id, int n_t) {
    const int row = blockIdx.x, tid = threadIdx.x;
    const size_t ro = static_cast<size_t>(row) * D;
    const float* pr = params + static_cast<size_t>(row) * 2 * n_t;
    float* dpr = dparams + static_cast<size_t>(row) * 2 * n_t;
    const float gl = dld != nullptr ? dld[row] : 0.0f;
    for (int j = tid; j < n_id; j += 128) dy[ro + id_idx[j]] = dx_out[ro + id_idx[j]];
    for (int j = tid; j < n_t; j += 128) {
        const int c = tr_idx[j];
        const float sg = 1.0f / (1.0f + expf(-(pr[n_t + j] + 2.0f))), s = sg + 1e-3f, g = dx_out[ro + c] / s;
        dy[ro + c] = g;
        dpr[j] = -g;
        dpr[n_t + j] = (-g * x_out[ro + c] - gl / s) * sg * (1.0f - sg);
    }
}

// One prefetch.global.L2 per 128-byte line: pulls a parameter range into the 126 MB L2 ahead of a chain of small latency-bound kernels
// that would otherwise each start with a cold HBM round trip (the 72 conditioner layers of a flow read 15 MB of weights once each).
__global__ void __launch_bounds__(256) l2_prefetch_kernel(const char* __restrict__ p, size_t lines) {
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < lines; i += 256ull * gridDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 128));
}

}  // namespace pgv

using namespace pgv;

extern "C" {

int pgv_preset_metrics(const float* v_out, const float* v_in, int B, int L, const int* kind, const int* col, const int* len, const int* card,
                       int P, int l1, float acc_scale, float* partial, float* out4, float* acc, pgv_stream_t stream) {
    PGV_CHECK_ARG(v_out && v_in && kind && col && len && card && partial && out4 && B > 0 && L > 0 && P > 0, "pgv_preset_metrics: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    preset_metrics_kernel<<<P, 128, 0, s>>>(v_out, v_in, B, L, kind, col, len, card, l1, partial);
    PGV_LAUNCH_CHECK();
    preset_metrics_finish_kernel<<<1, 128, 0, s>>>(partial, kind, P, B, acc_scale, out4, acc);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_learnable_to_full(const float* v, int B, int L, const int* kind, const int* col, const int* len, const int* card, const float* fill,
                          int P, float* full, pgv_stream_t stream) {
    PGV_CHECK_ARG(v && kind && col && len && card && fill && full && B > 0 && L > 0 && P > 0, "pgv_learnable_to_full: bad argument");
    const size_t total = static_cast<size_t>(B) * P;
    learnable_to_full_kernel<<<static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        v, B, L, kind, col, len, card, fill, P, full);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_flow_params_loss_fwd(const float* mu_logvar, const float* z0, const float* logdet_t, const float* logdet_u, int B, int D, float divisor,
                             float* loss_out, float* rows_ws, pgv_stream_t stream) {
    PGV_CHECK_ARG(mu_logvar && z0 && logdet_t && logdet_u && loss_out && rows_ws && B > 0 && D > 0 && divisor != 0.0f,
                  "pgv_flow_params_loss_fwd: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    flow_params_rows_kernel<<<B, 128, 0, s>>>(mu_logvar, z0, logdet_t, logdet_u, B, D, rows_ws);
    PGV_LAUNCH_CHECK();
    mean_scale_kernel<<<1, 256, 0, s>>>(rows_ws, B, -1.0f / divisor, loss_out);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_flow_params_loss_bwd(const float* grad_out, const float* mu_logvar, const float* z0, int B, int D, float divisor, float* d_mu_logvar,
                             float* dz0, float* dlogdet, pgv_stream_t stream) {
    PGV_CHECK_ARG(grad_out && mu_logvar && z0 && d_mu_logvar && dz0 && dlogdet && B > 0 && D > 0, "pgv_flow_params_loss_bwd: bad argument");
    const size_t n = static_cast<size_t>(B) * D;
    flow_params_bwd_kernel<<<static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        grad_out, mu_logvar, z0, B, D, -1.0f / divisor, d_mu_logvar, dz0, dlogdet);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_nan_flags(const float* s0, const float* s1, const float* s2, const float* s3, const float* s4, int* flags, pgv_stream_t stream) {
    PGV_CHECK_ARG(flags, "pgv_nan_flags: flags is NULL");
    NanArgs a;
    a.v[0] = s0; a.v[1] = s1; a.v[2] = s2; a.v[3] = s3; a.v[4] = s4; a.n = 5;
    nan_flag_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(a, flags);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_spectrogram_stats(const float* x, int N, size_t elems, float* per_item4, float* dataset4, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && per_item4 && N > 0 && elems > 0, "pgv_spectrogram_stats: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    spec_stats_kernel<<<N, 256, 0, s>>>(x, elems, per_item4);
    PGV_LAUNCH_CHECK();
    if (dataset4 != nullptr) {
        spec_stats_finish_kernel<<<1, 256, 0, s>>>(per_item4, N, dataset4);
        PGV_LAUNCH_CHECK();
    }
    return 0;
}

int pgv_l2_prefetch(const void* p, size_t bytes, pgv_stream_t stream) {
    PGV_CHECK_ARG(p != nullptr && bytes > 0, "pgv_l2_prefetch: bad argument");
    const size_t lines = (bytes + 127) / 128;
    l2_prefetch_kernel<<<static_cast<int>(std::min<size_t>((lines + 255) / 256, 148 * 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const char*>(p), lines);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_coupling_inv_bwd(const float* dx_out, const float* dlogdet, const float* x_out, const float* params, const int* id_idx,
                         const int* tr_idx, float* dy, float* dparams, int B, int D, int n_id, int n_t, pgv_stream_t stream) {
    PGV_CHECK_ARG(dx_out && x_out && params && id_idx && tr_idx && dy && dparams && B > 0 && n_id + n_t == D, "pgv_coupling_inv_bwd: bad argument");
    coupling_inv_bwd_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(dx_out, dlogdet, x_out, params, id_idx, tr_idx, dy, dparams, B, D,
                                                                              n_id, n_t);
    PGV_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
