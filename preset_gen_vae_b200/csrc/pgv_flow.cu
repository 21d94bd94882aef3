// Latent-space elementwise kernels on [B, D] tensors (D = 610 by default): reparameterisation (model/VAE.py:167-174),
// RealNVP affine coupling (nflows AffineCouplingTransform via model/flows.py:42-90), column gather/scatter around the
// conditioner network, Hardtanh (decoder.py:98, regression.py:22,51-52) and the per-group softmax activation
// (regression.py:47-50).  One warp per row where a row reduction (log|det J|) is needed.
#include "pgv_common.cuh"
#include "pgv_tc.cuh"

namespace pgv {

__device__ __forceinline__ float warp_sum_f(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

static inline int grid1d(size_t n, int block = 256) {
    size_t b = (n + block - 1) / block;
    return static_cast<int>(b < 1 ? 1 : (b > 148u * 16u ? 148u * 16u : b));
}

// z0 = mu + exp(logvar / 2) * eps ; mu_logvar is [B, 2, D] (row 0 = mu, row 1 = log variance)
__global__ void reparam_fwd_kernel(const float* __restrict__ ml, const float* __restrict__ eps, float* __restrict__ z, int B, int D) {
    const size_t n = static_cast<size_t>(B) * D;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t b = i / D, d = i % D;
        const float mu = ml[(b * 2) * D + d], lv = ml[(b * 2 + 1) * D + d];
        z[i] = eps != nullptr ? fmaf(expf(0.5f * lv), eps[i], mu) : mu;
    }
}
// d(mu_logvar) [B,2,D] = (dz, dz * eps * 0.5 * exp(logvar/2)) [+ d_ml_add if given]
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ ml, const float* __restrict__ eps,
                                   const float* __restrict__ add, float* __restrict__ dml, int B, int D) {
    const size_t n = static_cast<size_t>(B) * D;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t b = i / D, d = i % D, im = (b * 2) * D + d, il = (b * 2 + 1) * D + d;
        const float g = dz[i];
        float gm = g, gl = eps != nullptr ? g * eps[i] * 0.5f * expf(0.5f * ml[il]) : 0.0f;
        if (add != nullptr) { gm += add[im]; gl += add[il]; }
        dml[im] = gm;
        dml[il] = gl;
    }
}

__global__ void gather_cols_kernel(const float* __restrict__ x, const int* __restrict__ idx, float* __restrict__ out, int B, int D, int n) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t total = static_cast<size_t>(B) * n;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = x[(i / n) * D + idx[i % n]];
}
// dst[b, idx[j]] += src[b, j]   (idx has no duplicates)
__global__ void scatter_add_cols_kernel(float* __restrict__ dst, const int* __restrict__ idx, const float* __restrict__ src, int B, int D, int n) {
    griddep_launch_dependents();
    griddep_wait();
    const size_t total = static_cast<size_t>(B) * n;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[(i / n) * D + idx[i % n]] += src[i];
}

// Affine coupling.  params[b, j] = shift, params[b, n_t + j] = unconstrained scale; s = sigmoid(u + 2) + 1e-3.
// forward: y_t = x_t * s + t, logdet += sum log s ; inverse: y_t = (x_t - t) / s, logdet -= sum log s.  One 128-thread block per row.
constexpr int CPL_THREADS = 128;
__global__ void __launch_bounds__(CPL_THREADS) coupling_fwd_kernel(const float* __restrict__ x, const float* __restrict__ params,
                                                           const int* __restrict__ id_idx, const int* __restrict__ tr_idx,
                                                           float* __restrict__ y, const float* __restrict__ ld_in, float* __restrict__ ld_out,
                                                           int B, int D, int n_id, int n_t, int inverse) {
    __shared__ float red[CPL_THREADS / 32];
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* xr = x + static_cast<size_t>(row) * D;
    float* yr = y + static_cast<size_t>(row) * D;
    const float* pr = params + static_cast<size_t>(row) * 2 * n_t;
    for (int j = tid; j < n_id; j += CPL_THREADS) yr[id_idx[j]] = xr[id_idx[j]];
    float ld = 0.0f;
    for (int j = tid; j < n_t; j += CPL_THREADS) {
        const float s = sigmoidf_(pr[n_t + j] + 2.0f) + 1e-3f, t = pr[j];
        const int c = tr_idx[j];
        yr[c] = inverse ? (xr[c] - t) / s : fmaf(xr[c], s, t);
        ld += logf(s);
    }
    ld = warp_sum_f(ld);
    if ((tid & 31) == 0) red[tid >> 5] = ld;
    __syncthreads();
    if (tid == 0) {
        float tot = 0.0f;
        for (int w = 0; w < CPL_THREADS / 32; ++w) tot += red[w];
        ld_out[row] = (ld_in != nullptr ? ld_in[row] : 0.0f) + (inverse ? -tot : tot);
    }
}

// Backward of the forward direction.  dx gets the direct paths (identity columns pass through, transformed columns
// times s); dparams feeds the conditioner network, whose input gradient is scatter-added into dx afterwards.
__global__ void __launch_bounds__(CPL_THREADS) coupling_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dld,
                                                           const float* __restrict__ x, const float* __restrict__ params,
                                                           const int* __restrict__ id_idx, const int* __restrict__ tr_idx,
                                                           float* __restrict__ dx, float* __restrict__ dparams, int B, int D, int n_id, int n_t) {
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x, tid = threadIdx.x;
    const size_t ro = static_cast<size_t>(row) * D;
    const float* pr = params + static_cast<size_t>(row) * 2 * n_t;
    float* dpr = dparams + static_cast<size_t>(row) * 2 * n_t;
    const float gl = dld != nullptr ? dld[row] : 0.0f;
    for (int j = tid; j < n_id; j += CPL_THREADS) dx[ro + id_idx[j]] = dy[ro + id_idx[j]];
    for (int j = tid; j < n_t; j += CPL_THREADS) {
        const int c = tr_idx[j];
        const float sg = sigmoidf_(pr[n_t + j] + 2.0f), s = sg + 1e-3f, g = dy[ro + c];
        dx[ro + c] = g * s;
        dpr[j] = g;                                                   // d shift
        dpr[n_t + j] = (g * x[ro + c] + gl / s) * sg * (1.0f - sg);   // d unconstrained scale
    }
}

__global__ void hardtanh_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float lo, float hi, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        y[i] = clamp_nan(x[i], lo, hi);
}
__global__ void hardtanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, float lo, float hi, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dx[i] = (x[i] > lo && x[i] < hi) ? dy[i] : 0.0f;
}

// y = x * m (dropout with a pre-scaled mask); also its own backward.
__global__ void mul_kernel(const float* __restrict__ x, const float* __restrict__ m, float* __restrict__ y, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        y[i] = x[i] * m[i];
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        y[i] = a[i] + b[i];
}

__global__ void add_scalar_kernel(const float* __restrict__ a, const float* __restrict__ s, float* __restrict__ y, size_t n) {
    const float v = s[0];
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        y[i] = (a != nullptr ? a[i] : 0.0f) + v;
}

// PresetActivation with cat_softmax_activation=True (regression.py:47-50): Hardtanh(0,1) on numerical columns,
// softmax inside every categorical group.  kind[c] = -1 for numerical columns, else the group id; warp per row.
__global__ void __launch_bounds__(256) preset_act_softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                     const int* __restrict__ num_cols, int n_num,
                                                                     const int* __restrict__ grp_start, const int* __restrict__ grp_len,
                                                                     int n_grp, int B, int D) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* xr = x + static_cast<size_t>(row) * D;
    float* yr = y + static_cast<size_t>(row) * D;
    for (int j = lane; j < n_num; j += 32) yr[num_cols[j]] = clamp_nan(xr[num_cols[j]], 0.0f, 1.0f);
    for (int g = 0; g < n_grp; ++g) {
        const int s = grp_start[g], n = grp_len[g];
        float mx = -INFINITY;
        for (int j = lane; j < n; j += 32) mx = fmaxf(mx, xr[s + j]);
        mx = warp_max_f(mx);
        float sum = 0.0f;
        for (int j = lane; j < n; j += 32) sum += expf(xr[s + j] - mx);
        sum = warp_sum_f(sum);
        for (int j = lane; j < n; j += 32) yr[s + j] = expf(xr[s + j] - mx) / sum;
    }
}
__global__ void __launch_bounds__(256) preset_act_softmax_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                     const float* __restrict__ y, float* __restrict__ dx,
                                                                     const int* __restrict__ num_cols, int n_num,
                                                                     const int* __restrict__ grp_start, const int* __restrict__ grp_len,
                                                                     int n_grp, int B, int D) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const size_t ro = static_cast<size_t>(row) * D;
    for (int j = lane; j < n_num; j += 32) {
        const int c = num_cols[j];
        dx[ro + c] = (x[ro + c] > 0.0f && x[ro + c] < 1.0f) ? dy[ro + c] : 0.0f;
    }
    for (int g = 0; g < n_grp; ++g) {
        const int s = grp_start[g], n = grp_len[g];
        float dot = 0.0f;
        for (int j = lane; j < n; j += 32) dot += dy[ro + s + j] * y[ro + s + j];
        dot = warp_sum_f(dot);
        for (int j = lane; j < n; j += 32) dx[ro + s + j] = y[ro + s + j] * (dy[ro + s + j] - dot);
    }
}

// out[f] = sum_b x[b, f]   (bias gradients of the dense layers); block (32, 8) owns 32 columns
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int F) {
    __shared__ float sh[8][32];
    const int f = blockIdx.x * 32 + threadIdx.x;
    float s = 0.0f;
    if (f < F)
        for (int b = threadIdx.y; b < B; b += 8) s += x[static_cast<size_t>(b) * F + f];
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && f < F) {
        float t = 0.0f;
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        out[f] = t;
    }
}

}  // namespace pgv

using namespace pgv;

#define PGV_STREAM(s) static_cast<cudaStream_t>(s)

extern "C" {

int pgv_reparam_fwd(const float* mu_logvar, const float* eps, float* z0, int B, int D, pgv_stream_t stream) {
    PGV_CHECK_ARG(mu_logvar && z0 && B > 0 && D > 0, "pgv_reparam_fwd: bad argument");
    reparam_fwd_kernel<<<grid1d(static_cast<size_t>(B) * D), 256, 0, PGV_STREAM(stream)>>>(mu_logvar, eps, z0, B, D);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_reparam_bwd(const float* dz0, const float* mu_logvar, const float* eps, const float* d_mu_logvar_add, float* d_mu_logvar,
                    int B, int D, pgv_stream_t stream) {
    PGV_CHECK_ARG(dz0 && mu_logvar && d_mu_logvar && B > 0 && D > 0, "pgv_reparam_bwd: bad argument");
    reparam_bwd_kernel<<<grid1d(static_cast<size_t>(B) * D), 256, 0, PGV_STREAM(stream)>>>(dz0, mu_logvar, eps, d_mu_logvar_add, d_mu_logvar, B, D);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_gather_cols(const float* x, const int* idx, float* out, int B, int D, int n, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && idx && out && B > 0 && D > 0 && n > 0, "pgv_gather_cols: bad argument");
    PGV_CUDA(launch_pdl(gather_cols_kernel, dim3(grid1d(static_cast<size_t>(B) * n)), dim3(256), 0, PGV_STREAM(stream), x, idx, out, B, D, n));
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_scatter_add_cols(float* dst, const int* idx, const float* src, int B, int D, int n, pgv_stream_t stream) {
    PGV_CHECK_ARG(dst && idx && src && B > 0 && D > 0 && n > 0, "pgv_scatter_add_cols: bad argument");
    PGV_CUDA(launch_pdl(scatter_add_cols_kernel, dim3(grid1d(static_cast<size_t>(B) * n)), dim3(256), 0, PGV_STREAM(stream), dst, idx, src, B, D, n));
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_coupling_fwd(const float* x, const float* params, const int* identity_idx, const int* transform_idx, float* y,
                     const float* logdet_in, float* logdet_out, int B, int D, int n_identity, int n_transform, int inverse,
                     pgv_stream_t stream) {
    PGV_CHECK_ARG(x && params && identity_idx && transform_idx && y && logdet_out, "pgv_coupling_fwd: NULL argument");
    PGV_CHECK_ARG(n_identity + n_transform == D && B > 0, "pgv_coupling_fwd: index lists must partition the %d features", D);
    PGV_CUDA(launch_pdl(coupling_fwd_kernel, dim3(B), dim3(CPL_THREADS), 0, PGV_STREAM(stream), x, params, identity_idx, transform_idx, y,
                        logdet_in, logdet_out, B, D, n_identity, n_transform, inverse));
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_coupling_bwd(const float* dy, const float* dlogdet, const float* x, const float* params, const int* identity_idx,
                     const int* transform_idx, float* dx, float* dparams, int B, int D, int n_identity, int n_transform,
                     pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && x && params && identity_idx && transform_idx && dx && dparams, "pgv_coupling_bwd: NULL argument");
    PGV_CHECK_ARG(n_identity + n_transform == D && B > 0, "pgv_coupling_bwd: index lists must partition the features");
    PGV_CUDA(launch_pdl(coupling_bwd_kernel, dim3(B), dim3(CPL_THREADS), 0, PGV_STREAM(stream), dy, dlogdet, x, params, identity_idx,
                        transform_idx, dx, dparams, B, D, n_identity, n_transform));
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_hardtanh_fwd(const float* x, float* y, float lo, float hi, size_t n, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && y, "pgv_hardtanh_fwd: NULL argument");
    if (n == 0) return 0;
    hardtanh_fwd_kernel<<<grid1d(n), 256, 0, PGV_STREAM(stream)>>>(x, y, lo, hi, n);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_hardtanh_bwd(const float* dy, const float* x, float* dx, float lo, float hi, size_t n, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && x && dx, "pgv_hardtanh_bwd: NULL argument");
    if (n == 0) return 0;
    hardtanh_bwd_kernel<<<grid1d(n), 256, 0, PGV_STREAM(stream)>>>(dy, x, dx, lo, hi, n);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_mul(const float* x, const float* m, float* y, size_t n, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && m && y, "pgv_mul: NULL argument");
    if (n == 0) return 0;
    mul_kernel<<<grid1d(n), 256, 0, PGV_STREAM(stream)>>>(x, m, y, n);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_add(const float* a, const float* b, float* y, size_t n, pgv_stream_t stream) {
    PGV_CHECK_ARG(a && b && y, "pgv_add: NULL argument");
    if (n == 0) return 0;
    add_kernel<<<grid1d(n), 256, 0, PGV_STREAM(stream)>>>(a, b, y, n);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_add_scalar(const float* a, const float* scalar_dev, float* y, size_t n, pgv_stream_t stream) {
    PGV_CHECK_ARG(scalar_dev && y, "pgv_add_scalar: NULL argument");
    if (n == 0) return 0;
    add_scalar_kernel<<<grid1d(n), 256, 0, PGV_STREAM(stream)>>>(a, scalar_dev, y, n);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_preset_act_softmax_fwd(const float* x, float* y, const int* num_cols, int n_num, const int* grp_start, const int* grp_len,
                               int n_grp, int B, int D, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && y && B > 0 && D > 0, "pgv_preset_act_softmax_fwd: bad argument");
    preset_act_softmax_fwd_kernel<<<ceil_div(B, 8), 256, 0, PGV_STREAM(stream)>>>(x, y, num_cols, n_num, grp_start, grp_len, n_grp, B, D);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_preset_act_softmax_bwd(const float* dy, const float* x, const float* y, float* dx, const int* num_cols, int n_num,
                               const int* grp_start, const int* grp_len, int n_grp, int B, int D, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && x && y && dx && B > 0 && D > 0, "pgv_preset_act_softmax_bwd: bad argument");
    preset_act_softmax_bwd_kernel<<<ceil_div(B, 8), 256, 0, PGV_STREAM(stream)>>>(dy, x, y, dx, num_cols, n_num, grp_start, grp_len, n_grp, B, D);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_colsum(const float* x, float* out, int B, int F, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && out && B > 0 && F > 0, "pgv_colsum: bad argument");
    colsum_kernel<<<ceil_div(F, 32), dim3(32, 8), 0, PGV_STREAM(stream)>>>(x, out, B, F);
    PGV_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
