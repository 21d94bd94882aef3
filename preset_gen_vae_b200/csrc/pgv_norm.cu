// Normalisation layers.
//   BatchNorm2d (NCHW)  after LeakyReLU in every Conv2D/TConv2D block          model/layer.py:20-26, 39-46
//   BatchNorm1d ([B,F]) encoder 'lat_in_regularization' (encoder.py:86-87) and the conditioner ResidualBlocks of the
//                       flows (nflows ResidualBlock: BN(eps=1e-3) -> relu -> Linear -> BN -> relu -> dropout -> Linear)
//   Flow BatchNorm      nflows transforms.normalization.BatchNorm between regression-flow couplings (flows.py:87-88):
//                       batch mean / UNBIASED variance, weight = softplus(u) + eps, contributes to log|det J|
// Training-mode semantics are torch's: biased variance for normalisation, unbiased for the running estimate.
#include "pgv_common.cuh"

namespace pgv {

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ BatchNorm2d
// Pass 1: per-channel sum and sum of squares (fp32 per thread, fp64 across threads/blocks).  ws = double[2*C], zeroed.
__global__ void __launch_bounds__(256) bn2d_stats_kernel(const float* __restrict__ x, double* __restrict__ ws, int B, int C, int HW) {
    const int c = blockIdx.x;
    double s = 0.0, q = 0.0;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float* p = x + (static_cast<size_t>(b) * C + c) * HW;
        float ps = 0.0f, pq = 0.0f;
        int n = 0;
        for (int i = threadIdx.x; i < HW; i += 256) {
            const float v = p[i];
            ps += v; pq = fmaf(v, v, pq);
            if (++n == 64) { s += ps; q += pq; ps = pq = 0.0f; n = 0; }
        }
        s += ps; q += pq;
    }
    __shared__ double rs[8], rq[8];
    s = warp_sum(s); q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tq = 0.0;
        for (int i = 0; i < 8; ++i) { ts += rs[i]; tq += rq[i]; }
        atomicAdd(ws + 2 * c, ts);
        atomicAdd(ws + 2 * c + 1, tq);
    }
}

// Pass 2: normalise; block (c, 0) also publishes mean / rstd and updates the running statistics.
__global__ void __launch_bounds__(256) bn2d_apply_kernel(const float* __restrict__ x, const double* __restrict__ ws,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float* __restrict__ y, float* __restrict__ save_mean,
                                                         float* __restrict__ save_rstd, float* __restrict__ running_mean,
                                                         float* __restrict__ running_var, float momentum, float eps, int B, int C, int HW) {
    const int c = blockIdx.x;
    const double n = static_cast<double>(B) * HW;
    const double mean = ws[2 * c] / n;
    double var = ws[2 * c + 1] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float fmean = static_cast<float>(mean);
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        save_mean[c] = fmean;
        save_rstd[c] = rstd;
        if (running_mean != nullptr) {
            const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
            running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * fmean;
            running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
        }
    }
    const float g = gamma[c] * rstd, sh = beta[c] - fmean * g;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t off = (static_cast<size_t>(b) * C + c) * HW;
        for (int i = threadIdx.x; i < HW; i += 256) y[off + i] = fmaf(x[off + i], g, sh);
    }
}

__global__ void __launch_bounds__(256) bn2d_eval_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ rm,
                                                        const float* __restrict__ rv, float* __restrict__ y, float eps, int B, int C, int HW) {
    const int c = blockIdx.x;
    const float g = gamma[c] / sqrtf(rv[c] + eps), sh = beta[c] - rm[c] * g;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t off = (static_cast<size_t>(b) * C + c) * HW;
        for (int i = threadIdx.x; i < HW; i += 256) y[off + i] = fmaf(x[off + i], g, sh);
    }
}

// Backward pass 1: ws[2c] = sum(dy), ws[2c+1] = sum(dy * xhat)
__global__ void __launch_bounds__(256) bn2d_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                              const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                              double* __restrict__ ws, int B, int C, int HW) {
    const int c = blockIdx.x;
    const float m = save_mean[c], r = save_rstd[c];
    double s = 0.0, q = 0.0;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t off = (static_cast<size_t>(b) * C + c) * HW;
        float ps = 0.0f, pq = 0.0f;
        int n = 0;
        for (int i = threadIdx.x; i < HW; i += 256) {
            const float d = dy[off + i];
            ps += d; pq = fmaf(d, (x[off + i] - m) * r, pq);
            if (++n == 64) { s += ps; q += pq; ps = pq = 0.0f; n = 0; }
        }
        s += ps; q += pq;
    }
    __shared__ double rs[8], rq[8];
    s = warp_sum(s); q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tq = 0.0;
        for (int i = 0; i < 8; ++i) { ts += rs[i]; tq += rq[i]; }
        atomicAdd(ws + 2 * c, ts);
        atomicAdd(ws + 2 * c + 1, tq);
    }
}

// Backward pass 2: dx = gamma*rstd*(dy - mean(dy) - xhat*mean(dy*xhat)), then (optionally) through the LeakyReLU that
// produced x: multiply by 1 where x > 0, by `slope` elsewhere (x has the sign of the pre-activation).
__global__ void __launch_bounds__(256) bn2d_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                                             const float* __restrict__ save_rstd, const double* __restrict__ ws,
                                                             float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                             float slope, int B, int C, int HW) {
    const int c = blockIdx.x;
    const double n = static_cast<double>(B) * HW;
    const float m = save_mean[c], r = save_rstd[c];
    const float mean_dy = static_cast<float>(ws[2 * c] / n), mean_dyx = static_cast<float>(ws[2 * c + 1] / n);
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        dbeta[c] = static_cast<float>(ws[2 * c]);
        dgamma[c] = static_cast<float>(ws[2 * c + 1]);
    }
    const float gr = gamma[c] * r;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t off = (static_cast<size_t>(b) * C + c) * HW;
        for (int i = threadIdx.x; i < HW; i += 256) {
            const float xv = x[off + i];
            float d = gr * (dy[off + i] - mean_dy - (xv - m) * r * mean_dyx);
            if (slope >= 0.0f && !(xv > 0.0f)) d *= slope;
            dx[off + i] = d;
        }
    }
}

// y = (a > 0 ? 1 : slope) * dy   for conv blocks without BatchNorm (enc1, enc8)
__global__ void lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ a, float* __restrict__ dx, float slope, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dx[i] = a[i] > 0.0f ? dy[i] : dy[i] * slope;
}

// ------------------------------------------------------------------------------------------------ column-wise ([B, F]) norms
// One block owns 32 consecutive features for the whole batch: blockDim = (32, 8); rows strided by 8.
struct ColReduce {
    __device__ static void sum2(double& a, double& b, double (*sh)[32][2]) {
        sh[threadIdx.y][threadIdx.x][0] = a;
        sh[threadIdx.y][threadIdx.x][1] = b;
        __syncthreads();
        if (threadIdx.y == 0) {
            double ta = 0.0, tb = 0.0;
            for (int i = 0; i < 8; ++i) { ta += sh[i][threadIdx.x][0]; tb += sh[i][threadIdx.x][1]; }
            sh[0][threadIdx.x][0] = ta;
            sh[0][threadIdx.x][1] = tb;
        }
        __syncthreads();
        a = sh[0][threadIdx.x][0];
        b = sh[0][threadIdx.x][1];
        __syncthreads();
    }
};

// BatchNorm1d forward (training): y = [mask *] [relu] (gamma * xhat + beta)
__global__ void __launch_bounds__(256) bn1d_train_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ mask,
                                                             float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                             float* __restrict__ running_mean, float* __restrict__ running_var,
                                                             float momentum, float eps, int relu, int B, int F) {
    __shared__ double sh[8][32][2];
    const int f = blockIdx.x * 32 + threadIdx.x;
    const bool ok = f < F;
    double s = 0.0, q = 0.0;
    if (ok)
        for (int b = threadIdx.y; b < B; b += 8) { const double v = x[static_cast<size_t>(b) * F + f]; s += v; q += v * v; }
    ColReduce::sum2(s, q, sh);
    if (!ok) return;
    const double mean = s / B;
    double var = q / B - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))), fm = static_cast<float>(mean);
    if (threadIdx.y == 0) {
        save_mean[f] = fm;
        save_rstd[f] = rstd;
        if (running_mean != nullptr) {
            const double unbiased = B > 1 ? var * B / (B - 1.0) : var;
            running_mean[f] = (1.0f - momentum) * running_mean[f] + momentum * fm;
            running_var[f] = (1.0f - momentum) * running_var[f] + momentum * static_cast<float>(unbiased);
        }
    }
    const float g = gamma[f] * rstd, shf = beta[f] - fm * g;
    for (int b = threadIdx.y; b < B; b += 8) {
        const size_t i = static_cast<size_t>(b) * F + f;
        float v = fmaf(x[i], g, shf);
        if (relu) v = fmaxf(v, 0.0f);
        if (mask != nullptr) v *= mask[i];
        y[i] = v;
    }
}

__global__ void bn1d_eval_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ rm, const float* __restrict__ rv, float* __restrict__ y, float eps,
                                     int relu, size_t n, int F) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(i % F);
        float v = (x[i] - rm[f]) / sqrtf(rv[f] + eps) * gamma[f] + beta[f];
        y[i] = relu ? fmaxf(v, 0.0f) : v;
    }
}

__global__ void __launch_bounds__(256) bn1d_train_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                             const float* __restrict__ mask, float* __restrict__ dx,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta, int relu, int B, int F) {
    __shared__ double sh[8][32][2];
    const int f = blockIdx.x * 32 + threadIdx.x;
    const bool ok = f < F;
    float m = 0.f, r = 0.f, g = 0.f, bt = 0.f;
    if (ok) { m = save_mean[f]; r = save_rstd[f]; g = gamma[f]; bt = beta[f]; }
    auto upstream = [&](size_t i, float xh) {   // gradient w.r.t. the BN output, after mask and relu
        float d = dy[i];
        if (mask != nullptr) d *= mask[i];
        if (relu && !(fmaf(g, xh, bt) > 0.0f)) d = 0.0f;
        return d;
    };
    double s = 0.0, q = 0.0;
    if (ok)
        for (int b = threadIdx.y; b < B; b += 8) {
            const size_t i = static_cast<size_t>(b) * F + f;
            const float xh = (x[i] - m) * r, d = upstream(i, xh);
            s += d; q += static_cast<double>(d) * xh;
        }
    ColReduce::sum2(s, q, sh);
    if (!ok) return;
    if (threadIdx.y == 0) { dbeta[f] = static_cast<float>(s); dgamma[f] = static_cast<float>(q); }
    const float mean_d = static_cast<float>(s / B), mean_dx = static_cast<float>(q / B);
    for (int b = threadIdx.y; b < B; b += 8) {
        const size_t i = static_cast<size_t>(b) * F + f;
        const float xh = (x[i] - m) * r;
        dx[i] = g * r * (upstream(i, xh) - mean_d - xh * mean_dx);
    }
}

// nflows BatchNorm transform, training forward.  ld_out[0] = sum_f (log w_f - 0.5 log(var_f + eps)) (atomicAdd, zeroed by host).
__device__ __forceinline__ float softplus_f(float u) { return u > 20.0f ? u : log1pf(expf(u)); }

__global__ void __launch_bounds__(256) flowbn_train_fwd_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                                               const float* __restrict__ bias, float* __restrict__ y,
                                                               float* __restrict__ save_mean, float* __restrict__ save_var,
                                                               float* __restrict__ running_mean, float* __restrict__ running_var,
                                                               float* __restrict__ ld_out, float momentum, float eps, int B, int F) {
    __shared__ double sh[8][32][2];
    const int f = blockIdx.x * 32 + threadIdx.x;
    const bool ok = f < F;
    double s = 0.0, q = 0.0;
    if (ok)
        for (int b = threadIdx.y; b < B; b += 8) s += x[static_cast<size_t>(b) * F + f];
    ColReduce::sum2(s, q, sh);
    const double mean = s / B;
    q = 0.0;
    double dummy = 0.0;
    if (ok)
        for (int b = threadIdx.y; b < B; b += 8) { const double d = x[static_cast<size_t>(b) * F + f] - mean; q += d * d; }
    ColReduce::sum2(q, dummy, sh);
    float ld = 0.0f;
    if (ok) {
        const double var = q / (B - 1.0);     // torch.var: unbiased
        const float fm = static_cast<float>(mean), fv = static_cast<float>(var);
        const float w = softplus_f(u[f]) + eps, r = 1.0f / sqrtf(fv + eps);
        if (threadIdx.y == 0) {
            save_mean[f] = fm;
            save_var[f] = fv;
            running_mean[f] = running_mean[f] * (1.0f - momentum) + fm * momentum;
            running_var[f] = running_var[f] * (1.0f - momentum) + fv * momentum;
            ld = logf(w) - 0.5f * logf(fv + eps);
        }
        const float bs = bias[f];
        for (int b = threadIdx.y; b < B; b += 8) {
            const size_t i = static_cast<size_t>(b) * F + f;
            y[i] = w * ((x[i] - fm) * r) + bs;
        }
    }
    if (threadIdx.y == 0) {
        ld = warp_sum(ld);
        if (threadIdx.x == 0) atomicAdd(ld_out, ld);
    }
}

// Evaluation forward (running statistics) and inverse.  direction: 0 forward, 1 inverse.
__global__ void flowbn_eval_kernel(const float* __restrict__ x, const float* __restrict__ u, const float* __restrict__ bias,
                                   const float* __restrict__ rm, const float* __restrict__ rv, float* __restrict__ y,
                                   float* __restrict__ ld_out, float eps, int direction, int B, int F) {
    const size_t n = static_cast<size_t>(B) * F;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(i % F);
        const float w = softplus_f(u[f]) + eps, sd = sqrtf(rv[f] + eps);
        y[i] = direction == 0 ? w * ((x[i] - rm[f]) / sd) + bias[f] : sd * ((x[i] - bias[f]) / w) + rm[f];
    }
    if (blockIdx.x == 0) {
        float ld = 0.0f;
        for (int f = threadIdx.x; f < F; f += blockDim.x) ld += logf(softplus_f(u[f]) + eps) - 0.5f * logf(rv[f] + eps);
        __shared__ float red[32];
        ld = warp_sum(ld);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ld;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.0f;
            for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += red[i];
            ld_out[0] = direction == 0 ? t : -t;
        }
    }
}

// Training backward.  g_ld[0] = sum_b dL/dlogdet[b] (every row received the same scalar).
__global__ void __launch_bounds__(256) flowbn_train_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                               const float* __restrict__ u, const float* __restrict__ save_mean,
                                                               const float* __restrict__ save_var, const float* __restrict__ g_ld,
                                                               float* __restrict__ dx, float* __restrict__ du, float* __restrict__ dbias,
                                                               float eps, int B, int F) {
    __shared__ double sh[8][32][2];
    const int f = blockIdx.x * 32 + threadIdx.x;
    const bool ok = f < F;
    float m = 0.f, v = 1.f, w = 1.f, uu = 0.f;
    if (ok) { m = save_mean[f]; v = save_var[f]; uu = u[f]; w = softplus_f(uu) + eps; }
    const float r = 1.0f / sqrtf(v + eps);
    double s = 0.0, q = 0.0;   // sum(dy), sum(dy * (x - m))
    if (ok)
        for (int b = threadIdx.y; b < B; b += 8) {
            const size_t i = static_cast<size_t>(b) * F + f;
            const float d = dy[i];
            s += d; q += static_cast<double>(d) * (x[i] - m);
        }
    ColReduce::sum2(s, q, sh);
    if (!ok) return;
    const float gl = g_ld[0];
    // dL/dw = sum(dy * xhat) + g_ld / w ; dL/dvar = sum(dy*w*(x-m)) * (-0.5 r^3) + g_ld * (-0.5 / (v + eps))
    const float dw = static_cast<float>(q) * r + gl / w;
    const float dvar = static_cast<float>(q) * w * (-0.5f) * r * r * r - 0.5f * gl / (v + eps);
    const float dmean = -static_cast<float>(s) * w * r;      // the dvar term vanishes: sum(x - m) = 0
    if (threadIdx.y == 0) {
        dbias[f] = static_cast<float>(s);
        du[f] = dw / (1.0f + expf(-uu));                      // d softplus / du = sigmoid(u)
    }
    const float c1 = w * r, c2 = dvar * 2.0f / (B - 1.0f), c3 = dmean / B;
    for (int b = threadIdx.y; b < B; b += 8) {
        const size_t i = static_cast<size_t>(b) * F + f;
        dx[i] = dy[i] * c1 + (x[i] - m) * c2 + c3;
    }
}

static inline int grid_for(size_t n, int block = 256, int max_blocks = 148 * 16) {
    size_t b = (n + block - 1) / block;
    return static_cast<int>(b < 1 ? 1 : (b > static_cast<size_t>(max_blocks) ? max_blocks : b));
}

}  // namespace pgv

using namespace pgv;

extern "C" {

int pgv_bn2d_train_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean, float* save_rstd,
                       float* running_mean, float* running_var, float momentum, float eps, int B, int C, int HW, void* workspace,
                       pgv_stream_t stream) {
    PGV_CHECK_ARG(x && gamma && beta && y && save_mean && save_rstd && workspace, "pgv_bn2d_train_fwd: NULL argument");
    PGV_CHECK_ARG(B > 0 && C > 0 && HW > 0, "pgv_bn2d_train_fwd: empty tensor");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* ws = static_cast<double*>(workspace);
    PGV_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, s));
    int split = ceil_div(148 * 4, C);
    if (split > B) split = B;
    if (split < 1) split = 1;
    bn2d_stats_kernel<<<dim3(C, split), 256, 0, s>>>(x, ws, B, C, HW);
    PGV_LAUNCH_CHECK();
    bn2d_apply_kernel<<<dim3(C, split), 256, 0, s>>>(x, ws, gamma, beta, y, save_mean, save_rstd, running_mean, running_var, momentum,
                                                    eps, B, C, HW);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_bn2d_eval_fwd(const float* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float* y, float eps, int B, int C, int HW, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && gamma && beta && running_mean && running_var && y, "pgv_bn2d_eval_fwd: NULL argument");
    int split = ceil_div(148 * 4, C);
    if (split > B) split = B;
    if (split < 1) split = 1;
    bn2d_eval_kernel<<<dim3(C, split), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, running_mean, running_var, y, eps, B, C, HW);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_bn2d_train_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_rstd, float* dx,
                       float* dgamma, float* dbeta, float lrelu_slope, int B, int C, int HW, void* workspace, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && x && gamma && save_mean && save_rstd && dx && dgamma && dbeta && workspace, "pgv_bn2d_train_bwd: NULL argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* ws = static_cast<double*>(workspace);
    PGV_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, s));
    int split = ceil_div(148 * 4, C);
    if (split > B) split = B;
    if (split < 1) split = 1;
    bn2d_bwd_reduce_kernel<<<dim3(C, split), 256, 0, s>>>(dy, x, save_mean, save_rstd, ws, B, C, HW);
    PGV_LAUNCH_CHECK();
    bn2d_bwd_apply_kernel<<<dim3(C, split), 256, 0, s>>>(dy, x, gamma, save_mean, save_rstd, ws, dx, dgamma, dbeta, lrelu_slope, B, C, HW);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_lrelu_bwd(const float* dy, const float* a, float* dx, float slope, size_t n, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && a && dx, "pgv_lrelu_bwd: NULL argument");
    if (n == 0) return 0;
    lrelu_bwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, a, dx, slope, n);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_bn1d_train_fwd(const float* x, const float* gamma, const float* beta, const float* mask, float* y, float* save_mean,
                       float* save_rstd, float* running_mean, float* running_var, float momentum, float eps, int relu, int B, int F,
                       pgv_stream_t stream) {
    PGV_CHECK_ARG(x && gamma && beta && y && save_mean && save_rstd, "pgv_bn1d_train_fwd: NULL argument");
    PGV_CHECK_ARG(B > 0 && F > 0, "pgv_bn1d_train_fwd: empty tensor");
    bn1d_train_fwd_kernel<<<ceil_div(F, 32), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
        x, gamma, beta, mask, y, save_mean, save_rstd, running_mean, running_var, momentum, eps, relu, B, F);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_bn1d_eval_fwd(const float* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float* y, float eps, int relu, int B, int F, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && gamma && beta && running_mean && running_var && y, "pgv_bn1d_eval_fwd: NULL argument");
    const size_t n = static_cast<size_t>(B) * F;
    bn1d_eval_fwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, running_mean, running_var, y, eps, relu, n, F);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_bn1d_train_bwd(const float* dy, const float* x, const float* gamma, const float* beta, const float* save_mean,
                       const float* save_rstd, const float* mask, float* dx, float* dgamma, float* dbeta, int relu, int B, int F,
                       pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && x && gamma && beta && save_mean && save_rstd && dx && dgamma && dbeta, "pgv_bn1d_train_bwd: NULL argument");
    bn1d_train_bwd_kernel<<<ceil_div(F, 32), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
        dy, x, gamma, beta, save_mean, save_rstd, mask, dx, dgamma, dbeta, relu, B, F);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_flowbn_train_fwd(const float* x, const float* unconstrained_weight, const float* bias, float* y, float* save_mean,
                         float* save_var, float* running_mean, float* running_var, float* logdet_scalar, float momentum, float eps,
                         int B, int F, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && unconstrained_weight && bias && y && save_mean && save_var && running_mean && running_var && logdet_scalar,
                  "pgv_flowbn_train_fwd: NULL argument");
    PGV_CHECK_ARG(B > 1 && F > 0, "pgv_flowbn_train_fwd: needs at least 2 rows (unbiased variance)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PGV_CUDA(cudaMemsetAsync(logdet_scalar, 0, sizeof(float), s));
    flowbn_train_fwd_kernel<<<ceil_div(F, 32), dim3(32, 8), 0, s>>>(x, unconstrained_weight, bias, y, save_mean, save_var, running_mean,
                                                                     running_var, logdet_scalar, momentum, eps, B, F);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_flowbn_eval(const float* x, const float* unconstrained_weight, const float* bias, const float* running_mean,
                    const float* running_var, float* y, float* logdet_scalar, float eps, int inverse, int B, int F, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && unconstrained_weight && bias && running_mean && running_var && y && logdet_scalar, "pgv_flowbn_eval: NULL argument");
    flowbn_eval_kernel<<<grid_for(static_cast<size_t>(B) * F), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, unconstrained_weight, bias, running_mean, running_var, y, logdet_scalar, eps, inverse, B, F);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_flowbn_train_bwd(const float* dy, const float* x, const float* unconstrained_weight, const float* save_mean,
                         const float* save_var, const float* grad_logdet_sum, float* dx, float* d_unconstrained_weight, float* dbias,
                         float eps, int B, int F, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && x && unconstrained_weight && save_mean && save_var && grad_logdet_sum && dx && d_unconstrained_weight && dbias,
                  "pgv_flowbn_train_bwd: NULL argument");
    flowbn_train_bwd_kernel<<<ceil_div(F, 32), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
        dy, x, unconstrained_weight, save_mean, save_var, grad_logdet_sum, dx, d_unconstrained_weight, dbias, eps, B, F);
    PGV_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
