// Channels-last companions of the convolution path: BatchNorm2d on [P, C] matrices (P = B*H*W pixels, C channels
// contiguous), per-channel sums (bias gradients) and the NCHW <-> NHWC converters used at the Linear-layer boundaries.
//   BatchNorm2d after LeakyReLU in every Conv2D / TConv2D block            model/layer.py:20-26, 39-46
// Same arithmetic as the NCHW kernels in pgv_norm.cu (fp32 partial sums flushed into fp64, biased variance for the
// normalisation, unbiased for the running estimate); every thread owns 4 consecutive channels (float4 accesses) and a
// block covers 256 / (C/4) rows at a time, so global accesses are fully coalesced.  `round_out` rounds the result to
// TF32 (round-to-nearest) because its consumers are the cp.async-fed tensor-core kernels (pgv_conv_cl.cu).
#include <algorithm>

#include "pgv_common.cuh"
#include "pgv_tc.cuh"

namespace pgv {

struct ClMap { int tcb, rl_count; };            // float4 columns per block (power of two <= 256), row lanes per block
static ClMap cl_map(int C) {
    const int tc = C / 4;
    int tcb = 1;
    while (tcb < tc && tcb < 256) tcb <<= 1;
    return {tcb, 256 / tcb};
}
// Rows each thread walks before the grid is capped (pgv_debug_set_bn_rows_per_lane).  Swept on the captured training step at B = 160
// (tools/gpu_sweep_knob.py): 1: 6.74, 2: 6.78, 4: 6.64, 8: 6.57, 16: 6.49, 32: 6.59, 64: 7.03 ms - small grids win on the layers with few pixels
// because every CTA ends in 2C fp64 atomics on the same addresses and a sub-wave tail.
static int g_cl_rows_per_lane = 16;
static int cl_grid_rows(size_t P, int rl_count) {
    const size_t per = static_cast<size_t>(rl_count) * static_cast<size_t>(g_cl_rows_per_lane);
    const size_t want = (P + per - 1) / per;
    // cap: 4 blocks per SM, measured faster for the whole step than 8 (more blocks only add atomics and tail effects)
    return static_cast<int>(std::max<size_t>(1, std::min<size_t>(want, 148 * 4)));
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float rnd(float v, int round_out) { return round_out ? to_tf32_rna(v) : v; }

// Reduces, per channel, up to two quantities over the rows of a [P, C] matrix into ws (double[2*C], zero on entry):
//   MODE 0: sum(x), sum(x^2)      MODE 1: sum(dy), sum(dy * (x - mean) * rstd)       MODE 2: sum(x) only
template <int MODE>
__global__ void __launch_bounds__(256) cl_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                        double* __restrict__ ws, size_t P, int C, int tcb) {
    extern __shared__ double red[];                                    // [256][8]
    const int cg = blockIdx.y * tcb + (threadIdx.x % tcb), rl = threadIdx.x / tcb, RL = 256 / tcb;
    const bool col_ok = cg * 4 < C;
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (col_ok) {
        float4 mean = make_float4(0, 0, 0, 0), rstd = make_float4(1, 1, 1, 1);
        if (MODE == 1) { mean = ld4(save_mean + 4 * cg); rstd = ld4(save_rstd + 4 * cg); }
        float ps[4] = {0, 0, 0, 0}, pq[4] = {0, 0, 0, 0};
        int n = 0;
#pragma unroll 4
        for (size_t r = static_cast<size_t>(blockIdx.x) * RL + rl; r < P; r += static_cast<size_t>(gridDim.x) * RL) {
            const float4 v = ld4(x + r * C + 4 * cg);
            if (MODE == 0) {
                ps[0] += v.x; ps[1] += v.y; ps[2] += v.z; ps[3] += v.w;
                pq[0] = fmaf(v.x, v.x, pq[0]); pq[1] = fmaf(v.y, v.y, pq[1]); pq[2] = fmaf(v.z, v.z, pq[2]); pq[3] = fmaf(v.w, v.w, pq[3]);
            } else if (MODE == 1) {
                const float4 d = ld4(dy + r * C + 4 * cg);
                ps[0] += d.x; ps[1] += d.y; ps[2] += d.z; ps[3] += d.w;
                pq[0] = fmaf(d.x, (v.x - mean.x) * rstd.x, pq[0]); pq[1] = fmaf(d.y, (v.y - mean.y) * rstd.y, pq[1]);
                pq[2] = fmaf(d.z, (v.z - mean.z) * rstd.z, pq[2]); pq[3] = fmaf(d.w, (v.w - mean.w) * rstd.w, pq[3]);
            } else {
                ps[0] += v.x; ps[1] += v.y; ps[2] += v.z; ps[3] += v.w;
            }
            if (++n == 64) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { s[e] += ps[e]; q[e] += pq[e]; ps[e] = pq[e] = 0.0f; }
                n = 0;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[e] += ps[e]; q[e] += pq[e]; }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { red[threadIdx.x * 8 + e] = s[e]; red[threadIdx.x * 8 + 4 + e] = q[e]; }
    __syncthreads();
    if (rl == 0 && col_ok) {
        for (int o = 1; o < RL; ++o)
#pragma unroll
            for (int e = 0; e < 4; ++e) { s[e] += red[(o * tcb + threadIdx.x) * 8 + e]; q[e] += red[(o * tcb + threadIdx.x) * 8 + 4 + e]; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            atomicAdd(ws + 2 * (4 * cg + e), s[e]);
            if (MODE != 2) atomicAdd(ws + 2 * (4 * cg + e) + 1, q[e]);
        }
    }
}

// y = (x - mean) * rstd * gamma + beta from the sums in ws; block row 0 publishes mean / rstd and updates the running stats.
__global__ void __launch_bounds__(256) cl_bn_apply_kernel(const float* __restrict__ x, const double* __restrict__ ws,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                          float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                                          float eps, size_t P, int C, int tcb, int round_out) {
    const int cg = blockIdx.y * tcb + (threadIdx.x % tcb), rl = threadIdx.x / tcb, RL = 256 / tcb;
    if (cg * 4 >= C) return;
    const double n = static_cast<double>(P);
    float g[4], sh[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = 4 * cg + e;
        const double mean = ws[2 * c] / n;
        double var = ws[2 * c + 1] / n - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))), fmean = static_cast<float>(mean);
        if (blockIdx.x == 0 && rl == 0) {
            save_mean[c] = fmean;
            save_rstd[c] = rstd;
            if (running_mean != nullptr) {
                const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
                running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * fmean;
                running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
            }
        }
        g[e] = gamma[c] * rstd;
        sh[e] = beta[c] - fmean * g[e];
    }
#pragma unroll 4
    for (size_t r = static_cast<size_t>(blockIdx.x) * RL + rl; r < P; r += static_cast<size_t>(gridDim.x) * RL) {
        const float4 v = ld4(x + r * C + 4 * cg);
        *reinterpret_cast<float4*>(y + r * C + 4 * cg) =
            make_float4(rnd(fmaf(v.x, g[0], sh[0]), round_out), rnd(fmaf(v.y, g[1], sh[1]), round_out),
                        rnd(fmaf(v.z, g[2], sh[2]), round_out), rnd(fmaf(v.w, g[3], sh[3]), round_out));
    }
}

__global__ void __launch_bounds__(256) cl_bn_eval_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ rm,
                                                         const float* __restrict__ rv, float* __restrict__ y, float eps, size_t P, int C, int tcb,
                                                         int round_out) {
    const int cg = blockIdx.y * tcb + (threadIdx.x % tcb), rl = threadIdx.x / tcb, RL = 256 / tcb;
    if (cg * 4 >= C) return;
    float g[4], sh[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = 4 * cg + e;
        g[e] = gamma[c] / sqrtf(rv[c] + eps);
        sh[e] = beta[c] - rm[c] * g[e];
    }
#pragma unroll 4
    for (size_t r = static_cast<size_t>(blockIdx.x) * RL + rl; r < P; r += static_cast<size_t>(gridDim.x) * RL) {
        const float4 v = ld4(x + r * C + 4 * cg);
        *reinterpret_cast<float4*>(y + r * C + 4 * cg) =
            make_float4(rnd(fmaf(v.x, g[0], sh[0]), round_out), rnd(fmaf(v.y, g[1], sh[1]), round_out),
                        rnd(fmaf(v.z, g[2], sh[2]), round_out), rnd(fmaf(v.w, g[3], sh[3]), round_out));
    }
}

// dx = gamma*rstd*(dy - mean(dy) - xhat*mean(dy*xhat)), then through the LeakyReLU that produced x (x has the sign of the
// pre-activation): multiply by `slope` where x <= 0.
__global__ void __launch_bounds__(256) cl_bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                              const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                                              const float* __restrict__ save_rstd, const double* __restrict__ ws,
                                                              float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              float slope, size_t P, int C, int tcb, int round_out, double* __restrict__ colsum) {
    // colsum (optional, double[C], zero on entry): per-channel sums of the dx written here = the bias gradient of the convolution
    // that feeds this block, for free instead of another pass over dx
    extern __shared__ double red[];                                    // [256][4], only used with colsum
    const int cg = blockIdx.y * tcb + (threadIdx.x % tcb), rl = threadIdx.x / tcb, RL = 256 / tcb;
    const bool col_ok = cg * 4 < C;
    if (!col_ok && colsum == nullptr) return;
    const double n = static_cast<double>(P);
    float m[4], r_[4], gr[4], mdy[4], mdyx[4];
    double cs[4] = {0, 0, 0, 0};
    if (col_ok) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = 4 * cg + e;
        m[e] = save_mean[c]; r_[e] = save_rstd[c];
        mdy[e] = static_cast<float>(ws[2 * c] / n);
        mdyx[e] = static_cast<float>(ws[2 * c + 1] / n);
        gr[e] = gamma[c] * r_[e];
        if (blockIdx.x == 0 && rl == 0) {
            dbeta[c] = static_cast<float>(ws[2 * c]);
            dgamma[c] = static_cast<float>(ws[2 * c + 1]);
        }
    }
    float ps[4] = {0, 0, 0, 0};
    int np = 0;
#pragma unroll 4
    for (size_t r = static_cast<size_t>(blockIdx.x) * RL + rl; r < P; r += static_cast<size_t>(gridDim.x) * RL) {
        const float4 xv = ld4(x + r * C + 4 * cg), dv = ld4(dy + r * C + 4 * cg);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float d = gr[e] * (ds[e] - mdy[e] - (xs[e] - m[e]) * r_[e] * mdyx[e]);
            if (slope >= 0.0f && !(xs[e] > 0.0f)) d *= slope;
            o[e] = rnd(d, round_out);
            ps[e] += o[e];
        }
        *reinterpret_cast<float4*>(dx + r * C + 4 * cg) = make_float4(o[0], o[1], o[2], o[3]);
        if (++np == 64) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { cs[e] += ps[e]; ps[e] = 0.0f; }
            np = 0;
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) cs[e] += ps[e];
    }
    if (colsum == nullptr) return;
#pragma unroll
    for (int e = 0; e < 4; ++e) red[threadIdx.x * 4 + e] = cs[e];
    __syncthreads();
    if (rl == 0 && col_ok) {
        for (int o = 1; o < RL; ++o)
#pragma unroll
            for (int e = 0; e < 4; ++e) cs[e] += red[(o * tcb + threadIdx.x) * 4 + e];
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicAdd(colsum + 4 * cg + e, cs[e]);
    }
}

// ws[2c] = sum(dy), ws[2c + 1] = sum(dy * xhat) from the raw sums (sum(dy), sum(dy * x)) a convolution epilogue accumulated:
// sum(dy * (x - mean) * rstd) = rstd * (sum(dy * x) - mean * sum(dy)), in fp64; ws[2C + c] = 0 (the apply kernel's dx column sums).
__global__ void __launch_bounds__(256) cl_bnbwd_sums_kernel(const double* __restrict__ raw, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, double* __restrict__ ws, int C) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < C) {
        const double s = raw[2 * c];
        ws[2 * c] = s;
        ws[2 * c + 1] = static_cast<double>(rstd[c]) * (raw[2 * c + 1] - static_cast<double>(mean[c]) * s);
        ws[2 * C + c] = 0.0;
    }
}

__global__ void __launch_bounds__(256) cl_sum_finish_kernel(const double* __restrict__ ws, float* __restrict__ out, int C, int stride) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < C) out[c] = static_cast<float>(ws[stride * c]);
}

// dx = dy * (a > 0 ? 1 : slope) on flat arrays (layout-agnostic), optional TF32 rounding of the result
__global__ void __launch_bounds__(256) lrelu_bwd_round_kernel(const float* __restrict__ dy, const float* __restrict__ a, float* __restrict__ dx,
                                                              float slope, size_t n4, int round_out) {
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n4; i += 256ull * gridDim.x) {
        const float4 d = ld4(dy + 4 * i), v = ld4(a + 4 * i);
        *reinterpret_cast<float4*>(dx + 4 * i) =
            make_float4(rnd(v.x > 0.0f ? d.x : d.x * slope, round_out), rnd(v.y > 0.0f ? d.y : d.y * slope, round_out),
                        rnd(v.z > 0.0f ? d.z : d.z * slope, round_out), rnd(v.w > 0.0f ? d.w : d.w * slope, round_out));
    }
}

// [B, C, HW] <-> [B, HW, C] through a 32 x 32 shared-memory tile (both sides coalesced); optional TF32 rounding.
__global__ void __launch_bounds__(256) transpose_inner_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int S, int round_out) {
    // src: [batch][R][S]  ->  dst: [batch][S][R]
    __shared__ float tile[32][33];
    const size_t base = static_cast<size_t>(blockIdx.z) * R * S;
    const int r0 = blockIdx.y * 32, s0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, s = s0 + tx;
        if (r < R && s < S) tile[i][tx] = src[base + static_cast<size_t>(r) * S + s];
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int s = s0 + i, r = r0 + tx;
        if (r < R && s < S) dst[base + static_cast<size_t>(s) * R + r] = rnd(tile[tx][i], round_out);
    }
}

}  // namespace pgv

using namespace pgv;

extern "C" {

int pgv_debug_set_bn_rows_per_lane(int rows) { g_cl_rows_per_lane = rows < 1 ? 1 : rows; return 0; }

int pgv_bn_cl_train_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean, float* save_rstd, float* running_mean,
                        float* running_var, float momentum, float eps, size_t P, int C, int round_out, void* workspace, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && gamma && beta && y && save_mean && save_rstd && workspace, "pgv_bn_cl_train_fwd: NULL argument");
    PGV_CHECK_ARG(P > 0 && C > 0 && C % 4 == 0, "pgv_bn_cl_train_fwd: needs C %% 4 == 0 (C=%d)", C);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* ws = static_cast<double*>(workspace);
    PGV_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, s));
    const ClMap mp = cl_map(C);
    const dim3 grid(cl_grid_rows(P, mp.rl_count), ceil_div(C / 4, mp.tcb));
    cl_reduce_kernel<0><<<grid, 256, 256 * 8 * sizeof(double), s>>>(x, nullptr, nullptr, nullptr, ws, P, C, mp.tcb);
    PGV_LAUNCH_CHECK();
    cl_bn_apply_kernel<<<grid, 256, 0, s>>>(x, ws, gamma, beta, y, save_mean, save_rstd, running_mean, running_var, momentum, eps, P, C, mp.tcb,
                                            round_out);
    PGV_LAUNCH_CHECK();
    return 0;
}

/* Second half of pgv_bn_cl_train_fwd for statistics that a convolution epilogue already accumulated (pgv_conv_cl_fwd_bn / _dgrad_bn):
   sums[2c] = sum, sums[2c + 1] = sum of squares of channel c over the P rows of x. */
int pgv_bn_cl_train_apply(const float* x, const double* sums, const float* gamma, const float* beta, float* y, float* save_mean, float* save_rstd,
                          float* running_mean, float* running_var, float momentum, float eps, size_t P, int C, int round_out,
                          pgv_stream_t stream) {
    PGV_CHECK_ARG(x && sums && gamma && beta && y && save_mean && save_rstd, "pgv_bn_cl_train_apply: NULL argument");
    PGV_CHECK_ARG(P > 0 && C > 0 && C % 4 == 0, "pgv_bn_cl_train_apply: needs C %% 4 == 0 (C=%d)", C);
    const ClMap mp = cl_map(C);
    const dim3 grid(cl_grid_rows(P, mp.rl_count), ceil_div(C / 4, mp.tcb));
    cl_bn_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, sums, gamma, beta, y, save_mean, save_rstd, running_mean, running_var,
                                                                            momentum, eps, P, C, mp.tcb, round_out);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_bn_cl_eval_fwd(const float* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var, float* y,
                       float eps, size_t P, int C, int round_out, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && gamma && beta && running_mean && running_var && y, "pgv_bn_cl_eval_fwd: NULL argument");
    PGV_CHECK_ARG(P > 0 && C > 0 && C % 4 == 0, "pgv_bn_cl_eval_fwd: needs C %% 4 == 0 (C=%d)", C);
    const ClMap mp = cl_map(C);
    const dim3 grid(cl_grid_rows(P, mp.rl_count), ceil_div(C / 4, mp.tcb));
    cl_bn_eval_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, running_mean, running_var, y, eps, P, C, mp.tcb, round_out);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_bn_cl_train_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_rstd, float* dx,
                        float* dgamma, float* dbeta, float* dx_colsum, float lrelu_slope, size_t P, int C, int round_out, const double* raw_sums,
                        void* workspace, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && x && gamma && save_mean && save_rstd && dx && dgamma && dbeta && workspace, "pgv_bn_cl_train_bwd: NULL argument");
    PGV_CHECK_ARG(P > 0 && C > 0 && C % 4 == 0, "pgv_bn_cl_train_bwd: needs C %% 4 == 0 (C=%d)", C);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* ws = static_cast<double*>(workspace);
    const ClMap mp = cl_map(C);
    const dim3 grid(cl_grid_rows(P, mp.rl_count), ceil_div(C / 4, mp.tcb));
    if (raw_sums != nullptr) {
        // the convolution that produced dy already summed it: raw_sums[2c] = sum(dy), raw_sums[2c + 1] = sum(dy * x) (pgv_conv_cl_*_bn, bn_bwd_x)
        cl_bnbwd_sums_kernel<<<ceil_div(C, 256), 256, 0, s>>>(raw_sums, save_mean, save_rstd, ws, C);
    } else {
        PGV_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 3 * C, s));
        cl_reduce_kernel<1><<<grid, 256, 256 * 8 * sizeof(double), s>>>(x, dy, save_mean, save_rstd, ws, P, C, mp.tcb);
    }
    PGV_LAUNCH_CHECK();
    cl_bn_bwd_apply_kernel<<<grid, 256, dx_colsum ? 256 * 4 * sizeof(double) : 0, s>>>(dy, x, gamma, save_mean, save_rstd, ws, dx, dgamma, dbeta,
                                                                                       lrelu_slope, P, C, mp.tcb, round_out,
                                                                                       dx_colsum ? ws + 2 * C : nullptr);
    PGV_LAUNCH_CHECK();
    if (dx_colsum != nullptr) {
        cl_sum_finish_kernel<<<ceil_div(C, 256), 256, 0, s>>>(ws + 2 * C, dx_colsum, C, 1);
        PGV_LAUNCH_CHECK();
    }
    return 0;
}

/* out[c] = sum over the P rows of x[P, C] (bias gradient of a channels-last convolution). */
int pgv_colsum_cl(const float* x, float* out, size_t P, int C, void* workspace, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && out && workspace && P > 0 && C > 0 && C % 4 == 0, "pgv_colsum_cl: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* ws = static_cast<double*>(workspace);
    PGV_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, s));
    const ClMap mp = cl_map(C);
    const dim3 grid(cl_grid_rows(P, mp.rl_count), ceil_div(C / 4, mp.tcb));
    cl_reduce_kernel<2><<<grid, 256, 256 * 8 * sizeof(double), s>>>(x, nullptr, nullptr, nullptr, ws, P, C, mp.tcb);
    PGV_LAUNCH_CHECK();
    cl_sum_finish_kernel<<<ceil_div(C, 256), 256, 0, s>>>(ws, out, C, 2);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_lrelu_bwd_round(const float* dy, const float* a, float* dx, float slope, size_t n, int round_out, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && a && dx && n % 4 == 0, "pgv_lrelu_bwd_round: bad argument (n must be a multiple of 4)");
    if (n == 0) return 0;
    const int grid = static_cast<int>(std::min<size_t>((n / 4 + 255) / 256, 148 * 8));
    lrelu_bwd_round_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, a, dx, slope, n / 4, round_out);
    PGV_LAUNCH_CHECK();
    return 0;
}

/* src [batch, R, S] -> dst [batch, S, R].  NCHW -> NHWC: R = C, S = H*W;  NHWC -> NCHW: R = H*W, S = C. */
int pgv_transpose_inner(const float* src, float* dst, int batch, int R, int S, int round_out, pgv_stream_t stream) {
    PGV_CHECK_ARG(src && dst && batch > 0 && R > 0 && S > 0 && batch <= 65535, "pgv_transpose_inner: bad argument");
    const dim3 grid(ceil_div(S, 32), ceil_div(R, 32), batch);
    PGV_CHECK_ARG(grid.y <= 65535, "pgv_transpose_inner: R too large");
    transpose_inner_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, R, S, round_out);
    PGV_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
