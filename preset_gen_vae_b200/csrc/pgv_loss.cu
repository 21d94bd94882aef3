// Training losses (forward value + gradient kernels) and the fused optimizer step.
//   reconstruction  nn.MSELoss('mean') / L2Loss                train.py:103-106, model/loss.py:15-43
//   latent          FlowVAE.latent_loss                        model/VAE.py:183-193 + utils/probability.py:13-29
//   KL              GaussianDkl                                model/loss.py:46-66
//   controls        SynthParamsLoss                            model/loss.py:73-183 (+ data/preset.py:247-283)
//   Adam            torch.optim.Adam(weight_decay=L2-in-gradient)   train.py:165-167
// Every loss writes its scalar to device memory and every backward reads the upstream gradient from device memory,
// so a whole training step can be captured in one CUDA graph without host synchronisation.
#include "pgv_common.cuh"

namespace pgv {

__device__ __forceinline__ double warp_sum_d(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum_ff(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_ff(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum into one double atomic
__device__ __forceinline__ void block_atomic_add(double v, double* dst) {
    __shared__ double red[32];
    v = warp_sum_d(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += red[i];
        atomicAdd(dst, t);
    }
    __syncthreads();
}

static inline int grid1(size_t n, int block = 256) {
    size_t b = (n + block - 1) / block;
    return static_cast<int>(b < 1 ? 1 : (b > 148u * 8u ? 148u * 8u : b));
}

// ------------------------------------------------------------------------------------------------ squared error
__global__ void __launch_bounds__(256) sqerr_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ ws, size_t n) {
    double acc = 0.0;
    float part = 0.0f;
    int cnt = 0;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float d = a[i] - b[i];
        part = fmaf(d, d, part);
        if (++cnt == 64) { acc += part; part = 0.0f; cnt = 0; }
    }
    block_atomic_add(acc + part, ws);
}
__global__ void scale_to_float_kernel(const double* __restrict__ ws, float* __restrict__ out, double scale) { out[0] = static_cast<float>(ws[0] * scale); }
__global__ void sqerr_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gout, float* __restrict__ da,
                                 float scale2, size_t n) {
    const float g = gout[0] * scale2;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        da[i] = g * (a[i] - b[i]);
}

// ------------------------------------------------------------------------------------------------ latent loss
// per row: 0.5*sum(zK^2) - 0.5*sum(lv + (z0-mu)^2 * exp(-lv)) - logdet      (the D*log(2*pi) terms cancel)
__global__ void __launch_bounds__(256) latent_loss_fwd_kernel(const float* __restrict__ ml, const float* __restrict__ z0, const float* __restrict__ zk,
                                                              const float* __restrict__ logdet, double* __restrict__ ws, int B, int D) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    double acc = 0.0;
    if (row < B) {
        const float *mu = ml + static_cast<size_t>(row) * 2 * D, *lv = mu + D;
        const float *a = z0 + static_cast<size_t>(row) * D, *k = zk + static_cast<size_t>(row) * D;
        float s = 0.0f;
        for (int d = lane; d < D; d += 32) {
            const float diff = a[d] - mu[d];
            s += 0.5f * k[d] * k[d] - 0.5f * (lv[d] + diff * diff * expf(-lv[d]));
        }
        s = warp_sum_ff(s);
        if (lane == 0) acc = static_cast<double>(s) - logdet[row];
    }
    block_atomic_add(lane == 0 ? acc : 0.0, ws);
}
__global__ void latent_loss_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ ml, const float* __restrict__ z0,
                                       const float* __restrict__ zk, float* __restrict__ dml, float* __restrict__ dz0, float* __restrict__ dzk,
                                       float* __restrict__ dlogdet, float scale, int B, int D) {
    const float c = gout[0] * scale;
    const size_t n = static_cast<size_t>(B) * D;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t b = i / D, d = i % D, im = (b * 2) * D + d, il = im + D;
        const float diff = z0[i] - ml[im], e = expf(-ml[il]);
        dzk[i] = c * zk[i];
        dz0[i] = -c * diff * e;
        dml[im] = c * diff * e;
        dml[il] = -0.5f * c * (1.0f - diff * diff * e);
        if (d == 0) dlogdet[b] = -c;
    }
}

// ------------------------------------------------------------------------------------------------ KL to N(0, I)
__global__ void __launch_bounds__(256) dkl_fwd_kernel(const float* __restrict__ ml, double* __restrict__ ws, int B, int D) {
    const size_t n = static_cast<size_t>(B) * D;
    double acc = 0.0;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t b = i / D, d = i % D;
        const float mu = ml[(b * 2) * D + d], lv = ml[(b * 2 + 1) * D + d];
        acc += 0.5 * (expf(lv) + mu * mu - lv - 1.0f);
    }
    block_atomic_add(acc, ws);
}
__global__ void dkl_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ ml, float* __restrict__ dml, float scale, int B, int D) {
    const float c = gout[0] * scale;
    const size_t n = static_cast<size_t>(B) * D;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t b = i / D, d = i % D, im = (b * 2) * D + d, il = im + D;
        dml[im] = c * ml[im];
        dml[il] = 0.5f * c * (expf(ml[il]) - 1.0f);
    }
}

// ------------------------------------------------------------------------------------------------ synth parameters loss
// ws (double): [0] numerical squared-error sum, [1 .. G] per-group cross-entropy sums, [1+G .. 2G] per-group useful rows.
// A numerical output / categorical group whose `vol_col` target is < 1e-3 (silent Dexed operator) is ignored for
// that row.  Warp per row.
struct SynthTables {
    const int *num_cols, *num_vol, *grp_start, *grp_len, *grp_vol;
    int n_num, n_grp;
};
__global__ void __launch_bounds__(256) synth_loss_fwd_kernel(const float* __restrict__ vo, const float* __restrict__ vi, SynthTables t,
                                                             double* __restrict__ ws, float inv_temp, int use_softmax, int B, int L) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float *o = vo + static_cast<size_t>(row) * L, *in = vi + static_cast<size_t>(row) * L;
    float num = 0.0f;
    for (int j = lane; j < t.n_num; j += 32) {
        const int vol = t.num_vol[j];
        if (vol >= 0 && in[vol] < 1e-3f) continue;
        const float d = o[t.num_cols[j]] - in[t.num_cols[j]];
        num = fmaf(d, d, num);
    }
    num = warp_sum_ff(num);
    if (lane == 0) atomicAdd(ws, static_cast<double>(num));
    for (int g = 0; g < t.n_grp; ++g) {
        const int vol = t.grp_vol[g];
        if (vol >= 0 && in[vol] < 1e-3f) continue;
        const int s = t.grp_start[g], n = t.grp_len[g];
        float ce = 0.0f;
        if (use_softmax) {
            float mx = -INFINITY;
            for (int j = lane; j < n; j += 32) mx = fmaxf(mx, o[s + j] * inv_temp);
            mx = warp_max_ff(mx);
            float sum = 0.0f;
            for (int j = lane; j < n; j += 32) sum += expf(o[s + j] * inv_temp - mx);
            const float lse = mx + logf(warp_sum_ff(sum));
            for (int j = lane; j < n; j += 32)
                if (in[s + j] != 0.0f) ce -= o[s + j] * inv_temp - lse;
        } else {
            for (int j = lane; j < n; j += 32)
                if (in[s + j] != 0.0f) ce -= logf(o[s + j]);
        }
        ce = warp_sum_ff(ce);
        if (lane == 0) {
            atomicAdd(ws + 1 + g, static_cast<double>(ce));
            atomicAdd(ws + 1 + t.n_grp + g, 1.0);
        }
    }
}
// group_counts (optional): useful rows per group to normalise with instead of this batch's own counts - data-parallel training passes
// (global count) / world so that the mean over ranks equals the loss of the gathered batch (loss.py:172); they replace the counts in
// ws, which the backward kernel reads.
__global__ void synth_loss_finish_kernel(double* __restrict__ ws, float* __restrict__ out, double num_scale, double cat_scale, int n_grp,
                                         const double* __restrict__ group_counts) {
    double cat = 0.0;
    if (group_counts != nullptr)
        for (int g = 0; g < n_grp; ++g) ws[1 + n_grp + g] = group_counts[g];
    for (int g = 0; g < n_grp; ++g) cat += ws[1 + g] / ws[1 + n_grp + g];   // 0/0 -> NaN like the reference when a group has no useful row
    out[0] = static_cast<float>(ws[0] * num_scale + cat * cat_scale);
}
__global__ void __launch_bounds__(256) synth_loss_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ vo, const float* __restrict__ vi,
                                                             SynthTables t, const double* __restrict__ ws, float* __restrict__ dvo,
                                                             float num_scale2, float cat_scale, float inv_temp, int use_softmax, int B, int L) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float gup = gout[0];
    const float *o = vo + static_cast<size_t>(row) * L, *in = vi + static_cast<size_t>(row) * L;
    float* d = dvo + static_cast<size_t>(row) * L;
    for (int j = lane; j < t.n_num; j += 32) {
        const int vol = t.num_vol[j], c = t.num_cols[j];
        d[c] = (vol >= 0 && in[vol] < 1e-3f) ? 0.0f : gup * num_scale2 * (o[c] - in[c]);
    }
    for (int g = 0; g < t.n_grp; ++g) {
        const int vol = t.grp_vol[g], s = t.grp_start[g], n = t.grp_len[g];
        if (vol >= 0 && in[vol] < 1e-3f) {
            for (int j = lane; j < n; j += 32) d[s + j] = 0.0f;
            continue;
        }
        const float c = gup * cat_scale / static_cast<float>(ws[1 + t.n_grp + g]);
        if (use_softmax) {
            float mx = -INFINITY, cnt = 0.0f;
            for (int j = lane; j < n; j += 32) { mx = fmaxf(mx, o[s + j] * inv_temp); cnt += in[s + j] != 0.0f ? 1.0f : 0.0f; }
            mx = warp_max_ff(mx);
            cnt = warp_sum_ff(cnt);
            float sum = 0.0f;
            for (int j = lane; j < n; j += 32) sum += expf(o[s + j] * inv_temp - mx);
            sum = warp_sum_ff(sum);
            for (int j = lane; j < n; j += 32) {
                const float q = expf(o[s + j] * inv_temp - mx) / sum;
                d[s + j] = c * inv_temp * (cnt * q - (in[s + j] != 0.0f ? 1.0f : 0.0f));
            }
        } else {
            for (int j = lane; j < n; j += 32) d[s + j] = in[s + j] != 0.0f ? -c / o[s + j] : 0.0f;
        }
    }
}

__global__ void __launch_bounds__(128) useful_counts_kernel(const float* __restrict__ vi, int B, int L, const int* __restrict__ grp_vol,
                                                            double* __restrict__ counts) {
    __shared__ int red[4];
    const int vol = grp_vol[blockIdx.x];
    int c = 0;
    for (int b = threadIdx.x; b < B; b += 128) c += (vol >= 0 && vi[static_cast<size_t>(b) * L + vol] < 1e-3f) ? 0 : 1;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = static_cast<double>(red[0] + red[1] + red[2] + red[3]);
}

// ------------------------------------------------------------------------------------------------ Adam
// hyper (device, optional) = {lr, 1 - beta1^t, sqrt(1 - beta2^t), grad_scale}: read at run time so that a captured CUDA
// graph can be replayed with a new learning rate / step count.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                            float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt, float grad_scale,
                            const float* __restrict__ hyper) {
    if (hyper != nullptr) { lr = hyper[0]; bc1 = hyper[1]; bc2_sqrt = hyper[2]; grad_scale = hyper[3]; }
    const float step = lr / bc1;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float pi = p[i];
        const float gi = fmaf(weight_decay, pi, g[i] * grad_scale);
        const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

// Gather many tensors into one flat buffer (gradient bucket for the all-reduce / fused Adam).  table[3*i] = source
// address, table[3*i+1] = destination offset (elements), table[3*i+2] = element count; gridDim.y = tensors.
__global__ void multi_pack_kernel(const unsigned long long* __restrict__ table, float* __restrict__ flat, float scale) {
    const float* src = reinterpret_cast<const float*>(table[3 * blockIdx.y]);
    float* dst = flat + table[3 * blockIdx.y + 1];
    const size_t n = table[3 * blockIdx.y + 2];
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[i] = src[i] * scale;
}

}  // namespace pgv

using namespace pgv;
#define PGV_STREAM(s) static_cast<cudaStream_t>(s)

extern "C" {

/* loss = scale * sum((a-b)^2); workspace: >= 8 bytes */
int pgv_sqerr_fwd(const float* a, const float* b, size_t n, double scale, float* loss_out, void* workspace, pgv_stream_t stream) {
    PGV_CHECK_ARG(a && b && loss_out && workspace && n > 0, "pgv_sqerr_fwd: bad argument");
    cudaStream_t s = PGV_STREAM(stream);
    double* ws = static_cast<double*>(workspace);
    PGV_CUDA(cudaMemsetAsync(ws, 0, sizeof(double), s));
    sqerr_sum_kernel<<<grid1(n), 256, 0, s>>>(a, b, ws, n);
    PGV_LAUNCH_CHECK();
    scale_to_float_kernel<<<1, 1, 0, s>>>(ws, loss_out, scale);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_sqerr_bwd(const float* a, const float* b, size_t n, double scale, const float* grad_out, float* da, pgv_stream_t stream) {
    PGV_CHECK_ARG(a && b && grad_out && da && n > 0, "pgv_sqerr_bwd: bad argument");
    sqerr_bwd_kernel<<<grid1(n), 256, 0, PGV_STREAM(stream)>>>(a, b, grad_out, da, static_cast<float>(2.0 * scale), n);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_latent_loss_fwd(const float* mu_logvar, const float* z0, const float* zk, const float* logdet, int B, int D, int normalize,
                        float* loss_out, void* workspace, pgv_stream_t stream) {
    PGV_CHECK_ARG(mu_logvar && z0 && zk && logdet && loss_out && workspace && B > 0 && D > 0, "pgv_latent_loss_fwd: bad argument");
    cudaStream_t s = PGV_STREAM(stream);
    double* ws = static_cast<double*>(workspace);
    PGV_CUDA(cudaMemsetAsync(ws, 0, sizeof(double), s));
    latent_loss_fwd_kernel<<<ceil_div(B, 8), 256, 0, s>>>(mu_logvar, z0, zk, logdet, ws, B, D);
    PGV_LAUNCH_CHECK();
    scale_to_float_kernel<<<1, 1, 0, s>>>(ws, loss_out, 1.0 / B / (normalize ? D : 1));
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_latent_loss_bwd(const float* grad_out, const float* mu_logvar, const float* z0, const float* zk, int B, int D, int normalize,
                        float* d_mu_logvar, float* dz0, float* dzk, float* dlogdet, pgv_stream_t stream) {
    PGV_CHECK_ARG(grad_out && mu_logvar && z0 && zk && d_mu_logvar && dz0 && dzk && dlogdet, "pgv_latent_loss_bwd: NULL argument");
    latent_loss_bwd_kernel<<<grid1(static_cast<size_t>(B) * D), 256, 0, PGV_STREAM(stream)>>>(
        grad_out, mu_logvar, z0, zk, d_mu_logvar, dz0, dzk, dlogdet, 1.0f / B / (normalize ? D : 1), B, D);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_dkl_fwd(const float* mu_logvar, int B, int D, int normalize, float* loss_out, void* workspace, pgv_stream_t stream) {
    PGV_CHECK_ARG(mu_logvar && loss_out && workspace && B > 0 && D > 0, "pgv_dkl_fwd: bad argument");
    cudaStream_t s = PGV_STREAM(stream);
    double* ws = static_cast<double*>(workspace);
    PGV_CUDA(cudaMemsetAsync(ws, 0, sizeof(double), s));
    dkl_fwd_kernel<<<grid1(static_cast<size_t>(B) * D), 256, 0, s>>>(mu_logvar, ws, B, D);
    PGV_LAUNCH_CHECK();
    scale_to_float_kernel<<<1, 1, 0, s>>>(ws, loss_out, 1.0 / B / (normalize ? D : 1));
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_dkl_bwd(const float* grad_out, const float* mu_logvar, int B, int D, int normalize, float* d_mu_logvar, pgv_stream_t stream) {
    PGV_CHECK_ARG(grad_out && mu_logvar && d_mu_logvar, "pgv_dkl_bwd: NULL argument");
    dkl_bwd_kernel<<<grid1(static_cast<size_t>(B) * D), 256, 0, PGV_STREAM(stream)>>>(grad_out, mu_logvar, d_mu_logvar,
                                                                                   1.0f / B / (normalize ? D : 1), B, D);
    PGV_LAUNCH_CHECK();
    return 0;
}

size_t pgv_synth_loss_workspace_bytes(int n_groups) { return sizeof(double) * (1 + 2 * static_cast<size_t>(n_groups)); }

int pgv_synth_loss_fwd(const float* v_out, const float* v_in, int B, int L, const int* num_cols, const int* num_vol_col, int n_num,
                       const int* grp_start, const int* grp_len, const int* grp_vol_col, int n_grp, int normalize,
                       float cat_loss_factor, int cat_softmax, float softmax_temperature, const double* group_counts, float* loss_out,
                       void* workspace, pgv_stream_t stream) {
    PGV_CHECK_ARG(v_out && v_in && loss_out && workspace && B > 0 && L > 0, "pgv_synth_loss_fwd: bad argument");
    PGV_CHECK_ARG(n_num == 0 || (num_cols && num_vol_col), "pgv_synth_loss_fwd: numerical tables missing");
    PGV_CHECK_ARG(n_grp == 0 || (grp_start && grp_len && grp_vol_col), "pgv_synth_loss_fwd: categorical tables missing");
    cudaStream_t s = PGV_STREAM(stream);
    double* ws = static_cast<double*>(workspace);
    PGV_CUDA(cudaMemsetAsync(ws, 0, pgv_synth_loss_workspace_bytes(n_grp), s));
    SynthTables t{num_cols, num_vol_col, grp_start, grp_len, grp_vol_col, n_num, n_grp};
    synth_loss_fwd_kernel<<<ceil_div(B, 8), 256, 0, s>>>(v_out, v_in, t, ws, 1.0f / softmax_temperature, cat_softmax, B, L);
    PGV_LAUNCH_CHECK();
    // numerical: MSELoss(mean) over [B, n_num] when normalised, else L2Loss (sum / B)          loss.py:105-108, 136
    const double num_scale = n_num > 0 ? (normalize ? 1.0 / (static_cast<double>(B) * n_num) : 1.0 / B) : 0.0;
    const double cat_scale = n_grp > 0 ? cat_loss_factor / (normalize ? n_grp : 1) : 0.0;    // loss.py:180-183
    synth_loss_finish_kernel<<<1, 1, 0, s>>>(ws, loss_out, num_scale, cat_scale, n_grp, group_counts);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_synth_loss_bwd(const float* grad_out, const float* v_out, const float* v_in, int B, int L, const int* num_cols,
                       const int* num_vol_col, int n_num, const int* grp_start, const int* grp_len, const int* grp_vol_col, int n_grp,
                       int normalize, float cat_loss_factor, int cat_softmax, float softmax_temperature, const void* workspace,
                       float* d_v_out, pgv_stream_t stream) {
    PGV_CHECK_ARG(grad_out && v_out && v_in && workspace && d_v_out, "pgv_synth_loss_bwd: NULL argument");
    SynthTables t{num_cols, num_vol_col, grp_start, grp_len, grp_vol_col, n_num, n_grp};
    const double num_scale = n_num > 0 ? (normalize ? 1.0 / (static_cast<double>(B) * n_num) : 1.0 / B) : 0.0;
    const double cat_scale = n_grp > 0 ? cat_loss_factor / (normalize ? n_grp : 1) : 0.0;
    synth_loss_bwd_kernel<<<ceil_div(B, 8), 256, 0, PGV_STREAM(stream)>>>(grad_out, v_out, v_in, t, static_cast<const double*>(workspace), d_v_out,
                                                                         static_cast<float>(2.0 * num_scale), static_cast<float>(cat_scale),
                                                                         1.0f / softmax_temperature, cat_softmax, B, L);
    PGV_LAUNCH_CHECK();
    return 0;
}

/* counts[g] = rows of v_in whose operator-volume column does not silence categorical group g (data/preset.py:264-281) */
int pgv_synth_useful_counts(const float* v_in, int B, int L, const int* grp_vol_col, int n_grp, double* counts, pgv_stream_t stream) {
    PGV_CHECK_ARG(v_in && grp_vol_col && counts && B > 0 && L > 0 && n_grp > 0, "pgv_synth_useful_counts: bad argument");
    useful_counts_kernel<<<n_grp, 128, 0, PGV_STREAM(stream)>>>(v_in, B, L, grp_vol_col, counts);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_multi_pack(const void* table_dev, int n_tensors, size_t max_elems, float* flat, float scale, pgv_stream_t stream) {
    PGV_CHECK_ARG(table_dev && flat && n_tensors > 0, "pgv_multi_pack: bad argument");
    size_t bx = (max_elems + 256 * 16 - 1) / (256 * 16);     // blocks beyond a tensor's size exit at once
    if (bx < 1) bx = 1;
    if (bx > 2048) bx = 2048;
    multi_pack_kernel<<<dim3(static_cast<unsigned>(bx), n_tensors), 256, 0, PGV_STREAM(stream)>>>(
        static_cast<const unsigned long long*>(table_dev), flat, scale);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, pgv_stream_t stream) {
    PGV_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step >= 1, "pgv_adam_step: bad argument");
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step), bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    size_t blocks = (n + 255) / 256;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    adam_kernel<<<static_cast<int>(blocks), 256, 0, PGV_STREAM(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                                          static_cast<float>(bc1), static_cast<float>(sqrt(bc2)), grad_scale, nullptr);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, const float* hyper_dev, float beta1,
                      float beta2, float eps, float weight_decay, pgv_stream_t stream) {
    PGV_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && hyper_dev, "pgv_adam_step_dev: NULL argument");
    if (n == 0) return 0;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    adam_kernel<<<static_cast<int>(blocks), 256, 0, PGV_STREAM(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, 0.f, beta1, beta2, eps, weight_decay,
                                                                          1.f, 1.f, 1.f, hyper_dev);
    PGV_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
