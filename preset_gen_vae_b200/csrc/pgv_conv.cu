// Direct (CUDA-core, exact fp32) 2-D convolution family on NCHW tensors.
//   conv_fwd   : y = act(conv(x, w) + bias)                       nn.Conv2d           model/layer.py:19
//   conv_dgrad : dx = conv^T(dy, w)  (gather form, no atomics)     == nn.ConvTranspose2d forward, model/layer.py:38
//   conv_wgrad : dw += x (*) dy, db += sum(dy)
// A transposed convolution with weight [Cin, Cout, kh, kw] is the data-gradient of the convolution that has the same
// weight tensor, so TConv2D uses conv_dgrad as its forward, conv_fwd as its data-gradient and conv_wgrad with the
// roles of x and dy exchanged.  These kernels serve the thin HBM-bound layers (enc1, dec8), small batches, and are
// the on-device cross-check for the tcgen05 implicit-GEMM kernels in pgv_conv_tc.cu.
#include "pgv_common.cuh"

namespace pgv {

struct ConvGeom {
    int B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo;
};

__device__ __forceinline__ float apply_act(float v, float slope) { return (slope >= 0.0f && v < 0.0f) ? v * slope : v; }

// One thread: one output pixel, CO_T consecutive output channels.  grid.x covers B*Ho*Wo pixels, grid.y channel groups.
template <int CO_T>
__global__ void __launch_bounds__(128) conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y, ConvGeom g,
                                                       float slope) {
    const long long pix = blockIdx.x * 128LL + threadIdx.x;
    const long long npix = static_cast<long long>(g.B) * g.Ho * g.Wo;
    if (pix >= npix) return;
    const int ow = static_cast<int>(pix % g.Wo), oh = static_cast<int>((pix / g.Wo) % g.Ho), b = static_cast<int>(pix / (g.Wo * g.Ho));
    const int co0 = blockIdx.y * CO_T;
    float acc[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[j] = (bias != nullptr && co0 + j < g.Cout) ? bias[co0 + j] : 0.0f;
    const int ih0 = oh * g.stride - g.pad, iw0 = ow * g.stride - g.pad;
    const size_t wstride = static_cast<size_t>(g.Cin) * g.kh * g.kw;
    for (int ci = 0; ci < g.Cin; ++ci) {
        const float* xp = x + (static_cast<size_t>(b) * g.Cin + ci) * g.H * g.W;
        const float* wp = w + static_cast<size_t>(co0) * wstride + static_cast<size_t>(ci) * g.kh * g.kw;
        for (int r = 0; r < g.kh; ++r) {
            const int ih = ih0 + r;
            if (ih < 0 || ih >= g.H) continue;
            for (int s = 0; s < g.kw; ++s) {
                const int iw = iw0 + s;
                if (iw < 0 || iw >= g.W) continue;
                const float xv = xp[ih * g.W + iw];
#pragma unroll
                for (int j = 0; j < CO_T; ++j)
                    if (co0 + j < g.Cout) acc[j] = fmaf(xv, __ldg(wp + j * wstride + r * g.kw + s), acc[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CO_T; ++j)
        if (co0 + j < g.Cout) y[((static_cast<size_t>(b) * g.Cout + co0 + j) * g.Ho + oh) * g.Wo + ow] = apply_act(acc[j], slope);
}

// One thread: one input pixel, CI_T consecutive input channels; sums over the output pixels that read it.
// `bias`/`slope` are used when this kernel runs as the forward of a transposed convolution.
template <int CI_T>
__global__ void __launch_bounds__(128) conv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ dx, ConvGeom g,
                                                         float slope) {
    const long long pix = blockIdx.x * 128LL + threadIdx.x;
    const long long npix = static_cast<long long>(g.B) * g.H * g.W;
    if (pix >= npix) return;
    const int iw = static_cast<int>(pix % g.W), ih = static_cast<int>((pix / g.W) % g.H), b = static_cast<int>(pix / (g.W * g.H));
    const int ci0 = blockIdx.y * CI_T;
    float acc[CI_T];
#pragma unroll
    for (int j = 0; j < CI_T; ++j) acc[j] = (bias != nullptr && ci0 + j < g.Cin) ? bias[ci0 + j] : 0.0f;
    const int khw = g.kh * g.kw;
    for (int r = 0; r < g.kh; ++r) {
        const int th = ih + g.pad - r;
        if (th < 0 || th % g.stride != 0) continue;
        const int oh = th / g.stride;
        if (oh >= g.Ho) continue;
        for (int s = 0; s < g.kw; ++s) {
            const int tw = iw + g.pad - s;
            if (tw < 0 || tw % g.stride != 0) continue;
            const int ow = tw / g.stride;
            if (ow >= g.Wo) continue;
            const float* dyp = dy + (static_cast<size_t>(b) * g.Cout * g.Ho + oh) * g.Wo + ow;
            const float* wp = w + static_cast<size_t>(ci0) * khw + r * g.kw + s;
            for (int co = 0; co < g.Cout; ++co) {
                const float dv = dyp[static_cast<size_t>(co) * g.Ho * g.Wo];
                const float* wc = wp + static_cast<size_t>(co) * g.Cin * khw;
#pragma unroll
                for (int j = 0; j < CI_T; ++j)
                    if (ci0 + j < g.Cin) acc[j] = fmaf(dv, __ldg(wc + j * khw), acc[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CI_T; ++j)
        if (ci0 + j < g.Cin) dx[((static_cast<size_t>(b) * g.Cin + ci0 + j) * g.H + ih) * g.W + iw] = apply_act(acc[j], slope);
}

// dw[co, ci, r, s] += sum over (b, oh, ow) of dy[b,co,oh,ow] * x[b,ci,oh*stride-pad+r, ow*stride-pad+s].
// Block = one (co, ci) pair and one slice of the batch; each thread keeps kh*kw (<= 25) partial sums.
constexpr int WGRAD_MAX_TAPS = 25;
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dw, ConvGeom g, int b_per_block) {
    const int co = blockIdx.x / g.Cin, ci = blockIdx.x % g.Cin;
    const int b0 = blockIdx.y * b_per_block, b1 = min(g.B, b0 + b_per_block);
    const int taps = g.kh * g.kw;
    float acc[WGRAD_MAX_TAPS];
#pragma unroll
    for (int t = 0; t < WGRAD_MAX_TAPS; ++t) acc[t] = 0.0f;
    const int npix = g.Ho * g.Wo;
    for (int b = b0; b < b1; ++b) {
        const float* dyp = dy + (static_cast<size_t>(b) * g.Cout + co) * npix;
        const float* xp = x + (static_cast<size_t>(b) * g.Cin + ci) * g.H * g.W;
        for (int p = threadIdx.x; p < npix; p += 256) {
            const float dv = dyp[p];
            const int oh = p / g.Wo, ow = p % g.Wo;
            const int ih0 = oh * g.stride - g.pad, iw0 = ow * g.stride - g.pad;
#pragma unroll
            for (int t = 0; t < WGRAD_MAX_TAPS; ++t) {
                if (t < taps) {
                    const int ih = ih0 + t / g.kw, iw = iw0 + t % g.kw;
                    if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) acc[t] = fmaf(dv, xp[ih * g.W + iw], acc[t]);
                }
            }
        }
    }
    __shared__ float red[8][WGRAD_MAX_TAPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < WGRAD_MAX_TAPS; ++t) {
        float v = acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < taps) {
        float v = 0.0f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
        atomicAdd(dw + (static_cast<size_t>(co) * g.Cin + ci) * taps + threadIdx.x, v);
    }
}

// db[c] += sum over (b, h, w) of dy[b, c, h, w].  grid = (C, images, chunks of H*W): with one channel (the decoder's last layer:
// 57 MB per step) the images alone are too few blocks to pull the HBM bandwidth.
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ dy, float* __restrict__ db, int B, int C, int HW) {
    const int c = blockIdx.x;
    const int len = (HW + gridDim.z - 1) / gridDim.z, lo = blockIdx.z * len, hi = min(HW, lo + len);
    double acc = 0.0;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float* p = dy + (static_cast<size_t>(b) * C + c) * HW;
        float part = 0.0f;
        for (int i = lo + threadIdx.x; i < hi; i += 256) part += __ldg(p + i);
        acc += part;
    }
    __shared__ double red[8];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int i = 0; i < 8; ++i) v += red[i];
        atomicAdd(db + c, static_cast<float>(v));
    }
}

// grid for channel_sum_kernel: at least ~4 blocks per SM when the tensor is large enough to give each block >= 1024 elements
static dim3 channel_sum_grid(int B, int C, int HW) {
    const int nb = B < 64 ? B : 64;
    long long chunks = (148LL * 4 + static_cast<long long>(C) * nb - 1) / (static_cast<long long>(C) * nb);
    const long long max_chunks = (HW + 1023) / 1024;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    return dim3(C, nb, static_cast<unsigned>(chunks));
}

static int check_geom(const ConvGeom& g, const char* who) {
    if (g.B <= 0 || g.Cin <= 0 || g.Cout <= 0 || g.H <= 0 || g.W <= 0 || g.Ho <= 0 || g.Wo <= 0 || g.kh <= 0 || g.kw <= 0 ||
        g.stride <= 0 || g.pad < 0)
        return set_error(-1, "%s: bad geometry", who);
    if (g.kh * g.kw > WGRAD_MAX_TAPS) return set_error(-1, "%s: kernels larger than 5x5 are not supported", who);
    // (Ho-1)*stride - 2*pad + k <= H  must hold for the conv view; for the transposed view H may exceed it by output_padding
    if ((g.Ho - 1) * g.stride - 2 * g.pad + g.kh > g.H || (g.Wo - 1) * g.stride - 2 * g.pad + g.kw > g.W)
        return set_error(-1, "%s: output %dx%d does not fit input %dx%d", who, g.Ho, g.Wo, g.H, g.W);
    return 0;
}

}  // namespace pgv

using namespace pgv;

extern "C" {

int pgv_conv2d_fwd_f32(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H, int W, int Cout, int kh,
                       int kw, int stride, int pad, int Ho, int Wo, float lrelu_slope, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && w && y, "pgv_conv2d_fwd_f32: NULL argument");
    ConvGeom g{B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo};
    if (int rc = check_geom(g, "pgv_conv2d_fwd_f32")) return rc;
    const long long npix = static_cast<long long>(B) * Ho * Wo;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (Cout >= 8) {
        dim3 grid(static_cast<unsigned>((npix + 127) / 128), ceil_div(Cout, 8));
        conv_fwd_kernel<8><<<grid, 128, 0, s>>>(x, w, bias, y, g, lrelu_slope);
    } else {
        dim3 grid(static_cast<unsigned>((npix + 127) / 128), Cout);
        conv_fwd_kernel<1><<<grid, 128, 0, s>>>(x, w, bias, y, g, lrelu_slope);
    }
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_conv2d_dgrad_f32(const float* dy, const float* w, const float* bias, float* dx, int B, int Cin, int H, int W, int Cout,
                         int kh, int kw, int stride, int pad, int Ho, int Wo, float lrelu_slope, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && w && dx, "pgv_conv2d_dgrad_f32: NULL argument");
    ConvGeom g{B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo};
    if (int rc = check_geom(g, "pgv_conv2d_dgrad_f32")) return rc;
    const long long npix = static_cast<long long>(B) * H * W;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (Cin >= 8) {
        dim3 grid(static_cast<unsigned>((npix + 127) / 128), ceil_div(Cin, 8));
        conv_dgrad_kernel<8><<<grid, 128, 0, s>>>(dy, w, bias, dx, g, lrelu_slope);
    } else {
        dim3 grid(static_cast<unsigned>((npix + 127) / 128), Cin);
        conv_dgrad_kernel<1><<<grid, 128, 0, s>>>(dy, w, bias, dx, g, lrelu_slope);
    }
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int H, int W, int Cout, int kh,
                         int kw, int stride, int pad, int Ho, int Wo, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && dy && dw, "pgv_conv2d_wgrad_f32: NULL argument");
    ConvGeom g{B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo};
    if (int rc = check_geom(g, "pgv_conv2d_wgrad_f32")) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PGV_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * Cout * static_cast<size_t>(Cin) * kh * kw, s));
    // enough blocks to fill the chip: split the batch when there are few (co, ci) pairs
    const long long pairs = static_cast<long long>(Cout) * Cin;
    int splits = static_cast<int>((148LL * 8 + pairs - 1) / pairs);
    if (splits < 1) splits = 1;
    if (splits > B) splits = B;
    const int b_per_block = ceil_div(B, splits);
    dim3 grid(static_cast<unsigned>(pairs), ceil_div(B, b_per_block));
    conv_wgrad_kernel<<<grid, 256, 0, s>>>(x, dy, dw, g, b_per_block);
    PGV_LAUNCH_CHECK();
    if (db != nullptr) {
        PGV_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cout, s));
        channel_sum_kernel<<<channel_sum_grid(B, Cout, Ho * Wo), 256, 0, s>>>(dy, db, B, Cout, Ho * Wo);
        PGV_LAUNCH_CHECK();
    }
    return 0;
}

/* out[c] = sum over (b, h, w) of x[b, c, h, w] */
int pgv_channel_sum(const float* x, float* out, int B, int C, int HW, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && out && B > 0 && C > 0 && HW > 0, "pgv_channel_sum: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PGV_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, s));
    channel_sum_kernel<<<channel_sum_grid(B, C, HW), 256, 0, s>>>(x, out, B, C, HW);
    PGV_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
