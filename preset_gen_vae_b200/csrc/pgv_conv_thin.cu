// Direct kernels for the two thin 5x5 / stride-2 layers that touch the full-resolution spectrogram:
//   enc1  nn.Conv2d(1, 8, 5, 2, 2)            model/encoder.py:241      (1 x 257 x 347 -> 8 x 129 x 174)
//   dec8  nn.ConvTranspose2d(8, 1, 5, 2, 2)   model/decoder.py:218      (8 x 129 x 174 -> 1 x 257 x 347) + Hardtanh
// With one channel on one side these are 8-25 flop/byte: HBM-bound, so they are written as coalesced
// streaming kernels with the 200 weights in shared memory instead of 128-row tensor-core tiles with 1 useful column.
// Exact fp32 arithmetic.  "Conv view" geometry for all three: x [B,1,H,W], y [B,C,Ho,Wo], w [C,1,5,5], stride 2, pad 2.
#include <algorithm>

#include "pgv_common.cuh"
#include "pgv_tc.cuh"

namespace pgv {

constexpr int THIN_K = 5, THIN_TAPS = 25, THIN_MAXC = 8;

// The 200 weights (+ biases) of a thin layer, copied device-to-device into constant memory right before the launch (a memcpy
// node when captured): every thread uses the same weight at the same time, so the FMAs take it as a constant-bank operand
// instead of one shared-memory load per FMA.  One bank per kernel so that a forward and a transposed launch never share it.
__constant__ float c_thin_fwd_w[THIN_MAXC * THIN_TAPS + THIN_MAXC];
__constant__ float c_thin_quad_w[THIN_MAXC * THIN_TAPS + 1];       // + the single output-channel bias

__device__ __forceinline__ float thin_round(float v, int round_out) {
    if (!round_out) return v;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// y[b,c,oh,ow] = act(bias[c] + sum_{r,s} x[b,0,2oh-2+r,2ow-2+s] * w[c,0,r,s]).  One thread per output pixel, all C channels.
// CL: y is channels-last [B, Ho, Wo, C] (C == 8: two float4 stores per pixel), optionally rounded to TF32.
template <bool CL>
__global__ void __launch_bounds__(256) thin_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y, int B, int C, int H,
                                                            int W, int Ho, int Wo, float slope, int round_out) {
    __shared__ float sw[THIN_MAXC * THIN_TAPS], sb[THIN_MAXC];
    for (int i = threadIdx.x; i < C * THIN_TAPS; i += 256) sw[i] = w[i];
    if (threadIdx.x < C) sb[threadIdx.x] = bias != nullptr ? bias[threadIdx.x] : 0.0f;
    __syncthreads();
    const int HWo = Ho * Wo;
    const long long total = static_cast<long long>(B) * HWo;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int b = static_cast<int>(i / HWo), pix = static_cast<int>(i % HWo), oh = pix / Wo, ow = pix % Wo;
        const float* xb = x + static_cast<size_t>(b) * H * W;
        float in[THIN_TAPS];
#pragma unroll
        for (int r = 0; r < THIN_K; ++r) {
            const int ih = 2 * oh - 2 + r;
#pragma unroll
            for (int s = 0; s < THIN_K; ++s) {
                const int iw = 2 * ow - 2 + s;
                in[r * THIN_K + s] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(xb + ih * W + iw) : 0.0f;
            }
        }
        if (CL && C == THIN_MAXC) {
            float o[THIN_MAXC];
#pragma unroll
            for (int c = 0; c < THIN_MAXC; ++c) {
                float acc = sb[c];
#pragma unroll
                for (int t = 0; t < THIN_TAPS; ++t) acc = fmaf(in[t], sw[c * THIN_TAPS + t], acc);
                o[c] = thin_round((slope >= 0.0f && acc < 0.0f) ? acc * slope : acc, round_out);
            }
            float4* yo = reinterpret_cast<float4*>(y + static_cast<size_t>(i) * THIN_MAXC);
            yo[0] = make_float4(o[0], o[1], o[2], o[3]);
            yo[1] = make_float4(o[4], o[5], o[6], o[7]);
            continue;
        }
        float* yb = CL ? y + static_cast<size_t>(i) * C : y + static_cast<size_t>(b) * C * HWo + pix;
        const size_t cs = CL ? 1 : static_cast<size_t>(HWo);
        for (int c = 0; c < C; ++c) {
            float acc = sb[c];
#pragma unroll
            for (int t = 0; t < THIN_TAPS; ++t) acc = fmaf(in[t], sw[c * THIN_TAPS + t], acc);
            yb[c * cs] = thin_round((slope >= 0.0f && acc < 0.0f) ? acc * slope : acc, round_out);
        }
    }
}

// Channels-last, C == 8 forward: weights and biases as constant-bank operands (c_thin_fwd_w), two float4 stores per pixel.
__global__ void __launch_bounds__(256) thin_conv_fwd_cl8_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int Ho,
                                                                int Wo, float slope, int round_out) {
    const int HWo = Ho * Wo;
    const long long total = static_cast<long long>(B) * HWo;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int b = static_cast<int>(i / HWo), pix = static_cast<int>(i % HWo), oh = pix / Wo, ow = pix % Wo;
        const float* xb = x + static_cast<size_t>(b) * H * W;
        float in[THIN_TAPS];
#pragma unroll
        for (int r = 0; r < THIN_K; ++r) {
            const int ih = 2 * oh - 2 + r;
#pragma unroll
            for (int s = 0; s < THIN_K; ++s) {
                const int iw = 2 * ow - 2 + s;
                in[r * THIN_K + s] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(xb + ih * W + iw) : 0.0f;
            }
        }
        float o[THIN_MAXC];
#pragma unroll
        for (int c = 0; c < THIN_MAXC; ++c) {
            float acc = c_thin_fwd_w[THIN_MAXC * THIN_TAPS + c];
#pragma unroll
            for (int t = 0; t < THIN_TAPS; ++t) acc = fmaf(in[t], c_thin_fwd_w[c * THIN_TAPS + t], acc);
            o[c] = thin_round((slope >= 0.0f && acc < 0.0f) ? acc * slope : acc, round_out);
        }
        float4* yo = reinterpret_cast<float4*>(y + static_cast<size_t>(i) * THIN_MAXC);
        yo[0] = make_float4(o[0], o[1], o[2], o[3]);
        yo[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// Channels-last, C == 8 transposed form: one thread per 2x2 output quad (ih = 2i + ph, iw = 2j + pw).  The four pixels read the
// same 3x3 neighbourhood of y (rows i+1-a, columns j+1-d; kernel element (ph + 2a, pw + 2d)), loaded once as 18 float4;
// weights are constant-bank operands (c_thin_quad_w).
__global__ void __launch_bounds__(256) thin_conv_quad_cl8_kernel(const float* __restrict__ y, float* __restrict__ x, int B, int H, int W, int Ho,
                                                                 int Wo, float lo, float hi) {
    const float bias0 = c_thin_quad_w[THIN_MAXC * THIN_TAPS];
    const int Hq = (H + 1) >> 1, Wq = (W + 1) >> 1;
    const long long total = static_cast<long long>(B) * Hq * Wq;
    for (long long q = blockIdx.x * 256LL + threadIdx.x; q < total; q += 256LL * gridDim.x) {
        const int b = static_cast<int>(q / (Hq * Wq)), rq = static_cast<int>(q % (Hq * Wq)), i = rq / Wq, j = rq % Wq;
        const float* yb = y + static_cast<size_t>(b) * Ho * Wo * THIN_MAXC;
        float acc[2][2] = {{bias0, bias0}, {bias0, bias0}};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int oh = i + 1 - a;
            if (oh < 0 || oh >= Ho) continue;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int ow = j + 1 - d;
                if (ow < 0 || ow >= Wo) continue;
                const float4* src = reinterpret_cast<const float4*>(yb + (static_cast<size_t>(oh) * Wo + ow) * THIN_MAXC);
                const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
                const float vs[THIN_MAXC] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    if (ph + 2 * a >= THIN_K) continue;
#pragma unroll
                    for (int pw = 0; pw < 2; ++pw) {
                        if (pw + 2 * d >= THIN_K) continue;
#pragma unroll
                        for (int c = 0; c < THIN_MAXC; ++c)
                            acc[ph][pw] = fmaf(vs[c], c_thin_quad_w[c * THIN_TAPS + (ph + 2 * a) * THIN_K + pw + 2 * d], acc[ph][pw]);
                    }
                }
            }
        }
        float* xb = x + static_cast<size_t>(b) * H * W;
#pragma unroll
        for (int ph = 0; ph < 2; ++ph)
#pragma unroll
            for (int pw = 0; pw < 2; ++pw) {
                const int ih = 2 * i + ph, iw = 2 * j + pw;
                if (ih < H && iw < W) xb[ih * W + iw] = clamp_nan(acc[ph][pw], lo, hi);
            }
    }
}

// Transposed form: x[b,0,ih,iw] = clamp(bias + sum_{c,r,s: 2oh-2+r = ih, 2ow-2+s = iw} y[b,c,oh,ow] * w[c,0,r,s], lo, hi).
// One thread per full-resolution pixel; taps r = (ih & 1) + 2a.
template <bool CL>
__global__ void __launch_bounds__(256) thin_conv_dgrad_kernel(const float* __restrict__ y, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ x, int B, int C, int H,
                                                              int W, int Ho, int Wo, float lo, float hi) {
    __shared__ float sw[THIN_MAXC * THIN_TAPS];
    for (int i = threadIdx.x; i < C * THIN_TAPS; i += 256) sw[i] = w[i];
    __syncthreads();
    const float b0 = bias != nullptr ? bias[0] : 0.0f;
    const int HW = H * W, HWo = Ho * Wo;
    const long long total = static_cast<long long>(B) * HW;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int b = static_cast<int>(i / HW), pix = static_cast<int>(i % HW), ih = pix / W, iw = pix % W;
        const float* yb = y + static_cast<size_t>(b) * C * HWo;
        const int r0 = ih & 1, s0 = iw & 1, oh0 = (ih + 2) >> 1, ow0 = (iw + 2) >> 1;   // tap a: r = r0 + 2a, oh = oh0 - a
        float acc = b0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int r = r0 + 2 * a, oh = oh0 - a;
            if (r >= THIN_K || oh < 0 || oh >= Ho) continue;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int s = s0 + 2 * d, ow = ow0 - d;
                if (s >= THIN_K || ow < 0 || ow >= Wo) continue;
                if (CL) {                                    // y is [B, Ho, Wo, C]
                    const float* src = yb + (static_cast<size_t>(oh) * Wo + ow) * C;
                    if (C == THIN_MAXC) {
                        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                        const float vs[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                        for (int c = 0; c < THIN_MAXC; ++c) acc = fmaf(vs[c], sw[c * THIN_TAPS + r * THIN_K + s], acc);
                    } else {
                        for (int c = 0; c < C; ++c) acc = fmaf(__ldg(src + c), sw[c * THIN_TAPS + r * THIN_K + s], acc);
                    }
                } else {
                    const float* src = yb + oh * Wo + ow;
                    for (int c = 0; c < C; ++c) acc = fmaf(__ldg(src + static_cast<size_t>(c) * HWo), sw[c * THIN_TAPS + r * THIN_K + s], acc);
                }
            }
        }
        x[i] = clamp_nan(acc, lo, hi);
    }
}

// dw[c,0,r,s] = sum_{b,oh,ow} dy[b,c,oh,ow] * x[b,0,2oh-2+r,2ow-2+s].  One block per image slice: thread (c, tap) walks
// shared-memory tiles of dy (C x 8 x 32) and of the matching input patch (19 x 67); one atomicAdd per thread at the end.
constexpr int TW_TH = 8, TW_TW = 32, TW_PH = 2 * TW_TH + 3, TW_PW = 2 * TW_TW + 3;
__global__ void __launch_bounds__(256) thin_conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              float* __restrict__ dw, int B, int C, int H, int W, int Ho, int Wo,
                                                              int tiles_per_block, int channels_last, float* __restrict__ partial) {
    __shared__ float sdy[THIN_MAXC][TW_TH * TW_TW];
    __shared__ float sx[TW_PH * TW_PW];
    const int tiles_h = (Ho + TW_TH - 1) / TW_TH, tiles_w = (Wo + TW_TW - 1) / TW_TW, tiles_img = tiles_h * tiles_w;
    const long long n_tiles = static_cast<long long>(B) * tiles_img;
    const int c = threadIdx.x / THIN_TAPS, tap = threadIdx.x % THIN_TAPS, r = tap / THIN_K, s = tap % THIN_K;
    const bool worker = threadIdx.x < C * THIN_TAPS;
    float acc = 0.0f;
    const long long t0 = static_cast<long long>(blockIdx.x) * tiles_per_block;
    for (long long t = t0; t < t0 + tiles_per_block && t < n_tiles; ++t) {
        const int b = static_cast<int>(t / tiles_img), ti = static_cast<int>(t % tiles_img), oh0 = (ti / tiles_w) * TW_TH,
                  ow0 = (ti % tiles_w) * TW_TW;
        const float* xb = x + static_cast<size_t>(b) * H * W;
        const float* dyb = dy + static_cast<size_t>(b) * C * Ho * Wo;
        __syncthreads();
        for (int i = threadIdx.x; i < TW_PH * TW_PW; i += 256) {
            const int ih = 2 * oh0 - 2 + i / TW_PW, iw = 2 * ow0 - 2 + i % TW_PW;
            sx[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(xb + ih * W + iw) : 0.0f;
        }
        if (channels_last) {                             // dy is [B, Ho, Wo, C]: channel fastest
            for (int i = threadIdx.x; i < C * TW_TH * TW_TW; i += 256) {
                const int cc = i % C, p = i / C, oh = oh0 + p / TW_TW, ow = ow0 + p % TW_TW;
                sdy[cc][p] = (oh < Ho && ow < Wo) ? __ldg(dyb + (static_cast<size_t>(oh) * Wo + ow) * C + cc) : 0.0f;
            }
        } else {
            for (int i = threadIdx.x; i < C * TW_TH * TW_TW; i += 256) {
                const int cc = i / (TW_TH * TW_TW), p = i % (TW_TH * TW_TW), oh = oh0 + p / TW_TW, ow = ow0 + p % TW_TW;
                sdy[cc][p] = (oh < Ho && ow < Wo) ? __ldg(dyb + (static_cast<size_t>(cc) * Ho + oh) * Wo + ow) : 0.0f;
            }
        }
        __syncthreads();
        if (worker) {
            const float* px = sx + r * TW_PW + s;
#pragma unroll
            for (int py = 0; py < TW_TH; ++py)
#pragma unroll 8
                for (int qx = 0; qx < TW_TW; ++qx) acc = fmaf(sdy[c][py * TW_TW + qx], px[2 * py * TW_PW + 2 * qx], acc);
        }
    }
    if (worker) {
        if (partial != nullptr) partial[static_cast<size_t>(blockIdx.x) * (THIN_MAXC * THIN_TAPS) + c * THIN_TAPS + tap] = acc;
        else atomicAdd(dw + c * THIN_TAPS + tap, acc);
    }
}

// Channels-last, C == 8: thread = (output pixel, channel half), 25 taps x 4 channels = 100 accumulators in registers,
// so a pixel costs 25 conflict-free LDS.32 (its patch) + 1 LDS.128 (its 4 dy values) for 100 FMAs: the FMA pipe, not shared memory,
// is the on-chip bound (a lane = tap mapping needs 3 shared-memory reads per 8 FMAs and idles 7 of 32 lanes), and below both sits
// the HBM stream (x once with a 1.4x halo, dy once).  Tile = 4 rows x 32 pixels; warp w: row w % 4, channel half w / 4.  The patch is
// stored split into even and odd columns so that lanes (consecutive pixels, input columns 2 px + s) read consecutive words.  Tiles are
// fetched WG_STAGES - 1 ahead with cp.async (zero fill outside the image; a 7 KB tile computes in ~0.2 us but takes ~1.5 us to arrive, so
// one tile in flight per CTA is latency bound at 1.4 TB/s) and dealt round-robin to a fixed grid: per-CTA sums in a fixed order.
constexpr int WG_STAGES = 6, WG_TH = 4, WG_TW = 32, WG_PH = 2 * WG_TH + 3, WG_PW = 2 * WG_TW + 3, WG_PITCH = 36, WG_PIX = WG_TH * WG_TW;
struct WgStage {
    float xe[WG_PH][WG_PITCH], xo[WG_PH][WG_PITCH];
    float4 dy[2][WG_PIX];
};

__global__ void __launch_bounds__(256, 2) thin_conv_wgrad_cl8_reg_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                         float* __restrict__ dw, int B, int H, int W, int Ho, int Wo,
                                                                         FastDiv div_img, FastDiv div_w, int n_tiles,
                                                                         float* __restrict__ partial) {
    __shared__ __align__(16) WgStage st[WG_STAGES];
    __shared__ float sred[8][THIN_TAPS * 4];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, half = warp >> 2, py = warp & 3;

    // Per-thread copy plan, fixed for all tiles (the copy issue must stay well below the 126 instructions of the tile's arithmetic):
    // patch elements t, t + 256, t + 512 (row, column -> offset in the image relative to the patch corner, offset in the stage) and
    // the 16-byte half of dy pixel t / 2.
    int x_row[3], x_col[3], x_src[3];
    uint32_t x_dst[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int i = t + 256 * k, row = i / WG_PW, col = i - row * WG_PW;
        x_row[k] = i < WG_PH * WG_PW ? row : -100000;        // (never inside the image)
        x_col[k] = col;
        x_src[k] = row * W + col;
        x_dst[k] = static_cast<uint32_t>(((col & 1) ? offsetof(WgStage, xo) : offsetof(WgStage, xe)) + (row * WG_PITCH + (col >> 1)) * 4);
    }
    const int d_pix = t >> 1, d_h = t & 1, d_oh = d_pix / WG_TW, d_ow = d_pix % WG_TW;
    const uint32_t d_dst = static_cast<uint32_t>(offsetof(WgStage, dy) + (d_h * WG_PIX + d_pix) * 16);
    const uint32_t st_u32 = smem_u32(&st[0]);

    auto issue = [&](int tile, int s) {
        if (tile < n_tiles) {
            uint32_t b, ti, th, tw;
            div_img.divmod(static_cast<uint32_t>(tile), b, ti);
            div_w.divmod(ti, th, tw);
            const int oh0 = static_cast<int>(th) * WG_TH, ow0 = static_cast<int>(tw) * WG_TW, ih0 = 2 * oh0 - 2, iw0 = 2 * ow0 - 2;
            const float* xc = x + static_cast<size_t>(b) * H * W + ih0 * W + iw0;          // patch corner (may lie outside: never dereferenced there)
            const uint32_t base = st_u32 + s * static_cast<uint32_t>(sizeof(WgStage));
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const bool ok = static_cast<unsigned>(ih0 + x_row[k]) < static_cast<unsigned>(H) && static_cast<unsigned>(iw0 + x_col[k]) < static_cast<unsigned>(W);
                if (k < 2 || t + 512 < WG_PH * WG_PW) cp_async4(base + x_dst[k], ok ? xc + x_src[k] : x, ok ? 4u : 0u);
            }
            const int oh = oh0 + d_oh, ow = ow0 + d_ow;
            const bool ok = oh < Ho && ow < Wo;
            cp_async16_cg(base + d_dst, ok ? dy + ((static_cast<size_t>(b) * Ho + oh) * Wo + ow) * THIN_MAXC + 4 * d_h : dy, ok ? 16u : 0u);
        }
        cp_async_commit();
    };

    float acc[THIN_TAPS][4];
#pragma unroll
    for (int k = 0; k < THIN_TAPS; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[k][i] = 0.0f;
    int tile = blockIdx.x, s = 0, s_in = WG_STAGES - 1;
    const int grid = gridDim.x;
#pragma unroll 1
    for (int k = 0; k < WG_STAGES - 1; ++k) issue(tile + k * grid, k);
#pragma unroll 1
    for (; tile < n_tiles; tile += grid) {
        issue(tile + (WG_STAGES - 1) * grid, s_in);          // (that buffer was released by the last tile's barrier)
        cp_async_wait<WG_STAGES - 1>();
        __syncthreads();
        const float4 d = st[s].dy[half][py * WG_TW + lane];
#pragma unroll
        for (int r = 0; r < THIN_K; ++r)
#pragma unroll
            for (int q = 0; q < THIN_K; ++q) {
                const float xv = (q & 1) ? st[s].xo[2 * py + r][lane + (q >> 1)] : st[s].xe[2 * py + r][lane + (q >> 1)];
                float* a = acc[r * THIN_K + q];
                a[0] = fmaf(d.x, xv, a[0]); a[1] = fmaf(d.y, xv, a[1]); a[2] = fmaf(d.z, xv, a[2]); a[3] = fmaf(d.w, xv, a[3]);
            }
        __syncthreads();
        s = s + 1 == WG_STAGES ? 0 : s + 1;
        s_in = s_in + 1 == WG_STAGES ? 0 : s_in + 1;
    }
    cp_async_wait<0>();
#pragma unroll
    for (int k = 0; k < THIN_TAPS; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = acc[k][i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) sred[warp][k * 4 + i] = v;
        }
    __syncthreads();
    if (t < THIN_TAPS * THIN_MAXC) {
        const int c = t / THIN_TAPS, tap = t % THIN_TAPS, h = c >> 2, i = c & 3;
        const float v = ((sred[4 * h][tap * 4 + i] + sred[4 * h + 1][tap * 4 + i]) + sred[4 * h + 2][tap * 4 + i]) + sred[4 * h + 3][tap * 4 + i];
        if (partial != nullptr) partial[static_cast<size_t>(blockIdx.x) * (THIN_MAXC * THIN_TAPS) + t] = v;
        else atomicAdd(dw + t, v);
    }
}

// dw[i] = sum over blocks of partial[block][i], 8 lanes per output adding every 8th block in order, then combined in lane order
__global__ void __launch_bounds__(256) thin_wgrad_finish_kernel(const float* __restrict__ partial, int blocks, int n_out, float* __restrict__ dw) {
    __shared__ float part[8][33];
    const int ex = threadIdx.x & 31, sl = threadIdx.x >> 5, i = blockIdx.x * 32 + ex;
    float acc = 0.0f;
    if (i < n_out)
        for (int b = sl; b < blocks; b += 8) acc += __ldcg(partial + static_cast<size_t>(b) * (THIN_MAXC * THIN_TAPS) + i);
    part[sl][ex] = acc;
    __syncthreads();
    if (sl == 0 && i < n_out) {
        float v = part[0][ex];
#pragma unroll
        for (int l = 1; l < 8; ++l) v += part[l][ex];
        dw[i] = v;
    }
}

// Fills a constant bank: 200 weights from `w`, then `n_bias` biases from `bias` (zeros when NULL), all device-to-device on `st`.
template <typename Sym>
static int thin_load_constants(const Sym& symbol, const float* w, const float* bias, int n_bias, cudaStream_t st) {
    constexpr size_t WB = sizeof(float) * THIN_MAXC * THIN_TAPS;
    PGV_CUDA(cudaMemcpyToSymbolAsync(symbol, w, WB, 0, cudaMemcpyDeviceToDevice, st));
    if (bias != nullptr) {
        PGV_CUDA(cudaMemcpyToSymbolAsync(symbol, bias, sizeof(float) * n_bias, WB, cudaMemcpyDeviceToDevice, st));
    } else {
        void* base = nullptr;
        PGV_CUDA(cudaGetSymbolAddress(&base, symbol));
        PGV_CUDA(cudaMemsetAsync(static_cast<char*>(base) + WB, 0, sizeof(float) * n_bias, st));
    }
    return 0;
}

// CTAs per SM the one-thread-per-pixel forward / transposed kernels are capped at (pgv_debug_set_thin_grid_mult)
static int g_thin_grid_mult = 32;

static bool thin_geometry(int C, int kh, int kw, int stride, int pad, int H, int W, int Ho, int Wo) {
    return C >= 1 && C <= THIN_MAXC && kh == 5 && kw == 5 && stride == 2 && pad == 2 && Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1;
}

}  // namespace pgv

using namespace pgv;

extern "C" {

int pgv_debug_set_thin_grid_mult(int m) { g_thin_grid_mult = m < 1 ? 1 : m; return 0; }

int pgv_conv5x5s2_c1_supported(int Cin, int Cout, int kh, int kw, int stride, int pad, int H, int W, int Ho, int Wo) {
    return Cin == 1 && thin_geometry(Cout, kh, kw, stride, pad, H, W, Ho, Wo);
}

int pgv_conv5x5s2_c1_fwd(const float* x, const float* w, const float* bias, float* y, int B, int C, int H, int W, int Ho, int Wo,
                         float lrelu_slope, int channels_last, int round_out, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && w && y, "pgv_conv5x5s2_c1_fwd: NULL argument");
    PGV_CHECK_ARG(thin_geometry(C, 5, 5, 2, 2, H, W, Ho, Wo) && B > 0, "pgv_conv5x5s2_c1_fwd: unsupported geometry");
    const long long total = static_cast<long long>(B) * Ho * Wo;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * g_thin_grid_mult));
    if (channels_last && C == THIN_MAXC && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (int rc = thin_load_constants(c_thin_fwd_w, w, bias, THIN_MAXC, st)) return rc;
        thin_conv_fwd_cl8_kernel<<<grid, 256, 0, st>>>(x, y, B, H, W, Ho, Wo, lrelu_slope, round_out);
    } else if (channels_last)
        thin_conv_fwd_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, y, B, C, H, W, Ho, Wo, lrelu_slope, round_out);
    else
        thin_conv_fwd_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, y, B, C, H, W, Ho, Wo, lrelu_slope, round_out);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_conv5x5s2_c1_dgrad(const float* y, const float* w, const float* bias, float* x, int B, int C, int H, int W, int Ho, int Wo,
                           float clamp_lo, float clamp_hi, int channels_last, pgv_stream_t stream) {
    PGV_CHECK_ARG(y && w && x, "pgv_conv5x5s2_c1_dgrad: NULL argument");
    PGV_CHECK_ARG(thin_geometry(C, 5, 5, 2, 2, H, W, Ho, Wo) && B > 0, "pgv_conv5x5s2_c1_dgrad: unsupported geometry");
    const long long total = static_cast<long long>(B) * H * W;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * g_thin_grid_mult));
    if (channels_last && C == THIN_MAXC && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (int rc = thin_load_constants(c_thin_quad_w, w, bias, 1, st)) return rc;
        const long long quads = static_cast<long long>(B) * ((H + 1) / 2) * ((W + 1) / 2);
        const int qgrid = static_cast<int>(std::min<long long>((quads + 255) / 256, 148LL * g_thin_grid_mult));
        thin_conv_quad_cl8_kernel<<<qgrid, 256, 0, st>>>(y, x, B, H, W, Ho, Wo, clamp_lo, clamp_hi);
    } else if (channels_last)
        thin_conv_dgrad_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, w, bias, x, B, C, H, W, Ho, Wo, clamp_lo, clamp_hi);
    else
        thin_conv_dgrad_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, w, bias, x, B, C, H, W, Ho, Wo, clamp_lo, clamp_hi);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_conv5x5s2_c1_wgrad(const float* x, const float* dy, float* dw, int B, int C, int H, int W, int Ho, int Wo, int channels_last,
                           void* ws, size_t ws_bytes, pgv_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PGV_CHECK_ARG(x && dy && dw, "pgv_conv5x5s2_c1_wgrad: NULL argument");
    PGV_CHECK_ARG(thin_geometry(C, 5, 5, 2, 2, H, W, Ho, Wo) && B > 0, "pgv_conv5x5s2_c1_wgrad: unsupported geometry");
    const bool reg_form = channels_last && C == THIN_MAXC && (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
    long long n_tiles = static_cast<long long>(B) * ((Ho + TW_TH - 1) / TW_TH) * ((Wo + TW_TW - 1) / TW_TW);
    int per_block = static_cast<int>((n_tiles + 148 * 4 - 1) / (148 * 4));
    if (per_block < 1) per_block = 1;
    int grid = static_cast<int>((n_tiles + per_block - 1) / per_block);
    if (reg_form) {                                   // persistent: two CTAs per SM, tiles dealt round-robin
        n_tiles = static_cast<long long>(B) * ((Ho + WG_TH - 1) / WG_TH) * ((Wo + WG_TW - 1) / WG_TW);
        PGV_CHECK_ARG(n_tiles < (1LL << 30) && static_cast<long long>(B) * H * W < (1LL << 31), "pgv_conv5x5s2_c1_wgrad: too many tiles");
        grid = static_cast<int>(std::min<long long>(n_tiles, 2LL * 148));
    }
    // with a workspace (>= grid x 200 floats) the per-block sums are combined in a fixed order (deterministic); else fp32 atomics
    float* partial = (ws != nullptr && ws_bytes >= static_cast<size_t>(grid) * THIN_MAXC * THIN_TAPS * sizeof(float)) ? static_cast<float*>(ws) : nullptr;
    if (partial == nullptr) PGV_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * C * THIN_TAPS, stream));
    if (reg_form)
    {
        FastDiv div_img, div_w;
        div_img.init(static_cast<uint32_t>(((Ho + WG_TH - 1) / WG_TH) * ((Wo + WG_TW - 1) / WG_TW)));
        div_w.init(static_cast<uint32_t>((Wo + WG_TW - 1) / WG_TW));
        thin_conv_wgrad_cl8_reg_kernel<<<grid, 256, 0, stream>>>(x, dy, dw, B, H, W, Ho, Wo, div_img, div_w, static_cast<int>(n_tiles), partial);
    }
    else
        thin_conv_wgrad_kernel<<<grid, 256, 0, stream>>>(x, dy, dw, B, C, H, W, Ho, Wo, per_block, channels_last, partial);
    PGV_LAUNCH_CHECK();
    if (partial != nullptr) {
        thin_wgrad_finish_kernel<<<ceil_div(C * THIN_TAPS, 32), 256, 0, stream>>>(partial, grid, C * THIN_TAPS, dw);
        PGV_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
