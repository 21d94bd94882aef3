// Implicit-GEMM convolutions on the 5th-generation tensor cores (tcgen05, TF32 products, fp32 accumulation in TMEM).
//
//   mode FWD    y[b,co,oh,ow]  = act(bias[co] + sum_{ci,r,s} x[b,ci,oh*st-p+r,ow*st-p+s] * w[co,ci,r,s])
//               GEMM: M = B*Ho*Wo pixels, N = Cout, K = Cin*kh*kw
//   mode DGRAD  dx[b,ci,ih,iw] = act(bias[ci] + sum_{co,r,s : oh*st-p+r = ih, ...} dy[b,co,oh,ow] * w[co,ci,r,s])
//               = forward of nn.ConvTranspose2d(weight = w).  One GEMM per input-pixel parity class (st x st classes):
//               M = B*Hc*Wc pixels of the class, N = Cin, K = Cout * taps (taps = ceil(kh/st)*ceil(kw/st))
//   mode WGRAD  dw[co,ci,r,s] += sum_{b,oh,ow} dy[b,co,oh,ow] * x[b,ci,oh*st-p+r,ow*st-p+s]
//               GEMM: M = Cin*kh*kw, N = Cout, K = pixels; split along K across CTAs, fp32 atomic accumulation
//
// Neither operand of these GEMMs is a plain matrix in memory (NCHW activations, strided taps, zero padding, weight
// sub-lattices), so both smem tiles are written by 8 PRODUCER warps: each thread gathers 16-byte chunks (4 consecutive
// k of one row), rounds them to TF32 (cvt.rna) and stores them at the 128-byte-swizzled position the UMMA descriptor
// expects, then fence.proxy.async + mbarrier arrive.  One thread issues tcgen05.mma (M=128, N = tile width, K=8);
// accumulators are double-buffered in TMEM so the 4 epilogue warps drain tile i while tile i+1 is being multiplied.
#include <string.h>

#include "pgv_common.cuh"
#include "pgv_tc.cuh"

namespace pgv {

enum { CONV_FWD = 0, CONV_DGRAD = 1, CONV_WGRAD = 2, DENSE_WGRAD = 3 };

constexpr int CT_BLOCK_M = 128, CT_BLOCK_K = 32, CT_MAX_N = 128, CT_STAGES = 6;
constexpr int CT_A_BYTES = CT_BLOCK_M * 128, CT_B_BYTES = CT_MAX_N * 128, CT_STAGE_BYTES = CT_A_BYTES + CT_B_BYTES;
constexpr int CT_PRODUCER_WARPS = 8, CT_PRODUCERS = CT_PRODUCER_WARPS * 32;
constexpr int CT_THREADS = CT_PRODUCERS + 32 /*mma*/ + 128 /*epilogue*/;
constexpr int CT_SMEM = 1024 + CT_STAGES * CT_STAGE_BYTES + 256;

struct ConvTcParams {
    const float* x;      // FWD: input x        DGRAD: dy (conv output side)    WGRAD: x
    const float* w;      // weights [Cout, Cin, kh, kw]                          WGRAD: dy
    const float* bias;   // FWD: [Cout], DGRAD: [Cin] or NULL
    const float* residual;   // FWD only: same shape as out, added before the activation (or NULL)
    float* out;          // FWD: y              DGRAD: dx                        WGRAD: dw (pre-zeroed)
    int B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo;
    int n_tile, n_tiles;            // tile width (multiple of 16, <= 256) and count along N
    int gemm_n, gemm_k;             // logical N and K of one GEMM
    int kb_total, kb_per_split, k_splits;
    int classes;                    // DGRAD: stride*stride parity classes, else 1
    int m_tiles_class[4];           // M tiles per class (FWD / WGRAD use [0])
    int Hc[4], Wc[4];               // DGRAD: sub-grid size of every class
    int taps_h, taps_w;             // DGRAD: taps per class along h / w
    int pix_blocks;                 // WGRAD: k-blocks per image = ceil(Ho*Wo / 32)
    int fast;                       // 1: 4x4 kernel (FWD) / 2x2 taps (DGRAD); 2: 1x1 kernel; 0: generic
    // slot model (see Producer): element e of a chunk is at ptr + off[e-1]; ptr advances by step per k-block
    int a_pair;                     // 1: FWD 4x4/s2, 2: quad DGRAD -> adjacent lanes share two of the four chunk elements
    int slot_a, slot_b;             // 1: operand uses the slot model, 0: generic chunk_a / chunk_b gathers
    int a_off[3], b_off[3], a_vec, b_vec, a_kdim, b_kdim;   // kdim: valid length along k (tail masking)
    long long a_step, b_step;
    int quad;                       // DGRAD 4x4 / stride 2 / even pad: the 4 parity classes of a 2x2 pixel quad share one A row
                                    // (same 2x2 dy patch) and are stacked along N: column n = class * Cin + ci
    int sshift;                     // log2(stride) (stride is 1 or 2 on every tensor-core path)
    long long* trace;               // debug: clock64 timestamps of CTA 0 (NULL in production)
    int atomic_out;                 // FWD / DGRAD with k_splits > 1: accumulate into a pre-zeroed output, split 0 adds the bias
    int a_dense;                    // A is a plain row-major [M, K] matrix (Linear layers): K-contiguous vector loads
    float slope;
    FastDiv fd_Cin;
    FastDiv fd_HWo, fd_Wo, fd_taps, fd_kw, fd_dtaps, fd_dtapsw, fd_pixblocks, fd_ntile, fd_HcWc[4], fd_Wc[4];
};

__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
    return static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void st_chunk(uint8_t* tile, int row, int chunk, const float (&v)[4]) {
    *reinterpret_cast<float4*>(tile + sw128_offset(row, chunk)) =
        make_float4(to_tf32_rna(v[0]), to_tf32_rna(v[1]), to_tf32_rna(v[2]), to_tf32_rna(v[3]));
}

struct WorkItem { int cls, tm, tn, kb0, kb1; };

__device__ __forceinline__ int total_m_tiles(const ConvTcParams& p) {
    int t = 0;
    for (int c = 0; c < p.classes; ++c) t += p.m_tiles_class[c];
    return t;
}
__device__ __forceinline__ WorkItem decode_item(const ConvTcParams& p, int item) {
    WorkItem wi;
    const int split = item % p.k_splits;
    int t = item / p.k_splits;
    wi.tn = t % p.n_tiles;
    t /= p.n_tiles;
    wi.cls = 0;
    while (wi.cls < p.classes - 1 && t >= p.m_tiles_class[wi.cls]) { t -= p.m_tiles_class[wi.cls]; ++wi.cls; }
    wi.tm = t;
    wi.kb0 = split * p.kb_per_split;
    wi.kb1 = min(p.kb_total, wi.kb0 + p.kb_per_split);
    return wi;
}

// ------------------------------------------------------------------------------------------------ operand gathers
// Every gather produces 4 consecutive k values of one row (k0 = 4 * global chunk index).  The per-row state is computed
// once per tile (each producer thread owns one A row), so the per-chunk work is a few integer ops and predicated loads.
struct RowA {
    const float* base;
    int i0, j0;          // FWD: ih0, iw0   DGRAD: oh, ow of tap 0   WGRAD: r - pad, s - pad
    uint32_t hmask, wmask;   // validity of tap rows / columns
    bool ok;
};

template <int MODE>
__device__ __forceinline__ RowA row_a(const ConvTcParams& p, const WorkItem& wi, int row) {
    RowA r;
    r.base = nullptr; r.i0 = r.j0 = 0; r.hmask = r.wmask = 0; r.ok = false;
    if (MODE == CONV_FWD) {
        const uint32_t m = static_cast<uint32_t>(wi.tm) * CT_BLOCK_M + row;
        uint32_t b, pix, oh, ow;
        p.fd_HWo.divmod(m, b, pix);
        r.ok = b < static_cast<uint32_t>(p.B);
        if (r.ok) {
            p.fd_Wo.divmod(pix, oh, ow);
            r.i0 = static_cast<int>(oh) * p.stride - p.pad;
            r.j0 = static_cast<int>(ow) * p.stride - p.pad;
            r.base = p.x + static_cast<size_t>(b) * p.Cin * p.H * p.W;
            for (int t = 0; t < p.kh; ++t) if (r.i0 + t >= 0 && r.i0 + t < p.H) r.hmask |= 1u << t;
            for (int t = 0; t < p.kw; ++t) if (r.j0 + t >= 0 && r.j0 + t < p.W) r.wmask |= 1u << t;
        }
    } else if (MODE == CONV_DGRAD) {
        const uint32_t m = static_cast<uint32_t>(wi.tm) * CT_BLOCK_M + row;
        uint32_t b, pix, i, j;
        p.fd_HcWc[wi.cls].divmod(m, b, pix);
        r.ok = b < static_cast<uint32_t>(p.B);
        if (r.ok) {
            p.fd_Wc[wi.cls].divmod(pix, i, j);
            const int ph = wi.cls / p.stride, pw = wi.cls % p.stride;
            const int ih = static_cast<int>(i) * p.stride + ph, iw = static_cast<int>(j) * p.stride + pw;
            // tap (a, c) reads dy at oh = (ih + pad)/stride - a, kernel row r0 + a*stride with r0 = (ih + pad) % stride
            r.i0 = (ih + p.pad) / p.stride;
            r.j0 = (iw + p.pad) / p.stride;
            r.base = p.x + static_cast<size_t>(b) * p.Cout * p.Ho * p.Wo;
            for (int t = 0; t < p.taps_h; ++t) if (r.i0 - t >= 0 && r.i0 - t < p.Ho) r.hmask |= 1u << t;
            for (int t = 0; t < p.taps_w; ++t) if (r.j0 - t >= 0 && r.j0 - t < p.Wo) r.wmask |= 1u << t;
        }
    } else if (MODE == CONV_WGRAD) {   // row = (ci, r, s)
        const uint32_t m = static_cast<uint32_t>(wi.tm) * CT_BLOCK_M + row;
        r.ok = m < static_cast<uint32_t>(p.Cin * p.kh * p.kw);
        if (r.ok) {
            uint32_t ci, t, rr, ss;
            p.fd_taps.divmod(m, ci, t);
            p.fd_kw.divmod(t, rr, ss);
            r.i0 = static_cast<int>(rr) - p.pad;
            r.j0 = static_cast<int>(ss) - p.pad;
            r.base = p.x + static_cast<size_t>(ci) * p.H * p.W;
        }
    } else {                            // DENSE_WGRAD: row = input feature ci
        const int m = wi.tm * CT_BLOCK_M + row;
        r.ok = m < p.Cin;
        r.base = p.x + m;
    }
    return r;
}

template <int MODE>
__device__ __forceinline__ void chunk_a(const ConvTcParams& p, const WorkItem& wi, const RowA& r, int gchunk, float (&v)[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (!r.ok) return;
    const int k0 = gchunk * 4;
    if (MODE == CONV_FWD) {
        if (p.fast == 1) {                           // 4x4 kernel: chunk = one (ci, r), 4 consecutive input pixels
            const int ci = k0 >> 4, rr = (k0 >> 2) & 3;
            if (ci < p.Cin && ((r.hmask >> rr) & 1u)) {
                const float* src = r.base + (static_cast<size_t>(ci) * p.H + (r.i0 + rr)) * p.W + r.j0;
#pragma unroll
                for (int s = 0; s < 4; ++s) if ((r.wmask >> s) & 1u) v[s] = __ldg(src + s);
            }
        } else if (p.fast == 2) {                    // 1x1 kernel: chunk = 4 consecutive input channels of one pixel
            if (r.hmask & r.wmask & 1u) {
                const size_t hw = static_cast<size_t>(p.H) * p.W;
                const float* src = r.base + static_cast<size_t>(r.i0) * p.W + r.j0 + static_cast<size_t>(k0) * hw;
#pragma unroll
                for (int e = 0; e < 4; ++e) if (k0 + e < p.Cin) v[e] = __ldg(src + e * hw);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t k = k0 + e;
                if (k < static_cast<uint32_t>(p.gemm_k)) {
                    uint32_t ci, t, rr, ss;
                    p.fd_taps.divmod(k, ci, t);
                    p.fd_kw.divmod(t, rr, ss);
                    if (((r.hmask >> rr) & 1u) && ((r.wmask >> ss) & 1u))
                        v[e] = r.base[(static_cast<size_t>(ci) * p.H + (r.i0 + rr)) * p.W + r.j0 + ss];
                }
            }
        }
    } else if (MODE == CONV_DGRAD) {
        const int HWo = p.Ho * p.Wo;
        if (p.fast == 1) {                           // 2x2 taps: chunk = one output channel co
            const int co = k0 >> 2;
            if (co < p.Cout) {
                const float* src = r.base + static_cast<size_t>(co) * HWo + r.i0 * p.Wo + r.j0;
                const bool h0 = r.hmask & 1u, h1 = r.hmask & 2u, w0 = r.wmask & 1u, w1 = r.wmask & 2u;
                if (h0 && w0) v[0] = __ldg(src);
                if (h0 && w1) v[1] = __ldg(src - 1);
                if (h1 && w0) v[2] = __ldg(src - p.Wo);
                if (h1 && w1) v[3] = __ldg(src - p.Wo - 1);
            }
        } else if (p.fast == 2) {                    // single tap: chunk = 4 consecutive output channels
            if (r.hmask & r.wmask & 1u) {
                const float* src = r.base + static_cast<size_t>(k0) * HWo + r.i0 * p.Wo + r.j0;
#pragma unroll
                for (int e = 0; e < 4; ++e) if (k0 + e < p.Cout) v[e] = __ldg(src + static_cast<size_t>(e) * HWo);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t k = k0 + e;
                if (k < static_cast<uint32_t>(p.gemm_k)) {
                    uint32_t co, t, a, c;
                    p.fd_dtaps.divmod(k, co, t);
                    p.fd_dtapsw.divmod(t, a, c);
                    if (((r.hmask >> a) & 1u) && ((r.wmask >> c) & 1u))
                        v[e] = r.base[static_cast<size_t>(co) * HWo + (r.i0 - static_cast<int>(a)) * p.Wo + r.j0 - static_cast<int>(c)];
                }
            }
        }
    } else if (MODE == CONV_WGRAD) {   // k = pixel index inside image b
        uint32_t b, pb, oh, ow;
        p.fd_pixblocks.divmod(static_cast<uint32_t>(gchunk >> 3), b, pb);
        const int pix0 = static_cast<int>(pb) * CT_BLOCK_K + (gchunk & 7) * 4, HWo = p.Ho * p.Wo;
        if (pix0 >= HWo) return;
        const float* src = r.base + static_cast<size_t>(b) * p.Cin * p.H * p.W;
        p.fd_Wo.divmod(static_cast<uint32_t>(pix0), oh, ow);
        int ih = static_cast<int>(oh) * p.stride + r.i0, iw = static_cast<int>(ow) * p.stride + r.j0, col = static_cast<int>(ow);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (pix0 + e < HWo && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) v[e] = __ldg(src + ih * p.W + iw);
            iw += p.stride;
            if (++col == p.Wo) { col = 0; ih += p.stride; iw = r.j0; }
        }
    } else {                            // DENSE_WGRAD: k = batch row
#pragma unroll
        for (int e = 0; e < 4; ++e) if (k0 + e < p.B) v[e] = __ldg(r.base + static_cast<size_t>(k0 + e) * p.Cin);
    }
}

// B operand: row = output column n
template <int MODE>
__device__ __forceinline__ void chunk_b(const ConvTcParams& p, const WorkItem& wi, int row, int gchunk, float (&v)[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    const int n = wi.tn * p.n_tile + row;
    if (n >= p.gemm_n) return;
    const int k0 = gchunk * 4;
    if (MODE == CONV_FWD) {                          // w[co = n, k] contiguous in k
        const float* src = p.w + static_cast<size_t>(n) * p.gemm_k + k0;
        if (k0 + 4 <= p.gemm_k && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(src));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (k0 + e < p.gemm_k) v[e] = __ldg(src + e);
        }
    } else if (MODE == CONV_DGRAD) {                 // w[co, ci = n, r0 + a*stride, s0 + c*stride]
        const int khw = p.kh * p.kw;
        const int ph = wi.cls / p.stride, pw = wi.cls % p.stride;
        const int r0 = (ph + p.pad) % p.stride, s0 = (pw + p.pad) % p.stride;
        if (p.fast == 1) {
            const int co = k0 >> 2;
            if (co < p.Cout) {
                const float* src = p.w + (static_cast<size_t>(co) * p.Cin + n) * khw + r0 * p.kw + s0;
                v[0] = __ldg(src); v[1] = __ldg(src + p.stride);
                v[2] = __ldg(src + p.stride * p.kw); v[3] = __ldg(src + p.stride * p.kw + p.stride);
            }
        } else if (p.fast == 2) {
            const float* src = p.w + static_cast<size_t>(k0) * p.Cin + n;
#pragma unroll
            for (int e = 0; e < 4; ++e) if (k0 + e < p.Cout) v[e] = __ldg(src + static_cast<size_t>(e) * p.Cin);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t k = k0 + e;
                if (k < static_cast<uint32_t>(p.gemm_k)) {
                    uint32_t co, t, a, c;
                    p.fd_dtaps.divmod(k, co, t);
                    p.fd_dtapsw.divmod(t, a, c);
                    const int rr = r0 + static_cast<int>(a) * p.stride, ss = s0 + static_cast<int>(c) * p.stride;
                    if (rr < p.kh && ss < p.kw) v[e] = __ldg(p.w + (static_cast<size_t>(co) * p.Cin + n) * khw + rr * p.kw + ss);
                }
            }
        }
    } else if (MODE == CONV_WGRAD) {                 // dy[b, co = n, pixel]
        uint32_t b, pb;
        p.fd_pixblocks.divmod(static_cast<uint32_t>(gchunk >> 3), b, pb);
        const int pix0 = static_cast<int>(pb) * CT_BLOCK_K + (gchunk & 7) * 4, HWo = p.Ho * p.Wo;
        if (pix0 >= HWo) return;
        const float* src = p.w + (static_cast<size_t>(b) * p.Cout + n) * HWo + pix0;
        if (pix0 + 4 <= HWo && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(src));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (pix0 + e < HWo) v[e] = __ldg(src + e);
        }
    } else {                                         // DENSE_WGRAD: dy[b, co = n]
#pragma unroll
        for (int e = 0; e < 4; ++e) if (k0 + e < p.B) v[e] = __ldg(p.w + static_cast<size_t>(k0 + e) * p.Cout + n);
    }
}

// trace layout: [role][event_index][field]; role 0 producer thread 0, 1 MMA thread, 2 epilogue thread
__device__ __forceinline__ void trace_evt(const ConvTcParams& p, int role, int& n, int field, long long v) {
    if (p.trace != nullptr && blockIdx.x == 0 && n < 64) p.trace[(role * 64 + n) * 8 + field] = v;
}

__device__ __forceinline__ float act_lrelu(float v, float slope) { return (slope >= 0.0f && v < 0.0f) ? v * slope : v; }

template <int MODE>
__global__ void __launch_bounds__(CT_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + CT_STAGES * CT_STAGE_BYTES);
    uint64_t* bar_empty = bar_full + CT_STAGES;
    uint64_t* bar_tfull = bar_empty + CT_STAGES;
    uint64_t* bar_tempty = bar_tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int MMA_WARP = CT_PRODUCER_WARPS, EPI_WARP0 = CT_PRODUCER_WARPS + 1;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < CT_STAGES; ++s) { mbar_init(&bar_full[s], CT_PRODUCERS); mbar_init(&bar_empty[s], 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 4); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 2 * CT_MAX_N);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_items = total_m_tiles(p) * p.n_tiles * p.k_splits;

    if (warp < CT_PRODUCER_WARPS) {
        // ===================== producers: gather A (128 rows) and B (n_tile <= 128 rows), 8 chunks per row per k-block.
        // Each thread owns 4 A slots and up to 4 B slots (a slot = one 16-byte chunk position of the tile).  For the
        // common layer shapes a slot is {pointer, validity mask, swizzled smem offset}, set up once per tile; a
        // k-block then costs predicated loads + one pointer increment.  The loads of k-block i+1 are issued into a
        // second register set before k-block i is converted and stored, and the (tile, k-block) sequence is
        // flattened so the prefetch runs across tile boundaries.
        const int t = threadIdx.x;                   // 0 .. 255
        const float* ap[4]; uint32_t am[4];          // am / bm: mask (bits 0-3) | chunk (bits 4-6) | smem offset << 8
        const float* bp[4]; uint32_t bm[4];
        RowA ra;                                      // generic-gather state
        int b_row[4];
        bool has_left = false;                       // lane-1 holds the horizontally preceding pixel / quad of the same image row

        auto init_slots = [&](const WorkItem& wi) {
            const int a_row = t & 127, a_c0 = t >> 7;
            const int smask = p.stride - 1;            // stride is a power of two: divisions become shifts / masks
            // ---- row-level state, shared by the 4 A slots of this thread (non-dense operands) ----
            bool row_ok = false;
            const float* rbase = p.x;
            int r_i = 0, r_j = 0;                       // FWD: ih0, iw0      DGRAD: oh0, ow0
            uint32_t hm = 0, wm = 0;
            if (!p.slot_a) {
                ra = row_a<MODE>(p, wi, a_row);
            } else if (!p.a_dense && (MODE == CONV_FWD || MODE == CONV_DGRAD)) {
                const uint32_t m = static_cast<uint32_t>(wi.tm) * CT_BLOCK_M + a_row;
                uint32_t b, pix, i, jj;
                if (MODE == CONV_FWD) {
                    p.fd_HWo.divmod(m, b, pix);
                    row_ok = b < static_cast<uint32_t>(p.B);
                    p.fd_Wo.divmod(pix, i, jj);
                    r_i = (static_cast<int>(i) << p.sshift) - p.pad;
                    r_j = (static_cast<int>(jj) << p.sshift) - p.pad;
                    rbase = p.x + static_cast<size_t>(b) * p.Cin * p.H * p.W;
                    for (int e = 0; e < 4; ++e) {
                        if (r_i + e >= 0 && r_i + e < p.H) hm |= 1u << e;
                        if (r_j + e >= 0 && r_j + e < p.W) wm |= 1u << e;
                    }
                    has_left = row_ok && jj > 0 && (t & 31) != 0;
                } else {
                    p.fd_HcWc[wi.cls].divmod(m, b, pix);
                    row_ok = b < static_cast<uint32_t>(p.B);
                    p.fd_Wc[wi.cls].divmod(pix, i, jj);
                    const int ih = (static_cast<int>(i) << p.sshift) + (wi.cls >> p.sshift), iw = (static_cast<int>(jj) << p.sshift) + (wi.cls & smask);
                    r_i = (ih + p.pad) >> p.sshift;
                    r_j = (iw + p.pad) >> p.sshift;
                    rbase = p.x + static_cast<size_t>(b) * p.Cout * p.Ho * p.Wo;
                    for (int e = 0; e < 2; ++e) {
                        if (r_i - e >= 0 && r_i - e < p.Ho) hm |= 1u << e;
                        if (r_j - e >= 0 && r_j - e < p.Wo) wm |= 1u << e;
                    }
                    has_left = row_ok && jj > 0 && (t & 31) != 0 && p.quad;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int row = a_row, c = a_c0 + 2 * j;
                if (p.a_dense) { const int id = t + j * CT_PRODUCERS; row = id >> 3; c = id & 7; }
                uint32_t mask = 0;
                const float* ptr = p.x;
                if (p.slot_a) {
                    const int g0 = wi.kb0 * 8 + c;                 // first global chunk of this slot
                    if ((MODE == CONV_FWD || MODE == CONV_DGRAD) && p.a_dense) {
                        const int m = wi.tm * CT_BLOCK_M + row;
                        if (m < p.B) { mask = 15; ptr = p.x + static_cast<size_t>(m) * p.gemm_k + g0 * 4; }
                    } else if (MODE == CONV_FWD) {
                        if (row_ok) {
                            if (p.fast == 1) {                      // 4x4: chunk = (ci, r); elements = 4 consecutive iw
                                const int ci = g0 >> 2, r = g0 & 3;
                                if ((hm >> r) & 1u) mask = wm;
                                ptr = rbase + (static_cast<long long>(ci) * p.H + r_i + r) * p.W + r_j;
                            } else {                                // 1x1: chunk = 4 consecutive ci at one pixel
                                mask = 15;
                                ptr = rbase + static_cast<size_t>(g0) * 4 * p.H * p.W + r_i * p.W + r_j;
                            }
                        }
                    } else if (MODE == CONV_DGRAD) {
                        if (row_ok) {
                            if (p.fast == 1) {                      // 2x2 taps: (oh0,ow0), (oh0,ow0-1), (oh0-1,ow0), (oh0-1,ow0-1)
                                const bool h0 = hm & 1u, h1 = hm & 2u, w0 = wm & 1u, w1 = wm & 2u;
                                mask = (h0 && w0 ? 1u : 0u) | (h0 && w1 ? 2u : 0u) | (h1 && w0 ? 4u : 0u) | (h1 && w1 ? 8u : 0u);
                                ptr = rbase + static_cast<long long>(g0) * p.Ho * p.Wo + r_i * p.Wo + r_j;
                            } else {                                // single tap: chunk = 4 consecutive co
                                if ((hm & 1u) && (wm & 1u)) mask = 15;
                                ptr = rbase + static_cast<long long>(g0) * 4 * p.Ho * p.Wo + r_i * p.Wo + r_j;
                            }
                        }
                    } else if (MODE == DENSE_WGRAD) {
                        const int ci = wi.tm * CT_BLOCK_M + row;
                        if (ci < p.Cin) { mask = 15; ptr = p.x + ci + static_cast<size_t>(g0) * 4 * p.Cin; }
                    }
                }
                ap[j] = ptr;
                am[j] = mask | (static_cast<uint32_t>(c) << 4) | (sw128_offset(row, c) << 8);
            }
            const int b_chunks = p.n_tile * 8;
            const int r0s0 = (MODE == CONV_DGRAD) ? ((((wi.cls >> p.sshift) + p.pad) & smask) * p.kw + (((wi.cls & smask) + p.pad) & smask)) : 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int id = t + q * CT_PRODUCERS;
                uint32_t mask = 0; int row = -1, c = 0;
                const float* ptr = p.w;
                if (id < b_chunks) {
                    // operands that are contiguous along k (FWD weights, WGRAD dy): 8 consecutive threads read one 128-byte row
                    // segment; otherwise consecutive threads take consecutive rows (contiguous along n)
                    if (MODE == CONV_FWD || MODE == CONV_WGRAD) { row = id >> 3; c = id & 7; }
                    else { uint32_t qq, rem; p.fd_ntile.divmod(static_cast<uint32_t>(id), qq, rem); row = static_cast<int>(rem); c = static_cast<int>(qq); }
                    const int n = wi.tn * p.n_tile + row, g0 = wi.kb0 * 8 + c;
                    if (p.slot_b && n < p.gemm_n) {
                        mask = 15;
                        if (MODE == CONV_FWD) ptr = p.w + static_cast<size_t>(n) * p.gemm_k + g0 * 4;
                        else if (MODE == CONV_DGRAD) {
                            if (p.quad) {              // n = class * Cin + ci ; class (ph, pw) uses kernel rows ph, ph+2 and columns pw, pw+2
                                uint32_t cls, ci;
                                p.fd_Cin.divmod(static_cast<uint32_t>(n), cls, ci);
                                ptr = p.w + (static_cast<size_t>(g0) * p.Cin + ci) * (p.kh * p.kw) + (cls >> 1) * p.kw + (cls & 1);
                            } else if (p.fast == 1) ptr = p.w + (static_cast<size_t>(g0) * p.Cin + n) * (p.kh * p.kw) + r0s0;
                            else ptr = p.w + static_cast<size_t>(g0) * 4 * p.Cin + n;
                        } else if (MODE == DENSE_WGRAD) ptr = p.w + n + static_cast<size_t>(g0) * 4 * p.Cout;
                    }
                }
                b_row[q] = row;
                bp[q] = ptr;
                bm[q] = mask | (static_cast<uint32_t>(c) << 4) | ((row >= 0 ? sw128_offset(row, c) : 0u) << 8);
            }
        };

        // one slot: predicated loads of the 4 elements (vector load when contiguous, aligned and fully valid)
        auto load_slot = [&](const float*& ptr, uint32_t meta, int kb, const int (&off)[3], int vec, int kdim, long long step, float (&v)[4]) {
            uint32_t mask = meta & 15u;
            const int rem = kdim - (kb * 8 + static_cast<int>((meta >> 4) & 7u)) * 4;     // valid elements left along k
            if (rem < 4) mask &= rem <= 0 ? 0u : ((1u << rem) - 1u);
            v[0] = v[1] = v[2] = v[3] = 0.f;
            if (vec && mask == 15u && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0)) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(ptr));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
                if (mask & 1u) v[0] = __ldg(ptr);
                if (mask & 2u) v[1] = __ldg(ptr + off[0]);
                if (mask & 4u) v[2] = __ldg(ptr + off[1]);
                if (mask & 8u) v[3] = __ldg(ptr + off[2]);
            }
            ptr += step;
        };
        auto load_kb = [&](const WorkItem& wi, int kb, float (&va)[4][4], float (&vb)[4][4]) {
            if (p.slot_a && p.a_pair) {
                // 4x4 / stride-2 windows of horizontally adjacent pixels overlap by two elements: every lane loads only
                // its two NEW elements and takes the other two from lane-1 (warp shuffle); lanes without a left
                // neighbour in the same image row load all four.  Halves the LSU work of the A gather.
                const uint32_t own = (p.a_pair == 1) ? 0xCu : 0x5u;      // FWD: elements 2,3   DGRAD: elements 0,2
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t mask = am[j] & 15u;
                    const int rem = p.a_kdim - (kb * 8 + static_cast<int>((am[j] >> 4) & 7u)) * 4;
                    if (rem < 4) mask &= rem <= 0 ? 0u : ((1u << rem) - 1u);
                    const uint32_t ld = has_left ? (mask & own) : mask;
                    const float* ptr = ap[j];
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                    if (ld & 1u) v0 = __ldg(ptr);
                    if (ld & 2u) v1 = __ldg(ptr + p.a_off[0]);
                    if (ld & 4u) v2 = __ldg(ptr + p.a_off[1]);
                    if (ld & 8u) v3 = __ldg(ptr + p.a_off[2]);
                    if (p.a_pair == 1) {       // FWD: my (e0, e1) = left neighbour's (e2, e3)
                        const float n2 = __shfl_up_sync(0xffffffffu, v2, 1), n3 = __shfl_up_sync(0xffffffffu, v3, 1);
                        if (has_left) { v0 = n2; v1 = n3; }
                    } else {                   // quad DGRAD: my (e1, e3) = left neighbour's (e0, e2)
                        const float n0 = __shfl_up_sync(0xffffffffu, v0, 1), n2 = __shfl_up_sync(0xffffffffu, v2, 1);
                        if (has_left) { v1 = n0; v3 = n2; }
                    }
                    va[j][0] = v0; va[j][1] = v1; va[j][2] = v2; va[j][3] = v3;
                    ap[j] = ptr + p.a_step;
                }
            } else if (p.slot_a) {
#pragma unroll
                for (int j = 0; j < 4; ++j) load_slot(ap[j], am[j], kb, p.a_off, p.a_vec, p.a_kdim, p.a_step, va[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) chunk_a<MODE>(p, wi, ra, kb * 8 + static_cast<int>((am[j] >> 4) & 7u), va[j]);
            }
            if (p.slot_b) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (b_row[q] >= 0) load_slot(bp[q], bm[q], kb, p.b_off, p.b_vec, p.b_kdim, p.b_step, vb[q]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (b_row[q] >= 0) chunk_b<MODE>(p, wi, b_row[q], kb * 8 + static_cast<int>((bm[q] >> 4) & 7u), vb[q]);
            }
        };

        int stage = 0; uint32_t phase = 0;
        int item = blockIdx.x;
        int trace_n = 0;
        if (item < n_items) {
            WorkItem wi = decode_item(p, item);
            int kb = wi.kb0;
            init_slots(wi);
            float va[4][4], vb[4][4], na[4][4], nb[4][4];
            uint32_t cur_am[4], cur_bm[4];
            load_kb(wi, kb, va, vb);
#pragma unroll
            for (int j = 0; j < 4; ++j) { cur_am[j] = am[j]; cur_bm[j] = bm[j]; }
            bool cur_b_valid[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) cur_b_valid[q] = b_row[q] >= 0;
            while (true) {
                // advance the issue cursor and prefetch the next k-block (possibly of the next tile)
                bool have_next = true;
                int nkb = kb + 1;
                WorkItem nwi = wi;
                if (nkb >= wi.kb1) {
                    item += gridDim.x;
                    if (item < n_items) { nwi = decode_item(p, item); nkb = nwi.kb0; init_slots(nwi); }
                    else have_next = false;
                }
                long long t_a = 0, t_b = 0, t_c = 0, t_d = 0, t_e = 0;
                if (p.trace && t == 0) t_a = clock64();
                if (have_next) load_kb(nwi, nkb, na, nb);
                if (p.trace && t == 0) t_b = clock64();
                // store the current k-block
                mbar_wait(&bar_empty[stage], phase ^ 1);
                if (p.trace && t == 0) t_c = clock64();
                uint8_t* sA = smem + stage * CT_STAGE_BYTES;
                uint8_t* sB = sA + CT_A_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(sA + (cur_am[j] >> 8)) =
                        make_float4(to_tf32_rna(va[j][0]), to_tf32_rna(va[j][1]), to_tf32_rna(va[j][2]), to_tf32_rna(va[j][3]));
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (cur_b_valid[q])
                        *reinterpret_cast<float4*>(sB + (cur_bm[q] >> 8)) =
                            make_float4(to_tf32_rna(vb[q][0]), to_tf32_rna(vb[q][1]), to_tf32_rna(vb[q][2]), to_tf32_rna(vb[q][3]));
                if (p.trace && t == 0) t_d = clock64();
                fence_proxy_async_smem();
                if (p.trace && t == 0) t_e = clock64();
                mbar_arrive(&bar_full[stage]);
                if (p.trace && t == 0) {
                    trace_evt(p, 0, trace_n, 0, t_a); trace_evt(p, 0, trace_n, 1, t_b); trace_evt(p, 0, trace_n, 2, t_c);
                    trace_evt(p, 0, trace_n, 3, t_d); trace_evt(p, 0, trace_n, 4, t_e); trace_evt(p, 0, trace_n, 5, clock64());
                    trace_evt(p, 0, trace_n, 6, kb); ++trace_n;
                }
                if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                if (!have_next) break;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    cur_am[j] = am[j]; cur_bm[j] = bm[j]; cur_b_valid[j] = b_row[j] >= 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e) { va[j][e] = na[j][e]; vb[j][e] = nb[j][e]; }
                }
                wi = nwi; kb = nkb;
            }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(CT_BLOCK_M, p.n_tile);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            int trace_n = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const WorkItem wi = decode_item(p, item);
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * CT_MAX_N;
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    const long long t_a = p.trace ? clock64() : 0;
                    mbar_wait(&bar_full[stage], phase);
                    if (p.trace) { trace_evt(p, 1, trace_n, 0, t_a); trace_evt(p, 1, trace_n, 1, clock64()); trace_evt(p, 1, trace_n, 6, kb); }
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(smem + stage * CT_STAGE_BYTES);
                    const uint64_t da = umma_smem_desc_sw128(a_addr), db = umma_smem_desc_sw128(a_addr + CT_A_BYTES);
#pragma unroll
                    for (int k = 0; k < CT_BLOCK_K / 8; ++k) umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > wi.kb0 || k > 0) ? 1u : 0u);
                    umma_commit(&bar_empty[stage]);
                    if (p.trace) { trace_evt(p, 1, trace_n, 2, clock64()); ++trace_n; }
                    if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bar_tfull[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global
        const int quad = warp & 3, row = quad * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        int etrace_n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const WorkItem wi = decode_item(p, item);
            // destination of this thread's row
            float* dst = nullptr;      // address of column n = 0 of the tile; columns are `col_stride` apart
            size_t col_stride = 0;
            bool row_ok = false, q_h1 = false, q_w1 = false;
            if (MODE == CONV_FWD) {
                const uint32_t m = static_cast<uint32_t>(wi.tm) * CT_BLOCK_M + row;
                const int HW = p.Ho * p.Wo;
                uint32_t b, pix;
                p.fd_HWo.divmod(m, b, pix);
                row_ok = b < static_cast<uint32_t>(p.B);
                if (row_ok) {
                    dst = p.out + (static_cast<size_t>(b) * p.Cout + wi.tn * p.n_tile) * HW + pix;
                    col_stride = HW;
                }
            } else if (MODE == CONV_DGRAD) {
                const uint32_t m = static_cast<uint32_t>(wi.tm) * CT_BLOCK_M + row;
                uint32_t b, pix, i, j;
                p.fd_HcWc[wi.cls].divmod(m, b, pix);
                row_ok = b < static_cast<uint32_t>(p.B);
                if (row_ok) {
                    p.fd_Wc[wi.cls].divmod(pix, i, j);
                    const int ih = (static_cast<int>(i) << p.sshift) + (wi.cls >> p.sshift), iw = (static_cast<int>(j) << p.sshift) + (wi.cls & (p.stride - 1));
                    if (p.quad) {                  // dst = pixel (2i, 2j) of channel 0; class / channel offsets are added per column
                        dst = p.out + (static_cast<size_t>(b) * p.Cin * p.H + ih) * p.W + iw;
                        q_h1 = ih + 1 < p.H; q_w1 = iw + 1 < p.W;
                    } else {
                        dst = p.out + ((static_cast<size_t>(b) * p.Cin + wi.tn * p.n_tile) * p.H + ih) * p.W + iw;
                    }
                    col_stride = static_cast<size_t>(p.H) * p.W;
                }
            } else {
                const int m = wi.tm * CT_BLOCK_M + row, Kc = p.Cin * p.kh * p.kw;
                row_ok = m < Kc;
                if (row_ok) { dst = p.out + static_cast<size_t>(wi.tn * p.n_tile) * Kc + m; col_stride = Kc; }
            }
            const long long te_a = p.trace ? clock64() : 0;
            mbar_wait(&bar_tfull[acc], acc_phase);
            const long long te_b = p.trace ? clock64() : 0;
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + acc * CT_MAX_N + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < p.n_tile; c += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c, v);
                float bv[16];
                const int nbase = wi.tn * p.n_tile + c;
#pragma unroll
                for (int j = 0; j < 16; ++j) {  // all bias loads first: independent of the stores below
                    int bidx = nbase + j;
                    if (MODE == CONV_DGRAD && p.quad) bidx -= static_cast<int>(p.fd_Cin.div(static_cast<uint32_t>(bidx))) * p.Cin;
                    bv[j] = (MODE != CONV_WGRAD && MODE != DENSE_WGRAD && p.bias != nullptr && nbase + j < p.gemm_n &&
                             (!p.atomic_out || wi.kb0 == 0)) ? __ldg(p.bias + bidx) : 0.0f;
                }
                tmem_ld_wait();
                if (!row_ok) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = nbase + j;
                    if (n < p.gemm_n) {
                        float val = __uint_as_float(v[j]) + bv[j];
                        float* o = dst + static_cast<size_t>(c + j) * col_stride;
                        if (MODE == CONV_DGRAD && p.quad) {
                            uint32_t cls, ci;
                            p.fd_Cin.divmod(static_cast<uint32_t>(n), cls, ci);
                            if (((cls & 2u) && !q_h1) || ((cls & 1u) && !q_w1)) continue;       // odd H / W: last row / column has no partner
                            o = dst + static_cast<size_t>(ci) * col_stride + (cls >> 1) * p.W + (cls & 1u);
                        }
                        if (MODE == CONV_WGRAD) {
                            atomicAdd(o, val);
                        } else if (MODE == DENSE_WGRAD) {
                            *o = val;
                        } else if (p.atomic_out) {
                            atomicAdd(o, val);
                        } else {
                            if (MODE == CONV_FWD && p.residual != nullptr) val += __ldg(p.residual + (o - p.out));
                            *o = act_lrelu(val, p.slope);
                        }
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);
            if (p.trace && warp == EPI_WARP0 && lane == 0) {
                trace_evt(p, 2, etrace_n, 0, te_a); trace_evt(p, 2, etrace_n, 1, te_b); trace_evt(p, 2, etrace_n, 2, clock64()); ++etrace_n;
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 2 * CT_MAX_N);
}

extern long long* g_conv_trace_ptr;
static int pick_n_tile(int n) {
    int t = (n + 15) / 16 * 16;
    if (t > CT_MAX_N) {
        // split evenly into tiles of at most 256 columns, multiples of 16
        const int parts = ceil_div(n, CT_MAX_N);
        t = ceil_div(ceil_div(n, parts), 16) * 16;
    }
    return t;
}

// Long-K problems with few output tiles (deep conv layers, the two big FC layers): split K across CTAs when the epilogue is
// linear (no activation, no residual); partial sums are accumulated with fp32 atomics into a zero-filled output.
static int maybe_split_k(const pgv_handle* h, ConvTcParams& p, size_t out_elems, cudaStream_t stream) {
    int m_tiles = 0;
    for (int c = 0; c < p.classes; ++c) m_tiles += p.m_tiles_class[c];
    const long long tiles = static_cast<long long>(m_tiles) * p.n_tiles;
    if (p.slope >= 0.0f || p.residual != nullptr || tiles >= h->sm_count || p.kb_total < 16) return 0;
    int splits = static_cast<int>((h->sm_count * 2) / tiles);
    if (splits > p.kb_total / 8) splits = p.kb_total / 8;
    if (splits <= 1) return 0;
    p.kb_per_split = ceil_div(p.kb_total, splits);
    p.k_splits = ceil_div(p.kb_total, p.kb_per_split);
    p.atomic_out = 1;
    PGV_CUDA(cudaMemsetAsync(p.out, 0, sizeof(float) * out_elems, stream));
    return 0;
}

template <int MODE>
static int launch_conv_tc(const pgv_handle* h, ConvTcParams& p, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        PGV_CUDA(cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
        configured = true;
    }
    int m_tiles = 0;
    for (int c = 0; c < p.classes; ++c) m_tiles += p.m_tiles_class[c];
    const long long items = static_cast<long long>(m_tiles) * p.n_tiles * p.k_splits;
    const int grid = static_cast<int>(items < h->sm_count ? items : h->sm_count);
    p.trace = g_conv_trace_ptr;
    conv_tc_kernel<MODE><<<grid, CT_THREADS, CT_SMEM, stream>>>(p);
    PGV_LAUNCH_CHECK();
    return 0;
}

static int fill_common(ConvTcParams& p, const char* who, int B, int Cin, int H, int W, int Cout, int kh, int kw, int stride, int pad,
                       int Ho, int Wo) {
    if (B <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0)
        return set_error(-1, "%s: bad geometry", who);
    if ((Ho - 1) * stride - 2 * pad + kh > H || (Wo - 1) * stride - 2 * pad + kw > W)
        return set_error(-1, "%s: output %dx%d does not fit input %dx%d", who, Ho, Wo, H, W);
    memset(&p, 0, sizeof(p));
    p.B = B; p.Cin = Cin; p.H = H; p.W = W; p.Cout = Cout; p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.Ho = Ho; p.Wo = Wo;
    p.classes = 1; p.k_splits = 1; p.sshift = (stride == 2) ? 1 : 0;
    p.fd_HWo.init(Ho * Wo); p.fd_Wo.init(Wo); p.fd_taps.init(kh * kw); p.fd_kw.init(kw);
    p.fd_dtaps.init(1); p.fd_dtapsw.init(1); p.fd_pixblocks.init(1); p.fd_ntile.init(1);
    for (int c = 0; c < 4; ++c) { p.fd_HcWc[c].init(1); p.fd_Wc[c].init(1); }
    return 0;
}

long long* g_conv_trace_ptr = nullptr;

}  // namespace pgv

using namespace pgv;

extern "C" {

int pgv_conv2d_fwd_tf32(pgv_handle* h, const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H, int W,
                        int Cout, int kh, int kw, int stride, int pad, int Ho, int Wo, float lrelu_slope, pgv_stream_t stream) {
    PGV_CHECK_ARG(h && x && w && y, "pgv_conv2d_fwd_tf32: NULL argument");
    ConvTcParams p;
    if (int rc = fill_common(p, "pgv_conv2d_fwd_tf32", B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo)) return rc;
    p.x = x; p.w = w; p.bias = bias; p.out = y; p.slope = lrelu_slope;
    p.gemm_n = Cout; p.gemm_k = Cin * kh * kw;
    p.n_tile = pick_n_tile(Cout); p.n_tiles = ceil_div(Cout, p.n_tile);
    p.kb_total = ceil_div(p.gemm_k, CT_BLOCK_K); p.kb_per_split = p.kb_total;
    p.m_tiles_class[0] = static_cast<int>((static_cast<long long>(B) * Ho * Wo + CT_BLOCK_M - 1) / CT_BLOCK_M);
    p.fast = (kh == 4 && kw == 4) ? 1 : ((kh == 1 && kw == 1) ? 2 : 0);
    if (stride != 1 && stride != 2) p.fast = 0;      // the slot fast paths use shift arithmetic for the stride
    p.fd_ntile.init(p.n_tile);
    if (p.fast == 1) { p.slot_a = 1; p.a_off[0] = 1; p.a_off[1] = 2; p.a_off[2] = 3; p.a_step = 2LL * H * W; p.a_kdim = p.gemm_k; p.a_pair = 0; /* neighbour-shuffle reuse measured slower on B200 (layers6 vs layers5) */ }
    if (p.fast == 2) { p.slot_a = 1; p.a_off[0] = H * W; p.a_off[1] = 2 * H * W; p.a_off[2] = 3 * H * W; p.a_step = 32LL * H * W; p.a_kdim = p.gemm_k; }
    p.slot_b = 1; p.b_off[0] = 1; p.b_off[1] = 2; p.b_off[2] = 3; p.b_step = 32; p.b_kdim = p.gemm_k; p.b_vec = (p.gemm_k % 4 == 0);
    if (int rc = maybe_split_k(h, p, static_cast<size_t>(B) * Cout * Ho * Wo, static_cast<cudaStream_t>(stream))) return rc;
    return launch_conv_tc<CONV_FWD>(h, p, static_cast<cudaStream_t>(stream));
}

int pgv_conv2d_dgrad_tf32(pgv_handle* h, const float* dy, const float* w, const float* bias, float* dx, int B, int Cin, int H, int W,
                          int Cout, int kh, int kw, int stride, int pad, int Ho, int Wo, float lrelu_slope, pgv_stream_t stream) {
    PGV_CHECK_ARG(h && dy && w && dx, "pgv_conv2d_dgrad_tf32: NULL argument");
    PGV_CHECK_ARG(stride == 1 || stride == 2, "pgv_conv2d_dgrad_tf32: stride %d unsupported", stride);
    ConvTcParams p;
    if (int rc = fill_common(p, "pgv_conv2d_dgrad_tf32", B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo)) return rc;
    p.x = dy; p.w = w; p.bias = bias; p.out = dx; p.slope = lrelu_slope;
    p.taps_h = ceil_div(kh, stride); p.taps_w = ceil_div(kw, stride);
    p.quad = (stride == 2 && kh == 4 && kw == 4 && pad % 2 == 0) ? 1 : 0;
    p.fd_Cin.init(Cin);
    p.gemm_n = p.quad ? 4 * Cin : Cin; p.gemm_k = Cout * p.taps_h * p.taps_w;
    p.n_tile = pick_n_tile(p.gemm_n); p.n_tiles = ceil_div(p.gemm_n, p.n_tile);
    p.kb_total = ceil_div(p.gemm_k, CT_BLOCK_K); p.kb_per_split = p.kb_total;
    p.classes = p.quad ? 1 : stride * stride;
    for (int c = 0; c < p.classes; ++c) {
        const int ph = c / stride, pw = c % stride;     // quad mode: the single "class" enumerates the 2x2 quads (ph = pw = 0)
        p.Hc[c] = (H - ph + stride - 1) / stride;
        p.Wc[c] = (W - pw + stride - 1) / stride;
        p.m_tiles_class[c] = static_cast<int>((static_cast<long long>(B) * p.Hc[c] * p.Wc[c] + CT_BLOCK_M - 1) / CT_BLOCK_M);
        p.fd_HcWc[c].init(p.Hc[c] * p.Wc[c]);
        p.fd_Wc[c].init(p.Wc[c]);
    }
    p.fast = (p.taps_h == 2 && p.taps_w == 2) ? 1 : ((p.taps_h == 1 && p.taps_w == 1) ? 2 : 0);
    p.fd_dtaps.init(p.taps_h * p.taps_w); p.fd_dtapsw.init(p.taps_w); p.fd_ntile.init(p.n_tile);
    p.a_dense = (H == 1 && W == 1 && Ho == 1 && Wo == 1 && kh == 1 && kw == 1) ? 1 : 0;   // Linear dgrad: A = dy [M, N] row-major
    const int HWo = Ho * Wo, khw = kh * kw;
    if (p.a_dense) { p.slot_a = 1; p.a_off[0] = 1; p.a_off[1] = 2; p.a_off[2] = 3; p.a_step = 32; p.a_kdim = p.gemm_k; p.a_vec = (p.gemm_k % 4 == 0); }
    else if (p.fast == 1) { p.slot_a = 1; p.a_off[0] = -1; p.a_off[1] = -Wo; p.a_off[2] = -Wo - 1; p.a_step = 8LL * HWo; p.a_kdim = p.gemm_k; p.a_pair = 0; }
    else if (p.fast == 2) { p.slot_a = 1; p.a_off[0] = HWo; p.a_off[1] = 2 * HWo; p.a_off[2] = 3 * HWo; p.a_step = 32LL * HWo; p.a_kdim = p.gemm_k; }
    if (p.fast == 1) { p.slot_b = 1; p.b_off[0] = stride; p.b_off[1] = stride * kw; p.b_off[2] = stride * kw + stride; p.b_step = 8LL * Cin * khw; p.b_kdim = p.gemm_k; }
    else if (p.fast == 2) { p.slot_b = 1; p.b_off[0] = Cin; p.b_off[1] = 2 * Cin; p.b_off[2] = 3 * Cin; p.b_step = 32LL * Cin; p.b_kdim = p.gemm_k; }
    if (int rc = maybe_split_k(h, p, static_cast<size_t>(B) * Cin * H * W, static_cast<cudaStream_t>(stream))) return rc;
    return launch_conv_tc<CONV_DGRAD>(h, p, static_cast<cudaStream_t>(stream));
}

/* Debug hook (tools/gpu_trace_conv.py): the next pgv_conv2d_*_tf32 launches record CTA 0's pipeline timestamps into
 * `trace_dev` (3 roles x 64 events x 8 int64).  Pass NULL to disable. */
int pgv_debug_set_conv_trace(void* trace_dev) { pgv::g_conv_trace_ptr = static_cast<long long*>(trace_dev); return 0; }

int pgv_conv2d_wgrad_tf32(pgv_handle* h, const float* x, const float* dy, float* dw, int B, int Cin, int H, int W, int Cout, int kh,
                          int kw, int stride, int pad, int Ho, int Wo, pgv_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PGV_CHECK_ARG(h && x && dy && dw, "pgv_conv2d_wgrad_tf32: NULL argument");
    ConvTcParams p;
    if (int rc = fill_common(p, "pgv_conv2d_wgrad_tf32", B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo)) return rc;
    p.x = x; p.w = dy; p.out = dw; p.slope = -1.0f;
    const int Kc = Cin * kh * kw;
    p.gemm_n = Cout; p.gemm_k = 0;
    p.n_tile = pick_n_tile(Cout); p.n_tiles = ceil_div(Cout, p.n_tile);
    p.pix_blocks = ceil_div(Ho * Wo, CT_BLOCK_K);
    p.kb_total = B * p.pix_blocks;
    p.m_tiles_class[0] = ceil_div(Kc, CT_BLOCK_M);
    const int tiles = p.m_tiles_class[0] * p.n_tiles;
    int splits = ceil_div(h->sm_count * 2, tiles);          // ~2 work items per SM
    if (splits > p.kb_total / 4) splits = p.kb_total / 4;    // at least 4 k-blocks per item
    if (splits < 1) splits = 1;
    p.kb_per_split = ceil_div(p.kb_total, splits);
    p.k_splits = ceil_div(p.kb_total, p.kb_per_split);
    p.fd_pixblocks.init(p.pix_blocks); p.fd_ntile.init(p.n_tile);
    PGV_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * static_cast<size_t>(Cout) * Kc, stream));
    return launch_conv_tc<CONV_WGRAD>(h, p, stream);
}

/* Linear layers on the same kernel (no alignment requirement on K): x [M, K], w [N, K] (nn.Linear), y [M, N]. */
int pgv_linear_fwd_tf32(pgv_handle* h, const float* x, const float* w, const float* bias, const float* residual, float* y, int M, int N,
                        int K, int relu, pgv_stream_t stream) {
    PGV_CHECK_ARG(h && x && w && y && M > 0 && N > 0 && K > 0, "pgv_linear_fwd_tf32: bad argument");
    ConvTcParams p;
    if (int rc = fill_common(p, "pgv_linear_fwd_tf32", M, K, 1, 1, N, 1, 1, 1, 0, 1, 1)) return rc;
    p.x = x; p.w = w; p.bias = bias; p.residual = residual; p.out = y; p.slope = relu ? 0.0f : -1.0f;
    p.gemm_n = N; p.gemm_k = K;
    p.n_tile = pick_n_tile(N); p.n_tiles = ceil_div(N, p.n_tile);
    p.kb_total = ceil_div(K, CT_BLOCK_K); p.kb_per_split = p.kb_total;
    p.m_tiles_class[0] = ceil_div(M, CT_BLOCK_M);
    p.fast = 2; p.a_dense = 1;
    p.fd_ntile.init(p.n_tile);
    p.slot_a = 1; p.a_off[0] = 1; p.a_off[1] = 2; p.a_off[2] = 3; p.a_step = 32; p.a_kdim = K; p.a_vec = (K % 4 == 0);
    p.slot_b = 1; p.b_off[0] = 1; p.b_off[1] = 2; p.b_off[2] = 3; p.b_step = 32; p.b_kdim = K; p.b_vec = (K % 4 == 0);
    if (int rc = maybe_split_k(h, p, static_cast<size_t>(M) * N, static_cast<cudaStream_t>(stream))) return rc;
    return launch_conv_tc<CONV_FWD>(h, p, static_cast<cudaStream_t>(stream));
}

int pgv_linear_dgrad_tf32(pgv_handle* h, const float* dy, const float* w, float* dx, int M, int N, int K, pgv_stream_t stream) {
    return pgv_conv2d_dgrad_tf32(h, dy, w, nullptr, dx, M, K, 1, 1, N, 1, 1, 1, 0, 1, 1, -1.0f, stream);
}

int pgv_linear_wgrad_tf32(pgv_handle* h, const float* dy, const float* x, float* dw, int M, int N, int K, pgv_stream_t stream) {
    PGV_CHECK_ARG(h && dy && x && dw && M > 0 && N > 0 && K > 0, "pgv_linear_wgrad_tf32: bad argument");
    ConvTcParams p;
    if (int rc = fill_common(p, "pgv_linear_wgrad_tf32", M, K, 1, 1, N, 1, 1, 1, 0, 1, 1)) return rc;
    p.x = x; p.w = dy; p.out = dw; p.slope = -1.0f;
    p.gemm_n = N; p.gemm_k = M;
    p.n_tile = pick_n_tile(N); p.n_tiles = ceil_div(N, p.n_tile);
    p.kb_total = ceil_div(M, CT_BLOCK_K); p.kb_per_split = p.kb_total;
    p.m_tiles_class[0] = ceil_div(K, CT_BLOCK_M);
    p.fd_ntile.init(p.n_tile);
    p.slot_a = 1; p.a_off[0] = K; p.a_off[1] = 2 * K; p.a_off[2] = 3 * K; p.a_step = 32LL * K; p.a_kdim = M;
    p.slot_b = 1; p.b_off[0] = N; p.b_off[1] = 2 * N; p.b_off[2] = 3 * N; p.b_step = 32LL * N; p.b_kdim = M;
    return launch_conv_tc<DENSE_WGRAD>(h, p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
