// Channels-last (NHWC) implicit-GEMM convolutions on tcgen05 (TF32 products, fp32 accumulation in TMEM) fed by cp.async.
//
// With activations stored [B, H, W, C] (C a multiple of 4) and the reduction index ordered (kh, kw, c), every 16-byte
// chunk of every operand tile is 16 contiguous, 16-byte-aligned bytes of global memory:
//
//   GEMM mode  (convolution forward, and data-gradient / ConvTranspose2d forward through a re-packed weight matrix)
//       out[m, n] = act(bias[n] + sum_k A[m, k] * Bw[n, k])
//       m = (b, oh, ow) output pixel (or 2x2 pixel quad), k = (kh, kw, c): A[m, k] = in[b, oh*st-p+kh, ow*st-p+kw, c]
//       (for one kh the KW*C values are ONE contiguous run of memory); Bw = prepared weights [N][K], K contiguous.
//       Both tiles are K-major, 128-byte rows, SWIZZLE_128B.
//   WGRAD mode dWcl[n, m] += sum_k A[m, k] * Bd[n, k],  m = (kh, kw, ci), n = co, k = pixel:
//       A[m, k] = x[pixel k shifted by tap (kh, kw), ci] and Bd[n, k] = dy[pixel k, co] are contiguous along m / n for one
//       pixel, so both tiles are MN-major (atoms of 4 pixels x 128 bytes, SWIZZLE_128B_BASE32B).  The reduction runs over
//       (oh, ow, b) with b fastest, so that the 32 pixels of a k-block share their tap geometry.
//
// Nothing goes through registers: 8 producer warps issue cp.async (LDGSTS, zero-fill for padding / tails) straight into
// the swizzled tile and hand completion to the stage's mbarrier (cp.async.mbarrier.arrive.noinc), so up to CL_STAGES
// k-blocks (160 KB) are in flight per SM; in GEMM mode the weight tile of a k-block is one TMA box issued by a dedicated
// warp.  The operands are therefore NOT rounded on the way in: activations are rounded to TF32 (round-to-nearest) by the
// kernels that write them (BatchNorm apply, thin-layer kernels, layout converters: `round_out` flags) and weights by
// pgv_conv_cl_prep_weights, which is numerically identical to rounding in the gather and keeps the tensor core's own
// truncation out of the picture.
// One elected lane of a converged warp issues tcgen05.mma (M = 128, N = tile width, K = 8), two k-blocks per barrier round
// trip when both are ready; accumulators are double-buffered in TMEM; 4 epilogue warps drain them through a per-warp
// shared-memory tile (bias, LeakyReLU, optional TF32 rounding, float4 stores along the channel dimension) and, for
// launches with a single N tile, also accumulate the per-channel sum / sum of squares of what they store (the batch
// statistics of the BatchNorm2d that follows).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "pgv_common.cuh"
#include "pgv_tc.cuh"

namespace pgv {

enum { CL_GEMM = 0, CL_WGRAD = 1 };
enum { CL_EPI_ROWS = 0, CL_EPI_QUAD = 1 };

// The ring of k-block stages takes all the shared memory the epilogue does not need: a stage is the 16 KB activation tile plus a weight
// tile of n_tile x 128 bytes, so the 8 / 16 / 32-channel layers (whose per-SM throughput is bounded by the BYTES IN FLIGHT: four
// k-blocks per tile, ~2 us of memory latency) run 10-11 stages deep and the 128-column layers 6.
constexpr int CL_BLOCK_M = 128, CL_BLOCK_K = 32, CL_MAX_N = 128, CL_MAX_STAGES = 12;
constexpr int CL_A_BYTES = CL_BLOCK_M * 128;
constexpr int CL_PRODUCER_WARPS = 8, CL_PRODUCERS = CL_PRODUCER_WARPS * 32;
constexpr int CL_SLOTS = (CL_BLOCK_M * 8) / CL_PRODUCERS;          // 16-byte chunks of one operand tile per producer thread (4)
constexpr int CL_ROWS_PER_PASS = CL_PRODUCERS / 8;               // GEMM mode: rows covered by one pass of the producers (32)
// 16 warps (the most that fit at 128 registers): 0-7 cp.async producers, 8 MMA issuer, 9-12 epilogue group 0, 13 TMA, 14-15 TMA helpers.
// When the A operand comes through TMA (GEMM mode, a_mode 1 / 2) the producer warps have nothing to gather and become epilogue groups
// 1 (warps 0-3) and 2 (warps 4-7): three accumulator tiles are drained concurrently, which is what the 8 / 16 / 32-channel layers need -
// their tile period is the epilogue's instruction stream, one latency-bound warp per TMEM lane quarter (trace: 2 400 - 3 900 cycles
// per tile against ~900 of MMA issue).
constexpr int CL_THREADS = CL_PRODUCERS + 32 /*mma*/ + 128 /*epilogue*/ + 32 /*TMA*/ + 64 /*TMA helpers*/;
constexpr int CL_MAX_GROUPS = 3;
// epilogue staging: per epilogue warp 32 rows x 32 columns (+4 pad) of fp32 and one 64-bit destination offset per row
constexpr int CL_EPI_LD = 36, CL_EPI_WARP_BYTES = 32 * CL_EPI_LD * 4 + 32 * 8;
// Shared-memory budget: 188 KB, NOT the 227 KB maximum - the kernel shares its SMs with the CTAs of concurrent kernels (NCCL's
// all-reduce on the communication stream, the 27 KB conditioner kernels of the other branch), which need ~40 KB to be resident at all;
// with nothing left, each of them would serialise with a persistent 148-CTA convolution.  (A deeper ring bought nothing: measured.)
constexpr int CL_SMEM = 188 * 1024;
__host__ __device__ constexpr int cl_ring_bytes(int groups) { return (CL_SMEM - 1024 - 256 - 4 * groups * CL_EPI_WARP_BYTES) / 1024 * 1024; }

struct ConvClParams {
    const float* a;      // GEMM: gathered activations [B, H, W, C]      WGRAD: x [B, H, W, C]
    const float* b;      // GEMM: prepared weights [gemm_n][gemm_k]      WGRAD: dy [B, Hg, Wg, gemm_n]
    const float* bias;   // GEMM: per output channel (or NULL)
    float* out;
    int B, H, W, C, KH, KW, stride, pad, Hg, Wg;     // gather geometry: window KH x KW over [H, W, C], output grid Hg x Wg
    int gemm_m, gemm_n, gemm_k;
    int n_tile, n_tiles, m_tiles, kb_total, kb_per_split, k_splits;
    int stages, stage_bytes;       // ring geometry (cl_set_ring)
    int n_groups;                  // epilogue groups (1 or CL_MAX_GROUPS): three only pay when a CTA drains many small tiles
    int epi, ldo;        // CL_EPI_ROWS: out[m * ldo + n]          WGRAD: out[n * ldo + m] for m < m_valid
    int m_valid;
    int qH, qW, qC;      // CL_EPI_QUAD: out is [B, qH, qW, qC]; row m = quad (b, i, j), column n = (2*ph + pw) * qC + c
    int bblocks;         // WGRAD: ceil(B / 32) k-blocks per output position
    int round_out, atomic_out;
    double* stats;       // GEMM: if not NULL, stats[2c] += sum, stats[2c + 1] += sum of squares of the stored values of channel c
    const float* stat_x; // GEMM + stats: if not NULL (same layout as out), stats[2c + 1] += sum of stored value * stat_x instead of squares:
                         // the raw sums of the BatchNorm BACKWARD in front of this launch's output (pgv_bn_cl_train_bwd, raw_sums)
    int stat_c;          // number of channels (gemm_n, or qC with the quad epilogue)
    long long* trace;    // debug (tools/gpu_trace_conv.py): clock64 timestamps of CTA 0, [role][64 events][8]; NULL in production
    float slope;
    // A operand of GEMM mode: 0 = cp.async gather by the producer warps, 1 = one tiled TMA box per k-block (1x1 convolutions / Linear:
    // A is a plain [gemm_m][C] matrix), 2 = TMA im2col loads (one per tap and k-block: `a_sub` loads of 128 pixels x min(C, 32) channels)
    int a_mode, a_sub, a_sbo;      // a_sbo: byte distance of 8-row groups inside one load's tile (1024 / 512 / 256 for 128 / 64 / 32-byte rows)
    unsigned a_layout;             // UMMA layout type of the A tile: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
    // deterministic split-K: partial accumulator tiles go to `ws` ([tile][split][column][128 rows], fp32) instead of fp32 atomics.
    // GEMM mode: the CTA that finishes a tile last (ws_cnt[tile], self-resetting) sums the partials in split order and runs the
    // normal epilogue; WGRAD mode: a finish kernel sums them and writes the weight gradient in its final layout.
    float* ws;
    unsigned* ws_cnt;
    int wg_cin, wg_taps;           // WGRAD: if wg_taps > 0, out is the PyTorch weight layout [Cout][Cin][taps] (row m = tap * Cin + ci)
    FastDiv fd_HgWg, fd_Wg, fd_span /* KW*C */, fd_C, fd_bblocks, fd_qC, fd_splits, fd_ntiles;
    alignas(64) CUtensorMap tmap_b;      // GEMM mode: prepared weights [gemm_n][gemm_k], box 32 x n_tile, 128-byte swizzle
    alignas(64) CUtensorMap tmap_a;      // GEMM mode, a_mode 1: [gemm_m][C] tiled, box 32 x 128; a_mode 2: im2col map of the [B, H, W, C] activation
};

__device__ __forceinline__ uint32_t cl_sw128(int row, int chunk) {
    return static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// MN-major fp32 / TF32 operands must use the "128-byte swizzle with 32-byte atomicity" layout (the only MN-major layout the
// tensor core accepts for 32-bit types): atoms of 4 k-rows x 128 bytes in which the 32-byte unit index is XORed with the row
// index (byte-address bits [5,7) ^= bits [7,9)).  Offset of 16-byte chunk `c16` (0..7) inside row `row` of a 128-byte-row block:
__device__ __forceinline__ uint32_t cl_sw32(int c16, int row) {
    return static_cast<uint32_t>(((((c16 >> 1) ^ (row & 3)) << 1) | (c16 & 1)) << 4);
}

__device__ __forceinline__ void cl_trace(const ConvClParams& p, int role, int& n, long long a, long long b, long long c, int tag,
                                         long long d = 0, long long e2 = 0, long long f = 0) {
    if (p.trace != nullptr && blockIdx.x == 0 && n < 64) {
        long long* e = p.trace + (role * 64 + n) * 8;
        e[0] = a; e[1] = b; e[2] = c; e[3] = d; e[4] = e2; e[5] = f; e[6] = tag;
        ++n;
    }
}

struct ClItem { int tm, tn, kb0, kb1; };
__device__ __forceinline__ ClItem cl_decode(const ConvClParams& p, int item) {
    ClItem wi;
    uint32_t t, split, tm, tn;
    p.fd_splits.divmod(static_cast<uint32_t>(item), t, split);
    p.fd_ntiles.divmod(t, tm, tn);
    wi.tn = static_cast<int>(tn);
    wi.tm = static_cast<int>(tm);
    wi.kb0 = static_cast<int>(split) * p.kb_per_split;
    wi.kb1 = min(p.kb_total, wi.kb0 + p.kb_per_split);
    return wi;
}

__device__ __forceinline__ float cl_act(float v, float slope, int round_out) {
    if (slope >= 0.0f && v < 0.0f) v *= slope;
    return round_out ? to_tf32_rna(v) : v;
}

template <int MODE>
__global__ void __launch_bounds__(CL_THREADS, 1) conv_cl_kernel(const __grid_constant__ ConvClParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int CL_STAGES = p.stages, CL_STAGE_BYTES = p.stage_bytes;
    uint8_t* smem_ctl = smem + cl_ring_bytes(p.n_groups);                    // barriers, then the epilogue staging tiles
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_ctl);
    uint64_t* bar_empty = bar_full + CL_MAX_STAGES;
    uint64_t* bar_tfull = bar_empty + CL_MAX_STAGES;
    uint64_t* bar_tempty = bar_tfull + CL_MAX_GROUPS;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_tempty + CL_MAX_GROUPS);
    volatile uint32_t* ws_flags = tmem_ptr + 1;                             // split-K: "this CTA finishes the tile", one per epilogue group
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int MMA_WARP = CL_PRODUCER_WARPS, TMA_WARP = CL_PRODUCER_WARPS + 5;      // warps MMA_WARP + 1 .. + 4 are epilogue group 0
    const bool a_by_tma = MODE == CL_GEMM && p.a_mode != 0;
    // accumulator tiles in TMEM and epilogue groups: item j of this CTA uses accumulator j % n_acc and is drained by group j % n_groups
    const int n_groups = p.n_groups, n_acc = n_groups > 1 ? n_groups : 2;
    const bool with_stats = MODE == CL_GEMM && p.stats != nullptr;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            // arrivals per stage: the producer threads' cp.async completions (+ the TMA warp's expect_tx arrive in GEMM mode); with a
            // TMA-fed A operand the TMA warp is the only producer
            const uint32_t full_count = MODE == CL_WGRAD ? CL_PRODUCERS : (p.a_mode != 0 ? 1 : CL_PRODUCERS + 1);
            for (int s = 0; s < CL_STAGES; ++s) { mbar_init(&bar_full[s], full_count); mbar_init(&bar_empty[s], 1); }
            for (int a = 0; a < CL_MAX_GROUPS; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 4); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 4 * CL_MAX_N);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_items = p.m_tiles * p.n_tiles * p.k_splits;
    const uint32_t smem_base = smem_u32(smem);

    const int epi_group = (warp > MMA_WARP && warp < TMA_WARP) ? 0 : ((n_groups > 1 && warp < CL_PRODUCER_WARPS) ? 1 + (warp >> 2) : -1);
    if (warp < CL_PRODUCER_WARPS && !a_by_tma) {
        // ================================================================== producers
        const int t = threadIdx.x;
        int stage = 0; uint32_t phase = 0;
        int trace_n = 0;
        if (MODE == CL_GEMM) {
            // A operand: thread = (row group r0 = t / 8, chunk j = t % 8): 8 consecutive lanes copy one 128-byte row; rows
            // r0 + CL_ROWS_PER_PASS * i.  B operand (prepared weights, a plain K-major matrix): ONE TMA box per k-block, issued by
            // thread 0, which also posts the expected byte count on the stage's barrier.
            const int j = t & 7, r0 = t >> 3;
            const uint32_t dst0 = cl_sw128(r0, j);                    // rows r0 + 32 i: + i * 4096 (same row & 7)
            constexpr uint32_t PASS_BYTES = CL_ROWS_PER_PASS * 128;
            const uint32_t span = static_cast<uint32_t>(p.KW) * p.C;  // a multiple of 32: a k-block never straddles two kernel rows
            const long long row_pitch = static_cast<long long>(p.W) * p.C;
            // the ~170-cycle latency of testing the stage's "empty" barrier is taken off the critical path: the test for the NEXT
            // stage is issued right after the copies of the current one
            bool ready = mbar_test_wait(&bar_empty[0], 1);
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const ClItem wi = cl_decode(p, item);
                const float* a_ptr[CL_SLOTS]; uint32_t a_msk[CL_SLOTS];          // mask: bits 0-7 valid kh, bits 8-15 valid kw
#pragma unroll
                for (int i = 0; i < CL_SLOTS; ++i) {
                    const uint32_t m = static_cast<uint32_t>(wi.tm) * CL_BLOCK_M + r0 + CL_ROWS_PER_PASS * i;
                    uint32_t b, rem, oh, ow;
                    p.fd_HgWg.divmod(m, b, rem);
                    p.fd_Wg.divmod(rem, oh, ow);
                    const int ih0 = static_cast<int>(oh) * p.stride - p.pad, iw0 = static_cast<int>(ow) * p.stride - p.pad;
                    a_ptr[i] = p.a + ((static_cast<long long>(b) * p.H + ih0) * p.W + iw0) * p.C;     // only dereferenced where the masks allow
                    const int h_lo = max(0, -ih0), h_hi = min(p.KH, p.H - ih0), w_lo = max(0, -iw0), w_hi = min(p.KW, p.W - iw0);
                    const uint32_t hm = h_hi > h_lo ? ((1u << h_hi) - (1u << h_lo)) : 0u, wm = w_hi > w_lo ? ((1u << w_hi) - (1u << w_lo)) : 0u;
                    a_msk[i] = (m < static_cast<uint32_t>(p.gemm_m)) ? (hm | (wm << 8)) : 0u;
                }
                uint32_t kh, rem;                                   // this thread's chunk: k = 32 kb + 4 j = (kh, rem = kw * C + c)
                p.fd_span.divmod(static_cast<uint32_t>(wi.kb0) * CL_BLOCK_K + 4 * j, kh, rem);
                long long koff = static_cast<long long>(kh) * row_pitch + rem;
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    const uint32_t sel = (1u << kh) | (256u << p.fd_C.div(rem));
                    const long long tr_a = (p.trace && t == 0) ? clock64() : 0;
                    if (!ready) mbar_wait(&bar_empty[stage], phase ^ 1);
                    const long long tr_b = (p.trace && t == 0) ? clock64() : 0;
                    const int nstage = (stage + 1 == CL_STAGES) ? 0 : stage + 1;          // test the next stage before this stage's copies
                    const bool next_ready = mbar_test_wait(&bar_empty[nstage], (nstage == 0 ? (phase ^ 1) : phase) ^ 1);
                    const uint32_t sA = smem_base + stage * CL_STAGE_BYTES + dst0;
#pragma unroll
                    for (int i = 0; i < CL_SLOTS; ++i) {
                        const bool ok = (a_msk[i] & sel) == sel;
                        cp_async16_ca(sA + i * PASS_BYTES, ok ? a_ptr[i] + koff : p.a, ok ? 16u : 0u);
                    }
                    cp_async_mbar_arrive_noinc(&bar_full[stage]);
                    if (p.trace && t == 0) cl_trace(p, 0, trace_n, tr_a, tr_b, clock64(), kb);
                    if (++stage == CL_STAGES) { stage = 0; phase ^= 1; }
                    ready = next_ready;
                    rem += CL_BLOCK_K; koff += CL_BLOCK_K;
                    if (rem >= span) { rem -= span; ++kh; koff += row_pitch - span; }
                }
            }
        } else {
            // WGRAD.  A tile: pixel kk = (t / 32) + 8 i (i = k-group), 16-byte chunk cc = t % 32 along the 128 m of the tile:
            // a warp copies 512 contiguous bytes of one pixel.  B tile: n_tile / 4 chunks per pixel, same idea.
            // slot i of this thread: pixel kk = (t / 32) + CL_PRODUCER_WARPS * i of the k-block
            const int cc = t & 31, kk0 = t >> 5;
            uint32_t a_dst[CL_SLOTS];
#pragma unroll
            for (int i = 0; i < CL_SLOTS; ++i) {
                const int kk = kk0 + CL_PRODUCER_WARPS * i;
                a_dst[i] = static_cast<uint32_t>((kk >> 3) * 4096 + (cc >> 3) * 1024 + (kk & 7) * 128) + cl_sw32(cc & 7, kk & 7);
            }
            const int cpp = p.n_tile >> 2;                            // B chunks per pixel: 8, 16 or 32
            const int cpp_shift = (cpp == 8) ? 3 : ((cpp == 16) ? 4 : 5);
            const uint32_t b_group_bytes = static_cast<uint32_t>(p.n_tile >> 5) * 1024;    // one 8-pixel group of the B tile
            const long long img_a = static_cast<long long>(p.H) * p.W * p.C, img_b = static_cast<long long>(p.Hg) * p.Wg * p.gemm_n;
            bool wready = mbar_test_wait(&bar_empty[0], 1);
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const ClItem wi = cl_decode(p, item);
                // this thread's 4 consecutive m (fixed for the tile): tap (kh, kw) and channel
                const uint32_t m = static_cast<uint32_t>(wi.tm) * CL_BLOCK_M + 4 * cc;
                uint32_t kh, rem;
                p.fd_span.divmod(m, kh, rem);
                const uint32_t kw = p.fd_C.div(rem);
                const bool m_ok = m < static_cast<uint32_t>(p.gemm_m);
                const long long m_off = static_cast<long long>(kh) * p.W * p.C + rem;
                // B slots: id = t + 256 s -> (pixel, chunk)
                int b_kk[CL_SLOTS]; uint32_t b_dst[CL_SLOTS]; bool b_on[CL_SLOTS], b_ok[CL_SLOTS];
                const float* b_base[CL_SLOTS];                   // dy address of (image b_kk, position 0, this thread's 4 columns)
#pragma unroll
                for (int s = 0; s < CL_SLOTS; ++s) {
                    const int id = t + CL_PRODUCERS * s;
                    b_on[s] = id < 32 * cpp;
                    b_kk[s] = id >> cpp_shift;
                    const int kk = b_kk[s], nc = id & (cpp - 1);
                    b_dst[s] = static_cast<uint32_t>((kk >> 3) * b_group_bytes + (nc >> 3) * 1024 + (kk & 7) * 128) + cl_sw32(nc & 7, kk & 7);
                    b_ok[s] = b_on[s] && (wi.tn * p.n_tile + 4 * nc) < p.gemm_n;
                    b_base[s] = p.b + kk * img_b + wi.tn * p.n_tile + 4 * nc;
                }
                const float* a_base[CL_SLOTS];                   // x address of (image kk0 + 8 i, row 0, column 0, this thread's tap / channels)
#pragma unroll
                for (int i = 0; i < CL_SLOTS; ++i) a_base[i] = p.a + (kk0 + CL_PRODUCER_WARPS * i) * img_a + m_off;
                // k-block state: output position (oh, ow) and image block bb; pointers advance by 32 images per k-block and are
                // recomputed when the position changes (every `bblocks` k-blocks)
                uint32_t pos, bb, oh, ow;
                p.fd_bblocks.divmod(static_cast<uint32_t>(wi.kb0), pos, bb);
                p.fd_Wg.divmod(pos, oh, ow);
                const float* a_ptr[CL_SLOTS]; const float* b_ptr[CL_SLOTS];
                bool tap_ok = false;
                int img0 = 0;
                auto set_position = [&]() {
                    const int ih = static_cast<int>(oh) * p.stride - p.pad + static_cast<int>(kh);
                    const int iw = static_cast<int>(ow) * p.stride - p.pad + static_cast<int>(kw);
                    tap_ok = m_ok && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
                    img0 = static_cast<int>(bb) * 32;
                    const long long a_off = ((static_cast<long long>(oh) * p.stride - p.pad) * p.W + (static_cast<long long>(ow) * p.stride - p.pad)) * p.C +
                                            img0 * img_a;
                    const long long b_off = static_cast<long long>(pos) * p.gemm_n + img0 * img_b;
#pragma unroll
                    for (int i = 0; i < CL_SLOTS; ++i) { a_ptr[i] = a_base[i] + a_off; b_ptr[i] = b_base[i] + b_off; }
                };
                set_position();
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    if (!wready) mbar_wait(&bar_empty[stage], phase ^ 1);
                    const int nstage = (stage + 1 == CL_STAGES) ? 0 : stage + 1;
                    const bool next_ready = mbar_test_wait(&bar_empty[nstage], (nstage == 0 ? (phase ^ 1) : phase) ^ 1);
                    const uint32_t sA = smem_base + stage * CL_STAGE_BYTES, sB = sA + CL_A_BYTES;
#pragma unroll
                    for (int i = 0; i < CL_SLOTS; ++i) {
                        const bool ok = tap_ok && img0 + kk0 + CL_PRODUCER_WARPS * i < p.B;
                        cp_async16_ca(sA + a_dst[i], ok ? a_ptr[i] : p.a, ok ? 16u : 0u);
                    }
#pragma unroll
                    for (int s = 0; s < CL_SLOTS; ++s) {
                        if (b_on[s]) {
                            const bool ok = b_ok[s] && img0 + b_kk[s] < p.B;
                            cp_async16_cg(sB + b_dst[s], ok ? b_ptr[s] : p.b, ok ? 16u : 0u);
                        }
                    }
                    cp_async_mbar_arrive_noinc(&bar_full[stage]);
                    if (++stage == CL_STAGES) { stage = 0; phase ^= 1; }
                    wready = next_ready;
                    if (++bb < static_cast<uint32_t>(p.bblocks)) {
                        img0 += 32;
#pragma unroll
                        for (int i = 0; i < CL_SLOTS; ++i) { a_ptr[i] += 32 * img_a; b_ptr[i] += 32 * img_b; }
                    } else {
                        bb = 0; ++pos;
                        if (++ow == static_cast<uint32_t>(p.Wg)) { ow = 0; ++oh; }
                        set_position();
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ================================================================== MMA issuer
        // The whole warp runs the loop (waits, address arithmetic) so that every operand of tcgen05.mma stays in uniform registers;
        // one elected lane issues.  (Under `if (lane == 0)` the compiler wraps EVERY UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY
        // loop to make its operands uniform.)
        {
            const uint32_t idesc = umma_idesc_tf32(CL_BLOCK_M, p.n_tile) | (MODE == CL_WGRAD ? (UMMA_IDESC_A_MN | UMMA_IDESC_B_MN) : 0u);
            const uint32_t b_group_bytes = static_cast<uint32_t>(p.n_tile >> 5) * 1024;
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            int trace_n = 0;
            bool full_ready = false;
            // A descriptors of the 8 / 16-channel im2col tiles: constant upper part and the byte offset of every K = 8 slice
            const uint64_t a_desc_hi = umma_smem_desc_kmajor(0, p.a_sbo, p.a_layout);
            uint32_t a_koff[CL_BLOCK_K / 8];
            {
                const uint32_t sub_bytes = CL_A_BYTES / (p.a_sub > 0 ? p.a_sub : 1), per_sub = (CL_BLOCK_K / 8) / (p.a_sub > 0 ? p.a_sub : 1);
#pragma unroll
                for (int k = 0; k < CL_BLOCK_K / 8; ++k) a_koff[k] = (static_cast<uint32_t>(k) / per_sub) * sub_bytes + (static_cast<uint32_t>(k) % per_sub) * 32;
            }
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const ClItem wi = cl_decode(p, item);
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * CL_MAX_N;
                // Two k-blocks per round trip where the tile has them: both stages are waited for (the second test overlaps the first
                // wait), then ONE elected issue of 8 MMAs + 2 commits.  The fixed cost of an iteration (barrier tests, fence, elect,
                // re-convergence, loop) is of the order of the MMA issue time itself, so halving it per k-block matters.
                auto issue_stage = [&](int stg, bool accumulate) {
                    const uint32_t a_addr = smem_base + stg * CL_STAGE_BYTES, b_addr = a_addr + CL_A_BYTES;
                    if (MODE == CL_GEMM) {
                        const uint64_t db = umma_smem_desc_sw128(b_addr);
                        if (p.a_sub == 1) {
                            const uint64_t da = umma_smem_desc_sw128(a_addr);
#pragma unroll
                            for (int k = 0; k < CL_BLOCK_K / 8; ++k) umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
                        } else {
                            // im2col loads of C = 8 / 16 channels: the k-block holds 4 / 2 single-tap tiles of 128 rows x 32 / 64 bytes
                            // (SWIZZLE_32B / 64B), one K = 8 MMA per 32 bytes of a row; the weight tile keeps its 128-byte rows
#pragma unroll
                            for (int k = 0; k < CL_BLOCK_K / 8; ++k) {
                                const uint64_t da = a_desc_hi | static_cast<uint64_t>(((a_addr + a_koff[k]) >> 4) & 0x3FFF);
                                umma_tf32(tmem_d, da, db + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
                            }
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < CL_BLOCK_K / 8; ++g) {
                            // one MMA = 8 pixels = two 4-row atoms 512 bytes apart (SBO); 32-element groups along M / N are 1024 bytes apart (LBO)
                            const uint64_t da = umma_smem_desc_mn_sw128_32b(a_addr + g * 4096, 1024, 512);
                            const uint64_t db = umma_smem_desc_mn_sw128_32b(b_addr + g * b_group_bytes, 1024, 512);
                            umma_tf32(tmem_d, da, db, idesc, (accumulate || g > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&bar_empty[stg]);
                };
                for (int kb = wi.kb0; kb < wi.kb1;) {
                    const bool pair = kb + 1 < wi.kb1;
                    const int s0 = stage, s1 = (s0 + 1 == CL_STAGES) ? 0 : s0 + 1;
                    const uint32_t ph0 = phase, ph1 = (s1 == 0) ? ph0 ^ 1 : ph0;
                    const long long tr_a = p.trace ? clock64() : 0;
                    if (!full_ready) mbar_wait(&bar_full[s0], ph0);
                    // The phases complete through the producers' cp.async.mbarrier.arrive, i.e. only once their copies have landed;
                    // like CUTLASS's sm100 cp.async mainloop (sm100_mma_cpasync_warpspecialized.hpp) no proxy fence is issued here.
                    bool r1 = mbar_test_wait(&bar_full[s1], ph1);
                    if (pair && !r1) { mbar_wait(&bar_full[s1], ph1); r1 = true; }
                    int s2 = s1; uint32_t ph2 = ph1; bool r2 = r1;          // the stage after this round trip and whether it is already full
                    if (pair) {
                        s2 = (s1 + 1 == CL_STAGES) ? 0 : s1 + 1;
                        ph2 = (s2 == 0) ? ph1 ^ 1 : ph1;
                        r2 = mbar_test_wait(&bar_full[s2], ph2);
                    }
                    const long long tr_b = p.trace ? clock64() : 0;
                    tc_fence_after_sync();
                    if (elect_one()) {
                        issue_stage(s0, kb > wi.kb0);
                        if (pair) issue_stage(s1, true);
                    }
                    const long long tr_e = p.trace ? clock64() : 0;
                    __syncwarp();
                    if (p.trace && lane == 0) cl_trace(p, 1, trace_n, tr_a, tr_b, clock64(), kb, tr_b, tr_e);
                    stage = s2; phase = ph2; full_ready = r2;
                    kb += pair ? 2 : 1;
                }
                if (elect_one()) umma_commit(&bar_tfull[acc]);
                __syncwarp();
                if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= TMA_WARP) {
        // ================================================================== TMA (GEMM mode): the weight box of every k-block and, with a
        // TMA-fed A operand, the activation loads.  Issuing one cp.async.bulk.tensor costs its thread ~90 cycles, so the (up to four)
        // im2col loads of a k-block are spread over the TMA warp and its two helper warps: role 0 = TMA warp (expect_tx, weights, A load 0),
        // role 1 = A loads 1 and 3, role 2 = A load 2.  Bytes that land before the expect_tx is posted only drive the transaction count
        // negative; the phase cannot complete before the TMA warp's arrival.
        const int tma_role = warp - TMA_WARP;
        if (MODE == CL_GEMM && (tma_role == 0 || (p.a_mode == 2 && tma_role < p.a_sub))) {
            // (a_mode 4 is a measurement aid: no A loads at all, the MMAs multiply whatever the stage holds)
            const uint32_t tx_bytes = static_cast<uint32_t>(p.n_tile) * 128u + ((p.a_mode == 1 || p.a_mode == 2) ? static_cast<uint32_t>(CL_A_BYTES) : 0u);
            int trace_n = 0;
            if (lane == 0) { tma_prefetch_desc(&p.tmap_b); if (p.a_mode != 0) tma_prefetch_desc(&p.tmap_a); }
            const uint32_t span = static_cast<uint32_t>(p.KW) * p.C;
            const uint32_t sub_bytes = CL_A_BYTES / (p.a_sub > 0 ? p.a_sub : 1);
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const ClItem wi = cl_decode(p, item);
                // im2col: base pixel of the tile's first row = its output position scaled by the stride, minus the padding
                int w0 = 0, h0 = 0, b0 = 0;
                uint32_t kh = 0, rem = 0;
                if (p.a_mode == 2) {
                    uint32_t b, r, oh, ow;
                    p.fd_HgWg.divmod(static_cast<uint32_t>(wi.tm) * CL_BLOCK_M, b, r);
                    p.fd_Wg.divmod(r, oh, ow);
                    b0 = static_cast<int>(b); h0 = static_cast<int>(oh) * p.stride - p.pad; w0 = static_cast<int>(ow) * p.stride - p.pad;
                    p.fd_span.divmod(static_cast<uint32_t>(wi.kb0) * CL_BLOCK_K, kh, rem);           // k = (kh, rem = kw * C + c)
                }
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    const long long tr_a = (p.trace && p.a_mode != 0) ? clock64() : 0;
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    const long long tr_b = (p.trace && p.a_mode != 0) ? clock64() : 0;
                    if (elect_one()) {
                        uint8_t* st = smem + stage * CL_STAGE_BYTES;
                        if (tma_role == 0) {
                            mbar_arrive_expect_tx(&bar_full[stage], tx_bytes);
                            tma_load_2d(st + CL_A_BYTES, &p.tmap_b, &bar_full[stage], kb * CL_BLOCK_K, wi.tn * p.n_tile);
                        }
                        if (p.a_mode == 1) {
                            tma_load_2d(st, &p.tmap_a, &bar_full[stage], kb * CL_BLOCK_K, wi.tm * CL_BLOCK_M);
                        } else if (p.a_mode == 2) {
                            const uint32_t kw = p.fd_C.div(rem), c0 = rem - kw * static_cast<uint32_t>(p.C);
                            for (int sidx = 0; sidx < p.a_sub; ++sidx)
                                if ((sidx == 3 ? 1 : sidx) == tma_role)                 // role 0: load 0, role 1: loads 1 and 3, role 2: load 2
                                    tma_load_im2col_4d(smem_u32(st) + sidx * sub_bytes, &p.tmap_a, &bar_full[stage], static_cast<int>(c0), w0, h0,
                                                       b0, static_cast<uint16_t>(kw + sidx), static_cast<uint16_t>(kh));
                        }
                    }
                    __syncwarp();
                    if (p.trace && p.a_mode != 0 && lane == 0 && tma_role == 0) cl_trace(p, 0, trace_n, tr_a, tr_b, clock64(), kb);
                    if (++stage == CL_STAGES) { stage = 0; phase ^= 1; }
                    rem += CL_BLOCK_K;
                    if (rem >= span) { rem -= span; ++kh; }
                }
            }
        }
    } else if (epi_group >= 0) {
        // ================================================================== epilogue: TMEM -> registers -> (shared) -> global
        // A TMEM lane is an output row, so after tcgen05.ld a thread holds consecutive CHANNELS of one pixel while a coalesced
        // store wants consecutive lanes on consecutive channels.  GEMM mode therefore passes every 32-column chunk through a
        // per-warp shared-memory tile: thread = row on the way in, 8 lanes x float4 = 128 contiguous bytes of a row on the way out
        // (measured before: 16 k cycles to drain one 128 x 128 tile with row-strided float4 stores).
        const int quad = warp & 3, row = quad * 32 + lane;
        float* stg = reinterpret_cast<float*>(smem_ctl + 256 + (epi_group * 4 + quad) * CL_EPI_WARP_BYTES);
        long long* stg_dst = reinterpret_cast<long long*>(stg + 32 * CL_EPI_LD);     // per row: element offset of its destination, -1 = no row
        volatile uint32_t* ws_flag = ws_flags + epi_group;
        int etrace_n = 0;
        int j_local = -1;                                                            // index of the item among this CTA's items
        // BatchNorm statistics of the stored values (launches with one N tile only, so that a thread sees the same columns in
        // every tile): fp32 sums in registers over all tiles of the CTA, folded once at the end.  Direct epilogue: st_*[e] =
        // column e of the thread's rows; staged epilogue: st_*[4 * chunk + e] = column 32 * chunk + 4 * col4 + e of its 8 rows per tile.
        float st_s[16], st_q[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) st_s[e] = st_q[e] = 0.0f;
        // bias of this thread's columns (4 per column group / chunk), re-loaded only when the N tile changes and always BEFORE the
        // wait for the accumulator: a global load issued inside the drain costs 300-600 cycles of exposed latency per chunk
        float4 bvs[4];
        int bias_tn = -1;
        const int col4 = lane & 7, rsub = lane >> 3;                         // staged read-back role: float4 column group, row within a group of 4
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            ++j_local;
            if (j_local % n_groups != epi_group) continue;                           // another epilogue group drains this tile
            const int acc = j_local % n_acc;
            const uint32_t acc_phase = static_cast<uint32_t>(j_local / n_acc) & 1u;
            const ClItem wi = cl_decode(p, item);
            const uint32_t m = static_cast<uint32_t>(wi.tm) * CL_BLOCK_M + row;
            const bool row_ok = m < static_cast<uint32_t>(MODE == CL_WGRAD ? p.m_valid : p.gemm_m);
            long long dst_off = -1;
            int qflags = 0;                                                  // quad epilogue: bit 0 = row 2i+1 exists, bit 1 = column 2j+1 exists
            if (MODE == CL_WGRAD) {
                dst_off = m;                                                 // dWcl[n][m]: lanes write consecutive m
                if (p.wg_taps > 0) {                                         // PyTorch layout [n][ci][tap], m = tap * Cin + ci
                    uint32_t tap, ci;
                    p.fd_C.divmod(m, tap, ci);
                    dst_off = static_cast<long long>(ci) * p.wg_taps + tap;
                }
            } else if (row_ok) {
                if (p.epi == CL_EPI_ROWS) {
                    dst_off = static_cast<long long>(m) * p.ldo;
                } else {
                    uint32_t b, rem, i, j;
                    p.fd_HgWg.divmod(m, b, rem);
                    p.fd_Wg.divmod(rem, i, j);
                    const int qi = 2 * static_cast<int>(i), qj = 2 * static_cast<int>(j);
                    dst_off = ((static_cast<long long>(b) * p.qH + qi) * p.qW + qj) * p.qC;
                    qflags = (qi + 1 < p.qH ? 1 : 0) | (qj + 1 < p.qW ? 2 : 0);
                }
            }
            const bool add_bias = p.bias != nullptr && (!p.atomic_out || wi.kb0 == 0);     // workspace split-K: atomic_out = 0, the finishing CTA adds it
            if (MODE == CL_GEMM && p.bias != nullptr && wi.tn != bias_tn) {
                bias_tn = wi.tn;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int cn = p.n_tile <= 16 ? 4 * k : 32 * k + 4 * col4;          // column inside the tile: direct / staged epilogue
                    const int n = wi.tn * p.n_tile + cn;
                    bvs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cn < p.n_tile && n < p.gemm_n) {
                        uint32_t cls = 0, ch = static_cast<uint32_t>(n);
                        if (p.epi == CL_EPI_QUAD) p.fd_qC.divmod(static_cast<uint32_t>(n), cls, ch);
                        bvs[k] = __ldg(reinterpret_cast<const float4*>(p.bias + ch));
                    }
                }
            }
            if (MODE == CL_GEMM) {
                __syncwarp();
                stg_dst[lane] = row_ok ? (dst_off * 4 + qflags) : -1;        // destination element offset << 2 | quad flags
                __syncwarp();
            }
            const long long tr_a = p.trace ? clock64() : 0;
            mbar_wait(&bar_tfull[acc], acc_phase);
            const long long tr_b = p.trace ? clock64() : 0;
            long long tr_ld = 0;                                             // trace: arrival of the first tcgen05.ld
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + acc * CL_MAX_N + (static_cast<uint32_t>(quad * 32) << 16);
            // ---- deterministic split-K: park the partial tile in the workspace, column-major (lane = row: coalesced)
            const float* ws_tile = nullptr;                                  // != NULL: this CTA finishes the tile from the workspace
            if (p.ws != nullptr && p.k_splits > 1) {
                const int tile = item / p.k_splits, split = item - tile * p.k_splits;
                float* wsp = p.ws + (static_cast<size_t>(tile) * p.k_splits + split) * (static_cast<size_t>(p.n_tile) * CL_BLOCK_M) + row;
#pragma unroll 1
                for (int c = 0; c < p.n_tile; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) __stcg(wsp + (c + e) * CL_BLOCK_M, __uint_as_float(v[e]));
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tempty[acc]);                // the accumulator is free again
                if (MODE == CL_WGRAD) continue;                              // summed and laid out by wgrad_finish_kernel
                __threadfence();
                named_bar_sync(1 + epi_group, 128);
                if (quad == 0 && lane == 0) {
                    const unsigned old = atomicAdd(p.ws_cnt + tile, 1u);
                    const bool last = old == static_cast<unsigned>(p.k_splits - 1);
                    if (last) p.ws_cnt[tile] = 0u;                           // every contributor has arrived: ready for the next launch
                    *ws_flag = last ? 1u : 0u;
                }
                named_bar_sync(1 + epi_group, 128);
                if (*ws_flag == 0u) continue;
                __threadfence();
                ws_tile = p.ws + static_cast<size_t>(tile) * p.k_splits * (static_cast<size_t>(p.n_tile) * CL_BLOCK_M) + row;
            }
            // 16 accumulator columns of this thread's row: from TMEM, or summed over the splits in split order
            auto load16 = [&](int c, uint32_t (&v)[16]) {
                if (ws_tile == nullptr) {
                    tmem_ld16(taddr + c, v);
                } else {
                    float a[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) a[e] = 0.0f;
                    for (int sp = 0; sp < p.k_splits; ++sp) {
                        const float* src = ws_tile + static_cast<size_t>(sp) * (static_cast<size_t>(p.n_tile) * CL_BLOCK_M) + c * CL_BLOCK_M;
#pragma unroll
                        for (int e = 0; e < 16; ++e) a[e] += __ldcg(src + e * CL_BLOCK_M);
                    }
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(a[e]);
                }
            };
            if (MODE == CL_WGRAD) {
#pragma unroll 1
                for (int c = 0; c < p.n_tile; c += 16) {
                    uint32_t v[16];
                    load16(c, v);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    const int nbase = wi.tn * p.n_tile + c;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (nbase + e < p.gemm_n) {
                            float* o = p.out + dst_off + static_cast<size_t>(nbase + e) * p.ldo;
                            if (p.atomic_out) atomicAdd(o, __uint_as_float(v[e]));
                            else *o = __uint_as_float(v[e]);
                        }
                }
            } else if (p.n_tile <= 16) {
                // 16-column tiles (the 8 / 16-channel layers): a row is only 64 bytes, the thread stores its own row directly
                uint32_t v[16];
                load16(0, v);
                tmem_ld_wait();
                if (p.trace) tr_ld = clock64();
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int n = wi.tn * p.n_tile + 4 * g;
                        if (n >= p.gemm_n) break;
                        long long cls_off = n;
                        if (p.epi == CL_EPI_QUAD) {
                            uint32_t cls, ch;
                            p.fd_qC.divmod(static_cast<uint32_t>(n), cls, ch);
                            const int need = static_cast<int>(((cls >> 1) & 1u) | ((cls & 1u) << 1));
                            if ((need & qflags) != need) continue;
                            cls_off = (static_cast<long long>(cls >> 1) * p.qW + (cls & 1u)) * p.qC + ch;
                        }
                        float4 r = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                               __uint_as_float(v[4 * g + 3]));
                        if (add_bias) { r.x += bvs[g].x; r.y += bvs[g].y; r.z += bvs[g].z; r.w += bvs[g].w; }
                        float* o = p.out + dst_off + cls_off;
                        if (p.atomic_out) {
                            atomicAdd(o, r.x);
                            if (n + 1 < p.gemm_n) atomicAdd(o + 1, r.y);
                            if (n + 2 < p.gemm_n) atomicAdd(o + 2, r.z);
                            if (n + 3 < p.gemm_n) atomicAdd(o + 3, r.w);
                        } else {
                            r.x = cl_act(r.x, p.slope, p.round_out); r.y = cl_act(r.y, p.slope, p.round_out);
                            r.z = cl_act(r.z, p.slope, p.round_out); r.w = cl_act(r.w, p.slope, p.round_out);
                            *reinterpret_cast<float4*>(o) = r;
                            // second statistic: squares (BatchNorm forward) or products with the activation at the same place (backward)
                            const float4 xq = p.stat_x != nullptr ? __ldg(reinterpret_cast<const float4*>(p.stat_x + (o - p.out))) : r;
                            st_s[4 * g] += r.x; st_s[4 * g + 1] += r.y; st_s[4 * g + 2] += r.z; st_s[4 * g + 3] += r.w;
                            st_q[4 * g] = fmaf(r.x, xq.x, st_q[4 * g]); st_q[4 * g + 1] = fmaf(r.y, xq.y, st_q[4 * g + 1]);
                            st_q[4 * g + 2] = fmaf(r.z, xq.z, st_q[4 * g + 2]); st_q[4 * g + 3] = fmaf(r.w, xq.w, st_q[4 * g + 3]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int ci = 0; ci < CL_MAX_N / 32; ++ci) {
                    const int c = ci * 32;
                    if (c >= p.n_tile) break;
                    uint32_t v[32];
                    load16(c, *reinterpret_cast<uint32_t(*)[16]>(v));
                    if (c + 16 < p.n_tile) load16(c + 16, *reinterpret_cast<uint32_t(*)[16]>(v + 16));
                    tmem_ld_wait();
                    if (p.trace && ci == 0) tr_ld = clock64();
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        *reinterpret_cast<uint4*>(stg + lane * CL_EPI_LD + 4 * g) = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                    __syncwarp();
                    const int n = wi.tn * p.n_tile + c + 4 * col4;           // this lane's 4 columns
                    const bool col_ok = c + 4 * col4 < p.n_tile && n < p.gemm_n;
                    long long cls_off = n;                                   // column part of the destination offset
                    int need = 0;                                            // quad: flag bits the destination pixel needs
                    if (p.epi == CL_EPI_QUAD && col_ok) {
                        uint32_t cls, ch;
                        p.fd_qC.divmod(static_cast<uint32_t>(n), cls, ch);
                        cls_off = (static_cast<long long>(cls >> 1) * p.qW + (cls & 1u)) * p.qC + ch;
                        need = static_cast<int>(((cls >> 1) & 1u) | ((cls & 1u) << 1));
                    }
                    const float4 bv = add_bias ? bvs[ci] : make_float4(0.f, 0.f, 0.f, 0.f);
                    // the shared-memory reads of four row groups are issued back to back, ahead of the stores that use them
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        long long dd[4];
                        float4 rr[4];
#pragma unroll
                        for (int it = 0; it < 4; ++it) dd[it] = stg_dst[(half * 4 + it) * 4 + rsub];
#pragma unroll
                        for (int it = 0; it < 4; ++it)
                            rr[it] = *reinterpret_cast<const float4*>(stg + ((half * 4 + it) * 4 + rsub) * CL_EPI_LD + 4 * col4);
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const long long d = dd[it];
                            if (!col_ok || d < 0) continue;
                            const int flags = static_cast<int>(d & 3);
                            if ((need & flags) != need) continue;                // odd H / W: the last row / column of a quad may not exist
                            float4 r = rr[it];
                            r.x += bv.x; r.y += bv.y; r.z += bv.z; r.w += bv.w;
                            float* o = p.out + (d >> 2) + cls_off;
                            if (p.atomic_out) {
                                atomicAdd(o, r.x);
                                if (n + 1 < p.gemm_n) atomicAdd(o + 1, r.y);
                                if (n + 2 < p.gemm_n) atomicAdd(o + 2, r.z);
                                if (n + 3 < p.gemm_n) atomicAdd(o + 3, r.w);
                            } else {
                                r.x = cl_act(r.x, p.slope, p.round_out); r.y = cl_act(r.y, p.slope, p.round_out);
                                r.z = cl_act(r.z, p.slope, p.round_out); r.w = cl_act(r.w, p.slope, p.round_out);
                                *reinterpret_cast<float4*>(o) = r;
                                const float4 xq = p.stat_x != nullptr ? __ldg(reinterpret_cast<const float4*>(p.stat_x + (o - p.out))) : r;
                                st_s[4 * ci] += r.x; st_s[4 * ci + 1] += r.y; st_s[4 * ci + 2] += r.z; st_s[4 * ci + 3] += r.w;
                                st_q[4 * ci] = fmaf(r.x, xq.x, st_q[4 * ci]); st_q[4 * ci + 1] = fmaf(r.y, xq.y, st_q[4 * ci + 1]);
                                st_q[4 * ci + 2] = fmaf(r.z, xq.z, st_q[4 * ci + 2]); st_q[4 * ci + 3] = fmaf(r.w, xq.w, st_q[4 * ci + 3]);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            if (p.trace && quad == 1 && lane == 0 && epi_group == 0) cl_trace(p, 2, etrace_n, tr_a, tr_b, clock64(), item, tr_ld);
            if (ws_tile != nullptr) continue;                                // the accumulator was released when the partial was parked
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);
        }
        if (with_stats) {
            const bool direct = p.n_tile <= 16;
            // direct: fold all 32 lanes (rows), lane e then owns column e; staged: fold the 4 lanes (xor 8, 16) that share columns
#pragma unroll
            for (int e = 0; e < 16; ++e) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    if (o >= 8 || direct) {
                        st_s[e] += __shfl_xor_sync(0xffffffffu, st_s[e], o);
                        st_q[e] += __shfl_xor_sync(0xffffffffu, st_q[e], o);
                    }
            }
            // per-warp partial sums of column n -> this warp's (now idle) staging tile: stg[2n] = sum, stg[2n + 1] = sum of squares
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = direct ? e : 32 * (e >> 2) + 4 * (lane & 7) + (e & 3);       // column of entry e (the launch has one N tile)
                const bool mine = direct ? lane == e : lane < 8;
                if (mine && n < p.n_tile) { stg[2 * n] = st_s[e]; stg[2 * n + 1] = st_q[e]; }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 4 * CL_MAX_N);
    if (with_stats) {
        // fold the 4 epilogue warps (and, quad epilogue, the columns (class, c) of the 4 pixel classes) and add to the global sums
        const float* part = reinterpret_cast<const float*>(smem_ctl + 256);
        for (int i = threadIdx.x; i < 2 * p.stat_c; i += CL_THREADS) {
            double v = 0.0;
            for (int n = i >> 1; n < p.gemm_n; n += p.stat_c)
                for (int q = 0; q < 4 * n_groups; ++q) v += static_cast<double>(part[q * (CL_EPI_WARP_BYTES / 4) + 2 * n + (i & 1)]);
            atomicAdd(p.stats + i, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight re-packing
// PyTorch conv weight w[co][ci][kh][kw]  ->  TF32-rounded operand matrices:
//   wf[co][(kh, kw, ci)]                          forward  (K = KH*KW*Cin)
//   wq[(ph, pw, ci)][(a, bb, co)]                 data gradient of a 4x4 / stride 2 / even-pad convolution: output pixel parity
//                                                 (ph, pw), tap (a, bb) reads dy at (i + a, j + bb) and kernel element
//                                                 (ph + 2 (1 - a), pw + 2 (1 - bb));  K = 4 * Cout
//   wt[ci][co]                                    data gradient of a 1x1 convolution
__global__ void __launch_bounds__(256) cl_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wq,
                                                             int Cout, int Cin, int KH, int KW, int quad) {
    const int taps = KH * KW;
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        // i enumerates the DESTINATION wf index (co, kh, kw, ci) so that writes are coalesced
        const int ci = static_cast<int>(i % Cin);
        long long r = i / Cin;
        const int tap = static_cast<int>(r % taps), co = static_cast<int>(r / taps);
        const float v = to_tf32_rna(__ldg(w + (static_cast<long long>(co) * Cin + ci) * taps + tap));
        if (wf != nullptr) wf[i] = v;
        if (wq != nullptr) {
            if (quad) {
                const int kh = tap / KW, kw = tap % KW;
                const int ph = kh & 1, a = 1 - (kh >> 1), pw = kw & 1, bb = 1 - (kw >> 1);
                wq[(static_cast<long long>((ph * 2 + pw) * Cin + ci)) * (4LL * Cout) + (a * 2 + bb) * Cout + co] = v;
            } else {
                wq[static_cast<long long>(ci) * Cout + co] = v;          // 1x1: transpose
            }
        }
    }
}

// dWcl[co][(kh, kw, ci)] -> PyTorch layout dw[co][ci][kh][kw]
__global__ void __launch_bounds__(256) cl_unpack_dw_kernel(const float* __restrict__ dwcl, float* __restrict__ dw, int Cout, int Cin, int taps) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int tap = static_cast<int>(i % taps);
        long long r = i / taps;
        const int ci = static_cast<int>(r % Cin), co = static_cast<int>(r / Cin);
        dw[i] = __ldg(dwcl + (static_cast<long long>(co) * taps + tap) * Cin + ci);
    }
}

// dst[r][c] = TF32-rounded src[r][c] for c < cols, 0 for cols <= c < ldd (re-pitches rows to a 16-byte-aligned length)
__global__ void __launch_bounds__(256) round_copy_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows, int cols) {
    const long long total = static_cast<long long>(rows) * ldd;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int c = static_cast<int>(i % ldd);
        const long long r = i / ldd;
        dst[i] = c < cols ? to_tf32_rna(__ldg(src + r * lds + c)) : 0.0f;
    }
}

extern long long* g_conv_trace_ptr;          // set by pgv_debug_set_conv_trace (pgv_conv_tc.cu)

template <int MODE>
static int launch_conv_cl(const pgv_handle* h, ConvClParams& p, cudaStream_t stream) {
    p.trace = g_conv_trace_ptr;
    p.fd_splits.init(p.k_splits);
    static bool configured = false;
    if (!configured) {
        PGV_CUDA(cudaFuncSetAttribute(conv_cl_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, CL_SMEM));
        configured = true;
    }
    const long long items = static_cast<long long>(p.m_tiles) * p.n_tiles * p.k_splits;
    const int grid = static_cast<int>(items < h->sm_count ? items : h->sm_count);
    conv_cl_kernel<MODE><<<grid, CL_THREADS, CL_SMEM, stream>>>(p);
    PGV_LAUNCH_CHECK();
    return 0;
}

static int cl_pick_n_tile(int n, int granule) {
    int t = (n + granule - 1) / granule * granule;
    if (t > CL_MAX_N) {
        const int parts = ceil_div(n, CL_MAX_N);
        t = ceil_div(ceil_div(n, parts), granule) * granule;
    }
    return t;
}

int g_conv_groups = 0;        // debug: 0 = choose, 1 = one epilogue group everywhere, 2 = three groups for the quad epilogue only

static void cl_set_ring(ConvClParams& p, int sm_count) {
    // three epilogue groups (GEMM mode with a TMA-fed A operand only: the producer warps are free) when every CTA drains >= 6 tiles
    const long long items = static_cast<long long>(p.m_tiles) * p.n_tiles * std::max(p.k_splits, 1);
    p.n_groups = (p.a_mode != 0 && items >= 6LL * sm_count) ? CL_MAX_GROUPS : 1;
    if (g_conv_groups == 1 || (g_conv_groups == 2 && p.epi != CL_EPI_QUAD)) p.n_groups = 1;      // (pgv_debug_set_conv_groups)
    p.stage_bytes = static_cast<int>(align_up(static_cast<size_t>(CL_A_BYTES) + static_cast<size_t>(p.n_tile) * 128, 1024));
    p.stages = std::min(CL_MAX_STAGES, cl_ring_bytes(p.n_groups) / p.stage_bytes);
}

// Workspace of the deterministic split-K paths: [CL_WS_COUNTERS x u32 tile counters (zero on entry, self-resetting)][fp32 partial tiles].
constexpr size_t CL_WS_COUNTERS = 4096, CL_WS_HEADER = CL_WS_COUNTERS * sizeof(unsigned);

int g_conv_a_mode = -1;       // debug / A-B switch (pgv_debug_set_conv_a_mode): -1 = choose, 0 = force the cp.async gather

// im2col tensor map over a channels-last activation [B, H, W, C] (dims in (C, W, H, N) order, as the TMA unit wants them): base pixels
// run over the Hg x Wg output grid with the convolution stride, starting at -pad; a load fetches `cpp` channels of 128 consecutive
// base pixels shifted by the filter offset.  Parameters as CUTLASS derives them (cute/atom/copy_traits_sm90_im2col.hpp,
// cutlass/conv/collective/detail.hpp), with the upper corner chosen so that the box holds exactly Wg x Hg base pixels.
static int make_tmap_im2col(const pgv_handle* h, CUtensorMap* out, const float* base, int B, int H, int W, int C, int stride, int pad, int Hg,
                            int Wg, int cpp, CUtensorMapSwizzle swz) {
    if (!h->encode_im2col) return set_error(-2, "pgv handle has no cuTensorMapEncodeIm2col");
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(W) * C * 4, static_cast<cuuint64_t>(H) * W * C * 4};
    const int lower[2] = {-pad, -pad};
    const int upper[2] = {(Wg - 1) * stride - pad + 1 - W, (Hg - 1) * stride - pad + 1 - H};
    if (upper[0] < -128 || upper[0] > 127 || upper[1] < -128 || upper[1] > 127 || pad > 128) return set_error(-1, "im2col corners out of range");
    const cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
    CUresult r = h->encode_im2col(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, lower, upper,
                                  static_cast<cuuint32_t>(cpp), CL_BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(static_cast<int>(r), "cuTensorMapEncodeIm2col failed with CUresult %d", static_cast<int>(r));
    // drivers up to CUDA 13.1 set a descriptor bit that breaks im2col loads from tensors smaller than 128 KB (same fix-up as CUTLASS)
    if (h->driver_version <= 13010 && static_cast<size_t>(B) * H * W * C * 4 < 131072) reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
    return 0;
}

// Split count of a GEMM-mode launch with the workspace: minimises (waves of CTAs) x (cost of one item), in k-block units (~0.3 us):
// every item pays ~8 for pipeline fill + epilogue; a split item additionally ~6 for parking its partial tile, and the CTA that
// finishes the tile ~4 per split for reading them back (measured: 60 tiles x 128 k-blocks gain 25 % from 2 splits, 88 x 64 lose).
static int cl_pick_splits(int tiles, int kb_total, int sm_count, int max_splits) {
    int best = 1;
    long long best_cost = static_cast<long long>(ceil_div(tiles, sm_count)) * (kb_total + 8);
    for (int sp = 2; sp <= max_splits && sp * 8 <= kb_total; ++sp) {
        const int kbs = ceil_div(kb_total, sp), eff = ceil_div(kb_total, kbs);
        const long long cost = static_cast<long long>(ceil_div(tiles * eff, sm_count)) * (kbs + 8 + 6 + 4 * eff);
        if (cost * 10 < best_cost * 9) { best_cost = cost; best = eff; }
    }
    return best;
}

// Shared by forward and data gradient: out = act(bias + gather(in) * Bw^T).
static int conv_cl_gemm(const pgv_handle* h, const char* who, const float* in, const float* bw, const float* bias, float* out, int B, int H,
                        int W, int C, int KH, int KW, int stride, int pad, int Hg, int Wg, int N, int epi, int qH, int qW, int qC, float slope,
                        int round_out, size_t out_elems, double* stats, const float* stat_x, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (stat_x != nullptr && (stats == nullptr || (reinterpret_cast<uintptr_t>(stat_x) & 15)))
        return set_error(-1, "%s: the BatchNorm-backward sums need a statistics buffer and a 16-byte aligned activation", who);
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || N <= 0 || Hg <= 0 || Wg <= 0 || KH <= 0 || KH > 8 || KW <= 0 || KW > 8 || stride <= 0 || pad < 0)
        return set_error(-1, "%s: bad geometry", who);
    const bool unaligned_n = N % 4 != 0;           // only the atomic (scalar) epilogue can write rows whose pitch is not 16-byte aligned
    if (C % 4 != 0 || (unaligned_n && (slope >= 0.0f || round_out || bias != nullptr || epi != CL_EPI_ROWS)))
        return set_error(-1, "%s: channels-last kernels need C %% 4 == 0 and (N %% 4 == 0 or a linear, bias-free epilogue) (C=%d, N=%d)", who, C, N);
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(bw) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias) |
         reinterpret_cast<uintptr_t>(ws)) & 15)
        return set_error(-1, "%s: pointers must be 16-byte aligned", who);
    const long long m = static_cast<long long>(B) * Hg * Wg;
    if (m >= (1LL << 31) - CL_BLOCK_M) return set_error(-1, "%s: too many output pixels", who);
    ConvClParams p;
    memset(&p, 0, sizeof(p));
    p.a = in; p.b = bw; p.bias = bias; p.out = out;
    p.B = B; p.H = H; p.W = W; p.C = C; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad; p.Hg = Hg; p.Wg = Wg;
    p.gemm_m = static_cast<int>(m); p.gemm_n = N; p.gemm_k = KH * KW * C;
    p.n_tile = cl_pick_n_tile(N, 16); p.n_tiles = ceil_div(N, p.n_tile);
    p.m_tiles = ceil_div(p.gemm_m, CL_BLOCK_M);
    p.kb_total = ceil_div(p.gemm_k, CL_BLOCK_K); p.kb_per_split = p.kb_total; p.k_splits = 1;      // a K tail is zero-filled by the gather masks / TMA
    p.epi = epi; p.ldo = N; p.qH = qH; p.qW = qW; p.qC = qC;
    p.slope = slope; p.round_out = round_out;
    p.fd_HgWg.init(Hg * Wg); p.fd_Wg.init(Wg); p.fd_span.init(KW * C); p.fd_C.init(C); p.fd_bblocks.init(1); p.fd_qC.init(qC > 0 ? qC : 1); p.fd_ntiles.init(p.n_tiles);
    p.a_sub = 1; p.a_sbo = 1024; p.a_layout = 2;
    // ---- A operand: TMA where the geometry allows it
    const bool one_by_one = KH == 1 && KW == 1 && stride == 1 && pad == 0 && H == Hg && W == Wg;
    if (g_conv_a_mode == -2) {
        p.a_mode = 4;
    } else if (g_conv_a_mode != 0) {
        if (one_by_one) {
            const uint64_t ad[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(m)}, as[1] = {static_cast<uint64_t>(C) * 4};
            const uint32_t abox[2] = {CL_BLOCK_K, CL_BLOCK_M};
            if (int rc = make_tmap_f32(h, &p.tmap_a, in, 2, ad, as, abox)) return rc;
            p.a_mode = 1;
        } else if ((C % CL_BLOCK_K == 0 || C == 8 || C == 16) && (KW * C) % CL_BLOCK_K == 0 && h->encode_im2col != nullptr) {
            const int cpp = C < CL_BLOCK_K ? C : CL_BLOCK_K;
            const CUtensorMapSwizzle swz = cpp == 8 ? CU_TENSOR_MAP_SWIZZLE_32B : (cpp == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
            if (int rc = make_tmap_im2col(h, &p.tmap_a, in, B, H, W, C, stride, pad, Hg, Wg, cpp, swz)) return rc;
            p.a_mode = 2;
            p.a_sub = CL_BLOCK_K / cpp;
            p.a_sbo = 8 * cpp * 4;
            p.a_layout = cpp == 8 ? 6u : (cpp == 16 ? 4u : 2u);
        }
    }
    // ---- split K across CTAs when the tiles alone do not fill the GPU
    const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles;
    const bool linear_epi = slope < 0.0f && !round_out;
    if (ws != nullptr && !unaligned_n && tiles <= static_cast<long long>(CL_WS_COUNTERS) && ws_bytes > CL_WS_HEADER) {
        // deterministic: partial tiles through the workspace, the last CTA of a tile sums them in split order
        const size_t tile_bytes = static_cast<size_t>(p.n_tile) * CL_BLOCK_M * sizeof(float);
        const long long fit = static_cast<long long>((ws_bytes - CL_WS_HEADER) / (tile_bytes * tiles));
        const int max_splits = static_cast<int>(std::min<long long>(fit, 16));
        const int splits = stats != nullptr && p.n_tiles != 1 ? 1 : cl_pick_splits(static_cast<int>(tiles), p.kb_total, h->sm_count, max_splits);
        if (splits > 1) {
            p.kb_per_split = ceil_div(p.kb_total, splits);
            p.k_splits = ceil_div(p.kb_total, p.kb_per_split);
            p.ws_cnt = static_cast<unsigned*>(ws);
            p.ws = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + CL_WS_HEADER);
        }
    } else if (linear_epi && tiles < h->sm_count && p.kb_total >= 16) {
        // no workspace: long-K problems with few tiles and a linear epilogue use fp32 atomics into a zero-filled output
        int splits = static_cast<int>((h->sm_count * 2) / tiles);
        if (splits > p.kb_total / 8) splits = p.kb_total / 8;
        if (splits > 1) {
            p.kb_per_split = ceil_div(p.kb_total, splits);
            p.k_splits = ceil_div(p.kb_total, p.kb_per_split);
            p.atomic_out = 1;
        }
    }
    if (unaligned_n) p.atomic_out = 1;
    if (stats != nullptr) {
        p.stats = stats;
        p.stat_x = stat_x;
        p.stat_c = epi == CL_EPI_QUAD ? qC : N;
        if (p.atomic_out || p.n_tiles != 1 || p.stat_c % 4 != 0)
            return set_error(-1, "%s: output statistics need a non-atomic launch with one N tile (N=%d <= %d)", who, N, CL_MAX_N);
        PGV_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * p.stat_c, stream));
    }
    if (p.atomic_out) PGV_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * out_elems, stream));
    const uint64_t bd[2] = {static_cast<uint64_t>(p.gemm_k), static_cast<uint64_t>(N)}, bs[1] = {static_cast<uint64_t>(p.gemm_k) * 4};
    const uint32_t bbox[2] = {CL_BLOCK_K, static_cast<uint32_t>(p.n_tile)};
    if (int rc = make_tmap_f32(h, &p.tmap_b, bw, 2, bd, bs, bbox)) return rc;
    cl_set_ring(p, h->sm_count);
    return launch_conv_cl<CL_GEMM>(h, p, stream);
}

// Sums the split-K partials of a weight gradient and writes it in its final layout.  Block = 32 consecutive elements x 8 split lanes:
// lane l adds the partials of splits l, l + 8, ... in order, then the 8 lane sums are added in lane order - the same order in every
// run, so the result is reproducible bit for bit.
// Few splits (the deep layers: 2-16 partial tiles of up to 2 M elements): one thread per element, splits added in order.
__global__ void __launch_bounds__(256) wgrad_finish_few_kernel(const float* __restrict__ ws, float* __restrict__ out, int m_pad, int m_valid, int N,
                                                               int n_tile, int n_tiles, int splits, int ldo, int cin, int taps) {
    const long long total = static_cast<long long>(m_pad) * N;
    const size_t tile_elems = static_cast<size_t>(n_tile) * CL_BLOCK_M;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int m = static_cast<int>(i % m_pad), n = static_cast<int>(i / m_pad);
        if (m >= m_valid) continue;
        const int tm = m / CL_BLOCK_M, row = m % CL_BLOCK_M, tn = n / n_tile, col = n % n_tile;
        const float* src = ws + (static_cast<size_t>(tm) * n_tiles + tn) * splits * tile_elems + static_cast<size_t>(col) * CL_BLOCK_M + row;
        float acc = 0.0f;
#pragma unroll 4
        for (int sp = 0; sp < splits; ++sp) acc += __ldcg(src + sp * tile_elems);
        const long long dst = taps > 0 ? (static_cast<long long>(m % cin) * taps + m / cin) : m;
        out[dst + static_cast<long long>(n) * ldo] = acc;
    }
}

// Many splits (the thin layers: one or two tiles, ~300 partials each):
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ ws, float* __restrict__ out, int m_pad, int m_valid, int N,
                                                           int n_tile, int n_tiles, int splits, int ldo, int cin, int taps) {
    __shared__ float part[8][33];
    const int ex = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const long long total = static_cast<long long>(m_pad) * N;
    const size_t tile_elems = static_cast<size_t>(n_tile) * CL_BLOCK_M;
    for (long long base = blockIdx.x * 32LL; base < total; base += 32LL * gridDim.x) {
        const long long i = base + ex;                       // m_pad is a multiple of 128: the 32 elements of a block share n and the tile
        const int m = static_cast<int>(i % m_pad), n = static_cast<int>(i / m_pad);
        const int tm = m / CL_BLOCK_M, row = m % CL_BLOCK_M, tn = n / n_tile, col = n % n_tile;
        const float* src = ws + (static_cast<size_t>(tm) * n_tiles + tn) * splits * tile_elems + static_cast<size_t>(col) * CL_BLOCK_M + row;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;      // four independent chains keep the loads in flight; combined in a fixed order
        if (i < total) {
            int sp = sl;
            for (; sp + 24 < splits; sp += 32) {
                a0 += __ldcg(src + sp * tile_elems); a1 += __ldcg(src + (sp + 8) * tile_elems);
                a2 += __ldcg(src + (sp + 16) * tile_elems); a3 += __ldcg(src + (sp + 24) * tile_elems);
            }
            for (; sp < splits; sp += 8) a0 += __ldcg(src + sp * tile_elems);
        }
        const float acc = (a0 + a1) + (a2 + a3);
        part[sl][ex] = acc;
        __syncthreads();
        if (sl == 0 && i < total && m < m_valid) {
            float v = part[0][ex];
#pragma unroll
            for (int l = 1; l < 8; ++l) v += part[l][ex];
            const long long dst = taps > 0 ? (static_cast<long long>(m % cin) * taps + m / cin) : m;
            out[dst + static_cast<long long>(n) * ldo] = v;
        }
        __syncthreads();
    }
}

// Tiled form of cl_prep_weights_kernel for up to 16 taps: a CTA moves an 8 (co) x 32 (ci) x taps block through shared memory, so
// the source is read in contiguous runs (32 ci x taps floats per co) and both destinations are written in whole sectors
// (wf: 32 ci per (co, tap); wq: 8 co per (ci, tap)); the one-element-per-thread kernel above writes wq with a 4 Cout stride.
// 17 KB of shared memory on purpose: these launches run on a side stream UNDER the front end, whose persistent DFT GEMM CTAs hold
// 198 KB per SM - with a 35 KB tile nothing co-resided and the copies queued up behind the front end (encoder forward +0.15 ms).
constexpr int PW_TCO = 8, PW_TCI = 32, PW_MAXT = 16, PW_ROW = PW_TCI * (PW_MAXT + 1) + 1;
__global__ void __launch_bounds__(256) cl_prep_weights_tiled_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wq,
                                                                   int Cout, int Cin, int KH, int KW, int quad) {
    __shared__ float s[PW_TCO][PW_ROW];
    const int taps = KH * KW, tp = taps + 1, co0 = blockIdx.y * PW_TCO, ci0 = blockIdx.x * PW_TCI;
    const int nco = min(PW_TCO, Cout - co0), nci = min(PW_TCI, Cin - ci0), warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int co = warp; co < nco; co += 8) {
        const float* src = w + (static_cast<size_t>(co0 + co) * Cin + ci0) * taps;
        for (int e = lane; e < nci * taps; e += 32) s[co][(e / taps) * tp + e % taps] = to_tf32_rna(__ldg(src + e));
    }
    __syncthreads();
    if (wf != nullptr)
        for (int row = warp; row < nco * taps; row += 8) {
            const int co = row / taps, tap = row - co * taps;
            if (lane < nci) wf[(static_cast<size_t>(co0 + co) * taps + tap) * Cin + ci0 + lane] = s[co][lane * tp + tap];
        }
    if (wq != nullptr) {
        const int co = lane & 7;
        for (int row = warp * 4 + (lane >> 3); row < nci * taps; row += 32) {
            const int ci = row / taps, tap = row - ci * taps;
            if (co >= nco) continue;
            const float v = s[co][ci * tp + tap];
            if (quad) {
                const int kh = tap / KW, kw = tap % KW;
                const int ph = kh & 1, a = 1 - (kh >> 1), pw = kw & 1, bb = 1 - (kw >> 1);
                wq[(static_cast<size_t>((ph * 2 + pw) * Cin + ci0 + ci)) * (4 * static_cast<size_t>(Cout)) + (a * 2 + bb) * Cout + co0 + co] = v;
            } else {
                wq[static_cast<size_t>(ci0 + ci) * Cout + co0 + co] = v;
            }
        }
    }
}

}  // namespace pgv

using namespace pgv;

extern "C" {

int pgv_conv_cl_prep_weights(const float* w, float* wf, float* wq, int Cout, int Cin, int KH, int KW, int stride, int pad, pgv_stream_t stream) {
    PGV_CHECK_ARG(w && (wf || wq) && Cout > 0 && Cin > 0 && KH > 0 && KW > 0, "pgv_conv_cl_prep_weights: bad argument");
    const int quad = (KH == 4 && KW == 4 && stride == 2 && pad == 2) ? 1 : 0;
    PGV_CHECK_ARG(wq == nullptr || quad || (KH == 1 && KW == 1 && stride == 1 && pad == 0),
                  "pgv_conv_cl_prep_weights: the data-gradient matrix exists for 4x4/stride 2/pad 2 and 1x1/stride 1 only");
    if (KH * KW <= PW_MAXT && ceil_div(Cout, PW_TCO) <= 65535) {
        const dim3 grid(ceil_div(Cin, PW_TCI), ceil_div(Cout, PW_TCO));
        cl_prep_weights_tiled_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, wf, wq, Cout, Cin, KH, KW, quad);
    } else {
        const long long total = static_cast<long long>(Cout) * Cin * KH * KW;
        const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 8));
        cl_prep_weights_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, wf, wq, Cout, Cin, KH, KW, quad);
    }
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_debug_set_conv_groups(int mode) { g_conv_groups = mode; return 0; }

int pgv_conv_cl_supported(int Cin, int Cout, int KH, int KW, int stride, int pad) {
    if (Cin % 4 != 0 || Cout % 4 != 0 || (KH * KW * Cin) % 32 != 0) return 0;
    if (KH == 4 && KW == 4 && stride == 2 && pad == 2) return (Cout % 8 == 0 && Cin % 8 == 0) ? 1 : 0;
    if (KH == 1 && KW == 1 && stride == 1 && pad == 0) return (Cout % 32 == 0 && Cin % 32 == 0) ? 1 : 0;
    return 0;
}

int pgv_conv_cl_fwd_bn(pgv_handle* h, const float* x, const float* wf, const float* bias, float* y, int B, int H, int W, int Cin, int Cout, int KH,
                       int KW, int stride, int pad, int Ho, int Wo, float lrelu_slope, int round_out, double* bn_sums, const float* bn_bwd_x, void* ws,
                       size_t ws_bytes, pgv_stream_t stream) {
    PGV_CHECK_ARG(h && x && wf && y, "pgv_conv_cl_fwd: NULL argument");
    PGV_CHECK_ARG((Ho - 1) * stride - 2 * pad + KH <= H + stride && (Wo - 1) * stride - 2 * pad + KW <= W + stride,
                  "pgv_conv_cl_fwd: output %dx%d does not fit input %dx%d", Ho, Wo, H, W);
    return conv_cl_gemm(h, "pgv_conv_cl_fwd", x, wf, bias, y, B, H, W, Cin, KH, KW, stride, pad, Ho, Wo, Cout, CL_EPI_ROWS, 0, 0, 0, lrelu_slope,
                        round_out, static_cast<size_t>(B) * Ho * Wo * Cout, bn_sums, bn_bwd_x, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int pgv_conv_cl_fwd(pgv_handle* h, const float* x, const float* wf, const float* bias, float* y, int B, int H, int W, int Cin, int Cout, int KH,
                    int KW, int stride, int pad, int Ho, int Wo, float lrelu_slope, int round_out, pgv_stream_t stream) {
    return pgv_conv_cl_fwd_bn(h, x, wf, bias, y, B, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo, lrelu_slope, round_out, nullptr, nullptr, nullptr, 0, stream);
}

int pgv_conv_cl_dgrad_bn(pgv_handle* h, const float* dy, const float* wq, const float* bias, float* dx, int B, int H, int W, int Cin, int Cout,
                         int KH, int KW, int stride, int pad, int Ho, int Wo, float lrelu_slope, int round_out, double* bn_sums, const float* bn_bwd_x,
                         void* ws, size_t ws_bytes, pgv_stream_t stream) {
    PGV_CHECK_ARG(h && dy && wq && dx, "pgv_conv_cl_dgrad: NULL argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t out_elems = static_cast<size_t>(B) * H * W * Cin;
    if (KH == 1 && KW == 1 && stride == 1 && pad == 0) {
        PGV_CHECK_ARG(H == Ho && W == Wo, "pgv_conv_cl_dgrad: 1x1 geometry mismatch");
        return conv_cl_gemm(h, "pgv_conv_cl_dgrad", dy, wq, bias, dx, B, Ho, Wo, Cout, 1, 1, 1, 0, Ho, Wo, Cin, CL_EPI_ROWS, 0, 0, 0, lrelu_slope,
                            round_out, out_elems, bn_sums, bn_bwd_x, ws, ws_bytes, s);
    }
    PGV_CHECK_ARG(KH == 4 && KW == 4 && stride == 2, "pgv_conv_cl_dgrad: only 4x4/stride 2/pad 2 and 1x1/stride 1");
    // pad = 2: the four pixels (2i + ph, 2j + pw) of a quad all read the 2x2 patch of dy whose corner is (i, j)
    PGV_CHECK_ARG(pad == 2, "pgv_conv_cl_dgrad: pad %d not implemented", pad);
    const int Hq = (H + 1) / 2, Wq = (W + 1) / 2;
    return conv_cl_gemm(h, "pgv_conv_cl_dgrad", dy, wq, bias, dx, B, Ho, Wo, Cout, 2, 2, 1, 0, Hq, Wq, 4 * Cin, CL_EPI_QUAD, H, W, Cin, lrelu_slope,
                        round_out, out_elems, bn_sums, bn_bwd_x, ws, ws_bytes, s);
}

int pgv_conv_cl_dgrad(pgv_handle* h, const float* dy, const float* wq, const float* bias, float* dx, int B, int H, int W, int Cin, int Cout,
                      int KH, int KW, int stride, int pad, int Ho, int Wo, float lrelu_slope, int round_out, pgv_stream_t stream) {
    return pgv_conv_cl_dgrad_bn(h, dy, wq, bias, dx, B, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo, lrelu_slope, round_out, nullptr, nullptr, nullptr, 0,
                                stream);
}

// taps > 0: `out` is the PyTorch weight layout [Cout][Cin][taps]; else out[n * ldo + m] for m < m_valid.
static int conv_cl_wgrad_impl(pgv_handle* h, const char* who, const float* x, const float* dy, float* out, int ldo, int m_valid, int taps, int B,
                              int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo, void* ws, size_t ws_bytes,
                              cudaStream_t stream) {
    if (!(h && x && dy && out)) return set_error(-1, "%s: NULL argument", who);
    if (!(B > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0)) return set_error(-1, "%s: bad geometry", who);
    if (Cin % 4 != 0 || Cout % 4 != 0) return set_error(-1, "%s: needs Cin %% 4 == 0 and Cout %% 4 == 0 (Cin=%d, Cout=%d)", who, Cin, Cout);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(ws)) & 15)
        return set_error(-1, "%s: operand pointers must be 16-byte aligned", who);
    ConvClParams p;
    memset(&p, 0, sizeof(p));
    p.a = x; p.b = dy; p.out = out;
    p.B = B; p.H = H; p.W = W; p.C = Cin; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad; p.Hg = Ho; p.Wg = Wo;
    p.gemm_m = KH * KW * Cin; p.gemm_n = Cout; p.ldo = ldo; p.m_valid = m_valid;
    p.wg_cin = Cin; p.wg_taps = taps;
    p.n_tile = cl_pick_n_tile(Cout, 32); p.n_tiles = ceil_div(Cout, p.n_tile);
    if (p.n_tile != 32 && p.n_tile != 64 && p.n_tile != 128) { p.n_tile = p.n_tile <= 64 ? 64 : 128; p.n_tiles = ceil_div(Cout, p.n_tile); }
    p.m_tiles = ceil_div(p.gemm_m, CL_BLOCK_M);
    p.k_splits = 1;
    cl_set_ring(p, h->sm_count);
    p.bblocks = ceil_div(B, 32);
    p.kb_total = Ho * Wo * p.bblocks;
    const int tiles = p.m_tiles * p.n_tiles;
    int splits = ceil_div(h->sm_count * 2, tiles);
    if (splits > p.kb_total / 4) splits = p.kb_total / 4;
    const size_t tile_bytes = static_cast<size_t>(p.n_tile) * CL_BLOCK_M * sizeof(float);
    const bool use_ws = ws != nullptr && ws_bytes > CL_WS_HEADER + tile_bytes * tiles;
    if (use_ws) {                                       // deterministic: partials through the workspace, summed by wgrad_finish_kernel
        const long long fit = static_cast<long long>((ws_bytes - CL_WS_HEADER) / (tile_bytes * tiles));
        if (splits > fit) splits = static_cast<int>(fit);
    }
    if (splits < 1) splits = 1;
    p.kb_per_split = ceil_div(p.kb_total, splits);
    p.k_splits = ceil_div(p.kb_total, p.kb_per_split);
    p.slope = -1.0f;
    p.fd_HgWg.init(Ho * Wo); p.fd_Wg.init(Wo); p.fd_span.init(KW * Cin); p.fd_C.init(Cin); p.fd_bblocks.init(p.bblocks); p.fd_qC.init(1); p.fd_ntiles.init(p.n_tiles);
    if (p.k_splits > 1 && use_ws) {
        p.ws = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + CL_WS_HEADER);
        if (int rc = launch_conv_cl<CL_WGRAD>(h, p, stream)) return rc;
        const long long total = static_cast<long long>(p.m_tiles) * CL_BLOCK_M * Cout;
        if (p.k_splits <= 16) {
            const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 8));
            wgrad_finish_few_kernel<<<grid, 256, 0, stream>>>(p.ws, out, p.m_tiles * CL_BLOCK_M, m_valid, Cout, p.n_tile, p.n_tiles, p.k_splits,
                                                              ldo, Cin, taps);
        } else {
            const int grid = static_cast<int>(std::min<long long>((total + 31) / 32, 148LL * 16));
            wgrad_finish_kernel<<<grid, 256, 0, stream>>>(p.ws, out, p.m_tiles * CL_BLOCK_M, m_valid, Cout, p.n_tile, p.n_tiles, p.k_splits, ldo,
                                                          Cin, taps);
        }
        PGV_LAUNCH_CHECK();
        return 0;
    }
    p.atomic_out = p.k_splits > 1 ? 1 : 0;
    if (p.atomic_out) {
        if (taps > 0) PGV_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * static_cast<size_t>(ldo) * Cout, stream));
        else PGV_CUDA(cudaMemset2DAsync(out, sizeof(float) * ldo, 0, sizeof(float) * m_valid, Cout, stream));
    }
    return launch_conv_cl<CL_WGRAD>(h, p, stream);
}

/* dw in the PyTorch layout [Cout][Cin][KH][KW] (oihw_layout = 1) or as the forward matrix [Cout][(kh, kw, ci)] (oihw_layout = 0) */
int pgv_conv_cl_wgrad(pgv_handle* h, const float* x, const float* dy, float* dw, int oihw_layout, int B, int H, int W, int Cin, int Cout, int KH,
                      int KW, int stride, int pad, int Ho, int Wo, void* ws, size_t ws_bytes, pgv_stream_t stream) {
    return conv_cl_wgrad_impl(h, "pgv_conv_cl_wgrad", x, dy, dw, KH * KW * Cin, KH * KW * Cin, oihw_layout ? KH * KW : 0, B, H, W, Cin, Cout, KH, KW,
                              stride, pad, Ho, Wo, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

/* ---- nn.Linear on the same kernel (operands TF32-rounded by the caller, see pgv_round_copy / pgv_transpose_inner) ---- */
int pgv_linear_cl_fwd(pgv_handle* h, const float* x, const float* wr, const float* bias, float* y, int M, int N, int K, void* ws, size_t ws_bytes,
                      pgv_stream_t stream) {
    PGV_CHECK_ARG(h && x && wr && y && M > 0 && N > 0 && K > 0, "pgv_linear_cl_fwd: bad argument");
    return conv_cl_gemm(h, "pgv_linear_cl_fwd", x, wr, bias, y, M, 1, 1, K, 1, 1, 1, 0, 1, 1, N, CL_EPI_ROWS, 0, 0, 0, -1.0f, 0,
                        static_cast<size_t>(M) * N, nullptr, nullptr, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int pgv_linear_cl_dgrad(pgv_handle* h, const float* dy, const float* wt, float* dx, int M, int N, int K, void* ws, size_t ws_bytes,
                        pgv_stream_t stream) {
    PGV_CHECK_ARG(h && dy && wt && dx && M > 0 && N > 0 && K > 0, "pgv_linear_cl_dgrad: bad argument");
    return conv_cl_gemm(h, "pgv_linear_cl_dgrad", dy, wt, nullptr, dx, M, 1, 1, N, 1, 1, 1, 0, 1, 1, K, CL_EPI_ROWS, 0, 0, 0, -1.0f, 0,
                        static_cast<size_t>(M) * K, nullptr, nullptr, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int pgv_linear_cl_wgrad(pgv_handle* h, const float* dy, const float* x, float* dw, int lddw, int M, int N, int K, int k_valid, void* ws,
                        size_t ws_bytes, pgv_stream_t stream) {
    PGV_CHECK_ARG(k_valid > 0 && k_valid <= K && lddw >= k_valid, "pgv_linear_cl_wgrad: bad leading dimension");
    return conv_cl_wgrad_impl(h, "pgv_linear_cl_wgrad", x, dy, dw, lddw, k_valid, 0, M, 1, 1, K, N, 1, 1, 1, 0, 1, 1, ws, ws_bytes,
                              static_cast<cudaStream_t>(stream));
}

int pgv_conv_cl_workspace_bytes(void) { return static_cast<int>(CL_WS_HEADER); }

int pgv_debug_set_conv_a_mode(int mode) { g_conv_a_mode = mode; return 0; }

int pgv_round_copy(const float* src, int lds, float* dst, int ldd, int rows, int cols, pgv_stream_t stream) {
    PGV_CHECK_ARG(src && dst && rows > 0 && cols > 0 && lds >= cols && ldd >= cols, "pgv_round_copy: bad argument");
    const long long total = static_cast<long long>(rows) * ldd;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
    round_copy_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, lds, dst, ldd, rows, cols);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_conv_cl_unpack_dw(const float* dwcl, float* dw, int Cout, int Cin, int KH, int KW, pgv_stream_t stream) {
    PGV_CHECK_ARG(dwcl && dw && Cout > 0 && Cin > 0 && KH > 0 && KW > 0, "pgv_conv_cl_unpack_dw: bad argument");
    const long long total = static_cast<long long>(Cout) * Cin * KH * KW;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 8));
    cl_unpack_dw_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dwcl, dw, Cout, Cin, KH * KW);
    PGV_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
