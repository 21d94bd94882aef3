// Column-slice GEMM for the flow conditioners (nflows ResidualNet: Linear / BatchNorm1d / ReLU / Dropout chains on
// [batch, 300..610] activations; reference call sites VAE.py:118-125, flows.py:42-90, regression.py:142-148).
//
// These layers are far too small for a tensor-core tile pipeline (160 x 300 x 300) and are latency-bound as separate
// GEMM + BatchNorm launches.  Here a THREAD-BLOCK CLUSTER owns 16 output columns for all rows of the batch: CTA r of the
// cluster computes rows [32 r, 32 r + 32), the per-column batch statistics a BatchNorm1d needs are reduced across the
// cluster through distributed shared memory, and the normalisation is fused into the GEMM:
//   EPI_PLAIN   out = act(A * op(B) + bias + add_pre)                                   (no cluster needed)
//   EPI_BN_FWD  y = A * B^T + bias + add_pre (stored);  out = mask * relu(gamma * (y - mean) * rstd + beta), batch
//               statistics over the M rows, running-stat update            = Linear -> BatchNorm1d -> ReLU -> Dropout
//   EPI_BN_BWD  dt = A * B;  backward of that BatchNorm/ReLU/Dropout w.r.t. its input bn_x, + add_post (residual path);
//               also dgamma / dbeta                                          = Linear data-gradient -> BatchNorm1d backward
// Exact fp32 FMAs (the flows carry log-determinants; they stay off TF32).  Operands are staged with cp.async through a
// 4-stage shared-memory ring (27 KB, so a CTA fits next to a resident 194 KB convolution CTA of the concurrent decoder
// branch); each thread owns 1 row x 2 columns and reads shared memory with 128-bit loads along k.
#include <cooperative_groups.h>

#include <algorithm>
#include <string.h>

#include "pgv_common.cuh"
#include "pgv_tc.cuh"

namespace cg = cooperative_groups;

namespace pgv {

// 10 stages = the whole reduction of a 300-wide layer in flight at once: the K loop of such a tile is one L2 round trip plus the FMAs
// (measured with 4 stages: 0.6 us per 32-wide chunk, i.e. one exposed L2 latency per chunk, 6.1 us per tile).
constexpr int CS_COLS = 16, CS_ROWS = 32, CS_BK = 32, CS_LD = 36, CS_MAXM = 256, CS_THREADS = 256, CS_GROUP = 2, CS_NG = 5, CS_STAGES = CS_NG * CS_GROUP;
constexpr int CS_STAGE_FLOATS = (CS_ROWS + CS_COLS) * CS_LD;
constexpr int CS_SMEM = CS_STAGES * CS_STAGE_FLOATS * 4;
enum { EPI_PLAIN = 0, EPI_BN_FWD = 1, EPI_BN_BWD = 2 };

struct CsParams {
    const float* a; int lda;            // [M, Kd]
    const float* b; int ldb;            // TB = 0: [N, Kd] (Linear weight, forward)   TB = 1: [Kd, N] (Linear weight, data gradient)
    int M, N, Kd;
    const float* bias;                  // [N] or NULL
    const float* add_pre;               // [M, N] or NULL: added to the product before the epilogue
    float* out_pre;                     // EPI_BN_FWD: y (pre-normalisation), [M, N]
    float* out;                         // [M, N]
    const float* gamma; const float* beta; const float* mask;          // BatchNorm affine, Dropout keep-mask [M, N] or NULL
    float* save_mean; float* save_rstd; float* running_mean; float* running_var; float momentum, eps;     // EPI_BN_FWD
    const float* bn_x; const float* mean; const float* rstd; const float* add_post; float* dgamma; float* dbeta;   // EPI_BN_BWD
    int relu;                           // EPI_PLAIN: ReLU on the result
    int a_vec, b_vec;                   // operand rows are 16-byte aligned and Kd % 4 == 0: 16-byte cp.async
};

struct CsShared {
    double wred[8][CS_COLS][2];
    double gather[2][8][CS_COLS][2];    // [reduction parity][cluster rank]: every CTA of the cluster PUSHES its partial totals here
    double tot[CS_COLS][2];
    float stat[CS_COLS][2];
};

// Column totals over ALL rows of the batch of two per-thread quantities (thread = 1 row x columns tx, tx + 8):
// warp shuffle over the 4 rows of a warp, shared memory over the 8 warps, distributed shared memory over the cluster.
// On return tot[c][0..1] holds the totals of column c in every CTA of the cluster (summed in rank order everywhere: bit-identical).
// Each CTA stores its partial into every peer's `gather` slot (remote stores do not wait for a round trip, remote loads do) and ONE
// cluster barrier publishes them; the slots alternate with `phase` (the count of reductions this CTA has done, the same in every CTA
// of a cluster), so a CTA that races ahead to its next reduction writes the other buffer and needs no second barrier: to reach the
// reduction after that it has to pass a barrier its peers only arrive at once they have read the first one.
__device__ __forceinline__ void cs_cluster_col_reduce(double (&s)[2], double (&q)[2], CsShared& sh, int cluster_size, unsigned& phase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tx = threadIdx.x & 7;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8);  q[j] += __shfl_xor_sync(0xffffffffu, q[j], 8);
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16); q[j] += __shfl_xor_sync(0xffffffffu, q[j], 16);
        if (lane < 8) { sh.wred[warp][tx + 8 * j][0] = s[j]; sh.wred[warp][tx + 8 * j][1] = q[j]; }
    }
    __syncthreads();
    cg::cluster_group cluster = cg::this_cluster();
    double (*slot)[CS_COLS][2] = sh.gather[phase & 1u];
    if (threadIdx.x < 2 * CS_COLS) {
        const int c = threadIdx.x >> 1, w = threadIdx.x & 1;
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += sh.wred[g][c][w];
        double* mine = &slot[cluster.block_rank()][c][w];
        for (int r = 0; r < cluster_size; ++r) *cluster.map_shared_rank(mine, r) = t;
    }
    cluster.sync();                                       // (release / acquire) every CTA's partial has landed in every CTA
    if (threadIdx.x < 2 * CS_COLS) {
        const int c = threadIdx.x >> 1, w = threadIdx.x & 1;
        double t = 0.0;
        for (int r = 0; r < cluster_size; ++r) t += slot[r][c][w];
        sh.tot[c][w] = t;
    }
    __syncthreads();
    ++phase;
}

// One 32-row x 16-column tile: column slice `slice`, row block `rblock` of `cluster_size`.  PDL: the stand-alone kernels prefetch their
// weights before griddepcontrol.wait; inside the flow program kernel (pdl = false) the grid barrier has already ordered everything.
template <int TB, int EPI>
__device__ __forceinline__ void colslice_tile(const CsParams& p, int slice, int rblock, int cluster_size, uint8_t* cs_smem, CsShared& sh,
                                              bool pdl, unsigned& phase, unsigned long long* dbg = nullptr) {
    if (dbg != nullptr) dbg[0] = global_timer_ns();
    double (&tot)[CS_COLS][2] = sh.tot;
    float (&stat)[CS_COLS][2] = sh.stat;
    const int t = threadIdx.x, tx = t & 7, ty = t >> 3;           // row ty of this CTA's 32; columns tx, tx + 8
    const int n0 = slice * CS_COLS, m0 = rblock * CS_ROWS;
    float* const stage0 = reinterpret_cast<float*>(cs_smem);
    const uint32_t stage0_u32 = smem_u32(stage0);
    const int n_chunks = (p.Kd + CS_BK - 1) / CS_BK;

    // Per-thread copy plan, computed once (everything but k0 is loop invariant; recomputing it per chunk made the copy issue as long
    // as the arithmetic).  A: one 16-byte copy (row t / 8, floats 4 (t % 8) ..) or four 4-byte copies (rows t / 32 + 8 e, float t % 32).
    // B, [N, Kd]: the same shapes over 16 columns;  B, [Kd, N]: two 4-byte copies (k = t / 16 + 16 e, column t % 16), transposed on the way in.
    const int a_kk = p.a_vec ? 4 * (t & 7) : (t & 31), a_row = p.a_vec ? (t >> 3) : (t >> 5);
    const float* const a_src = p.a + static_cast<size_t>(m0 + a_row) * p.lda + a_kk;
    const uint32_t a_dst = (a_row * CS_LD + a_kk) * 4;
    uint32_t a_ok = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) a_ok |= (m0 + a_row + 8 * e < p.M ? 1u : 0u) << e;
    const size_t a_step = static_cast<size_t>(8) * p.lda;
    const bool b_v = TB == 0 && p.b_vec;
    const int b_kk = TB == 1 ? (t >> 4) : (b_v ? 4 * (t & 7) : (t & 31)), b_c = TB == 1 ? (t & 15) : (b_v ? (t >> 3) : (t >> 5));
    const float* const b_src = TB == 1 ? p.b + static_cast<size_t>(b_kk) * p.ldb + n0 + b_c : p.b + static_cast<size_t>(n0 + b_c) * p.ldb + b_kk;
    const uint32_t b_dst = (b_c * CS_LD + b_kk) * 4;
    const uint32_t b_ok = TB == 1 ? (n0 + b_c < p.N ? 3u : 0u)
                                  : (b_v ? (t < CS_COLS * 8 && n0 + b_c < p.N ? 1u : 0u) : ((n0 + b_c < p.N ? 1u : 0u) | (n0 + b_c + 8 < p.N ? 2u : 0u)));
    const size_t b_step = TB == 1 ? static_cast<size_t>(16) * p.ldb : static_cast<size_t>(8) * p.ldb;
    const size_t b_kstride = TB == 1 ? static_cast<size_t>(p.ldb) : 1;

    // parts: bit 0 = the A (activation) tile, bit 1 = the B (weight) tile of chunk ch; `commit` closes the chunk's cp.async group
    auto issue = [&](int ch, int parts, bool commit) {
        if (ch < n_chunks) {
            const int k0 = ch * CS_BK;
            const uint32_t sA = stage0_u32 + (ch % CS_STAGES) * CS_STAGE_FLOATS * 4, sB = sA + CS_ROWS * CS_LD * 4;
            if (parts & 1) {
                const bool kin = k0 + a_kk < p.Kd;
                if (p.a_vec) {
                    const bool ok = kin && (a_ok & 1u);
                    cp_async16_cg(sA + a_dst, ok ? a_src + k0 : p.a, ok ? 16u : 0u);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool ok = kin && ((a_ok >> e) & 1u);
                        cp_async4(sA + a_dst + e * 8 * CS_LD * 4, ok ? a_src + e * a_step + k0 : p.a, ok ? 4u : 0u);
                    }
                }
            }
            if (parts & 2) {
                if (TB == 1) {                              // B[k0 + kk][n0 + c], contiguous along c
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const bool ok = (b_ok & 1u) && k0 + b_kk + 16 * e < p.Kd;
                        cp_async4(sB + b_dst + e * 16 * 4, ok ? b_src + e * b_step + k0 * b_kstride : p.b, ok ? 4u : 0u);
                    }
                } else if (b_v) {                           // B[n0 + c][k0 + kk], contiguous along k
                    if (t < CS_COLS * 8) {
                        const bool ok = (b_ok & 1u) && k0 + b_kk < p.Kd;
                        cp_async16_cg(sB + b_dst, ok ? b_src + k0 : p.b, ok ? 16u : 0u);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const bool ok = ((b_ok >> e) & 1u) && k0 + b_kk < p.Kd;
                        cp_async4(sB + b_dst + e * 8 * CS_LD * 4, ok ? b_src + e * b_step + k0 : p.b, ok ? 4u : 0u);
                    }
                }
            }
        }
        if (commit) cp_async_commit();                      // one group per chunk slot, empty past the end
    };

    // Programmatic dependent launch: the weights do not depend on the kernel in front, so their first tiles are fetched while that
    // kernel is still draining; everything else (activations, every store) comes after griddepcontrol.wait.
    float acc[2] = {0.0f, 0.0f};
    // The ring holds CS_NG GROUPS of CS_GROUP chunks; one cp.async group and one pair of CTA barriers per chunk group (a barrier per
    // chunk put every warp's shared-memory loads and then every warp's FMAs in lock step).
    if (pdl) {
        griddep_launch_dependents();
#pragma unroll
        for (int s = 0; s < CS_STAGES; ++s) issue(s, 2, false);
        griddep_wait();
#pragma unroll
        for (int s = 0; s < CS_STAGES; ++s) issue(s, 1, s % CS_GROUP == CS_GROUP - 1);
    } else {
#pragma unroll
        for (int s = 0; s < CS_STAGES; ++s) issue(s, 3, s % CS_GROUP == CS_GROUP - 1);
    }
    // The epilogue's operands are independent of the product: fetch them now so their latency hides behind the K loop.
    const int row = m0 + ty;
    const bool row_ok = row < p.M;
    const int col[2] = {n0 + tx, n0 + tx + 8};
    const bool ok[2] = {row_ok && col[0] < p.N, row_ok && col[1] < p.N};
    float e_bias[2] = {0.0f, 0.0f}, e_add[2] = {0.0f, 0.0f}, e_gamma[2] = {0.0f, 0.0f}, e_beta[2] = {0.0f, 0.0f}, e_mask[2] = {1.0f, 1.0f};
    float e_x[2] = {0.0f, 0.0f}, e_mean[2] = {0.0f, 0.0f}, e_rstd[2] = {0.0f, 0.0f}, e_post[2] = {0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (ok[j]) {
            const size_t o = static_cast<size_t>(row) * p.N + col[j];
            if (p.bias != nullptr) e_bias[j] = __ldg(p.bias + col[j]);
            if (p.add_pre != nullptr) e_add[j] = p.add_pre[o];
            if (EPI != EPI_PLAIN) {
                e_gamma[j] = __ldg(p.gamma + col[j]); e_beta[j] = __ldg(p.beta + col[j]);
                if (p.mask != nullptr) e_mask[j] = p.mask[o];
            }
            if (EPI == EPI_BN_BWD) {
                e_x[j] = p.bn_x[o]; e_mean[j] = p.mean[col[j]]; e_rstd[j] = p.rstd[col[j]];
                if (p.add_post != nullptr) e_post[j] = p.add_post[o];
            }
        }
    // The K loop is split over the 8 warps: warp w multiplies the w-th group of 4 k values of every 32-wide chunk for the WHOLE 32 x 16
    // tile, each lane rows r4 + 8 i x columns c4 + 4 j (interleaved so that the 36-float row pitch spreads a quarter warp's float4 loads
    // over distinct banks; 8 LDS.128 per 64 FMAs; with 1 row x 2 columns per thread the loop was bound by shared-memory
    // wavefronts: 0.6 us per chunk), and the eight partial tiles are added through shared memory in warp order afterwards.
    const int lane_ = t & 31, warp_ = t >> 5, r4 = lane_ >> 2, c4 = lane_ & 3;
    float part16[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part16[i][j] = 0.0f;
    const int n_groups = (n_chunks + CS_GROUP - 1) / CS_GROUP;
    for (int g = 0; g < n_groups; ++g) {
        cp_async_wait<CS_NG - 1>();                         // group g has landed (this thread's copies) ...
        __syncthreads();                                    // ... and everybody's
        if (dbg != nullptr && g == 0) dbg[3] = global_timer_ns();
        const float* const ring = stage0 + (g % CS_NG) * CS_GROUP * CS_STAGE_FLOATS + 4 * warp_;
#pragma unroll
        for (int c = 0; c < CS_GROUP; ++c) {
            if (g * CS_GROUP + c < n_chunks) {
                const float* sA = ring + c * CS_STAGE_FLOATS;
                const float* sB = sA + CS_ROWS * CS_LD;
                float4 av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(sA + (r4 + 8 * i) * CS_LD);
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(sB + (c4 + 4 * j) * CS_LD);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        part16[i][j] = fmaf(av[i].x, bv[j].x, fmaf(av[i].y, bv[j].y, fmaf(av[i].z, bv[j].z, fmaf(av[i].w, bv[j].w, part16[i][j]))));
            }
        }
        if (g + CS_NG < n_groups) {                         // (uniform) refill this part of the ring once everybody has left it
            __syncthreads();
#pragma unroll
            for (int c = 0; c < CS_GROUP; ++c) issue((g + CS_NG) * CS_GROUP + c, 3, false);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();
    __syncthreads();                                        // every warp is done with the ring: reuse it for the partial tiles
    {
        float* red = stage0;                                // [warp][32 rows][16 columns]
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(warp_ * CS_ROWS + r4 + 8 * i) * CS_COLS + c4 + 4 * j] = part16[i][j];
        __syncthreads();
#pragma unroll
        for (int w8 = 0; w8 < CS_THREADS / 32; ++w8) {
            acc[0] += red[(w8 * CS_ROWS + ty) * CS_COLS + tx];
            acc[1] += red[(w8 * CS_ROWS + ty) * CS_COLS + tx + 8];
        }
        __syncthreads();                                    // (the next tile of a program refills the ring)
    }
    if (dbg != nullptr) dbg[1] = global_timer_ns();

    // ---------------------------------------------------------------- epilogue (operands were fetched before the K loop)
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (ok[j]) acc[j] = (acc[j] + e_bias[j]) + e_add[j];
    if (EPI == EPI_PLAIN) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (ok[j]) p.out[static_cast<size_t>(row) * p.N + col[j]] = p.relu ? fmaxf(acc[j], 0.0f) : acc[j];
        return;
    }
    if (EPI == EPI_BN_FWD) {
        double s[2] = {0.0, 0.0}, q[2] = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (ok[j]) {
                const double v = acc[j];
                s[j] = v; q[j] = v * v;
                if (p.out_pre != nullptr) p.out_pre[static_cast<size_t>(row) * p.N + col[j]] = acc[j];
            }
        cs_cluster_col_reduce(s, q, sh, cluster_size, phase);
        if (dbg != nullptr) dbg[2] = global_timer_ns();
        if (t < CS_COLS) {
            const int c = n0 + t;
            const double mean = tot[t][0] / p.M;
            double var = tot[t][1] / p.M - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps))), fm = static_cast<float>(mean);
            stat[t][0] = fm; stat[t][1] = rstd;
            if (c < p.N && rblock == 0) {
                p.save_mean[c] = fm; p.save_rstd[c] = rstd;
                if (p.running_mean != nullptr) {
                    const double unbiased = p.M > 1 ? var * p.M / (p.M - 1.0) : var;
                    p.running_mean[c] = (1.0f - p.momentum) * p.running_mean[c] + p.momentum * fm;
                    p.running_var[c] = (1.0f - p.momentum) * p.running_var[c] + p.momentum * static_cast<float>(unbiased);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (ok[j]) {
                const float g = e_gamma[j] * stat[tx + 8 * j][1], sh_ = e_beta[j] - stat[tx + 8 * j][0] * g;
                p.out[static_cast<size_t>(row) * p.N + col[j]] = fmaxf(fmaf(acc[j], g, sh_), 0.0f) * e_mask[j];
            }
        return;
    }
    // EPI_BN_BWD: acc = gradient w.r.t. the output of mask * relu(BN(bn_x))
    float xh[2] = {0.0f, 0.0f};
    double s[2] = {0.0, 0.0}, q[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (ok[j]) {
            const float h = (e_x[j] - e_mean[j]) * e_rstd[j];
            float d = acc[j] * e_mask[j];
            if (!(fmaf(e_gamma[j], h, e_beta[j]) > 0.0f)) d = 0.0f;
            acc[j] = d; xh[j] = h;
            s[j] = d; q[j] = static_cast<double>(d) * h;
        }
    cs_cluster_col_reduce(s, q, sh, cluster_size, phase);
    if (t < CS_COLS && n0 + t < p.N && rblock == 0) {
        p.dbeta[n0 + t] = static_cast<float>(tot[t][0]);
        p.dgamma[n0 + t] = static_cast<float>(tot[t][1]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (ok[j]) {
            const float gr = e_gamma[j] * e_rstd[j];
            const float mean_d = static_cast<float>(tot[tx + 8 * j][0] / p.M), mean_dx = static_cast<float>(tot[tx + 8 * j][1] / p.M);
            p.out[static_cast<size_t>(row) * p.N + col[j]] = gr * (acc[j] - mean_d - xh[j] * mean_dx) + e_post[j];
        }
}

template <int TB, int EPI>
__global__ void __launch_bounds__(CS_THREADS) colslice_gemm_kernel(const CsParams p) {
    extern __shared__ __align__(16) uint8_t cs_smem_k[];
    __shared__ CsShared sh;
    unsigned phase = 0;
    colslice_tile<TB, EPI>(p, blockIdx.x, blockIdx.y, gridDim.y, cs_smem_k, sh, true, phase);
}

template <int TB, int EPI>
static int cs_launch(CsParams& p, cudaStream_t stream) {
    p.a_vec = (p.Kd % 4 == 0 && p.lda % 4 == 0 && (reinterpret_cast<uintptr_t>(p.a) & 15) == 0) ? 1 : 0;
    p.b_vec = (TB == 0 && p.Kd % 4 == 0 && p.ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(p.b) & 15) == 0) ? 1 : 0;
    const int row_ctas = ceil_div(p.M, CS_ROWS);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ceil_div(p.N, CS_COLS), row_ctas, 1);
    cfg.blockDim = dim3(CS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = CS_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = (EPI == EPI_PLAIN) ? 1 : row_ctas;       // the row CTAs of one column slice exchange BatchNorm partial sums
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;    // see the prologue of the kernel
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 2 : 1;
    static bool configured = false;
    if (!configured) {
        PGV_CUDA(cudaFuncSetAttribute(colslice_gemm_kernel<TB, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
        configured = true;
    }
    PGV_CUDA(cudaLaunchKernelEx(&cfg, colslice_gemm_kernel<TB, EPI>, p));
    return 0;
}

// ------------------------------------------------------------------------------------------------ flow program kernel
// A RealNVP flow is a chain of ~50 (forward) / ~90 (backward) of the small kernels above and in pgv_flow.cu, each 5-10 us of mostly
// launch / drain / cold-start latency: on the training step's critical path the two flows took 3.6 of 7.4 ms.  The program kernel runs
// such a chain as ONE persistent launch: the host records the calls into a table (passed by value as a kernel parameter, up to 32 KB),
// every CTA walks the table and executes its share of each op - the same tile code as the stand-alone kernels - with a grid-wide
// barrier between dependent ops (release: bar.sync + fence + atomic; acquire: spin on ld.acquire + fence, which also invalidates L1).
// Grid = G clusters of `row_ctas` CTAs (the cluster that reduces BatchNorm statistics through DSMEM); all CTAs are co-resident
// (<= 1 per SM; 31 KB of shared memory each, so they fit beside a 188 KB convolution CTA of the concurrent decoder branch).
enum { MOP_GATHER = 0, MOP_CS_FWD = 1, MOP_CS_BN_FWD = 2, MOP_CS_DGRAD = 3, MOP_CS_BN_BWD = 4, MOP_COUPLING_FWD = 5, MOP_COUPLING_BWD = 6,
       MOP_SCATTER_ADD = 7, MOP_WGRAD = 8 };
struct MegaOp {
    int kind, barrier_after;
    CsParams cs;          // column-slice kinds: as for the stand-alone kernels; other kinds reuse the fields (see pgv_flow_program)
};
constexpr int MEGA_MAX_OPS = 150;
struct MegaProgram {
    int n_ops, row_ctas;
    unsigned* counter;    // grid barrier: zero on entry
    unsigned long long* trace;   // debug (pgv_debug_set_flow_trace): per op, globaltimer of CTA 0 at start / after the body / after the barrier
    MegaOp ops[MEGA_MAX_OPS];
};
static_assert(CS_SMEM >= 2 * CS_MAXM * 32 * 4, "the weight-gradient tile keeps two [CS_MAXM][32] panels in the ring");
static_assert(sizeof(MegaProgram) <= 32000, "the program must fit the kernel parameter space");

__device__ unsigned long long* g_mega_diag = nullptr;      // pinned host memory (pgv_debug_set_flow_diag): what a timed-out barrier saw

__device__ __forceinline__ void mega_grid_barrier(unsigned* counter, unsigned& target, unsigned n_ctas, int op_index) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += n_ctas;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v, spins = 0;
        uint64_t t0 = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v < target && (++spins & 4095u) == 0) {       // never wedge the device: a barrier that has not opened after 4 s traps
                const uint64_t now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) {
                    if (g_mega_diag != nullptr) {
                        g_mega_diag[0] = 0xdeadull; g_mega_diag[1] = blockIdx.z * gridDim.y + blockIdx.y; g_mega_diag[2] = op_index;
                        g_mega_diag[3] = v; g_mega_diag[4] = target; g_mega_diag[5] = n_ctas;
                        __threadfence_system();
                    }
                    __trap();
                }
            }
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float mega_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// One 32 x 32 tile of dW[N, Kd] = dy^T x (and the bias gradient, the column sums of dy, on the tiles of the first k column).
// Both operand panels ([M, 32] each, M <= CS_MAXM) are fetched in ONE round of cp.async (the batch loop used to pay a global-memory
// round trip per 32 rows); the batch rows are then split over the 8 warps, each lane a 4 (n) x 8 (k) register block (3 LDS.128 per 32
// FMAs), and the eight partial tiles are added through shared memory in warp order (deterministic).
__device__ __forceinline__ void mega_wgrad_tile(const CsParams& p, int tile_k, int tile_n, float* smem) {
    const float* dy = p.a; const float* x = p.b; float* dw = p.out; float* db = p.out_pre;
    const int M = p.M, N = p.N, K = p.Kd, t = threadIdx.x, lane = t & 31, warp = t >> 5, n0 = tile_n * 32, k0 = tile_k * 32;
    float* sa = smem;                                         // dy panel [M][32]; afterwards the partial tiles [8][32][32]
    float* sb = smem + CS_MAXM * 32;                          // x panel  [M][32]; afterwards the partial column sums [8][32]
    const uint32_t sa_u = smem_u32(sa), sb_u = smem_u32(sb);
    if (N % 4 == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
        for (int id = t; id < M * 8; id += CS_THREADS) {
            const int r = id >> 3, c = 4 * (id & 7);
            const bool ok = n0 + c < N;
            cp_async16_cg(sa_u + (r * 32 + c) * 4, ok ? dy + static_cast<size_t>(r) * N + n0 + c : dy, ok ? 16u : 0u);
        }
    } else {
        for (int id = t; id < M * 32; id += CS_THREADS) {
            const int r = id >> 5, c = id & 31;
            const bool ok = n0 + c < N;
            cp_async4(sa_u + (r * 32 + c) * 4, ok ? dy + static_cast<size_t>(r) * N + n0 + c : dy, ok ? 4u : 0u);
        }
    }
    if (K % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        for (int id = t; id < M * 8; id += CS_THREADS) {
            const int r = id >> 3, c = 4 * (id & 7);
            const bool ok = k0 + c < K;
            cp_async16_cg(sb_u + (r * 32 + c) * 4, ok ? x + static_cast<size_t>(r) * K + k0 + c : x, ok ? 16u : 0u);
        }
    } else {
        for (int id = t; id < M * 32; id += CS_THREADS) {
            const int r = id >> 5, c = id & 31;
            const bool ok = k0 + c < K;
            cp_async4(sb_u + (r * 32 + c) * 4, ok ? x + static_cast<size_t>(r) * K + k0 + c : x, ok ? 4u : 0u);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int n4 = lane >> 2, k8 = lane & 3;
    float acc[4][8], cs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
    for (int m = warp; m < M; m += CS_THREADS / 32) {
        const float4 a4 = *reinterpret_cast<const float4*>(sa + m * 32 + 4 * n4);
        const float4 b0 = *reinterpret_cast<const float4*>(sb + m * 32 + 8 * k8), b1 = *reinterpret_cast<const float4*>(sb + m * 32 + 8 * k8 + 4);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cs[i] += av[i];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    __syncthreads();                                          // every warp is done with the panels
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float* dst = sa + warp * 1024 + (4 * n4 + i) * 32 + 8 * k8;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        if (k8 == 0) sb[warp * 32 + 4 * n4 + i] = cs[i];
    }
    __syncthreads();
    float4 tot = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int w = 0; w < CS_THREADS / 32; ++w) {
        const float4 v = *reinterpret_cast<const float4*>(sa + w * 1024 + 4 * t);
        tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
    }
    const int n = n0 + (t >> 3), k = k0 + 4 * (t & 7);
    if (n < N) {
        float* o = dw + static_cast<size_t>(n) * K + k;
        if (k < K) o[0] = tot.x;
        if (k + 1 < K) o[1] = tot.y;
        if (k + 2 < K) o[2] = tot.z;
        if (k + 3 < K) o[3] = tot.w;
    }
    if (db != nullptr && tile_k == 0 && t < 32 && n0 + t < N) {
        float c = 0.0f;
#pragma unroll
        for (int w = 0; w < CS_THREADS / 32; ++w) c += sb[w * 32 + t];
        db[n0 + t] = c;
    }
}

__global__ void __launch_bounds__(CS_THREADS) flow_program_kernel(const __grid_constant__ MegaProgram prog) {
    extern __shared__ __align__(16) uint8_t cs_smem_m[];
    __shared__ CsShared sh;
    __shared__ float red[8];
    __shared__ int wg_tile[2];
    const int row_ctas = prog.row_ctas, rblock = blockIdx.y, cl = blockIdx.z, n_clusters = gridDim.z;
    const unsigned n_ctas = gridDim.y * gridDim.z;
    const int cta = cl * row_ctas + rblock, t = threadIdx.x;
    unsigned target = 0;
    const bool tracing = prog.trace != nullptr && cta == 0 && t == 0;
    unsigned bn_phase = 0;                                  // BatchNorm reductions done so far (the same in every CTA of a cluster)
    for (int oi = 0; oi < prog.n_ops; ++oi) {
        const MegaOp& op = prog.ops[oi];
        const CsParams& p = op.cs;
        if (tracing) { prog.trace[4 * oi] = global_timer_ns(); prog.trace[4 * oi + 3] = static_cast<unsigned long long>(op.kind); }
        switch (op.kind) {
        case MOP_CS_FWD: case MOP_CS_BN_FWD: case MOP_CS_DGRAD: case MOP_CS_BN_BWD: {
            const int slices = (p.N + CS_COLS - 1) / CS_COLS;
            for (int sl = cl; sl < slices; sl += n_clusters) {
                __syncthreads();                               // the previous tile's readers are done with the ring and the statistics
                unsigned long long* dbg = (tracing && sl == cl) ? prog.trace + 4 * MEGA_MAX_OPS + 4 * oi : nullptr;
                if (op.kind == MOP_CS_FWD) colslice_tile<0, EPI_PLAIN>(p, sl, rblock, row_ctas, cs_smem_m, sh, false, bn_phase, dbg);
                else if (op.kind == MOP_CS_BN_FWD) colslice_tile<0, EPI_BN_FWD>(p, sl, rblock, row_ctas, cs_smem_m, sh, false, bn_phase, dbg);
                else if (op.kind == MOP_CS_DGRAD) colslice_tile<1, EPI_PLAIN>(p, sl, rblock, row_ctas, cs_smem_m, sh, false, bn_phase);
                else colslice_tile<1, EPI_BN_BWD>(p, sl, rblock, row_ctas, cs_smem_m, sh, false, bn_phase);
            }
            break;
        }
        case MOP_GATHER: {                                     // out[b, j] = x[b, idx[j]]: a = x, b = idx, M = B, N = D, Kd = n
            const int* idx = reinterpret_cast<const int*>(p.b);
            const size_t total = static_cast<size_t>(p.M) * p.Kd;
            for (size_t i = static_cast<size_t>(cta) * CS_THREADS + t; i < total; i += static_cast<size_t>(n_ctas) * CS_THREADS)
                p.out[i] = p.a[(i / p.Kd) * p.N + idx[i % p.Kd]];
            break;
        }
        case MOP_SCATTER_ADD: {                                // dst[b, idx[j]] += src[b, j]: out = dst, b = idx, a = src
            const int* idx = reinterpret_cast<const int*>(p.b);
            const size_t total = static_cast<size_t>(p.M) * p.Kd;
            for (size_t i = static_cast<size_t>(cta) * CS_THREADS + t; i < total; i += static_cast<size_t>(n_ctas) * CS_THREADS)
                p.out[(i / p.Kd) * p.N + idx[i % p.Kd]] += p.a[i];
            break;
        }
        case MOP_COUPLING_FWD: {       // a = x, b = params, bias = id_idx, add_pre = tr_idx, out = y, out_pre = ld_out, gamma = ld_in; Kd = n_id, lda = n_t
            const int* id_idx = reinterpret_cast<const int*>(p.bias);
            const int* tr_idx = reinterpret_cast<const int*>(p.add_pre);
            const int D = p.N, n_id = p.Kd, n_t = p.lda, inverse = p.relu;
            for (int row = cta; row < p.M; row += static_cast<int>(n_ctas)) {
                const float* xr = p.a + static_cast<size_t>(row) * D;
                float* yr = p.out + static_cast<size_t>(row) * D;
                const float* pr = p.b + static_cast<size_t>(row) * 2 * n_t;
                for (int j = t; j < n_id; j += CS_THREADS) yr[id_idx[j]] = xr[id_idx[j]];
                float ld = 0.0f;
                for (int j = t; j < n_t; j += CS_THREADS) {
                    const float sc = mega_sigmoid(pr[n_t + j] + 2.0f) + 1e-3f, sf = pr[j];
                    const int c = tr_idx[j];
                    yr[c] = inverse ? (xr[c] - sf) / sc : fmaf(xr[c], sc, sf);
                    ld += logf(sc);
                }
                for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
                __syncthreads();
                if ((t & 31) == 0) red[t >> 5] = ld;
                __syncthreads();
                if (t == 0) {
                    float tot = 0.0f;
                    for (int w = 0; w < CS_THREADS / 32; ++w) tot += red[w];
                    p.out_pre[row] = (p.gamma != nullptr ? p.gamma[row] : 0.0f) + (inverse ? -tot : tot);
                }
            }
            break;
        }
        case MOP_COUPLING_BWD: {       // a = dy, b = dld, bias = x, add_pre = params, gamma = id_idx, beta = tr_idx, out = dx, out_pre = dparams
            const int* id_idx = reinterpret_cast<const int*>(p.gamma);
            const int* tr_idx = reinterpret_cast<const int*>(p.beta);
            const int D = p.N, n_id = p.Kd, n_t = p.lda;
            for (int row = cta; row < p.M; row += static_cast<int>(n_ctas)) {
                const size_t ro = static_cast<size_t>(row) * D;
                const float* pr = p.add_pre + static_cast<size_t>(row) * 2 * n_t;
                float* dpr = p.out_pre + static_cast<size_t>(row) * 2 * n_t;
                const float gl = p.b != nullptr ? p.b[row] : 0.0f;
                for (int j = t; j < n_id; j += CS_THREADS) p.out[ro + id_idx[j]] = p.a[ro + id_idx[j]];
                for (int j = t; j < n_t; j += CS_THREADS) {
                    const int c = tr_idx[j];
                    const float sg = mega_sigmoid(pr[n_t + j] + 2.0f), sc = sg + 1e-3f, g = p.a[ro + c];
                    p.out[ro + c] = g * sc;
                    dpr[j] = g;
                    dpr[n_t + j] = (g * p.bias[ro + c] + gl / sc) * sg * (1.0f - sg);
                }
            }
            break;
        }
        case MOP_WGRAD: {              // a = dy [M, N], b = x [M, Kd], out = dw [N, Kd], out_pre = db, relu = index of this op's tile counter
            // Tiles are handed out through an atomic counter: a weight gradient sits BEHIND the data-gradient op of its barrier
            // interval (pgv_flow_program orders them so), so the CTAs that op leaves idle start on the tiles at once and the busy
            // ones join when they are through.  Which CTA computes a tile does not change its value.
            const int tk = (p.Kd + 31) / 32, tn = (p.N + 31) / 32;
            unsigned* const ctr = prog.counter + 1 + p.relu;
            __syncthreads();                                   // the previous op is done with wg_tile
            if (t == 0) wg_tile[0] = static_cast<int>(atomicAdd(ctr, 1u));
            for (int it = 0;; ++it) {
                __syncthreads();                               // wg_tile[it & 1] is set; the previous tile is done with shared memory
                const int tile = wg_tile[it & 1];
                if (tile >= tk * tn) break;
                int next = 0;
                if (t == 0) next = static_cast<int>(atomicAdd(ctr, 1u));       // the next ticket travels while this tile is computed
                mega_wgrad_tile(p, tile % tk, tile / tk, reinterpret_cast<float*>(cs_smem_m));
                if (t == 0) wg_tile[(it + 1) & 1] = next;
            }
            break;
        }
        default: break;
        }
        if (prog.trace != nullptr) __syncthreads();        // (uniform condition) the body time of CTA 0 = its slowest thread
        if (tracing) prog.trace[4 * oi + 1] = global_timer_ns();
        if (op.barrier_after) mega_grid_barrier(prog.counter, target, n_ctas, oi);
        if (tracing) prog.trace[4 * oi + 2] = global_timer_ns();
    }
}

unsigned long long* g_flow_trace = nullptr;
int g_flow_clusters = 0;

}  // namespace pgv

using namespace pgv;

extern "C" {

/* Runs a recorded chain of flow ops as one persistent launch.  `ops` is a HOST array of n_ops records {int kind; int barrier_after;
 * CsParams} (layout as in this file; the Python side builds it with ctypes), `counter` a zero-filled device word (zeroed again by this
 * call), M the batch rows shared by the column-slice ops. */
int pgv_flow_program(pgv_handle* h, const void* ops, int n_ops, int M, unsigned* counter, pgv_stream_t stream_) {
    PGV_CHECK_ARG(h && ops && counter && n_ops > 0 && n_ops <= MEGA_MAX_OPS && M > 0 && M <= CS_MAXM, "pgv_flow_program: bad argument (%d ops)", n_ops);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    static MegaProgram prog;                                  // 30 KB: not on the stack; copied into the launch by cudaLaunchKernelEx
    prog.n_ops = n_ops;
    prog.row_ctas = ceil_div(M, CS_ROWS);
    prog.counter = counter;
    prog.trace = g_flow_trace;
    memcpy(prog.ops, ops, sizeof(MegaOp) * n_ops);
    for (int i = 0; i < n_ops; ++i) {
        CsParams& p = prog.ops[i].cs;
        const int k = prog.ops[i].kind;
        if (k == MOP_CS_FWD || k == MOP_CS_BN_FWD || k == MOP_CS_DGRAD || k == MOP_CS_BN_BWD) {
            const bool tb0 = k == MOP_CS_FWD || k == MOP_CS_BN_FWD;
            p.a_vec = (p.Kd % 4 == 0 && p.lda % 4 == 0 && (reinterpret_cast<uintptr_t>(p.a) & 15) == 0) ? 1 : 0;
            p.b_vec = (tb0 && p.Kd % 4 == 0 && p.ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(p.b) & 15) == 0) ? 1 : 0;
            if (p.M != M) return set_error(-1, "pgv_flow_program: op %d has %d rows, the program %d", i, p.M, M);
        }
    }
    // Inside a barrier interval the ops are independent by construction: put the weight gradients last (see MOP_WGRAD) and give
    // each one a tile counter; counter[0] is the grid barrier's.
    int n_wgrad = 0;
    for (int lo = 0; lo < n_ops;) {
        int hi = lo;
        while (hi < n_ops - 1 && !prog.ops[hi].barrier_after) ++hi;
        std::stable_partition(prog.ops + lo, prog.ops + hi + 1, [](const MegaOp& o) { return o.kind != MOP_WGRAD; });
        for (int i = lo; i <= hi; ++i) {
            prog.ops[i].barrier_after = i == hi ? 1 : 0;
            if (prog.ops[i].kind == MOP_WGRAD) prog.ops[i].cs.relu = n_wgrad++;
        }
        lo = hi + 1;
    }
    prog.ops[n_ops - 1].barrier_after = 0;
    static bool configured = false;
    if (!configured) {
        PGV_CUDA(cudaFuncSetAttribute(flow_program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
        configured = true;
    }
    PGV_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned) * (1 + n_wgrad), stream));
    // co-resident: at most one CTA per SM (pgv_debug_set_flow_clusters overrides the count for sweeps)
    const int clusters = std::max(1, std::min(g_flow_clusters > 0 ? g_flow_clusters : 39, (h->sm_count - 8) / prog.row_ctas));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(1, prog.row_ctas, clusters);
    cfg.blockDim = dim3(CS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = CS_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = prog.row_ctas;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PGV_CUDA(cudaLaunchKernelEx(&cfg, flow_program_kernel, prog));
    return 0;
}

int pgv_debug_set_flow_diag(void* pinned_host) {
    unsigned long long* p = static_cast<unsigned long long*>(pinned_host);
    PGV_CUDA(cudaMemcpyToSymbol(g_mega_diag, &p, sizeof(p)));
    return 0;
}

int pgv_debug_set_flow_clusters(int clusters) { g_flow_clusters = clusters; return 0; }
int pgv_debug_set_flow_trace(void* trace_dev) { g_flow_trace = static_cast<unsigned long long*>(trace_dev); return 0; }

int pgv_flow_program_op_bytes(void) { return static_cast<int>(sizeof(MegaOp)); }
int pgv_flow_program_max_ops(void) { return MEGA_MAX_OPS; }

int pgv_colslice_max_rows(void) { return CS_MAXM; }

int pgv_linear_cs_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y, int M, int N, int K, int relu,
                      pgv_stream_t stream) {
    PGV_CHECK_ARG(x && w && y && M > 0 && M <= 65535 * CS_ROWS && N > 0 && K > 0, "pgv_linear_cs_fwd: bad argument");
    CsParams p;
    memset(&p, 0, sizeof(p));
    p.a = x; p.lda = K; p.b = w; p.ldb = K; p.M = M; p.N = N; p.Kd = K; p.bias = bias; p.add_pre = residual; p.out = y; p.relu = relu;
    return cs_launch<0, EPI_PLAIN>(p, static_cast<cudaStream_t>(stream));
}

int pgv_linear_cs_dgrad(const float* dy, const float* w, float* dx, int M, int N, int K, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && w && dx && M > 0 && M <= 65535 * CS_ROWS && N > 0 && K > 0, "pgv_linear_cs_dgrad: bad argument");
    CsParams p;
    memset(&p, 0, sizeof(p));
    p.a = dy; p.lda = N; p.b = w; p.ldb = K; p.M = M; p.N = K; p.Kd = N; p.out = dx;
    return cs_launch<1, EPI_PLAIN>(p, static_cast<cudaStream_t>(stream));
}

int pgv_linear_bn_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y_pre, float* out,
                      const float* gamma, const float* beta, const float* mask, float* save_mean, float* save_rstd, float* running_mean,
                      float* running_var, float momentum, float eps, int M, int N, int K, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && w && out && gamma && beta && save_mean && save_rstd && M > 0 && M <= CS_MAXM && N > 0 && K > 0,
                  "pgv_linear_bn_fwd: bad argument (M <= %d)", CS_MAXM);
    CsParams p;
    memset(&p, 0, sizeof(p));
    p.a = x; p.lda = K; p.b = w; p.ldb = K; p.M = M; p.N = N; p.Kd = K; p.bias = bias; p.add_pre = residual; p.out_pre = y_pre; p.out = out;
    p.gamma = gamma; p.beta = beta; p.mask = mask; p.save_mean = save_mean; p.save_rstd = save_rstd; p.running_mean = running_mean;
    p.running_var = running_var; p.momentum = momentum; p.eps = eps;
    return cs_launch<0, EPI_BN_FWD>(p, static_cast<cudaStream_t>(stream));
}

int pgv_linear_dgrad_bn_bwd(const float* dy, const float* w, const float* bn_x, const float* gamma, const float* beta, const float* mean,
                            const float* rstd, const float* mask, const float* add_post, float* dx, float* dgamma, float* dbeta, int M,
                            int N, int K, pgv_stream_t stream) {
    PGV_CHECK_ARG(dy && w && bn_x && gamma && beta && mean && rstd && dx && dgamma && dbeta && M > 0 && M <= CS_MAXM && N > 0 && K > 0,
                  "pgv_linear_dgrad_bn_bwd: bad argument (M <= %d)", CS_MAXM);
    CsParams p;
    memset(&p, 0, sizeof(p));
    p.a = dy; p.lda = N; p.b = w; p.ldb = K; p.M = M; p.N = K; p.Kd = N; p.out = dx;
    p.gamma = gamma; p.beta = beta; p.mask = mask; p.bn_x = bn_x; p.mean = mean; p.rstd = rstd; p.add_post = add_post;
    p.dgamma = dgamma; p.dbeta = dbeta;
    return cs_launch<1, EPI_BN_BWD>(p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
