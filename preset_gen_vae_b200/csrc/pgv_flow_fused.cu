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
#include <string.h>

#include "pgv_common.cuh"
#include "pgv_tc.cuh"

namespace cg = cooperative_groups;

namespace pgv {

constexpr int CS_COLS = 16, CS_ROWS = 32, CS_BK = 32, CS_LD = 36, CS_MAXM = 256, CS_THREADS = 256, CS_STAGES = 4;
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

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Column totals over ALL rows of the batch of two per-thread quantities (thread = 1 row x columns tx, tx + 8):
// warp shuffle over the 4 rows of a warp, shared memory over the 8 warps, distributed shared memory over the cluster.
// On return tot[c][0..1] holds the totals of column c in every CTA of the cluster.
__device__ __forceinline__ void cs_cluster_col_reduce(double (&s)[2], double (&q)[2], double (*wred)[CS_COLS][2], double (*part)[2],
                                                      double (*tot)[2], int cluster_size) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tx = threadIdx.x & 7;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8);  q[j] += __shfl_xor_sync(0xffffffffu, q[j], 8);
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16); q[j] += __shfl_xor_sync(0xffffffffu, q[j], 16);
        if (lane < 8) { wred[warp][tx + 8 * j][0] = s[j]; wred[warp][tx + 8 * j][1] = q[j]; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * CS_COLS) {
        const int c = threadIdx.x >> 1, w = threadIdx.x & 1;
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += wred[g][c][w];
        part[c][w] = t;
    }
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                       // every CTA's `part` is complete and visible cluster-wide
    if (threadIdx.x < 2 * CS_COLS) {
        const int c = threadIdx.x >> 1, w = threadIdx.x & 1;
        double t = 0.0;
        for (int r = 0; r < cluster_size; ++r) {
            const double (*remote)[2] = cluster.map_shared_rank(part, r);
            t += remote[c][w];
        }
        tot[c][w] = t;
    }
    cluster.sync();                                       // nobody leaves (or overwrites `part`) while a peer still reads it
}

template <int TB, int EPI>
__global__ void __launch_bounds__(CS_THREADS) colslice_gemm_kernel(const CsParams p) {
    extern __shared__ __align__(16) uint8_t cs_smem[];
    __shared__ double wred[8][CS_COLS][2];
    __shared__ double part[CS_COLS][2], tot[CS_COLS][2];
    __shared__ float stat[CS_COLS][2];
    const int t = threadIdx.x, tx = t & 7, ty = t >> 3;           // row ty of this CTA's 32; columns tx, tx + 8
    const int n0 = blockIdx.x * CS_COLS, m0 = blockIdx.y * CS_ROWS;
    float* const stage0 = reinterpret_cast<float*>(cs_smem);
    const uint32_t stage0_u32 = smem_u32(stage0);
    const int n_chunks = (p.Kd + CS_BK - 1) / CS_BK;

    // parts: bit 0 = the A (activation) tile, bit 1 = the B (weight) tile of chunk ch; `commit` closes the chunk's cp.async group
    auto issue = [&](int ch, int parts, bool commit) {
        if (ch < n_chunks) {
            const int k0 = ch * CS_BK;
            const uint32_t sA = stage0_u32 + (ch % CS_STAGES) * CS_STAGE_FLOATS * 4, sB = sA + CS_ROWS * CS_LD * 4;
            if (!(parts & 1)) {
            } else if (p.a_vec) {                           // 32 rows x 8 float4: one per thread
                const int row = t >> 3, c4 = t & 7, k = k0 + 4 * c4;
                const bool ok = m0 + row < p.M && k < p.Kd;
                cp_async16_cg(sA + (row * CS_LD + 4 * c4) * 4, ok ? p.a + static_cast<size_t>(m0 + row) * p.lda + k : p.a, ok ? 16u : 0u);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int id = t + e * CS_THREADS, row = id >> 5, kk = id & 31, k = k0 + kk;
                    const bool ok = m0 + row < p.M && k < p.Kd;
                    cp_async4(sA + (row * CS_LD + kk) * 4, ok ? p.a + static_cast<size_t>(m0 + row) * p.lda + k : p.a, ok ? 4u : 0u);
                }
            }
            if (!(parts & 2)) {
            } else if (TB == 0) {                           // B[n0 + c][k0 + kk], contiguous along k
                if (p.b_vec) {
                    if (t < CS_COLS * 8) {
                        const int c = t >> 3, c4 = t & 7, k = k0 + 4 * c4;
                        const bool ok = n0 + c < p.N && k < p.Kd;
                        cp_async16_cg(sB + (c * CS_LD + 4 * c4) * 4, ok ? p.b + static_cast<size_t>(n0 + c) * p.ldb + k : p.b, ok ? 16u : 0u);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int id = t + e * CS_THREADS, c = id >> 5, kk = id & 31, k = k0 + kk;
                        const bool ok = n0 + c < p.N && k < p.Kd;
                        cp_async4(sB + (c * CS_LD + kk) * 4, ok ? p.b + static_cast<size_t>(n0 + c) * p.ldb + k : p.b, ok ? 4u : 0u);
                    }
                }
            } else {                                        // B[k0 + kk][n0 + c], contiguous along c: transposed on the way in
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int id = t + e * CS_THREADS, kk = id >> 4, c = id & 15, k = k0 + kk;
                    const bool ok = n0 + c < p.N && k < p.Kd;
                    cp_async4(sB + (c * CS_LD + kk) * 4, ok ? p.b + static_cast<size_t>(k) * p.ldb + n0 + c : p.b, ok ? 4u : 0u);
                }
            }
        }
        if (commit) cp_async_commit();                      // one group per chunk slot, empty past the end
    };

    // Programmatic dependent launch: the weights do not depend on the kernel in front, so their first tiles are fetched while that
    // kernel is still draining; everything else (activations, every store) comes after griddepcontrol.wait.
    griddep_launch_dependents();
    float acc[2] = {0.0f, 0.0f};
#pragma unroll
    for (int s = 0; s < CS_STAGES - 1; ++s) issue(s, 2, false);
    griddep_wait();
#pragma unroll
    for (int s = 0; s < CS_STAGES - 1; ++s) issue(s, 1, true);
    for (int ch = 0; ch < n_chunks; ++ch) {
        cp_async_wait<CS_STAGES - 2>();                     // chunk ch has landed (this thread's copies) ...
        __syncthreads();                                    // ... and everybody's; everybody is also done with chunk ch - 1
        issue(ch + CS_STAGES - 1, 3, true);                 // refill the slot chunk ch - 1 used
        const float* sA = stage0 + (ch % CS_STAGES) * CS_STAGE_FLOATS;
        const float* sB = sA + CS_ROWS * CS_LD;
#pragma unroll
        for (int k4 = 0; k4 < CS_BK / 4; ++k4) {
            const float4 a = *reinterpret_cast<const float4*>(sA + ty * CS_LD + 4 * k4);
            const float4 b0 = *reinterpret_cast<const float4*>(sB + tx * CS_LD + 4 * k4);
            const float4 b1 = *reinterpret_cast<const float4*>(sB + (tx + 8) * CS_LD + 4 * k4);
            acc[0] = fmaf(a.x, b0.x, fmaf(a.y, b0.y, fmaf(a.z, b0.z, fmaf(a.w, b0.w, acc[0]))));
            acc[1] = fmaf(a.x, b1.x, fmaf(a.y, b1.y, fmaf(a.z, b1.z, fmaf(a.w, b1.w, acc[1]))));
        }
    }
    cp_async_wait<0>();

    // ---------------------------------------------------------------- epilogue
    const int row = m0 + ty;
    const bool row_ok = row < p.M;
    const int col[2] = {n0 + tx, n0 + tx + 8};
    const bool ok[2] = {row_ok && col[0] < p.N, row_ok && col[1] < p.N};
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (ok[j]) {
            if (p.bias != nullptr) acc[j] += __ldg(p.bias + col[j]);
            if (p.add_pre != nullptr) acc[j] += __ldg(p.add_pre + static_cast<size_t>(row) * p.N + col[j]);
        }
    if (EPI == EPI_PLAIN) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (ok[j]) p.out[static_cast<size_t>(row) * p.N + col[j]] = p.relu ? fmaxf(acc[j], 0.0f) : acc[j];
        return;
    }
    const int cluster_size = gridDim.y;
    if (EPI == EPI_BN_FWD) {
        double s[2] = {0.0, 0.0}, q[2] = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (ok[j]) {
                const double v = acc[j];
                s[j] = v; q[j] = v * v;
                if (p.out_pre != nullptr) p.out_pre[static_cast<size_t>(row) * p.N + col[j]] = acc[j];
            }
        cs_cluster_col_reduce(s, q, wred, part, tot, cluster_size);
        if (t < CS_COLS) {
            const int c = n0 + t;
            const double mean = tot[t][0] / p.M;
            double var = tot[t][1] / p.M - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps))), fm = static_cast<float>(mean);
            stat[t][0] = fm; stat[t][1] = rstd;
            if (c < p.N && blockIdx.y == 0) {
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
                const float g = __ldg(p.gamma + col[j]) * stat[tx + 8 * j][1], sh = __ldg(p.beta + col[j]) - stat[tx + 8 * j][0] * g;
                const size_t o = static_cast<size_t>(row) * p.N + col[j];
                float v = fmaxf(fmaf(acc[j], g, sh), 0.0f);
                if (p.mask != nullptr) v *= __ldg(p.mask + o);
                p.out[o] = v;
            }
        return;
    }
    // EPI_BN_BWD: acc = gradient w.r.t. the output of mask * relu(BN(bn_x))
    float xh[2] = {0.0f, 0.0f};
    double s[2] = {0.0, 0.0}, q[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (ok[j]) {
            const size_t o = static_cast<size_t>(row) * p.N + col[j];
            const float h = (__ldg(p.bn_x + o) - __ldg(p.mean + col[j])) * __ldg(p.rstd + col[j]);
            float d = acc[j];
            if (p.mask != nullptr) d *= __ldg(p.mask + o);
            if (!(fmaf(__ldg(p.gamma + col[j]), h, __ldg(p.beta + col[j])) > 0.0f)) d = 0.0f;
            acc[j] = d; xh[j] = h;
            s[j] = d; q[j] = static_cast<double>(d) * h;
        }
    cs_cluster_col_reduce(s, q, wred, part, tot, cluster_size);
    if (t < CS_COLS && n0 + t < p.N && blockIdx.y == 0) {
        p.dbeta[n0 + t] = static_cast<float>(tot[t][0]);
        p.dgamma[n0 + t] = static_cast<float>(tot[t][1]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (ok[j]) {
            const float gr = __ldg(p.gamma + col[j]) * __ldg(p.rstd + col[j]);
            const float mean_d = static_cast<float>(tot[tx + 8 * j][0] / p.M), mean_dx = static_cast<float>(tot[tx + 8 * j][1] / p.M);
            const size_t o = static_cast<size_t>(row) * p.N + col[j];
            float v = gr * (acc[j] - mean_d - xh[j] * mean_dx);
            if (p.add_post != nullptr) v += __ldg(p.add_post + o);
            p.out[o] = v;
        }
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
    PGV_CUDA(cudaLaunchKernelEx(&cfg, colslice_gemm_kernel<TB, EPI>, p));
    return 0;
}

}  // namespace pgv

using namespace pgv;

extern "C" {

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
