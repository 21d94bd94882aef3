// Dense layer GEMMs:  C[M,N] = act(A[M,K] * B[N,K]^T + bias)   (nn.Linear layout: B is the [out, in] weight).
//   pgv_gemm_nt_tf32 : tcgen05 / TMEM / TMA path (1xTF32 or error-compensated 3xTF32)
//   pgv_gemm_nt_f32  : CUDA-core fp32 path (exact products) for tiny shapes and on-device cross-checks
#include <string.h>

#include "pgv_common.cuh"
#include "pgv_gemm.cuh"

namespace pgv {

struct DenseProblem {
    static constexpr int BLOCK_N = 128, STAGES = 6, ACC_STAGES = 2;
    struct Params {
        CUtensorMap a, a_lo, b, b_lo;
        int m, n, k, m_tiles, n_tiles, kb_per_pass, passes;
        float* c;
        int ldc;
        const float* bias;
        int act;
    };
    __device__ static void prefetch(const Params& p) {
        tma_prefetch_desc(&p.a); tma_prefetch_desc(&p.b);
        if (p.passes == 3) { tma_prefetch_desc(&p.a_lo); tma_prefetch_desc(&p.b_lo); }
    }
    __device__ static int num_tiles(const Params& p) { return p.m_tiles * p.n_tiles; }
    __device__ static int num_k_blocks(const Params& p) { return p.passes * p.kb_per_pass; }
    __device__ static void tile_coords(const Params& p, int tile, int& tm, int& tn) { tm = tile / p.n_tiles; tn = tile % p.n_tiles; }
    __device__ static void load(const Params& p, int tm, int tn, int kb, void* sA, void* sB, uint64_t* bar) {
        // 3-pass order: lo*hi, hi*lo, then hi*hi (corrections first, see pgv_gemm.cuh)
        const int pass = (p.passes == 3) ? kb / p.kb_per_pass : 2, k0 = (kb % p.kb_per_pass) * GEMM_BLOCK_K;
        tma_load_2d(sA, pass == 0 ? &p.a_lo : &p.a, bar, k0, tm * GEMM_BLOCK_M);
        tma_load_2d(sB, pass == 1 ? &p.b_lo : &p.b, bar, k0, tn * BLOCK_N);
    }
    __device__ static void epilogue(const Params& p, int tm, int tn, uint32_t taddr, int row) {
        const int r = tm * GEMM_BLOCK_M + row;
        const bool row_ok = r < p.m;
        float* crow = p.c + static_cast<size_t>(r) * p.ldc;
        const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c, v);
            tmem_ld_wait();
            const int n0 = tn * BLOCK_N + c;
            if (!row_ok || n0 >= p.n) continue;
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x = __uint_as_float(v[j]);
                if (p.bias != nullptr && n0 + j < p.n) x += p.bias[n0 + j];
                if (p.act == 1) x = fmaxf(x, 0.0f);
                o[j] = x;
            }
            if (vec_ok && n0 + 16 <= p.n) {
                float4* d = reinterpret_cast<float4*>(crow + n0);
#pragma unroll
                for (int j = 0; j < 4; ++j) d[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n0 + j < p.n) crow[n0 + j] = o[j];
            }
        }
    }
};

// Generic fp32 GEMM on the CUDA cores: C[M,N] = act(opA(A) * opB(B) + bias + residual), row-major storage.
//   TA = 0: A is [M, K]      TA = 1: A is [K, M]          TB = 0: B is [K, N]      TB = 1: B is [N, K] (nn.Linear weight)
// 64x64 output tile per block, 16x16 threads, 4x4 outputs per thread, K step 16; grid.z splits K (atomic accumulate).
template <int TA, int TB>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                                       float* __restrict__ c, int ldc, int m, int n, int k, const float* __restrict__ bias,
                                                       int act, const float* __restrict__ residual, int ldr, int k_per_split) {
    __shared__ float sa[16][64 + 4], sb[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int kbeg = blockIdx.z * k_per_split, kend = min(k, kbeg + k_per_split);
    float acc[4][4] = {};
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            {   // A tile: make the contiguous storage dimension the fastest-varying thread index
                const int r = TA ? (i & 63) : (i >> 4), kk = TA ? (i >> 6) : (i & 15);
                float v = 0.0f;
                if (m0 + r < m && k0 + kk < kend) v = TA ? a[static_cast<size_t>(k0 + kk) * lda + m0 + r] : a[static_cast<size_t>(m0 + r) * lda + k0 + kk];
                sa[kk][r] = v;
            }
            {
                const int r = TB ? (i >> 4) : (i & 63), kk = TB ? (i & 15) : (i >> 6);
                float v = 0.0f;
                if (n0 + r < n && k0 + kk < kend) v = TB ? b[static_cast<size_t>(n0 + r) * ldb + k0 + kk] : b[static_cast<size_t>(k0 + kk) * ldb + n0 + r];
                sb[kk][r] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = sa[kk][ty * 4 + i]; bv[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = m0 + ty * 4 + i, col = n0 + tx * 4 + j;
            if (r < m && col < n) {
                float x = acc[i][j];
                if (gridDim.z == 1) {
                    if (bias) x += bias[col];
                    if (residual) x += residual[static_cast<size_t>(r) * ldr + col];
                    if (act == 1) x = fmaxf(x, 0.0f);
                    c[static_cast<size_t>(r) * ldc + col] = x;
                } else {   // split-K: C was zero-filled; split 0 carries bias and residual
                    if (blockIdx.z == 0) {
                        if (bias) x += bias[col];
                        if (residual) x += residual[static_cast<size_t>(r) * ldr + col];
                    }
                    atomicAdd(c + static_cast<size_t>(r) * ldc + col, x);
                }
            }
        }
}

// Small-problem variant (flow conditioner layers: M = batch <= ~1k, N, K ~ 300-610): 32x32 output tile per block, 256
// threads with 2x2 outputs each, K step 32, next k-chunk prefetched into registers while the current one is multiplied.
// A 160x300x300 layer becomes 50 blocks of 10 short iterations instead of 15 blocks of 19 long ones.
template <int TA, int TB>
__global__ void __launch_bounds__(256) gemm_f32_small_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                                             float* __restrict__ c, int ldc, int m, int n, int k,
                                                             const float* __restrict__ bias, int act, const float* __restrict__ residual, int ldr,
                                                             float* __restrict__ a_colsum) {
    // a_colsum (optional, TA only): out[r] = sum_k A[k][r], r < m: the bias gradient of a Linear layer, for free next to dW = dY^T X
    __shared__ float sa[32][33], sb[32][33];          // [k][row]
    const int t = threadIdx.x, m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = t & 15, ty = t >> 4;               // outputs (ty*2 + i, tx*2 + j)
    // loader mapping: 1024 elements per operand per chunk -> 4 per thread; contiguous storage dimension fastest
    int ar[4], ak[4], br[4], bk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int id = t + q * 256;
        ar[q] = TA ? (id & 31) : (id >> 5); ak[q] = TA ? (id >> 5) : (id & 31);
        br[q] = TB ? (id >> 5) : (id & 31); bk[q] = TB ? (id & 31) : (id >> 5);
    }
    auto lda_ = [&](int q, int k0) -> float {
        const int r = m0 + ar[q], kk = k0 + ak[q];
        if (r >= m || kk >= k) return 0.0f;
        return TA ? __ldg(a + static_cast<size_t>(kk) * lda + r) : __ldg(a + static_cast<size_t>(r) * lda + kk);
    };
    auto ldb_ = [&](int q, int k0) -> float {
        const int r = n0 + br[q], kk = k0 + bk[q];
        if (r >= n || kk >= k) return 0.0f;
        return TB ? __ldg(b + static_cast<size_t>(r) * ldb + kk) : __ldg(b + static_cast<size_t>(kk) * ldb + r);
    };
    float pa[4], pb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { pa[q] = lda_(q, 0); pb[q] = ldb_(q, 0); }
    float acc[2][2] = {};
    float cs0 = 0.0f, cs1 = 0.0f;
    const bool do_colsum = TA && a_colsum != nullptr && blockIdx.x == 0 && tx == 0;
    for (int k0 = 0; k0 < k; k0 += 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { sa[ak[q]][ar[q]] = pa[q]; sb[bk[q]][br[q]] = pb[q]; }
        __syncthreads();
        if (k0 + 32 < k) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { pa[q] = lda_(q, k0 + 32); pb[q] = ldb_(q, k0 + 32); }
        }
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
            const float a0 = sa[kk][ty * 2], a1 = sa[kk][ty * 2 + 1], b0 = sb[kk][tx * 2], b1 = sb[kk][tx * 2 + 1];
            acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
            if (do_colsum) { cs0 += a0; cs1 += a1; }
        }
        __syncthreads();
    }
    if (do_colsum) {
        if (m0 + ty * 2 < m) a_colsum[m0 + ty * 2] = cs0;
        if (m0 + ty * 2 + 1 < m) a_colsum[m0 + ty * 2 + 1] = cs1;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = m0 + ty * 2 + i, col = n0 + tx * 2 + j;
            if (r < m && col < n) {
                float x = acc[i][j];
                if (bias) x += bias[col];
                if (residual) x += residual[static_cast<size_t>(r) * ldr + col];
                if (act == 1) x = fmaxf(x, 0.0f);
                c[static_cast<size_t>(r) * ldc + col] = x;
            }
        }
}

}  // namespace pgv

using namespace pgv;

extern "C" {

int pgv_gemm_nt_tf32(pgv_handle* h, const float* a, const float* a_lo, int lda, const float* b, const float* b_lo, int ldb,
                     float* c, int ldc, int m, int n, int k, const float* bias, int act, int three_pass, pgv_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PGV_CHECK_ARG(h && a && b && c, "pgv_gemm_nt_tf32: NULL argument");
    PGV_CHECK_ARG(m > 0 && n > 0 && k > 0, "pgv_gemm_nt_tf32: empty problem %dx%dx%d", m, n, k);
    PGV_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0 && lda >= k && ldb >= k, "pgv_gemm_nt_tf32: lda/ldb must be >= k and multiples of 4");
    PGV_CHECK_ARG(!three_pass || (a_lo && b_lo), "pgv_gemm_nt_tf32: three_pass needs a_lo and b_lo");
    PGV_CHECK_ARG(act == 0 || act == 1, "pgv_gemm_nt_tf32: unknown activation %d", act);
    DenseProblem::Params p;
    memset(&p, 0, sizeof(p));
    const uint64_t ad[2] = {static_cast<uint64_t>(k), static_cast<uint64_t>(m)}, as[1] = {static_cast<uint64_t>(lda) * 4};
    const uint64_t bd[2] = {static_cast<uint64_t>(k), static_cast<uint64_t>(n)}, bs[1] = {static_cast<uint64_t>(ldb) * 4};
    const uint32_t abox[2] = {GEMM_BLOCK_K, GEMM_BLOCK_M}, bbox[2] = {GEMM_BLOCK_K, DenseProblem::BLOCK_N};
    int rc;
    if ((rc = make_tmap_f32(h, &p.a, a, 2, ad, as, abox))) return rc;
    if ((rc = make_tmap_f32(h, &p.b, b, 2, bd, bs, bbox))) return rc;
    if (three_pass) {
        if ((rc = make_tmap_f32(h, &p.a_lo, a_lo, 2, ad, as, abox))) return rc;
        if ((rc = make_tmap_f32(h, &p.b_lo, b_lo, 2, bd, bs, bbox))) return rc;
    }
    p.m = m; p.n = n; p.k = k;
    p.m_tiles = ceil_div(m, GEMM_BLOCK_M); p.n_tiles = ceil_div(n, DenseProblem::BLOCK_N);
    p.kb_per_pass = ceil_div(k, GEMM_BLOCK_K); p.passes = three_pass ? 3 : 1;
    p.c = c; p.ldc = ldc; p.bias = bias; p.act = act;
    static bool configured = false;
    if (!configured) {
        PGV_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<DenseProblem>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GemmSmem<DenseProblem>::TOTAL));
        configured = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    gemm_tf32_kernel<DenseProblem><<<tiles < h->sm_count ? tiles : h->sm_count, GEMM_THREADS, GemmSmem<DenseProblem>::TOTAL, stream>>>(p);
    PGV_LAUNCH_CHECK();
    return 0;
}

int pgv_gemm_f32(pgv_handle* h, int trans_a, int trans_b, const float* a, int lda, const float* b, int ldb, float* c, int ldc, int m, int n,
                 int k, const float* bias, int act, const float* residual, int ldr, pgv_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PGV_CHECK_ARG(h && a && b && c, "pgv_gemm_f32: NULL argument");
    PGV_CHECK_ARG(m > 0 && n > 0 && k > 0, "pgv_gemm_f32: empty problem");
    PGV_CHECK_ARG(act == 0 || act == 1, "pgv_gemm_f32: unknown activation %d", act);
    if (static_cast<long long>(m) * n <= 1024LL * 1024 && k <= 4096) {   // small problem: many small tiles beat few big ones
        dim3 g(ceil_div(n, 32), ceil_div(m, 32));
#define PGV_SMALL_LAUNCH(TA, TB) gemm_f32_small_kernel<TA, TB><<<g, 256, 0, stream>>>(a, lda, b, ldb, c, ldc, m, n, k, bias, act, residual, ldr, nullptr)
        if (!trans_a && !trans_b) PGV_SMALL_LAUNCH(0, 0);
        else if (!trans_a && trans_b) PGV_SMALL_LAUNCH(0, 1);
        else if (trans_a && !trans_b) PGV_SMALL_LAUNCH(1, 0);
        else PGV_SMALL_LAUNCH(1, 1);
#undef PGV_SMALL_LAUNCH
        PGV_LAUNCH_CHECK();
        return 0;
    }
    const int tiles = ceil_div(n, 64) * ceil_div(m, 64);
    int splits = 1;
    if (act == 0 && k >= 2048 && tiles < h->sm_count * 2) {   // long-K, few tiles: split K to fill the chip
        splits = (h->sm_count * 4) / tiles;
        if (splits > k / 512) splits = k / 512;
        if (splits < 1) splits = 1;
    }
    const int k_per_split = ceil_div(ceil_div(k, splits), 16) * 16;
    splits = ceil_div(k, k_per_split);
    if (splits > 1) PGV_CUDA(cudaMemset2DAsync(c, sizeof(float) * ldc, 0, sizeof(float) * n, m, stream));
    dim3 grid(ceil_div(n, 64), ceil_div(m, 64), splits);
#define PGV_GEMM_LAUNCH(TA, TB) gemm_f32_kernel<TA, TB><<<grid, 256, 0, stream>>>(a, lda, b, ldb, c, ldc, m, n, k, bias, act, residual, ldr, k_per_split)
    if (!trans_a && !trans_b) PGV_GEMM_LAUNCH(0, 0);
    else if (!trans_a && trans_b) PGV_GEMM_LAUNCH(0, 1);
    else if (trans_a && !trans_b) PGV_GEMM_LAUNCH(1, 0);
    else PGV_GEMM_LAUNCH(1, 1);
#undef PGV_GEMM_LAUNCH
    PGV_LAUNCH_CHECK();
    return 0;
}

/* Linear-layer weight and bias gradients in exact fp32: dw [N, K] = dy^T x, db [N] = column sums of dy (dy [M, N], x [M, K]).
 * Small problems (the flow conditioner layers) get db from the GEMM kernel itself; db may be NULL. */
int pgv_linear_wgrad_f32(pgv_handle* h, const float* dy, const float* x, float* dw, float* db, int M, int N, int K, pgv_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PGV_CHECK_ARG(h && dy && x && dw && M > 0 && N > 0 && K > 0, "pgv_linear_wgrad_f32: bad argument");
    if (static_cast<long long>(N) * K <= 1024LL * 1024 && M <= 4096) {
        dim3 g(ceil_div(K, 32), ceil_div(N, 32));
        gemm_f32_small_kernel<1, 0><<<g, 256, 0, stream>>>(dy, N, x, K, dw, K, N, K, M, nullptr, 0, nullptr, 0, db);
        PGV_LAUNCH_CHECK();
        return 0;
    }
    if (int rc = pgv_gemm_f32(h, 1, 0, dy, N, x, K, dw, K, N, K, M, nullptr, 0, nullptr, 0, stream_)) return rc;
    return db != nullptr ? pgv_colsum(dy, db, M, N, stream_) : 0;
}

int pgv_gemm_nt_f32(pgv_handle* h, const float* a, int lda, const float* b, int ldb, float* c, int ldc, int m, int n, int k,
                    const float* bias, int act, pgv_stream_t stream) {
    return pgv_gemm_f32(h, 0, 1, a, lda, b, ldb, c, ldc, m, n, k, bias, act, nullptr, 0, stream);
}

}  // extern "C"
