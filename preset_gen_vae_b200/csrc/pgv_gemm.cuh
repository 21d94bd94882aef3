// Persistent, warp-specialised TF32 GEMM skeleton for sm_100a:  D[128 x BLOCK_N] (+)= A[128 x K] * B[BLOCK_N x K]^T
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : tcgen05.mma issuer (one thread), accumulators in TMEM, double-buffered across tiles
//   warps 2..5  : epilogue (tcgen05.ld -> registers -> problem-specific math -> global)
// Accumulation note (measured on B200): the fp32 accumulator in TMEM is updated with round-toward-zero, so the error
// of a long sum grows linearly with the number of tcgen05.mma instructions added into a LARGE accumulator.  Policies
// that use the error-compensated 3xTF32 product therefore issue the two small correction passes (lo*hi, hi*lo) first
// and the dominant hi*hi pass last, and long-K dense layers are split along K across CTAs.
// The "problem" policy P supplies the tile schedule, the TMA coordinates of every k-block and the epilogue, so the
// same skeleton runs the windowed-DFT contraction, the mel projection and the dense layers of the model.
#pragma once
#include "pgv_tc.cuh"

namespace pgv {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 32;                       // 32 fp32 = 128 B = one swizzle row
constexpr int GEMM_UMMA_K = 8;                         // kind::tf32
constexpr int GEMM_A_BYTES = GEMM_BLOCK_M * 128;       // 16 KB
constexpr int GEMM_THREADS = 192;

template <class P>
struct GemmSmem {
    static constexpr int B_BYTES = P::BLOCK_N * 128;
    static constexpr int STAGE_BYTES = GEMM_A_BYTES + B_BYTES;
    static constexpr int BAR_BYTES = 8 * (2 * P::STAGES + 2 * P::ACC_STAGES) + 16;
    static constexpr int TOTAL = 1024 /*alignment slack*/ + P::STAGES * STAGE_BYTES + BAR_BYTES;
    static_assert(B_BYTES % 1024 == 0, "BLOCK_N must be a multiple of 8 rows");
    static_assert(P::BLOCK_N % 16 == 0 && P::BLOCK_N >= 16 && P::BLOCK_N <= 256, "invalid UMMA N for M=128");
    static_assert(TOTAL <= 227 * 1024, "shared memory budget exceeded");
};

__host__ __device__ constexpr uint32_t tmem_cols_pow2(uint32_t n) {
    return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

template <class P>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tf32_kernel(const __grid_constant__ typename P::Params p) {
    using S = GemmSmem<P>;
    constexpr uint32_t TMEM_COLS = tmem_cols_pow2(P::ACC_STAGES * P::BLOCK_N);
    static_assert(P::ACC_STAGES * P::BLOCK_N <= 512, "TMEM budget exceeded");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + P::STAGES * S::STAGE_BYTES);
    uint64_t* bar_empty = bar_full + P::STAGES;
    uint64_t* bar_tfull = bar_empty + P::STAGES;
    uint64_t* bar_tempty = bar_tfull + P::ACC_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_tempty + P::ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) P::prefetch(p);
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < P::STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
            for (int a = 0; a < P::ACC_STAGES; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 4); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_tiles = P::num_tiles(p), n_kb = P::num_k_blocks(p);

    // Producer and MMA warps run their loops with all 32 lanes converged and let ONE elected lane issue: the operands of the TMA
    // / tcgen05.mma instructions then live in uniform registers (under `if (lane == 0)` the compiler has to wrap every such
    // instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop).
    if (warp == 0) {
        {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int tm, tn; P::tile_coords(p, tile, tm, tn);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&bar_full[stage], S::STAGE_BYTES);
                        uint8_t* sA = smem + stage * S::STAGE_BYTES;
                        P::load(p, tm, tn, kb, sA, sA + GEMM_A_BYTES, &bar_full[stage]);
                    }
                    __syncwarp();
                    if (++stage == P::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {
            constexpr uint32_t idesc = umma_idesc_tf32(GEMM_BLOCK_M, P::BLOCK_N);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * P::BLOCK_N;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&bar_full[stage], phase);
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint64_t da = umma_smem_desc_sw128(a_addr), db = umma_smem_desc_sw128(a_addr + GEMM_A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
                            // advance 32 B (= 2 x 16 B units) inside the 128 B swizzle row per UMMA_K step
                            umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit(&bar_empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == P::STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit(&bar_tfull[acc]);
                __syncwarp();
                if (++acc == P::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;   // TMEM lane quadrant this warp may read
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            int tm, tn; P::tile_coords(p, tile, tm, tn);
            mbar_wait(&bar_tfull[acc], acc_phase);
            tc_fence_after_sync();
            P::epilogue(p, tm, tn, tmem_base + acc * P::BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16), quad * 32 + lane);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);
            if (++acc == P::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace pgv
