// Spectrogram front end: audio -> (mel) dB spectrogram, for a batch of clips.
//   kernel 1  split_audio_kernel   x -> x_hi (TF32 rna) + x_lo (TF32 of residual), zero padded to whole hops
//   kernel 2  gemm_tf32_kernel<DftProblem>   frames x windowed-DFT basis on tcgen05, 3xTF32, |.|/norm epilogue
//   kernel 3  gemm_tf32_kernel<MelProblem>   magnitude x mel filterbank on tcgen05, 3xTF32, dB/min-max epilogue
// Reference behaviour: utils/audio.py:24-54, 80-87; data/abstractbasedataset.py:129-131.
#include <math.h>
#include <string.h>

#include <vector>

#include "pgv_common.cuh"
#include "pgv_gemm.cuh"

namespace pgv {

// ------------------------------------------------------------------------------------------------ kernel 1
__global__ void split_audio_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int n_clips,
                                   int n_samples, int padded) {
    const size_t total = static_cast<size_t>(n_clips) * padded;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int clip = static_cast<int>(i / padded), s = static_cast<int>(i % padded);
        float v = (s < n_samples) ? x[static_cast<size_t>(clip) * n_samples + s] : 0.0f;
        const float h = to_tf32_rna(v);
        hi[i] = h;
        lo[i] = to_tf32_rna(v - h);
    }
}

__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float v = x[i], h = to_tf32_rna(v);
        hi[i] = h;
        lo[i] = to_tf32_rna(v - h);
    }
}

// amplitude -> output value: linear_to_log_scale (audio.py:52-54) + optional min-max normalisation
__device__ __forceinline__ float finish(float amp, int log_scale, float floor_amp, float nrm_a, float nrm_b) {
    return log_scale ? nrm_a * (20.0f * log10f(fmaxf(amp, floor_amp))) + nrm_b : amp;
}

// ------------------------------------------------------------------------------------------------ kernel 2
// Frame t of a clip, sample k, is segment (t + k/hop - pad_segs) of the hop-sized segments of the clip: a [128 frames x
// 32 samples] operand tile is therefore ONE box of the [clips, segments, hop] view of the audio, and out-of-range
// segments (the centre padding of torch.stft) are zero-filled by TMA.  The basis rows are ordered so that a 256-column
// tile holds cos rows of 128 bins followed by the -sin rows of the same bins (the all-zero sin row of bin 0 carries
// the cos row of the Nyquist bin), so one epilogue thread owns re and im of a bin.
struct DftProblem {
    static constexpr int BLOCK_N = 256, STAGES = 4, ACC_STAGES = 2;
    struct Params {
        CUtensorMap x_hi, x_lo, w_hi, w_lo;
        int n_clips, n_frames, m_tiles, n_tiles, kb_per_pass, hop, pad_segs, n_bins;
        // output: either the magnitude split for the mel contraction ...
        float *mag_hi, *mag_lo;
        int mag_ld, frames_pad;
        // ... or the final linear-frequency dB spectrogram [clips, n_bins, n_frames]
        float* out_db;
        float norm, floor_amp, nrm_a, nrm_b;   // dB' = nrm_a * dB + nrm_b
        int log_scale;
    };
    __device__ static void prefetch(const Params& p) {
        tma_prefetch_desc(&p.x_hi); tma_prefetch_desc(&p.x_lo); tma_prefetch_desc(&p.w_hi); tma_prefetch_desc(&p.w_lo);
    }
    __device__ static int num_tiles(const Params& p) { return p.n_clips * p.m_tiles * p.n_tiles; }
    __device__ static int num_k_blocks(const Params& p) { return 3 * p.kb_per_pass; }
    __device__ static void tile_coords(const Params& p, int tile, int& tm, int& tn) { tm = tile / p.n_tiles; tn = tile % p.n_tiles; }
    __device__ static void load(const Params& p, int tm, int tn, int kb, void* sA, void* sB, uint64_t* bar) {
        const int pass = kb / p.kb_per_pass, k0 = (kb % p.kb_per_pass) * GEMM_BLOCK_K;   // pass 0: lo*hi, 1: hi*lo, 2: hi*hi (see pass order note)
        const int clip = tm / p.m_tiles, mt = tm % p.m_tiles;
        tma_load_3d(sA, pass == 0 ? &p.x_lo : &p.x_hi, bar, k0 % p.hop, mt * GEMM_BLOCK_M + k0 / p.hop - p.pad_segs, clip);
        tma_load_2d(sB, pass == 1 ? &p.w_lo : &p.w_hi, bar, k0, tn * BLOCK_N);
    }
    __device__ static void epilogue(const Params& p, int tm, int tn, uint32_t taddr, int row) {
        const int clip = tm / p.m_tiles, t = (tm % p.m_tiles) * GEMM_BLOCK_M + row;
        const bool valid = t < p.n_frames;
        const float inv_norm_div = p.norm;
#pragma unroll 1
        for (int c = 0; c < 128; c += 16) {
            uint32_t re[16], im[16];
            tmem_ld16(taddr + c, re);
            tmem_ld16(taddr + 128 + c, im);
            tmem_ld_wait();
            if (!valid) continue;
            float mag[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float a = __uint_as_float(re[j]), b = __uint_as_float(im[j]);
                mag[j] = sqrtf(a * a + b * b) / inv_norm_div;
            }
            float nyquist = 0.0f;
            const bool first = (tn == 0 && c == 0);
            if (first) {   // column 0 holds bin 0 (purely real); its "im" slot holds the Nyquist bin (purely real)
                mag[0] = fabsf(__uint_as_float(re[0])) / inv_norm_div;
                nyquist = fabsf(__uint_as_float(im[0])) / inv_norm_div;
            }
            const int bin0 = tn * 128 + c;
            if (p.out_db == nullptr) {
                const size_t base = (static_cast<size_t>(clip) * p.frames_pad + t) * p.mag_ld;
                float h[16], l[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { h[j] = to_tf32_rna(mag[j]); l[j] = to_tf32_rna(mag[j] - h[j]); }
                float4* dh = reinterpret_cast<float4*>(p.mag_hi + base + bin0);
                float4* dl = reinterpret_cast<float4*>(p.mag_lo + base + bin0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dh[j] = make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
                    dl[j] = make_float4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
                }
                if (first) {
                    const float nh = to_tf32_rna(nyquist);
                    p.mag_hi[base + p.n_bins - 1] = nh;
                    p.mag_lo[base + p.n_bins - 1] = to_tf32_rna(nyquist - nh);
                }
            } else {
                float* o = p.out_db + (static_cast<size_t>(clip) * p.n_bins + bin0) * p.n_frames + t;
#pragma unroll
                for (int j = 0; j < 16; ++j) o[static_cast<size_t>(j) * p.n_frames] = finish(mag[j], p.log_scale, p.floor_amp, p.nrm_a, p.nrm_b);
                if (first)
                    p.out_db[(static_cast<size_t>(clip) * p.n_bins + p.n_bins - 1) * p.n_frames + t] =
                        finish(nyquist, p.log_scale, p.floor_amp, p.nrm_a, p.nrm_b);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ kernel 3
struct MelProblem {
    static constexpr int BLOCK_N = 144, STAGES = 6, ACC_STAGES = 2;
    struct Params {
        CUtensorMap a_hi, a_lo, m_hi, m_lo;
        int n_clips, n_frames, m_tiles, n_tiles, kb_per_pass, n_mels;
        float* out;   // [clips, n_mels, n_frames]
        float floor_amp, nrm_a, nrm_b;
        int log_scale;
    };
    __device__ static void prefetch(const Params& p) {
        tma_prefetch_desc(&p.a_hi); tma_prefetch_desc(&p.a_lo); tma_prefetch_desc(&p.m_hi); tma_prefetch_desc(&p.m_lo);
    }
    __device__ static int num_tiles(const Params& p) { return p.n_clips * p.m_tiles * p.n_tiles; }
    __device__ static int num_k_blocks(const Params& p) { return 3 * p.kb_per_pass; }
    __device__ static void tile_coords(const Params& p, int tile, int& tm, int& tn) { tm = tile / p.n_tiles; tn = tile % p.n_tiles; }
    __device__ static void load(const Params& p, int tm, int tn, int kb, void* sA, void* sB, uint64_t* bar) {
        const int pass = kb / p.kb_per_pass, k0 = (kb % p.kb_per_pass) * GEMM_BLOCK_K;
        tma_load_3d(sA, pass == 0 ? &p.a_lo : &p.a_hi, bar, k0, (tm % p.m_tiles) * GEMM_BLOCK_M, tm / p.m_tiles);
        tma_load_2d(sB, pass == 1 ? &p.m_lo : &p.m_hi, bar, k0, tn * BLOCK_N);
    }
    __device__ static void epilogue(const Params& p, int tm, int tn, uint32_t taddr, int row) {
        const int clip = tm / p.m_tiles, t = (tm % p.m_tiles) * GEMM_BLOCK_M + row;
        const bool valid = t < p.n_frames;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c, v);
            tmem_ld_wait();
            if (!valid) continue;
            const int m0 = tn * BLOCK_N + c;
            float* o = p.out + (static_cast<size_t>(clip) * p.n_mels + m0) * p.n_frames + t;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (m0 + j < p.n_mels)
                    o[static_cast<size_t>(j) * p.n_frames] = finish(__uint_as_float(v[j]), p.log_scale, p.floor_amp, p.nrm_a, p.nrm_b);
        }
    }
};

struct FrontendGeometry {
    int n_frames, padded, n_segs, n_bins, m_tiles, frames_pad, mag_ld;
    size_t off_xhi, off_xlo, off_mhi, off_mlo, total;
};

static bool frontend_geometry(int n_clips, int n_samples, int n_fft, int hop, int n_mels, FrontendGeometry* g) {
    if (n_clips <= 0 || n_samples <= 0) return false;
    if (n_fft < 256 || n_fft > 4096 || (n_fft & (n_fft - 1)) != 0) return false;
    if (hop <= 0 || hop % 32 != 0 || (n_fft / 2) % hop != 0) return false;
    g->n_frames = 1 + n_samples / hop;
    g->padded = ceil_div(n_samples, hop) * hop;
    g->n_segs = g->padded / hop;
    g->n_bins = n_fft / 2 + 1;
    g->m_tiles = ceil_div(g->n_frames, GEMM_BLOCK_M);
    g->frames_pad = g->m_tiles * GEMM_BLOCK_M;
    g->mag_ld = pgv_frontend_mel_ld(n_fft);
    size_t off = 0;
    g->off_xhi = off; off += align_up(sizeof(float) * n_clips * g->padded, 256);
    g->off_xlo = off; off += align_up(sizeof(float) * n_clips * g->padded, 256);
    g->off_mhi = off;
    if (n_mels > 0) off += align_up(sizeof(float) * n_clips * g->frames_pad * g->mag_ld, 256);
    g->off_mlo = off;
    if (n_mels > 0) off += align_up(sizeof(float) * n_clips * g->frames_pad * g->mag_ld, 256);
    g->total = off;
    return true;
}

template <class P>
static int launch_gemm(const pgv_handle* h, const typename P::Params& prm, int n_tiles, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        PGV_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<P>::TOTAL));
        configured = true;
    }
    const int grid = n_tiles < h->sm_count ? n_tiles : h->sm_count;
    gemm_tf32_kernel<P><<<grid, GEMM_THREADS, GemmSmem<P>::TOTAL, stream>>>(prm);
    PGV_LAUNCH_CHECK();
    return 0;
}

}  // namespace pgv

using namespace pgv;

extern "C" {

int pgv_frontend_num_frames(int n_samples, int hop) { return hop > 0 ? 1 + n_samples / hop : 0; }

int pgv_frontend_mel_ld(int n_fft) { return (n_fft / 2 + 1 + 31) / 32 * 32; }   // 513 -> 544: whole 32-wide k-blocks

size_t pgv_frontend_workspace_bytes(int n_clips, int n_samples, int n_fft, int hop, int n_mels) {
    FrontendGeometry g;
    return frontend_geometry(n_clips, n_samples, n_fft, hop, n_mels, &g) ? g.total : 0;
}

int pgv_frontend_launch_count(int n_mels) { return n_mels > 0 ? 3 : 2; }

int pgv_frontend_init_constants(pgv_handle* h, const float* window_host, int n_fft, float* basis_hi, float* basis_lo,
                                const float* mel_host, int n_mels, float* mel_hi, float* mel_lo) {
    PGV_CHECK_ARG(h && window_host && basis_hi && basis_lo, "pgv_frontend_init_constants: NULL argument");
    PGV_CHECK_ARG(n_fft >= 256 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "n_fft=%d unsupported", n_fft);
    auto split = [](double v, float* hi, float* lo) {
        // round-to-nearest to 10 explicit mantissa bits (ties away, like cvt.rna.tf32.f32)
        auto rna = [](float f) {
            uint32_t u;
            memcpy(&u, &f, 4);
            u = (u + 0x1000u) & 0xFFFFE000u;
            float r;
            memcpy(&r, &u, 4);
            return r;
        };
        const float h = rna(static_cast<float>(v));
        *hi = h;
        *lo = rna(static_cast<float>(v - static_cast<double>(h)));
    };
    const int half = n_fft / 2;
    std::vector<float> hi(static_cast<size_t>(n_fft) * n_fft), lo(hi.size());
    const double two_pi = 6.283185307179586476925286766559;
    for (int tile = 0; tile < half / 128; ++tile) {
        for (int j = 0; j < 256; ++j) {
            const bool is_im = j >= 128;
            int bin = tile * 128 + (j & 127);
            bool cosine = !is_im;
            if (is_im && bin == 0) { bin = half; cosine = true; }   // Nyquist bin rides in bin 0's empty sin row
            float* rh = &hi[static_cast<size_t>(tile * 256 + j) * n_fft];
            float* rl = &lo[static_cast<size_t>(tile * 256 + j) * n_fft];
            for (int n = 0; n < n_fft; ++n) {
                const int phase = static_cast<int>((static_cast<long long>(bin) * n) % n_fft);   // exact argument reduction
                const double ang = two_pi * phase / n_fft;
                const double v = static_cast<double>(window_host[n]) * (cosine ? cos(ang) : -sin(ang));
                split(v, &rh[n], &rl[n]);
            }
        }
    }
    PGV_CUDA(cudaMemcpy(basis_hi, hi.data(), hi.size() * sizeof(float), cudaMemcpyHostToDevice));
    PGV_CUDA(cudaMemcpy(basis_lo, lo.data(), lo.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (mel_host && n_mels > 0) {
        PGV_CHECK_ARG(mel_hi && mel_lo, "mel_hi / mel_lo are NULL");
        const int n_bins = half + 1, ld = pgv_frontend_mel_ld(n_fft);
        std::vector<float> mh(static_cast<size_t>(n_mels) * ld, 0.0f), ml(mh.size(), 0.0f);
        for (int m = 0; m < n_mels; ++m)
            for (int k = 0; k < n_bins; ++k)
                split(mel_host[static_cast<size_t>(m) * n_bins + k], &mh[static_cast<size_t>(m) * ld + k], &ml[static_cast<size_t>(m) * ld + k]);
        PGV_CUDA(cudaMemcpy(mel_hi, mh.data(), mh.size() * sizeof(float), cudaMemcpyHostToDevice));
        PGV_CUDA(cudaMemcpy(mel_lo, ml.data(), ml.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

int pgv_frontend_fwd(pgv_handle* h, const float* audio, int n_clips, int n_samples, int n_fft, int hop, const float* basis_hi,
                     const float* basis_lo, const float* mel_hi, const float* mel_lo, int n_mels, float min_dB,
                     float norm_factor, int log_scale, int normalize, float spec_min, float spec_max, float* out,
                     void* workspace, size_t workspace_bytes, pgv_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PGV_CHECK_ARG(h && audio && basis_hi && basis_lo && out && workspace, "pgv_frontend_fwd: NULL argument");
    FrontendGeometry g;
    PGV_CHECK_ARG(frontend_geometry(n_clips, n_samples, n_fft, hop, n_mels, &g),
                  "pgv_frontend_fwd: unsupported geometry n_fft=%d hop=%d n_samples=%d", n_fft, hop, n_samples);
    PGV_CHECK_ARG(workspace_bytes >= g.total, "pgv_frontend_fwd: workspace too small (%zu < %zu)", workspace_bytes, g.total);
    PGV_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pgv_frontend_fwd: workspace must be 256-byte aligned");
    PGV_CHECK_ARG(n_mels <= 0 || (mel_hi && mel_lo), "pgv_frontend_fwd: mel operands are NULL");
    PGV_CHECK_ARG(norm_factor > 0.0f, "pgv_frontend_fwd: norm_factor must be positive");
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* x_hi = reinterpret_cast<float*>(ws + g.off_xhi);
    float* x_lo = reinterpret_cast<float*>(ws + g.off_xlo);
    float* mag_hi = reinterpret_cast<float*>(ws + g.off_mhi);
    float* mag_lo = reinterpret_cast<float*>(ws + g.off_mlo);
    const float floor_amp = powf(10.0f, min_dB / 20.0f);
    float nrm_a = 1.0f, nrm_b = 0.0f;
    if (normalize && log_scale) {   // -1 + (dB - min) / ((max - min)/2)
        nrm_a = 2.0f / (spec_max - spec_min);
        nrm_b = -1.0f - spec_min * nrm_a;
    }
    {
        const size_t total = static_cast<size_t>(n_clips) * g.padded;
        int blocks = static_cast<int>((total + 255) / 256);
        if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
        split_audio_kernel<<<blocks, 256, 0, stream>>>(audio, x_hi, x_lo, n_clips, n_samples, g.padded);
        PGV_LAUNCH_CHECK();
    }
    {
        DftProblem::Params p;
        memset(&p, 0, sizeof(p));
        const uint64_t xd[3] = {static_cast<uint64_t>(hop), static_cast<uint64_t>(g.n_segs), static_cast<uint64_t>(n_clips)};
        const uint64_t xs[2] = {static_cast<uint64_t>(hop) * 4, static_cast<uint64_t>(g.padded) * 4};
        const uint32_t xb[3] = {GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
        int rc;
        if ((rc = make_tmap_f32(h, &p.x_hi, x_hi, 3, xd, xs, xb))) return rc;
        if ((rc = make_tmap_f32(h, &p.x_lo, x_lo, 3, xd, xs, xb))) return rc;
        const uint64_t wd[2] = {static_cast<uint64_t>(n_fft), static_cast<uint64_t>(n_fft)};
        const uint64_t wst[1] = {static_cast<uint64_t>(n_fft) * 4};
        const uint32_t wb[2] = {GEMM_BLOCK_K, DftProblem::BLOCK_N};
        if ((rc = make_tmap_f32(h, &p.w_hi, basis_hi, 2, wd, wst, wb))) return rc;
        if ((rc = make_tmap_f32(h, &p.w_lo, basis_lo, 2, wd, wst, wb))) return rc;
        p.n_clips = n_clips; p.n_frames = g.n_frames; p.m_tiles = g.m_tiles; p.n_tiles = (n_fft / 2) / 128;
        p.kb_per_pass = n_fft / GEMM_BLOCK_K; p.hop = hop; p.pad_segs = (n_fft / 2) / hop; p.n_bins = g.n_bins;
        p.mag_hi = mag_hi; p.mag_lo = mag_lo; p.mag_ld = g.mag_ld; p.frames_pad = g.frames_pad;
        p.out_db = n_mels > 0 ? nullptr : out;
        p.norm = norm_factor; p.floor_amp = floor_amp; p.nrm_a = nrm_a; p.nrm_b = nrm_b; p.log_scale = log_scale;
        if ((rc = launch_gemm<DftProblem>(h, p, n_clips * p.m_tiles * p.n_tiles, stream))) return rc;
    }
    if (n_mels > 0) {
        MelProblem::Params p;
        memset(&p, 0, sizeof(p));
        const uint64_t ad[3] = {static_cast<uint64_t>(g.n_bins), static_cast<uint64_t>(g.n_frames), static_cast<uint64_t>(n_clips)};
        const uint64_t as[2] = {static_cast<uint64_t>(g.mag_ld) * 4, static_cast<uint64_t>(g.frames_pad) * g.mag_ld * 4};
        const uint32_t ab[3] = {GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
        int rc;
        if ((rc = make_tmap_f32(h, &p.a_hi, mag_hi, 3, ad, as, ab))) return rc;
        if ((rc = make_tmap_f32(h, &p.a_lo, mag_lo, 3, ad, as, ab))) return rc;
        const uint64_t md[2] = {static_cast<uint64_t>(g.n_bins), static_cast<uint64_t>(n_mels)};
        const uint64_t ms[1] = {static_cast<uint64_t>(g.mag_ld) * 4};
        const uint32_t mb[2] = {GEMM_BLOCK_K, MelProblem::BLOCK_N};
        if ((rc = make_tmap_f32(h, &p.m_hi, mel_hi, 2, md, ms, mb))) return rc;
        if ((rc = make_tmap_f32(h, &p.m_lo, mel_lo, 2, md, ms, mb))) return rc;
        p.n_clips = n_clips; p.n_frames = g.n_frames; p.m_tiles = g.m_tiles; p.n_tiles = ceil_div(n_mels, MelProblem::BLOCK_N);
        p.kb_per_pass = g.mag_ld / GEMM_BLOCK_K; p.n_mels = n_mels; p.out = out;
        p.floor_amp = floor_amp; p.nrm_a = nrm_a; p.nrm_b = nrm_b; p.log_scale = log_scale;
        if ((rc = launch_gemm<MelProblem>(h, p, n_clips * p.m_tiles * p.n_tiles, stream))) return rc;
    }
    return 0;
}

int pgv_frontend_fwd_host(pgv_handle* h, const float* audio_host, float* audio_dev, int n_clips, int n_samples, int n_fft, int hop,
                          const float* basis_hi, const float* basis_lo, const float* mel_hi, const float* mel_lo, int n_mels,
                          float min_dB, float norm_factor, int log_scale, int normalize, float spec_min, float spec_max,
                          float* out_dev, float* out_host, void* workspace, size_t workspace_bytes, pgv_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PGV_CHECK_ARG(audio_host && audio_dev && out_dev && out_host, "pgv_frontend_fwd_host: NULL argument");
    PGV_CUDA(cudaMemcpyAsync(audio_dev, audio_host, sizeof(float) * n_clips * static_cast<size_t>(n_samples),
                             cudaMemcpyHostToDevice, stream));
    int rc = pgv_frontend_fwd(h, audio_dev, n_clips, n_samples, n_fft, hop, basis_hi, basis_lo, mel_hi, mel_lo, n_mels, min_dB,
                              norm_factor, log_scale, normalize, spec_min, spec_max, out_dev, workspace, workspace_bytes, stream_);
    if (rc) return rc;
    const int F = n_mels > 0 ? n_mels : n_fft / 2 + 1;
    const size_t out_bytes = sizeof(float) * n_clips * static_cast<size_t>(F) * pgv_frontend_num_frames(n_samples, hop);
    PGV_CUDA(cudaMemcpyAsync(out_host, out_dev, out_bytes, cudaMemcpyDeviceToHost, stream));
    PGV_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

int pgv_split_tf32(const float* x, float* hi, float* lo, size_t n, pgv_stream_t stream) {
    PGV_CHECK_ARG(x && hi && lo, "pgv_split_tf32: NULL argument");
    if (n == 0) return 0;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_tf32_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, hi, lo, n);
    PGV_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
