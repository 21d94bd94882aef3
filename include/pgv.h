/* pgv.h — C ABI of libpgv.so, the B200 (sm_100a) hot path of gwendal-lv/preset-gen-vae.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; every entry point below replaces the library
 * call(s) the reference makes at the cited file:line (paths relative to the reference repository), and is what a
 * ctypes binding on the reference side would bind.
 *
 * Conventions
 *   - plain C types only: device/host pointers, explicit sizes and leading dimensions, scalar hyper-parameters;
 *   - every function returns 0 on success, <0 for a bad argument / unsupported shape, >0 for a cudaError_t or
 *     CUresult; pgv_last_error() returns the message of the last failure on the calling thread;
 *   - all device pointers (inputs, outputs, workspaces) are owned by the caller (PyTorch); the library never
 *     allocates device memory, never retains a pointer after the call, and never synchronises the device except
 *     in the *_host entry points, which say so;
 *   - `stream` is a cudaStream_t passed as void*; every device entry point is asynchronous on it and is
 *     CUDA-graph capturable.
 */
#ifndef PGV_H_
#define PGV_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGV_VERSION 100
#if defined(__GNUC__)
#define PGV_API __attribute__((visibility("default")))
#else
#define PGV_API
#endif

typedef struct pgv_handle pgv_handle;
typedef void* pgv_stream_t;

/* ------------------------------------------------------------------ library */
PGV_API int pgv_version(void);
PGV_API const char* pgv_last_error(void);
/* Binds to `device` (cudaSetDevice is NOT called; the caller's current device must be `device`), checks that it is
 * compute capability 10.x, resolves cuTensorMapEncodeTiled.  No device memory is allocated. */
PGV_API int pgv_init(pgv_handle** out, int device);
PGV_API void pgv_destroy(pgv_handle* h);
PGV_API int pgv_sm_count(const pgv_handle* h);

/* ------------------------------------------------------------------ spectrogram front end
 * Replaces, for a whole batch of clips at once:
 *   utils/audio.py:33-40  Spectrogram.get_stft        (torch.stft, centred, zero padded, symmetric Hann)
 *   utils/audio.py:42-54  Spectrogram.__call__        (|X| / norm, floor, 20*log10)
 *   utils/audio.py:80-87  MelSpectrogram.__call__     (mel_basis @ |X|/norm, floor, 20*log10)
 *   data/abstractbasedataset.py:129-131               (optional min-max normalisation to [-1, 1])
 * The STFT is evaluated as a windowed-DFT contraction on the tensor cores with an error-compensated 3xTF32
 * split (fp32-equivalent products, fp32 accumulation in TMEM); the mel projection is a second dense contraction.
 *
 * Supported: n_fft a power of two in [256, 4096]; hop a multiple of 32 that divides n_fft/2. */

/* Frames per clip = 1 + n_samples / hop (torch.stft, center=True). */
PGV_API int pgv_frontend_num_frames(int n_samples, int hop);
/* Bytes of device scratch pgv_frontend_fwd needs (0 if the arguments are unsupported). */
PGV_API size_t pgv_frontend_workspace_bytes(int n_clips, int n_samples, int n_fft, int hop, int n_mels);
/* Fills the constant operands.  window_host: n_fft floats (the reference's torch.hann_window(n_fft, periodic=False)).
 * basis_hi / basis_lo: device, [n_fft, n_fft] floats each (row = output column of the contraction).
 * mel_host: [n_mels, n_fft/2+1] floats row-major or NULL; mel_hi / mel_lo: device, [n_mels, mel_ld] floats each with
 * mel_ld = pgv_frontend_mel_ld(n_fft).  Uses synchronous cudaMemcpy (call once, outside any capture). */
PGV_API int pgv_frontend_mel_ld(int n_fft);
PGV_API int pgv_frontend_init_constants(pgv_handle* h, const float* window_host, int n_fft, float* basis_hi, float* basis_lo,
                                const float* mel_host, int n_mels, float* mel_hi, float* mel_lo);
/* audio: device [n_clips, n_samples] fp32.  out: device [n_clips, F, T] fp32 with F = n_mels if n_mels > 0 else
 * n_fft/2+1, T = pgv_frontend_num_frames().  norm_factor: max|rfft(window)| (audio.py:31).  log_scale == 0 returns
 * the linear magnitude (Spectrogram(log_scale=False), audio.py:47-50) and ignores min_dB / normalize.  If
 * normalize != 0 the result is -1 + (dB - spec_min) / ((spec_max - spec_min) / 2). */
PGV_API int pgv_frontend_fwd(pgv_handle* h, const float* audio, int n_clips, int n_samples, int n_fft, int hop,
                     const float* basis_hi, const float* basis_lo, const float* mel_hi, const float* mel_lo, int n_mels,
                     float min_dB, float norm_factor, int log_scale, int normalize, float spec_min, float spec_max,
                     float* out, void* workspace, size_t workspace_bytes, pgv_stream_t stream);
/* Same with HOST buffers (audio_host, out_host; pinned memory recommended): copies the audio to `audio_dev`,
 * runs pgv_frontend_fwd into `out_dev`, copies the result back and synchronises `stream` before returning. */
PGV_API int pgv_frontend_fwd_host(pgv_handle* h, const float* audio_host, float* audio_dev, int n_clips, int n_samples, int n_fft,
                          int hop, const float* basis_hi, const float* basis_lo, const float* mel_hi, const float* mel_lo,
                          int n_mels, float min_dB, float norm_factor, int log_scale, int normalize, float spec_min,
                          float spec_max, float* out_dev, float* out_host, void* workspace, size_t workspace_bytes,
                          pgv_stream_t stream);
/* Number of kernels one pgv_frontend_fwd call launches (for bench.py's gpu_launches). */
PGV_API int pgv_frontend_launch_count(int n_mels);

/* ------------------------------------------------------------------ dense layers (nn.Linear call sites:
 * model/encoder.py:85, model/decoder.py:64, nflows ResidualNet linears via model/flows.py:66-75)
 * C[M,N] = act(A[M,K] * B[N,K]^T + bias[N]); fp32 storage, TF32 tensor-core products, fp32 accumulation.
 * lda/ldb are in elements and must be multiples of 4 (16-byte TMA strides); act: 0 none, 1 ReLU.
 * three_pass != 0 selects the error-compensated 3xTF32 product (needs a_lo/b_lo = residual operands, may alias
 * NULL otherwise). */
PGV_API int pgv_gemm_nt_tf32(pgv_handle* h, const float* a, const float* a_lo, int lda, const float* b, const float* b_lo, int ldb,
                     float* c, int ldc, int m, int n, int k, const float* bias, int act, int three_pass,
                     pgv_stream_t stream);
/* fp32 CUDA-core GEMM with the same contract (exact fp32 products).  Used by the GPU tests to cross-check the
 * tensor-core path on the device and by layers too small for a 128-row tile. */
PGV_API int pgv_gemm_nt_f32(pgv_handle* h, const float* a, int lda, const float* b, int ldb, float* c, int ldc, int m, int n, int k,
                    const float* bias, int act, pgv_stream_t stream);
/* hi = round-to-nearest TF32 of x, lo = TF32 of (x - hi); n elements, device pointers. */
PGV_API int pgv_split_tf32(const float* x, float* hi, float* lo, size_t n, pgv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PGV_H_ */
