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

#define PGV_VERSION 102
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

/* General fp32 CUDA-core GEMM: C[M,N] = act(opA(A) * opB(B) + bias[N] + residual[M,N]), row-major.  trans_a = 0: A is
 * [M,K], 1: A is [K,M]; trans_b = 0: B is [K,N], 1: B is [N,K].  bias / residual may be NULL.  Long-K problems with few
 * output tiles are split along K (fp32 atomic accumulation).  Linear backward: dX = gemm(0,0,dY,W), dW = gemm(1,0,dY,X). */
PGV_API int pgv_gemm_f32(pgv_handle* h, int trans_a, int trans_b, const float* a, int lda, const float* b, int ldb, float* c, int ldc,
                         int m, int n, int k, const float* bias, int act, const float* residual, int ldr, pgv_stream_t stream);
/* out[f] = sum over rows of x[B,F] (bias gradients). */
PGV_API int pgv_colsum(const float* x, float* out, int B, int F, pgv_stream_t stream);
/* Linear-layer gradients in exact fp32 (nflows ResidualNet conditioners, VAE.py:118-125, flows.py:42-90): dw [N, K] = dy^T x,
 * db [N] = column sums of dy [M, N] (NULL to skip); the flow-sized problems get db from the GEMM kernel itself. */
PGV_API int pgv_linear_wgrad_f32(pgv_handle* h, const float* dy, const float* x, float* dw, float* db, int M, int N, int K,
                                 pgv_stream_t stream);

/* Column-slice GEMM family for the flow conditioners (nflows ResidualNet / ResidualBlock: VAE.py:118-125, flows.py:42-90,
 * regression.py:142-148): a thread-block cluster owns 16 output columns for all M <= pgv_colslice_max_rows() rows (32 rows per
 * CTA), so BatchNorm1d batch statistics stay on chip and the normalisation is fused into the GEMM.  Exact fp32.  x [M, K],
 * w [N, K] (nn.Linear).  The two plain entry points need no cluster and take any M.
 *   pgv_linear_cs_fwd       y = act(x w^T + bias + residual)
 *   pgv_linear_cs_dgrad     dx [M, K] = dy [M, N] w
 *   pgv_linear_bn_fwd       y_pre = x w^T + bias + residual (stored if non-NULL); out = mask * relu(BN(y_pre)) with batch
 *                           statistics (saved in save_mean / save_rstd) and running-stat update: Linear -> BatchNorm1d -> ReLU -> Dropout
 *   pgv_linear_dgrad_bn_bwd dt = dy w; dx = backward of mask * relu(BN(bn_x)) at dt, + add_post; dgamma / dbeta of that BatchNorm */
PGV_API int pgv_colslice_max_rows(void);
PGV_API int pgv_linear_cs_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y, int M, int N, int K,
                              int relu, pgv_stream_t stream);
PGV_API int pgv_linear_cs_dgrad(const float* dy, const float* w, float* dx, int M, int N, int K, pgv_stream_t stream);
PGV_API int pgv_linear_bn_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y_pre, float* out,
                              const float* gamma, const float* beta, const float* mask, float* save_mean, float* save_rstd,
                              float* running_mean, float* running_var, float momentum, float eps, int M, int N, int K,
                              pgv_stream_t stream);
PGV_API int pgv_linear_dgrad_bn_bwd(const float* dy, const float* w, const float* bn_x, const float* gamma, const float* beta,
                                    const float* mean, const float* rstd, const float* mask, const float* add_post, float* dx,
                                    float* dgamma, float* dbeta, int M, int N, int K, pgv_stream_t stream);

/* ------------------------------------------------------------------ convolutions (NCHW fp32)
 * nn.Conv2d / nn.ConvTranspose2d call sites: model/layer.py:19,38 (blocks of model/encoder.py:233-259 and
 * model/decoder.py:199-220).  All three take the geometry of the *convolution*: x [B,Cin,H,W], y/dy [B,Cout,Ho,Wo],
 * w [Cout,Cin,kh,kw].  A ConvTranspose2d with weight [Cin_t,Cout_t,kh,kw] is the data-gradient of the convolution with
 * that same weight tensor (Cout = Cin_t, Cin = Cout_t, H/W = the transposed conv's OUTPUT size incl. output_padding):
 *   tconv forward = pgv_conv2d_dgrad (with bias + activation),  tconv dgrad = pgv_conv2d_fwd,  tconv wgrad =
 *   pgv_conv2d_wgrad(x = tconv grad_output, dy = tconv input).
 * lrelu_slope < 0 disables the fused LeakyReLU.  kh*kw <= 25. */
PGV_API int pgv_conv2d_fwd_f32(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H, int W, int Cout, int kh,
                               int kw, int stride, int pad, int Ho, int Wo, float lrelu_slope, pgv_stream_t stream);
PGV_API int pgv_conv2d_dgrad_f32(const float* dy, const float* w, const float* bias, float* dx, int B, int Cin, int H, int W, int Cout,
                                 int kh, int kw, int stride, int pad, int Ho, int Wo, float lrelu_slope, pgv_stream_t stream);
/* dw [Cout,Cin,kh,kw] and (if db != NULL) db [Cout] are overwritten. */
PGV_API int pgv_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int H, int W, int Cout, int kh,
                                 int kw, int stride, int pad, int Ho, int Wo, pgv_stream_t stream);

/* The same three operations as implicit GEMMs on the tcgen05 tensor cores (TF32 products with round-to-nearest
 * operands, fp32 accumulation in TMEM; see csrc/pgv_conv_tc.cu).  Same argument meaning as the _f32 entry points.
 * dgrad supports stride 1 and 2; wgrad accumulates split-K partial sums with fp32 atomics (dw is zeroed first) and
 * does not produce db (use pgv_channel_sum). */
PGV_API int pgv_conv2d_fwd_tf32(pgv_handle* h, const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H, int W,
                                int Cout, int kh, int kw, int stride, int pad, int Ho, int Wo, float lrelu_slope, pgv_stream_t stream);
PGV_API int pgv_conv2d_dgrad_tf32(pgv_handle* h, const float* dy, const float* w, const float* bias, float* dx, int B, int Cin, int H, int W,
                                  int Cout, int kh, int kw, int stride, int pad, int Ho, int Wo, float lrelu_slope, pgv_stream_t stream);
PGV_API int pgv_conv2d_wgrad_tf32(pgv_handle* h, const float* x, const float* dy, float* dw, int B, int Cin, int H, int W, int Cout, int kh,
                                  int kw, int stride, int pad, int Ho, int Wo, pgv_stream_t stream);
/* Debug: subsequent pgv_conv2d_*_tf32 / pgv_linear_*_tf32 launches record pipeline timestamps of CTA 0 into trace_dev
 * (3 x 64 x 8 int64, device memory); NULL disables.  Used by tools/gpu_trace_conv.py only. */
PGV_API int pgv_debug_set_conv_trace(void* trace_dev);
/* Direct streaming kernels (exact fp32) for the two thin full-resolution layers, enc1 = Conv2d(1,8,5,2,2)
 * (encoder.py:241) and dec8 = ConvTranspose2d(8,1,5,2,2) (decoder.py:218), which are HBM-bound (8 flop/byte).  Conv-view
 * geometry: x [B,1,H,W], y [B,C<=8,Ho,Wo], w [C,1,5,5], stride 2, pad 2.  _dgrad is the transposed convolution and
 * clamps its result to [clamp_lo, clamp_hi] (the decoder's Hardtanh; pass -INF/+INF for none).  channels_last != 0: the
 * C-channel tensor (y / dy) is stored [B, Ho, Wo, C] instead of [B, C, Ho, Wo]; round_out rounds y to TF32 (nearest).
 * _wgrad: with a workspace of >= 592 x 200 floats the per-block partial sums are combined in a fixed order; without, fp32 atomics. */
PGV_API int pgv_conv5x5s2_c1_supported(int Cin, int Cout, int kh, int kw, int stride, int pad, int H, int W, int Ho, int Wo);
PGV_API int pgv_conv5x5s2_c1_fwd(const float* x, const float* w, const float* bias, float* y, int B, int C, int H, int W, int Ho, int Wo,
                                 float lrelu_slope, int channels_last, int round_out, pgv_stream_t stream);
PGV_API int pgv_conv5x5s2_c1_dgrad(const float* y, const float* w, const float* bias, float* x, int B, int C, int H, int W, int Ho, int Wo,
                                   float clamp_lo, float clamp_hi, int channels_last, pgv_stream_t stream);
PGV_API int pgv_conv5x5s2_c1_wgrad(const float* x, const float* dy, float* dw, int B, int C, int H, int W, int Ho, int Wo,
                                   int channels_last, void* ws, size_t ws_bytes, pgv_stream_t stream);
/* ---------------------------------------------------------------------------------------------------------------------
 * Channels-last (NHWC) convolution path: the default 'tf32' route of the Conv2D / TConv2D blocks (model/layer.py:10-46,
 * encoder.py:233-259, decoder.py:190-221).  Activations are [B, H, W, C]; the reduction index is (kh, kw, c), so every
 * operand tile is filled with 16-byte cp.async copies (no register staging) and multiplied by tcgen05.mma kind::tf32.
 * Operands are consumed as stored: callers pass TF32-rounded activations (the *_cl producers below have a round_out
 * flag) and the weight matrices made by pgv_conv_cl_prep_weights.  Supported: 4x4 / stride 2 / pad 2 and 1x1 / stride 1
 * (pgv_conv_cl_supported); channel counts multiples of 8 (4x4) or 32 (1x1); all pointers 16-byte aligned. */
PGV_API int pgv_conv_cl_supported(int Cin, int Cout, int KH, int KW, int stride, int pad);
/* w [Cout, Cin, KH, KW] (PyTorch) -> wf [Cout][(kh, kw, ci)] (forward operand) and / or wq (data-gradient operand:
 * [(ph, pw, ci)][(a, b, co)] = w[co, ci, ph + 2(1-a), pw + 2(1-b)] for 4x4/s2/p2, [ci][co] for 1x1); either may be NULL. */
PGV_API int pgv_conv_cl_prep_weights(const float* w, float* wf, float* wq, int Cout, int Cin, int KH, int KW, int stride, int pad,
                                     pgv_stream_t stream);
/* y [B, Ho, Wo, Cout] = lrelu(bias + conv(x [B, H, W, Cin], w)); lrelu_slope < 0: no activation. */
PGV_API int pgv_conv_cl_fwd(pgv_handle* h, const float* x, const float* wf, const float* bias, float* y, int B, int H, int W, int Cin,
                            int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo, float lrelu_slope, int round_out,
                            pgv_stream_t stream);
/* dx [B, H, W, Cin] = lrelu(bias + conv_transpose(dy [B, Ho, Wo, Cout], w)): data gradient of the convolution above and
 * forward of nn.ConvTranspose2d(weight = w) (bias has Cin entries or is NULL). */
PGV_API int pgv_conv_cl_dgrad(pgv_handle* h, const float* dy, const float* wq, const float* bias, float* dx, int B, int H, int W, int Cin,
                              int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo, float lrelu_slope, int round_out,
                              pgv_stream_t stream);
/* The same two calls for a convolution followed by (Leaky)ReLU and BatchNorm2d (model/layer.py:10-46 of the reference): the epilogue
 * also accumulates the batch statistics of what it stores, bn_sums[2c] = sum and bn_sums[2c + 1] = sum of squares of channel c
 * (zero-filled by the call; only for launches with one N tile: Cout <= 128, or 4 * Cin <= 128 for the 4x4 data gradient), which
 * pgv_bn_cl_train_apply then turns into the normalisation; bn_sums may be NULL.
 * bn_bwd_x (optional, needs bn_sums; a tensor of the output's shape and layout): the second statistic becomes the sum of
 * stored value * bn_bwd_x instead of the sum of squares.  In the backward pass the stored values are the gradient flowing into the
 * BatchNorm2d whose input bn_bwd_x was, so bn_sums then holds exactly what that BatchNorm's backward has to reduce over the batch
 * (pgv_bn_cl_train_bwd, raw_sums): its reduction pass over two activation-sized tensors disappears.
 * ws / ws_bytes (optional, 16-byte aligned, ws_bytes >= pgv_conv_cl_workspace_bytes() of ZERO-FILLED header + room for partial tiles;
 * the header is left zero by every call): with it, launches whose tiles do not fill the GPU split the reduction over several CTAs and
 * combine the partial accumulators in a fixed order (deterministic; no atomics).  One workspace per stream.
 * The activation operand is fetched by TMA: tiled boxes for 1x1 / stride 1, im2col-mode loads (cuTensorMapEncodeIm2col) for 4x4 /
 * stride 2 and for the 2x2 window of the data gradient when the channel count is 8, 16 or a multiple of 32; cp.async gathers otherwise. */
PGV_API int pgv_conv_cl_fwd_bn(pgv_handle* h, const float* x, const float* wf, const float* bias, float* y, int B, int H, int W, int Cin,
                               int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo, float lrelu_slope, int round_out,
                               double* bn_sums, const float* bn_bwd_x, void* ws, size_t ws_bytes, pgv_stream_t stream);
PGV_API int pgv_conv_cl_dgrad_bn(pgv_handle* h, const float* dy, const float* wq, const float* bias, float* dx, int B, int H, int W, int Cin,
                                 int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo, float lrelu_slope, int round_out,
                                 double* bn_sums, const float* bn_bwd_x, void* ws, size_t ws_bytes, pgv_stream_t stream);
/* Weight gradient of the convolution: sum over pixels of dy x patch(x), split over CTAs.  oihw_layout != 0: dw is the PyTorch tensor
 * [Cout, Cin, KH, KW] (e.g. a slice of a flat gradient buffer); else the forward-operand matrix [Cout][(kh, kw, ci)].
 * With a workspace the partial sums are combined in a fixed order by a finish kernel (deterministic); without one, fp32 atomics into
 * the zero-filled destination. */
PGV_API int pgv_conv_cl_wgrad(pgv_handle* h, const float* x, const float* dy, float* dw, int oihw_layout, int B, int H, int W, int Cin,
                              int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo, void* ws, size_t ws_bytes, pgv_stream_t stream);
/* nn.Linear on the same kernel, for layers large enough for the tensor cores (encoder / decoder FC, encoder.py:84, decoder.py:70).
 * Operands as stored: x / dy TF32-rounded with 16-byte-aligned rows, wr = rounded weights [N, K], wt = rounded transposed weights
 * [K, N]; K may carry zero padding (pgv_round_copy).  dw is written with row pitch lddw, columns < k_valid.  ws as above. */
PGV_API int pgv_linear_cl_fwd(pgv_handle* h, const float* x, const float* wr, const float* bias, float* y, int M, int N, int K,
                              void* ws, size_t ws_bytes, pgv_stream_t stream);
PGV_API int pgv_linear_cl_dgrad(pgv_handle* h, const float* dy, const float* wt, float* dx, int M, int N, int K, void* ws, size_t ws_bytes,
                                pgv_stream_t stream);
PGV_API int pgv_linear_cl_wgrad(pgv_handle* h, const float* dy, const float* x, float* dw, int lddw, int M, int N, int K, int k_valid,
                                void* ws, size_t ws_bytes, pgv_stream_t stream);
/* Bytes of the zero-filled header at the start of a channels-last workspace. */
PGV_API int pgv_conv_cl_workspace_bytes(void);
/* Debug / A-B switch: programmatic dependent launch of the stand-alone flow kernels (gather / column-slice GEMMs / couplings / scatter):
 * 1 lets each of them become resident and prefetch its weights while its predecessor drains; 0 (default: measured faster on B200)
 * launches them fully serialised. */
PGV_API int pgv_debug_set_pdl(int on);
/* Debug / A-B switch: 0 forces the cp.async gather of the activation operand, -1 (default) lets the library choose TMA where it can. */
PGV_API int pgv_debug_set_conv_a_mode(int mode);
/* dst [rows, ldd] = TF32-rounded src [rows, lds] (first `cols` columns), zero in columns cols..ldd-1. */
PGV_API int pgv_round_copy(const float* src, int lds, float* dst, int ldd, int rows, int cols, pgv_stream_t stream);
PGV_API int pgv_conv_cl_unpack_dw(const float* dwcl, float* dw, int Cout, int Cin, int KH, int KW, pgv_stream_t stream);
/* BatchNorm2d on channels-last tensors viewed as [P = B*H*W, C] (same semantics as pgv_bn2d_*; C % 4 == 0).
 * workspace: 24*C bytes.  dx_colsum (optional, [C]): per-channel sums of dx = bias gradient of the convolution feeding the block. */
PGV_API int pgv_bn_cl_train_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean, float* save_rstd,
                                float* running_mean, float* running_var, float momentum, float eps, size_t P, int C, int round_out,
                                void* workspace, pgv_stream_t stream);
PGV_API int pgv_bn_cl_train_apply(const float* x, const double* sums, const float* gamma, const float* beta, float* y, float* save_mean,
                                  float* save_rstd, float* running_mean, float* running_var, float momentum, float eps, size_t P, int C,
                                  int round_out, pgv_stream_t stream);
PGV_API int pgv_bn_cl_eval_fwd(const float* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                               float* y, float eps, size_t P, int C, int round_out, pgv_stream_t stream);
/* raw_sums (optional): raw_sums[2c] = sum(dy), raw_sums[2c + 1] = sum(dy * x) of channel c over the P rows, as accumulated by the
 * convolution that produced dy (pgv_conv_cl_fwd_bn / pgv_conv_cl_dgrad_bn with bn_bwd_x = x): the reduction pass is skipped. */
PGV_API int pgv_bn_cl_train_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_rstd,
                                float* dx, float* dgamma, float* dbeta, float* dx_colsum, float lrelu_slope, size_t P, int C, int round_out,
                                const double* raw_sums, void* workspace, pgv_stream_t stream);
/* out[c] = sum over rows of x [P, C] (bias gradient).  workspace: 16*C bytes. */
PGV_API int pgv_colsum_cl(const float* x, float* out, size_t P, int C, void* workspace, pgv_stream_t stream);
/* dx = dy * (a > 0 ? 1 : slope), flat arrays of n elements (n % 4 == 0), optionally rounded to TF32. */
PGV_API int pgv_lrelu_bwd_round(const float* dy, const float* a, float* dx, float slope, size_t n, int round_out, pgv_stream_t stream);
/* src [batch, R, S] -> dst [batch, S, R] (NCHW -> NHWC: R = C, S = H*W; NHWC -> NCHW: R = H*W, S = C). */
PGV_API int pgv_transpose_inner(const float* src, float* dst, int batch, int R, int S, int round_out, pgv_stream_t stream);

/* nn.Linear on the same tensor-core kernel (any K, no alignment requirement): x [M,K], w [N,K], y [M,N] = act(x w^T +
 * bias + residual) (bias / residual may be NULL, relu != 0 fuses a ReLU); dx [M,K] = dy w; dw [N,K] = dy^T x. */
PGV_API int pgv_linear_fwd_tf32(pgv_handle* h, const float* x, const float* w, const float* bias, const float* residual, float* y, int M,
                                int N, int K, int relu, pgv_stream_t stream);
PGV_API int pgv_linear_dgrad_tf32(pgv_handle* h, const float* dy, const float* w, float* dx, int M, int N, int K, pgv_stream_t stream);
PGV_API int pgv_linear_wgrad_tf32(pgv_handle* h, const float* dy, const float* x, float* dw, int M, int N, int K, pgv_stream_t stream);
/* out[c] = sum over (b, h, w) of x[b,c,h,w] (bias gradient of a transposed convolution). */
PGV_API int pgv_channel_sum(const float* x, float* out, int B, int C, int HW, pgv_stream_t stream);

/* ------------------------------------------------------------------ normalisation
 * BatchNorm2d after the activation of every conv block (model/layer.py:20-26): training mode uses batch statistics
 * (biased variance), updates running_mean / running_var (unbiased) with `momentum`, saves mean and 1/sqrt(var+eps).
 * workspace: 16*C bytes.  The backward optionally continues through the LeakyReLU that produced x (lrelu_slope >= 0). */
PGV_API int pgv_bn2d_train_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean, float* save_rstd,
                               float* running_mean, float* running_var, float momentum, float eps, int B, int C, int HW, void* workspace,
                               pgv_stream_t stream);
PGV_API int pgv_bn2d_eval_fwd(const float* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                              float* y, float eps, int B, int C, int HW, pgv_stream_t stream);
PGV_API int pgv_bn2d_train_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_rstd, float* dx,
                               float* dgamma, float* dbeta, float lrelu_slope, int B, int C, int HW, void* workspace, pgv_stream_t stream);
/* dx = dy * (a > 0 ? 1 : slope): LeakyReLU backward from its OUTPUT a (conv blocks without BatchNorm). */
PGV_API int pgv_lrelu_bwd(const float* dy, const float* a, float* dx, float slope, size_t n, pgv_stream_t stream);
/* BatchNorm1d on [B,F] (encoder.py:86-87 and the nflows ResidualBlock): y = mask * relu?(gamma*xhat+beta); mask NULL = none. */
PGV_API int pgv_bn1d_train_fwd(const float* x, const float* gamma, const float* beta, const float* mask, float* y, float* save_mean,
                               float* save_rstd, float* running_mean, float* running_var, float momentum, float eps, int relu, int B, int F,
                               pgv_stream_t stream);
PGV_API int pgv_bn1d_eval_fwd(const float* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                              float* y, float eps, int relu, int B, int F, pgv_stream_t stream);
PGV_API int pgv_bn1d_train_bwd(const float* dy, const float* x, const float* gamma, const float* beta, const float* save_mean,
                               const float* save_rstd, const float* mask, float* dx, float* dgamma, float* dbeta, int relu, int B, int F,
                               pgv_stream_t stream);
/* nflows transforms.normalization.BatchNorm (model/flows.py:87-88): batch mean / UNBIASED variance, weight =
 * softplus(u) + eps; logdet_scalar[0] = sum_f(log w_f - 0.5*log(var_f + eps)) is the value added to every row's log|det J|.
 * grad_logdet_sum[0] = sum over rows of dL/dlogdet. */
PGV_API int pgv_flowbn_train_fwd(const float* x, const float* unconstrained_weight, const float* bias, float* y, float* save_mean,
                                 float* save_var, float* running_mean, float* running_var, float* logdet_scalar, float momentum, float eps,
                                 int B, int F, pgv_stream_t stream);
PGV_API int pgv_flowbn_eval(const float* x, const float* unconstrained_weight, const float* bias, const float* running_mean,
                            const float* running_var, float* y, float* logdet_scalar, float eps, int inverse, int B, int F, pgv_stream_t stream);
PGV_API int pgv_flowbn_train_bwd(const float* dy, const float* x, const float* unconstrained_weight, const float* save_mean,
                                 const float* save_var, const float* grad_logdet_sum, float* dx, float* d_unconstrained_weight, float* dbias,
                                 float eps, int B, int F, pgv_stream_t stream);

/* ------------------------------------------------------------------ latent space / flows ([B, D] tensors)
 * Reparameterisation (model/VAE.py:167-174): z0 = mu + exp(logvar/2)*eps, mu_logvar is [B,2,D]; eps NULL = eval (z0 = mu).
 * The backward adds d_mu_logvar_add (e.g. the latent-loss gradient) when it is not NULL. */
PGV_API int pgv_reparam_fwd(const float* mu_logvar, const float* eps, float* z0, int B, int D, pgv_stream_t stream);
PGV_API int pgv_reparam_bwd(const float* dz0, const float* mu_logvar, const float* eps, const float* d_mu_logvar_add, float* d_mu_logvar,
                            int B, int D, pgv_stream_t stream);
PGV_API int pgv_gather_cols(const float* x, const int* idx, float* out, int B, int D, int n, pgv_stream_t stream);
PGV_API int pgv_scatter_add_cols(float* dst, const int* idx, const float* src, int B, int D, int n, pgv_stream_t stream);
/* nflows AffineCouplingTransform (via model/flows.py:42-90, model/VAE.py:118-125): params[b, :n_t] = shift,
 * params[b, n_t:] = unconstrained scale, s = sigmoid(u+2)+1e-3; forward y = x*s+t, logdet_out = logdet_in + sum log s
 * (logdet_in may be NULL); inverse != 0: y = (x-t)/s, logdet_out = logdet_in - sum log s. */
PGV_API int pgv_coupling_fwd(const float* x, const float* params, const int* identity_idx, const int* transform_idx, float* y,
                             const float* logdet_in, float* logdet_out, int B, int D, int n_identity, int n_transform, int inverse,
                             pgv_stream_t stream);
PGV_API int pgv_coupling_bwd(const float* dy, const float* dlogdet, const float* x, const float* params, const int* identity_idx,
                             const int* transform_idx, float* dx, float* dparams, int B, int D, int n_identity, int n_transform,
                             pgv_stream_t stream);
/* nn.Hardtanh (decoder.py:98 output activation, regression.py:22 PresetActivation). */
PGV_API int pgv_hardtanh_fwd(const float* x, float* y, float lo, float hi, size_t n, pgv_stream_t stream);
PGV_API int pgv_hardtanh_bwd(const float* dy, const float* x, float* dx, float lo, float hi, size_t n, pgv_stream_t stream);
/* y = x * m (dropout with a pre-scaled keep mask) and y = a + b. */
PGV_API int pgv_mul(const float* x, const float* m, float* y, size_t n, pgv_stream_t stream);
PGV_API int pgv_add(const float* a, const float* b, float* y, size_t n, pgv_stream_t stream);
/* y[i] = a[i] + scalar_dev[0] (a NULL = 0): adds the flow-BatchNorm log-det scalar to every row's log|det J|. */
PGV_API int pgv_add_scalar(const float* a, const float* scalar_dev, float* y, size_t n, pgv_stream_t stream);
/* PresetActivation with cat_softmax_activation=True (regression.py:47-50). */
PGV_API int pgv_preset_act_softmax_fwd(const float* x, float* y, const int* num_cols, int n_num, const int* grp_start, const int* grp_len,
                                       int n_grp, int B, int D, pgv_stream_t stream);
PGV_API int pgv_preset_act_softmax_bwd(const float* dy, const float* x, const float* y, float* dx, const int* num_cols, int n_num,
                                       const int* grp_start, const int* grp_len, int n_grp, int B, int D, pgv_stream_t stream);

/* ------------------------------------------------------------------ losses
 * Scalars live in device memory: *_fwd writes loss_out[0]; *_bwd reads the upstream gradient from grad_out[0].
 * pgv_sqerr: loss = scale * sum((a-b)^2)  (nn.MSELoss: scale = 1/n; L2Loss, model/loss.py:15-43: scale = 1/B).  workspace >= 8 B. */
PGV_API int pgv_sqerr_fwd(const float* a, const float* b, size_t n, double scale, float* loss_out, void* workspace, pgv_stream_t stream);
PGV_API int pgv_sqerr_bwd(const float* a, const float* b, size_t n, double scale, const float* grad_out, float* da, pgv_stream_t stream);
/* FlowVAE.latent_loss (model/VAE.py:183-193, utils/probability.py:13-29).  workspace >= 8 B. */
PGV_API int pgv_latent_loss_fwd(const float* mu_logvar, const float* z0, const float* zk, const float* logdet, int B, int D, int normalize,
                                float* loss_out, void* workspace, pgv_stream_t stream);
PGV_API int pgv_latent_loss_bwd(const float* grad_out, const float* mu_logvar, const float* z0, const float* zk, int B, int D, int normalize,
                                float* d_mu_logvar, float* dz0, float* dzk, float* dlogdet, pgv_stream_t stream);
/* GaussianDkl (model/loss.py:46-66).  workspace >= 8 B. */
PGV_API int pgv_dkl_fwd(const float* mu_logvar, int B, int D, int normalize, float* loss_out, void* workspace, pgv_stream_t stream);
PGV_API int pgv_dkl_bwd(const float* grad_out, const float* mu_logvar, int B, int D, int normalize, float* d_mu_logvar, pgv_stream_t stream);
/* ------------------------------------------------------------------ monitoring metrics, inference tail, inverse-flow loss
 * Per-VST-parameter tables (one entry per monitored parameter): kind = 0 numerical learned as numerical, 1 numerical learned as
 * one-hot, 2 categorical learned as numerical, 3 categorical learned as one-hot, 4 not learnable; col / len = learnable column range;
 * card = cardinality (<= 0: continuous).
 * pgv_preset_metrics = QuantizedNumericalParamsLoss (model/loss.py:187-261; l1 != 0: L1 instead of MSE) and CategoricalParamsAccuracy
 * (model/loss.py:265-315) in one pass, as train.py:232-233 calls them every step: out4 = (numerical loss, mean accuracy * acc_scale,
 * number of numerical, number of categorical parameters); acc (optional, [P]) = per-parameter accuracies (reduce=False), -1 elsewhere;
 * partial = [P] floats of workspace. */
PGV_API int pgv_preset_metrics(const float* v_out, const float* v_in, int B, int L, const int* kind, const int* col, const int* len,
                               const int* card, int P, int l1, float acc_scale, float* partial, float* out4, float* acc, pgv_stream_t stream);
/* PresetsParams.get_full from learnable presets (data/preset.py:350-369): full [B, P]; fill[p] = default value of a non-learnable
 * parameter (-0.1 when it has none). */
PGV_API int pgv_learnable_to_full(const float* v, int B, int L, const int* kind, const int* col, const int* len, const int* card,
                                  const float* fill, int P, float* full, pgv_stream_t stream);
/* FlowParamsLoss (model/loss.py:318-346): -mean_b(log N(z0; mu, exp(logvar)) + logdet_t + logdet_u) / divisor (1000 in the reference).
 * rows_ws: [B] floats. */
PGV_API int pgv_flow_params_loss_fwd(const float* mu_logvar, const float* z0, const float* logdet_t, const float* logdet_u, int B, int D,
                                     float divisor, float* loss_out, float* rows_ws, pgv_stream_t stream);
PGV_API int pgv_flow_params_loss_bwd(const float* grad_out, const float* mu_logvar, const float* z0, int B, int D, float divisor,
                                     float* d_mu_logvar, float* dz0, float* dlogdet, pgv_stream_t stream);
/* utils/exception.py:13-22 (train.py:245) without a host round trip: flags[0] |= 1 << i if scalar i (device pointer, may be NULL) is NaN. */
PGV_API int pgv_nan_flags(const float* s0, const float* s1, const float* s2, const float* s3, const float* s4, int* flags, pgv_stream_t stream);
/* data/abstractbasedataset.py:348-391: per_item4[i] = (min, max, mean, unbiased variance) of spectrogram i (x [N, elems]); dataset4
 * (optional) = (min of mins, max of maxes, mean of means, sqrt(mean of variances)). */
PGV_API int pgv_spectrogram_stats(const float* x, int N, size_t elems, float* per_item4, float* dataset4, pgv_stream_t stream);
/* A recorded chain of flow ops (gather / column-slice Linear with fused BatchNorm forward or backward / coupling / scatter / small
 * weight gradients) as ONE persistent launch with grid barriers between dependent ops: `ops` = host array of n_ops records of
 * pgv_flow_program_op_bytes() bytes each (kind, barrier_after, the parameter block of the stand-alone kernel: see csrc/pgv_flow_fused.cu
 * and model/ops.py), at most pgv_flow_program_max_ops(); M = batch rows (<= pgv_colslice_max_rows()); counter = pgv_flow_program_max_ops() + 1 device words (the grid barrier and one tile counter per weight-gradient op; zeroed by the call).
 * Replaces ~50 (forward) / ~90 (backward) launches of 5-10 us per RealNVP flow (model/flows.py:42-90, VAE.py:118-125). */
PGV_API int pgv_flow_program(pgv_handle* h, const void* ops, int n_ops, int M, unsigned* counter, pgv_stream_t stream);
/* Debug (tools/gpu_flow_trace.py): subsequent pgv_flow_program launches write, per op, 4 words {start ns, end of CTA 0's share ns, barrier open ns, kind} */
PGV_API int pgv_debug_set_flow_trace(void* trace_dev);
/* Experiment knob: epilogue groups of the channels-last conv kernel: 0 = choose (default), 1 = always one, 2 = three for the quad epilogue only. */
PGV_API int pgv_debug_set_conv_groups(int mode);
/* Experiment knob: CTAs per SM at which the grids of the thin 5x5 forward / transposed kernels are capped (default 32). */
PGV_API int pgv_debug_set_thin_grid_mult(int m);
/* Experiment knob: number of clusters (of ceil(M / 32) CTAs) of the flow program kernel; 0 = as many as fit (default). */
PGV_API int pgv_debug_set_flow_clusters(int clusters);
/* Experiment knob: rows of the [pixels, C] matrix each thread of the channels-last BatchNorm kernels walks before the grid is capped (default 16). */
PGV_API int pgv_debug_set_bn_rows_per_lane(int rows);
/* Debug: 6 words of pinned host memory that a timed-out grid barrier fills with {0xdead, cta, op, counter, target, n_ctas} before it traps. */
PGV_API int pgv_debug_set_flow_diag(void* pinned_host);
PGV_API int pgv_flow_program_op_bytes(void);
PGV_API int pgv_flow_program_max_ops(void);
/* Hint: pull [p, p + bytes) into L2 (one prefetch per 128-byte line); used on a flow's 15 MB of conditioner weights right before the
 * chain of small kernels that read them once each. */
PGV_API int pgv_l2_prefetch(const void* p, size_t bytes, pgv_stream_t stream);
/* Backward of the inverse direction of the affine coupling (pgv_coupling_fwd with inverse != 0), given its result x_out: needed by
 * FlowParamsLoss, which back-propagates through the inverse latent flow (model/VAE.py:128-131, regression.py:179-184). */
PGV_API int pgv_coupling_inv_bwd(const float* dx_out, const float* dlogdet, const float* x_out, const float* params, const int* id_idx,
                                 const int* tr_idx, float* dy, float* dparams, int B, int D, int n_id, int n_t, pgv_stream_t stream);
/* SynthParamsLoss (model/loss.py:73-183) with the useless-parameter rule of data/preset.py:247-283 evaluated on the
 * device from v_in: tables as returned by PresetIndexesHelper.device_tables().  cat_softmax != 0 applies
 * softmax(q / temperature) inside the loss.  The same workspace must be passed to fwd and bwd. */
PGV_API size_t pgv_synth_loss_workspace_bytes(int n_groups);
PGV_API int pgv_synth_loss_fwd(const float* v_out, const float* v_in, int B, int L, const int* num_cols, const int* num_vol_col, int n_num,
                               const int* grp_start, const int* grp_len, const int* grp_vol_col, int n_grp, int normalize,
                               float cat_loss_factor, int cat_softmax, float softmax_temperature, const double* group_counts,
                               float* loss_out, void* workspace, pgv_stream_t stream);
/* counts[g] = rows of v_in that are useful for categorical group g (its operator is not silent).  Data-parallel training all-reduces
 * them and passes (sum / world) as `group_counts` above, so that the mean of the per-rank losses is the loss of the gathered batch
 * (the reference's DataParallel computes the criterion on the gathered outputs, train.py:241 / loss.py:172); NULL = this batch's own. */
PGV_API int pgv_synth_useful_counts(const float* v_in, int B, int L, const int* grp_vol_col, int n_grp, double* counts, pgv_stream_t stream);
PGV_API int pgv_synth_loss_bwd(const float* grad_out, const float* v_out, const float* v_in, int B, int L, const int* num_cols,
                               const int* num_vol_col, int n_num, const int* grp_start, const int* grp_len, const int* grp_vol_col, int n_grp,
                               int normalize, float cat_loss_factor, int cat_softmax, float softmax_temperature, const void* workspace,
                               float* d_v_out, pgv_stream_t stream);

/* ------------------------------------------------------------------ optimizer (train.py:165-167)
 * torch.optim.Adam semantics with L2-in-gradient weight decay on one flat fp32 buffer; `step` counts from 1.
 * The _dev variant reads {lr, 1-beta1^t, sqrt(1-beta2^t), grad_scale} from device memory (CUDA-graph replay). */
/* flat[off_i .. off_i+n_i) = scale * src_i for n_tensors tensors in ONE launch; table_dev: device array of 3*n uint64
 * {source address, destination offset in elements, element count}; max_elems = largest n_i. */
PGV_API int pgv_multi_pack(const void* table_dev, int n_tensors, size_t max_elems, float* flat, float scale, pgv_stream_t stream);
PGV_API int pgv_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int step, float grad_scale, pgv_stream_t stream);
PGV_API int pgv_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, const float* hyper_dev,
                              float beta1, float beta2, float eps, float weight_decay, pgv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PGV_H_ */
