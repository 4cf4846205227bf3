/* se_b200.h -- C-ABI of the B200-native spectral front/back-end (libse_b200.so).
 *
 * The reference (ooshyun/Speech-Enhancement-Pytorch) has NO native / FFI layer for this path:
 * its "operator API" is four Python call signatures.  Each entry point below
 * names the reference interface it replaces; INTEGRATION.md shows the ctypes stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (unless stated), on the device that
 *     is current when the call is made; `stream` is a cudaStream_t passed as void*;
 *   - calls are asynchronous on `stream`, re-entrant, and keep no global scratch (constant
 *     tables per (device, n_fft, hop, win_length, scale) live behind a mutex-guarded cache);
 *   - inputs are never written; outputs are fully overwritten (unless `accumulate` != 0);
 *   - return value 0 = ok, negative = se_status; se_last_error() gives the message of the last
 *     failing call on the calling thread.  No exceptions cross this boundary;
 *   - `rows` = product of all leading dims (batch x [speakers x] channels), SURVEY.md 8 notation:
 *     N = nsample, n = n_fft, h = hop, F = n/2+1, T = nframe = 1 + N/h (centre=True).
 *   - supported: n_fft in {512,1024,2048}, hop in {n/4, n/2}, 2 <= win_length <= n_fft,
 *     centre=True only.  Anything else returns SE_ERR_UNSUPPORTED (there is no fallback).
 */
#ifndef SE_B200_H
#define SE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum se_status {
    SE_OK = 0,
    SE_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, inconsistent shape */
    SE_ERR_UNSUPPORTED = -2,  /* n_fft / hop / win_length / mode outside the supported set */
    SE_ERR_CUDA = -3,         /* launch or runtime failure; message holds cudaGetErrorString */
    SE_ERR_ENVELOPE = -4      /* window overlap-add envelope ~ 0 (torch.istft raises the same) */
} se_status;

typedef enum se_mask_mode {   /* SURVEY.md 8a row a5 */
    SE_MASK_REAL = 0,         /* Y = X * m           unet.py:62 dnn.py:140 stft_rnn.py:108 crn.py:139 */
    SE_MASK_E = 1,            /* polar               dcunet.py:136-155 dccrn.py:203-217 */
    SE_MASK_C = 2,            /* complex multiply    dcunet.py:156-157 dccrn.py:218-219 */
    SE_MASK_R = 3             /* re*mr, im*mi        dcunet.py:158-159 dccrn.py:220-221 */
} se_mask_mode;

int se_version(void);
const char* se_last_error(void);

/* ---- geometries.  The tuned engine is compiled for n_fft 512 / 1024 / 2048 at hop n_fft/4 or n_fft/2 (and DCCRN's
 * ConvSTFT / ConviSTFT 400/100/512); every other geometry the reference's signatures accept -- any even n_fft in
 * 8 .. 8192 (not a power of two: Bluestein), any 1 <= hop <= n_fft, any win_length <= n_fft; any win_len / win_inc /
 * even fft_len for the DCCRN transforms -- runs on the general path (csrc/se_generic.cuh) behind the SAME entry points: se_stft_fwd / se_stft_bwd /
 * se_istft_fwd / se_istft_bwd and se_conv_stft_fwd_w / se_conv_istft_fwd_w / se_conv_istft_bwd_w.  The fused entry
 * points (enhance, mask_istft, conv_mask_istft, the losses, the segment transforms) exist for tuned geometries only;
 * the host side composes them from the plain transforms otherwise.  These two report which case applies (1 = tuned). */
int se_geometry_tuned(int n_fft, int hop);
/* config.center = False (src/evaluate.py:116): torch.stft without padding, frame t = x[t hop : t hop + n_fft],
 * T = 1 + (N - n_fft) / hop; x [rows,N] -> spec [rows,F,T,2], and its adjoint.  General-geometry kernels for every size.
 * (istft_custom with center=False raises in the reference -- a Hann window's overlap-add envelope is zero at the first
 * sample -- and raises here.) */
int se_stft_nocenter_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                         float scale, void* stream);
int se_stft_nocenter_bwd(const float* gspec, float* gx, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                         float scale, int accumulate, void* stream);
int se_conv_geometry_tuned(int win_len, int win_inc, int fft_len);

/* ---- replaces torch.stft inside stft_custom(tensor, config), src/evaluate.py:101-128 --------
 * x [rows,N] -> spec [rows,F,T,2] = scale * rfft(hann(win_length) * reflect-padded frames).
 * The reference uses scale = 1/win_length (src/evaluate.py:120). */
int se_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int n_fft, int hop,
                int win_length, float scale, void* stream);

/* ---- magnitude features feeding the NN bodies, the reference's quirks kept:
 * kind 0 power |re^2+im^2| (src/model/unet.py:40), 1 magnitude sqrt(re^2+im^2) (dnn.py:98),
 * 2 amplitude |re^2-im^2| (dcunet.py:379, stft_rnn.py:119, mel_rnn.py:123), 3 crn sqrt(re^2-im^2)
 * (crn.py:101; NaN where |im| > |re|, like the reference).
 * se_magnitude_feature: spec [count,2] -> feat [count].
 * se_stft_feature_fwd: se_stft_fwd that also writes feat [rows,F,T] from the bins in registers. */
int se_magnitude_feature(const float* spec, float* feat, int64_t count, int kind, void* stream);
int se_stft_feature_fwd(const float* x, float* spec, float* feat, int64_t rows, int64_t nsample, int n_fft,
                        int hop, int win_length, float scale, int kind, void* stream);

/* ---- evaluate()'s "segment then STFT" (src/evaluate.py:29-39,164-183) without materialising the
 * overlapping segments: x [nclip, clip_stride] holds clips of clip_len valid samples; segment s of
 * clip c is x[c, s*seg_stride : s*seg_stride + nsample] (zero beyond clip_len, like the reference's
 * zero-filled pad) and is reflect-padded on its own.  spec [(nseg*nclip), F, T, 2], row = s*nclip + c
 * (the reference's reshape(num_segment*nbatch, nchannel, ...) order). */
int se_stft_segments_fwd(const float* x, float* spec, int64_t nseg, int64_t nclip, int64_t clip_len,
                         int64_t clip_stride, int64_t seg_stride, int64_t nsample, int n_fft, int hop,
                         int win_length, float scale, void* stream);

/* ---- evaluate()'s normalisation and stitch folded into the transforms (src/evaluate.py:18-21, 84-96).
 * se_row_stats: x [rows, row_stride] (len valid samples) -> stats [rows,4] = (mean, 1/(std+1e-9), std+1e-9, 0) with
 *   torch.mean / torch.std (unbiased) semantics, double accumulators.
 * se_stft_segments_norm_fwd: se_stft_segments_fwd with the z-score (x - mean) / (std + 1e-9) applied to the valid samples
 *   while they are staged (the zero-filled tail stays zero, like the reference's pad after normalisation).  Clip c reads
 *   statistics row (c / stats_div) * stats_c + c % stats_c; stats == NULL: no normalisation.
 * se_istft_stitch_fwd: spec [(nseg*nclip), F, T, 2] (row = s*nclip + c) -> out [nclip, out_stride]: istft_custom of every
 *   segment, then `enhanced[..., :num_feature] = output[0]` and the last `stride` samples of each later segment appended
 *   (:84-88), trimmed to out_len (:90) and de-normalised y * (std+1e-9) + mean (:92-93) -- in one launch that synthesises
 *   only the frames overlapping the kept samples (stride/hop blocks + halo per later segment instead of all T frames). */
int se_row_stats(const float* x, float* stats, int64_t rows, int64_t len, int64_t row_stride, void* stream);
int se_stft_segments_norm_fwd(const float* x, float* spec, const float* stats, int64_t stats_div, int64_t stats_c, int64_t nseg,
                              int64_t nclip, int64_t clip_len, int64_t clip_stride, int64_t seg_stride, int64_t nsample, int n_fft,
                              int hop, int win_length, float scale, void* stream);
/* se_stft_segments_shared_fwd: the same result as se_stft_segments_norm_fwd, computed with the frames that overlapping
 * segments SHARE transformed once: segment s, frame t is frame s * (seg_stride / hop) + t of one clip-level transform
 * whenever it does not touch the segment's reflect padding.  One clip-level launch into `scratch`
 * (se_stft_segments_scratch_bytes; 0 = not applicable: seg_stride must be a multiple of hop), one launch for the groups
 * of 16 frames that hold boundary frames (2 of T/16 per segment), one gather launch for the interior frames. */
int64_t se_stft_segments_scratch_bytes(int64_t nseg, int64_t nclip, int64_t seg_stride, int64_t nsample, int n_fft, int hop);
int se_stft_segments_shared_fwd(const float* x, float* spec, const float* stats, int64_t stats_div, int64_t stats_c, int64_t nseg,
                                int64_t nclip, int64_t clip_len, int64_t clip_stride, int64_t seg_stride, int64_t nsample, int n_fft,
                                int hop, int win_length, float scale, void* scratch, void* stream);
int se_istft_stitch_fwd(const float* spec, float* out, const float* stats, int64_t stats_div, int64_t stats_c, int64_t nseg,
                        int64_t nclip, int64_t nframe, int64_t num_feature, int64_t stride, int64_t out_len, int64_t out_stride,
                        int n_fft, int hop, int win_length, float scale, void* stream);

/* adjoint of se_stft_fwd (autograd of src/evaluate.py:109-120):
 * gspec [rows,F,T,2] (dL/dRe, dL/dIm) -> gx [rows,N]; accumulate != 0 adds into gx. */
int se_stft_bwd(const float* gspec, float* gx, int64_t rows, int64_t nsample, int n_fft, int hop,
                int win_length, float scale, int accumulate, void* stream);

/* ---- replaces torch.istft inside istft_custom(tensor, length, config), src/evaluate.py:130-162
 * spec [rows,F,T,2] -> y [rows,length] = OLA(hann * irfft(scale * spec)) / OLA(hann^2),
 * trimmed by n/2, zero-extended past the natural end.  Reference scale = win_length (:131). */
int se_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t length, int n_fft,
                 int hop, int win_length, float scale, void* stream);

/* adjoint of se_istft_fwd: gy [rows,length] -> gspec [rows,F,T,2] */
int se_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t length, int n_fft,
                 int hop, int win_length, float scale, void* stream);

/* ---- mask application (model forward tails) -----------------------------------
 * spec/out [count,2]; mask [count] (REAL) or [count,2] (E/C/R).  pre_tanh: mask = tanh(raw)
 * first (src/model/dcunet.py:131).  count = rows*F*T. */
int se_mask_fwd(const float* spec, const float* mask, float* out, int64_t count, int mode, int pre_tanh,
                void* stream);
/* gout [count,2] -> gmask (same shape as mask, gradient wrt the RAW mask) and, if gspec != NULL,
 * gspec [count,2]. */
int se_mask_bwd(const float* spec, const float* mask, const float* gout, float* gmask, float* gspec,
                int64_t count, int mode, int pre_tanh, void* stream);

/* ---- the same masks in DCCRN's layout (src/model/dccrn.py:147-223: real = specs[:, :F], imag = specs[:, F:],
 * masks padded at the DC bin to [B,F,T], out_spec = cat([real, imag], 1)): spec / out / gout / gspec are planar
 * [rows, 2*nbin, nframe], the masks and their gradients two planes [rows, nbin, nframe].  mode = SE_MASK_E / C / R. */
int se_mask_planar_fwd(const float* spec, const float* mask_re, const float* mask_im, float* out, int64_t rows,
                       int64_t nbin, int64_t nframe, int mode, void* stream);
int se_mask_planar_bwd(const float* spec, const float* mask_re, const float* mask_im, const float* gout,
                       float* gmask_re, float* gmask_im, float* gspec, int64_t rows, int64_t nbin, int64_t nframe,
                       int mode, void* stream);

/* ---- multi-resolution STFT loss (new component; calling convention loss_function(enhanced,
 * sources), src/solver.py:480; definition SURVEY.md 8c).  Resolutions are fixed:
 * (512,128,512) (1024,256,1024) (2048,512,2048).
 *   fwd   : est, ref [rows,N] -> sums[9] (device, double): per resolution
 *           sum (b-a)^2, sum b^2, sum |log b - log a|.  Deterministic two-stage reduction.
 *           workspace: se_mrstft_workspace_bytes(rows, nsample) bytes of device scratch; the
 *           forward pass also leaves the clamped reference magnitudes |B| there, so the backward
 *           pass transforms only the estimate.  Keep it alive (unmodified) until bwd has run.
 *           With SE_MRSTFT_SAVE_SPECTRUM=1 in the environment (read once per process) the forward
 *           pass also saves the estimate's spectrum (3x the workspace) and the backward pass runs
 *           one transform per resolution instead of two; same results, same signatures.
 *   value : sums (after the caller all-reduced them across ranks) -> loss (device float).
 *           global_rows = rows summed over ranks (sets the mean's denominator).
 *   bwd   : g_est [rows,N] = gout * dloss/dest, gout a DEVICE scalar (upstream gradient);
 *           workspace = the buffer the matching fwd call filled. */
int64_t se_mrstft_workspace_bytes(int64_t rows, int64_t nsample);
int se_mrstft_loss_fwd(const float* est, const float* ref, int64_t rows, int64_t nsample, double* sums,
                       void* workspace, void* stream);
int se_mrstft_loss_value(const double* sums, int64_t global_rows, int64_t nsample, float* loss, void* stream);
/* Single process (no exchange step): forward + value in the same launches -- the deterministic reduction also writes
 * the loss, so the step has one launch less than se_mrstft_loss_fwd followed by se_mrstft_loss_value(rows). */
int se_mrstft_loss_fwd_value(const float* est, const float* ref, int64_t rows, int64_t nsample, double* sums, float* loss,
                             void* workspace, void* stream);
/* Uneven shards (ranks holding different row counts): keep the count on the device instead of guessing it on the host.
 * sums10 = the 9 sums + [9] the row count; each rank stores its own count there before the exchange (all-reduce or
 * se_mrstft_exchange_rows_value), which leaves the GLOBAL count in [9].  se_mrstft_loss_value_dev reads it from there;
 * se_mrstft_loss_bwd does too when called with global_rows == 0 (its `sums` must then hold the 10 doubles). */
int se_mrstft_loss_value_dev(const double* sums10, int64_t nsample, float* loss, void* stream);
int se_mrstft_loss_bwd(const float* est, const void* workspace, const double* sums, const float* gout,
                       int64_t global_rows, int64_t rows, int64_t nsample, float* g_est, void* stream);

/* ---- the exchange step of the utterance-sharded MR-STFT loss over NVLink peer memory (SURVEY.md 8e;
 * the reference has no counterpart: its multi-GPU path is nn.DataParallel, src/solver.py:144-145).
 * One process per GPU.  se_p2p_create allocates this rank's exchange buffer on the current device and
 * returns its 64-byte cudaIpc handle; the host side gathers the handles (torch.distributed) and maps
 * every peer's buffer with se_p2p_open.  se_mrstft_exchange_value then replaces "ncclAllReduce(sums)
 * followed by se_mrstft_loss_value" with ONE single-CTA kernel on `stream`: it stores this rank's 9
 * sums into every peer's buffer, waits (on the device) until all ranks' sums have arrived, adds them in
 * rank order -- every rank gets the same bits -- and writes the global sums in place and, if loss !=
 * NULL, the loss.  bufs is a HOST array of `world` device pointers, bufs[rank] the local buffer.
 * Every rank of the group must make the call the same number of times.  A peer that does not show up within
 * SE_P2P_SPIN_SECONDS (environment, read once; default 600, 0 = wait forever like NCCL would) does not hang the
 * stream and does not kill the CUDA context: the kernel reports on stdout and poisons sums and loss with NaN.
 * se_mrstft_exchange_rows_value is the same step on 10 doubles (sums10[9] = row count in, global row count out). */
int se_p2p_create(void** local, unsigned char* handle64);
int se_p2p_open(const unsigned char* handle64, void** peer);
int se_p2p_close(void* peer);
int se_p2p_destroy(void* local);
int se_mrstft_exchange_value(double* sums, void* const* bufs, int world, int rank, int64_t global_rows,
                             int64_t nsample, float* loss, void* stream);
int se_mrstft_exchange_rows_value(double* sums10, void* const* bufs, int world, int rank, int64_t nsample, float* loss,
                                  void* stream);

/* ---- STFT-domain training losses against a WAVEFORM target: what
 * loss_function(enhanced, stft_custom(sources)) computes with torch's mse_loss / l1_loss on
 * [B,C,F,T,2] (src/solver.py:457-458,480; src/distrib.py:263-267), without materialising the target
 * spectrum.  kind 0 = mse, 1 = l1.  fwd: enh [rows,F,T,2], target [rows,N] -> *sum_out (device double)
 * = sum of squared / absolute differences over this rank's rows (the mean is sum / (global_rows*F*T*2);
 * all-reduce the sums across ranks first).  bwd: genh [rows,F,T,2] = gout * d mean / d enh. */
int64_t se_spectral_loss_workspace_bytes(int64_t rows, int64_t nsample, int hop);
int se_spectral_loss_fwd(const float* enh, const float* target, int64_t rows, int64_t nsample, int n_fft, int hop,
                         int win_length, float scale, int kind, double* sum_out, void* workspace, void* stream);
int se_spectral_loss_bwd(const float* enh, const float* target, const float* gout, int64_t global_rows, int64_t rows,
                         int64_t nsample, int n_fft, int hop, int win_length, float scale, int kind, float* genh,
                         void* stream);

/* ---- phase-sensitive spectral approximation, loss_phase_sensitive_spectral_approximation(enhance,
 * target, mixture), src/loss.py:32-56 (the `psa` key of src/distrib.py:271-272): three spectra [count,2].
 * fwd -> *sum_out (device double) = sum of squared residuals (mean = sum / global_count; all-reduce first
 * when sharded); bwd: genh [count,2] = gout * d mean / d enhance. */
int64_t se_psa_workspace_bytes(int64_t count);
int se_psa_loss_fwd(const float* enh, const float* tgt, const float* mix, int64_t count, double* sum_out,
                    void* workspace, void* stream);
int se_psa_loss_bwd(const float* enh, const float* tgt, const float* mix, const float* gout, int64_t global_count,
                    int64_t count, float* genh, void* stream);

/* ---- time-domain SI-SNR: si_snr(s1, s2, eps=1e-8) / loss_sisdr, src/loss.py:14-29.
 * fwd: s1, s2 [rows,N] -> dots [rows,3] (device double: <s1,s1>, <s1,s2>, <s2,s2>) and snr [rows]
 * (10 log10(|s_target|^2 / (|e|^2 + eps) + eps)); the caller takes the mean (and the sign for loss_sisdr).
 * bwd: g [rows,N] = gscale * gout * d snr_row / d s1 (gout a device scalar; gscale = -+1/rows). */
int se_sisnr_fwd(const float* s1, const float* s2, int64_t rows, int64_t nsample, double eps, double* dots,
                 float* snr, void* stream);
int se_sisnr_bwd(const float* s1, const float* s2, const double* dots, const float* gout, float gscale,
                 int64_t rows, int64_t nsample, double eps, float* g, void* stream);

/* ---- fused wave -> STFT -> mask -> iSTFT -> wave (stft_custom + model tail + istft_custom in
 * one launch; SURVEY.md 8b "se_enhance_fwd/bwd").  mask [rows,F,T] (REAL) or [rows,F,T,2].
 * The 1/win_length and win_length scales of the reference cancel and are not applied. */
int se_enhance_fwd(const float* x, const float* mask, float* y, int64_t rows, int64_t nsample, int n_fft,
                   int hop, int win_length, int mode, int pre_tanh, void* stream);
/* gy [rows,N], x, mask -> gmask (gradient wrt the raw mask) */
int se_enhance_bwd(const float* gy, const float* x, const float* mask, float* gmask, int64_t rows,
                   int64_t nsample, int n_fft, int hop, int win_length, int mode, int pre_tanh, void* stream);

/* ---- model tail + iSTFT in one launch: y = istft_custom(apply_mask(spec, mask), length)
 * (the mask tails listed at se_mask_fwd followed by src/evaluate.py:130-153 as evaluate() and
 * Solver._run_one_epoch chain them, src/evaluate.py:54-72, src/solver.py:466-480) without writing the
 * masked spectrum.  spec [rows,F,T,2], mask [rows,F,T] (REAL) or [rows,F,T,2], y [rows,length];
 * scale as se_istft_fwd (win_length).  bwd: gy [rows,length] -> gmask (gradient wrt the raw mask;
 * spec is treated as a constant, as it is when it comes from stft_custom(mixture)). */
int se_mask_istft_fwd(const float* spec, const float* mask, float* y, int64_t rows, int64_t nframe,
                      int64_t length, int n_fft, int hop, int win_length, float scale, int mode, int pre_tanh,
                      void* stream);
int se_mask_istft_bwd(const float* gy, const float* spec, const float* mask, float* gmask, int64_t rows,
                      int64_t nframe, int64_t length, int n_fft, int hop, int win_length, float scale, int mode,
                      int pre_tanh, void* stream);

/* ---- Conv-TasNet decoder tail: overlap_and_add(signal, frame_step), src/model/conv_tasnet.py:11-31
 * (called at :203 with frame_step = L/2).  signal [rows, frames, frame_length] ->
 * out [rows, frame_step*(frames-1) + frame_length]; contributions are summed in increasing frame order
 * (the order the reference's index_add_ applies them), deterministic, no atomics.
 * bwd: gout [rows, out_len] -> gsignal [rows, frames, frame_length]. */
int se_overlap_add_fwd(const float* signal, float* out, int64_t rows, int64_t frames, int frame_length,
                       int frame_step, void* stream);
int se_overlap_add_bwd(const float* gout, float* gsignal, int64_t rows, int64_t frames, int frame_length,
                       int frame_step, void* stream);

/* ---- DCCRN in-model transforms: ConvSTFT.forward / ConviSTFT.forward, src/model/dccrn.py:687-747
 * x [rows,N] -> spec [rows, 2F, T], T = (N + 2(win_len-win_inc) - win_len)/win_inc + 1, Hann
 * window, zero padding, frame zero-extended at the END to fft_len.  Supported: fft_len 512,
 * win_len <= fft_len, even win_inc. */
/* Any window instead of Hann (the reference takes a scipy.signal.get_window type, dccrn.py:651-655): the host computes
 * the window values once, se_register_window returns an id > 0 for them (identical values share an id), and the *_w
 * variants of the transforms take it (0 = the built-in periodic Hann). */
int se_register_window(const double* values, int win_len);
int se_conv_stft_fwd_w(const float* x, float* spec, int64_t rows, int64_t nsample, int win_len, int win_inc, int fft_len,
                       int window_id, void* stream);
int se_conv_istft_fwd_w(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc,
                        int fft_len, int window_id, void* stream);
int se_conv_istft_bwd_w(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc,
                        int fft_len, int window_id, void* stream);
int se_conv_mask_istft_fwd_w(const float* spec, const float* mask_re, const float* mask_im, float* y, int64_t rows,
                             int64_t nframe, int64_t out_len, int win_len, int win_inc, int fft_len, int mode, int window_id,
                             void* stream);
int se_conv_mask_istft_bwd_w(const float* gy, const float* spec, const float* mask_re, const float* mask_im, float* gmask_re,
                             float* gmask_im, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc,
                             int fft_len, int mode, int window_id, void* stream);
/* ConvSTFT(feature_type='real') epilogue (mags, phase = sqrt(re^2+im^2), atan2(im, re), dccrn.py:696-701) and
 * ConviSTFT(inputs, phase) prologue (cat([mags cos, mags sin], 1), :729-732) on the planar [rows, 2*nbin, nframe] layout,
 * one launch each; the last is the gradient of the prologue wrt (mags, phase). */
int se_polar_from_planar(const float* spec, float* mags, float* phase, int64_t rows, int64_t nbin, int64_t nframe, void* stream);
int se_planar_from_polar(const float* mags, const float* phase, float* spec, int64_t rows, int64_t nbin, int64_t nframe, void* stream);
int se_planar_from_polar_bwd(const float* mags, const float* phase, const float* gspec, float* gmags, float* gphase, int64_t rows,
                             int64_t nbin, int64_t nframe, void* stream);
int se_conv_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int win_len, int win_inc,
                     int fft_len, void* stream);
/* spec [rows,2F,T] -> y [rows,out_len]; out_len = length if length > 0 else natural
 * (win_inc*(T-1) + win_len - 2(win_len-win_inc)).  Least-squares (pinv) frame inverse,
 * window^2 overlap-add normalisation with the reference's +1e-8 (dccrn.py:733-745). */
int se_conv_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t out_len,
                      int win_len, int win_inc, int fft_len, void* stream);
int se_conv_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t out_len,
                      int win_len, int win_inc, int fft_len, void* stream);

/* ---- DCCRN's model tail + ConviSTFT in one launch each way (src/model/dccrn.py:203-224: mask, cat, self.istft):
 * spec [rows,2F,T] is the UNMASKED spectrum, mask_re / mask_im [rows,F,T] (padded at DC by the model, dccrn.py:200-201),
 * mode = SE_MASK_E / C / R.  fwd -> y [rows,out_len]; bwd: gy -> gradients wrt the two mask planes.  The masked
 * spectrum and its gradient never touch memory. */
int se_conv_mask_istft_fwd(const float* spec, const float* mask_re, const float* mask_im, float* y, int64_t rows,
                           int64_t nframe, int64_t out_len, int win_len, int win_inc, int fft_len, int mode,
                           void* stream);
int se_conv_mask_istft_bwd(const float* gy, const float* spec, const float* mask_re, const float* mask_im,
                           float* gmask_re, float* gmask_im, int64_t rows, int64_t nframe, int64_t out_len,
                           int win_len, int win_inc, int fft_len, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SE_B200_H */
