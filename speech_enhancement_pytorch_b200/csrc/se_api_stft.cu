// se_api_stft.cu -- STFT / iSTFT entry points (k_analysis / k_synthesis instantiations).
#include "se_host.h"

using namespace se;

// planning happens here, where the geometry (frames per group, CTAs per SM) is known
template <class G, int LMODE>
static cudaError_t run_analysis(AnaArgs a, int64_t rows, cudaStream_t st) {
    plan_analysis(rows, a.nframe, a.gpc, a.nchunks, G::FR);
    return launch(k_analysis<G, LMODE, false>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS, st, a);
}
template <class G, int EMODE>
static cudaError_t run_synthesis(SynArgs a, int64_t rows, cudaStream_t st) {
    a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, G::OLA, G::MINB, G::FR);
    return launch(k_synthesis<G, EMODE>, (unsigned)(rows * a.nchunks), G::NT,
                  EMODE == EMIT_ADJ ? Smem<G>::SYNTH_ADJ : Smem<G>::SYNTH_ISTFT, st, a);
}

extern "C" {

int se_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                float scale, void* stream) {
    if (!x || !spec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    AnaArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.in = x; a.out = spec; a.in_stride = nsample; a.seg_rows = 1; a.nsample = (int)nsample; a.in_len = (int)nsample;
    a.nframe = (int)(1 + nsample / hop); a.pad = 0; a.edge_scale = 1.0f;
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_analysis<G, LOAD_REFLECT>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_fwd launch");
}

int se_stft_feature_fwd(const float* x, float* spec, float* feat, int64_t rows, int64_t nsample, int n_fft, int hop,
                        int win_length, float scale, int kind, void* stream) {
    if (!x || !spec || !feat) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (kind < 0 || kind > 3) return fail(SE_ERR_UNSUPPORTED, "feature kind must be 0..3 (power, magnitude, amplitude, crn)");
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    AnaArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.in = x; a.out = spec; a.feat = feat; a.feat_kind = kind; a.in_stride = nsample; a.seg_rows = 1;
    a.nsample = (int)nsample; a.in_len = (int)nsample;
    a.nframe = (int)(1 + nsample / hop); a.pad = 0; a.edge_scale = 1.0f;
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_analysis<G, LOAD_REFLECT>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_feature_fwd launch");
}

int se_magnitude_feature(const float* spec, float* feat, int64_t count, int kind, void* stream) {
    if (!spec || !feat || count <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    if (kind < 0 || kind > 3) return fail(SE_ERR_UNSUPPORTED, "feature kind must be 0..3 (power, magnitude, amplitude, crn)");
    int64_t blocks = (count / 2 + 255) / 256 + 1;
    if (blocks > 148 * 12) blocks = 148 * 12;
    cudaError_t e = launch(k_feature, (unsigned)blocks, 256, 0, (cudaStream_t)stream,
                           reinterpret_cast<const float2*>(spec), feat, count, kind);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_magnitude_feature launch");
}

int se_stft_segments_fwd(const float* x, float* spec, int64_t nseg, int64_t nclip, int64_t clip_len, int64_t clip_stride,
                         int64_t seg_stride, int64_t nsample, int n_fft, int hop, int win_length, float scale, void* stream) {
    if (!x || !spec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nseg <= 0 || nclip <= 0 || clip_len <= 0 || seg_stride <= 0 || clip_stride < clip_len)
        return fail(SE_ERR_BAD_ARG, "bad segment geometry");
    const int64_t rows = nseg * nclip;
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    if ((nseg - 1) * seg_stride >= clip_len) return fail(SE_ERR_BAD_ARG, "last segment starts beyond the clip");
    AnaArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.in = x; a.out = spec; a.in_stride = seg_stride; a.clip_stride = clip_stride; a.seg_rows = (int)nclip;
    a.clip_len = (int)clip_len; a.nsample = (int)nsample; a.in_len = (int)nsample;
    a.nframe = (int)(1 + nsample / hop); a.pad = 0; a.edge_scale = 1.0f;
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_analysis<G, LOAD_REFLECT>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_segments_fwd launch");
}

int se_stft_bwd(const float* gspec, float* gx, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                float scale, int accumulate, void* stream) {
    if (!gspec || !gx) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample < n_fft) return fail(SE_ERR_UNSUPPORTED, "adjoint needs nsample >= n_fft");
    SynArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.in = gspec; a.out = gx; a.nsample = (int)nsample; a.out_len = (int)nsample;
    a.nframe = (int)(1 + nsample / hop);
    a.b_lo = 0; a.b_hi = (int)((nsample + n_fft + hop - 1) / hop);
    a.accumulate = accumulate; a.edge_scale = 2.0f;
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_synthesis<G, EMIT_ADJ>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_bwd launch");
}

int se_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop,
                 int win_length, float scale, void* stream) {
    if (!spec || !y) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nframe <= 0 || length <= 0) return fail(SE_ERR_BAD_ARG, "nframe and length must be positive");
    if (int rc = check_common(rows, length, n_fft, hop, win_length)) return rc;
    if (!envelope_ok(n_fft, hop, win_length, false, nframe, n_fft / 2, n_fft / 2 + length, 1e-11))
        return fail(SE_ERR_ENVELOPE, "window overlap add min < 1e-11 (torch.istft raises the same)");
    SynArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, scale / (float)n_fft, a.tb)) return rc;
    a.in = spec; a.out = y; a.nsample = (int)(n_fft + hop * (nframe - 1)); a.out_len = (int)length;
    a.nframe = (int)nframe;
    a.b_lo = (n_fft / 2) / hop; a.b_hi = (int)((n_fft / 2 + length + hop - 1) / hop);
    a.accumulate = 0; a.edge_scale = 1.0f;
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_synthesis<G, EMIT_ISTFT>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_istft_fwd launch");
}

int se_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop,
                 int win_length, float scale, void* stream) {
    if (!gy || !gspec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nframe <= 0 || length <= 0) return fail(SE_ERR_BAD_ARG, "nframe and length must be positive");
    if (int rc = check_common(rows, length, n_fft, hop, win_length)) return rc;
    AnaArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, scale / (float)n_fft, a.tb)) return rc;
    a.in = gy; a.out = gspec; a.in_stride = length; a.seg_rows = 1; a.in_len = (int)length;
    a.nsample = (int)(n_fft + hop * (nframe - 1)); a.nframe = (int)nframe; a.pad = 0; a.edge_scale = 0.5f;
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_analysis<G, LOAD_ENV>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_istft_bwd launch");
}

}  // extern "C"
