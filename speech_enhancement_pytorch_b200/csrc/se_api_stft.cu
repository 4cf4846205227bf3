// se_api_stft.cu -- STFT / iSTFT entry points (k_analysis / k_synthesis instantiations).
#include "se_host.h"

using namespace se;

// planning happens here, where the geometry (frames per group, CTAs per SM) is known
template <class G, int LMODE, bool NORM = false>
static cudaError_t run_analysis(AnaArgs a, int64_t rows, cudaStream_t st) {
    plan_analysis(rows, a.nframe, a.gpc, a.nchunks, G::FR);
    return launch(k_analysis<G, LMODE, false, NORM>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS, st, a);
}
template <class G, int EMODE>
static cudaError_t run_synthesis(SynArgs a, int64_t rows, cudaStream_t st) {
    a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, G::OLA, G::MINB, G::FR, (EMODE == EMIT_ADJ && G::FR == 8) ? 2 : 1, G::FR == 8 ? 16 : 8);
    return launch(k_synthesis<G, EMODE>, (unsigned)(rows * a.nchunks), G::NT,
                  EMODE == EMIT_ADJ ? Smem<G>::SYNTH_ADJ : Smem<G>::SYNTH_ISTFT, st, a);
}

// ---- signal-pair engine (two rows per CTA).  Tables are staged doubled when MINB CTAs still fit, undoubled
// otherwise; geometries that fit neither way stay on the scalar engine (returns false).
template <class G> constexpr bool fits2(size_t bytes) { return bytes <= 232448 && G::MINB * (bytes + 1024) <= 233472; }

template <int N, int HOP, int NT, int LMODE>
static bool run_analysis2(AnaArgs a, int64_t rows, cudaStream_t st, cudaError_t& e) {
    using GD = Geo2<N, HOP, NT, true>;
    using GN = Geo2<N, HOP, NT, false>;
    const int64_t prows = (rows + 1) / 2;
    plan_analysis(prows, a.nframe, a.gpc, a.nchunks, 8);
    if constexpr (fits2<GD>(Smem2<GD>::ANALYSIS)) {
        e = launch(k_analysis2<GD, LMODE>, (unsigned)(prows * a.nchunks), NT, Smem2<GD>::ANALYSIS, st, a, (int)rows);
        return true;
    } else if constexpr (Smem2<GN>::ANALYSIS <= 232448) {        // undoubled tables, possibly one CTA fewer per SM
        e = launch(k_analysis2<GN, LMODE>, (unsigned)(prows * a.nchunks), NT, Smem2<GN>::ANALYSIS, st, a, (int)rows);
        return true;
    }
    return false;
}
template <int N, int HOP, int NT, int EMODE>
static bool run_synthesis2(SynArgs a, int64_t rows, cudaStream_t st, cudaError_t& e) {
    using GD = Geo2<N, HOP, NT, true>;
    using GN = Geo2<N, HOP, NT, false>;
    const int64_t prows = (rows + 1) / 2;
    constexpr bool ADJ = EMODE == EMIT_ADJ;
    a.nchunks = plan_synthesis(prows, a.b_hi - a.b_lo, GD::OLA, GD::MINB, 8, ADJ ? 2 : 1, 16);
    constexpr size_t BD = ADJ ? Smem2<GD>::SYNTH_ADJ : Smem2<GD>::SYNTH_ISTFT;
    constexpr size_t BN = ADJ ? Smem2<GN>::SYNTH_ADJ : Smem2<GN>::SYNTH_ISTFT;
    if constexpr (fits2<GD>(BD)) {
        e = launch(k_synthesis2<GD, EMODE>, (unsigned)(prows * a.nchunks), NT, BD, st, a, (int)rows);
        return true;
    } else if constexpr (BN <= 232448) {
        e = launch(k_synthesis2<GN, EMODE>, (unsigned)(prows * a.nchunks), NT, BN, st, a, (int)rows);
        return true;
    }
    return false;
}
// CALL sees the compile-time constants N2, HOP2, NT2
#define SE_DISPATCH_GEO2(n_fft, hop, CALL)                                                              \
    do {                                                                                                \
        if (n_fft == 512 && hop == 128) { constexpr int N2 = 512, HOP2 = 128, NT2 = 128; CALL; }        \
        else if (n_fft == 512 && hop == 256) { constexpr int N2 = 512, HOP2 = 256, NT2 = 128; CALL; }   \
        else if (n_fft == 1024 && hop == 256) { constexpr int N2 = 1024, HOP2 = 256, NT2 = 256; CALL; } \
        else if (n_fft == 1024 && hop == 512) { constexpr int N2 = 1024, HOP2 = 512, NT2 = 256; CALL; } \
        else if (n_fft == 2048 && hop == 512) { constexpr int N2 = 2048, HOP2 = 512, NT2 = 512; CALL; } \
        else { constexpr int N2 = 2048, HOP2 = 1024, NT2 = 512; CALL; }                                 \
    } while (0)

// ---- two-pass engine (se_fft3.cuh): n_fft 512 / 1024 at hop n/4
template <class G3, int LMODE>
static cudaError_t run_analysis3(AnaArgs a, int64_t rows, cudaStream_t st) {
    plan_analysis(rows, a.nframe, a.gpc, a.nchunks, G3::FR);
    return launch(k_analysis3<G3, LMODE>, (unsigned)(rows * a.nchunks), G3::NT, Smem<typename G3::Base>::ANALYSIS, st, a);
}
template <class G3, int EMODE>
static cudaError_t run_synthesis3(SynArgs a, int64_t rows, cudaStream_t st) {
    a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, G3::OLA, G3::MINB, G3::FR);
    return launch(k_synthesis3<G3, EMODE>, (unsigned)(rows * a.nchunks), G3::NT,
                  EMODE == EMIT_ADJ ? Smem<typename G3::Base>::SYNTH_ADJ : Smem<typename G3::Base>::SYNTH_ISTFT, st, a);
}

template <int LMODE>
static cudaError_t dispatch_analysis(const AnaArgs& a, int64_t rows, int n_fft, int hop, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    bool done = false;
    if (engine_version() == 3 && !a.norm && n_fft == 1024 && hop == 256) return run_analysis3<Geo3<1024, 256, 256>, LMODE>(a, rows, st);
    if (engine_version() == 3 && !a.norm && n_fft == 512 && hop == 128) return run_analysis3<Geo3<512, 128, 128>, LMODE>(a, rows, st);
    if (engine_version() == 2) SE_DISPATCH_GEO2(n_fft, hop, (done = run_analysis2<N2, HOP2, NT2, LMODE>(a, rows, st, e)));
    if (!done) SE_DISPATCH_GEO(n_fft, hop, (e = run_analysis<G, LMODE>(a, rows, st)));
    return e;
}
template <int EMODE>
static cudaError_t dispatch_synthesis(const SynArgs& a, int64_t rows, int n_fft, int hop, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    bool done = false;
    if (engine_version() == 3 && n_fft == 1024 && hop == 256) return run_synthesis3<Geo3<1024, 256, 256>, EMODE>(a, rows, st);
    if (engine_version() == 3 && n_fft == 512 && hop == 128) return run_synthesis3<Geo3<512, 128, 128>, EMODE>(a, rows, st);
    if (engine_version() == 2) SE_DISPATCH_GEO2(n_fft, hop, (done = run_synthesis2<N2, HOP2, NT2, EMODE>(a, rows, st, e)));
    if (!done) SE_DISPATCH_GEO(n_fft, hop, (e = run_synthesis<G, EMODE>(a, rows, st)));
    return e;
}

template <class G>
static cudaError_t run_stitch(SynArgs a, StitchArgs s, cudaStream_t st) {
    s.chunks0 = plan_synthesis(s.nclip, a.b_hi - a.b_lo, G::OLA, G::MINB, G::FR);
    const unsigned grid = (unsigned)(s.nclip * s.chunks0 + (s.nseg - 1) * s.nclip);
    return launch(k_synthesis_stitch<G>, grid, G::NT, Smem<G>::SYNTH_ISTFT, st, a, s);
}

extern "C" {

int se_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                float scale, void* stream) {
    if (!x || !spec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (!geometry_tuned(n_fft, hop)) return gen_stft_fwd(x, spec, rows, nsample, n_fft, hop, win_length, scale, (cudaStream_t)stream);
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    AnaArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.in = x; a.out = spec; a.in_stride = nsample; a.seg_rows = 1; a.nsample = (int)nsample; a.in_len = (int)nsample;
    a.nframe = (int)(1 + nsample / hop); a.pad = 0; a.edge_scale = 1.0f;
    const cudaError_t e = dispatch_analysis<LOAD_REFLECT>(a, rows, n_fft, hop, (cudaStream_t)stream);
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
    return se_stft_segments_norm_fwd(x, spec, nullptr, 1, 1, nseg, nclip, clip_len, clip_stride, seg_stride, nsample, n_fft, hop,
                                     win_length, scale, stream);
}

int se_row_stats(const float* x, float* stats, int64_t rows, int64_t len, int64_t row_stride, void* stream) {
    if (!x || !stats) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (rows <= 0 || len <= 0 || row_stride < len) return fail(SE_ERR_BAD_ARG, "need rows > 0, len > 0, row_stride >= len");
    cudaError_t e = launch(k_row_stats, (unsigned)(rows * kStatsCluster), 1024u, 0, (cudaStream_t)stream, x, reinterpret_cast<float4*>(stats), len, row_stride);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_row_stats launch");
}

int se_stft_segments_norm_fwd(const float* x, float* spec, const float* stats, int64_t stats_div, int64_t stats_c, int64_t nseg,
                              int64_t nclip, int64_t clip_len, int64_t clip_stride, int64_t seg_stride, int64_t nsample, int n_fft,
                              int hop, int win_length, float scale, void* stream) {
    if (!x || !spec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (stats && (stats_div <= 0 || stats_c <= 0)) return fail(SE_ERR_BAD_ARG, "bad statistics indexing");
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
    a.norm = reinterpret_cast<const float4*>(stats); a.norm_div = (int)stats_div; a.norm_c = (int)stats_c;
    a.nframe = (int)(1 + nsample / hop); a.pad = 0; a.edge_scale = 1.0f;
    cudaError_t e;
    if (stats) SE_DISPATCH_GEO(n_fft, hop, (e = run_analysis<G, LOAD_REFLECT, true>(a, rows, (cudaStream_t)stream)));   // scalar engine: it has the normalising fill
    else e = dispatch_analysis<LOAD_REFLECT>(a, rows, n_fft, hop, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_segments_fwd launch");
}

int se_stft_bwd(const float* gspec, float* gx, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                float scale, int accumulate, void* stream) {
    if (!gspec || !gx) return fail(SE_ERR_BAD_ARG, "null pointer");
    // the tuned adjoint needs nsample >= n_fft (its reflect fold stays inside the edge chunks); shorter rows take the general path
    if (!geometry_tuned(n_fft, hop) || nsample < n_fft)
        return gen_stft_bwd(gspec, gx, rows, nsample, n_fft, hop, win_length, scale, accumulate, (cudaStream_t)stream);
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    SynArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.in = gspec; a.out = gx; a.nsample = (int)nsample; a.out_len = (int)nsample;
    a.nframe = (int)(1 + nsample / hop);
    a.b_lo = 0; a.b_hi = (int)((nsample + n_fft + hop - 1) / hop);
    a.accumulate = accumulate; a.edge_scale = 2.0f;
    const cudaError_t e = dispatch_synthesis<EMIT_ADJ>(a, rows, n_fft, hop, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_bwd launch");
}

int se_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop,
                 int win_length, float scale, void* stream) {
    if (!spec || !y) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (!geometry_tuned(n_fft, hop)) return gen_istft_fwd(spec, y, rows, nframe, length, n_fft, hop, win_length, scale, (cudaStream_t)stream);
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
    const cudaError_t e = dispatch_synthesis<EMIT_ISTFT>(a, rows, n_fft, hop, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_istft_fwd launch");
}

int se_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop,
                 int win_length, float scale, void* stream) {
    if (!gy || !gspec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (!geometry_tuned(n_fft, hop)) return gen_istft_bwd(gy, gspec, rows, nframe, length, n_fft, hop, win_length, scale, (cudaStream_t)stream);
    if (nframe <= 0 || length <= 0) return fail(SE_ERR_BAD_ARG, "nframe and length must be positive");
    if (int rc = check_common(rows, length, n_fft, hop, win_length)) return rc;
    AnaArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, scale / (float)n_fft, a.tb)) return rc;
    a.in = gy; a.out = gspec; a.in_stride = length; a.seg_rows = 1; a.in_len = (int)length;
    a.nsample = (int)(n_fft + hop * (nframe - 1)); a.nframe = (int)nframe; a.pad = 0; a.edge_scale = 0.5f;
    const cudaError_t e = dispatch_analysis<LOAD_ENV>(a, rows, n_fft, hop, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_istft_bwd launch");
}

int se_istft_stitch_fwd(const float* spec, float* out, const float* stats, int64_t stats_div, int64_t stats_c, int64_t nseg,
                        int64_t nclip, int64_t nframe, int64_t num_feature, int64_t stride, int64_t out_len, int64_t out_stride,
                        int n_fft, int hop, int win_length, float scale, void* stream) {
    if (!spec || !out) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nseg <= 0 || nclip <= 0 || nframe <= 0 || num_feature <= 0 || stride <= 0 || stride > num_feature || out_len <= 0 || out_stride < out_len)
        return fail(SE_ERR_BAD_ARG, "bad stitch geometry");
    if (out_len > num_feature + stride * (nseg - 1)) return fail(SE_ERR_BAD_ARG, "out_len exceeds the stitched length");
    if (stats && (stats_div <= 0 || stats_c <= 0)) return fail(SE_ERR_BAD_ARG, "bad statistics indexing");
    if (int rc = check_common(nseg * nclip, num_feature, n_fft, hop, win_length)) return rc;
    if (!envelope_ok(n_fft, hop, win_length, false, nframe, n_fft / 2, n_fft / 2 + num_feature, 1e-11))
        return fail(SE_ERR_ENVELOPE, "window overlap add min < 1e-11 (torch.istft raises the same)");
    SynArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, scale / (float)n_fft, a.tb)) return rc;
    a.in = spec; a.out = out; a.nsample = (int)(n_fft + hop * (nframe - 1)); a.out_len = (int)num_feature;
    a.nframe = (int)nframe;
    a.b_lo = (n_fft / 2) / hop; a.b_hi = (int)((n_fft / 2 + num_feature + hop - 1) / hop);
    a.accumulate = 0; a.edge_scale = 1.0f;
    StitchArgs s{};
    s.nclip = (int)nclip; s.nseg = (int)nseg; s.stride = (int)stride; s.num_feature = (int)num_feature;
    s.clip_len = (int)out_len; s.out_stride = out_stride;
    s.tail_b_lo = (int)((n_fft / 2 + num_feature - stride) / hop); s.tail_b_hi = a.b_hi;
    s.norm = reinterpret_cast<const float4*>(stats); s.norm_div = (int)stats_div; s.norm_c = (int)stats_c;
    // a later segment's tail is ONE chunk: its blocks plus the OLA halo must fit the kernel's group walk (any count does)
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_stitch<G>(a, s, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_istft_stitch_fwd launch");
}

// ---- evaluate(): segment STFT with the frames shared between overlapping segments computed once
// scratch = the clip-level spectrum [nclip][F][Tg]; 0 when the shared path does not apply (seg_stride not a multiple of hop)
static int64_t shared_frames(int64_t nseg, int64_t seg_stride, int64_t nsample, int n_fft, int hop, int64_t& t_lo, int64_t& t_hi) {
    if (seg_stride % hop) return 0;
    t_lo = (n_fft / 2 + hop - 1) / hop;                      // first frame that does not touch the left reflect padding
    t_hi = (nsample - n_fft / 2) / hop;                      // last frame that does not touch the right one
    if (t_hi < t_lo) return 0;
    return (nseg - 1) * (seg_stride / hop) + t_hi + 1;       // clip-level frames needed
}
int64_t se_stft_segments_scratch_bytes(int64_t nseg, int64_t nclip, int64_t seg_stride, int64_t nsample, int n_fft, int hop) {
    int64_t t_lo, t_hi;
    const int64_t Tg = shared_frames(nseg, seg_stride, nsample, n_fft, hop, t_lo, t_hi);
    return Tg <= 0 ? 0 : nclip * (int64_t)(n_fft / 2 + 1) * Tg * (int64_t)sizeof(float2);
}

int se_stft_segments_shared_fwd(const float* x, float* spec, const float* stats, int64_t stats_div, int64_t stats_c, int64_t nseg,
                                int64_t nclip, int64_t clip_len, int64_t clip_stride, int64_t seg_stride, int64_t nsample, int n_fft,
                                int hop, int win_length, float scale, void* scratch, void* stream) {
    if (!x || !spec || !scratch) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nseg <= 0 || nclip <= 0 || clip_len <= 0 || seg_stride <= 0 || clip_stride < clip_len) return fail(SE_ERR_BAD_ARG, "bad segment geometry");
    if (stats && (stats_div <= 0 || stats_c <= 0)) return fail(SE_ERR_BAD_ARG, "bad statistics indexing");
    const int64_t rows = nseg * nclip;
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    if ((nseg - 1) * seg_stride >= clip_len) return fail(SE_ERR_BAD_ARG, "last segment starts beyond the clip");
    int64_t t_lo, t_hi;
    const int64_t Tg = shared_frames(nseg, seg_stride, nsample, n_fft, hop, t_lo, t_hi);
    if (Tg <= 0) return fail(SE_ERR_UNSUPPORTED, "shared-frame segment STFT needs seg_stride to be a multiple of hop");
    const int64_t T = 1 + nsample / hop;
    const int FRG = (frames8() && ((n_fft == 512 && hop == 128) || (n_fft == 1024 && hop == 256))) ? 8 : 16;    // = G::FR of SE_DISPATCH_GEO
    const int64_t ng = (T + FRG - 1) / FRG;
    const int64_t lead = (t_lo + FRG - 1) / FRG;             // leading groups holding boundary frames
    const int64_t trail0 = (t_hi + 1) / FRG;                 // first trailing group holding a boundary frame
    if (lead >= trail0) {                                    // short segments: every group is an edge group
        return se_stft_segments_norm_fwd(x, spec, stats, stats_div, stats_c, nseg, nclip, clip_len, clip_stride, seg_stride, nsample,
                                         n_fft, hop, win_length, scale, stream);
    }
    cudaError_t e;
    // 1. clip-level transform: frame g covers clip samples [g hop - n/2, g hop + n/2), zero beyond the clip
    {
        AnaArgs a{};
        if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
        // rows of this launch are clips: seg_rows = nclip makes (seg, clip) = (0, row), which also indexes the statistics
        a.in = x; a.out = reinterpret_cast<float*>(scratch); a.in_stride = 0; a.clip_stride = clip_stride; a.seg_rows = (int)nclip;
        a.nsample = (int)clip_len; a.in_len = (int)clip_len; a.nframe = (int)Tg; a.pad = n_fft / 2; a.edge_scale = 1.0f;
        a.norm = reinterpret_cast<const float4*>(stats); a.norm_div = (int)stats_div; a.norm_c = (int)stats_c;
        if (stats) SE_DISPATCH_GEO(n_fft, hop, ({ plan_analysis(nclip, Tg, a.gpc, a.nchunks, G::FR);
                                                 e = launch(k_analysis<G, LOAD_ZEROPAD, false, true>, (unsigned)(nclip * a.nchunks), G::NT, Smem<G>::ANALYSIS, (cudaStream_t)stream, a); }));
        else SE_DISPATCH_GEO(n_fft, hop, ({ plan_analysis(nclip, Tg, a.gpc, a.nchunks, G::FR);
                                            e = launch(k_analysis<G, LOAD_ZEROPAD, false, false>, (unsigned)(nclip * a.nchunks), G::NT, Smem<G>::ANALYSIS, (cudaStream_t)stream, a); }));
        if (e != cudaSuccess) return cuda_fail(e, "se_stft_segments_shared_fwd (clip transform) launch");
    }
    // 2. per-segment transform of the groups that hold reflect-boundary frames
    {
        AnaArgs a{};
        if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
        a.in = x; a.out = spec; a.in_stride = seg_stride; a.clip_stride = clip_stride; a.seg_rows = (int)nclip;
        a.clip_len = (int)clip_len; a.nsample = (int)nsample; a.in_len = (int)nsample;
        a.norm = reinterpret_cast<const float4*>(stats); a.norm_div = (int)stats_div; a.norm_c = (int)stats_c;
        a.nframe = (int)T; a.pad = 0; a.edge_scale = 1.0f;
        a.gpc = 1; a.edge_lead = (int)lead; a.edge_trail0 = (int)trail0; a.nchunks = (int)(lead + (ng - trail0));
        if (stats) SE_DISPATCH_GEO(n_fft, hop, (e = launch(k_analysis<G, LOAD_REFLECT, false, true>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS, (cudaStream_t)stream, a)));
        else SE_DISPATCH_GEO(n_fft, hop, (e = launch(k_analysis<G, LOAD_REFLECT, false, false>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS, (cudaStream_t)stream, a)));
        if (e != cudaSuccess) return cuda_fail(e, "se_stft_segments_shared_fwd (edge groups) launch");
    }
    // 3. interior frames: copies out of the clip-level spectrum
    const int F = n_fft / 2 + 1;
    const int64_t grows = rows * F;
    e = launch(k_segment_gather, (unsigned)((grows + kGatherRows - 1) / kGatherRows), 256u, 0, (cudaStream_t)stream,
               reinterpret_cast<const float2*>(scratch), reinterpret_cast<float2*>(spec), (int)nclip, F, (int)T, (int)Tg,
               (int)(seg_stride / hop), (int)(lead * FRG), (int)(trail0 * FRG), (int)grows);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_segments_shared_fwd (gather) launch");
}

}  // extern "C"
