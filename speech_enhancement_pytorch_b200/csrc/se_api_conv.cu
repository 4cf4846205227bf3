// se_api_conv.cu -- DCCRN ConvSTFT / ConviSTFT entry points.
#include "se_host.h"
#include "se_conv.cuh"

using namespace se;

extern "C" int se_register_window(const double* values, int win_len) { return register_window(values, win_len); }

extern "C" int se_conv_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int win_len, int win_inc, int fft_len,
                     void* stream) {
    return se_conv_stft_fwd_w(x, spec, rows, nsample, win_len, win_inc, fft_len, 0, stream);
}
extern "C" int se_conv_stft_fwd_w(const float* x, float* spec, int64_t rows, int64_t nsample, int win_len, int win_inc, int fft_len,
                     int window_id, void* stream) {
    if (!x || !spec || rows <= 0 || nsample <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    if (fft_len != 512 || (win_inc != 100 && win_inc != 128) || win_len > fft_len || win_len < win_inc)
        return gen_conv_stft_fwd(x, spec, rows, nsample, win_len, win_inc, fft_len, window_id, (cudaStream_t)stream);
    const int pad = win_len - win_inc;
    const int64_t T = (nsample + 2 * pad - win_len) / win_inc + 1;
    if (T <= 0) return fail(SE_ERR_BAD_ARG, "ConvSTFT: input shorter than one frame");
    AnaArgs a{};
    if (int rc = get_tables(fft_len, win_inc, win_len, true, 0.5f, a.tb, window_id)) return rc;
    a.in = x; a.out = spec; a.in_stride = nsample; a.seg_rows = 1; a.nsample = (int)nsample; a.in_len = (int)nsample;
    a.nframe = (int)T; a.pad = pad; a.edge_scale = 1.0f;
    plan_analysis(rows, T, a.gpc, a.nchunks, 16);
    cudaError_t e;
    if (win_inc == 100) {
        using G = Geo<512, 100, 256>;
        e = launch(k_analysis<G, LOAD_ZEROPAD, true>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS,
                   (cudaStream_t)stream, a);
    } else if (win_inc == 128) {
        using G = Geo<512, 128, 256>;
        e = launch(k_analysis<G, LOAD_ZEROPAD, true>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS,
                   (cudaStream_t)stream, a);
    } else {
        return fail(SE_ERR_UNSUPPORTED, "ConvSTFT: unreachable geometry");
    }
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_conv_stft_fwd launch");
}
static int conv_args(ConvArgs& a, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc, int fft_len, int window_id) {
    if (rows <= 0 || nframe <= 0 || out_len <= 0) return fail(SE_ERR_BAD_ARG, "empty tensor");
    if (fft_len != 512 || win_inc != 100 || win_len > 4 * win_inc || win_len < win_inc || (win_len & 1))
        return fail(SE_ERR_UNSUPPORTED, "ConviSTFT: built for fft_len 512, win_inc 100, even win_len <= 400 (DCCRN's 400/100/512)");
    const int64_t total = win_len + (int64_t)win_inc * (nframe - 1);
    a.pad = win_len - win_inc;
    if (out_len > total - a.pad) return fail(SE_ERR_BAD_ARG, "ConviSTFT: out_len exceeds the overlap-added signal");
    // window at the front of the frame; scale = 1/2 (Hermitian weight) * 2/n (the (n/2)^-1 of the pinv)
    if (int rc = get_tables(fft_len, win_inc, win_len, true, 1.0f / (float)fft_len, a.tb, window_id)) return rc;
    a.win_len = win_len; a.nframe = (int)nframe; a.out_len = (int)out_len;
    a.inv_even = 1.0f / (float)(fft_len / 2 + (win_len + 1) / 2);
    a.inv_odd = 1.0f / (float)(fft_len / 2 + win_len / 2);
    return 0;
}

extern "C" int se_conv_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc, int fft_len, void* stream) {
    return se_conv_istft_fwd_w(spec, y, rows, nframe, out_len, win_len, win_inc, fft_len, 0, stream);
}
extern "C" int se_conv_istft_fwd_w(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t out_len, int win_len,
                      int win_inc, int fft_len, int window_id, void* stream) {
    if (!spec || !y) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (!se_conv_geometry_tuned(win_len, win_inc, fft_len))
        return gen_conv_istft_fwd(spec, y, rows, nframe, out_len, win_len, win_inc, fft_len, window_id, (cudaStream_t)stream);
    ConvArgs a{};
    if (int rc = conv_args(a, rows, nframe, out_len, win_len, win_inc, fft_len, window_id)) return rc;
    a.in = spec; a.out = y;
    a.b_lo = a.pad / win_inc; a.b_hi = (int)((a.pad + out_len + win_inc - 1) / win_inc);
    a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, 4, 2, 16);
    using G = Geo<512, 100, 256>;
    cudaError_t e = launch(k_conv_istft<G>, (unsigned)(rows * a.nchunks), G::NT, ConvGeo<G>::SYNTH, (cudaStream_t)stream, a);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_conv_istft_fwd launch");
}

extern "C" int se_conv_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc, int fft_len, void* stream) {
    return se_conv_istft_bwd_w(gy, gspec, rows, nframe, out_len, win_len, win_inc, fft_len, 0, stream);
}
extern "C" int se_conv_istft_bwd_w(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t out_len, int win_len,
                      int win_inc, int fft_len, int window_id, void* stream) {
    if (!gy || !gspec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (!se_conv_geometry_tuned(win_len, win_inc, fft_len))
        return gen_conv_istft_bwd(gy, gspec, rows, nframe, out_len, win_len, win_inc, fft_len, window_id, (cudaStream_t)stream);
    ConvArgs a{};
    if (int rc = conv_args(a, rows, nframe, out_len, win_len, win_inc, fft_len, window_id)) return rc;
    a.in = gy; a.out = gspec;
    plan_analysis(rows, nframe, a.gpc, a.nchunks, 16);
    using G = Geo<512, 100, 256>;
    cudaError_t e = launch(k_conv_istft_adj<G>, (unsigned)(rows * a.nchunks), G::NT, ConvGeo<G>::ADJ, (cudaStream_t)stream, a);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_conv_istft_bwd launch");
}

// ---- DCCRN model tail + ConviSTFT in one launch each way (src/model/dccrn.py:203-224): the masked spectrum and its
// gradient are never written
#define SE_DISPATCH_CONV_MODE(mode, CALL)                                                \
    do {                                                                                 \
        if (mode == 1) { constexpr int MODE = 1; CALL; }                                 \
        else if (mode == 2) { constexpr int MODE = 2; CALL; }                            \
        else { constexpr int MODE = 3; CALL; }                                           \
    } while (0)

extern "C" int se_conv_mask_istft_fwd(const float* spec, const float* mask_re, const float* mask_im, float* y, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc, int fft_len, int mode, void* stream) {
    return se_conv_mask_istft_fwd_w(spec, mask_re, mask_im, y, rows, nframe, out_len, win_len, win_inc, fft_len, mode, 0, stream);
}
extern "C" int se_conv_mask_istft_fwd_w(const float* spec, const float* mask_re, const float* mask_im, float* y, int64_t rows,
                                      int64_t nframe, int64_t out_len, int win_len, int win_inc, int fft_len, int mode, int window_id, void* stream) {
    if (!spec || !mask_re || !mask_im || !y) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (mode < 1 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "DCCRN mask mode must be E/C/R");
    ConvArgs a{};
    if (int rc = conv_args(a, rows, nframe, out_len, win_len, win_inc, fft_len, window_id)) return rc;
    a.in = spec; a.mre = mask_re; a.mim = mask_im; a.out = y;
    a.b_lo = a.pad / win_inc; a.b_hi = (int)((a.pad + out_len + win_inc - 1) / win_inc);
    a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, 4, 2, 16);
    using G = Geo<512, 100, 256>;
    cudaError_t e;
    SE_DISPATCH_CONV_MODE(mode, (e = launch(k_conv_istft<G, MODE>, (unsigned)(rows * a.nchunks), G::NT, ConvGeo<G>::SYNTH,
                                            (cudaStream_t)stream, a)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_conv_mask_istft_fwd launch");
}

extern "C" int se_conv_mask_istft_bwd(const float* gy, const float* spec, const float* mask_re, const float* mask_im, float* gmask_re, float* gmask_im, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc, int fft_len, int mode, void* stream) {
    return se_conv_mask_istft_bwd_w(gy, spec, mask_re, mask_im, gmask_re, gmask_im, rows, nframe, out_len, win_len, win_inc, fft_len, mode, 0, stream);
}
extern "C" int se_conv_mask_istft_bwd_w(const float* gy, const float* spec, const float* mask_re, const float* mask_im,
                                      float* gmask_re, float* gmask_im, int64_t rows, int64_t nframe, int64_t out_len, int win_len,
                                      int win_inc, int fft_len, int mode, int window_id, void* stream) {
    if (!gy || !spec || !mask_re || !mask_im || !gmask_re || !gmask_im) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (mode < 1 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "DCCRN mask mode must be E/C/R");
    ConvArgs a{};
    if (int rc = conv_args(a, rows, nframe, out_len, win_len, win_inc, fft_len, window_id)) return rc;
    a.in = gy; a.spec = spec; a.mre = mask_re; a.mim = mask_im; a.gre = gmask_re; a.gim = gmask_im;
    plan_analysis(rows, nframe, a.gpc, a.nchunks, 16);
    using G = Geo<512, 100, 256>;
    cudaError_t e;
    SE_DISPATCH_CONV_MODE(mode, (e = launch(k_conv_istft_adj<G, MODE>, (unsigned)(rows * a.nchunks), G::NT, ConvGeo<G>::ADJ,
                                            (cudaStream_t)stream, a)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_conv_mask_istft_bwd launch");
}
