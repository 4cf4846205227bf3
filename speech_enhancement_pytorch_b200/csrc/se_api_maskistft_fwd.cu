// se_api_maskistft_fwd.cu -- model tail + iSTFT in one launch: y = istft_custom(apply_mask(spec, mask)).
#include "se_host.h"
#include "se_fused.cuh"

using namespace se;

extern "C" int se_mask_istft_fwd(const float* spec, const float* mask, float* y, int64_t rows, int64_t nframe, int64_t length,
                                 int n_fft, int hop, int win_length, float scale, int mode, int pre_tanh, void* stream) {
    if (!spec || !mask || !y) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nframe <= 0 || length <= 0) return fail(SE_ERR_BAD_ARG, "nframe and length must be positive");
    if (int rc = check_common(rows, length, n_fft, hop, win_length)) return rc;
    if (mode < 0 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "mask mode must be REAL/E/C/R");
    if (!envelope_ok(n_fft, hop, win_length, false, nframe, n_fft / 2, n_fft / 2 + length, 1e-11))
        return fail(SE_ERR_ENVELOPE, "window overlap add min < 1e-11 (torch.istft raises the same)");
    MaskSynArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, scale / (float)n_fft, a.ts)) return rc;
    a.spec = spec; a.mask = mask; a.out = y; a.nframe = (int)nframe; a.length = (int)length;
    a.natural = (int)(n_fft + hop * (nframe - 1));
    a.b_lo = (n_fft / 2) / hop; a.b_hi = (int)((n_fft / 2 + length + hop - 1) / hop);
    cudaError_t e;
    SE_DISPATCH_MASK(mode, pre_tanh, SE_DISPATCH_GEO(n_fft, hop, (
        a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, G::OLA, G::MINB, G::FR),
        e = launch(k_mask_istft_fwd<G, MODE, TANH>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::SYNTH_ISTFT,
                   (cudaStream_t)stream, a))));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mask_istft_fwd launch");
}
