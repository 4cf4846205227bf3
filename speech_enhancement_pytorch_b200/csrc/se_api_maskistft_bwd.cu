// se_api_maskistft_bwd.cu -- backward of se_mask_istft_fwd to the raw mask: gmask = mask^T(spec, mask; iSTFT^T gy).
#include "se_host.h"
#include "se_fused.cuh"

using namespace se;

extern "C" int se_mask_istft_bwd(const float* gy, const float* spec, const float* mask, float* gmask, int64_t rows,
                                 int64_t nframe, int64_t length, int n_fft, int hop, int win_length, float scale, int mode,
                                 int pre_tanh, void* stream) {
    if (!gy || !spec || !mask || !gmask) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nframe <= 0 || length <= 0) return fail(SE_ERR_BAD_ARG, "nframe and length must be positive");
    if (int rc = check_common(rows, length, n_fft, hop, win_length)) return rc;
    if (mode < 0 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "mask mode must be REAL/E/C/R");
    MaskSynArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, scale / (float)n_fft, a.ts)) return rc;
    a.spec = spec; a.mask = mask; a.gy = gy; a.out = gmask; a.nframe = (int)nframe; a.length = (int)length;
    a.natural = (int)(n_fft + hop * (nframe - 1));
    cudaError_t e;
    SE_DISPATCH_MASK(mode, pre_tanh, SE_DISPATCH_GEO(n_fft, hop, (
        plan_analysis(rows, a.nframe, a.gpc, a.nchunks, G::FR),
        e = launch(k_mask_istft_bwd<G, MODE, TANH>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS,
                   (cudaStream_t)stream, a))));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mask_istft_bwd launch");
}
