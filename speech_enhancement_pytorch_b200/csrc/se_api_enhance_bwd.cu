// se_api_enhance_bwd.cu -- fused enhance, backward to the raw mask.
#include "se_host.h"
#include "se_fused.cuh"

using namespace se;

template <class G>
static cudaError_t run_enhance_bwd(EnhArgs a, int64_t rows, cudaStream_t st) {
    plan_analysis(rows, a.nframe, a.gpc, a.nchunks, G::FR);
    cudaError_t e;
    SE_DISPATCH_MASK(a.mode, a.pre_tanh, (e = launch(k_enhance_bwd<G, MODE, TANH>, (unsigned)(rows * a.nchunks), G::NT,
                                                     2 * Smem<G>::ZB + Smem<G>::STAGE + Smem<G>::TABLES + Smem<G>::WINDOW, st, a)));
    return e;
}

extern "C" int se_enhance_bwd(const float* gy, const float* x, const float* mask, float* gmask, int64_t rows, int64_t nsample,
                   int n_fft, int hop, int win_length, int mode, int pre_tanh, void* stream) {
    if (!gy || !x || !mask || !gmask) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (mode < 0 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "mask mode must be REAL/E/C/R");
    if (n_fft > 1024) return fail(SE_ERR_UNSUPPORTED, "se_enhance_bwd keeps two transforms in shared memory: n_fft <= 1024 "
                                                      "(compose se_stft_fwd + se_istft_bwd + se_mask_bwd for 2048)");
    const int64_t T = 1 + nsample / hop;
    EnhArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f / (float)win_length, a.ta)) return rc;
    if (int rc = get_tables(n_fft, hop, win_length, false, (float)win_length / (float)n_fft, a.ts)) return rc;
    a.x = x; a.mask = mask; a.gy = gy; a.out = gmask; a.nsample = (int)nsample; a.nframe = (int)T;
    a.mode = mode; a.pre_tanh = pre_tanh;
    cudaError_t e;
    if (n_fft == 512 && hop == 128) e = run_enhance_bwd<Geo<512, 128, 256>>(a, rows, (cudaStream_t)stream);
    else if (n_fft == 512) e = run_enhance_bwd<Geo<512, 256, 256>>(a, rows, (cudaStream_t)stream);
    else if (hop == 256) e = run_enhance_bwd<Geo<1024, 256, 256>>(a, rows, (cudaStream_t)stream);
    else e = run_enhance_bwd<Geo<1024, 512, 256>>(a, rows, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_enhance_bwd launch");
}
