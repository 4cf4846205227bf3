// se_api_enhance_fwd.cu -- fused wave -> STFT -> mask -> iSTFT -> wave, forward.
#include "se_host.h"
#include "se_fused.cuh"

using namespace se;

extern "C" int se_enhance_fwd(const float* x, const float* mask, float* y, int64_t rows, int64_t nsample, int n_fft, int hop,
                   int win_length, int mode, int pre_tanh, void* stream) {
    if (!x || !mask || !y) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (mode < 0 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "mask mode must be REAL/E/C/R");
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    const int64_t T = 1 + nsample / hop;
    if (!envelope_ok(n_fft, hop, win_length, false, T, n_fft / 2, n_fft / 2 + nsample, 1e-11))
        return fail(SE_ERR_ENVELOPE, "window overlap add min < 1e-11 (torch.istft raises the same)");
    EnhArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f / (float)win_length, a.ta)) return rc;
    if (int rc = get_tables(n_fft, hop, win_length, false, (float)win_length / (float)n_fft, a.ts)) return rc;
    a.x = x; a.mask = mask; a.out = y; a.nsample = (int)nsample; a.nframe = (int)T;
    a.b_lo = (n_fft / 2) / hop; a.b_hi = (int)((n_fft / 2 + nsample + hop - 1) / hop);
    a.mode = mode; a.pre_tanh = pre_tanh;
    cudaError_t e;
    SE_DISPATCH_MASK(mode, pre_tanh, SE_DISPATCH_GEO(n_fft, hop, (
        a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, G::OLA, G::MINB, G::FR),
        e = launch(k_enhance_fwd<G, MODE, TANH>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::FUSED_ISTFT,
                   (cudaStream_t)stream, a))));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_enhance_fwd launch");
}
