// se_api_mask.cu -- mask application entry points.
#include "se_host.h"

using namespace se;

extern "C" {

int se_mask_fwd(const float* spec, const float* mask, float* out, int64_t count, int mode, int pre_tanh, void* stream) {
    if (!spec || !mask || !out || count <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    if (mode < 0 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "mask mode must be REAL/E/C/R");
    int64_t blocks = (count / 2 + 255) / 256 + 1;
    if (blocks > 148 * 12) blocks = 148 * 12;    // grid-stride loop over a few waves
    cudaError_t e;
    SE_DISPATCH_MASK(mode, pre_tanh, (e = launch(k_mask_fwd_t<MODE, TANH>, (unsigned)blocks, 256, 0, (cudaStream_t)stream,
                                                 reinterpret_cast<const float2*>(spec), mask, reinterpret_cast<float2*>(out), count)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mask_fwd launch");
}

int se_mask_bwd(const float* spec, const float* mask, const float* gout, float* gmask, float* gspec, int64_t count,
                int mode, int pre_tanh, void* stream) {
    if (!spec || !mask || !gout || !gmask || count <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    if (mode < 0 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "mask mode must be REAL/E/C/R");
    int64_t blocks = (count / 2 + 255) / 256 + 1;
    if (blocks > 148 * 12) blocks = 148 * 12;    // grid-stride loop over a few waves
    cudaError_t e;
    SE_DISPATCH_MASK(mode, pre_tanh, (e = launch(k_mask_bwd_t<MODE, TANH>, (unsigned)blocks, 256, 0, (cudaStream_t)stream,
                                                 reinterpret_cast<const float2*>(spec), mask,
                                                 reinterpret_cast<const float2*>(gout), gmask,
                                                 reinterpret_cast<float2*>(gspec), count)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mask_bwd launch");
}

// blocks per row: enough to fill the GPU a few times over, at least one, never more than the row has 256-element tiles
static int planar_blocks_per_row(int64_t rows, int64_t plane) {
    const int64_t tiles = (plane + 255) / 256;
    int64_t want = (148 * 16 + rows - 1) / rows;
    want = want < 1 ? 1 : (want > tiles ? tiles : want);
    return (int)want;
}

#define SE_DISPATCH_PLANAR(mode, CALL)                                                   \
    do {                                                                                 \
        if (mode == 1) { constexpr int MODE = 1; CALL; }                                 \
        else if (mode == 2) { constexpr int MODE = 2; CALL; }                            \
        else { constexpr int MODE = 3; CALL; }                                           \
    } while (0)

int se_mask_planar_fwd(const float* spec, const float* mask_re, const float* mask_im, float* out, int64_t rows, int64_t nbin,
                       int64_t nframe, int mode, void* stream) {
    if (!spec || !mask_re || !mask_im || !out) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (rows <= 0 || nbin <= 0 || nframe <= 0) return fail(SE_ERR_BAD_ARG, "rows, nbin and nframe must be positive");
    if (mode < 1 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "DCCRN mask mode must be E/C/R");
    const int64_t plane = nbin * nframe;
    const int bpr = planar_blocks_per_row(rows, plane);
    if (rows * bpr > 0x7fffffffLL) return fail(SE_ERR_BAD_ARG, "problem too large for one launch");
    cudaError_t e;
    SE_DISPATCH_PLANAR(mode, (e = launch(k_mask_planar_fwd_t<MODE>, (unsigned)(rows * bpr), 256, 0, (cudaStream_t)stream, spec, mask_re,
                                         mask_im, out, plane, bpr)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mask_planar_fwd launch");
}

int se_mask_planar_bwd(const float* spec, const float* mask_re, const float* mask_im, const float* gout, float* gmask_re,
                       float* gmask_im, float* gspec, int64_t rows, int64_t nbin, int64_t nframe, int mode, void* stream) {
    if (!spec || !mask_re || !mask_im || !gout || !gmask_re || !gmask_im) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (rows <= 0 || nbin <= 0 || nframe <= 0) return fail(SE_ERR_BAD_ARG, "rows, nbin and nframe must be positive");
    if (mode < 1 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "DCCRN mask mode must be E/C/R");
    const int64_t plane = nbin * nframe;
    const int bpr = planar_blocks_per_row(rows, plane);
    if (rows * bpr > 0x7fffffffLL) return fail(SE_ERR_BAD_ARG, "problem too large for one launch");
    cudaError_t e;
    SE_DISPATCH_PLANAR(mode, (e = launch(k_mask_planar_bwd_t<MODE>, (unsigned)(rows * bpr), 256, 0, (cudaStream_t)stream, spec, mask_re,
                                         mask_im, gout, gmask_re, gmask_im, gspec, plane, bpr)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mask_planar_bwd launch");
}

}  // extern "C"
