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

}  // extern "C"
