// se_api_mask.cu -- mask application entry points.
#include "se_host.h"

using namespace se;

// ---- DCCRN's polar feature (ConvSTFT feature_type='real', src/model/dccrn.py:696-701) and its inverse
// (ConviSTFT(inputs, phase), :729-732) as single launches on the planar [rows, 2F, T] layout, full-precision
// sqrt / atan2 / sincos (no stack / cat / cos / sin / mul passes).  plane = F*T elements per row.
static __global__ void __launch_bounds__(256) k_polar_from_planar(const float* __restrict__ spec, float* __restrict__ mags,
                                                                  float* __restrict__ phase, int64_t plane, int bpr) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x / bpr;
    const int chunk = blockIdx.x - (int)row * bpr;
    for (int64_t i = (int64_t)chunk * blockDim.x + threadIdx.x; i < plane; i += (int64_t)bpr * blockDim.x) {
        const float re = __ldg(spec + row * 2 * plane + i), im = __ldg(spec + row * 2 * plane + plane + i);
        mags[row * plane + i] = sqrtf(re * re + im * im);
        phase[row * plane + i] = atan2f(im, re);
    }
}
static __global__ void __launch_bounds__(256) k_planar_from_polar(const float* __restrict__ mags, const float* __restrict__ phase,
                                                                  float* __restrict__ spec, int64_t plane, int bpr) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x / bpr;
    const int chunk = blockIdx.x - (int)row * bpr;
    for (int64_t i = (int64_t)chunk * blockDim.x + threadIdx.x; i < plane; i += (int64_t)bpr * blockDim.x) {
        float sn, cs;
        sincosf(__ldg(phase + row * plane + i), &sn, &cs);
        const float m = __ldg(mags + row * plane + i);
        spec[row * 2 * plane + i] = m * cs;
        spec[row * 2 * plane + plane + i] = m * sn;
    }
}
// gspec [rows, 2F, T] -> gmags = gre cos + gim sin, gphase = mags (gim cos - gre sin)
static __global__ void __launch_bounds__(256) k_planar_from_polar_bwd(const float* __restrict__ mags, const float* __restrict__ phase,
                                                                      const float* __restrict__ gspec, float* __restrict__ gmags,
                                                                      float* __restrict__ gphase, int64_t plane, int bpr) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x / bpr;
    const int chunk = blockIdx.x - (int)row * bpr;
    for (int64_t i = (int64_t)chunk * blockDim.x + threadIdx.x; i < plane; i += (int64_t)bpr * blockDim.x) {
        float sn, cs;
        sincosf(__ldg(phase + row * plane + i), &sn, &cs);
        const float gre = __ldg(gspec + row * 2 * plane + i), gim = __ldg(gspec + row * 2 * plane + plane + i);
        gmags[row * plane + i] = gre * cs + gim * sn;
        gphase[row * plane + i] = __ldg(mags + row * plane + i) * (gim * cs - gre * sn);
    }
}

static int planar_blocks_per_row(int64_t rows, int64_t plane);

extern "C" {

int se_polar_from_planar(const float* spec, float* mags, float* phase, int64_t rows, int64_t nbin, int64_t nframe, void* stream) {
    if (!spec || !mags || !phase || rows <= 0 || nbin <= 0 || nframe <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    const int64_t plane = nbin * nframe;
    const int bpr = planar_blocks_per_row(rows, plane);
    cudaError_t e = launch(k_polar_from_planar, (unsigned)(rows * bpr), 256u, 0, (cudaStream_t)stream, spec, mags, phase, plane, bpr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_polar_from_planar launch");
}
int se_planar_from_polar(const float* mags, const float* phase, float* spec, int64_t rows, int64_t nbin, int64_t nframe, void* stream) {
    if (!spec || !mags || !phase || rows <= 0 || nbin <= 0 || nframe <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    const int64_t plane = nbin * nframe;
    const int bpr = planar_blocks_per_row(rows, plane);
    cudaError_t e = launch(k_planar_from_polar, (unsigned)(rows * bpr), 256u, 0, (cudaStream_t)stream, mags, phase, spec, plane, bpr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_planar_from_polar launch");
}
int se_planar_from_polar_bwd(const float* mags, const float* phase, const float* gspec, float* gmags, float* gphase, int64_t rows,
                             int64_t nbin, int64_t nframe, void* stream) {
    if (!mags || !phase || !gspec || !gmags || !gphase || rows <= 0 || nbin <= 0 || nframe <= 0)
        return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    const int64_t plane = nbin * nframe;
    const int bpr = planar_blocks_per_row(rows, plane);
    cudaError_t e = launch(k_planar_from_polar_bwd, (unsigned)(rows * bpr), 256u, 0, (cudaStream_t)stream, mags, phase, gspec, gmags,
                           gphase, plane, bpr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_planar_from_polar_bwd launch");
}

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
// blocks per row: enough to fill the GPU a few times over, at least one, never more than the row has 256-element tiles
static int planar_blocks_per_row(int64_t rows, int64_t plane) {
    const int64_t tiles = (plane + 255) / 256;
    int64_t want = (148 * 16 + rows - 1) / rows;
    want = want < 1 ? 1 : (want > tiles ? tiles : want);
    return (int)want;
}
extern "C" {

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
