// se_api_ola.cu -- Conv-TasNet's decoder tail on device: overlap_and_add(signal, frame_step),
// src/model/conv_tasnet.py:11-31.  The reference scatters sub-frames with index_add_; here every output sample
// GATHERS its (at most ceil(L/step)) contributions in increasing frame order -- the order index_add_ applies them
// in -- so the result is deterministic and needs no atomics or zero-fill.  The backward is the plain gather
// gsignal[f][j] = gout[f*step + j].
#include "se_host.h"

using namespace se;

namespace {

__global__ void __launch_bounds__(256) k_overlap_add(const float* __restrict__ sig, float* __restrict__ out, int frames,
                                                     int len, int step, int64_t out_len, int bpr) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x / bpr;
    const int bx = blockIdx.x - (int)row * bpr;
    const float* s = sig + row * (int64_t)frames * len;
    float* o = out + row * out_len;
    for (int64_t i = (int64_t)bx * 256 + threadIdx.x; i < out_len; i += (int64_t)bpr * 256) {
        int64_t f_hi = i / step;
        if (f_hi > frames - 1) f_hi = frames - 1;
        int64_t f_lo = i - len + 1 <= 0 ? 0 : (i - len + step) / step;       // ceil((i - len + 1) / step)
        float acc = 0.f;
        for (int64_t f = f_lo; f <= f_hi; ++f) acc += __ldg(s + f * len + (i - f * step));
        o[i] = acc;
    }
}

__global__ void __launch_bounds__(256) k_overlap_add_bwd(const float* __restrict__ gout, float* __restrict__ gsig,
                                                         int frames, int len, int step, int64_t out_len, int bpr) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x / bpr;
    const int bx = blockIdx.x - (int)row * bpr;
    const float* g = gout + row * out_len;
    float* d = gsig + row * (int64_t)frames * len;
    const int64_t total = (int64_t)frames * len;
    for (int64_t i = (int64_t)bx * 256 + threadIdx.x; i < total; i += (int64_t)bpr * 256) {
        const int64_t f = i / len, j = i - f * len;
        d[i] = __ldg(g + f * step + j);
    }
}

int check_ola(int64_t rows, int64_t frames, int len, int step) {
    if (rows <= 0 || frames <= 0 || len <= 0 || step <= 0) return fail(SE_ERR_BAD_ARG, "rows, frames, frame_length and frame_step must be positive");
    if (rows > (1 << 22)) return fail(SE_ERR_UNSUPPORTED, "more than 4M rows per call");
    if (frames * (int64_t)len > (int64_t)1 << 40) return fail(SE_ERR_BAD_ARG, "signal too large");
    return 0;
}

unsigned ola_blocks(int64_t per_row, int64_t rows) {
    int64_t b = (per_row + 255) / 256;
    const int64_t cap = (148 * 8 + rows - 1) / rows;     // about 8 CTAs of 256 threads per SM over all rows
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int se_overlap_add_fwd(const float* signal, float* out, int64_t rows, int64_t frames, int frame_length,
                                  int frame_step, void* stream) {
    if (!signal || !out) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (int rc = check_ola(rows, frames, frame_length, frame_step)) return rc;
    const int64_t out_len = (int64_t)frame_step * (frames - 1) + frame_length;
    const unsigned bpr = ola_blocks(out_len, rows);
    cudaError_t e = launch(k_overlap_add, (unsigned)(rows * bpr), 256, 0, (cudaStream_t)stream, signal, out, (int)frames,
                           frame_length, frame_step, out_len, (int)bpr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_overlap_add_fwd launch");
}

extern "C" int se_overlap_add_bwd(const float* gout, float* gsignal, int64_t rows, int64_t frames, int frame_length,
                                  int frame_step, void* stream) {
    if (!gout || !gsignal) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (int rc = check_ola(rows, frames, frame_length, frame_step)) return rc;
    const int64_t out_len = (int64_t)frame_step * (frames - 1) + frame_length;
    const unsigned bpr = ola_blocks(frames * frame_length, rows);
    cudaError_t e = launch(k_overlap_add_bwd, (unsigned)(rows * bpr), 256, 0, (cudaStream_t)stream, gout, gsignal,
                           (int)frames, frame_length, frame_step, out_len, (int)bpr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_overlap_add_bwd launch");
}
