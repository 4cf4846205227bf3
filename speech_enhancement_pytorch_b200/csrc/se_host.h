// se_host.h -- host-side helpers shared by the C-ABI translation units (se_api_*.cu).
#pragma once
#include "../../include/se_b200.h"
#include "se_kernels2.cuh"
#include "se_kernels3.cuh"

#include <string>
#include <vector>

namespace se {

int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
const char* last_error();
// constant tables per (device, n, hop, win_len, placement, scale); `scale` multiplies the window,
// front=true puts a short window at the start of the frame (DCCRN)
int get_tables(int n, int hop, int win_len, bool front, float scale, Tables& out, int window_id = 0);
// window_id > 0: values registered with se_register_window (any window type); 0: periodic Hann
int register_window(const double* values, int win_len);
// torch.istft's "window overlap add min" check on the host (no device sync)
bool envelope_ok(int n, int hop, int win_len, bool front, int64_t T, int64_t lo, int64_t hi, double floor_);
void plan_analysis(int64_t rows, int64_t T, int& gpc, int& nchunks, int frames_per_group);
// g_min..g_max: groups per chunk considered (the pair engine's adjoint emitters need chunks of >= 7 blocks: g_min = 2)
int plan_synthesis(int64_t rows, int nb, int ola, int ctas_per_sm, int frames_per_group, int g_min = 1, int g_max = 8);
// SE_ENGINE=2 selects the signal-pair engine (se_fft2.cuh) where both exist; default 1: the scalar engine, which
// measured faster on B200 (profiles/r02_notes.md)
int engine_version();
int frames8();
int check_common(int64_t rows, int64_t nsample, int n_fft, int hop, int win_length);
bool host_window_values(int n, int win_len, bool front, std::vector<double>& w, int window_id);

// ---- general-geometry path (se_api_generic.cu): any power-of-two n_fft, any hop / win_length; the tuned engine only
// exists for n_fft 512 / 1024 / 2048 at hop n/4, n/2 (and DCCRN's 400/100/512)
bool geometry_tuned(int n_fft, int hop);
int gen_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length, float scale, cudaStream_t st);
int gen_stft_bwd(const float* gspec, float* gx, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length, float scale,
                 int accumulate, cudaStream_t st);
int gen_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop, int win_length,
                  float scale, cudaStream_t st);
int gen_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop, int win_length,
                  float scale, cudaStream_t st);
int gen_conv_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int win_len, int win_inc, int fft_len,
                      int window_id, cudaStream_t st);
int gen_conv_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc,
                       int fft_len, int window_id, cudaStream_t st);
int gen_conv_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc,
                       int fft_len, int window_id, cudaStream_t st);

// MODE (0..3) x TANH -> compile-time template arguments
#define SE_DISPATCH_MASK(mode, pre_tanh, CALL)                                         \
    do {                                                                               \
        if (pre_tanh) {                                                                \
            constexpr bool TANH = true;                                                \
            if (mode == 0) { constexpr int MODE = 0; CALL; } else if (mode == 1) { constexpr int MODE = 1; CALL; } \
            else if (mode == 2) { constexpr int MODE = 2; CALL; } else { constexpr int MODE = 3; CALL; }           \
        } else {                                                                       \
            constexpr bool TANH = false;                                               \
            if (mode == 0) { constexpr int MODE = 0; CALL; } else if (mode == 1) { constexpr int MODE = 1; CALL; } \
            else if (mode == 2) { constexpr int MODE = 2; CALL; } else { constexpr int MODE = 3; CALL; }           \
        }                                                                              \
    } while (0)

#define SE_DISPATCH_GEO(n_fft, hop, CALL)                                              \
    do {                                                                               \
        if (frames8() && n_fft == 512 && hop == 128) { using G = Geo<512, 128, 128, 8>; CALL; }        \
        else if (frames8() && n_fft == 1024 && hop == 256) { using G = Geo<1024, 256, 128, 8>; CALL; } \
        else if (n_fft == 512 && hop == 128) { using G = Geo<512, 128, 256>; CALL; }   \
        else if (n_fft == 512 && hop == 256) { using G = Geo<512, 256, 256>; CALL; }   \
        else if (n_fft == 1024 && hop == 256) { using G = Geo<1024, 256, 256>; CALL; } \
        else if (n_fft == 1024 && hop == 512) { using G = Geo<1024, 512, 256>; CALL; } \
        else if (n_fft == 2048 && hop == 512) { using G = Geo<2048, 512, 512>; CALL; } \
        else { using G = Geo<2048, 1024, 512>; CALL; }                                 \
    } while (0)


}  // namespace se
