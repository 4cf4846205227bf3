// se_capi.cu -- host side of libse_b200.so: constant-table cache, launch planning, C-ABI.
// See include/se_b200.h for the contract of every entry point.
#include "../../include/se_b200.h"
#include "se_kernels.cuh"
#include "se_conv.cuh"
#include "se_fused.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

namespace se {

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(SE_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------ constant tables
struct TableKey {
    int dev, n, hop, win_len, front;
    uint32_t scale_bits;
    bool operator<(const TableKey& o) const {
        return std::tie(dev, n, hop, win_len, front, scale_bits) <
               std::tie(o.dev, o.n, o.hop, o.win_len, o.front, o.scale_bits);
    }
};
static std::mutex g_mu;
static std::map<TableKey, Tables> g_tables;

static void host_window(int n, int win_len, bool front, std::vector<double>& w) {
    w.assign(n, 0.0);
    const int left = front ? 0 : (n - win_len) / 2;
    const double two_pi = 6.283185307179586476925286766559;
    for (int j = 0; j < win_len; ++j) w[left + j] = 0.5 - 0.5 * std::cos(two_pi * j / win_len);
}

// `scale` multiplies the window; front=true puts a short window at the start of the frame (DCCRN)
static int get_tables(int n, int hop, int win_len, bool front, float scale, Tables& out) {
    int dev = 0;
    cudaGetDevice(&dev);
    TableKey key{dev, n, hop, win_len, front ? 1 : 0, 0};
    std::memcpy(&key.scale_bits, &scale, 4);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) { out = it->second; return 0; }
    const int M = n / 2;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<double> w;
    host_window(n, win_len, front, w);
    std::vector<float> win(n), w2(n), inv_env(hop);
    std::vector<float2> tw(M), twn(M);
    for (int j = 0; j < n; ++j) {
        win[j] = (float)(w[j] * (double)scale);
        const float wf = (float)w[j];
        w2[j] = wf * wf;
    }
    for (int k = 0; k < M; ++k) {
        tw[k] = make_float2((float)std::cos(two_pi * k / M), (float)-std::sin(two_pi * k / M));
        twn[k] = make_float2((float)std::cos(two_pi * k / n), (float)-std::sin(two_pi * k / n));
    }
    for (int o = 0; o < hop; ++o) {
        float e = 0.f;
        for (int q = 0; o + q * hop < n; ++q) e += w2[o + q * hop];
        inv_env[o] = e > 0.f ? 1.0f / e : 0.f;
    }
    float *d_win = nullptr, *d_w2 = nullptr, *d_env = nullptr;
    float2 *d_tw = nullptr, *d_twn = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&d_win, n * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&d_w2, n * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&d_env, hop * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&d_tw, M * sizeof(float2))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&d_twn, M * sizeof(float2))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cudaMemcpy(d_win, win.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d_w2, w2.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d_env, inv_env.data(), hop * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d_tw, tw.data(), M * sizeof(float2), cudaMemcpyHostToDevice);
    e = cudaMemcpy(d_twn, twn.data(), M * sizeof(float2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(tables)");
    out = Tables{d_win, d_tw, d_twn, d_w2, d_env};
    g_tables[key] = out;
    return 0;
}

// torch.istft's "window overlap add min" check, evaluated on the host (no device sync):
// the envelope depends only on the configuration.
static bool envelope_ok(int n, int hop, int win_len, bool front, int64_t T, int64_t lo, int64_t hi, double floor_) {
    std::vector<double> w;
    host_window(n, win_len, front, w);
    const int ola = (n + hop - 1) / hop;
    // blocks whose set of contributing frames is not "all": first ola-1 and those past T-1
    auto env_at = [&](int64_t i) {
        const int64_t b = i / hop;
        const int o = (int)(i - b * hop);
        double e = 0.0;
        for (int q = 0; q < ola && o + q * hop < n; ++q) {
            const int64_t t = b - q;
            if (t >= 0 && t < T) { const float wf = (float)w[o + q * hop]; e += (double)(wf * wf); }
        }
        return e;
    };
    const int64_t natural = n + hop * (T - 1);
    hi = hi < natural ? hi : natural;
    if (hi <= lo) return true;
    const int64_t edge = (int64_t)ola * hop;
    for (int64_t i = lo; i < hi; ++i) {
        if (i >= lo + edge + hop && i < hi - edge - hop) { i = hi - edge - hop - 1; continue; }   // interior is periodic
        if (std::fabs(env_at(i)) < floor_) return false;
    }
    return true;
}

// ------------------------------------------------------------------ launch planning
static int g_target_ctas = 148 * 8;

static void plan_analysis(int64_t rows, int64_t T, int& gpc, int& nchunks) {
    const int64_t ng = (T + 15) / 16;
    int64_t g = (rows * ng) / g_target_ctas;
    g = g < 1 ? 1 : (g > 8 ? 8 : g);
    gpc = (int)g;
    nchunks = (int)((ng + g - 1) / g);
}
static int plan_synthesis(int64_t rows, int nb, int ola) {
    int64_t ng = (nb + 15) / 16;
    int64_t g = (rows * ng) / g_target_ctas;
    g = g < 1 ? 1 : (g > 8 ? 8 : g);
    const int cb_max = (int)(16 * g) - (ola - 1);
    int nchunks = (nb + cb_max - 1) / cb_max;
    return nchunks < 1 ? 1 : nchunks;
}

static int check_common(int64_t rows, int64_t nsample, int n_fft, int hop, int win_length) {
    if (rows <= 0 || nsample <= 0) return fail(SE_ERR_BAD_ARG, "rows and nsample must be positive");
    if (n_fft != 512 && n_fft != 1024 && n_fft != 2048)
        return fail(SE_ERR_UNSUPPORTED, "n_fft must be 512, 1024 or 2048 (no fallback path exists)");
    if (hop * 4 != n_fft && hop * 2 != n_fft)
        return fail(SE_ERR_UNSUPPORTED, "hop_length must be n_fft/4 or n_fft/2");
    if (win_length < 2 || win_length > n_fft) return fail(SE_ERR_UNSUPPORTED, "need 2 <= win_length <= n_fft");
    if (rows * ((nsample / hop + 1 + 15) / 16) > 0x7fffffffLL) return fail(SE_ERR_BAD_ARG, "problem too large for one launch");
    return 0;
}

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
        if (n_fft == 512 && hop == 128) { using G = Geo<512, 128, 256>; CALL; }        \
        else if (n_fft == 512 && hop == 256) { using G = Geo<512, 256, 256>; CALL; }   \
        else if (n_fft == 1024 && hop == 256) { using G = Geo<1024, 256, 256>; CALL; } \
        else if (n_fft == 1024 && hop == 512) { using G = Geo<1024, 512, 256>; CALL; } \
        else if (n_fft == 2048 && hop == 512) { using G = Geo<2048, 512, 512>; CALL; } \
        else { using G = Geo<2048, 1024, 512>; CALL; }                                 \
    } while (0)

// ------------------------------------------------------------------ launchers
template <class G, int LMODE>
static cudaError_t run_analysis(const AnaArgs& a, int64_t rows, cudaStream_t st) {
    return launch(k_analysis<G, LMODE, false>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS, st, a);
}
template <class G, int EMODE>
static cudaError_t run_synthesis(const SynArgs& a, int64_t rows, cudaStream_t st) {
    return launch(k_synthesis<G, EMODE>, (unsigned)(rows * a.nchunks), G::NT,
                  EMODE == EMIT_ADJ ? Smem<G>::SYNTH_ADJ : Smem<G>::SYNTH_ISTFT, st, a);
}
// loss kernels keep |B|^2 per pass-C task in registers: one task per thread (NT = M) for n <= 1024
#define SE_DISPATCH_LOSS_GEO(n_fft, CALL)                                              \
    do {                                                                               \
        if (n_fft == 512) { using G = Geo<512, 128, 256>; CALL; }                      \
        else if (n_fft == 1024) { using G = Geo<1024, 256, 512>; CALL; }               \
        else { using G = Geo<2048, 512, 512>; CALL; }                                  \
    } while (0)

template <class G>
static cudaError_t run_loss_fwd(const LossArgs& a, int64_t rows, cudaStream_t st) {
    return launch(k_loss_fwd<G>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::ANALYSIS, st, a);
}
template <class G>
static cudaError_t run_loss_bwd(const LossArgs& a, int64_t rows, cudaStream_t st) {
    return launch(k_loss_bwd<G>, (unsigned)(rows * a.nchunks), G::NT, Smem<G>::FUSED_ADJ, st, a);
}

static const int kRes[3][3] = {{512, 128, 512}, {1024, 256, 1024}, {2048, 512, 2048}};

}  // namespace se

using namespace se;

extern "C" {

int se_version(void) { return 100; }
const char* se_last_error(void) { return g_err.c_str(); }

int se_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                float scale, void* stream) {
    if (!x || !spec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    AnaArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.in = x; a.out = spec; a.in_stride = nsample; a.seg_rows = 1; a.nsample = (int)nsample; a.in_len = (int)nsample;
    a.nframe = (int)(1 + nsample / hop); a.pad = 0; a.edge_scale = 1.0f;
    plan_analysis(rows, a.nframe, a.gpc, a.nchunks);
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_analysis<G, LOAD_REFLECT>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_fwd launch");
}

int se_stft_segments_fwd(const float* x, float* spec, int64_t nseg, int64_t nclip, int64_t clip_len, int64_t clip_stride,
                         int64_t seg_stride, int64_t nsample, int n_fft, int hop, int win_length, float scale, void* stream) {
    if (!x || !spec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nseg <= 0 || nclip <= 0 || clip_len <= 0 || seg_stride <= 0 || clip_stride < clip_len)
        return fail(SE_ERR_BAD_ARG, "bad segment geometry");
    const int64_t rows = nseg * nclip;
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    if ((nseg - 1) * seg_stride >= clip_len) return fail(SE_ERR_BAD_ARG, "last segment starts beyond the clip");
    AnaArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.in = x; a.out = spec; a.in_stride = seg_stride; a.clip_stride = clip_stride; a.seg_rows = (int)nclip;
    a.clip_len = (int)clip_len; a.nsample = (int)nsample; a.in_len = (int)nsample;
    a.nframe = (int)(1 + nsample / hop); a.pad = 0; a.edge_scale = 1.0f;
    plan_analysis(rows, a.nframe, a.gpc, a.nchunks);
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_analysis<G, LOAD_REFLECT>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_segments_fwd launch");
}

int se_stft_bwd(const float* gspec, float* gx, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                float scale, int accumulate, void* stream) {
    if (!gspec || !gx) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (int rc = check_common(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample < n_fft) return fail(SE_ERR_UNSUPPORTED, "adjoint needs nsample >= n_fft");
    SynArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, 0.5f * scale, a.tb)) return rc;
    a.in = gspec; a.out = gx; a.nsample = (int)nsample; a.out_len = (int)nsample;
    a.nframe = (int)(1 + nsample / hop);
    a.b_lo = 0; a.b_hi = (int)((nsample + n_fft + hop - 1) / hop);
    a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, n_fft / hop);
    a.accumulate = accumulate; a.edge_scale = 2.0f;
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_synthesis<G, EMIT_ADJ>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_stft_bwd launch");
}

int se_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop,
                 int win_length, float scale, void* stream) {
    if (!spec || !y) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nframe <= 0 || length <= 0) return fail(SE_ERR_BAD_ARG, "nframe and length must be positive");
    if (int rc = check_common(rows, length, n_fft, hop, win_length)) return rc;
    if (!envelope_ok(n_fft, hop, win_length, false, nframe, n_fft / 2, n_fft / 2 + length, 1e-11))
        return fail(SE_ERR_ENVELOPE, "window overlap add min < 1e-11 (torch.istft raises the same)");
    SynArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, scale / (float)n_fft, a.tb)) return rc;
    a.in = spec; a.out = y; a.nsample = (int)(n_fft + hop * (nframe - 1)); a.out_len = (int)length;
    a.nframe = (int)nframe;
    a.b_lo = (n_fft / 2) / hop; a.b_hi = (int)((n_fft / 2 + length + hop - 1) / hop);
    a.nchunks = plan_synthesis(rows, a.b_hi - a.b_lo, n_fft / hop);
    a.accumulate = 0; a.edge_scale = 1.0f;
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_synthesis<G, EMIT_ISTFT>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_istft_fwd launch");
}

int se_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop,
                 int win_length, float scale, void* stream) {
    if (!gy || !gspec) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (nframe <= 0 || length <= 0) return fail(SE_ERR_BAD_ARG, "nframe and length must be positive");
    if (int rc = check_common(rows, length, n_fft, hop, win_length)) return rc;
    AnaArgs a{};
    if (int rc = get_tables(n_fft, hop, win_length, false, scale / (float)n_fft, a.tb)) return rc;
    a.in = gy; a.out = gspec; a.in_stride = length; a.seg_rows = 1; a.in_len = (int)length;
    a.nsample = (int)(n_fft + hop * (nframe - 1)); a.nframe = (int)nframe; a.pad = 0; a.edge_scale = 0.5f;
    plan_analysis(rows, a.nframe, a.gpc, a.nchunks);
    cudaError_t e;
    SE_DISPATCH_GEO(n_fft, hop, (e = run_analysis<G, LOAD_ENV>(a, rows, (cudaStream_t)stream)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_istft_bwd launch");
}

int se_mask_fwd(const float* spec, const float* mask, float* out, int64_t count, int mode, int pre_tanh, void* stream) {
    if (!spec || !mask || !out || count <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    if (mode < 0 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "mask mode must be REAL/E/C/R");
    int64_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaError_t e;
    SE_DISPATCH_MASK(mode, pre_tanh, (e = launch(k_mask_fwd_t<MODE, TANH>, (unsigned)blocks, 256, 0, (cudaStream_t)stream,
                                                 reinterpret_cast<const float2*>(spec), mask, reinterpret_cast<float2*>(out), count)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mask_fwd launch");
}

int se_mask_bwd(const float* spec, const float* mask, const float* gout, float* gmask, float* gspec, int64_t count,
                int mode, int pre_tanh, void* stream) {
    if (!spec || !mask || !gout || !gmask || count <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    if (mode < 0 || mode > 3) return fail(SE_ERR_UNSUPPORTED, "mask mode must be REAL/E/C/R");
    int64_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaError_t e;
    SE_DISPATCH_MASK(mode, pre_tanh, (e = launch(k_mask_bwd_t<MODE, TANH>, (unsigned)blocks, 256, 0, (cudaStream_t)stream,
                                                 reinterpret_cast<const float2*>(spec), mask,
                                                 reinterpret_cast<const float2*>(gout), gmask,
                                                 reinterpret_cast<float2*>(gspec), count)));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mask_bwd launch");
}

// ---------------------------------------------------------------- MR-STFT loss
static int loss_fwd_plan(int64_t rows, int64_t nsample, int r, int& gpc, int& nchunks) {
    const int64_t T = 1 + nsample / kRes[r][1];
    plan_analysis(rows, T, gpc, nchunks);
    return (int)(rows * nchunks);
}

// workspace layout: [per-CTA partial sums (double) for the 3 resolutions | |B| per resolution (float)]
static int64_t loss_partials_bytes(int64_t rows, int64_t nsample) {
    int64_t total = 0;
    for (int r = 0; r < 3; ++r) {
        int gpc, nchunks;
        total += (int64_t)loss_fwd_plan(rows, nsample, r, gpc, nchunks) * 3 * sizeof(double);
    }
    return (total + 255) / 256 * 256;
}
static int64_t loss_refmag_floats(int64_t rows, int64_t nsample, int r) {
    return rows * (kRes[r][0] / 2 + 1) * (1 + nsample / kRes[r][1]);
}

int64_t se_mrstft_workspace_bytes(int64_t rows, int64_t nsample) {
    int64_t total = loss_partials_bytes(rows, nsample);
    for (int r = 0; r < 3; ++r) total += loss_refmag_floats(rows, nsample, r) * (int64_t)sizeof(float);
    return total;
}

int se_mrstft_loss_fwd(const float* est, const float* ref, int64_t rows, int64_t nsample, double* sums,
                       void* workspace, void* stream) {
    if (!est || !ref || !sums || !workspace) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (rows <= 0 || nsample < 2048) return fail(SE_ERR_BAD_ARG, "need rows > 0 and nsample >= 2048");
    double* part = reinterpret_cast<double*>(workspace);
    float* refmag = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + loss_partials_bytes(rows, nsample));
    int nres[3] = {0, 0, 0};
    for (int r = 0; r < 3; ++r) {
        const int n = kRes[r][0], hop = kRes[r][1], win = kRes[r][2];
        if (int rc = check_common(rows, nsample, n, hop, win)) return rc;
        LossArgs a{};
        if (int rc = get_tables(n, hop, win, false, 0.5f, a.tb)) return rc;
        a.est = est; a.ref = ref; a.partials = part; a.refmag = refmag;
        a.nsample = (int)nsample; a.nframe = (int)(1 + nsample / hop);
        const int nctas = loss_fwd_plan(rows, nsample, r, a.gpc, a.nchunks);
        cudaError_t e;
        SE_DISPATCH_LOSS_GEO(n, (e = run_loss_fwd<G>(a, rows, (cudaStream_t)stream)));
        if (e != cudaSuccess) return cuda_fail(e, "se_mrstft_loss_fwd launch");
        nres[r] = nctas;
        part += (size_t)nctas * 3;
        refmag += loss_refmag_floats(rows, nsample, r);
    }
    cudaError_t e = launch(k_reduce_partials, 3u, 256u, 0, (cudaStream_t)stream,
                           (const double*)reinterpret_cast<double*>(workspace), nres[0], nres[1], nres[2], sums);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mrstft_loss_fwd reduce launch");
}

int se_mrstft_loss_value(const double* sums, int64_t global_rows, int64_t nsample, float* loss, void* stream) {
    if (!sums || !loss || global_rows <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty batch");
    double cnt[3];
    for (int r = 0; r < 3; ++r)
        cnt[r] = (double)global_rows * (kRes[r][0] / 2 + 1) * (double)(1 + nsample / kRes[r][1]);
    cudaError_t e = launch(k_loss_value, 1u, 32u, 0, (cudaStream_t)stream, sums, cnt[0], cnt[1], cnt[2], loss);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mrstft_loss_value launch");
}

int se_mrstft_loss_bwd(const float* est, const void* workspace, const double* sums, const float* gout, int64_t global_rows,
                       int64_t rows, int64_t nsample, float* g_est, void* stream) {
    if (!est || !workspace || !sums || !gout || !g_est) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (rows <= 0 || global_rows < rows || nsample < 2048) return fail(SE_ERR_BAD_ARG, "need 0 < rows <= global_rows, nsample >= 2048");
    const float* refmag = reinterpret_cast<const float*>(reinterpret_cast<const char*>(workspace) + loss_partials_bytes(rows, nsample));
    for (int r = 0; r < 3; ++r) {
        const int n = kRes[r][0], hop = kRes[r][1], win = kRes[r][2];
        if (int rc = check_common(rows, nsample, n, hop, win)) return rc;
        LossArgs a{};
        if (int rc = get_tables(n, hop, win, false, 0.5f, a.tb)) return rc;
        a.est = est; a.refmag = const_cast<float*>(refmag); a.g_est = g_est; a.sums = sums + 3 * r; a.gout = gout;
        a.nsample = (int)nsample; a.nframe = (int)(1 + nsample / hop);
        a.b_lo = 0; a.b_hi = (int)((nsample + n + hop - 1) / hop);
        a.nchunks = (a.b_hi + 12) / 13;          // single-group chunks: 16 - (OLA-1) blocks each, no carry
        a.accumulate = r > 0;
        a.inv_count = (float)(1.0 / ((double)global_rows * (n / 2 + 1) * (double)a.nframe));
        a.inv_res = 1.0f / 3.0f;
        cudaError_t e;
        SE_DISPATCH_LOSS_GEO(n, (e = run_loss_bwd<G>(a, rows, (cudaStream_t)stream)));
        if (e != cudaSuccess) return cuda_fail(e, "se_mrstft_loss_bwd launch");
        refmag += loss_refmag_floats(rows, nsample, r);
    }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------- fused enhance + DCCRN transforms
#include "se_capi_ext.inc"
