// se_api_generic.cu -- host side of the general-geometry path (se_generic.cuh): tables, scratch, launches.
// The public entry points (se_stft_fwd, se_istft_fwd, their adjoints, se_conv_*_w) route here when the geometry is not
// one the tuned engine is compiled for; the C-ABI does not change.
#include "se_host.h"
#include "se_generic.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <utility>
#include <vector>

namespace se {

bool geometry_tuned(int n_fft, int hop) {
    return (n_fft == 512 || n_fft == 1024 || n_fft == 2048) && (hop * 4 == n_fft || hop * 2 == n_fft);
}

int check_general(int64_t rows, int64_t nsample, int n_fft, int hop, int win_length) {
    if (rows <= 0 || nsample <= 0) return fail(SE_ERR_BAD_ARG, "rows and nsample must be positive");
    if (n_fft < 8 || n_fft > 8192 || (n_fft & 1)) return fail(SE_ERR_UNSUPPORTED, "n_fft must be even, 8 .. 8192");
    if (hop < 1 || hop > n_fft) return fail(SE_ERR_UNSUPPORTED, "need 1 <= hop_length <= n_fft");
    if (win_length < 2 || win_length > n_fft) return fail(SE_ERR_UNSUPPORTED, "need 2 <= win_length <= n_fft");
    if (nsample + 2LL * n_fft > 0x7fffffffLL || rows * (nsample / hop + 2) > 0x7fffffffLL)
        return fail(SE_ERR_BAD_ARG, "problem too large for one launch");
    return 0;
}

// ------------------------------------------------------------------ tables
struct GenKey {
    int dev, n, win_len, front, window_id;
    uint32_t scale_bits;
    bool operator<(const GenKey& o) const {
        return std::tie(dev, n, win_len, front, window_id, scale_bits) < std::tie(o.dev, o.n, o.win_len, o.front, o.window_id, o.scale_bits);
    }
};
static std::mutex g_gen_mu;
static std::map<GenKey, GenTables> g_gen_tables;

// in-place radix-2 FFT (forward, e^{-i}) in double: builds the Bluestein filter spectrum once per geometry
static void host_fft(std::vector<double>& re, std::vector<double>& im) {
    const size_t L = re.size();
    for (size_t i = 1, j = 0; i < L; ++i) {
        size_t bit = L >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    const double two_pi = 6.283185307179586476925286766559;
    for (size_t len = 2; len <= L; len <<= 1) {
        for (size_t i = 0; i < L; i += len) {
            for (size_t k = 0; k < len / 2; ++k) {
                const double ang = -two_pi * (double)k / (double)len, wr = std::cos(ang), wi = std::sin(ang);
                const size_t a = i + k, b = i + k + len / 2;
                const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
        }
    }
}

static int get_gen_tables(int n, int win_len, bool front, float scale, int window_id, GenTables& out) {
    int dev = 0;
    cudaGetDevice(&dev);
    GenKey key{dev, n, win_len, front ? 1 : 0, window_id, 0};
    std::memcpy(&key.scale_bits, &scale, 4);
    std::lock_guard<std::mutex> lock(g_gen_mu);
    auto it = g_gen_tables.find(key);
    if (it != g_gen_tables.end()) { out = it->second; return 0; }
    std::vector<double> w;
    if (!host_window_values(n, win_len, front, w, window_id))
        return fail(SE_ERR_BAD_ARG, "unknown window id (or its length differs from win_length)");
    const int M = n / 2;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float> win(n), w2(n);
    const int L = gen_bluestein_len(n);
    const int TW = L ? L : M;                        // the FFT that runs is L-point (Bluestein) or M-point
    std::vector<float2> tw(TW), twn(M + 1);
    for (int j = 0; j < n; ++j) {
        win[j] = (float)(w[j] * (double)scale);
        const float wf = (float)w[j];
        w2[j] = wf * wf;
    }
    for (int k = 0; k < TW; ++k) tw[k] = make_float2((float)std::cos(two_pi * k / TW), (float)-std::sin(two_pi * k / TW));
    for (int k = 0; k <= M; ++k) twn[k] = make_float2((float)std::cos(two_pi * k / n), (float)-std::sin(two_pi * k / n));
    float *d_win = nullptr, *d_w2 = nullptr;
    float2 *d_tw = nullptr, *d_twn = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&d_win, n * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&d_w2, n * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&d_tw, TW * sizeof(float2))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&d_twn, (M + 1) * sizeof(float2))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cudaMemcpy(d_win, win.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d_w2, w2.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d_tw, tw.data(), TW * sizeof(float2), cudaMemcpyHostToDevice);
    e = cudaMemcpy(d_twn, twn.data(), (M + 1) * sizeof(float2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(tables)");
    out = GenTables{d_win, d_w2, d_tw, d_twn, nullptr, nullptr, 0};
    if (L) {
        // chirp c_j = e^{i pi j^2 / M} (j^2 reduced mod 2M in integers) and the L-point FFT of the filter, in double
        std::vector<float2> chirp(M), bf(L);
        std::vector<double> br(L, 0.0), bi(L, 0.0);
        for (int j = 0; j < M; ++j) {
            const long long r = ((long long)j * j) % (2LL * M);
            const double ang = two_pi * 0.5 * (double)r / (double)M;
            chirp[j] = make_float2((float)std::cos(ang), (float)std::sin(ang));
            br[j] = std::cos(ang); bi[j] = std::sin(ang);
            if (j) { br[L - j] = br[j]; bi[L - j] = bi[j]; }
        }
        host_fft(br, bi);
        for (int k = 0; k < L; ++k) bf[k] = make_float2((float)(br[k] / L), (float)(bi[k] / L));
        float2 *d_chirp = nullptr, *d_bf = nullptr;
        if ((e = cudaMalloc((void**)&d_chirp, M * sizeof(float2))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        if ((e = cudaMalloc((void**)&d_bf, L * sizeof(float2))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        cudaMemcpy(d_chirp, chirp.data(), M * sizeof(float2), cudaMemcpyHostToDevice);
        e = cudaMemcpy(d_bf, bf.data(), L * sizeof(float2), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(tables)");
        out.chirp = d_chirp; out.bfft = d_bf; out.L = L;
    }
    g_gen_tables[key] = out;
    return 0;
}

// ------------------------------------------------------------------ launches
static int gen_warps(int n) {
    const size_t per_warp = gen_smem_bytes(n, 1);                 // at most 128 KB (Bluestein L = 8192 for n_fft > 4097)
    const size_t fit = (96 * 1024) / per_warp;
    int w = 8;                                                    // a power of two: the kernels index frames with shifts
    while (w > 1 && (size_t)w > fit) w >>= 1;
    return w;
}

template <int LMODE>
static int run_gen_analysis(GenArgs a, int64_t rows, cudaStream_t st, const char* what) {
    const int W = gen_warps(a.n);
    const int64_t cpr = (a.nframe + W - 1) / W;
    if (rows * cpr > 0x7fffffffLL) return fail(SE_ERR_BAD_ARG, "problem too large for one launch");
    const cudaError_t e = a.tb.L ? launch_ex(false, k_gen_analysis<LMODE, true>, (unsigned)(rows * cpr), 32u * W, gen_smem_bytes(a.n, W), st, a)
                                 : launch_ex(false, k_gen_analysis<LMODE, false>, (unsigned)(rows * cpr), 32u * W, gen_smem_bytes(a.n, W), st, a);
    return e == cudaSuccess ? 0 : cuda_fail(e, what);
}

#ifndef SE_EMULATE
// Stream-ordered scratch from a pool of our own that keeps its memory across synchronisations (the default pool hands
// everything back to the driver at every sync, which would make each call pay a fresh allocation).
static cudaError_t scratch_alloc(float** p, size_t bytes, cudaStream_t st) {
    static std::mutex mu;
    static std::map<int, cudaMemPool_t> pools;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = pools.find(dev);
        if (it == pools.end()) {
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaError_t e = cudaMemPoolCreate(&pool, &props);
            if (e != cudaSuccess) return e;
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            pools[dev] = pool;
        } else {
            pool = it->second;
        }
    }
    return cudaMallocFromPoolAsync((void**)p, bytes, pool, st);
}
#endif

// spectrum -> frames (scratch) -> overlap-add, in row batches that bound the scratch
template <int EMODE>
static int run_gen_synthesis(GenArgs a, int64_t rows, int64_t spec_row_floats, cudaStream_t st, const char* what) {
    const int W = gen_warps(a.n);
    const int64_t cpr = (a.nframe + W - 1) / W;
    const int64_t frame_floats = (int64_t)a.nframe * a.f_len;
    // <= 2 GiB of scratch (SE_GEN_SCRATCH_FLOATS overrides the cap: tests force the row-batch loop on tiny inputs)
    const char* cap_env = std::getenv("SE_GEN_SCRATCH_FLOATS");
    const int64_t cap = cap_env ? std::atoll(cap_env) : (int64_t)(1LL << 29);
    int64_t batch = cap / (frame_floats > 0 ? frame_floats : 1);
    batch = batch < 1 ? 1 : (batch > rows ? rows : batch);
    if (batch > 65535) batch = 65535;                                                  // gridDim.y
    if (batch * cpr > 0x7fffffffLL) batch = 0x7fffffffLL / cpr;
    float* scratch = nullptr;
#ifdef SE_EMULATE
    cudaError_t e = cudaMalloc((void**)&scratch, (size_t)batch * frame_floats * sizeof(float));
#else
    cudaError_t e = scratch_alloc(&scratch, (size_t)batch * frame_floats * sizeof(float), st);
#endif
    if (e != cudaSuccess) return cuda_fail(e, "scratch allocation (general-geometry synthesis)");
    const float* spec = a.in;
    float* out = a.out;
    for (int64_t r0 = 0; r0 < rows && e == cudaSuccess; r0 += batch) {
        const int64_t nr = rows - r0 < batch ? rows - r0 : batch;
        GenArgs f = a;
        f.in = spec + r0 * spec_row_floats;
        f.out = scratch;
        e = a.tb.L ? launch_ex(false, k_gen_frames<true>, (unsigned)(nr * cpr), 32u * W, gen_smem_bytes(a.n, W), st, f)
                   : launch_ex(false, k_gen_frames<false>, (unsigned)(nr * cpr), 32u * W, gen_smem_bytes(a.n, W), st, f);
        if (e != cudaSuccess) break;
        GenArgs o = a;
        o.in = scratch;
        o.out = out + r0 * a.out_len;
#ifdef SE_EMULATE
        emu::launch(dim3((a.out_len + 255) / 256, (unsigned)nr), dim3(256), 0, [&]() { k_gen_ola<EMODE>(o); });
#else
        k_gen_ola<EMODE><<<dim3((a.out_len + 255) / 256, (unsigned)nr), 256, 0, st>>>(o);
        e = cudaGetLastError();
#endif
    }
#ifdef SE_EMULATE
    cudaFree(scratch);
#else
    const cudaError_t e2 = cudaFreeAsync(scratch, st);
    if (e == cudaSuccess) e = e2;
#endif
    return e == cudaSuccess ? 0 : cuda_fail(e, what);
}

static void support(int n, int win_len, bool front, int& f_lo, int& f_len) {
    const int left = front ? 0 : (n - win_len) / 2;
    f_lo = left & ~1;
    const int hi = (left + win_len + 1) & ~1;
    f_len = (hi > n ? n : hi) - f_lo;
}

// ------------------------------------------------------------------ torch.stft / torch.istft convention
int gen_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length, float scale,
                 cudaStream_t st) {
    if (int rc = check_general(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    GenArgs a{};
    if (int rc = get_gen_tables(n_fft, win_length, false, scale, 0, a.tb)) return rc;
    a.in = x; a.out = spec; a.n = n_fft; a.hop = hop; a.nframe = (int)(1 + nsample / hop);
    a.in_stride = nsample; a.in_len = (int)nsample; a.pad = n_fft / 2; a.edge_w = a.mid_w = 1.0f;
    return run_gen_analysis<GEN_REFLECT>(a, rows, st, "se_stft_fwd (general geometry) launch");
}

int gen_stft_bwd(const float* gspec, float* gx, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length, float scale,
                 int accumulate, cudaStream_t st) {
    if (int rc = check_general(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample <= n_fft / 2) return fail(SE_ERR_BAD_ARG, "reflect padding needs nsample > n_fft/2");
    GenArgs a{};
    if (int rc = get_gen_tables(n_fft, win_length, false, scale, 0, a.tb)) return rc;
    a.in = gspec; a.out = gx; a.n = n_fft; a.hop = hop; a.nframe = (int)(1 + nsample / hop);
    a.pad = n_fft / 2; a.edge_w = 1.0f; a.mid_w = 0.5f;          // d/dx of sum_k: each of the F bins once (no Hermitian doubling)
    support(n_fft, win_length, false, a.f_lo, a.f_len);
    a.out_len = (int)nsample; a.nsample = (int)nsample; a.accumulate = accumulate;
    return run_gen_synthesis<GEN_OLA_ADJ>(a, rows, (int64_t)(n_fft / 2 + 1) * a.nframe * 2, st, "se_stft_bwd (general geometry) launch");
}

// center=False (torch.stft without padding, src/evaluate.py:116 passes config.center): frame t is x[t hop : t hop + n_fft]
int gen_stft_nocenter_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                          float scale, cudaStream_t st) {
    if (int rc = check_general(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample < n_fft) return fail(SE_ERR_BAD_ARG, "center=False needs nsample >= n_fft");
    GenArgs a{};
    if (int rc = get_gen_tables(n_fft, win_length, false, scale, 0, a.tb)) return rc;
    a.in = x; a.out = spec; a.n = n_fft; a.hop = hop; a.nframe = (int)(1 + (nsample - n_fft) / hop);
    a.in_stride = nsample; a.in_len = (int)nsample; a.pad = 0; a.edge_w = a.mid_w = 1.0f;
    return run_gen_analysis<GEN_ZEROPAD>(a, rows, st, "se_stft_nocenter_fwd launch");
}

int gen_stft_nocenter_bwd(const float* gspec, float* gx, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                          float scale, int accumulate, cudaStream_t st) {
    if (int rc = check_general(rows, nsample, n_fft, hop, win_length)) return rc;
    if (nsample < n_fft) return fail(SE_ERR_BAD_ARG, "center=False needs nsample >= n_fft");
    GenArgs a{};
    if (int rc = get_gen_tables(n_fft, win_length, false, scale, 0, a.tb)) return rc;
    a.in = gspec; a.out = gx; a.n = n_fft; a.hop = hop; a.nframe = (int)(1 + (nsample - n_fft) / hop);
    a.pad = 0; a.edge_w = 1.0f; a.mid_w = 0.5f;                  // no padding: the reflect fold of the adjoint is empty
    support(n_fft, win_length, false, a.f_lo, a.f_len);
    a.out_len = (int)nsample; a.nsample = (int)nsample; a.accumulate = accumulate;
    return run_gen_synthesis<GEN_OLA_ADJ>(a, rows, (int64_t)(n_fft / 2 + 1) * a.nframe * 2, st, "se_stft_nocenter_bwd launch");
}

int gen_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop, int win_length,
                  float scale, cudaStream_t st) {
    if (nframe <= 0 || length <= 0) return fail(SE_ERR_BAD_ARG, "nframe and length must be positive");
    if (int rc = check_general(rows, length, n_fft, hop, win_length)) return rc;
    if (!envelope_ok(n_fft, hop, win_length, false, nframe, n_fft / 2, n_fft / 2 + length, 1e-11))
        return fail(SE_ERR_ENVELOPE, "window overlap add min < 1e-11 (torch.istft raises the same)");
    GenArgs a{};
    if (int rc = get_gen_tables(n_fft, win_length, false, scale / (float)n_fft, 0, a.tb)) return rc;
    a.in = spec; a.out = y; a.n = n_fft; a.hop = hop; a.nframe = (int)nframe;
    a.pad = n_fft / 2; a.edge_w = a.mid_w = 1.0f; a.env_eps = 0.f;
    support(n_fft, win_length, false, a.f_lo, a.f_len);
    a.out_len = (int)length;
    return run_gen_synthesis<GEN_OLA_ISTFT>(a, rows, (int64_t)(n_fft / 2 + 1) * nframe * 2, st, "se_istft_fwd (general geometry) launch");
}

int gen_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t length, int n_fft, int hop, int win_length,
                  float scale, cudaStream_t st) {
    if (nframe <= 0 || length <= 0) return fail(SE_ERR_BAD_ARG, "nframe and length must be positive");
    if (int rc = check_general(rows, length, n_fft, hop, win_length)) return rc;
    GenArgs a{};
    if (int rc = get_gen_tables(n_fft, win_length, false, scale / (float)n_fft, 0, a.tb)) return rc;
    a.in = gy; a.out = gspec; a.n = n_fft; a.hop = hop; a.nframe = (int)nframe;
    a.in_stride = length; a.in_len = (int)length; a.pad = n_fft / 2; a.edge_w = 1.0f; a.mid_w = 2.0f; a.env_eps = 0.f;
    return run_gen_analysis<GEN_ENV>(a, rows, st, "se_istft_bwd (general geometry) launch");
}

// ------------------------------------------------------------------ DCCRN convention (window at the front of the frame)
static int conv_general_check(int64_t rows, int win_len, int win_inc, int fft_len) {
    if (rows <= 0) return fail(SE_ERR_BAD_ARG, "empty tensor");
    if (fft_len < 8 || fft_len > 8192 || (fft_len & 1))
        return fail(SE_ERR_UNSUPPORTED, "ConvSTFT / ConviSTFT: fft_len must be even, 8 .. 8192");
    if (win_inc < 1 || win_len < win_inc || win_len > fft_len)
        return fail(SE_ERR_UNSUPPORTED, "ConvSTFT / ConviSTFT: need 1 <= win_inc <= win_len <= fft_len");
    return 0;
}

int gen_conv_stft_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int win_len, int win_inc, int fft_len,
                      int window_id, cudaStream_t st) {
    if (int rc = conv_general_check(rows, win_len, win_inc, fft_len)) return rc;
    const int pad = win_len - win_inc;
    const int64_t T = (nsample + 2 * pad - win_len) / win_inc + 1;
    if (nsample <= 0 || T <= 0) return fail(SE_ERR_BAD_ARG, "ConvSTFT: input shorter than one frame");
    if (rows * (T + 8) > 0x7fffffffLL) return fail(SE_ERR_BAD_ARG, "problem too large for one launch");
    GenArgs a{};
    if (int rc = get_gen_tables(fft_len, win_len, true, 1.0f, window_id, a.tb)) return rc;
    a.in = x; a.out = spec; a.n = fft_len; a.hop = win_inc; a.nframe = (int)T;
    a.in_stride = nsample; a.in_len = (int)nsample; a.pad = pad; a.planar = 1; a.edge_w = a.mid_w = 1.0f;
    return run_gen_analysis<GEN_ZEROPAD>(a, rows, st, "se_conv_stft_fwd (general geometry) launch");
}

static int conv_inverse_args(GenArgs& a, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc, int fft_len,
                             int window_id) {
    if (int rc = conv_general_check(rows, win_len, win_inc, fft_len)) return rc;
    if (nframe <= 0 || out_len <= 0) return fail(SE_ERR_BAD_ARG, "empty tensor");
    const int64_t total = win_len + (int64_t)win_inc * (nframe - 1);
    const int pad = win_len - win_inc;
    if (out_len > total - pad) return fail(SE_ERR_BAD_ARG, "ConviSTFT: out_len exceeds the overlap-added signal");
    if (rows * (nframe + 8) > 0x7fffffffLL || total > 0x7fffffffLL) return fail(SE_ERR_BAD_ARG, "problem too large for one launch");
    // frame = pinv-corrected C2R x w x 2/n (closed form in se_conv.cuh)
    if (int rc = get_gen_tables(fft_len, win_len, true, 2.0f / (float)fft_len, window_id, a.tb)) return rc;
    a.n = fft_len; a.hop = win_inc; a.nframe = (int)nframe; a.pad = pad; a.planar = 1; a.env_eps = 1e-8f;
    a.parity_len = win_len;
    a.inv_even = 1.0f / (float)(fft_len / 2 + (win_len + 1) / 2);
    a.inv_odd = 1.0f / (float)(fft_len / 2 + win_len / 2);
    support(fft_len, win_len, true, a.f_lo, a.f_len);
    return 0;
}

int gen_conv_istft_fwd(const float* spec, float* y, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc,
                       int fft_len, int window_id, cudaStream_t st) {
    GenArgs a{};
    if (int rc = conv_inverse_args(a, rows, nframe, out_len, win_len, win_inc, fft_len, window_id)) return rc;
    a.in = spec; a.out = y; a.out_len = (int)out_len;
    a.edge_w = 1.0f; a.mid_w = 0.5f;                             // v[j] = Re sum_{k=0}^{n/2} Y_k e^{+i theta}: every bin once
    return run_gen_synthesis<GEN_OLA_ISTFT>(a, rows, (int64_t)(fft_len / 2 + 1) * nframe * 2, st, "se_conv_istft_fwd (general geometry) launch");
}

int gen_conv_istft_bwd(const float* gy, float* gspec, int64_t rows, int64_t nframe, int64_t out_len, int win_len, int win_inc,
                       int fft_len, int window_id, cudaStream_t st) {
    GenArgs a{};
    if (int rc = conv_inverse_args(a, rows, nframe, out_len, win_len, win_inc, fft_len, window_id)) return rc;
    a.in = gy; a.out = gspec; a.in_stride = out_len; a.in_len = (int)out_len;
    a.edge_w = a.mid_w = 1.0f;
    return run_gen_analysis<GEN_ENV>(a, rows, st, "se_conv_istft_bwd (general geometry) launch");
}

}  // namespace se

extern "C" int se_stft_nocenter_fwd(const float* x, float* spec, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                                    float scale, void* stream) {
    if (!x || !spec) return se::fail(SE_ERR_BAD_ARG, "null pointer");
    return se::gen_stft_nocenter_fwd(x, spec, rows, nsample, n_fft, hop, win_length, scale, (cudaStream_t)stream);
}
extern "C" int se_stft_nocenter_bwd(const float* gspec, float* gx, int64_t rows, int64_t nsample, int n_fft, int hop, int win_length,
                                    float scale, int accumulate, void* stream) {
    if (!gspec || !gx) return se::fail(SE_ERR_BAD_ARG, "null pointer");
    return se::gen_stft_nocenter_bwd(gspec, gx, rows, nsample, n_fft, hop, win_length, scale, accumulate, (cudaStream_t)stream);
}
extern "C" int se_geometry_tuned(int n_fft, int hop) { return se::geometry_tuned(n_fft, hop) ? 1 : 0; }
extern "C" int se_conv_geometry_tuned(int win_len, int win_inc, int fft_len) {
    return (fft_len == 512 && win_inc == 100 && win_len <= 4 * win_inc && win_len >= win_inc && !(win_len & 1)) ? 1 : 0;
}
