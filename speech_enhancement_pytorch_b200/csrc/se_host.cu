// se_host.cu -- constant-table cache, launch planning, error string (shared by every se_api_*.cu).
#include "se_host.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace se {

static thread_local std::string g_err;
const char* last_error() { return g_err.c_str(); }
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
    return fail(SE_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------ constant tables
struct TableKey {
    int dev, n, hop, win_len, front, window_id;
    uint32_t scale_bits;
    bool operator<(const TableKey& o) const {
        return std::tie(dev, n, hop, win_len, front, window_id, scale_bits) <
               std::tie(o.dev, o.n, o.hop, o.win_len, o.front, o.window_id, o.scale_bits);
    }
};
static std::mutex g_mu;
static std::map<TableKey, Tables> g_tables;

// Caller-supplied windows (DCCRN's ConvSTFT takes any scipy.signal.get_window type, src/model/dccrn.py:651-655): the
// host side computes the window once and registers its values; id 0 is the built-in periodic Hann.
static std::mutex g_win_mu;
static std::vector<std::vector<double>> g_windows;

int register_window(const double* values, int win_len) {
    if (!values || win_len < 2 || win_len > 8192) return fail(SE_ERR_BAD_ARG, "window: need 2 <= win_len <= 8192 values");
    std::lock_guard<std::mutex> lock(g_win_mu);
    for (size_t i = 0; i < g_windows.size(); ++i)
        if ((int)g_windows[i].size() == win_len && std::memcmp(g_windows[i].data(), values, sizeof(double) * win_len) == 0)
            return (int)i + 1;
    if (g_windows.size() >= 4096) return fail(SE_ERR_BAD_ARG, "too many distinct windows registered");
    g_windows.emplace_back(values, values + win_len);
    return (int)g_windows.size();
}

static bool host_window(int n, int win_len, bool front, std::vector<double>& w, int window_id = 0) {
    w.assign(n, 0.0);
    const int left = front ? 0 : (n - win_len) / 2;
    if (window_id > 0) {
        std::lock_guard<std::mutex> lock(g_win_mu);
        if (window_id > (int)g_windows.size() || (int)g_windows[window_id - 1].size() != win_len) return false;
        for (int j = 0; j < win_len; ++j) w[left + j] = g_windows[window_id - 1][j];
        return true;
    }
    const double two_pi = 6.283185307179586476925286766559;
    for (int j = 0; j < win_len; ++j) w[left + j] = 0.5 - 0.5 * std::cos(two_pi * j / win_len);
    return true;
}

bool host_window_values(int n, int win_len, bool front, std::vector<double>& w, int window_id) {
    return host_window(n, win_len, front, w, window_id);
}

// `scale` multiplies the window; front=true puts a short window at the start of the frame (DCCRN)
int get_tables(int n, int hop, int win_len, bool front, float scale, Tables& out, int window_id) {
    int dev = 0;
    cudaGetDevice(&dev);
    TableKey key{dev, n, hop, win_len, front ? 1 : 0, window_id, 0};
    std::memcpy(&key.scale_bits, &scale, 4);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) { out = it->second; return 0; }
    const int M = n / 2;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<double> w;
    if (!host_window(n, win_len, front, w, window_id)) return fail(SE_ERR_BAD_ARG, "unknown window id (or its length differs from win_length)");
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
static bool envelope_ok_uncached(int n, int hop, int win_len, bool front, int64_t T, int64_t lo, int64_t hi, double floor_);

bool envelope_ok(int n, int hop, int win_len, bool front, int64_t T, int64_t lo, int64_t hi, double floor_) {
    // the answer depends only on the configuration: remember it (the check runs on every iSTFT call)
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int, int64_t, int64_t, int64_t>, bool> seen;
    const auto key = std::make_tuple(n, hop, win_len, front ? 1 : 0, T, lo, hi);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = seen.find(key);
        if (it != seen.end()) return it->second;
    }
    const bool ok = envelope_ok_uncached(n, hop, win_len, front, T, lo, hi, floor_);
    std::lock_guard<std::mutex> lock(mu);
    if (seen.size() > 4096) seen.clear();
    seen[key] = ok;
    return ok;
}

static bool envelope_ok_uncached(int n, int hop, int win_len, bool front, int64_t T, int64_t lo, int64_t hi, double floor_) {
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

int engine_version() {
    auto read = [] { const char* e = std::getenv("SE_ENGINE"); return (e && e[0] == '2') ? 2 : ((e && e[0] == '3') ? 3 : 1); };
#ifdef SE_EMULATE
    return read();           // tests flip it between calls
#else
    static const int v = read();
    return v;
#endif
}

// SE_FR8=1: 8-frame groups (half the working set, twice the CTAs per SM) for the scalar engine's n <= 1024 kernels
int frames8() {
    auto read = [] { const char* e = std::getenv("SE_FR8"); return (e && e[0] == '1') ? 1 : 0; };
#ifdef SE_EMULATE
    return read();
#else
    static const int v = read();
    return v;
#endif
}

// ------------------------------------------------------------------ launch planning
static int g_target_ctas = 148 * 8;

void plan_analysis(int64_t rows, int64_t T, int& gpc, int& nchunks, int frames_per_group) {
    const int64_t ng = (T + frames_per_group - 1) / frames_per_group;
    static const int target = [] { const char* e = std::getenv("SE_TARGET_CTAS"); return e ? std::atoi(e) : 0; }();
    int64_t g = (rows * ng) / (target > 0 ? target : g_target_ctas);
    g = g < 1 ? 1 : (g > 8 ? 8 : g);
    gpc = (int)g;
    nchunks = (int)((ng + g - 1) / g);
}
// Pick the groups-per-chunk g that minimises (waves x g): a chunk of g groups emits 16 g - (ola-1)
// blocks (the first ola-1 frames are halo recompute), so larger g wastes less, but the grid must still
// fill the GPU in whole waves.  ctas_per_sm = resident CTAs of this kernel per SM.
int plan_synthesis(int64_t rows, int nb, int ola, int ctas_per_sm, int frames_per_group, int g_min, int g_max) {
    const int64_t slots = 148LL * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    int best_chunks = 1;
    double best_cost = 1e30;
    // SE_FORCE_GROUPS=g pins the choice (tests exercise the multi-group carry path on tiny inputs)
    const char* force = std::getenv("SE_FORCE_GROUPS");
    int g_lo = force ? std::atoi(force) : g_min, g_hi = force ? std::atoi(force) : g_max;
    g_lo = g_lo < g_min ? g_min : g_lo;
    g_hi = g_hi > g_max ? g_max : (g_hi < g_lo ? g_lo : g_hi);
    for (int g = g_lo; g <= g_hi; ++g) {
        const int cb_max = frames_per_group * g - (ola - 1);
        const int nchunks = (nb + cb_max - 1) / cb_max;
        const int cb = (nb + nchunks - 1) / nchunks;                 // even split (make_chunk)
        const int groups = (cb + ola - 1 + frames_per_group - 1) / frames_per_group;
        const int64_t waves = (rows * nchunks + slots - 1) / slots;
        const double cost = (double)waves * groups;
        if (cost < best_cost - 1e-9) { best_cost = cost; best_chunks = nchunks; }
    }
    return best_chunks < 1 ? 1 : best_chunks;
}

int check_common(int64_t rows, int64_t nsample, int n_fft, int hop, int win_length) {
    if (rows <= 0 || nsample <= 0) return fail(SE_ERR_BAD_ARG, "rows and nsample must be positive");
    if (n_fft != 512 && n_fft != 1024 && n_fft != 2048)
        return fail(SE_ERR_UNSUPPORTED, "n_fft must be 512, 1024 or 2048 (no fallback path exists)");
    if (hop * 4 != n_fft && hop * 2 != n_fft)
        return fail(SE_ERR_UNSUPPORTED, "hop_length must be n_fft/4 or n_fft/2");
    if (win_length < 2 || win_length > n_fft) return fail(SE_ERR_UNSUPPORTED, "need 2 <= win_length <= n_fft");
    if (rows * ((nsample / hop + 1 + 15) / 16) > 0x7fffffffLL) return fail(SE_ERR_BAD_ARG, "problem too large for one launch");
    return 0;
}

}  // namespace se

extern "C" int se_version(void) { return 100; }
extern "C" const char* se_last_error(void) { return se::last_error(); }
