// se_generic.cuh -- the general-geometry transform path: any power-of-two n_fft (8 .. 8192) and -- through Bluestein's
// chirp-z over a power-of-two length -- any even n_fft up to 4096 (the reference's commented CRN setting n_fft = 320,
// src/conf/config.yaml:78-80; the common 400-point speech front end); any hop 1 .. n_fft, any
// win_length <= n_fft, window centred (torch.stft, src/evaluate.py:109-119) or at the front of the frame (DCCRN's
// ConvSTFT / ConviSTFT with any win_len / win_inc / fft_len, src/model/dccrn.py:649-747).
//
// The tuned engine (se_fft.cuh) is compiled per (n_fft, hop) and only exists for the geometries the reference's configs
// use; everything else the reference's signatures accept lands here.  One WARP owns one frame:
//   * real n-point transform = complex M = n/2 point Stockham autosort FFT (radix 4, one radix-2 pass when log2 M is
//     odd), ping-pong between two shared-memory buffers, twiddles from a global table (L1-resident);
//   * a CTA is W consecutive frames of one row, so the spectrum is read / written through shared memory with the frame
//     index fastest (W x 8 bytes contiguous per bin) although the layout [rows, F, T, 2] has bins strided by T;
//   * synthesis writes windowed frames to a scratch [rows, T, flen] and a second kernel overlap-adds them by output
//     sample (deterministic, no atomics), divides by the window envelope or folds the reflect padding (STFT adjoint).
// Nothing here is tuned past coalescing; the cost model is ~3x the tuned engine's traffic and ~2x its time.
#pragma once
#include "se_platform.h"

namespace se {

enum { GEN_REFLECT = 0, GEN_ZEROPAD = 1, GEN_ENV = 2 };
enum { GEN_OLA_ISTFT = 0, GEN_OLA_ADJ = 1 };

struct GenTables {
    const float* win;     // [n] window x scale, zero outside its support
    const float* w2;      // [n] squared (unscaled, fp32-rounded) window: overlap-add envelope
    const float2* tw;     // [M] e^{-2 pi i k / M}
    const float2* twn;    // [M + 1] e^{-2 pi i k / n}
    // n_fft not a power of two (even n): the M = n/2 point complex DFT runs as Bluestein's chirp-z over a power-of-two
    // length L >= 2M - 1; tw is then [L]
    const float2* chirp;  // [M] c_j = e^{+i pi j^2 / M}
    const float2* bfft;   // [L] FFT_L of the chirp filter b (b_m = b_{L-m} = c_m, m < M), pre-divided by L
    int L;                // 0: power-of-two path
};

struct GenArgs {
    GenTables tb;
    const float* in;
    float* out;
    int n, hop, nframe;       // n_fft, hop, T
    int64_t in_stride;        // analysis: floats between input rows
    int in_len;               // analysis: valid input samples (REFLECT / ZEROPAD: N; ENV: length of gy rows)
    int pad;                  // REFLECT: n/2; ZEROPAD: zeros in front; ENV: samples dropped in front of the natural signal
    int planar;               // spectrum layout: 0 [rows, F, T, 2]; 1 [rows, 2F, T] (DCCRN)
    float edge_w, mid_w;      // weight of the DC / Nyquist bins and of the interior bins (output of analysis, input of synthesis)
    int parity_len;           // DCCRN pinv correction over the first parity_len samples of the frame; 0 = none
    float inv_even, inv_odd;  // 1 / (n/2 + #even), 1 / (n/2 + #odd)
    float env_eps;            // ENV / OLA_ISTFT: added to the envelope (1e-8 for ConviSTFT, 0 for torch.istft)
    int f_lo, f_len;          // synthesis: samples [f_lo, f_lo + f_len) of each frame are stored in the scratch (window support)
    // overlap-add
    int out_len;              // samples per output row
    int nsample;              // ADJ: N (the reflect fold needs it)
    int accumulate;
};

__device__ __forceinline__ float2 g_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 g_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 g_mul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
__device__ __forceinline__ float2 g_conj(float2 a) { return make_float2(a.x, -a.y); }

// M-point complex FFT of one warp's buffer: forward e^{-i}, inverse e^{+i} (unnormalised).  Returns the buffer that
// holds the natural-order result.  tw[k] = e^{-2 pi i k / M}.
template <bool INV>
__device__ __forceinline__ float2* warp_fft(float2* x, float2* y, const float2* __restrict__ tw, int M, int lane) {
    int n = M, ls = 0;                               // sub-transform length, log2 of the stride
    while (n > 1) {
        const int s = 1 << ls;
        if ((n & 3) == 0) {
            const int n1 = n >> 2, cnt = n1 << ls;   // M / 4 butterflies
            for (int i = lane; i < cnt; i += 32) {
                const int p = i >> ls, q = i & (s - 1);
                const float2 a = x[i], b = x[i + cnt], c = x[i + 2 * cnt], d = x[i + 3 * cnt];
                // one table load; w^2 and w^3 by two complex products (cheaper than two more dependent L1 loads)
                float2 w1 = __ldg(tw + (p << ls));
                if (INV) w1.y = -w1.y;
                const float2 w2 = g_mul(w1, w1), w3 = g_mul(w2, w1);
                const float2 apc = g_add(a, c), amc = g_sub(a, c), bpd = g_add(b, d), bmd = g_sub(b, d);
                // forward: -i (b - d); inverse: +i (b - d)
                const float2 jb = INV ? make_float2(-bmd.y, bmd.x) : make_float2(bmd.y, -bmd.x);
                float2* o = y + q + ((4 * p) << ls);
                o[0] = g_add(apc, bpd);
                o[s] = g_mul(g_add(amc, jb), w1);
                o[2 * s] = g_mul(g_sub(apc, bpd), w2);
                o[3 * s] = g_mul(g_sub(amc, jb), w3);
            }
            n >>= 2; ls += 2;
        } else {
            const int m = n >> 1, cnt = m << ls;
            for (int i = lane; i < cnt; i += 32) {
                const int p = i >> ls, q = i & (s - 1);
                const float2 a = x[i], b = x[i + cnt];
                float2 w = __ldg(tw + (p << ls));
                if (INV) w.y = -w.y;
                float2* o = y + q + ((2 * p) << ls);
                o[0] = g_add(a, b);
                o[s] = g_mul(g_sub(a, b), w);
            }
            n >>= 1; ls += 1;
        }
        __syncwarp();
        float2* t = x; x = y; y = t;
    }
    return x;
}

__host__ __device__ inline int gen_fft_passes(int M) {
    int passes = 0;
    for (int q = M; q > 1; q = (q & 3) == 0 ? q >> 2 : q >> 1) ++passes;
    return passes;
}

// overlap-add envelope at natural position i: sum of w2 over the frames that cover it
__device__ __forceinline__ float gen_env_at(const float* __restrict__ w2, int n, int hop, int T, int i) {
    int t_hi = i / hop;
    t_hi = t_hi < T - 1 ? t_hi : T - 1;
    int t_lo = i - n + 1;
    t_lo = t_lo <= 0 ? 0 : (t_lo + hop - 1) / hop;
    float e = 0.f;
    for (int t = t_hi; t >= t_lo; --t) e += __ldg(w2 + i - t * hop);     // same order as the tuned engine: q = 0, 1, ...
    return e;
}

template <int LMODE>
__device__ __forceinline__ float gen_sample(const GenArgs& a, const float* __restrict__ src, int i) {
    const int j = i - a.pad;
    if (LMODE == GEN_REFLECT) {
        int r = j < 0 ? -j : j;
        r = r >= a.in_len ? 2 * (a.in_len - 1) - r : r;
        return (r >= 0 && r < a.in_len) ? __ldg(src + r) : 0.f;
    } else if (LMODE == GEN_ZEROPAD) {
        return (j >= 0 && j < a.in_len) ? __ldg(src + j) : 0.f;
    } else {
        if (j < 0 || j >= a.in_len) return 0.f;
        const float e = gen_env_at(a.tb.w2, a.n, a.hop, a.nframe, i) + a.env_eps;
        return __ldg(src + j) / e;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// z[m] = (v[2m], v[2m+1]): subtract the parity means over samples j < parity_len (DCCRN's pinv closed form, se_conv.cuh)
__device__ __forceinline__ void parity_correct(float2* z, int M, int lane, int parity_len, float inv_even, float inv_odd) {
    float se = 0.f, so = 0.f;
    for (int m = lane; m < M; m += 32) {
        const float2 v = z[m];
        if (2 * m < parity_len) se += v.x;
        if (2 * m + 1 < parity_len) so += v.y;
    }
    se = warp_sum(se) * inv_even;
    so = warp_sum(so) * inv_odd;
    for (int m = lane; m < M; m += 32) {
        float2 v = z[m];
        if (2 * m < parity_len) v.x -= se;
        if (2 * m + 1 < parity_len) v.y -= so;
        z[m] = v;
    }
    __syncwarp();
}

// per-warp buffers: two of (M + 1) float2, padded to an even count of float2 (16-byte aligned rows)
// Bluestein length for n_fft that is not a power of two (0 for powers of two): power of two >= 2 (n/2) - 1
__host__ __device__ inline int gen_bluestein_len(int n) {
    if ((n & (n - 1)) == 0) return 0;
    int L = 1;
    while (L < n - 1) L <<= 1;
    return L;
}
__host__ __device__ inline int gen_buf_len(int n) {
    const int L = gen_bluestein_len(n);
    return L ? L : ((n / 2 + 2) & ~1);            // L >= n - 1 >= n/2 + 1 entries for the bins
}
inline size_t gen_smem_bytes(int n, int warps) { return (size_t)warps * 2 * gen_buf_len(n) * sizeof(float2); }

// ------------------------------------------------------------------ analysis: waveform-like -> spectrum
// grid = rows x ceil(T / W), block = 32 W
// M-point complex DFT of x (forward e^{-i}, inverse e^{+i}, unnormalised); returns the buffer holding the result.
// BZ = false: M is a power of two, Stockham FFT.  BZ = true: Bluestein -- the DFT as a circular convolution of length
// L >= 2M - 1 with the chirp c_j = e^{i pi j^2 / M}: X_k = conj(c_k) sum_j (x_j conj(c_j)) c_{k-j}; the filter's spectrum
// (pre-divided by L) is a table, so it costs two L-point FFTs.  The inverse is conj(DFT(conj x)).  Both FFTs take the
// same number of ping-pong passes, so the Bluestein result is always back in x.
template <bool INV, bool BZ>
__device__ __forceinline__ float2* gen_cdft(float2* x, float2* y, const GenTables& tb, int M, int lane) {
    if (!BZ) return warp_fft<INV>(x, y, tb.tw, M, lane);
    const int L = tb.L;
    for (int j = lane; j < L; j += 32) {
        float2 v = make_float2(0.f, 0.f);
        if (j < M) {
            v = x[j];
            if (INV) v.y = -v.y;
            v = g_mul(v, g_conj(__ldg(tb.chirp + j)));
        }
        x[j] = v;
    }
    __syncwarp();
    float2* z = warp_fft<false>(x, y, tb.tw, L, lane);
    float2* o = (z == x) ? y : x;
    for (int j = lane; j < L; j += 32) z[j] = g_mul(z[j], __ldg(tb.bfft + j));
    __syncwarp();
    warp_fft<true>(z, o, tb.tw, L, lane);
    for (int k = lane; k < M; k += 32) {
        float2 v = g_mul(x[k], g_conj(__ldg(tb.chirp + k)));
        if (INV) v.y = -v.y;
        x[k] = v;
    }
    __syncwarp();
    return x;
}

template <int LMODE, bool BZ>
__global__ void k_gen_analysis(GenArgs a) {
    SE_SMEM_DECL;
    pdl_wait();
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.n, M = n >> 1, BL = gen_buf_len(n);
    const int cpr = (a.nframe + W - 1) / W;                    // CTAs per row
    const int row = blockIdx.x / cpr, t0 = (blockIdx.x - row * cpr) * W;
    float2* bufs = reinterpret_cast<float2*>(se_smem);
    float2* x = bufs + (size_t)warp * 2 * BL;
    float2* y = x + BL;
    const int t = t0 + warp;
    const float* src = a.in + (size_t)row * a.in_stride;
    // the FFT ping-pongs once per pass: the split writes to the buffer the transform did NOT end in (same for every warp)
    const int spec_off = BZ ? BL : ((gen_fft_passes(M) & 1) ? 0 : BL);
    if (t < a.nframe) {
        const int base = t * a.hop;
        // frames that lie inside the row need no reflection / zero logic per sample
        const bool interior = LMODE != GEN_ENV && base - a.pad >= 0 && base - a.pad + n <= a.in_len;
        if (interior) {
            const float* f = src + (base - a.pad);
            const float2* w2p = reinterpret_cast<const float2*>(a.tb.win);
#pragma unroll 4
            for (int m = lane; m < M; m += 32) {
                const float2 w = __ldg(w2p + m);
                x[m] = make_float2(__ldg(f + 2 * m) * w.x, __ldg(f + 2 * m + 1) * w.y);
            }
        } else {
            for (int m = lane; m < M; m += 32) {
                const float w0 = __ldg(a.tb.win + 2 * m), w1 = __ldg(a.tb.win + 2 * m + 1);
                // zero-window samples are never fetched (front windows: win_len < n)
                const float v0 = w0 != 0.f ? gen_sample<LMODE>(a, src, base + 2 * m) * w0 : 0.f;
                const float v1 = w1 != 0.f ? gen_sample<LMODE>(a, src, base + 2 * m + 1) * w1 : 0.f;
                x[m] = make_float2(v0, v1);
            }
        }
        __syncwarp();
        if (a.parity_len > 0) parity_correct(x, M, lane, a.parity_len, a.inv_even, a.inv_odd);
        const float2* z = gen_cdft<false, BZ>(x, y, a.tb, M, lane);
        float2* o = x + spec_off;
        // real-FFT split: X[k] = 1/2 [(Z_k + conj Z_{M-k}) - i e^{-2 pi i k/n} (Z_k - conj Z_{M-k})], k = 0 .. M
        for (int k = lane; k <= M; k += 32) {
            const float2 zk = z[k == M ? 0 : k], zc = g_conj(z[k == 0 ? 0 : M - k]);
            const float2 s = g_add(zk, zc), d = g_sub(zk, zc);
            const float2 wd = g_mul(d, __ldg(a.tb.twn + k));
            const float wt = 0.5f * ((k == 0 || k == M) ? a.edge_w : a.mid_w);
            o[k] = make_float2(wt * (s.x + wd.y), wt * (s.y - wd.x));
        }
    }
    __syncthreads();
    const int F = M + 1, T = a.nframe;
    const int lw = (W > 4) ? 3 : (W > 2 ? 2 : W - 1);             // log2 W: W is 1, 2, 4 or 8 (gen_warps)
    for (int e = threadIdx.x; e < F * W; e += blockDim.x) {
        const int w = e & (W - 1), k = e >> lw, tt = t0 + w;
        if (tt >= T) continue;
        const float2 v = bufs[(size_t)w * 2 * BL + spec_off + k];
        if (a.planar) {
            float* o = a.out + (size_t)row * 2 * F * T;
            o[(size_t)k * T + tt] = v.x;
            o[(size_t)(F + k) * T + tt] = v.y;
        } else {
            reinterpret_cast<float2*>(a.out)[((size_t)row * F + k) * T + tt] = v;
        }
    }
}

// ------------------------------------------------------------------ synthesis 1: spectrum -> windowed frames
// out = scratch [rows, T, f_len]; grid = rows x ceil(T / W), block = 32 W
template <bool BZ>
__global__ void k_gen_frames(GenArgs a) {
    SE_SMEM_DECL;
    pdl_wait();
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.n, M = n >> 1, BL = gen_buf_len(n);
    const int cpr = (a.nframe + W - 1) / W;
    const int row = blockIdx.x / cpr, t0 = (blockIdx.x - row * cpr) * W;
    float2* bufs = reinterpret_cast<float2*>(se_smem);
    const int F = M + 1, T = a.nframe;
    // bins of W frames, frame index fastest in memory; lands in each warp's second buffer.  Four loads in flight per
    // thread: the phase is latency-bound otherwise (one dependent global load per iteration)
    const int lw = (W > 4) ? 3 : (W > 2 ? 2 : W - 1);             // log2 W: W is 1, 2, 4 or 8 (gen_warps)
    const int FW = F * W;
    for (int e0 = threadIdx.x; e0 < FW; e0 += 4 * blockDim.x) {
        float2 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = e0 + j * blockDim.x;
            const int w = e & (W - 1), k = e >> lw, tt = t0 + w;
            v[j] = make_float2(0.f, 0.f);
            if (e < FW && tt < T) {
                if (a.planar) {
                    const float* s = a.in + (size_t)row * 2 * F * T;
                    v[j] = make_float2(__ldg(s + (size_t)k * T + tt), __ldg(s + (size_t)(F + k) * T + tt));
                } else {
                    v[j] = __ldg(reinterpret_cast<const float2*>(a.in) + ((size_t)row * F + k) * T + tt);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = e0 + j * blockDim.x;
            if (e >= FW) continue;
            const int w = e & (W - 1), k = e >> lw;
            float2 u = v[j];
            if (k == 0 || k == M) u = make_float2(u.x * a.edge_w, 0.f);        // imaginary parts of DC / Nyquist are ignored
            else u = make_float2(u.x * a.mid_w, u.y * a.mid_w);
            bufs[(size_t)w * 2 * BL + BL + k] = u;
        }
    }
    __syncthreads();
    const int t = t0 + warp;
    if (t >= T) return;
    float2* x = bufs + (size_t)warp * 2 * BL;
    float2* y = x + BL;
    float* dst = a.out + ((size_t)row * T + t) * a.f_len;
    // Z_k = (X_k + conj X_{M-k}) + i e^{+2 pi i k/n} (X_k - conj X_{M-k}); inverse FFT gives z[m] = v[2m] + i v[2m+1],
    // v[j] = X_0 + (-1)^j X_M + 2 Re sum_{0<k<M} X_k e^{+2 pi i jk/n}
    for (int k = lane; k < M; k += 32) {
        const float2 xk = y[k], xc = g_conj(y[M - k]);
        const float2 s = g_add(xk, xc), d = g_sub(xk, xc);
        const float2 wd = g_mul(d, g_conj(__ldg(a.tb.twn + k)));
        x[k] = make_float2(s.x - wd.y, s.y + wd.x);
    }
    __syncwarp();
    float2* z = gen_cdft<true, BZ>(x, y, a.tb, M, lane);
    if (a.parity_len > 0) parity_correct(z, M, lane, a.parity_len, a.inv_even, a.inv_odd);
    float2* dst2 = reinterpret_cast<float2*>(dst);
    const int m_lo = a.f_lo >> 1, m_cnt = a.f_len >> 1;
    for (int m = lane; m < m_cnt; m += 32) {
        const float2 v = z[m_lo + m];
        const int j = 2 * (m_lo + m);
        dst2[m] = make_float2(v.x * __ldg(a.tb.win + j), v.y * __ldg(a.tb.win + j + 1));
    }
}

// Overlap-add and envelope at natural position i, walking the covering frames once for both.
// The envelope sums w2 over EVERY frame covering i; the signal only over the stored support [f_lo, f_lo + f_len) --
// outside it the window (and so the frame) is zero, so one loop over the support frames serves both.
__device__ __forceinline__ void gen_ola_env_at(const GenArgs& a, const float* __restrict__ frames, int i, bool want_env,
                                               float& acc, float& env) {
    const int T = a.nframe, hop = a.hop;
    acc = 0.f;
    env = 0.f;
    if (i < a.f_lo) return;
    int t_hi = (i - a.f_lo) / hop;
    t_hi = t_hi < T - 1 ? t_hi : T - 1;
    int t_lo = i - a.f_lo - a.f_len + 1;
    t_lo = t_lo <= 0 ? 0 : (t_lo + hop - 1) / hop;
#pragma unroll 4
    for (int t = t_hi; t >= t_lo; --t) {
        const int j = i - t * hop;
        acc += __ldg(frames + (size_t)t * a.f_len + (j - a.f_lo));
        if (want_env) env += __ldg(a.tb.w2 + j);
    }
}
__device__ __forceinline__ float gen_ola_at(const GenArgs& a, const float* __restrict__ frames, int i) {
    float acc, env;
    gen_ola_env_at(a, frames, i, false, acc, env);
    return acc;
}

// ------------------------------------------------------------------ synthesis 2: overlap-add by output sample
// in = scratch [rows, T, f_len]; grid.x covers out_len in blocks of blockDim.x, grid.y = rows
template <int EMODE>
__global__ void k_gen_ola(GenArgs a) {
    pdl_wait();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = blockIdx.y;
    if (s >= a.out_len) return;
    const float* frames = a.in + (size_t)row * a.nframe * a.f_len;
    float* dst = a.out + (size_t)row * a.out_len + s;
    float v;
    if (EMODE == GEN_OLA_ISTFT) {
        const int i = s + a.pad;
        const int natural = a.n + a.hop * (a.nframe - 1);
        if (i >= natural) {
            v = 0.f;                                                // `length` beyond the signal: zeros (torch.istft pads)
        } else {
            float acc, env;
            gen_ola_env_at(a, frames, i, true, acc, env);           // w2 is zero outside the support: same sum as gen_env_at
            v = acc / (env + a.env_eps);
        }
    } else {
        // adjoint of the reflect padding: sample s receives its own position and its mirror images
        const int N = a.nsample, pad = a.pad;
        v = gen_ola_at(a, frames, s + pad);
        if (s >= 1 && s <= pad) v += gen_ola_at(a, frames, pad - s);
        if (s <= N - 2 && s >= N - 1 - pad) v += gen_ola_at(a, frames, pad + 2 * (N - 1) - s);
        if (a.accumulate) v += *dst;
    }
    *dst = v;
}

}  // namespace se
