// se_kernels.cuh -- the spectral kernels built on the frame-interleaved FFT engine (se_fft.cuh).
//
// Coordinates: every transform works in PADDED coordinates i (reflect / zero padded signal, i = 0
// is the first sample of frame 0); frame t covers i in [t*HOP, t*HOP + N); "block" b is the
// hop-sized span [b*HOP, (b+1)*HOP).  A CTA owns one row and a contiguous chunk of it and walks
// the chunk in groups of 16 frames.
#pragma once
#include "se_fft.cuh"

// pass-C task loop: tasks share no state, so rolling it halves code size and register pressure
#ifndef SE_TC_PRAGMA
#define SE_TC_PRAGMA _Pragma("unroll 1")
#endif

namespace se {

enum LoadMode { LOAD_REFLECT = 0, LOAD_ZEROPAD = 1, LOAD_ENV = 2 };
enum EmitMode { EMIT_ISTFT = 0, EMIT_ADJ = 1 };

struct Tables {
    const float* win;       // window (placed in n_fft, scaled per op kind), N floats
    const float2* tw;       // exp(-2 pi i k / M), k < M
    const float2* twn;      // exp(-2 pi i k / N), k < M
    const float* w2;        // unscaled window^2, N floats
    const float* inv_env;   // 1 / sum_q w2[o + q*HOP], HOP floats
};

struct AnaArgs {            // analysis: waveform-like -> spectrum
    Tables tb;
    const float* in;        // [rows, in_len]
    float* out;             // spectrum
    int64_t in_stride;      // floats between rows of `in` (segments: between consecutive segments of a clip)
    int64_t clip_stride;    // segments only: floats between clips; row = seg * seg_rows + clip
    int seg_rows;           // clips per segment index (1 = plain rows)
    int clip_len;           // segments only: valid samples per clip (the rest reads as zero); 0 = no limit
    int nsample;            // N (REFLECT / ZEROPAD: valid input samples; ENV: natural padded length)
    int in_len;             // ENV: `length` of gy rows
    int nframe;             // T
    int pad;                // ZEROPAD: zeros in front
    int gpc, nchunks;       // groups per chunk, chunks per row
    float edge_scale;       // multiplies DC and Nyquist outputs
    float* feat;            // optional second output [rows, F, T]: magnitude feature of the spectrum (a6)
    int feat_kind;
    // evaluate()'s z-score folded into the fill (src/evaluate.py:18-21): per clip (mean, 1/(std+1e-9), std+1e-9, 0);
    // clip -> statistics row = (clip / norm_div) * norm_c + clip % norm_c.  nullptr: no normalisation.
    const float4* norm;
    int norm_div, norm_c;
    // evaluate()'s shared-frame path: only the groups that contain reflect-boundary frames are transformed per segment
    // (chunk c < edge_lead -> group c, else group edge_trail0 + c - edge_lead; one group per chunk); 0 = all groups
    int edge_lead, edge_trail0;
};

struct SynArgs {            // synthesis: spectrum -> waveform-like
    Tables tb;
    const float* in;        // spectrum [rows, F, T, 2]
    float* out;             // [rows, out_len]
    int nsample;            // ADJ: N ; ISTFT: natural padded length n + hop (T-1)
    int out_len;            // ISTFT: length ; ADJ: N
    int nframe;             // T
    int b_lo, b_hi;         // emitted block range per row
    int nchunks;
    int accumulate;
    float edge_scale;       // multiplies DC / Nyquist inputs
};

// ------------------------------------------------------------------ padded-signal samplers
__device__ __forceinline__ float sample_reflect(const float* __restrict__ x, int n_half, int N, int n_fft, int i, int nvalid,
                                                float mean = 0.f, float inv = 1.f) {
    if (i < 0 || i >= N + n_fft) return 0.f;
    int j = i - n_half;
    j = j < 0 ? -j : j;
    j = j >= N ? 2 * (N - 1) - j : j;
    return j < nvalid ? (__ldg(x + j) - mean) * inv : 0.f;      // nvalid < N: zero-filled tail of the last segments
}

template <class G>
__device__ __forceinline__ float inv_env_at(const Tables& tb, int T, int i) {
    const int b = i / G::HOP, o = i - b * G::HOP;
    if (b >= G::OLA - 1 && b <= T - 1) return __ldg(tb.inv_env + o);
    float e = 0.f;
#pragma unroll
    for (int q = 0; q < G::OLA; ++q) {
        const int t = b - q;
        if (t >= 0 && t < T) e += __ldg(tb.w2 + o + q * G::HOP);
    }
    return e > 0.f ? 1.0f / e : 0.f;
}

// Fill the hop-row-padded stage with padded coordinates [p0, p0 + SROWS*HOP).
// All global loads of a thread are issued back to back into registers before the first store, so a
// fill costs ONE memory round trip instead of one per loop iteration (ncu: the store that waited on
// the load held 90 % of the long-scoreboard samples before this change).
// NORM (evaluate()'s z-score) is a template switch: with the affine map compiled into every instance the plain STFT
// lost 15 % (28.7 -> 33.1 us at cfg2: different schedule, the loads no longer batch as well)
template <class G, int LMODE, bool NORM = false>
__device__ __forceinline__ void fill_stage(float* __restrict__ stage, const float* __restrict__ src,
                                           int p0, const AnaArgs& a, int tid, int nvalid = 0x7fffffff,
                                           float nm_mean = 0.f, float nm_inv = 1.f) {
    constexpr int SLOTS = G::SROWS * G::HOP / 2;                 // float2 slots
    constexpr int K = (SLOTS + G::NT - 1) / G::NT;
    float2 v[K];
    const int base = (LMODE == LOAD_ZEROPAD) ? p0 - a.pad : p0 - G::N / 2;   // source index of slot 0
    const int limit = (LMODE == LOAD_ENV) ? a.in_len : (a.nsample < nvalid ? a.nsample : nvalid);
    const bool interior = base >= 0 && base + 2 * SLOTS <= limit &&
                          (LMODE != LOAD_ENV || p0 + 2 * SLOTS <= a.nsample);
    if (interior) {                                              // uniform per CTA: plain streaming loads
        if ((reinterpret_cast<uintptr_t>(src + base) & 7) == 0) {
            const float2* s2 = reinterpret_cast<const float2*>(src + base);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int slot = tid + k * G::NT;
                if (slot < SLOTS) v[k] = __ldg(s2 + slot);
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int slot = tid + k * G::NT;
                if (slot < SLOTS) v[k] = make_float2(__ldg(src + base + 2 * slot), __ldg(src + base + 2 * slot + 1));
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int slot = tid + k * G::NT;
            if (slot >= SLOTS) continue;
            const int i = p0 + 2 * slot;
            if (LMODE == LOAD_REFLECT) {
                if (NORM) v[k] = make_float2(sample_reflect(src, G::N / 2, a.nsample, G::N, i, nvalid, nm_mean, nm_inv),
                                             sample_reflect(src, G::N / 2, a.nsample, G::N, i + 1, nvalid, nm_mean, nm_inv));
                else v[k] = make_float2(sample_reflect(src, G::N / 2, a.nsample, G::N, i, nvalid),
                                        sample_reflect(src, G::N / 2, a.nsample, G::N, i + 1, nvalid));
            } else if (LMODE == LOAD_ZEROPAD) {
                const int j = i - a.pad;
                v[k] = make_float2((j >= 0 && j < a.nsample) ? __ldg(src + j) : 0.f,
                                   (j + 1 >= 0 && j + 1 < a.nsample) ? __ldg(src + j + 1) : 0.f);
                if (NORM) {                                  // valid samples are normalised, the zero padding stays zero
                    if (j >= 0 && j < a.nsample) v[k].x = (v[k].x - nm_mean) * nm_inv;
                    if (j + 1 >= 0 && j + 1 < a.nsample) v[k].y = (v[k].y - nm_mean) * nm_inv;
                }
            } else {
                const int q = i - G::N / 2;
                v[k] = make_float2((q >= 0 && q < a.in_len && i < a.nsample) ? __ldg(src + q) : 0.f,
                                   (q + 1 >= 0 && q + 1 < a.in_len && i + 1 < a.nsample) ? __ldg(src + q + 1) : 0.f);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int slot = tid + k * G::NT;
        if (slot >= SLOTS) continue;
        const int rel = 2 * slot;
        float2 w = v[k];
        if (NORM && LMODE != LOAD_ENV && interior) w = make_float2((w.x - nm_mean) * nm_inv, (w.y - nm_mean) * nm_inv);
        if (LMODE == LOAD_ENV) {      // gy / envelope (zero where the envelope is empty)
            const int i = p0 + rel;
            w.x *= inv_env_at<G>(a.tb, a.nframe, i);
            w.y *= inv_env_at<G>(a.tb, a.nframe, i + 1);
        }
        *reinterpret_cast<float2*>(stage + (rel / G::HOP) * G::SROW + rel % G::HOP) = w;
    }
}

// Asynchronous variant for interior spans of the reflect-padded signal: issues the whole fill as cp.async copies and
// returns true; the caller waits (se_cp_async_wait_all + barrier) right before pass A reads the stage.  Returns false
// -- nothing issued -- when the span touches a signal edge or the source is not 8-byte aligned (the synchronous
// fill_stage handles those).  The stage is only read in pass A, so the fill of the NEXT transform can be issued as soon
// as pass A's barrier has passed and lands while passes B and C run (round 1 measured the exposed fill latency at
// 12 % of the loss forward: profiles/r01_notes.md (e)).
template <class G>
__device__ __forceinline__ bool fill_stage_async(float* __restrict__ stage, const float* __restrict__ src, int p0, int nsample,
                                                 int tid) {
    constexpr int SLOTS = G::SROWS * G::HOP / 2;
    constexpr int K = (SLOTS + G::NT - 1) / G::NT;
    const int base = p0 - G::N / 2;
    if (base < 0 || base + 2 * SLOTS > nsample || (reinterpret_cast<uintptr_t>(src + base) & 7) != 0) return false;
    const float2* s2 = reinterpret_cast<const float2*>(src + base);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int slot = tid + k * G::NT;
        if (slot < SLOTS) {
            const int rel = 2 * slot;
            se_cp_async8(stage + (rel / G::HOP) * G::SROW + rel % G::HOP, s2 + slot);
        }
    }
    se_cp_async_commit();
    return true;
}

// ------------------------------------------------------------------ magnitude features
// NN input features computed from the spectrum, quirks of the reference kept:
//   0 power     |re^2 + im^2|      src/model/unet.py:40
//   1 magnitude sqrt(re^2 + im^2)  src/model/dnn.py:98
//   2 amplitude |re^2 - im^2|      src/model/dcunet.py:379, stft_rnn.py:119, mel_rnn.py:123   (sic)
//   3 crn       sqrt(re^2 - im^2)  src/model/crn.py:101   (NaN where |im| > |re|, like the reference)
__device__ __forceinline__ float feature_of(float2 x, int kind) {
    const float a = x.x * x.x, b = x.y * x.y;
    if (kind == 0) return fabsf(a + b);
    if (kind == 1) return sqrtf(a + b);
    if (kind == 2) return fabsf(a - b);
    return sqrtf(a - b);
}

static __global__ void __launch_bounds__(256) k_feature(const float2* __restrict__ spec, float* __restrict__ feat,
                                                 int64_t count, int kind) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(spec) | reinterpret_cast<uintptr_t>(feat)) & 15) == 0;
    const int64_t pairs = vec ? count / 2 : 0;
    for (int64_t i = tid0; i < pairs; i += stride) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(spec) + i);
        reinterpret_cast<float2*>(feat)[i] = make_float2(feature_of(make_float2(x.x, x.y), kind),
                                                         feature_of(make_float2(x.z, x.w), kind));
    }
    for (int64_t i = 2 * pairs + tid0; i < count; i += stride) feat[i] = feature_of(__ldg(spec + i), kind);
}

// ------------------------------------------------------------------ spectrum row I/O
// interleaved [F][T] float2
// Addressing: the 8 bins of a unit are S*T float2 apart, so each unit walks ONE 64-bit pointer by a constant
// stride (two integer instructions per access).  Indexing every access as row[(q + S k) T + t] made the compiler
// rebuild the full 64-bit address each time -- 17 integer instructions per pair of stores in the ncu source view.
template <class G>
__device__ __forceinline__ void store_task_ft2(float2* __restrict__ row, int T, int t, int p,
                                               const float2* xa, const float2* xb, float2 nyq, float edge,
                                               float* __restrict__ frow = nullptr, int kind = 0) {
    if (t < 0 || t >= T) return;
    const size_t step = (size_t)G::S * T;
    const size_t ia = (size_t)task_qa<G>(p) * T + t, ib = (size_t)task_qb<G>(p) * T + t;
    float2* pa = row + ia;
    float2* pb = row + ib;
    const float2 dc = make_float2(xa[0].x * edge, 0.f);
    if (frow == nullptr) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            *pa = (p == 0 && k4 == 0) ? dc : xa[k4];
            *pb = xb[k4];
            pa += step;
            pb += step;
        }
    } else {                                          // the NN's input feature, written while the bin is in registers
        float* fa = frow + ia;
        float* fb = frow + ib;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            const float2 v = (p == 0 && k4 == 0) ? dc : xa[k4];
            *pa = v;
            *pb = xb[k4];
            *fa = feature_of(v, kind);
            *fb = feature_of(xb[k4], kind);
            pa += step; pb += step; fa += step; fb += step;
        }
    }
    if (p == 0) {
        const float2 v = make_float2(nyq.x * edge, 0.f);
        row[(size_t)G::M * T + t] = v;
        if (frow) frow[(size_t)G::M * T + t] = feature_of(v, kind);
    }
}
// planar [2F][T] floats (DCCRN)
template <class G>
__device__ __forceinline__ void store_task_planar(float* __restrict__ row, int T, int t, int p,
                                                  const float2* xa, const float2* xb, float2 nyq) {
    if (t < 0 || t >= T) return;
    const int qa = task_qa<G>(p), qb = task_qb<G>(p);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        const int ka = qa + G::S * k4, kb = qb + G::S * k4;
        row[(size_t)ka * T + t] = xa[k4].x;
        row[(size_t)(G::F + ka) * T + t] = xa[k4].y;
        row[(size_t)kb * T + t] = xb[k4].x;
        row[(size_t)(G::F + kb) * T + t] = xb[k4].y;
    }
    if (p == 0) {
        row[(size_t)G::M * T + t] = nyq.x;
        row[(size_t)(G::F + G::M) * T + t] = 0.f;
    }
}
template <class G>
__device__ __forceinline__ void load_task_ft2(const float2* __restrict__ row, int T, int t, int p,
                                              float2* ya, float2* yb, float2& nyq, float edge) {
    const bool ok = (t >= 0 && t < T);
    const int tc = ok ? t : 0;                                   // clamped: loads stay in bounds, result zeroed
    const size_t step = (size_t)G::S * T;
    const float2* pa = row + (size_t)task_qa<G>(p) * T + tc;
    const float2* pb = row + (size_t)task_qb<G>(p) * T + tc;
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        ya[k4] = __ldg(pa);
        yb[k4] = __ldg(pb);
        pa += step;
        pb += step;
    }
    nyq = make_float2(0.f, 0.f);
    if (p == 0) nyq = __ldg(row + (size_t)G::M * T + tc);
    if (!ok) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) ya[k4] = yb[k4] = make_float2(0.f, 0.f);
        nyq = make_float2(0.f, 0.f);
    }
    if (p == 0) {
        ya[0].x *= edge;
        nyq.x *= edge;
    }
}

// ------------------------------------------------------------------ analysis group (stage -> registers)
// After this call zb holds pass-B output; the caller runs pass C per paired task.
template <class G>
__device__ __forceinline__ void analysis_passes(const float* stage, const Tables& tb, float2* zb, int unit, int fr) {
    passA_fwd<G>(stage, tb.win, tb.tw, zb, unit, fr);
    __syncthreads();
    passB_fwd<G>(tb.tw, zb, unit, fr);
    __syncthreads();
}
template <class G>
__device__ __forceinline__ void analysis_task(const float2* zb, const Tables& tb, int p, int fr,
                                              float2* xa, float2* xb, float2& nyq) {
    passC_fwd_unit<G>(zb, task_qa<G>(p), fr, xa);
    passC_fwd_unit<G>(zb, task_qb<G>(p), fr, xb);
    split_task<G>(p, tb.twn, xa, xb, nyq);
}
// synthesis: registers -> zb (pass C'), then B', then per-task A' + OLA into ostage
template <class G>
__device__ __forceinline__ void synthesis_task(float2* zb, const Tables& tb, int p, int fr,
                                               float2* ya, float2* yb, float2 nyq) {
    merge_task<G>(p, tb.twn, ya, yb, nyq);
    passC_inv_unit<G>(zb, task_qa<G>(p), fr, ya);
    passC_inv_unit<G>(zb, task_qb<G>(p), fr, yb);
}
// CARRY=false: single-group chunks -- the wrapped contributions belong to blocks nobody emits
template <class G, bool CARRY = true>
__device__ __forceinline__ void synthesis_tail(float2* zb, const Tables& tb, float* ostage, int unit, int fr,
                                               float2 (*carry)[G::SEG]) {
    __syncthreads();
    passB_inv<G>(tb.tw, zb, unit, fr);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < G::TA; ++i) {
        const int u = unit + i * G::NU;
        float2 v[G::R1], acc[G::SEG];
        passA_inv_task<G>(zb, tb.win, tb.tw, u, fr, v);
        ola_rotate<G, CARRY>(v, fr, CARRY ? carry[i] : nullptr, acc);
#pragma unroll
        for (int s = 0; s < G::SEG; ++s)
            *reinterpret_cast<float2*>(ostage + fr * G::SROW + 2 * (u + 64 * s)) = acc[s];
    }
    __syncthreads();
}

// chunk geometry shared by all synthesis-type kernels
struct Chunk { int b0, b1, f0, ngroups; bool first, last; };
template <class G>
__device__ __forceinline__ Chunk make_chunk(int chunk, int nchunks, int b_lo, int b_hi) {
    Chunk c;
    const int nb = b_hi - b_lo;
    c.b0 = b_lo + (int)(((int64_t)chunk * nb) / nchunks);
    c.b1 = b_lo + (int)(((int64_t)(chunk + 1) * nb) / nchunks);
    c.f0 = c.b0 - (G::OLA - 1);
    c.ngroups = (c.b1 - c.f0 + G::FR - 1) / G::FR;
    c.first = chunk == 0;
    c.last = chunk == nchunks - 1;
    return c;
}

// ------------------------------------------------------------------ emitters
template <class G>
__device__ __forceinline__ void emit_istft_one(const float* __restrict__ ostage, float* __restrict__ y_row,
                                               int blk, int o, int b, const SynArgs& a) {
    const int i = b * G::HOP + o;
    const int s = i - G::N / 2;
    if (s < 0 || s >= a.out_len) return;
    float r = 0.f;
    if (i < a.nsample) r = ostage[blk * G::SROW + o] * inv_env_at<G>(a.tb, a.nframe, i);
    y_row[s] = r;
}

template <class G>
__device__ __forceinline__ void emit_istft(const float* __restrict__ ostage, float* __restrict__ y_row,
                                           int f_base, const Chunk& c, const SynArgs& a, int tid) {
    if ((reinterpret_cast<uintptr_t>(y_row) & 15) == 0 && (G::HOP & 3) == 0) {
        // 128-bit stores: 4 consecutive samples never straddle a hop block, and s = i - n/2 keeps 4-alignment
        for (int idx = tid; idx < G::FR * G::HOP / 4; idx += G::NT) {
            const int e = 4 * idx;
            const int blk = e / G::HOP, o = e - blk * G::HOP;
            const int b = f_base + blk;
            if (b < c.b0 || b >= c.b1) continue;
            const int i = b * G::HOP + o, s = i - G::N / 2;
            if (s >= 0 && s + 3 < a.out_len && i + 3 < a.nsample && b >= G::OLA - 1 && b <= a.nframe - 1) {
                const float2 v0 = *reinterpret_cast<const float2*>(ostage + blk * G::SROW + o);
                const float2 v1 = *reinterpret_cast<const float2*>(ostage + blk * G::SROW + o + 2);
                const float4 w = __ldg(reinterpret_cast<const float4*>(a.tb.inv_env + o));
                *reinterpret_cast<float4*>(y_row + s) = make_float4(v0.x * w.x, v0.y * w.y, v1.x * w.z, v1.y * w.w);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) emit_istft_one<G>(ostage, y_row, blk, o + k, b, a);
            }
        }
    } else {
        for (int idx = tid; idx < G::FR * G::HOP; idx += G::NT) {
            const int blk = idx / G::HOP, o = idx - blk * G::HOP;
            const int b = f_base + blk;
            if (b < c.b0 || b >= c.b1) continue;
            emit_istft_one<G>(ostage, y_row, blk, o, b, a);
        }
    }
}

// adjoint emitter: fold the reflect padding back.  hold[] keeps the right-edge zone until the
// row's last group (its mirror sources arrive later than its destinations).
template <class G>
__device__ __forceinline__ void emit_adj(const float* __restrict__ ostage, float* __restrict__ hold,
                                         float* __restrict__ gx_row, int f_base, const Chunk& c,
                                         int N, int accumulate, float gain, int tid) {
    constexpr int NH = G::N / 2;
    constexpr int K = G::FR * G::HOP / G::NT;                 // samples per thread
    static_assert(G::FR * G::HOP % G::NT == 0, "emit tiling");
    const int zs = ((N - 1) / G::HOP) * G::HOP;
    float v[K], old[K];
    int dst[K];                                               // >= 0: gx index, -1: nothing, <= -2: hold slot -(dst+2)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int idx = tid + k * G::NT;
        const int blk = idx / G::HOP, o = idx - blk * G::HOP;
        const int b = f_base + blk;
        const int i = b * G::HOP + o;
        dst[k] = -1;
        v[k] = 0.f;
        if (b < c.b0 || b >= c.b1 || i >= N + G::N) continue;
        float val = ostage[blk * G::SROW + o];
        if (i > NH && i <= G::N) {                       // left mirror: x[j] also fed p[n/2 - j]
            const int is = G::N - i;                      // source padded coordinate, < n/2
            const int sb = is / G::HOP - f_base;          // always inside this (first) group
            val += ostage[sb * G::SROW + is % G::HOP];
        }
        v[k] = val;
        if (c.last && i >= zs) dst[k] = -2 - (i - zs);
        else if (i >= NH) dst[k] = i - NH;
    }
    // read-modify-write of g_est (resolutions 2 and 3 accumulate): all loads in flight before the first store
    if (accumulate) {
#pragma unroll
        for (int k = 0; k < K; ++k) old[k] = dst[k] >= 0 ? gx_row[dst[k]] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (dst[k] >= 0) gx_row[dst[k]] = accumulate ? old[k] + v[k] * gain : v[k] * gain;
        else if (dst[k] <= -2) hold[-(dst[k] + 2)] = v[k];
    }
}
template <class G>
__device__ __forceinline__ void finish_adj(const float* __restrict__ hold, float* __restrict__ gx_row, int N,
                                           int accumulate, float gain, int tid) {
    constexpr int NH = G::N / 2;
    const int zs = ((N - 1) / G::HOP) * G::HOP;
    for (int i = zs + tid; i < N + NH; i += G::NT) {
        float v = hold[i - zs];
        if (i >= N - 1 && i <= N + NH - 2) v += hold[(2 * N + G::N - 2 - i) - zs];   // right mirror
        const int j = i - NH;
        gx_row[j] = accumulate ? gx_row[j] + v * gain : v * gain;
    }
}

// ------------------------------------------------------------------ tables in shared memory
// window (N floats), exp(-2 pi i k/M) and exp(-2 pi i k/n) (M float2 each) are copied into shared memory
// once per CTA (12 N bytes, 128-bit copies that hit L1 after the first CTA of an SM): every twiddle /
// window read inside the passes is then an LDS broadcast instead of an L1-tagged global load (ncu:
// long-scoreboard on those loads was the largest stall class, L1TEX the busiest unit).
template <class G>
__device__ __forceinline__ Tables stage_tables(const Tables& g, unsigned char* dst, int tid) {
    float4* d = reinterpret_cast<float4*>(dst);
    const float4* w = reinterpret_cast<const float4*>(g.win);
    const float4* t0 = reinterpret_cast<const float4*>(g.tw);
    const float4* t1 = reinterpret_cast<const float4*>(g.twn);
    constexpr int NW = G::N / 4, NTW = G::M / 2;
    for (int i = tid; i < NW; i += G::NT) d[i] = __ldg(w + i);
    for (int i = tid; i < NTW; i += G::NT) { d[NW + i] = __ldg(t0 + i); d[NW + NTW + i] = __ldg(t1 + i); }
    Tables r = g;
    r.win = reinterpret_cast<const float*>(dst);
    r.tw = reinterpret_cast<const float2*>(dst + sizeof(float) * G::N);
    r.twn = reinterpret_cast<const float2*>(dst + sizeof(float) * G::N + sizeof(float2) * G::M);
    return r;
}
// second window only (fused kernels use an analysis and a synthesis window with shared twiddles)
template <class G>
__device__ __forceinline__ Tables stage_window(const Tables& staged, const Tables& g, unsigned char* dst, int tid) {
    float4* d = reinterpret_cast<float4*>(dst);
    const float4* w = reinterpret_cast<const float4*>(g.win);
    for (int i = tid; i < G::N / 4; i += G::NT) d[i] = __ldg(w + i);
    Tables r = g;
    r.win = reinterpret_cast<const float*>(dst);
    r.tw = staged.tw;
    r.twn = staged.twn;
    return r;
}

// ------------------------------------------------------------------ smem carve-up
template <class G> struct Smem {
    static constexpr size_t al16(size_t b) { return (b + 15) / 16 * 16; }      // regions stay 16-byte aligned
    static constexpr size_t ZB = sizeof(float) * G::ZB_FLOATS;
    static constexpr size_t STAGE = al16(sizeof(float) * G::STAGE_FLOATS);
    static constexpr size_t OSTAGE = al16(sizeof(float) * G::OSTAGE_FLOATS);
    static constexpr size_t HOLD = al16(sizeof(float) * (G::N + 2 * G::HOP));
    static constexpr size_t TABLES = sizeof(float) * G::N + 2 * sizeof(float2) * G::M;   // window + 2 twiddle tables
    static constexpr size_t WINDOW = sizeof(float) * G::N;
    static constexpr size_t ANALYSIS = ZB + STAGE + TABLES;
    static constexpr size_t SYNTH_ISTFT = ZB + OSTAGE + TABLES;
    static constexpr size_t SYNTH_ADJ = ZB + OSTAGE + HOLD + TABLES;
    // fused analysis+synthesis: stage and ostage are never live together -> aliased
    static constexpr size_t IOBUF = STAGE > OSTAGE ? STAGE : OSTAGE;
    static constexpr size_t FUSED_ADJ = ZB + IOBUF + HOLD + TABLES;
    static constexpr size_t FUSED_ISTFT = ZB + IOBUF + TABLES + WINDOW;
};

// ================================================================== kernels
// wave-like rows -> spectrum.  LMODE picks the padded-signal definition, PLANAR the layout.
template <class G, int LMODE, bool PLANAR, bool NORM = false>
__global__ void __launch_bounds__(G::NT, G::MINB) k_analysis(const AnaArgs a) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* stage = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.tb, se_smem + Smem<G>::ZB + Smem<G>::STAGE, tid);   // visible after the fill's barrier
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const int seg = row / a.seg_rows, clip = row - seg * a.seg_rows;
    const float* src = a.in + (size_t)seg * a.in_stride + (size_t)clip * a.clip_stride;
    int nvalid = 0x7fffffff;
    if (a.clip_len > 0) {
        const int64_t left = (int64_t)a.clip_len - (int64_t)seg * a.in_stride;
        nvalid = left < 0 ? 0 : (left < a.nsample ? (int)left : a.nsample);
    }
    float nm_mean = 0.f, nm_inv = 1.f;
    if (NORM && a.norm) {
        const float4 st = __ldg(a.norm + (clip / a.norm_div) * a.norm_c + clip % a.norm_c);
        nm_mean = st.x;
        nm_inv = st.y;
    }
    for (int g = 0; g < a.gpc; ++g) {
        const int group = a.edge_lead ? (chunk < a.edge_lead ? chunk : a.edge_trail0 + chunk - a.edge_lead) : chunk * a.gpc + g;
        const int f_base = group * G::FR;
        if (f_base >= a.nframe) break;
        fill_stage<G, LMODE, NORM>(stage, src, f_base * G::HOP, a, tid, nvalid, nm_mean, nm_inv);
        __syncthreads();
        analysis_passes<G>(stage, tb, zb, unit, fr);
        const int t = f_base + fr;
        SE_TC_PRAGMA
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            float2 xa[8], xb[8], nyq;
            analysis_task<G>(zb, tb, p, fr, xa, xb, nyq);
            if (PLANAR) store_task_planar<G>(a.out + (size_t)row * 2 * G::F * a.nframe, a.nframe, t, p, xa, xb, nyq);
            else store_task_ft2<G>(reinterpret_cast<float2*>(a.out) + (size_t)row * G::F * a.nframe, a.nframe, t, p,
                                   xa, xb, nyq, a.edge_scale,
                                   a.feat ? a.feat + (size_t)row * G::F * a.nframe : nullptr, a.feat_kind);
        }
        __syncthreads();
    }
}

// spectrum -> wave-like rows.  EMODE: ISTFT (envelope + trim) or ADJ (reflect fold-back).
template <class G, int EMODE>
__global__ void __launch_bounds__(G::NT, G::MINB) k_synthesis(const SynArgs a) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* ostage = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    float* hold = reinterpret_cast<float*>(se_smem + Smem<G>::ZB + Smem<G>::OSTAGE);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.tb, se_smem + Smem<G>::ZB + Smem<G>::OSTAGE + (EMODE == EMIT_ADJ ? Smem<G>::HOLD : 0), tid);
    __syncthreads();
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const Chunk c = make_chunk<G>(chunk, a.nchunks, a.b_lo, a.b_hi);
    const float2* spec = reinterpret_cast<const float2*>(a.in) + (size_t)row * G::F * a.nframe;
    float* out_row = a.out + (size_t)row * a.out_len;
    float2 carry[G::TA][G::SEG];
#pragma unroll
    for (int i = 0; i < G::TA; ++i)
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) carry[i][s] = make_float2(0.f, 0.f);
    for (int g = 0; g < c.ngroups; ++g) {
        const int f_base = c.f0 + g * G::FR;
        const int t = f_base + fr;
        SE_TC_PRAGMA
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            float2 ya[8], yb[8], nyq;
            load_task_ft2<G>(spec, a.nframe, t, p, ya, yb, nyq, a.edge_scale);
            synthesis_task<G>(zb, tb, p, fr, ya, yb, nyq);
        }
        synthesis_tail<G>(zb, tb, ostage, unit, fr, carry);
        if (EMODE == EMIT_ISTFT) emit_istft<G>(ostage, out_row, f_base, c, a, tid);
        else emit_adj<G>(ostage, hold, out_row, f_base, c, a.nsample, a.accumulate, 1.0f, tid);
    }
    if (EMODE == EMIT_ADJ && c.last) {
        __syncthreads();
        finish_adj<G>(hold, out_row, a.nsample, a.accumulate, 1.0f, tid);
    }
}

// ------------------------------------------------------------------ evaluate(): iSTFT + stitch + de-normalise
// src/evaluate.py:72-96: every segment is inverse-transformed, then only segment 0 (whole) and the LAST `stride` samples
// of each later segment are kept and the z-score is undone.  Here one launch writes the stitched clip directly: CTAs
// of segment 0 run the usual chunked synthesis, one CTA per later segment synthesises only the frames that overlap its
// kept tail (stride / hop blocks + the OLA halo instead of all T frames), and the emitter applies
// y * (std + 1e-9) + mean and the final trim to the clip length.
struct StitchArgs {
    int nclip, nseg;            // rows of the spectrum = nseg * nclip, row = seg * nclip + clip
    int chunks0;                // chunks per segment-0 row
    int stride, num_feature;    // kept tail length, segment length
    int clip_len;               // samples written per clip (the mixture's length)
    int64_t out_stride;         // floats between clips of the output
    int tail_b_lo, tail_b_hi;   // block range of a later segment's kept tail
    const float4* norm;         // per-clip (mean, 1/scale, scale, 0) or nullptr
    int norm_div, norm_c;
};

template <class G>
__device__ __forceinline__ void emit_istft_stitch(const float* __restrict__ ostage, float* __restrict__ clip_out, int f_base,
                                                  const Chunk& c, const SynArgs& a, const StitchArgs& s, int seg,
                                                  float scale, float mean, int tid) {
    const int keep_lo = seg == 0 ? 0 : s.num_feature - s.stride;
    const int shift = seg == 0 ? 0 : s.num_feature + s.stride * (seg - 1) - keep_lo;     // segment sample -> clip position
    for (int idx = tid; idx < G::FR * G::HOP; idx += G::NT) {
        const int blk = idx / G::HOP, o = idx - blk * G::HOP;
        const int b = f_base + blk;
        if (b < c.b0 || b >= c.b1) continue;
        const int i = b * G::HOP + o, sl = i - G::N / 2;                                // segment-local sample
        if (sl < keep_lo || sl >= s.num_feature) continue;
        const int pos = sl + shift;
        if (pos >= s.clip_len) continue;
        float r = 0.f;
        if (i < a.nsample) r = ostage[blk * G::SROW + o] * inv_env_at<G>(a.tb, a.nframe, i);
        clip_out[pos] = r * scale + mean;
    }
}

template <class G>
__global__ void __launch_bounds__(G::NT, G::MINB) k_synthesis_stitch(const SynArgs a, const StitchArgs s) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* ostage = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.tb, se_smem + Smem<G>::ZB + Smem<G>::OSTAGE, tid);
    __syncthreads();
    pdl_wait();
    int seg, clip;
    Chunk c;
    const int head = s.nclip * s.chunks0;
    if ((int)blockIdx.x < head) {
        seg = 0;
        clip = blockIdx.x / s.chunks0;
        c = make_chunk<G>(blockIdx.x - clip * s.chunks0, s.chunks0, a.b_lo, a.b_hi);
    } else {
        const int idx = blockIdx.x - head;
        seg = 1 + idx / s.nclip;
        clip = idx - (seg - 1) * s.nclip;
        c = make_chunk<G>(0, 1, s.tail_b_lo, s.tail_b_hi);
    }
    float scale = 1.f, mean = 0.f;
    if (s.norm) {
        const float4 st = __ldg(s.norm + (clip / s.norm_div) * s.norm_c + clip % s.norm_c);
        mean = st.x;
        scale = st.z;
    }
    const float2* spec = reinterpret_cast<const float2*>(a.in) + ((size_t)seg * s.nclip + clip) * G::F * a.nframe;
    float* clip_out = a.out + (size_t)clip * s.out_stride;
    float2 carry[G::TA][G::SEG];
#pragma unroll
    for (int i = 0; i < G::TA; ++i)
#pragma unroll
        for (int q = 0; q < G::SEG; ++q) carry[i][q] = make_float2(0.f, 0.f);
    for (int g = 0; g < c.ngroups; ++g) {
        const int f_base = c.f0 + g * G::FR;
        const int t = f_base + fr;
        SE_TC_PRAGMA
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            float2 ya[8], yb[8], nyq;
            load_task_ft2<G>(spec, a.nframe, t, p, ya, yb, nyq, a.edge_scale);
            synthesis_task<G>(zb, tb, p, fr, ya, yb, nyq);
        }
        synthesis_tail<G>(zb, tb, ostage, unit, fr, carry);
        emit_istft_stitch<G>(ostage, clip_out, f_base, c, a, s, seg, scale, mean, tid);
    }
}

// per-row mean and unbiased standard deviation (torch.mean / torch.std, src/evaluate.py:19-20), double accumulators:
// stats[row] = (mean, 1 / (std + 1e-9), std + 1e-9, 0)
// A row (one channel of a clip, ~0.5 M samples) is summed by a thread-block CLUSTER of 8 CTAs: each CTA streams an eighth
// of the row (1024 threads x 8 independent 128-bit loads in flight), the partial sums meet in rank 0's shared memory
// through distributed shared memory and are added in rank order -- deterministic, no global scratch, no atomics.
#ifdef SE_EMULATE
constexpr int kStatsCluster = 1;
#define SE_CLUSTER_DIMS(n)
#else
constexpr int kStatsCluster = 8;
#define SE_CLUSTER_DIMS(n) __cluster_dims__(n, 1, 1)
#endif
static __global__ void SE_CLUSTER_DIMS(kStatsCluster) __launch_bounds__(1024) k_row_stats(const float* __restrict__ x,
                                                                                          float4* __restrict__ stats, int64_t len,
                                                                                          int64_t row_stride) {
    __shared__ double sh[2][32];
    __shared__ double parts[2][kStatsCluster];
    pdl_launch_dependents();
    pdl_wait();
#ifndef SE_EMULATE
    // a CTA may only write a peer's shared memory once that peer has started: arrive now, wait right before the stores
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#endif
    const int rowi = blockIdx.x / kStatsCluster, part = blockIdx.x % kStatsCluster;
    const float* row = x + (size_t)rowi * row_stride;
    double s1 = 0.0, s2 = 0.0;
    const bool vec = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
    const int64_t nvec = vec ? len / 4 : 0;
    const int64_t v0 = nvec * part / kStatsCluster, v1 = nvec * (part + 1) / kStatsCluster;
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (int64_t i = v0 + threadIdx.x; i < v1; i += 8 * 1024) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (i + k * 1024 < v1) ? __ldg(r4 + i + k * 1024) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double a = v[k].x, b = v[k].y, c = v[k].z, d = v[k].w;
            s1 += (a + b) + (c + d);
            s2 += (a * a + b * b) + (c * c + d * d);
        }
    }
    if (part == kStatsCluster - 1) {                          // scalar tail (or the whole row when it is not 16-byte aligned)
        for (int64_t i = 4 * nvec + threadIdx.x; i < len; i += 1024) {
            const double v = (double)__ldg(row + i);
            s1 += v;
            s2 += v * v;
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, m);
        s2 += __shfl_xor_sync(0xffffffffu, s2, m);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s1; sh[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
#ifndef SE_EMULATE
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");          // every CTA of the cluster is running
#endif
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 32; ++w) { a += sh[0][w]; b += sh[1][w]; }
#ifdef SE_EMULATE
        parts[0][part] = a;
        parts[1][part] = b;
#else
        // store this CTA's pair into rank 0's shared memory (DSMEM)
        unsigned local = (unsigned)__cvta_generic_to_shared(&parts[0][0]), remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
        asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote + 8u * part), "d"(a) : "memory");
        asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote + 8u * (kStatsCluster + part)), "d"(b) : "memory");
#endif
    }
#ifndef SE_EMULATE
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
    if (part == 0 && threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int c = 0; c < kStatsCluster; ++c) { a += parts[0][c]; b += parts[1][c]; }      // rank order: deterministic
        const double mean = a / (double)len;
        double var = len > 1 ? (b - (double)len * mean * mean) / (double)(len - 1) : 0.0;
        var = var > 0.0 ? var : 0.0;
        const double scale = sqrt(var) + 1e-9;
        stats[rowi] = make_float4((float)mean, (float)(1.0 / scale), (float)scale, 0.f);
    }
}

// evaluate(): interior frames of overlapping segments are frames of ONE clip-level transform (segment s, frame t =
// clip frame s * stride/hop + t whenever the frame does not touch the segment's reflect padding).  Copies
// the frame range [t0, t1) of every (segment, clip, bin) row out of the clip-level spectrum glob [nclip][F][Tg].
constexpr int kGatherRows = 16;                              // rows per block: 418 k one-row blocks were launch-rate bound (3.2 TB/s)
static __global__ void __launch_bounds__(256) k_segment_gather(const float2* __restrict__ glob, float2* __restrict__ out,
                                                                int nclip, int nbin, int T, int Tg, int frames_per_seg, int t0, int t1,
                                                                int nrows) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 8 warps x 2 rows each; a warp streams one row: 128-bit stores on the (DRAM-bound) output side -- rows of odd T
    // alternate their 16-byte phase, so a row may start with a 64-bit head -- and 64-bit loads from the L2-resident
    // clip spectrum
    for (int r = warp; r < kGatherRows; r += 8) {
        const int row = blockIdx.x * kGatherRows + r;        // (seg * nclip + clip) * nbin + bin
        if (row >= nrows) break;
        const int bin = row % nbin, sc = row / nbin, clip = sc % nclip, seg = sc / nclip;
        const float2* src = glob + ((size_t)clip * nbin + bin) * Tg + (size_t)seg * frames_per_seg;
        float2* dst = out + (size_t)row * T;
        int t = t0;
        if ((reinterpret_cast<uintptr_t>(dst + t) & 15) != 0) {
            if (lane == 0 && t < t1) dst[t] = __ldg(src + t);
            ++t;
        }
        const int npair = (t1 - t) / 2;
        for (int p = lane; p < npair; p += 128) {            // 4 independent pairs per lane in flight
            float2 a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int pp = p + 32 * k;
                if (pp < npair) { a[k] = __ldg(src + t + 2 * pp); b[k] = __ldg(src + t + 2 * pp + 1); }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int pp = p + 32 * k;
                if (pp < npair) *reinterpret_cast<float4*>(dst + t + 2 * pp) = make_float4(a[k].x, a[k].y, b[k].x, b[k].y);
            }
        }
        if (((t1 - t) & 1) && lane == 0 && t1 > t) dst[t1 - 1] = __ldg(src + t1 - 1);
    }
}

// ------------------------------------------------------------------ MR-STFT loss
struct LossArgs {
    Tables tb;
    const float* est;
    const float* ref;
    float* g_est;            // bwd
    double* partials;        // fwd: [grid][3]
    float* refmag;           // fwd writes / bwd reads |B| (clamped) as [rows][F][T] for this resolution
    float2* estspec;         // fwd writes / bwd reads the estimate's spectrum A as [rows][F][T] (nullptr: recompute mode)
    const double* sums;      // bwd: this resolution's 3 sums
    const float* gout;       // bwd: device scalar
    int nsample, nframe;
    int gpc, nchunks;        // fwd chunking (analysis style)
    int b_lo, b_hi;          // bwd chunking (synthesis style)
    int accumulate;
    int chained;             // fwd: 1 = independent follower of the previous loss-fwd kernel (see k_loss_fwd)
    float inv_count;         // 1 / (global_rows * F * T)
    float inv_res;           // 1 / number of resolutions
    const double* rows_dev;  // bwd, optional: global row count on the device (uneven shards); overrides inv_count
    double bins_per_row;     // F * T of this resolution
};
#define SE_MRSTFT_CLAMP 1e-7f

template <class G>
__global__ void __launch_bounds__(G::NT, G::MINB) k_loss_fwd(const LossArgs a) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* stage = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    __shared__ float red[3][32];
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    // The three resolutions are independent of each other (disjoint outputs): they are launched largest first
    // and chained so that the next one fills the SMs the previous one's last wave leaves idle.  The head of
    // the chain releases its follower only AFTER its own dependencies are met, so followers need not wait.
    if (a.chained) pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.tb, se_smem + Smem<G>::ZB + Smem<G>::STAGE, tid);
    if (!a.chained) { pdl_wait(); pdl_launch_dependents(); }
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    AnaArgs la;
    la.tb = a.tb; la.nsample = a.nsample; la.nframe = a.nframe; la.in_len = a.nsample; la.pad = 0;
    float s_d2 = 0.f, s_b2 = 0.f, s_lm = 0.f;
    bool prefetched = false;
    for (int g = 0; g < a.gpc; ++g) {
        const int f_base = (chunk * a.gpc + g) * G::FR;
        if (f_base >= a.nframe) break;
        const int t = f_base + fr;
        float pb[G::TC][17];
#pragma unroll 1
        for (int sig = 0; sig < 2; ++sig) {              // 0: reference magnitudes, 1: estimate + statistics
            if (prefetched) se_cp_async_wait_all();      // this transform's fill was issued during the previous one
            else fill_stage<G, LOAD_REFLECT>(stage, (sig ? a.est : a.ref) + (size_t)row * a.nsample, f_base * G::HOP, la, tid);
            __syncthreads();
            passA_fwd<G>(stage, tb.win, tb.tw, zb, unit, fr);
            __syncthreads();
            // the stage is free from here on: start the NEXT transform's fill (the estimate of this group, or the
            // reference of the next group) so that it lands while passes B and C run
            {
                const int nf_base = sig ? f_base + G::FR : f_base;
                const bool more = sig == 0 || (g + 1 < a.gpc && nf_base < a.nframe);
                prefetched = more && fill_stage_async<G>(stage, (sig ? a.ref : a.est) + (size_t)row * a.nsample, nf_base * G::HOP,
                                                         a.nsample, tid);
            }
            passB_fwd<G>(tb.tw, zb, unit, fr);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < G::TC; ++i) {
                const int p = unit + i * G::NU;
                float2 xa[8], xb[8], nyq;
                analysis_task<G>(zb, tb, p, fr, xa, xb, nyq);
                // |B| rows of this task: one pointer per unit walked by S*T (see store_task_ft2)
                float* mrow = a.refmag + (size_t)row * G::F * a.nframe + t;
                const size_t mstep = (size_t)G::S * a.nframe;
                float* wp = mrow + (size_t)task_qa<G>(p) * a.nframe;
#pragma unroll
                for (int k = 0; k < 17; ++k) {
                    const float2 v = k < 8 ? xa[k] : (k < 16 ? xb[k - 8] : nyq);
                    const float pw = v.x * v.x + v.y * v.y;
                    if (sig == 0) {
                        pb[i][k] = pw;
                        if (k == 8) wp = mrow + (size_t)task_qb<G>(p) * a.nframe;
                        if (k == 16) wp = mrow + (size_t)G::M * a.nframe;
                        if (t < a.nframe && !(k == 16 && p != 0)) {      // keep |B| for the backward pass
                            const float cb = fmaxf(pw, SE_MRSTFT_CLAMP);
                            *wp = cb * se_rsqrt(cb);
                        }
                        wp += mstep;
                    } else if (t < a.nframe && !(k == 16 && p != 0)) {
                        const float ca = fmaxf(pw, SE_MRSTFT_CLAMP);
                        const float cb = fmaxf(pb[i][k], SE_MRSTFT_CLAMP);
                        // rounded products: identical spectra must give d == 0 exactly (loss(x, x) = 0, zero gradient)
                        const float d = se_mul_rn(cb, se_rsqrt(cb)) - se_mul_rn(ca, se_rsqrt(ca));
                        s_d2 += d * d;
                        s_b2 += cb;
                        // |log b - log a| = ln2/2 |log2 cb - log2 ca|  (both clamped: normal range)
                        s_lm += 0.34657359f * fabsf(se_log2(cb) - se_log2(ca));
                    }
                }
                // keep A for the backward pass: it then runs one transform (the adjoint) instead of two
                if (sig == 1 && a.estspec && t < a.nframe) {
                    float2* srow = a.estspec + (size_t)row * G::F * a.nframe + t;
                    float2* sa = srow + (size_t)task_qa<G>(p) * a.nframe;
                    float2* sb = srow + (size_t)task_qb<G>(p) * a.nframe;
                    // evict-first stores: written once, read once much later -- keeps the scratch from displacing
                    // the step's reusable lines in L2 (measured: loss fwd 191 -> 184 us)
#pragma unroll
                    for (int k = 0; k < 8; ++k) { se_store_stream(sa, xa[k]); se_store_stream(sb, xb[k]); sa += mstep; sb += mstep; }
                    if (p == 0) se_store_stream(srow + (size_t)G::M * a.nframe, nyq);
                }
            }
        }
    }
    // block reduction -> one deterministic partial per CTA
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        s_d2 += __shfl_xor_sync(0xffffffffu, s_d2, m);
        s_b2 += __shfl_xor_sync(0xffffffffu, s_b2, m);
        s_lm += __shfl_xor_sync(0xffffffffu, s_lm, m);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = s_d2; red[1][tid >> 5] = s_b2; red[2][tid >> 5] = s_lm; }
    __syncthreads();
    if (tid < 3) {
        double acc = 0.0;
        for (int w = 0; w < G::NT / 32; ++w) acc += (double)red[tid][w];
        a.partials[(size_t)blockIdx.x * 3 + tid] = acc;
    }
}

template <class G>
__global__ void __launch_bounds__(G::NT, G::MINB) k_loss_bwd(const LossArgs a) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* iobuf = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);          // stage, later ostage
    float* hold = reinterpret_cast<float*>(se_smem + Smem<G>::ZB + Smem<G>::IOBUF);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    // Chained like the forward kernels (largest first).  Followers accumulate into g_est, which their
    // predecessor writes: they run their whole first group (analysis + synthesis) before waiting, right
    // in front of the first read-modify-write.
    if (a.chained) pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.tb, se_smem + Smem<G>::ZB + Smem<G>::IOBUF + Smem<G>::HOLD, tid);
    if (!a.chained) { pdl_wait(); pdl_launch_dependents(); }
    bool must_wait = a.chained != 0;
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const Chunk c = make_chunk<G>(chunk, a.nchunks, a.b_lo, a.b_hi);
    AnaArgs la;
    la.tb = a.tb; la.nsample = a.nsample; la.nframe = a.nframe; la.in_len = a.nsample; la.pad = 0;
    // dL/da = gs * [ alpha (a - b) + beta sign(a - b) / a ]
    const double d2 = a.sums[0], b2 = a.sums[1];
    const float gs = __ldg(a.gout) * a.inv_res;
    const float alpha = (d2 > 0.0 && b2 > 0.0) ? gs * (float)(1.0 / (sqrt(d2) * sqrt(b2))) : 0.f;
    const float beta = gs * (a.rows_dev ? (float)(1.0 / (a.rows_dev[0] * a.bins_per_row)) : a.inv_count);
    float* gx_row = a.g_est + (size_t)row * a.nsample;
    // n = 2048 runs single-group chunks (no OLA carry to keep live in its tighter register budget);
    // the smaller sizes carry the OLA tail across 2+ groups and recompute less halo
    constexpr bool CARRY = G::N <= 1024;
    float2 carry[CARRY ? G::TA : 1][G::SEG];
#pragma unroll
    for (int i = 0; i < (CARRY ? G::TA : 1); ++i)
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) carry[i][s] = make_float2(0.f, 0.f);
    for (int g = 0; g < c.ngroups; ++g) {
        const int f_base = c.f0 + g * G::FR;
        const int t = f_base + fr;
        const bool live = (t >= 0 && t < a.nframe);
        const int tc = live ? t : 0;
        const float alpha_l = live ? alpha : 0.f, beta_l = live ? beta : 0.f;
        const float* mrow = a.refmag + (size_t)row * G::F * a.nframe + tc;
        const size_t mstep = (size_t)G::S * a.nframe;
        fill_stage<G, LOAD_REFLECT>(iobuf, a.est + (size_t)row * a.nsample, f_base * G::HOP, la, tid);
        __syncthreads();
        // |B| for the first task is requested before the passes so its (L2 / DRAM) latency hides behind them;
        // later tasks are fetched one task ahead
        float mnext[17];
        {
            const int p = unit;
            const float* pa = mrow + (size_t)task_qa<G>(p) * a.nframe;
            const float* pb = mrow + (size_t)task_qb<G>(p) * a.nframe;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                mnext[k] = __ldg(pa);
                mnext[8 + k] = __ldg(pb);
                pa += mstep;
                pb += mstep;
            }
            mnext[16] = __ldg(mrow + (size_t)G::M * a.nframe);
        }
        analysis_passes<G>(iobuf, tb, zb, unit, fr);
        SE_TC_PRAGMA
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            float mb[17];                                    // |B| saved by the forward pass
#pragma unroll
            for (int k = 0; k < 17; ++k) mb[k] = mnext[k];
            if (i + 1 < G::TC) {
                const int pn = p + G::NU;
                const float* pa = mrow + (size_t)task_qa<G>(pn) * a.nframe;
                const float* pb = mrow + (size_t)task_qb<G>(pn) * a.nframe;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    mnext[k] = __ldg(pa);
                    mnext[8 + k] = __ldg(pb);
                    pa += mstep;
                    pb += mstep;
                }
            }
            float2 xa[8], xb[8], nyq;
            analysis_task<G>(zb, tb, p, fr, xa, xb, nyq);
#pragma unroll
            for (int k = 0; k < 17; ++k) {
                float2 v = k < 8 ? xa[k] : (k < 16 ? xb[k - 8] : nyq);
                const float pa = v.x * v.x + v.y * v.y;
                // branch-free: below the clamp the magnitude is constant (zero gradient), dead frames carry
                // zero weights (alpha_l, beta_l), and only p == 0 owns the Nyquist slot
                const float ia = se_rsqrt(fmaxf(pa, SE_MRSTFT_CLAMP));
                const float ma = pa * ia;
                const float sg = ma > mb[k] ? 1.f : (ma < mb[k] ? -1.f : 0.f);
                float coef = alpha_l * (ma - mb[k]) * ia + beta_l * sg * ia * ia;
                if (pa < SE_MRSTFT_CLAMP || (k == 16 && p != 0)) coef = 0.f;
                // edge bins enter the C2R with weight 2 (H = G / c_k, the 1/2 sits in the window)
                if (p == 0 && (k == 0 || k == 16)) coef *= 2.f;
                v = make_float2(v.x * coef, v.y * coef);
                if (k < 8) xa[k] = v; else if (k < 16) xb[k - 8] = v; else nyq = v;
            }
            synthesis_task<G>(zb, tb, p, fr, xa, xb, nyq);
        }
        synthesis_tail<G, CARRY>(zb, tb, iobuf, unit, fr, carry);
        if (must_wait) { pdl_wait(); must_wait = false; }
        emit_adj<G>(iobuf, hold, gx_row, f_base, c, a.nsample, a.accumulate, 1.0f, tid);
        __syncthreads();
    }
    if (c.last) {
        finish_adj<G>(hold, gx_row, a.nsample, a.accumulate, 1.0f, tid);
    }
}

// Backward from a spectrum saved by the forward pass (opt-in, SE_MRSTFT_SAVE_SPECTRUM=1): load A and |B| per bin, form
// G = coef * A in registers and run the STFT adjoint (synthesis + reflect fold).  One transform per resolution
// instead of two, at the price of 8 more bytes per bin of scratch written and read back through DRAM (300 MB per
// 64 x 4 s step, more than L2 holds).  Measured on B200 it is a wash (step 567-575 us against 571-575 us for the
// recompute kernel above: the forward pays 26-33 us for the stores, the backward gains 14-20 us), so recompute,
// with a third of the workspace, stays the default; profiles/r01_notes.md has the table.
template <class G>
__global__ void __launch_bounds__(G::NT, G::MINB) k_loss_bwd_saved(const LossArgs a) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* ostage = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    float* hold = reinterpret_cast<float*>(se_smem + Smem<G>::ZB + Smem<G>::OSTAGE);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    if (a.chained) pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.tb, se_smem + Smem<G>::ZB + Smem<G>::OSTAGE + Smem<G>::HOLD, tid);
    __syncthreads();
    if (!a.chained) { pdl_wait(); pdl_launch_dependents(); }
    bool must_wait = a.chained != 0;
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const Chunk c = make_chunk<G>(chunk, a.nchunks, a.b_lo, a.b_hi);
    const double d2 = a.sums[0], b2 = a.sums[1];
    const float gs = __ldg(a.gout) * a.inv_res;
    const float alpha = (d2 > 0.0 && b2 > 0.0) ? gs * (float)(1.0 / (sqrt(d2) * sqrt(b2))) : 0.f;
    const float beta = gs * (a.rows_dev ? (float)(1.0 / (a.rows_dev[0] * a.bins_per_row)) : a.inv_count);
    float* gx_row = a.g_est + (size_t)row * a.nsample;
    const size_t rbase = (size_t)row * G::F * a.nframe;
    const size_t step = (size_t)G::S * a.nframe;
    float2 carry[G::TA][G::SEG];
#pragma unroll
    for (int i = 0; i < G::TA; ++i)
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) carry[i][s] = make_float2(0.f, 0.f);
    for (int g = 0; g < c.ngroups; ++g) {
        const int f_base = c.f0 + g * G::FR;
        const int t = f_base + fr;
        const bool live = (t >= 0 && t < a.nframe);
        const int tc = live ? t : 0;                       // clamped: loads stay in bounds, weights zeroed
        const float alpha_l = live ? alpha : 0.f, beta_l = live ? beta : 0.f;
        const float2* srow = a.estspec + rbase + tc;
        const float* mrow = a.refmag + rbase + tc;
        SE_TC_PRAGMA
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            float2 ya[8], yb[8], nyq = make_float2(0.f, 0.f);
            float ma[8], mb[8], mn = 1.f;
            {
                const size_t ia = (size_t)task_qa<G>(p) * a.nframe, ib = (size_t)task_qb<G>(p) * a.nframe;
                const float2* sa = srow + ia;
                const float2* sb = srow + ib;
                const float* pa = mrow + ia;
                const float* pb = mrow + ib;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    ya[k] = __ldg(sa); yb[k] = __ldg(sb);
                    ma[k] = __ldg(pa); mb[k] = __ldg(pb);
                    sa += step; sb += step; pa += step; pb += step;
                }
                if (p == 0) {
                    nyq = __ldg(srow + (size_t)G::M * a.nframe);
                    mn = __ldg(mrow + (size_t)G::M * a.nframe);
                }
            }
#pragma unroll
            for (int k = 0; k < 17; ++k) {
                float2 v = k < 8 ? ya[k] : (k < 16 ? yb[k - 8] : nyq);
                const float m = k < 8 ? ma[k] : (k < 16 ? mb[k - 8] : mn);
                const float pw = v.x * v.x + v.y * v.y;
                const float ia = se_rsqrt(fmaxf(pw, SE_MRSTFT_CLAMP));
                const float mag = pw * ia;
                const float sg = mag > m ? 1.f : (mag < m ? -1.f : 0.f);
                float coef = alpha_l * (mag - m) * ia + beta_l * sg * ia * ia;
                if (pw < SE_MRSTFT_CLAMP || (k == 16 && p != 0)) coef = 0.f;
                if (p == 0 && (k == 0 || k == 16)) coef *= 2.f;     // edge bins enter the C2R with weight 2
                v = make_float2(v.x * coef, v.y * coef);
                if (k < 8) ya[k] = v; else if (k < 16) yb[k - 8] = v; else nyq = v;
            }
            synthesis_task<G>(zb, tb, p, fr, ya, yb, nyq);
        }
        synthesis_tail<G>(zb, tb, ostage, unit, fr, carry);
        if (must_wait) { pdl_wait(); must_wait = false; }
        emit_adj<G>(ostage, hold, gx_row, f_base, c, a.nsample, a.accumulate, 1.0f, tid);
    }
    if (c.last) {
        __syncthreads();
        finish_adj<G>(hold, gx_row, a.nsample, a.accumulate, 1.0f, tid);
    }
}

// reduce per-CTA partials -> sums[9] in a fixed order (deterministic).  gridDim.x == 3: one CTA per resolution;
// gridDim.x == 1: one CTA walks the three resolutions and, when `loss` is given, also evaluates the loss value
// (c0..c2 = global bins per resolution) -- the single-process step then needs no separate value launch.
static __global__ void k_reduce_partials(const double* __restrict__ partials, int n0, int n1, int n2, double* __restrict__ sums,
                                         float* __restrict__ loss, double c0, double c1, double c2) {
    __shared__ double sh[3][256];
    __shared__ double tot[9];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    const int r_lo = gridDim.x == 1 ? 0 : blockIdx.x, r_hi = gridDim.x == 1 ? 3 : blockIdx.x + 1;
    for (int r = r_lo; r < r_hi; ++r) {
        const int n = r == 0 ? n0 : (r == 1 ? n1 : n2);
        const double* part = partials + (size_t)3 * (r == 0 ? 0 : (r == 1 ? n0 : n0 + n1));
        double acc[3] = {0.0, 0.0, 0.0};
        for (int i = tid; i < n; i += 256)
            for (int j = 0; j < 3; ++j) acc[j] += part[(size_t)i * 3 + j];
        for (int j = 0; j < 3; ++j) sh[j][tid] = acc[j];
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (tid < s)
                for (int j = 0; j < 3; ++j) sh[j][tid] += sh[j][tid + s];
            __syncthreads();
        }
        if (tid < 3) { sums[3 * r + tid] = sh[tid][0]; tot[3 * r + tid] = sh[tid][0]; }
        __syncthreads();
    }
    if (gridDim.x == 1 && loss && tid == 0) {
        const double cnt[3] = {c0, c1, c2};
        double total = 0.0;
        for (int r = 0; r < 3; ++r) total += sqrt(tot[3 * r]) / sqrt(tot[3 * r + 1]) + tot[3 * r + 2] / cnt[r];
        *loss = (float)(total / 3.0);
    }
}

// c0..c2: global bins per resolution; rows_dev != nullptr: they are bins PER ROW and the global row count is read there
static __global__ void k_loss_value(const double* __restrict__ sums, double c0, double c1, double c2, float* __restrict__ loss,
                                    const double* __restrict__ rows_dev) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double rows = rows_dev ? rows_dev[0] : 1.0;
        const double cnt[3] = {c0 * rows, c1 * rows, c2 * rows};
        double total = 0.0;
        for (int r = 0; r < 3; ++r) {
            const double d2 = sums[3 * r], b2 = sums[3 * r + 1], lm = sums[3 * r + 2];
            total += sqrt(d2) / sqrt(b2) + lm / cnt[r];
        }
        *loss = (float)(total / 3.0);
    }
}

// ------------------------------------------------------------------ masks (elementwise)
// The bin ops are issue-bound (ncu: the precise-libm 'E' mask ran at 82 % issue utilisation and
// only 54 % of HBM), so they use MUFU-based forms whose error (<= 4e-7 relative) sits two orders
// below the 1e-4 budget: rsqrt for 1/|z| and sqrt, exp-based tanh with a series below 0.3.
__device__ __forceinline__ float fast_tanh(float x) {
    // branch-free 13/6 rational minimax (the classic single-precision form; max relative error
    // 4e-7 against float64 tanh, checked over [-10, 10] and 1e-8..10)
    x = fminf(fmaxf(x, -7.90531110763549805f), 7.90531110763549805f);
    const float x2 = x * x;
    float p = -2.76076847742355e-16f;
    p = fmaf(p, x2, 2.00018790482477e-13f);
    p = fmaf(p, x2, -8.60467152213735e-11f);
    p = fmaf(p, x2, 5.12229709037114e-08f);
    p = fmaf(p, x2, 1.48572235717979e-05f);
    p = fmaf(p, x2, 6.37261928875436e-04f);
    p = fmaf(p, x2, 4.89352455891786e-03f);
    float q = 1.19825839466702e-06f;
    q = fmaf(q, x2, 1.18534705686654e-04f);
    q = fmaf(q, x2, 2.26843463243900e-03f);
    q = fmaf(q, x2, 4.89352518554385e-03f);
    return p * x * se_rcp(q);                 // q in [4.9e-3, 0.9]
}

struct MaskMath {
    // unit vector of v (p = |v|^2); at the origin (+-1, 0) like atan2(+-0, +0) = 0 and atan2(+-0, -0) = +-pi
    __device__ static __forceinline__ float2 unit(float2 v, float p) {
        // branch-free: rescale (exactly, by 2^60) when |v|^2 would underflow, so the phase of tiny
        // non-zero values survives like it does through atan2
        const float sc = p < 1e-30f ? 1.15292150460684698e18f : 1.f;
        const float vx = v.x * sc, vy = v.y * sc;
        const float p2 = vx * vx + vy * vy;
        const float r = se_rsqrt(fmaxf(p2, 1e-37f));
        return p2 > 0.f ? make_float2(vx * r, vy * r) : make_float2(copysignf(1.f, v.x), 0.f);
    }
    // phi = tanh(r)/r as a function of pm = r^2: the same 13/6 rational as fast_tanh (tanh r = r P(r^2)/Q(r^2)),
    // so no square root is needed; beyond the rational's clamp tanh = 1 and phi = 1/r.
    __device__ static __forceinline__ float tanh_over_r(float pm) {
        const float x2 = fminf(pm, 62.4939437f);             // 7.90531110763549805^2
        float p = -2.76076847742355e-16f;
        p = fmaf(p, x2, 2.00018790482477e-13f);
        p = fmaf(p, x2, -8.60467152213735e-11f);
        p = fmaf(p, x2, 5.12229709037114e-08f);
        p = fmaf(p, x2, 1.48572235717979e-05f);
        p = fmaf(p, x2, 6.37261928875436e-04f);
        p = fmaf(p, x2, 4.89352455891786e-03f);
        float q = 1.19825839466702e-06f;
        q = fmaf(q, x2, 1.18534705686654e-04f);
        q = fmaf(q, x2, 2.26843463243900e-03f);
        q = fmaf(q, x2, 4.89352518554385e-03f);
        const float v = p * se_rcp(q);
        return pm > 62.4939437f ? se_rsqrt(pm) : v;
    }
    // y = f(x, m); m already squashed.  E: tanh(|m|) sqrt(|x|^2+1e-8) (x/|x|)(m/|m|) = phi(|m|^2) sqrt(|x|^2+1e-8)
    // (x/|x|) m: one unit vector, one complex product, two scalar factors.  Branch-free on purpose: inside the
    // FFT kernels the bins of a task are independent instruction streams the scheduler interleaves, and a branch
    // per bin serialises them (measured: 95 -> 101 us for the fused tail with an `if (x != 0)` fast path).
    template <int MODE>
    __device__ static __forceinline__ float2 fwd(float2 x, float2 m) {
        if (MODE == 2) return cmul(x, m);
        if (MODE == 3) return make_float2(x.x * m.x, x.y * m.y);
        const float px = x.x * x.x + x.y * x.y, pm = m.x * m.x + m.y * m.y;
        const float2 ux = unit(x, px);
        const float pe = px + 1e-8f;
        const float c = tanh_over_r(pm) * (pe * se_rsqrt(pe));
        const float2 w = cmul(ux, m);
        return make_float2(c * w.x, c * w.y);
    }
    // gradient wrt m (squashed) and x, given gy
    template <int MODE>
    __device__ static __forceinline__ void bwd(float2 x, float2 m, float2 gy, float2& gm, float2& gx) {
        if (MODE == 2) { gm = cmulc(gy, x); gx = cmulc(gy, m); return; }
        if (MODE == 3) { gm = make_float2(x.x * gy.x, x.y * gy.y); gx = make_float2(m.x * gy.x, m.y * gy.y); return; }
        // y = h k with h = phi(pm) m and k = mag(px) ux
        const float px = x.x * x.x + x.y * x.y, pm = m.x * m.x + m.y * m.y;
        const float2 ux = unit(x, px);
        const float pe = px + 1e-8f, imag = se_rsqrt(pe), mag = pe * imag;
        const float phi = tanh_over_r(pm);
        // phi'(r)/r = (1 - phi - pm phi^2)/pm, by its series where that cancels
        const float series = fmaf(pm, fmaf(pm, -34.f / 105.f, 8.f / 15.f), -2.f / 3.f);
        const float closed = (1.f - phi - pm * phi * phi) * se_rcp(fminf(fmaxf(pm, 1e-2f), 1e30f));
        const float dphi_r = pm < 1e-2f ? series : closed;
        const float2 gh = cmulc(gy, make_float2(mag * ux.x, mag * ux.y));      // conj(k) gy
        const float dot = m.x * gh.x + m.y * gh.y;
        gm = make_float2(phi * gh.x + dphi_r * dot * m.x, phi * gh.y + dphi_r * dot * m.y);
        // k = psi(rho) x, psi = mag / rho; zero at (numerically) zero x like the reference's atan2
        const bool big = px > 1e-30f;
        const float irho = se_rsqrt(big ? px : 1.f), psi = mag * irho, dpsi_r = -1e-8f * irho * irho * irho * imag;
        const float2 gk = cmulc(gy, make_float2(phi * m.x, phi * m.y));
        const float dx = x.x * gk.x + x.y * gk.y;
        gx = big ? make_float2(psi * gk.x + dpsi_r * dx * x.x, psi * gk.y + dpsi_r * dx * x.y) : make_float2(0.f, 0.f);
    }
    // one complex bin, runtime-free: MODE 0 keeps the real mask in m.x
    template <int MODE, bool TANH>
    __device__ static __forceinline__ float2 apply(float2 x, float2 m) {
        if (MODE == 0) { const float g = TANH ? fast_tanh(m.x) : m.x; return make_float2(x.x * g, x.y * g); }
        if (TANH) m = make_float2(fast_tanh(m.x), fast_tanh(m.y));
        return fwd<MODE>(x, m);
    }
    // gradient wrt the RAW mask (MODE 0: result in .x) and wrt x
    template <int MODE, bool TANH>
    __device__ static __forceinline__ void grad(float2 x, float2 m, float2 gy, float2& gm, float2& gx) {
        if (MODE == 0) {
            const float g = TANH ? fast_tanh(m.x) : m.x;
            float d = x.x * gy.x + x.y * gy.y;
            if (TANH) d *= (1.f - g * g);
            gm = make_float2(d, 0.f);
            gx = make_float2(gy.x * g, gy.y * g);
            return;
        }
        if (TANH) m = make_float2(fast_tanh(m.x), fast_tanh(m.y));
        bwd<MODE>(x, m, gy, gm, gx);
        if (TANH) gm = make_float2(gm.x * (1.f - m.x * m.x), gm.y * (1.f - m.y * m.y));
    }
};

// Two bins per thread per iteration with 128-bit accesses (count is even for every supported
// spectrum shape except odd F*T, handled by the scalar tail).
template <int MODE, bool TANH>
__global__ void __launch_bounds__(256) k_mask_fwd_t(const float2* __restrict__ spec, const float* __restrict__ mask,
                                                    float2* __restrict__ out, int64_t count) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(spec) | reinterpret_cast<uintptr_t>(mask) |
                       reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t pairs = vec ? count / 2 : 0;
    for (int64_t i = tid0; i < pairs; i += stride) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(spec) + i);
        float2 m0, m1;
        if (MODE == 0) {
            const float2 m = __ldg(reinterpret_cast<const float2*>(mask) + i);
            m0 = make_float2(m.x, 0.f); m1 = make_float2(m.y, 0.f);
        } else {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask) + i);
            m0 = make_float2(m.x, m.y); m1 = make_float2(m.z, m.w);
        }
        const float2 y0 = MaskMath::apply<MODE, TANH>(make_float2(x.x, x.y), m0);
        const float2 y1 = MaskMath::apply<MODE, TANH>(make_float2(x.z, x.w), m1);
        reinterpret_cast<float4*>(out)[i] = make_float4(y0.x, y0.y, y1.x, y1.y);
    }
    for (int64_t i = 2 * pairs + tid0; i < count; i += stride) {
        const float2 x = __ldg(spec + i);
        const float2 m = MODE == 0 ? make_float2(__ldg(mask + i), 0.f) : __ldg(reinterpret_cast<const float2*>(mask) + i);
        out[i] = MaskMath::apply<MODE, TANH>(x, m);
    }
}

template <int MODE, bool TANH>
__global__ void __launch_bounds__(256) k_mask_bwd_t(const float2* __restrict__ spec, const float* __restrict__ mask,
                                                    const float2* __restrict__ gout, float* __restrict__ gmask,
                                                    float2* __restrict__ gspec, int64_t count) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(spec) | reinterpret_cast<uintptr_t>(mask) | reinterpret_cast<uintptr_t>(gout) |
                       reinterpret_cast<uintptr_t>(gmask) | reinterpret_cast<uintptr_t>(gspec)) & 15) == 0;
    const int64_t pairs = vec ? count / 2 : 0;
    for (int64_t i = tid0; i < pairs; i += stride) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(spec) + i);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gout) + i);
        float2 m0, m1;
        if (MODE == 0) {
            const float2 m = __ldg(reinterpret_cast<const float2*>(mask) + i);
            m0 = make_float2(m.x, 0.f); m1 = make_float2(m.y, 0.f);
        } else {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask) + i);
            m0 = make_float2(m.x, m.y); m1 = make_float2(m.z, m.w);
        }
        float2 gm0, gx0, gm1, gx1;
        MaskMath::grad<MODE, TANH>(make_float2(x.x, x.y), m0, make_float2(g.x, g.y), gm0, gx0);
        MaskMath::grad<MODE, TANH>(make_float2(x.z, x.w), m1, make_float2(g.z, g.w), gm1, gx1);
        if (MODE == 0) reinterpret_cast<float2*>(gmask)[i] = make_float2(gm0.x, gm1.x);
        else reinterpret_cast<float4*>(gmask)[i] = make_float4(gm0.x, gm0.y, gm1.x, gm1.y);
        if (gspec) reinterpret_cast<float4*>(gspec)[i] = make_float4(gx0.x, gx0.y, gx1.x, gx1.y);
    }
    for (int64_t i = 2 * pairs + tid0; i < count; i += stride) {
        const float2 x = __ldg(spec + i), gy = __ldg(gout + i);
        const float2 m = MODE == 0 ? make_float2(__ldg(mask + i), 0.f) : __ldg(reinterpret_cast<const float2*>(mask) + i);
        float2 gm, gx;
        MaskMath::grad<MODE, TANH>(x, m, gy, gm, gx);
        if (MODE == 0) gmask[i] = gm.x; else reinterpret_cast<float2*>(gmask)[i] = gm;
        if (gspec) gspec[i] = gx;
    }
}


// DCCRN layout (src/model/dccrn.py:147-223): spectrum planar [rows][2F][T] (Re bins, then Im bins), masks two planes
// [rows][F][T].  Same per-bin math as above; no stack / cat copies around it.  plane = F*T elements.
template <int MODE>
__global__ void __launch_bounds__(256) k_mask_planar_fwd_t(const float* __restrict__ spec, const float* __restrict__ mre,
                                                           const float* __restrict__ mim, float* __restrict__ out,
                                                           int64_t plane, int bpr) {
    pdl_launch_dependents();
    pdl_wait();
    // bpr blocks per row: one 32-bit divide per block instead of a 64-bit divide per element
    const int64_t row = blockIdx.x / bpr;
    const int chunk = blockIdx.x - (int)row * bpr;
    for (int64_t i = (int64_t)chunk * blockDim.x + threadIdx.x; i < plane; i += (int64_t)bpr * blockDim.x) {
        const int64_t e = row * plane + i, re = e + row * plane, im = re + plane;  // row*2*plane + i (+ plane)
        const float2 y = MaskMath::apply<MODE, false>(make_float2(__ldg(spec + re), __ldg(spec + im)),
                                                      make_float2(__ldg(mre + e), __ldg(mim + e)));
        out[re] = y.x;
        out[im] = y.y;
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_mask_planar_bwd_t(const float* __restrict__ spec, const float* __restrict__ mre,
                                                           const float* __restrict__ mim, const float* __restrict__ gout,
                                                           float* __restrict__ gmre, float* __restrict__ gmim,
                                                           float* __restrict__ gspec, int64_t plane, int bpr) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x / bpr;
    const int chunk = blockIdx.x - (int)row * bpr;
    for (int64_t i = (int64_t)chunk * blockDim.x + threadIdx.x; i < plane; i += (int64_t)bpr * blockDim.x) {
        const int64_t e = row * plane + i, re = e + row * plane, im = re + plane;
        float2 gm, gx;
        MaskMath::grad<MODE, false>(make_float2(__ldg(spec + re), __ldg(spec + im)), make_float2(__ldg(mre + e), __ldg(mim + e)),
                                    make_float2(__ldg(gout + re), __ldg(gout + im)), gm, gx);
        gmre[e] = gm.x;
        gmim[e] = gm.y;
        if (gspec) { gspec[re] = gx.x; gspec[im] = gx.y; }
    }
}

}  // namespace se
