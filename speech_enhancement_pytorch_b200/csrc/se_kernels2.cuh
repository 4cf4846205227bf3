// se_kernels2.cuh -- spectral kernels on the signal-pair FFT engine (se_fft2.cuh).
//
// Same coordinates, chunking and PDL protocol as se_kernels.cuh; what changes is the unit of work: a CTA
// walks groups of 8 frames of TWO signals.  For the loss forward the two signals are the reference and the
// estimate of one row (their magnitudes meet in the two halves of one register pair, nothing is kept across
// transforms); everywhere else they are two consecutive rows (an odd last row is paired with itself and its
// duplicate outputs are not stored).
#pragma once
#include "se_fft2.cuh"
#include "se_kernels.cuh"

namespace se {

// ------------------------------------------------------------------ tables in shared memory
// window / exp(-2 pi i k/M) / exp(-2 pi i k/n): M float2 entries each in global memory (se_host.cu), staged as
// (a, a, b, b) when G::DUP
template <class G>
__device__ __forceinline__ Tables2 stage_tables2(const Tables& g, unsigned char* dst, int tid) {
    const float2* src[3] = {reinterpret_cast<const float2*>(g.win), g.tw, g.twn};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned char* d = dst + k * G::TABLE_BYTES;
        if (G::DUP) {
            for (int i = tid; i < G::M; i += G::NT) {
                const float2 t = __ldg(src[k] + i);
                reinterpret_cast<float4*>(d)[i] = make_float4(t.x, t.x, t.y, t.y);
            }
        } else {
            for (int i = tid; i < G::M / 2; i += G::NT)
                reinterpret_cast<float4*>(d)[i] = __ldg(reinterpret_cast<const float4*>(src[k]) + i);
        }
    }
    Tables2 r;
    r.win = dst;
    r.tw = dst + G::TABLE_BYTES;
    r.twn = dst + 2 * G::TABLE_BYTES;
    r.w2 = g.w2;
    r.inv_env = g.inv_env;
    return r;
}
// a second window (fused kernels: analysis and synthesis windows differ, twiddles are shared)
template <class G>
__device__ __forceinline__ Tables2 stage_window2(const Tables2& staged, const Tables& g, unsigned char* dst, int tid) {
    const float2* src = reinterpret_cast<const float2*>(g.win);
    if (G::DUP) {
        for (int i = tid; i < G::M; i += G::NT) {
            const float2 t = __ldg(src + i);
            reinterpret_cast<float4*>(dst)[i] = make_float4(t.x, t.x, t.y, t.y);
        }
    } else {
        for (int i = tid; i < G::M / 2; i += G::NT)
            reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    }
    Tables2 r = staged;
    r.win = dst;
    r.w2 = g.w2;
    r.inv_env = g.inv_env;
    return r;
}

template <class G>
__device__ __forceinline__ float inv_env_at2(const Tables2& tb, int T, int i) {
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

// ------------------------------------------------------------------ stage fill
// what the padded signal looks like around the valid samples
struct FillArgs {
    int nsample;            // REFLECT / ZEROPAD: valid input samples N; ENV: natural padded length
    int in_len;             // ENV: `length` of the gy rows
    int nframe;             // ENV: T
    int pad;                // ZEROPAD: zeros in front
};

template <class G, int LMODE>
__device__ __forceinline__ float sample_padded(const float* __restrict__ src, const FillArgs& a, int i, int nvalid) {
    if (LMODE == LOAD_REFLECT) return sample_reflect(src, G::N / 2, a.nsample, G::N, i, nvalid);
    if (LMODE == LOAD_ZEROPAD) {
        const int j = i - a.pad;
        return (j >= 0 && j < a.nsample) ? __ldg(src + j) : 0.f;
    }
    const int q = i - G::N / 2;
    return (q >= 0 && q < a.in_len && i < a.nsample) ? __ldg(src + q) : 0.f;
}

// Fill the stage with padded coordinates [p0, p0 + SROWS*HOP) of both signals, interleaved (s0[j], s1[j]).
// 128-bit global loads, all of a thread's loads in flight before its first shared-memory store.
template <class G, int LMODE>
__device__ __forceinline__ void fill_stage2(float2* __restrict__ stage, const float* __restrict__ s0, const float* __restrict__ s1,
                                            int p0, const FillArgs& a, const Tables2& tb, int tid,
                                            int nvalid0 = 0x7fffffff, int nvalid1 = 0x7fffffff) {
    constexpr int SLOTS = G::SROWS * G::HOP / 4;                 // float4 slots per signal
    constexpr int K = (SLOTS + G::NT - 1) / G::NT;
    float4 v0[K], v1[K];
    const int base = (LMODE == LOAD_ZEROPAD) ? p0 - a.pad : p0 - G::N / 2;   // source index of slot 0
    const int lim = (LMODE == LOAD_ENV) ? a.in_len : a.nsample;
    const int nv = nvalid0 < nvalid1 ? nvalid0 : nvalid1;
    const int limit = lim < nv ? lim : nv;
    const bool interior = base >= 0 && base + 4 * SLOTS <= limit && (LMODE != LOAD_ENV || p0 + 4 * SLOTS <= a.nsample) &&
                          ((reinterpret_cast<uintptr_t>(s0 + base) | reinterpret_cast<uintptr_t>(s1 + base)) & 15) == 0;
    if (interior) {                                              // uniform per CTA
        const float4* q0 = reinterpret_cast<const float4*>(s0 + base);
        const float4* q1 = reinterpret_cast<const float4*>(s1 + base);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int slot = tid + k * G::NT;
            if (slot < SLOTS) { v0[k] = __ldg(q0 + slot); v1[k] = __ldg(q1 + slot); }
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int slot = tid + k * G::NT;
            if (slot >= SLOTS) continue;
            const int i = p0 + 4 * slot;
            v0[k] = make_float4(sample_padded<G, LMODE>(s0, a, i, nvalid0), sample_padded<G, LMODE>(s0, a, i + 1, nvalid0),
                                sample_padded<G, LMODE>(s0, a, i + 2, nvalid0), sample_padded<G, LMODE>(s0, a, i + 3, nvalid0));
            v1[k] = make_float4(sample_padded<G, LMODE>(s1, a, i, nvalid1), sample_padded<G, LMODE>(s1, a, i + 1, nvalid1),
                                sample_padded<G, LMODE>(s1, a, i + 2, nvalid1), sample_padded<G, LMODE>(s1, a, i + 3, nvalid1));
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int slot = tid + k * G::NT;
        if (slot >= SLOTS) continue;
        const int rel = 4 * slot;
        float4 a0 = v0[k], a1 = v1[k];
        if (LMODE == LOAD_ENV) {                                 // gy / envelope (zero where the envelope is empty)
            const int i = p0 + rel;
            const float e0 = inv_env_at2<G>(tb, a.nframe, i), e1 = inv_env_at2<G>(tb, a.nframe, i + 1);
            const float e2 = inv_env_at2<G>(tb, a.nframe, i + 2), e3 = inv_env_at2<G>(tb, a.nframe, i + 3);
            a0 = make_float4(a0.x * e0, a0.y * e1, a0.z * e2, a0.w * e3);
            a1 = make_float4(a1.x * e0, a1.y * e1, a1.z * e2, a1.w * e3);
        }
        float4* d = reinterpret_cast<float4*>(stage + (rel / G::HOP) * G::SROW + rel % G::HOP);
        d[0] = make_float4(a0.x, a1.x, a0.y, a1.y);
        d[1] = make_float4(a0.z, a1.z, a0.w, a1.w);
    }
}

// ------------------------------------------------------------------ transform building blocks
template <class G>
__device__ __forceinline__ void analysis_passes2(const float2* stage, const Tables2& tb, float4* zb, int unit, int fr) {
    passA_fwd2<G>(stage, tb, zb, unit, fr);
    __syncthreads();
    passB2<G, false>(tb, zb, unit, fr);
    __syncthreads();
}
template <class G>
__device__ __forceinline__ void analysis_task2(const float4* zb, const Tables2& tb, int p, int fr, c2* xa, c2* xb, c2& nyq) {
    passC_fwd_unit2<G>(zb, task_qa2<G>(p), fr, xa);
    passC_fwd_unit2<G>(zb, task_qb2<G>(p), fr, xb);
    split_task2<G>(p, tb, xa, xb, nyq);
}
template <class G>
__device__ __forceinline__ void synthesis_task2(float4* zb, const Tables2& tb, int p, int fr, c2* ya, c2* yb, c2 nyq) {
    merge_task2<G>(p, tb, ya, yb, nyq);
    passC_inv_unit2<G>(zb, task_qa2<G>(p), fr, ya);
    passC_inv_unit2<G>(zb, task_qb2<G>(p), fr, yb);
}
template <class G, bool CARRY = true>
__device__ __forceinline__ void synthesis_tail2(float4* zb, const Tables2& tb, float2* ostage, int unit, int fr,
                                                c2 (*carry)[G::SEG]) {
    __syncthreads();
    passB2<G, true>(tb, zb, unit, fr);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < G::TA; ++i) {
        const int u = unit + i * G::NU;
        c2 v[G::R1], acc[G::SEG];
        passA_inv_task2<G>(zb, tb, u, fr, v);
        ola_rotate2<G, CARRY>(v, fr, CARRY ? carry[i] : nullptr, acc);
#pragma unroll
        for (int s = 0; s < G::SEG; ++s)
            *reinterpret_cast<float4*>(ostage + fr * G::SROW + 2 * (u + 64 * s)) = to4(acc[s]);
    }
    __syncthreads();
}

// ------------------------------------------------------------------ emitters (two output rows)
template <class G>
__device__ __forceinline__ void emit_istft2(const float2* __restrict__ ostage, float* __restrict__ y0, float* __restrict__ y1,
                                            int f_base, const Chunk& c, const SynArgs& a, const Tables2& tb, int tid) {
    const bool vec = ((reinterpret_cast<uintptr_t>(y0) | reinterpret_cast<uintptr_t>(y1 ? y1 : y0)) & 15) == 0;
    for (int idx = tid; idx < G::FR * G::HOP / 4; idx += G::NT) {
        const int e = 4 * idx;
        const int blk = e / G::HOP, o = e - blk * G::HOP;
        const int b = f_base + blk;
        if (b < c.b0 || b >= c.b1) continue;
        const int i = b * G::HOP + o, s = i - G::N / 2;
        const float4 p = *reinterpret_cast<const float4*>(ostage + blk * G::SROW + o);
        const float4 q = *reinterpret_cast<const float4*>(ostage + blk * G::SROW + o + 2);
        if (vec && s >= 0 && s + 3 < a.out_len && i + 3 < a.nsample && b >= G::OLA - 1 && b <= a.nframe - 1) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(tb.inv_env + o));
            *reinterpret_cast<float4*>(y0 + s) = make_float4(p.x * w.x, p.z * w.y, q.x * w.z, q.z * w.w);
            if (y1) *reinterpret_cast<float4*>(y1 + s) = make_float4(p.y * w.x, p.w * w.y, q.y * w.z, q.w * w.w);
        } else {
            const float v0[4] = {p.x, p.z, q.x, q.z}, v1[4] = {p.y, p.w, q.y, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int ii = i + k, ss = s + k;
                if (ss < 0 || ss >= a.out_len) continue;
                const float w = ii < a.nsample ? inv_env_at2<G>(tb, a.nframe, ii) : 0.f;
                y0[ss] = ii < a.nsample ? v0[k] * w : 0.f;
                if (y1) y1[ss] = ii < a.nsample ? v1[k] * w : 0.f;
            }
        }
    }
}

// adjoint emitter: fold the reflect padding back (see emit_adj in se_kernels.cuh).  Chunks hold >= 7 blocks
// (plan_synthesis2), so the left mirror's sources and destinations share the row's first group and the
// right-edge zone lies inside the last chunk.
template <class G>
__device__ __forceinline__ void emit_adj2(const float2* __restrict__ ostage, float2* __restrict__ hold,
                                          float* __restrict__ g0, float* __restrict__ g1, int f_base, const Chunk& c,
                                          int N, int accumulate, int tid) {
    constexpr int NH = G::N / 2;
    constexpr int K = G::FR * G::HOP / G::NT;
    static_assert(G::FR * G::HOP % G::NT == 0, "emit tiling");
    const int zs = ((N - 1) / G::HOP) * G::HOP;
    float2 v[K], old[K];
    int dst[K];                                               // >= 0: gx index, -1: nothing, <= -2: hold slot -(dst+2)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int idx = tid + k * G::NT;
        const int blk = idx / G::HOP, o = idx - blk * G::HOP;
        const int b = f_base + blk;
        const int i = b * G::HOP + o;
        dst[k] = -1;
        v[k] = make_float2(0.f, 0.f);
        if (b < c.b0 || b >= c.b1 || i >= N + G::N) continue;
        float2 val = ostage[blk * G::SROW + o];
        if (i > NH && i <= G::N) {                       // left mirror: x[j] also fed p[n/2 - j]
            const int is = G::N - i;
            const int sb = is / G::HOP - f_base;
            const float2 m = ostage[sb * G::SROW + is % G::HOP];
            val.x += m.x; val.y += m.y;
        }
        v[k] = val;
        if (c.last && i >= zs) dst[k] = -2 - (i - zs);
        else if (i >= NH) dst[k] = i - NH;
    }
    if (accumulate) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            old[k].x = dst[k] >= 0 ? g0[dst[k]] : 0.f;
            old[k].y = (dst[k] >= 0 && g1) ? g1[dst[k]] : 0.f;
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (dst[k] >= 0) {
            g0[dst[k]] = accumulate ? old[k].x + v[k].x : v[k].x;
            if (g1) g1[dst[k]] = accumulate ? old[k].y + v[k].y : v[k].y;
        } else if (dst[k] <= -2) hold[-(dst[k] + 2)] = v[k];
    }
}
template <class G>
__device__ __forceinline__ void finish_adj2(const float2* __restrict__ hold, float* __restrict__ g0, float* __restrict__ g1,
                                            int N, int accumulate, int tid) {
    constexpr int NH = G::N / 2;
    const int zs = ((N - 1) / G::HOP) * G::HOP;
    for (int i = zs + tid; i < N + NH; i += G::NT) {
        float2 v = hold[i - zs];
        if (i >= N - 1 && i <= N + NH - 2) {                       // right mirror
            const float2 m = hold[(2 * N + G::N - 2 - i) - zs];
            v.x += m.x; v.y += m.y;
        }
        const int j = i - NH;
        g0[j] = accumulate ? g0[j] + v.x : v.x;
        if (g1) g1[j] = accumulate ? g1[j] + v.y : v.y;
    }
}

// ------------------------------------------------------------------ spectrum rows of the two signals, [F][T] float2
template <class G>
__device__ __forceinline__ void store_task_pair(float2* __restrict__ r0, float2* __restrict__ r1, int T, int t, int p,
                                                const c2* xa, const c2* xb, c2 nyq, float edge) {
    if (t < 0 || t >= T) return;
    const size_t step = (size_t)G::S * T;
    const size_t ia = (size_t)task_qa2<G>(p) * T + t, ib = (size_t)task_qb2<G>(p) * T + t;
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        c2 va = xa[k4];
        if (p == 0 && k4 == 0) va = mk2(p_mul(va.re, p_dup(edge)), make_float2(0.f, 0.f));
        r0[ia + k4 * step] = make_float2(va.re.x, va.im.x);
        r0[ib + k4 * step] = make_float2(xb[k4].re.x, xb[k4].im.x);
        if (r1) {
            r1[ia + k4 * step] = make_float2(va.re.y, va.im.y);
            r1[ib + k4 * step] = make_float2(xb[k4].re.y, xb[k4].im.y);
        }
    }
    if (p == 0) {
        r0[(size_t)G::M * T + t] = make_float2(nyq.re.x * edge, 0.f);
        if (r1) r1[(size_t)G::M * T + t] = make_float2(nyq.re.y * edge, 0.f);
    }
}
template <class G>
__device__ __forceinline__ void load_task_pair(const float2* __restrict__ r0, const float2* __restrict__ r1, int T, int t, int p,
                                               c2* ya, c2* yb, c2& nyq, float edge) {
    const bool ok = (t >= 0 && t < T);
    const int tc = ok ? t : 0;
    const size_t step = (size_t)G::S * T;
    const size_t ia = (size_t)task_qa2<G>(p) * T + tc, ib = (size_t)task_qb2<G>(p) * T + tc;
    float2 a0[8], a1[8], b0[8], b1[8];
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        a0[k4] = __ldg(r0 + ia + k4 * step);
        b0[k4] = __ldg(r0 + ib + k4 * step);
        a1[k4] = __ldg(r1 + ia + k4 * step);
        b1[k4] = __ldg(r1 + ib + k4 * step);
    }
    float2 n0 = make_float2(0.f, 0.f), n1 = n0;
    if (p == 0) { n0 = __ldg(r0 + (size_t)G::M * T + tc); n1 = __ldg(r1 + (size_t)G::M * T + tc); }
    const float keep = ok ? 1.f : 0.f;
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        ya[k4] = mk2(make_float2(a0[k4].x * keep, a1[k4].x * keep), make_float2(a0[k4].y * keep, a1[k4].y * keep));
        yb[k4] = mk2(make_float2(b0[k4].x * keep, b1[k4].x * keep), make_float2(b0[k4].y * keep, b1[k4].y * keep));
    }
    nyq = mk2(make_float2(n0.x * keep, n1.x * keep), make_float2(n0.y * keep, n1.y * keep));
    if (p == 0) {
        ya[0].re = p_mul(ya[0].re, p_dup(edge));
        nyq.re = p_mul(nyq.re, p_dup(edge));
    }
}

// ------------------------------------------------------------------ smem carve-up (bytes)
template <class G> struct Smem2 {
    static constexpr size_t ANALYSIS = G::ZB_BYTES + G::STAGE_BYTES + G::TABLES_BYTES;
    static constexpr size_t SYNTH_ISTFT = G::ZB_BYTES + G::OSTAGE_BYTES + G::TABLES_BYTES;
    static constexpr size_t SYNTH_ADJ = G::ZB_BYTES + G::OSTAGE_BYTES + G::HOLD_BYTES + G::TABLES_BYTES;
    static constexpr size_t FUSED_ADJ = G::ZB_BYTES + G::IOBUF_BYTES + G::HOLD_BYTES + G::TABLES_BYTES;
    static constexpr size_t FUSED_ISTFT = G::ZB_BYTES + G::IOBUF_BYTES + G::TABLES_BYTES + G::TABLE_BYTES;
};

// ================================================================== kernels
// two wave-like rows -> two spectra
template <class G, int LMODE>
__global__ void __launch_bounds__(G::NT, G::MINB) k_analysis2(const AnaArgs a, int rows) {
    SE_SMEM_DECL;
    float4* zb = reinterpret_cast<float4*>(se_smem);
    float2* stage = reinterpret_cast<float2*>(se_smem + G::ZB_BYTES);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables2 tb = stage_tables2<G>(a.tb, se_smem + G::ZB_BYTES + G::STAGE_BYTES, tid);   // visible after the fill's barrier
    pdl_wait();
    const int prow = blockIdx.x / a.nchunks, chunk = blockIdx.x - prow * a.nchunks;
    const int r0 = 2 * prow, r1 = r0 + 1 < rows ? r0 + 1 : r0;
    const bool has1 = r0 + 1 < rows;
    // segments (evaluate()): row = seg * seg_rows + clip, strided views of the padded clips
    const float* src[2];
    int nvalid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int row = h ? r1 : r0;
        const int seg = row / a.seg_rows, clip = row - seg * a.seg_rows;
        src[h] = a.in + (size_t)seg * a.in_stride + (size_t)clip * a.clip_stride;
        nvalid[h] = 0x7fffffff;
        if (a.clip_len > 0) {
            const int64_t left = (int64_t)a.clip_len - (int64_t)seg * a.in_stride;
            nvalid[h] = left < 0 ? 0 : (left < a.nsample ? (int)left : a.nsample);
        }
    }
    FillArgs fa;
    fa.nsample = a.nsample; fa.in_len = a.in_len; fa.nframe = a.nframe; fa.pad = a.pad;
    float2* out0 = reinterpret_cast<float2*>(a.out) + (size_t)r0 * G::F * a.nframe;
    float2* out1 = has1 ? reinterpret_cast<float2*>(a.out) + (size_t)r1 * G::F * a.nframe : nullptr;
    for (int g = 0; g < a.gpc; ++g) {
        const int f_base = (chunk * a.gpc + g) * G::FR;
        if (f_base >= a.nframe) break;
        fill_stage2<G, LMODE>(stage, src[0], src[1], f_base * G::HOP, fa, tb, tid, nvalid[0], nvalid[1]);
        __syncthreads();
        analysis_passes2<G>(stage, tb, zb, unit, fr);
        const int t = f_base + fr;
#pragma unroll 1
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            c2 xa[8], xb[8], nyq;
            analysis_task2<G>(zb, tb, p, fr, xa, xb, nyq);
            store_task_pair<G>(out0, out1, a.nframe, t, p, xa, xb, nyq, a.edge_scale);
        }
        // no trailing barrier: the next fill only writes the stage (last read before pass A's barrier), and nobody
        // passes the fill's barrier before every thread has left pass C
    }
}

// two spectra -> two wave-like rows.  EMODE: ISTFT (envelope + trim) or ADJ (reflect fold-back).
template <class G, int EMODE>
__global__ void __launch_bounds__(G::NT, G::MINB) k_synthesis2(const SynArgs a, int rows) {
    SE_SMEM_DECL;
    float4* zb = reinterpret_cast<float4*>(se_smem);
    float2* ostage = reinterpret_cast<float2*>(se_smem + G::ZB_BYTES);
    float2* hold = reinterpret_cast<float2*>(se_smem + G::ZB_BYTES + G::OSTAGE_BYTES);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables2 tb = stage_tables2<G>(a.tb, se_smem + G::ZB_BYTES + G::OSTAGE_BYTES + (EMODE == EMIT_ADJ ? G::HOLD_BYTES : 0), tid);
    __syncthreads();
    pdl_wait();
    const int prow = blockIdx.x / a.nchunks, chunk = blockIdx.x - prow * a.nchunks;
    const int r0 = 2 * prow, r1 = r0 + 1 < rows ? r0 + 1 : r0;
    const bool has1 = r0 + 1 < rows;
    const Chunk c = make_chunk<G>(chunk, a.nchunks, a.b_lo, a.b_hi);
    const float2* spec0 = reinterpret_cast<const float2*>(a.in) + (size_t)r0 * G::F * a.nframe;
    const float2* spec1 = reinterpret_cast<const float2*>(a.in) + (size_t)r1 * G::F * a.nframe;
    float* out0 = a.out + (size_t)r0 * a.out_len;
    float* out1 = has1 ? a.out + (size_t)r1 * a.out_len : nullptr;
    c2 carry[G::TA][G::SEG];
#pragma unroll
    for (int i = 0; i < G::TA; ++i)
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) carry[i][s] = zero2();
    for (int g = 0; g < c.ngroups; ++g) {
        const int f_base = c.f0 + g * G::FR;
        const int t = f_base + fr;
#pragma unroll 1
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            c2 ya[8], yb[8], nyq;
            load_task_pair<G>(spec0, spec1, a.nframe, t, p, ya, yb, nyq, a.edge_scale);
            synthesis_task2<G>(zb, tb, p, fr, ya, yb, nyq);
        }
        synthesis_tail2<G>(zb, tb, ostage, unit, fr, carry);
        if (EMODE == EMIT_ISTFT) emit_istft2<G>(ostage, out0, out1, f_base, c, a, tb, tid);
        else emit_adj2<G>(ostage, hold, out0, out1, f_base, c, a.nsample, a.accumulate, tid);
    }
    if (EMODE == EMIT_ADJ && c.last) {
        __syncthreads();
        finish_adj2<G>(hold, out0, out1, a.nsample, a.accumulate, tid);
    }
}

// ------------------------------------------------------------------ MR-STFT loss
// Forward: the pair is (reference, estimate) of one row.  |B|^2 and |A|^2 of a bin arrive in the two halves of one
// register pair, so the statistics need no values kept across transforms (the scalar kernel holds TC*17 of them).
template <class G>
__global__ void __launch_bounds__(G::NT, G::MINB) k_loss_fwd2(const LossArgs a) {
    SE_SMEM_DECL;
    float4* zb = reinterpret_cast<float4*>(se_smem);
    float2* stage = reinterpret_cast<float2*>(se_smem + G::ZB_BYTES);
    __shared__ float red[3][32];
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    if (a.chained) pdl_launch_dependents();
    const Tables2 tb = stage_tables2<G>(a.tb, se_smem + G::ZB_BYTES + G::STAGE_BYTES, tid);
    if (!a.chained) { pdl_wait(); pdl_launch_dependents(); }
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    FillArgs fa;
    fa.nsample = a.nsample; fa.in_len = a.nsample; fa.nframe = a.nframe; fa.pad = 0;
    const float* ref = a.ref + (size_t)row * a.nsample;
    const float* est = a.est + (size_t)row * a.nsample;
    float s_d2 = 0.f, s_b2 = 0.f, s_lm = 0.f;
    for (int g = 0; g < a.gpc; ++g) {
        const int f_base = (chunk * a.gpc + g) * G::FR;
        if (f_base >= a.nframe) break;
        const int t = f_base + fr;
        fill_stage2<G, LOAD_REFLECT>(stage, ref, est, f_base * G::HOP, fa, tb, tid);
        __syncthreads();
        analysis_passes2<G>(stage, tb, zb, unit, fr);
#pragma unroll 1
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            c2 xa[8], xb[8], nyq;
            analysis_task2<G>(zb, tb, p, fr, xa, xb, nyq);
            float* mrow = a.refmag + (size_t)row * G::F * a.nframe + t;
            const size_t mstep = (size_t)G::S * a.nframe;
            float* wp = mrow + (size_t)task_qa2<G>(p) * a.nframe;
#pragma unroll
            for (int k = 0; k < 17; ++k) {
                const c2 v = k < 8 ? xa[k] : (k < 16 ? xb[k - 8] : nyq);
                const float2 pw = p_fma(v.im, v.im, p_mul(v.re, v.re));          // (|B|^2, |A|^2)
                if (k == 8) wp = mrow + (size_t)task_qb2<G>(p) * a.nframe;
                if (k == 16) wp = mrow + (size_t)G::M * a.nframe;
                if (t < a.nframe && !(k == 16 && p != 0)) {
                    const float cb = fmaxf(pw.x, SE_MRSTFT_CLAMP), ca = fmaxf(pw.y, SE_MRSTFT_CLAMP);
                    // rounded products: identical spectra must give d == 0 exactly (loss(x, x) = 0, zero gradient)
                    const float mb = se_mul_rn(cb, se_rsqrt(cb));
                    const float d = mb - se_mul_rn(ca, se_rsqrt(ca));
                    *wp = mb;                                                    // |B| for the backward pass
                    s_d2 += d * d;
                    s_b2 += cb;
                    s_lm += 0.34657359f * fabsf(se_log2(cb) - se_log2(ca));     // |log b - log a| = ln2/2 |log2 cb - log2 ca|
                }
                wp += mstep;
            }
        }
    }
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

// Backward: the pair is two rows of the estimate.  Re-transform, form G = coef * A per bin from |B| (saved by the
// forward pass) and the three global sums, run the STFT adjoint (synthesis + reflect fold) in the same kernel.
template <class G>
__global__ void __launch_bounds__(G::NT, G::MINB) k_loss_bwd2(const LossArgs a, int rows) {
    SE_SMEM_DECL;
    float4* zb = reinterpret_cast<float4*>(se_smem);
    float2* iobuf = reinterpret_cast<float2*>(se_smem + G::ZB_BYTES);            // stage, later ostage
    float2* hold = reinterpret_cast<float2*>(se_smem + G::ZB_BYTES + G::IOBUF_BYTES);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    if (a.chained) pdl_launch_dependents();
    const Tables2 tb = stage_tables2<G>(a.tb, se_smem + G::ZB_BYTES + G::IOBUF_BYTES + G::HOLD_BYTES, tid);
    if (!a.chained) { pdl_wait(); pdl_launch_dependents(); }
    bool must_wait = a.chained != 0;
    const int prow = blockIdx.x / a.nchunks, chunk = blockIdx.x - prow * a.nchunks;
    const int r0 = 2 * prow, r1 = r0 + 1 < rows ? r0 + 1 : r0;
    const bool has1 = r0 + 1 < rows;
    const Chunk c = make_chunk<G>(chunk, a.nchunks, a.b_lo, a.b_hi);
    FillArgs fa;
    fa.nsample = a.nsample; fa.in_len = a.nsample; fa.nframe = a.nframe; fa.pad = 0;
    // dL/da = gs * [ alpha (a - b) + beta sign(a - b) / a ]
    const double d2 = a.sums[0], b2 = a.sums[1];
    const float gs = __ldg(a.gout) * a.inv_res;
    const float alpha = (d2 > 0.0 && b2 > 0.0) ? gs * (float)(1.0 / (sqrt(d2) * sqrt(b2))) : 0.f;
    const float beta = gs * (a.rows_dev ? (float)(1.0 / (a.rows_dev[0] * a.bins_per_row)) : a.inv_count);
    const float* e0 = a.est + (size_t)r0 * a.nsample;
    const float* e1 = a.est + (size_t)r1 * a.nsample;
    float* g0 = a.g_est + (size_t)r0 * a.nsample;
    float* g1 = has1 ? a.g_est + (size_t)r1 * a.nsample : nullptr;
    c2 carry[G::TA][G::SEG];
#pragma unroll
    for (int i = 0; i < G::TA; ++i)
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) carry[i][s] = zero2();
    for (int g = 0; g < c.ngroups; ++g) {
        const int f_base = c.f0 + g * G::FR;
        const int t = f_base + fr;
        const bool live = (t >= 0 && t < a.nframe);
        const int tc = live ? t : 0;
        const float alpha_l = live ? alpha : 0.f, beta_l = live ? beta : 0.f;
        const float* m0 = a.refmag + (size_t)r0 * G::F * a.nframe + tc;
        const float* m1 = a.refmag + (size_t)r1 * G::F * a.nframe + tc;
        const size_t mstep = (size_t)G::S * a.nframe;
        fill_stage2<G, LOAD_REFLECT>(iobuf, e0, e1, f_base * G::HOP, fa, tb, tid);
        __syncthreads();
        analysis_passes2<G>(iobuf, tb, zb, unit, fr);
#pragma unroll 1
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            // |B| of both rows in two batches of 16 + 1 loads: the first is in flight while pass C runs, the second is
            // issued when the first has been consumed (all 34 at once cost 50 spilled registers under the 128 cap)
            const size_t oa = (size_t)task_qa2<G>(p) * a.nframe, ob = (size_t)task_qb2<G>(p) * a.nframe;
            float2 mb[8], mn;
#pragma unroll
            for (int k = 0; k < 8; ++k) mb[k] = make_float2(__ldg(m0 + oa + k * mstep), __ldg(m1 + oa + k * mstep));
            mn = make_float2(__ldg(m0 + (size_t)G::M * a.nframe), __ldg(m1 + (size_t)G::M * a.nframe));
            c2 xa[8], xb[8], nyq;
            analysis_task2<G>(zb, tb, p, fr, xa, xb, nyq);
            // G = coef * A: branch-free -- below the clamp the magnitude is constant (zero gradient), dead frames carry
            // zero weights (alpha_l, beta_l); coef = alpha (ma - mb) / ma + beta sign(ma - mb) / ma^2
            auto weigh = [&](c2 v, float2 m, float edge) {
                const float2 pa = p_fma(v.im, v.im, p_mul(v.re, v.re));
                const float2 ia = make_float2(se_rsqrt(fmaxf(pa.x, SE_MRSTFT_CLAMP)), se_rsqrt(fmaxf(pa.y, SE_MRSTFT_CLAMP)));
                const float2 df = p_sub(p_mul(pa, ia), m);
                const float2 sg = make_float2(df.x > 0.f ? beta_l : (df.x < 0.f ? -beta_l : 0.f),
                                              df.y > 0.f ? beta_l : (df.y < 0.f ? -beta_l : 0.f));
                float2 coef = p_mul(p_fma(sg, ia, p_mul(p_dup(alpha_l), df)), ia);
                coef.x = pa.x < SE_MRSTFT_CLAMP ? 0.f : coef.x * edge;
                coef.y = pa.y < SE_MRSTFT_CLAMP ? 0.f : coef.y * edge;
                return cscale(v, coef);
            };
            // edge bins enter the C2R with weight 2 (H = G / c_k, the 1/2 sits in the window); only p == 0 owns Nyquist
#pragma unroll
            for (int k = 0; k < 8; ++k) xa[k] = weigh(xa[k], mb[k], (p == 0 && k == 0) ? 2.f : 1.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) mb[k] = make_float2(__ldg(m0 + ob + k * mstep), __ldg(m1 + ob + k * mstep));
            nyq = weigh(nyq, mn, p == 0 ? 2.f : 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) xb[k] = weigh(xb[k], mb[k], 1.f);
            synthesis_task2<G>(zb, tb, p, fr, xa, xb, nyq);
        }
        synthesis_tail2<G>(zb, tb, iobuf, unit, fr, carry);
        if (must_wait) { pdl_wait(); must_wait = false; }
        emit_adj2<G>(iobuf, hold, g0, g1, f_base, c, a.nsample, a.accumulate, tid);
        __syncthreads();
    }
    if (c.last) finish_adj2<G>(hold, g0, g1, a.nsample, a.accumulate, tid);
}

}  // namespace se
