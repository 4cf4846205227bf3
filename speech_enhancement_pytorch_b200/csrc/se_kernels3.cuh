// se_kernels3.cuh -- spectral kernels on the two-pass FFT engine (se_fft3.cuh), n_fft 512 / 1024.
// Same coordinates, chunking, emitters and PDL protocol as se_kernels.cuh; a pass-C task now carries 2 x 16 bins.
#pragma once
#include "se_fft3.cuh"
#include "se_kernels.cuh"

namespace se {

template <class G>
__device__ __forceinline__ void analysis_task3(const float2* zb, const Tables& tb, int p, int fr, float2* xa, float2* xb, float2& nyq) {
    passC3_fwd_unit<G>(zb, task3_qa<G>(p), fr, xa);
    passC3_fwd_unit<G>(zb, task3_qb<G>(p), fr, xb);
    split_task3<G>(p, tb.twn, xa, xb, nyq);
}
template <class G>
__device__ __forceinline__ void synthesis_task3(float2* zb, const Tables& tb, int p, int fr, float2* ya, float2* yb, float2 nyq) {
    merge_task3<G>(p, tb.twn, ya, yb, nyq);
    passC3_inv_unit<G>(zb, task3_qa<G>(p), fr, ya);
    passC3_inv_unit<G>(zb, task3_qb<G>(p), fr, yb);
}
// pass A inverse + overlap-add by lane rotation into ostage (one barrier before, one after)
template <class G, bool CARRY = true>
__device__ __forceinline__ void synthesis_tail3(float2* zb, const Tables& tb, float* ostage, int unit, int fr,
                                                float2 (*carry)[G::SEG]) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < G::TA; ++i) {
        const int u = unit + i * G::NU;
        float2 v[G::R1], acc[G::SEG];
        passA3_inv_task<G>(zb, tb.win, tb.tw, u, fr, v);
        ola_rotate<G, CARRY>(v, fr, CARRY ? carry[i] : nullptr, acc);
#pragma unroll
        for (int s = 0; s < G::SEG; ++s)
            *reinterpret_cast<float2*>(ostage + fr * G::SROW + 2 * (u + G::RC * s)) = acc[s];
    }
    __syncthreads();
}

// spectrum rows, interleaved [F][T] float2: one pointer per unit walked by S*T (see store_task_ft2)
template <class G>
__device__ __forceinline__ void store_task3(float2* __restrict__ row, int T, int t, int p, const float2* xa, const float2* xb,
                                            float2 nyq, float edge) {
    if (t < 0 || t >= T) return;
    const size_t step = (size_t)G::S * T;
    float2* pa = row + (size_t)task3_qa<G>(p) * T + t;
    float2* pb = row + (size_t)task3_qb<G>(p) * T + t;
#pragma unroll
    for (int k = 0; k < G::RC; ++k) {
        *pa = (p == 0 && k == 0) ? make_float2(xa[0].x * edge, 0.f) : xa[k];
        *pb = xb[k];
        pa += step;
        pb += step;
    }
    if (p == 0) row[(size_t)G::M * T + t] = make_float2(nyq.x * edge, 0.f);
}
template <class G>
__device__ __forceinline__ void load_task3(const float2* __restrict__ row, int T, int t, int p, float2* ya, float2* yb,
                                           float2& nyq, float edge) {
    const bool ok = (t >= 0 && t < T);
    const int tc = ok ? t : 0;                                   // clamped: loads stay in bounds, result zeroed
    const size_t step = (size_t)G::S * T;
    const float2* pa = row + (size_t)task3_qa<G>(p) * T + tc;
    const float2* pb = row + (size_t)task3_qb<G>(p) * T + tc;
#pragma unroll
    for (int k = 0; k < G::RC; ++k) {
        ya[k] = __ldg(pa);
        yb[k] = __ldg(pb);
        pa += step;
        pb += step;
    }
    nyq = make_float2(0.f, 0.f);
    if (p == 0) nyq = __ldg(row + (size_t)G::M * T + tc);
    if (!ok) {
#pragma unroll
        for (int k = 0; k < G::RC; ++k) ya[k] = yb[k] = make_float2(0.f, 0.f);
        nyq = make_float2(0.f, 0.f);
    }
    if (p == 0) {
        ya[0].x *= edge;
        nyq.x *= edge;
    }
}

// ================================================================== kernels
template <class G, int LMODE>
__global__ void __launch_bounds__(G::NT, G::MINB) k_analysis3(const AnaArgs a) {
    using B = typename G::Base;
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* stage = reinterpret_cast<float*>(se_smem + Smem<B>::ZB);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<B>(a.tb, se_smem + Smem<B>::ZB + Smem<B>::STAGE, tid);
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const int seg = row / a.seg_rows, clip = row - seg * a.seg_rows;
    const float* src = a.in + (size_t)seg * a.in_stride + (size_t)clip * a.clip_stride;
    int nvalid = 0x7fffffff;
    if (a.clip_len > 0) {
        const int64_t left = (int64_t)a.clip_len - (int64_t)seg * a.in_stride;
        nvalid = left < 0 ? 0 : (left < a.nsample ? (int)left : a.nsample);
    }
    float2* out_row = reinterpret_cast<float2*>(a.out) + (size_t)row * G::F * a.nframe;
    for (int g = 0; g < a.gpc; ++g) {
        const int f_base = (chunk * a.gpc + g) * G::FR;
        if (f_base >= a.nframe) break;
        fill_stage<B, LMODE>(stage, src, f_base * G::HOP, a, tid, nvalid);
        __syncthreads();
        passA3_fwd<G>(stage, tb.win, tb.tw, zb, unit, fr);
        __syncthreads();
        const int t = f_base + fr;
#pragma unroll 1
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            float2 xa[G::RC], xb[G::RC], nyq;
            analysis_task3<G>(zb, tb, p, fr, xa, xb, nyq);
            store_task3<G>(out_row, a.nframe, t, p, xa, xb, nyq, a.edge_scale);
        }
        __syncthreads();        // pass C reads zb, the next pass A (after the next fill's barrier) writes it: the fill's barrier
                                // alone orders them only if every thread has left pass C -- keep the explicit one (cheap, 2 per group)
    }
}

template <class G, int EMODE>
__global__ void __launch_bounds__(G::NT, G::MINB) k_synthesis3(const SynArgs a) {
    using B = typename G::Base;
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* ostage = reinterpret_cast<float*>(se_smem + Smem<B>::ZB);
    float* hold = reinterpret_cast<float*>(se_smem + Smem<B>::ZB + Smem<B>::OSTAGE);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<B>(a.tb, se_smem + Smem<B>::ZB + Smem<B>::OSTAGE + (EMODE == EMIT_ADJ ? Smem<B>::HOLD : 0), tid);
    __syncthreads();
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const Chunk c = make_chunk<B>(chunk, a.nchunks, a.b_lo, a.b_hi);
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
#pragma unroll 1
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            float2 ya[G::RC], yb[G::RC], nyq;
            load_task3<G>(spec, a.nframe, t, p, ya, yb, nyq, a.edge_scale);
            synthesis_task3<G>(zb, tb, p, fr, ya, yb, nyq);
        }
        synthesis_tail3<G>(zb, tb, ostage, unit, fr, carry);
        if (EMODE == EMIT_ISTFT) emit_istft<B>(ostage, out_row, f_base, c, a, tid);
        else emit_adj<B>(ostage, hold, out_row, f_base, c, a.nsample, a.accumulate, 1.0f, tid);
        __syncthreads();        // the emitter reads ostage / the next group's pass C writes zb that pass A' just read
    }
    if (EMODE == EMIT_ADJ && c.last) finish_adj<B>(hold, out_row, a.nsample, a.accumulate, 1.0f, tid);
}

}  // namespace se
