// se_fused.cuh -- wave -> STFT -> mask -> iSTFT -> wave in ONE kernel (and its backward to the
// raw mask).  The spectrum never touches HBM: 2S + M bytes per row instead of 2S + 4P + M
// (SURVEY.md 8d), where P ~ 4S.  Same numerics as stft_custom -> model tail -> istft_custom:
// the analysis window carries the reference's 1/win_length (the polar mask is not scale
// invariant because of its 1e-8), the synthesis window carries win_length / n.
#pragma once
#include "se_kernels.cuh"

namespace se {

struct EnhArgs {
    Tables ta;               // analysis tables: window * 0.5 / win_length
    Tables ts;               // synthesis tables: window * win_length / n   (env tables valid here)
    const float* x;          // [rows, N]
    const float* mask;       // [rows, F, T] (REAL) or [rows, F, T, 2]
    const float* gy;         // bwd: [rows, N]
    float* out;              // fwd: y [rows, N]; bwd: gmask
    int nsample, nframe;
    int b_lo, b_hi, nchunks; // fwd (synthesis-style chunking)
    int gpc;                 // bwd (analysis-style chunking)
    int mode, pre_tanh;
};

template <class G, int MODE>
__device__ __forceinline__ float2 load_mask(const EnhArgs& a, size_t row, int k, int t) {
    const size_t idx = (row * G::F + k) * (size_t)a.nframe + t;
    if (MODE == 0) return make_float2(__ldg(a.mask + idx), 0.f);
    return __ldg(reinterpret_cast<const float2*>(a.mask) + idx);
}

// Raw-mask value of one bin through a pointer that walks a unit's 8 bins (S*T apart): real masks are float
// rows, the complex modes float2 rows (see store_task_ft2 for why pointers are walked instead of re-indexed).
template <int MODE>
struct MaskWalk {
    const float* p;
    size_t step;
    __device__ __forceinline__ MaskWalk(const float* mask, size_t idx, size_t step_) : p(mask + (MODE == 0 ? idx : 2 * idx)), step(MODE == 0 ? step_ : 2 * step_) {}
    __device__ __forceinline__ float2 next() {
        const float2 v = MODE == 0 ? make_float2(__ldg(p), 0.f) : __ldg(reinterpret_cast<const float2*>(p));
        p += step;
        return v;
    }
};

template <class G, int MODE, bool TANH>
__global__ void __launch_bounds__(G::NT, G::MINB) k_enhance_fwd(const EnhArgs a) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* iobuf = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables ta = stage_tables<G>(a.ta, se_smem + Smem<G>::ZB + Smem<G>::IOBUF, tid);
    const Tables ts = stage_window<G>(ta, a.ts, se_smem + Smem<G>::ZB + Smem<G>::IOBUF + Smem<G>::TABLES, tid);
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const Chunk c = make_chunk<G>(chunk, a.nchunks, a.b_lo, a.b_hi);
    AnaArgs la;
    la.tb = a.ta; la.nsample = a.nsample; la.nframe = a.nframe; la.in_len = a.nsample; la.pad = 0;
    SynArgs sa;
    sa.tb = a.ts; sa.nsample = G::N + G::HOP * (a.nframe - 1); sa.out_len = a.nsample; sa.nframe = a.nframe;
    float* out_row = a.out + (size_t)row * a.nsample;
    float2 carry[G::TA][G::SEG];
#pragma unroll
    for (int i = 0; i < G::TA; ++i)
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) carry[i][s] = make_float2(0.f, 0.f);
    for (int g = 0; g < c.ngroups; ++g) {
        const int f_base = c.f0 + g * G::FR;
        const int t = f_base + fr;
        const bool live = (t >= 0 && t < a.nframe);
        const int tc = live ? t : 0;                       // clamped: loads stay in bounds, result zeroed
        fill_stage<G, LOAD_REFLECT>(iobuf, a.x + (size_t)row * a.nsample, f_base * G::HOP, la, tid);
        __syncthreads();
        analysis_passes<G>(iobuf, ta, zb, unit, fr);
#pragma unroll
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            const int qa = task_qa<G>(p), qb = task_qb<G>(p);
            // mask values first: 17 independent loads in flight while pass C runs
            float2 ma[8], mb[8], mn;
            {
                const size_t rb = (size_t)row * G::F * a.nframe + tc, step = (size_t)G::S * a.nframe;
                MaskWalk<MODE> wa(a.mask, rb + (size_t)qa * a.nframe, step), wb(a.mask, rb + (size_t)qb * a.nframe, step);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    ma[k] = wa.next();
                    mb[k] = wb.next();
                }
            }
            mn = load_mask<G, MODE>(a, row, G::M, tc);
            float2 xa[8], xb[8], nyq;
            analysis_task<G>(zb, ta, p, fr, xa, xb, nyq);
            const float keep = live ? 1.f : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                xa[k] = MaskMath::apply<MODE, TANH>(xa[k], ma[k]);
                xb[k] = MaskMath::apply<MODE, TANH>(xb[k], mb[k]);
                xa[k] = make_float2(xa[k].x * keep, xa[k].y * keep);
                xb[k] = make_float2(xb[k].x * keep, xb[k].y * keep);
            }
            nyq = MaskMath::apply<MODE, TANH>(nyq, mn);
            nyq = (p == 0) ? make_float2(nyq.x * keep, nyq.y * keep) : make_float2(0.f, 0.f);
            synthesis_task<G>(zb, ts, p, fr, xa, xb, nyq);
        }
        synthesis_tail<G>(zb, ts, iobuf, unit, fr, carry);
        emit_istft<G>(iobuf, out_row, f_base, c, sa, tid);
        __syncthreads();
    }
}

// backward to the raw mask: X = STFT(x) recomputed, gY = iSTFT^T(gy), then the mask adjoint.
// Two working buffers (x and gy transforms are both needed per bin) -> n_fft <= 1024 only.
template <class G, int MODE, bool TANH>
__global__ void __launch_bounds__(G::NT) k_enhance_bwd(const EnhArgs a) {
    SE_SMEM_DECL;
    float2* zbx = reinterpret_cast<float2*>(se_smem);
    float2* zbg = reinterpret_cast<float2*>(se_smem + Smem<G>::ZB);
    float* stage = reinterpret_cast<float*>(se_smem + 2 * Smem<G>::ZB);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables ta = stage_tables<G>(a.ta, se_smem + 2 * Smem<G>::ZB + Smem<G>::STAGE, tid);
    const Tables ts = stage_window<G>(ta, a.ts, se_smem + 2 * Smem<G>::ZB + Smem<G>::STAGE + Smem<G>::TABLES, tid);
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    AnaArgs lx;
    lx.tb = a.ta; lx.nsample = a.nsample; lx.nframe = a.nframe; lx.in_len = a.nsample; lx.pad = 0;
    AnaArgs lg;
    lg.tb = a.ts; lg.nsample = G::N + G::HOP * (a.nframe - 1); lg.nframe = a.nframe; lg.in_len = a.nsample; lg.pad = 0;
    for (int g = 0; g < a.gpc; ++g) {
        const int f_base = (chunk * a.gpc + g) * G::FR;
        if (f_base >= a.nframe) break;
        const int t = f_base + fr;
        const bool live = t < a.nframe;
        const int tc = live ? t : 0;
        fill_stage<G, LOAD_REFLECT>(stage, a.x + (size_t)row * a.nsample, f_base * G::HOP, lx, tid);
        __syncthreads();
        analysis_passes<G>(stage, ta, zbx, unit, fr);          // ends with a barrier: stage is free
        fill_stage<G, LOAD_ENV>(stage, a.gy + (size_t)row * a.nsample, f_base * G::HOP, lg, tid);
        __syncthreads();
        analysis_passes<G>(stage, ts, zbg, unit, fr);
#pragma unroll
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            const int qa = task_qa<G>(p), qb = task_qb<G>(p);
#pragma unroll
            for (int half = 0; half < 2; ++half) {               // unit a, then unit b: 8 bins each (+ Nyquist)
                const int q = half ? qb : qa;
                float2 m[8], mn = make_float2(0.f, 0.f);
                {
                    MaskWalk<MODE> w(a.mask, (size_t)row * G::F * a.nframe + (size_t)q * a.nframe + tc, (size_t)G::S * a.nframe);
#pragma unroll
                    for (int k = 0; k < 8; ++k) m[k] = w.next();
                }
                if (half == 0) mn = load_mask<G, MODE>(a, row, G::M, tc);
                float2 xa[8], xb[8], xn, ga[8], gb[8], gn;
                analysis_task<G>(zbx, ta, p, fr, xa, xb, xn);
                analysis_task<G>(zbg, ts, p, fr, ga, gb, gn);
                if (!live) continue;
                const size_t obase = (size_t)row * G::F * a.nframe + (size_t)q * a.nframe + t, ostep = (size_t)G::S * a.nframe;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (k == 8 && (half != 0 || p != 0)) continue;
                    const float2 x = k < 8 ? (half ? xb[k] : xa[k]) : xn;
                    float2 gy = k < 8 ? (half ? gb[k] : ga[k]) : gn;
                    // iSTFT adjoint: c_k / n with the 2/n folded into the window -> edges get 1/2, real only
                    if (p == 0 && half == 0 && (k == 0 || k == 8)) gy = make_float2(0.5f * gy.x, 0.f);
                    const size_t idx = k < 8 ? obase + (size_t)k * ostep : (size_t)row * G::F * a.nframe + (size_t)G::M * a.nframe + t;
                    float2 gm, gx;
                    MaskMath::grad<MODE, TANH>(x, k < 8 ? m[k] : mn, gy, gm, gx);
                    if (MODE == 0) a.out[idx] = gm.x;
                    else reinterpret_cast<float2*>(a.out)[idx] = gm;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace se

// ------------------------------------------------------------------------------------------------------
// Model tail + iSTFT in one kernel: istft_custom(apply_mask(spec, mask)) -- what the reference does at the end
// of every STFT model's forward followed by src/evaluate.py:72 / src/model/dccrn.py:223-224 -- without writing
// the masked spectrum, and its backward (gy -> d/d raw mask) without writing d/d masked spectrum.  One working
// buffer, so two CTAs per SM like the plain transforms (the fully fused se_enhance_bwd needs two).
namespace se {

struct MaskSynArgs {
    Tables ts;               // window * win_length / n (env tables valid)
    const float* spec;       // [rows, F, T, 2]
    const float* mask;       // [rows, F, T] or [rows, F, T, 2]
    const float* gy;         // bwd: [rows, length]
    float* out;              // fwd: y [rows, length]; bwd: gmask
    int nframe, length, natural;
    int b_lo, b_hi, nchunks; // fwd
    int gpc;                 // bwd
};

template <class G, int MODE>
__device__ __forceinline__ float2 load_mask_at(const float* __restrict__ mask, size_t idx) {
    if (MODE == 0) return make_float2(__ldg(mask + idx), 0.f);
    return __ldg(reinterpret_cast<const float2*>(mask) + idx);
}

template <class G, int MODE, bool TANH>
__global__ void __launch_bounds__(G::NT, G::MINB) k_mask_istft_fwd(const MaskSynArgs a) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* ostage = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.ts, se_smem + Smem<G>::ZB + Smem<G>::OSTAGE, tid);
    __syncthreads();
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const Chunk c = make_chunk<G>(chunk, a.nchunks, a.b_lo, a.b_hi);
    SynArgs sa;
    sa.tb = a.ts; sa.nsample = a.natural; sa.out_len = a.length; sa.nframe = a.nframe;
    const size_t rbase = (size_t)row * G::F * a.nframe;
    const float2* spec = reinterpret_cast<const float2*>(a.spec) + rbase;
    float* out_row = a.out + (size_t)row * a.length;
    float2 carry[G::TA][G::SEG];
#pragma unroll
    for (int i = 0; i < G::TA; ++i)
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) carry[i][s] = make_float2(0.f, 0.f);
    for (int g = 0; g < c.ngroups; ++g) {
        const int f_base = c.f0 + g * G::FR;
        const int t = f_base + fr;
        const bool live = (t >= 0 && t < a.nframe);
        const int tc = live ? t : 0;                       // clamped: loads stay in bounds, result zeroed
        SE_TC_PRAGMA
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            const int qa = task_qa<G>(p), qb = task_qb<G>(p);
            float2 ya[8], yb[8], nyq;
            {
                const size_t step = (size_t)G::S * a.nframe, ia = (size_t)qa * a.nframe + tc, ib = (size_t)qb * a.nframe + tc;
                const float2* sa = spec + ia;
                const float2* sb = spec + ib;
                MaskWalk<MODE> wa(a.mask, rbase + ia, step), wb(a.mask, rbase + ib, step);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    ya[k] = MaskMath::apply<MODE, TANH>(__ldg(sa), wa.next());
                    yb[k] = MaskMath::apply<MODE, TANH>(__ldg(sb), wb.next());
                    if (!live) ya[k] = yb[k] = make_float2(0.f, 0.f);
                    sa += step;
                    sb += step;
                }
            }
            nyq = make_float2(0.f, 0.f);
            if (p == 0 && live) {
                const size_t in = (size_t)G::M * a.nframe + tc;
                nyq = MaskMath::apply<MODE, TANH>(__ldg(spec + in), load_mask_at<G, MODE>(a.mask, rbase + in));
            }
            synthesis_task<G>(zb, tb, p, fr, ya, yb, nyq);
        }
        synthesis_tail<G>(zb, tb, ostage, unit, fr, carry);
        emit_istft<G>(ostage, out_row, f_base, c, sa, tid);
    }
}

template <class G, int MODE, bool TANH>
__device__ __forceinline__ void mask_grad_half(const MaskSynArgs& a, const float2* x, const float2* m, const float2* gy,
                                               size_t rbase, int q, int t, bool edge) {
    size_t idx = rbase + (size_t)q * a.nframe + t;
    const size_t step = (size_t)G::S * a.nframe;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float2 g = gy[k];
        // iSTFT adjoint: c_k / n with the 2/n folded into the window -> DC gets 1/2, real part only
        if (k == 0 && edge) g = make_float2(0.5f * g.x, 0.f);
        float2 gm, gx;
        MaskMath::grad<MODE, TANH>(x[k], m[k], g, gm, gx);
        if (MODE == 0) a.out[idx] = gm.x;
        else reinterpret_cast<float2*>(a.out)[idx] = gm;
        idx += step;
    }
}

template <class G, int MODE, bool TANH>
__global__ void __launch_bounds__(G::NT, G::MINB) k_mask_istft_bwd(const MaskSynArgs a) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* stage = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.ts, se_smem + Smem<G>::ZB + Smem<G>::STAGE, tid);
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    AnaArgs lg;
    lg.tb = a.ts; lg.nsample = a.natural; lg.nframe = a.nframe; lg.in_len = a.length; lg.pad = 0;
    const size_t rbase = (size_t)row * G::F * a.nframe;
    const float2* spec = reinterpret_cast<const float2*>(a.spec) + rbase;
    for (int g = 0; g < a.gpc; ++g) {
        const int f_base = (chunk * a.gpc + g) * G::FR;
        if (f_base >= a.nframe) break;
        const int t = f_base + fr;
        const bool live = t < a.nframe;
        const int tc = live ? t : 0;                       // clamped: loads stay in bounds, nothing stored
        fill_stage<G, LOAD_ENV>(stage, a.gy + (size_t)row * a.length, f_base * G::HOP, lg, tid);
        __syncthreads();
        analysis_passes<G>(stage, tb, zb, unit, fr);
        SE_TC_PRAGMA
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            const int qa = task_qa<G>(p), qb = task_qb<G>(p);
            float2 x[8], m[8];
            const size_t step = (size_t)G::S * a.nframe;
            {                                              // unit a's operands in flight while pass C runs
                const size_t ia = (size_t)qa * a.nframe + tc;
                const float2* sa = spec + ia;
                MaskWalk<MODE> wa(a.mask, rbase + ia, step);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    x[k] = __ldg(sa);
                    m[k] = wa.next();
                    sa += step;
                }
            }
            float2 ga[8], gb[8], gn;
            analysis_task<G>(zb, tb, p, fr, ga, gb, gn);
            if (live) mask_grad_half<G, MODE, TANH>(a, x, m, ga, rbase, qa, t, p == 0);
            {
                const size_t ib = (size_t)qb * a.nframe + tc;
                const float2* sb = spec + ib;
                MaskWalk<MODE> wb(a.mask, rbase + ib, step);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    x[k] = __ldg(sb);
                    m[k] = wb.next();
                    sb += step;
                }
            }
            if (live) mask_grad_half<G, MODE, TANH>(a, x, m, gb, rbase, qb, t, false);
            if (p == 0 && live) {                          // Nyquist
                const size_t idx = (size_t)G::M * a.nframe + t;
                float2 gm, gx;
                MaskMath::grad<MODE, TANH>(__ldg(spec + idx), load_mask_at<G, MODE>(a.mask, rbase + idx),
                                           make_float2(0.5f * gn.x, 0.f), gm, gx);
                if (MODE == 0) a.out[rbase + idx] = gm.x;
                else reinterpret_cast<float2*>(a.out)[rbase + idx] = gm;
            }
        }
        __syncthreads();
    }
}

}  // namespace se
