// se_specloss.cuh -- STFT-domain training losses with the target spectrum never materialised
// (SURVEY.md 8f-2).  The reference computes `loss_function(enhanced, stft_custom(sources))` with
// torch's mse_loss / l1_loss on [B,C,F,T,2] (src/solver.py:457-458,480; src/distrib.py:263-267):
// here the target's STFT lives only in registers; the forward kernel reads the enhanced spectrum once
// and reduces, the backward kernel recomputes the target bins and writes d loss / d enhanced.
#pragma once
#include "se_kernels.cuh"

namespace se {

struct SpecLossArgs {
    Tables tb;
    const float* target;     // waveform rows [rows, N]
    const float* enh;        // enhanced spectrum [rows, F, T, 2]
    float* genh;             // bwd: [rows, F, T, 2]
    double* partials;        // fwd: one double per CTA
    const float* gout;       // bwd: device scalar
    int nsample, nframe, gpc, nchunks;
    int kind;                // 0: mse, 1: l1
    float inv_count;         // 1 / (global_rows * F * T * 2)
};

template <class G, bool BWD>
__global__ void __launch_bounds__(G::NT, G::MINB) k_spec_loss(const SpecLossArgs a) {
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* stage = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    __shared__ float red[32];
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.tb, se_smem + Smem<G>::ZB + Smem<G>::STAGE, tid);
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    AnaArgs la;
    la.tb = a.tb; la.nsample = a.nsample; la.nframe = a.nframe; la.in_len = a.nsample; la.pad = 0;
    const float2* erow = reinterpret_cast<const float2*>(a.enh) + (size_t)row * G::F * a.nframe;
    float2* grow = BWD ? reinterpret_cast<float2*>(a.genh) + (size_t)row * G::F * a.nframe : nullptr;
    const float gs = BWD ? __ldg(a.gout) * a.inv_count : 0.f;
    float acc = 0.f;
    for (int g = 0; g < a.gpc; ++g) {
        const int f_base = (chunk * a.gpc + g) * G::FR;
        if (f_base >= a.nframe) break;
        const int t = f_base + fr;
        const bool live = t < a.nframe;
        const int tc = live ? t : 0;
        fill_stage<G, LOAD_REFLECT>(stage, a.target + (size_t)row * a.nsample, f_base * G::HOP, la, tid);
        __syncthreads();
        analysis_passes<G>(stage, tb, zb, unit, fr);
        SE_TC_PRAGMA
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            const int qa = task_qa<G>(p), qb = task_qb<G>(p);
            float2 ea[8], eb[8], en;                      // enhanced bins first: 17 loads in flight during pass C
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                ea[k] = __ldg(erow + (size_t)(qa + G::S * k) * a.nframe + tc);
                eb[k] = __ldg(erow + (size_t)(qb + G::S * k) * a.nframe + tc);
            }
            en = __ldg(erow + (size_t)G::M * a.nframe + tc);
            float2 xa[8], xb[8], nyq;
            analysis_task<G>(zb, tb, p, fr, xa, xb, nyq);
            if (!live) continue;
#pragma unroll
            for (int k = 0; k < 17; ++k) {
                if (k == 16 && p != 0) continue;
                const int bin = k < 8 ? qa + G::S * k : (k < 16 ? qb + G::S * (k - 8) : G::M);
                float2 s = k < 8 ? xa[k] : (k < 16 ? xb[k - 8] : nyq);
                if (p == 0 && (k == 0 || k == 16)) s.y = 0.f;           // DC / Nyquist are real
                const float2 e = k < 8 ? ea[k] : (k < 16 ? eb[k - 8] : en);
                const float dx = e.x - s.x, dy = e.y - s.y;
                if (!BWD) {
                    acc += a.kind == 0 ? dx * dx + dy * dy : fabsf(dx) + fabsf(dy);
                } else {
                    float2 gv;
                    if (a.kind == 0) gv = make_float2(2.f * gs * dx, 2.f * gs * dy);
                    else gv = make_float2(dx > 0.f ? gs : (dx < 0.f ? -gs : 0.f), dy > 0.f ? gs : (dy < 0.f ? -gs : 0.f));
                    grow[(size_t)bin * a.nframe + t] = gv;
                }
            }
        }
        __syncthreads();
    }
    if (!BWD) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < G::NT / 32; ++w) tot += (double)red[w];
            a.partials[blockIdx.x] = tot;
        }
    }
}

// deterministic second stage: one CTA sums the per-CTA partials in a fixed order
static __global__ void k_sum_partials(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double sh[256];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    double acc = 0.0;
    for (int i = tid; i < n; i += 256) acc += partials[i];
    sh[tid] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    if (tid == 0) *out = sh[0];
}

}  // namespace se
