// se_conv.cuh -- DCCRN-convention synthesis (ConviSTFT.forward, src/model/dccrn.py:723-747) and
// its adjoint, computed with the FFT engine instead of a dense [2F x win_len] transposed conv.
//
// Closed form of the reference's pinv basis (K^T K = (n/2) I + parity blocks):
//   v[j]  = sum_k Re(Y_k e^{+2 pi i jk/n}),  j < win_len           (unnormalised C2R, H = Y / c_k)
//   v'[j] = v[j] - (sum of v over samples with the parity of j) / (n/2 + #samples of that parity)
//   frame = v' / (n/2) * w ;  y = OLA(frame) / (OLA(w^2) + 1e-8) ; drop win_len - hop in front.
// The analysis direction (ConvSTFT) is k_analysis<Geo<512,hop,256>, LOAD_ZEROPAD, planar>.
// hop does not divide n here, so overlap-add is a shared-memory gather over a ring of frames
// (16 current + 3 carried) rather than the lane rotation of se_fft.cuh.
#pragma once
#include "se_kernels.cuh"

namespace se {

struct ConvArgs {
    Tables tb;               // window placed at the FRONT of the frame, scale 0.5 * (2/n)
    const float* in;
    float* out;
    int win_len, nframe, out_len, pad;   // pad = win_len - hop
    int b_lo, b_hi, nchunks;             // synthesis chunking
    int gpc;                             // adjoint (analysis-style) chunking
    float inv_even, inv_odd;             // 1 / (n/2 + #even), 1 / (n/2 + #odd)
    // fused DCCRN tail (MODE >= 1): `in` / `spec` is the UNMASKED spectrum, the masks are two planes [rows,F,T]
    const float* spec;                   // bwd: unmasked spectrum (fwd reads it through `in`)
    const float* mre;
    const float* mim;
    float* gre;                          // bwd: gradient wrt the mask planes (replaces `out`)
    float* gim;
};

template <class G> struct ConvGeo {
    static_assert(G::FR == 16, "the per-frame parity reduction assumes 16 frames per half-warp");
    static constexpr int NOV = 4;                       // ceil(win_len / hop) supported: <= 4
    static constexpr int FROW = G::N + 2 + ((34 - (G::N + 2) % 32) % 32);   // frame row stride == 2 mod 32
    static constexpr int RING = G::FR + NOV - 1;
    static constexpr size_t FBUF = Smem<G>::al16(sizeof(float) * RING * FROW);
    static constexpr size_t RED = sizeof(float2) * (G::NT / 32) * G::FR;
    static constexpr size_t SYNTH = Smem<G>::ZB + FBUF + RED + Smem<G>::TABLES;
    static constexpr size_t ADJ = Smem<G>::ZB + Smem<G>::STAGE + RED + Smem<G>::TABLES;
};

// per-frame parity sums across the CTA: every thread contributes its partial (e, o) for frame fr
template <class G>
__device__ __forceinline__ float2 frame_parity_sums(float2 part, float2* red, int tid, int fr) {
    part.x += __shfl_xor_sync(0xffffffffu, part.x, 16);
    part.y += __shfl_xor_sync(0xffffffffu, part.y, 16);
    if ((tid & 31) < 16) red[(tid >> 5) * G::FR + fr] = part;
    __syncthreads();
    float2 tot = make_float2(0.f, 0.f);
#pragma unroll
    for (int w = 0; w < G::NT / 32; ++w) { const float2 v = red[w * G::FR + fr]; tot.x += v.x; tot.y += v.y; }
    return tot;
}

// MODE < 0: plain spectrum; MODE 1/2/3: the DCCRN mask tail (dccrn.py:203-221) applied to the bins as they are loaded
template <class G, int MODE>
__device__ __forceinline__ void load_task_planar(const float* __restrict__ row, const float* __restrict__ mre,
                                                 const float* __restrict__ mim, int T, int t, int p,
                                                 float2* ya, float2* yb, float2& nyq) {
    const bool ok = (t >= 0 && t < T);
    const int tc = ok ? t : 0;                                   // clamped: loads stay in bounds, result zeroed
    const int qa = task_qa<G>(p), qb = task_qb<G>(p);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        const size_t ia = (size_t)(qa + G::S * k4) * T + tc, ib = (size_t)(qb + G::S * k4) * T + tc;
        const size_t im = (size_t)G::F * T;
        ya[k4] = make_float2(__ldg(row + ia), __ldg(row + im + ia));
        yb[k4] = make_float2(__ldg(row + ib), __ldg(row + im + ib));
        if (MODE >= 0) {
            ya[k4] = MaskMath::apply<(MODE >= 0 ? MODE : 2), false>(ya[k4], make_float2(__ldg(mre + ia), __ldg(mim + ia)));
            yb[k4] = MaskMath::apply<(MODE >= 0 ? MODE : 2), false>(yb[k4], make_float2(__ldg(mre + ib), __ldg(mim + ib)));
        }
        if (!ok) ya[k4] = yb[k4] = make_float2(0.f, 0.f);
    }
    nyq = make_float2(0.f, 0.f);
    if (p == 0) {
        if (ok) {
            const size_t in = (size_t)G::M * T + tc;
            if (MODE >= 0) nyq = MaskMath::apply<(MODE >= 0 ? MODE : 2), false>(
                               make_float2(__ldg(row + in), __ldg(row + (size_t)G::F * T + in)), make_float2(__ldg(mre + in), __ldg(mim + in)));
            else nyq = make_float2(__ldg(row + in), 0.f);
            nyq.y = 0.f;
        }
        ya[0].x *= 2.f;          // H = Y / c_k with the 1/2 folded into the window: edges weigh 2
        nyq.x *= 2.f;
    }
}

template <class G, int MODE = -1>
__global__ void __launch_bounds__(G::NT, G::MINB) k_conv_istft(const ConvArgs a) {
    using C = ConvGeo<G>;
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* fbuf = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    float2* red = reinterpret_cast<float2*>(se_smem + Smem<G>::ZB + C::FBUF);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.tb, se_smem + Smem<G>::ZB + C::FBUF + C::RED, tid);   // visible after the ring-zero barrier below
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const int nb = a.b_hi - a.b_lo;
    const int b0 = a.b_lo + (int)(((int64_t)chunk * nb) / a.nchunks);
    const int b1 = a.b_lo + (int)(((int64_t)(chunk + 1) * nb) / a.nchunks);
    const int f0 = b0 - (C::NOV - 1);
    const int ngroups = (b1 - f0 + G::FR - 1) / G::FR;
    const float* spec = a.in + (size_t)row * 2 * G::F * a.nframe;
    float* out_row = a.out + (size_t)row * a.out_len;
    const int total = a.win_len + G::HOP * (a.nframe - 1);
    // ring slots 0..NOV-2 carry the previous group's last frames; zero them for the first group
    for (int i = tid; i < (C::NOV - 1) * C::FROW; i += G::NT) fbuf[i] = 0.f;
    __syncthreads();
    for (int g = 0; g < ngroups; ++g) {
        const int f_base = f0 + g * G::FR;
        const int t = f_base + fr;
#pragma unroll
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            float2 ya[8], yb[8], nyq;
            const size_t moff = (size_t)row * G::F * a.nframe;
            load_task_planar<G, MODE>(spec, MODE >= 0 ? a.mre + moff : nullptr, MODE >= 0 ? a.mim + moff : nullptr, a.nframe, t, p, ya, yb, nyq);
            synthesis_task<G>(zb, tb, p, fr, ya, yb, nyq);
        }
        __syncthreads();
        passB_inv<G>(tb.tw, zb, unit, fr);
        __syncthreads();
        // pass A' without the window: raw v[j]; parity sums over j < win_len
        float2 v[G::TA][G::R1];
        float2 part = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < G::TA; ++i) {
            const int u = unit + i * G::NU;
#pragma unroll
            for (int k = 0; k < G::R1; ++k) v[i][k] = zb[zidx<G>(u + 64 * k) + fr];
#pragma unroll
            for (int k = 1; k < G::R1; ++k) v[i][k] = cmulc(v[i][k], tb.tw[u * k]);
            dftR<G::R1, true>(v[i]);
#pragma unroll
            for (int r = 0; r < G::R1; ++r) {
                const int j = 2 * (u + 64 * r);
                if (j < a.win_len) part.x += v[i][r].x;
                if (j + 1 < a.win_len) part.y += v[i][r].y;
            }
        }
        const float2 tot = frame_parity_sums<G>(part, red, tid, fr);
        const float ce = tot.x * a.inv_even, co = tot.y * a.inv_odd;
#pragma unroll
        for (int i = 0; i < G::TA; ++i) {
            const int u = unit + i * G::NU;
#pragma unroll
            for (int r = 0; r < G::R1; ++r) {
                const int j = 2 * (u + 64 * r);
                const float2 w = *reinterpret_cast<const float2*>(tb.win + j);   // zero beyond win_len
                *reinterpret_cast<float2*>(fbuf + (C::NOV - 1 + fr) * C::FROW + j) =
                    make_float2((v[i][r].x - ce) * w.x, (v[i][r].y - co) * w.y);
            }
        }
        __syncthreads();
        // gather overlap-add: block b (= frame index) is complete once frames b-3..b are in the ring
        for (int idx = tid; idx < G::FR * G::HOP; idx += G::NT) {
            const int blk = idx / G::HOP, o = idx - blk * G::HOP;
            const int b = f_base + blk;
            if (b < b0 || b >= b1) continue;
            const int i = b * G::HOP + o;
            const int s = i - a.pad;
            if (s < 0 || s >= a.out_len || i >= total) continue;
            float acc = 0.f, env = 0.f;
#pragma unroll
            for (int q = 0; q < C::NOV; ++q) {
                const int j = o + q * G::HOP;
                const int tq = b - q;
                if (j < a.win_len && tq >= 0 && tq < a.nframe) {
                    acc += fbuf[(C::NOV - 1 + blk - q) * C::FROW + j];
                    env += __ldg(a.tb.w2 + j);
                }
            }
            out_row[s] = acc / (env + 1e-8f);
        }
        __syncthreads();
        // carry the last NOV-1 frames to the front of the ring
        for (int i = tid; i < (C::NOV - 1) * C::FROW; i += G::NT) fbuf[i] = fbuf[G::FR * C::FROW + i];
        __syncthreads();
    }
}

// adjoint of k_conv_istft: gy [rows,out_len] -> gspec [rows,2F,T] planar; MODE >= 1: continue through the mask tail
// to the gradient wrt the two mask planes (gspec is never written)
template <class G, int MODE = -1>
__global__ void __launch_bounds__(G::NT, G::MINB) k_conv_istft_adj(const ConvArgs a) {
    using C = ConvGeo<G>;
    SE_SMEM_DECL;
    float2* zb = reinterpret_cast<float2*>(se_smem);
    float* stage = reinterpret_cast<float*>(se_smem + Smem<G>::ZB);
    float2* red = reinterpret_cast<float2*>(se_smem + Smem<G>::ZB + Smem<G>::STAGE);
    const int tid = threadIdx.x, fr = tid % G::FR, unit = tid / G::FR;
    pdl_launch_dependents();
    const Tables tb = stage_tables<G>(a.tb, se_smem + Smem<G>::ZB + Smem<G>::STAGE + C::RED, tid);
    pdl_wait();
    const int row = blockIdx.x / a.nchunks, chunk = blockIdx.x - row * a.nchunks;
    const float* src = a.in + (size_t)row * a.out_len;
    float* out_row = a.out + (size_t)row * 2 * G::F * a.nframe;
    const int total = a.win_len + G::HOP * (a.nframe - 1);
    for (int g = 0; g < a.gpc; ++g) {
        const int f_base = (chunk * a.gpc + g) * G::FR;
        if (f_base >= a.nframe) break;
        // stage = gy / (env + 1e-8) in padded coordinates
        for (int rel = tid; rel < G::SROWS * G::HOP; rel += G::NT) {
            const int i = f_base * G::HOP + rel;
            const int s = i - a.pad;
            float val = 0.f;
            if (s >= 0 && s < a.out_len && i < total) {
                const int b = i / G::HOP, o = i - b * G::HOP;
                float env = 0.f;
#pragma unroll
                for (int q = 0; q < C::NOV; ++q) {
                    const int j = o + q * G::HOP, tq = b - q;
                    if (j < a.win_len && tq >= 0 && tq < a.nframe) env += __ldg(a.tb.w2 + j);
                }
                val = __ldg(src + s) / (env + 1e-8f);
            }
            stage[(rel / G::HOP) * G::SROW + rel % G::HOP] = val;
        }
        __syncthreads();
        // windowed frame, parity correction (I - P is symmetric), then the forward passes
        float2 v[G::TA][G::R1];
        float2 part = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < G::TA; ++i) {
            const int u = unit + i * G::NU;
#pragma unroll
            for (int r = 0; r < G::R1; ++r) {
                const int j = 2 * (u + 64 * r);
                const float2 x = *reinterpret_cast<const float2*>(stage + (fr + j / G::HOP) * G::SROW + j % G::HOP);
                const float2 w = *reinterpret_cast<const float2*>(tb.win + j);
                v[i][r] = make_float2(x.x * w.x, x.y * w.y);
                part.x += v[i][r].x;          // window is zero beyond win_len
                part.y += v[i][r].y;
            }
        }
        const float2 tot = frame_parity_sums<G>(part, red, tid, fr);
        const float ce = tot.x * a.inv_even, co = tot.y * a.inv_odd;
#pragma unroll
        for (int i = 0; i < G::TA; ++i) {
            const int u = unit + i * G::NU;
#pragma unroll
            for (int r = 0; r < G::R1; ++r) {
                const int j = 2 * (u + 64 * r);
                if (j < a.win_len) v[i][r].x -= ce;
                if (j + 1 < a.win_len) v[i][r].y -= co;
            }
            dftR<G::R1, false>(v[i]);
#pragma unroll
            for (int k = 1; k < G::R1; ++k) v[i][k] = cmul(v[i][k], tb.tw[u * k]);
#pragma unroll
            for (int k = 0; k < G::R1; ++k) zb[zidx<G>(u + 64 * k) + fr] = v[i][k];
        }
        __syncthreads();
        passB_fwd<G>(tb.tw, zb, unit, fr);
        __syncthreads();
        const int t = f_base + fr;
#pragma unroll
        for (int i = 0; i < G::TC; ++i) {
            const int p = unit + i * G::NU;
            float2 xa[8], xb[8], nyq;
            analysis_task<G>(zb, tb, p, fr, xa, xb, nyq);
            if (MODE < 0) {
                store_task_planar<G>(out_row, a.nframe, t, p, xa, xb, nyq);
            } else if (t < a.nframe) {
                constexpr int MM = MODE >= 0 ? MODE : 2;
                const size_t moff = (size_t)row * G::F * a.nframe, im = (size_t)G::F * a.nframe;
                const float* x_row = a.spec + 2 * moff;
                const int qa = task_qa<G>(p), qb = task_qb<G>(p);
#pragma unroll
                for (int k = 0; k < 17; ++k) {
                    if (k == 16 && p != 0) continue;
                    const int bin = k < 8 ? qa + G::S * k : (k < 16 ? qb + G::S * (k - 8) : G::M);
                    float2 gy = k < 8 ? xa[k] : (k < 16 ? xb[k - 8] : nyq);
                    if (bin == 0 || bin == G::M) gy.y = 0.f;              // imaginary grads at DC / Nyquist are exactly 0
                    const size_t idx = (size_t)bin * a.nframe + t;
                    float2 gm, gx;
                    MaskMath::grad<MM, false>(make_float2(__ldg(x_row + idx), __ldg(x_row + im + idx)),
                                              make_float2(__ldg(a.mre + moff + idx), __ldg(a.mim + moff + idx)), gy, gm, gx);
                    a.gre[moff + idx] = gm.x;
                    a.gim[moff + idx] = gm.y;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace se
