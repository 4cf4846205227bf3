// se_fft.cuh -- frame-interleaved batched real FFT engine for sm_100a.
//
// Layout idea (DESIGN.md "FFT engine"): a CTA works on FR = 16 consecutive STFT frames of one
// row at a time.  The n-point real transform is an M = n/2 point complex transform of
// z[m] = x[2m] + i x[2m+1] plus a split pass.  The M x 16 complex working set lives in shared
// memory as zb[point][frame]: the 16 frames are the fastest dimension, so a half-warp is always
// 16 frames of ONE butterfly:
//   * every shared-memory access is a contiguous 128-byte row  -> bank-conflict free by layout,
//   * twiddles are uniform across the half-warp                 -> broadcast loads,
//   * spectrum rows [F][T][2] are written/read as 16 consecutive t per bin -> 128-byte segments,
//   * overlap-add across frames is a lane rotation (__shfl) -- no atomics, no smem.
// M = R1 * 8 * 8 with R1 = 4 / 8 / 16 for n = 512 / 1024 / 2048: three in-place passes
// (DIF forward, DIT inverse), radix butterflies entirely in registers.
#pragma once
#include "se_platform.h"

namespace se {

// ------------------------------------------------------------------ complex helpers
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by the constant (c, -s) forward, (c, +s) inverse
template <bool INV>
__device__ __forceinline__ float2 ctw(float2 v, float c, float s) {
    if (INV) return make_float2(v.x * c - v.y * s, v.y * c + v.x * s);
    return make_float2(v.x * c + v.y * s, v.y * c - v.x * s);
}
// multiply by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 cmi(float2 v) {
    if (INV) return make_float2(-v.y, v.x);
    return make_float2(v.y, -v.x);
}

// ------------------------------------------------------------------ register butterflies
// All take natural-order input and produce natural-order output, X[k] = sum_m a[m] W^{mk},
// W = exp(-2 pi i / R) forward, conjugate for INV.
template <bool INV>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s0 = cadd(a0, a2), d0 = csub(a0, a2);
    const float2 s1 = cadd(a1, a3), d1 = cmi<INV>(csub(a1, a3));
    a0 = cadd(s0, s1);
    a2 = csub(s0, s1);
    a1 = cadd(d0, d1);
    a3 = csub(d0, d1);
}

template <bool INV>
__device__ __forceinline__ void dft8(float2* a) {
    constexpr float C = 0.70710678118654752440f;
    dft4<INV>(a[0], a[2], a[4], a[6]);          // E0..E3 in a0,a2,a4,a6
    dft4<INV>(a[1], a[3], a[5], a[7]);          // O0..O3 in a1,a3,a5,a7
    const float2 o0 = a[1];
    const float2 o1 = ctw<INV>(a[3], C, C);
    const float2 o2 = cmi<INV>(a[5]);
    const float2 o3 = ctw<INV>(a[7], -C, C);
    const float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}

template <bool INV>
__device__ __forceinline__ void dft16(float2* a) {
    constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f;   // cos/sin(pi/8)
    constexpr float C2 = 0.70710678118654752440f;
    // Y_b[k0] = dft4 over a of x[4a+b]; stored back in a[4*k0' + b] positions (k0 replaces a)
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4<INV>(a[b], a[4 + b], a[8 + b], a[12 + b]);
    // twiddle W16^{b*k0} on element (k0, b) = a[4*k0 + b]
    a[4 * 1 + 1] = ctw<INV>(a[4 * 1 + 1], C1, S1);      // m=1
    a[4 * 1 + 2] = ctw<INV>(a[4 * 1 + 2], C2, C2);      // m=2
    a[4 * 1 + 3] = ctw<INV>(a[4 * 1 + 3], S1, C1);      // m=3
    a[4 * 2 + 1] = ctw<INV>(a[4 * 2 + 1], C2, C2);      // m=2
    a[4 * 2 + 2] = cmi<INV>(a[4 * 2 + 2]);              // m=4
    a[4 * 2 + 3] = ctw<INV>(a[4 * 2 + 3], -C2, C2);     // m=6
    a[4 * 3 + 1] = ctw<INV>(a[4 * 3 + 1], S1, C1);      // m=3
    a[4 * 3 + 2] = ctw<INV>(a[4 * 3 + 2], -C2, C2);     // m=6
    a[4 * 3 + 3] = ctw<INV>(a[4 * 3 + 3], -C1, -S1);    // m=9: (cos(9pi/8), -sin(9pi/8)) = (-C1, +S1)
    // X[k0 + 4 k1] = dft4 over b of a[4*k0 + b]  -> lands in a[4*k0 + k1]; transpose to natural
#pragma unroll
    for (int k0 = 0; k0 < 4; ++k0) dft4<INV>(a[4 * k0], a[4 * k0 + 1], a[4 * k0 + 2], a[4 * k0 + 3]);
    // now a[4*k0 + k1] = X[k0 + 4*k1]: transpose the 4x4
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            const float2 t = a[4 * i + j];
            a[4 * i + j] = a[4 * j + i];
            a[4 * j + i] = t;
        }
}

template <int R, bool INV>
__device__ __forceinline__ void dftR(float2* a) {
    if (R == 4) dft4<INV>(a[0], a[1], a[2], a[3]);
    else if (R == 8) dft8<INV>(a);
    else dft16<INV>(a);
}

// ------------------------------------------------------------------ geometry
template <int N_, int HOP_, int NT_, int FR_ = 16, int MINB_ = 0>
struct Geo {
    static constexpr int N = N_, HOP = HOP_, NT = NT_;
    static constexpr int M = N / 2;            // complex points
    static constexpr int R1 = M / 64;          // radix of the time-side pass (4, 8, 16)
    static constexpr int FR = FR_;             // frames per group (= lanes per butterfly): 16, or 8 to halve the working set
    static constexpr int NU = NT / FR;         // butterfly units working in parallel per frame
    static constexpr int MINB = MINB_ > 0 ? MINB_ : (NT <= 128 ? 4 : (NT <= 256 ? 2 : 1));   // CTAs per SM the register budget is sized for
    static constexpr int TA = 64 / NU;         // pass-A tasks per thread   (64 radix-R1 butterflies)
    static constexpr int TB = 8 * R1 / NU;     // pass-B tasks per thread   (8*R1 radix-8 butterflies)
    static constexpr int TC = 4 * R1 / NU;     // pass-C paired tasks per thread (M/16 pairs)
    static constexpr int S = M / 8;            // bin stride between the 8 outputs of a pass-C unit
    static constexpr int F = M + 1;            // one-sided bins
    // waveform stage: hop-sized rows, padded so that the FR frames x 16/FR units of a half-warp (frame stride
    // = one row, unit stride = one float2) hit distinct banks: row stride == 2 * (16/FR) words (mod 32)
    static constexpr int PADW = (32 + 2 * (16 / FR) - HOP % 32) % 32;
    static constexpr int SROW = HOP + PADW;
    static constexpr int SPAN = N + (FR - 1) * HOP;
    static constexpr int SROWS = (SPAN + HOP - 1) / HOP;
    static constexpr int STAGE_FLOATS = SROWS * SROW;
    // FR = 8: a half-warp holds two butterflies; in pass C they sit 64 points apart (same banks), so every
    // 64-point block is shifted by FR float2 against its predecessor
    static constexpr int ZSKEW = FR < 16 ? FR : 0;
    static constexpr int ZB_FLOATS = 2 * (M * FR + (M / 64) * ZSKEW);
    // overlap-add (only when HOP divides N)
    static constexpr int OLA = N / HOP;
    static constexpr int SEG = (R1 / OLA) > 0 ? (R1 / OLA) : 1;   // float2 per hop segment per pass-A task
    static constexpr int OSTAGE_FLOATS = FR * SROW;
    static_assert(NU % 8 == 0, "pass B keeps its twiddles per thread: NU must be a multiple of 8");
    static_assert(TA >= 1 && TB >= 1 && TC >= 1, "too many threads for this size");
    static_assert(HOP % 2 == 0, "hop must be even (float2 staging)");
    static_assert(FR == 16 || FR == 8, "frames per group");
};

// float2 index of point p (frame 0) in the working buffer
template <class G>
__device__ __forceinline__ int zidx(int p) { return p * G::FR + (p >> 6) * G::ZSKEW; }

// position of pass-C unit q (bins q + S*k4) in zb
template <class G>
__device__ __forceinline__ int unit_base(int q) { return 64 * (q % G::R1) + 8 * (q / G::R1); }

// ------------------------------------------------------------------ forward passes (DIF)
// Pass A: reads the staged (padded-coordinate) waveform, applies the window, radix-R1.
template <class G>
__device__ __forceinline__ void passA_fwd(const float* __restrict__ stage, const float* __restrict__ win,
                                          const float2* __restrict__ tw, float2* __restrict__ zb,
                                          int unit, int fr) {
#pragma unroll
    for (int i = 0; i < G::TA; ++i) {
        const int u = unit + i * G::NU;
        float2 a[G::R1];
#pragma unroll
        for (int r = 0; r < G::R1; ++r) {
            const int j = 2 * (u + 64 * r);
            const float2 x = *reinterpret_cast<const float2*>(stage + (fr + j / G::HOP) * G::SROW + j % G::HOP);
            const float2 w = *reinterpret_cast<const float2*>(win + j);
            a[r] = make_float2(x.x * w.x, x.y * w.y);
        }
        dftR<G::R1, false>(a);
#pragma unroll
        for (int k = 1; k < G::R1; ++k) a[k] = cmul(a[k], tw[u * k]);
#pragma unroll
        for (int k = 0; k < G::R1; ++k) zb[zidx<G>(u + 64 * k) + fr] = a[k];
    }
}

template <class G>
__device__ __forceinline__ void passB_fwd(const float2* __restrict__ tw, float2* __restrict__ zb, int unit, int fr) {
    const int v = unit & 7;
    float2 t[8];
#pragma unroll
    for (int k = 1; k < 8; ++k) t[k] = tw[G::R1 * v * k];
#pragma unroll
    for (int i = 0; i < G::TB; ++i) {
        const int k1 = (unit >> 3) + i * (G::NU / 8);
        float2* p = zb + zidx<G>(64 * k1 + v) + fr;
        float2 a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = p[8 * r * G::FR];
        dft8<false>(a);
#pragma unroll
        for (int k = 1; k < 8; ++k) a[k] = cmul(a[k], t[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) p[8 * k * G::FR] = a[k];
    }
}

// Pass C for one unit: 8 contiguous points -> Z[q + S*k4], k4 = 0..7 (registers only)
template <class G>
__device__ __forceinline__ void passC_fwd_unit(const float2* __restrict__ zb, int q, int fr, float2* z) {
    const float2* p = zb + zidx<G>(unit_base<G>(q)) + fr;
#pragma unroll
    for (int r = 0; r < 8; ++r) z[r] = p[r * G::FR];
    dft8<false>(z);
}

// split step for the pair (k, M-k): in  zk = Z[k], zp = Z[M-k]  (Z pre-scaled by 1/2 via window)
//                                   out zk = X[k], zp = X[M-k];  w = exp(-2 pi i k / n)
__device__ __forceinline__ void split_pair(float2& zk, float2& zp, float2 w) {
    const float2 e = make_float2(zk.x + zp.x, zk.y - zp.y);     // Z[k] + conj Z[M-k]
    const float2 d = make_float2(zk.x - zp.x, zk.y + zp.y);     // Z[k] - conj Z[M-k]
    const float2 t = cmul(w, make_float2(d.y, -d.x));           // w * (-i d)
    zk = make_float2(e.x + t.x, e.y + t.y);
    zp = make_float2(e.x - t.x, t.y - e.y);                     // conj(e - t)
}
// inverse split: in yk = Y[k], yp = Y[M-k]; out zk = Z[k], zp = Z[M-k] (unnormalised C2R)
__device__ __forceinline__ void merge_pair(float2& yk, float2& yp, float2 w) {
    const float2 e = make_float2(yk.x + yp.x, yk.y - yp.y);
    const float2 d = make_float2(yk.x - yp.x, yk.y + yp.y);
    const float2 o = cmulc(d, w);                               // d * conj(w)
    yk = make_float2(e.x - o.y, e.y + o.x);                     // e + i o
    yp = make_float2(e.x + o.y, o.x - e.y);                     // conj(e) + i conj(o)
}

// Bins held by paired task p:  unit qa (bins qa + S*k4) in xa[], unit qb in xb[], Nyquist in nyq
// (p == 0 only).  p >= 1: qa = p, qb = S - p.  p == 0: qa = 0, qb = S/2 (both self-paired).
template <class G> __device__ __forceinline__ int task_qa(int p) { return p; }
template <class G> __device__ __forceinline__ int task_qb(int p) { return p == 0 ? G::S / 2 : G::S - p; }

template <class G>
__device__ __forceinline__ void split_task(int p, const float2* __restrict__ twn, float2* xa, float2* xb, float2& nyq) {
    if (p != 0) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) split_pair(xa[k4], xb[7 - k4], twn[p + G::S * k4]);
        nyq = make_float2(0.f, 0.f);
    } else {
        const float2 z0 = xa[0];
        xa[0] = make_float2(2.f * (z0.x + z0.y), 0.f);       // Z is pre-scaled by 1/2: E = 2 Re, O' = 2 Im
        nyq = make_float2(2.f * (z0.x - z0.y), 0.f);
#pragma unroll
        for (int k4 = 1; k4 < 4; ++k4) split_pair(xa[k4], xa[8 - k4], twn[G::S * k4]);
        float2 m0 = xa[4], m1 = xa[4];
        split_pair(m0, m1, twn[G::S * 4]);
        xa[4] = m0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) split_pair(xb[k4], xb[7 - k4], twn[G::S / 2 + G::S * k4]);
    }
}

template <class G>
__device__ __forceinline__ void merge_task(int p, const float2* __restrict__ twn, float2* ya, float2* yb, float2 nyq) {
    if (p != 0) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) merge_pair(ya[k4], yb[7 - k4], twn[p + G::S * k4]);
    } else {
        const float y0 = ya[0].x, ym = nyq.x;                   // imaginary parts of DC / Nyquist ignored
        ya[0] = make_float2(y0 + ym, y0 - ym);
#pragma unroll
        for (int k4 = 1; k4 < 4; ++k4) merge_pair(ya[k4], ya[8 - k4], twn[G::S * k4]);
        float2 m0 = ya[4], m1 = ya[4];
        merge_pair(m0, m1, twn[G::S * 4]);
        ya[4] = m0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) merge_pair(yb[k4], yb[7 - k4], twn[G::S / 2 + G::S * k4]);
    }
}

// ------------------------------------------------------------------ inverse passes (DIT)
template <class G>
__device__ __forceinline__ void passC_inv_unit(float2* __restrict__ zb, int q, int fr, float2* z) {
    dft8<true>(z);
    float2* p = zb + zidx<G>(unit_base<G>(q)) + fr;
#pragma unroll
    for (int r = 0; r < 8; ++r) p[r * G::FR] = z[r];
}

template <class G>
__device__ __forceinline__ void passB_inv(const float2* __restrict__ tw, float2* __restrict__ zb, int unit, int fr) {
    const int v = unit & 7;
    float2 t[8];
#pragma unroll
    for (int k = 1; k < 8; ++k) t[k] = tw[G::R1 * v * k];
#pragma unroll
    for (int i = 0; i < G::TB; ++i) {
        const int k1 = (unit >> 3) + i * (G::NU / 8);
        float2* p = zb + zidx<G>(64 * k1 + v) + fr;
        float2 a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = p[8 * r * G::FR];
#pragma unroll
        for (int k = 1; k < 8; ++k) a[k] = cmulc(a[k], t[k]);
        dft8<true>(a);
#pragma unroll
        for (int k = 0; k < 8; ++k) p[8 * k * G::FR] = a[k];
    }
}

// Pass A inverse for task i: returns the windowed time samples (y[2m], y[2m+1]), m = u + 64 r
template <class G>
__device__ __forceinline__ void passA_inv_task(const float2* __restrict__ zb, const float* __restrict__ win,
                                               const float2* __restrict__ tw, int u, int fr, float2* a) {
#pragma unroll
    for (int k = 0; k < G::R1; ++k) a[k] = zb[zidx<G>(u + 64 * k) + fr];
#pragma unroll
    for (int k = 1; k < G::R1; ++k) a[k] = cmulc(a[k], tw[u * k]);
    dftR<G::R1, true>(a);
#pragma unroll
    for (int r = 0; r < G::R1; ++r) {
        const float2 w = *reinterpret_cast<const float2*>(win + 2 * (u + 64 * r));
        a[r] = make_float2(a[r].x * w.x, a[r].y * w.y);
    }
}

// Overlap-add by lane rotation.  Thread (u, fr) holds frame fr's samples in OLA hop-segments of
// SEG float2 each; output block b = frame index gets segment q of frame b - q.  Contributions
// that wrap around the 16-lane group belong to the NEXT group's first OLA-1 blocks: they are
// returned in `carry` (registers) and added there.  acc[s] is block `fr`, offset 2*(u + 64 s).
template <class G, bool CARRY = true>
__device__ __forceinline__ void ola_rotate(const float2* a, int fr, float2* carry, float2* acc) {
    float2 nc[G::SEG];
#pragma unroll
    for (int s = 0; s < G::SEG; ++s) {
        acc[s] = CARRY ? cadd(a[s], carry[s]) : a[s];
        nc[s] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int q = 1; q < G::OLA; ++q) {
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) {
            const float rx = __shfl_sync(0xffffffffu, a[q * G::SEG + s].x, (fr - q) & (G::FR - 1), G::FR);
            const float ry = __shfl_sync(0xffffffffu, a[q * G::SEG + s].y, (fr - q) & (G::FR - 1), G::FR);
            if (fr >= q) { acc[s].x += rx; acc[s].y += ry; }
            else if (CARRY) { nc[s].x += rx; nc[s].y += ry; }
        }
    }
    if (CARRY) {
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) carry[s] = nc[s];
    }
}

}  // namespace se
