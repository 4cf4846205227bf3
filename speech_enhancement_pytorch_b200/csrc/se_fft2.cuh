// se_fft2.cuh -- "signal-pair" variant of the frame-interleaved FFT engine (se_fft.cuh) for sm_100a.
//
// Why (profiles/r02_ubench.jsonl, measured on B200): the scalar engine is issue-bound (ncu: 54-62 % issue
// active, ~1.3 thread-instructions per FFT flop, 70 % of them FADD/FMUL).  sm_100 has packed fp32
// instructions (FADD2 / FMUL2 / FFMA2: two fp32 lanes of a 64-bit register pair per issue slot, same lane
// throughput as the scalar forms) -- they halve the issue slots of the butterflies IF both halves of every
// pair run the same instruction stream.  A complex value as (re, im) does not (rotations and twiddle
// products swap halves), two independent SIGNALS do: every thread carries the same butterfly of two
// signals, c2 = {re = (re0, re1), im = (im0, im1)}.  The pair is whatever the caller makes it -- the
// reference and the estimate of the loss forward, two rows of a batch everywhere else.
//
// Layout: a CTA works on 8 consecutive frames x 2 signals of one row (pair) at a time.  The working set is
// zb[point][frame lane] of float4 = (re0, re1, im0, im1): one LDS.128 / STS.128 moves a complex point of
// both signals, and the 8 lanes of a butterfly unit always touch one contiguous 128-byte row -- 128-bit
// accesses are served per quarter-warp, so they are bank-conflict free for ANY row address (measured:
// 246 B/ns per SM for row strides 1..64).  Twiddle and window tables are staged in shared memory already
// duplicated, (c, c, s, s), so a broadcast LDS.128 delivers both packed operands.
// Pass structure (M = R1 * 8 * 8, DIF forward, DIT inverse, split in registers) is that of se_fft.cuh.
#pragma once
#include "se_fft.cuh"

namespace se {

#ifdef SE_EMULATE
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
#endif

// ------------------------------------------------------------------ packed helpers
__device__ __forceinline__ float2 p_neg(float2 a) { return make_float2(-a.x, -a.y); }      // folds into the operand modifier
__device__ __forceinline__ float2 p_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 p_sub(float2 a, float2 b) { return __fadd2_rn(a, p_neg(b)); }
__device__ __forceinline__ float2 p_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 p_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 p_dup(float v) { return make_float2(v, v); }

// one complex value of each of the two signals
struct c2 {
    float2 re, im;
};
__device__ __forceinline__ c2 mk2(float2 re, float2 im) { c2 r; r.re = re; r.im = im; return r; }
__device__ __forceinline__ c2 zero2() { return mk2(make_float2(0.f, 0.f), make_float2(0.f, 0.f)); }
__device__ __forceinline__ c2 from4(float4 v) { return mk2(make_float2(v.x, v.y), make_float2(v.z, v.w)); }
__device__ __forceinline__ float4 to4(c2 v) { return make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
__device__ __forceinline__ c2 cadd(c2 a, c2 b) { return mk2(p_add(a.re, b.re), p_add(a.im, b.im)); }
__device__ __forceinline__ c2 csub(c2 a, c2 b) { return mk2(p_sub(a.re, b.re), p_sub(a.im, b.im)); }
__device__ __forceinline__ c2 cscale(c2 a, float2 s) { return mk2(p_mul(a.re, s), p_mul(a.im, s)); }
// a * (c + i s), the factor given as broadcast pairs
__device__ __forceinline__ c2 cmul(c2 a, float2 c, float2 s) {
    return mk2(p_fma(p_neg(a.im), s, p_mul(a.re, c)), p_fma(a.re, s, p_mul(a.im, c)));
}
// a * (c - i s)
__device__ __forceinline__ c2 cmulc(c2 a, float2 c, float2 s) {
    return mk2(p_fma(a.im, s, p_mul(a.re, c)), p_fma(p_neg(a.re), s, p_mul(a.im, c)));
}
// multiply by the constant (c, -s) forward, (c, +s) inverse
template <bool INV>
__device__ __forceinline__ c2 ctw(c2 v, float c, float s) {
    return INV ? cmul(v, p_dup(c), p_dup(s)) : cmulc(v, p_dup(c), p_dup(s));
}
// multiply by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ c2 cmi(c2 v) {
    return INV ? mk2(p_neg(v.im), v.re) : mk2(v.im, p_neg(v.re));
}

// ------------------------------------------------------------------ register butterflies (see se_fft.cuh)
template <bool INV>
__device__ __forceinline__ void dft4(c2& a0, c2& a1, c2& a2, c2& a3) {
    const c2 s0 = cadd(a0, a2), d0 = csub(a0, a2);
    const c2 s1 = cadd(a1, a3), d1 = cmi<INV>(csub(a1, a3));
    a0 = cadd(s0, s1);
    a2 = csub(s0, s1);
    a1 = cadd(d0, d1);
    a3 = csub(d0, d1);
}

template <bool INV>
__device__ __forceinline__ void dft8(c2* a) {
    constexpr float C = 0.70710678118654752440f;
    dft4<INV>(a[0], a[2], a[4], a[6]);
    dft4<INV>(a[1], a[3], a[5], a[7]);
    const c2 o0 = a[1];
    const c2 o1 = ctw<INV>(a[3], C, C);
    const c2 o2 = cmi<INV>(a[5]);
    const c2 o3 = ctw<INV>(a[7], -C, C);
    const c2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}

template <bool INV>
__device__ __forceinline__ void dft16(c2* a) {
    constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f;
    constexpr float C2 = 0.70710678118654752440f;
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4<INV>(a[b], a[4 + b], a[8 + b], a[12 + b]);
    a[4 * 1 + 1] = ctw<INV>(a[4 * 1 + 1], C1, S1);
    a[4 * 1 + 2] = ctw<INV>(a[4 * 1 + 2], C2, C2);
    a[4 * 1 + 3] = ctw<INV>(a[4 * 1 + 3], S1, C1);
    a[4 * 2 + 1] = ctw<INV>(a[4 * 2 + 1], C2, C2);
    a[4 * 2 + 2] = cmi<INV>(a[4 * 2 + 2]);
    a[4 * 2 + 3] = ctw<INV>(a[4 * 2 + 3], -C2, C2);
    a[4 * 3 + 1] = ctw<INV>(a[4 * 3 + 1], S1, C1);
    a[4 * 3 + 2] = ctw<INV>(a[4 * 3 + 2], -C2, C2);
    a[4 * 3 + 3] = ctw<INV>(a[4 * 3 + 3], -C1, -S1);
#pragma unroll
    for (int k0 = 0; k0 < 4; ++k0) dft4<INV>(a[4 * k0], a[4 * k0 + 1], a[4 * k0 + 2], a[4 * k0 + 3]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            const c2 t = a[4 * i + j];
            a[4 * i + j] = a[4 * j + i];
            a[4 * j + i] = t;
        }
}

template <int R, bool INV>
__device__ __forceinline__ void dftR(c2* a) {
    if (R == 4) dft4<INV>(a[0], a[1], a[2], a[3]);
    else if (R == 8) dft8<INV>(a);
    else dft16<INV>(a);
}

// ------------------------------------------------------------------ geometry
// DUP: tables staged as (a, a, b, b) float4 (one broadcast LDS.128 per twiddle) instead of (a, b) float2 + two
// register moves; costs 12 N instead of 6 N bytes of shared memory, which the fused n = 2048 kernels do not have.
template <int N_, int HOP_, int NT_, bool DUP_ = true>
struct Geo2 {
    static constexpr int N = N_, HOP = HOP_, NT = NT_;
    static constexpr bool DUP = DUP_;
    static constexpr int M = N / 2;
    static constexpr int R1 = M / 64;
    static constexpr int FR = 8;               // frame lanes per butterfly unit (a quarter-warp)
    static constexpr int NU = NT / FR;
    static constexpr int MINB = NT <= 128 ? 4 : (NT <= 256 ? 2 : 1);
    static constexpr int TA = 64 / NU;
    static constexpr int TB = 8 * R1 / NU;
    static constexpr int TC = 4 * R1 / NU;
    static constexpr int S = M / 8;
    static constexpr int F = M + 1;
    // waveform stage: float2 (signal 0, signal 1) per sample, hop-sized rows; the 8 frame lanes of a unit read
    // 16 bytes each one row apart -> row pitch == 16 bytes (mod 128) makes the quarter-warp cover all banks once
    static constexpr int PAD = ((16 - (HOP * 8) % 128 + 128) % 128) / 8;
    static constexpr int SROW = HOP + PAD;     // float2 per row
    static constexpr int SPAN = N + (FR - 1) * HOP;
    static constexpr int SROWS = (SPAN + HOP - 1) / HOP;
    static constexpr int OLA = N / HOP;
    static constexpr int SEG = (R1 / OLA) > 0 ? (R1 / OLA) : 1;
    static constexpr size_t ZB_BYTES = (size_t)M * FR * 16;
    static constexpr size_t STAGE_BYTES = (size_t)SROWS * SROW * 8;
    static constexpr size_t OSTAGE_BYTES = (size_t)FR * SROW * 8;
    static constexpr size_t IOBUF_BYTES = STAGE_BYTES > OSTAGE_BYTES ? STAGE_BYTES : OSTAGE_BYTES;
    static constexpr size_t HOLD_BYTES = (size_t)(N + 2 * HOP) * 8;
    static constexpr size_t TABLE_BYTES = (size_t)M * (DUP ? 16 : 8);      // each of window / tw / twn
    static constexpr size_t TABLES_BYTES = 3 * TABLE_BYTES;
    static_assert(NU % 8 == 0, "pass B keeps v = unit & 7");
    static_assert(TA >= 1 && TB >= 1 && TC >= 1, "too many threads for this size");
    static_assert(HOP % 4 == 0 && SROW % 2 == 0, "128-bit staging");
    static_assert(OLA <= FR, "overlap-add rotates within the 8 frame lanes");
    static_assert(STAGE_BYTES % 16 == 0 && OSTAGE_BYTES % 16 == 0 && HOLD_BYTES % 16 == 0, "regions stay 16-byte aligned");
};

// table entry k as two broadcast pairs (a, a), (b, b)
template <class G>
__device__ __forceinline__ void pair_at(const void* __restrict__ tab, int k, float2& a, float2& b) {
    if (G::DUP) {
        const float4 t = reinterpret_cast<const float4*>(tab)[k];
        a = make_float2(t.x, t.y);
        b = make_float2(t.z, t.w);
    } else {
        const float2 t = reinterpret_cast<const float2*>(tab)[k];
        a = p_dup(t.x);
        b = p_dup(t.y);
    }
}

struct Tables2 {
    const void* win;        // (w[2m], w[2m+1]) per point m, scaled per op kind
    const void* tw;         // exp(-2 pi i k / M) as (cos, -sin)
    const void* twn;        // exp(-2 pi i k / N)
    const float* w2;        // global: unscaled window^2, N floats
    const float* inv_env;   // global: 1 / sum_q w2[o + q*HOP], HOP floats
};

template <class G>
__device__ __forceinline__ int unit_base2(int q) { return 64 * (q % G::R1) + 8 * (q / G::R1); }

// ------------------------------------------------------------------ forward passes (DIF)
// Pass A: staged waveform pair (padded coordinates) x window, radix-R1, twiddle, -> zb
template <class G>
__device__ __forceinline__ void passA_fwd2(const float2* __restrict__ stage, const Tables2& tb, float4* __restrict__ zb,
                                           int unit, int fr) {
#pragma unroll
    for (int i = 0; i < G::TA; ++i) {
        const int u = unit + i * G::NU;
        c2 a[G::R1];
#pragma unroll
        for (int r = 0; r < G::R1; ++r) {
            const int j = 2 * (u + 64 * r);
            const float4 x = *reinterpret_cast<const float4*>(stage + (fr + j / G::HOP) * G::SROW + j % G::HOP);
            float2 w0, w1;
            pair_at<G>(tb.win, u + 64 * r, w0, w1);
            a[r] = mk2(p_mul(make_float2(x.x, x.y), w0), p_mul(make_float2(x.z, x.w), w1));
        }
        dftR<G::R1, false>(a);
#pragma unroll
        for (int k = 1; k < G::R1; ++k) {
            float2 c, s;
            pair_at<G>(tb.tw, u * k, c, s);
            a[k] = cmul(a[k], c, s);
        }
#pragma unroll
        for (int k = 0; k < G::R1; ++k) zb[(u + 64 * k) * G::FR + fr] = to4(a[k]);
    }
}

template <class G, bool INV>
__device__ __forceinline__ void passB2(const Tables2& tb, float4* __restrict__ zb, int unit, int fr) {
    const int v = unit & 7;
#pragma unroll
    for (int i = 0; i < G::TB; ++i) {
        const int k1 = (unit >> 3) + i * (G::NU / 8);
        float4* p = zb + (64 * k1 + v) * G::FR + fr;
        c2 a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = from4(p[8 * r * G::FR]);
        if (INV) {
#pragma unroll
            for (int k = 1; k < 8; ++k) {
                float2 c, s;
                pair_at<G>(tb.tw, G::R1 * v * k, c, s);
                a[k] = cmulc(a[k], c, s);
            }
        }
        dft8<INV>(a);
        if (!INV) {
#pragma unroll
            for (int k = 1; k < 8; ++k) {
                float2 c, s;
                pair_at<G>(tb.tw, G::R1 * v * k, c, s);
                a[k] = cmul(a[k], c, s);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) p[8 * k * G::FR] = to4(a[k]);
    }
}

template <class G>
__device__ __forceinline__ void passC_fwd_unit2(const float4* __restrict__ zb, int q, int fr, c2* z) {
    const float4* p = zb + unit_base2<G>(q) * G::FR + fr;
#pragma unroll
    for (int r = 0; r < 8; ++r) z[r] = from4(p[r * G::FR]);
    dft8<false>(z);
}
template <class G>
__device__ __forceinline__ void passC_inv_unit2(float4* __restrict__ zb, int q, int fr, c2* z) {
    dft8<true>(z);
    float4* p = zb + unit_base2<G>(q) * G::FR + fr;
#pragma unroll
    for (int r = 0; r < 8; ++r) p[r * G::FR] = to4(z[r]);
}

// split / merge of the pair (k, M-k); (wc, ws) = exp(-2 pi i k / n) as broadcast pairs (see se_fft.cuh)
__device__ __forceinline__ void split_pair2(c2& zk, c2& zp, float2 wc, float2 ws) {
    const c2 e = mk2(p_add(zk.re, zp.re), p_sub(zk.im, zp.im));
    const c2 d = mk2(p_sub(zk.re, zp.re), p_add(zk.im, zp.im));
    // t = w * (-i d) = w * (d.im, -d.re)
    const float2 tr = p_fma(ws, d.re, p_mul(wc, d.im));
    const float2 ti = p_fma(ws, d.im, p_neg(p_mul(wc, d.re)));
    zk = mk2(p_add(e.re, tr), p_add(e.im, ti));
    zp = mk2(p_sub(e.re, tr), p_sub(ti, e.im));
}
__device__ __forceinline__ void merge_pair2(c2& yk, c2& yp, float2 wc, float2 ws) {
    const c2 e = mk2(p_add(yk.re, yp.re), p_sub(yk.im, yp.im));
    const c2 d = mk2(p_sub(yk.re, yp.re), p_add(yk.im, yp.im));
    // o = d * conj(w)
    const float2 orr = p_fma(d.im, ws, p_mul(d.re, wc));
    const float2 oi = p_fma(p_neg(d.re), ws, p_mul(d.im, wc));
    yk = mk2(p_sub(e.re, oi), p_add(e.im, orr));
    yp = mk2(p_add(e.re, oi), p_sub(orr, e.im));
}

template <class G> __device__ __forceinline__ int task_qa2(int p) { return p; }
template <class G> __device__ __forceinline__ int task_qb2(int p) { return p == 0 ? G::S / 2 : G::S - p; }

template <class G>
__device__ __forceinline__ void split_task2(int p, const Tables2& tb, c2* xa, c2* xb, c2& nyq) {
    float2 wc, ws;
    if (p != 0) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            pair_at<G>(tb.twn, p + G::S * k4, wc, ws);
            split_pair2(xa[k4], xb[7 - k4], wc, ws);
        }
        nyq = zero2();
    } else {
        const c2 z0 = xa[0];
        const float2 two = p_dup(2.f);
        xa[0] = mk2(p_mul(two, p_add(z0.re, z0.im)), make_float2(0.f, 0.f));
        nyq = mk2(p_mul(two, p_sub(z0.re, z0.im)), make_float2(0.f, 0.f));
#pragma unroll
        for (int k4 = 1; k4 < 4; ++k4) {
            pair_at<G>(tb.twn, G::S * k4, wc, ws);
            split_pair2(xa[k4], xa[8 - k4], wc, ws);
        }
        c2 m0 = xa[4], m1 = xa[4];
        pair_at<G>(tb.twn, G::S * 4, wc, ws);
        split_pair2(m0, m1, wc, ws);
        xa[4] = m0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            pair_at<G>(tb.twn, G::S / 2 + G::S * k4, wc, ws);
            split_pair2(xb[k4], xb[7 - k4], wc, ws);
        }
    }
}

template <class G>
__device__ __forceinline__ void merge_task2(int p, const Tables2& tb, c2* ya, c2* yb, c2 nyq) {
    float2 wc, ws;
    if (p != 0) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            pair_at<G>(tb.twn, p + G::S * k4, wc, ws);
            merge_pair2(ya[k4], yb[7 - k4], wc, ws);
        }
    } else {
        const float2 y0 = ya[0].re, ym = nyq.re;                 // imaginary parts of DC / Nyquist ignored
        ya[0] = mk2(p_add(y0, ym), p_sub(y0, ym));
#pragma unroll
        for (int k4 = 1; k4 < 4; ++k4) {
            pair_at<G>(tb.twn, G::S * k4, wc, ws);
            merge_pair2(ya[k4], ya[8 - k4], wc, ws);
        }
        c2 m0 = ya[4], m1 = ya[4];
        pair_at<G>(tb.twn, G::S * 4, wc, ws);
        merge_pair2(m0, m1, wc, ws);
        ya[4] = m0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            pair_at<G>(tb.twn, G::S / 2 + G::S * k4, wc, ws);
            merge_pair2(yb[k4], yb[7 - k4], wc, ws);
        }
    }
}

// Pass A inverse for task u: windowed time samples (y[2m], y[2m+1]) of both signals, m = u + 64 r
template <class G>
__device__ __forceinline__ void passA_inv_task2(const float4* __restrict__ zb, const Tables2& tb, int u, int fr, c2* a) {
#pragma unroll
    for (int k = 0; k < G::R1; ++k) a[k] = from4(zb[(u + 64 * k) * G::FR + fr]);
#pragma unroll
    for (int k = 1; k < G::R1; ++k) {
        float2 c, s;
        pair_at<G>(tb.tw, u * k, c, s);
        a[k] = cmulc(a[k], c, s);
    }
    dftR<G::R1, true>(a);
#pragma unroll
    for (int r = 0; r < G::R1; ++r) {
        float2 w0, w1;
        pair_at<G>(tb.win, u + 64 * r, w0, w1);
        a[r] = mk2(p_mul(a[r].re, w0), p_mul(a[r].im, w1));
    }
}

// Overlap-add by lane rotation within the 8 frame lanes (see se_fft.cuh ola_rotate)
template <class G, bool CARRY = true>
__device__ __forceinline__ void ola_rotate2(const c2* a, int fr, c2* carry, c2* acc) {
    c2 nc[G::SEG];
#pragma unroll
    for (int s = 0; s < G::SEG; ++s) {
        acc[s] = CARRY ? cadd(a[s], carry[s]) : a[s];
        nc[s] = zero2();
    }
#pragma unroll
    for (int q = 1; q < G::OLA; ++q) {
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) {
            const c2 v = a[q * G::SEG + s];
            const int src = (fr - q) & (G::FR - 1);
            c2 r;
            r.re.x = __shfl_sync(0xffffffffu, v.re.x, src, G::FR);
            r.re.y = __shfl_sync(0xffffffffu, v.re.y, src, G::FR);
            r.im.x = __shfl_sync(0xffffffffu, v.im.x, src, G::FR);
            r.im.y = __shfl_sync(0xffffffffu, v.im.y, src, G::FR);
            if (fr >= q) acc[s] = cadd(acc[s], r);
            else if (CARRY) nc[s] = cadd(nc[s], r);
        }
    }
    if (CARRY) {
#pragma unroll
        for (int s = 0; s < G::SEG; ++s) carry[s] = nc[s];
    }
}

}  // namespace se
