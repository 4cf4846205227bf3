// se_fft3.cuh -- two-pass variant of the frame-interleaved FFT engine (se_fft.cuh) for n_fft 512 and 1024.
//
// Round 2's measurements (profiles/r02_notes.md) say what bounds the three-pass engine: shared-memory
// wavefronts (about 229 per 1024-point frame, 160 of them the working set crossing shared memory three times) together
// with the issue slots of ~33 k thread instructions, at ~50 % overlap -- not issue width alone (packed fp32 did not
// help), not occupancy, not fill latency.  What is left is the NUMBER of passes: M = R1 x 16 with R1 = 16 (n = 512) or
// 32 (n = 1024) needs ONE intermediate round trip:
//
//   pass A  (DIF)  butterfly u < 16 takes z[u + 16 r], r < R1, from the staged waveform (windowed), radix-R1 in
//                  registers, twiddles W_M^{u k1}, writes Y[k1][u] to zb[(16 k1 + u)][frame];
//   pass C         unit q < R1 reads its 16 contiguous points, radix-16 in registers -> bins q + R1 k2, k2 < 16;
//                  units q and R1 - q are one task, so the real-FFT split Z[k], Z[M-k] -> X[k], X[M-k] stays in
//                  registers as in the three-pass engine (2 x 16 complex values per thread).
//
// Per 1024-point frame: 12 KB instead of 20 KB through shared memory, two barriers instead of three, and a radix-32
// butterfly costs ~2.8 flop per point per index bit against 3.7 for radix-8.  n = 2048 would need a paired radix-32 last
// pass (2 x 32 complex values per thread), which does not fit the register file: it stays on the three-pass engine.
// Layout, lane mapping (16 frames of one butterfly per half-warp), tables and the lane-rotation overlap-add are those of
// se_fft.cuh; Geo3 inherits its stage / working-set geometry.
#pragma once
#include "se_fft.cuh"

namespace se {

// natural-order radix-32: DIF radix-2 split into two radix-16 transforms
template <bool INV>
__device__ __forceinline__ void dft32(float2* a) {
    constexpr float C[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                             0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                             -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
    constexpr float S[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                             0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                             1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
    float2 e[16], o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        e[j] = cadd(a[j], a[j + 16]);
        const float2 d = csub(a[j], a[j + 16]);
        o[j] = j == 0 ? d : (j == 8 ? cmi<INV>(d) : ctw<INV>(d, C[j], S[j]));      // d * W32^{+-j}
    }
    dft16<INV>(e);
    dft16<INV>(o);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        a[2 * k] = e[k];
        a[2 * k + 1] = o[k];
    }
}

template <int R, bool INV>
__device__ __forceinline__ void dftR3(float2* a) {
    if (R == 16) dft16<INV>(a);
    else dft32<INV>(a);
}

// ------------------------------------------------------------------ geometry
template <int N_, int HOP_, int NT_>
struct Geo3 : Geo<N_, HOP_, NT_, 16> {
    using Base = Geo<N_, HOP_, NT_, 16>;
    static constexpr int RC = 16;                      // radix of the bin-side pass
    static constexpr int R1 = Base::M / RC;            // radix of the time-side pass: 16 (n = 512) or 32 (n = 1024)
    static constexpr int TA = RC / Base::NU;           // pass-A butterflies per thread (16 per frame)
    static constexpr int TC = (R1 / 2) / Base::NU;     // paired pass-C tasks per thread (R1 / 2 per frame)
    static constexpr int S = R1;                       // bin stride between the 16 outputs of a pass-C unit
    static constexpr int SEG = (R1 / Base::OLA) > 0 ? (R1 / Base::OLA) : 1;      // float2 per hop segment per pass-A butterfly
    static_assert(R1 == 16 || R1 == 32, "two-pass engine: n_fft 512 or 1024");
    static_assert(TA >= 1 && TC >= 1, "too many threads for this size");
    static_assert(Base::FR == 16 && Base::ZSKEW == 0, "frames in half-warps");
};

template <class G> __device__ __forceinline__ int task3_qa(int p) { return p; }
template <class G> __device__ __forceinline__ int task3_qb(int p) { return p == 0 ? G::S / 2 : G::S - p; }

// ------------------------------------------------------------------ forward (DIF)
template <class G>
__device__ __forceinline__ void passA3_fwd(const float* __restrict__ stage, const float* __restrict__ win,
                                           const float2* __restrict__ tw, float2* __restrict__ zb, int unit, int fr) {
#pragma unroll
    for (int i = 0; i < G::TA; ++i) {
        const int u = unit + i * G::NU;
        float2 a[G::R1];
#pragma unroll
        for (int r = 0; r < G::R1; ++r) {
            const int j = 2 * (u + G::RC * r);
            const float2 x = *reinterpret_cast<const float2*>(stage + (fr + j / G::HOP) * G::SROW + j % G::HOP);
            const float2 w = *reinterpret_cast<const float2*>(win + j);
            a[r] = make_float2(x.x * w.x, x.y * w.y);
        }
        dftR3<G::R1, false>(a);
#pragma unroll
        for (int k = 1; k < G::R1; ++k) a[k] = cmul(a[k], tw[u * k]);
#pragma unroll
        for (int k = 0; k < G::R1; ++k) zb[(G::RC * k + u) * G::FR + fr] = a[k];
    }
}

template <class G>
__device__ __forceinline__ void passC3_fwd_unit(const float2* __restrict__ zb, int q, int fr, float2* z) {
    const float2* p = zb + (G::RC * q) * G::FR + fr;
#pragma unroll
    for (int r = 0; r < G::RC; ++r) z[r] = p[r * G::FR];
    dft16<false>(z);
}

// bins of paired task p: unit qa (bins qa + S k) in xa[16], unit qb in xb[16], Nyquist in nyq (p == 0 only)
template <class G>
__device__ __forceinline__ void split_task3(int p, const float2* __restrict__ twn, float2* xa, float2* xb, float2& nyq) {
    constexpr int RC = G::RC;
    if (p != 0) {
#pragma unroll
        for (int k = 0; k < RC; ++k) split_pair(xa[k], xb[RC - 1 - k], twn[p + G::S * k]);
        nyq = make_float2(0.f, 0.f);
    } else {
        const float2 z0 = xa[0];
        xa[0] = make_float2(2.f * (z0.x + z0.y), 0.f);       // Z is pre-scaled by 1/2 through the window
        nyq = make_float2(2.f * (z0.x - z0.y), 0.f);
#pragma unroll
        for (int k = 1; k < RC / 2; ++k) split_pair(xa[k], xa[RC - k], twn[G::S * k]);
        float2 m0 = xa[RC / 2], m1 = xa[RC / 2];
        split_pair(m0, m1, twn[G::S * (RC / 2)]);
        xa[RC / 2] = m0;
#pragma unroll
        for (int k = 0; k < RC / 2; ++k) split_pair(xb[k], xb[RC - 1 - k], twn[G::S / 2 + G::S * k]);
    }
}
template <class G>
__device__ __forceinline__ void merge_task3(int p, const float2* __restrict__ twn, float2* ya, float2* yb, float2 nyq) {
    constexpr int RC = G::RC;
    if (p != 0) {
#pragma unroll
        for (int k = 0; k < RC; ++k) merge_pair(ya[k], yb[RC - 1 - k], twn[p + G::S * k]);
    } else {
        const float y0 = ya[0].x, ym = nyq.x;                   // imaginary parts of DC / Nyquist ignored
        ya[0] = make_float2(y0 + ym, y0 - ym);
#pragma unroll
        for (int k = 1; k < RC / 2; ++k) merge_pair(ya[k], ya[RC - k], twn[G::S * k]);
        float2 m0 = ya[RC / 2], m1 = ya[RC / 2];
        merge_pair(m0, m1, twn[G::S * (RC / 2)]);
        ya[RC / 2] = m0;
#pragma unroll
        for (int k = 0; k < RC / 2; ++k) merge_pair(yb[k], yb[RC - 1 - k], twn[G::S / 2 + G::S * k]);
    }
}

// ------------------------------------------------------------------ inverse (DIT)
template <class G>
__device__ __forceinline__ void passC3_inv_unit(float2* __restrict__ zb, int q, int fr, float2* z) {
    dft16<true>(z);
    float2* p = zb + (G::RC * q) * G::FR + fr;
#pragma unroll
    for (int r = 0; r < G::RC; ++r) p[r * G::FR] = z[r];
}
// pass A inverse for butterfly u: windowed time samples (y[2m], y[2m+1]), m = u + 16 r
template <class G>
__device__ __forceinline__ void passA3_inv_task(const float2* __restrict__ zb, const float* __restrict__ win,
                                                const float2* __restrict__ tw, int u, int fr, float2* a) {
#pragma unroll
    for (int k = 0; k < G::R1; ++k) a[k] = zb[(G::RC * k + u) * G::FR + fr];
#pragma unroll
    for (int k = 1; k < G::R1; ++k) a[k] = cmulc(a[k], tw[u * k]);
    dftR3<G::R1, true>(a);
#pragma unroll
    for (int r = 0; r < G::R1; ++r) {
        const float2 w = *reinterpret_cast<const float2*>(win + 2 * (u + G::RC * r));
        a[r] = make_float2(a[r].x * w.x, a[r].y * w.y);
    }
}

}  // namespace se
