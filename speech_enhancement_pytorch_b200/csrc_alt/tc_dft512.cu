// tc_dft512.cu -- MEASURED ALTERNATIVE, not on the product path: the n_fft = 512 / hop 128 STFT (reflect-centred,
// windowed, one-sided, reference layout [rows, F, T, 2]) as a DFT-as-GEMM on the tensor pipe, the variant
// BASELINE.json's north_star asks to be "tried as a measured alternative for small n_fft" and the reference's own
// formulation of DCCRN's transform (a dense basis GEMM, src/model/dccrn.py:649-666,691).
//
//   X[t, c] = sum_j p[128 t + j] * W[j, c],   W = window * scale * {cos, -sin}(2 pi j k / 512)
//
// A is never materialised: frame t is a 512-sample view of the padded signal at offset 128 t (a Toeplitz view of the
// staged span, rows skewed by 4 words so the 8 fragment rows hit distinct banks).  B (512 x 512) is precomputed on the
// host in double, split into TF32 hi + lo parts; fp32-grade accuracy comes from the 3xTF32 scheme
// (a_hi b_hi + a_lo b_hi + a_hi b_lo, error ~2^-21).  Columns are ordered (re_k, im_k) pairs, column 1 carries the
// Nyquist bin (im_0 = 0), so a warp's accumulator tile stores float2 runs along t.
// mma.sync.m16n8k8 TF32 (HMMA in SASS) is the SIMT-visible tensor path; profiles/r02_notes.md has the numbers
// (including why a tcgen05 version cannot change the verdict).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

constexpr int N = 512, HOP = 128, F = 257;
constexpr int MT = 64;                     // frames per CTA
constexpr int NTILE = 128;                 // output columns per CTA
constexpr int KC = 32;                     // K chunk staged per step
constexpr int SPAN = N + (MT - 1) * HOP;   // staged samples
constexpr int PROW = HOP + 4;              // span stored in rows of 128 samples, pitch 132 words
constexpr int BROW = NTILE + 8;            // B chunk [k][n], pitch 136 words
constexpr int NTHREADS = 128;              // 4 warps, 16 frames each

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float reflect_at(const float* __restrict__ x, int nsample, int i) {   // padded coordinate i
    int j = i - N / 2;
    j = j < 0 ? -j : j;
    j = j >= nsample ? 2 * (nsample - 1) - j : j;
    return (j >= 0 && j < nsample) ? __ldg(x + j) : 0.f;
}

// grid: (column tiles = 4, frame tiles, rows).  PASSES = 3: the 3xTF32 split; PASSES = 1: plain TF32 (for the accuracy figure)
template <int PASSES>
__global__ void __launch_bounds__(NTHREADS) k_tc_stft512(const float* __restrict__ x, float* __restrict__ spec,
                                                         const float* __restrict__ bhi, const float* __restrict__ blo,
                                                         int nsample, int nframe) {
    extern __shared__ __align__(16) float smem[];
    float* sp = smem;                                  // [ (SPAN/HOP) rows ][PROW]
    float* sb = smem + (SPAN / HOP) * PROW;            // [2 (hi, lo)][KC][BROW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int ct = blockIdx.x, ft = blockIdx.y, row = blockIdx.z;
    const int t0 = ft * MT;
    const float* xr = x + (size_t)row * nsample;
    for (int i = tid; i < SPAN; i += NTHREADS) sp[(i / HOP) * PROW + i % HOP] = reflect_at(xr, nsample, t0 * HOP + i);
    float acc[NTILE / 8][4];
#pragma unroll
    for (int n = 0; n < NTILE / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    const int fr = warp * 16 + g;                      // this lane's fragment rows: fr and fr + 8
    for (int k0 = 0; k0 < N; k0 += KC) {
        __syncthreads();                               // span ready (first pass) / previous chunk consumed
        for (int i = tid; i < KC * NTILE / 4; i += NTHREADS) {
            const int kk = i / (NTILE / 4), n4 = i - kk * (NTILE / 4);
            const size_t src = (size_t)(k0 + kk) * N + ct * NTILE + 4 * n4;
            *reinterpret_cast<float4*>(sb + kk * BROW + 4 * n4) = __ldg(reinterpret_cast<const float4*>(bhi + src));
            *reinterpret_cast<float4*>(sb + KC * BROW + kk * BROW + 4 * n4) = __ldg(reinterpret_cast<const float4*>(blo + src));
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < KC; ks += 8) {
            // A fragment: element (frame r, sample j) = sp[(r + j / 128) row][j % 128]
            uint32_t ah[4], al[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = fr + (q & 1) * 8, j = k0 + ks + t4 + (q >> 1) * 4;
                const float v = sp[(r + j / HOP) * PROW + j % HOP];
                ah[q] = to_tf32(v);
                al[q] = to_tf32(v - __uint_as_float(ah[q]));
            }
#pragma unroll
            for (int n = 0; n < NTILE / 8; ++n) {
                const float* bh = sb + (ks + t4) * BROW + n * 8 + g;
                const float* bl = bh + KC * BROW;
                const uint32_t h0 = __float_as_uint(bh[0]), h1 = __float_as_uint(bh[4 * BROW]);
                const uint32_t l0 = __float_as_uint(bl[0]), l1 = __float_as_uint(bl[4 * BROW]);
                if (PASSES == 3) {
                    mma_tf32(acc[n], al, h0, h1);      // small terms first
                    mma_tf32(acc[n], ah, l0, l1);
                }
                mma_tf32(acc[n], ah, h0, h1);
            }
        }
    }
    // accumulator (frame r, column c): c0 = n*8 + 2*t4 holds (re_k, im_k) of bin k = (ct*128 + c0) / 2 in d[0], d[1]
    // (frame fr) and d[2], d[3] (frame fr + 8); column 1 carries the Nyquist bin, im_0 = 0
    float2* out = reinterpret_cast<float2*>(spec) + (size_t)row * F * nframe;
#pragma unroll
    for (int n = 0; n < NTILE / 8; ++n) {
        const int k = (ct * NTILE + n * 8 + 2 * t4) / 2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = t0 + fr + 8 * h;
            if (t >= nframe) continue;
            const float re = acc[n][2 * h], im = acc[n][2 * h + 1];
            if (k == 0) {
                out[(size_t)0 * nframe + t] = make_float2(re, 0.f);
                out[(size_t)(N / 2) * nframe + t] = make_float2(im, 0.f);
            } else {
                out[(size_t)k * nframe + t] = make_float2(re, im);
            }
        }
    }
}

struct Basis { float *hi = nullptr, *lo = nullptr; int win = 0; float scale = 0.f; };
std::mutex g_mu;
Basis g_basis;

float tf32_round(float x) {                            // round to nearest, ties away (cvt.rna), on the host
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

}  // namespace

extern "C" int se_alt_tc_stft512(const float* x, float* spec, int64_t rows, int64_t nsample, int win_length, float scale,
                                 int passes, void* stream) {
    if (!x || !spec || rows <= 0 || nsample <= N / 2 || win_length < 2 || win_length > N) return -1;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (!g_basis.hi || g_basis.win != win_length || g_basis.scale != scale) {
            std::vector<float> hi((size_t)N * N), lo((size_t)N * N);
            const double two_pi = 6.283185307179586476925286766559;
            const int left = (N - win_length) / 2;
            for (int j = 0; j < N; ++j) {
                const int jw = j - left;
                const double w = (jw >= 0 && jw < win_length) ? (0.5 - 0.5 * std::cos(two_pi * jw / win_length)) * scale : 0.0;
                for (int c = 0; c < N; ++c) {
                    const int k = c / 2;
                    double v;
                    if (c == 1) v = w * std::cos(two_pi * j * (N / 2) / N);            // Nyquist, real
                    else if (c & 1) v = -w * std::sin(two_pi * ((int64_t)j * k % N) / N);
                    else v = w * std::cos(two_pi * ((int64_t)j * k % N) / N);
                    const float h = tf32_round((float)v);
                    hi[(size_t)j * N + c] = h;
                    lo[(size_t)j * N + c] = tf32_round((float)(v - (double)h));
                }
            }
            if (!g_basis.hi) {
                if (cudaMalloc((void**)&g_basis.hi, sizeof(float) * N * N) != cudaSuccess) return -3;
                if (cudaMalloc((void**)&g_basis.lo, sizeof(float) * N * N) != cudaSuccess) return -3;
            }
            cudaMemcpy(g_basis.hi, hi.data(), sizeof(float) * N * N, cudaMemcpyHostToDevice);
            cudaMemcpy(g_basis.lo, lo.data(), sizeof(float) * N * N, cudaMemcpyHostToDevice);
            g_basis.win = win_length;
            g_basis.scale = scale;
        }
    }
    const int nframe = (int)(1 + nsample / HOP);
    const size_t smem = sizeof(float) * ((SPAN / HOP) * PROW + 2 * KC * BROW);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_tc_stft512<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_tc_stft512<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    dim3 grid(N / NTILE, (nframe + MT - 1) / MT, (unsigned)rows);
    if (passes == 1) k_tc_stft512<1><<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(x, spec, g_basis.hi, g_basis.lo, (int)nsample, nframe);
    else k_tc_stft512<3><<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(x, spec, g_basis.hi, g_basis.lo, (int)nsample, nframe);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
