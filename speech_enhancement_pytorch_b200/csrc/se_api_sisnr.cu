// se_api_sisnr.cu -- scale-invariant SNR on device: si_snr / loss_sisdr,
// src/loss.py:14-29.  One pass over both waveforms per row: the three inner products <s1,s1>, <s1,s2>,
// <s2,s2> (accumulated in double) give the projection, the target and noise energies and the SNR;
// the gradient is A(row) s1 + B(row) s2, one elementwise pass.
#include "se_host.h"

using namespace se;

namespace {

struct SnrTerms { double alpha, T, Nn, r, snr; };

__device__ __host__ inline SnrTerms snr_terms(double d11, double d12, double d22, double eps) {
    SnrTerms t;
    t.alpha = d12 / (d22 + eps);                       // s_target = alpha * s2        (loss.py:23)
    t.T = t.alpha * t.alpha * d22;                     // <s_target, s_target>          (loss.py:25)
    t.Nn = d11 - 2.0 * t.alpha * d12 + t.T;            // <e, e>, e = s1 - s_target     (loss.py:24,26)
    if (t.Nn < 0.0) t.Nn = 0.0;
    t.r = t.T / (t.Nn + eps) + eps;
    t.snr = 10.0 * log10(t.r);                         // loss.py:27
    return t;
}

__global__ void __launch_bounds__(256) k_sisnr_fwd(const float* __restrict__ s1, const float* __restrict__ s2, int n,
                                                   double eps, double* __restrict__ dots, float* __restrict__ snr) {
    __shared__ double sh[3][8];
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* a = s1 + (size_t)row * n;
    const float* b = s2 + (size_t)row * n;
    double d11 = 0.0, d12 = 0.0, d22 = 0.0;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    const int n4 = vec ? n / 4 : 0;
    for (int i = tid; i < n4; i += 256) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i);
        const float4 y = __ldg(reinterpret_cast<const float4*>(b) + i);
        // fp32 products of four samples, then one double accumulate: keeps the fp64 pipe off the critical path
        d11 += (double)(x.x * x.x + x.y * x.y) + (double)(x.z * x.z + x.w * x.w);
        d12 += (double)(x.x * y.x + x.y * y.y) + (double)(x.z * y.z + x.w * y.w);
        d22 += (double)(y.x * y.x + y.y * y.y) + (double)(y.z * y.z + y.w * y.w);
    }
    for (int i = 4 * n4 + tid; i < n; i += 256) {
        const float x = __ldg(a + i), y = __ldg(b + i);
        d11 += (double)(x * x); d12 += (double)(x * y); d22 += (double)(y * y);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        d11 += __shfl_xor_sync(0xffffffffu, d11, m);
        d12 += __shfl_xor_sync(0xffffffffu, d12, m);
        d22 += __shfl_xor_sync(0xffffffffu, d22, m);
    }
    if ((tid & 31) == 0) { sh[0][tid >> 5] = d11; sh[1][tid >> 5] = d12; sh[2][tid >> 5] = d22; }
    __syncthreads();
    if (tid == 0) {
        double t11 = 0.0, t12 = 0.0, t22 = 0.0;
        for (int w = 0; w < 8; ++w) { t11 += sh[0][w]; t12 += sh[1][w]; t22 += sh[2][w]; }
        dots[3 * row] = t11; dots[3 * row + 1] = t12; dots[3 * row + 2] = t22;
        snr[row] = (float)snr_terms(t11, t12, t22, eps).snr;
    }
}

// g = gscale * gout * d snr_row / d s1 = A s1 + B s2
__global__ void __launch_bounds__(256) k_sisnr_bwd(const float* __restrict__ s1, const float* __restrict__ s2, int n,
                                                   double eps, const double* __restrict__ dots, const float* __restrict__ gout,
                                                   float gscale, float* __restrict__ g) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.y;
    const double d11 = dots[3 * row], d12 = dots[3 * row + 1], d22 = dots[3 * row + 2];
    const SnrTerms t = snr_terms(d11, d12, d22, eps);
    const double k = 10.0 / 2.302585092994046 / t.r * (double)gscale * (double)__ldg(gout);
    const double den = t.Nn + eps;
    const double a1 = 2.0 * t.alpha * d22 / (d22 + eps);             // dT/ds1 = a1 s2
    const double es2 = d12 - t.alpha * d22;                          // <e, s2>
    const double b1 = 2.0 * es2 / (d22 + eps);                       // dNn/ds1 = 2 e - b1 s2
    const double A = k * (-2.0 * t.T / (den * den));
    const double B = k * (a1 / den + t.T * b1 / (den * den)) - A * t.alpha;
    const float Af = (float)A, Bf = (float)B;
    const float* a = s1 + (size_t)row * n;
    const float* b = s2 + (size_t)row * n;
    float* o = g + (size_t)row * n;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) o[i] = Af * __ldg(a + i) + Bf * __ldg(b + i);
}

}  // namespace

extern "C" int se_sisnr_fwd(const float* s1, const float* s2, int64_t rows, int64_t nsample, double eps, double* dots,
                            float* snr, void* stream) {
    if (!s1 || !s2 || !dots || !snr) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (rows <= 0 || nsample <= 0 || rows > 0x7fffffffLL || nsample > 0x7fffffffLL) return fail(SE_ERR_BAD_ARG, "bad shape");
    cudaError_t e = launch(k_sisnr_fwd, (unsigned)rows, 256u, 0, (cudaStream_t)stream, s1, s2, (int)nsample, eps, dots, snr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_sisnr_fwd launch");
}

extern "C" int se_sisnr_bwd(const float* s1, const float* s2, const double* dots, const float* gout, float gscale, int64_t rows,
                            int64_t nsample, double eps, float* g, void* stream) {
    if (!s1 || !s2 || !dots || !gout || !g) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (rows <= 0 || nsample <= 0 || rows > 65535 || nsample > 0x7fffffffLL) return fail(SE_ERR_BAD_ARG, "bad shape (rows <= 65535)");
    unsigned bx = (unsigned)((nsample + 255) / 256);
    if (bx > 64) bx = 64;
#ifdef SE_EMULATE
    for (unsigned r = 0; r < (unsigned)rows; ++r)      // the emulator launches 1-D grids: one row at a time
        emu::launch(dim3(bx, 1), dim3(256), 0, [&]() {
            k_sisnr_bwd(s1 + (size_t)r * nsample, s2 + (size_t)r * nsample, (int)nsample, eps, dots + 3 * r, gout, gscale,
                        g + (size_t)r * nsample);
        });
    return 0;
#else
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(bx, (unsigned)rows);
    cfg.blockDim = dim3(256);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_sisnr_bwd, s1, s2, (int)nsample, eps, dots, gout, gscale, g);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_sisnr_bwd launch");
#endif
}
