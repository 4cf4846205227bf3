// se_api_psa.cu -- phase-sensitive spectral approximation loss,
// loss_phase_sensitive_spectral_approximation(enhance, target, mixture), src/loss.py:32-56:
//   angle_x = tanh(im / (re + 1e-9))      (sic: the reference squashes the tangent, it does not take atan)
//   loss = mean( (|enhance| - |target| cos(angle_target - angle_mixture))^2 )
// One elementwise pass over the three spectra with a deterministic two-stage reduction; the backward is a
// second elementwise pass (gradient to `enhance` only, like the reference's use).
#include "se_host.h"

using namespace se;

namespace {

__device__ __forceinline__ float psa_residual(float2 e, float2 t, float2 m, float& amp_e) {
    constexpr float EPS = 1e-9f;
    const float am = tanhf(m.y / (m.x + EPS));
    const float at = tanhf(t.y / (t.x + EPS));
    amp_e = sqrtf(e.y * e.y + e.x * e.x);
    const float amp_t = sqrtf(t.y * t.y + t.x * t.x);
    return amp_e - amp_t * cosf(at - am);
}

__global__ void __launch_bounds__(256) k_psa_fwd(const float2* __restrict__ enh, const float2* __restrict__ tgt,
                                                 const float2* __restrict__ mix, int64_t count, double* __restrict__ partials) {
    __shared__ double sh[8];
    pdl_launch_dependents();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        float amp;
        const float r = psa_residual(__ldg(enh + i), __ldg(tgt + i), __ldg(mix + i), amp);
        acc += r * r;
    }
    double d = (double)acc;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) d += __shfl_xor_sync(0xffffffffu, d, m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += sh[w];
        partials[blockIdx.x] = tot;
    }
}

__global__ void k_psa_sum(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double sh[256];
    pdl_launch_dependents();
    pdl_wait();
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += partials[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

__global__ void __launch_bounds__(256) k_psa_bwd(const float2* __restrict__ enh, const float2* __restrict__ tgt,
                                                 const float2* __restrict__ mix, const float* __restrict__ gout, float scale,
                                                 int64_t count, float2* __restrict__ genh) {
    pdl_launch_dependents();
    pdl_wait();
    const float gs = __ldg(gout) * scale;                     // scale = 2 / global_count
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float2 e = __ldg(enh + i);
        float amp;
        const float r = psa_residual(e, __ldg(tgt + i), __ldg(mix + i), amp);
        const float c = amp > 0.f ? gs * r / amp : 0.f;       // d|e|/de = e/|e| (0 at the origin; torch gives NaN there)
        genh[i] = make_float2(c * e.x, c * e.y);
    }
}

int psa_blocks(int64_t count) {
    int64_t b = (count + 255) / 256;
    return (int)(b > 148 * 8 ? 148 * 8 : b);
}

}  // namespace

extern "C" int64_t se_psa_workspace_bytes(int64_t count) { return (int64_t)psa_blocks(count) * (int64_t)sizeof(double); }

extern "C" int se_psa_loss_fwd(const float* enh, const float* tgt, const float* mix, int64_t count, double* sum_out,
                               void* workspace, void* stream) {
    if (!enh || !tgt || !mix || !sum_out || !workspace || count <= 0) return fail(SE_ERR_BAD_ARG, "null pointer or empty tensor");
    const int nb = psa_blocks(count);
    cudaError_t e = launch(k_psa_fwd, (unsigned)nb, 256u, 0, (cudaStream_t)stream, reinterpret_cast<const float2*>(enh),
                           reinterpret_cast<const float2*>(tgt), reinterpret_cast<const float2*>(mix), count,
                           reinterpret_cast<double*>(workspace));
    if (e != cudaSuccess) return cuda_fail(e, "se_psa_loss_fwd launch");
    e = launch(k_psa_sum, 1u, 256u, 0, (cudaStream_t)stream, (const double*)workspace, nb, sum_out);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_psa_loss_fwd reduce launch");
}

extern "C" int se_psa_loss_bwd(const float* enh, const float* tgt, const float* mix, const float* gout, int64_t global_count,
                               int64_t count, float* genh, void* stream) {
    if (!enh || !tgt || !mix || !gout || !genh || count <= 0 || global_count < count)
        return fail(SE_ERR_BAD_ARG, "null pointer or bad count");
    cudaError_t e = launch(k_psa_bwd, (unsigned)psa_blocks(count), 256u, 0, (cudaStream_t)stream,
                           reinterpret_cast<const float2*>(enh), reinterpret_cast<const float2*>(tgt),
                           reinterpret_cast<const float2*>(mix), gout, (float)(2.0 / (double)global_count), count,
                           reinterpret_cast<float2*>(genh));
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_psa_loss_bwd launch");
}
