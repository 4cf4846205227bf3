// ubench.cu -- sm_100a micro-benchmarks behind the engine design decisions of round 2:
//   * issue rate of scalar FADD/FFMA against the packed FADD2/FMUL2/FFMA2 forms,
//   * packed math co-issued with LDS.128 traffic (what an FFT pass looks like),
//   * LDS.128 where every quarter-warp reads one contiguous 128-byte row at an arbitrary row address,
//   * legacy mma.sync m16n8k8 TF32 rate (the SIMT-visible tensor path, for the DFT-as-GEMM costing).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run on one B200.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) k_math(float* out, float seed) {
    // 8 independent accumulator pairs per thread
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i, seed - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(seed * 1e-3f, -seed * 1e-3f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }      // 2 FFMA
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); }                                         // 1 FFMA2
            if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                           // 2 FADD
            if (MODE == 3) { a[i] = __fadd2_rn(a[i], c); }                                            // 1 FADD2
            if (MODE == 4) { a[i].x = a[i].x * m.x; a[i].y = a[i].y * m.y; }                           // 2 FMUL
            if (MODE == 5) { a[i] = __fmul2_rn(a[i], m); }                                            // 1 FMUL2
        }
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// radix-4-like pass: LDS.128 x4, packed math, STS.128 x4 per iteration (PACKED) against the scalar equivalent
// on the same bytes (LDS.64 x8, scalar math, STS.64 x8)
template <bool PACKED>
__global__ void __launch_bounds__(256) k_pass(float* out, int iters) {
    extern __shared__ float4 sm4[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += 256) sm4[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (PACKED) {
            float4 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = sm4[(tid + r * 1024 + it * 8) & 4095];
            float2 re[4], im[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { re[r] = make_float2(v[r].x, v[r].y); im[r] = make_float2(v[r].z, v[r].w); }
            // dft4 on packed pairs
            float2 s0r = __fadd2_rn(re[0], re[2]), s0i = __fadd2_rn(im[0], im[2]);
            float2 d0r = __fadd2_rn(re[0], make_float2(-re[2].x, -re[2].y)), d0i = __fadd2_rn(im[0], make_float2(-im[2].x, -im[2].y));
            float2 s1r = __fadd2_rn(re[1], re[3]), s1i = __fadd2_rn(im[1], im[3]);
            float2 d1r = __fadd2_rn(re[1], make_float2(-re[3].x, -re[3].y)), d1i = __fadd2_rn(im[1], make_float2(-im[3].x, -im[3].y));
            re[0] = __fadd2_rn(s0r, s1r); im[0] = __fadd2_rn(s0i, s1i);
            re[2] = __fadd2_rn(s0r, make_float2(-s1r.x, -s1r.y)); im[2] = __fadd2_rn(s0i, make_float2(-s1i.x, -s1i.y));
            re[1] = __fadd2_rn(d0r, d1i); im[1] = __fadd2_rn(d0i, make_float2(-d1r.x, -d1r.y));
            re[3] = __fadd2_rn(d0r, make_float2(-d1i.x, -d1i.y)); im[3] = __fadd2_rn(d0i, d1r);
            const float2 c = make_float2(0.6f, 0.6f), s = make_float2(0.8f, 0.8f);
#pragma unroll
            for (int r = 1; r < 4; ++r) {
                const float2 tr = __ffma2_rn(im[r], s, __fmul2_rn(re[r], c));
                const float2 ti = __ffma2_rn(make_float2(-re[r].x, -re[r].y), s, __fmul2_rn(im[r], c));
                re[r] = tr; im[r] = ti;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) sm4[(tid + r * 1024 + it * 8) & 4095] = make_float4(re[r].x, re[r].y, im[r].x, im[r].y);
            acc += re[0].x;
        } else {
            float2* sm2 = reinterpret_cast<float2*>(sm4);
            float2 a[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) a[r] = sm2[(tid + r * 1024 + it * 16) & 8191];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float2 &a0 = a[h], &a1 = a[2 + h], &a2 = a[4 + h], &a3 = a[6 + h];
                const float2 s0 = make_float2(a0.x + a2.x, a0.y + a2.y), d0 = make_float2(a0.x - a2.x, a0.y - a2.y);
                const float2 s1 = make_float2(a1.x + a3.x, a1.y + a3.y), d1 = make_float2(a1.y - a3.y, a3.x - a1.x);
                a0 = make_float2(s0.x + s1.x, s0.y + s1.y); a2 = make_float2(s0.x - s1.x, s0.y - s1.y);
                a1 = make_float2(d0.x + d1.x, d0.y + d1.y); a3 = make_float2(d0.x - d1.x, d0.y - d1.y);
                const float c = 0.6f, s = 0.8f;
                a1 = make_float2(a1.x * c + a1.y * s, a1.y * c - a1.x * s);
                a2 = make_float2(a2.x * c + a2.y * s, a2.y * c - a2.x * s);
                a3 = make_float2(a3.x * c + a3.y * s, a3.y * c - a3.x * s);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) sm2[(tid + r * 1024 + it * 16) & 8191] = a[r];
            acc += a[0].x;
        }
        __syncthreads();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// LDS.128: quarter-warp q reads the contiguous 128-byte row (row0 + q * rstride); rstride in float4 units
__global__ void __launch_bounds__(256) k_lds128(float* out, int rstride, int iters) {
    extern __shared__ float4 sm4[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += 256) sm4[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    const int lane = tid & 31, q = lane >> 3, l = lane & 7, w = tid >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int base = (w * 4 + q) * rstride * 8 + l;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float4 v = sm4[(base + r * 64) & 8191];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        base += 8;
    }
    out[blockIdx.x * blockDim.x + tid] = acc.x + acc.y + acc.z + acc.w;
}

// mma.sync m16n8k8 tf32: 4 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) k_mma_tf32(float* out, int iters) {
    unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f000000u, 0x3f000000u}, b[2] = {0x3f800000u, 0x3e800000u};
    float d[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x * 1e-6f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mma.sync m16n8k16 bf16
__global__ void __launch_bounds__(256) k_mma_bf16(float* out, int iters) {
    unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f003f00u, 0x3f003f00u}, b[2] = {0x3f803f80u, 0x3e803e80u};
    float d[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x * 1e-6f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, clk_khz);
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256));
    const char* names[6] = {"FFMA x2 (scalar)", "FFMA2", "FADD x2 (scalar)", "FADD2", "FMUL x2 (scalar)", "FMUL2"};
    for (int ctas = 2; ctas <= 8; ctas *= 2) {
        float ms[6];
        ms[0] = time_ms([&] { k_math<0><<<sms * ctas, 256>>>(out, 1.f); });
        ms[1] = time_ms([&] { k_math<1><<<sms * ctas, 256>>>(out, 1.f); });
        ms[2] = time_ms([&] { k_math<2><<<sms * ctas, 256>>>(out, 1.f); });
        ms[3] = time_ms([&] { k_math<3><<<sms * ctas, 256>>>(out, 1.f); });
        ms[4] = time_ms([&] { k_math<4><<<sms * ctas, 256>>>(out, 1.f); });
        ms[5] = time_ms([&] { k_math<5><<<sms * ctas, 256>>>(out, 1.f); });
        for (int m = 0; m < 6; ++m) {
            // lane-ops (one fp32 add/mul/fma on one lane) per SM per ns
            const double laneops = (double)ctas * 256 * ITERS * 8 * 2;
            printf("{\"bench\": \"math\", \"op\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.4f, \"lane_ops_per_sm_per_ns\": %.1f}\n",
                   names[m], ctas * 8, ms[m], laneops / (ms[m] * 1e6));
        }
    }
    CK(cudaFuncSetAttribute(k_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int ctas = 1; ctas <= 2; ++ctas) {
        const int iters = 2000;
        const float a = time_ms([&] { k_pass<true><<<sms * ctas, 256, 65536>>>(out, iters); });
        const float b = time_ms([&] { k_pass<false><<<sms * ctas, 256, 65536>>>(out, iters); });
        // both move 256 threads x 64 B x 2 (load + store) per iteration per CTA
        printf("{\"bench\": \"pass\", \"ctas_per_sm\": %d, \"packed_ms\": %.4f, \"scalar_ms\": %.4f, \"packed_ns_per_iter\": %.1f, \"scalar_ns_per_iter\": %.1f}\n",
               ctas, a, b, a * 1e6 / iters, b * 1e6 / iters);
    }
    CK(cudaFuncSetAttribute(k_lds128, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    for (int rs : {1, 2, 3, 4, 8, 16, 17, 32, 64}) {
        const int iters = 4000;
        const float ms = time_ms([&] { k_lds128<<<sms, 256, 131072>>>(out, rs, iters); });
        const double bytes = 256.0 * 16 * 8 * iters;
        printf("{\"bench\": \"lds128\", \"row_stride_x128B\": %d, \"ms\": %.4f, \"bytes_per_sm_per_ns\": %.1f}\n", rs, ms, bytes / (ms * 1e6));
    }
    {
        const int iters = 4096;
        for (int ctas = 1; ctas <= 2; ++ctas) {
            const float a = time_ms([&] { k_mma_tf32<<<sms * ctas, 256>>>(out, iters); });
            const float b = time_ms([&] { k_mma_bf16<<<sms * ctas, 256>>>(out, iters); });
            const double f_tf32 = (double)sms * ctas * 8 * iters * 4 * (2.0 * 16 * 8 * 8);
            const double f_bf16 = (double)sms * ctas * 8 * iters * 4 * (2.0 * 16 * 8 * 16);
            printf("{\"bench\": \"mma.sync\", \"warps_per_sm\": %d, \"tf32_tflops\": %.1f, \"bf16_tflops\": %.1f}\n", ctas * 8,
                   f_tf32 / (a * 1e9), f_bf16 / (b * 1e9));
        }
    }
    return 0;
}
