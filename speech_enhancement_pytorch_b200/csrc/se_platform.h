// se_platform.h -- CUDA runtime include + launch helper.
//
// The product is built with nvcc for sm_100a.  With -DSE_EMULATE (tests/cuda_emu, g++ only, no
// GPU) the same sources compile against a fiber-based emulator so kernel index math can be
// debugged in the build container; that build is test infrastructure and never shipped/loaded
// by the package.
#pragma once

#ifdef SE_EMULATE
#include "cuda_emu.h"
#define SE_SMEM_DECL unsigned char* se_smem = emu::g_dyn_smem
#define SE_NOINLINE __attribute__((noinline))
#else
#include <cuda_runtime.h>
#define SE_SMEM_DECL extern __shared__ __align__(16) unsigned char se_smem[]
#define SE_NOINLINE __noinline__
#endif

#include <cstddef>
#include <cstdint>
#include <mutex>
#include <map>
#include <set>
#include <utility>

namespace se {

// Programmatic dependent launch (sm_90+): every kernel lets its successor start launching as soon
// as all of its own CTAs are resident, and waits for its predecessors right before it first touches
// their output.  The successor's prologue (index setup, table staging into shared memory) and its
// ramp-up then overlap the predecessor's partially filled last wave.
__device__ __forceinline__ void pdl_launch_dependents() {
#ifndef SE_EMULATE
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#ifndef SE_EMULATE
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// Single-instruction MUFU forms (rsqrt / rcp / lg2 .approx.ftz).  The CUDA intrinsics rsqrtf, __fdividef and
// __log2f wrap the same MUFU in denormal/range fix-ups (a compare and two predicated multiplies each); every call
// site here passes a clamped, normal-range argument, so the per-bin math drops those instructions.  Denormal
// inputs are flushed to zero (rsqrt, rcp -> inf; lg2 -> -inf): callers guard with fmaxf / select.
__device__ __forceinline__ float se_rsqrt(float x) {
#ifdef SE_EMULATE
    return 1.0f / sqrtf(x);
#else
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float se_rcp(float x) {
#ifdef SE_EMULATE
    return 1.0f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float se_log2(float x) {
#ifdef SE_EMULATE
    return log2f(x);
#else
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// cp.async (LDGSTS): 8-byte global -> shared copies that bypass registers; the stage rows are 8-byte aligned (their pitch
// is == 2 words mod 32), so this is the widest form they take.  commit + wait bracket a fill.
__device__ __forceinline__ void se_cp_async8(void* smem_dst, const void* gmem_src) {
#ifdef SE_EMULATE
    *reinterpret_cast<float2*>(smem_dst) = *reinterpret_cast<const float2*>(gmem_src);
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void se_cp_async_commit() {
#ifndef SE_EMULATE
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void se_cp_async_wait_all() {
#ifndef SE_EMULATE
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// A product that must stay a rounded product: without this the compiler contracts `x*y - u*v` into an FMA whose
// inner product is unrounded, so the difference of two bit-identical products comes out as their rounding error, not 0.
__device__ __forceinline__ float se_mul_rn(float a, float b) {
#ifdef SE_EMULATE
    return a * b;
#else
    return __fmul_rn(a, b);
#endif
}

// Evict-first store for scratch that is written once and read once, much later (the loss workspace): keeps it from
// displacing the step's reusable lines in L2.
template <class T>
__device__ __forceinline__ void se_store_stream(T* p, T v) {
#ifdef SE_EMULATE
    *p = v;
#else
    __stcs(p, v);
#endif
}

// pdl = true: the kernel may start while its stream predecessor is still running (it must call
// pdl_wait() before touching the predecessor's output); pdl = false: classic stream serialisation.
template <class... Params, class... Args>
inline cudaError_t launch_ex(bool pdl, void (*kern)(Params...), unsigned grid, unsigned block, size_t smem,
                             cudaStream_t stream, Args... args) {
#ifdef SE_EMULATE
    (void)stream; (void)pdl;
    emu::launch(dim3(grid), dim3(block), smem, [&]() { kern(args...); });
    return cudaSuccess;
#else
    if (smem > 32 * 1024) {      // static __shared__ counts against the 48 KB default too: opt in early
        // opt in to large dynamic shared memory once per (device, kernel); keeps the steady-state
        // launch path free of driver calls (and CUDA-graph capturable)
        static std::mutex mu;
        static std::map<std::pair<int, const void*>, size_t> done;      // largest size opted in so far
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(mu);
        const auto key = std::make_pair(dev, reinterpret_cast<const void*>(kern));
        auto it = done.find(key);
        if (it == done.end() || it->second < smem) {          // kernels with a run-time working set (se_generic.cuh) grow it
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            done[key] = smem;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<Params>(args)...);
#endif
}

template <class... Params, class... Args>
inline cudaError_t launch(void (*kern)(Params...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                          Args... args) {
    return launch_ex(true, kern, grid, block, smem, stream, args...);
}

}  // namespace se
