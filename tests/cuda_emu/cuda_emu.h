// cuda_emu.h -- a tiny CUDA-semantics emulator for DEBUGGING KERNEL INDEX MATH WITHOUT A GPU.
//
// TEST INFRASTRUCTURE ONLY.  The build container has nvcc but no GPU; this header lets the very
// same .cu translation units compile with g++ (-DSE_EMULATE -x c++) into a host library with the
// same C-ABI, where "device pointers" are host pointers.  One CUDA thread = one ucontext fiber;
// __syncthreads / __syncwarp / __shfl_*_sync are cooperative yield points with the CUDA meaning
// (all live threads of the block / warp must arrive).  Blocks run one after another.
// The product package never loads the emulated library: tests/ builds it into
// tests/cuda_emu/_build/ and only tests/test_emu_*.py use it.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) alignas(n)

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emulated error" : "no error"; }

namespace emu {

enum State { READY, WAIT_BLOCK, WAIT_WARP, DONE };

struct Fiber {
    ucontext_t ctx;
    State state;
    unsigned gen_seen;     // generation of the barrier it waits on
    char* stack;
};

struct Block {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int cur = -1;
    int nthreads = 0, live = 0;
    int block_arrived = 0;
    unsigned block_gen = 0;
    std::vector<int> warp_arrived, warp_live;
    std::vector<unsigned> warp_gen;
    std::vector<uint32_t> slot;
    std::function<void()> body;
    dim3 bdim;
};

extern Block* g_blk;
extern unsigned char* g_dyn_smem;
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void yield_block();
void yield_warp();
uint32_t exchange(uint32_t v, int src_lane_abs);   // warp-collective: publish v, read lane's v

}  // namespace emu

#define threadIdx emu::g_threadIdx
#define blockIdx emu::g_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim

static inline void __syncthreads() { emu::yield_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield_warp(); }

static inline int emu_lane() { return (int)(threadIdx.x % 32); }
static inline int emu_warp_base() { return (int)(threadIdx.x - threadIdx.x % 32); }

template <class T>
static inline T emu_shfl_abs(T v, int src_lane) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "4- or 8-byte shuffles only");
    uint32_t b[2] = {0, 0};
    std::memcpy(b, &v, sizeof(T));
    for (size_t w = 0; w < sizeof(T) / 4; ++w) b[w] = emu::exchange(b[w], emu_warp_base() + src_lane);
    T r;
    std::memcpy(&r, b, sizeof(T));
    return r;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int lane = emu_lane();
    int base = lane - lane % width;
    return emu_shfl_abs(v, base + ((src % width) + width) % width);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    int lane = emu_lane();
    int base = lane - lane % width;
    int s = (lane % width) ^ m;
    return emu_shfl_abs(v, s < width ? base + s : lane);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu_lane();
    int base = lane - lane % width;
    int s = (lane % width) - (int)d;
    return emu_shfl_abs(v, s >= 0 ? base + s : lane);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu_lane();
    int base = lane - lane % width;
    int s = (lane % width) + (int)d;
    return emu_shfl_abs(v, s < width ? base + s : lane);
}

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline void __threadfence() {}
static inline float emu_expf(float x) { return std::exp(x); }
static inline float emu_logf(float x) { return std::log(x); }
static inline float emu_log2f(float x) { return std::log2(x); }
#define __expf emu_expf
#define __logf emu_logf
#define __log2f emu_log2f
