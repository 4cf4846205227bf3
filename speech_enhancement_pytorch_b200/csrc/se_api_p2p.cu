// se_api_p2p.cu -- the path's ONE exchange step through NVLink / NVSwitch peer memory.
//
// Utterance-sharded MR-STFT loss: between the forward kernels and the backward kernels every rank needs the
// batch-global sums (9 doubles).  The baseline is ncclAllReduce + a loss-value kernel; here one single-CTA kernel
// does the exchange and the value: it stores this rank's 9 sums straight into a slot of every peer's exchange
// buffer (posted NVLink writes), publishes a sequence number behind a system-scope fence, polls its own buffer
// until all ranks' slots carry that sequence number, adds the slots in rank order (so every rank gets the same
// bits) and writes the reduced sums and the loss.  No host round trip, no NCCL launch on the critical path,
// CUDA-graph capturable (the sequence counter lives in device memory).
//
// One process per GPU: buffers are cudaMalloc'ed here and shared with cudaIpc handles (the Python side gathers
// the 64-byte handles over torch.distributed once per process group).
#include "se_host.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace se;

namespace {

constexpr int kMaxWorld = 16;
constexpr int kSlot = 16;                      // doubles per slot (9 used): one 128-byte line

struct P2PBuf {                                // one per rank, in that rank's device memory
    double slots[2][kMaxWorld][kSlot];         // [parity of the sequence number][source rank]
    unsigned long long flags[2][kMaxWorld];    // sequence number of the data in the slot
    unsigned long long seq;                    // local launch counter
};
// Two parities are enough: a rank posts step s+2 into the slots of step s only after it left step s+1, which
// needs every peer's step-s+1 flag -- and a peer posts that only after it finished reading step s.

struct P2PArgs {
    P2PBuf* bufs[kMaxWorld];                   // bufs[rank] is the local one
    double* sums;                              // [9] (nval = 9) or [10] (nval = 10: [9] = row count) in: this rank's; out: global
    float* loss;                               // optional
    double cnt[3];                             // nval = 9: global bins per resolution; nval = 10: bins per row
    long long spin_limit;                      // clock64 ticks before a missing peer poisons the result (0: wait forever)
    int world, rank, nval;
};

#ifndef SE_EMULATE
__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(32) k_sums_exchange(const P2PArgs a) {
    __shared__ double in[kMaxWorld][10];
    __shared__ double tot[10];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    P2PBuf* me = a.bufs[a.rank];
    unsigned long long seq = 0;
    if (tid == 0) { seq = me->seq + 1; me->seq = seq; }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    const int par = (int)(seq & 1);
    if (tid < a.world) {
        // post: lane p writes this rank's sums into peer p's slot [par][rank], then the flag (release, system scope)
        P2PBuf* peer = a.bufs[tid];
        volatile double* dst = peer->slots[par][a.rank];
        for (int j = 0; j < a.nval; ++j) dst[j] = a.sums[j];
        __threadfence_system();
        st_flag(&peer->flags[par][a.rank], seq);
        // collect: lane r waits for rank r's flag in the LOCAL buffer, then reads its slot
        const long long t0 = clock64();
        bool timed_out = false;
        while (ld_flag(&me->flags[par][tid]) != seq) {
            if (a.spin_limit > 0 && clock64() - t0 > a.spin_limit) {
                // A peer that never arrives must not hang the stream, and must not kill the CUDA context either (the
                // NCCL path this replaces would only wait): report, poison the result with NaN so the step is visibly
                // invalid, and let the host decide.  SE_P2P_SPIN_SECONDS sets the limit (default 600 s, 0 = forever).
                printf("se_mrstft_exchange_value: rank %d gave up waiting for rank %d (step %llu); result poisoned with NaN\n",
                       a.rank, tid, seq);
                timed_out = true;
                break;
            }
        }
        const volatile double* src = me->slots[par][tid];
        for (int j = 0; j < a.nval; ++j) in[tid][j] = timed_out ? __longlong_as_double(0x7ff8000000000000LL) : src[j];
    }
    __syncwarp();
    if (tid < a.nval) {
        double acc = 0.0;
        for (int r = 0; r < a.world; ++r) acc += in[r][tid];    // rank order: identical bits on every rank
        tot[tid] = acc;
        a.sums[tid] = acc;
    }
    __syncwarp();
    if (tid == 0 && a.loss) {
        const double rows = a.nval == 10 ? tot[9] : 1.0;
        double total = 0.0;
        for (int r = 0; r < 3; ++r) total += sqrt(tot[3 * r]) / sqrt(tot[3 * r + 1]) + tot[3 * r + 2] / (a.cnt[r] * rows);
        *a.loss = (float)(total / 3.0);
    }
}
#endif

}  // namespace

extern "C" {

#ifdef SE_EMULATE
// the emulator has no peer memory; these entry points exist only in the product build
int se_p2p_create(void**, unsigned char*) { return fail(SE_ERR_UNSUPPORTED, "peer exchange is not emulated"); }
int se_p2p_open(const unsigned char*, void**) { return fail(SE_ERR_UNSUPPORTED, "peer exchange is not emulated"); }
int se_p2p_close(void*) { return fail(SE_ERR_UNSUPPORTED, "peer exchange is not emulated"); }
int se_p2p_destroy(void*) { return fail(SE_ERR_UNSUPPORTED, "peer exchange is not emulated"); }
int se_mrstft_exchange_value(double*, void* const*, int, int, int64_t, int64_t, float*, void*) {
    return fail(SE_ERR_UNSUPPORTED, "peer exchange is not emulated");
}
int se_mrstft_exchange_rows_value(double*, void* const*, int, int, int64_t, float*, void*) {
    return fail(SE_ERR_UNSUPPORTED, "peer exchange is not emulated");
}
#else

int se_p2p_create(void** local, unsigned char* handle64) {
    if (!local || !handle64) return fail(SE_ERR_BAD_ARG, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(P2PBuf));
    if (e != cudaSuccess) return cuda_fail(e, "se_p2p_create cudaMalloc");
    e = cudaMemset(p, 0, sizeof(P2PBuf));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "se_p2p_create"); }
    memcpy(handle64, &h, 64);
    *local = p;
    return 0;
}

int se_p2p_open(const unsigned char* handle64, void** peer) {
    if (!handle64 || !peer) return fail(SE_ERR_BAD_ARG, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(peer, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_p2p_open cudaIpcOpenMemHandle");
}

int se_p2p_close(void* peer) {
    if (!peer) return fail(SE_ERR_BAD_ARG, "null pointer");
    cudaError_t e = cudaIpcCloseMemHandle(peer);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_p2p_close");
}

int se_p2p_destroy(void* local) {
    if (!local) return fail(SE_ERR_BAD_ARG, "null pointer");
    cudaError_t e = cudaFree(local);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_p2p_destroy");
}

static long long spin_limit_ticks() {
    static const long long v = [] {
        const char* e = std::getenv("SE_P2P_SPIN_SECONDS");
        const double sec = e ? std::atof(e) : 600.0;           // minutes, like the NCCL watchdog; 0 = wait forever
        return sec <= 0.0 ? 0LL : (long long)(sec * 2.0e9);    // clock64 ticks at ~2 GHz
    }();
    return v;
}

static int exchange_impl(double* sums, void* const* bufs, int world, int rank, int64_t global_rows, int64_t nsample, float* loss,
                         void* stream, int nval);

int se_mrstft_exchange_value(double* sums, void* const* bufs, int world, int rank, int64_t global_rows, int64_t nsample,
                             float* loss, void* stream) {
    if (global_rows <= 0) return fail(SE_ERR_BAD_ARG, "need global_rows > 0");
    return exchange_impl(sums, bufs, world, rank, global_rows, nsample, loss, stream, 9);
}
// sums10[9] carries this rank's row count in and the global row count out: uneven shards need no host-side guess
int se_mrstft_exchange_rows_value(double* sums10, void* const* bufs, int world, int rank, int64_t nsample, float* loss, void* stream) {
    return exchange_impl(sums10, bufs, world, rank, 1, nsample, loss, stream, 10);
}

static int exchange_impl(double* sums, void* const* bufs, int world, int rank, int64_t global_rows, int64_t nsample, float* loss,
                         void* stream, int nval) {
    if (!sums || !bufs) return fail(SE_ERR_BAD_ARG, "null pointer");
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(SE_ERR_BAD_ARG, "need 1 <= world <= 16 and 0 <= rank < world");
    if (nsample < 2048) return fail(SE_ERR_BAD_ARG, "need nsample >= 2048");
    P2PArgs a{};
    for (int r = 0; r < world; ++r) {
        if (!bufs[r]) return fail(SE_ERR_BAD_ARG, "null exchange buffer");
        a.bufs[r] = reinterpret_cast<P2PBuf*>(bufs[r]);
    }
    static const int res[3][2] = {{512, 128}, {1024, 256}, {2048, 512}};
    for (int r = 0; r < 3; ++r) a.cnt[r] = (double)global_rows * (res[r][0] / 2 + 1) * (double)(1 + nsample / res[r][1]);
    a.sums = sums; a.loss = loss; a.world = world; a.rank = rank; a.nval = nval;
    a.spin_limit = spin_limit_ticks();
    cudaError_t e = launch(k_sums_exchange, 1u, 32u, 0, (cudaStream_t)stream, a);
    return e == cudaSuccess ? 0 : cuda_fail(e, "se_mrstft_exchange_value launch");
}
#endif

}  // extern "C"
