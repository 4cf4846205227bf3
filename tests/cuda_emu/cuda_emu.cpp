// Scheduler for the test-only CUDA emulator (see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {

Block* g_blk = nullptr;
unsigned char* g_dyn_smem = nullptr;
uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

static const size_t kStack = 256 * 1024;

static void release_block(Block* b) {
    b->block_gen++;
    b->block_arrived = 0;
    for (auto& f : b->fibers)
        if (f.state == WAIT_BLOCK) f.state = READY;
}
static void release_warp(Block* b, int w) {
    b->warp_gen[w]++;
    b->warp_arrived[w] = 0;
    int lo = w * 32, hi = lo + 32 < b->nthreads ? lo + 32 : b->nthreads;
    for (int i = lo; i < hi; ++i)
        if (b->fibers[i].state == WAIT_WARP) b->fibers[i].state = READY;
}

static void fiber_entry() {
    Block* b = g_blk;
    b->body();
    int me = b->cur;
    Fiber& f = b->fibers[me];
    f.state = DONE;
    b->live--;
    b->warp_live[me / 32]--;
    if (b->live > 0 && b->block_arrived == b->live) release_block(b);
    int w = me / 32;
    if (b->warp_live[w] > 0 && b->warp_arrived[w] == b->warp_live[w]) release_warp(b, w);
    swapcontext(&f.ctx, &b->sched);
}

void yield_block() {
    Block* b = g_blk;
    Fiber& f = b->fibers[b->cur];
    f.state = WAIT_BLOCK;
    if (++b->block_arrived == b->live) release_block(b);
    swapcontext(&f.ctx, &b->sched);
}

void yield_warp() {
    Block* b = g_blk;
    int w = b->cur / 32;
    Fiber& f = b->fibers[b->cur];
    f.state = WAIT_WARP;
    if (++b->warp_arrived[w] == b->warp_live[w]) release_warp(b, w);
    swapcontext(&f.ctx, &b->sched);
}

uint32_t exchange(uint32_t v, int src) {
    Block* b = g_blk;
    b->slot[b->cur] = v;
    yield_warp();
    uint32_t r = b->slot[src];
    yield_warp();
    return r;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    Block blk;
    int n = (int)(block.x * block.y * block.z);
    blk.nthreads = n;
    blk.fibers.resize(n);
    blk.slot.resize(n);
    blk.body = body;
    blk.bdim = block;
    int nw = (n + 31) / 32;
    std::vector<char> stacks((size_t)n * kStack);
    std::vector<unsigned char> dyn(smem + 256);
    g_dyn_smem = dyn.data() + (128 - ((uintptr_t)dyn.data() & 127));
    g_blockDim = block;
    g_gridDim = grid;
    g_blk = &blk;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = uint3{bx, by, bz};
                std::memset(dyn.data(), 0xFF, dyn.size());     // NaN-poison: catch uninitialised reads
                blk.live = n;
                blk.block_arrived = 0;
                blk.warp_arrived.assign(nw, 0);
                blk.warp_gen.assign(nw, 0);
                blk.warp_live.assign(nw, 0);
                for (int i = 0; i < n; ++i) {
                    blk.warp_live[i / 32]++;
                    Fiber& f = blk.fibers[i];
                    f.state = READY;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)i * kStack;
                    f.ctx.uc_stack.ss_size = kStack;
                    f.ctx.uc_link = &blk.sched;
                    makecontext(&f.ctx, fiber_entry, 0);
                }
                while (blk.live > 0) {
                    bool progressed = false;
                    for (int i = 0; i < n; ++i) {
                        if (blk.fibers[i].state != READY) continue;
                        blk.cur = i;
                        g_threadIdx = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y),
                                            (unsigned)(i / (block.x * block.y))};
                        swapcontext(&blk.sched, &blk.fibers[i].ctx);
                        progressed = true;
                    }
                    if (!progressed) {
                        std::fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): barrier divergence\n", bx, by, bz);
                        std::abort();
                    }
                }
            }
    g_blk = nullptr;
}

}  // namespace emu
