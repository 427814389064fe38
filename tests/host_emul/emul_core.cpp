// TEST INFRASTRUCTURE ONLY -- scheduler of the SIMT emulator declared in fake_cuda/cuda_runtime.h.
#include <ucontext.h>
#include "fake_cuda/cuda_runtime.h"

namespace cuda_emul {

uint3 threadIdx_, blockIdx_;
dim3 blockDim_, gridDim_;

namespace {
constexpr size_t STACK_BYTES = 512 * 1024;
struct Fibre {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    uint3 tid;
    int warp = 0, lane = 0;
};
struct WarpBar {
    int arrived = 0;
    unsigned gen = 0;
    int alive = 0;
    uint64_t slots[32];
};
struct Pending { void* dst; const void* src; uint32_t bytes; uint64_t* bar; };

std::vector<Fibre> fibres;
std::vector<WarpBar> wbars;
std::vector<Pending> pending;
int cur = -1, alive_total = 0, bar_arrived = 0;
unsigned bar_gen = 0;
ucontext_t sched_ctx;
const std::function<void()>* body = nullptr;
unsigned long long n_yields = 0, idle_yields = 0;
uint64_t rng_state = 0x9e3779b97f4a7c15ull;

uint64_t rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return rng_state;
}

void progress() { idle_yields = 0; }

void complete_one_pending() {
    if (pending.empty()) return;
    const size_t k = (size_t)(rnd() % pending.size());
    const Pending p = pending[k];
    pending.erase(pending.begin() + (long)k);
    memcpy(p.dst, p.src, p.bytes);
    *p.bar = (*p.bar & 1ull) ^ 1ull;     // the phase in progress completes, transaction count cleared
    progress();
}

void trampoline() {
    (*body)();
    fibres[(size_t)cur].done = true;
}
}  // namespace

unsigned long long yields() { return n_yields; }
int lane_id() { return fibres[(size_t)cur].lane; }
uint64_t* warp_slots() { return wbars[(size_t)fibres[(size_t)cur].warp].slots; }

void yield() {
    ++n_yields;
    if (++idle_yields > 30000000ull) {
        fprintf(stderr, "cuda_emul: no progress for 3e7 yields (deadlock?) block (%u,%u) thread %u, %zu copies pending\n",
                blockIdx_.x, blockIdx_.y, threadIdx_.x, pending.size());
        abort();
    }
    Fibre& f = fibres[(size_t)cur];
    swapcontext(&f.ctx, &sched_ctx);
}

void sync_block() {
    const unsigned g = bar_gen;
    ++bar_arrived;
    while (bar_gen == g) {
        if (bar_arrived >= alive_total) { bar_arrived = 0; ++bar_gen; progress(); break; }
        yield();
    }
}

void sync_warp() {
    WarpBar& w = wbars[(size_t)fibres[(size_t)cur].warp];
    const unsigned g = w.gen;
    ++w.arrived;
    while (w.gen == g) {
        if (w.arrived >= w.alive) { w.arrived = 0; ++w.gen; progress(); break; }
        yield();
    }
}

void bulk_copy_async(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    pending.push_back(Pending{dst, src, bytes, bar});
}

void launch(dim3 grid, dim3 block, const std::function<void()>& kernel_body) {
    body = &kernel_body;
    gridDim_ = grid;
    blockDim_ = block;
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    if ((int)fibres.size() < nthreads) fibres.resize((size_t)nthreads);
    for (int t = 0; t < nthreads; ++t)
        if (!fibres[(size_t)t].stack) fibres[(size_t)t].stack = (char*)malloc(STACK_BYTES);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx_ = uint3{bx, by, bz};
                wbars.assign((size_t)nwarps, WarpBar());
                bar_arrived = 0;
                alive_total = nthreads;
                for (int t = 0; t < nthreads; ++t) {
                    Fibre& f = fibres[(size_t)t];
                    f.done = false;
                    f.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                    f.warp = t / 32;
                    f.lane = t % 32;
                    ++wbars[(size_t)f.warp].alive;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = STACK_BYTES;
                    f.ctx.uc_link = &sched_ctx;
                    makecontext(&f.ctx, trampoline, 0);
                }
                int live = nthreads, next = 0;
                while (live > 0) {
                    // late, out-of-order completion of the queued bulk copies
                    if (!pending.empty() && (rnd() & 7) == 0) complete_one_pending();
                    // run a short random burst of consecutive fibres, so that the interleaving varies
                    int t = next;
                    while (fibres[(size_t)t].done) t = (t + 1) % nthreads;
                    cur = t;
                    threadIdx_ = fibres[(size_t)t].tid;
                    swapcontext(&sched_ctx, &fibres[(size_t)t].ctx);
                    if (fibres[(size_t)t].done) {
                        --live;
                        --alive_total;
                        --wbars[(size_t)fibres[(size_t)t].warp].alive;
                        progress();
                    }
                    next = (t + 1 + (int)(rnd() % 3)) % nthreads;
                }
                while (!pending.empty()) complete_one_pending();
            }
    body = nullptr;
    cur = -1;
}

}  // namespace cuda_emul
