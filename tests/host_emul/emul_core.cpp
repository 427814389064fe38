// TEST INFRASTRUCTURE ONLY -- scheduler of the SIMT emulator declared in fake_cuda/cuda_runtime.h.
#include "fake_cuda/cuda_runtime.h"

// Context switch between fibres.  glibc's swapcontext makes a signal-mask system call per switch, which dominates
// the run time of the emulated grid-side kernels; on x86-64 a 14-instruction switch of the callee-saved registers is
// used instead (the fibres never touch the signal mask or the floating-point control words).
#if defined(__x86_64__)
extern "C" void vggp_emul_ctx_switch(void** save_sp, void* load_sp);
asm(R"(
    .text
    .globl vggp_emul_ctx_switch
    .hidden vggp_emul_ctx_switch
    .type vggp_emul_ctx_switch,@function
vggp_emul_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size vggp_emul_ctx_switch, .-vggp_emul_ctx_switch
)");
#else
#include <ucontext.h>
#endif

namespace cuda_emul {

uint3 threadIdx_, blockIdx_;
dim3 blockDim_, gridDim_;

namespace {
constexpr size_t STACK_BYTES = 512 * 1024;
struct CpAsync { void* dst; const void* src; int bytes, src_bytes; };
struct Fibre {
#if defined(__x86_64__)
    void* sp = nullptr;
#else
    ucontext_t ctx;
#endif
    char* stack = nullptr;
    bool done = false;
    uint3 tid;
    int warp = 0, lane = 0;
    std::vector<CpAsync> cp_queue;       // cp.async copies not yet performed
    std::vector<size_t> cp_groups;       // queue length at every commit_group
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
#if defined(__x86_64__)
void* sched_sp = nullptr;
#else
ucontext_t sched_ctx;
#endif
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

void to_scheduler() {
    Fibre& f = fibres[(size_t)cur];
#if defined(__x86_64__)
    vggp_emul_ctx_switch(&f.sp, sched_sp);
#else
    swapcontext(&f.ctx, &sched_ctx);
#endif
}

void to_fibre(Fibre& f) {
#if defined(__x86_64__)
    vggp_emul_ctx_switch(&sched_sp, f.sp);
#else
    swapcontext(&sched_ctx, &f.ctx);
#endif
}

void trampoline() {
    (*body)();
    fibres[(size_t)cur].done = true;
#if defined(__x86_64__)
    to_scheduler();                      // a finished fibre is never resumed
    abort();
#endif
}

void prepare(Fibre& f) {
#if defined(__x86_64__)
    // initial frame: six callee-saved registers, the entry address popped by `ret`, one slot so that the entry
    // function starts with the stack alignment of a normal call
    uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
    void** sp = reinterpret_cast<void**>(top);
    *--sp = nullptr;
    *--sp = reinterpret_cast<void*>(&trampoline);
    for (int i = 0; i < 6; ++i) *--sp = nullptr;
    f.sp = sp;
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = STACK_BYTES;
    f.ctx.uc_link = &sched_ctx;
    makecontext(&f.ctx, trampoline, 0);
#endif
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
    to_scheduler();
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

void cp_async_enqueue(void* dst, const void* src, int bytes, int src_bytes) {
    fibres[(size_t)cur].cp_queue.push_back(CpAsync{dst, src, bytes, src_bytes});
}
void cp_async_commit_group() {
    Fibre& f = fibres[(size_t)cur];
    f.cp_groups.push_back(f.cp_queue.size());
}
void cp_async_wait_group(int keep) {
    Fibre& f = fibres[(size_t)cur];
    if ((int)f.cp_groups.size() <= keep) return;
    const size_t ngroups = f.cp_groups.size() - (size_t)keep;
    const size_t upto = f.cp_groups[ngroups - 1];
    for (size_t i = 0; i < upto; ++i) {
        const CpAsync& c = f.cp_queue[i];
        memset(c.dst, 0, (size_t)c.bytes);                       // src_bytes < bytes: zero fill
        if (c.src_bytes > 0) memcpy(c.dst, c.src, (size_t)c.src_bytes);
    }
    f.cp_queue.erase(f.cp_queue.begin(), f.cp_queue.begin() + (long)upto);
    f.cp_groups.erase(f.cp_groups.begin(), f.cp_groups.begin() + (long)ngroups);
    for (size_t& g : f.cp_groups) g -= upto;
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
                    f.cp_queue.clear();
                    f.cp_groups.clear();
                    f.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                    f.warp = t / 32;
                    f.lane = t % 32;
                    ++wbars[(size_t)f.warp].alive;
                    prepare(f);
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
                    to_fibre(fibres[(size_t)t]);
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
