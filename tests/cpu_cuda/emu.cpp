// CPU execution model for the library's plain-CUDA kernels: cooperative fibers (ucontext), one host thread.
// See stub/cuda_runtime.h.  TEST INFRASTRUCTURE.
#include <ucontext.h>
#include <sys/mman.h>
#include <map>
#include <utility>
#include <vector>

#include "cuda_runtime.h"

namespace emu { struct MBar { int count = 0, pending = 0; long long tx = 0; unsigned phase = 0; }; }

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {

unsigned char* dyn_smem = nullptr;
static unsigned long g_named_calls = 0, g_block_barriers = 0, g_launches = 0;
static const size_t STACK = 256 * 1024;

struct Coll { unsigned mask = 0, arrived = 0; unsigned long gen = 0; uint64_t vals[32], snap[2][32]; };
// Context switch: on x86-64 a dozen instructions (callee-saved registers + stack pointer) instead of swapcontext, whose
// two sigprocmask system calls per switch dominated the run time (a block barrier is blockDim switches).
#if defined(__x86_64__)
#define EMU_FAST_SWITCH 1
struct Ctx { void* sp; };
extern "C" void emu_switch(Ctx* from, Ctx* to);
asm(".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "pushq %rbp\npushq %rbx\npushq %r12\npushq %r13\npushq %r14\npushq %r15\n"
    "movq %rsp, (%rdi)\nmovq (%rsi), %rsp\n"
    "popq %r15\npopq %r14\npopq %r13\npopq %r12\npopq %rbx\npopq %rbp\nret\n"
    ".size emu_switch,.-emu_switch\n");
#else
#define EMU_FAST_SWITCH 0
struct Ctx { ucontext_t uc; };
static void emu_switch(Ctx* from, Ctx* to) { swapcontext(&from->uc, &to->uc); }
#endif
struct Fiber { Ctx ctx; bool done = false; };

static Ctx g_main;
static std::vector<Fiber> g_fib;
static std::vector<char*> g_stacks;
static std::vector<std::vector<Coll>> g_coll;        // per warp
static const std::function<void()>* g_body = nullptr;
static int g_cur = 0, g_n = 0, g_alive = 0, g_arrived = 0, g_or_acc = 0, g_or_res[2];
static unsigned long g_bar_gen = 0, g_events = 0;

// Thread and block schedules.  A kernel that is free of races between its barriers, and whose outputs do not depend on
// the order in which blocks run or atomics land, computes the same bytes under every schedule: 0 = forward (threads
// 0..n-1 inside a sweep, blocks in grid order), 1 = reverse, 2 = a fresh pseudo-random permutation for every sweep and
// every grid, 3 = the same at warp granularity (warps permuted, the lanes of a warp back to back in lane order: a
// converged warp polls an mbarrier as ONE instruction, so its lanes cannot observe different phases -- the
// "all lanes wait, one elected lane acts, __syncwarp" idiom of the tcgen05 kernels relies on that, and mode 2, which
// lets another warp run between two lanes' polls, deadlocks it by construction).
// SSG_EMU_SCHED=forward|reverse|random[:seed]|warps[:seed], or ssg_emu_set_sched() between launches.
static int g_sched = -1;
static unsigned g_rng = 1u;
static void sched_init() {
    if (g_sched >= 0) return;
    const char* e = getenv("SSG_EMU_SCHED");
    g_sched = 0;
    if (e && !strncmp(e, "reverse", 7)) g_sched = 1;
    if (e && !strncmp(e, "random", 6)) { g_sched = 2; g_rng = e[6] == ':' ? (unsigned)atoi(e + 7) * 2654435761u + 1u : 1u; }
    if (e && !strncmp(e, "warps", 5)) { g_sched = 3; g_rng = e[5] == ':' ? (unsigned)atoi(e + 6) * 2654435761u + 1u : 1u; }
}
static unsigned next_rand() { g_rng = g_rng * 1664525u + 1013904223u; return g_rng >> 8; }
static void make_order(std::vector<int>& o, int n) {
    o.resize(n);
    for (int i = 0; i < n; ++i) o[i] = g_sched == 1 ? n - 1 - i : i;
    if (g_sched >= 2) for (int i = n - 1; i > 0; --i) { const int j = (int)(next_rand() % (unsigned)(i + 1)); std::swap(o[i], o[j]); }
}
static void make_thread_order(std::vector<int>& o, int n) {
    if (g_sched != 3) { make_order(o, n); return; }
    const int nw = (n + 31) / 32;
    std::vector<int> w(nw);
    for (int i = 0; i < nw; ++i) w[i] = i;
    for (int i = nw - 1; i > 0; --i) { const int j = (int)(next_rand() % (unsigned)(i + 1)); std::swap(w[i], w[j]); }
    o.clear();
    for (int i = 0; i < nw; ++i) for (int t = w[i] * 32; t < n && t < w[i] * 32 + 32; ++t) o.push_back(t);
}

static void set_thread_idx(int t) {
    threadIdx.x = t % blockDim.x;
    threadIdx.y = (t / blockDim.x) % blockDim.y;
    threadIdx.z = t / (blockDim.x * blockDim.y);
}
static void yield_() { emu_switch(&g_fib[g_cur].ctx, &g_main); }
int lane_id() { return g_cur & 31; }

static void release_barrier_if_complete() {
    if (g_alive > 0 && g_arrived >= g_alive) {
        g_or_res[g_bar_gen & 1] = g_or_acc;
        g_or_acc = 0;
        g_arrived = 0;
        ++g_bar_gen;
        ++g_events;
    }
}
int syncthreads_or(int pred) {
    ++g_block_barriers;
    const unsigned long g = g_bar_gen;
    g_or_acc |= pred != 0;
    ++g_arrived;
    release_barrier_if_complete();
    // non-forward schedules: the thread that completes the barrier does not get a head start -- who runs first after a
    // barrier is decided by the sweep order alone
    if (g_sched > 0 && g_bar_gen != g) yield_();
    while (g_bar_gen == g) yield_();
    ++g_events;                             // a thread leaving a wait is progress (deadlock = nobody leaves one)
    return g_or_res[g & 1];
}
void syncthreads() { (void)syncthreads_or(0); }

const uint64_t* exchange(unsigned mask, uint64_t v) {
    std::vector<Coll>& wc = g_coll[g_cur >> 5];
    Coll* c = nullptr;
    for (Coll& e : wc) if (e.mask == mask) { c = &e; break; }
    if (!c) { wc.emplace_back(); c = &wc.back(); c->mask = mask; }
    const int lane = g_cur & 31;
    if (!((mask >> lane) & 1u)) { fprintf(stderr, "emu: lane %d calls a collective whose mask %08x excludes it\n", lane, mask); abort(); }
    const unsigned long g = c->gen;
    c->vals[lane] = v;
    c->arrived |= 1u << lane;
    // lanes beyond the block's last thread do not exist: a full mask means "every lane there is"
    unsigned need = mask;
    const int warp_first = (g_cur >> 5) << 5;
    if (warp_first + 32 > g_n) need &= (g_n - warp_first >= 32) ? 0xffffffffu : ((1u << (g_n - warp_first)) - 1u);
    if ((c->arrived & need) == need) {
        memcpy(c->snap[g & 1], c->vals, sizeof(c->vals));
        c->arrived = 0;
        ++c->gen;
        ++g_events;
        if (g_sched > 0) {                  // no head start for the lane that completes the collective (see syncthreads_or)
            yield_();
            std::vector<Coll>& w2 = g_coll[g_cur >> 5];
            for (Coll& e : w2) if (e.mask == mask) { c = &e; break; }
        }
    } else {
        // vector may grow (emplace_back by another lane with a new mask): re-find after every yield
        while (true) {
            yield_();
            std::vector<Coll>& w2 = g_coll[g_cur >> 5];
            c = nullptr;
            for (Coll& e : w2) if (e.mask == mask) { c = &e; break; }
            if (c->gen != g) break;
        }
    }
    ++g_events;
    return c->snap[g & 1];
}

void yield_now() { yield_(); }

// bar.sync id, count: a barrier among `count` threads of the block
struct NamedBar { int arrived = 0; unsigned long gen = 0; };
static NamedBar g_named[16];
static void print_stats() {
    if (getenv("SSG_EMU_STATS"))
        fprintf(stderr, "emu stats: %lu launches, %lu block-barrier arrivals, %lu named-barrier arrivals\n", g_launches,
                g_block_barriers, g_named_calls);
}
void named_barrier(int id, int count) {
    ++g_named_calls;
    NamedBar& b = g_named[id & 15];
    const unsigned long g = b.gen;
    if (++b.arrived >= count) { b.arrived = 0; ++b.gen; ++g_events; if (g_sched > 0) yield_(); }
    else while (b.gen == g) yield_();
    ++g_events;
}

float tmem[128][512];
static void (*g_tc_reset)() = nullptr;
void set_tc_reset(void (*f)()) { g_tc_reset = f; }
static void tc_reset_hook() { if (g_tc_reset) g_tc_reset(); }
std::map<const void*, MBar>& mbars() { static std::map<const void*, MBar> m; return m; }
void note_event() { ++g_events; }          // asynchronous-unit progress (mbarrier arrivals, TMA, MMA) counts as progress

static void trampoline() {
    (*g_body)();
    g_fib[g_cur].done = true;
    --g_alive;
    ++g_events;
    release_barrier_if_complete();          // threads that have exited count as arrived
    emu_switch(&g_fib[g_cur].ctx, &g_main);
    abort();                                // a finished fiber is never resumed
}

static void run_block(int n) {
    g_n = n; g_alive = n; g_arrived = 0; g_or_acc = 0;
    if ((int)g_fib.size() < n) g_fib.resize(n);
    while ((int)g_stacks.size() < n) {
        void* s = mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
        if (s == MAP_FAILED) { perror("emu: mmap"); abort(); }
        g_stacks.push_back((char*)s);
    }
    mbars().clear();
    tc_reset_hook();
    for (NamedBar& b : g_named) b = NamedBar();
    g_coll.assign((n + 31) / 32, std::vector<Coll>());
    for (auto& w : g_coll) w.reserve(8);
    for (int t = 0; t < n; ++t) {
        g_fib[t].done = false;
#if EMU_FAST_SWITCH
        // initial frame: six zeroed callee-saved registers, then the entry point emu_switch's `ret` jumps to, then a
        // null return address -- so that the entry sees rsp % 16 == 8 as after a call
        uintptr_t top = ((uintptr_t)g_stacks[t] + STACK) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;
        *--sp = (void*)trampoline;
        for (int r = 0; r < 6; ++r) *--sp = nullptr;
        g_fib[t].ctx.sp = sp;
#else
        getcontext(&g_fib[t].ctx.uc);
        g_fib[t].ctx.uc.uc_stack.ss_sp = g_stacks[t];
        g_fib[t].ctx.uc.uc_stack.ss_size = STACK;
        g_fib[t].ctx.uc.uc_link = &g_main.uc;
        makecontext(&g_fib[t].ctx.uc, trampoline, 0);
#endif
    }
    static std::vector<int> order;
    make_thread_order(order, n);
    while (g_alive > 0) {
        const unsigned long before = g_events;
        if (g_sched >= 2) make_thread_order(order, n);
        for (int i = 0; i < n; ++i) {
            const int t = order[i];
            if (g_fib[t].done) continue;
            g_cur = t;
            set_thread_idx(t);
            emu_switch(&g_main, &g_fib[t].ctx);
        }
        if (g_events == before && g_alive > 0) {
            fprintf(stderr, "emu: DEADLOCK in block (%u,%u,%u): %d threads alive, %d at the block barrier -- a barrier or warp "
                    "collective is not reached by all of its participants\n", blockIdx.x, blockIdx.y, blockIdx.z, g_alive, g_arrived);
            abort();
        }
    }
}

void launch(const std::function<void()>& fn, dim3 grid, dim3 block, size_t smem, cudaStream_t) {
    if (!dyn_smem) { dyn_smem = (unsigned char*)aligned_alloc(1024, 256 * 1024); atexit(print_stats); }
    ++g_launches;
    if (smem > 256 * 1024) { fprintf(stderr, "emu: %zu bytes of dynamic shared memory\n", smem); abort(); }
    gridDim = grid; blockDim = block;
    g_body = &fn;
    const int n = (int)(block.x * block.y * block.z);
    sched_init();
    if (g_sched == 0) {
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                    run_block(n);
                }
        return;
    }
    std::vector<int> blocks;                          // not the static one of run_block: that is rebuilt per sweep
    make_order(blocks, (int)(grid.x * grid.y * grid.z));
    for (int b : blocks) {
        blockIdx.x = (unsigned)b % grid.x; blockIdx.y = ((unsigned)b / grid.x) % grid.y; blockIdx.z = (unsigned)b / (grid.x * grid.y);
        run_block(n);
    }
}

}  // namespace emu

extern "C" void ssg_emu_set_sched(int mode, unsigned seed) { emu::g_sched = mode; emu::g_rng = seed * 2654435761u + 1u; }

// ---- "device memory" is host memory
extern "C" {
cudaError_t cudaMalloc(void** p, size_t bytes) { *p = aligned_alloc(256, (bytes + 255) / 256 * 256 + 256); return *p ? cudaSuccess : 2; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, enum cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, enum cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, enum cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; ++r) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp* p, int) { memset(p, 0, sizeof(*p)); p->major = 10; p->minor = 0; strcpy(p->name, "cpu-emulation"); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulation error"; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
}
