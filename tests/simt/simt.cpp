// simt.cpp -- scheduler of the SIMT-on-CPU emulator (TEST INFRASTRUCTURE, see simt.h).
#include "simt.h"

#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>

#include <vector>

#if !defined(__x86_64__)
#error "the fiber switch below is written for x86-64 (the build container and the GPU boxes)"
#endif

// Stackful context switch: saves the callee-saved registers of the System V ABI on the current stack, stores the stack
// pointer to *save_sp and continues on new_sp.
extern "C" void simt_switch(void** save_sp, void* new_sp);
asm(".text\n"
    ".globl simt_switch\n"
    ".type simt_switch,@function\n"
    "simt_switch:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size simt_switch,.-simt_switch\n");

namespace simt {

Ids cur;
int last_error = 0;

namespace {

constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    Ids ids;
    bool live = false;
};

struct Warp {
    unsigned exists = 0;       // lanes that belong to the block
    unsigned live = 0;         // lanes that have not returned from the kernel yet
    unsigned arrived = 0;
    unsigned gen = 0;
    unsigned members[2] = {0, 0};
    uint64_t slot[2][32];
};

struct Block {
    int nthreads = 0;
    int live = 0;
    int bar_arrived = 0;
    unsigned bar_gen = 0;
    Warp warps[kMaxThreads / 32];
};

std::vector<Fiber> g_fibers;
Block g_blk;
int g_running = -1;
void* g_main_sp = nullptr;
void (*g_thread_fn)(void*) = nullptr;
void* g_thread_ctx = nullptr;
std::vector<void*> g_dyn_shared;
unsigned long long g_spins = 0;       // consecutive hand-overs without any fiber making progress (dead-lock detector)

void switch_to(int next) {
    const int me = g_running;
    g_running = next;
    void** save = (me < 0) ? &g_main_sp : &g_fibers[me].sp;
    if (me >= 0) g_fibers[me].ids = cur;
    void* to;
    if (next < 0) to = g_main_sp;
    else { to = g_fibers[next].sp; cur = g_fibers[next].ids; }
    simt_switch(save, to);
}

// Hand-over order.  The schedule is deterministic, so a missing barrier only shows when the reader happens to run before
// the writer: SIMT_ORDER=reverse runs the fibers of a block from the last thread to the first (tests/simt/memcheck.sh
// runs the suite both ways) -- results that differ between the two orders point at a race.
// SIMT_ORDER=shuffle[:seed] picks the next fiber at random at every hand-over (any interleaving of the threads between
// their synchronisation points is a legal CUDA execution); slower -- a barrier completes only once every fiber has been
// drawn -- so it is meant for a subset of the tests.
const char* const g_order = getenv("SIMT_ORDER") ? getenv("SIMT_ORDER") : "forward";
const int g_step = (g_order[0] == 'r') ? -1 : 1;
const bool g_shuffle = (g_order[0] == 's');
unsigned long long g_rng = []() {
    const char* c = strchr(g_order, ':');
    return 0x9e3779b97f4a7c15ull ^ (c ? strtoull(c + 1, nullptr, 10) * 0xbf58476d1ce4e5b9ull : 0ull);
}();

int next_live(int from) {
    const int n = g_blk.nthreads;
    if (g_shuffle) {
        g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;          // xorshift64
        from = (int)(g_rng % (unsigned long long)n);
        if (g_fibers[from].live) return from;
    }
    for (int k = 1; k <= n; ++k) {
        const int j = ((from + g_step * k) % n + n) % n;
        if (g_fibers[j].live) return j;
    }
    return -1;
}

void release_barrier_if_complete() {
    if (g_blk.bar_arrived > 0 && g_blk.bar_arrived >= g_blk.live) {
        g_blk.bar_arrived = 0;
        g_blk.bar_gen++;
    }
}

void release_warp_if_complete(Warp& w) {
    // the pending exchange completes when every LIVE lane of its mask has arrived
    const unsigned need = w.members[w.gen & 1] & w.live;
    if (w.arrived != 0 && (w.arrived & need) == need) {
        w.members[w.gen & 1] = w.arrived;
        w.arrived = 0;
        w.gen++;
    }
}

void fiber_main() {
    g_thread_fn(g_thread_ctx);
    // the CUDA thread returned from the kernel: it no longer counts for barriers and collectives
    const int me = g_running;
    g_fibers[me].live = false;
    g_blk.live--;
    Warp& w = g_blk.warps[me >> 5];
    w.live &= ~(1u << (me & 31));
    release_barrier_if_complete();
    release_warp_if_complete(w);
    g_spins = 0;
    switch_to(next_live(me));      // -1 (back to run_grid) when this was the last one
    abort();                       // a finished fiber is never resumed
}

void prepare_fiber(Fiber& f) {
    if (!f.stack) {
        void* m = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) { perror("simt: mmap"); abort(); }
        f.stack = static_cast<char*>(m);
    }
    uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack) + kStackBytes) & ~uintptr_t(15);
    void** sp = reinterpret_cast<void**>(top);
    *--sp = nullptr;                                       // fake return address: fiber_main starts with rsp % 16 == 8
    *--sp = reinterpret_cast<void*>(&fiber_main);
    for (int i = 0; i < 6; ++i) *--sp = nullptr;          // rbp rbx r12 r13 r14 r15
    f.sp = sp;
}

}  // namespace

void yield() {
    if (++g_spins > 50ull * 1000 * 1000) {
        fprintf(stderr, "simt: dead-lock in block (%u,%u,%u): %d live threads, %d at the block barrier\n", cur.bid.x,
                cur.bid.y, cur.bid.z, g_blk.live, g_blk.bar_arrived);
        abort();
    }
    const int nxt = next_live(g_running);
    if (nxt >= 0 && nxt != g_running) switch_to(nxt);
}

void syncthreads() {
    const unsigned gen = g_blk.bar_gen;
    g_blk.bar_arrived++;
    g_spins = 0;
    release_barrier_if_complete();
    while (g_blk.bar_gen == gen) yield();
}

const uint64_t* exchange(unsigned mask, uint64_t v, unsigned* members) {
    Warp& w = g_blk.warps[cur.warp];
    const unsigned gen = w.gen;
    const int buf = gen & 1;
    const unsigned bit = 1u << cur.lane;
    if (!(mask & bit)) { fprintf(stderr, "simt: lane %d calls a collective whose mask %08x excludes it\n", cur.lane, mask); abort(); }
    if (w.arrived == 0) w.members[buf] = mask & w.exists;
    w.slot[buf][cur.lane] = v;
    w.arrived |= bit;
    g_spins = 0;
    release_warp_if_complete(w);
    while (w.gen == gen) yield();
    *members = w.members[buf];
    return w.slot[buf];
}

void register_dyn_shared(void* base) {
    memset(base, 0xff, kDynSharedBytes);
    g_dyn_shared.push_back(base);
}

namespace {
// A block may only touch the `smem` bytes it was launched with: everything behind them must still hold the poison.
void check_dyn_shared_tail(size_t smem) {
    const size_t lo = smem ? smem : 16, hi = lo + 65536 < kDynSharedBytes ? lo + 65536 : kDynSharedBytes;
    for (void* base : g_dyn_shared) {
        const unsigned char* b = static_cast<const unsigned char*>(base);
        for (size_t i = lo; i < hi; ++i) {
            if (b[i] != 0xff) {
                fprintf(stderr, "simt: block (%u,%u,%u) wrote dynamic shared memory at byte %zu, beyond the %zu bytes of its launch\n",
                        cur.bid.x, cur.bid.y, cur.bid.z, i, smem);
                abort();
            }
        }
    }
}
}  // namespace

void run_grid(dim3 grid, dim3 block, size_t smem, void (*thread_fn)(void*), void* ctx) {
    const int nthreads = (int)(block.x * block.y * block.z);
    // what the CUDA runtime rejects with cudaErrorInvalidConfiguration: the launch does not happen, the error is sticky
    // until cudaGetLastError() fetches it
    if (nthreads <= 0 || nthreads > kMaxThreads || smem > kDynSharedBytes || grid.x == 0 || grid.y == 0 || grid.z == 0 ||
        grid.y > 65535 || grid.z > 65535) {
        last_error = 9;
        return;
    }
    if (g_running >= 0) { fprintf(stderr, "simt: nested launch\n"); abort(); }
    if ((int)g_fibers.size() < nthreads) g_fibers.resize(nthreads);
    g_thread_fn = thread_fn;
    g_thread_ctx = ctx;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                // poison the dynamic shared memory: reading a word no thread of THIS block wrote yields NaN / 0xffff
                {   // (the tail behind `smem` is re-poisoned too: check_dyn_shared_tail looks at it after the block)
                    const size_t lo = smem ? smem : 16, n = lo + 65536 < kDynSharedBytes ? lo + 65536 : kDynSharedBytes;
                    for (void* base : g_dyn_shared) memset(base, 0xff, n);
                }
                g_blk.nthreads = nthreads;
                g_blk.live = nthreads;
                g_blk.bar_arrived = 0;
                const int nwarps = (nthreads + 31) / 32;
                for (int wi = 0; wi < nwarps; ++wi) {
                    Warp& w = g_blk.warps[wi];
                    const int lanes = nthreads - wi * 32 >= 32 ? 32 : nthreads - wi * 32;
                    w.exists = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
                    w.live = w.exists;
                    w.arrived = 0;
                }
                for (int t = 0; t < nthreads; ++t) {
                    Fiber& f = g_fibers[t];
                    prepare_fiber(f);
                    f.live = true;
                    f.ids.tid.x = t % block.x;
                    f.ids.tid.y = (t / block.x) % block.y;
                    f.ids.tid.z = t / (block.x * block.y);
                    f.ids.bid.x = bx; f.ids.bid.y = by; f.ids.bid.z = bz;
                    f.ids.bdim = block;
                    f.ids.gdim = grid;
                    f.ids.lane = t & 31;
                    f.ids.warp = t >> 5;
                }
                g_spins = 0;
                switch_to(g_step > 0 ? 0 : nthreads - 1);      // returns when the last fiber of the block has finished
                if (g_blk.live != 0) { fprintf(stderr, "simt: block ended with %d live threads\n", g_blk.live); abort(); }
                check_dyn_shared_tail(smem);
            }
}

}  // namespace simt
