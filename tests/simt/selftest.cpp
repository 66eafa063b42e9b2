// selftest.cpp -- known-answer checks of the SIMT-on-CPU emulator itself (TEST INFRASTRUCTURE, run by tests/test_simt_selftest.py):
// warp collectives against their PTX definitions, block barriers with early-exited threads, partial warps, shared-memory
// atomics, dynamic shared memory poisoning, the emulated mbarrier / bulk copy.
#include "simt.h"

#include <stdio.h>
#include <stdlib.h>

#define ICPF_DYN_SHARED extern
namespace icpf { ICPF_DYN_SHARED __align__(128) float4 g_tile[]; }

static int g_fail = 0;
#define CHECK(cond)                                                                            \
    do {                                                                                       \
        if (!(cond)) { ++g_fail; fprintf(stderr, "selftest: %s failed at line %d (thread %u)\n", #cond, __LINE__, threadIdx.x); } \
    } while (0)

__global__ void collectives_kernel(int* out) {
    const int tid = threadIdx.x, lane = tid & 31;
    // shuffles
    CHECK(__shfl_sync(0xffffffffu, tid, 5) == (tid & ~31) + 5);
    CHECK(__shfl_xor_sync(0xffffffffu, lane, 16) == (lane ^ 16));
    CHECK(__shfl_up_sync(0xffffffffu, lane, 3) == (lane >= 3 ? lane - 3 : lane));
    CHECK(__shfl_down_sync(0xffffffffu, lane, 30) == (lane + 30 < 32 ? lane + 30 : lane));
    CHECK(__shfl_sync(0xffffffffu, lane, 9, 8) == (lane & ~7) + 1);                       // width 8: source 9 % 8 within the segment
    const double dv = __shfl_xor_sync(0xffffffffu, 0.5 * lane, 1);
    CHECK(dv == 0.5 * (lane ^ 1));
    // votes, match, reduce
    CHECK(__ballot_sync(0xffffffffu, lane % 3 == 0) == 0x49249249u);
    CHECK(__any_sync(0xffffffffu, lane == 31) == 1 && __all_sync(0xffffffffu, lane < 31) == 0);
    const unsigned grp = __match_any_sync(0xffffffffu, lane / 4);
    CHECK(grp == (0xfu << (lane & ~3)));
    CHECK(__reduce_add_sync(0xffffffffu, lane) == 496);
    // warp-aggregated atomic on shared memory + block barrier
    __shared__ int counter;
    if (tid == 0) counter = 0;
    __syncthreads();
    const unsigned vote = __ballot_sync(0xffffffffu, (tid & 1) != 0);
    int base = 0;
    if (lane == 0) base = atomicAdd(&counter, __popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    __syncthreads();
    CHECK(counter == (int)blockDim.x / 2);
    CHECK(base % 16 == 0);
    // threads that leave early do not take part in later barriers
    if (tid >= 64) return;
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = counter + (int)gridDim.x;
}

__global__ void partial_warp_kernel(int* out) {
    // 40 threads: the second warp has 8 lanes; a full-mask collective involves the lanes that exist
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned b = __ballot_sync(0xffffffffu, 1);
    CHECK(b == (tid < 32 ? 0xffffffffu : 0xffu));
    int v = lane;
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, v, o);
        if ((lane ^ o) < (tid < 32 ? 32 : 8)) v += other;
    }
    if (tid == 0) out[0] = v;
    if (tid == 32) out[1] = v;
}

__global__ void dyn_shared_kernel(int* out, const float* src, int n) {
    float* sm = reinterpret_cast<float*>(icpf::g_tile);
    // poisoned at block start: nobody wrote sm[n + 1] in this block
    if (threadIdx.x == 0) out[0] = (sm[n + 1] != sm[n + 1]) ? 1 : 0;       // NaN pattern
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = src[i] * 2.0f;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += sm[i];
    if (threadIdx.x == blockDim.x - 1) out[1] = (int)s;
}

// a deliberate race: thread t reads what thread t + 1 writes, with no barrier in between.  The deterministic schedule
// makes the outcome a function of the hand-over order -- forward: the neighbour has not run yet (stale value); reverse:
// it has.  This is what running the suite under SIMT_ORDER=reverse relies on to expose missing barriers.
__global__ void race_kernel(int* out) {
    __shared__ int cell[64];
    const int t = threadIdx.x;
    cell[t] = -1;
    __syncthreads();
    cell[t] = t;
    out[t] = cell[(t + 1) % 64];
}

extern "C" int simt_selftest() {
    g_fail = 0;
    int out[4] = {0, 0, 0, 0};
    simt::bind(collectives_kernel, dim3(3), dim3(128), 0)(out);
    if (out[0] != 64 + 3 || out[2] != 64 + 3) { ++g_fail; fprintf(stderr, "selftest: collectives_kernel wrote %d %d %d\n", out[0], out[1], out[2]); }
    int pw[2] = {-1, -1};
    simt::bind(partial_warp_kernel, dim3(1), dim3(40), 0)(pw);
    if (pw[0] != 496 || pw[1] != 28) { ++g_fail; fprintf(stderr, "selftest: partial warp sums %d %d\n", pw[0], pw[1]); }
    float src[100];
    for (int i = 0; i < 100; ++i) src[i] = (float)i;
    int ds[2] = {0, 0};
    simt::bind(dyn_shared_kernel, dim3(2), dim3(96), 100 * sizeof(float) + 64)(ds, (const float*)src, 100);
    if (ds[0] != 1 || ds[1] != 9900) { ++g_fail; fprintf(stderr, "selftest: dynamic shared memory %d %d\n", ds[0], ds[1]); }
    {
        int seen[64];
        simt::bind(race_kernel, dim3(1), dim3(64), 0)(seen);
        const char* ord = getenv("SIMT_ORDER");
        const bool reverse = ord && ord[0] == 'r';
        // (the thread that completes the barrier goes on first, then the others in hand-over order)
        // forward: thread t runs before t + 1 -> stale -1, except thread 62 whose neighbour 63 completed the barrier;
        // reverse: thread t runs after t + 1 -> fresh value, except thread 0 which completed the barrier itself
        const bool ok = reverse ? (seen[5] == 6 && seen[63] == 0 && seen[0] == -1)
                                : (seen[5] == -1 && seen[62] == 63 && seen[0] == -1);
        if (!ok) { ++g_fail; fprintf(stderr, "selftest: race kernel saw %d %d %d %d\n", seen[0], seen[5], seen[62], seen[63]); }
    }
    // an empty grid / oversized block is rejected like the runtime does (the launch does not happen, the error is fetched once)
    int untouched[2] = {7, 7};
    simt::bind(partial_warp_kernel, dim3(0), dim3(40), 0)(untouched);
    if (cudaGetLastError() != 9 || cudaGetLastError() != 0 || untouched[0] != 7) { ++g_fail; fprintf(stderr, "selftest: empty grid\n"); }
    simt::bind(partial_warp_kernel, dim3(1), dim3(2048), 0)(untouched);
    if (cudaGetLastError() != 9 || untouched[0] != 7) { ++g_fail; fprintf(stderr, "selftest: oversized block\n"); }
    return g_fail;
}
