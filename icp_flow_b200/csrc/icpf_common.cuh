// icpf_common.cuh -- device helpers shared by the sm_100a kernels of the ICP-Flow registration engine.
//
//   * TMA 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) that stage one cluster's padded [N,4] fp32
//     row block from HBM into shared memory,
//   * deterministic block all-reduce built from warp shuffles,
//   * closed-form 3x3 Kabsch rotation in registers (one-sided Jacobi on the cross-covariance).
//
// Nothing here links against torch/ATen; the library boundary is the C ABI in include/icpflow_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// dynamic shared memory of a kernel (see ICPF_LAUNCH in icpf_internal.h)
#ifdef ICPF_SIMT_EMU
#define ICPF_DYN_SHARED extern
#else
#define ICPF_DYN_SHARED extern __shared__
#endif

namespace icpf {

constexpr unsigned FULL_MASK = 0xffffffffu;

// ------------------------------------------------------------------------------------------------ TMA / mbarrier
#ifdef ICPF_SIMT_EMU
// Emulated mbarrier (tests/simt/): the 64-bit word holds {phase parity, arrival count of a phase, pending arrivals,
// pending transaction bytes}; a bulk copy is a memcpy that completes its bytes at once.
struct EmuMbar { uint8_t phase; uint8_t unused; uint8_t count; uint8_t pending; int32_t tx; };
static_assert(sizeof(EmuMbar) == 8, "an mbarrier is one 64-bit word");
__device__ __forceinline__ void mbar_settle(EmuMbar* b) {
    if (b->pending == 0 && b->tx == 0) {      // phase complete: flip the parity, re-arm
        b->phase ^= 1;
        b->pending = b->count;
    }
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrive_count) {
    EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
    b->phase = 0; b->unused = 0; b->count = (uint8_t)arrive_count; b->pending = (uint8_t)arrive_count; b->tx = 0;
}
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
    b->tx += (int32_t)bytes;
    b->pending -= 1;
    mbar_settle(b);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    return reinterpret_cast<EmuMbar*>(bar)->phase != (parity & 1u);      // true once the phase `parity` has completed
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) simt::yield();
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
    memcpy(smem_dst, gmem_src, bytes);
    b->tx -= (int32_t)bytes;
    mbar_settle(b);
}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrive_count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrive_count) : "memory");
}

// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D bulk copy global -> shared (TMA engine; SASS: UBLKCP).  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#endif  // ICPF_SIMT_EMU

// ------------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// All threads of the block obtain the K sums.  Deterministic: xor-butterfly inside a warp, warps added in order.
// `scratch` holds NWARPS*K floats and must not be reused for another reduction before the next __syncthreads().
template <int K, int NWARPS>
__device__ __forceinline__ void block_allreduce_sum(float (&v)[K], float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) scratch[warp * K + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float s = scratch[k];
#pragma unroll
        for (int w = 1; w < NWARPS; ++w) s += scratch[w * K + k];
        v[k] = s;
    }
}

}  // namespace icpf

#include "icpf_kabsch.h"
