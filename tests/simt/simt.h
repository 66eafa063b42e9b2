// simt.h -- TEST INFRASTRUCTURE: a small SIMT-on-CPU emulator, so that the kernel SOURCES of icp_flow_b200/csrc can be
// compiled with g++ and executed without a GPU (tests/test_simt_*.py, `-m "not gpu"`).
//
// It is not a CPU implementation of the product and the product never loads it: icp_flow_b200 has no CPU path.  The
// emulator exists so that the logic of the CUDA kernels (block barriers, warp collectives, shared-memory carve-ups,
// atomics, the stopping rules) can be checked against the oracle in the build container, where there is no GPU.
//
// Model: one CTA at a time; every CUDA thread of the CTA is a stackful fiber on ONE host thread; a fiber runs until it
// reaches a block barrier or a warp collective, where it hands over to the next fiber (round robin).  Blocks of a
// grid run one after the other.  What this does and does not check:
//   + data flow through shared/global memory, barriers (a barrier some live thread never reaches dead-locks and is
//     reported), warp shuffles / votes / match / reduce with masks, atomics (trivially atomic), TMA bulk copies
//     (memcpy + an emulated mbarrier), dynamic shared memory poisoned with NaN bytes at every block start;
//   - no data races can be observed (the schedule is deterministic), no timing, and the transcendental / approximate
//     device functions (rsqrtf, __fdividef, atan2f ...) are the host's, i.e. they differ from the GPU in the last bits.
#pragma once

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <tuple>
#include <type_traits>
#include <utility>

#define ICPF_SIMT_EMU 1

// ------------------------------------------------------------------------------------------------ qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

// ------------------------------------------------------------------------------------------------ vector types
struct alignas(16) float4 { float x, y, z, w; };
struct float3 { float x, y, z; };
struct alignas(8) float2 { float x, y; };
struct uint3 { unsigned int x, y, z; };
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline float3 make_float3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }

// ------------------------------------------------------------------------------------------------ runtime API stubs
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
namespace simt { extern int last_error; }
// like the runtime: returns and clears the error of the last launch (9 = cudaErrorInvalidConfiguration: empty grid,
// empty or oversized block, too much dynamic shared memory)
inline cudaError_t cudaGetLastError() { const int e = simt::last_error; simt::last_error = 0; return e; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == 9 ? "invalid configuration argument (simt emulator)" : "simt emulator: no CUDA runtime"; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }

// ------------------------------------------------------------------------------------------------ scheduler interface
namespace simt {

constexpr size_t kDynSharedBytes = 232448;    // 227 KB, the opt-in maximum of sm_100

struct Ids { uint3 tid; uint3 bid; dim3 bdim; dim3 gdim; int lane; int warp; };
extern Ids cur;                                 // ids of the running fiber (rewritten at every switch)

void yield();                                   // hand over to the next live fiber of the block
void syncthreads();
// warp collective: deposit `v`, wait for the lanes of `mask` that are alive, return the slot array of this exchange
const uint64_t* exchange(unsigned mask, uint64_t v, unsigned* members);
void run_grid(dim3 grid, dim3 block, size_t smem, void (*thread_fn)(void*), void* ctx);
void register_dyn_shared(void* base);          // arrays poisoned at every block start

template <class K>
struct Bound {
    dim3 grid, block;
    size_t smem;
    K kernel;
    template <class... B>
    void operator()(B&&... b) const {
        struct Ctx { K k; std::tuple<std::decay_t<B>...> args; } ctx{kernel, std::tuple<std::decay_t<B>...>(b...)};
        run_grid(grid, block, smem,
                 [](void* p) {
                     Ctx* c = static_cast<Ctx*>(p);
                     std::apply([c](auto... a) { c->k(a...); }, c->args);     // kernel parameters are passed by value
                 },
                 &ctx);
    }
};
template <class K>
inline Bound<K> bind(K k, dim3 grid, dim3 block, size_t smem) { return Bound<K>{grid, block, smem, k}; }

template <class T> inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "warp collectives carry <= 8 bytes");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T> inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace simt

#define threadIdx (simt::cur.tid)
#define blockIdx (simt::cur.bid)
#define blockDim (simt::cur.bdim)
#define gridDim (simt::cur.gdim)
constexpr int warpSize = 32;

inline void __syncthreads() { simt::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { unsigned m; simt::exchange(mask, 0, &m); }

// ------------------------------------------------------------------------------------------------ warp collectives
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    unsigned m;
    const uint64_t* s = simt::exchange(mask, simt::to_bits(v), &m);
    const int lane = simt::cur.lane, base = lane & ~(width - 1);
    return simt::from_bits<T>(s[base + (src & (width - 1))]);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    unsigned m;
    const uint64_t* s = simt::exchange(mask, simt::to_bits(v), &m);
    const int lane = simt::cur.lane, src = lane ^ x;
    return (src / width == lane / width && src < 32) ? simt::from_bits<T>(s[src]) : v;
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned m;
    const uint64_t* s = simt::exchange(mask, simt::to_bits(v), &m);
    const int lane = simt::cur.lane, src = lane - (int)d;
    return (src >= (lane & ~(width - 1))) ? simt::from_bits<T>(s[src]) : v;
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned m;
    const uint64_t* s = simt::exchange(mask, simt::to_bits(v), &m);
    const int lane = simt::cur.lane, src = lane + (int)d;
    return (src < (lane & ~(width - 1)) + width) ? simt::from_bits<T>(s[src]) : v;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned m, r = 0;
    const uint64_t* s = simt::exchange(mask, pred ? 1u : 0u, &m);
    for (int l = 0; l < 32; ++l) if (((m >> l) & 1u) && s[l]) r |= 1u << l;
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) {
    unsigned m;
    const uint64_t* s = simt::exchange(mask, pred ? 1u : 0u, &m);
    for (int l = 0; l < 32; ++l) if (((m >> l) & 1u) && !s[l]) return 0;
    return 1;
}
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    unsigned m, r = 0;
    const uint64_t mine = simt::to_bits(v);
    const uint64_t* s = simt::exchange(mask, mine, &m);
    for (int l = 0; l < 32; ++l) if (((m >> l) & 1u) && s[l] == mine) r |= 1u << l;
    return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    unsigned m, r = 0;
    const uint64_t* s = simt::exchange(mask, v, &m);
    for (int l = 0; l < 32; ++l) if ((m >> l) & 1u) r += (unsigned)s[l];
    return r;
}
inline int __reduce_add_sync(unsigned mask, int v) { return (int)__reduce_add_sync(mask, (unsigned)v); }

// ------------------------------------------------------------------------------------------------ atomics (one host thread)
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, int v) { unsigned o = *p; *p = o + (unsigned)v; return o; }
template <class T> inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; *p = o > v ? o : v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; *p = o < v ? o : v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ------------------------------------------------------------------------------------------------ intrinsics
// Compiled with -ffp-contract=off -fno-fast-math on SSE2: +, -, *, /, sqrtf, fmaf are the IEEE round-to-nearest operations.
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __fdividef(float a, float b) { return a / b; }            // approximate on the device
inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }               // approximate (2 ulp) on the device
inline int __float2int_rd(float a) { return (int)floorf(a); }
inline int __float2int_rz(float a) { return (int)a; }
inline int __float2int_rn(float a) { return (int)rintf(a); }
inline float __int_as_float(int v) { return simt::from_bits<float>((uint32_t)v); }
inline float __uint_as_float(unsigned v) { return simt::from_bits<float>(v); }
inline int __float_as_int(float v) { return (int)(uint32_t)simt::to_bits(v); }
inline unsigned __float_as_uint(float v) { return (unsigned)simt::to_bits(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline void __threadfence() {}
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
inline double __longlong_as_double(long long v) { double d; __builtin_memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; __builtin_memcpy(&v, &d, 8); return v; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
