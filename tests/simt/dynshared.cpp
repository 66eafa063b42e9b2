// dynshared.cpp -- the dynamic shared-memory arrays the kernels declare with ICPF_DYN_SHARED (TEST INFRASTRUCTURE).
// In CUDA every kernel sees its own `extern __shared__` array; the emulator runs one block at a time, so one static
// array per declared name is enough.  They are poisoned at every block start (simt.cpp).
#include "simt.h"

namespace icpf {
alignas(128) float4 g_tile[simt::kDynSharedBytes / sizeof(float4)];
float sm[simt::kDynSharedBytes / sizeof(float)];
alignas(16) float4 fsm[simt::kDynSharedBytes / sizeof(float4)];
}  // namespace icpf

namespace {
struct Reg {
    Reg() {
        simt::register_dyn_shared(icpf::g_tile);
        simt::register_dyn_shared(icpf::sm);
        simt::register_dyn_shared(icpf::fsm);
    }
} g_reg;
}  // namespace
