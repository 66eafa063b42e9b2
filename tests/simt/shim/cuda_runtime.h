// stands in for <cuda_runtime.h> when the kernel sources are compiled for the SIMT-on-CPU emulator (tests/simt/simt.h)
#pragma once
#include "../simt.h"
