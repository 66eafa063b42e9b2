// Exhaustive check (GPU) of the division used by the fused histogram kernel: for a divisor b, every fp32 x in [0, b)
//     q = fma(fma(-(x * rcp(b)), b, x), rcp(b), x * rcp(b))          (Markstein's correction step)
// must equal the IEEE quotient __fdiv_rn(x, b) bit for bit (hist_cuda_core.cuh:54-56 divides by max - min).
//   nvcc -arch=sm_100a -o tools/check_fastdiv tools/check_fastdiv.cu && tools/check_fastdiv [F ...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__global__ void check(float b, unsigned int n_patterns, unsigned long long* bad, unsigned int* first_bad,
                      unsigned int* last_bad, unsigned long long* bad_bins) {
    const float inv = __frcp_rn(b);
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_patterns;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned int)i);
        const float q0 = __fmul_rn(x, inv);
        const float r = __fmaf_rn(-q0, b, x);
        const float q = __fmaf_rn(r, inv, q0);
        const float want = __fdiv_rn(x, b);
        if (__float_as_uint(q) != __float_as_uint(want)) {
            atomicAdd(bad, 1ull);
            atomicMin(first_bad, (unsigned int)i);
            atomicMax(last_bad, (unsigned int)i);
        }
        // what the kernel consumes: the bin index floor(q * len) for the histogram lengths in use
        const float lens[6] = {3.f, 41.f, 68.f, 135.f, 202.f, 269.f};
        for (int k = 0; k < 6; ++k)
            if (__float2int_rd(__fmul_rn(q, lens[k])) != __float2int_rd(__fmul_rn(want, lens[k]))) atomicAdd(bad_bins, 1ull);
    }
}

int main(int argc, char** argv) {
    unsigned long long *bad, *bad_bins;
    unsigned int *first, *last;
    cudaMallocManaged(&bad, 8);
    cudaMallocManaged(&bad_bins, 8);
    cudaMallocManaged(&first, 4);
    cudaMallocManaged(&last, 4);
    // divisors: bins.max() - bins.min() of torch.arange(-F, F + tau - 1e-8, tau) for the F values the reference uses
    // (demo.sh 2.0, --speed 1.67 -> 3.34, argparse default 3.333 and 6.666), the z axis (2 tau), and odd values
    float divisors[64];
    int nd = 0;
    const float tau = 0.1f;
    const double Fs[] = {2.0, 3.333, 3.34, 6.666, 1.0, 5.0, 10.0, 13.332};
    for (double F : Fs) {
        // torch.arange in fp32: start + i * step computed in double then cast (ATen uses accscalar_t = double on CPU,
        // float on CUDA); the max - min differences of both variants are covered by also checking neighbours
        const int len = (int)((F + tau - 1e-8 + F) / tau) + 1;
        const float mn = (float)(-F), mx = (float)(-F + (double)(len - 1) * (double)tau);
        const float d = mx - mn;
        for (int k = -2; k <= 2; ++k) {
            unsigned int u;
            memcpy(&u, &d, 4);
            u += k;
            memcpy(&divisors[nd++], &u, 4);
        }
    }
    divisors[nd++] = 0.2f;
    divisors[nd++] = 0.1f + 0.1f;
    divisors[nd++] = 0.30000001f;
    for (int i = 1; i < argc; ++i) divisors[nd++] = (float)atof(argv[i]);
    int fails = 0;
    for (int k = 0; k < nd; ++k) {
        const float b = divisors[k];
        unsigned int nb;
        memcpy(&nb, &b, 4);
        *bad = 0;
        *bad_bins = 0;
        *first = 0xffffffffu;
        *last = 0;
        check<<<148 * 8, 256>>>(b, nb, bad, first, last, bad_bins);
        cudaDeviceSynchronize();
        float lastf;
        memcpy(&lastf, last, 4);
        printf("divisor %.9g (0x%08x): %u dividends in [0, b): quotient mismatches %llu (largest dividend %.3g), bin "
               "mismatches %llu%s\n", b, nb, nb, *bad, *bad ? lastf : 0.f, *bad_bins, *bad_bins ? " <-- FAIL" : "");
        fails += *bad_bins != 0;
    }
    printf(fails ? "FAILED\n" : "all bins identical to the IEEE division\n");
    return fails != 0;
}
