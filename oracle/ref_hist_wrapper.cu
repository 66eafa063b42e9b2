// oracle/ref_hist_wrapper.cu -- TEST / BENCH INFRASTRUCTURE, never part of the product library.
//
// The reference's own CUDA vote kernel as the kernel-to-beat for the histogram stage: this file
// #includes /root/reference/hist_cuda/cpp/hist_cuda_core.cuh (the kernel `hist_cuda_kernel` and its launcher
// `hist_cuda_core`, hist_cuda_core.cuh:23-99) FROM THE REFERENCE TREE at build time -- nothing is copied into this
// repository -- and puts a plain C entry point in front of it.  The host loop below restates hist_cuda.cu:59-85
// (`at::zeros` + one launch per `mini_batch` pairs); calling `hist_cuda_core` directly also side-steps the one line of
// the reference that no longer compiles against torch 2.x (`AT_DISPATCH_FLOATING_TYPES(X.type(), ...)`,
// hist_cuda.cu:75), so the kernel is built unmodified, with the flags the reference uses (-O3, no arch-specific code
// beyond the target).  Built by `make -C oracle _ref/libref_hist.so` into oracle/_ref/ (git-ignored; it travels to the
// GPU box as a prebuilt file).
#include "hist_cuda_core.cuh"

extern "C" int icpf_ref_hist_f32(const float* X, const float* Y, int batch, int num_X, int num_Y, float min_x, float min_y,
                                 float min_z, float max_x, float max_y, float max_z, int len_x, int len_y, int len_z,
                                 int mini_batch, float* bins, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int dim = 4;
    cudaError_t err = cudaMemsetAsync(bins, 0, sizeof(float) * (size_t)batch * len_x * len_y * len_z, stream);   // at::zeros
    if (err != cudaSuccess) return (int)err;
    int iters = batch / mini_batch;
    if (batch % mini_batch != 0) iters += 1;
    for (int i = 0; i < iters; ++i) {
        int mini_batch_ = mini_batch;
        if ((i + 1) * mini_batch > batch) mini_batch_ = batch - i * mini_batch;
        hist_cuda_core<float>(stream, X + (size_t)i * mini_batch * num_X * dim, Y + (size_t)i * mini_batch * num_Y * dim,
                              mini_batch_, dim, num_X, num_Y, min_x, min_y, min_z, max_x, max_y, max_z, len_x, len_y, len_z,
                              bins + (size_t)i * mini_batch * len_x * len_y * len_z);
    }
    return (int)cudaGetLastError();
}
