// Test infrastructure (not part of libmaxib200.so): runs the straight-line division / reciprocal / oscillator increment of
// csrc/bank_kernels.cuh next to the compiler's own operators on the same operands (q, r: the unchecked sequences; i: osc_increment with its test), so that tests/test_gpu_ieee_div.py can compare the bits.
// Built by maximilian_b200.build.build_selftest() into tests/cuda/libmxbselftest.so with the product's flags (-fmad=false).
#include "../../maximilian_b200/csrc/bank_kernels.cuh"

namespace {
__global__ void div_check_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n, double* __restrict__ q_fn,
                                 double* __restrict__ q_op, double* __restrict__ r_fn, double* __restrict__ r_op, double* __restrict__ i_fn,
                                 double* __restrict__ i_op) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = a[i], y = b[i];
    q_fn[i] = mxb::div_rn_unchecked(x, y);
    q_op[i] = x / y;
    r_fn[i] = mxb::rcp_rn_unchecked(y);
    r_op[i] = 1.0 / y;
    i_fn[i] = mxb::osc_increment(x, y);            // 1./(sampleRate/frequency)
    i_op[i] = 1. / (x / (y));
}
}  // namespace

// host arrays in, host arrays out; returns 0 or the CUDA error code
extern "C" int mxbtest_div(const double* a, const double* b, long long n, double* q_fn, double* q_op, double* r_fn, double* r_op, double* i_fn,
                           double* i_op) {
    double* d[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const size_t bytes = sizeof(double) * (size_t)n;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 8 && e == cudaSuccess; ++k) e = cudaMalloc((void**)&d[k], bytes);
    if (e == cudaSuccess) e = cudaMemcpy(d[0], a, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d[1], b, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        div_check_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d[0], d[1], n, d[2], d[3], d[4], d[5], d[6], d[7]);
        e = cudaGetLastError();
    }
    double* host[6] = {q_fn, q_op, r_fn, r_op, i_fn, i_op};
    for (int k = 0; k < 6 && e == cudaSuccess; ++k) e = cudaMemcpy(host[k], d[2 + k], bytes, cudaMemcpyDeviceToHost);
    for (int k = 0; k < 8; ++k) cudaFree(d[k]);
    return (int)e;
}
