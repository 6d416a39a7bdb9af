// compat.cu -- the four element-wise helpers the reference's loader expects next to the paged-memory entry points
// (reference csrc/pythonInterface.cpp:586-592: CMAKE_ELEMENTWISE_FUNC fill / arange / _mul; kernel kfunc,
// csrc/kernels.cu:1569-1583; launcher csrc/ops.cu:453-460).  They are outside the hot path (SURVEY.md section 8b:
// "must exist though out of scope") but cost three lines each, so they are real kernels rather than stubs.
// Reference ABI: no stream argument -> legacy default stream.
#include "common.cuh"

namespace bnb200 {

namespace {

enum { kFill = 0, kArange = 1, kMul = 2 };

template <typename T, int FUNC> __global__ void elementwise_kernel(T* A, const T* B, T value, long n) {
    for (long i = (long)blockDim.x * blockIdx.x + threadIdx.x; i < n; i += (long)blockDim.x * gridDim.x) {
        if (FUNC == kFill) A[i] = value;
        if (FUNC == kArange) A[i] = (T)i;
        if (FUNC == kMul) A[i] = A[i] * B[i];
    }
}

} // namespace

template <typename T, int FUNC> void launch_elementwise(T* A, const T* B, T value, long n) {
    if (n <= 0) return;
    long blocks = (n + 511) / 512;
    if (blocks > 65535) blocks = 65535;
    elementwise_kernel<T, FUNC><<<(int)blocks, 512>>>(A, B, value, n);
    BNB200_CHECK_LAUNCH("elementwise");
}

template void launch_elementwise<float, kFill>(float*, const float*, float, long);
template void launch_elementwise<unsigned char, kFill>(unsigned char*, const unsigned char*, unsigned char, long);
template void launch_elementwise<float, kArange>(float*, const float*, float, long);
template void launch_elementwise<float, kMul>(float*, const float*, float, long);

} // namespace bnb200
