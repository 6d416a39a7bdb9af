// Micro-benchmark (developer tool): raw tcgen05.mma issue rate per SM for the operand modes the
// GEMM kernels use.  Shared memory / TMEM contents are garbage; only timing matters.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../bitsandbytes_b200/csrc mma_peak.cu -o mma_peak
#include "sm100_ptx.cuh"
#include <cstdio>
using namespace bnb200;

template <int MODE, int N>  // MODE 0: A from TMEM (TS), 1: A from smem (SS); UMMA 128 x N x 16, bf16
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, long long* cycles_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc<512>(&slot);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = ptx::make_idesc(1, 1, 1, 128, N);
        const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem));
        const uint64_t adesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + 64 * 1024));
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (MODE == 0)
                    ptx::mma_f16_ts(tmem, tmem + 256 + 8 * j + 32 * (i & 3), bdesc + 2 * j, idesc, 1u);
                else
                    ptx::mma_f16_ss(tmem, adesc + 2 * j, bdesc + 2 * j, idesc, 1u);
            }
        }
        ptx::tc_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        cycles_out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_dyn(tmem, 512);
    }
}

template <int MODE, int N> void run(const char* name, int iters) {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    auto k = peak_kernel<MODE, N>;
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<148, 128, smem>>>(10, d);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148, 128, smem>>>(iters, d);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    double macs = double(iters) * 4 * 128 * N * 16;
    printf("%-28s err=%d  %8.1f cycles/mma  %7.1f MAC/cycle/SM  chip %.1f TFLOPS (event time %.3f ms)\n", name, (int)err,
           avg / (iters * 4.0), macs / avg, 2.0 * macs * 148 / (ms * 1e-3) / 1e12, ms);
    cudaFree(d);
}

int main() {
    run<0, 256>("TS  128x256x16 (A in TMEM)", 20000);
    run<1, 256>("SS  128x256x16 (A in smem)", 20000);
    run<0, 128>("TS  128x128x16", 20000);
    run<1, 128>("SS  128x128x16", 20000);
    run<0, 64>("TS  128x64x16", 20000);
    run<0, 16>("TS  128x16x16", 20000);
    return 0;
}
