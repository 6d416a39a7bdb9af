// dram_pattern.cu -- does the DRAM read rate of a short streaming kernel depend on how many
// contiguous bytes each warp request takes from one weight row?  (decode-regime 4-bit GEMV question)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dram_pattern dram_pattern.cu
// A [rows x row_bytes] byte matrix is read once.  Each warp load instruction covers `piece` contiguous
// bytes in each of (512 / piece) rows; a warp issues 4 such loads (2 KB in flight) before consuming.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ src, unsigned* sink, int rows,
                                                   int row_vec /*16-byte vectors per row*/, int piece_vec) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int rows_per_req = 32 / piece_vec;           // rows touched by one warp load
    const int pieces_per_row = row_vec / piece_vec;    // column pieces
    const int row_groups = rows / rows_per_req;
    const long long units = (long long)row_groups * pieces_per_row;  // one unit = one warp load
    unsigned acc = 0;
    for (long long u0 = (long long)gw * 4; u0 < units; u0 += (long long)nw * 4) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long u = u0 + j;
            v[j] = make_uint4(0, 0, 0, 0);
            if (u < units) {
                // consecutive units walk along the row first (like consecutive k-chunks)
                const long long rg = u / pieces_per_row;
                const int pc = (int)(u % pieces_per_row);
                const int r = (int)rg * rows_per_req + lane / piece_vec;
                const int c = pc * piece_vec + lane % piece_vec;
                v[j] = __ldcs(src + (long long)r * row_vec + c);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main(int argc, char** argv) {
    const int rows = argc > 1 ? atoi(argv[1]) : 14336, row_bytes = argc > 2 ? atoi(argv[2]) : 2048;
    const size_t bytes = (size_t)rows * row_bytes;
    uint4* src;
    unsigned* sink;
    char* flush;
    const size_t flush_bytes = 512u << 20;
    cudaMalloc(&src, bytes);
    cudaMalloc(&sink, 4);
    cudaMalloc(&flush, flush_bytes);
    cudaMemset(src, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int piece = 16; piece <= 512; piece *= 2) {
        for (int ctas_per_sm = 2; ctas_per_sm <= 8; ctas_per_sm *= 2) {
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                cudaMemset(flush, it, flush_bytes);  // evict the matrix from L2
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                read_kernel<<<148 * ctas_per_sm, 256>>>(src, sink, rows, row_bytes / 16, piece / 16);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("piece %4d B x %2d rows per warp load, %d CTAs/SM: %7.1f us  %6.2f TB/s\n", piece, 512 / piece,
                   ctas_per_sm, best * 1e3, bytes / (best * 1e-3) / 1e12);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
