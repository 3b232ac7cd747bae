// Micro-benchmark (diagnostic): DRAM bandwidth of the tile-persistent access pattern -- every CTA streams R read
// vectors and W written vectors of `rows` x 32 B out of its own chunk, pass after pass -- against a grid-stride copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o streams streams.cu && ./streams
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int R, int W>
__global__ void __launch_bounds__(512) k_tile(double2 *base, size_t chunk2, int n2, int reps)
{
    double2 *c = base + (size_t)blockIdx.x * chunk2;
    for (int rep = 0; rep < reps; ++rep) {
        // rotate the vectors so that nothing stays in cache across reps
        const int rot = rep % 8;
#pragma unroll 4
        for (int i = threadIdx.x; i < n2; i += blockDim.x) {
            double2 s = make_double2(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double2 v = c[(size_t)((r + rot) % 8) * n2 + i];
                s.x += v.x;
                s.y += v.y;
            }
#pragma unroll
            for (int w = 0; w < W; ++w) c[(size_t)((w + rot + 4) % 8) * n2 + i] = s;
        }
        __syncthreads();
    }
}

__global__ void k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_read(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n)
{
    double2 s = make_double2(0, 0);
#pragma unroll 4
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { s.x += a[i].x; s.y += a[i].y; }
    if (s.x == 1.2345) b[0] = s;
}

template <int R, int W>
static void run_tile(double2 *base, int ctas, int nt, int rows, int reps)
{
    const int n2 = rows * 2;                       // double2 per vector (rows x 32 B)
    const size_t chunk2 = (size_t)8 * n2 + 1024;   // 8 vectors per chunk
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_tile<R, W><<<ctas, nt>>>(base, chunk2, n2, 2);
    cudaEventRecord(a);
    k_tile<R, W><<<ctas, nt>>>(base, chunk2, n2, reps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)ctas * reps * (R + W) * n2 * 16.0;
    printf("tile R=%d W=%d ctas=%4d nt=%4d rows=%6d: %8.2f ms  %7.0f GB/s   (%s)\n", R, W, ctas, nt, rows, ms, bytes / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const size_t total = (size_t)6 << 30;
    double2 *buf;
    if (cudaMalloc(&buf, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, total);
    {
        const size_t n = total / 32;  // two halves
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        for (int it = 0; it < 2; ++it) {
            cudaEventRecord(a);
            k_copy<<<148 * 16, 512>>>(buf, buf + n, n);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it) printf("grid-stride copy: %7.0f GB/s\n", 2.0 * n * 16 / ms / 1e6);
            cudaEventRecord(a);
            k_read<<<148 * 16, 512>>>(buf, buf + n, n);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            if (it) printf("grid-stride read: %7.0f GB/s\n", 1.0 * n * 16 / ms / 1e6);
        }
    }
    const int rows = 17152;
    for (int ctas : {148, 296, 592}) {
        run_tile<1, 1>(buf, ctas, 512, rows, 200);
        run_tile<3, 1>(buf, ctas, 512, rows, 100);
        run_tile<4, 2>(buf, ctas, 512, rows, 60);
        run_tile<2, 0>(buf, ctas, 512, rows, 150);
    }
    run_tile<3, 1>(buf, 592, 256, rows, 100);
    run_tile<3, 1>(buf, 1184, 256, rows, 50);
    run_tile<3, 1>(buf, 296, 1024, rows, 100);
    run_tile<3, 1>(buf, 296, 512, 4096, 400);
    run_tile<3, 1>(buf, 296, 512, 140000, 12);
    return 0;
}
