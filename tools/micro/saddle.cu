// Micro-benchmark (diagnostic): the product's sparse apply (program.cuh, op_spmm) on an operator dumped by the library
// (PMC_DUMP_SELL=<prefix>), stand-alone in the tile-persistent layout: y = A x with the fused dot product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o saddle.exe saddle.cu && ./saddle.exe <file> [nV]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../parelagmc_b200/csrc/program.cuh"

using namespace pmc;

template <int NT, bool WEIGHTED, bool STAGED>
__global__ void __launch_bounds__(NT, 2) k_apply(Op o0, double *base, long long stride, long long xoff, long long yoff, int reps)
{
    __shared__ Smem sm;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    StageCtx sc;
    sc.buf = reinterpret_cast<WarpStage *>(dyn_smem) + (threadIdx.x >> 5) * NSTAGE;
    sc.bar = smem_u32(&sm.bars[threadIdx.x >> 5][0]);
    sc.phase = 0u;
    if ((threadIdx.x & 31) == 0) {
        for (int j = 0; j < NSTAGE; ++j) mbar_init(sc.bar + 8 * j, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double *chunk = base + (size_t)blockIdx.x * stride;
    for (int rep = 0; rep < reps; ++rep) {
        Op o = o0;
        o.x.off = (rep & 1) ? yoff : xoff;
        o.y.off = (rep & 1) ? xoff : yoff;
        op_spmm<NT, 1, EP_AX, WEIGHTED, false, true, STAGED>(o, chunk, sm, sc);
        __syncthreads();
    }
}

int main(int argc, char **argv)
{
    if (argc < 2) { printf("usage: saddle.exe <dump.bin>\n"); return 1; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { printf("cannot open %s\n", argv[1]); return 1; }
    int hdr[4];
    long long nb;
    if (fread(hdr, sizeof hdr, 1, f) != 1) return 1;
    if (fread(&nb, sizeof nb, 1, f) != 1) return 1;
    const int rows = hdr[0], cols = hdr[1], weighted = hdr[2], noff = hdr[3];
    std::vector<int> off(noff);
    std::vector<unsigned char> pk(nb);
    if (fread(off.data(), sizeof(int), noff, f) != (size_t)noff) return 1;
    if (fread(pk.data(), 1, nb, f) != (size_t)nb) return 1;
    fclose(f);
    int maxw = 0;
    long long hist[16] = {0};
    for (int i = 0; i + 1 < noff; ++i) { const int w = off[i + 1] - off[i]; maxw = w > maxw ? w : maxw; hist[w < 15 ? w : 15]++; }
    printf("%s: %d x %d, %s, %d slices, widest %d, widths:", argv[1], rows, cols, weighted ? "weighted" : "plain", noff - 1, maxw);
    for (int w = 0; w < 16; ++w) if (hist[w]) printf(" %d:%lld", w, hist[w]);
    printf("\n");
    // weight rows: the largest weight index + 1
    int nV = 1;
    if (weighted)
        for (int sl = 0; sl + 1 < noff; ++sl) {
            const int w = off[sl + 1] - off[sl];
            const int *wi = (const int *)(pk.data() + (size_t)off[sl] * 16 * 16 + (size_t)w * 16 * 12);
            for (int k = 0; k < w * 16; ++k) nV = wi[k] + 1 > nV ? wi[k] + 1 : nV;
        }
    int *doff; unsigned char *dpk;
    cudaMalloc(&doff, off.size() * 4); cudaMalloc(&dpk, pk.size());
    cudaMemcpy(doff, off.data(), off.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dpk, pk.data(), pk.size(), cudaMemcpyHostToDevice);
    const int n = rows > cols ? rows : cols;
    long long offs = 0;
    auto take = [&](long long r) { long long o = offs; offs += r * TW; return o; };
    const long long xo = take(n), yo = take(n), vo = take(nV);
    offs += 10LL * n * TW;   // the rest of a tile's chunk
    const long long stride = offs;
    const int maxcta = 296;
    double *base;
    cudaMalloc(&base, (size_t)maxcta * stride * 8);
    {
        std::vector<double> h((size_t)stride, 0.0);
        for (long long i = 0; i < (long long)n * TW; ++i) h[xo + i] = 1e-3 * (i % 97);
        for (long long i = 0; i < (long long)nV * TW; ++i) h[vo + i] = 1.0 + 0.01 * (i % 13);
        for (int b = 0; b < maxcta; ++b) cudaMemcpy(base + (size_t)b * stride, h.data(), (size_t)stride * 8, cudaMemcpyHostToDevice);
    }
    Op o;
    memset(&o, 0, sizeof o);
    o.kind = OP_SPMM; o.n = rows; o.flags = 0; o.slot = 0;
    o.rowptr = doff; o.pk = dpk;
    o.x.off = xo; o.y.off = yo; o.v.off = vo; o.r.off = o.d.off = o.w.off = -1;
    const double opbytes = (2.0 * rows + (weighted ? nV : 0)) * 32;
    printf("n %d weight rows %d, algorithmic bytes per apply and tile %.0f KB\n", rows, nV, opbytes / 1e3);
    auto timeit = [&](const char *name, int ctas, auto launch) {
        const int reps = 100;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        launch(ctas, 4);
        cudaEventRecord(a);
        launch(ctas, reps);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("%-44s ctas %4d: %8.2f us/apply  %7.0f GB/s  (%s)\n", name, ctas, ms * 1e3 / reps, opbytes * ctas * reps / ms / 1e6,
               cudaGetErrorString(cudaGetLastError()));
    };
#define RUN(NAME, NT, W, ST)                                                                                             \
    {                                                                                                                    \
        const size_t dyn = (NT / 32) * NSTAGE * sizeof(WarpStage);                                                       \
        cudaFuncSetAttribute(k_apply<NT, W, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);                 \
        for (int ctas : {148, 296})                                                                                      \
            timeit(NAME, ctas, [&](int n_, int r_) { k_apply<NT, W, ST><<<n_, NT, dyn>>>(o, base, stride, xo, yo, r_); }); \
    }
    if (weighted) {
        RUN("product apply, 448 threads, staged", 448, true, true)
        RUN("product apply, 448 threads, entries from L2", 448, true, false)
        RUN("product apply, 512 threads, staged", 512, true, true)
    } else {
        RUN("product apply, 448 threads, staged", 448, false, true)
        RUN("product apply, 448 threads, entries from L2", 448, false, false)
        RUN("product apply, 512 threads, staged", 512, false, true)
    }
    return 0;
}
