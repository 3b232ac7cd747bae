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

// Experimental: the weighted apply (staged, unrolled widths 5/6) with the streaming solution update of the previous
// MINRES iteration attached to the same pass (w0 = a w0 + b w1 + c u; xs += e w0), its loads issued before the gathers
// (HOIST) or after them -- against the two operations run one after the other.
template <int NT, int MODE>   // MODE 0: apply only, 1: update only, 2: fused, update loads hoisted, 3: fused, update after the gathers
__global__ void __launch_bounds__(NT, 2) k_fused(Op o0, double *base, long long stride, long long xoff, long long yoff, long long woff, int n, int reps)
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
    constexpr int ES = 16, NW = NT / 32;
    const int sub = (threadIdx.x % LPR) * PW;
    const int *__restrict__ off = o0.rowptr;
    const unsigned char *__restrict__ pk = o0.pk;
    const double *__restrict__ V = chunk + o0.v.off + sub;
    double *__restrict__ w0 = chunk + woff + sub, *__restrict__ w1 = w0 + (size_t)n * TW, *__restrict__ u2 = w1 + (size_t)n * TW,
                        *__restrict__ xs = u2 + (size_t)n * TW;
    for (int rep = 0; rep < reps; ++rep) {
        const double *__restrict__ x = chunk + ((rep & 1) ? yoff : xoff) + sub;
        double *__restrict__ y = chunk + ((rep & 1) ? xoff : yoff) + sub;
        D2 acc = make_double2(0.0, 0.0);
        if (MODE == 1) {
#pragma unroll 4
            for (int row = threadIdx.x / LPR; row < n; row += NT / LPR) {
                const size_t ro = (size_t)row * TW;
                const D2 a = ld2c(w0 + ro), b = ld2c(w1 + ro), c = ld2c(u2 + ro);
                D2 xv = ld2c(xs + ro);
                const D2 wn = make_double2(fma(0.3, a.x, fma(0.2, b.x, 0.1 * c.x)), fma(0.3, a.y, fma(0.2, b.y, 0.1 * c.y)));
                xv.x = fma(0.05, wn.x, xv.x); xv.y = fma(0.05, wn.y, xv.y);
                st2(w0 + ro, wn); st2(xs + ro, xv);
            }
            __syncthreads();
            continue;
        }
        const int sl_end = (n + SLICE - 1) / SLICE, rs = (threadIdx.x & 31) / LPR;
        int sl = threadIdx.x >> 5, wcur = 0, wnext = 0, st = 0;
        for (int j = 0; j < NSTAGE; ++j) {
            const int slj = sl + j * NW;
            int k0 = 0, k1 = 0;
            if (slj < sl_end) { k0 = __ldg(off + slj); k1 = __ldg(off + slj + 1); stage_issue<ES>(pk, sc.buf + j, sc.bar + 8 * j, k0, k1); }
            if (j == 0) wcur = k1 - k0; else wnext = k1 - k0;
        }
        for (; sl < sl_end; sl += NW) {
            const int row = sl * SLICE + rs;
            const bool live = row < n;
            const size_t ro = (size_t)row * TW;
            const int sl2 = sl + NSTAGE * NW;
            int n0 = 0, n1 = 0;
            if (sl2 < sl_end) { n0 = __ldg(off + sl2); n1 = __ldg(off + sl2 + 1); }
            D2 a = make_double2(0, 0), b = a, c = a, xv = a;
            if (MODE == 2 && live) { a = ld2c(w0 + ro); b = ld2c(w1 + ro); c = ld2c(u2 + ro); xv = ld2c(xs + ro); }
            mbar_wait(sc.bar + 8 * st, (sc.phase >> st) & 1u);
            sc.phase ^= 1u << st;
            const unsigned char *bse = sc.buf[st].bytes;
            const int w = wcur;
            const double *__restrict__ eval = reinterpret_cast<const double *>(bse) + rs;
            const int *__restrict__ ecol = reinterpret_cast<const int *>(bse + (size_t)w * (SLICE * 8)) + rs;
            const int *__restrict__ ewid = ecol + w * SLICE;
            D2 s;
            if (w == 6) s = gather_fixed<true, 6>(eval, ecol, ewid, x, V);
            else if (w == 5) s = gather_fixed<true, 5>(eval, ecol, ewid, x, V);
            else s = make_double2(0, 0);
            if (live) {
                const D2 xr = ld2c(x + ro);
                st2(y + ro, s);
                acc.x = fma(s.x, xr.x, acc.x); acc.y = fma(s.y, xr.y, acc.y);
                if (MODE == 3) { a = ld2c(w0 + ro); b = ld2c(w1 + ro); c = ld2c(u2 + ro); xv = ld2c(xs + ro); }
                if (MODE >= 2) {
                    const D2 wn = make_double2(fma(0.3, a.x, fma(0.2, b.x, 0.1 * c.x)), fma(0.3, a.y, fma(0.2, b.y, 0.1 * c.y)));
                    xv.x = fma(0.05, wn.x, xv.x); xv.y = fma(0.05, wn.y, xv.y);
                    st2(w0 + ro, wn); st2(xs + ro, xv);
                }
            }
            __syncwarp();
            if (sl2 < sl_end) stage_issue<ES>(pk, sc.buf + st, sc.bar + 8 * st, n0, n1);
            wcur = wnext; wnext = n1 - n0; st ^= 1;
        }
        block_dot<NT, 1>(acc, sm, 0, false);
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
    if (weighted && maxw <= 6) {
        const size_t dyn = (448 / 32) * NSTAGE * sizeof(WarpStage);
        const long long woff = vo + (long long)nV * TW;   // four more vectors of n rows inside the rest of the chunk
        auto runf = [&](const char *name, auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            const int reps = 100, ctas = 296;
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            kern<<<ctas, 448, dyn>>>(o, base, stride, xo, yo, woff, rows, 4);
            cudaEventRecord(a);
            kern<<<ctas, 448, dyn>>>(o, base, stride, xo, yo, woff, rows, reps);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            printf("%-52s %8.2f us  (%s)\n", name, ms * 1e3 / reps, cudaGetErrorString(cudaGetLastError()));
        };
        runf("apply only (unrolled widths)", k_fused<448, 0>);
        runf("solution update only (6 N rows)", k_fused<448, 1>);
        runf("fused, update loads before the gathers", k_fused<448, 2>);
        runf("fused, update after the gathers", k_fused<448, 3>);
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
