// Micro-benchmark (diagnostic): the weighted Schur-complement apply of the bench's level 0 (7-point pattern on 16^3,
// per-sample values gathered from a vector of distinct values, Chebyshev epilogue) in the tile-persistent layout, to
// compare thread mappings / gather schedules outside the register budget of the interpreter kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o sapply.exe sapply.cu && ./sapply.exe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cstring>
#include "../../parelagmc_b200/csrc/program.cuh"

using pmc::TW;
using pmc::D2;
__device__ __forceinline__ D2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
using pmc::st2;

struct Mat {            // ELL, width W, column-major within slices of S rows: idx = (slice * W + k) * S + r
    int n, W;
    const int *col, *wid;   // S = 16 layout
    const int *col8, *wid8; // S = 8 layout
};
struct Chunk { long long x, y, r, d, dinv, V, stride; };

// V0: two lanes per row (16 B each), entries one after the other (the compiler's order), indices straight from global
__device__ int g_desync = 0;   // > 0: every CTA first waits a pseudo-random time up to g_desync microseconds
template <int W>
__global__ void __launch_bounds__(512, 2) k_v0(Mat A, double *base, Chunk c, int reps)
{
    double *ch = base + (size_t)blockIdx.x * c.stride;
    const int sub = (threadIdx.x & 1) * 2;
    if (g_desync > 0) {
        const unsigned h = (blockIdx.x * 2654435761u) >> 8;
        const long long wait = (long long)(h % 1000) * g_desync * 2;   // ~2 cycles per ns
        const long long t0 = clock64();
        while (clock64() - t0 < wait) { }
    }
    for (int rep = 0; rep < reps; ++rep) {
        const double *x = ch + ((rep & 1) ? c.y : c.x) + sub;
        double *y = ch + ((rep & 1) ? c.x : c.y) + sub;
        const double *r = ch + c.r + sub, *dinv = ch + c.dinv + sub, *V = ch + c.V + sub;
        double *d = ch + c.d + sub;
        for (int row = threadIdx.x / 2; row < A.n; row += 256) {
            const int sl = row / 16, rs = row % 16;
            D2 s = make_double2(0, 0);
#pragma unroll
            for (int k = 0; k < W; ++k) {
                const int idx = (sl * W + k) * 16 + rs;
                const D2 xv = ld2(x + (size_t)__ldg(A.col + idx) * TW), wv = ld2(V + (size_t)__ldg(A.wid + idx) * TW);
                s.x = fma(wv.x, xv.x, s.x);
                s.y = fma(wv.y, xv.y, s.y);
            }
            const size_t ro = (size_t)row * TW;
            const D2 rv = ld2(r + ro), di = ld2(dinv + ro), dv = ld2(d + ro), xr = ld2(x + ro);
            D2 dn = make_double2(fma(0.3, dv.x, 0.7 * di.x * (rv.x - s.x)), fma(0.3, dv.y, 0.7 * di.y * (rv.y - s.y)));
            st2(d + ro, dn);
            st2(y + ro, make_double2(xr.x + dn.x, xr.y + dn.y));
        }
        __syncthreads();
    }
}

// V1: two lanes per row, all index loads, then all gathers and streamed operands, then the arithmetic (register hungry)
template <int W, int MINB>
__global__ void __launch_bounds__(512, MINB) k_v1(Mat A, double *base, Chunk c, int reps)
{
    double *ch = base + (size_t)blockIdx.x * c.stride;
    const int sub = (threadIdx.x & 1) * 2;
    for (int rep = 0; rep < reps; ++rep) {
        const double *x = ch + ((rep & 1) ? c.y : c.x) + sub;
        double *y = ch + ((rep & 1) ? c.x : c.y) + sub;
        const double *r = ch + c.r + sub, *dinv = ch + c.dinv + sub, *V = ch + c.V + sub;
        double *d = ch + c.d + sub;
        for (int row = threadIdx.x / 2; row < A.n; row += 256) {
            const int sl = row / 16, rs = row % 16;
            int ci[W], wi[W];
#pragma unroll
            for (int k = 0; k < W; ++k) {
                const int idx = (sl * W + k) * 16 + rs;
                ci[k] = __ldg(A.col + idx);
                wi[k] = __ldg(A.wid + idx);
            }
            const size_t ro = (size_t)row * TW;
            const D2 rv = ld2(r + ro), di = ld2(dinv + ro), dv = ld2(d + ro), xr = ld2(x + ro);
            D2 xv[W], wv[W];
#pragma unroll
            for (int k = 0; k < W; ++k) {
                xv[k] = ld2(x + (size_t)ci[k] * TW);
                wv[k] = ld2(V + (size_t)wi[k] * TW);
            }
            D2 s = make_double2(0, 0);
#pragma unroll
            for (int k = 0; k < W; ++k) {
                s.x = fma(wv[k].x, xv[k].x, s.x);
                s.y = fma(wv[k].y, xv[k].y, s.y);
            }
            D2 dn = make_double2(fma(0.3, dv.x, 0.7 * di.x * (rv.x - s.x)), fma(0.3, dv.y, 0.7 * di.y * (rv.y - s.y)));
            st2(d + ro, dn);
            st2(y + ro, make_double2(xr.x + dn.x, xr.y + dn.y));
        }
        __syncthreads();
    }
}

// V2: four lanes per row (one sample = 8 B each), slices of 8 rows, everything issued before the arithmetic
template <int W, int MINB>
__global__ void __launch_bounds__(512, MINB) k_v2(Mat A, double *base, Chunk c, int reps)
{
    double *ch = base + (size_t)blockIdx.x * c.stride;
    const int sub = threadIdx.x & 3;
    for (int rep = 0; rep < reps; ++rep) {
        const double *x = ch + ((rep & 1) ? c.y : c.x) + sub;
        double *y = ch + ((rep & 1) ? c.x : c.y) + sub;
        const double *r = ch + c.r + sub, *dinv = ch + c.dinv + sub, *V = ch + c.V + sub;
        double *d = ch + c.d + sub;
        for (int row = threadIdx.x / 4; row < A.n; row += 128) {
            const int sl = row / 8, rs = row % 8;
            int ci[W], wi[W];
#pragma unroll
            for (int k = 0; k < W; ++k) {
                const int idx = (sl * W + k) * 8 + rs;
                ci[k] = __ldg(A.col8 + idx);
                wi[k] = __ldg(A.wid8 + idx);
            }
            const size_t ro = (size_t)row * TW;
            const double rv = r[ro], di = dinv[ro], dv = d[ro], xr = x[ro];
            double xv[W], wv[W];
#pragma unroll
            for (int k = 0; k < W; ++k) {
                xv[k] = x[(size_t)ci[k] * TW];
                wv[k] = V[(size_t)wi[k] * TW];
            }
            double s = 0;
#pragma unroll
            for (int k = 0; k < W; ++k) s = fma(wv[k], xv[k], s);
            const double dn = fma(0.3, dv, 0.7 * di * (rv - s));
            d[ro] = dn;
            y[ro] = xr + dn;
        }
        __syncthreads();
    }
}

// V3: streaming reference with the same bytes (no gathers): r, dinv, d, x, V (3.8 rows per row) read; d, y written
__global__ void __launch_bounds__(512, 2) k_stream(Mat A, double *base, Chunk c, int reps, int nU)
{
    double *ch = base + (size_t)blockIdx.x * c.stride;
    const int sub = (threadIdx.x & 1) * 2;
    for (int rep = 0; rep < reps; ++rep) {
        const double *x = ch + ((rep & 1) ? c.y : c.x) + sub;
        double *y = ch + ((rep & 1) ? c.x : c.y) + sub;
        const double *r = ch + c.r + sub, *dinv = ch + c.dinv + sub, *V = ch + c.V + sub;
        double *d = ch + c.d + sub;
#pragma unroll 2
        for (int row = threadIdx.x / 2; row < A.n; row += 256) {
            const size_t ro = (size_t)row * TW;
            const D2 rv = ld2(r + ro), di = ld2(dinv + ro), dv = ld2(d + ro), xr = ld2(x + ro);
            D2 s = make_double2(0, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int vr = row + k * A.n;
                if (vr < nU) { const D2 wv = ld2(V + (size_t)vr * TW); s.x += wv.x; s.y += wv.y; }
            }
            D2 dn = make_double2(fma(0.3, dv.x, 0.7 * di.x * (rv.x - s.x)), fma(0.3, dv.y, 0.7 * di.y * (rv.y - s.y)));
            st2(d + ro, dn);
            st2(y + ro, make_double2(xr.x + dn.x, xr.y + dn.y));
        }
        __syncthreads();
    }
}

// mixed: odd/even CTAs alternate out of phase between the sparse apply (v0 schedule) and a streaming update over a
// vector of NBIG rows (4 reads, 2 writes), the way tiles of the persistent kernel sit in different operations
template <int W>
__global__ void __launch_bounds__(512, 2) k_mix(Mat A, double *base, Chunk c, int reps, long long big, int nbig, int phase_mod)
{
    double *ch = base + (size_t)blockIdx.x * c.stride;
    const int sub = (threadIdx.x & 1) * 2;
    for (int rep = 0; rep < reps; ++rep) {
        if (((rep + (phase_mod ? blockIdx.x : 0)) & 1) == 0) {
            const double *x = ch + ((rep & 2) ? c.y : c.x) + sub;
            double *y = ch + ((rep & 2) ? c.x : c.y) + sub;
            const double *r = ch + c.r + sub, *dinv = ch + c.dinv + sub, *V = ch + c.V + sub;
            double *d = ch + c.d + sub;
            for (int row = threadIdx.x / 2; row < A.n; row += 256) {
                const int sl = row / 16, rs = row % 16;
                D2 s = make_double2(0, 0);
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const int idx = (sl * W + k) * 16 + rs;
                    const D2 xv = ld2(x + (size_t)__ldg(A.col + idx) * TW), wv = ld2(V + (size_t)__ldg(A.wid + idx) * TW);
                    s.x = fma(wv.x, xv.x, s.x);
                    s.y = fma(wv.y, xv.y, s.y);
                }
                const size_t ro = (size_t)row * TW;
                const D2 rv = ld2(r + ro), di = ld2(dinv + ro), dv = ld2(d + ro), xr = ld2(x + ro);
                D2 dn = make_double2(fma(0.3, dv.x, 0.7 * di.x * (rv.x - s.x)), fma(0.3, dv.y, 0.7 * di.y * (rv.y - s.y)));
                st2(d + ro, dn);
                st2(y + ro, make_double2(xr.x + dn.x, xr.y + dn.y));
            }
        } else {
            double *v = ch + big + sub;
#pragma unroll 4
            for (int row = threadIdx.x / 2; row < nbig; row += 256) {
                const size_t ro = (size_t)row * TW;
                const D2 a = ld2(v + ro), b = ld2(v + ro + (size_t)nbig * TW), e = ld2(v + ro + 2 * (size_t)nbig * TW), f = ld2(v + ro + 3 * (size_t)nbig * TW);
                st2(v + ro, make_double2(a.x + 0.5 * b.x + e.x, a.y + 0.5 * b.y + e.y));
                st2(v + ro + 3 * (size_t)nbig * TW, make_double2(f.x + a.x, f.y + a.y));
            }
        }
        __syncthreads();
    }
}

// the product's own sparse apply (program.cuh, op_spmm: weighted, Chebyshev epilogue) on the same data, outside the
// interpreter kernel
template <bool STAGED>
__global__ void __launch_bounds__(512, 2) k_product_op(pmc::Op o0, double *base, long long stride, long long xoff, long long yoff, int reps)
{
    using namespace pmc;
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
        op_spmm<512, 1, EP_CHEB, true, true, false, STAGED>(o, chunk, sm, sc);
        __syncthreads();
    }
}

// experimental copy of the staged apply: HOIST (epilogue operands before the gathers), WFIX (compile-time width, 0 = runtime),
// USEVAL (multiply by the stored coefficient)
template <bool HOIST, int WFIX, bool USEVAL>
__global__ void __launch_bounds__(512, 2) k_exp(pmc::Op o0, double *base, long long stride, long long xoff, long long yoff, int reps)
{
    using namespace pmc;
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
    constexpr int ES = 16, NW = 16;
    const int sub = (threadIdx.x % LPR) * PW;
    for (int rep = 0; rep < reps; ++rep) {
        const Op &o = o0;
        const double *__restrict__ x = chunk + ((rep & 1) ? yoff : xoff) + sub;
        double *__restrict__ y = chunk + ((rep & 1) ? xoff : yoff) + sub;
        const double *__restrict__ r = chunk + o.r.off + sub;
        double *__restrict__ d = chunk + o.d.off + sub;
        const double *__restrict__ V = chunk + o.v.off + sub;
        const double *__restrict__ dinvb = chunk + o.w.off + sub;
        const int *__restrict__ off = o.rowptr;
        const unsigned char *__restrict__ pk = o.pk;
        const double ca = o.ca, cb = o.cb;
        const int r1 = o.n, sl_end = (r1 + SLICE - 1) / SLICE;
        const int rs = (threadIdx.x & 31) / LPR;
        int sl = threadIdx.x >> 5;
        int wcur = 0, wnext = 0, st = 0;
        for (int j = 0; j < NSTAGE; ++j) {
            const int slj = sl + j * NW;
            int k0 = 0, k1 = 0;
            if (slj < sl_end) { k0 = __ldg(off + slj); k1 = __ldg(off + slj + 1); stage_issue<ES>(pk, sc.buf + j, sc.bar + 8 * j, k0, k1); }
            if (j == 0) wcur = k1 - k0; else wnext = k1 - k0;
        }
        for (; sl < sl_end; sl += NW) {
            const int row = sl * SLICE + rs;
            const bool live = row < r1;
            const size_t ro = (size_t)row * TW;
            const int sl2 = sl + NSTAGE * NW;
            int n0 = 0, n1 = 0;
            if (sl2 < sl_end) { n0 = __ldg(off + sl2); n1 = __ldg(off + sl2 + 1); }
            D2 rv = make_double2(0, 0), di = rv, dv = rv, xr = rv;
            if (HOIST && live) { rv = ld2c(r + ro); di = ld2c(dinvb + ro); dv = ld2c(d + ro); xr = ld2c(x + ro); }
            mbar_wait(sc.bar + 8 * st, (sc.phase >> st) & 1u);
            sc.phase ^= 1u << st;
            const unsigned char *bse = sc.buf[st].bytes;
            const int w = WFIX ? WFIX : wcur;
            const double *__restrict__ eval = reinterpret_cast<const double *>(bse) + rs;
            const int *__restrict__ ecol = reinterpret_cast<const int *>(bse + (size_t)w * (SLICE * 8)) + rs;
            const int *__restrict__ ewid = ecol + w * SLICE;
            D2 s = make_double2(0, 0);
            if (WFIX) {
#pragma unroll
                for (int k = 0; k < (WFIX ? WFIX : 1); ++k) {
                    const D2 xv = ld2c(x + (size_t)ecol[k * SLICE] * TW), wv = ld2c(V + (size_t)ewid[k * SLICE] * TW);
                    const double cc = USEVAL ? eval[k * SLICE] : 1.0;
                    s.x = fma(cc * wv.x, xv.x, s.x);
                    s.y = fma(cc * wv.y, xv.y, s.y);
                }
            } else {
#pragma unroll 4
                for (int k = 0; k < w; ++k) {
                    const D2 xv = ld2c(x + (size_t)ecol[k * SLICE] * TW), wv = ld2c(V + (size_t)ewid[k * SLICE] * TW);
                    const double cc = USEVAL ? eval[k * SLICE] : 1.0;
                    s.x = fma(cc * wv.x, xv.x, s.x);
                    s.y = fma(cc * wv.y, xv.y, s.y);
                }
            }
            if (live) {
                if (!HOIST) { rv = ld2c(r + ro); di = ld2c(dinvb + ro); dv = ld2c(d + ro); xr = ld2c(x + ro); }
                D2 dn = make_double2(cb * di.x * (rv.x - s.x), cb * di.y * (rv.y - s.y));
                dn.x = fma(ca, dv.x, dn.x);
                dn.y = fma(ca, dv.y, dn.y);
                st2(d + ro, dn);
                st2(y + ro, make_double2(xr.x + dn.x, xr.y + dn.y));
            }
            __syncwarp();
            if (sl2 < sl_end) stage_issue<ES>(pk, sc.buf + st, sc.bar + 8 * st, n0, n1);
            wcur = wnext;
            wnext = n1 - n0;
            st ^= 1;
        }
        __syncthreads();
    }
}

int main(int argc, char **argv)
{
    const int g = argc > 1 ? atoi(argv[1]) : 16;
    const int n = g * g * g, W = 7;
    // 7-point pattern, rank-major numbering of the distinct values (diag + upper)
    std::vector<std::vector<int>> cols(n);
    for (int k = 0; k < g; ++k) for (int j = 0; j < g; ++j) for (int i = 0; i < g; ++i) {
        const int e = (k * g + j) * g + i;
        if (k > 0) cols[e].push_back(e - g * g);
        if (j > 0) cols[e].push_back(e - g);
        if (i > 0) cols[e].push_back(e - 1);
        cols[e].push_back(e);
        if (i < g - 1) cols[e].push_back(e + 1);
        if (j < g - 1) cols[e].push_back(e + g);
        if (k < g - 1) cols[e].push_back(e + g * g);
    }
    std::vector<std::vector<int>> uid(n);
    for (int e = 0; e < n; ++e) uid[e].assign(cols[e].size(), -1);
    int nU = 0;
    for (int rank = 0; rank < 4; ++rank)
        for (int e = 0; e < n; ++e) {
            int fu = 0;
            while (cols[e][fu] < e) ++fu;
            if (fu + rank < (int)cols[e].size()) uid[e][fu + rank] = nU++;
        }
    for (int e = 0; e < n; ++e)
        for (size_t p = 0; p < cols[e].size(); ++p)
            if (cols[e][p] < e) {
                const int j = cols[e][p];
                const size_t q = std::find(cols[j].begin(), cols[j].end(), e) - cols[j].begin();
                uid[e][p] = uid[j][q];
            }
    auto pack = [&](int S, std::vector<int> &pc, std::vector<int> &pw) {
        const int nsl = (n + S - 1) / S;
        pc.assign((size_t)nsl * W * S, 0);
        pw.assign((size_t)nsl * W * S, 0);
        for (int e = 0; e < n; ++e)
            for (int k = 0; k < W; ++k) {
                const size_t idx = ((size_t)(e / S) * W + k) * S + e % S;
                const bool on = k < (int)cols[e].size();
                pc[idx] = on ? cols[e][k] : e;      // padding repeats the row itself with weight row nU (= 0)
                pw[idx] = on ? uid[e][k] : nU;
            }
    };
    std::vector<int> c16, w16, c8, w8;
    pack(16, c16, w16);
    pack(8, c8, w8);
    int *dc16, *dw16, *dc8, *dw8;
    cudaMalloc(&dc16, c16.size() * 4); cudaMalloc(&dw16, c16.size() * 4); cudaMalloc(&dc8, c8.size() * 4); cudaMalloc(&dw8, c8.size() * 4);
    cudaMemcpy(dc16, c16.data(), c16.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dw16, w16.data(), c16.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dc8, c8.data(), c8.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dw8, w8.data(), c8.size() * 4, cudaMemcpyHostToDevice);
    Mat A{n, W, dc16, dw16, dc8, dw8};
    Chunk c;
    long long off = 0;
    auto take = [&](long long rows) { long long o = off; off += rows * TW; return o; };
    c.x = take(n); c.y = take(n); c.r = take(n); c.d = take(n); c.dinv = take(n); c.V = take(nU + 1);
    off += 24 * (long long)n * TW;   // the rest of a tile's chunk (other vectors)
    c.stride = off;
    const int maxcta = 592;
    double *base;
    const size_t bytes = (size_t)maxcta * c.stride * 8;
    if (cudaMalloc(&base, bytes) != cudaSuccess) { printf("alloc of %zu MB failed\n", bytes >> 20); return 1; }
    cudaMemset(base, 0, bytes);
    printf("n %d nU %d chunk %.1f MB, bytes/op/tile %.0f KB\n", n, nU, c.stride * 8 / 1e6, (6.0 * n + nU) * 32 / 1e3);
    const double opbytes = (6.0 * n + nU) * 32;
    auto timeit = [&](const char *name, int ctas, auto launch) {
        const int reps = 200;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        launch(ctas, 4);
        cudaEventRecord(a);
        launch(ctas, reps);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("%-34s ctas %4d: %8.2f us/op  %7.0f GB/s  (%s)\n", name, ctas, ms * 1e3 / reps, opbytes * ctas * reps / ms / 1e6,
               cudaGetErrorString(cudaGetLastError()));
    };
    for (int ctas : {148, 296}) {
        timeit("v0 2 lanes/row, compiler order", ctas, [&](int n_, int r_) { k_v0<7><<<n_, 512>>>(A, base, c, r_); });
        cudaFuncSetAttribute(k_v0<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        timeit("v0 + 66 KB smem per CTA (L1 cut)", ctas, [&](int n_, int r_) { k_v0<7><<<n_, 512, 66 * 1024>>>(A, base, c, r_); });
        timeit("v0 + 100 KB smem per CTA", ctas, [&](int n_, int r_) { k_v0<7><<<n_, 512, 100 * 1024>>>(A, base, c, r_); });
        timeit("v1 2 lanes/row, batched, 64 regs", ctas, [&](int n_, int r_) { k_v1<7, 2><<<n_, 512>>>(A, base, c, r_); });
        timeit("v2 4 lanes/row, batched, 64 regs", ctas, [&](int n_, int r_) { k_v2<7, 2><<<n_, 512>>>(A, base, c, r_); });
        timeit("stream (same bytes, no gathers)", ctas, [&](int n_, int r_) { k_stream<<<n_, 512>>>(A, base, c, r_, nU); });
    }
    {
        const int nbig = 4 * n + 768;  // ~17152 rows for g = 16
        const long long big = c.V + (long long)(nU + 1) * TW;  // inside the "rest of the chunk"
        const double mixbytes = 0.5 * opbytes + 0.5 * 6.0 * nbig * 32;
        for (int pm : {0, 1})
            for (int ctas : {148, 296}) {
                const int reps = 200;
                cudaEvent_t a, b;
                cudaEventCreate(&a); cudaEventCreate(&b);
                k_mix<7><<<ctas, 512>>>(A, base, c, 4, big, nbig, pm);
                cudaEventRecord(a);
                k_mix<7><<<ctas, 512>>>(A, base, c, reps, big, nbig, pm);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                printf("mix apply/stream %s ctas %4d: %8.2f us/op  %7.0f GB/s  (%s)\n", pm ? "out of phase" : "in phase    ", ctas,
                       ms * 1e3 / reps, mixbytes * ctas * reps / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
    }
    for (int us : {30, 100}) {
        cudaMemcpyToSymbol(g_desync, &us, sizeof(int));
        timeit(us == 30 ? "v0, CTAs desynchronised up to 30 us" : "v0, CTAs desynchronised up to 100 us", 296,
               [&](int n_, int r_) { k_v0<7><<<n_, 512>>>(A, base, c, r_); });
    }
    { int z = 0; cudaMemcpyToSymbol(g_desync, &z, sizeof(int)); }
    {   // packed sliced ELL of the product (slice: [val w*16][col w*16][widx w*16]), width 7 everywhere
        const int nsl = (n + 15) / 16;
        std::vector<int> off(nsl + 1);
        for (int i = 0; i <= nsl; ++i) off[i] = i * W;
        std::vector<unsigned char> pk((size_t)nsl * W * 16 * 16, 0);
        for (int sl = 0; sl < nsl; ++sl) {
            unsigned char *b = pk.data() + (size_t)sl * W * 16 * 16;
            double *pv = (double *)b;
            int *pc = (int *)(b + W * 16 * 8), *pw = pc + W * 16;
            for (int k = 0; k < W; ++k)
                for (int r = 0; r < 16; ++r) {
                    const int e = sl * 16 + r;
                    if (e >= n) continue;
                    const bool on = k < (int)cols[e].size();
                    pv[k * 16 + r] = on ? 1.0 : 0.0;
                    pc[k * 16 + r] = on ? cols[e][k] : e;
                    pw[k * 16 + r] = on ? uid[e][k] : 0;
                }
        }
        int *doff; unsigned char *dpk;
        cudaMalloc(&doff, off.size() * 4); cudaMalloc(&dpk, pk.size());
        cudaMemcpy(doff, off.data(), off.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dpk, pk.data(), pk.size(), cudaMemcpyHostToDevice);
        pmc::Op o;
        memset(&o, 0, sizeof o);
        o.kind = pmc::OP_SPMM; o.n = n; o.flags = 0;
        o.rowptr = doff; o.pk = dpk;
        o.x.off = c.x; o.y.off = c.y; o.r.off = c.r; o.d.off = c.d; o.w.off = c.dinv; o.v.off = c.V;
        o.ca = 0.3; o.cb = 0.7;
        const size_t dyn = 16 * pmc::NSTAGE * sizeof(pmc::WarpStage);
        cudaFuncSetAttribute(k_product_op<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        cudaFuncSetAttribute(k_product_op<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
#define EXPRUN(NAME, ...)                                                                                      \
    cudaFuncSetAttribute(k_exp<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);                 \
    timeit(NAME, 296, [&](int n_, int r_) { k_exp<__VA_ARGS__><<<n_, 512, dyn>>>(o, base, c.stride, c.x, c.y, r_); });
        EXPRUN("exp: hoist, runtime w, val", true, 0, true)
        EXPRUN("exp: no hoist, runtime w, val", false, 0, true)
        EXPRUN("exp: hoist, W=7, val", true, 7, true)
        EXPRUN("exp: no hoist, W=7, val", false, 7, true)
        EXPRUN("exp: no hoist, W=7, no val", false, 7, false)
        EXPRUN("exp: hoist, W=7, no val", true, 7, false)
        for (int ctas : {148, 296}) {
            timeit("product op_spmm, staged (TMA)", ctas, [&](int n_, int r_) { k_product_op<true><<<n_, 512, dyn>>>(o, base, c.stride, c.x, c.y, r_); });
            timeit("product op_spmm, entries from L2", ctas, [&](int n_, int r_) { k_product_op<false><<<n_, 512, dyn>>>(o, base, c.stride, c.x, c.y, r_); });
        }
    }
    timeit("v1 2 lanes/row, batched, 128 regs", 148, [&](int n_, int r_) { k_v1<7, 1><<<n_, 512>>>(A, base, c, r_); });
    timeit("v2 4 lanes/row, batched, 128 regs", 148, [&](int n_, int r_) { k_v2<7, 1><<<n_, 512>>>(A, base, c, r_); });
    timeit("v2 4 lanes/row, 64 regs, 4 CTA/SM", 592, [&](int n_, int r_) { k_v2<7, 2><<<n_, 256>>>(A, base, c, r_); });
    return 0;
}
