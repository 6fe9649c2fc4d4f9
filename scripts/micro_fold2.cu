// micro-benchmark 2: the real phase_fold_large on synthetic long lists, 1 block x 8 warps (and 50 blocks)
#include <cstdio>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>
#include "gem_kernels.cuh"
using namespace gem;

__global__ void __launch_bounds__(256) bench(MapGeom g, MapLayers ml, Scratch sc, long long *cyc)
{
    __shared__ uint32_t s_key[8][FOLD_KMAX];
    const int w = threadIdx.x >> 5;
    long long t0 = clock64();
    phase_fold_large(g, ml, sc, true, true, s_key[w], blockIdx.x * 8 + w, gridDim.x * 8);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)cyc, (unsigned long long)(t1 - t0));
}

int main()
{
    const int L = 256, NC = L * L;
    for (int K : {16, 32, 72, 96, 128, 256, 512}) {
        for (int ncell : {8, 400}) {
            MapGeom g{}; g.L = L; g.res = 0.1f; g.rows = L; g.cols = L;
            MapLayers ml{}; Scratch sc{};
            cudaMalloc(&ml.ev, NC * 8); cudaMalloc(&ml.ci, NC * 8); cudaMalloc(&ml.lowest, NC * 4);
            cudaMalloc(&sc.cnt, NC * 4); cudaMalloc(&sc.tlarge, ncell * 16); cudaMalloc(&sc.ctr, sizeof(Counters));
            cudaMalloc(&sc.recA, (size_t)ncell * K * 16); cudaMalloc(&sc.recI, (size_t)ncell * K * 4);
            std::vector<float2> ev(NC, make_float2(0.1f, 0.01f));
            std::vector<float> low(NC, 100.f);
            cudaMemcpy(ml.ev, ev.data(), NC * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(ml.lowest, low.data(), NC * 4, cudaMemcpyHostToDevice);
            cudaMemset(ml.ci, 0, NC * 8);
            std::vector<int4> tl(ncell);
            std::vector<uint4> ra((size_t)ncell * K);
            std::vector<float> ri((size_t)ncell * K, 3.f);
            std::mt19937 rng(1);
            for (int c = 0; c < ncell; c++) {
                tl[c] = make_int4(c * 7 + 3, c * K, K, 0);
                std::vector<int> idx(K);
                std::iota(idx.begin(), idx.end(), c * K);
                std::shuffle(idx.begin(), idx.end(), rng);
                for (int e = 0; e < K; e++) {
                    float h = 0.1f + 0.01f * ((e * 37) % 11 - 5), v = 0.004f + 0.0001f * (e % 7);
                    uint4 r; r.x = idx[e]; r.y = *(uint32_t *)&h; r.z = *(uint32_t *)&v; r.w = 0x010203;
                    ra[(size_t)c * K + e] = r;
                }
            }
            cudaMemcpy(sc.tlarge, tl.data(), ncell * 16, cudaMemcpyHostToDevice);
            cudaMemcpy(sc.recA, ra.data(), ra.size() * 16, cudaMemcpyHostToDevice);
            cudaMemcpy(sc.recI, ri.data(), ri.size() * 4, cudaMemcpyHostToDevice);
            Counters h{}; h.nlarge = ncell;
            cudaMemcpy(sc.ctr, &h, sizeof h, cudaMemcpyHostToDevice);
            long long *cyc; cudaMallocManaged(&cyc, 8);
            for (int rep = 0; rep < 2; rep++) {
                *cyc = 0;
                bench<<<(ncell + 7) / 8, 256>>>(g, ml, sc, cyc);
                cudaDeviceSynchronize();
            }
            printf("K=%4d cells=%4d: slowest warp %8lld cycles = %.1f us (%.0f cycles/record)  err=%s\n", K, ncell, *cyc,
                   *cyc / 1965.0, (double)*cyc / K, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
