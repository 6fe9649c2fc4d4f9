// micro-benchmark of the serial per-cell fold step (build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I gem_b200/csrc)
#include <cstdio>
#include "gem_kernels.cuh"
using namespace gem;

__global__ void bench(int K, int variant, float *out, long long *cyc, const float *hs, const float *vs)
{
    __shared__ float sh[1024], sv[1024];
    for (int i = threadIdx.x; i < K; i += 32) { sh[i] = hs[i]; sv[i] = vs[i]; }
    __syncwarp();
    CellState s;
    s.elev = 0.1f; s.var = 0.01f; s.inten = 0; s.rgb = 0; s.ci_dirty = false; s.minh = 0; s.minhv = 0; s.any = false;
    const unsigned lane = threadIdx.x;
    long long t0 = clock64();
    if (variant == 0) { // shuffles (chunks of 32) + fold_step, as phase_fold_large
        for (int c0 = 0; c0 < K; c0 += 32) {
            float h = sh[c0 + lane], v = sv[c0 + lane];
            for (int t = 0; t < 32; t++) {
                float hh = __shfl_sync(0xffffffffu, h, t), vv = __shfl_sync(0xffffffffu, v, t);
                fold_step(s, hh, vv, 0x1010101u, 1.0f, true);
            }
        }
    } else if (variant == 1) { // smem broadcast + fold_step
        for (int t = 0; t < K; t++) fold_step(s, sh[t], sv[t], 0x1010101u, 1.0f, true);
    } else if (variant == 2) { // only the Kalman arithmetic, literal reference form
        float e = s.elev, var = s.var;
        for (int t = 0; t < K; t++) {
            float h = sh[t], v = sv[t];
            float ne = (var * h + v * e) / (var + v);
            var = (v * var) / (v + var);
            e = ne;
        }
        s.elev = e; s.var = var;
    } else { // div2_rn form
        float e = s.elev, var = s.var;
        for (int t = 0; t < K; t++) {
            float h = sh[t], v = sv[t], qe, qv;
            div2_rn(var * h + v * e, v * var, var + v, qe, qv);
            e = qe; var = qv;
        }
        s.elev = e; s.var = var;
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = s.elev; out[1] = s.var; cyc[0] = t1 - t0; }
}

int main()
{
    const int K = 512;
    float *hs, *vs, *out; long long *cyc;
    cudaMallocManaged(&hs, K * 4); cudaMallocManaged(&vs, K * 4); cudaMallocManaged(&out, 8); cudaMallocManaged(&cyc, 8);
    for (int i = 0; i < K; i++) { hs[i] = 0.1f + 0.01f * ((i * 37) % 11 - 5); vs[i] = 0.004f + 0.0001f * (i % 7); }
    for (int variant = 0; variant < 4; variant++) {
        for (int rep = 0; rep < 2; rep++) {
            bench<<<1, 32>>>(K, variant, out, cyc, hs, vs);
            cudaDeviceSynchronize();
        }
        printf("variant %d: %.1f cycles/step  (e=%g var=%g)\n", variant, (double)cyc[0] / K, out[0], out[1]);
    }
    return 0;
}
