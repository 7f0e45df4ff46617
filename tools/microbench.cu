// Latency microbenchmarks on the target GPU (dependent-issue cycles per op, one warp unless stated).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, int iters, const int* chase, double2* gtab) {
    __shared__ double sm[1024];
    __shared__ int ich[1024];
    int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += blockDim.x) { sm[i] = 1.0 + i * 1e-9; ich[i] = chase[i]; }
    __syncthreads();
    double a = 1.0 + tid * 1e-9, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // 0: dependent DFMA
    t0 = clock64();
    for (int i = 0; i < iters; i++) a = fma(a, b, c);
    t1 = clock64(); if (tid == 0) cyc[0] = t1 - t0;
    // 1: dependent DADD
    t0 = clock64();
    for (int i = 0; i < iters; i++) a = a + c;
    t1 = clock64(); if (tid == 0) cyc[1] = t1 - t0;
    // 2: dependent DMUL
    t0 = clock64();
    for (int i = 0; i < iters; i++) a = a * b;
    t1 = clock64(); if (tid == 0) cyc[2] = t1 - t0;
    // 3: dependent FFMA (fp32) for comparison
    float fa = (float)a, fb = 1.0000001f, fc = 1e-9f;
    t0 = clock64();
    for (int i = 0; i < iters; i++) fa = fmaf(fa, fb, fc);
    t1 = clock64(); if (tid == 0) cyc[3] = t1 - t0;
    // 4: dependent double division
    t0 = clock64();
    for (int i = 0; i < iters; i++) a = 1.0000001 / a;
    t1 = clock64(); if (tid == 0) cyc[4] = t1 - t0;
    // 5: dependent LDS (pointer chase in smem)
    int j = tid & 1023;
    t0 = clock64();
    for (int i = 0; i < iters; i++) j = ich[j];
    t1 = clock64(); if (tid == 0) cyc[5] = t1 - t0;
    // 6: LDS.64 -> DFMA -> STS chain (same address per thread)
    t0 = clock64();
    for (int i = 0; i < iters; i++) { double v = sm[tid & 1023]; v = fma(v, b, c); sm[tid & 1023] = v; }
    t1 = clock64(); if (tid == 0) cyc[6] = t1 - t0;
    // 7: __syncthreads
    t0 = clock64();
    for (int i = 0; i < iters; i++) __syncthreads();
    t1 = clock64(); if (tid == 0) cyc[7] = t1 - t0;
    // 8: __syncwarp
    t0 = clock64();
    for (int i = 0; i < iters; i++) __syncwarp();
    t1 = clock64(); if (tid == 0) cyc[8] = t1 - t0;
    // 9: dependent global load (L2/L1 hit) pointer chase
    int g = tid & 1023;
    t0 = clock64();
    for (int i = 0; i < iters; i++) g = __ldg(chase + g);
    t1 = clock64(); if (tid == 0) cyc[9] = t1 - t0;
    // 10: 4 independent DFMA chains (ILP)
    double a1 = a + 1, a2 = a + 2, a3 = a + 3;
    t0 = clock64();
    for (int i = 0; i < iters; i++) { a = fma(a, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
    t1 = clock64(); if (tid == 0) cyc[10] = t1 - t0;
    // 11: exp()
    t0 = clock64();
    for (int i = 0; i < iters; i++) a = exp(a * 1e-3);
    t1 = clock64(); if (tid == 0) cyc[11] = t1 - t0;
    out[blockIdx.x * blockDim.x + tid] = a + fa + j + g + a1 + a2 + a3 + sm[tid & 1023];
}
int main() {
    const char* names[] = {"dep DFMA", "dep DADD", "dep DMUL", "dep FFMA", "dep DDIV", "dep LDS chase", "LDS->DFMA->STS",
                           "__syncthreads", "__syncwarp", "dep LDG chase (L1/L2 hit)", "4x indep DFMA (per iter)", "dep exp()"};
    int h[1024];
    for (int i = 0; i < 1024; i++) h[i] = (i * 37 + 11) & 1023;
    int* dch; double* out; long long* cyc; double2* gt;
    cudaMalloc(&dch, sizeof(h)); cudaMemcpy(dch, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&gt, 1 << 16);
    for (int nt : {32, 128, 256}) {
        const int iters = 2000;
        k_lat<<<1, nt>>>(out, cyc, iters, dch, gt);
        cudaDeviceSynchronize();
        k_lat<<<1, nt>>>(out, cyc, iters, dch, gt);
        cudaDeviceSynchronize();
        long long hc[16];
        cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        printf("--- block of %d threads, cycles per iteration ---\n", nt);
        for (int i = 0; i < 12; i++) printf("%-28s %8.1f\n", names[i], (double)hc[i] / iters);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
