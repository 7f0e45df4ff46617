// K3 — deterministic reduction over families + conditioning (src/core.jl:54,63 ; src/condition.jl), and
// the DFMA microbenchmark that provides the fp64 roofline denominator.
#pragma once
#include "whale_common.cuh"

__global__ void __launch_bounds__(256) k_reduce1(const double* __restrict__ out_fam, int F, int K, int chunk,
                                                 double* __restrict__ partial) {
    __shared__ double sh[256];
    const int b = blockIdx.x, f0 = b * chunk, f1 = min(F, f0 + chunk);
    for (int k = 0; k < K; k++) {
        double s = 0.0;
        for (int f = f0 + threadIdx.x; f < f1; f += 256) s += out_fam[(size_t)f * K + k];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[(size_t)b * K + k] = sh[0];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_reduce2(const double* __restrict__ partial, int nb, int K, int F,
                                                 int cond_kind, PlanDev PL, int root, int first,
                                                 double* __restrict__ out) {
    __shared__ double tot[256];
    __shared__ int finite;
    for (int k = threadIdx.x; k < K; k += 256) {
        double s = 0.0;
        for (int b = 0; b < nb; b++) s += partial[(size_t)b * K + k];
        s -= (double)F * PL.cond[cond_kind * PL.Kmax + k];
        tot[k] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) finite = isfinite(tot[0]) ? 1 : 0;  // ℓhood src/core.jl:15
    __syncthreads();
    // `out` was zeroed by the host; every gradient pass (parameter chunk) writes its own parameters, the first
    // pass also the log-likelihood
    for (int k = threadIdx.x; k < K; k += 256) {
        if (k == 0) { if (first) out[0] = finite ? tot[0] : -dinf(); }
        else out[1 + PL.act[root * PL.Kmax + k]] = finite ? tot[k] : 0.0;
    }
}

// dependent-free DFMA microbenchmark (fp64 roofline denominator)
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
