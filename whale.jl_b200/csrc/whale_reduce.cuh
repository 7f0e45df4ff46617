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

// ---------------------------------------------------------------------------------------------------------
// Mixture of WhaleModels (src/core.jl:66-76): ℓ = Σ_i logsumexp_j (ℓ_ij + log p_j − condition_j), with the
// gradient w.r.t. every component's raw parameters and w.r.t. log p_j through the responsibilities.
// ---------------------------------------------------------------------------------------------------------
// scatter one tangent plan's per-family outputs (component layout of the root) into dense [F][1+P] rows
__global__ void __launch_bounds__(256) k_mix_gather(const double* __restrict__ out_fam, int F, int KR,
                                                    const int* __restrict__ act_root, int first,
                                                    const double* __restrict__ cond, double* __restrict__ dense,
                                                    double* __restrict__ condv, int P) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (long long)F * KR) {
        const int f = (int)(idx / KR), k = (int)(idx - (long long)f * KR);
        if (k == 0) { if (first) dense[(size_t)f * (1 + P)] = out_fam[idx]; }
        else dense[(size_t)f * (1 + P) + 1 + act_root[k]] = out_fam[idx];
    }
    if (idx < KR) {
        const int k = (int)idx;
        if (k == 0) { if (first) condv[0] = cond[0]; }
        else condv[1 + act_root[k]] = cond[k];
    }
}

// per family: lse_i and the responsibilities r_ij = exp(M_ij − lse_i)
__global__ void __launch_bounds__(256) k_mix_resp(const double* __restrict__ dense, const double* __restrict__ condv,
                                                  const double* __restrict__ logw, int F, int P, int J,
                                                  double* __restrict__ lse, double* __restrict__ resp) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const size_t comp = (size_t)F * (1 + P);
    double mx = -dinf();
    for (int j = 0; j < J; j++) {
        const double v = dense[j * comp + (size_t)f * (1 + P)] + logw[j] - condv[(size_t)j * (1 + P)];
        mx = v > mx ? v : mx;
    }
    if (!(mx > -dinf())) {  // every component gives L <= 0 for this family
        lse[f] = -dinf();
        for (int j = 0; j < J; j++) resp[(size_t)f * J + j] = 0.0;
        return;
    }
    double s = 0.0;
    for (int j = 0; j < J; j++) {
        const double v = dense[j * comp + (size_t)f * (1 + P)] + logw[j] - condv[(size_t)j * (1 + P)];
        s += exp(v - mx);
    }
    const double l = mx + log(s);
    lse[f] = l;
    for (int j = 0; j < J; j++) {
        const double v = dense[j * comp + (size_t)f * (1 + P)] + logw[j] - condv[(size_t)j * (1 + P)];
        resp[(size_t)f * J + j] = exp(v - l);
    }
}

// one block per output: q = 0 the total, q = 1 + j*(1+P) + 0 the gradient w.r.t. log p_j, + 1 + p w.r.t. x_jp;
// fixed-order strided sums + tree (deterministic bits)
__global__ void __launch_bounds__(256) k_mix_reduce(const double* __restrict__ dense, const double* __restrict__ condv,
                                                    const double* __restrict__ lse, const double* __restrict__ resp,
                                                    int F, int P, int J, double* __restrict__ out) {
    __shared__ double sh[256];
    const int q = blockIdx.x;
    const size_t comp = (size_t)F * (1 + P);
    double s = 0.0;
    if (q == 0) {
        for (int f = threadIdx.x; f < F; f += 256) s += lse[f];
    } else {
        const int j = (q - 1) / (1 + P), c = (q - 1) - j * (1 + P);
        if (c == 0) {
            for (int f = threadIdx.x; f < F; f += 256) s += resp[(size_t)f * J + j];
        } else {
            const double cv = condv[(size_t)j * (1 + P) + c];
            for (int f = threadIdx.x; f < F; f += 256) {
                const double r = resp[(size_t)f * J + j];
                if (r != 0.0) s += r * (dense[j * comp + (size_t)f * (1 + P) + c] - cv);
            }
        }
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[q] = sh[0];
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

// the exchange as its own launch (evaluations in several passes, whale_peer_sum_async); see peer_exchange in whale_common.cuh
__global__ void __launch_bounds__(128) k_peer_sum(PeerDev* PD, double* out) {
    EXTERN_SHARED(psm);
    peer_exchange<128>(PD, out, reinterpret_cast<unsigned*>(psm));
}
