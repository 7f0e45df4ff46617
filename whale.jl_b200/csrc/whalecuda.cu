// libwhalecuda — B200 (sm_100a) engine for Whale.jl's ALE/DLWGD likelihood + forward-mode gradient.
//
// Hot path replaced (reference paths relative to the reference checkout):
//   slice tables   src/model.jl:162-191, src/bdputil.jl:6-11          -> k_tables
//   the DP         src/core.jl:83-199 (whale!/whalewgd!/whaleroot!)   -> k_dp
//   Σ − N·cond     src/core.jl:46-64, src/condition.jl:11-29          -> k_reduce1/2
// Forward tangents replace ForwardDiff duals: every ℓ cell carries K_e = 1 + (#parameters that can
// influence branch e) components; lanes span (clade cell × component).
//
// Data layout in HBM (built once by whale_data_create, the read_ale-time packer):
//   FamHdr[F]   : per family {arena byte offset, Γ, #root levels, ℓ offset, work}
//   arena blob  : per family  NodeRec[n_nodes] | per-node u32 CSR pointers | 16-byte triple entries
//                 {u16 i1, u16 i2, f64 p} with indices already resolved to the *local* cell index of the
//                 branch they are used in (the reference resolves index[γ,e] at run time, src/ccd.jl:41-49)
// On chip: one CTA per family; the last row of every branch (C_e × K_e doubles) lives in shared memory,
// slices ping-pong between that row and a scratch row, one barrier per slice.  No tensor cores: the DP
// is an irregular gather–multiply–accumulate in fp64.
//
// There is NO CPU fallback in this file: every entry point that computes needs a CUDA device.
#ifdef WHALE_EMU
// Test-only build: tests/emu/cuda_emu.h maps the CUDA constructs used here onto host threads so the
// kernel logic can be exercised on a machine without a GPU.  Never shipped, never loaded by the package.
#include "cuda_emu.h"
#define LAUNCH(kern, grid, block, smem, st, ...) emu::launch(grid, block, smem, [=]() { kern(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define LAUNCH(kern, grid, block, smem, st, ...) kern<<<grid, block, smem, st>>>(__VA_ARGS__)
#endif

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/whalecuda.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int32_t fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(WHALE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),    \
                        __FILE__, __LINE__);                                                       \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device-side structures
// ------------------------------------------------------------------------------------------------
struct __align__(16) Ent {  // one resolved clade-split term: p * X[i1] * Y[i2]
    uint16_t i1, i2;
    uint32_t pad;
    double p;
};
static_assert(sizeof(Ent) == 16, "Ent must be 16 bytes");

#ifndef WHALE_EMU
#define EXTERN_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#endif

struct NodeRec {        // per family, per species-tree node (all offsets relative to the family blob)
    uint32_t C;         // compatible clades (columns of ℓ[e])
    uint32_t nonleaf;   // how many of them are non-leaf clades (they come last: clades are size-sorted)
    uint32_t dptr_off;  // u32-word offset of dptr[C+1]  (within-branch / WGD / Πroot terms)
    uint32_t dent_off;  // 16-byte-entry offset of the entries dptr points into
    uint32_t tptr_off;  // u32-word offset of tptr[C+1]  (speciation terms at row 1; internal/root only)
    uint32_t tent_off;  // 16-byte-entry offset of the speciation entries
    uint32_t loss_off;  // u32-word offset of lossF[C], lossG[C] (int32 local index in child or -1)
    uint32_t lev_off;   // root only: u32-word offset of level pointers lev[nlev+1]
};

struct FamHdr {
    uint64_t base;     // byte offset of the blob in the arena
    uint64_t ell_off;  // offset (doubles) of this family's ℓ in the keep_ell buffer
    uint32_t G;        // clades
    uint32_t nlev;     // root levels (distinct clade sizes)
    uint32_t sumC;     // Σ_e C_e
    uint32_t maxC;     // max_e C_e
};

struct ModelDev {  // structure arrays (device pointers), node index = id-1
    int nn;
    const int* order;
    const int* child0;
    const int* child1;
    const int* kind;
    const int* nsl;
    const double* dt;
    const double* leafP;
    const int* lam_slot;
    const int* mu_slot;
    const int* q_slot;
    int eta_slot;
    int log_scale;
    int root;
    // tables-kernel schedule: nodes grouped by height
    int nlvl;
    const int* lvl_off;
    const int* lvl_nodes;
};

struct PlanDev {  // tangent plan: which raw parameters each branch carries
    int Kmax;
    const int* K;          // [nn]
    const int* act;        // [nn*Kmax] global parameter id of component k (act[e][0] = -1)
    const int16_t* cmap;   // [nn*2*Kmax] component index in child j of parent's component k, or -1
    const uint8_t* role;   // [nn*Kmax] bit0 own λ, bit1 own μ, bit2 own q, bit3 η
    const int* toff;       // [nn] offset (doubles) of node e's table: (n_e+1) rows × K_e
    // tables written by k_tables
    double* eps;           // [tab_len]           ϵ rows (component-major within a row)
    double2* pp;           // [tab_len]           (ϕ, ψ) rows
    double* cx;            // [nn*Kmax]  row-1 coefficient X (WGD: 1−q+2qϵ_f ; root: (1−η)ξ/η)
    double* cy;            // [nn*Kmax]  row-1 coefficient Y (WGD: q ; root: η(1−ϵ)/ξ²)
    double* leaf;          // [nn*Kmax]  last-row value of a leaf clade on leaf branch e
    double* cond;          // [3*Kmax]   condition() per kind, components of the root
};

// ------------------------------------------------------------------------------------------------
// one-partial dual number: each lane carries the value and ITS tangent component
// ------------------------------------------------------------------------------------------------
struct D1 {
    double v, d;
};
__device__ __forceinline__ D1 mk(double v, double d = 0.0) { return D1{v, d}; }
__device__ __forceinline__ D1 operator+(D1 a, D1 b) { return D1{a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ D1 operator-(D1 a, D1 b) { return D1{a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ D1 operator*(D1 a, D1 b) { return D1{a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ D1 operator/(D1 a, D1 b) {
    double q = a.v / b.v;
    return D1{q, (a.d - q * b.d) / b.v};
}
__device__ __forceinline__ D1 operator+(double a, D1 b) { return D1{a + b.v, b.d}; }
__device__ __forceinline__ D1 operator-(double a, D1 b) { return D1{a - b.v, -b.d}; }
__device__ __forceinline__ D1 operator-(D1 a, double b) { return D1{a.v - b, a.d}; }
__device__ __forceinline__ D1 operator*(double a, D1 b) { return D1{a * b.v, a * b.d}; }
__device__ __forceinline__ D1 dexp(D1 a) {
    double e = exp(a.v);
    return D1{e, e * a.d};
}
__device__ __forceinline__ D1 dlog(D1 a) { return D1{log(a.v), a.d / a.v}; }

// ------------------------------------------------------------------------------------------------
// K1: slice tables (ϵ, ϕ, ψ) with tangents.  src/model.jl:162-191, src/bdputil.jl:6-11.
// One CTA; nodes of equal height are independent -> one warp per node, lanes over components.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ D1 child_eps_last(const ModelDev& M, const PlanDev& PL, int e, int j, int child,
                                              int k) {
    int Kc = PL.K[child];
    const double* row = PL.eps + PL.toff[child] + (size_t)M.nsl[child] * Kc;
    int kc = k == 0 ? 0 : PL.cmap[(e * 2 + j) * PL.Kmax + k];
    return mk(row[0], (k == 0 || kc < 0) ? 0.0 : row[kc]);
}

__global__ void __launch_bounds__(1024) k_tables(ModelDev M, PlanDev PL, const double* __restrict__ x,
                                                 const double* __restrict__ pleaf) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    for (int L = 0; L < M.nlvl; L++) {
        int n0 = M.lvl_off[L], n1 = M.lvl_off[L + 1];
        for (int j = n0 + warp; j < n1; j += nwarp) {
            const int e = M.lvl_nodes[j];
            const int K = PL.K[e], kind = M.kind[e], n = M.nsl[e];
            for (int k = lane; k < K; k += 32) {
                const unsigned role = k == 0 ? 0u : PL.role[e * PL.Kmax + k];
                // getθ (src/rmodels.jl:31-33,55-64): raw -> rate, with the chain factor for the log scale
                D1 lam, mu;
                {
                    int ls = M.lam_slot[e], ms = M.mu_slot[e];
                    double lv = ls < 0 ? NaN : (M.log_scale ? exp(x[ls]) : x[ls]);
                    double mv = ms < 0 ? NaN : (M.log_scale ? exp(x[ms]) : x[ms]);
                    lam = mk(lv, (role & 1u) ? (M.log_scale ? lv : 1.0) : 0.0);
                    mu = mk(mv, (role & 2u) ? (M.log_scale ? mv : 1.0) : 0.0);
                }
                D1 ep;
                D1 q = mk(0.0);
                if (kind == WHALE_LEAF) {  // setnode! src/model.jl:170
                    ep = mk(pleaf ? pleaf[e] : 0.0);
                } else if (kind == WHALE_WGD) {  // setwgdnode! src/model.jl:175-180
                    q = mk(x[M.q_slot[e]], (role & 4u) ? 1.0 : 0.0);
                    D1 ec = child_eps_last(M, PL, e, 0, M.child0[e], k);
                    ep = q * (ec * ec) + (1.0 - q) * ec;
                    D1 w = (1.0 - q) + 2.0 * (q * ec);  // Πwgdloss coefficient src/core.jl:198
                    PL.cx[e * PL.Kmax + k] = k == 0 ? w.v : w.d;
                    PL.cy[e * PL.Kmax + k] = k == 0 ? q.v : q.d;
                } else {  // internal / root: product of the children's last ϵ
                    D1 ef = child_eps_last(M, PL, e, 0, M.child0[e], k);
                    D1 eg = child_eps_last(M, PL, e, 1, M.child1[e], k);
                    ep = ef * eg;
                    if (kind == WHALE_ROOT) {  // whaleroot! src/core.jl:131-147 ; condition src/condition.jl
                        D1 eta = mk(x[M.eta_slot], (role & 8u) ? 1.0 : 0.0);
                        D1 xi = 1.0 - (1.0 - eta) * ep;
                        D1 A = (1.0 - eta) * xi / eta;
                        D1 B = eta * (1.0 - ep) / (xi * xi);
                        PL.cx[e * PL.Kmax + k] = k == 0 ? A.v : A.d;
                        PL.cy[e * PL.Kmax + k] = k == 0 ? B.v : B.d;
                        // geompgf(η, s) = ηs/(1−(1−η)s)  src/bdputil.jl:67
                        D1 gr = eta * ep / (1.0 - (1.0 - eta) * ep);
                        D1 gf = eta * ef / (1.0 - (1.0 - eta) * ef);
                        D1 gg = eta * eg / (1.0 - (1.0 - eta) * eg);
                        D1 pr = ((1.0 - gf) - gg) + gr;          // RootCondition :21-29
                        D1 pn = 1.0 - gr;                        // NonExtinctCondition :15-18
                        double inf = __longlong_as_double(0x7ff0000000000000LL);
                        D1 cr = pr.v > 0.0 ? dlog(pr) : mk(-inf, 0.0);
                        D1 cn = dlog(pn);
                        PL.cond[0 * PL.Kmax + k] = 0.0;
                        PL.cond[1 * PL.Kmax + k] = k == 0 ? cr.v : cr.d;
                        PL.cond[2 * PL.Kmax + k] = k == 0 ? cn.v : cn.d;
                    }
                }
                double* erow = PL.eps + PL.toff[e];
                double2* prow = PL.pp + PL.toff[e];
                erow[k] = k == 0 ? ep.v : ep.d;
                prow[k] = make_double2(k == 0 ? 1.0 : 0.0, k == 0 ? 1.0 : 0.0);
                if (n > 0) {
                    // getα src/bdputil.jl:6-7 (critical branch decided on VALUES, like isapprox on Duals)
                    const double t = M.dt[e];
                    D1 a;
                    if (fabs(lam.v - mu.v) <= 1e-6) {
                        a = (lam * mk(t)) / (1.0 + lam * mk(t));
                    } else {
                        D1 ex = dexp(mk(t) * (lam - mu));
                        a = mu * (ex - 1.0) / (lam * ex - mu);
                    }
                    D1 b = (lam / mu) * a;
                    D1 oma = 1.0 - a, omb = 1.0 - b;
                    D1 g = oma * omb;
                    D1 lf = mk(M.leafP[e]);  // leaf clade on a leaf branch: ℓ_i = ϕ_i ℓ_{i−1} (src/core.jl:94,123)
                    for (int i = 1; i <= n; i++) {  // setslices! src/model.jl:182-191
                        D1 den = 1.0 - b * ep;
                        D1 inv = mk(1.0) / den;
                        D1 phi = g * (inv * inv);
                        D1 psi = (g * b) * (inv * inv * inv);
                        ep = (a + (oma - b) * ep) * inv;
                        erow[(size_t)i * K + k] = k == 0 ? ep.v : ep.d;
                        prow[(size_t)i * K + k] = make_double2(k == 0 ? phi.v : phi.d, k == 0 ? psi.v : psi.d);
                        lf = phi * lf;
                    }
                    if (kind == WHALE_LEAF) PL.leaf[e * PL.Kmax + k] = k == 0 ? lf.v : lf.d;
                } else if (kind == WHALE_LEAF) {
                    PL.leaf[e * PL.Kmax + k] = k == 0 ? M.leafP[e] : 0.0;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K2: the DP.  One CTA per family, NT threads; lanes span (cell, component).
// ------------------------------------------------------------------------------------------------
struct DPArgs {
    ModelDev M;
    PlanDev PL;
    const unsigned char* arena;
    const FamHdr* hdr;
    const int* perm;   // launch order (decreasing work)
    double* out_fam;   // [F * Kroot]  (log L_f, ∂ log L_f / ∂ component)
    double* ell;       // keep_ell buffer or nullptr
    int nfam;
    int skip_leaf;     // share family-independent leaf-branch rows (off in keep_ell mode)
};

// Σ_t p_t X[i1] Y[i2] with the product rule for the lane's component; m = 0 for the value lane.
__device__ __forceinline__ void pairsum(const Ent* __restrict__ ents, uint32_t tb, uint32_t te,
                                        const double* __restrict__ X, int KX, int kx,
                                        const double* __restrict__ Y, int KY, int ky, double m, double& S0,
                                        double& Sk) {
    double s0 = 0.0, sk = 0.0;
    for (uint32_t t = tb; t < te; t++) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(ents + t));
        const double p = __hiloint2double((int)raw.w, (int)raw.z);
        const double* xp = X + (size_t)(raw.x & 0xffffu) * KX;
        const double* yp = Y + (size_t)(raw.x >> 16) * KY;
        double x0 = xp[0], y0 = yp[0];
        double xk = kx >= 0 ? xp[kx] : 0.0;
        double yk = ky >= 0 ? yp[ky] : 0.0;
        double px = p * x0;
        s0 = fma(px, y0, s0);
        sk = fma(px, yk, sk);
        sk = fma(m * (p * y0), xk, sk);
    }
    S0 = s0;
    Sk = sk;
}

template <int NT>
__global__ void __launch_bounds__(NT) k_dp(DPArgs A) {
    EXTERN_SHARED(smem_raw);
    const int tid = threadIdx.x;
    const ModelDev& M = A.M;
    const PlanDev& PL = A.PL;
    const int nn = M.nn, Kmax = PL.Kmax;
    const int fam = A.perm[blockIdx.x];
    const FamHdr H = A.hdr[fam];
    const unsigned char* blob = A.arena + H.base;
    const NodeRec* nrec = reinterpret_cast<const NodeRec*>(blob);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(blob);
    const Ent* ents = reinterpret_cast<const Ent*>(blob);

    int* s_roff = reinterpret_cast<int*>(smem_raw);          // [nn+1] row offsets (doubles)
    double* rows = reinterpret_cast<double*>(smem_raw + (((nn + 1) * sizeof(int) + 15) & ~size_t(15)));
    if (tid == 0) {
        int o = 0;
        for (int e = 0; e < nn; e++) {
            s_roff[e] = o;
            o += (int)nrec[e].C * PL.K[e];
        }
        s_roff[nn] = o;
    }
    __syncthreads();
    double* scr = rows + s_roff[nn];

    for (int oi = 0; oi < nn; oi++) {
        const int e = M.order[oi];
        const NodeRec R = nrec[e];
        const int C = (int)R.C;
        if (C == 0) continue;
        const int kind = M.kind[e], K = PL.K[e], n = M.nsl[e];
        double* fin = rows + s_roff[e];
        double* ellp = A.ell ? A.ell + H.ell_off : nullptr;
        size_t ell_e = 0;
        if (ellp) {  // offset of node e's matrix in the family's ℓ (node-index order)
            for (int e2 = 0; e2 < e; e2++) ell_e += (size_t)(M.nsl[e2] + 1) * nrec[e2].C;
            ellp += ell_e;
        }
        // lane -> (cell group, component): groups of K lanes, GP groups per pass
        const int GP = NT / K;            // K <= NT is guaranteed by the host
        const int grp = tid / K, k = tid - grp * K;
        const bool lane_on = grp < GP;
        const double m = k == 0 ? 0.0 : 1.0;

        if (kind == WHALE_LEAF && R.nonleaf == 0 && A.skip_leaf) {
            // every compatible clade is a leaf clade: the last row is family-independent (k_tables)
            if (lane_on)
                for (int c = grp; c < C; c += GP) fin[c * K + k] = PL.leaf[e * Kmax + k];
            __syncthreads();
            continue;
        }

        if (kind == WHALE_ROOT) {
            // whaleroot! src/core.jl:130-149: clades ascending in size, level-synchronous
            const int f = M.child0[e], g = M.child1[e];
            const int KF = PL.K[f], KG = PL.K[g];
            const double* finF = rows + s_roff[f];
            const double* finG = rows + s_roff[g];
            const int kf = k == 0 ? 0 : PL.cmap[(e * 2 + 0) * Kmax + k];
            const int kg = k == 0 ? 0 : PL.cmap[(e * 2 + 1) * Kmax + k];
            const double* epsF = PL.eps + PL.toff[f] + (size_t)M.nsl[f] * KF;
            const double* epsG = PL.eps + PL.toff[g] + (size_t)M.nsl[g] * KG;
            const double ef0 = epsF[0], eg0 = epsG[0];
            const double efk = (k > 0 && kf >= 0) ? epsF[kf] : 0.0;
            const double egk = (k > 0 && kg >= 0) ? epsG[kg] : 0.0;
            const double cx0 = PL.cx[e * Kmax], cy0 = PL.cy[e * Kmax];
            const double cxk = PL.cx[e * Kmax + k], cyk = PL.cy[e * Kmax + k];
            const uint32_t* dptr = words + R.dptr_off;
            const uint32_t* tptr = words + R.tptr_off;
            const int32_t* lossF = reinterpret_cast<const int32_t*>(words + R.loss_off);
            const int32_t* lossG = lossF + C;
            const uint32_t* lev = words + R.lev_off;
            for (uint32_t L = 0; L < H.nlev; L++) {
                const int c0 = (int)lev[L], c1 = (int)lev[L + 1];
                if (lane_on)
                    for (int c = c0 + grp; c < c1; c += GP) {
                        double a0, ak, b0, bk;
                        pairsum(ents + R.dent_off, dptr[c], dptr[c + 1], fin, K, k, fin, K, k, m, a0, ak);
                        pairsum(ents + R.tent_off, tptr[c], tptr[c + 1], finF, KF, kf, finG, KG, kg, m, b0, bk);
                        const int lf = lossF[c], lg = lossG[c];
                        double f0 = 0.0, fk = 0.0, g0 = 0.0, gk = 0.0;
                        if (lf >= 0) { f0 = finF[lf * KF]; fk = kf >= 0 ? finF[lf * KF + kf] : 0.0; }
                        if (lg >= 0) { g0 = finG[lg * KG]; gk = kg >= 0 ? finG[lg * KG + kg] : 0.0; }
                        // Πloss src/core.jl:172-176
                        double c0v = f0 * eg0 + g0 * ef0;
                        double ckv = fk * eg0 + gk * ef0 + m * (f0 * egk + g0 * efk);
                        double u0 = b0 + c0v, uk = bk + ckv;
                        double r = cx0 * ak + cy0 * uk + m * (cxk * a0 + cyk * u0);
                        fin[c * K + k] = r;
                        if (ellp && k == 0) ellp[c] = r;
                    }
                __syncthreads();
            }
            // log L and its gradient  (src/core.jl:35-36)
            if (tid < K) {
                const double Lv = fin[(C - 1) * K];
                double o;
                if (Lv > 0.0) o = tid == 0 ? log(Lv) : fin[(C - 1) * K + tid] / Lv;
                else o = tid == 0 ? -__longlong_as_double(0x7ff0000000000000LL) : 0.0;
                A.out_fam[(size_t)fam * K + tid] = o;
            }
            continue;
        }

        // ---- row 1 of a non-root branch ----
        double* cur = (n & 1) ? scr : fin;  // row i lives in fin iff (n − i) is even
        if (kind == WHALE_LEAF) {           // src/core.jl:93-94
            const int nleafc = C - (int)R.nonleaf;
            if (lane_on)
                for (int c = grp; c < C; c += GP) {
                    double v = (c < nleafc && k == 0) ? M.leafP[e] : 0.0;
                    cur[c * K + k] = v;
                    if (ellp && k == 0) ellp[c] = v;
                }
        } else if (kind == WHALE_INTERNAL) {  // Πspeciation + Πloss, src/core.jl:95-98,160-176
            const int f = M.child0[e], g = M.child1[e];
            const int KF = PL.K[f], KG = PL.K[g];
            const double* finF = rows + s_roff[f];
            const double* finG = rows + s_roff[g];
            const int kf = k == 0 ? 0 : PL.cmap[(e * 2 + 0) * Kmax + k];
            const int kg = k == 0 ? 0 : PL.cmap[(e * 2 + 1) * Kmax + k];
            const double* epsF = PL.eps + PL.toff[f] + (size_t)M.nsl[f] * KF;
            const double* epsG = PL.eps + PL.toff[g] + (size_t)M.nsl[g] * KG;
            const double ef0 = epsF[0], eg0 = epsG[0];
            const double efk = (k > 0 && kf >= 0) ? epsF[kf] : 0.0;
            const double egk = (k > 0 && kg >= 0) ? epsG[kg] : 0.0;
            const uint32_t* tptr = words + R.tptr_off;
            const int32_t* lossF = reinterpret_cast<const int32_t*>(words + R.loss_off);
            const int32_t* lossG = lossF + C;
            if (lane_on)
                for (int c = grp; c < C; c += GP) {
                    double b0, bk;
                    pairsum(ents + R.tent_off, tptr[c], tptr[c + 1], finF, KF, kf, finG, KG, kg, m, b0, bk);
                    const int lf = lossF[c], lg = lossG[c];
                    double f0 = 0.0, fk = 0.0, g0 = 0.0, gk = 0.0;
                    if (lf >= 0) { f0 = finF[lf * KF]; fk = kf >= 0 ? finF[lf * KF + kf] : 0.0; }
                    if (lg >= 0) { g0 = finG[lg * KG]; gk = kg >= 0 ? finG[lg * KG + kg] : 0.0; }
                    double c0v = f0 * eg0 + g0 * ef0;
                    double ckv = fk * eg0 + gk * ef0 + m * (f0 * egk + g0 * efk);
                    double r = k == 0 ? b0 + c0v : bk + ckv;
                    cur[c * K + k] = r;
                    if (ellp && k == 0) ellp[c] = r;
                }
        } else {  // WGD: q·Σ p ℓ_f[γ1]ℓ_f[γ2] + (1−q+2qϵ_f)·ℓ_f[γ]   src/core.jl:103-119,187-199
            const int f = M.child0[e];
            const int KF = PL.K[f];
            const double* finF = rows + s_roff[f];
            const int kf = k == 0 ? 0 : PL.cmap[(e * 2 + 0) * Kmax + k];
            const double cx0 = PL.cx[e * Kmax], cy0 = PL.cy[e * Kmax];
            const double cxk = PL.cx[e * Kmax + k], cyk = PL.cy[e * Kmax + k];
            const uint32_t* dptr = words + R.dptr_off;
            if (lane_on)
                for (int c = grp; c < C; c += GP) {
                    double s0, sk;
                    pairsum(ents + R.dent_off, dptr[c], dptr[c + 1], finF, KF, kf, finF, KF, kf, m, s0, sk);
                    double u0 = finF[c * KF];
                    double uk = kf >= 0 ? finF[c * KF + kf] : 0.0;
                    double r = cy0 * sk + cx0 * uk + m * (cyk * s0 + cxk * u0);
                    cur[c * K + k] = r;
                    if (ellp && k == 0) ellp[c] = r;
                }
        }
        __syncthreads();

        // ---- slices: ℓ_i = ϕ_i ℓ_{i−1} + ψ_i Σ_t p ℓ_{i−1}[γ1] ℓ_{i−1}[γ2]   src/core.jl:121-128,178-185 ----
        const uint32_t* dptr = words + R.dptr_off;
        const Ent* dents = ents + R.dent_off;
        const double2* pprow = PL.pp + PL.toff[e];
        // the lane's first cell keeps its triple range in registers across all slices
        uint32_t tb0 = 0, te0 = 0;
        if (lane_on && grp < C) { tb0 = dptr[grp]; te0 = dptr[grp + 1]; }
        for (int i = 1; i <= n; i++) {
            const double* src = cur;
            double* dst = (cur == fin) ? scr : fin;
            if (lane_on && grp < C) {
                const double2 c0 = pprow[(size_t)i * K];
                const double2 ck = pprow[(size_t)i * K + k];
                for (int c = grp; c < C; c += GP) {
                    uint32_t tb = tb0, te = te0;
                    if (c != grp) { tb = dptr[c]; te = dptr[c + 1]; }
                    double s0, sk;
                    pairsum(dents, tb, te, src, K, k, src, K, k, m, s0, sk);
                    double o0 = src[c * K], ok = src[c * K + k];
                    double r = c0.x * ok + c0.y * sk + m * (ck.x * o0 + ck.y * s0);
                    dst[c * K + k] = r;
                    if (ellp && k == 0) ellp[(size_t)i * C + c] = r;
                }
            }
            cur = dst;
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3: deterministic reduction over families + conditioning.  src/core.jl:54,63 ; src/condition.jl
// ------------------------------------------------------------------------------------------------
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
            if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[(size_t)b * K + k] = sh[0];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_reduce2(const double* __restrict__ partial, int nb, int K, int F,
                                                 int cond_kind, PlanDev PL, int root, int P,
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
    for (int i = threadIdx.x; i <= P; i += 256) out[i] = 0.0;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) {
        if (k == 0) out[0] = finite ? tot[0] : -__longlong_as_double(0x7ff0000000000000LL);
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <class T>
static cudaError_t upload(const std::vector<T>& v, T** d) {
    *d = nullptr;
    size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)d, n);
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

struct Plan {
    int Kmax = 1;
    std::vector<int> K, act, toff;
    std::vector<int16_t> cmap;
    std::vector<uint8_t> role;
    size_t tab_len = 0;
    PlanDev dev{};
    std::vector<void*> owned;
};

struct whale_model {
    int device = 0;
    int nn = 0, P = 0, root = -1;
    std::vector<int> order, child0, child1, kind, nsl, lam_slot, mu_slot, q_slot;
    std::vector<double> dt, leafP;
    int eta_slot = 0, log_scale = 0;
    std::vector<int> lvl_off, lvl_nodes;
    ModelDev dev{};
    std::vector<void*> owned;
    Plan plan[2];  // 0: value only, 1: all raw parameters
    double* d_x = nullptr;      // staging for host-pointer calls
    double* d_pleaf = nullptr;  // [nn]
    double* d_out = nullptr;    // [1+P]
    double* h_pin = nullptr;    // pinned staging (x in, out back)
    cudaStream_t stream = nullptr;
};

struct whale_data {
    whale_model* m = nullptr;
    int F = 0;
    std::vector<FamHdr> hdr;
    std::vector<int> perm;
    std::vector<unsigned char> arena_host;  // packing buffer (released after upload)
    size_t arena_bytes = 0;
    unsigned char* d_arena = nullptr;
    FamHdr* d_hdr = nullptr;
    int* d_perm = nullptr;
    double* d_out_fam = nullptr;  // [F*Kmax(plan1)]
    double* d_partial = nullptr;
    double* d_ell = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    uint64_t ell_total = 0;
    bool ell_valid = false;
    int maxSumCK[2] = {0, 0};   // max over families of Σ_e C_e K_e + max_e C_e K_e, per plan (doubles)
    // aggregated work counters per node (over families): Σ C_e, Σ T_e (unfiltered triples of compat clades)
    std::vector<double> aggC, aggT;
    double aggG = 0, aggTroot = 0;
    int64_t algo_bytes = 0;
    std::vector<std::vector<uint32_t>> famC;  // [F][nn] compat counts (for ℓ layout)
};

static void build_plan(const whale_model& m, bool grad, Plan& pl) {
    const int nn = m.nn;
    std::vector<std::vector<int>> act(nn);
    for (int oi = 0; oi < nn; oi++) {
        int e = m.order[oi];
        std::vector<int> a;
        if (grad) {
            auto add = [&](int s) { if (s >= 0) a.push_back(s); };
            if (m.child0[e] >= 0) a.insert(a.end(), act[m.child0[e]].begin(), act[m.child0[e]].end());
            if (m.child1[e] >= 0) a.insert(a.end(), act[m.child1[e]].begin(), act[m.child1[e]].end());
            if (m.kind[e] == WHALE_ROOT) add(m.eta_slot);
            else { add(m.lam_slot[e]); add(m.mu_slot[e]); }
            if (m.kind[e] == WHALE_WGD) add(m.q_slot[e]);
            std::sort(a.begin(), a.end());
            a.erase(std::unique(a.begin(), a.end()), a.end());
        }
        act[e] = a;
    }
    pl.Kmax = 1;
    for (int e = 0; e < nn; e++) pl.Kmax = std::max(pl.Kmax, 1 + (int)act[e].size());
    const int Kmax = pl.Kmax;
    pl.K.assign(nn, 1);
    pl.act.assign((size_t)nn * Kmax, -1);
    pl.cmap.assign((size_t)nn * 2 * Kmax, -1);
    pl.role.assign((size_t)nn * Kmax, 0);
    pl.toff.assign(nn, 0);
    size_t off = 0;
    for (int e = 0; e < nn; e++) {
        pl.K[e] = 1 + (int)act[e].size();
        pl.toff[e] = (int)off;
        off += (size_t)(m.nsl[e] + 1) * pl.K[e];
        for (int k = 1; k < pl.K[e]; k++) {
            int gp = act[e][k - 1];
            pl.act[(size_t)e * Kmax + k] = gp;
            uint8_t r = 0;
            if (m.kind[e] != WHALE_ROOT) {
                if (gp == m.lam_slot[e]) r |= 1;
                if (gp == m.mu_slot[e]) r |= 2;
            }
            if (m.kind[e] == WHALE_WGD && gp == m.q_slot[e]) r |= 4;
            if (m.kind[e] == WHALE_ROOT && gp == m.eta_slot) r |= 8;
            pl.role[(size_t)e * Kmax + k] = r;
            for (int j = 0; j < 2; j++) {
                int c = j == 0 ? m.child0[e] : m.child1[e];
                if (c < 0) continue;
                auto it = std::lower_bound(act[c].begin(), act[c].end(), gp);
                if (it != act[c].end() && *it == gp)
                    pl.cmap[((size_t)e * 2 + j) * Kmax + k] = (int16_t)(1 + (it - act[c].begin()));
            }
        }
        pl.cmap[((size_t)e * 2 + 0) * Kmax + 0] = 0;
        pl.cmap[((size_t)e * 2 + 1) * Kmax + 0] = 0;
    }
    pl.tab_len = off;
}

static cudaError_t upload_plan(Plan& pl, int nn) {
    cudaError_t e;
    int *dK, *dact, *dtoff;
    int16_t* dcmap;
    uint8_t* drole;
    if ((e = upload(pl.K, &dK)) != cudaSuccess) return e;
    if ((e = upload(pl.act, &dact)) != cudaSuccess) return e;
    if ((e = upload(pl.toff, &dtoff)) != cudaSuccess) return e;
    if ((e = upload(pl.cmap, &dcmap)) != cudaSuccess) return e;
    if ((e = upload(pl.role, &drole)) != cudaSuccess) return e;
    double *eps, *cx, *cy, *leaf, *cond;
    double2* pp;
    if ((e = cudaMalloc((void**)&eps, std::max<size_t>(pl.tab_len, 1) * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&pp, std::max<size_t>(pl.tab_len, 1) * sizeof(double2))) != cudaSuccess) return e;
    size_t nk = (size_t)nn * pl.Kmax * sizeof(double);
    if ((e = cudaMalloc((void**)&cx, nk)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&cy, nk)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&leaf, nk)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&cond, 3 * pl.Kmax * sizeof(double))) != cudaSuccess) return e;
    cudaMemset(cx, 0, nk); cudaMemset(cy, 0, nk); cudaMemset(leaf, 0, nk);
    cudaMemset(cond, 0, 3 * pl.Kmax * sizeof(double));
    pl.owned = {dK, dact, dtoff, dcmap, drole, eps, pp, cx, cy, leaf, cond};
    pl.dev = PlanDev{pl.Kmax, dK, dact, dcmap, drole, dtoff, eps, pp, cx, cy, leaf, cond};
    return cudaSuccess;
}

static int g_device = 0;

extern "C" {

int32_t whale_version(void) { return 100; }

int32_t whale_last_error(char* buf, size_t n) {
    if (buf && n) {
        strncpy(buf, g_err, n - 1);
        buf[n - 1] = 0;
    }
    return WHALE_OK;
}

int32_t whale_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t whale_set_device(int32_t device) {
    CU(cudaSetDevice(device));
    g_device = device;
    return WHALE_OK;
}

int64_t whale_launch_count(void) { return g_launches.load(); }

int32_t whale_model_create(const whale_model_desc* d, whale_model_t* out) {
    if (!d || !out) return fail(WHALE_ERR_ARG, "null argument");
    const int nn = d->n_nodes;
    if (nn < 3) return fail(WHALE_ERR_ARG, "n_nodes must be >= 3");
    if (whale_device_count() <= 0) return fail(WHALE_ERR_CUDA, "no CUDA device available (libwhalecuda has no CPU fallback)");
    CU(cudaSetDevice(g_device));
    auto* m = new whale_model();
    m->device = g_device;
    m->nn = nn;
    m->P = d->n_params;
    m->order.assign(d->order, d->order + nn);
    m->child0.assign(d->child0, d->child0 + nn);
    m->child1.assign(d->child1, d->child1 + nn);
    m->kind.assign(d->kind, d->kind + nn);
    m->nsl.assign(d->n_slices, d->n_slices + nn);
    m->dt.assign(d->slice_dt, d->slice_dt + nn);
    m->leafP.assign(d->leafP, d->leafP + nn);
    m->lam_slot.assign(d->lam_slot, d->lam_slot + nn);
    m->mu_slot.assign(d->mu_slot, d->mu_slot + nn);
    m->q_slot.assign(d->q_slot, d->q_slot + nn);
    m->eta_slot = d->eta_slot;
    m->log_scale = d->log_scale;
    // validation
    std::vector<int> seen(nn, 0), height(nn, 0);
    for (int oi = 0; oi < nn; oi++) {
        int e = m->order[oi];
        if (e < 0 || e >= nn || seen[e]) { delete m; return fail(WHALE_ERR_ARG, "order is not a permutation"); }
        int k = m->kind[e];
        int c0 = m->child0[e], c1 = m->child1[e];
        bool ok = (k == WHALE_LEAF && c0 < 0 && c1 < 0) || (k == WHALE_WGD && c0 >= 0 && c1 < 0) ||
                  ((k == WHALE_INTERNAL || k == WHALE_ROOT) && c0 >= 0 && c1 >= 0);
        if (!ok) { delete m; return fail(WHALE_ERR_ARG, "node %d: kind/children mismatch", e); }
        if ((c0 >= 0 && !seen[c0]) || (c1 >= 0 && !seen[c1])) { delete m; return fail(WHALE_ERR_ARG, "order is not children-first at node %d", e); }
        if (m->nsl[e] < 0) { delete m; return fail(WHALE_ERR_ARG, "negative slice count"); }
        if (k == WHALE_WGD && (m->q_slot[e] < 0 || m->q_slot[e] >= m->P)) { delete m; return fail(WHALE_ERR_ARG, "wgd node %d without q slot", e); }
        if (m->lam_slot[e] >= m->P || m->mu_slot[e] >= m->P) { delete m; return fail(WHALE_ERR_ARG, "rate slot out of range"); }
        seen[e] = 1;
        height[e] = std::max(c0 >= 0 ? height[c0] + 1 : 0, c1 >= 0 ? height[c1] + 1 : 0);
        if (k == WHALE_ROOT) m->root = e;
    }
    if (m->root != m->order[nn - 1] || m->nsl[m->root] != 0) { delete m; return fail(WHALE_ERR_ARG, "root must be last in order and have 0 slices"); }
    if (m->eta_slot < 0 || m->eta_slot >= m->P) { delete m; return fail(WHALE_ERR_ARG, "eta slot out of range"); }
    int maxh = *std::max_element(height.begin(), height.end());
    m->lvl_off.assign(1, 0);
    for (int h = 0; h <= maxh; h++) {
        for (int oi = 0; oi < nn; oi++) if (height[m->order[oi]] == h) m->lvl_nodes.push_back(m->order[oi]);
        m->lvl_off.push_back((int)m->lvl_nodes.size());
    }
    int *o, *c0, *c1, *kd, *ns, *ls, *ms, *qs, *lo, *ln;
    double *dt, *lp;
    CU(upload(m->order, &o)); CU(upload(m->child0, &c0)); CU(upload(m->child1, &c1)); CU(upload(m->kind, &kd));
    CU(upload(m->nsl, &ns)); CU(upload(m->lam_slot, &ls)); CU(upload(m->mu_slot, &ms)); CU(upload(m->q_slot, &qs));
    CU(upload(m->lvl_off, &lo)); CU(upload(m->lvl_nodes, &ln)); CU(upload(m->dt, &dt)); CU(upload(m->leafP, &lp));
    m->owned = {o, c0, c1, kd, ns, ls, ms, qs, lo, ln, dt, lp};
    m->dev = ModelDev{nn, o, c0, c1, kd, ns, dt, lp, ls, ms, qs, m->eta_slot, m->log_scale, m->root,
                      (int)m->lvl_off.size() - 1, lo, ln};
    for (int g = 0; g < 2; g++) {
        build_plan(*m, g == 1, m->plan[g]);
        CU(upload_plan(m->plan[g], nn));
    }
    CU(cudaMalloc((void**)&m->d_x, std::max(1, m->P) * sizeof(double)));
    CU(cudaMalloc((void**)&m->d_pleaf, nn * sizeof(double)));
    CU(cudaMemset(m->d_pleaf, 0, nn * sizeof(double)));
    CU(cudaMalloc((void**)&m->d_out, (1 + m->P) * sizeof(double)));
    CU(cudaMallocHost((void**)&m->h_pin, (2 + 2 * m->P + nn) * sizeof(double)));
    CU(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    *out = m;
    return WHALE_OK;
}

int32_t whale_model_destroy(whale_model_t m) {
    if (!m) return WHALE_OK;
    cudaSetDevice(m->device);
    for (void* p : m->owned) cudaFree(p);
    for (int g = 0; g < 2; g++) for (void* p : m->plan[g].owned) cudaFree(p);
    cudaFree(m->d_x); cudaFree(m->d_pleaf); cudaFree(m->d_out);
    if (m->h_pin) cudaFreeHost(m->h_pin);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return WHALE_OK;
}

// ---- the packer: reference-layout CSR -> per-branch resolved device arena ----
int32_t whale_data_create(whale_model_t m, const whale_ccd_desc* d, whale_data_t* out) {
    if (!m || !d || !out) return fail(WHALE_ERR_ARG, "null argument");
    const int nn = m->nn, F = d->n_fam;
    if (F <= 0) return fail(WHALE_ERR_ARG, "n_fam must be positive");
    CU(cudaSetDevice(m->device));
    auto* D = new whale_data();
    D->m = m;
    D->F = F;
    D->hdr.resize(F);
    D->aggC.assign(nn, 0.0);
    D->aggT.assign(nn, 0.0);
    D->famC.resize(F);
    std::vector<unsigned char>& A = D->arena_host;
    std::vector<double> work(F, 0.0);
    std::vector<int32_t> lidx;  // local index of clade γ at node e: lidx[e*G + γ]
    uint64_t ell_total = 0;
    int64_t algo_bytes = 0;
    for (int f = 0; f < F; f++) {
        const int64_t cb = d->clade_off[f];
        const int G = (int)(d->clade_off[f + 1] - cb);
        if (G < 1 || G > 65535) { delete D; return fail(WHALE_ERR_ARG, "family %d: %d clades (must be 1..65535, UInt16 ids)", f, G); }
        const int32_t* nleaf = d->clade_nleaf + cb;
        const int64_t* soff = d->split_off + cb;
        const int64_t* coff = d->compat_off + (int64_t)f * nn;
        lidx.assign((size_t)nn * G, -1);
        std::vector<uint32_t>& Cs = D->famC[f];
        Cs.assign(nn, 0);
        const uint64_t fam_ell0 = ell_total;
        for (int e = 0; e < nn; e++) {
            int C = (int)(coff[e + 1] - coff[e]);
            Cs[e] = C;
            int prev = -1;
            for (int j = 0; j < C; j++) {
                int g = d->compat[coff[e] + j];
                if (g < 0 || g >= G || g <= prev) { delete D; return fail(WHALE_ERR_ARG, "family %d node %d: compat list must be ascending clade ids", f, e); }
                prev = g;
                lidx[(size_t)e * G + g] = j;
            }
        }
        if ((int)Cs[m->root] != G) { delete D; return fail(WHALE_ERR_ARG, "family %d: every clade must be compatible with the root", f); }
        for (int c = 1; c < G; c++)
            if (nleaf[c] < nleaf[c - 1]) { delete D; return fail(WHALE_ERR_ARG, "family %d: clades must be sorted by size", f); }
        // blob assembly
        std::vector<NodeRec> recs(nn);
        std::vector<uint32_t> wordsv;   // pointer/loss/level words
        std::vector<Ent> entsv;         // entries
        uint32_t sumC = 0, maxC = 0;
        double wk = 0.0;
        for (int e = 0; e < nn; e++) {
            NodeRec& R = recs[e];
            memset(&R, 0, sizeof(R));
            const int C = (int)Cs[e];
            R.C = C;
            sumC += C;
            maxC = std::max<uint32_t>(maxC, C);
            const int kind = m->kind[e];
            int nonleaf = 0;
            double Te = 0;
            for (int j = 0; j < C; j++) {
                int g = d->compat[coff[e] + j];
                if (nleaf[g] > 1) nonleaf++;
                Te += (double)(soff[g + 1] - soff[g]);
            }
            R.nonleaf = nonleaf;
            D->aggC[e] += C;
            D->aggT[e] += Te;
            // (1) same-branch terms: within-branch duplication (src/core.jl:178-185), Πroot (:151-158);
            //     for WGD nodes the same list drives Πwgdretention on the child's row (:187-194), the
            //     child's compat list being identical.
            R.dptr_off = (uint32_t)wordsv.size();
            R.dent_off = (uint32_t)entsv.size();
            {
                const int src = (kind == WHALE_WGD) ? m->child0[e] : e;
                if (kind == WHALE_WGD && Cs[src] != (uint32_t)C) { delete D; return fail(WHALE_ERR_ARG, "family %d: WGD node %d and its child must have identical compat lists", f, e); }
                for (int j = 0; j < C; j++) {
                    int g = d->compat[coff[e] + j];
                    wordsv.push_back((uint32_t)entsv.size() - R.dent_off);
                    for (int64_t t = soff[g]; t < soff[g + 1]; t++) {
                        int g1 = d->g1[t], g2 = d->g2[t];
                        if (g1 < 0 || g1 >= G || g2 < 0 || g2 >= G) { delete D; return fail(WHALE_ERR_ARG, "family %d: triple out of range", f); }
                        int i1 = lidx[(size_t)src * G + g1], i2 = lidx[(size_t)src * G + g2];
                        if (i1 < 0 || i2 < 0) continue;  // getl == 0 (src/ccd.jl:43)
                        entsv.push_back(Ent{(uint16_t)i1, (uint16_t)i2, 0u, d->p[t]});
                    }
                }
                wordsv.push_back((uint32_t)entsv.size() - R.dent_off);
                wk += (double)(entsv.size() - R.dent_off) * (m->nsl[e] + 1) + (double)C * (m->nsl[e] + 1);
            }
            // (2) speciation terms at row 1 (src/core.jl:160-170) and loss indices (:172-176)
            if (kind == WHALE_INTERNAL || kind == WHALE_ROOT) {
                const int fch = m->child0[e], gch = m->child1[e];
                R.tptr_off = (uint32_t)wordsv.size();
                R.tent_off = (uint32_t)entsv.size();
                for (int j = 0; j < C; j++) {
                    int g = d->compat[coff[e] + j];
                    wordsv.push_back((uint32_t)entsv.size() - R.tent_off);
                    for (int64_t t = soff[g]; t < soff[g + 1]; t++) {
                        int g1 = d->g1[t], g2 = d->g2[t];
                        int f1 = lidx[(size_t)fch * G + g1], gg2 = lidx[(size_t)gch * G + g2];
                        int gg1 = lidx[(size_t)gch * G + g1], f2 = lidx[(size_t)fch * G + g2];
                        if (f1 >= 0 && gg2 >= 0) entsv.push_back(Ent{(uint16_t)f1, (uint16_t)gg2, 0u, d->p[t]});
                        if (gg1 >= 0 && f2 >= 0) entsv.push_back(Ent{(uint16_t)f2, (uint16_t)gg1, 0u, d->p[t]});
                    }
                }
                wordsv.push_back((uint32_t)entsv.size() - R.tent_off);
                wk += (double)(entsv.size() - R.tent_off);
                R.loss_off = (uint32_t)wordsv.size();
                for (int j = 0; j < C; j++) wordsv.push_back((uint32_t)lidx[(size_t)fch * G + d->compat[coff[e] + j]]);
                for (int j = 0; j < C; j++) wordsv.push_back((uint32_t)lidx[(size_t)gch * G + d->compat[coff[e] + j]]);
            }
            if (kind == WHALE_ROOT) {
                R.lev_off = (uint32_t)wordsv.size();
                uint32_t nlev = 0;
                for (int c = 0; c < G; c++)
                    if (c == 0 || nleaf[c] != nleaf[c - 1]) { wordsv.push_back((uint32_t)c); nlev++; }
                wordsv.push_back((uint32_t)G);
                D->hdr[f].nlev = nlev;
                D->aggTroot += (double)(soff[G] - soff[0]);
            }
            ell_total += (uint64_t)(m->nsl[e] + 1) * C;
        }
        D->aggG += G;
        // serialise: NodeRec[nn] | words | (pad to 16) | entries ; fix offsets to be blob-relative
        size_t base = (A.size() + 15) & ~size_t(15);
        size_t rec_bytes = (size_t)nn * sizeof(NodeRec);
        size_t words_at = rec_bytes;  // NodeRec is 32 bytes -> stays 4-byte aligned
        size_t ents_at = (words_at + wordsv.size() * 4 + 15) & ~size_t(15);
        size_t total = ents_at + entsv.size() * sizeof(Ent);
        A.resize(base + total, 0);
        for (int e = 0; e < nn; e++) {
            NodeRec& R = recs[e];
            R.dptr_off += (uint32_t)(words_at / 4);
            R.tptr_off += (uint32_t)(words_at / 4);
            R.loss_off += (uint32_t)(words_at / 4);
            R.lev_off += (uint32_t)(words_at / 4);
            R.dent_off += (uint32_t)(ents_at / 16);
            R.tent_off += (uint32_t)(ents_at / 16);
        }
        memcpy(A.data() + base, recs.data(), rec_bytes);
        if (!wordsv.empty()) memcpy(A.data() + base + words_at, wordsv.data(), wordsv.size() * 4);
        if (!entsv.empty()) memcpy(A.data() + base + ents_at, entsv.data(), entsv.size() * sizeof(Ent));
        D->hdr[f].base = base;
        D->hdr[f].G = G;
        D->hdr[f].sumC = sumC;
        D->hdr[f].maxC = maxC;
        D->hdr[f].ell_off = fam_ell0;
        work[f] = wk;
        // SURVEY §8d algorithmic bytes per evaluation: 12·T + 2·Γ + 4·Σ_e C_e
        algo_bytes += 12 * (soff[G] - soff[0]) + 2 * (int64_t)G + 4 * (int64_t)sumC;
        for (int g = 0; g < 2; g++) {
            const Plan& pl = m->plan[g];
            int s = 0, mx = 0;
            for (int e = 0; e < nn; e++) { int ck = (int)Cs[e] * pl.K[e]; s += ck; mx = std::max(mx, ck); }
            D->maxSumCK[g] = std::max(D->maxSumCK[g], s + mx);
        }
    }
    D->ell_total = ell_total;
    D->algo_bytes = algo_bytes;
    D->perm.resize(F);
    std::iota(D->perm.begin(), D->perm.end(), 0);
    std::stable_sort(D->perm.begin(), D->perm.end(), [&](int a, int b) { return work[a] > work[b]; });
    CU(cudaMalloc((void**)&D->d_arena, std::max<size_t>(A.size(), 16)));
    CU(cudaMemcpy(D->d_arena, A.data(), A.size(), cudaMemcpyHostToDevice));
    D->arena_bytes = A.size();
    std::vector<unsigned char>().swap(A);
    CU(upload(D->hdr, &D->d_hdr));
    CU(upload(D->perm, &D->d_perm));
    CU(cudaMalloc((void**)&D->d_out_fam, (size_t)F * m->plan[1].Kmax * sizeof(double)));
    CU(cudaMalloc((void**)&D->d_partial, (size_t)1024 * m->plan[1].Kmax * sizeof(double)));
    *out = D;
    return WHALE_OK;
}

int32_t whale_data_destroy(whale_data_t d) {
    if (!d) return WHALE_OK;
    cudaSetDevice(d->m->device);
    cudaFree(d->d_arena); cudaFree(d->d_hdr); cudaFree(d->d_perm); cudaFree(d->d_out_fam); cudaFree(d->d_partial);
    cudaFree(d->d_ell);
    for (int i = 0; i < 4; i++) if (d->ev[i]) cudaEventDestroy(d->ev[i]);
    delete d;
    return WHALE_OK;
}

int32_t whale_data_nfam(whale_data_t d) { return d ? d->F : 0; }
int64_t whale_data_arena_bytes(whale_data_t d) { return d ? (int64_t)d->arena_bytes : 0; }
int64_t whale_data_arena_dump(whale_data_t d, void* buf, int64_t cap) {
    if (!d) return 0;
    int64_t n = (int64_t)d->arena_bytes;
    if (buf && cap >= n) {
        // read back from the DEVICE copy: this is what the kernels see
        if (cudaMemcpy(buf, d->d_arena, n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    }
    return n;
}

// enqueue tables -> DP -> reduction on `st`; result in d_out (1+P doubles)
static int32_t enqueue_eval(whale_model* m, whale_data* D, const double* d_x, int32_t condition, uint32_t flags,
                            double* d_out, cudaStream_t st) {
    if (condition < 0 || condition > 2) return fail(WHALE_ERR_ARG, "unknown condition kind %d", condition);
    const int g = (flags & WHALE_WANT_GRAD) ? 1 : 0;
    Plan& pl = m->plan[g];
    const int nn = m->nn, F = D->F;
    const bool keep = (flags & WHALE_KEEP_ELL) != 0;
    if (keep && !D->d_ell) CU(cudaMalloc((void**)&D->d_ell, std::max<uint64_t>(D->ell_total, 1) * sizeof(double)));
    const bool prof = (flags & WHALE_PROFILE) != 0;
    if (prof && !D->ev[0]) for (int i = 0; i < 4; i++) CU(cudaEventCreate(&D->ev[i]));
    if (prof) CU(cudaEventRecord(D->ev[0], st));
    // K1
    int nw = std::min(32, std::max(1, nn));
    LAUNCH(k_tables, 1, nw * 32, 0, st, m->dev, pl.dev, d_x, m->d_pleaf);
    g_launches++;
    if (prof) CU(cudaEventRecord(D->ev[1], st));
    // K2
    constexpr int NT = 128;
    if (pl.Kmax > NT) return fail(WHALE_ERR_CAPACITY, "K=%d tangent components exceed %d lanes (parameter chunking not built yet)", pl.Kmax, NT);
    size_t smem = (((nn + 1) * sizeof(int) + 15) & ~size_t(15)) + (size_t)D->maxSumCK[g] * sizeof(double);
    if (smem > 227 * 1024) return fail(WHALE_ERR_CAPACITY, "a family needs %zu bytes of shared memory (> 227 KB)", smem);
    static thread_local size_t smem_set = 0;
    if (smem > smem_set) {
        CU(cudaFuncSetAttribute(k_dp<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        smem_set = smem;
    }
    DPArgs a{m->dev, pl.dev, D->d_arena, D->d_hdr, D->d_perm, D->d_out_fam, keep ? D->d_ell : nullptr, F, keep ? 0 : 1};
    LAUNCH(k_dp<NT>, F, NT, smem, st, a);
    g_launches++;
    if (prof) CU(cudaEventRecord(D->ev[2], st));
    // K3
    const int KR = pl.K[m->root];
    int nb = std::min(1024, (F + 255) / 256);
    int chunk = (F + nb - 1) / nb;
    LAUNCH(k_reduce1, nb, 256, 0, st, D->d_out_fam, F, KR, chunk, D->d_partial);
    LAUNCH(k_reduce2, 1, 256, 0, st, D->d_partial, nb, KR, F, condition, pl.dev, m->root, m->P, d_out);
    g_launches += 2;
    if (prof) CU(cudaEventRecord(D->ev[3], st));
    D->ev_valid = prof;
    CU(cudaGetLastError());
    D->ell_valid = keep;
    return WHALE_OK;
}

int32_t whale_logpdf_grad_async(whale_model_t m, whale_data_t d, const double* d_x, int32_t condition,
                                uint32_t flags, double* d_out, void* stream) {
    if (!m || !d || !d_x || !d_out) return fail(WHALE_ERR_ARG, "null argument");
    if (d->m != m) return fail(WHALE_ERR_ARG, "data handle belongs to another model");
    CU(cudaSetDevice(m->device));
    return enqueue_eval(m, d, d_x, condition, flags, d_out, (cudaStream_t)stream);
}

int32_t whale_logpdf_grad(whale_model_t m, whale_data_t d, const double* x, const double* p_leaf, int32_t condition,
                          uint32_t flags, double* loglik, double* grad, double* ll_fam, double* grad_fam) {
    if (!m || !d || !x || !loglik) return fail(WHALE_ERR_ARG, "null argument");
    if (d->m != m) return fail(WHALE_ERR_ARG, "data handle belongs to another model");
    if ((grad || grad_fam) && !(flags & WHALE_WANT_GRAD)) return fail(WHALE_ERR_ARG, "grad requested without WHALE_WANT_GRAD");
    CU(cudaSetDevice(m->device));
    const int P = m->P, nn = m->nn;
    double* hp = m->h_pin;
    memcpy(hp, x, P * sizeof(double));
    for (int e = 0; e < nn; e++) hp[P + e] = p_leaf ? p_leaf[e] : 0.0;
    CU(cudaMemcpyAsync(m->d_x, hp, P * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(m->d_pleaf, hp + P, nn * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    int32_t rc = enqueue_eval(m, d, m->d_x, condition, flags, m->d_out, m->stream);
    if (rc != WHALE_OK) return rc;
    double* ho = hp + P + nn;
    CU(cudaMemcpyAsync(ho, m->d_out, (1 + P) * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    *loglik = ho[0];
    if (grad) memcpy(grad, ho + 1, P * sizeof(double));
    if (ll_fam || grad_fam) {
        const Plan& pl = m->plan[(flags & WHALE_WANT_GRAD) ? 1 : 0];
        const int KR = pl.K[m->root];
        std::vector<double> tmp((size_t)d->F * KR);
        CU(cudaMemcpy(tmp.data(), d->d_out_fam, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int f = 0; f < d->F; f++) {
            if (ll_fam) ll_fam[f] = tmp[(size_t)f * KR];
            if (grad_fam) {
                for (int p = 0; p < P; p++) grad_fam[(size_t)f * P + p] = 0.0;
                for (int k = 1; k < KR; k++) grad_fam[(size_t)f * P + pl.act[(size_t)m->root * pl.Kmax + k]] = tmp[(size_t)f * KR + k];
            }
        }
    }
    return WHALE_OK;
}

int32_t whale_slices(whale_model_t m, const double* x, const double* p_leaf, double* eps, double* phi, double* psi) {
    if (!m || !x || !eps || !phi || !psi) return fail(WHALE_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->device));
    const int P = m->P, nn = m->nn;
    std::vector<double> pl(nn, 0.0);
    if (p_leaf) pl.assign(p_leaf, p_leaf + nn);
    CU(cudaMemcpy(m->d_x, x, P * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m->d_pleaf, pl.data(), nn * sizeof(double), cudaMemcpyHostToDevice));
    Plan& p0 = m->plan[0];
    LAUNCH(k_tables, 1, std::min(32, nn) * 32, 0, m->stream, m->dev, p0.dev, m->d_x, m->d_pleaf);
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(m->stream));
    std::vector<double> he(p0.tab_len);
    std::vector<double2> hp(p0.tab_len);
    CU(cudaMemcpy(he.data(), p0.dev.eps, p0.tab_len * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hp.data(), p0.dev.pp, p0.tab_len * sizeof(double2), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < p0.tab_len; i++) { eps[i] = he[i]; phi[i] = hp[i].x; psi[i] = hp[i].y; }
    return WHALE_OK;
}

int64_t whale_ell_size(whale_data_t d, int32_t fam) {
    if (!d || fam < 0 || fam >= d->F) return -1;
    int64_t s = 0;
    for (int e = 0; e < d->m->nn; e++) s += (int64_t)(d->m->nsl[e] + 1) * d->famC[fam][e];
    return s;
}

int32_t whale_ell_get(whale_data_t d, int32_t fam, double* out) {
    if (!d || !out || fam < 0 || fam >= d->F) return fail(WHALE_ERR_ARG, "bad argument");
    if (!d->ell_valid || !d->d_ell) return fail(WHALE_ERR_STATE, "no ℓ kept: evaluate with WHALE_KEEP_ELL first");
    CU(cudaSetDevice(d->m->device));
    CU(cudaMemcpy(out, d->d_ell + d->hdr[fam].ell_off, whale_ell_size(d, fam) * sizeof(double), cudaMemcpyDeviceToHost));
    return WHALE_OK;
}

int32_t whale_backtrack(whale_model_t, whale_data_t, int32_t, const double*, int64_t, int32_t, int32_t*, int32_t*,
                        int32_t*, int32_t*, int32_t*, int32_t*) {
    return fail(WHALE_ERR_STATE, "whale_backtrack: kernel not built yet");
}

int32_t whale_work_estimate(whale_model_t m, whale_data_t d, uint32_t flags, double* flops, double* bytes) {
    if (!m || !d) return fail(WHALE_ERR_ARG, "null argument");
    const Plan& pl = m->plan[(flags & WHALE_WANT_GRAD) ? 1 : 0];
    double fl = 0.0;
    for (int e = 0; e < m->nn; e++) {  // SURVEY §8d fixed coefficients
        const double C = d->aggC[e], T = d->aggT[e], Pe = pl.K[e] - 1, n = m->nsl[e];
        fl += n * (3 * T + 3 * C + Pe * (5 * T + 8 * C));
        if (m->kind[e] == WHALE_INTERNAL) fl += 6 * T + 4 * C + Pe * (10 * T + 8 * C);
        else if (m->kind[e] == WHALE_WGD) fl += 3 * T + 5 * C + Pe * (5 * T + 8 * C);
        else if (m->kind[e] == WHALE_ROOT) fl += 9 * d->aggTroot + 10 * d->aggG + Pe * (15 * d->aggTroot + 16 * d->aggG);
    }
    if (flops) *flops = fl;
    if (bytes) *bytes = (double)d->algo_bytes + 8.0 * d->F * pl.K[m->root];
    return WHALE_OK;
}

int32_t whale_last_kernel_ms(whale_data_t d, double* tables_ms, double* dp_ms, double* reduce_ms) {
    if (!d) return fail(WHALE_ERR_ARG, "null argument");
    if (!d->ev_valid) return fail(WHALE_ERR_STATE, "last evaluation was not run with WHALE_PROFILE");
    CU(cudaSetDevice(d->m->device));
    CU(cudaEventSynchronize(d->ev[3]));
    float a = 0, b = 0, c = 0;
    CU(cudaEventElapsedTime(&a, d->ev[0], d->ev[1]));
    CU(cudaEventElapsedTime(&b, d->ev[1], d->ev[2]));
    CU(cudaEventElapsedTime(&c, d->ev[2], d->ev[3]));
    if (tables_ms) *tables_ms = a;
    if (dp_ms) *dp_ms = b;
    if (reduce_ms) *reduce_ms = c;
    return WHALE_OK;
}

int32_t whale_fp64_peak(double* tflops) {
    if (!tflops) return fail(WHALE_ERR_ARG, "null argument");
    CU(cudaSetDevice(g_device));
    cudaDeviceProp pr;
    CU(cudaGetDeviceProperties(&pr, g_device));
    const int blocks = pr.multiProcessorCount * 8, iters = 1 << 14;
    double* d;
    CU(cudaMalloc((void**)&d, (size_t)blocks * 256 * sizeof(double)));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    LAUNCH(k_dfma, blocks, 256, 0, 0, d, iters);
    CU(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CU(cudaEventRecord(a));
        LAUNCH(k_dfma, blocks, 256, 0, 0, d, iters);
        CU(cudaEventRecord(b));
        CU(cudaEventSynchronize(b));
        float ms;
        CU(cudaEventElapsedTime(&ms, a, b));
        best = std::min(best, ms);
    }
    g_launches += 6;
    cudaFree(d); cudaEventDestroy(a); cudaEventDestroy(b);
    *tflops = 2.0 * 8.0 * iters * (double)blocks * 256 / (best * 1e-3) / 1e12;
    return WHALE_OK;
}

}  // extern "C"
