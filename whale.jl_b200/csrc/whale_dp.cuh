// K2 — the ALE recursion with forward tangents.  src/core.jl:83-199.
//
// One CTA (NT threads) per family.  The LAST row of every branch (C_e × K_e doubles: K_e = 1 + #raw
// parameters that can influence branch e) stays in shared memory for the parent; slices ping-pong between
// that row and a scratch row (one barrier per slice).  Lanes span (clade cell × component); every lane also
// recomputes the value part it needs, so no shuffles are required and the k = 0 lane is the plain logpdf.
//   phase A  leaf branches are independent of each other: one WARP per leaf branch, warp-level sync only;
//            branches whose compatible clades are all leaf clades are family-independent and are filled
//            from the table k_tables prepared (ℓ_n = leafℙ·Πϕ_i).
//   phase B  internal / WGD / root nodes in the reference's order (children first), all warps cooperating;
//            the node's pointer arrays and within-branch terms are staged in shared memory with cp.async
//            while row 1 (speciation + loss) is being computed.
#pragma once
#include "whale_common.cuh"

struct DPArgs {
    ModelDev M;
    PlanDev PL;
    const unsigned char* arena;
    const FamHdr* hdr;
    const int* perm;   // launch order
    double* out_fam;   // [F * Kroot]  (log L_f, ∂ log L_f / ∂ component)
    double* ell;       // keep_ell buffer or nullptr
    int plan;          // 0 value only, 1 with tangents
    int skip_leaf;     // share family-independent leaf-branch rows (off in keep_ell mode)
};

#ifdef WHALE_EMU
#define CP_ASYNC16(dst, src) (*(uint4*)(dst) = *(const uint4*)(src))
#define CP_ASYNC_WAIT() ((void)0)
#define PREFETCH_L2(p) ((void)0)
#else
#define CP_ASYNC16(dst, src)                                                                        \
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), \
                 "l"(src) : "memory")
#define CP_ASYNC_WAIT() asm volatile("cp.async.wait_all;" ::: "memory")
#define PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#endif

// Σ_t p_t X[i1] Y[i2] with the product rule for the lane's component; m = 0 on the value lane (k = 0,
// kx = ky = 0), 1 on tangent lanes; a component absent from a child (kx < 0) contributes a zero tangent.
template <bool GLOBAL_ENTS>
__device__ __forceinline__ void pairsum(const Ent* __restrict__ ents, uint32_t tb, uint32_t te,
                                        const double* __restrict__ X, int KX, int kx,
                                        const double* __restrict__ Y, int KY, int ky, double m, double& S0,
                                        double& Sk) {
    double s0 = 0.0, sk = 0.0;
    for (uint32_t t = tb; t < te; t++) {
        uint4 raw;
        if (GLOBAL_ENTS) raw = __ldg(reinterpret_cast<const uint4*>(ents + t));
        else raw = *reinterpret_cast<const uint4*>(ents + t);
        const double p = __hiloint2double((int)raw.w, (int)raw.z);
        const double* xp = X + (raw.x & 0xffffu) * KX;
        const double* yp = Y + (raw.x >> 16) * KY;
        const double x0 = xp[0], y0 = yp[0];
        const double xk = kx >= 0 ? xp[kx] : 0.0;
        const double yk = ky >= 0 ? yp[ky] : 0.0;
        const double px = p * x0;
        s0 = fma(px, y0, s0);
        sk = fma(px, yk, sk);
        sk = fma(m * (p * y0), xk, sk);
    }
    S0 = s0;
    Sk = sk;
}

// Πloss (src/core.jl:172-176) for cell c with child indices lf/lg (−1: incompatible -> getl = 0)
__device__ __forceinline__ void loss_term(int lf, int lg, const double* finF, int KF, int kf, const double* finG,
                                          int KG, int kg, double ef0, double efk, double eg0, double egk, double m,
                                          double& c0, double& ck) {
    double f0 = 0.0, fk = 0.0, g0 = 0.0, gk = 0.0;
    if (lf >= 0) { f0 = finF[lf * KF]; fk = kf >= 0 ? finF[lf * KF + kf] : 0.0; }
    if (lg >= 0) { g0 = finG[lg * KG]; gk = kg >= 0 ? finG[lg * KG + kg] : 0.0; }
    c0 = f0 * eg0 + g0 * ef0;
    ck = fk * eg0 + gk * ef0 + m * (f0 * egk + g0 * efk);
}

template <int NT>
__global__ void __launch_bounds__(NT) k_dp(DPArgs A, int perm_off) {
    EXTERN_SHARED(smem_raw);
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ModelDev& M = A.M;
    const PlanDev& PL = A.PL;
    const int nn = M.nn, Kmax = PL.Kmax;
    const int fam = A.perm[perm_off + blockIdx.x];
    const FamHdr* Hp = A.hdr + fam;
    struct { uint64_t base, ell_off; uint32_t nlev, rows_len, scr_len, leafmax, blob_bytes; } H;
    H.base = Hp->base; H.ell_off = Hp->ell_off; H.nlev = Hp->nlev; H.blob_bytes = Hp->blob_bytes;
    H.rows_len = Hp->rows_len[A.plan]; H.scr_len = Hp->scr_len[A.plan]; H.leafmax = Hp->leafmax[A.plan];
    const unsigned char* blob = A.arena + H.base;
    const NodeRec* nrec = reinterpret_cast<const NodeRec*>(blob);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(blob);
    const Ent* ents = reinterpret_cast<const Ent*>(blob);

    // pull the whole blob towards L2 now; it is consumed node by node below
    for (uint32_t o = tid * 128u; o < H.blob_bytes; o += NT * 128u) PREFETCH_L2(blob + o);

    int* s_roff = reinterpret_cast<int*>(smem_raw);  // [nn+1] row offsets (doubles)
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
    double* scr = rows + H.rows_len;
    unsigned char* stage = reinterpret_cast<unsigned char*>(scr + H.scr_len);
    double* const ell_base = A.ell ? A.ell + H.ell_off : nullptr;
    auto ell_of = [&](int e) -> double* {  // node e's matrix inside the family's ℓ (node-index order)
        if (!ell_base) return nullptr;
        size_t o = 0;
        for (int e2 = 0; e2 < e; e2++) o += (size_t)(M.nsl[e2] + 1) * nrec[e2].C;
        return ell_base + o;
    };

    // ================= phase A: leaf branches, one warp each (src/core.jl:83-101,121-128) =================
    for (int li = warp; li < M.nleafnodes; li += NW) {
        const int e = M.leafnodes[li];
        const NodeRec R = nrec[e];
        const int C = (int)R.C;
        if (C == 0) continue;
        const int K = PL.K[e], n = M.nsl[e];
        double* fin = rows + s_roff[e];
        const int GP = 32 / K;  // K_leaf <= 3
        const int grp = lane / K, k = lane - grp * K;
        const bool on = grp < GP;
        const double m = k == 0 ? 0.0 : 1.0;
        if (R.nonleaf == 0 && A.skip_leaf) {  // family-independent: ℓ_n = leafℙ·Πϕ_i from k_tables
            if (on)
                for (int c = grp; c < C; c += GP) fin[c * K + k] = PL.leaf[e * Kmax + k];
            continue;
        }
        double* ellp = ell_of(e);
        double* wscr = scr + warp * H.leafmax;
        double* cur = (n & 1) ? wscr : fin;  // row i lives in fin iff (n − i) is even
        const int nleafc = C - (int)R.nonleaf;
        if (on)
            for (int c = grp; c < C; c += GP) {
                const double v = (c < nleafc && k == 0) ? M.leafP[e] : 0.0;
                cur[c * K + k] = v;
                if (ellp && k == 0) ellp[c] = v;
            }
        __syncwarp();
        const uint32_t* dptr = words + R.dptr_off;
        const Ent* dents = ents + R.dent_off;
        const double2* pprow = PL.pp + PL.toff[e];
        for (int i = 1; i <= n; i++) {
            const double* src = cur;
            double* dst = (cur == fin) ? wscr : fin;
            if (on && grp < C) {
                const double2 c0 = pprow[(size_t)i * K];
                const double2 ck = pprow[(size_t)i * K + k];
                for (int c = grp; c < C; c += GP) {
                    double s0 = 0.0, sk = 0.0;
                    if (c >= nleafc) pairsum<true>(dents, dptr[c], dptr[c + 1], src, K, k, src, K, k, m, s0, sk);
                    const double o0 = src[c * K], ok = src[c * K + k];
                    const double r = c0.x * ok + c0.y * sk + m * (ck.x * o0 + ck.y * s0);
                    dst[c * K + k] = r;
                    if (ellp && k == 0) ellp[(size_t)i * C + c] = r;
                }
            }
            cur = dst;
            __syncwarp();
        }
    }
    __syncthreads();

    // ================= phase B: internal, WGD and root nodes, whole CTA =================
    for (int oi = 0; oi < M.ninner; oi++) {
        const int e = M.inner[oi];
        const NodeRec R = nrec[e];
        const int C = (int)R.C;
        if (C == 0) continue;
        const int kind = M.kind[e], K = PL.K[e], n = M.nsl[e];
        double* fin = rows + s_roff[e];
        double* ellp = ell_of(e);
        // lane -> (cell group, component): groups of K lanes, GP groups per pass (K <= NT checked on the host)
        const int GP = NT / K;
        const int grp = tid / K, k = tid - grp * K;
        const bool on = grp < GP;
        const double m = k == 0 ? 0.0 : 1.0;

        // ---- stage this node's lists: [dents | dptr | tptr,lossF,lossG,lev] ----
        const int nd16 = (kind == WHALE_ROOT) ? 0 : (int)R.ndent;     // Πroot terms are read once: stay global
        const int dp16 = (C + 1 + 3) >> 2;
        const int tp16 = (kind == WHALE_WGD) ? 0 : ((3 * C + 1 + (kind == WHALE_ROOT ? (int)H.nlev + 1 : 0) + 3) >> 2);
        uint4* st4 = reinterpret_cast<uint4*>(stage);
        {
            const uint4* g_de = reinterpret_cast<const uint4*>(ents + R.dent_off);
            const uint4* g_dp = reinterpret_cast<const uint4*>(words + R.dptr_off);
            const uint4* g_tp = reinterpret_cast<const uint4*>(words + R.tptr_off);
            for (int i = tid; i < nd16; i += NT) CP_ASYNC16(st4 + i, g_de + i);
            for (int i = tid; i < dp16; i += NT) CP_ASYNC16(st4 + nd16 + i, g_dp + i);
            for (int i = tid; i < tp16; i += NT) CP_ASYNC16(st4 + nd16 + dp16 + i, g_tp + i);
        }
        const Ent* s_dents = reinterpret_cast<const Ent*>(stage);
        const uint32_t* s_dptr = reinterpret_cast<const uint32_t*>(st4 + nd16);
        const uint32_t* s_tptr = reinterpret_cast<const uint32_t*>(st4 + nd16 + dp16);
        const int32_t* s_lossF = reinterpret_cast<const int32_t*>(s_tptr + C + 1);
        const int32_t* s_lossG = s_lossF + C;
        const uint32_t* s_lev = reinterpret_cast<const uint32_t*>(s_lossG + C);
        CP_ASYNC_WAIT();
        __syncthreads();

        // children, shared by row-1 formulas
        const int f = M.child0[e], g = M.child1[e];
        const int KF = PL.K[f];
        const double* finF = rows + s_roff[f];
        const int kf = k == 0 ? 0 : PL.cmap[(e * 2 + 0) * Kmax + k];

        if (kind == WHALE_ROOT) {
            // whaleroot! src/core.jl:130-149: clades ascending in size, level-synchronous, in place
            const int KG = PL.K[g];
            const double* finG = rows + s_roff[g];
            const int kg = k == 0 ? 0 : PL.cmap[(e * 2 + 1) * Kmax + k];
            const double* epsF = PL.eps + PL.toff[f] + (size_t)M.nsl[f] * KF;
            const double* epsG = PL.eps + PL.toff[g] + (size_t)M.nsl[g] * KG;
            const double ef0 = epsF[0], eg0 = epsG[0];
            const double efk = (k > 0 && kf >= 0) ? epsF[kf] : 0.0;
            const double egk = (k > 0 && kg >= 0) ? epsG[kg] : 0.0;
            const double cx0 = PL.cx[e * Kmax], cy0 = PL.cy[e * Kmax];
            const double cxk = PL.cx[e * Kmax + k], cyk = PL.cy[e * Kmax + k];
            const Ent* g_dents = ents + R.dent_off;
            const Ent* g_tents = ents + R.tent_off;
            for (uint32_t L = 0; L < H.nlev; L++) {
                const int c0 = (int)s_lev[L], c1 = (int)s_lev[L + 1];
                if (on)
                    for (int c = c0 + grp; c < c1; c += GP) {
                        double a0, ak, b0, bk, l0, lk;
                        pairsum<true>(g_dents, s_dptr[c], s_dptr[c + 1], fin, K, k, fin, K, k, m, a0, ak);
                        pairsum<true>(g_tents, s_tptr[c], s_tptr[c + 1], finF, KF, kf, finG, KG, kg, m, b0, bk);
                        loss_term(s_lossF[c], s_lossG[c], finF, KF, kf, finG, KG, kg, ef0, efk, eg0, egk, m, l0, lk);
                        const double u0 = b0 + l0, uk = bk + lk;
                        const double r = cx0 * ak + cy0 * uk + m * (cxk * a0 + cyk * u0);
                        fin[c * K + k] = r;
                        if (ellp && k == 0) ellp[c] = r;
                    }
                __syncthreads();
            }
            if (tid < K) {  // log L and its gradient (src/core.jl:35-36)
                const double Lv = fin[(C - 1) * K];
                double o;
                if (Lv > 0.0) o = tid == 0 ? log(Lv) : fin[(C - 1) * K + tid] / Lv;
                else o = tid == 0 ? -dinf() : 0.0;
                A.out_fam[(size_t)fam * K + tid] = o;
            }
            continue;
        }

        // ---- row 1 of an internal / WGD branch ----
        double* cur = (n & 1) ? scr : fin;  // row i lives in fin iff (n − i) is even
        if (kind == WHALE_INTERNAL) {  // Πspeciation + Πloss, src/core.jl:95-98,160-176
            const int KG = PL.K[g];
            const double* finG = rows + s_roff[g];
            const int kg = k == 0 ? 0 : PL.cmap[(e * 2 + 1) * Kmax + k];
            const double* epsF = PL.eps + PL.toff[f] + (size_t)M.nsl[f] * KF;
            const double* epsG = PL.eps + PL.toff[g] + (size_t)M.nsl[g] * KG;
            const double ef0 = epsF[0], eg0 = epsG[0];
            const double efk = (k > 0 && kf >= 0) ? epsF[kf] : 0.0;
            const double egk = (k > 0 && kg >= 0) ? epsG[kg] : 0.0;
            const Ent* g_tents = ents + R.tent_off;
            if (on)
                for (int c = grp; c < C; c += GP) {
                    double b0, bk, l0, lk;
                    pairsum<true>(g_tents, s_tptr[c], s_tptr[c + 1], finF, KF, kf, finG, KG, kg, m, b0, bk);
                    loss_term(s_lossF[c], s_lossG[c], finF, KF, kf, finG, KG, kg, ef0, efk, eg0, egk, m, l0, lk);
                    const double r = k == 0 ? b0 + l0 : bk + lk;
                    cur[c * K + k] = r;
                    if (ellp && k == 0) ellp[c] = r;
                }
        } else {  // WGD: q·Σ p ℓ_f[γ1]ℓ_f[γ2] + (1−q+2qϵ_f)·ℓ_f[γ]   src/core.jl:103-119,187-199
            const double cx0 = PL.cx[e * Kmax], cy0 = PL.cy[e * Kmax];
            const double cxk = PL.cx[e * Kmax + k], cyk = PL.cy[e * Kmax + k];
            if (on)
                for (int c = grp; c < C; c += GP) {
                    double s0, sk;
                    pairsum<false>(s_dents, s_dptr[c], s_dptr[c + 1], finF, KF, kf, finF, KF, kf, m, s0, sk);
                    const double u0 = finF[c * KF];
                    const double uk = kf >= 0 ? finF[c * KF + kf] : 0.0;
                    const double r = cy0 * sk + cx0 * uk + m * (cyk * s0 + cxk * u0);
                    cur[c * K + k] = r;
                    if (ellp && k == 0) ellp[c] = r;
                }
        }
        __syncthreads();

        // ---- slices: ℓ_i = ϕ_i ℓ_{i−1} + ψ_i Σ_t p ℓ_{i−1}[γ1] ℓ_{i−1}[γ2]   src/core.jl:121-128,178-185 ----
        const double2* pprow = PL.pp + PL.toff[e];
        uint32_t tb0 = 0, te0 = 0;  // the lane's first cell keeps its term range in registers
        if (on && grp < C) { tb0 = s_dptr[grp]; te0 = s_dptr[grp + 1]; }
        for (int i = 1; i <= n; i++) {
            const double* src = cur;
            double* dst = (cur == fin) ? scr : fin;
            if (on && grp < C) {
                const double2 c0 = pprow[(size_t)i * K];
                const double2 ck = pprow[(size_t)i * K + k];
                for (int c = grp; c < C; c += GP) {
                    uint32_t tb = tb0, te = te0;
                    if (c != grp) { tb = s_dptr[c]; te = s_dptr[c + 1]; }
                    double s0, sk;
                    pairsum<false>(s_dents, tb, te, src, K, k, src, K, k, m, s0, sk);
                    const double o0 = src[c * K], ok = src[c * K + k];
                    const double r = c0.x * ok + c0.y * sk + m * (ck.x * o0 + ck.y * s0);
                    dst[c * K + k] = r;
                    if (ellp && k == 0) ellp[(size_t)i * C + c] = r;
                }
            }
            cur = dst;
            __syncthreads();
        }
    }
}
