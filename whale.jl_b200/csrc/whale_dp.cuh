// K2 — the ALE recursion with forward tangents.  src/core.jl:83-199.
//
// One CTA (NT threads) per family.  The LAST row of every branch (C_e × K_e doubles: K_e = 1 + #raw
// parameters that can influence branch e) stays in shared memory for the parent; slices ping-pong between
// that row and a scratch row.  Every Σ_t p_t·ℓ[γ1]·ℓ[γ2] of the reference (Πduplication, Πspeciation,
// Πwgdretention, Πroot) is evaluated in two balanced phases instead of one serial loop per clade:
//   P1  one thread per TERM t: products p·x·y with the product rule for all K components -> prod[k][t]
//   P2  one lane per (cell, component): sums its contiguous range of prod (the reference's summation
//       order), applies the row formula, writes the new row.
// That removes the load imbalance of CCDs (a branch's largest clade typically owns ~10× the mean number of
// splits; the ubiquitous clade ~100) from the critical path.
//   phase A  leaf branches are independent of each other: one WARP per leaf branch, warp-level sync only;
//            branches whose compatible clades are all leaf clades are family-independent and are filled
//            from the table k_tables prepared (ℓ_n = leafℙ·Πϕ_i).
//   phase B  internal / WGD / root nodes in the reference's order (children first), all warps cooperating;
//            the node's pointer arrays and within-branch terms are staged in shared memory.
// ϕ/ψ rows are prefetched one slice ahead into registers, so the slice loop touches shared memory only.
#pragma once
#include "whale_common.cuh"

struct DPArgs {
    ModelDev M;
    PlanDev PL;
    const unsigned char* arena;
    const FamHdr* hdr;
    const int* perm;   // launch order
    const uint32_t* roff;  // [F][nn] row offsets (doubles) of this plan, from place_rows on the host
    double* out_fam;   // [F * Kroot]  (log L_f, ∂ log L_f / ∂ component)
    double* ell;       // keep_ell buffer or nullptr
    int plan;          // 0 value only, 1 with tangents
    int skip_leaf;     // share family-independent leaf-branch rows (off in keep_ell mode)
    long long* tim;    // optional per-family phase cycle counters [F][8] (profiling aid) or nullptr
};



#ifdef WHALE_EMU
#define PREFETCH_L2(p) ((void)0)
#else
#define PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#endif

template <bool WARP>
__device__ __forceinline__ void scope_sync() {
    if (WARP) __syncwarp();
    else __syncthreads();
}

// copy n16 16-byte words global -> shared with all threads of the scope.  cp.async keeps every copy of a
// staging batch in flight at once (one memory latency per node instead of one per loop iteration); the batch
// is completed by stage_wait() before the scope's barrier.
#ifdef WHALE_EMU
__device__ __forceinline__ void copy16(uint4* dst, const uint4* __restrict__ src, int n16, int tid, int nt) {
    for (int i = tid; i < n16; i += nt) dst[i] = src[i];
}
__device__ __forceinline__ void stage_wait() {}
#else
__device__ __forceinline__ void copy16(uint4* dst, const uint4* __restrict__ src, int n16, int tid, int nt) {
    for (int i = tid; i < n16; i += nt)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + i)),
                     "l"(src + i) : "memory");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

// P1: terms [ta, tb) of `ents` -> prod[k*cap + (t - ta)], k = 0..K-1.
// X/Y are rows with KX/KY components; mapX/mapY translate the output component k into the component of
// X/Y (identity if null; -1: the row does not depend on that parameter -> zero tangent).
template <bool GLOBAL_ENTS>
__device__ __forceinline__ void terms(const Ent* __restrict__ ents, uint32_t ta, uint32_t tb,
                                      const double* __restrict__ X, int KX, const int16_t* __restrict__ mapX,
                                      const double* __restrict__ Y, int KY, const int16_t* __restrict__ mapY,
                                      int K, double* __restrict__ prod, int cap, int tid, int nt) {
    for (uint32_t t = ta + tid; t < tb; t += nt) {
        uint4 raw;
        if (GLOBAL_ENTS) raw = __ldg(reinterpret_cast<const uint4*>(ents + t));
        else raw = *reinterpret_cast<const uint4*>(ents + t);
        const double p = __hiloint2double((int)raw.w, (int)raw.z);
        const double* xp = X + (raw.x & 0xffffu) * KX;
        const double* yp = Y + (raw.x >> 16) * KY;
        const double px = p * xp[0], py = p * yp[0];
        double* o = prod + (t - ta);
        o[0] = px * yp[0];
#pragma unroll 4
        for (int k = 1; k < K; k++) {
            const int kx = mapX ? mapX[k] : k, ky = mapY ? mapY[k] : k;
            const double xv = kx >= 0 ? xp[kx] : 0.0, yv = ky >= 0 ? yp[ky] : 0.0;
            o[(size_t)k * cap] = fma(px, yv, py * xv);
        }
    }
}

// P2 helper: (S0, Sk) of cell range [tb, te) (indices relative to the P1 window)
__device__ __forceinline__ void cellsum(const double* __restrict__ prod, int cap, int k, uint32_t tb, uint32_t te,
                                        double& S0, double& Sk) {
    double s0 = 0.0, sk = 0.0;
    const double* pk = prod + (size_t)k * cap;
    for (uint32_t t = tb; t < te; t++) {
        s0 += prod[t];
        sk += pk[t];
    }
    S0 = s0;
    Sk = sk;
}

// Πloss (src/core.jl:172-176) for a cell with child indices lf/lg (−1: incompatible -> getl = 0)
__device__ __forceinline__ void loss_term(int lf, int lg, const double* finF, int KF, int kf, const double* finG,
                                          int KG, int kg, double ef0, double efk, double eg0, double egk, double m,
                                          double& c0, double& ck) {
    double f0 = 0.0, fk = 0.0, g0 = 0.0, gk = 0.0;
    if (lf >= 0) { f0 = finF[lf * KF]; fk = kf >= 0 ? finF[lf * KF + kf] : 0.0; }
    if (lg >= 0) { g0 = finG[lg * KG]; gk = kg >= 0 ? finG[lg * KG + kg] : 0.0; }
    c0 = f0 * eg0 + g0 * ef0;
    ck = fk * eg0 + gk * ef0 + m * (f0 * egk + g0 * efk);
}

// The n slices of one branch (src/core.jl:121-128,178-185):
//   ℓ_i[γ] = ϕ_i ℓ_{i−1}[γ] + ψ_i Σ_t p_t ℓ_{i−1}[γ1] ℓ_{i−1}[γ2]
// `cur` holds row 1 on entry; rows alternate between fin and scr so that row n+1 lands in fin.
template <bool WARP>
__device__ __forceinline__ void run_slices(int n, int C, int K, double* fin, double* scr, double* cur,
                                           const Ent* s_dents, const uint32_t* s_dptr, uint32_t nd,
                                           const double2* __restrict__ pprow, double* prod, int cap, double* ellp,
                                           int tid, int nt) {
    const int GP = nt / K;
    const int grp = tid / K, k = tid - grp * K;
    const bool on = grp < GP && grp < C;
    const double m = k == 0 ? 0.0 : 1.0;
    uint32_t tb0 = 0, te0 = 0;  // the lane's first cell keeps its term range in registers
    if (on) { tb0 = s_dptr[grp]; te0 = s_dptr[grp + 1]; }
    double2 c0 = make_double2(0, 0), ck = c0;
    if (on && n >= 1) { c0 = __ldg(pprow + K); ck = __ldg(pprow + K + k); }
    for (int i = 1; i <= n; i++) {
        const double* src = cur;
        double* dst = (cur == fin) ? scr : fin;
        terms<false>(s_dents, 0, nd, src, K, nullptr, src, K, nullptr, K, prod, cap, tid, nt);
        double2 n0 = c0, nk = ck;  // prefetch the next slice's ϕ/ψ while this one is computed
        if (on && i < n) { n0 = __ldg(pprow + (size_t)(i + 1) * K); nk = __ldg(pprow + (size_t)(i + 1) * K + k); }
        scope_sync<WARP>();
        if (on)
            for (int c = grp; c < C; c += GP) {
                uint32_t tb = tb0, te = te0;
                if (c != grp) { tb = s_dptr[c]; te = s_dptr[c + 1]; }
                double s0, sk;
                cellsum(prod, cap, k, tb, te, s0, sk);
                const double o0 = src[c * K], ok = src[c * K + k];
                const double r = c0.x * ok + c0.y * sk + m * (ck.x * o0 + ck.y * s0);
                dst[c * K + k] = r;
                if (ellp && k == 0) ellp[(size_t)i * C + c] = r;
            }
        c0 = n0;
        ck = nk;
        cur = dst;
        scope_sync<WARP>();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fused slice loop for K <= 8 components (the common case: constant rates, or few branch-wise parameters).
// Each lane owns one Slot of the packer's lane table for the whole branch and keeps its (up to two) terms in
// registers, so a slice is: gather 2·K operands per term from the previous row, K fused multiply-adds per
// term, a shuffle reduction inside teams that share a heavy clade, the ϕ/ψ update, one barrier.
// ---------------------------------------------------------------------------------------------------------
constexpr int KMAX_FUSED = 8;

#ifdef WHALE_EMU
#define SHFL_DOWN(v, d) emu::shfl_down(v, d)
#else
#define SHFL_DOWN(v, d) __shfl_down_sync(0xffffffffu, v, d)
#endif

template <int K>
struct LaneWork {
    int cell, cnt, gsz, first;
    int a1, a2, b1, b2;  // row offsets (elements) of the operands of the lane's first two terms
    double pa, pb;
};

template <int K>
__device__ __forceinline__ LaneWork<K> load_work(const Slot* s_slots, int nslots, const Ent* s_dents, int sidx) {
    LaneWork<K> w;
    w.cell = -1; w.cnt = 0; w.gsz = 1; w.first = 0;
    w.a1 = w.a2 = w.b1 = w.b2 = 0;
    w.pa = w.pb = 0.0;
    if (sidx < nslots) {
        const Slot sl = s_slots[sidx];
        w.cell = sl.cell == 0xFFFFu ? -1 : (int)sl.cell;
        w.cnt = sl.cnt; w.gsz = 1 << sl.glog; w.first = sl.first;
        if (w.cnt > 0) { const Ent en = s_dents[w.first]; w.a1 = en.i1 * K; w.a2 = en.i2 * K; w.pa = en.p; }
        if (w.cnt > 1) { const Ent en = s_dents[w.first + w.gsz]; w.b1 = en.i1 * K; w.b2 = en.i2 * K; w.pb = en.p; }
    }
    return w;
}

template <int K>
__device__ __forceinline__ void accum(const double* __restrict__ src, int o1, int o2, double p, double (&s)[K]) {
    const double* x = src + o1;
    const double* y = src + o2;
    const double px = p * x[0], py = p * y[0];
    s[0] = fma(px, y[0], s[0]);
#pragma unroll
    for (int k = 1; k < K; k++) s[k] = fma(px, y[k], fma(py, x[k], s[k]));
}

// one pass of one slice for this lane; wg = team size of the warp's first lane (warp-uniform, 0: nothing to do)
template <int K>
__device__ __forceinline__ void slice_pass(const LaneWork<K>& w, int wg, int sidx, const double* __restrict__ src,
                                           double* __restrict__ dst, const double2* ppi, const Ent* s_dents, int C,
                                           int i, double* ellp) {
    if (wg == 0) return;
    double s[K];
#pragma unroll
    for (int k = 0; k < K; k++) s[k] = 0.0;
    if (w.cnt > 0) accum<K>(src, w.a1, w.a2, w.pa, s);
    if (w.cnt > 1) accum<K>(src, w.b1, w.b2, w.pb, s);
    for (int j = 2; j < w.cnt; j++) {
        const Ent en = s_dents[w.first + j * w.gsz];
        accum<K>(src, en.i1 * K, en.i2 * K, en.p, s);
    }
    for (int step = 1; step < wg; step <<= 1) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            const double t = SHFL_DOWN(s[k], step);
            if (step < w.gsz) s[k] += t;
        }
    }
    if (w.cell >= 0 && (sidx & (w.gsz - 1)) == 0) {  // team leader: ℓ_i = ϕ_i ℓ_{i−1} + ψ_i Σ  (with tangents)
        const int c = w.cell;
        const double2 c0 = ppi[0];
        const double o0 = src[c * K];
        const double r0 = fma(c0.x, o0, c0.y * s[0]);
        dst[c * K] = r0;
        if (ellp) ellp[(size_t)i * C + c] = r0;
#pragma unroll
        for (int k = 1; k < K; k++) {
            const double2 ck = ppi[k];
            dst[c * K + k] = fma(c0.x, src[c * K + k], fma(c0.y, s[k], fma(ck.x, o0, ck.y * s[0])));
        }
    }
}

// Not inlined on purpose: each component count gets its own register allocation; inlining all variants into
// k_dp makes ptxas demote the per-lane accumulator arrays to local memory.
template <int K, bool WARP>
__device__ __noinline__ void run_slices_fused(int n, int C, double* fin, double* scr, double* cur,
                                                 const Slot* s_slots, int nslots, const Ent* s_dents,
                                                 const double2* pprow, double* ellp, int tid, int nt) {
    // the first two passes keep their descriptors in registers
    const int wbase = tid & ~31;
    const LaneWork<K> w0 = load_work<K>(s_slots, nslots, s_dents, tid);
    const LaneWork<K> w1 = load_work<K>(s_slots, nslots, s_dents, tid + nt);
    // slots are sorted by team size (descending): the warp's first lane carries the warp's largest team
    const int wg0 = wbase < nslots ? (1 << s_slots[wbase].glog) : 0;
    const int wg1 = (wbase + nt) < nslots ? (1 << s_slots[wbase + nt].glog) : 0;
    const int npass = (nslots + nt - 1) / nt;
    // warp scope: ϕ/ψ rows come from global memory -> keep the current row in registers and fetch the next
    // one while the slice is computed; block scope: rows were staged in shared memory
    double2 pc[K], pn[K];
    if (WARP) {
#pragma unroll
        for (int k = 0; k < K; k++) pn[k] = n >= 1 ? __ldg(pprow + K + k) : make_double2(0.0, 0.0);
    }
    for (int i = 1; i <= n; i++) {
        const double* src = cur;
        double* dst = (cur == fin) ? scr : fin;
        const double2* ppi = pprow + (size_t)i * K;
        if (WARP) {
#pragma unroll
            for (int k = 0; k < K; k++) pc[k] = pn[k];
            if (i < n) {
#pragma unroll
                for (int k = 0; k < K; k++) pn[k] = __ldg(pprow + (size_t)(i + 1) * K + k);
            }
            ppi = pc;
        }
        slice_pass<K>(w0, wg0, tid, src, dst, ppi, s_dents, C, i, ellp);
        slice_pass<K>(w1, wg1, tid + nt, src, dst, ppi, s_dents, C, i, ellp);
        for (int q = 2; q < npass; q++) {  // oversized rows: descriptors reloaded from shared memory
            const LaneWork<K> wq = load_work<K>(s_slots, nslots, s_dents, tid + q * nt);
            const int wgq = (wbase + q * nt) < nslots ? (1 << s_slots[wbase + q * nt].glog) : 0;
            slice_pass<K>(wq, wgq, tid + q * nt, src, dst, ppi, s_dents, C, i, ellp);
        }
        cur = dst;
        scope_sync<WARP>();
    }
}

// dispatch on the branch's component count (leaf branches carry at most value + own λ, μ: K <= 3)
template <bool WARP, int KCAP>
__device__ __forceinline__ bool run_slices_fused_k(int K, int n, int C, double* fin, double* scr, double* cur,
                                                   const Slot* s_slots, int nslots, const Ent* s_dents,
                                                   const double2* pprow, double* ellp, int tid, int nt) {
#define CASEK(KK) case KK: run_slices_fused<KK, WARP>(n, C, fin, scr, cur, s_slots, nslots, s_dents, pprow, ellp, tid, nt); return true;
    if (WARP) {
        switch (K) {
            CASEK(1) CASEK(2) CASEK(3)
            default: return false;
        }
    } else {
        if (KCAP <= 6) {
            switch (K) {
                CASEK(1) CASEK(2) CASEK(3) CASEK(4) CASEK(5) CASEK(6)
                default: return false;
            }
        } else {
            switch (K) {
                CASEK(1) CASEK(2) CASEK(3) CASEK(4) CASEK(5) CASEK(6) CASEK(7) CASEK(8)
                default: return false;
            }
        }
    }
#undef CASEK
}

// KCAP: largest component count with a fused slice loop compiled in (6 or 8); plans with larger K use the
// generic two-phase slice loop
template <int NT, int MINB, int KCAP>
__global__ void __launch_bounds__(NT, MINB) k_dp(DPArgs A, int perm_off) {
    EXTERN_SHARED(smem_raw);
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ModelDev& M = A.M;
    const PlanDev& PL = A.PL;
    const int nn = M.nn, Kmax = PL.Kmax;
    const int fam = A.perm[perm_off + blockIdx.x];
    const FamHdr* Hp = A.hdr + fam;
    const uint64_t base = Hp->base;
    const uint32_t nlev = Hp->nlev, blob_bytes = Hp->blob_bytes;
    const uint32_t rows_len = Hp->rows_len[A.plan], scr_len = Hp->scr_len[A.plan], prod_len = Hp->prod_len[A.plan];
    const uint32_t leafmax = Hp->leafmax[A.plan];
    const uint32_t stage_bytes = Hp->stage_bytes[A.plan], leaf_stage = Hp->leaf_stage;
    const unsigned char* blob = A.arena + base;
    const NodeRec* nrec = reinterpret_cast<const NodeRec*>(blob);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(blob);
    const Ent* ents = reinterpret_cast<const Ent*>(blob);

    long long tc0 = CLOCK64(), tc1 = 0, acc_row1 = 0, acc_slices = 0, acc_stage = 0;
    // pull the whole blob towards L2 now; it is consumed node by node below
    for (uint32_t o = tid * 128u; o < blob_bytes; o += NT * 128u) PREFETCH_L2(blob + o);

    // ---- shared memory carve-up (mirrored by smem_need() on the host) ----
    // species-tree metadata (one coalesced read instead of dependent global loads at every node)
    int* s_kind = reinterpret_cast<int*>(smem_raw);
    int* s_nsl = s_kind + nn;
    int* s_ch0 = s_nsl + nn;
    int* s_ch1 = s_ch0 + nn;
    int* s_K = s_ch1 + nn;
    int* s_toff = s_K + nn;
    int* s_roff = s_toff + nn;                                           // [nn+1] row offsets (doubles)
    int16_t* s_cmap = reinterpret_cast<int16_t*>(s_roff + nn + 1);       // [nn*2*Kmax]
    const size_t hdr_bytes = (((7 * nn + 1) * sizeof(int) + (size_t)nn * 2 * Kmax * sizeof(int16_t)) + 15) & ~size_t(15);
    double* rows = reinterpret_cast<double*>(smem_raw + hdr_bytes);
    double* scr = rows + rows_len;
    double* prod = scr + scr_len;
    unsigned char* stage = reinterpret_cast<unsigned char*>(prod + prod_len);
    unsigned char* leaf_area = stage + stage_bytes;
    const size_t leaf_area_bytes = (size_t)leafmax * sizeof(double) + leaf_stage;
    for (int i = tid; i < nn; i += NT) {
        s_kind[i] = M.kind[i]; s_nsl[i] = M.nsl[i]; s_ch0[i] = M.child0[i]; s_ch1[i] = M.child1[i];
        s_K[i] = PL.K[i]; s_toff[i] = PL.toff[i];
    }
    for (int i = tid; i < nn * 2 * Kmax; i += NT) s_cmap[i] = PL.cmap[i];
    __syncthreads();
    for (int i = tid; i < nn; i += NT) s_roff[i] = (int)A.roff[(size_t)fam * nn + i];
    __syncthreads();
    double* const ell_base = A.ell ? A.ell + Hp->ell_off : nullptr;
    auto ell_of = [&](int e) -> double* {  // node e's matrix inside the family's ℓ (node-index order)
        if (!ell_base) return nullptr;
        size_t o = 0;
        for (int e2 = 0; e2 < e; e2++) o += (size_t)(s_nsl[e2] + 1) * nrec[e2].C;
        return ell_base + o;
    };

    const long long tcA = CLOCK64();
    // ================= phase A: leaf branches, one warp each (src/core.jl:83-101,121-128) =================
    for (int li = warp; li < M.nleafnodes; li += NW) {
        const int e = M.leafnodes[li];
        const NodeRec R = nrec[e];
        const int C = (int)R.C;
        if (C == 0) continue;
        const int K = s_K[e], n = s_nsl[e];
        double* fin = rows + s_roff[e];
        if (A.skip_leaf && R.sptr_off) {
            // closed form over tree shapes: ℓ_n[γ] = Σ_σ C_σ[γ]·wσ_n  (C from the packer, w from k_leafshapes)
            const uint32_t* sptr = words + R.sptr_off;
            const Ent* sent = ents + R.sent_off;
            const double* W = PL.shapeW + (size_t)e * NSHAPE * Kmax;
            for (int i = lane; i < C * K; i += 32) {
                const int c = i / K, k = i - c * K;
                double v = 0.0;
                for (uint32_t t = sptr[c]; t < sptr[c + 1]; t++) {
                    const Ent en = sent[t];
                    v = fma(en.p, W[en.i1 * Kmax + k], v);
                }
                fin[i] = v;
            }
            continue;
        }
        if (R.nslots > 32 && !(R.nonleaf == 0 && A.skip_leaf)) continue;  // heavy branch: whole CTA, below
        if (R.nonleaf == 0 && A.skip_leaf) {  // family-independent: ℓ_n = leafℙ·Πϕ_i from k_tables
            for (int i = lane; i < C * K; i += 32) fin[i] = PL.leaf[e * Kmax + (i % K)];
            continue;
        }
        double* ellp = ell_of(e);
        double* wscr = reinterpret_cast<double*>(leaf_area + warp * leaf_area_bytes);
        uint4* wst = reinterpret_cast<uint4*>(wscr + leafmax);
        const int nd16 = (int)R.ndent, sl16 = ((int)R.nslots + 1) >> 1;
        copy16(wst, reinterpret_cast<const uint4*>(ents + R.dent_off), nd16, lane, 32);
        copy16(wst + nd16, reinterpret_cast<const uint4*>(words + R.slot_off), sl16, lane, 32);
        double* cur = (n & 1) ? wscr : fin;  // row i lives in fin iff (n − i) is even
        const int nleafc = C - (int)R.nonleaf;
        for (int i = lane; i < C * K; i += 32) {
            const int c = i / K, k = i - c * K;
            const double v = (c < nleafc && k == 0) ? M.leafP[e] : 0.0;
            cur[i] = v;
            if (ellp && k == 0) ellp[c] = v;
        }
        stage_wait();
        __syncwarp();
        run_slices_fused_k<true, KCAP>(K, n, C, fin, wscr, cur, reinterpret_cast<const Slot*>(wst + nd16), (int)R.nslots,
                                 reinterpret_cast<const Ent*>(wst), PL.pp + s_toff[e], ellp, lane, 32);
    }
    __syncthreads();
    // leaf branches with many in-paralog clades (more lanes of work than one warp): all warps cooperate
    for (int li = 0; li < M.nleafnodes; li++) {
        const int e = M.leafnodes[li];
        const NodeRec R = nrec[e];
        const int C = (int)R.C;
        if (C == 0 || R.nslots <= 32 || (R.nonleaf == 0 && A.skip_leaf) || (A.skip_leaf && R.sptr_off)) continue;
        const int K = s_K[e], n = s_nsl[e];
        double* fin = rows + s_roff[e];
        double* ellp = ell_of(e);
        const int nd16 = (int)R.ndent, sl16 = ((int)R.nslots + 1) >> 1, pp16 = (n + 1) * K;
        uint4* st4 = reinterpret_cast<uint4*>(stage);
        copy16(st4, reinterpret_cast<const uint4*>(ents + R.dent_off), nd16, tid, NT);
        copy16(st4 + nd16, reinterpret_cast<const uint4*>(words + R.slot_off), sl16, tid, NT);
        copy16(st4 + nd16 + sl16, reinterpret_cast<const uint4*>(PL.pp + s_toff[e]), pp16, tid, NT);
        double* cur = (n & 1) ? scr : fin;
        const int nleafc = C - (int)R.nonleaf;
        for (int i = tid; i < C * K; i += NT) {
            const int c = i / K, k = i - c * K;
            const double v = (c < nleafc && k == 0) ? M.leafP[e] : 0.0;
            cur[i] = v;
            if (ellp && k == 0) ellp[c] = v;
        }
        stage_wait();
        __syncthreads();
        run_slices_fused_k<false, KCAP>(K, n, C, fin, scr, cur, reinterpret_cast<const Slot*>(st4 + nd16), (int)R.nslots,
                                  reinterpret_cast<const Ent*>(st4), reinterpret_cast<const double2*>(st4 + nd16 + sl16),
                                  ellp, tid, NT);
    }
    const long long tcB = CLOCK64();

    // ================= phase B: internal, WGD and root nodes, whole CTA =================
    for (int oi = 0; oi < M.ninner; oi++) {
        const int e = M.inner[oi];
        const NodeRec R = nrec[e];
        const int C = (int)R.C;
        if (C == 0) continue;
        const int kind = s_kind[e], K = s_K[e], n = s_nsl[e];
        double* fin = rows + s_roff[e];
        double* ellp = ell_of(e);
        const int cap = (int)(prod_len / K);
        const bool fused = K <= KCAP;
        // lane -> (cell group, component) for the P2 passes of row 1
        const int GP = NT / K;
        const int grp = tid / K, k = tid - grp * K;
        const bool on = grp < GP;
        const double m = k == 0 ? 0.0 : 1.0;

        const long long tst = CLOCK64();
        // ---- stage this node's lists: [dents | slots | dptr | tptr,lossF,lossG,lev | ϕψ rows] ----
        const int nd16 = (kind == WHALE_ROOT) ? 0 : (int)R.ndent;  // Πroot terms are read once: stay global
        const int sl16 = (kind == WHALE_ROOT) ? 0 : (((int)R.nslots + 1) >> 1);
        const int dp16 = (C + 1 + 3) >> 2;
        const int tp16 = (kind == WHALE_WGD) ? 0 : ((3 * C + 1 + (kind == WHALE_ROOT ? (int)nlev + 1 : 0) + 3) >> 2);
        const int pp16 = fused ? (n + 1) * K : 0;
        uint4* st4 = reinterpret_cast<uint4*>(stage);
        copy16(st4, reinterpret_cast<const uint4*>(ents + R.dent_off), nd16, tid, NT);
        copy16(st4 + nd16, reinterpret_cast<const uint4*>(words + R.slot_off), sl16, tid, NT);
        copy16(st4 + nd16 + sl16, reinterpret_cast<const uint4*>(words + R.dptr_off), dp16, tid, NT);
        copy16(st4 + nd16 + sl16 + dp16, reinterpret_cast<const uint4*>(words + R.tptr_off), tp16, tid, NT);
        copy16(st4 + nd16 + sl16 + dp16 + tp16, reinterpret_cast<const uint4*>(PL.pp + s_toff[e]), pp16, tid, NT);
        stage_wait();
        const Ent* s_dents = reinterpret_cast<const Ent*>(stage);
        const Slot* s_slots = reinterpret_cast<const Slot*>(st4 + nd16);
        const uint32_t* s_dptr = reinterpret_cast<const uint32_t*>(st4 + nd16 + sl16);
        const uint32_t* s_tptr = reinterpret_cast<const uint32_t*>(st4 + nd16 + sl16 + dp16);
        const int32_t* s_lossF = reinterpret_cast<const int32_t*>(s_tptr + C + 1);
        const int32_t* s_lossG = s_lossF + C;
        const uint32_t* s_lev = reinterpret_cast<const uint32_t*>(s_lossG + C);
        const double2* s_pp = reinterpret_cast<const double2*>(st4 + nd16 + sl16 + dp16 + tp16);
        auto slices = [&](double* cur) {
            const long long ts = CLOCK64();
            acc_row1 += ts - tc1;
            if (!run_slices_fused_k<false, KCAP>(K, n, C, fin, scr, cur, s_slots, (int)R.nslots, s_dents, s_pp, ellp, tid, NT))
                run_slices<false>(n, C, K, fin, scr, cur, s_dents, s_dptr, R.ndent, PL.pp + s_toff[e], prod, cap, ellp,
                                  tid, NT);
            acc_slices += CLOCK64() - ts;
        };

        // children, shared by the row-1 formulas
        tc1 = CLOCK64();
        acc_stage += tc1 - tst;
        const int f = s_ch0[e], g = s_ch1[e];
        const int KF = s_K[f];
        const double* finF = rows + s_roff[f];
        const int16_t* mapF = s_cmap + (e * 2 + 0) * Kmax;
        const int16_t* mapG = s_cmap + (e * 2 + 1) * Kmax;
        const int kf = mapF[k];
        double* cur = (n & 1) ? scr : fin;  // row i lives in fin iff (n − i) is even

        if (kind == WHALE_WGD) {
            // q·Σ p ℓ_f[γ1]ℓ_f[γ2] + (1−q+2qϵ_f)·ℓ_f[γ]   src/core.jl:103-119,187-199
            __syncthreads();
            const double cx0 = PL.cx[e * Kmax], cy0 = PL.cy[e * Kmax];
            const double cxk = PL.cx[e * Kmax + k], cyk = PL.cy[e * Kmax + k];
            // retention terms in windows of `cap` products, whole cells at a time
            int cA = 0;
            while (cA < C) {
                const uint32_t ta = s_dptr[cA];
                int cB = cA;
                while (cB < C && s_dptr[cB + 1] - ta <= (uint32_t)cap) cB++;
                const bool big = cB == cA;  // a single cell with more than `cap` terms: serial sum
                if (big) cB = cA + 1;
                else terms<false>(s_dents, ta, s_dptr[cB], finF, KF, mapF, finF, KF, mapF, K, prod, cap, tid, NT);
                __syncthreads();
                if (on)
                    for (int c = cA + grp; c < cB; c += GP) {
                        double s0 = 0.0, sk = 0.0;
                        if (!big) cellsum(prod, cap, k, s_dptr[c] - ta, s_dptr[c + 1] - ta, s0, sk);
                        else {
                            for (uint32_t t = s_dptr[c]; t < s_dptr[c + 1]; t++) {
                                const Ent en = s_dents[t];
                                const double* xp = finF + en.i1 * KF;
                                const double* yp = finF + en.i2 * KF;
                                const double xv = kf >= 0 ? xp[kf] : 0.0, yv = kf >= 0 ? yp[kf] : 0.0;
                                s0 = fma(en.p * xp[0], yp[0], s0);
                                sk += fma(en.p * xp[0], yv, (en.p * yp[0]) * xv);
                            }
                            if (k == 0) sk = s0;
                        }
                        const double u0 = finF[c * KF];
                        const double uk = kf >= 0 ? finF[c * KF + kf] : 0.0;
                        const double r = cy0 * sk + cx0 * uk + m * (cyk * s0 + cxk * u0);
                        cur[c * K + k] = r;
                        if (ellp && k == 0) ellp[c] = r;
                    }
                __syncthreads();
                cA = cB;
            }
            slices(cur);
            continue;
        }

        // internal node or root: speciation + loss from the children's last rows (src/core.jl:160-176)
        const int KG = s_K[g];
        const double* finG = rows + s_roff[g];
        const int kg = mapG[k];
        const double* epsF = PL.eps + s_toff[f] + (size_t)s_nsl[f] * KF;
        const double* epsG = PL.eps + s_toff[g] + (size_t)s_nsl[g] * KG;
        const double ef0 = epsF[0], eg0 = epsG[0];
        const double efk = (k > 0 && kf >= 0) ? epsF[kf] : 0.0;
        const double egk = (k > 0 && kg >= 0) ? epsG[kg] : 0.0;
        const Ent* g_tents = ents + R.tent_off;
        __syncthreads();  // staged lists visible
        if (kind == WHALE_INTERNAL) {
            // speciation terms may exceed the product window: process them in windows of `cap` terms, whole
            // cells at a time (a single cell with more than `cap` terms falls back to a serial sum)
            int cA = 0;
            while (cA < C) {
                const uint32_t ta = s_tptr[cA];
                int cB = cA;
                while (cB < C && s_tptr[cB + 1] - ta <= (uint32_t)cap) cB++;
                if (cB == cA) {
                    if (tid < K) {
                        double s0 = 0.0, sk = 0.0;
                        for (uint32_t t = ta; t < s_tptr[cA + 1]; t++) {
                            const Ent en = g_tents[t];
                            const double* xp = finF + en.i1 * KF;
                            const double* yp = finG + en.i2 * KG;
                            const double xv = kf >= 0 ? xp[kf] : 0.0, yv = kg >= 0 ? yp[kg] : 0.0;
                            s0 = fma(en.p * xp[0], yp[0], s0);
                            sk += fma(en.p * xp[0], yv, (en.p * yp[0]) * xv);
                        }
                        double l0, lk;
                        loss_term(s_lossF[cA], s_lossG[cA], finF, KF, kf, finG, KG, kg, ef0, efk, eg0, egk, m, l0, lk);
                        const double r = k == 0 ? s0 + l0 : sk + lk;
                        cur[cA * K + k] = r;
                        if (ellp && k == 0) ellp[cA] = r;
                    }
                    cB = cA + 1;
                } else {
                    terms<true>(g_tents, ta, s_tptr[cB], finF, KF, mapF, finG, KG, mapG, K, prod, cap, tid, NT);
                    __syncthreads();
                    if (on)
                        for (int c = cA + grp; c < cB; c += GP) {
                            double b0, bk, l0, lk;
                            cellsum(prod, cap, k, s_tptr[c] - ta, s_tptr[c + 1] - ta, b0, bk);
                            loss_term(s_lossF[c], s_lossG[c], finF, KF, kf, finG, KG, kg, ef0, efk, eg0, egk, m, l0, lk);
                            const double r = k == 0 ? b0 + l0 : bk + lk;
                            cur[c * K + k] = r;
                            if (ellp && k == 0) ellp[c] = r;
                        }
                }
                __syncthreads();
                cA = cB;
            }
            slices(cur);
            continue;
        }

        // ---- root: ℓ_r[γ] = (1−η)ξ/η·a + η(1−ϵ)/ξ²·(b + c),  a = Σ p ℓ_r[γ1]ℓ_r[γ2] over the SAME row
        //      (src/core.jl:130-158) => clades in ascending size, one level (= clade size) at a time.
        //      Per level P1 covers the level's Πroot terms (window [0, na)) and speciation terms ([na, na+nb)).
        const double cx0 = PL.cx[e * Kmax], cy0 = PL.cy[e * Kmax];
        const double cxk = PL.cx[e * Kmax + k], cyk = PL.cy[e * Kmax + k];
        const Ent* g_dents = ents + R.dent_off;
        // the root works in place, so its product window spans the (idle) scratch row and the product window
        double* const rprod = scr;
        const int rcap = (int)((scr_len + prod_len) / K);
        for (uint32_t L = 0; L < nlev; L++) {
            const int c0 = (int)s_lev[L], c1 = (int)s_lev[L + 1];
            const uint32_t ta = s_dptr[c0], na = s_dptr[c1] - ta;
            const uint32_t ua = s_tptr[c0], nb = s_tptr[c1] - ua;
            const bool fits = na + nb <= (uint32_t)rcap;
            if (fits) {
                terms<true>(g_dents, ta, ta + na, fin, K, nullptr, fin, K, nullptr, K, rprod, rcap, tid, NT);
                terms<true>(g_tents, ua, ua + nb, finF, KF, mapF, finG, KG, mapG, K, rprod + na, rcap, tid, NT);
                __syncthreads();
            }
            if (on)
                for (int c = c0 + grp; c < c1; c += GP) {
                    double a0 = 0.0, ak = 0.0, b0 = 0.0, bk = 0.0, l0, lk;
                    if (fits) {
                        cellsum(rprod, rcap, k, s_dptr[c] - ta, s_dptr[c + 1] - ta, a0, ak);
                        cellsum(rprod, rcap, k, na + s_tptr[c] - ua, na + s_tptr[c + 1] - ua, b0, bk);
                    } else {  // oversized level: serial sums per lane
                        for (uint32_t t = s_dptr[c]; t < s_dptr[c + 1]; t++) {
                            const Ent en = g_dents[t];
                            const double* xp = fin + en.i1 * K;
                            const double* yp = fin + en.i2 * K;
                            a0 = fma(en.p * xp[0], yp[0], a0);
                            ak += fma(en.p * xp[0], yp[k], (en.p * yp[0]) * xp[k]);
                        }
                        for (uint32_t t = s_tptr[c]; t < s_tptr[c + 1]; t++) {
                            const Ent en = g_tents[t];
                            const double* xp = finF + en.i1 * KF;
                            const double* yp = finG + en.i2 * KG;
                            const double xv = kf >= 0 ? xp[kf] : 0.0, yv = kg >= 0 ? yp[kg] : 0.0;
                            b0 = fma(en.p * xp[0], yp[0], b0);
                            bk += fma(en.p * xp[0], yv, (en.p * yp[0]) * xv);
                        }
                        if (k == 0) { ak = a0; bk = b0; }
                    }
                    loss_term(s_lossF[c], s_lossG[c], finF, KF, kf, finG, KG, kg, ef0, efk, eg0, egk, m, l0, lk);
                    const double u0 = b0 + l0, uk = k == 0 ? u0 : bk + lk;
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
        if (A.tim && tid == 0) {
            const long long te = CLOCK64();
            long long* T = A.tim + (size_t)fam * 8;
            T[0] = tcA - tc0;            // prologue
            T[1] = tcB - tcA;            // leaf phase (incl. waiting for the slowest warp)
            T[2] = acc_stage;            // staging copies issued (phase B)
            T[3] = acc_row1;             // row 1 of internal/WGD nodes (incl. staging wait)
            T[4] = acc_slices;           // slices of internal/WGD nodes
            T[5] = te - tc1;             // root
            T[6] = te - tc0;             // total
            T[7] = 0;
        }
    }
}
