// K2r — the ALE recursion differentiated in REVERSE mode (adjoints).  src/core.jl:83-199 and, transposed, the same.
//
// What it replaces: ForwardDiff.gradient over logpdf (test/runtests.jl:36-42).  k_dp (whale_dp.cuh) carries K_e
// forward tangents per ℓ cell, so its cost grows with the number of parameters and branch-wise rates need several
// passes over parameter chunks.  Here the DP itself sees no parameter at all:
//   forward   value-only recursion; every row of every internal/WGD branch is kept in a per-CTA history buffer in
//             global memory (rewritten by every family the CTA processes, so it lives in L2).  Leaf branches keep
//             the K <= 3 forward tangents of their own λ, μ (they have no children: 82 % of all slices stay as
//             cheap as before, incl. the family-independent rows and the tree-shape closed form).
//   backward  ℓ̄ rows from the root down with the TRANSPOSED term lists (RevRec): for a slice
//                 ℓ̄_{i−1}[γ] = ϕ_i ℓ̄_i[γ] + ψ_i Σ_{(π,σ,p) ∈ T'(γ)} p·ℓ̄_i[π]·ℓ_{i−1}[σ]
//             — the same gather / team-reduce structure as the forward slice loop, reading the kept row i−1 (staged
//             three rows ahead with cp.async).  The adjoints of the slice tables are folded on the fly into the
//             three LOCAL parameters of the branch:  with w_ϕ = ℓ̄_i[γ]ℓ_{i−1}[γ] and w_ψ = ½ℓ_{i−1}[γ]·Σ_{T'(γ)} (each
//             forward term appears under both of its operands, hence the ½) every leader lane accumulates
//                 (λ̄_e, μ̄_e, ϵ̄⁰_e) += w_ϕ·∂ϕ_i/∂(λ_e, μ_e, ϵ⁰_e) + w_ψ·∂ψ_i/∂(λ_e, μ_e, ϵ⁰_e)
//             from the local-tangent table of k_tables3 (ϕ_i, ψ_i depend on nothing else), one block reduction per
//             branch.  Row 1 of internal / WGD nodes and the root are transposed the same way and add the adjoints of
//             the children's last ϵ and of the row-1 coefficients.
//   contract  ∂ log L_f/∂θ_k = Σ_e λ̄_e·[k = λ_e] + μ̄_e·[k = μ_e] + ϵ̄⁰_e·∂ϵ⁰_e/∂θ_k + ϵ̄ⁿ_c·∂ϵⁿ_c/∂θ_k + c̄x·∂cx/∂θ_k + c̄y·∂cy/∂θ_k
//             + (leaf branches) Σ_γ ℓ̄_n[γ]·∂ℓ_n[γ]/∂θ_k, with the GLOBAL tangents of the full plan's tables.
// The output has the layout of k_dp's full plan ([F][K_root]: log L_f, ∂ log L_f/∂ component), so the fused
// reduction, the mixture kernels and the per-family read-back are shared.  Cost: 2 gathers per term and slice forward
// + 4 backward, whatever P is (forward tangents: 2·K_e).
//
// Launch: persistent CTAs (as many as fit the GPU), each pulling families from the launch order with an atomic
// counter — no tail wave, species-tree metadata read once per CTA, one history slot per CTA.
#pragma once
#include "whale_dp.cuh"

struct RevArgs {
    ModelDev M;
    PlanDev PR;  // hybrid plan the DP runs on: leaf branches K <= 3 (own λ, μ), every other node K = 1
    PlanDev PG;  // full plan: global tangents of ϵ rows, cx, cy, condition
    PlanDev PL;  // local plan: K = 4 on internal/WGD branches — (ϕ_i, ψ_i) and ∂/∂(own λ, own μ, ϵ_0)
    const unsigned char* arena;
    const FamHdr* hdr;
    const unsigned char* rarena;  // reverse blobs
    const RevHdr* rhdr;
    const int* perm;
    const uint32_t* roff;  // [F][nn] last-row offsets (doubles) in `rows`
    const uint32_t* aoff;  // [F][nn] adjoint-row offsets (doubles) in `arows`
    double* out_fam;       // [F * KR]
    double* hist;          // per-CTA history slots
    unsigned long long hist_stride;  // doubles per slot
    unsigned int* next;    // family counter of this launch (zeroed by the host)
    int bin_off, bin_count, slot0;
    long long* tim;
    unsigned int* done;
    int n_total, cond_kind;
    double* out;
    PeerDev* peer;  // fused exchange over ranks behind the fused reduction (nullptr: off)
};

// Σ_t p_t·X[i1·sx]·Y[i2·sy] over entries tb, tb+step, ... < te; two entries in flight
template <bool GLOBAL_ENTS>
__device__ __forceinline__ double vsum(const Ent* __restrict__ ents, uint32_t tb, uint32_t te, uint32_t step,
                                       const double* __restrict__ X, int sx, const double* __restrict__ Y, int sy) {
    double a = 0.0, b = 0.0;
    uint32_t t = tb;
    for (; t + step < te; t += 2 * step) {
        uint4 r1, r2;
        if (GLOBAL_ENTS) { r1 = __ldg(reinterpret_cast<const uint4*>(ents + t)); r2 = __ldg(reinterpret_cast<const uint4*>(ents + t + step)); }
        else { r1 = *reinterpret_cast<const uint4*>(ents + t); r2 = *reinterpret_cast<const uint4*>(ents + t + step); }
        const double x1 = X[(r1.x & 0xffffu) * sx], y1 = Y[(r1.x >> 16) * sy];
        const double x2 = X[(r2.x & 0xffffu) * sx], y2 = Y[(r2.x >> 16) * sy];
        const double p1 = __hiloint2double((int)r1.w, (int)r1.z), p2 = __hiloint2double((int)r2.w, (int)r2.z);
        a = fma(p1 * x1, y1, a);
        b = fma(p2 * x2, y2, b);
    }
    if (t < te) {
        uint4 r1;
        if (GLOBAL_ENTS) r1 = __ldg(reinterpret_cast<const uint4*>(ents + t));
        else r1 = *reinterpret_cast<const uint4*>(ents + t);
        const double x1 = X[(r1.x & 0xffffu) * sx], y1 = Y[(r1.x >> 16) * sy];
        const double p1 = __hiloint2double((int)r1.w, (int)r1.z);
        a = fma(p1 * x1, y1, a);
    }
    return a + b;
}

// deterministic block sum of N per-thread scalars: shuffle tree inside each warp, then thread j adds the warps' partial
// sums in warp order; dst[j] (+)= the total.  Every thread of the CTA must call it.
template <int N, int NT>
__device__ __forceinline__ void block_sum(double (&v)[N], double* s_red, double* dst, bool accumulate, int tid) {
    constexpr int NW = NT / 32;
#pragma unroll
    for (int j = 0; j < N; j++)
        for (int step = 16; step > 0; step >>= 1) v[j] += SHFL_DOWN(v[j], step);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int j = 0; j < N; j++) s_red[(tid >> 5) * N + j] = v[j];
    }
    __syncthreads();
    if (tid < N) {
        double s = 0.0;
        for (int w = 0; w < NW; w++) s += s_red[w * N + tid];
        dst[tid] = accumulate ? dst[tid] + s : s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// Slice loops of the reverse-mode kernel (K = 1: one double per cell).  A thread owns up to two slots of the packer's
// lane table per slice ("passes" 0 and 1, descriptors in registers); both passes are carried through the gathers, the
// team reduction and the leader's update TOGETHER, so their shuffle chains overlap instead of running back to back
// (the slice is a latency chain, not a throughput problem: profiles/r2_ncu_k_dp_rev_v1_*).
// ---------------------------------------------------------------------------------------------------------
struct Pass2 {
    LaneWork<1> w0, w1;
    int wg0, wg1;          // team size of the warp's first lane in each pass (0: the warp owns no slot of that pass)
    bool two0, more0, two1, more1;
    bool lead0, lead1;
    int c0, c1;
};
__device__ __forceinline__ Pass2 load_pass2(const Slot* s_slots, int nslots, const Ent* s_ents, int tid, int nt) {
    Pass2 P;
    const int wbase = tid & ~31;
    P.w0 = load_work<1>(s_slots, nslots, s_ents, tid);
    P.w1 = load_work<1>(s_slots, nslots, s_ents, tid + nt);
    P.wg0 = wbase < nslots ? (1 << s_slots[wbase].glog) : 0;
    P.wg1 = (wbase + nt) < nslots ? (1 << s_slots[wbase + nt].glog) : 0;
    P.two0 = WARP_ANY(P.w0.cnt > 1); P.more0 = WARP_ANY(P.w0.cnt > 2);
    P.two1 = WARP_ANY(P.w1.cnt > 1); P.more1 = WARP_ANY(P.w1.cnt > 2);
    P.lead0 = P.w0.cell >= 0 && (tid & (P.w0.gsz - 1)) == 0;
    P.lead1 = P.w1.cell >= 0 && ((tid + nt) & (P.w1.gsz - 1)) == 0;
    P.c0 = max(P.w0.cell, 0); P.c1 = max(P.w1.cell, 0);
    return P;
}
// Σ over the lane's share of one cell's terms: p·X[i1]·Y[i2] (the first two terms from registers)
__device__ __forceinline__ double lane_terms(const LaneWork<1>& w, bool two, bool more, const double* __restrict__ X,
                                             const double* __restrict__ Y, const Ent* s_ents) {
    double s;
    if (two) s = fma(w.pb * X[w.b1], Y[w.b2], (w.pa * X[w.a1]) * Y[w.a2]);
    else s = (w.pa * X[w.a1]) * Y[w.a2];
    if (more) {
#pragma unroll 1
        for (int j = 2; j < w.cnt; j++) {
            const Ent en = s_ents[w.first + j * w.gsz];
            s = fma(en.p * X[en.i1], Y[en.i2], s);
        }
    }
    return s;
}
// team reduction of both passes' partial sums, interleaved (lane j adds lane j+step while step is inside its own team;
// the mask multiplies instead of selecting: fma(t, 1, s) is the correctly rounded s + t)
__device__ __forceinline__ void team_reduce2(const Pass2& P, double& s0, double& s1) {
    for (int step = 1; step < P.wg0; step <<= 1) {  // slots are sorted by team size: wg1 <= wg0
        const double m0 = step < P.w0.gsz ? 1.0 : 0.0, m1 = step < P.w1.gsz ? 1.0 : 0.0;
        const double t0 = SHFL_DOWN(s0, step), t1 = SHFL_DOWN(s1, step);
        s0 = fma(t0, m0, s0);
        s1 = fma(t1, m1, s1);
    }
}
__device__ __forceinline__ void team_reduce1(const LaneWork<1>& w, int wg, double& s) {
    for (int step = 1; step < wg; step <<= 1) {
        const double m = step < w.gsz ? 1.0 : 0.0;
        s = fma(SHFL_DOWN(s, step), m, s);
    }
}

// coalesced copy of one finished row (Cp doubles, 16-byte aligned on both sides) from shared memory to the history
__device__ __forceinline__ void keep_row(double* __restrict__ g, const double* __restrict__ srow, int Cp, int tid, int nt) {
#pragma unroll 1
    for (int c = tid; c < (Cp >> 1); c += nt) reinterpret_cast<double2*>(g)[c] = reinterpret_cast<const double2*>(srow)[c];
}

// Forward: ℓ_i[γ] = ϕ_i ℓ_{i−1}[γ] + ψ_i Σ_t p_t ℓ_{i−1}[γ1] ℓ_{i−1}[γ2]  (src/core.jl:121-128,178-185), every row kept:
// row i−1 is copied to the history at the top of slice i (it is complete since the last barrier), row n after the loop.
template <int NT>
__device__ __noinline__ void run_slices_fwd1(int n, int Cp, double* fin, double* scr, double* cur, const Slot* s_slots,
                                             int nslots, const Ent* s_dents, const double2* pprow, double* hist, int tid) {
    const Pass2 P = load_pass2(s_slots, nslots, s_dents, tid, NT);
    const int npass = (nslots + NT - 1) / NT;
    const int wbase = tid & ~31;
    for (int i = 1; i <= n; i++) {
        const double* src = cur;
        double* dst = (cur == fin) ? scr : fin;
        keep_row(hist + (size_t)(i - 1) * Cp, src, Cp, tid, NT);
        const double2 pp = pprow[i];
        if (P.wg1) {
            double s0 = lane_terms(P.w0, P.two0, P.more0, src, src, s_dents);
            double s1 = lane_terms(P.w1, P.two1, P.more1, src, src, s_dents);
            const double o0 = src[P.c0], o1 = src[P.c1];
            team_reduce2(P, s0, s1);
            if (P.lead0) dst[P.c0] = fma(pp.x, o0, pp.y * s0);
            if (P.lead1) dst[P.c1] = fma(pp.x, o1, pp.y * s1);
        } else if (P.wg0) {
            double s0 = lane_terms(P.w0, P.two0, P.more0, src, src, s_dents);
            const double o0 = src[P.c0];
            team_reduce1(P.w0, P.wg0, s0);
            if (P.lead0) dst[P.c0] = fma(pp.x, o0, pp.y * s0);
        }
        for (int q = 2; q < npass; q++) {  // oversized rows: descriptors reloaded from shared memory
            const int wgq = (wbase + q * NT) < nslots ? (1 << s_slots[wbase + q * NT].glog) : 0;
            if (wgq == 0) continue;
            const LaneWork<1> wq = load_work<1>(s_slots, nslots, s_dents, tid + q * NT);
            double sq = lane_terms(wq, WARP_ANY(wq.cnt > 1), WARP_ANY(wq.cnt > 2), src, src, s_dents);
            const int cq = max(wq.cell, 0);
            const double oq = src[cq];
            team_reduce1(wq, wgq, sq);
            if (wq.cell >= 0 && ((tid + q * NT) & (wq.gsz - 1)) == 0) dst[cq] = fma(pp.x, oq, pp.y * sq);
        }
        cur = dst;
        __syncthreads();
    }
    keep_row(hist + (size_t)n * Cp, cur, Cp, tid, NT);
}

// Backward (see the header): `cur` holds ℓ̄_n on entry; rows alternate between arow and scr; returns the buffer holding
// ℓ̄_0.  hist = the branch's kept rows (global, row stride Cp), staged into the three rotating buffers hb[0..2] (hlen
// doubles each): row i−1 is in use by slice i, row i−2 is landing, row i−3 is being issued — one barrier per slice.
// lpp = the local table rows (4 double2 per row: value, ∂λ, ∂μ, ∂ϵ_0).  acc[0..2] += this thread's share of (λ̄, μ̄, ϵ̄⁰).
__device__ __forceinline__ void bwd_leader(double* __restrict__ dst, int c, double o, double y, double s, const double2 t0,
                                           const double2 t1, const double2 t2, const double2 t3, double (&acc)[3]) {
    dst[c] = fma(t0.x, o, t0.y * s);
    const double wphi = o * y, wpsi = 0.5 * (y * s);
    acc[0] = fma(wphi, t1.x, fma(wpsi, t1.y, acc[0]));
    acc[1] = fma(wphi, t2.x, fma(wpsi, t2.y, acc[1]));
    acc[2] = fma(wphi, t3.x, fma(wpsi, t3.y, acc[2]));
}

template <int NT>
__device__ __noinline__ double* run_slices_bwd(int n, int Cp, double* arow, double* scr, double* cur,
                                               const double* __restrict__ hist, double* hb, int hlen,
                                               const Slot* s_slots, int nslots, const Ent* s_ents,
                                               const double2* s_lpp, double (&acc)[3], int tid) {
    const Pass2 P = load_pass2(s_slots, nslots, s_ents, tid, NT);
    const int npass = (nslots + NT - 1) / NT;
    const int wbase = tid & ~31;
    const int n16 = Cp >> 1;  // 16-byte words per kept row
    auto issue = [&](int row, int buf) {  // row < 0: an empty group keeps the group arithmetic uniform
        if (row >= 0) copy16(reinterpret_cast<uint4*>(hb + (size_t)buf * hlen), reinterpret_cast<const uint4*>(hist + (size_t)row * Cp), n16, tid, NT);
        stage_commit();
    };
    // rows n−1 and n−2 before the loop; slice i (n .. 1) reads row i−1 from buffer (n−i) % 3
    issue(n - 1, 0);
    issue(n - 2, 1);
    stage_wait_prev();  // row n−1 has landed (this thread's copies)
    __syncthreads();
    for (int i = n; i >= 1; i--) {
        const int b = (n - i) % 3;
        issue(i - 3, (b + 2) % 3);  // that buffer held row i (slice i+1): free since the last barrier
        const double* src = cur;
        double* dst = (cur == arow) ? scr : arow;
        const double* val = hb + (size_t)b * hlen;
        const double2* lp = s_lpp + (size_t)i * 4;
        const double2 t0 = lp[0], t1 = lp[1], t2 = lp[2], t3 = lp[3];
        if (P.wg1) {
            double s0 = lane_terms(P.w0, P.two0, P.more0, src, val, s_ents);
            double s1 = lane_terms(P.w1, P.two1, P.more1, src, val, s_ents);
            const double o0 = src[P.c0], y0 = val[P.c0], o1 = src[P.c1], y1 = val[P.c1];
            team_reduce2(P, s0, s1);
            if (P.lead0) bwd_leader(dst, P.c0, o0, y0, s0, t0, t1, t2, t3, acc);
            if (P.lead1) bwd_leader(dst, P.c1, o1, y1, s1, t0, t1, t2, t3, acc);
        } else if (P.wg0) {
            double s0 = lane_terms(P.w0, P.two0, P.more0, src, val, s_ents);
            const double o0 = src[P.c0], y0 = val[P.c0];
            team_reduce1(P.w0, P.wg0, s0);
            if (P.lead0) bwd_leader(dst, P.c0, o0, y0, s0, t0, t1, t2, t3, acc);
        }
        for (int q = 2; q < npass; q++) {  // oversized rows: descriptors reloaded from shared memory
            const int wgq = (wbase + q * NT) < nslots ? (1 << s_slots[wbase + q * NT].glog) : 0;
            if (wgq == 0) continue;
            const LaneWork<1> wq = load_work<1>(s_slots, nslots, s_ents, tid + q * NT);
            double sq = lane_terms(wq, WARP_ANY(wq.cnt > 1), WARP_ANY(wq.cnt > 2), src, val, s_ents);
            const int cq = max(wq.cell, 0);
            const double oq = src[cq], yq = val[cq];
            team_reduce1(wq, wgq, sq);
            if (wq.cell >= 0 && ((tid + q * NT) & (wq.gsz - 1)) == 0) bwd_leader(dst, cq, oq, yq, sq, t0, t1, t2, t3, acc);
        }
        cur = dst;
        stage_wait_prev();  // everything but the newest group: row i−2 has landed
        __syncthreads();
    }
    stage_wait();
    return cur;
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_dp_rev(RevArgs A) {
    EXTERN_SHARED(smem_raw);
    constexpr int NW = NT / 32;
    // (rotating the warp roles per CTA so that co-resident CTAs put their critical warps on different schedulers measured
    // 6 % SLOWER on the B200, profiles/r2_rev_rotation_ab.txt: roles stay put)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ModelDev& M = A.M;
    const PlanDev& PR = A.PR;
    const PlanDev& PG = A.PG;
    const int nn = M.nn, KmaxR = PR.Kmax, KmaxG = PG.Kmax, root = M.root;
    const int KR = PG.K[root];

    // ---- shared memory carve-up (mirrored by smem_need_rev() on the host) ----
    int* s_kind = reinterpret_cast<int*>(smem_raw);
    int* s_nsl = s_kind + nn;
    int* s_ch0 = s_nsl + nn;
    int* s_ch1 = s_ch0 + nn;
    int* s_K = s_ch1 + nn;       // hybrid plan
    int* s_toff = s_K + nn;      // hybrid plan
    int* s_ltoff = s_toff + nn;  // local plan
    int* s_roff = s_ltoff + nn;
    int* s_aoff = s_roff + nn;
    int* s_misc = s_aoff + nn;   // [4]: next family index
    const size_t meta_bytes = (((9 * nn + 4) * sizeof(int)) + 15) & ~size_t(15);
    NodeRec* s_nrec = reinterpret_cast<NodeRec*>(smem_raw + meta_bytes);
    RevRec* s_rrec = reinterpret_cast<RevRec*>(s_nrec + nn);
    // [nn][8] local adjoints: internal/WGD/root {λ̄, μ̄, ϵ̄⁰ (slices), c̄x, c̄y (WGD, root), ϵ̄ⁿ of child 0, ϵ̄ⁿ of child 1, -};
    // leaf branches {Σ_γ ℓ̄_n[γ]·∂ℓ_n[γ]/∂(component 1), ·/∂(component 2), ...}: row r of the Jacobian table k_tables3 built
    double* zloc = reinterpret_cast<double*>(s_rrec + nn);
    double* s_red = zloc + 8 * nn;                          // [NW*8] block_sum scratch
    double* s_res = s_red + NW * 8;                         // [8] block_sum results
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_res + 8);  // [2] mbarrier of the list prefetches
    unsigned char* dyn = reinterpret_cast<unsigned char*>(s_bar + 2);
    if (tid == 0) mbar_init(s_bar, 1);
    unsigned s2_phase = 0;
    for (int i = tid; i < nn; i += NT) {
        s_kind[i] = M.kind[i]; s_nsl[i] = M.nsl[i]; s_ch0[i] = M.child0[i]; s_ch1[i] = M.child1[i];
        s_K[i] = PR.K[i]; s_toff[i] = PR.toff[i]; s_ltoff[i] = A.PL.toff[i];
    }
    int nfam_done = 0;
    for (;;) {
        __syncthreads();  // the previous family is finished with shared memory
        if (tid == 0) s_misc[0] = (int)atomicAdd(A.next, 1u);
        __syncthreads();
        const int pidx = s_misc[0];
        if (pidx >= A.bin_count) break;
        nfam_done++;
        const int fam = A.perm[A.bin_off + pidx];
        const FamHdr* Hp = A.hdr + fam;
        const RevHdr* Rp = A.rhdr + fam;
        const unsigned char* blob = A.arena + Hp->base;
        const unsigned char* rblob = A.rarena + Rp->base;
        const uint32_t nlev = Hp->nlev, blob_bytes = Hp->blob_bytes, rblob_bytes = Rp->blob_bytes;
        const uint32_t rows_len = Rp->rows_len, scr_len = Rp->scr_len, leafmax = Rp->leafmax;
        const uint32_t stage_bytes = Rp->stage_bytes, leaf_stage = Hp->leaf_stage;
        const uint32_t hlen = Rp->hbuf_len;
        const uint32_t* words = reinterpret_cast<const uint32_t*>(blob);
        const Ent* ents = reinterpret_cast<const Ent*>(blob);
        const uint32_t* rwords = reinterpret_cast<const uint32_t*>(rblob);
        const Ent* rents = reinterpret_cast<const Ent*>(rblob);
        long long tc0 = CLOCK64(), acc_fsl = 0, acc_bsl = 0, acc_froot = 0;
        for (uint32_t o = tid * 128u; o < blob_bytes; o += NT * 128u) PREFETCH_L2(blob + o);
        for (uint32_t o = tid * 128u; o < rblob_bytes; o += NT * 128u) PREFETCH_L2(rblob + o);

        double* rows = reinterpret_cast<double*>(dyn);
        double* scr = rows + rows_len;
        unsigned char* stage = reinterpret_cast<unsigned char*>(scr + scr_len);
        unsigned char* stage2 = stage + stage_bytes;     // row-1 / root lists, prefetched one node ahead (bulk copies)
        const uint32_t stage2_bytes = Rp->stage2_bytes;
        const bool staged2 = stage2_bytes != 0;
        // the root has no slices: its (largest) lists use `stage` and `stage2` together, requested once the last slice
        // loop is over; stage2 alone only has to hold the row-1 lists of one internal/WGD node
        const bool root_staged = Rp->root_staged != 0;
        unsigned char* leaf_area = stage2 + stage2_bytes;  // phase A only; the backward pass reuses it:
        const size_t leaf_area_bytes = (size_t)leafmax * sizeof(double) + leaf_stage;
        double* arows = reinterpret_cast<double*>(leaf_area);
        double* hb = arows + Rp->arows_len;  // [3][hlen]
        const bool staged = stage_bytes != 0;
        for (int i = tid; i < nn; i += NT) {
            s_roff[i] = (int)A.roff[(size_t)fam * nn + i];
            s_aoff[i] = (int)A.aoff[(size_t)fam * nn + i];
        }
        for (int i = tid; i < nn * 3; i += NT) {
            reinterpret_cast<uint4*>(s_nrec)[i] = __ldg(reinterpret_cast<const uint4*>(blob) + i);
            reinterpret_cast<uint4*>(s_rrec)[i] = __ldg(reinterpret_cast<const uint4*>(rblob) + i);
        }
        for (int i = tid; i < 8 * nn; i += NT) zloc[i] = 0.0;
        __syncthreads();
        const NodeRec* const nrec = s_nrec;
        const RevRec* const rrec = s_rrec;
        double* const hist = A.hist + (size_t)(A.slot0 + blockIdx.x) * A.hist_stride;
        // The row-1 / root lists of a node (read once per evaluation) are brought into shared memory by the copy engine
        // one node ahead — thread 0 issues cp.async.bulk copies that complete on an mbarrier while the current node's
        // slices run.  A request lists up to six contiguous segments of the family's blobs; the consumer derives the same
        // layout.  stage2_bytes == 0 (lists too long: large CCDs): the segments are read in place from global memory.
        // (segment sizes are recomputed from the node records on both sides: scalars only, nothing indexed dynamically)
        struct Lay { uint32_t a, b, c, d, e, f; };  // byte sizes of the (up to six) segments, in order
        auto p4 = [](uint32_t w) -> uint32_t { return ((w + 3u) & ~3u) * 4u; };
        auto fwd_lay = [&](int e) -> Lay {  // forward: [dptr] [tptr|lossF|lossG|lev] [dents] [tents]
            const NodeRec& Q = nrec[e];
            const int kd = s_kind[e];
            Lay y;
            y.a = kd != WHALE_INTERNAL ? p4(Q.C + 1) : 0u;
            y.b = kd != WHALE_WGD ? p4(3 * Q.C + 1 + (kd == WHALE_ROOT ? nlev + 1 : 0)) : 0u;
            y.c = kd != WHALE_INTERNAL ? Q.ndent * 16u : 0u;
            y.d = kd != WHALE_WGD ? Q.ntent * 16u : 0u;
            y.e = y.f = 0u;
            return y;
        };
        auto bwd_lay = [&](int e) -> Lay {  // transposed: [bptr] [sF|upF] [sG|upG] [bents] [sFents] [sGents]
            const RevRec& Q = rrec[e];
            const int kd = s_kind[e];
            Lay y;
            y.a = kd != WHALE_INTERNAL ? p4(nrec[e].C + 1) : 0u;
            y.b = kd != WHALE_WGD ? p4(2 * nrec[s_ch0[e]].C + 1) : 0u;
            y.c = kd != WHALE_WGD ? p4(2 * nrec[s_ch1[e]].C + 1) : 0u;
            y.d = kd != WHALE_INTERNAL ? Q.nbent * 16u : 0u;
            y.e = kd != WHALE_WGD ? Q.nsFent * 16u : 0u;
            y.f = kd != WHALE_WGD ? Q.nsGent * 16u : 0u;
            return y;
        };
        // requests must follow a barrier after which nobody reads the previous contents of stage2
        auto s2_on = [&](int e) -> bool { return e == root ? root_staged : staged2; };
        auto s2_base = [&](int e) -> unsigned char* { return e == root ? stage : stage2; };
        unsigned char* s2_dst = stage2;
        auto s2_copy = [&](uint32_t& off, const unsigned char* src, uint32_t bytes) {
            if (bytes) bulk_g2s(s2_dst + off, src, bytes, s_bar);
            off += bytes;
        };
        auto s2_request_fwd = [&](int e) {
            if (!s2_on(e) || tid != 0) return;
            s2_dst = s2_base(e);
            const Lay y = fwd_lay(e);
            const NodeRec& Q = nrec[e];
            fence_proxy_async();
            mbar_expect_tx(s_bar, y.a + y.b + y.c + y.d);
            uint32_t off = 0;
            s2_copy(off, blob + (size_t)Q.dptr_off * 4, y.a);
            s2_copy(off, blob + (size_t)Q.tptr_off * 4, y.b);
            s2_copy(off, blob + (size_t)Q.dent_off * 16, y.c);
            s2_copy(off, blob + (size_t)Q.tent_off * 16, y.d);
        };
        auto s2_request_bwd = [&](int e) {
            if (!s2_on(e) || tid != 0) return;
            s2_dst = s2_base(e);
            const Lay y = bwd_lay(e);
            const RevRec& Q = rrec[e];
            fence_proxy_async();
            mbar_expect_tx(s_bar, y.a + y.b + y.c + y.d + y.e + y.f);
            uint32_t off = 0;
            s2_copy(off, rblob + (size_t)Q.bptr_off * 4, y.a);
            s2_copy(off, rblob + (size_t)Q.sF_off * 4, y.b);
            s2_copy(off, rblob + (size_t)Q.sG_off * 4, y.c);
            s2_copy(off, rblob + (size_t)Q.bent_off * 16, y.d);
            s2_copy(off, rblob + (size_t)Q.sFent_off * 16, y.e);
            s2_copy(off, rblob + (size_t)Q.sGent_off * 16, y.f);
        };
        auto s2_wait = [&](int e) {
            if (!s2_on(e)) return;
            mbar_wait(s_bar, s2_phase);
            s2_phase ^= 1u;
        };
        auto next_fwd = [&](int oi) -> int { for (int j = oi + 1; j < M.ninner; j++) if (nrec[M.inner[j]].C) return j; return -1; };
        auto next_bwd = [&](int oi) -> int { for (int j = oi - 1; j >= 0; j--) if (nrec[M.inner[j]].C) return j; return -1; };
        const int j_first = next_fwd(-1);
        if (j_first >= 0 && M.inner[j_first] != root) s2_request_fwd(M.inner[j_first]);  // streams in during the leaf phase
        GRID_DEP_WAIT();  // programmatic dependent launch: the prologue of the CTA's first family overlapped k_tables3
        const long long tcA = CLOCK64();
        // ================= phase A: leaf branches (as in k_dp, hybrid plan: value + own λ, μ) =================
        for (int li = warp; li < M.nleafnodes; li += NW) {
            const int e = M.leafnodes[li];
            const NodeRec R = nrec[e];
            const int C = (int)R.C;
            if (C == 0) continue;
            const int K = s_K[e], n = s_nsl[e];
            double* fin = rows + s_roff[e];
            if (R.sptr_off) {  // closed form over tree shapes
                const uint32_t* sptr = words + R.sptr_off;
                const Ent* sent = ents + R.sent_off;
                const double* W = PR.shapeW + (size_t)e * NSHAPE * KmaxR;
                for (int i = lane; i < C * K; i += 32) {
                    const int c = i / K, k = i - c * K;
                    double v = 0.0;
                    for (uint32_t t = sptr[c]; t < sptr[c + 1]; t++) {
                        const Ent en = sent[t];
                        v = fma(en.p, W[en.i1 * KmaxR + k], v);
                    }
                    fin[c * RS(K) + k] = v;
                }
                continue;
            }
            if (R.nonleaf == 0) {  // family-independent: ℓ_n = leafℙ·Πϕ_i from k_tables
                for (int i = lane; i < C * K; i += 32) fin[(i / K) * RS(K) + (i % K)] = PR.leaf[e * KmaxR + (i % K)];
                continue;
            }
            if (R.nslots > HEAVY_SLOTS) continue;  // heavy branch: whole CTA, below
            double* wscr = reinterpret_cast<double*>(leaf_area + warp * leaf_area_bytes);
            uint4* wst = reinterpret_cast<uint4*>(wscr + leafmax);
            const int nd16 = (int)R.ndent, sl16 = ((int)R.nslots + 1) >> 1;
            copy16(wst, reinterpret_cast<const uint4*>(ents + R.dent_off), nd16, lane, 32);
            copy16(wst + nd16, reinterpret_cast<const uint4*>(words + R.slot_off), sl16, lane, 32);
            double* cur = (n & 1) ? wscr : fin;
            const int nleafc = C - (int)R.nonleaf;
            for (int i = lane; i < C * K; i += 32) {
                const int c = i / K, k = i - c * K;
                cur[c * RS(K) + k] = (c < nleafc && k == 0) ? M.leafP[e] : 0.0;
            }
            stage_wait();
            __syncwarp();
            if (!run_slices_fused_k<true>(K, n, C, fin, wscr, cur, reinterpret_cast<const Slot*>(wst + nd16), (int)R.nslots,
                                          reinterpret_cast<const Ent*>(wst), PR.pp + s_toff[e], nullptr, lane, 32))
                run_slices_kc<1, true>(n, C, K, fin, wscr, cur, reinterpret_cast<const Slot*>(wst + nd16), (int)R.nslots,
                                       reinterpret_cast<const Ent*>(wst), PR.pp + s_toff[e], nullptr, lane);
        }
        __syncthreads();
        for (int li = 0; li < M.nleafnodes; li++) {  // leaf branches with many in-paralog clades: whole CTA
            const int e = M.leafnodes[li];
            const NodeRec R = nrec[e];
            const int C = (int)R.C;
            if (C == 0 || R.nslots <= HEAVY_SLOTS || R.nonleaf == 0 || R.sptr_off) continue;
            const int K = s_K[e], n = s_nsl[e];
            double* fin = rows + s_roff[e];
            const int nd16 = (int)R.ndent, sl16 = ((int)R.nslots + 1) >> 1, pp16 = (n + 1) * K;
            const Ent* h_dents = ents + R.dent_off;
            const Slot* h_slots = reinterpret_cast<const Slot*>(words + R.slot_off);
            const double2* h_pp = PR.pp + s_toff[e];
            if (staged) {
                uint4* st4 = reinterpret_cast<uint4*>(stage);
                copy16(st4, reinterpret_cast<const uint4*>(h_dents), nd16, tid, NT);
                copy16(st4 + nd16, reinterpret_cast<const uint4*>(h_slots), sl16, tid, NT);
                copy16(st4 + nd16 + sl16, reinterpret_cast<const uint4*>(h_pp), pp16, tid, NT);
                h_dents = reinterpret_cast<const Ent*>(st4);
                h_slots = reinterpret_cast<const Slot*>(st4 + nd16);
                h_pp = reinterpret_cast<const double2*>(st4 + nd16 + sl16);
            }
            double* cur = (n & 1) ? scr : fin;
            const int nleafc = C - (int)R.nonleaf;
            for (int i = tid; i < C * K; i += NT) {
                const int c = i / K, k = i - c * K;
                cur[c * RS(K) + k] = (c < nleafc && k == 0) ? M.leafP[e] : 0.0;
            }
            stage_wait();
            __syncthreads();
            run_slices_fused_k<false>(K, n, C, fin, scr, cur, h_slots, (int)R.nslots, h_dents, h_pp, nullptr, tid, NT);
            __syncthreads();
        }
        const long long tcB = CLOCK64();
        if (j_first >= 0 && M.inner[j_first] == root) s2_request_fwd(root);  // (a tree without internal branches)

        // ================= phase B forward: internal, WGD and root nodes, value only, rows kept =================
        for (int oi = 0; oi < M.ninner; oi++) {
            const int e = M.inner[oi];
            const NodeRec R = nrec[e];
            const int C = (int)R.C;
            if (C == 0) continue;
            const int kind = s_kind[e], n = s_nsl[e];
            const int Cp = (C + 1) & ~1;
            const long long tn0 = CLOCK64();
            double* fin = rows + s_roff[e];
            double* hrow = hist + rrec[e].hoff;
            const int f = s_ch0[e], g = s_ch1[e];
            const int sF = RS(s_K[f]);
            const double* finF = rows + s_roff[f];
            // ---- stage this node's lists: [dents | slots | ϕψ rows]; the row-1 lists are read once: in place ----
            const int nd16 = (kind == WHALE_ROOT) ? 0 : (int)R.ndent;
            const int sl16 = (kind == WHALE_ROOT) ? 0 : (((int)R.nslots + 1) >> 1);
            const int pp16 = (kind == WHALE_ROOT) ? 0 : (n + 1);
            const Ent* s_dents = ents + R.dent_off;
            const Slot* s_slots = reinterpret_cast<const Slot*>(words + R.slot_off);
            const double2* s_pp = PR.pp + s_toff[e];
            // row-1 / root lists: the segments requested one node ahead (see fwd_segs for their order)
            const Lay ly = fwd_lay(e);
            const bool st2 = s2_on(e);
            const unsigned char* sb = s2_base(e);
            const uint32_t* g_dptr = st2 ? reinterpret_cast<const uint32_t*>(sb) : words + R.dptr_off;
            const uint32_t* g_tptr = st2 ? reinterpret_cast<const uint32_t*>(sb + ly.a) : words + R.tptr_off;
            const Ent* g_dents = st2 ? reinterpret_cast<const Ent*>(sb + ly.a + ly.b) : ents + R.dent_off;
            const Ent* g_tents = st2 ? reinterpret_cast<const Ent*>(sb + ly.a + ly.b + ly.c) : ents + R.tent_off;
            if (staged && kind != WHALE_ROOT) {
                uint4* st4 = reinterpret_cast<uint4*>(stage);
                copy16(st4, reinterpret_cast<const uint4*>(s_dents), nd16, tid, NT);
                copy16(st4 + nd16, reinterpret_cast<const uint4*>(s_slots), sl16, tid, NT);
                copy16(st4 + nd16 + sl16, reinterpret_cast<const uint4*>(s_pp), pp16, tid, NT);
                s_dents = reinterpret_cast<const Ent*>(st4);
                s_slots = reinterpret_cast<const Slot*>(st4 + nd16);
                s_pp = reinterpret_cast<const double2*>(st4 + nd16 + sl16);
            }
            double* cur = (n & 1) ? scr : fin;
            s2_wait(e);  // this node's row-1 lists have landed
            if (kind == WHALE_WGD) {  // q·Σ p ℓ_f[γ1]ℓ_f[γ2] + (1−q+2qϵ_f)·ℓ_f[γ]   src/core.jl:103-119,187-199
                const double cx0 = PR.cx[e * KmaxR], cy0 = PR.cy[e * KmaxR];
                for (int c = tid; c < C; c += NT) {
                    const double s0 = vsum<false>(g_dents, g_dptr[c], g_dptr[c + 1], 1u, finF, sF, finF, sF);
                    cur[c] = fma(cy0, s0, cx0 * finF[c * sF]);
                }
            } else {
                const int sG = RS(s_K[g]);
                const double* finG = rows + s_roff[g];
                const double ef0 = PR.eps[s_toff[f] + s_nsl[f] * s_K[f]], eg0 = PR.eps[s_toff[g] + s_nsl[g] * s_K[g]];
                const int32_t* g_lossF = reinterpret_cast<const int32_t*>(g_tptr + C + 1);
                const int32_t* g_lossG = g_lossF + C;
                if (kind == WHALE_INTERNAL) {  // Πspeciation + Πloss  src/core.jl:160-176
                    for (int c = tid; c < C; c += NT) {
                        const double s0 = vsum<false>(g_tents, g_tptr[c], g_tptr[c + 1], 1u, finF, sF, finG, sG);
                        const int lf = g_lossF[c], lg = g_lossG[c];
                        const double l0 = (lf >= 0 ? finF[lf * sF] : 0.0) * eg0 + (lg >= 0 ? finG[lg * sG] : 0.0) * ef0;
                        cur[c] = s0 + l0;
                    }
                } else {  // root (src/core.jl:130-158): clades ascending in size, one level at a time
                    const double cx0 = PR.cx[e * KmaxR], cy0 = PR.cy[e * KmaxR];
                    const uint32_t* g_lev = reinterpret_cast<const uint32_t*>(g_lossG + C);
                    for (uint32_t L = 0; L < nlev; L++) {
                        const int c0 = (int)g_lev[L], c1 = (int)g_lev[L + 1];
                        int glog = 0;
                        {   // team size from the level's shape only: as many lanes per cell as the CTA has, while a lane
                            // still gets at least two terms
                            const uint32_t nt_ = (g_dptr[c1] - g_dptr[c0]) + (g_tptr[c1] - g_tptr[c0]);
                            while (glog < 5 && ((c1 - c0) << (glog + 1)) <= NT && (2u << glog) * (uint32_t)(c1 - c0) <= nt_) glog++;
                        }
                        const int G = 1 << glog;
                        const int lanes = (c1 - c0) << glog;
                        for (int ub = 0; ub < lanes; ub += NT) {  // uniform trip count: the shuffles need whole warps
                            const int u = ub + tid;
                            const int c = c0 + (u >> glog), j = u & (G - 1);
                            const bool valid = u < lanes;
                            double v = 0.0;
                            if (valid) {
                                const double a0 = vsum<false>(g_dents, g_dptr[c] + j, g_dptr[c + 1], (uint32_t)G, fin, 1, fin, 1);
                                const double b0 = vsum<false>(g_tents, g_tptr[c] + j, g_tptr[c + 1], (uint32_t)G, finF, sF, finG, sG);
                                v = fma(cx0, a0, cy0 * b0);
                            }
                            for (int step = 1; step < G; step <<= 1) v += SHFL_DOWN(v, step);
                            if (valid && j == 0) {
                                const int lf = g_lossF[c], lg = g_lossG[c];
                                const double l0 = (lf >= 0 ? finF[lf * sF] : 0.0) * eg0 + (lg >= 0 ? finG[lg * sG] : 0.0) * ef0;
                                fin[c] = fma(cy0, l0, v);
                            }
                        }
                        __syncthreads();
                    }
                    s2_request_bwd(e);  // the root's transposed lists, for the backward pass
                    acc_froot += CLOCK64() - tn0;
                    continue;
                }
            }
            stage_wait();
            __syncthreads();  // row 1 and the staged lists are visible
            const long long ts0 = CLOCK64();
            const int jn = next_fwd(oi);
            // row 1 is done: the next node's row-1 lists stream in while this node's slices run
            if (jn >= 0 && M.inner[jn] != root) s2_request_fwd(M.inner[jn]);
            run_slices_fwd1<NT>(n, Cp, fin, scr, cur, s_slots, (int)R.nslots, s_dents, s_pp, hrow, tid);
            __syncthreads();  // the last row is complete; the staging buffer may be reused
            if (jn >= 0 && M.inner[jn] == root) s2_request_fwd(root);  // (the root's lists also use this node's slice buffer)
            const long long ts1 = CLOCK64();
            acc_fsl += ts1 - ts0;
            if (A.tim && tid == 0 && oi < TIMN) A.tim[(size_t)fam * TIMW + 8 + oi] = ts1 - tn0;
        }
        const long long tcC = CLOCK64();

        // ================= backward =================
        const int CR = (int)nrec[root].C;
        const double Lv = rows[s_roff[root] + CR - 1];
        if (!(Lv > 0.0)) {  // L <= 0 -> −Inf, zero gradient (src/core.jl:36)
            for (int k = tid; k < KR; k += NT) A.out_fam[(size_t)fam * KR + k] = k == 0 ? -dinf() : 0.0;
            s2_wait(root);  // (the root's transposed lists were requested: keep the barrier's phase in step)
            continue;
        }
        // per-node local adjoints: zloc[e] = {λ̄, μ̄, ϵ̄⁰ (slices), c̄x, c̄y (WGD, root), ϵ̄ⁿ of child 0, ϵ̄ⁿ of child 1, -}
        auto leaf_or_inner_child = [&](int c, int buf, const double*& V, int& sV) {
            // value row of child c for a row-1 transposition: leaf rows are resident; internal/WGD rows come back
            // from the history (last row) into staging buffer `buf`
            if (s_kind[c] == WHALE_LEAF) { V = rows + s_roff[c]; sV = RS(s_K[c]); return; }
            const int Cc = (int)nrec[c].C, Ccp = (Cc + 1) & ~1;
            copy16(reinterpret_cast<uint4*>(hb + (size_t)buf * hlen),
                   reinterpret_cast<const uint4*>(hist + rrec[c].hoff + (size_t)s_nsl[c] * Ccp), Ccp >> 1, tid, NT);
            V = hb + (size_t)buf * hlen; sV = 1;
        };
        // transposed Πspeciation + Πloss towards one child (src/core.jl:160-176): for every cell γ' of child `ch`
        //   ℓ̄_ch[γ'] = coef·(Σ_{(π,σ,p)} p·Ā[π]·V_other[σ] + Ā[up(γ')]·ϵ_other);  returns this thread's shares of
        //   Σ V_ch[γ']·Σ_{...} (→ c̄y at the root), Σ Ā[up(γ')]·V_ch[γ'] (→ ϵ̄ of the OTHER child) and, for a leaf child,
        //   of Σ_γ' ℓ̄_ch[γ']·∂ℓ_ch[γ'] (components 1, 2)
        auto spec_down = [&](int ch, const uint32_t* sptr, const Ent* sent, const double* Abar, const double* Vch, int sVch,
                             const double* Vo, int sVo, double eps_o, double coef, double& a_sv, double& a_up,
                             double& a_l1, double& a_l2) {
            const int Cc = (int)nrec[ch].C;
            const int32_t* up = reinterpret_cast<const int32_t*>(sptr + Cc + 1);
            const bool isleaf = s_kind[ch] == WHALE_LEAF;
            const int Kc = s_K[ch];
            double* Ach = isleaf ? nullptr : arows + s_aoff[ch];
            for (int c = tid; c < Cc; c += NT) {
                const double s = vsum<false>(sent, sptr[c], sptr[c + 1], 1u, Abar, 1, Vo, sVo);
                const double au = Abar[up[c]];
                const double v = Vch[c * sVch];
                const double ab = coef * fma(au, eps_o, s);
                a_sv = fma(v, s, a_sv);
                a_up = fma(au, v, a_up);
                if (isleaf) {
                    if (Kc > 1) a_l1 = fma(ab, Vch[c * sVch + 1], a_l1);
                    if (Kc > 2) a_l2 = fma(ab, Vch[c * sVch + 2], a_l2);
                } else {
                    Ach[c] = ab;
                }
            }
        };

        {   // ---- root, transposed: levels in DESCENDING clade size ----
            const int e = root;
            const NodeRec R = nrec[e];
            const RevRec RR = rrec[e];
            const int f = s_ch0[e], g = s_ch1[e];
            const int sF = RS(s_K[f]), sG = RS(s_K[g]);
            const double* finF = rows + s_roff[f];
            const double* finG = rows + s_roff[g];
            const double* fin = rows + s_roff[e];
            double* Ab = arows + s_aoff[e];
            const double cx0 = PR.cx[e * KmaxR], cy0 = PR.cy[e * KmaxR];
            const double ef0 = PR.eps[s_toff[f] + s_nsl[f] * s_K[f]], eg0 = PR.eps[s_toff[g] + s_nsl[g] * s_K[g]];
            const uint32_t* g_tptr = words + R.tptr_off;
            const uint32_t* g_lev = g_tptr + 3 * CR + 1;
            const Lay ly = bwd_lay(e);  // [bptr] [sF|upF] [sG|upG] [bents] [sFents] [sGents]
            const unsigned char* sp0 = root_staged ? stage : rblob + (size_t)RR.bptr_off * 4;
            const unsigned char* sp1 = root_staged ? stage + ly.a : rblob + (size_t)RR.sF_off * 4;
            const unsigned char* sp2 = root_staged ? stage + ly.a + ly.b : rblob + (size_t)RR.sG_off * 4;
            const unsigned char* sp3 = root_staged ? stage + ly.a + ly.b + ly.c : rblob + (size_t)RR.bent_off * 16;
            const unsigned char* sp4 = root_staged ? stage + ly.a + ly.b + ly.c + ly.d : rblob + (size_t)RR.sFent_off * 16;
            const unsigned char* sp5 = root_staged ? stage + ly.a + ly.b + ly.c + ly.d + ly.e : rblob + (size_t)RR.sGent_off * 16;
            s2_wait(root);
            const uint32_t* bptr = reinterpret_cast<const uint32_t*>(sp0);
            const Ent* bent = reinterpret_cast<const Ent*>(sp3);
            const double seed = 1.0 / Lv;  // d log L / dL
            double acx = 0.0;
            for (int L = (int)nlev - 1; L >= 0; L--) {
                const int c0 = (int)g_lev[L], c1 = (int)g_lev[L + 1];
                const uint32_t ne = bptr[c1] - bptr[c0];
                int glog = 0;  // team size from the level's shape only
                while (glog < 5 && ((c1 - c0) << (glog + 1)) <= NT && (2u << glog) * (uint32_t)(c1 - c0) <= 2u * ne) glog++;
                const int G = 1 << glog;
                const int lanes = (c1 - c0) << glog;
                for (int ub = 0; ub < lanes; ub += NT) {
                    const int u = ub + tid;
                    const int c = c0 + (u >> glog), j = u & (G - 1);
                    const bool valid = u < lanes;
                    double v = 0.0;
                    if (valid) v = vsum<false>(bent, bptr[c] + j, bptr[c + 1], (uint32_t)G, Ab, 1, fin, 1);
                    for (int step = 1; step < G; step <<= 1) v += SHFL_DOWN(v, step);
                    if (valid && j == 0) {
                        Ab[c] = fma(cx0, v, c == CR - 1 ? seed : 0.0);
                        acx = fma(0.5 * fin[c], v, acx);
                    }
                }
                __syncthreads();
            }
            double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // Σ V_F·s_F, ū_G, lF1, lF2, (unused), ū_F, lG1, lG2
            double dummy = 0.0;
            spec_down(f, reinterpret_cast<const uint32_t*>(sp1), reinterpret_cast<const Ent*>(sp4), Ab, finF, sF, finG, sG, eg0,
                      cy0, a[0], a[1], a[2], a[3]);
            spec_down(g, reinterpret_cast<const uint32_t*>(sp2), reinterpret_cast<const Ent*>(sp5), Ab, finG, sG, finF, sF, ef0,
                      cy0, dummy, a[5], a[6], a[7]);
            a[4] = acx;
            block_sum<8, NT>(a, s_red, s_res, false, tid);
            if (tid == 0) {
                // ū_G = Σ_c Ā[c]·ℓ_F[lf(c)] is the adjoint share of ϵ_G (Πloss: ℓ_F[γ]·ϵ_G), ū_F that of ϵ_F
                const double uG = s_res[1], uF = s_res[5];
                zloc[e * 8 + 3] = s_res[4];                                // c̄x
                zloc[e * 8 + 4] = s_res[0] + uG * eg0 + uF * ef0;          // c̄y = Σ Ā·(b + loss)
                zloc[e * 8 + 5] = cy0 * uF;                                // ϵ̄ⁿ of child 0
                zloc[e * 8 + 6] = cy0 * uG;                                // ϵ̄ⁿ of child 1
                if (s_kind[f] == WHALE_LEAF) { zloc[f * 8 + 0] = s_res[2]; zloc[f * 8 + 1] = s_res[3]; }
                if (s_kind[g] == WHALE_LEAF) { zloc[g * 8 + 0] = s_res[6]; zloc[g * 8 + 1] = s_res[7]; }
            }
            __syncthreads();
            {   // the first backward node's transposed row-1 lists stream in during its slices
                const int jn = next_bwd(M.ninner - 1);
                if (jn >= 0) s2_request_bwd(M.inner[jn]);
            }
        }
        const long long tcD = CLOCK64();

        for (int oi = M.ninner - 2; oi >= 0; oi--) {
            const int e = M.inner[oi];
            const NodeRec R = nrec[e];
            const RevRec RR = rrec[e];
            const int C = (int)R.C;
            if (C == 0) continue;
            const int kind = s_kind[e], n = s_nsl[e];
            const int Cp = (C + 1) & ~1;
            double* arow = arows + s_aoff[e];
            double* cur = arow;
            const long long tn0 = CLOCK64();
            if (n > 0) {  // ---- slices n .. 1, transposed ----
                const int nb16 = (int)RR.nbent, sl16 = ((int)RR.nbslots + 1) >> 1, lp16 = 4 * (n + 1);
                const Ent* s_bents = rents + RR.bent_off;
                const Slot* s_bslots = reinterpret_cast<const Slot*>(rwords + RR.bslot_off);
                const double2* s_lpp = A.PL.pp + s_ltoff[e];
                if (staged) {
                    uint4* st4 = reinterpret_cast<uint4*>(stage);
                    copy16(st4, reinterpret_cast<const uint4*>(s_bents), nb16, tid, NT);
                    copy16(st4 + nb16, reinterpret_cast<const uint4*>(s_bslots), sl16, tid, NT);
                    copy16(st4 + nb16 + sl16, reinterpret_cast<const uint4*>(s_lpp), lp16, tid, NT);
                    stage_wait();
                    s_bents = reinterpret_cast<const Ent*>(st4);
                    s_bslots = reinterpret_cast<const Slot*>(st4 + nb16);
                    s_lpp = reinterpret_cast<const double2*>(st4 + nb16 + sl16);
                }
                __syncthreads();
                double acc[3] = {0.0, 0.0, 0.0};
                cur = run_slices_bwd<NT>(n, Cp, arow, scr, arow, hist + RR.hoff, hb, (int)hlen, s_bslots, (int)RR.nbslots,
                                         s_bents, s_lpp, acc, tid);
                block_sum<3, NT>(acc, s_red, zloc + e * 8, false, tid);
                acc_bsl += CLOCK64() - tn0;
            }
            // ---- row 1, transposed ----
            const int f = s_ch0[e], g = s_ch1[e];
            const Lay ly = bwd_lay(e);  // WGD: [bptr] [bents]; internal: [sF|upF] [sG|upG] [sFents] [sGents]
            if (kind == WHALE_WGD) {
                // ℓ_0[c] = cy·Σ_{D(c)} p ℓ_f[i1]ℓ_f[i2] + cx·ℓ_f[c]   (the child's compat list is this node's)
                const double* VF; int sVF;
                leaf_or_inner_child(f, 0, VF, sVF);
                stage_wait();
                s2_wait(e);
                __syncthreads();
                const double cx0 = PR.cx[e * KmaxR], cy0 = PR.cy[e * KmaxR];
                const uint32_t* bptr = staged2 ? reinterpret_cast<const uint32_t*>(stage2) : rwords + RR.bptr_off;
                const Ent* bent = staged2 ? reinterpret_cast<const Ent*>(stage2 + ly.a) : rents + RR.bent_off;
                const bool isleaf = s_kind[f] == WHALE_LEAF;
                const int Kc = s_K[f];
                double* Ach = isleaf ? nullptr : arows + s_aoff[f];
                double a[4] = {0.0, 0.0, 0.0, 0.0};  // c̄x, c̄y, l1, l2
                for (int c = tid; c < C; c += NT) {
                    const double s = vsum<false>(bent, bptr[c], bptr[c + 1], 1u, cur, 1, VF, sVF);
                    const double a0 = cur[c], v = VF[c * sVF];
                    const double ab = fma(cy0, s, cx0 * a0);
                    a[0] = fma(a0, v, a[0]);
                    a[1] = fma(0.5 * v, s, a[1]);
                    if (isleaf) {
                        if (Kc > 1) a[2] = fma(ab, VF[c * sVF + 1], a[2]);
                        if (Kc > 2) a[3] = fma(ab, VF[c * sVF + 2], a[3]);
                    } else {
                        Ach[c] = ab;
                    }
                }
                block_sum<4, NT>(a, s_red, s_res, false, tid);
                if (tid == 0) {
                    zloc[e * 8 + 3] = s_res[0];
                    zloc[e * 8 + 4] = s_res[1];
                    if (isleaf) { zloc[f * 8 + 0] = s_res[2]; zloc[f * 8 + 1] = s_res[3]; }
                }
                __syncthreads();
            } else {  // internal node
                const double* VF; const double* VG; int sVF, sVG;
                leaf_or_inner_child(f, 0, VF, sVF);
                leaf_or_inner_child(g, 1, VG, sVG);
                stage_wait();
                s2_wait(e);
                __syncthreads();
                const double ef0 = PR.eps[s_toff[f] + s_nsl[f] * s_K[f]], eg0 = PR.eps[s_toff[g] + s_nsl[g] * s_K[g]];
                double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                double d0 = 0.0, d1 = 0.0;
                const uint32_t* pF = staged2 ? reinterpret_cast<const uint32_t*>(stage2) : rwords + RR.sF_off;
                const uint32_t* pG = staged2 ? reinterpret_cast<const uint32_t*>(stage2 + ly.b) : rwords + RR.sG_off;
                const Ent* eF = staged2 ? reinterpret_cast<const Ent*>(stage2 + ly.b + ly.c) : rents + RR.sFent_off;
                const Ent* eG = staged2 ? reinterpret_cast<const Ent*>(stage2 + ly.b + ly.c + ly.e) : rents + RR.sGent_off;
                spec_down(f, pF, eF, cur, VF, sVF, VG, sVG, eg0, 1.0, d0, a[1], a[2], a[3]);
                spec_down(g, pG, eG, cur, VG, sVG, VF, sVF, ef0, 1.0, d1, a[5], a[6], a[7]);
                block_sum<8, NT>(a, s_red, s_res, false, tid);
                if (tid == 0) {
                    zloc[e * 8 + 5] = s_res[5];  // ϵ̄ⁿ of child 0 = Σ_c Ā[c]·ℓ_G[lg(c)]
                    zloc[e * 8 + 6] = s_res[1];  // ϵ̄ⁿ of child 1 = Σ_c Ā[c]·ℓ_F[lf(c)]
                    if (s_kind[f] == WHALE_LEAF) { zloc[f * 8 + 0] = s_res[2]; zloc[f * 8 + 1] = s_res[3]; }
                    if (s_kind[g] == WHALE_LEAF) { zloc[g * 8 + 0] = s_res[6]; zloc[g * 8 + 1] = s_res[7]; }
                }
                __syncthreads();
            }
            {   // (both branches above end with a barrier) the next backward node's lists
                const int jn = next_bwd(oi);
                if (jn >= 0) s2_request_bwd(M.inner[jn]);
            }
            if (A.tim && tid == 0 && oi < TIMN) A.tim[(size_t)fam * TIMW + 8 + TIMN + oi] = CLOCK64() - tn0;
        }
        const long long tcE = CLOCK64();

        // ================= contraction with the table tangents =================
        // ∂ log L_f/∂θ_k = Σ_r z_r·J[r][k]: J (k_tables3, plan G: one row per local adjoint, one column per component of the
        // root's list) holds 0/1 for a branch's own λ, μ and the global tangents of ϵ_0, ϵ_n, cx, cy otherwise
        __syncthreads();
        for (int k = tid; k < KR; k += NT) {
            double gk;
            if (k == 0) {
                gk = log(Lv);
            } else {
                gk = 0.0;
                const double* J = PG.jac + k;
                for (int r = 0; r < 8 * nn; r++) {
                    const double z = zloc[r];
                    if (z != 0.0) gk = fma(z, __ldg(J + (size_t)r * KR), gk);
                }
            }
            A.out_fam[(size_t)fam * KR + k] = gk;
        }
        if (A.tim && tid == 0) {
            const long long te = CLOCK64();
            long long* T = A.tim + (size_t)fam * TIMW;
            T[0] = tcA - tc0;                       // prologue
            T[1] = tcB - tcA;                       // leaf phase
            T[2] = (tcC - tcB) - acc_fsl - acc_froot;  // forward: staging + row 1 of internal/WGD nodes
            T[3] = (tcE - tcD) - acc_bsl;           // backward: row 1 of internal/WGD nodes, transposed
            T[4] = acc_fsl + acc_bsl;               // slices, forward + transposed
            T[5] = acc_froot + (tcD - tcC);         // root, forward + transposed
            T[6] = te - tc0;                        // total
            T[7] = te - tcE;                        // contraction with the table tangents
            if (M.ninner < TIMN) {                  // spare per-node slots: the split of [4] and [5]
                T[8 + TIMN - 1] = acc_fsl;
                T[8 + 2 * TIMN - 1] = acc_froot;
            }
        }
    }
    if (A.done) {
        TailArgs TA{A.done, A.n_total, root, PG.K, KmaxG, A.out_fam, PG.cond, A.cond_kind, 1, A.out, PG.act, A.peer};
        dp_tail_reduce<NT>(TA, nfam_done, reinterpret_cast<double*>(smem_raw));
    }
}
