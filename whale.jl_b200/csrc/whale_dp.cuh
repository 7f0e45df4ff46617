// K2 — the ALE recursion with forward tangents.  src/core.jl:83-199.
//
// One CTA (NT threads) per family.  The LAST row of every branch (C_e × K_e doubles: K_e = 1 + #raw
// parameters that can influence branch e) stays in shared memory for the parent; slices ping-pong between
// that row and a scratch row.
//
// Work decomposition: one LANE per (clade cell, component k) of the row being formed.  The lane walks (its
// share of) the cell's clade-split terms Σ_t p_t·X[γ1]·Y[γ2] (Πduplication, Πspeciation, Πwgdretention, Πroot)
// and forms ITS component with the product rule (value lanes k = 0; tangent lanes k > 0 need the operands'
// value and their own component only), folds it into the row formula's linear part — a scalar — and writes
// one double.  No products go through shared memory and K is a run-time number (one code path for every
// tangent plan).  Where a cell has many terms (the slice loop's heavy clades, the root's last levels with the
// ubiquitous clade) its terms are split over a team of 2^g adjacent lanes and the scalar partial results are
// shuffle-reduced.
//   phase A  leaf branches are independent of each other: one WARP per leaf branch, warp-level sync only;
//            branches whose compatible clades are all leaf clades are family-independent and are filled
//            from the table k_tables prepared (ℓ_n = leafℙ·Πϕ_i); small in-paralog clades use the closed
//            form over tree shapes (leaf-shape CTAs of k_tables).
//   phase B  internal / WGD / root nodes in the reference's order (children first); the node's lists and
//            ϕ/ψ rows are staged in shared memory; only as many warps as the row has lanes take part in the
//            slice loop (one warp: warp-level sync; several: a named barrier among them).
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
    long long* tim;    // optional per-family phase cycle counters [F][TIMW] (profiling aid) or nullptr
    // fused K3: the CTA that finishes last sums the per-family outputs (fixed order), subtracts N·condition and
    // scatters the plan's components into `out`  (src/core.jl:54,63 ; src/condition.jl).  done == nullptr: off
    unsigned int* done;  // CTAs of this plan finished so far (all bins); re-armed to 0 by the last one
    int n_total;         // F
    int cond_kind, first;
    double* out;         // [1+P]
    PeerDev* peer;       // fused exchange over ranks behind the fused reduction (nullptr: off)
};
// tim layout: [0..7] phases, [8 + oi] slice-loop cycles of inner node oi, [8 + TIMN + oi] staging + row-1 cycles
constexpr int TIMN = 32, TIMW = 8 + 2 * TIMN;

#ifdef WHALE_EMU
#define PREFETCH_L2(p) ((void)0)
#define PREFETCH_L1(p) ((void)0)
#else
#define PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#define PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#endif

// copy n16 16-byte words global -> shared with all threads of the scope.  cp.async keeps every copy of a
// staging batch in flight at once (one memory latency per node instead of one per loop iteration); the batch
// is completed by stage_wait() before the scope's barrier.
#ifdef WHALE_EMU
__device__ __forceinline__ void copy16(uint4* dst, const uint4* __restrict__ src, int n16, int tid, int nt) {
    for (int i = tid; i < n16; i += nt) dst[i] = src[i];
}
__device__ __forceinline__ void stage_wait() {}
__device__ __forceinline__ void stage_commit() {}
__device__ __forceinline__ void stage_wait_prev() {}
#else
__device__ __forceinline__ void copy16(uint4* dst, const uint4* __restrict__ src, int n16, int tid, int nt) {
    for (int i = tid; i < n16; i += nt)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + i)),
                     "l"(src + i) : "memory");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// double-buffered batches: commit the batch just issued / wait for everything but the newest batch
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void stage_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
#endif

// ---- bulk asynchronous copies (cp.async.bulk: the copy engine, one issuing thread) completed on an mbarrier ----
#ifdef WHALE_EMU
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) { *bar = 0; (void)count; }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long*, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(unsigned long long*, unsigned) { __syncthreads(); }  // orders the fibers behind the issuer
__device__ __forceinline__ void fence_proxy_async() {}
#else
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WHALE_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WHALE_MBAR_DONE;\n"
        "bra WHALE_MBAR_WAIT;\n"
        "WHALE_MBAR_DONE:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

#ifdef WHALE_EMU
#define SHFL_DOWN(v, d) emu::shfl_down(v, d)
#define WARP_ANY(p) emu::warp_any(p)
#else
#define SHFL_DOWN(v, d) __shfl_down_sync(0xffffffffu, v, d)
#define WARP_ANY(p) (__any_sync(0xffffffffu, (p)) != 0)
#endif

// barrier among the first `nw` warps of the CTA (the warps that own lanes of the current row)
template <int NW>
__device__ __forceinline__ void part_sync(int nw) {
#ifdef WHALE_EMU
    if (NW == 1) __syncwarp();
    else __syncthreads();  // the emulation keeps every warp in the loop
#else
    if (NW == 1 || nw == 1) __syncwarp();
    else if (nw == NW) __syncthreads();
    else asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
#endif
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

// Σ_t p_t·X[i1]·Y[i2] over terms tb, tb+step, ... < te for ONE component: value s0 and the lane's own
// component sk (product rule).  kx/ky: the component's index in X/Y (−1: that row does not depend on the
// parameter -> zero tangent).  Two terms are in flight per iteration so the loads of one overlap the
// arithmetic of the other.
template <bool GLOBAL_ENTS>
__device__ __forceinline__ void term_sum(const Ent* __restrict__ ents, uint32_t tb, uint32_t te, uint32_t step,
                                         const double* __restrict__ X, int KX, int kx, const double* __restrict__ Y,
                                         int KY, int ky, double& s0, double& sk) {
    double a0 = 0.0, ak = 0.0, b0 = 0.0, bk = 0.0;
    uint32_t t = tb;
    for (; t + step < te; t += 2 * step) {
        uint4 r1, r2;
        if (GLOBAL_ENTS) { r1 = __ldg(reinterpret_cast<const uint4*>(ents + t)); r2 = __ldg(reinterpret_cast<const uint4*>(ents + t + step)); }
        else { r1 = *reinterpret_cast<const uint4*>(ents + t); r2 = *reinterpret_cast<const uint4*>(ents + t + step); }
        const double* x1 = X + (r1.x & 0xffffu) * KX;
        const double* y1 = Y + (r1.x >> 16) * KY;
        const double* x2 = X + (r2.x & 0xffffu) * KX;
        const double* y2 = Y + (r2.x >> 16) * KY;
        const double x10 = x1[0], y10 = y1[0], x20 = x2[0], y20 = y2[0];
        const double x1k = kx >= 0 ? x1[kx] : 0.0, y1k = ky >= 0 ? y1[ky] : 0.0;
        const double x2k = kx >= 0 ? x2[kx] : 0.0, y2k = ky >= 0 ? y2[ky] : 0.0;
        const double p1 = __hiloint2double((int)r1.w, (int)r1.z), p2 = __hiloint2double((int)r2.w, (int)r2.z);
        const double px1 = p1 * x10, px2 = p2 * x20;
        a0 = fma(px1, y10, a0);
        b0 = fma(px2, y20, b0);
        ak = fma(px1, y1k, fma(p1 * y10, x1k, ak));
        bk = fma(px2, y2k, fma(p2 * y20, x2k, bk));
    }
    if (t < te) {
        uint4 r1;
        if (GLOBAL_ENTS) r1 = __ldg(reinterpret_cast<const uint4*>(ents + t));
        else r1 = *reinterpret_cast<const uint4*>(ents + t);
        const double* x1 = X + (r1.x & 0xffffu) * KX;
        const double* y1 = Y + (r1.x >> 16) * KY;
        const double x10 = x1[0], y10 = y1[0];
        const double x1k = kx >= 0 ? x1[kx] : 0.0, y1k = ky >= 0 ? y1[ky] : 0.0;
        const double p1 = __hiloint2double((int)r1.w, (int)r1.z);
        const double px1 = p1 * x10;
        a0 = fma(px1, y10, a0);
        ak = fma(px1, y1k, fma(p1 * y10, x1k, ak));
    }
    s0 = a0 + b0;
    sk = ak + bk;
}

// ---------------------------------------------------------------------------------------------------------
// The n slices of one branch (src/core.jl:121-128,178-185):
//   ℓ_i[γ] = ϕ_i ℓ_{i−1}[γ] + ψ_i Σ_t p_t ℓ_{i−1}[γ1] ℓ_{i−1}[γ2]        (+ tangents)
// `cur` holds row 1 on entry; rows alternate between fin and scr so that row n+1 lands in fin.
// Lanes = (slot of the packer's lane table) × (group of KC components), group-major in blocks of `spad` lanes,
// so a team's slots are adjacent lanes.  KC is chosen per branch so that the row needs at most two lane sets
// per thread: short rows get one component per lane (the shortest dependent chain), long rows up to three.
// Every lane keeps the row offsets and probabilities of its (at most two) terms in registers for the whole
// branch and runs the same branch-free code: operand loads, the product rule for its components, the row
// formula's linear part (ψ_i·Σ_k + ψ'_i·Σ_0) folded into one scalar per component, a shuffle reduction of those
// scalars inside the team, and the leader's ϕ part + store.  The two lane sets of a thread are issued together
// so their latencies overlap.
// ---------------------------------------------------------------------------------------------------------
struct LaneSK {
    int c0k;         // offset (elements) of the cell's value component
    int kb;          // first component of the lane's group
    int gsz;         // team size (0: idle lane)
    int lead;        // team leader: writes the cell
    int cnt, first;  // the lane's share of the cell's terms: first, first+gsz, ...
    int a1, a2, b1, b2;  // row offsets (elements) of the operands of its two terms
    double pa, pb;
};

__device__ __forceinline__ LaneSK load_lane(const Slot* s_slots, int nslots, const Ent* s_dents, int spad, int K,
                                            int KC, int total, int u) {
    LaneSK w;
    w.c0k = 0; w.kb = 0; w.gsz = 0; w.lead = 0; w.cnt = 0; w.first = 0;
    w.a1 = w.a2 = w.b1 = w.b2 = 0;
    w.pa = w.pb = 0.0;
    if (u < total) {
        const int g = u / spad, s = u - g * spad;
        if (s < nslots) {
            const Slot sl = s_slots[s];
            w.kb = g * KC;
            w.gsz = 1 << sl.glog;
            w.lead = (s & (w.gsz - 1)) == 0;
            w.c0k = (int)sl.cell * RS(K);
            w.cnt = sl.cnt; w.first = sl.first;
            if (w.cnt > 0) { const Ent en = s_dents[w.first]; w.a1 = en.i1 * RS(K); w.a2 = en.i2 * RS(K); w.pa = en.p; }
            if (w.cnt > 1) { const Ent en = s_dents[w.first + w.gsz]; w.b1 = en.i1 * RS(K); w.b2 = en.i2 * RS(K); w.pb = en.p; }
        }
    }
    return w;
}

// team size of the warp's first lane = the warp's largest (slots are sorted by team size, descending)
__device__ __forceinline__ int warp_team(const Slot* s_slots, int nslots, int spad, int total, int u0) {
    if (u0 >= total) return 0;
    if (spad < 32) return 1 << s_slots[0].glog;
    const int s0 = u0 % spad;
    return s0 < nslots ? (1 << s_slots[s0].glog) : 0;
}

// the lane's partials of ψ_i·Σ_k + ψ'_i·Σ_0 for one slice (branch-free for cnt <= 2)
template <int KC>
__device__ __forceinline__ void lane_partial(const LaneSK& w, const double* __restrict__ src, double2 c0,
                                             const double2 (&ck)[KC], const Ent* s_dents, int K, double (&part)[KC]) {
    const double* xa = src + w.a1;
    const double* ya = src + w.a2;
    const double* xb = src + w.b1;
    const double* yb = src + w.b2;
    const double x0 = xa[0], y0 = ya[0], u0 = xb[0], v0 = yb[0];
    double xk[KC], yk[KC], uk[KC], vk[KC];
#pragma unroll
    for (int j = 0; j < KC; j++) {
        const int k = min(w.kb + j, K - 1);
        xk[j] = xa[k]; yk[j] = ya[k]; uk[j] = xb[k]; vk[j] = yb[k];
    }
    const double px = w.pa * x0, py = w.pa * y0, pu = w.pb * u0, pv = w.pb * v0;
    double s0 = fma(px, y0, pu * v0);
    double sk[KC];
#pragma unroll
    for (int j = 0; j < KC; j++) sk[j] = fma(px, yk[j], py * xk[j]) + fma(pu, vk[j], pv * uk[j]);
    if (w.cnt > 2) {  // clades with more than 64 same-branch terms
#pragma unroll
        for (int j = 0; j < KC; j++) {
            double r0, rk;
            term_sum<false>(s_dents, (uint32_t)(w.first + 2 * w.gsz), (uint32_t)(w.first + w.cnt * w.gsz), (uint32_t)w.gsz,
                            src, RS(K), min(w.kb + j, K - 1), src, RS(K), min(w.kb + j, K - 1), r0, rk);
            if (j == 0) s0 += r0;
            sk[j] += rk;
        }
    }
#pragma unroll
    for (int j = 0; j < KC; j++) part[j] = (w.kb + j == 0) ? c0.y * s0 : fma(c0.y, sk[j], ck[j].y * s0);
}

template <int KC>
__device__ __forceinline__ void lane_finish(const LaneSK& w, int wg, double (&part)[KC], const double* __restrict__ src,
                                            double* __restrict__ dst, double2 c0, const double2 (&ck)[KC], int K, int C,
                                            int i, double* ellp) {
    const double o0 = src[w.c0k];
    double ok[KC];
#pragma unroll
    for (int j = 0; j < KC; j++) ok[j] = src[w.c0k + min(w.kb + j, K - 1)];
    for (int step = 1; step < wg; step <<= 1) {
#pragma unroll
        for (int j = 0; j < KC; j++) {
            const double t = SHFL_DOWN(part[j], step);
            if (step < w.gsz) part[j] += t;
        }
    }
    if (w.lead) {
#pragma unroll
        for (int j = 0; j < KC; j++) {
            const int k = w.kb + j;
            if (k < K) {
                const double r = k == 0 ? fma(c0.x, o0, part[j]) : fma(c0.x, ok[j], fma(ck[j].x, o0, part[j]));
                dst[w.c0k + k] = r;
                if (ellp && k == 0) ellp[(size_t)i * C + w.c0k / RS(K)] = r;
            }
        }
    }
}

// NWB: warps of the scope (1: a warp working alone — leaf branches; else the CTA's warp count).
// PPG: the ϕ/ψ rows are read from global memory (prefetched one slice ahead into registers) instead of the
// staging buffer.  KC: components per lane.
template <int NWB, bool PPG, int KC>
__device__ __noinline__ void run_slices(int n, int C, int K, double* fin, double* scr, double* cur,
                                           const Slot* s_slots, int nslots, const Ent* s_dents, const double2* pprow,
                                           double* ellp, int tid, int spad) {
    constexpr int nt = NWB * 32;
    const int total = spad * ((K + KC - 1) / KC);
#ifdef WHALE_EMU
    const int nwa = NWB;
#else
    const int nwa = min(NWB, (total + 31) >> 5);
#endif
    if ((tid >> 5) >= nwa) return;  // this warp owns no lane of the row
    const int npass = (total + nt - 1) / nt;
    const int wbase = tid & ~31;
    const LaneSK w0 = load_lane(s_slots, nslots, s_dents, spad, K, KC, total, tid);
    const LaneSK w1 = load_lane(s_slots, nslots, s_dents, spad, K, KC, total, tid + nt);
    const int wg0 = warp_team(s_slots, nslots, spad, total, wbase);
    const int wg1 = warp_team(s_slots, nslots, spad, total, wbase + nt);
    int k0[KC], k1[KC];
#pragma unroll
    for (int j = 0; j < KC; j++) { k0[j] = min(w0.kb + j, K - 1); k1[j] = min(w1.kb + j, K - 1); }
    double2 n0 = make_double2(0.0, 0.0), nk0[KC], nk1[KC];
    if (PPG) {
#pragma unroll
        for (int j = 0; j < KC; j++) { nk0[j] = n0; nk1[j] = n0; }
        if (n >= 1) {
            n0 = __ldg(pprow + K);
#pragma unroll
            for (int j = 0; j < KC; j++) { nk0[j] = __ldg(pprow + K + k0[j]); nk1[j] = __ldg(pprow + K + k1[j]); }
        }
    }
    for (int i = 1; i <= n; i++) {
        const double* src = cur;
        double* dst = (cur == fin) ? scr : fin;
        const double2* ppi = pprow + (size_t)i * K;
        double2 c0, ck0[KC], ck1[KC];
        if (PPG) {
            c0 = n0;
#pragma unroll
            for (int j = 0; j < KC; j++) { ck0[j] = nk0[j]; ck1[j] = nk1[j]; }
            if (i < n) {
                n0 = __ldg(ppi + K);
#pragma unroll
                for (int j = 0; j < KC; j++) { nk0[j] = __ldg(ppi + K + k0[j]); nk1[j] = __ldg(ppi + K + k1[j]); }
            }
        } else {
            c0 = ppi[0];
#pragma unroll
            for (int j = 0; j < KC; j++) { ck0[j] = ppi[k0[j]]; ck1[j] = ppi[k1[j]]; }
        }
        if (wg1) {  // both lane sets: issue the loads of both before either reduction
            double p0[KC], p1[KC];
            lane_partial<KC>(w0, src, c0, ck0, s_dents, K, p0);
            lane_partial<KC>(w1, src, c0, ck1, s_dents, K, p1);
            lane_finish<KC>(w0, wg0, p0, src, dst, c0, ck0, K, C, i, ellp);
            lane_finish<KC>(w1, wg1, p1, src, dst, c0, ck1, K, C, i, ellp);
        } else if (wg0) {
            double p0[KC];
            lane_partial<KC>(w0, src, c0, ck0, s_dents, K, p0);
            lane_finish<KC>(w0, wg0, p0, src, dst, c0, ck0, K, C, i, ellp);
        }
        for (int q = 2; q < npass; q++) {  // very long rows: descriptors reloaded from shared memory
            const int wgq = warp_team(s_slots, nslots, spad, total, wbase + q * nt);
            if (wgq == 0) continue;
            const LaneSK wq = load_lane(s_slots, nslots, s_dents, spad, K, KC, total, tid + q * nt);
            double2 ckq[KC];
            double pq[KC];
#pragma unroll
            for (int j = 0; j < KC; j++) ckq[j] = PPG ? __ldg(ppi + min(wq.kb + j, K - 1)) : ppi[min(wq.kb + j, K - 1)];
            lane_partial<KC>(wq, src, c0, ckq, s_dents, K, pq);
            lane_finish<KC>(wq, wgq, pq, src, dst, c0, ckq, K, C, i, ellp);
        }
        cur = dst;
        part_sync<NWB>(nwa);
    }
}

// components per lane: the fewest that keep the row within two lane sets per thread (at most 3)
template <int NWB, bool PPG>
__device__ __forceinline__ void run_slices_kc(int n, int C, int K, double* fin, double* scr, double* cur,
                                              const Slot* s_slots, int nslots, const Ent* s_dents, const double2* pprow,
                                              double* ellp, int tid) {
    int spad = (nslots + 31) & ~31;
    if (nslots < 32) { spad = 1; while (spad < nslots) spad <<= 1; }
    const int lanes2 = 2 * NWB * 32;
    if (spad * K <= lanes2) run_slices<NWB, PPG, 1>(n, C, K, fin, scr, cur, s_slots, nslots, s_dents, pprow, ellp, tid, spad);
    else if (spad * ((K + 1) >> 1) <= lanes2) run_slices<NWB, PPG, 2>(n, C, K, fin, scr, cur, s_slots, nslots, s_dents, pprow, ellp, tid, spad);
    else run_slices<NWB, PPG, 3>(n, C, K, fin, scr, cur, s_slots, nslots, s_dents, pprow, ellp, tid, spad);
}

template <bool WARP>
__device__ __forceinline__ void scope_sync() {
    if (WARP) __syncwarp();
    else __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// Fused slice loop for K <= 8 components (the common case: constant rates, or few branch-wise parameters).
// Each lane owns one Slot of the packer's lane table for the whole branch and keeps its (up to two) terms in
// registers, so a slice is: gather 2·K operands per term from the previous row, K fused multiply-adds per
// term, a shuffle reduction inside teams that share a heavy clade, the ϕ/ψ update, one barrier.
// ---------------------------------------------------------------------------------------------------------
constexpr int KMAX_FUSED = 8;

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
        if (w.cnt > 0) { const Ent en = s_dents[w.first]; w.a1 = en.i1 * RS(K); w.a2 = en.i2 * RS(K); w.pa = en.p; }
        if (w.cnt > 1) { const Ent en = s_dents[w.first + w.gsz]; w.b1 = en.i1 * RS(K); w.b2 = en.i2 * RS(K); w.pb = en.p; }
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

// one pass of one slice for this lane; wg = team size of the warp's first lane (warp-uniform, 0: nothing to do).
// two = some lane of the warp owns a second term, more = some lane owns more than two (both warp-uniform, fixed
// for the branch): the first two terms run branch-free for every lane (an absent term has p = 0 and reads cell 0),
// so the 4·K operand loads of a slice are in flight together and the warp never diverges before the reduction.
template <int K>
__device__ __forceinline__ void slice_pass(const LaneWork<K>& w, int wg, bool two, bool more, int sidx,
                                           const double* __restrict__ src, double* __restrict__ dst, const double2* ppi,
                                           const Ent* s_dents, int C, int i, double* ellp) {
    if (wg == 0) return;
    double s[K];
#pragma unroll
    for (int k = 0; k < K; k++) s[k] = 0.0;
#ifdef WHALE_SLICE_V1
    if (w.cnt > 0) accum<K>(src, w.a1, w.a2, w.pa, s);
    if (w.cnt > 1) accum<K>(src, w.b1, w.b2, w.pb, s);
    for (int j = 2; j < w.cnt; j++) {
        const Ent en = s_dents[w.first + j * w.gsz];
        accum<K>(src, en.i1 * RS(K), en.i2 * RS(K), en.p, s);
    }
    for (int step = 1; step < wg; step <<= 1) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            const double t = SHFL_DOWN(s[k], step);
            if (step < w.gsz) s[k] += t;
        }
    }
#else
    if (two) {
        const double* xa = src + w.a1;
        const double* ya = src + w.a2;
        const double* xb = src + w.b1;
        const double* yb = src + w.b2;
        const double pxa = w.pa * xa[0], pya = w.pa * ya[0], pxb = w.pb * xb[0], pyb = w.pb * yb[0];
        s[0] = fma(pxb, yb[0], pxa * ya[0]);
#pragma unroll
        for (int k = 1; k < K; k++) s[k] = fma(pxa, ya[k], pya * xa[k]) + fma(pxb, yb[k], pyb * xb[k]);
    } else {
        accum<K>(src, w.a1, w.a2, w.pa, s);
    }
    if (more) {
        for (int j = 2; j < w.cnt; j++) {
            const Ent en = s_dents[w.first + j * w.gsz];
            accum<K>(src, en.i1 * RS(K), en.i2 * RS(K), en.p, s);
        }
    }
    // team reduction: lane j adds lane j+step while step is inside its own team.  The mask multiplies instead of
    // selecting (one DFMA per component and step; fma(t, 1, s) is the correctly rounded s + t)
#define TEAM_REDUCE()                                                               \
    for (int step = 1; step < wg; step <<= 1) {                                     \
        const double m = step < w.gsz ? 1.0 : 0.0;                                  \
        _Pragma("unroll") for (int k = 0; k < K; k++) s[k] = fma(SHFL_DOWN(s[k], step), m, s[k]); \
    }
#ifndef WHALE_SLICE_NOHOIST
    // the leader's operands (ϕ_i, ψ_i with tangents, the cell's previous value) are fetched before the reduction so
    // their latency hides behind the shuffles; every lane loads (valid addresses), only leaders use them.  (Letting
    // the non-leaders read one broadcast cell instead of their own measured 1 % slower.)
    const bool lead = w.cell >= 0 && (sidx & (w.gsz - 1)) == 0;
    const int cK = max(w.cell, 0) * RS(K);
    double2 pk[K];
    double o[K];
#pragma unroll
    for (int k = 0; k < K; k++) { pk[k] = ppi[k]; o[k] = src[cK + k]; }
    TEAM_REDUCE()
    if (lead) {  // team leader: ℓ_i = ϕ_i ℓ_{i−1} + ψ_i Σ  (with tangents)
        const double r0 = fma(pk[0].x, o[0], pk[0].y * s[0]);
        dst[cK] = r0;
        if (ellp) ellp[(size_t)i * C + w.cell] = r0;
#pragma unroll
        for (int k = 1; k < K; k++)
            dst[cK + k] = fma(pk[0].x, o[k], fma(pk[0].y, s[k], fma(pk[k].x, o[0], pk[k].y * s[0])));
    }
    return;
#else
    TEAM_REDUCE()
#endif
#undef TEAM_REDUCE
#endif
    if (w.cell >= 0 && (sidx & (w.gsz - 1)) == 0) {  // team leader: ℓ_i = ϕ_i ℓ_{i−1} + ψ_i Σ  (with tangents)
        const int c = w.cell;
        const double2 c0 = ppi[0];
        const double o0 = src[c * RS(K)];
        const double r0 = fma(c0.x, o0, c0.y * s[0]);
        dst[c * RS(K)] = r0;
        if (ellp) ellp[(size_t)i * C + c] = r0;
#pragma unroll
        for (int k = 1; k < K; k++) {
            const double2 ck = ppi[k];
            dst[c * RS(K) + k] = fma(c0.x, src[c * RS(K) + k], fma(c0.y, s[k], fma(ck.x, o0, ck.y * s[0])));
        }
    }
}

// Not inlined on purpose: each component count gets its own register allocation; inlining all variants into
// k_dp makes ptxas demote the per-lane accumulator arrays to local memory.
template <int K, bool WARP>
__device__ __noinline__ void run_slices_fused(int n, int C, double* fin, double* scr, double* cur,
                                                 const Slot* s_slots, int nslots, const Ent* s_dents,
                                                 const double2* pprow, double* ellp, int tid, int nt) {
    // every warp of the scope takes part in the per-slice barrier, also those that own no slot of the row: letting
    // them skip the loop (named barrier / warp-level sync among the owners) measured 10 % SLOWER on the B200
    // the first two passes keep their descriptors in registers
    const int wbase = tid & ~31;
    const LaneWork<K> w0 = load_work<K>(s_slots, nslots, s_dents, tid);
    const LaneWork<K> w1 = load_work<K>(s_slots, nslots, s_dents, tid + nt);
    // slots are sorted by team size (descending): the warp's first lane carries the warp's largest team
    const int wg0 = wbase < nslots ? (1 << s_slots[wbase].glog) : 0;
    const int wg1 = (wbase + nt) < nslots ? (1 << s_slots[wbase + nt].glog) : 0;
    const int npass = (nslots + nt - 1) / nt;
    const bool two0 = WARP_ANY(w0.cnt > 1), more0 = WARP_ANY(w0.cnt > 2);
    const bool two1 = WARP_ANY(w1.cnt > 1), more1 = WARP_ANY(w1.cnt > 2);
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
        slice_pass<K>(w0, wg0, two0, more0, tid, src, dst, ppi, s_dents, C, i, ellp);
        slice_pass<K>(w1, wg1, two1, more1, tid + nt, src, dst, ppi, s_dents, C, i, ellp);
        for (int q = 2; q < npass; q++) {  // oversized rows: descriptors reloaded from shared memory
            const LaneWork<K> wq = load_work<K>(s_slots, nslots, s_dents, tid + q * nt);
            const int wgq = (wbase + q * nt) < nslots ? (1 << s_slots[wbase + q * nt].glog) : 0;
            slice_pass<K>(wq, wgq, WARP_ANY(wq.cnt > 1), WARP_ANY(wq.cnt > 2), tid + q * nt, src, dst, ppi, s_dents, C, i, ellp);
        }
        cur = dst;
        scope_sync<WARP>();
    }
}

// dispatch on the branch's component count (leaf branches carry at most value + own λ, μ: K <= 3)
template <bool WARP>
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
        switch (K) {
            CASEK(1) CASEK(2) CASEK(3) CASEK(4) CASEK(5) CASEK(6) CASEK(7) CASEK(8)
            default: return false;
        }
    }
#undef CASEK
}

// ---------------------------------------------------------------------------------------------------------
// Fused K3.  Every CTA publishes its family's (log L, ∇) and bumps a counter; the CTA that sees the last count
// reduces: thread (r, k) sums component k over the families f ≡ r (mod NT/KR), thread k adds those partial sums —
// the order of the additions depends on (F, NT, KR) only, never on which CTA happens to be last (deterministic bits).
// ---------------------------------------------------------------------------------------------------------
#ifdef WHALE_EMU
#define LDCG(p) (*(p))
#else
#define LDCG(p) __ldcg(p)
#endif
struct TailArgs {
    unsigned int* done;     // families finished so far (all launches of this evaluation); re-armed to 0 by the last CTA
    int n_total;            // F
    int root;
    const int* K;           // the output plan's component counts
    int Kmax;
    const double* out_fam;  // [F * K[root]]
    const double* cond;     // [4 * Kmax]
    int cond_kind, first;
    double* out;            // [1+P]
    const int* act;         // [nn * Kmax]
    PeerDev* peer;          // one process per GPU: the sum over ranks follows in the same CTA (nullptr: off)
};
// `count` = families this CTA finished (1 for k_dp; the persistent CTAs of k_dp_rev report their share on exit)
template <int NT>
__device__ __forceinline__ void dp_tail_reduce(const TailArgs& A, int count, double* s_tot) {
    constexpr int NW = NT / 32;  // (no static shared memory here: k_dp opts in to the full 227 KB dynamically)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __syncthreads();  // this CTA's outputs are written
    int last = 0;
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(A.done, (unsigned)count);
        last = (prev + (unsigned)count == (unsigned)A.n_total) ? 1 : 0;
    }
    if (!__syncthreads_or(last)) return;
    __threadfence();
    const int root = A.root, KR = A.K[root], F = A.n_total, Kmax = A.Kmax;
    if (KR <= NT / 2) {
        // R = NT/KR row groups: thread (r, k) sums component k of the families f ≡ r (mod R), eight loads in flight;
        // thread k then adds the R partial sums in order
        const int R = NT / KR, r = tid / KR, k = tid - r * KR;
        double* part = s_tot + KR;  // [R*KR]
        if (r < R) {
            const double* p = A.out_fam + (size_t)r * KR + k;
            const size_t st = (size_t)R * KR;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0, a7 = 0.0;
            int f = r;
            for (; f + 7 * R < F; f += 8 * R, p += 8 * st) {
                a0 += LDCG(p); a1 += LDCG(p + st); a2 += LDCG(p + 2 * st); a3 += LDCG(p + 3 * st);
                a4 += LDCG(p + 4 * st); a5 += LDCG(p + 5 * st); a6 += LDCG(p + 6 * st); a7 += LDCG(p + 7 * st);
            }
            for (; f < F; f += R, p += st) a0 += LDCG(p);
            part[tid] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
        }
        __syncthreads();
        if (tid < KR) {
            double s = 0.0;
            for (int q = 0; q < R; q++) s += part[q * KR + tid];
            s_tot[tid] = s - (double)F * A.cond[A.cond_kind * Kmax + tid];
        }
    } else {
        for (int k = warp; k < KR; k += NW) {
            double s = 0.0;
            for (int f = lane; f < F; f += 32) s += LDCG(A.out_fam + (size_t)f * KR + k);
            for (int step = 16; step > 0; step >>= 1) s += SHFL_DOWN(s, step);
            if (lane == 0) s_tot[k] = s - (double)F * A.cond[A.cond_kind * Kmax + k];
        }
    }
    __syncthreads();
    const bool finite = isfinite(s_tot[0]);  // ℓhood src/core.jl:15
    // `out` was zeroed by the host; every gradient pass (parameter chunk) writes its own parameters, the first
    // pass also the log-likelihood
    for (int k = tid; k < KR; k += NT) {
        if (k == 0) { if (A.first) A.out[0] = finite ? s_tot[0] : -dinf(); }
        else A.out[1 + A.act[root * Kmax + k]] = finite ? s_tot[k] : 0.0;
    }
    if (tid == 0) *A.done = 0u;
    if (A.peer) {  // WHALE_PEER_SUM of a one-pass evaluation: `out` is complete, exchange it right here
        __syncthreads();
        peer_exchange<NT>(A.peer, A.out, reinterpret_cast<unsigned*>(s_tot));
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_dp(DPArgs A, int perm_off) {
    EXTERN_SHARED(smem_raw);
    constexpr int NW = NT / 32;
    // Warps are bound to the SM's four schedulers by their index, and the work of a row fills warps from the
    // first one up, so co-resident families would all load the same scheduler: rotate the warp roles per CTA.
    const int lane = threadIdx.x & 31;
    const int warp = (int)(((threadIdx.x >> 5) + ((blockIdx.x * 2654435761u) >> 13)) % NW);
    const int tid = warp * 32 + lane;
    const ModelDev& M = A.M;
    const PlanDev& PL = A.PL;
    const int nn = M.nn, Kmax = PL.Kmax;
    const int fam = A.perm[perm_off + blockIdx.x];
    const FamHdr* Hp = A.hdr + fam;
    const uint64_t base = Hp->base;
    const uint32_t nlev = Hp->nlev, blob_bytes = Hp->blob_bytes;
    const uint32_t rows_len = Hp->rows_len[A.plan], scr_len = Hp->scr_len[A.plan];
    const uint32_t leafmax = Hp->leafmax[A.plan];
    const uint32_t stage_bytes = Hp->stage_bytes[A.plan], leaf_stage = Hp->leaf_stage;
    const unsigned char* blob = A.arena + base;
    const NodeRec* const g_nrec = reinterpret_cast<const NodeRec*>(blob);
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
    const size_t meta_bytes = (((7 * nn + 1) * sizeof(int) + (size_t)nn * 2 * Kmax * sizeof(int16_t)) + 15) & ~size_t(15);
    // the family's node records (48 B per node, the head of its blob): one coalesced read here instead of a
    // dependent L2 round trip at every node of phase B
    NodeRec* s_nrec = reinterpret_cast<NodeRec*>(smem_raw + meta_bytes);
    const NodeRec* const nrec = s_nrec;
    const size_t hdr_bytes = meta_bytes + (size_t)nn * sizeof(NodeRec);
    double* rows = reinterpret_cast<double*>(smem_raw + hdr_bytes);
    double* scr = rows + rows_len;
    unsigned char* stage = reinterpret_cast<unsigned char*>(scr + scr_len);
    unsigned char* leaf_area = stage + stage_bytes;
    const size_t leaf_area_bytes = (size_t)leafmax * sizeof(double) + leaf_stage;
    // families whose lists are too long to stage (large CCDs) read them from global memory (L2) in place
    const bool staged = stage_bytes != 0;
    for (int i = tid; i < nn; i += NT) {
        s_kind[i] = M.kind[i]; s_nsl[i] = M.nsl[i]; s_ch0[i] = M.child0[i]; s_ch1[i] = M.child1[i];
        s_K[i] = PL.K[i]; s_toff[i] = PL.toff[i];
        s_roff[i] = (int)A.roff[(size_t)fam * nn + i];
    }
    for (int i = tid; i < nn * 2 * Kmax; i += NT) s_cmap[i] = PL.cmap[i];
    for (int i = tid; i < nn * 3; i += NT)
        reinterpret_cast<uint4*>(s_nrec)[i] = __ldg(reinterpret_cast<const uint4*>(g_nrec) + i);
    __syncthreads();
    double* const ell_base = A.ell ? A.ell + Hp->ell_off : nullptr;
    auto ell_of = [&](int e) -> double* {  // node e's matrix inside the family's ℓ (node-index order)
        if (!ell_base) return nullptr;
        size_t o = 0;
        for (int e2 = 0; e2 < e; e2++) o += (size_t)(s_nsl[e2] + 1) * nrec[e2].C;
        return ell_base + o;
    };

    GRID_DEP_WAIT();  // programmatic dependent launch: everything above overlapped the table kernel
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
            // closed form over tree shapes: ℓ_n[γ] = Σ_σ C_σ[γ]·wσ_n  (C from the packer, w from the leaf-shape CTAs of k_tables)
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
                fin[c * RS(K) + k] = v;
            }
            continue;
        }
        if (R.nslots > HEAVY_SLOTS && !(R.nonleaf == 0 && A.skip_leaf)) continue;  // heavy branch: whole CTA, below
        if (R.nonleaf == 0 && A.skip_leaf) {  // family-independent: ℓ_n = leafℙ·Πϕ_i from k_tables
            for (int i = lane; i < C * K; i += 32) fin[(i / K) * RS(K) + (i % K)] = PL.leaf[e * Kmax + (i % K)];
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
            cur[c * RS(K) + k] = v;
            if (ellp && k == 0) ellp[c] = v;
        }
        stage_wait();
        __syncwarp();
        if (!run_slices_fused_k<true>(K, n, C, fin, wscr, cur, reinterpret_cast<const Slot*>(wst + nd16), (int)R.nslots,
                                      reinterpret_cast<const Ent*>(wst), PL.pp + s_toff[e], ellp, lane, 32))
            run_slices_kc<1, true>(n, C, K, fin, wscr, cur, reinterpret_cast<const Slot*>(wst + nd16), (int)R.nslots,
                                   reinterpret_cast<const Ent*>(wst), PL.pp + s_toff[e], ellp, lane);
    }
    __syncthreads();
    // leaf branches with many in-paralog clades (more lanes of work than one warp handles well): whole CTA
    for (int li = 0; li < M.nleafnodes; li++) {
        const int e = M.leafnodes[li];
        const NodeRec R = nrec[e];
        const int C = (int)R.C;
        if (C == 0 || R.nslots <= HEAVY_SLOTS || (R.nonleaf == 0 && A.skip_leaf) || (A.skip_leaf && R.sptr_off)) continue;
        const int K = s_K[e], n = s_nsl[e];
        double* fin = rows + s_roff[e];
        double* ellp = ell_of(e);
        const int nd16 = (int)R.ndent, sl16 = ((int)R.nslots + 1) >> 1, pp16 = (n + 1) * K;
        const Ent* h_dents = ents + R.dent_off;
        const Slot* h_slots = reinterpret_cast<const Slot*>(words + R.slot_off);
        const double2* h_pp = PL.pp + s_toff[e];
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
            const double v = (c < nleafc && k == 0) ? M.leafP[e] : 0.0;
            cur[c * RS(K) + k] = v;
            if (ellp && k == 0) ellp[c] = v;
        }
        stage_wait();
        __syncthreads();
        run_slices_fused_k<false>(K, n, C, fin, scr, cur, h_slots, (int)R.nslots, h_dents, h_pp, ellp, tid, NT);
        __syncthreads();
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
        const bool pps = K <= 8;  // ϕ/ψ rows staged in shared memory (else read from global, prefetched)

        const long long tst = CLOCK64();
        // ---- stage this node's lists: [dents | slots | dptr | tptr,lossF,lossG,lev | tents | ϕψ rows] ----
        const int nd16 = (kind == WHALE_ROOT) ? 0 : (int)R.ndent;  // Πroot terms are read once: stay global
        const int sl16 = (kind == WHALE_ROOT) ? 0 : (((int)R.nslots + 1) >> 1);
        const int dp16 = (C + 1 + 3) >> 2;
        const int tp16 = (kind == WHALE_WGD) ? 0 : ((3 * C + 1 + (kind == WHALE_ROOT ? (int)nlev + 1 : 0) + 3) >> 2);
        const int tn16 = (kind == WHALE_INTERNAL) ? (int)R.ntent : 0;
        const int pp16 = pps ? (n + 1) * K : 0;
        uint4* st4 = reinterpret_cast<uint4*>(stage);
        uint4* st5 = st4 + nd16 + sl16;
        const Ent* s_dents = ents + R.dent_off;
        const Slot* s_slots = reinterpret_cast<const Slot*>(words + R.slot_off);
        const uint32_t* s_dptr = words + R.dptr_off;
        const uint32_t* s_tptr = words + R.tptr_off;
        const Ent* s_tents = ents + R.tent_off;
        const double2* s_pp = PL.pp + s_toff[e];
        if (staged) {
            copy16(st4, reinterpret_cast<const uint4*>(s_dents), nd16, tid, NT);
            copy16(st4 + nd16, reinterpret_cast<const uint4*>(s_slots), sl16, tid, NT);
            copy16(st5, reinterpret_cast<const uint4*>(s_dptr), dp16, tid, NT);
            copy16(st5 + dp16, reinterpret_cast<const uint4*>(s_tptr), tp16, tid, NT);
            copy16(st5 + dp16 + tp16, reinterpret_cast<const uint4*>(s_tents), tn16, tid, NT);
            copy16(st5 + dp16 + tp16 + tn16, reinterpret_cast<const uint4*>(s_pp), pp16, tid, NT);
            stage_wait();
            s_dents = reinterpret_cast<const Ent*>(stage);
            s_slots = reinterpret_cast<const Slot*>(st4 + nd16);
            s_dptr = reinterpret_cast<const uint32_t*>(st5);
            s_tptr = reinterpret_cast<const uint32_t*>(st5 + dp16);
            s_tents = reinterpret_cast<const Ent*>(st5 + dp16 + tp16);
            s_pp = reinterpret_cast<const double2*>(st5 + dp16 + tp16 + tn16);
        }
        const int32_t* s_lossF = reinterpret_cast<const int32_t*>(s_tptr + C + 1);
        const int32_t* s_lossG = s_lossF + C;
        const uint32_t* s_lev = reinterpret_cast<const uint32_t*>(s_lossG + C);
        auto slices = [&](double* cur) {
            const long long ts = CLOCK64();
            acc_row1 += ts - tc1;
            if (A.tim && tid == 0 && oi < TIMN) A.tim[(size_t)fam * TIMW + 8 + TIMN + oi] = ts - tst;
            if (pps) run_slices_fused_k<false>(K, n, C, fin, scr, cur, s_slots, (int)R.nslots, s_dents, s_pp, ellp, tid, NT);
            else run_slices_kc<NW, true>(n, C, K, fin, scr, cur, s_slots, (int)R.nslots, s_dents, PL.pp + s_toff[e], ellp, tid);
            __syncthreads();  // the last row is complete for every warp; the staging buffer may be reused
            const long long tse = CLOCK64();
            acc_slices += tse - ts;
            if (A.tim && tid == 0 && oi < TIMN) A.tim[(size_t)fam * TIMW + 8 + oi] = tse - ts;
        };

        // children, shared by the row-1 formulas
        tc1 = CLOCK64();
        acc_stage += tc1 - tst;
        const int f = s_ch0[e], g = s_ch1[e];
        const int KF = s_K[f];
        const double* finF = rows + s_roff[f];
        const int16_t* mapF = s_cmap + (e * 2 + 0) * Kmax;
        const int16_t* mapG = s_cmap + (e * 2 + 1) * Kmax;
        double* cur = (n & 1) ? scr : fin;  // row i lives in fin iff (n − i) is even
        __syncthreads();  // staged lists visible

        if (kind == WHALE_WGD) {
            // q·Σ p ℓ_f[γ1]ℓ_f[γ2] + (1−q+2qϵ_f)·ℓ_f[γ]   src/core.jl:103-119,187-199
            const double cx0 = PL.cx[e * Kmax], cy0 = PL.cy[e * Kmax];
            for (int u = tid; u < C * K; u += NT) {
                const int c = u / K, k = u - c * K;
                const int kf = mapF[k];
                double s0, sk;
                term_sum<false>(s_dents, s_dptr[c], s_dptr[c + 1], 1u, finF, RS(KF), kf, finF, RS(KF), kf, s0, sk);
                const double u0 = finF[c * RS(KF)];
                double r;
                if (k == 0) r = fma(cy0, s0, cx0 * u0);
                else {
                    const double uk = kf >= 0 ? finF[c * RS(KF) + kf] : 0.0;
                    r = fma(cy0, sk, fma(cx0, uk, fma(PL.cy[e * Kmax + k], s0, PL.cx[e * Kmax + k] * u0)));
                }
                cur[c * RS(K) + k] = r;
                if (ellp && k == 0) ellp[c] = r;
            }
            __syncthreads();
            slices(cur);
            continue;
        }

        // internal node or root: speciation + loss from the children's last rows (src/core.jl:160-176)
        const int KG = s_K[g];
        const double* finG = rows + s_roff[g];
        const double* epsF = PL.eps + s_toff[f] + (size_t)s_nsl[f] * KF;
        const double* epsG = PL.eps + s_toff[g] + (size_t)s_nsl[g] * KG;
        const double ef0 = epsF[0], eg0 = epsG[0];
        if (kind == WHALE_INTERNAL) {
            for (int u = tid; u < C * K; u += NT) {
                const int c = u / K, k = u - c * K;
                const int kf = mapF[k], kg = mapG[k];
                const double efk = (k > 0 && kf >= 0) ? epsF[kf] : 0.0;
                const double egk = (k > 0 && kg >= 0) ? epsG[kg] : 0.0;
                double s0, sk, l0, lk;
                term_sum<false>(s_tents, s_tptr[c], s_tptr[c + 1], 1u, finF, RS(KF), kf, finG, RS(KG), kg, s0, sk);
                loss_term(s_lossF[c], s_lossG[c], finF, RS(KF), kf, finG, RS(KG), kg, ef0, efk, eg0, egk, k == 0 ? 0.0 : 1.0, l0, lk);
                const double r = k == 0 ? s0 + l0 : sk + lk;
                cur[c * RS(K) + k] = r;
                if (ellp && k == 0) ellp[c] = r;
            }
            __syncthreads();
            slices(cur);
            continue;
        }

        // ---- root: ℓ_r[γ] = (1−η)ξ/η·a + η(1−ϵ)/ξ²·(b + c),  a = Σ p ℓ_r[γ1]ℓ_r[γ2] over the SAME row
        //      (src/core.jl:130-158) => clades in ascending size, one level (= clade size) at a time.
        //      A level with few cells gives each (cell, component) a team of G adjacent lanes; every lane forms the
        //      combined partial result of its share of the Πroot and speciation terms, so the team reduces ONE
        //      double by shuffles.
        //      The level's Πroot and speciation terms are staged in shared memory one level ahead (two buffers), so the
        //      per-level critical path holds no global-memory latency.
        const double cx0 = PL.cx[e * Kmax], cy0 = PL.cy[e * Kmax];
        const Ent* g_dents = ents + R.dent_off;
        const Ent* g_tents = ents + R.tent_off;
        uint4* const lbuf = st5 + dp16 + tp16 + tn16 + pp16;
        const uint32_t rootwin = Hp->rootwin;
        auto issue_level = [&](uint32_t L) {
            const int c0 = (int)s_lev[L], c1 = (int)s_lev[L + 1];
            const uint32_t da = s_dptr[c0], na = s_dptr[c1] - da, ta = s_tptr[c0], nb = s_tptr[c1] - ta;
            uint4* dst = lbuf + (L & 1u) * rootwin;
            copy16(dst, reinterpret_cast<const uint4*>(g_dents + da), (int)na, tid, NT);
            copy16(dst + na, reinterpret_cast<const uint4*>(g_tents + ta), (int)nb, tid, NT);
        };
        if (staged) { issue_level(0); stage_commit(); }
        for (uint32_t L = 0; L < nlev; L++) {
            if (staged) {
                if (L + 1 < nlev) issue_level(L + 1);
                stage_commit();
                stage_wait_prev();
                __syncthreads();  // level L's terms are in shared memory for every thread
            }
            const int c0 = (int)s_lev[L], c1 = (int)s_lev[L + 1];
            const uint32_t da = s_dptr[c0], ta = s_tptr[c0];
            const Ent* l_dents = staged ? reinterpret_cast<const Ent*>(lbuf + (L & 1u) * rootwin) : g_dents + da;
            const Ent* l_tents = staged ? l_dents + (s_dptr[c1] - da) : g_tents + ta;
            const int groups = (c1 - c0) * K;
            int glog = 0;  // team size from the number of cells only: the value path must not depend on the plan's K
            while (glog < 4 && ((c1 - c0) << (glog + 1)) <= 16) glog++;
            const int G = 1 << glog;
            const int lanes = groups << glog;
            for (int ub = 0; ub < lanes; ub += NT) {  // uniform trip count: the shuffles below need whole warps
                const int u = ub + tid;
                const int gi = u >> glog, j = u & (G - 1);
                const bool valid = gi < groups;
                double v = 0.0;
                int c = 0, k = 0, kf = 0, kg = 0;
                if (valid) {
                    const int cc = gi / K;
                    c = c0 + cc; k = gi - cc * K;
                    kf = mapF[k]; kg = mapG[k];
                    double a0, ak, b0, bk;
                    term_sum<false>(l_dents, s_dptr[c] - da + j, s_dptr[c + 1] - da, (uint32_t)G, fin, RS(K), k, fin, RS(K), k, a0, ak);
                    term_sum<false>(l_tents, s_tptr[c] - ta + j, s_tptr[c + 1] - ta, (uint32_t)G, finF, RS(KF), kf, finG, RS(KG), kg, b0, bk);
                    if (k == 0) v = fma(cx0, a0, cy0 * b0);
                    else v = fma(cx0, ak, fma(cy0, bk, fma(PL.cx[e * Kmax + k], a0, PL.cy[e * Kmax + k] * b0)));
                }
                for (int step = 1; step < G; step <<= 1) v += SHFL_DOWN(v, step);
                if (valid && j == 0) {
                    const double efk = (k > 0 && kf >= 0) ? epsF[kf] : 0.0;
                    const double egk = (k > 0 && kg >= 0) ? epsG[kg] : 0.0;
                    double l0, lk;
                    loss_term(s_lossF[c], s_lossG[c], finF, RS(KF), kf, finG, RS(KG), kg, ef0, efk, eg0, egk, k == 0 ? 0.0 : 1.0, l0, lk);
                    if (k == 0) v = fma(cy0, l0, v);
                    else v = fma(cy0, lk, fma(PL.cy[e * Kmax + k], l0, v));
                    fin[c * RS(K) + k] = v;
                    if (ellp && k == 0) ellp[c] = v;
                }
            }
            __syncthreads();
        }
        if (tid < K) {  // log L and its gradient (src/core.jl:35-36)
            const double Lv = fin[(C - 1) * RS(K)];
            double o;
            if (Lv > 0.0) o = tid == 0 ? log(Lv) : fin[(C - 1) * RS(K) + tid] / Lv;
            else o = tid == 0 ? -dinf() : 0.0;
            A.out_fam[(size_t)fam * K + tid] = o;
        }
        if (A.tim && tid == 0) {
            const long long te = CLOCK64();
            long long* T = A.tim + (size_t)fam * TIMW;
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
    if (A.done) {
        const TailArgs TA{A.done, A.n_total, M.root, PL.K, PL.Kmax, A.out_fam, PL.cond, A.cond_kind, A.first, A.out, PL.act, A.peer};
        dp_tail_reduce<NT>(TA, 1, reinterpret_cast<double*>(smem_raw));
    }
}
