// K1 — slice tables (ϵ, ϕ, ψ) with forward tangents.  src/model.jl:162-191, src/bdputil.jl:6-11.
//
// The reference iterates ϵ_i = (α + (1−α−β)ϵ_{i−1}) / (1 − βϵ_{i−1}) slice by slice.  All slices of a branch share
// (α, β) (uniform Δt, src/model.jl:20), so the step is ONE Möbius map M = [[1−α−β, α], [−β, 1]] and row i is M^i ϵ_0.
//   * closed form (the normal case).  For the linear birth–death process the i-fold composition is the same map
//     with (α, β) evaluated at time i·Δt:  ϵ_i = (α_i + (1−α_i−β_i)ϵ_0) / (1 − β_i ϵ_0),  α_i = getα(λ, μ, iΔt).
//     Every row of every branch is then independent: the only serial part left is the chain over the species
//     tree's height (ϵ_0 of a branch is the product of its children's last ϵ), a handful of flops per level.
//     Row i is formed from ϵ_{i−1} exactly as the reference does (ϕ_i, ψ_i, ϵ_i from 1 − βϵ_{i−1}).
//   * chain (fallback).  Near the critical case the reference's own arithmetic is what has to be mirrored: its
//     getα switches to the λ = μ formula for |λ−μ| ≤ 1e-6 (an approximation whose composition is NOT the
//     time-i·Δt map) and loses digits to cancellation in exp(Δt(λ−μ)) − 1 just above it.  Branches with
//     |Δt(λ−μ)| < 1e-4 therefore run the per-slice recurrence, in projective form ϵ = u/v (division-free:
//     u_i = (1−α−β)u_{i−1} + αv_{i−1},  v_i = v_{i−1} − βu_{i−1},  1 − βϵ_{i−1} = v_i/v_{i−1}).
// Launch: G table CTAs + one CTA per leaf node (tree-shape rows, below).  Every table CTA runs the cheap
// per-level chain redundantly (identical values) and takes a 1/G share of the rows; no inter-CTA dependency.
#pragma once
#include "whale_common.cuh"

#ifdef WHALE_EMU
#define SHFL_IDX(v, src) emu::shfl_idx(v, src)
#else
#define SHFL_IDX(v, src) __shfl_sync(0xffffffffu, v, src)
#endif

constexpr unsigned TAB_FORCE_CHAIN = 1u;  // flags: run the per-slice recurrence on every branch (tests, A/B)
constexpr unsigned TAB_JACOBIAN = 2u;     // this plan: no slice rows, the reverse pass's Jacobian table instead (see below)

// species-tree metadata staged in shared memory: after an L2 flush every dependent global load of these tiny
// arrays costs a DRAM round trip per level
struct TabMeta {
    const int *kind, *nsl, *ch0, *ch1, *ls, *ms, *qs, *K, *toff, *lvl_off, *lvl_nodes, *koff;
    int* mode;  // 1: closed form, 0: chain
    const double *dt, *leafP, *pleaf, *x;
    const int16_t* cmap;
    const uint8_t* role;
};

// getα, β of the linear BDP over time t, non-critical branch (src/bdputil.jl:6-7; β = (λ/μ)α src/model.jl:186)
__device__ __forceinline__ void bdp_ab_t(D1 lam, D1 mu, double t, D1& a, D1& b) {
    const D1 ex = dexp(mk(t) * (lam - mu));
    a = mu * (ex - 1.0) / (lam * ex - mu);
    b = (lam / mu) * a;
}
// one (α, β) map applied to ϵ
__device__ __forceinline__ D1 eps_map(D1 a, D1 b, D1 e0) { return (a + ((1.0 - a) - b) * e0) / (1.0 - b * e0); }

__device__ __forceinline__ void rates_of(const TabMeta& T, const ModelDev& M, int e, unsigned role, D1& lam, D1& mu) {
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    const int ls = T.ls[e], ms = T.ms[e];
    const double lv = ls < 0 ? NaN : (M.log_scale ? exp(T.x[ls]) : T.x[ls]);
    const double mv = ms < 0 ? NaN : (M.log_scale ? exp(T.x[ms]) : T.x[ms]);
    lam = mk(lv, (role & 1u) ? (M.log_scale ? lv : 1.0) : 0.0);
    mu = mk(mv, (role & 2u) ? (M.log_scale ? mv : 1.0) : 0.0);
}

__device__ void leafshapes_block(const ModelDev& M, const PlanDev& PL, const double* __restrict__ x,
                                 const double* __restrict__ pleaf, int li, unsigned char* tsm);

// one CTA of the table launch for plan PL: CTA `bid` of G table CTAs, or (bid >= G) the leaf-shape CTA of leaf bid − G
__device__ __forceinline__ void tables_block(const ModelDev& M, const PlanDev& PL, const double* __restrict__ x,
                                             const double* __restrict__ pleaf, int G, unsigned flags, int bid,
                                             unsigned char* tsm) {
    GRID_DEP_LAUNCH();  // the DP kernel behind this one may start its prologue now (it waits before reading the tables)
    if (bid >= G) {  // tree-shape rows of one leaf branch
        leafshapes_block(M, PL, x, pleaf, bid - G, tsm);
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const bool writer = bid == 0;  // the per-branch outputs are identical in every table CTA: one writes
    const long long tk0 = CLOCK64();
    const int nn = M.nn, Kmax = PL.Kmax, P = M.n_params;
    // dt, leafP, pleaf [nn each], x [P]; then per (node, component), packed with the plan's own offsets koff[e] (KT = Σ_e K_e
    // entries, not nn·Kmax: the full plan of a branch-wise model has Kmax ≈ P but few components on most nodes):
    // (α, β) [2·KT], whole-branch (α_n, β_n) [2·KT], ϵ_0 [KT], ϵ_n [KT]
    const int KT = PL.ktot;
    double* sd = reinterpret_cast<double*>(tsm);
    double* s_ab = sd + 3 * nn + P;           // per-slice (α, β) components
    double* s_abn = s_ab + 2 * KT;            // whole-branch (α_n, β_n) components (closed-form branches)
    double* s_e0 = s_abn + 2 * KT;
    double* s_en = s_e0 + KT;
    int* si = reinterpret_cast<int*>(s_en + KT);
    const double* const s_eps0 = s_e0;
    // 10 arrays [nn], lvl_off [nlvl+1], lvl_nodes [nn]
    int16_t* s_cm = reinterpret_cast<int16_t*>(si + 12 * nn + M.nlvl + 1);
    uint8_t* s_ro = reinterpret_cast<uint8_t*>(s_cm + nn * 2 * Kmax);
    {
        // one round trip: the first blockDim elements of every array are requested before anything is stored
        // (after an L2 flush each dependent round costs a DRAM latency); longer arrays finish in the loops below
        const int t = threadIdx.x, nt = blockDim.x;
        double r_dt = 0.0, r_lp = 0.0, r_pl = 0.0, r_x = 0.0;
        int r_kind = 0, r_nsl = 0, r_c0 = 0, r_c1 = 0, r_ls = 0, r_ms = 0, r_qs = 0, r_K = 0, r_to = 0, r_ln = 0, r_lo = 0, r_ko = 0;
        int16_t r_cm = 0;
        uint8_t r_ro = 0;
        if (t < nn) {
            r_dt = M.dt[t]; r_lp = M.leafP[t]; r_pl = pleaf ? pleaf[t] : 0.0;
            r_kind = M.kind[t]; r_nsl = M.nsl[t]; r_c0 = M.child0[t]; r_c1 = M.child1[t];
            r_ls = M.lam_slot[t]; r_ms = M.mu_slot[t]; r_qs = M.q_slot[t];
            r_K = PL.K[t]; r_to = PL.toff[t]; r_ln = M.lvl_nodes[t]; r_ko = PL.koff[t];
        }
        if (t <= M.nlvl) r_lo = M.lvl_off[t];
        if (t < P) r_x = x[t];
        if (t < nn * 2 * Kmax) r_cm = PL.cmap[t];
        if (t < nn * Kmax) r_ro = PL.role[t];
        if (t < nn) {
            sd[t] = r_dt; sd[nn + t] = r_lp; sd[2 * nn + t] = r_pl;
            si[t] = r_kind; si[nn + t] = r_nsl; si[2 * nn + t] = r_c0; si[3 * nn + t] = r_c1;
            si[4 * nn + t] = r_ls; si[5 * nn + t] = r_ms; si[6 * nn + t] = r_qs;
            si[7 * nn + t] = r_K; si[8 * nn + t] = r_to; si[10 * nn + M.nlvl + 1 + t] = r_ln; si[11 * nn + M.nlvl + 1 + t] = r_ko;
        }
        if (t <= M.nlvl) si[10 * nn + t] = r_lo;
        if (t < P) sd[3 * nn + t] = r_x;
        if (t < nn * 2 * Kmax) s_cm[t] = r_cm;
        if (t < nn * Kmax) s_ro[t] = r_ro;
        for (int i = t + nt; i < nn; i += nt) {
            sd[i] = M.dt[i]; sd[nn + i] = M.leafP[i]; sd[2 * nn + i] = pleaf ? pleaf[i] : 0.0;
            si[i] = M.kind[i]; si[nn + i] = M.nsl[i]; si[2 * nn + i] = M.child0[i]; si[3 * nn + i] = M.child1[i];
            si[4 * nn + i] = M.lam_slot[i]; si[5 * nn + i] = M.mu_slot[i]; si[6 * nn + i] = M.q_slot[i];
            si[7 * nn + i] = PL.K[i]; si[8 * nn + i] = PL.toff[i]; si[10 * nn + M.nlvl + 1 + i] = M.lvl_nodes[i];
            si[11 * nn + M.nlvl + 1 + i] = PL.koff[i];
        }
        for (int i = t + nt; i <= M.nlvl; i += nt) si[10 * nn + i] = M.lvl_off[i];
        for (int i = t + nt; i < P; i += nt) sd[3 * nn + i] = x[i];
        for (int i = t + nt; i < nn * 2 * Kmax; i += nt) s_cm[i] = PL.cmap[i];
        for (int i = t + nt; i < nn * Kmax; i += nt) s_ro[i] = PL.role[i];
    }
    __syncthreads();
    TabMeta T{si, si + nn, si + 2 * nn, si + 3 * nn, si + 4 * nn, si + 5 * nn, si + 6 * nn, si + 7 * nn, si + 8 * nn,
              si + 10 * nn, si + 10 * nn + M.nlvl + 1, si + 11 * nn + M.nlvl + 1, si + 9 * nn, sd, sd + nn, sd + 2 * nn, sd + 3 * nn, s_cm, s_ro};
    // ---- phase A0: per-slice (α, β) of every branch and its mode — they depend on the branch's own rates only,
    //      so all nodes go in parallel (one warp per node) ----
    for (int e = warp; e < nn; e += nwarp) {
        const int K = T.K[e];
        for (int k = lane; k < K; k += 32) {
            const unsigned role = k == 0 ? 0u : T.role[e * PL.Kmax + k];
            D1 lam, mu;
            rates_of(T, M, e, role, lam, mu);
            D1 a = mk(0.0), b = mk(0.0);
            int closed = 0;
            if (T.nsl[e] > 0) {
                // getα src/bdputil.jl:6-7 (critical branch decided on VALUES, like isapprox on Duals)
                const double t = T.dt[e];
                if (fabs(lam.v - mu.v) <= 1e-6) {
                    a = (lam * mk(t)) / (1.0 + lam * mk(t));
                } else {
                    const D1 ex = dexp(mk(t) * (lam - mu));
                    a = mu * (ex - 1.0) / (lam * ex - mu);
                    closed = (fabs(t * (lam.v - mu.v)) >= 1e-4 && !(flags & TAB_FORCE_CHAIN)) ? 1 : 0;
                }
                b = (lam / mu) * a;
                if (closed) {  // the whole branch as one map (time n·Δt): all the chain over levels needs from this branch
                    D1 an, bn;
                    bdp_ab_t(lam, mu, t * (double)T.nsl[e], an, bn);
                    s_abn[(T.koff[e] + k) * 2 + 0] = k == 0 ? an.v : an.d;
                    s_abn[(T.koff[e] + k) * 2 + 1] = k == 0 ? bn.v : bn.d;
                }
            }
            s_ab[(T.koff[e] + k) * 2 + 0] = k == 0 ? a.v : a.d;
            s_ab[(T.koff[e] + k) * 2 + 1] = k == 0 ? b.v : b.d;
            if (k == 0) T.mode[e] = closed;
            if (writer) {
                PL.ab[(e * PL.Kmax + k) * 2 + 0] = k == 0 ? a.v : a.d;
                PL.ab[(e * PL.Kmax + k) * 2 + 1] = k == 0 ? b.v : b.d;
            }
        }
    }
    __syncthreads();
    if (writer && threadIdx.x == 0) PL.tim[0] = CLOCK64() - tk0;
    // ---- phase A: the chain over the tree's height.  Per level, one warp per node, lanes over components:
    //      ϵ_0 from the children's last ϵ (shared memory), then the branch's last ϵ ----
    for (int L = 0; L < M.nlvl; L++) {
        const int n0 = T.lvl_off[L], n1 = T.lvl_off[L + 1];
        for (int j = n0 + warp; j < n1; j += nwarp) {
            const int e = T.lvl_nodes[j];
            const int K = T.K[e], kind = T.kind[e], n = T.nsl[e];
            const int closed = T.mode[e];
            for (int k = lane; k < K; k += 32) {
                const unsigned role = k == 0 ? 0u : T.role[e * PL.Kmax + k];
                auto child_last = [&](int jc, int child) -> D1 {  // the child's last ϵ, this lane's component
                    const int kc = k == 0 ? 0 : T.cmap[(e * 2 + jc) * Kmax + k];
                    return mk(s_en[T.koff[child]], (k > 0 && kc >= 0) ? s_en[T.koff[child] + kc] : 0.0);
                };
                D1 ep;
                if (kind == WHALE_LEAF) {  // setnode! src/model.jl:170
                    ep = mk(T.pleaf[e]);
                } else if (kind == WHALE_WGD) {  // setwgdnode! src/model.jl:175-180
                    const D1 q = mk(T.x[T.qs[e]], (role & 4u) ? 1.0 : 0.0);
                    const D1 ec = child_last(0, T.ch0[e]);
                    ep = q * (ec * ec) + (1.0 - q) * ec;
                    const D1 w = (1.0 - q) + 2.0 * (q * ec);  // Πwgdloss coefficient src/core.jl:198
                    if (writer) {
                        PL.cx[e * PL.Kmax + k] = k == 0 ? w.v : w.d;
                        PL.cy[e * PL.Kmax + k] = k == 0 ? q.v : q.d;
                    }
                } else {  // internal / root: product of the children's last ϵ
                    const D1 ef = child_last(0, T.ch0[e]);
                    const D1 eg = child_last(1, T.ch1[e]);
                    ep = ef * eg;
                    if (kind == WHALE_ROOT && writer) {  // whaleroot! src/core.jl:131-147 ; condition src/condition.jl
                        const D1 eta = mk(T.x[M.eta_slot], (role & 8u) ? 1.0 : 0.0);
                        const D1 xi = 1.0 - (1.0 - eta) * ep;
                        const D1 A = (1.0 - eta) * xi / eta;
                        const D1 B = eta * (1.0 - ep) / (xi * xi);
                        PL.cx[e * PL.Kmax + k] = k == 0 ? A.v : A.d;
                        PL.cy[e * PL.Kmax + k] = k == 0 ? B.v : B.d;
                        // geompgf(η, s) = ηs/(1−(1−η)s)  src/bdputil.jl:67
                        const D1 gr = eta * ep / (1.0 - (1.0 - eta) * ep);
                        const D1 gf = eta * ef / (1.0 - (1.0 - eta) * ef);
                        const D1 gg = eta * eg / (1.0 - (1.0 - eta) * eg);
                        const D1 pr = ((1.0 - gf) - gg) + gr;  // RootCondition :21-29
                        const D1 pn = 1.0 - gr;                // NonExtinctCondition :15-18
                        const D1 cr = pr.v > 0.0 ? dlog(pr) : mk(-dinf(), 0.0);
                        const D1 cn = dlog(pn);
                        PL.cond[0 * PL.Kmax + k] = 0.0;
                        PL.cond[1 * PL.Kmax + k] = k == 0 ? cr.v : cr.d;
                        PL.cond[2 * PL.Kmax + k] = k == 0 ? cn.v : cn.d;
                    }
                }
                if (role & 16u) ep = mk(ep.v, 1.0);  // local plan: this component is ∂/∂ϵ_0 of the branch itself
                s_e0[T.koff[e] + k] = k == 0 ? ep.v : ep.d;
                D1 en = ep, lf = mk(0.0);
                if (n > 0 && closed) {
                    const D1 an = mk(s_abn[(T.koff[e]) * 2], k == 0 ? 0.0 : s_abn[(T.koff[e] + k) * 2]);
                    const D1 bn = mk(s_abn[(T.koff[e]) * 2 + 1], k == 0 ? 0.0 : s_abn[(T.koff[e] + k) * 2 + 1]);
                    const D1 r = mk(1.0) / (1.0 - bn * ep);
                    en = (an + ((1.0 - an) - bn) * ep) * r;
                    // leaf clade on a leaf branch: ℓ_n = leafℙ·Π_i ϕ_i (src/core.jl:94,123) = leafℙ·ϕ over the whole
                    // branch: Π_i ϕ_i = (1−α_n)(1−β_n)/(1−β_n ϵ_0)²  (det M^n = ((1−α)(1−β))^n)
                    if (kind == WHALE_LEAF) lf = mk(T.leafP[e]) * (((1.0 - an) * (1.0 - bn)) * (r * r));
                } else if (n > 0) {
                    double2* uvrow = PL.uv + T.toff[e];  // every table CTA writes the same values and reads its own
                    D1 u = ep, v = mk(1.0);
                    uvrow[k] = k == 0 ? make_double2(u.v, v.v) : make_double2(u.d, v.d);
                    const D1 a = mk(s_ab[(T.koff[e]) * 2], k == 0 ? 0.0 : s_ab[(T.koff[e] + k) * 2]);
                    const D1 b = mk(s_ab[(T.koff[e]) * 2 + 1], k == 0 ? 0.0 : s_ab[(T.koff[e] + k) * 2 + 1]);
                    const D1 c = (1.0 - a) - b;
                    for (int i = 1; i <= n; i++) {
                        const D1 un = c * u + a * v;
                        const D1 vn = v - b * u;
                        u = un;
                        v = vn;
                        uvrow[(size_t)i * K + k] = k == 0 ? make_double2(u.v, v.v) : make_double2(u.d, v.d);
                    }
                    en = u / v;
                    if (kind == WHALE_LEAF) {  // ℓ_n = leafℙ·gⁿ·(v_0/v_n)²
                        const D1 g = (1.0 - a) * (1.0 - b);
                        const D1 r = mk(1.0) / v;
                        lf = mk(T.leafP[e]) * dpowi(g, n) * (r * r);
                    }
                } else if (kind == WHALE_LEAF) {
                    lf = mk(T.leafP[e]);
                }
                s_en[T.koff[e] + k] = k == 0 ? en.v : en.d;
                if (kind == WHALE_LEAF && writer) PL.leaf[e * PL.Kmax + k] = k == 0 ? lf.v : lf.d;
            }
        }
        __syncthreads();
        if (writer && threadIdx.x == 0 && L < 28) PL.tim[1 + L] = CLOCK64() - tk0;
    }
    if (flags & TAB_JACOBIAN) {
        // plan G of a reverse-mode evaluation (one table CTA): the slice rows are not needed, only the Jacobian of the
        // quantities whose adjoints k_dp_rev accumulates per node — row r = e*8 + j (whale_rev.cuh: zloc), column k =
        // component of the root's list:  j = 0, 1 own λ, μ (0/1; leaf branches: their components 1, 2);  2 ϵ_0 of the branch;
        // 3, 4 the row-1 coefficients cx, cy (WGD, root);  5, 6 the last ϵ of child 0 / child 1 (Πloss)
        const int KRr = T.K[M.root];
        for (int idx = threadIdx.x; idx < nn * 8 * KRr; idx += blockDim.x) {
            const int k = idx % KRr, r = idx / KRr, j = r & 7, e = r >> 3;
            double v = 0.0;
            if (k > 0) {
                const int gp = PL.act[M.root * PL.Kmax + k];
                const int kind = T.kind[e];
                const int ke = PL.rinv[e * KRr + k];
                if (kind == WHALE_LEAF) {
                    if (j < 2 && j + 1 < T.K[e] && PL.act[e * PL.Kmax + j + 1] == gp) v = 1.0;
                } else {
                    if (kind != WHALE_ROOT) {
                        if (j == 0 && T.ls[e] == gp) v = 1.0;
                        if (j == 1 && T.ms[e] == gp) v = 1.0;
                        if (j == 2 && ke > 0) v = s_e0[T.koff[e] + ke];
                    }
                    if (kind != WHALE_INTERNAL && ke > 0) {
                        if (j == 3) v = PL.cx[e * PL.Kmax + ke];
                        if (j == 4) v = PL.cy[e * PL.Kmax + ke];
                    }
                    if (kind != WHALE_WGD && (j == 5 || j == 6)) {
                        const int c = j == 5 ? T.ch0[e] : T.ch1[e];
                        const int kc = PL.rinv[c * KRr + k];
                        if (kc > 0) v = s_en[T.koff[c] + kc];
                    }
                }
            }
            PL.jac[idx] = v;
        }
        if (threadIdx.x == 0) { PL.tim[30] = CLOCK64() - tk0; PL.tim[31] = M.nlvl; }
        return;
    }
    // ---- phase B: every (node, row, component) of the tables in parallel (one flat index space, 1/G per CTA) ----
    const int total = T.toff[nn - 1] + (T.nsl[nn - 1] + 1) * T.K[nn - 1];  // toff is ascending in node index
    for (int t = bid * blockDim.x + threadIdx.x; t < total; t += G * blockDim.x) {
        int lo = 0, hi = nn - 1;  // node e with toff[e] <= t < toff[e+1]
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (T.toff[mid] <= t) lo = mid; else hi = mid - 1;
        }
        const int e = lo, K = T.K[e];
        const int idx = t - T.toff[e];
        const int i = idx / K, k = idx - i * K;
        if (i == 0) {
            PL.eps[t] = s_eps0[T.koff[e] + k];
            PL.pp[t] = make_double2(k == 0 ? 1.0 : 0.0, k == 0 ? 1.0 : 0.0);  // ϕ_1 = 1 (src/model.jl:171)
            continue;
        }
        const D1 a = mk(s_ab[(T.koff[e]) * 2], k == 0 ? 0.0 : s_ab[(T.koff[e] + k) * 2]);
        const D1 b = mk(s_ab[(T.koff[e]) * 2 + 1], k == 0 ? 0.0 : s_ab[(T.koff[e] + k) * 2 + 1]);
        const D1 g = (1.0 - a) * (1.0 - b);
        D1 ep, r;  // ϵ_i and 1 / (1 − βϵ_{i−1})
        if (T.mode[e]) {
            D1 prev = mk(s_eps0[T.koff[e]], k == 0 ? 0.0 : s_eps0[T.koff[e] + k]);
            if (i > 1) {
                const unsigned role = k == 0 ? 0u : T.role[e * PL.Kmax + k];
                D1 lam, mu, ai, bi;
                rates_of(T, M, e, role, lam, mu);
                bdp_ab_t(lam, mu, T.dt[e] * (double)(i - 1), ai, bi);
                prev = eps_map(ai, bi, prev);
            }
            r = mk(1.0) / (1.0 - b * prev);
            ep = (a + ((1.0 - a) - b) * prev) * r;
        } else {
            const double2* uvrow = PL.uv + T.toff[e];
            const double2 w0 = uvrow[(size_t)i * K], wk = uvrow[(size_t)i * K + k];
            const D1 u = mk(w0.x, k == 0 ? 0.0 : wk.x), v = mk(w0.y, k == 0 ? 0.0 : wk.y);
            const double2 p0 = uvrow[(size_t)(i - 1) * K], pk = uvrow[(size_t)(i - 1) * K + k];
            const D1 vp = mk(p0.y, k == 0 ? 0.0 : pk.y);
            ep = u / v;
            r = vp / v;
        }
        PL.eps[t] = k == 0 ? ep.v : ep.d;
        const D1 phi = g * (r * r);
        const D1 psi = (g * b) * (r * r * r);
        PL.pp[t] = make_double2(k == 0 ? phi.v : phi.d, k == 0 ? psi.v : psi.d);
    }
    __syncthreads();
    if (writer && threadIdx.x == 0) { PL.tim[30] = CLOCK64() - tk0; PL.tim[31] = M.nlvl; }
}

__global__ void __launch_bounds__(TABLES_NT) k_tables(ModelDev M, PlanDev PL, const double* __restrict__ x,
                                                      const double* __restrict__ pleaf, int G, unsigned flags) {
    EXTERN_SHARED(tsm);
    tables_block(M, PL, x, pleaf, G, flags, (int)blockIdx.x, tsm);
}

// The reverse-mode evaluation needs three table sets for the same θ — the hybrid plan the DP runs on (leaf branches
// with their own λ, μ tangents, everything else value only; with the leaf-shape CTAs), the full plan (global tangents
// of ϵ and of the row-1 coefficients, contracted with the adjoints at the end) and the local plan (∂ϕ_i, ∂ψ_i w.r.t.
// the branch's own λ, μ and ϵ_0) — in ONE launch: CTAs [0, n0) serve plan 0, [n0, n0+n1) plan 1, the rest plan 2.
struct Tables3 {
    PlanDev PL[3];
    int G[3];   // table CTAs per plan
    int n[3];   // CTAs per plan (G + leaf-shape CTAs)
};
__global__ void __launch_bounds__(TABLES_NT) k_tables3(ModelDev M, Tables3 T3, const double* __restrict__ x,
                                                       const double* __restrict__ pleaf, unsigned flags) {
    EXTERN_SHARED(tsm);
    int bid = (int)blockIdx.x, pi = 0;
    while (pi < 2 && bid >= T3.n[pi]) { bid -= T3.n[pi]; pi++; }
    tables_block(M, T3.PL[pi], x, pleaf, T3.G[pi], flags | (pi == 1 ? TAB_JACOBIAN : 0u), bid, tsm);
}

// ---------------------------------------------------------------------------------------------------------
// Tree-shape rows — last-row values wσ_n of the tree shapes on a leaf branch (see SHAPES in whale_common.cuh).
// One CTA per leaf node beside the table CTAs of k_tables; a leaf branch depends on nothing below it, so the
// block recomputes its own branch's slice rows in shared memory: warp 0 runs the projective chain (lanes of
// shape 0), all threads turn it into (ϕ_i, ψ_i) rows, warp 0 runs the shape sequences in lock-step with
// lanes = (shape σ, component k).
// ---------------------------------------------------------------------------------------------------------
__device__ void leafshapes_block(const ModelDev& M, const PlanDev& PL, const double* __restrict__ x,
                                 const double* __restrict__ pleaf, int li, unsigned char* tsm) {
    const int e = M.leafnodes[li];
    const int tid = threadIdx.x, lane = tid & 31;
    const bool w0 = tid < 32;
    const int K = PL.K[e], n = M.nsl[e];
    const int sh = lane / K, k = lane - sh * K;
    const bool on = sh < NSHAPE;  // 8·K <= 32 lanes (leaf branches have K <= 3)
    double2* uv = reinterpret_cast<double2*>(tsm);  // [(n+1)*K] projective ϵ rows
    double2* pp = uv + (size_t)(n + 1) * K;         // [(n+1)*K] (ϕ, ψ) rows
    __shared__ double s_ab[8][4];  // (α, β) value / tangent per component, from the lanes of shape 0
    if (w0 && sh == 0) {
        const double NaN = __longlong_as_double(0x7ff8000000000000LL);
        const unsigned role = k == 0 ? 0u : PL.role[e * PL.Kmax + k];
        const int ls = M.lam_slot[e], ms = M.mu_slot[e];
        const double lv = ls < 0 ? NaN : (M.log_scale ? exp(x[ls]) : x[ls]);
        const double mv = ms < 0 ? NaN : (M.log_scale ? exp(x[ms]) : x[ms]);
        const D1 lam = mk(lv, (role & 1u) ? (M.log_scale ? lv : 1.0) : 0.0);
        const D1 mu = mk(mv, (role & 2u) ? (M.log_scale ? mv : 1.0) : 0.0);
        D1 a = mk(0.0), b = mk(0.0);
        if (n > 0) {
            const double t = M.dt[e];
            if (fabs(lam.v - mu.v) <= 1e-6) a = (lam * mk(t)) / (1.0 + lam * mk(t));
            else { const D1 ex = dexp(mk(t) * (lam - mu)); a = mu * (ex - 1.0) / (lam * ex - mu); }
            b = (lam / mu) * a;
        }
        // pass 1: the projective chain, rows to shared memory
        s_ab[k][0] = a.v; s_ab[k][1] = a.d; s_ab[k][2] = b.v; s_ab[k][3] = b.d;
        D1 u = mk(pleaf ? pleaf[e] : 0.0), v = mk(1.0);
        const D1 c = (1.0 - a) - b;
        uv[k] = k == 0 ? make_double2(u.v, v.v) : make_double2(u.d, v.d);
        for (int i = 1; i <= n; i++) {
            const D1 un = c * u + a * v, vn = v - b * u;
            u = un; v = vn;
            uv[(size_t)i * K + k] = k == 0 ? make_double2(u.v, v.v) : make_double2(u.d, v.d);
        }
    }
    __syncthreads();
    // pass 2: ϕ_i, ψ_i of every row in parallel (all threads)
    for (int idx = tid; idx < n * K; idx += blockDim.x) {
        const int i = 1 + idx / K, kk = idx - (i - 1) * K;
        const D1 aa = mk(s_ab[0][0], kk == 0 ? 0.0 : s_ab[kk][1]), bb = mk(s_ab[0][2], kk == 0 ? 0.0 : s_ab[kk][3]);
        const D1 gg = (1.0 - aa) * (1.0 - bb);
        const double2 w0_ = uv[(size_t)i * K], wk = uv[(size_t)i * K + kk];
        const double2 p0 = uv[(size_t)(i - 1) * K], pk = uv[(size_t)(i - 1) * K + kk];
        const D1 v = mk(w0_.y, kk == 0 ? 0.0 : wk.y), vp = mk(p0.y, kk == 0 ? 0.0 : pk.y);
        const D1 r = vp / v;
        const D1 phi = gg * (r * r), psi = (gg * bb) * (r * r * r);
        pp[(size_t)i * K + kk] = make_double2(kk == 0 ? phi.v : phi.d, kk == 0 ? psi.v : psi.d);
    }
    __syncthreads();
    if (!w0) return;
    // pass 3: the shape sequences, all shapes in lock-step (operands of shape σ come from the lanes of a, b)
    const int SA[NSHAPE] = SHAPE_A, SB[NSHAPE] = SHAPE_B;
    const int la_ = on ? SA[sh] * K + k : lane, lb_ = on ? SB[sh] * K + k : lane;
    D1 w = mk((on && sh == 0) ? M.leafP[e] : 0.0, 0.0);
    for (int i = 1; i <= n; i++) {
        const double2 q0 = pp[(size_t)i * K], qk = pp[(size_t)i * K + (on ? k : 0)];
        const D1 phi = mk(q0.x, k == 0 ? 0.0 : qk.x), psi = mk(q0.y, k == 0 ? 0.0 : qk.y);
        const D1 wa = mk(SHFL_IDX(w.v, la_), SHFL_IDX(w.d, la_));
        const D1 wb = mk(SHFL_IDX(w.v, lb_), SHFL_IDX(w.d, lb_));
        D1 nw = phi * w;
        if (on && sh > 0) nw = nw + psi * (wa * wb);
        w = nw;
    }
    if (on) PL.shapeW[((size_t)e * NSHAPE + sh) * PL.Kmax + k] = k == 0 ? w.v : w.d;
}

// ---------------------------------------------------------------------------------------------------------
// k_nowhere — NowhereExtinctCondition (src/condition.jl:5-9,31-36): the probability that no leaf of the species
// tree is left without a gene, by inclusion–exclusion over the tree pgf at all 2^L binary arguments
// (treepgf_allbinary, src/bdputil.jl:109-133).  Node e carries the vector f_e(x) for x ∈ {0,1}^{leaves below e};
// a parent's vector is the outer product of its children's (first child's index fastest) pushed through its own
// branch pgf — LinearBDP(λ, μ, t) (:58-64), composed with the WGD pgf (:73) at WGD nodes, Geometric(η) at the
// root.  Lanes = (argument index, component); tangents ride along as in k_tables.  One CTA; nodes in the model's
// order (children first).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ D1 bdp_pgf(D1 lam, D1 mu, double t, D1 s) {
    if (fabs(lam.v - mu.v) <= 1e-6) {  // isapprox(λ, μ, atol=ΛMATOL) on values
        const D1 lt = lam * mk(t);
        return (1.0 - (lt - 1.0) * (s - 1.0)) / (1.0 - lt * (s - 1.0));
    }
    const D1 rho = dexp((mu - lam) * mk(t));
    const D1 a = rho * (lam * s - mu);
    return (a - mu * (s - 1.0)) / (a - lam * (s - 1.0));
}

__global__ void __launch_bounds__(256) k_nowhere(ModelDev M, PlanDev PL, const double* __restrict__ x) {
    __shared__ double sh[256];
    const int Kmax = PL.Kmax;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    for (int oi = 0; oi < M.nn; oi++) {
        const int e = M.order[oi];
        const int K = PL.K[e], kind = M.kind[e], Le = PL.nwL[e];
        const long long len = 1LL << Le;
        double* ve = PL.nwvec + PL.nwoff[e];
        const int c0 = M.child0[e], c1 = M.child1[e];
        const int K0 = c0 >= 0 ? PL.K[c0] : 0, K1 = c1 >= 0 ? PL.K[c1] : 0;
        const int L0 = c0 >= 0 ? PL.nwL[c0] : 0;
        const double* v0 = c0 >= 0 ? PL.nwvec + PL.nwoff[c0] : nullptr;
        const double* v1 = c1 >= 0 ? PL.nwvec + PL.nwoff[c1] : nullptr;
        const int ls = M.lam_slot[e], ms = M.mu_slot[e];
        const double lv = ls < 0 ? NaN : (M.log_scale ? exp(x[ls]) : x[ls]);
        const double mv = ms < 0 ? NaN : (M.log_scale ? exp(x[ms]) : x[ms]);
        const double t = M.dt[e] * (double)M.nsl[e];
        for (long long idx = threadIdx.x; idx < len * K; idx += blockDim.x) {
            const long long i = idx / K;
            const int k = (int)(idx - i * K);
            const unsigned role = k == 0 ? 0u : PL.role[e * Kmax + k];
            D1 s;
            if (kind == WHALE_LEAF) {
                s = mk(i == 0 ? 0.0 : 1.0);
            } else {
                const long long i0 = c1 >= 0 ? (i & ((1LL << L0) - 1)) : i, i1 = c1 >= 0 ? (i >> L0) : 0;
                const int k0 = k == 0 ? 0 : PL.cmap[(e * 2 + 0) * Kmax + k];
                D1 a = mk(v0[i0 * K0], (k > 0 && k0 >= 0) ? v0[i0 * K0 + k0] : 0.0);
                if (c1 >= 0) {
                    const int k1 = k == 0 ? 0 : PL.cmap[(e * 2 + 1) * Kmax + k];
                    const D1 b = mk(v1[i1 * K1], (k > 0 && k1 >= 0) ? v1[i1 * K1 + k1] : 0.0);
                    a = a * b;
                }
                if (kind == WHALE_WGD) {  // wgdpgf(q, s) = s(1 − q + s q)
                    const D1 q = mk(x[M.q_slot[e]], (role & 4u) ? 1.0 : 0.0);
                    a = a * ((1.0 - q) + a * q);
                }
                s = a;
            }
            D1 r;
            if (kind == WHALE_ROOT) {
                const D1 eta = mk(x[M.eta_slot], (role & 8u) ? 1.0 : 0.0);
                r = eta * s / (1.0 - (1.0 - eta) * s);  // geompgf
            } else {
                const D1 lam = mk(lv, (role & 1u) ? (M.log_scale ? lv : 1.0) : 0.0);
                const D1 mu = mk(mv, (role & 2u) ? (M.log_scale ? mv : 1.0) : 0.0);
                r = bdp_pgf(lam, mu, t, s);
            }
            ve[idx] = k == 0 ? r.v : r.d;
        }
        __syncthreads();
    }
    // c = 1 − Σ_{i < 2^L − 1} (−1)^{popcount(i)} p_i ; condition = log c for 0 < c < 1, else −Inf
    const int root = M.root, KR = PL.K[root];
    const long long len = 1LL << PL.nwL[root];
    const double* p = PL.nwvec + PL.nwoff[root];
    __shared__ double c0s;
    for (int k = 0; k < KR; k++) {
        double acc = 0.0;
        for (long long i = threadIdx.x; i < len - 1; i += blockDim.x) {
            const double v = p[i * KR + k];
            acc += (__popcll((unsigned long long)i) & 1) ? -v : v;
        }
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (k == 0) {
                c0s = 1.0 - sh[0];
                PL.cond[3 * Kmax] = (c0s > 0.0 && c0s < 1.0) ? log(c0s) : -dinf();
            } else {
                PL.cond[3 * Kmax + k] = (c0s > 0.0 && c0s < 1.0) ? -sh[0] / c0s : 0.0;
            }
        }
        __syncthreads();
    }
}
