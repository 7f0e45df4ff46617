// libwhalecuda — B200 (sm_100a) engine for Whale.jl's ALE/DLWGD likelihood + forward-mode gradient.
//
// Hot path replaced (reference paths relative to the reference checkout):
//   slice tables   src/model.jl:162-191, src/bdputil.jl:6-11          -> k_tables   (whale_tables.cuh; closed-form rows)
//   the DP         src/core.jl:83-199 (whale!/whalewgd!/whaleroot!)   -> k_dp       (whale_dp.cuh)
//   Σ − N·cond     src/core.jl:46-64, src/condition.jl:11-29          -> tail of k_dp (dp_tail_reduce; k_reduce1/2 in
//                                                                        whale_reduce.cuh for very large batches)
// Forward tangents replace ForwardDiff duals: every ℓ cell carries K_e = 1 + (#parameters that can
// influence branch e) components; lanes span (clade cell × component).
//
// Data layout in HBM (built once by whale_data_create, the read_ale-time packer):
//   FamHdr[F]   : per family {arena byte offset, Γ, #root levels, ℓ offset, shared-memory budget}
//   arena blob  : per family  NodeRec[n_nodes] | per-node u32 CSR pointers | 16-byte term entries
//                 {u16 i1, u16 i2, f64 p} with indices already resolved to the *local* cell index of the
//                 branch they are used in (the reference resolves index[γ,e] at run time, src/ccd.jl:41-49)
// On chip: one CTA per family; the last row of every branch (C_e × K_e doubles) lives in shared memory,
// slices ping-pong between that row and a scratch row.  Families are binned by shared-memory need and the
// bins launched concurrently so ragged inputs do not drag occupancy down to the largest family.
// No tensor cores: the DP is an irregular gather–multiply–accumulate in fp64.
//
// There is NO CPU fallback in this file: every entry point that computes needs a CUDA device.
#include "whale_common.cuh"
#include "whale_tables.cuh"
#include "whale_dp.cuh"
#include "whale_rev.cuh"
#include "whale_reduce.cuh"
#include "whale_track.cuh"
#include "whale_ale.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

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
    std::vector<int> lvl_off, lvl_nodes, leafnodes, inner;
    ModelDev dev{};
    std::vector<void*> owned;
    Plan plan[2];  // 0: value only, 1: all raw parameters
    // reverse-mode gradient (whale_rev.cuh): the hybrid plan the DP runs on (leaf branches: own λ, μ; else value only),
    // the local plan (K = 4 on internal/WGD branches: ∂/∂ own λ, own μ, ϵ_0) and, for the full plan, the component of
    // every root component in every node's list
    Plan planR, planL;
    std::vector<int16_t> rinv;
    int16_t* d_rinv = nullptr;
    int n_sm = 1;
    bool attr_set = false;  // the kernels' dynamic shared-memory opt-in was done on this model's device
    bool x_host_valid = false;  // d_x / d_pleaf hold the parameters of the last host-pointer evaluation
    double* d_x = nullptr;      // staging for host-pointer calls
    double* d_pleaf = nullptr;  // [nn]
    double* d_out = nullptr;    // [1+P]
    double* h_pin = nullptr;    // pinned staging (x in, out back)
    cudaStream_t stream = nullptr;
};

// k_dp launch shape: threads per family CTA and the resident-CTA target that caps registers.
// Defaults chosen from B200 measurements; WHALE_NT / WHALE_MINB override for experiments.
static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}
static int dp_nt() {
    static int nt = [] { int v = env_int("WHALE_NT", 128); return (v == 64 || v == 256) ? v : 128; }();
    return nt;
}
static int dp_minb() {  // resident CTAs per SM the register cap is chosen for
    static int mb = [] {
        const int nt = dp_nt(), v = env_int("WHALE_MINB", 0);
        if (nt == 64) return 8;
        if (nt == 256) return 2;
        return (v == 3 || v == 5 || v == 6) ? v : 4;
    }();
    return mb;
}
// launch shapes compiled in (k_dp and k_dp_rev); round 1's sweep over 13 shapes found nothing better than 128 x 4
// (profiles/r1_occupancy_sweep_v8.txt), the others stay for experiments
#ifdef WHALE_DEV_BUILD  // quick experiment builds: the default variant only
#define DP_VARIANTS(X) X(128, 4)
#else
#define DP_VARIANTS(X) X(64, 8) X(128, 3) X(128, 4) X(128, 5) X(128, 6) X(256, 2)
#endif
// k_dp_rev: WHALE_REV_NT / WHALE_REV_MINB select among the compiled shapes (threads per family CTA, resident CTAs per SM)
#ifdef WHALE_DEV_BUILD
#define REV_VARIANTS(X) X(128, 4) X(128, 5) X(128, 6) X(128, 8)
#else
#define REV_VARIANTS(X) X(32, 16) X(64, 8) X(128, 3) X(128, 4) X(128, 5) X(128, 6) X(128, 8) X(256, 2)
#endif
static int rev_nt() {
    static int nt = [] { int v = env_int("WHALE_REV_NT", 128); return (v == 32 || v == 64 || v == 256) ? v : 128; }();
    return nt;
}
static int rev_minb() {
    static int mb = [] {
        const int nt = rev_nt(), v = env_int("WHALE_REV_MINB", 0);
        if (nt == 32) return 16;
        if (nt == 64) return 8;
        if (nt == 256) return 2;
#ifdef WHALE_DEV_BUILD
        return (v == 5 || v == 6 || v == 8) ? v : 4;
#else
        return (v == 3 || v == 5 || v == 6 || v == 8) ? v : 4;
#endif
    }();
    return mb;
}
constexpr int MAX_BINS = 8;
struct Bin {
    int off, count;
    size_t smem;
};

struct GraphSlot {  // one captured evaluation per (flags, condition)
    uint32_t key;
    int state;  // 0 new, 1 ran once uncaptured, 2 graph ready, 3 capture failed
#ifndef WHALE_EMU
    cudaGraphExec_t exec;
#else
    void* exec;
#endif
    int64_t launches = 0;
};
// Host-pointer evaluations replayed as ONE CUDA graph (H2D of θ, the kernels, D2H of the result).  WHALE_GRAPHS=1 / 0
// forces it on / off; unset = automatic: on when the evaluation is a single chain of launches (every plan in one
// shared-memory bin) — measured on the B200, C2: 0.2986 vs 0.3112 ms per call end to end (+4.2 %).  With several bins the
// evaluation forks onto side streams, and replaying THAT as a graph measured slower than plain launches in round 1
// (0.53 vs 0.48 ms), so it stays on the plain path.
static int graphs_mode() {  // -1 automatic, 0 off, 1 on
    static int mode = [] { const char* s = getenv("WHALE_GRAPHS"); return s ? (atoi(s) != 0 ? 1 : 0) : -1; }();
    return mode;
}

struct whale_data {
    whale_model* m = nullptr;
    int F = 0;
    std::vector<FamHdr> hdr;
    // tangent plans used for evaluations: plans[0] = value only; plans[1..] = gradient passes (one when the
    // full-tangent working set fits shared memory, else several passes over parameter chunks)
    std::vector<Plan*> plans;
    std::vector<Plan> chunk_plans;       // owned chunk plans (empty when plans[1] is the model's full plan)
    std::vector<size_t> out_off;         // offset (doubles) of each plan's per-family outputs in d_out_fam
    std::vector<int> perm[MAXPLAN];
    std::vector<Bin> bins[MAXPLAN];
    // per family x node facts kept from packing (shared-memory budgets are recomputed per plan)
    std::vector<uint32_t> f_ndent, f_ntent, f_heavy, f_stage16, f_rootwin;
    std::vector<GraphSlot> graphs;
    double last_bt_ms = 0.0;
    std::vector<uint32_t> roff_host[MAXPLAN];
    uint32_t* d_roff[MAXPLAN] = {};
    std::vector<double> work;           // predicted flops per family (pack time)
    std::vector<double> work_measured;  // SM cycles per family from the calibration pass (empty: not calibrated)
    size_t out_total = 0;
    cudaStream_t side[MAX_BINS] = {};
    cudaEvent_t ev_join[MAX_BINS] = {};
    cudaEvent_t ev_fork = nullptr;
    std::vector<unsigned char> arena_host;  // packing buffer (released after upload)
    size_t arena_bytes = 0;
    unsigned char* d_arena = nullptr;
    FamHdr* d_hdr = nullptr;
    int* d_perm[MAXPLAN] = {};
    double* d_out_fam = nullptr;  // [F*Kmax(plan1)]
    double* d_partial = nullptr;
    unsigned int* d_done = nullptr;  // k_dp's finished-CTA counter (fused reduction)
    double* d_ell = nullptr;
    long long* d_tim = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    uint64_t ell_total = 0;
    bool ell_valid = false;
    double* d_x_keep = nullptr;      // raw parameters and p_leaf of the evaluation that produced the kept ℓ: backtracking
    double* d_pleaf_keep = nullptr;  //   must use THESE (a model handle is shared by wm(θ) copies and evaluated at many θ)
    // aggregated work counters per node (over families): Σ C_e, Σ T_e (unfiltered triples of compat clades)
    std::vector<double> aggC, aggT;
    double aggG = 0, aggTroot = 0;
    int64_t algo_bytes = 0;
    std::vector<std::vector<uint32_t>> famC;  // [F][nn] compat counts (for ℓ layout)
    // reverse-mode gradient: transposed lists (a second arena), per-family budgets, adjoint-row offsets, the per-CTA
    // history slots and the family counters of the persistent launches
    bool rev = false;
    std::vector<RevHdr> rhdr;
    std::vector<unsigned char> rarena_host;
    std::vector<uint32_t> aoff_host;
    std::vector<int> rev_grid, rev_slot0;  // per bin: CTAs, first history slot
    RevHdr* d_rhdr = nullptr;
    unsigned char* d_rarena = nullptr;
    size_t rarena_bytes = 0;
    uint32_t* d_aoff = nullptr;
    double* d_hist = nullptr;
    size_t hist_stride = 0;
    unsigned int* d_next = nullptr;
    // backtracked trees of the last whale_backtrack / whale_track_sample call: persistent (grow-only) device and pinned host
    // buffers; the padded per-walk rows stay on the device for whale_trees_get / whale_trees_summary
    struct TreeBuf {
        long long W = 0;
        int S = 0, max_nodes = 0;
        bool valid = false;
        int32_t *d_cnt = nullptr, *d_st = nullptr, *d_g = nullptr, *d_e = nullptr, *d_t = nullptr, *d_p = nullptr;
        size_t capW = 0, capN = 0;           // walks, walks*max_nodes
        int4* d_stack = nullptr; size_t cap_stack = 0;
        double* d_u = nullptr; size_t cap_u = 0;
        double* d_xs = nullptr; size_t cap_xs = 0;
        int2* d_pairs = nullptr; int* d_famlist = nullptr; size_t cap_pairs = 0;
        long long* d_off = nullptr; size_t cap_off = 0;
        int4* d_packed = nullptr; size_t cap_packed = 0;
        unsigned long long* d_hash = nullptr; unsigned long long* d_dh = nullptr; int32_t *d_dc = nullptr, *d_df = nullptr, *d_nd = nullptr;
        size_t cap_hash = 0, cap_nd = 0;
        int32_t *h_cnt = nullptr, *h_st = nullptr; size_t hcapW = 0;      // pinned
        long long* h_off = nullptr; size_t hcap_off = 0;
        int32_t* h_packed = nullptr; size_t hcap_packed = 0;
        long long total_nodes = 0;
        bool packed_valid = false;
        // batched whale_track_sample: value-only plan clones (own slice-table buffers) and ℓ buffers for `nslot` posterior
        // draws evaluated back to back before ONE walk launch over all their (family, sample) pairs
        std::vector<Plan> slot_plans;
        double* d_ell_slots = nullptr; size_t cap_ell_slots = 0;
        const double** d_slot_eps = nullptr; const double2** d_slot_pp = nullptr; int* d_slot_off = nullptr;
        int nslot = 0;  // capacity of d_slot_off (ints)
    } tb;
    // peer-memory exchange (one process per GPU): own buffer, the peers' buffers as mapped here, step counter
    int peer_rank = -1, peer_world = 0;
    double* peer_bufs[16] = {};
    bool peer_open[16] = {};
    PeerDev* d_peer = nullptr;   // device-resident exchange state (buffers of all ranks, step counter)
    bool peer_uploaded = false;
    bool peer_fused = false;     // the evaluation just enqueued carried the exchange in the DP kernel's tail
};

// `subset`: which raw parameters this plan differentiates (empty = none: the value-only plan)
static void build_plan(const whale_model& m, const std::vector<char>& subset, Plan& pl) {
    const bool grad = !subset.empty();
    const int nn = m.nn;
    std::vector<std::vector<int>> act(nn);
    for (int oi = 0; oi < nn; oi++) {
        int e = m.order[oi];
        std::vector<int> a;
        if (grad) {
            auto add = [&](int s) { if (s >= 0 && subset[s]) a.push_back(s); };
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

// The plan the reverse-mode DP runs on: a leaf branch carries its own λ, μ (it has no children, so K <= 3 whatever P
// is), every other node is value only.
static void build_plan_hybrid(const whale_model& m, Plan& pl) {
    const int nn = m.nn;
    std::vector<std::vector<int>> act(nn);
    for (int e = 0; e < nn; e++) {
        if (m.kind[e] != WHALE_LEAF) continue;
        if (m.lam_slot[e] >= 0) act[e].push_back(m.lam_slot[e]);
        if (m.mu_slot[e] >= 0) act[e].push_back(m.mu_slot[e]);
        std::sort(act[e].begin(), act[e].end());
        act[e].erase(std::unique(act[e].begin(), act[e].end()), act[e].end());
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
            const int gp = act[e][k - 1];
            pl.act[(size_t)e * Kmax + k] = gp;
            pl.role[(size_t)e * Kmax + k] = (uint8_t)((gp == m.lam_slot[e] ? 1 : 0) | (gp == m.mu_slot[e] ? 2 : 0));
        }
        pl.cmap[((size_t)e * 2 + 0) * Kmax + 0] = 0;
        pl.cmap[((size_t)e * 2 + 1) * Kmax + 0] = 0;
    }
    pl.tab_len = off;
}

// The local plan: on an internal/WGD branch with slices, component 1 = the branch's own λ, 2 = its own μ (raw scale,
// role bits as in the other plans) and 3 = ϵ_0 of the branch itself (role bit 16: k_tables seeds ∂ϵ_0 = 1).  ϕ_i and ψ_i
// of a branch depend on nothing else, so these three partials are all the reverse pass needs from the slice tables.
static void build_plan_local(const whale_model& m, Plan& pl) {
    const int nn = m.nn;
    pl.Kmax = 4;
    pl.K.assign(nn, 1);
    pl.act.assign((size_t)nn * 4, -1);
    pl.cmap.assign((size_t)nn * 2 * 4, -1);
    pl.role.assign((size_t)nn * 4, 0);
    pl.toff.assign(nn, 0);
    size_t off = 0;
    for (int e = 0; e < nn; e++) {
        const bool on = m.kind[e] != WHALE_LEAF && m.kind[e] != WHALE_ROOT && m.nsl[e] > 0;
        pl.K[e] = on ? 4 : 1;
        pl.toff[e] = (int)off;
        off += (size_t)(m.nsl[e] + 1) * pl.K[e];
        if (on) {
            pl.act[(size_t)e * 4 + 1] = m.lam_slot[e]; pl.role[(size_t)e * 4 + 1] = 1;
            pl.act[(size_t)e * 4 + 2] = m.mu_slot[e];  pl.role[(size_t)e * 4 + 2] = 2;
            pl.role[(size_t)e * 4 + 3] = 16;
        }
        pl.cmap[((size_t)e * 2 + 0) * 4 + 0] = 0;
        pl.cmap[((size_t)e * 2 + 1) * 4 + 0] = 0;
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
    double *eps, *cx, *cy, *leaf, *cond, *ab, *shapeW;
    double2 *pp, *uv;
    long long* tim;
    if ((e = cudaMalloc((void**)&tim, 32 * sizeof(long long))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&eps, std::max<size_t>(pl.tab_len, 1) * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&pp, std::max<size_t>(pl.tab_len, 1) * sizeof(double2))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&uv, std::max<size_t>(pl.tab_len, 1) * sizeof(double2))) != cudaSuccess) return e;
    size_t nk = (size_t)nn * pl.Kmax * sizeof(double);
    if ((e = cudaMalloc((void**)&ab, 2 * nk)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&shapeW, NSHAPE * nk)) != cudaSuccess) return e;
    cudaMemset(shapeW, 0, NSHAPE * nk);
    if ((e = cudaMalloc((void**)&cx, nk)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&cy, nk)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&leaf, nk)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&cond, 4 * pl.Kmax * sizeof(double))) != cudaSuccess) return e;
    cudaMemset(cx, 0, nk); cudaMemset(cy, 0, nk); cudaMemset(leaf, 0, nk);
    cudaMemset(cond, 0, 4 * pl.Kmax * sizeof(double));
    std::vector<int> koff(nn, 0);
    int ktot = 0;
    for (int e2 = 0; e2 < nn; e2++) { koff[e2] = ktot; ktot += pl.K[e2]; }
    int* dkoff;
    if ((e = upload(koff, &dkoff)) != cudaSuccess) return e;
    pl.owned = {dK, dact, dtoff, dcmap, drole, eps, pp, uv, ab, cx, cy, leaf, shapeW, cond, tim, dkoff};
    pl.dev = PlanDev{pl.Kmax, dK, dact, dcmap, drole, dtoff, eps, pp, uv, ab, cx, cy, leaf, shapeW, cond, nullptr, nullptr, nullptr, tim};
    pl.dev.koff = dkoff;
    pl.dev.ktot = ktot;
    return cudaSuccess;
}

// fn(lo, hi) over [0, n) in chunks on the host threads (packer / budget loops over 10^5 families)
template <class Fn>
static void parallel_for(int n, int grain, Fn fn) {
    static const int hw = [] { const char* s = getenv("WHALE_PACK_THREADS"); int v = s ? atoi(s) : (int)std::thread::hardware_concurrency(); return std::max(1, std::min(v, 64)); }();
    const int nthr = std::max(1, std::min(hw, (n + grain - 1) / grain));
    if (nthr == 1) { fn(0, n); return; }
    std::atomic<int> next{0};
    auto worker = [&]() {
        for (;;) {
            const int lo = next.fetch_add(grain);
            if (lo >= n) break;
            fn(lo, std::min(n, lo + grain));
        }
    };
    std::vector<std::thread> th;
    for (int i = 0; i < nthr; i++) th.emplace_back(worker);
    for (auto& t : th) t.join();
}

static int g_device = 0;
static double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// grow-only device / pinned buffers (backtracking results)
template <class T>
static cudaError_t grow_dev(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    const size_t want = need + need / 4;
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(want, 1) * sizeof(T));
    if (e == cudaSuccess) *cap = want;
    return e;
}
template <class T>
static cudaError_t grow_pinned(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    const size_t want = need + need / 4;
    cudaError_t e = cudaMallocHost((void**)p, std::max<size_t>(want, 1) * sizeof(T));
    if (e == cudaSuccess) *cap = want;
    return e;
}



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
    for (int oi = 0; oi < nn; oi++) {
        int e = m->order[oi];
        (m->kind[e] == WHALE_LEAF ? m->leafnodes : m->inner).push_back(e);
    }
    int *o, *c0, *c1, *kd, *ns, *ls, *ms, *qs, *lo, *ln, *lfn, *inn;
    double *dt, *lp;
    CU(upload(m->leafnodes, &lfn)); CU(upload(m->inner, &inn));
    CU(upload(m->order, &o)); CU(upload(m->child0, &c0)); CU(upload(m->child1, &c1)); CU(upload(m->kind, &kd));
    CU(upload(m->nsl, &ns)); CU(upload(m->lam_slot, &ls)); CU(upload(m->mu_slot, &ms)); CU(upload(m->q_slot, &qs));
    CU(upload(m->lvl_off, &lo)); CU(upload(m->lvl_nodes, &ln)); CU(upload(m->dt, &dt)); CU(upload(m->leafP, &lp));
    m->owned = {o, c0, c1, kd, ns, ls, ms, qs, lo, ln, dt, lp, lfn, inn};
    m->dev = ModelDev{nn, o, c0, c1, kd, ns, dt, lp, ls, ms, qs, m->eta_slot, m->log_scale, m->root, m->P,
                      (int)m->lvl_off.size() - 1, lo, ln, (int)m->leafnodes.size(), lfn, (int)m->inner.size(), inn};
    for (int g = 0; g < 2; g++) {
        build_plan(*m, g == 1 ? std::vector<char>(m->P, 1) : std::vector<char>(), m->plan[g]);
        CU(upload_plan(m->plan[g], nn));
    }
    build_plan_hybrid(*m, m->planR);
    CU(upload_plan(m->planR, nn));
    build_plan_local(*m, m->planL);
    CU(upload_plan(m->planL, nn));
    {
        const Plan& pg = m->plan[1];
        const int KR = pg.K[m->root];
        m->rinv.assign((size_t)nn * KR, -1);
        for (int e = 0; e < nn; e++) {
            m->rinv[(size_t)e * KR] = 0;
            for (int k = 1; k < KR; k++) {
                const int gp = pg.act[(size_t)m->root * pg.Kmax + k];
                for (int j = 1; j < pg.K[e]; j++)
                    if (pg.act[(size_t)e * pg.Kmax + j] == gp) m->rinv[(size_t)e * KR + k] = (int16_t)j;
            }
        }
        CU(upload(m->rinv, &m->d_rinv));
        double* jac = nullptr;
        CU(cudaMalloc((void**)&jac, (size_t)nn * 8 * KR * sizeof(double)));
        m->plan[1].owned.push_back(jac);
        m->plan[1].dev.rinv = m->d_rinv;
        m->plan[1].dev.jac = jac;
    }
    {
        cudaDeviceProp pr;
        CU(cudaGetDeviceProperties(&pr, g_device));
        m->n_sm = std::max(1, pr.multiProcessorCount);
    }
    // θ and the leaf parameters live in ONE allocation ([P] | [nn]) so a host-pointer evaluation uploads both with one copy
    CU(cudaMalloc((void**)&m->d_x, (std::max(1, m->P) + nn) * sizeof(double)));
    m->d_pleaf = m->d_x + std::max(1, m->P);
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
    for (void* p : m->planR.owned) cudaFree(p);
    for (void* p : m->planL.owned) cudaFree(p);
    cudaFree(m->d_rinv);
    cudaFree(m->d_x); cudaFree(m->d_out);
    if (m->h_pin) cudaFreeHost(m->h_pin);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return WHALE_OK;
}

// ---- the packer: reference-layout CSR -> per-branch resolved device arena ----
static inline void pad4(std::vector<uint32_t>& w) { while (w.size() & 3) w.push_back(0); }

static size_t smem_need(const whale_model* m, const FamHdr& h, int plan, int Kmax_) {  // mirrors the carve-up in k_dp
    const size_t nn = m->nn, Kmax = Kmax_, NW = dp_nt() / 32;
    const size_t hdr = ((((7 * nn + 1) * sizeof(int) + nn * 2 * Kmax * sizeof(int16_t)) + 15) & ~size_t(15)) + nn * sizeof(NodeRec);
    return hdr + ((size_t)h.rows_len[plan] + h.scr_len[plan] + h.prod_len[plan]) * sizeof(double) +
           h.stage_bytes[plan] + NW * ((size_t)h.leafmax[plan] * sizeof(double) + h.leaf_stage);
}

static size_t tables_smem(const whale_model* m, const Plan& pl, bool shapes) {  // mirrors the carve-ups in k_tables
    const size_t nn = m->nn, nlvl = m->lvl_off.size() - 1;
    size_t ktot = 0;
    for (int e = 0; e < m->nn; e++) ktot += (size_t)pl.K[e];
    size_t need = (3 * nn + m->P + 6 * ktot) * sizeof(double) + (12 * nn + nlvl + 1) * sizeof(int) +
                  nn * 2 * pl.Kmax * sizeof(int16_t) + nn * pl.Kmax + 16;
    if (shapes)  // a leaf-shape CTA keeps its branch's projective and (ϕ, ψ) rows in shared memory
        for (int e : m->leafnodes) need = std::max(need, 2 * (size_t)(m->nsl[e] + 1) * pl.K[e] * sizeof(double2));
    return need;
}

// K1 launch: G table CTAs (each takes 1/G of the rows) + one CTA per leaf node for the tree-shape rows
static bool pdl_enabled() {  // programmatic dependent launch of the DP kernel behind the table kernel (WHALE_PDL=0: off)
    static bool f = env_int("WHALE_PDL", 1) != 0;
    return f;
}
static bool peer_fusion() {  // WHALE_PEER_FUSE=0: the exchange over ranks always as its own launch (k_peer_sum)
    static bool f = env_int("WHALE_PEER_FUSE", 1) != 0;
    return f;
}
static bool fused_reduce() {
    static bool f = env_int("WHALE_FUSED_REDUCE", 1) != 0;
    return f;
}
static unsigned tables_flags() {
    static unsigned f = env_int("WHALE_TABLES_CHAIN", 0) ? TAB_FORCE_CHAIN : 0u;
    return f;
}
static cudaError_t launch_tables(whale_model* m, Plan& pl, const double* d_x, const double* d_pleaf, cudaStream_t st,
                                 bool shapes) {
    shapes = shapes && !m->leafnodes.empty();
    const int G = (int)std::max<size_t>(1, std::min<size_t>(32, (pl.tab_len + TABLES_NT - 1) / TABLES_NT));
    const size_t smem = tables_smem(m, pl, shapes);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    LAUNCH(k_tables, G + (shapes ? (int)m->leafnodes.size() : 0), TABLES_NT, smem, st, m->dev, pl.dev, d_x, d_pleaf, G,
           tables_flags());
    g_launches++;
    return cudaSuccess;
}

// the three table sets of a reverse-mode evaluation in one launch (see Tables3 in whale_tables.cuh)
static cudaError_t launch_tables3(whale_model* m, const double* d_x, const double* d_pleaf, cudaStream_t st) {
    Plan* pls[3] = {&m->planR, &m->plan[1], &m->planL};
    Tables3 T3;
    size_t smem = 0;
    int total = 0;
    for (int i = 0; i < 3; i++) {
        const bool shapes = i == 0 && !m->leafnodes.empty();
        T3.PL[i] = pls[i]->dev;
        // (the full plan only runs the chain over the tree's height and writes the Jacobian table: one CTA)
        T3.G[i] = i == 1 ? 1 : (int)std::max<size_t>(1, std::min<size_t>(32, (pls[i]->tab_len + TABLES_NT - 1) / TABLES_NT));
        T3.n[i] = T3.G[i] + (shapes ? (int)m->leafnodes.size() : 0);
        total += T3.n[i];
        smem = std::max(smem, tables_smem(m, *pls[i], shapes));
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_tables3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    LAUNCH(k_tables3, total, TABLES_NT, smem, st, m->dev, T3, d_x, d_pleaf, tables_flags());
    g_launches++;
    return cudaSuccess;
}

// largest dynamic shared-memory request that still lets `per_sm` CTAs share an SM (228 KB per SM, 1 KB reserved per CTA)
static size_t occupancy_line(int per_sm) {  // (WHALE_OCC_LINE: another line, for tests of the unstaging rule on small families)
    static const int forced = env_int("WHALE_OCC_LINE", 0);
    return forced > 0 ? (size_t)forced : (size_t)(228 * 1024) / (size_t)std::max(1, per_sm) - 1024;
}
static size_t ctas_per_sm_by_smem(size_t need) { return (size_t)(228 * 1024) / (std::min<size_t>(need, 227 * 1024) + 1024); }

// shared-memory budget of every family under tangent plan `pl` (stored at index g); returns the largest need
static size_t set_budgets(whale_data* D, int g, const Plan& pl) {
    const whale_model* m = D->m;
    const int nn = m->nn;
    std::atomic<size_t> worst_a{0};
    D->roff_host[g].assign((size_t)D->F * nn, 0);
    parallel_for(D->F, 256, [&](int f_lo, int f_hi) {
    size_t worst = 0;
    for (int f = f_lo; f < f_hi; f++) {
        FamHdr& H = D->hdr[f];
        const std::vector<uint32_t>& Cs = D->famC[f];
        const uint32_t* ns = D->f_heavy.data() + (size_t)f * nn;  // heavy-leaf flags
        const uint32_t* s16 = D->f_stage16.data() + (size_t)f * nn;
        uint32_t mxinner = 0, mxleaf = 0;
        const uint32_t prod = 0;
        size_t stg = 0;
        for (int e = 0; e < nn; e++) {
            const uint32_t K = (uint32_t)pl.K[e], ck = Cs[e] * (uint32_t)RS(K);
            if (s16[e] > 0)  // lists + (inner nodes with K <= 8) the ϕ/ψ rows
                stg = std::max(stg, (size_t)s16[e] + (K <= 8 ? (size_t)(m->nsl[e] + 1) * K : 0));
            if (m->kind[e] == WHALE_LEAF) {
                if (!ns[e]) mxleaf = std::max(mxleaf, ck);
                else mxinner = std::max(mxinner, ck);  // heavy leaf branch: block-scope scratch row
            } else if (m->kind[e] != WHALE_ROOT) {
                mxinner = std::max(mxinner, ck);
            }
        }
        auto even = [](uint32_t v) { return (v + 1) & ~1u; };
        const uint32_t scr = even(mxinner);
        std::vector<int> roff(nn + 1);
        const int rows = place_rows(nn, (int)m->leafnodes.size(), m->leafnodes.data(), (int)m->inner.size(), m->inner.data(),
                                    m->child0.data(), m->child1.data(), m->kind.data(),
                                    [&](int e2) { return (int)(Cs[e2] * (uint32_t)RS(pl.K[e2])); }, roff.data());
        H.rows_len[g] = even((uint32_t)rows);
        for (int e = 0; e < nn; e++) D->roff_host[g][(size_t)f * nn + e] = (uint32_t)roff[e];
        H.scr_len[g] = scr;
        H.prod_len[g] = even(prod);
        H.leafmax[g] = even(mxleaf);
        // long lists (large CCDs) are not staged: the kernel then reads them from global memory in place
        // (a batch smaller than the GPU is latency-bound: stage whatever fits, +20 % on 64 C4 families; a large batch is
        //  bound by how many families an SM holds: keep the staging buffer small)
        const size_t STAGE_MAX = (size_t)env_int("WHALE_STAGE_MAX", D->F <= 160 ? 160 * 1024 : 24 * 1024);
        H.stage_bytes[g] = 16 * stg > STAGE_MAX ? 0u : (uint32_t)(16 * stg);
        worst = std::max(worst, smem_need(m, H, g, pl.Kmax));
        if (env_int("WHALE_DEBUG", 0) >= 2)
            fprintf(stderr, "[whale] fam %d plan %d: G %u root levels %u (window %u) rows %u scr %u prod %u leafmax %u stage %u leaf_stage %u -> %zu B\n", f, g,
                    H.G, H.nlev, H.rootwin, H.rows_len[g], H.scr_len[g], H.prod_len[g], H.leafmax[g], H.stage_bytes[g], H.leaf_stage,
                    smem_need(m, H, g, pl.Kmax));
    }
    size_t cur = worst_a.load();
    while (worst > cur && !worst_a.compare_exchange_weak(cur, worst)) {}
    });
    // A launch requests the largest need of its bin for every CTA: a handful of families just above the size that still lets
    // dp_minb() families share an SM would cost the whole batch a quarter of its resident families (measured on the B200:
    // a 1000-family shard whose largest family needs 57.7 KB ran at 3 per SM, k_dp 0.338 instead of 0.280 ms).  Those few
    // read their lists in place instead (stage_bytes = 0) when that brings them under the line.
    const size_t line = occupancy_line(dp_minb());
    if (worst_a.load() > line) {
        int over = 0;
        for (int f = 0; f < D->F; f++) over += smem_need(m, D->hdr[f], g, pl.Kmax) > line;
        if (over * 4 <= D->F) {
            size_t worst = 0;
            for (int f = 0; f < D->F; f++) {
                FamHdr& H = D->hdr[f];
                if (smem_need(m, H, g, pl.Kmax) > line && smem_need(m, H, g, pl.Kmax) - H.stage_bytes[g] <= line) H.stage_bytes[g] = 0;
                worst = std::max(worst, smem_need(m, H, g, pl.Kmax));
            }
            worst_a.store(worst);
        }
    }
    return worst_a.load();
}

static int32_t enqueue_eval(whale_model* m, whale_data* D, const double* d_x, int32_t condition, uint32_t flags,
                            double* d_out, cudaStream_t st);
static int32_t enqueue_peer_sum(whale_data* D, double* d_out, cudaStream_t st);
static int32_t peer_state(whale_data* D, PeerDev** out);
static int32_t finalize_data(whale_model* m, whale_data* D, whale_data_t* out);

// Launch order of plan g's families for a given per-family work vector (predicted flops at pack time, measured SM
// cycles after the calibration pass).  Bins by shared-memory need (one per occupancy class, launched concurrently);
// inside a bin the heaviest families go first (the hardware hands CTAs to SM slots in index order = LPT list
// scheduling).  A "paired" order for the 1–2 wave case (heaviest alone, lightest + middle paired) was measured 4 %
// SLOWER on the B200: an SM with fewer resident families runs each of them faster (the SM's shared-memory/issue
// throughput is what is conserved, not the slot count), so LPT's ragged tail costs little.
static size_t smem_need_rev(const whale_model* m, const FamHdr& h, const RevHdr& r);
static cudaError_t order_families(whale_data* D, int g, const std::vector<double>& work) {
    const whale_model* m = D->m;
    const Plan& pl = *D->plans[g];
    const int F = D->F;
    std::vector<int>& perm = D->perm[g];
    perm.resize(F);
    std::iota(perm.begin(), perm.end(), 0);
    std::vector<size_t> need(F);
    const bool rev = D->rev && g == 1;
    for (int f = 0; f < F; f++) need[f] = rev ? smem_need_rev(m, D->hdr[f], D->rhdr[f]) : smem_need(m, D->hdr[f], g, pl.Kmax);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return need[a] != need[b] ? need[a] > need[b] : work[a] > work[b]; });
    std::vector<Bin>& bins = D->bins[g];
    bins.clear();
    int i = 0;
    // a bin = families that allow the same number of resident CTAs per SM (capped by the register limit):
    // a smaller shared-memory request buys nothing once registers are the limiter
    auto cls = [&](size_t nd) { return std::min<size_t>((size_t)(rev ? rev_minb() : dp_minb()), ctas_per_sm_by_smem(nd)); };
    while (i < F) {
        Bin b{i, 0, need[perm[i]]};
        while (i < F && (cls(need[perm[i]]) == cls(b.smem) || b.count < 64 || (int)bins.size() >= MAX_BINS - 1)) { i++; b.count++; }
        bins.push_back(b);
    }
    // merge a tiny last bin into its predecessor
    if (bins.size() > 1 && bins.back().count < 64) { bins[bins.size() - 2].count += bins.back().count; bins.pop_back(); }
    for (const Bin& b : bins)
        std::stable_sort(perm.begin() + b.off, perm.begin() + b.off + b.count, [&](int x, int y) { return work[x] > work[y]; });
    if (env_int("WHALE_DEBUG_BINS", 0))
        for (const Bin& b : bins)
            fprintf(stderr, "[whale] plan %d%s bin: off %d count %d smem %zu (families per SM %zu); largest need %zu, smallest %zu\n", g,
                    rev ? " (reverse)" : "", b.off, b.count, b.smem, cls(b.smem), need[perm[0]], need[perm[F - 1]]);
    if (rev) {  // persistent CTAs: as many as the GPU holds at this bin's shared-memory need, one history slot each
        D->rev_grid.clear(); D->rev_slot0.clear();
        int slot = 0;
        for (const Bin& b : bins) {
            const int per_sm = (int)std::max<size_t>(1, cls(b.smem));
            const int cap = env_int("WHALE_REV_GRID", per_sm * m->n_sm);
            const int grid = std::max(1, std::min(b.count, cap));
            D->rev_grid.push_back(grid);
            D->rev_slot0.push_back(slot);
            slot += grid;
        }
        if (D->d_hist) { cudaFree(D->d_hist); D->d_hist = nullptr; }
        cudaError_t e = cudaMalloc((void**)&D->d_hist, std::max<size_t>((size_t)slot * D->hist_stride, 1) * sizeof(double));
        if (e != cudaSuccess) return e;
    }
    if (!D->d_perm[g]) {
        cudaError_t e = cudaMalloc((void**)&D->d_perm[g], std::max<size_t>(F, 1) * sizeof(int));
        if (e != cudaSuccess) return e;
    }
    return cudaMemcpy(D->d_perm[g], perm.data(), (size_t)F * sizeof(int), cudaMemcpyHostToDevice);
}

// One family of the packer: everything it contributes, self-contained, so that families can be packed on all host threads
struct FamPack {
    std::vector<unsigned char> blob;
    FamHdr H;
    std::vector<uint32_t> Cs, ndent, ntent, heavy, stage16;
    std::vector<double> aggC, aggT;
    double aggTroot = 0.0, work = 0.0;
    int64_t algo = 0;
    uint64_t ell = 0;
    uint32_t rootwin = 0;
};
static int32_t pack_family(const whale_model* m, const whale_ccd_desc* d, int f, FamPack& O, std::vector<int32_t>& lidx) {
    const int nn = m->nn;
    O.ndent.assign(nn, 0); O.ntent.assign(nn, 0); O.heavy.assign(nn, 0); O.stage16.assign(nn, 0);
    O.aggC.assign(nn, 0.0); O.aggT.assign(nn, 0.0);
    O.blob.clear();
    {
        const int64_t cb = d->clade_off[f];
        const int G = (int)(d->clade_off[f + 1] - cb);
        if (G < 1 || G > 65535) { return fail(WHALE_ERR_ARG, "family %d: %d clades (must be 1..65535, UInt16 ids)", f, G); }
        const int32_t* nleaf = d->clade_nleaf + cb;
        const int64_t* soff = d->split_off + cb;
        const int64_t* coff = d->compat_off + (int64_t)f * nn;
        lidx.assign((size_t)nn * G, -1);
        std::vector<uint32_t>& Cs = O.Cs;
        Cs.assign(nn, 0);
        FamHdr& H = O.H;
        memset(&H, 0, sizeof(H));
        for (int e = 0; e < nn; e++) {
            int C = (int)(coff[e + 1] - coff[e]);
            Cs[e] = C;
            int prev = -1;
            for (int j = 0; j < C; j++) {
                int g = d->compat[coff[e] + j];
                if (g < 0 || g >= G || g <= prev) { return fail(WHALE_ERR_ARG, "family %d node %d: compat list must be ascending clade ids", f, e); }
                prev = g;
                lidx[(size_t)e * G + g] = j;
            }
        }
        if ((int)Cs[m->root] != G) { return fail(WHALE_ERR_ARG, "family %d: every clade must be compatible with the root", f); }
        for (int c = 1; c < G; c++)
            if (nleaf[c] < nleaf[c - 1]) { return fail(WHALE_ERR_ARG, "family %d: clades must be sorted by size", f); }
        // blob assembly
        std::vector<NodeRec> recs(nn);
        std::vector<uint32_t> wordsv;   // pointer / loss / level words; every array starts on a 16-byte boundary
        std::vector<Ent> entsv;         // term entries
        uint32_t sumC = 0, nlev = 0;
        size_t leaf_stage = 0;
        uint32_t rootwin = 0;
        std::vector<size_t> stage16(nn, 0);  // 16-byte words staged per node, excluding the ϕ/ψ rows
        double wk = 0.0;
        for (int e = 0; e < nn; e++) {
            NodeRec& R = recs[e];
            memset(&R, 0, sizeof(R));
            const int C = (int)Cs[e];
            R.C = C;
            sumC += C;
            const int kind = m->kind[e];
            int nonleaf = 0;
            double Te = 0;
            for (int j = 0; j < C; j++) {
                int g = d->compat[coff[e] + j];
                if (nleaf[g] > 1) nonleaf++;
                Te += (double)(soff[g + 1] - soff[g]);
            }
            R.nonleaf = nonleaf;
            O.aggC[e] = C;
            O.aggT[e] = Te;
            // (1) same-branch terms: within-branch duplication (src/core.jl:178-185), Πroot (:151-158);
            //     for WGD nodes the same list drives Πwgdretention on the child's row (:187-194), the
            //     child's compat list being identical.
            pad4(wordsv);
            R.dptr_off = (uint32_t)wordsv.size();
            R.dent_off = (uint32_t)entsv.size();
            {
                const int src = (kind == WHALE_WGD) ? m->child0[e] : e;
                if (kind == WHALE_WGD && Cs[src] != (uint32_t)C) { return fail(WHALE_ERR_ARG, "family %d: WGD node %d and its child must have identical compat lists", f, e); }
                for (int j = 0; j < C; j++) {
                    int g = d->compat[coff[e] + j];
                    wordsv.push_back((uint32_t)entsv.size() - R.dent_off);
                    for (int64_t t = soff[g]; t < soff[g + 1]; t++) {
                        int g1 = d->g1[t], g2 = d->g2[t];
                        if (g1 < 0 || g1 >= G || g2 < 0 || g2 >= G) { return fail(WHALE_ERR_ARG, "family %d: triple out of range", f); }
                        int i1 = lidx[(size_t)src * G + g1], i2 = lidx[(size_t)src * G + g2];
                        if (i1 < 0 || i2 < 0) continue;  // getl == 0 (src/ccd.jl:43)
                        entsv.push_back(Ent{(uint16_t)i1, (uint16_t)i2, 0u, d->p[t]});
                    }
                }
                wordsv.push_back((uint32_t)entsv.size() - R.dent_off);
                R.ndent = (uint32_t)entsv.size() - R.dent_off;
                wk += (double)R.ndent * (m->nsl[e] + 1) + (double)C * (m->nsl[e] + 1);
            }
            // (1b) the slice loop's lane table: teams of 2^glog slots per clade, at most two terms per slot (E <= 64),
            //      largest teams first (keeps teams aligned and makes the team size warp-uniform-monotone)
            pad4(wordsv);
            R.slot_off = (uint32_t)wordsv.size();
            if (kind != WHALE_ROOT && m->nsl[e] > 0) {
                if (R.ndent > 65535) { return fail(WHALE_ERR_CAPACITY, "family %d node %d: %u same-branch terms (> 65535)", f, e, R.ndent); }
                const uint32_t* dp = wordsv.data() + R.dptr_off;
                std::vector<std::pair<int, int>> order;  // (-glog, cell)
                std::vector<int> glogs(C);
                for (int j = 0; j < C; j++) {
                    const uint32_t E = dp[j + 1] - dp[j];
                    int gl = 0;
                    while (gl < 5 && (2u << gl) < E) gl++;  // team size G = 2^gl >= ceil(E/2), capped at 32
                    glogs[j] = gl;
                    order.push_back({-gl, j});
                }
                std::stable_sort(order.begin(), order.end());
                std::vector<Slot> slots;
                for (auto& oc : order) {
                    const int j = oc.second, gl = glogs[j], G = 1 << gl;
                    const uint32_t E = dp[j + 1] - dp[j], first = dp[j];
                    for (int l = 0; l < G; l++) {
                        const uint32_t cnt = (uint32_t)l < E ? (E - l + G - 1) / G : 0;
                        if (cnt > 255) { return fail(WHALE_ERR_CAPACITY, "family %d node %d: clade with %u terms", f, e, E); }
                        slots.push_back(Slot{(uint16_t)j, (uint8_t)gl, (uint8_t)cnt, (uint16_t)(first + l), (uint16_t)G});
                    }
                }
                R.nslots = (uint32_t)slots.size();
                const size_t w0 = wordsv.size();
                wordsv.resize(w0 + 2 * slots.size());
                if (!slots.empty()) memcpy(wordsv.data() + w0, slots.data(), slots.size() * sizeof(Slot));
            }
            // (1c) leaf branch in closed form: coefficients C_σ[γ] over tree shapes (see SHAPES in whale_common.cuh)
            if (kind == WHALE_LEAF && nonleaf > 0) {
                int maxsz = 0;
                for (int j = 0; j < C; j++) maxsz = std::max(maxsz, (int)nleaf[d->compat[coff[e] + j]]);
                if (maxsz <= SHAPE_MAXLEAVES) {
                    static const int SA[NSHAPE] = SHAPE_A, SB[NSHAPE] = SHAPE_B;
                    auto compose = [&](int s1, int s2) -> int {
                        const int lo = std::min(s1, s2), hi = std::max(s1, s2);
                        for (int sg = 1; sg < NSHAPE; sg++) if (SA[sg] == lo && SB[sg] == hi) return sg;
                        return -1;
                    };
                    std::vector<double> coef((size_t)C * NSHAPE, 0.0);
                    const uint32_t* dp = wordsv.data() + R.dptr_off;
                    for (int j = 0; j < C; j++) {
                        if (dp[j + 1] == dp[j]) { if (nleaf[d->compat[coff[e] + j]] == 1) coef[(size_t)j * NSHAPE] = 1.0; continue; }
                        for (uint32_t t = dp[j]; t < dp[j + 1]; t++) {
                            const Ent& en = entsv[R.dent_off + t];
                            for (int s1 = 0; s1 < NSHAPE; s1++) {
                                const double c1 = coef[(size_t)en.i1 * NSHAPE + s1];
                                if (c1 == 0.0) continue;
                                for (int s2 = 0; s2 < NSHAPE; s2++) {
                                    const double c2 = coef[(size_t)en.i2 * NSHAPE + s2];
                                    if (c2 == 0.0) continue;
                                    const int sg = compose(s1, s2);
                                    if (sg >= 0) coef[(size_t)j * NSHAPE + sg] += en.p * c1 * c2;
                                }
                            }
                        }
                    }
                    pad4(wordsv);
                    R.sptr_off = (uint32_t)wordsv.size();
                    R.sent_off = (uint32_t)entsv.size();
                    for (int j = 0; j < C; j++) {
                        wordsv.push_back((uint32_t)entsv.size() - R.sent_off);
                        for (int sg = 0; sg < NSHAPE; sg++)
                            if (coef[(size_t)j * NSHAPE + sg] != 0.0) entsv.push_back(Ent{(uint16_t)sg, 0, 0u, coef[(size_t)j * NSHAPE + sg]});
                    }
                    wordsv.push_back((uint32_t)entsv.size() - R.sent_off);
                }
            }
            // (2) speciation terms at row 1 (src/core.jl:160-170), Πloss child indices (:172-176), root levels
            pad4(wordsv);
            R.tptr_off = (uint32_t)wordsv.size();
            R.tent_off = (uint32_t)entsv.size();
            if (kind == WHALE_INTERNAL || kind == WHALE_ROOT) {
                const int fch = m->child0[e], gch = m->child1[e];
                for (int j = 0; j < C; j++) {
                    int g = d->compat[coff[e] + j];
                    wordsv.push_back((uint32_t)entsv.size() - R.tent_off);
                    for (int64_t t = soff[g]; t < soff[g + 1]; t++) {
                        int g1 = d->g1[t], g2 = d->g2[t];
                        int f1 = lidx[(size_t)fch * G + g1], gg2 = lidx[(size_t)gch * G + g2];
                        int gg1 = lidx[(size_t)gch * G + g1], f2 = lidx[(size_t)fch * G + g2];
                        // pad = (ordinal of the triple within its clade) << 1 | ("gf": γ1 goes to child g) — used by
                        // the backtracker to keep the reference's per-triple order (src/track.jl:304-324,369-397)
                        const uint32_t ord = (uint32_t)(t - soff[g]) << 1;
                        if (f1 >= 0 && gg2 >= 0) entsv.push_back(Ent{(uint16_t)f1, (uint16_t)gg2, ord, d->p[t]});
                        if (gg1 >= 0 && f2 >= 0) entsv.push_back(Ent{(uint16_t)f2, (uint16_t)gg1, ord | 1u, d->p[t]});
                    }
                }
                wordsv.push_back((uint32_t)entsv.size() - R.tent_off);
                R.ntent = (uint32_t)entsv.size() - R.tent_off;
                wk += (double)R.ntent;
                for (int j = 0; j < C; j++) wordsv.push_back((uint32_t)lidx[(size_t)fch * G + d->compat[coff[e] + j]]);
                for (int j = 0; j < C; j++) wordsv.push_back((uint32_t)lidx[(size_t)gch * G + d->compat[coff[e] + j]]);
                if (kind == WHALE_ROOT) {
                    for (int c = 0; c < G; c++)
                        if (c == 0 || nleaf[c] != nleaf[c - 1]) { wordsv.push_back((uint32_t)c); nlev++; }
                    wordsv.push_back((uint32_t)G);
                    O.aggTroot = (double)(soff[G] - soff[0]);
                    // largest number of products one root level needs at once (Πroot + speciation terms)
                    const uint32_t* dpr = wordsv.data() + R.dptr_off;
                    const uint32_t* tpr = wordsv.data() + R.tptr_off;
                    uint32_t lvmax = 0;
                    int cprev = 0;
                    for (int c = 1; c <= G; c++)
                        if (c == G || nleaf[c] != nleaf[c - 1]) {
                            lvmax = std::max(lvmax, (dpr[c] - dpr[cprev]) + (tpr[c] - tpr[cprev]));
                            cprev = c;
                        }
                    rootwin = lvmax;
                }
            }
            // local cell -> clade id (the compat list itself), for the backtracker's output
            pad4(wordsv);
            for (int j = 0; j < C; j++) wordsv.push_back((uint32_t)d->compat[coff[e] + j]);
            if (kind == WHALE_LEAF) {
                const size_t l16 = (size_t)R.ndent + ((size_t)R.nslots + 1) / 2;  // [dents | slots]
                if (R.nslots <= HEAVY_SLOTS) leaf_stage = std::max(leaf_stage, 16 * l16);
                else stage16[e] = l16;  // heavy leaf branch: block scope
            } else {  // what k_dp stages in shared memory for this node: [dents | slots | dptr | tptr.. | tents]
                size_t nd16 = kind == WHALE_ROOT ? 0 : R.ndent;
                size_t dp16 = ((size_t)C + 1 + 3) / 4 + (kind == WHALE_ROOT ? 0 : ((size_t)R.nslots + 1) / 2);
                size_t tp16 = kind == WHALE_WGD ? 0 : (3 * (size_t)C + 1 + (kind == WHALE_ROOT ? nlev + 1 : 0) + 3) / 4;
                size_t tn16 = kind == WHALE_INTERNAL ? R.ntent : 0;
                stage16[e] = nd16 + dp16 + tp16 + tn16 + (kind == WHALE_ROOT ? 2 * (size_t)rootwin : 0);
            }
            O.ell += (uint64_t)(m->nsl[e] + 1) * C;
        }
        pad4(wordsv);
        // serialise: NodeRec[nn] | words | entries ; make offsets blob-relative
        std::vector<unsigned char>& A = O.blob;
        size_t base = 0;
        size_t words_at = (size_t)nn * sizeof(NodeRec);  // 32-byte records: 16-byte aligned
        size_t ents_at = words_at + wordsv.size() * 4;   // words padded to 4 -> 16-byte aligned
        size_t total = ents_at + entsv.size() * sizeof(Ent);
        A.resize(base + total, 0);
        for (int e = 0; e < nn; e++) {
            NodeRec& R = recs[e];
            R.dptr_off += (uint32_t)(words_at / 4);
            R.slot_off += (uint32_t)(words_at / 4);
            R.tptr_off += (uint32_t)(words_at / 4);
            if (R.sptr_off) { R.sptr_off += (uint32_t)(words_at / 4); R.sent_off += (uint32_t)(ents_at / 16); }
            R.dent_off += (uint32_t)(ents_at / 16);
            R.tent_off += (uint32_t)(ents_at / 16);
        }
        memcpy(A.data() + base, recs.data(), words_at);
        if (!wordsv.empty()) memcpy(A.data() + base + words_at, wordsv.data(), wordsv.size() * 4);
        if (!entsv.empty()) memcpy(A.data() + base + ents_at, entsv.data(), entsv.size() * sizeof(Ent));
        H.G = G;
        H.nlev = nlev;
        H.leaf_stage = (uint32_t)leaf_stage;
        H.blob_bytes = (uint32_t)total;
        H.rootwin = rootwin;
        O.work = wk;
        // SURVEY §8d algorithmic bytes per evaluation: 12·T + 2·Γ + 4·Σ_e C_e
        O.algo = 12 * (soff[G] - soff[0]) + 2 * (int64_t)G + 4 * (int64_t)sumC;
        for (int e = 0; e < nn; e++) {
            O.ndent[e] = recs[e].ndent;
            O.ntent[e] = recs[e].ntent;
            O.heavy[e] = recs[e].nslots > HEAVY_SLOTS ? 1u : 0u;
            O.stage16[e] = (uint32_t)stage16[e];
        }
        O.rootwin = rootwin;
    
    }
    return WHALE_OK;
}

int32_t whale_data_create(whale_model_t m, const whale_ccd_desc* d, whale_data_t* out) {
    if (!m || !d || !out) return fail(WHALE_ERR_ARG, "null argument");
    const int nn = m->nn, F = d->n_fam;
    if (F <= 0) return fail(WHALE_ERR_ARG, "n_fam must be positive");
    CU(cudaSetDevice(m->device));
    const double t0 = now_s();
    // pack the families on all host threads (each family's blob is self-contained), then lay the blobs out in order
    std::vector<FamPack> packs(F);
    int nthr = env_int("WHALE_PACK_THREADS", (int)std::thread::hardware_concurrency());
    nthr = std::max(1, std::min(nthr, std::min(64, (F + 31) / 32)));
    std::atomic<int> next{0};
    std::atomic<int> failed{0};
    std::string err;
    int32_t err_code = WHALE_OK;
    std::mutex err_mu;
    auto worker = [&]() {
        std::vector<int32_t> lidx;  // local index of clade γ at node e: lidx[e*G + γ]
        for (;;) {
            const int f0 = next.fetch_add(16);
            if (f0 >= F || failed.load()) break;
            for (int f = f0; f < std::min(F, f0 + 16); f++) {
                const int32_t rc = pack_family(m, d, f, packs[f], lidx);
                if (rc != WHALE_OK) {
                    std::lock_guard<std::mutex> lk(err_mu);
                    if (!failed.exchange(1)) { err = g_err; err_code = rc; }
                    return;
                }
            }
        }
    };
    if (nthr == 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthr; i++) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    if (failed.load()) return fail(err_code, "%s", err.c_str());
    auto* D = new whale_data();
    D->m = m;
    D->F = F;
    D->hdr.resize(F);
    D->aggC.assign(nn, 0.0);
    D->aggT.assign(nn, 0.0);
    D->famC.resize(F);
    D->work.assign(F, 0.0);
    D->f_ndent.resize((size_t)F * nn); D->f_ntent.resize((size_t)F * nn); D->f_heavy.resize((size_t)F * nn);
    D->f_stage16.resize((size_t)F * nn); D->f_rootwin.resize(F);
    uint64_t ell_total = 0;
    int64_t algo_bytes = 0;
    size_t bytes = 0;
    for (int f = 0; f < F; f++) {
        FamPack& O = packs[f];
        D->hdr[f] = O.H;
        D->hdr[f].base = bytes;
        D->hdr[f].ell_off = ell_total;
        bytes += (O.blob.size() + 15) & ~size_t(15);
        ell_total += O.ell;
        algo_bytes += O.algo;
        D->famC[f].swap(O.Cs);
        D->work[f] = O.work;
        D->aggG += D->hdr[f].G;
        D->aggTroot += O.aggTroot;
        for (int e = 0; e < nn; e++) {
            D->aggC[e] += O.aggC[e]; D->aggT[e] += O.aggT[e];
            D->f_ndent[(size_t)f * nn + e] = O.ndent[e]; D->f_ntent[(size_t)f * nn + e] = O.ntent[e];
            D->f_heavy[(size_t)f * nn + e] = O.heavy[e]; D->f_stage16[(size_t)f * nn + e] = O.stage16[e];
        }
        D->f_rootwin[f] = O.rootwin;
    }
    std::vector<unsigned char>& A = D->arena_host;
    A.assign(bytes, 0);
    {
        std::atomic<int> nx{0};
        auto copier = [&]() {
            for (;;) {
                const int f0 = nx.fetch_add(64);
                if (f0 >= F) break;
                for (int f = f0; f < std::min(F, f0 + 64); f++) {
                    if (!packs[f].blob.empty()) memcpy(A.data() + D->hdr[f].base, packs[f].blob.data(), packs[f].blob.size());
                    std::vector<unsigned char>().swap(packs[f].blob);
                }
            }
        };
        if (nthr == 1) copier();
        else {
            std::vector<std::thread> th;
            for (int i = 0; i < nthr; i++) th.emplace_back(copier);
            for (auto& t : th) t.join();
        }
    }
    std::vector<FamPack>().swap(packs);
    D->ell_total = ell_total;
    D->algo_bytes = algo_bytes;
    if (env_int("WHALE_DEBUG", 0) >= 1) fprintf(stderr, "[whale] packed %d families on %d threads in %.2f s (%zu bytes)\n", F, nthr, now_s() - t0, bytes);
    return finalize_data(m, D, out);
}

// ---- reverse-mode gradient: transposed lists, built from the forward blobs (so the arena cache needs nothing new) ----
// lane table of a slice loop: teams of 2^glog slots per cell, at most two terms per slot while E <= 64, largest teams
// first (keeps teams aligned and makes the team size monotone over the warps)
static bool make_slots(const uint32_t* ptr, int C, std::vector<Slot>& slots) {
    std::vector<std::pair<int, int>> order;  // (-glog, cell)
    std::vector<int> glogs(C);
    for (int j = 0; j < C; j++) {
        const uint32_t E = ptr[j + 1] - ptr[j];
        int gl = 0;
        while (gl < 5 && (2u << gl) < E) gl++;
        glogs[j] = gl;
        order.push_back({-gl, j});
    }
    std::stable_sort(order.begin(), order.end());
    slots.clear();
    for (auto& oc : order) {
        const int j = oc.second, gl = glogs[j], G = 1 << gl;
        const uint32_t E = ptr[j + 1] - ptr[j], first = ptr[j];
        for (int l = 0; l < G; l++) {
            const uint32_t cnt = (uint32_t)l < E ? (E - l + G - 1) / G : 0;
            if (cnt > 255 || first + l > 65535u) return false;
            slots.push_back(Slot{(uint16_t)j, (uint8_t)gl, (uint8_t)cnt, (uint16_t)(first + l), (uint16_t)G});
        }
    }
    return true;
}

// returns false when a family cannot be represented (u16 entry indices): the handle then keeps forward tangents
// transposed lists of ONE family from its forward blob (self-contained: families are processed on all host threads)
static bool reverse_family(const whale_model* m, const unsigned char* blobF, int nn, std::vector<unsigned char>& blob,
                           uint32_t& hist_len) {
    {
        const NodeRec* recs = reinterpret_cast<const NodeRec*>(blobF);
        const uint32_t* words = reinterpret_cast<const uint32_t*>(blobF);
        const Ent* ents = reinterpret_cast<const Ent*>(blobF);
        std::vector<RevRec> rr(nn);
        std::vector<uint32_t> wv;
        std::vector<Ent> ev;
        uint64_t hoff = 0;
        for (int e = 0; e < nn; e++) {
            RevRec& Q = rr[e];
            memset(&Q, 0, sizeof(Q));
            const NodeRec& R = recs[e];
            const int C = (int)R.C, kind = m->kind[e];
            if (kind == WHALE_LEAF) continue;
            // (1) same-branch terms, transposed: the term (c; i1, i2, p) appears under i1 as (c, i2) and under i2 as (c, i1)
            {
                const uint32_t* dptr = words + R.dptr_off;
                const Ent* dents = ents + R.dent_off;
                std::vector<std::vector<Ent>> lists(C);
                for (int c = 0; c < C; c++)
                    for (uint32_t t = dptr[c]; t < dptr[c + 1]; t++) {
                        const Ent& en = dents[t];
                        lists[en.i1].push_back(Ent{(uint16_t)c, en.i2, 0u, en.p});
                        lists[en.i2].push_back(Ent{(uint16_t)c, en.i1, 0u, en.p});
                    }
                pad4(wv);
                Q.bptr_off = (uint32_t)wv.size();
                Q.bent_off = (uint32_t)ev.size();
                for (int c = 0; c < C; c++) {
                    wv.push_back((uint32_t)ev.size() - Q.bent_off);
                    ev.insert(ev.end(), lists[c].begin(), lists[c].end());
                }
                wv.push_back((uint32_t)ev.size() - Q.bent_off);
                Q.nbent = (uint32_t)ev.size() - Q.bent_off;
                pad4(wv);
                Q.bslot_off = (uint32_t)wv.size();
                if (kind != WHALE_ROOT && m->nsl[e] > 0) {
                    std::vector<Slot> slots;
                    if (!make_slots(wv.data() + Q.bptr_off, C, slots)) return false;
                    Q.nbslots = (uint32_t)slots.size();
                    const size_t w0 = wv.size();
                    wv.resize(w0 + 2 * slots.size());
                    if (!slots.empty()) memcpy(wv.data() + w0, slots.data(), slots.size() * sizeof(Slot));
                }
            }
            // (2) speciation terms at row 1, grouped by the child's cell; Πloss transposed (the parent's cell of a child's cell)
            if (kind == WHALE_INTERNAL || kind == WHALE_ROOT) {
                const uint32_t* tptr = words + R.tptr_off;
                const Ent* tents = ents + R.tent_off;
                const int32_t* lossF = reinterpret_cast<const int32_t*>(tptr + C + 1);
                const int32_t* lossG = lossF + C;
                for (int j = 0; j < 2; j++) {
                    const int ch = j == 0 ? m->child0[e] : m->child1[e];
                    const int Cc = (int)recs[ch].C;
                    std::vector<std::vector<Ent>> lists(Cc);
                    std::vector<uint32_t> up(Cc, 0xffffffffu);
                    for (int c = 0; c < C; c++) {
                        const int32_t lc = j == 0 ? lossF[c] : lossG[c];
                        if (lc >= 0) up[lc] = (uint32_t)c;
                        for (uint32_t t = tptr[c]; t < tptr[c + 1]; t++) {
                            const Ent& en = tents[t];  // i1 = first child's cell, i2 = second child's cell
                            if (j == 0) lists[en.i1].push_back(Ent{(uint16_t)c, en.i2, 0u, en.p});
                            else lists[en.i2].push_back(Ent{(uint16_t)c, en.i1, 0u, en.p});
                        }
                    }
                    for (int c = 0; c < Cc; c++) if (up[c] == 0xffffffffu) return false;  // (a child's clade is always the parent's)
                    pad4(wv);
                    const uint32_t woff = (uint32_t)wv.size(), eoff = (uint32_t)ev.size();
                    for (int c = 0; c < Cc; c++) {
                        wv.push_back((uint32_t)ev.size() - eoff);
                        ev.insert(ev.end(), lists[c].begin(), lists[c].end());
                    }
                    wv.push_back((uint32_t)ev.size() - eoff);
                    wv.insert(wv.end(), up.begin(), up.end());
                    if (j == 0) { Q.sF_off = woff; Q.sFent_off = eoff; Q.nsFent = (uint32_t)ev.size() - eoff; }
                    else { Q.sG_off = woff; Q.sGent_off = eoff; Q.nsGent = (uint32_t)ev.size() - eoff; }
                }
            }
            if (kind != WHALE_ROOT) {
                if (hoff > 0xffffffffull) return false;
                Q.hoff = (uint32_t)hoff;
                hoff += (uint64_t)(m->nsl[e] + 1) * ((C + 1) & ~1);
            }
        }
        pad4(wv);
        if (hoff > 0xffffffffull) return false;
        const size_t words_at = (size_t)nn * sizeof(RevRec);
        const size_t ents_at = words_at + wv.size() * 4;
        const size_t total = ents_at + ev.size() * sizeof(Ent);
        if (total > 0xffffffffull) return false;
        blob.assign(total, 0);
        for (int e = 0; e < nn; e++) {
            RevRec& Q = rr[e];
            Q.bptr_off += (uint32_t)(words_at / 4); Q.bslot_off += (uint32_t)(words_at / 4);
            Q.sF_off += (uint32_t)(words_at / 4); Q.sG_off += (uint32_t)(words_at / 4);
            Q.bent_off += (uint32_t)(ents_at / 16); Q.sFent_off += (uint32_t)(ents_at / 16); Q.sGent_off += (uint32_t)(ents_at / 16);
        }
        memcpy(blob.data(), rr.data(), words_at);
        if (!wv.empty()) memcpy(blob.data() + words_at, wv.data(), wv.size() * 4);
        if (!ev.empty()) memcpy(blob.data() + ents_at, ev.data(), ev.size() * sizeof(Ent));
        hist_len = (uint32_t)hoff;
    }
    return true;
}

static bool build_reverse(whale_data* D) {
    const whale_model* m = D->m;
    const int nn = m->nn, F = D->F;
    const std::vector<unsigned char>& A = D->arena_host;
    std::vector<unsigned char>& RA = D->rarena_host;
    RA.clear();
    D->rhdr.assign(F, RevHdr{});
    std::vector<std::vector<unsigned char>> blobs(F);
    int nthr = env_int("WHALE_PACK_THREADS", (int)std::thread::hardware_concurrency());
    nthr = std::max(1, std::min(nthr, std::min(64, (F + 31) / 32)));
    std::atomic<int> next{0}, bad{0};
    auto worker = [&]() {
        for (;;) {
            const int f0 = next.fetch_add(16);
            if (f0 >= F || bad.load()) break;
            for (int f = f0; f < std::min(F, f0 + 16); f++)
                if (!reverse_family(m, A.data() + D->hdr[f].base, nn, blobs[f], D->rhdr[f].hist_len)) { bad.store(1); return; }
        }
    };
    if (nthr == 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthr; i++) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    if (bad.load()) return false;
    size_t bytes = 0;
    for (int f = 0; f < F; f++) {
        D->rhdr[f].base = bytes;
        D->rhdr[f].blob_bytes = (uint32_t)blobs[f].size();
        bytes += (blobs[f].size() + 15) & ~size_t(15);
    }
    RA.assign(bytes, 0);
    for (int f = 0; f < F; f++) {
        if (!blobs[f].empty()) memcpy(RA.data() + D->rhdr[f].base, blobs[f].data(), blobs[f].size());
        std::vector<unsigned char>().swap(blobs[f]);
    }
    return true;
}

static size_t smem_need_rev(const whale_model* m, const FamHdr& h, const RevHdr& r) {  // mirrors the carve-up in k_dp_rev
    const size_t nn = m->nn, NW = rev_nt() / 32;
    const size_t hdr = ((((9 * nn + 4) * sizeof(int)) + 15) & ~size_t(15)) + nn * (sizeof(NodeRec) + sizeof(RevRec)) +
                       (8 * nn + NW * 8 + 8 + 2) * sizeof(double);
    const size_t leaf = NW * ((size_t)r.leafmax * sizeof(double) + h.leaf_stage);
    const size_t back = ((size_t)r.arows_len + 3 * (size_t)r.hbuf_len) * sizeof(double);
    return hdr + ((size_t)r.rows_len + r.scr_len) * sizeof(double) + r.stage_bytes + r.stage2_bytes + std::max(leaf, back);
}

// shared-memory budgets of the reverse-mode kernel (hybrid plan), row and adjoint-row placement; returns the largest need
static size_t set_budgets_rev(whale_data* D) {
    const whale_model* m = D->m;
    const Plan& pl = m->planR;
    const int nn = m->nn;
    std::atomic<size_t> worst_a{0}, hist_a{0};
    D->roff_host[1].assign((size_t)D->F * nn, 0);
    D->aoff_host.assign((size_t)D->F * nn, 0);
    auto even = [](uint32_t v) { return (v + 1) & ~1u; };
    auto amax = [](std::atomic<size_t>& a, size_t v) { size_t cur = a.load(); while (v > cur && !a.compare_exchange_weak(cur, v)) {} };
    parallel_for(D->F, 256, [&](int f_lo, int f_hi) {
    size_t worst = 0, hist_max = 0;
    for (int f = f_lo; f < f_hi; f++) {
        RevHdr& H = D->rhdr[f];
        const std::vector<uint32_t>& Cs = D->famC[f];
        const NodeRec* recs = reinterpret_cast<const NodeRec*>(D->arena_host.data() + D->hdr[f].base);
        const RevRec* rrs = reinterpret_cast<const RevRec*>(D->rarena_host.data() + H.base);
        uint32_t mxinner = 0, mxleaf = 0, hbuf = 0;
        size_t stg = 0, stg2 = 0, stg_root = 0;
        auto p4 = [](size_t w) { return (w + 3) & ~size_t(3); };
        for (int e = 0; e < nn; e++) {
            if (m->kind[e] != WHALE_LEAF) {  // row-1 / root lists, forward and transposed (the segments of fwd_segs / bwd_segs)
                const size_t C = Cs[e], CF = Cs[m->child0[e]], CG = m->child1[e] >= 0 ? Cs[m->child1[e]] : 0;
                const int kd = m->kind[e];
                size_t fw = 0, bw = 0;
                if (kd != WHALE_INTERNAL) { fw += p4(C + 1) * 4 + (size_t)recs[e].ndent * 16; bw += p4(C + 1) * 4 + (size_t)rrs[e].nbent * 16; }
                if (kd != WHALE_WGD) {
                    fw += p4(3 * C + 1 + (kd == WHALE_ROOT ? D->hdr[f].nlev + 1 : 0)) * 4 + (size_t)recs[e].ntent * 16;
                    bw += p4(2 * CF + 1) * 4 + p4(2 * CG + 1) * 4 + ((size_t)rrs[e].nsFent + rrs[e].nsGent) * 16;
                }
                if (kd == WHALE_ROOT) stg_root = std::max(fw, bw);
                else stg2 = std::max(stg2, std::max(fw, bw));
            }
            const uint32_t K = (uint32_t)pl.K[e], ck = Cs[e] * (uint32_t)RS(K);
            const size_t n1 = (size_t)m->nsl[e] + 1;
            if (m->kind[e] == WHALE_LEAF) {
                if (recs[e].nslots > HEAVY_SLOTS) {
                    mxinner = std::max(mxinner, ck);
                    stg = std::max(stg, (size_t)recs[e].ndent + (recs[e].nslots + 1) / 2 + n1 * K);
                } else {
                    mxleaf = std::max(mxleaf, ck);
                }
            } else if (m->kind[e] != WHALE_ROOT) {
                mxinner = std::max(mxinner, even(Cs[e]));
                hbuf = std::max(hbuf, even(Cs[e]));
                stg = std::max(stg, (size_t)recs[e].ndent + (recs[e].nslots + 1) / 2 + n1);            // forward lists + ϕ/ψ rows
                stg = std::max(stg, (size_t)rrs[e].nbent + (rrs[e].nbslots + 1) / 2 + 4 * n1);         // backward lists + local rows
            }
        }
        std::vector<int> roff(nn + 1), aoff(nn + 1, 0);
        const int rows = place_rows(nn, (int)m->leafnodes.size(), m->leafnodes.data(), (int)m->inner.size(), m->inner.data(),
                                    m->child0.data(), m->child1.data(), m->kind.data(),
                                    [&](int e2) { return (int)even(Cs[e2] * (uint32_t)RS(pl.K[e2])); }, roff.data(), true);
        const int arows = place_arows((int)m->inner.size(), m->inner.data(), m->child0.data(), m->child1.data(), m->kind.data(),
                                      [&](int e2) { return (int)Cs[e2]; }, aoff.data());
        for (int e = 0; e < nn; e++) {
            D->roff_host[1][(size_t)f * nn + e] = (uint32_t)roff[e];
            D->aoff_host[(size_t)f * nn + e] = (uint32_t)aoff[e];
        }
        H.rows_len = even((uint32_t)rows);
        H.scr_len = even(mxinner);
        H.leafmax = even(mxleaf);
        H.arows_len = even((uint32_t)arows);
        H.hbuf_len = hbuf;
        const size_t STAGE_MAX = (size_t)env_int("WHALE_STAGE_MAX", D->F <= 160 ? 160 * 1024 : 24 * 1024);
        H.stage_bytes = 16 * stg > STAGE_MAX ? 0u : (uint32_t)(16 * stg);
        // The row-1 lists of the next internal/WGD node are prefetched into `stage2` while the current node's slices run; the
        // root (no slices) uses `stage` + `stage2` together, so stage2 is sized for max(one non-root node, root − stage):
        // 20–30 KB per C2 family, four families still resident per SM.  It shortens a family's critical path (−15 % cycles:
        // root 109 k -> 44 k) — +14 % evals/s at 1000 families per GPU, +16 % at 300 — but on a full GPU (12 500 families)
        // the latency it removes is hidden by the other resident families anyway and the extra shared-memory traffic
        // costs 3.5 % (profiles/r2_rev_stage2_sweep.txt): on for batches of up to a few waves of CTAs.
        const size_t STAGE2_MAX = (size_t)env_int("WHALE_STAGE2_MAX", D->F <= 160 ? 100 * 1024 : D->F <= 2500 ? 32 * 1024 : 0);
        const size_t with_root = std::max(stg2, stg_root > H.stage_bytes ? stg_root - H.stage_bytes : (size_t)0);
        if (H.stage_bytes == 0) { H.stage2_bytes = 0; H.root_staged = 0; }
        else if (with_root <= STAGE2_MAX) { H.stage2_bytes = (uint32_t)std::max<size_t>(with_root, 16); H.root_staged = 1; }
        else if (stg2 <= STAGE2_MAX) { H.stage2_bytes = (uint32_t)std::max<size_t>(stg2, 16); H.root_staged = 0; }
        else { H.stage2_bytes = 0; H.root_staged = 0; }
        hist_max = std::max(hist_max, (size_t)H.hist_len);
        worst = std::max(worst, smem_need_rev(m, D->hdr[f], H));
        if (env_int("WHALE_DEBUG", 0) >= 2)
            fprintf(stderr, "[whale] fam %d reverse: rows %u scr %u leafmax %u stage %u arows %u hbuf %u hist %u -> %zu B\n", f,
                    H.rows_len, H.scr_len, H.leafmax, H.stage_bytes, H.arows_len, H.hbuf_len, H.hist_len, smem_need_rev(m, D->hdr[f], H));
    }
    amax(worst_a, worst);
    amax(hist_a, hist_max);
    });
    D->hist_stride = (hist_a.load() + 1) & ~size_t(1);
    // the few families just above the rev_minb()-per-SM line read their lists in place (see set_budgets)
    const size_t line = occupancy_line(rev_minb());
    if (worst_a.load() > line) {
        int over = 0;
        for (int f = 0; f < D->F; f++) over += smem_need_rev(m, D->hdr[f], D->rhdr[f]) > line;
        if (over * 4 <= D->F) {
            size_t worst = 0;
            for (int f = 0; f < D->F; f++) {
                RevHdr& H = D->rhdr[f];
                const size_t nd = smem_need_rev(m, D->hdr[f], H);
                if (nd > line && nd - H.stage2_bytes <= line) { H.stage2_bytes = 0; H.root_staged = 0; }
                else if (nd > line && nd - H.stage2_bytes - H.stage_bytes <= line) { H.stage_bytes = 0; H.stage2_bytes = 0; H.root_staged = 0; }
                worst = std::max(worst, smem_need_rev(m, D->hdr[f], H));
            }
            worst_a.store(worst);
        }
    }
    return worst_a.load();
}

// Second half of whale_data_create, shared with whale_data_load: tangent plans and shared-memory budgets for this
// model/plan (they depend on the runtime configuration, not on the CCDs), launch order, upload, calibration.
// Takes ownership of D (deleted on failure).
static int32_t finalize_data(whale_model* m, whale_data* D, whale_data_t* out) {
    const int nn = m->nn, F = D->F;
    const double tf0 = now_s();
    const bool dbg = env_int("WHALE_DEBUG", 0) >= 1;
    std::vector<unsigned char>& A = D->arena_host;
    // ---- tangent plans: one gradient pass if every family's working set fits, else parameter chunks ----
    D->plans = {&m->plan[0], &m->plan[1]};
    set_budgets(D, 0, m->plan[0]);
    const size_t SMEM_MAX = 227 * 1024;
    const size_t SMEM_GOAL = (size_t)env_int("WHALE_SMEM_GOAL", 80000);  // default: at least two families per SM
    // Gradient mode: reverse (adjoint) DP by default — one pass whatever P is; WHALE_GRAD_MODE=fwd keeps the forward
    // tangents of k_dp (also the fallback when a family does not fit the reverse kernel's working set)
    // Default (WHALE_GRAD_MODE unset or "auto"): reverse mode, except for the small-P / small-family corner where one
    // forward pass is faster because a slice is a latency chain whose length does not depend on K — measured on the B200
    // (profiles/r2_modes_*.json): C2 (P = 5, ~200 clades) 3.3 M evals/s forward vs 2.1 M reverse; C3 (P = 37) 0.60 M
    // forward (6 passes) vs 2.5 M reverse; C4 (P = 8, ~2000 clades) 1.1e4 forward vs 2.9e4 reverse.
    {
        const char* gm = getenv("WHALE_GRAD_MODE");
        bool want_rev = !(gm && !strcmp(gm, "fwd"));
        if (want_rev && !(gm && !strcmp(gm, "rev"))) {
            double clades = 0.0;
            for (int f = 0; f < F; f++) clades += D->hdr[f].G;
            const size_t need_fwd = set_budgets(D, 1, m->plan[1]);
            if (m->plan[1].Kmax <= 6 && need_fwd <= SMEM_GOAL && clades / F <= 600.0) want_rev = false;
        }
        if (want_rev && tables_smem(m, m->plan[1], false) <= SMEM_MAX && build_reverse(D)) {
            const size_t need_rev = set_budgets_rev(D);
            D->rev = need_rev <= SMEM_MAX;
            if (env_int("WHALE_DEBUG", 0) >= 1)
                fprintf(stderr, "[whale] %d families, P = %d: reverse-mode gradient, worst family %zu B of shared memory%s\n", F, m->P,
                        need_rev, D->rev ? "" : " (does not fit: forward tangents)");
        }
        if (!D->rev) { std::vector<unsigned char>().swap(D->rarena_host); D->rhdr.clear(); }
    }
    size_t need1 = D->rev ? 0 : set_budgets(D, 1, m->plan[1]);
    if (!D->rev && (need1 > SMEM_GOAL || m->plan[1].Kmax > dp_nt())) {
        // parameters ordered by the node that owns them (subtrees stay together -> sparse chunks)
        std::vector<int> porder;
        std::vector<char> seen(m->P, 0);
        auto addp = [&](int sl) { if (sl >= 0 && !seen[sl]) { seen[sl] = 1; porder.push_back(sl); } };
        for (int oi = 0; oi < nn; oi++) {
            int e = m->order[oi];
            if (m->kind[e] != WHALE_ROOT) { addp(m->lam_slot[e]); addp(m->mu_slot[e]); }
            if (m->kind[e] == WHALE_WGD) addp(m->q_slot[e]);
        }
        addp(m->eta_slot);
        for (int p = 0; p < m->P; p++) addp(p);  // parameters no node reads (e.g. the root's rates): zero gradient
        // smallest chunk count (from a geometric ladder) whose worst family meets the goal; the last rung is taken
        // as long as it fits the hardware limit at all
        std::vector<int> ladder;
        const int top = std::max(2, std::min(MAXPLAN - 1, m->P));
        for (int v = 2; v < top; v = std::max(v + 1, v * 3 / 2)) ladder.push_back(v);
        ladder.push_back(top);
        bool ok = false;
        for (size_t li = 0; li < ladder.size() && !ok; li++) {
            const int nch = ladder[li];
            const bool last = li + 1 == ladder.size();
            for (Plan& cp : D->chunk_plans) for (void* q : cp.owned) cudaFree(q);
            D->chunk_plans.assign(nch, Plan());
            D->plans.assign(1, &m->plan[0]);
            size_t worst = 0;
            int kmax = 0;
            for (int c = 0; c < nch; c++) {
                std::vector<char> sub(m->P, 0);
                const size_t lo = porder.size() * c / nch, hi = porder.size() * (c + 1) / nch;
                for (size_t i = lo; i < hi; i++) sub[porder[i]] = 1;
                build_plan(*m, sub, D->chunk_plans[c]);
                worst = std::max(worst, set_budgets(D, 1 + c, D->chunk_plans[c]));
                kmax = std::max(kmax, D->chunk_plans[c].Kmax);
            }
            ok = (worst <= SMEM_GOAL || (last && worst <= SMEM_MAX)) && kmax <= dp_nt();
            if (ok || last) {
                for (int c = 0; c < nch; c++) {
                    CU(upload_plan(D->chunk_plans[c], nn));
                    D->plans.push_back(&D->chunk_plans[c]);
                }
                need1 = worst;
            }
        }
    }
    if (env_int("WHALE_DEBUG", 0) >= 1 && !D->rev)
        fprintf(stderr, "[whale] %d families, P = %d: %zu gradient pass(es), worst family %zu B of shared memory\n", F, m->P,
                D->plans.size() - 1, need1);
    if (need1 > SMEM_MAX) { delete D; return fail(WHALE_ERR_CAPACITY, "a family needs %zu bytes of shared memory (> 227 KB) even with %d parameter chunks", need1, MAXPLAN - 1); }
    if (dbg) fprintf(stderr, "[whale] plans, reverse lists and budgets: %.2f s\n", now_s() - tf0);
    size_t outsz = 0;
    for (size_t g = 0; g < D->plans.size(); g++) {
        const Plan& pl = *D->plans[g];
        D->out_off.push_back(g == 0 ? 0 : outsz);  // the value plan shares the first gradient slot's space
        if (g >= 1) outsz += (size_t)F * pl.K[m->root];
        CU(order_families(D, (int)g, D->work));
        CU(upload(D->roff_host[g], &D->d_roff[g]));
    }
    D->out_total = std::max<size_t>(outsz, (size_t)F);
    CU(cudaMalloc((void**)&D->d_arena, std::max<size_t>(A.size(), 16)));
    CU(cudaMemcpy(D->d_arena, A.data(), A.size(), cudaMemcpyHostToDevice));
    D->arena_bytes = A.size();
    std::vector<unsigned char>().swap(A);
    if (D->rev) {
        CU(cudaMalloc((void**)&D->d_rarena, std::max<size_t>(D->rarena_host.size(), 16)));
        CU(cudaMemcpy(D->d_rarena, D->rarena_host.data(), D->rarena_host.size(), cudaMemcpyHostToDevice));
        D->rarena_bytes = D->rarena_host.size();
        std::vector<unsigned char>().swap(D->rarena_host);
        CU(upload(D->rhdr, &D->d_rhdr));
        CU(upload(D->aoff_host, &D->d_aoff));
        CU(cudaMalloc((void**)&D->d_next, MAX_BINS * sizeof(unsigned int)));
        CU(cudaMemset(D->d_next, 0, MAX_BINS * sizeof(unsigned int)));
    }
    CU(upload(D->hdr, &D->d_hdr));
    CU(cudaMalloc((void**)&D->d_out_fam, D->out_total * sizeof(double)));
    CU(cudaMalloc((void**)&D->d_partial, (size_t)1024 * m->plan[1].Kmax * sizeof(double)));
    CU(cudaMalloc((void**)&D->d_done, 16));
    CU(cudaMemset(D->d_done, 0, 16));
    for (int i = 0; i < MAX_BINS; i++) {
        CU(cudaStreamCreateWithFlags(&D->side[i], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&D->ev_join[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&D->ev_fork, cudaEventDisableTiming));
    *out = D;
    if (dbg) fprintf(stderr, "[whale] ... + launch order and upload: %.2f s\n", now_s() - tf0);
#ifndef WHALE_EMU
    // Calibration pass: one profiled evaluation at a benign parameter point; the measured SM cycles of every family
    // replace the packer's flop estimate as the scheduling weight (the estimate misses e.g. long leaf-branch chains).
    // Results never depend on the launch order: per-family outputs are indexed by family, the sum runs in family order.
    if (F > 1 && env_int("WHALE_CALIBRATE", 1)) {
        std::vector<double> x(std::max(m->P, 1), 0.0);
        for (int e = 0; e < nn; e++) {
            if (m->lam_slot[e] >= 0) x[m->lam_slot[e]] = m->log_scale ? log(0.2) : 0.2;
            if (m->mu_slot[e] >= 0) x[m->mu_slot[e]] = m->log_scale ? log(0.3) : 0.3;
            if (m->q_slot[e] >= 0) x[m->q_slot[e]] = 0.2;
        }
        if (m->eta_slot >= 0) x[m->eta_slot] = 0.67;
        CU(cudaMemcpy(m->d_x, x.data(), m->P * sizeof(double), cudaMemcpyHostToDevice));
        CU(cudaMemset(m->d_pleaf, 0, nn * sizeof(double)));
        m->x_host_valid = false;
        for (int rep = 0; rep < 2; rep++) {  // the first run pays cold caches and lazy allocations
            int32_t rc = enqueue_eval(m, D, m->d_x, 0, WHALE_WANT_GRAD | WHALE_PROFILE, m->d_out, m->stream);
            if (rc != WHALE_OK) return rc;
            CU(cudaStreamSynchronize(m->stream));
        }
        std::vector<long long> tim((size_t)F * TIMW);
        CU(cudaMemcpy(tim.data(), D->d_tim, tim.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        std::vector<double> cyc(F);
        bool ok = true;
        for (int f = 0; f < F; f++) { cyc[f] = (double)tim[(size_t)f * TIMW + 6]; ok = ok && cyc[f] > 0.0; }
        if (ok) {
            D->work_measured = cyc;
            for (size_t g = 0; g < D->plans.size(); g++) CU(order_families(D, (int)g, cyc));
        }
        D->ev_valid = false;
    }
#endif
    return WHALE_OK;
}

int32_t whale_read_ale(whale_model_t m, int32_t n_files, const char* const* paths, int32_t n_species,
                       const char* const* species_names, const int32_t* species_ids, const int64_t* node_clade_off,
                       const int32_t* node_clade, int32_t n_threads, whale_data_t* out, int32_t* n_clades) {
    if (!m || !paths || !species_names || !species_ids || !node_clade_off || !node_clade || !out || n_files <= 0)
        return fail(WHALE_ERR_ARG, "null argument");
    const int nn = m->nn;
    whale_ale::Species sp;
    for (int i = 0; i < n_species; i++) {
        if (species_ids[i] < 0) return fail(WHALE_ERR_ARG, "negative species id");
        sp.id_of[species_names[i]] = species_ids[i];
        sp.n_ids = std::max(sp.n_ids, (int)species_ids[i]);
    }
    for (int64_t i = 0; i < node_clade_off[nn]; i++) sp.n_ids = std::max(sp.n_ids, (int)node_clade[i]);
    sp.words = (sp.n_ids + 64) / 64;
    sp.node_mask.assign(nn, std::vector<uint64_t>(sp.words, 0));
    for (int e = 0; e < nn; e++)
        for (int64_t i = node_clade_off[e]; i < node_clade_off[e + 1]; i++) {
            if (node_clade[i] < 0) return fail(WHALE_ERR_ARG, "negative species id");
            sp.node_mask[e][node_clade[i] >> 6] |= 1ull << (node_clade[i] & 63);
        }
    std::vector<std::string> files(paths, paths + n_files);
    std::vector<whale_ale::Family> fams;
    const double tp0 = now_s();
    whale_ale::parse_all(files, sp, nn, fams, n_threads);
    if (env_int("WHALE_DEBUG", 0) >= 1) fprintf(stderr, "[whale] parsed %d .ale files in %.2f s\n", n_files, now_s() - tp0);
    for (int f = 0; f < n_files; f++)
        if (!fams[f].error.empty()) return fail(WHALE_ERR_ARG, "%s", fams[f].error.c_str());
    // concatenate into the flattened reference layout and pack
    std::vector<int64_t> clade_off(1, 0), split_off(1, 0), compat_off(1, 0);
    std::vector<int32_t> nleaf, g1, g2, compat;
    std::vector<double> p;
    for (int f = 0; f < n_files; f++) {
        const whale_ale::Family& F = fams[f];
        const int64_t sbase = (int64_t)g1.size(), cbase = (int64_t)compat.size();
        clade_off.push_back(clade_off.back() + (int64_t)F.nleaf.size());
        nleaf.insert(nleaf.end(), F.nleaf.begin(), F.nleaf.end());
        for (size_t i = 1; i < F.split_off.size(); i++) split_off.push_back(sbase + F.split_off[i]);
        g1.insert(g1.end(), F.g1.begin(), F.g1.end());
        g2.insert(g2.end(), F.g2.begin(), F.g2.end());
        p.insert(p.end(), F.p.begin(), F.p.end());
        for (size_t i = 1; i < F.compat_off.size(); i++) compat_off.push_back(cbase + F.compat_off[i]);
        compat.insert(compat.end(), F.compat.begin(), F.compat.end());
        if (n_clades) n_clades[f] = (int32_t)F.nleaf.size();
        fams[f] = whale_ale::Family();  // release as we go
    }
    whale_ccd_desc d{n_files, clade_off.data(), nleaf.data(), split_off.data(), g1.data(), g2.data(), p.data(),
                     compat_off.data(), compat.data()};
    return whale_data_create(m, &d, out);
}

// ---- binary arena cache: the packed state of a data handle, so a later process skips parsing the
//      .ale files and running the packer.  Layout: CacheHdr | FamHdr[F] | arena bytes | famC | per-node packer facts |
//      work | aggregates.  Plans, shared-memory budgets, launch order and calibration are rebuilt on load (they depend
//      on the runtime configuration).  A cache is valid for the model it was packed for (structure fingerprint). ----
constexpr uint32_t CACHE_VERSION = 2;   // file layout
constexpr uint32_t PACKER_VERSION = 2;  // meaning of the packed bytes (NodeRec / Ent / Slot / list order); bump on any packer change
struct CacheHdr {
    char magic[8];       // "WHALEAR1"
    uint32_t version, nn;
    uint64_t model_fp;   // species-tree structure + slicing + record sizes + packer version
    uint64_t F, arena_bytes, ell_total;
    int64_t algo_bytes;
    double aggG, aggTroot;
    uint64_t content_hash;  // FNV-1a over the family headers and the arena
};
static uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
// Structural validation of a loaded arena: every offset the kernels will follow must stay inside its family's blob
// (a truncated-then-padded, corrupted or stale-layout cache must be refused, not read out of bounds on the device).
static const char* validate_arena(const whale_model* m, const whale_data* D, const std::vector<uint32_t>& famC) {
    const size_t nn = (size_t)m->nn, F = (size_t)D->F, A = D->arena_host.size();
    uint64_t ell = 0;
    for (size_t f = 0; f < F; f++) {
        const FamHdr& H = D->hdr[f];
        if ((H.base & 15) || H.base > A || (uint64_t)H.blob_bytes > A - H.base) return "family blob outside the arena";
        if ((size_t)H.blob_bytes < nn * sizeof(NodeRec) || (H.blob_bytes & 15)) return "family blob too small";
        if (H.ell_off != ell) return "ℓ offsets inconsistent";
        const unsigned char* blob = D->arena_host.data() + H.base;
        const NodeRec* recs = reinterpret_cast<const NodeRec*>(blob);
        const uint64_t W = H.blob_bytes / 4, E = H.blob_bytes / 16;
        for (size_t e = 0; e < nn; e++) {
            const NodeRec& R = recs[e];
            const uint64_t C = R.C;
            if (C != famC[f * nn + e] || C > 65535 || R.nonleaf > C) return "compat counts inconsistent";
            const int kind = m->kind[e];
            uint64_t tw = (kind == WHALE_INTERNAL || kind == WHALE_ROOT) ? 3 * C + 1 : 0;
            if (kind == WHALE_ROOT) tw += (uint64_t)H.nlev + 1;
            tw = (tw + 3) & ~3ull;
            if ((R.dptr_off & 3) || (R.tptr_off & 3) || (R.slot_off & 3)) return "misaligned list";
            if (R.dptr_off + C + 1 > W || R.slot_off + 2ull * R.nslots > W || R.tptr_off + tw + C > W) return "word list outside the blob";
            if ((uint64_t)R.dent_off + R.ndent > E || (uint64_t)R.tent_off + R.ntent > E) return "entry list outside the blob";
            if (R.sptr_off && ((uint64_t)R.sptr_off + C + 1 > W || R.sent_off > E)) return "shape list outside the blob";
            const uint32_t* words = reinterpret_cast<const uint32_t*>(blob);
            if (words[R.dptr_off + C] != R.ndent) return "same-branch term count inconsistent";
            if (tw && words[R.tptr_off + C] != R.ntent) return "speciation term count inconsistent";
            if (R.sptr_off && (uint64_t)R.sent_off + words[R.sptr_off + C] > E) return "shape list outside the blob";
            const Ent* ents = reinterpret_cast<const Ent*>(blob);
            const uint64_t srcC = kind == WHALE_WGD ? famC[f * nn + m->child0[e]] : C;
            for (uint32_t t = 0; t < R.ndent; t++) if (ents[R.dent_off + t].i1 >= srcC || ents[R.dent_off + t].i2 >= srcC) return "term index out of range";
            if (tw) {
                const uint64_t CF = famC[f * nn + m->child0[e]], CG = famC[f * nn + m->child1[e]];
                for (uint32_t t = 0; t < R.ntent; t++) if (ents[R.tent_off + t].i1 >= CF || ents[R.tent_off + t].i2 >= CG) return "term index out of range";
                const int32_t* lossF = reinterpret_cast<const int32_t*>(words + R.tptr_off + C + 1);
                for (uint64_t c = 0; c < C; c++) if (lossF[c] >= (int64_t)CF || lossF[C + c] >= (int64_t)CG) return "loss index out of range";
            }
            const Slot* sl = reinterpret_cast<const Slot*>(words + R.slot_off);
            for (uint32_t i = 0; i < R.nslots; i++)
                if (sl[i].cell >= C || (uint64_t)sl[i].first + (uint64_t)(sl[i].cnt ? sl[i].cnt - 1 : 0) * sl[i].stride >= std::max<uint64_t>(R.ndent, 1) + (sl[i].cnt ? 0 : 65536))
                    return "lane table out of range";
            ell += (uint64_t)(m->nsl[e] + 1) * C;
        }
        if (recs[m->root].C != H.G) return "root compat list is not the clade list";
        {   // root levels: every Πroot term of a cell must point below the cell's level (the kernels rely on it)
            const NodeRec& R = recs[m->root];
            const uint32_t* words = reinterpret_cast<const uint32_t*>(blob);
            const Ent* ents = reinterpret_cast<const Ent*>(blob);
            const uint32_t* dptr = words + R.dptr_off;
            const uint32_t* lev = words + R.tptr_off + 3 * (size_t)R.C + 1;
            if (H.nlev == 0 || lev[0] != 0 || lev[H.nlev] != R.C) return "root levels inconsistent";
            for (uint32_t L = 0; L < H.nlev; L++) {
                if (lev[L] >= lev[L + 1]) return "root levels inconsistent";
                for (uint32_t c = lev[L]; c < lev[L + 1]; c++)
                    for (uint32_t t = dptr[c]; t < dptr[c + 1]; t++)
                        if (ents[R.dent_off + t].i1 >= lev[L] || ents[R.dent_off + t].i2 >= lev[L]) return "root term above its level";
            }
        }
    }
    if (ell != D->ell_total) return "ℓ size inconsistent";
    return nullptr;
}
static uint64_t model_fingerprint(const whale_model* m) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { h ^= v; h *= 1099511628211ull; };
    mix((uint64_t)m->nn);
    for (int e = 0; e < m->nn; e++) {
        mix((uint64_t)(uint32_t)m->kind[e]); mix((uint64_t)(uint32_t)m->nsl[e]);
        mix((uint64_t)(uint32_t)m->child0[e]); mix((uint64_t)(uint32_t)m->child1[e]); mix((uint64_t)(uint32_t)m->order[e]);
    }
    mix((uint64_t)sizeof(FamHdr)); mix((uint64_t)sizeof(NodeRec)); mix((uint64_t)HEAVY_SLOTS); mix((uint64_t)NSHAPE);
    mix((uint64_t)sizeof(Ent)); mix((uint64_t)sizeof(Slot)); mix((uint64_t)PACKER_VERSION);
    return h;
}
static bool put_raw(FILE* fp, const void* p, size_t bytes) { return bytes == 0 || fwrite(p, 1, bytes, fp) == bytes; }
static bool get_raw(FILE* fp, void* p, size_t bytes) { return bytes == 0 || fread(p, 1, bytes, fp) == bytes; }
#define put(fp, v) put_raw(fp, (v).data(), (v).size() * sizeof((v)[0]))
#define get(fp, v, n) ((v).resize(n), get_raw(fp, (v).data(), (size_t)(n) * sizeof((v)[0])))

int32_t whale_data_save(whale_data_t d, const char* path) {
    if (!d || !path) return fail(WHALE_ERR_ARG, "null argument");
    const whale_model* m = d->m;
    const size_t F = (size_t)d->F, nn = (size_t)m->nn;
    CU(cudaSetDevice(m->device));
    std::vector<unsigned char> arena(d->arena_bytes);
    if (d->arena_bytes) CU(cudaMemcpy(arena.data(), d->d_arena, d->arena_bytes, cudaMemcpyDeviceToHost));
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(WHALE_ERR_ARG, "cannot open %s for writing", path);
    CacheHdr h{};
    memcpy(h.magic, "WHALEAR1", 8);
    h.version = CACHE_VERSION; h.nn = (uint32_t)nn; h.model_fp = model_fingerprint(m);
    h.content_hash = fnv1a(arena.data(), arena.size(), fnv1a(d->hdr.data(), d->hdr.size() * sizeof(FamHdr)));
    h.F = F; h.arena_bytes = d->arena_bytes; h.ell_total = d->ell_total; h.algo_bytes = d->algo_bytes;
    h.aggG = d->aggG; h.aggTroot = d->aggTroot;
    std::vector<uint32_t> famC(F * nn);
    for (size_t f = 0; f < F; f++) for (size_t e = 0; e < nn; e++) famC[f * nn + e] = d->famC[f][e];
    bool ok = fwrite(&h, sizeof(h), 1, fp) == 1 && put(fp, d->hdr) && put(fp, arena) && put(fp, famC) && put(fp, d->f_ndent) &&
              put(fp, d->f_ntent) && put(fp, d->f_heavy) && put(fp, d->f_stage16) && put(fp, d->f_rootwin) && put(fp, d->work) &&
              put(fp, d->aggC) && put(fp, d->aggT);
    ok = (fclose(fp) == 0) && ok;
    return ok ? WHALE_OK : fail(WHALE_ERR_ARG, "short write to %s", path);
}

int32_t whale_data_load(whale_model_t m, const char* path, whale_data_t* out) {
    if (!m || !path || !out) return fail(WHALE_ERR_ARG, "null argument");
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(WHALE_ERR_ARG, "cannot open %s", path);
    CacheHdr h{};
    if (fread(&h, sizeof(h), 1, fp) != 1 || memcmp(h.magic, "WHALEAR1", 8) != 0 || h.version != CACHE_VERSION) {
        fclose(fp);
        return fail(WHALE_ERR_ARG, "%s is not a whalecuda arena cache (version %u)", path, CACHE_VERSION);
    }
    if (h.nn != (uint32_t)m->nn || h.model_fp != model_fingerprint(m) || h.F == 0 || h.F > 0x7fffffffull) {
        fclose(fp);
        return fail(WHALE_ERR_ARG, "%s was packed for another model (species tree / slicing differ)", path);
    }
    CU(cudaSetDevice(m->device));
    auto* D = new whale_data();
    D->m = m;
    D->F = (int)h.F;
    const size_t F = (size_t)h.F, nn = (size_t)m->nn;
    std::vector<uint32_t> famC;
    bool ok = get(fp, D->hdr, F) && get(fp, D->arena_host, (size_t)h.arena_bytes) && get(fp, famC, F * nn) &&
              get(fp, D->f_ndent, F * nn) && get(fp, D->f_ntent, F * nn) && get(fp, D->f_heavy, F * nn) &&
              get(fp, D->f_stage16, F * nn) && get(fp, D->f_rootwin, F) && get(fp, D->work, F) && get(fp, D->aggC, nn) &&
              get(fp, D->aggT, nn);
    ok = ok && fgetc(fp) == EOF;
    fclose(fp);
    if (!ok) { delete D; return fail(WHALE_ERR_ARG, "%s is truncated or has trailing bytes", path); }
    if (fnv1a(D->arena_host.data(), D->arena_host.size(), fnv1a(D->hdr.data(), D->hdr.size() * sizeof(FamHdr))) != h.content_hash) {
        delete D;
        return fail(WHALE_ERR_ARG, "%s is corrupted (content hash mismatch)", path);
    }
    D->ell_total = h.ell_total;
    if (const char* why = validate_arena(m, D, famC)) { delete D; return fail(WHALE_ERR_ARG, "%s is not a valid arena for this model: %s", path, why); }
    D->famC.resize(F);
    for (size_t f = 0; f < F; f++) D->famC[f].assign(famC.begin() + f * nn, famC.begin() + (f + 1) * nn);
    D->ell_total = h.ell_total;
    D->algo_bytes = h.algo_bytes;
    D->aggG = h.aggG;
    D->aggTroot = h.aggTroot;
    return finalize_data(m, D, out);
}
#undef put
#undef get

int32_t whale_data_destroy(whale_data_t d) {
    if (!d) return WHALE_OK;
    cudaSetDevice(d->m->device);
    cudaFree(d->d_arena); cudaFree(d->d_hdr);
    for (int g = 0; g < MAXPLAN; g++) { cudaFree(d->d_perm[g]); cudaFree(d->d_roff[g]); }
    for (Plan& cp : d->chunk_plans) for (void* q : cp.owned) cudaFree(q);
    cudaFree(d->d_out_fam); cudaFree(d->d_partial); cudaFree(d->d_done);
    {
        auto& T = d->tb;
        cudaFree(T.d_cnt); cudaFree(T.d_st); cudaFree(T.d_g); cudaFree(T.d_e); cudaFree(T.d_t); cudaFree(T.d_p); cudaFree(T.d_stack);
        cudaFree(T.d_u); cudaFree(T.d_xs); cudaFree(T.d_pairs); cudaFree(T.d_famlist); cudaFree(T.d_off); cudaFree(T.d_packed);
        cudaFree(T.d_hash); cudaFree(T.d_dh); cudaFree(T.d_dc); cudaFree(T.d_df); cudaFree(T.d_nd);
        if (T.h_cnt) cudaFreeHost(T.h_cnt);
        if (T.h_st) cudaFreeHost(T.h_st);
        if (T.h_off) cudaFreeHost(T.h_off);
        if (T.h_packed) cudaFreeHost(T.h_packed);
        for (Plan& sp : T.slot_plans) for (void* p : sp.owned) cudaFree(p);
        cudaFree(T.d_ell_slots); cudaFree((void*)T.d_slot_eps); cudaFree((void*)T.d_slot_pp); cudaFree(T.d_slot_off);
    }
#ifndef WHALE_EMU
    for (int q = 0; q < 16; q++) if (d->peer_open[q] && d->peer_bufs[q]) cudaIpcCloseMemHandle(d->peer_bufs[q]);
#endif
    if (d->peer_rank >= 0) cudaFree(d->peer_bufs[d->peer_rank]);
    cudaFree(d->d_peer);
    cudaFree(d->d_rarena); cudaFree(d->d_rhdr); cudaFree(d->d_aoff); cudaFree(d->d_hist); cudaFree(d->d_next);
    for (int i = 0; i < MAX_BINS; i++) {
        if (d->side[i]) cudaStreamDestroy(d->side[i]);
        if (d->ev_join[i]) cudaEventDestroy(d->ev_join[i]);
    }
#ifndef WHALE_EMU
    for (auto& g : d->graphs) if (g.state == 2 && g.exec) cudaGraphExecDestroy(g.exec);
#endif
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    cudaFree(d->d_ell); cudaFree(d->d_tim); cudaFree(d->d_x_keep); cudaFree(d->d_pleaf_keep);
    for (int i = 0; i < 4; i++) if (d->ev[i]) cudaEventDestroy(d->ev[i]);
    delete d;
    return WHALE_OK;
}

int32_t whale_data_nfam(whale_data_t d) { return d ? d->F : 0; }
int32_t whale_data_grad_mode(whale_data_t d) { return d ? (d->rev ? 1 : 0) : -1; }
int32_t whale_data_grad_passes(whale_data_t d) { return d ? (int32_t)d->plans.size() - 1 : -1; }
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

// NowhereExtinctCondition scratch of a plan: Σ_e 2^L_e · K_e doubles (L_e = leaves below node e), on first use
static int32_t ensure_nowhere(whale_model* m, Plan& pl) {
    if (pl.dev.nwvec) return WHALE_OK;
    const int nn = m->nn;
    std::vector<int> L(nn, 0);
    std::vector<long long> off(nn, 0);
    long long tot = 0;
    for (int oi = 0; oi < nn; oi++) {
        const int e = m->order[oi];
        L[e] = m->kind[e] == WHALE_LEAF ? 1 : (m->child0[e] >= 0 ? L[m->child0[e]] : 0) + (m->child1[e] >= 0 ? L[m->child1[e]] : 0);
        if (L[e] > 20) return fail(WHALE_ERR_CAPACITY, "NowhereExtinctCondition needs 2^%d terms (more than 20 leaves)", L[e]);
        off[e] = tot;
        tot += (1LL << L[e]) * pl.K[e];
    }
    int* dL = nullptr; long long* doff = nullptr; double* dv = nullptr;
    CU(upload(L, &dL));
    CU(upload(off, &doff));
    CU(cudaMalloc((void**)&dv, (size_t)tot * sizeof(double)));
    pl.owned.push_back(dL); pl.owned.push_back(doff); pl.owned.push_back(dv);
    pl.dev.nwL = dL; pl.dev.nwoff = doff; pl.dev.nwvec = dv;
    return WHALE_OK;
}

// reverse-mode gradient evaluation: [three table sets in one launch -> k_dp_rev (persistent CTAs) -> reduction in its tail]
static int32_t enqueue_eval_rev(whale_model* m, whale_data* D, const double* d_x, int32_t condition, uint32_t flags,
                                double* d_out, cudaStream_t st) {
    const int F = D->F, NT = rev_nt(), MB = rev_minb();
    const bool prof = (flags & WHALE_PROFILE) != 0;
    Plan& pg = m->plan[1];
    CU(cudaMemsetAsync(d_out, 0, (1 + m->P) * sizeof(double), st));
    CU(cudaMemsetAsync(D->d_next, 0, MAX_BINS * sizeof(unsigned int), st));
    if (prof) CU(cudaEventRecord(D->ev[0], st));
    CU(launch_tables3(m, d_x, m->d_pleaf, st));
    if (condition == WHALE_COND_NOWHERE) {
        int32_t rcn = ensure_nowhere(m, pg);
        if (rcn != WHALE_OK) return rcn;
        LAUNCH(k_nowhere, 1, 256, 0, st, m->dev, pg.dev, d_x);
        g_launches++;
    }
    if (prof) CU(cudaEventRecord(D->ev[1], st));
    const int KR = pg.K[m->root];
    double* out_fam = D->d_out_fam + D->out_off[1];
    const bool fused = fused_reduce() && (size_t)F * KR <= ((size_t)1 << 20) && KR <= 1024;
    const std::vector<Bin>& bins = D->bins[1];
    size_t tail_smem = (size_t)(NT + 2 * KR + 2) * sizeof(double);
    PeerDev* peer = nullptr;  // the sum over ranks rides in the tail of the same kernel (one pass, fused reduction)
    if ((flags & WHALE_PEER_SUM) && fused && peer_fusion()) {
        int32_t rcp = peer_state(D, &peer);
        if (rcp != WHALE_OK) return rcp;
        tail_smem = std::max(tail_smem, peer_smem_bytes(D->peer_world, 1 + m->P));
        D->peer_fused = true;
    }
    auto launch_bin = [&](size_t b, cudaStream_t s) {
        RevArgs a{m->dev, m->planR.dev, pg.dev, m->planL.dev, D->d_arena, D->d_hdr, D->d_rarena, D->d_rhdr, D->d_perm[1],
                  D->d_roff[1], D->d_aoff, out_fam, D->d_hist, (unsigned long long)D->hist_stride, D->d_next + b,
                  bins[b].off, bins[b].count, D->rev_slot0[b], prof ? D->d_tim : nullptr, fused ? D->d_done : nullptr, F,
                  condition, d_out, peer};
        const bool pdl = pdl_enabled() && bins.size() == 1;
#define LAUNCHV(NTV, MBV)                                                                                                   \
    if (NT == NTV && MB == MBV) {                                                                                           \
        if (pdl) LAUNCH_PDL((k_dp_rev<NTV, MBV>), D->rev_grid[b], NTV, std::max(bins[b].smem, tail_smem), s, a);            \
        else LAUNCH((k_dp_rev<NTV, MBV>), D->rev_grid[b], NTV, std::max(bins[b].smem, tail_smem), s, a);                    \
    }
        REV_VARIANTS(LAUNCHV)
#undef LAUNCHV
        g_launches++;
    };
    if (bins.size() == 1) {
        launch_bin(0, st);
    } else {
        CU(cudaEventRecord(D->ev_fork, st));
        for (size_t b = 0; b < bins.size(); b++) {
            CU(cudaStreamWaitEvent(D->side[b], D->ev_fork, 0));
            launch_bin(b, D->side[b]);
            CU(cudaEventRecord(D->ev_join[b], D->side[b]));
            CU(cudaStreamWaitEvent(st, D->ev_join[b], 0));
        }
    }
    if (prof) CU(cudaEventRecord(D->ev[2], st));
    if (!fused) {
        const int nb = std::min(1024, (F + 255) / 256), chunk = (F + nb - 1) / nb;
        LAUNCH(k_reduce1, nb, 256, 0, st, out_fam, F, KR, chunk, D->d_partial);
        LAUNCH(k_reduce2, 1, 256, 0, st, D->d_partial, nb, KR, F, condition, pg.dev, m->root, 1, d_out);
        g_launches += 2;
    }
    if (prof) CU(cudaEventRecord(D->ev[3], st));
    D->ev_valid = prof;
    CU(cudaGetLastError());
    return WHALE_OK;
}

// enqueue [tables -> DP -> reduction] for every tangent plan of this evaluation on `st`; result in d_out
static int32_t enqueue_eval(whale_model* m, whale_data* D, const double* d_x, int32_t condition, uint32_t flags,
                            double* d_out, cudaStream_t st) {
    if (condition < 0 || condition > 3) return fail(WHALE_ERR_ARG, "unknown condition kind %d", condition);
    const bool grad = (flags & WHALE_WANT_GRAD) != 0;
    const int F = D->F;
    const bool keep = (flags & WHALE_KEEP_ELL) != 0;
    if (keep && !D->d_ell) CU(cudaMalloc((void**)&D->d_ell, std::max<uint64_t>(D->ell_total, 1) * sizeof(double)));
    if (keep) {
        if (!D->d_x_keep) {
            CU(cudaMalloc((void**)&D->d_x_keep, std::max(1, m->P) * sizeof(double)));
            CU(cudaMalloc((void**)&D->d_pleaf_keep, m->nn * sizeof(double)));
        }
        CU(cudaMemcpyAsync(D->d_x_keep, d_x, m->P * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(D->d_pleaf_keep, m->d_pleaf, m->nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    const bool prof = (flags & WHALE_PROFILE) != 0;
    if (prof && !D->ev[0]) for (int i = 0; i < 4; i++) CU(cudaEventCreate(&D->ev[i]));
    if (prof && !D->d_tim) {
        CU(cudaMalloc((void**)&D->d_tim, (size_t)F * TIMW * sizeof(long long)));
        CU(cudaMemset(D->d_tim, 0, (size_t)F * TIMW * sizeof(long long)));
    }
    if (!m->attr_set) {  // function attributes are per device: once per model (a model lives on one device)
#define SETATTR(NTV, MB) CU(cudaFuncSetAttribute(k_dp<NTV, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DP_VARIANTS(SETATTR)
#undef SETATTR
#define SETATTR(NTV, MB) CU(cudaFuncSetAttribute(k_dp_rev<NTV, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        REV_VARIANTS(SETATTR)
#undef SETATTR
        m->attr_set = true;
    }
    const int NT = dp_nt();
    if (grad && D->rev) {
        if (keep) {  // logpdf! + gradient: ℓ comes from the value-only forward kernel, the gradient from the reverse pass
            int32_t rck = enqueue_eval(m, D, d_x, condition, flags & ~(WHALE_WANT_GRAD | WHALE_PEER_SUM), d_out, st);
            if (rck != WHALE_OK) return rck;
        }
        return enqueue_eval_rev(m, D, d_x, condition, flags, d_out, st);
    }
    const size_t g0 = grad ? 1 : 0, g1 = grad ? D->plans.size() : 1;
    CU(cudaMemsetAsync(d_out, 0, (1 + m->P) * sizeof(double), st));
    for (size_t g = g0; g < g1; g++) {
        Plan& pl = *D->plans[g];
        const bool first = g == g0;
        if (prof && first) CU(cudaEventRecord(D->ev[0], st));
        // K1: slice tables of this plan (+ the leaf-branch shape tables, extra CTAs of the same launch)
        CU(launch_tables(m, pl, d_x, m->d_pleaf, st, !keep));
        if (condition == WHALE_COND_NOWHERE) {
            int32_t rcn = ensure_nowhere(m, pl);
            if (rcn != WHALE_OK) return rcn;
            LAUNCH(k_nowhere, 1, 256, 0, st, m->dev, pl.dev, d_x);
            g_launches++;
        }
        if (prof && first) CU(cudaEventRecord(D->ev[1], st));
        // K2: one launch per shared-memory bin, concurrently on side streams
        const std::vector<Bin>& bins = D->bins[g];
        double* out_fam = D->d_out_fam + D->out_off[g];
        // the ℓ kept for backtracking is written by the first pass only (values do not depend on the chunk)
        DPArgs a{m->dev, pl.dev, D->d_arena, D->d_hdr, D->d_perm[g], D->d_roff[g], out_fam, (keep && first) ? D->d_ell : nullptr, (int)g,
                 keep ? 0 : 1, (prof && first) ? D->d_tim : nullptr, nullptr, F, condition, first ? 1 : 0, d_out, nullptr};
        // K3 rides in the tail of K2 (the last CTA to finish reduces) unless the sum is too long for one CTA
        const bool fused = fused_reduce() && (size_t)F * pl.K[m->root] <= ((size_t)1 << 20) && pl.K[m->root] <= 1024;
        if (fused) a.done = D->d_done;
        const int MB = dp_minb();
        size_t tail_smem = (size_t)(NT + 2 * pl.K[m->root] + 2) * sizeof(double);  // the fused reduction's scratch
        if ((flags & WHALE_PEER_SUM) && fused && g1 - g0 == 1 && peer_fusion()) {  // one pass: the sum over ranks rides in the tail
            int32_t rcp = peer_state(D, &a.peer);
            if (rcp != WHALE_OK) return rcp;
            tail_smem = std::max(tail_smem, peer_smem_bytes(D->peer_world, 1 + m->P));
            D->peer_fused = true;
        }
        auto launch_bin = [&](const Bin& b, cudaStream_t s) {
            const bool pdl = pdl_enabled() && bins.size() == 1;
#define LAUNCHV(NTV, MBV)                                                                                                   \
    if (NT == NTV && MB == MBV) {                                                                                           \
        if (pdl) LAUNCH_PDL((k_dp<NTV, MBV>), b.count, NTV, std::max(b.smem, tail_smem), s, a, b.off);                      \
        else LAUNCH((k_dp<NTV, MBV>), b.count, NTV, std::max(b.smem, tail_smem), s, a, b.off);                              \
    }
            DP_VARIANTS(LAUNCHV)
#undef LAUNCHV
            g_launches++;
        };
        if (bins.size() == 1) {
            launch_bin(bins[0], st);
        } else {
            CU(cudaEventRecord(D->ev_fork, st));
            for (size_t b = 0; b < bins.size(); b++) {
                CU(cudaStreamWaitEvent(D->side[b], D->ev_fork, 0));
                launch_bin(bins[b], D->side[b]);
                CU(cudaEventRecord(D->ev_join[b], D->side[b]));
                CU(cudaStreamWaitEvent(st, D->ev_join[b], 0));
            }
        }
        if (prof && first) CU(cudaEventRecord(D->ev[2], st));
        // K3: Σ over families − N·condition; this plan's parameters scattered into d_out
        const int KR = pl.K[m->root];
        int nb = std::min(1024, (F + 255) / 256);
        int chunk = (F + nb - 1) / nb;
        if (!fused) {
            LAUNCH(k_reduce1, nb, 256, 0, st, out_fam, F, KR, chunk, D->d_partial);
            LAUNCH(k_reduce2, 1, 256, 0, st, D->d_partial, nb, KR, F, condition, pl.dev, m->root, first ? 1 : 0, d_out);
            g_launches += 2;
        }
        if (prof && first) CU(cudaEventRecord(D->ev[3], st));
    }
    D->ev_valid = prof;
    CU(cudaGetLastError());
    D->ell_valid = keep;
    return WHALE_OK;
}

// host-pointer evaluations: hp = [x (P) | p_leaf (nn)] in pinned memory -> d_x | d_pleaf (adjacent when P >= 1)
static cudaError_t upload_theta(whale_model* m, const double* hp) {
    if (m->P >= 1) return cudaMemcpyAsync(m->d_x, hp, (size_t)(m->P + m->nn) * sizeof(double), cudaMemcpyHostToDevice, m->stream);
    return cudaMemcpyAsync(m->d_pleaf, hp, (size_t)m->nn * sizeof(double), cudaMemcpyHostToDevice, m->stream);
}

int32_t whale_logpdf_grad_async(whale_model_t m, whale_data_t d, const double* d_x, int32_t condition,
                                uint32_t flags, double* d_out, void* stream) {
    if (!m || !d || !d_x || !d_out) return fail(WHALE_ERR_ARG, "null argument");
    if (d->m != m) return fail(WHALE_ERR_ARG, "data handle belongs to another model");
    CU(cudaSetDevice(m->device));
    d->peer_fused = false;
    int32_t rc = enqueue_eval(m, d, d_x, condition, flags, d_out, (cudaStream_t)stream);
    if (rc == WHALE_OK && (flags & WHALE_PEER_SUM) && !d->peer_fused) rc = enqueue_peer_sum(d, d_out, (cudaStream_t)stream);
    return rc;
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
    double* ho = hp + P + nn;
    // The whole evaluation (H2D of θ, ~10 kernel launches on several streams, D2H of the result) is replayed as
    // ONE CUDA graph from the second call with the same (flags, condition) on: the host-side launch cost is
    // what separates the end-to-end rate from the device-timed rate for small batches.
    bool done = false;
    // the evaluation + (one process per GPU, WHALE_PEER_SUM) the sum over ranks, in the DP kernel's tail or as its own launch;
    // the exchange keeps its step counter on the device, so it can be part of a captured graph like everything else
    auto enqueue_all = [&]() -> int32_t {
        d->peer_fused = false;
        int32_t rc = enqueue_eval(m, d, m->d_x, condition, flags, m->d_out, m->stream);
        if (rc == WHALE_OK && (flags & WHALE_PEER_SUM) && !d->peer_fused) rc = enqueue_peer_sum(d, m->d_out, m->stream);
        return rc;
    };
#ifndef WHALE_EMU
    bool use_graph = !(flags & WHALE_PROFILE) && graphs_mode() != 0;
    if (use_graph && graphs_mode() < 0)  // automatic: single-bin plans only
        for (size_t g = 0; g < d->plans.size() && g < (size_t)MAXPLAN; g++) use_graph = use_graph && d->bins[g].size() <= 1;
    if (use_graph) {
        const uint32_t key = (flags & (WHALE_WANT_GRAD | WHALE_KEEP_ELL | WHALE_PEER_SUM)) | ((uint32_t)condition << 8);
        GraphSlot* gs = nullptr;
        for (auto& g : d->graphs) if (g.key == key) gs = &g;
        if (!gs) { d->graphs.push_back(GraphSlot{key, 0, nullptr}); gs = &d->graphs.back(); }
        if (gs->state == 1) {  // second call: capture (everything the path allocates lazily exists by now)
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                ok = upload_theta(m, hp) == cudaSuccess && enqueue_all() == WHALE_OK &&
                     cudaMemcpyAsync(ho, m->d_out, (1 + P) * sizeof(double), cudaMemcpyDeviceToHost, m->stream) == cudaSuccess;
                ok = (cudaStreamEndCapture(m->stream, &graph) == cudaSuccess) && ok && graph;
            }
            if (ok && cudaGraphInstantiate(&gs->exec, graph, 0) == cudaSuccess) gs->state = 2;
            else { gs->state = 3; cudaGetLastError(); }  // not capturable here: stay on the plain path
            if (graph) cudaGraphDestroy(graph);
        }
        if (gs->state == 2) {
            CU(cudaGraphLaunch(gs->exec, m->stream));
            CU(cudaStreamSynchronize(m->stream));
            d->ell_valid = (flags & WHALE_KEEP_ELL) != 0;
            d->ev_valid = false;
            g_launches += gs->launches;
            done = true;
        } else if (gs->state == 0) {
            gs->state = 1;
            const int64_t l0 = g_launches.load();
            CU(upload_theta(m, hp));
            int32_t rc0 = enqueue_all();
            if (rc0 != WHALE_OK) return rc0;
            gs->launches = g_launches.load() - l0;
            CU(cudaMemcpyAsync(ho, m->d_out, (1 + P) * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
            CU(cudaStreamSynchronize(m->stream));
            done = true;
        }
    }
#endif
    if (!done) {
        CU(upload_theta(m, hp));
        int32_t rc = enqueue_all();
        if (rc != WHALE_OK) return rc;
        CU(cudaMemcpyAsync(ho, m->d_out, (1 + P) * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        CU(cudaStreamSynchronize(m->stream));
    }
    m->x_host_valid = true;
    *loglik = ho[0];
    if (grad) memcpy(grad, ho + 1, P * sizeof(double));
    if (ll_fam || grad_fam) {
        const bool grad_ = (flags & WHALE_WANT_GRAD) != 0;
        if (grad_fam) for (size_t i = 0; i < (size_t)d->F * P; i++) grad_fam[i] = 0.0;
        for (size_t g = grad_ ? 1 : 0; g < (grad_ ? d->plans.size() : 1); g++) {
            const Plan& pl = *d->plans[g];
            const int KR = pl.K[m->root];
            std::vector<double> tmp((size_t)d->F * KR);
            CU(cudaMemcpy(tmp.data(), d->d_out_fam + d->out_off[g], tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int f = 0; f < d->F; f++) {
                if (ll_fam) ll_fam[f] = tmp[(size_t)f * KR];
                if (grad_fam)
                    for (int k = 1; k < KR; k++) grad_fam[(size_t)f * P + pl.act[(size_t)m->root * pl.Kmax + k]] = tmp[(size_t)f * KR + k];
            }
        }
    }
    return WHALE_OK;
}

int32_t whale_mixture_logpdf_grad(whale_model_t m, whale_data_t d, int32_t n_comp, const double* x, const double* log_w,
                                  const double* p_leaf, int32_t condition, uint32_t flags, double* loglik,
                                  double* grad_x, double* grad_logw) {
    if (!m || !d || !x || !log_w || !loglik || n_comp < 1) return fail(WHALE_ERR_ARG, "null argument");
    if (d->m != m) return fail(WHALE_ERR_ARG, "data handle belongs to another model");
    if (condition < 0 || condition > 3) return fail(WHALE_ERR_ARG, "unknown condition kind %d", condition);
    const bool grad = (flags & WHALE_WANT_GRAD) != 0;
    if ((grad_x || grad_logw) && !grad) return fail(WHALE_ERR_ARG, "grad requested without WHALE_WANT_GRAD");
    CU(cudaSetDevice(m->device));
    const int P = m->P, nn = m->nn, F = d->F, J = n_comp;
    const size_t comp = (size_t)F * (1 + P);
    double *dense = nullptr, *condv = nullptr, *logw = nullptr, *lse = nullptr, *resp = nullptr, *out = nullptr, *dx = nullptr;
    const int Q = 1 + J * (1 + P);
    auto cleanup = [&]() { cudaFree(dense); cudaFree(condv); cudaFree(logw); cudaFree(lse); cudaFree(resp); cudaFree(out); cudaFree(dx); };
#define CUM(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(WHALE_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
    CUM(cudaMalloc((void**)&dense, J * comp * sizeof(double)));
    CUM(cudaMalloc((void**)&condv, (size_t)J * (1 + P) * sizeof(double)));
    CUM(cudaMalloc((void**)&logw, J * sizeof(double)));
    CUM(cudaMalloc((void**)&lse, (size_t)F * sizeof(double)));
    CUM(cudaMalloc((void**)&resp, (size_t)F * J * sizeof(double)));
    CUM(cudaMalloc((void**)&out, Q * sizeof(double)));
    CUM(cudaMalloc((void**)&dx, (size_t)J * P * sizeof(double)));
    cudaStream_t st = m->stream;
    CUM(cudaMemsetAsync(dense, 0, J * comp * sizeof(double), st));
    CUM(cudaMemsetAsync(condv, 0, (size_t)J * (1 + P) * sizeof(double), st));
    CUM(cudaMemcpyAsync(logw, log_w, J * sizeof(double), cudaMemcpyHostToDevice, st));
    CUM(cudaMemcpyAsync(dx, x, (size_t)J * P * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<double> pl(nn, 0.0);
    if (p_leaf) pl.assign(p_leaf, p_leaf + nn);
    CUM(cudaMemcpyAsync(m->d_pleaf, pl.data(), nn * sizeof(double), cudaMemcpyHostToDevice, st));
    CUM(cudaStreamSynchronize(st));  // pl is a stack-lifetime buffer
    for (int j = 0; j < J; j++) {
        // per-family outputs of component j (every tangent plan), then scattered into its dense rows
        int32_t rc = enqueue_eval(m, d, dx + (size_t)j * P, condition, flags & WHALE_WANT_GRAD, m->d_out, st);
        if (rc != WHALE_OK) { cleanup(); return rc; }
        const size_t g0 = grad ? 1 : 0, g1 = grad ? d->plans.size() : 1;
        for (size_t g = g0; g < g1; g++) {
            const Plan& plg = *d->plans[g];
            const int KR = plg.K[m->root];
            const long long tot = (long long)F * KR;
            LAUNCH(k_mix_gather, (int)((tot + 255) / 256), 256, 0, st, d->d_out_fam + d->out_off[g], F, KR,
                   plg.dev.act + (size_t)m->root * plg.Kmax, g == g0 ? 1 : 0, plg.dev.cond + (size_t)condition * plg.Kmax,
                   dense + j * comp, condv + (size_t)j * (1 + P), P);
            g_launches++;
        }
        // NOTE: with several gradient passes the plans' tables are reused by the next component only after these
        // launches (same stream), so reading plg.dev.cond here is ordered correctly.
    }
    LAUNCH(k_mix_resp, (F + 255) / 256, 256, 0, st, dense, condv, logw, F, P, J, lse, resp);
    LAUNCH(k_mix_reduce, grad ? Q : 1, 256, 0, st, dense, condv, lse, resp, F, P, J, out);
    g_launches += 2;
    std::vector<double> ho(Q, 0.0);
    CUM(cudaMemcpyAsync(ho.data(), out, (grad ? Q : 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUM(cudaStreamSynchronize(st));
    CUM(cudaGetLastError());
#undef CUM
    cleanup();
    d->ell_valid = false;
    d->ev_valid = false;
    const bool fin = std::isfinite(ho[0]);
    *loglik = fin ? ho[0] : -INFINITY;  // ℓhood src/core.jl:15
    for (int j = 0; j < J; j++) {
        if (grad_logw) grad_logw[j] = fin ? ho[1 + (size_t)j * (1 + P)] : 0.0;
        if (grad_x) for (int p2 = 0; p2 < P; p2++) grad_x[(size_t)j * P + p2] = fin ? ho[1 + (size_t)j * (1 + P) + 1 + p2] : 0.0;
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
    CU(launch_tables(m, p0, m->d_x, m->d_pleaf, m->stream, false));
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

// ---- backtracking: persistent buffers, walks, compaction, summaries ----
static int32_t ensure_tree_bufs(whale_data* D, long long W, int max_nodes, size_t stack_walks) {
    auto& T = D->tb;
    const size_t w = (size_t)W, n = (size_t)W * max_nodes;
    if (w > T.capW) {
        size_t c1 = T.capW, c2 = T.capW;
        CU(grow_dev(&T.d_cnt, &c1, w));
        CU(grow_dev(&T.d_st, &c2, w));
        T.capW = std::min(c1, c2);
    }
    if (n > T.capN) {
        size_t c[4] = {T.capN, T.capN, T.capN, T.capN};
        CU(grow_dev(&T.d_g, &c[0], n)); CU(grow_dev(&T.d_e, &c[1], n)); CU(grow_dev(&T.d_t, &c[2], n)); CU(grow_dev(&T.d_p, &c[3], n));
        T.capN = std::min(std::min(c[0], c[1]), std::min(c[2], c[3]));
    }
    CU(grow_dev(&T.d_stack, &T.cap_stack, stack_walks * (size_t)max_nodes));
    if (w > T.hcapW) {
        size_t c1 = T.hcapW, c2 = T.hcapW;
        CU(grow_pinned(&T.h_cnt, &c1, w));
        CU(grow_pinned(&T.h_st, &c2, w));
        T.hcapW = std::min(c1, c2);
    }
    T.W = W; T.max_nodes = max_nodes; T.valid = false; T.packed_valid = false;
    return WHALE_OK;
}

// after the walks: counts and statuses to the host (pinned), total node count
static int32_t finish_walks(whale_data* D, cudaStream_t st, int S) {
    auto& T = D->tb;
    CU(cudaMemcpyAsync(T.h_cnt, T.d_cnt, (size_t)T.W * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(T.h_st, T.d_st, (size_t)T.W * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    long long tot = 0;
    for (long long w = 0; w < T.W; w++) tot += T.h_cnt[w];
    T.total_nodes = tot;
    T.S = S;
    T.valid = true;
    return WHALE_OK;
}

int32_t whale_backtrack(whale_model_t m, whale_data_t d, int32_t n_samples, const double* uniforms, int64_t stride,
                        int32_t max_nodes, int32_t* node_count, int32_t* gamma, int32_t* e, int32_t* t, int32_t* parent,
                        int32_t* status) {
    if (!m || !d || !uniforms || !node_count || !gamma || !e || !t || !parent || !status)
        return fail(WHALE_ERR_ARG, "null argument");
    int32_t rc = whale_backtrack_device(m, d, n_samples, uniforms, stride, 0ull, max_nodes, nullptr);
    if (rc != WHALE_OK) return rc;
    auto& T = d->tb;
    const size_t W = (size_t)T.W, N = W * (size_t)max_nodes;
    memcpy(node_count, T.h_cnt, W * 4);
    memcpy(status, T.h_st, W * 4);
    CU(cudaMemcpy(gamma, T.d_g, N * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(e, T.d_e, N * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(t, T.d_t, N * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(parent, T.d_p, N * 4, cudaMemcpyDeviceToHost));
    return WHALE_OK;
}

int32_t whale_backtrack_device(whale_model_t m, whale_data_t d, int32_t n_samples, const double* uniforms, int64_t stride,
                               uint64_t seed, int32_t max_nodes, int64_t* total_nodes) {
    if (!m || !d) return fail(WHALE_ERR_ARG, "null argument");
    if (d->m != m) return fail(WHALE_ERR_ARG, "data handle belongs to another model");
    if (n_samples <= 0 || max_nodes <= 1 || (uniforms && stride <= 0)) return fail(WHALE_ERR_ARG, "n_samples, stride and max_nodes must be positive");
    if (!d->ell_valid || !d->d_ell) return fail(WHALE_ERR_STATE, "no ℓ kept: evaluate with WHALE_KEEP_ELL (logpdf!) first");
    CU(cudaSetDevice(m->device));
    const long long W = (long long)d->F * n_samples;
    int32_t rc = ensure_tree_bufs(d, W, max_nodes, (size_t)W);
    if (rc != WHALE_OK) return rc;
    auto& T = d->tb;
    cudaStream_t st = m->stream;
    if (uniforms) {
        CU(grow_dev(&T.d_u, &T.cap_u, (size_t)W * stride));
        CU(cudaMemcpyAsync(T.d_u, uniforms, (size_t)W * stride * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    Plan& p0 = m->plan[0];
    CU(launch_tables(m, p0, d->d_x_keep, d->d_pleaf_keep, st, false));
    BTArgs a{m->dev, p0.dev, d->d_arena, d->d_hdr, d->d_ell, d->d_x_keep, uniforms ? T.d_u : nullptr,
             uniforms ? (long long)stride : (1LL << 40), d->F, n_samples, max_nodes, 0, n_samples, T.d_cnt, T.d_g, T.d_e,
             T.d_t, T.d_p, T.d_st, T.d_stack, nullptr, 0, (unsigned long long)seed};
    if (!d->ev[0]) for (int i = 0; i < 4; i++) CU(cudaEventCreate(&d->ev[i]));
    CU(cudaEventRecord(d->ev[0], st));
    LAUNCH(k_backtrack, (int)((W + 127) / 128), 128, 0, st, a);
    CU(cudaEventRecord(d->ev[1], st));
    g_launches += 2;
    rc = finish_walks(d, st, n_samples);
    if (rc != WHALE_OK) return rc;
    { float ms = 0; cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]); d->last_bt_ms = ms; d->ev_valid = false; }
    if (total_nodes) *total_nodes = T.total_nodes;
    return WHALE_OK;
}

// value-only DP keeping ℓ for a SUBSET of the families (the launch order is the given list): logpdf! for the families
// whose sample uses this posterior draw
static int32_t enqueue_keep_subset(whale_model* m, whale_data* D, const double* d_x, const int* d_famlist, int count, cudaStream_t st,
                                   Plan* slot_plan = nullptr, double* slot_ell = nullptr) {
    if (count <= 0) return WHALE_OK;
    if (!slot_ell && !D->d_ell) CU(cudaMalloc((void**)&D->d_ell, std::max<uint64_t>(D->ell_total, 1) * sizeof(double)));
    if (!m->attr_set) {
#define SETATTR(NTV, MB) CU(cudaFuncSetAttribute(k_dp<NTV, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DP_VARIANTS(SETATTR)
#undef SETATTR
#define SETATTR(NTV, MB) CU(cudaFuncSetAttribute(k_dp_rev<NTV, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        REV_VARIANTS(SETATTR)
#undef SETATTR
        m->attr_set = true;
    }
    Plan& p0 = slot_plan ? *slot_plan : m->plan[0];
    CU(launch_tables(m, p0, d_x, m->d_pleaf, st, false));
    size_t smem = 0;
    for (const Bin& b : D->bins[0]) smem = std::max(smem, b.smem);
    DPArgs a{m->dev, p0.dev, D->d_arena, D->d_hdr, d_famlist, D->d_roff[0], D->d_out_fam + D->out_off[0], slot_ell ? slot_ell : D->d_ell, 0, 0,
             nullptr, nullptr, D->F, 0, 1, m->d_out};
    const int NT = dp_nt(), MB = dp_minb();
#define LAUNCHV(NTV, MBV) if (NT == NTV && MB == MBV) LAUNCH((k_dp<NTV, MBV>), count, NTV, smem, st, a, 0);
    DP_VARIANTS(LAUNCHV)
#undef LAUNCHV
    g_launches++;
    return WHALE_OK;
}

int32_t whale_track_sample(whale_model_t m, whale_data_t d, int32_t n_theta, const double* x, const double* p_leaf,
                           int32_t n_samples, const int32_t* theta_index, const double* uniforms, int64_t stride,
                           uint64_t seed, int32_t max_nodes, int64_t* total_nodes) {
    if (!m || !d || !x) return fail(WHALE_ERR_ARG, "null argument");
    if (d->m != m) return fail(WHALE_ERR_ARG, "data handle belongs to another model");
    if (n_theta <= 0 || n_samples <= 0 || max_nodes <= 1 || (uniforms && stride <= 0)) return fail(WHALE_ERR_ARG, "n_theta, n_samples, stride and max_nodes must be positive");
    CU(cudaSetDevice(m->device));
    const int P = m->P, nn = m->nn, F = d->F, S = n_samples;
    const long long W = (long long)F * S;
    // group the (family, sample) pairs by posterior draw
    std::vector<int> cnt(n_theta + 1, 0);
    auto idx = [&](long long w) -> int { return theta_index ? theta_index[w] : (int)(w % S) % n_theta; };
    for (long long w = 0; w < W; w++) {
        const int j = idx(w);
        if (j < 0 || j >= n_theta) return fail(WHALE_ERR_ARG, "theta_index out of range");
        cnt[j + 1]++;
    }
    for (int j = 0; j < n_theta; j++) cnt[j + 1] += cnt[j];
    std::vector<int2> pairs((size_t)W);
    std::vector<int> fill(cnt.begin(), cnt.end() - 1), famlist((size_t)W), famoff(n_theta + 1, 0);
    for (long long w = 0; w < W; w++) pairs[fill[idx(w)]++] = make_int2((int)(w / S), (int)(w % S));  // (family, sample), family-major
    size_t nf = 0, maxgroup = 0;
    for (int j = 0; j < n_theta; j++) {
        famoff[j] = (int)nf;
        for (int i = cnt[j]; i < cnt[j + 1]; i++)
            if (i == cnt[j] || pairs[i].x != pairs[i - 1].x) famlist[nf++] = pairs[i].x;
        maxgroup = std::max(maxgroup, (size_t)(cnt[j + 1] - cnt[j]));
    }
    famoff[n_theta] = (int)nf;
    auto& T = d->tb;
    // posterior draws handled per batch (see below): limited by free device memory (an ℓ buffer each)
    int nslot = std::min(n_theta, env_int("WHALE_TRACK_SLOTS", 16));
    {
        size_t fre = 0, tot = 0;
        const size_t ellb = std::max<uint64_t>(d->ell_total, 1) * sizeof(double);
        if (cudaMemGetInfo(&fre, &tot) != cudaSuccess) { cudaGetLastError(); fre = 0; }
        fre += T.cap_ell_slots * sizeof(double);  // what we already hold counts as available
        nslot = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(nslot, 1), (fre / 2) / ellb));
    }
    size_t maxbatch = 0;
    for (int j0 = 0; j0 < n_theta; j0 += nslot) maxbatch = std::max(maxbatch, (size_t)(cnt[std::min(n_theta, j0 + nslot)] - cnt[j0]));
    int32_t rc = ensure_tree_bufs(d, W, max_nodes, std::max(maxgroup, maxbatch));
    if (rc != WHALE_OK) return rc;
    cudaStream_t st = m->stream;
    if (uniforms) {
        CU(grow_dev(&T.d_u, &T.cap_u, (size_t)W * stride));
        CU(cudaMemcpyAsync(T.d_u, uniforms, (size_t)W * stride * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CU(grow_dev(&T.d_xs, &T.cap_xs, (size_t)n_theta * P));
    CU(cudaMemcpyAsync(T.d_xs, x, (size_t)n_theta * P * sizeof(double), cudaMemcpyHostToDevice, st));
    if (T.cap_pairs < (size_t)W) {
        if (T.d_pairs) cudaFree(T.d_pairs);
        if (T.d_famlist) cudaFree(T.d_famlist);
        T.d_pairs = nullptr; T.d_famlist = nullptr; T.cap_pairs = 0;
        CU(cudaMalloc((void**)&T.d_pairs, (size_t)W * sizeof(int2)));
        CU(cudaMalloc((void**)&T.d_famlist, (size_t)W * sizeof(int)));
        T.cap_pairs = (size_t)W;
    }
    CU(cudaMemcpyAsync(T.d_pairs, pairs.data(), (size_t)W * sizeof(int2), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(T.d_famlist, famlist.data(), nf * sizeof(int), cudaMemcpyHostToDevice, st));
    std::vector<double> pl(nn, 0.0);
    if (p_leaf) pl.assign(p_leaf, p_leaf + nn);
    CU(cudaMemcpyAsync(m->d_pleaf, pl.data(), nn * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));  // the host vectors above go out of scope
    if (!d->ev[0]) for (int i = 0; i < 4; i++) CU(cudaEventCreate(&d->ev[i]));
    CU(cudaEventRecord(d->ev[0], st));
    Plan& p0 = m->plan[0];
    // Batches of `nslot` posterior draws: logpdf!(model(θ_j), ·) for the families that drew row j into slot j's own ℓ and
    // slice tables, back to back, then ONE walk launch over all the batch's pairs (a launch per draw holds ~W/n_theta
    // walks — a few CTAs per SM — and cost 2.5 ms per draw on C5 where all 10^6 walks together take 28 ms).  Nothing
    // returns to the host in between.  Slots are limited by free device memory (an ℓ buffer each); one slot = the
    // draw-by-draw path.
    if (nslot > 1) {
        CU(grow_dev(&T.d_ell_slots, &T.cap_ell_slots, (size_t)nslot * d->ell_total));
        if ((int)T.slot_plans.size() < nslot) {
            const size_t have = T.slot_plans.size();
            T.slot_plans.resize(nslot);
            for (size_t q = have; q < (size_t)nslot; q++) {
                build_plan(*m, std::vector<char>(), T.slot_plans[q]);
                CU(upload_plan(T.slot_plans[q], nn));
            }
            cudaFree((void*)T.d_slot_eps); cudaFree((void*)T.d_slot_pp);
            T.d_slot_eps = nullptr; T.d_slot_pp = nullptr;
            std::vector<const double*> he(nslot);
            std::vector<const double2*> hp2(nslot);
            for (int q = 0; q < nslot; q++) { he[q] = T.slot_plans[q].dev.eps; hp2[q] = T.slot_plans[q].dev.pp; }
            CU(cudaMalloc((void**)&T.d_slot_eps, nslot * sizeof(double*)));
            CU(cudaMalloc((void**)&T.d_slot_pp, nslot * sizeof(double2*)));
            CU(cudaMemcpy((void*)T.d_slot_eps, he.data(), nslot * sizeof(double*), cudaMemcpyHostToDevice));
            CU(cudaMemcpy((void*)T.d_slot_pp, hp2.data(), nslot * sizeof(double2*), cudaMemcpyHostToDevice));
        }
        // per batch: the pair offsets of its draws relative to the batch's first pair, nb + 1 entries
        std::vector<int> relx, batch_at;
        for (int j0 = 0; j0 < n_theta; j0 += nslot) {
            const int j1 = std::min(n_theta, j0 + nslot);
            batch_at.push_back((int)relx.size());
            for (int j = j0; j <= j1; j++) relx.push_back(cnt[j] - cnt[j0]);
        }
        if ((int)relx.size() > T.nslot) {
            cudaFree(T.d_slot_off); T.d_slot_off = nullptr; T.nslot = 0;
            CU(cudaMalloc((void**)&T.d_slot_off, relx.size() * sizeof(int)));
            T.nslot = (int)relx.size();
        }
        CU(cudaMemcpy(T.d_slot_off, relx.data(), relx.size() * sizeof(int), cudaMemcpyHostToDevice));
        int bi = 0;
        for (int j0 = 0; j0 < n_theta; j0 += nslot, bi++) {
            const int j1 = std::min(n_theta, j0 + nslot), nb = j1 - j0;
            const int np = cnt[j1] - cnt[j0];
            if (np == 0) continue;
            for (int j = j0; j < j1; j++) {
                rc = enqueue_keep_subset(m, d, T.d_xs + (size_t)j * P, T.d_famlist + famoff[j], famoff[j + 1] - famoff[j], st,
                                         &T.slot_plans[j - j0], T.d_ell_slots + (size_t)(j - j0) * d->ell_total);
                if (rc != WHALE_OK) return rc;
            }
            BTArgs a{m->dev, p0.dev, d->d_arena, d->d_hdr, T.d_ell_slots, T.d_xs + (size_t)j0 * P, uniforms ? T.d_u : nullptr,
                     uniforms ? (long long)stride : (1LL << 40), F, S, max_nodes, 0, S, T.d_cnt, T.d_g, T.d_e, T.d_t, T.d_p, T.d_st,
                     T.d_stack, T.d_pairs + cnt[j0], np, (unsigned long long)seed,
                     nb, T.d_slot_off + batch_at[bi], T.d_slot_eps, T.d_slot_pp, P, (unsigned long long)d->ell_total};
            LAUNCH(k_backtrack, (np + 127) / 128, 128, 0, st, a);
            g_launches++;
        }
    } else {
        for (int j = 0; j < n_theta; j++) {
            const int np = cnt[j + 1] - cnt[j];
            if (np == 0) continue;
            const double* d_xj = T.d_xs + (size_t)j * P;
            rc = enqueue_keep_subset(m, d, d_xj, T.d_famlist + famoff[j], famoff[j + 1] - famoff[j], st);
            if (rc != WHALE_OK) return rc;
            BTArgs a{m->dev, p0.dev, d->d_arena, d->d_hdr, d->d_ell, d_xj, uniforms ? T.d_u : nullptr,
                     uniforms ? (long long)stride : (1LL << 40), F, S, max_nodes, 0, S, T.d_cnt, T.d_g, T.d_e, T.d_t, T.d_p, T.d_st,
                     T.d_stack, T.d_pairs + cnt[j], np, (unsigned long long)seed};
            LAUNCH(k_backtrack, (np + 127) / 128, 128, 0, st, a);
            g_launches++;
        }
    }
    CU(cudaEventRecord(d->ev[1], st));
    rc = finish_walks(d, st, S);
    if (rc != WHALE_OK) return rc;
    { float ms = 0; cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]); d->last_bt_ms = ms; d->ev_valid = false; }
    d->ell_valid = false;  // ℓ holds a mixture of draws
    m->x_host_valid = false;
    if (total_nodes) *total_nodes = T.total_nodes;
    return WHALE_OK;
}

int32_t whale_trees_counts(whale_data_t d, int32_t* node_count, int32_t* status) {
    if (!d) return fail(WHALE_ERR_ARG, "null argument");
    if (!d->tb.valid) return fail(WHALE_ERR_STATE, "no backtracked trees on this handle");
    if (node_count) memcpy(node_count, d->tb.h_cnt, (size_t)d->tb.W * 4);
    if (status) memcpy(status, d->tb.h_st, (size_t)d->tb.W * 4);
    return WHALE_OK;
}

// compact (γ, e, t, parent) rows of all walks in (family, sample) order: offsets[W+1] and total_nodes rows, in
// library-owned pinned host memory (valid until the next backtracking call on this handle)
int32_t whale_trees_view(whale_data_t d, const int64_t** offsets, const int32_t** nodes4, int64_t* total_nodes) {
    if (!d) return fail(WHALE_ERR_ARG, "null argument");
    auto& T = d->tb;
    if (!T.valid) return fail(WHALE_ERR_STATE, "no backtracked trees on this handle");
    whale_model* m = d->m;
    CU(cudaSetDevice(m->device));
    if (!T.packed_valid) {
        const size_t W = (size_t)T.W;
        CU(grow_pinned(&T.h_off, &T.hcap_off, W + 1));
        long long o = 0;
        for (size_t w = 0; w < W; w++) { T.h_off[w] = o; o += T.h_cnt[w]; }
        T.h_off[W] = o;
        CU(grow_dev(&T.d_off, &T.cap_off, W + 1));
        CU(grow_dev(&T.d_packed, &T.cap_packed, (size_t)std::max<long long>(o, 1)));
        CU(grow_pinned(&T.h_packed, &T.hcap_packed, (size_t)std::max<long long>(o, 1) * 4));
        cudaStream_t st = m->stream;
        CU(cudaMemcpyAsync(T.d_off, T.h_off, (W + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
        LAUNCH(k_tree_pack, (int)((W * 32 + 127) / 128), 128, 0, st, T.d_cnt, T.d_off, T.d_g, T.d_e, T.d_t, T.d_p, T.max_nodes,
               (long long)W, T.d_packed);
        g_launches++;
        CU(cudaMemcpyAsync(T.h_packed, T.d_packed, (size_t)o * sizeof(int4), cudaMemcpyDeviceToHost, st));
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));
        T.packed_valid = true;
    }
    static_assert(sizeof(long long) == sizeof(int64_t), "offsets are int64");
    if (offsets) *offsets = reinterpret_cast<const int64_t*>(T.h_off);
    if (nodes4) *nodes4 = T.h_packed;
    if (total_nodes) *total_nodes = T.total_nodes;
    return WHALE_OK;
}

int32_t whale_trees_get(whale_data_t d, int64_t* offsets, int32_t* nodes4) {
    const int64_t* o = nullptr;
    const int32_t* n = nullptr;
    int64_t tot = 0;
    int32_t rc = whale_trees_view(d, &o, &n, &tot);
    if (rc != WHALE_OK) return rc;
    if (offsets) memcpy(offsets, o, ((size_t)d->tb.W + 1) * sizeof(int64_t));
    if (nodes4) memcpy(nodes4, n, (size_t)tot * 16);
    return WHALE_OK;
}

// sumtrees on the device: identity hash of every tree, then per family the distinct trees with counts and first samples
int32_t whale_trees_summary(whale_data_t d, int32_t* n_distinct, uint64_t* hash, int32_t* count, int32_t* first,
                            uint64_t* tree_hash) {
    if (!d || !n_distinct || !hash || !count || !first) return fail(WHALE_ERR_ARG, "null argument");
    auto& T = d->tb;
    if (!T.valid) return fail(WHALE_ERR_STATE, "no backtracked trees on this handle");
    whale_model* m = d->m;
    CU(cudaSetDevice(m->device));
    const size_t W = (size_t)T.W;
    const int S = T.S, F = (int)(W / S);
    int Spad = 1;
    while (Spad < S) Spad <<= 1;
    const size_t smem = (size_t)Spad * 12;
    if (smem > 200 * 1024) return fail(WHALE_ERR_CAPACITY, "%d samples per family: summarise on the host", S);
    if (W > T.cap_hash) {
        size_t c[4] = {T.cap_hash, T.cap_hash, T.cap_hash, T.cap_hash};
        CU(grow_dev(&T.d_hash, &c[0], W)); CU(grow_dev(&T.d_dh, &c[1], W)); CU(grow_dev(&T.d_dc, &c[2], W)); CU(grow_dev(&T.d_df, &c[3], W));
        T.cap_hash = std::min(std::min(c[0], c[1]), std::min(c[2], c[3]));
    }
    CU(grow_dev(&T.d_nd, &T.cap_nd, (size_t)F));
    CU(grow_dev(&T.d_stack, &T.cap_stack, W * (size_t)T.max_nodes));  // scratch of the hash kernel
    cudaStream_t st = m->stream;
    HashArgs a{T.d_cnt, T.d_g, T.d_e, T.d_p, T.d_st, T.d_stack, T.d_hash, (long long)W, T.max_nodes};
    LAUNCH(k_tree_hash, (int)((W + 127) / 128), 128, 0, st, a);
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_tree_dedup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(k_tree_dedup, F, 256, smem, st, T.d_hash, S, Spad, T.d_nd, T.d_dh, T.d_dc, T.d_df);
    g_launches += 2;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(n_distinct, T.d_nd, (size_t)F * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hash, T.d_dh, W * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(count, T.d_dc, W * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(first, T.d_df, W * 4, cudaMemcpyDeviceToHost));
    if (tree_hash) CU(cudaMemcpy(tree_hash, T.d_hash, W * 8, cudaMemcpyDeviceToHost));
    return WHALE_OK;
}

int32_t whale_track(whale_model_t m, whale_data_t d, int32_t n_theta, const double* x, const double* p_leaf,
                    int32_t condition, const double* uniforms, int64_t stride, int32_t max_nodes, int32_t* node_count,
                    int32_t* gamma, int32_t* e, int32_t* t, int32_t* parent, int32_t* status, double* loglik) {
    if (!m || !d || !x || !uniforms || !node_count || !gamma || !e || !t || !parent || !status)
        return fail(WHALE_ERR_ARG, "null argument");
    if (d->m != m) return fail(WHALE_ERR_ARG, "data handle belongs to another model");
    if (n_theta <= 0 || stride <= 0 || max_nodes <= 1) return fail(WHALE_ERR_ARG, "n_theta, stride and max_nodes must be positive");
    CU(cudaSetDevice(m->device));
    const int P = m->P, nn = m->nn;
    const size_t W = (size_t)d->F * n_theta;
    double *d_u = nullptr, *d_xs = nullptr, *d_outs = nullptr;
    int32_t *d_cnt = nullptr, *d_g = nullptr, *d_e = nullptr, *d_t = nullptr, *d_p = nullptr, *d_st = nullptr;
    int4* d_stack = nullptr;
    cudaEvent_t eb0 = nullptr, eb1 = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_u); cudaFree(d_xs); cudaFree(d_outs); cudaFree(d_cnt); cudaFree(d_g); cudaFree(d_e); cudaFree(d_t);
        cudaFree(d_p); cudaFree(d_st); cudaFree(d_stack);
        if (eb0) cudaEventDestroy(eb0);
        if (eb1) cudaEventDestroy(eb1);
    };
#define CUB(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return fail(WHALE_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); } } while (0)
    cudaStream_t st = m->stream;
    CUB(cudaMalloc((void**)&d_u, W * stride * sizeof(double)));
    CUB(cudaMemcpyAsync(d_u, uniforms, W * stride * sizeof(double), cudaMemcpyHostToDevice, st));
    CUB(cudaMalloc((void**)&d_xs, (size_t)n_theta * P * sizeof(double)));
    CUB(cudaMemcpyAsync(d_xs, x, (size_t)n_theta * P * sizeof(double), cudaMemcpyHostToDevice, st));
    CUB(cudaMalloc((void**)&d_outs, (size_t)n_theta * (1 + P) * sizeof(double)));
    CUB(cudaMalloc((void**)&d_cnt, W * 4)); CUB(cudaMalloc((void**)&d_st, W * 4));
    CUB(cudaMalloc((void**)&d_g, W * max_nodes * 4)); CUB(cudaMalloc((void**)&d_e, W * max_nodes * 4));
    CUB(cudaMalloc((void**)&d_t, W * max_nodes * 4)); CUB(cudaMalloc((void**)&d_p, W * max_nodes * 4));
    CUB(cudaMalloc((void**)&d_stack, (size_t)d->F * max_nodes * sizeof(int4)));  // one sample in flight at a time
    std::vector<double> pl(nn, 0.0);
    if (p_leaf) pl.assign(p_leaf, p_leaf + nn);
    CUB(cudaMemcpyAsync(m->d_pleaf, pl.data(), nn * sizeof(double), cudaMemcpyHostToDevice, st));
    CUB(cudaStreamSynchronize(st));
    CUB(cudaEventCreate(&eb0)); CUB(cudaEventCreate(&eb1));
    CUB(cudaEventRecord(eb0, st));
    Plan& p0 = m->plan[0];
    for (int s = 0; s < n_theta; s++) {
        // logpdf!(model(θ_s), ccds) keeping ℓ, then one walk per family from it — nothing returns to the host in between
        int32_t rc = enqueue_eval(m, d, d_xs + (size_t)s * P, condition, WHALE_KEEP_ELL, d_outs + (size_t)s * (1 + P), st);
        if (rc != WHALE_OK) { cleanup(); return rc; }
        BTArgs a{m->dev, p0.dev, d->d_arena, d->d_hdr, d->d_ell, d_xs + (size_t)s * P, d_u, (long long)stride, d->F, 1, max_nodes,
                 s, n_theta, d_cnt, d_g, d_e, d_t, d_p, d_st, d_stack};
        LAUNCH(k_backtrack, (d->F + 127) / 128, 128, 0, st, a);
        g_launches++;
    }
    CUB(cudaEventRecord(eb1, st));
    CUB(cudaGetLastError());
    CUB(cudaStreamSynchronize(st));
    { float ms = 0; cudaEventElapsedTime(&ms, eb0, eb1); d->last_bt_ms = ms; }
    CUB(cudaMemcpy(node_count, d_cnt, W * 4, cudaMemcpyDeviceToHost));
    CUB(cudaMemcpy(status, d_st, W * 4, cudaMemcpyDeviceToHost));
    CUB(cudaMemcpy(gamma, d_g, W * max_nodes * 4, cudaMemcpyDeviceToHost));
    CUB(cudaMemcpy(e, d_e, W * max_nodes * 4, cudaMemcpyDeviceToHost));
    CUB(cudaMemcpy(t, d_t, W * max_nodes * 4, cudaMemcpyDeviceToHost));
    CUB(cudaMemcpy(parent, d_p, W * max_nodes * 4, cudaMemcpyDeviceToHost));
    if (loglik) {
        std::vector<double> ho((size_t)n_theta * (1 + P));
        CUB(cudaMemcpy(ho.data(), d_outs, ho.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int s = 0; s < n_theta; s++) loglik[s] = ho[(size_t)s * (1 + P)];
    }
#undef CUB
    cleanup();
    d->ell_valid = true;      // ℓ of the last θ stays on the device
    m->x_host_valid = false;  // ... but m->d_x does not hold it
    return WHALE_OK;
}

int32_t whale_work_estimate(whale_model_t m, whale_data_t d, uint32_t flags, double* flops, double* bytes) {
    if (!m || !d) return fail(WHALE_ERR_ARG, "null argument");
    const Plan& pl = m->plan[(flags & WHALE_WANT_GRAD) ? 1 : 0];  // the algorithmic count ignores chunking
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

int32_t whale_last_backtrack_ms(whale_data_t d, double* ms) {
    if (!d || !ms) return fail(WHALE_ERR_ARG, "null argument");
    *ms = d->last_bt_ms;
    return WHALE_OK;
}

int32_t whale_last_phase_cycles(whale_data_t d, double* mean8, double* max8) {
    if (!d || !mean8 || !max8) return fail(WHALE_ERR_ARG, "null argument");
    if (!d->ev_valid || !d->d_tim) return fail(WHALE_ERR_STATE, "last evaluation was not run with WHALE_PROFILE");
    CU(cudaSetDevice(d->m->device));
    CU(cudaEventSynchronize(d->ev[3]));
    std::vector<long long> h((size_t)d->F * TIMW);
    CU(cudaMemcpy(h.data(), d->d_tim, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int j = 0; j < 8; j++) {
        double s = 0, mx = 0;
        for (int f = 0; f < d->F; f++) { s += (double)h[(size_t)f * TIMW + j]; mx = std::max(mx, (double)h[(size_t)f * TIMW + j]); }
        mean8[j] = s / d->F;
        max8[j] = mx;
    }
    return WHALE_OK;
}

int32_t whale_last_family_cycles(whale_data_t d, double* out8) {
    if (!d || !out8) return fail(WHALE_ERR_ARG, "null argument");
    if (!d->ev_valid || !d->d_tim) return fail(WHALE_ERR_STATE, "last evaluation was not run with WHALE_PROFILE");
    CU(cudaSetDevice(d->m->device));
    CU(cudaEventSynchronize(d->ev[3]));
    std::vector<long long> h((size_t)d->F * TIMW);
    CU(cudaMemcpy(h.data(), d->d_tim, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int f = 0; f < d->F; f++)
        for (int j = 0; j < 8; j++) out8[(size_t)f * 8 + j] = (double)h[(size_t)f * TIMW + j];
    return WHALE_OK;
}

int32_t whale_last_node_cycles(whale_data_t d, double* slices_mean, double* row1_mean, int32_t cap) {
    if (!d || !slices_mean || !row1_mean) return fail(WHALE_ERR_ARG, "null argument");
    if (!d->ev_valid || !d->d_tim) return fail(WHALE_ERR_STATE, "last evaluation was not run with WHALE_PROFILE");
    CU(cudaSetDevice(d->m->device));
    CU(cudaEventSynchronize(d->ev[3]));
    std::vector<long long> h((size_t)d->F * TIMW);
    CU(cudaMemcpy(h.data(), d->d_tim, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    const int n = std::min<int>({cap, TIMN, (int)d->m->inner.size()});
    for (int j = 0; j < n; j++) {
        double a = 0, b = 0;
        for (int f = 0; f < d->F; f++) { a += (double)h[(size_t)f * TIMW + 8 + j]; b += (double)h[(size_t)f * TIMW + 8 + TIMN + j]; }
        slices_mean[j] = a / d->F;
        row1_mean[j] = b / d->F;
    }
    return n;
}

int32_t whale_last_tables_cycles(whale_model_t m, int32_t with_grad, double* out32) {
    if (!m || !out32) return fail(WHALE_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->device));
    CU(cudaDeviceSynchronize());
    long long h[32];
    CU(cudaMemcpy(h, m->plan[with_grad ? 1 : 0].dev.tim, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 32; i++) out32[i] = (double)h[i];
    return WHALE_OK;
}

// ------------------------------------------------------------------------------------------------
// Several GPUs behind ONE handle: whale_set_devices picks the devices; whale_multi_create packs the
// model once per device and shards the families over them by predicted work (longest processing time first); an
// evaluation enqueues [θ H2D -> tables -> DP -> reduction -> result D2H] on every device's own stream, waits for all of
// them and adds the per-device results in device-list order (fixed order: deterministic bits).  The host needs the
// 1+P doubles anyway, so the exchange is n pinned D2H copies in flight together — no device-side collective.
// (One process per GPU instead — the benchmark's layout — exchanges through peer memory: whale_peer_* below.)
// ------------------------------------------------------------------------------------------------
static std::vector<int> g_devices;

struct whale_multi {
    std::vector<whale_model*> models;
    std::vector<whale_data*> datas;
    std::vector<std::vector<int>> fams;  // original family indices of every shard
    int F = 0, P = 0;
};

int32_t whale_set_devices(int32_t n, const int32_t* ids) {
    if (n < 0 || (n > 0 && !ids)) return fail(WHALE_ERR_ARG, "bad device list");
    const int have = whale_device_count();
    for (int i = 0; i < n; i++)
        if (ids[i] < 0 || ids[i] >= have) return fail(WHALE_ERR_ARG, "device %d not available (%d devices)", ids[i], have);
    g_devices.assign(ids, ids + n);
    return WHALE_OK;
}

int32_t whale_multi_destroy(whale_multi_t h) {
    if (!h) return WHALE_OK;
    for (whale_data* d : h->datas) whale_data_destroy(d);
    for (whale_model* m : h->models) whale_model_destroy(m);
    delete h;
    return WHALE_OK;
}

int32_t whale_multi_create(const whale_model_desc* md, const whale_ccd_desc* cd, whale_multi_t* out) {
    if (!md || !cd || !out) return fail(WHALE_ERR_ARG, "null argument");
    std::vector<int> devs = g_devices;
    if (devs.empty()) devs.push_back(g_device);
    const int nd = (int)devs.size(), F = cd->n_fam, nn = md->n_nodes;
    if (F < nd) return fail(WHALE_ERR_ARG, "%d families for %d devices", F, nd);
    // LPT on a pack-independent work proxy: triples + clades of the family
    std::vector<double> w(F);
    for (int f = 0; f < F; f++) {
        const int64_t c0 = cd->clade_off[f], c1 = cd->clade_off[f + 1];
        w[f] = (double)(cd->split_off[c1] - cd->split_off[c0]) + (double)(c1 - c0);
    }
    std::vector<int> order(F);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w[a] > w[b]; });
    auto* H = new whale_multi();
    H->F = F; H->P = md->n_params;
    H->fams.assign(nd, {});
    std::vector<double> load(nd, 0.0);
    for (int f : order) {
        int best = 0;
        for (int i = 1; i < nd; i++) if (load[i] < load[best]) best = i;
        H->fams[best].push_back(f);
        load[best] += w[f];
    }
    const int saved = g_device;
    for (int i = 0; i < nd; i++) {
        std::sort(H->fams[i].begin(), H->fams[i].end());  // family order inside a shard = original order
        // the shard's CSR arrays
        std::vector<int64_t> clade_off(1, 0), split_off(1, 0), compat_off(1, 0);
        std::vector<int32_t> nleaf, g1, g2, compat;
        std::vector<double> p;
        for (int f : H->fams[i]) {
            const int64_t c0 = cd->clade_off[f], c1 = cd->clade_off[f + 1];
            const int64_t s0 = cd->split_off[c0], s1 = cd->split_off[c1];
            const int64_t sb = (int64_t)g1.size();
            for (int64_t c = c0; c < c1; c++) { nleaf.push_back(cd->clade_nleaf[c]); split_off.push_back(sb + cd->split_off[c + 1] - s0); }
            g1.insert(g1.end(), cd->g1 + s0, cd->g1 + s1);
            g2.insert(g2.end(), cd->g2 + s0, cd->g2 + s1);
            p.insert(p.end(), cd->p + s0, cd->p + s1);
            clade_off.push_back(clade_off.back() + (c1 - c0));
            const int64_t k0 = cd->compat_off[(int64_t)f * nn], cb = (int64_t)compat.size();
            for (int e = 0; e < nn; e++) compat_off.push_back(cb + cd->compat_off[(int64_t)f * nn + e + 1] - k0);
            compat.insert(compat.end(), cd->compat + k0, cd->compat + cd->compat_off[(int64_t)(f + 1) * nn]);
        }
        whale_ccd_desc sd{(int32_t)H->fams[i].size(), clade_off.data(), nleaf.data(), split_off.data(), g1.data(), g2.data(),
                          p.data(), compat_off.data(), compat.data()};
        whale_model* m = nullptr;
        whale_data* d = nullptr;
        int32_t rc = whale_set_device(devs[i]);
        if (rc == WHALE_OK) rc = whale_model_create(md, &m);
        if (rc == WHALE_OK) { H->models.push_back(m); rc = whale_data_create(m, &sd, &d); }
        if (rc == WHALE_OK) H->datas.push_back(d);
        if (rc != WHALE_OK) { whale_multi_destroy(H); whale_set_device(saved); return rc; }
    }
    whale_set_device(saved);
    *out = H;
    return WHALE_OK;
}

int32_t whale_multi_ndev(whale_multi_t h) { return h ? (int32_t)h->models.size() : 0; }
int32_t whale_multi_shard_size(whale_multi_t h, int32_t i) {
    return (h && i >= 0 && i < (int)h->fams.size()) ? (int32_t)h->fams[i].size() : -1;
}

int32_t whale_multi_logpdf_grad(whale_multi_t h, const double* x, const double* p_leaf, int32_t condition, uint32_t flags,
                                double* loglik, double* grad, double* ll_fam) {
    if (!h || !x || !loglik) return fail(WHALE_ERR_ARG, "null argument");
    if (grad && !(flags & WHALE_WANT_GRAD)) return fail(WHALE_ERR_ARG, "grad requested without WHALE_WANT_GRAD");
    if (flags & (WHALE_KEEP_ELL | WHALE_PROFILE)) return fail(WHALE_ERR_ARG, "keep_ell / profile: use the per-device handles");
    const int nd = (int)h->models.size(), P = h->P;
    // enqueue everything on every device before waiting for any
    for (int i = 0; i < nd; i++) {
        whale_model* m = h->models[i];
        const int nn = m->nn;
        CU(cudaSetDevice(m->device));
        double* hp = m->h_pin;
        memcpy(hp, x, P * sizeof(double));
        for (int e = 0; e < nn; e++) hp[P + e] = p_leaf ? p_leaf[e] : 0.0;
        CU(upload_theta(m, hp));
        int32_t rc = enqueue_eval(m, h->datas[i], m->d_x, condition, flags, m->d_out, m->stream);
        if (rc != WHALE_OK) return rc;
        CU(cudaMemcpyAsync(hp + P + nn, m->d_out, (1 + P) * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    }
    double tot = 0.0;
    bool finite = true;
    if (grad) for (int k = 0; k < P; k++) grad[k] = 0.0;
    for (int i = 0; i < nd; i++) {  // fixed order: device-list order
        whale_model* m = h->models[i];
        CU(cudaSetDevice(m->device));
        CU(cudaStreamSynchronize(m->stream));
        m->x_host_valid = true;
        const double* ho = m->h_pin + P + m->nn;
        finite = finite && std::isfinite(ho[0]);
        tot += ho[0];
        if (grad) for (int k = 0; k < P; k++) grad[k] += ho[1 + k];
    }
    finite = finite && std::isfinite(tot);  // ℓhood src/core.jl:15 on the total
    *loglik = finite ? tot : -INFINITY;
    if (grad && !finite) for (int k = 0; k < P; k++) grad[k] = 0.0;
    if (ll_fam) {
        const bool g_ = (flags & WHALE_WANT_GRAD) != 0;
        for (int i = 0; i < nd; i++) {
            whale_data* d = h->datas[i];
            whale_model* m = h->models[i];
            CU(cudaSetDevice(m->device));
            const size_t g = g_ ? 1 : 0;
            const int KR = d->plans[g]->K[m->root];
            std::vector<double> tmp((size_t)d->F * KR);
            CU(cudaMemcpy(tmp.data(), d->d_out_fam + d->out_off[g], tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int f = 0; f < d->F; f++) ll_fam[h->fams[i][f]] = tmp[(size_t)f * KR];
        }
    }
    return WHALE_OK;
}

// ------------------------------------------------------------------------------------------------
// One process per GPU: the sum over ranks through peer memory (see k_peer_sum in whale_reduce.cuh).
// ------------------------------------------------------------------------------------------------
int32_t whale_peer_export(whale_data_t d, int32_t rank, int32_t world, void* handle64) {
    if (!d || !handle64) return fail(WHALE_ERR_ARG, "null argument");
    if (world < 1 || world > 16 || rank < 0 || rank >= world) return fail(WHALE_ERR_ARG, "rank %d of %d (at most 16 ranks)", rank, world);
    whale_model* m = d->m;
    CU(cudaSetDevice(m->device));
    const size_t n = 1 + (size_t)m->P;
    const size_t bytes = peer_buf_bytes(world, (int)n);
    if (!d->peer_bufs[rank] || d->peer_world != world || d->peer_rank != rank) {
        if (d->peer_rank >= 0 && d->peer_bufs[d->peer_rank]) cudaFree(d->peer_bufs[d->peer_rank]);
        memset(d->peer_bufs, 0, sizeof(d->peer_bufs));
        memset(d->peer_open, 0, sizeof(d->peer_open));
        CU(cudaMalloc((void**)&d->peer_bufs[rank], bytes));
        CU(cudaMemset(d->peer_bufs[rank], 0, bytes));
        if (!d->d_peer) CU(cudaMalloc((void**)&d->d_peer, sizeof(PeerDev)));
        d->peer_rank = rank; d->peer_world = world; d->peer_uploaded = false;
    }
#ifndef WHALE_EMU
    cudaIpcMemHandle_t hnd;
    CU(cudaIpcGetMemHandle(&hnd, d->peer_bufs[rank]));
    static_assert(sizeof(hnd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &hnd, 64);
#else
    memset(handle64, 0, 64);
    memcpy(handle64, &d->peer_bufs[rank], sizeof(void*));
#endif
    return WHALE_OK;
}

int32_t whale_peer_import(whale_data_t d, int32_t peer, const void* handle64) {
    if (!d || !handle64) return fail(WHALE_ERR_ARG, "null argument");
    if (d->peer_rank < 0) return fail(WHALE_ERR_STATE, "whale_peer_export first");
    if (peer < 0 || peer >= d->peer_world) return fail(WHALE_ERR_ARG, "peer %d of %d", peer, d->peer_world);
    if (peer == d->peer_rank) return WHALE_OK;
    CU(cudaSetDevice(d->m->device));
#ifndef WHALE_EMU
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle64, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
    d->peer_bufs[peer] = (double*)p;
    d->peer_open[peer] = true;
    d->peer_uploaded = false;
#else
    memcpy(&d->peer_bufs[peer], handle64, sizeof(void*));
#endif
    return WHALE_OK;
}

int32_t whale_peer_ready(whale_data_t d) {
    if (!d || d->peer_rank < 0) return 0;
    for (int q = 0; q < d->peer_world; q++) if (!d->peer_bufs[q]) return 0;
    return 1;
}

// the device-resident exchange state, uploaded once every peer's buffer is known
static int32_t peer_state(whale_data* D, PeerDev** out) {
    if (!whale_peer_ready(D)) return fail(WHALE_ERR_STATE, "WHALE_PEER_SUM without a complete whale_peer_export / whale_peer_import exchange");
    if (!D->peer_uploaded) {
        PeerDev h{};
        for (int q = 0; q < D->peer_world; q++) h.bufs[q] = D->peer_bufs[q];
        h.rank = D->peer_rank; h.world = D->peer_world; h.n = 1 + D->m->P; h.status = 0; h.seq = 0;
        h.timeout_cycles = (long long)std::max(1, env_int("WHALE_PEER_TIMEOUT_S", 60)) * 2000000000LL;  // ~2 GHz SM clock
        CU(cudaMemcpy(D->d_peer, &h, sizeof(h), cudaMemcpyHostToDevice));
        D->peer_uploaded = true;
    }
    *out = D->d_peer;
    return WHALE_OK;
}

// enqueue the exchange behind an evaluation that left its local (1+P) result in d_out
static int32_t enqueue_peer_sum(whale_data* D, double* d_out, cudaStream_t st) {
    PeerDev* pd = nullptr;
    int32_t rc = peer_state(D, &pd);
    if (rc != WHALE_OK) return rc;
    LAUNCH(k_peer_sum, 1, 128, peer_smem_bytes(D->peer_world, 1 + D->m->P), st, pd, d_out);
    g_launches++;
    CU(cudaGetLastError());
    return WHALE_OK;
}

int32_t whale_peer_sum_async(whale_data_t d, double* d_out, void* stream) {
    if (!d || !d_out) return fail(WHALE_ERR_ARG, "null argument");
    CU(cudaSetDevice(d->m->device));
    return enqueue_peer_sum(d, d_out, (cudaStream_t)stream);
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
