// K4 — stochastic backtracking of reconciled trees.  src/track.jl:190-414.
//
// One thread per (family, sample) walk over the ℓ kept by a WHALE_KEEP_ELL evaluation.  The walk is the
// reference's depth-first recursion with an explicit stack; every decision draws r = U·ℓ[e][t,γ] from the
// host-supplied uniform stream (exactly where the reference calls rand(): :217 and :274) and subtracts the
// event weights in the reference's order until r < 0:
//   internal  sploss(f,g) -> per triple: speciation fg, gf          (:234-242, :304-340)
//   WGD       wgdloss -> per triple: retention                      (:258-266, :342-367)
//   root      per triple: root duplication, fg, gf ; then rootloss  (:245-256, :369-414)
//   in-branch ϕ·stay -> per triple: duplication                     (:268-302)
// The packed lists drop terms whose weight is exactly 0 (an incompatible sub-clade, getl = 0); subtracting
// 0 can never make r negative, so the decisions are unchanged.  Speciation entries carry the triple's ordinal
// and a "gf" flag so the root's per-triple interleaving and the child order of the reference are preserved.
// Products use explicit round-to-nearest multiplies (no FMA contraction) in the reference's association.
#pragma once
#include "whale_common.cuh"

struct BTArgs {
    ModelDev M;
    PlanDev PL;              // value-only plan tables (ϕ, ψ, ϵ)
    const unsigned char* arena;
    const FamHdr* hdr;
    const double* ell;       // kept ℓ
    const double* x;         // raw parameters (q, η)
    const double* uniforms;  // [(f*S + s) * stride ...]
    long long stride;
    int nfam, nsamp, max_nodes;
    int samp_off, samp_total;  // this launch covers samples [samp_off, samp_off + nsamp) of samp_total per family
    int32_t* node_count;     // [F*S]
    int32_t* gamma;          // [F*S*max_nodes]
    int32_t* enode;
    int32_t* trow;
    int32_t* parent;
    int32_t* status;         // [F*S]
    int4* stack;             // [F*nsamp*max_nodes] scratch of this launch
    // general form (whale_track_sample): this launch walks the listed (family, sample) pairs — the samples whose
    // posterior draw is the θ the kept ℓ was computed for — instead of a sample range of every family
    const int2* pairs;       // nullptr: the range form above
    int npairs;
    // uniforms == nullptr: draw k of walk w is a counter-based generator's output for (seed, w, k)  (the reference calls
    // the global rand(), src/track.jl:217,274; with host-supplied uniforms the stream is explicit instead)
    unsigned long long seed;
    // batched form (whale_track_sample): the launch covers the pairs of `nslots` consecutive posterior draws; slot s owns
    // the pairs [slot_off[s], slot_off[s+1]) of this launch, its own slice tables, parameter row x + s·P and kept ℓ at
    // ell + s·ell_stride (nslots == 0: one draw, the members above)
    int nslots;
    const int* slot_off;            // [nslots+1]
    const double* const* slot_eps;  // [nslots]
    const double2* const* slot_pp;  // [nslots]
    int P;
    unsigned long long ell_stride;  // doubles
};

// counter-based uniform in [0, 1): splitmix64 of (seed, walk, draw), top 53 bits
__device__ __forceinline__ double rng_u01(unsigned long long seed, unsigned long long w, unsigned long long k) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (w + 1) + 0xD1B54A32D192ED03ull * (k + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ double mul3(double a, double b, double c) { return __dmul_rn(__dmul_rn(a, b), c); }

__global__ void __launch_bounds__(128) k_backtrack(BTArgs A) {
    const long long wl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int fam;
    long long w;
    if (A.pairs) {
        if (wl >= A.npairs) return;
        const int2 pr = A.pairs[wl];
        fam = pr.x;
        w = (long long)fam * A.samp_total + pr.y;
    } else {
        if (wl >= (long long)A.nfam * A.nsamp) return;
        fam = (int)(wl / A.nsamp);
        w = (long long)fam * A.samp_total + A.samp_off + (wl - (long long)fam * A.nsamp);
    }
    const ModelDev& M = A.M;
    const PlanDev& PL = A.PL;
    const double* eps_tab = PL.eps;
    const double2* pp_tab = PL.pp;
    const double* xv = A.x;
    const double* ell_base = A.ell;
    if (A.nslots > 0) {  // the draw this pair belongs to: last slot whose first pair is <= wl
        int lo = 0, hi = A.nslots - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((long long)A.slot_off[mid] <= wl) lo = mid; else hi = mid - 1;
        }
        eps_tab = A.slot_eps[lo];
        pp_tab = A.slot_pp[lo];
        xv = A.x + (size_t)lo * A.P;
        ell_base = A.ell + (size_t)lo * A.ell_stride;
    }
    const FamHdr* Hp = A.hdr + fam;
    const unsigned char* blob = A.arena + Hp->base;
    const NodeRec* nrec = reinterpret_cast<const NodeRec*>(blob);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(blob);
    const Ent* ents = reinterpret_cast<const Ent*>(blob);
    const double* ellf = ell_base + Hp->ell_off;
    const double* U = A.uniforms ? A.uniforms + w * A.stride : nullptr;
    auto next_u = [&](int k) -> double { return U ? U[k] : rng_u01(A.seed, (unsigned long long)w, (unsigned long long)k); };
    int32_t* o_g = A.gamma + w * A.max_nodes;
    int32_t* o_e = A.enode + w * A.max_nodes;
    int32_t* o_t = A.trow + w * A.max_nodes;
    int32_t* o_p = A.parent + w * A.max_nodes;
    int4* stk = A.stack + wl * A.max_nodes;  // scratch is per launch
    const int nn = M.nn, root = M.root;
    const uint32_t nlev = Hp->nlev;

    auto ell_node = [&](int e) -> const double* {
        size_t o = 0;
        for (int e2 = 0; e2 < e; e2++) o += (size_t)(M.nsl[e2] + 1) * nrec[e2].C;
        return ellf + o;
    };
    auto tpwords = [&](int e) -> uint32_t {
        const int kd = M.kind[e];
        uint32_t wds = (kd == WHALE_INTERNAL || kd == WHALE_ROOT) ? 3 * nrec[e].C + 1 : 0;
        if (kd == WHALE_ROOT) wds += nlev + 1;
        return (wds + 3) & ~3u;
    };
    auto cmp_of = [&](int e) -> const uint32_t* { return words + nrec[e].tptr_off + tpwords(e); };
    auto L = [&](int e, int c, int t) -> double { return ell_node(e)[(size_t)t * nrec[e].C + c]; };
    auto Llast = [&](int e, int c) -> double { return L(e, c, M.nsl[e]); };
    auto eps_last = [&](int e) -> double { return eps_tab[PL.toff[e] + M.nsl[e]]; };  // plan 0: K = 1

    int nnodes = 0, used = 0, sp = 0, st = 0;
    auto add_node = [&](int g, int e, int t, int par) -> int {
        if (nnodes >= A.max_nodes) return -1;
        o_g[nnodes] = g; o_e[nnodes] = e; o_t[nnodes] = t; o_p[nnodes] = par;
        return nnodes++;
    };
    auto push = [&](int e, int c, int t, int node) {
        if (sp < A.max_nodes) stk[sp++] = make_int4(e, c, t, node);
        else st = 2;
    };
    // BackTracker(model, ccd) :157-160
    add_node((int)Hp->G - 1, root, 1, -1);
    push(root, (int)Hp->G - 1, 0, 0);
    while (sp > 0 && st == 0) {
        const int4 s = stk[--sp];
        const int e = s.x, c = s.y, t = s.z;
        int node = s.w;
        const int g = c < 0 ? -1 : (int)cmp_of(e)[c];
        if (g != o_g[node] || e != o_e[node]) {  // b(newstate) :162-170
            node = add_node(g, e, c < 0 ? 0 : t + 1, node);
            if (node < 0) { st = 2; break; }
        }
        if (c < 0) continue;  // loss node :206-207
        const NodeRec R = nrec[e];
        const int kind = M.kind[e];
        int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
        int nnext = 0;
        if (t == 0) {  // inter-node :213-225
            if (kind == WHALE_LEAF) continue;
            if (used >= A.stride) { st = 3; break; }
            double r = __dmul_rn(next_u(used++), L(e, c, 0));
            const int f = M.child0[e], h = M.child1[e];
            const int lf_ = M.nsl[f];
            if (kind == WHALE_WGD) {  // :258-266
                const double q = xv[M.q_slot[e]];
                const double wgt = __dadd_rn(__dadd_rn(1.0, -q), mul3(2.0, q, eps_last(f)));
                r = __dadd_rn(r, -__dmul_rn(wgt, Llast(f, c)));
                if (r < 0.0) { n0 = make_int4(f, c, lf_, 0); nnext = 1; }
                const uint32_t* dptr = words + R.dptr_off;
                for (uint32_t k = dptr[c]; k < dptr[c + 1] && !nnext; k++) {
                    const Ent en = ents[R.dent_off + k];
                    r = __dadd_rn(r, -__dmul_rn(mul3(q, en.p, Llast(f, en.i1)), Llast(f, en.i2)));
                    if (r < 0.0) { n0 = make_int4(f, en.i1, lf_, 0); n1 = make_int4(f, en.i2, lf_, 0); nnext = 2; }
                }
            } else {
                const int lh_ = M.nsl[h];
                const uint32_t* tptr = words + R.tptr_off;
                const int32_t* lossF = reinterpret_cast<const int32_t*>(tptr + R.C + 1);
                const int32_t* lossG = lossF + R.C;
                const Ent* te = ents + R.tent_off;
                if (kind == WHALE_INTERNAL) {  // sploss then speciation :234-242
                    if (lossF[c] >= 0) {
                        r = __dadd_rn(r, -__dmul_rn(Llast(f, lossF[c]), eps_last(h)));
                        if (r < 0.0) { n0 = make_int4(f, lossF[c], lf_, 0); n1 = make_int4(h, -1, 0, 0); nnext = 2; }
                    }
                    if (!nnext && lossG[c] >= 0) {
                        r = __dadd_rn(r, -__dmul_rn(Llast(h, lossG[c]), eps_last(f)));
                        if (r < 0.0) { n0 = make_int4(h, lossG[c], lh_, 0); n1 = make_int4(f, -1, 0, 0); nnext = 2; }
                    }
                    for (uint32_t k = tptr[c]; k < tptr[c + 1] && !nnext; k++) {
                        const Ent en = te[k];  // (i1 in f, i2 in g); pad bit0: γ1 went to g ("gf")
                        r = __dadd_rn(r, -__dmul_rn(__dmul_rn(en.p, (en.pad & 1u) ? Llast(h, en.i2) : Llast(f, en.i1)),
                                                    (en.pad & 1u) ? Llast(f, en.i1) : Llast(h, en.i2)));
                        if (r < 0.0) {
                            if (en.pad & 1u) { n0 = make_int4(h, en.i2, lh_, 0); n1 = make_int4(f, en.i1, lf_, 0); }
                            else { n0 = make_int4(f, en.i1, lf_, 0); n1 = make_int4(h, en.i2, lh_, 0); }
                            nnext = 2;
                        }
                    }
                } else {  // root :245-256, :369-414
                    const double eta = xv[M.eta_slot];
                    const double eps = eps_tab[PL.toff[e]];
                    const double xi = __dadd_rn(1.0, -__dmul_rn(__dadd_rn(1.0, -eta), eps));
                    const double ome = __dadd_rn(1.0, -eta), omeps = __dadd_rn(1.0, -eps), xi2 = __dmul_rn(xi, xi);
                    const uint32_t* dptr = words + R.dptr_off;
                    uint32_t kt = tptr[c];
                    const uint32_t kte = tptr[c + 1];
                    for (uint32_t k = dptr[c]; k < dptr[c + 1] && !nnext; k++) {
                        const Ent en = ents[R.dent_off + k];
                        double wv = __dmul_rn(__dmul_rn(en.p, L(e, en.i1, 0)), L(e, en.i2, 0));
                        wv = __ddiv_rn(__dmul_rn(__dmul_rn(wv, xi), ome), eta);
                        r = __dadd_rn(r, -wv);
                        if (r < 0.0) { n0 = make_int4(e, en.i1, 0, 0); n1 = make_int4(e, en.i2, 0, 0); nnext = 2; break; }
                        const uint32_t j = k - dptr[c];
                        while (kt < kte && (te[kt].pad >> 1) == j && !nnext) {
                            const Ent sn = te[kt++];
                            const bool gf = sn.pad & 1u;
                            double sv = __dmul_rn(__dmul_rn(sn.p, gf ? Llast(h, sn.i2) : Llast(f, sn.i1)),
                                                  gf ? Llast(f, sn.i1) : Llast(h, sn.i2));
                            sv = __ddiv_rn(__dmul_rn(__dmul_rn(sv, eta), omeps), xi2);
                            r = __dadd_rn(r, -sv);
                            if (r < 0.0) {
                                if (gf) { n0 = make_int4(h, sn.i2, lh_, 0); n1 = make_int4(f, sn.i1, lf_, 0); }
                                else { n0 = make_int4(f, sn.i1, lf_, 0); n1 = make_int4(h, sn.i2, lh_, 0); }
                                nnext = 2;
                            }
                        }
                    }
                    if (!nnext && lossF[c] >= 0) {
                        double lv = __dmul_rn(Llast(f, lossF[c]), eps_last(h));
                        lv = __ddiv_rn(__dmul_rn(__dmul_rn(lv, eta), omeps), xi2);
                        r = __dadd_rn(r, -lv);
                        if (r < 0.0) { n0 = make_int4(f, lossF[c], lf_, 0); n1 = make_int4(h, -1, 0, 0); nnext = 2; }
                    }
                    if (!nnext && lossG[c] >= 0) {
                        double lv = __dmul_rn(Llast(h, lossG[c]), eps_last(f));
                        lv = __ddiv_rn(__dmul_rn(__dmul_rn(lv, eta), omeps), xi2);
                        r = __dadd_rn(r, -lv);
                        if (r < 0.0) { n0 = make_int4(h, lossG[c], lh_, 0); n1 = make_int4(f, -1, 0, 0); nnext = 2; }
                    }
                }
            }
        } else {  // intra-branch :268-281
            if ((uint32_t)c < R.C - R.nonleaf) {  // leaf clade: straight to the leafward end
                n0 = make_int4(e, c, 0, 0);
                nnext = 1;
            } else {
                if (used >= A.stride) { st = 3; break; }
                double r = __dmul_rn(next_u(used++), L(e, c, t));
                const double2 pp = pp_tab[PL.toff[e] + t];
                r = __dadd_rn(r, -__dmul_rn(pp.x, L(e, c, t - 1)));
                if (r < 0.0) { n0 = make_int4(e, c, t - 1, 0); nnext = 1; }
                const uint32_t* dptr = words + R.dptr_off;
                for (uint32_t k = dptr[c]; k < dptr[c + 1] && !nnext; k++) {  // duplication :286-302
                    const Ent en = ents[R.dent_off + k];
                    r = __dadd_rn(r, -__dmul_rn(__dmul_rn(__dmul_rn(en.p, L(e, en.i1, t - 1)), L(e, en.i2, t - 1)), pp.y));
                    if (r < 0.0) { n0 = make_int4(e, en.i1, t - 1, 0); n1 = make_int4(e, en.i2, t - 1, 0); nnext = 2; }
                }
            }
        }
        if (!nnext) { st = 1; break; }  // "Backtracking failed" :150
        if (nnext == 2) { n1.w = node; push(n1.x, n1.y, n1.z, n1.w); }
        n0.w = node;
        push(n0.x, n0.y, n0.z, n0.w);
    }
    A.node_count[w] = nnodes;
    A.status[w] = st;
    (void)nn;
}


// ---------------------------------------------------------------------------------------------------------
// Tree identity and the per-family summary of backtracked samples (src/track.jl:95-113 nodehash / cladehash,
// src/rectree.jl:113-133 sumtrees).  The reference identifies a reconciled tree with the SET over its nodes of
//     (γ, e, {(γ, e) of the children})        loss nodes: (loss, e, γ of the sister)
// — the slice t of an event is not part of it.  k_tree_hash forms a 64-bit hash of that set, one thread per tree:
// a node's children are folded commutatively into the parent (nodes are in creation order: parents first), each node's
// key is mixed and the keys are added up (a set hash; a (γ, e) state occurs once per tree, so set = multiset).
// k_tree_dedup sorts the n_samples hashes of one family in shared memory and writes the runs: distinct trees with
// their counts and the first sample that showed them.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long key2(int g, int e) {
    return mix64(((unsigned long long)(unsigned)(g + 2) << 32) | (unsigned long long)(unsigned)e);
}

struct HashArgs {
    const int32_t* node_count;  // [W]
    const int32_t* gamma;       // [W*max_nodes]
    const int32_t* enode;
    const int32_t* parent;
    const int32_t* status;
    int4* scratch;              // [W*max_nodes]: .x,.y = children key sum (u64), .z = Σ children γ, .w = #children
    unsigned long long* hash;   // [W]
    long long W;
    int max_nodes;
};
__global__ void __launch_bounds__(128) k_tree_hash(HashArgs A) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= A.W) return;
    const int n = A.node_count[w];
    const int32_t* g = A.gamma + w * A.max_nodes;
    const int32_t* e = A.enode + w * A.max_nodes;
    const int32_t* par = A.parent + w * A.max_nodes;
    int4* sc = A.scratch + w * A.max_nodes;
    for (int i = 0; i < n; i++) sc[i] = make_int4(0, 0, 0, 0);
    for (int i = 1; i < n; i++) {
        const int p = par[i];
        int4 v = sc[p];
        unsigned long long acc = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
        acc += key2(g[i], e[i]);
        sc[p] = make_int4((int)(unsigned)(acc & 0xffffffffull), (int)(unsigned)(acc >> 32), v.z + g[i], v.w + 1);
    }
    unsigned long long h = 0x243F6A8885A308D3ull + (unsigned long long)n;
    for (int i = 0; i < n; i++) {
        unsigned long long k;
        if (g[i] < 0) {  // loss node: (loss, e, γ of its sister)
            const int p = par[i];
            const int sis = (p >= 0 && sc[p].w == 2) ? sc[p].z - g[i] : -2;
            k = mix64(key2(-1, e[i]) ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(unsigned)(sis + 2)));
        } else {
            const int4 v = sc[i];
            const unsigned long long acc = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
            k = mix64(key2(g[i], e[i]) + 0xC2B2AE3D27D4EB4Full * (acc + (unsigned long long)v.w));
        }
        h += mix64(k);
    }
    A.hash[w] = A.status[w] == 0 ? h : 0xFFFFFFFFFFFFFFFFull;  // failed walks share one bucket
}

// one CTA per family: sort (hash, sample) pairs (bitonic, padded to a power of two), then emit the runs
__global__ void __launch_bounds__(256) k_tree_dedup(const unsigned long long* __restrict__ hash, int S, int Spad,
                                                    int32_t* __restrict__ n_distinct, unsigned long long* __restrict__ out_hash,
                                                    int32_t* __restrict__ out_count, int32_t* __restrict__ out_first) {
    EXTERN_SHARED(dsm);
    unsigned long long* hk = reinterpret_cast<unsigned long long*>(dsm);  // [Spad]
    int* hv = reinterpret_cast<int*>(hk + Spad);                           // [Spad]
    const int f = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < Spad; i += nt) {
        hk[i] = i < S ? hash[(size_t)f * S + i] : 0xFFFFFFFFFFFFFFFFull;
        hv[i] = i < S ? i : 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= Spad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Spad; i += nt) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = hk[i], b = hk[l];
                    const int va = hv[i], vb = hv[l];
                    const bool gt = a > b || (a == b && va > vb);  // ties: the earlier sample first
                    if (gt == up) { hk[i] = b; hk[l] = a; hv[i] = vb; hv[l] = va; }
                }
            }
            __syncthreads();
        }
    // runs over the first S sorted entries (padding sorts last: its sample index is larger than any real one)
    if (tid == 0) {
        int nd = 0;
        for (int i = 0; i < S;) {
            int j = i + 1;
            while (j < S && hk[j] == hk[i]) j++;
            out_hash[(size_t)f * S + nd] = hk[i];
            out_count[(size_t)f * S + nd] = j - i;
            out_first[(size_t)f * S + nd] = hv[i];
            nd++;
            i = j;
        }
        n_distinct[f] = nd;
    }
}

// pack the walks' node rows: one warp per tree copies its node_count rows (γ, e, t, parent) to offsets[w]
__global__ void __launch_bounds__(128) k_tree_pack(const int32_t* __restrict__ node_count, const long long* __restrict__ offsets,
                                                   const int32_t* __restrict__ gamma, const int32_t* __restrict__ enode,
                                                   const int32_t* __restrict__ trow, const int32_t* __restrict__ parent,
                                                   int max_nodes, long long W, int4* __restrict__ packed) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= W) return;
    const int n = node_count[w];
    const long long o = offsets[w], b = w * max_nodes;
    for (int i = lane; i < n; i += 32) packed[o + i] = make_int4(gamma[b + i], enode[b + i], trow[b + i], parent[b + i]);
}
