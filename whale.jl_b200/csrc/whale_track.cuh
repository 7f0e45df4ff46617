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
};

__device__ __forceinline__ double mul3(double a, double b, double c) { return __dmul_rn(__dmul_rn(a, b), c); }

__global__ void __launch_bounds__(128) k_backtrack(BTArgs A) {
    const long long wl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (wl >= (long long)A.nfam * A.nsamp) return;
    const int fam = (int)(wl / A.nsamp);
    const long long w = (long long)fam * A.samp_total + A.samp_off + (wl - (long long)fam * A.nsamp);
    const ModelDev& M = A.M;
    const PlanDev& PL = A.PL;
    const FamHdr* Hp = A.hdr + fam;
    const unsigned char* blob = A.arena + Hp->base;
    const NodeRec* nrec = reinterpret_cast<const NodeRec*>(blob);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(blob);
    const Ent* ents = reinterpret_cast<const Ent*>(blob);
    const double* ellf = A.ell + Hp->ell_off;
    const double* U = A.uniforms + w * A.stride;
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
    auto eps_last = [&](int e) -> double { return PL.eps[PL.toff[e] + M.nsl[e]]; };  // plan 0: K = 1

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
            double r = __dmul_rn(U[used++], L(e, c, 0));
            const int f = M.child0[e], h = M.child1[e];
            const int lf_ = M.nsl[f];
            if (kind == WHALE_WGD) {  // :258-266
                const double q = A.x[M.q_slot[e]];
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
                    const double eta = A.x[M.eta_slot];
                    const double eps = PL.eps[PL.toff[e]];
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
                double r = __dmul_rn(U[used++], L(e, c, t));
                const double2 pp = PL.pp[PL.toff[e] + t];
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
