// ORACLE — fast CPU restatement of Whale.jl's ALE/DLWGD hot path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library; the product (libwhalecuda) never links or calls it.
//
// Parity status: PINNED through oracle/whale_oracle.py (which reproduces test/runtests.jl:19 and
// :32-34 exactly); tests/test_oracle.py checks this library against both known answers too.
//
// One scalar-templated implementation serves `double` (logpdf) and `Dual<N>` (ForwardDiff.jl-style
// forward-mode dual numbers, run in chunks of <= 12 partials exactly like ForwardDiff's chunked
// gradient, test/runtests.jl:36-42).  The loops follow the Julia source line by line
// (paths relative to /root/reference):
//   slice tables  src/model.jl:162-191, src/bdputil.jl:6-11, src/rmodels.jl:14,31-33,55-64
//   DP            src/core.jl:29-64,83-199  (getl: src/ccd.jl:41-49)
//   conditioning  src/condition.jl:11-29, src/bdputil.jl:67
//   backtracking  src/track.jl:190-414
// Families are threaded like the reference (Threads.@threads, src/core.jl:60) with OpenMP.
//
// All indices are 0-based here; the Python side (oracle/whale_oracle.py, oracle/flat.py) flattens
// the reference-layout structures (clades sorted by size, triples in file order, compat lists).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {
struct OModel {
    int32_t n_nodes;
    const int32_t* order;     // node ids in processing order
    const int32_t* child0;    // -1 if none
    const int32_t* child1;    // -1 if none (wgd nodes have one child)
    const int32_t* kind;      // 0 leaf, 1 internal, 2 wgd, 3 root
    const int32_t* nslices;   // n (rows = n+1)
    const double* dt;         // slice length per node (0 for root)
    const double* leafP;      // src/model.jl:110-111
    const double* pleaf;      // sampling-failure probability for leaves (eps row 1), src/rmodels.jl:14
    const int32_t* lam_slot;  // index into the raw parameter vector, or -1 (NaN rates)
    const int32_t* mu_slot;
    const int32_t* q_slot;    // wgd nodes only
    int32_t eta_slot;
    int32_t log_scale;        // 1: DLWGD (rates = exp(raw)), 0: ConstantDLWGD
    int32_t condition;        // 0 none, 1 root, 2 nonextinct
};
struct OFams {
    int32_t n_fam;
    const int64_t* clade_off;   // [n_fam+1] into per-clade arrays
    const int32_t* clade_nleaf; // number of gene leaves per clade
    const int64_t* split_off;   // [total_clades+1] CSR into g1/g2/p
    const int32_t* g1;
    const int32_t* g2;
    const double* p;
    const int64_t* compat_off;  // [n_fam*n_nodes+1] CSR into compat
    const int32_t* compat;      // ascending local clade ids
};
}

namespace {

constexpr double LMATOL = 1e-6;  // src/bdputil.jl:3

template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0) { for (int i = 0; i < N; i++) d[i] = 0; }
    Dual(double x) : v(x) { for (int i = 0; i < N; i++) d[i] = 0; }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v / b.v; for (int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double b, const Dual<N>& a) { return a + b; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double b, const Dual<N>& a) { Dual<N> r; r.v = b - a.v; for (int i = 0; i < N; i++) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b; return r; }
template <int N> inline Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v / b; for (int i = 0; i < N; i++) r.d[i] = a.d[i] / b; return r; }
template <int N> inline Dual<N> operator/(double b, const Dual<N>& a) { return Dual<N>(b) / a; }
template <int N> inline Dual<N> xexp(const Dual<N>& a) { Dual<N> r; r.v = std::exp(a.v); for (int i = 0; i < N; i++) r.d[i] = r.v * a.d[i]; return r; }
template <int N> inline Dual<N> xlog(const Dual<N>& a) { Dual<N> r; r.v = std::log(a.v); for (int i = 0; i < N; i++) r.d[i] = a.d[i] / a.v; return r; }
inline double xexp(double a) { return std::exp(a); }
inline double xlog(double a) { return std::log(a); }
template <int N> inline double val(const Dual<N>& a) { return a.v; }
inline double val(double a) { return a; }

// ---- slice tables: per node rows 0..n of (eps, phi, psi) ----
template <class T>
struct Tables {
    std::vector<std::vector<T>> eps, phi, psi;
    std::vector<T> q;  // per node (wgd only)
    T eta;
};

template <class T>
inline T getalpha(const T& lam, const T& mu, double t) {  // src/bdputil.jl:6-7
    if (std::fabs(val(lam) - val(mu)) <= LMATOL) return lam * t / (1.0 + lam * t);
    T e = xexp(t * (lam - mu));
    return mu * (e - 1.0) / (lam * e - mu);
}

template <class T>
void setmodel(const OModel& m, const T* x, Tables<T>& tb) {  // src/model.jl:162-191
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    int nn = m.n_nodes;
    tb.eps.assign(nn, {}); tb.phi.assign(nn, {}); tb.psi.assign(nn, {});
    tb.q.assign(nn, T(NaN));
    tb.eta = x[m.eta_slot];
    for (int k = 0; k < nn; k++) {
        int e = m.order[k];
        int rows = m.nslices[e] + 1;
        tb.eps[e].assign(rows, T(1.0)); tb.phi[e].assign(rows, T(1.0)); tb.psi[e].assign(rows, T(1.0));
        // getθ: src/rmodels.jl:31-33,55-64 (slots already resolve nonwgdchild / constant rates)
        T lam = m.lam_slot[e] < 0 ? T(NaN) : (m.log_scale ? xexp(x[m.lam_slot[e]]) : x[m.lam_slot[e]]);
        T mu = m.mu_slot[e] < 0 ? T(NaN) : (m.log_scale ? xexp(x[m.mu_slot[e]]) : x[m.mu_slot[e]]);
        if (m.kind[e] == 2) {  // setwgdnode! :175-180
            T q = x[m.q_slot[e]];
            tb.q[e] = q;
            T ec = tb.eps[m.child0[e]].back();
            tb.eps[e][0] = q * (ec * ec) + (1.0 - q) * ec;
        } else if (m.kind[e] == 0) {  // setnode! :167-173
            tb.eps[e][0] = T(m.pleaf[e]);
        } else {
            tb.eps[e][0] = tb.eps[m.child0[e]].back() * tb.eps[m.child1[e]].back();
        }
        for (int i = 1; i < rows; i++) {  // setslices! :182-191
            T a = getalpha(lam, mu, m.dt[e]);
            T b = (lam / mu) * a;
            T ep = tb.eps[e][i - 1];
            T den = 1.0 - b * ep;
            tb.eps[e][i] = (a + (1.0 - a - b) * ep) / den;
            tb.phi[e][i] = (1.0 - a) * (1.0 - b) / (den * den);
            tb.psi[e][i] = (1.0 - a) * (1.0 - b) * b / (den * den * den);
        }
    }
}

template <class T>
inline T geompgf(const T& p, const T& s) { return p * s / (1.0 - (1.0 - p) * s); }  // src/bdputil.jl:67

template <class T>
T condition(const OModel& m, const Tables<T>& tb, bool& neginf) {  // src/condition.jl:11-29
    neginf = false;
    if (m.condition == 0) return T(0.0);
    int r = m.order[m.n_nodes - 1];
    if (m.condition == 2) return xlog(1.0 - geompgf(tb.eta, tb.eps[r].back()));
    int f = m.child0[r], g = m.child1[r];
    T er = geompgf(tb.eta, tb.eps[r].back());
    T ef = geompgf(tb.eta, tb.eps[f].back());
    T eg = geompgf(tb.eta, tb.eps[g].back());
    T p = 1.0 - ef - eg + er;
    if (val(p) > 0.0) return xlog(p);
    neginf = true;
    return T(0.0);
}

// ---- one family's view of the flattened CCD arena ----
struct Fam {
    int G;
    const int32_t* nleaf;
    const int64_t* soff;  // G+1 (absolute offsets)
    const int32_t *g1, *g2;
    const double* p;
    const int64_t* coff;  // n_nodes+1 (absolute offsets)
    const int32_t* compat;
};

inline Fam getfam(const OFams& fs, int f, int nn) {
    Fam x;
    int64_t c0 = fs.clade_off[f];
    x.G = (int)(fs.clade_off[f + 1] - c0);
    x.nleaf = fs.clade_nleaf + c0;
    x.soff = fs.split_off + c0;
    x.g1 = fs.g1; x.g2 = fs.g2; x.p = fs.p;
    x.coff = fs.compat_off + (int64_t)f * nn;
    x.compat = fs.compat;
    return x;
}

template <class T>
struct Ell {  // ℓ[e] : rows x cols, plus index[γ][e] (src/ccd.jl:59-75)
    std::vector<std::vector<T>> m;
    std::vector<int> cols;
    std::vector<int32_t> index;  // G x nn, -1 = incompatible
    int nn;
    inline T get(int e, int g, int row) const {  // getl src/ccd.jl:41-49
        int i = index[(size_t)g * nn + e];
        return i < 0 ? T(0.0) : m[e][(size_t)row * cols[e] + i];
    }
};

template <class T>
T family_L(const OModel& m, const Tables<T>& tb, const Fam& x, Ell<T>& L) {
    int nn = m.n_nodes;
    L.nn = nn;
    L.m.assign(nn, {});
    L.cols.assign(nn, 0);
    L.index.assign((size_t)x.G * nn, -1);
    for (int e = 0; e < nn; e++) {
        int C = (int)(x.coff[e + 1] - x.coff[e]);
        L.cols[e] = C;
        for (int j = 0; j < C; j++) L.index[(size_t)x.compat[x.coff[e] + j] * nn + e] = j;
        L.m[e].assign((size_t)(m.nslices[e] + 1) * C, T(0.0));
    }
    auto last = [&](int e) { return m.nslices[e]; };
    for (int k = 0; k < nn; k++) {
        int e = m.order[k];
        int C = L.cols[e];
        int rows = m.nslices[e] + 1;
        T* M = L.m[e].data();
        int kind = m.kind[e];
        if (kind == 3) {  // whaleroot! src/core.jl:130-149
            int f = m.child0[e], g = m.child1[e];
            T eta = tb.eta;
            T eps = tb.eps[e].back();
            T xi = 1.0 - (1.0 - eta) * eps;
            for (int c = 0; c < x.G; c++) {
                bool leaf = x.nleaf[c] == 1;
                T a(0.0), b(0.0);
                T cc = L.get(f, c, last(f)) * tb.eps[g].back() + L.get(g, c, last(g)) * tb.eps[f].back();
                if (!leaf) {
                    for (int64_t t = x.soff[c]; t < x.soff[c + 1]; t++) {
                        a = a + x.p[t] * L.get(e, x.g1[t], 0) * L.get(e, x.g2[t], 0);
                    }
                    for (int64_t t = x.soff[c]; t < x.soff[c + 1]; t++) {
                        b = b + x.p[t] * (L.get(f, x.g1[t], last(f)) * L.get(g, x.g2[t], last(g)) +
                                          L.get(g, x.g1[t], last(g)) * L.get(f, x.g2[t], last(f)));
                    }
                }
                M[c] = (1.0 - eta) * xi * a / eta + (eta * (1.0 - eps) / (xi * xi)) * (b + cc);
            }
            continue;
        }
        for (int j = 0; j < C; j++) {
            int c = x.compat[x.coff[e] + j];
            bool leaf = x.nleaf[c] == 1;
            if (kind == 2) {  // whalewgd! src/core.jl:103-119
                int f = m.child0[e];
                T q = tb.q[e];
                T p(0.0);
                if (!leaf) {
                    T s(0.0);
                    for (int64_t t = x.soff[c]; t < x.soff[c + 1]; t++)
                        s = s + x.p[t] * L.get(f, x.g1[t], last(f)) * L.get(f, x.g2[t], last(f));
                    p = p + s * q;
                }
                T lf = L.get(f, c, last(f));
                p = p + ((1.0 - q) * lf + 2.0 * q * tb.eps[f].back() * lf);
                M[j] = p;
            } else if (kind == 0) {  // whale! leaf branch src/core.jl:93-94
                if (leaf) M[j] = T(m.leafP[e]);
            } else {  // whale! internal branch src/core.jl:95-98
                int f = m.child0[e], g = m.child1[e];
                T s(0.0);
                for (int64_t t = x.soff[c]; t < x.soff[c + 1]; t++)
                    s = s + x.p[t] * (L.get(f, x.g1[t], last(f)) * L.get(g, x.g2[t], last(g)) +
                                      L.get(g, x.g1[t], last(g)) * L.get(f, x.g2[t], last(f)));
                T lo = L.get(f, c, last(f)) * tb.eps[g].back() + L.get(g, c, last(g)) * tb.eps[f].back();
                M[j] = M[j] + (s + lo);
            }
            for (int i = 1; i < rows; i++) {  // within_branch! src/core.jl:121-128
                T v = M[(size_t)i * C + j] + tb.phi[e][i] * M[(size_t)(i - 1) * C + j];
                if (!leaf) {
                    T s(0.0);
                    for (int64_t t = x.soff[c]; t < x.soff[c + 1]; t++)
                        s = s + x.p[t] * L.get(e, x.g1[t], i - 1) * L.get(e, x.g2[t], i - 1);
                    v = v + tb.psi[e][i] * s;
                }
                M[(size_t)i * C + j] = v;
            }
        }
    }
    int r = m.order[nn - 1];
    return L.m[r][L.cols[r] - 1];  // ℓ[root][1,end] src/core.jl:35
}

inline int nthreads_or_default(int n) {
#ifdef _OPENMP
    return n > 0 ? n : omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}

// value-only pass
double run_value(const OModel& m, const OFams& fs, const double* x, double* ll_fam, int nthreads) {
    Tables<double> tb;
    setmodel<double>(m, x, tb);
    int F = fs.n_fam;
    std::vector<double> ll(F);
#pragma omp parallel num_threads(nthreads_or_default(nthreads))
    {
        Ell<double> L;
#pragma omp for schedule(dynamic)
        for (int f = 0; f < F; f++) {
            Fam fx = getfam(fs, f, m.n_nodes);
            double Lr = family_L<double>(m, tb, fx, L);
            ll[f] = Lr > 0.0 ? std::log(Lr) : -INFINITY;  // src/core.jl:36
        }
    }
    double tot = 0.0;
    for (int f = 0; f < F; f++) { tot += ll[f]; if (ll_fam) ll_fam[f] = ll[f]; }
    bool ninf;
    double c = condition<double>(m, tb, ninf);
    if (ninf) c = -INFINITY;
    tot -= F * c;
    return std::isfinite(tot) ? tot : -INFINITY;  // ℓhood src/core.jl:15
}

// one ForwardDiff-style chunk: partials for raw parameters [p0, p0+N)
template <int N>
void run_chunk(const OModel& m, const OFams& fs, const double* x, int P, int p0, int np, double* grad,
               double* grad_fam, int nthreads) {
    using D = Dual<N>;
    std::vector<D> xd(P);
    for (int i = 0; i < P; i++) { xd[i] = D(x[i]); }
    for (int j = 0; j < np; j++) xd[p0 + j].d[j] = 1.0;
    Tables<D> tb;
    setmodel<D>(m, xd.data(), tb);
    int F = fs.n_fam;
    std::vector<double> g((size_t)F * N, 0.0);
#pragma omp parallel num_threads(nthreads_or_default(nthreads))
    {
        Ell<D> L;
#pragma omp for schedule(dynamic)
        for (int f = 0; f < F; f++) {
            Fam fx = getfam(fs, f, m.n_nodes);
            D Lr = family_L<D>(m, tb, fx, L);
            if (Lr.v > 0.0) {
                D l = xlog(Lr);
                for (int j = 0; j < N; j++) g[(size_t)f * N + j] = l.d[j];
            }
        }
    }
    bool ninf;
    D c = condition<D>(m, tb, ninf);
    for (int j = 0; j < np; j++) {
        double s = 0.0;
        for (int f = 0; f < F; f++) {
            s += g[(size_t)f * N + j];
            if (grad_fam) grad_fam[(size_t)f * P + p0 + j] = g[(size_t)f * N + j];
        }
        grad[p0 + j] = s - (ninf ? 0.0 : F * c.d[j]);
    }
}

template <int N>
void dispatch_chunk(int n, const OModel& m, const OFams& fs, const double* x, int P, int p0, int np,
                    double* grad, double* grad_fam, int nthreads) {
    if constexpr (N == 0) { (void)n; }
    else {
        if (n == N) run_chunk<N>(m, fs, x, P, p0, np, grad, grad_fam, nthreads);
        else dispatch_chunk<N - 1>(n, m, fs, x, P, p0, np, grad, grad_fam, nthreads);
    }
}

}  // namespace

extern "C" {

// Σ_i logpdf_i − N·condition (src/core.jl:58-64); per-family unconditioned ℓ_i in ll_fam (nullable).
// If grad != NULL also the ForwardDiff-style gradient wrt the P raw parameters (chunked, <= 12).
double oracle_logpdf(const OModel* m, const OFams* fs, const double* x, int32_t P, double* ll_fam,
                     double* grad, double* grad_fam, int32_t nthreads) {
    double tot = run_value(*m, *fs, x, ll_fam, nthreads);
    if (grad) {
        // ForwardDiff.pickchunksize: P <= 12 -> P; else ceil(P / ceil(P/12))
        int nchunks = (P + 11) / 12;
        int chunk = (P + nchunks - 1) / nchunks;
        for (int p0 = 0; p0 < P; p0 += chunk) {
            int np = P - p0 < chunk ? P - p0 : chunk;
            dispatch_chunk<12>(chunk, *m, *fs, x, P, p0, np, grad, grad_fam, nthreads);
        }
        if (!std::isfinite(tot)) for (int i = 0; i < P; i++) grad[i] = 0.0;
    }
    return tot;
}

// slice tables (eps, phi, psi) concatenated over nodes in id order, rows 0..n each
void oracle_slices(const OModel* m, const double* x, double* eps, double* phi, double* psi) {
    Tables<double> tb;
    setmodel<double>(*m, x, tb);
    size_t o = 0;
    for (int e = 0; e < m->n_nodes; e++)
        for (int i = 0; i <= m->nslices[e]; i++, o++) { eps[o] = tb.eps[e][i]; phi[o] = tb.phi[e][i]; psi[o] = tb.psi[e][i]; }
}

// full ℓ of one family (logpdf! src/core.jl:29): out = concatenation over nodes (id order) of
// row-major (n_e+1) x C_e matrices; returns log L
double oracle_ell(const OModel* m, const OFams* fs, int32_t fam, const double* x, double* out) {
    Tables<double> tb;
    setmodel<double>(*m, x, tb);
    Ell<double> L;
    Fam fx = getfam(*fs, fam, m->n_nodes);
    double Lr = family_L<double>(*m, tb, fx, L);
    size_t o = 0;
    for (int e = 0; e < m->n_nodes; e++) {
        std::memcpy(out + o, L.m[e].data(), L.m[e].size() * sizeof(double));
        o += L.m[e].size();
    }
    return Lr > 0.0 ? std::log(Lr) : -INFINITY;
}

// backtrack(wm, ccd) src/track.jl:190-414 for one family with an explicit uniform stream.
// nodes out: gamma (clade id, -1 for loss nodes), e, t (1-based row index, 0 for loss), parent (-1 root).
// returns number of nodes (>=0) or -1 on "Backtracking failed" (src/track.jl:150) / -2 overflow.
int32_t oracle_backtrack(const OModel* mp, const OFams* fs, int32_t fam, const double* x, const double* u,
                         int32_t n_u, int32_t max_nodes, int32_t* o_gamma, int32_t* o_e, int32_t* o_t,
                         int32_t* o_parent, int32_t* n_used) {
    const OModel& m = *mp;
    Tables<double> tb;
    setmodel<double>(m, x, tb);
    Ell<double> L;
    Fam c = getfam(*fs, fam, m.n_nodes);
    family_L<double>(m, tb, c, L);
    struct St { int e, g, t, node; };
    std::vector<St> stack;
    int nn = 0, used = 0;
    int root = m.order[m.n_nodes - 1];
    auto last = [&](int e) { return m.nslices[e]; };      // 0-based last row
    auto push_node = [&](int g, int e, int t, int par) -> int {
        if (nn >= max_nodes) return -1;
        o_gamma[nn] = g; o_e[nn] = e; o_t[nn] = t; o_parent[nn] = par;
        return nn++;
    };
    push_node(c.G - 1, root, 1, -1);
    stack.push_back({root, c.G - 1, 0, 0});  // t is 0-based row here; reported 1-based
    while (!stack.empty()) {
        St s = stack.back(); stack.pop_back();
        int node = s.node;
        if (s.g != o_gamma[node] || s.e != o_e[node]) {  // src/track.jl:162-170
            node = push_node(s.g, s.e, s.g < 0 ? 0 : s.t + 1, node);
            if (node < 0) return -2;
        }
        if (s.g < 0) continue;  // loss node :206-207
        int e = s.e, g = s.g, t = s.t;
        St nxt[2]; int nnxt = 0;
        bool ok = false;
        if (t == 0) {  // inter-node :213-225
            int kind = m.kind[e];
            if (kind == 0) continue;
            if (used >= n_u) return -3;
            double r = u[used++] * L.get(e, g, t);
            if (kind == 3) {  // :245-256, :369-414
                double eta = tb.eta, eps = tb.eps[e].back();
                double xi = 1.0 - (1.0 - eta) * eps;
                int f = m.child0[e], h = m.child1[e];
                for (int64_t k = c.soff[g]; k < c.soff[g + 1] && !ok; k++) {
                    int g1 = c.g1[k], g2 = c.g2[k]; double p = c.p[k];
                    r -= p * L.get(e, g1, t) * L.get(e, g2, t) * xi * (1.0 - eta) / eta;
                    if (r < 0.0) { nxt[0] = {e, g1, t, 0}; nxt[1] = {e, g2, t, 0}; nnxt = 2; ok = true; break; }
                    r -= p * L.get(f, g1, last(f)) * L.get(h, g2, last(h)) * eta * (1.0 - eps) / (xi * xi);
                    if (r < 0.0) { nxt[0] = {f, g1, last(f), 0}; nxt[1] = {h, g2, last(h), 0}; nnxt = 2; ok = true; break; }
                    r -= p * L.get(h, g1, last(h)) * L.get(f, g2, last(f)) * eta * (1.0 - eps) / (xi * xi);
                    if (r < 0.0) { nxt[0] = {h, g1, last(h), 0}; nxt[1] = {f, g2, last(f), 0}; nnxt = 2; ok = true; break; }
                }
                if (!ok) {
                    r -= L.get(f, g, last(f)) * tb.eps[h].back() * eta * (1.0 - eps) / (xi * xi);
                    if (r < 0.0) { nxt[0] = {f, g, last(f), 0}; nxt[1] = {h, -1, 0, 0}; nnxt = 2; ok = true; }
                }
                if (!ok) {
                    r -= L.get(h, g, last(h)) * tb.eps[f].back() * eta * (1.0 - eps) / (xi * xi);
                    if (r < 0.0) { nxt[0] = {h, g, last(h), 0}; nxt[1] = {f, -1, 0, 0}; nnxt = 2; ok = true; }
                }
            } else if (kind == 2) {  // :258-266, :342-367
                double q = tb.q[e];
                int f = m.child0[e];
                r -= (1.0 - q + 2.0 * q * tb.eps[f].back()) * L.get(f, g, last(f));
                if (r < 0.0) { nxt[0] = {f, g, last(f), 0}; nnxt = 1; ok = true; }
                for (int64_t k = c.soff[g]; k < c.soff[g + 1] && !ok; k++) {
                    r -= q * c.p[k] * L.get(f, c.g1[k], last(f)) * L.get(f, c.g2[k], last(f));
                    if (r < 0.0) { nxt[0] = {f, c.g1[k], last(f), 0}; nxt[1] = {f, c.g2[k], last(f), 0}; nnxt = 2; ok = true; }
                }
            } else {  // :234-242, :304-340
                int f = m.child0[e], h = m.child1[e];
                r -= L.get(f, g, last(f)) * tb.eps[h].back();
                if (r < 0.0) { nxt[0] = {f, g, last(f), 0}; nxt[1] = {h, -1, 0, 0}; nnxt = 2; ok = true; }
                if (!ok) {
                    r -= L.get(h, g, last(h)) * tb.eps[f].back();
                    if (r < 0.0) { nxt[0] = {h, g, last(h), 0}; nxt[1] = {f, -1, 0, 0}; nnxt = 2; ok = true; }
                }
                for (int64_t k = c.soff[g]; k < c.soff[g + 1] && !ok; k++) {
                    int g1 = c.g1[k], g2 = c.g2[k]; double p = c.p[k];
                    r -= p * L.get(f, g1, last(f)) * L.get(h, g2, last(h));
                    if (r < 0.0) { nxt[0] = {f, g1, last(f), 0}; nxt[1] = {h, g2, last(h), 0}; nnxt = 2; ok = true; break; }
                    r -= p * L.get(h, g1, last(h)) * L.get(f, g2, last(f));
                    if (r < 0.0) { nxt[0] = {h, g1, last(h), 0}; nxt[1] = {f, g2, last(f), 0}; nnxt = 2; ok = true; break; }
                }
            }
        } else {  // intra-branch :268-281
            if (c.nleaf[g] == 1) { nxt[0] = {e, g, 0, 0}; nnxt = 1; ok = true; }
            else {
                if (used >= n_u) return -3;
                double r = u[used++] * L.get(e, g, t);
                r -= tb.phi[e][t] * L.get(e, g, t - 1);
                if (r < 0.0) { nxt[0] = {e, g, t - 1, 0}; nnxt = 1; ok = true; }
                for (int64_t k = c.soff[g]; k < c.soff[g + 1] && !ok; k++) {  // duplication :286-302
                    r -= c.p[k] * L.get(e, c.g1[k], t - 1) * L.get(e, c.g2[k], t - 1) * tb.psi[e][t];
                    if (r < 0.0) { nxt[0] = {e, c.g1[k], t - 1, 0}; nxt[1] = {e, c.g2[k], t - 1, 0}; nnxt = 2; ok = true; }
                }
            }
        }
        if (!ok) { if (n_used) *n_used = used; return -1; }
        for (int k = nnxt - 1; k >= 0; k--) { nxt[k].node = node; stack.push_back(nxt[k]); }
    }
    if (n_used) *n_used = used;
    return nn;
}

int32_t oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
}
