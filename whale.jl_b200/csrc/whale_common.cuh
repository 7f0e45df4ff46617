// Shared device-side structures of libwhalecuda (see whalecuda.cu for the overview).
#pragma once
#ifdef WHALE_EMU
// Test-only build: tests/emu/cuda_emu.h maps the CUDA constructs used here onto host threads so the
// kernel logic can be exercised on a machine without a GPU.  Never shipped, never loaded by the package.
#include "cuda_emu.h"
#define LAUNCH(kern, grid, block, smem, st, ...) emu::launch(grid, block, smem, [=]() { kern(__VA_ARGS__); })
#define LAUNCH_PDL(kern, grid, block, smem, st, ...) LAUNCH(kern, grid, block, smem, st, __VA_ARGS__)
#define DYN_SMEM_BYTES() ((unsigned)emu::g_smem_bytes)
#define GRID_DEP_WAIT() ((void)0)
#define GRID_DEP_LAUNCH() ((void)0)
#else
#include <cuda_runtime.h>
#define LAUNCH(kern, grid, block, smem, st, ...) kern<<<grid, block, smem, st>>>(__VA_ARGS__)
#define EXTERN_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
// Programmatic dependent launch: the kernel may start while its predecessor in the stream (the table kernel) is still
// running — its prologue (species-tree metadata, node records, L2 prefetch of the family's lists) overlaps the tables;
// GRID_DEP_WAIT() (griddepcontrol.wait) blocks until the predecessor has completed and its writes are visible.
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
#define LAUNCH_PDL(kern, grid, block, smem, st, ...) launch_pdl(kern, grid, block, smem, st, __VA_ARGS__)
// dynamic shared memory of this launch (the bin's largest need: a smaller family has spare room behind its own carve-up)
static __device__ __forceinline__ unsigned dyn_smem_bytes_() {
    unsigned v;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(v));
    return v;
}
#define DYN_SMEM_BYTES() dyn_smem_bytes_()
#define GRID_DEP_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define GRID_DEP_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#endif
#include <cstdint>
#include <utility>

#include "../../include/whalecuda.h"

// one resolved clade-split term p * X[i1] * Y[i2]; indices are LOCAL cell indices of the branch row they
// address (the reference resolves index[γ,e] at run time, src/ccd.jl:41-49; the packer does it once)
struct __align__(16) Ent {
    uint16_t i1, i2;
    uint32_t pad;
    double p;
};
static_assert(sizeof(Ent) == 16, "Ent must be 16 bytes");

// per family, per species-tree node; offsets are relative to the family blob
struct NodeRec {
    uint32_t C;         // compatible clades (columns of ℓ[e])
    uint32_t nonleaf;   // how many are non-leaf clades (they come last: clades are size-sorted)
    uint32_t dptr_off;  // u32-word offset (multiple of 4) of dptr[C+1]: same-branch terms
                        //   (within-branch duplication :178-185, Πwgdretention :187-194, Πroot :151-158)
    uint32_t dent_off;  // 16-byte-entry offset of those terms
    uint32_t ndent;
    uint32_t tptr_off;  // u32-word offset (multiple of 4) of [tptr[C+1] | lossF[C] | lossG[C] | lev[nlev+1]] pad4
                        //   | cmp[C]: speciation terms at row 1 (:160-170), Πloss child indices (:172-176),
                        //   root levels, local cell -> clade id
    uint32_t tent_off;
    uint32_t ntent;
    uint32_t slot_off;  // u32-word offset (multiple of 4) of the slice-loop lane table (Slot[nslots])
    uint32_t nslots;    //   (a leaf branch with more than HEAVY_SLOTS slots is processed by the whole CTA)
    uint32_t sptr_off;  // leaf branches in shape mode (see SHAPES below): u32-word offset of sptr[C+1], else 0
    uint32_t sent_off;  //   16-byte-entry offset of {i1 = shape id, p = coefficient}
};
static_assert(sizeof(NodeRec) == 48, "NodeRec must be 48 bytes");

// Slice-loop work descriptor (built by the packer; src/core.jl:178-185 balanced over lanes): a clade cell with E
// same-branch terms is owned by a team of G = 2^glog adjacent slots (G = 1 for E <= 2) so that no lane sums
// more than two terms (E <= 64); slot j of the team sums terms first, first+G, ... (cnt of them).  At run time
// every slot is expanded into K lanes (one per component), the team reduces its scalar partial results by
// shuffles and the leader (slot index multiple of G) writes the cell component.
struct Slot {
    uint16_t cell;    // local cell index
    uint8_t glog;
    uint8_t cnt;
    uint16_t first;   // index into the node's term list
    uint16_t stride;  // = G
};
static_assert(sizeof(Slot) == 8, "Slot must be 8 bytes");
constexpr uint32_t HEAVY_SLOTS = 24;

#ifdef WHALE_EMU
constexpr int TABLES_NT = 128;  // emulation build: keep the fiber count small
#else
constexpr int TABLES_NT = 512;
#endif
// Row stride (doubles) of a branch row with K components per cell: even K is padded to K+1 so that cells do not
// collide on shared-memory banks in the slice gathers (K = 4: cells ≡ mod 4 share banks, 2.1 wavefronts per ideal one
// — profiles/r1_ncu_k_dp_v8_summary.txt; measured on the B200 in round 2: +3.2 % evals/s on C2,
// profiles/r2_ab_odd_stride.json vs r2_ab_base.json).  WHALE_EVEN_STRIDE restores the dense layout (A/B, tests).
#ifdef WHALE_EVEN_STRIDE
#define RS(K) (K)
#else
#define RS(K) ((K) | 1)
#endif
constexpr int MAXPLAN = 64;  // tangent plans per data handle: [0] value only, [1..] gradient (parameter chunks)

// Tree shapes for the closed form of leaf branches.  On a leaf branch e every leaf clade has the same
// family-independent sequence w•_i = leafℙ·Π_{j<=i} ϕ_j, hence by induction over src/core.jl:121-128,178-185 every
// clade γ made of in-paralogs satisfies ℓ_i[γ] = Σ_σ C_σ[γ]·wσ_i over rooted binary tree shapes σ with |γ| leaves,
//   wσ_i = ϕ_i wσ_{i−1} + ψ_i w^a_{i−1} w^b_{i−1}   (σ = {a, b}),     C_σ[γ] = Σ_splits p Σ_{(σ1,σ2)→σ} C_σ1[γ1] C_σ2[γ2].
// C depends only on the CCD (packer), w only on θ (leaf-shape CTAs of k_tables).  Shapes with <= 5 leaves:
//   0 •   1 (•,•)   2 (•,1)   3 (•,2)   4 (1,1)   5 (•,3)   6 (•,4)   7 (1,2)
constexpr int NSHAPE = 8;
constexpr int SHAPE_MAXLEAVES = 5;
#define SHAPE_A {0, 0, 0, 0, 1, 0, 0, 1}
#define SHAPE_B {0, 0, 1, 2, 1, 3, 4, 2}

struct FamHdr {
    uint64_t base;           // byte offset of the blob in the arena (16-byte aligned)
    uint64_t ell_off;        // offset (doubles) of this family's ℓ in the keep_ell buffer
    uint32_t G;              // clades
    uint32_t nlev;           // root levels (distinct clade sizes)
    // shared-memory budget (doubles unless stated), per tangent plan where it depends on K_e
    uint32_t rows_len[MAXPLAN];    // Σ_e C_e K_e (even)
    uint32_t scr_len[MAXPLAN];     // scratch row: max over internal/WGD nodes of C_e K_e (even)
    uint32_t prod_len[MAXPLAN];    // (unused: no products go through shared memory any more; kept 0)
    uint32_t leafmax[MAXPLAN];     // per-warp scratch row: max over leaf branches of C_e K_e (even)
    uint32_t stage_bytes[MAXPLAN]; // staging buffer for one internal node's lists + ϕ/ψ rows (bytes, multiple of 16)
    uint32_t leaf_stage;     // per-warp staging buffer for one leaf branch's lists (bytes, multiple of 16)
    uint32_t blob_bytes;
    uint32_t rootwin;        // most terms (Πroot + speciation) any one root level holds: size of a level buffer
};

// ---- reverse-mode (adjoint) gradient: the transposed lists of a family, built from its forward blob ----
// Per node, relative to the family's REVERSE blob (RevRec[nn] | u32 words | 16-byte entries).  An entry
// {i1 = π, i2 = σ, p} under cell γ' stands for the term p·ℓ̄[π]·ℓ[σ] of ℓ̄[γ']: π indexes the adjoint row of the
// node that owns the term, σ the value row of the term's other operand.
struct RevRec {
    uint32_t bptr_off;   // u32-word offset of bptr[C+1]: transposed same-branch terms (slices; WGD row 1; Πroot)
    uint32_t bent_off;   // 16-byte-entry offset of those entries (each forward term appears under both operands)
    uint32_t nbent;
    uint32_t bslot_off;  // lane table of the backward slice loop (Slot[nbslots]; non-root nodes with slices)
    uint32_t nbslots;
    uint32_t sF_off;     // internal/root: u32-word offset of [sFptr[C_F+1] | upF[C_F]]: transposed speciation terms
    uint32_t sFent_off;  //   grouped by the FIRST child's cell (π = this node's cell, σ = second child's cell);
    uint32_t nsFent;     //   upF = this node's cell holding the same clade (Πloss transposed)
    uint32_t sG_off, sGent_off, nsGent;  // same for the second child
    uint32_t hoff;       // offset (doubles) of this node's rows in the family's history: (n+1) rows of stride Cp
};
static_assert(sizeof(RevRec) == 48, "RevRec must be 48 bytes");

struct RevHdr {
    uint64_t base;          // byte offset of the reverse blob in the reverse arena (16-byte aligned)
    uint32_t blob_bytes;
    uint32_t hist_len;      // doubles: every row of every internal/WGD branch (the forward pass keeps them)
    uint32_t rows_len;      // doubles: last rows under the hybrid plan (leaf branches keep theirs until the end)
    uint32_t scr_len;       // scratch row
    uint32_t leafmax;       // per-warp scratch row of a leaf branch
    uint32_t stage_bytes;   // staging of one node's forward lists + ϕ/ψ rows, or of its backward slice lists
    uint32_t arows_len;     // doubles: adjoint last rows, placed by (reverse) lifetime
    uint32_t hbuf_len;      // doubles: one staged history row (largest padded C of an internal/WGD node)
    uint32_t stage2_bytes;  // row-1 lists of one internal/WGD node, prefetched by bulk copies (0: read in place)
    uint32_t root_staged;   // the root's lists fit stage + stage2 together (they are contiguous) and are prefetched too
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
    int n_params;
    int nlvl;              // nodes grouped by height (k_tables schedule)
    const int* lvl_off;
    const int* lvl_nodes;
    int nleafnodes;        // leaf nodes, then the remaining nodes in processing order (k_dp schedule)
    const int* leafnodes;
    int ninner;
    const int* inner;
};

struct PlanDev {  // tangent plan: which raw parameters each branch carries
    int Kmax;
    const int* K;          // [nn]
    const int* act;        // [nn*Kmax] global parameter id of component k (act[e][0] = -1)
    const int16_t* cmap;   // [nn*2*Kmax] component index in child j of parent's component k, or -1
    const uint8_t* role;   // [nn*Kmax] bit0 own λ, bit1 own μ, bit2 own q, bit3 η
    const int* toff;       // [nn] offset (doubles) of node e's table: (n_e+1) rows × K_e
    double* eps;           // [tab_len]  ϵ rows (component-major within a row)
    double2* pp;           // [tab_len]  (ϕ, ψ) rows
    double2* uv;           // [tab_len]  projective ϵ = u/v (k_tables scratch of chain-mode branches)
    double* ab;            // [nn*Kmax*2] per-branch (α, β) components
    double* cx;            // [nn*Kmax]  row-1 coefficient X (WGD: 1−q+2qϵ_f ; root: (1−η)ξ/η)
    double* cy;            // [nn*Kmax]  row-1 coefficient Y (WGD: q ; root: η(1−ϵ)/ξ²)
    double* leaf;          // [nn*Kmax]  last-row value of a leaf clade on leaf branch e
    double* shapeW;        // [nn*NSHAPE*Kmax] last-row value wσ_n of every tree shape on leaf branch e
    double* cond;          // [4*Kmax]   condition() per kind (none, root, nonextinct, nowhere), components of the root
    const int* nwL;        // [nn] leaves below node e         } NowhereExtinctCondition scratch, allocated on
    const long long* nwoff;  // [nn] offset of node e's pgf vector  } first use (k_nowhere)
    double* nwvec;         // Σ_e 2^L_e · K_e doubles
    long long* tim;        // [32] k_tables cycle stamps (profiling aid)
    // reverse-mode gradient, full plan only: k_tables3 stops after the chain over the tree's height and writes the
    // Jacobian of the local quantities the adjoint pass differentiates with respect to — jac[(e*8 + j)*KR + k], KR = K[root]
    const int16_t* rinv;   // [nn][KR] component of root component k in node e's list (−1: none)
    double* jac;
    const int* koff;       // [nn] prefix sums of K: k_tables packs its per-(node, component) shared arrays with them
    int ktot;              // Σ_e K_e
};

#ifdef WHALE_EMU
#define CLOCK64() 0LL
#else
#define CLOCK64() clock64()
#endif

// one-partial dual number: each lane carries the value and ITS tangent component
struct D1 {
    double v, d;
};
__device__ __forceinline__ D1 mk(double v, double d = 0.0) { return D1{v, d}; }
__device__ __forceinline__ D1 operator+(D1 a, D1 b) { return D1{a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ D1 operator-(D1 a, D1 b) { return D1{a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ D1 operator*(D1 a, D1 b) { return D1{a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ D1 operator/(D1 a, D1 b) {  // one reciprocal instead of two divisions (≤ 1.5 ulp)
    const double inv = 1.0 / b.v;
    const double q = a.v * inv;
    return D1{q, (a.d - q * b.d) * inv};
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
__device__ __forceinline__ D1 dpowi(D1 a, int n) {  // a^n, n >= 0
    double p = pow(a.v, (double)n);
    return D1{p, n == 0 ? 0.0 : (double)n * (p / a.v) * a.d};
}
__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

// ---------------------------------------------------------------------------------------------------------
// One-shot sum over ranks through peer memory (one process per GPU).  Every rank owns an exchange buffer
//   [2 parities][world slots][2n packets of 8 bytes]
// that all peers have mapped (CUDA IPC over NVLink).  A packet carries half a double and the step's 32-bit tag in ONE
// 64-bit store (data and flag cannot be seen apart, so no fence and no second round trip is needed — the "LL" idea of
// collective libraries): rank r stores the 2n packets of its n = 1+P doubles into slot r of EVERY rank's buffer, then polls
// the world's packets in its own buffer until they carry the tag and adds the slots in rank order — the same order on
// every rank, so all ranks hold bit-identical totals.  Parity double-buffering is enough: a rank can start step s+1 only
// after every peer has entered the exchange of step s.  The step counter lives on the device (advanced here), so the
// exchange can ride in the tail of the DP kernel (dp_tail_reduce) or run as its own launch (k_peer_sum) alike.
// ---------------------------------------------------------------------------------------------------------
struct PeerDev {
    double* bufs[16];  // exchange buffer of every rank, as mapped in THIS process (bufs[rank] = own)
    int rank, world, n;
    int status;        // set to 1 when a peer did not show up in time
    unsigned long long seq;  // exchanges completed
    long long timeout_cycles;  // how long to wait for a peer's packets (SM clock cycles; WHALE_PEER_TIMEOUT_S, default 60 s)
};
#ifdef WHALE_EMU
#define ST_RELAXED_SYS(p, v) (*(volatile unsigned long long*)(p) = (v))
#define LD_RELAXED_SYS(p) (*(volatile const unsigned long long*)(p))
#else
#define ST_RELAXED_SYS(p, v) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory")
__device__ __forceinline__ unsigned long long ld_relaxed_sys_(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
#define LD_RELAXED_SYS(p) ld_relaxed_sys_(p)
#endif
__host__ __device__ inline size_t peer_buf_bytes(int world, int n) { return (size_t)2 * world * 2 * n * sizeof(unsigned long long); }
__host__ __device__ inline size_t peer_smem_bytes(int world, int n) { return (size_t)world * 2 * n * sizeof(unsigned); }

// whole CTA (NT threads); `out` holds this rank's n doubles (visible to the CTA) and receives the world's total;
// s_w: world·2n words of shared memory
template <int NT>
__device__ __forceinline__ void peer_exchange(PeerDev* PD, double* out, unsigned* s_w) {
    const int tid = threadIdx.x, W = PD->world, n = PD->n, rank = PD->rank, n2 = 2 * n;
    const unsigned long long seq = PD->seq + 1;
    const int par = (int)(seq & 1ull);
    const unsigned tag = (unsigned)(seq % 0xFFFFFFFFull) + 1u;  // never 0 (the buffers start zeroed)
    for (int idx = tid; idx < W * n2; idx += NT) {
        const int q = idx / n2, j = idx - q * n2;
        const double v = out[j >> 1];
        const unsigned w = (j & 1) ? (unsigned)__double2hiint(v) : (unsigned)__double2loint(v);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(PD->bufs[q]) + ((size_t)par * W + rank) * n2 + j;
        ST_RELAXED_SYS(dst, ((unsigned long long)tag << 32) | w);
    }
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(PD->bufs[rank]) + (size_t)par * W * n2;
    int bad = 0;
    for (int idx = tid; idx < W * n2; idx += NT) {
        const long long t0 = CLOCK64();
        unsigned long long pk;
        while ((unsigned)((pk = LD_RELAXED_SYS(mine + idx)) >> 32) != tag) {
#ifdef WHALE_EMU
            break;  // (the emulation runs the ranks one after another: no waiting)
#endif
            if (CLOCK64() - t0 > PD->timeout_cycles) { bad = 1; break; }  // a peer is gone
        }
        s_w[idx] = (unsigned)pk;
    }
    bad = __syncthreads_or(bad);
    double t0v = 0.0;
    for (int q = 0; q < W; q++) t0v += __hiloint2double((int)s_w[q * n2 + 1], (int)s_w[q * n2]);
    const bool fin = isfinite(t0v) && !bad;  // ℓhood (src/core.jl:15) on the world's total
    for (int j = tid; j < n; j += NT) {
        double t = 0.0;
        for (int q = 0; q < W; q++) t += __hiloint2double((int)s_w[q * n2 + 2 * j + 1], (int)s_w[q * n2 + 2 * j]);
        out[j] = fin ? t : (j == 0 ? -dinf() : 0.0);
    }
    if (tid == 0) {
        PD->seq = seq;
        if (bad) PD->status = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Row placement in shared memory.  A branch's last row is live from the moment it is computed until its
// parent's row 1 has been formed (the root's children until the end), so rows are placed first-fit into the
// gaps left by rows that are already dead; the peak is ~60 % of the plain sum Σ_e C_e K_e for a 9-taxon tree.
// Runs identically on the host (budget, `set_budgets`) and on the device (thread 0 of k_dp).
// ---------------------------------------------------------------------------------------------------------
#ifndef WHALE_EMU
#define WHALE_HD
#else
#define WHALE_HD
#endif
constexpr int ROWALLOC_MAXBLK = 64;

struct RowAlloc {  // first-fit allocator over one linear region; free blocks below `top`, sorted by offset
    int foff[ROWALLOC_MAXBLK], flen[ROWALLOC_MAXBLK];
    int nfree = 0, top = 0;
    int take(int need) {
        if (need == 0) return 0;
        for (int i = 0; i < nfree; i++)
            if (flen[i] >= need) {
                const int o = foff[i];
                foff[i] += need; flen[i] -= need;
                if (flen[i] == 0) { for (int j = i + 1; j < nfree; j++) { foff[j - 1] = foff[j]; flen[j - 1] = flen[j]; } nfree--; }
                return o;
            }
        const int o = top;
        top += need;
        return o;
    }
    void give(int o, int len) {
        if (len == 0) return;
        int i = 0;
        while (i < nfree && foff[i] < o) i++;
        if (nfree >= ROWALLOC_MAXBLK) return;  // table full: leak the block (still correct, just less reuse)
        for (int j = nfree; j > i; j--) { foff[j] = foff[j - 1]; flen[j] = flen[j - 1]; }
        foff[i] = o; flen[i] = len; nfree++;
        if (i + 1 < nfree && foff[i] + flen[i] == foff[i + 1]) {  // merge with the next block
            flen[i] += flen[i + 1];
            for (int j = i + 2; j < nfree; j++) { foff[j - 1] = foff[j]; flen[j - 1] = flen[j]; }
            nfree--;
        }
        if (i > 0 && foff[i - 1] + flen[i - 1] == foff[i]) {  // merge with the previous block
            flen[i - 1] += flen[i];
            for (int j = i + 1; j < nfree; j++) { foff[j - 1] = foff[j]; flen[j - 1] = flen[j]; }
            nfree--;
        }
    }
};

template <class CKfn>
WHALE_HD inline int place_rows(int nn, int nleafnodes, const int* leafnodes, int ninner, const int* inner,
                               const int* child0, const int* child1, const int* kind, CKfn ck, int* roff,
                               bool keep_leaf = false) {
    RowAlloc al;
    for (int i = 0; i < nleafnodes; i++) roff[leafnodes[i]] = al.take(ck(leafnodes[i]));
    for (int i = 0; i < ninner; i++) {
        const int e = inner[i];
        roff[e] = al.take(ck(e));
        if (kind[e] == WHALE_ROOT) continue;  // the root reads its children level by level: they stay
        // the children are read for the last time while e's row 1 is formed; e's own slices then run in e's
        // row and the scratch row, so later nodes may reuse the children's space
        // (reverse mode keeps the leaf branches' rows: their tangents are contracted in the backward pass)
        if (child0[e] >= 0 && !(keep_leaf && kind[child0[e]] == WHALE_LEAF)) al.give(roff[child0[e]], ck(child0[e]));
        if (child1[e] >= 0 && !(keep_leaf && kind[child1[e]] == WHALE_LEAF)) al.give(roff[child1[e]], ck(child1[e]));
    }
    return al.top;
}

// Adjoint last rows ℓ̄_{e,n} of the internal/WGD/root nodes in the backward pass (nodes in reverse order): a node's
// row is written while its parent's row 1 is transposed and dead once its own row 1 has been transposed.
template <class Cfn>
WHALE_HD inline int place_arows(int ninner, const int* inner, const int* child0, const int* child1, const int* kind,
                                Cfn cf, int* aoff) {
    RowAlloc al;
    if (ninner > 0) aoff[inner[ninner - 1]] = al.take(cf(inner[ninner - 1]));  // the root
    for (int i = ninner - 1; i >= 0; i--) {
        const int e = inner[i];
        for (int j = 0; j < 2; j++) {
            const int c = j == 0 ? child0[e] : child1[e];
            if (c >= 0 && kind[c] != WHALE_LEAF) aoff[c] = al.take(cf(c));
        }
        al.give(aoff[e], cf(e));
    }
    return al.top;
}
