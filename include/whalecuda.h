/*
 * libwhalecuda — C ABI of the B200-native ALE/DLWGD likelihood engine.
 *
 * This is the drop-in boundary for ONE hot path of arzwa/Whale.jl (reference paths are relative to the
 * reference checkout): the per-family amalgamated-likelihood recursion `logpdf`/`logpdf!` -> `whale!`
 * (src/core.jl:29-199), the slice tables it consumes (src/model.jl:162-191, src/bdputil.jl:6-11), the
 * conditioning term (src/condition.jl:11-29) and the stochastic backtracker (src/track.jl:190-414).
 * The reference is pure Julia and has no FFI; the seam is cut at Julia method dispatch (see
 * INTEGRATION.md for the `ccall` glue a maintainer would add).  Conventions:
 *
 *   - plain C types only; every entry point returns a status (0 = WHALE_OK) and never throws/aborts;
 *     the message of the last failure on the calling thread is available via whale_last_error();
 *   - all indices crossing the ABI are 0-based (the Julia glue subtracts 1 from ids);
 *   - the caller owns every buffer it passes (valid for the duration of the call); the library copies
 *     what it needs at *_create time and owns device memory behind the opaque handles;
 *   - a handle is used by one host thread at a time; calls synchronise before returning unless the
 *     name ends in `_async`;
 *   - numerical non-events follow the reference: L <= 0 or a non-finite total gives -Inf
 *     (src/core.jl:15,36), never an error.
 *
 * There is no CPU fallback: every compute entry point fails with WHALE_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef WHALECUDA_H
#define WHALECUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WHALE_OK 0
#define WHALE_ERR_ARG 1      /* invalid argument / inconsistent description            */
#define WHALE_ERR_CUDA 2     /* CUDA runtime failure (message has the CUDA error text) */
#define WHALE_ERR_CAPACITY 3 /* a family does not fit the on-chip working set          */
#define WHALE_ERR_STATE 4    /* call sequence error (e.g. backtrack without keep_ell)  */

/* node kinds (src/model.jl:54, NewickTree isleaf/isroot) */
#define WHALE_LEAF 0
#define WHALE_INTERNAL 1
#define WHALE_WGD 2
#define WHALE_ROOT 3

/* conditioning (src/condition.jl:11-29) */
#define WHALE_COND_NONE 0       /* NoCondition          */
#define WHALE_COND_ROOT 1       /* RootCondition        */
#define WHALE_COND_NONEXTINCT 2 /* NonExtinctCondition  */
#define WHALE_COND_NOWHERE 3    /* NowhereExtinctCondition (src/condition.jl:5-9,31-36; 2^L terms, L <= 20 leaves) */

/* flags for whale_logpdf_grad */
#define WHALE_WANT_GRAD 1u  /* also return d loglik / d raw parameter                          */
#define WHALE_KEEP_ELL 2u   /* logpdf! semantics: keep the full ℓ on the device (src/core.jl:29) */
#define WHALE_PROFILE 4u    /* record CUDA events around each kernel (see whale_last_kernel_ms)   */
#define WHALE_PEER_SUM 8u   /* one process per GPU: return the sum over all ranks (see whale_peer_export) */

typedef struct whale_model* whale_model_t;
typedef struct whale_data* whale_data_t;

/*
 * Species-tree model structure + parameter layout.
 * Replaces: WhaleModel construction (src/model.jl:96-145: node order, ids, slices per branch),
 * Slices (src/model.jl:16-22) and the rates structs' per-node lookup getθ (src/rmodels.jl:31-33,55-64).
 * Node index = Julia node id - 1.  The raw parameter vector x has n_params entries on the scale the
 * rates struct stores them (log scale for DLWGD, natural for ConstantDLWGD); *_slot give, per node,
 * which entry of x it reads (-1: NaN rates, e.g. a root omitted from a DLWGD vector).  WGD nodes carry
 * the slots of their nonwgdchild (src/rmodels.jl:56-58) and their own q slot (index by wgdid).
 */
typedef struct {
    int32_t n_nodes;
    const int32_t* order;    /* [n_nodes] processing order: leaves, then postorder (src/model.jl:124) */
    const int32_t* child0;   /* [n_nodes] first child or -1                                            */
    const int32_t* child1;   /* [n_nodes] second child or -1 (WGD nodes have one child)                */
    const int32_t* kind;     /* [n_nodes] WHALE_LEAF / INTERNAL / WGD / ROOT                           */
    const int32_t* n_slices; /* [n_nodes] n (the slices matrix has n+1 rows; root: 0)                  */
    const double* slice_dt;  /* [n_nodes] slice length t/n (0 for the root)                            */
    const double* leafP;     /* [n_nodes] Slices.leafℙ (src/model.jl:110-111)                          */
    int32_t n_params;        /* P = length of x                                                        */
    const int32_t* lam_slot; /* [n_nodes]                                                              */
    const int32_t* mu_slot;  /* [n_nodes]                                                              */
    const int32_t* q_slot;   /* [n_nodes] -1 unless WGD                                                */
    int32_t eta_slot;
    int32_t log_scale;       /* 1: λ = exp(x[lam_slot]) (DLWGD); 0: λ = x[lam_slot] (ConstantDLWGD)    */
} whale_model_desc;

/*
 * A batch of CCDs in the reference's own layout, flattened (src/ccd.jl:80-89,102-121):
 * clades of family f are clade_off[f] .. clade_off[f+1]-1, sorted by size (leaves first, the
 * ubiquitous clade last); clade c's triples are split_off[c] .. split_off[c+1]-1 in file order with
 * family-local clade ids g1/g2 and probability p; compat lists (ascending family-local clade ids) of
 * family f at node e are compat_off[f*n_nodes+e] .. compat_off[f*n_nodes+e+1]-1.
 * Clade ids must fit UInt16 like the reference's Triple{UInt16} (src/ccd.jl:13-20).
 */
typedef struct {
    int32_t n_fam;
    const int64_t* clade_off;   /* [n_fam+1]           */
    const int32_t* clade_nleaf; /* [total clades]      */
    const int64_t* split_off;   /* [total clades + 1]  */
    const int32_t* g1;          /* [total triples]     */
    const int32_t* g2;
    const double* p;
    const int64_t* compat_off;  /* [n_fam*n_nodes + 1] */
    const int32_t* compat;
} whale_ccd_desc;

int32_t whale_version(void);
/* copies the calling thread's last error message (NUL-terminated, truncated to n) */
int32_t whale_last_error(char* buf, size_t n);
/* number of usable CUDA devices (0 when none); selects the device used by subsequent *_create calls */
int32_t whale_device_count(void);
int32_t whale_set_device(int32_t device);

/*
 * Several GPUs behind one handle (whale_set_devices(n, ids[])).  Families are i.i.d. terms of a sum
 * (src/core.jl:54,63; the reference's parallel axis is Threads.@threads over families, :58-64): whale_multi_create packs
 * the model on every listed device and shards the families over them by predicted work (longest processing time first);
 * whale_multi_logpdf_grad evaluates all shards concurrently and adds the per-device (loglik, grad) in device-list order,
 * so the result is bit-reproducible.  ll_fam (nullable, [n_fam]) is indexed by the ORIGINAL family order.  A device may be
 * listed more than once (two shards on one GPU).  Without whale_set_devices the current device (whale_set_device) is used.
 */
typedef struct whale_multi* whale_multi_t;
int32_t whale_set_devices(int32_t n, const int32_t* ids);
int32_t whale_multi_create(const whale_model_desc* model, const whale_ccd_desc* data, whale_multi_t* out);
int32_t whale_multi_destroy(whale_multi_t h);
int32_t whale_multi_ndev(whale_multi_t h);
int32_t whale_multi_shard_size(whale_multi_t h, int32_t i);
int32_t whale_multi_logpdf_grad(whale_multi_t h, const double* x, const double* p_leaf, int32_t condition, uint32_t flags,
                                double* loglik, double* grad, double* ll_fam);

/*
 * One process per GPU (torch.distributed / MPI style drivers; families sharded over the ranks like Threads.@threads shards
 * them over threads, src/core.jl:58-64): the sum of the ranks' (loglik, grad) WITHOUT a collective library.  Each rank
 * calls whale_peer_export (allocates its exchange buffer, returns a 64-byte CUDA IPC handle), the driver all-gathers the
 * handles by any means, every rank calls whale_peer_import for every other rank.  From then on an evaluation with
 * WHALE_PEER_SUM ends with the exchange: the rank stores its 1+P doubles into every peer's buffer through NVLink peer
 * memory (64-bit packets carrying data and the step's tag together: no fence, one hop), waits for the world's packets in
 * its own buffer and adds the contributions in rank order: every rank gets the same bits.  One-pass evaluations run it in
 * the last CTA of the DP kernel (no extra launch), others as a one-CTA kernel.  Like a collective, all ranks must issue the
 * same sequence of WHALE_PEER_SUM evaluations.  At most 16 ranks.  A peer that does not show up within WHALE_PEER_TIMEOUT_S
 * (environment, default 60) seconds makes the result -Inf (and zero gradient) instead of hanging: synchronise the ranks
 * (a barrier of the driver's own) before the first such evaluation if their set-up times differ by more than that.
 */
int32_t whale_peer_export(whale_data_t d, int32_t rank, int32_t world, void* handle64);
int32_t whale_peer_import(whale_data_t d, int32_t peer, const void* handle64);
int32_t whale_peer_ready(whale_data_t d);
/* the exchange alone: d_out ([1 + n_params] doubles on the device, e.g. a prior's value and gradient the driver wants
 * summed the same way) is replaced by the sum over ranks; ordered on `stream`; counts as one WHALE_PEER_SUM step */
int32_t whale_peer_sum_async(whale_data_t d, double* d_out, void* stream);

int32_t whale_model_create(const whale_model_desc* desc, whale_model_t* out);
int32_t whale_model_destroy(whale_model_t m);

/* read_ale-time packing (src/ccd.jl:126-137 + CCD ctor): builds the device arena once.  Ends with a calibration
 * pass (two profiled evaluations at a benign parameter point on the model's stream) whose measured per-family SM
 * cycles order the launch; results never depend on that order.  WHALE_CALIBRATE=0 skips it. */
int32_t whale_data_create(whale_model_t m, const whale_ccd_desc* desc, whale_data_t* out);
/*
 * read_ale in one native call (src/ccd.jl:126-248): parse n_files ALEobserve `.ale` files on n_threads host
 * threads (0 = all), build each CCD in the reference's layout (leaf clades, the ubiquitous clade, ids by size,
 * split probabilities, compat lists) and pack the device arena.  The species tree contributes the gene-name
 * prefix -> species id map (species_names/species_ids, `n_species` entries; MUL trees give several leaves one id,
 * src/model.jl:133-135) and, per node, the species ids below it (node_clade_off[n_nodes+1], node_clade).
 * n_clades (nullable, [n_files]) receives Γ of every family.
 */
int32_t whale_read_ale(whale_model_t m, int32_t n_files, const char* const* paths, int32_t n_species,
                       const char* const* species_names, const int32_t* species_ids, const int64_t* node_clade_off,
                       const int32_t* node_clade, int32_t n_threads, whale_data_t* out, int32_t* n_clades);
/*
 * Binary arena cache (the "binary arena cache format" of the ingest row): whale_data_save writes the packed state
 * of a handle (family headers, arena, packer facts); whale_data_load rebuilds a handle from it for a model with the
 * same species tree and slicing (checked by a structure fingerprint) without parsing .ale files or re-packing —
 * tangent plans, shared-memory budgets, launch order and calibration are redone for the loading process.
 * Replaces re-running read_ale (src/ccd.jl:126-137) on an unchanged data set.
 */
int32_t whale_data_save(whale_data_t d, const char* path);
int32_t whale_data_load(whale_model_t m, const char* path, whale_data_t* out);
int32_t whale_data_destroy(whale_data_t d);
int32_t whale_data_nfam(whale_data_t d);
/* how this handle computes gradients: 1 = reverse-mode (adjoint) DP, one pass whatever P is (the default);
 * 0 = forward tangents (WHALE_GRAD_MODE=fwd, or a family that does not fit the reverse kernel's working set), in
 * whale_data_grad_passes() passes over parameter chunks */
int32_t whale_data_grad_mode(whale_data_t d);
int32_t whale_data_grad_passes(whale_data_t d);
/* bytes of the packed arena resident in HBM, and the algorithmic bytes one evaluation reads */
int64_t whale_data_arena_bytes(whale_data_t d);
/* copies the packed arena back (for the bit-exact packing tests); buf may be NULL to query the size */
int64_t whale_data_arena_dump(whale_data_t d, void* buf, int64_t cap);

/*
 * logpdf / logpdf! over a vector of CCDs, fused with the forward-mode gradient
 * (replaces src/core.jl:46-64 + ForwardDiff.gradient over it, test/runtests.jl:36-38):
 *   *loglik  = Σ_f log L_f − n_fam·condition(model)          (ℓhood-guarded, src/core.jl:15)
 *   grad[P]  = ∂ *loglik / ∂ x                                (NULL unless WHALE_WANT_GRAD)
 *   ll_fam[F]= unconditioned per-family log L_f (nullable)    (for mixtures, src/core.jl:66-79)
 *   grad_fam[F*P] (nullable) per-family ∂ log L_f / ∂ x
 * p_leaf (nullable, [n_nodes]) = sampling-failure probabilities getp (src/rmodels.jl:14).
 */
int32_t whale_logpdf_grad(whale_model_t m, whale_data_t d, const double* x, const double* p_leaf,
                          int32_t condition, uint32_t flags, double* loglik, double* grad, double* ll_fam,
                          double* grad_fam);

/*
 * Mixture of n_comp WhaleModels that share the species-tree structure and differ in their raw parameters
 * (logpdf(mm::MixtureModel{…,<:WhaleModel}, xs), src/core.jl:66-76):
 *     ℓ = Σ_i logsumexp_j ( log L_i(x_j) + log_w[j] − condition(x_j) )
 * evaluated on the device: every component runs the fused tables -> DP pass, the F × n_comp matrix, its row-wise
 * logsumexp and the responsibilities stay in HBM.  x is [n_comp][P] (row-major), log_w [n_comp].
 * With WHALE_WANT_GRAD: grad_x [n_comp][P] = ∂ℓ/∂x_j and grad_logw [n_comp] = ∂ℓ/∂log_w[j] (= Σ_i responsibility_ij).
 */
int32_t whale_mixture_logpdf_grad(whale_model_t m, whale_data_t d, int32_t n_comp, const double* x, const double* log_w,
                                  const double* p_leaf, int32_t condition, uint32_t flags, double* loglik,
                                  double* grad_x, double* grad_logw);

/*
 * Device-resident variant: d_x (P doubles) and d_out (1+P doubles: loglik, grad) live in device memory;
 * work is enqueued on `stream` (a cudaStream_t) and NOT synchronised.  Used by the benchmark's
 * device-timed leg and by multi-GPU drivers that all-reduce d_out with NCCL on the same stream.
 */
int32_t whale_logpdf_grad_async(whale_model_t m, whale_data_t d, const double* d_x, int32_t condition,
                                uint32_t flags, double* d_out, void* stream);

/* slice tables for table-parity tests: rows of all nodes concatenated in node-index order,
 * (n_slices[e]+1) rows each (src/model.jl:182-191) */
int32_t whale_slices(whale_model_t m, const double* x, const double* p_leaf, double* eps, double* phi,
                     double* psi);

/* the full ℓ of family `fam` kept by the last WHALE_KEEP_ELL evaluation: concatenation over nodes
 * (index order) of row-major (n_slices[e]+1) x C_e matrices (ccd.ℓ, src/ccd.jl:59-75) */
int64_t whale_ell_size(whale_data_t d, int32_t fam);
int32_t whale_ell_get(whale_data_t d, int32_t fam, double* out);

/*
 * backtrack(wm, ccd) (src/track.jl:190-414) for n_samples reconciliations per family, on the ℓ kept by
 * the last WHALE_KEEP_ELL evaluation, with host-supplied uniforms instead of rand():
 * sample s of family f consumes uniforms[(f*n_samples+s)*stride ...] in the order the reference calls
 * rand().  Output per (f,s): node_count, and at most max_nodes nodes (gamma = clade id or -1 for loss
 * nodes, e = node index, t = 1-based row index (0 for loss nodes), parent = index of the parent node in
 * this tree or -1) in creation (DFS) order; status 0 ok, 1 "Backtracking failed" (src/track.jl:150),
 * 2 node overflow, 3 uniforms exhausted.
 */
int32_t whale_backtrack(whale_model_t m, whale_data_t d, int32_t n_samples, const double* uniforms,
                        int64_t stride, int32_t max_nodes, int32_t* node_count, int32_t* gamma, int32_t* e,
                        int32_t* t, int32_t* parent, int32_t* status);

/*
 * The same walks with the results kept on the device (persistent buffers, no padded transfers): n_samples walks per
 * family from the kept ℓ.  uniforms may be NULL: draw k of walk w then comes from a counter-based generator keyed by
 * (seed, w, k) on the device (the reference calls the global rand(), src/track.jl:217,274 — with host-supplied uniforms
 * the stream is explicit and reproducible against the oracle instead).  total_nodes (nullable) = Σ node counts.
 * Fetch with whale_trees_counts / whale_trees_get / whale_trees_view, summarise with whale_trees_summary.
 */
int32_t whale_backtrack_device(whale_model_t m, whale_data_t d, int32_t n_samples, const double* uniforms,
                               int64_t stride, uint64_t seed, int32_t max_nodes, int64_t* total_nodes);
/*
 * track_and_sum's sampling loop in its general form (src/track.jl:47-63): sample s of family f uses posterior row
 * theta_index[f*n_samples + s] of x[n_theta][P] (the reference draws `i = rand(1:length(df))` per family and sample;
 * NULL: row s mod n_theta).  For every row that is used, logpdf! runs for exactly the families that drew it (value-only
 * DP keeping ℓ, launch order = that family list) and their walks follow on the same stream; nothing returns to the host
 * in between.  Results stay on the device like whale_backtrack_device's.
 */
int32_t whale_track_sample(whale_model_t m, whale_data_t d, int32_t n_theta, const double* x, const double* p_leaf,
                           int32_t n_samples, const int32_t* theta_index, const double* uniforms, int64_t stride,
                           uint64_t seed, int32_t max_nodes, int64_t* total_nodes);
/* node counts and statuses ([F*n_samples], family-major) of the trees on the handle */
int32_t whale_trees_counts(whale_data_t d, int32_t* node_count, int32_t* status);
/* the trees in compact form: offsets[F*n_samples + 1] (rows) and total_nodes rows of (gamma, e, t, parent) — packed on
 * the device (one warp per tree), one pinned D2H copy.  _view returns pointers into library-owned pinned memory (valid
 * until the next backtracking call on the handle), _get copies into caller buffers. */
int32_t whale_trees_view(whale_data_t d, const int64_t** offsets, const int32_t** nodes4, int64_t* total_nodes);
int32_t whale_trees_get(whale_data_t d, int64_t* offsets, int32_t* nodes4);
/*
 * sumtrees on the device (src/rectree.jl:113-133 with the tree identity of src/track.jl:95-113: the set over nodes of
 * (γ, e, {(γ, e) of the children}), loss nodes (loss, e, γ of the sister); the slice t is not part of it): a 64-bit
 * identity hash per tree, then per family the distinct trees — n_distinct[F]; for family f the first n_distinct[f]
 * entries of hash/count/first[f*n_samples ...] hold the identity, how many samples showed it and the first such sample
 * (ascending hash; sort by count for the reference's order).  tree_hash (nullable, [F*n_samples]) = every tree's hash.
 */
int32_t whale_trees_summary(whale_data_t d, int32_t* n_distinct, uint64_t* hash, int32_t* count, int32_t* first,
                            uint64_t* tree_hash);

/* counters for benchmarks: kernels launched by this library since load, and the last evaluation's
 * algorithmic flop / byte counts (SURVEY §8d coefficients) */
/*
 * Fused track_and_sum inner loop (src/track.jl:47-63): for each of the n_theta parameter vectors (posterior draws,
 * x is [n_theta][P]) run logpdf! over all families keeping ℓ and backtrack ONE reconciled tree per family from it,
 * everything enqueued back to back on the device (no host round trip between the draws).  uniforms is
 * [F][n_theta][stride]; outputs are laid out like whale_backtrack's with n_samples = n_theta.  loglik (nullable,
 * [n_theta]) receives Σ_f log L_f − F·condition for every draw.  The reference draws a posterior row per (family,
 * sample); here a draw is shared by all families of the batch (same marginal distribution per family).
 */
int32_t whale_track(whale_model_t m, whale_data_t d, int32_t n_theta, const double* x, const double* p_leaf,
                    int32_t condition, const double* uniforms, int64_t stride, int32_t max_nodes, int32_t* node_count,
                    int32_t* gamma, int32_t* e, int32_t* t, int32_t* parent, int32_t* status, double* loglik);

int64_t whale_launch_count(void);
int32_t whale_work_estimate(whale_model_t m, whale_data_t d, uint32_t flags, double* flops, double* bytes);

/* device time of the kernels of the last WHALE_PROFILE evaluation on this data handle, measured with CUDA
 * events on the stream they were launched on (waits for that evaluation to finish) */
int32_t whale_last_kernel_ms(whale_data_t d, double* tables_ms, double* dp_ms, double* reduce_ms);
/* (the sum over families runs in the tail of k_dp: reduce_ms is then the gap between the end of k_dp and the end of
 * the evaluation; with WHALE_FUSED_REDUCE=0 it times the separate reduction kernels) */

/* device time of the k_backtrack launch of the last whale_backtrack call on this handle (CUDA events) */
int32_t whale_last_backtrack_ms(whale_data_t d, double* ms);

/* SM-cycle breakdown of the DP kernel over the families of the last WHALE_PROFILE evaluation (mean and max
 * over families): forward tangents [prologue, leaf phase, staging, row 1, slices, root, total, 0]; reverse mode
 * [prologue, leaf phase, forward staging + row 1, transposed row 1, slices forward + transposed, root forward +
 * transposed, total, contraction] */
int32_t whale_last_phase_cycles(whale_data_t d, double* mean8, double* max8);

/* same evaluation, per inner node in processing order (internal / WGD nodes, the root last): mean SM cycles of
 * the slice loop and of staging + row 1; returns the number of entries written (<= cap) or a negative status */
/* same evaluation, the eight phase counters of every family: out8[f*8 + j] */
int32_t whale_last_family_cycles(whale_data_t d, double* out8);
int32_t whale_last_node_cycles(whale_data_t d, double* slices_mean, double* row1_mean, int32_t cap);

/* SM-cycle stamps of the last k_tables launch of the model's value / full-gradient plan: [0] metadata staged,
 * [1+L] level L done, [30] all rows written, [31] number of levels */
int32_t whale_last_tables_cycles(whale_model_t m, int32_t with_grad, double* out32);

/* measured fp64 FMA peak of the current device (dependent-free DFMA microbenchmark), TFLOP/s */
int32_t whale_fp64_peak(double* tflops);

#ifdef __cplusplus
}
#endif
#endif
