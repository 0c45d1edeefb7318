/*
 * deephall_b200 -- C ABI of the B200-native walker-evaluation engine for DeepHall.
 *
 * The reference (peterzjx/DeepHall) is pure Python on JAX and has no FFI of its own; its
 * seams are Python factories returning closures.  Each entry point below is what a
 * `jax.ffi` / ctypes binding for that seam would call (INTEGRATION.md shows the stubs):
 *
 *   dh_logpsi          <- model.apply(params, x)            networks/psiformer.py:72-76
 *                         (vmapped: train.py:69)             types.py:68-70
 *   dh_local_energy    <- local_energy(f, system)._e_l      hamiltonian.py:175-212
 *                         (vmapped: loss.py:51,67)           types.py:49-65
 *   dh_potential       <- make_potential(...).potential     hamiltonian.py:63-80
 *   dh_mcmc_sweep      <- make_mcmc_step(...).mcmc_step     mcmc.py:105-150
 *   dh_mcmc_propose    <- sph_sampling                      mcmc.py:67-102
 *   dh_mcmc_accept     <- mh_update (accept/select part)    mcmc.py:56-62
 *   dh_logpsi_vjp      <- df_real/df_imag + loss_prod       loss.py:53-64,96-106
 *   dh_kfac_factors    <- kfac_jax curvature estimation     optimizers/kfac.py:42-102, loss.py:98
 *   dh_slogdet         <- jnp.linalg.slogdet + tail         psiformer.py:74-76
 *   dh_param_layout    <- the flax parameter tree           psiformer.py, blocks.py
 *   dh_init_walkers    <- init_guess                        train.py:40-54
 *   dh_pair_correlation, dh_density_histogram, dh_overlap_sum/_ratio, dh_lll_orbitals, dh_one_rdm_scatter/_product
 *                      <- the NetObs estimators            netobs_bridge/observables/*.py (see the end of this file)
 *   network_type = 1   <- Laughlin(nspins, flux).apply      networks/laughlin.py:19-100 (ground state, quasihole, quasiparticle;
 *                         no parameters: dh_param_count = 0, dh_logpsi_vjp writes nothing)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - all floating-point data is fp32; complex data is interleaved (re, im) fp32;
 *   - walkers `x` are (B, N, 2) = (theta, phi), exactly the reference layout (mcmc.py:68);
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), performs
 *     no allocation and no host synchronisation; the caller owns every buffer including
 *     the workspace (size from dh_workspace_bytes).  A pass over more walkers than one internal chunk
 *     alternates its chunks between `stream` and a stream the plan owns (created by dh_plan_create,
 *     forked from and joined back into `stream` by events): to the caller the call is still ordered on `stream`;
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = DH_E_* argument error.  Numerical
 *     pathologies follow the reference: NaN/-inf propagate into the outputs
 *     (singular determinant -> log|psi| = -inf, NaN proposal -> rejected, mcmc.py:59).
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef DEEPHALL_B200_H
#define DEEPHALL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DH_E_BADARG (-1)
#define DH_E_UNSUPPORTED (-2)
#define DH_E_WORKSPACE (-3)

/* System + network hyper-parameters: config.py:56-104 (System, PsiformerNetwork). */
typedef struct dh_config {
  int32_t n_up, n_dn;       /* system.nspins; n_up >= 1, n_dn >= 0 (Laughlin: n_dn = 0)  */
  int32_t flux;             /* system.flux = 2Q                                        */
  int32_t ndets;            /* network.psiformer.determinants                         */
  int32_t num_heads;        /* network.psiformer.num_heads                            */
  int32_t heads_dim;        /* network.psiformer.heads_dim                            */
  int32_t num_layers;       /* network.psiformer.num_layers                           */
  int32_t interaction_type; /* 0 = coulomb, 1 = harmonic (config.py:51-53)            */
  float interaction_strength; /* system.interaction_strength                          */
  float radius;             /* system.radius; <= 0 means sqrt(Q) (hamiltonian.py:189) */
  int32_t chunk_walkers;    /* walkers per internal pass (0 = library default)        */
  int32_t network_type;     /* network.type: 0 = psiformer, 1 = laughlin (config.py:82-84)          */
  int32_t cf_flux;          /* laughlin: composite-fermion flux p (networks/laughlin.py:25), 0 -> 1  */
  int32_t orbital_type;     /* network.orbital: 0 = full, 1 = sparse (config.py:87-89, blocks.py:47-62) */
  int32_t contraction;      /* arithmetic of the dense / attention contractions: 0 = fp16 hi/lo pieces on the tensor cores (default),
                               1 = TF32 hi/lo pieces (fp32 exponent range; the fallback when dh_plan_status reports saturation),
                               2 = plain fp32 FMA (cross-check) */
  float excitation_lz;      /* laughlin quasihole (N = 2 Q1) / quasiparticle (N = 2 Q1 + 2): L_z of the excitation = system.lz_center (networks/__init__.py:25-27) */
} dh_config;

typedef struct dh_plan dh_plan;

/* One entry of the flat parameter layout (mirrors the flax tree path). */
typedef struct dh_param_entry {
  char name[96];    /* e.g. "PsiformerLayers_0/MultiHeadAttention_0/query/kernel" */
  int64_t offset;   /* in floats, into the flat parameter vector                  */
  int32_t ndim;
  int32_t shape[4];
} dh_param_entry;

enum dh_op { DH_OP_LOGPSI = 0, DH_OP_LOCAL_ENERGY = 1, DH_OP_MCMC = 2, DH_OP_VJP = 3, DH_OP_KFAC = 4 };

/* One curvature block of the KFAC optimizer (optimizers/kfac.py:33-241; the reference's default optimizer).
 * kind 0: a dense layer seen as a `repeated_dense` block (kfac.py:42-102: the electron axis is extra batch) with
 *         Kronecker factors A = mean_rows[x~ x~^T] (x~ = layer input, plus a 1 when the layer has a bias) and
 *         G = mean_rows[g g^T] (g = d Re log psi_b / d layer output).  dh_kfac_factors delivers the SUMS over rows
 *         sum x x^T [in][in] at xtx_offset, sum x [in] at xsum_offset (has_bias only) and sum g g^T [out][out] at
 *         gtg_offset; rows = B * rows_per_walker.  xtx_offset = -1: the caller forms it itself (Dense_0, whose
 *         inputs are the four features of psiformer.py:51-60).  Layers that share their input share xtx / xsum.
 * kind 1: LayerNorm scale / bias (kfac_jax's scale-and-shift blocks): diagonal Fisher from per-walker gradients,
 *         sum_b (d Re log psi_b / d theta)^2 at diag_offset, `size` entries.
 * kind 2: a parameter no layer pattern matches (Jastrow ee_par / ee_anti, and lll_weight kernel / bias of the sparse
 *         orbitals, blocks.py:57: a DenseGeneral over axis 1; kfac_jax's generic tag with its "naive" diagonal = square
 *         of the batch-summed gradient): sum_b d Re log psi_b / d theta at diag_offset, `size` entries.
 * Sparse orbitals: the kind-0 blocks of the orbital projections have out_dim = 8 N K (their own 8-feature output). */
typedef struct dh_kfac_entry {
  char name[96];             /* parameter path of the kernel (kind 0) or of the parameter (kind 1)      */
  int32_t kind, in_dim, out_dim, has_bias, rows_per_walker;
  int64_t kernel_offset, bias_offset;         /* into the flat parameter vector; bias_offset = -1: none */
  int64_t xtx_offset, xsum_offset, gtg_offset, diag_offset; /* into the factor vector; -1: not present  */
  int64_t size;              /* kind 1: number of parameters                                            */
} dh_kfac_entry;

int dh_plan_create(const dh_config* cfg, dh_plan** out);
int dh_plan_destroy(dh_plan* plan);
const char* dh_version(void);

/* Flat parameter vector: number of floats, and the name/offset/shape table. */
int64_t dh_param_count(const dh_plan* plan);
int dh_param_layout(const dh_plan* plan, dh_param_entry* entries_host, int32_t* n_entries_host);

/* Prepared weights.  The tensor-core contractions use transposed, two-piece-split copies of the dense kernels
 * (plus two folded products), kept in a plan-owned buffer.  By default (auto_prepare = 1) every op rebuilds them
 * from `params` (about 40 tiny launches), because the library cannot see in-place updates.  A caller that
 * updates the parameters once per optimisation step (train.py:140) can switch that off and call
 * dh_params_prepare after each update; ops then reuse the copies as long as they are given the same pointer. */
int dh_params_prepare(dh_plan* plan, const float* params, void* stream);
int dh_plan_set_auto_prepare(dh_plan* plan, int32_t on);

/* Bytes of device workspace an op needs for a batch of B walkers. */
int dh_workspace_bytes(const dh_plan* plan, int op, int64_t B, size_t* bytes_host);

/* log psi of B walkers.  out_logpsi: (B) complex64 = (log|psi|, phase). */
int dh_logpsi(dh_plan* plan, const float* params, const float* x, int64_t B, float* out_logpsi,
              void* ws, size_t ws_bytes, void* stream);

/* Local energy + observables (hamiltonian.py:207-210).  Any output may be NULL.
 *   out_el, out_kinetic: (B) complex64;  out_potential, out_lz, out_lz2, out_l2: (B) f32;
 *   out_logpsi: (B) complex64 (free by-product). */
int dh_local_energy(dh_plan* plan, const float* params, const float* x, int64_t B, float* out_el,
                    float* out_kinetic, float* out_potential, float* out_lz, float* out_lz2,
                    float* out_l2, float* out_logpsi, void* ws, size_t ws_bytes, void* stream);

/* interaction_strength is NOT applied here (hamiltonian.py:78 vs :205). */
int dh_potential(dh_plan* plan, const float* x, int64_t B, float* out, void* stream);

/* `steps` Metropolis-Hastings all-electron moves (mcmc.py:122-148).
 *   x_inout (B,N,2) is updated in place (the reference donates it, train.py:75);
 *   randoms: NULL -> in-kernel Philox4x32-10 keyed by (seed, offset); walker b uses
 *            subsequence `subsequence0 + b`.  Otherwise "injected randoms":
 *            (steps, B, 2N+1) f32 = [N normals | N uniforms (phi') | 1 uniform (accept)].
 *   out_naccept: one int64 on the device, overwritten with the number of accepted moves
 *            (pmove = naccept / (steps * B), mcmc.py:146).
 *   out_lp: optional (B) f32, 2 Re log psi of the final configurations. */
int dh_mcmc_sweep(dh_plan* plan, const float* params, float* x_inout, int64_t B, int32_t steps,
                  float width, uint64_t seed, uint64_t offset, uint64_t subsequence0,
                  const float* randoms, long long* out_naccept, float* out_lp, void* ws,
                  size_t ws_bytes, void* stream);

/* The same sweep with its traced arguments on the DEVICE: width_dev (1 float) and key_dev (2 x uint64: Philox seed, offset)
 * are read by the kernels, not by the host -- the form a jit-compiled caller needs (the reference's `mcmc_width` lives in the
 * CheckpointState and its key is split every step: both are traced values under `jax.jit` / `pmap`, train.py:126-131), and
 * the form the XLA FFI shim binds (integration/jax_ffi/dh_xla_ffi.cc).  In-kernel Philox only (no injected randoms). */
int dh_mcmc_sweep_dev(dh_plan* plan, const float* params, float* x_inout, int64_t B, int32_t steps, const float* width_dev,
                      const uint64_t* key_dev, uint64_t subsequence0, long long* out_naccept, float* out_lp, void* ws,
                      size_t ws_bytes, void* stream);

/* One proposal (mcmc.py:67-102) and one accept/select (mcmc.py:56-62), exposed for parity
 * tests.  `randoms` as above with steps = 1 (NULL -> Philox). */
int dh_mcmc_propose(dh_plan* plan, const float* x1, int64_t B, float width, uint64_t seed,
                    uint64_t offset, uint64_t subsequence0, const float* randoms, float* x2,
                    void* stream);
int dh_mcmc_accept(dh_plan* plan, float* x1_inout, const float* x2, float* lp1_inout,
                   const float* lp2, int64_t B, uint64_t seed, uint64_t offset,
                   uint64_t subsequence0, const float* randoms, long long* naccept_inout,
                   void* stream);

/* Uniform points on the sphere (train.py:40-54) from Philox. */
int dh_init_walkers(dh_plan* plan, float* x, int64_t B, uint64_t seed, uint64_t subsequence0,
                    void* stream);

/* grad_flat (P floats) = sum_b [ cot[b,0] * d Re logpsi_b / d params
 *                              + cot[b,1] * d Im logpsi_b / d params ]        (loss.py:99-106:
 * the caller passes cot = (2/B) * (Re diff, Im diff)).  grad_flat is overwritten.
 * out_logpsi optional. */
int dh_logpsi_vjp(dh_plan* plan, const float* params, const float* x, int64_t B,
                  const float* cot, float* grad_flat, float* out_logpsi, void* ws,
                  size_t ws_bytes, void* stream);

/* Energy statistics of a walker batch, replaces loss.py:30-38,66-92 (iqr_clip, nanmean, nanquantile on the rank's shard).
 * dh_energy_stats: the rank-local means the reference `pmean`s, packed in out16 (device, 16 floats):
 *   [0],[1] mean kinetic (re, im)   [2] mean potential   [3] mean L_z   [4] mean L_z^2   [5] mean L^2     (plain means, :68-71)
 *   [6],[7] nanmean E_L (:73)       [8],[9] nanmean iqr_clip(E_L) (:74)                  [10] nanmean (Re E_L)^2 (:91)
 *   [11] nanmean iqr_clip(L_z^2)    [12] nanmean iqr_clip(L_z)    [13] nanmean iqr_clip(L^2)              (:79,:80,:87)
 *   The caller all-reduces (means) the vector over ranks -- ONE collective -- and hands it to
 * dh_energy_diff: out_diff (B complex) = iqr_clip(E_L - clipped [+ lz_penalty ((L_z^2 - c) - 2 lz_center (L_z - c)) + l2_penalty
 *   (L^2 - c)]) (:75-89); out_cot (B x 2) = (2 / n_ok) diff for walkers whose diff is a number and whose log psi (optional,
 *   B complex, NULL = not checked) is finite, 0 otherwise: the cotangent of the gradient's VJP (:60-64,99-106);
 *   out_ok (B floats, 1 / 0, may be NULL); out_counts (2 floats: walkers with a numeric diff, n_ok; may be NULL).
 * Quantiles are rank-local (the reference's nanquantile runs on the local shard).  E_L / kinetic / logpsi are complex64
 * (interleaved), the observables f32.  B <= 32768 walkers per rank (one shared-memory sort per array). */
int dh_energy_stats(const float* el, const float* kinetic, const float* potential, const float* lz, const float* lz2,
                    const float* l2, int64_t B, float* out16, void* stream);
int dh_energy_diff(const float* el, const float* lz, const float* lz2, const float* l2, const float* logpsi, int64_t B,
                   const float* reduced16, float lz_penalty, float lz_center, float l2_penalty, float* out_diff,
                   float* out_cot, float* out_ok, float* out_counts, void* stream);

/* Range guard of the fp16-piece contractions.  The dense and attention contractions split every fp32 operand into two
 * fp16 pieces; fp16 tops out at 65504 where the reference's fp32 does not.  A piece that saturates is never silent: the
 * kernel that saturates it ORs bit 0 into the plan's device status word.
 *   dh_plan_status      copies the word to the host (synchronises `stream`), optionally clearing it.
 *   dh_plan_status_copy copies it device-to-device, stream-ordered, into `dst_device` (no synchronisation): lets a
 *                       caller fold the flag into statistics it reads back anyway.
 * A plan created with dh_config.contraction = 1 runs every contraction on TF32 pieces (fp32 exponent range) instead;
 * callers re-run an op that reported saturation on such a plan (deephall_b200/train.py does so automatically). */
int dh_plan_status(dh_plan* plan, int32_t clear, uint32_t* out_bits, void* stream);
int dh_plan_status_copy(dh_plan* plan, uint32_t* dst_device, void* stream);

/* Curvature statistics for KFAC with the reference's registration (loss.py:98: a unit-variance normal predictive
 * distribution on Re log psi, `fisher_exact`): one forward + one reverse pass with cotangent (1, 0) per walker.
 * dh_kfac_layout: entries == NULL -> *n = number of blocks; *factor_floats = length of the factor vector.
 * dh_kfac_factors overwrites `factors` (factor_floats floats) with the sums described at dh_kfac_entry (full and sparse
 * orbitals; DH_E_UNSUPPORTED for the parameter-free Laughlin network and for sparse orbitals with 8 N K > 3 D).
 * Workspace: DH_OP_KFAC.  Not available for the Laughlin network (no parameters) and sparse orbitals. */
int dh_kfac_layout(const dh_plan* plan, dh_kfac_entry* entries, int32_t* n, int64_t* factor_floats);
int dh_kfac_factors(dh_plan* plan, const float* params, const float* x, int64_t B, float* factors, void* ws,
                    size_t ws_bytes, void* stream);
/* The same, allowed to skip its forward pass: when the plan's previous op was dh_logpsi_vjp (or dh_kfac_factors) on the
 * same params, x, B and ws, the batch fitted one chunk, and the caller has not written to params, x or ws since, that
 * op's activations are still in the workspace (the KFAC step calls it right after the gradient's VJP, optimizers/kfac.py
 * estimates the curvature on the batch of the gradient).  In every other case it is dh_kfac_factors. */
int dh_kfac_factors_reuse_forward(dh_plan* plan, const float* params, const float* x, int64_t B, float* factors, void* ws,
                                  size_t ws_bytes, void* stream);

/* The KFAC update from the moving-average statistics (optimizers/kfac.py:202-219 hands the loss to kfac_jax.Optimizer;
 * its update rule is restated in oracle/kfac.py).  `stats` is the moving average of the factor vector of dh_kfac_factors
 * (each entry already divided by its row count), `dense0_xtx` the same for Dense_0's 4 x 4 input factor, `weight` the
 * moving-average weight both are divided by.
 *   dh_kfac_update_shape    -> number and padded size of the damped Kronecker factors in the "small" (<= 288 rows) and the
 *                              "large" batch, the number of dense blocks and the floats of one gather buffer.  First call
 *                              builds the plan's descriptor tables (allocates; later calls and the two ops below do not).
 *                              DH_E_UNSUPPORTED: a factor has more than 1024 rows, Laughlin.
 *   dh_kfac_damped_factors  -> coef [8 n_blocks] = {tr_avg(A), tr_avg(G), d, c_k, ok, -, -, -} per block and every
 *                              A / tr_avg(A) + d I, G / tr_avg(G) + d I (kfac_jax's pi-adjusted damping with average-trace
 *                              norms, d = sqrt(damping / rows_per_walker / (tr_avg(A) tr_avg(G)))), each embedded as
 *                              diag(M, I) in its batch: mats_small [n_small][dim_small^2], mats_large likewise.
 *   (the caller inverts the batches with dh_spd_inverse -- sharded over ranks + all-gathered when there are several)
 *   dh_kfac_update          -> out [num_params] = A~^-1 V G~^-1 / (c_k^2 rows_per_walker) for the dense blocks,
 *                              g / (F + damping) for the diagonal ones.  ws: 3 * gather_floats floats, 16-byte aligned. */
int dh_kfac_update_shape(dh_plan* plan, int32_t* n_small, int32_t* dim_small, int32_t* n_large, int32_t* dim_large,
                         int32_t* n_blocks, int64_t* gather_floats);
int dh_kfac_damped_factors(dh_plan* plan, const float* stats, const float* dense0_xtx, float weight, float damping,
                           float* coef, float* mats_small, float* mats_large, void* stream);
int dh_kfac_update(dh_plan* plan, const float* inv_small, const float* inv_large, const float* coef, const float* stats,
                   float weight, float damping, const float* grads, float* out, void* ws, size_t ws_bytes, void* stream);

/* In-place inverse of `batch` symmetric positive-definite n x n fp32 matrices (row-major, contiguous), n <= 1024: the
 * damped Kronecker factors of the KFAC update (Gauss-Jordan without pivoting, one block per matrix). */
int dh_spd_inverse(float* mats, int32_t n, int32_t batch, void* stream);

/* Batched complex slogdet with the reference's multi-determinant tail (psiformer.py:74-76).
 *   mats: (B, K, n, n) complex64, row-major.
 *   out_sign (B,K) complex64 and out_logabs (B,K) f32 may be NULL;
 *   out_logpsi (B) complex64 = log sum_k sign_k exp(logabs_k) may be NULL. */
int dh_slogdet(const float* mats, int64_t B, int32_t K, int32_t n, float* out_sign,
               float* out_logabs, float* out_logpsi, void* stream);

/* fp32-accurate GEMM used by the network (exposed for parity tests, the roofline bench and the KFAC preconditioner):
 *   C[M,N] = A[M,K] @ W[K,N] (+ bias[N] on rows r with r % rows_per_group == 0) (+ C if accumulate)
 *   impl: 0 = fp32 FMA (any shape, no workspace);
 *         1 = tcgen05 / TMEM / TMA with fp16 hi/lo pieces, one accumulator per tile ("3xFP16", the plan ops' default);
 *         2 = tcgen05 with TF32 hi/lo pieces ("3xTF32");  3 = fp16 pieces, separate main / correction accumulators.
 *   impl >= 1 needs K % 32 == 0 and a caller workspace of dh_gemm_workspace_bytes (16-byte aligned) for the split
 *   weight planes.  Stream-ordered; no allocation and no host synchronisation. */
int dh_gemm_workspace_bytes(int32_t N, int32_t K, int32_t impl, size_t* bytes);
int dh_gemm(const float* A, const float* W, const float* bias, float* C, int64_t M, int32_t N,
            int32_t K, int32_t rows_per_group, int32_t accumulate, int32_t impl, void* ws, size_t ws_bytes,
            void* stream);

/* Debug: after dh_logpsi / dh_local_energy with B <= chunk, intermediate buffers live in
 * the workspace.  Returns the float offset and count of a named buffer ("h", "qkv", "attn",
 * "orb", "ld", "lpjet"); <0 if unknown. */
int dh_debug_buffer(const dh_plan* plan, int op, int64_t B, const char* name, int64_t* offset,
                    int64_t* count);

/* Instrumentation used by bench.py.  dh_launch_count: kernels launched through this plan so
 * far.  dh_profile_begin/end: bracket every kernel launch of the following calls with CUDA
 * events on the launching stream; dh_profile_end synchronises and returns, per category
 * (0 dense contractions, 1 attention, 2 layernorm, 3 envelope+logdet+assembly, 4 mcmc, 5 other),
 * the summed milliseconds, the number of timed scopes and the algorithmic flops. Arrays of 6. */
long long dh_launch_count(const dh_plan* plan);
int dh_profile_begin(dh_plan* plan, int32_t max_launches);
int dh_profile_end(dh_plan* plan, double* ms_host, int32_t* count_host, double* flops_host);

/* ---- walker-ensemble estimators (netobs_bridge/observables).  Plan-free: they read walkers or
 * log-amplitudes that the entry points above produced.
 *
 * dh_pair_correlation  <- PairCorrelationEstimator.evaluate (netobs_bridge/observables/pair_corr.py:42-60):
 *   state_inout[bins] (f32, device) += histogram over [0, pi] of theta_12 = arccos(r_i . r_j), all pairs i < j of
 *   all B walkers, weights 1 / sin(theta_12), times 4 bins / (batch_norm N^2 pi).  batch_norm: the batch size the
 *   normalisation divides by (the global batch when walkers are sharded over ranks and the states are summed);
 *   <= 0 -> B.  hist_ws: `bins` doubles of device scratch.  bins <= 4096.
 * dh_density_histogram <- DensityEstimator.evaluate (netobs_bridge/observables/density.py:42-49):
 *   counts_inout[bins] (u64, device) += counts of theta over all B N electrons, `bins` equal bins over [0, pi].
 * dh_overlap_sum / dh_overlap_ratio <- OverlapEstimator.evaluate (netobs_bridge/observables/overlap.py:55-63):
 *   logphi, logpsi: (B,2) f32 = (Re, Im) as dh_logpsi writes them.  dh_overlap_sum: out_sum[2] (f64, device) =
 *   sum_b (logphi_b - logpsi_b); the host divides by the (global) batch to get `shift` (after an all-reduce when
 *   sharded).  dh_overlap_ratio: ratio_b = exp(logphi_b - logpsi_b - shift) as (B,2) f32 and |ratio_b|^2 as (B) f32
 *   (either output may be NULL); shift: 2 doubles on the device. */
int dh_pair_correlation(const float* x, int64_t B, int32_t N, int32_t bins, int64_t batch_norm,
                        float* state_inout, double* hist_ws, void* stream);
int dh_density_histogram(const float* x, int64_t B, int32_t N, int32_t bins,
                         unsigned long long* counts_inout, void* stream);
int dh_overlap_sum(const float* logphi, const float* logpsi, int64_t B, double* out_sum, void* stream);
int dh_overlap_ratio(const float* logphi, const float* logpsi, int64_t B, const double* shift,
                     float* out_ratio, float* out_ratio_square, void* stream);

/* One-body reduced density matrix <- OneRDMEstimator.eval_product (netobs_bridge/observables/one_rdm.py:91-109).
 * dh_lll_orbitals:    out_phi (n, flux+1, 2) f32 = Y_{Q,Q,m}(points), m = -Q..Q, Q = flux/2, exactly as
 *                     make_monopole_harm(Q, Q, m) defines them (one_rdm.py:34-58, incl. the clip of cos theta).
 * dh_one_rdm_scatter: out_x_prime (B, N, N, 2): copy a of walker b = the walker with electron a moved to
 *                     r_prime[b] (B, 2) (one_rdm.py:92-94); feed it to dh_logpsi as B*N walkers.
 * dh_one_rdm_product: per walker rdm_ij = 4 pi sum_a exp(logpsi'_a - logpsi) phi_i(r_a) conj(phi_j(r')) with
 *                     logpsi (B,2), logpsi_prime (B,N,2), phi (B,N,L,2), phi_prime (B,L,2);
 *                     out_rdm (B,L,L,2) f32 and/or out_sum_inout (L,L,2) f64 += the sum over non-NaN walkers. */
int dh_lll_orbitals(const float* points, int64_t n, int32_t flux, float* out_phi, void* stream);
int dh_one_rdm_scatter(const float* x, const float* r_prime, int64_t B, int32_t N, float* out_x_prime,
                       void* stream);
int dh_one_rdm_product(const float* logpsi, const float* logpsi_prime, const float* phi,
                       const float* phi_prime, int64_t B, int32_t N, int32_t L, float* out_rdm,
                       double* out_sum_inout, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPHALL_B200_H */
