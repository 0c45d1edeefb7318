// Internal: plan object, workspace carving and profiling hooks shared by api.cu / api_vjp.cu.
#pragma once
#include <string.h>

#include <string>
#include <vector>

#include "../../include/deephall_b200.h"
#include "kernels.h"

using namespace dh;

struct LayerOff {
  int64_t q_k, q_b, k_k, k_b, v_k, v_b, o_k, o_b, d1_k, ln0_s, ln0_b, d2_k, d2_b, ln1_s, ln1_b;
};

struct KfLayer { int q, k, v, o, d1, d2, ln0s, ln0b, ln1s, ln1b; };  // indices into dh_plan::kfac

struct dh_plan {
  dh_config cfg;
  // Metropolis sweep as a replayed CUDA graph: one captured move (propose -> log psi -> accept -> advance) per
  // (walker buffer, workspace, batch); the move's scalar arguments live in d_mcmc, the parameters it reads in raw_params
  struct MoveGraph { const void* x; const void* ws; int64_t B; cudaGraphExec_t exec; long long launches; };
  std::vector<MoveGraph> move_graphs;
  McmcDev* d_mcmc = nullptr;
  float* raw_params = nullptr;   // plan-owned copy of the parameter vector (address-stable across optimizer steps)
  int graphs_ok = 1;             // cleared if a capture ever fails: the sweep then launches its kernels one by one
  unsigned* d_status = nullptr;  // device status word (dh_plan_status): bit 0 = an fp16 operand piece saturated
  int N, L, K, D, H, hd, nl, twoQ, LNK;
  int nsb;   // spin blocks with their own orbital projections (blocks.py:29-34): 1 (n_dn = 0) or 2
  int orbN;  // columns of the orbital-coefficient tensor: 2 * nsb * LNK = [re | im] per spin block
  int laughlin;  // analytic Laughlin ground state instead of the Psiformer (no parameters)
  int twoQ1;     // laughlin: 2 Q1 = flux - 2 p (N - 1) = N - 1 (ground state), N (quasihole) or N - 2 (quasiparticle)
  int lskip;     // laughlin quasihole: exponent index Q1 - lz of the orbital that is left out (-1: ground state)
  int lqp = 0, lqp_a = 0;  // laughlin quasiparticle (N = 2 Q1 + 2): flag and the u exponent Q1 + lz of the projected orbital
  float Q, radius;
  std::vector<dh_param_entry> entries;
  int64_t nparams;
  int64_t off_W0;
  std::vector<LayerOff> layer;
  int64_t orb_k[4], orb_b[4];  // DenseGeneral_{2 sb + part}: spin block sb, part 0 = real, 1 = imaginary
  int64_t ee_par, ee_anti;
  // sparse orbitals (blocks.py:52-62): the projections have 8 features instead of L and a real (8, L) map
  // `lll_weight` follows; both are linear, so prepare_weights folds them into effective full kernels
  // [2 nsb][(D + 1) rows: D kernel rows + the bias row][LNK] kept at prep + orb_eff (gradients: prep + orb_geff)
  int sparse;
  int64_t lll_k, lll_b;
  size_t orb_eff, orb_geff;
  // KFAC curvature blocks (dh_kfac_layout / dh_kfac_factors)
  std::vector<dh_kfac_entry> kfac;
  std::vector<KfLayer> kf_layer;
  int kf_dense0, kf_orb[4], kf_eepar, kf_eeanti, kf_lllk, kf_lllb;
  int64_t kfac_floats;
  // The activations of the last reverse pass's forward (dh_logpsi_vjp / dh_kfac_factors, single chunk) are still in the
  // caller's workspace: dh_kfac_factors_reuse_forward may skip its forward.  Cleared by every other op of the plan.
  struct FwdKey { const float* P = nullptr; const float* x = nullptr; int64_t B = 0; void* ws = nullptr; bool valid = false; } vjp_fwd;
  // KFAC update tables (dh_kfac_update_shape builds them on first use; device copies freed with the plan)
  struct KfUpdate {
    int ready = 0;  // 0 not built, 1 built, -1 unsupported (a factor with more than 1024 rows)
    std::vector<KfBlkDesc> blk;
    std::vector<KfMatDesc> mat;
    std::vector<KfDiagDesc> diag;
    KfBlkDesc* d_blk = nullptr;
    KfMatDesc* d_mat = nullptr;
    KfDiagDesc* d_diag = nullptr;
    int n_small = 0, dim_small = 0, n_large = 0, dim_large = 0;
    int64_t gather_floats = 0;
  } kfu;
  double* d_normfac;
  int gemm_impl;  // 0 = SIMT fp32 FMA, 1 = tcgen05 (two-piece operand split)
  int tc_merged;  // tcgen05 path: 1 = one double-buffered accumulator per tile, 0 = main + correction accumulators
  int tc_f16;     // tcgen05 path: 1 = kind::f16 pieces (D % 64 == 0), 0 = kind::tf32 pieces
  int ln_value_fuse = 0;  // value-only passes: residual + LayerNorm as the epilogue of the 256-wide contractions
  int orb_fuse = 0;    // jet passes at N = 12: the envelope contraction runs as the epilogue of the orbital projection
  size_t orb_perm = 0; // prep offset of the permuted fp32 orbital kernel + bias (staging for the split)
  int a_planes;   // jet passes keep the activations that feed a contraction as fp16 hi / lo planes (fp16 pieces, D = 256)
  // prepared (pre-split, transposed) weights for the tcgen05 path
  float* prep;            // device buffer owned by the plan
  size_t prep_floats;
  const float* prep_src;  // params pointer the preparation was made from
  int auto_prepare;       // 1 (default): every op refreshes the prepared weights itself; 0: the caller does
                          // (dh_params_prepare) after each parameter update
  bool prep_fwd_valid, prep_vjp_valid;  // what dh_params_prepare has produced for prep_src
  struct Slot { int Nout; size_t hi, lo, bias, scale; int ldw; };  // float offsets into prep (bias: SIZE_MAX = none); scale: 3 floats; ldw: plane row length
  std::vector<Slot> slots;  // per layer: qkv, o, d1, d2, od (= o folded into d1) ; last: orbitals (re | im)
  // reverse pass (dX = G @ W^T): planes of W itself, [D][Kpad]; per layer d2, d1, o, qkv ; last: orbitals
  std::vector<Slot> vslots;
  size_t cot_scale;         // float offset into prep: {s, 1/s} for the cotangents of the current dh_logpsi_vjp call
  size_t w0qkv;             // float offset into prep: [4][3D] = W0 @ (Wq|Wk|Wv) of layer 0 (fp32)
  size_t fold_tmp;          // float offset into prep: [D][D] scratch for Wo @ W1 (fp32)
  // ---- instrumentation (dh_profile_*): CUDA-event timing of kernel categories, launch count
  bool prof_on;
  std::vector<cudaEvent_t> prof_ev;   // pool, pairs (start, stop)
  std::vector<int> prof_cat;          // category of pair i
  std::vector<double> prof_flops;     // algorithmic flops of pair i
  size_t prof_used;
  long long launches;
  // ---- chunk interleave: a pass of two or more chunks alternates them between the caller's stream and this one
  // (each with its own activation workspace), so one chunk's launch gaps and wave tails are filled by the other's kernels
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

// 0 disables the chunk interleave (DH_DUAL_STREAM=0)
static inline bool dual_stream_env() {
  static const bool on = !(dbg_env("DH_DUAL_STREAM") && atoi(dbg_env("DH_DUAL_STREAM")) == 0);
  return on;
}

enum ProfCat { PC_GEMM = 0, PC_ATTENTION = 1, PC_LAYERNORM = 2, PC_TAIL = 3, PC_MCMC = 4, PC_OTHER = 5, PC_COUNT = 6 };

// Wraps one kernel launch: counts it and, when profiling, brackets it with events.
struct ProfScope {
  dh_plan* p; cudaStream_t s; bool active;
  ProfScope(const dh_plan* plan, int cat, double flops, cudaStream_t st, int nlaunch = 1)
      : p(const_cast<dh_plan*>(plan)), s(st), active(false) {
    p->launches += nlaunch;
    if (p->prof_on && p->prof_used + 2 <= p->prof_ev.size()) {
      active = true;
      p->prof_cat[p->prof_used / 2] = cat;
      p->prof_flops[p->prof_used / 2] = flops;
      cudaEventRecord(p->prof_ev[p->prof_used], s);
    }
  }
  ~ProfScope() {
    if (active) { cudaEventRecord(p->prof_ev[p->prof_used + 1], s); p->prof_used += 2; }
  }
};

// --------------------------------------------------------------------------------- workspace
struct FwdWs {
  float *h, *t1, *t2, *qkv, *att, *cbuf, *Mj, *ld, *lpjet, *Minv;
  size_t floats;
};

static inline size_t al(size_t n) { return (n + 63) / 64 * 64; }  // 256-byte granules

// Walkers per internal pass.  Default: as many as keep the pass's activation workspace near 6 GiB
// (jets: 1024 at c3, 256 at c5 with 4 determinants), so the workspace does not grow with B.
static inline int64_t pick_chunk(const dh_plan* p, bool jets, int64_t B) {
  int64_t c = p->cfg.chunk_walkers;
  if (c <= 0) {
    const int R = jets ? 2 * p->N + 8 : 1;
    const double per_walker = (double)p->N * R * (7.0 * p->D + (double)p->orbN) * sizeof(float) +
                              (double)p->K * R * p->N * p->N * 2 * sizeof(float);
    c = (int64_t)(6.0 * 1024 * 1024 * 1024 / per_walker);
    const int64_t cap = jets ? 1024 : 16384;
    if (c > cap) c = cap;
    c = c / 32 * 32;
    if (c < 32) c = 32;
  }
  return B < c ? B : c;
}

// Chunk interleave plan of a forward pass: walkers per chunk and the number of activation workspaces (1 or 2).
// Two or more chunks alternate between two streams.  A value-only pass (Metropolis sweep, log psi) that fits one chunk
// is cut in two halves when it has at least DH_DUAL_VALUE_MIN walkers (default 2048; 0 = never), so that the two
// halves' short launches fill each other's gaps.
static inline int64_t plan_chunks(const dh_plan* p, bool jets, int64_t B, int* copies) {
  int64_t chunk = pick_chunk(p, jets, B);
  *copies = 1;
  if (!dual_stream_env()) return chunk;
  if (B > chunk) { *copies = 2; return chunk; }
  static const long vmin = dbg_env("DH_DUAL_VALUE_MIN") ? atol(dbg_env("DH_DUAL_VALUE_MIN")) : 2048;
  if (!jets && vmin > 0 && B >= vmin && p->cfg.chunk_walkers <= 0) {
    chunk = ((B + 1) / 2 + 31) / 32 * 32;
    *copies = 2;
  }
  return chunk;
}

static inline FwdWs carve_fwd(const dh_plan* p, float* base, int64_t Bc, bool jets, bool keep_inverse) {
  const int R = jets ? 2 * p->N + 8 : 1;
  const size_t rows = p->laughlin ? 0 : (size_t)Bc * p->N * R;  // the analytic network has no body activations
  FwdWs w;
  size_t off = 0;
  auto take = [&](size_t n) { float* q = base ? base + off : nullptr; off += al(n); return q; };
  w.h = take(rows * p->D);
  w.t1 = take(rows * p->D);
  w.t2 = take(rows * p->D);
  w.qkv = take(rows * 3 * p->D);
  w.att = take(rows * p->D);
  w.cbuf = take(rows * (size_t)p->orbN);
  w.Mj = take((size_t)Bc * p->K * R * p->N * p->N * 2);
  w.ld = take((size_t)Bc * p->K * R * 2);
  w.lpjet = take((size_t)Bc * R * 2);
  w.Minv = keep_inverse ? take((size_t)Bc * p->K * p->N * p->N * 2) : nullptr;
  w.floats = off;
  return w;
}

struct McmcWs {
  float *x2, *lp1, *logpsi2;
  size_t floats;
};
static inline McmcWs carve_mcmc(const dh_plan* p, float* base, int64_t B) {
  McmcWs w;
  size_t off = 0;
  auto take = [&](size_t n) { float* q = base ? base + off : nullptr; off += al(n); return q; };
  w.x2 = take((size_t)B * p->N * 2);
  w.lp1 = take((size_t)B);
  w.logpsi2 = take((size_t)B * 2);
  w.floats = off;
  return w;
}

static inline float* align_ws(void* ws) {
  uintptr_t a = (reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255;
  return reinterpret_cast<float*>(a);
}

// kernel [D][LNK] and bias [LNK] of orbital projection t = 2 sb + part, as the contractions see them
static inline const float* orbW(const dh_plan* p, const float* P, int t) {
  return p->sparse ? p->prep + p->orb_eff + (size_t)t * (p->D + 1) * p->LNK : P + p->orb_k[t];
}
static inline const float* orbB(const dh_plan* p, const float* P, int t) {
  return p->sparse ? orbW(p, P, t) + (size_t)p->D * p->LNK : P + p->orb_b[t];
}
// where the reverse pass accumulates d/d(kernel) and d/d(bias) of projection t
static inline float* orbGW(const dh_plan* p, float* grad, int t) {
  return p->sparse ? p->prep + p->orb_geff + (size_t)t * (p->D + 1) * p->LNK : grad + p->orb_k[t];
}
static inline float* orbGB(const dh_plan* p, float* grad, int t) {
  return p->sparse ? orbGW(p, grad, t) + (size_t)p->D * p->LNK : grad + p->orb_b[t];
}

// --------------------------------------------------------------------------------- forward
enum { SL_QKV = 0, SL_O = 1, SL_D1 = 2, SL_D2 = 3, SL_OD = 4, SL_PER_LAYER = 5 };
enum { VS_D2 = 0, VS_D1 = 1, VS_O = 2, VS_QKV = 3, VS_PER_LAYER = 4 };

// tcgen05 path: C[rows, Nout] (ldc) = A[rows, D] @ W_slot (+ bias on value rows).
// a_planes: A is the fp16 hi / lo plane view of a [rows][D] buffer (common.cuh), written by the producing kernel.

// A Metropolis move evaluated by a value-only pass (forward_chunk): current configuration / log-probability of the chunk's
// walkers (both updated by the accept step), the device block of move arguments, the chunk's first walker in the rank's batch.
struct MoveChunk { float* x1; float* lp1; McmcDev* dv; int64_t walker0; };
struct LnArgs { const float* gamma; const float* beta; int tanh_mode; };  // fused value LayerNorm epilogue: C = LN(C + f(A W + b)) in place
static inline int dense_tc(const dh_plan* p, const float* A, int slot, float* C, int64_t rows, int64_t ldc, int R,
                           cudaStream_t s, bool a_planes = false, const LnArgs* ln = nullptr) {
  const dh_plan::Slot& sl = p->slots[slot];
  ProfScope ps(p, PC_GEMM, 2.0 * (double)rows * sl.Nout * p->D, s);
  TcGemm g;
  g.A = A; g.lda = p->D; g.Wt_hi = p->prep + sl.hi; g.Wt_lo = p->prep + sl.lo; g.ldw = p->D;
  g.bias = sl.bias == SIZE_MAX ? nullptr : p->prep + sl.bias;
  g.inv_scale = p->tc_f16 ? p->prep + sl.scale + 1 : nullptr;
  g.C = C; g.ldc = ldc; g.M = rows; g.N = sl.Nout; g.K = p->D; g.rpg = R;
  g.f16 = p->tc_f16; g.merged = p->tc_merged; g.reduce_add = 0; g.a_scale = nullptr;
  g.A_lo = a_planes ? reinterpret_cast<const __half*>(A) + rows * p->D : nullptr;
  g.orb_env = nullptr; g.orb_Mj = nullptr; g.orb_L = 0;
  g.ln_res = ln ? C : nullptr; g.ln_gamma = ln ? ln->gamma : nullptr; g.ln_beta = ln ? ln->beta : nullptr; g.ln_tanh = ln ? ln->tanh_mode : 0;
  return gemm_tc_ex(g, s);
}

// SIMT path: C[rows, Nout] (ldc) = A[rows, D] @ W[D, Nout] (+ bias on value rows)
static inline int dense(const dh_plan* p, const float* A, const float* W, const float* bias, float* C, int64_t rows,
                 int Nout, int64_t ldc, int R, cudaStream_t s) {
  ProfScope ps(p, PC_GEMM, 2.0 * (double)rows * Nout * p->D, s);
  return gemm_simt(A, W, bias, C, rows, Nout, p->D, p->D, 1, Nout, 1, ldc, R, 0, 1, s);
}


// The five dense contractions of the network, on whichever implementation the plan selected.
static inline int dense_qkv(const dh_plan* p, const float* P, int l, const float* A, float* qkv, int64_t rows, int R,
                            cudaStream_t s, bool a_planes = false) {
  const int D = p->D;
  if (p->gemm_impl == 1) return dense_tc(p, A, l * SL_PER_LAYER + SL_QKV, qkv, rows, 3 * D, R, s, a_planes);
  const LayerOff& o = p->layer[l];
  int rc;
  if ((rc = dense(p, A, P + o.q_k, P + o.q_b, qkv, rows, D, 3 * D, R, s))) return rc;
  if ((rc = dense(p, A, P + o.k_k, P + o.k_b, qkv + D, rows, D, 3 * D, R, s))) return rc;
  return dense(p, A, P + o.v_k, P + o.v_b, qkv + 2 * D, rows, D, 3 * D, R, s);
}
static inline int dense_layer(const dh_plan* p, const float* P, int l, int which, const float* A, float* C,
                              int64_t rows, int R, cudaStream_t s, bool a_planes = false) {
  const int D = p->D;
  if (p->gemm_impl == 1) return dense_tc(p, A, l * SL_PER_LAYER + which, C, rows, D, R, s, a_planes);
  const LayerOff& o = p->layer[l];
  if (which == SL_O) return dense(p, A, P + o.o_k, P + o.o_b, C, rows, D, D, R, s);
  if (which == SL_D1) return dense(p, A, P + o.d1_k, nullptr, C, rows, D, D, R, s);
  return dense(p, A, P + o.d2_k, P + o.d2_b, C, rows, D, D, R, s);
}
static inline int dense_orb(const dh_plan* p, const float* P, const float* A, float* cbuf, int64_t rows, int R,
                            cudaStream_t s, bool a_planes = false) {
  const int LNK = p->LNK;
  // every row gets the projections of every spin block (columns [2 sb LNK, 2 (sb+1) LNK)); the tail kernels read
  // the block of the row's electron
  if (p->gemm_impl == 1) return dense_tc(p, A, p->nl * SL_PER_LAYER, cbuf, rows, p->orbN, R, s, a_planes);
  for (int t = 0; t < 2 * p->nsb; ++t) {
    int rc = dense(p, A, orbW(p, P, t), orbB(p, P, t), cbuf + (size_t)t * LNK, rows, LNK, p->orbN, R, s);
    if (rc) return rc;
  }
  return 0;
}

int prepare_weights(dh_plan* p, const float* P, cudaStream_t s);      // api.cu
int prepare_weights_vjp(dh_plan* p, const float* P, cudaStream_t s);  // api.cu
