// dh_logpsi_vjp: one reverse pass through the value-only network for a batch of walkers,
// contracting d(Re, Im log psi_b)/d params with per-walker cotangents (loss.py:53-64,96-106).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "plan.h"

namespace {

struct VjpWs {
  std::vector<float*> hs, qkv, att, t1, t2, hA, z;
  float *cbuf, *Mj, *ld, *Minv, *logpsi, *lpjet;
  float *gH, *gA, *gB, *gC, *gQKV, *gCb;
  size_t floats;
};

VjpWs carve_vjp(const dh_plan* p, float* base, int64_t Bc) {
  const size_t rows = (size_t)Bc * p->N;
  const int D = p->D, nl = p->nl;
  VjpWs w;
  size_t off = 0;
  auto take = [&](size_t n) { float* q = base ? base + off : nullptr; off += al(n); return q; };
  w.hs.resize(nl + 1);
  w.qkv.resize(nl); w.att.resize(nl); w.t1.resize(nl); w.t2.resize(nl); w.hA.resize(nl); w.z.resize(nl);
  for (int l = 0; l <= nl; ++l) w.hs[l] = take(rows * D);
  for (int l = 0; l < nl; ++l) {
    w.qkv[l] = take(rows * 3 * D);
    w.att[l] = take(rows * D);
    w.t1[l] = take(rows * D);
    w.t2[l] = take(rows * D);
    w.hA[l] = take(rows * D);
    w.z[l] = take(rows * D);
  }
  w.cbuf = take(rows * (size_t)p->orbN);
  w.Mj = take((size_t)Bc * p->K * p->N * p->N * 2);
  w.ld = take((size_t)Bc * p->K * 2);
  w.Minv = take((size_t)Bc * p->K * p->N * p->N * 2);
  w.logpsi = take((size_t)Bc * 2);
  w.lpjet = take((size_t)Bc * 2);
  w.gH = take(rows * D);
  w.gA = take(rows * D);
  w.gB = take(rows * D);
  w.gC = take(rows * D);
  w.gQKV = take(rows * 3 * D);
  w.gCb = take(rows * (size_t)p->orbN);
  w.floats = off;
  return w;
}

// C[rows, Din] (=|+=) G[rows, Nout] (ldg) @ W[Din, Nout]^T
int dense_bwd_x(const dh_plan* p, const float* G, int64_t ldg, const float* W, int Nout, float* C, int64_t rows,
                int Din, int accumulate, cudaStream_t s) {
  ProfScope ps(p, PC_GEMM, 2.0 * (double)rows * Nout * Din, s);
  return gemm_simt(G, W, nullptr, C, rows, Din, Nout, ldg, 1, 1, Nout, Din, 1, accumulate, 1, s);
}

// tensor-core form of the same product through the prepared reverse-pass planes of slot `vs` (K = Nout columns of G)
bool bwd_x_tc_ok(const dh_plan* p, const float* G, int64_t ldg, const float* C) {
  return p->gemm_impl == 1 && (ldg % 4) == 0 && ((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(C)) & 15) == 0;
}
int dense_bwd_x_tc(const dh_plan* p, const float* G, int64_t ldg, int vs, int K, float* C, int64_t rows, int accumulate,
                   cudaStream_t s) {
  const dh_plan::Slot& sl = p->vslots[vs];
  ProfScope ps(p, PC_GEMM, 2.0 * (double)rows * K * p->D, s);
  TcGemm g;
  g.A = G; g.lda = ldg;
  g.Wt_hi = p->prep + sl.hi; g.Wt_lo = p->prep + sl.lo; g.ldw = sl.ldw;
  g.bias = nullptr; g.inv_scale = p->tc_f16 ? p->prep + sl.scale + 1 : nullptr;
  g.C = C; g.ldc = p->D; g.M = rows; g.N = p->D; g.K = K; g.rpg = 1;
  g.f16 = p->tc_f16; g.merged = p->tc_merged; g.reduce_add = accumulate;
  g.A_lo = nullptr;
  g.orb_env = nullptr; g.orb_Mj = nullptr; g.orb_L = 0; g.ln_res = nullptr; g.ln_gamma = nullptr; g.ln_beta = nullptr; g.ln_tanh = 0;
  g.a_scale = p->prep + p->cot_scale;  // gradients scale with the cotangents (O(1/B)): keep the fp16 pieces in range
  return gemm_tc_ex(g, s);
}

// dW[Din, Nout] += X[rows, Din]^T @ G[rows, Nout] (ldg)
// The "TN" contractions of the reverse pass (rows = the contracted index) run on the tensor cores when the operands
// allow it (16-byte alignment, fp16 pieces); DH_VJP_DW=simt forces the fp32-FMA split-K kernel.
bool tn_tc_ok(const dh_plan* p, const float* A, int64_t lda, const float* B, int64_t ldb, const float* C, int64_t ldc, int Ma) {
  static const bool simt = dbg_env("DH_VJP_DW") && strcmp(dbg_env("DH_VJP_DW"), "simt") == 0;
  return !simt && p->gemm_impl == 1 && p->tc_f16 && Ma >= 32 && gemm_tn_tc_ok(A, lda, B, ldb, C, ldc);
}

// dW[Din, Nout] += X[rows, Din]^T @ G[rows, Nout] (ldg); G is cotangent-sized, scaled by the plan's cotangent scale
int dense_bwd_w(const dh_plan* p, const float* X, int Din, const float* G, int64_t ldg, int Nout, float* dW,
                int64_t rows, cudaStream_t s) {
  ProfScope ps(p, PC_GEMM, 2.0 * (double)rows * Nout * Din, s);
  if (tn_tc_ok(p, X, Din, G, ldg, dW, Nout, Din))
    return gemm_tn_tc(X, Din, Din, G, ldg, Nout, dW, Nout, rows, nullptr, p->prep + p->cot_scale, s);
  const int tiles = ((Din + 127) / 128) * ((Nout + 127) / 128);
  int split = 296 / tiles;
  if (split < 1) split = 1;
  const int64_t max_split = (rows + 127) / 128;
  if (split > max_split) split = (int)max_split;
  return gemm_simt(X, G, nullptr, dW, Din, Nout, rows, 1, Din, ldg, 1, Nout, 1, 1, split, s);
}

// out[dim][dim] += X^T X over `rows` rows of X (row stride ldx): a Kronecker factor sum of the KFAC curvature blocks.
// cot_sized: X is a gradient (scaled like the cotangents before the fp16 split)
int gram(const dh_plan* p, const float* X, int64_t ldx, int dim, float* out, int64_t rows, cudaStream_t s, bool cot_sized = false) {
  ProfScope ps(p, PC_GEMM, 2.0 * (double)rows * dim * dim, s);
  if (tn_tc_ok(p, X, ldx, X, ldx, out, dim, dim)) {
    const float* sc = cot_sized ? p->prep + p->cot_scale : nullptr;
    return gemm_tn_tc(X, ldx, dim, X, ldx, dim, out, dim, rows, sc, sc, s);
  }
  const int tiles = ((dim + 127) / 128) * ((dim + 127) / 128);
  int split = 296 / tiles;
  if (split < 1) split = 1;
  const int64_t max_split = (rows + 127) / 128;
  if (split > max_split) split = (int)max_split;
  return gemm_simt(X, X, nullptr, out, dim, dim, rows, 1, ldx, ldx, 1, dim, 1, 1, split, s);
}

}  // namespace

size_t vjp_ws_floats(const dh_plan* p, int64_t Bc) { return carve_vjp(p, nullptr, Bc).floats; }

// One forward + reverse pass.  kf == nullptr: the parameter gradient for per-walker cotangents `cot` (dh_logpsi_vjp).
// kf != nullptr: cotangent (1, 0) for every walker and, instead of parameter gradients, the factor sums of the KFAC
// curvature blocks (dh_kfac_factors); `cot` is ignored and the parameter-shaped by-products go to a scratch vector.
static int vjp_core(dh_plan* p, const float* P, const float* x, int64_t B, const float* cot, float* grad, float* kf,
                    float* out_logpsi, void* ws, size_t ws_bytes, cudaStream_t s, bool reuse_forward = false) {
  // the forward of the previous reverse pass may be reused when it ran on the same parameters, walkers and workspace, as one chunk
  const bool skip_fwd = reuse_forward && p->vjp_fwd.valid && p->vjp_fwd.P == P && p->vjp_fwd.x == x && p->vjp_fwd.B == B &&
                        p->vjp_fwd.ws == ws && !out_logpsi;
  p->vjp_fwd.valid = false;
  if (kf) DH_CHECK(cudaMemsetAsync(kf, 0, (size_t)p->kfac_floats * sizeof(float), s));
  if (!kf) DH_CHECK(cudaMemsetAsync(grad, 0, p->nparams * sizeof(float), s));
  if (B == 0) return 0;
  if (p->sparse)
    DH_CHECK(cudaMemsetAsync(p->prep + p->orb_geff, 0, (size_t)2 * p->nsb * (p->D + 1) * p->LNK * sizeof(float), s));
  const int64_t chunk = pick_chunk(p, false, B);
  float* base = align_ws(ws);
  VjpWs w = carve_vjp(p, base, chunk);
  size_t need = w.floats;
  float* unit_cot = nullptr;
  if (kf) {  // extra workspace: the unit cotangents and a parameter-shaped scratch vector
    unit_cot = base + need; need += al((size_t)chunk * 2);
    grad = base + need; need += al((size_t)p->nparams);
  }
  if (!ws || (size_t)((char*)(base + need) - (char*)ws) > ws_bytes) return DH_E_WORKSPACE;
  if (kf) {
    if (int rcu = fill_unit_cot(unit_cot, chunk, s)) return rcu;
    DH_CHECK(cudaMemsetAsync(grad, 0, p->nparams * sizeof(float), s));
  }
  auto K = [&](int idx) -> const dh_kfac_entry& { return p->kfac[idx]; };
  const int N = p->N, D = p->D, nl = p->nl, LNK = p->LNK;
  NetDims nd{N, 1, D, p->H, p->hd, p->cfg.n_up};
  TailDims td{N, 1, p->L, p->K, p->twoQ, p->cfg.n_up, p->cfg.n_dn};
  int rc;
  if ((rc = prepare_weights(p, P, s))) return rc;
  if ((rc = prepare_weights_vjp(p, P, s))) return rc;
  if (p->gemm_impl == 1 && (rc = pow2_scale_tc(kf ? unit_cot : cot, 2 * (kf ? chunk : B), p->prep + p->cot_scale, s))) return rc;
#define RUN(cat, call)                          \
  do {                                          \
    ProfScope _ps(p, cat, 0, s);                \
    if ((rc = (call))) return rc;               \
  } while (0)

  for (int64_t b0 = 0; b0 < B; b0 += chunk) {
    const int64_t Bc = (B - b0) < chunk ? (B - b0) : chunk;
    const int64_t rows = Bc * N;
    const float* xc = x + b0 * N * 2;
    const float* cotc = kf ? unit_cot : cot + b0 * 2;
    // ------------------------------------------------------------ forward, keeping activations
    const float* hf = w.hs[nl];
    if (!skip_fwd) {
    RUN(PC_OTHER, features_dense0(xc, P + p->off_W0, w.hs[0], Bc, nd, s));
    for (int l = 0; l < nl; ++l) {
      const LayerOff& o = p->layer[l];
      if ((rc = dense_qkv(p, P, l, w.hs[l], w.qkv[l], rows, 1, s))) return rc;
      RUN(PC_ATTENTION, attention_value(w.qkv[l], w.att[l], Bc, nd, s));
      if ((rc = dense_layer(p, P, l, SL_O, w.att[l], w.t1[l], rows, 1, s))) return rc;
      if ((rc = dense_layer(p, P, l, SL_D1, w.t1[l], w.t2[l], rows, 1, s))) return rc;
      RUN(PC_LAYERNORM, residual_layernorm(w.hs[l], w.t2[l], P + o.ln0_s, P + o.ln0_b, w.hA[l], Bc, nd, 0, s));
      if ((rc = dense_layer(p, P, l, SL_D2, w.hA[l], w.z[l], rows, 1, s))) return rc;
      RUN(PC_LAYERNORM, residual_layernorm(w.hA[l], w.z[l], P + o.ln1_s, P + o.ln1_b, w.hs[l + 1], Bc, nd, 1, s));
    }
    if ((rc = dense_orb(p, P, hf, w.cbuf, rows, 1, s))) return rc;
    RUN(PC_TAIL, orbital_contract(w.cbuf, xc, p->d_normfac, w.Mj, Bc, td, s));
    RUN(PC_TAIL, logdet_jets_impl(w.Mj, w.ld, w.Minv, Bc, td, s));
    }
    if (out_logpsi) {
      FinalizeArgs fa;
      memset(&fa, 0, sizeof(fa));
      fa.ld = w.ld; fa.x = xc;
      fa.ee_par = p->ee_par >= 0 ? P + p->ee_par : nullptr;
      fa.ee_anti = p->ee_anti >= 0 ? P + p->ee_anti : nullptr;
      fa.Q = p->Q; fa.radius = p->radius;
      fa.interaction_strength = p->cfg.interaction_strength;
      fa.interaction_type = p->cfg.interaction_type;
      fa.out_logpsi = out_logpsi + b0 * 2;
      RUN(PC_TAIL, finalize(fa, Bc, td, s));
    }
    // ------------------------------------------------------------ backward
    RUN(PC_TAIL, tail_bwd(cotc, w.ld, w.Minv, xc, p->d_normfac, w.gCb, Bc, td, s));
    if (p->ee_par >= 0 || p->ee_anti >= 0)
      RUN(PC_TAIL, jastrow_bwd(cotc, xc, p->ee_par >= 0 ? P + p->ee_par : nullptr, p->ee_anti >= 0 ? P + p->ee_anti : nullptr,
                               p->ee_par >= 0 ? grad + p->ee_par : nullptr, p->ee_anti >= 0 ? grad + p->ee_anti : nullptr,
                               nullptr, nullptr, Bc, N, p->cfg.n_up, s));
    // orbital projections (tail_bwd leaves zeros in the columns of the spin block a row's electron does not use)
    const int64_t ldg = p->orbN;
    if (!kf) {
      for (int t = 0; t < 2 * p->nsb; ++t) {  // (sparse orbitals: into the effective-kernel gradients, unfolded after the loop)
        if ((rc = dense_bwd_w(p, hf, D, w.gCb + (size_t)t * LNK, ldg, LNK, orbGW(p, grad, t), rows, s))) return rc;
        RUN(PC_OTHER, colsum_add(w.gCb + (size_t)t * LNK, orbGB(p, grad, t), rows, LNK, ldg, s));
      }
    } else {
      for (int sbk = 0; sbk < p->nsb; ++sbk) {
        const dh_kfac_entry& e = K(p->kf_orb[2 * sbk]);
        const float* xin = hf;
        if (p->nsb == 2) {  // only this spin block's electrons pass through its projections (blocks.py:29-34)
          RUN(PC_OTHER, mask_rows_by_spin(hf, w.gA, rows, D, N, p->cfg.n_up, sbk, s));
          xin = w.gA;
        }
        if ((rc = gram(p, xin, D, D, kf + e.xtx_offset, rows, s))) return rc;
        RUN(PC_OTHER, colsum_add(xin, kf + e.xsum_offset, rows, D, D, s));
        for (int part = 0; part < 2; ++part) {
          const int t = 2 * sbk + part;
          if (!p->sparse) {
            if ((rc = gram(p, w.gCb + (size_t)t * LNK, ldg, LNK, kf + K(p->kf_orb[t]).gtg_offset, rows, s, true))) return rc;
          } else {
            // sparse orbitals: the block's output is the 8-feature tensor; its gradient is the effective coefficients'
            // gradient contracted with lll_weight (gQKV is free until the layer loop)
            const int F8 = 8 * N * p->K;
            RUN(PC_OTHER, sparse_g8(w.gCb + (size_t)t * LNK, ldg, P + p->lll_k, w.gQKV, rows, p->L, N * p->K, s));
            if ((rc = gram(p, w.gQKV, F8, F8, kf + K(p->kf_orb[t]).gtg_offset, rows, s, true))) return rc;
            // the batch-summed gradient of lll_weight (naive-diagonal blocks) comes out of the effective kernels' gradients
            if ((rc = dense_bwd_w(p, hf, D, w.gCb + (size_t)t * LNK, ldg, LNK, orbGW(p, grad, t), rows, s))) return rc;
            RUN(PC_OTHER, colsum_add(w.gCb + (size_t)t * LNK, orbGB(p, grad, t), rows, LNK, ldg, s));
          }
        }
      }
    }
    if (bwd_x_tc_ok(p, w.gCb, ldg, w.gH)) {
      if ((rc = dense_bwd_x_tc(p, w.gCb, ldg, nl * VS_PER_LAYER, p->orbN, w.gH, rows, 0, s))) return rc;
    } else {
      for (int t = 0; t < 2 * p->nsb; ++t)
        if ((rc = dense_bwd_x(p, w.gCb + (size_t)t * LNK, ldg, orbW(p, P, t), LNK, w.gH, rows, D, t > 0 ? 1 : 0, s))) return rc;
    }
    for (int l = nl - 1; l >= 0; --l) {
      const LayerOff& o = p->layer[l];
      const KfLayer* kl = kf ? &p->kf_layer[l] : nullptr;
      // h_out = LN1(hA + tanh(z)) : gH -> (gA = d/d hA, gB = d/d z)
      if (kf) RUN(PC_LAYERNORM, ln_fisher_diag(w.hA[l], w.z[l], w.gH, kf + K(kl->ln1s).diag_offset, kf + K(kl->ln1b).diag_offset, Bc, N, D, 1, s));
      RUN(PC_LAYERNORM, residual_layernorm_bwd(w.hA[l], w.z[l], P + o.ln1_s, w.gH, w.gA, w.gB, grad + o.ln1_s,
                                               grad + o.ln1_b, rows, D, 1, s));
      if (!kf) {
        if ((rc = dense_bwd_w(p, w.hA[l], D, w.gB, D, D, grad + o.d2_k, rows, s))) return rc;
        RUN(PC_OTHER, colsum_add(w.gB, grad + o.d2_b, rows, D, D, s));
      } else {
        if ((rc = gram(p, w.hA[l], D, D, kf + K(kl->d2).xtx_offset, rows, s))) return rc;
        RUN(PC_OTHER, colsum_add(w.hA[l], kf + K(kl->d2).xsum_offset, rows, D, D, s));
        if ((rc = gram(p, w.gB, D, D, kf + K(kl->d2).gtg_offset, rows, s, true))) return rc;
      }
      const bool tc = bwd_x_tc_ok(p, w.gB, D, w.gA) && bwd_x_tc_ok(p, w.gC, D, w.gH) && bwd_x_tc_ok(p, w.gQKV, 3 * D, w.gH);
      if (tc) { if ((rc = dense_bwd_x_tc(p, w.gB, D, l * VS_PER_LAYER + VS_D2, D, w.gA, rows, 1, s))) return rc; }
      else if ((rc = dense_bwd_x(p, w.gB, D, P + o.d2_k, D, w.gA, rows, D, 1, s))) return rc;
      // hA = LN0(h_in + t2) : gA -> (gH = d/d h_in, gB = d/d t2)
      if (kf) RUN(PC_LAYERNORM, ln_fisher_diag(w.hs[l], w.t2[l], w.gA, kf + K(kl->ln0s).diag_offset, kf + K(kl->ln0b).diag_offset, Bc, N, D, 0, s));
      RUN(PC_LAYERNORM, residual_layernorm_bwd(w.hs[l], w.t2[l], P + o.ln0_s, w.gA, w.gH, w.gB, grad + o.ln0_s,
                                               grad + o.ln0_b, rows, D, 0, s));
      if (!kf) {
        if ((rc = dense_bwd_w(p, w.t1[l], D, w.gB, D, D, grad + o.d1_k, rows, s))) return rc;
      } else {
        if ((rc = gram(p, w.t1[l], D, D, kf + K(kl->d1).xtx_offset, rows, s))) return rc;
        if ((rc = gram(p, w.gB, D, D, kf + K(kl->d1).gtg_offset, rows, s, true))) return rc;
      }
      if (tc) { if ((rc = dense_bwd_x_tc(p, w.gB, D, l * VS_PER_LAYER + VS_D1, D, w.gC, rows, 0, s))) return rc; }
      else if ((rc = dense_bwd_x(p, w.gB, D, P + o.d1_k, D, w.gC, rows, D, 0, s))) return rc;   // gC = d/d t1
      if (!kf) {
        if ((rc = dense_bwd_w(p, w.att[l], D, w.gC, D, D, grad + o.o_k, rows, s))) return rc;
        RUN(PC_OTHER, colsum_add(w.gC, grad + o.o_b, rows, D, D, s));
      } else {
        if ((rc = gram(p, w.att[l], D, D, kf + K(kl->o).xtx_offset, rows, s))) return rc;
        RUN(PC_OTHER, colsum_add(w.att[l], kf + K(kl->o).xsum_offset, rows, D, D, s));
        if ((rc = gram(p, w.gC, D, D, kf + K(kl->o).gtg_offset, rows, s, true))) return rc;
      }
      if (tc) { if ((rc = dense_bwd_x_tc(p, w.gC, D, l * VS_PER_LAYER + VS_O, D, w.gB, rows, 0, s))) return rc; }
      else if ((rc = dense_bwd_x(p, w.gC, D, P + o.o_k, D, w.gB, rows, D, 0, s))) return rc;    // gB = d/d att
      RUN(PC_ATTENTION, attention_value_bwd(w.qkv[l], w.gB, w.gQKV, Bc, nd, s));
      const int64_t qk[3] = {o.q_k, o.k_k, o.v_k};
      const int64_t qb[3] = {o.q_b, o.k_b, o.v_b};
      if (kf) {
        if ((rc = gram(p, w.hs[l], D, D, kf + K(kl->q).xtx_offset, rows, s))) return rc;
        RUN(PC_OTHER, colsum_add(w.hs[l], kf + K(kl->q).xsum_offset, rows, D, D, s));
        const int kq[3] = {kl->q, kl->k, kl->v};
        for (int t = 0; t < 3; ++t)
          if ((rc = gram(p, w.gQKV + t * D, 3 * D, D, kf + K(kq[t]).gtg_offset, rows, s, true))) return rc;
      }
      for (int t = 0; t < 3; ++t) {
        if (!kf) {
          if ((rc = dense_bwd_w(p, w.hs[l], D, w.gQKV + t * D, 3 * D, D, grad + qk[t], rows, s))) return rc;
          RUN(PC_OTHER, colsum_add(w.gQKV + t * D, grad + qb[t], rows, D, 3 * D, s));
        }
        if (!tc && (rc = dense_bwd_x(p, w.gQKV + t * D, 3 * D, P + qk[t], D, w.gH, rows, D, 1, s))) return rc;
      }
      // gH += gQKV[rows, 3D] @ (Wq | Wk | Wv)^T : one K = 3D contraction added to the residual-path gradient
      if (tc && (rc = dense_bwd_x_tc(p, w.gQKV, 3 * D, l * VS_PER_LAYER + VS_QKV, 3 * D, w.gH, rows, 1, s))) return rc;
    }
    if (kf) {  // Dense_0: the caller forms sum feat feat^T itself; here the output-gradient factor
      if ((rc = gram(p, w.gH, D, D, kf + K(p->kf_dense0).gtg_offset, rows, s, true))) return rc;
      continue;
    }
    RUN(PC_OTHER, features_dense0_bwd(xc, w.gH, grad + p->off_W0, Bc, nd, s));
  }
  if (kf) {  // naive-diagonal blocks: the batch-summed gradient of Re log psi (accumulated in the scratch vector)
    const int idx[2] = {p->kf_eepar, p->kf_eeanti};
    const int64_t src[2] = {p->ee_par, p->ee_anti};
    for (int t = 0; t < 2; ++t)
      if (idx[t] >= 0)
        DH_CHECK(cudaMemcpyAsync(kf + K(idx[t]).diag_offset, grad + src[t], sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  if (p->sparse) {  // (factor pass: only the lll_weight part of the result is used, below)
    for (int t = 0; t < 2 * p->nsb; ++t) {
      ProfScope ps(p, PC_OTHER, 0, s, 2);
      if ((rc = sparse_fold_bwd(orbGW(p, grad, t), P + p->orb_k[t], P + p->orb_b[t], P + p->lll_k, (t & 1) == 0 ? 1 : 0,
                                grad + p->orb_k[t], grad + p->orb_b[t], grad + p->lll_k, grad + p->lll_b, D, p->L, N * p->K, s)))
        return rc;
    }
  }
  if (kf && p->sparse) {
    DH_CHECK(cudaMemcpyAsync(kf + K(p->kf_lllk).diag_offset, grad + p->lll_k, (size_t)8 * p->L * sizeof(float), cudaMemcpyDeviceToDevice, s));
    DH_CHECK(cudaMemcpyAsync(kf + K(p->kf_lllb).diag_offset, grad + p->lll_b, (size_t)p->L * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
#undef RUN
  if (B <= chunk) p->vjp_fwd = dh_plan::FwdKey{P, x, B, ws, true};  // one chunk: its activations stay in the workspace
  return 0;
}

extern "C" int dh_logpsi_vjp(dh_plan* p, const float* P, const float* x, int64_t B, const float* cot,
                             float* grad, float* out_logpsi, void* ws, size_t ws_bytes, void* stream) {
  if (!p) return DH_E_BADARG;
  if (p->laughlin) {  // no parameters: nothing to differentiate; log psi on request
    if (out_logpsi && B > 0) return dh_logpsi(p, P, x, B, out_logpsi, ws, ws_bytes, stream);
    return 0;
  }
  if (!P || !grad || B < 0 || (B > 0 && (!x || !cot))) return DH_E_BADARG;
  return vjp_core(p, P, x, B, cot, grad, nullptr, out_logpsi, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dh_kfac_layout(const dh_plan* p, dh_kfac_entry* entries, int32_t* n, int64_t* factor_floats) {
  if (!p || !n) return DH_E_BADARG;
  if (entries) {
    if (*n < (int32_t)p->kfac.size()) return DH_E_BADARG;
    memcpy(entries, p->kfac.data(), p->kfac.size() * sizeof(dh_kfac_entry));
  }
  *n = (int32_t)p->kfac.size();
  if (factor_floats) *factor_floats = p->kfac_floats;
  return 0;
}

extern "C" int dh_kfac_factors(dh_plan* p, const float* P, const float* x, int64_t B, float* factors, void* ws,
                               size_t ws_bytes, void* stream) {
  if (!p) return DH_E_BADARG;
  if (p->laughlin || (p->sparse && 8 * p->N * p->K > 3 * p->D)) return DH_E_UNSUPPORTED;
  if (!P || !factors || B < 0 || (B > 0 && !x)) return DH_E_BADARG;
  return vjp_core(p, P, x, B, nullptr, nullptr, factors, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dh_kfac_factors_reuse_forward(dh_plan* p, const float* P, const float* x, int64_t B, float* factors, void* ws,
                                             size_t ws_bytes, void* stream) {
  if (!p) return DH_E_BADARG;
  if (p->laughlin || (p->sparse && 8 * p->N * p->K > 3 * p->D)) return DH_E_UNSUPPORTED;
  if (!P || !factors || B < 0 || (B > 0 && !x)) return DH_E_BADARG;
  return vjp_core(p, P, x, B, nullptr, nullptr, factors, nullptr, ws, ws_bytes, (cudaStream_t)stream, true);
}

// --------------------------------------------------------------------------------- KFAC update (optimizers/kfac.py:202-219)
// Factors up to KFU_SMALL rows form the "small" batch (inverted by 4-CTA clusters), the others the "large" one.
static constexpr int KFU_SMALL = 288;

static int kfac_update_tables(dh_plan* p) {
  auto& u = p->kfu;
  if (u.ready) return u.ready > 0 ? 0 : DH_E_UNSUPPORTED;
  if (p->laughlin || p->kfac.empty()) { u.ready = -1; return DH_E_UNSUPPORTED; }
  for (const auto& e : p->kfac)
    if (e.kind == 0 && (e.in_dim + e.has_bias > 1024 || e.out_dim > 1024)) { u.ready = -1; return DH_E_UNSUPPORTED; }
  for (const auto& e : p->kfac) {
    if (e.kind != 0) { u.diag.push_back(KfDiagDesc{e.kernel_offset, e.diag_offset, e.size}); continue; }
    const int blk = (int)u.blk.size();
    KfBlkDesc b{e.kernel_offset, e.has_bias ? e.bias_offset : -1, e.xtx_offset, e.gtg_offset, u.gather_floats,
                e.in_dim, e.out_dim, e.has_bias, e.rows_per_walker};
    u.gather_floats += (int64_t)(e.in_dim + e.has_bias) * e.out_dim;
    u.gather_floats = (u.gather_floats + 3) & ~(int64_t)3;  // 16-byte aligned operands for the contractions
    u.blk.push_back(b);
    KfMatDesc a{e.xtx_offset, e.has_bias ? e.xsum_offset : -1, 0, e.in_dim, e.in_dim + e.has_bias, 0, 0, blk, 0};
    KfMatDesc g{e.gtg_offset, -1, 0, e.out_dim, e.out_dim, 0, 0, blk, 1};
    u.mat.push_back(a);  // mat[2 blk] = A, mat[2 blk + 1] = G
    u.mat.push_back(g);
  }
  for (auto& m : u.mat) {
    m.cls = m.n > KFU_SMALL ? 1 : 0;
    int& dim = m.cls ? u.dim_large : u.dim_small;
    dim = m.n > dim ? m.n : dim;
  }
  u.dim_small = (u.dim_small + 3) & ~3;  // row strides the 16-byte loads of the contractions can use
  u.dim_large = (u.dim_large + 3) & ~3;
  for (auto& m : u.mat) {
    int& cnt = m.cls ? u.n_large : u.n_small;
    m.dim = m.cls ? u.dim_large : u.dim_small;
    m.dst = (long long)cnt * m.dim * m.dim;
    ++cnt;
  }
  auto up = [](const void* src, size_t bytes, void** dst) -> bool {
    if (bytes == 0) { *dst = nullptr; return true; }
    if (cudaMalloc(dst, bytes) != cudaSuccess) return false;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
  };
  if (!up(u.blk.data(), u.blk.size() * sizeof(KfBlkDesc), (void**)&u.d_blk) ||
      !up(u.mat.data(), u.mat.size() * sizeof(KfMatDesc), (void**)&u.d_mat) ||
      !up(u.diag.data(), u.diag.size() * sizeof(KfDiagDesc), (void**)&u.d_diag)) {
    u.ready = -1;
    return (int)cudaGetLastError();
  }
  u.ready = 1;
  return 0;
}

extern "C" int dh_kfac_update_shape(dh_plan* p, int32_t* n_small, int32_t* dim_small, int32_t* n_large, int32_t* dim_large,
                                    int32_t* n_blocks, int64_t* gather_floats) {
  if (!p) return DH_E_BADARG;
  int rc = kfac_update_tables(p);
  if (rc) return rc;
  if (n_small) *n_small = p->kfu.n_small;
  if (dim_small) *dim_small = p->kfu.dim_small;
  if (n_large) *n_large = p->kfu.n_large;
  if (dim_large) *dim_large = p->kfu.dim_large;
  if (n_blocks) *n_blocks = (int32_t)p->kfu.blk.size();
  if (gather_floats) *gather_floats = p->kfu.gather_floats;
  return 0;
}

extern "C" int dh_kfac_damped_factors(dh_plan* p, const float* stats, const float* dense0_xtx, float weight, float damping,
                                      float* coef, float* mats_small, float* mats_large, void* stream) {
  if (!p || !stats || !dense0_xtx || !coef || !(weight > 0.f)) return DH_E_BADARG;
  int rc = kfac_update_tables(p);
  if (rc) return rc;
  const auto& u = p->kfu;
  if ((u.n_small && !mats_small) || (u.n_large && !mats_large)) return DH_E_BADARG;
  p->launches += 2;
  return kfac_damped_factors(u.d_blk, (int)u.blk.size(), u.d_mat, (int)u.mat.size(), stats, dense0_xtx, weight, damping, coef,
                             mats_small, mats_large, (cudaStream_t)stream);
}

extern "C" int dh_kfac_update(dh_plan* p, const float* inv_small, const float* inv_large, const float* coef, const float* stats,
                              float weight, float damping, const float* grads, float* out, void* ws, size_t ws_bytes,
                              void* stream) {
  if (!p || !coef || !stats || !grads || !out || !(weight > 0.f)) return DH_E_BADARG;
  int rc = kfac_update_tables(p);
  if (rc) return rc;
  const auto& u = p->kfu;
  if ((u.n_small && !inv_small) || (u.n_large && !inv_large)) return DH_E_BADARG;
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 15) || ws_bytes < (size_t)(3 * u.gather_floats) * sizeof(float)) return DH_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float* V = static_cast<float*>(ws);
  float* T = V + u.gather_floats;
  float* U = T + u.gather_floats;
  DH_CHECK(cudaMemsetAsync(out, 0, (size_t)p->nparams * sizeof(float), s));
  if ((rc = kfac_gather(u.d_blk, (int)u.blk.size(), grads, V, s))) return rc;
  // T = A~^-1 V~ ; U = T G~^-1 for all blocks: two grouped launches of the fp32 FMA contraction
  int max_rows = 0, max_cols = 0;
  for (const KfBlkDesc& bd : u.blk) {
    max_rows = bd.din + bd.hb > max_rows ? bd.din + bd.hb : max_rows;
    max_cols = bd.dout > max_cols ? bd.dout : max_cols;
  }
  if ((rc = kfac_grouped_gemm(u.d_blk, u.d_mat, (int)u.blk.size(), max_rows, max_cols, inv_small, inv_large, V, T, 0, s))) return rc;
  if ((rc = kfac_grouped_gemm(u.d_blk, u.d_mat, (int)u.blk.size(), max_rows, max_cols, inv_small, inv_large, T, U, 1, s))) return rc;
  p->launches += 5;
  return kfac_scatter(u.d_blk, (int)u.blk.size(), u.d_diag, (int)u.diag.size(), U, coef, stats, weight, damping, grads, out, s);
}
