// C ABI (include/deephall_b200.h): plan, parameter layout, workspace carving and the launch
// sequences of the four hot-path operations.
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/deephall_b200.h"
#include "plan.h"

using namespace dh;

static void add_entry(dh_plan* p, const std::string& name, std::vector<int> shape, int64_t* off_out) {
  dh_param_entry e;
  memset(&e, 0, sizeof(e));
  snprintf(e.name, sizeof(e.name), "%s", name.c_str());
  e.offset = p->nparams;
  e.ndim = (int)shape.size();
  int64_t n = 1;
  for (size_t i = 0; i < shape.size(); ++i) { e.shape[i] = shape[i]; n *= shape[i]; }
  p->entries.push_back(e);
  if (off_out) *off_out = p->nparams;
  p->nparams += n;
}

static double binom(int n, int k) {
  double r = 1.0;
  for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
  return r;
}

extern "C" const char* dh_version(void) { return "deephall_b200 0.1 (sm_100a)"; }

// side stream + fork / join events of the chunk interleave; made at plan creation so that the ops never allocate
static bool ensure_side_stream(dh_plan* p) {
  if (p->side_stream) return true;
  if (cudaStreamCreateWithFlags(&p->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    if (p->side_stream) cudaStreamDestroy(p->side_stream);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    p->side_stream = nullptr;
    p->ev_fork = p->ev_join = nullptr;
    return false;
  }
  return true;
}

extern "C" int dh_plan_create(const dh_config* cfg, dh_plan** out) {
  if (!cfg || !out) return DH_E_BADARG;
  if (cfg->n_up < 1 || cfg->flux < 0 || cfg->ndets < 1 || cfg->num_heads < 1 || cfg->heads_dim < 1 ||
      cfg->num_layers < 0)
    return DH_E_BADARG;
  if (cfg->n_dn < 0 || (cfg->network_type == 1 && cfg->n_dn != 0)) return DH_E_BADARG;
  dh_plan* p = new dh_plan();
  if (cudaMalloc(&p->d_status, sizeof(unsigned)) != cudaSuccess || cudaMemset(p->d_status, 0, sizeof(unsigned)) != cudaSuccess ||
      cudaMalloc(&p->d_mcmc, sizeof(McmcDev)) != cudaSuccess) {
    delete p;
    return (int)cudaGetLastError();
  }
  p->cfg = *cfg;
  p->N = cfg->n_up + cfg->n_dn;
  p->twoQ = cfg->flux;
  p->L = cfg->flux + 1;
  p->K = cfg->ndets;
  p->H = cfg->num_heads;
  p->hd = cfg->heads_dim;
  p->D = p->H * p->hd;
  p->nl = cfg->num_layers;
  p->Q = 0.5f * cfg->flux;
  p->radius = cfg->radius > 0.f ? cfg->radius : sqrtf(p->Q);
  p->LNK = p->L * p->N * p->K;
  p->nsb = cfg->n_dn > 0 ? 2 : 1;
  p->orbN = 2 * p->nsb * p->LNK;
  p->sparse = cfg->orbital_type == 1 ? 1 : 0;
  p->lll_k = p->lll_b = -1;
  p->orb_eff = p->orb_geff = 0;
  if (cfg->orbital_type != 0 && cfg->orbital_type != 1) { delete p; return DH_E_BADARG; }
  p->nparams = 0;
  p->prep = nullptr;
  p->prep_floats = 0;
  p->prep_src = nullptr;
  p->auto_prepare = 1;
  p->prep_fwd_valid = false;
  p->prep_vjp_valid = false;
  p->prof_on = false;
  p->prof_used = 0;
  p->launches = 0;
  p->kfac_floats = 0;
  p->kf_dense0 = p->kf_eepar = p->kf_eeanti = -1;
  p->laughlin = cfg->network_type == 1 ? 1 : 0;
  p->twoQ1 = 0;
  p->lskip = -1;
  if (cfg->network_type != 0 && cfg->network_type != 1) { delete p; return DH_E_BADARG; }
  if (p->N > 16 || p->D % 32 != 0 || p->D > 256 || p->hd % 4 != 0) { delete p; return DH_E_UNSUPPORTED; }
  if (p->laughlin) {
    // networks/laughlin.py:33-47: Q1 = flux/2 - p (N - 1); N = 2 Q1 + 1 ground state, 2 Q1 quasihole, 2 Q1 + 2 quasiparticle
    const int pf = cfg->cf_flux > 0 ? cfg->cf_flux : 1;
    p->twoQ1 = cfg->flux - 2 * pf * (p->N - 1);
    p->lskip = -1;
    if (p->twoQ1 == p->N) {  // quasihole (laughlin.py:38-41,73-83): the orbital m = -lz is left out of the N + 1
      const double sk = 0.5 * p->twoQ1 - (double)cfg->excitation_lz;
      const long ski = lround(sk);
      if (fabs(sk - (double)ski) > 1e-6 || ski < 0 || ski > p->twoQ1) { delete p; return DH_E_BADARG; }
      p->lskip = (int)ski;
    } else if (p->twoQ1 == p->N - 2 && p->twoQ1 >= 0) {  // quasiparticle (laughlin.py:42-46,85-100): shell + projected orbital
      const double ea = 0.5 * p->twoQ1 + (double)cfg->excitation_lz;  // u exponent Q1 + lz of the excited orbital
      const long eai = lround(ea);
      if (fabs(ea - (double)eai) > 1e-6 || eai < -1 || eai > p->twoQ1 + 1) { delete p; return DH_E_BADARG; }
      p->lqp = 1;
      p->lqp_a = (int)eai;
    } else if (p->twoQ1 != p->N - 1) { delete p; return DH_E_UNSUPPORTED; }
    p->L = p->twoQ1 + 1;
    p->K = 1;
    p->nl = 0;
    p->LNK = p->L * p->N * p->K;
    p->orbN = 2 * p->LNK;
    p->sparse = 0;
    p->gemm_impl = 0;
    p->tc_f16 = 0;
    p->tc_merged = 0;
    p->a_planes = 0;
    p->orb_fuse = 0;
    p->ee_par = p->ee_anti = -1;
    for (int t = 0; t < 4; ++t) p->orb_k[t] = p->orb_b[t] = -1;
    p->off_W0 = -1;
    p->w0qkv = p->fold_tmp = p->cot_scale = 0;
    std::vector<double> ones(p->L, 1.0);
    cudaError_t e1 = cudaMalloc(&p->d_normfac, p->L * sizeof(double));
    if (e1 != cudaSuccess) { delete p; return (int)e1; }
    e1 = cudaMemcpy(p->d_normfac, ones.data(), p->L * sizeof(double), cudaMemcpyHostToDevice);
    if (e1 != cudaSuccess) { cudaFree(p->d_normfac); delete p; return (int)e1; }
    if (dual_stream_env()) ensure_side_stream(p);
    *out = p;
    return 0;
  }
  const int D = p->D, H = p->H, hd = p->hd, N = p->N, L = p->L, K = p->K;
  const std::string pl = "PsiformerLayers_0/";
  add_entry(p, pl + "Dense_0/kernel", {4, D}, &p->off_W0);
  p->layer.resize(p->nl);
  for (int l = 0; l < p->nl; ++l) {
    LayerOff& o = p->layer[l];
    const std::string a = pl + "MultiHeadAttention_" + std::to_string(l) + "/";
    add_entry(p, a + "query/kernel", {D, H, hd}, &o.q_k);
    add_entry(p, a + "query/bias", {H, hd}, &o.q_b);
    add_entry(p, a + "key/kernel", {D, H, hd}, &o.k_k);
    add_entry(p, a + "key/bias", {H, hd}, &o.k_b);
    add_entry(p, a + "value/kernel", {D, H, hd}, &o.v_k);
    add_entry(p, a + "value/bias", {H, hd}, &o.v_b);
    add_entry(p, a + "out/kernel", {H, hd, D}, &o.o_k);
    add_entry(p, a + "out/bias", {D}, &o.o_b);
    add_entry(p, pl + "Dense_" + std::to_string(1 + 2 * l) + "/kernel", {D, D}, &o.d1_k);
    add_entry(p, pl + "LayerNorm_" + std::to_string(2 * l) + "/scale", {D}, &o.ln0_s);
    add_entry(p, pl + "LayerNorm_" + std::to_string(2 * l) + "/bias", {D}, &o.ln0_b);
    add_entry(p, pl + "Dense_" + std::to_string(2 + 2 * l) + "/kernel", {D, D}, &o.d2_k);
    add_entry(p, pl + "Dense_" + std::to_string(2 + 2 * l) + "/bias", {D}, &o.d2_b);
    add_entry(p, pl + "LayerNorm_" + std::to_string(2 * l + 1) + "/scale", {D}, &o.ln1_s);
    add_entry(p, pl + "LayerNorm_" + std::to_string(2 * l + 1) + "/bias", {D}, &o.ln1_b);
  }
  const std::string ob = "Orbitals_0/featured_orbitals/";
  for (int t = 0; t < 4; ++t) p->orb_k[t] = p->orb_b[t] = -1;
  const int F = p->sparse ? 8 : L;  // blocks.py:47-56: sparse orbitals project to 8 features
  for (int t = 0; t < 2 * p->nsb; ++t) {  // blocks.py:29-34: one (re, im) pair of DenseGeneral per non-empty spin block
    add_entry(p, ob + "DenseGeneral_" + std::to_string(t) + "/kernel", {D, F, N, K}, &p->orb_k[t]);
    add_entry(p, ob + "DenseGeneral_" + std::to_string(t) + "/bias", {F, N, K}, &p->orb_b[t]);
  }
  if (p->sparse) {  // blocks.py:57: nn.DenseGeneral(2Q + 1, axis=1) on the complex 8-feature orbitals
    add_entry(p, "Orbitals_0/lll_weight/kernel", {8, L}, &p->lll_k);
    add_entry(p, "Orbitals_0/lll_weight/bias", {L}, &p->lll_b);
  }
  p->ee_par = p->ee_anti = -1;
  const int nu = cfg->n_up, nd = cfg->n_dn;
  if (nu * (nu - 1) / 2 + nd * (nd - 1) / 2 > 0) add_entry(p, "Jastrow_0/ee_par", {1}, &p->ee_par);  // blocks.py:91
  // blocks.py:99 tests r_ees[0][1].shape[0] > 0, and that block has shape (n_up, n_dn): the leaf exists whenever
  // n_up > 0, also for spin-polarised systems, where it multiplies an empty sum (inert, zero gradient)
  if (nu > 0) add_entry(p, "Jastrow_0/ee_anti", {1}, &p->ee_anti);

  // ---- KFAC curvature blocks, in parameter order (optimizers/kfac.py: every Dense / DenseGeneral is a repeated-dense
  // block over the electron axis; everything else gets a diagonal block)
  {
    int64_t off = 0;
    auto take = [&](int64_t n) { int64_t o = off; off += (n + 63) / 64 * 64; return o; };
    auto dense = [&](const std::string& name, int64_t k_off, int64_t b_off, int in, int out, int rpw, int64_t xtx,
                     int64_t xsum) {
      dh_kfac_entry e;
      memset(&e, 0, sizeof(e));
      snprintf(e.name, sizeof(e.name), "%s", name.c_str());
      e.kind = 0; e.in_dim = in; e.out_dim = out; e.has_bias = b_off >= 0 ? 1 : 0; e.rows_per_walker = rpw;
      e.kernel_offset = k_off; e.bias_offset = b_off;
      e.xtx_offset = xtx; e.xsum_offset = b_off >= 0 ? xsum : -1;
      e.gtg_offset = take((int64_t)out * out);
      e.diag_offset = -1; e.size = (int64_t)in * out;
      p->kfac.push_back(e);
      return (int)p->kfac.size() - 1;
    };
    auto diag = [&](const std::string& name, int64_t p_off, int n) {
      dh_kfac_entry e;
      memset(&e, 0, sizeof(e));
      snprintf(e.name, sizeof(e.name), "%s", name.c_str());
      e.kind = 1; e.in_dim = e.out_dim = 0; e.has_bias = 0; e.rows_per_walker = 1;
      e.kernel_offset = p_off; e.bias_offset = -1;
      e.xtx_offset = e.xsum_offset = e.gtg_offset = -1;
      e.diag_offset = take(n); e.size = n;
      p->kfac.push_back(e);
      return (int)p->kfac.size() - 1;
    };
    p->kf_dense0 = dense(pl + "Dense_0/kernel", p->off_W0, -1, 4, D, N, -1, -1);
    p->kf_layer.resize(p->nl);
    for (int l = 0; l < p->nl; ++l) {
      const LayerOff& o = p->layer[l];
      KfLayer& k = p->kf_layer[l];
      const std::string a = pl + "MultiHeadAttention_" + std::to_string(l) + "/";
      const int64_t xh = take((int64_t)D * D), sh = take(D);  // q, k, v share their input h
      k.q = dense(a + "query/kernel", o.q_k, o.q_b, D, D, N, xh, sh);
      k.k = dense(a + "key/kernel", o.k_k, o.k_b, D, D, N, xh, sh);
      k.v = dense(a + "value/kernel", o.v_k, o.v_b, D, D, N, xh, sh);
      const int64_t xo = take((int64_t)D * D), so = take(D);
      k.o = dense(a + "out/kernel", o.o_k, o.o_b, D, D, N, xo, so);
      k.d1 = dense(pl + "Dense_" + std::to_string(1 + 2 * l) + "/kernel", o.d1_k, -1, D, D, N, take((int64_t)D * D), -1);
      k.ln0s = diag(pl + "LayerNorm_" + std::to_string(2 * l) + "/scale", o.ln0_s, D);
      k.ln0b = diag(pl + "LayerNorm_" + std::to_string(2 * l) + "/bias", o.ln0_b, D);
      const int64_t x2 = take((int64_t)D * D), s2 = take(D);
      k.d2 = dense(pl + "Dense_" + std::to_string(2 + 2 * l) + "/kernel", o.d2_k, o.d2_b, D, D, N, x2, s2);
      k.ln1s = diag(pl + "LayerNorm_" + std::to_string(2 * l + 1) + "/scale", o.ln1_s, D);
      k.ln1b = diag(pl + "LayerNorm_" + std::to_string(2 * l + 1) + "/bias", o.ln1_b, D);
    }
    for (int t = 0; t < 4; ++t) p->kf_orb[t] = -1;
    for (int sbk = 0; sbk < p->nsb; ++sbk) {  // the (re, im) projections of a spin block share its electrons' h
      const int64_t xb = take((int64_t)D * D), sbs = take(D);
      const int n_alpha = p->nsb == 1 ? N : (sbk == 0 ? cfg->n_up : cfg->n_dn);
      for (int part = 0; part < 2; ++part) {
        const int t = 2 * sbk + part;
        // (sparse orbitals, blocks.py:52-56: the projection's own output is the 8-feature tensor [8][N][K])
        p->kf_orb[t] = dense(ob + "DenseGeneral_" + std::to_string(t) + "/kernel", p->orb_k[t], p->orb_b[t], D,
                             p->sparse ? 8 * N * K : p->LNK, n_alpha, xb, sbs);
      }
    }
    // sparse orbitals: `lll_weight` is a DenseGeneral over axis 1 of a complex tensor (blocks.py:57): its dot_general
    // has the dimension numbers of none of optimizers/kfac.py:148-195's patterns, so kernel and bias fall to kfac_jax's
    // generic tag like the Jastrow parameters -- naive diagonal, the factor vector carries the batch-summed gradient
    p->kf_lllk = p->kf_lllb = -1;
    if (p->sparse) {
      p->kf_lllk = diag("Orbitals_0/lll_weight/kernel", p->lll_k, 8 * L);
      p->kf_lllb = diag("Orbitals_0/lll_weight/bias", p->lll_b, L);
      p->kfac[p->kf_lllk].kind = 2;
      p->kfac[p->kf_lllb].kind = 2;
    }
    // no layer pattern matches the Jastrow parameters: kfac_jax's generic tag, whose diagonal is "naive" -- the square of
    // the batch-summed gradient; the factor vector carries that sum (kind 2)
    p->kf_eepar = p->ee_par >= 0 ? diag("Jastrow_0/ee_par", p->ee_par, 1) : -1;
    p->kf_eeanti = p->ee_anti >= 0 ? diag("Jastrow_0/ee_anti", p->ee_anti, 1) : -1;
    if (p->kf_eepar >= 0) p->kfac[p->kf_eepar].kind = 2;
    if (p->kf_eeanti >= 0) p->kfac[p->kf_eeanti].kind = 2;
    p->kfac_floats = off;
  }

  // prepared-weight slots for the tcgen05 path: [Npad][D] hi | lo, plus fused biases
  {
    // dh_config.contraction: 0 fp16 pieces (one accumulator per tile, double-buffered), 1 TF32 pieces (main + correction
    // accumulators), 2 plain fp32 FMA
    const int mode = cfg->contraction;
    if (mode < 0 || mode > 2) { delete p; return DH_E_BADARG; }
    p->gemm_impl = mode == 2 ? 0 : 1;
    p->tc_f16 = (gemm_tc_f16_ok(D) && mode == 0) ? 1 : 0;
    p->tc_merged = p->tc_f16;
    p->a_planes = 0;
    // envelope contraction as the epilogue of the orbital projection (gemm_tc.cu, ORB): 32 jet rows per electron (N = 12),
    // one determinant, one spin block, full orbitals, at most 48 orbitals
    p->ln_value_fuse = !(dbg_env("DH_LNV_FUSE") && atoi(dbg_env("DH_LNV_FUSE")) == 0) ? 1 : 0;
    p->orb_fuse = (p->gemm_impl == 1 && p->tc_f16 && D == 256 && N == 12 && K == 1 && p->nsb == 1 && !p->sparse && L <= 48 &&
                   !(dbg_env("DH_ORB_FUSE") && atoi(dbg_env("DH_ORB_FUSE")) == 0)) ? 1 : 0;
    size_t off = 0;
    auto slot = [&](int Nout, bool has_bias) {
      dh_plan::Slot sl;
      const size_t npad = (size_t)((Nout + 15) & ~15);
      sl.Nout = Nout;
      sl.hi = off; off += al(npad * D);
      sl.lo = off; off += al(npad * D);
      sl.bias = SIZE_MAX;
      if (has_bias) { sl.bias = off; off += al((size_t)Nout); }
      sl.scale = off; off += al(3);
      sl.ldw = D;
      p->slots.push_back(sl);
    };
    for (int l = 0; l < p->nl; ++l) { slot(3 * D, true); slot(D, true); slot(D, false); slot(D, true); slot(D, true); }
    slot(p->orbN, true);
    if (p->orb_fuse) {  // the same projection with its output columns permuted to [tile][m][re | im][column], orb_per_tile(L) m per tile
      const int nperm = orb_columns(L);
      slot(nperm, true);
      p->orb_perm = off; off += al((size_t)(D + 1) * nperm);  // fp32 staging of the permuted kernel [D][nperm] and bias [nperm]
    }
    // reverse-pass planes: [D rows][Kpad], Kpad = K rounded up to 32
    auto vslot = [&](int K) {
      dh_plan::Slot sl;
      sl.Nout = D;
      sl.ldw = (K + 31) & ~31;
      sl.hi = off; off += al((size_t)D * sl.ldw);
      sl.lo = off; off += al((size_t)D * sl.ldw);
      sl.bias = SIZE_MAX;
      sl.scale = off; off += al(3);
      p->vslots.push_back(sl);
    };
    for (int l = 0; l < p->nl; ++l) { vslot(D); vslot(D); vslot(D); vslot(3 * D); }
    vslot(p->orbN);
    p->cot_scale = off; off += al(2);
    p->w0qkv = off; off += al((size_t)4 * 3 * D);
    p->fold_tmp = off; off += al((size_t)D * D);
    if (p->sparse) {
      p->orb_eff = off; off += al((size_t)2 * p->nsb * (D + 1) * p->LNK);
      p->orb_geff = off; off += al((size_t)2 * p->nsb * (D + 1) * p->LNK);
    }
    p->prep_floats = off;
    cudaError_t e0 = cudaMalloc(&p->prep, off * sizeof(float));
    if (e0 == cudaSuccess) e0 = cudaMalloc(&p->raw_params, (size_t)p->nparams * sizeof(float));
    if (e0 != cudaSuccess) { delete p; return (int)e0; }
  }

  // sqrt(C(2Q, Q-m)) for m = -Q..Q  (blocks.py:45-46); index a = Q+m -> C(2Q, 2Q-a) = C(2Q, a)
  std::vector<double> nf(L);
  for (int a = 0; a < L; ++a) nf[a] = sqrt(binom(p->twoQ, a));
  cudaError_t e = cudaMalloc(&p->d_normfac, L * sizeof(double));
  if (e != cudaSuccess) { delete p; return (int)e; }
  e = cudaMemcpy(p->d_normfac, nf.data(), L * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(p->d_normfac); delete p; return (int)e; }
  if (dual_stream_env()) ensure_side_stream(p);
  *out = p;
  return 0;
}

extern "C" int dh_plan_status(dh_plan* p, int32_t clear, uint32_t* out_bits, void* stream) {
  if (!p || !out_bits) return DH_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  *out_bits = 0;
  if (!p->d_status) return 0;
  DH_CHECK(cudaMemcpyAsync(out_bits, p->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  if (clear) DH_CHECK(cudaMemsetAsync(p->d_status, 0, sizeof(uint32_t), s));
  DH_CHECK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int dh_plan_status_copy(dh_plan* p, uint32_t* dst_device, void* stream) {
  if (!p || !dst_device) return DH_E_BADARG;
  if (!p->d_status) return (int)cudaMemsetAsync(dst_device, 0, sizeof(uint32_t), (cudaStream_t)stream);
  return (int)cudaMemcpyAsync(dst_device, p->d_status, sizeof(uint32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}

extern "C" int dh_plan_destroy(dh_plan* p) {
  if (!p) return DH_E_BADARG;
  if (p->d_status) {
    if (range_flag_get() == p->d_status) range_flag_set(nullptr);  // never leave a dangling status pointer behind
    cudaFree(p->d_status);
  }
  for (auto& g : p->move_graphs) cudaGraphExecDestroy(g.exec);
  if (p->kfu.d_blk) cudaFree(p->kfu.d_blk);
  if (p->kfu.d_mat) cudaFree(p->kfu.d_mat);
  if (p->kfu.d_diag) cudaFree(p->kfu.d_diag);
  if (p->d_mcmc) cudaFree(p->d_mcmc);
  if (p->raw_params) cudaFree(p->raw_params);
  if (p->d_normfac) cudaFree(p->d_normfac);
  if (p->prep) cudaFree(p->prep);
  for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
  if (p->side_stream) cudaStreamDestroy(p->side_stream);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  delete p;
  return 0;
}

extern "C" int64_t dh_param_count(const dh_plan* p) { return p ? p->nparams : -1; }

extern "C" int dh_param_layout(const dh_plan* p, dh_param_entry* entries, int32_t* n) {
  if (!p || !n) return DH_E_BADARG;
  if (entries) {
    if (*n < (int32_t)p->entries.size()) return DH_E_BADARG;
    memcpy(entries, p->entries.data(), p->entries.size() * sizeof(dh_param_entry));
  }
  *n = (int32_t)p->entries.size();
  return 0;
}

size_t vjp_ws_floats(const dh_plan* p, int64_t Bc);  // api_vjp.cu

// floats of forward workspace for a pass over B walkers (all copies of the chunk interleave)
static inline size_t fwd_ws_floats(const dh_plan* p, bool jets, int64_t B) {
  int copies;
  const int64_t chunk = plan_chunks(p, jets, B, &copies);
  const size_t one = carve_fwd(p, nullptr, pick_chunk(p, jets, B), jets, false).floats;  // single-stream fallback always fits
  const size_t two = carve_fwd(p, nullptr, chunk, jets, false).floats * copies;
  return one > two ? one : two;
}

extern "C" int dh_workspace_bytes(const dh_plan* p, int op, int64_t B, size_t* bytes) {
  if (!p || !bytes || B < 0) return DH_E_BADARG;
  size_t fl = 0;
  switch (op) {
    // (passes of two or more chunks interleave them on two streams: two activation workspaces)
    case DH_OP_LOGPSI: fl = fwd_ws_floats(p, false, B); break;
    case DH_OP_LOCAL_ENERGY: fl = fwd_ws_floats(p, true, B); break;
    case DH_OP_MCMC: fl = carve_mcmc(p, nullptr, B).floats + fwd_ws_floats(p, false, B); break;
    case DH_OP_VJP: fl = vjp_ws_floats(p, pick_chunk(p, false, B)); break;
    case DH_OP_KFAC:
      fl = vjp_ws_floats(p, pick_chunk(p, false, B)) + al((size_t)pick_chunk(p, false, B) * 2) + al((size_t)p->nparams);
      break;
    default: return DH_E_BADARG;
  }
  *bytes = fl * sizeof(float) + 256;
  return 0;
}

// Runs the network body + tail for Bc walkers whose coordinates are x; fills w.ld / outputs.
// mv (value-only passes): the pass evaluates a Metropolis move of these walkers -- x is then the OUTPUT buffer of the proposal
// drawn from mv->x1 and the walkers' accept / select step runs at the end (both inside the pass's first / last kernel where
// the fused forms apply, as launches of their own otherwise).
int forward_chunk(const dh_plan* p, const float* P, const float* x, int64_t Bc, bool jets, const FwdWs& w,
                  FinalizeArgs fa, cudaStream_t s, const MoveChunk* mv) {
  const int N = p->N, D = p->D;
  if (mv && jets) return DH_E_BADARG;
  // value-only passes on the tensor-core path: proposal, features, both first-layer maps and the envelope table are ONE launch
  const bool vprologue = !jets && !p->laughlin && p->gemm_impl == 1 && p->nl > 0;
  const bool vtail = !jets && N <= 16 && w.Minv == nullptr;  // log-determinants (+ accept) inside the finalize launch
  if (mv && !vprologue) {
    ProfScope ps(p, PC_MCMC, 0, s);
    int rcp = mcmc_propose_dev(mv->x1, const_cast<float*>(x), Bc, N, mv->dv, mv->walker0, s);
    if (rcp) return rcp;
  }
  auto move_tail = [&](FinalizeArgs& f) {
    if (mv && vtail) { f.mv_x1 = mv->x1; f.mv_lp1 = mv->lp1; f.mv_dv = mv->dv; f.mv_walker0 = mv->walker0; }
  };
  auto move_accept = [&]() -> int {  // separate accept launch when the fused tail does not apply
    if (!mv || vtail) return 0;
    ProfScope ps(p, PC_MCMC, 0, s);
    return mcmc_accept_dev(mv->x1, x, mv->lp1, fa.out_logpsi, 2, Bc, N, mv->dv, mv->walker0, s);
  };
  const int R = jets ? 2 * N + 8 : 1;
  const int64_t rows = Bc * N * R;
  NetDims nd{N, R, D, p->H, p->hd, p->cfg.n_up};
  TailDims td{N, R, p->L, p->K, p->twoQ, p->cfg.n_up, p->cfg.n_dn};
  int rc;
  if (p->laughlin) {
    // analytic Laughlin ground state: orbital-matrix jets straight from the coordinates, then the same tail
    TailDims tl{N, R, p->L, 1, p->twoQ1, p->cfg.n_up, 0};
    tl.lskip = p->lskip;
    tl.qp = p->lqp;
    tl.qp_a = p->lqp_a;
    ProfScope pst(p, PC_TAIL, 0, s, vtail ? 2 : 3);
    if ((rc = laughlin_orbital_jets(x, p->d_normfac, w.Mj, Bc, tl, s))) return rc;
    if (vtail) fa.Mj_value = w.Mj;
    else if ((rc = logdet_jets_impl(w.Mj, w.ld, w.Minv, Bc, tl, s))) return rc;
    move_tail(fa);
    fa.ld = w.ld;
    fa.x = x;
    fa.ee_par = nullptr;
    fa.ee_anti = nullptr;
    fa.Q = p->Q;
    fa.radius = p->radius;
    fa.interaction_strength = p->cfg.interaction_strength;
    fa.interaction_type = p->cfg.interaction_type;
    fa.lpjet = jets ? w.lpjet : nullptr;
    if ((rc = finalize(fa, Bc, tl, s))) return rc;
    return move_accept();
  }
  // jets on the tensor-core path: Dense_0's output has 10 non-zero jet rows per electron; it is written in that
  // compressed form (into t1, which is free until the second Dense of the layer) and expanded by the first LayerNorm
  const bool h0_comp = jets && p->gemm_impl == 1 && p->nl > 0;
  const bool orb_epi = p->orb_fuse && (!jets || rows % 128 == 0);  // envelope contraction as the orbital projection's epilogue
  const bool jprologue = jets && h0_comp && orb_epi && (D & 3) == 0;
  if (vprologue) {
    ProfScope ps(p, PC_OTHER, 0, s);
    const dh_plan::Slot& q = p->slots[SL_QKV];
    ValuePrologue vp;
    memset(&vp, 0, sizeof(vp));
    vp.x = mv ? mv->x1 : x;
    vp.x_new = mv ? const_cast<float*>(x) : nullptr;
    vp.dv = mv ? mv->dv : nullptr;
    vp.walker0 = mv ? mv->walker0 : 0;
    vp.W0 = P + p->off_W0; vp.h = w.h; vp.n0 = D;
    vp.W1 = p->prep + p->w0qkv; vp.b1 = p->prep + q.bias; vp.q = w.qkv; vp.n1 = 3 * D;
    if (orb_epi) {
      const dh_plan::Slot& sl = p->slots[p->nl * SL_PER_LAYER + 1];
      vp.normfac = p->d_normfac; vp.bre = orbB(p, P, 0); vp.bim = orbB(p, P, 1); vp.unscale = p->prep + sl.scale + 1;
      vp.tab = w.cbuf; vp.L = p->L; vp.NK = N * p->K; vp.twoQ = p->twoQ;
    }
    if ((rc = value_prologue(vp, Bc * N, N, p->cfg.n_up, s))) return rc;
  } else if (jprologue) {
    // jet passes: compressed Dense_0 output, compressed first-layer q|k|v and the envelope table in one launch
    ProfScope ps(p, PC_OTHER, 0, s);
    const dh_plan::Slot& q = p->slots[SL_QKV];
    const dh_plan::Slot& sl = p->slots[p->nl * SL_PER_LAYER + 1];
    if ((rc = jets_prologue(x, P + p->off_W0, w.t1, D, p->prep + p->w0qkv, p->prep + q.bias, w.qkv, 3 * D, p->d_normfac, orbB(p, P, 0),
                            orbB(p, P, 1), p->prep + sl.scale + 1, w.cbuf, Bc, p->cfg.n_up, td, s)))
      return rc;
  } else {
    ProfScope ps(p, PC_OTHER, 0, s);
    if ((rc = features_linear(x, P + p->off_W0, nullptr, h0_comp ? w.t1 : w.h, D, Bc, nd, h0_comp ? 1 : 0, s))) return rc;
  }
  // jet passes with fp16 pieces: every tensor that is the left operand of a contraction (att, h) lives as fp16
  // hi / lo planes in its buffer -- written so by the attention and LayerNorm kernels, read back as hi + lo by the
  // next LayerNorm's residual -- and the contraction takes its operand tiles straight from them by TMA
  const bool pl = jets && p->a_planes;
  for (int l = 0; l < p->nl; ++l) {
    const LayerOff& o = p->layer[l];
    if (l == 0 && (vprologue || jprologue)) {
      // q|k|v of the first layer came with the prologue
    } else if (l == 0 && p->gemm_impl == 1) {
      // h = feat @ W0 is linear in the features: q|k|v = feat @ (W0 Wqkv) + b, a 4-deep contraction
      ProfScope ps(p, PC_OTHER, 0, s);
      const dh_plan::Slot& q = p->slots[SL_QKV];
      // (jets: only the 10 non-zero rows per electron are written; attention_jets' first-layer form reads them)
      if ((rc = features_linear(x, p->prep + p->w0qkv, p->prep + q.bias, w.qkv, 3 * D, Bc, nd, jets ? 1 : 0, s))) return rc;
    } else if ((rc = dense_qkv(p, P, l, w.h, w.qkv, rows, R, s, pl))) return rc;
    { ProfScope ps(p, PC_ATTENTION, 0, s);
      if (jets) rc = attention_jets(w.qkv, w.att, Bc, nd, (l == 0 && p->gemm_impl == 1) ? 1 : 0, p->tc_f16 ? 0 : 1, s);
      else rc = attention_value(w.qkv, w.att, Bc, nd, s);
      if (rc) return rc; }
    // value-only passes on the fp16-piece tensor-core path: the residual + LayerNorm (+ tanh) that follows a 256-wide
    // contraction runs as its epilogue (gemm_tc.cu, LNV) -- no separate LayerNorm launch, the contraction's output never
    // goes to HBM.
    const bool lnv = !jets && p->gemm_impl == 1 && p->tc_f16 && D == 256 && p->ln_value_fuse;
    if (lnv) {
      const LnArgs la0{P + o.ln0_s, P + o.ln0_b, 0}, la1{P + o.ln1_s, P + o.ln1_b, 1};
      if ((rc = dense_tc(p, w.att, l * SL_PER_LAYER + SL_OD, w.h, rows, D, R, s, false, &la0))) return rc;  // h = LN0(h + att Wod + bod)
      if ((rc = dense_tc(p, w.h, l * SL_PER_LAYER + SL_D2, w.h, rows, D, R, s, false, &la1))) return rc;    // h = LN1(h + tanh(h W2 + b2))
      continue;
    }
    if (p->gemm_impl == 1) {
      // MHA out-projection and the bias-free Dense that follows it are one linear map (Wo W1, bo W1)
      if ((rc = dense_tc(p, w.att, l * SL_PER_LAYER + SL_OD, w.t2, rows, D, R, s, pl))) return rc;
    } else {
      if ((rc = dense_layer(p, P, l, SL_O, w.att, w.t1, rows, R, s))) return rc;
      if ((rc = dense_layer(p, P, l, SL_D1, w.t1, w.t2, rows, R, s))) return rc;
    }
    { ProfScope ps(p, PC_LAYERNORM, 0, s);
      if (l == 0 && h0_comp) rc = residual_layernorm_ex(w.t1, w.t2, P + o.ln0_s, P + o.ln0_b, w.h, Bc, nd, 0, 1, 0, pl ? 1 : 0, s);
      else rc = residual_layernorm_ex(w.h, w.t2, P + o.ln0_s, P + o.ln0_b, w.h, Bc, nd, 0, 0, pl ? 1 : 0, pl ? 1 : 0, s);
      if (rc) return rc; }
    if ((rc = dense_layer(p, P, l, SL_D2, w.h, w.t1, rows, R, s, pl))) return rc;
    { ProfScope ps(p, PC_LAYERNORM, 0, s);
      if ((rc = residual_layernorm_ex(w.h, w.t1, P + o.ln1_s, P + o.ln1_b, w.h, Bc, nd, 1, 0, pl ? 1 : 0, pl ? 1 : 0, s))) return rc; }
  }
  if (orb_epi) {
    // The envelope contraction (blocks.py:59-70) is the EPILOGUE of the orbital projection (blocks.py:28-35): the per-electron
    // envelope values (jet passes: jets) go to a small table first (w.cbuf, which the coefficient tensor no longer needs), the
    // contraction reads its coefficients out of tensor memory and writes the orbital matrices -- c[rows][2 L N] never exists in HBM
    const dh_plan::Slot& sl = p->slots[p->nl * SL_PER_LAYER + 1];
    if (!jprologue && (jets || !vprologue)) {  // (the table normally comes with the prologue)
      ProfScope pse(p, PC_TAIL, 0, s);
      if (jets) rc = envelope_table(x, p->d_normfac, orbB(p, P, 0), orbB(p, P, 1), p->prep + sl.scale + 1, w.cbuf, Bc, td, s);
      else rc = envelope_value_table(x, p->d_normfac, orbB(p, P, 0), orbB(p, P, 1), p->prep + sl.scale + 1, w.cbuf, Bc, td, s);
      if (rc) return rc;
    }
    ProfScope ps(p, PC_GEMM, 2.0 * (double)rows * p->orbN * p->D, s);
    TcGemm g;
    g.A = w.h; g.lda = p->D; g.Wt_hi = p->prep + sl.hi; g.Wt_lo = p->prep + sl.lo; g.ldw = p->D;
    g.bias = nullptr; g.inv_scale = nullptr;  // both are folded into the envelope table
    g.C = w.Mj; g.ldc = sl.Nout; g.M = rows; g.N = sl.Nout; g.K = p->D; g.rpg = R;
    g.f16 = 1; g.merged = 1; g.reduce_add = 0; g.a_scale = nullptr; g.A_lo = nullptr;
    g.orb_env = w.cbuf; g.orb_Mj = w.Mj; g.orb_L = p->L; g.ln_res = nullptr; g.ln_gamma = nullptr; g.ln_beta = nullptr; g.ln_tanh = 0;
    if ((rc = gemm_tc_ex(g, s))) return rc;
  } else {
    if ((rc = dense_orb(p, P, w.h, w.cbuf, rows, R, s, pl))) return rc;
    ProfScope psc(p, PC_TAIL, 0, s);
    if ((rc = orbital_contract(w.cbuf, x, p->d_normfac, w.Mj, Bc, td, s))) return rc;
  }
  ProfScope pst(p, PC_TAIL, 0, s, vtail ? 1 : 2);
  if (vtail) fa.Mj_value = w.Mj;
  else if ((rc = logdet_jets_impl(w.Mj, w.ld, w.Minv, Bc, td, s))) return rc;
  move_tail(fa);
  fa.ld = w.ld;
  fa.x = x;
  fa.ee_par = p->ee_par >= 0 ? P + p->ee_par : nullptr;
  fa.ee_anti = p->ee_anti >= 0 ? P + p->ee_anti : nullptr;
  fa.Q = p->Q;
  fa.radius = p->radius;
  fa.interaction_strength = p->cfg.interaction_strength;
  fa.interaction_type = p->cfg.interaction_type;
  fa.lpjet = jets ? w.lpjet : nullptr;
  if ((rc = finalize(fa, Bc, td, s))) return rc;
  return move_accept();
}

static int prepare_weights_now(dh_plan* p, const float* P, cudaStream_t s) {
  p->vjp_fwd.valid = false;  // the parameters may have changed
  const int D = p->D, LNK = p->LNK, f16 = p->tc_f16;
  int rc;
  if (p->sparse) {  // effective full projections of the sparse orbitals (both contraction implementations read them)
    for (int t = 0; t < 2 * p->nsb; ++t)
      if ((rc = sparse_fold(P + p->orb_k[t], P + p->orb_b[t], P + p->lll_k, P + p->lll_b, (t & 1) == 0 ? 1 : 0,
                            const_cast<float*>(orbW(p, P, t)), D, p->L, p->N * p->K, s)))
        return rc;
    p->launches += 2 * p->nsb;
  }
  if (p->gemm_impl != 1) return 0;
  auto cp = [&](size_t dst, int64_t src, size_t n) {
    return (int)cudaMemcpyAsync(p->prep + dst, P + src, n * sizeof(float), cudaMemcpyDeviceToDevice, s);
  };
  // element `e` of a slot's hi / lo plane (fp16 planes pack two elements per float of `prep`)
  auto plane = [&](size_t off_floats, size_t e) -> void* {
    return f16 ? (void*)(reinterpret_cast<__half*>(p->prep + off_floats) + e) : (void*)(p->prep + off_floats + e);
  };
  // one slot = `nparts` blocks W_t[D][n_t] stacked along the output dimension
  struct Part { const float* W; int64_t ldw; int n; };
  auto fill = [&](const dh_plan::Slot& sl, const Part* parts, int nparts, bool zero_first) -> int {
    float* slot = p->prep + sl.scale;
    if (zero_first) {
      const size_t bytes = (size_t)((sl.Nout + 15) & ~15) * D * (f16 ? 2 : 4);
      DH_CHECK(cudaMemsetAsync(plane(sl.hi, 0), 0, bytes, s));
      DH_CHECK(cudaMemsetAsync(plane(sl.lo, 0), 0, bytes, s));
    }
    if (f16) {
      DH_CHECK(cudaMemsetAsync(slot, 0, 3 * sizeof(float), s));
      for (int t = 0; t < nparts; ++t)
        if ((rc = weight_maxabs_tc(parts[t].W, parts[t].ldw, D, parts[t].n, slot, s))) return rc;
    }
    size_t row = 0;
    for (int t = 0; t < nparts; ++t) {
      const bool last = t == nparts - 1;
      if ((rc = split_weight_tc(parts[t].W, parts[t].ldw, D, parts[t].n, last && !zero_first ? 1 : 0,
                                plane(sl.hi, row * D), plane(sl.lo, row * D), slot, f16, s)))
        return rc;
      row += parts[t].n;
    }
    return 0;
  };
  for (int l = 0; l < p->nl; ++l) {
    const LayerOff& o = p->layer[l];
    const dh_plan::Slot& q = p->slots[l * SL_PER_LAYER + SL_QKV];
    const int64_t wk[3] = {o.q_k, o.k_k, o.v_k}, wb[3] = {o.q_b, o.k_b, o.v_b};
    const Part pq[3] = {{P + wk[0], D, D}, {P + wk[1], D, D}, {P + wk[2], D, D}};
    if ((rc = fill(q, pq, 3, false))) return rc;
    for (int t = 0; t < 3; ++t)
      if ((rc = cp(q.bias + (size_t)t * D, wb[t], D))) return rc;
    const dh_plan::Slot& so = p->slots[l * SL_PER_LAYER + SL_O];
    const Part po = {P + o.o_k, D, D};
    if ((rc = fill(so, &po, 1, false))) return rc;
    if ((rc = cp(so.bias, o.o_b, D))) return rc;
    const dh_plan::Slot& s1 = p->slots[l * SL_PER_LAYER + SL_D1];
    const Part p1 = {P + o.d1_k, D, D};
    if ((rc = fill(s1, &p1, 1, false))) return rc;
    // folded out-projection . Dense_{1+2l}:  Wc = Wo @ W1 (fp32 FMA), bc = bo @ W1
    const dh_plan::Slot& sod = p->slots[l * SL_PER_LAYER + SL_OD];
    float* tmp = p->prep + p->fold_tmp;
    if ((rc = gemm_simt(P + o.o_k, P + o.d1_k, nullptr, tmp, D, D, D, D, 1, D, 1, D, 1, 0, 1, s))) return rc;
    const Part pod = {tmp, D, D};
    if ((rc = fill(sod, &pod, 1, false))) return rc;
    if ((rc = gemm_simt(P + o.o_b, P + o.d1_k, nullptr, p->prep + sod.bias, 1, D, D, D, 1, D, 1, D, 1, 0, 1, s))) return rc;
    if (l == 0) {
      // W0 @ (Wq | Wk | Wv): [4][3D]
      for (int t = 0; t < 3; ++t)
        if ((rc = gemm_simt(P + p->off_W0, P + wk[t], nullptr, p->prep + p->w0qkv + (size_t)t * D, 4, D, D, D, 1, D, 1,
                            3 * D, 1, 0, 1, s)))
          return rc;
    }
    const dh_plan::Slot& s2 = p->slots[l * SL_PER_LAYER + SL_D2];
    const Part p2 = {P + o.d2_k, D, D};
    if ((rc = fill(s2, &p2, 1, false))) return rc;
    if ((rc = cp(s2.bias, o.d2_b, D))) return rc;
  }
  // orbital projections, per spin block sb: rows [2 sb LNK, +LNK) = real part, the next LNK = imaginary part; pad rows stay zero
  const dh_plan::Slot& sb = p->slots[p->nl * SL_PER_LAYER];
  Part pb[4];
  for (int t = 0; t < 2 * p->nsb; ++t) pb[t] = {orbW(p, P, t), LNK, LNK};
  if ((rc = fill(sb, pb, 2 * p->nsb, true))) return rc;
  for (int t = 0; t < 2 * p->nsb; ++t)
    DH_CHECK(cudaMemcpyAsync(p->prep + sb.bias + (size_t)t * LNK, orbB(p, P, t), LNK * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (p->orb_fuse) {  // the same projection with permuted output columns (fused envelope contraction of the jet passes)
    const dh_plan::Slot& sp = p->slots[p->nl * SL_PER_LAYER + 1];
    float* Wp = p->prep + p->orb_perm;
    float* bp = Wp + (size_t)D * sp.Nout;
    if ((rc = orb_permute_weights(orbW(p, P, 0), orbW(p, P, 1), orbB(p, P, 0), orbB(p, P, 1), Wp, bp, D, p->L, p->N * p->K, sp.Nout, s)))
      return rc;
    const Part pp = {Wp, sp.Nout, sp.Nout};
    if ((rc = fill(sp, &pp, 1, true))) return rc;
    DH_CHECK(cudaMemcpyAsync(p->prep + sp.bias, bp, sp.Nout * sizeof(float), cudaMemcpyDeviceToDevice, s));
    p->launches += 4;
  }
  p->launches += p->nl * (f16 ? 18 : 11) + 5 + (f16 ? 4 : 2);
  return 0;
}

// Planes for the reverse pass: dX = G @ W^T is C = A . B with A = G and the [N][K] operand = W as stored.
static int prepare_weights_vjp_now(dh_plan* p, const float* P, cudaStream_t s) {
  if (p->gemm_impl != 1) return 0;
  const int D = p->D, LNK = p->LNK, f16 = p->tc_f16;
  int rc;
  auto plane = [&](size_t off_floats) -> void* { return (void*)(p->prep + off_floats); };
  struct Part { const float* W; int64_t ldw; int k; };
  auto fill = [&](const dh_plan::Slot& sl, const Part* parts, int nparts) -> int {
    float* slot = p->prep + sl.scale;
    const size_t bytes = (size_t)D * sl.ldw * (f16 ? 2 : 4);
    DH_CHECK(cudaMemsetAsync(plane(sl.hi), 0, bytes, s));
    DH_CHECK(cudaMemsetAsync(plane(sl.lo), 0, bytes, s));
    DH_CHECK(cudaMemsetAsync(slot, 0, 3 * sizeof(float), s));
    if (f16)
      for (int t = 0; t < nparts; ++t)
        if ((rc = weight_maxabs_tc(parts[t].W, parts[t].ldw, D, parts[t].k, slot, s))) return rc;
    int koff = 0;
    for (int t = 0; t < nparts; ++t) {
      if ((rc = split_weight_nt_tc(parts[t].W, parts[t].ldw, D, parts[t].k, koff, sl.ldw, plane(sl.hi), plane(sl.lo), slot, f16, s)))
        return rc;
      koff += parts[t].k;
    }
    return 0;
  };
  for (int l = 0; l < p->nl; ++l) {
    const LayerOff& o = p->layer[l];
    const Part pd2 = {P + o.d2_k, D, D}, pd1 = {P + o.d1_k, D, D}, po = {P + o.o_k, D, D};
    const Part pq[3] = {{P + o.q_k, D, D}, {P + o.k_k, D, D}, {P + o.v_k, D, D}};
    if ((rc = fill(p->vslots[l * VS_PER_LAYER + VS_D2], &pd2, 1))) return rc;
    if ((rc = fill(p->vslots[l * VS_PER_LAYER + VS_D1], &pd1, 1))) return rc;
    if ((rc = fill(p->vslots[l * VS_PER_LAYER + VS_O], &po, 1))) return rc;
    if ((rc = fill(p->vslots[l * VS_PER_LAYER + VS_QKV], pq, 3))) return rc;
  }
  Part pb[4];
  for (int t = 0; t < 2 * p->nsb; ++t) pb[t] = {orbW(p, P, t), LNK, LNK};
  if ((rc = fill(p->vslots[p->nl * VS_PER_LAYER], pb, 2 * p->nsb))) return rc;
  p->launches += (p->nl * 6 + 2) * (f16 ? 2 : 1);
  return 0;
}

// Called by every op.  With auto_prepare the prepared weights are rebuilt from `P` each time (the library
// cannot see in-place parameter updates); otherwise the caller promised to call dh_params_prepare after
// every update and the planes made there are reused (the reverse-pass planes lazily, on the first VJP).
int prepare_weights(dh_plan* p, const float* P, cudaStream_t s) {
  range_flag_set(p->d_status);  // every op of this plan starts here: its kernels report into this plan's status word
  if (p->auto_prepare) return prepare_weights_now(p, P, s);
  if (p->prep_fwd_valid && p->prep_src == P) return 0;
  int rc = prepare_weights_now(p, P, s);
  if (!rc) { p->prep_src = P; p->prep_fwd_valid = true; p->prep_vjp_valid = false; }
  return rc;
}
int prepare_weights_vjp(dh_plan* p, const float* P, cudaStream_t s) {
  if (p->auto_prepare) return prepare_weights_vjp_now(p, P, s);
  if (p->prep_vjp_valid && p->prep_src == P) return 0;
  int rc = prepare_weights_vjp_now(p, P, s);
  if (!rc) p->prep_vjp_valid = true;
  return rc;
}

extern "C" int dh_params_prepare(dh_plan* p, const float* params, void* stream) {
  if (!p) return DH_E_BADARG;
  if (p->laughlin) return 0;
  if (!params) return DH_E_BADARG;
  int rc = prepare_weights_now(p, params, (cudaStream_t)stream);
  if (rc) return rc;
  p->prep_src = params;
  p->prep_fwd_valid = true;
  p->prep_vjp_valid = false;
  return 0;
}

extern "C" int dh_plan_set_auto_prepare(dh_plan* p, int32_t on) {
  if (!p) return DH_E_BADARG;
  p->auto_prepare = on ? 1 : 0;
  p->prep_fwd_valid = false;
  p->prep_vjp_valid = false;
  return 0;
}

// mv: the (value-only) pass evaluates a Metropolis move -- x receives the proposal drawn from mv->x1, and mv->x1 / mv->lp1 are
// updated by the accept step (forward_chunk); walker0 is ignored on input (set per chunk).
static int run_forward(dh_plan* p, const float* params, const float* x, int64_t B, bool jets, float* out_el,
                       float* out_kin, float* out_pot, float* out_lz, float* out_lz2, float* out_l2,
                       float* out_logpsi, void* ws, size_t ws_bytes, cudaStream_t s, const MoveChunk* mv = nullptr) {
  if (!p || B < 0) return DH_E_BADARG;
  if (B == 0) return 0;
  if ((!params && p->nparams > 0) || !x) return DH_E_BADARG;
  p->vjp_fwd.valid = false;  // this pass takes the workspace
  int copies = 1;
  int64_t chunk = plan_chunks(p, jets, B, &copies);
  float* base = align_ws(ws);
  FwdWs w = carve_fwd(p, base, chunk, jets, false);
  // Chunk interleave: with room for a second activation workspace, odd chunks run on the plan's side stream
  // (forked from and joined back into `s` by events, so the call keeps stream semantics and stays capturable).
  // Chunks are independent; one chunk's launch gaps and wave tails are filled by the other's kernels.
  FwdWs w2 = w;
  const bool dual = copies == 2 && !p->prof_on && p->side_stream && ws &&
                    (size_t)((char*)(base + 2 * w.floats) - (char*)ws) <= ws_bytes;
  if (!dual) {  // single stream: the plain chunking
    chunk = pick_chunk(p, jets, B);
    w = carve_fwd(p, base, chunk, jets, false);
    w2 = w;
  }
  if (!ws || (size_t)((char*)(base + w.floats) - (char*)ws) > ws_bytes) return DH_E_WORKSPACE;
  if (dual) {
    w2 = carve_fwd(p, base + w.floats, chunk, jets, false);
    if (cudaEventRecord(p->ev_fork, s) != cudaSuccess || cudaStreamWaitEvent(p->side_stream, p->ev_fork, 0) != cudaSuccess)
      return (int)cudaGetLastError();
  }
  int rc = 0;
  int64_t k = 0;
  for (int64_t b0 = 0; b0 < B && !rc; b0 += chunk, ++k) {
    const int64_t Bc = (B - b0) < chunk ? (B - b0) : chunk;
    FinalizeArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.out_logpsi = out_logpsi ? out_logpsi + b0 * 2 : nullptr;
    fa.out_el = out_el ? out_el + b0 * 2 : nullptr;
    fa.out_kin = out_kin ? out_kin + b0 * 2 : nullptr;
    fa.out_pot = out_pot ? out_pot + b0 : nullptr;
    fa.out_lz = out_lz ? out_lz + b0 : nullptr;
    fa.out_lz2 = out_lz2 ? out_lz2 + b0 : nullptr;
    fa.out_l2 = out_l2 ? out_l2 + b0 : nullptr;
    const bool odd = dual && (k & 1);
    MoveChunk mc;
    if (mv) mc = MoveChunk{mv->x1 + b0 * p->N * 2, mv->lp1 + b0, mv->dv, b0};
    rc = forward_chunk(p, params, x + b0 * p->N * 2, Bc, jets, odd ? w2 : w, fa, odd ? p->side_stream : s, mv ? &mc : nullptr);
  }
  if (dual) {  // join even after an error, so the side stream never outlives the call
    cudaError_t e = cudaEventRecord(p->ev_join, p->side_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, p->ev_join, 0);
    if (!rc && e != cudaSuccess) rc = (int)e;
  }
  return rc;
}

extern "C" int dh_logpsi(dh_plan* p, const float* params, const float* x, int64_t B, float* out_logpsi,
                         void* ws, size_t ws_bytes, void* stream) {
  if (!out_logpsi && B > 0) return DH_E_BADARG;
  if (p && params && B > 0 && !p->laughlin) { int rc = prepare_weights(p, params, (cudaStream_t)stream); if (rc) return rc; }
  return run_forward(p, params, x, B, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, out_logpsi, ws,
                     ws_bytes, (cudaStream_t)stream);
}

extern "C" int dh_local_energy(dh_plan* p, const float* params, const float* x, int64_t B, float* out_el,
                               float* out_kinetic, float* out_potential, float* out_lz, float* out_lz2,
                               float* out_l2, float* out_logpsi, void* ws, size_t ws_bytes, void* stream) {
  if (p && params && B > 0 && !p->laughlin) { int rc = prepare_weights(p, params, (cudaStream_t)stream); if (rc) return rc; }
  return run_forward(p, params, x, B, true, out_el, out_kinetic, out_potential, out_lz, out_lz2, out_l2,
                     out_logpsi, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dh_potential(dh_plan* p, const float* x, int64_t B, float* out, void* stream) {
  if (!p || !x || !out) return DH_E_BADARG;
  if (B == 0) return 0;
  return potential(x, out, B, p->N, p->Q, p->radius, p->cfg.interaction_type, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------------- MCMC
extern "C" int dh_mcmc_propose(dh_plan* p, const float* x1, int64_t B, float width, uint64_t seed,
                               uint64_t offset, uint64_t subsequence0, const float* randoms, float* x2,
                               void* stream) {
  if (!p || !x1 || !x2) return DH_E_BADARG;
  if (B == 0) return 0;
  return mcmc_propose(x1, x2, B, p->N, width, seed, offset, subsequence0, randoms, (cudaStream_t)stream);
}

extern "C" int dh_mcmc_accept(dh_plan* p, float* x1, const float* x2, float* lp1, const float* lp2, int64_t B,
                              uint64_t seed, uint64_t offset, uint64_t subsequence0, const float* randoms,
                              long long* naccept, void* stream) {
  if (!p || !x1 || !x2 || !lp1 || !lp2 || !naccept) return DH_E_BADARG;
  if (B == 0) return 0;
  return mcmc_accept(x1, x2, lp1, lp2, 1, B, p->N, seed, offset, subsequence0, randoms,
                     reinterpret_cast<unsigned long long*>(naccept), (cudaStream_t)stream);
}

extern "C" int dh_init_walkers(dh_plan* p, float* x, int64_t B, uint64_t seed, uint64_t subsequence0,
                               void* stream) {
  if (!p || !x) return DH_E_BADARG;
  if (B == 0) return 0;
  return init_walkers(x, B, p->N, seed, subsequence0, (cudaStream_t)stream);
}

// The sweep proper.  dev_args: the move arguments (Philox key / offset, width, first subsequence) are already in the plan's
// device block (dh_mcmc_sweep_dev); otherwise they are the host scalars given here.
static int mcmc_sweep_impl(dh_plan* p, const float* params, float* x, int64_t B, int32_t steps, float width, uint64_t seed,
                           uint64_t offset, uint64_t subsequence0, const float* randoms, bool dev_args, long long* out_naccept,
                           float* out_lp, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (B == 0) return (int)cudaMemsetAsync(out_naccept, 0, sizeof(long long), s);
  float* base = align_ws(ws);
  McmcWs mw = carve_mcmc(p, base, B);
  float* fbase = base + mw.floats;
  const int64_t chunk = pick_chunk(p, false, B);
  FwdWs fw = carve_fwd(p, fbase, chunk, false, false);
  if (!ws || (size_t)((char*)(fbase + fw.floats) - (char*)ws) > ws_bytes) return DH_E_WORKSPACE;
  size_t fwd_bytes = ws_bytes - (size_t)((char*)fbase - (char*)ws);
  int rc;
  DH_CHECK(cudaMemsetAsync(out_naccept, 0, sizeof(long long), s));
  // mcmc.py:142 -- log-probability of the incoming configurations
  if (!p->laughlin && (rc = prepare_weights(p, params, s))) return rc;
  if ((rc = run_forward(p, params, x, B, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, mw.logpsi2, fbase,
                        fwd_bytes, s)))
    return rc;
  if ((rc = lp_from_logpsi(mw.logpsi2, mw.lp1, B, s))) return rc;
  const int64_t rstride = B * (2 * (int64_t)p->N + 1);
  // ---- production path: in-kernel Philox, the move replayed from a captured CUDA graph with its arguments in a device block
  static const bool graphs_env = !(dbg_env("DH_MCMC_GRAPH") && atoi(dbg_env("DH_MCMC_GRAPH")) == 0);
  const bool dev_ok = !randoms && p->d_mcmc && (p->laughlin || p->raw_params);
  if (dev_args && !dev_ok) return DH_E_UNSUPPORTED;
  if (dev_ok) {
    const float* Pg = p->laughlin ? params : p->raw_params;
    if (!p->laughlin) DH_CHECK(cudaMemcpyAsync(p->raw_params, params, p->nparams * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (!dev_args && (rc = mcmc_dev_init(p->d_mcmc, seed, offset, subsequence0, width, s))) return rc;
    const MoveChunk move{x, mw.lp1, p->d_mcmc, 0};  // proposal and accept / select run inside the pass (forward_chunk)
    dh_plan::MoveGraph* mg = nullptr;
    const bool want_graph = steps >= 2 && !p->prof_on && p->graphs_ok && graphs_env;
    if (want_graph)
      for (auto& g : p->move_graphs) if (g.x == x && g.ws == ws && g.B == B) mg = &g;
    if (want_graph && !mg) {
      cudaGraph_t graph = nullptr;
      const long long l0 = p->launches;
      if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        rc = run_forward(p, Pg, mw.x2, B, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, mw.logpsi2, fbase, fwd_bytes, s, &move);
        if (!rc) rc = mcmc_dev_advance(p->d_mcmc, s);
        const cudaError_t ce = cudaStreamEndCapture(s, &graph);
        cudaGraphExec_t exec = nullptr;
        if (!rc && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
          if (p->move_graphs.size() >= 4) { cudaGraphExecDestroy(p->move_graphs.front().exec); p->move_graphs.erase(p->move_graphs.begin()); }
          p->move_graphs.push_back({x, ws, B, exec, p->launches - l0 + 1});
          mg = &p->move_graphs.back();
        } else {
          p->graphs_ok = 0;
          cudaGetLastError();
        }
        if (graph) cudaGraphDestroy(graph);
        p->launches = l0;
      } else {
        p->graphs_ok = 0;
        cudaGetLastError();
      }
    }
    if (mg) {
      for (int st = 0; st < steps; ++st) DH_CHECK(cudaGraphLaunch(mg->exec, s));
      p->launches += mg->launches * steps;
    } else {  // no graph (one move, profiling, or capture refused): the same kernels launched one by one
      for (int st = 0; st < steps; ++st) {
        if ((rc = run_forward(p, Pg, mw.x2, B, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, mw.logpsi2, fbase, fwd_bytes, s, &move))) return rc;
        { ProfScope ps(p, PC_MCMC, 0, s);
          if ((rc = mcmc_dev_advance(p->d_mcmc, s))) return rc; }
      }
    }
    DH_CHECK(cudaMemcpyAsync(out_naccept, &p->d_mcmc->naccept, sizeof(long long), cudaMemcpyDeviceToDevice, s));
    if (out_lp) DH_CHECK(cudaMemcpyAsync(out_lp, mw.lp1, B * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  // injected randoms (bit-exact parity tests against the oracle's decisions): the three public steps' kernels, launch by launch
  for (int st = 0; st < steps; ++st) {
    const float* rnd = randoms ? randoms + st * rstride : nullptr;
    { ProfScope ps(p, PC_MCMC, 0, s); if ((rc = mcmc_propose(x, mw.x2, B, p->N, width, seed, offset + st, subsequence0, rnd, s))) return rc; }
    if ((rc = run_forward(p, params, mw.x2, B, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, mw.logpsi2,
                          fbase, fwd_bytes, s)))
      return rc;
    { ProfScope ps(p, PC_MCMC, 0, s);
      if ((rc = mcmc_accept(x, mw.x2, mw.lp1, mw.logpsi2, 2, B, p->N, seed, offset + st, subsequence0, rnd,
                            reinterpret_cast<unsigned long long*>(out_naccept), s)))
        return rc; }
  }
  if (out_lp) DH_CHECK(cudaMemcpyAsync(out_lp, mw.lp1, B * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" int dh_mcmc_sweep(dh_plan* p, const float* params, float* x, int64_t B, int32_t steps, float width,
                             uint64_t seed, uint64_t offset, uint64_t subsequence0, const float* randoms,
                             long long* out_naccept, float* out_lp, void* ws, size_t ws_bytes, void* stream) {
  if (!p || (!params && p->nparams > 0) || !x || !out_naccept || steps < 0) return DH_E_BADARG;
  return mcmc_sweep_impl(p, params, x, B, steps, width, seed, offset, subsequence0, randoms, false, out_naccept, out_lp, ws, ws_bytes,
                         (cudaStream_t)stream);
}

extern "C" int dh_mcmc_sweep_dev(dh_plan* p, const float* params, float* x, int64_t B, int32_t steps, const float* width_dev,
                                 const uint64_t* key_dev, uint64_t subsequence0, long long* out_naccept, float* out_lp, void* ws,
                                 size_t ws_bytes, void* stream) {
  if (!p || (!params && p->nparams > 0) || !x || !out_naccept || steps < 0 || !width_dev || !key_dev) return DH_E_BADARG;
  if (!p->d_mcmc) return DH_E_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = mcmc_dev_init_from(p->d_mcmc, reinterpret_cast<const unsigned long long*>(key_dev), width_dev, subsequence0, s);
  if (rc) return rc;
  return mcmc_sweep_impl(p, params, x, B, steps, 0.f, 0, 0, subsequence0, nullptr, true, out_naccept, out_lp, ws, ws_bytes, s);
}

// --------------------------------------------------------------------------------- energy statistics (loss.py:66-92)
extern "C" int dh_energy_stats(const float* el, const float* kinetic, const float* potential, const float* lz, const float* lz2,
                               const float* l2, int64_t B, float* out16, void* stream) {
  if (!el || !kinetic || !potential || !lz || !lz2 || !l2 || !out16) return DH_E_BADARG;
  if (B < 1 || B > 32768) return DH_E_UNSUPPORTED;
  return energy_stats(el, kinetic, potential, lz, lz2, l2, B, out16, (cudaStream_t)stream);
}

extern "C" int dh_energy_diff(const float* el, const float* lz, const float* lz2, const float* l2, const float* logpsi, int64_t B,
                              const float* reduced16, float lz_penalty, float lz_center, float l2_penalty, float* out_diff,
                              float* out_cot, float* out_ok, float* out_counts, void* stream) {
  if (!el || !reduced16 || !out_diff || !out_cot) return DH_E_BADARG;
  if ((lz_penalty != 0.f && (!lz || !lz2)) || (l2_penalty != 0.f && !l2)) return DH_E_BADARG;
  if (B < 1 || B > 32768) return DH_E_UNSUPPORTED;
  return energy_diff(el, lz, lz2, l2, logpsi, B, reduced16, lz_penalty, lz_center, l2_penalty, out_diff, out_cot, out_ok, out_counts,
                     (cudaStream_t)stream);
}

extern "C" int dh_slogdet(const float* mats, int64_t B, int32_t K, int32_t n, float* out_sign,
                          float* out_logabs, float* out_logpsi, void* stream) {
  if (!mats) return DH_E_BADARG;
  if (B == 0) return 0;
  return slogdet_batched(mats, B, K, n, out_sign, out_logabs, out_logpsi, (cudaStream_t)stream);
}

extern "C" int dh_spd_inverse(float* mats, int32_t n, int32_t batch, void* stream) {
  if (!mats && batch > 0) return DH_E_BADARG;
  if (n > 1024) return DH_E_UNSUPPORTED;
  return spd_inverse_batched(mats, n, batch, (cudaStream_t)stream);
}

extern "C" int dh_gemm_workspace_bytes(int32_t N, int32_t K, int32_t impl, size_t* bytes) {
  if (!bytes || N < 1 || K < 1) return DH_E_BADARG;
  const size_t npad = (size_t)((N + 15) & ~15);
  *bytes = impl == 0 ? 0 : (2 * npad * K + 64) * sizeof(float);  // split weight planes + the scale slot
  return 0;
}

extern "C" int dh_gemm(const float* A, const float* W, const float* bias, float* C, int64_t M, int32_t N,
                       int32_t K, int32_t rows_per_group, int32_t accumulate, int32_t impl, void* ws, size_t ws_bytes,
                       void* stream) {
  if (!A || !W || !C) return DH_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  range_flag_set(nullptr);  // plan-free entry point: no status word to report into
  if (impl == 0) return gemm_simt(A, W, bias, C, M, N, K, K, 1, N, 1, N, rows_per_group, accumulate, 1, s);
  if (!gemm_tc_supported(N, K) || accumulate) return DH_E_UNSUPPORTED;
  // test / bench / optimizer entry point: the split weights are made on the fly in the caller's workspace (the plan
  // ops keep theirs cached).  Stream-ordered, no allocation, no synchronisation -- like every other entry point.
  // impl 1: tcgen05 with fp16 hi/lo pieces, one accumulator per tile (the default of the plan ops) when K % 32 == 0;
  // impl 2: tcgen05 with TF32 hi/lo pieces;  impl 3: fp16 pieces with separate main / correction accumulators.
  size_t need = 0;
  dh_gemm_workspace_bytes(N, K, impl, &need);
  if (!ws || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 15)) return DH_E_WORKSPACE;
  const int f16 = ((impl == 1 || impl == 3) && gemm_tc_f16_ok(K)) ? 1 : 0;
  const int merged = (impl == 1 && f16) ? 1 : 0;
  float* wt = static_cast<float*>(ws);
  const size_t npad = (size_t)((N + 15) & ~15);
  float* slot = wt + 2 * npad * K;
  int rc = (int)cudaMemsetAsync(slot, 0, 3 * sizeof(float), s);
  if (!rc && f16) rc = weight_maxabs_tc(W, N, K, N, slot, s);
  if (!rc) rc = split_weight_tc(W, N, K, N, 1, wt, wt + npad * K, slot, f16, s);
  if (!rc) rc = gemm_tc(A, wt, wt + npad * K, bias, f16 ? slot + 1 : nullptr, C, M, N, K, N, rows_per_group, f16, merged, s);
  return rc;
}

extern "C" int dh_debug_buffer(const dh_plan* p, int op, int64_t B, const char* name, int64_t* offset,
                               int64_t* count) {
  if (!p || !name || !offset || !count) return DH_E_BADARG;
  const bool jets = op == DH_OP_LOCAL_ENERGY;
  if (op != DH_OP_LOCAL_ENERGY && op != DH_OP_LOGPSI) return DH_E_BADARG;
  const int64_t Bc = pick_chunk(p, jets, B);
  float* zero = reinterpret_cast<float*>(uintptr_t(256));
  FwdWs w = carve_fwd(p, zero, Bc, jets, false);
  const int R = jets ? 2 * p->N + 8 : 1;
  const int64_t rows = Bc * p->N * R;
  const std::string n(name);
  const float* ptr = nullptr;
  int64_t cnt = 0;
  if (n == "h") { ptr = w.h; cnt = rows * p->D; }
  else if (n == "qkv") { ptr = w.qkv; cnt = rows * 3 * p->D; }
  else if (n == "attn") { ptr = w.att; cnt = rows * p->D; }
  else if (n == "t1") { ptr = w.t1; cnt = rows * p->D; }
  else if (n == "t2") { ptr = w.t2; cnt = rows * p->D; }
  else if (n == "c") { ptr = w.cbuf; cnt = rows * (int64_t)p->orbN; }
  else if (n == "orb") { ptr = w.Mj; cnt = Bc * p->K * R * p->N * p->N * 2; }
  else if (n == "ld") { ptr = w.ld; cnt = Bc * p->K * R * 2; }
  else if (n == "lpjet") { ptr = w.lpjet; cnt = Bc * R * 2; }
  else return DH_E_BADARG;
  *offset = ptr - zero;  // floats from the 256-byte-aligned workspace base
  *count = cnt;
  return 0;
}

// --------------------------------------------------------------------------------- instrumentation
extern "C" long long dh_launch_count(const dh_plan* p) { return p ? p->launches : -1; }

extern "C" int dh_profile_begin(dh_plan* p, int32_t max_launches) {
  if (!p || max_launches < 1) return DH_E_BADARG;
  while (p->prof_ev.size() < (size_t)max_launches * 2) {
    cudaEvent_t e;
    DH_CHECK(cudaEventCreate(&e));
    p->prof_ev.push_back(e);
  }
  p->prof_cat.assign(p->prof_ev.size() / 2, 0);
  p->prof_flops.assign(p->prof_ev.size() / 2, 0.0);
  p->prof_used = 0;
  p->prof_on = true;
  return 0;
}

// Synchronises the device, then reports per category: total ms, number of timed scopes, flops.
extern "C" int dh_profile_end(dh_plan* p, double* ms_host, int32_t* count_host, double* flops_host) {
  if (!p || !ms_host || !count_host || !flops_host) return DH_E_BADARG;
  p->prof_on = false;
  DH_CHECK(cudaDeviceSynchronize());
  for (int c = 0; c < PC_COUNT; ++c) { ms_host[c] = 0; count_host[c] = 0; flops_host[c] = 0; }
  for (size_t i = 0; i + 1 < p->prof_used; i += 2) {
    float ms = 0.f;
    DH_CHECK(cudaEventElapsedTime(&ms, p->prof_ev[i], p->prof_ev[i + 1]));
    const int c = p->prof_cat[i / 2];
    ms_host[c] += ms;
    count_host[c] += 1;
    flops_host[c] += p->prof_flops[i / 2];
  }
  p->prof_used = 0;
  return 0;
}
