// Internal launch interface between api.cu and the kernel translation units.
#pragma once
#include "common.cuh"

namespace dh {

// ---- gemm_simt.cu
int gemm_simt(const float* A, const float* B, const float* bias, float* C, int64_t M, int N, int64_t K,
              int64_t a_sm, int64_t a_sk, int64_t b_sk, int64_t b_sn, int64_t ldc, int rpg, int accumulate,
              int split_k, cudaStream_t stream);

// ---- gemm_tc.cu (tcgen05, two-piece operand split): C[M,N] = A[M,K] @ W[K,N] * (*inv_scale) (+bias on value rows)
//      Wt_hi / Wt_lo are the pre-split, K-major (transposed, [Npad][K]) copies of W: TF32-valued fp32
//      (f16 = 0) or fp16 with a per-slot power-of-two scale (f16 = 1; needs K % 64 == 0).
//      merged = 1: one TMEM accumulator per tile, double-buffered (epilogue overlaps the next tile);
//      merged = 0: separate main / correction accumulators (fewer truncating additions, no overlap).
//      scale slot = 3 floats {max|W| bits, 1/scale, scale}; zero it, run weight_maxabs_tc over every
//      block stacked into the slot, then split_weight_tc each block.
int gemm_tc_supported(int N, int K);
int gemm_tc_f16_ok(int K);
int gemm_tc(const float* A, const void* Wt_hi, const void* Wt_lo, const float* bias, const float* inv_scale, float* C,
            int64_t M, int N, int K, int64_t ldc, int rpg, int f16, int merged, cudaStream_t stream);
// General form: A row stride lda (floats, multiple of 4), any K >= 1 (the planes are [Npad][ldw] with ldw = K rounded
// up to 32 and the tail zero), reduce_add = 1 adds the product to C (TMA reduce) instead of storing it.
struct TcGemm {
  const float* A; int64_t lda;
  const void *Wt_hi, *Wt_lo; int64_t ldw;
  const float *bias, *inv_scale;
  float* C; int64_t ldc;
  int64_t M; int N, K, rpg, f16, merged, reduce_add;
  const float* a_scale;  // optional device {s, 1/s}: A is multiplied by s before the split, C by 1/s
  // fused envelope contraction (jet passes at N = 12, one determinant, pair form with resident A): the contraction is the
  // orbital projection with its weights in the permuted [tile][m][re | im][column] layout (orb_permute_weights) and the
  // epilogue contracts the coefficients with the per-electron envelope table orb_env [electrons][10][orb_L] complex into the
  // orbital-matrix jets orb_Mj -- C is not written; null = off
  const float* orb_env; float* orb_Mj; int orb_L;
  // fused value LayerNorm epilogue (value-only passes, pair form, N = 256, rpg = 1): C = LN(ln_res + acc + bias) or
  // LN(ln_res + tanh(acc + bias)), written in place over ln_res (= C); null = off
  const float* ln_res; const float* ln_gamma; const float* ln_beta; int ln_tanh;
  const void* A_lo;      // non-null (fp16 pieces only): A and A_lo are fp16 hi / lo planes [M][lda] written by the
                         // producing kernel; the in-kernel split is skipped
};
// "TN" contraction on the tensor cores (reverse-pass dW = X^T dY, KFAC Gram sums): C[Ma, Nb] (ldc) += A[rows, Ma]^T B[rows, Nb]
// / (a_scale * b_scale); a_scale / b_scale: optional device {s, 1/s} applied to the operand before the fp16 split
int gemm_tn_tc_ok(const float* A, int64_t lda, const float* B, int64_t ldb, const float* C, int64_t ldc);
int gemm_tn_tc(const float* A, int64_t lda, int Ma, const float* B, int64_t ldb, int Nb, float* C, int64_t ldc, int64_t rows,
               const float* a_scale, const float* b_scale, cudaStream_t stream);
// slot = {s, 1/s}, s = power of two bringing max|v| into [1, 2)
int pow2_scale_tc(const float* v, int64_t n, float* slot, cudaStream_t stream);
int gemm_tc_ex(const TcGemm& g, cudaStream_t stream);
// planes[n][k_off + k] = W[n*ldw + k] * scale (no transpose); the slot must hold max|W| (weight_maxabs_tc)
int split_weight_nt_tc(const float* W, int64_t ldw, int N, int Kpart, int k_off, int64_t ldp, void* hi, void* lo,
                       float* scale_slot, int f16, cudaStream_t stream);
int weight_maxabs_tc(const float* W, int64_t ldw, int K, int N, float* scale_slot, cudaStream_t stream);
int split_weight_tc(const float* W, int64_t ldw, int K, int N, int pad_rows, void* Wt_hi, void* Wt_lo,
                    float* scale_slot, int f16, cudaStream_t stream);

// ---- net_kernels.cu
struct NetDims {
  int N;       // electrons
  int R;       // rows per electron (1 or 2N+8)
  int D;       // model width = H*hd
  int H, hd;
  int n_up;    // spin-up count (for the spin feature)
};
int features_dense0(const float* x, const float* W0, float* h, int64_t B, NetDims d, cudaStream_t s);
// out[rows, Nout] = feature jets @ W[4][Nout] (+ bias on value rows)
//   compressed != 0 (jets only): only the 10 rows per electron that are non-zero are written, in the order
//   value | own tangent flows (2) | S | D_a (3) | T_a (3)
int features_linear(const float* x, const float* W, const float* bias, float* out, int Nout, int64_t B, NetDims d,
                    int compressed, cudaStream_t s);
// out = LN(a + (tanh_mode ? tanh(b) : b)) with jets; a/b/out are [B*N*R, D]
int residual_layernorm(const float* a, const float* b, const float* scale, const float* bias, float* out,
                       int64_t B, NetDims d, int tanh_mode, cudaStream_t s);
// a_comp != 0 (jets): `a` holds only the 10 non-zero rows per electron of the first layer's Dense_0 output
// a_pl / o_pl != 0 (D = 256): `a` / `out` are fp16 hi / lo planes (common.cuh) instead of fp32 rows
int residual_layernorm_ex(const float* a, const float* b, const float* scale, const float* bias, float* out,
                          int64_t B, NetDims d, int tanh_mode, int a_comp, int a_pl, int o_pl, cudaStream_t s);
int attention_value(const float* qkv, float* o, int64_t B, NetDims d, cudaStream_t s);
// layer0 != 0: qkv is the compressed first-layer tensor [B*N*10][3D] (features_linear with compressed = 1)
// fp32_only != 0: never the fp16-piece tensor-core form (plans whose contraction mode is tf32 / fp32)
int attention_jets(const float* qkv, float* o, int64_t B, NetDims d, int layer0, int fp32_only, cudaStream_t s);
size_t attention_jets_smem(NetDims d);
// Device status word of the running plan: the fp16-piece kernels (gemm_tc.cu, attention_tc.cu) OR bit 0 into it when an
// operand piece saturates fp16's range (dh_plan_status).  Set by the API entry points; nullptr = no reporting.
void range_flag_set(unsigned* flag);
unsigned* range_flag_get();
// attention_tc.cu: the same op with its contractions as mma.sync (fp16 hi / lo split), hd = 64, N in {3, 6, 10, 12, 16}
bool attention_jets_tc_ok(NetDims d);
int attention_jets_tc(const float* qkv, float* o, int64_t B, NetDims d, int layer0, cudaStream_t s);

// ---- tail_kernels.cu
struct TailDims {
  int N, R, L, K;   // electrons, rows, orbitals (2Q+1), determinants
  int twoQ;
  int n_up, n_dn;   // n_dn > 0: two spin blocks of orbital coefficients per row, electron i >= n_up reads the second
  int lskip = -1;   // Laughlin quasihole: exponent index left out of the L = N + 1 orbitals (-1: none, L = N)
  int qp = 0;       // Laughlin quasiparticle (L = N - 1 shell orbitals + one LLL-projected orbital)
  int qp_a = 0;     // its u exponent Q1 + lz, in [-1, 2 Q1 + 1]
};
// c: [B*N*R, 2*nsb*L*N*K] (per spin block: re block | im block) -> M jets [B][K][R][N][N] complex
// per electron: envelope jets [10][L] complex, multiplied by *unscale, followed by the bias products [10][N K] complex
// (sum_m bias(m, j) env_s[m]): the right operand of the fused envelope contraction (gemm_tc.cu, ORB)
int envelope_table(const float* x, const double* normfac, const float* bre, const float* bim, const float* unscale, float* tab,
                   int64_t B, TailDims d, cudaStream_t s);
// jet-pass prologue in one launch: compressed Dense_0 output h [B N][10][n0], compressed first-layer q|k|v [B N][10][n1]
// (W1 [4][n1] + bias b1) and the envelope table
int jets_prologue(const float* x, const float* W0, float* h, int n0, const float* W1, const float* b1, float* q, int n1,
                  const double* normfac, const float* bre, const float* bim, const float* unscale, float* tab, int64_t B, int n_up,
                  TailDims d, cudaStream_t s);
// value-only form: per electron [L] complex envelope (times *unscale) + [N K] complex bias products
int envelope_value_table(const float* x, const double* normfac, const float* bre, const float* bim, const float* unscale, float* tab,
                         int64_t B, TailDims d, cudaStream_t s);
// Column tiles of the fused envelope contraction (gemm_tc.cu, ORB): the L orbitals are dealt out EVENLY over ceil(L / 10) tiles
// of 256 weight columns (24 columns per orbital), so that no tile's tensor work is much shorter than its neighbour's epilogue
// (L = 34: 9, 9, 9, 7 orbitals = 224, 224, 224, 176 tensor columns instead of 256, 256, 256, 96)
inline int orb_tiles(int L) { return (L + 9) / 10; }
inline int orb_per_tile(int L) { return (L + orb_tiles(L) - 1) / orb_tiles(L); }
inline int orb_columns(int L) { return (orb_tiles(L) - 1) * 256 + (L - orb_per_tile(L) * (orb_tiles(L) - 1)) * 24; }
// orbital-projection kernels / biases -> fp32 [D][ncol] + [ncol] with columns ordered [tile][m (10)][re | im][NK]
int orb_permute_weights(const float* Wre, const float* Wim, const float* bre, const float* bim, float* Wp, float* bp, int D, int L,
                        int NK, int ncol, cudaStream_t s);  // ncol = orb_columns(L); orb_per_tile(L) orbitals per 256-column tile
int orbital_contract(const float* c, const float* x, const double* normfac, float* Mj, int64_t B, TailDims d,
                     cudaStream_t s);
// Laughlin ground-state orbital matrix jets (networks/laughlin.py:59-71): Mj [B][1][R][N][N] complex; d.L == d.N,
// d.twoQ = 2 Q1 = N - 1, `ones` = L doubles equal to 1
int laughlin_orbital_jets(const float* x, const double* ones, float* Mj, int64_t B, TailDims d, cudaStream_t s);
// M jets -> per-det logdet jets ld [B][K][R] complex
int logdet_jets(const float* Mj, float* ld, int64_t B, TailDims d, cudaStream_t s);
// same, also writing the inverse of the value matrix ([b][kd][N][N] complex) when Minv != nullptr
int logdet_jets_impl(const float* Mj, float* ld, float* Minv, int64_t B, TailDims d, cudaStream_t s);
struct FinalizeArgs {
  const float* ld;       // [B][K][R] complex
  const float* x;        // [B][N][2]
  const float* ee_par;   // jastrow parameter of the parallel-spin pairs (device scalar) or nullptr when there are none
  const float* ee_anti;  // same for the anti-parallel pairs
  float Q, radius, interaction_strength;
  int interaction_type;
  float* out_logpsi;     // [B] complex
  float* out_el;         // [B] complex   (jets only)
  float* out_kin;        // [B] complex
  float* out_pot, *out_lz, *out_lz2, *out_l2;  // [B]
  float* lpjet;          // optional [B][R] complex debug copy of the log psi jets
  // value-only passes (R = 1, N <= 16): the orbital matrices [B][K][N][N] complex -- the warp takes the log-determinants
  // itself (no separate launch, `ld` unused)
  const float* Mj_value;
  // Metropolis move evaluated by this pass (value-only): accept / select (mcmc.py:56-62) in the same warp -- the "accept ->
  // forward epilogue" fusion.  x = the proposal; mv_x1 / mv_lp1 = current configuration / log-probability of the chunk's
  // walkers, updated in place; mv_walker0 = the chunk's first walker in the rank's batch (Philox subsequence).
  float* mv_x1;
  float* mv_lp1;
  struct McmcDev* mv_dv;
  int64_t mv_walker0;
};
int finalize(FinalizeArgs a, int64_t B, TailDims d, cudaStream_t s);
// sparse orbitals (blocks.py:52-62).  W8 [D][8][NK], b8 [8][NK], Wl [8][L], bl [L] (added when add_bl != 0: the real
// part) -> effective full projection Weff [(D + 1)][L][NK] whose last row is the bias; and the reverse map, which
// ADDS into g_W8, g_b8, g_Wl (and g_bl when add_bl != 0) the gradients implied by g_Weff [(D + 1)][L][NK].
int sparse_fold(const float* W8, const float* b8, const float* Wl, const float* bl, int add_bl, float* Weff, int D, int L,
                int NK, cudaStream_t s);
int sparse_fold_bwd(const float* g_Weff, const float* W8, const float* b8, const float* Wl, int add_bl, float* g_W8,
                    float* g_b8, float* g_Wl, float* g_bl, int D, int L, int NK, cudaStream_t s);
// g8[row][s][jk] = sum_m g[row][m][jk] Wl[s][m]: the gradient with respect to the 8-feature projection's own output
// (the KFAC output factor of the sparse orbitals); g rows have stride ldg, g8 rows 8 NK
int sparse_g8(const float* g, int64_t ldg, const float* Wl, float* g8, int64_t rows, int L, int NK, cudaStream_t s);
int potential(const float* x, float* out, int64_t B, int N, float Q, float radius, int interaction_type,
              cudaStream_t s);
int slogdet_batched(const float* mats, int64_t B, int K, int n, float* out_sign, float* out_logabs,
                    float* out_logpsi, cudaStream_t s);

// ---- mcmc_kernels.cu
int mcmc_propose(const float* x1, float* x2, int64_t B, int N, float width, uint64_t seed, uint64_t offset,
                 uint64_t subseq0, const float* randoms, cudaStream_t s);
int mcmc_accept(float* x1, const float* x2, float* lp1, const float* lp2c, int lp2_stride, int64_t B, int N,
                uint64_t seed, uint64_t offset, uint64_t subseq0, const float* randoms,
                unsigned long long* naccept, cudaStream_t s);
int init_walkers(float* x, int64_t B, int N, uint64_t seed, uint64_t subseq0, cudaStream_t s);
int lp_from_logpsi(const float* logpsi_c, float* lp, int64_t B, cudaStream_t s);
// ---- stats_kernels.cu: energy statistics of a walker batch (loss.py:30-38,66-92); B <= 32768 per rank
int energy_stats(const float* el, const float* kin, const float* pot, const float* lz, const float* lz2, const float* l2, int64_t B,
                 float* out16, cudaStream_t s);
int energy_diff(const float* el, const float* lz, const float* lz2, const float* l2, const float* lp, int64_t B, const float* red_stats,
                float lz_penalty, float lz_center, float l2_penalty, float* diff, float* cot, float* ok, float* counts, cudaStream_t s);
// device-resident arguments of a Metropolis move: lets one captured CUDA graph of a move be replayed for every move
struct McmcDev { unsigned long long seed, offset, subseq0, naccept; float width; };
int mcmc_dev_init(McmcDev* dv, uint64_t seed, uint64_t offset, uint64_t subseq0, float width, cudaStream_t s);
int mcmc_dev_init_from(McmcDev* dv, const unsigned long long* key, const float* width, uint64_t subseq0, cudaStream_t s);
int mcmc_dev_advance(McmcDev* dv, cudaStream_t s);
// walker0: index of x1's first walker in the rank's batch (its Philox subsequence is dv->subseq0 + walker0 + b)
int mcmc_propose_dev(const float* x1, float* x2, int64_t B, int N, const McmcDev* dv, int64_t walker0, cudaStream_t s);
int mcmc_accept_dev(float* x1, const float* x2, float* lp1, const float* lp2c, int lp2_stride, int64_t B, int N, McmcDev* dv,
                    int64_t walker0, cudaStream_t s);

// ---- value_path.cu: the coordinate-only prologue of a value-only pass in one launch
struct ValuePrologue {
  const float* x;        // [rows][2] coordinates (a move: the CURRENT configuration)
  float* x_new;          // a move: the proposal is written here and the pass evaluates it; else nullptr
  const McmcDev* dv;     // move arguments (device block), needed with x_new
  int64_t walker0;       // index of the first walker in the rank's batch (Philox subsequence)
  const float* W0; float* h; int n0;                     // first map [4][n0] (no bias) -> h [rows][n0]
  const float* W1; const float* b1; float* q; int n1;    // second map [4][n1] + bias -> q [rows][n1]; W1 = nullptr: none
  // envelope table of the fused orbital epilogue [rows][2 (L + NK)]; tab = nullptr: none
  const double* normfac; const float* bre; const float* bim; const float* unscale; float* tab; int L, NK, twoQ;
};
int value_prologue(const ValuePrologue& a, int64_t rows, int N, int n_up, cudaStream_t s);

// ---- vjp_kernels.cu
int tail_bwd(const float* cot, const float* ld, const float* Minv, const float* x, const double* normfac,
             float* g_c, int64_t B, TailDims d, cudaStream_t s);
int jastrow_bwd(const float* cot, const float* x, const float* ee_par, const float* ee_anti, float* g_eepar,
                float* g_eeanti, float* sq_eepar, float* sq_eeanti, int64_t B, int N, int n_up, cudaStream_t s);
int spd_inverse_batched(float* mats, int n, int batch, cudaStream_t s);

// ---- kfac_kernels.cu: the KFAC update from the moving-average statistics (descriptor tables: plan.h / api_vjp.cu)
struct KfBlkDesc {   // one dense curvature block
  long long ko, bo;        // kernel / bias offset in the parameter vector (bo < 0: no bias)
  long long xtx, gtg;      // factor-vector offsets of sum x x^T (< 0: Dense_0, the caller's 4 x 4 matrix) and sum g g^T
  long long v_off;         // float offset of the block's [din + hb][dout] gradient / update in the gather buffers
  int din, dout, hb, npw;
};
struct KfMatDesc {   // one damped Kronecker factor
  long long src, xsum;     // factor-vector offsets of the core matrix (< 0: Dense_0's) and of sum x (bias blocks)
  long long dst;           // float offset inside its batch (slot * dim * dim)
  int n_core, n, dim, cls, blk, is_g;   // core size, size with the bias row, padded size, batch (0 small / 1 large), block
};
struct KfDiagDesc { long long ko, o, n; };   // diagonal block: parameter offset, factor-vector offset, size
int kfac_damped_factors(const KfBlkDesc* bd, int nblk, const KfMatDesc* md, int nmat, const float* stats, const float* xtx0,
                        float weight, float damping, float* coef, float* batch_s, float* batch_l, cudaStream_t s);
int kfac_gather(const KfBlkDesc* bd, int nblk, const float* grads, float* V, cudaStream_t s);
int kfac_grouped_gemm(const KfBlkDesc* bd, const KfMatDesc* md, int nblk, int max_rows, int max_cols, const float* inv_s,
                      const float* inv_l, const float* X, float* Y, int stage, cudaStream_t s);
int kfac_scatter(const KfBlkDesc* bd, int nblk, const KfDiagDesc* dd, int ndiag, const float* U, const float* coef,
                 const float* stats, float weight, float damping, const float* grads, float* out, cudaStream_t s);
// KFAC factor pass (dh_kfac_factors)
int fill_unit_cot(float* cot, int64_t n, cudaStream_t s);
int mask_rows_by_spin(const float* src, float* dst, int64_t rows, int D, int N, int n_up, int sb, cudaStream_t s);
int ln_fisher_diag(const float* a, const float* b, const float* gy, float* sq_scale, float* sq_bias, int64_t B, int N, int D,
                   int tanh_mode, cudaStream_t s);
int residual_layernorm_bwd(const float* a, const float* b, const float* scale, const float* g_out, float* g_a,
                           float* g_b, float* g_scale, float* g_bias, int64_t rows, int D, int tanh_mode,
                           cudaStream_t s);
int attention_value_bwd(const float* qkv, const float* g_o, float* g_qkv, int64_t B, NetDims d, cudaStream_t s);
int features_dense0_bwd(const float* x, const float* g_h, float* g_W0, int64_t B, NetDims d, cudaStream_t s);
int colsum_add(const float* g, float* out, int64_t M, int N, int64_t ld, cudaStream_t s);
int add_inplace(float* dst, const float* src, int64_t n, cudaStream_t s);

}  // namespace dh
