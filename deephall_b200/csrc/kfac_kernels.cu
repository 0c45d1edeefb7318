// KFAC update from the moving-average curvature statistics (optimizers/kfac.py:202-219 hands this to kfac_jax; the
// rule is restated in oracle/kfac.py): for every dense block the pi-adjusted, damped Kronecker factors
//     A~ = A / tr_avg(A) + d I,   G~ = G / tr_avg(G) + d I,   d = sqrt((damping / npw) / (tr_avg(A) tr_avg(G)))
// (kfac_jax.utils.pi_adjusted_kronecker_inverse with average-trace norms), their inverses (spd_inverse_batched), and
//     U = A~^-1 V G~^-1 / (tr_avg(A) tr_avg(G) npw);
// LayerNorm / Jastrow parameters: diagonal blocks, U = g / (F + damping).
// Everything that is not an inverse or a contraction is FOUR launches here (coefficients, factor assembly, gradient
// gather, update scatter) instead of ~20 small tensor ops per block.
#include "kernels.h"

namespace dh {

namespace {

__device__ __forceinline__ float block_sum256(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) { t = warp_sum(t); if (threadIdx.x == 0) red[0] = t; }
  __syncthreads();
  const float out = red[0];
  __syncthreads();
  return out;
}

// coef[blk] = {tr_avg(A), tr_avg(G), d, c_k, ok}: one block of 256 threads per dense curvature block
__global__ void __launch_bounds__(256)
kfac_coef_kernel(const KfBlkDesc* __restrict__ bd, const float* __restrict__ stats, const float* __restrict__ xtx0, float w,
                 float damping, float* __restrict__ coef) {
  __shared__ float red[8];
  const KfBlkDesc b = bd[blockIdx.x];
  const float* xtx = b.xtx >= 0 ? stats + b.xtx : xtx0;
  float ta = 0.f, tg = 0.f;
  for (int i = threadIdx.x; i < b.din; i += 256) ta += xtx[(int64_t)i * b.din + i] / w;
  for (int i = threadIdx.x; i < b.dout; i += 256) tg += stats[b.gtg + (int64_t)i * b.dout + i] / w;
  ta = block_sum256(ta, red);
  tg = block_sum256(tg, red);
  if (threadIdx.x == 0) {
    if (b.hb) ta += 1.f;
    const float ca = ta / (float)(b.din + b.hb), cg = tg / (float)b.dout;
    const float c = ca * cg, dn = damping / (float)b.npw;
    const bool ok = c > 0.f;  // a factor that is still zero: plain damping
    float* o = coef + 8 * blockIdx.x;
    o[0] = ok ? ca : 1.f;
    o[1] = ok ? cg : 1.f;
    o[2] = ok ? sqrtf(dn / c) : 1.f;
    o[3] = ok ? sqrtf(c) : sqrtf(dn);
    o[4] = ok ? 1.f : 0.f;
  }
}

// one damped factor per blockIdx.x, embedded as diag(M~, I) in the padded size of its batch
__global__ void __launch_bounds__(256)
kfac_build_kernel(const KfMatDesc* __restrict__ md, const float* __restrict__ stats, const float* __restrict__ xtx0, float w,
                  const float* __restrict__ coef, float* __restrict__ batch_s, float* __restrict__ batch_l) {
  const KfMatDesc m = md[blockIdx.x];
  float* dst = (m.cls ? batch_l : batch_s) + m.dst;
  const float* cf = coef + 8 * m.blk;
  const float c = m.is_g ? cf[1] : cf[0], d_hat = cf[2];
  const bool ok = cf[4] != 0.f;
  const float* src = m.src >= 0 ? stats + m.src : xtx0;
  const int dim = m.dim, n = m.n, nc = m.n_core;
  for (int idx = blockIdx.y * 256 + threadIdx.x; idx < dim * dim; idx += gridDim.y * 256) {
    const int i = idx / dim, j = idx - i * dim;
    float v;
    if (i < n && j < n) {
      if (i < nc && j < nc) v = src[(int64_t)i * nc + j] / w;
      else if (i == nc && j == nc) v = 1.f;               // the bias input is the constant 1
      else v = stats[m.xsum + (i == nc ? j : i)] / w;
      v = ok ? v / c : 0.f;
      if (i == j) v += d_hat;
    } else {
      v = i == j ? 1.f : 0.f;
    }
    dst[idx] = v;
  }
}

// V~ = [kernel gradient ; bias gradient] of every dense block, contiguous
__global__ void __launch_bounds__(256)
kfac_gather_kernel(const KfBlkDesc* __restrict__ bd, const float* __restrict__ grads, float* __restrict__ V) {
  const KfBlkDesc b = bd[blockIdx.x];
  const int64_t nk = (int64_t)b.din * b.dout, tot = nk + (b.hb ? b.dout : 0);
  for (int64_t t = (int64_t)blockIdx.y * 256 + threadIdx.x; t < tot; t += (int64_t)gridDim.y * 256)
    V[b.v_off + t] = t < nk ? grads[b.ko + t] : grads[b.bo + (t - nk)];
}

__global__ void __launch_bounds__(256)
kfac_scatter_kernel(const KfBlkDesc* __restrict__ bd, const float* __restrict__ U, const float* __restrict__ coef,
                    float* __restrict__ out) {
  const KfBlkDesc b = bd[blockIdx.x];
  const float ck = coef[8 * blockIdx.x + 3];
  const float den = ck * ck * (float)b.npw;
  const int64_t nk = (int64_t)b.din * b.dout, tot = nk + (b.hb ? b.dout : 0);
  for (int64_t t = (int64_t)blockIdx.y * 256 + threadIdx.x; t < tot; t += (int64_t)gridDim.y * 256) {
    const float v = U[b.v_off + t] / den;
    if (t < nk) out[b.ko + t] = v;
    else out[b.bo + (t - nk)] = v;
  }
}

__global__ void __launch_bounds__(256)
kfac_diag_kernel(const KfDiagDesc* __restrict__ dd, const float* __restrict__ stats, float w, float damping,
                 const float* __restrict__ grads, float* __restrict__ out) {
  const KfDiagDesc d = dd[blockIdx.x];
  for (int64_t t = threadIdx.x; t < d.n; t += 256) out[d.ko + t] = grads[d.ko + t] / (stats[d.o + t] / w + damping);
}

// Both products of U = A~^-1 V~ G~^-1 for ALL dense blocks in one launch each (grid.z = block): a plain fp32 FMA contraction,
// 64 x 64 tiles, 4 x 4 outputs per thread.  stage 0: T = A~^-1 V~ ([na x na] . [na x dout]); stage 1: U = T G~^-1.
// (As 30 separate launches of 6-12 CTAs each these products took 0.9 ms of the update; the contractions are 1.8 GFLOP.)
constexpr int KG_T = 64, KG_K = 16;
__global__ void __launch_bounds__(256)
kfac_grouped_gemm_kernel(const KfBlkDesc* __restrict__ bd, const KfMatDesc* __restrict__ md, const float* __restrict__ inv_s,
                         const float* __restrict__ inv_l, const float* __restrict__ X, float* __restrict__ Y, int stage) {
  __shared__ float As[KG_K][KG_T + 4], Bs[KG_K][KG_T + 4];
  const KfBlkDesc b = bd[blockIdx.z];
  const KfMatDesc mat = md[2 * blockIdx.z + stage];  // stage 0: the A factor, stage 1: the G factor
  const float* F = (mat.cls ? inv_l : inv_s) + mat.dst;
  const int na = b.din + b.hb, M = na, N = b.dout, Kd = stage == 0 ? na : b.dout;
  const int m0 = blockIdx.y * KG_T, n0 = blockIdx.x * KG_T;
  if (m0 >= M || n0 >= N) return;
  // stage 0: A = F (ld = dim), B = X_b (ld = dout);  stage 1: A = X_b (ld = dout), B = F (ld = dim)
  const float* A = stage == 0 ? F : X + b.v_off;
  const float* B = stage == 0 ? X + b.v_off : F;
  const int lda = stage == 0 ? mat.dim : b.dout, ldb = stage == 0 ? b.dout : mat.dim;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < Kd; k0 += KG_K) {
    for (int t = threadIdx.x; t < KG_T * KG_K; t += 256) {
      const int i = t / KG_K, k = t - i * KG_K;  // A tile: rows m0 + i, columns k0 + k (k fastest: contiguous reads)
      As[k][i] = (m0 + i < M && k0 + k < Kd) ? A[(int64_t)(m0 + i) * lda + k0 + k] : 0.f;
      const int kb = t / KG_T, j = t - kb * KG_T;  // B tile: rows k0 + kb, columns n0 + j
      Bs[kb][j] = (k0 + kb < Kd && n0 + j < N) ? B[(int64_t)(k0 + kb) * ldb + n0 + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KG_K; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][4 * ty]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][4 * tx]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* C = Y + b.v_off;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + 4 * ty + i, n = n0 + 4 * tx + j;
      if (m < M && n < N) C[(int64_t)m * N + n] = acc[i][j];
    }
}

}  // namespace

// Y_b = A~_b^-1 X_b (stage 0) or X_b G~_b^-1 (stage 1) for every dense block; max_rows / max_cols bound the tile grid
int kfac_grouped_gemm(const KfBlkDesc* bd, const KfMatDesc* md, int nblk, int max_rows, int max_cols, const float* inv_s,
                      const float* inv_l, const float* X, float* Y, int stage, cudaStream_t s) {
  if (nblk <= 0) return 0;
  dim3 grid((unsigned)((max_cols + KG_T - 1) / KG_T), (unsigned)((max_rows + KG_T - 1) / KG_T), (unsigned)nblk);
  kfac_grouped_gemm_kernel<<<grid, 256, 0, s>>>(bd, md, inv_s, inv_l, X, Y, stage);
  return (int)cudaGetLastError();
}

int kfac_damped_factors(const KfBlkDesc* bd, int nblk, const KfMatDesc* md, int nmat, const float* stats, const float* xtx0,
                        float weight, float damping, float* coef, float* batch_s, float* batch_l, cudaStream_t s) {
  if (nblk <= 0) return 0;
  kfac_coef_kernel<<<nblk, 256, 0, s>>>(bd, stats, xtx0, weight, damping, coef);
  kfac_build_kernel<<<dim3(nmat, 16), 256, 0, s>>>(md, stats, xtx0, weight, coef, batch_s, batch_l);
  return (int)cudaGetLastError();
}

int kfac_gather(const KfBlkDesc* bd, int nblk, const float* grads, float* V, cudaStream_t s) {
  if (nblk <= 0) return 0;
  kfac_gather_kernel<<<dim3(nblk, 16), 256, 0, s>>>(bd, grads, V);
  return (int)cudaGetLastError();
}

int kfac_scatter(const KfBlkDesc* bd, int nblk, const KfDiagDesc* dd, int ndiag, const float* U, const float* coef,
                 const float* stats, float weight, float damping, const float* grads, float* out, cudaStream_t s) {
  if (nblk > 0) kfac_scatter_kernel<<<dim3(nblk, 16), 256, 0, s>>>(bd, U, coef, out);
  if (ndiag > 0) kfac_diag_kernel<<<ndiag, 256, 0, s>>>(dd, stats, weight, damping, grads, out);
  return (int)cudaGetLastError();
}

}  // namespace dh
