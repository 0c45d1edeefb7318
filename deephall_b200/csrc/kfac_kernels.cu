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

}  // namespace

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
