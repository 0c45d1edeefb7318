// Prologue of a value-only pass (log psi; every move of the Metropolis sweep) as ONE launch: everything that depends on the
// electron coordinates alone.
//   * the Metropolis proposal x' = sph_sampling(x) (mcmc.py:67-102) when the pass evaluates a move -- the "proposal ->
//     forward prologue" fusion: the proposal never makes a round trip through its own launch;
//   * the input features (psiformer.py:51-60) and the two linear maps of them: Dense_0 (h) and the first layer's q|k|v
//     (W0 . Wqkv, folded at prepare time);
//   * the per-electron envelope table of the fused orbital epilogue (gemm_tc.cu, ORB): [L] complex envelope values times
//     the weight un-scale factor + [N K] complex bias products.
// The feature maps are store-bound (4 KB per electron); the trigonometry, the fp64 envelope magnitudes and the bias products
// of the other two parts run while those stores drain.
//
// Block = 256 threads = VP_ROWS (64) consecutive (walker, electron) rows.
//   phase A  thread t < 64: row t's scalar work (proposal, sincos of the angles, sincos of theta / 2 in double)
//   phase B  warp w: rows 8 w .. 8 w + 7 of both feature maps, every 16-byte group of weight columns loaded once for 8 rows
//   phase C  items (row, m): envelope value -> shared memory + table
//   phase D  items (row, column): bias products -> table
#include "kernels.h"

namespace dh {

namespace {

constexpr int VP_ROWS = 64;

__device__ __forceinline__ void feature_map_rows(const float (&f)[8][4], int nrow, const float* __restrict__ W,
                                                 const float* __restrict__ bias, float* __restrict__ o, int Nout, int lane) {
  if ((Nout & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0) {
    for (int d = 4 * lane; d < Nout; d += 128) {  // 16-byte accesses
      const float4 w0 = *reinterpret_cast<const float4*>(W + d), w1 = *reinterpret_cast<const float4*>(W + Nout + d);
      const float4 w2 = *reinterpret_cast<const float4*>(W + 2 * Nout + d), w3 = *reinterpret_cast<const float4*>(W + 3 * Nout + d);
      const float4 bb = bias != nullptr ? *reinterpret_cast<const float4*>(bias + d) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r >= nrow) break;
        float4 v;
        v.x = fmaf(f[r][0], w0.x, fmaf(f[r][1], w1.x, fmaf(f[r][2], w2.x, f[r][3] * w3.x))) + bb.x;
        v.y = fmaf(f[r][0], w0.y, fmaf(f[r][1], w1.y, fmaf(f[r][2], w2.y, f[r][3] * w3.y))) + bb.y;
        v.z = fmaf(f[r][0], w0.z, fmaf(f[r][1], w1.z, fmaf(f[r][2], w2.z, f[r][3] * w3.z))) + bb.z;
        v.w = fmaf(f[r][0], w0.w, fmaf(f[r][1], w1.w, fmaf(f[r][2], w2.w, f[r][3] * w3.w))) + bb.w;
        *reinterpret_cast<float4*>(o + (int64_t)r * Nout + d) = v;
      }
    }
    return;
  }
  for (int d = lane; d < Nout; d += 32) {
    const float w0 = W[d], w1 = W[Nout + d], w2 = W[2 * Nout + d], w3 = W[3 * Nout + d];
    const float bb = bias != nullptr ? bias[d] : 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r >= nrow) break;
      o[(int64_t)r * Nout + d] = fmaf(f[r][0], w0, fmaf(f[r][1], w1, fmaf(f[r][2], w2, f[r][3] * w3))) + bb;
    }
  }
}

__global__ void __launch_bounds__(256)
value_prologue_kernel(ValuePrologue a, int64_t rows, int N, int n_up) {
  extern __shared__ __align__(16) unsigned char vp_smem[];
  __shared__ float feat[VP_ROWS][4];
  __shared__ double half_ang[VP_ROWS][2];  // cos, sin of theta / 2
  __shared__ float phis[VP_ROWS];
  cplx* env = reinterpret_cast<cplx*>(vp_smem);  // [VP_ROWS][L]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * VP_ROWS;
  const int nvalid = rows - row0 < VP_ROWS ? (int)(rows - row0) : VP_ROWS;
  if (t < nvalid) {
    const int64_t row = row0 + t;
    float theta = a.x[row * 2], phi = a.x[row * 2 + 1];
    if (a.x_new != nullptr) {
      const int64_t b = row / N;
      float nrm, uph, th2, ph2;
      propose_draws(a.dv->seed, a.dv->offset, a.dv->subseq0 + (uint64_t)(a.walker0 + b), (int)(row % N), nrm, uph);
      propose_point(theta, phi, nrm, uph, a.dv->width, th2, ph2);
      a.x_new[row * 2] = th2;
      a.x_new[row * 2 + 1] = ph2;
      theta = th2;
      phi = ph2;
    }
    float st, ct, sp, cp;
    sincosf(theta, &st, &ct);
    sincosf(phi, &sp, &cp);
    feat[t][0] = ct; feat[t][1] = st * cp; feat[t][2] = st * sp; feat[t][3] = ((int)(row % N) < n_up) ? 1.f : -1.f;  // (z, x, y, spin)
    if (a.tab != nullptr) {
      double sh, ch;
      sincos(0.5 * (double)theta, &sh, &ch);
      half_ang[t][0] = ch; half_ang[t][1] = sh;
      phis[t] = phi;
    }
  }
  __syncthreads();
  // ---- phase B: the feature maps
  if (8 * warp < nvalid) {
    float f[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) f[r][k] = feat[8 * warp + r][k];
    const int nrow = nvalid - 8 * warp < 8 ? nvalid - 8 * warp : 8;
    const int64_t r0 = row0 + 8 * warp;
    feature_map_rows(f, nrow, a.W0, nullptr, a.h + r0 * a.n0, a.n0, lane);
    if (a.W1 != nullptr) feature_map_rows(f, nrow, a.W1, a.b1, a.q + r0 * a.n1, a.n1, lane);
  }
  if (a.tab == nullptr) return;
  // ---- phase C: envelope values (the arithmetic of orbital_value_kernel: magnitudes by repeated squaring in double, the
  // phase angle reduced in double and evaluated in fp32)
  const int L = a.L, NK = a.NK, twoQ = a.twoQ;
  const float us = a.unscale ? __ldg(a.unscale) : 1.f;
  const int stride = 2 * (L + NK);
  for (int it = t; it < nvalid * L; it += 256) {
    const int e = it / L, m = it - e * L;
    const double mag = a.normfac[m] * dpow_int(half_ang[e][0], m) * dpow_int(half_ang[e][1], twoQ - m);
    double psi = (double)(2 * m - twoQ) * 0.5 * (double)phis[e];
    psi -= 6.283185307179586476925287 * rint(psi * 0.15915494309189533576888);
    float sp, cp;
    sincosf((float)psi, &sp, &cp);
    const cplx ev = make_float2((float)(mag * (double)cp), (float)(mag * (double)sp));
    env[it] = ev;
    *reinterpret_cast<float2*>(a.tab + (row0 + e) * stride + 2 * m) = make_float2(ev.x * us, ev.y * us);
  }
  __syncthreads();
  // ---- phase D: bias products sum_m b(m, j) env[m]
  for (int it = t; it < nvalid * NK; it += 256) {
    const int e = it / NK, j = it - e * NK;
    const cplx* ee = env + e * L;
    cplx acc = cmake(0.f, 0.f);
    for (int m = 0; m < L; ++m) acc = cfma(cmake(__ldg(a.bre + m * NK + j), __ldg(a.bim + m * NK + j)), ee[m], acc);
    *reinterpret_cast<float2*>(a.tab + (row0 + e) * stride + 2 * L + 2 * j) = acc;
  }
}

}  // namespace

int value_prologue(const ValuePrologue& a, int64_t rows, int N, int n_up, cudaStream_t s) {
  if (rows <= 0) return 0;
  if (!a.x || !a.W0 || !a.h || (a.x_new && !a.dv) || (a.W1 && !a.q) || (a.tab && (a.L < 1 || a.L > 96 || !a.normfac || !a.bre || !a.bim)))
    return -2;
  const size_t smem = a.tab ? (size_t)VP_ROWS * a.L * sizeof(cplx) : 0;  // <= 48 KB for L <= 96
  value_prologue_kernel<<<(unsigned)((rows + VP_ROWS - 1) / VP_ROWS), 256, smem, s>>>(a, rows, N, n_up);
  return (int)cudaGetLastError();
}

}  // namespace dh
