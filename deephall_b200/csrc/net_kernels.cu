// Psiformer body kernels other than the dense contractions: input features + Dense_0,
// residual + (tanh) + LayerNorm, self-attention -- each in a value-only (R = 1) and a
// forward-Laplacian "jet" form (R = 2N + 8 rows per electron, see common.cuh / oracle/jets.py).
//
// Reference semantics: networks/psiformer.py:37-60 and flax 0.10.2 MultiHeadAttention /
// LayerNorm(epsilon=1e-5).
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace dh {

// =============================================================================================
// features (psiformer.py:51-60) fused with Dense_0 (psiformer.py:42, no bias)
// grid: B*N blocks; thread d loops over model columns.
// =============================================================================================
// Generalised to any linear map of the features: out[row, 0..Nout) = feat_row @ W[4][Nout]
// (+ bias on the value row) -- also used for the FIRST layer's q|k|v, whose input h = feat @ W0
// is linear in the features, so q|k|v = feat @ (W0 Wqkv) + b needs no 256-deep contraction.
__global__ void features_dense0_kernel(const float* __restrict__ x, const float* __restrict__ W0,
                                       const float* __restrict__ bias, float* __restrict__ h, int Nout, NetDims dm,
                                       int compressed) {
  extern __shared__ __align__(16) float feat[];  // [R][4]
  const int N = dm.N, R = dm.R, D = Nout;
  const int64_t bi = blockIdx.x;
  const int i = (int)(bi % N);
  for (int t = threadIdx.x; t < R * 4; t += blockDim.x) feat[t] = 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float th = x[bi * 2], ph = x[bi * 2 + 1];
    float st, ct, sp, cp;
    sincosf(th, &st, &ct);
    sincosf(ph, &sp, &cp);
    float rx = st * cp, ry = st * sp, rz = ct;
    float spin = i < dm.n_up ? 1.f : -1.f;
    // feature order: (z, x, y, spin)
    feat[0] = rz; feat[1] = rx; feat[2] = ry; feat[3] = spin;
    if (R > 1) {
      Rows rw(N, true);
      float* f;
      f = feat + rw.J(2 * i) * 4;      // theta_hat x r = -phi_hat = (sp, -cp, 0)
      f[0] = 0.f; f[1] = sp; f[2] = -cp;
      f = feat + rw.J(2 * i + 1) * 4;  // phi_hat x r = theta_hat = (ct cp, ct sp, -st)
      f[0] = -st; f[1] = ct * cp; f[2] = ct * sp;
      f = feat + rw.S() * 4;           // -2 r
      f[0] = -2.f * rz; f[1] = -2.f * rx; f[2] = -2.f * ry;
      f = feat + rw.D(0) * 4; f[0] = ry;  f[1] = 0.f;  f[2] = -rz;   // e_x x r = (0,-rz,ry)
      f = feat + rw.D(1) * 4; f[0] = -rx; f[1] = rz;   f[2] = 0.f;   // e_y x r = (rz,0,-rx)
      f = feat + rw.D(2) * 4; f[0] = 0.f; f[1] = -ry;  f[2] = rx;    // e_z x r = (-ry,rx,0)
      f = feat + rw.T(0) * 4; f[0] = -rz; f[1] = 0.f;  f[2] = -ry;   // e_x r_x - r
      f = feat + rw.T(1) * 4; f[0] = -rz; f[1] = -rx;  f[2] = 0.f;
      f = feat + rw.T(2) * 4; f[0] = 0.f; f[1] = -rx;  f[2] = -ry;
    }
  }
  __syncthreads();
  const int Rout = compressed ? 10 : R;
  float* out = h + bi * Rout * D;
  Rows rwc(N, true);
  if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(W0) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0) {
    // 16-byte form: thread = four columns, all rows (a quarter of the instructions per output of the scalar form)
    for (int d = 4 * threadIdx.x; d < D; d += 4 * blockDim.x) {
      const float4 w0 = *reinterpret_cast<const float4*>(W0 + d), w1 = *reinterpret_cast<const float4*>(W0 + D + d);
      const float4 w2 = *reinterpret_cast<const float4*>(W0 + 2 * D + d), w3 = *reinterpret_cast<const float4*>(W0 + 3 * D + d);
      const float4 bb = bias != nullptr ? *reinterpret_cast<const float4*>(bias + d) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int ro = 0; ro < Rout; ++ro) {
        const int r = !compressed ? ro : (ro == 0 ? 0 : (ro <= 2 ? rwc.J(2 * i + ro - 1) : rwc.S() + (ro - 3)));
        const float4 f = *reinterpret_cast<const float4*>(feat + r * 4);
        float4 v;
        v.x = fmaf(f.x, w0.x, fmaf(f.y, w1.x, fmaf(f.z, w2.x, f.w * w3.x)));
        v.y = fmaf(f.x, w0.y, fmaf(f.y, w1.y, fmaf(f.z, w2.y, f.w * w3.y)));
        v.z = fmaf(f.x, w0.z, fmaf(f.y, w1.z, fmaf(f.z, w2.z, f.w * w3.z)));
        v.w = fmaf(f.x, w0.w, fmaf(f.y, w1.w, fmaf(f.z, w2.w, f.w * w3.w)));
        if (r == 0) { v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w; }
        *reinterpret_cast<float4*>(out + (int64_t)ro * D + d) = v;
      }
    }
    return;
  }
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float w0 = W0[d], w1 = W0[D + d], w2 = W0[2 * D + d], w3 = W0[3 * D + d];
    for (int ro = 0; ro < Rout; ++ro) {
      // compressed: value | own tangent flows | S | D_a | T_a  (every other row of the full layout is zero)
      const int r = !compressed ? ro : (ro == 0 ? 0 : (ro <= 2 ? rwc.J(2 * i + ro - 1) : rwc.S() + (ro - 3)));
      const float* f = feat + r * 4;
      float v = fmaf(f[0], w0, fmaf(f[1], w1, fmaf(f[2], w2, f[3] * w3)));
      if (r == 0 && bias != nullptr) v += bias[d];
      out[(int64_t)ro * D + d] = v;
    }
  }
}

int features_dense0(const float* x, const float* W0, float* h, int64_t B, NetDims d, cudaStream_t s) {
  return features_linear(x, W0, nullptr, h, d.D, B, d, 0, s);
}

// value-only form (R = 1): one warp per FV_ROWS consecutive (walker, electron) rows, eight warps per block.  Lane r
// evaluates the features of row r, the warp shares them by shuffles, and every 16-byte group of weight columns is
// loaded once for all FV_ROWS rows (one warp per row re-read the weights from L1 for each row: L1 was 84 % busy).
constexpr int FV_ROWS = 8;

__global__ void __launch_bounds__(256)
features_value_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                      float* __restrict__ out, int Nout, int64_t rows, NetDims dm) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * FV_ROWS;
  if (row0 >= rows) return;
  float fl[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane < FV_ROWS && row0 + lane < rows) {
    const int64_t row = row0 + lane;
    float st, ct, sp, cp;
    sincosf(x[row * 2], &st, &ct);
    sincosf(x[row * 2 + 1], &sp, &cp);
    fl[0] = ct; fl[1] = st * cp; fl[2] = st * sp; fl[3] = ((int)(row % dm.N) < dm.n_up) ? 1.f : -1.f;  // (z, x, y, spin)
  }
  float f[FV_ROWS][4];
#pragma unroll
  for (int r = 0; r < FV_ROWS; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) f[r][k] = __shfl_sync(0xffffffffu, fl[k], r);
  const int nrow = rows - row0 < FV_ROWS ? (int)(rows - row0) : FV_ROWS;
  float* o = out + row0 * Nout;
  if ((Nout & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0) {
    for (int d = 4 * lane; d < Nout; d += 128) {  // 16-byte accesses
      const float4 w0 = *reinterpret_cast<const float4*>(W + d), w1 = *reinterpret_cast<const float4*>(W + Nout + d);
      const float4 w2 = *reinterpret_cast<const float4*>(W + 2 * Nout + d), w3 = *reinterpret_cast<const float4*>(W + 3 * Nout + d);
      const float4 bb = bias != nullptr ? *reinterpret_cast<const float4*>(bias + d) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < FV_ROWS; ++r) {
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
    for (int r = 0; r < FV_ROWS; ++r) {
      if (r >= nrow) break;
      o[(int64_t)r * Nout + d] = fmaf(f[r][0], w0, fmaf(f[r][1], w1, fmaf(f[r][2], w2, f[r][3] * w3))) + bb;
    }
  }
}

int features_linear(const float* x, const float* W, const float* bias, float* out, int Nout, int64_t B, NetDims d,
                    int compressed, cudaStream_t s) {
  if (d.R == 1) {
    const int64_t rows = B * d.N;
    features_value_kernel<<<(unsigned)((rows + 8 * FV_ROWS - 1) / (8 * FV_ROWS)), 256, 0, s>>>(x, W, bias, out, Nout, rows, d);
    return (int)cudaGetLastError();
  }
  int threads = Nout >= 256 ? 256 : ((Nout + 31) / 32 * 32);
  if ((Nout & 3) == 0) threads = ((Nout / 4 + 31) / 32 * 32) < 256 ? ((Nout / 4 + 31) / 32 * 32) : 256;  // one float4 column group per thread
  features_dense0_kernel<<<(unsigned)(B * d.N), threads, d.R * 4 * sizeof(float), s>>>(x, W, bias, out, Nout, d, compressed);
  return (int)cudaGetLastError();
}

// =============================================================================================
// out = LayerNorm(a + b) or LayerNorm(a + tanh(b)), with jets.  One warp per (walker, electron);
// lane l owns columns l, l+32, ...  (D <= 256, D % 32 == 0).
// =============================================================================================
#ifndef DH_TANH
#define DH_TANH tanhf
#endif
constexpr int LN_VPL = 8;

// tanh with absolute error < 2e-7 (odd polynomial below 0.3, 1 - 2/(e^{2|x|}+1) with the fast exponential above)
__device__ __forceinline__ float tanh_fast(float x) {
  const float ax = fabsf(x);
  const float x2 = x * x;
  const float p = x * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 0.021869488536155203f, -0.053968253968253971f), 0.13333333333333333f),
                               -0.33333333333333331f), 1.0f);
  const float e = __expf(2.0f * ax);
  const float t = copysignf(1.0f - __fdividef(2.0f, e + 1.0f), x);
  return ax < 0.3f ? p : t;
}

// VEC4 (D = 256, 16-byte aligned tensors): lane l owns columns 4l..4l+3 and 128+4l..128+4l+3, moved as two
// float4 per row (512-byte requests); otherwise lane l owns columns l, l+32, ...
// (at most 80 registers: a block of this kernel fits next to a persistent contraction CTA of the other chunk stream,
// 576 threads x 96 registers, see gemm_tc.cu)
template <bool TANH, bool VEC4>
__global__ void __launch_bounds__(128, 6)
residual_layernorm_kernel(const float* __restrict__ a, const float* __restrict__ b,
                          const float* __restrict__ scale, const float* __restrict__ bias,
                          float* __restrict__ out, int64_t groups, NetDims dm, int a_comp, int a_pl, int o_pl) {
  const int lane = threadIdx.x & 31;
  const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= groups) return;
  const int R = dm.R, D = dm.D, N = dm.N;
  const int vpl = VEC4 ? LN_VPL : (D >> 5);
  const float invD = 1.0f / (float)D;
  // a_comp: `a` is the first layer's Dense_0 output in compressed form, 10 rows per electron
  // (value | own tangent flows | S | D_a | T_a); every other jet row of it is zero
  const int iel = (int)(g % N);
  const float* ga = a + g * (a_comp ? 10 : R) * D;
  const float* gb = b + g * R * D;
  float* go = out + g * R * D;
  // a_pl / o_pl (VEC4 only): `a` / `out` are fp16 hi / lo planes (common.cuh), the form the tcgen05 contraction
  // takes its left operand in
  const int64_t plane = groups * R * D;
  const __half* pa = reinterpret_cast<const __half*>(a) + g * R * D;
  __half* po = reinterpret_cast<__half*>(out) + g * R * D;

  // one row (D floats at `base`) <-> the LN_VPL values this lane owns
  auto ldrow = [&](const float* base, float (&dst)[LN_VPL]) {
    if (VEC4) {
      const float4 p = *reinterpret_cast<const float4*>(base + 4 * lane);
      const float4 q = *reinterpret_cast<const float4*>(base + 128 + 4 * lane);
      dst[0] = p.x; dst[1] = p.y; dst[2] = p.z; dst[3] = p.w;
      dst[4] = q.x; dst[5] = q.y; dst[6] = q.z; dst[7] = q.w;
    } else {
#pragma unroll
      for (int v = 0; v < LN_VPL; ++v) dst[v] = v < vpl ? base[lane + 32 * v] : 0.f;
    }
  };
  auto strow = [&](float* base, const float (&src)[LN_VPL]) {
    if (VEC4) {
      *reinterpret_cast<float4*>(base + 4 * lane) = make_float4(src[0], src[1], src[2], src[3]);
      *reinterpret_cast<float4*>(base + 128 + 4 * lane) = make_float4(src[4], src[5], src[6], src[7]);
    } else {
#pragma unroll
      for (int v = 0; v < LN_VPL; ++v)
        if (v < vpl) base[lane + 32 * v] = src[v];
    }
  };

  auto sto = [&](int r, const float (&src)[LN_VPL]) {  // row r of `out`
    if (VEC4 && o_pl) {
      st_planes4(po, plane, (int64_t)r * D + 4 * lane, make_float4(src[0], src[1], src[2], src[3]));
      st_planes4(po, plane, (int64_t)r * D + 128 + 4 * lane, make_float4(src[4], src[5], src[6], src[7]));
    } else {
      strow(go + (int64_t)r * D, src);
    }
  };
  auto lda = [&](int r, float (&dst)[LN_VPL]) {  // row r of `a`
    if (VEC4 && a_pl) {
      const float4 p = ld_planes4(pa, plane, (int64_t)r * D + 4 * lane);
      const float4 q = ld_planes4(pa, plane, (int64_t)r * D + 128 + 4 * lane);
      dst[0] = p.x; dst[1] = p.y; dst[2] = p.z; dst[3] = p.w;
      dst[4] = q.x; dst[5] = q.y; dst[6] = q.z; dst[7] = q.w;
      return;
    }
    if (!a_comp) { ldrow(ga + (int64_t)r * D, dst); return; }
    int rc;
    if (r == 0) rc = 0;
    else if (r <= 2 * N) rc = ((r - 1) >> 1) == iel ? 1 + ((r - 1) & 1) : -1;
    else rc = 3 + (r - (2 * N + 1));
    if (rc < 0) {
#pragma unroll
      for (int v = 0; v < LN_VPL; ++v) dst[v] = 0.f;
    } else {
      ldrow(ga + (int64_t)rc * D, dst);
    }
  };

  float gam[LN_VPL], bet[LN_VPL];
  float c0[LN_VPL], t1[LN_VPL], t2[LN_VPL];  // centred value row; tanh' and tanh''
  ldrow(scale, gam);
  ldrow(bias, bet);
#pragma unroll
  for (int v = 0; v < LN_VPL; ++v) { t1[v] = 1.f; t2[v] = 0.f; }
  // ---- value row
  float xr[LN_VPL], ra[LN_VPL], rb[LN_VPL], ro[LN_VPL];
  lda(0, ra);
  ldrow(gb, rb);
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < LN_VPL; ++v) {
    xr[v] = 0.f;
    if (v < vpl) {
      float bv = rb[v];
      if (TANH) {
        float t = DH_TANH(bv);
        t1[v] = 1.f - t * t;
        t2[v] = -2.f * t * t1[v];
        bv = t;
      }
      xr[v] = ra[v] + bv;
      sum += xr[v];
    }
  }
  float mu = warp_sum(sum) * invD;
  float sq = 0.f;
#pragma unroll
  for (int v = 0; v < LN_VPL; ++v) {
    c0[v] = (v < vpl) ? xr[v] - mu : 0.f;
    sq += c0[v] * c0[v];
  }
  const float var = warp_sum(sq) * invD + 1e-5f;
  const float rho0 = rsqrtf(var);
  const float rho1 = -0.5f * rho0 / var;          // d rho / d var
  const float rho2 = 0.75f * rho0 / (var * var);  // d2 rho / d var2
#pragma unroll
  for (int v = 0; v < LN_VPL; ++v) ro[v] = fmaf(c0[v] * rho0, gam[v], bet[v]);
  sto(0, ro);
  if (R == 1) return;

  Rows rw(N, true);
  float accS[LN_VPL], bsq[LN_VPL];
#pragma unroll
  for (int v = 0; v < LN_VPL; ++v) { accS[v] = 0.f; bsq[v] = 0.f; }
  float sum_cc = 0.f, sum_vv = 0.f;  // sum_k mean(cJk^2), sum_k vJk^2

  // first-order row helper: loads x_r, centres it, returns mean(c0*c_r), mean(c_r^2)
  auto load_first = [&](int r, float (&cr)[LN_VPL], float (&braw)[LN_VPL], float& m_c0c, float& m_cc) {
    float s1 = 0.f;
    float av[LN_VPL];
    lda(r, av);
    ldrow(gb + (int64_t)r * D, braw);
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) {
      cr[v] = 0.f;
      if (v < vpl) {
        cr[v] = av[v] + (TANH ? t1[v] * braw[v] : braw[v]);
        s1 += cr[v];
      }
    }
    float m = warp_sum(s1) * invD;
    float d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) {
      if (v < vpl) cr[v] -= m;
      d1 += c0[v] * cr[v];
      d2 += cr[v] * cr[v];
    }
    m_c0c = warp_sum(d1) * invD;
    m_cc = warp_sum(d2) * invD;
  };
  // second-order row helper: x_r = a_r + t1*b_r + t2*extra ; returns centred row and mean(c0*c_r)
  auto load_second = [&](int r, const float (&extra)[LN_VPL], float (&cr)[LN_VPL], float& m_c0c) {
    float s1 = 0.f;
    float av[LN_VPL], bv[LN_VPL];
    lda(r, av);
    ldrow(gb + (int64_t)r * D, bv);
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) {
      cr[v] = 0.f;
      if (v < vpl) {
        cr[v] = av[v] + (TANH ? fmaf(t1[v], bv[v], t2[v] * extra[v]) : bv[v]);
        s1 += cr[v];
      }
    }
    float m = warp_sum(s1) * invD;
    float d1 = 0.f;
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) {
      if (v < vpl) cr[v] -= m;
      d1 += c0[v] * cr[v];
    }
    m_c0c = warp_sum(d1) * invD;
  };

  float cr[LN_VPL], braw[LN_VPL];
  // ---- J rows
  for (int k = 0; k < 2 * N; ++k) {
    const int r = rw.J(k);
    float m_c0c, m_cc;
    load_first(r, cr, braw, m_c0c, m_cc);
    const float vJ = 2.f * m_c0c;
    const float rhoJ = rho1 * vJ;
    sum_cc += m_cc;
    sum_vv += vJ * vJ;
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) {
      ro[v] = gam[v] * fmaf(cr[v], rho0, c0[v] * rhoJ);
      accS[v] = fmaf(cr[v], rhoJ, accS[v]);
      if (TANH) bsq[v] = fmaf(braw[v], braw[v], bsq[v]);
    }
    sto(r, ro);
  }
  // ---- S row
  {
    const int r = rw.S();
    float m_c0c;
    load_second(r, bsq, cr, m_c0c);
    const float vS = 2.f * m_c0c + 2.f * sum_cc;
    const float rhoS = rho1 * vS + rho2 * sum_vv;
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) ro[v] = gam[v] * (fmaf(cr[v], rho0, c0[v] * rhoS) + 2.f * accS[v]);
    sto(r, ro);
  }
  // ---- D_a / T_a rows
  for (int a3 = 0; a3 < 3; ++a3) {
    float cD[LN_VPL], bD[LN_VPL], bD2[LN_VPL];
    float m_c0c, m_cc;
    load_first(rw.D(a3), cD, bD, m_c0c, m_cc);
    const float vD = 2.f * m_c0c;
    const float rhoD = rho1 * vD;
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) {
      bD2[v] = bD[v] * bD[v];
      ro[v] = gam[v] * fmaf(cD[v], rho0, c0[v] * rhoD);
    }
    sto(rw.D(a3), ro);
    float m2;
    load_second(rw.T(a3), bD2, cr, m2);
    const float vT = 2.f * m2 + 2.f * m_cc;
    const float rhoT = rho1 * vT + rho2 * vD * vD;
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) ro[v] = gam[v] * (fmaf(cr[v], rho0, c0[v] * rhoT) + 2.f * cD[v] * rhoD);
    sto(rw.T(a3), ro);
  }
}

// value-only form (R = 1, D = 256): one warp per row, lane l owns columns 4l..4l+3 and 128+4l..128+4l+3
// (two 512-byte requests per tensor), two rows in flight per warp.
template <bool TANH>
__global__ void __launch_bounds__(256)
layernorm_value256_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ scale,
                          const float* __restrict__ bias, float* __restrict__ out, int64_t rows) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const float4 g0 = *reinterpret_cast<const float4*>(scale + 4 * lane), g1 = *reinterpret_cast<const float4*>(scale + 128 + 4 * lane);
  const float4 b0 = *reinterpret_cast<const float4*>(bias + 4 * lane), b1 = *reinterpret_cast<const float4*>(bias + 128 + 4 * lane);
  const int64_t r0 = warp * 2;
  if (r0 >= rows) return;
  const bool two = r0 + 1 < rows;
  float4 xa[2][2], xb[2][2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int64_t r = (u == 0 || two) ? r0 + u : r0;
    xa[u][0] = *reinterpret_cast<const float4*>(a + r * 256 + 4 * lane);
    xa[u][1] = *reinterpret_cast<const float4*>(a + r * 256 + 128 + 4 * lane);
    xb[u][0] = *reinterpret_cast<const float4*>(b + r * 256 + 4 * lane);
    xb[u][1] = *reinterpret_cast<const float4*>(b + r * 256 + 128 + 4 * lane);
  }
  float x[2][8], sum[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const float av[8] = {xa[u][0].x, xa[u][0].y, xa[u][0].z, xa[u][0].w, xa[u][1].x, xa[u][1].y, xa[u][1].z, xa[u][1].w};
    const float bv[8] = {xb[u][0].x, xb[u][0].y, xb[u][0].z, xb[u][0].w, xb[u][1].x, xb[u][1].y, xb[u][1].z, xb[u][1].w};
    sum[u] = 0.f;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      x[u][v] = av[v] + (TANH ? DH_TANH(bv[v]) : bv[v]);
      sum[u] += x[u][v];
    }
  }
  sum[0] = warp_sum(sum[0]);
  sum[1] = warp_sum(sum[1]);
  float sq[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const float mu = sum[u] * (1.0f / 256.0f);
    sq[u] = 0.f;
#pragma unroll
    for (int v = 0; v < 8; ++v) { x[u][v] -= mu; sq[u] = fmaf(x[u][v], x[u][v], sq[u]); }
  }
  sq[0] = warp_sum(sq[0]);
  sq[1] = warp_sum(sq[1]);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (u == 1 && !two) break;
    const float rho = rsqrtf(sq[u] * (1.0f / 256.0f) + 1e-5f);
    float4 o0, o1;
    o0.x = fmaf(x[u][0] * rho, g0.x, b0.x); o0.y = fmaf(x[u][1] * rho, g0.y, b0.y);
    o0.z = fmaf(x[u][2] * rho, g0.z, b0.z); o0.w = fmaf(x[u][3] * rho, g0.w, b0.w);
    o1.x = fmaf(x[u][4] * rho, g1.x, b1.x); o1.y = fmaf(x[u][5] * rho, g1.y, b1.y);
    o1.z = fmaf(x[u][6] * rho, g1.z, b1.z); o1.w = fmaf(x[u][7] * rho, g1.w, b1.w);
    *reinterpret_cast<float4*>(out + (r0 + u) * 256 + 4 * lane) = o0;
    *reinterpret_cast<float4*>(out + (r0 + u) * 256 + 128 + 4 * lane) = o1;
  }
}

int residual_layernorm(const float* a, const float* b, const float* scale, const float* bias, float* out,
                       int64_t B, NetDims d, int tanh_mode, cudaStream_t s) {
  return residual_layernorm_ex(a, b, scale, bias, out, B, d, tanh_mode, 0, 0, 0, s);
}

int residual_layernorm_ex(const float* a, const float* b, const float* scale, const float* bias, float* out,
                          int64_t B, NetDims d, int tanh_mode, int a_comp, int a_pl, int o_pl, cudaStream_t s) {
  if (d.D % 32 != 0 || d.D > 32 * LN_VPL || (a_comp && d.R == 1) || (a_comp && a_pl)) return -2;
  if (!a_pl && !o_pl && d.R == 1 && d.D == 256 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out) |
                                  reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0) {
    const int64_t rows = B * d.N;
    const unsigned grid = (unsigned)((rows + 15) / 16);
    if (tanh_mode) layernorm_value256_kernel<true><<<grid, 256, 0, s>>>(a, b, scale, bias, out, rows);
    else layernorm_value256_kernel<false><<<grid, 256, 0, s>>>(a, b, scale, bias, out, rows);
    return (int)cudaGetLastError();
  }
  const int64_t groups = B * d.N;
  const int wpb = 4;
  unsigned grid = (unsigned)((groups + wpb - 1) / wpb);
  const bool vec4 = d.D == 256 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out) |
                                    reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0;
  if ((a_pl || o_pl) && !vec4) return -2;
  if (tanh_mode) {
    if (vec4) residual_layernorm_kernel<true, true><<<grid, wpb * 32, 0, s>>>(a, b, scale, bias, out, groups, d, a_comp, a_pl, o_pl);
    else residual_layernorm_kernel<true, false><<<grid, wpb * 32, 0, s>>>(a, b, scale, bias, out, groups, d, a_comp, a_pl, o_pl);
  } else {
    if (vec4) residual_layernorm_kernel<false, true><<<grid, wpb * 32, 0, s>>>(a, b, scale, bias, out, groups, d, a_comp, a_pl, o_pl);
    else residual_layernorm_kernel<false, false><<<grid, wpb * 32, 0, s>>>(a, b, scale, bias, out, groups, d, a_comp, a_pl, o_pl);
  }
  return (int)cudaGetLastError();
}

// =============================================================================================
// value-only attention: one block of 256 threads per walker.
// qkv rows: [q (D) | k (D) | v (D)], head h owns columns h*hd .. (h+1)*hd of each.
//   stage q|k|v of the walker in shared memory (row stride 3D + 4 floats: conflict-free float4 reads)
//   scores   thread = (head, 2 queries, 2 keys): four hd-long dot products from four float4 streams
//   softmax  thread = (head, query)
//   output   thread = (2 queries, head, 4 head-dim columns)
// =============================================================================================
constexpr int ATT_NMAX = 32;

__global__ void __launch_bounds__(256)
attention_value_kernel(const float* __restrict__ qkv, float* __restrict__ o, NetDims dm) {
  extern __shared__ __align__(16) float sm[];  // [N][3D + 4] then scores [H][N][N]
  const int N = dm.N, D = dm.D, H = dm.H, hd = dm.hd;
  const int ld = 3 * D + 4;
  float* sc = sm + N * ld;
  const int64_t b = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(qkv + b * N * 3 * D);
  const int row4 = 3 * D / 4;
  // asynchronous copies: all of a thread's 16-byte pieces are in flight at once
  for (int t = threadIdx.x; t < N * row4; t += blockDim.x) {
    const int r = t / row4, c = t % row4;
    const unsigned da = (unsigned)__cvta_generic_to_shared(sm + r * ld + 4 * c);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(src + t) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const float scl = rsqrtf((float)hd);
  // scores: one thread per (head, 2 queries, 2 keys): 4 float4 loads feed 16 FMAs
  const int N2 = (N + 1) / 2;
  for (int t = threadIdx.x; t < H * N2 * N2; t += blockDim.x) {
    const int jb = t % N2, ib = (t / N2) % N2, hh = t / (N2 * N2);
    const int i0 = 2 * ib, i1 = min(i0 + 1, N - 1), j0 = 2 * jb, j1 = min(j0 + 1, N - 1);
    const float4* qa = reinterpret_cast<const float4*>(sm + i0 * ld + hh * hd);
    const float4* qb = reinterpret_cast<const float4*>(sm + i1 * ld + hh * hd);
    const float4* ka = reinterpret_cast<const float4*>(sm + j0 * ld + D + hh * hd);
    const float4* kb = reinterpret_cast<const float4*>(sm + j1 * ld + D + hh * hd);
    float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll 4
    for (int d = 0; d < hd / 4; ++d) {
      const float4 a0 = qa[d], a1 = qb[d], b0 = ka[d], b1 = kb[d];
      s00 = fmaf(a0.x, b0.x, s00); s00 = fmaf(a0.y, b0.y, s00); s00 = fmaf(a0.z, b0.z, s00); s00 = fmaf(a0.w, b0.w, s00);
      s01 = fmaf(a0.x, b1.x, s01); s01 = fmaf(a0.y, b1.y, s01); s01 = fmaf(a0.z, b1.z, s01); s01 = fmaf(a0.w, b1.w, s01);
      s10 = fmaf(a1.x, b0.x, s10); s10 = fmaf(a1.y, b0.y, s10); s10 = fmaf(a1.z, b0.z, s10); s10 = fmaf(a1.w, b0.w, s10);
      s11 = fmaf(a1.x, b1.x, s11); s11 = fmaf(a1.y, b1.y, s11); s11 = fmaf(a1.z, b1.z, s11); s11 = fmaf(a1.w, b1.w, s11);
    }
    float* s0 = sc + (hh * N + i0) * N;
    float* s1 = sc + (hh * N + i0 + 1) * N;
    s0[j0] = s00 * scl;
    if (j0 + 1 < N) s0[j0 + 1] = s01 * scl;
    if (i0 + 1 < N) {
      s1[j0] = s10 * scl;
      if (j0 + 1 < N) s1[j0 + 1] = s11 * scl;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < H * N; t += blockDim.x) {
    float* s = sc + t * N;
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, s[j]);
    float Z = 0.f;
    for (int j = 0; j < N; ++j) { const float e = expf(s[j] - mx); s[j] = e; Z += e; }
    const float iz = 1.f / Z;
    for (int j = 0; j < N; ++j) s[j] *= iz;
  }
  __syncthreads();
  // output: one thread per (2 queries, head, 4 head-dim columns)
  const int hd4 = hd / 4;
  for (int t = threadIdx.x; t < N2 * H * hd4; t += blockDim.x) {
    const int d4 = t % hd4, hh = (t / hd4) % H, i0 = 2 * (t / (hd4 * H));
    const bool has1 = i0 + 1 < N;
    const float* p0 = sc + (hh * N + i0) * N;
    const float* p1 = sc + (hh * N + (has1 ? i0 + 1 : i0)) * N;
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
    for (int j = 0; j < N; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(sm + j * ld + 2 * D + hh * hd + 4 * d4);
      const float pa = p0[j], pb = p1[j];
      acc0.x = fmaf(pa, v.x, acc0.x); acc0.y = fmaf(pa, v.y, acc0.y); acc0.z = fmaf(pa, v.z, acc0.z); acc0.w = fmaf(pa, v.w, acc0.w);
      acc1.x = fmaf(pb, v.x, acc1.x); acc1.y = fmaf(pb, v.y, acc1.y); acc1.z = fmaf(pb, v.z, acc1.z); acc1.w = fmaf(pb, v.w, acc1.w);
    }
    *reinterpret_cast<float4*>(o + (b * N + i0) * D + hh * hd + 4 * d4) = acc0;
    if (has1) *reinterpret_cast<float4*>(o + (b * N + i0 + 1) * D + hh * hd + 4 * d4) = acc1;
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous form (N <= 4 T, hd % 8 == 0, hd <= 64): ONE WARP per (walker, head), no block-level
// synchronisation -- every warp stages its own 3 x N x hd floats with cp.async and runs on while its neighbours
// are still loading, so a resident set of 16-24 warps per SM keeps ~200 KB of loads in flight.
//   scores   lane = (d-half, T x T tile of (query, key)): T*T dot products over half of the head dimension from
//            2T float4 streams, the two halves combined with one shuffle per product
//   softmax  in registers, row maxima / sums over the four lanes that share a query tile
//   output   lane = (query half, 4 head-dim columns): probabilities broadcast from shared memory
// ---------------------------------------------------------------------------------------------
constexpr int AVW_WARPS = 8;       // warps per block
constexpr int AVW_LD = 64 + 4;     // row stride of the staged q / k / v rows (floats): conflict-free float4 tile reads

// NT / HD > 0: electron count / head dimension as compile-time constants (the kernel is instruction-issue bound:
// index arithmetic folds and every loop unrolls); 0: taken from dm at run time.
template <int T, int NT, int HD>
__global__ void __launch_bounds__(AVW_WARPS * 32)
attention_value_warp_kernel(const float* __restrict__ qkv, float* __restrict__ o, int64_t n_items, NetDims dm) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NP = 4 * T;  // padded electron count
  const int N = NT > 0 ? NT : dm.N, D = dm.D, H = dm.H, hd = HD > 0 ? HD : dm.hd;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t item = (int64_t)blockIdx.x * AVW_WARPS + warp;
  if (item >= n_items) return;  // whole warps leave; only __syncwarp below
  const int64_t b = item / H;
  const int hh = (int)(item % H);
  float* qs = sm + (size_t)warp * (3 * NP * AVW_LD + NP * NP);
  float* ks = qs + NP * AVW_LD;
  float* vs = ks + NP * AVW_LD;
  float* ps = vs + NP * AVW_LD;  // [NP][NP] probabilities
  const int hd4 = hd >> 2;
  // ---- stage q, k, v of this head (rows beyond N: zeros, so that padded tiles read finite numbers)
  {
    const float* src = qkv + b * N * 3 * (int64_t)D + hh * hd;
    const int per = NP * hd4;
#pragma unroll
    for (int t = lane; t < 3 * per; t += 32) {
      const int part = t / per, r = (t % per) / hd4, c = t % hd4;
      float* dst = qs + (part * NP + r) * AVW_LD + 4 * c;
      if (r < N) {
        const unsigned da = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(src + (int64_t)r * 3 * D + part * D + 4 * c) : "memory");
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  // ---- scores
  const int half = lane >> 4, ti = (lane & 15) >> 2, tj = lane & 3;
  float sc[T][T];
#pragma unroll
  for (int a = 0; a < T; ++a)
#pragma unroll
    for (int c = 0; c < T; ++c) sc[a][c] = 0.f;
  {
    const int dh4 = hd4 >> 1;  // float4 steps per half
    const float* qb = qs + (T * ti) * AVW_LD + half * (hd >> 1);
    const float* kb = ks + (T * tj) * AVW_LD + half * (hd >> 1);
#pragma unroll
    for (int d = 0; d < dh4; ++d) {
      float4 qv[T], kv[T];
#pragma unroll
      for (int a = 0; a < T; ++a) {
        qv[a] = *reinterpret_cast<const float4*>(qb + a * AVW_LD + 4 * d);
        kv[a] = *reinterpret_cast<const float4*>(kb + a * AVW_LD + 4 * d);
      }
#pragma unroll
      for (int a = 0; a < T; ++a)
#pragma unroll
        for (int c = 0; c < T; ++c) {
          sc[a][c] = fmaf(qv[a].x, kv[c].x, sc[a][c]); sc[a][c] = fmaf(qv[a].y, kv[c].y, sc[a][c]);
          sc[a][c] = fmaf(qv[a].z, kv[c].z, sc[a][c]); sc[a][c] = fmaf(qv[a].w, kv[c].w, sc[a][c]);
        }
    }
  }
  const float scl = rsqrtf((float)hd);
  // ---- softmax over the keys of every query row: the row is spread over the tj = 0..3 lanes of this (half, ti)
#pragma unroll
  for (int a = 0; a < T; ++a) {
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < T; ++c) {
      sc[a][c] = (sc[a][c] + __shfl_xor_sync(0xffffffffu, sc[a][c], 16)) * scl;  // the two d-halves
      if (T * tj + c >= N) sc[a][c] = -INFINITY;                                 // padded keys
      mx = fmaxf(mx, sc[a][c]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float Z = 0.f;
#pragma unroll
    for (int c = 0; c < T; ++c) { sc[a][c] = __expf(sc[a][c] - mx); Z += sc[a][c]; }
    Z += __shfl_xor_sync(0xffffffffu, Z, 1);
    Z += __shfl_xor_sync(0xffffffffu, Z, 2);
    const float iz = 1.f / Z;
    if (half == 0) {
#pragma unroll
      for (int c = 0; c < T; ++c) ps[(T * ti + a) * NP + T * tj + c] = sc[a][c] * iz;
    }
  }
  __syncwarp();
  // ---- output: lane = (query half qh, float4 column d4); queries qh * NP/2 .. + NP/2 - 1
  {
    constexpr int QH = NP / 2;
    const int qh = lane >> 4, d4 = lane & 15;
    if (d4 < hd4) {
      float4 acc[QH];
#pragma unroll
      for (int a = 0; a < QH; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(vs + j * AVW_LD + 4 * d4);
#pragma unroll
        for (int a = 0; a < QH; ++a) {
          const float pw = ps[(qh * QH + a) * NP + j];
          acc[a].x = fmaf(pw, v.x, acc[a].x); acc[a].y = fmaf(pw, v.y, acc[a].y);
          acc[a].z = fmaf(pw, v.z, acc[a].z); acc[a].w = fmaf(pw, v.w, acc[a].w);
        }
      }
#pragma unroll
      for (int a = 0; a < QH; ++a) {
        const int i = qh * QH + a;
        if (i < N) *reinterpret_cast<float4*>(o + (b * N + i) * (int64_t)D + hh * hd + 4 * d4) = acc[a];
      }
    }
  }
}

template <int T, int NT, int HD>
static int attention_value_warp(const float* qkv, float* o, int64_t B, NetDims d, cudaStream_t s) {
  constexpr int NP = 4 * T;
  const size_t smem = (size_t)AVW_WARPS * (3 * NP * AVW_LD + NP * NP) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_value_warp_kernel<T, NT, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int64_t items = B * d.H;
  attention_value_warp_kernel<T, NT, HD><<<(unsigned)((items + AVW_WARPS - 1) / AVW_WARPS), AVW_WARPS * 32, smem, s>>>(qkv, o, items, d);
  return (int)cudaGetLastError();
}

int attention_value(const float* qkv, float* o, int64_t B, NetDims d, cudaStream_t s) {
  if (d.N > ATT_NMAX || (d.hd % 4) != 0) return -2;
  static const bool block_form = dbg_env("DH_ATTN_VALUE") && strcmp(dbg_env("DH_ATTN_VALUE"), "block") == 0;
  if (!block_form && d.hd % 8 == 0 && d.hd <= 64 && (d.D % 4) == 0 && ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
    if (d.hd == 64) {  // the BASELINE configurations: everything a compile-time constant
      if (d.N == 6) return attention_value_warp<2, 6, 64>(qkv, o, B, d, s);
      if (d.N == 10) return attention_value_warp<3, 10, 64>(qkv, o, B, d, s);
      if (d.N == 12) return attention_value_warp<3, 12, 64>(qkv, o, B, d, s);
      if (d.N == 16) return attention_value_warp<4, 16, 64>(qkv, o, B, d, s);
    }
    if (d.N <= 8) return attention_value_warp<2, 0, 0>(qkv, o, B, d, s);
    if (d.N <= 12) return attention_value_warp<3, 0, 0>(qkv, o, B, d, s);
    if (d.N <= 16) return attention_value_warp<4, 0, 0>(qkv, o, B, d, s);
  }
  const size_t smem = ((size_t)d.N * (3 * d.D + 4) + (size_t)d.H * d.N * d.N) * sizeof(float);
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(attention_value_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_smem = smem;
  }
  attention_value_kernel<<<(unsigned)B, 256, smem, s>>>(qkv, o, d);
  return (int)cudaGetLastError();
}


}  // namespace dh
