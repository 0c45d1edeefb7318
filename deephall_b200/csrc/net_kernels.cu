// Psiformer body kernels other than the dense contractions: input features + Dense_0,
// residual + (tanh) + LayerNorm, self-attention -- each in a value-only (R = 1) and a
// forward-Laplacian "jet" form (R = 2N + 8 rows per electron, see common.cuh / oracle/jets.py).
//
// Reference semantics: networks/psiformer.py:37-60 and flax 0.10.2 MultiHeadAttention /
// LayerNorm(epsilon=1e-5).
#include "kernels.h"

namespace dh {

// =============================================================================================
// features (psiformer.py:51-60) fused with Dense_0 (psiformer.py:42, no bias)
// grid: B*N blocks; thread d loops over model columns.
// =============================================================================================
__global__ void features_dense0_kernel(const float* __restrict__ x, const float* __restrict__ W0,
                                       float* __restrict__ h, NetDims dm) {
  extern __shared__ float feat[];  // [R][4]
  const int N = dm.N, R = dm.R, D = dm.D;
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
  float* out = h + bi * R * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float w0 = W0[d], w1 = W0[D + d], w2 = W0[2 * D + d], w3 = W0[3 * D + d];
    for (int r = 0; r < R; ++r) {
      const float* f = feat + r * 4;
      out[(int64_t)r * D + d] = fmaf(f[0], w0, fmaf(f[1], w1, fmaf(f[2], w2, f[3] * w3)));
    }
  }
}

int features_dense0(const float* x, const float* W0, float* h, int64_t B, NetDims d, cudaStream_t s) {
  int threads = d.D >= 256 ? 256 : ((d.D + 31) / 32 * 32);
  features_dense0_kernel<<<(unsigned)(B * d.N), threads, d.R * 4 * sizeof(float), s>>>(x, W0, h, d);
  return (int)cudaGetLastError();
}

// =============================================================================================
// out = LayerNorm(a + b) or LayerNorm(a + tanh(b)), with jets.  One warp per (walker, electron);
// lane l owns columns l, l+32, ...  (D <= 256, D % 32 == 0).
// =============================================================================================
constexpr int LN_VPL = 8;

template <bool TANH>
__global__ void __launch_bounds__(128)
residual_layernorm_kernel(const float* __restrict__ a, const float* __restrict__ b,
                          const float* __restrict__ scale, const float* __restrict__ bias,
                          float* __restrict__ out, int64_t groups, NetDims dm) {
  const int lane = threadIdx.x & 31;
  const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= groups) return;
  const int R = dm.R, D = dm.D, N = dm.N;
  const int vpl = D >> 5;
  const float invD = 1.0f / (float)D;
  const float* ga = a + g * R * D;
  const float* gb = b + g * R * D;
  float* go = out + g * R * D;

  float gam[LN_VPL], bet[LN_VPL];
  float c0[LN_VPL], t1[LN_VPL], t2[LN_VPL];  // centred value row; tanh' and tanh''
#pragma unroll
  for (int v = 0; v < LN_VPL; ++v) {
    if (v < vpl) { gam[v] = scale[lane + 32 * v]; bet[v] = bias[lane + 32 * v]; }
    else { gam[v] = 0.f; bet[v] = 0.f; }
    t1[v] = 1.f; t2[v] = 0.f;
  }
  // ---- value row
  float xr[LN_VPL];
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < LN_VPL; ++v) {
    xr[v] = 0.f;
    if (v < vpl) {
      float av = ga[lane + 32 * v], bv = gb[lane + 32 * v];
      if (TANH) {
        float t = tanhf(bv);
        t1[v] = 1.f - t * t;
        t2[v] = -2.f * t * t1[v];
        bv = t;
      }
      xr[v] = av + bv;
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
  for (int v = 0; v < LN_VPL; ++v)
    if (v < vpl) go[lane + 32 * v] = fmaf(c0[v] * rho0, gam[v], bet[v]);
  if (R == 1) return;

  Rows rw(N, true);
  float accS[LN_VPL], bsq[LN_VPL];
#pragma unroll
  for (int v = 0; v < LN_VPL; ++v) { accS[v] = 0.f; bsq[v] = 0.f; }
  float sum_cc = 0.f, sum_vv = 0.f;  // sum_k mean(cJk^2), sum_k vJk^2

  // first-order row helper: loads x_r, centres it, returns mean(c0*c_r), mean(c_r^2)
  auto load_first = [&](int r, float (&cr)[LN_VPL], float (&braw)[LN_VPL], float& m_c0c, float& m_cc) {
    float s1 = 0.f;
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) {
      cr[v] = 0.f; braw[v] = 0.f;
      if (v < vpl) {
        float av = ga[(int64_t)r * D + lane + 32 * v], bv = gb[(int64_t)r * D + lane + 32 * v];
        braw[v] = bv;
        cr[v] = av + (TANH ? t1[v] * bv : bv);
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
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v) {
      cr[v] = 0.f;
      if (v < vpl) {
        float av = ga[(int64_t)r * D + lane + 32 * v], bv = gb[(int64_t)r * D + lane + 32 * v];
        cr[v] = av + (TANH ? fmaf(t1[v], bv, t2[v] * extra[v]) : bv);
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
      if (v < vpl) {
        go[(int64_t)r * D + lane + 32 * v] = gam[v] * fmaf(cr[v], rho0, c0[v] * rhoJ);
        accS[v] = fmaf(cr[v], rhoJ, accS[v]);
        if (TANH) bsq[v] = fmaf(braw[v], braw[v], bsq[v]);
      }
    }
  }
  // ---- S row
  {
    const int r = rw.S();
    float m_c0c;
    load_second(r, bsq, cr, m_c0c);
    const float vS = 2.f * m_c0c + 2.f * sum_cc;
    const float rhoS = rho1 * vS + rho2 * sum_vv;
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v)
      if (v < vpl)
        go[(int64_t)r * D + lane + 32 * v] = gam[v] * (fmaf(cr[v], rho0, c0[v] * rhoS) + 2.f * accS[v]);
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
      if (v < vpl) go[(int64_t)rw.D(a3) * D + lane + 32 * v] = gam[v] * fmaf(cD[v], rho0, c0[v] * rhoD);
    }
    float m2;
    load_second(rw.T(a3), bD2, cr, m2);
    const float vT = 2.f * m2 + 2.f * m_cc;
    const float rhoT = rho1 * vT + rho2 * vD * vD;
#pragma unroll
    for (int v = 0; v < LN_VPL; ++v)
      if (v < vpl)
        go[(int64_t)rw.T(a3) * D + lane + 32 * v] =
            gam[v] * (fmaf(cr[v], rho0, c0[v] * rhoT) + 2.f * cD[v] * rhoD);
  }
}

int residual_layernorm(const float* a, const float* b, const float* scale, const float* bias, float* out,
                       int64_t B, NetDims d, int tanh_mode, cudaStream_t s) {
  if (d.D % 32 != 0 || d.D > 32 * LN_VPL) return -2;
  const int64_t groups = B * d.N;
  const int wpb = 4;
  unsigned grid = (unsigned)((groups + wpb - 1) / wpb);
  if (tanh_mode)
    residual_layernorm_kernel<true><<<grid, wpb * 32, 0, s>>>(a, b, scale, bias, out, groups, d);
  else
    residual_layernorm_kernel<false><<<grid, wpb * 32, 0, s>>>(a, b, scale, bias, out, groups, d);
  return (int)cudaGetLastError();
}

// =============================================================================================
// value-only attention: one block per walker, one warp per (head, query).
// qkv rows: [q (D) | k (D) | v (D)], head h owns columns h*hd .. (h+1)*hd of each.
// =============================================================================================
constexpr int ATT_NMAX = 32;

__global__ void __launch_bounds__(256)
attention_value_kernel(const float* __restrict__ qkv, float* __restrict__ o, NetDims dm) {
  extern __shared__ float sm[];  // [N][3D]
  const int N = dm.N, D = dm.D, H = dm.H, hd = dm.hd;
  const int64_t b = blockIdx.x;
  const float* src = qkv + b * N * 3 * D;
  for (int t = threadIdx.x; t < N * 3 * D; t += blockDim.x) sm[t] = src[t];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float scl = rsqrtf((float)hd);
  for (int w = warp; w < H * N; w += nwarp) {
    const int hh = w / N, i = w % N;
    const float* q = sm + i * 3 * D + hh * hd;
    float sc[ATT_NMAX];
    float mx = -INFINITY;
#pragma unroll 4
    for (int j = 0; j < N; ++j) {
      const float* k = sm + j * 3 * D + D + hh * hd;
      float p = 0.f;
      for (int d = lane; d < hd; d += 32) p = fmaf(q[d], k[d], p);
      p = warp_sum(p) * scl;
      sc[j] = p;
      mx = fmaxf(mx, p);
    }
    float Z = 0.f;
    for (int j = 0; j < N; ++j) { sc[j] = __expf(sc[j] - mx); Z += sc[j]; }
    const float iz = 1.f / Z;
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(sc[j] * iz, sm[j * 3 * D + 2 * D + hh * hd + d], acc);
      o[(b * N + i) * D + hh * hd + d] = acc;
    }
  }
}

int attention_value(const float* qkv, float* o, int64_t B, NetDims d, cudaStream_t s) {
  if (d.N > ATT_NMAX) return -2;
  size_t smem = (size_t)d.N * 3 * d.D * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attention_value_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  attention_value_kernel<<<(unsigned)B, 256, smem, s>>>(qkv, o, d);
  return (int)cudaGetLastError();
}

// =============================================================================================
// jet attention: one block per (head, walker).
// =============================================================================================
constexpr int AJ_THREADS = 256;
constexpr int AJ_CH = 16;       // head-dim chunk staged in shared memory
constexpr int AJ_STRIDE = 20;   // padded row stride of the staged chunk (bank-conflict-free float4)

struct AJSmem {
  float* qs;   // [N*R][AJ_STRIDE]     staged q (phase 1) / v (phase 2) chunk
  float* sj;   // [N][N][R]            score -> log-softmax -> probability jets
  float* cr;   // [N][N][R]            q_r . k_r cross products
  float* p0;   // [N][N]
  float* qq;   // [N][N]               sum_k l_Jk^2
  float* dd;   // [3][N][N]            l_Da^2
};
__host__ __device__ inline size_t aj_smem_floats(int N, int R) {
  return (size_t)N * R * AJ_STRIDE + 2 * (size_t)N * N * R + 5 * (size_t)N * N;
}
size_t attention_jets_smem(NetDims d) { return aj_smem_floats(d.N, d.R) * sizeof(float); }

template <int NMAX>
__global__ void __launch_bounds__(AJ_THREADS)
attention_jets_kernel(const float* __restrict__ qkv, float* __restrict__ o, NetDims dm) {
  extern __shared__ __align__(16) float smem[];
  const int N = dm.N, R = dm.R, D = dm.D, hd = dm.hd;
  const int hh = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int tid = threadIdx.x;
  const int NR = N * R;
  Rows rw(N, true);
  AJSmem S;
  S.qs = smem;
  S.sj = S.qs + (size_t)NR * AJ_STRIDE;
  S.cr = S.sj + (size_t)N * N * R;
  S.p0 = S.cr + (size_t)N * N * R;
  S.qq = S.p0 + N * N;
  S.dd = S.qq + N * N;
  const int64_t ld = 3 * (int64_t)D;
  const float* base = qkv + b * NR * ld + hh * hd;  // q columns of this head
  const float scl = rsqrtf((float)hd);
  const int nchunk = (hd + AJ_CH - 1) / AJ_CH;

  // ------------------------------------------------------------------ phase 1: score jets
  for (int p0i = 0; p0i < NR; p0i += AJ_THREADS) {
    const int item = p0i + tid;
    const bool active = item < NR;
    const int j = active ? item / R : 0, r = active ? item % R : 0;
    float acc[NMAX][3];
#pragma unroll
    for (int i = 0; i < NMAX; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; }
    for (int ch = 0; ch < nchunk; ++ch) {
      __syncthreads();
      for (int t = tid; t < NR * 4; t += AJ_THREADS) {
        const int row = t >> 2, f4 = t & 3;
        const int dcol = ch * AJ_CH + f4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dcol + 3 < hd) v = *reinterpret_cast<const float4*>(base + row * ld + dcol);
        else {
          const float* pp = base + row * ld + dcol;
          if (dcol < hd) v.x = pp[0];
          if (dcol + 1 < hd) v.y = pp[1];
          if (dcol + 2 < hd) v.z = pp[2];
        }
        *reinterpret_cast<float4*>(S.qs + row * AJ_STRIDE + f4 * 4) = v;
      }
      __syncthreads();
      if (active) {
        float kr[AJ_CH], k0[AJ_CH];
        const float* pkr = base + D + (int64_t)(j * R + r) * ld + ch * AJ_CH;
        const float* pk0 = base + D + (int64_t)(j * R) * ld + ch * AJ_CH;
#pragma unroll
        for (int c = 0; c < AJ_CH; c += 4) {
          const int dcol = ch * AJ_CH + c;
          float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a4;
          if (dcol + 3 < hd) {
            a4 = *reinterpret_cast<const float4*>(pkr + c);
            b4 = *reinterpret_cast<const float4*>(pk0 + c);
          } else {
            if (dcol < hd) { a4.x = pkr[c]; b4.x = pk0[c]; }
            if (dcol + 1 < hd) { a4.y = pkr[c + 1]; b4.y = pk0[c + 1]; }
            if (dcol + 2 < hd) { a4.z = pkr[c + 2]; b4.z = pk0[c + 2]; }
          }
          kr[c] = a4.x; kr[c + 1] = a4.y; kr[c + 2] = a4.z; kr[c + 3] = a4.w;
          k0[c] = b4.x; k0[c + 1] = b4.y; k0[c + 2] = b4.z; k0[c + 3] = b4.w;
        }
#pragma unroll
        for (int i = 0; i < NMAX; ++i) {
          if (i < N) {
            const float* qr = S.qs + (i * R + r) * AJ_STRIDE;
            const float* q0 = S.qs + (i * R) * AJ_STRIDE;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int c = 0; c < AJ_CH; c += 4) {
              float4 x4 = *reinterpret_cast<const float4*>(qr + c);
              float4 y4 = *reinterpret_cast<const float4*>(q0 + c);
              a0 = fmaf(x4.x, k0[c], a0); a0 = fmaf(x4.y, k0[c + 1], a0);
              a0 = fmaf(x4.z, k0[c + 2], a0); a0 = fmaf(x4.w, k0[c + 3], a0);
              a1 = fmaf(y4.x, kr[c], a1); a1 = fmaf(y4.y, kr[c + 1], a1);
              a1 = fmaf(y4.z, kr[c + 2], a1); a1 = fmaf(y4.w, kr[c + 3], a1);
              a2 = fmaf(x4.x, kr[c], a2); a2 = fmaf(x4.y, kr[c + 1], a2);
              a2 = fmaf(x4.z, kr[c + 2], a2); a2 = fmaf(x4.w, kr[c + 3], a2);
            }
            acc[i][0] += a0; acc[i][1] += a1; acc[i][2] += a2;
          }
        }
      }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        if (i < N) {
          const int idx = (i * N + j) * R + r;
          S.sj[idx] = (r == 0) ? acc[i][2] * scl : (acc[i][0] + acc[i][1]) * scl;
          S.cr[idx] = acc[i][2] * scl;
        }
      }
    }
  }
  __syncthreads();
  // second-order rows pick up the cross products: S += 2 sum_k qJk.kJk ; T_a += 2 qDa.kDa
  for (int t = tid; t < N * N * 4; t += AJ_THREADS) {
    const int ij = t >> 2, w = t & 3;
    float* sp = S.sj + ij * R;
    const float* cp = S.cr + ij * R;
    if (w == 0) {
      float s2 = 0.f;
      for (int k = 0; k < 2 * N; ++k) s2 += cp[rw.J(k)];
      sp[rw.S()] += 2.f * s2;
    } else {
      sp[rw.T(w - 1)] += 2.f * cp[rw.D(w - 1)];
    }
  }
  __syncthreads();
  // ------------------------------------------------------------------ softmax jets
  for (int i = tid; i < N; i += AJ_THREADS) {
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, S.sj[(i * N + j) * R]);
    float Z = 0.f;
    for (int j = 0; j < N; ++j) { float e = __expf(S.sj[(i * N + j) * R] - mx); S.p0[i * N + j] = e; Z += e; }
    const float iz = 1.f / Z;
    for (int j = 0; j < N; ++j) S.p0[i * N + j] *= iz;
  }
  __syncthreads();
  const int nfirst = 2 * N + 3;
  for (int t = tid; t < N * nfirst; t += AJ_THREADS) {  // first-order rows: l = s - lse
    const int i = t / nfirst, q = t % nfirst;
    const int r = q < 2 * N ? rw.J(q) : rw.D(q - 2 * N);
    float lse = 0.f;
    for (int j = 0; j < N; ++j) lse = fmaf(S.p0[i * N + j], S.sj[(i * N + j) * R + r], lse);
    for (int j = 0; j < N; ++j) S.sj[(i * N + j) * R + r] -= lse;
  }
  __syncthreads();
  for (int t = tid; t < N * N; t += AJ_THREADS) {
    const float* sp = S.sj + t * R;
    float s2 = 0.f;
    for (int k = 0; k < 2 * N; ++k) s2 = fmaf(sp[rw.J(k)], sp[rw.J(k)], s2);
    S.qq[t] = s2;
    for (int a3 = 0; a3 < 3; ++a3) S.dd[a3 * N * N + t] = sp[rw.D(a3)] * sp[rw.D(a3)];
  }
  __syncthreads();
  for (int t = tid; t < N * 4; t += AJ_THREADS) {  // second-order rows
    const int i = t >> 2, w = t & 3;
    const int r = w == 0 ? rw.S() : rw.T(w - 1);
    const float* extra = w == 0 ? S.qq : S.dd + (w - 1) * N * N;
    float lse = 0.f;
    for (int j = 0; j < N; ++j) {
      float v = S.sj[(i * N + j) * R + r] + extra[i * N + j];
      S.sj[(i * N + j) * R + r] = v;
      lse = fmaf(S.p0[i * N + j], v, lse);
    }
    for (int j = 0; j < N; ++j) S.sj[(i * N + j) * R + r] -= lse;
  }
  __syncthreads();
  for (int t = tid; t < N * N * R; t += AJ_THREADS) {  // l -> p jets
    const int ij = t / R, r = t % R;
    S.sj[t] = r == 0 ? S.p0[ij] : S.p0[ij] * S.sj[t];
  }
  // ------------------------------------------------------------------ phase 2: o = P V jets
  const float* vbase = base + 2 * D;
  float* obase = o + b * NR * (int64_t)D + hh * hd;
  for (int ch = 0; ch < nchunk; ++ch) {
    __syncthreads();
    for (int t = tid; t < NR * 4; t += AJ_THREADS) {
      const int row = t >> 2, f4 = t & 3;
      const int dcol = ch * AJ_CH + f4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dcol + 3 < hd) v = *reinterpret_cast<const float4*>(vbase + row * ld + dcol);
      else {
        const float* pp = vbase + row * ld + dcol;
        if (dcol < hd) v.x = pp[0];
        if (dcol + 1 < hd) v.y = pp[1];
        if (dcol + 2 < hd) v.z = pp[2];
      }
      *reinterpret_cast<float4*>(S.qs + row * AJ_STRIDE + f4 * 4) = v;
    }
    __syncthreads();
    for (int t = tid; t < NR * 4; t += AJ_THREADS) {
      const int f4 = t & 3, ir = t >> 2;
      const int i = ir / R, r = ir % R;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      auto fma4 = [&](float p, const float* vrow) {
        float4 v = *reinterpret_cast<const float4*>(vrow + f4 * 4);
        acc.x = fmaf(p, v.x, acc.x); acc.y = fmaf(p, v.y, acc.y);
        acc.z = fmaf(p, v.z, acc.z); acc.w = fmaf(p, v.w, acc.w);
      };
      for (int j = 0; j < N; ++j) {
        const float* pj = S.sj + (i * N + j) * R;
        const float* vj = S.qs + (size_t)(j * R) * AJ_STRIDE;
        if (r == 0) {
          fma4(pj[0], vj);
        } else {
          fma4(pj[r], vj);
          fma4(pj[0], vj + r * AJ_STRIDE);
          if (r == rw.S()) {
            for (int k = 0; k < 2 * N; ++k) fma4(2.f * pj[rw.J(k)], vj + rw.J(k) * AJ_STRIDE);
          } else if (r >= rw.T(0)) {
            const int rd = rw.D(r - rw.T(0));
            fma4(2.f * pj[rd], vj + rd * AJ_STRIDE);
          }
        }
      }
      const int dcol = ch * AJ_CH + f4 * 4;
      float* dst = obase + (int64_t)ir * D + dcol;
      if (dcol + 3 < hd) *reinterpret_cast<float4*>(dst) = acc;
      else {
        if (dcol < hd) dst[0] = acc.x;
        if (dcol + 1 < hd) dst[1] = acc.y;
        if (dcol + 2 < hd) dst[2] = acc.z;
      }
    }
  }
}

int attention_jets(const float* qkv, float* o, int64_t B, NetDims d, cudaStream_t s) {
  if (d.N > 16 || d.R != 2 * d.N + 8 || (d.hd % 4) != 0 || (d.D % 4) != 0) return -2;
  size_t smem = attention_jets_smem(d);
  dim3 grid((unsigned)d.H, (unsigned)B);
#define DH_AJ(NM)                                                                                           \
  do {                                                                                                      \
    cudaError_t e = cudaFuncSetAttribute(attention_jets_kernel<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return (int)e;                                                                    \
    attention_jets_kernel<NM><<<grid, AJ_THREADS, smem, s>>>(qkv, o, d);                                    \
  } while (0)
  if (d.N <= 4) DH_AJ(4);
  else if (d.N <= 8) DH_AJ(8);
  else if (d.N <= 12) DH_AJ(12);
  else DH_AJ(16);
#undef DH_AJ
  return (int)cudaGetLastError();
}

}  // namespace dh
