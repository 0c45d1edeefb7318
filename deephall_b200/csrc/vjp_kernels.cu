// Reverse-mode kernels for d(Re, Im log psi)/d params contracted with per-walker cotangents
// (loss.py:53-64,96-106 computes the same number from materialised per-walker gradients).
// Value-only activations (R = 1).  Convention: for a complex intermediate z the "gradient"
// G_z satisfies dL = Re(G_z dz); for log psi, G = cot_re - i cot_im.
#include "kernels.h"

namespace dh {

typedef double2 dcplx;

// =============================================================================================
// tail backward: (cot, ld, Minv, x) -> g_c [rows, 2*L*N*K].  One block per (walker, electron).
//   w_k = softmax_k(ld_k) (complex), G_k = G w_k, G^M_ij = G_k (M_k^-1)_ji,
//   dL/dc_re[i,m,j,k] = Re(G^M_ij env_im), dL/dc_im = -Im(G^M_ij env_im).
// =============================================================================================
__global__ void __launch_bounds__(128)
tail_bwd_kernel(const float* __restrict__ cot, const float* __restrict__ ld, const float* __restrict__ Minv,
                const float* __restrict__ x, const double* __restrict__ normfac, float* __restrict__ g_c,
                TailDims dm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = dm.N, L = dm.L, K = dm.K, twoQ = dm.twoQ;
  dcplx* upow = reinterpret_cast<dcplx*>(smraw);
  dcplx* vpow = upow + L;
  cplx* env = reinterpret_cast<cplx*>(vpow + L);  // [L]
  cplx* gm = env + L;                             // [N*K]  G^M_{i j} for this electron i, per (j,k)
  const int64_t bi = blockIdx.x;
  const int64_t b = bi / N;
  const int i = (int)(bi % N);
  const int tid = threadIdx.x;
  // envelope values (double powers, as in the forward)
  {
    double sh, ch, sph, cph;
    sincos(0.5 * (double)x[bi * 2], &sh, &ch);
    sincos(0.5 * (double)x[bi * 2 + 1], &sph, &cph);
    const dcplx u = make_double2(ch * cph, ch * sph);
    const dcplx v = make_double2(sh * cph, -sh * sph);
    if (tid < 2) {
      dcplx z = tid == 0 ? u : v;
      dcplx* tab = tid == 0 ? upow : vpow;
      dcplx p = make_double2(1.0, 0.0);
      for (int e = 0; e <= twoQ; ++e) {
        tab[e] = p;
        p = make_double2(p.x * z.x - p.y * z.y, p.x * z.y + p.y * z.x);
      }
    }
    __syncthreads();
    for (int m = tid; m < L; m += blockDim.x) {
      dcplx a = upow[m], c = vpow[twoQ - m];
      const double nf = normfac[m];
      env[m] = make_float2((float)(nf * (a.x * c.x - a.y * c.y)), (float)(nf * (a.x * c.y + a.y * c.x)));
    }
  }
  // determinant weights and G^M
  const cplx* ldb = reinterpret_cast<const cplx*>(ld) + b * K;
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, ldb[k].x);
  float sr = 0.f, si = 0.f;
  for (int k = 0; k < K; ++k) {
    float e = expf(ldb[k].x - mx), s_, c_;
    sincosf(ldb[k].y, &s_, &c_);
    sr += e * c_; si += e * s_;
  }
  const cplx sinv = cinv(cmake(sr, si));
  const cplx G = cmake(cot[b * 2], -cot[b * 2 + 1]);
  for (int t = tid; t < N * K; t += blockDim.x) {
    const int j = t / K, k = t % K;
    float e = expf(ldb[k].x - mx), s_, c_;
    sincosf(ldb[k].y, &s_, &c_);
    const cplx wk = cmul(cmake(e * c_, e * s_), sinv);
    const cplx mi = reinterpret_cast<const cplx*>(Minv)[((b * K + k) * N + j) * N + i];  // (M^-1)_{j i}
    gm[t] = cmul(cmul(G, wk), mi);
  }
  __syncthreads();
  const int NK = N * K, LNK = L * NK;
  // columns of this electron's spin block; the other block (if any) gets zeros
  const int nsb = dm.n_dn > 0 ? 2 : 1, sbi = (dm.n_dn > 0 && i >= dm.n_up) ? 1 : 0;
  float* row0 = g_c + bi * 2 * (int64_t)nsb * LNK;
  float* row = row0 + 2 * (int64_t)sbi * LNK;
  for (int t = tid; t < LNK; t += blockDim.x) {
    const int m = t / NK, jk = t % NK;
    const cplx val = cmul(gm[jk], env[m]);
    row[t] = val.x;
    row[LNK + t] = -val.y;
  }
  if (nsb == 2) {
    float* oth = row0 + 2 * (int64_t)(1 - sbi) * LNK;
    for (int t = tid; t < 2 * LNK; t += blockDim.x) oth[t] = 0.f;
  }
}

int tail_bwd(const float* cot, const float* ld, const float* Minv, const float* x, const double* normfac,
             float* g_c, int64_t B, TailDims d, cudaStream_t s) {
  size_t smem = 2 * d.L * sizeof(dcplx) + (size_t)(d.L + d.N * d.K) * sizeof(cplx);
  tail_bwd_kernel<<<(unsigned)(B * d.N), 128, smem, s>>>(cot, ld, Minv, x, normfac, g_c, d);
  return (int)cudaGetLastError();
}

// d Jastrow / d ee_par and d ee_anti summed over walkers with weight cot_re (the Jastrow is real):
// parallel pairs -a^2/4/(a+r) with a = ee_par, anti-parallel pairs -a^2/2/(a+r) with a = ee_anti (blocks.py:91-105).
__global__ void jastrow_bwd_kernel(const float* __restrict__ cot, const float* __restrict__ x,
                                   const float* __restrict__ ee_par, const float* __restrict__ ee_anti,
                                   float* __restrict__ g_eepar, float* __restrict__ g_eeanti, float* __restrict__ sq_eepar,
                                   float* __restrict__ sq_eeanti, int64_t B, int N, int n_up) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  float accp = 0.f, acca = 0.f;
  if (b < B && lane < N) {
    const float ap = ee_par ? ee_par[0] : 0.f, aa = ee_anti ? ee_anti[0] : 0.f;
    const float* xw = x + b * N * 2;
    float st, ct, sp, cp;
    sincosf(xw[lane * 2], &st, &ct);
    sincosf(xw[lane * 2 + 1], &sp, &cp);
    const float rx = st * cp, ry = st * sp, rz = ct;
    for (int j = lane + 1; j < N; ++j) {
      float sj, cj, spj, cpj;
      sincosf(xw[j * 2], &sj, &cj);
      sincosf(xw[j * 2 + 1], &spj, &cpj);
      const float dx = rx - sj * cpj, dy = ry - sj * spj, dz = rz - cj;
      const float r = sqrtf(dx * dx + dy * dy + dz * dz);
      const bool par = (lane < n_up) == (j < n_up);
      const float a = par ? ap : aa;
      const float ar = a + r;
      const float d = -(a * a + 2.f * a * r) / (ar * ar);  // d/da [-a^2/(a+r)]
      if (par) accp += 0.25f * d; else acca += 0.5f * d;
    }
    accp *= cot[b * 2];
    acca *= cot[b * 2];
  }
  accp = warp_sum(accp);
  acca = warp_sum(acca);
  if (lane == 0 && b < B) {
    if (g_eepar) atomicAdd(g_eepar, accp);
    if (g_eeanti) atomicAdd(g_eeanti, acca);
    // KFAC diagonal blocks: sum over walkers of the squared per-walker gradient
    if (sq_eepar) atomicAdd(sq_eepar, accp * accp);
    if (sq_eeanti) atomicAdd(sq_eeanti, acca * acca);
  }
}

int jastrow_bwd(const float* cot, const float* x, const float* ee_par, const float* ee_anti, float* g_eepar,
                float* g_eeanti, float* sq_eepar, float* sq_eeanti, int64_t B, int N, int n_up, cudaStream_t s) {
  const int wpb = 4;
  jastrow_bwd_kernel<<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, s>>>(cot, x, ee_par, ee_anti, g_eepar, g_eeanti, sq_eepar,
                                                                      sq_eeanti, B, N, n_up);
  return (int)cudaGetLastError();
}

// ---- helpers of the KFAC factor pass (dh_kfac_factors)
__global__ void fill_unit_cot_kernel(float* __restrict__ cot, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { cot[2 * i] = 1.f; cot[2 * i + 1] = 0.f; }
}
int fill_unit_cot(float* cot, int64_t n, cudaStream_t s) {
  fill_unit_cot_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cot, n);
  return (int)cudaGetLastError();
}
// dst = src on the rows whose electron belongs to spin block sb (0: i < n_up, 1: i >= n_up), zero elsewhere
__global__ void mask_rows_by_spin_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int D, int N, int n_up,
                                         int sb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int el = (int)((i / D) % N);
  dst[i] = ((el >= n_up) == (sb == 1)) ? src[i] : 0.f;
}
int mask_rows_by_spin(const float* src, float* dst, int64_t rows, int D, int N, int n_up, int sb, cudaStream_t s) {
  const int64_t n = rows * D;
  mask_rows_by_spin_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, n, D, N, n_up, sb);
  return (int)cudaGetLastError();
}
// Diagonal Fisher of a LayerNorm's scale and bias: y = xhat * scale + bias with xhat = LN(a + (TANH ? tanh b : b)), so the
// per-walker gradients are sum_i xhat_i gy_i and sum_i gy_i over the walker's electrons; their squares are summed
// over walkers.  One block per walker, thread = column (D <= 1024, D % 32 == 0).
__global__ void ln_fisher_diag_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gy,
                                      float* __restrict__ sq_scale, float* __restrict__ sq_bias, int N, int D, int tanh_mode) {
  __shared__ float red[2][32];
  const int64_t w = blockIdx.x;
  const int d = threadIdx.x, lane = d & 31, wid = d >> 5, nw = blockDim.x >> 5;
  float ss = 0.f, sb = 0.f;
  for (int i = 0; i < N; ++i) {
    const int64_t idx = (w * N + i) * D + d;
    float bv = b[idx];
    if (tanh_mode) bv = tanhf(bv);
    const float u = a[idx] + bv;
    float s1 = warp_sum(u), s2 = warp_sum(u * u);
    __syncthreads();
    if (lane == 0) { red[0][wid] = s1; red[1][wid] = s2; }
    __syncthreads();
    float t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < nw; ++k) { t1 += red[0][k]; t2 += red[1][k]; }
    const float mu = t1 / (float)D;
    const float var = fmaxf(t2 / (float)D - mu * mu, 0.f);
    const float xhat = (u - mu) * rsqrtf(var + 1e-5f);
    const float g = gy[idx];
    ss = fmaf(xhat, g, ss);
    sb += g;
  }
  atomicAdd(sq_scale + d, ss * ss);
  atomicAdd(sq_bias + d, sb * sb);
}
int ln_fisher_diag(const float* a, const float* b, const float* gy, float* sq_scale, float* sq_bias, int64_t B, int N, int D,
                   int tanh_mode, cudaStream_t s) {
  if (D % 32 != 0 || D > 1024) return -2;
  ln_fisher_diag_kernel<<<(unsigned)B, D, 0, s>>>(a, b, gy, sq_scale, sq_bias, N, D, tanh_mode);
  return (int)cudaGetLastError();
}

// =============================================================================================
// In-place inverse of a batch of symmetric positive-definite matrices (the damped Kronecker factors of the KFAC step:
// ~30 matrices of 256-410 rows per iteration, for which library batched LU / Cholesky calls cost 3-100 ms).
// Gauss-Jordan without pivoting (stable for SPD), one block per matrix; the matrix stays in global memory / L2,
// pivot row and column of a step are staged in shared memory.  n <= 1024.
// =============================================================================================
__global__ void __launch_bounds__(1024)
spd_inverse_kernel(float* __restrict__ mats, int n) {
  __shared__ float rowk[1024], colk[1024];
  float* a = mats + (size_t)blockIdx.x * n * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = 0; k < n; ++k) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) { rowk[j] = a[(size_t)k * n + j]; colk[j] = a[(size_t)j * n + k]; }
    __syncthreads();
    const float d = 1.f / rowk[k];
    for (int i = warp; i < n; i += nw) {
      float* ai = a + (size_t)i * n;
      if (i == k) {
        for (int j = lane; j < n; j += 32) ai[j] = j == k ? d : rowk[j] * d;
      } else {
        const float f = colk[i] * d;
        for (int j = lane; j < n; j += 32) ai[j] = j == k ? -f : fmaf(-f, rowk[j], ai[j]);
      }
    }
    __syncthreads();
  }
}
// Cluster form (the one the KFAC step runs): a thread-block CLUSTER of 8 CTAs per matrix, the matrix resident in the
// cluster's distributed shared memory (CTA c holds rows [c rpc, (c + 1) rpc)), so a step costs one pivot-row
// broadcast through DSMEM + one cluster barrier instead of a round trip of the whole matrix through L2:
//   step k: the CTA that owns row k writes the OLD row k into every CTA's row buffer (double-buffered by k & 1),
//           barrier.cluster, then every CTA updates its own rows (it holds its own entries of the old column k).
// One barrier per step is enough: the buffer of step k is rewritten at step k + 2, after barrier k + 1, which a CTA
// only passes once every CTA has finished the update of step k.
constexpr int SPD_CL = 8;        // matrices of more than SPD_SMALL rows
constexpr int SPD_CL_SMALL = 4;  // up to SPD_SMALL rows: four CTAs hold the matrix, twice as many matrices per wave
constexpr int SPD_SMALL = 288;
__device__ __forceinline__ uint32_t spd_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(1024, 1)
spd_inverse_cluster_kernel(float* __restrict__ mats, int n, int rpc, int ldn) {
  extern __shared__ __align__(16) float spd_smem[];
  float* rows = spd_smem;                 // [rpc][ldn]  this CTA's rows
  float* rowbuf = spd_smem + rpc * ldn;   // [2][ldn]    pivot row of the current / next step
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  float* a = mats + (size_t)(blockIdx.x / CL) * n * n;
  const int r0 = (int)rank * rpc, nr = max(0, min(rpc, n - r0));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < nr; i += nw)
    for (int j = lane; j < n; j += 32) rows[i * ldn + j] = a[(size_t)(r0 + i) * n + j];
  __syncthreads();
  const uint32_t rowbuf_s = (uint32_t)__cvta_generic_to_shared(rowbuf);
  for (int k = 0; k < n; ++k) {
    const int owner = k / rpc;
    if ((int)rank == owner) {  // broadcast the old row k (own copy included)
      const float* src = rows + (k - r0) * ldn;
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float v = src[j];
        const uint32_t dst = rowbuf_s + (uint32_t)(((k & 1) * ldn + j) * 4);
#pragma unroll
        for (int c = 0; c < CL; ++c) asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(spd_mapa(dst, c)), "f"(v) : "memory");
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    const float* rk = rowbuf + (k & 1) * ldn;
    const float d = 1.f / rk[k];
    for (int i = warp; i < nr; i += nw) {
      float* ai = rows + i * ldn;
      if (r0 + i == k) {
        for (int j = lane; j < n; j += 32) ai[j] = j == k ? d : rk[j] * d;
      } else {
        const float f = ai[k] * d;
        __syncwarp();  // every lane has read the old a[i][k] before lane k % 32 overwrites it
        for (int j = lane; j < n; j += 32) ai[j] = j == k ? -f : fmaf(-f, rk[j], ai[j]);
      }
    }
    __syncthreads();  // this CTA's rows are complete before the owner of step k + 1 reads its row
  }
  for (int i = warp; i < nr; i += nw)
    for (int j = lane; j < n; j += 32) a[(size_t)(r0 + i) * n + j] = rows[i * ldn + j];
  // no CTA may exit while another one can still write into its row buffer (the last broadcast precedes the last
  // barrier, so this is already guaranteed; the trailing barrier keeps the invariant explicit)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CL>
static int spd_inverse_cluster_launch(float* mats, int n, int batch, cudaStream_t s, bool* fits) {
  const int rpc = (n + CL - 1) / CL, ldn = (n + 3) & ~3;
  const size_t smem = (size_t)(rpc + 2) * ldn * sizeof(float);
  *fits = smem <= 220 * 1024;
  if (!*fits) return 0;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(spd_inverse_cluster_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = smem;
  }
  spd_inverse_cluster_kernel<CL><<<batch * CL, 1024, smem, s>>>(mats, n, rpc, ldn);
  return (int)cudaGetLastError();
}
int spd_inverse_batched(float* mats, int n, int batch, cudaStream_t s) {
  if (n < 1 || n > 1024 || batch < 0) return -2;
  if (batch == 0) return 0;
  bool fits = false;
  int rc = n <= SPD_SMALL ? spd_inverse_cluster_launch<SPD_CL_SMALL>(mats, n, batch, s, &fits)
                          : spd_inverse_cluster_launch<SPD_CL>(mats, n, batch, s, &fits);
  if (fits) return rc;
  spd_inverse_kernel<<<batch, 1024, 0, s>>>(mats, n);
  return (int)cudaGetLastError();
}

// =============================================================================================
// LayerNorm backward: u = a + (TANH ? tanh(b) : b); y = LN(u).  One warp per row, grid-stride.
// g_a = dL/du ; g_b = g_a or g_a * (1 - tanh^2 b) ; g_scale/g_bias accumulated with atomics.
// =============================================================================================
constexpr int LNB_VPL = 8;

template <bool TANH>
__global__ void __launch_bounds__(256)
residual_layernorm_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                              const float* __restrict__ scale, const float* __restrict__ g_out,
                              float* __restrict__ g_a, float* __restrict__ g_b, float* __restrict__ g_scale,
                              float* __restrict__ g_bias, int64_t rows, int D) {
  __shared__ float red[2][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int vpl = D >> 5;
  const float invD = 1.f / (float)D;
  float gam[LNB_VPL], gs[LNB_VPL], gb[LNB_VPL];
#pragma unroll
  for (int v = 0; v < LNB_VPL; ++v) {
    gam[v] = v < vpl ? scale[lane + 32 * v] : 0.f;
    gs[v] = 0.f; gb[v] = 0.f;
  }
  for (int64_t row = (int64_t)blockIdx.x * nwarp + warp; row < rows; row += (int64_t)gridDim.x * nwarp) {
    float u[LNB_VPL], tp[LNB_VPL], gy[LNB_VPL];
    float sum = 0.f;
#pragma unroll
    for (int v = 0; v < LNB_VPL; ++v) {
      u[v] = 0.f; tp[v] = 1.f; gy[v] = 0.f;
      if (v < vpl) {
        const int64_t idx = row * D + lane + 32 * v;
        float bv = b[idx];
        if (TANH) { float t = tanhf(bv); tp[v] = 1.f - t * t; bv = t; }
        u[v] = a[idx] + bv;
        gy[v] = g_out[idx];
        sum += u[v];
      }
    }
    const float mu = warp_sum(sum) * invD;
    float sq = 0.f;
#pragma unroll
    for (int v = 0; v < LNB_VPL; ++v) { u[v] = v < vpl ? u[v] - mu : 0.f; sq += u[v] * u[v]; }
    const float rho = rsqrtf(warp_sum(sq) * invD + 1e-5f);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int v = 0; v < LNB_VPL; ++v) {
      u[v] *= rho;  // x_hat
      const float gx = gy[v] * gam[v];
      m1 += gx; m2 += gx * u[v];
      gs[v] += gy[v] * u[v];
      gb[v] += gy[v];
    }
    m1 = warp_sum(m1) * invD;
    m2 = warp_sum(m2) * invD;
#pragma unroll
    for (int v = 0; v < LNB_VPL; ++v) {
      if (v < vpl) {
        const int64_t idx = row * D + lane + 32 * v;
        const float gu = rho * (gy[v] * gam[v] - m1 - u[v] * m2);
        g_a[idx] = gu;
        g_b[idx] = TANH ? gu * tp[v] : gu;
      }
    }
  }
  // block reduction of g_scale / g_bias partials, then one atomic per column per block
  for (int v = 0; v < vpl; ++v) {
    __syncthreads();
    red[0][threadIdx.x] = gs[v];
    red[1][threadIdx.x] = gb[v];
    __syncthreads();
    if (warp == 0) {
      float s0 = 0.f, s1 = 0.f;
      for (int w = 0; w < nwarp; ++w) { s0 += red[0][w * 32 + lane]; s1 += red[1][w * 32 + lane]; }
      atomicAdd(g_scale + lane + 32 * v, s0);
      atomicAdd(g_bias + lane + 32 * v, s1);
    }
  }
}

int residual_layernorm_bwd(const float* a, const float* b, const float* scale, const float* g_out, float* g_a,
                           float* g_b, float* g_scale, float* g_bias, int64_t rows, int D, int tanh_mode,
                           cudaStream_t s) {
  if (D % 32 != 0 || D > 32 * LNB_VPL) return -2;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (tanh_mode)
    residual_layernorm_bwd_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(a, b, scale, g_out, g_a, g_b, g_scale, g_bias, rows, D);
  else
    residual_layernorm_bwd_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(a, b, scale, g_out, g_a, g_b, g_scale, g_bias, rows, D);
  return (int)cudaGetLastError();
}

// =============================================================================================
// attention backward (value-only): one block per walker.
// =============================================================================================
__global__ void __launch_bounds__(256)
attention_value_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ g_o,
                           float* __restrict__ g_qkv, NetDims dm) {
  extern __shared__ __align__(16) float sm[];
  const int N = dm.N, D = dm.D, H = dm.H, hd = dm.hd;
  const int ldq = 3 * D + 4, ldg = D + 4;  // padded rows: conflict-free float4 reads across electrons
  float* sq = sm;                   // [N][3D + 4]
  float* sgo = sq + N * ldq;        // [N][D + 4]
  float* sp = sgo + N * ldg;        // [H][N][N] scores -> probabilities
  float* sgs = sp + H * N * N;      // [H][N][N] dL/dp -> dL/ds
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  {
    const float4* src = reinterpret_cast<const float4*>(qkv + b * N * 3 * D);
    const int row4 = 3 * D / 4;
    for (int t = tid; t < N * row4; t += blockDim.x) {
      const unsigned da = (unsigned)__cvta_generic_to_shared(sq + (t / row4) * ldq + 4 * (t % row4));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(src + t) : "memory");
    }
    const float4* srg = reinterpret_cast<const float4*>(g_o + b * N * D);
    const int rg4 = D / 4;
    for (int t = tid; t < N * rg4; t += blockDim.x) {
      const unsigned da = (unsigned)__cvta_generic_to_shared(sgo + (t / rg4) * ldg + 4 * (t % rg4));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(srg + t) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const float scl = rsqrtf((float)hd);
  // scores s_ij = q_i.k_j * scl and dL/dp_ij = go_i.v_j : one thread per (head, i, j)
  for (int t = tid; t < H * N * N; t += blockDim.x) {
    const int j = t % N, i = (t / N) % N, hh = t / (N * N);
    const float4* q = reinterpret_cast<const float4*>(sq + i * ldq + hh * hd);
    const float4* k = reinterpret_cast<const float4*>(sq + j * ldq + D + hh * hd);
    const float4* v = reinterpret_cast<const float4*>(sq + j * ldq + 2 * D + hh * hd);
    const float4* go = reinterpret_cast<const float4*>(sgo + i * ldg + hh * hd);
    float p = 0.f, g = 0.f;
    for (int d = 0; d < hd / 4; ++d) {
      const float4 a = q[d], bb = k[d], c = go[d], e = v[d];
      p = fmaf(a.x, bb.x, p); p = fmaf(a.y, bb.y, p); p = fmaf(a.z, bb.z, p); p = fmaf(a.w, bb.w, p);
      g = fmaf(c.x, e.x, g); g = fmaf(c.y, e.y, g); g = fmaf(c.z, e.z, g); g = fmaf(c.w, e.w, g);
    }
    sp[t] = p * scl;
    sgs[t] = g;
  }
  __syncthreads();
  // softmax and dL/ds_ij = p_ij (dL/dp_ij - sum_j p_ij dL/dp_ij) * scl : one thread per (head, i)
  for (int t = tid; t < H * N; t += blockDim.x) {
    float* s = sp + t * N;
    float* gs = sgs + t * N;
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, s[j]);
    float Z = 0.f;
    for (int j = 0; j < N; ++j) { const float ex = expf(s[j] - mx); s[j] = ex; Z += ex; }
    const float iz = 1.f / Z;
    float dot = 0.f;
    for (int j = 0; j < N; ++j) { s[j] *= iz; dot = fmaf(s[j], gs[j], dot); }
    for (int j = 0; j < N; ++j) gs[j] = s[j] * (gs[j] - dot) * scl;
  }
  __syncthreads();
  // g_q_n = sum_m gs_nm k_m ; g_k_n = sum_m gs_mn q_m ; g_v_n = sum_m p_mn go_m : one thread per (n, 4 columns)
  float* out = g_qkv + b * N * 3 * D;
  const int D4 = D / 4;
  for (int t = tid; t < N * D4; t += blockDim.x) {
    const int n = t / D4, c = 4 * (t % D4), hh = c / hd;
    float4 gq = make_float4(0.f, 0.f, 0.f, 0.f), gk = gq, gv = gq;
    for (int m = 0; m < N; ++m) {
      const float a = sgs[(hh * N + n) * N + m], bb = sgs[(hh * N + m) * N + n], pp = sp[(hh * N + m) * N + n];
      const float4 km = *reinterpret_cast<const float4*>(sq + m * ldq + D + c);
      const float4 qm = *reinterpret_cast<const float4*>(sq + m * ldq + c);
      const float4 gm = *reinterpret_cast<const float4*>(sgo + m * ldg + c);
      gq.x = fmaf(a, km.x, gq.x); gq.y = fmaf(a, km.y, gq.y); gq.z = fmaf(a, km.z, gq.z); gq.w = fmaf(a, km.w, gq.w);
      gk.x = fmaf(bb, qm.x, gk.x); gk.y = fmaf(bb, qm.y, gk.y); gk.z = fmaf(bb, qm.z, gk.z); gk.w = fmaf(bb, qm.w, gk.w);
      gv.x = fmaf(pp, gm.x, gv.x); gv.y = fmaf(pp, gm.y, gv.y); gv.z = fmaf(pp, gm.z, gv.z); gv.w = fmaf(pp, gm.w, gv.w);
    }
    *reinterpret_cast<float4*>(out + n * 3 * D + c) = gq;
    *reinterpret_cast<float4*>(out + n * 3 * D + D + c) = gk;
    *reinterpret_cast<float4*>(out + n * 3 * D + 2 * D + c) = gv;
  }
}

int attention_value_bwd(const float* qkv, const float* g_o, float* g_qkv, int64_t B, NetDims d, cudaStream_t s) {
  if (d.N > 32 || (d.hd % 4) != 0) return -2;
  const size_t smem = ((size_t)d.N * (3 * d.D + 4) + (size_t)d.N * (d.D + 4) + 2 * (size_t)d.H * d.N * d.N) * sizeof(float);
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(attention_value_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_smem = smem;
  }
  attention_value_bwd_kernel<<<(unsigned)B, 256, smem, s>>>(qkv, g_o, g_qkv, d);
  return (int)cudaGetLastError();
}

// =============================================================================================
// Dense_0 backward: g_W0[c, d] += sum_rows feat[row, c] g_h[row, d]   (features recomputed)
// =============================================================================================
__global__ void __launch_bounds__(256)
features_dense0_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g_h, float* __restrict__ g_W0,
                           int64_t rows, NetDims dm, int rows_per_block) {
  const int D = dm.D, N = dm.N;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      float st, ct, sp, cp;
      sincosf(x[r * 2], &st, &ct);
      sincosf(x[r * 2 + 1], &sp, &cp);
      const float g = g_h[r * D + d];
      a0 = fmaf(ct, g, a0);
      a1 = fmaf(st * cp, g, a1);
      a2 = fmaf(st * sp, g, a2);
      a3 = fmaf((int)(r % N) < dm.n_up ? 1.f : -1.f, g, a3);
    }
    atomicAdd(g_W0 + d, a0);
    atomicAdd(g_W0 + D + d, a1);
    atomicAdd(g_W0 + 2 * D + d, a2);
    atomicAdd(g_W0 + 3 * D + d, a3);
  }
}

int features_dense0_bwd(const float* x, const float* g_h, float* g_W0, int64_t B, NetDims d, cudaStream_t s) {
  const int64_t rows = B * d.N;
  const int rpb = 64;
  int threads = d.D >= 256 ? 256 : ((d.D + 31) / 32 * 32);
  features_dense0_bwd_kernel<<<(unsigned)((rows + rpb - 1) / rpb), threads, 0, s>>>(x, g_h, g_W0, rows, d, rpb);
  return (int)cudaGetLastError();
}

// out[n] += sum_m g[m*ld + n]
__global__ void __launch_bounds__(256)
colsum_add_kernel(const float* __restrict__ g, float* __restrict__ out, int64_t M, int N, int64_t ld,
                  int rows_per_block) {
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += g[r * ld + n];
  atomicAdd(out + n, acc);
}

int colsum_add(const float* g, float* out, int64_t M, int N, int64_t ld, cudaStream_t s) {
  const int rpb = 256;
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)((M + rpb - 1) / rpb));
  colsum_add_kernel<<<grid, 256, 0, s>>>(g, out, M, N, ld, rpb);
  return (int)cudaGetLastError();
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
int add_inplace(float* dst, const float* src, int64_t n, cudaStream_t s) {
  add_inplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, src, n);
  return (int)cudaGetLastError();
}

}  // namespace dh
