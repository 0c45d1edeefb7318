// Wavefunction tail: monopole-harmonic envelope contraction (networks/blocks.py:59-70), batched
// complex log-determinant with phase (jnp.linalg.slogdet at networks/psiformer.py:74), the
// multi-determinant log-sum-exp (psiformer.py:75-76), the Jastrow factor (blocks.py:77-121), the
// Coulomb / harmonic potential (hamiltonian.py:27-80) and the local-energy assembly
// (hamiltonian.py:121-133,165-169 re-expressed on rotation flows, see oracle/jets.py).
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace dh {

typedef double2 dcplx;
__device__ inline dcplx dmul(dcplx a, dcplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ inline dcplx dadd(dcplx a, dcplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ inline dcplx dscale(dcplx a, double s) { return make_double2(a.x * s, a.y * s); }

// =============================================================================================
// Envelope jets of one electron, in double (exponents reach 2Q = 45; norm factors 1e6).
// env slots: 0 value | 1,2 own tangent flows | 3 S | 4..6 D_a | 7..9 T_a      -> [10][L] cplx
// =============================================================================================
constexpr int ENV_SLOTS = 10;

__device__ void envelope_jets(float theta, float phi, int twoQ, const double* __restrict__ normfac,
                              dcplx* upow, dcplx* vpow, cplx* env, int nslots) {
  const int L = twoQ + 1;
  const int tid = threadIdx.x;
  // the four double-precision sincos are done once (threads 0..3) and shared through upow[0..3] scratch
  __shared__ double trig[8];
  if (tid < 4) {
    const double ang = tid == 0 ? (double)theta : tid == 1 ? (double)phi : tid == 2 ? 0.5 * (double)theta : 0.5 * (double)phi;
    sincos(ang, &trig[2 * tid], &trig[2 * tid + 1]);
  }
  __syncthreads();
  const double st = trig[0], ct = trig[1], sp = trig[2], cp = trig[3], sh = trig[4], ch = trig[5], sph = trig[6], cph = trig[7];
  const dcplx u = make_double2(ch * cph, ch * sph);
  const dcplx v = make_double2(sh * cph, -sh * sph);
  if (tid < 2) {
    dcplx z = tid == 0 ? u : v;
    dcplx* tab = tid == 0 ? upow : vpow;
    dcplx p = make_double2(1.0, 0.0);
    for (int e = 0; e <= twoQ; ++e) { tab[e] = p; p = dmul(p, z); }
  }
  __syncthreads();
  for (int m = tid; m < L; m += blockDim.x) {
    const int a = m, b = twoQ - m;
    const double nf = normfac[m];
    auto pw = [&](const dcplx* tab, int e) { return e >= 0 ? tab[e] : make_double2(0.0, 0.0); };
    const dcplx e0 = dscale(dmul(upow[a], vpow[b]), nf);
    env[0 * L + m] = make_float2((float)e0.x, (float)e0.y);
    if (nslots == 1) continue;
    const dcplx eu = dscale(dmul(pw(upow, a - 1), vpow[b]), nf * a);
    const dcplx ev = dscale(dmul(upow[a], pw(vpow, b - 1)), nf * b);
    const dcplx euu = dscale(dmul(pw(upow, a - 2), vpow[b]), nf * a * (a - 1));
    const dcplx euv = dscale(dmul(pw(upow, a - 1), pw(vpow, b - 1)), nf * a * b);
    const dcplx evv = dscale(dmul(upow[a], pw(vpow, b - 2)), nf * b * (b - 1));
    const dcplx quarter = dscale(dadd(dmul(eu, u), dmul(ev, v)), -0.25);
    // axes: 0 theta_hat, 1 phi_hat, 2..4 x,y,z
    const double ax[5][3] = {{ct * cp, ct * sp, -st}, {-sp, cp, 0.0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    dcplx Ssum = make_double2(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const double nx = ax[q][0], ny = ax[q][1], nz = ax[q][2];
      // (i/2) (n.sigma)^T (u, v)
      dcplx t1 = dadd(dscale(u, nz), dmul(make_double2(nx, ny), v));
      dcplx t2 = dadd(dmul(make_double2(nx, -ny), u), dscale(v, -nz));
      const dcplx du = make_double2(-0.5 * t1.y, 0.5 * t1.x);
      const dcplx dv = make_double2(-0.5 * t2.y, 0.5 * t2.x);
      const dcplx f1 = dadd(dmul(eu, du), dmul(ev, dv));
      dcplx f2 = dadd(dmul(euu, dmul(du, du)), dscale(dmul(euv, dmul(du, dv)), 2.0));
      f2 = dadd(f2, dadd(dmul(evv, dmul(dv, dv)), quarter));
      if (q < 2) {
        env[(1 + q) * L + m] = make_float2((float)f1.x, (float)f1.y);
        Ssum = dadd(Ssum, f2);
      } else {
        env[(4 + q - 2) * L + m] = make_float2((float)f1.x, (float)f1.y);
        env[(7 + q - 2) * L + m] = make_float2((float)f2.x, (float)f2.y);
      }
    }
    env[3 * L + m] = make_float2((float)Ssum.x, (float)Ssum.y);
  }
  __syncthreads();
}

// c rows: [(b,i,r)][ re: (m, j, kdet) | im: (m, j, kdet) ]  ->  Mj[b][kdet][r][i][j] complex
__global__ void __launch_bounds__(128, 8)  // <= 64 registers: the kernel is load-latency bound, occupancy matters
orbital_contract_kernel(const float* __restrict__ c, const float* __restrict__ x,
                        const double* __restrict__ normfac, float* __restrict__ Mj, TailDims dm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = dm.N, R = dm.R, L = dm.L, K = dm.K;
  dcplx* upow = reinterpret_cast<dcplx*>(smraw);
  dcplx* vpow = upow + L;
  cplx* env = reinterpret_cast<cplx*>(vpow + L);
  const int64_t bi = blockIdx.x;
  const int64_t b = bi / N;
  const int i = (int)(bi % N);
  const int nslots = R > 1 ? ENV_SLOTS : 1;
  envelope_jets(x[bi * 2], x[bi * 2 + 1], dm.twoQ, normfac, upow, vpow, env, nslots);
  const int NK = N * K;
  const int LNK = L * NK;
  const int64_t ldc = (dm.n_dn > 0 ? 4 : 2) * (int64_t)LNK;
  const float* cbase = c + bi * R * ldc + ((dm.n_dn > 0 && i >= dm.n_up) ? 2 * LNK : 0);  // this electron's spin block
  Rows rw(N, R > 1);
  for (int t = threadIdx.x; t < R * NK; t += blockDim.x) {
    const int r = t / NK, jk = t % NK;
    const int j = jk / K, kd = jk % K;
    cplx acc = cmake(0.f, 0.f);
    auto dot = [&](int crow, int slot, float w) {
      const float* cr = cbase + (int64_t)crow * ldc + jk;
      const cplx* e = env + slot * L;
      cplx s = cmake(0.f, 0.f);
      int m = 0;
      for (; m + 8 <= L; m += 8) {  // eight independent load pairs in flight before the first FMA needs one
        float re[8], im[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { re[u] = cr[(m + u) * NK]; im[u] = cr[LNK + (m + u) * NK]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) s = cfma(cmake(re[u], im[u]), e[m + u], s);
      }
      for (; m < L; ++m) s = cfma(cmake(cr[m * NK], cr[LNK + m * NK]), e[m], s);
      acc.x = fmaf(w, s.x, acc.x);
      acc.y = fmaf(w, s.y, acc.y);
    };
    dot(r, 0, 1.f);
    if (r > 0) {
      if (r == rw.J(2 * i)) dot(0, 1, 1.f);
      else if (r == rw.J(2 * i + 1)) dot(0, 2, 1.f);
      else if (r == rw.S()) {
        dot(0, 3, 1.f);
        dot(rw.J(2 * i), 1, 2.f);
        dot(rw.J(2 * i + 1), 2, 2.f);
      } else if (r >= rw.D(0) && r < rw.T(0)) {
        dot(0, 4 + (r - rw.D(0)), 1.f);
      } else if (r >= rw.T(0)) {
        const int a3 = r - rw.T(0);
        dot(0, 7 + a3, 1.f);
        dot(rw.D(a3), 4 + a3, 2.f);
      }
    }
    float* dst = Mj + ((((b * K + kd) * R + r) * N + i) * N + j) * 2;
    dst[0] = acc.x;
    dst[1] = acc.y;
  }
}

// =============================================================================================
// Sparse orbitals (blocks.py:52-62): c[i, m, j, k] = sum_s c8[i, s, j, k] Wl[s, m] + bl[m] with c8 the 8-feature
// projection of h.  Both maps are linear, so they fold into one full projection (row d < D: kernel, row D: bias):
//   Weff[d][m][jk] = sum_s W8[d][s][jk] Wl[s][m]          beff[m][jk] = sum_s b8[s][jk] Wl[s][m] (+ bl[m], real part)
// =============================================================================================
__global__ void sparse_fold_kernel(const float* __restrict__ W8, const float* __restrict__ b8, const float* __restrict__ Wl,
                                   const float* __restrict__ bl, int add_bl, float* __restrict__ Weff, int D, int L, int NK) {
  const int d = blockIdx.x;  // D = the bias row
  const float* src = d < D ? W8 + (size_t)d * 8 * NK : b8;
  float* dst = Weff + (size_t)d * L * NK;
  for (int t = threadIdx.x; t < L * NK; t += blockDim.x) {
    const int m = t / NK, jk = t % NK;
    float acc = (d == D && add_bl) ? bl[m] : 0.f;
#pragma unroll
    for (int sft = 0; sft < 8; ++sft) acc = fmaf(src[sft * NK + jk], Wl[sft * L + m], acc);
    dst[t] = acc;
  }
}
// g_W8[d][s][jk] += sum_m g[d][m][jk] Wl[s][m]   (row D: g_b8)
__global__ void sparse_fold_bwd_w8_kernel(const float* __restrict__ g, const float* __restrict__ Wl, float* __restrict__ g_W8,
                                          float* __restrict__ g_b8, int D, int L, int NK) {
  const int d = blockIdx.x;
  const float* gr = g + (size_t)d * L * NK;
  float* dst = d < D ? g_W8 + (size_t)d * 8 * NK : g_b8;
  for (int t = threadIdx.x; t < 8 * NK; t += blockDim.x) {
    const int sft = t / NK, jk = t % NK;
    float acc = 0.f;
    for (int m = 0; m < L; ++m) acc = fmaf(gr[m * NK + jk], Wl[sft * L + m], acc);
    dst[t] += acc;
  }
}
// g_Wl[s][m] += sum_{d <= D, jk} src[d][s][jk] g[d][m][jk]  (src row D = b8);  g_bl[m] += sum_jk g[D][m][jk]
__global__ void sparse_fold_bwd_wl_kernel(const float* __restrict__ g, const float* __restrict__ W8, const float* __restrict__ b8,
                                          int add_bl, float* __restrict__ g_Wl, float* __restrict__ g_bl, int D, int L, int NK) {
  __shared__ float red[8];
  const int sft = blockIdx.x / L, m = blockIdx.x % L;
  float acc = 0.f, accb = 0.f;
  for (int t = threadIdx.x; t < (D + 1) * NK; t += blockDim.x) {
    const int d = t / NK, jk = t % NK;
    const float gv = g[((size_t)d * L + m) * NK + jk];
    const float w = d < D ? W8[((size_t)d * 8 + sft) * NK + jk] : b8[sft * NK + jk];
    acc = fmaf(w, gv, acc);
    if (d == D) accb += gv;
  }
  for (int pass = 0; pass < 2; ++pass) {
    float v = warp_sum(pass == 0 ? acc : accb);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
      if (pass == 0) g_Wl[sft * L + m] += tot;
      else if (add_bl && sft == 0) g_bl[m] += tot;
    }
  }
}
__global__ void __launch_bounds__(256)
sparse_g8_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ Wl, float* __restrict__ g8, int64_t rows,
                 int L, int NK) {
  extern __shared__ float wl_s[];  // [8][L]
  for (int t = threadIdx.x; t < 8 * L; t += blockDim.x) wl_s[t] = Wl[t];
  __syncthreads();
  const int64_t total = rows * 8 * NK;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / (8 * NK);
    const int r = (int)(t - row * 8 * NK), sft = r / NK, jk = r - sft * NK;
    const float* gr = g + row * ldg + jk;
    float acc = 0.f;
    for (int m = 0; m < L; ++m) acc = fmaf(gr[(size_t)m * NK], wl_s[sft * L + m], acc);
    g8[t] = acc;
  }
}
int sparse_g8(const float* g, int64_t ldg, const float* Wl, float* g8, int64_t rows, int L, int NK, cudaStream_t s) {
  const int64_t total = rows * 8 * NK;
  if (total <= 0) return 0;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sparse_g8_kernel<<<(unsigned)blocks, 256, 8 * L * sizeof(float), s>>>(g, ldg, Wl, g8, rows, L, NK);
  return (int)cudaGetLastError();
}
int sparse_fold(const float* W8, const float* b8, const float* Wl, const float* bl, int add_bl, float* Weff, int D, int L,
                int NK, cudaStream_t s) {
  sparse_fold_kernel<<<D + 1, 256, 0, s>>>(W8, b8, Wl, bl, add_bl, Weff, D, L, NK);
  return (int)cudaGetLastError();
}
int sparse_fold_bwd(const float* g_Weff, const float* W8, const float* b8, const float* Wl, int add_bl, float* g_W8,
                    float* g_b8, float* g_Wl, float* g_bl, int D, int L, int NK, cudaStream_t s) {
  sparse_fold_bwd_w8_kernel<<<D + 1, 256, 0, s>>>(g_Weff, Wl, g_W8, g_b8, D, L, NK);
  sparse_fold_bwd_wl_kernel<<<8 * L, 256, 0, s>>>(g_Weff, W8, b8, add_bl, g_Wl, g_bl, D, L, NK);
  return (int)cudaGetLastError();
}

// =============================================================================================
// Laughlin ground state (networks/laughlin.py:59-71, `full_orbitals`): orbital matrix
//   O[i][m] = u_i^m v_i^(2Q1-m) * Jas_i,   Jas_i = prod_{j != i} e_ij,   e_ij = u_i v_j - u_j v_i
// with its jets along the rotation flows.  One block per (walker, electron i).
//   * P_im = u_i^m v_i^(2Q1-m) and its jets come from envelope_jets (unit norm factors);
//   * e_ij is an SU(2) singlet: global rotations leave Jas_i unchanged (D_a, T_a act on P only);
//   * a flow on electron e != i touches one factor: delta Jas_i / Jas_i = g = (u_i dv_e - du_e v_i)/e_ie and
//     delta^2 Jas_i / Jas_i = -1/4  (delta^2 (u, v) = -(u, v)/4);
//   * a flow on electron i touches every factor: with g_j = (du_i v_j - u_j dv_i)/e_ij,
//     delta Jas/Jas = sum_j g_j,  delta^2 Jas/Jas = -(N-1)/4 - sum_j g_j^2 + (sum_j g_j)^2.
// =============================================================================================
__device__ inline dcplx ddiv(dcplx a, dcplx b) {
  const double d = 1.0 / (b.x * b.x + b.y * b.y);
  return make_double2((a.x * b.x + a.y * b.y) * d, (a.y * b.x - a.x * b.y) * d);
}
__device__ inline dcplx dsub(dcplx a, dcplx b) { return make_double2(a.x - b.x, a.y - b.y); }

// Second-order jet of a complex function along ONE flow: (f, delta f, delta^2 f).  Used for the quasiparticle's
// LLL-projected orbital (networks/laughlin.py:85-100), which is not a product of a one-electron factor and an SU(2)
// singlet: its rows are evaluated flow by flow with product / quotient rules instead of closed forms.
struct Jet2 { dcplx f, d, dd; };
__device__ inline Jet2 jconst(dcplx c) { return Jet2{c, make_double2(0, 0), make_double2(0, 0)}; }
__device__ inline Jet2 jadd(Jet2 a, Jet2 b) { return Jet2{dadd(a.f, b.f), dadd(a.d, b.d), dadd(a.dd, b.dd)}; }
__device__ inline Jet2 jsub(Jet2 a, Jet2 b) { return Jet2{dsub(a.f, b.f), dsub(a.d, b.d), dsub(a.dd, b.dd)}; }
__device__ inline Jet2 jscale(Jet2 a, double s) { return Jet2{dscale(a.f, s), dscale(a.d, s), dscale(a.dd, s)}; }
__device__ inline Jet2 jmul(Jet2 a, Jet2 b) {
  return Jet2{dmul(a.f, b.f), dadd(dmul(a.d, b.f), dmul(a.f, b.d)),
              dadd(dadd(dmul(a.dd, b.f), dmul(a.f, b.dd)), dscale(dmul(a.d, b.d), 2.0))};
}
__device__ inline Jet2 jinv(Jet2 a) {  // 1/a: (1/f, -d/f^2, -dd/f^2 + 2 d^2/f^3)
  const dcplx r = ddiv(make_double2(1.0, 0.0), a.f);
  const dcplx r2 = dmul(r, r);
  const dcplx dr = dscale(dmul(a.d, r2), -1.0);
  return Jet2{r, dr, dadd(dscale(dmul(a.dd, r2), -1.0), dscale(dmul(dmul(a.d, a.d), dmul(r2, r)), 2.0))};
}
__device__ inline Jet2 jpow(Jet2 a, int n) {  // n >= 0
  Jet2 p = jconst(make_double2(1.0, 0.0));
  for (int k = 0; k < n; ++k) p = jmul(p, a);
  return p;
}

__global__ void __launch_bounds__(128)
laughlin_orbital_jets_kernel(const float* __restrict__ x, const double* __restrict__ ones, float* __restrict__ Mj,
                             TailDims dm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = dm.N, R = dm.R, L = dm.L;
  dcplx* upow = reinterpret_cast<dcplx*>(smraw);
  dcplx* vpow = upow + L;
  dcplx* U = vpow + L;        // [N]
  dcplx* V = U + N;           // [N]
  dcplx* dU = V + N;          // [N][2] own tangent flows of every electron
  dcplx* dV = dU + 2 * N;     // [N][2]
  dcplx* gE = dV + 2 * N;     // [N][2] delta Jas_i / Jas_i for a flow on electron e
  dcplx* sc = gE + 2 * N;     // [0] Jas, [1..2] dL_t, [3..4] q_t
  dcplx* exd = sc + 5;        // [2N+3] quasiparticle column: delta along flow f (own flows of every electron, then x, y, z)
  dcplx* exdd = exd + 2 * N + 3;  // [2N+3] delta^2 along flow f; [2N+3]: the value
  cplx* env = reinterpret_cast<cplx*>(exdd + 2 * N + 4);
  const int64_t bi = blockIdx.x;
  const int64_t b = bi / N;
  const int i = (int)(bi % N);
  const int tid = threadIdx.x;
  const int nslots = R > 1 ? ENV_SLOTS : 1;
  const float* xw = x + b * N * 2;
  if (tid < N) {
    double st, ct, sp, cp, sh, ch, sph, cph;
    sincos((double)xw[2 * tid], &st, &ct);
    sincos((double)xw[2 * tid + 1], &sp, &cp);
    sincos(0.5 * (double)xw[2 * tid], &sh, &ch);
    sincos(0.5 * (double)xw[2 * tid + 1], &sph, &cph);
    const dcplx u = make_double2(ch * cph, ch * sph), v = make_double2(sh * cph, -sh * sph);
    U[tid] = u; V[tid] = v;
    const double ax[2][3] = {{ct * cp, ct * sp, -st}, {-sp, cp, 0.0}};  // theta_hat, phi_hat
    for (int t = 0; t < 2; ++t) {
      const double nx = ax[t][0], ny = ax[t][1], nz = ax[t][2];
      const dcplx t1 = dadd(dscale(u, nz), dmul(make_double2(nx, ny), v));
      const dcplx t2 = dadd(dmul(make_double2(nx, -ny), u), dscale(v, -nz));
      dU[2 * tid + t] = make_double2(-0.5 * t1.y, 0.5 * t1.x);
      dV[2 * tid + t] = make_double2(-0.5 * t2.y, 0.5 * t2.x);
    }
  }
  envelope_jets(xw[2 * i], xw[2 * i + 1], dm.twoQ, ones, upow, vpow, env, nslots);  // (contains __syncthreads)
  if (tid == 0) {
    dcplx jas = make_double2(1.0, 0.0), dl0 = make_double2(0, 0), dl1 = dl0, s0 = dl0, s1 = dl0;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const dcplx e = dsub(dmul(U[i], V[j]), dmul(U[j], V[i]));
      jas = dmul(jas, e);
      const dcplx g0 = ddiv(dsub(dmul(dU[2 * i], V[j]), dmul(U[j], dV[2 * i])), e);
      const dcplx g1 = ddiv(dsub(dmul(dU[2 * i + 1], V[j]), dmul(U[j], dV[2 * i + 1])), e);
      dl0 = dadd(dl0, g0); dl1 = dadd(dl1, g1);
      s0 = dadd(s0, dmul(g0, g0)); s1 = dadd(s1, dmul(g1, g1));
      gE[2 * j] = ddiv(dsub(dmul(U[i], dV[2 * j]), dmul(dU[2 * j], V[i])), e);
      gE[2 * j + 1] = ddiv(dsub(dmul(U[i], dV[2 * j + 1]), dmul(dU[2 * j + 1], V[i])), e);
    }
    const double c = -0.25 * (N - 1);
    sc[0] = jas; sc[1] = dl0; sc[2] = dl1;
    sc[3] = dadd(make_double2(c, 0.0), dsub(dmul(dl0, dl0), s0));
    sc[4] = dadd(make_double2(c, 0.0), dsub(dmul(dl1, dl1), s1));
  }
  __syncthreads();
  const dcplx jas = sc[0];
  Rows rw(N, R > 1);
  auto E = [&](int slot, int m) { const cplx e = env[slot * L + m]; return make_double2((double)e.x, (double)e.y); };
  // columns: ground state a = 0 .. N-1; quasihole (laughlin.py:75-80): a = 0 .. skip-1, then 2Q1 down to skip+1
  // quasiparticle (laughlin.py:85-100): the N - 1 shell columns a = 0 .. 2Q1, then the projected orbital (below)
  const int ncol = dm.qp ? N - 1 : N;
  for (int t = tid; t < R * ncol; t += blockDim.x) {
    const int r = t / ncol, col = t % ncol;
    const int m = (dm.lskip < 0 || col < dm.lskip) ? col : dm.twoQ - (col - dm.lskip);
    const dcplx P = E(0, m);
    dcplx val;
    if (r == 0) val = dmul(P, jas);
    else if (r <= 2 * N) {
      const int e = (r - 1) >> 1, tt = (r - 1) & 1;
      if (e == i) val = dmul(dadd(E(1 + tt, m), dmul(P, sc[1 + tt])), jas);
      else val = dmul(dmul(P, jas), gE[2 * e + tt]);
    } else if (r == rw.S()) {
      dcplx acc = E(3, m);
      acc = dadd(acc, dscale(dadd(dmul(E(1, m), sc[1]), dmul(E(2, m), sc[2])), 2.0));
      acc = dadd(acc, dmul(P, dadd(sc[3], sc[4])));
      acc = dadd(acc, dscale(P, -0.5 * (N - 1)));
      val = dmul(acc, jas);
    } else if (r < rw.T(0)) val = dmul(E(4 + (r - rw.D(0)), m), jas);
    else val = dmul(E(7 + (r - rw.T(0)), m), jas);
    float* dst = Mj + (((b * R + r) * N + i) * N + col) * 2;
    dst[0] = (float)val.x;
    dst[1] = (float)val.y;
  }
  if (!dm.qp) return;
  // Projected orbital of electron i (laughlin.py:89-99), with a = Q1 + lz, b = 2 Q1 - a:
  //   X_i = Jas_i [ (a+1) u_i^a v_i^(b+1) A_i - (b+1) u_i^(a+1) v_i^b C_i ],  A_i = -sum_{j!=i} u_j / e_ij,  C_i = sum_{j!=i} v_j / e_ij
  // (the reference's u^a v^b ((Q+1+lz) v dJas/dv - (Q+1-lz) u dJas/du); a or b = -1 only with a vanishing coefficient).
  // Thread f < 2N + 3 carries (X, delta X, delta^2 X) along flow f: every spinor moves as delta (u, v) = (dU, dV) (own
  // flows: one electron; f >= 2N: all electrons about the x, y, z axis), delta^2 (u, v) = -(u, v) / 4.
  if (tid < 2 * N + 3) {
    const int f = tid;
    auto spinor = [&](int k, Jet2& ju, Jet2& jv) {
      dcplx du = make_double2(0, 0), dv = du;
      bool moves = false;
      if (f < 2 * N) {
        if ((f >> 1) == k) { du = dU[f]; dv = dV[f]; moves = true; }
      } else {
        const int a3 = f - 2 * N;
        const double nx = a3 == 0, ny = a3 == 1, nz = a3 == 2;
        const dcplx t1 = dadd(dscale(U[k], nz), dmul(make_double2(nx, ny), V[k]));
        const dcplx t2 = dadd(dmul(make_double2(nx, -ny), U[k]), dscale(V[k], -nz));
        du = make_double2(-0.5 * t1.y, 0.5 * t1.x);
        dv = make_double2(-0.5 * t2.y, 0.5 * t2.x);
        moves = true;
      }
      ju = Jet2{U[k], du, moves ? dscale(U[k], -0.25) : make_double2(0, 0)};
      jv = Jet2{V[k], dv, moves ? dscale(V[k], -0.25) : make_double2(0, 0)};
    };
    Jet2 ui, vi;
    spinor(i, ui, vi);
    Jet2 jasj = jconst(make_double2(1.0, 0.0)), A = jconst(make_double2(0, 0)), Cc = A;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      Jet2 uj, vj;
      spinor(j, uj, vj);
      const Jet2 e = jsub(jmul(ui, vj), jmul(uj, vi));
      const Jet2 ie = jinv(e);
      jasj = jmul(jasj, e);
      A = jsub(A, jmul(uj, ie));
      Cc = jadd(Cc, jmul(vj, ie));
    }
    const int a = dm.qp_a, bb = dm.twoQ - a;
    Jet2 X = jconst(make_double2(0, 0));
    if (a + 1 != 0) X = jadd(X, jscale(jmul(jmul(jpow(ui, a), jpow(vi, bb + 1)), A), (double)(a + 1)));
    if (bb + 1 != 0) X = jsub(X, jscale(jmul(jmul(jpow(ui, a + 1), jpow(vi, bb)), Cc), (double)(bb + 1)));
    X = jmul(X, jasj);
    exd[f] = X.d;
    exdd[f] = X.dd;
    if (f == 0) exdd[2 * N + 3] = X.f;
  }
  __syncthreads();
  for (int r = tid; r < R; r += blockDim.x) {
    dcplx val;
    if (r == 0) val = exdd[2 * N + 3];
    else if (r <= 2 * N) val = exd[r - 1];
    else if (r == rw.S()) { val = make_double2(0, 0); for (int f = 0; f < 2 * N; ++f) val = dadd(val, exdd[f]); }
    else if (r < rw.T(0)) val = exd[2 * N + (r - rw.D(0))];
    else val = exdd[2 * N + (r - rw.T(0))];
    float* dst = Mj + (((b * R + r) * N + i) * N + (N - 1)) * 2;
    dst[0] = (float)val.x;
    dst[1] = (float)val.y;
  }
}

int laughlin_orbital_jets(const float* x, const double* ones, float* Mj, int64_t B, TailDims d, cudaStream_t s) {
  if ((d.qp ? d.L != d.N - 1 : d.lskip < 0 ? d.L != d.N : d.L != d.N + 1) || d.K != 1) return -2;
  if (d.qp && (d.qp_a < -1 || d.qp_a > d.twoQ + 1 || 2 * d.N + 3 > 128)) return -2;
  const size_t smem = (2 * (size_t)d.L + 12 * (size_t)d.N + 12) * sizeof(dcplx) + (size_t)ENV_SLOTS * d.L * sizeof(cplx);
  laughlin_orbital_jets_kernel<<<(unsigned)(B * d.N), 128, smem, s>>>(x, ones, Mj, d);
  return (int)cudaGetLastError();
}

// Value-only form (R = 1): one warp per (walker, electron), eight per block.
//   env[m] = sqrt(C(2Q,m)) cos^m(theta/2) sin^(2Q-m)(theta/2) e^{i (m - Q) phi}: magnitudes by repeated squaring
//   in double (exponents reach 2Q), phase angle reduced in double and evaluated in fp32;
//   M[i][j,kd] = sum_m c[m][j,kd] env[m] with lanes = (m-group, column), coalesced 4-byte loads.

__global__ void __launch_bounds__(256)
orbital_value_kernel(const float* __restrict__ c, const float* __restrict__ x, const double* __restrict__ normfac,
                     float* __restrict__ Mj, int64_t rows, TailDims dm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = dm.N, L = dm.L, K = dm.K, twoQ = dm.twoQ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cplx* env = reinterpret_cast<cplx*>(smraw) + warp * L;
  const int64_t bi = (int64_t)blockIdx.x * 8 + warp;
  if (bi >= rows) return;  // whole warps leave; only __syncwarp is used below
  const int64_t b = bi / N;
  const int i = (int)(bi % N);
  const int NK = N * K, LNK = L * NK;
  const float* cr = c + bi * (dm.n_dn > 0 ? 4 : 2) * (int64_t)LNK + ((dm.n_dn > 0 && i >= dm.n_up) ? 2 * LNK : 0);
  const double phi = (double)x[bi * 2 + 1];
  double sh = 0.0, ch = 0.0;
  if (lane == 0) sincos(0.5 * (double)x[bi * 2], &sh, &ch);
  sh = __shfl_sync(0xffffffffu, sh, 0);
  ch = __shfl_sync(0xffffffffu, ch, 0);
  for (int m = lane; m < L; m += 32) {
    const double mag = normfac[m] * dpow_int(ch, m) * dpow_int(sh, twoQ - m);
    double psi = (double)(2 * m - twoQ) * 0.5 * phi;
    psi -= 6.283185307179586476925287 * rint(psi * 0.15915494309189533576888);
    float sp, cp;
    sincosf((float)psi, &sp, &cp);
    env[m] = make_float2((float)(mag * (double)cp), (float)(mag * (double)sp));
  }
  __syncwarp();
  for (int jk0 = 0; jk0 < NK; jk0 += 32) {
    const int width = NK - jk0 < 32 ? NK - jk0 : 32;
    int NKp = 1;
    while (NKp < width) NKp <<= 1;
    const int G = 32 / NKp, g = lane / NKp, jl = lane % NKp;
    cplx acc = cmake(0.f, 0.f);
    if (jl < width) {
      const float* p = cr + jk0 + jl;
      // batches of 6 m-values: the 12 loads of a batch are issued before the first FMA needs one
      for (int m0 = g; m0 < L; m0 += 6 * G) {
        float re[6], im[6];
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          const int m = m0 + u * G;
          re[u] = m < L ? p[m * NK] : 0.f;
          im[u] = m < L ? p[LNK + m * NK] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          const int m = m0 + u * G;
          if (m < L) acc = cfma(cmake(re[u], im[u]), env[m], acc);
        }
      }
    }
    for (int off = NKp; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    }
    if (lane < width) {
      const int jk = jk0 + lane, j = jk / K, kd = jk % K;
      float* dst = Mj + ((((b * K + kd)) * N + i) * N + j) * 2;
      dst[0] = acc.x;
      dst[1] = acc.y;
    }
  }
}

// Vectorised jet form (N K % VEC == 0).  The 46 (c-row, envelope-slot) products of an electron -- one per output row
// plus the 14 product-rule terms -- are spread as (product, VEC columns) items over the threads of the block, every
// item a dot over the L orbitals from 16-byte (VEC = 4) loads, and the weighted results are summed into the output
// rows in shared memory.  Against one thread per (row, column) with 4-byte loads: a quarter of the load
// instructions, no idle tail (the S row alone has four terms), a fraction of the code size (the scalar kernel
// stalled 3.8 issue slots per instruction on instruction fetch).
constexpr int OCV_THREADS = 160;

template <int VEC>
__global__ void __launch_bounds__(OCV_THREADS, 5)
orbital_contract_vec_kernel(const float* __restrict__ c, const float* __restrict__ x, const double* __restrict__ normfac,
                            float* __restrict__ Mj, TailDims dm, int prefetch) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = dm.N, R = dm.R, L = dm.L, K = dm.K;
  dcplx* upow = reinterpret_cast<dcplx*>(smraw);
  dcplx* vpow = upow + L;
  cplx* env = reinterpret_cast<cplx*>(vpow + L);          // [ENV_SLOTS][L]
  const int NK = N * K;
  float* outs = reinterpret_cast<float*>(env + ENV_SLOTS * L);  // [R + 14][NK][2]: one slot per product (summed in a fixed order below)
  const int64_t bi = blockIdx.x;
  const int64_t b = bi / N;
  const int i = (int)(bi % N);
  const int LNK = L * NK;
  const int64_t ldc = (dm.n_dn > 0 ? 4 : 2) * (int64_t)LNK;
  const float* cbase = c + bi * R * ldc + ((dm.n_dn > 0 && i >= dm.n_up) ? 2 * LNK : 0);
  if (prefetch) {
    // the electron's coefficient rows (R x 2 LNK floats, contiguous per row) start their way from HBM to L2 while the
    // fp64 envelope prologue runs: the dot loops below then stream from L2
    const int lines_per_row = (2 * LNK * (int)sizeof(float) + 127) / 128;
    for (int t = threadIdx.x; t < R * lines_per_row; t += blockDim.x) {
      const int r = t / lines_per_row, ln = t - r * lines_per_row;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(cbase + (int64_t)r * ldc) + (size_t)ln * 128));
    }
  }
  envelope_jets(x[bi * 2], x[bi * 2 + 1], dm.twoQ, normfac, upow, vpow, env, ENV_SLOTS);  // ends with __syncthreads
  Rows rw(N, true);
  const int q = NK / VEC;
  const int ndots = R + 14;
  for (int item = threadIdx.x; item < ndots * q; item += blockDim.x) {
    const int d = item / q, jq = item - d * q;
    // product d: c-row `crow` . envelope slot `slot`, times w, into output row `orow`
    int crow, slot, orow;
    float w = 1.f;
    if (d < R) { crow = d; slot = 0; orow = d; }
    else {
      const int e = d - R;
      if (e < 2) { crow = 0; slot = 1 + e; orow = rw.J(2 * i + e); }
      else if (e == 2) { crow = 0; slot = 3; orow = rw.S(); }
      else if (e < 5) { crow = rw.J(2 * i + (e - 3)); slot = 1 + (e - 3); orow = rw.S(); w = 2.f; }
      else if (e < 8) { crow = 0; slot = 4 + (e - 5); orow = rw.D(e - 5); }
      else if (e < 11) { crow = 0; slot = 7 + (e - 8); orow = rw.T(e - 8); }
      else { crow = rw.D(e - 11); slot = 4 + (e - 11); orow = rw.T(e - 11); w = 2.f; }
    }
    const float* cr = cbase + (int64_t)crow * ldc + VEC * jq;
    const cplx* ev = env + slot * L;
    float ar[VEC], ai[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { ar[v] = 0.f; ai[v] = 0.f; }
    auto ld = [&](const float* p, float (&dst)[VEC]) {
      if (VEC == 4) { const float4 a = *reinterpret_cast<const float4*>(p); dst[0] = a.x; dst[1] = a.y; dst[2 % VEC] = a.z; dst[3 % VEC] = a.w; }
      else { const float2 a = *reinterpret_cast<const float2*>(p); dst[0] = a.x; dst[1] = a.y; }
    };
    int m = 0;
    for (; m + 4 <= L; m += 4) {  // eight 16-byte loads in flight; two consecutive m cover whole 32-byte sectors
      float re[4][VEC], im[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) { ld(cr + (m + u) * NK, re[u]); ld(cr + LNK + (m + u) * NK, im[u]); }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const cplx e = ev[m + u];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          ar[v] = fmaf(re[u][v], e.x, ar[v]); ar[v] = fmaf(-im[u][v], e.y, ar[v]);
          ai[v] = fmaf(re[u][v], e.y, ai[v]); ai[v] = fmaf(im[u][v], e.x, ai[v]);
        }
      }
    }
    for (; m < L; ++m) {
      float re[VEC], im[VEC];
      ld(cr + m * NK, re);
      ld(cr + LNK + m * NK, im);
      const cplx e = ev[m];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        ar[v] = fmaf(re[v], e.x, ar[v]); ar[v] = fmaf(-im[v], e.y, ar[v]);
        ai[v] = fmaf(re[v], e.y, ai[v]); ai[v] = fmaf(im[v], e.x, ai[v]);
      }
    }
    (void)orow;
    float* o = outs + ((size_t)d * NK + VEC * jq) * 2;
#pragma unroll
    for (int v = 0; v < VEC; ++v) { o[2 * v] = w * ar[v]; o[2 * v + 1] = w * ai[v]; }
  }
  __syncthreads();
  // output row r = its own product + the product-rule terms that belong to it, in a fixed order (bitwise reproducible)
  for (int t = threadIdx.x; t < R * NK; t += blockDim.x) {
    const int r = t / NK, jk = t - r * NK;
    const int j = jk / K, kd = jk - j * K;
    float re = outs[2 * t], im = outs[2 * t + 1];
    auto add = [&](int e) { re += outs[2 * ((R + e) * NK + jk)]; im += outs[2 * ((R + e) * NK + jk) + 1]; };
    if (r == rw.J(2 * i)) add(0);
    else if (r == rw.J(2 * i + 1)) add(1);
    else if (r == rw.S()) { add(2); add(3); add(4); }
    else if (r >= rw.D(0) && r < rw.T(0)) add(5 + (r - rw.D(0)));
    else if (r >= rw.T(0)) { add(8 + (r - rw.T(0))); add(11 + (r - rw.T(0))); }
    float* dst = Mj + ((((b * K + kd) * R + r) * N + i) * N + j) * 2;
    dst[0] = re;
    dst[1] = im;
  }
}

// Envelope jets of every electron as a table [electrons][ENV_SLOTS][L] complex: the right operand of the envelope
// contraction when it runs as the epilogue of the orbital projection (gemm_tc.cu, ORB).  One block per electron.
__global__ void __launch_bounds__(64)
envelope_table_kernel(const float* __restrict__ x, const double* __restrict__ normfac, const float* __restrict__ bre,
                      const float* __restrict__ bim, const float* __restrict__ unscale, float* __restrict__ tab, TailDims dm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int L = dm.L, NK = dm.N * dm.K;
  dcplx* upow = reinterpret_cast<dcplx*>(smraw);
  dcplx* vpow = upow + L;
  cplx* env = reinterpret_cast<cplx*>(vpow + L);  // [ENV_SLOTS][L]
  const int64_t bi = blockIdx.x;
  envelope_jets(x[bi * 2], x[bi * 2 + 1], dm.twoQ, normfac, upow, vpow, env, ENV_SLOTS);  // ends with __syncthreads
  float* dst = tab + bi * (int64_t)(ENV_SLOTS * 2 * L + ENV_SLOTS * 2 * NK);
  const float us = unscale ? __ldg(unscale) : 1.f;  // the contraction's accumulators are (weights x power of two): undone here
  const float* src = reinterpret_cast<const float*>(env);
  for (int t = threadIdx.x; t < ENV_SLOTS * 2 * L; t += blockDim.x) dst[t] = src[t] * us;
  // bias products: the bias sits on the value row, so its share of every output is sum_m bias(m, j) env_s[m]
  for (int t = threadIdx.x; t < ENV_SLOTS * NK; t += blockDim.x) {
    const int sl = t / NK, j = t - sl * NK;
    cplx acc = cmake(0.f, 0.f);
    for (int m = 0; m < L; ++m) acc = cfma(cmake(bre[m * NK + j], bim[m * NK + j]), env[sl * L + m], acc);
    dst[ENV_SLOTS * 2 * L + 2 * t] = acc.x;
    dst[ENV_SLOTS * 2 * L + 2 * t + 1] = acc.y;
  }
}
int envelope_table(const float* x, const double* normfac, const float* bre, const float* bim, const float* unscale, float* tab,
                   int64_t B, TailDims d, cudaStream_t s) {
  const size_t smem = 2 * d.L * sizeof(dcplx) + (size_t)ENV_SLOTS * d.L * sizeof(cplx);
  envelope_table_kernel<<<(unsigned)(B * d.N), 64, smem, s>>>(x, normfac, bre, bim, unscale, tab, d);
  return (int)cudaGetLastError();
}

// Prologue of a jet pass in ONE launch, one block per electron: the compressed feature maps (features_dense0_kernel's
// 10 non-zero jet rows: Dense_0 -> h and the first layer's q|k|v = feat (W0 Wqkv) + b) followed by the envelope table above.
// The maps are store-bound (40 KB per electron); the fp64 envelope jets run while those stores drain.
__global__ void __launch_bounds__(64)
jets_prologue_kernel(const float* __restrict__ x, const float* __restrict__ W0, float* __restrict__ h, int n0,
                     const float* __restrict__ W1, const float* __restrict__ b1, float* __restrict__ q, int n1,
                     const double* __restrict__ normfac, const float* __restrict__ bre, const float* __restrict__ bim,
                     const float* __restrict__ unscale, float* __restrict__ tab, int n_up, TailDims dm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ __align__(16) float feat[ENV_SLOTS][4];  // value | own theta, phi flows | S | D_a | T_a
  const int64_t bi = blockIdx.x;
  const int i = (int)(bi % dm.N);
  const float th = x[bi * 2], ph = x[bi * 2 + 1];
  if (threadIdx.x == 0) {
    float st, ct, sp, cp;
    sincosf(th, &st, &ct);
    sincosf(ph, &sp, &cp);
    const float rx = st * cp, ry = st * sp, rz = ct;
    const float f[ENV_SLOTS][4] = {{rz, rx, ry, i < n_up ? 1.f : -1.f},  // feature order: (z, x, y, spin)
                                   {0.f, sp, -cp, 0.f},                   // theta_hat x r = -phi_hat
                                   {-st, ct * cp, ct * sp, 0.f},          // phi_hat x r = theta_hat
                                   {-2.f * rz, -2.f * rx, -2.f * ry, 0.f},
                                   {ry, 0.f, -rz, 0.f}, {-rx, rz, 0.f, 0.f}, {0.f, -ry, rx, 0.f},   // e_a x r
                                   {-rz, 0.f, -ry, 0.f}, {-rz, -rx, 0.f, 0.f}, {0.f, -rx, -ry, 0.f}};  // e_a r_a - r
#pragma unroll
    for (int r = 0; r < ENV_SLOTS; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) feat[r][k] = f[r][k];
  }
  __syncthreads();
  for (int g = threadIdx.x; g < (n0 + n1) / 4; g += blockDim.x) {
    const bool second = g >= n0 / 4;
    const int d = 4 * (second ? g - n0 / 4 : g), nout = second ? n1 : n0;
    const float* W = second ? W1 : W0;
    float* out = (second ? q : h) + bi * ENV_SLOTS * nout + d;
    const float4 w0 = *reinterpret_cast<const float4*>(W + d), w1 = *reinterpret_cast<const float4*>(W + nout + d);
    const float4 w2 = *reinterpret_cast<const float4*>(W + 2 * nout + d), w3 = *reinterpret_cast<const float4*>(W + 3 * nout + d);
    const float4 bb = second && b1 != nullptr ? *reinterpret_cast<const float4*>(b1 + d) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < ENV_SLOTS; ++r) {
      const float4 f = *reinterpret_cast<const float4*>(feat[r]);
      float4 v;
      v.x = fmaf(f.x, w0.x, fmaf(f.y, w1.x, fmaf(f.z, w2.x, f.w * w3.x)));
      v.y = fmaf(f.x, w0.y, fmaf(f.y, w1.y, fmaf(f.z, w2.y, f.w * w3.y)));
      v.z = fmaf(f.x, w0.z, fmaf(f.y, w1.z, fmaf(f.z, w2.z, f.w * w3.z)));
      v.w = fmaf(f.x, w0.w, fmaf(f.y, w1.w, fmaf(f.z, w2.w, f.w * w3.w)));
      if (r == 0) { v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w; }
      *reinterpret_cast<float4*>(out + (int64_t)r * nout) = v;
    }
  }
  // ---- envelope table (envelope_table_kernel's arithmetic)
  const int L = dm.L, NK = dm.N * dm.K;
  dcplx* upow = reinterpret_cast<dcplx*>(smraw);
  dcplx* vpow = upow + L;
  cplx* env = reinterpret_cast<cplx*>(vpow + L);  // [ENV_SLOTS][L]
  envelope_jets(th, ph, dm.twoQ, normfac, upow, vpow, env, ENV_SLOTS);  // ends with __syncthreads
  float* dst = tab + bi * (int64_t)(ENV_SLOTS * 2 * L + ENV_SLOTS * 2 * NK);
  const float us = unscale ? __ldg(unscale) : 1.f;
  const float* src = reinterpret_cast<const float*>(env);
  for (int t = threadIdx.x; t < ENV_SLOTS * 2 * L; t += blockDim.x) dst[t] = src[t] * us;
  for (int t = threadIdx.x; t < ENV_SLOTS * NK; t += blockDim.x) {
    const int sl = t / NK, j = t - sl * NK;
    cplx acc = cmake(0.f, 0.f);
    for (int m = 0; m < L; ++m) acc = cfma(cmake(bre[m * NK + j], bim[m * NK + j]), env[sl * L + m], acc);
    dst[ENV_SLOTS * 2 * L + 2 * t] = acc.x;
    dst[ENV_SLOTS * 2 * L + 2 * t + 1] = acc.y;
  }
}
int jets_prologue(const float* x, const float* W0, float* h, int n0, const float* W1, const float* b1, float* q, int n1,
                  const double* normfac, const float* bre, const float* bim, const float* unscale, float* tab, int64_t B, int n_up,
                  TailDims d, cudaStream_t s) {
  if ((n0 | n1) & 3) return -2;
  if ((reinterpret_cast<uintptr_t>(W0) | reinterpret_cast<uintptr_t>(W1) | reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(h) |
       reinterpret_cast<uintptr_t>(q)) & 15)
    return -2;
  const size_t smem = 2 * d.L * sizeof(dcplx) + (size_t)ENV_SLOTS * d.L * sizeof(cplx);
  // 64 threads per electron: the envelope jets are a chain of short serial fp64 phases, so many small blocks per SM overlap
  // better than few large ones (measured per 1024-walker chunk: 64 threads 123 us, 128: 146 us, 256: 207 us; the three
  // separate launches this replaces: 148 us)
  jets_prologue_kernel<<<(unsigned)(B * d.N), 64, smem, s>>>(x, W0, h, n0, W1, b1, q, n1, normfac, bre, bim, unscale, tab, n_up, d);
  return (int)cudaGetLastError();
}

// Value-only form of the table: per electron [L] complex envelope values (times *unscale) followed by [N K] complex bias
// products.  One warp per electron (the envelope of orbital_value_kernel).
__global__ void __launch_bounds__(256)
envelope_value_table_kernel(const float* __restrict__ x, const double* __restrict__ normfac, const float* __restrict__ bre,
                            const float* __restrict__ bim, const float* __restrict__ unscale, float* __restrict__ tab, int64_t rows,
                            TailDims dm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int L = dm.L, NK = dm.N * dm.K, twoQ = dm.twoQ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cplx* env = reinterpret_cast<cplx*>(smraw) + warp * L;
  const int64_t bi = (int64_t)blockIdx.x * 8 + warp;
  if (bi >= rows) return;  // whole warps leave; only __syncwarp is used below
  const double phi = (double)x[bi * 2 + 1];
  double sh = 0.0, ch = 0.0;
  if (lane == 0) sincos(0.5 * (double)x[bi * 2], &sh, &ch);
  sh = __shfl_sync(0xffffffffu, sh, 0);
  ch = __shfl_sync(0xffffffffu, ch, 0);
  const float us = unscale ? __ldg(unscale) : 1.f;
  float* dst = tab + bi * (int64_t)(2 * (L + NK));
  for (int m = lane; m < L; m += 32) {
    const double mag = normfac[m] * dpow_int(ch, m) * dpow_int(sh, twoQ - m);
    double psi = (double)(2 * m - twoQ) * 0.5 * phi;
    psi -= 6.283185307179586476925287 * rint(psi * 0.15915494309189533576888);
    float sp, cp;
    sincosf((float)psi, &sp, &cp);
    const cplx e = make_float2((float)(mag * (double)cp), (float)(mag * (double)sp));
    env[m] = e;
    dst[2 * m] = e.x * us;
    dst[2 * m + 1] = e.y * us;
  }
  __syncwarp();
  for (int j = lane; j < NK; j += 32) {
    cplx acc = cmake(0.f, 0.f);
    for (int m = 0; m < L; ++m) acc = cfma(cmake(bre[m * NK + j], bim[m * NK + j]), env[m], acc);
    dst[2 * L + 2 * j] = acc.x;
    dst[2 * L + 2 * j + 1] = acc.y;
  }
}
int envelope_value_table(const float* x, const double* normfac, const float* bre, const float* bim, const float* unscale, float* tab,
                         int64_t B, TailDims d, cudaStream_t s) {
  const int64_t rows = B * d.N;
  envelope_value_table_kernel<<<(unsigned)((rows + 7) / 8), 256, 8 * d.L * sizeof(cplx), s>>>(x, normfac, bre, bim, unscale, tab, rows, d);
  return (int)cudaGetLastError();
}

// Orbital-projection kernels [D][L N] (real part, imaginary part) and biases -> ONE fp32 matrix [D][ncol] and bias [ncol]
// whose columns are ordered for the fused envelope contraction: tile t (256 columns) holds the orbitals m = g t .. g t + g - 1
// (g = orb_per_tile(L)), each as 24 columns [re(m, 0..11) | im(m, 0..11)]; the remaining columns of a tile are zero.
__global__ void orb_permute_weights_kernel(const float* __restrict__ Wre, const float* __restrict__ Wim, const float* __restrict__ bre,
                                           const float* __restrict__ bim, float* __restrict__ Wp, float* __restrict__ bp, int D,
                                           int L, int NK, int ncol, int mpt) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)(D + 1) * ncol) return;
  const int k = (int)(t / ncol), n = (int)(t % ncol);
  const int tile = n >> 8, rem = n & 255, ml = rem / (2 * NK), w = rem - ml * 2 * NK;
  const int m = mpt * tile + ml, part = w / NK, col = w - part * NK;
  const bool valid = ml < mpt && m < L;
  const int src = m * NK + col;
  if (k < D) Wp[(int64_t)k * ncol + n] = valid ? (part ? Wim : Wre)[(int64_t)k * L * NK + src] : 0.f;
  else bp[n] = valid ? (part ? bim : bre)[src] : 0.f;
}
int orb_permute_weights(const float* Wre, const float* Wim, const float* bre, const float* bim, float* Wp, float* bp, int D, int L,
                        int NK, int ncol, cudaStream_t s) {
  const int64_t n = (int64_t)(D + 1) * ncol;
  orb_permute_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(Wre, Wim, bre, bim, Wp, bp, D, L, NK, ncol, orb_per_tile(L));
  return (int)cudaGetLastError();
}

int orbital_contract(const float* c, const float* x, const double* normfac, float* Mj, int64_t B, TailDims d,
                     cudaStream_t s) {
  // (the value-only kernel measured 1 % slower with the same prefetch: its prologue is one warp's, not a block's)
  static const int prefetch = !(dbg_env("DH_ORB_PREFETCH") && atoi(dbg_env("DH_ORB_PREFETCH")) == 0);
  if (d.R == 1) {
    const int64_t rows = B * d.N;
    orbital_value_kernel<<<(unsigned)((rows + 7) / 8), 256, 8 * d.L * sizeof(cplx), s>>>(c, x, normfac, Mj, rows, d);
    return (int)cudaGetLastError();
  }
  size_t smem = 2 * d.L * sizeof(dcplx) + (size_t)ENV_SLOTS * d.L * sizeof(cplx);
  static const bool scalar_form = dbg_env("DH_ORB_CONTRACT") && strcmp(dbg_env("DH_ORB_CONTRACT"), "scalar") == 0;
  const int NK = d.N * d.K;
  const size_t smem_v = smem + (size_t)(d.R + 14) * NK * 2 * sizeof(float);
  if (!scalar_form && (reinterpret_cast<uintptr_t>(c) & 15) == 0 && smem_v <= 48 * 1024) {
    if (NK % 4 == 0) {
      orbital_contract_vec_kernel<4><<<(unsigned)(B * d.N), OCV_THREADS, smem_v, s>>>(c, x, normfac, Mj, d, prefetch);
      return (int)cudaGetLastError();
    }
    if (NK % 2 == 0) {
      orbital_contract_vec_kernel<2><<<(unsigned)(B * d.N), OCV_THREADS, smem_v, s>>>(c, x, normfac, Mj, d, prefetch);
      return (int)cudaGetLastError();
    }
  }
  orbital_contract_kernel<<<(unsigned)(B * d.N), 128, smem, s>>>(c, x, normfac, Mj, d);
  return (int)cudaGetLastError();
}

// =============================================================================================
// Warp-level Gauss-Jordan with partial pivoting on A[n][ncols] (complex, shared memory, row
// stride `ld`).  The leading n x n block is reduced to identity; if ncols == 2n and the right
// block started as I, it ends as A^-1.  Returns log|det| and the unit-modulus phase
// (jax slogdet semantics: singular -> (-inf, 0)).
// =============================================================================================
__device__ void warp_gauss_jordan(cplx* A, int n, int ncols, int ld, float& logabs, cplx& phase) {
  const int lane = threadIdx.x & 31;
  float la = 0.f;
  cplx ph = cmake(1.f, 0.f);
  bool singular = false;
  for (int p = 0; p < n; ++p) {
    // pivot search over rows p..n-1 of column p
    float best = -1.f;
    int bi = p;
    for (int i = p + lane; i < n; i += 32) {
      float m = cabs2(A[i * ld + p]);
      if (m > best) { best = m; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (bi != p) {
      for (int cidx = lane; cidx < ncols; cidx += 32) {
        cplx t = A[p * ld + cidx];
        A[p * ld + cidx] = A[bi * ld + cidx];
        A[bi * ld + cidx] = t;
      }
      ph = cmake(-ph.x, -ph.y);
    }
    __syncwarp();
    const cplx d = A[p * ld + p];
    const float ad = hypotf(d.x, d.y);
    la += logf(ad);
    if (ad == 0.f && !isnan(la)) singular = true;  // first exact-zero pivot: jax slogdet returns (0, -inf)
    ph = ad > 0.f ? cmul(ph, cmake(d.x / ad, d.y / ad)) : cmake(0.f, 0.f);
    const cplx dinv = cinv(d);
    __syncwarp();
    for (int cidx = lane; cidx < ncols; cidx += 32) A[p * ld + cidx] = cmul(A[p * ld + cidx], dinv);
    __syncwarp();
    for (int i = 0; i < n; ++i) {
      if (i == p) continue;
      const cplx f = A[i * ld + p];
      __syncwarp();
      for (int cidx = lane; cidx < ncols; cidx += 32) {
        cplx pv = A[p * ld + cidx];
        cplx cur = A[i * ld + cidx];
        A[i * ld + cidx] = cmake(cur.x - (f.x * pv.x - f.y * pv.y), cur.y - (f.x * pv.y + f.y * pv.x));
      }
      __syncwarp();
    }
  }
  // renormalise the phase (product of n unit numbers drifts by O(n eps))
  const float pn = hypotf(ph.x, ph.y);
  if (pn > 0.f) ph = cmake(ph.x / pn, ph.y / pn);
  if (singular) { la = -INFINITY; ph = cmake(0.f, 0.f); }
  logabs = la;
  phase = ph;
}

// Mj[b][kd][r][N][N] -> ld[b][kd][r] complex jets of log det ; also writes the inverse of the value
// matrix to Minv (optional, [b][kd][N][N]) for the VJP.
__global__ void __launch_bounds__(128)
logdet_jets_kernel(const float* __restrict__ Mj, float* __restrict__ ldout, float* __restrict__ Minv_out,
                   TailDims dm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = dm.N, R = dm.R;
  const int NN = N * N;
  cplx* aug = reinterpret_cast<cplx*>(smraw);        // [N][2N]
  cplx* wbuf = aug + (size_t)N * 2 * N;              // [nwarp][2][N][N]
  cplx* trsq = wbuf + (size_t)(blockDim.x >> 5) * 2 * NN;  // [R]
  cplx* trv = trsq + R;                              // [R]
  const int64_t bk = blockIdx.x;
  const cplx* M0 = reinterpret_cast<const cplx*>(Mj) + bk * R * NN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  for (int t = tid; t < NN; t += blockDim.x) {
    const int i = t / N, j = t % N;
    aug[i * 2 * N + j] = M0[t];
    aug[i * 2 * N + N + j] = cmake(i == j ? 1.f : 0.f, 0.f);
  }
  __syncthreads();
  if (warp == 0) {
    float la;
    cplx ph;
    warp_gauss_jordan(aug, N, 2 * N, 2 * N, la, ph);
    if (lane == 0) {
      float* o = ldout + bk * R * 2;
      o[0] = la;
      o[1] = atan2f(ph.y, ph.x);
      if (ph.x == 0.f && ph.y == 0.f) o[1] = 0.f;
    }
  }
  __syncthreads();
  if (Minv_out != nullptr) {
    cplx* mo = reinterpret_cast<cplx*>(Minv_out) + bk * NN;
    for (int t = tid; t < NN; t += blockDim.x) mo[t] = aug[(t / N) * 2 * N + N + (t % N)];
  }
  if (R == 1) return;
  Rows rw(N, true);
  cplx* Ms = wbuf + (size_t)warp * 2 * NN;
  cplx* Xs = Ms + NN;
  for (int r = 1 + warp; r < R; r += nwarp) {
    const cplx* Mr = M0 + (size_t)r * NN;
    for (int t = lane; t < NN; t += 32) Ms[t] = Mr[t];
    __syncwarp();
    for (int t = lane; t < NN; t += 32) {
      const int i = t / N, j = t % N;
      cplx s = cmake(0.f, 0.f);
      for (int l = 0; l < N; ++l) s = cfma(aug[i * 2 * N + N + l], Ms[l * N + j], s);
      Xs[t] = s;
    }
    __syncwarp();
    float tx = 0.f, ty = 0.f, qx = 0.f, qy = 0.f;
    for (int t = lane; t < NN; t += 32) {
      const int i = t / N, j = t % N;
      cplx a = Xs[t];
      if (i == j) { tx += a.x; ty += a.y; }
      cplx pr = cmul(a, Xs[j * N + i]);
      qx += pr.x; qy += pr.y;
    }
    tx = warp_sum(tx); ty = warp_sum(ty); qx = warp_sum(qx); qy = warp_sum(qy);
    if (lane == 0) { trv[r] = cmake(tx, ty); trsq[r] = cmake(qx, qy); }
    __syncwarp();
  }
  __syncthreads();
  for (int r = 1 + tid; r < R; r += blockDim.x) {
    cplx val = trv[r];
    if (r == rw.S()) {
      for (int k = 0; k < 2 * N; ++k) val = csub(val, trsq[rw.J(k)]);
    } else if (r >= rw.T(0)) {
      val = csub(val, trsq[rw.D(r - rw.T(0))]);
    }
    float* o = ldout + (bk * R + r) * 2;
    o[0] = val.x;
    o[1] = val.y;
  }
}

// Register-resident LU with partial pivoting for n <= 16: lane i holds row i (16 complex registers), pivots are
// found by a warp arg-max, rows swapped and the pivot row broadcast by shuffles.  Returns log|det| and the phase
// with jax slogdet semantics (singular -> (-inf, 0)).  All loops are fully unrolled (no dynamic register indexing).
__device__ __forceinline__ void warp_lu_regs(cplx (&a)[16], int n, float& logabs, cplx& phase) {
  const int lane = threadIdx.x & 31;
  float la = 0.f;
  cplx ph = cmake(1.f, 0.f);
  bool singular = false;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    if (p < n) {
      // pivot: largest |a[i][p]| over rows i >= p (ties -> smallest row)
      float best = (lane >= p && lane < n) ? cabs2(a[p]) : -1.f;
      int bi = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (bi != p) {
        const int src = lane == p ? bi : (lane == bi ? p : lane);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          a[c].x = __shfl_sync(0xffffffffu, a[c].x, src);
          a[c].y = __shfl_sync(0xffffffffu, a[c].y, src);
        }
        ph = cmake(-ph.x, -ph.y);
      }
      const cplx d = cmake(__shfl_sync(0xffffffffu, a[p].x, p), __shfl_sync(0xffffffffu, a[p].y, p));
      const float ad = hypotf(d.x, d.y);
      la += logf(ad);
      if (ad == 0.f && !isnan(la)) singular = true;
      ph = ad > 0.f ? cmul(ph, cmake(d.x / ad, d.y / ad)) : cmake(0.f, 0.f);
      const cplx f = (lane > p && lane < n) ? cmul(a[p], cinv(d)) : cmake(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (c > p) {
          const cplx pv = cmake(__shfl_sync(0xffffffffu, a[c].x, p), __shfl_sync(0xffffffffu, a[c].y, p));
          a[c] = cmake(a[c].x - (f.x * pv.x - f.y * pv.y), a[c].y - (f.x * pv.y + f.y * pv.x));
        }
      }
    }
  }
  const float pn = hypotf(ph.x, ph.y);
  if (pn > 0.f) ph = cmake(ph.x / pn, ph.y / pn);
  if (singular) { la = -INFINITY; ph = cmake(0.f, 0.f); }
  logabs = la;
  phase = ph;
}

// value-only form without the inverse (log psi / Metropolis passes): one warp per matrix, eight per block
__global__ void __launch_bounds__(256)
logdet_value_kernel(const float* __restrict__ Mj, float* __restrict__ ldout, int64_t nmat, int N) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t bk = (int64_t)blockIdx.x * 8 + warp;
  if (bk >= nmat) return;  // whole warps leave; only warp-level primitives below
  const cplx* M0 = reinterpret_cast<const cplx*>(Mj) + bk * N * N;
  cplx a[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) a[c] = (lane < N && c < N) ? M0[lane * N + c] : cmake(0.f, 0.f);
  float la;
  cplx ph;
  warp_lu_regs(a, N, la, ph);
  if (lane == 0) {
    float* o = ldout + bk * 2;
    o[0] = la;
    o[1] = (ph.x == 0.f && ph.y == 0.f) ? 0.f : atan2f(ph.y, ph.x);
  }
}

int logdet_jets_impl(const float* Mj, float* ld, float* Minv, int64_t B, TailDims d, cudaStream_t s) {
  if (d.R == 1 && Minv == nullptr) {
    const int64_t nmat = B * d.K;
    if (d.N > 16) return -2;
    logdet_value_kernel<<<(unsigned)((nmat + 7) / 8), 256, 0, s>>>(Mj, ld, nmat, d.N);
    return (int)cudaGetLastError();
  }
  const int nwarp = 4;
  size_t smem = ((size_t)d.N * 2 * d.N + (size_t)nwarp * 2 * d.N * d.N + 2 * (size_t)d.R) * sizeof(cplx);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(logdet_jets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  logdet_jets_kernel<<<(unsigned)(B * d.K), nwarp * 32, smem, s>>>(Mj, ld, Minv, d);
  return (int)cudaGetLastError();
}
int logdet_jets(const float* Mj, float* ld, int64_t B, TailDims d, cudaStream_t s) {
  return logdet_jets_impl(Mj, ld, nullptr, B, d, s);
}

// =============================================================================================
// finalize: one warp per walker.  log-sum-exp over determinants (with jets), Jastrow jets,
// potential, local-energy assembly.
// =============================================================================================
struct PairTerms {  // per-walker pair sums computed by pair_terms()
  float jas, jasS, coul, harm;
};

// lane i < N handles electron i; returns warp-reduced sums and the per-lane J rows of the Jastrow.
// Jastrow (blocks.py:91-105): parallel-spin pairs -a^2/4/(a + r) with a = ee_par, anti-parallel pairs
// -a^2/2/(a + r) with a = ee_anti; a class without pairs has no parameter (pointer null, terms skipped).
__device__ inline PairTerms pair_terms(const float* __restrict__ xw, int N, int n_up, const float* __restrict__ ee_par,
                                       const float* __restrict__ ee_anti, float Q, float& jJ0, float& jJ1) {
  const int lane = threadIdx.x & 31;
  float jas = 0.f, jasS = 0.f, coul = 0.f, harm = 0.f;
  jJ0 = 0.f; jJ1 = 0.f;
  const float a_par = ee_par ? ee_par[0] : 0.f, a_anti = ee_anti ? ee_anti[0] : 0.f;
  // every lane evaluates the trigonometry of ITS electron once; the partners' unit vectors come by shuffle
  float st = 0.f, ct = 1.f, sp = 0.f, cp = 1.f;
  if (lane < N) {
    sincosf(xw[lane * 2], &st, &ct);
    sincosf(xw[lane * 2 + 1], &sp, &cp);
  }
  const float rx = st * cp, ry = st * sp, rz = ct;
  // tangent-flow velocities of r_i: theta_hat x r = -phi_hat ; phi_hat x r = theta_hat
  const float w0x = sp, w0y = -cp, w0z = 0.f;
  const float w1x = ct * cp, w1y = ct * sp, w1z = -st;
  for (int j = 0; j < N; ++j) {
    const float qx = __shfl_sync(0xffffffffu, rx, j), qy = __shfl_sync(0xffffffffu, ry, j), qz = __shfl_sync(0xffffffffu, rz, j);
    if (lane >= N || j == lane) continue;
    const float dx = rx - qx, dy = ry - qy, dz = rz - qz;
    const float r2 = dx * dx + dy * dy + dz * dz;
    const float r = sqrtf(r2);
    const float cth = 1.f - 0.5f * r2;  // cos(theta_12)
    const bool par = (lane < n_up) == (j < n_up);
    if (par ? ee_par != nullptr : ee_anti != nullptr) {
      // f(r) = -w a^2/(a+r), w = 1/4 (parallel) or 1/2 (anti-parallel) ;  c = r_i.r_j ; dr/dc = -1/r ; d2r/dc2 = -1/r^3
      const float alpha = par ? a_par : a_anti, w = par ? 0.25f : 0.5f;
      const float ar = alpha + r;
      const float fr = w * alpha * alpha / (ar * ar);                // df/dr
      const float frr = -2.f * w * alpha * alpha / (ar * ar * ar);   // d2f/dr2
      const float fc = -fr / r;                                      // df/dc
      const float fcc = frr / r2 - fr / (r2 * r);                    // d2f/dc2
      jJ0 = fmaf(fc, w0x * qx + w0y * qy + w0z * qz, jJ0);
      jJ1 = fmaf(fc, w1x * qx + w1y * qy + w1z * qz, jJ1);
      if (j > lane) {
        jas += -w * alpha * alpha / ar;
        jasS += fc * (-4.f * cth) + fcc * 2.f * (1.f - cth * cth);
      }
    }
    if (j > lane) {
      coul += 1.f / r;
      harm += 1.f + (Q + 1.f) / Q * cth;
    }
  }
  PairTerms out;
  out.jas = warp_sum(jas);
  out.jasS = warp_sum(jasS);
  out.coul = warp_sum(coul);
  out.harm = warp_sum(harm);
  return out;
}

__global__ void __launch_bounds__(128)
finalize_kernel(FinalizeArgs a, int64_t B, TailDims dm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = dm.N, R = dm.R, K = dm.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= B) return;
  cplx* lp = reinterpret_cast<cplx*>(smraw) + (size_t)warp * (R + 2 * K);  // [R] log psi jets
  cplx* wk = lp + R;                                                      // [K] determinant weights
  const cplx* ld = reinterpret_cast<const cplx*>(a.ld) + b * K * R;
  if (a.Mj_value != nullptr) {  // value-only pass: log-determinants of the K orbital matrices taken here (logdet_value_kernel's arithmetic)
    cplx* lds = wk + K;
    for (int k = 0; k < K; ++k) {
      const cplx* M0 = reinterpret_cast<const cplx*>(a.Mj_value) + (b * K + k) * N * N;
      cplx m[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) m[c] = (lane < N && c < N) ? M0[lane * N + c] : cmake(0.f, 0.f);
      float la;
      cplx ph;
      warp_lu_regs(m, N, la, ph);
      if (lane == 0) lds[k] = cmake(la, (ph.x == 0.f && ph.y == 0.f) ? 0.f : atan2f(ph.y, ph.x));
    }
    __syncwarp();
    ld = lds;
  }
  // ---- value: log sum_k exp(ld_k)
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, ld[k * R].x);
  float sr = 0.f, si = 0.f;
  for (int k = 0; k < K; ++k) {
    float e = expf(ld[k * R].x - mx), s_, c_;
    sincosf(ld[k * R].y, &s_, &c_);
    sr += e * c_; si += e * s_;
  }
  cplx lp0;
  if (K == 1) lp0 = ld[0];
  else lp0 = cmake(logf(hypotf(sr, si)) + mx, atan2f(si, sr));
  if (lane < K || K > 32) {
    for (int k = lane; k < K; k += 32) {
      float e = expf(ld[k * R].x - lp0.x), s_, c_;
      sincosf(ld[k * R].y - lp0.y, &s_, &c_);
      wk[k] = cmake(e * c_, e * s_);
    }
  }
  __syncwarp();
  float jJ0, jJ1;
  PairTerms pt = pair_terms(a.x + b * N * 2, N, dm.n_up, a.ee_par, a.ee_anti, a.Q, jJ0, jJ1);
  const float pot_raw = a.interaction_type == 0 ? pt.coul / a.radius : pt.harm;
  lp0.x += pt.jas;
  if (lane == 0 && a.out_logpsi) { a.out_logpsi[b * 2] = lp0.x; a.out_logpsi[b * 2 + 1] = lp0.y; }
  if (R == 1) {
    if (lane == 0 && a.out_pot) a.out_pot[b] = pot_raw * a.interaction_strength;
    if (a.mv_x1 != nullptr) {  // Metropolis accept / select of this walker's move (mcmc_accept_kernel's arithmetic)
      int acc = 0;
      if (lane == 0) {
        const float logu = accept_log_uniform(a.mv_dv->seed, a.mv_dv->offset, a.mv_dv->subseq0 + (uint64_t)(a.mv_walker0 + b));
        const float l2 = 2.0f * lp0.x, l1 = a.mv_lp1[b];
        acc = (l2 - l1) > logu ? 1 : 0;  // NaN -> false (mcmc.py:59)
        if (acc) {
          a.mv_lp1[b] = l2;
          atomicAdd(&a.mv_dv->naccept, 1ull);
        }
      }
      acc = __shfl_sync(0xffffffffu, acc, 0);
      if (acc)
        for (int q = lane; q < 2 * N; q += 32) a.mv_x1[b * 2 * N + q] = a.x[b * 2 * N + q];
    }
    return;
  }
  Rows rw(N, true);
  // ---- first-order rows + their second-order corrections
  for (int r = 1 + lane; r < R; r += 32) {
    if (r == rw.S() || r >= rw.T(0)) continue;
    cplx v = cmake(0.f, 0.f), sq = cmake(0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      cplx l = ld[k * R + r];
      v = cfma(wk[k], l, v);
      sq = cfma(wk[k], cmul(l, l), sq);
    }
    // extra = sum_k w_k l_k^2 - (sum_k w_k l_k)^2   (Hessian of log-sum-exp)
    cplx extra = csub(sq, cmul(v, v));
    lp[r] = v;
    // stash the correction in the slot of the matching second-order row via shared scratch:
    // J rows accumulate into lp[S] later, D_a into lp[T_a]; keep it in registers for now.
    // (re-computed below to avoid extra shared arrays)
    (void)extra;
  }
  __syncwarp();
  // second-order rows (lanes 0..3): S, T_x, T_y, T_z
  if (lane < 4) {
    const int r = lane == 0 ? rw.S() : rw.T(lane - 1);
    cplx v = cmake(0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      cplx acc = ld[k * R + r];
      if (K > 1) {
        if (lane == 0) {
          for (int q = 0; q < 2 * N; ++q) { cplx l = ld[k * R + rw.J(q)]; acc = cadd(acc, cmul(l, l)); }
        } else {
          cplx l = ld[k * R + rw.D(lane - 1)];
          acc = cadd(acc, cmul(l, l));
        }
      }
      v = cfma(wk[k], acc, v);
    }
    if (K > 1) {
      if (lane == 0) {
        for (int q = 0; q < 2 * N; ++q) { cplx l = lp[rw.J(q)]; v = csub(v, cmul(l, l)); }
      } else {
        cplx l = lp[rw.D(lane - 1)];
        v = csub(v, cmul(l, l));
      }
    }
    lp[r] = v;
  }
  __syncwarp();
  // ---- add the Jastrow jets (real): J rows of own electron, S row; D/T rows vanish
  if ((a.ee_par != nullptr || a.ee_anti != nullptr) && lane < N) {
    lp[rw.J(2 * lane)].x += jJ0;
    lp[rw.J(2 * lane + 1)].x += jJ1;
  }
  if (lane == 0) lp[rw.S()].x += pt.jasS;
  if (lane == 0) lp[0] = lp0;
  __syncwarp();
  if (a.lpjet) {
    for (int r = lane; r < R; r += 32) { a.lpjet[(b * R + r) * 2] = lp[r].x; a.lpjet[(b * R + r) * 2 + 1] = lp[r].y; }
  }
  // ---- assemble
  float jx = 0.f, jy = 0.f;
  for (int q = lane; q < 2 * N; q += 32) { cplx l = lp[rw.J(q)]; cplx s2 = cmul(l, l); jx += s2.x; jy += s2.y; }
  jx = warp_sum(jx); jy = warp_sum(jy);
  if (lane == 0) {
    const float inv2r2 = 0.5f / (a.radius * a.radius);
    const cplx S = lp[rw.S()];
    const cplx kin = cmake(-(S.x + jx) * inv2r2, -(S.y + jy) * inv2r2);
    const float pot = pot_raw * a.interaction_strength;
    float l2 = 0.f, lz2 = 0.f;
    for (int a3 = 0; a3 < 3; ++a3) {
      cplx Dv = lp[rw.D(a3)], Tv = lp[rw.T(a3)];
      float term = -(Tv.x + (Dv.x * Dv.x - Dv.y * Dv.y));
      l2 += term;
      if (a3 == 2) lz2 = term;
    }
    if (a.out_kin) { a.out_kin[b * 2] = kin.x; a.out_kin[b * 2 + 1] = kin.y; }
    if (a.out_el) { a.out_el[b * 2] = kin.x + pot; a.out_el[b * 2 + 1] = kin.y; }
    if (a.out_pot) a.out_pot[b] = pot;
    if (a.out_lz) a.out_lz[b] = lp[rw.D(2)].y;
    if (a.out_lz2) a.out_lz2[b] = lz2;
    if (a.out_l2) a.out_l2[b] = l2;
  }
}

int finalize(FinalizeArgs a, int64_t B, TailDims d, cudaStream_t s) {
  if (d.N > 32) return -2;
  const int wpb = 4;
  if (a.Mj_value != nullptr && (d.R != 1 || d.N > 16)) return -2;
  if (a.mv_x1 != nullptr && (d.R != 1 || !a.mv_lp1 || !a.mv_dv)) return -2;
  size_t smem = (size_t)wpb * (d.R + 2 * d.K) * sizeof(cplx);
  finalize_kernel<<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, smem, s>>>(a, B, d);
  return (int)cudaGetLastError();
}

__global__ void potential_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t B, int N,
                                 float Q, float radius, int itype) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= B) return;
  float j0, j1;
  PairTerms pt = pair_terms(x + b * N * 2, N, N, nullptr, nullptr, Q, j0, j1);
  if (lane == 0) out[b] = itype == 0 ? pt.coul / radius : pt.harm;
}

int potential(const float* x, float* out, int64_t B, int N, float Q, float radius, int interaction_type,
              cudaStream_t s) {
  if (N > 32) return -2;
  const int wpb = 4;
  potential_kernel<<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, s>>>(x, out, B, N, Q, radius, interaction_type);
  return (int)cudaGetLastError();
}

// =============================================================================================
// Public batched slogdet (+ multi-determinant tail): one warp per walker, K matrices in turn.
// =============================================================================================
__global__ void __launch_bounds__(128)
slogdet_kernel(const float* __restrict__ mats, int64_t B, int K, int n, float* __restrict__ out_sign,
               float* __restrict__ out_logabs, float* __restrict__ out_logpsi) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= B) return;
  cplx* A = reinterpret_cast<cplx*>(smraw) + (size_t)warp * n * n;
  float mx = -INFINITY, sr = 0.f, si = 0.f;
  for (int k = 0; k < K; ++k) {
    const cplx* src = reinterpret_cast<const cplx*>(mats) + (b * K + k) * n * n;
    float la;
    cplx ph;
    if (n <= 16) {  // register-resident LU
      cplx a[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) a[c] = (lane < n && c < n) ? src[lane * n + c] : cmake(0.f, 0.f);
      warp_lu_regs(a, n, la, ph);
    } else {
      for (int t = lane; t < n * n; t += 32) A[t] = src[t];
      __syncwarp();
      warp_gauss_jordan(A, n, n, n, la, ph);
    }
    if (lane == 0) {
      if (out_sign) { out_sign[(b * K + k) * 2] = ph.x; out_sign[(b * K + k) * 2 + 1] = ph.y; }
      if (out_logabs) out_logabs[b * K + k] = la;
    }
    // running log-sum-exp
    if (la > mx) {
      const float sc = expf(mx - la);  // exp(-inf) = 0 on the first pass
      sr *= sc; si *= sc; mx = la;
    }
    if (la != -INFINITY) {
      const float e = expf(la - mx);
      sr += e * ph.x; si += e * ph.y;
    }
    __syncwarp();
  }
  if (lane == 0 && out_logpsi) {
    out_logpsi[b * 2] = logf(hypotf(sr, si)) + mx;
    out_logpsi[b * 2 + 1] = atan2f(si, sr);
  }
}

int slogdet_batched(const float* mats, int64_t B, int K, int n, float* out_sign, float* out_logabs,
                    float* out_logpsi, cudaStream_t s) {
  if (n < 1 || n > 64 || K < 1) return -1;
  const int wpb = 4;
  size_t smem = (size_t)wpb * n * n * sizeof(cplx);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(slogdet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  slogdet_kernel<<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, smem, s>>>(mats, B, K, n, out_sign, out_logabs, out_logpsi);
  return (int)cudaGetLastError();
}

}  // namespace dh
