// Walker-ensemble estimators (the reference's netobs_bridge/observables): pair correlation g(theta_12), polar
// density histogram, and the overlap ratio against a second wavefunction.  All are reductions over the walker batch:
// HBM-bound reads of the (B, N, 2) coordinates (or of two (B) complex log-amplitudes), histogram / sum accumulators in
// shared memory, one global atomic per bin and block.
//
//   dh_pair_correlation   <- PairCorrelationEstimator.evaluate   netobs_bridge/observables/pair_corr.py:42-60
//   dh_density_histogram  <- DensityEstimator.evaluate           netobs_bridge/observables/density.py:42-49
//   dh_overlap_sum / dh_overlap_ratio <- OverlapEstimator.evaluate  netobs_bridge/observables/overlap.py:55-63
#include <math.h>

#include "../../include/deephall_b200.h"
#include "kernels.h"

namespace dh {

constexpr int OBS_THREADS = 256;
constexpr int OBS_MAX_BINS = 4096;

// theta_12 = arccos(r_i . r_j) of every pair i < j of every walker, weight 1 / sin(theta_12), `bins` equal bins over
// [0, pi] (numpy / jnp.histogram: the last bin is closed on the right).  The geometry runs in fp64 from the fp32
// coordinates, so that bin decisions do not depend on fp32 rounding of arccos; accumulation in fp64.
__global__ void __launch_bounds__(OBS_THREADS)
pair_correlation_kernel(const float* __restrict__ x, int64_t B, int N, int bins, double* __restrict__ hist) {
  extern __shared__ double sh[];
  for (int t = threadIdx.x; t < bins; t += blockDim.x) sh[t] = 0.0;
  __syncthreads();
  const int npair = N * (N - 1) / 2;
  const int64_t items = B * npair;
  const double scale = (double)bins / M_PI;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = it / npair;
    int q = (int)(it - b * npair);
    int i = 0;
    while (q >= N - 1 - i) { q -= N - 1 - i; ++i; }  // triu_indices(N, 1) order: (0,1) (0,2) .. (1,2) ..
    const int j = i + 1 + q;
    const float* xw = x + b * N * 2;
    double si, ci, spi, cpi, sj, cj, spj, cpj;
    sincos((double)xw[2 * i], &si, &ci);
    sincos((double)xw[2 * i + 1], &spi, &cpi);
    sincos((double)xw[2 * j], &sj, &cj);
    sincos((double)xw[2 * j + 1], &spj, &cpj);
    double c12 = si * cpi * sj * cpj + si * spi * sj * spj + ci * cj;
    c12 = fmin(1.0, fmax(-1.0, c12));
    const double th = acos(c12);
    int bin = (int)floor(th * scale);
    if (bin >= bins) bin = bins - 1;  // th == pi belongs to the last bin
    const double w = 1.0 / sin(th);
    atomicAdd(&sh[bin], w);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < bins; t += blockDim.x)
    if (sh[t] != 0.0) atomicAdd(&hist[t], sh[t]);
}

// state += hist * 4 bins / (B N^2 pi)   (pair_corr.py:59)
__global__ void pair_correlation_scale_kernel(const double* __restrict__ hist, int bins, double factor,
                                              float* __restrict__ state_inout) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < bins) state_inout[t] = (float)((double)state_inout[t] + hist[t] * factor);
}

// counts of theta over every electron of every walker, `bins` equal bins over [0, pi] (density.py:46-48); integer
// accumulation: exact and order-independent.  bin = floor(theta * bins / pi) in fp64, theta == pi in the last bin.
__global__ void __launch_bounds__(OBS_THREADS)
density_histogram_kernel(const float* __restrict__ x, int64_t n_elec, int bins, unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned int shc[];
  for (int t = threadIdx.x; t < bins; t += blockDim.x) shc[t] = 0u;
  __syncthreads();
  const double scale = (double)bins / M_PI;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_elec; it += (int64_t)gridDim.x * blockDim.x) {
    const double th = (double)x[2 * it];
    if (!(th >= 0.0) || th > M_PI) continue;  // outside the range (or NaN): not counted
    int bin = (int)floor(th * scale);
    if (bin >= bins) bin = bins - 1;
    atomicAdd(&shc[bin], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < bins; t += blockDim.x)
    if (shc[t]) atomicAdd(&counts[t], (unsigned long long)shc[t]);
}

// sum_b (logphi_b - logpsi_b), complex, fp64, one block and a fixed summation order (deterministic)
__global__ void __launch_bounds__(1024)
overlap_sum_kernel(const float* __restrict__ logphi, const float* __restrict__ logpsi, int64_t B, double* __restrict__ out) {
  __shared__ double sr[1024], si[1024];
  double ar = 0.0, ai = 0.0;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    ar += (double)logphi[2 * b] - (double)logpsi[2 * b];
    ai += (double)logphi[2 * b + 1] - (double)logpsi[2 * b + 1];
  }
  sr[threadIdx.x] = ar; si[threadIdx.x] = ai;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { sr[threadIdx.x] += sr[threadIdx.x + s]; si[threadIdx.x] += si[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sr[0]; out[1] = si[0]; }
}

// ratio_b = exp(logphi_b - logpsi_b - shift), ratio_square_b = |ratio_b|^2   (overlap.py:60-62)
__global__ void overlap_ratio_kernel(const float* __restrict__ logphi, const float* __restrict__ logpsi, int64_t B,
                                     const double* __restrict__ shift, float* __restrict__ ratio,
                                     float* __restrict__ ratio_sq) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double dr = (double)logphi[2 * b] - (double)logpsi[2 * b] - shift[0];
  const double di = (double)logphi[2 * b + 1] - (double)logpsi[2 * b + 1] - shift[1];
  const double m = exp(dr);
  double s, c;
  sincos(di, &s, &c);
  if (ratio) { ratio[2 * b] = (float)(m * c); ratio[2 * b + 1] = (float)(m * s); }
  if (ratio_sq) ratio_sq[b] = (float)(m * m);
}

// ------------------------------------------------------------------ one-body reduced density matrix (one_rdm.py)
// Y_{Q,Q,m}(theta, phi), m = -Q .. Q, as make_monopole_harm(Q, Q, m) builds it (one_rdm.py:34-58; with l = q the sum
// has the single term s = 0): sqrt((2Q+1)/(4 pi) C(2Q, Q-m)^-1 ... ) folded into
//   Y = sqrt((2Q+1)/(4 pi) * (Q-m)! (Q+m)! / (2Q)!) / 2^Q * (-1)^(Q-m) C(2Q, Q-m) (1-x)^((Q-m)/2) (1+x)^((Q+m)/2) e^{i m phi},
// x = clip(cos theta, -1 + 1e-4, 1 - 1e-4); fp64.
// One block = 8 points x all orbitals.  Per block the L values 0.5 ln C(2Q, a) go into shared memory once; per point
// the fp64 transcendental work is one cos, two logs and one sincos (by lanes 0..7 of the block's first warp, one
// point each); per (point, orbital) it is one exp and one step of the phase table e^{i m phi} = e^{-i Q phi} (e^{i phi})^a,
// built per point by repeated complex multiplication in fp64 (2Q + 1 <= 256 steps).
constexpr int LLL_PTS = 8;
__global__ void __launch_bounds__(256)
lll_orbitals_kernel(const float* __restrict__ pts, int64_t n, int twoQ, float* __restrict__ out) {
  extern __shared__ double shl[];  // lc[L] | per point: l1, l2 | phase table [LLL_PTS][L][2]
  const int L = twoQ + 1;
  double* lc = shl;
  double* pl = lc + L;                  // [LLL_PTS][2]
  double* ph = pl + 2 * LLL_PTS;        // [LLL_PTS][L][2]
  const int64_t p0 = (int64_t)blockIdx.x * LLL_PTS;
  const int tid = threadIdx.x;
  for (int a = tid; a < L; a += blockDim.x)
    lc[a] = 0.5 * (lgamma((double)twoQ + 1.0) - lgamma((double)a + 1.0) - lgamma((double)(twoQ - a) + 1.0));
  if (tid < LLL_PTS && p0 + tid < n) {
    const double th = (double)pts[2 * (p0 + tid)], phi = (double)pts[2 * (p0 + tid) + 1];
    const double x = fmin(1.0 - 1e-4, fmax(-1.0 + 1e-4, cos(th)));  // one_rdm.py:49
    pl[2 * tid] = 0.5 * log(0.5 * (1.0 - x));
    pl[2 * tid + 1] = 0.5 * log(0.5 * (1.0 + x));
    double s1, c1, s0, c0;
    sincos(phi, &s1, &c1);
    sincos(-0.5 * twoQ * phi, &s0, &c0);
    double* t = ph + (size_t)tid * L * 2;
    for (int a = 0; a < L; ++a) {
      t[2 * a] = c0; t[2 * a + 1] = s0;
      const double cn = c0 * c1 - s0 * s1;
      s0 = c0 * s1 + s0 * c1;
      c0 = cn;
    }
  }
  __syncthreads();
  const double norm = sqrt((twoQ + 1.0) / (4.0 * M_PI));
  for (int t = tid; t < LLL_PTS * L; t += blockDim.x) {
    const int q = t / L, a = t - q * L;
    if (p0 + q >= n) break;
    const int bq = twoQ - a;
    // Y = sqrt((2Q+1)/(4 pi) C(2Q,a)) (-1)^(Q-m) ((1-x)/2)^((Q-m)/2) ((1+x)/2)^((Q+m)/2) e^{i m phi},  a = Q + m
    const double mag = norm * exp(lc[a] + bq * pl[2 * q] + a * pl[2 * q + 1]) * ((bq & 1) ? -1.0 : 1.0);
    const double* e = ph + ((size_t)q * L + a) * 2;
    float* o = out + ((p0 + q) * L + a) * 2;  // consecutive t -> consecutive addresses
    o[0] = (float)(mag * e[0]);
    o[1] = (float)(mag * e[1]);
  }
}

// x' (B, N, N, 2): copy a of walker b is the walker with electron a moved to r'_b (one_rdm.py:92-94)
__global__ void one_rdm_scatter_kernel(const float* __restrict__ x, const float* __restrict__ rp, int64_t B, int N,
                                       float* __restrict__ xp) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * N * N) return;
  const int e = (int)(t % N);
  const int a = (int)((t / N) % N);
  const int64_t b = t / ((int64_t)N * N);
  const float2 v = e == a ? reinterpret_cast<const float2*>(rp)[b] : reinterpret_cast<const float2*>(x)[b * N + e];
  reinterpret_cast<float2*>(xp)[t] = v;
}

// per walker: w_i = sum_a exp(logpsi'_a - logpsi) phi_i(r_a);  rdm_ij = 4 pi w_i conj(phi_j(r'))   (one_rdm.py:101-109)
// One block per walker; optional per-walker output (B, L, L) c64 and batch sum (L, L) complex fp64 (atomics).
__global__ void __launch_bounds__(256)
one_rdm_product_kernel(const float* __restrict__ logpsi, const float* __restrict__ logpsi_prime,
                       const float* __restrict__ phi, const float* __restrict__ phi_prime, int N, int L,
                       float* __restrict__ out, double* __restrict__ out_sum) {
  extern __shared__ double shw[];  // ratio [N][2] | w [L][2] | phi' [L][2]
  double* ratio = shw;
  double* w = ratio + 2 * N;
  double* pp = w + 2 * L;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid < N) {
    const double dr = (double)logpsi_prime[2 * (b * N + tid)] - (double)logpsi[2 * b];
    const double di = (double)logpsi_prime[2 * (b * N + tid) + 1] - (double)logpsi[2 * b + 1];
    const double mg = exp(dr);
    double sn, cs;
    sincos(di, &sn, &cs);
    ratio[2 * tid] = mg * cs;
    ratio[2 * tid + 1] = mg * sn;
  }
  for (int j = tid; j < L; j += blockDim.x) {
    pp[2 * j] = (double)phi_prime[2 * (b * L + j)];
    pp[2 * j + 1] = (double)phi_prime[2 * (b * L + j) + 1];
  }
  __syncthreads();
  for (int i = tid; i < L; i += blockDim.x) {
    double wr = 0.0, wi = 0.0;
    for (int a = 0; a < N; ++a) {
      const double pr = (double)phi[2 * ((b * N + a) * L + i)], pi = (double)phi[2 * ((b * N + a) * L + i) + 1];
      wr += ratio[2 * a] * pr - ratio[2 * a + 1] * pi;
      wi += ratio[2 * a] * pi + ratio[2 * a + 1] * pr;
    }
    w[2 * i] = 4.0 * M_PI * wr;
    w[2 * i + 1] = 4.0 * M_PI * wi;
  }
  __syncthreads();
  for (int t = tid; t < L * L; t += blockDim.x) {
    const int i = t / L, j = t % L;
    const double re = w[2 * i] * pp[2 * j] + w[2 * i + 1] * pp[2 * j + 1];   // w_i * conj(p_j)
    const double im = w[2 * i + 1] * pp[2 * j] - w[2 * i] * pp[2 * j + 1];
    if (out) { out[2 * (b * L * L + t)] = (float)re; out[2 * (b * L * L + t) + 1] = (float)im; }
    if (out_sum && !(isnan(re) || isnan(im))) { atomicAdd(&out_sum[2 * t], re); atomicAdd(&out_sum[2 * t + 1], im); }
  }
}

static inline int obs_grid(int64_t items) {
  int64_t g = (items + OBS_THREADS - 1) / OBS_THREADS;
  const int64_t cap = 148 * 8;  // eight resident blocks of 256 threads on each of the 148 SMs
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dh

extern "C" int dh_pair_correlation(const float* x, int64_t B, int32_t N, int32_t bins, int64_t batch_norm,
                                   float* state_inout, double* hist_ws, void* stream) {
  if (!x || !state_inout || !hist_ws || B < 0 || N < 1 || bins < 1 || bins > dh::OBS_MAX_BINS) return DH_E_BADARG;
  if (batch_norm <= 0) batch_norm = B;
  if (B == 0 || N < 2) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(hist_ws, 0, (size_t)bins * sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  const int64_t items = B * (N * (N - 1) / 2);
  dh::pair_correlation_kernel<<<dh::obs_grid(items), dh::OBS_THREADS, (size_t)bins * sizeof(double), s>>>(x, B, N, bins, hist_ws);
  const double factor = 4.0 * bins / ((double)batch_norm * N * N * M_PI);
  dh::pair_correlation_scale_kernel<<<(bins + 255) / 256, 256, 0, s>>>(hist_ws, bins, factor, state_inout);
  return (int)cudaGetLastError();
}

extern "C" int dh_density_histogram(const float* x, int64_t B, int32_t N, int32_t bins,
                                    unsigned long long* counts_inout, void* stream) {
  if (!x || !counts_inout || B < 0 || N < 1 || bins < 1 || bins > dh::OBS_MAX_BINS) return DH_E_BADARG;
  if (B == 0) return 0;
  dh::density_histogram_kernel<<<dh::obs_grid(B * N), dh::OBS_THREADS, (size_t)bins * sizeof(unsigned int), (cudaStream_t)stream>>>(
      x, B * N, bins, counts_inout);
  return (int)cudaGetLastError();
}

extern "C" int dh_overlap_sum(const float* logphi, const float* logpsi, int64_t B, double* out_sum, void* stream) {
  if (!logphi || !logpsi || !out_sum || B < 0) return DH_E_BADARG;
  dh::overlap_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logphi, logpsi, B, out_sum);
  return (int)cudaGetLastError();
}

extern "C" int dh_overlap_ratio(const float* logphi, const float* logpsi, int64_t B, const double* shift,
                                float* out_ratio, float* out_ratio_square, void* stream) {
  if (!logphi || !logpsi || !shift || B < 0) return DH_E_BADARG;
  if (B == 0) return 0;
  dh::overlap_ratio_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logphi, logpsi, B, shift, out_ratio,
                                                                                       out_ratio_square);
  return (int)cudaGetLastError();
}

extern "C" int dh_lll_orbitals(const float* points, int64_t n, int32_t flux, float* out_phi, void* stream) {
  if (!points || !out_phi || n < 0 || flux < 0 || flux > 255) return DH_E_BADARG;
  if (n == 0) return 0;
  const int L = flux + 1;
  const size_t smem = (size_t)(L + 2 * dh::LLL_PTS + 2 * dh::LLL_PTS * L) * sizeof(double);
  dh::lll_orbitals_kernel<<<(unsigned)((n + dh::LLL_PTS - 1) / dh::LLL_PTS), 256, smem, (cudaStream_t)stream>>>(points, n, flux, out_phi);
  return (int)cudaGetLastError();
}

extern "C" int dh_one_rdm_scatter(const float* x, const float* r_prime, int64_t B, int32_t N, float* out_x_prime,
                                  void* stream) {
  if (!x || !r_prime || !out_x_prime || B < 0 || N < 1) return DH_E_BADARG;
  if (B == 0) return 0;
  const int64_t items = B * N * N;
  dh::one_rdm_scatter_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, r_prime, B, N, out_x_prime);
  return (int)cudaGetLastError();
}

extern "C" int dh_one_rdm_product(const float* logpsi, const float* logpsi_prime, const float* phi,
                                  const float* phi_prime, int64_t B, int32_t N, int32_t L, float* out_rdm,
                                  double* out_sum_inout, void* stream) {
  if (!logpsi || !logpsi_prime || !phi || !phi_prime || B < 0 || N < 1 || N > 256 || L < 1 || L > 256) return DH_E_BADARG;
  if (B == 0 || (!out_rdm && !out_sum_inout)) return 0;
  const size_t smem = (size_t)(2 * N + 4 * L) * sizeof(double);
  dh::one_rdm_product_kernel<<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(logpsi, logpsi_prime, phi, phi_prime, N, L,
                                                                             out_rdm, out_sum_inout);
  return (int)cudaGetLastError();
}
