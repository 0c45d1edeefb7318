// Metropolis-Hastings walker update: on-sphere rotation proposal (mcmc.py:67-102), accept /
// select (mcmc.py:56-62), uniform initial walkers (train.py:40-54).  Random numbers come from an
// in-kernel Philox4x32-10 stream or, for bit-exact parity tests, from an injected array.
//
// Philox addressing: key = seed; counter = (offset * 64 + slot, subsequence0 + walker) where
// slot = electron index for the proposal draws and 63 for the accept draw; `offset` advances by
// one per MH move.
#include "kernels.h"

namespace dh {


// Device-resident arguments of a Metropolis move (McmcDev, kernels.h): when a kernel is given this block it reads the
// Philox key / offset / width from it instead of from its launch arguments, so ONE captured CUDA graph of a move
// (propose -> log psi -> accept -> advance) can be replayed for every move of every sweep.
__global__ void mcmc_dev_init_kernel(McmcDev* dv, unsigned long long seed, unsigned long long offset,
                                     unsigned long long subseq0, float width) {
  dv->seed = seed; dv->offset = offset; dv->subseq0 = subseq0; dv->naccept = 0ull; dv->width = width;
}
__global__ void mcmc_dev_advance_kernel(McmcDev* dv) { dv->offset += 1ull; }
int mcmc_dev_init(McmcDev* dv, uint64_t seed, uint64_t offset, uint64_t subseq0, float width, cudaStream_t s) {
  mcmc_dev_init_kernel<<<1, 1, 0, s>>>(dv, seed, offset, subseq0, width);
  return (int)cudaGetLastError();
}
// the same block filled from DEVICE scalars (traced values of a jit-compiled caller: the Philox key and the proposal width)
__global__ void mcmc_dev_init_from_kernel(McmcDev* dv, const unsigned long long* __restrict__ key, const float* __restrict__ width,
                                          unsigned long long subseq0) {
  dv->seed = key[0]; dv->offset = key[1]; dv->subseq0 = subseq0; dv->naccept = 0ull; dv->width = width[0];
}
int mcmc_dev_init_from(McmcDev* dv, const unsigned long long* key, const float* width, uint64_t subseq0, cudaStream_t s) {
  mcmc_dev_init_from_kernel<<<1, 1, 0, s>>>(dv, key, width, subseq0);
  return (int)cudaGetLastError();
}
int mcmc_dev_advance(McmcDev* dv, cudaStream_t s) {
  mcmc_dev_advance_kernel<<<1, 1, 0, s>>>(dv);
  return (int)cudaGetLastError();
}

__global__ void mcmc_propose_kernel(const float* __restrict__ x1, float* __restrict__ x2, int64_t total, int N,
                                    float width, uint64_t seed, uint64_t offset, uint64_t subseq0,
                                    const float* __restrict__ randoms, const McmcDev* __restrict__ dv) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  if (dv != nullptr) { seed = dv->seed; offset = dv->offset; subseq0 += dv->subseq0; width = dv->width; }
  const int64_t b = t / N;
  const int i = (int)(t % N);
  float nrm, uph;
  if (randoms != nullptr) {
    nrm = randoms[b * (2 * N + 1) + i];
    uph = randoms[b * (2 * N + 1) + N + i];
  } else {
    propose_draws(seed, offset, subseq0 + (uint64_t)b, i, nrm, uph);
  }
  float th2, ph2;
  propose_point(x1[t * 2], x1[t * 2 + 1], nrm, uph, width, th2, ph2);
  x2[t * 2] = th2;
  x2[t * 2 + 1] = ph2;
}

int mcmc_propose(const float* x1, float* x2, int64_t B, int N, float width, uint64_t seed, uint64_t offset,
                 uint64_t subseq0, const float* randoms, cudaStream_t s) {
  const int64_t total = B * N;
  mcmc_propose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x1, x2, total, N, width, seed, offset,
                                                                      subseq0, randoms, nullptr);
  return (int)cudaGetLastError();
}
int mcmc_propose_dev(const float* x1, float* x2, int64_t B, int N, const McmcDev* dv, int64_t walker0, cudaStream_t s) {
  const int64_t total = B * N;
  mcmc_propose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x1, x2, total, N, 0.f, 0, 0, (uint64_t)walker0, nullptr, dv);
  return (int)cudaGetLastError();
}

// lp2c: pointer to 2 Re log psi source; if lp2_stride == 2 it is the complex log psi array
// (lp = 2 * re), if 1 it is already lp.
__global__ void mcmc_accept_kernel(float* __restrict__ x1, const float* __restrict__ x2, float* __restrict__ lp1,
                                   const float* __restrict__ lp2c, int lp2_stride, int64_t B, int N, uint64_t seed,
                                   uint64_t offset, uint64_t subseq0, const float* __restrict__ randoms,
                                   unsigned long long* __restrict__ naccept, McmcDev* __restrict__ dv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool acc = false;
  if (dv != nullptr) { seed = dv->seed; offset = dv->offset; subseq0 += dv->subseq0; naccept = &dv->naccept; }
  if (b < B) {
    float logu;
    if (randoms != nullptr) logu = (float)log((double)randoms[b * (2 * N + 1) + 2 * N]);  // one rounding, as oracle.mcmc.log_uniform
    else logu = accept_log_uniform(seed, offset, subseq0 + (uint64_t)b);
    const float l2 = lp2_stride == 2 ? 2.0f * lp2c[b * 2] : lp2c[b];
    const float l1 = lp1[b];
    acc = (l2 - l1) > logu;  // NaN -> false (mcmc.py:59)
    if (acc) {
      lp1[b] = l2;
      for (int t = 0; t < 2 * N; ++t) x1[b * 2 * N + t] = x2[b * 2 * N + t];
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(naccept, (unsigned long long)__popc(m));
}

int mcmc_accept(float* x1, const float* x2, float* lp1, const float* lp2c, int lp2_stride, int64_t B, int N,
                uint64_t seed, uint64_t offset, uint64_t subseq0, const float* randoms,
                unsigned long long* naccept, cudaStream_t s) {
  mcmc_accept_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(x1, x2, lp1, lp2c, lp2_stride, B, N, seed, offset,
                                                                subseq0, randoms, naccept, nullptr);
  return (int)cudaGetLastError();
}
int mcmc_accept_dev(float* x1, const float* x2, float* lp1, const float* lp2c, int lp2_stride, int64_t B, int N, McmcDev* dv,
                    int64_t walker0, cudaStream_t s) {
  mcmc_accept_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(x1, x2, lp1, lp2c, lp2_stride, B, N, 0, 0, (uint64_t)walker0, nullptr, nullptr, dv);
  return (int)cudaGetLastError();
}

__global__ void init_walkers_kernel(float* __restrict__ x, int64_t total, int N, uint64_t seed, uint64_t subseq0) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t b = t / N;
  const int i = (int)(t % N);
  Philox ph(seed);
  // own counter domain (top bit set): the Metropolis moves use counters offset * 64 + slot < 2^63 on the same
  // (seed, walker), so the first proposal is independent of the walker's starting point
  uint4 r = ph((1ull << 63) | (uint64_t)i, subseq0 + (uint64_t)b);
  x[t * 2] = acosf(2.f * u01(r.x) - 1.f);
  x[t * 2 + 1] = (2.f * u01(r.y) - 1.f) * 3.14159265358979f;
}

int init_walkers(float* x, int64_t B, int N, uint64_t seed, uint64_t subseq0, cudaStream_t s) {
  const int64_t total = B * N;
  init_walkers_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, total, N, seed, subseq0);
  return (int)cudaGetLastError();
}

__global__ void lp_from_logpsi_kernel(const float* __restrict__ lc, float* __restrict__ lp, int64_t B) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) lp[b] = 2.0f * lc[b * 2];
}
int lp_from_logpsi(const float* logpsi_c, float* lp, int64_t B, cudaStream_t s) {
  lp_from_logpsi_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(logpsi_c, lp, B);
  return (int)cudaGetLastError();
}

}  // namespace dh
