// Energy statistics of a walker batch (deephall/loss.py:30-38,66-92) as two kernels instead of ~170 small tensor ops:
//   energy_stats : the rank-local means that the reference `pmean`s -- plain means of the five observables (loss.py:68-71),
//                  nanmean(E_L) (:73), nanmean(iqr_clip(E_L)) (:74), nanmean(Re E_L^2) (:91), and the clipped means of
//                  L_z^2, L_z, L^2 that the penalties use (:79,:80,:87) -- as ONE packed vector (one all-reduce);
//   energy_diff  : diff = iqr_clip(E_L - clipped [+ penalties]) (:75-89) and the cotangent (2 / n_valid) diff of the
//                  gradient's vector-Jacobian product (:60-64,99-106), NaN walkers masked.
// iqr_clip (:30-38) needs the 25 % / 75 % `nanquantile`s (linear interpolation) of the real and of the imaginary part: one
// block per array sorts its non-NaN values in shared memory (bitonic, NaN -> +inf goes to the end), so a batch of up to
// 32768 walkers per rank needs no global scratch.  Quantiles stay rank-local, as in the reference.
#include <math.h>

#include "kernels.h"

namespace dh {

namespace {

constexpr int ST_THREADS = 1024;

__device__ __forceinline__ float block_sum(float v, float* red) {
  __syncthreads();
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < ST_THREADS / 32 ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) { t = warp_sum(t); if (threadIdx.x == 0) red[0] = t; }
  __syncthreads();
  const float out = red[0];
  __syncthreads();
  return out;
}

// ascending bitonic sort of keys[0..npad), npad a power of two
__device__ void bitonic_sort(float* keys, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < npad / 2; t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
        const int p = i | j;
        const bool up = (i & k) == 0;
        const float a = keys[i], b = keys[p];
        if ((a > b) == up) { keys[i] = b; keys[p] = a; }
      }
    }
  }
  __syncthreads();
}

// nanquantile with linear interpolation over the n smallest (= non-NaN) sorted keys; NaN if there is none
__device__ __forceinline__ float quantile_sorted(const float* keys, int n, float q) {
  if (n <= 0) return __int_as_float(0x7fc00000);
  const float pos = q * (float)(n - 1);
  const int lo = (int)floorf(pos);
  const int hi = min(lo + 1, n - 1);
  return keys[lo] + (pos - (float)lo) * (keys[hi] - keys[lo]);
}

// loads x_b = src[b * stride] (+ optional per-walker term), sorts the non-NaN values, returns the clip bounds (loss.py:30-34)
template <class F>
__device__ void clip_bounds(F value, int64_t B, int npad, float* keys, float* red, float& lo, float& hi) {
  float cnt = 0.f;
  for (int t = threadIdx.x; t < npad; t += blockDim.x) {
    float v = INFINITY;
    if (t < B) { const float x = value(t); if (!isnan(x)) { v = x; cnt += 1.f; } }
    keys[t] = v;
  }
  const int n = (int)block_sum(cnt, red);
  bitonic_sort(keys, npad);
  const float q1 = quantile_sorted(keys, n, 0.25f), q3 = quantile_sorted(keys, n, 0.75f);
  const float iqr = q3 - q1;
  lo = q1 - 100.f * iqr;
  hi = q3 + 100.f * iqr;
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return isnan(x) ? x : fminf(fmaxf(x, lo), hi); }

// tasks (blockIdx.x): 0 Re E_L | 1 Im E_L | 2 L_z^2 | 3 L_z | 4 L^2 | 5 plain means of kinetic (re, im) and potential
__global__ void __launch_bounds__(ST_THREADS)
energy_stats_kernel(const float* __restrict__ el, const float* __restrict__ kin, const float* __restrict__ pot,
                    const float* __restrict__ lz, const float* __restrict__ lz2, const float* __restrict__ l2, int64_t B, int npad,
                    float* __restrict__ out) {
  extern __shared__ float st_keys[];
  __shared__ float red[32];
  const int task = blockIdx.x;
  const float invB = 1.f / (float)B;
  if (task == 5) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int64_t t = threadIdx.x; t < B; t += blockDim.x) { a += kin[2 * t]; b += kin[2 * t + 1]; c += pot[t]; }
    a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red);
    if (threadIdx.x == 0) { out[0] = a * invB; out[1] = b * invB; out[2] = c * invB; }
    return;
  }
  float lo, hi;
  if (task <= 1) {
    const int part = task;
    clip_bounds([&](int t) { return el[2 * t + part]; }, B, npad, st_keys, red, lo, hi);
    // complex nanmean: an element drops out if either part is NaN (jnp.nanmean on complex)
    float s = 0.f, sc = 0.f, n = 0.f, sq = 0.f, nr = 0.f;
    for (int64_t t = threadIdx.x; t < B; t += blockDim.x) {
      const float re = el[2 * t], im = el[2 * t + 1], x = part ? im : re;
      if (!isnan(re) && !isnan(im)) { s += x; sc += clampf(x, lo, hi); n += 1.f; }
      if (part == 0 && !isnan(re)) { sq += re * re; nr += 1.f; }
    }
    s = block_sum(s, red); sc = block_sum(sc, red); n = block_sum(n, red);
    if (part == 0) { sq = block_sum(sq, red); nr = block_sum(nr, red); }
    if (threadIdx.x == 0) {
      out[6 + part] = s / n;
      out[8 + part] = sc / n;
      if (part == 0) out[10] = sq / nr;
    }
    return;
  }
  const float* src = task == 2 ? lz2 : (task == 3 ? lz : l2);
  clip_bounds([&](int t) { return src[t]; }, B, npad, st_keys, red, lo, hi);
  float s = 0.f, sc = 0.f, n = 0.f;
  for (int64_t t = threadIdx.x; t < B; t += blockDim.x) {
    const float x = src[t];
    s += x;  // plain mean: a NaN walker contaminates it, as jnp.mean does (loss.py:68-71)
    if (!isnan(x)) { sc += clampf(x, lo, hi); n += 1.f; }
  }
  s = block_sum(s, red); sc = block_sum(sc, red); n = block_sum(n, red);
  if (threadIdx.x == 0) {
    out[task == 2 ? 4 : (task == 3 ? 3 : 5)] = s * invB;
    out[task == 2 ? 11 : (task == 3 ? 12 : 13)] = sc / n;
  }
}

// block p: 0 = real part, 1 = imaginary part of diff_to_clip = E_L - clipped (+ real penalty terms)
__global__ void __launch_bounds__(ST_THREADS)
energy_diff_kernel(const float* __restrict__ el, const float* __restrict__ lz, const float* __restrict__ lz2,
                   const float* __restrict__ l2, const float* __restrict__ lp, int64_t B, int npad, const float* __restrict__ red_stats,
                   float lz_penalty, float lz_center, float l2_penalty, float* __restrict__ diff, float* __restrict__ cot,
                   float* __restrict__ ok_out, float* __restrict__ counts) {
  extern __shared__ float st_keys[];
  __shared__ float red[32];
  const int part = blockIdx.x;
  const float c_re = red_stats[8], c_im = red_stats[9], c_lz2 = red_stats[11], c_lz = red_stats[12], c_l2 = red_stats[13];
  auto d_re = [&](int64_t t) {
    float d = el[2 * t] - c_re;
    if (lz_penalty != 0.f) d += lz_penalty * ((lz2[t] - c_lz2) - 2.f * lz_center * (lz[t] - c_lz));
    if (l2_penalty != 0.f) d += l2_penalty * (l2[t] - c_l2);
    return d;
  };
  auto d_im = [&](int64_t t) { return el[2 * t + 1] - c_im; };
  float lo, hi;
  if (part == 0) clip_bounds([&](int t) { return d_re(t); }, B, npad, st_keys, red, lo, hi);
  else clip_bounds([&](int t) { return d_im(t); }, B, npad, st_keys, red, lo, hi);
  // valid: both parts of diff are numbers (loss.py:60-64 nanmean over walkers); ok: additionally a finite log psi
  float nv = 0.f, nok = 0.f;
  for (int64_t t = threadIdx.x; t < B; t += blockDim.x) {
    const bool valid = !isnan(d_re(t)) && !isnan(d_im(t));
    const bool okb = valid && (lp == nullptr || (isfinite(lp[2 * t]) && isfinite(lp[2 * t + 1])));
    nv += valid ? 1.f : 0.f;
    nok += okb ? 1.f : 0.f;
  }
  nv = block_sum(nv, red);
  nok = block_sum(nok, red);
  const float scale = 2.f / fmaxf(nok, 1.f);
  for (int64_t t = threadIdx.x; t < B; t += blockDim.x) {
    const float dr = d_re(t), di = d_im(t);
    const bool valid = !isnan(dr) && !isnan(di);
    const bool okb = valid && (lp == nullptr || (isfinite(lp[2 * t]) && isfinite(lp[2 * t + 1])));
    const float d = clampf(part ? di : dr, lo, hi);
    diff[2 * t + part] = d;
    cot[2 * t + part] = okb ? d * scale : 0.f;
    if (part == 0 && ok_out) ok_out[t] = okb ? 1.f : 0.f;
  }
  if (part == 0 && threadIdx.x == 0 && counts) { counts[0] = nv; counts[1] = nok; }
}

int pad_pow2(int64_t B) { int n = 2; while (n < B) n <<= 1; return n; }

}  // namespace

int energy_stats(const float* el, const float* kin, const float* pot, const float* lz, const float* lz2, const float* l2, int64_t B,
                 float* out16, cudaStream_t s) {
  if (B < 1 || B > 32768) return -2;
  const int npad = pad_pow2(B);
  const size_t smem = (size_t)npad * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(energy_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = smem;
  }
  cudaError_t e = cudaMemsetAsync(out16, 0, 16 * sizeof(float), s);
  if (e != cudaSuccess) return (int)e;
  energy_stats_kernel<<<6, ST_THREADS, smem, s>>>(el, kin, pot, lz, lz2, l2, B, npad, out16);
  return (int)cudaGetLastError();
}

int energy_diff(const float* el, const float* lz, const float* lz2, const float* l2, const float* lp, int64_t B, const float* red_stats,
                float lz_penalty, float lz_center, float l2_penalty, float* diff, float* cot, float* ok, float* counts, cudaStream_t s) {
  if (B < 1 || B > 32768) return -2;
  const int npad = pad_pow2(B);
  const size_t smem = (size_t)npad * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(energy_diff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = smem;
  }
  energy_diff_kernel<<<2, ST_THREADS, smem, s>>>(el, lz, lz2, l2, lp, B, npad, red_stats, lz_penalty, lz_center, l2_penalty, diff, cot,
                                                 ok, counts);
  return (int)cudaGetLastError();
}

}  // namespace dh
