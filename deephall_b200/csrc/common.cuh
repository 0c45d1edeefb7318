// Shared device helpers for the deephall_b200 kernels (sm_100a).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define DH_CHECK(expr)                         \
  do {                                         \
    cudaError_t _e = (expr);                   \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

#define DH_LAUNCH_CHECK()                      \
  do {                                         \
    cudaError_t _e = cudaGetLastError();       \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

namespace dh {

// Developer A/B switches read from the environment exist only in builds made with -DDH_DEBUG_SWITCHES (make DEBUG=1);
// the shipped library has ONE code path per operation and reads no environment variable.
#ifdef DH_DEBUG_SWITCHES
inline const char* dbg_env(const char* name) { return getenv(name); }
#else
inline const char* dbg_env(const char*) { return nullptr; }
#endif

// Row layout of a jet group (see oracle/jets.py): R = 2N + 8 rows per electron.
//   0: value | 1..2N: J_k | 2N+1: S | 2N+2..4: D_a | 2N+5..7: T_a
struct Rows {
  int N, R;
  __host__ __device__ explicit Rows(int n, bool jets) : N(n), R(jets ? 2 * n + 8 : 1) {}
  __host__ __device__ int J(int k) const { return 1 + k; }
  __host__ __device__ int S() const { return 2 * N + 1; }
  __host__ __device__ int D(int a) const { return 2 * N + 2 + a; }
  __host__ __device__ int T(int a) const { return 2 * N + 5 + a; }
};

typedef float2 cplx;
__host__ __device__ inline cplx cmake(float re, float im) { return make_float2(re, im); }
__host__ __device__ inline cplx cadd(cplx a, cplx b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ inline cplx csub(cplx a, cplx b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ inline cplx cmul(cplx a, cplx b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ inline cplx cscale(cplx a, float s) { return make_float2(a.x * s, a.y * s); }
__host__ __device__ inline cplx cconj(cplx a) { return make_float2(a.x, -a.y); }
__device__ inline cplx cfma(cplx a, cplx b, cplx c) {  // a*b + c
  return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, c.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, c.y)));
}
__device__ inline float cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
__device__ inline cplx cinv(cplx a) {
  float d = 1.0f / (a.x * a.x + a.y * a.y);
  return make_float2(a.x * d, -a.y * d);
}

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- two-piece fp16 split of an fp32 number (operands of the tcgen05 contraction, gemm_tc.cu)
// fp16 piece of x with saturation to the largest finite fp16 (NaN stays NaN)
__device__ __forceinline__ float sat_f16_range(float x) { return fabsf(x) > 65504.f ? copysignf(65504.f, x) : x; }
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(sat_f16_range(x));
  lo = __float2half_rn(sat_f16_range(x - __half2float(hi)));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {  // a in the low half
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// two floats -> packed fp16 pair (x0 in the low half), round-to-nearest, saturating to +-65504, NaN kept
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float x0, float x1) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
  return r;
}
// (x0, x1) -> hi pair, lo pair:  hi = fp16(x), lo = fp16(x - hi)
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = cvt_f16x2_sat(x0, x1);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = cvt_f16x2_sat(x0 - hf.x, x1 - hf.y);
}

// Activation tensors that feed the tcgen05 contraction can be kept as its operands: a [rows][D] fp32 buffer viewed
// as two fp16 planes [rows][D], hi at the start of the buffer and lo `plane` = rows * D halves later (same bytes).
// Four consecutive columns starting at element index idx:
__device__ __forceinline__ void st_planes4(__half* hi, int64_t plane, int64_t idx, const float4& v) {
  uint2 h, l;
  split_f16x2(v.x, v.y, h.x, l.x);
  split_f16x2(v.z, v.w, h.y, l.y);
  *reinterpret_cast<uint2*>(hi + idx) = h;
  *reinterpret_cast<uint2*>(hi + plane + idx) = l;
}
__device__ __forceinline__ float4 ld_planes4(const __half* hi, int64_t plane, int64_t idx) {
  const uint2 h = *reinterpret_cast<const uint2*>(hi + idx);
  const uint2 l = *reinterpret_cast<const uint2*>(hi + plane + idx);
  const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
  const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
  return make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
}

// Philox4x32-10 (Salmon et al. 2011), counter = (offset_lo, offset_hi, subseq_lo, subseq_hi).
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ explicit Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  __host__ __device__ inline uint4 operator()(uint64_t offset, uint64_t subseq) const {
    uint32_t c0 = (uint32_t)offset, c1 = (uint32_t)(offset >> 32);
    uint32_t c2 = (uint32_t)subseq, c3 = (uint32_t)(subseq >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// 24-bit mantissa uniform in [0, 1)
__host__ __device__ inline float u01(uint32_t r) { return (r >> 8) * (1.0f / 16777216.0f); }
// uniform in (0, 1] for Box-Muller
__host__ __device__ inline float u01_open0(uint32_t r) { return ((r >> 8) + 1) * (1.0f / 16777216.0f); }

constexpr float kTwoPi = 6.283185307179586f;

// The proposal draws of electron i of (global) walker `subseq` in move `offset`: a standard normal and a uniform in [0, 1)
// (Philox addressing: mcmc_kernels.cu).
__device__ inline void propose_draws(uint64_t seed, uint64_t offset, uint64_t subseq, int i, float& nrm, float& uph) {
  Philox ph(seed);
  const uint4 r = ph(offset * 64 + (uint64_t)i, subseq);
  const float u1 = u01_open0(r.x), u2 = u01(r.y);
  float s_, c_;
  sincosf(kTwoPi * u2, &s_, &c_);
  nrm = sqrtf(-2.f * logf(u1)) * c_;
  uph = u01(r.z);
}
// sph_sampling (mcmc.py:67-102): the point at polar angle atan(nrm width) / azimuth 2 pi uph around the north pole, carried
// to the frame of (theta, phi) by R_z(phi) R_y(theta).
__device__ inline void propose_point(float theta, float phi, float nrm, float uph, float width, float& th2, float& ph2) {
  const float theta_p = atanf(nrm * width);
  const float phi_p = uph * kTwoPi;
  float stp, ctp, spp, cpp, st, ct, sp, cp;
  sincosf(theta_p, &stp, &ctp);
  sincosf(phi_p, &spp, &cpp);
  sincosf(theta, &st, &ct);
  sincosf(phi, &sp, &cp);
  const float xp = stp * cpp, yp = stp * spp, zp = ctp;
  const float X = ct * xp + st * zp;
  const float Y = yp;
  const float Z = -st * xp + ct * zp;
  const float x2v = cp * X - sp * Y;
  const float y2v = sp * X + cp * Y;
  const float z2v = Z;
  th2 = acosf(fminf(fmaxf(z2v, -1.f), 1.f));
  const float sgn = (y2v > 0.f) ? 1.f : ((y2v < 0.f) ? -1.f : 0.f);
  ph2 = sgn * acosf(fminf(fmaxf(x2v / sinf(th2), -1.f), 1.f));
}
// log of the accept draw of (global) walker `subseq` in move `offset` (one rounding: oracle.mcmc.log_uniform)
__device__ inline float accept_log_uniform(uint64_t seed, uint64_t offset, uint64_t subseq) {
  Philox ph(seed);
  const float u = u01(ph(offset * 64 + 63, subseq).x);
  return (float)log((double)u);
}

__device__ inline double dpow_int(double z, int e) {
  double r = 1.0;
  while (e) {
    if (e & 1) r *= z;
    z *= z;
    e >>= 1;
  }
  return r;
}

}  // namespace dh
