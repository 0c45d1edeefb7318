// fp32 SIMT GEMM with arbitrary strides (NN / NT / TN), split-K and jet-aware bias.
// Used for the VJP contractions and odd shapes; the forward contractions of the hot path
// run on tcgen05 (gemm_tc.cu).
#include "kernels.h"

namespace dh {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

// element (m,k) of A at A[m*a_sm + k*a_sk]; element (k,n) of B at B[k*b_sk + n*b_sn]
template <bool A_KCONTIG, bool B_NCONTIG, bool VEC>
__global__ void __launch_bounds__(256, 2)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ bias,
                 float* __restrict__ C, int64_t M, int N, int64_t K, int64_t a_sm, int64_t a_sk, int64_t b_sk,
                 int64_t b_sn, int64_t ldc, int rpg, int accumulate, int64_t k_per_split) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];

  auto load_tiles = [&](int64_t k0) {
    // ---- A tile: BM x BK
    if (A_KCONTIG) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int idx = tid + t * 256;
        int row = idx >> 2, kq = (idx & 3) * 4;
        int64_t m = m0 + row, k = k0 + kq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M) {
          const float* p = A + m * a_sm + k;
          if (VEC && k + 3 < kend) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (k + 0 < kend) v.x = p[0];
            if (k + 1 < kend) v.y = p[1];
            if (k + 2 < kend) v.z = p[2];
            if (k + 3 < kend) v.w = p[3];
          }
        }
        ra[t * 4 + 0] = v.x; ra[t * 4 + 1] = v.y; ra[t * 4 + 2] = v.z; ra[t * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int idx = tid + t * 256;
        int kk = idx >> 5, mq = (idx & 31) * 4;
        int64_t m = m0 + mq, k = k0 + kk;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend) {
          const float* p = A + k * a_sk + m;
          if (VEC && m + 3 < M) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (m + 0 < M) v.x = p[0];
            if (m + 1 < M) v.y = p[1];
            if (m + 2 < M) v.z = p[2];
            if (m + 3 < M) v.w = p[3];
          }
        }
        ra[t * 4 + 0] = v.x; ra[t * 4 + 1] = v.y; ra[t * 4 + 2] = v.z; ra[t * 4 + 3] = v.w;
      }
    }
    // ---- B tile: BK x BN
    if (B_NCONTIG) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int idx = tid + t * 256;
        int kk = idx >> 5, nq = (idx & 31) * 4;
        int64_t k = k0 + kk;
        int n = n0 + nq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend) {
          const float* p = Bm + k * b_sk + n;
          if (VEC && n + 3 < N) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (n + 0 < N) v.x = p[0];
            if (n + 1 < N) v.y = p[1];
            if (n + 2 < N) v.z = p[2];
            if (n + 3 < N) v.w = p[3];
          }
        }
        rb[t * 4 + 0] = v.x; rb[t * 4 + 1] = v.y; rb[t * 4 + 2] = v.z; rb[t * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int idx = tid + t * 256;
        int col = idx >> 2, kq = (idx & 3) * 4;
        int n = n0 + col;
        int64_t k = k0 + kq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N) {
          const float* p = Bm + (int64_t)n * b_sn + k;
          if (VEC && k + 3 < kend) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (k + 0 < kend) v.x = p[0];
            if (k + 1 < kend) v.y = p[1];
            if (k + 2 < kend) v.z = p[2];
            if (k + 3 < kend) v.w = p[3];
          }
        }
        rb[t * 4 + 0] = v.x; rb[t * 4 + 1] = v.y; rb[t * 4 + 2] = v.z; rb[t * 4 + 3] = v.w;
      }
    }
  };

  auto store_tiles = [&](int buf) {
    if (A_KCONTIG) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int idx = tid + t * 256;
        int row = idx >> 2, kq = (idx & 3) * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) As[buf][kq + c][row] = ra[t * 4 + c];
      }
    } else {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int idx = tid + t * 256;
        int kk = idx >> 5, mq = (idx & 31) * 4;
        *reinterpret_cast<float4*>(&As[buf][kk][mq]) = make_float4(ra[t * 4], ra[t * 4 + 1], ra[t * 4 + 2], ra[t * 4 + 3]);
      }
    }
    if (B_NCONTIG) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int idx = tid + t * 256;
        int kk = idx >> 5, nq = (idx & 31) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][kk][nq]) = make_float4(rb[t * 4], rb[t * 4 + 1], rb[t * 4 + 2], rb[t * 4 + 3]);
      }
    } else {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int idx = tid + t * 256;
        int col = idx >> 2, kq = (idx & 3) * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) Bs[buf][kq + c][col] = rb[t * 4 + c];
      }
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool has_next = (k0 + BK) < kend;
    if (has_next) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const bool add_bias = bias != nullptr && (rpg <= 1 || (m % rpg) == 0) && blockIdx.z == 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      if (add_bias) v += bias[n];
      float* c = C + m * ldc + n;
      if (split) {
        atomicAdd(c, v);
      } else {
        *c = accumulate ? (*c + v) : v;
      }
    }
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// C[M,N] (ldc) = A(M,K) . B(K,N); strides as documented above. split_k > 1 requires C
// pre-initialised (zero or the accumulate target) because partial sums are atomically added.
int gemm_simt(const float* A, const float* B, const float* bias, float* C, int64_t M, int N, int64_t K,
              int64_t a_sm, int64_t a_sk, int64_t b_sk, int64_t b_sn, int64_t ldc, int rpg, int accumulate,
              int split_k, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if (!((a_sm == 1) || (a_sk == 1)) || !((b_sk == 1) || (b_sn == 1))) return -1;
  const bool a_kc = (a_sk == 1);
  const bool b_nc = (b_sn == 1);
  if (split_k < 1) split_k = 1;
  int64_t kps = (K + split_k - 1) / split_k;
  kps = (kps + BK - 1) / BK * BK;
  split_k = (int)((K + kps - 1) / kps);
  if (split_k < 1) split_k = 1;
  const int64_t a_ld = a_kc ? a_sm : a_sk;
  const int64_t b_ld = b_nc ? b_sk : b_sn;
  const bool vec = aligned16(A) && aligned16(B) && (a_ld % 4 == 0) && (b_ld % 4 == 0);
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)split_k);
  if (grid.y > 65535 || grid.z > 65535) return -1;
#define DH_GEMM_LAUNCH(AK, BN_, V)                                                                          \
  gemm_simt_kernel<AK, BN_, V><<<grid, 256, 0, stream>>>(A, B, bias, C, M, N, K, a_sm, a_sk, b_sk, b_sn, \
                                                         ldc, rpg, accumulate, kps)
  if (a_kc && b_nc) { if (vec) DH_GEMM_LAUNCH(true, true, true); else DH_GEMM_LAUNCH(true, true, false); }
  else if (a_kc && !b_nc) { if (vec) DH_GEMM_LAUNCH(true, false, true); else DH_GEMM_LAUNCH(true, false, false); }
  else if (!a_kc && b_nc) { if (vec) DH_GEMM_LAUNCH(false, true, true); else DH_GEMM_LAUNCH(false, true, false); }
  else { if (vec) DH_GEMM_LAUNCH(false, false, true); else DH_GEMM_LAUNCH(false, false, false); }
#undef DH_GEMM_LAUNCH
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

}  // namespace dh
