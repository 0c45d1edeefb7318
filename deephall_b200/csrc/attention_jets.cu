// Self-attention on forward-Laplacian jets (flax MultiHeadAttention semantics, networks/psiformer.py:44).
// One block of 256 threads per (head, walker).  Rows per electron: R = 2N + 8 (common.cuh).
//
// Phase 1  score jets  s_ij^(r) = scale * sum_d [ q_i^(r) k_j^(0) + q_i^(0) k_j^(r) (+ 2 q_i^(r) k_j^(r) cross terms) ]
//          register-tiled: one thread owns (row r, 3 queries, 6 keys) = 54 accumulators; q and k stream
//          through shared memory 8 head-dim columns at a time, double-buffered with cp.async so that the
//          global-load latency of sub-chunk s+1 hides behind the FMAs of sub-chunk s (rows of 32 B, the
//          16-byte half h of row n stored at half h ^ ((n >> 2) & 1): conflict-free float4 reads).
// Phase 2  softmax jets (log-sum-exp Hessian = diag(p) - p p^T) in shared memory.
// Phase 3  output jets o_i^(r) = sum_j [ p^(r) v^(0) + p^(0) v^(r) (+ cross terms) ]
//          one thread owns (row r, 4 head-dim columns, 6 queries); the S-row cross term
//          2 sum_{j,k} p^(Jk) v^(Jk) is computed by warps whose lanes run over k and warp-reduced.
// FIRST-LAYER form (L0): q, k, v of the first layer depend on their own electron only, so their jets have
// just 10 non-zero rows per electron: value | own tangent flows (2) | S | D_a (3) | T_a (3).  The kernel then
// takes that compressed [B*N*10][3D] tensor (written by the feature kernel), does the score products on the
// 10 rows, scatters the own-flow products into the full row set (row J(2i+t) gets q_i^(t).k_j, row J(2j+t)
// gets q_i.k_j^(t)), runs the same softmax jets, and in P.V only touches the v rows that exist.
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace dh {

constexpr int AJ_THREADS = 256;
constexpr int AJ_RC = 10;      // compressed first-layer rows per electron
constexpr int AJ_CH = 16;      // head-dim columns per phase-3 step (two staged sub-chunks)
constexpr int AJ_SUB = 8;      // head-dim columns per staged sub-chunk (one 32-byte row)
constexpr int AJ_IB = 3, AJ_JB = 6, AJ_OB = 6;

__host__ __device__ inline int aj_np(int N) { return (N + AJ_OB - 1) / AJ_OB * AJ_OB; }  // padded query count
__host__ __device__ inline size_t aj_smem_floats(int N, int R, int RI) {
  const int NP = aj_np(N);
  // four staging regions [N*RI][SUB] ; sj, cr : [N][R][NP] ; p0, qq : [N][NP] ; dd : [3][N][NP] ; xs : [2][NP][CH]
  return 4 * (size_t)N * RI * AJ_SUB + 2 * (size_t)N * R * NP + ((5 * (size_t)N * NP + 3) & ~(size_t)3) + 2 * (size_t)NP * AJ_CH;
}
size_t attention_jets_smem(NetDims d) { return aj_smem_floats(d.N, d.R, d.R) * sizeof(float); }

// asynchronous copy of head-dim columns [sub*8, sub*8+8) of NR rows into a staging region
__device__ __forceinline__ void stage_async(float* dst, const float* __restrict__ src, int64_t ld, int NR, int sub, int hd) {
  // thread t takes the 16-byte half h = t & 1 of rows t >> 1, t >> 1 + 128, ...: with 256 threads h, the swizzle bit
  // ((row >> 2) & 1) and the column are the same for every row of a thread, so both addresses advance by constants
  static_assert(AJ_THREADS == 256, "the row stride of a thread must be a multiple of 8");
  const int h = threadIdx.x & 1, row0 = threadIdx.x >> 1;
  const int dcol = sub * AJ_SUB + h * 4;
  float* d = dst + row0 * AJ_SUB + ((h ^ ((row0 >> 2) & 1)) << 2);
  const float* sp = src + row0 * ld + dcol;
  if (dcol < hd) {  // hd % 4 == 0: a 16-byte piece is entirely inside or entirely outside the head
    for (int row = row0; row < NR; row += AJ_THREADS / 2, d += (AJ_THREADS / 2) * AJ_SUB, sp += (AJ_THREADS / 2) * ld) {
      const unsigned da = (unsigned)__cvta_generic_to_shared(d);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(sp) : "memory");
    }
  } else {
    for (int row = row0; row < NR; row += AJ_THREADS / 2, d += (AJ_THREADS / 2) * AJ_SUB)
      *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void stage_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// float4 number f (0 or 1) of staged row `row`
__device__ __forceinline__ float4 staged(const float* buf, int row, int f) {
  return *reinterpret_cast<const float4*>(buf + row * AJ_SUB + ((f ^ ((row >> 2) & 1)) << 2));
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}
__device__ __forceinline__ void axpy4(float4& acc, float p, const float4& v) {
  acc.x = fmaf(p, v.x, acc.x); acc.y = fmaf(p, v.y, acc.y);
  acc.z = fmaf(p, v.z, acc.z); acc.w = fmaf(p, v.w, acc.w);
}

// Sum of 24 per-lane values (6 queries x 4 columns) over the 32 lanes of a warp by recursive halving: each of the
// first three exchange rounds halves the number of values a lane keeps (12, 6, 3 shuffles), two more rounds finish the
// three that are left (27 shuffles in all, against 24 x 5 for one butterfly per value).  On return the lanes with
// (lane & 3) == 0 hold, in z[0..2], the complete sums of values base .. base + 2, base = 12 [lane & 16] + 6 [lane & 8]
// + 3 [lane & 4]; value index = query * 4 + column.
__device__ __forceinline__ int warp_sum24(const float4 (&acc)[6], float (&z)[3]) {
  const int lane = threadIdx.x & 31;
  const float v[24] = {acc[0].x, acc[0].y, acc[0].z, acc[0].w, acc[1].x, acc[1].y, acc[1].z, acc[1].w,
                       acc[2].x, acc[2].y, acc[2].z, acc[2].w, acc[3].x, acc[3].y, acc[3].z, acc[3].w,
                       acc[4].x, acc[4].y, acc[4].z, acc[4].w, acc[5].x, acc[5].y, acc[5].z, acc[5].w};
  const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
  float w[12], u[6];
#pragma unroll
  for (int t = 0; t < 12; ++t) {
    const float keep = h16 ? v[12 + t] : v[t], send = h16 ? v[t] : v[12 + t];
    w[t] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const float keep = h8 ? w[6 + t] : w[t], send = h8 ? w[t] : w[6 + t];
    u[t] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const float keep = h4 ? u[3 + t] : u[t], send = h4 ? u[t] : u[3 + t];
    z[t] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    z[t] += __shfl_xor_sync(0xffffffffu, z[t], 2);
    z[t] += __shfl_xor_sync(0xffffffffu, z[t], 1);
  }
  return (h16 ? 12 : 0) + (h8 ? 6 : 0) + (h4 ? 3 : 0);
}

// NT > 0: the electron count is a compile-time constant (index arithmetic folds, j-loops unroll);
// NT == 0: generic.
// Q12 (7 <= N <= 12): phase 3 gives every thread all 12 (padded) queries of one (row, 4 columns) -- twice the FMAs
// per shared-memory wavefront -- and the two halves of the block work on two 16-column steps at once.
template <int NT, bool L0, bool Q12>
__global__ void __launch_bounds__(AJ_THREADS, 2)
attention_jets_kernel(const float* __restrict__ qkv, float* __restrict__ o, NetDims dm, int o_flags) {
  extern __shared__ __align__(16) float smem[];
  const int o_pl = o_flags & 1;        // output as fp16 hi / lo planes
  const int row_rot = o_flags & 2;     // phase 3 (Q12): the second half of the block takes its rows rotated by 8
  const int N = NT > 0 ? NT : dm.N, R = 2 * N + 8, D = dm.D, hd = dm.hd;
  const int RI = L0 ? AJ_RC : R;  // rows per electron of the input tensor
  constexpr int JB = L0 ? 2 : AJ_JB;  // keys per thread in phase 1 (fewer rows -> smaller tiles keep all threads busy)
  const int NP = aj_np(N);
  const int hh = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NR = N * R;    // output rows of one walker
  const int NRI = N * RI;  // input rows of one walker
  Rows rw(N, true);
  // four staging regions: phase 1 uses (q, k) x 2 buffers, phase 3 two 16-column steps of v
  auto stg = [&](int u) { return smem + (size_t)u * NRI * AJ_SUB; };
  float* sj = smem + (size_t)4 * NRI * AJ_SUB;  // SJ(i,j,r)
  float* cr = sj + (size_t)N * R * NP;
  float* p0 = cr + (size_t)N * R * NP;      // [j][NP] (query index fastest)
  float* qq = p0 + N * NP;
  float* dd = qq + N * NP;                  // [3][j][NP]
  float* xs = p0 + ((5 * N * NP + 3) & ~3);  // [NP][CH] S-row cross term of the current chunk (16-byte aligned)
#define SJ(i, j, r) (((j) * R + (r)) * NP + (i))
  const int64_t ld = 3 * (int64_t)D;
  const float* qbase = qkv + b * NRI * ld + hh * hd;
  const float* kbase = qbase + D;
  const float* vbase = qbase + 2 * D;
  const float scl = rsqrtf((float)hd);
  const int nchunk = (hd + AJ_CH - 1) / AJ_CH;
  const int nsub = (hd + AJ_SUB - 1) / AJ_SUB;
  const int IBN = (N + AJ_IB - 1) / AJ_IB, JBN = (N + JB - 1) / JB, OBN = NP / AJ_OB;

  // zero the padded query slots so that vectorised reads of sj never see garbage
  for (int t = tid; t < 2 * N * R * NP; t += AJ_THREADS) sj[t] = 0.f;  // sj and cr are contiguous

  // ------------------------------------------------------------------ phase 1: score jets
  const int items1 = RI * IBN * JBN;
  for (int it0 = 0; it0 < items1; it0 += AJ_THREADS) {
    const int item = it0 + tid;
    const bool active = item < items1;
    const int r = active ? item % RI : 0;  // input row (L0: compressed row index)
    const int ib = active ? (item / RI) % IBN : 0;
    const int jb = active ? item / (RI * IBN) : 0;
    const int i0 = ib * AJ_IB, j0 = jb * JB;
    float acc[AJ_IB][JB][3];
#pragma unroll
    for (int a = 0; a < AJ_IB; ++a)
#pragma unroll
      for (int c = 0; c < JB; ++c) { acc[a][c][0] = 0.f; acc[a][c][1] = 0.f; acc[a][c][2] = 0.f; }
    __syncthreads();  // staging regions are free (previous pass / zero-fill above)
    stage_async(stg(0), qbase, ld, NRI, 0, hd);
    stage_async(stg(1), kbase, ld, NRI, 0, hd);  // (one commit group each: q and k of a sub-chunk = 2 groups)
    for (int sb = 0; sb < nsub; ++sb) {
      const float* qs = stg(2 * (sb & 1));
      const float* ks = stg(2 * (sb & 1) + 1);
      if (sb + 1 < nsub) {
        stage_async(stg(2 * ((sb + 1) & 1)), qbase, ld, NRI, sb + 1, hd);
        stage_async(stg(2 * ((sb + 1) & 1) + 1), kbase, ld, NRI, sb + 1, hd);
        asm volatile("cp.async.wait_group 2;" ::: "memory");  // everything but the two groups just issued
      } else {
        stage_wait_all();
      }
      __syncthreads();
      if (active) {
#pragma unroll
        for (int f = 0; f < AJ_SUB / 4; ++f) {
          float4 qr[AJ_IB], q0[AJ_IB];
#pragma unroll
          for (int a = 0; a < AJ_IB; ++a) {
            const int i = min(i0 + a, N - 1);
            qr[a] = staged(qs, i * RI + r, f);
            q0[a] = staged(qs, i * RI, f);
          }
#pragma unroll
          for (int c = 0; c < JB; ++c) {
            const int j = min(j0 + c, N - 1);
            const float4 kr = staged(ks, j * RI + r, f);
            const float4 k0 = staged(ks, j * RI, f);
#pragma unroll
            for (int a = 0; a < AJ_IB; ++a) {
              acc[a][c][0] = dot4(qr[a], k0, acc[a][c][0]);
              acc[a][c][1] = dot4(q0[a], kr, acc[a][c][1]);
              acc[a][c][2] = dot4(qr[a], kr, acc[a][c][2]);
            }
          }
        }
      }
      __syncthreads();  // all reads of this buffer are done before it is refilled two iterations later
    }
    if (active) {
#pragma unroll
      for (int a = 0; a < AJ_IB; ++a)
#pragma unroll
        for (int c = 0; c < JB; ++c) {
          const int i = i0 + a, j = j0 + c;
          if (i < N && j < N) {
            if (!L0) {
              sj[SJ(i, j, r)] = (r == 0) ? acc[a][c][2] * scl : (acc[a][c][0] + acc[a][c][1]) * scl;
              cr[SJ(i, j, r)] = acc[a][c][2] * scl;
            } else if (r == 0) {
              sj[SJ(i, j, 0)] = acc[a][c][2] * scl;
            } else if (r <= 2) {
              // own tangent flow t of the query electron moves q_i only, of the key electron k_j only; no other
              // thread writes these two entries (they coincide when i == j)
              const int t = r - 1;
              sj[SJ(i, j, rw.J(2 * i + t))] += acc[a][c][0] * scl;
              sj[SJ(i, j, rw.J(2 * j + t))] += acc[a][c][1] * scl;
              if (i == j) cr[SJ(i, j, rw.J(2 * i + t))] = acc[a][c][2] * scl;
            } else {
              const int rf = r == 3 ? rw.S() : (r <= 6 ? rw.D(r - 4) : rw.T(r - 7));
              sj[SJ(i, j, rf)] = (acc[a][c][0] + acc[a][c][1]) * scl;
              cr[SJ(i, j, rf)] = acc[a][c][2] * scl;
            }
          }
        }
    }
  }
  __syncthreads();
  // second-order rows pick up the cross products: S += 2 sum_k qJk.kJk ; T_a += 2 qDa.kDa
  for (int t = tid; t < N * N * 4; t += AJ_THREADS) {
    const int w = t & 3, ij = t >> 2;
    const int i = ij % N, j = ij / N;
    if (w == 0) {
      float s2 = 0.f;
      for (int k = 0; k < 2 * N; ++k) s2 += cr[SJ(i, j, rw.J(k))];
      sj[SJ(i, j, rw.S())] += 2.f * s2;
    } else {
      sj[SJ(i, j, rw.T(w - 1))] += 2.f * cr[SJ(i, j, rw.D(w - 1))];
    }
  }
  __syncthreads();
  // ------------------------------------------------------------------ phase 2: softmax jets
  for (int i = tid; i < N; i += AJ_THREADS) {
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, sj[SJ(i, j, 0)]);
    float Z = 0.f;
    for (int j = 0; j < N; ++j) { const float e = expf(sj[SJ(i, j, 0)] - mx); p0[j * NP + i] = e; Z += e; }
    const float iz = 1.f / Z;
    for (int j = 0; j < N; ++j) p0[j * NP + i] *= iz;
  }
  __syncthreads();
  const int nfirst = 2 * N + 3;
  for (int t = tid; t < N * nfirst; t += AJ_THREADS) {  // first-order rows: l = s - lse
    const int i = t % N, q = t / N;
    const int r = q < 2 * N ? rw.J(q) : rw.D(q - 2 * N);
    float lse = 0.f;
    for (int j = 0; j < N; ++j) lse = fmaf(p0[j * NP + i], sj[SJ(i, j, r)], lse);
    for (int j = 0; j < N; ++j) sj[SJ(i, j, r)] -= lse;
  }
  __syncthreads();
  for (int t = tid; t < N * N; t += AJ_THREADS) {
    const int i = t % N, j = t / N;
    float s2 = 0.f;
    for (int k = 0; k < 2 * N; ++k) { const float l = sj[SJ(i, j, rw.J(k))]; s2 = fmaf(l, l, s2); }
    qq[j * NP + i] = s2;
    for (int a3 = 0; a3 < 3; ++a3) { const float l = sj[SJ(i, j, rw.D(a3))]; dd[(a3 * N + j) * NP + i] = l * l; }
  }
  __syncthreads();
  for (int t = tid; t < N * 4; t += AJ_THREADS) {  // second-order rows
    const int i = t % N, w = t / N;
    const int r = w == 0 ? rw.S() : rw.T(w - 1);
    const float* extra = w == 0 ? qq : dd + (w - 1) * N * NP;
    float lse = 0.f;
    for (int j = 0; j < N; ++j) {
      const float v = sj[SJ(i, j, r)] + extra[j * NP + i];
      sj[SJ(i, j, r)] = v;
      lse = fmaf(p0[j * NP + i], v, lse);
    }
    for (int j = 0; j < N; ++j) sj[SJ(i, j, r)] -= lse;
  }
  __syncthreads();
  for (int t = tid; t < N * R * N; t += AJ_THREADS) {  // l -> p jets
    const int i = t % N, jr = t / N;
    const int r = jr % R, j = jr / R;
    const float pv = p0[j * NP + i];
    sj[SJ(i, j, r)] = r == 0 ? pv : pv * sj[SJ(i, j, r)];
  }
  // ------------------------------------------------------------------ phase 3: o = P V jets
  float* obase = o + b * NR * (int64_t)D + hh * hd;
  // o_pl: the output goes out as fp16 hi / lo planes, the left operand of the contraction that follows
  __half* opl = reinterpret_cast<__half*>(o) + b * NR * (int64_t)D + hh * hd;
  const int64_t oplane = (int64_t)gridDim.y * NR * D;
  const int rS = rw.S(), rT0 = rw.T(0), rD0 = rw.D(0);
  if (Q12) {
    // two 16-column steps at a time: threads 0..127 take step cp (regions 0, 1), threads 128..255 step cp + 1
    // (regions 2, 3); thread = (f4, r) owns all 12 query slots
    for (int cp = 0; cp < nchunk; cp += 2) {
      __syncthreads();  // staging regions and xs are free
      for (int u = 0; u < 4; ++u)
        if (cp + (u >> 1) < nchunk) stage_async(stg(u), vbase, ld, NRI, 2 * cp + u, hd);
      stage_wait_all();
      __syncthreads();
      // ---- S-row cross term of both steps: warp item = (step, query block, float4), lanes = k
      for (int wi = warp; wi < 2 * OBN * 4; wi += AJ_THREADS / 32) {
        const int c2 = wi / (OBN * 4), ob = (wi >> 2) % OBN, f4 = wi & 3;
        if (cp + c2 >= nchunk) continue;
        const float* vpair = stg(2 * c2);
        float4 acc[AJ_OB];
#pragma unroll
        for (int a = 0; a < AJ_OB; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kk = lane; kk < 2 * N; kk += 32) {
          const int r = rw.J(kk);
          for (int j = L0 ? (kk >> 1) : 0; j < (L0 ? (kk >> 1) + 1 : N); ++j) {
            const float4 v = staged(vpair + (size_t)(f4 >> 1) * NRI * AJ_SUB, L0 ? j * RI + 1 + (kk & 1) : j * R + r, f4 & 1);
            const float2* pp = reinterpret_cast<const float2*>(sj + SJ(ob * AJ_OB, j, r));
            const float2 pa = pp[0], pb = pp[1], pc = pp[2];
            axpy4(acc[0], pa.x, v); axpy4(acc[1], pa.y, v);
            axpy4(acc[2], pb.x, v); axpy4(acc[3], pb.y, v);
            axpy4(acc[4], pc.x, v); axpy4(acc[5], pc.y, v);
          }
        }
        float z[3];
        const int zb = warp_sum24(acc, z);
        if ((lane & 3) == 0) {
#pragma unroll
          for (int t = 0; t < 3; ++t)
            xs[(c2 * NP + ob * AJ_OB + ((zb + t) >> 2)) * AJ_CH + f4 * 4 + ((zb + t) & 3)] = 2.f * z[t];
        }
      }
      __syncthreads();
      // ---- all rows
      const int c2 = tid >> 7, idx = tid & 127;
      if (idx < 4 * R && cp + c2 < nchunk) {
        // The T rows carry a third product loop (36 key passes against 24), so the warp that holds them is the slow
        // one of its half.  Warps 3 and 7 share a scheduler (warp % 4): with the same row order in both halves that
        // scheduler would issue 72 units per step against 48 for the others.  The second half therefore takes its
        // rows rotated by one warp (8 rows): its slow warp is warp 6, and the per-scheduler maximum drops to 60.
        // Blocks alternate (by block number / 128, which separates most pairs of blocks that share an SM in the first
        // waves) between slow warps on schedulers {3, 2} and {1, 0}.
        const int f4 = idx & 3;
        int r = idx >> 2;
        if (row_rot) {
          const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
          r += 8 * ((c2 + 2 * ((bid >> 7) & 1)) & 3);
          while (r >= R) r -= R;
        }
        const float* vsub = stg(2 * c2) + (size_t)(f4 >> 1) * NRI * AJ_SUB;
        const int fh = f4 & 1;
        float4 acc[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool isT = r >= rT0;
        const int rd = isT ? rD0 + (r - rT0) : 0;
        const bool isJ = r >= 1 && r <= 2 * N;
        const int jown = isJ ? (r - 1) >> 1 : -1;
        const int rin = !L0 ? r : (isJ ? 1 + ((r - 1) & 1) : (r == rS ? 3 : (isT ? 7 + (r - rT0) : (r == 0 ? 0 : 4 + (r - rD0)))));
        const int rdin = L0 ? 4 + (r - rT0) : rd;
        auto axpy12 = [&](const float* prow, float w, const float4& v) {  // acc[i] += w * prow[i] * v, i < 12
          const float4* p4 = reinterpret_cast<const float4*>(prow);
          const float4 a0 = p4[0], a1 = p4[1], a2 = p4[2];
          axpy4(acc[0], w * a0.x, v); axpy4(acc[1], w * a0.y, v); axpy4(acc[2], w * a0.z, v); axpy4(acc[3], w * a0.w, v);
          axpy4(acc[4], w * a1.x, v); axpy4(acc[5], w * a1.y, v); axpy4(acc[6], w * a1.z, v); axpy4(acc[7], w * a1.w, v);
          axpy4(acc[8], w * a2.x, v); axpy4(acc[9], w * a2.y, v); axpy4(acc[10], w * a2.z, v); axpy4(acc[11], w * a2.w, v);
        };
        // (the row-dependent terms are separate loops: one branch per thread instead of one per key, and in the
        // first-layer form an own-flow row touches a single key)
        if (isT) {  // 2 sum_j p^(D_a) v^(D_a) first, doubled once in the accumulators (not once per product)
          for (int j = 0; j < N; ++j) axpy12(sj + SJ(0, j, rd), 1.f, staged(vsub, j * RI + rdin, fh));
#pragma unroll
          for (int a = 0; a < 12; ++a) { acc[a].x *= 2.f; acc[a].y *= 2.f; acc[a].z *= 2.f; acc[a].w *= 2.f; }
        }
        for (int j = 0; j < N; ++j) axpy12(sj + SJ(0, j, r), 1.f, staged(vsub, j * RI, fh));
        if (L0 && isJ) {
          axpy12(sj + SJ(0, jown, 0), 1.f, staged(vsub, jown * RI + rin, fh));
        } else if (r != 0) {
          for (int j = 0; j < N; ++j) axpy12(sj + SJ(0, j, 0), 1.f, staged(vsub, j * RI + rin, fh));
        }
        const int dcol = (cp + c2) * AJ_CH + f4 * 4;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          if (i < N) {
            float4 out = acc[i];
            if (r == rS) {
              const float4 x4 = *reinterpret_cast<const float4*>(xs + (c2 * NP + i) * AJ_CH + f4 * 4);
              out.x += x4.x; out.y += x4.y; out.z += x4.z; out.w += x4.w;
            }
            float* dst = obase + (int64_t)(i * R + r) * D + dcol;
            if (o_pl) { if (dcol < hd) st_planes4(opl, oplane, (int64_t)(i * R + r) * D + dcol, out); }
            else if (dcol + 3 < hd) *reinterpret_cast<float4*>(dst) = out;
            else {
              if (dcol < hd) dst[0] = out.x;
              if (dcol + 1 < hd) dst[1] = out.y;
              if (dcol + 2 < hd) dst[2] = out.z;
            }
          }
        }
      }
    }
    return;
  }
  // a 16-column step of v = two staged sub-chunks; steps are double-buffered over the four regions
  __syncthreads();
  stage_async(stg(0), vbase, ld, NRI, 0, hd);
  stage_async(stg(1), vbase, ld, NRI, 1, hd);
  for (int ch = 0; ch < nchunk; ++ch) {
    const float* vpair = stg(2 * (ch & 1));  // sub-chunk u of this step at vpair + u * NRI * AJ_SUB
    if (ch + 1 < nchunk) {
      stage_async(stg(2 * ((ch + 1) & 1)), vbase, ld, NRI, 2 * (ch + 1), hd);
      stage_async(stg(2 * ((ch + 1) & 1) + 1), vbase, ld, NRI, 2 * (ch + 1) + 1, hd);
      asm volatile("cp.async.wait_group 2;" ::: "memory");
    } else {
      stage_wait_all();
    }
    __syncthreads();
    // ---- S-row cross term: xs[i][c] = 2 sum_j sum_k p_ij^(Jk) v_j^(Jk)[c] ; warp = (query block, float4), lanes = k
    for (int wi = warp; wi < OBN * 4; wi += AJ_THREADS / 32) {
      const int ob = wi >> 2, f4 = wi & 3;
      float4 acc[AJ_OB];
#pragma unroll
      for (int a = 0; a < AJ_OB; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int kk = lane; kk < 2 * N; kk += 32) {
        const int r = rw.J(kk);
        // L0: v_j moves along flow kk only if j is that flow's electron
        for (int j = L0 ? (kk >> 1) : 0; j < (L0 ? (kk >> 1) + 1 : N); ++j) {
          const float4 v = staged(vpair + (size_t)(f4 >> 1) * NRI * AJ_SUB, L0 ? j * RI + 1 + (kk & 1) : j * R + r, f4 & 1);
          const float2* pp = reinterpret_cast<const float2*>(sj + SJ(ob * AJ_OB, j, r));
          const float2 pa = pp[0], pb = pp[1], pc = pp[2];
          axpy4(acc[0], pa.x, v); axpy4(acc[1], pa.y, v);
          axpy4(acc[2], pb.x, v); axpy4(acc[3], pb.y, v);
          axpy4(acc[4], pc.x, v); axpy4(acc[5], pc.y, v);
        }
      }
      float z[3];
      const int zb = warp_sum24(acc, z);
      if ((lane & 3) == 0) {
#pragma unroll
        for (int t = 0; t < 3; ++t)
          xs[(ob * AJ_OB + ((zb + t) >> 2)) * AJ_CH + f4 * 4 + ((zb + t) & 3)] = 2.f * z[t];
      }
    }
    __syncthreads();
    // ---- all rows: thread = (f4, r, query block)
    for (int t = tid; t < R * 4 * OBN; t += AJ_THREADS) {
      const int f4 = t & 3, r = (t >> 2) % R, ob = (t >> 2) / R;
      float4 acc[AJ_OB];
#pragma unroll
      for (int a = 0; a < AJ_OB; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool isT = r >= rT0;
      const int rd = isT ? rD0 + (r - rT0) : 0;
      // L0: input row of v that carries output row r (own-flow rows exist for one electron only)
      const bool isJ = r >= 1 && r <= 2 * N;
      const int jown = isJ ? (r - 1) >> 1 : -1;
      const int rin = !L0 ? r : (isJ ? 1 + ((r - 1) & 1) : (r == rS ? 3 : (isT ? 7 + (r - rT0) : (r == 0 ? 0 : 4 + (r - rD0)))));
      const int rdin = L0 ? 4 + (r - rT0) : rd;
      for (int j = 0; j < N; ++j) {
        const float4 v0 = staged(vpair + (size_t)(f4 >> 1) * NRI * AJ_SUB, j * RI, f4 & 1);
        const float2* pr = reinterpret_cast<const float2*>(sj + SJ(ob * AJ_OB, j, r));
        const float2 a0 = pr[0], a1 = pr[1], a2 = pr[2];
        axpy4(acc[0], a0.x, v0); axpy4(acc[1], a0.y, v0);
        axpy4(acc[2], a1.x, v0); axpy4(acc[3], a1.y, v0);
        axpy4(acc[4], a2.x, v0); axpy4(acc[5], a2.y, v0);
        if (r != 0 && (!L0 || !isJ || j == jown)) {
          const float4 vr = staged(vpair + (size_t)(f4 >> 1) * NRI * AJ_SUB, j * RI + rin, f4 & 1);
          const float2* pz = reinterpret_cast<const float2*>(sj + SJ(ob * AJ_OB, j, 0));
          const float2 b0 = pz[0], b1 = pz[1], b2 = pz[2];
          axpy4(acc[0], b0.x, vr); axpy4(acc[1], b0.y, vr);
          axpy4(acc[2], b1.x, vr); axpy4(acc[3], b1.y, vr);
          axpy4(acc[4], b2.x, vr); axpy4(acc[5], b2.y, vr);
          if (isT) {
            const float4 vd = staged(vpair + (size_t)(f4 >> 1) * NRI * AJ_SUB, j * RI + rdin, f4 & 1);
            const float2* pd = reinterpret_cast<const float2*>(sj + SJ(ob * AJ_OB, j, rd));
            const float2 c0 = pd[0], c1 = pd[1], c2 = pd[2];
            axpy4(acc[0], 2.f * c0.x, vd); axpy4(acc[1], 2.f * c0.y, vd);
            axpy4(acc[2], 2.f * c1.x, vd); axpy4(acc[3], 2.f * c1.y, vd);
            axpy4(acc[4], 2.f * c2.x, vd); axpy4(acc[5], 2.f * c2.y, vd);
          }
        }
      }
      const int dcol = ch * AJ_CH + f4 * 4;
#pragma unroll
      for (int a = 0; a < AJ_OB; ++a) {
        const int i = ob * AJ_OB + a;
        if (i < N) {
          float4 out = acc[a];
          if (r == rS) {
            const float4 x4 = *reinterpret_cast<const float4*>(xs + i * AJ_CH + f4 * 4);
            out.x += x4.x; out.y += x4.y; out.z += x4.z; out.w += x4.w;
          }
          float* dst = obase + (int64_t)(i * R + r) * D + dcol;
          if (o_pl) { if (dcol < hd) st_planes4(opl, oplane, (int64_t)(i * R + r) * D + dcol, out); }
          else if (dcol + 3 < hd) *reinterpret_cast<float4*>(dst) = out;
          else {
            if (dcol < hd) dst[0] = out.x;
            if (dcol + 1 < hd) dst[1] = out.y;
            if (dcol + 2 < hd) dst[2] = out.z;
          }
        }
      }
    }
    __syncthreads();  // v buffers of this step and xs are free again
  }
#undef SJ
}

int attention_jets(const float* qkv, float* o, int64_t B, NetDims d, int layer0, int fp32_only, cudaStream_t s) {
  // head size 64 and the electron counts of the BASELINE configurations: contractions on the tensor cores with fp16
  // pieces (attention_tc.cu).  fp32_only (plans with contraction = tf32 / fp32): this file's fp32-FMA form.
  if (!fp32_only && attention_jets_tc_ok(d)) return attention_jets_tc(qkv, o, B, d, layer0, s);
  const int o_pl = 0;
  if (o_pl && ((d.D % 8) != 0 || (reinterpret_cast<uintptr_t>(o) & 15))) return -2;
  if (d.N > 16 || d.R != 2 * d.N + 8 || (d.hd % 4) != 0 || (d.D % 4) != 0) return -2;
  const size_t smem = aj_smem_floats(d.N, d.R, layer0 ? AJ_RC : d.R) * sizeof(float);
  if (smem > 227 * 1024) return -2;
  dim3 grid((unsigned)d.H, (unsigned)B);
  static const int row_rot = (dbg_env("DH_ATT_ROT") && atoi(dbg_env("DH_ATT_ROT")) == 0) ? 0 : 2;
  const int o_flags = (o_pl ? 1 : 0) | row_rot;
#define DH_AJ_LAUNCH(NT, LZ)                                                                                          \
  do {                                                                                                                \
    static size_t attr_smem = 0;                                                                                      \
    if (smem > attr_smem) {                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(attention_jets_kernel<NT, LZ, (NT >= 7 && NT <= 12)>,                    \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
      if (e != cudaSuccess) return (int)e;                                                                            \
      attr_smem = smem;                                                                                               \
    }                                                                                                                 \
    attention_jets_kernel<NT, LZ, (NT >= 7 && NT <= 12)><<<grid, AJ_THREADS, smem, s>>>(qkv, o, d, o_flags);                \
  } while (0)
#define DH_AJ_BOTH(NT) do { if (layer0) DH_AJ_LAUNCH(NT, true); else DH_AJ_LAUNCH(NT, false); } while (0)
  switch (d.N) {
    case 6: DH_AJ_BOTH(6); break;
    case 10: DH_AJ_BOTH(10); break;
    case 12: DH_AJ_BOTH(12); break;
    case 16: DH_AJ_BOTH(16); break;
    default: DH_AJ_BOTH(0); break;
  }
#undef DH_AJ_BOTH
#undef DH_AJ_LAUNCH
  return (int)cudaGetLastError();
}

}  // namespace dh
