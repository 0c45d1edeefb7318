// Self-attention on forward-Laplacian jets with the contractions on the tensor cores (flax MultiHeadAttention
// semantics, networks/psiformer.py:44).  Same contract as attention_jets.cu (which stays as the generic form for
// head sizes other than 64): one block of 256 threads per (head, walker), rows per electron R = 2N + 8.
//
// Per (walker, head) the jet algebra is a handful of small-N matrix products (N <= 16 electrons, hd = 64):
//   G1  U[(i,r), j] = q_i^(r) . k_j^(0)          [N R x 64] x [64 x N]      first-order + linear second-order terms
//   G2  W[(j,r), i] = k_j^(r) . q_i^(0)          [N R x 64] x [64 x N]
//   X   X[i, j]     = sum_k q_i^(Jk) . k_j^(Jk)  [N x 64 (2N)] x [64 (2N) x N]   S-row cross term;  same per D_a -> T_a
//   PV  o_i^(r)     = P^(r) V^(0) + P^(0) V^(r) (+ 2 sum_k P^(Jk) V^(Jk) for S, + 2 P^(Da) V^(Da) for T_a)
// All of them run as `mma.sync.m16n8k16` (fp16 operands, fp32 accumulation) with fp32 accuracy from the same
// two-piece split as the dense contractions (gemm_tc.cu): x = hi + lo, products lo*hi + hi*lo + hi*hi.  q, k, v are
// split ONCE while they are staged: 16 head-dim columns at a time arrive by `cp.async` in a raw fp32 area WHILE the
// previous 16 columns are multiplied, then every thread rewrites its own pieces as fp16 hi / lo planes; fragments
// come from the planes by `ldmatrix` (transposed for V).  The softmax jets in between (log-sum-exp Hessian =
// diag(p) - p p^T) are the same shared-memory pass as in attention_jets.cu.
//
// Plane layout: row (electron e, jet row r) of a plane with RB bytes per row lives at
//   e * (R * RB + 16) + r * RB + 16 * (chunk ^ swizzle(r)):
// the 16-byte pad per electron makes the 8 rows {e R + r, e = e0 .. e0 + 7} of an ldmatrix (cross terms, P.V) fall
// into 8 different bank groups, the swizzle does the same for 8 consecutive rows of one electron (G1, G2).
// FIRST-LAYER form (L0): q, k, v arrive compressed to their 10 non-zero rows per electron (value | own flows (2) |
// S | D_a | T_a, written by the feature kernel); the staging step expands them into the full row set (zeros
// elsewhere) and everything after it is the same code.
#include "kernels.h"

namespace dh {

namespace {

constexpr int AT_THREADS = 256;
constexpr int AT_WARPS = AT_THREADS / 32;
constexpr int AT_NP = 16;   // padded electron count (one m16 / two n8 tiles)
constexpr int AT_HD = 64;   // head size this kernel is built for
constexpr int AT_RC = 10;   // compressed first-layer rows per electron

#ifdef DH_DEBUG_SWITCHES
// developer builds (make DEBUG=1): per-phase clock sums of thread 0 of every block (both kernels add into the same slots), read by dh_debug_at_prof
// [0] blocks [1] whole kernel [2] score steps: wait + barrier + conversion [3] score steps: multiply [4] score write-out
// [5] softmax [6] P fragments [7] P.V steps: wait + barrier + conversion [8] P.V steps: multiply + stores
__device__ unsigned long long g_at_prof[16];
#define AT_PROF_DECL long long tp_ = clock64(), tp0_ = tp_; (void)tp0_;
#define AT_LAP(slot) if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(g_at_prof + (slot), (unsigned long long)(t_ - tp_)); tp_ = t_; }
#define AT_PROF_END if (threadIdx.x == 0) { atomicAdd(g_at_prof + 1, (unsigned long long)(clock64() - tp0_)); atomicAdd(g_at_prof, 1ull); }
#else
#define AT_PROF_DECL
#define AT_LAP(slot)
#define AT_PROF_END
#endif

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// the three products of the split, small terms first
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
  mma16816(c, al, bh0, bh1);
  mma16816(c, ah, bl0, bl1);
  mma16816(c, ah, bh0, bh1);
}

template <int R, int RB>
__device__ __forceinline__ uint32_t plane_off(int row, int chunk) {
  const int e = row / R, r = row - e * R;
  const int sw = RB == 32 ? ((r >> 2) & 1) : ((r >> 1) & 3);
  return (uint32_t)(e * (R * RB + 16) + r * RB + ((chunk ^ sw) << 4));
}

template <int NT>
struct AtGeom {
  static constexpr int N = NT, R = 2 * NT + 8, NR = N * R;
  static constexpr int MT = (NR + 15) / 16;                 // m16 tiles over the flat (electron, row) index
  static constexpr int TPW = (MT + AT_WARPS - 1) / AT_WARPS;  // tiles per warp
  static constexpr int RPW = (R + AT_WARPS - 1) / AT_WARPS;   // jet rows per warp in P.V
  static constexpr int PL = N * (R * 32 + 16);              // bytes of a 16-column fp16 plane
  static constexpr int SCRATCH = 4 * (5 * N * AT_NP + 11 * 256);  // qq, dd, xw, dw (softmax scratch, in the idle planes)
  static constexpr int PLANES = 4 * PL > SCRATCH ? 4 * PL : SCRATCH;  // q hi | q lo | k hi | k lo  (P.V: v hi | v lo)
  static constexpr int RAW = NR * 64;                       // bytes of 16 fp32 columns of one tensor, as they arrive
  // sj[j][r][i]: jet-row stride 17 and floats per key electron = 4 (mod 16).  The score accumulators leave the fragments as
  // (row g, column pair t4) per lane: with these strides the 32 lanes of a store hit 32 different banks (strides 16 / 516
  // put them into 4 banks: 8-way conflicts on every store of the write-out)
  static constexpr int SJ_R = AT_NP + 1;
  static constexpr int SJ_J = R * SJ_R + ((4 - R * SJ_R) % 16 + 16) % 16;
  static constexpr int SJ_FLOATS = (N * SJ_J + 3) & ~3;
  // shared memory: planes | raw | sj | p0 [N][16] | red [warp][lane][8]
  //   scores: the raw area of q AND k (2 RAW bytes) runs over sj, which is not live before the last stage is converted
  //   softmax: qq, dd and xw, dw live in the (idle) planes
  static constexpr int OFF_RAW = PLANES, OFF_SJ = OFF_RAW + RAW, OFF_P0 = OFF_SJ + 4 * SJ_FLOATS;
  static constexpr int OFF_RED = OFF_P0 + 4 * N * AT_NP, SMEM_USED = OFF_RED + 4 * AT_WARPS * 256;
  static constexpr size_t SMEM = SMEM_USED > OFF_RAW + 2 * RAW ? SMEM_USED : OFF_RAW + 2 * RAW;
  // staging: thread = (jet row r, 16-byte piece q4) of EPP electrons at a time
  static constexpr int SLOTS = R * 4, EPP = AT_THREADS / SLOTS >= 1 ? AT_THREADS / SLOTS : 1;
  static_assert(SLOTS <= AT_THREADS, "one pass covers at least one electron");
};

// compressed first-layer row of full row r of electron e, or -1 when that row is identically zero
template <int NT>
__device__ __forceinline__ int l0_row(int e, int r) {
  if (r == 0) return 0;
  if (r <= 2 * NT) { const int k = r - 1; return (k >> 1) == e ? 1 + (k & 1) : -1; }
  return r - 2 * NT + 2;  // S -> 3, D_a -> 4 + a, T_a -> 7 + a
}

// Staging of 16 head-dim columns of one tensor.  A thread owns the 16-byte piece q4 of jet row r of electrons
// e = egrp, egrp + EPP, ...: it issues the asynchronous copies of its pieces (global fp32 -> raw area), and later
// converts the same pieces (raw fp32 -> fp16 hi / lo planes), so a thread only ever waits for its own copies.
template <int NT, bool L0>
struct AtStager {
  using G = AtGeom<NT>;
  int r, q4, egrp;
  bool active;
  float amax;                 // largest |x| converted by this thread (fp16 pieces saturate beyond 65504)
  uint32_t raw_off, pl_off;   // of electron egrp
  int64_t g_own;              // global float offset of this thread's piece of electron egrp (full form with a constant stride)
  __device__ __forceinline__ void init(int64_t ld) {
    const int slot = threadIdx.x % G::SLOTS;
    egrp = threadIdx.x / G::SLOTS;
    active = egrp < G::EPP;
    amax = 0.f;
    r = slot >> 2;
    q4 = slot & 3;
    raw_off = (uint32_t)(((egrp * G::R + r) * 4 + q4) * 16);
    pl_off = (uint32_t)(egrp * (G::R * 32 + 16) + r * 32 + (((q4 >> 1) ^ ((r >> 2) & 1)) << 4) + ((q4 & 1) << 3));
    g_own = (int64_t)(egrp * G::R + r) * ld + 4 * q4;
  }
  // global -> raw (asynchronous); src points at column col0 of row 0 of this walker's tensor.  LDT = the row stride as a
  // compile-time constant (0: use ld): the copies of a thread then differ by immediates from ONE address (g_own)
  template <int LDT>
  __device__ __forceinline__ void issue(uint32_t raw_s, const float* __restrict__ src, int64_t ld) const {
    if (active) {
      if constexpr (!L0 && LDT > 0) {
        const float* gp0 = src + g_own;
        const uint32_t dst0 = raw_s + raw_off;
#pragma unroll
        for (int k = 0; k * G::EPP < G::N; ++k) {
          if (egrp + k * G::EPP < G::N)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(k * G::EPP * G::R * 64)),
                         "l"(gp0 + (int64_t)k * (G::EPP * G::R) * LDT) : "memory");
        }
      } else {
#pragma unroll
        for (int e = egrp, k = 0; e < G::N; e += G::EPP, ++k) {
          const int rc = L0 ? l0_row<NT>(e, r) : r;
          if (rc >= 0) {
            const float* gp = src + (int64_t)(e * (L0 ? AT_RC : G::R) + rc) * ld + 4 * q4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(raw_s + raw_off + (uint32_t)(k * G::EPP * G::R * 64)), "l"(gp) : "memory");
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // raw -> planes (own pieces; call after cp.async.wait_group 0)
  __device__ __forceinline__ void convert(const uint8_t* raw, uint8_t* hi, uint8_t* lo) {
    if (active) {
#pragma unroll
      for (int e = egrp, k = 0; e < G::N; e += G::EPP, ++k) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!L0 || l0_row<NT>(e, r) >= 0) v = *reinterpret_cast<const float4*>(raw + raw_off + k * G::EPP * G::R * 64);
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        uint2 h, l;
        split_f16x2(v.x, v.y, h.x, l.x);
        split_f16x2(v.z, v.w, h.y, l.y);
        const uint32_t off = pl_off + (uint32_t)(k * G::EPP * (G::R * 32 + 16));
        *reinterpret_cast<uint2*>(hi + off) = h;
        *reinterpret_cast<uint2*>(lo + off) = l;
      }
    }
  }
};

// Softmax jets (log-sum-exp Hessian = diag(p) - p p^T) on the score jets in shared memory: on return sj holds
// l^(r) = p^(r) / p for r > 0 and p0 the probabilities; qq, dd are scratch.  Ends with a barrier.
template <int NT>
__device__ __forceinline__ void softmax_jets(float* sj, float* p0, float* qq, float* dd) {
  using G = AtGeom<NT>;
  constexpr int N = G::N, NP = AT_NP;
  const int tid = threadIdx.x;
  Rows rw(N, true);
#define SJ(i, j, r) ((j) * G::SJ_J + (r) * G::SJ_R + (i))
  for (int i = tid; i < N; i += AT_THREADS) {
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, sj[SJ(i, j, 0)]);
    float Z = 0.f;
    for (int j = 0; j < N; ++j) { const float e = expf(sj[SJ(i, j, 0)] - mx); p0[j * NP + i] = e; Z += e; }
    const float iz = 1.f / Z;
    for (int j = 0; j < N; ++j) p0[j * NP + i] *= iz;
  }
  if constexpr (NP > N) {  // padded query slots
    for (int t = tid; t < N * (NP - N); t += AT_THREADS) p0[(t / (NP - N)) * NP + N + t % (NP - N)] = 0.f;
  }
  __syncthreads();
  constexpr int nfirst = 2 * N + 3;
  for (int t = tid; t < N * nfirst; t += AT_THREADS) {  // first-order rows: l = s - lse
    const int i = t % N, q = t / N;
    const int r = q < 2 * N ? rw.J(q) : rw.D(q - 2 * N);
    float lse = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) lse = fmaf(p0[j * NP + i], sj[SJ(i, j, r)], lse);
#pragma unroll
    for (int j = 0; j < N; ++j) sj[SJ(i, j, r)] -= lse;
  }
  __syncthreads();
  for (int t = tid; t < N * N; t += AT_THREADS) {
    const int i = t % N, j = t / N;
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * N; ++k) { const float l = sj[SJ(i, j, 1 + k)]; s2 = fmaf(l, l, s2); }
    qq[j * NP + i] = s2;
    for (int a3 = 0; a3 < 3; ++a3) { const float l = sj[SJ(i, j, rw.D(a3))]; dd[(a3 * N + j) * NP + i] = l * l; }
  }
  __syncthreads();
  for (int t = tid; t < N * 4; t += AT_THREADS) {  // second-order rows
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
  __syncthreads();  // sj holds l^(r) = p^(r) / p for r > 0; the p jets are formed in the fragments: p^(r) = p l^(r)
#undef SJ
}

template <int NT, bool L0, int DT>
__global__ void __launch_bounds__(AT_THREADS, NT <= 12 ? 2 : 1)
attention_jets_tc_kernel(const float* __restrict__ qkv, float* __restrict__ o, NetDims dm, unsigned* __restrict__ rflag) {
  using G = AtGeom<NT>;
  constexpr int N = G::N, R = G::R, NR = G::NR, NP = AT_NP;
  extern __shared__ __align__(128) uint8_t smem_at[];
  uint8_t* planes = smem_at;
  uint8_t* raw = smem_at + G::OFF_RAW;
  float* sj = reinterpret_cast<float*>(smem_at + G::OFF_SJ);
  float* p0 = reinterpret_cast<float*>(smem_at + G::OFF_P0);   // [j][NP]
  float* red = reinterpret_cast<float*>(smem_at + G::OFF_RED);  // [warp][lane][8]
  float* qq = reinterpret_cast<float*>(planes);                  // [j][NP]      (softmax scratch, in the idle planes)
  float* dd = qq + N * NP;                                       // [3][j][NP]
  float* xw = dd + 3 * N * NP;                                   // [warp][16 x 16]
  float* dw = xw + AT_WARPS * 256;                               // [3][16 x 16]
#define SJ(i, j, r) ((j) * G::SJ_J + (r) * G::SJ_R + (i))
  // DT = model width as a compile-time constant (0 = read it from dm): with constant row strides the staging copies and the
  // output stores of a thread differ by immediates, one address register pair serves them all -- the copies of a step used
  // to wait for one another's address registers (long-scoreboard stalls on the address arithmetic, 10 % of the samples)
  const int D = DT ? DT : dm.D;
  const int hh = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  constexpr int RI = L0 ? AT_RC : R;
  const int64_t ld = 3 * (int64_t)D;
  const float* qbase = qkv + b * (int64_t)(N * RI) * ld + hh * AT_HD;
  const float* kbase = qbase + D;
  const float* vbase = qbase + 2 * D;
  const float scl = rsqrtf((float)AT_HD);
  Rows rw(N, true);
  AtStager<NT, L0> stg;
  stg.init(ld);
  const uint32_t raw_s = sm_u32(raw), pl_s = sm_u32(planes);
  AT_PROF_DECL

  // ------------------------------------------------------------------ phase 1: score jets on the tensor cores
  {
    const uint32_t qh_s = pl_s, ql_s = pl_s + G::PL, kh_s = pl_s + 2 * G::PL, kl_s = pl_s + 3 * G::PL;
    float g1[G::TPW][2][4], g2[G::TPW][2][4], xa[2][4], da[2][4];
#pragma unroll
    for (int ts = 0; ts < G::TPW; ++ts)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) { g1[ts][n][c] = 0.f; g2[ts][n][c] = 0.f; }
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) { xa[n][c] = 0.f; da[n][c] = 0.f; }
    // lane roles of the two ldmatrix address patterns
    const int a_r = (lane & 7) + ((lane >> 3) & 1) * 8, a_c = lane >> 4;        // A: matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15)
    const int b_n = (lane & 7) + (lane >> 4) * 8, b_c = (lane >> 3) & 1;        // B: matrices (n 0-7: k 0-7, k 8-15 | n 8-15: ...)
    const int b_e = b_n < N ? b_n : N - 1, a_e = a_r < N ? a_r : N - 1;
    const uint32_t b0_off = plane_off<R, 32>(b_e * R, b_c);  // value rows of the electrons: right operands of G1 / G2
    uint32_t t_off[G::TPW];
#pragma unroll
    for (int ts = 0; ts < G::TPW; ++ts) {
      int arow = (warp + AT_WARPS * ts) * 16 + a_r;
      arow = arow < NR ? arow : NR - 1;
      t_off[ts] = plane_off<R, 32>(arow, a_c);
    }
    const uint32_t xa_off = (uint32_t)(a_e * (R * 32 + 16)), xb_off = (uint32_t)(b_e * (R * 32 + 16));  // + row term of the flow
    stg.template issue<3 * DT>(raw_s, qbase, ld);
    stg.template issue<3 * DT>(raw_s + G::RAW, kbase, ld);
#pragma unroll 1
    for (int st = 0; st < AT_HD / 16; ++st) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();  // the planes are free
      stg.convert(raw, planes, planes + G::PL);
      stg.convert(raw + G::RAW, planes + 2 * G::PL, planes + 3 * G::PL);
      __syncthreads();
      AT_LAP(2)
      if (st + 1 < AT_HD / 16) {  // the next 16 columns arrive while this stage is multiplied
        stg.template issue<3 * DT>(raw_s, qbase + (st + 1) * 16, ld);
        stg.template issue<3 * DT>(raw_s + G::RAW, kbase + (st + 1) * 16, ld);
      } else {
        for (int t = tid; t < G::SJ_FLOATS; t += AT_THREADS) sj[t] = 0.f;  // (the raw area of k ran over sj until now)
      }
      uint32_t k0h[4], k0l[4], q0h[4], q0l[4];
      ldsm_x4(kh_s + b0_off, k0h); ldsm_x4(kl_s + b0_off, k0l);
      ldsm_x4(qh_s + b0_off, q0h); ldsm_x4(ql_s + b0_off, q0l);
#pragma unroll
      for (int ts = 0; ts < G::TPW; ++ts) {
        if (warp + AT_WARPS * ts < G::MT) {
          uint32_t ah[4], al[4], ch[4], cl[4];
          ldsm_x4(qh_s + t_off[ts], ah); ldsm_x4(ql_s + t_off[ts], al);
          ldsm_x4(kh_s + t_off[ts], ch); ldsm_x4(kl_s + t_off[ts], cl);
          // four independent accumulators, the three products of each interleaved
          mma16816(g1[ts][0], al, k0h[0], k0h[1]); mma16816(g1[ts][1], al, k0h[2], k0h[3]);
          mma16816(g2[ts][0], cl, q0h[0], q0h[1]); mma16816(g2[ts][1], cl, q0h[2], q0h[3]);
          mma16816(g1[ts][0], ah, k0l[0], k0l[1]); mma16816(g1[ts][1], ah, k0l[2], k0l[3]);
          mma16816(g2[ts][0], ch, q0l[0], q0l[1]); mma16816(g2[ts][1], ch, q0l[2], q0l[3]);
          mma16816(g1[ts][0], ah, k0h[0], k0h[1]); mma16816(g1[ts][1], ah, k0h[2], k0h[3]);
          mma16816(g2[ts][0], ch, q0h[0], q0h[1]); mma16816(g2[ts][1], ch, q0h[2], q0h[3]);
        }
      }
      // cross terms: flow f < 2N -> X (S row), flow 2N + a -> D_a (T_a row)
#pragma unroll 1
      for (int f = warp; f < 2 * N + 3; f += AT_WARPS) {
        const int rf = f < 2 * N ? 1 + f : f + 2;  // J(f) | D(f - 2N)
        uint32_t ah[4], al[4], bh[4], bl[4];
        const uint32_t aoff = xa_off + (uint32_t)(rf * 32 + ((a_c ^ ((rf >> 2) & 1)) << 4));
        const uint32_t boff = xb_off + (uint32_t)(rf * 32 + ((b_c ^ ((rf >> 2) & 1)) << 4));
        ldsm_x4(qh_s + aoff, ah); ldsm_x4(ql_s + aoff, al);
        ldsm_x4(kh_s + boff, bh); ldsm_x4(kl_s + boff, bl);
        if (f < 2 * N) {
          mma16816(xa[0], al, bh[0], bh[1]); mma16816(xa[1], al, bh[2], bh[3]);
          mma16816(xa[0], ah, bl[0], bl[1]); mma16816(xa[1], ah, bl[2], bl[3]);
          mma16816(xa[0], ah, bh[0], bh[1]); mma16816(xa[1], ah, bh[2], bh[3]);
        } else {
          mma16816(da[0], al, bh[0], bh[1]); mma16816(da[1], al, bh[2], bh[3]);
          mma16816(da[0], ah, bl[0], bl[1]); mma16816(da[1], ah, bl[2], bl[3]);
          mma16816(da[0], ah, bh[0], bh[1]); mma16816(da[1], ah, bh[2], bh[3]);
        }
      }
      AT_LAP(3)
    }
    __syncthreads();  // every warp is done with the planes (xw, dw live there) and sj is zeroed
    // accumulators -> sj: every entry has exactly one G1 term (stored first) and, for r > 0, one G2 term (added after a
    // barrier by the thread that owns it) -- no atomics, a fixed order of additions
#pragma unroll
    for (int ts = 0; ts < G::TPW; ++ts) {
      const int mt = warp + AT_WARPS * ts;
      if (mt < G::MT) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = mt * 16 + g + 8 * h;
          if (row < NR) {
            const int e = row / R, r = row - e * R;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const int x = nt * 8 + 2 * t4 + c;
                if (x < N) sj[SJ(e, x, r)] = g1[ts][nt][2 * h + c] * scl;  // q_e^(r) . k_x
              }
          }
        }
      }
    }
    const int fd = (warp - (2 * N) % AT_WARPS + AT_WARPS) % AT_WARPS;  // this warp's D flow, if < 3
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int idx = (g + 8 * (c >> 1)) * 16 + nt * 8 + 2 * t4 + (c & 1);  // (i, j)
        xw[warp * 256 + idx] = xa[nt][c];
        if (fd < 3) dw[fd * 256 + idx] = da[nt][c];
      }
    stg.template issue<3 * DT>(raw_s, vbase, ld);  // the first 16 columns of v arrive during the softmax
    __syncthreads();
#pragma unroll
    for (int ts = 0; ts < G::TPW; ++ts) {
      const int mt = warp + AT_WARPS * ts;
      if (mt < G::MT) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = mt * 16 + g + 8 * h;
          if (row < NR) {
            const int e = row / R, r = row - e * R;
            if (r != 0) {
              float* dst = sj + SJ(0, e, r);  // k_e^(r) . q_x : consecutive x
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const int x = nt * 8 + 2 * t4 + c;
                  if (x < N) dst[x] += g2[ts][nt][2 * h + c] * scl;
                }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // second-order rows pick up the cross products: S += 2 sum_k qJk.kJk ; T_a += 2 qDa.kDa
  for (int t = tid; t < N * N * 4; t += AT_THREADS) {
    const int w = t & 3, ij = t >> 2;
    const int i = ij % N, j = ij / N;
    if (w == 0) {
      float s2 = 0.f;
#pragma unroll
      for (int u = 0; u < AT_WARPS; ++u) s2 += xw[u * 256 + i * 16 + j];
      sj[SJ(i, j, rw.S())] += 2.f * scl * s2;
    } else {
      sj[SJ(i, j, rw.T(w - 1))] += 2.f * scl * dw[(w - 1) * 256 + i * 16 + j];
    }
  }
  __syncthreads();
  // ------------------------------------------------------------------ phase 2: softmax jets (as attention_jets.cu)
  AT_LAP(4)
  softmax_jets<NT>(sj, p0, qq, dd);
  AT_LAP(5)
  // ------------------------------------------------------------------ phase 3: o = P V jets on the tensor cores
  {
    const uint32_t vh_s = pl_s, vl_s = pl_s + G::PL;
    float* obase = o + b * (int64_t)NR * D + hh * AT_HD + 2 * t4;  // this lane's column pair; + ((i R + r) D + 16 qt)
    const int rS = rw.S(), rT0 = rw.T(0);
    const int ob0 = g * R * D, ob1 = (g + 8) * R * D;
    // p of this lane's fragment slots: (i = g | g + 8) x (j = 2t, 2t + 1, 2t + 8, 2t + 9); zero outside the N x N block
    float pw[2][4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 2 * t4 + (jj & 1) + 8 * (jj >> 1);
      pw[0][jj] = j < N ? p0[j * NP + g] : 0.f;
      pw[1][jj] = j < N ? p0[j * NP + g + 8] : 0.f;
    }
    // A fragment of P^(r): a0 (i = g, j = 2t, 2t+1) a1 (i = g + 8, same j) a2 (i = g, j = 2t + 8, 2t + 9) a3 (i = g + 8, ...)
    auto pfrag = [&](int r, uint32_t (&ph)[4], uint32_t (&pl)[4]) {
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int j0 = 2 * t4 + 8 * jh;
        float x0 = pw[0][2 * jh], x1 = pw[0][2 * jh + 1], y0 = pw[1][2 * jh], y1 = pw[1][2 * jh + 1];
        if (r != 0) {
          if (j0 < N) { x0 *= sj[SJ(g, j0, r)]; y0 *= sj[SJ(g + 8, j0, r)]; }
          if (j0 + 1 < N) { x1 *= sj[SJ(g, j0 + 1, r)]; y1 *= sj[SJ(g + 8, j0 + 1, r)]; }
        }
        stg.amax = fmaxf(stg.amax, fmaxf(fmaxf(fabsf(x0), fabsf(x1)), fmaxf(fabsf(y0), fabsf(y1))));
        split_f16x2(x0, x1, ph[2 * jh], pl[2 * jh]);
        split_f16x2(y0, y1, ph[2 * jh + 1], pl[2 * jh + 1]);
      }
    };
    // the P fragments of this warp's rows r = warp + 8 k stay in registers for the four column steps
    uint32_t p0h[4], p0l[4], prh[G::RPW][4], prl[G::RPW][4], pdh[4], pdl[4];
    pfrag(0, p0h, p0l);
#pragma unroll
    for (int k = 0; k < G::RPW; ++k) {
      const int r = warp + AT_WARPS * k;
      if (r < R) pfrag(r, prh[k], prl[k]);
      if (r < R && r >= rT0) pfrag(r - 3, pdh, pdl);  // (at most one T row per warp: they are three consecutive rows)
    }
    // B fragments of V^(r), 16 columns, transposed ldmatrix: matrices (j 0-7 | 8-15) x (d 0-7 | 8-15)
    const int v_j = (lane & 7) + ((lane >> 3) & 1) * 8, v_c = lane >> 4;
    const int v_e = v_j < N ? v_j : N - 1;
    const uint32_t v_off = (uint32_t)(v_e * (R * 32 + 16));
    auto vfrag = [&](int r, uint32_t (&fh)[4], uint32_t (&fl)[4]) {
      const uint32_t off = v_off + (uint32_t)(r * 32 + ((v_c ^ ((r >> 2) & 1)) << 4));
      ldsm_x4_t(vh_s + off, fh);
      ldsm_x4_t(vl_s + off, fl);
    };
    float accS[2][4];  // the S row of this column step, held by the warp that owns row S until the partials are complete
    auto finish_s = [&](int qt) {
      if (warp == rS % AT_WARPS) {
#pragma unroll
        for (int u = 0; u < AT_WARPS; ++u) {
          const float4 x = *reinterpret_cast<const float4*>(red + (u * 32 + lane) * 8);
          const float4 y = *reinterpret_cast<const float4*>(red + (u * 32 + lane) * 8 + 4);
          accS[0][0] = fmaf(2.f, x.x, accS[0][0]); accS[0][1] = fmaf(2.f, x.y, accS[0][1]);
          accS[0][2] = fmaf(2.f, x.z, accS[0][2]); accS[0][3] = fmaf(2.f, x.w, accS[0][3]);
          accS[1][0] = fmaf(2.f, y.x, accS[1][0]); accS[1][1] = fmaf(2.f, y.y, accS[1][1]);
          accS[1][2] = fmaf(2.f, y.z, accS[1][2]); accS[1][3] = fmaf(2.f, y.w, accS[1][3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (g + 8 * h < N) {
            float* dst = obase + ((h ? ob1 : ob0) + rS * D + qt * 16);
            *reinterpret_cast<float2*>(dst) = make_float2(accS[0][2 * h], accS[0][2 * h + 1]);
            *reinterpret_cast<float2*>(dst + 8) = make_float2(accS[1][2 * h], accS[1][2 * h + 1]);
          }
        }
      }
    };
#pragma unroll 1
    for (int qt = 0; qt < AT_HD / 16; ++qt) {
      if (qt == 0) { AT_LAP(6) } else { AT_LAP(8) }
      // v steps arrive two deep: step qt lies in raw half qt & 1 (the second half runs over sj, which is dead once the P
      // fragments are in registers), steps qt + 1 and, from here on, qt + 2 are in flight while step qt is multiplied
      if (qt == 0 || qt + 1 == AT_HD / 16) asm volatile("cp.async.wait_group 0;" ::: "memory");
      else asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();  // the planes are free (softmax scratch / previous step's fragments) and `red` is complete
      if (qt > 0) finish_s(qt - 1);
      stg.convert(raw + (qt & 1) * G::RAW, planes, planes + G::PL);
      __syncthreads();  // (also: `red` has been read before this step's partials overwrite it)
      AT_LAP(7)
      if (qt == 0 && AT_HD / 16 > 1) stg.template issue<3 * DT>(raw_s + G::RAW, vbase + 16, ld);
      if (qt + 2 < AT_HD / 16) stg.template issue<3 * DT>(raw_s + (qt & 1) * G::RAW, vbase + (qt + 2) * 16, ld);
      uint32_t v0h[4], v0l[4];
      vfrag(0, v0h, v0l);
      float sx[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) { sx[n][c] = 0.f; accS[n][c] = 0.f; }
#pragma unroll
      for (int k = 0; k < G::RPW; ++k) {
        const int r = warp + AT_WARPS * k;
        const bool allJ = (AT_WARPS * k >= 1) && (AT_WARPS * k + AT_WARPS - 1 <= 2 * N);  // compile-time per k: pure J-row group
        if (allJ || r < R) {
          float acc[2][4];
#pragma unroll
          for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[n][c] = 0.f;
          if (!allJ && r == 0) {
            mma16816(acc[0], p0l, v0h[0], v0h[1]); mma16816(acc[1], p0l, v0h[2], v0h[3]);
            mma16816(acc[0], p0h, v0l[0], v0l[1]); mma16816(acc[1], p0h, v0l[2], v0l[3]);
            mma16816(acc[0], p0h, v0h[0], v0h[1]); mma16816(acc[1], p0h, v0h[2], v0h[3]);
          } else {
            uint32_t fh[4], fl[4];
            if (!allJ && r >= rT0) {  // 2 P^(D_a) V^(D_a) first, doubled once in the accumulators
              vfrag(r - 3, fh, fl);
              mma16816(acc[0], pdl, fh[0], fh[1]); mma16816(acc[1], pdl, fh[2], fh[3]);
              mma16816(acc[0], pdh, fl[0], fl[1]); mma16816(acc[1], pdh, fl[2], fl[3]);
              mma16816(acc[0], pdh, fh[0], fh[1]); mma16816(acc[1], pdh, fh[2], fh[3]);
#pragma unroll
              for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[n][c] *= 2.f;
            }
            vfrag(r, fh, fl);
            const bool isJ = allJ || r <= 2 * N;
            // P^(r) V^(0) + P^(0) V^(r)  (+ S-row cross term P^(Jk) V^(Jk)), small products first, accumulators interleaved
            mma16816(acc[0], prl[k], v0h[0], v0h[1]); mma16816(acc[1], prl[k], v0h[2], v0h[3]);
            if (isJ) { mma16816(sx[0], prl[k], fh[0], fh[1]); mma16816(sx[1], prl[k], fh[2], fh[3]); }
            mma16816(acc[0], p0l, fh[0], fh[1]); mma16816(acc[1], p0l, fh[2], fh[3]);
            if (isJ) { mma16816(sx[0], prh[k], fl[0], fl[1]); mma16816(sx[1], prh[k], fl[2], fl[3]); }
            mma16816(acc[0], prh[k], v0l[0], v0l[1]); mma16816(acc[1], prh[k], v0l[2], v0l[3]);
            mma16816(acc[0], p0h, fl[0], fl[1]); mma16816(acc[1], p0h, fl[2], fl[3]);
            if (isJ) { mma16816(sx[0], prh[k], fh[0], fh[1]); mma16816(sx[1], prh[k], fh[2], fh[3]); }
            mma16816(acc[0], prh[k], v0h[0], v0h[1]); mma16816(acc[1], prh[k], v0h[2], v0h[3]);
            mma16816(acc[0], p0h, fh[0], fh[1]); mma16816(acc[1], p0h, fh[2], fh[3]);
          }
          if (!allJ && r == rS) {
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
              for (int c = 0; c < 4; ++c) accS[n][c] = acc[n][c];
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (g + 8 * h < N) {
                float* dst = obase + ((h ? ob1 : ob0) + r * D + qt * 16);
                *reinterpret_cast<float2*>(dst) = make_float2(acc[0][2 * h], acc[0][2 * h + 1]);
                *reinterpret_cast<float2*>(dst + 8) = make_float2(acc[1][2 * h], acc[1][2 * h + 1]);
              }
            }
          }
        }
      }
      // S row: + 2 sum over all warps' cross partials, summed in warp order (fragment order: 8 floats per lane)
      *reinterpret_cast<float4*>(red + (warp * 32 + lane) * 8) = make_float4(sx[0][0], sx[0][1], sx[0][2], sx[0][3]);
      *reinterpret_cast<float4*>(red + (warp * 32 + lane) * 8 + 4) = make_float4(sx[1][0], sx[1][1], sx[1][2], sx[1][3]);
    }
    __syncthreads();
    finish_s(AT_HD / 16 - 1);
  }
  // the fp16 pieces have a narrower range than the reference's fp32: a saturated operand is reported (dh_plan_status)
  if (rflag != nullptr && !(stg.amax <= 65504.f)) atomicOr(rflag, 1u);
  AT_LAP(8)
  AT_PROF_END
#undef SJ
}

// =============================================================================================================
// FIRST-LAYER kernel.  q, k, v of the first layer depend on their own electron only, so their jets have just 10
// non-zero rows per electron (value | own flows (2) | S | D_a | T_a) and arrive compressed [B N 10][3 D].  Instead of
// expanding them to the full row set (two thirds of the products would multiply zeros) the contractions run on the
// compressed rows:
//   scores   U_c[(i,c), j] = q_i^(c) . k_j,  W_c[(j,c), i] = k_j^(c) . q_i      [10 N x 64] x [64 x N]   (3.2x fewer rows)
//            s^(J(2e+t))_ij = [i = e] U_c[(e,1+t), j] + [j = e] W_c[(e,1+t), i];   S, D_a, T_a rows as in the full form;
//            cross terms: own flows only touch the diagonal (S_ii += 2 sum_t q_i^(1+t) . k_i^(1+t)), D_a as before
//   P.V      o^(r) = P^(r) V^(0) on the tensor cores for every row; P^(0) V^(r) is a rank-one update for the 2N own-flow
//            rows (v^(J(2e+t)) lives on electron e only) and a tensor-core product for S, D_a, T_a; the S-row cross
//            term is the sum of 2N rank-one updates.
// 1,536 HMMA per block instead of 4,776, a third of the staging conversions.
// =============================================================================================================
template <int NT>
struct AtGeomC {
  static constexpr int N = NT, R = 2 * NT + 8, RC = AT_RC, NRC = N * RC, NR = N * R;
  static constexpr int MT = (NRC + 15) / 16, TPW = (MT + AT_WARPS - 1) / AT_WARPS;
  static constexpr int RPW = (R + AT_WARPS - 1) / AT_WARPS;
  static constexpr int EB = RC * 32 + 16;                  // bytes per electron of a 16-column fp16 plane
  static constexpr int PL = N * EB;
  static constexpr int SCRATCH = 4 * (5 * N * AT_NP + 11 * 256);
  static constexpr int PLANES = 4 * PL > SCRATCH ? 4 * PL : SCRATCH;
  static constexpr int RAW = NRC * 64;
  static constexpr int OFF_RAW = PLANES, OFF_SJ = OFF_RAW + 2 * RAW, OFF_P0 = OFF_SJ + 4 * AtGeom<NT>::SJ_FLOATS;
  static constexpr int OFF_RED = OFF_P0 + 4 * N * AT_NP;
  static constexpr size_t SMEM = OFF_RED + 4 * AT_WARPS * 256;
  static constexpr int SLOTS = RC * 4, EPP = AT_THREADS / SLOTS;
};
template <int NT>
__device__ __forceinline__ uint32_t plane_off_c(int e, int c, int chunk) {
  return (uint32_t)(e * AtGeomC<NT>::EB + c * 32 + ((chunk ^ ((c >> 2) & 1)) << 4));
}
template <int NT>
struct AtStagerC {
  using G = AtGeomC<NT>;
  int c, q4, egrp;
  bool active;
  float amax;
  uint32_t raw_off, pl_off;
  int64_t g_own;
  __device__ __forceinline__ void init(int64_t ld) {
    const int slot = threadIdx.x % G::SLOTS;
    egrp = threadIdx.x / G::SLOTS;
    active = egrp < G::EPP;
    amax = 0.f;
    c = slot >> 2;
    q4 = slot & 3;
    raw_off = (uint32_t)(((egrp * G::RC + c) * 4 + q4) * 16);
    pl_off = plane_off_c<NT>(egrp, c, q4 >> 1) + (uint32_t)((q4 & 1) << 3);
    g_own = (int64_t)(egrp * G::RC + c) * ld + 4 * q4;
  }
  template <int LDT>
  __device__ __forceinline__ void issue(uint32_t raw_s, const float* __restrict__ src, int64_t ld) const {
    if (active) {
      if constexpr (LDT > 0) {  // constant row stride: one address, immediates (see AtStager)
        const float* gp0 = src + g_own;
        const uint32_t dst0 = raw_s + raw_off;
#pragma unroll
        for (int k = 0; k * G::EPP < G::N; ++k) {
          if (egrp + k * G::EPP < G::N)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(k * G::EPP * G::RC * 64)),
                         "l"(gp0 + (int64_t)k * (G::EPP * G::RC) * LDT) : "memory");
        }
      } else {
#pragma unroll
        for (int e = egrp, k = 0; e < G::N; e += G::EPP, ++k) {
          const float* gp = src + (int64_t)(e * G::RC + c) * ld + 4 * q4;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(raw_s + raw_off + (uint32_t)(k * G::EPP * G::RC * 64)), "l"(gp) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __device__ __forceinline__ void convert(const uint8_t* raw, uint8_t* hi, uint8_t* lo) {
    if (active) {
#pragma unroll
      for (int e = egrp, k = 0; e < G::N; e += G::EPP, ++k) {
        const float4 v = *reinterpret_cast<const float4*>(raw + raw_off + k * G::EPP * G::RC * 64);
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        uint2 h, l;
        split_f16x2(v.x, v.y, h.x, l.x);
        split_f16x2(v.z, v.w, h.y, l.y);
        const uint32_t off = pl_off + (uint32_t)(k * G::EPP * G::EB);
        *reinterpret_cast<uint2*>(hi + off) = h;
        *reinterpret_cast<uint2*>(lo + off) = l;
      }
    }
  }
};

template <int NT, int DT>
__global__ void __launch_bounds__(AT_THREADS, NT <= 12 ? 2 : 1)
attention_jets_tc_l0_kernel(const float* __restrict__ qkv, float* __restrict__ o, NetDims dm, unsigned* __restrict__ rflag) {
  using G = AtGeomC<NT>;
  using GF = AtGeom<NT>;
  constexpr int N = G::N, R = G::R, RC = G::RC, NRC = G::NRC, NP = AT_NP;
  extern __shared__ __align__(128) uint8_t smem_at[];
  uint8_t* planes = smem_at;
  uint8_t* raw = smem_at + G::OFF_RAW;
  float* sj = reinterpret_cast<float*>(smem_at + G::OFF_SJ);
  float* p0 = reinterpret_cast<float*>(smem_at + G::OFF_P0);
  float* red = reinterpret_cast<float*>(smem_at + G::OFF_RED);
  float* qq = reinterpret_cast<float*>(planes);
  float* dd = qq + N * NP;
  float* xw = dd + 3 * N * NP;      // [warp][16 x 16]: own-flow cross products (warps 0, 1)
  float* dw = xw + AT_WARPS * 256;  // [3][16 x 16]
#define SJ(i, j, r) ((j) * GF::SJ_J + (r) * GF::SJ_R + (i))
  const int D = DT ? DT : dm.D;  // (compile-time row strides: see attention_jets_tc_kernel)
  const int hh = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int64_t ld = 3 * (int64_t)D;
  const float* qbase = qkv + b * (int64_t)NRC * ld + hh * AT_HD;
  const float* kbase = qbase + D;
  const float* vbase = qbase + 2 * D;
  const float scl = rsqrtf((float)AT_HD);
  Rows rw(N, true);
  const int rS = rw.S(), rD0 = rw.D(0), rT0 = rw.T(0);
  // full jet row of compressed row c of electron e
  auto full_row = [&](int e, int c) { return c == 0 ? 0 : (c <= 2 ? 2 * e + c : (c == 3 ? rS : (c <= 6 ? rD0 + (c - 4) : rT0 + (c - 7)))); };
  AtStagerC<NT> stg;
  stg.init(ld);
  AT_PROF_DECL
  const uint32_t raw_s = sm_u32(raw), pl_s = sm_u32(planes);
  for (int t = tid; t < GF::SJ_FLOATS; t += AT_THREADS) sj[t] = 0.f;

  // ------------------------------------------------------------------ phase 1: score jets of the compressed rows
  {
    const uint32_t qh_s = pl_s, ql_s = pl_s + G::PL, kh_s = pl_s + 2 * G::PL, kl_s = pl_s + 3 * G::PL;
    float g1[G::TPW][2][4], g2[G::TPW][2][4], xa[2][4], da[2][4];
#pragma unroll
    for (int ts = 0; ts < G::TPW; ++ts)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) { g1[ts][n][c] = 0.f; g2[ts][n][c] = 0.f; }
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) { xa[n][c] = 0.f; da[n][c] = 0.f; }
    const int a_r = (lane & 7) + ((lane >> 3) & 1) * 8, a_c = lane >> 4;
    const int b_n = (lane & 7) + (lane >> 4) * 8, b_c = (lane >> 3) & 1;
    const int b_e = b_n < N ? b_n : N - 1, a_e = a_r < N ? a_r : N - 1;
    const uint32_t b0_off = plane_off_c<NT>(b_e, 0, b_c);
    uint32_t t_off[G::TPW];
#pragma unroll
    for (int ts = 0; ts < G::TPW; ++ts) {
      int arow = (warp + AT_WARPS * ts) * 16 + a_r;
      arow = arow < NRC ? arow : NRC - 1;
      t_off[ts] = plane_off_c<NT>(arow / RC, arow % RC, a_c);
    }
    // cross flows: warp 0, 1 -> own flows t = 0, 1 (compressed rows 1, 2; only the diagonal is used);
    //              warp 2 + a -> D_a (compressed row 4 + a)
    const int xc = warp < 2 ? 1 + warp : 4 + (warp - 2);
    const uint32_t xa_off = plane_off_c<NT>(a_e, xc, a_c), xb_off = plane_off_c<NT>(b_e, xc, b_c);
    stg.template issue<3 * DT>(raw_s, qbase, ld);
    stg.template issue<3 * DT>(raw_s + G::RAW, kbase, ld);
#pragma unroll 1
    for (int st = 0; st < AT_HD / 16; ++st) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      stg.convert(raw, planes, planes + G::PL);
      stg.convert(raw + G::RAW, planes + 2 * G::PL, planes + 3 * G::PL);
      __syncthreads();
      AT_LAP(2)
      if (st + 1 < AT_HD / 16) {
        stg.template issue<3 * DT>(raw_s, qbase + (st + 1) * 16, ld);
        stg.template issue<3 * DT>(raw_s + G::RAW, kbase + (st + 1) * 16, ld);
      }
      uint32_t k0h[4], k0l[4], q0h[4], q0l[4];
      ldsm_x4(kh_s + b0_off, k0h); ldsm_x4(kl_s + b0_off, k0l);
      ldsm_x4(qh_s + b0_off, q0h); ldsm_x4(ql_s + b0_off, q0l);
#pragma unroll
      for (int ts = 0; ts < G::TPW; ++ts) {
        if (warp + AT_WARPS * ts < G::MT) {
          uint32_t ah[4], al[4], ch[4], cl[4];
          ldsm_x4(qh_s + t_off[ts], ah); ldsm_x4(ql_s + t_off[ts], al);
          ldsm_x4(kh_s + t_off[ts], ch); ldsm_x4(kl_s + t_off[ts], cl);
          mma16816(g1[ts][0], al, k0h[0], k0h[1]); mma16816(g1[ts][1], al, k0h[2], k0h[3]);
          mma16816(g2[ts][0], cl, q0h[0], q0h[1]); mma16816(g2[ts][1], cl, q0h[2], q0h[3]);
          mma16816(g1[ts][0], ah, k0l[0], k0l[1]); mma16816(g1[ts][1], ah, k0l[2], k0l[3]);
          mma16816(g2[ts][0], ch, q0l[0], q0l[1]); mma16816(g2[ts][1], ch, q0l[2], q0l[3]);
          mma16816(g1[ts][0], ah, k0h[0], k0h[1]); mma16816(g1[ts][1], ah, k0h[2], k0h[3]);
          mma16816(g2[ts][0], ch, q0h[0], q0h[1]); mma16816(g2[ts][1], ch, q0h[2], q0h[3]);
        }
      }
      if (warp < 5) {
        uint32_t ah[4], al[4], bh[4], bl[4];
        ldsm_x4(qh_s + xa_off, ah); ldsm_x4(ql_s + xa_off, al);
        ldsm_x4(kh_s + xb_off, bh); ldsm_x4(kl_s + xb_off, bl);
        if (warp < 2) {
          mma16816(xa[0], al, bh[0], bh[1]); mma16816(xa[1], al, bh[2], bh[3]);
          mma16816(xa[0], ah, bl[0], bl[1]); mma16816(xa[1], ah, bl[2], bl[3]);
          mma16816(xa[0], ah, bh[0], bh[1]); mma16816(xa[1], ah, bh[2], bh[3]);
        } else {
          mma16816(da[0], al, bh[0], bh[1]); mma16816(da[1], al, bh[2], bh[3]);
          mma16816(da[0], ah, bl[0], bl[1]); mma16816(da[1], ah, bl[2], bl[3]);
          mma16816(da[0], ah, bh[0], bh[1]); mma16816(da[1], ah, bh[2], bh[3]);
        }
      }
      AT_LAP(3)
    }
    __syncthreads();  // every warp is done with the planes (xw, dw live there); sj was zeroed at the start
    // G1: q_e^(c) . k_x -> s^(r)_{e x}
#pragma unroll
    for (int ts = 0; ts < G::TPW; ++ts) {
      const int mt = warp + AT_WARPS * ts;
      if (mt < G::MT) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = mt * 16 + g + 8 * h;
          if (row < NRC) {
            const int e = row / RC, r = full_row(e, row - e * RC);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const int x = nt * 8 + 2 * t4 + c;
                if (x < N) sj[SJ(e, x, r)] = g1[ts][nt][2 * h + c] * scl;
              }
          }
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int idx = (g + 8 * (c >> 1)) * 16 + nt * 8 + 2 * t4 + (c & 1);  // (i, j)
        xw[warp * 256 + idx] = xa[nt][c];
        if (warp >= 2 && warp < 5) dw[(warp - 2) * 256 + idx] = da[nt][c];
      }
    stg.template issue<3 * DT>(raw_s, vbase, ld);
    __syncthreads();
    // G2: k_e^(c) . q_x -> s^(r)_{x e}   (r > 0; the entry (e, e, r) already holds its G1 term)
#pragma unroll
    for (int ts = 0; ts < G::TPW; ++ts) {
      const int mt = warp + AT_WARPS * ts;
      if (mt < G::MT) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = mt * 16 + g + 8 * h;
          if (row < NRC) {
            const int e = row / RC, c0 = row - e * RC;
            if (c0 != 0) {
              float* dst = sj + SJ(0, e, full_row(e, c0));
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const int x = nt * 8 + 2 * t4 + c;
                  if (x < N) dst[x] += g2[ts][nt][2 * h + c] * scl;
                }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // second-order rows pick up the cross products: S_ii += 2 sum_t q_i^(t) . k_i^(t) ; T_a += 2 qDa . kDa
  for (int t = tid; t < N * N * 4; t += AT_THREADS) {
    const int w = t & 3, ij = t >> 2;
    const int i = ij % N, j = ij / N;
    if (w == 0) {
      if (i == j) sj[SJ(i, i, rS)] += 2.f * scl * (xw[i * 16 + i] + xw[256 + i * 16 + i]);
    } else {
      sj[SJ(i, j, rw.T(w - 1))] += 2.f * scl * dw[(w - 1) * 256 + i * 16 + j];
    }
  }
  __syncthreads();
  // ------------------------------------------------------------------ phase 2: softmax jets
  AT_LAP(4)
  softmax_jets<NT>(sj, p0, qq, dd);
  AT_LAP(5)
  // ------------------------------------------------------------------ phase 3: o = P V jets
  {
    const uint32_t vh_s = pl_s, vl_s = pl_s + G::PL;
    const uint8_t* vh = planes;
    const uint8_t* vl = planes + G::PL;
    float* obase = o + b * (int64_t)G::NR * D + hh * AT_HD + 2 * t4;
    const int ob0 = g * R * D, ob1 = (g + 8) * R * D;
    float pw[2][4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 2 * t4 + (jj & 1) + 8 * (jj >> 1);
      pw[0][jj] = j < N ? p0[j * NP + g] : 0.f;
      pw[1][jj] = j < N ? p0[j * NP + g + 8] : 0.f;
    }
    auto pfrag = [&](int r, uint32_t (&ph)[4], uint32_t (&pl)[4]) {
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int j0 = 2 * t4 + 8 * jh;
        float x0 = pw[0][2 * jh], x1 = pw[0][2 * jh + 1], y0 = pw[1][2 * jh], y1 = pw[1][2 * jh + 1];
        if (r != 0) {
          if (j0 < N) { x0 *= sj[SJ(g, j0, r)]; y0 *= sj[SJ(g + 8, j0, r)]; }
          if (j0 + 1 < N) { x1 *= sj[SJ(g, j0 + 1, r)]; y1 *= sj[SJ(g + 8, j0 + 1, r)]; }
        }
        stg.amax = fmaxf(stg.amax, fmaxf(fmaxf(fabsf(x0), fabsf(x1)), fmaxf(fabsf(y0), fabsf(y1))));
        split_f16x2(x0, x1, ph[2 * jh], pl[2 * jh]);
        split_f16x2(y0, y1, ph[2 * jh + 1], pl[2 * jh + 1]);
      }
    };
    uint32_t p0h[4], p0l[4], prh[G::RPW][4], prl[G::RPW][4], pdh[4], pdl[4];
    pfrag(0, p0h, p0l);
#pragma unroll
    for (int k = 0; k < G::RPW; ++k) {
      const int r = warp + AT_WARPS * k;
      if (r < R) pfrag(r, prh[k], prl[k]);
      if (r < R && r >= rT0) pfrag(r - 3, pdh, pdl);
    }
    const int v_j = (lane & 7) + ((lane >> 3) & 1) * 8, v_c = lane >> 4;
    const int v_e = v_j < N ? v_j : N - 1;
    auto vfrag = [&](int c, uint32_t (&fh)[4], uint32_t (&fl)[4]) {  // compressed row c of every electron
      const uint32_t off = plane_off_c<NT>(v_e, c, v_c);
      ldsm_x4_t(vh_s + off, fh);
      ldsm_x4_t(vl_s + off, fl);
    };
    float accS[2][4];
    auto finish_s = [&](int qt) {
      if (warp == rS % AT_WARPS) {
#pragma unroll
        for (int u = 0; u < AT_WARPS; ++u) {
          const float4 x = *reinterpret_cast<const float4*>(red + (u * 32 + lane) * 8);
          const float4 y = *reinterpret_cast<const float4*>(red + (u * 32 + lane) * 8 + 4);
          accS[0][0] = fmaf(2.f, x.x, accS[0][0]); accS[0][1] = fmaf(2.f, x.y, accS[0][1]);
          accS[0][2] = fmaf(2.f, x.z, accS[0][2]); accS[0][3] = fmaf(2.f, x.w, accS[0][3]);
          accS[1][0] = fmaf(2.f, y.x, accS[1][0]); accS[1][1] = fmaf(2.f, y.y, accS[1][1]);
          accS[1][2] = fmaf(2.f, y.z, accS[1][2]); accS[1][3] = fmaf(2.f, y.w, accS[1][3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (g + 8 * h < N) {
            float* dst = obase + ((h ? ob1 : ob0) + rS * D + qt * 16);
            *reinterpret_cast<float2*>(dst) = make_float2(accS[0][2 * h], accS[0][2 * h + 1]);
            *reinterpret_cast<float2*>(dst + 8) = make_float2(accS[1][2 * h], accS[1][2 * h + 1]);
          }
        }
      }
    };
#pragma unroll 1
    for (int qt = 0; qt < AT_HD / 16; ++qt) {
      // v steps arrive two deep, step qt in raw half qt & 1 (see the full form)
      if (qt == 0) { AT_LAP(6) } else { AT_LAP(8) }
      if (qt == 0 || qt + 1 == AT_HD / 16) asm volatile("cp.async.wait_group 0;" ::: "memory");
      else asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();
      if (qt > 0) finish_s(qt - 1);
      stg.convert(raw + (qt & 1) * G::RAW, planes, planes + G::PL);
      __syncthreads();
      AT_LAP(7)
      if (qt == 0 && AT_HD / 16 > 1) stg.template issue<3 * DT>(raw_s + G::RAW, vbase + 16, ld);
      if (qt + 2 < AT_HD / 16) stg.template issue<3 * DT>(raw_s + (qt & 1) * G::RAW, vbase + (qt + 2) * 16, ld);
      uint32_t v0h[4], v0l[4];
      vfrag(0, v0h, v0l);
      float sx[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) { sx[n][c] = 0.f; accS[n][c] = 0.f; }
#pragma unroll
      for (int k = 0; k < G::RPW; ++k) {
        const int r = warp + AT_WARPS * k;
        if (r < R) {
          float acc[2][4];
#pragma unroll
          for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[n][c] = 0.f;
          if (r == 0) {
            mma16816(acc[0], p0l, v0h[0], v0h[1]); mma16816(acc[1], p0l, v0h[2], v0h[3]);
            mma16816(acc[0], p0h, v0l[0], v0l[1]); mma16816(acc[1], p0h, v0l[2], v0l[3]);
            mma16816(acc[0], p0h, v0h[0], v0h[1]); mma16816(acc[1], p0h, v0h[2], v0h[3]);
          } else if (r <= 2 * N) {
            // own-flow row of electron e: P^(r) V^(0) on the tensor cores, P^(0) V^(r) and the S-row cross term as rank-one
            // updates with v_e^(1+t) (this lane's four columns, hi + lo pieces back to fp32)
            mma16816(acc[0], prl[k], v0h[0], v0h[1]); mma16816(acc[1], prl[k], v0h[2], v0h[3]);
            mma16816(acc[0], prh[k], v0l[0], v0l[1]); mma16816(acc[1], prh[k], v0l[2], v0l[3]);
            mma16816(acc[0], prh[k], v0h[0], v0h[1]); mma16816(acc[1], prh[k], v0h[2], v0h[3]);
            const int e = (r - 1) >> 1, c = 1 + ((r - 1) & 1);
            float vv[2][2];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const uint32_t off = plane_off_c<NT>(e, c, nt) + (uint32_t)(4 * t4);  // columns 8 nt + 2 t4, + 1
              const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(vh + off));
              const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(vl + off));
              vv[nt][0] = fh.x + fl.x; vv[nt][1] = fh.y + fl.y;
            }
            const float pa = p0[e * NP + g], pb = p0[e * NP + g + 8];
            const float qa = pa * sj[SJ(g, e, r)], qb = pb * sj[SJ(g + 8, e, r)];  // P^(r)[i, e]
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int c2 = 0; c2 < 2; ++c2) {
                acc[nt][c2] = fmaf(pa, vv[nt][c2], acc[nt][c2]);
                acc[nt][2 + c2] = fmaf(pb, vv[nt][c2], acc[nt][2 + c2]);
                sx[nt][c2] = fmaf(qa, vv[nt][c2], sx[nt][c2]);
                sx[nt][2 + c2] = fmaf(qb, vv[nt][c2], sx[nt][2 + c2]);
              }
          } else {
            uint32_t fh[4], fl[4];
            const int c = r == rS ? 3 : (r < rT0 ? 4 + (r - rD0) : 7 + (r - rT0));
            if (r >= rT0) {  // 2 P^(D_a) V^(D_a) first, doubled once in the accumulators
              vfrag(c - 3, fh, fl);
              mma16816(acc[0], pdl, fh[0], fh[1]); mma16816(acc[1], pdl, fh[2], fh[3]);
              mma16816(acc[0], pdh, fl[0], fl[1]); mma16816(acc[1], pdh, fl[2], fl[3]);
              mma16816(acc[0], pdh, fh[0], fh[1]); mma16816(acc[1], pdh, fh[2], fh[3]);
#pragma unroll
              for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) acc[n][c2] *= 2.f;
            }
            vfrag(c, fh, fl);
            mma16816(acc[0], prl[k], v0h[0], v0h[1]); mma16816(acc[1], prl[k], v0h[2], v0h[3]);
            mma16816(acc[0], p0l, fh[0], fh[1]); mma16816(acc[1], p0l, fh[2], fh[3]);
            mma16816(acc[0], prh[k], v0l[0], v0l[1]); mma16816(acc[1], prh[k], v0l[2], v0l[3]);
            mma16816(acc[0], p0h, fl[0], fl[1]); mma16816(acc[1], p0h, fl[2], fl[3]);
            mma16816(acc[0], prh[k], v0h[0], v0h[1]); mma16816(acc[1], prh[k], v0h[2], v0h[3]);
            mma16816(acc[0], p0h, fh[0], fh[1]); mma16816(acc[1], p0h, fh[2], fh[3]);
          }
          if (r == rS) {
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
              for (int c2 = 0; c2 < 4; ++c2) accS[n][c2] = acc[n][c2];
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (g + 8 * h < N) {
                float* dst = obase + ((h ? ob1 : ob0) + r * D + qt * 16);
                *reinterpret_cast<float2*>(dst) = make_float2(acc[0][2 * h], acc[0][2 * h + 1]);
                *reinterpret_cast<float2*>(dst + 8) = make_float2(acc[1][2 * h], acc[1][2 * h + 1]);
              }
            }
          }
        }
      }
      *reinterpret_cast<float4*>(red + (warp * 32 + lane) * 8) = make_float4(sx[0][0], sx[0][1], sx[0][2], sx[0][3]);
      *reinterpret_cast<float4*>(red + (warp * 32 + lane) * 8 + 4) = make_float4(sx[1][0], sx[1][1], sx[1][2], sx[1][3]);
    }
    __syncthreads();
    finish_s(AT_HD / 16 - 1);
  }
  if (rflag != nullptr && !(stg.amax <= 65504.f)) atomicOr(rflag, 1u);
  AT_LAP(8)
  AT_PROF_END
#undef SJ
}

template <int NT, int DT>
int launch_at_l0(const float* qkv, float* o, int64_t B, NetDims d, cudaStream_t s) {
  using G = AtGeomC<NT>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attention_jets_tc_l0_kernel<NT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  dim3 grid((unsigned)d.H, (unsigned)B);
  attention_jets_tc_l0_kernel<NT, DT><<<grid, AT_THREADS, G::SMEM, s>>>(qkv, o, d, range_flag_get());
  return (int)cudaGetLastError();
}

template <int NT, int DT>
int launch_at(const float* qkv, float* o, int64_t B, NetDims d, cudaStream_t s) {
  using G = AtGeom<NT>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attention_jets_tc_kernel<NT, false, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  dim3 grid((unsigned)d.H, (unsigned)B);
  attention_jets_tc_kernel<NT, false, DT><<<grid, AT_THREADS, G::SMEM, s>>>(qkv, o, d, range_flag_get());
  return (int)cudaGetLastError();
}


}  // namespace

#ifdef DH_DEBUG_SWITCHES
extern "C" int dh_debug_at_prof(unsigned long long* host16, int reset) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess && host16) e = cudaMemcpyFromSymbol(host16, g_at_prof, 16 * sizeof(unsigned long long));
  if (e == cudaSuccess && reset) { unsigned long long z[16] = {0}; e = cudaMemcpyToSymbol(g_at_prof, z, sizeof(z)); }
  return (int)e;
}
#endif

// The tensor-core form exists for head size 64 and the electron counts of the BASELINE configurations.
bool attention_jets_tc_ok(NetDims d) {
  return d.hd == AT_HD && d.R == 2 * d.N + 8 && (d.D % 4) == 0 && (d.N == 3 || d.N == 6 || d.N == 10 || d.N == 12 || d.N == 16);
}

int attention_jets_tc(const float* qkv, float* o, int64_t B, NetDims d, int layer0, cudaStream_t s) {
  if (!attention_jets_tc_ok(d)) return -2;
  // the default width (D = 256) has its own instantiation with compile-time row strides
#define DH_AT(NT) (d.D == 256 ? (layer0 ? launch_at_l0<NT, 256>(qkv, o, B, d, s) : launch_at<NT, 256>(qkv, o, B, d, s)) \
                              : (layer0 ? launch_at_l0<NT, 0>(qkv, o, B, d, s) : launch_at<NT, 0>(qkv, o, B, d, s)))
  switch (d.N) {
    case 3: return DH_AT(3);
    case 6: return DH_AT(6);
    case 10: return DH_AT(10);
    case 12: return DH_AT(12);
    case 16: return DH_AT(16);
    default: return -2;
  }
#undef DH_AT
}

}  // namespace dh
