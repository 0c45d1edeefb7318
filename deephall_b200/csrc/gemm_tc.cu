// tcgen05 / TMEM / TMA dense contraction with fp32 accuracy from a two-piece operand split (sm_100a).
//
//   C[M, N] = A[M, K] @ W[K, N] (+ bias[N] on rows m with m % rpg == 0)
//
// Every operand x is written x = hi + lo with 11-bit-significand pieces and the product is taken as
// hi*hi + lo*hi + hi*lo on the tensor cores with fp32 accumulation in tensor memory (the dropped
// lo*lo term is 2^-22 relative).  Two instantiations of the same kernel:
//   * kind::tf32 ("3xTF32"):  pieces are TF32 numbers, K step 8 per MMA; 2-stage operand ring of 96 KB stages;
//   * kind::f16  ("3xFP16"):  pieces are fp16 numbers -- the same 11-bit significand, so inside fp16's
//     normal range the three products are bit-identical to the TF32 ones -- K step 16 per MMA at twice
//     the TF32 rate, half the operand bytes per flop, 4-stage operand ring of 48 KB stages.  fp16's narrow exponent
//     range is handled by a per-matrix power-of-two weight scale (undone exactly in the epilogue),
//     saturating conversions, and gradual underflow of `lo` (absolute error <= 2^-25, below fp32
//     rounding of O(1) activations; oracle emulation in DESIGN.md).
// W is given pre-transposed ([Npad][K], K-major) and pre-split (split_weight_tc, once per parameter
// update).  A is split inside the kernel: TMA lands the fp32 tile in shared memory (128B swizzle),
// four warps rewrite it in place as the hi / lo operand tiles, one elected thread issues the MMAs.
//
// PERSISTENT kernel, one CTA per SM, 576 threads, CTA tile 128 x (<=256) x K:
//   warp 0      TMA producer           (192 KB operand ring of 32-K-element stages, runs ahead across tiles)
//   warp 1      TMEM allocator + MMA issuer
//   warps 2..9   splitter              (fp32 -> hi / lo pieces, in shared memory)
//   warps 10..17 epilogue              (tcgen05.ld -> scale, +bias -> swizzled smem staging -> TMA store)
// Launched as clusters of `cl` CTAs (1, 2 or 4) working on different row bands but the same column tile at the
// same time: each CTA fetches 1/cl of the weight tile and TMA-multicasts it to the whole cluster, which cuts the
// L2 -> SM traffic of the re-streamed weights (the measured limiter at cl = 1) by cl.
// The loads and the split of tile t+1 overlap the drain and the stores of tile t.  The two
// accumulators (main, correction) fill all 512 TMEM columns, so the first MMA of tile t+1 waits
// until the epilogue has read tile t out of TMEM (not until its stores have landed).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace dh {

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 32;             // K elements per pipeline stage (both kinds)
constexpr int A_BYTES = BLOCK_M * 128;  // 16 KB: the fp32 A tile TMA lands (128 rows x 32 floats, 128B swizzle)
constexpr int RING_BYTES = 192 * 1024;  // operand ring: tf32 pieces 2 stages x 96 KB, fp16 pieces 4 stages x 48 KB
constexpr int MAX_STAGES = 6;
constexpr int EPI_CHUNK = 32;                            // accumulator columns per epilogue step
constexpr int EPI_BUF_BYTES = 32 * EPI_CHUNK * 4;        // 32 rows x 32 floats = 4 KB (one warp, one step)
constexpr int EPI_WARPS = 8;                             // two per TMEM lane quarter, alternating 32-column chunks
constexpr int EPI_BYTES = EPI_WARPS * EPI_BUF_BYTES;     // one staging buffer per warp = 32 KB
constexpr int SMEM_BYTES = RING_BYTES + EPI_BYTES + 512 /*barriers*/;  // (the window is 1024-byte aligned: no slack)
constexpr int SPLIT_WARPS = 8;
constexpr int EPI_WARP0 = 2 + SPLIT_WARPS;  // first epilogue warp
constexpr int THREADS = 32 * (EPI_WARP0 + EPI_WARPS);  // 576
constexpr int TMEM_COLS = 512;  // [0,256): hi*hi accumulator, [256,512): correction (lo*hi + hi*lo) accumulator

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// wait with back-off: a warp that spins on try_wait issues instructions all the time and takes issue slots from the
// warps that have work (the fused-LayerNorm epilogue shares its schedulers with eight waiting splitter warps)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, bool cluster_scope, unsigned ns) {
  uint32_t ok;
  for (;;) {
    if (cluster_scope)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// same, delivered to the same shared-memory offsets (and signalling the same barrier offset) in every CTA of
// the cluster selected by cta_mask
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// same, but the tile is ADDED to global memory (fp32 reduction performed by the TMA unit / L2)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
template <bool SW64>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  // K-major operand tile whose rows are one swizzle span: SWIZZLE_128B (layout 2, 8-row groups 1024 B apart) or
  // SWIZZLE_64B (layout 4, 8-row groups 512 B apart); SBO = group stride, LBO unused (=1), version 1 (sm_100)
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = ((SW64 ? 512u : 1024u) >> 4) | (1u << 14) | ((SW64 ? 4u : 2u) << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
template <bool F16>
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  // c_format F32 (1) @4, a/b format @7/@10 (kind::tf32: TF32 = 2; kind::f16: F16 = 0), K-major both,
  // N>>3 @17, M>>4 @24
  const uint32_t ab = F16 ? 0u : 2u;
  return (1u << 4) | (ab << 7) | (ab << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool F16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                     uint32_t accumulate) {
  if (F16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same offset in every CTA of cta_mask when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
// ---- CTA-pair (cta_group::2) forms: one MMA spans the two CTAs of a cluster (M = 256, 128 rows each); every CTA
// stages its own A rows and HALF of the weight tile, so a CTA's shared memory serves 8 KB instead of 12 KB of
// operand reads per MMA and receives 32 KB instead of 48 KB per pipeline stage
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta_rank) {  // shared::cluster address of `addr` in CTA cta_rank
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
// (default semantics, release at CTA scope: a cluster-scope release costs ~1000 cycles per arrival and the data it
// would publish -- operand tiles behind a fence.proxy.async, accumulators behind tcgen05 fences -- is read by the
// tensor core / after the tcgen05 fence, not by generic loads of the other CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // acquires what remote arrivals released
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA load into this CTA's shared memory whose completion bytes are counted on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_pair_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// explicit shared-space accesses (pointers derived from the aligned dynamic shared-memory base are generic to the compiler)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
        "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
        "f"(v[30]), "f"(v[31])
      : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// tma_store != 0: C is written through tmC (box 32 x 32 floats, 128B swizzle); otherwise (ldc not a
// multiple of 4 floats, or misaligned C) by direct global stores.
// PAIR (fp16 pieces, merged accumulator): clusters of two CTAs run `tcgen05.mma.cta_group::2` -- see the helpers above.
// ORB (pair form, resident A, 32 jet rows per electron, N K = 12): the contraction is the orbital projection
// (blocks.py:28-35) and its epilogue is the envelope contraction that follows it (blocks.py:59-70) -- the coefficient
// tensor c[rows][2 L N] (the largest activation of the pass) never goes to HBM.  See the epilogue.
struct OrbFuse {
  const float* env;  // per electron: [10 slots][L] complex envelope jets, pre-multiplied by the weight un-scale factor, then
                     // [10][12] complex bias products sum_m bias(m, j) env_s[m]  (envelope_table, tail_kernels.cu)
  float* Mj;         // out: orbital-matrix jets [walkers][32 rows][12 electrons][12 orbitals] complex
  int L;             // orbitals 2Q + 1
  int mpt;           // orbitals per 256-column weight tile (orb_per_tile(L), kernels.h)
};
// LNV (pair form, value-only passes: one row per electron, N = 256): the epilogue is the residual + (tanh) + LayerNorm that
// follows the contraction in the network (psiformer.py:45-48) -- C = LN(res + acc + bias) or LN(res + tanh(acc + bias)),
// in place over res.  Value rows only need their own statistics (no jet couplings), so the fused form is two sweeps over the
// tile's accumulator columns: x = res + f(acc) back into tensor memory with the row sums, then normalise and store.
struct LnFuse {
  const float* res;    // residual = output tensor [M][ldc] (in place)
  const float* gamma;  // LayerNorm scale [256]
  const float* beta;   // LayerNorm bias  [256]
  int tanh_mode;
};
constexpr int ORB_NK = 12;               // orbital columns per (part, m) group
constexpr int ORB_GW = 2 * ORB_NK;       // accumulator columns per m: [re (12) | im (12)]
constexpr int ORB_MPT = BLOCK_N / ORB_GW;  // at most 10 m per column tile; OrbFuse::mpt = orb_per_tile(L) are used
// tensor columns of column tile nt: its orbitals' 24 columns each, rounded up to the 16-column granule of the pair MMA
__device__ __forceinline__ int orb_tile_cols(const OrbFuse& o, int nt) {
  return (ORB_GW * min(o.mpt, o.L - o.mpt * nt) + 15) & ~15;
}

template <bool F16, bool MERGED, bool PAIR, bool ORB, bool LNV>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmC,
               const float* __restrict__ bias, const float* __restrict__ inv_scale_ptr, float* __restrict__ C, int64_t M,
               int N, int K, int64_t ldc, int rpg, int tma_store, int reduce_add, int cl, const float* __restrict__ a_scale_ptr,
               int a_pre, int res, OrbFuse orb, LnFuse ln, unsigned long long* __restrict__ prof, unsigned* __restrict__ rflag) {
  constexpr int merged = MERGED ? 1 : 0;
  constexpr int UMMA_K = F16 ? 16 : 8;    // K elements per MMA (32 bytes of operand row)
  // Stage layout.  tf32 pieces: [A hi 16 KB | A lo 16 KB | B hi 32 KB | B lo 32 KB], rows of 128 B (SWIZZLE_128B).
  //                fp16 pieces: [A hi 8 KB | A lo 8 KB (together = the landed fp32 tile) | B hi 16 KB | B lo 16 KB],
  //                operand rows of 64 B (SWIZZLE_64B).
  //                PAIR:        [A hi 8 KB | A lo 8 KB | B hi 8 KB | B lo 8 KB]: this CTA's 128 of the tile's 256 weight rows.
  static_assert(!PAIR || (F16 && MERGED), "the CTA-pair form exists for fp16 pieces with the merged accumulator");
  static_assert(!ORB || PAIR, "the fused orbital-contraction epilogue exists for the pair form");
  static_assert(!LNV || (PAIR && !ORB), "the fused value LayerNorm epilogue exists for the pair form");
  constexpr int STAGES = PAIR ? 6 : (F16 ? 4 : 2);
  constexpr int ROW_BYTES = F16 ? 64 : 128;            // operand tile row = BLOCK_K pieces
  constexpr int AP_BYTES = BLOCK_M * ROW_BYTES;        // one A piece tile
  constexpr int B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * ROW_BYTES;  // one B piece tile (PAIR: this CTA's half)
  constexpr int STAGE_BYTES = 2 * AP_BYTES + 2 * B_BYTES;
  static_assert(STAGES * STAGE_BYTES == RING_BYTES && 2 * AP_BYTES >= A_BYTES, "operand ring geometry");
  constexpr int A_TX_BYTES = A_BYTES;
  // The dynamic shared-memory window is declared 1024-byte aligned (the operand tiles' swizzle atoms need it), so no slack
  // is allocated for rounding the base up: with 230,912 bytes + the 1 KB the system reserves per block, one 128-thread
  // block without shared memory of ANOTHER launch still fits the SM's 228 KB next to this CTA -- the LayerNorm kernels of the
  // other chunk stream run in the contraction's shadow on the registers and threads it leaves free.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* smem = smem_raw;
  uint8_t* epi_smem = smem + RING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + EPI_BYTES);
  // bars: [0..S) full, [S..2S) split, [2S..3S) empty, [3S] tmem_full, [3S+1] tmem_empty ; then tmem ptr
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto split_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  // merged == 0: main accumulator in TMEM columns [0,256), correction accumulator in [256,512), one tile in
  //              flight (the issuer waits for the drain of the previous tile);
  // merged != 0: all three products of a tile go to ONE accumulator, tiles alternate between columns [0,256)
  //              and [256,512), so tile t+1 is computed while tile t is drained -- at the price of 3x as many
  //              truncating additions into the accumulator ("accumulator modes").
  auto tmem_full_bar = [&](int b) { return bar_base + 8u * (3 * MAX_STAGES + b); };
  auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (3 * MAX_STAGES + 2 + b); };
  auto a_hi = [&](int s) { return smem_base + s * STAGE_BYTES; };
  auto a_lo = [&](int s) { return smem_base + s * STAGE_BYTES + AP_BYTES; };
  auto b_hi = [&](int s) { return smem_base + s * STAGE_BYTES + 2 * AP_BYTES; };
  auto b_lo = [&](int s) { return smem_base + s * STAGE_BYTES + 2 * AP_BYTES + B_BYTES; };
  // res (PAIR, K <= 256, several column tiles per band): the band's A operand tiles stay RESIDENT for all its column
  // tiles -- every (band, k-block) of A lands in one of RES_AR 16 KB regions (a ring over the sequence of k-blocks)
  // and is split in place ONCE per band instead of once per column tile; the weights stream through a separate ring
  // of RES_BST 16 KB stages.  The producer reloads a region for the NEXT band as soon as the last column tile of
  // the current band has released it (non-blocking test between its weight loads), i.e. ~7 k-blocks before the MMA
  // needs it.  Takes the repeated A loads and splits (half of a tile's shared-memory traffic) off every column tile
  // but the first.
  constexpr int RES_AR = 8, RES_BST = 4;
  static_assert((RES_AR + RES_BST) * 16 * 1024 <= RING_BYTES, "resident-A layout");
  auto ra_reg = [&](int r) { return smem_base + r * 2 * AP_BYTES; };
  auto rb_hi = [&](int s) { return smem_base + RES_AR * 2 * AP_BYTES + s * 2 * B_BYTES; };
  auto rb_lo = [&](int s) { return smem_base + RES_AR * 2 * AP_BYTES + s * 2 * B_BYTES + B_BYTES; };
  auto ra_full = [&](int r) { return bar_base + 8u * (24 + r); };
  auto ra_split = [&](int r) { return bar_base + 8u * (24 + RES_AR + r); };
  auto ra_empty = [&](int r) { return bar_base + 8u * (24 + 2 * RES_AR + r); };
  auto rb_full = [&](int s) { return bar_base + 8u * (24 + 3 * RES_AR + s); };
  auto rb_empty = [&](int s) { return bar_base + 8u * (24 + 3 * RES_AR + RES_BST + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;  // a K tail is zero-filled by TMA (A) and zero-padded (W planes)
  const int n_ntiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int64_t n_mtiles = (M + BLOCK_M - 1) / BLOCK_M;
  // Tile order: a CTA takes 128-row bands blockIdx.x, blockIdx.x + gridDim.x, ... and walks all column tiles
  // of a band before moving on, so a band's A rows are fetched from HBM once and re-read from L2.
  // Every CTA of a cluster runs the same number of tiles (the weight multicast is collective): bands past
  // the end of the matrix load zeros and their stores are clipped.
  uint32_t crank = 0;
  if (PAIR || cl > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const uint16_t cmask = (uint16_t)((1u << cl) - 1u);
  const int64_t cbase = (int64_t)blockIdx.x - crank;  // first CTA of this cluster
  // PAIR: the cluster takes PAIRS of adjacent bands (2u, 2u + 1), u = cluster index, + number of clusters, ...
  const int64_t n_units = PAIR ? (n_mtiles + 1) / 2 : n_mtiles;
  const int64_t unit0 = PAIR ? (int64_t)(blockIdx.x >> 1) : cbase;
  const int64_t ustride = PAIR ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  const int64_t my_bands = unit0 < n_units ? (n_units - unit0 + ustride - 1) / ustride : 0;
  const int64_t my_tiles = my_bands * n_ntiles;
  auto band_of = [&](int64_t k) {
    return PAIR ? 2 * (unit0 + (k / n_ntiles) * ustride) + crank : (int64_t)blockIdx.x + (k / n_ntiles) * gridDim.x;
  };
  auto ntile_of = [&](int64_t k) { return (int)(k % n_ntiles); };
  // PAIR: barriers of the leader CTA (rank 0) that the other CTA signals
  auto leader = [&](uint32_t bar) { return mapa_u32(bar, 0); };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
    if (tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), PAIR ? 2 * SPLIT_WARPS : SPLIT_WARPS);  // PAIR: the leader's barrier counts both CTAs' splitters
      mbar_init(empty_bar(s), PAIR ? 1u : (uint32_t)cl);  // one tcgen05.commit arrival from every CTA of the cluster
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), PAIR ? 2 * EPI_WARPS : EPI_WARPS);
    }
    if (PAIR && res) {
      for (int kb = 0; kb < RES_AR; ++kb) { mbar_init(ra_full(kb), 1); mbar_init(ra_split(kb), 2 * SPLIT_WARPS); mbar_init(ra_empty(kb), 1); }
      for (int st = 0; st < RES_BST; ++st) { mbar_init(rb_full(st), 1); mbar_init(rb_empty(st), 1); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {  // collective over the pair: the same columns in both CTAs' tensor memory
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR || cl > 1) cluster_sync_all();  // every CTA's barriers are initialised before any remote arrival
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // developer instrumentation (DH_GEMM_PROF=1): cycles each role spends blocked, summed over CTAs
  unsigned long long pw[3] = {0, 0, 0}, pe[5] = {0, 0, 0, 0, 0};
  const long long t_begin = prof ? clock64() : 0;
  auto timed_wait = [&](uint32_t bar, uint32_t parity, int slot) {
    const long long t0 = prof ? clock64() : 0;
    if (PAIR) mbar_wait_cluster(bar, parity);
    else mbar_wait(bar, parity);
    if (prof) pw[slot] += (unsigned long long)(clock64() - t0);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0 && PAIR && res) {
      uint32_t itb = 0;
      int64_t ja = 0;  // A tiles issued so far, in (band, k-block) order
      const int64_t total_a = my_bands * num_kb;
      auto issue_a = [&](bool blocking) -> bool {
        if (ja >= total_a) return false;
        const int r = (int)(ja % RES_AR);
        const uint32_t par = (uint32_t)(((ja / RES_AR) & 1) ^ 1);
        if (blocking) mbar_wait(ra_empty(r), par); else if (!mbar_test(ra_empty(r), par)) return false;
        const int64_t bi2 = ja / num_kb;
        const int kb2 = (int)(ja % num_kb);
        mbar_arrive_expect_tx(ra_full(r), A_TX_BYTES);
        tma_load_2d(ra_reg(r), &tmA, ra_full(r), kb2 * BLOCK_K, (int)(band_of(bi2 * n_ntiles) * BLOCK_M));
        ++ja;
        return true;
      };
      for (int64_t bi = 0; bi < my_bands; ++bi) {
        for (int nt = 0; nt < n_ntiles; ++nt) {
          const int n0 = nt * BLOCK_N;
          const int n_tile = ORB ? orb_tile_cols(orb, nt) : (N - n0 < BLOCK_N ? ((N - n0 + 15) & ~15) : BLOCK_N);
          const int nb = n0 + (int)crank * (n_tile / 2);
          for (int kb = 0; kb < num_kb; ++kb, ++itb) {
            if (nt == 0) while (ja <= bi * num_kb + kb) issue_a(true);   // this step's A tile at the latest now
            const int st = itb % RES_BST;
            mbar_wait(rb_empty(st), ((itb / RES_BST) & 1) ^ 1);
            if (crank == 0) mbar_arrive_expect_tx(rb_full(st), 4 * B_BYTES);  // both CTAs' halves, hi and lo
            tma_load_2d_pair(rb_hi(st), &tmBhi, leader(rb_full(st)), kb * BLOCK_K, nb);
            tma_load_2d_pair(rb_lo(st), &tmBlo, leader(rb_full(st)), kb * BLOCK_K, nb);
            // one tile of the next band per step, as regions come free, queued BEHIND the latency-critical weight
            // loads (a burst of HBM-bound A loads in front of them starves the MMA)
            if (ja < (bi + 2) * num_kb) issue_a(false);
          }
        }
      }
    } else if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tk = 0; tk < my_tiles; ++tk) {
        const int m0 = (int)(band_of(tk) * BLOCK_M);
        const int n0 = ntile_of(tk) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          timed_wait(empty_bar(s), ph ^ 1, 0);
          if (PAIR) {
            // A: this CTA's band, counted on its own barrier (its splitter waits there).  Weights: this CTA's half
            // of the tile's rows, counted on the LEADER's barrier together with the leader's own loads.
            const int n_tile = N - n0 < BLOCK_N ? ((N - n0 + 15) & ~15) : BLOCK_N;
            const int nb = n0 + (int)crank * (n_tile / 2);
            mbar_arrive_expect_tx(full_bar(s), crank == 0 ? A_TX_BYTES + 4 * B_BYTES : A_TX_BYTES);
            tma_load_2d(a_hi(s), &tmA, full_bar(s), kb * BLOCK_K, m0);
            tma_load_2d_pair(b_hi(s), &tmBhi, leader(full_bar(s)), kb * BLOCK_K, nb);
            tma_load_2d_pair(b_lo(s), &tmBlo, leader(full_bar(s)), kb * BLOCK_K, nb);
            continue;
          }
          if (a_pre) {
            // A arrives already split (fp16 hi / lo planes written by the producing kernel): operand tiles by TMA
            mbar_arrive_expect_tx(full_bar(s), 2 * AP_BYTES + 2 * B_BYTES);
            tma_load_2d(a_hi(s), &tmA, full_bar(s), kb * BLOCK_K, m0);
            tma_load_2d(a_lo(s), &tmAlo, full_bar(s), kb * BLOCK_K, m0);
          } else {
            mbar_arrive_expect_tx(full_bar(s), A_TX_BYTES + 2 * B_BYTES);
            tma_load_2d(a_hi(s), &tmA, full_bar(s), kb * BLOCK_K, m0);
          }
          if (cl == 1) {
            tma_load_2d(b_hi(s), &tmBhi, full_bar(s), kb * BLOCK_K, n0);
            tma_load_2d(b_lo(s), &tmBlo, full_bar(s), kb * BLOCK_K, n0);
          } else {
            // rows [crank * 256/cl, (crank+1) * 256/cl) of the weight tile, delivered to every CTA of the cluster
            const int brows = BLOCK_N / cl;
            tma_load_2d_mc(b_hi(s) + crank * brows * ROW_BYTES, &tmBhi, full_bar(s), kb * BLOCK_K, n0 + (int)crank * brows, cmask);
            tma_load_2d_mc(b_lo(s) + crank * brows * ROW_BYTES, &tmBlo, full_bar(s), kb * BLOCK_K, n0 + (int)crank * brows, cmask);
          }
        }
      }
      if (prof) { atomicAdd(prof + 0, pw[0]); atomicAdd(prof + 10, (unsigned long long)(clock64() - t_begin)); }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (PAIR: the leader CTA's, for both)
    if (lane == 0 && PAIR && res && crank == 0) {
      uint32_t itb = 0, tl = 0;
      for (int64_t bi = 0; bi < my_bands; ++bi) {
        for (int nt = 0; nt < n_ntiles; ++nt, ++tl) {
          const int n0 = nt * BLOCK_N;
          const int n_tile = ORB ? orb_tile_cols(orb, nt) : (N - n0 < BLOCK_N ? ((N - n0 + 15) & ~15) : BLOCK_N);
          const uint32_t idesc = make_idesc<F16>(2 * BLOCK_M, n_tile);
          const int ab = (int)(tl & 1);
          const uint32_t acc = tmem_base + (uint32_t)ab * BLOCK_N;
          timed_wait(tmem_empty_bar(ab), ((tl >> 1) & 1) ^ 1, 0);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int kb = 0; kb < num_kb; ++kb, ++itb) {
            const uint32_t ja = (uint32_t)(bi * num_kb + kb);  // position of this (band, k-block) in the A ring
            const int r = ja % RES_AR;
            if (nt == 0) timed_wait(ra_split(r), (ja / RES_AR) & 1, 2);  // both CTAs' A tiles of this k-block landed and split
            const int st = itb % RES_BST;
            timed_wait(rb_full(st), (itb / RES_BST) & 1, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t dah = make_smem_desc<F16>(ra_reg(r)), dal = make_smem_desc<F16>(ra_reg(r) + AP_BYTES);
            const uint64_t dbh = make_smem_desc<F16>(rb_hi(st)), dbl = make_smem_desc<F16>(rb_lo(st));
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);
              umma_pair_f16(acc, dal + adv, dbh + adv, idesc, (kb | k) != 0);
              umma_pair_f16(acc, dah + adv, dbl + adv, idesc, 1);
              umma_pair_f16(acc, dah + adv, dbh + adv, idesc, 1);
            }
            umma_commit_pair(rb_empty(st), 3);
            if (nt == n_ntiles - 1) umma_commit_pair(ra_empty(r), 3);  // the band is done with this k-block of A
          }
          umma_commit_pair(tmem_full_bar(ab), 3);
        }
      }
      if (prof) { atomicAdd(prof + 1, pw[0]); atomicAdd(prof + 2, pw[1]); atomicAdd(prof + 3, pw[2]);
                  atomicAdd(prof + 11, (unsigned long long)(clock64() - t_begin)); }
    } else if (lane == 0 && (!PAIR || crank == 0)) {
      uint32_t it = 0, tl = 0;
      for (int64_t tk = 0; tk < my_tiles; ++tk, ++tl) {
        const int n0 = ntile_of(tk) * BLOCK_N;
        int n_tile = N - n0 < BLOCK_N ? N - n0 : BLOCK_N;
        n_tile = (n_tile + 15) & ~15;  // Wt buffers are zero-padded to a multiple of 16 rows
        const uint32_t idesc = make_idesc<F16>(PAIR ? 2 * BLOCK_M : BLOCK_M, n_tile);
        // the epilogue has read the previous user of this accumulator out of TMEM
        const int ab = merged ? (int)(tl & 1) : 0;
        const uint32_t au = merged ? (tl >> 1) : tl;  // how many times this accumulator has been used before
        const uint32_t acc = tmem_base + (uint32_t)ab * BLOCK_N;
        timed_wait(tmem_empty_bar(ab), (au & 1) ^ 1, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          timed_wait(full_bar(s), ph, 1);
          if (!a_pre) timed_wait(split_bar(s), ph, 2);  // PAIR: both CTAs' A tiles have landed and are split
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t dah = make_smem_desc<F16>(a_hi(s)), dal = make_smem_desc<F16>(a_lo(s));
          const uint64_t dbh = make_smem_desc<F16>(b_hi(s)), dbl = make_smem_desc<F16>(b_lo(s));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);  // 32 operand bytes per MMA; start-address field is in 16-byte units
            // The tensor core rounds toward zero when it adds into the fp32 accumulator, a bias that
            // grows with the number of additions: keep the small correction terms in their own
            // accumulator so the main one sees K/8 additions instead of 3K/8.
            if (PAIR) {
              umma_pair_f16(acc, dal + adv, dbh + adv, idesc, (kb | k) != 0);
              umma_pair_f16(acc, dah + adv, dbl + adv, idesc, 1);
              umma_pair_f16(acc, dah + adv, dbh + adv, idesc, 1);
            } else if (merged) {
              umma<F16>(acc, dal + adv, dbh + adv, idesc, (kb | k) != 0);
              umma<F16>(acc, dah + adv, dbl + adv, idesc, 1);
              umma<F16>(acc, dah + adv, dbh + adv, idesc, 1);
            } else {
              umma<F16>(tmem_base + BLOCK_N, dal + adv, dbh + adv, idesc, (kb | k) != 0);
              umma<F16>(tmem_base + BLOCK_N, dah + adv, dbl + adv, idesc, 1);
              umma<F16>(tmem_base, dah + adv, dbh + adv, idesc, (kb | k) != 0);
            }
          }
          // arrives (in every CTA of the cluster: their TMA writes into this stage too) when the MMAs reading
          // this stage have completed
          if (PAIR) umma_commit_pair(empty_bar(s), 3);
          else if (cl == 1) umma_commit(empty_bar(s)); else umma_commit_mc(empty_bar(s), cmask);
        }
        if (PAIR) umma_commit_pair(tmem_full_bar(ab), 3); else umma_commit(tmem_full_bar(ab));
      }
      if (prof) { atomicAdd(prof + 1, pw[0]); atomicAdd(prof + 2, pw[1]); atomicAdd(prof + 3, pw[2]);
                  atomicAdd(prof + 11, (unsigned long long)(clock64() - t_begin)); }
    }
  } else if (warp < EPI_WARP0) {
    // ------------------------------------------------------------------ splitter (8 warps)
    const int t = threadIdx.x - 64;  // 0..255
    // optional power-of-two scale of A (reverse pass: gradients are O(1/B) and would underflow fp16 pieces);
    // a_scale_ptr = {scale, 1/scale}, the epilogue undoes it
    const float asc = a_scale_ptr ? __ldg(a_scale_ptr) : 1.f;
    float amax = 0.f;  // largest |a| this thread split: an fp16 piece saturates beyond 65504 (reported through rflag)
    uint32_t it = 0;
    const bool rs = PAIR && res;  // resident A: one split per (band, k-block), in region kb
    for (int64_t tk = 0; tk < (a_pre ? 0 : (rs ? my_bands : my_tiles)); ++tk) {
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const int rr = it % RES_AR;  // rs: `it` counts the (band, k-block) pairs = position in the A ring
        const uint32_t ph = rs ? (it / RES_AR) & 1 : (it / STAGES) & 1;
        const uint32_t wbar = rs ? ra_full(rr) : full_bar(s);
        uint8_t* const stage_reg = rs ? smem + rr * 2 * AP_BYTES : smem + s * STAGE_BYTES;
        if (t == 0) timed_wait(wbar, ph, 0); else mbar_wait(wbar, ph);
        const long long ts0 = (prof && t == 0) ? clock64() : 0;
        if (F16) {
          // Two threads per tile row r: thread (r, h) converts the 16 floats k = 16h .. 16h+15 of the landed fp32
          // tile (row r = 128 B at r*128, 16-byte piece j at position j ^ (r & 7)).  The fp16 hi tile (rows of
          // 64 B at r*64, piece c at position c ^ ((r >> 1) & 3)) overwrites the first 8 KB of the same region and
          // the lo tile the second 8 KB, so all 256 splitter threads read before any of them writes.
          const int r = t >> 1, h = t & 1, sw = r & 7;
          uint8_t* reg = stage_reg;
          const uint8_t* src = reg + r * 128;
          float4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j] = *reinterpret_cast<const float4*>(src + (((4 * h + j) ^ sw) << 4));
            v[j].x *= asc; v[j].y *= asc; v[j].z *= asc; v[j].w *= asc;
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * SPLIT_WARPS) : "memory");
          const int sw2 = (r >> 1) & 3;
#pragma unroll
          for (int c = 0; c < 2; ++c) {  // piece 2h + c of the fp16 row = floats 8c .. 8c+7 of this thread
            const float4 a = v[2 * c], b = v[2 * c + 1];
            uint4 hi, lo;
            split_f16x2(a.x, a.y, hi.x, lo.x);
            split_f16x2(a.z, a.w, hi.y, lo.y);
            split_f16x2(b.x, b.y, hi.z, lo.z);
            split_f16x2(b.z, b.w, hi.w, lo.w);
            const int pos = r * 64 + (((2 * h + c) ^ sw2) << 4);
            *reinterpret_cast<uint4*>(reg + pos) = hi;
            *reinterpret_cast<uint4*>(reg + AP_BYTES + pos) = lo;
          }
        } else {
          float4* hi = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
          float4* lo = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + AP_BYTES);
#pragma unroll
          for (int i = 0; i < A_BYTES / 16 / (32 * SPLIT_WARPS); ++i) {
            const int idx = t + i * 32 * SPLIT_WARPS;
            float4 v = hi[idx], h, l;
            v.x *= asc; v.y *= asc; v.z *= asc; v.w *= asc;
            h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
            l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
            hi[idx] = h;
            lo[idx] = l;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(leader(rs ? ra_split(rr) : split_bar(s))); else mbar_arrive(split_bar(s)); }
        if (prof && t == 0) pw[1] += (unsigned long long)(clock64() - ts0);
      }
    }
    if (prof && t == 0) { atomicAdd(prof + 4, pw[0]); atomicAdd(prof + 5, pw[1]); }
    // fp16 pieces have a narrower range than the reference's fp32: a saturated piece is reported, never silent
    if (F16 && rflag != nullptr && !(amax <= 65504.f)) atomicOr(rflag, 1u);
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int chalf = (warp - EPI_WARP0) >> 2;    // this warp takes the 32-column chunks with index % 2 == chalf
    const int row = q * 32 + lane;
    const bool lead = warp == EPI_WARP0 && lane == 0;
    uint8_t* buf = epi_smem + (warp - EPI_WARP0) * EPI_BUF_BYTES;
    uint32_t tl = 0;
    // undoes the fp16 weight scale and the optional A scale (powers of two)
    const float inv_scale = (inv_scale_ptr ? __ldg(inv_scale_ptr) : 1.f) * (a_scale_ptr ? __ldg(a_scale_ptr + 1) : 1.f);
    float oacc[ORB ? ORB_GW : 1], oacc2[ORB ? ORB_GW : 1];  // ORB: this row's 12 complex outputs, carried over a band's column tiles
    (void)oacc; (void)oacc2;
    for (int64_t tk = 0; tk < my_tiles; ++tk, ++tl) {
      const int64_t m0 = band_of(tk) * BLOCK_M;
      const int n0 = ntile_of(tk) * BLOCK_N;
      int n_tile = N - n0 < BLOCK_N ? N - n0 : BLOCK_N;
      n_tile = (n_tile + 15) & ~15;
      const int64_t m = m0 + row;
      const bool add_bias = bias != nullptr && (rpg <= 1 || (m % rpg) == 0);
      // first row of this warp's 32-row slice that is the value row of an electron (jet passes, rpg > 1)
      const int first_vrow = rpg > 1 ? (int)((rpg - (m0 + q * 32) % rpg) % rpg) : 0;
      const int ab = merged ? (int)(tl & 1) : 0;
      const uint32_t au = merged ? (tl >> 1) : tl;
      if (lead) timed_wait(tmem_full_bar(ab), au & 1, 0); else mbar_wait(tmem_full_bar(ab), au & 1);
      const long long te0 = (prof && lead) ? clock64() : 0;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ab * BLOCK_N;
      if constexpr (ORB) {
        if (rpg <= 1) {
          // -------------------------------------------------------------- value-only passes: one row per electron
          // lane = one electron; its envelope (L complex numbers, pre-multiplied by the weight un-scale factor) and its bias
          // products (12 complex) come from the per-electron table; the two warps of a lane quarter take the even / odd m.
          const int nt = ntile_of(tk);
          const int L = orb.L;
          const bool live = m < M;
          const float* et = orb.env + (live ? m : 0) * (int64_t)(2 * (L + ORB_NK));
          float* xsum = reinterpret_cast<float*>(epi_smem + q * (2 * EPI_BUF_BYTES));  // [24][32] partial sums of the odd warp
          if (nt == 0) {
            asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");  // the previous band's partial sums have been read
#pragma unroll
            for (int j = 0; j < ORB_GW; ++j) oacc[j] = 0.f;
          }
          const int mt = min(orb.mpt, L - orb.mpt * nt);
#pragma unroll 1
          for (int ml = chalf; ml < mt; ml += 2) {
            uint32_t v[24];
            tmem_ld16(tbase + (uint32_t)(ORB_GW * ml), *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
            tmem_ld8(tbase + (uint32_t)(ORB_GW * ml + 16), *reinterpret_cast<uint32_t(*)[8]>(&v[16]));
            const float2 e0 = __ldg(reinterpret_cast<const float2*>(et + 2 * (orb.mpt * nt + ml)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ml + 2 >= mt) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(leader(tmem_empty_bar(ab)));
            }
#pragma unroll
            for (int j = 0; j < ORB_NK; ++j) {
              const float cr = __uint_as_float(v[j]), ci = __uint_as_float(v[ORB_NK + j]);
              oacc[j] = fmaf(cr, e0.x, fmaf(-ci, e0.y, oacc[j]));
              oacc[ORB_NK + j] = fmaf(cr, e0.y, fmaf(ci, e0.x, oacc[ORB_NK + j]));
            }
          }
          if (mt <= chalf) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader(tmem_empty_bar(ab)));
          }
          if (nt == n_ntiles - 1) {
            if (chalf == 1) {
#pragma unroll
              for (int j = 0; j < ORB_GW; ++j) xsum[j * 32 + lane] = oacc[j];
            }
            asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
            if (chalf == 0 && live) {
              float* dst = orb.Mj + m * (int64_t)(2 * ORB_NK);
#pragma unroll
              for (int j = 0; j < ORB_NK; j += 2) {
                const float4 bp = __ldg(reinterpret_cast<const float4*>(et + 2 * L + 2 * j));  // bias products of columns j, j + 1
                *reinterpret_cast<float4*>(dst + 2 * j) =
                    make_float4(oacc[j] + xsum[j * 32 + lane] + bp.x, oacc[ORB_NK + j] + xsum[(ORB_NK + j) * 32 + lane] + bp.y,
                                oacc[j + 1] + xsum[(j + 1) * 32 + lane] + bp.z, oacc[ORB_NK + j + 1] + xsum[(ORB_NK + j + 1) * 32 + lane] + bp.w);
              }
            }
          }
          continue;
        }
        // ---------------------------------------------------------------- fused envelope contraction (orbital matrices)
        // Accumulator columns of tile nt: m = 10 nt + ml, ml < 10, at [24 ml, 24 ml + 24) = [re c(m, 0..11) | im c(m, 0..11)]
        // (the weights were laid out so by orb_permute_weights).  The 32 lanes of this warp are the 32 jet rows of ONE
        // electron (value | 24 J | S | 3 D | 3 T); the two warps of a lane quarter take the even / odd ml.
        //   out_r[j] = sum_m c_r[m, j] env_0[m]                                  every row
        //            + sum_m c_0[m, j] env_s[m]   s = slot of row r               own-flow J rows, S, D_a, T_a  (product rule)
        //   S  += 2 sum_e (c_J(2i+e) . env_(1+e)),  T_a += 2 (c_D_a . env_(4+a))  second accumulator of rows J(2i+e), D_a,
        //                                                                        handed to S / T_a at the end of the band
        // env_t[m]: envelope jets of THIS electron (a per-electron table in global memory, staged in shared memory per band).
        const int nt = ntile_of(tk);
        const int64_t ge = (m0 >> 5) + q;            // global electron index (M % 128 == 0: bands hold whole electrons)
        const int ei = (int)(ge % 12);               // electron within its walker
        const int L = orb.L;
        // this warp pair's scratch: env table [10][L] complex (<= 10 * 48 * 8 B) | partial sums of the odd warp [32][48]
        float* envs = reinterpret_cast<float*>(epi_smem + q * (2 * EPI_BUF_BYTES));   // 8 KB per quarter
        float* xsum = envs;  // [48][32] partial sums of the odd warp: reuses the table's space once the band's groups are read
        static_assert(2 * EPI_BUF_BYTES >= 32 * 2 * ORB_GW * 4 && 2 * EPI_BUF_BYTES >= 10 * 96 * 4, "per-quarter scratch");
        const int r = lane;
        // slot whose envelope jets multiply the VALUE row into this row (0 = none), and the slot of the second accumulator
        int s1 = 0, s2 = 0;
        if (r == 1 + 2 * ei) { s1 = 1; s2 = 1; }
        else if (r == 2 + 2 * ei) { s1 = 2; s2 = 2; }
        else if (r == 25) s1 = 3;
        else if (r >= 26 && r <= 28) { s1 = 4 + (r - 26); s2 = s1; }
        else if (r >= 29) s1 = 7 + (r - 29);
        if (nt == 0) {
          // stage this electron's envelope table; both warps of the quarter fill and read it
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");  // the previous band's reads are done
          const float* eg = orb.env + ge * (int64_t)(10 * 2 * L + 10 * 2 * ORB_NK);
          const bool live = m0 + q * 32 < M;
          for (int t = chalf * 32 + lane; t < 10 * 2 * L; t += 64) {
            const int sl = t / (2 * L), rest = t - sl * 2 * L;
            envs[sl * 96 + rest] = live ? __ldg(eg + t) : 0.f;
          }
          for (int t = chalf * 32 + lane; t < 10 * 2 * ORB_NK; t += 64) envs[1536 + t] = live ? __ldg(eg + 10 * 2 * L + t) : 0.f;
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
#pragma unroll
          for (int j = 0; j < ORB_GW; ++j) { oacc[j] = 0.f; oacc2[j] = 0.f; }
        }
        const int mt = min(orb.mpt, L - orb.mpt * nt);  // m's in this tile
#pragma unroll 1
        for (int ml = chalf; ml < mt; ml += 2) {
          const int mm = orb.mpt * nt + ml;
          uint32_t v[24];
          tmem_ld16(tbase + (uint32_t)(ORB_GW * ml), *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          tmem_ld8(tbase + (uint32_t)(ORB_GW * ml + 16), *reinterpret_cast<uint32_t(*)[8]>(&v[16]));
          // (the table carries the weight un-scale factor; rows without a product-rule term read slot 0 with weight 0)
          const float2 e0 = *reinterpret_cast<const float2*>(envs + 2 * mm);
          float2 e1 = *reinterpret_cast<const float2*>(envs + s1 * 96 + 2 * mm);
          float2 e2 = *reinterpret_cast<const float2*>(envs + s2 * 96 + 2 * mm);
          if (!s1) e1 = make_float2(0.f, 0.f);
          if (!s2) e2 = make_float2(0.f, 0.f);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (ml + 2 >= mt) {  // this warp's last read of the tile's accumulators: hand TMEM back to the MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader(tmem_empty_bar(ab)));
          }
#pragma unroll
          for (int j = 0; j < ORB_NK; ++j) {
            const float cr = __uint_as_float(v[j]), ci = __uint_as_float(v[ORB_NK + j]);
            const float c0r = __shfl_sync(0xffffffffu, cr, 0), c0i = __shfl_sync(0xffffffffu, ci, 0);
            // own row x env_0
            oacc[j] = fmaf(cr, e0.x, fmaf(-ci, e0.y, oacc[j]));
            oacc[ORB_NK + j] = fmaf(cr, e0.y, fmaf(ci, e0.x, oacc[ORB_NK + j]));
            // value row x env_s1 (product rule)
            oacc[j] = fmaf(c0r, e1.x, fmaf(-c0i, e1.y, oacc[j]));
            oacc[ORB_NK + j] = fmaf(c0r, e1.y, fmaf(c0i, e1.x, oacc[ORB_NK + j]));
            // own row x env_s2 (second accumulator: goes to S / T_a doubled)
            oacc2[j] = fmaf(cr, e2.x, fmaf(-ci, e2.y, oacc2[j]));
            oacc2[ORB_NK + j] = fmaf(cr, e2.y, fmaf(ci, e2.x, oacc2[ORB_NK + j]));
          }
        }
        if (mt <= chalf) {  // a narrow last tile can leave the odd warp without a group
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader(tmem_empty_bar(ab)));
        }
        if (nt == n_ntiles - 1) {
          // the band is complete: the odd warp hands its partial sums to the even one, which assembles and stores
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");  // both warps are done with the envelope table
          if (chalf == 1) {
#pragma unroll
            for (int j = 0; j < ORB_GW; ++j) { xsum[j * 32 + lane] = oacc[j]; xsum[(ORB_GW + j) * 32 + lane] = oacc2[j]; }
          }
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
          if (chalf == 0) {
#pragma unroll
            for (int j = 0; j < ORB_GW; ++j) { oacc[j] += xsum[j * 32 + lane]; oacc2[j] += xsum[(ORB_GW + j) * 32 + lane]; }
            // S += 2 (acc2 of J(2i) + acc2 of J(2i+1)); T_a += 2 acc2 of D_a
            const int srcA = r == 25 ? 1 + 2 * ei : (r >= 29 ? r - 3 : 0), srcB = 2 + 2 * ei;
            const float wA = (r == 25 || r >= 29) ? 2.f : 0.f, wB = r == 25 ? 2.f : 0.f;
#pragma unroll
            for (int j = 0; j < ORB_GW; ++j) {
              const float a = __shfl_sync(0xffffffffu, oacc2[j], srcA), b2 = __shfl_sync(0xffffffffu, oacc2[j], srcB);
              oacc[j] = fmaf(wA, a, fmaf(wB, b2, oacc[j]));
            }
            // the bias sits on the value row only: its products with the envelope slots come ready-made from the table
            {
              const float* bp0 = envs + 1536;              // slot 0: the value row's own output
              const float* bp1 = envs + 1536 + s1 * 2 * ORB_NK;  // slot s1: the product-rule term of this row
              const float u0 = r == 0 ? 1.f : 0.f, u1 = s1 ? 1.f : 0.f;
#pragma unroll
              for (int j = 0; j < ORB_NK; ++j) {
                oacc[j] = fmaf(u0, bp0[2 * j], fmaf(u1, bp1[2 * j], oacc[j]));
                oacc[ORB_NK + j] = fmaf(u0, bp0[2 * j + 1], fmaf(u1, bp1[2 * j + 1], oacc[ORB_NK + j]));
              }
            }
            if (m < M) {
              const int64_t wb = ge / 12;
              float* dst = orb.Mj + (((wb * 32 + r) * 12 + ei) * 12) * 2;
#pragma unroll
              for (int j = 0; j < ORB_NK; j += 2)
                *reinterpret_cast<float4*>(dst + 2 * j) = make_float4(oacc[j], oacc[ORB_NK + j], oacc[j + 1], oacc[ORB_NK + j + 1]);
            }
          }
        }
        continue;
      }
      if constexpr (LNV) {
        // ---------------------------------------------------------------- fused residual + (tanh) + LayerNorm, value rows
        // lane = one row (one electron's value row); the two warps of a lane quarter take the even / odd 32-column chunks.
        const bool mvalid = m < M;
        const float* rrow = ln.res + (mvalid ? m : 0) * ldc;
        float s_x = 0.f, s_xx = 0.f;
#pragma unroll 1
        for (int c0 = chalf * EPI_CHUNK; c0 < BLOCK_N; c0 += 2 * EPI_CHUNK) {
          uint32_t v[32];
          tmem_ld32(tbase + (uint32_t)c0, v);
          float x[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a4 = mvalid ? __ldg(reinterpret_cast<const float4*>(rrow + c0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            x[4 * j] = a4.x; x[4 * j + 1] = a4.y; x[4 * j + 2] = a4.z; x[4 * j + 3] = a4.w;
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float f = fmaf(__uint_as_float(v[4 * j + u]), inv_scale, bb[u]);
              if (ln.tanh_mode) {
                // tanh = 1 - 2 / (1 + e^{2|x|}) with the fast exponential and division: absolute error ~1e-7, the level of the
                // contraction's own rounding, at a quarter of tanhf's instructions (32 per lane and chunk: the epilogue's main cost)
                const float e2 = __expf(2.f * fabsf(f));
                f = copysignf(1.f - __fdividef(2.f, e2 + 1.f), f);
              }
              const float xv = x[4 * j + u] + f;
              x[4 * j + u] = xv;
              s_x += xv;
              s_xx = fmaf(xv, xv, s_xx);
            }
          }
          tmem_st32(tbase + (uint32_t)c0, x);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // the two warps of the quarter exchange their halves of the row sums through their staging buffers
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous tile's stores have read them
        __syncwarp();
        reinterpret_cast<float2*>(buf)[lane] = make_float2(s_x, s_xx);
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        {
          const uint8_t* other = epi_smem + ((warp - EPI_WARP0) ^ 4) * EPI_BUF_BYTES;
          const float2 o2 = reinterpret_cast<const float2*>(other)[lane];
          s_x += o2.x;
          s_xx += o2.y;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");  // both have read before the buffers take the outputs
        constexpr float invD = 1.f / 256.f;
        const float mu = s_x * invD;
        const float rho = rsqrtf(fmaxf(s_xx * invD - mu * mu, 0.f) + 1e-5f);  // flax LayerNorm: fast variance, eps 1e-5
#pragma unroll 1
        for (int c0 = chalf * EPI_CHUNK; c0 < BLOCK_N; c0 += 2 * EPI_CHUNK) {
          uint32_t v[32];
          tmem_ld32(tbase + (uint32_t)c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 2 * EPI_CHUNK >= BLOCK_N) {  // last read of this accumulator
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader(tmem_empty_bar(ab)));
          }
          float y[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(ln.gamma + c0 + 4 * j));
            const float4 e4 = __ldg(reinterpret_cast<const float4*>(ln.beta + c0 + 4 * j));
            y[4 * j] = fmaf((__uint_as_float(v[4 * j]) - mu) * rho, g4.x, e4.x);
            y[4 * j + 1] = fmaf((__uint_as_float(v[4 * j + 1]) - mu) * rho, g4.y, e4.y);
            y[4 * j + 2] = fmaf((__uint_as_float(v[4 * j + 2]) - mu) * rho, g4.z, e4.z);
            y[4 * j + 3] = fmaf((__uint_as_float(v[4 * j + 3]) - mu) * rho, g4.w, e4.w);
          }
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, smem_u32(buf), n0 + c0, (int)(m0 + q * 32));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        continue;
      }
      bool released = false;
      for (int c0 = chalf * EPI_CHUNK; c0 < n_tile; c0 += 2 * EPI_CHUNK) {
        uint32_t v[32];
        long long tp0 = (prof && lead) ? clock64() : 0, tp1;
#define DH_LAP(slot) if (prof && lead) { tp1 = clock64(); pe[slot] += (unsigned long long)(tp1 - tp0); tp0 = tp1; }
        float o[32];
        tmem_ld32(tbase + (uint32_t)c0, v);
        // lane j fetches bias column j of this step (coalesced)
        const float bl = (bias != nullptr && n0 + c0 + lane < N) ? __ldg(bias + n0 + c0 + lane) : 0.f;
        if (!MERGED) {
          uint32_t w[32];
          tmem_ld32(tbase + (uint32_t)c0 + BLOCK_N, w);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = (__uint_as_float(v[j]) + __uint_as_float(w[j])) * inv_scale;
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) * inv_scale;
        }
        DH_LAP(0)
        if (c0 + 2 * EPI_CHUNK >= n_tile) {
          // this warp's last read of the tile's accumulators: hand TMEM back to the MMA issuer
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_cluster(leader(tmem_empty_bar(ab))); else mbar_arrive(tmem_empty_bar(ab)); }
          released = true;
          if (prof && lead) pw[1] += (unsigned long long)(clock64() - te0);
        }
        if (bias != nullptr && (rpg <= 1 || !tma_store)) {
          // every row takes the bias (value-only passes), or no staging buffer to patch: per-lane adds
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float bj = __shfl_sync(0xffffffffu, bl, j);
            if (add_bias) o[j] += bj;
          }
        }
        DH_LAP(1)
        if (tma_store) {
          // the staging buffer of this warp: wait until the previous TMA store has read it
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          DH_LAP(2)
          // row `lane` of the 32 x 32 chunk; 16-byte piece j lives at piece position j ^ (lane & 7) (128B swizzle)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          if (bias != nullptr && rpg > 1) {
            // jet passes: only every rpg-th row (the value row of an electron) takes the bias; patch those
            // rows in the staging buffer, lane j adding bias column j
            __syncwarp();
            for (int r = first_vrow; r < 32; r += rpg) {
              float* pz = reinterpret_cast<float*>(buf + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
              *pz += bl;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          DH_LAP(3)
          if (lane == 0) {
            if (reduce_add) tma_reduce_add_2d(&tmC, smem_u32(buf), n0 + c0, (int)(m0 + q * 32));
            else tma_store_2d(&tmC, smem_u32(buf), n0 + c0, (int)(m0 + q * 32));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          DH_LAP(4)
        } else if (m < M) {
          float* crow = C + m * ldc + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < N) {
              if (reduce_add) atomicAdd(crow + j, o[j]); else crow[j] = o[j];
            }
        }
      }
      if (!released) {  // a narrow tile left this warp without a chunk
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(leader(tmem_empty_bar(ab))); else mbar_arrive(tmem_empty_bar(ab)); }
      }
    }
    if (tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (prof && lead) { for (int i = 0; i < 5; ++i) atomicAdd(prof + 16 + i, pe[i]);
                        atomicAdd(prof + 6, pw[0]); atomicAdd(prof + 7, pw[1]);
                        atomicAdd(prof + 12, (unsigned long long)(clock64() - t_begin)); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR || cl > 1) cluster_sync_all();  // no CTA leaves while a peer can still write into its shared memory
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// =============================================================================================================
// "TN" contraction for the reverse pass and the KFAC factor sums:   C[Ma, Nb] += A[rows, Ma]^T  B[rows, Nb]
// (dW = X^T dY, Gram matrices X^T X and G^T G): the contracted index is the ROW index of both row-major operands, so
// neither is K-major in memory.  One CTA owns one 256 x 256 block of C and a slice of the rows (split-K over the
// grid); per 32-row step TMA lands the two fp32 tiles [32][256] as they lie, the 256 splitter threads -- thread t
// = column t -- read their column (conflict-free), split it into fp16 hi / lo pieces and write ROW t of the K-major,
// 64B-swizzled operand tiles (the transposition happens in this register pass), and one elected thread issues
// 2 M-tiles x 2 K-steps x 3 products of tcgen05.mma into two 128 x 256 fp32 accumulators (all 512 TMEM columns).
// After its last step the CTA adds its block into C with TMA reduce-add.  Optional power-of-two scales keep
// small-magnitude operands (cotangent-sized gradients) inside fp16's range; the epilogue divides them out.
// =============================================================================================================
constexpr int TN_BK = 32;                        // rows (contracted index) per step
constexpr int TN_LAND_BYTES = TN_BK * 256 * 4;    // one landed fp32 tile [32][256]
constexpr int TN_OP_BYTES = 256 * 64;             // one fp16 operand piece tile: 256 rows of 64 B
constexpr int TN_SMEM_BYTES = 4 * TN_LAND_BYTES + 4 * TN_OP_BYTES + EPI_BYTES + 1024 + 256;

__global__ void __launch_bounds__(THREADS, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, int64_t rows, int nblk_n, int ksplit,
                  const float* __restrict__ a_scale_ptr, const float* __restrict__ b_scale_ptr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // [landing: 2 stages x (A 32 KB | B 32 KB)] [operands: A hi | A lo | B hi | B lo, 16 KB each] [epilogue staging] [barriers]
  uint8_t* land = smem;
  uint8_t* ops = smem + 4 * TN_LAND_BYTES;
  uint8_t* epi_smem = ops + 4 * TN_OP_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + EPI_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };         // TMA landed stage s
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 + s); };  // splitters have read stage s
  const uint32_t op_full = bar_base + 8u * 4;                        // operand tiles written
  const uint32_t op_empty = bar_base + 8u * 5;                       // MMAs have read the operand tiles
  const uint32_t acc_done = bar_base + 8u * 6;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item: block (bm, bn) of C and k-slice ks
  const int blk = blockIdx.x / ksplit, ks = blockIdx.x % ksplit;
  const int bm = blk / nblk_n, bn = blk % nblk_n;
  const int64_t nkb = (rows + TN_BK - 1) / TN_BK;
  const int64_t per = (nkb + ksplit - 1) / ksplit;
  const int64_t kb0 = ks * per, kb1 = (kb0 + per < nkb) ? kb0 + per : nkb;
  const int64_t my_kb = kb1 > kb0 ? kb1 - kb0 : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < 2; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), SPLIT_WARPS); }
    mbar_init(op_full, SPLIT_WARPS);
    mbar_init(op_empty, 1);
    mbar_init(acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int64_t it = 0; it < my_kb; ++it) {
        const int s = (int)(it & 1);
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), 2 * TN_LAND_BYTES);
        const int r0 = (int)((kb0 + it) * TN_BK);
        tma_load_2d(smem_u32(land + s * 2 * TN_LAND_BYTES), &tmA, full_bar(s), bm * 256, r0);
        tma_load_2d(smem_u32(land + s * 2 * TN_LAND_BYTES + TN_LAND_BYTES), &tmB, full_bar(s), bn * 256, r0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && my_kb > 0) {
      const uint32_t idesc = make_idesc<true>(BLOCK_M, 256);
      const uint32_t a_hi = smem_u32(ops), a_lo = a_hi + TN_OP_BYTES, b_hi = a_lo + TN_OP_BYTES, b_lo = b_hi + TN_OP_BYTES;
      for (int64_t it = 0; it < my_kb; ++it) {
        mbar_wait(op_full, (uint32_t)(it & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t dah = make_smem_desc<true>(a_hi + mt * 128 * 64), dal = make_smem_desc<true>(a_lo + mt * 128 * 64);
          const uint64_t dbh = make_smem_desc<true>(b_hi), dbl = make_smem_desc<true>(b_lo);
          const uint32_t acc = tmem_base + (uint32_t)mt * 256;
#pragma unroll
          for (int k = 0; k < TN_BK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            umma<true>(acc, dal + adv, dbh + adv, idesc, (it | k) != 0);
            umma<true>(acc, dah + adv, dbl + adv, idesc, 1);
            umma<true>(acc, dah + adv, dbh + adv, idesc, 1);
          }
        }
        umma_commit(op_empty);
      }
      umma_commit(acc_done);
    }
  } else if (warp < EPI_WARP0) {
    // ------------------------------------------------------------------ splitter / transposer (256 threads)
    const int t = threadIdx.x - 64;  // column of the landed tiles = row of the operand tiles
    const float asc = a_scale_ptr ? __ldg(a_scale_ptr) : 1.f, bsc = b_scale_ptr ? __ldg(b_scale_ptr) : 1.f;
    const int sw = (t >> 1) & 3;
    for (int64_t it = 0; it < my_kb; ++it) {
      const int s = (int)(it & 1);
      mbar_wait(full_bar(s), (uint32_t)((it >> 1) & 1));
#pragma unroll
      for (int op = 0; op < 2; ++op) {
        const float* src = reinterpret_cast<const float*>(land + s * 2 * TN_LAND_BYTES + op * TN_LAND_BYTES) + t;
        const float sc = op == 0 ? asc : bsc;
        float v[TN_BK];
#pragma unroll
        for (int k = 0; k < TN_BK; ++k) v[k] = src[k * 256] * sc;
        // the operand tiles are free once the MMAs of the previous step have completed
        if (op == 0 && it > 0) mbar_wait(op_empty, (uint32_t)((it - 1) & 1));
        uint8_t* hi_t = ops + op * 2 * TN_OP_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 16-byte piece c of row t = k elements 8c .. 8c+7
          uint4 hi, lo;
          split_f16x2(v[8 * c], v[8 * c + 1], hi.x, lo.x);
          split_f16x2(v[8 * c + 2], v[8 * c + 3], hi.y, lo.y);
          split_f16x2(v[8 * c + 4], v[8 * c + 5], hi.z, lo.z);
          split_f16x2(v[8 * c + 6], v[8 * c + 7], hi.w, lo.w);
          const int pos = t * 64 + ((c ^ sw) << 4);
          *reinterpret_cast<uint4*>(hi_t + pos) = hi;
          *reinterpret_cast<uint4*>(hi_t + TN_OP_BYTES + pos) = lo;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(s));  // both landed tiles of this stage are in registers / written out
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(op_full);
    }
  } else if (my_kb > 0) {
    // ------------------------------------------------------------------ epilogue (8 warps): C block += accumulators
    const int q = warp & 3, chalf = (warp - EPI_WARP0) >> 2;
    uint8_t* buf = epi_smem + (warp - EPI_WARP0) * EPI_BUF_BYTES;
    const float inv = (a_scale_ptr ? __ldg(a_scale_ptr + 1) : 1.f) * (b_scale_ptr ? __ldg(b_scale_ptr + 1) : 1.f);
    mbar_wait(acc_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mt * 256;
      for (int c0 = chalf * EPI_CHUNK; c0 < 256; c0 += 2 * EPI_CHUNK) {
        uint32_t v[32];
        tmem_ld32(tbase + (uint32_t)c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_float4(__uint_as_float(v[4 * j]) * inv, __uint_as_float(v[4 * j + 1]) * inv,
                          __uint_as_float(v[4 * j + 2]) * inv, __uint_as_float(v[4 * j + 3]) * inv);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmC, smem_u32(buf), bn * 256 + c0, bm * 256 + mt * 128 + q * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- weight preparation ---------------------------------------------------------------------
// scale slot (3 floats): [0] max |W| over the slot (as ordered uint bits), [1] 1/scale, [2] scale.
__global__ void weight_maxabs_kernel(const float* __restrict__ W, int64_t ldw, int K, int N, unsigned* __restrict__ slot) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)K * N; i += (int64_t)gridDim.x * blockDim.x) {
    const float w = fabsf(W[(i / N) * ldw + (i % N)]);
    if (w < INFINITY) m = fmaxf(m, w);  // NaN / inf weights do not define the scale
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(slot, __float_as_uint(m));
}
// power-of-two scale that brings max |W| into [2^13, 2^14): far from fp16 overflow, and the lo pieces
// of all but the (2^-17 relative) smallest weights stay normal
__device__ __forceinline__ float f16_weight_scale(const unsigned* slot) {
  const float m = __uint_as_float(*slot);
  if (!(m > 0.f)) return 1.f;
  int e;
  frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, 14 - e);
}
// W[K][N] (row stride ldw) -> Wt_hi, Wt_lo [Npad][K] (K-major); rows n >= N are zero.
template <bool F16>
__global__ void split_weight_kernel(const float* __restrict__ W, int64_t ldw, int K, int N, int Npad,
                                    void* __restrict__ hi_, void* __restrict__ lo_, float* __restrict__ slot) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  float scale = 1.f;
  if (F16) {
    scale = f16_weight_scale(reinterpret_cast<const unsigned*>(slot));
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { slot[1] = 1.f / scale; slot[2] = scale; }
  }
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i, n = n0 + tx;
    tile[i][tx] = (k < K && n < N) ? W[(int64_t)k * ldw + n] * scale : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + tx;
    if (n < Npad && k < K) {
      const float w = tile[tx][i];
      if (F16) {
        __half h, l;
        split_f16(w, h, l);
        reinterpret_cast<__half*>(hi_)[(int64_t)n * K + k] = h;
        reinterpret_cast<__half*>(lo_)[(int64_t)n * K + k] = l;
      } else {
        const float h = rna_tf32(w);
        reinterpret_cast<float*>(hi_)[(int64_t)n * K + k] = h;
        reinterpret_cast<float*>(lo_)[(int64_t)n * K + k] = rna_tf32(w - h);
      }
    }
  }
}

// planes[n][k_off + k] = W[n][k] * scale for n < N, k < Kpart (no transpose: W is already [N][K], row stride ldw)
template <bool F16>
__global__ void split_weight_nt_kernel(const float* __restrict__ W, int64_t ldw, int N, int Kpart, int k_off, int64_t ldp,
                                       void* __restrict__ hi_, void* __restrict__ lo_, float* __restrict__ slot) {
  float scale = 1.f;
  if (F16) {
    scale = f16_weight_scale(reinterpret_cast<const unsigned*>(slot));
    if (blockIdx.x == 0 && threadIdx.x == 0) { slot[1] = 1.f / scale; slot[2] = scale; }
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)N * Kpart; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / Kpart), k = (int)(i % Kpart);
    const float w = W[n * ldw + k] * scale;
    const int64_t o = n * ldp + k_off + k;
    if (F16) {
      __half h, l;
      split_f16(w, h, l);
      reinterpret_cast<__half*>(hi_)[o] = h;
      reinterpret_cast<__half*>(lo_)[o] = l;
    } else {
      const float h = rna_tf32(w);
      reinterpret_cast<float*>(hi_)[o] = h;
      reinterpret_cast<float*>(lo_)[o] = rna_tf32(w - h);
    }
  }
}

// slot = {s, 1/s} with s the power of two that brings max|v| into [1, 2)  (one block)
__global__ void pow2_scale_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ slot) {
  __shared__ float red[32];
  float m = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = fabsf(v[i]);
    if (a < INFINITY) m = fmaxf(m, a);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) {
      float sc = 1.f;
      if (m > 0.f) { int e; frexpf(m, &e); sc = ldexpf(1.f, 1 - e); }
      slot[0] = sc;
      slot[1] = 1.f / sc;
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D tensor [rows][cols] of fp32 (or fp16) with row stride ld (elements); box = 32 elements x box_rows, the box
// row being one swizzle span (128 B of fp32 / 64 B of fp16)
static int make_map(CUtensorMap* tm, const void* base, bool half, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -2;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (half ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 700 + (int)r;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace tc

// Tensor map over a row-major fp32 matrix [rows][cols] (row stride ld) with an arbitrary box and no swizzle (the
// landed tiles of the TN contraction) or the 32 x 32, 128B-swizzled box of the epilogue staging buffers.
static int make_map_box(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                        uint32_t box_rows, bool swizzle128) {
  tc::EncodeTiledFn enc = tc::get_encode();
  if (!enc) return -2;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 700 + (int)r;
}

int gemm_tn_tc_ok(const float* A, int64_t lda, const float* B, int64_t ldb, const float* C, int64_t ldc) {
  return (lda % 4) == 0 && (ldb % 4) == 0 && (ldc % 4) == 0 &&
         ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0;
}

// C[Ma, Nb] (ldc) += A[rows, Ma]^T (lda) @ B[rows, Nb] (ldb) / (a_scale * b_scale); scales = device {s, 1/s} or null
int gemm_tn_tc(const float* A, int64_t lda, int Ma, const float* B, int64_t ldb, int Nb, float* C, int64_t ldc, int64_t rows,
               const float* a_scale, const float* b_scale, cudaStream_t stream) {
  if (rows <= 0 || Ma <= 0 || Nb <= 0) return 0;
  if (!gemm_tn_tc_ok(A, lda, B, ldb, C, ldc) || rows > 0x7fffff00LL) return -2;
  CUtensorMap tmA, tmB, tmC;
  int rc;
  if ((rc = make_map_box(&tmA, A, (uint64_t)rows, (uint64_t)Ma, (uint64_t)lda, 256, tc::TN_BK, false))) return rc;
  if ((rc = make_map_box(&tmB, B, (uint64_t)rows, (uint64_t)Nb, (uint64_t)ldb, 256, tc::TN_BK, false))) return rc;
  if ((rc = make_map_box(&tmC, C, (uint64_t)Ma, (uint64_t)Nb, (uint64_t)ldc, 32, 32, true))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc::gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TN_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int nbm = (Ma + 255) / 256, nbn = (Nb + 255) / 256, nblk = nbm * nbn;
  const int64_t nkb = (rows + tc::TN_BK - 1) / tc::TN_BK;
  int ksplit = tc::num_sms() / nblk;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > nkb) ksplit = (int)nkb;
  tc::gemm_tn_tc_kernel<<<nblk * ksplit, tc::THREADS, tc::TN_SMEM_BYTES, stream>>>(tmA, tmB, tmC, rows, nbn, ksplit, a_scale, b_scale);
  return (int)cudaGetLastError();
}

// Device word that the fp16-piece kernels OR a bit into when an operand piece saturates (dh_plan_status); set by the
// API entry points for the plan they run (host launches of one plan are sequential).
static thread_local unsigned* g_range_flag = nullptr;
void range_flag_set(unsigned* flag) { g_range_flag = flag; }
unsigned* range_flag_get() { return g_range_flag; }

int gemm_tc_supported(int N, int K) { return N >= 1 && K >= 32 && K % 32 == 0; }
int gemm_tc_f16_ok(int K) { return K >= 32 && K % 32 == 0; }

int weight_maxabs_tc(const float* W, int64_t ldw, int K, int N, float* scale_slot, cudaStream_t stream) {
  const int64_t n = (int64_t)K * N;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 296) blocks = 296;
  tc::weight_maxabs_kernel<<<blocks, 256, 0, stream>>>(W, ldw, K, N, reinterpret_cast<unsigned*>(scale_slot));
  return (int)cudaGetLastError();
}

// pad_rows != 0: rows N..Npad-1 (Npad = N rounded up to 16) are written as zeros; otherwise exactly N
// rows are written so that blocks can be stacked.
int split_weight_tc(const float* W, int64_t ldw, int K, int N, int pad_rows, void* Wt_hi, void* Wt_lo,
                    float* scale_slot, int f16, cudaStream_t stream) {
  const int Npad = pad_rows ? ((N + 15) & ~15) : N;
  dim3 grid((K + 31) / 32, (Npad + 31) / 32);
  if (f16) tc::split_weight_kernel<true><<<grid, 256, 0, stream>>>(W, ldw, K, N, Npad, Wt_hi, Wt_lo, scale_slot);
  else tc::split_weight_kernel<false><<<grid, 256, 0, stream>>>(W, ldw, K, N, Npad, Wt_hi, Wt_lo, scale_slot);
  return (int)cudaGetLastError();
}

int pow2_scale_tc(const float* v, int64_t n, float* slot, cudaStream_t stream) {
  tc::pow2_scale_kernel<<<1, 1024, 0, stream>>>(v, n, slot);
  return (int)cudaGetLastError();
}

int split_weight_nt_tc(const float* W, int64_t ldw, int N, int Kpart, int k_off, int64_t ldp, void* hi, void* lo,
                       float* scale_slot, int f16, cudaStream_t stream) {
  const int64_t n = (int64_t)N * Kpart;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 592) blocks = 592;
  if (f16) tc::split_weight_nt_kernel<true><<<blocks, 256, 0, stream>>>(W, ldw, N, Kpart, k_off, ldp, hi, lo, scale_slot);
  else tc::split_weight_nt_kernel<false><<<blocks, 256, 0, stream>>>(W, ldw, N, Kpart, k_off, ldp, hi, lo, scale_slot);
  return (int)cudaGetLastError();
}

int gemm_tc(const float* A, const void* Wt_hi, const void* Wt_lo, const float* bias, const float* inv_scale, float* C,
            int64_t M, int N, int K, int64_t ldc, int rpg, int f16, int merged, cudaStream_t stream) {
  if (!gemm_tc_supported(N, K)) return -2;
  TcGemm g;
  g.A = A; g.lda = K; g.Wt_hi = Wt_hi; g.Wt_lo = Wt_lo; g.ldw = K; g.bias = bias; g.inv_scale = inv_scale;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.rpg = rpg; g.f16 = f16; g.merged = merged; g.reduce_add = 0;
  g.a_scale = nullptr;
  g.A_lo = nullptr;
  g.orb_env = nullptr; g.orb_Mj = nullptr; g.orb_L = 0;
  g.ln_res = nullptr; g.ln_gamma = nullptr; g.ln_beta = nullptr; g.ln_tanh = 0;
  return gemm_tc_ex(g, stream);
}

int gemm_tc_ex(const TcGemm& gm, cudaStream_t stream) {
  const float* A = gm.A; const void* Wt_hi = gm.Wt_hi; const void* Wt_lo = gm.Wt_lo;
  const float* bias = gm.bias; const float* inv_scale = gm.inv_scale; float* C = gm.C;
  const int64_t M = gm.M, ldc = gm.ldc; const int N = gm.N, K = gm.K, rpg = gm.rpg, f16 = gm.f16, merged = gm.merged;
  const int reduce_add = gm.reduce_add;
  if (M <= 0) return 0;
  if (N < 1 || K < 1 || gm.ldw % 32 != 0 || gm.ldw < K || (gm.lda % 4) != 0 || M > 0x7fffff00LL) return -2;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Wt_hi) & 15) ||
      (reinterpret_cast<uintptr_t>(Wt_lo) & 15))
    return -1;
  const int Npad = (N + 15) & ~15;
  CUtensorMap tmA, tmBh, tmBl, tmC;
  int rc;
  const int a_pre = gm.A_lo != nullptr ? 1 : 0;  // A given as fp16 hi / lo planes ([M][lda] halves each)
  if (a_pre && (!f16 || gm.a_scale != nullptr || (reinterpret_cast<uintptr_t>(gm.A_lo) & 15) || (gm.lda % 8) != 0)) return -2;
  CUtensorMap tmAlo;
  if ((rc = tc::make_map(&tmA, A, a_pre != 0, (uint64_t)M, (uint64_t)K, (uint64_t)gm.lda, tc::BLOCK_M))) return rc;
  if (a_pre) { if ((rc = tc::make_map(&tmAlo, gm.A_lo, true, (uint64_t)M, (uint64_t)K, (uint64_t)gm.lda, tc::BLOCK_M))) return rc; }
  else tmAlo = tmA;
  const int tma_store = ((reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc % 4) == 0) ? 1 : 0;
  if (tma_store) {
    if ((rc = tc::make_map(&tmC, C, false, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32))) return rc;
  } else {
    tmC = tmA;  // unused
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc::gemm_tc_kernel<false, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tc_kernel<false, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tc_kernel<true, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tc_kernel<true, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tc_kernel<true, true, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tc_kernel<true, true, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tc_kernel<true, true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int64_t bands = (M + tc::BLOCK_M - 1) / tc::BLOCK_M;
  const int64_t tiles = bands * ((N + tc::BLOCK_N - 1) / tc::BLOCK_N);
  const int sms = tc::num_sms();
  // CTA pairs (tcgen05.mma.cta_group::2): fp16 pieces with the merged accumulator, fp32 A; DH_GEMM_PAIR=0 disables
  static const bool pair_env = !(dbg_env("DH_GEMM_PAIR") && atoi(dbg_env("DH_GEMM_PAIR")) == 0);
  // (the fused value LayerNorm exists for the pair form only: a single band runs as a pair whose second band is out of range --
  // its loads are zero-filled and its stores clipped -- so that every value pass takes the same arithmetic path)
  const bool pair = pair_env && f16 && merged && !a_pre && (bands >= 2 || gm.ln_res != nullptr || gm.orb_env != nullptr);
  // cluster size: DH_GEMM_CLUSTER = 1 | 2 | 4; small problems run un-clustered.  Default 1: with the
  // 192 KB operand ring the kernel is bound by shared-memory bandwidth, not by L2 -> SM traffic, and the
  // multicast measured no faster at 2 and slower at 4 (profiles/r1_gemm_tc_notes.md).
  static const int cl_env = dbg_env("DH_GEMM_CLUSTER") ? atoi(dbg_env("DH_GEMM_CLUSTER")) : 1;
  int cl = (cl_env == 4 || cl_env == 2) ? cl_env : 1;
  if (bands < 2 * sms || pair) cl = 1;
  int64_t g = bands < sms ? bands : sms;
  g = g / cl * cl;
  if (pair) {  // one cluster of two CTAs per pair of bands, at most one CTA per SM
    const int64_t units = (bands + 1) / 2;
    g = 2 * (units < sms / 2 ? units : sms / 2);
  }
  dim3 grid((unsigned)g);
  const uint32_t b_box = pair ? tc::BLOCK_N / 2 : tc::BLOCK_N / cl;
  if ((rc = tc::make_map(&tmBh, Wt_hi, f16 != 0, (uint64_t)Npad, (uint64_t)gm.ldw, (uint64_t)gm.ldw, b_box))) return rc;
  if ((rc = tc::make_map(&tmBl, Wt_lo, f16 != 0, (uint64_t)Npad, (uint64_t)gm.ldw, (uint64_t)gm.ldw, b_box))) return rc;
  static const bool want_prof = dbg_env("DH_GEMM_PROF") != nullptr;
  unsigned long long* prof = nullptr;
  if (want_prof) {
    static unsigned long long* buf = nullptr;
    if (!buf) cudaMalloc(&buf, 40 * sizeof(unsigned long long));
    cudaMemsetAsync(buf, 0, 40 * sizeof(unsigned long long), stream);
    prof = buf;
  }
  // resident A (pair form): K <= 256 and more than one column tile per band; DH_GEMM_RES=0 disables
  static const bool res_env = !(dbg_env("DH_GEMM_RES") && atoi(dbg_env("DH_GEMM_RES")) == 0);
  const int res = (pair && res_env && K <= 8 * tc::BLOCK_K && N > tc::BLOCK_N) ? 1 : 0;
  // fused envelope contraction (orbital matrices): pair form with resident A, 32 jet rows per electron, whole electrons per
  // band, weights in the permuted [tile][m][re | im] layout (10 m per 256-column tile)
  tc::OrbFuse orb;
  orb.env = gm.orb_env; orb.Mj = gm.orb_Mj; orb.L = gm.orb_L; orb.mpt = gm.orb_L > 0 ? orb_per_tile(gm.orb_L) : tc::ORB_MPT;
  const bool orb_on = gm.orb_env != nullptr;
  if (orb_on && !(pair && res && ((rpg == 32 && M % 128 == 0) || rpg <= 1) && !reduce_add && gm.a_scale == nullptr && gm.orb_L >= 1 && gm.orb_L <= 48 &&
                  N == orb_columns(gm.orb_L)))
    return -2;
  // fused value LayerNorm epilogue: pair form, one 256-wide column tile, every row a value row, TMA-stored output in place
  tc::LnFuse lnf;
  lnf.res = gm.ln_res; lnf.gamma = gm.ln_gamma; lnf.beta = gm.ln_beta; lnf.tanh_mode = gm.ln_tanh;
  const bool ln_on = gm.ln_res != nullptr;
  if (ln_on && !(pair && !orb_on && N == tc::BLOCK_N && rpg <= 1 && tma_store && !reduce_add && gm.a_scale == nullptr &&
                 gm.ln_gamma && gm.ln_beta && ((reinterpret_cast<uintptr_t>(gm.ln_res) | reinterpret_cast<uintptr_t>(gm.ln_gamma) |
                                                reinterpret_cast<uintptr_t>(gm.ln_beta) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0))
    return -2;
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = grid;
  lc.blockDim = dim3(tc::THREADS);
  lc.dynamicSmemBytes = tc::SMEM_BYTES;
  lc.stream = stream;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeClusterDimension;
  lattr[0].val.clusterDim.x = pair ? 2u : (unsigned)cl;
  lattr[0].val.clusterDim.y = 1;
  lattr[0].val.clusterDim.z = 1;
  lc.attrs = lattr;
  lc.numAttrs = 1;
  cudaError_t le;
#define DH_LAUNCH_TC(F, MG, PR) le = cudaLaunchKernelEx(&lc, tc::gemm_tc_kernel<F, MG, PR, false, false>, tmA, tmAlo, tmBh, tmBl, tmC, bias, inv_scale, \
                                                        C, M, N, K, ldc, rpg, tma_store, reduce_add, cl, gm.a_scale, a_pre, res, orb, lnf, prof, \
                                                        g_range_flag)
  if (orb_on)
    le = cudaLaunchKernelEx(&lc, tc::gemm_tc_kernel<true, true, true, true, false>, tmA, tmAlo, tmBh, tmBl, tmC, bias, inv_scale, C, M, N, K,
                            ldc, rpg, tma_store, reduce_add, cl, gm.a_scale, a_pre, res, orb, lnf, prof, g_range_flag);
  else if (ln_on)
    le = cudaLaunchKernelEx(&lc, tc::gemm_tc_kernel<true, true, true, false, true>, tmA, tmAlo, tmBh, tmBl, tmC, bias, inv_scale, C, M, N, K,
                            ldc, rpg, tma_store, reduce_add, cl, gm.a_scale, a_pre, res, orb, lnf, prof, g_range_flag);
  else if (pair) DH_LAUNCH_TC(true, true, true);
  else if (f16) { if (merged) DH_LAUNCH_TC(true, true, false); else DH_LAUNCH_TC(true, false, false); }
  else { if (merged) DH_LAUNCH_TC(false, true, false); else DH_LAUNCH_TC(false, false, false); }
#undef DH_LAUNCH_TC
  if (le != cudaSuccess) return (int)le;
  if (want_prof) {
    unsigned long long h[40];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
    const double g = (double)grid.x, tp = (double)tiles / g;
    fprintf(stderr, "[gemm_tc M=%lld N=%d f16=%d] per CTA: tiles %.1f | total cyc %.0f | producer wait empty %.0f | "
            "mma wait tmem_empty %.0f full %.0f split %.0f | splitter wait full %.0f work %.0f | epi wait tmem_full %.0f drain %.0f "
            "[tmem ld+wait %.0f, math+bias %.0f, wait_group.read %.0f, st.shared+fence %.0f, tma issue %.0f]\n",
            (long long)M, N, f16, tp, h[11] / g, h[0] / g, h[1] / g, h[2] / g, h[3] / g, h[4] / g, h[5] / g, h[6] / g, h[7] / g,
            h[16] / g, h[17] / g, h[18] / g, h[19] / g, h[20] / g);
  }
  return (int)cudaGetLastError();
}

}  // namespace dh
