// tcgen05 / TMEM / TMA dense contraction with fp32 accuracy by 3xTF32 splitting (sm_100a).
//
//   C[M, N] = A[M, K] @ W[K, N] (+ bias[N] on rows m with m % rpg == 0)
//
// A is the fp32 activation matrix (row-major, K contiguous).  W is given pre-transposed and
// pre-split: Wt_hi / Wt_lo are [Npad][K] fp32 arrays holding tf32-exact values with
// W = hi + lo + O(2^-22 |W|)  (split_weight_tc, once per parameter update).
// A is split inside the kernel: TMA lands the fp32 tile in shared memory (128B swizzle), four
// warps rewrite it in place as hi = rna_tf32(a) and write lo = rna_tf32(a - hi) to a twin buffer;
// one elected thread then issues, per 32-float K block, 4 x 3 tcgen05.mma.kind::tf32
// (lo*hi, hi*lo, hi*hi) accumulating in fp32 in tensor memory.
//
// PERSISTENT kernel, one CTA per SM, 320 threads, CTA tile 128 x (<=256) x K:
//   warp 0      TMA producer           (2-stage ring of 96 KB stages, runs ahead across tiles)
//   warp 1      TMEM allocator + MMA issuer
//   warps 2..5  splitter               (fp32 -> tf32 hi / lo, in shared memory)
//   warps 6..9  epilogue               (tcgen05.ld -> +bias -> swizzled smem staging -> TMA store)
// so the loads and the split of tile t+1 overlap the drain and the stores of tile t.  The two
// accumulators (main, correction) fill all 512 TMEM columns, so the first MMA of tile t+1 waits
// until the epilogue has read tile t out of TMEM (not until its stores have landed).
#include <cuda.h>

#include "kernels.h"

namespace dh {

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 32;  // floats = one 128-byte swizzle row
constexpr int UMMA_K = 8;    // tf32
constexpr int STAGES = 2;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 4;  // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 96 KB
constexpr int EPI_CHUNK = 32;                            // accumulator columns per epilogue step
constexpr int EPI_BUF_BYTES = 32 * EPI_CHUNK * 4;        // 32 rows x 32 floats = 4 KB (one warp, one step)
constexpr int EPI_BYTES = 4 * 2 * EPI_BUF_BYTES;         // 4 warps x 2 buffers = 32 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int THREADS = 320;
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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B: 8-row groups 1024 B apart (SBO), LBO unused (=1), version 1 (sm_100)
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  // c_format F32 (1) @4, a/b format TF32 (2) @7/@10, K-major both, N>>3 @17, M>>4 @24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

// tma_store != 0: C is written through tmC (box 32 x 32 floats, 128B swizzle); otherwise (ldc not a
// multiple of 4 floats, or misaligned C) by direct global stores.
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmC,
               const float* __restrict__ bias, float* __restrict__ C, int64_t M, int N, int K, int64_t ldc, int rpg,
               int tma_store) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + EPI_BYTES);
  // bars: [0..S) full, [S..2S) split, [2S..3S) empty, [3S] tmem_full, [3S+1] tmem_empty ; then tmem ptr
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 2);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto split_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (3 * STAGES);
  const uint32_t tmem_empty_bar = bar_base + 8u * (3 * STAGES + 1);
  auto a_hi = [&](int s) { return smem_base + s * STAGE_BYTES; };
  auto a_lo = [&](int s) { return smem_base + s * STAGE_BYTES + A_BYTES; };
  auto b_hi = [&](int s) { return smem_base + s * STAGE_BYTES + 2 * A_BYTES; };
  auto b_lo = [&](int s) { return smem_base + s * STAGE_BYTES + 2 * A_BYTES + B_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = K / BLOCK_K;
  const int n_ntiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int64_t n_mtiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int64_t total_tiles = n_mtiles * n_ntiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
    if (tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (int)((tile / n_ntiles) * BLOCK_M);
        const int n0 = (int)(tile % n_ntiles) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_arrive_expect_tx(full_bar(s), A_BYTES + 2 * B_BYTES);
          tma_load_2d(a_hi(s), &tmA, full_bar(s), kb * BLOCK_K, m0);
          tma_load_2d(b_hi(s), &tmBhi, full_bar(s), kb * BLOCK_K, n0);
          tma_load_2d(b_lo(s), &tmBlo, full_bar(s), kb * BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const int n0 = (int)(tile % n_ntiles) * BLOCK_N;
        int n_tile = N - n0 < BLOCK_N ? N - n0 : BLOCK_N;
        n_tile = (n_tile + 15) & ~15;  // Wt buffers are zero-padded to a multiple of 16 rows
        const uint32_t idesc = make_idesc(BLOCK_M, n_tile);
        mbar_wait(tmem_empty_bar, (tl & 1) ^ 1);  // epilogue has read the previous tile out of TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          mbar_wait(split_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t dah = make_smem_desc(a_hi(s)), dal = make_smem_desc(a_lo(s));
          const uint64_t dbh = make_smem_desc(b_hi(s)), dbl = make_smem_desc(b_lo(s));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);  // start-address field is in 16-byte units
            // The tensor core rounds toward zero when it adds into the fp32 accumulator, a bias that
            // grows with the number of additions: keep the small correction terms in their own
            // accumulator so the main one sees K/8 additions instead of 3K/8.
            umma_tf32(tmem_base + BLOCK_N, dal + adv, dbh + adv, idesc, (kb | k) != 0);
            umma_tf32(tmem_base + BLOCK_N, dah + adv, dbl + adv, idesc, 1);
            umma_tf32(tmem_base, dah + adv, dbh + adv, idesc, (kb | k) != 0);
          }
          umma_commit(empty_bar(s));  // arrives when the MMAs reading this stage have completed
        }
        umma_commit(tmem_full_bar);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ splitter
    const int t = threadIdx.x - 64;  // 0..127
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        float4* hi = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + A_BYTES);
#pragma unroll
        for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
          const int idx = t + i * 128;
          float4 v = hi[idx], h, l;
          h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
          l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
          hi[idx] = h;
          lo[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(s));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    uint8_t* my_epi = epi_smem + (warp - 6) * 2 * EPI_BUF_BYTES;
    uint32_t tl = 0, nstore = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      const int64_t m0 = (tile / n_ntiles) * BLOCK_M;
      const int n0 = (int)(tile % n_ntiles) * BLOCK_N;
      int n_tile = N - n0 < BLOCK_N ? N - n0 : BLOCK_N;
      n_tile = (n_tile + 15) & ~15;
      const int64_t m = m0 + row;
      const bool add_bias = bias != nullptr && (rpg <= 1 || (m % rpg) == 0);
      mbar_wait(tmem_full_bar, tl & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = 0; c0 < n_tile; c0 += EPI_CHUNK) {
        uint32_t v[32], w[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + BLOCK_N, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + EPI_CHUNK >= n_tile) {
          // last read of this tile's accumulators: hand TMEM back to the MMA issuer
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        if (add_bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < N) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(bias + n0 + c0 + j));
        }
        if (tma_store) {
          // staging buffer `nstore & 1` of this warp: wait until the TMA store issued two steps ago has read it
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          uint8_t* buf = my_epi + (nstore & 1) * EPI_BUF_BYTES;
          // row `lane` of the 32 x 32 chunk; 16-byte piece j lives at piece position j ^ (lane & 7) (128B swizzle)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                   __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, smem_u32(buf), n0 + c0, (int)(m0 + q * 32));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++nstore;
        } else if (m < M) {
          float* crow = C + m * ldc + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < N) crow[j] = __uint_as_float(v[j]);
        }
      }
    }
    if (tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// W[K][N] (row stride ldw) -> Wt_hi, Wt_lo [Npad][K]; rows n >= N are zero.
__global__ void split_weight_kernel(const float* __restrict__ W, int64_t ldw, int K, int N, int Npad,
                                    float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i, n = n0 + tx;
    tile[i][tx] = (k < K && n < N) ? W[(int64_t)k * ldw + n] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + tx;
    if (n < Npad && k < K) {
      const float w = tile[tx][i];
      const float h = rna_tf32(w);
      hi[(int64_t)n * K + k] = h;
      lo[(int64_t)n * K + k] = rna_tf32(w - h);
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

// 2-D fp32 tensor [rows][cols] with row stride ld (floats); box = 32 floats x box_rows, 128B swizzle
static int make_map(CUtensorMap* tm, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -2;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

int gemm_tc_supported(int N, int K) { return N >= 1 && K >= tc::BLOCK_K && K % tc::BLOCK_K == 0; }

int split_weight_tc(const float* W, int64_t ldw, int K, int N, float* Wt_hi, float* Wt_lo, cudaStream_t stream) {
  const int Npad = (N + 15) & ~15;
  dim3 grid((K + 31) / 32, (Npad + 31) / 32);
  tc::split_weight_kernel<<<grid, 256, 0, stream>>>(W, ldw, K, N, Npad, Wt_hi, Wt_lo);
  return (int)cudaGetLastError();
}

// same, but writes exactly N rows (no zero pad rows) so that blocks can be stacked
int split_weight_tc_rows(const float* W, int64_t ldw, int K, int N, float* Wt_hi, float* Wt_lo, cudaStream_t stream) {
  dim3 grid((K + 31) / 32, (N + 31) / 32);
  tc::split_weight_kernel<<<grid, 256, 0, stream>>>(W, ldw, K, N, N, Wt_hi, Wt_lo);
  return (int)cudaGetLastError();
}

int gemm_tc(const float* A, const float* Wt_hi, const float* Wt_lo, const float* bias, float* C, int64_t M, int N,
            int K, int64_t ldc, int rpg, int accumulate, cudaStream_t stream) {
  if (M <= 0) return 0;
  if (!gemm_tc_supported(N, K) || accumulate || M > 0x7fffff00LL) return -2;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Wt_hi) & 15) ||
      (reinterpret_cast<uintptr_t>(Wt_lo) & 15))
    return -1;
  const int Npad = (N + 15) & ~15;
  CUtensorMap tmA, tmBh, tmBl, tmC;
  int rc;
  if ((rc = tc::make_map(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)K, tc::BLOCK_M))) return rc;
  if ((rc = tc::make_map(&tmBh, Wt_hi, (uint64_t)Npad, (uint64_t)K, (uint64_t)K, tc::BLOCK_N))) return rc;
  if ((rc = tc::make_map(&tmBl, Wt_lo, (uint64_t)Npad, (uint64_t)K, (uint64_t)K, tc::BLOCK_N))) return rc;
  const int tma_store = ((reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc % 4) == 0) ? 1 : 0;
  if (tma_store) {
    if ((rc = tc::make_map(&tmC, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32))) return rc;
  } else {
    tmC = tmA;  // unused
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc::gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int64_t tiles = ((M + tc::BLOCK_M - 1) / tc::BLOCK_M) * ((N + tc::BLOCK_N - 1) / tc::BLOCK_N);
  const int sms = tc::num_sms();
  dim3 grid((unsigned)(tiles < sms ? tiles : sms));
  tc::gemm_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(tmA, tmBh, tmBl, tmC, bias, C, M, N, K, ldc, rpg, tma_store);
  return (int)cudaGetLastError();
}

}  // namespace dh
