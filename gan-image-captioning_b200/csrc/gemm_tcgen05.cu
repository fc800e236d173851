// Blackwell tensor-core GEMM for the dense contractions of the hot path (sm_100a only):
//   * operands fp32 in HBM, fed as TF32 (kind::tf32) -- no conversion pass, fp32 accumulation in TMEM
//   * TMA (cp.async.bulk.tensor, 128B swizzle) stages A/B tiles into a 6-deep shared-memory ring
//   * one elected thread issues tcgen05.mma (UMMA 128 x BN x 8), accumulator 128 lanes x BN columns in TMEM
//   * 4 epilogue warps read TMEM with tcgen05.ld and apply alpha / beta / bias before storing fp32
//   * both operand layouts are supported natively: K-major (A[M,K], B[N,K]) and MN-major (A^T stored [K,M],
//     B stored [K,N]) through the UMMA descriptor "major" bits, so the backward GEMMs need no transposes
// C[M,N] = alpha * op(A) * op(B) + beta * C + bias[N]
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (TMEM lane quarter = warp_id % 4).  One 128 x BN output tile per CTA.
#include <cuda.h>

#include <map>
#include <tuple>

#include "gic_internal.cuh"

namespace gic {

namespace tc {

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BK = 32;         // fp32 elements per stage along K = one 128-byte swizzle row
constexpr int UMMA_K = 8;      // kind::tf32: 32 bytes of K per instruction
constexpr int NTHREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// UMMA shared-memory descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (sm_100) | [61,64) layout type = 2 (SWIZZLE_128B)
//   layout type 2 = SWIZZLE_128B (16-byte atoms; K-major tiles), 1 = SWIZZLE_128B_BASE32B (32-byte atoms: the only
//   layout the hardware accepts for MN-major tf32 operands, cutlass sm100_common.inl:92)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10),
// [10,13), a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int a_mn, int b_mn, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 4;     // 16 KB
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 128) ? 6 : 8;
  static constexpr int TOTAL = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
};

// A_MN / B_MN: operand is MN-major in global memory (contiguous along M resp. N), loaded as 32-wide MN slabs.
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                 float alpha, float beta, float* __restrict__ C, int ldc, const float* __restrict__ bias, int vecC) {
  using S = Smem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE);
  uint64_t* empty = full + S::STAGES;
  uint64_t* tmem_full = empty + S::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // split-K: gridDim.z CTAs share one output tile, each owns a contiguous range of k-blocks and adds its partial
  // sum with fp32 atomics (the host pre-scales C by beta).  kb_lo/kb_hi are this CTA's range.
  const int nkb_all = (K + BK - 1) / BK;
  const int kb_lo = (int)(((long long)nkb_all * blockIdx.z) / gridDim.z);
  const int kb_hi = (int)(((long long)nkb_all * (blockIdx.z + 1)) / gridDim.z);
  const int nkb = kb_hi - kb_lo;
  const bool split = gridDim.z > 1;
  constexpr uint32_t TMEM_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < S::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S::STAGES;
        const uint32_t ph = (kb / S::STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * S::STAGE;
        uint8_t* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full[s], S::STAGE);
        const int k0 = (kb_lo + kb) * BK;
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + j * (BK * 128), &tmA, &full[s], m0 + 32 * j, k0);
        } else {
          tma_load_2d(sa, &tmA, &full[s], k0, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + j * (BK * 128), &tmB, &full[s], n0 + 32 * j, k0);
        } else {
          tma_load_2d(sb, &tmB, &full[s], k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(A_MN ? 1 : 0, B_MN ? 1 : 0, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S::STAGES;
        const uint32_t ph = (kb / S::STAGES) & 1;
        mbar_wait(&full[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::STAGE);
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major (SWIZZLE_128B): 8 tf32 = 32 bytes further along the swizzled 128-byte row; groups of 8 rows
          //   are 1024 B apart (SBO).
          // MN-major (SWIZZLE_128B_BASE32B): a row is 32 MN elements (128 B) of one k; the swizzle atom is 4 k-rows
          //   = 512 B (SBO); 32-wide MN slabs are BK*128 B apart (LBO); one MMA consumes 8 k-rows = 1024 B.
          const uint64_t da = A_MN ? make_desc(sa + k * 1024, BK * 128, 512, 1) : make_desc(sa + k * 32, 16, 1024, 2);
          const uint64_t db = B_MN ? make_desc(sb + k * 1024, BK * 128, 512, 1) : make_desc(sb + k * 32, 16, 1024, 2);
          umma_tf32(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);          // frees the smem slot when these MMAs retire
      }
      umma_commit(tmem_full);            // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> smem transpose -> coalesced global stores =====
    // tcgen05.ld hands every lane one accumulator ROW (32 consecutive columns); storing that directly would make
    // each warp store touch 32 different rows.  Each warp transposes its 32 x 32 block through a padded smem
    // staging tile (the pipeline buffers are idle once tmem_full has fired) so that one store / atomic instruction
    // covers 128 contiguous bytes of one output row.
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    float* stg = reinterpret_cast<float*>(smem) + q * (32 * 33);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (n0 + c0 >= N) continue;        // warp-uniform: nothing to store from this chunk
      if (split) {
        // split-K: transpose the warp's 32 x 32 block through smem so one atomic instruction covers 128 contiguous
        // bytes of one output row (8x fewer L2 atomic sector operations than one row per lane)
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = alpha * __uint_as_float(r[j]);
        __syncwarp();
        const int n = n0 + c0 + lane;
        const bool n_ok = n < N;
        const float bv = (bias != nullptr && n_ok && blockIdx.z == 0) ? bias[n] : 0.f;
        const int m_base = m0 + q * 32;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
          const int m = m_base + i;
          if (m < M && n_ok) atomicAdd(C + (size_t)m * ldc + n, stg[i * 33 + lane] + bv);
        }
        __syncwarp();
      } else {
        // direct: each lane owns one output row and writes 32 consecutive floats as 8 float4 stores
        const int m = m0 + q * 32 + lane;
        if (m < M) {
          float* crow = C + (size_t)m * ldc;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = n0 + c0 + j;
            if (n >= N) break;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = alpha * __uint_as_float(r[j + e]);
            if (vecC && n + 3 < N) {
              if (bias) {
                const float4 bb = *reinterpret_cast<const float4*>(bias + n);
                v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
              }
              if (beta != 0.f) {
                const float4 cc = *reinterpret_cast<const float4*>(crow + n);
                v[0] += beta * cc.x; v[1] += beta * cc.y; v[2] += beta * cc.z; v[3] += beta * cc.w;
              }
              *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (n + e < N) {
                  float x = v[e];
                  if (bias) x += bias[n + e];
                  if (beta != 0.f) x += beta * crow[n + e];
                  crow[n + e] = x;
                }
              }
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point fetched at run time: the library does not link libcuda)
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] (cols contiguous, leading dimension ld) with a box of box_cols x box_rows.
// dtype TFLOAT32: the TMA unit rounds fp32 -> tf32 while staging (unbiased, unlike the MMA's own truncation).
static bool make_map(CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_cols, int box_rows, bool rn,
                     bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, rn ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, float alpha, float beta, float* C,
                  int ldc, const float* bias, int splits, cudaStream_t s) {
  using S = Smem<BN>;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gemm_tf32_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    attr = true;
  }
  const int vecC = (aligned16(C) && (ldc % 4 == 0) && (!bias || aligned16(bias))) ? 1 : 0;
  dim3 grid(cdiv(N, BN), cdiv(M, BM), splits);
  gemm_tf32_kernel<BN, A_MN, B_MN><<<grid, NTHREADS, S::TOTAL, s>>>(ta, tb, M, N, K, alpha, beta, C, ldc, bias, vecC);
  return check_launch("gemm_tf32_kernel");
}

static bool tf32_round_in_tma() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GIC_TMA_TF32_RN");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace tc

// Returns handled = false (and launches nothing) when the operands do not meet TMA's constraints
// (16-byte aligned base, leading dimension a multiple of 4 floats); the caller then uses the FFMA kernel.
int gemm_tc(int mode, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
            const float* B, int ldb, float beta, float* C, int ldc, const float* bias, cudaStream_t stream,
            bool* handled) {
  using namespace tc;
  *handled = false;
  if (mode != GEMM_TF32) return GIC_OK;                 // GEMM_TF32X3 is served by the exact-fp32 kernel for now
  if (M <= 0 || N <= 0 || K <= 0) return GIC_OK;
  if (!aligned16(A) || !aligned16(B) || (lda % 4) || (ldb % 4)) return GIC_OK;
  if ((long long)M * N < 64 * 64 || K < 32) return GIC_OK;   // tiny problems: launch-latency bound either way
  const bool rn = tf32_round_in_tma();
  const bool a_mn = transA;       // A stored [K, M]: contiguous along M
  const bool b_mn = !transB;      // B stored [K, N]: contiguous along N
  const int BN = (N <= 64) ? 64 : 128;
  CUtensorMap ta, tb;
  bool ok;
  if (a_mn) ok = make_map(&ta, A, K, M, lda, 32, BK, rn, true);           // [K rows, M cols], box 32 (M) x 32 (K)
  else      ok = make_map(&ta, A, M, K, lda, BK, BM, rn, false);          // [M rows, K cols], box 32 (K) x 128 (M)
  if (ok) {
    if (b_mn) ok = make_map(&tb, B, K, N, ldb, 32, BK, rn, true);         // [K rows, N cols], box 32 (N) x 32 (K)
    else      ok = make_map(&tb, B, N, K, ldb, BK, BN, rn, false);        // [N rows, K cols], box 32 (K) x BN (N)
  }
  if (!ok) return GIC_OK;
  // split-K when the output tiles alone cannot fill the machine (long-K / small-MN shapes of the backward pass)
  int splits = 1;
  {
    const int tiles = cdiv(N, BN) * cdiv(M, BM), nkb = cdiv(K, BK);
    if (tiles * 2 <= num_sms() && nkb >= 8 && (beta == 0.f || beta == 1.f)) {
      splits = min(min(cdiv(num_sms(), tiles), nkb / 4), 32);
      if (splits < 1) splits = 1;
    }
    if (splits > 1 && beta == 0.f) {
      cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, stream);
      if (e != cudaSuccess) { set_error("memset2D: %s", cudaGetErrorString(e)); return GIC_ERR_CUDA; }
    }
  }
  int rc;
#define GIC_TC(BN_)                                                                                        \
  do {                                                                                                     \
    if (!a_mn && !b_mn) rc = launch<BN_, false, false>(ta, tb, M, N, K, alpha, beta, C, ldc, bias, splits, stream); \
    else if (!a_mn && b_mn) rc = launch<BN_, false, true>(ta, tb, M, N, K, alpha, beta, C, ldc, bias, splits, stream); \
    else if (a_mn && !b_mn) rc = launch<BN_, true, false>(ta, tb, M, N, K, alpha, beta, C, ldc, bias, splits, stream); \
    else rc = launch<BN_, true, true>(ta, tb, M, N, K, alpha, beta, C, ldc, bias, splits, stream);                  \
  } while (0)
  if (BN == 64) GIC_TC(64); else GIC_TC(128);
#undef GIC_TC
  if (rc == GIC_OK) *handled = true;
  return rc;
}

}  // namespace gic
