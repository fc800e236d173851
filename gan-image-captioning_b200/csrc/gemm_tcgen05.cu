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
#include "tcgen05_common.cuh"

namespace gic {

namespace tc {

template <int BN>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 4;     // 16 KB
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = (196 * 1024 / STAGE) > 8 ? 8 : (196 * 1024 / STAGE);
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
    float* stg = reinterpret_cast<float*>(smem) + q * (32 * 17);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      if (n0 + c0 >= N) continue;        // warp-uniform: nothing to store from this chunk
      if (split) {
        // split-K: transpose the warp's 32 x 16 block through smem so that one atomic instruction covers two
        // 64-byte row segments instead of 32 different rows (fewer L2 atomic sector operations)
#pragma unroll
        for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = alpha * __uint_as_float(r[j]);
        __syncwarp();
        const int cj = lane & 15, rh = lane >> 4;      // column within the chunk, row parity
        const int n = n0 + c0 + cj;
        const bool n_ok = n < N;
        const float bv = (bias != nullptr && n_ok && blockIdx.z == 0) ? bias[n] : 0.f;
        const int m_base = m0 + q * 32;
#pragma unroll 4
        for (int i = 0; i < 32; i += 2) {
          const int m = m_base + i + rh;
          if (m < M && n_ok) atomicAdd(C + (size_t)m * ldc + n, stg[(i + rh) * 17 + cj] + bv);
        }
        __syncwarp();
      } else {
        // direct: each lane owns one output row and writes 16 consecutive floats as 4 float4 stores
        const int m = m0 + q * 32 + lane;
        if (m < M) {
          float* crow = C + (size_t)m * ldc;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
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

}  // namespace tc

// Returns handled = false (and launches nothing) when the operands do not meet TMA's constraints
// (16-byte aligned base, leading dimension a multiple of 4 floats); the caller then uses the FFMA kernel.
int gemm_tc(int mode, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
            const float* B, int ldb, float beta, float* C, int ldc, const float* bias, cudaStream_t stream,
            bool* handled) {
  using namespace tc;
  *handled = false;
  if (mode != GEMM_TF32) return GIC_OK;
  if (M <= 0 || N <= 0 || K <= 0) return GIC_OK;
  if (!aligned16(A) || !aligned16(B) || (lda % 4) || (ldb % 4)) return GIC_OK;
  if ((long long)M * N < 64 * 64 || K < 32) return GIC_OK;   // tiny problems: launch-latency bound either way
  const bool rn = tf32_round_in_tma();
  const bool a_mn = transA;       // A stored [K, M]: contiguous along M
  const bool b_mn = !transB;      // B stored [K, N]: contiguous along N
  // tile width: minimise waves x bytes-per-k-block (the main loop is bound by L2 -> smem traffic, (128 + BN) rows of
  // 128 B per k-block), e.g. N = 10000, M = 256 -> BN = 144 (140 CTAs, one wave) instead of 128 (158 CTAs, two waves)
  static const int kBN[6] = {64, 128, 144, 192, 240, 256};
  int BN = 128;
  {
    long long best = -1;
    for (int i = 0; i < 6; ++i) {
      const int bn = kBN[i];
      if (b_mn && (bn % 32)) continue;                 // MN-major B is loaded as 32-wide slabs
      const long long tiles = (long long)cdiv(N, bn) * cdiv(M, BM);
      const long long cost = (long long)cdiv(tiles, num_sms()) * (BM + bn) + bn / 64;   // tie-break: smaller tile
      if (best < 0 || cost < best) { best = cost; BN = bn; }
    }
  }
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
  int rc = GIC_OK;
#define GIC_TC(BN_)                                                                                                \
  do {                                                                                                             \
    if (!a_mn && !b_mn) rc = launch<BN_, false, false>(ta, tb, M, N, K, alpha, beta, C, ldc, bias, splits, stream); \
    else if (a_mn && !b_mn) rc = launch<BN_, true, false>(ta, tb, M, N, K, alpha, beta, C, ldc, bias, splits, stream); \
    else if constexpr ((BN_ % 32) == 0) {                                                                          \
      if (!a_mn) rc = launch<BN_, false, true>(ta, tb, M, N, K, alpha, beta, C, ldc, bias, splits, stream);          \
      else rc = launch<BN_, true, true>(ta, tb, M, N, K, alpha, beta, C, ldc, bias, splits, stream);                 \
    }                                                                                                              \
  } while (0)
  switch (BN) {
    case 64: GIC_TC(64); break;
    case 128: GIC_TC(128); break;
    case 144: GIC_TC(144); break;
    case 192: GIC_TC(192); break;
    case 240: GIC_TC(240); break;
    default: GIC_TC(256); break;
  }
#undef GIC_TC
  if (rc == GIC_OK) *handled = true;
  return rc;
}

}  // namespace gic
