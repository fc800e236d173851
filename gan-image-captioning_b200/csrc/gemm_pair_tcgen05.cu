// CTA-pair (cta_group::2) bf16 GEMM for the discriminator's [N*R, F] x [F, F] contractions (sm_100a only).
//
// Why: with one CTA per 128 x 256 tile the persistent kernel of gemm_persistent.cu pulls 48 KB of operands per 64-deep
// k-block (A 16 KB + B 32 KB) through an SM that ingests ~44 B/clk (profiles/README.md): 1 090 clk per k-block against
// 861 clk of tensor-core work -- the highway / dx GEMMs are operand-ingest bound.  Two CTAs of a cluster (the two SMs of
// a TPC) share one 256 x 256 tile here: each loads its own 128 rows of A and HALF of the B tile (32 KB per k-block and
// SM), one elected thread of the even CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads both halves of B from
// the two shared memories and accumulates each CTA's 128 rows in that CTA's own TMEM.
//
// Protocol (barriers live at the same shared-memory offsets in both CTAs):
//   full[s]        leader only.  One arrival (the leader's expect_tx of BOTH CTAs' bytes); the TMA loads of both CTAs
//                  complete on it (cp.async.bulk.tensor .cta_group::2, peer bit of the barrier address cleared).
//   empty[s]       each CTA.  tcgen05.commit .cta_group::2 .multicast::cluster (mask 0b11) after the k-block's MMAs.
//   tmem_full[a]   each CTA.  Multicast commit after the tile's last MMA: both epilogues start.
//   tmem_empty[a]  leader only, 8 arrivals: the four epilogue warps of each CTA (the peer's arrive through mapa).
// Scope: A K-major bf16, B K-major ([N, K]) or MN-major ([K, N]) bf16, fp32 output through TMA stores (beta = 0), optional
// bias; whole tiles only (no stream-K).  Everything else stays on gemm_p_kernel.
#include "tcgen05_common.cuh"

namespace gic {
namespace tc {

constexpr int PBN = 256;                                  // pair-tile width
constexpr int P_A_BYTES = BM * 128;                       // 128 rows x 64 bf16
constexpr int P_B_BYTES = (PBN / 2) * 128;                // this CTA's half of the B tile
constexpr int P_STAGE = P_A_BYTES + P_B_BYTES;            // 32 KB
constexpr int P_STAGES = 5;
constexpr int P_STG_WARP = 4096;                          // per-warp 32 x 32 fp32 TMA-store tile (128B-swizzled)
constexpr int P_BIAS = PBN * 4;                           // the tile's slice of the bias, staged once per tile
constexpr int P_SMEM = P_STAGES * P_STAGE + 4 * P_STG_WARP + 512 + P_BIAS + 1024;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of a CTA pair: the bytes are counted on the LEADER's barrier (peer bit of the shared::cluster address cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// instruction descriptor of the pair MMA: M = 256
__host__ __device__ constexpr uint32_t make_idesc_bf16_pair(int a_mn, int b_mn, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

struct PairArgs {
  int M, N, K, tiles_n, tiles;
  float alpha;
  const float* bias;       // [N] or null
};

template <bool B_MN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, PairArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stg_all = smem + P_STAGES * P_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(stg_all + 4 * P_STG_WARP);
  uint64_t* empty = full + P_STAGES;
  uint64_t* tmem_full = empty + P_STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(stg_all + 4 * P_STG_WARP + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int nkb = (a.K + 63) / 64;
  constexpr uint32_t TMEM_COLS = 512;          // two 256-column accumulators

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();                              // both CTAs' barriers exist before anything remote touches them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();                              // both allocations done before the first pair MMA
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of A, own half of the B tile =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = cid; tile < a.tiles; tile += ncl) {
        const int m0 = (tile / a.tiles_n) * 256 + (int)rank * BM;
        const int n0 = (tile % a.tiles_n) * PBN + (int)rank * (PBN / 2);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % P_STAGES;
          const uint32_t ph = (it / P_STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * P_STAGE;
          uint8_t* sb = sa + P_A_BYTES;
          if (rank == 0) mbar_expect_tx(&full[s], 2 * P_STAGE);
          const int k0 = kb * 64;
          tma_load_2d_pair(sa, &tmA, &full[s], k0, m0);
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < PBN / 2 / 64; ++j) tma_load_2d_pair(sb + j * 8192, &tmB, &full[s], n0 + 64 * j, k0);
          } else {
            tma_load_2d_pair(sb, &tmB, &full[s], k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_pair(0, B_MN ? 1 : 0, PBN);
      uint32_t it = 0, seg = 0;
      for (int tile = cid; tile < a.tiles; tile += ncl, ++seg) {
        const uint32_t acc = seg & 1, accph = (seg >> 1) & 1;
        mbar_wait(&tmem_empty[acc], accph ^ 1);          // both epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PBN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % P_STAGES;
          const uint32_t ph = (it / P_STAGES) & 1;
          mbar_wait(&full[s], ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + s * P_STAGE);
          const uint32_t sb = sa + P_A_BYTES;
          // the last k-block issues only the MMAs (16 elements of K each) that hold data: K = 900 leaves 4 of 64
          // elements in block 15, TMA zero-fills the rest and three of its four MMAs would multiply zeros
          const int kmax = min(4, (a.K - kb * 64 + 15) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k >= kmax) break;
            const uint64_t da = make_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t db = B_MN ? make_desc(sb + k * 2048, 8192, 1024, 2) : make_desc(sb + k * 32, 16, 1024, 2);
            umma_bf16_pair(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty[s]);                   // both CTAs may refill the stage
        }
        umma_commit_pair(&tmem_full[acc]);
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): own 128 rows x 256 columns, TMA stores of 32 x 32 boxes =====
    const int q = warp & 3;
    uint8_t* stg = stg_all + q * P_STG_WARP;
    uint32_t seg = 0;
    for (int tile = cid; tile < a.tiles; tile += ncl, ++seg) {
      const uint32_t acc = seg & 1, accph = (seg >> 1) & 1;
      const int m_base = (tile / a.tiles_n) * 256 + (int)rank * BM + q * 32;
      const int n0 = (tile % a.tiles_n) * PBN;
      if (a.bias != nullptr) {
        // bias slice -> shared memory while the tile's MMAs still run (not a dependent global load per 32-column block)
        asm volatile("bar.sync 1, 128;" ::: "memory");        // the previous tile's readers are done
        for (int j = (int)threadIdx.x - 64; j < PBN; j += 128) bias_s[j] = (n0 + j < a.N) ? a.bias[n0 + j] : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(&tmem_full[acc], accph);
      tcgen05_fence_after();
      const uint32_t t_addr = tmem_base + acc * PBN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < PBN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_addr + c0, r);
        if (c0 + 32 >= PBN) {                              // last chunk read: hand the accumulator back to the leader
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
        }
        if (n0 + c0 >= a.N || m_base >= a.M) continue;     // warp-uniform
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile free again
        __syncwarp();
        uint8_t* srow = stg + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 v;
          v.x = a.alpha * __uint_as_float(r[4 * c + 0]); v.y = a.alpha * __uint_as_float(r[4 * c + 1]);
          v.z = a.alpha * __uint_as_float(r[4 * c + 2]); v.w = a.alpha * __uint_as_float(r[4 * c + 3]);
          if (a.bias != nullptr) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * c);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          }
          *reinterpret_cast<float4*>(srow + ((c ^ (lane & 7)) << 4)) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&tmC), "r"(smem_u32(stg)), "r"(n0 + c0), "r"(m_base) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();                              // the peer has finished with this CTA's shared memory / barriers
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

}  // namespace tc

// C[M,N] (fp32, TMA-stored) = alpha * A[M,K] (bf16, K-major) * op(B) + bias, op(B) = B[N,K]^T (b_mn = false) or B[K,N]
// (b_mn = true).  handled = false when the shape is not worth / not fit for a CTA pair.
int gemm_tc_bf16_pair(bool b_mn, int M, int N, int K, float alpha, const void* A, int lda, const void* B, int ldb, float* C,
                      int ldc, const float* bias, cudaStream_t stream, bool* handled) {
  using namespace tc;
  *handled = false;
  if (option("GIC_GEMM_2CTA", 1) == 0) return GIC_OK;
  if (M < 512 || N < 192 || K < 64) return GIC_OK;
  if (!aligned16(A) || !aligned16(B) || !aligned16(C) || (lda % 8) || (ldb % 8) || (ldc % 4) || (bias && !aligned16(bias))) return GIC_OK;
  const int tiles_n = cdiv(N, PBN), tiles = tiles_n * cdiv(M, 256);
  const int pairs = num_sms() / 2;
  if (tiles < pairs) return GIC_OK;             // not enough whole pair-tiles to fill the machine: stream-K kernel's job
  CUtensorMap ta, tb, tcm;
  bool ok = make_map_bf16(&ta, A, M, K, lda, 64, BM);
  if (ok) ok = b_mn ? make_map_bf16(&tb, B, K, N, ldb, 64, 64) : make_map_bf16(&tb, B, N, K, ldb, 64, PBN / 2);
  if (ok) ok = make_map(&tcm, C, M, N, ldc, 32, 32, false, false);
  if (!ok) return GIC_OK;
  PairArgs pa;
  pa.M = M; pa.N = N; pa.K = K; pa.tiles_n = tiles_n; pa.tiles = tiles; pa.alpha = alpha; pa.bias = bias;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gemm_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM);
    cudaFuncSetAttribute(gemm_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM);
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * min(pairs, tiles), 1, 1);
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = P_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = b_mn ? cudaLaunchKernelEx(&cfg, gemm_pair_kernel<true>, ta, tb, tcm, pa)
                       : cudaLaunchKernelEx(&cfg, gemm_pair_kernel<false>, ta, tb, tcm, pa);
  if (e != cudaSuccess) { set_error("gemm_pair_kernel launch: %s", cudaGetErrorString(e)); return GIC_ERR_CUDA; }
  int rc = check_launch("gemm_pair_kernel");
  if (rc == GIC_OK) *handled = true;
  return rc;
}

}  // namespace gic
