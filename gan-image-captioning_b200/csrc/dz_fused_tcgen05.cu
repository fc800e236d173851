// Generator backward through the discriminator's embedding layer and the tempered softmax, in ONE streaming kernel
// (sm_100a only; autograd of src/discriminator.py:40 and src/generator.py:69):
//   dz[r, v]  = T * p[r, v] * (sum_k demb[r, k] W_e[k, v] - dot[r])        dot[r] = <demb[r, :], emb[r, :]> = sum_v p dp
//   db_out[v] (+)= sum_r dz[r, v]                                          (bias gradient of the vocab projection)
// dz is written once, as bf16 (it is only read by the two tensor-core contractions dW_out = dz^T htop, dhtop = dz W_out).
// The dense d(probs)[B*L, V] never exists: its 128 x 64 tile is a K = De (<= 64) tcgen05.mma product held in TMEM.
// The kernel is bound by HBM: it reads p once (TMA, 128-byte-swizzled tiles, three in flight per SM) and writes dz once
// (TMA store of a bf16 staging tile): 6 bytes per element against 4 + 4 + 4 + 4 + 2 + 2 for the separate GEMM,
// elementwise and column-sum kernels it replaces.
// Persistent: every CTA walks a contiguous run of the (row block, column tile) space, so demb (the A operand, 32 KB per
// row block) is loaded when the row block changes, not per tile.
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..17 = two epilogue
// groups of 8 warps that take alternate tiles (thread = one row x 32 columns; two warps per TMEM lane quarter), so one
// group's TMEM read / convert / column-sum chain overlaps the other's.
#include "tcgen05_common.cuh"

namespace gic {
namespace tc {

constexpr int DZ_BN = 64;
constexpr int DZ_STAGES = 3;
constexpr int DZ_EPI_THREADS = 256;
constexpr int DZ_GROUPS = 2;                                // epilogue groups alternating tiles (TMEM buffer = staging tile = group)
constexpr int DZ_THREADS = 64 + DZ_GROUPS * DZ_EPI_THREADS;
constexpr int DZ_A_BYTES = 2 * BM * BK * 4;                 // two k-blocks of the row block: 32 KB
constexpr int DZ_B_BYTES = 2 * DZ_BN * BK * 4;              // two k-blocks of W_e[:, n0:n0+64]: 16 KB
constexpr int DZ_P_BYTES = BM * DZ_BN * 4;                  // p tile: 32 KB = 2 boxes of [128 rows][128 B]
constexpr int DZ_O_BYTES = BM * DZ_BN * 2;                  // bf16 staging tile: 16 KB = [128 rows][128 B]
constexpr int DZ_SMEM = DZ_A_BYTES + DZ_STAGES * (DZ_B_BYTES + DZ_P_BYTES) + 2 * DZ_O_BYTES + DZ_GROUPS * 4 * DZ_BN * 4 + 512 + 1024;

struct DZArgs {
  int M, N, tiles_n, tiles, nkb;
  const float* dot;          // [M]
  float T;
  const float* T_dev;
  float* db_out;             // [N], zeroed / holding the value to accumulate onto
};

__device__ __forceinline__ void dz_bar(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(DZ_EPI_THREADS) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned int bf16_rn(float x) {
  unsigned int u = __float_as_uint(x);
  u += 0x7fffu + ((u >> 16) & 1u);
  return u >> 16;
}

__global__ void __launch_bounds__(DZ_THREADS, 1)
dz_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmO, DZArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + DZ_A_BYTES;
  uint8_t* sP = sB + DZ_STAGES * DZ_B_BYTES;
  uint8_t* sO = sP + DZ_STAGES * DZ_P_BYTES;
  float* s_cs = reinterpret_cast<float*>(sO + 2 * DZ_O_BYTES);          // [group][4][64] column partial sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_cs + DZ_GROUPS * 4 * DZ_BN);
  uint64_t* a_full = bars;                 // [1]
  uint64_t* a_free = bars + 1;             // [1]
  uint64_t* b_full = bars + 2;             // [3]
  uint64_t* b_empty = bars + 5;            // [3]
  // p_full: one barrier per (stage, consuming group).  With three stages and two groups the tiles of a stage alternate
  // between the groups; on a barrier shared by both, a group would see only every other phase, and a parity wait cannot tell
  // "my phase k has completed" from "phase k - 1 has not completed yet" (same parity of the phase in progress).  That is
  // what happened when the p load of tile it - 3 (HBM, the other group's) was still in flight while B / MMA / TMEM of tile
  // it (L2-resident operands) were already through: the group consumed a stage that had not landed, arrived on p_empty a
  // phase early, the producer's next expect_tx hit a barrier whose phase was still open -- an exception or a stalled
  // pipeline once in ~30 000 training steps under the discriminator chain's HBM load (profiles/stress.py).  Per-group
  // barriers make every waiter see every phase of the barrier it waits on.
  uint64_t* p_full = bars + 8;             // [3][2]
  uint64_t* p_empty = bars + 14;           // [3]
  uint64_t* t_full = bars + 17;            // [2]
  uint64_t* t_empty = bars + 19;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long lo = (long long)a.tiles * blockIdx.x / gridDim.x;
  const long long hi = (long long)a.tiles * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    mbar_init(a_full, 1); mbar_init(a_free, 1);
    for (int s = 0; s < DZ_STAGES; ++s) {
      mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1);
      mbar_init(&p_full[2 * s], 1); mbar_init(&p_full[2 * s + 1], 1); mbar_init(&p_empty[s], DZ_EPI_THREADS / 32);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], DZ_EPI_THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int cur_m = -1, a_loads = 0;
      uint32_t it = 0;
      for (long long tile = lo; tile < hi; ++tile, ++it) {
        const int mtile = (int)(tile / a.tiles_n), ntile = (int)(tile % a.tiles_n);
        const int m0 = mtile * BM, n0 = ntile * DZ_BN;
        if (mtile != cur_m) {
          if (a_loads > 0) mbar_wait(a_free, (a_loads - 1) & 1);       // every MMA of the previous row block has retired
          mbar_expect_tx(a_full, DZ_A_BYTES);
          tma_load_2d(sA, &tmA, a_full, 0, m0);
          tma_load_2d(sA + BM * BK * 4, &tmA, a_full, BK, m0);
          cur_m = mtile; ++a_loads;
        }
        const int s = it % DZ_STAGES;
        const uint32_t ph = (it / DZ_STAGES) & 1;
        mbar_wait(&b_empty[s], ph ^ 1);
        mbar_expect_tx(&b_full[s], DZ_B_BYTES);
        uint8_t* sb = sB + s * DZ_B_BYTES;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int j = 0; j < DZ_BN / 32; ++j)
            tma_load_2d(sb + (kb * (DZ_BN / 32) + j) * 4096, &tmB, &b_full[s], n0 + 32 * j, kb * BK);
        mbar_wait(&p_empty[s], ph ^ 1);
        uint64_t* pf = &p_full[2 * s + (it & 1)];                       // the barrier of the group that takes this tile
        mbar_expect_tx(pf, DZ_P_BYTES);
        uint8_t* sp = sP + s * DZ_P_BYTES;
        tma_load_2d(sp, &tmP, pf, n0, m0);
        tma_load_2d(sp + BM * 128, &tmP, pf, n0 + 32, m0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: K <= 64 -> at most 8 instructions per tile =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(0, 1, DZ_BN);
      int cur_m = -1, a_loads = 0;
      uint32_t it = 0;
      for (long long tile = lo; tile < hi; ++tile, ++it) {
        const int mtile = (int)(tile / a.tiles_n);
        if (mtile != cur_m) { mbar_wait(a_full, a_loads & 1); cur_m = mtile; ++a_loads; }
        const int s = it % DZ_STAGES;
        const uint32_t ph = (it / DZ_STAGES) & 1;
        const uint32_t acc = it & 1, accph = (it >> 1) & 1;
        mbar_wait(&t_empty[acc], accph ^ 1);
        mbar_wait(&b_full[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(sA), sb = smem_u32(sB + s * DZ_B_BYTES);
        const uint32_t d_tmem = tmem_base + acc * DZ_BN;
        for (int kb = 0; kb < a.nkb; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_desc(sa + kb * (BM * BK * 4) + k * 32, 16, 1024, 2);
            const uint64_t db = make_desc(sb + kb * ((DZ_BN / 32) * 4096) + k * 1024, 4096, 512, 1);
            umma_tf32(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&b_empty[s]);
        umma_commit(&t_full[acc]);
        const bool last_of_block = (tile + 1 == hi) || ((int)((tile + 1) / a.tiles_n) != mtile);
        if (last_of_block) umma_commit(a_free);
      }
    }
  } else {
    // ===== epilogue: two groups of 8 warps take alternate tiles; thread = (row, 32-column half g) =====
    const int grp = (warp - 2) >> 3;
    const int w8 = (warp - 2) & 7;
    const int q = warp & 3;
    const int g = w8 >> 2;
    const int row = q * 32 + lane;
    const int etid = w8 * 32 + lane;                 // 0..255 inside the group
    const int sw = row & 7;
    const float T = a.T_dev ? __ldg(a.T_dev) : a.T;
    float* cs = s_cs + grp * 4 * DZ_BN;
    uint8_t* stage_o = sO + grp * DZ_O_BYTES;
    int cur_m = -1;
    float dotv = 0.f;
    for (uint32_t it = grp; lo + it < hi; it += DZ_GROUPS) {
      const long long tile = lo + it;
      const int mtile = (int)(tile / a.tiles_n), ntile = (int)(tile % a.tiles_n);
      const int m0 = mtile * BM, n0 = ntile * DZ_BN;
      if (mtile != cur_m) { cur_m = mtile; dotv = (m0 + row < a.M) ? __ldg(a.dot + m0 + row) : 0.f; }
      const int s = it % DZ_STAGES;
      const uint32_t acc = it & 1, accph = (it >> 1) & 1;                   // acc == grp
      uint8_t* prow = sP + s * DZ_P_BYTES + g * (BM * 128) + row * 128;      // this thread's 32 fp32 of the p tile
      uint8_t* orow = stage_o + row * 128;                                  // this row's 64 bf16 of the staging tile
      mbar_wait(&t_full[acc], accph);
      tcgen05_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + acc * DZ_BN + 32 * g + ((uint32_t)(q * 32) << 16), r);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&t_empty[acc]);
      mbar_wait(&p_full[2 * s + grp], (it / (2 * DZ_STAGES)) & 1);      // this group's uses of stage s are 6 tiles apart
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4* pp = reinterpret_cast<float4*>(prow + ((k ^ sw) << 4));
        float4 v = *pp;
        v.x = T * v.x * (__uint_as_float(r[4 * k + 0]) - dotv);
        v.y = T * v.y * (__uint_as_float(r[4 * k + 1]) - dotv);
        v.z = T * v.z * (__uint_as_float(r[4 * k + 2]) - dotv);
        v.w = T * v.w * (__uint_as_float(r[4 * k + 3]) - dotv);
        *pp = v;                                                        // fp32 dz stays for the column sums
        r[4 * k + 0] = bf16_rn(v.x) | (bf16_rn(v.y) << 16);
        r[4 * k + 1] = bf16_rn(v.z) | (bf16_rn(v.w) << 16);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)                                          // 32 bf16 = four 16-byte chunks 4g .. 4g+3
        *reinterpret_cast<uint4*>(orow + (((4 * g + c) ^ sw) << 4)) =
            make_uint4(r[8 * c + 0], r[8 * c + 1], r[8 * c + 4], r[8 * c + 5]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      dz_bar(grp);
      if (etid == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(&tmO), "r"(smem_u32(stage_o)), "r"(n0), "r"(m0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      // column sums: thread = (column e & 63, row quarter e >> 6), fp32 dz from the swizzled p tile; four independent
      // accumulators keep eight loads in flight
      {
        const int col = etid & 63, rq = etid >> 6;
        const uint8_t* base = sP + s * DZ_P_BYTES + (col >> 5) * (BM * 128) + (col & 3) * 4;
        const int chunk = (col & 31) >> 2;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int rr = rq * 32 + i;
          s0 += *reinterpret_cast<const float*>(base + (rr + 0) * 128 + ((chunk ^ ((rr + 0) & 7)) << 4));
          s1 += *reinterpret_cast<const float*>(base + (rr + 1) * 128 + ((chunk ^ ((rr + 1) & 7)) << 4));
          s2 += *reinterpret_cast<const float*>(base + (rr + 2) * 128 + ((chunk ^ ((rr + 2) & 7)) << 4));
          s3 += *reinterpret_cast<const float*>(base + (rr + 3) * 128 + ((chunk ^ ((rr + 3) & 7)) << 4));
        }
        cs[rq * DZ_BN + col] = (s0 + s1) + (s2 + s3);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&p_empty[s]);
      if (etid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile free for this group's next tile
      dz_bar(grp);
      if (etid < DZ_BN) {
        const int n = n0 + etid;
        if (n < a.N) atomicAdd(a.db_out + n, (cs[etid] + cs[DZ_BN + etid]) + (cs[2 * DZ_BN + etid] + cs[3 * DZ_BN + etid]));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
  }
}

}  // namespace tc

// dz (bf16, [M, Vp]) and db_out from demb [M, K], W_e [K, V], p [M, V] (row pitch V), dot [M].  handled = false
// (nothing launched) when the operands do not fit the kernel; the caller then runs the separate kernels.
int dz_fused_tc(const float* demb, int K, const float* W_e, const float* p, const float* dot, float T, const float* T_dev,
                int M, int V, void* dz_bf, int Vp, float* db_out, int accumulate, cudaStream_t stream, bool* handled) {
  using namespace tc;
  *handled = false;
  if (option("GIC_FUSED_DZ_BF16", 1) == 0) return GIC_OK;
  if (M <= 0 || V <= 0 || K <= 0 || K > 64 || (K % 4) || (V % 4) || (Vp % 8)) return GIC_OK;
  if (!aligned16(demb) || !aligned16(W_e) || !aligned16(p) || !aligned16(dz_bf)) return GIC_OK;
  const bool rn = tf32_round_in_tma();
  CUtensorMap ta, tb, tp, to;
  bool ok = make_map(&ta, demb, M, K, K, BK, BM, rn, false) && make_map(&tb, W_e, K, V, V, 32, BK, rn, true) &&
            make_map(&tp, p, M, V, V, 32, BM, false, false) && make_map_bf16(&to, dz_bf, M, V, Vp, DZ_BN, BM);
  if (!ok) return GIC_OK;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(db_out, 0, (size_t)V * sizeof(float), stream);
    if (e != cudaSuccess) { set_error("dz_fused memset: %s", cudaGetErrorString(e)); return GIC_ERR_CUDA; }
  }
  DZArgs a;
  a.M = M; a.N = V; a.tiles_n = cdiv(V, DZ_BN); a.tiles = cdiv(M, BM) * a.tiles_n; a.nkb = cdiv(K, BK);
  a.dot = dot; a.T = T; a.T_dev = T_dev; a.db_out = db_out;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(dz_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DZ_SMEM);
    attr = true;
  }
  const int grid = a.tiles < num_sms() ? a.tiles : num_sms();
  dz_fused_kernel<<<grid, DZ_THREADS, DZ_SMEM, stream>>>(ta, tb, tp, to, a);
  int rc = check_launch("dz_fused_kernel");
  if (rc == GIC_OK) *handled = true;
  return rc;
}

}  // namespace gic
