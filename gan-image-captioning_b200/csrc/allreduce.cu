// Data-parallel gradient exchange over NVLink 5 / NVSwitch peer memory (the reference is single-GPU,
// src/args.py:213-216,276).  ONE kernel per call does what the step needs between the backward pass and the optimizer:
//   all-reduce(sum) of a flat fp32 gradient buffer, in place, and -- fused into the same pass -- the global square norm of
//   the REDUCED gradient that clip_grad_norm_ (src/training.py:198) needs next.
// There is no staging copy and no library collective: every rank's gradient buffer lives in a symmetric allocation that all
// ranks of the node map (CUDA IPC), and the kernel reads and writes peer memory directly:
//   start barrier   per CTA c: "my gradients are final" -> every peer's ready[c][me]; wait for every peer's flag
//   reduce-scatter  rank r owns slice r of the buffer; CTA c owns chunk c of that slice: for each 16-byte element, load it
//   + all-gather    from all W ranks (peer loads over NVLink), add in RANK ORDER (every rank computes bit-identical sums),
//                   accumulate its square, and store the sum into all W buffers (peer stores)
//   end barrier     per CTA c: "chunk c of my slice is everywhere" -> every peer's done[c][me]; wait for every peer's flag:
//                   then chunk c of EVERY slice has landed here, and every peer has finished reading this rank's buffer
//   square norm     each CTA pushes its partial to sq[me][c] on every rank; the last CTA of a rank to finish waits for all
//                   flags and adds the W x CTAS partials in a fixed order: the same number on every rank
// Flags carry a per-(channel, CTA) epoch that the kernel itself increments, so a captured CUDA graph can be replayed.
// Waits are bounded (2 s): a lost peer sets comm->error instead of hanging the GPU.
// Traffic per rank for n floats over W ranks: reads 4 n (W-1)/W from peers, writes the same; at c2 (58 MB, W = 8) that is
// 51 MB each way: ~70 us at NVLink rate, against 0.42 ms exposed by the NCCL call it replaces (SCALE_r01.json).
#include <stdlib.h>
#include <string.h>

#include "gic_internal.cuh"

namespace gic {

constexpr int AR_MAX_WORLD = 8;              // one NVSwitch box
// CTAs per rank and call.  Footprint matters more than width: the exchange of the early buckets runs UNDER the backward pass,
// next to persistent one-CTA-per-SM kernels.  The first version (48 CTAs x 512 threads x 128 registers = a whole register
// file per CTA) could only be scheduled on an EMPTY SM and kept everything else off it: a 20 MB bucket took 150 - 200 us
// under the BPTT kernel (45 us alone) and the chains beside it lost ~100 us.  256 threads x <= 64 registers (16 K registers,
// 1 KB shared memory) co-resides with every kernel of the step; 64 x 256 x 8 x 16 B = 2 MB of peer loads in flight still
// cover NVLink's latency-bandwidth product (~2 us x 750 GB/s).
constexpr int AR_CTAS = 64;
constexpr int AR_THREADS = 256;
constexpr int AR_CTAS_PER_SM = 4;
constexpr int AR_CHANNELS = 4;         // independent flag sets: calls on different streams may overlap in time

struct ArFlags {                       // lives behind the data in every rank's symmetric allocation
  unsigned int ready[AR_CHANNELS][AR_CTAS][AR_MAX_WORLD];     // written by peers
  unsigned int done[AR_CHANNELS][AR_CTAS][AR_MAX_WORLD];      // written by peers
  float sq[AR_CHANNELS][AR_MAX_WORLD][AR_CTAS];               // written by peers: partial square norms of rank r's slice
  unsigned int epoch[AR_CHANNELS][AR_CTAS];                   // local
  unsigned int ticket[AR_CHANNELS];                           // local: CTAs of this rank that finished the current call
  unsigned int error;                                         // local: a bounded wait expired
};

}  // namespace gic

struct gic_comm {
  int rank, world;
  size_t data_bytes;                   // bytes of the gradient region (16-byte multiple)
  char* base;                          // this rank's allocation: [data | ArFlags]
  char* peer[gic::AR_MAX_WORLD];       // every rank's allocation as mapped here (peer[rank] == base)
  bool opened[gic::AR_MAX_WORLD];      // mapped through cudaIpcOpenMemHandle (to be closed)
  bool local_group;                    // in-process group (one GPU, tests): peers are plain allocations of this process
};

namespace gic {

struct ArRank {                        // what one rank needs inside the kernel
  float* data[AR_MAX_WORLD];           // the buffer being reduced, in every rank's memory
  ArFlags* flags[AR_MAX_WORLD];
  float* sqnorm_out;                   // local, may be null
};
struct ArArgs {
  ArRank r[AR_MAX_WORLD];              // r[0] only for a multi-process call; r[0..world) for an in-process group (gridDim.y)
  int rank_base, world, channel;
  size_t n4;                           // float4 elements of the buffer
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer memory: never from a stale cache line.  No "memory" clobber: these loads only have to stay behind the start barrier
// (a __syncthreads) and ahead of the stores that consume them, and they must be free to overlap one another.
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ar_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait until *flag >= e; false (and the error word set) after 2 s
__device__ __forceinline__ bool ar_wait(const unsigned int* flag, unsigned int e, unsigned int* err) {
  const unsigned long long t0 = ar_now();
  while ((int)(ld_acquire_sys(flag) - e) < 0) {
    __nanosleep(64);
    if (ar_now() - t0 > 2000000000ull) { atomicExch(err, 1u); return false; }
  }
  return true;
}

// U elements per thread and trip: U x W loads in flight, sums in rank order, U x W stores
template <int U>
__device__ __forceinline__ float ar_body(const ArRank& me, int W, size_t i0, size_t s1, size_t stride) {
  float sq = 0.f;
  for (size_t i = i0; i < s1; i += U * stride) {
    float4 v[U][AR_MAX_WORLD];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t iu = i + u * stride;
#pragma unroll
      for (int p = 0; p < AR_MAX_WORLD; ++p)
        if (p < W && iu < s1) v[u][p] = ld_peer_f4(reinterpret_cast<const float4*>(me.data[p]) + iu);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t iu = i + u * stride;
      if (iu < s1) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < AR_MAX_WORLD; ++p)                   // rank order: the same sum on every rank
          if (p < W) { acc.x += v[u][p].x; acc.y += v[u][p].y; acc.z += v[u][p].z; acc.w += v[u][p].w; }
        sq = fmaf(acc.x, acc.x, fmaf(acc.y, acc.y, fmaf(acc.z, acc.z, fmaf(acc.w, acc.w, sq))));
#pragma unroll
        for (int p = 0; p < AR_MAX_WORLD; ++p)
          if (p < W) reinterpret_cast<float4*>(me.data[p])[iu] = acc;
      }
    }
  }
  return sq;
}

__global__ void __launch_bounds__(AR_THREADS, AR_CTAS_PER_SM) allreduce_p2p_kernel(const __grid_constant__ ArArgs a) {
  const int W = a.world, ch = a.channel, c = blockIdx.x, tid = threadIdx.x;
  const int ly = blockIdx.y;                                   // 0 for a multi-process call
  const ArRank& me = a.r[ly];
  const int rank = a.rank_base + ly;
  ArFlags* mine = me.flags[rank];
  __shared__ unsigned int s_epoch;
  __shared__ float s_red[AR_THREADS / 32];
  if (tid == 0) s_epoch = ++mine->epoch[ch][c];
  __syncthreads();
  const unsigned int e = s_epoch;

  // ---- start barrier: this kernel runs after the backward pass in its stream, so this rank's gradients are final
  if (tid < W) st_release_sys(&me.flags[tid]->ready[ch][c][rank], e);
  if (tid < W) ar_wait(&mine->ready[ch][c][tid], e, &mine->error);
  __syncthreads();

  // ---- reduce-scatter + all-gather of chunk c of slice `rank`
  const size_t per = (a.n4 + W - 1) / W;                       // float4 per slice
  const size_t s0 = (size_t)rank * per;
  const size_t s1 = (s0 + per < a.n4) ? s0 + per : a.n4;
  float sq = 0.f;
  const size_t stride = (size_t)AR_CTAS * AR_THREADS;
  const size_t i0 = s0 + (size_t)c * AR_THREADS + tid;
  // NVLink's latency-bandwidth product (~2 us x 750 GB/s) needs ~1.5 MB of loads in flight per rank, and a trip also has to
  // hide its own W stores: every thread keeps 8 16-byte peer loads outstanding -- U elements x W ranks, all issued before
  // the first one is consumed (8 x 4 data registers: the kernel stays within 64 registers per thread)
  if (W <= 2) sq = ar_body<4>(me, W, i0, s1, stride);
  else if (W <= 4) sq = ar_body<2>(me, W, i0, s1, stride);
  else sq = ar_body<1>(me, W, i0, s1, stride);
  // partial square norm of this chunk: block reduction in a fixed order, pushed to every rank
  sq = warp_sum(sq);
  if ((tid & 31) == 0) s_red[tid >> 5] = sq;
  __threadfence_system();                                      // this thread's peer stores before the flags below
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < AR_THREADS / 32; ++w) t += s_red[w];
    for (int p = 0; p < W; ++p) me.flags[p]->sq[ch][rank][c] = t;
    __threadfence_system();
  }
  __syncthreads();

  // ---- end barrier: chunk c of my slice is in every buffer; wait until chunk c of every slice is in mine
  if (tid < W) st_release_sys(&me.flags[tid]->done[ch][c][rank], e);
  if (tid < W) ar_wait(&mine->done[ch][c][tid], e, &mine->error);
  __syncthreads();

  // ---- global square norm by the last CTA of this rank to get here
  if (me.sqnorm_out != nullptr) {
    __shared__ unsigned int s_last;
    if (tid == 0) s_last = (atomicAdd(&mine->ticket[ch], 1u) == (unsigned int)AR_CTAS - 1u) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      for (int k = tid; k < AR_CTAS * W; k += AR_THREADS) ar_wait(&mine->done[ch][k / W][k % W], e, &mine->error);
      __syncthreads();
      if (tid == 0) {
        float t = 0.f;
        for (int r = 0; r < W; ++r)
          for (int k = 0; k < AR_CTAS; ++k) t += *reinterpret_cast<volatile float*>(&mine->sq[ch][r][k]);
        atomicAdd(me.sqnorm_out, t);                           // accumulates: several buffers / overlapping calls may share one norm
        mine->ticket[ch] = 0u;
      }
    }
  } else if (tid == 0) {
    if (atomicAdd(&mine->ticket[ch], 1u) == (unsigned int)AR_CTAS - 1u) mine->ticket[ch] = 0u;
  }
}

static size_t ar_data_bytes(size_t b) { return (b + 255) & ~(size_t)255; }

static int ar_fill(ArRank* r, gic_comm* comm, float* buf, float* sqnorm_out) {
  const ptrdiff_t off = reinterpret_cast<char*>(buf) - comm->base;
  for (int p = 0; p < comm->world; ++p) {
    r->data[p] = reinterpret_cast<float*>(comm->peer[p] + off);
    r->flags[p] = reinterpret_cast<ArFlags*>(comm->peer[p] + ar_data_bytes(comm->data_bytes));
  }
  r->sqnorm_out = sqnorm_out;
  return GIC_OK;
}

static int ar_check(gic_comm* comm, const float* buf, size_t n, int channel) {
  GIC_REQUIRE(comm && buf, GIC_ERR_NULL, "allreduce: NULL pointer");
  GIC_REQUIRE(channel >= 0 && channel < AR_CHANNELS, GIC_ERR_SHAPE, "allreduce: channel %d out of range", channel);
  const char* b = reinterpret_cast<const char*>(buf);
  GIC_REQUIRE(b >= comm->base && b + n * sizeof(float) <= comm->base + comm->data_bytes, GIC_ERR_SHAPE,
              "allreduce: the buffer must lie inside the communicator's symmetric allocation (gic_comm_buffer)");
  GIC_REQUIRE(aligned16(buf) && (n % 4) == 0, GIC_ERR_SHAPE, "allreduce: buffer must be 16-byte aligned with n %% 4 == 0");
  return GIC_OK;
}

}  // namespace gic

using namespace gic;

extern "C" {

gic_comm_t* gic_comm_create(int rank, int world, size_t data_bytes) {
  if (world < 1 || world > AR_MAX_WORLD || rank < 0 || rank >= world || data_bytes == 0) {
    set_error("gic_comm_create: bad rank / world / size (world <= %d)", AR_MAX_WORLD);
    return nullptr;
  }
  gic_comm* c = static_cast<gic_comm*>(calloc(1, sizeof(gic_comm)));
  if (!c) return nullptr;
  c->rank = rank; c->world = world; c->data_bytes = (data_bytes + 15) & ~(size_t)15;
  const size_t total = ar_data_bytes(c->data_bytes) + sizeof(ArFlags);
  // the one allocation this library makes: peers must be able to map it, which a caller-owned (caching-allocator) buffer
  // does not allow
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->base), total);
  if (e != cudaSuccess) { set_error("gic_comm_create: cudaMalloc(%zu): %s", total, cudaGetErrorString(e)); free(c); return nullptr; }
  cudaMemset(c->base, 0, total);
  cudaDeviceSynchronize();
  c->peer[rank] = c->base;
  return c;
}

size_t gic_comm_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int gic_comm_ipc_handle(gic_comm_t* c, void* out) {
  GIC_REQUIRE(c && out, GIC_ERR_NULL, "gic_comm_ipc_handle: NULL pointer");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, c->base);
  if (e != cudaSuccess) { set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return GIC_ERR_CUDA; }
  memcpy(out, &h, sizeof(h));
  return GIC_OK;
}

int gic_comm_open(gic_comm_t* c, const void* handles) {
  GIC_REQUIRE(c && handles, GIC_ERR_NULL, "gic_comm_open: NULL pointer");
  const char* hs = static_cast<const char*>(handles);
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + (size_t)p * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e)); return GIC_ERR_CUDA; }
    c->peer[p] = static_cast<char*>(ptr);
    c->opened[p] = true;
  }
  return GIC_OK;
}

int gic_comm_local_group(gic_comm_t* const* comms, int world) {
  GIC_REQUIRE(comms && world >= 1 && world <= AR_MAX_WORLD, GIC_ERR_SHAPE, "gic_comm_local_group: bad group");
  for (int r = 0; r < world; ++r) {
    GIC_REQUIRE(comms[r] && comms[r]->world == world && comms[r]->rank == r && comms[r]->data_bytes == comms[0]->data_bytes,
                GIC_ERR_SHAPE, "gic_comm_local_group: communicator %d does not match the group", r);
    for (int p = 0; p < world; ++p) comms[r]->peer[p] = comms[p]->base;
    comms[r]->local_group = true;
  }
  return GIC_OK;
}

void* gic_comm_buffer(gic_comm_t* c) { return c ? c->base : nullptr; }
size_t gic_comm_buffer_bytes(gic_comm_t* c) { return c ? c->data_bytes : 0; }

int gic_comm_error(gic_comm_t* c) {
  if (!c) return 1;
  unsigned int v = 0;
  const ArFlags* f = reinterpret_cast<const ArFlags*>(c->base + ar_data_bytes(c->data_bytes));
  cudaMemcpy(&v, &f->error, sizeof(v), cudaMemcpyDeviceToHost);
  return (int)v;
}

void gic_comm_destroy(gic_comm_t* c) {
  if (!c) return;
  for (int p = 0; p < c->world; ++p)
    if (c->opened[p]) cudaIpcCloseMemHandle(c->peer[p]);
  cudaFree(c->base);
  free(c);
}

int gic_allreduce(float* buf, size_t n, gic_comm_t* comm, int channel, float* sqnorm, gic_stream_t stream) {
  GIC_TRY(ar_check(comm, buf, n, channel));
  GIC_REQUIRE(!comm->local_group, GIC_ERR_UNSUPPORTED, "gic_allreduce: in-process group: use gic_allreduce_local_group");
  if (n == 0) return GIC_OK;
  ArArgs a;
  memset(&a, 0, sizeof(a));
  ar_fill(&a.r[0], comm, buf, sqnorm);
  a.rank_base = comm->rank; a.world = comm->world; a.channel = channel; a.n4 = n / 4;
  allreduce_p2p_kernel<<<dim3(AR_CTAS, 1), AR_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("allreduce_p2p_kernel");
}

int gic_allreduce_local_group(gic_comm_t* const* comms, float* const* bufs, float* const* sqnorms, size_t n, int world, int channel,
                              gic_stream_t stream) {
  GIC_REQUIRE(comms && bufs && world >= 1 && world <= AR_MAX_WORLD, GIC_ERR_SHAPE, "allreduce_local_group: bad group");
  GIC_REQUIRE(world * AR_CTAS <= num_sms() * AR_CTAS_PER_SM, GIC_ERR_UNSUPPORTED, "allreduce_local_group: %d ranks x %d CTAs must be co-resident", world, AR_CTAS);
  ArArgs a;
  memset(&a, 0, sizeof(a));
  for (int r = 0; r < world; ++r) {
    GIC_TRY(ar_check(comms[r], bufs[r], n, channel));
    GIC_REQUIRE(comms[r]->local_group, GIC_ERR_UNSUPPORTED, "allreduce_local_group: call gic_comm_local_group first");
    ar_fill(&a.r[r], comms[r], bufs[r], sqnorms ? sqnorms[r] : nullptr);
  }
  if (n == 0) return GIC_OK;
  a.rank_base = 0; a.world = world; a.channel = channel; a.n4 = n / 4;
  allreduce_p2p_kernel<<<dim3(AR_CTAS, world), AR_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("allreduce_p2p_kernel");
}

}  // extern "C"
