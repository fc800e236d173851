// fp32 CUDA-core GEMM (FFMA) + column reductions.  This is the exact-fp32 parity path and the
// fallback for shapes the tcgen05 kernels do not take; the tensor-core path lives in
// gemm_tcgen05.cu.  Replaces the reference's nn.Linear / cuBLAS calls
// (src/generator.py:61,64,68; src/discriminator.py:40,53,58,60) and their autograd backward.
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

#include "gic_internal.cuh"

namespace gic {

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static unsigned long long g_launches = 0;
unsigned long long launch_count() { return g_launches; }

// per-kernel launch counters (gic_kernel_launches): the tests assert that a fused kernel actually ran instead of its
// fallback.  `what` is always a string literal, so the table is keyed by name and stays tiny.
struct KernelCount { const char* name; unsigned long long n; };
static KernelCount g_kcounts[96];
static int g_nk = 0;
static void count_kernel(const char* what) {
  for (int i = 0; i < g_nk; ++i)
    if (g_kcounts[i].name == what || strcmp(g_kcounts[i].name, what) == 0) { ++g_kcounts[i].n; return; }
  if (g_nk < 96) { g_kcounts[g_nk].name = what; g_kcounts[g_nk].n = 1; ++g_nk; }
}
unsigned long long kernel_launches(const char* name) {
  unsigned long long n = 0;
  if (!name) return launch_count();
  for (int i = 0; i < g_nk; ++i)
    if (strcmp(g_kcounts[i].name, name) == 0) n += g_kcounts[i].n;
  return n;
}
int kernel_names(char* buf, int cap) {
  int used = 0;
  for (int i = 0; i < g_nk; ++i) {
    const int len = (int)strlen(g_kcounts[i].name);
    if (used + len + 2 > cap) break;
    memcpy(buf + used, g_kcounts[i].name, len); used += len; buf[used++] = '\n';
  }
  if (cap > 0) buf[used < cap ? used : cap - 1] = 0;
  return g_nk;
}

int check_launch(const char* what) {
  ++g_launches;
  count_kernel(what);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return GIC_ERR_CUDA;
  }
  return GIC_OK;
}

static thread_local Ctx tl_default_ctx;
static thread_local Ctx* tl_ctx = nullptr;
Ctx& ctx() { return tl_ctx ? *tl_ctx : tl_default_ctx; }
Ctx* ctx_set_current(Ctx* c) { Ctx* prev = tl_ctx; tl_ctx = c; return prev; }

static Ctx::Opt* option_find(Ctx& c, const char* name) {
  for (int i = 0; i < c.n_opts; ++i)
    if (strcmp(c.opts[i].name, name) == 0) return &c.opts[i];
  return nullptr;
}
static Ctx::Opt* option_slot(Ctx& c, const char* name) {
  if (Ctx::Opt* o = option_find(c, name)) return o;
  if (c.n_opts >= Ctx::MAX_OPTS || strlen(name) >= sizeof(c.opts[0].name)) return nullptr;
  Ctx::Opt* o = &c.opts[c.n_opts++];
  strcpy(o->name, name);
  o->has = false; o->value = 0;
  return o;
}
static Ctx::Opt* option_lookup(const char* name) {
  Ctx& c = ctx();
  if (Ctx::Opt* o = option_find(c, name)) return o;
  Ctx::Opt* o = option_slot(c, name);
  if (o) {                                         // first lookup in this context: the environment, once
    const char* e = getenv(name);
    if (e && *e) { o->has = true; o->value = atoi(e); }
  }
  return o;
}
int option(const char* name, int dflt) {
  const Ctx::Opt* o = option_lookup(name);
  return (o && o->has) ? o->value : dflt;
}
bool option_is_set(const char* name) {
  const Ctx::Opt* o = option_lookup(name);
  return o && o->has;
}
void option_set(const char* name, int value) {
  if (Ctx::Opt* o = option_slot(ctx(), name)) { o->has = true; o->value = value; }
}
void option_clear(const char* name) {
  Ctx& c = ctx();
  if (Ctx::Opt* o = option_find(c, name)) { *o = c.opts[c.n_opts - 1]; --c.n_opts; }
}

unsigned long long* trap_slot() {
  static unsigned long long* p = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = nullptr;
    if (cudaHostAlloc(&h, 4096, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      memset(h, 0, 4096);
      p = static_cast<unsigned long long*>(h);
    } else {
      cudaGetLastError();
    }
  }
  return p;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

bool pdl_enabled() {
  return option("GIC_PDL", 1) != 0;
}

// ---------------------------------------------------------------------------------------
// tile loaders.  smem tile layout is [BK][BMN + PAD] (k-major) for both operands.
//   KCONTIG = true : global memory is contiguous along k   (A not transposed / B transposed)
//   KCONTIG = false: global memory is contiguous along m|n (A transposed / B not transposed)
// ---------------------------------------------------------------------------------------
constexpr int PAD = 4;

template <bool KCONTIG, int BMN, int BK, int NT>
__device__ __forceinline__ void load_tile(float (*dst)[BMN + PAD], const float* __restrict__ src,
                                          int ld, int mn0, int k0, int MN, int K, bool vec) {
  if (KCONTIG) {
    constexpr int KV = BK / 4;             // float4 per row
    constexpr int TOT = BMN * KV;
#pragma unroll
    for (int v = threadIdx.x; v < TOT; v += NT) {
      const int row = v / KV, kq = (v % KV) * 4;
      const int gm = mn0 + row, gk = k0 + kq;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gm < MN) {
        const float* p = src + (size_t)gm * ld + gk;
        if (vec && gk + 3 < K) {
          x = *reinterpret_cast<const float4*>(p);
        } else {
          if (gk + 0 < K) x.x = p[0];
          if (gk + 1 < K) x.y = p[1];
          if (gk + 2 < K) x.z = p[2];
          if (gk + 3 < K) x.w = p[3];
        }
      }
      dst[kq + 0][row] = x.x;
      dst[kq + 1][row] = x.y;
      dst[kq + 2][row] = x.z;
      dst[kq + 3][row] = x.w;
    }
  } else {
    constexpr int MV = BMN / 4;            // float4 per k-row
    constexpr int TOT = BK * MV;
#pragma unroll
    for (int v = threadIdx.x; v < TOT; v += NT) {
      const int k = v / MV, mq = (v % MV) * 4;
      const int gk = k0 + k, gm = mn0 + mq;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gk < K) {
        const float* p = src + (size_t)gk * ld + gm;
        if (vec && gm + 3 < MN) {
          x = *reinterpret_cast<const float4*>(p);
        } else {
          if (gm + 0 < MN) x.x = p[0];
          if (gm + 1 < MN) x.y = p[1];
          if (gm + 2 < MN) x.z = p[2];
          if (gm + 3 < MN) x.w = p[3];
        }
      }
      *reinterpret_cast<float4*>(&dst[k][mq]) = x;
    }
  }
}

// Each thread computes a (2*TH) x (2*TH) micro-tile split in two halves along both dims
// (rows {ty*TH.., BM/2 + ty*TH..}, cols {tx*TH.., BN/2 + tx*TH..}) so that the float4 smem
// reads of a warp hit distinct banks.
template <int BM, int BN, int BK, int TH, bool TA, bool TB>
__global__ void __launch_bounds__((BM / (2 * TH)) * (BN / (2 * TH)), (BM >= 128 ? 2 : 4))
sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
             const float* __restrict__ B, int ldb, float beta, float* __restrict__ C, int ldc,
             const float* __restrict__ bias, bool vecA, bool vecB, bool vecC) {
  constexpr int TX = BN / (2 * TH), TY = BM / (2 * TH), NT = TX * TY;
  static_assert(TH == 4, "micro tile is float4 based");
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[2 * TH][2 * TH];
#pragma unroll
  for (int i = 0; i < 2 * TH; ++i)
#pragma unroll
    for (int j = 0; j < 2 * TH; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    load_tile<!TA, BM, BK, NT>(As, A, lda, m0, k0, M, K, vecA);
    load_tile<TB, BN, BK, NT>(Bs, B, ldb, n0, k0, N, K, vecB);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TH]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][BM / 2 + ty * TH]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * TH]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][BN / 2 + tx * TH]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 2 * TH; ++i) {
    const int m = m0 + (i < TH ? ty * TH + i : BM / 2 + ty * TH + (i - TH));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + (jh == 0 ? tx * TH : BN / 2 + tx * TH);
      float* cp = C + (size_t)m * ldc + n;
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = alpha * acc[i][jh * TH + j];
      if (vecC && n + 3 < N) {
        if (bias) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + n);
          r[0] += bb.x; r[1] += bb.y; r[2] += bb.z; r[3] += bb.w;
        }
        if (beta != 0.f) {
          const float4 cc = *reinterpret_cast<const float4*>(cp);
          r[0] += beta * cc.x; r[1] += beta * cc.y; r[2] += beta * cc.z; r[3] += beta * cc.w;
        }
        *reinterpret_cast<float4*>(cp) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j < N) {
            float v = r[j];
            if (bias) v += bias[n + j];
            if (beta != 0.f) v += beta * cp[j];
            cp[j] = v;
          }
        }
      }
    }
  }
}

template <int BM, int BN, bool TA, bool TB>
static void launch_sgemm(int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                         int ldb, float beta, float* C, int ldc, const float* bias, bool vA,
                         bool vB, bool vC, cudaStream_t s) {
  constexpr int BK = 16, TH = 4;
  dim3 grid(cdiv(N, BN), cdiv(M, BM));
  dim3 block((BM / (2 * TH)) * (BN / (2 * TH)));
  sgemm_kernel<BM, BN, BK, TH, TA, TB><<<grid, block, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C,
                                                            ldc, bias, vA, vB, vC);
}

int gemm_f32(bool tA, bool tB, int M, int N, int K, float alpha, const float* A, int lda,
             const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
             cudaStream_t s) {
  GIC_REQUIRE(M >= 0 && N >= 0 && K >= 0, GIC_ERR_SHAPE, "gemm_f32: negative dimension");
  if (M == 0 || N == 0) return GIC_OK;
  GIC_REQUIRE(A && B && C, GIC_ERR_NULL, "gemm_f32: NULL operand");
  const bool vA = aligned16(A) && (lda % 4 == 0);
  const bool vB = aligned16(B) && (ldb % 4 == 0);
  const bool vC = aligned16(C) && (ldc % 4 == 0) && (!bias || aligned16(bias));
  const long long big = (long long)cdiv(M, 128) * cdiv(N, 128);
  const bool use128 = big >= (num_sms() * 3) / 4;
#define GIC_SGEMM(TA_, TB_)                                                                        \
  do {                                                                                             \
    if (use128)                                                                                    \
      launch_sgemm<128, 128, TA_, TB_>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, vA, vB, \
                                       vC, s);                                                     \
    else                                                                                           \
      launch_sgemm<64, 64, TA_, TB_>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, vA, vB,   \
                                     vC, s);                                                       \
  } while (0)
  if (!tA && !tB) GIC_SGEMM(false, false);
  else if (!tA && tB) GIC_SGEMM(false, true);
  else if (tA && !tB) GIC_SGEMM(true, false);
  else GIC_SGEMM(true, true);
#undef GIC_SGEMM
  return check_launch("sgemm_kernel");
}

// ---------------------------------------------------------------------------------------
// column sums: out[n] (+)= scale * sum_m A[m, n]      (bias gradients)
// grid.x tiles 32 columns; 32x32 threads, rows strided by 32, smem tree over the row lanes.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
colsum_kernel(const float* __restrict__ A, int M, int N, int lda, float scale, int accumulate,
              float* __restrict__ out) {
  __shared__ float part[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N)
    for (int m = ty; m < M; m += 32) s += A[(size_t)m * lda + n];
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += part[i][tx];
    if (n < N) out[n] = (accumulate ? out[n] : 0.f) + scale * t;
  }
}

// column sums of a bf16 matrix (fp32 accumulation): bias gradients in GIC_GEMM_BF16 mode
__global__ void __launch_bounds__(1024)
colsum_bf16_kernel(const unsigned short* __restrict__ A, int M, int N, int lda, int accumulate, float* __restrict__ out) {
  __shared__ float part[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N)
    for (int m = ty; m < M; m += 32) s += __uint_as_float((unsigned int)A[(size_t)m * lda + n] << 16);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += part[i][tx];
    if (n < N) out[n] = (accumulate ? out[n] : 0.f) + t;
  }
}
int colsum_bf16(const void* A, int M, int N, int lda, bool accumulate, float* out, cudaStream_t s) {
  if (N == 0) return GIC_OK;
  colsum_bf16_kernel<<<cdiv(N, 32), 1024, 0, s>>>(reinterpret_cast<const unsigned short*>(A), M, N, lda, accumulate ? 1 : 0, out);
  return check_launch("colsum_bf16_kernel");
}

int colsum_f32(const float* A, int M, int N, int lda, float scale, bool accumulate, float* out,
               cudaStream_t s) {
  if (N == 0) return GIC_OK;
  GIC_REQUIRE(A && out, GIC_ERR_NULL, "colsum_f32: NULL operand");
  colsum_kernel<<<cdiv(N, 32), 1024, 0, s>>>(A, M, N, lda, scale, accumulate ? 1 : 0, out);
  return check_launch("colsum_kernel");
}

}  // namespace gic
