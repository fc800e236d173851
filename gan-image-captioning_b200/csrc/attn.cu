// EXTENSION (north-star stage 1; not in the reference): additive attention over the CNN feature
// grid inside the decode step.  Definition (oracle/ref_ext.py, "parity unpinned by reference"):
//   once per image:  Ak = grid W_k^T [B,P,Da],  Av = grid W_v^T [B,P,E]      (tcgen05 GEMMs, gemm())
//   every step:      q = h_{t-1} W_q^T (GEMM);  s_l = w_e . tanh(Ak_l + q);  alpha = softmax_l(s);
//                    x'_t = x_t + sum_l alpha_l Av_l                           (attn_fwd_kernel: one CTA per caption)
// The score dot products and the softmax over the P locations are warp-shuffle reductions; Ak / Av (38 MB at c2) stay
// L2-resident across the L steps.
#include "gic_internal.cuh"

namespace gic {

// grid = B, block = 256.  smem: q[Da] | w_e[Da] | sc[P]
__global__ void __launch_bounds__(256)
attn_fwd_kernel(const float* __restrict__ Ak, const float* __restrict__ Av, const float* __restrict__ q,
                const float* __restrict__ w_e, int P, int Da, int E, float* __restrict__ x /*[B,E], += ctx*/,
                float* __restrict__ alpha_out /*[B,P]*/) {
  extern __shared__ float sm[];
  float* q_s = sm;
  float* we_s = q_s + Da;
  float* sc = we_s + Da;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int d = threadIdx.x; d < Da; d += blockDim.x) { q_s[d] = q[(size_t)b * Da + d]; we_s[d] = w_e[d]; }
  __syncthreads();
  const float* Akb = Ak + (size_t)b * P * Da;
  for (int l = warp; l < P; l += nw) {
    float s = 0.f;
    for (int d = lane; d < Da; d += 32) s = fmaf(we_s[d], tanhf(Akb[(size_t)l * Da + d] + q_s[d]), s);
    s = warp_sum(s);
    if (lane == 0) sc[l] = s;
  }
  __syncthreads();
  if (warp == 0) {                                  // softmax over the P locations
    float mx = -INFINITY;
    for (int l = lane; l < P; l += 32) mx = fmaxf(mx, sc[l]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int l = lane; l < P; l += 32) { const float e = expf(sc[l] - mx); sc[l] = e; sum += e; }
    sum = warp_sum(sum);
    for (int l = lane; l < P; l += 32) { const float a = sc[l] / sum; sc[l] = a; alpha_out[(size_t)b * P + l] = a; }
  }
  __syncthreads();
  const float* Avb = Av + (size_t)b * P * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < P; ++l) acc = fmaf(sc[l], Avb[(size_t)l * E + e], acc);
    x[(size_t)b * E + e] += acc;
  }
}

// Backward of the attention, split so that nothing of size [B, P, .] is read-modify-written per step (the first version
// accumulated dAk and dAv -- 38 MB at c2 -- inside every step's kernel: 63 us per step, 23 % of the c2a step):
//   per step (serial: dq_t feeds the recurrent gradient)        attn_bwd_step_kernel, grid = B
//     dalpha_l = <dx'_t, Av_l>;  ds = alpha (dalpha - <alpha, dalpha>)  -> ds_out[t, b, :]
//     dq_t = sum_l ds_l w_e (1 - tanh^2(Ak_l + q_t))
//   once, after the loop                                          attn_bwd_accum_kernel, grid = B
//     dAv_l = sum_t alpha_{t,l} dx'_t;   dAk_l = sum_t ds_{t,l} w_e (1 - tanh^2(Ak_l + q_t));   dw_e += sum_{t,l} ds_{t,l} tanh(.)
//   (sums over t in registers; the caption's alpha, ds, q and dx' of all L steps are staged in shared memory first)
// smem of the step kernel: dx[E] | alpha[P] | ds[P] | q[Da] | w_e[Da]
__global__ void __launch_bounds__(256)
attn_bwd_step_kernel(const float* __restrict__ dxp, const float* __restrict__ alpha, const float* __restrict__ q,
                     const float* __restrict__ Ak, const float* __restrict__ Av, const float* __restrict__ w_e, int P, int Da,
                     int E, float* __restrict__ dq /*[B,Da]*/, float* __restrict__ ds_out /*[B,P]*/) {
  extern __shared__ float sm[];
  float* dx_s = sm;
  float* al_s = dx_s + E;
  float* ds_s = al_s + P;
  float* q_s = ds_s + P;
  float* we_s = q_s + Da;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int e = threadIdx.x; e < E; e += blockDim.x) dx_s[e] = dxp[(size_t)b * E + e];
  for (int l = threadIdx.x; l < P; l += blockDim.x) al_s[l] = alpha[(size_t)b * P + l];
  for (int d = threadIdx.x; d < Da; d += blockDim.x) { q_s[d] = q[(size_t)b * Da + d]; we_s[d] = w_e[d]; }
  __syncthreads();
  const float* Avb = Av + (size_t)b * P * E;
  for (int l = warp; l < P; l += nw) {
    float s = 0.f;
    for (int e = lane; e < E; e += 32) s = fmaf(dx_s[e], Avb[(size_t)l * E + e], s);
    s = warp_sum(s);
    if (lane == 0) ds_s[l] = s;                       // dalpha_l for now
  }
  __syncthreads();
  if (warp == 0) {
    float dot = 0.f;
    for (int l = lane; l < P; l += 32) dot = fmaf(al_s[l], ds_s[l], dot);
    dot = warp_sum(dot);
    for (int l = lane; l < P; l += 32) {
      const float v = al_s[l] * (ds_s[l] - dot);
      ds_s[l] = v;
      ds_out[(size_t)b * P + l] = v;
    }
  }
  __syncthreads();
  const float* Akb = Ak + (size_t)b * P * Da;
  for (int d = threadIdx.x; d < Da; d += blockDim.x) {
    float dqv = 0.f;
    const float qd = q_s[d], wd = we_s[d];
    for (int l = 0; l < P; ++l) {
      const float th = tanhf(Akb[(size_t)l * Da + d] + qd);
      dqv = fmaf(ds_s[l] * wd, 1.f - th * th, dqv);
    }
    dq[(size_t)b * Da + d] = dqv;
  }
}

// grid = B, block = 256.  smem: alpha[L][P] | ds[L][P] | q[L][Da] | dx[L][E] | w_e[Da]
__global__ void __launch_bounds__(256)
attn_bwd_accum_kernel(const float* __restrict__ alpha /*[L,B,P]*/, const float* __restrict__ ds /*[L,B,P]*/,
                      const float* __restrict__ q /*[L,B,Da]*/, const float* __restrict__ dX /*[L,B,E]*/,
                      const float* __restrict__ Ak, const float* __restrict__ w_e, int B, int L, int P, int Da, int E,
                      float* __restrict__ dAk, float* __restrict__ dAv, float* __restrict__ dw_e /*[Da], atomics*/) {
  extern __shared__ float sm[];
  float* al_s = sm;                       // [L][P]
  float* ds_s = al_s + L * P;             // [L][P]
  float* q_s = ds_s + L * P;              // [L][Da]
  float* dx_s = q_s + L * Da;             // [L][E]
  float* we_s = dx_s + L * E;             // [Da]
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < L * P; i += blockDim.x) {
    const int t = i / P, l = i % P;
    al_s[i] = alpha[((size_t)t * B + b) * P + l];
    ds_s[i] = ds[((size_t)t * B + b) * P + l];
  }
  for (int i = threadIdx.x; i < L * Da; i += blockDim.x) q_s[i] = q[((size_t)(i / Da) * B + b) * Da + i % Da];
  for (int i = threadIdx.x; i < L * E; i += blockDim.x) dx_s[i] = dX[((size_t)(i / E) * B + b) * E + i % E];
  for (int d = threadIdx.x; d < Da; d += blockDim.x) we_s[d] = w_e[d];
  __syncthreads();
  // dAv[b, l, e] = sum_t alpha[t, l] dx'[t, e]
  float* dAvb = dAv + (size_t)b * P * E;
  for (int i = threadIdx.x; i < P * E; i += blockDim.x) {
    const int l = i / E, e = i % E;
    float acc = 0.f;
    for (int t = 0; t < L; ++t) acc = fmaf(al_s[t * P + l], dx_s[t * E + e], acc);
    dAvb[i] = acc;
  }
  // dAk[b, l, d] = w_e[d] sum_t ds[t, l] (1 - tanh^2(Ak[l, d] + q[t, d]));  dw_e[d] += sum_{t, l} ds[t, l] tanh(.)
  const float* Akb = Ak + (size_t)b * P * Da;
  float* dAkb = dAk + (size_t)b * P * Da;
  for (int d = threadIdx.x; d < Da; d += blockDim.x) {         // thread = attention unit d (coalesced over d), loop over l
    float dwe = 0.f;
    const float wd = we_s[d];
    for (int l = 0; l < P; ++l) {
      const float ak = Akb[(size_t)l * Da + d];
      float acc = 0.f;
      for (int t = 0; t < L; ++t) {
        const float th = tanhf(ak + q_s[t * Da + d]);
        const float s = ds_s[t * P + l];
        acc = fmaf(s, 1.f - th * th, acc);
        dwe = fmaf(s, th, dwe);
      }
      dAkb[(size_t)l * Da + d] = acc * wd;
    }
    atomicAdd(dw_e + d, dwe);
  }
}

int attn_fwd(const float* Ak, const float* Av, const float* q, const float* w_e, int B, int P, int Da, int E, float* x,
             float* alpha, cudaStream_t s) {
  if (B == 0) return GIC_OK;
  const size_t smem = (size_t)(2 * Da + P) * sizeof(float);
  GIC_REQUIRE(smem <= 48 * 1024, GIC_ERR_SHAPE, "attention: attn_dim / locations too large for shared memory");
  attn_fwd_kernel<<<B, 256, smem, s>>>(Ak, Av, q, w_e, P, Da, E, x, alpha);
  return check_launch("attn_fwd_kernel");
}

int attn_bwd_step(const float* dxp, const float* alpha, const float* q, const float* Ak, const float* Av, const float* w_e,
                  int B, int P, int Da, int E, float* dq, float* ds_out, cudaStream_t s) {
  if (B == 0) return GIC_OK;
  const size_t smem = (size_t)(E + 2 * P + 2 * Da) * sizeof(float);
  GIC_REQUIRE(smem <= 48 * 1024, GIC_ERR_SHAPE, "attention: dims too large for shared memory");
  attn_bwd_step_kernel<<<B, 256, smem, s>>>(dxp, alpha, q, Ak, Av, w_e, P, Da, E, dq, ds_out);
  return check_launch("attn_bwd_step_kernel");
}

int attn_bwd_accum(const float* alpha, const float* ds, const float* q, const float* dX, const float* Ak, const float* w_e,
                   int B, int L, int P, int Da, int E, float* dAk, float* dAv, float* dw_e, cudaStream_t s) {
  if (B == 0) return GIC_OK;
  const size_t smem = ((size_t)L * (2 * P + Da + E) + Da) * sizeof(float);
  GIC_REQUIRE(smem <= 200 * 1024, GIC_ERR_SHAPE, "attention: L * (2 P + Da + E) floats exceed shared memory");
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(attn_bwd_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  attn_bwd_accum_kernel<<<B, 256, smem, s>>>(alpha, ds, q, dX, Ak, w_e, B, L, P, Da, E, dAk, dAv, dw_e);
  return check_launch("attn_bwd_accum_kernel");
}

}  // namespace gic
