// GAN losses with their backward seeds (src/utils.py:10-53) and the optimizer step
// (clip_grad_norm_ + Adam, src/training.py:194-199 with the Adam ctor at :24-26).
#include <cooperative_groups.h>

#include "gic_internal.cuh"

namespace cg = cooperative_groups;

namespace gic {

enum LossType : int { LOSS_STANDARD = 0, LOSS_JS = 1, LOSS_KL = 2, LOSS_HINGE = 3, LOSS_TV = 4, LOSS_RSGAN = 5 };

// BCEWithLogits(x, y) = max(x,0) - x*y + log1p(exp(-|x|))
__device__ __forceinline__ float bce(float x, float y) { return fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))); }

// One thread-block cluster (1 or 8 CTAs): every CTA reduces its slice, the partial sums meet in CTA 0's shared memory
// over DSMEM and are added in rank order, so the result does not depend on scheduling (no atomics, no scratch buffer).
// losses[0] = g_loss, losses[1] = d_loss (get_losses returns g first, utils.py:53).
// Seeds: dd_real = d d_loss / d d_out_real, dd_fake = d d_loss / d d_out_fake,
//        dg_out = d g_loss / d g_out.  (rsgan: g_loss only sees detached D outputs -> dg_out = 0.)
__global__ void __launch_bounds__(1024)
gan_loss_kernel(int type, const float* __restrict__ d_real, const float* __restrict__ d_fake,
                const float* __restrict__ g_out, int n, float* __restrict__ losses,
                float* __restrict__ dd_real, float* __restrict__ dd_fake, float* __restrict__ dg_out) {
  __shared__ float red[32];
  __shared__ float cpart[2][8];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank(), csize = cluster.num_blocks();
  const float inv = 1.0f / (float)n;
  float sd = 0.f, sg = 0.f;
  for (int i = crank * blockDim.x + threadIdx.x; i < n; i += csize * blockDim.x) {
    const float r = d_real[i], f = d_fake[i], g = g_out[i];
    float gr = 0.f, gf = 0.f, gg = 0.f;
    switch (type) {
      case LOSS_STANDARD:
      case LOSS_JS:
      case LOSS_KL:
        sd += bce(r, 1.f) + bce(f, 0.f);
        gr = (sigmoidf_acc(r) - 1.f) * inv;
        gf = sigmoidf_acc(f) * inv;
        if (type == LOSS_STANDARD) { sg += bce(g, 1.f); gg = (sigmoidf_acc(g) - 1.f) * inv; }
        else if (type == LOSS_JS) { sg -= bce(g, 0.f); gg = -sigmoidf_acc(g) * inv; }
        else { sg -= g; gg = -inv; }
        break;
      case LOSS_HINGE:   // Q4: F.relu intent
        sd += fmaxf(1.f - r, 0.f) + fmaxf(1.f + f, 0.f);
        gr = (1.f - r > 0.f) ? -inv : 0.f;
        gf = (1.f + f > 0.f) ? inv : 0.f;
        sg -= g; gg = -inv;
        break;
      case LOSS_TV: {    // Q4: torch.tanh intent
        const float tr = tanhf(r), tf = tanhf(f), tg = tanhf(g);
        sd += tf - tr;
        gr = -(1.f - tr * tr) * inv;
        gf = (1.f - tf * tf) * inv;
        sg -= tg; gg = -(1.f - tg * tg) * inv;
        break;
      }
      case LOSS_RSGAN: {
        const float x = r - f;
        sd += bce(x, 1.f);
        gr = (sigmoidf_acc(x) - 1.f) * inv;
        gf = -gr;
        sg += bce(-x, 1.f);
        gg = 0.f;
        break;
      }
    }
    if (dd_real) dd_real[i] = gr;
    if (dd_fake) dd_fake[i] = gf;
    if (dg_out) dg_out[i] = gg;
  }
  sd = block_sum(sd, red);
  sg = block_sum(sg, red);
  if (csize == 1) {
    if (threadIdx.x == 0) { losses[0] = sg * inv; losses[1] = sd * inv; }
    return;
  }
  if (threadIdx.x == 0) {
    float* dst = cluster.map_shared_rank(&cpart[0][0], 0);
    dst[crank] = sg;
    dst[8 + crank] = sd;
  }
  cluster.sync();
  if (crank == 0 && threadIdx.x == 0) {
    float tg = 0.f, td = 0.f;
    for (unsigned i = 0; i < csize; ++i) { tg += cpart[0][i]; td += cpart[1][i]; }
    losses[0] = tg * inv; losses[1] = td * inv;
  }
}

int gan_loss(int type, const float* d_real, const float* d_fake, const float* g_out, int n, float* losses,
             float* dd_real, float* dd_fake, float* dg_out, cudaStream_t s) {
  GIC_REQUIRE(type >= 0 && type <= LOSS_RSGAN, GIC_ERR_UNSUPPORTED, "Divergence type %d is not implemented", type);
  GIC_REQUIRE(n > 0, GIC_ERR_SHAPE, "gan_loss: empty batch");
  GIC_REQUIRE(d_real && d_fake && g_out && losses, GIC_ERR_NULL, "gan_loss: NULL operand");
  const int csize = (n >= 4096) ? 8 : 1;
  const int threads = (n >= 8192) ? 1024 : 256;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gan_loss_kernel, type, d_real, d_fake, g_out, n, losses, dd_real, dd_fake, dg_out);
  if (e != cudaSuccess) { set_error("gan_loss_kernel launch: %s", cudaGetErrorString(e)); return GIC_ERR_CUDA; }
  return check_launch("gan_loss_kernel");
}

// ---------------------------------------------------------------------------------------
// EXTENSION (north-star stage 4, not in the reference: SURVEY.md 8a rows B2/B3) -- SeqGAN-style reward, baseline and
// log-prob-weighted policy-gradient loss with its backward.
//   reward of a scored caption  = mean over the R representations of sigmoid(D logit)
//   Q[b, t-1] (value of the prefix of length t, t = 1..L-1) = mean over the n rollouts started from that prefix;
//   Q[b, L-1] = reward of the sampled caption itself.
// Rollout rows are ordered (t-1) * B*n + b*n + j (gic_decode_rollouts).
// ---------------------------------------------------------------------------------------
__global__ void rollout_q_kernel(const float* __restrict__ roll_logits /*[(L-1)*B*n*R]*/,
                                 const float* __restrict__ main_logits /*[B*R]*/, int B, int L, int n, int R,
                                 float* __restrict__ Q /*[B,L]*/) {
  const int b = blockIdx.x / L, t = blockIdx.x % L;       // t = prefix length - 1
  __shared__ float red[32];
  float s = 0.f;
  int cnt;
  if (t == L - 1) {
    cnt = R;
    for (int i = threadIdx.x; i < R; i += blockDim.x) s += sigmoidf_acc(main_logits[(size_t)b * R + i]);
  } else {
    cnt = n * R;
    const float* base = roll_logits + ((size_t)t * B * n + (size_t)b * n) * R;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) s += sigmoidf_acc(base[i]);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) Q[(size_t)b * L + t] = s / (float)cnt;
}

int rollout_q(const float* roll_logits, const float* main_logits, int B, int L, int n, int R, float* Q, cudaStream_t s) {
  GIC_REQUIRE(B >= 1 && L >= 1 && n >= 1 && R >= 1, GIC_ERR_SHAPE, "rollout_rewards: bad shape");
  GIC_REQUIRE((L == 1 || roll_logits) && main_logits && Q, GIC_ERR_NULL, "rollout_rewards: NULL operand");
  rollout_q_kernel<<<B * L, 128, 0, s>>>(roll_logits, main_logits, B, L, n, R, Q);
  return check_launch("rollout_q_kernel");
}

// loss = -(1 / (B L)) sum_{b,t} log pi(y_bt) (Q_bt - base_t),  base_t = mean_b Q_bt (baseline_mode 1) or 0;
// dlogits[b,t,v] = (A_bt / (B L)) (softmax(logits_bt)_v - [v == y_bt]).  One CTA per (b, t) row; losses[0] += row loss.
__global__ void __launch_bounds__(256)
pg_loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ ids, const float* __restrict__ Q,
               int baseline_mode, int B, int L, int V, float* __restrict__ loss, float* __restrict__ dlogits,
               float* __restrict__ logp_out /*[B,L] or null*/) {
  __shared__ float red[32];
  const int row = blockIdx.x, t = row % L;
  const float* lr = logits + (size_t)row * V;
  float base = 0.f;
  if (baseline_mode == 1 && Q) {
    float s = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) s += Q[(size_t)i * L + t];
    base = block_sum(s, red) / (float)B;
  }
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, lr[v]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) sum += expf(lr[v] - mx);
  sum = block_sum(sum, red);
  int64_t y = ids[row];
  if (y < 0 || y >= V) y = 0;
  const float adv = Q ? (Q[row] - base) : 1.f;            // Q == nullptr: plain cross entropy (weight 1)
  const float scale = adv / (float)((size_t)B * L);
  if (dlogits) {
    float* dr = dlogits + (size_t)row * V;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float p = expf(lr[v] - mx) / sum;
      dr[v] = scale * (p - (v == (int)y ? 1.f : 0.f));
    }
  }
  if (threadIdx.x == 0) {
    const float lp = (lr[y] - mx) - logf(sum);
    if (logp_out) logp_out[row] = lp;
    atomicAdd(loss, -lp * scale);
  }
}

int pg_loss(const float* logits, const int64_t* ids, const float* Q, int baseline_mode, int B, int L, int V, float* loss,
            float* dlogits, float* logp, cudaStream_t s) {
  GIC_REQUIRE(B >= 1 && L >= 1 && V >= 1, GIC_ERR_SHAPE, "pg_loss: bad shape");
  GIC_REQUIRE(logits && ids && loss, GIC_ERR_NULL, "pg_loss: NULL operand");
  GIC_REQUIRE(baseline_mode == 0 || baseline_mode == 1, GIC_ERR_UNSUPPORTED, "pg_loss: baseline mode %d", baseline_mode);
  cudaMemsetAsync(loss, 0, sizeof(float), s);
  pg_loss_kernel<<<B * L, 256, 0, s>>>(logits, ids, Q, baseline_mode, B, L, V, loss, dlogits, logp);
  return check_launch("pg_loss_kernel");
}

// ---------------------------------------------------------------------------------------
// global L2 norm (squared, accumulated into *out which the caller zeroes) over a flat buffer
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sqnorm_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) s += g[i] * g[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

__global__ void sqnorm_scalar_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += g[i] * g[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

int grad_sqnorm(const float* g, size_t n, float* out, cudaStream_t s) {
  if (n == 0) return GIC_OK;
  GIC_REQUIRE(g && out, GIC_ERR_NULL, "grad_sqnorm: NULL operand");
  const int grid = min(cdiv((long long)n, 1024), 2 * num_sms());
  if (aligned16(g)) sqnorm_kernel<<<grid, 256, 0, s>>>(g, n, out);
  else sqnorm_scalar_kernel<<<grid, 256, 0, s>>>(g, n, out);
  return check_launch("sqnorm_kernel");
}

// ---------------------------------------------------------------------------------------
// fused clip + Adam over a flat parameter buffer.
//   coef = min(1, max_norm / (sqrt(*sqnorm) + 1e-6))          (clip_grad_norm_)
//   g' = coef*g; m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)    (torch.optim.Adam)
// sqnorm is read from device memory so the step needs no host sync; max_norm <= 0 disables
// clipping.  grad_scale pre-multiplies g (1/world for a summed data-parallel all-reduce).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, size_t n, const float* __restrict__ sqnorm, float max_norm,
                 float grad_scale, float lr_over_bc1, float inv_sqrt_bc2, const float* __restrict__ bc_dev, float b1,
                 float b2, float eps) {
  if (bc_dev) { lr_over_bc1 = __ldg(bc_dev); inv_sqrt_bc2 = __ldg(bc_dev + 1); }
  float coef = grad_scale;
  if (max_norm > 0.f && sqnorm) {
    const float nrm = sqrtf(*sqnorm) * grad_scale;
    coef *= fminf(1.f, max_norm / (nrm + 1e-6f));
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_over_bc1 * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

// The same update on 16-byte vectors, two vectors per thread and trip: eight 16-byte loads in flight per thread before the
// first use (the scalar kernel kept four 4-byte loads in flight and reached 67 - 74 % of the HBM rate; this is the tail of
// the step, nothing overlaps it).  Element-wise arithmetic is unchanged, so the result is bit-identical.
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float coef, float lr_over_bc1, float inv_sqrt_bc2,
                                          float b1, float b2, float eps) {
  const float gi = g * coef;
  const float mi = b1 * m + (1.f - b1) * gi;
  const float vi = b2 * v + (1.f - b2) * gi * gi;
  m = mi;
  v = vi;
  p -= lr_over_bc1 * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
}
__global__ void __launch_bounds__(256)
clip_adam_vec_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, size_t n4,
                     const float* __restrict__ sqnorm, float max_norm, float grad_scale, float lr_over_bc1, float inv_sqrt_bc2,
                     const float* __restrict__ bc_dev, float b1, float b2, float eps) {
  if (bc_dev) { lr_over_bc1 = __ldg(bc_dev); inv_sqrt_bc2 = __ldg(bc_dev + 1); }
  float coef = grad_scale;
  if (max_norm > 0.f && sqnorm) {
    const float nrm = sqrtf(*sqnorm) * grad_scale;
    coef *= fminf(1.f, max_norm / (nrm + 1e-6f));
  }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
    const size_t j = i + stride;
    const bool two = j < n4;
    float4 g0 = g[i], m0 = m[i], v0 = v[i], p0 = p[i];
    float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = g1, v1 = g1, p1 = g1;
    if (two) { g1 = g[j]; m1 = m[j]; v1 = v[j]; p1 = p[j]; }
    adam_elem(p0.x, g0.x, m0.x, v0.x, coef, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
    adam_elem(p0.y, g0.y, m0.y, v0.y, coef, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
    adam_elem(p0.z, g0.z, m0.z, v0.z, coef, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
    adam_elem(p0.w, g0.w, m0.w, v0.w, coef, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
    m[i] = m0; v[i] = v0; p[i] = p0;
    if (two) {
      adam_elem(p1.x, g1.x, m1.x, v1.x, coef, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
      adam_elem(p1.y, g1.y, m1.y, v1.y, coef, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
      adam_elem(p1.z, g1.z, m1.z, v1.z, coef, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
      adam_elem(p1.w, g1.w, m1.w, v1.w, coef, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
      m[j] = m1; v[j] = v1; p[j] = p1;
    }
  }
}

int clip_adam(float* p, const float* g, float* m, float* v, size_t n, const float* sqnorm, float max_norm,
              float grad_scale, int step, float lr, float b1, float b2, float eps, const float* bc_dev, cudaStream_t s) {
  if (n == 0) return GIC_OK;
  GIC_REQUIRE(p && g && m && v, GIC_ERR_NULL, "clip_adam: NULL operand");
  GIC_REQUIRE(step >= 1 || bc_dev, GIC_ERR_SHAPE, "clip_adam: step must be >= 1");
  if (step < 1) step = 1;
  const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
  ProfScope prof(PROF_ADAM, 28.0 * n, s);                              // read p,g,m,v; write p,m,v
  const size_t n4 = (aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v)) ? (n >> 2) : 0;
  if (n4 > 0) {
    const int grid = (int)min((long long)cdiv((long long)n4, 512), (long long)8 * num_sms());
    clip_adam_vec_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
                                              reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4, sqnorm, max_norm,
                                              grad_scale, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), bc_dev, b1, b2, eps);
    GIC_TRY(check_launch("clip_adam_kernel"));
  }
  const size_t done = n4 << 2;
  if (done < n) {                                                      // unaligned buffers, or the last n % 4 elements
    const int grid = min(cdiv((long long)(n - done), 256), 8 * num_sms());
    clip_adam_kernel<<<grid, 256, 0, s>>>(p + done, g + done, m + done, v + done, n - done, sqnorm, max_norm, grad_scale,
                                          (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), bc_dev, b1, b2, eps);
    if (n4 == 0) return check_launch("clip_adam_kernel");
    return check_launch("clip_adam_tail_kernel");
  }
  return GIC_OK;
}

}  // namespace gic
