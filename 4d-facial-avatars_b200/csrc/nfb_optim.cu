// nfb_optim.cu — the training step's tail as two launches: the loss of
// train_transformed_rays.py:355-389 (mse(rgb_coarse) + mse(rgb_fine); the latent-code regulariser joins in the optimizer
// kernel) as a gradient w.r.t. the rendered colours, and torch.optim.Adam (:391, YAML optimizer block) over ONE flat FP32
// bucket holding both networks and the latent-code table, with optimizer.zero_grad() fused in.  The FP32 -> kernel-layout
// re-pack that follows is fold_feat_kernel + repack_kernel (nfb_pack.cu), two more launches.
#include <cuda_runtime.h>

#include "nfb_internal.h"
#include "nfb_layout.h"

namespace nfb {

// d/d rgb of  mean((rgb - target)^2)  taken over n_total * 3 elements (n_total = the GLOBAL batch when the rays of this call
// are one shard of it: a SUM all-reduce of the parameter gradients then yields the single-process gradient).
// loss[0] += sum((rgb_c - t)^2) / (3 n_total), loss[1] likewise for the fine pass (each call adds its shard's share).
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ rgb_c, const float* __restrict__ rgb_f,
                                                        const float* __restrict__ target, int n_elems, float inv_count,
                                                        float* __restrict__ g_c, float* __restrict__ g_f, float* __restrict__ loss) {
  __shared__ float part[2][8];
  float sc = 0.f, sf = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += gridDim.x * blockDim.x) {
    const float t = target[i];
    const float dc = rgb_c[i] - t;
    g_c[i] = 2.f * dc * inv_count;
    sc = fmaf(dc, dc, sc);
    if (rgb_f) {
      const float df = rgb_f[i] - t;
      g_f[i] = 2.f * df * inv_count;
      sf = fmaf(df, df, sf);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sc += __shfl_xor_sync(0xffffffffu, sc, o);
    sf += __shfl_xor_sync(0xffffffffu, sf, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { part[0][w] = sc; part[1][w] = sf; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += part[threadIdx.x][k];
    atomicAdd(loss + threadIdx.x, s * inv_count);
  }
}

// torch.optim.Adam (no weight decay, no amsgrad), element-wise over the flat bucket, same operation order as torch's
// single-tensor implementation:  m.lerp_(g, 1 - b1);  v.mul_(b2).addcmul_(g, g, 1 - b2);
// denom = sqrt(v) / sqrt(1 - b2^t) + eps;  p -= (lr / (1 - b1^t)) * m / denom.   Gradients are zeroed after use.
// Latent-code regulariser (train_transformed_rays.py:369-372,386: 10 * 0.0005 * ||latent||_2 on the frame's row of the
// table): its gradient reg_w * l / ||l|| (0 at l == 0, as torch.norm's backward gives) is added to that row's gradient here,
// after any all-reduce, so every rank adds it exactly once.
struct AdamArgs {
  float* p; float* g; float* m; float* v;
  long long n;
  float lr_over_bc1, sqrt_bc2, b1, b2, eps, grad_scale;
  long long reg_off;   // float offset of the regularised 32-vector inside the bucket; < 0: none
  float reg_w;
};
__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
  __shared__ float reg_inv_norm;
  if (a.reg_off >= 0) {  // every block that touches the row needs 1 / ||l||: 32 values, recomputed per block (cheap, uniform)
    if (threadIdx.x < 32) {
      const float l = a.p[a.reg_off + threadIdx.x];
      float s = l * l;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (threadIdx.x == 0) reg_inv_norm = s > 0.f ? rsqrtf(s) : 0.f;
    }
    __syncthreads();
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    float g = a.g[i] * a.grad_scale;
    const float p = a.p[i];
    if (a.reg_off >= 0 && i >= a.reg_off && i < a.reg_off + kDimLatent) g = fmaf(a.reg_w * reg_inv_norm, p, g);
    float m = a.m[i], v = a.v[i];
    m = fmaf(g - m, 1.f - a.b1, m);
    v = fmaf(g * g, 1.f - a.b2, v * a.b2);
    const float denom = sqrtf(v) / a.sqrt_bc2 + a.eps;
    a.p[i] = p - a.lr_over_bc1 * (m / denom);
    a.m[i] = m;
    a.v[i] = v;
    a.g[i] = 0.f;
  }
}

// Graph-capturable variant: the step counter, the learning-rate schedule and the regularised row live in DEVICE memory, so one
// captured CUDA graph replays every iteration.  adam_prepare_kernel (1 thread) advances the step and derives this step's
// scalars exactly as launch_adam does on the host (double precision); adam_dev_kernel is adam_kernel reading them.
struct AdamDevState {  // mirrors NfbAdamDev (include/nfb.h)
  int step, pad;
  float lr0, decay_factor, decay_steps, b1, b2, eps, grad_scale, reg_w;
  long long table_off;       // float offset of the latent table in the bucket (< 0: no regulariser)
  const long long* row;      // device pointer to the current row index
  float lr_over_bc1, sqrt_bc2;
  long long reg_off;
};
__global__ void adam_prepare_kernel(AdamDevState* st) {
  const int step = ++st->step;  // 1-based number of the step being taken
  const int i = step - 1;       // the reference's loop index (train_transformed_rays.py:393-399: lr set AFTER step i)
  const double lr = (i <= 0) ? (double)st->lr0 : (double)st->lr0 * pow((double)st->decay_factor, (double)(i - 1) / (double)st->decay_steps);
  const double bc1 = 1.0 - pow((double)st->b1, (double)step), bc2 = 1.0 - pow((double)st->b2, (double)step);
  st->lr_over_bc1 = (float)(lr / bc1);
  st->sqrt_bc2 = (float)sqrt(bc2);
  st->reg_off = (st->table_off >= 0 && st->row) ? st->table_off + (long long)kDimLatent * st->row[0] : -1;
}
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ P, float* __restrict__ G, float* __restrict__ M, float* __restrict__ V,
                                                       long long n, const AdamDevState* __restrict__ st) {
  __shared__ float reg_inv_norm;
  const long long reg_off = st->reg_off;
  const float lr_over_bc1 = st->lr_over_bc1, sqrt_bc2 = st->sqrt_bc2, b1 = st->b1, b2 = st->b2, eps = st->eps, gs = st->grad_scale, reg_w = st->reg_w;
  if (reg_off >= 0) {
    if (threadIdx.x < 32) {
      const float l = P[reg_off + threadIdx.x];
      float s = l * l;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (threadIdx.x == 0) reg_inv_norm = s > 0.f ? rsqrtf(s) : 0.f;
    }
    __syncthreads();
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = G[i] * gs;
    const float p = P[i];
    if (reg_off >= 0 && i >= reg_off && i < reg_off + kDimLatent) g = fmaf(reg_w * reg_inv_norm, p, g);
    float m = M[i], v = V[i];
    m = fmaf(g - m, 1.f - b1, m);
    v = fmaf(g * g, 1.f - b2, v * b2);
    const float denom = sqrtf(v) / sqrt_bc2 + eps;
    P[i] = p - lr_over_bc1 * (m / denom);
    M[i] = m;
    V[i] = v;
    G[i] = 0.f;
  }
}
cudaError_t launch_adam_dev(float* p, float* g, float* m, float* v, long long n, void* dev_state, cudaStream_t st, long long* launches) {
  AdamDevState* s = static_cast<AdamDevState*>(dev_state);
  adam_prepare_kernel<<<1, 1, 0, st>>>(s);
  ++*launches;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_dev_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, n, s);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_loss_grad(const float* rgb_c, const float* rgb_f, const float* target, int n_rays, long long n_total, float* g_c,
                             float* g_f, float* loss, cudaStream_t st, long long* launches) {
  const int n = 3 * n_rays;
  if (n <= 0) return cudaSuccess;
  int blocks = (n + 255) / 256;
  if (blocks > 296) blocks = 296;
  loss_grad_kernel<<<blocks, 256, 0, st>>>(rgb_c, rgb_f, target, n, 1.f / (3.f * (float)n_total), g_c, g_f, loss);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_adam(float* p, float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, int step,
                        float grad_scale, long long reg_off, float reg_w, cudaStream_t st, long long* launches) {
  AdamArgs a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.n = n;
  const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  a.lr_over_bc1 = (float)((double)lr / bc1);
  a.sqrt_bc2 = (float)sqrt(bc2);
  a.b1 = b1; a.b2 = b2; a.eps = eps; a.grad_scale = grad_scale;
  a.reg_off = reg_off; a.reg_w = reg_w;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<(int)blocks, 256, 0, st>>>(a);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace nfb
