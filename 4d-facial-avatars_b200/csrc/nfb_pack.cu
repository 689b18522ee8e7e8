// nfb_pack.cu — load-time and per-frame preparation kernels (not on the per-ray hot path):
//   * fold_feat_kernel : fc_feat pre-multiplied into fc_alpha and layers_dir.0[:, :256]  (FP64 accumulate)
//   * pack_step_kernel : FP32 weights -> FP16 hi/lo, written as the swizzled shared-memory image
//                        (nfb_layout.h) the render kernel bulk-copies
//   * gather_kernel    : static biases, conditioning columns, transposed direction columns
//   * frame_fold_kernel: per-frame expression/latent fold into the layer-0 / layer-3 biases
// Reference semantics: nerf/models.py:236-261 (forward), :218-233 (parameter shapes).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "nfb_internal.h"
#include "nfb_layout.h"

namespace nfb {

// W6[144][256]: rows 0..127 = Wd0[:, :256] @ Wf, row 128 = wa @ Wf, rows 129..143 = 0.
// b6[144]:      rows 0..127 = bd0 + Wd0[:, :256] @ bf, row 128 = ba + wa . bf.
__global__ void fold_feat_kernel(const float* __restrict__ Wf, const float* __restrict__ bf, const float* __restrict__ wa,
                                 const float* __restrict__ ba, const float* __restrict__ Wd0, const float* __restrict__ bd0,
                                 float* __restrict__ W6, float* __restrict__ b6) {
  const int n = blockIdx.x;   // 0..143
  const int k = threadIdx.x;  // 0..255
  if (n > 128) {
    W6[n * 256 + k] = 0.f;
    if (k == 0) b6[n] = 0.f;
    return;
  }
  const float* left = (n < 128) ? (Wd0 + (size_t)n * 280) : wa;
  double acc = 0.0;
  for (int j = 0; j < 256; ++j) acc += (double)left[j] * (double)Wf[j * 256 + k];
  W6[n * 256 + k] = (float)acc;
  if (k == 0) {
    double b = (n < 128) ? (double)bd0[n] : (double)ba[0];
    for (int j = 0; j < 256; ++j) b += (double)left[j] * (double)bf[j];
    b6[n] = (float)b;
  }
}

// One thread per 16-byte chunk (8 consecutive K) of one weight row of unit `u` (blockIdx.y) of step `s` (blockIdx.z).
struct PackArgs {
  const float* src[kNumSteps];  // step -> source matrix (row-major [out, in])
  int ld[kNumSteps];            // leading dimension
  int n_valid[kNumSteps];       // valid output rows
};
__global__ void pack_step_kernel(PackArgs a, uint8_t* __restrict__ dst_x1, uint8_t* __restrict__ dst_x3) {
  const int s = blockIdx.z;
  const StepInfo si = step_info(s);
  const int u = blockIdx.y;
  if (u >= num_units(s)) return;
  const float* __restrict__ src = a.src[s];
  const int ld = a.ld[s], n_valid = a.n_valid[s];
  const UnitInfo ui = unit_info(s, u);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ui.rows * 8) return;
  const int c16 = idx & 7;
  const int n_local = idx >> 3;
  const int n = (ui.h ? si.nh0 : 0) + n_local;  // row of the step's logical weight matrix
  __align__(16) __half hi[8];
  __align__(16) __half lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = ui.ka * 64 + c16 * 8 + e;  // logical K index of this step
    float w = 0.f;
    if (n < n_valid) {
      if (si.pe_first) {
        if (k < kDimXyz) w = src[(size_t)n * ld + k];
        else if (k >= 64) w = src[(size_t)n * ld + (kDimXyz + kDimCond) + (k - 64)];
      } else {
        w = src[(size_t)n * ld + k];
      }
    }
    hi[e] = __float2half_rn(w);
    lo[e] = __float2half_rn(w - __half2float(hi[e]));
  }
  const size_t unit_x1 = (size_t)step_offset_x1(s) + unit_offset_in_step(s, u);
  const int inner = n_local * 128 + ((c16 ^ (n_local & 7)) << 4);
  *reinterpret_cast<uint4*>(dst_x1 + unit_x1 + inner) = *reinterpret_cast<const uint4*>(hi);
  const size_t unit_x3 = 2 * unit_x1;
  *reinterpret_cast<uint4*>(dst_x3 + unit_x3 + inner) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(dst_x3 + unit_x3 + (size_t)ui.rows * 128 + inner) = *reinterpret_cast<const uint4*>(lo);
}

// Static bias block, the 108 conditioning columns of layers_xyz.0/.3 and the transposed direction
// columns of layers_dir.0.  `p` = the 26 parameter pointers, `b6` = folded step-6 bias.
struct GatherArgs {
  const float* p[26];
};
__global__ void gather_kernel(GatherArgs g, const float* __restrict__ b6, float* __restrict__ bias_static,
                              float* __restrict__ w0c, float* __restrict__ w3c, float* __restrict__ wd0b_t) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < kBiasFloats) {
    float v = 0.f;
    if (t < 1536) v = g.p[2 * (t / 256) + 1][t % 256];       // layers_xyz.{0..5}.bias
    else if (t < 1680) v = b6[t - 1536];                      // folded layers_dir.0 / fc_alpha
    else if (t < 1808) v = g.p[19][t - 1680];                 // layers_dir.1.bias
    else if (t < 1936) v = g.p[21][t - 1808];                 // layers_dir.2.bias
    else if (t < 1939) v = g.p[25][t - 1936];                 // fc_rgb.bias
    bias_static[t] = v;
  }
  if (t < 256 * kDimCond) {
    const int n = t / kDimCond, j = t % kDimCond;
    w0c[t] = g.p[0][(size_t)n * 171 + kDimXyz + j];
    w3c[t] = g.p[6][(size_t)n * 427 + kDimXyz + j];
  }
  if (t < kDimDir * 128) {
    const int j = t / 128, n = t % 128;
    wd0b_t[t] = g.p[16][(size_t)n * 280 + 256 + j];  // layers_dir.0.weight[:, 256 + j]
  }
}

// bias_frame = bias_static, then rows of step 0 and step 3 += W[:, 63:171] . [expr/3 ; latent].
__global__ void frame_fold_kernel(const float* __restrict__ expr, const float* __restrict__ latent,
                                  const float* __restrict__ bias_static, const float* __restrict__ w0c,
                                  const float* __restrict__ w3c, float* __restrict__ bias_frame) {
  __shared__ float c[kDimCond];
  const int t = threadIdx.x;
  if (t < kDimExpr) c[t] = __fdiv_rn(expr[t], 3.0f);  // (expr * 1 / 3), models.py:241
  else if (t < kDimCond) c[t] = latent[t - kDimExpr];
  __syncthreads();
  for (int i = t; i < kBiasFloats; i += blockDim.x) {
    float v = bias_static[i];
    const float* w = nullptr;
    int n = 0;
    if (i < 256) { w = w0c; n = i; }
    else if (i >= 768 && i < 1024) { w = w3c; n = i - 768; }
    if (w) {
      float acc = 0.f;
      for (int j = 0; j < kDimCond; ++j) acc = fmaf(w[n * kDimCond + j], c[j], acc);
      v += acc;
    }
    bias_frame[i] = v;
  }
}

cudaError_t launch_load_weights(NetBuffers& nb, const float* const params[26], cudaStream_t st, long long* launches) {
  fold_feat_kernel<<<144, 256, 0, st>>>(params[12], params[13], params[14], params[15], params[16], params[17], nb.w6,
                                        nb.b6);
  ++*launches;
  GatherArgs g;
  for (int i = 0; i < 26; ++i) g.p[i] = params[i];
  gather_kernel<<<(256 * kDimCond + 255) / 256, 256, 0, st>>>(g, nb.b6, nb.bias_static, nb.w0c, nb.w3c, nb.wd0b_t);
  ++*launches;
  // step -> (source matrix, leading dimension, valid rows); one launch packs every unit of every step
  PackArgs a;
  const float* src[kNumSteps] = {params[0], params[2], params[4], params[6], params[8], params[10], nb.w6, params[18], params[20], params[24]};
  const int ld[kNumSteps] = {171, 256, 256, 427, 256, 256, 256, 128, 128, 128};
  const int nv[kNumSteps] = {256, 256, 256, 256, 256, 256, 129, 128, 128, 3};
  int max_units = 0;
  for (int s = 0; s < kNumSteps; ++s) {
    a.src[s] = src[s]; a.ld[s] = ld[s]; a.n_valid[s] = nv[s];
    if (num_units(s) > max_units) max_units = num_units(s);
  }
  pack_step_kernel<<<dim3(8, max_units, kNumSteps), 256, 0, st>>>(a, nb.stream_x1, nb.stream_x3);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_frame_fold(NetBuffers& nb, const float* expr, const float* latent, cudaStream_t st, long long* launches) {
  frame_fold_kernel<<<1, 256, 0, st>>>(expr, latent, nb.bias_static, nb.w0c, nb.w3c, nb.bias_frame);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace nfb
