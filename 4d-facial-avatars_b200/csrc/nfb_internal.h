// nfb_internal.h — structures shared between the C-ABI layer (nfb_api.cu), the preparation kernels
// (nfb_pack.cu) and the render kernel (nfb_render.cu).  Not part of the public interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nfb {

// Device buffers of one loaded network.
struct NetBuffers {
  uint8_t* stream_x1 = nullptr;  // kStreamBytesX1: FP16 weights, swizzled units in execution order
  uint8_t* stream_x3 = nullptr;  // kStreamBytesX3: hi unit, lo unit, ...
  float* w6 = nullptr;           // [144,256] folded layers_dir.0 / fc_alpha
  float* b6 = nullptr;           // [144]
  float* bias_static = nullptr;  // [kBiasFloats]
  float* bias_frame = nullptr;   // [kBiasFloats] bias_static + per-frame fold (what the kernel reads)
  float* w0c = nullptr;          // [256,108] conditioning columns of layers_xyz.0
  float* w3c = nullptr;          // [256,108] conditioning columns of layers_xyz.3
  float* wd0b_t = nullptr;       // [24,128] direction columns of layers_dir.0, transposed
  bool loaded = false;
};

// Everything the render kernel needs, passed by value (__grid_constant__).
struct RenderParams {
  // rays
  const float* o;
  const float* d;
  int n_rays;
  float pose[12];
  float fx, fy, wcx, hcy;  // intrinsics as FP32; wcx = width*cx, hcy = height*cy (rounded like torch does)
  int width, row_begin;
  float near_, far_;
  const float* dir_z;
  const float* bg;
  // sampling
  int nc, nf, s_fine;     // s_fine = nc + nf
  int rays_per_unit;      // R
  int tiles_c, tiles_f;   // 128-row tiles per coarse / fine pass of one unit
  int n_units;
  int perturb;
  float noise_std;
  int white_bkgd;
  const float* t_coarse;  // [nc]
  const float* u_fine;    // [nf]
  const float *t_rand, *noise_c, *u_rand, *noise_f;
  // networks
  const uint8_t* wstream[2];
  const float* bias[2];
  const float* wd0b_t[2];
  // outputs
  float *rgb_c, *disp_c, *acc_c, *rgb_f, *disp_f, *acc_f, *w_last;
  // debug
  float *dbg_z_c, *dbg_raw_c, *dbg_z_f, *dbg_raw_f, *dbg_act;
  int dbg_act_step;
  unsigned long long* prof;  // optional [64] phase-cycle counters (see PhaseTimer in nfb_render.cu)
};

cudaError_t launch_load_weights(NetBuffers& nb, const float* const params[26], cudaStream_t st, long long* launches);
cudaError_t launch_frame_fold(NetBuffers& nb, const float* expr, const float* latent, cudaStream_t st, long long* launches);
// precision: 0 = fast (x1), 1 = exact (x3).  num_sms = CTAs to launch at most.
cudaError_t launch_render(const RenderParams& p, int precision, int num_sms, cudaStream_t st, long long* launches);
cudaError_t render_kernel_setup();  // opt-in to the large dynamic shared memory size

}  // namespace nfb
