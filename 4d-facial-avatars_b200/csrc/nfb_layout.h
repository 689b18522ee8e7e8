// nfb_layout.h — the MLP as the kernel executes it: 10 tensor-core "steps", each a GEMM
//   D[128 rows, N] = A[128 rows, K] * W[N, K]^T   (FP16 operands, FP32 accumulate in TMEM)
// and the byte layout of the packed weight stream the kernel bulk-copies into shared memory.
// Shared by host (packing, tests) and device code.
//
// Reference model: ConditionalBlendshapePaperNeRFModel.forward (nerf/models.py:236-261).  Exact
// algebra applied at load / per frame (SURVEY.md §8 a''):
//   * the 108 expression/latent columns of layers_xyz.0 and layers_xyz.3 multiply a per-frame constant
//     vector -> folded into those layers' biases (nfb_set_frame);
//   * fc_feat has no activation -> pre-multiplied into fc_alpha and layers_dir.0[:, :256] (FP64 fold);
//   * the 24 direction-encoding columns of layers_dir.0 depend on the ray only -> a per-ray bias.
//
//   step  reference layer            N (half 0 + half 1)  K (atoms of 64)            A operand
//   0     layers_xyz.0               128 + 128            64  = PE(63)+pad           SMEM (PE buffer)
//   1,2   layers_xyz.1,2             128 + 128            256                        TMEM
//   3     layers_xyz.3               128 + 128            320 = PE(63)+pad | h(256)  SMEM atom + TMEM
//   4,5   layers_xyz.4,5             128 + 128            256                        TMEM
//   6     layers_dir.0∘fc_feat | σ   128 + 16 (σ)         256                        TMEM
//   7,8   layers_dir.1,2             128                  128                        TMEM
//   9     fc_rgb                     16 (3 used)          128                        TMEM
//
// Every step is issued as (up to) two N-halves with separate accumulators so that the epilogue of one half
// overlaps the MMAs of the other and the next step can start on the K atoms that are already converted.
// A weight "unit" = the rows of one half x one 64-wide K atom (<= 16 KB), in the order the MMA warp consumes them:
//   [PE atom of half 0, of half 1] [hidden atoms 0,1 of half 0, of half 1]  |  [hidden atoms 2,3 of half 0, of half 1]
//   `-------------------------- group 1 ---------------------------------'     `------------ group 2 -------------'
// group 1 needs only what the half-0 epilogue of the previous step produced, group 2 also its half-1 epilogue.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NFB_HD __host__ __device__
#else
#define NFB_HD
#endif

namespace nfb {

constexpr int kTileM = 128;      // rows (samples) per tensor-core tile == TMEM lanes
constexpr int kAtomK = 64;       // fp16 elements per 128-byte swizzle row
constexpr int kNumSteps = 10;
constexpr int kMaxUnitBytes = 128 * 128;  // one weight unit: <=128 output rows x 64 K x 2 B
constexpr int kDimXyz = 63, kDimDir = 24, kDimExpr = 76, kDimLatent = 32, kDimCond = 108;

struct StepInfo {
  int16_t nh0, nh1;   // output columns (= MMA N) of half 0 / half 1; nh1 == 0: single half
  int16_t k_atoms;    // K / 64 including the PE atom
  int16_t pe_first;   // 1: the first K atom is the positional encoding (A from shared memory)
  int16_t bias_off;   // float offset of this step's bias vector in the per-network bias block
  int16_t n_bias;     // bias floats
};

NFB_HD constexpr StepInfo step_info(int s) {
  return s == 0   ? StepInfo{128, 128, 1, 1, 0, 256}
         : s == 1 ? StepInfo{128, 128, 4, 0, 256, 256}
         : s == 2 ? StepInfo{128, 128, 4, 0, 512, 256}
         : s == 3 ? StepInfo{128, 128, 5, 1, 768, 256}
         : s == 4 ? StepInfo{128, 128, 4, 0, 1024, 256}
         : s == 5 ? StepInfo{128, 128, 4, 0, 1280, 256}
         : s == 6 ? StepInfo{128, 16, 4, 0, 1536, 144}
         : s == 7 ? StepInfo{128, 0, 2, 0, 1680, 128}
         : s == 8 ? StepInfo{128, 0, 2, 0, 1808, 128}
                  : StepInfo{16, 0, 2, 0, 1936, 16};
}
constexpr int kBiasFloats = 1952;  // 6*256 + 144 + 128 + 128 + 16

// Bytes of the FP16 "hi" weights of one step (x1 stream); the x3 stream stores hi then lo per unit.
NFB_HD constexpr int step_bytes_x1(int s) {
  return (step_info(s).nh0 + step_info(s).nh1) * step_info(s).k_atoms * 128;
}
NFB_HD constexpr int step_offset_x1(int s) {
  int off = 0;
  for (int i = 0; i < s; ++i) off += step_bytes_x1(i);
  return off;
}
constexpr int kStreamBytesX1 = step_offset_x1(kNumSteps);  // 864256
constexpr int kStreamBytesX3 = 2 * kStreamBytesX1;

// The u-th weight unit of step s in consumption order.
struct UnitInfo {
  int16_t h;        // N-half (0/1): accumulator half and row block [h ? nh0 : 0, ...) of the layer's weight matrix
  int16_t ka;       // K atom of the step's logical K axis (0 = the PE atom when the step has one)
  int16_t from_pe;  // A operand comes from the shared-memory PE buffer
  int16_t group;    // 1: needs the previous step's half-0 epilogue only; 2: also its half-1 epilogue
  int16_t rows;     // output rows (MMA N) of the unit
  int16_t last;     // last unit of its half in this step (-> commit "accumulator half complete")
};
NFB_HD constexpr int num_units(int s) {
  return step_info(s).k_atoms * (step_info(s).nh1 > 0 ? 2 : 1);
}
NFB_HD constexpr UnitInfo unit_info(int s, int u) {
  const StepInfo si = step_info(s);
  const int halves = si.nh1 > 0 ? 2 : 1;
  const int hid = si.k_atoms - si.pe_first;          // hidden (TMEM) atoms: 0, 2 or 4
  const int npe = si.pe_first ? halves : 0;
  int h = 0, ka = 0, from_pe = 0, group = 1, last = 0;
  if (u < npe) {
    h = u; ka = 0; from_pe = 1; last = (hid == 0);
  } else {
    const int v = u - npe;
    const int g1 = hid < 2 ? hid : 2;
    if (v < g1 * halves) {
      h = v / g1; ka = si.pe_first + v % g1; last = (hid <= 2 && v % g1 == g1 - 1);
    } else {
      const int w = v - g1 * halves;
      h = w / 2; ka = si.pe_first + 2 + w % 2; group = 2; last = (w % 2 == 1);
    }
  }
  return UnitInfo{(int16_t)h, (int16_t)ka, (int16_t)from_pe, (int16_t)group, (int16_t)(h ? si.nh1 : si.nh0), (int16_t)last};
}
// Byte offset of unit u inside its step, x1 stream.
NFB_HD constexpr int unit_offset_in_step(int s, int u) {
  int off = 0;
  for (int i = 0; i < u; ++i) off += unit_info(s, i).rows * 128;
  return off;
}
// Byte offset of element (row n, k in [0,64)) inside one swizzled unit: 128-byte rows, 16-byte chunks
// XORed with (row & 7) — the SWIZZLE_128B pattern the UMMA shared-memory descriptor expects.
NFB_HD constexpr int sw128_offset(int n, int k) { return n * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + ((k & 7) << 1); }

// Algorithmic cost used for the roofline: 550,016 MAC per MLP evaluation.
constexpr long long kAlgoFlopPerEval = 1100032LL;
// MACs the kernel actually issues per evaluation after the folds above.
constexpr long long kExecMacPerEval = 64LL * 256 + 4LL * 256 * 256 + 320LL * 256 + 256LL * 144 + 2LL * 128 * 128 + 128LL * 16;

}  // namespace nfb
