// v3d_fused.cuh -- parameter block shared by the fused front-end kernels: k_fused (v3d_frontend.cu, the reference's
// operator sequence) and k_fused_ring (v3d_fused_ring.cu, algebraic patch means + TMA token ring).
#pragma once

#include "v3d_common.cuh"
#include "v3d_pe.cuh"

namespace v3d {

static constexpr int kFusedThreads = 256;

struct FusedParams {
  const void* depth;
  const float* poses;
  const float* intr;
  const float* coords_in;  // FROM_DEPTH == false: (V,S,S,3) materialised coordinates
  int n_frames, H, W, S;
  float depth_scale;
  const int32_t* row_map;
  const int32_t* col_map;
  int patch, grid, Sp, Sp_pad;  // Sp = patch*grid columns/rows used
  int pool_mode, n_points;
  Clamp clamp;
  Discretize disc;
  int pe_mode;
  PeGeom g;
  const float* dim_t;
  const void* table;
  const void* tokens_in;
  void* pe_out;
  float* pooled_out;
  float* pe_coords_out;
  // TMA staging: bytes [col_lo_byte, col_lo_byte + span) of each needed depth row
  int col_lo_byte, span;
  // depth is already resampled: sample (r, c) of the (H, W) buffer belongs to dst pixel (r, c), whose source pixel
  // coordinates (for u - cx, v - cy) are still (row_map[r], col_map[c])
  int resampled;
  // dtype of the coordinate tensor the reference would hold (bf16 under the HF Trainer, fp16 in the eval scripts): each
  // unprojected (and clamped) coordinate and each pooled mean is rounded to it; V3D_F32 = the fp32 contract
  int coord_dtype;
};


__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}

// v3d_fused_ring.cu: true when the ring kernel can take this configuration
bool fused_ring_supported(const FusedParams& p, int depth_dtype, bool use_tma, int out_dtype);
int launch_fused_ring(const FusedParams& p, int out_dtype, cudaStream_t s);

}  // namespace v3d
