// v3d_objects.cu -- object-proposal patch masks and masked token means.
//
// Replaces the per-object Python loops of llava/model/llava_arch.py:
//   :351-378  for every object box: which patches of every frame have enough pixels inside the box
//             ("patch14": >= int(14*14*0.5) of the 196 pixels of a 14x14 patch on the 27x27 grid;
//              "patch27": >= int(27*27*0.25) of the 729 pixels of a 27x27 window on the 14x14 grid)
//   :479-496  object feature = mean of the token features of its patches (zeros if it has none)
// One launch each for all objects instead of ~6 ATen kernels per object.
#include <math_constants.h>

#include "v3d_common.cuh"

namespace v3d {

static constexpr int kObjThreads = 128;

// the patch's pixels are in shared memory ([3][n_px]); every warp tests its share of the boxes.  The patch's own
// bounding box (NaN coordinates ignored: they are inside no box) rejects most boxes without looking at the pixels.
__device__ __forceinline__ void count_boxes(const float* s_px, float* s_bbox, int n_px, const float* __restrict__ boxes, int K, int bounds, int threshold,
                                            uint8_t* __restrict__ masks, size_t cell_stride, size_t cell_off) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 3) {  // warp a reduces axis a
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    for (int i = lane; i < n_px; i += 32) {
      const float v = s_px[warp * n_px + i];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    for (int o = 16; o; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
      s_bbox[2 * warp] = mn;
      s_bbox[2 * warp + 1] = mx;
    }
  }
  __syncthreads();
  for (int k = warp; k < K; k += kObjThreads / 32) {
    // min = centre - size/2, max = centre + size/2 in fp32, as the reference computes them per object (:363-364)
    float lo[3], hi[3];
    bool disjoint = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (bounds) {  // (min xyz, max xyz) as the caller computed them
        lo[a] = __ldg(boxes + 6 * k + a);
        hi[a] = __ldg(boxes + 6 * k + 3 + a);
      } else {
        const float c = __ldg(boxes + 6 * k + a), h = __fdiv_rn(__ldg(boxes + 6 * k + 3 + a), 2.0f);
        lo[a] = __fsub_rn(c, h);
        hi[a] = __fadd_rn(c, h);
      }
      disjoint = disjoint || hi[a] < s_bbox[2 * a] || lo[a] > s_bbox[2 * a + 1];
    }
    int cnt = 0;
    if (!disjoint) {
      for (int i = lane; i < n_px; i += 32) {
        const float x = s_px[i], y = s_px[n_px + i], z = s_px[2 * n_px + i];
        cnt += (lo[0] <= x && x <= hi[0] && lo[1] <= y && y <= hi[1] && lo[2] <= z && z <= hi[2]) ? 1 : 0;
      }
      for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) masks[(size_t)k * cell_stride + cell_off] = cnt >= threshold ? 1 : 0;
  }
}

// one CTA per (frame, patch): the patch's pixels are staged once, then every warp tests its share of the boxes
__global__ void __launch_bounds__(kObjThreads) k_object_masks(const float* __restrict__ coords, int V, int S, int patch, int grid,
                                                             const float* __restrict__ boxes, int K, int bounds, int threshold,
                                                             uint8_t* __restrict__ masks) {
  extern __shared__ float s_px[];  // [3][patch*patch] + the patch's bounding box [6]
  const int n_px = patch * patch;
  const int cell = blockIdx.x;  // (frame, gi, gj)
  const int gj = cell % grid, gi = (cell / grid) % grid, f = cell / (grid * grid);
  for (int i = threadIdx.x; i < n_px; i += kObjThreads) {
    const int dy = i / patch, dx = i - dy * patch;
    const float* src = coords + (((size_t)f * S + gi * patch + dy) * S + gj * patch + dx) * 3;
    s_px[i] = __ldg(src);
    s_px[n_px + i] = __ldg(src + 1);
    s_px[2 * n_px + i] = __ldg(src + 2);
  }
  __syncthreads();
  count_boxes(s_px, s_px + 3 * n_px, n_px, boxes, K, bounds, threshold, masks, (size_t)V * grid * grid, (size_t)f * grid * grid + gi * grid + gj);
}

// The same masks straight from depth: the patch's pixels are unprojected in the kernel (same device function and
// operation order as v3d_unproject, so the coordinates -- and the masks -- are bit-identical to the materialised
// path) instead of being read from a (V,S,S,3) tensor that would have to be built first (56.6 MB per 32 frames).
// coord_dtype: the dtype the reference's tensor would have at this point (bf16 under the HF Trainer, fp16 in the
// eval scripts, fp32 otherwise); coordinates are rounded to it before the comparison.
template <int DT>
__global__ void __launch_bounds__(kObjThreads) k_object_masks_depth(const void* __restrict__ depth, const float* __restrict__ poses,
                                                                   const float* __restrict__ intr, int V, int H, int W, float depth_scale,
                                                                   const int32_t* __restrict__ row_map, const int32_t* __restrict__ col_map,
                                                                   Clamp clamp, int coord_dtype, int patch, int grid,
                                                                   const float* __restrict__ boxes, int K, int bounds, int threshold,
                                                                   uint8_t* __restrict__ masks) {
  extern __shared__ float s_px[];  // [3][patch*patch] + the patch's bounding box [6]
  const int n_px = patch * patch;
  const int cell = blockIdx.x;  // (frame, gi, gj)
  const int gj = cell % grid, gi = (cell / grid) % grid, f = cell / (grid * grid);
  const Camera cam = load_camera(poses, intr, f);
  const Divisor ds = make_divisor(depth_scale);
  for (int i = threadIdx.x; i < n_px; i += kObjThreads) {
    const int dy = i / patch, dx = i - dy * patch;
    const int sr = __ldg(row_map + gi * patch + dy), sc = __ldg(col_map + gj * patch + dx);
    float x, y, z;
    unproject_px(cam, load_depth<DT>(depth, ((size_t)f * H + sr) * W + sc), __fsub_rn((float)sc, cam.cx), __fsub_rn((float)sr, cam.cy), ds,
                 x, y, z);
    if (clamp.on) {
      x = apply_clamp(clamp, x, 0);
      y = apply_clamp(clamp, y, 1);
      z = apply_clamp(clamp, z, 2);
    }
    if (coord_dtype == V3D_BF16) {
      x = __bfloat162float(__float2bfloat16_rn(x));
      y = __bfloat162float(__float2bfloat16_rn(y));
      z = __bfloat162float(__float2bfloat16_rn(z));
    } else if (coord_dtype == V3D_F16) {
      x = __half2float(__float2half_rn(x));
      y = __half2float(__float2half_rn(y));
      z = __half2float(__float2half_rn(z));
    }
    s_px[i] = x;
    s_px[n_px + i] = y;
    s_px[2 * n_px + i] = z;
  }
  __syncthreads();
  count_boxes(s_px, s_px + 3 * n_px, n_px, boxes, K, bounds, threshold, masks, (size_t)V * grid * grid, (size_t)f * grid * grid + gi * grid + gj);
}

// one CTA per (object, block of 256*8 channels): compact the selected tokens chunk by chunk, accumulate in fp32
static constexpr int kMeanThreads = 256;
static constexpr int kMeanChunk = 2048;

__global__ void __launch_bounds__(kMeanThreads) k_masked_token_mean(const __nv_bfloat16* __restrict__ tokens, int T, int C,
                                                                    const uint8_t* __restrict__ masks, __nv_bfloat16* __restrict__ out) {
  __shared__ int s_idx[kMeanChunk];
  __shared__ int s_n;
  const int k = blockIdx.x;
  const int c0 = (blockIdx.y * kMeanThreads + threadIdx.x) * 8;
  const bool active = c0 < C;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int total = 0;
  const uint8_t* m = masks + (size_t)k * T;
  for (int t0 = 0; t0 < T; t0 += kMeanChunk) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int t = t0 + threadIdx.x; t < min(T, t0 + kMeanChunk); t += kMeanThreads)
      if (m[t]) s_idx[atomicAdd(&s_n, 1)] = t;
    __syncthreads();
    const int n = s_n;
    total += n;
    if (active) {
      for (int i = 0; i < n; ++i) {  // order within a chunk is arbitrary: fp32 sums, result rounded to bf16
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(tokens + (size_t)s_idx[i] * C + c0));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f2 = __bfloat1622float2(h[j]);
          acc[2 * j] += f2.x;
          acc[2 * j + 1] += f2.y;
        }
      }
    }
    __syncthreads();
  }
  if (active) {
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = total > 0 ? (float)total : 1.0f;
      const __nv_bfloat162 b = __floats2bfloat162_rn(total > 0 ? acc[2 * j] / d : 0.f, total > 0 ? acc[2 * j + 1] / d : 0.f);
      o[j] = *reinterpret_cast<const uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(out + (size_t)k * C + c0) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace v3d

using namespace v3d;

extern "C" {

int v3d_object_patch_masks(const float* coords, int V, int S, int patch, int grid, const float* boxes, int K, int threshold,
                           uint8_t* masks_out, v3d_stream_t stream) {
  if (V == 0 || K == 0) return V3D_OK;
  V3D_REQUIRE(coords && boxes && masks_out, V3D_E_ARG, "v3d_object_patch_masks: null pointer");
  V3D_REQUIRE(S > 0 && patch > 0 && grid > 0 && patch * grid <= S, V3D_E_ARG, "v3d_object_patch_masks: patch*grid=%d exceeds S=%d", patch * grid, S);
  const size_t smem = ((size_t)3 * patch * patch + 8) * sizeof(float);
  V3D_REQUIRE(smem <= 48 * 1024, V3D_E_UNSUPPORTED, "v3d_object_patch_masks: patch %d too large", patch);
  k_object_masks<<<V * grid * grid, kObjThreads, smem, (cudaStream_t)stream>>>(coords, V, S, patch, grid, boxes, K, 0, threshold, masks_out);
  return check_launch("k_object_masks");
}

static int object_masks_depth(const void* depth, int depth_dtype, const float* poses, const float* intr, int V, int H, int W,
                              float depth_scale, const int32_t* row_map, const int32_t* col_map, const float* clamp_lo_hi_host,
                              int coord_dtype, int patch, int grid, const float* boxes, int K, int bounds, int threshold, uint8_t* masks_out,
                              v3d_stream_t stream) {
  if (V == 0 || K == 0) return V3D_OK;
  V3D_REQUIRE(depth && poses && intr && row_map && col_map && boxes && masks_out, V3D_E_ARG, "v3d_object_patch_masks_depth: null pointer");
  V3D_REQUIRE(H > 0 && W > 0 && patch > 0 && grid > 0, V3D_E_ARG, "v3d_object_patch_masks_depth: bad sizes");
  V3D_REQUIRE(coord_dtype == V3D_BF16 || coord_dtype == V3D_F16 || coord_dtype == V3D_F32, V3D_E_ARG, "v3d_object_patch_masks_depth: bad coord dtype");
  const size_t smem = ((size_t)3 * patch * patch + 8) * sizeof(float);
  V3D_REQUIRE(smem <= 48 * 1024, V3D_E_UNSUPPORTED, "v3d_object_patch_masks_depth: patch %d too large", patch);
  Clamp cl;
  cl.on = clamp_lo_hi_host != nullptr;
  for (int a = 0; a < 3; ++a) {
    cl.lo[a] = cl.on ? clamp_lo_hi_host[a] : 0.f;
    cl.hi[a] = cl.on ? clamp_lo_hi_host[3 + a] : 0.f;
  }
  const int blocks = V * grid * grid;
  cudaStream_t s = (cudaStream_t)stream;
  switch (depth_dtype) {
    case V3D_DEPTH_U16:
      k_object_masks_depth<V3D_DEPTH_U16><<<blocks, kObjThreads, smem, s>>>(depth, poses, intr, V, H, W, depth_scale, row_map, col_map, cl, coord_dtype, patch, grid, boxes, K, bounds, threshold, masks_out);
      break;
    case V3D_DEPTH_I32:
      k_object_masks_depth<V3D_DEPTH_I32><<<blocks, kObjThreads, smem, s>>>(depth, poses, intr, V, H, W, depth_scale, row_map, col_map, cl, coord_dtype, patch, grid, boxes, K, bounds, threshold, masks_out);
      break;
    case V3D_DEPTH_F32:
      k_object_masks_depth<V3D_DEPTH_F32><<<blocks, kObjThreads, smem, s>>>(depth, poses, intr, V, H, W, depth_scale, row_map, col_map, cl, coord_dtype, patch, grid, boxes, K, bounds, threshold, masks_out);
      break;
    default: set_error("v3d_object_patch_masks_depth: unknown depth dtype %d", depth_dtype); return V3D_E_ARG;
  }
  return check_launch("k_object_masks_depth");
}

int v3d_object_patch_masks_depth(const void* depth, int depth_dtype, const float* poses, const float* intr, int V, int H, int W,
                                 float depth_scale, const int32_t* row_map, const int32_t* col_map, const float* clamp_lo_hi_host,
                                 int coord_dtype, int patch, int grid, const float* boxes, int K, int threshold, uint8_t* masks_out,
                                 v3d_stream_t stream) {
  return object_masks_depth(depth, depth_dtype, poses, intr, V, H, W, depth_scale, row_map, col_map, clamp_lo_hi_host, coord_dtype, patch, grid,
                            boxes, K, 0, threshold, masks_out, stream);
}

int v3d_object_patch_masks_depth_bounds(const void* depth, int depth_dtype, const float* poses, const float* intr, int V, int H, int W,
                                        float depth_scale, const int32_t* row_map, const int32_t* col_map, const float* clamp_lo_hi_host,
                                        int coord_dtype, int patch, int grid, const float* min_max_xyz, int K, int threshold,
                                        uint8_t* masks_out, v3d_stream_t stream) {
  return object_masks_depth(depth, depth_dtype, poses, intr, V, H, W, depth_scale, row_map, col_map, clamp_lo_hi_host, coord_dtype, patch, grid,
                            min_max_xyz, K, 1, threshold, masks_out, stream);
}

int v3d_masked_token_mean(const void* tokens_bf16, int T, int C, const uint8_t* masks, int K, void* out_bf16, v3d_stream_t stream) {
  if (K == 0) return V3D_OK;
  V3D_REQUIRE(tokens_bf16 && masks && out_bf16, V3D_E_ARG, "v3d_masked_token_mean: null pointer");
  V3D_REQUIRE(T > 0 && C > 0 && C % 8 == 0, V3D_E_UNSUPPORTED, "v3d_masked_token_mean: C=%d must be a multiple of 8", C);
  V3D_REQUIRE(((reinterpret_cast<uintptr_t>(tokens_bf16) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0, V3D_E_ALIGN,
              "v3d_masked_token_mean: tokens/out must be 16-byte aligned");
  dim3 grid(K, (C / 8 + kMeanThreads - 1) / kMeanThreads);
  k_masked_token_mean<<<grid, kMeanThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(tokens_bf16), T, C, masks,
                                                                      reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return check_launch("k_masked_token_mean");
}

}  // extern "C"
