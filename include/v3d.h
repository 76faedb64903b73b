/*
 * v3d.h -- C ABI of the B200-native (sm_100a) position-aware video front end.
 *
 * The reference (zd11024/Video-3D-LLM) has no FFI seam for this path: its boundary is a set
 * of Python call signatures.  The host mirror in video-3d-llm_b200/
 * keeps those signatures and binds the entry points below through ctypes; each entry point
 * cites the reference interface it replaces (paths relative to the upstream repo root).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host.
 *   - the caller owns every buffer (inputs, outputs, workspace, tables); the library keeps no
 *     state besides one-time per-device function attributes.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden sync.
 *   - return value: V3D_OK or a negative V3D_E_* code; v3d_last_error() gives the message of
 *     the calling thread's last failure.  Nothing throws or aborts across the ABI.
 *   - outputs are fully overwritten (including the zero tail channels of the PE).
 */
#ifndef V3D_H_
#define V3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V3D_VERSION 102 /* 0.1.2 */

enum v3d_status { V3D_OK = 0, V3D_E_ARG = -1, V3D_E_ALIGN = -2, V3D_E_UNSUPPORTED = -3, V3D_E_CUDA = -4 };

/* depth element type: uint16 millimetres as stored in the ScanNet PNGs, int32 as the reference
 * keeps it after np.int32 (llava/video_utils.py:217), or float32 (its .float(), :230). */
enum v3d_depth_dtype { V3D_DEPTH_U16 = 0, V3D_DEPTH_I32 = 1, V3D_DEPTH_F32 = 2 };
/* element type of PE / token buffers */
enum v3d_dtype { V3D_BF16 = 0, V3D_F16 = 1, V3D_F32 = 2 };
/* x / voxel_size: IEEE division (torch CPU, the parity target) or x * (1/voxel_size)
 * (what ATen's CUDA kernel does for a scalar divisor). */
enum v3d_div_mode { V3D_DIV_IEEE = 0, V3D_DIV_RECIPROCAL = 1 };
/* per-patch pooling of the coordinates, selected by substring of
 * world_position_embedding_type (llava/model/llava_arch.py:395-407). */
enum v3d_pool_mode { V3D_POOL_AVG = 0, V3D_POOL_MINMAX = 1, V3D_POOL_SAMPLE9 = 2, V3D_POOL_SAMPLE5 = 3, V3D_POOL_SAMPLE1 = 4 };
enum v3d_pe_mode { V3D_PE_NONE = 0, V3D_PE_SINE3D = 1 };

typedef void* v3d_stream_t;

int v3d_version(void);
const char* v3d_last_error(void);
/* Diagnostic: the kernels divide by per-frame constants (depth scale, fx, fy, voxel size) with a
 * hoisted-reciprocal sequence that must equal IEEE division bit for bit.  This runs that sequence
 * against div.rn.f32 for all 2^32 dividends of `divisor` and writes the number of mismatches
 * (device uint64) -- expected 0. */
int v3d_selftest_division(float divisor, uint64_t* mismatches, v3d_stream_t stream);
/* number of points per token for a pool mode: avg 1, minmax 2, sample9/5/1 -> 9/5/1 */
int v3d_pool_points(int pool_mode);

/* ---- unprojection ------------------------------------------------------------------------
 * Replaces llava/video_utils.py:38-68 `unproject(intrinsics, poses, depths)`.
 *   depth (V,H,W) of depth_dtype; poses, intr (V,4,4) fp32 row-major; depth_scale = 1000.
 * With row_map/col_map == NULL writes world (V,H,W,3) fp32.  With maps (int32, lengths
 * out_h/out_w) it writes only the gathered pixels world[v,row_map[r],col_map[c]] as
 * (V,out_h,out_w,3): the cv2.INTER_NEAREST resize + centre crop of :291-308 fused in.
 * clamp_lo_hi_host: NULL or 6 floats {lo xyz, hi xyz}: the 'norm' clamp of :232-234. */
int v3d_unproject(const void* depth, int depth_dtype, const float* poses, const float* intr, int V, int H, int W,
                  float depth_scale, const int32_t* row_map, const int32_t* col_map, int out_h, int out_w,
                  const float* clamp_lo_hi_host, float* world_out, v3d_stream_t stream);

/* Scene AABB [xmin,xmax,ymin,ymax,zmin,zmax] over every pixel of every frame: the "boundry"
 * entry of llava/video_utils.py:273-277, computed straight from depth.  out6: 6 floats. */
int v3d_boundary(const void* depth, int depth_dtype, const float* poses, const float* intr, int V, int H, int W,
                 float depth_scale, const float* clamp_lo_hi_host, float* out6, v3d_stream_t stream);

/* Host -> device upload of ONLY the depth samples the front end reads.  The nearest-neighbour resize + centre
 * crop (llava/video_utils.py:291-308) references a column window and, for ScanNet's 480 -> 384 rows, 4 of every
 * 5 rows: 58 % of the bytes of a frame.  depth_host (pinned) and depth_dev have the same (n_rows_total, pitch)
 * layout; the copy moves, for each group of `row_period` rows, the first `rows_keep` rows, bytes
 * [x_off_bytes, x_off_bytes + width_bytes) of each -- one strided DMA (cudaMemcpy3DAsync), stream-ordered.
 * Everything outside the window is left untouched on the device and is never read by v3d_frontend_pe. */
int v3d_upload_depth_window(const void* depth_host, void* depth_dev, size_t pitch_bytes, size_t n_rows_total,
                            size_t x_off_bytes, size_t width_bytes, size_t row_period, size_t rows_keep,
                            v3d_stream_t stream);

/* The same window in PACKED form: a loader that already skips the unread samples (scene_pack.py, FrontEnd.pack_window)
 * keeps, per frame, rows_keep of every row_period rows and width_bytes of each, contiguously -- so the host -> device
 * transfer is ONE contiguous copy (55 GB/s on this pool's PCIe instead of 47 GB/s for the strided DMA).  This expands a
 * packed DEVICE buffer into the (n_rows_total, pitch) frame layout the kernels read (device-to-device strided copy). */
int v3d_expand_depth_window(const void* packed_dev, void* depth_dev, size_t pitch_bytes, size_t n_rows_total, size_t x_off_bytes,
                            size_t width_bytes, size_t row_period, size_t rows_keep, v3d_stream_t stream);

/* Depth that the LOADER already resampled: packed_dev (V,out_h,out_w) holds depth[f, row_map[r], col_map[c]] -- the
 * nearest-neighbour resize + centre crop of llava/video_utils.py:291-308 applied to the raw depth (48 % of a ScanNet frame;
 * scene_pack.ScenePack.load(resample=...), FrontEnd.pack_resampled).  This scatters it back into the (V,H,W) layout
 * (depth_dev[f, row_map[r], col_map[c]] = packed[f, r, c]); everything else in depth_dev is left untouched and is never
 * read by v3d_frontend_pe configured with the same maps. */
int v3d_scatter_resampled_depth(const void* packed_dev, void* depth_dev, int depth_dtype, int V, int H, int W, const int32_t* row_map,
                                const int32_t* col_map, int out_h, int out_w, v3d_stream_t stream);

/* Loader helper, HOST memory only (no device work, `stream` not involved): copy n_blocks blocks of bytes_each bytes from
 * src_host[i] to dst_host + i*bytes_each on up to n_threads native threads -- the frames of a batch from the page cache
 * (packed scenes, scene_pack.py) into one pinned buffer in a single call.  Replaces the per-frame `Image.open` / `np.array`
 * loop of llava/video_utils.py:212-218 on the packed-scene path. */
int v3d_host_gather(void* dst_host, const void* const* src_host, size_t bytes_each, int n_blocks, int n_threads);

/* The same gather with the copy done by the DMA engines: n_blocks blocks of bytes_each bytes, block i from the PINNED host
 * address src_host[i], to dst_dev + i * bytes_each, stream-ordered (one cudaMemcpyAsync per run of blocks that are adjacent
 * in host memory).  For loaders that keep their packed scenes in pinned memory (scene_pack.ScenePack.pin()): the CPU
 * then moves no depth at all.  No reference equivalent. */
int v3d_h2d_gather(void* dst_dev, const void* const* src_host, size_t bytes_each, int n_blocks, v3d_stream_t stream);

/* ---- helpers on materialised coordinates (API parity with the reference's call structure) --
 * v3d_pool_coords replaces llava/model/llava_arch.py:213-223 `average_coordinate_in_patch`,
 * :225-239 `minmax_coordinate_in_patch`, :241-257 `sample_n_points`.
 *   coords (V,S,S,3) fp32; rows/cols >= grid*patch are dropped (the [:-6,:-6] trim);
 *   out (V,grid,grid,n_points,3) fp32.  Averages are sequential row-major fp32 sums / p^2,
 *   bit-identical to ATen's CPU avg_pool2d. */
int v3d_pool_coords(const float* coords, int V, int S, int patch, int grid, int pool_mode, float* out,
                    v3d_stream_t stream);
/* replaces llava/model/llava_arch.py:259-272 `discrete_coords` and
 * llava/video_utils.py:348-358 `VideoProcessor.discrete_point`:
 *   out = rint((min(max(c, lo), hi) - lo) / voxel), n_points*3 floats.  has_range = 0 skips
 *   the clamp and the shift (discrete_point with min/max_xyz_range None).
 *   coord_dtype (v3d_dtype): the dtype of the reference's tensor; for bf16/fp16 the subtraction and the quotient are
 *   each rounded to it (ATen rounds every element-wise op on 16-bit tensors); coords must hold values of that dtype. */
int v3d_discretize(const float* coords, size_t n_points, const float* min_xyz_host, const float* max_xyz_host,
                   int has_range, float voxel_size, int div_mode, int coord_dtype, float* out, v3d_stream_t stream);

/* ---- sinusoidal 3D position encoding ------------------------------------------------------
 * Lookup table behind the discretised PE: for slot s (= point*3 + axis) and
 * integer coordinate r in [0, rows) the nf channels sin/cos(r / dim_t[j]), laid out as
 * [n_slots][rows+1][Wt] elements of `dtype`, each slot shifted so that 16-byte chunks of the
 * output row map to 16-byte chunks of the table; row `rows` holds NaN (non-finite input).
 * dim_t: nf device floats = temperature ** (2*(j//2)/nf) (llava/model/position_encoding.py:24-25;
 * taken from the caller because torch's pow is not correctly rounded). */
size_t v3d_pe_table_bytes(int E, int n_points, int rows, int dtype);
int v3d_pe_table_build(const float* dim_t, int E, int n_points, int rows, int dtype, void* table, v3d_stream_t stream);

/* Replaces llava/model/position_encoding.py:17-49 `PositionEmbeddingSine3D.forward` (+ the add
 * of llava/model/llava_arch.py:515 when tokens_in != NULL).
 *   coords (n_tokens, n_points, 3) fp32; out (n_tokens, E) of `dtype`; channels past
 *   3*nf*n_points are zero.  table/rows: optional LUT (v3d_pe_table_build); integer-valued
 *   coordinates in [0,rows) are gathered from it, anything else is evaluated directly
 *   (IEEE x/dim_t, sinf/cosf).  tokens_in (same shape/dtype as out, may alias out) is added
 *   after the PE has been rounded to `dtype`. */
int v3d_pe_sine3d(const float* coords, size_t n_tokens, int n_points, int E, const float* dim_t, const void* table,
                  int rows, int dtype, const void* tokens_in, void* out, v3d_stream_t stream);

/* ---- the fused front end -------------------------------------------------------------------
 * depth/pose/intrinsics -> unproject -> nearest resample + crop -> per-patch pooling ->
 * [clamp + discretise] -> sin-3D PE -> [+ tokens] in ONE kernel: the chain
 * video_utils.py:230,296-308 -> llava_arch.py:396-420 -> position_encoding.py:17-49 ->
 * llava_arch.py:515.  One CTA per (frame, patch row); depth rows are staged in shared memory
 * by the TMA unit (cp.async.bulk), pooled sums keep ATen's sequential order. */
typedef struct v3d_frontend_desc {
  /* inputs */
  const void* depth;      /* (n_frames,H,W) */
  int depth_dtype;        /* v3d_depth_dtype */
  const float* poses;     /* (n_frames,16) */
  const float* intr;      /* (n_frames,16) */
  int n_frames, H, W;
  float depth_scale;      /* 1000 */
  /* geometry: dst pixel (r,c) reads depth[row_map[r], col_map[c]]; the first grid*patch
   * entries of each map are used */
  const int32_t* row_map;
  const int32_t* col_map;
  int src_col_min;        /* smallest / largest source column among col_map[0 .. grid*patch): the */
  int src_col_max;        /* host knows the maps; they bound the row segment the TMA unit stages */
  int patch;              /* p */
  int grid;               /* G: tokens per frame = G*G */
  int pool_mode;          /* v3d_pool_mode; SAMPLE* need patch 27 / grid 14 */
  int has_clamp;          /* 'norm': clamp world coords to clamp_lo/hi before pooling */
  float clamp_lo[3], clamp_hi[3];
  /* discretisation (the 'discrete' token) */
  int discrete;
  float min_xyz[3], max_xyz[3], voxel_size;
  int div_mode;           /* v3d_div_mode */
  /* position encoding */
  int pe_mode;            /* v3d_pe_mode */
  int E;                  /* embedding size */
  const float* dim_t;     /* (E / (3*n_points)) floats */
  const void* table;      /* LUT of v3d_pe_table_build in `dtype`, required when discrete */
  int table_rows;
  int dtype;              /* v3d_dtype of tokens_in / pe_out */
  const void* tokens_in;  /* NULL or (n_frames, G*G, E): fused add; may alias pe_out */
  void* pe_out;           /* (n_frames, G*G, E); NULL with pe_mode NONE */
  /* optional side outputs, fp32 */
  float* pooled_out;      /* (n_frames,G,G,n_points,3) pooled coordinates */
  float* pe_coords_out;   /* same shape: what the PE was evaluated on (discretised or not) */
  /* depth that the LOADER already resampled (scene_pack.py, FrontEnd.pack_resampled): `depth` is (n_frames,H,W) with
   * depth[f,r,c] = raw[f,row_map[r],col_map[c]] for r,c < grid*patch (H, W >= grid*patch; rows of W elements, so a
   * 16-byte multiple keeps the TMA path: 384 columns for the ScanNet recipe).  row_map/col_map still give the SOURCE
   * pixel coordinates u, v of every sample; src_col_min/max are ignored. */
  int depth_resampled;
  /* v3d_dtype of the coordinate tensor the reference would hold at llava_arch.py:396 (bf16 under the HF Trainer's
   * cast, fp16 after the eval scripts' .half(), model_scanqa.py:164-165): every unprojected (and clamped) coordinate,
   * every pooled mean and every step of the discretisation is rounded to it, like ATen's element-wise kernels do.
   * V3D_F32 = the fp32 contract of the oracle.  NOTE: V3D_BF16 is 0 -- set this field explicitly. */
  int coord_dtype;
  /* 0 (default): the recipe configuration (uint16 depth, avg pooling, discretised coordinates, table-backed PE added to
   * bf16/fp16 tokens, no clamp, pooled_out == NULL) runs as a streaming kernel that forms every patch mean algebraically
   * from exact integer sums over the window and accepts it only if a proven error band around it discretises to a single
   * integer; coordinates inside the band are recomputed with the reference's operator sequence in the same launch.  The
   * results are the same bit for bit either way; 1 forces the operator-sequence kernel everywhere (tests, profiling). */
  int exact_pooling;
} v3d_frontend_desc;

int v3d_frontend_pe(const v3d_frontend_desc* desc, v3d_stream_t stream);

/* ---- maximum-coverage frame sampler ---------------------------------------------------------
 * Replaces scripts/3d/preprocessing/max_coverage_sampling.py:39-104 (voxelise + greedy set
 * cover) for a batch of scenes.  See csrc/v3d_mc.cu. */
typedef struct v3d_mc_scene {
  const void* depth;        /* (n_frames,H,W) candidate frames */
  const float* poses;       /* (n_frames,16) */
  const float* intr;        /* (n_frames,16) */
  const int32_t* pc_voxels; /* (n_pc,3) mesh universe, int triples (convert_pcd_to_voxel.py:69-82) */
  int n_frames;
  int n_pc;
  uint64_t seed;            /* random.seed(seed) right before the greedy loop */
  /* workspace, caller-owned */
  int32_t* hash_table;      /* 4*hash_slots int32, 16-byte aligned: {i,j,k,index} per slot */
  int hash_slots;           /* v3d_mc_hash_slots(n_pc): power of two >= 4*n_pc */
  int pad_;
  uint32_t* bitsets;        /* (n_frames, words), words = ceil(n_pc/32): frame coverage over pc order */
  /* output */
  int32_t* record;          /* V3D_MC_RECORD_INTS int32: order[32], gains[32] (-1 padded), n_picked, n_all, n_sel, 0.
                               The voxelise stage uses record[0..7] as workspace (bounding box of the universe, flags);
                               the greedy stage overwrites the whole record afterwards. */
} v3d_mc_scene;

#define V3D_MC_MAX_PICK 32
#define V3D_MC_RECORD_INTS (2 * V3D_MC_MAX_PICK + 4)

/* hash_slots for a universe of n_pc voxels */
int v3d_mc_hash_slots(int n_pc);
/* scenes_host: array of n_scenes descriptors in HOST memory (device pointers inside).
 * scratch: device buffer of at least n_scenes*sizeof(v3d_mc_scene) bytes the descriptors are
 * staged in (stream-ordered copy).  All scenes share H, W, depth dtype, voxel size. */
int v3d_mc_sample(const v3d_mc_scene* scenes_host, int n_scenes, int H, int W, int depth_dtype, float depth_scale,
                  float voxel_size, int div_mode, int max_pick, void* scratch, v3d_stream_t stream);
/* the two stages separately (tests, profiling).  Voxelise = k_mc_clear, k_mc_build, k_mc_voxelize_dense (uint16 depth in
 * whole 16-pixel rows, universe box small enough for one bit per cell in shared memory) and k_mc_voxelize (everything the
 * dense kernel left: other scenes / frames, other depth types).  Environment V3D_MC_DENSE=0 runs k_mc_voxelize alone. */
int v3d_mc_voxelize(const v3d_mc_scene* scenes_host, int n_scenes, int H, int W, int depth_dtype, float depth_scale,
                    float voxel_size, int div_mode, void* scratch, v3d_stream_t stream);
int v3d_mc_greedy(const v3d_mc_scene* scenes_host, int n_scenes, int max_pick, void* scratch, v3d_stream_t stream);
/* voxel ids of materialised world coordinates: int(round(w / voxel)) (max_coverage_sampling.py:44-45) */
int v3d_mc_voxel_ids(const float* world, size_t n_points, float voxel_size, int div_mode, int32_t* ids, v3d_stream_t stream);

/* ---- projector tail ------------------------------------------------------
 * One kernel for llava/model/llava_arch.py:191-210 `get_2dPool` (the 2x2 spatial pooling of the projected video
 * tokens: bilinear interpolate to ceil(G/stride), or avg/max pooling to floor(G/stride)), :506-517 (the add of
 * world_position_embedding(coords)) and :307-328 `add_token_per_grid` (image_newline after every grid row).
 *   tokens (n_frames, src_grid*src_grid, E) of `dtype`, row-major grid like the vision tower emits it;
 *   coords NULL (no PE) or (n_frames, out_grid*out_grid, n_points, 3) fp32 -- what the PE is evaluated on
 *     (discretised or not), e.g. the pe_coords_out of v3d_frontend_pe; dim_t/table/table_rows as for v3d_pe_sine3d;
 *   newline NULL or E elements: out is (n_frames, out_grid, out_grid + (newline != NULL), E).
 * The pooled value is rounded to `dtype` before the PE is added (the reference materialises it), so the result has
 * the reference's two roundings. */
/* NONE keeps the grid (stride ignored): PE add and/or newline insertion only */
enum v3d_token_pool_mode { V3D_TOKEN_POOL_BILINEAR = 0, V3D_TOKEN_POOL_AVERAGE = 1, V3D_TOKEN_POOL_MAX = 2, V3D_TOKEN_POOL_NONE = 3 };
typedef struct v3d_pool_desc {
  const void* tokens;
  int n_frames, src_grid, E, dtype;
  int pool_mode; /* v3d_token_pool_mode (config.mm_spatial_pool_mode) */
  int stride;    /* config.mm_spatial_pool_stride */
  const float* coords;
  int n_points;
  const float* dim_t;
  const void* table;
  int table_rows;
  const void* newline;
  void* out;
} v3d_pool_desc;
/* tokens per side after pooling, -1 for bad arguments */
int v3d_token_pool_out_grid(int src_grid, int stride, int pool_mode);
int v3d_pool_pe_grid(const v3d_pool_desc* desc, v3d_stream_t stream);

/* ---- object proposals -------------------------------------------------------
 * v3d_object_patch_masks replaces the per-object loop of llava/model/llava_arch.py:351-378: for each of K
 * boxes (cx,cy,cz,w,h,l) and each patch of each frame, is the number of pixels whose world coordinate lies
 * inside the box (closed interval per axis) >= threshold?  coords (V,S,S,3) fp32; masks_out (K,V,grid,grid)
 * uint8.  "patch14": patch 14, grid 27, threshold int(14*14*0.5) = 98; "patch27": patch 27, grid 14,
 * threshold int(27*27*0.25) = 182.
 * v3d_masked_token_mean replaces :479-496: out[k] = mean over the tokens t with masks[k,t] != 0 of tokens[t]
 * (fp32 accumulation, bf16 result; zeros when no token is selected).  tokens (T,C) bf16, masks (K,T) uint8. */
int v3d_object_patch_masks(const float* coords, int V, int S, int patch, int grid, const float* boxes, int K, int threshold,
                           uint8_t* masks_out, v3d_stream_t stream);
/* the same masks straight from depth/pose/intrinsics (row_map/col_map as for v3d_unproject: the resize + crop; the maps must
 * hold at least grid*patch entries; clamp_lo_hi_host: the 'norm' clamp or NULL): the (V,S,S,3) tensor is never built.
 * coord_dtype (v3d_dtype): the dtype the reference's world_coords has when it reaches llava_arch.py:360 -- coordinates
 * are rounded to it before the comparison, so the result equals v3d_object_patch_masks on the materialised tensor. */
int v3d_object_patch_masks_depth(const void* depth, int depth_dtype, const float* poses, const float* intr, int V, int H, int W,
                                 float depth_scale, const int32_t* row_map, const int32_t* col_map, const float* clamp_lo_hi_host,
                                 int coord_dtype, int patch, int grid, const float* boxes, int K, int threshold, uint8_t* masks_out,
                                 v3d_stream_t stream);
/* the same with the K boxes given as (min x, min y, min z, max x, max y, max z) -- the bounds the reference's loop forms per
 * object (llava_arch.py:363-364), in whatever dtype it forms them, widened to fp32 by the caller: what the overlay's lazy
 * handle launches when the unchanged loop compares against a slice of world_coords */
int v3d_object_patch_masks_depth_bounds(const void* depth, int depth_dtype, const float* poses, const float* intr, int V, int H, int W,
                                        float depth_scale, const int32_t* row_map, const int32_t* col_map, const float* clamp_lo_hi_host,
                                        int coord_dtype, int patch, int grid, const float* min_max_xyz, int K, int threshold,
                                        uint8_t* masks_out, v3d_stream_t stream);
int v3d_masked_token_mean(const void* tokens_bf16, int T, int C, const uint8_t* masks, int K, void* out_bf16, v3d_stream_t stream);

/* ---- learned-projection PE (tensor cores) ----------------------------------------------------
 * Replaces llava/model/position_encoding.py:52-84 `PositionEmbeddingMLP.forward` (+ the add of
 * llava_arch.py:515): Linear(3,512) -> LayerNorm(512) -> ReLU -> Linear(512,E) (+ tokens).
 * The 512-wide hidden activations are produced in-kernel and never touch HBM; the 512 -> E
 * contraction runs on tcgen05 with TMEM accumulators, bias + token add in the epilogue.
 *   coords (M,3) fp32; w1 (512,3), b1, gamma, beta (512) fp32; w2 (E,512) bf16 row-major
 *   (nn.Linear weight layout), b2 (E) fp32; tokens_in NULL or (M,E) bf16; out (M,E) bf16. */
int v3d_pe_mlp(const float* coords, int M, int E, const float* w1, const float* b1, const float* gamma,
               const float* beta, float ln_eps, const void* w2_bf16, const float* b2, const void* tokens_in,
               void* out, v3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* V3D_H_ */
