// v3d_frontend.cu -- unprojection, patch pooling, discretisation, sin-3D position encoding and
// the fused depth -> PE (+token add) kernel for sm_100a.  C ABI: include/v3d.h.
//
// Arithmetic contract: every fp32 step that decides a bit of the
// reference's result is an explicitly rounded intrinsic (v3d_common.cuh); pooled sums keep
// ATen's sequential row-major order; PE values are sinf/cosf of an IEEE quotient.
#include <math_constants.h>
#include <stdarg.h>

#include <type_traits>

#include "v3d_common.cuh"
#include "v3d_fused.cuh"
#include "v3d_pe.cuh"

namespace v3d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return V3D_E_CUDA;
}

static constexpr int kThreads = 256;

// ============================================================================================
// 1. stand-alone unprojection (API parity with llava/video_utils.py:38-68, optional gather)
// ============================================================================================
template <int DT>
__global__ void __launch_bounds__(kThreads) k_unproject(const void* __restrict__ depth, const float* __restrict__ poses,
                                                        const float* __restrict__ intr, int V, int H, int W,
                                                        float depth_scale, const int32_t* __restrict__ row_map,
                                                        const int32_t* __restrict__ col_map, int OH, int OW, Clamp clamp,
                                                        float* __restrict__ out) {
  // each thread: 4 consecutive output pixels; a warp's 384 floats are transposed through
  // shared memory so that global stores are contiguous float4s.
  __shared__ __align__(16) float stage[kThreads / 32][4 * 32 * 3];
  const size_t total = (size_t)V * OH * OW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t n_groups = (total + 3) / 4;
  const size_t groups_per_block = kThreads;
  const Divisor ds = make_divisor(depth_scale);
  for (size_t gbase = (size_t)blockIdx.x * groups_per_block; gbase < n_groups; gbase += (size_t)gridDim.x * groups_per_block) {
    const size_t g = gbase + threadIdx.x;
    float vals[12];
    if (g < n_groups) {
      size_t px = g * 4;
      int f = (int)(px / ((size_t)OH * OW));
      int rem = (int)(px - (size_t)f * OH * OW);
      int r = rem / OW, c = rem - r * OW;
      Camera cam = load_camera(poses, intr, f);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (px + e < total) {
          const int sr = row_map ? __ldg(row_map + r) : r;
          const int sc = col_map ? __ldg(col_map + c) : c;
          const float d = load_depth<DT>(depth, ((size_t)f * H + sr) * W + sc);
          unproject_px(cam, d, __fsub_rn((float)sc, cam.cx), __fsub_rn((float)sr, cam.cy), ds, x, y, z);
          x = apply_clamp(clamp, x, 0);
          y = apply_clamp(clamp, y, 1);
          z = apply_clamp(clamp, z, 2);
          if (++c == OW) {
            c = 0;
            if (++r == OH) {
              r = 0;
              ++f;
              if (f < V) cam = load_camera(poses, intr, f);
            }
          }
        }
        vals[3 * e + 0] = x;
        vals[3 * e + 1] = y;
        vals[3 * e + 2] = z;
      }
    }
    float* st = stage[warp];
    if (g < n_groups) {
#pragma unroll
      for (int k = 0; k < 12; ++k) st[lane * 12 + k] = vals[k];
    }
    __syncwarp();
    // warp writes floats [3*4*warp_first_group, +384)
    const size_t warp_first = (gbase + (size_t)warp * 32) * 12;
    const size_t total_f = total * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int idx = (k * 32 + lane) * 4;
      const size_t o = warp_first + idx;
      if (o + 4 <= total_f) {
        *reinterpret_cast<float4*>(out + o) = *reinterpret_cast<const float4*>(st + idx);
      } else {
        for (int e = 0; e < 4; ++e)
          if (o + e < total_f) out[o + e] = st[idx + e];
      }
    }
    __syncwarp();
  }
}

// Rows of whole 4-pixel groups (OW % 4 == 0, every shape of the reference): one CTA per (frame, block of rows), so the
// camera is loaded and its divisors are prepared once per CTA instead of once per group, and the pixel index needs no
// 64-bit division.  Same staging: a warp's 32 groups are 384 consecutive output floats.
template <int DT>
__global__ void __launch_bounds__(kThreads) k_unproject_rows(const void* __restrict__ depth, const float* __restrict__ poses,
                                                             const float* __restrict__ intr, int V, int H, int W, float depth_scale,
                                                             const int32_t* __restrict__ row_map, const int32_t* __restrict__ col_map,
                                                             int OH, int OW, int rows_per_block, uint32_t gpr_magic, Clamp clamp,
                                                             float* __restrict__ out) {
  __shared__ __align__(16) float stage[kThreads / 32][4 * 32 * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bpf = (OH + rows_per_block - 1) / rows_per_block;
  const int f = blockIdx.x / bpf, r0 = (blockIdx.x - f * bpf) * rows_per_block;
  const int r1 = min(OH, r0 + rows_per_block);
  const int gpr = OW >> 2;
  const int n = (r1 - r0) * gpr;
  const Camera cam = load_camera(poses, intr, f);
  const Divisor ds = make_divisor(depth_scale);
  float* obase = out + ((size_t)f * OH + r0) * OW * 3;
  const size_t total_f = (size_t)n * 12;
  // four uint16 depths with one 64-bit load when the frame is read in place
  const bool vec = DT == V3D_DEPTH_U16 && col_map == nullptr && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(depth) & 7) == 0;
  for (int base = 0; base < n; base += kThreads) {
    const int idx = base + (int)threadIdx.x;
    float vals[12];
    if (idx < n) {
      const int rr = (int)__umulhi((uint32_t)idx, gpr_magic);  // idx / gpr (idx < 2^16)
      const int c = (idx - rr * gpr) << 2;
      const int r = r0 + rr;
      const int sr = row_map ? __ldg(row_map + r) : r;
      const float vm = __fsub_rn((float)sr, cam.cy);
      const size_t row_base = ((size_t)f * H + sr) * W;
      float d[4];
      int sc[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[e] = col_map ? __ldg(col_map + c + e) : c + e;
      if (vec) {
        if constexpr (DT == V3D_DEPTH_U16) {
          const uint2 q = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(depth) + row_base + c));
          d[0] = (float)(q.x & 0xffffu);
          d[1] = (float)(q.x >> 16);
          d[2] = (float)(q.y & 0xffffu);
          d[3] = (float)(q.y >> 16);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] = load_depth<DT>(depth, row_base + sc[e]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float x, y, z;
        unproject_px(cam, d[e], __fsub_rn((float)sc[e], cam.cx), vm, ds, x, y, z);
        vals[3 * e + 0] = apply_clamp(clamp, x, 0);
        vals[3 * e + 1] = apply_clamp(clamp, y, 1);
        vals[3 * e + 2] = apply_clamp(clamp, z, 2);
      }
    }
    float* st = stage[warp];
    if (idx < n) {
#pragma unroll
      for (int k = 0; k < 12; ++k) st[lane * 12 + k] = vals[k];
    }
    __syncwarp();
    const size_t warp_first = (size_t)(base + warp * 32) * 12;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int j = (k * 32 + lane) * 4;
      const size_t o = warp_first + j;
      if (o + 4 <= total_f) st_stream_v4(obase + o, *reinterpret_cast<const uint4*>(st + j));
    }
    __syncwarp();
  }
}

// Depth that the loader already resampled (nearest-neighbour resize + crop of llava/video_utils.py:291-308 applied to the
// uint16 depth instead of the fp32 coordinates: out_h x out_w samples per frame, 48 % of a ScanNet frame) is scattered
// back to the pixels it was taken from, so that every kernel keeps reading the (V,H,W) layout with the original pixel
// coordinates.  Pixels that the maps do not reference are left untouched (never read by v3d_frontend_pe).
template <typename T>
__global__ void __launch_bounds__(kThreads) k_scatter_resampled(const T* __restrict__ packed, T* __restrict__ depth, int V, int H, int W,
                                                                const int32_t* __restrict__ row_map, const int32_t* __restrict__ col_map,
                                                                int OH, int OW) {
  for (long row = blockIdx.x; row < (long)V * OH; row += gridDim.x) {
    const int f = (int)(row / OH), r = (int)(row - (long)f * OH);
    const T* src = packed + (size_t)row * OW;
    T* dst = depth + ((size_t)f * H + __ldg(row_map + r)) * W;
    for (int c = threadIdx.x; c < OW; c += kThreads) dst[__ldg(col_map + c)] = src[c];
  }
}

// ============================================================================================
// 2. scene AABB ("boundry") straight from depth
// ============================================================================================
__device__ __forceinline__ int fkey_max(float f) {  // monotone int key, NaN -> INT_MAX
  if (f != f) return 0x7fffffff;
  int b = __float_as_int(f);
  return b >= 0 ? b : (b ^ 0x7fffffff);
}
__device__ __forceinline__ int fkey_min(float f) {  // NaN -> INT_MIN
  if (f != f) return (int)0x80000000;
  int b = __float_as_int(f);
  return b >= 0 ? b : (b ^ 0x7fffffff);
}
__device__ __forceinline__ float fkey_decode(int k) {
  if (k == 0x7fffffff || k == (int)0x80000000) return CUDART_NAN_F;
  return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff));
}

__global__ void k_boundary_init(int* keys) {
  if (threadIdx.x < 6) keys[threadIdx.x] = (threadIdx.x & 1) ? fkey_max(-CUDART_INF_F) : fkey_min(CUDART_INF_F);
}
__global__ void k_boundary_final(int* keys) {
  if (threadIdx.x < 6) reinterpret_cast<float*>(keys)[threadIdx.x] = fkey_decode(keys[threadIdx.x]);
}

template <int DT>
__global__ void __launch_bounds__(kThreads) k_boundary(const void* __restrict__ depth, const float* __restrict__ poses,
                                                       const float* __restrict__ intr, int V, int H, int W,
                                                       float depth_scale, Clamp clamp, int* __restrict__ keys) {
  // one CTA works on rows of one frame at a time: blockIdx -> (frame, row-slab)
  const int rows_per_cta = 8;
  const int slabs = (H + rows_per_cta - 1) / rows_per_cta;
  const Divisor ds = make_divisor(depth_scale);
  int mn[3] = {0x7ffffffe, 0x7ffffffe, 0x7ffffffe}, mx[3] = {(int)0x80000001, (int)0x80000001, (int)0x80000001};
  for (int item = blockIdx.x; item < V * slabs; item += gridDim.x) {
    const int f = item / slabs, r0 = (item - f * slabs) * rows_per_cta;
    const Camera cam = load_camera(poses, intr, f);
    const int r1 = min(r0 + rows_per_cta, H);
    if (DT == V3D_DEPTH_U16 && (W & 7) == 0 && W <= 65528 && (reinterpret_cast<uintptr_t>(depth) & 15) == 0) {
      // eight pixels per 128-bit load; group index -> (row, column) by multiplication (index < 2^16)
      if constexpr (DT == V3D_DEPTH_U16) {
        const int gpr = W >> 3;
        const uint32_t magic = gpr > 1 ? 0xffffffffu / (uint32_t)gpr + 1u : 0u;
        for (int i = threadIdx.x; i < (r1 - r0) * gpr; i += kThreads) {
          const int rr = gpr > 1 ? (int)__umulhi((uint32_t)i, magic) : i;
          const int r = r0 + rr, c0 = (i - rr * gpr) << 3;
          const float vm = __fsub_rn((float)r, cam.cy);
          const uint4 q = ld_stream_v4(reinterpret_cast<const uint16_t*>(depth) + ((size_t)f * H + r) * W + c0);
          const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float d = (float)((e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xffffu));
            float p[3];
            unproject_px(cam, d, __fsub_rn((float)(c0 + e), cam.cx), vm, ds, p[0], p[1], p[2]);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const float v = apply_clamp(clamp, p[a], a);
              mn[a] = min(mn[a], fkey_min(v));
              mx[a] = max(mx[a], fkey_max(v));
            }
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < (r1 - r0) * W; i += kThreads) {
        const int r = r0 + i / W, c = i - (i / W) * W;
        float p[3];
        unproject_px(cam, load_depth<DT>(depth, ((size_t)f * H + r) * W + c), __fsub_rn((float)c, cam.cx),
                     __fsub_rn((float)r, cam.cy), ds, p[0], p[1], p[2]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float v = apply_clamp(clamp, p[a], a);
          mn[a] = min(mn[a], fkey_min(v));
          mx[a] = max(mx[a], fkey_max(v));
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o; o >>= 1) {
      mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(keys + 2 * a, mn[a]);
      atomicMax(keys + 2 * a + 1, mx[a]);
    }
  }
}

// ============================================================================================
// 3. sin-3D PE building blocks (v3d_pe.cuh) and the table kernel
// ============================================================================================
template <typename OutT>
__global__ void __launch_bounds__(kThreads) k_pe_table(const float* __restrict__ dim_t, PeGeom g, OutT* __restrict__ table) {
  const size_t total = pe_table_elems(g);
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    const int k = (int)(i % g.Wt);
    const int r = (int)((i / g.Wt) % (g.rows + 1));
    const int s = (int)(i / ((size_t)g.Wt * (g.rows + 1)));
    const int j = ((s * g.nf) & ~7) + k - s * g.nf;
    float v = 0.0f;
    if (s < g.n_slots && j >= 0 && j < g.nf) v = (r == g.rows) ? CUDART_NAN_F : pe_value((float)r, j, dim_t);
    table[i] = Elem<OutT>::from(v);
  }
}


// A whole token row evaluated without the table wherever a slot needs it (non-integer / out-of-range
// coordinates): kept out of line so that the sinf/cosf slow paths do not bloat the streaming loop.
// One warp, one token.  off: n_slots+1 table row offsets (-1 = evaluate directly).
template <typename OutT>
__device__ __noinline__ void stream_pe_direct(const PeGeom g, const float* __restrict__ dim_t, const OutT* __restrict__ table,
                                              const float* pts, const int* off, const OutT* trow, OutT* orow) {
  const int lane = threadIdx.x & 31;
  for (int c8 = lane; c8 < g.cpt; c8 += 32) {
    Chunk<OutT> pe = chunk_zero<OutT>();
#pragma unroll 1
    for (int e = 0; e < 8; ++e) {
      const int ch = c8 * 8 + e;
      const int s = ch / g.nf;
      if (s >= g.n_slots) continue;
      const int j = ch - s * g.nf;
      const int o = off[s];
      pe.v[e] = (o >= 0) ? table[o + (s * g.nf - ((s * g.nf) & ~7)) + j] : Elem<OutT>::from(pe_value(pts[s], j, dim_t));
    }
    if (trow) pe = chunk_add(chunk_ld_stream(trow + c8 * 8), pe);
    chunk_st_stream(orow + c8 * 8, pe);
  }
}

// One pass of the streaming loop.  Work is laid out over PAIRS of tokens: the 2*cpt chunks of two
// consecutive token rows form one linear range L (cpt = 144 -> 288 = 9 x 32: no ragged tail), lane l
// owns L = l0 + 32u for u < U (descriptor and token parity in registers), warps stride the pairs.  Per
// pair all loads (tokens + table rows) are issued before the first add/store -- the compiler cannot
// reorder them itself because tokens may alias out.  GUARD: L may be past 2*cpt.  STRADDLE: chunks may
// touch two slots.  ADD: fused token add.
template <typename OutT, int U, bool GUARD, bool STRADDLE, bool ADD>
__device__ __forceinline__ void stream_pass(int l0, int cpt, int E, int n_tok, int n_off, const OutT* __restrict__ table,
                                            const int* s_off, const int* s_flag, const ChunkDesc* s_desc, const OutT* tokens,
                                            OutT* out) {
  constexpr int kWarps = kThreads / 32;
  const int warp = threadIdx.x >> 5;
  uint32_t sa[U], ka[U], sb[U], kb[U];
  uint32_t rel[U];   // element offset of the chunk relative to the pair's first row
  uint32_t osel[U];  // byte offset of the token's offset list relative to the pair's first list
  int tsel[U];
  bool act[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int L = l0 + 32 * u;
    tsel[u] = L >= cpt;
    const int c8 = L - (tsel[u] ? cpt : 0);
    act[u] = !GUARD || c8 < cpt;
    const uint2 d = act[u] ? lds_u2(s_desc + c8) : make_uint2(0u, 0u);
    sa[u] = (d.x >> 16) * 4u;  // byte offset into the token's offset list
    ka[u] = d.x & 0xffffu;
    sb[u] = (d.y >> 16) * 4u;
    kb[u] = d.y & 0xffffu;
    rel[u] = (uint32_t)(tsel[u] * E + c8 * 8);
    osel[u] = (uint32_t)(tsel[u] * n_off) * 4u;
  }
  const uint32_t off_base = smem_u32(s_off);
  const uint32_t flag_base = smem_u32(s_flag);
  for (int t0 = 2 * warp; t0 < n_tok; t0 += 2 * kWarps) {
    const uint32_t off = off_base + (uint32_t)(t0 * n_off) * 4u;
    const size_t row = (size_t)t0 * E;
    const OutT* tp = ADD ? tokens + row : nullptr;
    OutT* op = out + row;
    Chunk<OutT> tv[U], pe[U];
    bool go[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      go[u] = act[u] && (t0 + tsel[u] < n_tok);
      if (go[u]) {
        int flag;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(flag) : "r"(flag_base + 4u * (uint32_t)(t0 + tsel[u])));
        go[u] = flag == 0;
      }
      if (go[u]) {
        if (ADD) tv[u] = chunk_ld_stream(tp + rel[u]);
        int o;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(o) : "r"(off + osel[u] + sa[u]));
        pe[u] = chunk_ldg(table + (uint32_t)o + ka[u]);
        if (STRADDLE) {
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(o) : "r"(off + osel[u] + sb[u]));
          chunk_or(pe[u], chunk_ldg(table + (uint32_t)o + kb[u]));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (go[u]) chunk_st_stream(op + rel[u], ADD ? chunk_add(tv[u], pe[u]) : pe[u]);
    }
  }
}

template <typename OutT, int U, bool STRADDLE, bool ADD>
__device__ __forceinline__ void stream_rows_u(int cpt, int E, int n_tok, int n_off, const OutT* __restrict__ table, const int* s_off,
                                              const int* s_flag, const ChunkDesc* s_desc, const OutT* tokens, OutT* out) {
  const int lane = threadIdx.x & 31;
  const int total = 2 * cpt;
  const int full = total / (32 * U) * (32 * U);
  for (int l0 = 0; l0 < full; l0 += 32 * U)
    stream_pass<OutT, U, false, STRADDLE, ADD>(l0 + lane, cpt, E, n_tok, n_off, table, s_off, s_flag, s_desc, tokens, out);
  if (full < total) stream_pass<OutT, U, true, STRADDLE, ADD>(full + lane, cpt, E, n_tok, n_off, table, s_off, s_flag, s_desc, tokens, out);
}

template <typename OutT, bool STRADDLE, bool ADD>
__device__ __forceinline__ void stream_rows(int cpt, int E, int n_tok, int n_off, const OutT* __restrict__ table, const int* s_off,
                                            const int* s_flag, const ChunkDesc* s_desc, const OutT* tokens, OutT* out) {
  // pick the batch size that leaves no ragged pass: 2*cpt/32 steps (9 for E=1152, 28 for E=3584)
  if ((2 * cpt) % 96 == 0) stream_rows_u<OutT, 3, STRADDLE, ADD>(cpt, E, n_tok, n_off, table, s_off, s_flag, s_desc, tokens, out);
  else stream_rows_u<OutT, 2, STRADDLE, ADD>(cpt, E, n_tok, n_off, table, s_off, s_flag, s_desc, tokens, out);
}

// The streaming loop when no token of the CTA needs direct evaluation (the normal case with
// discretised coordinates): same token-pair layout as stream_pass, stripped of every per-token test.
// Per chunk: one shared load (row offset), two 128-bit global loads, four packed adds, one store.
template <typename OutT, int U, bool STRADDLE, bool ADD>
__device__ __forceinline__ void stream_pairs_clean(int l0, int cpt, int E, int n_pairs, int n_off, const OutT* __restrict__ table,
                                                   const int* s_off, const ChunkDesc* s_desc, const OutT* tokens, OutT* out) {
  constexpr int kWarps = kThreads / 32;
  const int warp = threadIdx.x >> 5;
  uint32_t offa[U], offb[U], ka[U], kb[U], rel[U];
  bool two[U];  // this chunk column touches a second slot (only the 2-3 columns at slot boundaries do)
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int L = l0 + 32 * u;
    const int ts = L >= cpt;
    const int c8 = L - (ts ? cpt : 0);
    const uint2 d = lds_u2(s_desc + c8);
    offa[u] = (uint32_t)(ts * n_off + (int)(d.x >> 16)) * 4u;
    ka[u] = d.x & 0xffffu;
    offb[u] = (uint32_t)(ts * n_off + (int)(d.y >> 16)) * 4u;
    kb[u] = d.y & 0xffffu;
    two[u] = STRADDLE && (int)(d.y >> 16) != n_off - 1;
    rel[u] = (uint32_t)(ts * E + c8 * 8);
  }
  uint32_t off = smem_u32(s_off) + (uint32_t)(warp * 2 * n_off) * 4u;
  const uint32_t off_step = (uint32_t)(kWarps * 2 * n_off) * 4u;
  const size_t row_step = (size_t)kWarps * 2 * E;
  const OutT* tp = ADD ? tokens + (size_t)warp * 2 * E : nullptr;
  OutT* op = out + (size_t)warp * 2 * E;
  for (int pair = warp; pair < n_pairs; pair += kWarps, off += off_step, op += row_step) {
    Chunk<OutT> tv[U], pe[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ADD) tv[u] = chunk_ld_stream(tp + rel[u]);
      int o;
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(o) : "r"(off + offa[u]));
      pe[u] = chunk_ldg(table + ((uint32_t)o + ka[u]));
      if (STRADDLE) {
        if (two[u]) {
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(o) : "r"(off + offb[u]));
          chunk_or(pe[u], chunk_ldg(table + ((uint32_t)o + kb[u])));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) chunk_st_stream(op + rel[u], ADD ? chunk_add(tv[u], pe[u]) : pe[u]);
    if (ADD) tp += row_step;
  }
}

// one token row, lanes stride its chunks (the odd last token of a CTA)
template <typename OutT, bool STRADDLE, bool ADD>
__device__ __forceinline__ void stream_single_clean(int tok, int cpt, int E, int n_off, const OutT* __restrict__ table,
                                                    const int* s_off, const ChunkDesc* s_desc, const OutT* tokens, OutT* out) {
  const int lane = threadIdx.x & 31;
  const uint32_t off = smem_u32(s_off) + (uint32_t)(tok * n_off) * 4u;
  const OutT* tp = ADD ? tokens + (size_t)tok * E : nullptr;
  OutT* op = out + (size_t)tok * E;
#pragma unroll 2
  for (int c8 = lane; c8 < cpt; c8 += 32) {
    const uint2 d = lds_u2(s_desc + c8);
    Chunk<OutT> tv;
    if (ADD) tv = chunk_ld_stream(tp + c8 * 8);
    int o;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(o) : "r"(off + (d.x >> 16) * 4u));
    Chunk<OutT> pe = chunk_ldg(table + ((uint32_t)o + (d.x & 0xffffu)));
    if (STRADDLE) {
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(o) : "r"(off + (d.y >> 16) * 4u));
      chunk_or(pe, chunk_ldg(table + ((uint32_t)o + (d.y & 0xffffu))));
    }
    chunk_st_stream(op + c8 * 8, ADD ? chunk_add(tv, pe) : pe);
  }
}

// A pair of token rows per warp, lanes stride the chunk columns: the column's descriptor is read once for both rows
// and the inner loop carries no per-pair address arithmetic.  For CTAs with few pairs (7 for the 14x14 grid) whose
// slots do not start on 16-byte boundaries (nf = 1194 for E = 3584).
template <typename OutT, bool STRADDLE, bool ADD>
__device__ __forceinline__ void stream_pair_rows(int pair, int cpt, int E, int n_off, const OutT* __restrict__ table, const int* s_off,
                                                 const ChunkDesc* s_desc, const OutT* tokens, OutT* out) {
  const int lane = threadIdx.x & 31;
  const uint32_t off0 = smem_u32(s_off) + (uint32_t)(2 * pair * n_off) * 4u, off1 = off0 + (uint32_t)n_off * 4u;
  const OutT* t0 = ADD ? tokens + (size_t)(2 * pair) * E : nullptr;
  const OutT* t1 = ADD ? t0 + E : nullptr;
  OutT* o0 = out + (size_t)(2 * pair) * E;
  OutT* o1 = o0 + E;
  const uint32_t none = (uint32_t)(n_off - 1);
#pragma unroll 2
  for (int c8 = lane; c8 < cpt; c8 += 32) {
    const uint2 d = lds_u2(s_desc + c8);
    Chunk<OutT> tv0, tv1;
    if (ADD) {
      tv0 = chunk_ld_stream(t0 + c8 * 8);
      tv1 = chunk_ld_stream(t1 + c8 * 8);
    }
    const uint32_t sa = (d.x >> 16) * 4u, ka = d.x & 0xffffu;
    int a0, a1;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a0) : "r"(off0 + sa));
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a1) : "r"(off1 + sa));
    Chunk<OutT> pe0 = chunk_ldg(table + ((uint32_t)a0 + ka));
    Chunk<OutT> pe1 = chunk_ldg(table + ((uint32_t)a1 + ka));
    if (STRADDLE) {
      if ((d.y >> 16) != none) {  // only the columns at slot boundaries touch a second slot
        const uint32_t sb = (d.y >> 16) * 4u, kb = d.y & 0xffffu;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a0) : "r"(off0 + sb));
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a1) : "r"(off1 + sb));
        chunk_or(pe0, chunk_ldg(table + ((uint32_t)a0 + kb)));
        chunk_or(pe1, chunk_ldg(table + ((uint32_t)a1 + kb)));
      }
    }
    chunk_st_stream(o0 + c8 * 8, ADD ? chunk_add(tv0, pe0) : pe0);
    chunk_st_stream(o1 + c8 * 8, ADD ? chunk_add(tv1, pe1) : pe1);
  }
}

template <typename OutT, bool STRADDLE, bool ADD, bool PAIR_ROWS = false>
__device__ __forceinline__ void stream_rows_clean(int cpt, int E, int n_tok, int n_off, const OutT* __restrict__ table,
                                                  const int* s_off, const ChunkDesc* s_desc, const OutT* tokens, OutT* out) {
  constexpr int kWarps = kThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_pairs = n_tok >> 1;
  const int total = 2 * cpt;
  if (STRADDLE && PAIR_ROWS) {
    // slots that do not start on a 16-byte boundary (nf = 1194 for E = 3584): the pair layout spends ~60 instructions
    // per chunk on descriptor and address arithmetic; one warp per token with per-chunk descriptors from shared memory
    for (int pair = warp; pair < n_pairs; pair += kWarps)
      stream_pair_rows<OutT, STRADDLE, ADD>(pair, cpt, E, n_off, table, s_off, s_desc, tokens, out);
    if ((n_tok & 1) && warp == (n_pairs % kWarps))
      stream_single_clean<OutT, STRADDLE, ADD>(n_tok - 1, cpt, E, n_off, table, s_off, s_desc, tokens, out);
    return;
  }
  if (total % 96 == 0) {
    for (int l0 = 0; l0 < total; l0 += 96)
      stream_pairs_clean<OutT, 3, STRADDLE, ADD>(l0 + lane, cpt, E, n_pairs, n_off, table, s_off, s_desc, tokens, out);
  } else if (total % 64 == 0) {
    for (int l0 = 0; l0 < total; l0 += 64)
      stream_pairs_clean<OutT, 2, STRADDLE, ADD>(l0 + lane, cpt, E, n_pairs, n_off, table, s_off, s_desc, tokens, out);
  } else {
    for (int tok = warp; tok < 2 * n_pairs; tok += kWarps)
      stream_single_clean<OutT, STRADDLE, ADD>(tok, cpt, E, n_off, table, s_off, s_desc, tokens, out);
  }
  if ((n_tok & 1) && warp == (n_pairs % kWarps))
    stream_single_clean<OutT, STRADDLE, ADD>(n_tok - 1, cpt, E, n_off, table, s_off, s_desc, tokens, out);
}

// PE rows of `n_tok` tokens; must be called by every thread of the CTA.
//   s_pts  [n_tok][n_slots]    coordinates the PE is evaluated on
//   s_off  [n_tok][n_slots+1]  element offset of each slot's table row (-1 = evaluate directly);
//                              entry n_slots is the zero row
//   s_flag [n_tok]             1 = some slot of the token needs direct evaluation
// PAIR_ROWS selects stream_pair_rows for slots that do not start on 16-byte boundaries.  It is a separate kernel
// instantiation so that the code generated for the aligned geometries (E = 1152) is exactly what it was tuned to be.
template <typename OutT, bool PAIR_ROWS = false>
__device__ __forceinline__ void stream_pe(const PeGeom& g, const float* __restrict__ dim_t, const OutT* __restrict__ table,
                                          const float* s_pts, const int* s_off, const int* s_flag, const ChunkDesc* s_desc,
                                          int n_tok, const OutT* tokens, OutT* out) {
  const int warp = threadIdx.x >> 5;
  constexpr int kWarps = kThreads / 32;
  const int n_off = g.n_slots + 1;
  if ((g.E & 7) == 0) {
    int mine = 0;
    for (int i = threadIdx.x; i < n_tok; i += kThreads) mine |= s_flag[i];
    const int any_direct = __syncthreads_or(mine);
    const bool straddle = (g.nf & 7) != 0;
    if (table != nullptr && !any_direct) {
      if (tokens) {
        if (straddle) stream_rows_clean<OutT, true, true, PAIR_ROWS>(g.cpt, g.E, n_tok, n_off, table, s_off, s_desc, tokens, out);
        else stream_rows_clean<OutT, false, true>(g.cpt, g.E, n_tok, n_off, table, s_off, s_desc, tokens, out);
      } else {
        if (straddle) stream_rows_clean<OutT, true, false, PAIR_ROWS>(g.cpt, g.E, n_tok, n_off, table, s_off, s_desc, tokens, out);
        else stream_rows_clean<OutT, false, false>(g.cpt, g.E, n_tok, n_off, table, s_off, s_desc, tokens, out);
      }
      return;
    }
    if (table != nullptr) {
      if (tokens) {
        if (straddle) stream_rows<OutT, true, true>(g.cpt, g.E, n_tok, n_off, table, s_off, s_flag, s_desc, tokens, out);
        else stream_rows<OutT, false, true>(g.cpt, g.E, n_tok, n_off, table, s_off, s_flag, s_desc, tokens, out);
      } else {
        if (straddle) stream_rows<OutT, true, false>(g.cpt, g.E, n_tok, n_off, table, s_off, s_flag, s_desc, tokens, out);
        else stream_rows<OutT, false, false>(g.cpt, g.E, n_tok, n_off, table, s_off, s_flag, s_desc, tokens, out);
      }
    }
    for (int tok = warp; tok < n_tok; tok += kWarps) {
      if (s_flag[tok])
        stream_pe_direct<OutT>(g, dim_t, table, s_pts + tok * g.n_slots, s_off + tok * n_off,
                               tokens ? tokens + (size_t)tok * g.E : nullptr, out + (size_t)tok * g.E);
    }
  } else {
    // rows that are not a whole number of 16-byte chunks: element-wise (rare configuration)
    for (int i = threadIdx.x; i < n_tok * g.E; i += kThreads) {
      const int tok = i / g.E, ch = i - tok * g.E;
      const int s = ch / g.nf;
      float v = 0.0f;
      if (s < g.n_slots) {
        const int o = s_off[tok * n_off + s];
        const int j = ch - s * g.nf;
        v = (o >= 0) ? Elem<OutT>::to(table[o + (s * g.nf - ((s * g.nf) & ~7)) + j]) : Elem<OutT>::to(Elem<OutT>::from(pe_value(s_pts[tok * g.n_slots + s], j, dim_t)));
      }
      const size_t oidx = (size_t)tok * g.E + ch;
      if (tokens) v = __fadd_rn(Elem<OutT>::to(tokens[oidx]), v);
      out[oidx] = Elem<OutT>::from(v);
    }
  }
}

// stand-alone PE: llava/model/position_encoding.py:17-49 (+ llava_arch.py:515 add)
static constexpr int kPeTokPerCta = 16;

template <typename OutT>
__global__ void __launch_bounds__(kThreads) k_pe_sine3d(const float* __restrict__ coords, size_t n_tokens, PeGeom g,
                                                        const float* __restrict__ dim_t, const OutT* __restrict__ table,
                                                        const OutT* tokens, OutT* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChunkDesc* s_desc = reinterpret_cast<ChunkDesc*>(smem_raw);
  float* s_pts = reinterpret_cast<float*>(s_desc + g.cpt);
  int* s_off = reinterpret_cast<int*>(s_pts + kPeTokPerCta * g.n_slots);
  int* s_flag = s_off + kPeTokPerCta * (g.n_slots + 1);
  const int n_off = g.n_slots + 1;
  for (int i = threadIdx.x; i < g.cpt; i += kThreads) s_desc[i] = make_chunk_desc(i, g);
  for (size_t t0 = (size_t)blockIdx.x * kPeTokPerCta; t0 < n_tokens; t0 += (size_t)gridDim.x * kPeTokPerCta) {
    const int n_tok = (int)min((size_t)kPeTokPerCta, n_tokens - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < n_tok; i += kThreads) {
      int flag = 0;
      for (int sl = 0; sl < g.n_slots; ++sl) {
        const float x = __ldg(coords + (t0 + i) * g.n_slots + sl);
        const int o = pe_row_offset(x, sl, g, table != nullptr);
        s_pts[i * g.n_slots + sl] = x;
        s_off[i * n_off + sl] = o;
        flag |= (o < 0);
      }
      s_off[i * n_off + g.n_slots] = pe_zero_row_offset(g);
      s_flag[i] = flag;
    }
    __syncthreads();
    stream_pe<OutT>(g, dim_t, table, s_pts, s_off, s_flag, s_desc, n_tok, tokens ? tokens + t0 * g.E : nullptr, out + t0 * g.E);
  }
}

// ============================================================================================
// 4. the fused kernel: one CTA per (frame, patch row)
// ============================================================================================
// xyz of one resampled row live in three planes of kPlane floats (x | y | z), double buffered.  The
// plane stride is a compile-time constant so that the three stores of a pixel and the chain loads use
// immediate offsets; 520 % 32 == 8 keeps the planes on different banks.
#ifndef V3D_FUSED_MIN_BLOCKS
#define V3D_FUSED_MIN_BLOCKS 4
#endif
static constexpr int kPlane = 520;  // >= 2 * kThreads columns, even

__host__ __device__ inline size_t fused_smem_bytes(const FusedParams& p, bool stage_depth) {
  const size_t n_slots = (size_t)3 * p.n_points;
  size_t b = 16;                                              // mbarrier
  if (stage_depth) b += (size_t)p.patch * p.span;             // depth rows
  b += (size_t)4 * 3 * kPlane * sizeof(float);                // xyz rows: 2 buffers x 2 rows
  b += (size_t)p.patch * 2 * sizeof(float2);                  // {row - cy, fast flag} and source row per patch row
  b += (size_t)p.grid * (n_slots * sizeof(float) + (n_slots + 1) * sizeof(int) + sizeof(int));  // PE coords, row offsets, flags
  b = (b + 7) & ~(size_t)7;
  if (p.pe_mode != V3D_PE_NONE) b += (size_t)p.g.cpt * sizeof(ChunkDesc);
  return b;
}

// sequential sum of PW consecutive floats (ATen's avg_pool2d order); PW = 0: runtime width
template <int PW>
__device__ __forceinline__ float chain_add(float acc, const float* src, int P) {
  if constexpr (PW > 0) {
    if constexpr ((PW & 1) == 0) {
      // even width: rows of a patch start 8-byte aligned (kPlane is even) -> 64-bit shared loads
      if ((smem_u32(src) & 7) == 0) {
#pragma unroll
        for (int dx = 0; dx < PW; dx += 2) {
          const float2 v = *reinterpret_cast<const float2*>(src + dx);
          acc = __fadd_rn(__fadd_rn(acc, v.x), v.y);
        }
        return acc;
      }
    }
#pragma unroll
    for (int dx = 0; dx < PW; ++dx) acc = __fadd_rn(acc, src[dx]);
  } else {
#pragma unroll 4
    for (int dx = 0; dx < P; ++dx) acc = __fadd_rn(acc, src[dx]);
  }
  return acc;
}

__device__ __forceinline__ void sts_xyz(uint32_t addr, float x, float y, float z) {
  asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+%4], %2;\n\tst.shared.f32 [%0+%5], %3;" ::"r"(addr), "f"(x), "f"(y),
               "f"(z), "n"(kPlane * 4), "n"(2 * kPlane * 4)
               : "memory");
}

// State of one chain thread (axis a, patch j): running sum / max / max of negatives.
struct ChainAcc {
  float acc, mx, mneg;
};

__device__ __forceinline__ void chain_row(ChainAcc& c, const float* src, int P, int pool_mode) {
  if (pool_mode == V3D_POOL_AVG) {
    if (P == 14) c.acc = chain_add<14>(c.acc, src, P);
    else if (P == 27) c.acc = chain_add<27>(c.acc, src, P);
    else c.acc = chain_add<0>(c.acc, src, P);
  } else {
    for (int dx = 0; dx < P; ++dx) {
      const float v = src[dx];
      if (v > c.mx || v != v) c.mx = v;
      const float nv = -v;
      if (nv > c.mneg || nv != nv) c.mneg = nv;
    }
  }
}

// The row loop when EVERY pixel of the CTA qualifies for the branch-free path: uint16 depth staged by
// TMA, rigid pose, depth scale / fx / fy and all (u-cx), (v-cy) well inside the normal range.  Each
// dividend of the three divisions is then either exactly zero (d == 0, where a*r is the correctly
// signed zero) or inside div_exact's fast range, and every camera-space value is finite, so no range
// test, no fallback and no w-division remain in the loop.
template <bool ROUND>
__device__ __forceinline__ void pool_rows_fast(const Camera& cam, const Divisor& ds, int P, uint32_t span, int n_px,
                                               const float um[2], const uint32_t d_addr[2], const uint32_t w_addr[2],
                                               const float2* s_vm, bool chain, const float* chain_src, int pool_mode,
                                               const Clamp& clamp, ChainAcc& c, const unsigned char* pf_src, size_t pf_bytes,
                                               int coord_dtype) {
  // the token rows are pulled into L2 ~8 row pairs before the streaming phase needs them: early enough to hide
  // DRAM latency, late enough that the output stream of the co-resident CTAs has not evicted them again
#ifndef V3D_PF_LEAD
#define V3D_PF_LEAD 8
#endif
  const int pf_dy = (P > V3D_PF_LEAD) ? ((P - V3D_PF_LEAD) & ~1) : 0;
  // two resampled rows per barrier: the row buffers hold 2 x 2 rows, halving the number of block-wide
  // synchronisations (the chain threads add row dy, then row dy+1, so the summation order is unchanged)
  uint32_t row_off = 0;
  for (int dy = 0; dy < P; dy += 2) {
    const uint32_t bsel = ((dy >> 1) & 1) ? (uint32_t)(6 * kPlane * 4) : 0u;
    const int n_rows = (dy + 1 < P) ? 2 : 1;
    if (dy == pf_dy && pf_src != nullptr && threadIdx.x == 32) {
      for (size_t o = 0; o < pf_bytes; o += 32768) tma_prefetch_l2(pf_src + o, (uint32_t)min((size_t)32768, pf_bytes - o));
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      if (rr < n_rows) {
        const float vm = s_vm[dy + rr].x;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k < n_px) {
            const uint32_t raw = lds_u16(d_addr[k] + row_off);
            const float d = (float)raw;
            const bool nz = raw != 0;
            float q0 = __fmul_rn(d, ds.r);
            float q1 = __fmaf_rn(ds.r, __fmaf_rn(-ds.b, q0, d), q0);
            const float cz = nz ? q1 : q0;
            const float ax = __fmul_rn(um[k], cz);
            q0 = __fmul_rn(ax, cam.fx.r);
            q1 = __fmaf_rn(cam.fx.r, __fmaf_rn(-cam.fx.b, q0, ax), q0);
            const float cx_ = nz ? q1 : q0;
            const float ay = __fmul_rn(vm, cz);
            q0 = __fmul_rn(ay, cam.fy.r);
            q1 = __fmaf_rn(cam.fy.r, __fmaf_rn(-cam.fy.b, q0, ay), q0);
            const float cy_ = nz ? q1 : q0;
            float x = __fadd_rn(__fmaf_rn(cam.P[2], cz, __fmaf_rn(cam.P[1], cy_, __fmul_rn(cam.P[0], cx_))), cam.P[3]);
            float y = __fadd_rn(__fmaf_rn(cam.P[6], cz, __fmaf_rn(cam.P[5], cy_, __fmul_rn(cam.P[4], cx_))), cam.P[7]);
            float z = __fadd_rn(__fmaf_rn(cam.P[10], cz, __fmaf_rn(cam.P[9], cy_, __fmul_rn(cam.P[8], cx_))), cam.P[11]);
            if (clamp.on) {
              x = apply_clamp(clamp, x, 0);
              y = apply_clamp(clamp, y, 1);
              z = apply_clamp(clamp, z, 2);
            }
            if constexpr (ROUND) {
              x = round_to_dtype(x, coord_dtype);
              y = round_to_dtype(y, coord_dtype);
              z = round_to_dtype(z, coord_dtype);
            }
            sts_xyz(w_addr[k] + bsel + (uint32_t)(rr * 3 * kPlane * 4), x, y, z);
          }
        }
        row_off += span;
      }
    }
    __syncthreads();
    if (chain) {
      const float* src = chain_src + (((dy >> 1) & 1) ? 6 * kPlane : 0);
      chain_row(c, src, P, pool_mode);
      if (n_rows == 2) chain_row(c, src + 3 * kPlane, P, pool_mode);
    }
  }
}

template <int DT, bool FROM_DEPTH, bool USE_TMA, typename OutT, bool PAIR_ROWS = false>
__global__ void __launch_bounds__(kThreads, V3D_FUSED_MIN_BLOCKS) k_fused(const FusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kEsz = (DT == V3D_DEPTH_U16) ? 2 : 4;
  const int t = threadIdx.x;
  const int f = blockIdx.x / p.grid, gi = blockIdx.x - f * p.grid;
  const int G = p.grid, P = p.patch, n_slots = 3 * p.n_points, n_off = n_slots + 1;

  // ---- shared memory carve-up
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  unsigned char* s_depth = smem_raw + 16;
  float* s_row = reinterpret_cast<float*>(s_depth + ((FROM_DEPTH && USE_TMA) ? (size_t)P * p.span : 0));
  float2* s_vm = reinterpret_cast<float2*>(s_row + 4 * 3 * kPlane);  // {row - cy, unused}
  int* s_srow = reinterpret_cast<int*>(s_vm + P);
  float* s_pts = reinterpret_cast<float*>(s_srow + 2 * P);
  int* s_off = reinterpret_cast<int*>(s_pts + G * n_slots);
  int* s_flag = s_off + G * n_off;
  ChunkDesc* s_desc = reinterpret_cast<ChunkDesc*>((reinterpret_cast<uintptr_t>(s_flag + G) + 7) & ~(uintptr_t)7);

  // ---- the token rows this CTA will add the PE to are independent of everything computed below: the TMA unit
  // pulls them into L2 ahead of the streaming phase (issued inside the row loop on the fast path, see
  // pool_rows_fast, otherwise right before the generic loop)
  const unsigned char* pf_src = nullptr;
  size_t pf_bytes = 0;
  if (p.pe_mode != V3D_PE_NONE && p.tokens_in != nullptr) {
    const size_t row_bytes = (size_t)G * p.g.E * sizeof(OutT);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.tokens_in) + ((size_t)f * G + gi) * row_bytes;
    if ((row_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      pf_src = src;
      pf_bytes = row_bytes;
    }
  }
  // ---- stage the P depth rows of this patch row with the TMA unit
  if constexpr (FROM_DEPTH && USE_TMA) {
    if (t == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (t < 32) {
      if (t == 0) mbar_arrive_expect_tx(bar, (uint32_t)(P * p.span));
      __syncwarp();
      for (int dy = t; dy < P; dy += 32) {
        const int sr = p.resampled ? gi * P + dy : __ldg(p.row_map + gi * P + dy);
        const unsigned char* src =
            reinterpret_cast<const unsigned char*>(p.depth) + ((size_t)f * p.H + sr) * p.W * kEsz + p.col_lo_byte;
        tma_bulk_g2s(s_depth + (size_t)dy * p.span, src, (uint32_t)p.span, bar);
      }
    }
  }
  if (p.pe_mode != V3D_PE_NONE) {
    for (int i = t; i < p.g.cpt; i += kThreads) s_desc[i] = make_chunk_desc(i, p.g);
  }

  Camera cam;
  Divisor ds;
  bool fast_ok = false;  // this thread's pixels qualify for pool_rows_fast
  if constexpr (FROM_DEPTH) {
    cam = load_camera(p.poses, p.intr, f);
    ds = make_divisor(p.depth_scale);
    if constexpr (DT == V3D_DEPTH_U16 && USE_TMA) fast_ok = cam.rigid && ds.ok && cam.fx.ok && cam.fy.ok && ds.b > 0.f;
  }
  for (int dy = t; dy < P; dy += kThreads) {
    const int sr = FROM_DEPTH ? __ldg(p.row_map + gi * P + dy) : gi * P + dy;
    const float vm = FROM_DEPTH ? __fsub_rn((float)sr, cam.cy) : 0.0f;
    s_srow[dy] = (FROM_DEPTH && p.resampled) ? gi * P + dy : sr;  // the row the sample is stored in
    s_vm[dy] = make_float2(vm, 0.0f);
    fast_ok = fast_ok && fabsf(vm) >= 0x1p-20f && fabsf(vm) <= 0x1p20f;
  }

  // per-thread column constants: thread t owns dst column t and (second round, taken from the top of
  // the block so that the chain warps at the bottom get less unprojection work) 511 - t
  int col[2];
  float um[2];
  size_t src_off[2];     // FROM_DEPTH: byte offset inside a staged row / element offset inside a global row
  uint32_t w_addr[2];    // shared address of this column's x slot in buffer 0
  uint32_t d_addr[2];    // shared address of this column's depth sample in staged row 0
  col[0] = (t < p.Sp) ? t : -1;
  col[1] = (kThreads + (kThreads - 1 - t) < p.Sp) ? kThreads + (kThreads - 1 - t) : -1;
  if (p.Sp < kThreads) fast_ok = false;  // the fast loop assumes every thread owns column t
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    um[k] = 0.f;
    src_off[k] = 0;
    w_addr[k] = smem_u32(s_row + (col[k] >= 0 ? col[k] : 0));
    d_addr[k] = smem_u32(s_depth);
    if (col[k] >= 0) {
      const int sc = FROM_DEPTH ? __ldg(p.col_map + col[k]) : col[k];
      if constexpr (FROM_DEPTH) {
        um[k] = __fsub_rn((float)sc, cam.cx);
        fast_ok = fast_ok && fabsf(um[k]) >= 0x1p-20f && fabsf(um[k]) <= 0x1p20f;
        const int stored = p.resampled ? col[k] : sc;  // the column the sample is stored in
        src_off[k] = USE_TMA ? (size_t)(stored * kEsz - p.col_lo_byte) : (size_t)stored;
        d_addr[k] += (uint32_t)src_off[k];
      } else {
        src_off[k] = (size_t)sc * 3;
      }
    }
  }
  const int all_fast = __syncthreads_and(fast_ok);  // also makes s_srow / s_vm / s_desc visible
  if constexpr (FROM_DEPTH && USE_TMA) mbar_wait(bar, 0);

  const bool is_sample = p.pool_mode >= V3D_POOL_SAMPLE9;
  if (!is_sample) {
    // ---- chains: thread q < 3G owns (axis a = q / G, patch j = q % G)
    const bool chain = t < 3 * G;
    const int a = chain ? t / G : 0, j = chain ? t - a * G : 0;
    ChainAcc ca = {0.0f, -CUDART_INF_F, -CUDART_INF_F};
    const float* chain_src = s_row + a * kPlane + j * P;
    if (all_fast) {
      if (p.coord_dtype == V3D_F32)
        pool_rows_fast<false>(cam, ds, P, (uint32_t)p.span, col[1] >= 0 ? 2 : 1, um, d_addr, w_addr, s_vm, chain, chain_src,
                              p.pool_mode, p.clamp, ca, pf_src, pf_bytes, V3D_F32);
      else
        pool_rows_fast<true>(cam, ds, P, (uint32_t)p.span, col[1] >= 0 ? 2 : 1, um, d_addr, w_addr, s_vm, chain, chain_src,
                             p.pool_mode, p.clamp, ca, pf_src, pf_bytes, p.coord_dtype);
    } else {
      if (pf_src != nullptr && t == 32) {
        for (size_t o = 0; o < pf_bytes; o += 32768) tma_prefetch_l2(pf_src + o, (uint32_t)min((size_t)32768, pf_bytes - o));
      }
      uint32_t row_off = 0;  // byte offset of staged depth row dy
      for (int dy = 0; dy < P; ++dy, row_off += (uint32_t)p.span) {
        const uint32_t bsel = (dy & 1) ? (uint32_t)(3 * kPlane * 4) : 0u;
        const float vm = s_vm[dy].x;
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
          if (col[k] < 0) continue;
          float x, y, z;
          if constexpr (FROM_DEPTH) {
            float d;
            if constexpr (USE_TMA) {
              const unsigned char* sp = s_depth + row_off + src_off[k];
              if constexpr (DT == V3D_DEPTH_U16) d = (float)*reinterpret_cast<const uint16_t*>(sp);
              else if constexpr (DT == V3D_DEPTH_I32) d = (float)*reinterpret_cast<const int32_t*>(sp);
              else d = *reinterpret_cast<const float*>(sp);
            } else {
              d = load_depth<DT>(p.depth, ((size_t)f * p.H + s_srow[dy]) * p.W + src_off[k]);
            }
            unproject_px(cam, d, um[k], vm, ds, x, y, z);
            if (p.clamp.on) {
              x = apply_clamp(p.clamp, x, 0);
              y = apply_clamp(p.clamp, y, 1);
              z = apply_clamp(p.clamp, z, 2);
            }
            if (p.coord_dtype != V3D_F32) {
              x = round_to_dtype(x, p.coord_dtype);
              y = round_to_dtype(y, p.coord_dtype);
              z = round_to_dtype(z, p.coord_dtype);
            }
          } else {
            const float* src = p.coords_in + ((size_t)f * p.S + s_srow[dy]) * p.S * 3 + src_off[k];
            x = __ldg(src);
            y = __ldg(src + 1);
            z = __ldg(src + 2);
          }
          sts_xyz(w_addr[k] + bsel, x, y, z);
        }
        __syncthreads();
        if (chain) chain_row(ca, chain_src + ((dy & 1) ? 3 * kPlane : 0), P, p.pool_mode);
      }
    }
    const float acc = ca.acc, mx = ca.mx, mneg = ca.mneg;
    if (chain) {
      const size_t tok = ((size_t)f * G + gi) * G + j;
      const bool have_table = p.table != nullptr;
      if (p.pool_mode == V3D_POOL_AVG) {
        const float v = round_to_dtype(__fdiv_rn(acc, (float)(P * P)), p.coord_dtype);
        const float pv = p.disc.on ? discretize1(p.disc, v, a) : v;
        if (p.pooled_out) p.pooled_out[tok * 3 + a] = v;
        if (p.pe_coords_out) p.pe_coords_out[tok * 3 + a] = pv;
        s_pts[j * n_slots + a] = pv;
        s_off[j * n_off + a] = pe_row_offset(pv, a, p.g, have_table);
      } else {  // minmax: point 0 = min, point 1 = max (llava_arch.py:237)
        const float v2[2] = {-mneg, mx};
#pragma unroll
        for (int pt = 0; pt < 2; ++pt) {
          const float pv = p.disc.on ? discretize1(p.disc, v2[pt], a) : v2[pt];
          if (p.pooled_out) p.pooled_out[(tok * 2 + pt) * 3 + a] = v2[pt];
          if (p.pe_coords_out) p.pe_coords_out[(tok * 2 + pt) * 3 + a] = pv;
          s_pts[j * n_slots + pt * 3 + a] = pv;
          s_off[j * n_off + pt * 3 + a] = pe_row_offset(pv, pt * 3 + a, p.g, have_table);
        }
      }
    }
  } else {
    // ---- sample modes (llava_arch.py:241-257): offsets {4,13,22}^2 of the 27-pixel window
    const int np = p.n_points;
    for (int i = t; i < G * np; i += kThreads) {
      const int j = i / np, k = i - j * np;
      const int q9 = (np == 9) ? k : (np == 5 ? 2 * k : 4);
      const int r = gi * P + 4 + 9 * (q9 / 3), c = j * P + 4 + 9 * (q9 % 3);
      float v[3];
      if constexpr (FROM_DEPTH) {
        const int sr = __ldg(p.row_map + r), sc = __ldg(p.col_map + c);
        const float d = load_depth<DT>(p.depth, ((size_t)f * p.H + (p.resampled ? r : sr)) * p.W + (p.resampled ? c : sc));
        unproject_px(cam, d, __fsub_rn((float)sc, cam.cx), __fsub_rn((float)sr, cam.cy), ds, v[0], v[1], v[2]);
#pragma unroll
        for (int a = 0; a < 3; ++a) v[a] = round_to_dtype(apply_clamp(p.clamp, v[a], a), p.coord_dtype);
      } else {
        const float* src = p.coords_in + (((size_t)f * p.S + r) * p.S + c) * 3;
        v[0] = __ldg(src);
        v[1] = __ldg(src + 1);
        v[2] = __ldg(src + 2);
      }
      const size_t tok = ((size_t)f * G + gi) * G + j;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float pv = p.disc.on ? discretize1(p.disc, v[a], a) : v[a];
        if (p.pooled_out) p.pooled_out[(tok * np + k) * 3 + a] = v[a];
        if (p.pe_coords_out) p.pe_coords_out[(tok * np + k) * 3 + a] = pv;
        s_pts[j * n_slots + k * 3 + a] = pv;
        s_off[j * n_off + k * 3 + a] = pe_row_offset(pv, k * 3 + a, p.g, p.table != nullptr);
      }
    }
  }
  if (p.pe_mode == V3D_PE_NONE) return;
  __syncthreads();
  // per-token bookkeeping for the streaming loop: zero-row offset, "needs direct evaluation" flag
  for (int j = t; j < G; j += kThreads) {
    int flag = 0;
    for (int sl = 0; sl < n_slots; ++sl) flag |= (s_off[j * n_off + sl] < 0);
    s_off[j * n_off + n_slots] = pe_zero_row_offset(p.g);
    s_flag[j] = flag;
  }
  __syncthreads();

  // ---- stream the PE rows of this patch row (G tokens x E channels), fused token add
  const size_t row0 = ((size_t)f * G + gi) * G;
  const OutT* tokens = reinterpret_cast<const OutT*>(p.tokens_in);
  stream_pe<OutT, PAIR_ROWS>(p.g, p.dim_t, reinterpret_cast<const OutT*>(p.table), s_pts, s_off, s_flag, s_desc, G,
                     tokens ? tokens + row0 * p.g.E : nullptr, reinterpret_cast<OutT*>(p.pe_out) + row0 * p.g.E);
}

// ============================================================================================
// 5. element-wise discretisation
// ============================================================================================
__global__ void __launch_bounds__(kThreads) k_discretize(const float* __restrict__ c, size_t n, Discretize d, int has_range,
                                                         float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
    const int a = (int)(i % 3);
    float x = c[i];
    if (has_range) {
      out[i] = discretize1(d, x, a);
    } else {
      x = (d.div_mode == V3D_DIV_RECIPROCAL) ? __fmul_rn(x, d.rcp_voxel) : __fdiv_rn(x, d.voxel);
      out[i] = rintf(round_to_dtype(x, d.coord_dtype));
    }
  }
}

// ============================================================================================
// 6. self-test: div_exact(a, b) against div.rn.f32 for EVERY fp32 dividend a
// ============================================================================================
__global__ void __launch_bounds__(kThreads) k_selftest_div(float b, unsigned long long* mismatches) {
  const Divisor d = make_divisor(b);
  unsigned long long bad = 0;
  const unsigned long long total = 1ull << 32;
  for (unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * kThreads) {
    const float a = __uint_as_float((unsigned)i);
    const unsigned got = __float_as_uint(div_exact(a, d)), ref = __float_as_uint(__fdiv_rn(a, b));
    const bool both_nan = ((got & 0x7fffffffu) > 0x7f800000u) && ((ref & 0x7fffffffu) > 0x7f800000u);
    if (got != ref && !both_nan) ++bad;
  }
  for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}

// ============================================================================================
// host side
// ============================================================================================
static Clamp make_clamp(const float* lo_hi) {
  Clamp c;
  memset(&c, 0, sizeof(c));
  if (lo_hi) {
    c.on = 1;
    for (int a = 0; a < 3; ++a) {
      c.lo[a] = lo_hi[a];
      c.hi[a] = lo_hi[3 + a];
    }
  }
  return c;
}
static Discretize make_disc(int on, const float* lo, const float* hi, float voxel, int div_mode, int coord_dtype = V3D_F32) {
  Discretize d;
  memset(&d, 0, sizeof(d));
  d.on = on;
  for (int a = 0; a < 3; ++a) {
    d.lo[a] = lo ? lo[a] : 0.f;
    d.hi[a] = hi ? hi[a] : 0.f;
  }
  d.voxel = voxel;
  d.rcp_voxel = 1.0f / voxel;
  d.div_mode = div_mode;
  d.coord_dtype = coord_dtype;
  return d;
}

template <typename F>
static int dispatch_out(int dtype, F&& f) {
  switch (dtype) {
    case V3D_BF16: return f(__nv_bfloat16());
    case V3D_F16: return f(__half());
    case V3D_F32: return f(float());
    default: set_error("unknown dtype %d", dtype); return V3D_E_ARG;
  }
}

static int grid_for(size_t work_items, int per_block, int max_waves = 8) {
  size_t blocks = (work_items + per_block - 1) / per_block;
  size_t cap = (size_t)num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace v3d

using namespace v3d;

extern "C" {

int v3d_version(void) { return V3D_VERSION; }

int v3d_selftest_division(float divisor, uint64_t* mismatches, v3d_stream_t stream) {
  V3D_REQUIRE(mismatches, V3D_E_ARG, "v3d_selftest_division: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  V3D_CUDA(cudaMemsetAsync(mismatches, 0, sizeof(uint64_t), s));
  k_selftest_div<<<num_sms() * 16, kThreads, 0, s>>>(divisor, reinterpret_cast<unsigned long long*>(mismatches));
  return check_launch("k_selftest_div");
}
const char* v3d_last_error(void) { return g_err; }

int v3d_pool_points(int pool_mode) {
  switch (pool_mode) {
    case V3D_POOL_AVG: return 1;
    case V3D_POOL_MINMAX: return 2;
    case V3D_POOL_SAMPLE9: return 9;
    case V3D_POOL_SAMPLE5: return 5;
    case V3D_POOL_SAMPLE1: return 1;
    default: return -1;
  }
}

int v3d_unproject(const void* depth, int depth_dtype, const float* poses, const float* intr, int V, int H, int W,
                  float depth_scale, const int32_t* row_map, const int32_t* col_map, int out_h, int out_w,
                  const float* clamp_lo_hi_host, float* world_out, v3d_stream_t stream) {
  if (V == 0) return V3D_OK;  // empty batch: nothing to read or write (pointers may be null)
  V3D_REQUIRE(depth && poses && intr && world_out, V3D_E_ARG, "v3d_unproject: null pointer");
  V3D_REQUIRE(V >= 0 && H > 0 && W > 0, V3D_E_ARG, "v3d_unproject: bad shape V=%d H=%d W=%d", V, H, W);
  V3D_REQUIRE((row_map == nullptr) == (col_map == nullptr), V3D_E_ARG, "v3d_unproject: give both maps or none");
  V3D_REQUIRE((reinterpret_cast<uintptr_t>(world_out) & 15) == 0, V3D_E_ALIGN, "v3d_unproject: output not 16-byte aligned");
  const int OH = row_map ? out_h : H, OW = col_map ? out_w : W;
  V3D_REQUIRE(OH > 0 && OW > 0, V3D_E_ARG, "v3d_unproject: bad output size %dx%d", OH, OW);
  if (V == 0) return V3D_OK;
  const Clamp cl = make_clamp(clamp_lo_hi_host);
  const size_t groups = ((size_t)V * OH * OW + 3) / 4;
  const int grid = grid_for(groups, kThreads, 16);
  cudaStream_t s = (cudaStream_t)stream;
  const long row_blocks = (long)V * ((OH + 7) / 8);
  if ((OW & 3) == 0 && (OW >> 2) >= 2 && (OW >> 2) <= 8191 && row_blocks <= 0x7fffffffL) {
    // rows of whole 4-pixel groups: one CTA per (frame, 8 rows); idx / gpr by multiplication (idx < 2^16)
    const int rpb = 8;
    const uint32_t gpr = (uint32_t)(OW >> 2);
    const uint32_t magic = (uint32_t)(0xffffffffu / gpr + 1u);
    const long blocks = row_blocks;
    switch (depth_dtype) {
      case V3D_DEPTH_U16:
        k_unproject_rows<V3D_DEPTH_U16><<<(unsigned)blocks, kThreads, 0, s>>>(depth, poses, intr, V, H, W, depth_scale, row_map, col_map, OH, OW, rpb, magic, cl, world_out);
        break;
      case V3D_DEPTH_I32:
        k_unproject_rows<V3D_DEPTH_I32><<<(unsigned)blocks, kThreads, 0, s>>>(depth, poses, intr, V, H, W, depth_scale, row_map, col_map, OH, OW, rpb, magic, cl, world_out);
        break;
      case V3D_DEPTH_F32:
        k_unproject_rows<V3D_DEPTH_F32><<<(unsigned)blocks, kThreads, 0, s>>>(depth, poses, intr, V, H, W, depth_scale, row_map, col_map, OH, OW, rpb, magic, cl, world_out);
        break;
      default: set_error("v3d_unproject: unknown depth dtype %d", depth_dtype); return V3D_E_ARG;
    }
    return check_launch("k_unproject_rows");
  }
  switch (depth_dtype) {
    case V3D_DEPTH_U16:
      k_unproject<V3D_DEPTH_U16><<<grid, kThreads, 0, s>>>(depth, poses, intr, V, H, W, depth_scale, row_map, col_map, OH, OW, cl, world_out);
      break;
    case V3D_DEPTH_I32:
      k_unproject<V3D_DEPTH_I32><<<grid, kThreads, 0, s>>>(depth, poses, intr, V, H, W, depth_scale, row_map, col_map, OH, OW, cl, world_out);
      break;
    case V3D_DEPTH_F32:
      k_unproject<V3D_DEPTH_F32><<<grid, kThreads, 0, s>>>(depth, poses, intr, V, H, W, depth_scale, row_map, col_map, OH, OW, cl, world_out);
      break;
    default: set_error("v3d_unproject: unknown depth dtype %d", depth_dtype); return V3D_E_ARG;
  }
  return check_launch("k_unproject");
}

int v3d_upload_depth_window(const void* depth_host, void* depth_dev, size_t pitch_bytes, size_t n_rows_total,
                            size_t x_off_bytes, size_t width_bytes, size_t row_period, size_t rows_keep,
                            v3d_stream_t stream) {
  V3D_REQUIRE(depth_host && depth_dev, V3D_E_ARG, "v3d_upload_depth_window: null pointer");
  V3D_REQUIRE(pitch_bytes > 0 && width_bytes > 0 && x_off_bytes + width_bytes <= pitch_bytes, V3D_E_ARG,
              "v3d_upload_depth_window: window [%zu,+%zu) outside a row of %zu bytes", x_off_bytes, width_bytes, pitch_bytes);
  V3D_REQUIRE(row_period > 0 && rows_keep > 0 && rows_keep <= row_period && n_rows_total % row_period == 0, V3D_E_ARG,
              "v3d_upload_depth_window: %zu rows are not a whole number of groups of %zu", n_rows_total, row_period);
  if (n_rows_total == 0) return V3D_OK;
  cudaMemcpy3DParms prm;
  memset(&prm, 0, sizeof(prm));
  prm.srcPtr = make_cudaPitchedPtr(const_cast<void*>(depth_host), pitch_bytes, pitch_bytes, row_period);
  prm.dstPtr = make_cudaPitchedPtr(depth_dev, pitch_bytes, pitch_bytes, row_period);
  prm.srcPos = make_cudaPos(x_off_bytes, 0, 0);
  prm.dstPos = make_cudaPos(x_off_bytes, 0, 0);
  prm.extent = make_cudaExtent(width_bytes, rows_keep, n_rows_total / row_period);
  prm.kind = cudaMemcpyHostToDevice;
  V3D_CUDA(cudaMemcpy3DAsync(&prm, (cudaStream_t)stream));
  return V3D_OK;
}

int v3d_expand_depth_window(const void* packed_dev, void* depth_dev, size_t pitch_bytes, size_t n_rows_total, size_t x_off_bytes,
                            size_t width_bytes, size_t row_period, size_t rows_keep, v3d_stream_t stream) {
  V3D_REQUIRE(packed_dev && depth_dev, V3D_E_ARG, "v3d_expand_depth_window: null pointer");
  V3D_REQUIRE(pitch_bytes > 0 && width_bytes > 0 && x_off_bytes + width_bytes <= pitch_bytes, V3D_E_ARG,
              "v3d_expand_depth_window: window [%zu,+%zu) outside a row of %zu bytes", x_off_bytes, width_bytes, pitch_bytes);
  V3D_REQUIRE(row_period > 0 && rows_keep > 0 && rows_keep <= row_period && n_rows_total % row_period == 0, V3D_E_ARG,
              "v3d_expand_depth_window: %zu rows are not a whole number of groups of %zu", n_rows_total, row_period);
  if (n_rows_total == 0) return V3D_OK;
  cudaMemcpy3DParms prm;
  memset(&prm, 0, sizeof(prm));
  prm.srcPtr = make_cudaPitchedPtr(const_cast<void*>(packed_dev), width_bytes, width_bytes, rows_keep);
  prm.dstPtr = make_cudaPitchedPtr(depth_dev, pitch_bytes, pitch_bytes, row_period);
  prm.srcPos = make_cudaPos(0, 0, 0);
  prm.dstPos = make_cudaPos(x_off_bytes, 0, 0);
  prm.extent = make_cudaExtent(width_bytes, rows_keep, n_rows_total / row_period);
  prm.kind = cudaMemcpyDeviceToDevice;
  V3D_CUDA(cudaMemcpy3DAsync(&prm, (cudaStream_t)stream));
  return V3D_OK;
}

int v3d_scatter_resampled_depth(const void* packed_dev, void* depth_dev, int depth_dtype, int V, int H, int W, const int32_t* row_map,
                                const int32_t* col_map, int out_h, int out_w, v3d_stream_t stream) {
  if (V == 0) return V3D_OK;
  V3D_REQUIRE(packed_dev && depth_dev && row_map && col_map, V3D_E_ARG, "v3d_scatter_resampled_depth: null pointer");
  V3D_REQUIRE(V > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0, V3D_E_ARG, "v3d_scatter_resampled_depth: bad sizes");
  long blocks = (long)V * out_h;
  const long cap = (long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (depth_dtype == V3D_DEPTH_U16)
    k_scatter_resampled<uint16_t><<<(unsigned)blocks, kThreads, 0, s>>>(reinterpret_cast<const uint16_t*>(packed_dev), reinterpret_cast<uint16_t*>(depth_dev), V, H, W, row_map, col_map, out_h, out_w);
  else if (depth_dtype == V3D_DEPTH_I32 || depth_dtype == V3D_DEPTH_F32)
    k_scatter_resampled<uint32_t><<<(unsigned)blocks, kThreads, 0, s>>>(reinterpret_cast<const uint32_t*>(packed_dev), reinterpret_cast<uint32_t*>(depth_dev), V, H, W, row_map, col_map, out_h, out_w);
  else {
    set_error("v3d_scatter_resampled_depth: unknown depth dtype %d", depth_dtype);
    return V3D_E_ARG;
  }
  return check_launch("k_scatter_resampled");
}

int v3d_boundary(const void* depth, int depth_dtype, const float* poses, const float* intr, int V, int H, int W,
                 float depth_scale, const float* clamp_lo_hi_host, float* out6, v3d_stream_t stream) {
  V3D_REQUIRE(depth && poses && intr && out6, V3D_E_ARG, "v3d_boundary: null pointer");
  V3D_REQUIRE(V > 0 && H > 0 && W > 0, V3D_E_ARG, "v3d_boundary: bad shape");
  const Clamp cl = make_clamp(clamp_lo_hi_host);
  cudaStream_t s = (cudaStream_t)stream;
  int* keys = reinterpret_cast<int*>(out6);
  k_boundary_init<<<1, 32, 0, s>>>(keys);
  const int grid = grid_for((size_t)V * ((H + 7) / 8), 1, 8);
  switch (depth_dtype) {
    case V3D_DEPTH_U16: k_boundary<V3D_DEPTH_U16><<<grid, kThreads, 0, s>>>(depth, poses, intr, V, H, W, depth_scale, cl, keys); break;
    case V3D_DEPTH_I32: k_boundary<V3D_DEPTH_I32><<<grid, kThreads, 0, s>>>(depth, poses, intr, V, H, W, depth_scale, cl, keys); break;
    case V3D_DEPTH_F32: k_boundary<V3D_DEPTH_F32><<<grid, kThreads, 0, s>>>(depth, poses, intr, V, H, W, depth_scale, cl, keys); break;
    default: set_error("v3d_boundary: unknown depth dtype %d", depth_dtype); return V3D_E_ARG;
  }
  k_boundary_final<<<1, 32, 0, s>>>(keys);
  return check_launch("k_boundary");
}

int v3d_discretize(const float* coords, size_t n_points, const float* min_xyz_host, const float* max_xyz_host,
                   int has_range, float voxel_size, int div_mode, int coord_dtype, float* out, v3d_stream_t stream) {
  if (n_points == 0) return V3D_OK;
  V3D_REQUIRE(coords && out, V3D_E_ARG, "v3d_discretize: null pointer");
  V3D_REQUIRE(!has_range || (min_xyz_host && max_xyz_host), V3D_E_ARG, "v3d_discretize: range requested but not given");
  V3D_REQUIRE(voxel_size > 0.f, V3D_E_ARG, "v3d_discretize: voxel_size must be positive");
  V3D_REQUIRE(coord_dtype == V3D_BF16 || coord_dtype == V3D_F16 || coord_dtype == V3D_F32, V3D_E_ARG, "v3d_discretize: bad coord dtype %d", coord_dtype);
  const Discretize d = make_disc(1, min_xyz_host, max_xyz_host, voxel_size, div_mode, coord_dtype);
  const size_t n = n_points * 3;
  k_discretize<<<grid_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(coords, n, d, has_range, out);
  return check_launch("k_discretize");
}

size_t v3d_pe_table_bytes(int E, int n_points, int rows, int dtype) {
  PeGeom g;
  if (make_pe_geom(E, n_points, rows, &g) != V3D_OK || rows <= 0) return 0;
  return pe_table_elems(g) * dtype_size(dtype);
}

int v3d_pe_table_build(const float* dim_t, int E, int n_points, int rows, int dtype, void* table, v3d_stream_t stream) {
  V3D_REQUIRE(dim_t && table, V3D_E_ARG, "v3d_pe_table_build: null pointer");
  V3D_REQUIRE(rows > 0, V3D_E_ARG, "v3d_pe_table_build: rows must be positive");
  V3D_REQUIRE((reinterpret_cast<uintptr_t>(table) & 15) == 0, V3D_E_ALIGN, "v3d_pe_table_build: table not 16-byte aligned");
  PeGeom g;
  int rc = make_pe_geom(E, n_points, rows, &g);
  if (rc != V3D_OK) return rc;
  const size_t total = pe_table_elems(g);
  cudaStream_t s = (cudaStream_t)stream;
  return dispatch_out(dtype, [&](auto tag) {
    using T = decltype(tag);
    k_pe_table<T><<<grid_for(total, kThreads), kThreads, 0, s>>>(dim_t, g, reinterpret_cast<T*>(table));
    return check_launch("k_pe_table");
  });
}

int v3d_pe_sine3d(const float* coords, size_t n_tokens, int n_points, int E, const float* dim_t, const void* table,
                  int rows, int dtype, const void* tokens_in, void* out, v3d_stream_t stream) {
  if (n_tokens == 0) return V3D_OK;
  V3D_REQUIRE(coords && dim_t && out, V3D_E_ARG, "v3d_pe_sine3d: null pointer");
  PeGeom g;
  int rc = make_pe_geom(E, n_points, table ? rows : 0, &g);
  if (rc != V3D_OK) return rc;
  V3D_REQUIRE(!table || rows > 0, V3D_E_ARG, "v3d_pe_sine3d: table given with rows=%d", rows);
  const bool vec = (E & 7) == 0;
  V3D_REQUIRE(!vec || ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(tokens_in) | reinterpret_cast<uintptr_t>(table)) & 15) == 0,
              V3D_E_ALIGN, "v3d_pe_sine3d: out/tokens/table must be 16-byte aligned");
  if (n_tokens == 0) return V3D_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec) return launch_pe_flat(coords, n_tokens, g, dim_t, table, dtype, tokens_in, out, s);
  const size_t smem = (size_t)g.cpt * sizeof(ChunkDesc) + (size_t)kPeTokPerCta * ((size_t)g.n_slots * 8 + 8);
  const int grid = grid_for((n_tokens + kPeTokPerCta - 1) / kPeTokPerCta, 1, 16);
  return dispatch_out(dtype, [&](auto tag) {
    using T = decltype(tag);
    if (smem > 48 * 1024) V3D_CUDA(cudaFuncSetAttribute(k_pe_sine3d<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_pe_sine3d<T><<<grid, kThreads, smem, s>>>(coords, n_tokens, g, dim_t, reinterpret_cast<const T*>(table),
                                                reinterpret_cast<const T*>(tokens_in), reinterpret_cast<T*>(out));
    return check_launch("k_pe_sine3d");
  });
}

}  // extern "C"

// ---- fused launcher ---------------------------------------------------------------------------
template <int DT, bool FROM_DEPTH, bool USE_TMA, typename OutT, bool PAIR_ROWS = false>
static int launch_fused(const FusedParams& p, cudaStream_t s) {
  const size_t smem = fused_smem_bytes(p, FROM_DEPTH && USE_TMA);
  V3D_REQUIRE(smem <= 227 * 1024, V3D_E_UNSUPPORTED, "fused kernel needs %zu bytes of shared memory", smem);
  auto kern = k_fused<DT, FROM_DEPTH, USE_TMA, OutT, PAIR_ROWS>;
  if (smem > 48 * 1024) V3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<p.n_frames * p.grid, kThreads, smem, s>>>(p);
  return check_launch("k_fused");
}

template <bool FROM_DEPTH, typename OutT>
static int launch_fused_dt(int depth_dtype, bool use_tma, const FusedParams& p, cudaStream_t s) {
  if constexpr (!FROM_DEPTH) {
    return launch_fused<V3D_DEPTH_F32, false, false, OutT>(p, s);
  } else {
    switch (depth_dtype) {
      case V3D_DEPTH_U16:
        // the main variant comes in two instantiations: the second streams pairs of token rows per warp when the slots do
        // not start on 16-byte boundaries (nf % 8 != 0, e.g. nf = 1194 for E = 3584)
        if (use_tma && p.pe_mode != V3D_PE_NONE && (p.g.nf & 7) != 0) return launch_fused<V3D_DEPTH_U16, true, true, OutT, true>(p, s);
        return use_tma ? launch_fused<V3D_DEPTH_U16, true, true, OutT>(p, s) : launch_fused<V3D_DEPTH_U16, true, false, OutT>(p, s);
      // int32 / float32 depth (the reference's in-memory dtypes) are API-compatibility inputs: they take the
      // plain-load variant, which keeps the number of kernel instantiations (and the build time) down
      case V3D_DEPTH_I32: return launch_fused<V3D_DEPTH_I32, true, false, OutT>(p, s);
      case V3D_DEPTH_F32: return launch_fused<V3D_DEPTH_F32, true, false, OutT>(p, s);
      default: set_error("unknown depth dtype %d", depth_dtype); return V3D_E_ARG;
    }
  }
}

static int check_pool(int pool_mode, int patch, int grid, int* n_points) {
  *n_points = v3d_pool_points(pool_mode);
  V3D_REQUIRE(*n_points > 0, V3D_E_ARG, "unknown pool mode %d", pool_mode);
  V3D_REQUIRE(patch > 0 && grid > 0, V3D_E_ARG, "bad patch=%d grid=%d", patch, grid);
  V3D_REQUIRE(patch * grid <= 2 * kThreads && patch * grid <= kPlane, V3D_E_UNSUPPORTED, "patch*grid=%d exceeds %d columns", patch * grid, 2 * kThreads);
  V3D_REQUIRE(3 * grid <= kThreads, V3D_E_UNSUPPORTED, "grid=%d too large", grid);
  if (pool_mode >= V3D_POOL_SAMPLE9)
    V3D_REQUIRE(patch == 27 && grid == 14, V3D_E_UNSUPPORTED, "sample pooling is defined for patch 27 / grid 14 only (llava_arch.py:245)");
  return V3D_OK;
}

extern "C" {

int v3d_pool_coords(const float* coords, int V, int S, int patch, int grid, int pool_mode, float* out, v3d_stream_t stream) {
  if (V == 0) return V3D_OK;
  V3D_REQUIRE(coords && out, V3D_E_ARG, "v3d_pool_coords: null pointer");
  V3D_REQUIRE(V >= 0 && S > 0, V3D_E_ARG, "v3d_pool_coords: bad shape");
  int n_points;
  int rc = check_pool(pool_mode, patch, grid, &n_points);
  if (rc != V3D_OK) return rc;
  V3D_REQUIRE(patch * grid <= S, V3D_E_ARG, "v3d_pool_coords: patch*grid=%d exceeds S=%d", patch * grid, S);
  if (V == 0) return V3D_OK;
  FusedParams p;
  memset(&p, 0, sizeof(p));
  p.coords_in = coords;
  p.n_frames = V;
  p.S = S;
  p.patch = patch;
  p.grid = grid;
  p.Sp = patch * grid;
  p.Sp_pad = (p.Sp + 2) & ~1;
  p.pool_mode = pool_mode;
  p.n_points = n_points;
  p.pe_mode = V3D_PE_NONE;
  p.pooled_out = out;
  p.coord_dtype = V3D_F32;  // the caller's tensor already holds values of its dtype and rounds the result itself
  p.disc.coord_dtype = V3D_F32;
  return launch_fused_dt<false, __nv_bfloat16>(0, false, p, (cudaStream_t)stream);
}

int v3d_frontend_pe(const v3d_frontend_desc* d, v3d_stream_t stream) {
  V3D_REQUIRE(d, V3D_E_ARG, "v3d_frontend_pe: null descriptor");
  if (d->n_frames == 0) return V3D_OK;
  V3D_REQUIRE(d->depth && d->poses && d->intr && d->row_map && d->col_map, V3D_E_ARG, "v3d_frontend_pe: null input pointer");
  V3D_REQUIRE(d->n_frames >= 0 && d->H > 0 && d->W > 0, V3D_E_ARG, "v3d_frontend_pe: bad shape");
  int n_points;
  int rc = check_pool(d->pool_mode, d->patch, d->grid, &n_points);
  if (rc != V3D_OK) return rc;
  V3D_REQUIRE(d->pe_mode == V3D_PE_NONE || d->pe_mode == V3D_PE_SINE3D, V3D_E_ARG, "unknown pe_mode %d", d->pe_mode);
  V3D_REQUIRE(d->pe_mode != V3D_PE_NONE || d->pooled_out || d->pe_coords_out, V3D_E_ARG, "v3d_frontend_pe: nothing to compute");
  FusedParams p;
  memset(&p, 0, sizeof(p));
  p.depth = d->depth;
  p.poses = d->poses;
  p.intr = d->intr;
  p.n_frames = d->n_frames;
  p.H = d->H;
  p.W = d->W;
  p.depth_scale = d->depth_scale;
  p.row_map = d->row_map;
  p.col_map = d->col_map;
  p.patch = d->patch;
  p.grid = d->grid;
  p.Sp = d->patch * d->grid;
  p.Sp_pad = (p.Sp + 2) & ~1;
  p.pool_mode = d->pool_mode;
  p.n_points = n_points;
  if (d->has_clamp) {
    float lh[6] = {d->clamp_lo[0], d->clamp_lo[1], d->clamp_lo[2], d->clamp_hi[0], d->clamp_hi[1], d->clamp_hi[2]};
    p.clamp = make_clamp(lh);
  }
  V3D_REQUIRE(d->coord_dtype == V3D_BF16 || d->coord_dtype == V3D_F16 || d->coord_dtype == V3D_F32, V3D_E_ARG,
              "v3d_frontend_pe: bad coord_dtype %d", d->coord_dtype);
  p.coord_dtype = d->coord_dtype;
  p.disc.coord_dtype = d->coord_dtype;
  p.resampled = d->depth_resampled ? 1 : 0;
  if (p.resampled)
    V3D_REQUIRE(d->H >= p.Sp && d->W >= p.Sp, V3D_E_ARG, "v3d_frontend_pe: resampled depth is %dx%d, the grid needs %dx%d samples", d->H, d->W,
                p.Sp, p.Sp);
  if (d->discrete) {
    V3D_REQUIRE(d->voxel_size > 0.f, V3D_E_ARG, "v3d_frontend_pe: voxel_size must be positive");
    p.disc = make_disc(1, d->min_xyz, d->max_xyz, d->voxel_size, d->div_mode, d->coord_dtype);
  }
  p.pe_mode = d->pe_mode;
  p.pooled_out = d->pooled_out;
  p.pe_coords_out = d->pe_coords_out;
  if (d->pe_mode == V3D_PE_SINE3D) {
    V3D_REQUIRE(d->pe_out && d->dim_t, V3D_E_ARG, "v3d_frontend_pe: pe_out / dim_t missing");
    rc = make_pe_geom(d->E, n_points, d->table ? d->table_rows : 0, &p.g);
    if (rc != V3D_OK) return rc;
    V3D_REQUIRE(!d->table || d->table_rows > 0, V3D_E_ARG, "v3d_frontend_pe: table given with rows=%d", d->table_rows);
    if (d->discrete && d->table) {
      for (int a = 0; a < 3; ++a) {
        const float top = rintf((d->max_xyz[a] - d->min_xyz[a]) / d->voxel_size);
        V3D_REQUIRE(top < (float)d->table_rows, V3D_E_ARG, "v3d_frontend_pe: table has %d rows, axis %d needs %d", d->table_rows, a, (int)top + 1);
      }
    }
    if ((d->E & 7) == 0)
      V3D_REQUIRE(((reinterpret_cast<uintptr_t>(d->pe_out) | reinterpret_cast<uintptr_t>(d->tokens_in) | reinterpret_cast<uintptr_t>(d->table)) & 15) == 0,
                  V3D_E_ALIGN, "v3d_frontend_pe: pe_out/tokens_in/table must be 16-byte aligned");
    p.dim_t = d->dim_t;
    p.table = d->table;
    p.tokens_in = d->tokens_in;
    p.pe_out = d->pe_out;
  }
  if (d->n_frames == 0) return V3D_OK;

  // TMA staging geometry: bytes [lo, hi) of a depth row that cover every column the maps use.
  const size_t esz = depth_size(d->depth_dtype);
  V3D_REQUIRE(p.resampled || (d->src_col_min >= 0 && d->src_col_max >= d->src_col_min && d->src_col_max < d->W), V3D_E_ARG,
              "v3d_frontend_pe: src_col_min/max [%d,%d] outside [0,%d)", d->src_col_min, d->src_col_max, d->W);
  const bool use_tma = ((size_t)d->W * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->depth) & 15) == 0;
  if (use_tma) {
    // resampled depth: the first Sp stored columns of every row, whatever the source columns were
    const int lo = p.resampled ? 0 : (int)((d->src_col_min * esz) & ~(size_t)15);
    const int hi = (int)((((p.resampled ? p.Sp : d->src_col_max + 1)) * esz + 15) & ~(size_t)15);
    p.col_lo_byte = lo;
    p.span = hi - lo;
  }
  cudaStream_t s = (cudaStream_t)stream;
  // the recipe configuration streams through k_fused_ring (v3d_fused_ring.cu); same results bit for bit
  if (!d->exact_pooling && fused_ring_supported(p, d->depth_dtype, use_tma, d->dtype)) return launch_fused_ring(p, d->dtype, s);
  return dispatch_out(d->pe_mode == V3D_PE_NONE ? V3D_BF16 : d->dtype, [&](auto tag) {
    using T = decltype(tag);
    return launch_fused_dt<true, T>(d->depth_dtype, use_tma, p, s);
  });
}

}  // extern "C"
