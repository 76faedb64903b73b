// v3d_pool.cu -- the projector tail on sm_100a (C ABI: include/v3d.h).
//
// One kernel for what the reference does to the projected vision tokens of a video between mm_projector and the
// language model (llava/model/llava_arch.py):
//   get_2dPool          :191-210  (V,27*27,E) -> permute -> interpolate(bilinear) | avg_pool2d | max_pool2d -> (V,14*14,E)
//   PE add              :506-517  + world_position_embedding(coords)   (position_encoding.py:17-49)
//   add_token_per_grid  :307-328  image_newline appended to every grid row -> (V*14*15, E)
// The reference runs ~30 ATen launches that move ~25 MB per frame (two permute+contiguous copies of the 5 MB input,
// the PE tensor, the cat); here every input token row is read once and every output row written once.
//
// One CTA per (frame, output grid row): its two source rows are streamed with 128-bit loads, four taps are combined
// in fp32 and rounded to the token dtype (the tensor get_2dPool returns), the PE chunk comes from the L2-resident
// lookup table (v3d_pe.cuh) and is added with one more rounding, exactly the reference's two roundings.
#include "v3d_pe.cuh"

namespace v3d {

static constexpr int kPoolThreads = 256;
static constexpr int kPoolMaxGrid = 32;

struct PoolParams {
  const void* tokens;
  void* out;
  const float* coords;
  const float* dim_t;
  const void* table;
  const void* newline;
  PeGeom g;
  int n_frames, src_grid, out_grid, mode, has_pe, has_table;
  // taps of output index i (rows and columns share them: the grids are square)
  int i0[kPoolMaxGrid], i1[kPoolMaxGrid];
  float w0[kPoolMaxGrid], w1[kPoolMaxGrid];
};

template <typename OutT>
__device__ __forceinline__ void chunk_to_float(const Chunk<OutT>& c, float (&f)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = Elem<OutT>::to(c.v[e]);
}

// the four taps of one output element -> the value get_2dPool returns (rounded to OutT by the caller)
__device__ __forceinline__ float pool_taps(int mode, float a, float b, float c, float d, float wy0, float wy1, float wx0, float wx1) {
  if (mode == V3D_TOKEN_POOL_BILINEAR) {
    // columns first, each step mul + fma (upsample_bilinear2d accumulates in fp32 for 16-bit inputs)
    const float h0 = __fmaf_rn(wy1, c, __fmul_rn(wy0, a));
    const float h1 = __fmaf_rn(wy1, d, __fmul_rn(wy0, b));
    return __fmaf_rn(wx1, h1, __fmul_rn(wx0, h0));
  }
  if (mode == V3D_TOKEN_POOL_NONE) return a;
  if (mode == V3D_TOKEN_POOL_AVERAGE) {
    // avg_pool2d: row-major fp32 sum of the window, divided by its size
    return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d), 0.25f);
  }
  // max_pool2d: a later element replaces the maximum if it is larger or NaN
  float m = a;
  m = (b > m || b != b) ? b : m;
  m = (c > m || c != c) ? c : m;
  m = (d > m || d != d) ? d : m;
  return m;
}

// one 8-channel chunk of one output token: taps -> pooled value (rounded to OutT) -> + PE -> store
template <typename OutT>
__device__ __forceinline__ void pool_chunk(const PoolParams& p, const PeGeom& g, const OutT* __restrict__ table, const int* off,
                                           const ChunkDesc* s_desc, const float* pts, int c8, const Chunk<OutT>& ca, const Chunk<OutT>& cb,
                                           const Chunk<OutT>& cc, const Chunk<OutT>& cd, float wy0, float wy1, float wx0, float wx1,
                                           OutT* po) {
  const int ch = c8 * 8;
  const int mode = p.mode;
  Chunk<OutT> pe;
  bool direct = false;
  if (p.has_pe) {
    const ChunkDesc d = s_desc[c8];
    const int oa = off[d.a >> 16], ob = off[d.b >> 16];
    direct = (oa < 0) || (ob < 0);
    if (!direct) {
      pe = chunk_ldg(table + (uint32_t)oa + (d.a & 0xffffu));
      if ((g.nf & 7) != 0) chunk_or(pe, chunk_ldg(table + (uint32_t)ob + (d.b & 0xffffu)));
    }
  }
  float fa[8], fb[8], fc[8], fd[8];
  chunk_to_float(ca, fa);
  chunk_to_float(cb, fb);
  chunk_to_float(cc, fc);
  chunk_to_float(cd, fd);
  Chunk<OutT> pooled;
#pragma unroll
  for (int e = 0; e < 8; ++e) pooled.v[e] = Elem<OutT>::from(pool_taps(mode, fa[e], fb[e], fc[e], fd[e], wy0, wy1, wx0, wx1));
  if (p.has_pe) {
    if (direct) {  // non-integer / out-of-range coordinates: evaluate the slot(s) of this chunk directly
#pragma unroll 1
      for (int e = 0; e < 8; ++e) {
        const int s = (ch + e) / g.nf;
        float val = 0.0f;
        if (s < g.n_slots) {
          const int j = ch + e - s * g.nf;
          const int o = off[s];
          val = (o >= 0) ? Elem<OutT>::to(table[o + (s * g.nf - ((s * g.nf) & ~7)) + j]) : pe_value(pts[s], j, p.dim_t);
        }
        pe.v[e] = Elem<OutT>::from(val);
      }
    }
    pooled = chunk_add(pooled, pe);
  }
  chunk_st_stream(po + ch, pooled);
}

template <typename OutT, bool ONE_TAP>
__global__ void __launch_bounds__(kPoolThreads) k_pool_pe_grid(const PoolParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PeGeom g = p.g;
  const int Go = p.out_grid, Gs = p.src_grid, E = g.E, cpt = g.cpt;
  const int n_off = g.n_slots + 1;
  ChunkDesc* s_desc = reinterpret_cast<ChunkDesc*>(smem_raw);
  float* s_pts = reinterpret_cast<float*>(s_desc + cpt);
  int* s_off = reinterpret_cast<int*>(s_pts + Go * g.n_slots);
  const int v = blockIdx.x / Go, oy = blockIdx.x - v * Go;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kPoolThreads / 32;
  const OutT* table = reinterpret_cast<const OutT*>(p.table);

  if (p.has_pe) {
    for (int i = threadIdx.x; i < cpt; i += kPoolThreads) s_desc[i] = make_chunk_desc(i, g);
    for (int i = threadIdx.x; i < Go * g.n_slots; i += kPoolThreads) {
      const int ox = i / g.n_slots, sl = i - ox * g.n_slots;
      const float x = __ldg(p.coords + ((size_t)(v * Go + oy) * Go + ox) * g.n_slots + sl);
      s_pts[i] = x;
      s_off[ox * n_off + sl] = pe_row_offset(x, sl, g, p.has_table != 0);
    }
    for (int ox = threadIdx.x; ox < Go; ox += kPoolThreads) s_off[ox * n_off + g.n_slots] = pe_zero_row_offset(g);
    __syncthreads();
  }

  const int nl = p.newline != nullptr;
  const OutT* src = reinterpret_cast<const OutT*>(p.tokens) + (size_t)v * Gs * Gs * E;
  OutT* orow = reinterpret_cast<OutT*>(p.out) + ((size_t)(v * Go + oy) * (Go + nl)) * E;
  const int y0 = p.i0[oy], y1 = p.i1[oy];
  const float wy0 = p.w0[oy], wy1 = p.w1[oy];
  const OutT* r0 = src + (size_t)y0 * Gs * E;
  const OutT* r1 = src + (size_t)y1 * Gs * E;
  for (int ox = warp; ox < Go; ox += kWarps) {
    const int x0 = p.i0[ox], x1 = p.i1[ox];
    const float wx0 = p.w0[ox], wx1 = p.w1[ox];
    const OutT *pa = r0 + (size_t)x0 * E, *pb = r0 + (size_t)x1 * E, *pc = r1 + (size_t)x0 * E, *pd = r1 + (size_t)x1 * E;
    OutT* po = orow + (size_t)ox * E;
    const int* off = s_off + ox * n_off;
    // two chunks per lane and iteration: eight 128-bit token loads in flight before the first use
    for (int c8a = lane; c8a < cpt; c8a += 64) {
      const int c8b = c8a + 32;
      const bool has_b = c8b < cpt;
      const int cha = c8a * 8, chb = (has_b ? c8b : c8a) * 8;
      const Chunk<OutT> a0 = chunk_ld_stream(pa + cha), a1 = chunk_ld_stream(pa + chb);
      Chunk<OutT> b0 = a0, c0 = a0, d0 = a0, b1 = a1, c1 = a1, d1 = a1;
      if constexpr (!ONE_TAP) {  // one tap when the grid is kept (V3D_TOKEN_POOL_NONE)
        b0 = chunk_ld_stream(pb + cha);
        c0 = chunk_ld_stream(pc + cha);
        d0 = chunk_ld_stream(pd + cha);
        b1 = chunk_ld_stream(pb + chb);
        c1 = chunk_ld_stream(pc + chb);
        d1 = chunk_ld_stream(pd + chb);
      }
      pool_chunk<OutT>(p, g, table, off, s_desc, s_pts + ox * g.n_slots, c8a, a0, b0, c0, d0, wy0, wy1, wx0, wx1, po);
      if (has_b) pool_chunk<OutT>(p, g, table, off, s_desc, s_pts + ox * g.n_slots, c8b, a1, b1, c1, d1, wy0, wy1, wx0, wx1, po);
    }
  }
  if (nl) {  // add_token_per_grid: image_newline closes the grid row
    const OutT* nlp = reinterpret_cast<const OutT*>(p.newline);
    OutT* po = orow + (size_t)Go * E;
    for (int c8 = threadIdx.x; c8 < cpt; c8 += kPoolThreads) chunk_st_stream(po + c8 * 8, chunk_ldg(nlp + c8 * 8));
  }
}

// ---- stand-alone PE (+ add) over a flat list of tokens -----------------------------------------------------
// position_encoding.py:17-49 (+ the add of llava_arch.py:515) with the loop structure of the kernel above: 16 tokens
// per CTA, one warp per token, two 8-channel chunks per lane and iteration (the token loads are issued before the
// table loads and the adds).  PE only: 2 bytes written per element; with the add 2 read + 2 written.
static constexpr int kFlatTok = 16;

template <typename OutT>
__global__ void __launch_bounds__(kPoolThreads) k_pe_flat(const float* __restrict__ coords, size_t n_tokens, const PeGeom g,
                                                          const float* __restrict__ dim_t, const OutT* __restrict__ table,
                                                          const OutT* tokens, OutT* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int cpt = g.cpt, E = g.E, n_off = g.n_slots + 1;
  ChunkDesc* s_desc = reinterpret_cast<ChunkDesc*>(smem_raw);
  float* s_pts = reinterpret_cast<float*>(s_desc + cpt);
  int* s_off = reinterpret_cast<int*>(s_pts + kFlatTok * g.n_slots);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kPoolThreads / 32;
  for (int i = threadIdx.x; i < cpt; i += kPoolThreads) s_desc[i] = make_chunk_desc(i, g);
  for (size_t t0 = (size_t)blockIdx.x * kFlatTok; t0 < n_tokens; t0 += (size_t)gridDim.x * kFlatTok) {
    const int n_tok = (int)min((size_t)kFlatTok, n_tokens - t0);
    __syncthreads();  // the previous block of tokens is done with s_off / s_pts
    for (int i = threadIdx.x; i < n_tok * g.n_slots; i += kPoolThreads) {
      const int tk = i / g.n_slots, sl = i - tk * g.n_slots;
      const float x = __ldg(coords + (t0 + tk) * g.n_slots + sl);
      s_pts[i] = x;
      s_off[tk * n_off + sl] = pe_row_offset(x, sl, g, table != nullptr);
    }
    for (int tk = threadIdx.x; tk < n_tok; tk += kPoolThreads) s_off[tk * n_off + g.n_slots] = pe_zero_row_offset(g);
    __syncthreads();
    for (int tk = warp; tk < n_tok; tk += kWarps) {
      const int* off = s_off + tk * n_off;
      const float* pts = s_pts + tk * g.n_slots;
      const OutT* tp = tokens ? tokens + (t0 + tk) * E : nullptr;
      OutT* op = out + (t0 + tk) * E;
      for (int c8a = lane; c8a < cpt; c8a += 64) {
        const int c8b = c8a + 32;
        const bool has_b = c8b < cpt;
        Chunk<OutT> ta, tb;
        if (tp) {
          ta = chunk_ld_stream(tp + c8a * 8);
          if (has_b) tb = chunk_ld_stream(tp + c8b * 8);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !has_b) break;
          const int c8 = h ? c8b : c8a, ch = c8 * 8;
          const ChunkDesc d = s_desc[c8];
          const int oa = off[d.a >> 16], ob = off[d.b >> 16];
          Chunk<OutT> pe;
          if (oa >= 0 && ob >= 0) {
            pe = chunk_ldg(table + (uint32_t)oa + (d.a & 0xffffu));
            if ((g.nf & 7) != 0) chunk_or(pe, chunk_ldg(table + (uint32_t)ob + (d.b & 0xffffu)));
          } else {  // non-integer / out-of-range coordinates: evaluate the slot(s) of this chunk directly
#pragma unroll 1
            for (int e = 0; e < 8; ++e) {
              const int s = (ch + e) / g.nf;
              float val = 0.0f;
              if (s < g.n_slots) {
                const int j = ch + e - s * g.nf;
                const int o = off[s];
                val = (o >= 0) ? Elem<OutT>::to(table[o + (s * g.nf - ((s * g.nf) & ~7)) + j]) : pe_value(pts[s], j, dim_t);
              }
              pe.v[e] = Elem<OutT>::from(val);
            }
          }
          chunk_st_stream(op + ch, tp ? chunk_add(h ? tb : ta, pe) : pe);
        }
      }
    }
  }
}

template <typename OutT>
static int launch_pe_flat_t(const float* coords, size_t n_tokens, const PeGeom& g, const float* dim_t, const void* table,
                            const void* tokens, void* out, cudaStream_t s) {
  const size_t smem = sizeof(ChunkDesc) * g.cpt + (size_t)kFlatTok * (g.n_slots * sizeof(float) + (g.n_slots + 1) * sizeof(int)) + 16;
  V3D_REQUIRE(smem <= 200 * 1024, V3D_E_UNSUPPORTED, "v3d_pe_sine3d: %zu bytes of shared memory", smem);
  if (smem > 48 * 1024) V3D_CUDA(cudaFuncSetAttribute(k_pe_flat<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  size_t blocks = (n_tokens + kFlatTok - 1) / kFlatTok;
  const size_t cap = (size_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  k_pe_flat<OutT><<<(unsigned)blocks, kPoolThreads, smem, s>>>(coords, n_tokens, g, dim_t, reinterpret_cast<const OutT*>(table),
                                                                reinterpret_cast<const OutT*>(tokens), reinterpret_cast<OutT*>(out));
  return check_launch("k_pe_flat");
}

// E % 8 == 0 path of v3d_pe_sine3d (v3d_frontend.cu)
int launch_pe_flat(const float* coords, size_t n_tokens, const PeGeom& g, const float* dim_t, const void* table, int dtype,
                   const void* tokens, void* out, cudaStream_t s) {
  switch (dtype) {
    case V3D_BF16: return launch_pe_flat_t<__nv_bfloat16>(coords, n_tokens, g, dim_t, table, tokens, out, s);
    case V3D_F16: return launch_pe_flat_t<__half>(coords, n_tokens, g, dim_t, table, tokens, out, s);
    case V3D_F32: return launch_pe_flat_t<float>(coords, n_tokens, g, dim_t, table, tokens, out, s);
    default: set_error("v3d_pe_sine3d: unknown dtype %d", dtype); return V3D_E_ARG;
  }
}

// taps of F.interpolate(mode="bilinear", align_corners=False) from `in` to `out` samples: ATen's
// area_pixel_compute_source_index in fp32 -- scale = float(in)/out, src = fma(scale, dst + 0.5, -0.5) clamped at 0.
static void bilinear_taps(int in, int out, int* i0, int* i1, float* w0, float* w1) {
  const float scale = (float)in / (float)out;
  for (int o = 0; o < out; ++o) {
    float s = (float)((double)scale * ((double)o + 0.5) - 0.5);  // exact in double, rounded once: the fused multiply-add
    if (s < 0.0f) s = 0.0f;
    int a = (int)s;
    if (a > in - 1) a = in - 1;
    const float l1 = s - (float)a;
    i0[o] = a;
    i1[o] = a + (a < in - 1 ? 1 : 0);
    w0[o] = 1.0f - l1;
    w1[o] = l1;
  }
}

}  // namespace v3d

using namespace v3d;

extern "C" {

int v3d_token_pool_out_grid(int src_grid, int stride, int pool_mode) {
  if (src_grid <= 0 || stride <= 0) return -1;
  if (pool_mode == V3D_TOKEN_POOL_BILINEAR) return (src_grid + stride - 1) / stride;  // math.ceil(height / stride), llava_arch.py:203
  if (pool_mode == V3D_TOKEN_POOL_AVERAGE || pool_mode == V3D_TOKEN_POOL_MAX) return src_grid / stride;
  if (pool_mode == V3D_TOKEN_POOL_NONE) return src_grid;
  return -1;
}

int v3d_pool_pe_grid(const v3d_pool_desc* d, v3d_stream_t stream) {
  V3D_REQUIRE(d != nullptr, V3D_E_ARG, "v3d_pool_pe_grid: null descriptor");
  V3D_REQUIRE(d->n_frames >= 0 && d->src_grid > 0 && d->E > 0, V3D_E_ARG, "v3d_pool_pe_grid: bad sizes");
  if (d->n_frames == 0) return V3D_OK;
  V3D_REQUIRE(d->tokens && d->out, V3D_E_ARG, "v3d_pool_pe_grid: null tokens / out");
  V3D_REQUIRE((d->E & 7) == 0, V3D_E_UNSUPPORTED, "v3d_pool_pe_grid: E=%d must be a multiple of 8", d->E);
  const int Go = v3d_token_pool_out_grid(d->src_grid, d->stride, d->pool_mode);
  V3D_REQUIRE(Go > 0, V3D_E_ARG, "v3d_pool_pe_grid: bad pool mode %d / stride %d", d->pool_mode, d->stride);
  V3D_REQUIRE(Go <= kPoolMaxGrid, V3D_E_UNSUPPORTED, "v3d_pool_pe_grid: output grid %d > %d", Go, kPoolMaxGrid);
  V3D_REQUIRE(d->pool_mode == V3D_TOKEN_POOL_BILINEAR || d->pool_mode == V3D_TOKEN_POOL_NONE || d->stride == 2, V3D_E_UNSUPPORTED,
              "v3d_pool_pe_grid: average/max pooling is implemented for stride 2 (the reference's setting)");
  V3D_REQUIRE(d->dtype == V3D_BF16 || d->dtype == V3D_F16 || d->dtype == V3D_F32, V3D_E_ARG, "v3d_pool_pe_grid: bad dtype");
  V3D_REQUIRE(((reinterpret_cast<uintptr_t>(d->tokens) | reinterpret_cast<uintptr_t>(d->out) | reinterpret_cast<uintptr_t>(d->newline)) & 15) == 0,
              V3D_E_ALIGN, "v3d_pool_pe_grid: tokens / out / newline must be 16-byte aligned");
  PoolParams p = {};
  p.tokens = d->tokens;
  p.out = d->out;
  p.newline = d->newline;
  p.n_frames = d->n_frames;
  p.src_grid = d->src_grid;
  p.out_grid = Go;
  p.mode = d->pool_mode;
  p.has_pe = d->coords != nullptr;
  if (p.has_pe) {
    V3D_REQUIRE(d->dim_t != nullptr, V3D_E_ARG, "v3d_pool_pe_grid: dim_t is required with coords");
    int rc = make_pe_geom(d->E, d->n_points, d->table ? d->table_rows : 1, &p.g);
    if (rc != V3D_OK) return rc;
    V3D_REQUIRE(!d->table || (reinterpret_cast<uintptr_t>(d->table) & 15) == 0, V3D_E_ALIGN, "v3d_pool_pe_grid: table not 16-byte aligned");
    p.coords = d->coords;
    p.dim_t = d->dim_t;
    p.table = d->table;
    p.has_table = d->table != nullptr;
  } else {
    p.g = PeGeom{d->E, d->E, 0, d->E / 8, 1, 8};
  }
  if (d->pool_mode == V3D_TOKEN_POOL_BILINEAR) {
    bilinear_taps(d->src_grid, Go, p.i0, p.i1, p.w0, p.w1);
  } else {
    const int step = d->pool_mode == V3D_TOKEN_POOL_NONE ? 1 : 2;
    for (int o = 0; o < Go; ++o) {
      p.i0[o] = step * o;
      p.i1[o] = step * o + step - 1;
      p.w0[o] = p.w1[o] = 0.5f;
    }
  }
  const size_t smem = p.has_pe ? sizeof(ChunkDesc) * p.g.cpt + sizeof(float) * Go * p.g.n_slots + sizeof(int) * Go * (p.g.n_slots + 1) + 16 : 16;
  V3D_REQUIRE(smem <= 48 * 1024, V3D_E_UNSUPPORTED, "v3d_pool_pe_grid: %zu bytes of shared memory", smem);
  const dim3 grid((unsigned)d->n_frames * Go);
  cudaStream_t s = (cudaStream_t)stream;
  const bool one = d->pool_mode == V3D_TOKEN_POOL_NONE;
  switch (d->dtype) {
    case V3D_BF16:
      if (one) k_pool_pe_grid<__nv_bfloat16, true><<<grid, kPoolThreads, smem, s>>>(p);
      else k_pool_pe_grid<__nv_bfloat16, false><<<grid, kPoolThreads, smem, s>>>(p);
      break;
    case V3D_F16:
      if (one) k_pool_pe_grid<__half, true><<<grid, kPoolThreads, smem, s>>>(p);
      else k_pool_pe_grid<__half, false><<<grid, kPoolThreads, smem, s>>>(p);
      break;
    default:
      if (one) k_pool_pe_grid<float, true><<<grid, kPoolThreads, smem, s>>>(p);
      else k_pool_pe_grid<float, false><<<grid, kPoolThreads, smem, s>>>(p);
      break;
  }
  return check_launch("k_pool_pe_grid");
}

}  // extern "C"
