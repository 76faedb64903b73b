// v3d_pe.cuh -- sin-3D position-encoding building blocks shared by the front-end kernels (v3d_frontend.cu) and the
// projector-tail kernel (v3d_pool.cu): the lookup-table geometry, 8-channel chunks, packed adds.
#pragma once

#include <math_constants.h>

#include <type_traits>

#include "v3d_common.cuh"

namespace v3d {

// llava/model/position_encoding.py:27-42: feature j of a coordinate x
__device__ __forceinline__ float pe_value(float x, int j, const float* __restrict__ dim_t) {
  const float pos = __fdiv_rn(x, __ldg(dim_t + j));
  return (j & 1) ? cosf(pos) : sinf(pos);
}

// Table geometry shared by host and device: slot s covers output channels [s*nf, (s+1)*nf);
// its table rows start at channel lo_s = floor8(s*nf) and are Wt = ceil8(nf+7) wide.
__host__ __device__ __forceinline__ int pe_table_width(int nf) { return (nf + 7 + 7) & ~7; }

struct PeGeom {
  int E, nf, n_slots, cpt;  // cpt = chunks (of 8 channels) per token = ceil(E/8)
  int rows, Wt;             // table rows (without the NaN row), row width in elements
};

// per-chunk descriptor: which slots a chunk of 8 channels touches and where in their table rows.
// Each half is (slot << 16) | element offset inside the slot's row; slot == n_slots is the virtual
// "zero" slot whose row (the last row of the table) is all zeros, so that the streaming loop needs no
// branches: a chunk is always row(s0)[k0..] | row(s1)[k1..].
struct __align__(8) ChunkDesc {
  uint32_t a, b;
};

__device__ __forceinline__ ChunkDesc make_chunk_desc(int c8, const PeGeom& g) {
  ChunkDesc d;
  const int c = c8 * 8;
  const int s0 = c / g.nf, s1 = (c + 7) / g.nf;
  const uint32_t none = (uint32_t)g.n_slots << 16;
  d.a = (s0 < g.n_slots) ? (((uint32_t)s0 << 16) | (uint32_t)(c - ((s0 * g.nf) & ~7))) : none;
  d.b = (s1 != s0 && s1 < g.n_slots) ? (((uint32_t)s1 << 16) | (uint32_t)(c - ((s1 * g.nf) & ~7))) : none;
  return d;
}

__host__ __device__ __forceinline__ size_t pe_table_elems(const PeGeom& g) {
  return ((size_t)g.n_slots * (g.rows + 1) + 1) * g.Wt;  // + the zero row
}

// one chunk = 8 consecutive channels
template <typename OutT>
struct alignas(sizeof(OutT) * 8) Chunk {
  OutT v[8];
};

template <typename OutT>
__device__ __forceinline__ Chunk<OutT> chunk_zero() {
  Chunk<OutT> c;
#pragma unroll
  for (int e = 0; e < 8; ++e) c.v[e] = Elem<OutT>::from(0.0f);
  return c;
}
template <typename OutT>
__device__ __forceinline__ Chunk<OutT> chunk_ldg(const OutT* p) {
  Chunk<OutT> c;
  if constexpr (sizeof(OutT) == 2) {
    *reinterpret_cast<uint4*>(&c) = __ldg(reinterpret_cast<const uint4*>(p));
  } else {
    reinterpret_cast<uint4*>(&c)[0] = __ldg(reinterpret_cast<const uint4*>(p));
    reinterpret_cast<uint4*>(&c)[1] = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  }
  return c;
}
template <typename OutT>
__device__ __forceinline__ Chunk<OutT> chunk_ld_stream(const OutT* p) {
  Chunk<OutT> c;
  if constexpr (sizeof(OutT) == 2) {
    *reinterpret_cast<uint4*>(&c) = ld_stream_v4(p);
  } else {
    reinterpret_cast<uint4*>(&c)[0] = ld_stream_v4(p);
    reinterpret_cast<uint4*>(&c)[1] = ld_stream_v4(reinterpret_cast<const uint4*>(p) + 1);
  }
  return c;
}
template <typename OutT>
__device__ __forceinline__ void chunk_st_stream(OutT* p, const Chunk<OutT>& c) {
  if constexpr (sizeof(OutT) == 2) {
    st_stream_v4(p, *reinterpret_cast<const uint4*>(&c));
  } else {
    st_stream_v4(p, reinterpret_cast<const uint4*>(&c)[0]);
    st_stream_v4(reinterpret_cast<uint4*>(p) + 1, reinterpret_cast<const uint4*>(&c)[1]);
  }
}
template <typename OutT>
__device__ __forceinline__ void chunk_or(Chunk<OutT>& a, const Chunk<OutT>& b) {
  constexpr int n = sizeof(Chunk<OutT>) / 4;
#pragma unroll
  for (int i = 0; i < n; ++i) reinterpret_cast<uint32_t*>(&a)[i] |= reinterpret_cast<const uint32_t*>(&b)[i];
}

template <typename OutT>
__device__ __forceinline__ Chunk<OutT> chunk_add(const Chunk<OutT>& a, const Chunk<OutT>& b) {
  // a + b with ONE rounding to OutT per element.  For 16-bit types the packed add rounds the exact
  // sum once, which equals torch's "add in fp32, round to bf16/fp16" for every pair of inputs: the
  // fp32 sum of two 8/11-bit significands is exact unless their exponents differ by more than 16/13,
  // and then the smaller term is below a quarter ulp of the result either way.
  Chunk<OutT> c;
  if constexpr (std::is_same<OutT, __nv_bfloat16>::value) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<__nv_bfloat162*>(&c)[i] = __hadd2(reinterpret_cast<const __nv_bfloat162*>(&a)[i], reinterpret_cast<const __nv_bfloat162*>(&b)[i]);
  } else if constexpr (std::is_same<OutT, __half>::value) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<__half2*>(&c)[i] = __hadd2(reinterpret_cast<const __half2*>(&a)[i], reinterpret_cast<const __half2*>(&b)[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) c.v[i] = __fadd_rn(a.v[i], b.v[i]);
  }
  return c;
}

__device__ __forceinline__ uint2 lds_u2(const void* p) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(smem_u32(p)));
  return r;
}
__device__ __forceinline__ int lds_s32(const void* p) {
  int r;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(smem_u32(p)));
  return r;
}

// element offset of the table row serving coordinate x in slot s: integer-valued in [0,rows) -> that
// row, NaN -> the NaN row, anything else -> -1 (direct evaluation)
__device__ __forceinline__ int pe_row_offset(float x, int s, const PeGeom& g, bool have_table) {
  if (!have_table) return -1;
  int row;
  if (x != x) row = g.rows;
  else if (x >= 0.0f && x < (float)g.rows && x == rintf(x)) row = (int)x;
  else return -1;
  return (s * (g.rows + 1) + row) * g.Wt;
}
__device__ __forceinline__ int pe_zero_row_offset(const PeGeom& g) { return g.n_slots * (g.rows + 1) * g.Wt; }


// host: table geometry of an (E, n_points, rows) configuration
inline int make_pe_geom(int E, int n_points, int rows, PeGeom* g) {
  V3D_REQUIRE(E > 0 && n_points > 0, V3D_E_ARG, "bad E=%d / n_points=%d", E, n_points);
  const int nf = E / (3 * n_points);
  V3D_REQUIRE(nf >= 8, V3D_E_UNSUPPORTED, "embedding %d too small for %d points (nf=%d < 8)", E, n_points, nf);
  V3D_REQUIRE(3 * n_points <= 254, V3D_E_UNSUPPORTED, "too many points per token (%d)", n_points);
  g->E = E;
  g->nf = nf;
  g->n_slots = 3 * n_points;
  g->cpt = (E + 7) / 8;
  g->rows = rows;
  g->Wt = pe_table_width(nf);
  V3D_REQUIRE(nf + 14 < 65536 && E < 65536 * 8, V3D_E_UNSUPPORTED, "embedding too large");
  return V3D_OK;
}

// flat streaming PE (+ add) kernel of v3d_pool.cu, for rows that are whole 16-byte chunks
int launch_pe_flat(const float* coords, size_t n_tokens, const PeGeom& g, const float* dim_t, const void* table, int dtype,
                   const void* tokens, void* out, cudaStream_t s);

}  // namespace v3d
