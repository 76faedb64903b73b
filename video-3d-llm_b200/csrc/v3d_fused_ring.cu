// v3d_fused_ring.cu -- the fused front end as a streaming kernel: algebraic patch means (approximate, then verify) and a
// TMA ring for the token rows.  Same contract as k_fused (v3d_frontend.cu) for the configuration the recipes run:
// uint16 depth, 'avg' pooling, discretised coordinates, table-backed sin-3D PE added to 16-bit tokens
// (video_utils.py:230,296-308 -> llava_arch.py:213-223,259-272 -> position_encoding.py:17-49 -> llava_arch.py:515).
//
// Why a second kernel.  k_fused evaluates the reference's operator sequence per pixel (23 fp32 operations plus ATen's
// strictly sequential window sum), which made the 27-pixel geometry issue-bound, and it reads its token rows with plain
// 128-bit loads, whose latency its streaming loop cannot cover on its own (ncu: long-scoreboard stalls, DRAM at 60-64 %).
// Here
//   * the patch mean comes from three exact integer sums per patch (two integer operations per pixel) and is accepted
//     only when a proven error band around it discretises to one integer -- then that integer IS the reference's result;
//     a coordinate inside the band (a few per cent of the patch rows hold one) is recomputed on the spot with the
//     reference's operator sequence (exact_coord), and so is every coordinate of a frame whose camera or resampling maps
//     do not qualify for the algebra;
//   * persistent CTAs pull the token rows through a ring of shared-memory stages with cp.async.bulk (TMA) running
//     S-1 stages ahead -- across patch rows -- so that ~100 KB per SM are in flight without a single scoreboard stall;
//     the add reads the stage with 128-bit shared loads and streams the result out with 128-bit stores;
//   * the depth rows of the patch rows ahead are staged (TMA) by the pool warps while the current ones stream.
// Measurements, the ablation and what was tried and rejected: DESIGN.md section 4.1.
#include <math_constants.h>
#include <stdlib.h>

#include "v3d_fused.cuh"

namespace v3d {

static constexpr int kStreamWarps = 16;  // the add: stage -> registers -> global
static constexpr int kPoolWarps = 8;     // patch means of the patch rows ahead
static constexpr int kStreamThreads = kStreamWarps * 32, kPoolThreads = kPoolWarps * 32;
static constexpr int kRingThreads = kStreamThreads + kPoolThreads + 32;  // + one producer warp (TMA issue)
static constexpr int kMaxRingSlots = 12;
static constexpr int kItemSlots = 8;                // patch rows the pool warps may run ahead of the stream warps
static constexpr int kChunksPerStageTarget = 1024;  // 16 KB stages: two 16-byte chunks per stream thread
static constexpr int kMaxPoolGroups = 4;

struct RingGeom {
  int n_items;       // n_frames * grid patch rows
  int item_chunks;   // 16-byte chunks of one patch row of tokens (grid * cpt)
  int stage_rows;    // whole token rows per ring stage
  int stage_chunks;  // = stage_rows * cpt
  int n_stage;       // stages per item
  int slots;         // ring slots
  int span_limit;    // largest (u_last - u0) / (v_last - v0) for which the 32-bit window sums are exact
  double inv_n;      // 1 / pixels per patch
  int pool_groups;   // independent groups of pool warps
  int depth_bufs;    // depth buffers per group (1: the next patch row's depth rows are requested when the current one is done)
};

struct RingSmem {
  size_t bars, fail, off, cs, du, dv, mono, coef, scratch, depth, depth_stride, ring, total;
};
__host__ __device__ inline RingSmem ring_smem(const FusedParams& p, const RingGeom& r) {
  RingSmem m;
  size_t b = 0;
  m.bars = b;
  b += 8 * (2 * kMaxRingSlots + 2 * kItemSlots + 2 * kMaxPoolGroups);
  m.fail = b;
  b += (size_t)r.pool_groups * (1 + 3 * p.grid) * sizeof(int);  // per group: count + (axis, patch) of the coordinates inside the band
  b = (b + 15) & ~(size_t)15;
  m.off = b;
  b += (size_t)kItemSlots * p.grid * 4 * sizeof(int);  // G tokens x (3 axes + zero row) byte offsets into the table, per item slot
  m.cs = b;
  b += (size_t)r.pool_groups * 3 * p.Sp * sizeof(unsigned);
  m.du = b;
  b += (size_t)p.Sp * sizeof(int);
  m.dv = b;
  b += (size_t)p.Sp * sizeof(int);
  m.mono = b;
  b += (size_t)2 * p.grid * sizeof(int);  // windows of patch column j / patch row gi qualify for the integer sums
  b = (b + 15) & ~(size_t)15;
  m.coef = b;
  b += (size_t)kMaxPoolGroups * 3 * 40;  // AxisCoef
  b = (b + 15) & ~(size_t)15;
  m.scratch = b;
  b += (size_t)kPoolWarps * p.patch * p.patch * sizeof(float);  // one window per pool warp (exact_coord)
  b = (b + 15) & ~(size_t)15;
  m.depth = b;
  m.depth_stride = (size_t)p.patch * p.span;
  b += (size_t)r.depth_bufs * r.pool_groups * m.depth_stride;
  b = (b + 127) & ~(size_t)127;
  m.ring = b;
  b += (size_t)r.slots * r.stage_chunks * 16;
  m.total = b;
  return m;
}

__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ int lds_i32(uint32_t addr) {
  int r;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint4 ldg_v4(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <typename OutT>
__device__ __forceinline__ uint4 add_packed(const uint4& a, const uint4& b) {
  Chunk<OutT> x, y;
  *reinterpret_cast<uint4*>(&x) = a;
  *reinterpret_cast<uint4*>(&y) = b;
  const Chunk<OutT> z = chunk_add(x, y);
  return *reinterpret_cast<const uint4*>(&z);
}

// ---- algebraic patch means (approximate, then verify) ----------------------------------------------------------------
// For a rigid pose the world coordinate of a pixel is, in real arithmetic, linear in (d, u*d, v*d):
//     w_i = a_i*um*d + b_i*vm*d + c_i*d + P_i3,   a_i = P_i0/(fx*scale), b_i = P_i1/(fy*scale), c_i = P_i2/scale,
// so the mean over a patch needs only the integer sums S0 = sum d, S1 = sum (u-u0)*d, S2 = sum (v-v0)*d (exact in 32
// bits: checked against RingGeom::span_limit).  The reference's fp32 result deviates from that real mean by at most
//     eps_i = u*1.01*(((n+1)/2 + 12)*Smax_i + ((n+1)/2 + 6)*|P_i3|) + 2^-130,   u = 2^-24, n = pixels per patch,
//     Smax_i = dmax*(|a_i|*um_max + |b_i|*vm_max + |c_i|)   (largest |w_i - P_i3| in the patch):
// per pixel the operator sequence (video_utils.py:57-65: three divisions, mul, fma, fma, add, and the rounding of u - cx)
// is off by <= u*(10*Smax + 3|P_i3|); ATen's sequential fp32 sum of n values of magnitude <= V = Smax + |P_i3|
// contributes <= u*V*n(n+1)/2, i.e. u*V*(n+1)/2 in the mean; the division by n u*V.  The camera test below keeps every
// intermediate inside the normal fp32 range (no overflow; underflow is covered by the absolute term).  Discretisation
// (llava_arch.py:264-270) is monotone, so when it maps mean - eps and mean + eps (rounded outwards to fp32) to the same
// integer, that integer is the reference's result.
__device__ __forceinline__ bool in_range(float v, float lo, float hi) { return fabsf(v) >= lo && fabsf(v) <= hi; }  // false for NaN

__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar_addr),
      "r"(parity)
      : "memory");
}

struct __align__(8) AxisCoef {  // per (patch row of this CTA, axis): the linear form of the world coordinate, in fp64
  double aa, bb, cc, p3;        // aa = NaN: the camera does not qualify, every coordinate of the frame goes to the work list
  float cx, cy;
};

// the camera of frame f, axis a
__device__ __noinline__ void make_axis_coef(const float* __restrict__ intr, const float* __restrict__ poses, int f, int a, float depth_scale,
                                            AxisCoef* out) {
  const float* K = intr + 16 * (size_t)f;
  const float* Pm = poses + 16 * (size_t)f;
  const float fx = __ldg(K + 0), fy = __ldg(K + 5), cx = __ldg(K + 2), cy = __ldg(K + 6);
  const float p0 = __ldg(Pm + 4 * a), p1 = __ldg(Pm + 4 * a + 1), p2 = __ldg(Pm + 4 * a + 2), p3 = __ldg(Pm + 4 * a + 3);
  // the operator sequence stays inside the normal fp32 range (|w| < 2^100) and the homogeneous coordinate is exactly 1
  bool cam_ok = __ldg(Pm + 12) == 0.0f && __ldg(Pm + 13) == 0.0f && __ldg(Pm + 14) == 0.0f && __ldg(Pm + 15) == 1.0f;
  cam_ok = cam_ok && in_range(fx, 0x1p-16f, 0x1p24f) && in_range(fy, 0x1p-16f, 0x1p24f) && depth_scale >= 0x1p-16f && depth_scale <= 0x1p24f;
  cam_ok = cam_ok && fabsf(cx) <= 0x1p20f && fabsf(cy) <= 0x1p20f;
#pragma unroll
  for (int i = 0; i < 12; ++i) cam_ok = cam_ok && fabsf(__ldg(Pm + i)) <= 0x1p30f;
  const double sc_ = (double)depth_scale;
  AxisCoef c;
  c.aa = cam_ok ? (double)p0 / ((double)fx * sc_) : (double)CUDART_NAN_F;
  c.bb = (double)p1 / ((double)fy * sc_);
  c.cc = (double)p2 / sc_;
  c.p3 = (double)p3;
  c.cx = cx;
  c.cy = cy;
  *out = c;
}

// One coordinate (axis a of patch j) of a patch row: window sums from the column sums, the verified mean, the table row.
// Out of line: the fp64 work that only 3*G threads run once per patch row must not set the register budget of the
// streaming loop.  inv_n = 1 / pixels per patch (the product differs from the quotient by 2^-53 relative: covered by
// the 1 % slack of eps).
__device__ __noinline__ bool finalize_patch(int a, int j, int item, int gi, int G, int P, int Sp, const unsigned* s_cs, const int* s_du,
                                            const int* s_dv, const int32_t* __restrict__ col_map, const int32_t* __restrict__ row_map,
                                            const AxisCoef* coef, double inv_n, const Discretize disc, const PeGeom g, const int* s_mono, int* s_off_item,
                                            float* pe_coords_out, int* s_fail) {
  const int c_lo = j * P;
  unsigned S0 = 0u, S1 = 0u, S2 = 0u, dmax = 0u;
#pragma unroll 2
  for (int dx = 0; dx < P; ++dx) {
    const unsigned v0 = s_cs[c_lo + dx];
    S0 += v0;
    S1 += v0 * (unsigned)s_du[c_lo + dx];
    S2 += s_cs[Sp + c_lo + dx];
    dmax = max(dmax, s_cs[2 * Sp + c_lo + dx]);
  }
  const int u0 = __ldg(col_map + c_lo), u_span = s_du[c_lo + P - 1];
  const int v0 = __ldg(row_map + gi * P), v_span = s_dv[gi * P + P - 1];
  const AxisCoef c = *coef;
  const double n_px = (double)(P * P);
  const double um0 = (double)u0 - (double)c.cx, vm0 = (double)v0 - (double)c.cy;
  const double s0 = (double)S0;
  const double mean = (c.aa * ((double)S1 + um0 * s0) + c.bb * ((double)S2 + vm0 * s0) + c.cc * s0) * inv_n + c.p3;
  const double um_max = fmax(fabs(um0), fabs(um0 + (double)u_span)), vm_max = fmax(fabs(vm0), fabs(vm0 + (double)v_span));
  const double smax = (double)dmax * (fabs(c.aa) * um_max + fabs(c.bb) * vm_max + fabs(c.cc));
  const double half_n = 0.5 * (n_px + 1.0);
  const double eps = (1.01 / 16777216.0) * ((half_n + 12.0) * smax + (half_n + 6.0) * fabs(c.p3)) + 0x1p-130;
  const float lo = __double2float_rd(mean - eps), hi = __double2float_ru(mean + eps);
  const float pv = discretize1(disc, lo, a), pv_hi = discretize1(disc, hi, a);
  const bool mono = s_mono[j] != 0 && s_mono[G + gi] != 0;  // the windows qualify for the 32-bit sums (checked once per CTA)
  const int o = pe_row_offset(pv, a, g, true);
  const bool ok = mono && pv == pv_hi && o >= 0;  // NaN (camera not ok, non-finite pose) fails the comparison
  if (ok) {
    V3D_CHECK(j >= 0 && j < G && a >= 0 && a < 3);
    s_off_item[j * 4 + a] = o * 2;  // byte offset into the table
    if (pe_coords_out) pe_coords_out[((size_t)item * G + j) * 3 + a] = pv;
  } else {
    s_fail[1 + atomicAdd(s_fail, 1)] = (a << 16) | j;  // exact_coord takes over
  }
  return ok;
}

// One coordinate by the reference's operator sequence (video_utils.py:57-65 per pixel, ATen's sequential window sum,
// llava_arch.py:221-223, 264-270), called by a whole warp: the lanes unproject the patch's pixels, lane 0 adds them in
// avg_pool2d's order.  Same device functions and same order as k_fused, so the result is the same bit for bit.
__device__ __noinline__ void exact_coord(int a, int j, int item, int f, int gi, int G, int P, const unsigned char* s_depth_item, int span,
                                         int col_lo_byte, int resampled, const int32_t* __restrict__ col_map, const int32_t* __restrict__ row_map,
                                         const float* __restrict__ intr, const float* __restrict__ poses, float depth_scale, const Discretize disc,
                                         const PeGeom g, float* scratch, int* s_off_item, float* pe_coords_out) {
  const int lane = threadIdx.x & 31;
  const Camera cam = load_camera(poses, intr, f);
  const Divisor ds = make_divisor(depth_scale);
  const int n = P * P;
  for (int px = lane; px < n; px += 32) {
    const int dy = px / P, dx = px - dy * P;
    const int c = j * P + dx;
    const int sc = __ldg(col_map + c), sr = __ldg(row_map + gi * P + dy);
    const int stored = resampled ? c : sc;
    const float d = (float)*reinterpret_cast<const uint16_t*>(s_depth_item + (size_t)dy * span + (stored * 2 - col_lo_byte));
    float w[3];
    unproject_px(cam, d, __fsub_rn((float)sc, cam.cx), __fsub_rn((float)sr, cam.cy), ds, w[0], w[1], w[2]);
    V3D_CHECK(stored * 2 - col_lo_byte >= 0 && stored * 2 - col_lo_byte + 2 <= span && dy < P);
    scratch[px] = (a == 0) ? w[0] : (a == 1 ? w[1] : w[2]);
  }
  __syncwarp();
  if (lane == 0) {
    float acc = 0.0f;
    for (int i = 0; i < n; ++i) acc = __fadd_rn(acc, scratch[i]);
    const float pv = discretize1(disc, __fdiv_rn(acc, (float)n), a);
    s_off_item[j * 4 + a] = pe_row_offset(pv, a, g, true) * 2;  // discretised (or NaN): always a table row
    if (pe_coords_out) pe_coords_out[((size_t)item * G + j) * 3 + a] = pv;
  }
  __syncwarp();
}

// Warp-specialised, one CTA per SM:
//   * 8 pool warps in 4 independent groups form the patch means of the patch rows up to kItemSlots ahead (depth rows
//     staged by TMA, one buffer per group: a group requests its next patch row's depth when it has published the current
//     one), recompute the coordinates the algebra cannot settle with the operator sequence, and publish the table-row
//     offsets of every token (`iready`);
//   * the producer warp's lane 0 keeps the token ring full, looking across patch rows (`full` / `empty`);
//   * 16 stream warps add: table rows of the NEXT stage are requested before the current stage is waited for, a warp
//     never waits for another warp -- it reads its chunks of a stage, arrives on `empty`, adds and stores.
template <typename OutT, bool STRADDLE>
__global__ void __launch_bounds__(kRingThreads, 1) k_fused_ring(const FusedParams p, const RingGeom r) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const RingSmem m = ring_smem(p, r);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int G = p.grid, P = p.patch;
  constexpr int n_off = 4;  // 3 axes + the zero row

  const uint32_t full0 = smem_u32(smem_raw + m.bars);                 // [kMaxRingSlots]
  const uint32_t empty0 = full0 + 8u * kMaxRingSlots;                 // [kMaxRingSlots]
  const uint32_t iready0 = empty0 + 8u * kMaxRingSlots;               // [kItemSlots]
  const uint32_t ifree0 = iready0 + 8u * kItemSlots;                  // [kItemSlots]
  const uint32_t dbar0 = ifree0 + 8u * kItemSlots;                    // [2 * pool groups]
  int* s_fail = reinterpret_cast<int*>(smem_raw + m.fail);
  int* s_off = reinterpret_cast<int*>(smem_raw + m.off);
  unsigned* s_cs = reinterpret_cast<unsigned*>(smem_raw + m.cs);
  int* s_du = reinterpret_cast<int*>(smem_raw + m.du);
  int* s_dv = reinterpret_cast<int*>(smem_raw + m.dv);
  AxisCoef* s_coef = reinterpret_cast<AxisCoef*>(smem_raw + m.coef);
  unsigned char* s_depth = smem_raw + m.depth;
  const uint32_t ring_base = smem_u32(smem_raw + m.ring);
  const uint32_t stage_bytes = (uint32_t)r.stage_chunks * 16u;

  const int nk = (r.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items of this CTA (>= 1)
  const size_t item_bytes = (size_t)r.item_chunks * 16;

  // ---- per-CTA constants
  if (t == 0) {
    for (int i = 0; i < r.slots; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + m.bars) + i, 1);
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + m.bars) + kMaxRingSlots + i, kStreamWarps);
    }
    for (int i = 0; i < kItemSlots; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + m.bars) + 2 * kMaxRingSlots + i, 1);
      mbar_init(reinterpret_cast<uint64_t*>(smem_raw + m.bars) + 2 * kMaxRingSlots + kItemSlots + i, kStreamWarps);
    }
    for (int i = 0; i < 2 * kMaxPoolGroups; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem_raw + m.bars) + 2 * kMaxRingSlots + 2 * kItemSlots + i, 1);
    mbar_fence_init();
  }
  for (int c = t; c < p.Sp; c += kRingThreads) {
    s_du[c] = __ldg(p.col_map + c) - __ldg(p.col_map + (c / P) * P);
    s_dv[c] = __ldg(p.row_map + c) - __ldg(p.row_map + (c / P) * P);
  }
  for (int i = t; i < kItemSlots * G; i += kRingThreads) s_off[i * n_off + 3] = pe_zero_row_offset(p.g) * 2;
  // source columns / rows must be non-decreasing inside a window (nearest resize + crop maps are) and close enough for
  // the 32-bit sums
  int* s_mono = reinterpret_cast<int*>(smem_raw + m.mono);
  for (int i = t; i < 2 * G; i += kRingThreads) {
    const int32_t* map = (i < G) ? p.col_map : p.row_map;
    const int w = (i < G) ? i : i - G;
    const int first = __ldg(map + w * P);
    bool ok = first >= 0 && first < 65536 && __ldg(map + w * P + P - 1) - first <= r.span_limit;
    for (int d = 1; d < P; ++d) ok = ok && __ldg(map + w * P + d) >= __ldg(map + w * P + d - 1);
    s_mono[i] = ok ? 1 : 0;
  }
  __syncthreads();

  // ================================= producer warp =================================
  if (warp == kStreamWarps + kPoolWarps) {
    if (lane != 0) return;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.tokens_in) + (size_t)blockIdx.x * item_bytes;
    const size_t src_step = (size_t)gridDim.x * item_bytes;
    // The depth rows of the first patch rows are on the critical path of the whole CTA (nothing streams before the pool
    // groups have published a row): let them reach DRAM before the ring's 100+ KB of token loads queue up in front
    mbar_wait_addr(dbar0, 0u);
    uint32_t slot = 0, use_parity = 0;  // parity of the `empty` phase that frees the slot for its next use (first lap: none)
    bool first_lap = true;
    for (int ord = 0; ord < nk; ++ord, src += src_step) {
      int c0 = 0;
      for (int j = 0; j < r.n_stage; ++j, c0 += r.stage_chunks) {
        if (!first_lap) mbar_wait_addr(empty0 + 8u * slot, use_parity);
        const uint32_t bytes = (uint32_t)min(r.stage_chunks, r.item_chunks - c0) * 16u;
        V3D_CHECK(slot < (uint32_t)r.slots && bytes <= stage_bytes && bytes > 0u);
        V3D_CHECK((size_t)(src - reinterpret_cast<const unsigned char*>(p.tokens_in)) + (size_t)c0 * 16 + bytes <= (size_t)r.n_items * item_bytes);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8u * slot), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring_base + slot * stage_bytes),
                     "l"(src + (size_t)c0 * 16), "r"(bytes), "r"(full0 + 8u * slot)
                     : "memory");
        if (++slot == (uint32_t)r.slots) {
          slot = 0;
          if (first_lap) first_lap = false;
          else use_parity ^= 1u;
        }
      }
    }
    return;
  }

  // ================================= pool warps =================================
  // r.pool_groups independent groups; group g forms the patch means of this CTA's patch rows g, g + groups, ...: the work
  // of one patch row is a chain of dependent steps (stage, column sums, barrier, window sums and fp64 check, barrier),
  // so throughput comes from having several rows in flight, not from more threads per row.
  if (warp >= kStreamWarps) {
    const int NG = r.pool_groups, GT = kPoolThreads / NG;
    const int tp = t - kStreamThreads;
    const int grp = tp / GT, tg = tp - grp * GT, gwarp = tg >> 5;
    const uint32_t bar_id = 1u + (uint32_t)grp;
    auto bar_group = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"((uint32_t)GT) : "memory"); };
    auto bar_group_and = [&](bool pred) -> bool {
      uint32_t res;
      asm volatile("{\n.reg .pred p, q;\nsetp.ne.u32 p, %1, 0;\nbar.red.and.pred q, %2, %3, p;\nselp.u32 %0, 1, 0, q;\n}\n"
                   : "=r"(res)
                   : "r"((uint32_t)pred), "r"(bar_id), "r"((uint32_t)GT)
                   : "memory");
      return res != 0;
    };
    unsigned* cs = s_cs + (size_t)grp * 3 * p.Sp;
    int* fail = s_fail + (size_t)grp * (1 + 3 * G);
    AxisCoef* coef = s_coef + 3 * grp;
    float* scratch = reinterpret_cast<float*>(smem_raw + m.scratch) + (size_t)(tp >> 5) * P * P;
    if (tg == 0) fail[0] = 0;  // visible to the group after its first barrier
    auto issue_depth = [&](int n) {  // the group's first warp: stage the P depth rows of the group's n-th item
      const int item = (int)blockIdx.x + (grp + n * NG) * (int)gridDim.x;
      const int f = item / G, gi = item - f * G;
      const uint32_t b = (uint32_t)(grp * r.depth_bufs + (r.depth_bufs == 2 ? (n & 1) : 0));
      const uint32_t bar = dbar0 + 8u * b;
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(P * p.span)) : "memory");
      __syncwarp();
      unsigned char* dst = s_depth + (size_t)b * m.depth_stride;
      for (int dy = lane; dy < P; dy += 32) {
        const int sr = p.resampled ? gi * P + dy : __ldg(p.row_map + gi * P + dy);
        const unsigned char* g = reinterpret_cast<const unsigned char*>(p.depth) + ((size_t)f * p.H + sr) * p.W * 2 + p.col_lo_byte;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + (size_t)dy * p.span)),
                     "l"(g), "r"((uint32_t)p.span), "r"(bar)
                     : "memory");
      }
    };
    if (gwarp == 0) {
      if (grp < nk) issue_depth(0);
      if (r.depth_bufs == 2 && grp + NG < nk) issue_depth(1);
    }
    int n = 0;
    for (int ord = grp; ord < nk; ord += NG, ++n) {
      const int item = (int)blockIdx.x + ord * (int)gridDim.x;
      const int f = item / G, gi = item - f * G;
      const uint32_t is = (uint32_t)ord % kItemSlots;
      const uint32_t b = (uint32_t)(grp * r.depth_bufs + (r.depth_bufs == 2 ? (n & 1) : 0));
      // camera coefficients of the frame: the last three threads of the group, while the others wait for the depth rows
      if (tg >= GT - 3) make_axis_coef(p.intr, p.poses, f, tg - (GT - 3), p.depth_scale, coef + (tg - (GT - 3)));
      if (ord >= kItemSlots) mbar_wait_addr(ifree0 + 8u * is, (((uint32_t)ord / kItemSlots) - 1u) & 1u);
      mbar_wait_addr(dbar0 + 8u * b, (r.depth_bufs == 2 ? ((uint32_t)n >> 1) : (uint32_t)n) & 1u);
      // column sums over the P staged rows: S0 = sum d, S2 = sum (v - v0)*d, and the largest depth; two columns at a time
      {
        const uint32_t dbase = smem_u32(s_depth) + (uint32_t)(b * m.depth_stride) - (uint32_t)p.col_lo_byte;
        const uint32_t dv0 = smem_u32(s_dv + gi * P);
        for (int c0 = tg; c0 < p.Sp; c0 += 2 * GT) {
          const int c1 = c0 + GT;
          const bool has1 = c1 < p.Sp;
          uint32_t a0 = dbase + 2u * (uint32_t)(p.resampled ? c0 : __ldg(p.col_map + c0));
          uint32_t a1 = has1 ? dbase + 2u * (uint32_t)(p.resampled ? c1 : __ldg(p.col_map + c1)) : a0;  // no second column: re-read the first
          unsigned s00 = 0u, s20 = 0u, mx0 = 0u, s01 = 0u, s21 = 0u, mx1 = 0u;
          uint32_t dvp = dv0;
#pragma unroll 4
          for (int dy = 0; dy < P; ++dy, dvp += 4u, a0 += (uint32_t)p.span, a1 += (uint32_t)p.span) {
            V3D_CHECK(a0 >= smem_u32(s_depth) + (uint32_t)(b * m.depth_stride) && a0 + 2u <= smem_u32(s_depth) + (uint32_t)((b + 1) * m.depth_stride));
            V3D_CHECK(a1 >= smem_u32(s_depth) + (uint32_t)(b * m.depth_stride) && a1 + 2u <= smem_u32(s_depth) + (uint32_t)((b + 1) * m.depth_stride));
            V3D_CHECK(dvp + 4u <= smem_u32(s_dv + p.Sp));
            const unsigned dv = (unsigned)lds_i32(dvp);
            const unsigned d0 = lds_u16(a0), d1 = lds_u16(a1);
            s00 += d0;
            s20 += d0 * dv;
            mx0 = max(mx0, d0);
            s01 += d1;
            s21 += d1 * dv;
            mx1 = max(mx1, d1);
          }
          V3D_CHECK(c0 < p.Sp && grp < r.pool_groups);
          cs[c0] = s00;
          cs[p.Sp + c0] = s20;
          cs[2 * p.Sp + c0] = mx0;
          if (has1) {
            cs[c1] = s01;
            cs[p.Sp + c1] = s21;
            cs[2 * p.Sp + c1] = mx1;
          }
        }
      }
      bar_group();
      bool ok = true;
      for (int q = tg; q < 3 * G; q += GT) {
        const int a = q / G;
        ok = finalize_patch(a, q - a * G, item, gi, G, P, p.Sp, cs, s_du, s_dv, p.col_map, p.row_map, coef + a, r.inv_n, p.disc, p.g, s_mono,
                            s_off + (size_t)is * G * n_off, p.pe_coords_out, fail) && ok;
      }
      if (!bar_group_and(ok)) {
        // coordinates inside the band (or a camera / map the algebra does not cover): the operator sequence, one
        // coordinate per warp at a time
        const int n_fail = *reinterpret_cast<volatile int*>(fail);
        for (int i = gwarp; i < n_fail; i += GT / 32) {
          const int e = reinterpret_cast<volatile int*>(fail)[1 + i];
          exact_coord(e >> 16, e & 0xffff, item, f, gi, G, P, s_depth + (size_t)b * m.depth_stride, p.span, p.col_lo_byte, p.resampled, p.col_map,
                      p.row_map, p.intr, p.poses, p.depth_scale, p.disc, p.g, scratch, s_off + (size_t)is * G * n_off, p.pe_coords_out);
        }
        bar_group();
        if (tg == 0) fail[0] = 0;  // next written by the finalize threads of the group's next item, after its first barrier
      }
      // s_off of this slot is complete; this depth buffer, the coefficients and the column sums are free again
      if (tg == 0) mbar_arrive_addr(iready0 + 8u * is);  // release: visible to the waiters
      if (gwarp == 0 && ord + r.depth_bufs * NG < nk) issue_depth(n + r.depth_bufs);
    }
    return;
  }

  // ================================= stream warps =================================
  unsigned char* out_ptr = reinterpret_cast<unsigned char*>(p.pe_out) + (size_t)blockIdx.x * item_bytes + (size_t)t * 16;
  const size_t out_step = (size_t)gridDim.x * item_bytes;
  const unsigned char* table_bytes = reinterpret_cast<const unsigned char*>(p.table);

  // this thread's chunks of a stage (whole token rows per stage, so the chunk's place in its row never changes)
  uint32_t oa[2], ka[2], ob[2], kb[2];
  int tl[2];
  bool valid[2], two[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int ci = t + u * kStreamThreads;
    valid[u] = ci < r.stage_chunks;
    tl[u] = valid[u] ? ci / p.g.cpt : 0;
    const int c8 = valid[u] ? ci - tl[u] * p.g.cpt : 0;
    const ChunkDesc d = make_chunk_desc(c8, p.g);
    oa[u] = (uint32_t)(tl[u] * n_off + (int)(d.a >> 16)) * 4u;
    ka[u] = (d.a & 0xffffu) * 2u;
    ob[u] = (uint32_t)(tl[u] * n_off + (int)(d.b >> 16)) * 4u;
    kb[u] = (d.b & 0xffffu) * 2u;
    two[u] = STRADDLE && (d.b >> 16) != 3u;  // only the chunks at slot boundaries touch a second slot
  }
  const uint32_t ring_t = ring_base + (uint32_t)t * 16u;
  const uint32_t off0 = smem_u32(s_off);
  const uint32_t off_stage_step = (uint32_t)(r.stage_rows * n_off) * 4u;

  auto load_pe = [&](uint32_t off_stage, int tok0, uint4 (&pe)[2], bool (&act)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      act[u] = valid[u] && tok0 + tl[u] < G;
      if (act[u]) {
        V3D_CHECK(off_stage + oa[u] + 4u <= smem_u32(s_off) + (uint32_t)(kItemSlots * G * n_off) * 4u);
        V3D_CHECK((size_t)(uint32_t)lds_i32(off_stage + oa[u]) + ka[u] + 16 <= pe_table_elems(p.g) * 2);
        pe[u] = ldg_v4(table_bytes + ((uint32_t)lds_i32(off_stage + oa[u]) + ka[u]));
        if (STRADDLE) {
          if (two[u]) {
            V3D_CHECK((size_t)(uint32_t)lds_i32(off_stage + ob[u]) + kb[u] + 16 <= pe_table_elems(p.g) * 2);
            const uint4 q = ldg_v4(table_bytes + ((uint32_t)lds_i32(off_stage + ob[u]) + kb[u]));
            pe[u].x |= q.x;
            pe[u].y |= q.y;
            pe[u].z |= q.z;
            pe[u].w |= q.w;
          }
        }
      }
    }
  };

  uint32_t slot = 0, parity = 0;
  for (int ord = 0; ord < nk; ++ord, out_ptr += out_step) {
    const uint32_t is = (uint32_t)ord % kItemSlots;
    mbar_wait_addr(iready0 + 8u * is, ((uint32_t)ord / kItemSlots) & 1u);
    {
      uint32_t off_stage = off0 + is * (uint32_t)(G * n_off) * 4u;
      unsigned char* out_stage = out_ptr;
      uint4 pe[2], pn[2], tv[2];
      bool act[2], actn[2] = {false, false};
      load_pe(off_stage, 0, pe, act);
      int tok0 = 0;
      for (int j = 0; j < r.n_stage; ++j) {
        // table rows of the next stage first: they do not depend on the arrival of any stage
        if (j + 1 < r.n_stage) load_pe(off_stage + off_stage_step, tok0 + r.stage_rows, pn, actn);
        mbar_wait_addr(full0 + 8u * slot, parity);
        const uint32_t src = ring_t + slot * stage_bytes;
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (act[u]) {
            V3D_CHECK(slot < (uint32_t)r.slots && src + (uint32_t)(u * kStreamThreads * 16) + 16u <= ring_base + (uint32_t)r.slots * stage_bytes);
            V3D_CHECK((size_t)(out_stage - reinterpret_cast<unsigned char*>(p.pe_out)) + (size_t)(u * kStreamThreads * 16) + 16 <=
                      (size_t)r.n_items * item_bytes);
            tv[u] = lds_v4(src + (uint32_t)(u * kStreamThreads * 16));
          }
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(empty0 + 8u * slot);  // this warp has read its chunks of the slot
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (act[u]) st_stream_v4(out_stage + (size_t)(u * kStreamThreads * 16), add_packed<OutT>(tv[u], pe[u]));
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          pe[u] = pn[u];
          act[u] = actn[u];
        }
        if (++slot == (uint32_t)r.slots) {
          slot = 0;
          parity ^= 1u;
        }
        off_stage += off_stage_step;
        tok0 += r.stage_rows;
        out_stage += stage_bytes;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_addr(ifree0 + 8u * is);
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------
static int ring_geom(const FusedParams& p, RingGeom* r) {
  memset(r, 0, sizeof(*r));
  r->n_items = p.n_frames * p.grid;
  r->item_chunks = p.grid * p.g.cpt;
  r->stage_rows = kChunksPerStageTarget / p.g.cpt;
  if (r->stage_rows < 1) return 0;
  if (r->stage_rows > p.grid) r->stage_rows = p.grid;
  r->stage_chunks = r->stage_rows * p.g.cpt;
  r->n_stage = (p.grid + r->stage_rows - 1) / r->stage_rows;
  r->inv_n = 1.0 / ((double)p.patch * (double)p.patch);
  // four groups when their eight depth buffers leave most of the shared memory to the ring
  r->pool_groups = 4;
  r->depth_bufs = 1;
  if (const char* e = getenv("V3D_RING_POOL")) {  // experiments: "<groups><depth buffers>", e.g. 42
    r->pool_groups = (e[0] == '2') ? 2 : 4;
    r->depth_bufs = (e[1] == '2') ? 2 : 1;
  }
  r->span_limit = (int)min((unsigned long long)4095, 0xffffffffull / ((unsigned long long)p.patch * p.patch * 65535ull));
  return 1;
}

bool fused_ring_supported(const FusedParams& p, int depth_dtype, bool use_tma, int out_dtype) {
  if (depth_dtype != V3D_DEPTH_U16 || !use_tma) return false;
  if (p.pool_mode != V3D_POOL_AVG || p.clamp.on || p.coord_dtype != V3D_F32 || !p.disc.on || p.pooled_out != nullptr) return false;
  if (p.pe_mode != V3D_PE_SINE3D || p.table == nullptr || p.tokens_in == nullptr) return false;
  if (out_dtype != V3D_BF16 && out_dtype != V3D_F16) return false;
  if ((p.g.E & 7) != 0 || p.g.n_slots != 3 || p.patch > 64 || p.grid > 128) return false;
  if ((reinterpret_cast<uintptr_t>(p.tokens_in) & 15) != 0) return false;
  RingGeom r;
  if (!ring_geom(p, &r)) return false;
  // the pool groups' depth buffers and at least three ring slots must fit in one CTA's shared memory (large patches do not)
  r.slots = 3;
  return ring_smem(p, r).total <= (size_t)227 * 1024;
}

template <typename OutT, bool STRADDLE>
static int launch_ring(const FusedParams& p, RingGeom r, cudaStream_t s) {
  auto kern = k_fused_ring<OutT, STRADDLE>;
  static int max_smem = 0;
  if (max_smem == 0) {
    int dev = 0;
    V3D_CUDA(cudaGetDevice(&dev));
    V3D_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  // one CTA per SM with as many ring slots as fit (the slots are what keeps bytes in flight)
  r.slots = 3;
  for (int sl = kMaxRingSlots; sl >= 3; --sl) {
    r.slots = sl;
    if (ring_smem(p, r).total <= (size_t)max_smem) break;
  }
  const size_t smem = ring_smem(p, r).total;
  V3D_REQUIRE(smem <= (size_t)max_smem, V3D_E_UNSUPPORTED, "ring kernel needs %zu bytes of shared memory", smem);
  V3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = min(r.n_items, num_sms());
  kern<<<grid, kRingThreads, smem, s>>>(p, r);
  return check_launch("k_fused_ring");
}

int launch_fused_ring(const FusedParams& p, int out_dtype, cudaStream_t s) {
  RingGeom r;
  V3D_REQUIRE(ring_geom(p, &r), V3D_E_UNSUPPORTED, "ring kernel: embedding too wide");
  const bool straddle = (p.g.nf & 7) != 0;
  if (out_dtype == V3D_BF16)
    return straddle ? launch_ring<__nv_bfloat16, true>(p, r, s) : launch_ring<__nv_bfloat16, false>(p, r, s);
  return straddle ? launch_ring<__half, true>(p, r, s) : launch_ring<__half, false>(p, r, s);
}

}  // namespace v3d
