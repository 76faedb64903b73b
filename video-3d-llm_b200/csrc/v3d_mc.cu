// v3d_mc.cu -- maximum-coverage frame sampler on sm_100a (C ABI: include/v3d.h).
//
// Replaces scripts/3d/preprocessing/max_coverage_sampling.py:39-104:
//   k_mc_build    : open-addressing hash of the mesh universe pc_voxel (:65), clears bitsets
//   k_mc_voxelize_dense / k_mc_voxelize :
//                   unproject every pixel of every candidate frame, vox = int(round(w/voxel))
//                   (:44-45), look it up (a bit per cell of the universe's box in shared memory / the
//                   hash), set the frame's coverage bit -> per-frame bitsets = set(frame voxels) &
//                   pc_voxel (:47-50, :75)
//   k_mc_greedy   : one 8-CTA cluster per scene; each round = popcount(bits & ~used) per remaining frame,
//                   argmax with the ascending-order tie list (:72-82), CPython-compatible
//                   MT19937 random.choice (:84), used |= bits (:86) -- bit-exact selection.
#include <cooperative_groups.h>
#include <cstdlib>
#include <math_constants.h>

#include "v3d_common.cuh"

namespace v3d {

static constexpr int kMcThreads = 256;
static constexpr int kGreedyThreads = 512;
static constexpr int kGreedyCluster = 8;  // CTAs per scene (portable cluster size)
static constexpr int kMaxWordsSmem = 16384;  // `used` bitset lives in shared memory (64 KB)

__device__ __forceinline__ uint32_t hash3(int i, int j, int k) {
  uint32_t h = (uint32_t)i * 73856093u ^ (uint32_t)j * 19349663u ^ (uint32_t)k * 83492791u;
  h ^= h >> 15;
  h *= 0x2c1b3c6du;
  h ^= h >> 12;
  return h;
}

// ---- stage 0: clear bitsets, build the hash ---------------------------------------------------
__global__ void __launch_bounds__(kMcThreads) k_mc_clear(const v3d_mc_scene* __restrict__ scenes) {
  const v3d_mc_scene sc = scenes[blockIdx.y];
  const int words = (sc.n_pc + 31) / 32;
  const size_t nb = (size_t)sc.n_frames * words;
  for (size_t i = (size_t)blockIdx.x * kMcThreads + threadIdx.x; i < nb; i += (size_t)gridDim.x * kMcThreads) sc.bitsets[i] = 0u;
  int4* tab = reinterpret_cast<int4*>(sc.hash_table);
  for (size_t i = (size_t)blockIdx.x * kMcThreads + threadIdx.x; i < (size_t)sc.hash_slots; i += (size_t)gridDim.x * kMcThreads)
    tab[i] = make_int4(0, 0, 0, -1);
  // workspace of the dense voxeliser inside `record` (the greedy stage overwrites the whole record later):
  // [0..2] / [3..5] = smallest / largest voxel triple of the universe, [6] = the universe has duplicates,
  // [7] = frames the dense kernel left to the hash kernel
  if (blockIdx.x == 0 && threadIdx.x < 8)
    sc.record[threadIdx.x] = threadIdx.x < 3 ? 0x7fffffff : (threadIdx.x < 6 ? (int)0x80000000 : 0);
}

__global__ void __launch_bounds__(kMcThreads) k_mc_build(const v3d_mc_scene* __restrict__ scenes) {
  const v3d_mc_scene sc = scenes[blockIdx.y];
  const uint32_t mask = (uint32_t)sc.hash_slots - 1u;
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  for (int b = blockIdx.x * kMcThreads + threadIdx.x; b < sc.n_pc; b += gridDim.x * kMcThreads) {
    const int i = sc.pc_voxels[3 * b], j = sc.pc_voxels[3 * b + 1], k = sc.pc_voxels[3 * b + 2];
    lo[0] = min(lo[0], i), lo[1] = min(lo[1], j), lo[2] = min(lo[2], k);
    hi[0] = max(hi[0], i), hi[1] = max(hi[1], j), hi[2] = max(hi[2], k);
    uint32_t h = hash3(i, j, k) & mask;
    for (;;) {
      const int prev = atomicCAS(sc.hash_table + 4 * (size_t)h + 3, -1, b);
      if (prev == -1) {  // claimed: publish the triple (read only by later kernels)
        sc.hash_table[4 * (size_t)h + 0] = i;
        sc.hash_table[4 * (size_t)h + 1] = j;
        sc.hash_table[4 * (size_t)h + 2] = k;
        break;
      }
      // occupied: a duplicate of the same voxel keeps the first index (set semantics)
      if (sc.pc_voxels[3 * prev] == i && sc.pc_voxels[3 * prev + 1] == j && sc.pc_voxels[3 * prev + 2] == k) {
        sc.record[6] = 1;
        break;
      }
      h = (h + 1) & mask;
    }
  }
  // bounding box of the universe (all lanes take part: a lane without a voxel contributes the neutral element)
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int l = __reduce_min_sync(0xffffffffu, lo[a]), h = __reduce_max_sync(0xffffffffu, hi[a]);
    if ((threadIdx.x & 31) == 0 && l <= h) {
      atomicMin(sc.record + a, l);
      atomicMax(sc.record + 3 + a, h);
    }
  }
}

__device__ __forceinline__ int hash_find(const int4* __restrict__ tab, uint32_t mask, int i, int j, int k) {
  uint32_t h = hash3(i, j, k) & mask;
  for (;;) {
    const int4 e = __ldg(tab + h);
    if (e.w < 0) return -1;
    if (e.x == i && e.y == j && e.z == k) return e.w;
    h = (h + 1) & mask;
  }
}

// ---- stage 1: voxelise -------------------------------------------------------------------------
// torch: (w / voxel).round().int(); float->int of non-finite / out-of-range values is INT_MIN on x86
__device__ __forceinline__ int voxel_id(float w, const Divisor& voxel, float rcp, int div_mode) {
  float x = (div_mode == V3D_DIV_RECIPROCAL) ? __fmul_rn(w, rcp) : div_exact(w, voxel);
  x = rintf(x);
  return (x >= -2147483648.0f && x < 2147483648.0f) ? (int)x : (int)0x80000000;
}

static constexpr int kPxPerThread = 8;
#ifndef V3D_MC_MIN_BLOCKS
#define V3D_MC_MIN_BLOCKS 4
#endif

struct Vox {
  int i, j, k;
};
__device__ __forceinline__ bool vox_eq(const Vox& a, const Vox& b) { return a.i == b.i && a.j == b.j && a.k == b.k; }

// Candidate voxels are queued in shared memory so that the hash probes run with dense lanes (a probe is a
// dependent chain of L2 loads; run inline, 2-4 of 32 lanes would be active).
__device__ __forceinline__ void mc_push(int3* s_q, int* s_qn, const Vox& v) {
  const int pos = atomicAdd(s_qn, 1);
  s_q[pos] = make_int3(v.i, v.j, v.k);
}

__device__ __forceinline__ void mc_probe(const int4* __restrict__ tab, uint32_t mask, uint32_t* __restrict__ bits, int i, int j, int k) {
  const int b = hash_find(tab, mask, i, j, k);
  if (b >= 0) {
    const uint32_t m = 1u << (b & 31);
    if (!(bits[b >> 5] & m)) atomicOr(bits + (b >> 5), m);
  }
}

// ---- approximate-then-verify voxel ids ------------------------------------------------------------------
// For a rigid pose the voxel coordinate of a pixel is, in real arithmetic,
//     t_i = d * (A_i*um + B_i*vm + C_i) + D_i,   A_i = P_i0/(fx*scale*voxel), B_i = P_i1/(fy*scale*voxel),
//                                                 C_i = P_i2/(scale*voxel),    D_i = P_i3/voxel,
// six FMAs per pixel instead of the reference's 3 divisions + 12 pose operations + 3 divisions.  The reference
// result q_i (every operation rounded to fp32, video_utils.py:57-65 then max_coverage_sampling.py:44) deviates
// from t_i by at most u*(9*S + 3*|D_i|), u = 2^-24, S = d*(|A_i||um| + |B_i||vm| + |C_i|) (three roundings
// per camera-space coordinate, four pose operations, the final division); the FMA evaluation with coefficients
// rounded from fp64 deviates by at most u*(8*S + 2*|D_i|).  With k1 = 40u*max_i(|A_i|um_max + |B_i|vm_max +
// |C_i|) and k0 = 12u*max_i|D_i| + 2^-20 (2.3x the sum of the two bounds) a pixel whose three t_i are farther than
// k1*d + k0 from every half-integer has rint(q_i) == rint(t_i); the others (2-3 of 10^4 on ScanNet-shaped
// scenes, tools/check_mc_guard_band.py) run the exact operator sequence.  rint is the 1.5*2^23 add/subtract
// (round-half-even for |t| < 2^22, which the frame test guarantees for every uint16 depth).
struct McSetup {
  float coef[12];  // A_i, B_i, C_i, D_i for i = x, y, z
  float k1, half_minus_k0;
  int fast;
  Camera cam;  // for the exact path
  Divisor ds, dv;
  float rcp;
  int div_mode;
};

static constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
static constexpr int kWarpQueue = 32 + 32 * kPxPerThread;  // leftovers of the previous iteration + one iteration

// `base` (NULL or three integers): the estimate is made in the shifted frame t_i - base_i (the dense voxeliser's cell
// coordinates).  The shifted constant D_i - base_i is rounded from fp64 like D_i, so the bound above holds with
// max(|D_i|, |D_i - base_i|) in the place of |D_i|.
__device__ __noinline__ void mc_setup_frame(McSetup* s, const v3d_mc_scene& sc, int f, int H, int W, float depth_scale, float voxel, int div_mode,
                               bool layout_ok, const int* base = nullptr) {
  const Camera cam = load_camera(sc.poses, sc.intr, f);
  s->cam = cam;
  s->ds = make_divisor(depth_scale);
  s->dv = make_divisor(voxel);
  s->rcp = 1.0f / voxel;
  s->div_mode = div_mode;
  const double sv = (double)depth_scale * (double)voxel;
  const double den[4] = {(double)cam.fx.b * sv, (double)cam.fy.b * sv, sv, (double)voxel};
  const double um_max = fmax(fabs(0.0 - (double)cam.cx), fabs((double)(W - 1) - (double)cam.cx));
  const double vm_max = fmax(fabs(0.0 - (double)cam.cy), fabs((double)(H - 1) - (double)cam.cy));
  double r_max = 0.0, d_max = 0.0, shifted_max = 0.0;
  bool finite = true;
  for (int i = 0; i < 3; ++i) {
    double c[4];
    for (int j = 0; j < 4; ++j) {
      c[j] = (double)cam.P[4 * i + j] / den[j];
      finite = finite && (fabs(c[j]) < 1.0e30);
    }
    d_max = fmax(d_max, fabs(c[3]));
    if (base) {
      c[3] -= (double)base[i];
      shifted_max = fmax(shifted_max, fabs(c[3]));
    }
    for (int j = 0; j < 4; ++j) s->coef[4 * i + j] = (float)c[j];
    r_max = fmax(r_max, fabs(c[0]) * um_max + fabs(c[1]) * vm_max + fabs(c[2]));
    d_max = fmax(d_max, fabs(c[3]));
  }
  const double u = 1.0 / 16777216.0;
  const double k1 = 40.0 * u * r_max, k0 = 12.0 * u * d_max + 1.0 / 1048576.0;
  s->k1 = (float)(k1 * 1.000001);
  s->half_minus_k0 = (float)(0.5 - k0 * 1.000001);
  // every |t| stays below 2^21 for any uint16 depth; divisors far inside the normal range (no overflow or
  // underflow inside the reference sequence that the bound would not see)
  s->fast = layout_ok && finite && cam.rigid && s->ds.ok && cam.fx.ok && cam.fy.ok && s->dv.ok && (65535.0 * r_max + d_max < 2097152.0) &&
            (um_max < 1.0e6) && (vm_max < 1.0e6);
  // The dense voxeliser's box test (one compare for x and y together, mc_dense_part) is exact only while no coordinate
  // can carry a bit at or above 2^12, where the bits of the bitset's address sit in x: every cell coordinate any uint16
  // depth can produce must stay below 4096 in magnitude (a 65.5 m ray is 655 cells at 0.1 m).  Other frames are not
  // `fast` for either kernel and take the exact path of k_mc_voxelize.
  if (base) s->fast = s->fast && (65535.0 * r_max + shifted_max < 4095.0);
}

// exact operator sequence of one pixel (the guard-band pixels), probed right away
__device__ __noinline__ void mc_exact_probe(const McSetup* s, const v3d_mc_scene* __restrict__ scp, int f, uint32_t raw, int col, int row) {
  const Camera cam = s->cam;
  float x, y, z;
  unproject_px(cam, (float)raw, __fsub_rn((float)col, cam.cx), __fsub_rn((float)row, cam.cy), s->ds, x, y, z);
  const int i = voxel_id(x, s->dv, s->rcp, s->div_mode), j = voxel_id(y, s->dv, s->rcp, s->div_mode), k = voxel_id(z, s->dv, s->rcp, s->div_mode);
  mc_probe(reinterpret_cast<const int4*>(scp->hash_table), (uint32_t)scp->hash_slots - 1u,
           scp->bitsets + (size_t)f * ((scp->n_pc + 31) / 32), i, j, k);
}

// Append (x,y,z) to the warp's queue if it differs from (px,py,pz): predicated, no branch (some lane of the warp
// appends on 99 % of the steps).
__device__ __forceinline__ void mc_push_if_changed(uint32_t q_addr, uint32_t cnt_addr, float x, float y, float z, float px, float py,
                                                   float pz) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 pos, a;\n\t"
      "mov.u32 pos, 0;\n\t"
      "setp.neu.f32 p, %2, %5;\n\t"
      "setp.neu.or.f32 p, %3, %6, p;\n\t"
      "setp.neu.or.f32 p, %4, %7, p;\n\t"
      "@p atom.shared.add.u32 pos, [%1], 1;\n\t"
      "mad.lo.u32 a, pos, 12, %0;\n\t"
      "@p st.shared.f32 [a], %2;\n\t"
      "@p st.shared.f32 [a+4], %3;\n\t"
      "@p st.shared.f32 [a+8], %4;\n\t"
      "}" ::"r"(q_addr),
      "r"(cnt_addr), "f"(x), "f"(y), "f"(z), "f"(px), "f"(py), "f"(pz)
      : "memory");
}

struct McPixel {
  float n[3];
  bool ok;
};
// voxel estimate of one pixel: n = rint(t), ok = every t is outside the guard band
__device__ __forceinline__ McPixel mc_estimate(float df, const float (&ray)[3], const float (&D)[3], float k1n, float hmk0) {
  McPixel r;
  const float m = __fmaf_rn(df, k1n, hmk0);
  r.ok = true;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float t = __fmaf_rn(df, ray[i], D[i]);
    r.n[i] = __fsub_rn(__fadd_rn(t, kMagic), kMagic);
    r.ok = r.ok && (fabsf(__fsub_rn(t, r.n[i])) < m);
  }
  return r;
}

// One warp voxelises rows [rw0, rw1) of frame f: 8 pixels per lane from one 128-bit load, run-length dedup
// against the previous pixel (for the first pixel of a group: the left lane's last pixel, re-estimated from its
// shuffled depth), candidates appended to the warp's own shared-memory queue (no block barriers) and probed
// 32 at a time from the top of the queue.  A guard-band pixel is entered as NaN (never equal to a neighbour, skipped
// by the probes) and evaluated exactly after the eight steps.
__device__ __forceinline__ void mc_rows_fast(const McSetup* __restrict__ s, const v3d_mc_scene* __restrict__ scp, int f, int H, int W,
                                             int rw0, int rw1, float* __restrict__ q, int* __restrict__ q_cnt) {
  const int lane = threadIdx.x & 31;
  const int gpr = W / kPxPerThread;
  const int n_groups = (rw1 - rw0) * gpr;
  if (n_groups <= 0) return;
  const uint32_t gpr_magic = 0xffffffffu / (uint32_t)gpr + 1u;  // g / gpr == umulhi(g, magic) for g < 2^16
  float A[3], B[3], Cc[3], D[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A[i] = s->coef[4 * i];
    B[i] = s->coef[4 * i + 1];
    Cc[i] = s->coef[4 * i + 2];
    D[i] = s->coef[4 * i + 3];
  }
  const float k1n = -s->k1, hmk0 = s->half_minus_k0, cx = s->cam.cx, cy = s->cam.cy;
  uint32_t q_addr = smem_u32(q), cnt_addr = smem_u32(q_cnt);
  // opaque to the compiler: keep these in registers instead of re-deriving them from %tid at every use
  asm volatile("" : "+r"(q_addr), "+r"(cnt_addr));
  const uint16_t* frame = reinterpret_cast<const uint16_t*>(scp->depth) + ((size_t)f * H + rw0) * W;
  // lanes past the end of a ragged last iteration redo the last group (duplicate candidates are harmless)
  int g = min(lane, n_groups - 1);
  int r = (int)__umulhi((uint32_t)g, gpr_magic);
  int c0 = (g - r * gpr) * kPxPerThread;
  uint4 w = ld_stream_v4(frame + (size_t)r * W + c0);
  for (int g0 = 0; g0 < n_groups; g0 += 32) {
    const uint32_t w4[4] = {w.x, w.y, w.z, w.w};
    const float vm = __fsub_rn((float)(rw0 + r), cy), um0 = __fsub_rn((float)c0, cx);
    float ray0[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) ray0[i] = __fmaf_rn(A[i], um0, __fmaf_rn(B[i], vm, Cc[i]));
    // previous pixel of the group's first pixel: the left lane's last pixel (same row when c0 > 0)
    float px, py, pz;
    {
      const uint32_t left = __shfl_up_sync(0xffffffffu, w4[3], 1);
      const float df = (float)(left >> 16);
      float ray[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) ray[i] = __fsub_rn(ray0[i], A[i]);
      const McPixel p = mc_estimate(df, ray, D, k1n, hmk0);
      const bool usable = p.ok && lane != 0 && c0 != 0;
      px = usable ? p.n[0] : __int_as_float(0x7fc00000);
      py = p.n[1];
      pz = p.n[2];
    }
    uint32_t slow = 0;
#pragma unroll
    for (int e = 0; e < kPxPerThread; ++e) {
      const float df = (float)((e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xffffu));  // I2F.U16 reads either half
      float ray[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) ray[i] = __fmaf_rn(A[i], (float)e, ray0[i]);
      const McPixel p = mc_estimate(df, ray, D, k1n, hmk0);
      const float nx = p.ok ? p.n[0] : __int_as_float(0x7fc00000);
      slow |= p.ok ? 0u : (1u << e);
      mc_push_if_changed(q_addr, cnt_addr, nx, p.n[1], p.n[2], px, py, pz);
      px = nx;
      py = p.n[1];
      pz = p.n[2];
    }
    if (slow) {  // 2-3 pixels in 10^4: exact operator sequence, probed directly
#pragma unroll 1
      for (uint32_t rest = slow; rest; rest &= rest - 1) {
        const int e = __ffs(rest) - 1;
        const uint32_t lo = (e & 4) ? w4[2] : w4[0], hi = (e & 4) ? w4[3] : w4[1];
        const uint32_t word = (e & 2) ? hi : lo;
        mc_exact_probe(s, scp, f, (e & 1) ? (word >> 16) : (word & 0xffffu), c0 + e, rw0 + r);
      }
    }
    // next iteration's position; its load is in flight during the probes below
    const bool last = g0 + 32 >= n_groups;
    if (!last) {
      g = min(g0 + 32 + lane, n_groups - 1);
      r = (int)__umulhi((uint32_t)g, gpr_magic);
      c0 = (g - r * gpr) * kPxPerThread;
      w = ld_stream_v4(frame + (size_t)r * W + c0);
    }
    __syncwarp();
    // probe whole batches of 32 from the top of the queue; fewer than 32 candidates stay for the next iteration
    const int qn = *q_cnt;
    const int keep = last ? 0 : (qn & 31);
    if (qn > keep) {
      const int4* tab = reinterpret_cast<const int4*>(scp->hash_table);
      const uint32_t mask = (uint32_t)scp->hash_slots - 1u;
      uint32_t* bits = scp->bitsets + (size_t)f * ((scp->n_pc + 31) / 32);
      for (int i = keep + lane; i < qn; i += 32) {
        const float* v = q + 3 * i;
        const float vx = v[0];
        if (vx == vx) mc_probe(tab, mask, bits, (int)vx, (int)v[1], (int)v[2]);
      }
      __syncwarp();
      if (lane == 0) *q_cnt = keep;
    }
    __syncwarp();
  }
}

template <int DT>
__device__ __forceinline__ void mc_slab_generic(const v3d_mc_scene& sc, const McSetup* __restrict__ s, int f, int H, int W, int r0, int r1,
                                                int3* s_q, int* s_qn, int& phase, const int4* __restrict__ tab, uint32_t mask,
                                                uint32_t* __restrict__ bits) {
  const Camera cam = s->cam;
  const Divisor ds = s->ds, dv = s->dv;
  const float rcp = s->rcp;
  const int div_mode = s->div_mode;
  const int groups_per_row = (W + kPxPerThread - 1) / kPxPerThread;
  const int lane = threadIdx.x & 31;
  const int n_groups = (r1 - r0) * groups_per_row;
  // all threads of the block iterate together (barriers and shuffles below); whole warps past the end idle
  // two queue counters used alternately: the idle one is cleared between the barriers of an iteration
  for (int g0 = 0; g0 < n_groups; g0 += kMcThreads, phase ^= 1) {
    int* cnt = s_qn + phase;
    const int g = g0 + threadIdx.x;
    const bool valid = g < n_groups;
    Vox first = {0x7fffffff, 0x7fffffff, 0x7fffffff}, prev = first;
    int n_px = 0, c0 = 0;
    if (valid) {
      const int gr = g / groups_per_row;
      const int r = r0 + gr;
      c0 = (g - gr * groups_per_row) * kPxPerThread;
      n_px = min(kPxPerThread, W - c0);
      const float vm = __fsub_rn((float)r, cam.cy);
      const size_t base = ((size_t)f * H + r) * W + c0;
#pragma unroll 1
      for (int e = 0; e < n_px; ++e) {
        float x, y, z;
        unproject_px(cam, load_depth<DT>(sc.depth, base + e), __fsub_rn((float)(c0 + e), cam.cx), vm, ds, x, y, z);
        const Vox v = {voxel_id(x, dv, rcp, div_mode), voxel_id(y, dv, rcp, div_mode), voxel_id(z, dv, rcp, div_mode)};
        if (e == 0) first = v;
        else if (!vox_eq(v, prev)) mc_push(s_q, cnt, v);
        prev = v;
      }
    }
    Vox left;
    left.i = __shfl_up_sync(0xffffffffu, prev.i, 1);
    left.j = __shfl_up_sync(0xffffffffu, prev.j, 1);
    left.k = __shfl_up_sync(0xffffffffu, prev.k, 1);
    const int left_px = __shfl_up_sync(0xffffffffu, n_px, 1);
    if (n_px > 0 && (lane == 0 || left_px != kPxPerThread || c0 == 0 || !vox_eq(first, left))) mc_push(s_q, cnt, first);
    __syncthreads();
    const int qn = *cnt;
    if (threadIdx.x == 0) s_qn[phase ^ 1] = 0;
    for (int i = threadIdx.x; i < qn; i += kMcThreads) {
      const int3 v = s_q[i];
      mc_probe(tab, mask, bits, v.x, v.y, v.z);
    }
    __syncthreads();
  }
}

// ---- dense voxeliser ---------------------------------------------------------------------------------------
// The universe of a scene (mesh voxels of one room) sits in a small box: 128 x 128 x 33 cells for a 12.8 m x 12.8 m x 3.3 m
// room at 0.1 m.  One bit per CELL of that box fits in shared memory, so a pixel needs no hash probe, no queue and no
// global atomic: estimate the cell (the same verified FMA estimate as above, made in the box's own frame), test that it
// lies inside the box, set its bit with one shared-memory atomic unless the previous pixel of the lane set the same one.
// After the last row of a frame the CTA walks the universe once (n_pc cells) and ORs the coverage words into the frame's
// bitset.  Pixels outside the box cannot be in the universe and are dropped; guard-band pixels take the exact operator
// sequence and the hash probe as before.  Scenes whose box does not fit (or whose frames are not `fast`) are left to
// k_mc_voxelize, which skips everything this kernel has done -- both derive the split from mc_dense_plan.
// 512 threads x 2 CTAs per SM at 64 registers is the measured best (384 x 2 at 80 registers and one CTA per SM are slower).
#ifndef V3D_DENSE_THREADS
#define V3D_DENSE_THREADS 512
#endif
#ifndef V3D_DENSE_BLOCKS
#define V3D_DENSE_BLOCKS 2
#endif
static constexpr int kDenseThreads = V3D_DENSE_THREADS;
static constexpr int kDenseRows = 2;  // rows per warp in one item of a CTA's range: items are blocks of 32 rows
static constexpr int kDenseSmemBytes = (V3D_DENSE_BLOCKS == 1 ? 160 : 80) * 1024;  // cell bitset; two CTAs per SM with their staging rings
static constexpr uint32_t kMagicBits = 0x4B400000u;  // bits of kMagic: (t + kMagic) has the bits kMagicBits + rint(t)

struct McDense {
  int ok;
  int base[3];   // smallest voxel triple of the universe
  int px, pxy;   // cells per x row (power of two), per z slice (px * py, py a power of two)
  int dimz;
  int words;     // 32-bit words of the cell bitset, a multiple of 4
};

__device__ __forceinline__ int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

__device__ __forceinline__ McDense mc_dense_plan(const v3d_mc_scene& sc, int smem_bits, int frame_pixels) {
  McDense d;
  d.ok = 0;
  if (sc.n_pc <= 0 || smem_bits <= 0) return d;
  // the walk over the universe after every frame must stay small next to the frame itself
  if (sc.n_pc > 65536 && sc.n_pc > frame_pixels / 2) return d;
  long ext[3];
  for (int a = 0; a < 3; ++a) {
    const int lo = sc.record[a], hi = sc.record[3 + a];
    if (lo > hi || lo < -(1 << 20) || hi > (1 << 20)) return d;
    d.base[a] = lo;
    ext[a] = (long)hi - lo + 1;
  }
  if (ext[0] > 4096 || ext[1] > 4096 || ext[2] > 4096) return d;
  d.px = pow2_ceil((int)ext[0]);
  d.pxy = d.px * pow2_ceil((int)ext[1]);
  d.dimz = (int)ext[2];
  const long bits = (long)d.pxy * d.dimz;
  if (bits > smem_bits - 8 * 1024) return d;  // alignment of the bitset (512 B) and, behind it, the 128 B of dummy words of mc_dense_set
  d.words = (int)((bits + 127) / 128) * 4;
  d.ok = 1;
  return d;
}

// Set bit (key & 31) of the shared-memory word at byte address (key >> 3) & 0x7fffc unless key is the invalid key (all
// ones) or equal to the previous pixel's.  No predicate on the atomic (ptxas wraps a predicated shared atomic into a
// divergent branch: three more issue slots per pixel): a lane with nothing to set ORs into its own dummy word.
__device__ __forceinline__ void mc_dense_set(uint32_t key, uint32_t prev, uint32_t dummy_addr) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      ".reg .u32 a, b;\n\t"
      "setp.ne.u32 q, %0, %1;\n\t"
      "setp.ne.and.u32 q, %0, 0xffffffff, q;\n\t"
      "shr.u32 a, %0, 3;\n\t"
      "and.b32 a, a, 0x7fffc;\n\t"
      "selp.u32 a, a, %2, q;\n\t"
      "shf.l.wrap.b32 b, 0, 1, %0;\n\t"  // (1 << (key & 31)): high word of {1, 0} << n
      "red.shared.or.b32 [a], b;\n\t"
      "}" ::"r"(key),
      "r"(prev), "r"(dummy_addr)
      : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}

// packed fp32 pairs (sm_100: fma/add/sub.rn.f32x2)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float half2f(uint64_t v, int hi) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return hi ? b : a;
}
__device__ __forceinline__ uint32_t half2u(uint64_t v, int hi) { return __float_as_uint(half2f(v, hi)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

static constexpr int kDensePx = 16;        // pixels of a lane per warp iteration: two 16-byte pieces
static constexpr int kDenseTaskIters = 2;  // warp iterations per task: 2 x 32 x 16 = 1024 pixels

// One warp voxelises its share of rows [ra, rb) of frame f into the CTA's cell bitset.  The rows are one flat run of
// 16-pixel chunks; warps take tasks of 64 chunks from a shared counter as they finish (no end-of-frame skew), and a lane
// stages the chunk of its next iteration -- also across tasks -- in shared memory with cp.async while it works on the
// current one (no registers held by the load in flight).  gpr = W / 16 chunks per row, gpr_magic its reciprocal.
__device__ __forceinline__ void mc_dense_part(const McSetup* __restrict__ s, const v3d_mc_scene* __restrict__ scp, const McDense& dn, int f, int H,
                                              int W, int gpr, uint32_t gpr_magic, int ra, int rb, uint32_t* __restrict__ s_bits,
                                              int* __restrict__ s_next, uint4* __restrict__ stage) {
  const int lane = threadIdx.x & 31;
  const int n = (rb - ra) * gpr;
  float A[3], D[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A[i] = s->coef[4 * i];
    D[i] = s->coef[4 * i + 3];
  }
  const float k1n = -s->k1, hmk0 = s->half_minus_k0;
  const uint32_t mx = ~(uint32_t)(dn.px - 1), my = ~(uint32_t)(dn.pxy / dn.px - 1);
  const uint32_t px = (uint32_t)dn.px, pxy = (uint32_t)dn.pxy, dimz = (uint32_t)dn.dimz;
  // The x estimate is rounded with kMagic + 8*sbase (sbase = shared-memory address of the bitset, a multiple of 512 below
  // 2^18): exact like kMagic itself (integers below 2^24), and the bits of the sum are kMagicBits + 8*sbase + x -- the
  // key below then holds the word's address without another add.
  const uint32_t sbase8 = smem_u32(s_bits) << 3;
  float magic_x = kMagic + (float)sbase8;
  uint32_t bits_x = kMagicBits | sbase8;
  uint32_t dummy = smem_u32(s_bits) + (uint32_t)dn.words * 4u + (uint32_t)lane * 4u;
  // a stage = two planes of 512 bytes: piece p (pixels 8p..8p+7) of lane l at p*512 + l*16
  uint32_t st_addr = smem_u32(stage) + (uint32_t)lane * 16u;
  const uint16_t* rows = reinterpret_cast<const uint16_t*>(scp->depth) + ((size_t)f * H + ra) * W;
  // plain registers (otherwise re-derived from uniform registers and special registers in every iteration)
  asm volatile("" : "+r"(dummy), "+r"(st_addr), "+f"(magic_x), "+r"(bits_x), "+l"(rows));
  int cur;  // first chunk of the warp's task
  {
    int t = 0;
    if (lane == 0) t = atomicAdd(s_next, 1);
    cur = __shfl_sync(0xffffffffu, t, 0) * (32 * kDenseTaskIters);
  }
  if (cur >= n) return;
  // lanes past the end of the run redo its last chunk (setting a bit twice is harmless)
  {
    const uint16_t* src = rows + (size_t)min(cur + lane, n - 1) * kDensePx;
    cp_async16(st_addr, src);
    cp_async16(st_addr + 512u, src + 8);
    cp_async_commit();
  }
  uint32_t k = 0;
  while (cur < n) {
    int nxt;
    {
      int t = 0;
      if (lane == 0) t = atomicAdd(s_next, 1);
      nxt = __shfl_sync(0xffffffffu, t, 0) * (32 * kDenseTaskIters);
    }
    const int iters = min(kDenseTaskIters, (n - cur + 31) / 32);
#pragma unroll 1
    for (int it = 0; it < iters; ++it, ++k) {
      const int pf = it + 1 < iters ? cur + (it + 1) * 32 : nxt;
      if (pf < n) {
        const uint16_t* src = rows + (size_t)min(pf + lane, n - 1) * kDensePx;
        const uint32_t dst = st_addr + (((k + 1u) & 1u) << 10);
        cp_async16(dst, src);
        cp_async16(dst + 512u, src + 8);
      }
      cp_async_commit();
      cp_async_wait<1>();
      const uint32_t st_cur = st_addr + ((k & 1u) << 10);
      const int ch = min(cur + it * 32 + lane, n - 1);
      const int r = (int)__umulhi((uint32_t)ch, gpr_magic);
      const int c0 = (ch - r * gpr) * kDensePx;
      const int y0 = ra;
      float ray0[3];
      {
        const float vm = __fsub_rn((float)(y0 + r), s->cam.cy), um0 = __fsub_rn((float)c0, s->cam.cx);
#pragma unroll
        for (int i = 0; i < 3; ++i) ray0[i] = __fmaf_rn(A[i], um0, __fmaf_rn(s->coef[4 * i + 1], vm, s->coef[4 * i + 2]));
      }
      uint32_t prev = 0xffffffffu, slow = 0;
      // Four quarters of four pixels, not unrolled (more pixels in flight need more than 64 registers); inside a quarter
      // the pixels go in pairs through the packed fp32 instructions (fma/add/sub.rn.f32x2, SASS FFMA2 / FADD2: two lanes
      // of the same IEEE operation, so the estimate is bit for bit mc_estimate's) -- half the issue slots for the arithmetic.
#pragma unroll 1
      for (int h = 0; h < kDensePx / 4; ++h) {
        uint32_t wa, wb;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wa), "=r"(wb) : "r"(st_cur + ((uint32_t)(h >> 1) << 9) + ((uint32_t)(h & 1) << 3)) : "memory");
        const float fh = (float)(4 * h);
        slow <<= 4;  // in the end bits 15..12 = pixels 0..3, ..., bits 3..0 = pixels 12..15
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t word = j ? wb : wa;
          const uint64_t df2 = pack2((float)(word & 0xffffu), (float)(word >> 16));
          const uint64_t fe2 = pack2(fh + (float)(2 * j), fh + (float)(2 * j + 1));  // exact
          const uint64_t m2 = fma2(df2, pack2(k1n, k1n), pack2(hmk0, hmk0));
          uint64_t t2[3], tm2[3], dv2[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float mg = i == 0 ? magic_x : kMagic;
            t2[i] = fma2(df2, fma2(pack2(A[i], A[i]), fe2, pack2(ray0[i], ray0[i])), pack2(D[i], D[i]));
            tm2[i] = add2(t2[i], pack2(mg, mg));
            dv2[i] = sub2(t2[i], sub2(tm2[i], pack2(mg, mg)));
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int e = 2 * j + q;
            const float dev = fmaxf(fmaxf(fabsf(half2f(dv2[0], q)), fabsf(half2f(dv2[1], q))), fabsf(half2f(dv2[2], q)));
            const bool ok = dev < half2f(m2, q);
            if (!ok) slow |= 1u << e;
            // tm = kMagicBits + cell coordinate (+ 8*sbase for x, a multiple of 4096): inside the box iff x and y have no
            // bit above their power-of-two extent other than those constants and 0 <= z < dimz.  The fast-frame test
            // keeps every |coordinate| below 4096: a coordinate outside its extent then shows a bit below 2^12 (too
            // large) or clears bit 22 and sets bits 12..21 (negative) -- neither can be mistaken for the bits of
            // 8*sbase, which are all at or above 2^12 and never all of 12..21
            const uint32_t bx = half2u(tm2[0], q), by = half2u(tm2[1], q), bz = half2u(tm2[2], q);
            const bool good = ok && (((bx & mx) | (by & my)) == bits_x) && ((bz - kMagicBits) < dimz);
            // kMagicBits * (1 + px + pxy) is a multiple of 2^22: the low 22 bits of the key are 8*sbase + x + px*y + pxy*z,
            // i.e. 8 * (byte address of the cell's word) + bit number
            const uint32_t key = good ? bx + by * px + bz * pxy : 0xffffffffu;
            V3D_CHECK(!good || (((key >> 3) & 0x7fffcu) - smem_u32(s_bits)) < (uint32_t)dn.words * 4u);
            mc_dense_set(key, prev, dummy);
            prev = key;
          }
        }
      }
      if (slow) {  // 2-3 pixels in 10^4: exact operator sequence, probed through the hash
#pragma unroll 1
        for (uint32_t rest = slow; rest; rest &= rest - 1) {
          const int e = (__ffs(rest) - 1) ^ 12;
          uint32_t raw;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(raw) : "r"(st_cur + ((uint32_t)(e >> 3) << 9) + ((uint32_t)(e & 7) << 1)) : "memory");
          mc_exact_probe(s, scp, f, raw, c0 + e, y0 + r);
        }
      }
    }
    cur = nxt;
  }
  cp_async_wait<0>();
}

// Coverage words of frame f from the cell bitset, in universe order; clears the cell bitset afterwards.
__device__ __forceinline__ void mc_dense_flush(const v3d_mc_scene& sc, const McDense& dn, int f, uint32_t* __restrict__ s_bits, int has_dup) {
  __syncthreads();  // every warp has set its bits
  const int lane = threadIdx.x & 31;
  uint32_t* bits = sc.bitsets + (size_t)f * ((sc.n_pc + 31) / 32);
#pragma unroll 2
  for (int p0 = 0; p0 < sc.n_pc; p0 += kDenseThreads) {
    const int p = p0 + threadIdx.x;
    bool b = false;
    if (p < sc.n_pc) {
      const int i = sc.pc_voxels[3 * p], j = sc.pc_voxels[3 * p + 1], k = sc.pc_voxels[3 * p + 2];
      const uint32_t cell = (uint32_t)(i - dn.base[0]) + (uint32_t)(j - dn.base[1]) * (uint32_t)dn.px + (uint32_t)(k - dn.base[2]) * (uint32_t)dn.pxy;
      V3D_CHECK(cell < (uint32_t)dn.words * 32u);
      b = (s_bits[cell >> 5] >> (cell & 31u)) & 1u;
      // a duplicated voxel owns one bit: that of the index the hash kept
      if (b && has_dup) b = hash_find(reinterpret_cast<const int4*>(sc.hash_table), (uint32_t)sc.hash_slots - 1u, i, j, k) == p;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, b);
    if (lane == 0 && word) atomicOr(bits + (p >> 5), word);
  }
  __syncthreads();
  uint4* z = reinterpret_cast<uint4*>(s_bits);
  for (int i = threadIdx.x; i < dn.words / 4; i += kDenseThreads) z[i] = make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(kDenseThreads, V3D_DENSE_BLOCKS) k_mc_voxelize_dense(const v3d_mc_scene* __restrict__ scenes, int n_scenes, int H, int W,
                                                                         float depth_scale, float voxel, int div_mode, int items_per_frame,
                                                                         int rows_per_item, long total_items, int smem_bits,
                                                                         uint32_t gpr_magic) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  // 512-byte aligned: 8 * (shared-memory address of the bitset) is then a multiple of every power-of-two x extent up to
  // 4096, which lets the x coordinate carry the address (mc_dense_part)
  uint32_t* s_bits = s_dyn + ((512u - (smem_u32(s_dyn) & 511u)) & 511u) / 4u;
  __shared__ McSetup s_set;
  __shared__ int s_next;  // next warp task of the frame: warps take tasks as they finish
  __shared__ __align__(16) uint4 s_stage[kDenseThreads / 32][2][64];
  // a CTA takes a contiguous range of the (scene, frame, row block) items of the whole batch
  const long item0 = total_items * blockIdx.x / gridDim.x, item1 = total_items * (blockIdx.x + 1) / gridDim.x;
  long item = item0, first = 0;  // first = first item of scene s
  int s = 0;
  while (item < item1 && s < n_scenes) {
    const long cnt = (long)scenes[s].n_frames * items_per_frame;
    if (item >= first + cnt) {
      first += cnt;
      ++s;
      continue;
    }
    const long send = item1 < first + cnt ? item1 : first + cnt;
    const v3d_mc_scene sc = scenes[s];
    const McDense dn = mc_dense_plan(sc, smem_bits, H * W);
    if (!dn.ok) {  // k_mc_voxelize takes the whole scene
      item = send;
      continue;
    }
    const bool layout_ok = (W % kPxPerThread) == 0 && (reinterpret_cast<uintptr_t>(sc.depth) & 15) == 0;
    const int has_dup = sc.record[6];
    __syncthreads();
    {
      uint4* z = reinterpret_cast<uint4*>(s_bits);
      for (int i = threadIdx.x; i < dn.words / 4; i += kDenseThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    while (item < send) {
      const int li = (int)(item - first);
      const int f = li / items_per_frame, i0 = li - f * items_per_frame;
      const long left = send - item;
      const int i1 = left < (long)(items_per_frame - i0) ? i0 + (int)left : items_per_frame;
      const int ra = i0 * rows_per_item, rb = min(i1 * rows_per_item, H);
      __syncthreads();  // the previous frame's readers of s_set are done; the cleared bitset is visible
      if (threadIdx.x == 0) {
        mc_setup_frame(&s_set, sc, f, H, W, depth_scale, voxel, div_mode, layout_ok, dn.base);
        s_next = 0;
      }
      __syncthreads();
      if (s_set.fast) {
        mc_dense_part(&s_set, scenes + s, dn, f, H, W, W / kDensePx, gpr_magic, ra, rb, s_bits, &s_next, &s_stage[threadIdx.x >> 5][0][0]);
        mc_dense_flush(sc, dn, f, s_bits, has_dup);
      } else if (threadIdx.x == 0 && i0 == 0) {
        atomicAdd(sc.record + 7, 1);  // a frame for k_mc_voxelize
      }
      item += i1 - i0;
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(kMcThreads, V3D_MC_MIN_BLOCKS) k_mc_voxelize(const v3d_mc_scene* __restrict__ scenes, int H, int W,
                                                            float depth_scale, float voxel, int div_mode, int slabs_per_frame,
                                                            int rows_per_slab, int dense_smem_bits) {
  const v3d_mc_scene sc = scenes[blockIdx.y];
  // what k_mc_voxelize_dense has done already (dense_smem_bits = 0: it did not run): every `fast` frame of a scene
  // whose universe box fits its shared-memory cell bitset
  const McDense dn = mc_dense_plan(sc, dense_smem_bits, H * W);
  if (dn.ok && sc.record[7] == 0) return;
  const int words = (sc.n_pc + 31) / 32;
  const uint32_t mask = (uint32_t)sc.hash_slots - 1u;
  const int4* tab = reinterpret_cast<const int4*>(sc.hash_table);
  const int warp = threadIdx.x >> 5;
  __shared__ __align__(16) float s_q[(kMcThreads / 32) * kWarpQueue * 3];  // 27 KB; the generic path uses it as one int3 queue
  __shared__ McSetup s_set;
  __shared__ int s_qn[2];
  __shared__ int s_wqn[kMcThreads / 32];
  int phase = 0;
  static_assert(sizeof(float) * (kMcThreads / 32) * kWarpQueue * 3 >= sizeof(int3) * kMcThreads * kPxPerThread, "queue size");
  // 8 pixels per 128-bit load: rows must be whole groups and 16-byte aligned
  const bool layout_ok = DT == V3D_DEPTH_U16 && (W % kPxPerThread) == 0 && (reinterpret_cast<uintptr_t>(sc.depth) & 15) == 0;
  const int rows_per_warp = (rows_per_slab + kMcThreads / 32 - 1) / (kMcThreads / 32);
  if (threadIdx.x == 0) s_qn[0] = s_qn[1] = 0;
  if (threadIdx.x < kMcThreads / 32) s_wqn[threadIdx.x] = 0;
  // A CTA takes a CONTIGUOUS range of (frame, slab) items: consecutive slabs of a frame share the frame's setup (camera,
  // fp64 coefficient divisions, guard-band constants: ~2 k cycles on one thread with the other 255 waiting), which
  // strided items recomputed for every slab -- 12 % of all stall samples sat at these two barriers.
  const long total_items = (long)sc.n_frames * slabs_per_frame;
  const int item0 = (int)(total_items * blockIdx.x / gridDim.x), item1 = (int)(total_items * (blockIdx.x + 1) / gridDim.x);
  int f_set = -1;
  for (int item = item0; item < item1; ++item) {
    const int f = item / slabs_per_frame, r0 = (item - f * slabs_per_frame) * rows_per_slab;
    const int r1 = min(r0 + rows_per_slab, H);
    if (f != f_set) {
      __syncthreads();  // the previous frame's readers of s_set are done
      if (threadIdx.x == 0) mc_setup_frame(&s_set, sc, f, H, W, depth_scale, voxel, div_mode, layout_ok, dn.ok ? dn.base : nullptr);
      __syncthreads();
      f_set = f;
    }
    if (dn.ok && s_set.fast) continue;
    uint32_t* bits = sc.bitsets + (size_t)f * words;
    if (s_set.fast) {
      if constexpr (DT == V3D_DEPTH_U16) {
        const int rw0 = r0 + warp * rows_per_warp;
        mc_rows_fast(&s_set, scenes + blockIdx.y, f, H, W, rw0, min(rw0 + rows_per_warp, r1), s_q + warp * kWarpQueue * 3, s_wqn + warp);
      }
    } else {
      mc_slab_generic<DT>(sc, &s_set, f, H, W, r0, r1, reinterpret_cast<int3*>(s_q), s_qn, phase, tab, mask, bits);
    }
  }
}

__global__ void __launch_bounds__(kMcThreads) k_mc_voxel_ids(const float* __restrict__ world, size_t n, float voxel, int div_mode,
                                                             int32_t* __restrict__ ids) {
  const float rcp = 1.0f / voxel;
  const Divisor dv = make_divisor(voxel);
  for (size_t i = (size_t)blockIdx.x * kMcThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kMcThreads)
    ids[i] = voxel_id(world[i], dv, rcp, div_mode);
}

// ---- stage 2: greedy set cover -------------------------------------------------------------------
// CPython's MT19937 (Modules/_randommodule.c) driven by one thread; state in shared memory.
struct Mt {
  uint32_t mt[624];
  int idx;
};
// One thread; the recurrence is serial, so the previous word stays in a register (a reload of the word just stored
// would put a shared-memory round trip on every one of the 1871 steps).
__device__ void mt_seed(Mt* s, uint64_t seed) {  // random.seed(int): init_by_array(32-bit words of |seed|)
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const int klen = key[1] ? 2 : 1;
  uint32_t* mt = s->mt;
  uint32_t p = 19650218u;
  mt[0] = p;
  for (int i = 1; i < 624; ++i) {
    p = 1812433253u * (p ^ (p >> 30)) + (uint32_t)i;
    mt[i] = p;
  }
  int i = 1, j = 0;
  p = 19650218u;  // mt[i - 1] for i = 1
  for (int k = 624; k; --k) {
    p = (mt[i] ^ ((p ^ (p >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    mt[i] = p;
    ++i;
    ++j;
    if (i >= 624) { mt[0] = p; i = 1; }
    if (j >= klen) j = 0;
  }
  for (int k = 623; k; --k) {
    p = (mt[i] ^ ((p ^ (p >> 30)) * 1566083941u)) - (uint32_t)i;
    mt[i] = p;
    ++i;
    if (i >= 624) { mt[0] = p; i = 1; }
  }
  mt[0] = 0x80000000u;
  s->idx = 624;
}
// The first regeneration of the state after seeding, by the whole CTA (mt_next does it with one thread: 624 steps of three
// shared-memory loads each).  Words [0,227) depend on old words only, [227,454) on the new [0,227), [454,623) on the new
// [227,396), the last word on the new words 396 and 0; inside a phase every thread reads before anybody writes.
__device__ void mt_first_twist(Mt* s, int t, int n_threads) {
  uint32_t* mt = s->mt;
  auto mix = [](uint32_t hi, uint32_t lo) {
    const uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  };
  const int lo[3] = {0, 227, 454}, hi[3] = {227, 454, 623};
  for (int ph = 0; ph < 3; ++ph) {
    for (int k0 = lo[ph]; k0 < hi[ph]; k0 += n_threads) {  // one pass for 512 threads
      const int kk = k0 + t;
      uint32_t v = 0;
      if (kk < hi[ph]) v = mt[kk < 227 ? kk + 397 : kk - 227] ^ mix(mt[kk], mt[kk + 1]);
      __syncthreads();
      if (kk < hi[ph]) mt[kk] = v;
      __syncthreads();
    }
  }
  if (t == 0) {
    mt[623] = mt[396] ^ mix(mt[623], mt[0]);
    s->idx = 0;
  }
  __syncthreads();
}
__device__ uint32_t mt_next(Mt* s) {
  uint32_t* mt = s->mt;
  if (s->idx >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; ++kk) {
      const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < 623; ++kk) {
      const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s->idx = 0;
  }
  uint32_t y = mt[s->idx++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
// Random._randbelow_with_getrandbits: k = n.bit_length(); r = getrandbits(k); while r >= n: retry
__device__ uint32_t mt_randbelow(Mt* s, uint32_t n) {
  const int k = 32 - __clz(n);
  uint32_t r = mt_next(s) >> (32 - k);
  while (r >= n) r = mt_next(s) >> (32 - k);
  return r;
}

// One cluster of kGreedyCluster CTAs per scene.  CTA `rank` owns a contiguous block of frames whose coverage
// bitsets it keeps in shared memory for all rounds (when they fit); every CTA keeps its own copy of `used` and of
// the MT19937 state, so a round needs one exchange only: the owners' gains, read through distributed shared
// memory after one cluster barrier (double-buffered by round parity), after which every CTA derives the same
// pick (max gain, ascending tie list, random.choice).
__global__ void __launch_bounds__(kGreedyThreads) k_mc_greedy(const v3d_mc_scene* __restrict__ scenes, int max_pick, int per_max,
                                                              int bits_in_smem) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const v3d_mc_scene sc = scenes[blockIdx.x / kGreedyCluster];
  const int F = sc.n_frames, words = (sc.n_pc + 31) / 32;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, n_warps = kGreedyThreads / 32;
  const int per = max((F + kGreedyCluster - 1) / kGreedyCluster, 1);
  const int f_lo = min(F, rank * per), nf = min(F, f_lo + per) - f_lo;
  // layout (identical in every CTA of the cluster: remote addresses are computed from the local ones)
  int* s_gain_own = reinterpret_cast<int*>(smem_raw);              // [2][per_max], -2 = already taken
  int* s_part = s_gain_own + 2 * per_max;                            // [1] partial popcount of the union
  int* s_taken = s_part + 1;                                         // [per_max]
  int* s_gain = s_taken + per_max;                                   // [F] gathered gains
  uint32_t* s_used = reinterpret_cast<uint32_t*>(s_gain + kGreedyCluster * per_max);
  Mt* s_mt = reinterpret_cast<Mt*>(s_used + words);
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_mt + 1);          // [nf][words] when bits_in_smem
  __shared__ int s_red[kGreedyThreads / 32];
  __shared__ int s_pick, s_max;

  for (int w = t; w < words; w += kGreedyThreads) s_used[w] = 0u;
  for (int f = t; f < per_max; f += kGreedyThreads) s_taken[f] = 0;
  if (rank == 0)
    for (int i = t; i < V3D_MC_RECORD_INTS; i += kGreedyThreads) sc.record[i] = (i < 2 * V3D_MC_MAX_PICK) ? -1 : 0;
  if (t == 0) mt_seed(s_mt, sc.seed);  // overlaps with the loads below on the other warps
  const uint32_t* g_bits = sc.bitsets + (size_t)f_lo * words;
  const uint32_t* my_bits = g_bits;
  if (bits_in_smem) {
    for (int i = t; i < nf * words; i += kGreedyThreads) s_bits[i] = g_bits[i];
    my_bits = s_bits;
  }
  // n_all = |(union of frames) & pc| (:102): this CTA counts its slice of the words of the OR over all frames
  {
    const int wper = (words + kGreedyCluster - 1) / kGreedyCluster;
    const int w_hi = min(words, (rank + 1) * wper);
    int cnt = 0;
    for (int w = rank * wper + t; w < w_hi; w += kGreedyThreads) {
      uint32_t u = 0u;
      for (int f = 0; f < F; ++f) u |= sc.bitsets[(size_t)f * words + w];
      cnt += __popc(u);
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_red[warp] = cnt;
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int i = 0; i < n_warps; ++i) tot += s_red[i];
      *s_part = tot;
    }
  }
  __syncthreads();
  mt_first_twist(s_mt, t, kGreedyThreads);

  int n_picked = 0;
  const int rounds = min(F, max_pick);
  for (int round = 0; round < rounds; ++round) {
    int* own = s_gain_own + (round & 1) * per_max;
    // gains of this CTA's remaining frames: one warp per frame, lanes stride the words
    for (int fl = warp; fl < nf; fl += n_warps) {
      int g = -2;
      if (!s_taken[fl]) {
        const uint32_t* bs = my_bits + (size_t)fl * words;
        g = 0;
        for (int w = lane; w < words; w += 32) g += __popc(bs[w] & ~s_used[w]);
        for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      }
      if (lane == 0) own[fl] = g;
    }
    cluster.sync();
    for (int f = t; f < F; f += kGreedyThreads) {
      const int c = f / per;
      s_gain[f] = cluster.map_shared_rank(own, c)[f - c * per];
    }
    __syncthreads();
    if (warp == 0) {
      // max gain, then the r-th frame (ascending) among those that reach it
      int m = -1;
      for (int f = lane; f < F; f += 32) m = max(m, s_gain[f]);
      for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
      int n_ties = 0;
      for (int f0 = 0; f0 < F; f0 += 32) {
        const int f = f0 + lane;
        n_ties += __popc(__ballot_sync(0xffffffffu, f < F && s_gain[f] == m));
      }
      uint32_t r = 0;
      if (lane == 0) r = mt_randbelow(s_mt, (uint32_t)n_ties);
      r = __shfl_sync(0xffffffffu, r, 0);
      int pick = -1, seen = 0;
      for (int f0 = 0; f0 < F && pick < 0; f0 += 32) {
        const int f = f0 + lane;
        const uint32_t b = __ballot_sync(0xffffffffu, f < F && s_gain[f] == m);
        const int c = __popc(b);
        if ((int)r < seen + c) {
          uint32_t bb = b;
          for (int k = (int)r - seen; k > 0; --k) bb &= bb - 1;  // drop the k lowest set bits
          pick = f0 + __ffs(bb) - 1;
        }
        seen += c;
      }
      if (lane == 0) {
        s_pick = pick;
        s_max = m;
      }
    }
    __syncthreads();
    const int pick = s_pick;
    const uint32_t* bs = (bits_in_smem && pick >= f_lo && pick < f_lo + nf) ? s_bits + (size_t)(pick - f_lo) * words
                                                                            : sc.bitsets + (size_t)pick * words;
    for (int w = t; w < words; w += kGreedyThreads) s_used[w] |= bs[w];
    if (t == 0) {
      if (pick >= f_lo && pick < f_lo + nf) s_taken[pick - f_lo] = 1;
      if (rank == 0) {
        sc.record[n_picked] = pick;
        sc.record[V3D_MC_MAX_PICK + n_picked] = s_max;
      }
    }
    ++n_picked;
    __syncthreads();
  }
  cluster.sync();  // every CTA's s_part is written; nobody reads remote gains any more
  if (rank == 0) {
    // n_sel = |used & pc| (:103)
    int cnt = 0;
    for (int w = t; w < words; w += kGreedyThreads) cnt += __popc(s_used[w]);
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_red[warp] = cnt;
    __syncthreads();
    if (t == 0) {
      int tot = 0, n_all = 0;
      for (int i = 0; i < n_warps; ++i) tot += s_red[i];
      for (int c = 0; c < kGreedyCluster; ++c) n_all += *cluster.map_shared_rank(s_part, c);
      sc.record[2 * V3D_MC_MAX_PICK + 0] = n_picked;
      sc.record[2 * V3D_MC_MAX_PICK + 1] = n_all;
      sc.record[2 * V3D_MC_MAX_PICK + 2] = tot;
    }
  }
  cluster.sync();  // shared memory of a CTA must outlive the remote reads of it
}

// ---- host side -------------------------------------------------------------------------------------
static int stage_scenes(const v3d_mc_scene* scenes_host, int n_scenes, void* scratch, cudaStream_t s, int* max_frames,
                        int* max_pc, long* total_frames = nullptr) {
  V3D_REQUIRE(scenes_host && scratch, V3D_E_ARG, "mc: null scene array / scratch");
  V3D_REQUIRE(n_scenes > 0 && n_scenes <= 65535, V3D_E_ARG, "mc: n_scenes=%d out of range", n_scenes);
  *max_frames = 0;
  *max_pc = 0;
  if (total_frames) *total_frames = 0;
  for (int i = 0; i < n_scenes; ++i) {
    const v3d_mc_scene& sc = scenes_host[i];
    V3D_REQUIRE(sc.depth && sc.poses && sc.intr && sc.bitsets && sc.hash_table && sc.record, V3D_E_ARG, "mc: scene %d has a null pointer", i);
    V3D_REQUIRE(sc.n_pc == 0 || sc.pc_voxels, V3D_E_ARG, "mc: scene %d has no pc_voxels", i);
    V3D_REQUIRE(sc.n_frames >= 0 && sc.n_pc >= 0, V3D_E_ARG, "mc: scene %d has negative sizes", i);
    V3D_REQUIRE(sc.hash_slots >= v3d_mc_hash_slots(sc.n_pc) && (sc.hash_slots & (sc.hash_slots - 1)) == 0, V3D_E_ARG,
                "mc: scene %d hash_slots=%d must be a power of two >= %d", i, sc.hash_slots, v3d_mc_hash_slots(sc.n_pc));
    V3D_REQUIRE((reinterpret_cast<uintptr_t>(sc.hash_table) & 15) == 0, V3D_E_ALIGN, "mc: scene %d hash_table not 16-byte aligned", i);
    if (total_frames) *total_frames += sc.n_frames;
    if (sc.n_frames > *max_frames) *max_frames = sc.n_frames;
    if (sc.n_pc > *max_pc) *max_pc = sc.n_pc;
  }
  V3D_CUDA(cudaMemcpyAsync(scratch, scenes_host, sizeof(v3d_mc_scene) * (size_t)n_scenes, cudaMemcpyHostToDevice, s));
  return V3D_OK;
}

static int launch_voxelize(const v3d_mc_scene* dev_scenes, int n_scenes, int max_frames, long total_frames, int max_pc, int H, int W,
                           int depth_dtype, float depth_scale, float voxel, int div_mode, cudaStream_t s) {
  V3D_REQUIRE(H > 0 && W > 0 && voxel > 0.f, V3D_E_ARG, "mc: bad H/W/voxel");
  if (max_frames == 0) return V3D_OK;
  // clear + build
  {
    const size_t per_scene = (size_t)max_frames * ((max_pc + 31) / 32) + (size_t)v3d_mc_hash_slots(max_pc);
    int gx = (int)((per_scene + kMcThreads * 4 - 1) / (kMcThreads * 4));
    gx = gx < 1 ? 1 : (gx > 1024 ? 1024 : gx);
    k_mc_clear<<<dim3(gx, n_scenes), kMcThreads, 0, s>>>(dev_scenes);
    int bx = (max_pc + kMcThreads - 1) / kMcThreads;
    bx = bx < 1 ? 1 : (bx > 1024 ? 1024 : bx);
    k_mc_build<<<dim3(bx, n_scenes), kMcThreads, 0, s>>>(dev_scenes);
  }
  // dense voxeliser first (uint16 depth in whole 16-pixel chunks; anything it leaves is done by k_mc_voxelize below)
  int dense_bits = 0;
  const char* dense_env = getenv("V3D_MC_DENSE");  // "0": hash kernel only (tests compare the two, tools/time_mc.py times both)
  if (!(dense_env && dense_env[0] == '0') && depth_dtype == V3D_DEPTH_U16 && (W % kDensePx) == 0 && (long)H * (W / kDensePx) * (W / kDensePx) < (1L << 32) && total_frames > 0 &&
      n_scenes <= 4096 /* a CTA finds its first scene by walking the descriptors */) {
    // per device and cheap: set at every launch (a process may drive several devices)
    V3D_CUDA(cudaFuncSetAttribute(k_mc_voxelize_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmemBytes));
    dense_bits = kDenseSmemBytes * 8;
    const int rows_per_item = (kDenseThreads / 32) * kDenseRows;
    const int items_per_frame = (H + rows_per_item - 1) / rows_per_item;
    const long total_items = total_frames * items_per_frame;
    long g = (long)num_sms() * V3D_DENSE_BLOCKS;
    if (g > total_items) g = total_items;
    const uint32_t gpr_magic = 0xffffffffu / (uint32_t)(W / kDensePx) + 1u;  // g / gpr == umulhi(g, magic) while g * gpr < 2^32
    k_mc_voxelize_dense<<<(unsigned)g, kDenseThreads, kDenseSmemBytes, s>>>(dev_scenes, n_scenes, H, W, depth_scale, voxel, div_mode, items_per_frame,
                                                                           rows_per_item, total_items, dense_bits, gpr_magic);
    int rc = check_launch("k_mc_voxelize_dense");
    if (rc != V3D_OK) return rc;
  }
  // 48 rows = 6 per warp; 6 x (640/8) groups = 15 full warp iterations for ScanNet-shaped depth
  const int rows_per_slab = 48;
  const int slabs = (H + rows_per_slab - 1) / rows_per_slab;
  // enough CTAs per scene to fill the machine a few times over, grid.y = scene
  long want = ((long)num_sms() * 16 + n_scenes - 1) / n_scenes;
  long items = (long)max_frames * slabs;
  int gx = (int)(want < items ? want : items);
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_scenes);
  switch (depth_dtype) {
    case V3D_DEPTH_U16: k_mc_voxelize<V3D_DEPTH_U16><<<grid, kMcThreads, 0, s>>>(dev_scenes, H, W, depth_scale, voxel, div_mode, slabs, rows_per_slab, dense_bits); break;
    case V3D_DEPTH_I32: k_mc_voxelize<V3D_DEPTH_I32><<<grid, kMcThreads, 0, s>>>(dev_scenes, H, W, depth_scale, voxel, div_mode, slabs, rows_per_slab, dense_bits); break;
    case V3D_DEPTH_F32: k_mc_voxelize<V3D_DEPTH_F32><<<grid, kMcThreads, 0, s>>>(dev_scenes, H, W, depth_scale, voxel, div_mode, slabs, rows_per_slab, dense_bits); break;
    default: set_error("mc: unknown depth dtype %d", depth_dtype); return V3D_E_ARG;
  }
  return check_launch("k_mc_voxelize");
}

static int launch_greedy(const v3d_mc_scene* dev_scenes, int n_scenes, int max_frames, int max_pc, int max_pick, cudaStream_t s) {
  V3D_REQUIRE(max_pick > 0 && max_pick <= V3D_MC_MAX_PICK, V3D_E_ARG, "mc: max_pick=%d must be in [1,%d]", max_pick, V3D_MC_MAX_PICK);
  const int words = (max_pc + 31) / 32;
  V3D_REQUIRE(words <= kMaxWordsSmem, V3D_E_UNSUPPORTED, "mc: universe of %d voxels exceeds the %d supported", max_pc, kMaxWordsSmem * 32);
  int per_max = (max_frames + kGreedyCluster - 1) / kGreedyCluster;
  if (per_max < 1) per_max = 1;
  const size_t fixed = (size_t)(3 * per_max + 1 + kGreedyCluster * per_max) * 4 + (size_t)words * 4 + sizeof(Mt) + 16;
  V3D_REQUIRE(fixed <= 200 * 1024, V3D_E_UNSUPPORTED, "mc: greedy needs %zu bytes of shared memory", fixed);
  const size_t with_bits = fixed + (size_t)per_max * words * 4;
  const int bits_in_smem = with_bits <= 200 * 1024;
  const size_t smem = bits_in_smem ? with_bits : fixed;
  V3D_CUDA(cudaFuncSetAttribute(k_mc_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_scenes * kGreedyCluster);
  cfg.blockDim = dim3(kGreedyThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kGreedyCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  V3D_CUDA(cudaLaunchKernelEx(&cfg, k_mc_greedy, dev_scenes, max_pick, per_max, bits_in_smem));
  return check_launch("k_mc_greedy");
}

}  // namespace v3d

using namespace v3d;

extern "C" {

int v3d_mc_hash_slots(int n_pc) {
  int s = 64;
  while (s < 4 * (long long)n_pc && s < (1 << 30)) s <<= 1;  // load factor <= 0.25: probe chains of 1-2 slots
  return s;
}

int v3d_mc_voxelize(const v3d_mc_scene* scenes_host, int n_scenes, int H, int W, int depth_dtype, float depth_scale,
                    float voxel_size, int div_mode, void* scratch, v3d_stream_t stream) {
  int mf, mp;
  long tf;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = stage_scenes(scenes_host, n_scenes, scratch, s, &mf, &mp, &tf);
  if (rc != V3D_OK) return rc;
  return launch_voxelize(reinterpret_cast<const v3d_mc_scene*>(scratch), n_scenes, mf, tf, mp, H, W, depth_dtype, depth_scale, voxel_size, div_mode, s);
}

int v3d_mc_greedy(const v3d_mc_scene* scenes_host, int n_scenes, int max_pick, void* scratch, v3d_stream_t stream) {
  int mf, mp;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = stage_scenes(scenes_host, n_scenes, scratch, s, &mf, &mp);
  if (rc != V3D_OK) return rc;
  return launch_greedy(reinterpret_cast<const v3d_mc_scene*>(scratch), n_scenes, mf, mp, max_pick, s);
}

int v3d_mc_sample(const v3d_mc_scene* scenes_host, int n_scenes, int H, int W, int depth_dtype, float depth_scale,
                  float voxel_size, int div_mode, int max_pick, void* scratch, v3d_stream_t stream) {
  int mf, mp;
  long tf;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = stage_scenes(scenes_host, n_scenes, scratch, s, &mf, &mp, &tf);
  if (rc != V3D_OK) return rc;
  const v3d_mc_scene* dev = reinterpret_cast<const v3d_mc_scene*>(scratch);
  rc = launch_voxelize(dev, n_scenes, mf, tf, mp, H, W, depth_dtype, depth_scale, voxel_size, div_mode, s);
  if (rc != V3D_OK) return rc;
  return launch_greedy(dev, n_scenes, mf, mp, max_pick, s);
}

int v3d_mc_voxel_ids(const float* world, size_t n_points, float voxel_size, int div_mode, int32_t* ids, v3d_stream_t stream) {
  V3D_REQUIRE(world && ids, V3D_E_ARG, "v3d_mc_voxel_ids: null pointer");
  V3D_REQUIRE(voxel_size > 0.f, V3D_E_ARG, "v3d_mc_voxel_ids: voxel_size must be positive");
  if (n_points == 0) return V3D_OK;
  const size_t n = n_points * 3;
  size_t blocks = (n + kMcThreads - 1) / kMcThreads;
  const size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  k_mc_voxel_ids<<<(int)blocks, kMcThreads, 0, (cudaStream_t)stream>>>(world, n, voxel_size, div_mode, ids);
  return check_launch("k_mc_voxel_ids");
}

}  // extern "C"
