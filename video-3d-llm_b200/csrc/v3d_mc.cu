// v3d_mc.cu -- maximum-coverage frame sampler on sm_100a (C ABI: include/v3d.h).
//
// Replaces scripts/3d/preprocessing/max_coverage_sampling.py:39-104:
//   k_mc_build    : open-addressing hash of the mesh universe pc_voxel (:65), clears bitsets
//   k_mc_voxelize : unproject every pixel of every candidate frame, vox = int(round(w/voxel))
//                   (:44-45), look it up, atomicOr the frame's coverage bit -> per-frame
//                   bitsets = set(frame voxels) & pc_voxel (:47-50, :75)
//   k_mc_greedy   : one CTA per scene; each round = popcount(bits & ~used) per remaining frame,
//                   argmax with the ascending-order tie list (:72-82), CPython-compatible
//                   MT19937 random.choice (:84), used |= bits (:86) -- bit-exact selection.
#include <math_constants.h>

#include "v3d_common.cuh"

namespace v3d {

static constexpr int kMcThreads = 256;
static constexpr int kGreedyThreads = 512;
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
}

__global__ void __launch_bounds__(kMcThreads) k_mc_build(const v3d_mc_scene* __restrict__ scenes) {
  const v3d_mc_scene sc = scenes[blockIdx.y];
  const uint32_t mask = (uint32_t)sc.hash_slots - 1u;
  for (int b = blockIdx.x * kMcThreads + threadIdx.x; b < sc.n_pc; b += gridDim.x * kMcThreads) {
    const int i = sc.pc_voxels[3 * b], j = sc.pc_voxels[3 * b + 1], k = sc.pc_voxels[3 * b + 2];
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
      if (sc.pc_voxels[3 * prev] == i && sc.pc_voxels[3 * prev + 1] == j && sc.pc_voxels[3 * prev + 2] == k) break;
      h = (h + 1) & mask;
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

// rint(RN(w / voxel)) without the division when it provably does not matter: t = w * r approximates the
// IEEE quotient q within |t| * 2^-22 (r is within an ulp of 1/voxel, one more rounding for the product),
// so if t is farther than that (with a 2x safety factor) from every half-integer, rint(q) == rint(t).
// About one coordinate in 10^4 lands in the guard band and takes the exact division.
__device__ __forceinline__ int voxel_id_fast(float w, const Divisor& voxel) {
  const float t = __fmul_rn(w, voxel.r);
  const float n = rintf(t);
  const float margin = __fmaf_rn(fabsf(t), -0x1p-21f, 0.5f);
  if (fabsf(__fsub_rn(t, n)) < margin && fabsf(t) < 1.0e9f) return (int)n;
  return voxel_id(w, voxel, 0.f, V3D_DIV_IEEE);
}

struct Vox {
  int i, j, k;
};
__device__ __forceinline__ bool vox_eq(const Vox& a, const Vox& b) { return a.i == b.i && a.j == b.j && a.k == b.k; }

// Candidate voxels of one block iteration are queued in shared memory so that the hash probes run with
// dense lanes (a probe is a dependent chain of L2 loads; run inline, 2-4 of 32 lanes would be active).
__device__ __forceinline__ void mc_push(int3* s_q, int* s_qn, const Vox& v) {
  const int pos = atomicAdd(s_qn, 1);
  s_q[pos] = make_int3(v.i, v.j, v.k);
}

template <int DT>
__global__ void __launch_bounds__(kMcThreads) k_mc_voxelize(const v3d_mc_scene* __restrict__ scenes, int H, int W,
                                                            float depth_scale, float voxel, int div_mode, int slabs_per_frame,
                                                            int rows_per_slab) {
  const v3d_mc_scene sc = scenes[blockIdx.y];
  const int words = (sc.n_pc + 31) / 32;
  const uint32_t mask = (uint32_t)sc.hash_slots - 1u;
  const int4* tab = reinterpret_cast<const int4*>(sc.hash_table);
  const float rcp = 1.0f / voxel;
  const Divisor ds = make_divisor(depth_scale);
  const Divisor dv = make_divisor(voxel);
  const int groups_per_row = (W + kPxPerThread - 1) / kPxPerThread;
  const int lane = threadIdx.x & 31;
  __shared__ int3 s_q[kMcThreads * kPxPerThread];  // 24 KB
  __shared__ int s_qn;
  if (threadIdx.x == 0) s_qn = 0;
  __syncthreads();
  for (int item = blockIdx.x; item < sc.n_frames * slabs_per_frame; item += gridDim.x) {
    const int f = item / slabs_per_frame, r0 = (item - f * slabs_per_frame) * rows_per_slab;
    const int r1 = min(r0 + rows_per_slab, H);
    const Camera cam = load_camera(sc.poses, sc.intr, f);
    // uint16 depth + rigid pose + divisors well inside the normal range: branch-free unprojection (see
    // pool_rows_fast in v3d_frontend.cu) and division-free voxel ids
    const bool frame_fast = DT == V3D_DEPTH_U16 && div_mode == V3D_DIV_IEEE && cam.rigid && ds.ok && cam.fx.ok && cam.fy.ok &&
                            ds.b > 0.f && dv.ok;
    uint32_t* bits = sc.bitsets + (size_t)f * words;
    const int n_groups = (r1 - r0) * groups_per_row;
    // all threads of the block iterate together (barriers and shuffles below); whole warps past the end idle
    for (int g0 = 0; g0 < n_groups; g0 += kMcThreads) {
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
        bool fast = false;
        if constexpr (DT == V3D_DEPTH_U16) {
          const float um0 = __fsub_rn((float)c0, cam.cx), um7 = __fsub_rn((float)(c0 + kPxPerThread - 1), cam.cx);
          // (u - cx) of the 8 columns lies between um0 and um7: all are in the fast range unless that
          // interval comes within 2^-20 of zero
          fast = frame_fast && n_px == kPxPerThread && ((base * 2 + reinterpret_cast<uintptr_t>(sc.depth)) & 15) == 0 &&
                 fabsf(vm) >= 0x1p-20f && fabsf(vm) <= 0x1p20f && fabsf(um0) <= 0x1p20f && fabsf(um7) <= 0x1p20f &&
                 (um0 > 0x1p-20f || um7 < -0x1p-20f);
        }
        if (fast) {
          const uint4 q = ld_stream_v4(reinterpret_cast<const uint16_t*>(sc.depth) + base);
          const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < kPxPerThread; ++e) {
            const uint32_t raw = (e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xffffu);
            const float df = (float)raw;
            const bool nz = raw != 0;
            const float um = __fsub_rn((float)(c0 + e), cam.cx);
            float q0 = __fmul_rn(df, ds.r);
            float q1 = __fmaf_rn(ds.r, __fmaf_rn(-ds.b, q0, df), q0);
            const float cz = nz ? q1 : q0;
            const float ax = __fmul_rn(um, cz);
            q0 = __fmul_rn(ax, cam.fx.r);
            q1 = __fmaf_rn(cam.fx.r, __fmaf_rn(-cam.fx.b, q0, ax), q0);
            const float cx_ = nz ? q1 : q0;
            const float ay = __fmul_rn(vm, cz);
            q0 = __fmul_rn(ay, cam.fy.r);
            q1 = __fmaf_rn(cam.fy.r, __fmaf_rn(-cam.fy.b, q0, ay), q0);
            const float cy_ = nz ? q1 : q0;
            const float x = __fadd_rn(__fmaf_rn(cam.P[2], cz, __fmaf_rn(cam.P[1], cy_, __fmul_rn(cam.P[0], cx_))), cam.P[3]);
            const float y = __fadd_rn(__fmaf_rn(cam.P[6], cz, __fmaf_rn(cam.P[5], cy_, __fmul_rn(cam.P[4], cx_))), cam.P[7]);
            const float z = __fadd_rn(__fmaf_rn(cam.P[10], cz, __fmaf_rn(cam.P[9], cy_, __fmul_rn(cam.P[8], cx_))), cam.P[11]);
            const Vox v = {voxel_id_fast(x, dv), voxel_id_fast(y, dv), voxel_id_fast(z, dv)};
            if (e == 0) first = v;
            else if (!vox_eq(v, prev)) mc_push(s_q, &s_qn, v);
            prev = v;
          }
        } else {
#pragma unroll 1
          for (int e = 0; e < n_px; ++e) {
            float x, y, z;
            unproject_px(cam, load_depth<DT>(sc.depth, base + e), __fsub_rn((float)(c0 + e), cam.cx), vm, ds, x, y, z);
            const Vox v = {voxel_id(x, dv, rcp, div_mode), voxel_id(y, dv, rcp, div_mode), voxel_id(z, dv, rcp, div_mode)};
            if (e == 0) first = v;
            else if (!vox_eq(v, prev)) mc_push(s_q, &s_qn, v);
            prev = v;
          }
        }
      }
      // the first pixel of a group is a duplicate if it equals the last pixel of the left neighbour lane
      Vox left;
      left.i = __shfl_up_sync(0xffffffffu, prev.i, 1);
      left.j = __shfl_up_sync(0xffffffffu, prev.j, 1);
      left.k = __shfl_up_sync(0xffffffffu, prev.k, 1);
      const int left_px = __shfl_up_sync(0xffffffffu, n_px, 1);
      if (n_px > 0 && (lane == 0 || left_px != kPxPerThread || c0 == 0 || !vox_eq(first, left))) mc_push(s_q, &s_qn, first);
      __syncthreads();
      const int qn = s_qn;
      for (int i = threadIdx.x; i < qn; i += kMcThreads) {
        const int3 v = s_q[i];
        const int b = hash_find(tab, mask, v.x, v.y, v.z);
        if (b >= 0) {
          const uint32_t m = 1u << (b & 31);
          if (!(bits[b >> 5] & m)) atomicOr(bits + (b >> 5), m);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) s_qn = 0;
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
__device__ void mt_seed(Mt* s, uint64_t seed) {  // random.seed(int): init_by_array(32-bit words of |seed|)
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const int klen = key[1] ? 2 : 1;
  uint32_t* mt = s->mt;
  mt[0] = 19650218u;
  for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  int i = 1, j = 0;
  for (int k = 624; k; --k) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    ++i;
    ++j;
    if (i >= 624) { mt[0] = mt[623]; i = 1; }
    if (j >= klen) j = 0;
  }
  for (int k = 623; k; --k) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    ++i;
    if (i >= 624) { mt[0] = mt[623]; i = 1; }
  }
  mt[0] = 0x80000000u;
  s->idx = 624;
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

__global__ void __launch_bounds__(kGreedyThreads) k_mc_greedy(const v3d_mc_scene* __restrict__ scenes, int max_pick) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const v3d_mc_scene sc = scenes[blockIdx.x];
  const int F = sc.n_frames, words = (sc.n_pc + 31) / 32;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, n_warps = kGreedyThreads / 32;
  uint32_t* s_used = reinterpret_cast<uint32_t*>(smem_raw);
  int* s_gain = reinterpret_cast<int*>(s_used + words);  // -2 = already taken
  Mt* s_mt = reinterpret_cast<Mt*>(s_gain + F);
  __shared__ int s_red[kGreedyThreads / 32];
  __shared__ int s_pick, s_max, s_n_all;

  for (int w = t; w < words; w += kGreedyThreads) s_used[w] = 0u;
  for (int f = t; f < F; f += kGreedyThreads) s_gain[f] = 0;
  for (int i = t; i < V3D_MC_RECORD_INTS; i += kGreedyThreads) sc.record[i] = (i < 2 * V3D_MC_MAX_PICK) ? -1 : 0;
  if (t == 0) mt_seed(s_mt, sc.seed);  // overlaps with the union popcount below on the other warps

  // n_all = |(union of frames) & pc| (:102): popcount of the OR over frames
  int cnt = 0;
  if (t >= 32 || F == 0) {
    for (int w = t - (F ? 32 : 0); w < words; w += kGreedyThreads - (F ? 32 : 0)) {
      uint32_t u = 0u;
      for (int f = 0; f < F; ++f) u |= sc.bitsets[(size_t)f * words + w];
      cnt += __popc(u);
    }
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_red[warp] = cnt;
  __syncthreads();
  if (t == 0) {
    int tot = 0;
    for (int i = 0; i < n_warps; ++i) tot += s_red[i];
    s_n_all = tot;
  }
  __syncthreads();

  int n_picked = 0;
  const int rounds = min(F, max_pick);
  for (int round = 0; round < rounds; ++round) {
    // gains of the remaining frames: one warp per frame, lanes stride the words
    for (int f = warp; f < F; f += n_warps) {
      if (s_gain[f] == -2) continue;
      const uint32_t* bs = sc.bitsets + (size_t)f * words;
      int g = 0;
      for (int w = lane; w < words; w += 32) g += __popc(bs[w] & ~s_used[w]);
      for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      if (lane == 0) s_gain[f] = g;
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
    const uint32_t* bs = sc.bitsets + (size_t)pick * words;
    for (int w = t; w < words; w += kGreedyThreads) s_used[w] |= bs[w];
    if (t == 0) {
      sc.record[n_picked] = pick;
      sc.record[V3D_MC_MAX_PICK + n_picked] = s_max;
      s_gain[pick] = -2;
    }
    ++n_picked;
    __syncthreads();
  }
  // n_sel = |used & pc| (:103)
  cnt = 0;
  for (int w = t; w < words; w += kGreedyThreads) cnt += __popc(s_used[w]);
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = cnt;
  __syncthreads();
  if (t == 0) {
    int tot = 0;
    for (int i = 0; i < n_warps; ++i) tot += s_red[i];
    sc.record[2 * V3D_MC_MAX_PICK + 0] = n_picked;
    sc.record[2 * V3D_MC_MAX_PICK + 1] = s_n_all;
    sc.record[2 * V3D_MC_MAX_PICK + 2] = tot;
  }
}

// ---- host side -------------------------------------------------------------------------------------
static int stage_scenes(const v3d_mc_scene* scenes_host, int n_scenes, void* scratch, cudaStream_t s, int* max_frames,
                        int* max_pc) {
  V3D_REQUIRE(scenes_host && scratch, V3D_E_ARG, "mc: null scene array / scratch");
  V3D_REQUIRE(n_scenes > 0 && n_scenes <= 65535, V3D_E_ARG, "mc: n_scenes=%d out of range", n_scenes);
  *max_frames = 0;
  *max_pc = 0;
  for (int i = 0; i < n_scenes; ++i) {
    const v3d_mc_scene& sc = scenes_host[i];
    V3D_REQUIRE(sc.depth && sc.poses && sc.intr && sc.bitsets && sc.hash_table && sc.record, V3D_E_ARG, "mc: scene %d has a null pointer", i);
    V3D_REQUIRE(sc.n_pc == 0 || sc.pc_voxels, V3D_E_ARG, "mc: scene %d has no pc_voxels", i);
    V3D_REQUIRE(sc.n_frames >= 0 && sc.n_pc >= 0, V3D_E_ARG, "mc: scene %d has negative sizes", i);
    V3D_REQUIRE(sc.hash_slots >= v3d_mc_hash_slots(sc.n_pc) && (sc.hash_slots & (sc.hash_slots - 1)) == 0, V3D_E_ARG,
                "mc: scene %d hash_slots=%d must be a power of two >= %d", i, sc.hash_slots, v3d_mc_hash_slots(sc.n_pc));
    V3D_REQUIRE((reinterpret_cast<uintptr_t>(sc.hash_table) & 15) == 0, V3D_E_ALIGN, "mc: scene %d hash_table not 16-byte aligned", i);
    if (sc.n_frames > *max_frames) *max_frames = sc.n_frames;
    if (sc.n_pc > *max_pc) *max_pc = sc.n_pc;
  }
  V3D_CUDA(cudaMemcpyAsync(scratch, scenes_host, sizeof(v3d_mc_scene) * (size_t)n_scenes, cudaMemcpyHostToDevice, s));
  return V3D_OK;
}

static int launch_voxelize(const v3d_mc_scene* dev_scenes, int n_scenes, int max_frames, int max_pc, int H, int W, int depth_dtype,
                           float depth_scale, float voxel, int div_mode, cudaStream_t s) {
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
  // 16 rows x (640/8) groups = 1280 = 5 full block iterations for ScanNet-shaped depth
  const int rows_per_slab = 16;
  const int slabs = (H + rows_per_slab - 1) / rows_per_slab;
  // enough CTAs per scene to fill the machine a few times over, grid.y = scene
  long want = ((long)num_sms() * 16 + n_scenes - 1) / n_scenes;
  long items = (long)max_frames * slabs;
  int gx = (int)(want < items ? want : items);
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_scenes);
  switch (depth_dtype) {
    case V3D_DEPTH_U16: k_mc_voxelize<V3D_DEPTH_U16><<<grid, kMcThreads, 0, s>>>(dev_scenes, H, W, depth_scale, voxel, div_mode, slabs, rows_per_slab); break;
    case V3D_DEPTH_I32: k_mc_voxelize<V3D_DEPTH_I32><<<grid, kMcThreads, 0, s>>>(dev_scenes, H, W, depth_scale, voxel, div_mode, slabs, rows_per_slab); break;
    case V3D_DEPTH_F32: k_mc_voxelize<V3D_DEPTH_F32><<<grid, kMcThreads, 0, s>>>(dev_scenes, H, W, depth_scale, voxel, div_mode, slabs, rows_per_slab); break;
    default: set_error("mc: unknown depth dtype %d", depth_dtype); return V3D_E_ARG;
  }
  return check_launch("k_mc_voxelize");
}

static int launch_greedy(const v3d_mc_scene* dev_scenes, int n_scenes, int max_frames, int max_pc, int max_pick, cudaStream_t s) {
  V3D_REQUIRE(max_pick > 0 && max_pick <= V3D_MC_MAX_PICK, V3D_E_ARG, "mc: max_pick=%d must be in [1,%d]", max_pick, V3D_MC_MAX_PICK);
  const int words = (max_pc + 31) / 32;
  V3D_REQUIRE(words <= kMaxWordsSmem, V3D_E_UNSUPPORTED, "mc: universe of %d voxels exceeds the %d supported", max_pc, kMaxWordsSmem * 32);
  const size_t smem = (size_t)words * 4 + (size_t)max_frames * 4 + sizeof(Mt) + 16;
  V3D_REQUIRE(smem <= 200 * 1024, V3D_E_UNSUPPORTED, "mc: greedy needs %zu bytes of shared memory", smem);
  if (smem > 48 * 1024) V3D_CUDA(cudaFuncSetAttribute(k_mc_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mc_greedy<<<n_scenes, kGreedyThreads, smem, s>>>(dev_scenes, max_pick);
  return check_launch("k_mc_greedy");
}

}  // namespace v3d

using namespace v3d;

extern "C" {

int v3d_mc_hash_slots(int n_pc) {
  int s = 64;
  while (s < 2 * n_pc) s <<= 1;
  return s;
}

int v3d_mc_voxelize(const v3d_mc_scene* scenes_host, int n_scenes, int H, int W, int depth_dtype, float depth_scale,
                    float voxel_size, int div_mode, void* scratch, v3d_stream_t stream) {
  int mf, mp;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = stage_scenes(scenes_host, n_scenes, scratch, s, &mf, &mp);
  if (rc != V3D_OK) return rc;
  return launch_voxelize(reinterpret_cast<const v3d_mc_scene*>(scratch), n_scenes, mf, mp, H, W, depth_dtype, depth_scale, voxel_size, div_mode, s);
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
  cudaStream_t s = (cudaStream_t)stream;
  int rc = stage_scenes(scenes_host, n_scenes, scratch, s, &mf, &mp);
  if (rc != V3D_OK) return rc;
  const v3d_mc_scene* dev = reinterpret_cast<const v3d_mc_scene*>(scratch);
  rc = launch_voxelize(dev, n_scenes, mf, mp, H, W, depth_dtype, depth_scale, voxel_size, div_mode, s);
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
