/*
 * v3d_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the position-aware video front end of
 * zd11024/Video-3D-LLM, written so that every rounding step is explicit.  It is
 * the checker the CUDA path is compared against; only tests/, the smoke() of
 * __graft_entry__ and bench.py's cpu_baseline/reference legs may load it.
 *
 * Parity status: PINNED against outputs of the reference itself (imported by
 * path in the build container, see oracle/gen_golden.py -> tests/golden/).  The
 * reference ships no tests or golden vectors of its own.
 *
 * Build (oracle/Makefile):
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC
 * -ffp-contract=off matters: every fused multiply-add below is an explicit
 * fmaf(), everything else must stay an individually rounded fp32 operation.
 *
 * Reference citations are relative to the upstream repository root.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

enum { V3D_DEPTH_U16 = 0, V3D_DEPTH_I32 = 1, V3D_DEPTH_F32 = 2 };
enum { V3D_DIV_IEEE = 0, V3D_DIV_RECIPROCAL = 1 };

int v3d_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void v3d_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static inline float load_depth(const void* p, int dt, size_t i) {
  switch (dt) {
    case V3D_DEPTH_U16: return (float)((const uint16_t*)p)[i];
    case V3D_DEPTH_I32: return (float)((const int32_t*)p)[i];
    default: return ((const float*)p)[i];
  }
}

/* torch.maximum / torch.minimum propagate NaN (fmaxf does not). */
static inline float t_maximum(float a, float b) {
  if (isnan(a) || isnan(b)) return NAN;
  return a > b ? a : b;
}
static inline float t_minimum(float a, float b) {
  if (isnan(a) || isnan(b)) return NAN;
  return a < b ? a : b;
}

/* One pixel of llava/video_utils.py:38-68 (`unproject`).
 *   z = d / depth_scale                     (:57, IEEE division, scale 1000)
 *   x = ((u - cx) * z) / fx                 (:58, u = column index)
 *   y = ((v - cy) * z) / fy                 (:59, v = row index)
 *   w = P @ [x, y, z, 1]                    (:64, a K=4 sgemm: mul, fma, fma, +d)
 *   out = w[:3] / w[3]                      (:65)
 */
static inline void unproject_px(float d, float u, float v, const float* K, const float* P,
                                float depth_scale, float* out) {
  const float fx = K[0], fy = K[5], cx = K[2], cy = K[6];
  const float z = d / depth_scale;
  const float x = ((u - cx) * z) / fx;
  const float y = ((v - cy) * z) / fy;
  float w[4];
  for (int i = 0; i < 4; ++i) {
    float t = P[4 * i + 0] * x;
    t = fmaf(P[4 * i + 1], y, t);
    t = fmaf(P[4 * i + 2], z, t);
    t = t + P[4 * i + 3];
    w[i] = t;
  }
  out[0] = w[0] / w[3];
  out[1] = w[1] / w[3];
  out[2] = w[2] / w[3];
}

/* llava/video_utils.py:38-68.  depth (V,H,W) -> world (V,H,W,3) fp32. */
void v3d_oracle_unproject(const void* depth, int depth_dtype, const float* poses, const float* intr,
                          int V, int H, int W, float depth_scale, float* world) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int v = 0; v < V; ++v) {
    for (int r = 0; r < H; ++r) {
      const float* K = intr + 16 * (size_t)v;
      const float* P = poses + 16 * (size_t)v;
      for (int c = 0; c < W; ++c) {
        size_t i = ((size_t)v * H + r) * W + c;
        unproject_px(load_depth(depth, depth_dtype, i), (float)c, (float)r, K, P, depth_scale,
                     world + 3 * i);
      }
    }
  }
}

/* Optional clamp of world coordinates to the mesh AABB
 * (llava/video_utils.py:232-234, the 'norm' strategy token). */
void v3d_oracle_clamp(float* world, size_t n_points, const float* lo, const float* hi) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n_points; ++i)
    for (int a = 0; a < 3; ++a)
      world[3 * i + a] = t_minimum(t_maximum(world[3 * i + a], lo[a]), hi[a]);
}

/* Scene AABB "boundry" of llava/video_utils.py:273-277: [xmin,xmax,ymin,ymax,zmin,zmax]. */
void v3d_oracle_boundary(const float* world, size_t n_points, float* out6) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  int has_nan[3] = {0, 0, 0};
  for (size_t i = 0; i < n_points; ++i)
    for (int a = 0; a < 3; ++a) {
      float x = world[3 * i + a];
      if (isnan(x)) has_nan[a] = 1;
      if (x < mn[a]) mn[a] = x;
      if (x > mx[a]) mx[a] = x;
    }
  for (int a = 0; a < 3; ++a) {
    out6[2 * a] = has_nan[a] ? NAN : mn[a];
    out6[2 * a + 1] = has_nan[a] ? NAN : mx[a];
  }
}

/* Nearest-neighbour resize + centre crop of llava/video_utils.py:293-308, expressed as
 * a gather: out[v,r,c] = world[v,row_map[r],col_map[c]].  For the default geometry
 * row_map[r] = floor(1.25 r), col_map[c] = floor(1.25 (c+64)) (cv2.INTER_NEAREST). */
void v3d_oracle_resample(const float* world, int V, int H, int W, const int32_t* row_map,
                         const int32_t* col_map, int OH, int OW, float* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int v = 0; v < V; ++v)
    for (int r = 0; r < OH; ++r)
      for (int c = 0; c < OW; ++c) {
        const float* s = world + 3 * (((size_t)v * H + row_map[r]) * W + col_map[c]);
        float* d = out + 3 * (((size_t)v * OH + r) * OW + c);
        d[0] = s[0];
        d[1] = s[1];
        d[2] = s[2];
      }
}

/* llava/model/llava_arch.py:213-223 (`average_coordinate_in_patch`): drop the last `trim`
 * rows/cols, avg_pool2d(kernel=stride=p).  ATen's CPU kernel is a sequential row-major
 * fp32 sum starting from 0, divided by p*p.  coords (V,S,S,3) -> out (V,G,G,3). */
void v3d_oracle_pool_avg(const float* coords, int V, int S, int trim, int p, float* out) {
  const int G = (S - trim) / p;
  const float div = (float)(p * p);
#pragma omp parallel for collapse(2) schedule(static)
  for (int v = 0; v < V; ++v)
    for (int gi = 0; gi < G; ++gi)
      for (int gj = 0; gj < G; ++gj)
        for (int a = 0; a < 3; ++a) {
          float s = 0.0f;
          for (int dy = 0; dy < p; ++dy)
            for (int dx = 0; dx < p; ++dx)
              s = s + coords[3 * (((size_t)v * S + gi * p + dy) * S + gj * p + dx) + a];
          out[3 * (((size_t)v * G + gi) * G + gj) + a] = s / div;
        }
}

/* llava/model/llava_arch.py:225-239 (`minmax_coordinate_in_patch`): out (V,G,G,2,3),
 * [..,0,:] = -maxpool(-x), [..,1,:] = maxpool(x).  ATen max_pool2d takes a value when it
 * is larger than the running max or NaN. */
void v3d_oracle_pool_minmax(const float* coords, int V, int S, int trim, int p, float* out) {
  const int G = (S - trim) / p;
#pragma omp parallel for collapse(2) schedule(static)
  for (int v = 0; v < V; ++v)
    for (int gi = 0; gi < G; ++gi)
      for (int gj = 0; gj < G; ++gj)
        for (int a = 0; a < 3; ++a) {
          float mx = -INFINITY, mneg = -INFINITY;
          for (int dy = 0; dy < p; ++dy)
            for (int dx = 0; dx < p; ++dx) {
              float x = coords[3 * (((size_t)v * S + gi * p + dy) * S + gj * p + dx) + a];
              if (x > mx || isnan(x)) mx = x;
              float nx = -x;
              if (nx > mneg || isnan(nx)) mneg = nx;
            }
          float* o = out + 6 * (((size_t)v * G + gi) * G + gj);
          o[a] = -mneg;
          o[3 + a] = mx;
        }
}

/* llava/model/llava_arch.py:241-257 (`sample_n_points`): samples at offsets {4,13,22}^2
 * of each 27-pixel window of the 378x378 trimmed grid; n_points 9 -> all, 5 -> every
 * second of the nine, 1 -> the centre.  coords (V,S,S,3) -> out (V,14,14,n,3). */
void v3d_oracle_sample_points(const float* coords, int V, int S, int n_points, float* out) {
  const int G = 14, p = 27;
  for (int v = 0; v < V; ++v)
    for (int gi = 0; gi < G; ++gi)
      for (int gj = 0; gj < G; ++gj) {
        int k = 0;
        for (int q = 0; q < 9; ++q) {
          int keep = (n_points == 9) || (n_points == 5 && (q % 2) == 0) || (n_points == 1 && q == 4);
          if (!keep) continue;
          int dy = 4 + 9 * (q / 3), dx = 4 + 9 * (q % 3);
          const float* s = coords + 3 * (((size_t)v * S + gi * p + dy) * S + gj * p + dx);
          float* o = out + 3 * ((((size_t)v * G + gi) * G + gj) * n_points + k);
          o[0] = s[0];
          o[1] = s[1];
          o[2] = s[2];
          ++k;
        }
      }
}

/* llava/model/llava_arch.py:259-272 (`discrete_coords`) and llava/video_utils.py:348-358:
 *   round((min(max(c, lo), hi) - lo) / voxel_size), round = half-to-even, result float.
 * div_mode IEEE is torch-CPU; RECIPROCAL (x * (1/voxel)) is what ATen's CUDA kernel does. */
void v3d_oracle_discretize(const float* c, size_t n_points, const float* lo, const float* hi,
                           float voxel_size, int div_mode, float* out) {
  const float rcp = 1.0f / voxel_size;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n_points; ++i)
    for (int a = 0; a < 3; ++a) {
      float x = t_minimum(t_maximum(c[3 * i + a], lo[a]), hi[a]) - lo[a];
      x = (div_mode == V3D_DIV_RECIPROCAL) ? x * rcp : x / voxel_size;
      out[3 * i + a] = rintf(x);
    }
}

/* llava/model/position_encoding.py:17-49 (`PositionEmbeddingSine3D.forward`).
 * x: (B*N_tok*n_points, 3) fp32; dim_t: (nf) fp32, nf = E // (3*n_points), supplied by the
 * caller because torch's powf is not correctly rounded (:24-25).  For each point and axis a,
 * feature j = sin(x_a/dim_t[j]) for even j, cos(x_a/dim_t[j]) for odd j (:27-42); a token
 * concatenates its n_points points (:43-44); channels past 3*nf*n_points are zero (:46-47).
 * out: (B*N_tok, E) fp32. */
void v3d_oracle_pe_sine3d(const float* x, size_t n_tokens, int n_points, int E, const float* dim_t,
                          int nf, float* out) {
#pragma omp parallel for schedule(static)
  for (size_t t = 0; t < n_tokens; ++t) {
    float* o = out + t * (size_t)E;
    memset(o, 0, sizeof(float) * (size_t)E);
    for (int pt = 0; pt < n_points; ++pt)
      for (int a = 0; a < 3; ++a) {
        float xa = x[3 * (t * n_points + pt) + a];
        float* oa = o + (size_t)(pt * 3 + a) * nf;
        for (int j = 0; j < nf; ++j) {
          float pos = xa / dim_t[j];
          oa[j] = (j & 1) ? cosf(pos) : sinf(pos);
        }
      }
  }
}

/* fp32 -> bf16 bits, round-to-nearest-even (torch's .to(torch.bfloat16)). */
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

void v3d_oracle_f32_to_bf16(const float* in, size_t n, uint16_t* out) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) out[i] = f32_to_bf16(in[i]);
}

/* llava/model/llava_arch.py:515: image_feat + PE, both bf16: the PE is rounded to bf16
 * first, the add is done in fp32 and rounded to bf16 again. */
void v3d_oracle_add_bf16(const uint16_t* tokens, const float* pe_f32, size_t n, uint16_t* out) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i)
    out[i] = f32_to_bf16(bf16_to_f32(tokens[i]) + bf16_to_f32(f32_to_bf16(pe_f32[i])));
}

/* ------------------------------------------------------------------------------------------
 * Maximum-coverage sampler, scripts/3d/preprocessing/max_coverage_sampling.py:39-104.
 * ---------------------------------------------------------------------------------------- */

/* :44-45  vox = int(round(world / voxel_size)), no clamp and no offset (:41-43 commented). */
void v3d_oracle_mc_voxel_ids(const float* world, size_t n_points, float voxel_size, int div_mode,
                             int32_t* ids) {
  const float rcp = 1.0f / voxel_size;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < 3 * n_points; ++i) {
    float x = (div_mode == V3D_DIV_RECIPROCAL) ? world[i] * rcp : world[i] / voxel_size;
    x = rintf(x);
    /* torch .int() of a non-finite / out-of-range float on x86 yields INT32_MIN */
    ids[i] = (x >= -2147483648.0f && x < 2147483648.0f) ? (int32_t)x : INT32_MIN;
  }
}

/* voxel triples are compared as Python compares tuples of ints: no range limit (a non-finite coordinate
 * becomes INT32_MIN at :45 and is a set element like any other) */
typedef struct {
  int32_t v[3];
  int32_t idx;
} kv_t;
static int key_cmp(const int32_t* x, const int32_t* y) {
  for (int a = 0; a < 3; ++a)
    if (x[a] != y[a]) return x[a] < y[a] ? -1 : 1;
  return 0;
}
static int kv_cmp(const void* a, const void* b) {
  int c = key_cmp(((const kv_t*)a)->v, ((const kv_t*)b)->v);
  if (c) return c;
  /* duplicates of a voxel: the first index of the list comes first (set semantics keep one element) */
  return ((const kv_t*)a)->idx < ((const kv_t*)b)->idx ? -1 : (((const kv_t*)a)->idx > ((const kv_t*)b)->idx);
}

/* Per-frame coverage bitsets over the mesh universe `pc` (n_pc int triples,
 * scripts/3d/preprocessing/convert_pcd_to_voxel.py:69-82): bit b of frame f is set iff voxel
 * pc[b] occurs among frame f's voxel ids, i.e. set(frame) & pc_voxel (:47-50, :75); a voxel listed twice
 * in pc is one set element, represented by its first index.
 * ids: (F, n_px, 3) int32.  bitsets: (F, words) uint32, words = ceil(n_pc/32). */
int v3d_oracle_mc_bitsets(const int32_t* ids, int F, size_t n_px, const int32_t* pc, int n_pc,
                          uint32_t* bitsets) {
  const int words = (n_pc + 31) / 32;
  kv_t* tab = (kv_t*)malloc(sizeof(kv_t) * (size_t)(n_pc > 0 ? n_pc : 1));
  if (!tab) return -1;
  for (int b = 0; b < n_pc; ++b) {
    memcpy(tab[b].v, pc + 3 * b, sizeof(int32_t) * 3);
    tab[b].idx = b;
  }
  qsort(tab, (size_t)n_pc, sizeof(kv_t), kv_cmp);
  memset(bitsets, 0, sizeof(uint32_t) * (size_t)F * words);
#pragma omp parallel for schedule(dynamic, 1)
  for (int f = 0; f < F; ++f) {
    uint32_t* bs = bitsets + (size_t)f * words;
    const int32_t* last = NULL;
    for (size_t i = 0; i < n_px; ++i) {
      const int32_t* v = ids + 3 * ((size_t)f * n_px + i);
      if (last && key_cmp(v, last) == 0) continue;
      last = v;
      int lo = 0, hi = n_pc - 1, hit = -1;
      while (lo <= hi) { /* leftmost match = smallest index among duplicates */
        int mid = (lo + hi) >> 1;
        int c = key_cmp(tab[mid].v, v);
        if (c < 0) lo = mid + 1;
        else {
          if (c == 0) hit = mid;
          hi = mid - 1;
        }
      }
      if (hit >= 0) {
        int b = tab[hit].idx;
        bs[b >> 5] |= 1u << (b & 31);
      }
    }
  }
  free(tab);
  return 0;
}

/* CPython's MT19937 (Modules/_randommodule.c): random.seed(int) = init_by_array(key words of
 * abs(seed)), getrandbits(k<=32) = genrand_uint32() >> (32-k), Random._randbelow(n) =
 * rejection loop on getrandbits(n.bit_length()); random.choice(seq) = seq[_randbelow(len)]. */
typedef struct {
  uint32_t mt[624];
  int idx;
} mt_t;

static void mt_init_genrand(mt_t* s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; ++i) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->idx = 624;
}
static void mt_init_by_array(mt_t* s, const uint32_t* key, int klen) {
  mt_init_genrand(s, 19650218u);
  int i = 1, j = 0;
  int k = 624 > klen ? 624 : klen;
  for (; k; --k) {
    s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    ++i; ++j;
    if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    if (j >= klen) j = 0;
  }
  for (k = 623; k; --k) {
    s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    ++i;
    if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
  }
  s->mt[0] = 0x80000000u;
}
static void mt_seed_u64(mt_t* s, uint64_t seed) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  mt_init_by_array(s, key, key[1] ? 2 : 1);
}
static uint32_t mt_next(mt_t* s) {
  if (s->idx >= 624) {
    uint32_t* mt = s->mt;
    int kk;
    for (kk = 0; kk < 624 - 397; ++kk) {
      uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < 623; ++kk) {
      uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s->idx = 0;
  }
  uint32_t y = s->mt[s->idx++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
static uint32_t mt_randbelow(mt_t* s, uint32_t n) {
  int k = 32 - __builtin_clz(n); /* n.bit_length(), n >= 1 */
  uint32_t r = mt_next(s) >> (32 - k);
  while (r >= n) r = mt_next(s) >> (32 - k);
  return r;
}

/* Exposed for the known-answer test against CPython's `random` module. */
void v3d_oracle_mt_randbelow_seq(uint64_t seed, const uint32_t* ns, int count, uint32_t* out) {
  mt_t s;
  mt_seed_u64(&s, seed);
  for (int i = 0; i < count; ++i) out[i] = mt_randbelow(&s, ns[i]);
}

/* Greedy set cover, max_coverage_sampling.py:68-104.  Each round scans the remaining
 * candidates in ascending order, gain = |S_f \ used| (pc-restricted, :75-77); strictly
 * larger gain restarts the tie list, equal gain appends (:78-82); the pick is
 * random.choice(ties) (:84) under the protocol random.seed(seed) right before the loop;
 * stops after max_pick picks (:93, 32 upstream) or when candidates run out (:72).
 * Outputs: order/gains (max_pick, -1 padded), counts = {n_picked, n_all, n_sel}
 * where n_all = |union of all frames| (:102) and n_sel = |used| (:103). */
void v3d_oracle_mc_greedy(const uint32_t* bitsets, int F, int words, int max_pick, uint64_t seed,
                          int32_t* order, int32_t* gains, int32_t* counts) {
  uint32_t* used = (uint32_t*)calloc((size_t)(words > 0 ? words : 1), sizeof(uint32_t));
  uint8_t* taken = (uint8_t*)calloc((size_t)(F > 0 ? F : 1), 1);
  int32_t* ties = (int32_t*)malloc(sizeof(int32_t) * (size_t)(F > 0 ? F : 1));
  mt_t rng;
  mt_seed_u64(&rng, seed);
  for (int i = 0; i < max_pick; ++i) order[i] = gains[i] = -1;
  int n_picked = 0;
  for (int round = 0; round < F && n_picked < max_pick; ++round) {
    int maxv = -1, n_ties = 0;
    for (int f = 0; f < F; ++f) {
      if (taken[f]) continue;
      const uint32_t* bs = bitsets + (size_t)f * words;
      int g = 0;
      for (int w = 0; w < words; ++w) g += __builtin_popcount(bs[w] & ~used[w]);
      if (g > maxv) { maxv = g; n_ties = 0; ties[n_ties++] = f; }
      else if (g == maxv) ties[n_ties++] = f;
    }
    int pick = ties[mt_randbelow(&rng, (uint32_t)n_ties)];
    const uint32_t* bs = bitsets + (size_t)pick * words;
    for (int w = 0; w < words; ++w) used[w] |= bs[w];
    taken[pick] = 1;
    order[n_picked] = pick;
    gains[n_picked] = maxv;
    ++n_picked;
  }
  int n_all = 0, n_sel = 0;
  for (int w = 0; w < words; ++w) {
    uint32_t u = 0;
    for (int f = 0; f < F; ++f) u |= bitsets[(size_t)f * words + w];
    n_all += __builtin_popcount(u);
    n_sel += __builtin_popcount(used[w]);
  }
  counts[0] = n_picked;
  counts[1] = n_all;
  counts[2] = n_sel;
  free(used);
  free(taken);
  free(ties);
}
