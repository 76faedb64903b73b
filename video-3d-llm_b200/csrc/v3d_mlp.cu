// v3d_mlp.cu -- learned-projection PE on tcgen05 / TMEM (C ABI: v3d_pe_mlp, include/v3d.h).
//
// Replaces llava/model/position_encoding.py:52-84 (`PositionEmbeddingMLP.forward`) + the add of
// llava/model/llava_arch.py:515:
//     out = tokens + Linear(512,E)( ReLU( LayerNorm(512)( Linear(3,512)(coords) ) ) )
// The 512-wide hidden activations are produced inside the kernel and never touch HBM; the 512 -> E
// contraction (the only dense contraction on the path) runs on the 5th-generation tensor cores.
//
// One persistent CTA per SM, warp-specialised (640 threads in the shipped variant):
//   warp 0    TMA producer: W2 tiles (BLOCK_N rows x 64 k, 128-byte swizzle) through a ring of mbarriers
//   warp 1    MMA issuer  : tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16; owns TMEM alloc
//   warps 4.. (PARTS per TMEM lane quadrant -- 4 in the shipped variant -- each with a 1/PARTS share of the k range when the
//             A operand is built and of an N tile's columns in the epilogue)
//             (a) build the A operand of an M tile (128 tokens x 512 hidden, bf16): Linear(3,512) + LayerNorm
//             + ReLU per token, one thread per token row = one TMEM lane, written with tcgen05.st straight
//             into TENSOR MEMORY (256 columns: two bf16 per 32-bit cell), from where the MMA reads it.  The A operand of
//             the NEXT M tile is built as soon as the last accumulator of the current one is full, before it is drained;
//             (b) epilogue: tcgen05.ld the fp32 accumulator, + bias, round to bf16, + tokens, store.  Token blocks arrive
//             in per-warp staging tiles by cp.async.bulk.tensor (two tiles per warp, the next N tile's block in flight),
//             results leave from the same tiles by TMA stores; the (128 x BLOCK_N) token block of an N tile is pulled
//             into L2 two tiles ahead by cp.async.bulk.prefetch.tensor.
// TMEM map (512 columns): [0, 2*BLOCK_N) double-buffered fp32 accumulators (the MMAs of N tile i+1 overlap
// the epilogue of N tile i), [256, 512) the A operand, resident for all N tiles of its M tile.  Keeping A out
// of shared memory leaves room for a deep W2 ring: with A in smem (128 KB) only 3 x 16 KB of W2 fitted in
// flight and the tensor pipe starved on W2 tiles (24 % active, profiles/r1i_*).  Measured limits and what was tried:
// DESIGN.md section 4.4.
#include <cuda.h>

#include "v3d_common.cuh"

namespace v3d {

// warps: TMA, MMA, 2 spare, then PARTS producer/epilogue warps per TMEM lane quadrant (each a 1/PARTS share of the k range
// when the A operand is built and of an N tile's columns in the epilogue)
__host__ __device__ constexpr int mlp_threads(int parts) { return 128 + 128 * parts; }
static constexpr int kHidden = 512;          // nn.Linear(3, 512) of the reference module
static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;           // 64 bf16 = 128 bytes = one swizzle atom
static constexpr int kNumKBlocks = kHidden / kBlockK;  // 8
#ifndef V3D_MLP_STAGES
#define V3D_MLP_STAGES 8
#endif
#ifndef V3D_MLP_PREFETCH
#define V3D_MLP_PREFETCH 0
#endif
#ifndef V3D_MLP_TMA_STAGE
#define V3D_MLP_TMA_STAGE 1  // epilogue staging tiles filled / drained by TMA (0: plain loads and stores)
#endif
#ifndef V3D_MLP_PARTS
#define V3D_MLP_PARTS 4  // producer/epilogue warps per TMEM lane quadrant of the TMA-staged variant (2 or 4)
#endif
#ifndef V3D_MLP_MCAST
#define V3D_MLP_MCAST 0  // 1: W2 stages multicast to the two CTAs of a cluster (measured: no gain)
#endif
#ifndef V3D_MLP_TOK_PF
#define V3D_MLP_TOK_PF 2  // N tiles ahead of the epilogue at which a tile's token block is pulled into L2 (0 = off)
#endif
static constexpr int kStages = V3D_MLP_STAGES;  // W2 ring (16 KB per stage at BLOCK_N = 128)
static constexpr uint32_t kTmemColsA = 256;  // first TMEM column of the A operand (128 lanes x 256 cols x 2 bf16)

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4
// in bits [0,14), LBO (unused for one swizzle atom along K) in [16,30), SBO = 8 rows x 128 B = 1024 B
// (>> 4) in [32,46), version 1 in [46,48), layout type SWIZZLE_128B (2) in [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(const void* smem_ptr) {
  const uint64_t addr = (uint64_t)((smem_u32(smem_ptr) & 0x3FFFFu) >> 4);
  return addr | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = fp32, A = B = bf16, both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory (lane = row, 32-bit column c = k elements 2c, 2c+1)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#ifdef V3D_MLP_TRACE
// diagnostic build: cycles the MMA thread of CTA 0 spends waiting for (0) the A operand, (1) a drained accumulator, (2) a W2
// stage, and (3) its total; (4) cycles an epilogue warp of CTA 0 spends producing A
__device__ unsigned long long g_mlp_trace[8];
#define MLP_T0() const long long t0__ = clock64()
#define MLP_ACC(k) g_mlp_trace_local[k] += clock64() - t0__
#else
#define MLP_T0() do { } while (0)
#define MLP_ACC(k) do { } while (0)
#endif

struct MlpParams {
  const float* coords;   // (M,3)
  const float* w1;       // (512,3)
  const float* b1;       // (512)
  const float* gamma;    // (512)
  const float* beta;     // (512)
  const float* b2;       // (E)
  const __nv_bfloat16* tokens;  // (M,E) or null
  __nv_bfloat16* out;    // (M,E)
  int M, E;
  float eps;
};

// shared memory map (dynamic, 1024-byte aligned base):
//   [0, S*BN*128)            W2 ring: S stages of [BLOCK_N rows][128 B], SWIZZLE_128B
//   then: float4 gw1[512], float2 gb[512], barriers, tmem base, LN moments, epilogue staging
static constexpr int kMaxE = 8192;  // bias staged in shared memory

// TMA_STAGE: the epilogue's token blocks arrive by TMA (cp.async.bulk.tensor, SWIZZLE_128B = the staging tile's XOR
// pattern) into two staging tiles per warp and the results leave by TMA stores from the same tiles.
template <int BLOCK_N, bool TMA_STAGE, int PARTS>
struct MlpSmem {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kOffB = 0;
  static constexpr int kOffW1 = kOffB + kStages * kBBytes;
  static constexpr int kOffGB = kOffW1 + kHidden * 16;
  static constexpr int kOffBar = kOffGB + kHidden * 8;
  // barriers: b_full[S], b_empty[S], a_full, tmem_full[2], tmem_empty[2], tok_full[4*PARTS warps][2]
  static constexpr int kNumBars = 2 * kStages + 1 + 4 + 8 * PARTS;
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kOffMom = kOffTmem + 16;   // 16 floats: LayerNorm moment data
  static constexpr int kOffStage = (kOffMom + 64 + 1023) & ~1023;  // epilogue staging: 8 warps x (1 or 2) x 32 rows x BLOCK_N/2 bf16
  static constexpr int kStageBytesPerWarp = 32 * (BLOCK_N / PARTS) * 2;  // 32 rows x this warp's share of an N tile, bf16
  static constexpr int kStageBufs = TMA_STAGE ? 2 : 1;
  static constexpr int kOffBias = kOffStage + 4 * PARTS * kStageBufs * kStageBytesPerWarp;  // b2 (E floats)
  static constexpr int total(int E) { return kOffBias + E * 4; }
};

// MCAST: launched in clusters of two CTAs that walk their M tiles in lockstep; every W2 stage is fetched from L2 once and
// delivered to both CTAs' rings by one multicast TMA load (the CTAs issue alternate stages), which halves the W2 stream
// out of L2 -- with one CTA per SM each re-reading all of W2 for every 128 tokens that stream, not HBM, is what the
// tensor pipe waits for.
template <int BLOCK_N, bool TMA_STAGE, int PARTS, bool MCAST>
__global__ void __launch_bounds__(mlp_threads(PARTS), 1) k_pe_mlp(const __grid_constant__ CUtensorMap w2_map, const __grid_constant__ CUtensorMap tok_map,
                                                           const __grid_constant__ CUtensorMap tok_io_map,
                                                           const __grid_constant__ CUtensorMap out_io_map, const MlpParams p) {
  using S = MlpSmem<BLOCK_N, TMA_STAGE, PARTS>;
  constexpr int kMlpThreads = mlp_threads(PARTS), kEpiThreads = 128 * PARTS;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  unsigned char* s_b = smem + S::kOffB;
  float4* s_w1 = reinterpret_cast<float4*>(smem + S::kOffW1);
  float2* s_gb = reinterpret_cast<float2*>(smem + S::kOffGB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* b_full = bars;
  uint64_t* b_empty = bars + kStages;
  uint64_t* a_full = bars + 2 * kStages;
  uint64_t* t_full = bars + 2 * kStages + 1;
  uint64_t* t_empty = bars + 2 * kStages + 3;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + S::kOffTmem);

  static_assert(2 * BLOCK_N <= (int)kTmemColsA, "accumulators overlap the A operand");
  constexpr uint32_t kTmemCols = 512;

  // ---- one-time setup
  // LayerNorm statistics in closed form.  h_k = w_k . c + b_k is affine in the token's coordinates c, so
  //   mean_k h      = wbar . c + bbar
  //   var_k  h      = c^T S c + 2 u . c + v,   S = mean (w_k-wbar)(w_k-wbar)^T, u = mean (w_k-wbar)(b_k-bbar), v = mean (b_k-bbar)^2
  // (all moments centred, so nothing cancels against the mean).  One warp reduces the 512 rows once per CTA;
  // per token the statistics then cost ~20 flops instead of a 512-long pass.
  float* s_mom = reinterpret_cast<float*>(smem + S::kOffMom);
  if (warp == 3) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < kHidden; k += 32) {
      a[0] += __ldg(p.w1 + 3 * k);
      a[1] += __ldg(p.w1 + 3 * k + 1);
      a[2] += __ldg(p.w1 + 3 * k + 2);
      a[3] += __ldg(p.b1 + k);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      for (int o = 16; o; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
      a[i] *= (1.0f / kHidden);
    }
    float m[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // S00 S01 S02 S11 S12 S22 u0 u1 u2 v
    for (int k = lane; k < kHidden; k += 32) {
      const float d0 = __ldg(p.w1 + 3 * k) - a[0], d1 = __ldg(p.w1 + 3 * k + 1) - a[1], d2 = __ldg(p.w1 + 3 * k + 2) - a[2];
      const float db = __ldg(p.b1 + k) - a[3];
      m[0] += d0 * d0; m[1] += d0 * d1; m[2] += d0 * d2; m[3] += d1 * d1; m[4] += d1 * d2; m[5] += d2 * d2;
      m[6] += d0 * db; m[7] += d1 * db; m[8] += d2 * db; m[9] += db * db;
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      for (int o = 16; o; o >>= 1) m[i] += __shfl_xor_sync(0xffffffffu, m[i], o);
      if (lane == 0) s_mom[4 + i] = m[i] * (1.0f / kHidden);
    }
    if (lane == 0) {
      s_mom[0] = a[0]; s_mom[1] = a[1]; s_mom[2] = a[2]; s_mom[3] = a[3];
    }
  }
  float* s_bias = reinterpret_cast<float*>(smem + S::kOffBias);
  for (int e = threadIdx.x; e < p.E; e += kMlpThreads) s_bias[e] = __ldg(p.b2 + e);
  for (int k = threadIdx.x; k < kHidden; k += kMlpThreads) {
    const float g = __ldg(p.gamma + k);
    // gamma folded into the first Linear: y_k = relu(rstd * (gw_k . c + gb_k) + (beta_k - rstd * mean * gamma_k))
    s_w1[k] = make_float4(g * __ldg(p.w1 + 3 * k), g * __ldg(p.w1 + 3 * k + 1), g * __ldg(p.w1 + 3 * k + 2), g * __ldg(p.b1 + k));
    s_gb[k] = make_float2(g, __ldg(p.beta + k));
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&w2_map) : "memory");
    for (int i = 0; i < kStages; ++i) {
      mbar_init(b_full + i, 1);
      mbar_init(b_empty + i, MCAST ? 2 : 1);  // MCAST: the MMAs of both CTAs have read the stage
    }
    mbar_init(a_full, kEpiThreads);
    for (int i = 0; i < 2; ++i) {
      mbar_init(t_full + i, 1);
      mbar_init(t_empty + i, kEpiThreads);
    }
    for (int i = 0; i < 8 * PARTS; ++i) mbar_init(bars + 2 * kStages + 5 + i, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int n_tiles = p.E / BLOCK_N;
  // M tiles of this CTA: blockIdx.x + j * gridDim.x.  With MCAST both CTAs of a cluster make the same number of passes;
  // a pass past the end recomputes the last tile without storing it.
  const int n_iter = MCAST ? (m_tiles + (int)gridDim.x - 1) / (int)gridDim.x : (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int j) { return min((int)blockIdx.x + j * (int)gridDim.x, m_tiles - 1); };
  uint32_t cta_rank = 0;
  if constexpr (MCAST) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    // the peer's barriers are initialised before anything of ours reaches them
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }

  if (warp == 0) {
    // ================= TMA producer: W2 tiles for every (m tile, n tile, k block) =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int j = 0; j < n_iter; ++j)
        for (int nt = 0; nt < n_tiles; ++nt)
          for (int kb = 0; kb < kNumKBlocks; ++kb, ++it) {
            const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
            mbar_wait(b_empty + stage, phase ^ 1);
            mbar_arrive_expect_tx(b_full + stage, (uint32_t)S::kBBytes);
            if constexpr (MCAST) {
              if ((it & 1u) == cta_rank)  // one load, delivered to the same slot and barrier of both CTAs
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
                        smem_u32(s_b + stage * S::kBBytes)),
                    "l"(&w2_map), "r"(kb * kBlockK), "r"(nt * BLOCK_N), "r"(smem_u32(b_full + stage)), "h"((uint16_t)3)
                    : "memory");
            } else {
              tma_load_2d(s_b + stage * S::kBBytes, &w2_map, kb * kBlockK, nt * BLOCK_N, b_full + stage);
            }
          }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N);
      uint32_t it = 0, acc_it = 0, a_phase = 0;
#ifdef V3D_MLP_TRACE
      long long g_mlp_trace_local[4] = {0, 0, 0, 0};
      const long long t_start = clock64();
#endif
      for (int j = 0; j < n_iter; ++j, a_phase ^= 1) {
        { MLP_T0(); mbar_wait(a_full, a_phase); MLP_ACC(0); }  // A operand of this M tile is in tensor memory
        tcgen05_fence_after();
        for (int nt = 0; nt < n_tiles; ++nt, ++acc_it) {
          const uint32_t acc = acc_it & 1, acc_phase = (acc_it >> 1) & 1;
          { MLP_T0(); mbar_wait(t_empty + acc, acc_phase ^ 1); MLP_ACC(1); }  // epilogue has drained this accumulator
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
          for (int kb = 0; kb < kNumKBlocks; ++kb, ++it) {
            const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
            { MLP_T0(); mbar_wait(b_full + stage, phase); MLP_ACC(2); }
            tcgen05_fence_after();
            const uint32_t ta = tmem_base + kTmemColsA + kb * (kBlockK / 2);
            const uint64_t db = make_kmajor_sw128_desc(s_b + stage * S::kBBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)  // UMMA_K = 16 bf16: 8 TMEM columns of A, 32 bytes (+2 in the >>4 field) of B
              umma_bf16_ts(tmem_d, ta + 8 * k, db + 2 * k, idesc, (kb | k) != 0);
            // frees the W2 slot when these MMAs have read it (MCAST: in both CTAs, whose loads fill both rings)
            if constexpr (MCAST)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                               smem_u32(b_empty + stage)),
                           "h"((uint16_t)3)
                           : "memory");
            else
              umma_commit(b_empty + stage);
          }
          umma_commit(t_full + acc);  // accumulator complete
        }
      }
#ifdef V3D_MLP_TRACE
      if (blockIdx.x == 0) {
        for (int i = 0; i < 3; ++i) g_mlp_trace[i] = (unsigned long long)g_mlp_trace_local[i];
        g_mlp_trace[3] = (unsigned long long)(clock64() - t_start);
      }
#endif
    }
  } else if (warp >= 4) {
    // ================= A producer + epilogue: thread = token row = TMEM lane =====================
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;  // which share of the k range (A production) / of the N tile's columns (epilogue)
    const int r = q * 32 + lane;       // row inside the M tile
    uint32_t acc_it = 0, tile_seq = 0;
    // The epilogue reads the tokens of an N tile as 128 rows x 256 bytes, 2304+ bytes apart: left to plain loads that is
    // one DRAM round trip per tile on the critical path of the accumulator hand-over (ncu: a fifth of all stall samples
    // on the first use of those loads).  One thread asks the TMA unit to pull the (128 x BLOCK_N) block into L2
    // V3D_MLP_TOK_PF tiles ahead -- across M tiles -- so that the loads, issued one tile ahead, hit L2.
    auto prefetch_tokens = [&](int mt_, int nt_) {
      if (V3D_MLP_TOK_PF > 0 && threadIdx.x == 128 && p.tokens != nullptr) {
        if (nt_ >= n_tiles) {
          nt_ -= n_tiles;
          mt_ += (int)gridDim.x;
        }
        if (mt_ < m_tiles && nt_ < n_tiles)
          asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tok_map), "r"(nt_ * BLOCK_N), "r"(mt_ * kBlockM)
                       : "memory");
      }
    };
    for (int i = 0; i < V3D_MLP_TOK_PF; ++i) prefetch_tokens((int)blockIdx.x, i);
    // (a) the A operand of M tile mt_: hidden activations of this thread's token (Linear(3,512) -> LayerNorm -> ReLU -> bf16)
    // into tensor memory, then one arrival on a_full
    auto produce_a = [&](int mt_) {
      const int m = mt_ * kBlockM + r;
      const bool row_ok = m < p.M;
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (row_ok) {
        c0 = __ldg(p.coords + 3 * (size_t)m);
        c1 = __ldg(p.coords + 3 * (size_t)m + 1);
        c2 = __ldg(p.coords + 3 * (size_t)m + 2);
      }
      // LayerNorm statistics of this token in closed form (biased variance, as torch)
      const float mean = fmaf(s_mom[0], c0, fmaf(s_mom[1], c1, fmaf(s_mom[2], c2, s_mom[3])));
      float var = s_mom[4] * c0 * c0 + s_mom[7] * c1 * c1 + s_mom[9] * c2 * c2 +
                  2.0f * (s_mom[5] * c0 * c1 + s_mom[6] * c0 * c2 + s_mom[8] * c1 * c2) +
                  2.0f * (s_mom[10] * c0 + s_mom[11] * c1 + s_mom[12] * c2) + s_mom[13];
      var = fmaxf(var, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      const float nmr = -mean * rstd;
      const uint32_t ta_row = tmem_base + ((uint32_t)(q * 32) << 16) + kTmemColsA;
#pragma unroll 1
      for (int kb = half * (kNumKBlocks / PARTS); kb < (half + 1) * (kNumKBlocks / PARTS); ++kb) {  // 64 features = 32 TMEM columns per store
        uint32_t packed[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float y[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int k = kb * kBlockK + 2 * i + j;
            const float4 w = s_w1[k];   // gamma-scaled first layer
            const float2 gb = s_gb[k];  // {gamma, beta}
            const float t = fmaf(w.x, c0, fmaf(w.y, c1, fmaf(w.z, c2, w.w)));
            y[j] = fmaxf(fmaf(rstd, t, fmaf(nmr, gb.x, gb.y)), 0.f);
          }
          const __nv_bfloat162 b = __floats2bfloat162_rn(y[0], y[1]);  // low half = even k
          packed[i] = *reinterpret_cast<const uint32_t*>(&b);
        }
        tmem_st_32x32b_x32(ta_row + kb * (kBlockK / 2), packed);
      }
      tmem_st_wait();
      tcgen05_fence_before();  // order the TMEM stores before the MMA thread's reads
      mbar_arrive(a_full);
    };
    for (int j = 0; j < n_iter; ++j) {
      const int mt = tile_of(j);
      const bool real = (int)blockIdx.x + j * (int)gridDim.x < m_tiles;  // a pass past the end (MCAST) stores nothing
      // L2 prefetch of the token rows by the TMA unit, one M tile ahead of their use in the epilogue (the very
      // first tile is requested here, before its A operand is built)
      if (V3D_MLP_PREFETCH && threadIdx.x == 128 && p.tokens != nullptr) {  // (off by default)
        for (int t = (mt == (int)blockIdx.x ? mt : mt + (int)gridDim.x); t <= mt + (int)gridDim.x && t < m_tiles; t += gridDim.x) {
          const size_t rows = (size_t)min(kBlockM, p.M - t * kBlockM);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(p.tokens + (size_t)t * kBlockM * p.E);
          const size_t bytes = rows * (size_t)p.E * 2;
          for (size_t o = 0; o < bytes; o += 32768) tma_prefetch_l2(src + o, (uint32_t)min((size_t)32768, bytes - o));
        }
      }
      // With the TMA-staged epilogue the A operand of the NEXT M tile is built before the last N tile of this one is
      // drained (below), so only the first M tile is built here.
      if (!TMA_STAGE || j == 0) produce_a(mt);

      // ---- (b) epilogue of every N tile of this M tile.
      // Global traffic is warp-coalesced: a warp instruction moves 512 contiguous bytes (2 rows x 256 B at
      // BLOCK_N = 128) of the warp's own 32 rows, and a per-warp staging tile in shared memory (XOR-swizzled,
      // conflict-free both ways) turns that into the row-per-thread order of the TMEM accumulator and back.
      // The token segment of the NEXT N tile is requested before the current accumulator is touched, so
      // its latency hides behind this tile's TMEM loads, math and stores.
      constexpr int kHalfN = BLOCK_N / PARTS;         // columns of an N tile this warp owns
      constexpr int kChunks = kHalfN / 8;             // 16-byte chunks per row
      const int col0 = half * kHalfN;                // first column (inside the N tile) of this warp
      const int row0 = mt * kBlockM + q * 32;        // first global row of this warp
      const bool add = p.tokens != nullptr;
      if constexpr (TMA_STAGE) {
        // ---- staging tiles filled and drained by the TMA unit.  Tile sequence of this warp: every (M tile, N tile) of the CTA
        // in order, `seq` counts them; tile s uses staging buffer s & 1.  The token block of tile s+1 is requested at the
        // start of tile s, into the buffer tile s-1 has just been stored from.
        static_assert(!TMA_STAGE || kChunks == 8 || kChunks == 4, "staging rows are one 128-byte or 64-byte swizzle atom");
        uint64_t* tokbar = bars + 2 * kStages + 5 + (warp - 4) * 2;
        unsigned char* stage0 = smem + S::kOffStage + (warp - 4) * 2 * S::kStageBytesPerWarp;
        auto issue_load = [&](int mt_, int nt_, uint32_t buf) {  // lane 0
          mbar_arrive_expect_tx(tokbar + buf, (uint32_t)S::kStageBytesPerWarp);
          tma_load_2d(stage0 + buf * S::kStageBytesPerWarp, &tok_io_map, nt_ * BLOCK_N + col0, mt_ * kBlockM + q * 32, tokbar + buf);
        };
        if (add && lane == 0 && j == 0) {  // the first two tiles of this warp
          issue_load(mt, 0, 0);
          if (n_tiles > 1) issue_load(mt, 1, 1);
          else if (j + 1 < n_iter) issue_load(tile_of(j + 1), 0, 1);
        }
        for (int nt = 0; nt < n_tiles; ++nt, ++tile_seq) {
          const uint32_t buf = tile_seq & 1u;
          unsigned char* stage = stage0 + buf * S::kStageBytesPerWarp;
          const uint32_t acc = acc_it & 1, acc_phase = (acc_it >> 1) & 1;
          prefetch_tokens(mt, nt + V3D_MLP_TOK_PF);
          if (nt == n_tiles - 1 && j + 1 < n_iter) {
            // Every MMA of this M tile has completed once the last accumulator is full: the A operand is free.  Build the
            // next M tile's A first -- its MMAs then run (into the other accumulator) while this last N tile is drained --
            // instead of leaving the tensor pipe idle through the last epilogue AND the A production.
            mbar_wait(t_full + acc, acc_phase);
            tcgen05_fence_after();
            produce_a(tile_of(j + 1));
          }
          if (lane == 0 && tile_seq >= 1) {
            // the previous tile's store has read its buffer: refill it with the next tile's tokens
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (add) {
              if (nt + 1 < n_tiles) issue_load(mt, nt + 1, buf ^ 1u);
              else if (j + 1 < n_iter) issue_load(tile_of(j + 1), 0, buf ^ 1u);
            }
          }
          if (add) mbar_wait(tokbar + buf, (tile_seq >> 1) & 1u);
          mbar_wait(t_full + acc, acc_phase);
          tcgen05_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + col0;
#pragma unroll
          for (int cc = 0; cc < kHalfN / 32; ++cc) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + cc * 32, v);
            tmem_ld_wait();
            const float* bias = s_bias + nt * BLOCK_N + col0 + cc * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t o[4];
              const int chunk = cc * 4 + i;
              // the TMA unit's swizzle: 16-byte chunk index ^ address bits [7,10) (128-byte rows) / [7,9) (64-byte rows)
              const int sw = (kChunks == 8) ? (lane & 7) : ((lane >> 1) & 3);
              uint4* slot = reinterpret_cast<uint4*>(stage + lane * (kHalfN * 2) + ((chunk ^ sw) << 4));
              uint4 tk = make_uint4(0u, 0u, 0u, 0u);
              if (add) tk = *slot;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = i * 8 + 2 * j;
                // Linear output in bf16 (fp32 accumulate + bias, one rounding), then the bf16 add of llava_arch.py:515
                const float2 bb = *reinterpret_cast<const float2*>(bias + e);  // same address for the whole warp: broadcast
                __nv_bfloat162 pe = __floats2bfloat162_rn(__uint_as_float(v[e]) + bb.x, __uint_as_float(v[e + 1]) + bb.y);
                if (add) {
                  const uint32_t tw = (j == 0) ? tk.x : (j == 1) ? tk.y : (j == 2) ? tk.z : tk.w;
                  pe = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&tw), pe);
                }
                o[j] = *reinterpret_cast<const uint32_t*>(&pe);
              }
              *slot = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
          tcgen05_fence_before();
          mbar_arrive(t_empty + acc);  // 256 arrivals release the accumulator
          fence_proxy_async_smem();    // the tile's shared-memory writes are visible to the TMA store
          __syncwarp();
          if (lane == 0 && real) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&out_io_map), "r"(smem_u32(stage)),
                         "r"(nt * BLOCK_N + col0), "r"(row0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++acc_it;
        }
      } else {
      constexpr int kRowsPerInst = 32 / kChunks;      // rows covered by one warp instruction
      constexpr int kInst = 32 / kRowsPerInst;        // warp instructions per N tile (= kChunks)
      unsigned char* stage = smem + S::kOffStage + (warp - 4) * S::kStageBytesPerWarp;
      const int my_chunk = lane % kChunks, my_rsub = lane / kChunks;
      uint4 tok_a[kInst], tok_b[kInst];
      auto load_tokens = [&](uint4 (&t)[kInst], int nt) {
        if (add && nt < n_tiles) {
#pragma unroll
          for (int i = 0; i < kInst; ++i) {
            const int gr = row0 + i * kRowsPerInst + my_rsub;
            if (gr < p.M) t[i] = ld_stream_v4(p.tokens + (size_t)gr * p.E + (size_t)nt * BLOCK_N + col0 + my_chunk * 8);
          }
        }
      };
      auto stage_addr = [&](int row, int chunk) -> unsigned char* {
        return stage + row * (kHalfN * 2) + ((chunk ^ (row & (kChunks - 1))) << 4);
      };
      auto epilogue_tile = [&](const uint4 (&t)[kInst], int nt) {
        const uint32_t acc = acc_it & 1, acc_phase = (acc_it >> 1) & 1;
        prefetch_tokens(mt, nt + V3D_MLP_TOK_PF);
        if (add) {
#pragma unroll
          for (int i = 0; i < kInst; ++i) *reinterpret_cast<uint4*>(stage_addr(i * kRowsPerInst + my_rsub, my_chunk)) = t[i];
          __syncwarp();
        }
        mbar_wait(t_full + acc, acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + col0;
#pragma unroll
        for (int cc = 0; cc < kHalfN / 32; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + cc * 32, v);
          tmem_ld_wait();
          const float* bias = s_bias + nt * BLOCK_N + col0 + cc * 32;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t o[4];
            uint4* slot = reinterpret_cast<uint4*>(stage_addr(lane, cc * 4 + i));
            uint4 tk = make_uint4(0u, 0u, 0u, 0u);
            if (add) tk = *slot;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = i * 8 + 2 * j;
              // Linear output in bf16 (fp32 accumulate + bias, one rounding), then the bf16 add of llava_arch.py:515
              const float2 bb = *reinterpret_cast<const float2*>(bias + e);  // same address for the whole warp: broadcast
              __nv_bfloat162 pe = __floats2bfloat162_rn(__uint_as_float(v[e]) + bb.x, __uint_as_float(v[e + 1]) + bb.y);
              if (add) {
                const uint32_t tw = (j == 0) ? tk.x : (j == 1) ? tk.y : (j == 2) ? tk.z : tk.w;
                pe = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&tw), pe);
              }
              o[j] = *reinterpret_cast<const uint32_t*>(&pe);
            }
            *slot = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        tcgen05_fence_before();
        mbar_arrive(t_empty + acc);  // 256 arrivals release the accumulator
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kInst; ++i) {
          const int gr = row0 + i * kRowsPerInst + my_rsub;
          const uint4 val = *reinterpret_cast<const uint4*>(stage_addr(i * kRowsPerInst + my_rsub, my_chunk));
          if (gr < p.M) st_stream_v4(p.out + (size_t)gr * p.E + (size_t)nt * BLOCK_N + col0 + my_chunk * 8, val);
        }
        __syncwarp();
        ++acc_it;
      };
      load_tokens(tok_a, 0);
      for (int nt = 0; nt < n_tiles; nt += 2) {
        load_tokens(tok_b, nt + 1);
        epilogue_tile(tok_a, nt);
        if (nt + 1 < n_tiles) {
          load_tokens(tok_a, nt + 2);
          epilogue_tile(tok_b, nt + 1);
        }
      }
      }
    }
  }
  // ---- teardown
  if constexpr (MCAST)  // the peer may still multicast into this CTA's ring / arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (TMA_STAGE && warp >= 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // this warp's TMA stores are done
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

template <int BLOCK_N, bool TMA_STAGE, int PARTS, bool MCAST>
static int launch_mlp(const MlpParams& p, const void* w2, cudaStream_t s) {
  EncodeTiledFn encode = get_encode_tiled();
  V3D_REQUIRE(encode != nullptr, V3D_E_CUDA, "v3d_pe_mlp: cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)kHidden, (cuuint64_t)p.E};        // inner = k, outer = output channel
  const cuuint64_t gstride[1] = {(cuuint64_t)kHidden * 2};                  // bytes between rows of W2
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)BLOCK_N};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w2), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  V3D_REQUIRE(r == CUDA_SUCCESS, V3D_E_CUDA, "v3d_pe_mlp: cuTensorMapEncodeTiled failed (%d)", (int)r);
  // the token / output matrices as 2-D tensors: (128 rows x BLOCK_N columns) blocks for the L2 prefetch, (32 rows x 64
  // columns, 128-byte swizzle) blocks for the epilogue warps' staging tiles
  const cuuint64_t tdim[2] = {(cuuint64_t)p.E, (cuuint64_t)p.M};
  const cuuint64_t tstride[1] = {(cuuint64_t)p.E * 2};
  auto encode_io = [&](CUtensorMap* m, const void* base, cuuint32_t bw, cuuint32_t bh, CUtensorMapSwizzle sw) -> CUresult {
    const cuuint32_t tbox[2] = {bw, bh};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), tdim, tstride, tbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  const CUtensorMapSwizzle io_swizzle = (BLOCK_N / PARTS) * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tok_map, tok_io_map, out_io_map;
  memset(&tok_map, 0, sizeof(tok_map));
  memset(&tok_io_map, 0, sizeof(tok_io_map));
  memset(&out_io_map, 0, sizeof(out_io_map));
  if (p.tokens != nullptr) {
    V3D_REQUIRE(encode_io(&tok_map, p.tokens, BLOCK_N, kBlockM, CU_TENSOR_MAP_SWIZZLE_NONE) == CUDA_SUCCESS, V3D_E_CUDA,
                "v3d_pe_mlp: cuTensorMapEncodeTiled (token prefetch) failed");
    if (TMA_STAGE)
      V3D_REQUIRE(encode_io(&tok_io_map, p.tokens, BLOCK_N / PARTS, 32, io_swizzle) == CUDA_SUCCESS, V3D_E_CUDA,
                  "v3d_pe_mlp: cuTensorMapEncodeTiled (token tiles) failed");
  }
  if (TMA_STAGE)
    V3D_REQUIRE(encode_io(&out_io_map, p.out, BLOCK_N / PARTS, 32, io_swizzle) == CUDA_SUCCESS, V3D_E_CUDA,
                "v3d_pe_mlp: cuTensorMapEncodeTiled (output tiles) failed");
  using S = MlpSmem<BLOCK_N, TMA_STAGE, PARTS>;
  const int smem = S::total(p.E) + 1024;  // slack for the 1024-byte alignment of the dynamic window
  auto kern = k_pe_mlp<BLOCK_N, TMA_STAGE, PARTS, MCAST>;
  V3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  int grid = m_tiles < num_sms() ? m_tiles : num_sms();
  if (MCAST) {
    grid &= ~1;  // whole clusters of two
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)mlp_threads(PARTS));
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    V3D_CUDA(cudaLaunchKernelEx(&cfg, kern, map, tok_map, tok_io_map, out_io_map, p));
    return check_launch("k_pe_mlp");
  }
  kern<<<grid, mlp_threads(PARTS), smem, s>>>(map, tok_map, tok_io_map, out_io_map, p);
  return check_launch("k_pe_mlp");
}

}  // namespace v3d

using namespace v3d;

#ifdef V3D_MLP_TRACE
extern "C" int v3d_debug_mlp_trace(unsigned long long* out_host) {
  V3D_CUDA(cudaDeviceSynchronize());
  V3D_CUDA(cudaMemcpyFromSymbol(out_host, g_mlp_trace, sizeof(unsigned long long) * 8));
  return V3D_OK;
}
#endif

extern "C" int v3d_pe_mlp(const float* coords, int M, int E, const float* w1, const float* b1, const float* gamma,
                          const float* beta, float ln_eps, const void* w2_bf16, const float* b2, const void* tokens_in,
                          void* out, v3d_stream_t stream) {
  V3D_REQUIRE(coords && w1 && b1 && gamma && beta && w2_bf16 && b2 && out, V3D_E_ARG, "v3d_pe_mlp: null pointer");
  V3D_REQUIRE(M >= 0 && E > 0, V3D_E_ARG, "v3d_pe_mlp: bad shape M=%d E=%d", M, E);
  V3D_REQUIRE(E % 64 == 0, V3D_E_UNSUPPORTED, "v3d_pe_mlp: embedding size %d is not a multiple of 64", E);
  V3D_REQUIRE(E <= kMaxE, V3D_E_UNSUPPORTED, "v3d_pe_mlp: embedding size %d exceeds %d", E, kMaxE);
  V3D_REQUIRE(((reinterpret_cast<uintptr_t>(w2_bf16) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(tokens_in)) & 15) == 0,
              V3D_E_ALIGN, "v3d_pe_mlp: w2/out/tokens must be 16-byte aligned");
  if (M == 0) return V3D_OK;
  MlpParams p;
  p.coords = coords;
  p.w1 = w1;
  p.b1 = b1;
  p.gamma = gamma;
  p.beta = beta;
  p.b2 = b2;
  p.tokens = reinterpret_cast<const __nv_bfloat16*>(tokens_in);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.M = M;
  p.E = E;
  p.eps = ln_eps;
  cudaStream_t s = (cudaStream_t)stream;
  if (E % 128 == 0) {
    // staging tiles by TMA need two tiles per warp: fits next to the 8-stage W2 ring up to E = 5120
    if (V3D_MLP_TMA_STAGE && MlpSmem<128, true, V3D_MLP_PARTS>::total(E) + 1024 <= 227 * 1024) {
      // clusters of two need at least two M tiles
      if (V3D_MLP_MCAST && (M + kBlockM - 1) / kBlockM >= 2) return launch_mlp<128, true, V3D_MLP_PARTS, true>(p, w2_bf16, s);
      return launch_mlp<128, true, V3D_MLP_PARTS, false>(p, w2_bf16, s);
    }
    return launch_mlp<128, false, 2, false>(p, w2_bf16, s);
  }
  return launch_mlp<64, false, 2, false>(p, w2_bf16, s);
}
