// tcgen05 / TMEM / TMA GEMM for sm_100a:   C[M,N] = epilogue(A[M,K] * W[N,K]^T + bias) (+ residual)
//
// This is the one dense-contraction kernel of the hot path. It replaces every nn.Linear / Conv2d-as-GEMM
// the reference dispatches to cuBLAS/cuDNN: eva_vit.py:124 (qkv), :146 (proj), :55-59 (fc1/GELU/fc2),
// :196 (patch-embed conv as im2col GEMM), Qformer.py query/key/value/dense, st_llm.py:368 (llama_proj),
// the HF Llama q/k/v/o/gate/up/down projections and lm_head (st_llm.py:108,122).
//
// Design (B200-first, persistent, warp-specialised):
//   * grid = min(#tiles, #SMs) persistent CTAs, 1 CTA/SM, 320 threads:
//       warp 0      : TMA producer   (one elected lane; A and W tiles, 128B-swizzled, mbarrier complete_tx)
//       warp 1      : TMEM allocator + MMA issuer (one lane issues tcgen05.mma kind::f16, fp32 accumulate)
//       warps 2..9  : epilogue (tcgen05.ld -> bias / GELU / SwiGLU / residual -> fp16 -> global)
//   * smem ring of STAGES x {A 128x64, W BNx64} fp16 tiles; TMEM holds 2 accumulator stages (2 x BN
//     columns) so the epilogue of tile i overlaps the mainloop of tile i+1.
//   * Both operands are K-major (activations row-major, nn.Linear weights [N,K]) so no transposes exist
//     anywhere on the path; ragged M/N/K edges are zero-filled by TMA and predicated in the epilogue.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 B = one swizzle-128B row
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;

enum Epilogue : int { EPI_NONE = 0, EPI_GELU = 1, EPI_SWIGLU = 2, EPI_RELU = 3 };

struct GemmParams {
  void* C;
  const __half* bias;      // [N] or null
  const __half* residual;  // [M, ldr] or null (added after the activation)
  long long ldc, ldr;
  int M, N, K;
  int m_tiles, n_tiles;
  int out_f32;  // 0: fp16 output, 1: fp32 output
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
};

DEVI float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
DEVI float silu(float x) { return x / (1.0f + __expf(-x)); }

// One 32-column chunk of one accumulator row: bias / activation / SwiGLU / residual / convert / store.
// bias and residual are fetched BEFORE the TMEM load is waited on, so their global-load latency overlaps it.
template <int EPI>
DEVI void epilogue_chunk(const GemmParams& p, uint32_t taddr, int row, bool row_ok, int col0) {
  constexpr int NC = (EPI == EPI_SWIGLU) ? 16 : 32;  // output columns produced by this chunk
  const int ocol0 = (EPI == EPI_SWIGLU) ? (col0 >> 1) : col0;
  const int n_out = (EPI == EPI_SWIGLU) ? (p.N >> 1) : p.N;
  const bool active = row_ok && col0 < p.N;
  const bool full_chunk = (col0 + 32 <= p.N);
  const bool vec_ok = (ocol0 + NC <= n_out);
  uint32_t r[32];
  tmem_ld_32x32b_x32(taddr, r);
  uint4 bq[4], rq[NC / 8];
  const bool bias_vec = active && p.bias != nullptr && full_chunk;
  const __half* rp = active && p.residual != nullptr ? p.residual + (long long)row * p.ldr + ocol0 : nullptr;
  const bool res_vec = rp != nullptr && vec_ok && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0);
  if (bias_vec) {
#pragma unroll
    for (int q = 0; q < 4; ++q) bq[q] = __ldg(reinterpret_cast<const uint4*>(p.bias + col0) + q);
  }
  if (res_vec) {
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) rq[q] = __ldg(reinterpret_cast<const uint4*>(rp) + q);
  }
  tmem_ld_wait();
  if (!active) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (p.bias != nullptr) {
    if (bias_vec) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __half2* bh = reinterpret_cast<const __half2*>(&bq[q]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(bh[j]);
          v[q * 8 + 2 * j] += f.x;
          v[q * 8 + 2 * j + 1] += f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] += __half2float(p.bias[col0 + j]);
    }
  }
  if (EPI == EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  }
  if (EPI == EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (EPI == EPI_SWIGLU) {
    // interleaved weights: accumulator column 2j = gate_j, 2j+1 = up_j -> output column j
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = silu(v[2 * j]) * v[2 * j + 1];
  }
  if (rp != nullptr) {
    if (res_vec) {
#pragma unroll
      for (int q = 0; q < NC / 8; ++q) {
        const __half2* bh = reinterpret_cast<const __half2*>(&rq[q]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(bh[j]);
          v[q * 8 + 2 * j] += f.x;
          v[q * 8 + 2 * j + 1] += f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (ocol0 + j < n_out) v[j] += __half2float(rp[j]);
    }
  }
  if (p.out_f32) {
    float* cp = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + ocol0;
    if (vec_ok && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
      for (int q = 0; q < NC / 4; ++q)
        reinterpret_cast<float4*>(cp)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (ocol0 + j < n_out) cp[j] = v[j];
    }
  } else {
    __half* cp = reinterpret_cast<__half*>(p.C) + (long long)row * p.ldc + ocol0;
    if (vec_ok && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
      for (int q = 0; q < NC / 8; ++q) {
        uint4 o;
        o.x = pack_half2(v[8 * q + 0], v[8 * q + 1]);
        o.y = pack_half2(v[8 * q + 2], v[8 * q + 3]);
        o.z = pack_half2(v[8 * q + 4], v[8 * q + 5]);
        o.w = pack_half2(v[8 * q + 6], v[8 * q + 7]);
        reinterpret_cast<uint4*>(cp)[q] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (ocol0 + j < n_out) cp[j] = __float2half_rn(v[j]);
    }
  }
}

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const GemmParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // swizzle-128B tiles need 1024-B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t % p.m_tiles, n_blk = t / p.m_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, m_blk * BM, TMA_EVICT_NORMAL);
          tma_load_2d(&tmB, &full_bar[stage], sb, kb * BK, n_blk * BN, TMA_EVICT_NORMAL);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da = make_sdesc_k_sw128(sa + kk * 32);
            const uint64_t db = make_sdesc_k_sw128(sb + kk * 32);
            umma_f16_ss(d_tmem, da, db, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int e = warp - 2;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int half_idx = e >> 2;         // which half of the BN columns
    constexpr int COLS_PER_WARP = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t % p.m_tiles, n_blk = t / p.m_tiles;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half_idx * COLS_PER_WARP);
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 32) {
        __syncwarp();
        const int col0 = n_blk * BN + half_idx * COLS_PER_WARP + c;  // first accumulator column of the chunk
        epilogue_chunk<EPI>(p, t_row + c, row, row_ok, col0);
        if (c + 32 >= COLS_PER_WARP) {
          // all TMEM reads of this warp for this tile are done -> release the accumulator stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ==============================================================================================
// CTA-pair variant (cta_group::2): a cluster of 2 CTAs computes a 256 x 256 output tile.
//   * each CTA TMA-loads its own 128 rows of A and its own 128 rows (= half of the N tile) of W per k-block:
//     32 KB per CTA per 512-cycle k-block = 64 B/clk of L2->SMEM ingest instead of 96 B/clk for the 128x256
//     one-CTA tile (ncu: the one-CTA kernel sits at 0.67 tensor-pipe utilisation = 64/96).
//   * the LEADER CTA's elected thread issues tcgen05.mma.cta_group::2 (M=256, N=256); the hardware reads A/B
//     from both CTAs' shared memory and writes 128 accumulator rows into each CTA's TMEM.
//   * barriers: full[s] lives in the leader (both CTAs' TMA complete_tx on it); empty[s] and tmem_full[a] are
//     multicast-committed to both CTAs; tmem_empty[a] lives in the leader and collects both CTAs' epilogue warps.
// ==============================================================================================
constexpr int BN2 = 256;
constexpr int STAGES2 = 6;
struct SmemLayout2 {
  static constexpr int A_BYTES = BM * BK * 2;          // 16 KB: this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN2 / 2) * BK * 2;   // 16 KB: this CTA's half of the W tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES2 * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_f16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const GemmParams p) {
  using L = SmemLayout2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES2;
  uint64_t* tmem_full_bar = empty_bar + STAGES2;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool is_leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = p.m_tiles * p.n_tiles;   // 256 x 256 pair tiles
  const int k_blocks = (p.K + BK - 1) / BK;
  constexpr uint32_t TMEM_COLS = 512;            // 2 accumulator stages x 256 fp32 columns

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full_bar[s], 1);     // leader's copy is the one in use: 1 arrival (leader producer) + tx bytes
      mbar_init(&empty_bar[s], 1);    // one multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);                    // one multicast tcgen05.commit
      mbar_init(&tmem_empty_bar[a], 2 * NUM_EPI_WARPS);   // leader's copy: epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_base_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();   // barriers of BOTH CTAs initialised before any remote arrive / TMA complete_tx
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int m_blk = t % p.m_tiles, n_blk = t / p.m_tiles;
        const int row_a = m_blk * 256 + (int)cta_rank * 128;
        const int row_b = n_blk * BN2 + (int)cta_rank * (BN2 / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);  // bytes of both CTAs
          tma_load_2d_2sm(&tmA, &full_bar[stage], sa, kb * BK, row_a, TMA_EVICT_NORMAL);
          tma_load_2d_2sm(&tmB, &full_bar[stage], sb, kb * BK, row_b, TMA_EVICT_NORMAL);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (is_leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, BN2, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da = make_sdesc_k_sw128(sa + kk * 32);
            const uint64_t db = make_sdesc_k_sw128(sb + kk * 32);
            umma_f16_ss_2sm(d_tmem, da, db, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage], 3);   // frees this smem slot in BOTH CTAs
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tmem_full_bar[acc], 3);   // accumulators of both CTAs complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, 128 rows each) =====================
    const int e = warp - 2;
    const int quad = warp & 3;
    const int half_idx = e >> 2;
    constexpr int COLS_PER_WARP = BN2 / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int m_blk = t % p.m_tiles, n_blk = t / p.m_tiles;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * 256 + (int)cta_rank * 128 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN2 + half_idx * COLS_PER_WARP);
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 32) {
        __syncwarp();
        const int col0 = n_blk * BN2 + half_idx * COLS_PER_WARP + c;
        epilogue_chunk<EPI>(p, t_row + c, row, row_ok, col0);
        if (c + 32 >= COLS_PER_WARP) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0));  // leader's barrier
        }
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's smem/TMEM/barriers stay alive until the leader's last MMA and commit landed
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(sym);
  }
  return fn;
}

// 2-D fp16 row-major [rows, cols] tensor, row stride ld elements; box = {64 cols, box_rows}; 128B swizzle
int make_tmap(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    stllm_set_last_error("cuTensorMapEncodeTiled driver entry point not available");
    return STLLM_ERR_DRIVER;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    stllm_set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (ptr %p rows %lld cols %lld ld %lld)", (int)r, ptr,
                         rows, cols, ld);
    return STLLM_ERR_DRIVER;
  }
  return STLLM_OK;
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int STAGES, int EPI>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = gemm_f16_tcgen05_kernel<BN, STAGES, EPI>;
  static bool configured = false;
  if (!configured) {
    STLLM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, NUM_THREADS, L::TOTAL, stream>>>(tmA, tmB, p);
  STLLM_CHECK_CUDA(cudaGetLastError());
  return STLLM_OK;
}

template <int BN, int STAGES>
int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case EPI_NONE: return launch_gemm<BN, STAGES, EPI_NONE>(tmA, tmB, p, s);
    case EPI_GELU: return launch_gemm<BN, STAGES, EPI_GELU>(tmA, tmB, p, s);
    case EPI_SWIGLU: return launch_gemm<BN, STAGES, EPI_SWIGLU>(tmA, tmB, p, s);
    case EPI_RELU: return launch_gemm<BN, STAGES, EPI_RELU>(tmA, tmB, p, s);
    default: stllm_set_last_error("unknown epilogue %d", epi); return STLLM_ERR_BAD_ARG;
  }
}

template <int EPI>
int launch_gemm_2cta(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_f16_tcgen05_2cta_kernel<EPI>;
  static bool configured = false;
  if (!configured) {
    STLLM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout2::TOTAL));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  int clusters = num_sms() / 2;
  if (tiles < clusters) clusters = tiles;
  kern<<<2 * clusters, NUM_THREADS, SmemLayout2::TOTAL, stream>>>(tmA, tmB, p);
  STLLM_CHECK_CUDA(cudaGetLastError());
  return STLLM_OK;
}

int dispatch_epi_2cta(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case EPI_NONE: return launch_gemm_2cta<EPI_NONE>(tmA, tmB, p, s);
    case EPI_GELU: return launch_gemm_2cta<EPI_GELU>(tmA, tmB, p, s);
    case EPI_SWIGLU: return launch_gemm_2cta<EPI_SWIGLU>(tmA, tmB, p, s);
    case EPI_RELU: return launch_gemm_2cta<EPI_RELU>(tmA, tmB, p, s);
    default: stllm_set_last_error("unknown epilogue %d", epi); return STLLM_ERR_BAD_ARG;
  }
}

}  // namespace

extern "C" int stllm_gemm_f16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc,
                              int M, int N, int K, const void* bias, const void* residual, long long ldr,
                              int epilogue, int out_f32, int tile_n, void* stream) {
  STLLM_CHECK_ARG(A != nullptr && W != nullptr && C != nullptr);
  STLLM_CHECK_ARG(M > 0 && N > 0 && K > 0);
  STLLM_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0);  // TMA: 16-byte global strides
  STLLM_CHECK_ARG(lda >= K && ldw >= K);
  STLLM_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  STLLM_CHECK_ARG(epilogue != EPI_SWIGLU || (N % 2 == 0));
  STLLM_CHECK_ARG(ldc >= (epilogue == EPI_SWIGLU ? N / 2 : N));
  STLLM_CHECK_ARG(residual == nullptr || ldr >= (epilogue == EPI_SWIGLU ? N / 2 : N));
  int bn = tile_n;
  if (bn == 0) {
    // heuristic: wide tiles when there are enough of them to fill the machine, else narrower tiles
    const long long m_tiles = (M + BM - 1) / BM;
    const long long t256 = m_tiles * ((N + 255) / 256);
    bn = (t256 >= num_sms()) ? 256 : 128;
  }
  STLLM_CHECK_ARG(bn == 128 || bn == 256 || bn == 512);   // 512 = CTA-pair kernel (256 x 256 tile per 2 SMs)
  CUtensorMap tmA, tmB;
  int rc = make_tmap(&tmA, A, M, K, lda, BM);
  if (rc != STLLM_OK) return rc;
  rc = make_tmap(&tmB, W, N, K, ldw, bn == 512 ? 128 : bn);
  if (rc != STLLM_OK) return rc;
  GemmParams p;
  p.C = C;
  p.bias = reinterpret_cast<const __half*>(bias);
  p.residual = reinterpret_cast<const __half*>(residual);
  p.ldc = ldc;
  p.ldr = ldr;
  p.M = M;
  p.N = N;
  p.K = K;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + bn - 1) / bn;
  p.out_f32 = out_f32;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (bn == 512) {
    p.m_tiles = (M + 255) / 256;
    p.n_tiles = (N + 255) / 256;
    return dispatch_epi_2cta(epilogue, tmA, tmB, p, s);
  }
  if (bn == 256) return dispatch_epi<256, 4>(epilogue, tmA, tmB, p, s);
  return dispatch_epi<128, 6>(epilogue, tmA, tmB, p, s);
}
