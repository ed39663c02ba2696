// tcgen05 / TMEM attention for the two prefill shapes that carry the attention time of the path:
//   * EVA-ViT self-attention: head_dim 88 (contracted as 96), seq 257, q pre-scaled in fp16   eva_vit.py:118-148
//   * Llama causal self-attention: head_dim 128, key-padding mask, fp32 softmax              modeling_llama_mem.py:172-248
// (decode steps, the Q-Former's 32-query attention and every other shape stay on attention.cu.)
//
// One CTA = one (batch, head, 128-query tile); 192 threads:
//   warp 4     : TMA producer - Q once, then K/V tiles of 128 keys through a 2-stage ring (3-D tensor maps over the
//                packed q|k|v activations: {feature, token, batch}; 128B swizzle; two 64-column boxes per operand)
//   warp 5     : TMEM allocator + MMA issuer:  S = Q K^T (tcgen05.mma, fp32 S in TMEM columns 0..127), then
//                O += P V with P read from shared memory (K-major) and V used in place as an MN-major B operand
//                (no transpose of V anywhere), O in TMEM columns 128..255
//   warps 0..3 : softmax, one thread per query row: tcgen05.ld the S row, scale/mask, online max/sum in fp32, rescale
//                O in TMEM (tcgen05.ld/st) when the running max moved, write P (fp16) into the swizzled smem tile,
//                finally O / l -> fp16 -> global.
// head_dim 88: the second 64-column box of a head spills into the next head's columns; those 8 extra contraction
// columns (88..95) are zeroed in the Q tile (so whatever K holds there is multiplied by 0) and the 8 extra output
// columns of P V are simply not stored.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TQ = 128;    // query rows per CTA
constexpr int TKV = 64;    // keys per tile
constexpr int TC_THREADS = 192;
constexpr int BLK_BYTES = 128 * 128;   // one [128 rows x 64 fp16] swizzled block = 16 KB (Q, P)
constexpr int KBLK = TKV * 128;        // one [64 keys x 64 fp16] swizzled block = 8 KB (K, V)
// shared memory: Q 2 x 16 KB | K 2 stages x 2 x 8 KB | V 2 stages x 2 x 8 KB  = 96 KB -> 2 CTAs per SM.
// P_j (128 x 64 fp16 = 16 KB) is written over the K stage of tile j: K_j is dead once S_j = Q K_j^T has completed and
// the stage is only handed back to the TMA producer after P_j V_j has completed.

struct AttnTcParams {
  __half* o;
  const unsigned char* kmask;  // [B, Skv] 1 = attend, 0 = masked (additive mask_value); may be null
  long long o_sb, o_ss, o_sh;
  int B, H, Sq, Skv;
  float scale;       // applied to fp32 scores (1.0 when q_prescale)
  float q_prescale;  // if != 0: q <- fp16(q * q_prescale) before QK^T (eva_vit.py:128)
  float mask_value;
  int causal;
  // wide-tile kernel (attention_tc_wide_kernel) only: raw q/k/v for the rows / keys that do not fill a tensor-core tile
  const __half* q; const __half* k; const __half* v;
  long long q_sb, q_ss, k_sb, k_ss, v_sb, v_ss;   // batch / token strides in elements (heads contiguous: h * DH)
  int n_left;       // trailing keys (Skv % 128, <= LEFT_MAX) scored on CUDA cores by the softmax threads
  int n_tail;       // trailing query rows (Sq % 128, <= TAIL_MAX) handled by their own CUDA-core CTAs (blockIdx.y = Sq / 128)
};

DEVI void tma_load_3d(const CUtensorMap* desc, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "l"(TMA_EVICT_NORMAL)
      : "memory");
}
DEVI void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
DEVI void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major B operand (V tile [keys][64 features] blocks, 128B swizzle): LBO = distance between 64-feature blocks,
// SBO = distance between 8-key groups (cute::UMMA canonical ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
DEVI uint64_t make_sdesc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16_bmn(int M, int N) {   // A K-major, B MN-major
  return (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk `c` (8 fp16) of row `r` inside a swizzled [128 x 64] block
DEVI uint32_t sw_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

template <int DH>
__device__ __forceinline__ void attention_tail_rows(const AttnTcParams& p, uint8_t* smem, int h, int b);   // defined below

template <int DH>
__global__ void __launch_bounds__(TC_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  constexpr int DK = (DH == 88) ? 96 : 128;   // contraction length of Q K^T
  constexpr int DV = (DH == 88) ? 96 : 128;   // N of P V (columns >= DH are not stored)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // 2 blocks of 16 KB
  uint8_t* sK = sQ + 2 * BLK_BYTES;         // 2 stages x 2 blocks of 8 KB   (stage j&1 doubles as P_j)
  uint8_t* sV = sK + 4 * KBLK;              // 2 stages x 2 blocks of 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 4 * KBLK);
  uint64_t* q_full = bars + 0;
  uint64_t* q_ready = bars + 1;
  uint64_t* kv_full = bars + 2;    // [2]
  uint64_t* kv_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;     // [2]
  uint64_t* p_full = bars + 8;
  uint64_t* pv_done = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Trailing query rows (p.n_tail > 0: the ViT's 257th row) can get their own short CUDA-core CTAs in the SAME grid
  // (trailing-row mode 1) instead of a third 128-row tile per (frame, head) that holds one row and costs a full
  // TMA -> MMA -> softmax chain.  With tail CTAs the grid is re-indexed linearly: ALL tensor-core CTAs first (same order as without tail CTAs), the
  // short tail CTAs last, so they are dispatched after the last tensor-core CTA and fill the slots the ragged final wave
  // leaves free instead of delaying the first wave (interleaved per batch they cost more than they saved).
  int bx = (int)blockIdx.x, by = (int)blockIdx.y, bz = (int)blockIdx.z;
  int n_qt = (int)gridDim.y;
  if (p.n_tail > 0) {
    n_qt -= 1;
    const int H = (int)gridDim.x, L = bx + H * (by + (n_qt + 1) * bz), main_total = H * n_qt * (int)gridDim.z;
    if (L >= main_total) {
      const int T = L - main_total;
      pdl_launch_dependents();
      pdl_wait();
      attention_tail_rows<DH>(p, smem, T % H, T / H);
      return;
    }
    bx = L % H;
    by = (L / H) % n_qt;
    bz = L / (H * n_qt);
  }
  // grid = (heads, query tiles, batch): CTAs are dispatched x-fastest, so all heads of one query tile start together.
  // Causal: the LAST query tile has the most key tiles -> dispatch it first (longest-processing-time-first);
  // non-causal: ragged last tile (1 row of 257 in the ViT) goes last and fills the tail of the final wave.
  const int qt = p.causal ? (n_qt - 1 - by) : by;
  const int q0 = qt * TQ, h = bx, b = bz;
  const int kv_off = p.Skv - p.Sq;
  int kv_end = p.Skv;
  if (p.causal) {
    const int last = q0 + TQ + kv_off;
    kv_end = last < kv_end ? last : kv_end;
    if (kv_end < 1) kv_end = 1;
  }
  // Leftover keys (p.n_left > 0: the ViT's 257th key, Skv = 4 x 64 + 1) never become a tile of their own -- a fifth
  // TMA -> MMA -> softmax -> MMA round trip for ONE key cost a fifth of the CTA's serial chain.  The softmax threads score
  // them on CUDA cores during the Q fix-up pass (one DH-long dot product per row and key), open the online softmax with
  // them and add their weighted values to O in the epilogue (same scheme as attention_tc_wide_kernel).
  constexpr int LEFT_CAP = 4;                                  // == LEFT_MAX (defined below with the wide kernel)
  const int n_left = (DH == 88) ? p.n_left : 0;
  kv_end -= n_left;
  const int kv_mma = kv_end;
  __half* s_left_k = reinterpret_cast<__half*>(bars + 16);      // [LEFT_CAP][96] leftover keys
  __half* s_left_v = s_left_k + LEFT_CAP * 96;                  // [LEFT_CAP][96] leftover values
  const int n_tiles = (kv_end + TKV - 1) / TKV;
  const int col0 = h * DH;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_ready, 128);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 2 * TKV;   // S0: 0..63, S1: 64..127, O: 128..255

  if (warp == 4) {
    // ===================== TMA producer (warp 4: highest id on scheduler 0, see gemm_tcgen05.cu) =====================
    // warp-uniform loop, one elected lane issues (elect_one(), common.cuh)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * BLK_BYTES);
      tma_load_3d(&tmQ, q_full, sQ, col0, q0, b);
      tma_load_3d(&tmQ, q_full, sQ + BLK_BYTES, col0 + 64, q0, b);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&kv_empty[stage], phase ^ 1);
      uint8_t* k = sK + stage * 2 * KBLK;
      uint8_t* v = sV + stage * 2 * KBLK;
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[stage], 4 * KBLK);
        tma_load_3d(&tmK, &kv_full[stage], k, col0, j * TKV, b);
        tma_load_3d(&tmK, &kv_full[stage], k + KBLK, col0 + 64, j * TKV, b);
        tma_load_3d(&tmV, &kv_full[stage], v, col0, j * TKV, b);
        tma_load_3d(&tmV, &kv_full[stage], v + KBLK, col0 + 64, j * TKV, b);
      }
      __syncwarp();
      if (++stage == 2) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (warp 5: highest id on scheduler 1) =====================
    // Warp-uniform control flow; the elected lane issues tcgen05.mma / commit from uniform registers.  From a divergent
    // lane-0 region every UTCHMMA sits in a waterfall loop (~87 cycles per issue, profiles/r02_mma_pace.log) -- more than
    // the 32-64 tensor-pipe cycles of the N = 64 / 96 / 128 MMAs here, i.e. the kernel was issue-bound.
    constexpr uint32_t idesc_pv = make_idesc_f16_bmn(TQ, DV);
    mbar_wait(q_ready, 0);
    tc_fence_after();
    const uint32_t aq = smem_u32(sQ);
    // S_j = Q K_j^T into S buffer (j & 1); only ceil16(valid keys) columns are computed for a ragged last tile
    auto issue_qk = [&](int j) {
      const int stage = j & 1;
      mbar_wait(&kv_full[stage], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      int nk = kv_end - j * TKV;
      nk = nk > TKV ? TKV : ((nk + 15) & ~15);
      const uint32_t idesc_qk = make_idesc_f16(TQ, nk, false);
      const uint32_t bk = smem_u32(sK + stage * 2 * KBLK);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < DK / 16; ++ks) {
          const uint32_t aoff = (uint32_t)((ks >> 2) * BLK_BYTES + (ks & 3) * 32);
          const uint32_t boff = (uint32_t)((ks >> 2) * KBLK + (ks & 3) * 32);
          umma_f16_ss(tmem_S + (uint32_t)(stage * TKV), make_sdesc_k_sw128(aq + aoff), make_sdesc_k_sw128(bk + boff),
                      idesc_qk, ks > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[stage]);
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int stage = j & 1;
      if (j + 1 < n_tiles) issue_qk(j + 1);      // overlaps the softmax of tile j
      mbar_wait(p_full, (uint32_t)(j & 1));      // P_j written to smem, O rescaled
      tc_fence_after();
      int nk = kv_end - j * TKV;
      nk = nk > TKV ? TKV : ((nk + 15) & ~15);
      const uint32_t ap = smem_u32(sK + stage * 2 * KBLK);    // P_j lives in the consumed K stage
      const uint32_t bv = smem_u32(sV + stage * 2 * KBLK);
      if (elect_one()) {
#pragma unroll 4
        for (int ks = 0; ks < nk / 16; ++ks) {
          umma_f16_ss(tmem_O, make_sdesc_k_sw128(ap + ks * 32), make_sdesc_mn_sw128(bv + ks * 2048, KBLK), idesc_pv,
                      (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[stage]);   // K/V stage free once QK^T and PV of this tile are done
        umma_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax warps: thread = query row =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                // row inside the tile == TMEM lane
    const int qrow = q0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    // leftover keys / values -> smem (one 16-byte chunk per thread), zero padded to 96 columns
    if (DH == 88 && n_left > 0) {
      const int chunks = n_left * 12;
      for (int i = r; i < 2 * chunks; i += 128) {
        const int which = i >= chunks, ii = which ? i - chunks : i;
        const int li = ii / 12, c = ii % 12;
        const int key = kv_mma + li;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (c * 8 < DH) {
          const __half* src = which ? p.v + (long long)b * p.v_sb + (long long)key * p.v_ss + (long long)h * DH
                                    : p.k + (long long)b * p.k_sb + (long long)key * p.k_ss + (long long)h * DH;
          u = __ldg(reinterpret_cast<const uint4*>(src) + c);
        }
        *reinterpret_cast<uint4*>((which ? s_left_v : s_left_k) + li * 96 + c * 8) = u;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // ---- Q fix-up: optional fp16 pre-scale (ViT), zeroing of the padded contraction columns, leftover-key scores ----
    mbar_wait(q_full, 0);
    float s_left[LEFT_CAP];
#pragma unroll
    for (int i = 0; i < LEFT_CAP; ++i) s_left[i] = 0.f;
    if (DH == 88 || p.q_prescale != 0.f) {
#pragma unroll 1
      for (int c = 0; c < 16; ++c) {               // 16 chunks of 8 columns = 128 contraction columns
        uint4* ptr = reinterpret_cast<uint4*>(sQ + (c >> 3) * BLK_BYTES + sw_off(r, c & 7));
        if (c * 8 >= DH) {
          if (c * 8 < DK) *ptr = make_uint4(0, 0, 0, 0);
        } else if (p.q_prescale != 0.f || (DH == 88 && n_left > 0)) {
          uint4 u = *ptr;
          __half2* hh = reinterpret_cast<__half2*>(&u);
          if (p.q_prescale != 0.f) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(hh[i]);
              hh[i] = __floats2half2_rn(f.x * p.q_prescale, f.y * p.q_prescale);
            }
            *ptr = u;
          }
          if (DH == 88) {
#pragma unroll
            for (int li = 0; li < LEFT_CAP; ++li) {
              if (li < n_left) {
                const uint4 kk = *reinterpret_cast<const uint4*>(s_left_k + li * 96 + c * 8);   // broadcast read
                const __half2* kh = reinterpret_cast<const __half2*>(&kk);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 a = __half22float2(hh[i]), bb = __half22float2(kh[i]);
                  s_left[li] = fmaf(a.x, bb.x, fmaf(a.y, bb.y, s_left[li]));
                }
              }
            }
          }
        }
      }
      fence_proxy_async_smem();
    }
    mbar_arrive(q_ready);

    const float LOG2E = 1.4426950408889634f;
    const float sl2 = p.scale * LOG2E;
    const float mask_l2 = p.mask_value * LOG2E;
    const unsigned char* kmask = p.kmask ? p.kmask + (long long)b * p.Skv : nullptr;
    // the leftover keys open the online softmax: m = max of their scores, weights w_i = exp2(s_i - m)
    float m_run = -INFINITY, l_run = 0.f;
    float w_left[LEFT_CAP];
#pragma unroll
    for (int i = 0; i < LEFT_CAP; ++i) {
      s_left[i] *= sl2;
      if (DH == 88 && i < n_left) m_run = fmaxf(m_run, s_left[i]);
    }
#pragma unroll
    for (int i = 0; i < LEFT_CAP; ++i) {
      w_left[i] = (DH == 88 && i < n_left) ? exp2f(s_left[i] - m_run) : 0.f;
      l_run += w_left[i];
    }
    const bool warp_valid = (q0 + quad * 32) < p.Sq;   // warps whose 32 rows are all beyond Sq do no math at all
    for (int j = 0; j < n_tiles; ++j) {
      const int kbase = j * TKV;
      int nk = kv_end - kbase;                         // valid keys of this tile (for this CTA)
      nk = nk > TKV ? TKV : nk;
      const int n_chunks = (nk + 31) >> 5;             // 32-key chunks that contain at least one valid key
      mbar_wait(&s_full[j & 1], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      constexpr int NCH = TKV / 32;
      uint32_t s[NCH][32];
      float alpha = 1.f;
      if (warp_valid) {
        const uint32_t s_addr = tmem_S + (uint32_t)((j & 1) * TKV) + lane_addr;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          if (c < n_chunks) tmem_ld_32x32b_x32(s_addr + c * 32, s[c]);
        tmem_ld_wait();
        const bool need_mask = (kbase + TKV > p.Skv) || (p.causal && (kbase + TKV - 1 > q0 + kv_off)) || kmask != nullptr;
        float mx = -INFINITY;
        if (need_mask) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (c < n_chunks) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int key = kbase + c * 32 + i;
                float x = __uint_as_float(s[c][i]) * sl2;
                if (kmask != nullptr && key < p.Skv && kmask[key] == 0) x += mask_l2;
                if (key >= p.Skv || (p.causal && key > qrow + kv_off)) x = -INFINITY;
                s[c][i] = __float_as_uint(x);
                mx = fmaxf(mx, x);
              }
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (c < n_chunks) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[c][i]));
            }
          }
          mx *= sl2;                                   // sl2 > 0: max commutes with the scale
        }
        const float m_new = fmaxf(m_run, mx);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        alpha = exp2f(m_run - m_use);
        float sum = 0.f;
        // ---- P = exp2(S - m) -> fp16 pairs (kept in the low 16 words of each chunk) ----
        if (need_mask) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (c < n_chunks) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = exp2f(__uint_as_float(s[c][i]) - m_use);
                const float p1 = exp2f(__uint_as_float(s[c][i + 1]) - m_use);
                sum += p0 + p1;
                s[c][i >> 1] = pack_half2(p0, p1);
              }
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (c < n_chunks) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = exp2f(fmaf(__uint_as_float(s[c][i]), sl2, -m_use));
                const float p1 = exp2f(fmaf(__uint_as_float(s[c][i + 1]), sl2, -m_use));
                sum += p0 + p1;
                s[c][i >> 1] = pack_half2(p0, p1);
              }
            }
          }
        }
        l_run = l_run * alpha + sum;
        m_run = m_new;
        if (DH == 88) {
#pragma unroll
          for (int i = 0; i < LEFT_CAP; ++i) w_left[i] *= alpha;
        }
      }
      // ---- previous P V must be complete before O is rescaled and before sP is overwritten ----
      if (j > 0) {
        mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
        tc_fence_after();
        if (warp_valid && !__all_sync(0xffffffffu, alpha == 1.f)) {   // warp-uniform: skip when no row's max moved
#pragma unroll 1
          for (int c = 0; c < DV; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_O + lane_addr + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(tmem_O + lane_addr + c, o);
          }
          tmem_st_wait();
        }
      }
      // ---- P row -> swizzled smem (A operand of P V, K-major over keys); only the 16-key groups the MMA reads ----
      if (warp_valid) {
        uint8_t* sP = sK + (j & 1) * 2 * KBLK;         // the K stage of this tile: S_j = Q K_j^T has completed
        const int n_q = (((nk + 15) & ~15) + 7) >> 3;  // 8-key chunks to write
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = c * 4 + q;               // 8 chunks of 8 keys
            if (chunk < n_q)
              *reinterpret_cast<uint4*>(sP + sw_off(r, chunk)) =
                  make_uint4(s[c][4 * q], s[c][4 * q + 1], s[c][4 * q + 2], s[c][4 * q + 3]);
          }
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> fp16 -> global ----
    mbar_wait(pv_done, (uint32_t)((n_tiles - 1) & 1));
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    __half* orow = p.o + (long long)b * p.o_sb + (long long)qrow * p.o_ss + (long long)h * p.o_sh;
#pragma unroll 1
    for (int c = 0; c < DV; c += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_O + lane_addr + c, o);
      tmem_ld_wait();
      if (qrow < p.Sq) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int colq = c + q * 8;
          if (colq < DH) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[8 * q + i]);
            if (DH == 88) {
#pragma unroll
              for (int li = 0; li < LEFT_CAP; ++li) {
                if (li < n_left) {
                  const uint4 vv = *reinterpret_cast<const uint4*>(s_left_v + li * 96 + colq);
                  const __half2* vh = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 x = __half22float2(vh[i]);
                    f[2 * i] = fmaf(w_left[li], x.x, f[2 * i]);
                    f[2 * i + 1] = fmaf(w_left[li], x.y, f[2 * i + 1]);
                  }
                }
              }
            }
            uint4 w;
            w.x = pack_half2(f[0] * inv, f[1] * inv);
            w.y = pack_half2(f[2] * inv, f[3] * inv);
            w.z = pack_half2(f[4] * inv, f[5] * inv);
            w.w = pack_half2(f[6] * inv, f[7] * inv);
            *reinterpret_cast<uint4*>(orow + colq) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ================================================================================================
// Wide-tile variant for the ViT shape (non-causal, no key mask, sequence 257 = 2 x 128 + 1; eva_vit.py:128-145).
//
// Round 1's kernel walks the 257 keys in five 64-key tiles (the fifth holds ONE key) and gives the 257th query row a
// whole 128-row tile: 768 CTAs, each a chain of five serial TMA -> MMA -> softmax -> MMA round trips.  It ran at 41 us per
// layer with the tensor pipe ~10 % busy: latency-bound.  Here
//   * K/V tiles are 128 keys: two round trips per CTA; one K/V stage and one S buffer per CTA (96 KB smem, 224 TMEM
//     columns) so that two CTAs still share an SM and fill each other's waits;
//   * the keys left over after the last full tile (257 = 2 x 128 + 1) never become a tile: the softmax threads score them on
//     CUDA cores (one 88-long dot product per row and key) and carry their weights next to the TMEM accumulator;
//   * the query rows left over after the last full 128-row tile (one row per head and frame) go to their own small CTAs
//     (blockIdx.y == Sq / 128) that never touch the tensor core: q.K^T, softmax and P.V for <= TAIL_MAX rows.
// Softmax reads S from TMEM twice (row maximum, then exponentials) instead of holding 128 scores per thread in registers.
// ================================================================================================
constexpr int WKV = 128;                 // keys per tile
constexpr int WKBLK = WKV * 128;         // one [128 keys x 64 fp16] swizzled block = 16 KB
constexpr int LEFT_MAX = 4;
constexpr int TAIL_MAX = 4;
constexpr int WIDE_SMEM = 2 * BLK_BYTES + 2 * WKBLK + 2 * WKBLK + 4096 + 1024;   // Q | K (= P) | V | barriers + leftovers

// ---- CUDA-core path for the <= TAIL_MAX query rows after the last full 128-row tile (ViT: row 257 of 257) ----
// One CTA per (head, batch).  The head's K and V ([Skv][DH] fp16, 2 x 45 KB for the ViT) are staged in shared memory with
// one sweep of 16-byte cp.async copies (everything in flight at once: one L2 round trip instead of the ~30 dependent ones of
// the first version, profiles/r02_attention.md), then q.K^T (thread = key), softmax, P.V ((key group, 8-feature chunk) per
// thread) run out of shared memory.  Row stride = DH * 2 bytes when that is an odd number of 16-byte chunks (88 -> 11),
// else one chunk of padding: the 8 threads of an LDS.128 phase then hit 8 different bank groups.
template <int DH>
struct TailCfg {
  static constexpr int NCHUNK = DH / 8;                                   // 16-byte chunks per row
  static constexpr int RS = (NCHUNK & 1) ? DH * 2 : DH * 2 + 16;          // row stride in bytes
  __host__ __device__ static int smem_bytes(int Skv, int threads) {
    const int kg = threads / NCHUNK;
    return 2 * Skv * RS + (128 + ((Skv + 31) & ~31) + 64 + kg * DH) * 4 + 256;
  }
};
DEVI void cp_async16_tail(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
DEVI uint4 lds_u4(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
DEVI float lds_f1(uint32_t a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
  return r;
}
DEVI void sts_f1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

template <int DH>
__device__ __forceinline__ void attention_tail_rows(const AttnTcParams& p, uint8_t* smem, int h, int b) {
  constexpr int NCHUNK = TailCfg<DH>::NCHUNK, RS = TailCfg<DH>::RS;
  const int nt = (int)blockDim.x, nwarp = nt >> 5, t = threadIdx.x;
  const int KG = nt / NCHUNK;                            // key groups of the P V phase
  // shared-window addresses (the aligned `smem` pointer is generic: explicit ld/st.shared keep ptxas from emitting LD.E)
  const uint32_t s_k = smem_u32(smem);                   // [Skv][RS] keys
  const uint32_t s_v = s_k + (uint32_t)(p.Skv * RS);     // [Skv][RS] values
  const uint32_t s_q = s_v + (uint32_t)(p.Skv * RS);     // [128] floats: the (pre-scaled, fp16-rounded) query row
  const uint32_t s_p = s_q + 128 * 4;                    // [Skv] probabilities (unnormalised)
  const uint32_t s_red = s_p + (uint32_t)(((p.Skv + 31) & ~31) * 4);   // [2][32] reduction scratch
  const uint32_t s_part = s_red + 64 * 4;                // [KG][DH] partial outputs
  {
    const char* kb = reinterpret_cast<const char*>(p.k + (long long)b * p.k_sb + (long long)h * DH);
    const char* vb = reinterpret_cast<const char*>(p.v + (long long)b * p.v_sb + (long long)h * DH);
    const long long kss = p.k_ss * 2, vss = p.v_ss * 2;
    for (int i = t; i < p.Skv * NCHUNK; i += nt) {
      const int key = i / NCHUNK, c = i - key * NCHUNK;
      cp_async16_tail(s_k + (uint32_t)(key * RS + c * 16), kb + key * kss + c * 16);
      cp_async16_tail(s_v + (uint32_t)(key * RS + c * 16), vb + key * vss + c * 16);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const float sl2 = p.scale * 1.4426950408889634f;
  const int row0 = (p.Sq / TQ) * TQ;
  for (int r = 0; r < p.n_tail; ++r) {
    const int qrow = row0 + r;
    const __half* qp = p.q + (long long)b * p.q_sb + (long long)qrow * p.q_ss + (long long)h * DH;
    if (t < DH) {
      float x = __half2float(qp[t]);
      if (p.q_prescale != 0.f) x = __half2float(__float2half_rn(x * p.q_prescale));
      sts_f1(s_q + 4 * t, x);
    }
    if (r == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- scores: thread = key ----
    float mx = -INFINITY;
    for (int key = t; key < p.Skv; key += nt) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const uint4 u = lds_u4(s_k + (uint32_t)(key * RS + c * 16));
        const float4 qa = lds_f4(s_q + c * 32), qb = lds_f4(s_q + c * 32 + 16);   // broadcast
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
        const float2 f0 = __half22float2(hh[0]), f1 = __half22float2(hh[1]), f2 = __half22float2(hh[2]), f3 = __half22float2(hh[3]);
        a0 = fmaf(qa.x, f0.x, a0);
        a1 = fmaf(qa.y, f0.y, a1);
        a0 = fmaf(qa.z, f1.x, a0);
        a1 = fmaf(qa.w, f1.y, a1);
        a0 = fmaf(qb.x, f2.x, a0);
        a1 = fmaf(qb.y, f2.y, a1);
        a0 = fmaf(qb.z, f3.x, a0);
        a1 = fmaf(qb.w, f3.y, a1);
      }
      const float acc = (a0 + a1) * sl2;
      sts_f1(s_p + 4 * key, acc);
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    if ((t & 31) == 0) sts_f1(s_red + 4 * (t >> 5), mx);
    __syncthreads();
    mx = -INFINITY;
    for (int w = 0; w < nwarp; ++w) mx = fmaxf(mx, lds_f1(s_red + 4 * w));
    float sum = 0.f;
    for (int key = t; key < p.Skv; key += nt) {
      const float e = exp2f(lds_f1(s_p + 4 * key) - mx);
      sts_f1(s_p + 4 * key, e);
      sum += e;
    }
    sum = warp_sum(sum);
    if ((t & 31) == 0) sts_f1(s_red + 4 * (32 + (t >> 5)), sum);
    __syncthreads();
    float total = 0.f;
    for (int w = 0; w < nwarp; ++w) total += lds_f1(s_red + 4 * (32 + w));
    const float inv = 1.f / total;
    // ---- P V: thread = (key group, 8-feature chunk) ----
    if (t < KG * NCHUNK) {
      const int kg = t / NCHUNK, c = t - kg * NCHUNK;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int key = kg; key < p.Skv; key += KG) {
        const float w = lds_f1(s_p + 4 * key);
        const uint4 u = lds_u4(s_v + (uint32_t)(key * RS + c * 16));
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(hh[i]);
          acc[2 * i] = fmaf(w, f.x, acc[2 * i]);
          acc[2 * i + 1] = fmaf(w, f.y, acc[2 * i + 1]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sts_f1(s_part + 4 * (kg * DH + c * 8 + i), acc[i]);
    }
    __syncthreads();
    if (t < DH) {
      float a = 0.f;
      for (int kg = 0; kg < KG; ++kg) a += lds_f1(s_part + 4 * (kg * DH + t));
      p.o[(long long)b * p.o_sb + (long long)qrow * p.o_ss + (long long)h * p.o_sh + t] = __float2half_rn(a * inv);
    }
    __syncthreads();
  }
}

// The <= TAIL_MAX query rows after the last full 128-row tile as their own launch (grid = heads x 1 x batch): with it the
// tensor-core kernel runs on Sq - tail rows -- for the ViT (257 = 2 x 128 + 1) two query tiles per (frame, head) instead of
// three, the third of which held ONE row and cost a quarter of the kernel (39.0 -> 29.0 us, profiles/r02_attention.md).
constexpr int TAIL_THREADS = 288;   // one key per thread for the 257 keys of the ViT
template <int DH>
__global__ void __launch_bounds__(TAIL_THREADS) attention_tail_kernel(const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  pdl_launch_dependents();
  pdl_wait();
  attention_tail_rows<DH>(p, smem, (int)blockIdx.x, (int)blockIdx.z);
}

template <int DH>
__global__ void __launch_bounds__(TC_THREADS, 2)
attention_tc_wide_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  constexpr int DK = (DH == 88) ? 96 : 128;
  constexpr int DV = (DH == 88) ? 96 : 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.z;
  const int n_full_qt = p.Sq / TQ + ((p.Sq % TQ) > p.n_tail ? 1 : 0);   // tiles handled on the tensor core
  if ((int)blockIdx.y >= n_full_qt) {
    pdl_launch_dependents();
    pdl_wait();
    attention_tail_rows<DH>(p, smem, h, b);
    return;
  }
  uint8_t* sQ = smem;                       // 2 blocks of 16 KB
  uint8_t* sK = sQ + 2 * BLK_BYTES;         // 2 blocks of 16 KB (128 keys); doubles as P once S = Q K^T has completed
  uint8_t* sV = sK + 2 * WKBLK;             // 2 blocks of 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * WKBLK);
  uint64_t* q_full = bars + 0;
  uint64_t* q_ready = bars + 1;
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = bars + 3;
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* pv_done = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  __half* s_left_k = reinterpret_cast<__half*>(bars + 16);             // [LEFT_MAX][96] leftover keys
  __half* s_left_v = s_left_k + LEFT_MAX * 96;                         // [LEFT_MAX][96] leftover values

  const int q0 = (int)blockIdx.y * TQ;
  const int kv_mma = p.Skv - p.n_left;                                 // keys that go through the tensor core
  const int n_tiles = (kv_mma + WKV - 1) / WKV;
  const int col0 = h * DH;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_ready, 128);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + WKV;         // S: columns 0..127, O: 128..223

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * BLK_BYTES);
      tma_load_3d(&tmQ, q_full, sQ, col0, q0, b);
      tma_load_3d(&tmQ, q_full, sQ + BLK_BYTES, col0 + 64, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(kv_empty, (uint32_t)((j & 1) ^ 1));
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 4 * WKBLK);
        tma_load_3d(&tmK, kv_full, sK, col0, j * WKV, b);
        tma_load_3d(&tmK, kv_full, sK + WKBLK, col0 + 64, j * WKV, b);
        tma_load_3d(&tmV, kv_full, sV, col0, j * WKV, b);
        tma_load_3d(&tmV, kv_full, sV + WKBLK, col0 + 64, j * WKV, b);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_pv = make_idesc_f16_bmn(TQ, DV);
    mbar_wait(q_ready, 0);
    tc_fence_after();
    const uint32_t aq = smem_u32(sQ), bk = smem_u32(sK), bv = smem_u32(sV);
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(kv_full, (uint32_t)(j & 1));
      tc_fence_after();
      int nk = kv_mma - j * WKV;
      nk = nk > WKV ? WKV : ((nk + 15) & ~15);
      const uint32_t idesc_qk = make_idesc_f16(TQ, nk, false);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < DK / 16; ++ks) {
          const uint32_t aoff = (uint32_t)((ks >> 2) * BLK_BYTES + (ks & 3) * 32);
          const uint32_t boff = (uint32_t)((ks >> 2) * WKBLK + (ks & 3) * 32);
          umma_f16_ss(tmem_S, make_sdesc_k_sw128(aq + aoff), make_sdesc_k_sw128(bk + boff), idesc_qk, ks > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, (uint32_t)(j & 1));      // P_j written over the K tile, O rescaled
      tc_fence_after();
      if (elect_one()) {
#pragma unroll 4
        for (int ks = 0; ks < nk / 16; ++ks) {
          const uint32_t aoff = (uint32_t)((ks >> 2) * BLK_BYTES + (ks & 3) * 32);   // P: [128 rows x 64 keys] blocks
          umma_f16_ss(tmem_O, make_sdesc_k_sw128(bk + aoff), make_sdesc_mn_sw128(bv + ks * 2048, WKBLK), idesc_pv,
                      (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(kv_empty);                   // the K (= P) and V tiles are free once P V has completed
        umma_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax warps: thread = query row =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int qrow = q0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    // leftover keys / values -> smem (one 16-byte chunk per thread), zero padded to 96 columns
    {
      const int chunks = p.n_left * 12;
      for (int i = r; i < 2 * chunks; i += 128) {
        const int which = i >= chunks, ii = which ? i - chunks : i;
        const int li = ii / 12, c = ii % 12;
        const int key = kv_mma + li;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (c * 8 < DH) {
          const __half* src = which ? p.v + (long long)b * p.v_sb + (long long)key * p.v_ss + (long long)h * DH
                                    : p.k + (long long)b * p.k_sb + (long long)key * p.k_ss + (long long)h * DH;
          u = __ldg(reinterpret_cast<const uint4*>(src) + c);
        }
        *reinterpret_cast<uint4*>((which ? s_left_v : s_left_k) + li * 96 + c * 8) = u;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // ---- Q fix-up: optional fp16 pre-scale (ViT) and zeroing of the padded contraction columns; leftover scores ----
    mbar_wait(q_full, 0);
    float s_left[LEFT_MAX];
#pragma unroll
    for (int i = 0; i < LEFT_MAX; ++i) s_left[i] = 0.f;
#pragma unroll 1
    for (int c = 0; c < 16; ++c) {               // 16 chunks of 8 columns = 128 contraction columns
      uint4* ptr = reinterpret_cast<uint4*>(sQ + (c >> 3) * BLK_BYTES + sw_off(r, c & 7));
      if (c * 8 >= DH) {
        if (c * 8 < DK) *ptr = make_uint4(0, 0, 0, 0);
      } else {
        uint4 u = *ptr;
        __half2* hh = reinterpret_cast<__half2*>(&u);
        if (p.q_prescale != 0.f) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(hh[i]);
            hh[i] = __floats2half2_rn(f.x * p.q_prescale, f.y * p.q_prescale);
          }
          *ptr = u;
        }
#pragma unroll
        for (int li = 0; li < LEFT_MAX; ++li) {
          if (li < p.n_left) {
            const uint4 kk = *reinterpret_cast<const uint4*>(s_left_k + li * 96 + c * 8);   // broadcast read
            const __half2* kh = reinterpret_cast<const __half2*>(&kk);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a = __half22float2(hh[i]), bb = __half22float2(kh[i]);
              s_left[li] = fmaf(a.x, bb.x, fmaf(a.y, bb.y, s_left[li]));
            }
          }
        }
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(q_ready);

    const float LOG2E = 1.4426950408889634f;
    const float sl2 = p.scale * LOG2E;
    const bool warp_valid = (q0 + quad * 32) < p.Sq;
    // the leftover keys open the online softmax: m = max of their scores, weights w_i = exp2(s_i - m)
    float m_run = -INFINITY, l_run = 0.f;
    float w_left[LEFT_MAX];
#pragma unroll
    for (int i = 0; i < LEFT_MAX; ++i) {
      s_left[i] *= sl2;
      if (i < p.n_left) m_run = fmaxf(m_run, s_left[i]);
    }
#pragma unroll
    for (int i = 0; i < LEFT_MAX; ++i) {
      w_left[i] = (i < p.n_left) ? exp2f(s_left[i] - m_run) : 0.f;
      l_run += w_left[i];
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int kbase = j * WKV;
      int nk = kv_mma - kbase;
      nk = nk > WKV ? WKV : nk;
      const int n_chunks = (nk + 31) >> 5;
      mbar_wait(s_full, (uint32_t)(j & 1));
      tc_fence_after();
      float alpha = 1.f;
      if (warp_valid) {
        const uint32_t s_addr = tmem_S + lane_addr;
        // pass 1: row maximum
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t s[32];
          tmem_ld_32x32b_x32(s_addr + c * 32, s);
          tmem_ld_wait();
          const int valid = nk - c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < valid) mx = fmaxf(mx, __uint_as_float(s[i]));
        }
        mx *= sl2;                                     // sl2 > 0: max commutes with the scale
        const float m_new = fmaxf(m_run, mx);
        alpha = exp2f(m_run - m_new);                  // m_run = -inf (no leftover keys, first tile) -> 0
        // pass 2: P = exp2(S * sl2 - m) -> fp16 -> swizzled smem over the K tile (S_j = Q K_j^T has completed)
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t s[32];
          tmem_ld_32x32b_x32(s_addr + c * 32, s);
          tmem_ld_wait();
          const int valid = nk - c * 32;
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (i < valid) ? exp2f(fmaf(__uint_as_float(s[i]), sl2, -m_new)) : 0.f;
            const float p1 = (i + 1 < valid) ? exp2f(fmaf(__uint_as_float(s[i + 1]), sl2, -m_new)) : 0.f;
            sum += p0 + p1;
            pk[i >> 1] = pack_half2(p0, p1);
          }
          // the K tile may still be read by Q K^T of THIS tile?  No: s_full fired, i.e. the MMAs that read K have completed.
          uint8_t* blk = sK + (c >> 1) * BLK_BYTES;      // P block of 64 keys: chunk c -> 16-byte chunks (c & 1) * 4 .. + 3
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(blk + sw_off(r, (c & 1) * 4 + q)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        l_run = l_run * alpha + sum;
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < LEFT_MAX; ++i) w_left[i] *= alpha;
        fence_proxy_async_smem();
      }
      // ---- previous P V must be complete before O is rescaled (S and the K tile were released by kv_empty already) ----
      if (j > 0) {
        mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
        tc_fence_after();
        if (warp_valid && !__all_sync(0xffffffffu, alpha == 1.f)) {
#pragma unroll 1
          for (int c = 0; c < DV; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_O + lane_addr + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(tmem_O + lane_addr + c, o);
          }
          tmem_st_wait();
        }
      }
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: (O + sum_i w_i v_left_i) / l -> fp16 -> global ----
    mbar_wait(pv_done, (uint32_t)((n_tiles - 1) & 1));
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    __half* orow = p.o + (long long)b * p.o_sb + (long long)qrow * p.o_ss + (long long)h * p.o_sh;
#pragma unroll 1
    for (int c = 0; c < DV; c += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_O + lane_addr + c, o);
      tmem_ld_wait();
      if (qrow < p.Sq) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int colq = c + q * 8;
          if (colq < DH) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[8 * q + i]);
#pragma unroll
            for (int li = 0; li < LEFT_MAX; ++li) {
              if (li < p.n_left) {
                const uint4 vv = *reinterpret_cast<const uint4*>(s_left_v + li * 96 + colq);
                const __half2* vh = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 x = __half22float2(vh[i]);
                  f[2 * i] = fmaf(w_left[li], x.x, f[2 * i]);
                  f[2 * i + 1] = fmaf(w_left[li], x.y, f[2 * i + 1]);
                }
              }
            }
            uint4 w;
            w.x = pack_half2(f[0] * inv, f[1] * inv);
            w.y = pack_half2(f[2] * inv, f[3] * inv);
            w.z = pack_half2(f[4] * inv, f[5] * inv);
            w.w = pack_half2(f[6] * inv, f[7] * inv);
            *reinterpret_cast<uint4*>(orow + colq) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// {features (H*D), tokens (S, stride ss), batch (B, stride sb)} fp16, box {64, 128, 1}, 128B swizzle
int make_tmap3(CUtensorMap* tm, const void* ptr, int feat, int S, int B, long long ss, long long sb, int box_rows) {
  static PFN_encodeTiled enc = nullptr;
  if (enc == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      stllm_set_last_error("cuTensorMapEncodeTiled driver entry point not available");
      return STLLM_ERR_DRIVER;
    }
    enc = reinterpret_cast<PFN_encodeTiled>(sym);
  }
  cuuint64_t gdim[3] = {(cuuint64_t)feat, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t gstride[2] = {(cuuint64_t)ss * 2, (cuuint64_t)(B > 1 ? sb : ss * S) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    stllm_set_last_error("cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)r);
    return STLLM_ERR_DRIVER;
  }
  return STLLM_OK;
}

template <int DH>
int launch_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p,
              cudaStream_t stream) {
  constexpr int SMEM_TC = 2 * BLK_BYTES + 8 * KBLK + 128 + 2048 + 1024;   // Q | K | V | barriers | leftover keys/values | slack
  // with tail CTAs (p.n_tail > 0) every CTA of the grid gets the larger of the two footprints (still two CTAs per SM)
  const int smem_tail = p.n_tail > 0 ? TailCfg<DH>::smem_bytes(p.Skv, TC_THREADS) + 1024 : 0;
  const int smem = smem_tail > SMEM_TC ? smem_tail : SMEM_TC;
  auto kern = attention_tc_kernel<DH>;
  static int configured = 0;
  if (smem > configured) {
    STLLM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(p.H, p.n_tail > 0 ? p.Sq / TQ + 1 : (p.Sq + TQ - 1) / TQ, p.B);
  STLLM_CHECK_CUDA(stllm_launch(kern, grid, dim3(TC_THREADS), smem, stream, tq, tk, tv, p));
  return STLLM_OK;
}

template <int DH>
int launch_tc_wide(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p,
                   cudaStream_t stream) {
  auto kern = attention_tc_wide_kernel<DH>;
  static bool configured = false;
  if (!configured) {
    STLLM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WIDE_SMEM));
    configured = true;
  }
  const int n_full_qt = p.Sq / TQ + ((p.Sq % TQ) > p.n_tail ? 1 : 0);
  dim3 grid(p.H, n_full_qt + (p.n_tail > 0 ? 1 : 0), p.B);
  STLLM_CHECK_CUDA(stllm_launch(kern, grid, dim3(TC_THREADS), WIDE_SMEM, stream, tq, tk, tv, p));
  return STLLM_OK;
}

}  // namespace

static int g_attn_tc = -1;   // -1: read STLLM_ATTN_TC on first use (default on)
static int g_attn_wide = -1;  // -1: read STLLM_ATTN_WIDE on first use (default off)
static int g_attn_tail = -1;  // -1: read STLLM_ATTN_TAIL on first use (default 2: tail launch)
static int g_attn_last_launches = 1;
// kernels launched by the most recent attention call of this process (1, or 2 with the trailing-row launch): the host
// side counts launches per step from it (bench.py gpu_launches)
extern "C" int stllm_attention_last_launches(void) { return g_attn_last_launches; }
// experiments kept in the library (profiles/r02_attention.md): wide = 128-key tiles + leftover key / tail rows on CUDA cores
// in one kernel; tail = round-1 kernel on the full query tiles + a CUDA-core launch for the trailing rows.  -1 = environment.
extern "C" int stllm_set_attention_variant(int wide, int tail) {
  g_attn_wide = wide;
  g_attn_tail = tail;
  return STLLM_OK;
}
// debug / A-B switch: 1 = use the tcgen05 attention kernel where it applies, 0 = always the mma.sync kernel
extern "C" int stllm_set_attention_tc(int on) {
  g_attn_tc = on ? 1 : 0;
  return STLLM_OK;
}

// returns STLLM_ERR_UNSUPPORTED (without setting an error) when the shape/layout is not one this kernel handles
int stllm_attention_tc_try(const void* q, const void* k, const void* v, void* out, int B, int H, int Sq, int Skv, int D,
                           const long long* qs, const long long* ks, const long long* vs, const long long* os,
                           float scale, float q_prescale, int causal, const void* key_mask, float mask_value,
                           cudaStream_t stream) {
  g_attn_last_launches = 1;
  if (g_attn_tc < 0) {
    const char* e = getenv("STLLM_ATTN_TC");
    g_attn_tc = (e != nullptr && atoi(e) == 0) ? 0 : 1;
  }
  if (!g_attn_tc) return STLLM_ERR_UNSUPPORTED;
  if (!(D == 88 || D == 128) || Sq < 32) return STLLM_ERR_UNSUPPORTED;
  // heads contiguous in the feature dimension, 16-byte aligned token / batch strides
  if (qs[2] != D || ks[2] != D || vs[2] != D) return STLLM_ERR_UNSUPPORTED;
  if ((qs[1] % 8) || (ks[1] % 8) || (vs[1] % 8) || (qs[0] % 8) || (ks[0] % 8) || (vs[0] % 8) || (os[1] % 8) || (os[2] % 8) ||
      (os[0] % 8))
    return STLLM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) return STLLM_ERR_UNSUPPORTED;
  // wide-tile kernel: the ViT shape (non-causal, no key mask, head dim 88, at least one full 128-key tile)
  // Measured SLOWER than the 64-key kernel (49 vs 39 us on the ViT shape, profiles/r02_attention.md): kept as an experiment
  static const int wide_env = [] { const char* e = getenv("STLLM_ATTN_WIDE"); return (e != nullptr && atoi(e) == 1) ? 1 : 0; }();
  const int wide_ok = g_attn_wide >= 0 ? g_attn_wide : wide_env;
  const bool wide = wide_ok && D == 88 && !causal && key_mask == nullptr && Skv >= WKV && Skv <= 4096 && Sq >= TQ;
  CUtensorMap tq, tk, tv;
  int rc = make_tmap3(&tq, q, H * D, Sq, B, qs[1], qs[0], TQ);
  if (rc != STLLM_OK) return rc;
  if (wide) {
    rc = make_tmap3(&tk, k, H * D, Skv, B, ks[1], ks[0], WKV);
    if (rc != STLLM_OK) return rc;
    rc = make_tmap3(&tv, v, H * D, Skv, B, vs[1], vs[0], WKV);
    if (rc != STLLM_OK) return rc;
    AttnTcParams p;
    p.o = (__half*)out;
    p.kmask = nullptr;
    p.o_sb = os[0]; p.o_ss = os[1]; p.o_sh = os[2];
    p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv;
    p.scale = scale; p.q_prescale = q_prescale; p.mask_value = mask_value; p.causal = 0;
    p.q = (const __half*)q; p.k = (const __half*)k; p.v = (const __half*)v;
    p.q_sb = qs[0]; p.q_ss = qs[1]; p.k_sb = ks[0]; p.k_ss = ks[1]; p.v_sb = vs[0]; p.v_ss = vs[1];
    p.n_left = (Skv % WKV) <= LEFT_MAX ? (Skv % WKV) : 0;
    p.n_tail = ((Sq % TQ) <= TAIL_MAX && TailCfg<88>::smem_bytes(Skv, TC_THREADS) + 1024 <= WIDE_SMEM) ? (Sq % TQ) : 0;
    return launch_tc_wide<88>(tq, tk, tv, p, stream);
  }
  rc = make_tmap3(&tk, k, H * D, Skv, B, ks[1], ks[0], TKV);
  if (rc != STLLM_OK) return rc;
  rc = make_tmap3(&tv, v, H * D, Skv, B, vs[1], vs[0], TKV);
  if (rc != STLLM_OK) return rc;
  AttnTcParams p;
  p.o = (__half*)out;
  p.kmask = (const unsigned char*)key_mask;
  p.o_sb = os[0]; p.o_ss = os[1]; p.o_sh = os[2];
  p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv;
  p.scale = scale; p.q_prescale = q_prescale; p.mask_value = mask_value; p.causal = causal;
  p.q = (const __half*)q; p.k = (const __half*)k; p.v = (const __half*)v;
  p.q_sb = qs[0]; p.q_ss = qs[1]; p.k_sb = ks[0]; p.k_ss = ks[1]; p.v_sb = vs[0]; p.v_ss = vs[1];
  p.n_left = p.n_tail = 0;
  // ViT shape (257 keys = 4 x 64 + 1): the key(s) after the last full 64-key tile are scored on CUDA cores by the softmax
  // threads instead of costing a fifth tile round trip (STLLM_ATTN_LEFT=0 restores the fifth tile)
  static const int left_env = [] { const char* e = getenv("STLLM_ATTN_LEFT"); return e != nullptr ? atoi(e) : 1; }();
  if (left_env && D == 88 && !causal && key_mask == nullptr && Skv > TKV && (Skv % TKV) >= 1 && (Skv % TKV) <= LEFT_MAX)
    p.n_left = Skv % TKV;
  // ViT shape (257 = 2 x 128 + 1): the query rows after the last full tile go to CUDA-core tail CTAs (K/V of the head staged
  // in shared memory by one cp.async sweep), the tensor-core kernel runs on the full tiles only.  STLLM_ATTN_TAIL /
  // stllm_set_attention_variant: 0 = third 128-row tile (round 1), 1 = tail CTAs inside the same grid, 2 (default) = a
  // second small launch.  Measured in the step on one box, three alternating runs each (profiles/r02_attention.md):
  // 806 / 808 / 816 frames/s; standalone 39.3 / 40.9 / 35.2 us per layer.
  static const int tail_env = [] { const char* e = getenv("STLLM_ATTN_TAIL"); return e != nullptr ? atoi(e) : 2; }();
  const int tail_ok = g_attn_tail >= 0 ? g_attn_tail : tail_env;
  const int tail = Sq % TQ;
  // mode 1: tail CTAs inside the tensor-core kernel's grid (one launch); mode 2: a separate tail launch behind it (A/B)
  const bool tail_fits = TailCfg<88>::smem_bytes(Skv, TC_THREADS) + 1024 <= 112 * 1024;   // two CTAs per SM must still fit
  if (tail_ok && D == 88 && !causal && key_mask == nullptr && Sq > TQ && tail >= 1 && tail <= TAIL_MAX && tail_fits) {
    if (tail_ok == 2) {
      AttnTcParams pm = p;
      pm.Sq = Sq - tail;
      rc = launch_tc<88>(tq, tk, tv, pm, stream);
      if (rc != STLLM_OK) return rc;
      p.n_tail = tail;
      auto kern = attention_tail_kernel<88>;
      const int smem = TailCfg<88>::smem_bytes(Skv, TAIL_THREADS) + 256;
      static int configured_smem = 0;
      if (smem > 48 * 1024 && smem > configured_smem) {
        STLLM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_smem = smem;
      }
      STLLM_CHECK_CUDA(stllm_launch(kern, dim3(H, 1, B), dim3(TAIL_THREADS), smem, stream, p));
      g_attn_last_launches = 2;
      return STLLM_OK;
    }
    p.n_tail = tail;
    return launch_tc<88>(tq, tk, tv, p, stream);
  }
  if (D == 88) return launch_tc<88>(tq, tk, tv, p, stream);
  return launch_tc<128>(tq, tk, tv, p, stream);
}
