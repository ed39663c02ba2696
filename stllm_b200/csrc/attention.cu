// Fused (flash-style) multi-head attention, fp16 in / fp32 softmax / fp16 out, for sm_100a.
//
// One kernel covers the four attention shapes of the hot path (reference call sites):
//   * EVA-ViT self-attention, 16 heads x 88, seq 257, non-causal, q pre-scaled in fp16   eva_vit.py:118-148
//   * Q-Former self-attention (12 x 64, seq 32 + L_txt, additive -10000 key mask)         Qformer.py:169-275
//   * Q-Former cross-attention (32 queries x 257 image tokens)                            Qformer.py:186-193
//   * Llama causal self-attention (32 x 128, fp32 softmax, key-padding mask)              modeling_llama_mem.py:172-248
// The reference materialises the [B,H,S,S] score matrix in HBM (eva_vit.py:129,143); here scores never
// leave registers: S = Q K^T with mma.sync m16n8k16 (fp32 accumulate), online softmax in fp32, P V with the
// accumulator-to-A-fragment register reuse. K/V tiles are double-buffered in shared memory with cp.async.
// (Attention is ~2 % of the path's FLOPs - SURVEY 8d; the 98 % that is GEMM runs on tcgen05, gemm_tcgen05.cu.)
#include "common.cuh"

namespace {

constexpr int BKV = 64;   // keys per smem tile
// BQ (template): query rows per CTA = 16 per warp; 64 everywhere (see the note at the dispatch).

struct AttnParams {
  const __half* q;
  const __half* k;
  const __half* v;
  __half* o;
  const unsigned char* kmask;  // [B, Skv] 1 = attend, 0 = masked (additive mask_value); may be null
  long long q_sb, q_ss, q_sh, k_sb, k_ss, k_sh, v_sb, v_ss, v_sh, o_sb, o_ss, o_sh;
  int B, H, Sq, Skv;
  float scale;       // applied to fp32 scores (1.0 when q_prescale)
  float q_prescale;  // if != 0: q <- fp16(q * q_prescale) before QK^T (eva_vit.py:128)
  float mask_value;  // additive value for masked keys (-10000 Q-Former, -65504 Llama)
  int causal;
};

DEVI void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
DEVI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
DEVI void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

DEVI void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
DEVI void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
DEVI void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// DH: real head dim (multiple of 8); DP: padded to a multiple of 16
template <int DH, int DP, int BQ>
__global__ void __launch_bounds__(BQ * 2)
attention_f16_kernel(const AttnParams p) {
  constexpr int ATT_THREADS = BQ * 2;
  constexpr int LD = DP + 8;            // smem row stride in halves (conflict-free ldmatrix)
  constexpr int CHUNKS = DH / 8;        // 16-byte chunks per row actually loaded
  constexpr int KSTEPS = DP / 16;       // k-steps of QK^T
  constexpr int NT_O = DP / 8;          // n-tiles of the output
  extern __shared__ __align__(16) unsigned char att_smem[];
  __half* sQ = reinterpret_cast<__half*>(att_smem);
  __half* sK = sQ + BQ * LD;            // [2][BKV][LD]
  __half* sV = sK + 2 * BKV * LD;       // [2][BKV][LD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  pdl_launch_dependents();
  const __half* qg = p.q + b * p.q_sb + h * p.q_sh;
  const __half* kg = p.k + b * p.k_sb + h * p.k_sh;
  const __half* vg = p.v + b * p.v_sb + h * p.v_sh;
  const int kv_off = p.Skv - p.Sq;  // causal: key j visible to query i iff j <= i + kv_off

  // zero the padding columns [DH, DP) (never touched by cp.async) and the +8 tail
  if (DP + 8 > DH) {
    constexpr int PADW = LD - DH;
    for (int i = tid; i < (BQ + 4 * BKV) * PADW; i += ATT_THREADS) {
      const int r = i / PADW, c = DH + i % PADW;
      sQ[r * LD + c] = __float2half(0.f);  // sQ, sK, sV are contiguous with the same stride
    }
  }

  int kv_end = p.Skv;
  if (p.causal) {
    const int last = q0 + BQ - 1 + kv_off + 1;
    kv_end = last < kv_end ? last : kv_end;
    if (kv_end < 1) kv_end = 1;
  }
  const int n_tiles = (kv_end + BKV - 1) / BKV;

  auto load_q = [&]() {
    for (int i = tid; i < BQ * CHUNKS; i += ATT_THREADS) {
      const int r = i / CHUNKS, c = i % CHUNKS;
      const bool ok = (q0 + r) < p.Sq;
      cp_async16(sQ + r * LD + c * 8, qg + (long long)(ok ? q0 + r : 0) * p.q_ss + c * 8, ok);
    }
  };
  auto load_kv = [&](int t, int buf) {
    const int k0 = t * BKV;
    __half* dk = sK + buf * BKV * LD;
    __half* dv = sV + buf * BKV * LD;
    for (int i = tid; i < BKV * CHUNKS; i += ATT_THREADS) {
      const int r = i / CHUNKS, c = i % CHUNKS;
      const bool ok = (k0 + r) < p.Skv;
      const long long rr = ok ? k0 + r : 0;
      cp_async16(dk + r * LD + c * 8, kg + rr * p.k_ss + c * 8, ok);
      cp_async16(dv + r * LD + c * 8, vg + rr * p.v_ss + c * 8, ok);
    }
  };

  pdl_wait();   // q/k/v come from the previous kernel
  load_q();
  cp_async_commit();
  load_kv(0, 0);
  cp_async_commit();

  cp_async_wait<1>();  // Q landed
  __syncthreads();
  if (p.q_prescale != 0.f) {
    for (int i = tid; i < BQ * DH; i += ATT_THREADS) {
      const int r = i / DH, c = i % DH;
      sQ[r * LD + c] = __float2half_rn(__half2float(sQ[r * LD + c]) * p.q_prescale);
    }
    __syncthreads();
  }

  // Q fragments (A operand), kept in registers for the whole KV sweep
  uint32_t qf[KSTEPS][4];
  {
    const int m = lane >> 3, r = lane & 7;
    const int row = warp * 16 + (m & 1) * 8 + r;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const int col = ks * 16 + (m >> 1) * 8;
      ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], sQ + row * LD + col);
    }
  }

  float o_acc[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float row_m[2] = {-INFINITY, -INFINITY};
  float row_l[2] = {0.f, 0.f};

  const float LOG2E = 1.4426950408889634f;
  const float sl2 = p.scale * LOG2E;
  const float mask_l2 = p.mask_value * LOG2E;
  const int qrow0 = q0 + warp * 16 + (lane >> 2);  // this thread's two query rows: qrow0, qrow0 + 8
  const bool warp_active = (q0 + warp * 16) < p.Sq;
  const unsigned char* kmask = p.kmask ? p.kmask + (long long)b * p.Skv : nullptr;

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      load_kv(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __half* tK = sK + buf * BKV * LD;
    const __half* tV = sV + buf * BKV * LD;
    // ragged edges (S = 257 = 4*64 + 1 in the ViT): skip 16-key groups beyond the last valid key and warps whose
    // 16 query rows are all beyond Sq, instead of computing on zero padding
    const int valid_keys = (kv_end - t * BKV) < BKV ? (kv_end - t * BKV) : BKV;
    const int n_groups = (valid_keys + 15) >> 4;
    if (warp_active) {

    // ---- S = Q K^T  (16 x 64 per warp) ----
    float s[BKV / 8][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    {
      const int m = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
        for (int np = 0; np < BKV / 16; ++np) {
          if (np >= n_groups) continue;
          uint32_t b0, b1, b2, b3;
          const int key = np * 16 + (m >> 1) * 8 + r;
          const int col = ks * 16 + (m & 1) * 8;
          ldmatrix_x4(b0, b1, b2, b3, tK + key * LD + col);
          mma_16816(s[2 * np], qf[ks], b0, b1);
          mma_16816(s[2 * np + 1], qf[ks], b2, b3);
        }
      }
    }

    // ---- scale, mask, online softmax ----
    const int kbase = t * BKV + (lane & 3) * 2;
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase + i * 8 + (e & 1);
        const int qrow = qrow0 + (e >> 1) * 8;
        float x = s[i][e] * sl2;
        if (kmask != nullptr && key < p.Skv && kmask[key] == 0) x += mask_l2;
        if (key >= p.Skv || (p.causal && key > qrow + kv_off)) x = -INFINITY;
        s[i][e] = x;
        tmax[e >> 1] = fmaxf(tmax[e >> 1], x);
      }
    }
    float corr[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float m = tmax[rr];
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      const float m_new = fmaxf(row_m[rr], m);
      // fully masked so far (only possible for rows >= Sq): keep everything at zero, avoid inf-inf
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      corr[rr] = exp2f(row_m[rr] - m_use);
      row_m[rr] = m_new;
      tmax[rr] = m_use;
    }
    float tsum[2] = {0.f, 0.f};
    uint32_t pf[BKV / 16][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) {
      const float p0 = exp2f(s[i][0] - tmax[0]);
      const float p1 = exp2f(s[i][1] - tmax[0]);
      const float p2 = exp2f(s[i][2] - tmax[1]);
      const float p3 = exp2f(s[i][3] - tmax[1]);
      tsum[0] += p0 + p1;
      tsum[1] += p2 + p3;
      pf[i >> 1][(i & 1) * 2 + 0] = pack_half2(p0, p1);
      pf[i >> 1][(i & 1) * 2 + 1] = pack_half2(p2, p3);
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float x = tsum[rr];
      x += __shfl_xor_sync(0xffffffffu, x, 1);
      x += __shfl_xor_sync(0xffffffffu, x, 2);
      row_l[rr] = row_l[rr] * corr[rr] + x;
    }
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
      o_acc[i][0] *= corr[0];
      o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1];
      o_acc[i][3] *= corr[1];
    }

    // ---- O += P V ----
    {
      const int m = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < BKV / 16; ++ks) {
        if (ks >= n_groups) continue;
#pragma unroll
        for (int np = 0; np < NT_O / 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const int key = ks * 16 + (m & 1) * 8 + r;
          const int col = np * 16 + (m >> 1) * 8;
          ldmatrix_x4_trans(b0, b1, b2, b3, tV + key * LD + col);
          mma_16816(o_acc[2 * np], pf[ks], b0, b1);
          mma_16816(o_acc[2 * np + 1], pf[ks], b2, b3);
        }
      }
    }
    }  // warp_active
    __syncthreads();  // everyone done with buf before it is refilled
  }

  // ---- normalise and store ----
  __half* og = p.o + b * p.o_sb + h * p.o_sh;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int qrow = qrow0 + rr * 8;
    if (qrow >= p.Sq) continue;
    const float inv = row_l[rr] > 0.f ? 1.f / row_l[rr] : 0.f;
    __half* orow = og + (long long)qrow * p.o_ss;
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
      const int col = i * 8 + (lane & 3) * 2;
      if (col < DH) {
        *reinterpret_cast<uint32_t*>(orow + col) = pack_half2(o_acc[i][rr * 2] * inv, o_acc[i][rr * 2 + 1] * inv);
      }
    }
  }
}

template <int DH, int DP, int BQ>
int launch_attention(const AttnParams& p, cudaStream_t stream) {
  constexpr int LD = DP + 8;
  constexpr int SMEM = (BQ + 4 * BKV) * LD * 2;
  constexpr int ATT_THREADS = BQ * 2;
  auto kern = attention_f16_kernel<DH, DP, BQ>;
  static bool configured = false;
  if (!configured) {
    STLLM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid((p.Sq + BQ - 1) / BQ, p.H, p.B);
  STLLM_CHECK_CUDA(stllm_launch(kern, grid, dim3(ATT_THREADS), SMEM, stream, p));
  return STLLM_OK;
}

}  // namespace

int stllm_attention_tc_try(const void* q, const void* k, const void* v, void* out, int B, int H, int Sq, int Skv, int D,
                           const long long* qs, const long long* ks, const long long* vs, const long long* os,
                           float scale, float q_prescale, int causal, const void* key_mask, float mask_value,
                           cudaStream_t stream);

extern "C" int stllm_attention_f16(const void* q, const void* k, const void* v, void* out, int B, int H, int Sq,
                                   int Skv, int D, const long long* q_strides, const long long* k_strides,
                                   const long long* v_strides, const long long* o_strides, float scale,
                                   float q_prescale, int causal, const void* key_mask, float mask_value,
                                   void* stream) {
  STLLM_CHECK_ARG(q && k && v && out && q_strides && k_strides && v_strides && o_strides);
  STLLM_CHECK_ARG(B > 0 && H > 0 && Sq > 0 && Skv > 0);
  STLLM_CHECK_ARG(H <= 65535 && B <= 65535);
  AttnParams p;
  p.q = (const __half*)q; p.k = (const __half*)k; p.v = (const __half*)v; p.o = (__half*)out;
  p.kmask = (const unsigned char*)key_mask;
  p.q_sb = q_strides[0]; p.q_ss = q_strides[1]; p.q_sh = q_strides[2];
  p.k_sb = k_strides[0]; p.k_ss = k_strides[1]; p.k_sh = k_strides[2];
  p.v_sb = v_strides[0]; p.v_ss = v_strides[1]; p.v_sh = v_strides[2];
  p.o_sb = o_strides[0]; p.o_ss = o_strides[1]; p.o_sh = o_strides[2];
  for (int i = 0; i < 3; ++i) {
    STLLM_CHECK_ARG(q_strides[i] % 8 == 0 && k_strides[i] % 8 == 0 && v_strides[i] % 8 == 0);
    STLLM_CHECK_ARG(o_strides[i] % 2 == 0);
  }
  STLLM_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0 &&
                  ((uintptr_t)out & 3) == 0);
  p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv;
  p.scale = scale; p.q_prescale = q_prescale; p.mask_value = mask_value; p.causal = causal;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  {
    // prefill shapes of the ViT / Llama go to the tcgen05 kernel (attention_tc.cu); everything else stays here
    const int rc_tc = stllm_attention_tc_try(q, k, v, out, B, H, Sq, Skv, D, q_strides, k_strides, v_strides, o_strides,
                                             scale, q_prescale, causal, key_mask, mask_value, s);
    if (rc_tc != STLLM_ERR_UNSUPPORTED) return rc_tc;
  }
  switch (D) {
    case 64: return launch_attention<64, 64, 64>(p, s);
    case 88:
      // BQ = 96 (3 CTAs per (frame, head) at S = 257 instead of 5) was measured SLOWER: 168 registers x 192 threads
      // limits occupancy to 2 CTAs/SM (ncu: 84.8 us vs 63.7 us per ViT layer, profiles/r01_ncu_summary.md)
      return launch_attention<88, 96, 64>(p, s);
    case 128: return launch_attention<128, 128, 64>(p, s);
    default:
      stllm_set_last_error("attention: unsupported head_dim %d (supported: 64, 88, 128)", D);
      return STLLM_ERR_UNSUPPORTED;
  }
}
