// Single-token decode step of the Llama decoder for sm_100a (the `generate` loop every harness of the
// reference runs after the prefill: conversation.py:231-243, mvbench_conversation.py:114-125).
//
// A decode step multiplies R <= 8 activation rows (R = beams) with every weight matrix once: the step is bound by
// STREAMING THE WEIGHTS FROM HBM (13.5 GB per token for Vicuna-7B), not by the tensor pipe.  The kernels here are
// therefore byte movers: 16-byte streaming loads with many bytes in flight per SM, fp32 accumulation, no reshaping into
// 128-row MMA tiles.
//
//   decode_linear_kernel      y[R,N] = epi(norm(x)[R,K] * W[N,K]^T) (+ residual)      modeling_llama_mem.py:61-78 (RMSNorm),
//                             160-163 (q/k/v/o), 139-143 (gate/up/down, SiLU);          st_llm.py:108,122 (lm_head)
//       persistent CTAs of 16 warps; a CTA owns 16 weight rows at a time, its warps split K (interleaved 32-element
//       steps -> adjacent warps read adjacent 64-byte pieces of the same rows); the products run on mma.sync m16n8k16 with
//       the ACTIVATIONS as the 16-row operand (rows >= R unused) and 8 weight rows as the n operand, so one lane loads 16
//       contiguous bytes of a weight row and feeds two MMAs (the k permutation inside a step is the same for both
//       operands).  The RMSNorm in front of the Linear is part of the prologue: either explicit (w (.) fp16(x rstd), the
//       reference's rounding) or folded (the packed weights already carry the norm weight; the result is scaled by rstd).
//       The first batch of a CTA's weight loads is issued BEFORE griddepcontrol.wait (and the next block's before the
//       reduction barrier): HBM keeps streaming across kernel boundaries.
//   decode_attention_fused_kernel   one launch per layer: RoPE of q / k, append of the new token's k / v to the per-beam
//       suffix cache, softmax(scale q K^T) V for one query token per beam over [shared prompt KV || per-beam generated KV],
//       merge of the key splits by the last CTA of a head (atomic ticket).  The prompt's K/V is ONE copy read once for all
//       beams (HF 4.28 prefills and re-reads num_beams copies); generated K/V is addressed through an ancestry table
//       src[beam][j], so a beam re-ordering permutes a few ints instead of copying the cache (HF _reorder_cache
//       index_selects all of it).  Prompt-key scores run on mma.sync with the beams as the 16-row operand.
//                                                                                      modeling_llama_mem.py:81-127,205-244
//   decode_rope_append_kernel / decode_attention_kernel / decode_attention_combine_kernel   the same work as three
//       launches (first version, kept as the A/B reference: STLLM_DECODE_FUSED_ATTN=0).
#include "common.cuh"

namespace {

constexpr int DL_THREADS = 512;
constexpr int DL_WARPS = DL_THREADS / 32;
constexpr int DL_RMAX = 8;

struct DecLinParams {
  const __half* x;
  long long ldx;
  const __half* W;
  long long ldw;
  __half* y;
  long long ldy;
  const __half* residual;
  long long ldr;
  const __half* norm_w;
  int R, N, K, Kp;   // Kp: row stride (halves) of the staged activations; (2 Kp) % 128 == 64 -> conflict-free LDS.128
  int norm_mode;     // 0 none | 1 folded RMSNorm (result * rstd) | 2 explicit RMSNorm (norm_w (.) fp16(x rstd))
  int epi;           // 0 none | 2 SwiGLU (rows interleaved gate_j, up_j; N/2 outputs)
  float eps;
  int prefetch;
};

DEVI void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// weights are read exactly once per token and are far larger than L2: streaming (evict-first) loads
DEVI uint4 ld_stream(const __half* p) {
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
DEVI uint4 ld_stream_nol1(const __half* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
// one instruction pulls `bytes` (multiple of 16) into L2
DEVI void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
DEVI void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
DEVI float dec_silu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.44269504088896340736f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;   // same arithmetic as the SwiGLU epilogue of the prefill GEMM (gemm_tcgen05.cu: silu)
}

// DL_UNROLL 32-element steps are in flight per warp: 4 with two CTAs per SM (<= 64 registers), 8 when the staged
// activations leave room for one CTA only (down_proj with several beams)
// X_DIRECT: the activations are NOT staged in shared memory but read through L1 (read-only path) by every lane, 16 bytes
// per step -- for wide inputs with several beams (down_proj: 5 x 11008 halves = 110 KB would leave room for one CTA per SM
// only).  No normalisation in this mode.  The weights bypass L1 (no_allocate) so the activations stay resident there.
template <int DL_UNROLL, int MIN_CTAS, bool X_DIRECT>
__global__ void __launch_bounds__(DL_THREADS, MIN_CTAS) decode_linear_kernel(const DecLinParams p) {
  extern __shared__ __align__(16) uint8_t dl_smem[];
  __half* xs = reinterpret_cast<__half*>(dl_smem);                                   // [R][Kp]
  float* part = reinterpret_cast<float*>(dl_smem + (X_DIRECT ? (size_t)0 : (size_t)p.R * p.Kp * 2));   // [2][16 warps][8 x 16]
  float* ssq_s = part + 2 * DL_WARPS * 128;                                          // [16]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nblk = (p.N + 15) >> 4;
  const int nsteps = p.K >> 5;
  pdl_launch_dependents();
  // The weights do not depend on the previous kernel: the first batch of this CTA's weight loads (DL_UNROLL steps of
  // 2 x 16 bytes per lane: 64 KB per CTA) is requested BEFORE griddepcontrol.wait and stays in registers while the
  // predecessor drains and the activations are staged -- HBM keeps streaming across the kernel boundary.  p.prefetch == 2
  // additionally pulls the rest of the first 16 rows into L2.
  uint4 wa[DL_UNROLL], wb[DL_UNROLL];
  bool preloaded = false;
  auto load_batch = [&](int blk, int s0) {
    const int n0 = blk << 4;
    const int rowA = n0 + g, rowB = n0 + 8 + g;
    const __half* wA = p.W + (long long)(rowA < p.N ? rowA : 0) * p.ldw + 8 * t;
    const __half* wB = p.W + (long long)(rowB < p.N ? rowB : 0) * p.ldw + 8 * t;
#pragma unroll
    for (int u = 0; u < DL_UNROLL; ++u) {
      const int s = s0 + u * DL_WARPS;
      if (s < nsteps) {
        wa[u] = X_DIRECT ? ld_stream_nol1(wA + (s << 5)) : ld_stream(wA + (s << 5));
        wb[u] = X_DIRECT ? ld_stream_nol1(wB + (s << 5)) : ld_stream(wB + (s << 5));
      } else {
        wa[u] = make_uint4(0u, 0u, 0u, 0u);
        wb[u] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  if (p.prefetch && (int)blockIdx.x < nblk) {
    load_batch(blockIdx.x, warp);
    preloaded = true;
    if (p.prefetch == 2) {
      const int sectors_per_row = p.K >> 4;   // 32-byte sectors
      const int total = 16 * sectors_per_row;
      const int n0 = (int)blockIdx.x << 4;
      for (int i = tid; i < total; i += DL_THREADS) {
        const int r = i / sectors_per_row, s = i - r * sectors_per_row;
        if (n0 + r < p.N) prefetch_l2(p.W + (long long)(n0 + r) * p.ldw + (s << 4));
      }
    }
  }
  pdl_wait();
  // ---- stage the activations: one flat, coalesced copy by all threads (independent 16-byte loads, one round trip) ----
  if (!X_DIRECT) {
    const int nvec = p.K >> 3;
    const int total = p.R * nvec;
    for (int f0 = tid; f0 < total; f0 += DL_THREADS * 4) {
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = f0 + j * DL_THREADS;
        if (f < total) {
          const int r = f / nvec, v = f - r * nvec;
          u[j] = *(reinterpret_cast<const uint4*>(p.x + (long long)r * p.ldx) + v);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = f0 + j * DL_THREADS;
        if (f < total) {
          const int r = f / nvec, v = f - r * nvec;
          *(reinterpret_cast<uint4*>(xs + (size_t)r * p.Kp) + v) = u[j];
        }
      }
    }
    __syncthreads();
    // row statistics from shared memory: warp w -> row w & 7, half (w >> 3) of the row
    if (p.norm_mode != 0) {
      const int r = warp & 7, hh = warp >> 3;
      const int v0 = hh ? (nvec >> 1) : 0, v1 = hh ? nvec : (nvec >> 1);
      float sq = 0.f;
      uint4* row = reinterpret_cast<uint4*>(xs + (size_t)r * p.Kp);
      if (r < p.R) {
        for (int v = v0 + lane; v < v1; v += 32) {
          const uint4 u = row[v];
          const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            sq += f.x * f.x + f.y * f.y;
          }
        }
      }
      sq = warp_sum(sq);
      if (lane == 0) ssq_s[warp] = sq;
      __syncthreads();
      if (p.norm_mode == 2) {
        if (r < p.R) {
          const float rstd = rsqrtf((ssq_s[r] + ssq_s[r + 8]) / (float)p.K + p.eps);
          const uint4* wsrc = reinterpret_cast<const uint4*>(p.norm_w);
          for (int v = v0 + lane; v < v1; v += 32) {
            uint4 u = row[v];
            const uint4 wv = __ldg(wsrc + v);
            __half2* h = reinterpret_cast<__half2*>(&u);
            const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(h[j]);
              h[j] = __hmul2(wh[j], __floats2half2_rn(f.x * rstd, f.y * rstd));   // rounding of norm.cu / the reference
            }
            row[v] = u;
          }
        }
        __syncthreads();
      }
    }
  }
  const __half* xr = X_DIRECT ? p.x + (long long)(g < p.R ? g : 0) * p.ldx + 8 * t
                              : xs + (size_t)(g < p.R ? g : 0) * p.Kp + 8 * t;
  int it = 0;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x, ++it) {
    const int n0 = blk << 4;
    float accA[4] = {0.f, 0.f, 0.f, 0.f}, accB[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s0 = warp; s0 < nsteps; s0 += DL_WARPS * DL_UNROLL) {
      if (!preloaded) load_batch(blk, s0);
      preloaded = false;
#pragma unroll
      for (int u = 0; u < DL_UNROLL; ++u) {
        const int s = s0 + u * DL_WARPS;
        const uint4 xa = X_DIRECT ? __ldg(reinterpret_cast<const uint4*>(xr + ((s < nsteps ? s : 0) << 5)))
                                  : *reinterpret_cast<const uint4*>(xr + ((s < nsteps ? s : 0) << 5));
        mma16816(accA, xa.x, 0u, xa.y, 0u, wa[u].x, wa[u].y);
        mma16816(accA, xa.z, 0u, xa.w, 0u, wa[u].z, wa[u].w);
        mma16816(accB, xa.x, 0u, xa.y, 0u, wb[u].x, wb[u].y);
        mma16816(accB, xa.z, 0u, xa.w, 0u, wb[u].z, wb[u].w);
      }
    }
    // the next block's first batch is requested before this block's reduction (barrier + epilogue) starts
    if (p.prefetch && blk + (int)gridDim.x < nblk) {
      load_batch(blk + gridDim.x, warp);
      preloaded = true;
    }
    // D[beam g][weight row]: lane holds rows n0 + 2t, 2t+1 (group A) and n0 + 8 + 2t, 2t+1 (group B) of beam g
    float* pw = part + ((it & 1) * DL_WARPS + warp) * 128 + g * 16 + 2 * t;
    *reinterpret_cast<float2*>(pw) = make_float2(accA[0], accA[1]);
    *reinterpret_cast<float2*>(pw + 8) = make_float2(accB[0], accB[1]);
    __syncthreads();
    if (tid < 128) {
      const int r = tid >> 4, c = tid & 15;
      const float* pr = part + (it & 1) * DL_WARPS * 128 + tid;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < DL_WARPS; ++w) v += pr[w * 128];
      if (p.norm_mode == 1) v *= rsqrtf((ssq_s[r] + ssq_s[r + 8]) / (float)p.K + p.eps);
      const int n = n0 + c;
      if (p.epi == 2) {
        const float up = __shfl_down_sync(0xffffffffu, v, 1);
        if (!(c & 1) && r < p.R && n + 1 < p.N) {
          float o = dec_silu(v) * up;
          const int oc = n >> 1;
          if (p.residual) o += __half2float(p.residual[(long long)r * p.ldr + oc]);
          p.y[(long long)r * p.ldy + oc] = __float2half_rn(o);
        }
      } else if (r < p.R && n < p.N) {
        if (p.residual) v += __half2float(p.residual[(long long)r * p.ldr + n]);
        p.y[(long long)r * p.ldy + n] = __float2half_rn(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct DecRopeParams {
  __half* qkv;
  long long ldq;       // row = [q | k | v], each H*D
  const float* cos_t;
  const float* sin_t;
  const int* n_gen;    // device scalar: tokens generated before this step
  __half* ks;
  __half* vs;
  long long s_beam_stride, s_tok_stride;
  int* src;
  int src_ld;
  int prefix_len, max_pos, l_max, R, H, D;
};

__global__ void decode_rope_append_kernel(const DecRopeParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int j_suf = *p.n_gen;
  const int pos = p.prefix_len + j_suf;
  if (j_suf >= p.l_max || pos >= p.max_pos) return;   // the host never launches past the capacity
  const int half_d = p.D >> 1, groups = half_d >> 3;
  const int total = p.R * p.H * groups;
  const long long HD = (long long)p.H * p.D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = (i % groups) * 8;
    const int rh = i / groups;
    const int h = rh % p.H, r = rh / p.H;
    if (h == 0 && j == 0) p.src[r * p.src_ld + j_suf] = r;   // the new token of beam r lives in physical row r
    const float4* cp = reinterpret_cast<const float4*>(p.cos_t + (long long)pos * half_d + j);
    const float4* sp = reinterpret_cast<const float4*>(p.sin_t + (long long)pos * half_d + j);
    const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
    const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    __half* row = p.qkv + (long long)r * p.ldq + (long long)h * p.D;
    const long long cache_off = (long long)r * p.s_beam_stride + (long long)j_suf * p.s_tok_stride + (long long)h * p.D;
#pragma unroll
    for (int which = 0; which < 2; ++which) {   // 0: q (in place), 1: k (-> cache)
      __half* base = row + which * HD;
      const uint4 ulo = *reinterpret_cast<const uint4*>(base + j);
      const uint4 uhi = *reinterpret_cast<const uint4*>(base + j + half_d);
      const __half2* hlo = reinterpret_cast<const __half2*>(&ulo);
      const __half2* hhi = reinterpret_cast<const __half2*>(&uhi);
      uint4 olo, ohi;
      uint32_t* plo = reinterpret_cast<uint32_t*>(&olo);
      uint32_t* phi = reinterpret_cast<uint32_t*>(&ohi);
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // same arithmetic as rope_kernel (elementwise.cu)
        const float2 lo = __half22float2(hlo[q]);
        const float2 hi = __half22float2(hhi[q]);
        plo[q] = pack_half2(lo.x * c[2 * q] - hi.x * sn[2 * q], lo.y * c[2 * q + 1] - hi.y * sn[2 * q + 1]);
        phi[q] = pack_half2(hi.x * c[2 * q] + lo.x * sn[2 * q], hi.y * c[2 * q + 1] + lo.y * sn[2 * q + 1]);
      }
      __half* dst = which == 0 ? base : p.ks + cache_off;
      *reinterpret_cast<uint4*>(dst + j) = olo;
      *reinterpret_cast<uint4*>(dst + j + half_d) = ohi;
    }
    const __half* vsrc = row + 2 * HD;
    __half* vdst = p.vs + cache_off;
    *reinterpret_cast<uint4*>(vdst + j) = *reinterpret_cast<const uint4*>(vsrc + j);
    *reinterpret_cast<uint4*>(vdst + j + half_d) = *reinterpret_cast<const uint4*>(vsrc + j + half_d);
  }
}

// ------------------------------------------------------------------------------------------------------------------
constexpr int DA_THREADS = 512;
constexpr int DA_WARPS = DA_THREADS / 32;
constexpr int DA_D = 128;        // head dim (Vicuna / Llama: 4096 / 32)
constexpr int DA_WS_ROW = 132;   // floats per (head, split, beam) in the workspace: o[128], m, l, pad

struct DecAttnParams {
  const __half* q;
  long long ldq;
  const __half* kp;
  const __half* vp;
  long long p_stride;
  const uint8_t* p_mask;
  const __half* ks;
  const __half* vs;
  long long s_beam_stride, s_tok_stride;
  const int* src;
  int src_ld;
  const int* n_gen;
  float* ws;
  int R, H, Sp, NS, cap;
  float scale;
};

DEVI void half8_to_float(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 v = __half22float2(h[j]);
    f[2 * j] = v.x;
    f[2 * j + 1] = v.y;
  }
}
DEVI float sum16(float v) {   // over the 16 lanes of a half-warp
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__global__ void __launch_bounds__(DA_THREADS, 1) decode_attention_kernel(const DecAttnParams p) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  float* sc = reinterpret_cast<float*>(da_smem);            // [8][cap] scores, then probabilities
  float* red = sc + (size_t)DL_RMAX * p.cap;                // [16 warps][R][128] prefix partials
  float* red_s = red + (size_t)DA_WARPS * p.R * DA_D;       // [16 warps][128] suffix partials (warp w -> beam w % R)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, hw = lane >> 4, c = lane & 15;
  const int h = blockIdx.x, sp = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  const int n_suf = *p.n_gen + 1;   // the token appended by this step is a key too
  const int pk0 = (int)((long long)p.Sp * sp / p.NS), pk1 = (int)((long long)p.Sp * (sp + 1) / p.NS);
  const int sk0 = (int)((long long)n_suf * sp / p.NS), sk1 = (int)((long long)n_suf * (sp + 1) / p.NS);
  const int np = pk1 - pk0, ns = sk1 - sk0, ntot = np + ns;
  const long long hoff = (long long)h * DA_D + 8 * c;
  constexpr int U = 4;
  constexpr float NEG_INF = -INFINITY;
  // generated keys differ per beam: the warps are dealt out to the beams (warp w -> beam w % R, slice w / R of nsub), so
  // all beams walk their ancestry chains at the same time
  const int nsub = DA_WARPS / p.R;
  const int my_r = warp % p.R, my_sub = warp / p.R;
  const bool suf_warp = my_sub < nsub;

  // ---- scores: half-warp per key, lane = 16-byte chunk of the head dim ----
  {
    float qf[DL_RMAX][8];
#pragma unroll
    for (int r = 0; r < DL_RMAX; ++r) {
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (r < p.R) u = *reinterpret_cast<const uint4*>(p.q + (long long)r * p.ldq + hoff);
      half8_to_float(u, qf[r]);
    }
    for (int b0 = 2 * warp; b0 < np; b0 += 2 * DA_WARPS * U) {   // warp-uniform trip count (shuffles inside)
      const int i0 = b0 + hw;
      uint4 kv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + 2 * DA_WARPS * u;
        kv[u] = i < np ? __ldg(reinterpret_cast<const uint4*>(p.kp + (long long)(pk0 + i) * p.p_stride + hoff))
                       : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + 2 * DA_WARPS * u;
        float kf[8];
        half8_to_float(kv[u], kf);
        const bool masked = p.p_mask != nullptr && i < np && p.p_mask[pk0 + i] == 0;
#pragma unroll
        for (int r = 0; r < DL_RMAX; ++r) {
          if (r < p.R) {
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) d = fmaf(qf[r][e], kf[e], d);
            d = sum16(d);
            if (c == 0 && i < np) sc[r * p.cap + i] = masked ? NEG_INF : d * p.scale;
          }
        }
      }
    }
    if (suf_warp) {
      float qs[8];
      half8_to_float(*reinterpret_cast<const uint4*>(p.q + (long long)my_r * p.ldq + hoff), qs);
      for (int c0 = sk0 + 2 * my_sub; c0 < sk1; c0 += 2 * nsub * U) {
        const int j0 = c0 + hw;
        uint4 kv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + 2 * nsub * u;
          kv[u] = make_uint4(0u, 0u, 0u, 0u);
          if (j < sk1) {
            const int pr = p.src[my_r * p.src_ld + j];
            kv[u] = *reinterpret_cast<const uint4*>(p.ks + (long long)pr * p.s_beam_stride +
                                                    (long long)j * p.s_tok_stride + hoff);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + 2 * nsub * u;
          float kf[8];
          half8_to_float(kv[u], kf);
          float d = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) d = fmaf(qs[e], kf[e], d);
          d = sum16(d);
          if (c == 0 && j < sk1) sc[my_r * p.cap + np + (j - sk0)] = d * p.scale;
        }
      }
    }
  }
  __syncthreads();
  // ---- fp32 softmax of this split: warp r -> beam r ----
  float* wsb = p.ws + ((size_t)(h * p.NS + sp) * DL_RMAX) * DA_WS_ROW;
  if (warp < p.R) {
    float* s = sc + warp * p.cap;
    float m = NEG_INF;
    for (int i = lane; i < ntot; i += 32) m = fmaxf(m, s[i]);
    m = warp_max(m);
    float l = 0.f;
    for (int i = lane; i < ntot; i += 32) {
      const float e = (m == NEG_INF) ? 0.f : __expf(s[i] - m);
      s[i] = e;
      l += e;
    }
    l = warp_sum(l);
    if (lane == 0) {
      wsb[warp * DA_WS_ROW + 128] = m;
      wsb[warp * DA_WS_ROW + 129] = l;
    }
  }
  __syncthreads();
  // ---- o = P V (unnormalised) ----
  {
    float acc[DL_RMAX][8];
#pragma unroll
    for (int r = 0; r < DL_RMAX; ++r)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[r][e] = 0.f;
    for (int b0 = 2 * warp; b0 < np; b0 += 2 * DA_WARPS * U) {
      const int i0 = b0 + hw;
      uint4 vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + 2 * DA_WARPS * u;
        vv[u] = i < np ? __ldg(reinterpret_cast<const uint4*>(p.vp + (long long)(pk0 + i) * p.p_stride + hoff))
                       : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + 2 * DA_WARPS * u;
        if (i < np) {
          float vf[8];
          half8_to_float(vv[u], vf);
#pragma unroll
          for (int r = 0; r < DL_RMAX; ++r) {
            if (r < p.R) {
              const float pr = sc[r * p.cap + i];
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[r][e] = fmaf(pr, vf[e], acc[r][e]);
            }
          }
        }
      }
    }
    // the two half-warps of a warp hold different keys: fold them, then one partial per warp
#pragma unroll
    for (int r = 0; r < DL_RMAX; ++r) {
      if (r < p.R) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[r][e] += __shfl_xor_sync(0xffffffffu, acc[r][e], 16);
        if (hw == 0) {
          float* dst = red + ((size_t)warp * p.R + r) * DA_D + 8 * c;
          *reinterpret_cast<float4*>(dst) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
        }
      }
    }
    // generated keys of this warp's beam
    float as[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) as[e] = 0.f;
    if (suf_warp) {
      for (int c0 = sk0 + 2 * my_sub; c0 < sk1; c0 += 2 * nsub * U) {
        const int j0 = c0 + hw;
        uint4 vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + 2 * nsub * u;
          vv[u] = make_uint4(0u, 0u, 0u, 0u);
          if (j < sk1) {
            const int pr = p.src[my_r * p.src_ld + j];
            vv[u] = *reinterpret_cast<const uint4*>(p.vs + (long long)pr * p.s_beam_stride +
                                                    (long long)j * p.s_tok_stride + hoff);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + 2 * nsub * u;
          if (j < sk1) {
            float vf[8];
            half8_to_float(vv[u], vf);
            const float pr = sc[my_r * p.cap + np + (j - sk0)];
#pragma unroll
            for (int e = 0; e < 8; ++e) as[e] = fmaf(pr, vf[e], as[e]);
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) as[e] += __shfl_xor_sync(0xffffffffu, as[e], 16);
    if (hw == 0) {
      float* dst = red_s + (size_t)warp * DA_D + 8 * c;
      *reinterpret_cast<float4*>(dst) = make_float4(as[0], as[1], as[2], as[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(as[4], as[5], as[6], as[7]);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < p.R * DA_D; idx += DA_THREADS) {
    const int r = idx >> 7, d = idx & 127;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) s += red[((size_t)w * p.R + r) * DA_D + d];
    for (int w = r; w < nsub * p.R; w += p.R) s += red_s[(size_t)w * DA_D + d];
    wsb[r * DA_WS_ROW + d] = s;
  }
}

// out[r, h*128 + d] = sum_s o_s[d] e^(m_s - M) / sum_s l_s e^(m_s - M)
__global__ void __launch_bounds__(DA_D) decode_attention_combine_kernel(const float* __restrict__ ws, __half* out,
                                                                        long long ldo, int NS) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, r = blockIdx.y, d = threadIdx.x;
  const float* base = ws + ((size_t)h * NS * DL_RMAX + r) * DA_WS_ROW;
  const size_t split_stride = (size_t)DL_RMAX * DA_WS_ROW;
  float M = -INFINITY;
  for (int s = 0; s < NS; ++s) M = fmaxf(M, base[s * split_stride + 128]);
  float L = 0.f, o = 0.f;
  if (M != -INFINITY) {
    for (int s = 0; s < NS; ++s) {
      const float w = __expf(base[s * split_stride + 128] - M);
      L = fmaf(base[s * split_stride + 129], w, L);
      o = fmaf(base[s * split_stride + d], w, o);
    }
  }
  out[(long long)r * ldo + (long long)h * DA_D + d] = __float2half_rn(L > 0.f ? o / L : 0.f);
}

// ------------------------------------------------------------------------------------------------------------------
// Fused variant: RoPE of q / k, the cache append, the attention of one (head, key split) and the merge of the splits in
// ONE launch (3 launches -> 1 per layer).  Every CTA ropes the R query rows of its head itself (R x 128 values); the
// CTA of the LAST split owns the new token: it ropes k, stores k / v into the suffix caches and scores the new key from
// shared memory, so no CTA reads a cache slot written by this launch.  The last CTA of a head to finish (atomic
// ticket) merges the splits.  All global loads of a CTA are issued up front (prefix K and V, the ancestry entries, then
// the generated K and V): two dependent round trips instead of eight.
struct DecAttnFusedParams {
  __half* qkv;
  long long ldq;
  const float* cos_t;
  const float* sin_t;
  const __half* kp;
  const __half* vp;
  long long p_stride;
  const uint8_t* p_mask;
  __half* ks;
  __half* vs;
  long long s_beam_stride, s_tok_stride;
  int* src;
  int src_ld;
  const int* n_gen;
  __half* out;
  long long ldo;
  float* ws;
  int* tickets;   // [H], zero before the first launch; reset by the merging CTA
  const uint8_t* pf_ptr[2];   // optional: memory the NEXT kernels stream (o_proj weights, the head of gate|up) -> pulled
  long long pf_bytes[2];      // into L2 while this kernel's latency chain leaves HBM idle
  int R, H, Sp, NS, cap, l_max, max_pos;
  float scale;
};

__global__ void __launch_bounds__(DA_THREADS, 1) decode_attention_fused_kernel(const DecAttnFusedParams p) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  float* sc = reinterpret_cast<float*>(da_smem);            // [8][cap]
  float* red = sc + (size_t)DL_RMAX * p.cap;                // [16 warps][R][128]
  float* red_s = red + (size_t)DA_WARPS * p.R * DA_D;       // [16 warps][128]
  float* qs = red_s + (size_t)DA_WARPS * DA_D;              // [8][128] roped q (fp16-rounded values)
  float* kn = qs + DL_RMAX * DA_D;                          // [8][128] roped k of the new token (last split only)
  float* vn = kn + DL_RMAX * DA_D;                          // [8][128] v of the new token
  __half* qh = reinterpret_cast<__half*>(vn + DL_RMAX * DA_D);   // [8][128] roped q in fp16 (rows >= R zero): MMA A operand
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, hw = lane >> 4, c = lane & 15;
  const int g = lane >> 2, t = lane & 3;   // mma.sync fragment coordinates
  const int h = blockIdx.x, sp = blockIdx.y;
  const bool last_split = sp == p.NS - 1;
  pdl_launch_dependents();
  const int pk0 = (int)((long long)p.Sp * sp / p.NS), pk1 = (int)((long long)p.Sp * (sp + 1) / p.NS);
  const int np = pk1 - pk0;
  const long long hoff = (long long)h * DA_D + 8 * c;
  constexpr int U = 4;
  constexpr float NEG_INF = -INFINITY;
  const int nsub = DA_WARPS / p.R;
  const int my_r = warp % p.R, my_sub = warp / p.R;
  const bool suf_warp = my_sub < nsub;
  // ---- (a) the prompt's K and V of the first batch: static data, requested before the predecessor has finished ----
  // K in the mma.sync B-fragment layout (lane (g, t): key g of an 8-key group, dims 32 s + 8 t .. + 7 of step s), NG groups
  // per warp = 256 keys per CTA; V in the half-warp-per-key layout of the P V loop, UV keys per half-warp = 160 per CTA
  constexpr int NG = 2, UV = 5;
  uint4 kq[NG][4], vv0[UV];
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) {
    const int i = 8 * (warp + DA_WARPS * gi) + g;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      kq[gi][st] = make_uint4(0u, 0u, 0u, 0u);
      if (i < np)
        kq[gi][st] = __ldg(reinterpret_cast<const uint4*>(p.kp + (long long)(pk0 + i) * p.p_stride + (long long)h * DA_D +
                                                          32 * st + 8 * t));
    }
  }
#pragma unroll
  for (int u = 0; u < UV; ++u) {
    const int i = 2 * warp + hw + 2 * DA_WARPS * u;
    vv0[u] = make_uint4(0u, 0u, 0u, 0u);
    if (i < np) vv0[u] = __ldg(reinterpret_cast<const uint4*>(p.vp + (long long)(pk0 + i) * p.p_stride + hoff));
  }
  // bulk L2 prefetch (8 KB per instruction, one lane per warp, consecutive warps of the grid take consecutive chunks)
  if (lane == 0) {
    constexpr long long CH = 8192;
    const long long nwarps = (long long)gridDim.x * gridDim.y * DA_WARPS;
    const long long me = ((long long)blockIdx.y * gridDim.x + blockIdx.x) + (long long)warp * gridDim.x * gridDim.y;
#pragma unroll
    for (int reg = 0; reg < 2; ++reg) {
      if (p.pf_ptr[reg] != nullptr) {
        const long long nch = (p.pf_bytes[reg] + CH - 1) / CH;
        for (long long i = me; i < nch; i += nwarps) {
          const long long left = p.pf_bytes[reg] - i * CH;
          prefetch_l2_bulk(p.pf_ptr[reg] + i * CH, (uint32_t)((left < CH ? left : CH) & ~15LL));
        }
      }
    }
  }
  pdl_wait();
  const int n_gen = *p.n_gen;
  const int n_suf = n_gen + 1;
  const int sk0 = (int)((long long)n_suf * sp / p.NS);
  const int sk1 = (int)((long long)n_suf * (sp + 1) / p.NS);
  const int sk1c = last_split ? sk1 - 1 : sk1;   // keys read from the cache; the new token (j = n_gen) comes from smem
  const int ns = sk1 - sk0, ntot = np + ns;
  const int pos = p.Sp + n_gen;
  if (n_gen >= p.l_max || pos >= p.max_pos) return;   // the host never launches past the capacity
  // ---- (b) ancestry entries of the first batch of generated keys ----
  int pr0[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int j = sk0 + 2 * my_sub + hw + 2 * nsub * u;
    pr0[u] = (suf_warp && j < sk1c) ? p.src[my_r * p.src_ld + j] : 0;
  }
  // ---- (c) RoPE of q (and of k for the new token) into shared memory ----
  for (int idx = tid; idx < DL_RMAX * 64; idx += DA_THREADS) {
    const int r = idx >> 6, i = idx & 63;
    if (r >= p.R) {
      qh[r * DA_D + i] = __float2half_rn(0.f);
      qh[r * DA_D + i + 64] = __float2half_rn(0.f);
      continue;
    }
    const float cs = __ldg(p.cos_t + (long long)pos * 64 + i), sn = __ldg(p.sin_t + (long long)pos * 64 + i);
    const __half* row = p.qkv + (long long)r * p.ldq + (long long)h * DA_D;
    const float lo = __half2float(row[i]), hi = __half2float(row[i + 64]);
    // explicit contraction = what nvcc emits for rope_kernel (elementwise.cu): fma(lo, c, -rn(hi s)), fma(hi, c, rn(lo s))
    const __half qlo = __float2half_rn(__fmaf_rn(lo, cs, -__fmul_rn(hi, sn)));
    const __half qhi = __float2half_rn(__fmaf_rn(hi, cs, __fmul_rn(lo, sn)));
    qh[r * DA_D + i] = qlo;
    qh[r * DA_D + i + 64] = qhi;
    qs[r * DA_D + i] = __half2float(qlo);
    qs[r * DA_D + i + 64] = __half2float(qhi);
    if (last_split) {
      const long long HD = (long long)p.H * DA_D;
      const float klo = __half2float(row[HD + i]), khi = __half2float(row[HD + i + 64]);
      const __half rlo = __float2half_rn(__fmaf_rn(klo, cs, -__fmul_rn(khi, sn)));
      const __half rhi = __float2half_rn(__fmaf_rn(khi, cs, __fmul_rn(klo, sn)));
      const __half vlo = row[2 * HD + i], vhi = row[2 * HD + i + 64];
      kn[r * DA_D + i] = __half2float(rlo);
      kn[r * DA_D + i + 64] = __half2float(rhi);
      vn[r * DA_D + i] = __half2float(vlo);
      vn[r * DA_D + i + 64] = __half2float(vhi);
      const long long off = (long long)r * p.s_beam_stride + (long long)n_gen * p.s_tok_stride + (long long)h * DA_D;
      p.ks[off + i] = rlo;
      p.ks[off + i + 64] = rhi;
      p.vs[off + i] = vlo;
      p.vs[off + i + 64] = vhi;
      if (h == 0 && i == 0) p.src[r * p.src_ld + n_gen] = r;
    }
  }
  // ---- (d) generated K and V of the first batch ----
  uint4 ks0[U], vs0[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int j = sk0 + 2 * my_sub + hw + 2 * nsub * u;
    ks0[u] = make_uint4(0u, 0u, 0u, 0u);
    vs0[u] = make_uint4(0u, 0u, 0u, 0u);
    if (suf_warp && j < sk1c) {
      const long long off = (long long)pr0[u] * p.s_beam_stride + (long long)j * p.s_tok_stride + hoff;
      ks0[u] = *reinterpret_cast<const uint4*>(p.ks + off);
      vs0[u] = *reinterpret_cast<const uint4*>(p.vs + off);
    }
  }
  __syncthreads();
  // ---- scores ----
  {
    // prompt keys on mma.sync m16n8k16: A = roped q (row = beam, rows >= R zero), B = 8 keys; D[beam g][key 2t, 2t+1].
    // One 16-byte load of a key feeds two MMAs (the k permutation inside a 32-dim step is the same for q and k).
    // First NG groups per warp from registers, further groups (np > 256) loaded here.
    {
      uint4 qa[4];
#pragma unroll
      for (int st = 0; st < 4; ++st) qa[st] = *reinterpret_cast<const uint4*>(qh + g * DA_D + 32 * st + 8 * t);
      for (int g0 = 0; 8 * g0 < np; g0 += DA_WARPS * NG) {   // block-uniform trip count (warp-collective MMAs inside)
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          const int grp = g0 + warp + DA_WARPS * gi;
          const int i = 8 * grp + g;
          uint4 kk[4];
#pragma unroll
          for (int st = 0; st < 4; ++st) {
            if (g0 == 0) kk[st] = kq[gi][st];
            else {
              kk[st] = make_uint4(0u, 0u, 0u, 0u);
              if (i < np)
                kk[st] = __ldg(reinterpret_cast<const uint4*>(p.kp + (long long)(pk0 + i) * p.p_stride +
                                                              (long long)h * DA_D + 32 * st + 8 * t));
            }
          }
          float cc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int st = 0; st < 4; ++st) {
            mma16816(cc, qa[st].x, 0u, qa[st].y, 0u, kk[st].x, kk[st].y);
            mma16816(cc, qa[st].z, 0u, qa[st].w, 0u, kk[st].z, kk[st].w);
          }
          if (g < p.R) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int key = 8 * grp + 2 * t + e;
              if (key < np) {
                const bool masked = p.p_mask != nullptr && p.p_mask[pk0 + key] == 0;
                sc[g * p.cap + key] = masked ? NEG_INF : cc[e] * p.scale;
              }
            }
          }
        }
      }
    }
    // generated keys of this warp's beam (warp-uniform loops)
    if (suf_warp) {
      float qr[8];
      {
        const float4 q0 = *reinterpret_cast<const float4*>(qs + my_r * DA_D + 8 * c);
        const float4 q1 = *reinterpret_cast<const float4*>(qs + my_r * DA_D + 8 * c + 4);
        qr[0] = q0.x; qr[1] = q0.y; qr[2] = q0.z; qr[3] = q0.w; qr[4] = q1.x; qr[5] = q1.y; qr[6] = q1.z; qr[7] = q1.w;
      }
      for (int c0 = sk0 + 2 * my_sub; c0 < sk1c; c0 += 2 * nsub * U) {
        const bool first = c0 == sk0 + 2 * my_sub;
        uint4 kv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = c0 + hw + 2 * nsub * u;
          if (first) kv[u] = ks0[u];
          else {
            kv[u] = make_uint4(0u, 0u, 0u, 0u);
            if (j < sk1c) {
              const int pr = p.src[my_r * p.src_ld + j];
              kv[u] = *reinterpret_cast<const uint4*>(p.ks + (long long)pr * p.s_beam_stride +
                                                      (long long)j * p.s_tok_stride + hoff);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = c0 + hw + 2 * nsub * u;
          float kf[8];
          half8_to_float(kv[u], kf);
          float d = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) d = fmaf(qr[e], kf[e], d);
          d = sum16(d);
          if (c == 0 && j < sk1c) sc[my_r * p.cap + np + (j - sk0)] = d * p.scale;
        }
      }
    }
    // the new token's key (last split): warp r -> beam r, from shared memory
    if (last_split && warp < p.R) {
      float d = 0.f;
      for (int i = lane; i < DA_D; i += 32) d = fmaf(qs[warp * DA_D + i], kn[warp * DA_D + i], d);
      d = warp_sum(d);
      if (lane == 0) sc[warp * p.cap + np + (n_gen - sk0)] = d * p.scale;
    }
  }
  __syncthreads();
  // ---- fp32 softmax of this split: warp r -> beam r ----
  float* wsb = p.ws + ((size_t)(h * p.NS + sp) * DL_RMAX) * DA_WS_ROW;
  if (warp < p.R) {
    float* s = sc + warp * p.cap;
    float m = NEG_INF;
    for (int i = lane; i < ntot; i += 32) m = fmaxf(m, s[i]);
    m = warp_max(m);
    float l = 0.f;
    for (int i = lane; i < ntot; i += 32) {
      const float e = (m == NEG_INF) ? 0.f : __expf(s[i] - m);
      s[i] = e;
      l += e;
    }
    l = warp_sum(l);
    if (lane == 0) {
      wsb[warp * DA_WS_ROW + 128] = m;
      wsb[warp * DA_WS_ROW + 129] = l;
    }
  }
  __syncthreads();
  // ---- o = P V (unnormalised) ----
  {
    float acc[DL_RMAX][8];
#pragma unroll
    for (int r = 0; r < DL_RMAX; ++r)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[r][e] = 0.f;
    for (int b0 = 0; b0 < np; b0 += 2 * DA_WARPS * UV) {
      uint4 vv[UV];
#pragma unroll
      for (int u = 0; u < UV; ++u) {
        const int i = b0 + 2 * warp + hw + 2 * DA_WARPS * u;
        if (b0 == 0) vv[u] = vv0[u];
        else vv[u] = i < np ? __ldg(reinterpret_cast<const uint4*>(p.vp + (long long)(pk0 + i) * p.p_stride + hoff))
                            : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < UV; ++u) {
        const int i = b0 + 2 * warp + hw + 2 * DA_WARPS * u;
        if (i < np) {
          float vf[8];
          half8_to_float(vv[u], vf);
#pragma unroll
          for (int r = 0; r < DL_RMAX; ++r) {
            if (r < p.R) {
              const float pr = sc[r * p.cap + i];
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[r][e] = fmaf(pr, vf[e], acc[r][e]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < DL_RMAX; ++r) {
      if (r < p.R) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[r][e] += __shfl_xor_sync(0xffffffffu, acc[r][e], 16);
        if (hw == 0) {
          float* dst = red + ((size_t)warp * p.R + r) * DA_D + 8 * c;
          *reinterpret_cast<float4*>(dst) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
        }
      }
    }
    float as[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) as[e] = 0.f;
    if (suf_warp) {
      for (int c0 = sk0 + 2 * my_sub; c0 < sk1c; c0 += 2 * nsub * U) {
        const bool first = c0 == sk0 + 2 * my_sub;
        uint4 vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = c0 + hw + 2 * nsub * u;
          if (first) vv[u] = vs0[u];
          else {
            vv[u] = make_uint4(0u, 0u, 0u, 0u);
            if (j < sk1c) {
              const int pr = p.src[my_r * p.src_ld + j];
              vv[u] = *reinterpret_cast<const uint4*>(p.vs + (long long)pr * p.s_beam_stride +
                                                      (long long)j * p.s_tok_stride + hoff);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = c0 + hw + 2 * nsub * u;
          if (j < sk1c) {
            float vf[8];
            half8_to_float(vv[u], vf);
            const float pr = sc[my_r * p.cap + np + (j - sk0)];
#pragma unroll
            for (int e = 0; e < 8; ++e) as[e] = fmaf(pr, vf[e], as[e]);
          }
        }
      }
      // the new token's value: added once per beam by the first of the beam's warps (half-warp 0 only: folded below)
      if (last_split && my_sub == 0 && hw == 0) {
        const float pr = sc[my_r * p.cap + np + (n_gen - sk0)];
#pragma unroll
        for (int e = 0; e < 8; ++e) as[e] = fmaf(pr, vn[my_r * DA_D + 8 * c + e], as[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) as[e] += __shfl_xor_sync(0xffffffffu, as[e], 16);
    if (hw == 0) {
      float* dst = red_s + (size_t)warp * DA_D + 8 * c;
      *reinterpret_cast<float4*>(dst) = make_float4(as[0], as[1], as[2], as[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(as[4], as[5], as[6], as[7]);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < p.R * DA_D; idx += DA_THREADS) {
    const int r = idx >> 7, d = idx & 127;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) s += red[((size_t)w * p.R + r) * DA_D + d];
    for (int w = r; w < nsub * p.R; w += p.R) s += red_s[(size_t)w * DA_D + d];
    wsb[r * DA_WS_ROW + d] = s;
  }
  // ---- merge of the splits by the last CTA of this head ----
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(p.tickets + h, 1);
    s_last = prev == p.NS - 1;
    if (s_last) p.tickets[h] = 0;   // every CTA of this head has arrived: ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* base = p.ws + (size_t)h * p.NS * DL_RMAX * DA_WS_ROW;
  const size_t split_stride = (size_t)DL_RMAX * DA_WS_ROW;
  // thread -> (beam, 4 output values); every partial of every split is requested before the first is used
  for (int idx = tid; idx < p.R * (DA_D / 4); idx += DA_THREADS) {
    const int r = idx >> 5, d = (idx & 31) * 4;
    const float* br = base + (size_t)r * DA_WS_ROW;
    constexpr int NSMAX = 16;
    float ms[NSMAX], ls[NSMAX];
    float4 os[NSMAX];
#pragma unroll
    for (int s = 0; s < NSMAX; ++s) {
      if (s < p.NS) {
        ms[s] = __ldcg(br + s * split_stride + 128);
        ls[s] = __ldcg(br + s * split_stride + 129);
        os[s] = __ldcg(reinterpret_cast<const float4*>(br + s * split_stride + d));
      }
    }
    float M = NEG_INF;
#pragma unroll
    for (int s = 0; s < NSMAX; ++s)
      if (s < p.NS) M = fmaxf(M, ms[s]);
    float L = 0.f;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (M != NEG_INF) {
#pragma unroll
      for (int s = 0; s < NSMAX; ++s) {
        if (s < p.NS) {
          const float w = __expf(ms[s] - M);
          L = fmaf(ls[s], w, L);
          o.x = fmaf(os[s].x, w, o.x); o.y = fmaf(os[s].y, w, o.y);
          o.z = fmaf(os[s].z, w, o.z); o.w = fmaf(os[s].w, w, o.w);
        }
      }
    }
    const float inv = L > 0.f ? 1.0f / L : 0.f;
    __half2* dst = reinterpret_cast<__half2*>(p.out + (long long)r * p.ldo + (long long)h * DA_D + d);
    dst[0] = __floats2half2_rn(o.x * inv, o.y * inv);
    dst[1] = __floats2half2_rn(o.z * inv, o.w * inv);
  }
}

int env_int_dec(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

int num_sms_dec() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// Hypothesis for the bimodal 5-beam timings (profiles/r01_decode_summary.md section 3), UNTESTED: consecutive kernels of a
// step ask for different shared-memory sizes (25 ... 127 KB), so the driver may pick different L1 / shared carve-outs and
// has to drain the SM between them.  STLLM_DECODE_CARVEOUT=1 pins every decode kernel to the maximum shared carve-out
// (none of them relies on L1 hits).  Off by default until it has been A/B-ed on the device.
template <typename K>
cudaError_t dec_configure(K kernel, int max_dyn_smem) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem);
  if (e == cudaSuccess && env_int_dec("STLLM_DECODE_CARVEOUT", 0) != 0)
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  return e;
}

constexpr int DL_SMEM_MAX = 220 * 1024;
constexpr int DA_SMEM_MAX = 200 * 1024;

}  // namespace

extern "C" int stllm_decode_linear_f16(const void* x, long long ldx, const void* W, long long ldw, void* y,
                                       long long ldy, int R, int N, int K, int norm_mode, const void* norm_weight,
                                       float norm_eps, const void* residual, long long ldr, int epilogue,
                                       void* stream) {
  STLLM_CHECK_ARG(x && W && y && R >= 1 && R <= DL_RMAX && N >= 1 && K >= 32);
  STLLM_CHECK_ARG(norm_mode >= 0 && norm_mode <= 2 && (norm_mode != 2 || norm_weight != nullptr));
  STLLM_CHECK_ARG(epilogue == 0 || epilogue == 2);
  STLLM_CHECK_ARG(epilogue != 2 || N % 2 == 0);
  STLLM_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0 && ldx >= K && ldw >= K);
  STLLM_CHECK_ARG((((uintptr_t)x | (uintptr_t)W | (uintptr_t)norm_weight) & 15) == 0);
  if (K % 32 != 0) {
    stllm_set_last_error("stllm_decode_linear_f16: K=%d is not a multiple of 32", K);
    return STLLM_ERR_UNSUPPORTED;
  }
  DecLinParams p;
  p.x = (const __half*)x; p.ldx = ldx;
  p.W = (const __half*)W; p.ldw = ldw;
  p.y = (__half*)y; p.ldy = ldy;
  p.residual = (const __half*)residual; p.ldr = ldr;
  p.norm_w = (const __half*)norm_weight;
  p.R = R; p.N = N; p.K = K;
  p.Kp = K + (((K * 2) % 128 == 64) ? 0 : 32);
  p.norm_mode = norm_mode; p.epi = epilogue; p.eps = norm_eps;
  const int prefetch = env_int_dec("STLLM_DECODE_PREFETCH", 1);   // read per call: A/B inside one process
  p.prefetch = prefetch;
  const size_t part_bytes = (size_t)(2 * DL_WARPS * 128 + 16) * sizeof(float);
  size_t smem = (size_t)R * p.Kp * 2 + part_bytes;
  if (smem > (size_t)DL_SMEM_MAX) {
    stllm_set_last_error("stllm_decode_linear_f16: R=%d x K=%d activations do not fit in shared memory", R, K);
    return STLLM_ERR_UNSUPPORTED;
  }
  static bool configured = false;
  if (!configured) {
    STLLM_CHECK_CUDA(dec_configure(decode_linear_kernel<4, 2, false>, DL_SMEM_MAX / 2));
    STLLM_CHECK_CUDA(dec_configure(decode_linear_kernel<8, 1, false>, DL_SMEM_MAX));
    STLLM_CHECK_CUDA(dec_configure(decode_linear_kernel<4, 2, true>, 48 * 1024));
    configured = true;
  }
  const int direct_ok = env_int_dec("STLLM_DECODE_XDIRECT", 0);   // A/B: 3.49 (on) vs 3.37 ms/token (off) at 5 beams
  const int nblk = (N + 15) / 16;
  const bool two = smem <= (size_t)(DL_SMEM_MAX / 2);
  const bool direct = !two && norm_mode == 0 && direct_ok;
  int grid = num_sms_dec() * ((two || direct) ? 2 : 1);
  if (grid > nblk) grid = nblk;
  if (two)
    STLLM_CHECK_CUDA(stllm_launch(decode_linear_kernel<4, 2, false>, dim3(grid), dim3(DL_THREADS), smem, (cudaStream_t)stream, p));
  else if (direct) {
    p.prefetch = 0;   // the register preload spills in this variant
    STLLM_CHECK_CUDA(stllm_launch(decode_linear_kernel<4, 2, true>, dim3(grid), dim3(DL_THREADS), part_bytes, (cudaStream_t)stream, p));
  } else
    STLLM_CHECK_CUDA(stllm_launch(decode_linear_kernel<8, 1, false>, dim3(grid), dim3(DL_THREADS), smem, (cudaStream_t)stream, p));
  return STLLM_OK;
}

extern "C" int stllm_decode_rope_append_f16(void* qkv, long long ldq, const float* cos_t, const float* sin_t,
                                            int max_pos, int prefix_len, const int* n_gen, void* k_suffix,
                                            void* v_suffix, long long s_beam_stride, long long s_tok_stride, int l_max,
                                            int* src, int src_ld, int R, int H, int D, void* stream) {
  STLLM_CHECK_ARG(qkv && cos_t && sin_t && n_gen && k_suffix && v_suffix && src);
  STLLM_CHECK_ARG(R >= 1 && R <= DL_RMAX && H >= 1 && D % 16 == 0 && ldq >= 3LL * H * D && ldq % 8 == 0);
  STLLM_CHECK_ARG(s_beam_stride % 8 == 0 && s_tok_stride % 8 == 0 && s_tok_stride >= (long long)H * D && src_ld >= l_max);
  STLLM_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)k_suffix | (uintptr_t)v_suffix | (uintptr_t)cos_t | (uintptr_t)sin_t) & 15) == 0);
  DecRopeParams p;
  p.qkv = (__half*)qkv; p.ldq = ldq; p.cos_t = cos_t; p.sin_t = sin_t; p.n_gen = n_gen;
  p.ks = (__half*)k_suffix; p.vs = (__half*)v_suffix;
  p.s_beam_stride = s_beam_stride; p.s_tok_stride = s_tok_stride;
  p.src = src; p.src_ld = src_ld;
  p.prefix_len = prefix_len; p.max_pos = max_pos; p.l_max = l_max; p.R = R; p.H = H; p.D = D;
  const int total = R * H * (D / 16);
  STLLM_CHECK_CUDA(stllm_launch(decode_rope_append_kernel, dim3((total + 127) / 128), dim3(128), 0, (cudaStream_t)stream, p));
  return STLLM_OK;
}

extern "C" long long stllm_decode_attention_workspace_bytes(int H, int n_split) {
  if (H <= 0 || n_split <= 0) return 0;
  return (long long)H * n_split * DL_RMAX * DA_WS_ROW * (long long)sizeof(float);
}

extern "C" int stllm_decode_attention_f16(const void* q, long long ldq, const void* k_prefix, const void* v_prefix,
                                          long long p_stride, const void* prefix_mask, int prefix_len,
                                          const void* k_suffix, const void* v_suffix, long long s_beam_stride,
                                          long long s_tok_stride, int l_max, const int* src, int src_ld,
                                          const int* n_gen, void* out, long long ldo, void* workspace,
                                          long long workspace_bytes, int R, int H, int D, int n_split, float scale,
                                          void* stream) {
  STLLM_CHECK_ARG(q && k_suffix && v_suffix && src && n_gen && out && workspace);
  STLLM_CHECK_ARG(prefix_len == 0 || (k_prefix && v_prefix));
  STLLM_CHECK_ARG(R >= 1 && R <= DL_RMAX && H >= 1 && prefix_len >= 0 && l_max >= 1 && n_split >= 1 && n_split <= 64);
  STLLM_CHECK_ARG(ldq % 8 == 0 && p_stride % 8 == 0 && s_beam_stride % 8 == 0 && s_tok_stride % 8 == 0 && src_ld >= l_max);
  STLLM_CHECK_ARG((((uintptr_t)q | (uintptr_t)k_prefix | (uintptr_t)v_prefix | (uintptr_t)k_suffix | (uintptr_t)v_suffix) & 15) == 0);
  STLLM_CHECK_ARG(workspace_bytes >= stllm_decode_attention_workspace_bytes(H, n_split) && ((uintptr_t)workspace & 15) == 0);
  if (D != DA_D) {
    stllm_set_last_error("stllm_decode_attention_f16: head dim %d unsupported (128 only)", D);
    return STLLM_ERR_UNSUPPORTED;
  }
  DecAttnParams p;
  p.q = (const __half*)q; p.ldq = ldq;
  p.kp = (const __half*)k_prefix; p.vp = (const __half*)v_prefix; p.p_stride = p_stride;
  p.p_mask = (const uint8_t*)prefix_mask;
  p.ks = (const __half*)k_suffix; p.vs = (const __half*)v_suffix;
  p.s_beam_stride = s_beam_stride; p.s_tok_stride = s_tok_stride;
  p.src = src; p.src_ld = src_ld; p.n_gen = n_gen; p.ws = (float*)workspace;
  p.R = R; p.H = H; p.Sp = prefix_len; p.NS = n_split;
  p.cap = (((prefix_len + n_split - 1) / n_split + (l_max + n_split - 1) / n_split + 2) + 3) & ~3;
  p.scale = scale;
  const size_t smem = ((size_t)DL_RMAX * p.cap + (size_t)DA_WARPS * (R + 1) * DA_D) * sizeof(float);
  if (smem > (size_t)DA_SMEM_MAX) {
    stllm_set_last_error("stllm_decode_attention_f16: %d + %d keys over %d splits do not fit in shared memory",
                         prefix_len, l_max, n_split);
    return STLLM_ERR_UNSUPPORTED;
  }
  static bool configured = false;
  if (!configured) {
    STLLM_CHECK_CUDA(dec_configure(decode_attention_kernel, DA_SMEM_MAX));
    configured = true;
  }
  STLLM_CHECK_CUDA(stllm_launch(decode_attention_kernel, dim3(H, n_split), dim3(DA_THREADS), smem, (cudaStream_t)stream, p));
  STLLM_CHECK_CUDA(stllm_launch(decode_attention_combine_kernel, dim3(H, R), dim3(DA_D), 0, (cudaStream_t)stream,
                                (const float*)workspace, (__half*)out, ldo, n_split));
  return STLLM_OK;
}

extern "C" long long stllm_decode_attention_fused_workspace_bytes(int H, int n_split) {
  if (H <= 0 || n_split <= 0) return 0;
  return stllm_decode_attention_workspace_bytes(H, n_split) + (((long long)H * (long long)sizeof(int) + 15) & ~15LL);
}

extern "C" int stllm_decode_attention_fused_f16(void* qkv, long long ldq, const float* cos_t, const float* sin_t,
                                                int max_pos, const void* k_prefix, const void* v_prefix,
                                                long long p_stride, const void* prefix_mask, int prefix_len,
                                                void* k_suffix, void* v_suffix, long long s_beam_stride,
                                                long long s_tok_stride, int l_max, int* src, int src_ld,
                                                const int* n_gen, void* out, long long ldo, void* workspace,
                                                long long workspace_bytes, int R, int H, int D, int n_split,
                                                float scale, const void* prefetch_ptr, long long prefetch_bytes,
                                                const void* prefetch_ptr2, long long prefetch_bytes2, void* stream) {
  STLLM_CHECK_ARG(qkv && cos_t && sin_t && k_suffix && v_suffix && src && n_gen && out && workspace);
  STLLM_CHECK_ARG(prefetch_bytes >= 0 && (((uintptr_t)prefetch_ptr) & 15) == 0);
  STLLM_CHECK_ARG(prefetch_bytes2 >= 0 && (((uintptr_t)prefetch_ptr2) & 15) == 0);
  STLLM_CHECK_ARG(prefix_len == 0 || (k_prefix && v_prefix));
  STLLM_CHECK_ARG(R >= 1 && R <= DL_RMAX && H >= 1 && prefix_len >= 0 && l_max >= 1 && n_split >= 1 && n_split <= 64);
  STLLM_CHECK_ARG(ldq % 8 == 0 && ldq >= 3LL * H * D && p_stride % 8 == 0 && s_beam_stride % 8 == 0 &&
                  s_tok_stride % 8 == 0 && src_ld >= l_max);
  STLLM_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)k_prefix | (uintptr_t)v_prefix | (uintptr_t)k_suffix | (uintptr_t)v_suffix) & 15) == 0);
  STLLM_CHECK_ARG(workspace_bytes >= stllm_decode_attention_fused_workspace_bytes(H, n_split) && ((uintptr_t)workspace & 15) == 0);
  if (D != DA_D) {
    stllm_set_last_error("stllm_decode_attention_fused_f16: head dim %d unsupported (128 only)", D);
    return STLLM_ERR_UNSUPPORTED;
  }
  DecAttnFusedParams p;
  p.qkv = (__half*)qkv; p.ldq = ldq; p.cos_t = cos_t; p.sin_t = sin_t;
  p.kp = (const __half*)k_prefix; p.vp = (const __half*)v_prefix; p.p_stride = p_stride;
  p.p_mask = (const uint8_t*)prefix_mask;
  p.ks = (__half*)k_suffix; p.vs = (__half*)v_suffix;
  p.s_beam_stride = s_beam_stride; p.s_tok_stride = s_tok_stride;
  p.src = src; p.src_ld = src_ld; p.n_gen = n_gen; p.out = (__half*)out; p.ldo = ldo;
  p.ws = (float*)workspace;
  p.tickets = (int*)((char*)workspace + stllm_decode_attention_workspace_bytes(H, n_split));
  p.R = R; p.H = H; p.Sp = prefix_len; p.NS = n_split; p.l_max = l_max; p.max_pos = max_pos;
  p.cap = (((prefix_len + n_split - 1) / n_split + (l_max + n_split - 1) / n_split + 2) + 3) & ~3;
  p.scale = scale;
  p.pf_ptr[0] = prefetch_bytes > 0 ? (const uint8_t*)prefetch_ptr : nullptr;
  p.pf_bytes[0] = prefetch_bytes;
  p.pf_ptr[1] = prefetch_bytes2 > 0 ? (const uint8_t*)prefetch_ptr2 : nullptr;
  p.pf_bytes[1] = prefetch_bytes2;
  const size_t smem = ((size_t)DL_RMAX * p.cap + (size_t)DA_WARPS * (R + 1) * DA_D + 3 * DL_RMAX * DA_D) * sizeof(float) +
                      (size_t)DL_RMAX * DA_D * sizeof(__half);
  if (smem > (size_t)DA_SMEM_MAX) {
    stllm_set_last_error("stllm_decode_attention_fused_f16: %d + %d keys over %d splits do not fit in shared memory",
                         prefix_len, l_max, n_split);
    return STLLM_ERR_UNSUPPORTED;
  }
  static bool configured = false;
  if (!configured) {
    STLLM_CHECK_CUDA(dec_configure(decode_attention_fused_kernel, DA_SMEM_MAX));
    configured = true;
  }
  STLLM_CHECK_CUDA(stllm_launch(decode_attention_fused_kernel, dim3(H, n_split), dim3(DA_THREADS), smem, (cudaStream_t)stream, p));
  return STLLM_OK;
}
