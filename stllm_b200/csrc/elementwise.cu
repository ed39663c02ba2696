// HBM-bound data-movement kernels of the video-token path (coalesced, 16-byte vectors where the layout allows).
//   im2col            eva_vit.py:196,203   Conv2d(3, D, k=14, s=14) == GEMM over 14x14x3 patches
//   vit_assemble      eva_vit.py:328-332   [cls ; patches] + pos_embed
//   rope              modeling_llama_mem.py:81-127  rotate-half RoPE on q,k (positions 0..S-1)
//   gather_rows       st_llm.py:491 (img_embeds[~mask]), nn.Embedding lookups (st_llm.py:392-393,510)
//   mean_frames       st_llm.py:466,470   mean over T
//   residual_merge    st_llm.py:472-476   local[:, idx] + global (global-local long-video path)
//   cosine_loss       st_llm.py:89-91     MVM loss  mean(2 - 2 cos)
//   cross_entropy     st_llm.py:124-138   shifted CE, ignore_index -100
#include "common.cuh"

namespace {

template <typename TIN>
__global__ void im2col_patch_kernel(const TIN* __restrict__ img, __half* __restrict__ out, int n_img, int C, int H,
                                    int W, int P, int kpad) {
  pdl_launch_dependents();
  pdl_wait();
  const int gw = W / P, gh = H / P;
  const int kreal = C * P * P;
  const long long total = (long long)n_img * gh * gw * kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % kpad);
    const long long rowi = i / kpad;
    float v = 0.f;
    if (col < kreal) {
      const int px = col % P, py = (col / P) % P, c = col / (P * P);
      const int gx = (int)(rowi % gw), gy = (int)((rowi / gw) % gh);
      const long long n = rowi / ((long long)gw * gh);
      v = (float)img[((n * C + c) * H + (gy * P + py)) * (long long)W + gx * P + px];
    }
    out[i] = __float2half_rn(v);
  }
}

// x[n,0,:] = cls + pos[0];  x[n,1+p,:] = patches[n*NP+p,:] + pos[1+p]   (fp16 add, one rounding)
__global__ void vit_assemble_kernel(const __half* __restrict__ patches, long long ldp, const __half* __restrict__ cls,
                                    const __half* __restrict__ pos, __half* __restrict__ x, int n_img, int np,
                                    int dim) {
  pdl_launch_dependents();
  pdl_wait();
  const int dv = dim >> 3;
  const long long total = (long long)n_img * (np + 1) * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv);
    const long long tok = i / dv;
    const int t = (int)(tok % (np + 1));
    const long long n = tok / (np + 1);
    uint4 a = (t == 0) ? __ldg(reinterpret_cast<const uint4*>(cls) + c)
                       : __ldg(reinterpret_cast<const uint4*>(patches + (n * np + t - 1) * ldp) + c);
    uint4 b = __ldg(reinterpret_cast<const uint4*>(pos + (long long)t * dim) + c);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __hadd2(ah[j], bh[j]);
    reinterpret_cast<uint4*>(x + tok * dim)[c] = o;
  }
}

// in-place rotate-half RoPE on `nheads` consecutive heads of each row: x[row, h*D + j], j in [0,D)
// cos/sin: fp32 tables [max_pos, D/2]; pos[row] int32 position ids
__global__ void rope_kernel(__half* __restrict__ x, long long ldx, const int* __restrict__ pos,
                            const float* __restrict__ cos_t, const float* __restrict__ sin_t, int rows, int nheads,
                            int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int half_d = D >> 1;
  const int groups = half_d >> 3;                                   // 8 rotation pairs (16 bytes + 16 bytes) per thread
  const long long total = (long long)rows * nheads * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % groups) * 8;
    const long long rh = i / groups;
    const int h = (int)(rh % nheads);
    const long long row = rh / nheads;
    const int ppos = pos[row];
    __half* base = x + row * ldx + (long long)h * D;
    const uint4 ulo = *reinterpret_cast<const uint4*>(base + j);
    const uint4 uhi = *reinterpret_cast<const uint4*>(base + j + half_d);
    const float4* cp = reinterpret_cast<const float4*>(cos_t + (long long)ppos * half_d + j);
    const float4* sp = reinterpret_cast<const float4*>(sin_t + (long long)ppos * half_d + j);
    const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
    const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const __half2* hlo = reinterpret_cast<const __half2*>(&ulo);
    const __half2* hhi = reinterpret_cast<const __half2*>(&uhi);
    uint4 olo, ohi;
    uint32_t* plo = reinterpret_cast<uint32_t*>(&olo);
    uint32_t* phi = reinterpret_cast<uint32_t*>(&ohi);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 lo = __half22float2(hlo[q]);
      const float2 hi = __half22float2(hhi[q]);
      // out_lo = lo*cos - hi*sin ; out_hi = hi*cos + lo*sin   (cos/sin identical for j and j + D/2)
      plo[q] = pack_half2(lo.x * c[2 * q] - hi.x * sn[2 * q], lo.y * c[2 * q + 1] - hi.y * sn[2 * q + 1]);
      phi[q] = pack_half2(hi.x * c[2 * q] + lo.x * sn[2 * q], hi.y * c[2 * q + 1] + lo.y * sn[2 * q + 1]);
    }
    *reinterpret_cast<uint4*>(base + j) = olo;
    *reinterpret_cast<uint4*>(base + j + half_d) = ohi;
  }
}

__global__ void gather_rows_kernel(const __half* __restrict__ src, long long lds, const long long* __restrict__ idx,
                                   __half* __restrict__ dst, long long ldd, int n_rows, int cols, long long src_rows,
                                   const __half* __restrict__ addend, int add_period, int* __restrict__ err) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = cols >> 3;
  const long long total = (long long)n_rows * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long r = i / cv;
    const long long s = idx[r];
    if (s < 0 || s >= src_rows) {
      if (err != nullptr) *err = 1;
      continue;
    }
    uint4 v = __ldg(reinterpret_cast<const uint4*>(src + s * lds) + c);
    if (addend != nullptr) {  // dst[r] = src[idx[r]] + addend[r % add_period]   (BertEmbeddings word + position)
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(addend + (r % add_period) * (long long)cols) + c);
      __half2* vh = reinterpret_cast<__half2*>(&v);
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
      for (int j = 0; j < 4; ++j) vh[j] = __hadd2(vh[j], ah[j]);
    }
    reinterpret_cast<uint4*>(dst + r * ldd)[c] = v;
  }
}

// dst[r] = idx[r] >= 0 ? src0[idx[r]] : src1[-1 - idx[r]]   -- the LLM input of a batch in ONE launch: token-embedding
// rows and video-token rows selected by the host-built row plan (stllm_b200/prompt_assembly.py; st_llm.py:379-432,519-530)
__global__ void gather_rows2_kernel(const __half* __restrict__ src0, long long lds0, long long rows0,
                                    const __half* __restrict__ src1, long long lds1, long long rows1,
                                    const long long* __restrict__ idx, __half* __restrict__ dst, long long ldd, int n_rows,
                                    int cols, int* __restrict__ err) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = cols >> 3;
  const long long total = (long long)n_rows * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long r = i / cv;
    const long long s = idx[r];
    const long long s1 = -1 - s;
    const bool first = s >= 0;
    if (first ? (s >= rows0) : (s1 >= rows1)) {
      if (err != nullptr) *err = 1;
      continue;
    }
    const __half* row = first ? src0 + s * lds0 : src1 + s1 * lds1;
    reinterpret_cast<uint4*>(dst + r * ldd)[c] = __ldg(reinterpret_cast<const uint4*>(row) + c);
  }
}

// x: [B, T, inner] -> y: [B, inner], mean over T with fp32 accumulation (torch.mean on fp16 does the same)
__global__ void mean_frames_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int T,
                                   long long inner) {
  pdl_launch_dependents();
  pdl_wait();
  const long long iv = inner >> 1;
  const long long total = (long long)B * iv;
  const float inv = 1.f / (float)T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i % iv, b = i / iv;
    float2 acc = make_float2(0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(x + ((long long)b * T + t) * inner + 2 * c));
      acc.x += f.x;
      acc.y += f.y;
    }
    *reinterpret_cast<__half2*>(y + b * inner + 2 * c) = __floats2half2_rn(acc.x * inv, acc.y * inv);
  }
}

// out[b, r, :] = x[b, idx[r], :] + g[b, :]      x: [B, T, inner], g: [B, inner], out: [B, R, inner]
__global__ void residual_merge_kernel(const __half* __restrict__ x, const __half* __restrict__ g,
                                      const long long* __restrict__ idx, __half* __restrict__ out, int B, int T,
                                      int R, long long inner) {
  pdl_launch_dependents();
  pdl_wait();
  const long long iv = inner >> 3;
  const long long total = (long long)B * R * iv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i % iv;
    const long long br = i / iv;
    const int r = (int)(br % R);
    const long long b = br / R;
    const long long t = idx[r];
    uint4 a = __ldg(reinterpret_cast<const uint4*>(x + (b * T + t) * inner) + c);
    uint4 gg = __ldg(reinterpret_cast<const uint4*>(g + b * inner) + c);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* gh = reinterpret_cast<const __half2*>(&gg);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __hadd2(ah[j], gh[j]);
    reinterpret_cast<uint4*>(out + br * inner)[c] = o;
  }
}

// per row: 2 - 2 * <a/|a|, b/|b|>; accumulated (sum) into out[0]; host divides by rows. One warp per row.
__global__ void cosine_loss_kernel(const __half* __restrict__ a, long long lda, const __half* __restrict__ b,
                                   long long ldb, int rows, int cols, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __half* ar = a + (long long)row * lda;
  const __half* br = b + (long long)row * ldb;
  float aa = 0.f, bb = 0.f, ab = 0.f;
  for (int i = lane * 2; i < cols; i += 64) {
    const float2 x = __half22float2(*reinterpret_cast<const __half2*>(ar + i));
    const float2 y = __half22float2(*reinterpret_cast<const __half2*>(br + i));
    aa += x.x * x.x + x.y * x.y;
    bb += y.x * y.x + y.y * y.y;
    ab += x.x * y.x + x.y * y.y;
  }
  aa = warp_sum(aa);
  bb = warp_sum(bb);
  ab = warp_sum(ab);
  if (lane == 0) atomicAdd(out, (2.f - 2.f * ab * rsqrtf(aa) * rsqrtf(bb)) / (float)rows);
}

// per row: logsumexp(logits) - logits[label] for label != ignore; out[0] += loss, out[1] += count
__global__ void cross_entropy_kernel(const __half* __restrict__ logits, long long ldl,
                                     const long long* __restrict__ labels, int rows, int V, long long ignore_index,
                                     float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  if (row >= rows) return;
  const long long lab = labels[row];
  if (lab == ignore_index || lab < 0 || lab >= V) return;
  const __half* lr = logits + (long long)row * ldl;
  __shared__ float red[32];
  __shared__ float bcast;
  const int nwarps = blockDim.x >> 5;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) m = fmaxf(m, __half2float(lr[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float x = (threadIdx.x < nwarps) ? red[threadIdx.x] : -INFINITY;
    x = warp_max(x);
    if (threadIdx.x == 0) bcast = x;
  }
  __syncthreads();
  m = bcast;
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) s += __expf(__half2float(lr[i]) - m);
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < nwarps) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      atomicAdd(out, logf(s) + m - __half2float(lr[lab]));
      atomicAdd(out + 1, 1.f);
    }
  }
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ------------------------------------------------------------------------------------------------
// BT-Adapter layout changes (eva_btadapter.py).  The ViT stream x is frame-major: row (b*T + t)*(P+1) + l (l = 0 the
// cls token).  The adapter branch z is patch-major:  rows [0, B) = one cls token per clip, then row B + (b*P + l)*T + t
// = patch l of frame t  ('b (p t) m', eva_btadapter.py:188,220), so the temporal attention of one patch position is a
// contiguous run of T rows.  Every row is `dim` contiguous fp16 values: the transposes move whole rows (coalesced).
// ------------------------------------------------------------------------------------------------
// z_out = T(x) + z_prev          (forward_branch :184-193 and the residual of BTAdapter_Spatial :273-276)
//   patch rows: z_out = x_row + z_prev_row                         cls rows: z_out = fp16(mean_t x_cls) + z_prev_row
// init (z_prev == nullptr, init_input :203-226):
//   patch rows: z_out = fp16(fp16(x_row + pos[1+l]) + time[t])     cls rows: (fp16(mean_t x_cls) + fp16(bt_cls + pos[0])) / 2
__global__ void bt_vit_to_branch_kernel(const __half* __restrict__ x, const __half* z_prev, __half* z_out,
                                        const __half* __restrict__ pos, const __half* __restrict__ time_emb,
                                        const __half* __restrict__ bt_cls, int B, int T, int P, int dim) {
  pdl_launch_dependents();
  pdl_wait();
  const int dv = dim >> 3;
  const long long rows = (long long)B * (1 + (long long)P * T);
  const long long total = rows * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv);
    const long long zr = i / dv;
    float v[8];
    if (zr < B) {
      // cls row of clip b: mean over the T frames of the ViT cls tokens (fp32 accumulate, one fp16 rounding)
      const long long b = zr;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int t = 0; t < T; ++t) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + ((b * T + t) * (long long)(P + 1)) * dim) + c);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __half2float(__float2half_rn(acc[j] / (float)T));
      if (z_prev != nullptr) {
        const uint4 u = *(reinterpret_cast<const uint4*>(z_prev + zr * dim) + c);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          v[2 * j] += f.x;
          v[2 * j + 1] += f.y;
        }
      } else {
        const uint4 uc = __ldg(reinterpret_cast<const uint4*>(bt_cls) + c);
        const uint4 up = __ldg(reinterpret_cast<const uint4*>(pos) + c);
        const __half2* hc = reinterpret_cast<const __half2*>(&uc);
        const __half2* hp = reinterpret_cast<const __half2*>(&up);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(__hadd2(hc[j], hp[j]));     // cls_branch = fp16(BTAdapter_cls + pos[0])
          v[2 * j] = __half2float(__float2half_rn(v[2 * j] + f.x)) * 0.5f;
          v[2 * j + 1] = __half2float(__float2half_rn(v[2 * j + 1] + f.y)) * 0.5f;
        }
      }
    } else {
      const long long pr = zr - B;
      const int t = (int)(pr % T);
      const int l = (int)((pr / T) % P);
      const long long b = pr / ((long long)T * P);
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + ((b * T + t) * (long long)(P + 1) + 1 + l) * dim) + c);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
      }
      if (z_prev != nullptr) {
        const uint4 uz = *(reinterpret_cast<const uint4*>(z_prev + zr * dim) + c);
        const __half2* hz = reinterpret_cast<const __half2*>(&uz);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(hz[j]);
          v[2 * j] += f.x;
          v[2 * j + 1] += f.y;
        }
      } else {
        const uint4 up = __ldg(reinterpret_cast<const uint4*>(pos + (long long)(1 + l) * dim) + c);
        const uint4 ut = __ldg(reinterpret_cast<const uint4*>(time_emb + (long long)t * dim) + c);
        const __half2* hp = reinterpret_cast<const __half2*>(&up);
        const __half2* ht = reinterpret_cast<const __half2*>(&ut);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fp = __half22float2(hp[j]);
          const float2 ft = __half22float2(ht[j]);
          v[2 * j] = __half2float(__float2half_rn(v[2 * j] + fp.x)) + ft.x;
          v[2 * j + 1] = __half2float(__float2half_rn(v[2 * j + 1] + fp.y)) + ft.y;
        }
      }
    }
    uint4 o;
    o.x = pack_half2(v[0], v[1]);
    o.y = pack_half2(v[2], v[3]);
    o.z = pack_half2(v[4], v[5]);
    o.w = pack_half2(v[6], v[7]);
    reinterpret_cast<uint4*>(z_out + zr * dim)[c] = o;
  }
}

// out = (x + T^-1(z)) / 2   (forward_features :178-181): patch rows take their branch row, every frame's cls row takes
// the clip's branch cls row.  out may alias x.
__global__ void bt_branch_to_vit_avg_kernel(const __half* x, const __half* __restrict__ z, __half* out, int B, int T,
                                            int P, int dim) {
  pdl_launch_dependents();
  pdl_wait();
  const int dv = dim >> 3;
  const long long rows = (long long)B * T * (P + 1);
  const long long total = rows * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv);
    const long long xr = i / dv;
    const int l1 = (int)(xr % (P + 1));
    const long long bt = xr / (P + 1);
    const int t = (int)(bt % T);
    const long long b = bt / T;
    const long long zr = (l1 == 0) ? b : (long long)B + (b * P + (l1 - 1)) * T + t;
    const uint4 ux = *(reinterpret_cast<const uint4*>(x + xr * dim) + c);
    const uint4 uz = __ldg(reinterpret_cast<const uint4*>(z + zr * dim) + c);
    const __half2* hx = reinterpret_cast<const __half2*>(&ux);
    const __half2* hz = reinterpret_cast<const __half2*>(&uz);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __hmul2(__hadd2(hx[j], hz[j]), __float2half2_rn(0.5f));
    reinterpret_cast<uint4*>(out + xr * dim)[c] = o;
  }
}

}  // namespace

extern "C" int stllm_bt_vit_to_branch_f16(const void* x, const void* z_prev, void* z_out, const void* pos,
                                          const void* time_emb, const void* bt_cls, int B, int T, int P, int dim,
                                          void* stream) {
  STLLM_CHECK_ARG(x && z_out && B > 0 && T > 0 && P > 0 && dim > 0 && dim % 8 == 0);
  STLLM_CHECK_ARG(z_prev != nullptr || (pos != nullptr && time_emb != nullptr && bt_cls != nullptr));
  const long long total = (long long)B * (1 + (long long)P * T) * (dim / 8);
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  STLLM_CHECK_CUDA(stllm_launch(bt_vit_to_branch_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream,
                                (const __half*)x, (const __half*)z_prev, (__half*)z_out, (const __half*)pos,
                                (const __half*)time_emb, (const __half*)bt_cls, B, T, P, dim));
  return STLLM_OK;
}

extern "C" int stllm_bt_branch_to_vit_avg_f16(const void* x, const void* z, void* out, int B, int T, int P, int dim,
                                              void* stream) {
  STLLM_CHECK_ARG(x && z && out && B > 0 && T > 0 && P > 0 && dim > 0 && dim % 8 == 0);
  const long long total = (long long)B * T * (P + 1) * (dim / 8);
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  STLLM_CHECK_CUDA(stllm_launch(bt_branch_to_vit_avg_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream,
                                (const __half*)x, (const __half*)z, (__half*)out, B, T, P, dim));
  return STLLM_OK;
}

extern "C" int stllm_im2col_patch(const void* img, int img_is_f32, void* out, int n_img, int C, int H, int W, int P,
                                  int kpad, void* stream) {
  STLLM_CHECK_ARG(img && out && n_img > 0 && C > 0 && P > 0 && H % P == 0 && W % P == 0 && kpad >= C * P * P);
  const long long total = (long long)n_img * (H / P) * (W / P) * kpad;
  const int grid = grid_for(total, 256);
  if (img_is_f32)
    STLLM_CHECK_CUDA(stllm_launch(im2col_patch_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)img, (__half*)out, n_img, C, H, W, P, kpad));
  else
    STLLM_CHECK_CUDA(stllm_launch(im2col_patch_kernel<__half>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __half*)img, (__half*)out, n_img, C, H, W, P, kpad));
    return STLLM_OK;
}

extern "C" int stllm_vit_assemble_f16(const void* patches, long long ldp, const void* cls, const void* pos, void* x,
                                      int n_img, int n_patches, int dim, void* stream) {
  STLLM_CHECK_ARG(patches && cls && pos && x && n_img > 0 && n_patches > 0 && dim % 8 == 0 && ldp % 8 == 0);
  STLLM_CHECK_ARG((((uintptr_t)patches | (uintptr_t)cls | (uintptr_t)pos | (uintptr_t)x) & 15) == 0);
  const long long total = (long long)n_img * (n_patches + 1) * (dim / 8);
  STLLM_CHECK_CUDA(stllm_launch(vit_assemble_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const __half*)patches, ldp, (const __half*)cls, (const __half*)pos, (__half*)x, n_img, n_patches, dim));
    return STLLM_OK;
}

extern "C" int stllm_rope_f16(void* x, long long ldx, const int* pos, const float* cos_t, const float* sin_t, int rows,
                              int nheads, int head_dim, int max_pos, void* stream) {
  STLLM_CHECK_ARG(x && pos && cos_t && sin_t && rows > 0 && nheads > 0 && head_dim % 16 == 0 && ldx % 8 == 0);
  STLLM_CHECK_ARG(max_pos > 0 && (((uintptr_t)x) & 15) == 0 && (((uintptr_t)cos_t | (uintptr_t)sin_t) & 15) == 0);
  const long long total = (long long)rows * nheads * (head_dim / 16);
  STLLM_CHECK_CUDA(stllm_launch(rope_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (__half*)x, ldx, pos, cos_t, sin_t, rows, nheads,
                                                                      head_dim));
    return STLLM_OK;
}

extern "C" int stllm_gather_rows_f16(const void* src, long long lds, long long src_rows, const long long* idx, void* dst,
                                     long long ldd, int n_rows, int cols, const void* addend, int add_period,
                                     int* err_flag, void* stream) {
  STLLM_CHECK_ARG(src && idx && dst && n_rows > 0 && cols > 0 && cols % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0);
  STLLM_CHECK_ARG((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)addend) & 15) == 0);
  STLLM_CHECK_ARG(addend == nullptr || add_period > 0);
  const long long total = (long long)n_rows * (cols / 8);
  STLLM_CHECK_CUDA(stllm_launch(gather_rows_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const __half*)src, lds, idx, (__half*)dst, ldd, n_rows, cols, src_rows, (const __half*)addend, add_period, err_flag));
    return STLLM_OK;
}

extern "C" int stllm_gather_rows2_f16(const void* src0, long long lds0, long long rows0, const void* src1, long long lds1,
                                      long long rows1, const long long* idx, void* dst, long long ldd, int n_rows, int cols,
                                      int* err_flag, void* stream) {
  STLLM_CHECK_ARG(src0 && src1 && idx && dst && n_rows > 0 && cols > 0 && cols % 8 == 0);
  STLLM_CHECK_ARG(lds0 % 8 == 0 && lds1 % 8 == 0 && ldd % 8 == 0 && rows0 >= 0 && rows1 >= 0);
  STLLM_CHECK_ARG((((uintptr_t)src0 | (uintptr_t)src1 | (uintptr_t)dst) & 15) == 0);
  const long long total = (long long)n_rows * (cols / 8);
  STLLM_CHECK_CUDA(stllm_launch(gather_rows2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream,
                                (const __half*)src0, lds0, rows0, (const __half*)src1, lds1, rows1, idx, (__half*)dst, ldd,
                                n_rows, cols, err_flag));
  return STLLM_OK;
}

extern "C" int stllm_mean_frames_f16(const void* x, void* y, int B, int T, long long inner, void* stream) {
  STLLM_CHECK_ARG(x && y && B > 0 && T > 0 && inner > 0 && inner % 2 == 0);
  const long long total = (long long)B * (inner / 2);
  STLLM_CHECK_CUDA(stllm_launch(mean_frames_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const __half*)x, (__half*)y, B, T, inner));
    return STLLM_OK;
}

extern "C" int stllm_residual_merge_f16(const void* x, const void* g, const long long* idx, void* out, int B, int T,
                                        int R, long long inner, void* stream) {
  STLLM_CHECK_ARG(x && g && idx && out && B > 0 && T > 0 && R > 0 && inner > 0 && inner % 8 == 0);
  STLLM_CHECK_ARG((((uintptr_t)x | (uintptr_t)g | (uintptr_t)out) & 15) == 0);
  const long long total = (long long)B * R * (inner / 8);
  STLLM_CHECK_CUDA(stllm_launch(residual_merge_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const __half*)x, (const __half*)g, idx, (__half*)out, B, T, R, inner));
    return STLLM_OK;
}

extern "C" int stllm_cosine_loss_f16(const void* a, long long lda, const void* b, long long ldb, int rows, int cols,
                                     float* out_sum, void* stream) {
  STLLM_CHECK_ARG(a && b && out_sum && rows > 0 && cols > 0 && cols % 2 == 0 && lda % 2 == 0 && ldb % 2 == 0);
  STLLM_CHECK_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(float), (cudaStream_t)stream));
  STLLM_CHECK_CUDA(stllm_launch(cosine_loss_kernel, dim3((rows + 3) / 4), dim3(128), 0, (cudaStream_t)stream, (const __half*)a, lda, (const __half*)b, ldb,
                                                                       rows, cols, out_sum));
    return STLLM_OK;
}

extern "C" int stllm_cross_entropy_f16(const void* logits, long long ldl, const long long* labels, int rows, int V,
                                       long long ignore_index, float* out_sum_count, void* stream) {
  STLLM_CHECK_ARG(logits && labels && out_sum_count && rows > 0 && V > 0);
  STLLM_CHECK_CUDA(cudaMemsetAsync(out_sum_count, 0, 2 * sizeof(float), (cudaStream_t)stream));
  STLLM_CHECK_CUDA(stllm_launch(cross_entropy_kernel, dim3(rows), dim3(256), 0, (cudaStream_t)stream, (const __half*)logits, ldl, labels, rows, V,
                                                               ignore_index, out_sum_count));
    return STLLM_OK;
}
