// Row normalisations with fp32 statistics (HBM-bound; one warp per row, 16-byte vector loads).
//   LayerNorm : eva_vit.py:157,163 (eps 1e-6), blip2.py:103-109 ln_vision (fp32 LN, eps 1e-5),
//               Qformer.py:65,282,366 (post-LN eps 1e-12), st_llm.py:39 (MVM head LN)
//   RMSNorm   : modeling_llama_mem.py:61-78 (fp32 variance, weight * x.to(fp16))
#include "common.cuh"

namespace {

constexpr int ROWS_PER_CTA = 4;  // 4 warps, one row each

// y = (x - mean) * rstd * gamma + beta ; statistics in fp32 over the fp16 input; output rounded once.
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
layernorm_f16_kernel(const __half* __restrict__ x, long long ldx, __half* __restrict__ y, long long ldy,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int cols,
                     float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __half* xr = x + (long long)row * ldx;
  __half* yr = y + (long long)row * ldy;
  const int nvec = cols >> 3;
  // pass 1: mean
  float sum = 0.f;
  for (int i = lane; i < nvec; i += 32) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(xr) + i);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      sum += f.x + f.y;
    }
  }
  for (int i = (nvec << 3) + lane; i < cols; i += 32) sum += __half2float(xr[i]);
  const float mean = warp_sum(sum) / (float)cols;
  // pass 2: variance (two-pass: matches torch's numerics better than E[x^2]-mean^2; row is L1/L2 hot)
  float sq = 0.f;
  for (int i = lane; i < nvec; i += 32) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(xr) + i);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      const float a = f.x - mean, b = f.y - mean;
      sq += a * a + b * b;
    }
  }
  for (int i = (nvec << 3) + lane; i < cols; i += 32) {
    const float a = __half2float(xr[i]) - mean;
    sq += a * a;
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)cols + eps);
  // pass 3: normalise
  for (int i = lane; i < nvec; i += 32) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(xr) + i);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * i);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * i + 1);
    const float gs[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      o[2 * j] = (f.x - mean) * rstd * gs[2 * j] + bs[2 * j];
      o[2 * j + 1] = (f.y - mean) * rstd * gs[2 * j + 1] + bs[2 * j + 1];
    }
    uint4 w;
    w.x = pack_half2(o[0], o[1]);
    w.y = pack_half2(o[2], o[3]);
    w.z = pack_half2(o[4], o[5]);
    w.w = pack_half2(o[6], o[7]);
    reinterpret_cast<uint4*>(yr)[i] = w;
  }
  for (int i = (nvec << 3) + lane; i < cols; i += 32)
    yr[i] = __float2half_rn((__half2float(xr[i]) - mean) * rstd * gamma[i] + beta[i]);
}

// Llama RMSNorm: h = fp16(x * rsqrt(mean(x^2) + eps)); y = w * h (fp16 multiply, rounded)
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
rmsnorm_f16_kernel(const __half* __restrict__ x, long long ldx, __half* __restrict__ y, long long ldy,
                   const __half* __restrict__ w, int rows, int cols, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __half* xr = x + (long long)row * ldx;
  __half* yr = y + (long long)row * ldy;
  const int nvec = cols >> 3;
  float sq = 0.f;
  for (int i = lane; i < nvec; i += 32) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(xr) + i);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      sq += f.x * f.x + f.y * f.y;
    }
  }
  for (int i = (nvec << 3) + lane; i < cols; i += 32) {
    const float a = __half2float(xr[i]);
    sq += a * a;
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)cols + eps);
  for (int i = lane; i < nvec; i += 32) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(xr) + i);
    uint4 wv = __ldg(reinterpret_cast<const uint4*>(w) + i);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    const __half2* wh = reinterpret_cast<const __half2*>(&wv);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      const __half2 n = __floats2half2_rn(f.x * rstd, f.y * rstd);
      oh[j] = __hmul2(wh[j], n);
    }
    reinterpret_cast<uint4*>(yr)[i] = o;
  }
  for (int i = (nvec << 3) + lane; i < cols; i += 32)
    yr[i] = __hmul(w[i], __float2half_rn(__half2float(xr[i]) * rstd));
}


// Register-cached LayerNorm: one warp per row, the row (<= 32 * MAXV * 8 elements) is read from global ONCE.
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_f16_regs_kernel(const __half* __restrict__ x, long long ldx, __half* __restrict__ y, long long ldy,
                          const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int cols,
                          float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  uint4* yr = reinterpret_cast<uint4*>(y + (long long)row * ldy);
  const int nvec = cols >> 3;
  float v[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 u = __ldg(xr + vi);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        v[i][2 * j] = f.x;
        v[i][2 * j + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
  const float mean = warp_sum(sum) / (float)cols;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = v[i][j] - mean;
        sq += a * a;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)cols + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi + 1);
      uint4 w;
      w.x = pack_half2((v[i][0] - mean) * rstd * g0.x + b0.x, (v[i][1] - mean) * rstd * g0.y + b0.y);
      w.y = pack_half2((v[i][2] - mean) * rstd * g0.z + b0.z, (v[i][3] - mean) * rstd * g0.w + b0.w);
      w.z = pack_half2((v[i][4] - mean) * rstd * g1.x + b1.x, (v[i][5] - mean) * rstd * g1.y + b1.y);
      w.w = pack_half2((v[i][6] - mean) * rstd * g1.z + b1.z, (v[i][7] - mean) * rstd * g1.w + b1.w);
      yr[vi] = w;
    }
  }
}

// Block-per-row RMSNorm for wide rows (Llama hidden 4096): 256 threads x VPT 16-byte vectors, one global read.
template <int VPT>
__global__ void __launch_bounds__(256)
rmsnorm_f16_block_kernel(const __half* __restrict__ x, long long ldx, __half* __restrict__ y, long long ldy,
                         const __half* __restrict__ w, int rows, int cols, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  uint4* yr = reinterpret_cast<uint4*>(y + (long long)row * ldy);
  const int nvec = cols >> 3;
  __shared__ float red[8];
  uint4 u[VPT];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      u[i] = __ldg(xr + vi);
      const __half2* h = reinterpret_cast<const __half2*>(&u[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        sq += f.x * f.x + f.y * f.y;
      }
    }
  }
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rstd = rsqrtf(tot / (float)cols + eps);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w) + vi);
      const __half2* h = reinterpret_cast<const __half2*>(&u[i]);
      const __half2* wh = reinterpret_cast<const __half2*>(&wv);
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        oh[j] = __hmul2(wh[j], __floats2half2_rn(f.x * rstd, f.y * rstd));
      }
      yr[vi] = o;
    }
  }
}

}  // namespace

extern "C" int stllm_layernorm_f16(const void* x, long long ldx, void* y, long long ldy, const float* gamma,
                                   const float* beta, int rows, int cols, float eps, void* stream) {
  STLLM_CHECK_ARG(x && y && gamma && beta && rows > 0 && cols > 0);
  STLLM_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0);
  STLLM_CHECK_ARG(((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0);
  if (cols % 8 == 0 && cols <= 32 * 8 * 6) {
    STLLM_CHECK_CUDA(stllm_launch(layernorm_f16_regs_kernel<6>, dim3((rows + 7) / 8), dim3(256), 0, (cudaStream_t)stream, (const __half*)x, ldx, (__half*)y, ldy, gamma, beta, rows, cols, eps));
  } else if (cols % 8 == 0 && cols <= 32 * 8 * 16) {
    STLLM_CHECK_CUDA(stllm_launch(layernorm_f16_regs_kernel<16>, dim3((rows + 7) / 8), dim3(256), 0, (cudaStream_t)stream, (const __half*)x, ldx, (__half*)y, ldy, gamma, beta, rows, cols, eps));
  } else {
    const int grid = (rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
    STLLM_CHECK_CUDA(stllm_launch(layernorm_f16_kernel, dim3(grid), dim3(ROWS_PER_CTA * 32), 0, (cudaStream_t)stream, (const __half*)x, ldx, (__half*)y, ldy, gamma, beta, rows, cols, eps));
  }
    return STLLM_OK;
}

extern "C" int stllm_rmsnorm_f16(const void* x, long long ldx, void* y, long long ldy, const void* weight, int rows,
                                 int cols, float eps, void* stream) {
  STLLM_CHECK_ARG(x && y && weight && rows > 0 && cols > 0);
  STLLM_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                  ((uintptr_t)weight & 15) == 0);
  if (cols % 8 == 0 && cols >= 2048 && cols <= 256 * 8 * 2) {
    STLLM_CHECK_CUDA(stllm_launch(rmsnorm_f16_block_kernel<2>, dim3(rows), dim3(256), 0, (cudaStream_t)stream, (const __half*)x, ldx, (__half*)y, ldy, (const __half*)weight, rows, cols, eps));
  } else {
    const int grid = (rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
    STLLM_CHECK_CUDA(stllm_launch(rmsnorm_f16_kernel, dim3(grid), dim3(ROWS_PER_CTA * 32), 0, (cudaStream_t)stream, (const __half*)x, ldx, (__half*)y, ldy, (const __half*)weight, rows, cols, eps));
  }
    return STLLM_OK;
}
