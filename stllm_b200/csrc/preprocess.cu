// Frame preprocessing of the inference path on the GPU: decoded RGB uint8 frames [n, H, W, 3]
// -> fp16 [n, 3, S, S] encoder input, bit-exact with the reference's CPU pipeline
//     GroupScale(224, BICUBIC) -> GroupCenterCrop(224) -> Stack -> ToTorchFormatTensor -> GroupNormalize
// (conversation.py:192-200,274-278; video_transforms.py:66-71,94-123,367-406), whose resize is Pillow's 8-bit two-pass
// resampler (Resample.c): per output pixel a window of 22-bit fixed-point weights, int32 accumulation started at
// 1 << 21 (round half up), arithmetic shift, clamp to uint8 -- after the horizontal AND after the vertical pass.
// The weight tables are computed by the host in double precision exactly as Pillow does (stllm_b200/video_preprocess.py);
// here is only the integer arithmetic.  HBM-bound byte work: one thread per output pixel, channel-interleaved reads.
#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

DEVI int clip8(int v) {
  v >>= PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// pass 1 (horizontal): tmp[n][r][xc][c], r = input row y_first + r, xc = crop column -> resized column crop_left + xc
__global__ void resample_h_kernel(const uint8_t* __restrict__ in, int H, int W, const int* __restrict__ bounds,
                                  const int* __restrict__ kk, int ksize, int crop_left, int S, int y_first, int rows,
                                  uint8_t* __restrict__ tmp, int n_img) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)n_img * rows * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xc = (int)(i % S);
    const int r = (int)((i / S) % rows);
    const long long n = i / ((long long)S * rows);
    const int xx = crop_left + xc;
    const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    const int* k = kk + (long long)xx * ksize;
    const uint8_t* src = in + ((n * H + (y_first + r)) * (long long)W + xmin) * 3;
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < cnt; ++x) {
      const int w = __ldg(k + x);
      a0 += (int)src[3 * x] * w;
      a1 += (int)src[3 * x + 1] * w;
      a2 += (int)src[3 * x + 2] * w;
    }
    uint8_t* dst = tmp + i * 3;
    dst[0] = (uint8_t)clip8(a0);
    dst[1] = (uint8_t)clip8(a1);
    dst[2] = (uint8_t)clip8(a2);
  }
}

// pass 2 (vertical) + centre crop + x/255 + (x - mean)/std in fp32 (IEEE, the reference's operation order) + fp16
__global__ void resample_v_norm_kernel(const uint8_t* __restrict__ tmp, int rows, int y_first,
                                       const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                       int crop_top, int S, float m0, float m1, float m2, float s0, float s1, float s2,
                                       __half* __restrict__ out, int n_img) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)n_img * S * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xc = (int)(i % S);
    const int yc = (int)((i / S) % S);
    const long long n = i / ((long long)S * S);
    const int yy = crop_top + yc;
    const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    const int* k = kk + (long long)yy * ksize;
    const uint8_t* src = tmp + ((n * rows + (ymin - y_first)) * (long long)S + xc) * 3;
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int y = 0; y < cnt; ++y) {
      const int w = __ldg(k + y);
      const uint8_t* p = src + (long long)y * S * 3;
      a0 += (int)p[0] * w;
      a1 += (int)p[1] * w;
      a2 += (int)p[2] * w;
    }
    const float v0 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(a0), 255.0f), m0), s0);
    const float v1 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(a1), 255.0f), m1), s1);
    const float v2 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(a2), 255.0f), m2), s2);
    const long long plane = (long long)S * S;
    __half* o = out + n * 3 * plane + (long long)yc * S + xc;
    o[0] = __float2half_rn(v0);
    o[plane] = __float2half_rn(v1);
    o[2 * plane] = __float2half_rn(v2);
  }
}

}  // namespace

// frames [n, H, W, 3] uint8 -> out [n, 3, S, S] fp16.  xbounds/xcoef: [new_w][2] / [new_w][kx] and ybounds/ycoef:
// [new_h][2] / [new_h][ky] int32 device tables of Pillow's resampling windows for W -> new_w and H -> new_h (identity
// windows when a side is not resized); crop_left/crop_top: offsets of the S x S centre crop in the resized image.
// tmp: caller-owned scratch of at least n * rows * S * 3 bytes where rows = y_last - y_first is the span of input
// rows the cropped output rows read.  mean/std: host pointers to 3 floats.
extern "C" int stllm_preprocess_frames_u8(const void* frames, int n_img, int H, int W, int S, const int* xbounds,
                                          const int* xcoef, int kx, int crop_left, const int* ybounds,
                                          const int* ycoef, int ky, int crop_top, int y_first, int y_last,
                                          const float* mean3, const float* std3, void* tmp, long long tmp_bytes,
                                          void* out, void* stream) {
  STLLM_CHECK_ARG(frames && xbounds && xcoef && ybounds && ycoef && mean3 && std3 && tmp && out);
  STLLM_CHECK_ARG(n_img > 0 && H > 0 && W > 0 && S > 0 && kx > 0 && ky > 0 && crop_left >= 0 && crop_top >= 0);
  STLLM_CHECK_ARG(y_first >= 0 && y_last > y_first && y_last <= H);
  const int rows = y_last - y_first;
  STLLM_CHECK_ARG(tmp_bytes >= (long long)n_img * rows * S * 3);
  const long long t1 = (long long)n_img * rows * S, t2 = (long long)n_img * S * S;
  const int g1 = (int)((t1 + 255) / 256 < 148 * 32 ? (t1 + 255) / 256 : 148 * 32);
  const int g2 = (int)((t2 + 255) / 256 < 148 * 32 ? (t2 + 255) / 256 : 148 * 32);
  cudaStream_t s = (cudaStream_t)stream;
  STLLM_CHECK_CUDA(stllm_launch(resample_h_kernel, dim3(g1), dim3(256), 0, s, (const uint8_t*)frames, H, W, xbounds,
                                xcoef, kx, crop_left, S, y_first, rows, (uint8_t*)tmp, n_img));
  STLLM_CHECK_CUDA(stllm_launch(resample_v_norm_kernel, dim3(g2), dim3(256), 0, s, (const uint8_t*)tmp, rows, y_first,
                                ybounds, ycoef, ky, crop_top, S, mean3[0], mean3[1], mean3[2], std3[0], std3[1],
                                std3[2], (__half*)out, n_img));
  return STLLM_OK;
}
