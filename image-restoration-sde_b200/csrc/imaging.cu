// Image <-> tensor conversion and the full-reference metrics the reference's test loops compute on the CPU
// (codes/utils/img_utils.py:136-163 tensor2img, :171-180 img2tensor, :182-190 calculate_psnr, :193-234 ssim /
// calculate_ssim).  Integer work (uint8 quantisation, BGR/HWC index maps, squared-error sums) is bit exact;
// SSIM is fp64 like the reference.  All kernels are HBM-bound streaming passes.
#include <math.h>

#include "common.cuh"

namespace irsde {

// ---------------------------------------------------------------------------------------------
// tensor2img: fp32 [B][C][H][W] (RGB) -> uint8 [B][H][W][C] (BGR for C == 3)
//   clamp to [lo,hi], (t - lo) / (hi - lo), * 255, round half to even (numpy .round()), cast
//   (img_utils.py:142-143,159-162).  Every step is a single fp32 operation, like the torch / numpy chain.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned char quant_u8(float v, float lo, float hi, float range) {  // range = fp32(hi - lo)
  v = fminf(fmaxf(v, lo), hi);
  v = __fdiv_rn(__fsub_rn(v, lo), range);
  return (unsigned char)__float2int_rn(__fmul_rn(v, 255.0f));
}

__global__ void __launch_bounds__(256) tensor2img_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, int C,
                                                         long long HW, float lo, float hi, float range) {
  const int b = blockIdx.y;
  const float* src = in + (long long)b * C * HW;
  unsigned char* dst = out + (long long)b * C * HW;
  // 4 pixels per thread: 3 x float4 loads, 12 bytes out
  const long long p0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (p0 >= HW) return;
  if (C == 3 && p0 + 3 < HW && (HW & 3) == 0) {
    const float4 r = *reinterpret_cast<const float4*>(src + p0);
    const float4 g = *reinterpret_cast<const float4*>(src + HW + p0);
    const float4 bl = *reinterpret_cast<const float4*>(src + 2 * HW + p0);
    const float rr[4] = {r.x, r.y, r.z, r.w}, gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {bl.x, bl.y, bl.z, bl.w};
    unsigned char o[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[3 * i + 0] = quant_u8(bb[i], lo, hi, range);  // BGR
      o[3 * i + 1] = quant_u8(gg[i], lo, hi, range);
      o[3 * i + 2] = quant_u8(rr[i], lo, hi, range);
    }
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + p0 * 3);  // p0 % 4 == 0 and b*3*HW % 4 == 0 => 4-byte aligned
#pragma unroll
    for (int i = 0; i < 3; ++i)
      d[i] = (uint32_t)o[4 * i] | ((uint32_t)o[4 * i + 1] << 8) | ((uint32_t)o[4 * i + 2] << 16) | ((uint32_t)o[4 * i + 3] << 24);
    return;
  }
  for (long long p = p0; p < p0 + 4 && p < HW; ++p)
    for (int c = 0; c < C; ++c) {
      const int cs = (C == 3) ? 2 - c : c;
      dst[p * C + c] = quant_u8(src[(long long)cs * HW + p], lo, hi, range);
    }
}

// img2tensor / read_img: uint8 [B][H][W][C] (BGR) -> fp32 [B][C][H][W] (RGB), v / 255 (img_utils.py:176-179,
// codes/data/util.py:72).
__global__ void __launch_bounds__(256) img2tensor_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, int C,
                                                         long long HW) {
  const int b = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const unsigned char* src = in + ((long long)b * HW + p) * C;
  float* dst = out + (long long)b * C * HW + p;
  for (int c = 0; c < C; ++c) {
    const int cs = (C == 3) ? 2 - c : c;
    dst[(long long)c * HW] = __fdiv_rn((float)src[cs], 255.0f);
  }
}

// ---------------------------------------------------------------------------------------------
// PSNR numerator: sum over the cropped region of (a - b)^2 on uint8 images, exact in uint64
// (calculate_psnr: mean of squared differences in float64 - every partial sum is an integer < 2^53, so the
// float64 mean equals sum / n exactly, whatever the summation order).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sqerr_u8_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b,
                                                       int H, int W, int C, int crop, unsigned long long* __restrict__ out) {
  const int img = blockIdx.y;
  const int h = H - 2 * crop, w = W - 2 * crop;
  const long long rowlen = (long long)w * C, total = (long long)h * rowlen;
  const unsigned char* pa = a + (long long)img * H * W * C;
  const unsigned char* pb = b + (long long)img * H * W * C;
  unsigned long long acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long y = i / rowlen, x = i - y * rowlen;
    const long long off = ((y + crop) * W + crop) * C + x;
    const int d = (int)pa[off] - (int)pb[off];
    acc += (unsigned long long)(d * d);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ unsigned long long red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out + img, t);  // integer: order independent
  }
}

// ---------------------------------------------------------------------------------------------
// SSIM (img_utils.py:193-214): 11x11 Gaussian window (sigma 1.5, cv2.getGaussianKernel), "valid" region
// [5:-5, 5:-5], C1 = (0.01*255)^2, C2 = (0.03*255)^2, fp64; for HxWx3 inputs the map covers all three
// channels and calculate_ssim's three identical passes average to the same number (:225-229).
// One block = 16x16 outputs of one channel of one image; block partial sums go to `partial` and are
// reduced in a fixed order by ssim_reduce_kernel (deterministic).
// ---------------------------------------------------------------------------------------------
struct SsimWin { double k[11]; };

__global__ void __launch_bounds__(256) ssim_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b,
                                                   int H, int W, int C, int crop, SsimWin win, double* __restrict__ partial) {
  __shared__ float sa[26][26], sb[26][26];
  __shared__ double red[8];
  const int h = H - 2 * crop, w = W - 2 * crop;  // metric runs on the cropped image
  const int vh = h - 10, vw = w - 10;
  const int tiles_x = (vw + 15) / 16;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int c = blockIdx.y, img = blockIdx.z;
  const unsigned char* pa = a + (long long)img * H * W * C;
  const unsigned char* pb = b + (long long)img * H * W * C;
  for (int i = threadIdx.x; i < 26 * 26; i += 256) {
    const int yy = i / 26, xx = i - yy * 26;
    const int y = ty * 16 + yy, x = tx * 16 + xx;  // cropped-image coordinates of the window origin + (yy,xx)
    float va = 0.f, vb = 0.f;
    if (y < h && x < w) {
      const long long off = ((long long)(y + crop) * W + (x + crop)) * C + c;
      va = (float)pa[off]; vb = (float)pb[off];
    }
    sa[yy][xx] = va; sb[yy][xx] = vb;
  }
  __syncthreads();
  const int oy = threadIdx.x >> 4, ox = threadIdx.x & 15;
  double val = 0.0;
  if (ty * 16 + oy < vh && tx * 16 + ox < vw) {
    double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int r = 0; r < 11; ++r) {
      for (int q = 0; q < 11; ++q) {
        const double wv = win.k[r] * win.k[q];
        const double x1 = (double)sa[oy + r][ox + q], x2 = (double)sb[oy + r][ox + q];
        m1 += wv * x1; m2 += wv * x2;
        s11 += wv * (x1 * x1); s22 += wv * (x2 * x2); s12 += wv * (x1 * x2);
      }
    }
    const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
    const double m1s = m1 * m1, m2s = m2 * m2, m12 = m1 * m2;
    const double v1 = s11 - m1s, v2 = s22 - m2s, v12 = s12 - m12;
    val = ((2 * m12 + C1) * (2 * v12 + C2)) / ((m1s + m2s + C1) * (v1 + v2 + C2));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) val += __shfl_xor_sync(0xffffffffu, val, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[((long long)img * gridDim.y + c) * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void ssim_reduce_kernel(const double* __restrict__ partial, int n, double denom, double* __restrict__ out) {
  __shared__ double red[256];
  const int img = blockIdx.x;
  double t = 0;
  for (int i = threadIdx.x; i < n; i += 256) t += partial[(long long)img * n + i];
  red[threadIdx.x] = t;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[img] = red[0] / denom;
}

// ---- launchers ------------------------------------------------------------------------------
void launch_tensor2img(const float* in, unsigned char* out, int B, int C, int H, int W, double lo, double hi, cudaStream_t st) {
  const long long HW = (long long)H * W;
  // python scalars: `hi - lo` is evaluated in double, then each scalar enters the fp32 tensor op rounded to fp32
  tensor2img_kernel<<<dim3((unsigned)((HW + 1023) / 1024), B), 256, 0, st>>>(in, out, C, HW, (float)lo, (float)hi, (float)(hi - lo));
}
void launch_img2tensor(const unsigned char* in, float* out, int B, int C, int H, int W, cudaStream_t st) {
  const long long HW = (long long)H * W;
  img2tensor_kernel<<<dim3((unsigned)((HW + 255) / 256), B), 256, 0, st>>>(in, out, C, HW);
}
void launch_sqerr_u8(const unsigned char* a, const unsigned char* b, int B, int H, int W, int C, int crop, unsigned long long* out,
                     cudaStream_t st) {
  cudaMemsetAsync(out, 0, sizeof(unsigned long long) * B, st);
  const long long total = (long long)(H - 2 * crop) * (W - 2 * crop) * C;
  int blocks = (int)((total + 256 * 8 - 1) / (256 * 8));
  blocks = blocks < 1 ? 1 : (blocks > 592 ? 592 : blocks);
  sqerr_u8_kernel<<<dim3(blocks, B), 256, 0, st>>>(a, b, H, W, C, crop, out);
}
long long ssim_partial_count(int H, int W, int C, int crop) {
  const int vh = H - 2 * crop - 10, vw = W - 2 * crop - 10;
  return (long long)((vh + 15) / 16) * ((vw + 15) / 16) * C;
}
void launch_ssim_u8(const unsigned char* a, const unsigned char* b, int B, int H, int W, int C, int crop, double* partial, double* out,
                    cudaStream_t st) {
  SsimWin win;
  double sum = 0;
  for (int i = 0; i < 11; ++i) { win.k[i] = exp(-((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); sum += win.k[i]; }
  for (int i = 0; i < 11; ++i) win.k[i] /= sum;
  const int vh = H - 2 * crop - 10, vw = W - 2 * crop - 10;
  const int tiles = ((vh + 15) / 16) * ((vw + 15) / 16);
  ssim_kernel<<<dim3(tiles, C, B), 256, 0, st>>>(a, b, H, W, C, crop, win, partial);
  ssim_reduce_kernel<<<B, 256, 0, st>>>(partial, tiles * C, (double)vh * vw * C, out);
}

}  // namespace irsde
