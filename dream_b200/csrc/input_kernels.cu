// input_kernels.cu -- the device side of the dataset's per-sample tensor work:
//   * dreamb200_normalize_u8 : uint8 HWC frames (PIL / decoder layout) -> fp32 NCHW, exactly
//     torchvision ToTensor (x/255) followed by Normalize ((x-mean)/std), the transform the reference builds at
//     dream/datasets.py:60-75 and applies at :177-179;
//   * dreamb200_belief_targets : keypoints -> training belief maps, the (4*sigma+1)^2 Gaussian stamp of
//     dream/image_proc.py:866-910 (centre truncated to an integer pixel, stamp dropped unless its window is
//     strictly inside the frame).
// Both are pure HBM writers (3 B -> 12 B per pixel; 4 B per target element).
#include "common.cuh"
#include "dreamb200.h"

namespace db200 {
int device_sm_count();

struct NormParams {
  const uint8_t* x;
  float* y;
  long long plane;       // H*W
  long long groups;      // ceil(plane/4) per image
  int B;
  float mean[3], stdv[3];
};

__device__ __forceinline__ float norm_one(uint32_t u, float mean, float stdv) {
  // same fp32 operations in the same order as ToTensor().div(255) and Normalize's sub_().div_()
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), mean), stdv);
}

__global__ void __launch_bounds__(256) normalize_u8_kernel(const __grid_constant__ NormParams p) {
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i >> 8][i & 255] = norm_one(i & 255, p.mean[i >> 8], p.stdv[i >> 8]);
  __syncthreads();
  const long long total = p.groups * p.B;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long b = g / p.groups, q = g - b * p.groups;
    const long long px = q * 4;
    const uint8_t* src = p.x + (b * p.plane + px) * 3;
    float* dst = p.y + b * 3 * p.plane + px;
    if (px + 4 <= p.plane && (p.plane & 3) == 0) {
      // 12 bytes = 4 pixels; the address is 4-byte aligned because plane*3 and px*3 are multiples of 4
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
      const uint32_t w0 = __ldg(s32), w1 = __ldg(s32 + 1), w2 = __ldg(s32 + 2);
      uint8_t v[12];
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i] = (w0 >> (8 * i)) & 255u; v[4 + i] = (w1 >> (8 * i)) & 255u; v[8 + i] = (w2 >> (8 * i)) & 255u; }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float4 o;
        o.x = lut[c][v[c]]; o.y = lut[c][v[3 + c]]; o.z = lut[c][v[6 + c]]; o.w = lut[c][v[9 + c]];
        *reinterpret_cast<float4*>(dst + c * p.plane) = o;
      }
    } else {
      for (int i = 0; i < 4 && px + i < p.plane; ++i)
        for (int c = 0; c < 3; ++c) dst[c * p.plane + i] = lut[c][src[i * 3 + c]];
    }
  }
}

struct TargetParams {
  const float* pts;   // [n,2] (x, y) in the target frame
  float* out;         // [n,h,w]
  int n, h, w, wr;    // wr = int(2*sigma)
  float table[64];    // exp(-d2/(2 sigma^2)) for d2 = 0 .. 2*wr*wr, computed by the caller in fp64 then rounded
};

__global__ void __launch_bounds__(256) belief_targets_kernel(const __grid_constant__ TargetParams p) {
  const int map = blockIdx.y;
  const float fx = __ldg(p.pts + 2 * map), fy = __ldg(p.pts + 2 * map + 1);
  bool valid = fabsf(fx) < 1.0e9f && fabsf(fy) < 1.0e9f;     // also false for NaN
  const int u = valid ? (int)truncf(fx) : 0, v = valid ? (int)truncf(fy) : 0;   // python int(): toward zero
  valid = valid && u - p.wr >= 0 && u + p.wr + 1 < p.w && v - p.wr >= 0 && v + p.wr + 1 < p.h;
  float* dst = p.out + (size_t)map * p.h * p.w;
  const int plane = p.h * p.w;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) * 4; e < plane; e += gridDim.x * blockDim.x * 4) {
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = e + i;
      const int j = idx / p.w, ii = idx - j * p.w;
      const int dx = ii - u, dy = j - v;
      const bool in = valid && dx >= -p.wr && dx <= p.wr && dy >= -p.wr && dy <= p.wr;
      o[i] = in ? p.table[dx * dx + dy * dy] : 0.0f;
    }
    if (e + 4 <= plane && (plane & 3) == 0) {
      *reinterpret_cast<float4*>(dst + e) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      for (int i = 0; i < 4 && e + i < plane; ++i) dst[e + i] = o[i];
    }
  }
}

}  // namespace db200

using namespace db200;

extern "C" int dreamb200_normalize_u8(const void* x_u8_nhwc, float* y_nchw, int B, int H, int W, const float* mean3,
                                      const float* std3, void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  DB_REQUIRE(x_u8_nhwc && y_nchw && mean3 && std3, "normalize_u8: null pointer");
  DB_REQUIRE(B >= 0 && H > 0 && W > 0, "normalize_u8: bad shape");
  if (B == 0) return 0;
  NormParams p;
  p.x = (const uint8_t*)x_u8_nhwc; p.y = y_nchw; p.plane = (long long)H * W; p.groups = (p.plane + 3) / 4; p.B = B;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.stdv[c] = std3[c]; }
  const long long total = p.groups * B;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  normalize_u8_kernel<<<(int)blocks, 256, 0, stream>>>(p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_belief_targets(const float* pts, int n_maps, int h, int w, int window_radius,
                                        const float* table, float* out, void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  DB_REQUIRE(pts && table && out, "belief_targets: null pointer");
  DB_REQUIRE(n_maps >= 0 && n_maps <= 65535 * 64 && h > 0 && w > 0, "belief_targets: bad shape");
  DB_REQUIRE(window_radius >= 0 && 2 * window_radius * window_radius + 1 <= 64, "belief_targets: window radius must be <= 5");
  if (n_maps == 0) return 0;
  TargetParams p;
  p.pts = pts; p.out = out; p.h = h; p.w = w; p.wr = window_radius;
  for (int i = 0; i < 64; ++i) p.table[i] = i <= 2 * window_radius * window_radius ? table[i] : 0.0f;
  const int plane = h * w;
  int bx = (plane / 4 + 255) / 256;
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  for (int done = 0; done < n_maps; done += 65535) {       // gridDim.y limit
    const int chunk = n_maps - done < 65535 ? n_maps - done : 65535;
    p.n = chunk; p.pts = pts + 2 * (size_t)done; p.out = out + (size_t)done * plane;
    belief_targets_kernel<<<dim3(bx, chunk), 256, 0, stream>>>(p);
    DB_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}
