// aux_kernels.cu -- HBM-bound helpers around the tensor-core convolutions: first-layer patch
// gather, max pooling, nearest upsampling and the NCHW<->NHWC boundary conversions.
// All are pure streaming kernels: 16-byte vector accesses along the channel axis, grid-stride
// loops sized to a multiple of the SM count.
#include "common.cuh"
#include "dreamb200.h"

#include <stdarg.h>

#include <atomic>

namespace db200 {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n); }

int device_sm_count();   // conv_tc.cu

static int grid_for(long long work_items, int threads) {
  const int sms = device_sm_count();
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// First-layer patch gather: x fp32 NCHW [B,3,H,W] -> fp16 [B,Ho,Wo,Kpad], k = (r*S+s)*3 + c.
// A block step handles 32 consecutive output pixels: warp w gathers k-group w (8 consecutive k) for the 32
// pixels (lane = pixel, so the fp32 reads are coalesced along x), the 16-byte results cross through shared
// memory, and the block then writes the 32 x Kpad/8 chunks as one contiguous, fully coalesced run.
// ---------------------------------------------------------------------------------------------
template <int R, int S>
__global__ void __launch_bounds__(256)
im2col_first_kernel(const float* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int stride, int pad,
                    int Ho, int Wo, int Kpad) {
  constexpr int K = R * S * 3;
  const int kgroups = Kpad / 8;                       // <= 24 (Kpad <= 192)
  __shared__ uint4 tile[32][25];
  const long long npix = (long long)B * Ho * Wo;
  const long long nsteps = (npix + 31) / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t plane = (size_t)H * W;
  for (long long step = blockIdx.x; step < nsteps; step += gridDim.x) {
    const long long pix = step * 32 + lane;
    const bool pvalid = pix < npix;
    int ox = 0, oy = 0, b = 0;
    if (pvalid) {
      ox = (int)(pix % Wo);
      const long long r = pix / Wo;
      oy = (int)(r % Ho);
      b = (int)(r / Ho);
    }
    const float* xb = x + (size_t)b * 3 * plane;
    for (int kg = warp; kg < kgroups; kg += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = kg * 8 + j;
        float f = 0.0f;
        if (pvalid && k < K) {
          const int c = k % 3;
          const int rs = k / 3;
          const int r = rs / S, s_ = rs - r * S;
          const int iy = oy * stride - pad + r, ix = ox * stride - pad + s_;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) f = __ldg(xb + c * plane + (size_t)iy * W + ix);
        }
        v[j] = f;
      }
      uint4 o;
      __half2 h;
      h = __floats2half2_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2half2_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2half2_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2half2_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
      tile[lane][kg] = o;
    }
    __syncthreads();
    const int chunks = 32 * kgroups;
    for (int i = threadIdx.x; i < chunks; i += 256) {
      const int pl = i / kgroups, kg = i - pl * kgroups;
      if (step * 32 + pl < npix) out[(size_t)step * 32 * kgroups + i] = tile[pl][kg];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// k x k max pool, NHWC fp16, 8 channels per thread.  Padding cells never win (-inf), like ATen.
// ---------------------------------------------------------------------------------------------
__global__ void maxpool_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W,
                                    int C8, int k, int s, int p, int Ho, int Wo) {
  const long long total = (long long)B * Ho * Wo * C8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8);
    long long pix = idx / C8;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    const __half2 ninf = __float2half2_rn(-65504.0f);
    __half2 m[4] = {ninf, ninf, ninf, ninf};
    for (int r = 0; r < k; ++r) {
      const int iy = oy * s - p + r;
      if (iy < 0 || iy >= H) continue;
      for (int q = 0; q < k; ++q) {
        const int ix = ox * s - p + q;
        if (ix < 0 || ix >= W) continue;
        const uint4 v = __ldg(x + ((size_t)((size_t)b * H + iy) * W + ix) * C8 + c);
        const __half2* vh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], vh[j]);
      }
    }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&m[0]);
    o.y = *reinterpret_cast<uint32_t*>(&m[1]);
    o.z = *reinterpret_cast<uint32_t*>(&m[2]);
    o.w = *reinterpret_cast<uint32_t*>(&m[3]);
    y[idx] = o;
  }
}

// nearest x2 upsample, NHWC fp16
__global__ void upsample2_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W,
                                      int C8) {
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)B * Ho * Wo * C8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8);
    long long pix = idx / C8;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    y[idx] = __ldg(x + ((size_t)((size_t)b * H + (oy >> 1)) * W + (ox >> 1)) * C8 + c);
  }
}

// y += x, fp16, 8 elements per thread (hourglass skip connections, models.py:775-799)
__global__ void add_f16_kernel(uint4* __restrict__ y, const uint4* __restrict__ x, long long n8) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n8;
       idx += (long long)gridDim.x * blockDim.x) {
    uint4 a = y[idx];
    const uint4 b = __ldg(x + idx);
    __half2* ah = reinterpret_cast<__half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int j = 0; j < 4; ++j) ah[j] = __hadd2(ah[j], bh[j]);
    y[idx] = a;
  }
}

// fp16 NHWC [B,H,W,Cpad] -> fp32 NCHW [B,C,H,W]; a 32-pixel x 32-channel smem transpose per block step
__global__ void nhwc_f16_to_nchw_f32_kernel(const __half* __restrict__ x, float* __restrict__ y, int B,
                                            long long HW, int Cpad, int C) {
  __shared__ float tile[32][33];
  const long long ptiles = (HW + 31) / 32;
  const int ctiles = (C + 31) / 32;
  const long long total = (long long)B * ptiles * ctiles;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ct = (int)(t % ctiles);
    long long r = t / ctiles;
    const long long pt = r % ptiles;
    const int b = (int)(r / ptiles);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
      const long long pix = pt * 32 + i;
      const int c = ct * 32 + tx;
      float v = 0.0f;
      if (pix < HW && c < C) v = __half2float(x[((size_t)b * HW + pix) * Cpad + c]);
      tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = ct * 32 + i;
      const long long pix = pt * 32 + tx;
      if (pix < HW && c < C) y[((size_t)b * C + c) * HW + pix] = tile[tx][i];
    }
    __syncthreads();
  }
}

__global__ void nchw_f32_to_nhwc_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, int B,
                                            long long HW, int C, int Cpad) {
  __shared__ float tile[32][33];
  const long long ptiles = (HW + 31) / 32;
  const int ctiles = (Cpad + 31) / 32;
  const long long total = (long long)B * ptiles * ctiles;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ct = (int)(t % ctiles);
    long long r = t / ctiles;
    const long long pt = r % ptiles;
    const int b = (int)(r / ptiles);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
      const int c = ct * 32 + i;
      const long long pix = pt * 32 + tx;
      float v = 0.0f;
      if (pix < HW && c < C) v = x[((size_t)b * C + c) * HW + pix];
      tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const long long pix = pt * 32 + i;
      const int c = ct * 32 + tx;
      if (pix < HW && c < Cpad) y[((size_t)b * HW + pix) * Cpad + c] = __float2half_rn(tile[tx][i]);
    }
    __syncthreads();
  }
}

}  // namespace db200

using namespace db200;

extern "C" const char* dreamb200_last_error(void) { return g_err; }
extern "C" int dreamb200_version(void) { return 100; }
extern "C" int64_t dreamb200_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int dreamb200_im2col_first(const float* x, void* out, int B, int H, int W, int R, int S,
                                      int stride, int pad, int Ho, int Wo, int Kpad, void* stream) {
  DB_REQUIRE(x && out, "im2col_first: null pointer");
  DB_REQUIRE(Kpad % 8 == 0 && Kpad >= R * S * 3, "im2col_first: Kpad=%d too small / unaligned", Kpad);
  DB_REQUIRE(B > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "im2col_first: empty tensor");
  DB_REQUIRE(Kpad <= 192, "im2col_first: Kpad=%d > 192", Kpad);
  DB_REQUIRE((R == 3 && S == 3) || (R == 7 && S == 7), "im2col_first: only 3x3 and 7x7 first layers exist (got %dx%d)",
             R, S);
  const long long steps = ((long long)B * Ho * Wo + 31) / 32;
  const int grid = grid_for(steps * 256, 256);
  if (R == 3)
    im2col_first_kernel<3, 3><<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<uint4*>(out), B, H, W,
                                                                     stride, pad, Ho, Wo, Kpad);
  else
    im2col_first_kernel<7, 7><<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<uint4*>(out), B, H, W,
                                                                     stride, pad, Ho, Wo, Kpad);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_maxpool_nhwc(const void* x, void* y, int B, int H, int W, int C, int k, int s, int p,
                                      int Ho, int Wo, void* stream) {
  DB_REQUIRE(x && y, "maxpool: null pointer");
  DB_REQUIRE(C % 8 == 0, "maxpool: C=%d must be a multiple of 8", C);
  DB_REQUIRE(B > 0 && Ho > 0 && Wo > 0, "maxpool: empty tensor");
  const long long total = (long long)B * Ho * Wo * (C / 8);
  maxpool_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), B, H, W, C / 8, k, s, p, Ho, Wo);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_upsample2_nhwc(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  DB_REQUIRE(x && y, "upsample2: null pointer");
  DB_REQUIRE(C % 8 == 0, "upsample2: C=%d must be a multiple of 8", C);
  const long long total = (long long)B * 4 * H * W * (C / 8);
  upsample2_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), B, H, W, C / 8);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_nhwc_f16_to_nchw_f32(const void* x, float* y, int B, int H, int W, int Cpad, int C,
                                              void* stream) {
  DB_REQUIRE(x && y, "nhwc->nchw: null pointer");
  const long long HW = (long long)H * W;
  const long long tiles = (long long)B * ((HW + 31) / 32) * ((C + 31) / 32);
  nhwc_f16_to_nchw_f32_kernel<<<grid_for(tiles * 256, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __half*>(x), y, B, HW, Cpad, C);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_nchw_f32_to_nhwc_f16(const float* x, void* y, int B, int H, int W, int C, int Cpad,
                                              void* stream) {
  DB_REQUIRE(x && y, "nchw->nhwc: null pointer");
  const long long HW = (long long)H * W;
  const long long tiles = (long long)B * ((HW + 31) / 32) * ((Cpad + 31) / 32);
  nchw_f32_to_nhwc_f16_kernel<<<grid_for(tiles * 256, 256), 256, 0, (cudaStream_t)stream>>>(
      x, reinterpret_cast<__half*>(y), B, HW, C, Cpad);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_add_f16(void* y, const void* x, long long n, void* stream) {
  DB_REQUIRE(y && x, "add_f16: null pointer");
  DB_REQUIRE(n > 0 && n % 8 == 0, "add_f16: n=%lld must be a positive multiple of 8", n);
  add_f16_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<uint4*>(y), reinterpret_cast<const uint4*>(x), n / 8);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
