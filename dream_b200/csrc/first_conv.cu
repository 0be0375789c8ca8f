// first_conv.cu -- the network's first convolution (3 -> 64 channels, 3x3, stride 1, padding 1 + bias + ReLU;
// reference: dream/models.py:591-599 `layer_0_1_down.0`) fused with the fp32 NCHW -> fp16 NHWC input pack.
//
// K = 27 is far too small for an im2col-free TMA operand (3 channels = 6 bytes per pixel), and a
// materialised patch tensor costs 2 x 128 B per pixel of HBM traffic.  Here producer warps gather each output
// pixel's 27 inputs straight from the fp32 NCHW image (L1/L2 absorb the 9x reuse), round them to fp16 and write
// the A tile (128 pixels x 32 k, zero padded) into shared memory in the canonical SWIZZLE_128B K-major layout
// by hand; one thread issues two K=16 tcgen05.mma per tile against the 64x32 weight tile that stays resident
// in shared memory; the epilogue is the conv_tc one (TMEM -> bias + ReLU -> fp16 -> swizzled smem -> TMA
// store).  HBM-bound: 12 B read + 128 B written per pixel.
//
// Warps: 0-3 producer group 0, 4-7 producer group 1 (tiles alternate between the groups / A stages),
//        8 = MMA issuer + TMEM owner + weight load, 9-12 epilogue.  Persistent, grid = #SMs.
#include "common.cuh"
#include "dreamb200.h"

namespace db200 {

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                        const char* what);
int device_sm_count();

struct FirstConvParams {
  const float* x;     // [B,3,H,W] fp32 (already normalised) ...
  const uint8_t* xu8; // ... or [B,H,W,3] uint8 frames, normalised on the fly with (u/255 - mean)/std
  float mean[3], stdv[3];
  const float* bias;  // [64] fp32
  int B, H, W;
  int tiles_x, tiles_y, total_tiles;
};

constexpr int kFcThreads = 13 * 32;
constexpr int kFcTw = 16, kFcTh = 8;

template <bool U8>
__global__ void __launch_bounds__(kFcThreads, 1)
first_conv_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmC,
                  const __grid_constant__ FirstConvParams p) {
  constexpr int kABytes = 128 * 128;
  constexpr uint32_t kIdesc = umma_idesc_f16_m128(64);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;                  // 2 x 16 KB
  const uint32_t smem_w = smem_a + 2 * kABytes;       // 8 KB (64 co x 128 B)
  const uint32_t smem_out = smem_w + 8192;            // 2 x 16 KB
  const uint32_t bar_base = smem_out + 2 * kABytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };          // 128 producer arrivals
  auto empty_bar = [&](int s) { return bar_base + 16u + 8u * s; };   // MMA commit
  auto tfull_bar = [&](int a) { return bar_base + 32u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 48u + 8u * a; };
  const uint32_t w_bar = bar_base + 64u;
  const uint32_t tmem_ptr_smem = bar_base + 72u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  float* lut = reinterpret_cast<float*>(smem_gen + (bar_base + 128u - smem_base));   // [3][256], U8 only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (U8) {
    // the same fp32 operations as torchvision's ToTensor + Normalize (dream/datasets.py:60-75), so the fp16
    // operand is bit-identical to packing the host-normalised fp32 tensor
    for (int i = threadIdx.x; i < 768; i += kFcThreads)
      lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), p.mean[i >> 8]), p.stdv[i >> 8]);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), 128);
      mbar_init(empty_bar(s), 1);
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<128>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp < 8) {
    // ===================== producers: gather + pack the A tile =====================
    const int grp = warp >> 2;                 // A stage owned by this group
    const int t = threadIdx.x & 127;           // row of the tile == pixel
    const int ly = t / kFcTw, lx = t % kFcTw;
    const size_t plane = (size_t)p.H * p.W;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      int r = tile;
      const int tx = r % p.tiles_x; r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      const int b = r / p.tiles_y;
      const int ox = tx * kFcTw + lx, oy = ty * kFcTh + ly;
      const float* xb = p.x + (size_t)b * 3 * plane;
      const uint8_t* ub = p.xu8 + (size_t)b * 3 * plane;
      float v[32];
#pragma unroll
      for (int k = 27; k < 32; ++k) v[k] = 0.0f;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const int iy = oy + rr - 1;
        const bool yok = iy >= 0 && iy < p.H;
#pragma unroll
        for (int ss = 0; ss < 3; ++ss) {
          const int ix = ox + ss - 1;
          const bool ok = yok && ix >= 0 && ix < p.W;
          const size_t off = (size_t)(ok ? iy : 0) * p.W + (ok ? ix : 0);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float f = U8 ? lut[c * 256 + __ldg(ub + off * 3 + c)] : __ldg(xb + c * plane + off);
            v[(rr * 3 + ss) * 3 + c] = ok ? f : 0.0f;
          }
        }
      }
      mbar_wait(empty_bar(grp), phase ^ 1u);
      const uint32_t row_addr = smem_a + grp * kABytes + (uint32_t)t * 128u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half2 h0 = __floats2half2_rn(v[8 * j + 0], v[8 * j + 1]);
        __half2 h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
        __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]);
        __half2 h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
        const uint32_t dst = row_addr + (((uint32_t)j ^ (uint32_t)(t & 7)) * 16u);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                     "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                     "r"(*reinterpret_cast<uint32_t*>(&h3))
                     : "memory");
      }
      fence_proxy_async_smem();
      mbar_arrive(full_bar(grp));
      phase ^= 1u;
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      mbar_expect_tx(w_bar, 8192u);
      tma_load_3d(smem_w, &tmW, w_bar, 0, 0, 0);
      mbar_wait(w_bar, 0);
      uint32_t ph[2] = {0u, 0u};
      int as = 0;
      uint32_t aphase = 0;
      int it = 0;
      const uint64_t bdesc = umma_desc_k_sw128(smem_w);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        mbar_wait(full_bar(s), ph[s]);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(smem_a + s * kABytes);
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 64);
        umma_f16(d_tmem, adesc, bdesc, kIdesc, 0u);
        umma_f16(d_tmem, adesc + 2u, bdesc + 2u, kIdesc, 1u);
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(as));
        ph[s] ^= 1u;
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 9..12) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - 9 * 32;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t ctr = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ctr) {
      int r = tile;
      const int tx = r % p.tiles_x; r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      const int b = r / p.tiles_y;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 64);
      const uint32_t obuf = smem_out + (ctr & 1u) * kABytes;
      if (epi_tid == 0) tma_store_wait_read<1>();
      named_bar_sync(1, 128);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + (uint32_t)(h * 32), v);
        tmem_wait_ld();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + h * 32 + i));
          f[i] = fmaxf(__uint_as_float(v[i]) + bv.x, 0.0f);
          f[i + 1] = fmaxf(__uint_as_float(v[i + 1]) + bv.y, 0.0f);
          f[i + 2] = fmaxf(__uint_as_float(v[i + 2]) + bv.z, 0.0f);
          f[i + 3] = fmaxf(__uint_as_float(v[i + 3]) + bv.w, 0.0f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __half2 h0 = __floats2half2_rn(f[8 * j + 0], f[8 * j + 1]);
          __half2 h1 = __floats2half2_rn(f[8 * j + 2], f[8 * j + 3]);
          __half2 h2 = __floats2half2_rn(f[8 * j + 4], f[8 * j + 5]);
          __half2 h3 = __floats2half2_rn(f[8 * j + 6], f[8 * j + 7]);
          const uint32_t chunk16 = (uint32_t)(h * 4 + j) ^ (uint32_t)(row & 7);
          const uint32_t dst = obuf + (uint32_t)row * 128u + chunk16 * 16u;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                       "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)),
                       "r"(*reinterpret_cast<uint32_t*>(&h2)), "r"(*reinterpret_cast<uint32_t*>(&h3))
                       : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (epi_tid == 0) {
        tma_store_4d(&tmC, obuf, 0, tx * kFcTw, ty * kFcTh, b);
        tma_store_commit();
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if (epi_tid == 0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace db200

using namespace db200;

static int first_conv_launch(const float* x, const uint8_t* xu8, const float* mean3, const float* std3,
                             const void* w, const float* bias, void* y, int B, int H, int W, cudaStream_t stream) {
  DB_REQUIRE((x || xu8) && w && bias && y, "first_conv: null pointer");
  DB_REQUIRE(B > 0 && H > 0 && W > 0, "first_conv: empty input");
  FirstConvParams p;
  p.x = x; p.xu8 = xu8; p.bias = bias; p.B = B; p.H = H; p.W = W;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3 ? mean3[c] : 0.0f; p.stdv[c] = std3 ? std3[c] : 1.0f; }
  p.tiles_x = (W + kFcTw - 1) / kFcTw;
  p.tiles_y = (H + kFcTh - 1) / kFcTh;
  p.total_tiles = p.tiles_x * p.tiles_y * B;
  CUtensorMap tmW, tmC;
  {
    uint64_t dims[3] = {64, 64, 1};
    uint64_t str[2] = {128, 64 * 128};
    uint32_t box[3] = {64, 64, 1};
    uint32_t es[3] = {1, 1, 1};
    if (make_tensor_map_f16(&tmW, w, 3, dims, str, box, es, "first-conv weights")) return -1;
  }
  {
    uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
    uint32_t box[4] = {64, kFcTw, kFcTh, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (make_tensor_map_f16(&tmC, y, 4, dims, str, box, es, "first-conv output")) return -1;
  }
  const int smem_bytes = 1024 + 2 * 16384 + 8192 + 2 * 16384 + 128 + 3072;
  static bool attr_set = false;
  if (!attr_set) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    DB_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  const int sms = device_sm_count();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (xu8) first_conv_kernel<true><<<grid, kFcThreads, smem_bytes, stream>>>(tmW, tmC, p);
  else first_conv_kernel<false><<<grid, kFcThreads, smem_bytes, stream>>>(tmW, tmC, p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// x fp32 NCHW [B,3,H,W]; w fp16 [1][64][64] (k=(r*3+s)*3+c, zero padded); bias fp32 [64]; y fp16 NHWC [B,H,W,64]
extern "C" int dreamb200_first_conv3x3(const float* x, const void* w, const float* bias, void* y, int B, int H,
                                       int W, void* stream_v) {
  DB_REQUIRE(x != nullptr, "first_conv: null input");
  return first_conv_launch(x, nullptr, nullptr, nullptr, w, bias, y, B, H, W, (cudaStream_t)stream_v);
}

// same layer fed with raw uint8 HWC frames [B,H,W,3]; mean3/std3 are host pointers (image_normalization)
extern "C" int dreamb200_first_conv3x3_u8(const void* x_u8, const float* mean3, const float* std3, const void* w,
                                          const float* bias, void* y, int B, int H, int W, void* stream_v) {
  DB_REQUIRE(x_u8 && mean3 && std3, "first_conv_u8: null pointer");
  return first_conv_launch(nullptr, (const uint8_t*)x_u8, mean3, std3, w, bias, y, B, H, W, (cudaStream_t)stream_v);
}
