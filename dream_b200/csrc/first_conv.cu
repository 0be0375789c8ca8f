// first_conv.cu -- the network's first convolution (3 -> 64 channels, 3x3, stride 1, padding 1 + bias + ReLU;
// reference: dream/models.py:591-599 `layer_0_1_down.0`) fused with the fp32 NCHW -> fp16 NHWC input pack, or --
// for raw uint8 HWC frames -- with the dataset's ToTensor + Normalize (dream/datasets.py:60-75,177-179).
//
// K = 27 is far too small for an im2col-free TMA operand (3 channels = 6 bytes per pixel), and a
// materialised patch tensor costs 2 x 128 B per pixel of HBM traffic.  Here the input patch of every 16x8 output
// tile (10 rows x 24 columns around it, zero filled outside the image by TMA) is staged in shared memory by a TMA
// warp running 4 tiles ahead; producer warps read each output pixel's 27 inputs from that patch, round them to
// fp16 and write the A tile (128 pixels x 32 k, zero padded) in the canonical SWIZZLE_128B K-major layout by
// hand; one thread issues two K=16 tcgen05.mma per tile against the 64x32 weight tile that stays resident in
// shared memory; the epilogue is the conv_tc one (TMEM -> bias + ReLU -> fp16 -> swizzled smem -> TMA store).
// HBM-bound: 12 B (3 B for uint8 frames) read + 128 B written per pixel.
//
// Images whose row pitch is not a multiple of 16 bytes (W % 4 for fp32, W % 16 for uint8) cannot be described
// to TMA; for them the producers gather straight from global memory (STAGED = false).
//
// Warps: 0-3 producer group 0, 4-7 producer group 1 (tiles alternate between the groups / A stages),
//        8 = MMA issuer + TMEM owner + weight load, 9 = input TMA, 10-13 / 14-17 epilogue groups 0 / 1 (tiles
//        alternate: one group drains accumulator stage 0, the other stage 1, each with its own two staging
//        buffers, so the TMEM-load -> pack -> store latency of consecutive tiles overlaps).
//        Persistent, grid = #SMs.
#include "common.cuh"
#include "dreamb200.h"

namespace db200 {

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                        const char* what);
int make_tensor_map_plain(CUtensorMap* tm, int kind, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, const char* what);
int device_sm_count();

struct FirstConvParams {
  const float* x;     // [B,3,H,W] fp32 (already normalised) ...
  const uint8_t* xu8; // ... or [B,H,W,3] uint8 frames, normalised on the fly with (u/255 - mean)/std
  float mean[3], stdv[3];
  const float* bias;  // [64] fp32
  int B, H, W;
  int tiles_x, tiles_y, total_tiles;
};

constexpr int kFcThreads = 18 * 32;
constexpr int kFcTw = 16, kFcTh = 8;
constexpr int kFcInStages = 4;
// staged input patch: fp32 [3][10][24] starting at column x0-4 (16-byte aligned), uint8 [10][80] starting at byte
// 3*x0-16 of the row
constexpr int kFcPatchW = 24, kFcPatchH = kFcTh + 2, kFcPatchU8 = 80;
constexpr uint32_t kFcInBytesF32 = 3 * kFcPatchH * kFcPatchW * 4;   // 2880
constexpr uint32_t kFcInBytesU8 = kFcPatchH * kFcPatchU8;           // 800
constexpr uint32_t kFcInStride = 3072;

template <bool U8, bool STAGED>
__global__ void __launch_bounds__(kFcThreads, 1)
first_conv_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmC,
                  const __grid_constant__ CUtensorMap tmX, const __grid_constant__ FirstConvParams p) {
  constexpr int kABytes = 128 * 128;
  constexpr uint32_t kIdesc = umma_idesc_f16_m128(64);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;                  // 2 x 16 KB
  const uint32_t smem_w = smem_a + 2 * kABytes;       // 8 KB (64 co x 128 B)
  const uint32_t smem_out = smem_w + 8192;            // 2 groups x 2 x 16 KB
  const uint32_t smem_in = smem_out + 4 * kABytes;    // kFcInStages x 3 KB
  const uint32_t bar_base = smem_in + kFcInStages * kFcInStride;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };          // 128 producer arrivals
  auto empty_bar = [&](int s) { return bar_base + 16u + 8u * s; };   // MMA commit
  auto tfull_bar = [&](int a) { return bar_base + 32u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 48u + 8u * a; };
  const uint32_t w_bar = bar_base + 64u;
  const uint32_t tmem_ptr_smem = bar_base + 72u;
  auto in_full = [&](int s) { return bar_base + 80u + 8u * s; };     // TMA transaction
  auto in_empty = [&](int s) { return bar_base + 112u + 8u * s; };   // 4 producer warps
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  float* lut = reinterpret_cast<float*>(smem_gen + (bar_base + 256u - smem_base));        // [3][256], U8 only
  const uint8_t* in_gen = smem_gen + (smem_in - smem_base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (U8) {
    // the same fp32 operations as torchvision's ToTensor + Normalize, so the fp16 operand is bit-identical to
    // packing the host-normalised fp32 tensor
    for (int i = threadIdx.x; i < 768; i += kFcThreads)
      lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), p.mean[i >> 8]), p.stdv[i >> 8]);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmC);
    if (STAGED) tma_prefetch_desc(&tmX);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), 128);
      mbar_init(empty_bar(s), 1);
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    for (int s = 0; s < kFcInStages; ++s) {
      mbar_init(in_full(s), 1);
      mbar_init(in_empty(s), 4);
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
    // ===================== producers: build the A tile =====================
    const int grp = warp >> 2;                 // A stage owned by this group
    // lanes 0-15 / 16-31 of a warp take tile rows two apart (0|2, 1|3, 4|6, 5|7): the patch row pitch is 24 words,
    // so the two half-warps then read disjoint shared-memory banks
    const int pw = (threadIdx.x >> 5) & 3, lx = threadIdx.x & 15;
    const int ly = (pw >> 1) * 4 + (pw & 1) + 2 * ((threadIdx.x >> 4) & 1);
    const int t = ly * kFcTw + lx;             // row of the A tile == pixel of the output tile
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
      float v[32];
      v[27] = 1.0f; v[28] = 1.0f;               // x bias_hi, bias_lo rows of the weight tile (see the MMA warp)
#pragma unroll
      for (int k = 29; k < 32; ++k) v[k] = 0.0f;
      if (STAGED) {
        const int is = it % kFcInStages;
        mbar_wait(in_full(is), (uint32_t)((it / kFcInStages) & 1));
        const uint8_t* patch = in_gen + is * kFcInStride;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
          for (int ss = 0; ss < 3; ++ss) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float f;
              if (U8) {
                // TMA zero-fills bytes outside the frame, but the conv pads the NORMALISED image with 0.0
                const int iy = oy + rr - 1, ix = ox + ss - 1;
                const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                const uint32_t u = patch[(ly + rr) * kFcPatchU8 + (lx + ss) * 3 + 13 + c];
                f = ok ? lut[c * 256 + u] : 0.0f;
              } else {
                f = reinterpret_cast<const float*>(patch)[(c * kFcPatchH + ly + rr) * kFcPatchW + lx + ss + 3];
              }
              v[(rr * 3 + ss) * 3 + c] = f;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(in_empty(is));       // values are in registers: the patch slot can be refilled
      } else {
        const float* xb = p.x + (size_t)b * 3 * plane;
        const uint8_t* ub = p.xu8 + (size_t)b * 3 * plane;
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
      // bias rides in the GEMM: k = 27 / 28 of every A row is 1.0 and the weight tile gets the bias split into two
      // fp16 terms there (hi + lo reproduces the fp32 value to ~2^-22), so the epilogue has no bias traffic at all
      for (int co = 0; co < 64; ++co) {
        const float bf = __ldg(p.bias + co);
        const __half hi = __float2half_rn(bf);
        const __half lo = __float2half_rn(bf - __half2float(hi));
        // element (row co, k) of the SWIZZLE_128B K-major tile: chunk (k/8) ^ (co & 7), k = 27 -> chunk 3 elem 3
        const uint32_t a = smem_w + (uint32_t)co * 128u + ((3u ^ (uint32_t)(co & 7)) * 16u);
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(a + 6u), "h"(__half_as_ushort(hi)) : "memory");
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(a + 8u), "h"(__half_as_ushort(lo)) : "memory");
      }
      fence_proxy_async_smem();
    }
    __syncwarp();
    {   // whole warp in uniform control flow, one elected lane issues (see elect_one in common.cuh)
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
        if (elect_one()) {
          umma_f16(d_tmem, adesc, bdesc, kIdesc, 0u);
          umma_f16(d_tmem, adesc + 2u, bdesc + 2u, kIdesc, 1u);
          umma_commit(empty_bar(s));
          umma_commit(tfull_bar(as));
        }
        ph[s] ^= 1u;
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ===================== input TMA: patches of the next kFcInStages tiles =====================
    if (STAGED) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        int r = tile;
        const int tx = r % p.tiles_x; r /= p.tiles_x;
        const int ty = r % p.tiles_y;
        const int b = r / p.tiles_y;
        const int is = it % kFcInStages;
        mbar_wait(in_empty(is), (uint32_t)(((it / kFcInStages) & 1) ^ 1));
        if (elect_one()) {
          mbar_expect_tx(in_full(is), U8 ? kFcInBytesU8 : kFcInBytesF32);
          if (U8) tma_load_3d(smem_in + is * kFcInStride, &tmX, in_full(is), tx * kFcTw * 3 - 16, ty * kFcTh - 1, b);
          else tma_load_4d(smem_in + is * kFcInStride, &tmX, in_full(is), tx * kFcTw - 4, ty * kFcTh - 1, 0, b);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 10..13 and 14..17) =====================
    const int eg = (warp - 10) >> 2;               // epilogue group == accumulator stage it drains
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - (10 + 4 * eg) * 32;
    uint32_t aphase = 0;
    uint32_t ctr = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != eg) continue;
      int r = tile;
      const int tx = r % p.tiles_x; r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      const int b = r / p.tiles_y;
      mbar_wait(tfull_bar(eg), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(eg * 64);
      const uint32_t obuf = smem_out + (uint32_t)(eg * 2 + (ctr & 1u)) * kABytes;
      if (epi_tid == 0) tma_store_wait_read<1>();
      named_bar_sync(1 + eg, 128);
      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(t_row, va);
      tmem_ld_32x32b_x32(t_row + 32u, vb);
      tmem_wait_ld();
      // all 64 accumulator columns of this row are in registers: hand the stage back before the pack / store work
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(eg));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t* v = (j < 4 ? va : vb) - (j < 4 ? 0 : 32);     // v[8j..8j+7] below (fully unrolled)
        __half2 h0 = __floats2half2_rn(fmaxf(__uint_as_float(v[8 * j + 0]), 0.0f),
                                       fmaxf(__uint_as_float(v[8 * j + 1]), 0.0f));
        __half2 h1 = __floats2half2_rn(fmaxf(__uint_as_float(v[8 * j + 2]), 0.0f),
                                       fmaxf(__uint_as_float(v[8 * j + 3]), 0.0f));
        __half2 h2 = __floats2half2_rn(fmaxf(__uint_as_float(v[8 * j + 4]), 0.0f),
                                       fmaxf(__uint_as_float(v[8 * j + 5]), 0.0f));
        __half2 h3 = __floats2half2_rn(fmaxf(__uint_as_float(v[8 * j + 6]), 0.0f),
                                       fmaxf(__uint_as_float(v[8 * j + 7]), 0.0f));
        const uint32_t chunk16 = (uint32_t)j ^ (uint32_t)(row & 7);
        const uint32_t dst = obuf + (uint32_t)row * 128u + chunk16 * 16u;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                     "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)),
                     "r"(*reinterpret_cast<uint32_t*>(&h2)), "r"(*reinterpret_cast<uint32_t*>(&h3))
                     : "memory");
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + eg, 128);
      if (epi_tid == 0) {
        tma_store_4d(&tmC, obuf, 0, tx * kFcTw, ty * kFcTh, b);
        tma_store_commit();
      }
      aphase ^= 1u;
      ++ctr;
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

template <bool U8, bool STAGED>
static int first_conv_go(const CUtensorMap& tmW, const CUtensorMap& tmC, const CUtensorMap& tmX,
                         const FirstConvParams& p, cudaStream_t stream) {
  const int smem_bytes = 1024 + 2 * 16384 + 8192 + 4 * 16384 + kFcInStages * (int)kFcInStride + 256 + 256 + 3072;
  auto kern = first_conv_kernel<U8, STAGED>;
  static bool attr_set = false;          // per instantiation
  if (!attr_set) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  const int sms = device_sm_count();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  kern<<<grid, kFcThreads, smem_bytes, stream>>>(tmW, tmC, tmX, p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

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
  CUtensorMap tmW, tmC, tmX;
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
  // TMA needs 16-byte aligned bases and row pitches; otherwise the producers gather from global memory
  static const bool allow_staged = [] { const char* e = getenv("DREAMB200_FC_STAGED"); return !(e && e[0] == '0'); }();
  bool staged = allow_staged;
  if (xu8) staged = staged && (W % 16 == 0) && (((uintptr_t)xu8 & 15) == 0);
  else staged = staged && (W % 4 == 0) && (((uintptr_t)x & 15) == 0);
  tmX = tmW;
  if (staged && xu8) {
    uint64_t dims[3] = {(uint64_t)W * 3, (uint64_t)H, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)W * 3, (uint64_t)H * W * 3};
    uint32_t box[3] = {kFcPatchU8, kFcPatchH, 1};
    if (make_tensor_map_plain(&tmX, 1, xu8, 3, dims, str, box, "first-conv uint8 frames")) return -1;
  } else if (staged) {
    uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)H * W * 12};
    uint32_t box[4] = {kFcPatchW, kFcPatchH, 3, 1};
    if (make_tensor_map_plain(&tmX, 0, x, 4, dims, str, box, "first-conv fp32 input")) return -1;
  }
  if (xu8) return staged ? first_conv_go<true, true>(tmW, tmC, tmX, p, stream)
                         : first_conv_go<true, false>(tmW, tmC, tmX, p, stream);
  return staged ? first_conv_go<false, true>(tmW, tmC, tmX, p, stream)
                : first_conv_go<false, false>(tmW, tmC, tmX, p, stream);
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
