// conv_common.cuh -- pieces shared by the tensor-core convolution kernels (conv_tc.cu, conv_rs.cu):
// the parameter block and the NHWC epilogue (TMEM -> +bias (+residual) -> ReLU -> fp16 -> 128B-swizzled smem ->
// TMA store, with the optional fused 2x2 max pool and the optional fp32 stream copy).
#pragma once
#include "common.cuh"
#include "dreamb200.h"

namespace db200 {

struct ConvParams {
  int tiles_x, tiles_y, n_tiles, total_tiles;
  int tw, th;
  int B, Ho, Wo;
  int in_stride;
  int taps, kchunks;
  int8_t dy[DREAMB200_MAX_TAPS], dx[DREAMB200_MAX_TAPS];
  const float* bias;
  const __half* residual;
  const float* residual_f32;   // fp32 NHWC residual stream (ResNet identity path), or NULL
  float* y_f32;                // optional fp32 NHWC copy of the output (next block's identity)
  int Cout_pad;
  int relu;
  float* out_f32;
  int cout_real;
  int stages;
  int pool;         // fused 2x2/s2 max pool of the output tile (tw, th even): pooled tile -> tmP
  int store_full;   // also store the un-pooled tile through tmC
};

constexpr int kThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 rows x 64 fp16
constexpr int kStageOutBytes = 128 * 128;
constexpr int kPoolBytes = 32 * 128;    // pooled tile: <= 32 rows x 64 fp16

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}


// One output tile (128 accumulator rows x BLOCK_N columns) of the NHWC fp16 path; called by the 4 epilogue
// warps (128 threads, named barrier 1).  `t_row` = TMEM address of this thread's lane quarter / accumulator stage.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_nhwc_tile(const ConvParams& p, const CUtensorMap* tmC, const CUtensorMap* tmP,
                                                   uint32_t t_row, uint32_t smem_out, uint32_t smem_pool,
                                                   uint32_t tempty_bar_addr, int n, int tx, int ty, int b, int ox,
                                                   int oy, bool valid, int row, int lane, int epi_tid,
                                                   uint32_t& chunk_ctr) {
    const __half* res_row = nullptr;
    const float* res32_row = nullptr;
    float* y32_row = nullptr;
    const size_t row_off = ((size_t)((size_t)b * p.Ho + oy) * p.Wo + ox) * p.Cout_pad + n * BLOCK_N;
    if (p.residual != nullptr && valid) res_row = p.residual + row_off;
    if (p.residual_f32 != nullptr && valid) res32_row = p.residual_f32 + row_off;
    if (p.y_f32 != nullptr && valid) y32_row = p.y_f32 + row_off;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 64; ++c, ++chunk_ctr) {
      const uint32_t obuf = smem_out + (chunk_ctr & 1u) * kStageOutBytes;
      const uint32_t pbuf = smem_pool + (chunk_ctr & 1u) * kPoolBytes;
      if (epi_tid == 0) {                            // stores that used obuf / pbuf two chunks ago have read them
        if (p.pool && p.store_full) tma_store_wait_read<2>(); else tma_store_wait_read<1>();
      }
      named_bar_sync(1, 128);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 64 + h * 32), v);
        tmem_wait_ld();
        const int ch0 = n * BLOCK_N + c * 64 + h * 32;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + i));
            f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
          }
        }
        if (res_row != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const uint4 rv = __ldg(reinterpret_cast<const uint4*>(res_row + c * 64 + h * 32 + i));
            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 rf = __half22float2(rh[j]);
              f[i + 2 * j] += rf.x;
              f[i + 2 * j + 1] += rf.y;
            }
          }
        }
        if (res32_row != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 rv = __ldg(reinterpret_cast<const float4*>(res32_row + c * 64 + h * 32 + i));
            f[i] += rv.x; f[i + 1] += rv.y; f[i + 2] += rv.z; f[i + 3] += rv.w;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
        }
        if (y32_row != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(y32_row + c * 64 + h * 32 + i) =
                make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w0 = pack_h2(f[8 * j + 0], f[8 * j + 1]);
          const uint32_t w1 = pack_h2(f[8 * j + 2], f[8 * j + 3]);
          const uint32_t w2 = pack_h2(f[8 * j + 4], f[8 * j + 5]);
          const uint32_t w3 = pack_h2(f[8 * j + 6], f[8 * j + 7]);
          const uint32_t chunk16 = (uint32_t)(h * 4 + j) ^ (uint32_t)(row & 7);
          const uint32_t dst = obuf + (uint32_t)row * 128u + chunk16 * 16u;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w0), "r"(w1),
                       "r"(w2), "r"(w3)
                       : "memory");
        }
      }
      if (c == BLOCK_N / 64 - 1) {
        // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar_addr);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (epi_tid == 0 && (!p.pool || p.store_full)) {
        tma_store_4d(tmC, obuf, n * BLOCK_N + c * 64, tx * p.tw, ty * p.th, b);
        tma_store_commit();
      }
      if (p.pool) {
        // 2x2 max over the tile that now sits in obuf: pooled row pr, 16 B chunk ch per work item
        const int ptw = p.tw >> 1;
        const int items = ptw * (p.th >> 1) * 8;
        for (int item = epi_tid; item < items; item += 128) {
          const int pr = item >> 3, ch = item & 7;
          const int py = pr / ptw, px = pr - py * ptw;
          const int r00 = (2 * py) * p.tw + 2 * px;
          __half2 m[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = r00 + (k >> 1) * p.tw + (k & 1);
            uint32_t a0, a1, a2, a3;
            const uint32_t src = obuf + (uint32_t)r * 128u + (((uint32_t)ch ^ (uint32_t)(r & 7)) * 16u);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(src) : "memory");
            const __half2 h0 = *reinterpret_cast<__half2*>(&a0), h1 = *reinterpret_cast<__half2*>(&a1);
            const __half2 h2 = *reinterpret_cast<__half2*>(&a2), h3 = *reinterpret_cast<__half2*>(&a3);
            if (k == 0) { m[0] = h0; m[1] = h1; m[2] = h2; m[3] = h3; }
            else { m[0] = __hmax2(m[0], h0); m[1] = __hmax2(m[1], h1); m[2] = __hmax2(m[2], h2); m[3] = __hmax2(m[3], h3); }
          }
          const uint32_t dst = pbuf + (uint32_t)pr * 128u + (((uint32_t)ch ^ (uint32_t)(pr & 7)) * 16u);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                       "r"(*reinterpret_cast<uint32_t*>(&m[0])), "r"(*reinterpret_cast<uint32_t*>(&m[1])),
                       "r"(*reinterpret_cast<uint32_t*>(&m[2])), "r"(*reinterpret_cast<uint32_t*>(&m[3]))
                       : "memory");
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (epi_tid == 0) {
          tma_store_4d(tmP, pbuf, n * BLOCK_N + c * 64, tx * ptw, ty * (p.th >> 1), b);
          tma_store_commit();
        }
      }
    }
}

}  // namespace db200
