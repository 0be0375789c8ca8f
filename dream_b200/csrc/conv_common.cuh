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
  // division-free tile decode: q = (x * magic) >> 40 is exact for x < 2^24, divisor < 2^16 (host: 2^40/d + 1)
  unsigned long long mg_n, mg_x, mg_y;
  float* absmax;    // optional: running max |output| (bits of a non-negative float), see dreamb200_conv_desc
  const __half* gate;      // optional ReLU gate of the backward pass: output zeroed where gate <= 0
  const float* out_scale;  // optional device scalar multiplied into every output
  float* colsum;           // optional [Cout_pad]: += sum over output pixels (after gate / scale): a bias gradient
  // phase groups (conv_tc2.cu): `phases` sub-pixel convolutions in one launch, tile index = phase * tiles_per_phase + q;
  // the tap tables above then hold phases * taps entries, phase-major.  0 / 1 = an ordinary convolution.
  int phases, tiles_per_phase;
  unsigned long long mg_phase;
  int res_slots;    // conv_tc2: > 0 = the fp32 residual is staged through a shared-memory ring of this many chunk slots
  int res_inplace;  // ... and the fp32 output leaves through the same slots (TMA stores)
};

__host__ __device__ inline unsigned long long div_magic(int d) { return (1ull << 40) / (unsigned long long)d + 1ull; }
__device__ __forceinline__ int fast_div(int x, unsigned long long magic) {
  return (int)(((unsigned long long)(unsigned)x * magic) >> 40);
}
// tile index -> (output-channel tile n, patch column tx, patch row ty, image b); n varies fastest
__device__ __forceinline__ void decode_tile(const ConvParams& p, int tile, int& n, int& tx, int& ty, int& b) {
  int t = fast_div(tile, p.mg_n);
  n = tile - t * p.n_tiles;
  int t2 = fast_div(t, p.mg_x);
  tx = t - t2 * p.tiles_x;
  b = fast_div(t2, p.mg_y);
  ty = t2 - b * p.tiles_y;
}

constexpr int kThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 rows x 64 fp16
constexpr int kStageOutBytes = 128 * 128;
constexpr int kPoolBytes = 32 * 128;    // pooled tile: <= 32 rows x 64 fp16

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}


// fp32 residual tiles staged in shared memory by the TMA producer warp (conv_tc2.cu): a ring of `slots` chunk slots,
// each two 128-row x 32-float sub-tiles (16 KB, SWIZZLE_128B) = the residual of one 64-column output chunk.  Every
// epilogue warp keeps its own (idx, phase) cursor; a slot is handed back once all `4 * SPLIT` warps have read it.
struct ResRing {
  uint32_t smem;          // shared address of slot 0 (1024-byte aligned)
  uint32_t full_bar;      // mbarrier of slot 0 (TMA transaction bytes); slot s at + 8 * s
  uint32_t empty_bar;     // mbarrier of slot 0 (count = epilogue warps)
  int slots;
  int idx;
  uint32_t phase;
  // in-place output (y_tm != nullptr): the fp32 result of a chunk is written back over its residual in the slot and
  // leaves through TMA stores issued with the chunk's fp16 store; the slot returns to the producer only when that
  // bulk group has been read (one arrival, by the thread that issues the stores), two chunks later.
  const CUtensorMap* y_tm;
  int rel_idx;            // next slot to release
  int pending;            // chunks whose stores have been issued but whose slots are not released yet
};
constexpr uint32_t kResSubBytes = 128 * 128;            // one sub-tile: 128 rows x 32 fp32
constexpr uint32_t kResSlotBytes = 2 * kResSubBytes;

// ReLU gate tiles of a data-gradient launch staged in shared memory by a dedicated TMA warp (conv_rs2.cu): a ring of
// `slots` 64-column chunks (128 rows x 128 B, SWIZZLE_128B -- the layout of the output staging buffer).  Every
// epilogue warp keeps its own cursor; a slot goes back to the TMA warp once all epilogue warps have read it.
struct GateRing {
  uint32_t smem, full_bar, empty_bar;
  int slots, idx;
  uint32_t phase;
};
constexpr uint32_t kGateSlotBytes = 128 * 128;

// One output tile (128 accumulator rows x BLOCK_N columns) of the NHWC fp16 path; called by the 4 epilogue
// warps (128 threads, named barrier 1).  `t_row` = TMEM address of this thread's lane quarter / accumulator stage,
// `smem_bias` = shared-memory copy of the current channel tile's fp32 bias (BLOCK_N floats, zeros when none).
// Per 64-column chunk: TMEM -> registers -> +bias (+residuals) -> ReLU -> fp16 (the accumulator stage is handed
// back to the MMA warp right after the last TMEM read, BEFORE any wait on the staging buffers) -> staging smem
// (128B-swizzled) -> TMA store.  The fused 2x2 max pool runs on the packed fp16 values with warp shuffles when
// the tile is 8 or 16 pixels wide (window partners are lanes ^1 and ^tw), else through a second smem pass.
// SPLIT = 1: 4 epilogue warps, each thread handles all 64 columns of a chunk.  SPLIT = 2: 8 epilogue warps, two
// per TMEM lane quarter, each thread handles 32 of the 64 columns (`hsel`) -- halves the per-tile epilogue latency
// for the narrow layers where the epilogue, not the MMA, paces the tile.
// CLUSTER_ARRIVE: `tempty_bar_addr` is a shared::cluster address (the leader CTA's barrier of a cta_group::2 pair,
// see conv_rs2.cu) and the accumulator hand-back is a cluster-scope arrive.
// PLAIN: the launch has no residual / fp32 stream / gate / out_scale / colsum / absmax (every inference layer of the
// vgg networks): those options compile away -- the epilogue of the 64-channel layers paces the kernel (two warps
// per scheduler, one dependent chain per tile), so every runtime test on the chain counts.
// GRING: the ReLU gate always arrives through a GateRing (`gr`), the per-thread gate loads compile away.
// hand_back = false: the caller has more columns of this accumulator stage to drain (conv_rs3.cu) and returns it itself.
template <int BLOCK_N, int SPLIT = 1, bool CLUSTER_ARRIVE = false, bool PLAIN = false, bool GRING = false>
__device__ __forceinline__ void epilogue_nhwc_tile(const ConvParams& p, const CUtensorMap* tmC, const CUtensorMap* tmP,
                                                   uint32_t t_row, uint32_t smem_out, uint32_t smem_pool,
                                                   uint32_t smem_bias, float* smem_bias_gen, uint32_t tempty_bar_addr,
                                                   int n, int tx, int ty, int b, int ox, int oy, bool valid, int row,
                                                   int lane, int epi_tid, uint32_t& chunk_ctr, int hsel = 0,
                                                   float* csum = nullptr, const float* breg = nullptr,
                                                   ResRing* rr = nullptr, uint32_t tfull_addr = 0u,
                                                   uint32_t tfull_phase = 0u, GateRing* gr = nullptr,
                                                   bool hand_back = true) {
  constexpr int kEpiThreads = 128 * SPLIT;
  constexpr int kRegs = 32 / SPLIT;                    // packed fp16 pairs per thread per chunk
  if (p.n_tiles > 1) {
    // several output-channel tiles per CTA: re-stage this tile's BLOCK_N bias values (smem holds one tile's worth)
    named_bar_sync(1, kEpiThreads);
    for (int i = epi_tid; i < BLOCK_N; i += kEpiThreads)
      smem_bias_gen[i] = p.bias != nullptr ? __ldg(p.bias + n * BLOCK_N + i) : 0.0f;
    named_bar_sync(1, kEpiThreads);
  }
  const __half* res_row = nullptr;
  const float* res32_row = nullptr;
  float* y32_row = nullptr;
  const __half* gate_row = nullptr;
  float out_scale = 1.0f;
  if constexpr (!PLAIN) {
    const size_t row_off = ((size_t)((size_t)b * p.Ho + oy) * p.Wo + ox) * p.Cout_pad + n * BLOCK_N;
    if (p.residual != nullptr && valid) res_row = p.residual + row_off;
    if (p.residual_f32 != nullptr && valid) res32_row = p.residual_f32 + row_off;
    if (p.y_f32 != nullptr && valid) y32_row = p.y_f32 + row_off;
    gate_row = (p.gate != nullptr && valid) ? p.gate + row_off : nullptr;
    out_scale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.0f;
  }
  const bool shfl_pool = p.pool && (p.tw == 8 || p.tw == 16);
  const bool write_full = !p.pool || p.store_full || !shfl_pool;
  // fp32 identity stream (ResNet bottlenecks): the 1x1 expansion layers are HBM-bound on this read -- 4 bytes per
  // output element against 2 * Cin MACs -- and the loads used to be issued only after the TMEM read of the same
  // chunk, one dependent round trip per 64 columns.  Now the residual of chunk c+1 is in flight while chunk c is
  // converted, stored and handed to the TMA (double-buffered in registers).
  float4 rpre[2 / SPLIT][8];
  auto prefetch_res32 = [&](int c) {
#pragma unroll
    for (int hh = 0; hh < 2 / SPLIT; ++hh) {
      const int h = SPLIT == 2 ? hsel : hh;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        rpre[hh][i] = res32_row != nullptr ? __ldg(reinterpret_cast<const float4*>(res32_row + c * 64 + h * 32) + i)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if constexpr (!PLAIN) {
    if (p.residual_f32 != nullptr && rr == nullptr) prefetch_res32(0);
  }
  // ReLU gate of a data-gradient launch (the forward output of the layer below, long evicted from L2): one HBM round
  // trip per chunk when it is read after the TMEM load -- the gated 64-channel data gradients ran at half the speed of
  // the same convolution forward.  The gate does not depend on the accumulator, so chunk 0's values are requested
  // BEFORE the wait for the tile's MMAs (`tfull_addr`, when the caller delegates that wait) and chunk c+1's while
  // chunk c is finished.
  uint4 gpre[2 / SPLIT][4];
  auto prefetch_gate = [&](int c) {
#pragma unroll
    for (int hh = 0; hh < 2 / SPLIT; ++hh) {
      const int h = SPLIT == 2 ? hsel : hh;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        gpre[hh][i] = gate_row != nullptr ? __ldg(reinterpret_cast<const uint4*>(gate_row + c * 64 + h * 32) + i)
                                          : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  if constexpr (!PLAIN && !GRING) {
    if (p.gate != nullptr && gr == nullptr) prefetch_gate(0);
  }
  if (tfull_addr != 0u) {
    mbar_wait(tfull_addr, tfull_phase);
    tc_fence_after();
  }
#pragma unroll 1
  for (int c = 0; c < BLOCK_N / 64; ++c, ++chunk_ctr) {
    const uint32_t obuf = smem_out + (chunk_ctr & 1u) * kStageOutBytes;
    const uint32_t pbuf = smem_pool + (chunk_ctr & 1u) * kPoolBytes;
    if constexpr (PLAIN) {
      if (shfl_pool && !p.store_full) {
        // Pool-only tiles (every pooled layer in inference): max FIRST, on the fp32 accumulators, then bias + ReLU +
        // fp16 on the pooled quarter only.  Bit-identical to pooling the finished fp16 values -- x -> fp16(relu(x + b))
        // is monotonic, so it commutes with max -- at less than half the instructions: the 2x2 window's four lanes
        // (l, l^1, l^tw, l^tw^1) swap HALVES of their 32 columns (16 + 8 shuffles) so each ends up owning 8 pooled
        // channels, instead of every lane finishing all 32 channels and three of four throwing them away.
        uint32_t packed[2 / SPLIT][4];
        const uint32_t odd = (uint32_t)lane & 1u, up = ((uint32_t)lane & (uint32_t)p.tw) ? 1u : 0u;
#pragma unroll
        for (int hh = 0; hh < 2 / SPLIT; ++hh) {
          const int h = SPLIT == 2 ? hsel : hh;
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 64 + h * 32), v);
          tmem_wait_ld();
          float a[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float send = __uint_as_float(odd ? v[i] : v[16 + i]);
            const float keep = __uint_as_float(odd ? v[16 + i] : v[i]);
            a[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
          }
          float q[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float send = up ? a[i] : a[8 + i];
            const float keep = up ? a[8 + i] : a[i];
            q[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, p.tw));
          }
          if (p.bias != nullptr) {
            const uint32_t bias_addr = smem_bias + (uint32_t)(c * 64 + h * 32 + (int)(odd * 16u + up * 8u)) * 4u;
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
              float b0, b1, b2, b3;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(bias_addr + (uint32_t)i * 4u));
              q[i] += b0; q[i + 1] += b1; q[i + 2] += b2; q[i + 3] += b3;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = fmaxf(q[i], 0.0f);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) packed[hh][i] = pack_h2(q[2 * i], q[2 * i + 1]);
        }
        if (c == BLOCK_N / 64 - 1 && hand_back) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CLUSTER_ARRIVE) mbar_arrive_cluster(tempty_bar_addr); else mbar_arrive(tempty_bar_addr);
          }
        }
        if (epi_tid < 32) {                              // the store that used pbuf two chunks ago has read it
          if (elect_one()) tma_store_wait_read<1>();
        }
        named_bar_sync(1, kEpiThreads);
        {
          const int sh = p.tw == 8 ? 3 : 4;
          const int ly = row >> sh, lx = row & (p.tw - 1);
          const uint32_t pr = (uint32_t)((ly >> 1) * (p.tw >> 1) + (lx >> 1));
#pragma unroll
          for (int hh = 0; hh < 2 / SPLIT; ++hh) {
            const uint32_t h = SPLIT == 2 ? (uint32_t)hsel : (uint32_t)hh;
            const uint32_t j = h * 4u + odd * 2u + up;    // 16-byte chunk (8 channels) of the pooled pixel's row
            const uint32_t dst = pbuf + pr * 128u + ((j ^ (pr & 7u)) * 16u);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[hh][0]), "r"(packed[hh][1]),
                         "r"(packed[hh][2]), "r"(packed[hh][3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, kEpiThreads);
        if (epi_tid < 32) {
          if (elect_one()) {
            tma_store_4d(tmP, pbuf, n * BLOCK_N + c * 64, tx * (p.tw >> 1), ty * (p.th >> 1), b);
            tma_store_commit();
          }
        }
        continue;
      }
    }
    uint32_t hv[kRegs];                                // this thread's output channels of the pixel, packed fp16
#pragma unroll
    for (int hh = 0; hh < 2 / SPLIT; ++hh) {
      const int h = SPLIT == 2 ? hsel : hh;
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 64 + h * 32), v);
      tmem_wait_ld();
      float f[32];
      if (p.bias == nullptr) {
        // (data-gradient launches have no bias: skip the shared-memory reads -- a broadcast LDS.128 still costs
        //  four wavefronts, and the 64-channel layers are bound by shared-memory bandwidth)
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
      } else if (BLOCK_N == 64 && SPLIT == 2 && breg != nullptr) {
        // single channel tile, 32 columns per thread: the bias lives in registers for the CTA's lifetime
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + breg[i];
      } else {
        const uint32_t bias_addr = smem_bias + (uint32_t)(c * 64 + h * 32) * 4u;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float b0, b1, b2, b3;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(bias_addr + (uint32_t)i * 4u));
          f[i] = __uint_as_float(v[i]) + b0;
          f[i + 1] = __uint_as_float(v[i + 1]) + b1;
          f[i + 2] = __uint_as_float(v[i + 2]) + b2;
          f[i + 3] = __uint_as_float(v[i + 3]) + b3;
        }
      }
      if (!PLAIN && res_row != nullptr) {
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
      if (!PLAIN && p.residual_f32 != nullptr && rr != nullptr) {
        // residual chunk staged by the producer warp: this thread's 32 floats are 8 swizzled 16-byte pieces of its row
        if (hh == 0) mbar_wait(rr->full_bar + 8u * (uint32_t)rr->idx, rr->phase);
        const uint32_t rbase = rr->smem + (uint32_t)rr->idx * kResSlotBytes + (uint32_t)h * kResSubBytes +
                               (uint32_t)row * 128u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float r0, r1, r2, r3;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(r3)
                       : "r"(rbase + (((uint32_t)i ^ (uint32_t)(row & 7)) * 16u)));
          f[4 * i] += r0; f[4 * i + 1] += r1; f[4 * i + 2] += r2; f[4 * i + 3] += r3;
        }
        if (hh == 2 / SPLIT - 1 && rr->y_tm == nullptr) {
          __syncwarp();
          if (lane == 0) mbar_arrive(rr->empty_bar + 8u * (uint32_t)rr->idx);
          if (++rr->idx == rr->slots) { rr->idx = 0; rr->phase ^= 1u; }
        }
      } else if (!PLAIN && p.residual_f32 != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 rv = rpre[hh][i >> 2];
          f[i] += rv.x; f[i + 1] += rv.y; f[i + 2] += rv.z; f[i + 3] += rv.w;
        }
        if (hh == 2 / SPLIT - 1 && c + 1 < BLOCK_N / 64) prefetch_res32(c + 1);
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
      }
      if (!PLAIN && p.gate != nullptr) {
        // data gradient leaving through the previous layer's ReLU: keep it where that layer's output was > 0
        if ((GRING || gr != nullptr) && hh == 0) mbar_wait(gr->full_bar + 8u * (uint32_t)gr->idx, gr->phase);
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 gv;
          if (GRING || gr != nullptr) {
            // staged tile: this thread's row, 16-byte piece h * 4 + i / 8 (swizzled like the output staging buffer);
            // rows outside the image were zero-filled by the TMA, i.e. gated off like `gate_row == nullptr`
            const uint32_t src = gr->smem + (uint32_t)gr->idx * kGateSlotBytes + (uint32_t)row * 128u +
                                 ((((uint32_t)h * 4u + (uint32_t)(i >> 3)) ^ (uint32_t)(row & 7)) * 16u);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(gv.x), "=r"(gv.y), "=r"(gv.z), "=r"(gv.w) : "r"(src) : "memory");
          } else {
            gv = gpre[hh][i >> 3];
          }
          const __half2* gh = reinterpret_cast<const __half2*>(&gv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 gf = __half22float2(gh[j]);
            f[i + 2 * j] = gf.x > 0.0f ? f[i + 2 * j] * out_scale : 0.0f;
            f[i + 2 * j + 1] = gf.y > 0.0f ? f[i + 2 * j + 1] * out_scale : 0.0f;
          }
        }
        if (GRING || gr != nullptr) {
          if (hh == 2 / SPLIT - 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(gr->empty_bar + 8u * (uint32_t)gr->idx);
            if (++gr->idx == gr->slots) { gr->idx = 0; gr->phase ^= 1u; }
          }
        } else if (!GRING && hh == 2 / SPLIT - 1 && c + 1 < BLOCK_N / 64) {
          prefetch_gate(c + 1);
        }
      } else if (!PLAIN && p.out_scale != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] *= out_scale;
      }
      if (!PLAIN && rr != nullptr && rr->y_tm != nullptr) {
        // fp32 copy of the output: back into this thread's own 8 pieces of the residual slot (read above), stored by
        // TMA below.  (Per-thread STG.128s -- 32 rows x 16 bytes per instruction -- back-pressured the whole epilogue:
        // ncu showed it waiting on the store queue to release the source registers.)
        const uint32_t rbase = rr->smem + (uint32_t)rr->idx * kResSlotBytes + (uint32_t)h * kResSubBytes +
                               (uint32_t)row * 128u;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (((uint32_t)i ^ (uint32_t)(row & 7)) * 16u)),
                       "f"(f[4 * i]), "f"(f[4 * i + 1]), "f"(f[4 * i + 2]), "f"(f[4 * i + 3])
                       : "memory");
      } else if (!PLAIN && y32_row != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(y32_row + c * 64 + h * 32 + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
      }
      if (!PLAIN && p.colsum != nullptr) {
        // column sums over the warp's 32 rows by a halving butterfly (31 shuffles): lane i ends up with column i
        auto warp_colsum = [&]() {
          float r[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = valid ? f[i] : 0.0f;
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int k = 0; k < off; ++k) {
              const float send = up ? r[k] : r[k + off];
              const float keep = up ? r[k + off] : r[k];
              r[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          return r[0];
        };
        if constexpr (BLOCK_N == 64 && SPLIT == 2) {
          // single 64-channel tile, 32 fixed columns per thread for the CTA's whole life: keep per-thread running
          // column sums in registers and fold the warp's 32 rows ONCE, in flush_colsum, instead of a butterfly per
          // tile (the 64-channel data gradients at 400x400 were paced by exactly that: 3.6 ms against 1.6 ms forward)
          if (p.n_tiles == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) csum[i] += valid ? f[i] : 0.0f;
          } else {
            atomicAdd(p.colsum + n * BLOCK_N + c * 64 + h * 32 + lane, warp_colsum());
          }
        } else {
          const float r0 = warp_colsum();
          if (p.n_tiles == 1) csum[c * 2 + h] += r0;       // single channel tile: accumulate across the CTA's tiles
          else atomicAdd(p.colsum + n * BLOCK_N + c * 64 + h * 32 + lane, r0);
        }
      }
      bool running_absmax = false;
      if constexpr (!PLAIN && BLOCK_N == 64 && SPLIT == 2) {
        if (p.absmax != nullptr && p.n_tiles == 1) {
          // (same idea: a per-thread running maximum, reduced and published once per CTA in flush_colsum)
          running_absmax = true;
          float m = csum[32];
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) m = fmaxf(m, fabsf(f[i]));
          }
          csum[32] = m;
        }
      }
      if (!PLAIN && p.absmax != nullptr && !running_absmax) {
        float m = 0.0f;
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, fabsf(f[i]));
        }
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) {
          if (!(m <= 3.0e38f)) m = 3.0e38f;
          atomicMax(reinterpret_cast<int*>(p.absmax), __float_as_int(m));
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) hv[hh * 16 + i] = pack_h2(f[2 * i], f[2 * i + 1]);
    }
    if (c == BLOCK_N / 64 - 1 && hand_back) {
      // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp right away
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CLUSTER_ARRIVE) mbar_arrive_cluster(tempty_bar_addr); else mbar_arrive(tempty_bar_addr);
      }
    }
    // (bulk-store bookkeeping by ONE lane of the first epilogue warp: elect.sync picks the same lane every time, and
    //  issued from warp-uniform control flow the TMA instructions need no per-instruction election loop)
    if (epi_tid < 32) {                                // stores that used obuf / pbuf two chunks ago have read them
      if (elect_one()) {
        if (p.pool && p.store_full) tma_store_wait_read<2>(); else tma_store_wait_read<1>();
        if constexpr (!PLAIN) {
          if (rr != nullptr && rr->y_tm != nullptr && rr->pending == 2) {
            // ... and so has the fp32 store of the chunk two back: its slot goes back to the producer
            mbar_arrive(rr->empty_bar + 8u * (uint32_t)rr->rel_idx);
          }
        }
      }
    }
    if constexpr (!PLAIN) {
      if (rr != nullptr && rr->y_tm != nullptr && rr->pending == 2) {   // (cursor kept by every thread alike)
        if (++rr->rel_idx == rr->slots) rr->rel_idx = 0;
        rr->pending = 1;
      }
    }
    named_bar_sync(1, kEpiThreads);
    const uint32_t j0 = SPLIT == 2 ? (uint32_t)hsel * 4u : 0u;      // first 16-byte chunk this thread owns
    if (write_full) {
#pragma unroll
      for (int j = 0; j < 8 / SPLIT; ++j) {
        const uint32_t dst = obuf + (uint32_t)row * 128u + (((j0 + (uint32_t)j) ^ (uint32_t)(row & 7)) * 16u);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hv[4 * j]), "r"(hv[4 * j + 1]),
                     "r"(hv[4 * j + 2]), "r"(hv[4 * j + 3])
                     : "memory");
      }
    }
    if (shfl_pool) {
      // 2x2 window = lanes {l, l^1, l^tw, l^tw^1}; the even-x / even-y lane keeps the maximum
#pragma unroll
      for (int i = 0; i < kRegs; ++i) {
        __half2 m = *reinterpret_cast<__half2*>(&hv[i]);
        uint32_t o = __shfl_xor_sync(0xffffffffu, hv[i], 1);
        m = __hmax2(m, *reinterpret_cast<__half2*>(&o));
        uint32_t mm = *reinterpret_cast<uint32_t*>(&m);
        o = __shfl_xor_sync(0xffffffffu, mm, p.tw);
        m = __hmax2(m, *reinterpret_cast<__half2*>(&o));
        hv[i] = *reinterpret_cast<uint32_t*>(&m);
      }
      if ((lane & 1) == 0 && (lane & p.tw) == 0) {
        const int sh = p.tw == 8 ? 3 : 4;                       // tw is 8 or 16 here
        const int ly = row >> sh, lx = row & (p.tw - 1);
        const int pr = (ly >> 1) * (p.tw >> 1) + (lx >> 1);
#pragma unroll
        for (int j = 0; j < 8 / SPLIT; ++j) {
          const uint32_t dst = pbuf + (uint32_t)pr * 128u + (((j0 + (uint32_t)j) ^ (uint32_t)(pr & 7)) * 16u);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hv[4 * j]), "r"(hv[4 * j + 1]),
                       "r"(hv[4 * j + 2]), "r"(hv[4 * j + 3])
                       : "memory");
        }
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, kEpiThreads);
    if (epi_tid < 32 && (!p.pool || p.store_full)) {
      if (elect_one()) {
        tma_store_4d(tmC, obuf, n * BLOCK_N + c * 64, tx * p.tw, ty * p.th, b);
        if constexpr (!PLAIN) {
          if (rr != nullptr && rr->y_tm != nullptr) {
            const uint32_t src = rr->smem + (uint32_t)rr->idx * kResSlotBytes;
            tma_store_4d(rr->y_tm, src, n * BLOCK_N + c * 64, tx * p.tw, ty * p.th, b);
            tma_store_4d(rr->y_tm, src + kResSubBytes, n * BLOCK_N + c * 64 + 32, tx * p.tw, ty * p.th, b);
          }
        }
        tma_store_commit();
      }
    }
    if constexpr (!PLAIN) {
      if (rr != nullptr && rr->y_tm != nullptr) {
        if (++rr->idx == rr->slots) { rr->idx = 0; rr->phase ^= 1u; }
        ++rr->pending;
      }
    }
    if (p.pool && !shfl_pool) {
      // generic tile shape: 2x2 max over the tile that now sits in obuf (pooled row pr, 16 B chunk ch per item)
      const int ptw = p.tw >> 1;
      const int items = ptw * (p.th >> 1) * 8;
      for (int item = epi_tid; item < items; item += kEpiThreads) {
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
      named_bar_sync(1, kEpiThreads);
    }
    if (p.pool && epi_tid < 32) {
      if (elect_one()) {
        tma_store_4d(tmP, pbuf, n * BLOCK_N + c * 64, tx * (p.tw >> 1), ty * (p.th >> 1), b);
        tma_store_commit();
      }
    }
  }
}

// After a CTA's last tile: add the per-warp column sums kept in `csum` (see epilogue_nhwc_tile) to p.colsum.
// csum holds kCsumSize<BLOCK_N, SPLIT> floats: 8 butterfly results, or -- single 64-channel tile with two warps per lane
// quarter -- 32 per-thread running column sums + 1 running max |output|.
template <int BLOCK_N, int SPLIT>
constexpr int kCsumSize = (BLOCK_N == 64 && SPLIT == 2) ? 33 : 8;

template <int BLOCK_N, int SPLIT>
__device__ __forceinline__ void flush_colsum(const ConvParams& p, float* csum, int lane, int hsel) {
  if constexpr (BLOCK_N == 64 && SPLIT == 2) {
    if (p.n_tiles != 1) return;                          // several channel tiles: the epilogue added per tile
    if (p.colsum != nullptr) {
      float r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = csum[i];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
          const float send = up ? r[k] : r[k + off];
          const float keep = up ? r[k + off] : r[k];
          r[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      atomicAdd(p.colsum + hsel * 32 + lane, r[0]);
    }
    if (p.absmax != nullptr) {
      float m = csum[32];
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) {
        if (!(m <= 3.0e38f)) m = 3.0e38f;
        atomicMax(reinterpret_cast<int*>(p.absmax), __float_as_int(m));
      }
    }
  } else {
    if (p.colsum == nullptr || p.n_tiles != 1) return;
#pragma unroll
    for (int c = 0; c < BLOCK_N / 64; ++c)
#pragma unroll
      for (int hh = 0; hh < 2 / SPLIT; ++hh) {
        const int h = SPLIT == 2 ? hsel : hh;
        atomicAdd(p.colsum + c * 64 + h * 32 + lane, csum[c * 2 + h]);
      }
  }
}

// Cooperative copy of the first output-channel tile's fp32 bias into shared memory (call before the prologue
// __syncthreads(); layers with a single channel tile never touch it again).
__device__ __forceinline__ void stage_bias(const ConvParams& p, float* sb, int block_n) {
  for (int i = threadIdx.x; i < block_n; i += blockDim.x) sb[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.0f;
}

}  // namespace db200
