// Weight gradients of the three fused MLPs on the tensor cores (training side of kernel 3):
//
//     dW_l = (1 / loss_scale) * G_l^T . X_l          for the nine layers l,   summed over all n samples
//
// where X [n, 624] fp16 holds the forward activations saved by apnerf_field_forward_train and G [n, 576] fp16 the
// activation gradients x loss_scale written by apnerf_field_backward (column maps below; both row-major, one row per
// sample).  Replaces the weight-gradient half of tcnn's FullyFusedMLP backward behind the reference's loss.backward()
// (scripts/pipeline.py:518; modules perception/models/radiance_fields/ngp.py:123-169), which tcnn computes with CUTLASS
// split-K GEMMs; round 1 used one library GEMM (torch.bmm) here.
//
// Shape of the problem: nine small outputs (at most 128 x 128), a huge reduction dimension (the samples) -- a pure
// split-K job.  Persistent CTAs (one per SM) each stream a slice of the samples through shared memory in stages of 32
// rows (cp.async, two stages in flight) and keep ALL nine accumulators in TMEM (432 of the 512 columns) for their
// whole slice; one elected thread issues 18 tcgen05.mma (M = 128, K = 16) per stage.  Both operands are read
// "MN-major": a stage is stored as [16-byte feature chunk][sample][8 features], which is a straight 16-byte copy of
// the row-major global data (no transpose) and is the canonical no-swizzle MN-major UMMA layout with
// SBO = 512 B (between feature chunks) and LBO = 128 B (between groups of 8 samples)
// (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>; instruction-descriptor bits 15 / 16).
// The M side of every product is 128 features wide; where the layer has only 64 (or the block sits at the end of a
// row) the extra accumulator rows are fed by whatever follows in shared memory and are never read back.  Layers whose
// OUTPUT width is the small dimension (16 / 32) are computed transposed (M = inputs, N = outputs).
// At the end every CTA adds its partial sums to the fp32 gradient vectors with vector atomics (red.global.add.v4.f32).
#pragma once
#include "field.cuh"
#include "ptx.cuh"

namespace apnerf {

constexpr int WG_THREADS = 256;
constexpr int WG_STAGE_ROWS = 32;                      // samples per stage (two K = 16 MMA steps)
constexpr int WG_G_WIDTH = 576, WG_X_WIDTH = 624;      // halves per row (radiance_fields/ngp.py: _G_COLS / _X_COLS)
constexpr int WG_G_CHUNKS = WG_G_WIDTH / 8, WG_X_CHUNKS = WG_X_WIDTH / 8;  // 72, 78
constexpr int WG_CHUNK_BYTES = WG_STAGE_ROWS * 16;     // 512: one feature chunk of a stage = SBO
constexpr int WG_G_BYTES = WG_G_CHUNKS * WG_CHUNK_BYTES, WG_X_BYTES = WG_X_CHUNKS * WG_CHUNK_BYTES;
constexpr int WG_STAGE_BYTES = WG_G_BYTES + WG_X_BYTES;  // 76 800
constexpr int WG_STAGES = 2;
constexpr int WG_PAD_BYTES = 8 * WG_CHUNK_BYTES;       // the last M-side operand reads 8 chunks past the end of X
constexpr int WG_SM_BAR = WG_STAGES * WG_STAGE_BYTES + WG_PAD_BYTES;
constexpr int WG_SMEM = WG_SM_BAR + 64;
constexpr uint32_t WG_TM_COLS = 512;

struct WgradBlock {
  int a_is_x, a_col;   // M-side operand: matrix and first column (128 columns are read)
  int b_col;           // N-side operand: first column in the OTHER matrix
  int n;               // MMA N (multiple of 16)
  int tm_col;          // first TMEM column of the accumulator
  int m_valid;         // accumulator rows that belong to the layer
  int transposed;      // 1: rows = inputs, columns = outputs (dW[out][in] = D[in][out])
  int dst, dst_off;    // destination vector (0 base, 1 head, 2 sem) and offset of the layer's matrix in it
};

// column maps: G = [g_h1 0 | g_h2 128 | g_base 256 | g_hh1 272 | g_hh2 336 | g_out_h 400 | g_hs1 416 | g_hs2 480 | g_out_s 544]
//              X = [enc 0 | h1 64 | h2 192 | xh 320 | xs 352 | hh1 368 | hh2 432 | hs1 496 | hs2 560]
__constant__ WgradBlock WG_BLOCKS[9] = {
    {0, 0, 0, 64, 0, 128, 0, 0, 0},                         // W1  [128 x 64]  = g_h1^T  . enc
    {0, 128, 64, 128, 64, 128, 0, 0, 128 * 64},             // W2  [128 x 128] = g_h2^T  . h1
    {1, 192, 256, 16, 192, 128, 1, 0, 128 * 64 + 128 * 128},  // W3  [16 x 128]  = g_base^T . h2   (transposed)
    {0, 272, 320, 32, 208, 64, 0, 1, 0},                    // WH1 [64 x 32]   = g_hh1^T . xh
    {0, 336, 368, 64, 240, 64, 0, 1, 64 * 32},              // WH2 [64 x 64]   = g_hh2^T . hh1
    {1, 432, 400, 16, 304, 64, 1, 1, 64 * 32 + 64 * 64},    // WH3 [16 x 64]   = g_out_h^T . hh2 (transposed)
    {0, 416, 352, 16, 320, 64, 0, 2, 0},                    // WS1 [64 x 16]   = g_hs1^T . xs
    {0, 480, 496, 64, 336, 64, 0, 2, 64 * 16},              // WS2 [64 x 64]   = g_hs2^T . hs1
    {1, 560, 544, 32, 400, 64, 1, 2, 64 * 16 + 64 * 64},    // WS3 [<=32 x 64] = g_out_s^T . hs2 (transposed)
};

struct WgradIO {
  long long n;          // samples; both matrices are zero-padded to a multiple of WG_STAGE_ROWS rows
  const __half* G;      // [n_pad, 576]
  const __half* X;      // [n_pad, 624]
  float scale;          // 1 / loss_scale
  float* d_base;        // += [W1 | W2 | W3]
  float* d_head;        // += [WH1 | WH2 | WH3]
  float* d_sem;         // += [WS1 | WS2 | WS3]  (nullptr: no semantic head)
  int sem_out_rows;     // rows of WS3 in the flat parameter vector (tcnn pads the class count to 16)
};

// MN-major, K-major select in the instruction descriptor: bits 15 (A) and 16 (B)
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(uint32_t M, uint32_t N) {
  return ptx::make_idesc_f16(M, N) | (1u << 15) | (1u << 16);
}

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) field_wgrad_kernel(const WgradIO io) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_free = smem_base + WG_SM_BAR;          // [WG_STAGES]: the MMAs that read a stage have completed
  const uint32_t bar_done = bar_free + 8 * WG_STAGES;       // all MMAs of this CTA have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WG_SM_BAR + 8 * (WG_STAGES + 1));

  const long long n_stages = (io.n + WG_STAGE_ROWS - 1) / WG_STAGE_ROWS;
  // contiguous slice of stages for this CTA
  const long long per = n_stages / gridDim.x, rem = n_stages % gridDim.x;
  const long long s_begin = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem);
  const long long s_count = per + (blockIdx.x < rem ? 1 : 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) ptx::mbar_init(bar_free + 8 * i, 1);
    ptx::mbar_init(bar_done, 1);
    ptx::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < WG_PAD_BYTES / 16; i += WG_THREADS)  // the pad is read by the tensor core: keep it finite
    reinterpret_cast<uint4*>(smem + WG_STAGES * WG_STAGE_BYTES)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 4) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), WG_TM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // one stage: 72 + 78 chunks x 32 rows of 16 bytes.  Lane pairs copy the two halves of a 32-byte sector of one row.
  auto load_stage = [&](long long stage, int buf) {
    const long long row0 = stage * WG_STAGE_ROWS;
    const uint32_t dst0 = smem_base + buf * WG_STAGE_BYTES;
    for (int idx = threadIdx.x; idx < (WG_G_CHUNKS + WG_X_CHUNKS) * WG_STAGE_ROWS; idx += WG_THREADS) {
      const int c_lo = idx & 1, s = (idx >> 1) & (WG_STAGE_ROWS - 1), c = ((idx >> 6) << 1) | c_lo;  // c in [0, 150)
      if (c < WG_G_CHUNKS)
        cp_async16(dst0 + c * WG_CHUNK_BYTES + s * 16, io.G + (row0 + s) * WG_G_WIDTH + c * 8);
      else
        cp_async16(dst0 + WG_G_BYTES + (c - WG_G_CHUNKS) * WG_CHUNK_BYTES + s * 16,
                   io.X + (row0 + s) * WG_X_WIDTH + (c - WG_G_CHUNKS) * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (s_count > 0) load_stage(s_begin, 0);
  for (long long i = 0; i < s_count; ++i) {
    const int buf = (int)(i % WG_STAGES);
    if (i + 1 < s_count) {
      const int nb = (int)((i + 1) % WG_STAGES);
      if (i + 1 >= WG_STAGES) {  // the buffer was used by stage i + 1 - WG_STAGES: wait for its MMAs
        ptx::mbar_wait(bar_free + 8 * nb, (uint32_t)(((i + 1) / WG_STAGES - 1) & 1));
      }
      load_stage(s_begin + i + 1, nb);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    ptx::fence_proxy_async_smem();  // cp.async wrote through the generic proxy; the MMA reads through the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      ptx::tc_fence_after();
      const uint32_t g0 = smem_base + buf * WG_STAGE_BYTES, x0 = g0 + WG_G_BYTES;
#pragma unroll 1
      for (int b = 0; b < 9; ++b) {
        const WgradBlock& blk = WG_BLOCKS[b];
        if (blk.dst == 2 && io.d_sem == nullptr) continue;
        const uint32_t a0 = (blk.a_is_x ? x0 : g0) + (blk.a_col / 8) * WG_CHUNK_BYTES;
        const uint32_t b0 = (blk.a_is_x ? g0 : x0) + (blk.b_col / 8) * WG_CHUNK_BYTES;
        const uint32_t idesc = make_idesc_f16_mn(128, blk.n);
#pragma unroll
        for (int k = 0; k < WG_STAGE_ROWS / 16; ++k) {
          const uint64_t ad = ptx::make_smem_desc(a0 + k * 256, 128, WG_CHUNK_BYTES);
          const uint64_t bd = ptx::make_smem_desc(b0 + k * 256, 128, WG_CHUNK_BYTES);
          ptx::mma_f16_ss(tmem + blk.tm_col, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
      }
      ptx::mma_commit(bar_free + 8 * buf);
      if (i + 1 == s_count) ptx::mma_commit(bar_done);
    }
  }

  // ---- epilogue: TMEM -> scaled atomic adds into the flat fp32 gradients ----
  if (warp < 4 && s_count > 0) {
    ptx::mbar_wait(bar_done, 0);
    ptx::tc_fence_after();
    const int m = threadIdx.x;  // accumulator row = TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int b = 0; b < 9; ++b) {
      const WgradBlock& blk = WG_BLOCKS[b];
      float* dst = blk.dst == 0 ? io.d_base : (blk.dst == 1 ? io.d_head : io.d_sem);
      if (dst == nullptr) continue;
      dst += blk.dst_off;
      const int n_out = (b == 8) ? io.sem_out_rows : blk.n;  // WS3: only the flat vector's rows
#pragma unroll 1
      for (int c0 = 0; c0 < blk.n; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld_x16(trow + blk.tm_col + c0, v);  // warp-collective: every lane takes part
        ptx::tmem_wait_ld();
        if (m >= blk.m_valid) continue;
        if (!blk.transposed) {  // row m = output feature, columns = input features (contiguous in dW[out][in])
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 q = make_float4(__uint_as_float(v[j]) * io.scale, __uint_as_float(v[j + 1]) * io.scale,
                                   __uint_as_float(v[j + 2]) * io.scale, __uint_as_float(v[j + 3]) * io.scale);
            atomicAdd(reinterpret_cast<float4*>(dst + (size_t)m * blk.n + c0 + j), q);
          }
        } else {  // row m = input feature, column = output feature
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < n_out) atomicAdd(dst + (size_t)(c0 + j) * blk.m_valid + m, __uint_as_float(v[j]) * io.scale);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem, WG_TM_COLS);
}

}  // namespace apnerf
