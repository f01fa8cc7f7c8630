// Backward of the three fused MLPs (training side of kernels 2 + 3): for every tile of 128 samples the
// chain of "dX = dY . W" products runs on the tensor cores with the activation gradients staying in shared
// memory / TMEM, exactly like the forward kernel runs its layers:
//
//   d rgb logits [128 x 16] . WH3 -> (x relu'(hh2)) . WH2 -> (x relu'(hh1)) . WH1 -> d head input [128 x 32]
//   d sem logits [128 x 32] . WS3 -> (x relu'(hs2)) . WS2 -> (x relu'(hs1)) . WS1 -> d sem input  [128 x 16]
//   d base out = [d density logit | d feat(head) + d feat(sem)] [128 x 16]
//                . W3 -> (x relu'(h2)) . W2 -> (x relu'(h1)) . W1 -> d encoding [128 x 64]  (-> hash-grid backward)
//
// Replaces the backward of tcnn's FullyFusedMLP behind the reference's loss.backward()
// (scripts/pipeline.py:518; modules perception/models/radiance_fields/ngp.py:123-169).  Like tcnn, the
// gradients are multiplied by a loss scale so that they survive fp16 (operands fp16, accumulation fp32 in
// TMEM), the activation gradients of every layer are written out in fp16, and the weight gradients
// dW = dY^T . X are plain GEMMs over all samples done by the caller (library GEMM, as tcnn does with CUTLASS).
//
//   warps 0-3  thread r <-> sample row r <-> TMEM lane r: build the A operand of the next product from the
//              accumulators (ReLU mask from the saved forward activation, fp16), store it to global too
//   warp  4    one thread issues the tcgen05.mma chain; the warp owns the TMEM allocation
//
// Weights: the forward blob's matrices TRANSPOSED ([in, out], K-major over `out`), same offsets.
#pragma once
#include "field_kernel.cuh"  // issue_layer, the shared geometry

namespace apnerf {

constexpr int BWD_THREADS = 160;
constexpr int BW_SM_W = 0;
constexpr int BW_SM_ACT = BW_SM_W + W_BYTES;          // 32 KB: A operand of the current product(s)
constexpr int BW_SM_BAR = BW_SM_ACT + TILE_M * HID * 2;
constexpr int BWD_SMEM = BW_SM_BAR + 64;
constexpr uint32_t BW_TM_COLS = 128;

struct FieldBwdIO {
  long long n;
  // incoming gradients (fp32, unscaled) w.r.t. the raw network outputs
  const float* d_dens;   // [n]      density logit
  const float* d_rgb;    // [n, 3]   rgb logits
  const float* d_sem;    // [n, n_sem] semantic logits (may be nullptr)
  int n_sem;
  // forward activations saved by the forward kernel (fp16, rows `act_stride` elements apart)
  long long act_stride;
  const __half* h1;      // [n, 128]
  const __half* h2;      // [n, 128]
  const __half* hh1;     // [n, 64]
  const __half* hh2;     // [n, 64]
  const __half* hs1;     // [n, 64]
  const __half* hs2;     // [n, 64]
  const uint4* weights_t;  // transposed blob
  float loss_scale;
  // outputs: activation gradients x loss_scale in fp16 (inputs of the weight-gradient GEMMs; rows `g_stride`
  // elements apart) ...
  long long g_stride;
  __half* g_out_h;       // [n, 16]  incoming rgb-logit gradient, padded
  __half* g_out_s;       // [n, 32]  incoming semantic-logit gradient, padded
  __half* g_hh2;         // [n, 64]
  __half* g_hs2;         // [n, 64]
  __half* g_hh1;         // [n, 64]
  __half* g_hs1;         // [n, 64]
  __half* g_base;        // [n, 16]
  __half* g_h2;          // [n, 128]
  __half* g_h1;          // [n, 128]
  // ... and the gradient w.r.t. the encoding, unscaled fp32
  float* d_enc;          // [n, 64]
};

// accumulator columns [col0, col0 + 32) of this thread's row -> x (saved activation > 0) -> fp16 ->
// K-chunks col0/8 .. col0/8+3 of the next A operand (+ the same 64 bytes to global when g != nullptr)
__device__ __forceinline__ void mask_store_32(uint32_t taddr, uint8_t* dst, int row, int col0, const __half* act_row,
                                              __half* g_row) {
  uint32_t v[32];
  ptx::tmem_ld_x32(taddr + col0, v);
  uint4 m[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
  if (act_row) {
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = __ldg(reinterpret_cast<const uint4*>(act_row + col0) + j);
  }
  ptx::tmem_wait_ld();
  const __half* mh = reinterpret_cast<const __half*>(m);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __align__(16) __half2 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 8 * j + 2 * e;
      const float a = (act_row && !(__half2float(mh[c]) > 0.f)) ? 0.f : __uint_as_float(v[c]);
      const float b = (act_row && !(__half2float(mh[c + 1]) > 0.f)) ? 0.f : __uint_as_float(v[c + 1]);
      h[e] = __floats2half2_rn(a, b);
    }
    const uint4 q = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(dst + (col0 / 8 + j) * (TILE_M * 16) + row * 16) = q;
    if (g_row) *reinterpret_cast<uint4*>(g_row + col0 + 8 * j) = q;
  }
}

__global__ void __launch_bounds__(BWD_THREADS, 1) field_backward_kernel(const FieldBwdIO io) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_mma = smem_base + BW_SM_BAR, bar_epi = bar_mma + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BW_SM_BAR + 16);
  const long long n = io.n;
  const long long n_tiles = (n + TILE_M - 1) / TILE_M;

  for (int i = threadIdx.x; i < W_BYTES / 16; i += BWD_THREADS)
    reinterpret_cast<uint4*>(smem + BW_SM_W)[i] = __ldg(io.weights_t + i);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_mma, 1);
    ptx::mbar_init(bar_epi, 128);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), BW_TM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sW = smem_base + BW_SM_W, act_s = smem_base + BW_SM_ACT;

  if (warp == 4) {
    // ===================== MMA issuer: six products per tile, each after the rows are in place =====================
    if (lane == 0) {
      uint32_t ph = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        auto wait_rows = [&]() {
          ptx::mbar_wait(bar_epi, ph);
          ph ^= 1;
          ptx::tc_fence_after();
        };
        wait_rows();  // d rgb logits at +0 [128 x 16], d sem logits at +4096 [128 x 32]
        issue_layer(tmem + 0, act_s, sW + WH3_OFF, HID2, HEAD_OUT);
        issue_layer(tmem + 64, act_s + TILE_M * HEAD_OUT * 2, sW + WS3_OFF, HID2, SEM_OUT);
        ptx::mma_commit(bar_mma);
        wait_rows();  // g_hh2 at +0, g_hs2 at +16384 (each [128 x 64])
        issue_layer(tmem + 0, act_s, sW + WH2_OFF, HID2, HID2);
        issue_layer(tmem + 64, act_s + TILE_M * HID2 * 2, sW + WS2_OFF, HID2, HID2);
        ptx::mma_commit(bar_mma);
        wait_rows();  // g_hh1, g_hs1
        issue_layer(tmem + 0, act_s, sW + WH1_OFF, HEAD_IN, HID2);
        issue_layer(tmem + 32, act_s + TILE_M * HID2 * 2, sW + WS1_OFF, SEM_IN, HID2);
        ptx::mma_commit(bar_mma);
        wait_rows();  // g_base [128 x 16]
        issue_layer(tmem + 0, act_s, sW + W3_OFF, HID, BASE_OUT);
        ptx::mma_commit(bar_mma);
        wait_rows();  // g_h2 [128 x 128]
        issue_layer(tmem + 0, act_s, sW + W2_OFF, HID, HID);
        ptx::mma_commit(bar_mma);
        wait_rows();  // g_h1 [128 x 128]
        issue_layer(tmem + 0, act_s, sW + W1_OFF, ENC_DIM, HID);
        ptx::mma_commit(bar_mma);
        // (the next tile's first wait_rows also orders this tile's d_enc reads before the accumulators are reused)
      }
    }
  } else {
    // ===================== row threads =====================
    const int row = threadIdx.x;  // 0..127 == TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint8_t* act = smem + BW_SM_ACT;
    uint32_t ph = 0;
    auto rows_ready = [&]() {
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_epi);
    };
    auto wait_mma = [&]() {
      ptx::mbar_wait(bar_mma, ph);
      ph ^= 1;
      ptx::tc_fence_after();
    };
    const float scale = io.loss_scale, inv_scale = 1.0f / io.loss_scale;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long s = tile * TILE_M + row;
      const bool valid = s < n;
      // ---- incoming gradients -> fp16 A operands
      {
        __align__(16) __half g[HEAD_OUT + SEM_OUT];
#pragma unroll
        for (int c = 0; c < HEAD_OUT + SEM_OUT; ++c) g[c] = __float2half_rn(0.f);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) g[c] = __float2half_rn(io.d_rgb[s * 3 + c] * scale);
          if (io.d_sem)
            for (int c = 0; c < io.n_sem; ++c) g[HEAD_OUT + c] = __float2half_rn(io.d_sem[s * io.n_sem + c] * scale);
        }
        const uint4* q = reinterpret_cast<const uint4*>(g);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          *reinterpret_cast<uint4*>(act + j * (TILE_M * 16) + row * 16) = q[j];
          if (valid) reinterpret_cast<uint4*>(io.g_out_h + s * io.g_stride)[j] = q[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(act + TILE_M * HEAD_OUT * 2 + j * (TILE_M * 16) + row * 16) = q[2 + j];
          if (valid) reinterpret_cast<uint4*>(io.g_out_s + s * io.g_stride)[j] = q[2 + j];
        }
      }
      rows_ready();
      // ---- hidden layers 2 and 1 of the head / semantic networks
#pragma unroll 1
      for (int layer = 2; layer >= 1; --layer) {
        const __half* ah = valid ? (layer == 2 ? io.hh2 : io.hh1) + s * io.act_stride : nullptr;
        const __half* as = valid ? (layer == 2 ? io.hs2 : io.hs1) + s * io.act_stride : nullptr;
        __half* gh = valid ? (layer == 2 ? io.g_hh2 : io.g_hh1) + s * io.g_stride : nullptr;
        __half* gs = valid ? (layer == 2 ? io.g_hs2 : io.g_hs1) + s * io.g_stride : nullptr;
        wait_mma();
        // invalid rows: mask pointer nullptr would pass the values through; their inputs were zero, so they stay zero
        mask_store_32(trow + 0, act, row, 0, ah, gh);
        mask_store_32(trow + 0, act, row, 32, ah, gh);
        mask_store_32(trow + 64, act + TILE_M * HID2 * 2, row, 0, as, gs);
        mask_store_32(trow + 64, act + TILE_M * HID2 * 2, row, 32, as, gs);
        rows_ready();
      }
      // ---- d base output = [d density logit | d feat from the head input (cols 16..30) + from the semantic input (0..14)]
      wait_mma();
      {
        uint32_t xh[16], xs[16];
        ptx::tmem_ld_x16(trow + 16, xh);
        ptx::tmem_ld_x16(trow + 32, xs);
        ptx::tmem_wait_ld();
        __align__(16) __half g[BASE_OUT];
        g[0] = __float2half_rn(valid ? io.d_dens[s] * scale : 0.f);
#pragma unroll
        for (int i = 0; i < 15; ++i) g[1 + i] = __float2half_rn(__uint_as_float(xh[i]) + __uint_as_float(xs[i]));
        const uint4* q = reinterpret_cast<const uint4*>(g);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          *reinterpret_cast<uint4*>(act + j * (TILE_M * 16) + row * 16) = q[j];
          if (valid) *reinterpret_cast<uint4*>(io.g_base + s * io.g_stride + 8 * j) = q[j];
        }
      }
      rows_ready();
      // ---- base hidden layers 2 and 1
#pragma unroll 1
      for (int layer = 2; layer >= 1; --layer) {
        const __half* a = valid ? (layer == 2 ? io.h2 : io.h1) + s * io.act_stride : nullptr;
        __half* g = valid ? (layer == 2 ? io.g_h2 : io.g_h1) + s * io.g_stride : nullptr;
        wait_mma();
#pragma unroll 1
        for (int c = 0; c < HID; c += 32) mask_store_32(trow + 0, act, row, c, a, g);
        rows_ready();
      }
      // ---- d encoding, unscaled fp32
      wait_mma();
#pragma unroll 1
      for (int c = 0; c < ENC_DIM; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_x32(trow + c, v);
        ptx::tmem_wait_ld();
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(io.d_enc + s * ENC_DIM + c);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]) * inv_scale, __uint_as_float(v[4 * j + 1]) * inv_scale,
                                 __uint_as_float(v[4 * j + 2]) * inv_scale, __uint_as_float(v[4 * j + 3]) * inv_scale);
        }
      }
      // no arrival here: every row thread passes the wait above before its next arrival, so an mbarrier phase
      // never sees two arrivals of one thread
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem, BW_TM_COLS);
}

}  // namespace apnerf
