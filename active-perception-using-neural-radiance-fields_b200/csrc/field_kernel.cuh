// The fused radiance-field kernel (kernels 2 + 3): hash-grid gather -> base MLP -> {colour head,
// semantic head}, 128 samples per tile, persistent CTAs (one per SM), warp-specialised:
//
//   warps 0-3   epilogue    thread r <-> sample row r <-> TMEM lane r: TMEM -> registers
//                           (tcgen05.ld), ReLU, fp16, next layer's A operand -> shared memory;
//                           SH-4 of the view direction; final activations and output.
//   warp  4     MMA issuer  one thread issues every tcgen05.mma / tcgen05.commit; owns TMEM.
//   warps 5-12  encoders    256 threads: 8-byte hash-table gathers (L2-resident table),
//                           trilinear blend, fp16 features straight into the MMA's A tile.
//
// The 64-wide encoding and all activations stay in shared memory / TMEM; only positions come
// in and (density, rgb, semantic logits) go out.  All nine weight matrices (80 KB fp16) sit in
// shared memory for the lifetime of the CTA in the UMMA K-major no-swizzle layout.
//
// Numerics (the specification the oracle restates): fp16 operands, fp32 accumulation in TMEM,
// every layer output rounded to fp16 (tcnn FullyFusedMLP stores fp16 activations), ReLU on
// hidden layers, inputs of the head / semantic networks padded to a multiple of 16 with 1.0.
#pragma once
#include "field.cuh"
#include "ptx.cuh"

namespace apnerf {

constexpr int N_EPI_WARPS = 4, N_ENC_WARPS = 8;
constexpr int FIELD_THREADS = (N_EPI_WARPS + 1 + N_ENC_WARPS) * 32;  // 416
constexpr int N_ENC_THREADS = N_ENC_WARPS * 32;                      // 256
constexpr int A0_STAGES = 2;

// shared-memory map (bytes)
constexpr int SM_W = 0;
constexpr int SM_A0 = SM_W + W_BYTES;                             // A0_STAGES x [128 x 64] fp16
constexpr int SM_H = SM_A0 + A0_STAGES * TILE_M * ENC_DIM * 2;    // [128 x 128]
constexpr int SM_XH = SM_H + TILE_M * HID * 2;                    // [128 x 32]
constexpr int SM_XS = SM_XH + TILE_M * HEAD_IN * 2;               // [128 x 16]
constexpr int SM_HH = SM_XS + TILE_M * SEM_IN * 2;                // [128 x 64]
constexpr int SM_HS = SM_HH + TILE_M * HID2 * 2;                  // [128 x 64]
constexpr int SM_BAR = SM_HS + TILE_M * HID2 * 2;                 // mbarriers + tmem base
constexpr int FIELD_SMEM = SM_BAR + 128;

// TMEM column map (fp32 accumulators, 128 lanes)
constexpr uint32_t TM_MAIN = 0;    // 128 cols: base layer 1 / 2 outputs
constexpr uint32_t TM_OUT3 = 128;  // 16 cols : base output (density, geo features)
constexpr uint32_t TM_H = 160;     // 64 cols : head hidden
constexpr uint32_t TM_S = 224;     // 64 cols : semantic hidden
constexpr uint32_t TM_HO = 288;    // 16 cols : rgb (padded)
constexpr uint32_t TM_SO = 320;    // 32 cols : semantic logits (padded)
constexpr uint32_t TM_COLS = 512;

struct FieldIO {
  // --- inputs: either explicit points (positions/directions) or ray samples ---
  long long n;                  // number of samples (used when n_dev == nullptr)
  const int* n_dev;             // optional device-side sample count (fused renderer)
  const float* positions;       // [n, 3] or nullptr
  const float* directions;      // [n, 3] or nullptr
  const int* ray_idx;           // [n]   (ray-sample mode)
  const float* t_starts;        // [n]
  const float* t_ends;          // [n]
  const float* rays_o;          // [n_rays, 3]
  const float* rays_d;          // [n_rays, 3]
  // --- parameters ---
  const uint2* table;           // fp16 [entries, 4]
  const uint4* weights;         // W_BYTES blob in UMMA layout
  // --- outputs (element (row, ch) at base[ch * ch_stride + row * row_stride]) ---
  float* density;               // [n]
  float* rgb;
  long long rgb_row, rgb_ch;
  float* sem;
  long long sem_row, sem_ch;
  __half* feat;                 // optional [n, 15] geo features (query_density(return_feat=True))
  int n_sem;                    // number of semantic classes actually written (<= 32), 0 = none
  int density_only;             // stop after the base MLP
};

__device__ __forceinline__ uint32_t pack_relu_h2(uint32_t a_bits, uint32_t b_bits) {
  const float a = fmaxf(__uint_as_float(a_bits), 0.f), b = fmaxf(__uint_as_float(b_bits), 0.f);
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// TMEM accumulator columns [col0, col0 + 32) of this thread's row -> ReLU -> fp16 -> the A tile
// of the next layer (K-chunks col0/8 .. col0/8+3); `rows16` = TILE_M * 16 bytes per K-chunk.
__device__ __forceinline__ void relu_store_32(uint32_t taddr, uint8_t* dst, int row, int col0) {
  uint32_t v[32];
  ptx::tmem_ld_x32(taddr + col0, v);
  ptx::tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q;
    q.x = pack_relu_h2(v[8 * j + 0], v[8 * j + 1]);
    q.y = pack_relu_h2(v[8 * j + 2], v[8 * j + 3]);
    q.z = pack_relu_h2(v[8 * j + 4], v[8 * j + 5]);
    q.w = pack_relu_h2(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(dst + (col0 / 8 + j) * (TILE_M * 16) + row * 16) = q;
  }
}

// Issue one layer: D[128 x N] (+)= A[128 x K] * W[N x K]^T as K/16 tcgen05.mma instructions.
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, uint32_t a_smem, uint32_t w_smem, int N, int K) {
  const uint32_t idesc = ptx::make_idesc_f16(TILE_M, N);
  const uint32_t a_lbo = TILE_M * 16, b_lbo = N * 16;
  for (int k = 0; k < K / 16; ++k) {
    const uint64_t ad = ptx::make_smem_desc(a_smem + k * 2 * a_lbo, a_lbo, 128);
    const uint64_t bd = ptx::make_smem_desc(w_smem + k * 2 * b_lbo, b_lbo, 128);
    ptx::mma_f16_ss(d_tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
  }
}

__device__ __forceinline__ void sample_point(const FieldIO& io, long long s, float p[3], float d[3], bool want_dir) {
  if (io.positions) {
    p[0] = io.positions[3 * s], p[1] = io.positions[3 * s + 1], p[2] = io.positions[3 * s + 2];
    if (want_dir) d[0] = io.directions[3 * s], d[1] = io.directions[3 * s + 1], d[2] = io.directions[3 * s + 2];
  } else {
    // positions = o + d * (t_start + t_end) / 2, op for op as perception/models/utils.py:833-836
    const long long r = io.ray_idx[s];
    const float tsum = __fadd_rn(io.t_starts[s], io.t_ends[s]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float da = io.rays_d[3 * r + a];
      p[a] = __fadd_rn(io.rays_o[3 * r + a], __fmul_rn(__fmul_rn(da, tsum), 0.5f));
      d[a] = da;
    }
  }
}

__global__ void __launch_bounds__(FIELD_THREADS, 1)
field_forward_kernel(const FieldIO io, const HashGridMeta meta, const FieldConst fc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = ptx::smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  // bars[0..1] a0_full, bars[2..3] a0_empty, bars[4] mma_done, bars[5] epi_done, then tmem base
  const uint32_t bar_full = smem_base + SM_BAR, bar_empty = bar_full + 16, bar_mma = bar_full + 32,
                 bar_epi = bar_full + 40;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 64);

  const long long n = io.n_dev ? (long long)*io.n_dev : io.n;
  const long long n_tiles = (n + TILE_M - 1) / TILE_M;

  // ---- one-time setup: weights -> smem, barriers, TMEM ----
  for (int i = threadIdx.x; i < W_BYTES / 16; i += FIELD_THREADS)
    reinterpret_cast<uint4*>(smem + SM_W)[i] = __ldg(io.weights + i);
  if (threadIdx.x == 0) {
    for (int i = 0; i < A0_STAGES; ++i) {
      ptx::mbar_init(bar_full + 8 * i, N_ENC_THREADS);
      ptx::mbar_init(bar_empty + 8 * i, 1);
    }
    ptx::mbar_init(bar_mma, 1);
    ptx::mbar_init(bar_epi, N_EPI_WARPS * 32);
    ptx::fence_barrier_init();
  }
  if (warp == N_EPI_WARPS) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), TM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();  // weights were written through the generic proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= N_EPI_WARPS + 1) {
    // =============================== encoders ===============================
    const int e = threadIdx.x - (N_EPI_WARPS + 1) * 32;
    const int row = e & (TILE_M - 1), half = e >> 7;  // this thread does levels [8*half, 8*half+8)
    const float ext[3] = {__fsub_rn(fc.aabb[3], fc.aabb[0]), __fsub_rn(fc.aabb[4], fc.aabb[1]),
                          __fsub_rn(fc.aabb[5], fc.aabb[2])};
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it % A0_STAGES;
      const uint32_t ph = (it / A0_STAGES) & 1;
      const long long s = tile * TILE_M + row;
      float x[3] = {0.5f, 0.5f, 0.5f};
      const bool valid = s < n;
      if (valid) {
        float p[3], d[3];
        sample_point(io, s, p, d, false);
#pragma unroll
        for (int a = 0; a < 3; ++a) x[a] = __fdiv_rn(__fsub_rn(p[a], fc.aabb[a]), ext[a]);
      }
      uint4 q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int l = half * 8 + 2 * j;
        uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
        if (valid) {
          if (l < meta.n_levels) lo = encode_level(meta, l, x, io.table);
          if (l + 1 < meta.n_levels) hi = encode_level(meta, l + 1, x, io.table);
        }
        q[j] = make_uint4(lo.x, lo.y, hi.x, hi.y);
      }
      ptx::mbar_wait(bar_empty + 8 * buf, ph ^ 1);  // MMA of the tile that used this slot is done
      uint8_t* a0 = smem + SM_A0 + buf * (TILE_M * ENC_DIM * 2);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(a0 + (half * 4 + j) * (TILE_M * 16) + row * 16) = q[j];
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(bar_full + 8 * buf);
    }
  } else if (warp == N_EPI_WARPS) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      uint32_t epi_ph = 0;
      int it = 0;
      const uint32_t sW = smem_base + SM_W;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it % A0_STAGES;
        const uint32_t ph = (it / A0_STAGES) & 1;
        // base layer 1: enc[128x64] x W1^T -> TM_MAIN
        ptx::mbar_wait(bar_full + 8 * buf, ph);
        ptx::tc_fence_after();
        issue_layer(tmem + TM_MAIN, smem_base + SM_A0 + buf * (TILE_M * ENC_DIM * 2), sW + W1_OFF, HID, ENC_DIM);
        ptx::mma_commit(bar_empty + 8 * buf);
        ptx::mma_commit(bar_mma);
        // base layer 2
        ptx::mbar_wait(bar_epi, epi_ph), epi_ph ^= 1;
        ptx::tc_fence_after();
        issue_layer(tmem + TM_MAIN, smem_base + SM_H, sW + W2_OFF, HID, HID);
        ptx::mma_commit(bar_mma);
        // base output
        ptx::mbar_wait(bar_epi, epi_ph), epi_ph ^= 1;
        ptx::tc_fence_after();
        issue_layer(tmem + TM_OUT3, smem_base + SM_H, sW + W3_OFF, BASE_OUT, HID);
        ptx::mma_commit(bar_mma);
        if (!io.density_only) {
          // head / semantic layer 1
          ptx::mbar_wait(bar_epi, epi_ph), epi_ph ^= 1;
          ptx::tc_fence_after();
          issue_layer(tmem + TM_H, smem_base + SM_XH, sW + WH1_OFF, HID2, HEAD_IN);
          issue_layer(tmem + TM_S, smem_base + SM_XS, sW + WS1_OFF, HID2, SEM_IN);
          ptx::mma_commit(bar_mma);
          // layer 2
          ptx::mbar_wait(bar_epi, epi_ph), epi_ph ^= 1;
          ptx::tc_fence_after();
          issue_layer(tmem + TM_H, smem_base + SM_HH, sW + WH2_OFF, HID2, HID2);
          issue_layer(tmem + TM_S, smem_base + SM_HS, sW + WS2_OFF, HID2, HID2);
          ptx::mma_commit(bar_mma);
          // outputs
          ptx::mbar_wait(bar_epi, epi_ph), epi_ph ^= 1;
          ptx::tc_fence_after();
          issue_layer(tmem + TM_HO, smem_base + SM_HH, sW + WH3_OFF, HEAD_OUT, HID2);
          issue_layer(tmem + TM_SO, smem_base + SM_HS, sW + WS3_OFF, SEM_OUT, HID2);
          ptx::mma_commit(bar_mma);
        }
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int row = threadIdx.x;  // 0..127 == TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t mma_ph = 0;
    const float ext[3] = {__fsub_rn(fc.aabb[3], fc.aabb[0]), __fsub_rn(fc.aabb[4], fc.aabb[1]),
                          __fsub_rn(fc.aabb[5], fc.aabb[2])};
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long s = tile * TILE_M + row;
      const bool valid = s < n;
      // ---- base layer 1 -> H
      ptx::mbar_wait(bar_mma, mma_ph), mma_ph ^= 1;
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < HID; c += 32) relu_store_32(trow + TM_MAIN, smem + SM_H, row, c);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_epi);
      // ---- base layer 2 -> H
      ptx::mbar_wait(bar_mma, mma_ph), mma_ph ^= 1;
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < HID; c += 32) relu_store_32(trow + TM_MAIN, smem + SM_H, row, c);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_epi);
      // ---- base output: density + geo features; build the head / semantic inputs
      ptx::mbar_wait(bar_mma, mma_ph), mma_ph ^= 1;
      ptx::tc_fence_after();
      uint32_t o3[16];
      ptx::tmem_ld_x16(trow + TM_OUT3, o3);
      ptx::tmem_wait_ld();
      __half hb[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) hb[i] = __float2half_rn(__uint_as_float(o3[i]));
      float p[3], d[3] = {0.f, 0.f, 1.f};
      bool inside = false;
      if (valid) {
        sample_point(io, s, p, d, !io.density_only);
        inside = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float xa = __fdiv_rn(__fsub_rn(p[a], fc.aabb[a]), ext[a]);
          inside = inside && (xa > 0.0f) && (xa < 1.0f);
        }
        // density = exp(x - 1) * selector  (ngp.py:79,191-193; fp16 network output upcast first)
        const float dens = inside ? expf(__fsub_rn(__half2float(hb[0]), 1.0f)) : 0.0f;
        io.density[s] = dens;
        if (io.feat) {
#pragma unroll
          for (int i = 0; i < 15; ++i) io.feat[s * 15 + i] = hb[1 + i];
        }
      }
      if (io.density_only) continue;
      {
        float sh[16];
        sh4(d, sh);
        __align__(16) __half hx[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) hx[i] = __float2half_rn(sh[i]);
#pragma unroll
        for (int i = 0; i < 15; ++i) hx[16 + i] = hb[1 + i];
        hx[31] = __float2half_rn(1.0f);
        const uint4* hq = reinterpret_cast<const uint4*>(hx);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(smem + SM_XH + j * (TILE_M * 16) + row * 16) = hq[j];
#pragma unroll
        for (int j = 0; j < 2; ++j)
          *reinterpret_cast<uint4*>(smem + SM_XS + j * (TILE_M * 16) + row * 16) = hq[2 + j];
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_epi);
      // ---- head / semantic hidden layers 1 and 2
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        ptx::mbar_wait(bar_mma, mma_ph), mma_ph ^= 1;
        ptx::tc_fence_after();
        relu_store_32(trow + TM_H, smem + SM_HH, row, 0);
        relu_store_32(trow + TM_H, smem + SM_HH, row, 32);
        relu_store_32(trow + TM_S, smem + SM_HS, row, 0);
        relu_store_32(trow + TM_S, smem + SM_HS, row, 32);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_epi);
      }
      // ---- outputs: rgb = sigmoid(head), semantic logits raw (ngp.py:210-220)
      ptx::mbar_wait(bar_mma, mma_ph), mma_ph ^= 1;
      ptx::tc_fence_after();
      uint32_t oh[16], os[32];
      ptx::tmem_ld_x16(trow + TM_HO, oh);
      ptx::tmem_ld_x32(trow + TM_SO, os);
      ptx::tmem_wait_ld();
      ptx::tc_fence_before();
      if (valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = __half2float(__float2half_rn(__uint_as_float(oh[c])));
          io.rgb[c * io.rgb_ch + s * io.rgb_row] = 1.0f / (1.0f + expf(-v));
        }
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < io.n_sem)
            io.sem[c * io.sem_ch + s * io.sem_row] = __half2float(__float2half_rn(__uint_as_float(os[c])));
      }
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == N_EPI_WARPS) ptx::tmem_dealloc(tmem, TM_COLS);
}

}  // namespace apnerf
