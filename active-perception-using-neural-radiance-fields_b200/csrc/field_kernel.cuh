// The fused radiance-field kernel (kernels 2 + 3): hash-grid gather -> base MLP -> {colour head,
// semantic head}, 128 samples per tile, persistent CTAs (one per SM), warp-specialised:
//
//   warps 0-7   epilogue    thread r <-> sample row r <-> TMEM lane r: TMEM -> registers
//                           (tcgen05.ld), ReLU, fp16, next layer's A operand -> shared memory;
//                           SH-4 of the view direction; final activations and output.
//   warps 8-9   MMA issuers one thread per chain issues tcgen05.mma / tcgen05.commit; warp 8 owns TMEM.
//   warps 10-25 encoders    512 threads (4 levels of one sample each): 8-byte hash-table gathers (L2-resident table),
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

// Two MLP "chains" ping-pong over alternate tiles: while chain 0's epilogue warps convert one
// layer's accumulators, chain 1's MMAs run (ncu on the single-chain version: encoders and the
// serial MMA -> epilogue chain were both ~90 % busy at ~14.7 k cycles per tile).
constexpr int N_CHAINS = 2;
constexpr int N_EPI_WARPS = 4 * N_CHAINS, N_MMA_WARPS = N_CHAINS, N_ENC_WARPS = 16;
constexpr int FIELD_THREADS = (N_EPI_WARPS + N_MMA_WARPS + N_ENC_WARPS) * 32;  // 832
constexpr int N_ENC_THREADS = N_ENC_WARPS * 32;                                // 512
constexpr int LEVELS_PER_ENC_THREAD = MAX_LEVELS * TILE_M / N_ENC_THREADS;     // 4
constexpr int A0_STAGES = 3;

// shared-memory map (bytes).  Per chain ONE 32 KB activation region is reused by every layer:
//   H  [128 x 128] at +0                      (base layers 1, 2 outputs)
//   XH [128 x 32]  at +0,  XS [128 x 16] at +8192    (head / semantic inputs; H is dead by then)
//   HH [128 x 64]  at +0,  HS [128 x 64] at +16384   (head / semantic hidden; XH / XS are dead)
constexpr int SM_W = 0;
constexpr int SM_A0 = SM_W + W_BYTES;                             // A0_STAGES x [128 x 64] fp16
constexpr int SM_ACT = SM_A0 + A0_STAGES * TILE_M * ENC_DIM * 2;  // N_CHAINS x 32 KB
constexpr int ACT_BYTES = TILE_M * HID * 2;
constexpr int ACT_XH = 0, ACT_XS = TILE_M * HEAD_IN * 2, ACT_HH = 0, ACT_HS = TILE_M * HID2 * 2;
constexpr int SM_BAR = SM_ACT + N_CHAINS * ACT_BYTES;             // mbarriers + tmem base
constexpr int FIELD_SMEM = SM_BAR + 128;

// TMEM column map per chain (fp32 accumulators, 128 lanes, 128 columns, reused layer by layer)
constexpr uint32_t TM_CHAIN = 128;
constexpr uint32_t TM_MAIN = 0;   // 128 cols: base layer 1 / 2 outputs
constexpr uint32_t TM_OUT3 = 0;   // 16 cols : base output (density, geo features)
constexpr uint32_t TM_H = 0;      // 64 cols : head hidden
constexpr uint32_t TM_S = 64;     // 64 cols : semantic hidden
constexpr uint32_t TM_HO = 0;     // 16 cols : rgb (padded)
constexpr uint32_t TM_SO = 16;    // 32 cols : semantic logits (padded)
constexpr uint32_t TM_COLS = 256;

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
  uint4* packed;                // optional [n][5] x 16 B rows of 40 fp16: raw network outputs for the fused
                                // renderer {density logit (-inf outside the aabb), rgb logits x3, sigma fp32, pad, 32 sem logits}
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
  // barriers: a0_full[3] | a0_empty[3] | mma_done[2] | epi_done[2] | tmem base slot
  const uint32_t bar_full = smem_base + SM_BAR, bar_empty = bar_full + 8 * A0_STAGES,
                 bar_mma = bar_empty + 8 * A0_STAGES, bar_epi = bar_mma + 8 * N_CHAINS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * (2 * A0_STAGES + 2 * N_CHAINS));

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
    for (int c = 0; c < N_CHAINS; ++c) {
      ptx::mbar_init(bar_mma + 8 * c, 1);
      ptx::mbar_init(bar_epi + 8 * c, 128);
    }
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
  const float ext[3] = {__fsub_rn(fc.aabb[3], fc.aabb[0]), __fsub_rn(fc.aabb[4], fc.aabb[1]),
                        __fsub_rn(fc.aabb[5], fc.aabb[2])};

  if (warp >= N_EPI_WARPS + N_MMA_WARPS) {
    // =============================== encoders ===============================
    const int e = threadIdx.x - (N_EPI_WARPS + N_MMA_WARPS) * 32;
    const int row = e & (TILE_M - 1), part = e >> 7;  // this thread does levels [4*part, 4*part+4)
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
      uint4 q[LEVELS_PER_ENC_THREAD / 2];
#pragma unroll
      for (int j = 0; j < LEVELS_PER_ENC_THREAD / 2; ++j) {
        const int l = part * LEVELS_PER_ENC_THREAD + 2 * j;
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
      for (int j = 0; j < LEVELS_PER_ENC_THREAD / 2; ++j)
        *reinterpret_cast<uint4*>(a0 + (part * (LEVELS_PER_ENC_THREAD / 2) + j) * (TILE_M * 16) + row * 16) = q[j];
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(bar_full + 8 * buf);
    }
  } else if (warp >= N_EPI_WARPS) {
    // =============================== MMA issuers (one per chain) ===============================
    const int chain = warp - N_EPI_WARPS;
    if (lane == 0) {
      uint32_t epi_ph = 0;
      bool first = true;
      const uint32_t sW = smem_base + SM_W;
      const uint32_t act = smem_base + SM_ACT + chain * ACT_BYTES;
      const uint32_t tm = tmem + chain * TM_CHAIN;
      const uint32_t my_mma = bar_mma + 8 * chain, my_epi = bar_epi + 8 * chain;
      int it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        if ((it % N_CHAINS) != chain) continue;
        const int buf = it % A0_STAGES;
        const uint32_t ph = (it / A0_STAGES) & 1;
        // base layer 1: enc[128x64] x W1^T -> TM_MAIN (after the previous tile's outputs were read)
        ptx::mbar_wait(bar_full + 8 * buf, ph);
        if (!first) ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
        first = false;
        ptx::tc_fence_after();
        issue_layer(tm + TM_MAIN, smem_base + SM_A0 + buf * (TILE_M * ENC_DIM * 2), sW + W1_OFF, HID, ENC_DIM);
        ptx::mma_commit(bar_empty + 8 * buf);
        ptx::mma_commit(my_mma);
        // base layer 2
        ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
        ptx::tc_fence_after();
        issue_layer(tm + TM_MAIN, act, sW + W2_OFF, HID, HID);
        ptx::mma_commit(my_mma);
        // base output
        ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
        ptx::tc_fence_after();
        issue_layer(tm + TM_OUT3, act, sW + W3_OFF, BASE_OUT, HID);
        ptx::mma_commit(my_mma);
        if (!io.density_only) {
          // head / semantic layer 1
          ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
          ptx::tc_fence_after();
          issue_layer(tm + TM_H, act + ACT_XH, sW + WH1_OFF, HID2, HEAD_IN);
          issue_layer(tm + TM_S, act + ACT_XS, sW + WS1_OFF, HID2, SEM_IN);
          ptx::mma_commit(my_mma);
          // layer 2
          ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
          ptx::tc_fence_after();
          issue_layer(tm + TM_H, act + ACT_HH, sW + WH2_OFF, HID2, HID2);
          issue_layer(tm + TM_S, act + ACT_HS, sW + WS2_OFF, HID2, HID2);
          ptx::mma_commit(my_mma);
          // outputs
          ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
          ptx::tc_fence_after();
          issue_layer(tm + TM_HO, act + ACT_HH, sW + WH3_OFF, HEAD_OUT, HID2);
          issue_layer(tm + TM_SO, act + ACT_HS, sW + WS3_OFF, SEM_OUT, HID2);
          ptx::mma_commit(my_mma);
        }
      }
    }
  } else {
    // =============================== epilogue (4 warps per chain) ===============================
    const int chain = warp >> 2;
    const int row = threadIdx.x & (TILE_M - 1);  // == TMEM lane
    const uint32_t trow = tmem + chain * TM_CHAIN + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* act = smem + SM_ACT + chain * ACT_BYTES;
    const uint32_t my_mma = bar_mma + 8 * chain, my_epi = bar_epi + 8 * chain;
    uint32_t mma_ph = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      if ((it % N_CHAINS) != chain) continue;
      const long long s = tile * TILE_M + row;
      const bool valid = s < n;
      // the sample's point / direction are fetched now and consumed after two MMA round trips
      float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f};
      if (valid) sample_point(io, s, p, d, !io.density_only);
      // ---- base layers 1 and 2 -> H
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        ptx::mbar_wait(my_mma, mma_ph), mma_ph ^= 1;
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HID; c += 32) relu_store_32(trow + TM_MAIN, act, row, c);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        ptx::mbar_arrive(my_epi);
      }
      // ---- base output: density + geo features; build the head / semantic inputs
      ptx::mbar_wait(my_mma, mma_ph), mma_ph ^= 1;
      ptx::tc_fence_after();
      uint32_t o3[16];
      ptx::tmem_ld_x16(trow + TM_OUT3, o3);
      ptx::tmem_wait_ld();
      __half hb[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) hb[i] = __float2half_rn(__uint_as_float(o3[i]));
      bool inside = false;
      if (valid) {
        inside = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float xa = __fdiv_rn(__fsub_rn(p[a], fc.aabb[a]), ext[a]);
          inside = inside && (xa > 0.0f) && (xa < 1.0f);
        }
        // density = exp(x - 1) * selector  (ngp.py:79,191-193; fp16 network output upcast first)
        if (io.density) io.density[s] = inside ? expf(__fsub_rn(__half2float(hb[0]), 1.0f)) : 0.0f;
        if (io.feat) {
#pragma unroll
          for (int i = 0; i < 15; ++i) io.feat[s * 15 + i] = hb[1 + i];
        }
      }
      if (io.density_only) {
        ptx::tc_fence_before();
        ptx::mbar_arrive(my_epi);  // TMEM columns may be overwritten by this chain's next tile
        continue;
      }
      const __half dens_logit = inside ? hb[0] : __ushort_as_half((unsigned short)0xFC00);  // -inf -> sigma 0
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
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(act + ACT_XH + j * (TILE_M * 16) + row * 16) = hq[j];
#pragma unroll
        for (int j = 0; j < 2; ++j) *reinterpret_cast<uint4*>(act + ACT_XS + j * (TILE_M * 16) + row * 16) = hq[2 + j];
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(my_epi);
      // ---- head / semantic hidden layers 1 and 2
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        ptx::mbar_wait(my_mma, mma_ph), mma_ph ^= 1;
        ptx::tc_fence_after();
        relu_store_32(trow + TM_H, act + ACT_HH, row, 0);
        relu_store_32(trow + TM_H, act + ACT_HH, row, 32);
        relu_store_32(trow + TM_S, act + ACT_HS, row, 0);
        relu_store_32(trow + TM_S, act + ACT_HS, row, 32);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        ptx::mbar_arrive(my_epi);
      }
      // ---- outputs: rgb = sigmoid(head), semantic logits raw (ngp.py:210-220)
      ptx::mbar_wait(my_mma, mma_ph), mma_ph ^= 1;
      ptx::tc_fence_after();
      uint32_t oh[16], os[32];
      ptx::tmem_ld_x16(trow + TM_HO, oh);
      ptx::tmem_ld_x32(trow + TM_SO, os);
      ptx::tmem_wait_ld();
      ptx::tc_fence_before();
      ptx::mbar_arrive(my_epi);  // outputs are in registers: the next tile's layer 1 may overwrite TMEM
      if (valid && io.packed) {
        __align__(16) __half row_h[40];
        row_h[0] = dens_logit;
#pragma unroll
        for (int c = 0; c < 3; ++c) row_h[1 + c] = __float2half_rn(__uint_as_float(oh[c]));
        // halves 4-5 carry sigma = exp(logit - 1) * selector as fp32 so the compositor needs no exp for it
        reinterpret_cast<float*>(row_h)[2] = expf(__fsub_rn(__half2float(dens_logit), 1.0f));
        row_h[6] = row_h[7] = __ushort_as_half((unsigned short)0);
#pragma unroll
        for (int c = 0; c < 32; ++c) row_h[8 + c] = __float2half_rn(__uint_as_float(os[c]));
        uint4* dst = io.packed + (size_t)s * 5;
#pragma unroll
        for (int j = 0; j < 5; ++j) dst[j] = reinterpret_cast<const uint4*>(row_h)[j];
      } else if (valid) {
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
