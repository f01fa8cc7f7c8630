// The fused radiance-field kernel (kernels 2 + 3): hash-grid gather -> base MLP -> {colour head,
// semantic head}, 128 samples per tile, persistent CTAs (one per SM), warp-specialised:
//
//   warps 0-7   epilogue    thread r <-> sample row r <-> TMEM lane r: TMEM -> registers
//                           (tcgen05.ld), ReLU, fp16, next layer's A operand -> TMEM (tcgen05.st);
//                           SH-4 of the view direction; final activations and output.
//   warps 8-9   MMA issuers one thread per chain issues tcgen05.mma / tcgen05.commit; warp 8 owns TMEM.
//   warps 10-25 encoders    512 threads (4 levels of one sample each): 8-byte hash-table gathers (L2-resident table),
//                           trilinear blend, fp16 features straight into the MMA's A tile.
//
// The 64-wide encoding stays in shared memory and all activations in TMEM; only positions come
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
constexpr int A0_STAGES = 2;

// shared-memory map (bytes): all nine weight matrices, the encoders' A-tile ring, barriers = 112 KB, which leaves
// 96 KB of the SM's unified array as L1 for the hash-grid gathers.  Activations between layers live in TMEM (below);
// only the variant with the compositor fused into the epilogue adds a 32 KB scratch region per chain.
constexpr int SM_W = 0;
constexpr int SM_A0 = SM_W + W_BYTES;                             // A0_STAGES x [128 x 64] fp16
constexpr int SM_BAR = SM_A0 + A0_STAGES * TILE_M * ENC_DIM * 2;  // mbarriers + tmem base
constexpr int SM_ACT = SM_BAR + 128;                              // N_CHAINS x 32 KB, fused compositor only
constexpr int ACT_BYTES = TILE_M * HID * 2;
constexpr int FIELD_SMEM_MIN = SM_ACT;                            // 112 KB: weights + encode ring + barriers
constexpr int FIELD_SMEM = SM_ACT + N_CHAINS * ACT_BYTES;         // + the fused compositor's per-chain scratch
constexpr int FIELD_L0_BYTES = 32768;                             // optional staging of hash-grid level 0 (16^3 entries x 8 B)
// Fused compositing scratch, one region per chain:
constexpr int ROWBUF_BYTES = 5 * TILE_M * 16;                     // 5 chunks x [128 x 16 B] raw fp16 rows
constexpr int ACT_ROWBUF = 0;
constexpr int ACT_FBUF = ACT_ROWBUF + ROWBUF_BYTES;               // [6][128] f32 per-sample terms
constexpr int ACT_WBUF = ACT_FBUF + 6 * TILE_M * 4;               // [128] f32 sample weights
constexpr int ACT_CBUF = ACT_WBUF + TILE_M * 4;                   // [128] u8 row codes
static_assert(ACT_CBUF + TILE_M <= ACT_BYTES, "compositing scratch must fit the activation region");
static_assert(FIELD_SMEM <= 232448, "field kernel shared memory exceeds 227 KB");

// TMEM column map per chain: fp32 accumulators (128 lanes x 128 columns, reused layer by layer) and, next to them,
// the fp16 A operand of the NEXT layer (two halves per column): the epilogue hands activations to the tensor core
// through TMEM (tcgen05.st -> tcgen05.mma with A from TMEM), not through shared memory, which takes the
// activation stores off the L1 / shared-memory pipe that bounds this kernel.
constexpr uint32_t TM_CHAIN = 256;
constexpr uint32_t TM_MAIN = 0;   // 128 cols: base layer 1 / 2 outputs
constexpr uint32_t TM_OUT3 = 0;   // 16 cols : base output (density, geo features)
constexpr uint32_t TM_H = 0;      // 64 cols : head hidden
constexpr uint32_t TM_S = 64;     // 64 cols : semantic hidden
constexpr uint32_t TM_HO = 0;     // 16 cols : rgb (padded)
constexpr uint32_t TM_SO = 16;    // 32 cols : semantic logits (padded)
constexpr uint32_t TM_A = 128;          // A operands start here
constexpr uint32_t TM_A_H = TM_A;       // 64 cols: H [128 x 128] fp16
constexpr uint32_t TM_A_XH = TM_A;      // 16 cols: head input [128 x 32]   (H is dead by then)
constexpr uint32_t TM_A_XS = TM_A + 16; //  8 cols: semantic input [128 x 16]
constexpr uint32_t TM_A_HH = TM_A;      // 32 cols: head hidden [128 x 64]  (XH / XS are dead)
constexpr uint32_t TM_A_HS = TM_A + 32; // 32 cols: semantic hidden [128 x 64]
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
  const float4* x01;            // optional [n]: the samples' aabb-normalised points, written by the renderer's marcher
                                // (same arithmetic as below, so the kernel need not re-derive them 5x per sample)
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
  uint4* packed;                // optional [n][5] x 16 B rows for the renderer's compositor:
                                // {sigma, r, g, b as fp32 (activated) | 32 semantic logits fp16}
  int n_sem;                    // number of semantic classes actually written (<= 32), 0 = none
  int density_only;             // stop after the base MLP
  // --- fused compositing (device-driven renderer): when `state` is set, the epilogue composites each
  // ray's samples (rows of one tile) into the per-ray state instead of writing per-sample outputs ---
  float* state;                 // [9 + n_sem][n_rays_total] structure-of-arrays
  int n_rays_total;
  int rays_per_call;
  float alpha_thre, opc_thre;
  const int* n_samp;            // [n_calls] samples per live ray this iteration
  const int* iter_samples;      // [n_calls]
  int max_samples;
  const uint8_t* s_cnt;         // [n] number of samples of the ray that STARTS at this row, else 0
  uint8_t* keep_flag;           // [n_rays_total] out: ray stays live for the next iteration
  int* total_samples;           // [n_calls] composited (alpha_thre-visible) samples
  int probabilistic;            // accumulate the variance terms
  int* ray_counts;              // optional [2][n_rays_total]: += samples evaluated / composited per ray (tests)
  int pair_x;                   // 1: x-neighbour entries through one 16-byte load where they share an aligned pair (ldg_entry_pair)
  int stage_level0;             // 1: copy level 0 of the table (<= 32 KB) into shared memory behind the kernel's own
                                // regions and gather it from there (launch with FIELD_SMEM_MIN + FIELD_L0_BYTES)
  // --- training forward (kernel instantiation TRAIN): density / rgb get the raw fp16 logits (no exp / sigmoid /
  // selector) and the activations the backward kernel needs are saved (fp16, row-major) ---
  long long save_stride;        // elements between consecutive rows of every save_* matrix (they may be column
                                // slices of one [n, 624] matrix, which makes the weight gradients ONE library GEMM)
  __half* save_enc;             // [n, 64]
  __half* save_h1;              // [n, 128]
  __half* save_h2;              // [n, 128]
  __half* save_xh;              // [n, 32]  head input  (16 SH | 15 geo | 1.0)
  __half* save_xs;              // [n, 16]  semantic input (15 geo | 1.0)
  __half* save_hh1;             // [n, 64]
  __half* save_hh2;             // [n, 64]
  __half* save_hs1;             // [n, 64]
  __half* save_hs2;             // [n, 64]
  // --- occupancy-grid update (OccGridEstimator._update, occ_grid.py:377-437): the points are jittered cells
  // of one grid level and the epilogue applies the EMA-max directly, density_only = 1 ---
  const long long* cell_ids;    // [n] cell index inside the level (x slowest, as grid_coords)
  const float* jitter;          // [n, 3] U[0,1) offsets inside the cell
  float cell_lo[3], cell_ext[3];
  int cell_res[3];
  const float* occs_old;        // [cells] snapshot of the level's occupancy values
  float* occs_new;              // [cells] updated in place of the snapshot's values
  float occ_scale, ema_decay;   // occ = density * occ_scale;  new = max(old * ema_decay, occ)
};

// (lo, hi) fp32 bit patterns -> ReLU -> one half2 word: a single F2FP with the .relu modifier (round-to-nearest
// then clamp at zero equals clamp then round) instead of two FMNMX and a pack.
__device__ __forceinline__ uint32_t pack_relu_h2(uint32_t a_bits, uint32_t b_bits) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(b_bits)), "f"(__uint_as_float(a_bits)));
  return r;
}

// TMEM accumulator columns [col0, col0 + 32) of this thread's row -> ReLU -> fp16 -> columns [col0/2, col0/2 + 16)
// of the next layer's A operand in TMEM (a_row = this thread's lane of that operand).
__device__ __forceinline__ void relu_store_32(uint32_t taddr, uint32_t a_row, int col0, __half* save_row = nullptr) {
  uint32_t v[32];
  ptx::tmem_ld_x32(taddr + col0, v);
  ptx::tmem_wait_ld();
  uint32_t q[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) q[j] = pack_relu_h2(v[2 * j], v[2 * j + 1]);
  ptx::tmem_st_x16(a_row + col0 / 2, q);
  if (save_row) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<uint4*>(save_row + col0 + 8 * j) = make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
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

// The same with the A operand in TMEM (a_tmem: first column of the operand, 8 columns per K = 16 step).
__device__ __forceinline__ void issue_layer_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t w_smem, int N, int K) {
  const uint32_t idesc = ptx::make_idesc_f16(TILE_M, N);
  const uint32_t b_lbo = N * 16;
  for (int k = 0; k < K / 16; ++k) {
    const uint64_t bd = ptx::make_smem_desc(w_smem + k * 2 * b_lbo, b_lbo, 128);
    ptx::mma_f16_ts(d_tmem, a_tmem + k * 8, bd, idesc, k > 0 ? 1u : 0u);
  }
}

__device__ __forceinline__ void sample_point(const FieldIO& io, long long s, float p[3], float d[3], bool want_dir) {
  if (io.cell_ids) {
    // x = (grid_coords + rand) / resolution;  x = aabb_lo + x * (aabb_hi - aabb_lo)   (occ_grid.py:398-403)
    const long long id = io.cell_ids[s];
    const int yz = io.cell_res[1] * io.cell_res[2];
    const int c[3] = {(int)(id / yz), (int)((id / io.cell_res[2]) % io.cell_res[1]), (int)(id % io.cell_res[2])};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float u = __fdiv_rn(__fadd_rn((float)c[a], io.jitter[3 * s + a]), (float)io.cell_res[a]);
      p[a] = __fadd_rn(io.cell_lo[a], __fmul_rn(u, io.cell_ext[a]));
    }
  } else if (io.positions) {
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


__device__ __forceinline__ void chain_bar_sync(int chain) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + chain) : "memory");
}

// Fused compositing of one tile (128 rows = whole rays, never straddling).  Same arithmetic, in the
// same order, as render_composite_kernel: transmittance weights with prefix = 1 - accumulated opacity
// (utils.py:937-944), alpha_thre filter, rgb / opacity / depth / semantic accumulation, variance against
// the UPDATED running rgb / depth (utils.py:957-999), next ray mask (utils.py:1004-1009).
// Three phases separated by chain-wide named barriers:
//   A (every row, parallel)  own sample: sigma*dt, alpha, sigmoid colours, midpoint -> shared memory;
//   B (ray's first row)      the short serial part: exclusive sum of sigma*dt, w = exp(-sum)*prefix*alpha
//                            (-> wbuf), rgb / opacity / depth sums, variances, state update, keep flag;
//   C (ray's first min(k,4) rows) semantic logits in four groups of 8 channels.
// `code` = s_cnt of this row.  All 128 threads of the chain call this.
struct CompositeSmem {
  uint8_t* rowbuf;  // 5 chunks x [128 x 16 B] raw fp16 rows (chunk-major)
  float* wbuf;      // [128] sample weights (-1: filtered by alpha_thre)
  float* fbuf;      // [6][128] sdt, alpha, tmid, col r, col g, col b
  uint8_t* cbuf;    // [128] row codes
};

__device__ __forceinline__ void composite_tile(const FieldIO& io, const CompositeSmem& sm, int chain, int row,
                                               long long tile_base, int code, float t0, float t1,
                                               const uint4 (&my_row)[5], int lane) {
  const size_t NR = (size_t)io.n_rays_total;
  const bool in_ray = code != 0, leader = in_ray && !(code & 0x80);
  const int j = leader ? 0 : (code & 0x7f);
  const int ray = in_ray ? io.ray_idx[tile_base + row] : 0;
  float* st = io.state + ray;
  // prefetch this thread's share of the ray state (latency overlaps phase A and the barriers)
  float opac = 0.f, rgb[3] = {0.f, 0.f, 0.f}, depth = 0.f, rv[3] = {0.f, 0.f, 0.f}, dv = 0.f, acc[8];
  if (leader) {
    opac = st[ST_OPA * NR], rgb[0] = st[0], rgb[1] = st[NR], rgb[2] = st[2 * NR], depth = st[ST_DEPTH * NR];
    if (io.probabilistic)
      rv[0] = st[ST_RGBVAR * NR], rv[1] = st[(ST_RGBVAR + 1) * NR], rv[2] = st[(ST_RGBVAR + 2) * NR], dv = st[ST_DVAR * NR];
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = (in_ray && j < 4 && 8 * j + c < io.n_sem) ? st[(ST_SEM + 8 * j + c) * NR] : 0.f;

  // ---- phase A (the scratch is free: the previous tile's compositing ended with a chain barrier and this
  // tile's last MMA, the final reader of the activation region it aliases, has completed)
#pragma unroll
  for (int q = 0; q < 5; ++q) *reinterpret_cast<uint4*>(sm.rowbuf + q * (TILE_M * 16) + row * 16) = my_row[q];
  sm.cbuf[row] = (uint8_t)code;
  {
    const float sdt = __fmul_rn(__uint_as_float(my_row[0].x), __fsub_rn(t1, t0));
    sm.fbuf[0 * TILE_M + row] = sdt;
    sm.fbuf[1 * TILE_M + row] = __fsub_rn(1.0f, expf(-sdt));
    sm.fbuf[2 * TILE_M + row] = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
    sm.fbuf[3 * TILE_M + row] = __uint_as_float(my_row[0].y);
    sm.fbuf[4 * TILE_M + row] = __uint_as_float(my_row[0].z);
    sm.fbuf[5 * TILE_M + row] = __uint_as_float(my_row[0].w);
  }
  chain_bar_sync(chain);
  const int k = leader ? code : (in_ray ? (int)sm.cbuf[row - j] : 0);

  // ---- phase B
  int call = -1, n_vis = 0;
  if (leader) {
    call = ray / io.rays_per_call;
    const float prefix = __fsub_rn(1.0f, opac);
    float esum = 0.f;
    for (int i = 0; i < k; ++i) {
      const int r = row + i;
      const float alpha = sm.fbuf[1 * TILE_M + r];
      const float w = __fmul_rn(__fmul_rn(expf(-esum), prefix), alpha);
      esum = __fadd_rn(esum, sm.fbuf[0 * TILE_M + r]);
      const bool vis = !(io.alpha_thre > 0.f && !(alpha >= io.alpha_thre));
      sm.wbuf[r] = vis ? w : -1.0f;  // negative marks "filtered out" (weights are >= 0)
      if (!vis) continue;
      ++n_vis;
#pragma unroll
      for (int c = 0; c < 3; ++c) rgb[c] = __fadd_rn(rgb[c], __fmul_rn(w, sm.fbuf[(3 + c) * TILE_M + r]));
      opac = __fadd_rn(opac, w);
      depth = __fadd_rn(depth, __fmul_rn(w, sm.fbuf[2 * TILE_M + r]));
    }
    if (io.probabilistic) {
      for (int i = 0; i < k; ++i) {
        const int r = row + i;
        const float w = sm.wbuf[r];
        if (w < 0.f) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float df = __fsub_rn(sm.fbuf[(3 + c) * TILE_M + r], rgb[c]);
          rv[c] = __fadd_rn(rv[c], __fmul_rn(w, __fmul_rn(df, df)));
        }
        const float dd = __fsub_rn(sm.fbuf[2 * TILE_M + r], depth);
        dv = __fadd_rn(dv, __fmul_rn(w, __fmul_rn(dd, dd)));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) st[(ST_RGBVAR + c) * NR] = rv[c];
      st[ST_DVAR * NR] = dv;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) st[c * NR] = rgb[c];
    st[ST_OPA * NR] = opac;
    st[ST_DEPTH * NR] = depth;
    const int n = io.n_samp[call];
    const bool keep = (n > 0) && (opac <= io.opc_thre) && (k == n) && (io.iter_samples[call] < io.max_samples);
    io.keep_flag[ray] = keep ? 1 : 0;
    if (io.ray_counts) io.ray_counts[ray] += k, io.ray_counts[NR + ray] += n_vis;
  }
  chain_bar_sync(chain);  // weights of every ray of the tile are in wbuf
  // ---- phase C: group g = 8 channels; rows j < min(k, 4) of the ray take groups j, j + min(k,4), ...
  if (in_ray && j < 4) {
    const int stride = k < 1 ? 1 : (k < 4 ? k : 4);
    const int r0row = row - j;
#pragma unroll 1
    for (int g = j; g < 4 && 8 * g < io.n_sem; g += stride) {
      if (g != j) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = (8 * g + c < io.n_sem) ? st[(ST_SEM + 8 * g + c) * NR] : 0.f;
      }
      for (int i = 0; i < k; ++i) {
        const float w = sm.wbuf[r0row + i];
        if (w < 0.f) continue;
        const uint4 rq = *reinterpret_cast<const uint4*>(sm.rowbuf + (1 + g) * (TILE_M * 16) + (r0row + i) * 16);
        const __half* hq = reinterpret_cast<const __half*>(&rq);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(w, __half2float(hq[c])));
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (8 * g + c < io.n_sem) st[(ST_SEM + 8 * g + c) * NR] = acc[c];
    }
  }
  const unsigned peers = __match_any_sync(0xffffffffu, call);
  const int vis_sum = __reduce_add_sync(peers, n_vis);
  if (call >= 0 && vis_sum && lane == __ffs(peers) - 1) atomicAdd(io.total_samples + call, vis_sum);
  chain_bar_sync(chain);  // every reader of the scratch is done before the next tile's activations overwrite it
}

// MODE 0: inference on explicit points / ray samples / occupancy cells (apnerf_field_forward, apnerf_occ_update);
// 1: training forward (raw outputs + saved activations, apnerf_field_forward_train); 2: the renderer's sample rows
// with the compositor fused into the epilogue (apnerf_field_forward_fused); 3: the renderer's sample rows -> packed
// fp16 rows for the stand-alone compositor (apnerf_field_forward_rows).  Each
// instantiation carries only its own code: registers and instruction-cache footprint of the hot inference kernel
// do not pay for the other two.
template <int MODE>
__global__ void __launch_bounds__(FIELD_THREADS, 1)
field_forward_kernel(const FieldIO io, const HashGridMeta meta, const FieldConst fc) {
  constexpr bool TRAIN = MODE == 1, FUSED = MODE == 2;
  constexpr bool XROWS = MODE == 2 || MODE == 3;  // renderer rows: (ray, marcher-written point) in, packed rows / state out
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = ptx::smem_u32(smem);
  // barriers: a0_full[A0_STAGES] | a0_empty[A0_STAGES] | mma_done[2] | epi_done[2] | tmem base slot
  const uint32_t bar_full = smem_base + SM_BAR, bar_empty = bar_full + 8 * A0_STAGES,
                 bar_mma = bar_empty + 8 * A0_STAGES, bar_epi = bar_mma + 8 * N_CHAINS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * (2 * A0_STAGES + 2 * N_CHAINS));

  long long n = io.n_dev ? (long long)*io.n_dev : io.n;
  if (io.n_dev && io.n > 0 && n > io.n) n = io.n;  // device-side count clamped to the buffer capacity
  const long long n_tiles = (n + TILE_M - 1) / TILE_M;

  // ---- one-time setup: weights -> smem, barriers, TMEM ----
  for (int i = threadIdx.x; i < W_BYTES / 16; i += FIELD_THREADS)
    reinterpret_cast<uint4*>(smem + SM_W)[i] = __ldg(io.weights + i);
  // optional: level 0 of the hash grid (the one level every sample of a tile shares cells of) staged in shared memory
  const uint2* staged_l0 = nullptr;
  if (MODE == 3 && io.stage_level0 && meta.size[0] * 8u <= (uint32_t)FIELD_L0_BYTES) {
    uint2* dst = reinterpret_cast<uint2*>(smem + SM_ACT);
    for (uint32_t i = threadIdx.x; i < meta.size[0]; i += FIELD_THREADS) dst[i] = __ldg(io.table + meta.offset[0] + i);
    staged_l0 = dst;
  }
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
    const int row = e & (TILE_M - 1), part = e >> 7;  // levels (2p, 2p+1, 2p+8, 2p+9) of sample `row`
    int it = 0;
    // marcher-written points: the next tile's point is fetched while this tile is encoded (one coalesced 16-byte
    // load per row instead of the ray index -> origin / direction chain and three IEEE divisions per part)
    float4 xn = make_float4(0.5f, 0.5f, 0.5f, 0.f);
    if (XROWS && (long long)blockIdx.x * TILE_M + row < n) xn = __ldg(io.x01 + (long long)blockIdx.x * TILE_M + row);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it % A0_STAGES;
      const uint32_t ph = (it / A0_STAGES) & 1;
      const long long s = tile * TILE_M + row;
      float x[3] = {0.5f, 0.5f, 0.5f};
      if (XROWS) {
        x[0] = xn.x, x[1] = xn.y, x[2] = xn.z;  // padding rows carry (0.5, 0.5, 0.5)
        const long long sn = s + (long long)gridDim.x * TILE_M;
        if (sn < n) xn = __ldg(io.x01 + sn);
      } else {
        const bool valid = s < n && (io.ray_idx == nullptr || io.ray_idx[s] >= 0);  // padding rows carry ray -1
        if (valid) {
          float p[3], d[3];
          sample_point(io, s, p, d, false);
#pragma unroll
          for (int a = 0; a < 3; ++a) x[a] = __fdiv_rn(__fsub_rn(p[a], fc.aabb[a]), ext[a]);
        }
      }
      // This thread's four levels are two 16-byte chunks of the A tile: levels (2p, 2p+1) and (2p+8, 2p+9) for
      // part p, i.e. every part has coarse (L1-resident) and fine (L2) levels and the parts finish together.
      // Software pipeline: two levels' gathers are always in flight while the previous level is blended.
      // Rows past the end / padding rows encode the point (0.5, 0.5, 0.5): their A rows are never read back.
      const int nl1 = meta.n_levels - 1;
      const int lv[4] = {2 * part, 2 * part + 1, 2 * part + 8, 2 * part + 9};
      uint2 f[4], va[8], vb[8];
      float wa[3], wb[3];
      const bool px = io.pair_x != 0;
      gather_level(meta, min(lv[0], nl1), x, io.table, va, wa, part == 0 ? staged_l0 : nullptr, px);
      gather_level(meta, min(lv[1], nl1), x, io.table, vb, wb, nullptr, px);
      f[0] = blend_level(wa, va);
      gather_level(meta, min(lv[2], nl1), x, io.table, va, wa, nullptr, px);
      f[1] = blend_level(wb, vb);
      gather_level(meta, min(lv[3], nl1), x, io.table, vb, wb, nullptr, px);
      f[2] = blend_level(wa, va);
      f[3] = blend_level(wb, vb);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lv[j] > nl1) f[j] = make_uint2(0u, 0u);  // levels the grid does not have contribute zeros
      const uint4 q[2] = {make_uint4(f[0].x, f[0].y, f[1].x, f[1].y), make_uint4(f[2].x, f[2].y, f[3].x, f[3].y)};
      ptx::mbar_wait(bar_empty + 8 * buf, ph ^ 1);  // MMA of the tile that used this slot is done
      uint8_t* a0 = smem + SM_A0 + buf * (TILE_M * ENC_DIM * 2);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int chunk = part + 4 * j;
        *reinterpret_cast<uint4*>(a0 + chunk * (TILE_M * 16) + row * 16) = q[j];
        if (TRAIN && s < n) reinterpret_cast<uint4*>(io.save_enc + s * io.save_stride)[chunk] = q[j];
      }
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
        issue_layer_ts(tm + TM_MAIN, tm + TM_A_H, sW + W2_OFF, HID, HID);
        ptx::mma_commit(my_mma);
        // base output
        ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
        ptx::tc_fence_after();
        issue_layer_ts(tm + TM_OUT3, tm + TM_A_H, sW + W3_OFF, BASE_OUT, HID);
        ptx::mma_commit(my_mma);
        if (XROWS || !io.density_only) {
          // head / semantic layer 1
          ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
          ptx::tc_fence_after();
          issue_layer_ts(tm + TM_H, tm + TM_A_XH, sW + WH1_OFF, HID2, HEAD_IN);
          issue_layer_ts(tm + TM_S, tm + TM_A_XS, sW + WS1_OFF, HID2, SEM_IN);
          ptx::mma_commit(my_mma);
          // layer 2
          ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
          ptx::tc_fence_after();
          issue_layer_ts(tm + TM_H, tm + TM_A_HH, sW + WH2_OFF, HID2, HID2);
          issue_layer_ts(tm + TM_S, tm + TM_A_HS, sW + WS2_OFF, HID2, HID2);
          ptx::mma_commit(my_mma);
          // outputs
          ptx::mbar_wait(my_epi, epi_ph), epi_ph ^= 1;
          ptx::tc_fence_after();
          issue_layer_ts(tm + TM_HO, tm + TM_A_HH, sW + WH3_OFF, HEAD_OUT, HID2);
          issue_layer_ts(tm + TM_SO, tm + TM_A_HS, sW + WS3_OFF, SEM_OUT, HID2);
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
      // the sample's point / direction are fetched early and consumed after two MMA round trips; renderer rows
      // (XROWS) read the marcher-written normalised point and only need the ray's direction
      float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f};
      bool valid;
      if (XROWS) {
        const int ray = s < n ? io.ray_idx[s] : -1;  // padding rows carry ray -1
        valid = ray >= 0;
        if (valid) {
          const float4 xs = __ldg(io.x01 + s);
          p[0] = xs.x, p[1] = xs.y, p[2] = xs.z;  // already aabb-normalised
          d[0] = io.rays_d[3 * (long long)ray], d[1] = io.rays_d[3 * (long long)ray + 1], d[2] = io.rays_d[3 * (long long)ray + 2];
        }
      } else {
        valid = s < n && (io.ray_idx == nullptr || io.ray_idx[s] >= 0);
        if (valid) sample_point(io, s, p, d, !io.density_only);
      }
      // ---- base layers 1 and 2 -> H
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        ptx::mbar_wait(my_mma, mma_ph), mma_ph ^= 1;
        ptx::tc_fence_after();
        {
          __half* save = (TRAIN && s < n) ? (layer == 0 ? io.save_h1 : io.save_h2) + s * io.save_stride : nullptr;
#pragma unroll 1
          for (int c = 0; c < HID; c += 32) relu_store_32(trow + TM_MAIN, trow + TM_A_H, c, save);
        }
        ptx::tmem_wait_st();
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
          const float xa = XROWS ? p[a] : __fdiv_rn(__fsub_rn(p[a], fc.aabb[a]), ext[a]);
          inside = inside && (xa > 0.0f) && (xa < 1.0f);
        }
        // density = exp(x - 1) * selector  (ngp.py:79,191-193; fp16 network output upcast first)
        const float dens = inside ? expf(__fsub_rn(__half2float(hb[0]), 1.0f)) : 0.0f;
        if (!XROWS && io.density) io.density[s] = TRAIN ? __half2float(hb[0]) : dens;
        if (!XROWS && io.occs_new) {
          // occs[cell] = maximum(occs[cell] * ema_decay, occ); a NaN result restores the old value (:405-434)
          const long long id = io.cell_ids[s];
          const float old = io.occs_old[id];
          const float od = __fmul_rn(old, io.ema_decay), v = __fmul_rn(dens, io.occ_scale);
          io.occs_new[id] = (od != od || v != v) ? old : fmaxf(od, v);
        }
        if (!XROWS && io.feat) {
#pragma unroll
          for (int i = 0; i < 15; ++i) io.feat[s * 15 + i] = hb[1 + i];
        }
      }
      if (!XROWS && io.density_only) {
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
        {
          const uint32_t* hw = reinterpret_cast<const uint32_t*>(hx);  // 16 words = [16 SH | 15 geo | 1.0]
          ptx::tmem_st_x16(trow + TM_A_XH, hw);
          ptx::tmem_st_x8(trow + TM_A_XS, hw + 8);  // the semantic input is the second half: [15 geo | 1.0]
        }
        if (TRAIN && s < n) {
#pragma unroll
          for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(io.save_xh + s * io.save_stride)[j] = hq[j];
#pragma unroll
          for (int j = 0; j < 2; ++j) reinterpret_cast<uint4*>(io.save_xs + s * io.save_stride)[j] = hq[2 + j];
        }
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(my_epi);
      // ---- head / semantic hidden layers 1 and 2
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        ptx::mbar_wait(my_mma, mma_ph), mma_ph ^= 1;
        ptx::tc_fence_after();
        __half* save_h = (TRAIN && s < n) ? (layer == 0 ? io.save_hh1 : io.save_hh2) + s * io.save_stride : nullptr;
        __half* save_s = (TRAIN && s < n) ? (layer == 0 ? io.save_hs1 : io.save_hs2) + s * io.save_stride : nullptr;
        relu_store_32(trow + TM_H, trow + TM_A_HH, 0, save_h);
        relu_store_32(trow + TM_H, trow + TM_A_HH, 32, save_h);
        relu_store_32(trow + TM_S, trow + TM_A_HS, 0, save_s);
        relu_store_32(trow + TM_S, trow + TM_A_HS, 32, save_s);
        ptx::tmem_wait_st();
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
      if (FUSED || (valid && (XROWS || io.packed))) {
        // row = {sigma, r, g, b as fp32 | 32 semantic logits fp16}: the activations run HERE, sample-parallel with all
        // 128 lanes busy -- sigma = exp(fp16 logit - 1) * selector (ngp.py:79,191-193), rgb = sigmoid(fp16 logit)
        // (ngp.py:211-212) -- so the ray-parallel compositor, whose lanes diverge, is left with two exps per sample
        __align__(16) __half row_h[40];
        float* row_f = reinterpret_cast<float*>(row_h);
        row_f[0] = expf(__fsub_rn(__half2float(dens_logit), 1.0f));
#pragma unroll
        for (int c = 0; c < 3; ++c)
          row_f[1 + c] = 1.0f / (1.0f + expf(-__half2float(__float2half_rn(__uint_as_float(oh[c])))));
#pragma unroll
        for (int c = 0; c < 32; ++c) row_h[8 + c] = __float2half_rn(__uint_as_float(os[c]));
        if (FUSED) {
          // ---- fused compositing: the tile's rows are exchanged through shared memory
          CompositeSmem cs;
          cs.rowbuf = act + ACT_ROWBUF;
          cs.wbuf = reinterpret_cast<float*>(act + ACT_WBUF);
          cs.fbuf = reinterpret_cast<float*>(act + ACT_FBUF);
          cs.cbuf = act + ACT_CBUF;
          uint4 my_row[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) my_row[j] = reinterpret_cast<const uint4*>(row_h)[j];
          composite_tile(io, cs, chain, row, (long long)(tile * TILE_M), valid ? (int)io.s_cnt[s] : 0,
                         valid ? io.t_starts[s] : 0.f, valid ? io.t_ends[s] : 0.f, my_row, lane);
        } else {
          uint4* dst = io.packed + (size_t)s * 5;
#pragma unroll
          for (int j = 0; j < 5; ++j) dst[j] = reinterpret_cast<const uint4*>(row_h)[j];
        }
      } else if (!XROWS && valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = __half2float(__float2half_rn(__uint_as_float(oh[c])));
          io.rgb[c * io.rgb_ch + s * io.rgb_row] = TRAIN ? v : 1.0f / (1.0f + expf(-v));
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
