// Radiance-field device code shared by field.cu (drop-in query kernels) and render.cu (the
// fused test-mode renderer): multiresolution hash-grid addressing / interpolation (kernel 2),
// SH-4 direction encoding, and the layout constants of the tcgen05 MLP (kernel 3).
//
// Specification: tiny-cuda-nn's HashGrid / SphericalHarmonics / FullyFusedMLP as configured by
// the reference at perception/models/radiance_fields/ngp.py:107-169.
// tiny-cuda-nn is an external, unpinned dependency that is absent from the reference tree, so
// parity is checked against the CPU restatement in oracle/ ("parity unpinned").
#pragma once
#include "common.cuh"

namespace apnerf {

// per-ray running state of the test-mode renderer, structure-of-arrays [channel][ray]
constexpr int ST_RGB = 0, ST_OPA = 3, ST_DEPTH = 4, ST_RGBVAR = 5, ST_DVAR = 8, ST_SEM = 9;
constexpr int MAX_ITER_SAMPLES = 64;  // the reference caps n at 64 (utils.py:902)

constexpr int MAX_LEVELS = 16;
constexpr int FEATS = 4;  // features per level (ngp.py:128)

struct HashGridMeta {  // by-value kernel parameter (constant bank)
  float scale[MAX_LEVELS];
  uint32_t res[MAX_LEVELS];
  uint32_t size[MAX_LEVELS];    // entries in the level
  uint32_t offset[MAX_LEVELS];  // first entry of the level in the table
  uint32_t hashed[MAX_LEVELS];  // 1: spatial hash, 0: dense (x fastest)
  int n_levels;
};

struct FieldConst {
  float aabb[6];
};

// x in aabb-normalised coordinates -> table entry index of corner c (bit a of c = +1 on axis a)
// and the level's interpolation weights.  Integer math is uint32 with wrap-around, as in tcnn.
__device__ __forceinline__ void level_cell(const HashGridMeta& m, int l, const float x[3], uint32_t cell[3],
                                           float w[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pos = __fmaf_rn(m.scale[l], x[a], 0.5f);
    const float fl = floorf(pos);
    cell[a] = (uint32_t)__float2int_rz(fl);
    w[a] = __fsub_rn(pos, fl);
  }
}

__device__ __forceinline__ uint32_t corner_index(const HashGridMeta& m, int l, const uint32_t cell[3], int c) {
  const uint32_t gx = cell[0] + (c & 1), gy = cell[1] + ((c >> 1) & 1), gz = cell[2] + ((c >> 2) & 1);
  uint32_t idx;
  if (m.hashed[l]) {
    idx = gx ^ (gy * 2654435761u) ^ (gz * 805459861u);
    idx &= (m.size[l] - 1u);  // hashed levels have a power-of-two size (2^log2_hashmap_size)
  } else {
    const uint32_t r = m.res[l];
    idx = gx + gy * r + gz * r * r;
    if (idx >= m.size[l]) idx %= m.size[l];  // only for points outside the unit cube
  }
  return m.offset[l] + idx;
}

__device__ __forceinline__ float corner_weight(const float w[3], int c) {
  float wt = (c & 1) ? w[0] : __fsub_rn(1.0f, w[0]);
  wt = __fmul_rn(wt, (c & 2) ? w[1] : __fsub_rn(1.0f, w[1]));
  wt = __fmul_rn(wt, (c & 4) ? w[2] : __fsub_rn(1.0f, w[2]));
  return wt;
}

// Interpolate one level: 8 x 8-byte gathers (read-only path), fp32 blend in corner order,
// result rounded to fp16 and packed as 2 x u32 (features 0,1 | 2,3).
__device__ __forceinline__ uint2 encode_level(const HashGridMeta& m, int l, const float x[3],
                                              const uint2* __restrict__ table) {
  uint32_t cell[3];
  float w[3];
  level_cell(m, l, x, cell, w);
  uint2 v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = __ldg(table + corner_index(m, l, cell, c));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float wt = corner_weight(w, c);
    const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&v[c].x));
    const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&v[c].y));
    acc[0] = __fmaf_rn(wt, f01.x, acc[0]);
    acc[1] = __fmaf_rn(wt, f01.y, acc[1]);
    acc[2] = __fmaf_rn(wt, f23.x, acc[2]);
    acc[3] = __fmaf_rn(wt, f23.y, acc[3]);
  }
  const __half2 h01 = __floats2half2_rn(acc[0], acc[1]);
  const __half2 h23 = __floats2half2_rn(acc[2], acc[3]);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&h01);
  o.y = *reinterpret_cast<const uint32_t*>(&h23);
  return o;
}


// ---- the encoder of the fused field kernel: gather and blend as two separately schedulable halves -------------
// The same arithmetic as level_cell / corner_index / corner_weight above (identical table indices, identical
// roundings), organised for instruction count: the level's hashed / dense decision is ONE warp-uniform branch, the
// y / z hash products and the dense strides are shared by the eight corners, the dense wrap-around is a conditional
// subtract (idx < 2 * size whenever the cell lies inside the level, which is every sample the marcher emits; the
// general modulo is kept for points outside the unit cube), and the xy weight products are shared between the two
// z planes.  (The generic form above compiled to ~375 SASS instructions per level; this one to about a third.)
__device__ __forceinline__ uint2 ldg_entry(const uint2* __restrict__ level_base, uint32_t idx) {
  unsigned long long addr;  // one IMAD.WIDE per corner instead of a 64-bit add + shift chain
  asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(addr) : "r"(idx), "l"(level_base));
  return __ldg(reinterpret_cast<const uint2*>(addr));
}

// `staged`: optional copy of THIS level's entries in shared memory (the coarse-level staging of the north star; see
// field_forward_kernel) -- the same entries, so the same result.
// One 16-byte load of the aligned entry pair that holds entry ia; entry ib = the x-neighbour comes out of the same load
// when it is the pair's other half (every second sample: hashed levels multiply x by 1, so x even <=> the two indices
// differ in bit 0 only; dense levels store x contiguously), else it is loaded on its own.  Same entries, same result; a
// quarter fewer load instructions and L1 tag look-ups per level -- and measured 14 % slower than eight single 8-byte
// loads (the selects and the data-dependent second load lengthen the gather -> blend chain the kernel is bound by), so
// it is an experiment switch (APNERF_FIELD_PAIR_X), not the default.
__device__ __forceinline__ void ldg_entry_pair(const uint2* __restrict__ level_base, uint32_t ia, uint32_t ib, uint2& va,
                                               uint2& vb) {
  unsigned long long addr;
  asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(addr) : "r"(ia & ~1u), "l"(level_base));
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(addr));
  const bool hi = ia & 1u;
  va = hi ? make_uint2(q.z, q.w) : make_uint2(q.x, q.y);
  if ((ia ^ ib) == 1u) vb = hi ? make_uint2(q.x, q.y) : make_uint2(q.z, q.w);
  else vb = ldg_entry(level_base, ib);
}

__device__ __forceinline__ void gather_level(const HashGridMeta& m, int l, const float x[3],
                                             const uint2* __restrict__ table, uint2 (&v)[8], float (&w)[3],
                                             const uint2* staged = nullptr, bool pair_x = false) {
  uint32_t cell[3];
  level_cell(m, l, x, cell, w);
  const uint2* __restrict__ tl = table + m.offset[l];
  const uint32_t size = m.size[l];
  uint32_t i0, i1, i2, i3, i4, i5, i6, i7;
  if (m.hashed[l]) {
    const uint32_t mask = size - 1u;
    const uint32_t hy0 = cell[1] * 2654435761u, hy1 = hy0 + 2654435761u;
    const uint32_t hz0 = cell[2] * 805459861u, hz1 = hz0 + 805459861u;
    const uint32_t x0 = cell[0], x1 = cell[0] + 1u;
    i0 = (x0 ^ hy0 ^ hz0) & mask, i1 = (x1 ^ hy0 ^ hz0) & mask;
    i2 = (x0 ^ hy1 ^ hz0) & mask, i3 = (x1 ^ hy1 ^ hz0) & mask;
    i4 = (x0 ^ hy0 ^ hz1) & mask, i5 = (x1 ^ hy0 ^ hz1) & mask;
    i6 = (x0 ^ hy1 ^ hz1) & mask, i7 = (x1 ^ hy1 ^ hz1) & mask;
  } else {
    const uint32_t r = m.res[l], r2 = r * r;
    i0 = cell[0] + cell[1] * r + cell[2] * r2;
    i1 = i0 + 1u, i2 = i0 + r, i3 = i2 + 1u, i4 = i0 + r2, i5 = i4 + 1u, i6 = i4 + r, i7 = i6 + 1u;
    if (cell[0] < r && cell[1] < r && cell[2] < r) {  // inside the level: every index < 2 * size
      i0 -= i0 >= size ? size : 0u, i1 -= i1 >= size ? size : 0u, i2 -= i2 >= size ? size : 0u;
      i3 -= i3 >= size ? size : 0u, i4 -= i4 >= size ? size : 0u, i5 -= i5 >= size ? size : 0u;
      i6 -= i6 >= size ? size : 0u, i7 -= i7 >= size ? size : 0u;
    } else {
      i0 %= size, i1 %= size, i2 %= size, i3 %= size, i4 %= size, i5 %= size, i6 %= size, i7 %= size;
    }
  }
  if (staged) {
    v[0] = staged[i0], v[1] = staged[i1], v[2] = staged[i2], v[3] = staged[i3];
    v[4] = staged[i4], v[5] = staged[i5], v[6] = staged[i6], v[7] = staged[i7];
    return;
  }
  if (pair_x) {
    ldg_entry_pair(tl, i0, i1, v[0], v[1]);
    ldg_entry_pair(tl, i2, i3, v[2], v[3]);
    ldg_entry_pair(tl, i4, i5, v[4], v[5]);
    ldg_entry_pair(tl, i6, i7, v[6], v[7]);
    return;
  }
  v[0] = ldg_entry(tl, i0), v[1] = ldg_entry(tl, i1), v[2] = ldg_entry(tl, i2), v[3] = ldg_entry(tl, i3);
  v[4] = ldg_entry(tl, i4), v[5] = ldg_entry(tl, i5), v[6] = ldg_entry(tl, i6), v[7] = ldg_entry(tl, i7);
}

// fp32 blend in corner order, one FMA per feature per corner -- issued as packed FFMA2 (fma.rn.f32x2: two
// independent round-to-nearest FMAs per instruction, bit-identical to two scalar FFMAs), result rounded to fp16.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}

__device__ __forceinline__ uint2 blend_level(const float w[3], const uint2 v[8]) {
  const float u0 = __fsub_rn(1.0f, w[0]), u1 = __fsub_rn(1.0f, w[1]), u2 = __fsub_rn(1.0f, w[2]);
  // corner c: ((x weight * y weight) * z weight), bit a of c selects w[a] over 1 - w[a]  (== corner_weight)
  const float xy[4] = {__fmul_rn(u0, u1), __fmul_rn(w[0], u1), __fmul_rn(u0, w[1]), __fmul_rn(w[0], w[1])};
  unsigned long long acc01 = 0ull, acc23 = 0ull;  // (+0.f, +0.f)
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float wt = __fmul_rn(xy[c & 3], (c & 4) ? w[2] : u2);
    const unsigned long long ww = pack_f32x2(wt, wt);
    const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&v[c].x));
    const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&v[c].y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc01) : "l"(ww), "l"(pack_f32x2(f01.x, f01.y)));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc23) : "l"(ww), "l"(pack_f32x2(f23.x, f23.y)));
  }
  float a0, a1, a2, a3;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc01));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(acc23));
  const __half2 h01 = __floats2half2_rn(a0, a1);
  const __half2 h23 = __floats2half2_rn(a2, a3);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&h01);
  o.y = *reinterpret_cast<const uint32_t*>(&h23);
  return o;
}

// Real spherical harmonics, degree 4 (16 coefficients), evaluated on d itself: the reference
// feeds (d + 1) / 2 (ngp.py:205) and tcnn maps it back with 2u - 1.  Every product / sum is
// individually rounded (the file is compiled with --fmad=false).
__device__ __forceinline__ void sh4(const float d[3], float o[16]) {
  const float u0 = (d[0] + 1.0f) / 2.0f, u1 = (d[1] + 1.0f) / 2.0f, u2 = (d[2] + 1.0f) / 2.0f;
  const float x = u0 * 2.0f - 1.0f, y = u1 * 2.0f - 1.0f, z = u2 * 2.0f - 1.0f;
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = (0.59004358992664352f * y) * (-3.0f * x2 + y2);
  o[10] = (2.8906114426405538f * xy) * z;
  o[11] = (0.45704579946446572f * y) * (1.0f - 5.0f * z2);
  o[12] = (0.3731763325901154f * z) * (5.0f * z2 - 3.0f);
  o[13] = (0.45704579946446572f * x) * (1.0f - 5.0f * z2);
  o[14] = (1.4453057213202769f * z) * (x2 - y2);
  o[15] = (0.59004358992664352f * x) * (-x2 + 3.0f * y2);
}

// ---- MLP geometry (ngp.py:134-169 with neurons = 128, layers = 2, geo_feat_dim = 15) -------
// Matrices are stored [out, in] (K-major for the MMA's B operand) as fp16 in the UMMA
// "interleaved" no-swizzle layout: element (n, k) at byte (k / 8) * (N * 16) + n * 16 + (k % 8) * 2.
constexpr int TILE_M = 128;                  // samples per tile = MMA M = TMEM lanes
constexpr int ENC_DIM = MAX_LEVELS * FEATS;  // 64
constexpr int HID = 128;                     // base MLP width
constexpr int BASE_OUT = 16;                 // 1 density + 15 geo features
constexpr int HEAD_IN = 32;                  // 16 SH + 15 geo + 1.0 pad
constexpr int SEM_IN = 16;                   // 15 geo + 1.0 pad
constexpr int HID2 = 64;                     // head / semantic MLP width
constexpr int HEAD_OUT = 16;                 // 3 rgb padded to 16
constexpr int SEM_OUT = 32;                  // <= 32 classes padded to 32

constexpr int W1_OFF = 0;                                // [128 x 64]
constexpr int W2_OFF = W1_OFF + HID * ENC_DIM * 2;       // [128 x 128]
constexpr int W3_OFF = W2_OFF + HID * HID * 2;           // [16 x 128]
constexpr int WH1_OFF = W3_OFF + BASE_OUT * HID * 2;     // [64 x 32]
constexpr int WH2_OFF = WH1_OFF + HID2 * HEAD_IN * 2;    // [64 x 64]
constexpr int WH3_OFF = WH2_OFF + HID2 * HID2 * 2;       // [16 x 64]
constexpr int WS1_OFF = WH3_OFF + HEAD_OUT * HID2 * 2;   // [64 x 16]
constexpr int WS2_OFF = WS1_OFF + HID2 * SEM_IN * 2;     // [64 x 64]
constexpr int WS3_OFF = WS2_OFF + HID2 * HID2 * 2;       // [32 x 64]
constexpr int W_BYTES = WS3_OFF + SEM_OUT * HID2 * 2;    // 81920

}  // namespace apnerf
