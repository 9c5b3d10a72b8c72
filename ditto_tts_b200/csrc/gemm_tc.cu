// bf16 GEMM on the 5th-generation tensor cores of sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> 4-stage shared-memory ring -> tcgen05.mma (one elected thread,
//   UMMA 128x256x16) -> fp32 accumulators in TMEM (2 x 256 columns, double buffered) -> tcgen05.ld -> fused epilogue.
// Persistent: one CTA per SM loops over (batch, m, n) tiles, n fastest so that concurrently running CTAs share the
// A row-block in L2.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue.
//
// Epilogues (fused, so that no elementwise pass re-reads the GEMM output from HBM):
//   STORE     out = alpha*acc (+bias) (+fp32 residual), fp32 or bf16, optional extra bf16 copy
//   GEGLU     columns interleaved [fc1(16) | gate(16)]: hid = GELU_erf(a+ba) * sigmoid(g+bg)          (DiT.py:153-155)
//   QKV_ROPE  column-permuted q/k so that the RoPE partner (j, j+d/2) sits PD columns away in the tile  (DiT.py:52-72)
#include <cstdlib>

#include "kernels.cuh"

namespace ditto {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;
constexpr int BARRIER_BYTES = 256;
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;  // per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BARRIER_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES + 1024 /*align slack*/;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB shared memory of an sm_100 CTA");

struct DevParams {
  int M, N, K;
  int batch_inner, batch_outer;
  int m_tiles, n_tiles, num_tiles, num_kb;
  int a_bi, a_bo, b_bi, b_bo;  // which batch coordinates each operand consumes
  float alpha;
  const float* bias; long long sb_inner, sb_outer;
  void* out; int out_bf16;
  long long ldo, so_inner, so_outer;
  const float* resid;
  long long ldr, sr_inner, sr_outer, resid_row_mod;
  bf16* out2; long long ldo2;
  const float* rope_cos; const float* rope_sin;
  int rope_half, rope_pd, seq_T, hidden;
  int stages;  // smem ring depth actually used (<= STAGES / P_STAGES)
};

// ---------------------------------------------------------------------------------------------------
// epilogue math
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GELU_erf(a) * sigmoid(g) with 3 MUFU ops (2x ex2, 1x rcp shared by both factors).
//   Phi(a) = 0.5 (1 + erf(a / sqrt 2)), erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7), exact-erf semantics of nn.GELU()
__device__ __forceinline__ float geglu_fast(float a, float g) {
  const float z = fabsf(a) * 0.70710678118654752440f;
  const float e1 = ex2_approx(-z * z * 1.4426950408889634f);                 // exp(-z^2)
  const float e2 = ex2_approx(fminf(-g * 1.4426950408889634f, 80.f));        // exp(-g), clamped: no inf * 0
  const float da = fmaf(0.3275911f, z, 1.0f);                                // 1 + p z
  const float db = 1.0f + e2;                                                // 1 + exp(-g)
  const float r = rcp_approx(da * db);
  const float t = db * r;                                                    // 1 / (1 + p z)
  const float sig = da * r;                                                  // sigmoid(g)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float half_erfc = 0.5f * t * poly * e1;                              // 0.5 erfc(z)
  const float phi = a >= 0.f ? 1.0f - half_erfc : half_erfc;
  return a * phi * sig;
}

// ---------------------------------------------------------------------------------------------------
// Epilogue staging: each epilogue warp owns a 32 x 32 fp32 tile in shared memory (4 KiB, 16-byte chunks XOR-swizzled
// by row) used to turn the TMEM layout (thread = row) into a coalesced global layout (quarter-warp = 128 B of a row).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int stage_idx(int row, int chunk) { return row * 8 + (chunk ^ (row & 7)); }

__device__ __forceinline__ void stage_put_row(float4* stage, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) stage[stage_idx(lane, j)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// Residual prefetch for one 32 x ncols chunk in the coalesced layout (lane -> 4 columns x 8 rows).  Issued BEFORE the
// TMEM load / staging of the chunk so that the (HBM-latency) loads overlap them; out and resid alias for the in-place
// residual update, so loads and stores of one chunk must not be interleaved.
__device__ __forceinline__ void resid_prefetch(const DevParams& p, int lane, long long row0, int ocol0, int ncols,
                                               long long res_off, float4 (&rv)[8]) {
  const int c4 = lane & 7, rsub = lane >> 3;
  const int c = c4 * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c >= ncols) return;
  const bool full = c + 3 < ncols;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long row = row0 + i * 4 + rsub;
    if (row >= p.M) break;
    const long long rrow = p.resid_row_mod > 0 ? static_cast<long long>(static_cast<unsigned>(row) % static_cast<unsigned>(p.resid_row_mod)) : row;
    const float* rp = p.resid + res_off + rrow * p.ldr + ocol0 + c;
    if (full) {
      rv[i] = *reinterpret_cast<const float4*>(rp);
    } else {
      rv[i].x = rp[0];
      if (c + 1 < ncols) rv[i].y = rp[1];
      if (c + 2 < ncols) rv[i].z = rp[2];
    }
  }
}

// Flush a staged 32 x ncols tile: v = alpha * staged (+bias) (+rv) -> out (fp32 or bf16) (+ bf16 copy).
// Row r of the tile is global row row0 + r; column c is output column ocol0 + c (bias indexed by bias + c).
__device__ __forceinline__ void stage_flush(const DevParams& p, const float4* stage, int lane, long long row0, int ocol0,
                                            const float* bias, int ncols, long long out_off, bool out_bf16,
                                            const float4 (&rv)[8]) {
  const int c4 = lane & 7, rsub = lane >> 3;
  const int c = c4 * 4;
  if (c >= ncols) return;                      // after the caller's __syncwarp; no further warp-collectives inside
  const bool full = c + 3 < ncols;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias != nullptr) {
    if (full) {
      b4 = *reinterpret_cast<const float4*>(bias + c);
    } else {
      b4.x = bias[c];
      if (c + 1 < ncols) b4.y = bias[c + 1];
      if (c + 2 < ncols) b4.z = bias[c + 2];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rsub;
    const long long row = row0 + r;
    if (row >= p.M) break;
    float4 v = stage[stage_idx(r, c4)];
    v.x = fmaf(p.alpha, v.x, b4.x) + rv[i].x; v.y = fmaf(p.alpha, v.y, b4.y) + rv[i].y;
    v.z = fmaf(p.alpha, v.z, b4.z) + rv[i].z; v.w = fmaf(p.alpha, v.w, b4.w) + rv[i].w;
    const long long o = out_off + row * p.ldo + ocol0 + c;
    if (out_bf16) {
      bf16* dst = static_cast<bf16*>(p.out) + o;
      if (full) {
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      } else {
        dst[0] = __float2bfloat16_rn(v.x);
        if (c + 1 < ncols) dst[1] = __float2bfloat16_rn(v.y);
        if (c + 2 < ncols) dst[2] = __float2bfloat16_rn(v.z);
      }
    } else {
      float* dst = static_cast<float*>(p.out) + o;
      if (full) {
        *reinterpret_cast<float4*>(dst) = v;
      } else {
        dst[0] = v.x;
        if (c + 1 < ncols) dst[1] = v.y;
        if (c + 2 < ncols) dst[2] = v.z;
      }
    }
    if (p.out2) {
      bf16* d2 = p.out2 + out_off + row * p.ldo2 + ocol0 + c;
      if (full) {
        *reinterpret_cast<uint2*>(d2) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      } else {
        d2[0] = __float2bfloat16_rn(v.x);
        if (c + 1 < ncols) d2[1] = __float2bfloat16_rn(v.y);
        if (c + 2 < ncols) d2[2] = __float2bfloat16_rn(v.z);
      }
    }
  }
}

// One accumulator tile (this warp's 32 rows x 128 columns of it): TMEM -> registers -> swizzled smem -> coalesced global.
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const DevParams& p, float4* stage, int lane, int half_sel, uint32_t t_row,
                                              long long row0, int n_blk, long long out_off, long long res_off,
                                              long long bias_off) {
  float4 zero8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) zero8[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool rows_live = row0 < p.M;  // warp-uniform: nothing to store for a fully out-of-range row group

  if (EPI == TC_EPI_STORE) {
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int tcol = (half_sel * 4 + c) * 32;
      const int col0 = n_blk * BLOCK_N + tcol;
      if (col0 >= p.N || !rows_live) break;  // warp-uniform
      const int ncols = min(32, p.N - col0);
      float4 rv[8];
      if (p.resid != nullptr) {
        resid_prefetch(p, lane, row0, col0, ncols, res_off, rv);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint32_t r[32];
      tmem_ld_32x32(t_row + tcol, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      stage_put_row(stage, lane, v);
      __syncwarp();
      stage_flush(p, stage, lane, row0, col0, p.bias ? p.bias + bias_off + col0 : nullptr, ncols, out_off, p.out_bf16 != 0, rv);
      __syncwarp();
    }
  } else if (EPI == TC_EPI_GEGLU) {
    // two 32-column accumulator chunks [a16|g16][a16|g16] -> 32 hidden values per row -> one staged tile
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int tcol = (half_sel * 4 + c * 2) * 32;
      const int col0 = n_blk * BLOCK_N + tcol;
      if (col0 >= p.N || !rows_live) break;
      const bool second = col0 + 32 < p.N;  // N % 32 == 0, so the second chunk is all-or-nothing
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(t_row + tcol, r0);
      if (second) tmem_ld_32x32(t_row + tcol + 32, r1);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        v[j] = geglu_fast(__uint_as_float(r0[j]) + __ldg(p.bias + col0 + j), __uint_as_float(r0[16 + j]) + __ldg(p.bias + col0 + 16 + j));
      if (second) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[16 + j] = geglu_fast(__uint_as_float(r1[j]) + __ldg(p.bias + col0 + 32 + j),
                                 __uint_as_float(r1[16 + j]) + __ldg(p.bias + col0 + 48 + j));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 + j] = 0.f;
      }
      stage_put_row(stage, lane, v);
      __syncwarp();
      stage_flush(p, stage, lane, row0, col0 >> 1, nullptr, second ? 32 : 16, out_off, true, zero8);
      __syncwarp();
    }
  } else {  // TC_EPI_QKV_ROPE
    const int cpp = p.rope_pd >> 5;  // 32-column chunks per PD block
    const int c4 = lane & 7, rsub = lane >> 3;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int pi = half_sel * 2 + c;
      const int ch1 = (pi / cpp) * (2 * cpp) + (pi % cpp);
      const int ch2 = ch1 + cpp;
      const int pc1 = n_blk * BLOCK_N + ch1 * 32, pc2 = n_blk * BLOCK_N + ch2 * 32;
      if (pc1 >= p.N || !rows_live) continue;  // warp-uniform
      uint32_t r1[32], r2[32];
      tmem_ld_32x32(t_row + ch1 * 32, r1);
      tmem_ld_32x32(t_row + ch2 * 32, r2);
      tmem_ld_wait();
      float v[32];
      if (pc1 >= 2 * p.hidden) {  // v third: identity layout, plain bias store of both chunks
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r1[j]);
        stage_put_row(stage, lane, v);
        __syncwarp();
        stage_flush(p, stage, lane, row0, pc1, p.bias + pc1, 32, out_off, true, zero8);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r2[j]);
        stage_put_row(stage, lane, v);
        __syncwarp();
        stage_flush(p, stage, lane, row0, pc2, p.bias + pc2, 32, out_off, true, zero8);
        __syncwarp();
        continue;
      }
      // q / k thirds: x1 chunk and its RoPE partner chunk -> coalesced layout, then rotate
      float4 xa[8], xb[8];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r1[j]);
      stage_put_row(stage, lane, v);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) xa[i] = stage[stage_idx(i * 4 + rsub, c4)];
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r2[j]);
      stage_put_row(stage, lane, v);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) xb[i] = stage[stage_idx(i * 4 + rsub, c4)];
      __syncwarp();
      const int region = pc1 / p.hidden;              // 0 = q, 1 = k
      const int lp = pc1 - region * p.hidden;         // permuted column inside the region
      const int g = lp / (2 * p.rope_pd), w = lp - g * 2 * p.rope_pd;  // w < PD by construction
      const int e0 = g * p.rope_pd;                   // first "x1 element" index of this group
      const int head = e0 / p.rope_half;
      const int j0 = e0 - head * p.rope_half + w + c4 * 4;  // rotary frequency index of this lane's 4 columns
      const int dest1 = region * p.hidden + head * 2 * p.rope_half + j0;
      const float4 b1 = *reinterpret_cast<const float4*>(p.bias + pc1 + c4 * 4);
      const float4 b2 = *reinterpret_cast<const float4*>(p.bias + pc2 + c4 * 4);
      const unsigned pos_base = static_cast<unsigned>(row0) % static_cast<unsigned>(p.seq_T);  // M < 2^31 (host check)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long row = row0 + i * 4 + rsub;
        if (row >= p.M) break;
        unsigned pos_u = pos_base + static_cast<unsigned>(i * 4 + rsub);
        if (pos_u >= static_cast<unsigned>(p.seq_T)) pos_u %= static_cast<unsigned>(p.seq_T);
        const int pos = static_cast<int>(pos_u);
        const float4 cs = *reinterpret_cast<const float4*>(p.rope_cos + static_cast<long long>(pos) * p.rope_half + j0);
        const float4 sn = *reinterpret_cast<const float4*>(p.rope_sin + static_cast<long long>(pos) * p.rope_half + j0);
        const float x1x = xa[i].x + b1.x, x1y = xa[i].y + b1.y, x1z = xa[i].z + b1.z, x1w = xa[i].w + b1.w;
        const float x2x = xb[i].x + b2.x, x2y = xb[i].y + b2.y, x2z = xb[i].z + b2.z, x2w = xb[i].w + b2.w;
        bf16* dst = static_cast<bf16*>(p.out) + out_off + row * p.ldo + dest1;
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(x1x * cs.x - x2x * sn.x, x1y * cs.y - x2y * sn.y),
                                                    pack_bf16x2(x1z * cs.z - x2z * sn.z, x1w * cs.w - x2w * sn.w));
        *reinterpret_cast<uint2*>(dst + p.rope_half) =
            make_uint2(pack_bf16x2(x2x * cs.x + x1x * sn.x, x2y * cs.y + x1y * sn.y),
                       pack_bf16x2(x2z * cs.z + x1z * sn.z, x2w * cs.w + x1w * sn.w));
      }
    }
  }
}

template <int EPI, bool B_KN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const DevParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B needs 1024-B aligned stages
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float4* stage_base = reinterpret_cast<float4*>(smem + STAGES * STAGE_BYTES + BARRIER_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int m_blk = rest % p.m_tiles;
        const int b = rest / p.m_tiles;
        const int bo = b / p.batch_inner, bi = b - bo * p.batch_inner;
        const int abi = p.a_bi ? bi : 0, abo = p.a_bo ? bo : 0;
        const int bbi = p.b_bi ? bi : 0, bbo = p.b_bo ? bo : 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_4d(&tmap_a, &full_bar[stage], sa, kb * BLOCK_K, m_blk * BLOCK_M, abi, abo);
          if (!B_KN) {
            tma_load_4d(&tmap_b, &full_bar[stage], sb, kb * BLOCK_K, n_blk * BLOCK_N, bbi, bbo);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)  // [64 k-rows x 64 n] boxes, N-contiguous (MN-major operand)
              tma_load_4d(&tmap_b, &full_bar[stage], sb + j * (BLOCK_K * 128), n_blk * BLOCK_N + j * 64, kb * BLOCK_K, bbi, bbo);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (single thread) ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, false, B_KN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1u);  // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BLOCK_N);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_KN ? umma_smem_desc(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                     : umma_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full[as]);  // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =========================== epilogue: TMEM -> registers -> swizzled smem -> coalesced global ===========================
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the ones this warp may touch
    const int half_sel = ew >> 2;  // which 4 of the 8 column chunks
    float4* stage = stage_base + ew * 256;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      const int m_blk = rest % p.m_tiles;
      const int b = rest / p.m_tiles;
      const int bo = b / p.batch_inner, bi = b - bo * p.batch_inner;
      const long long out_off = bo * p.so_outer + bi * p.so_inner;
      const long long res_off = bo * p.sr_outer + bi * p.sr_inner;
      const long long bias_off = bo * p.sb_outer + bi * p.sb_inner;
      const long long row0 = static_cast<long long>(m_blk) * BLOCK_M + quarter * 32;
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * BLOCK_N);
      epilogue_tile<EPI>(p, stage, lane, half_sel, t_row, row0, n_blk, out_off, res_off, bias_off);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// =====================================================================================================
// cta_group::2 variant: a cluster of two CTAs (one SM pair) computes a 256 x 256 tile.  Each CTA loads its own 128 rows
// of A and HALF of the B tile (128 of the 256 N-rows); the leader CTA issues tcgen05.mma.cta_group::2 (UMMA 256x256x16)
// which reads both halves, so B is fetched into shared memory once per SM pair: 32 KiB per stage and CTA -> 6 stages.
// Accumulators: each CTA's TMEM holds its own 128 rows.  K-major ("NK") B only, no batching (weight GEMMs).
// =====================================================================================================
constexpr int P_STAGES = 6;
constexpr int P_B_STAGE_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;  // 16 KiB
constexpr int P_STAGE_BYTES = A_STAGE_BYTES + P_B_STAGE_BYTES;  // 32 KiB
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + BARRIER_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES + 1024;
static_assert(P_SMEM_BYTES <= 232448, "exceeds the 227 KiB shared memory of an sm_100 CTA");
static_assert((2 * P_STAGES + 4) * 8 + 4 <= BARRIER_BYTES, "barrier block too small");
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive (once the issued MMAs retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
    tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const DevParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + P_STAGES;
  uint64_t* tmem_full = empty_bar + P_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float4* stage_base = reinterpret_cast<float4*>(smem + P_STAGES * P_STAGE_BYTES + BARRIER_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_pairs = gridDim.x >> 1;
  const int pair_id = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's copy is the one in use: 1 arrive (expect_tx) + bytes of both CTAs
      mbar_init(&empty_bar[s], 1);  // one multicast commit per use, delivered to both CTAs
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * NUM_EPI_WARPS);  // leader's copy: epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();  // peer barriers initialised and TMEM allocated on both SMs before any remote arrive / MMA
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // tile = (m_pair, n_blk), n fastest; M is covered in 256-row pair tiles
  const int m_pairs = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_tiles = m_pairs * p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        const int n_blk = tile % p.n_tiles, m_pair = tile / p.n_tiles;
        const int m0 = m_pair * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M;
        const int n0 = n_blk * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * P_STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * P_STAGE_BYTES);
          tma_load_4d_2sm(&tmap_a, &full_bar[stage], sa, kb * BLOCK_K, m0, 0, 0);
          tma_load_4d_2sm(&tmap_b, &full_bar[stage], sb, kb * BLOCK_K, n0, 0, 0);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BLOCK_M, BLOCK_N, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[as], aphase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BLOCK_N);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * P_STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = umma_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
            umma_bf16_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&tmem_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;
    const int half_sel = ew >> 2;
    float4* stage = stage_base + ew * 256;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
      const int n_blk = tile % p.n_tiles, m_pair = tile / p.n_tiles;
      const long long row0 = static_cast<long long>(m_pair) * 2 * BLOCK_M + rank * BLOCK_M + quarter * 32;
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * BLOCK_N);
      epilogue_tile<EPI>(p, stage, lane, half_sel, t_row, row0, n_blk, 0, 0, 0);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();  // nobody leaves while the peer may still signal its barriers / read its shared memory
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
bool g_init_done = false;
bool g_use_pair = true;
int g_stages_1cta = STAGES, g_stages_pair = P_STAGES;

// 4-D bf16 map: dims (cols, rows, inner, outer), box (box_cols, box_rows, 1, 1), 128B swizzle, zero OOB fill.
int make_map(CUtensorMap* m, const TcOperand& op, int64_t n_inner, int64_t n_outer, int box_cols, int box_rows) {
  DITTO_REQUIRE((reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0, DITTO_E_BADARG, "tc_gemm: operand base must be 16-B aligned");
  DITTO_REQUIRE(op.ld % 8 == 0 && op.s_inner % 8 == 0 && op.s_outer % 8 == 0, DITTO_E_UNSUPPORTED,
                "tc_gemm: operand strides must be multiples of 8 bf16 (16 B)");
  const int64_t row_bytes = op.ld * 2;
  // a dimension of extent 1 still needs a legal (non-zero, 16-B multiple) stride
  const int64_t inner_bytes = (op.s_inner ? op.s_inner : op.rows * op.ld) * 2;
  const int64_t outer_bytes = (op.s_outer ? op.s_outer : std::max<int64_t>(n_inner, 1) * (op.s_inner ? op.s_inner : op.rows * op.ld)) * 2;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(op.cols), static_cast<cuuint64_t>(op.rows),
                        static_cast<cuuint64_t>(op.s_inner ? n_inner : 1), static_cast<cuuint64_t>(op.s_outer ? n_outer : 1)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(row_bytes), static_cast<cuuint64_t>(inner_bytes),
                           static_cast<cuuint64_t>(outer_bytes)};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(op.ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) + " (cols=" +
              std::to_string(op.cols) + " rows=" + std::to_string(op.rows) + " ld=" + std::to_string(op.ld) + ")");
    return DITTO_E_CUDA;
  }
  return 0;
}

template <int EPI>
int set_attr_pair() {
  DITTO_CUDA(cudaFuncSetAttribute(tc_gemm_pair_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
  return 0;
}

template <int EPI, bool B_KN>
int set_attr() {
  DITTO_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EPI, B_KN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return 0;
}

}  // namespace

int tc_gemm_init() {
  if (g_init_done) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DITTO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  DITTO_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, DITTO_E_CUDA, "cuTensorMapEncodeTiled not available");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  DITTO_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DITTO_CUDA(cudaGetDeviceProperties(&prop, dev));
  DITTO_REQUIRE(prop.major == 10, DITTO_E_UNSUPPORTED, "libditto_b200 needs an sm_100a device (B200)");
  g_num_sms = prop.multiProcessorCount;
  DITTO_TRY((set_attr<TC_EPI_STORE, false>()));
  DITTO_TRY((set_attr<TC_EPI_STORE, true>()));
  DITTO_TRY((set_attr<TC_EPI_GEGLU, false>()));
  DITTO_TRY((set_attr<TC_EPI_QKV_ROPE, false>()));
  DITTO_TRY((set_attr_pair<TC_EPI_STORE>()));
  DITTO_TRY((set_attr_pair<TC_EPI_GEGLU>()));
  DITTO_TRY((set_attr_pair<TC_EPI_QKV_ROPE>()));
  {
    const char* env = getenv("DITTO_NO_PAIR");
    g_use_pair = !(env && env[0] == '1');
    if (const char* e1 = getenv("DITTO_STAGES_1CTA")) g_stages_1cta = std::max(2, std::min(STAGES, atoi(e1)));
    if (const char* e2 = getenv("DITTO_STAGES_PAIR")) g_stages_pair = std::max(2, std::min(P_STAGES, atoi(e2)));
  }
  g_init_done = true;
  return 0;
}

int launch_tc_gemm(const TcGemmParams& q, cudaStream_t st) {
  DITTO_TRY(tc_gemm_init());
  DITTO_REQUIRE(q.M > 0 && q.N > 0 && q.K > 0 && q.batch_inner >= 1 && q.batch_outer >= 1, DITTO_E_BADARG, "tc_gemm: bad sizes");
  DITTO_REQUIRE(q.A.ptr && q.B.ptr && q.out, DITTO_E_BADARG, "tc_gemm: null operand");
  const int elt = q.out_bf16 ? 8 : 4;
  DITTO_REQUIRE(q.ldo % elt == 0 && q.so_inner % elt == 0 && q.so_outer % elt == 0, DITTO_E_UNSUPPORTED,
                "tc_gemm: output strides must keep 16-B alignment");
  if (q.resid) DITTO_REQUIRE(q.ldr % 4 == 0 && q.sr_inner % 4 == 0 && q.sr_outer % 4 == 0, DITTO_E_UNSUPPORTED, "tc_gemm: resid strides");
  if (q.out2) DITTO_REQUIRE(q.ldo2 % 8 == 0 && !q.out_bf16, DITTO_E_UNSUPPORTED, "tc_gemm: out2 needs fp32 primary output");
  if (q.epilogue == TC_EPI_GEGLU)
    DITTO_REQUIRE(q.N % 32 == 0 && q.bias && q.out_bf16 && !q.b_kn, DITTO_E_UNSUPPORTED, "tc_gemm: GEGLU epilogue constraints");
  if (q.epilogue == TC_EPI_QKV_ROPE)
    DITTO_REQUIRE(q.bias && q.out_bf16 && !q.b_kn && q.rope_cos && q.rope_sin && q.N == 3 * q.hidden &&
                      (q.rope_pd == 32 || q.rope_pd == 64 || q.rope_pd == 128) && q.rope_half % q.rope_pd == 0 &&
                      q.hidden % (2 * q.rope_pd) == 0 && q.seq_T > 0 && q.alpha == 1.f && !q.resid && !q.out2,
                  DITTO_E_UNSUPPORTED, "tc_gemm: QKV_ROPE epilogue constraints");

  CUtensorMap ma, mb;
  TcOperand A = q.A, B = q.B;
  // paired (cta_group::2) kernel for the large un-batched weight GEMMs
  const bool pair = g_use_pair && !q.b_kn && q.batch_inner == 1 && q.batch_outer == 1 && q.M >= 2 * BLOCK_M && (g_num_sms % 2 == 0);
  DITTO_TRY(make_map(&ma, A, q.batch_inner, q.batch_outer, BLOCK_K, BLOCK_M));
  if (pair)
    DITTO_TRY(make_map(&mb, B, 1, 1, BLOCK_K, BLOCK_N / 2));
  else if (!q.b_kn)
    DITTO_TRY(make_map(&mb, B, q.batch_inner, q.batch_outer, BLOCK_K, BLOCK_N));
  else
    DITTO_TRY(make_map(&mb, B, q.batch_inner, q.batch_outer, 64, BLOCK_K));

  DevParams p;
  p.M = q.M; p.N = q.N; p.K = q.K;
  p.batch_inner = q.batch_inner; p.batch_outer = q.batch_outer;
  p.m_tiles = static_cast<int>(ceil_div(q.M, BLOCK_M));
  p.n_tiles = static_cast<int>(ceil_div(q.N, BLOCK_N));
  const int64_t tiles = static_cast<int64_t>(p.m_tiles) * p.n_tiles * q.batch_inner * q.batch_outer;
  DITTO_REQUIRE(tiles < (1ll << 31), DITTO_E_UNSUPPORTED, "tc_gemm: too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  p.num_kb = static_cast<int>(ceil_div(q.K, BLOCK_K));
  p.a_bi = q.A.s_inner ? 1 : 0; p.a_bo = q.A.s_outer ? 1 : 0;
  p.b_bi = q.B.s_inner ? 1 : 0; p.b_bo = q.B.s_outer ? 1 : 0;
  p.alpha = q.alpha; p.bias = q.bias; p.sb_inner = q.sb_inner; p.sb_outer = q.sb_outer;
  p.out = q.out; p.out_bf16 = q.out_bf16 ? 1 : 0;
  p.ldo = q.ldo; p.so_inner = q.so_inner; p.so_outer = q.so_outer;
  p.resid = q.resid; p.ldr = q.ldr; p.sr_inner = q.sr_inner; p.sr_outer = q.sr_outer; p.resid_row_mod = q.resid_row_mod;
  p.out2 = q.out2; p.ldo2 = q.ldo2;
  p.rope_cos = q.rope_cos; p.rope_sin = q.rope_sin;
  p.rope_half = q.rope_half; p.rope_pd = q.rope_pd; p.seq_T = q.seq_T; p.hidden = q.hidden;

  unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, g_num_sms));
  ProfScope prof(q.tag, st, 2.0 * q.M * q.N * q.K * q.batch_inner * q.batch_outer, 0.0);
  p.stages = pair ? g_stages_pair : g_stages_1cta;
  if (pair) {
    const int64_t pair_tiles = ceil_div(q.M, 2 * BLOCK_M) * p.n_tiles;
    grid = static_cast<unsigned>(2 * std::min<int64_t>(pair_tiles, g_num_sms / 2));
    if (q.epilogue == TC_EPI_STORE)
      tc_gemm_pair_kernel<TC_EPI_STORE><<<grid, NUM_THREADS, P_SMEM_BYTES, st>>>(ma, mb, p);
    else if (q.epilogue == TC_EPI_GEGLU)
      tc_gemm_pair_kernel<TC_EPI_GEGLU><<<grid, NUM_THREADS, P_SMEM_BYTES, st>>>(ma, mb, p);
    else
      tc_gemm_pair_kernel<TC_EPI_QKV_ROPE><<<grid, NUM_THREADS, P_SMEM_BYTES, st>>>(ma, mb, p);
    DITTO_LAUNCH_CHECK();
    return 0;
  }
  if (q.epilogue == TC_EPI_STORE && !q.b_kn)
    tc_gemm_kernel<TC_EPI_STORE, false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, p);
  else if (q.epilogue == TC_EPI_STORE && q.b_kn)
    tc_gemm_kernel<TC_EPI_STORE, true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, p);
  else if (q.epilogue == TC_EPI_GEGLU)
    tc_gemm_kernel<TC_EPI_GEGLU, false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, p);
  else if (q.epilogue == TC_EPI_QKV_ROPE)
    tc_gemm_kernel<TC_EPI_QKV_ROPE, false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, p);
  else {
    set_error("tc_gemm: unknown epilogue");
    return DITTO_E_BADARG;
  }
  DITTO_LAUNCH_CHECK();
  return 0;
}

}  // namespace ditto
