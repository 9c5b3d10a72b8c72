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
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct DevParams {
  int M, N, K;
  int batch_inner, batch_outer;
  int m_tiles, n_tiles, num_tiles, num_kb;
  int a_bi, a_bo, b_bi, b_bo;  // which batch coordinates each operand consumes
  float alpha;
  const float* bias;
  void* out; int out_bf16;
  long long ldo, so_inner, so_outer;
  const float* resid;
  long long ldr, sr_inner, sr_outer, resid_row_mod;
  bf16* out2; long long ldo2;
  const float* rope_cos; const float* rope_sin;
  int rope_half, rope_pd, seq_T, hidden;
};

__device__ __forceinline__ void store_f32x32(float* dst, const float (&v)[32], int ncols) {
  if (ncols == 32) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) dst[j] = v[j];
  }
}
__device__ __forceinline__ void store_bf16x32(bf16* dst, const float (&v)[32], int ncols) {
  if (ncols == 32) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 o;
      o.x = pack_bf16x2(v[j], v[j + 1]);
      o.y = pack_bf16x2(v[j + 2], v[j + 3]);
      o.z = pack_bf16x2(v[j + 4], v[j + 5]);
      o.w = pack_bf16x2(v[j + 6], v[j + 7]);
      *reinterpret_cast<uint4*>(dst + j) = o;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) dst[j] = __float2bfloat16_rn(v[j]);
  }
}

// generic chunk store used by STORE and by the v-part / fallback of QKV_ROPE
__device__ __forceinline__ void epilogue_store_chunk(const DevParams& p, const uint32_t (&r)[32], long long row, int col0,
                                                     long long out_off, long long res_off, bool row_ok) {
  const int ncols = min(32, p.N - col0);
  if (!row_ok || ncols <= 0) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = p.alpha * __uint_as_float(r[j]);
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) v[j] += __ldg(p.bias + col0 + j);
  }
  if (p.resid) {
    const long long rrow = p.resid_row_mod > 0 ? row % p.resid_row_mod : row;
    const float* rp = p.resid + res_off + rrow * p.ldr + col0;
    if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(rp + j);
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) v[j] += rp[j];
    }
  }
  if (p.out_bf16)
    store_bf16x32(static_cast<bf16*>(p.out) + out_off + row * p.ldo + col0, v, ncols);
  else
    store_f32x32(static_cast<float*>(p.out) + out_off + row * p.ldo + col0, v, ncols);
  if (p.out2) store_bf16x32(p.out2 + out_off + row * p.ldo2 + col0, v, ncols);
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
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
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
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full[as]);  // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =========================== epilogue: TMEM -> registers -> global ===========================
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the ones this warp may touch
    const int half_sel = ew >> 2;  // which 4 of the 8 column chunks
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
      const long long row = static_cast<long long>(m_blk) * BLOCK_M + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * BLOCK_N);

      if (EPI == TC_EPI_STORE) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int tcol = (half_sel * 4 + c) * 32;
          const int col0 = n_blk * BLOCK_N + tcol;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32(t_row + tcol, r);
          tmem_ld_wait();
          epilogue_store_chunk(p, r, row, col0, out_off, res_off, row_ok);
        }
      } else if (EPI == TC_EPI_GEGLU) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int tcol = (half_sel * 4 + c) * 32;
          const int col0 = n_blk * BLOCK_N + tcol;
          if (col0 >= p.N) break;
          uint32_t r[32];
          tmem_ld_32x32(t_row + tcol, r);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float a0 = __uint_as_float(r[j]) + __ldg(p.bias + col0 + j);
              const float a1 = __uint_as_float(r[j + 1]) + __ldg(p.bias + col0 + j + 1);
              const float g0 = __uint_as_float(r[16 + j]) + __ldg(p.bias + col0 + 16 + j);
              const float g1 = __uint_as_float(r[16 + j + 1]) + __ldg(p.bias + col0 + 16 + j + 1);
              o[j >> 1] = pack_bf16x2(gelu_erf_f(a0) * sigmoid_f(g0), gelu_erf_f(a1) * sigmoid_f(g1));
            }
            bf16* dst = static_cast<bf16*>(p.out) + out_off + row * p.ldo + (col0 >> 1);
            *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
      } else {  // TC_EPI_QKV_ROPE
        const int cpp = p.rope_pd >> 5;  // 32-column chunks per PD block
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int pi = half_sel * 2 + c;
          const int ch1 = (pi / cpp) * (2 * cpp) + (pi % cpp);
          const int ch2 = ch1 + cpp;
          const int pc1 = n_blk * BLOCK_N + ch1 * 32, pc2 = n_blk * BLOCK_N + ch2 * 32;
          if (pc1 >= p.N) continue;  // warp-uniform
          uint32_t r1[32], r2[32];
          tmem_ld_32x32(t_row + ch1 * 32, r1);
          tmem_ld_32x32(t_row + ch2 * 32, r2);
          tmem_ld_wait();
          if (pc1 >= 2 * p.hidden) {  // v third: identity layout, plain bias store
            epilogue_store_chunk(p, r1, row, pc1, out_off, res_off, row_ok);
            epilogue_store_chunk(p, r2, row, pc2, out_off, res_off, row_ok);
            continue;
          }
          if (!row_ok) continue;
          const int region = pc1 / p.hidden;              // 0 = q, 1 = k
          const int lp = pc1 - region * p.hidden;         // permuted column inside the region
          const int g = lp / (2 * p.rope_pd), w = lp - g * 2 * p.rope_pd;  // w < PD by construction
          const int e0 = g * p.rope_pd;                   // first "x1 element" index of this group
          const int head = e0 / p.rope_half;
          const int j0 = e0 - head * p.rope_half + w;     // rotary frequency index of column 0 of the chunk
          const int dest1 = region * p.hidden + head * 2 * p.rope_half + j0;
          const int pos = static_cast<int>(row % p.seq_T);
          const float* cp = p.rope_cos + static_cast<long long>(pos) * p.rope_half + j0;
          const float* sp = p.rope_sin + static_cast<long long>(pos) * p.rope_half + j0;
          float o1[32], o2[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 cs = *reinterpret_cast<const float4*>(cp + j);
            const float4 sn = *reinterpret_cast<const float4*>(sp + j);
            const float cc[4] = {cs.x, cs.y, cs.z, cs.w}, ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float x1 = __uint_as_float(r1[j + q]) + __ldg(p.bias + pc1 + j + q);
              const float x2 = __uint_as_float(r2[j + q]) + __ldg(p.bias + pc2 + j + q);
              o1[j + q] = x1 * cc[q] - x2 * ss[q];
              o2[j + q] = x2 * cc[q] + x1 * ss[q];
            }
          }
          bf16* dst = static_cast<bf16*>(p.out) + out_off + row * p.ldo + dest1;
          store_bf16x32(dst, o1, 32);
          store_bf16x32(dst + p.rope_half, o2, 32);
        }
      }
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

// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
bool g_init_done = false;

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
  DITTO_TRY(make_map(&ma, A, q.batch_inner, q.batch_outer, BLOCK_K, BLOCK_M));
  if (!q.b_kn)
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
  p.alpha = q.alpha; p.bias = q.bias;
  p.out = q.out; p.out_bf16 = q.out_bf16 ? 1 : 0;
  p.ldo = q.ldo; p.so_inner = q.so_inner; p.so_outer = q.so_outer;
  p.resid = q.resid; p.ldr = q.ldr; p.sr_inner = q.sr_inner; p.sr_outer = q.sr_outer; p.resid_row_mod = q.resid_row_mod;
  p.out2 = q.out2; p.ldo2 = q.ldo2;
  p.rope_cos = q.rope_cos; p.rope_sin = q.rope_sin;
  p.rope_half = q.rope_half; p.rope_pd = q.rope_pd; p.seq_T = q.seq_T; p.hidden = q.hidden;

  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, g_num_sms));
  ProfScope prof(q.tag, st, 2.0 * q.M * q.N * q.K * q.batch_inner * q.batch_outer, 0.0);
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
