// Flash-style self-attention for head_dim 64 (multi-head models, e.g. the DiTTO constructor defaults h = 12, d = 64;
// reference: src/components/DiT.py:117-139):
//
//     h[:, head] += softmax(alpha q k^T) v            per (utterance, head), q / k already rotated by the QKV epilogue
//
// Neither the scores nor the probabilities touch HBM (the GEMM formulation writes and re-reads 2 T^2 bytes per head: 433 MB
// per layer at C2, 4-byte stores).  TWO PASSES over the keys instead of an online rescale: pass 1 computes the score tiles
// and keeps only the row maxima, pass 2 recomputes them (a 128 x 128 x 64 MMA costs 128 cycles; the exponentials, MUFU-bound
// at 16 / clk / SM, cost 1000), writes P = exp2(s - m) as the bf16 A operand into shared memory and accumulates O += P V in
// TMEM.  No accumulator correction, no dependence of the MMA stream on the softmax statistics.
//
// Per CTA (384 threads, persistent over (utterance, head, 128-query tile) items, query tile fastest so that the K / V of one
// head stay in L2):
//   warp 0   TMA producer: Q tile once per item; K tiles [128 keys x 64] through a 3-stage ring (12 per item: 6 per pass at
//            T = 750); V tiles [128 keys x 64] (MN-major B operand, two 64-key boxes) through a 2-stage ring (pass 2)
//   warp 1   MMA issuer: S = Q K^T into two 128-column TMEM buffers (the next tile's S is issued before the current tile's
//            P.V, so the softmax of tile t+1 overlaps the P.V of tile t); O[128 x 64] += P[128 x 128] V[128 x 64]
//   warps 4-11  softmax: warp (quarter, half) owns 16 complete rows (tcgen05.ld.16x256b fragments, quad reductions only);
//            pass 2 writes P with the 128-B swizzle by hand (two K-major 64-key sub-tiles, double-buffered); final:
//            O / sum + fp32 residual -> h
// TMEM: [0, 256) two score buffers, [256, 320) O.
// Measured at C2 with the constructor-default model (L = 12, h = 12): 176 us per layer against 381 us for scores+softmax and
// P.V of the GEMM formulation.  Tried, no gain: 16 softmax warps (196 us), all key tiles of a head resident in shared memory
// and shared by two query tiles (175 us) -- the kernel is paced by the softmax warps (MUFU ex2 at 16 / clk / SM is ~45 % of
// their time), not by operand traffic or warp count.  ncu: XU pipe 90 % busy (ex2 + the bf16 packs), FMA 10 %; moving half of
// the exponentials to a degree-3 polynomial on the FMA pipe (FA4-style) made it slower (201 us): with two softmax warps per
// scheduler the 9 extra dependent instructions per element cost more issue latency than the MUFU slots they free.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace ditto {
namespace {

constexpr int FA_THREADS = 384;
constexpr int FA_EPI_WARP0 = 4;
constexpr int FA_EPI_WARPS = 8;
constexpr int FA_BM = 128, FA_BN = 128, FA_D = 64;
constexpr int FA_Q_BYTES = FA_BM * FA_D * 2;       // 16 KiB
constexpr int FA_K_BYTES = FA_BN * FA_D * 2;       // 16 KiB
constexpr int FA_V_BYTES = FA_BN * FA_D * 2;       // 16 KiB (two [64 keys x 64] boxes)
constexpr int FA_P_BYTES = FA_BM * FA_BN * 2;      // 32 KiB (two K-major 64-key sub-tiles)
constexpr int FA_K_STAGES = 3, FA_V_STAGES = 2, FA_P_BUFS = 2;
constexpr int FA_OFF_Q = 0;
constexpr int FA_OFF_K = FA_OFF_Q + FA_Q_BYTES;
constexpr int FA_OFF_V = FA_OFF_K + FA_K_STAGES * FA_K_BYTES;
constexpr int FA_OFF_P = FA_OFF_V + FA_V_STAGES * FA_V_BYTES;
constexpr int FA_OFF_BAR = FA_OFF_P + FA_P_BUFS * FA_P_BYTES;
constexpr int FA_SMEM_BYTES = FA_OFF_BAR + 256 + 1024;
static_assert(FA_SMEM_BYTES <= 232448, "exceeds the 227 KiB shared memory of an sm_100 CTA");
constexpr int FA_TMEM_COLS = 512;
constexpr int FA_O_COL = 256;

struct FaDev {
  int n_seq, heads, Tq, Tk;
  int m_tiles, k_tiles, num_items;
  float alpha2;                       // alpha * log2(e)
  float* out; const float* resid;     // fp32 [n_seq, Tq, ld]; head h occupies columns [64 h, 64 h + 64)
  long long ldo, o_seq, ldr, r_seq;
};

__global__ void __launch_bounds__(FA_THREADS, 1)
    flash_attn_d64_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                          const __grid_constant__ CUtensorMap tmap_v, const FaDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* q_smem = smem + FA_OFF_Q;
  uint8_t* k_smem = smem + FA_OFF_K;
  uint8_t* v_smem = smem + FA_OFF_V;
  uint8_t* p_smem = smem + FA_OFF_P;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + FA_OFF_BAR);
  uint64_t* q_empty = q_full + 1;
  uint64_t* k_full = q_empty + 1;               // [3]
  uint64_t* k_empty = k_full + FA_K_STAGES;     // [3]
  uint64_t* v_full = k_empty + FA_K_STAGES;     // [2]
  uint64_t* v_empty = v_full + FA_V_STAGES;     // [2]
  uint64_t* s_full = v_empty + FA_V_STAGES;     // [2] score tile complete      (MMA -> softmax)
  uint64_t* s_empty = s_full + 2;               // [2] score tile read          (softmax -> MMA)
  uint64_t* p_full = s_empty + 2;               // [2] probabilities written    (softmax -> MMA)
  uint64_t* p_empty = p_full + FA_P_BUFS;       // [2] P.V of that buffer done  (MMA -> softmax)
  uint64_t* o_full = p_empty + FA_P_BUFS;       //     O complete               (MMA -> softmax)
  uint64_t* o_empty = o_full + 1;               //     O read                   (softmax -> MMA)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + 1);
  static_assert((2 + 2 * FA_K_STAGES + 2 * FA_V_STAGES + 4 + 2 * FA_P_BUFS + 2) * 8 + 4 <= 256, "barrier block too small");

  const int warp = warp_id_uniform();   // control warps run on warp-uniform values, one elected lane issues (common.cuh: elect_one)
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < FA_K_STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < FA_V_STAGES; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], FA_EPI_WARPS); }
    for (int s = 0; s < FA_P_BUFS; ++s) { mbar_init(&p_full[s], FA_EPI_WARPS); mbar_init(&p_empty[s], 1); }
    mbar_init(o_full, 1);
    mbar_init(o_empty, FA_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<FA_TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int first = blockIdx.x, step = gridDim.x;
  const int KT = p.k_tiles;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    regs_shrink_ctrl();
    {
      const bool leader = elect_one();
      uint32_t it = 0, kc = 0, vc = 0;
      for (int item = first; item < p.num_items; item += step, ++it) {
        const int qt = item % p.m_tiles;
        const int sh = item / p.m_tiles;
        const int head = sh % p.heads, seq = sh / p.heads;
        mbar_wait(q_empty, (it & 1u) ^ 1u);
        if (leader) {
          mbar_expect_tx(q_full, FA_Q_BYTES);
          tma_load_4d(&tmap_q, q_full, q_smem, 0, qt * FA_BM, head, seq);
        }
        __syncwarp();
        for (int pass = 0; pass < 2; ++pass)
          for (int kt = 0; kt < KT; ++kt) {
            const uint32_t ks = kc % FA_K_STAGES;
            mbar_wait(&k_empty[ks], ((kc / FA_K_STAGES) & 1u) ^ 1u);
            if (leader) {
              mbar_expect_tx(&k_full[ks], FA_K_BYTES);
              tma_load_4d(&tmap_k, &k_full[ks], k_smem + ks * FA_K_BYTES, 0, kt * FA_BN, head, seq);   // keys >= Tk: zero-filled
            }
            __syncwarp();
            ++kc;
            if (pass == 1) {
              const uint32_t vs = vc % FA_V_STAGES;
              mbar_wait(&v_empty[vs], ((vc / FA_V_STAGES) & 1u) ^ 1u);
              if (leader) {
                mbar_expect_tx(&v_full[vs], FA_V_BYTES);
                // two [64 key-rows x 64 n] boxes, n contiguous (MN-major operand)
                tma_load_4d(&tmap_v, &v_full[vs], v_smem + vs * FA_V_BYTES, 0, kt * FA_BN, head, seq);
                tma_load_4d(&tmap_v, &v_full[vs], v_smem + vs * FA_V_BYTES + 64 * 128, 0, kt * FA_BN + 64, head, seq);
              }
              __syncwarp();
              ++vc;
            }
          }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (whole warp on warp-uniform values, one elected lane issues) ===========================
    regs_shrink_ctrl();
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = umma_idesc_bf16(FA_BM, FA_BN, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(FA_BM, FA_D, false, true);
      const uint64_t dq = umma_smem_desc(smem_u32(q_smem), 16, 1024);
      const uint64_t dk0 = umma_smem_desc(smem_u32(k_smem), 16, 1024);
      const uint64_t dp0 = umma_smem_desc(smem_u32(p_smem), 16, 1024);
      const uint64_t dv0 = umma_smem_desc(smem_u32(v_smem), 64 * 128, 1024);
      uint32_t it = 0, kc = 0, vc = 0, sc = 0, pc = 0;
      auto issue_s = [&]() {   // next score tile of the K stream: S[sc & 1] = Q K^T
        const uint32_t ks = kc % FA_K_STAGES, sb = sc & 1u;
        mbar_wait(&k_full[ks], (kc / FA_K_STAGES) & 1u);
        mbar_wait(&s_empty[sb], ((sc >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        if (leader) {
          const uint64_t dk = dk0 + static_cast<uint64_t>((ks * FA_K_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < FA_D / 16; ++k)
            umma_bf16(tmem_base + sb * FA_BN, dq + ((k * 32) >> 4), dk + ((k * 32) >> 4), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&k_empty[ks]);
          umma_commit(&s_full[sb]);
        }
        __syncwarp();
        ++kc; ++sc;
      };
      for (int item = first; item < p.num_items; item += step, ++it) {
        mbar_wait(q_full, it & 1u);
        tcgen05_fence_after();
        for (int kt = 0; kt < KT; ++kt) issue_s();          // pass 1: row maxima only
        issue_s();                                          // pass 2, tile 0
        for (int kt = 0; kt < KT; ++kt) {
          if (kt + 1 < KT) {
            issue_s();                                      // the next tile's scores go ahead of this tile's P.V
          } else {
            if (leader) umma_commit(q_empty);               // every Q K^T of the item has been issued: Q may be reloaded once they retire
            __syncwarp();
          }
          const uint32_t pb = pc % FA_P_BUFS, vs = vc % FA_V_STAGES;
          mbar_wait(&p_full[pb], (pc / FA_P_BUFS) & 1u);
          mbar_wait(&v_full[vs], (vc / FA_V_STAGES) & 1u);
          if (kt == 0) mbar_wait(o_empty, (it & 1u) ^ 1u);  // the previous item's O has been read
          tcgen05_fence_after();
          if (leader) {
            const uint64_t dp = dp0 + static_cast<uint64_t>((pb * FA_P_BYTES) >> 4), dv = dv0 + static_cast<uint64_t>((vs * FA_V_BYTES) >> 4);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)                    // two 64-key sub-tiles
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + FA_O_COL, dp + ((kb * (FA_BM * 128) + k * 32) >> 4), dv + ((kb * (64 * 128) + k * (16 * 128)) >> 4),
                          idesc_o, (kt | kb | k) != 0 ? 1u : 0u);
            umma_commit(&v_empty[vs]);
            umma_commit(&p_empty[pb]);
          }
          __syncwarp();
          ++pc; ++vc;
        }
        if (leader) umma_commit(o_full);
        __syncwarp();
      }
    }
  } else if (warp >= FA_EPI_WARP0) {
    // =========================== softmax + output ===========================
    regs_grow_epi();
    const int ew = warp - FA_EPI_WARP0;
    const int quarter = warp & 3, hsel = ew >> 2;
    const int g = lane >> 2, q2 = (lane & 3) * 2;
    const int trow = quarter * 32 + hsel * 16;     // this warp's 16 rows: trow + g, trow + g + 8
    uint32_t it = 0, sc = 0, pc = 0;
    for (int item = first; item < p.num_items; item += step, ++it) {
      const int qt = item % p.m_tiles;
      const int sh = item / p.m_tiles;
      const int head = sh % p.heads, seq = sh / p.heads;
      // ---------------- pass 1: row maxima (raw accumulators; alpha2 > 0 is applied once at the end) ----------------
      float mA = -INFINITY, mB = -INFINITY;
      for (int kt = 0; kt < KT; ++kt, ++sc) {
        const uint32_t sb = sc & 1u;
        mbar_wait(&s_full[sb], (sc >> 1) & 1u);
        tcgen05_fence_after();
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          uint32_t r[32];
          tmem_ld_16x64(tmem_base + (static_cast<uint32_t>(trow) << 16) + sb * FA_BN + cb * 64, r);
          tmem_ld_wait();
          const int c0 = kt * FA_BN + cb * 64 + q2;
          if (c0 - q2 + 64 <= p.Tk) {
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
              mA = fmaxf(mA, fmaxf(__uint_as_float(r[4 * kb]), __uint_as_float(r[4 * kb + 1])));
              mB = fmaxf(mB, fmaxf(__uint_as_float(r[4 * kb + 2]), __uint_as_float(r[4 * kb + 3])));
            }
          } else {
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
              const int c = c0 + kb * 8;
              if (c < p.Tk) { mA = fmaxf(mA, __uint_as_float(r[4 * kb])); mB = fmaxf(mB, __uint_as_float(r[4 * kb + 2])); }
              if (c + 1 < p.Tk) { mA = fmaxf(mA, __uint_as_float(r[4 * kb + 1])); mB = fmaxf(mB, __uint_as_float(r[4 * kb + 3])); }
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);
      }
      mA = quad_max(mA) * p.alpha2;
      mB = quad_max(mB) * p.alpha2;
      // ---------------- pass 2: P = exp2(alpha2 s - m) -> shared memory, row sums ----------------
      float2 sA2 = make_float2(0.f, 0.f), sB2 = make_float2(0.f, 0.f);
      const int rA = trow + g, rB = rA + 8;
      for (int kt = 0; kt < KT; ++kt, ++sc, ++pc) {
        const uint32_t sb = sc & 1u, pb = pc % FA_P_BUFS;
        mbar_wait(&s_full[sb], (sc >> 1) & 1u);
        tcgen05_fence_after();
        mbar_wait(&p_empty[pb], ((pc / FA_P_BUFS) & 1u) ^ 1u);   // the P.V that read this buffer two tiles ago has retired
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {                          // 64-key sub-tile cb of the probability buffer
          uint32_t r[32];
          tmem_ld_16x64(tmem_base + (static_cast<uint32_t>(trow) << 16) + sb * FA_BN + cb * 64, r);
          tmem_ld_wait();
          const int c0 = kt * FA_BN + cb * 64 + q2;
          const bool fullblk = c0 - q2 + 64 <= p.Tk;
          uint8_t* pa = p_smem + pb * FA_P_BYTES + cb * (FA_BM * 128) + rA * 128 + q2 * 2;
          uint8_t* pbp = p_smem + pb * FA_P_BYTES + cb * (FA_BM * 128) + rB * 128 + q2 * 2;
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {
            const int c = c0 + kb * 8;
            float2 eA = make_float2(ex2_approx(fmaf(__uint_as_float(r[4 * kb]), p.alpha2, -mA)),
                                    ex2_approx(fmaf(__uint_as_float(r[4 * kb + 1]), p.alpha2, -mA)));
            float2 eB = make_float2(ex2_approx(fmaf(__uint_as_float(r[4 * kb + 2]), p.alpha2, -mB)),
                                    ex2_approx(fmaf(__uint_as_float(r[4 * kb + 3]), p.alpha2, -mB)));
            if (!fullblk) {   // keys past the sequence contribute nothing
              if (c >= p.Tk) { eA.x = 0.f; eB.x = 0.f; }
              if (c + 1 >= p.Tk) { eA.y = 0.f; eB.y = 0.f; }
            }
            sA2 = __fadd2_rn(sA2, eA);
            sB2 = __fadd2_rn(sB2, eB);
            *reinterpret_cast<uint32_t*>(pa + ((kb ^ (rA & 7)) << 4)) = pack_bf16x2(eA.x, eA.y);
            *reinterpret_cast<uint32_t*>(pbp + ((kb ^ (rB & 7)) << 4)) = pack_bf16x2(eB.x, eB.y);
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_empty[sb]);
          mbar_arrive(&p_full[pb]);
        }
      }
      // ---------------- output: O / sum + residual ----------------
      const float iA = 1.0f / quad_sum(sA2.x + sA2.y), iB = 1.0f / quad_sum(sB2.x + sB2.y);
      const long long rowA = static_cast<long long>(qt) * FA_BM + rA;
      const bool okA = rowA < p.Tq, okB = rowA + 8 < p.Tq;
      const float* rp = p.resid + seq * p.r_seq + rowA * p.ldr + head * FA_D + q2;
      float* op = p.out + seq * p.o_seq + rowA * p.ldo + head * FA_D + q2;
      float2 fA[8], fB[8];
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
        fA[kb] = okA ? *reinterpret_cast<const float2*>(rp + kb * 8) : make_float2(0.f, 0.f);
        fB[kb] = okB ? *reinterpret_cast<const float2*>(rp + 8 * p.ldr + kb * 8) : make_float2(0.f, 0.f);
      }
      mbar_wait(o_full, it & 1u);
      tcgen05_fence_after();
      uint32_t r[32];
      tmem_ld_16x64(tmem_base + (static_cast<uint32_t>(trow) << 16) + FA_O_COL, r);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
        if (okA)
          *reinterpret_cast<float2*>(op + kb * 8) =
              make_float2(fmaf(iA, __uint_as_float(r[4 * kb]), fA[kb].x), fmaf(iA, __uint_as_float(r[4 * kb + 1]), fA[kb].y));
        if (okB)
          *reinterpret_cast<float2*>(op + 8 * p.ldo + kb * 8) =
              make_float2(fmaf(iB, __uint_as_float(r[4 * kb + 2]), fB[kb].x), fmaf(iB, __uint_as_float(r[4 * kb + 3]), fB[kb].y));
      }
    }
  } else {
    regs_shrink_ctrl();  // warps 2 and 3 idle; the whole warpgroup has to execute the setmaxnreg
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<FA_TMEM_COLS>(tmem_base);
  }
}

}  // namespace

bool flash_attn_supported(int d, int Tq, int Tk) { return d == FA_D && Tq >= 1 && Tk >= 1; }

int launch_flash_attn(const FlashAttnParams& q, cudaStream_t st) {
  DITTO_TRY(tc_gemm_init());
  DITTO_REQUIRE(flash_attn_supported(q.d, q.Tq, q.Tk), DITTO_E_UNSUPPORTED, "flash_attn: head_dim must be 64");
  DITTO_REQUIRE(q.Q.ptr && q.Km.ptr && q.V.ptr && q.out && q.resid, DITTO_E_BADARG, "flash_attn: null argument");
  DITTO_REQUIRE(q.n_seq >= 1 && q.heads >= 1 && q.ldo % 2 == 0 && q.ldr % 2 == 0 && q.o_seq % 2 == 0 && q.r_seq % 2 == 0, DITTO_E_BADARG,
                "flash_attn: bad sizes / strides");
  DeviceState* ds = device_state();
  if (ds == nullptr) return DITTO_E_CUDA;
  if (!ds->fa_attr) {  // per device
    DITTO_CUDA(cudaFuncSetAttribute(flash_attn_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
    ds->fa_attr = true;
  }
  CUtensorMap mq, mk, mv;
  DITTO_TRY(tc_make_map(&mq, q.Q, q.heads, q.n_seq, FA_D, FA_BM));
  DITTO_TRY(tc_make_map(&mk, q.Km, q.heads, q.n_seq, FA_D, FA_BN));
  DITTO_TRY(tc_make_map(&mv, q.V, q.heads, q.n_seq, 64, 64));
  FaDev p;
  p.n_seq = static_cast<int>(q.n_seq); p.heads = q.heads; p.Tq = q.Tq; p.Tk = q.Tk;
  p.m_tiles = static_cast<int>(ceil_div(q.Tq, FA_BM));
  p.k_tiles = static_cast<int>(ceil_div(q.Tk, FA_BN));
  const int64_t items = static_cast<int64_t>(p.m_tiles) * q.heads * q.n_seq;
  DITTO_REQUIRE(items < (1ll << 31), DITTO_E_UNSUPPORTED, "flash_attn: too many work items");
  p.num_items = static_cast<int>(items);
  p.alpha2 = q.alpha * 1.4426950408889634f;
  p.out = q.out; p.resid = q.resid; p.ldo = q.ldo; p.o_seq = q.o_seq; p.ldr = q.ldr; p.r_seq = q.r_seq;
  // algorithmic flops: 4 Tq Tk d per (utterance, head) (the recomputed Q K^T of the second pass is not counted)
  ProfScope prof(q.tag, st, 4.0 * q.Tq * q.Tk * q.d * q.heads * q.n_seq, 0.0);
  const int grid = static_cast<int>(std::min<int64_t>(tc_num_sms(), items));
  flash_attn_d64_kernel<<<grid, FA_THREADS, FA_SMEM_BYTES, st>>>(mq, mk, mv, p);
  DITTO_LAUNCH_CHECK();
  return 0;
}

}  // namespace ditto
