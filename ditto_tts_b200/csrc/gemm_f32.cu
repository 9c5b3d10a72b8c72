// fp32 CUDA-core GEMM (FFMA): the arithmetic of the fp32 parity path (rel-L2 <= 1e-4 vs the reference) and of the
// tiny step-invariant products (time/text modulation tables).  128x128x16 tiles, 8x8 register blocking.
//   C[b] = alpha * A[b] @ op(B[b]) (+ bias[col]) (+ resid[b])
#include "kernels.cuh"

namespace ditto {

namespace {
constexpr int BM = 128, BN = 128, BK = 16, TPB = 256;

template <bool kBNK>
__global__ void __launch_bounds__(TPB) sgemm_kernel(SgemmParams p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z;
  const int bo = bz / p.batch_inner, bi = bz - bo * p.batch_inner;
  const float* __restrict__ A = p.A + bo * p.sA_outer + bi * p.sA_inner;
  const float* __restrict__ B = p.B + bo * p.sB_outer + bi * p.sB_inner;
  float* __restrict__ C = p.C + bo * p.sC_outer + bi * p.sC_inner;
  const float* __restrict__ R = p.resid ? p.resid + bo * p.sR_outer + bi * p.sR_inner : nullptr;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    // ---- A tile: 128 rows x 16 k, each thread 2 x (1 row, 4 k)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int r = (tid >> 2) + it * 64, kk = (tid & 3) * 4;
      const int gm = m0 + r, gk = k0 + kk;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gm < p.M) {
        const float* src = A + static_cast<int64_t>(gm) * p.lda + gk;
        if (gk + 3 < p.K && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          const float4 t = *reinterpret_cast<const float4*>(src);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (gk + q < p.K) v[q] = src[q];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) As[kk + q][r] = v[q];
    }
    // ---- B tile
    if (kBNK) {  // B [N, K]: same pattern as A
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int r = (tid >> 2) + it * 64, kk = (tid & 3) * 4;
        const int gn = n0 + r, gk = k0 + kk;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gn < p.N) {
          const float* src = B + static_cast<int64_t>(gn) * p.ldb + gk;
          if (gk + 3 < p.K && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const float4 t = *reinterpret_cast<const float4*>(src);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (gk + q < p.K) v[q] = src[q];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) Bs[kk + q][r] = v[q];
      }
    } else {  // B [K, N]: 16 k-rows x 128 n, each thread 2 x (1 k, 4 n)
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int kk = (tid >> 5) + it * 8, nn = (tid & 31) * 4;
        const int gk = k0 + kk, gn = n0 + nn;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gk < p.K) {
          const float* src = B + static_cast<int64_t>(gk) * p.ldb + gn;
          if (gn + 3 < p.N && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const float4 t = *reinterpret_cast<const float4*>(src);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (gn + q < p.N) v[q] = src[q];
          }
        }
        *reinterpret_cast<float4*>(&Bs[kk][nn]) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= p.N) continue;
      float v = p.alpha * acc[i][j];
      if (p.bias) v += p.bias[gn];
      if (R) v += R[static_cast<int64_t>(gm) * p.ldr + gn];
      C[static_cast<int64_t>(gm) * p.ldc + gn] = v;
    }
  }
}
}  // namespace

int launch_sgemm(const SgemmParams& p, cudaStream_t st) {
  DITTO_REQUIRE(p.M >= 0 && p.N >= 0 && p.K >= 0 && p.batch_inner >= 1 && p.batch_outer >= 1, DITTO_E_BADARG,
                "sgemm: bad sizes");
  if (p.M == 0 || p.N == 0) return 0;
  const int64_t batch = static_cast<int64_t>(p.batch_inner) * p.batch_outer;
  DITTO_REQUIRE(batch <= 65535 && ceil_div(p.M, BM) <= 65535, DITTO_E_UNSUPPORTED, "sgemm: grid too large");
  ProfScope prof(PC_SGEMM, st, 2.0 * p.M * p.N * p.K * batch, 0.0);
  dim3 grid(static_cast<unsigned>(ceil_div(p.N, BN)), static_cast<unsigned>(ceil_div(p.M, BM)), static_cast<unsigned>(batch));
  if (p.b_is_nk)
    sgemm_kernel<true><<<grid, TPB, 0, st>>>(p);
  else
    sgemm_kernel<false><<<grid, TPB, 0, st>>>(p);
  DITTO_LAUNCH_CHECK();
  return 0;
}

}  // namespace ditto
