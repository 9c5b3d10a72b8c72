// The hand-off kernels either side of the sampling loop, all fp32 like the reference:
//   * latents -> codebook indices   VectorQuantizer.forward, src/components/VectorQuantizer.py:22-43
//                                   (called on the sampled latents at src/model/SpeechGenerator.py:117-118)
//   * channel-mean latent pooling   audio_latents[:, :, :max_length].mean(dim=1), src/TrainDiTTO.py:70-71 / :113-114
//   * MSE between eps_hat and noise nn.MSELoss(), src/TrainDiTTO.py:51,87,126
// The argmin is index-exact work: the distances are formed in fp32 in the reference's order ((|z|^2 - 2 z.c) + |c|^2) on the
// CUDA cores -- bf16 tensor-core products would flip the winner on near ties, so this GEMM deliberately stays FFMA.
#include "kernels.cuh"

namespace ditto {
namespace {

constexpr int VBM = 128, VBN = 128, VBK = 16, VTPB = 256;

// One CTA owns 128 latent rows and walks over ALL code tiles, keeping the running (distance, index) minimum of its rows
// in registers: the [rows, codes] distance matrix never exists in memory.  Ties resolve to the lowest index (torch.argmin).
__global__ void __launch_bounds__(VTPB) vq_argmin_kernel(const float* __restrict__ Z, int64_t ldz, int M, const float* __restrict__ Cb,
                                                         int64_t ldc, int K, int D, const float* __restrict__ cnorm,
                                                         long long* __restrict__ out, int channels, int64_t frames) {
  __shared__ __align__(16) float As[VBK][VBM + 4];
  __shared__ __align__(16) float Bs[VBK][VBN + 4];
  __shared__ float zz[VBM];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * VBM;
  const int tx = tid & 15, ty = tid >> 4;
  const bool vec_ok = ((ldz | ldc) & 3) == 0 && ((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(Cb)) & 15) == 0;

  float best[8];
  int best_i[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { best[i] = __int_as_float(0x7f800000); best_i[i] = 0; }
  float zpart[2] = {0.f, 0.f};

  for (int n0 = 0; n0 < K; n0 += VBN) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += VBK) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int r = (tid >> 2) + it * 64, kk = (tid & 3) * 4;
        const int gk = k0 + kk;
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (m0 + r < M) {
          const float* src = Z + static_cast<int64_t>(m0 + r) * ldz + gk;
          if (vec_ok && gk + 3 < D) {
            const float4 t = *reinterpret_cast<const float4*>(src);
            a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (gk + q < D) a[q] = src[q];
          }
        }
        if (n0 + r < K) {
          const float* src = Cb + static_cast<int64_t>(n0 + r) * ldc + gk;
          if (vec_ok && gk + 3 < D) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src));
            b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (gk + q < D) b[q] = src[q];
          }
        }
        if (n0 == 0) zpart[it] += a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];  // |z|^2 rides along with the first pass
#pragma unroll
        for (int q = 0; q < 4; ++q) { As[kk + q][r] = a[q]; Bs[kk + q][r] = b[q]; }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < VBK; ++kk) {
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
    if (n0 == 0) {  // the 4 threads that loaded a row (tid & 3) are neighbouring lanes
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        float s = zpart[it];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if ((tid & 3) == 0) zz[(tid >> 2) + it * 64] = s;
      }
      __syncthreads();
    }
    // distances of this code tile in the reference's association: (|z|^2 - 2 z.c) + |c|^2; columns scanned in ascending order
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int code = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (code >= K) continue;
      const float cn = __ldg(cnorm + code);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        const float dist = fmaf(-2.f, acc[i][j], zz[r]) + cn;
        if (dist < best[i]) { best[i] = dist; best_i[i] = code; }
      }
    }
  }
  // the 16 threads sharing `ty` are 16 consecutive lanes: butterfly over them, lowest index wins a tie
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float d = best[i];
    int c = best_i[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float d2 = __shfl_xor_sync(0xffffffffu, d, o);
      const int c2 = __shfl_xor_sync(0xffffffffu, c, o);
      if (d2 < d || (d2 == d && c2 < c)) { d = d2; c = c2; }
    }
    const int64_t row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (tx == 0 && row < M) {
      const int64_t b = row / frames, t = row - b * frames;
      for (int ch = 0; ch < channels; ++ch) out[(b * channels + ch) * frames + t] = c;
    }
  }
}

// |c|^2 per code: one warp per row, fp32
__global__ void row_sqnorm_kernel(const float* __restrict__ x, int64_t ld, int rows, int D, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = x + static_cast<int64_t>(row) * ld;
  float s = 0.f;
  for (int k = threadIdx.x & 31; k < D; k += 32) s = fmaf(p[k], p[k], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[row] = s;
}

// out[b, t, :] = (sum_c lat[b, c, t, :]) / C for t < Tout: sum in channel order then ONE division, like torch's mean
__global__ void pool_latents_kernel(const float4* __restrict__ lat, float4* __restrict__ out, int64_t C, int64_t T, int64_t Tout,
                                    int64_t D4, int64_t total) {
  const float divisor = static_cast<float>(C);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t per_b = Tout * D4;
    const int64_t b = i / per_b, r = i - b * per_b;  // r = t * D4 + k
    const float4* src = lat + b * C * T * D4 + r;
    float4 s = src[0];
    for (int64_t c = 1; c < C; ++c) {
      const float4 v = src[c * T * D4];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[i] = make_float4(s.x / divisor, s.y / divisor, s.z / divisor, s.w / divisor);
  }
}

constexpr int kMseBlocks = 592;  // 4 x 148 SMs
__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                          double* __restrict__ partial) {
  __shared__ double red[8];
  // fp32 partial sums over short runs (4 independent chains), flushed into a double every 32 iterations: B200's fp64 rate is
  // far too low to square-and-add every element in double, and short fp32 runs lose nothing that matters
  double s = 0.0;
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
  int run = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n4 = (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) ? n / 4 : 0;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 u = __ldg(a4 + i), v = __ldg(b4 + i);
    const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
    f0 = fmaf(d0, d0, f0); f1 = fmaf(d1, d1, f1); f2 = fmaf(d2, d2, f2); f3 = fmaf(d3, d3, f3);
    if (++run == 32) {
      s += static_cast<double>(f0 + f1) + static_cast<double>(f2 + f3);
      f0 = f1 = f2 = f3 = 0.f; run = 0;
    }
  }
  for (int64_t i = n4 * 4 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const float d = a[i] - b[i];
    s += static_cast<double>(d * d);
  }
  s += static_cast<double>(f0 + f1) + static_cast<double>(f2 + f3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void mse_final_kernel(const double* __restrict__ partial, int blocks, int64_t n, float* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < blocks; i += blockDim.x) s += partial[i];  // fixed order: deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += red[w];
    out[0] = static_cast<float>(t / static_cast<double>(n));
  }
}

}  // namespace
}  // namespace ditto

using namespace ditto;

extern "C" {
#pragma GCC visibility push(default)

int32_t ditto_vq_code_sqnorm(const float* codebook, int64_t codes, int64_t dim, float* sqnorm, void* stream) {
  DITTO_REQUIRE(codebook && sqnorm && codes > 0 && dim > 0 && codes < (1ll << 31) && dim < (1ll << 31), DITTO_E_BADARG,
                "vq_code_sqnorm: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(PC_ELEMENTWISE, st, 0.0, static_cast<double>(codes) * dim * 4.0);
  row_sqnorm_kernel<<<static_cast<unsigned>(ceil_div(codes, 8)), 256, 0, st>>>(codebook, dim, static_cast<int>(codes), static_cast<int>(dim),
                                                                                sqnorm);
  DITTO_LAUNCH_CHECK();
  return 0;
}

int32_t ditto_vq_encode(const float* latents, int64_t batch, int64_t frames, int64_t dim, const float* codebook, int64_t codes,
                        const float* code_sqnorm, int64_t repeat_channels, int64_t* indices, void* stream) {
  DITTO_REQUIRE(batch >= 0 && frames >= 0 && dim > 0 && codes > 0 && repeat_channels >= 1, DITTO_E_BADARG, "vq_encode: bad sizes");
  const int64_t M = batch * frames;
  DITTO_REQUIRE(M == 0 || (latents && indices), DITTO_E_BADARG, "vq_encode: null argument");
  DITTO_REQUIRE(codebook && code_sqnorm, DITTO_E_BADARG, "vq_encode: null codebook");
  DITTO_REQUIRE(M < (1ll << 31) - VBM && codes < (1ll << 31) && dim < (1ll << 31), DITTO_E_UNSUPPORTED, "vq_encode: too many rows for one call");
  if (M == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(PC_SGEMM, st, 2.0 * M * codes * dim, 0.0);
  vq_argmin_kernel<<<static_cast<unsigned>(ceil_div(M, VBM)), VTPB, 0, st>>>(latents, dim, static_cast<int>(M), codebook, dim,
                                                                             static_cast<int>(codes), static_cast<int>(dim), code_sqnorm,
                                                                             reinterpret_cast<long long*>(indices),
                                                                             static_cast<int>(repeat_channels), frames);
  DITTO_LAUNCH_CHECK();
  return 0;
}

int32_t ditto_pool_latents(const float* latents, int64_t batch, int64_t channels, int64_t frames, int64_t dim, int64_t max_frames,
                           float* out, void* stream) {
  DITTO_REQUIRE(latents && out, DITTO_E_BADARG, "pool_latents: null argument");
  DITTO_REQUIRE(batch >= 0 && channels >= 1 && frames >= 0 && dim > 0 && max_frames >= 0, DITTO_E_BADARG, "pool_latents: bad sizes");
  DITTO_REQUIRE(dim % 4 == 0 && ((reinterpret_cast<uintptr_t>(latents) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, DITTO_E_UNSUPPORTED,
                "pool_latents: dim must be a multiple of 4 and the buffers 16-byte aligned");
  const int64_t Tout = std::min(frames, max_frames);
  const int64_t total = batch * Tout * (dim / 4);
  if (total == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(PC_ELEMENTWISE, st, 0.0, static_cast<double>(total) * 16.0 * static_cast<double>(channels + 1));
  const int64_t blocks = std::min<int64_t>(ceil_div(total, 256), 148 * 16);
  pool_latents_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(reinterpret_cast<const float4*>(latents), reinterpret_cast<float4*>(out),
                                                                     channels, frames, Tout, dim / 4, total);
  DITTO_LAUNCH_CHECK();
  return 0;
}

int64_t ditto_mse_workspace_bytes(void) { return static_cast<int64_t>(kMseBlocks) * sizeof(double); }

int32_t ditto_mse_loss(const float* a, const float* b, int64_t n, float* loss, void* workspace, int64_t workspace_bytes, void* stream) {
  DITTO_REQUIRE(a && b && loss && workspace && n > 0, DITTO_E_BADARG, "mse_loss: bad argument");
  DITTO_REQUIRE(workspace_bytes >= ditto_mse_workspace_bytes(), DITTO_E_WORKSPACE, "mse_loss: workspace too small");
  DITTO_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, DITTO_E_BADARG, "mse_loss: workspace must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(PC_ELEMENTWISE, st, 0.0, static_cast<double>(n) * 8.0);
  const int blocks = static_cast<int>(std::min<int64_t>(kMseBlocks, ceil_div(n, 1024)));
  mse_partial_kernel<<<blocks, 256, 0, st>>>(a, b, n, static_cast<double*>(workspace));
  DITTO_LAUNCH_CHECK();
  mse_final_kernel<<<1, 256, 0, st>>>(static_cast<const double*>(workspace), blocks, n, loss);
  DITTO_LAUNCH_CHECK();
  return 0;
}

#pragma GCC visibility pop
}  // extern "C"
