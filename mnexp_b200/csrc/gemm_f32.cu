// Generic fp32 GEMM (FFMA) with split-K — the exact-precision building block.
//
//   C[M,N] (+)= op(A)[M,K] . op(B)[K,N] (+ bias[N]) (relu)
//
// Used for every "plain" dense contraction of the LSTUR path that is not the
// title Conv1D tensor-core kernel: Dense(F->U) (task/paper.py:159), the GRU
// input projection and all weight/input gradients (keras GRU, task/paper.py:612),
// and — in the fp32 verification mode — the Conv1D itself expressed as a GEMM
// over the zero-padded title buffer (A row m = Xp[m .. m+2], lda = E, K = 3E).
#include "common.cuh"

namespace lstur {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, GEMM_THREADS = 256;

template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, long long lda, const float* __restrict__ B,
                long long ldb, float* __restrict__ C, long long ldc, const float* __restrict__ bias, int flags,
                int k_per_split, float* __restrict__ partial) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, each 8x8 as 2x2 blocks of 4x4
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile (BM x BK) into As[k][m]
    if (!TA) {
      // k contiguous in memory: thread -> (row = tid/4 + 64*r, 4 consecutive k)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        int row = tid / 4 + 64 * r, kk = (tid % 4) * 4;
        int gm = m0 + row;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int gk = k0 + kk + q;
          As[kk + q][row] = (gm < M && gk < kend) ? __ldg(A + (long long)gm * lda + gk) : 0.f;
        }
      }
    } else {
      // m contiguous: thread -> (k = tid/32 + 8*r, 4 consecutive m)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        int kk = tid / 32 + 8 * r, mm = (tid % 32) * 4;
        int gk = k0 + kk;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int gm = m0 + mm + q;
          As[kk][mm + q] = (gm < M && gk < kend) ? __ldg(A + (long long)gk * lda + gm) : 0.f;
        }
      }
    }
    // ---- stage B tile (BK x BN) into Bs[k][n]
    if (!TB) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        int kk = tid / 32 + 8 * r, nn = (tid % 32) * 4;
        int gk = k0 + kk;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int gn = n0 + nn + q;
          Bs[kk][nn + q] = (gn < N && gk < kend) ? __ldg(B + (long long)gk * ldb + gn) : 0.f;
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        int col = tid / 4 + 64 * r, kk = (tid % 4) * 4;
        int gn = n0 + col;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int gk = k0 + kk + q;
          Bs[kk + q][col] = (gn < N && gk < kend) ? __ldg(B + (long long)gn * ldb + gk) : 0.f;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      *(float4*)(a) = *(const float4*)(&As[kk][ty * 4]);
      *(float4*)(a + 4) = *(const float4*)(&As[kk][64 + ty * 4]);
      *(float4*)(b) = *(const float4*)(&Bs[kk][tx * 4]);
      *(float4*)(b + 4) = *(const float4*)(&Bs[kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float v = acc[i][j];
      if (split) {
        partial[((long long)blockIdx.z * M + gm) * N + gn] = v;
      } else {
        if (bias) v += bias[gn];
        if (flags & LSTUR_GEMM_ACCUM) v += C[(long long)gm * ldc + gn];
        if (flags & LSTUR_GEMM_RELU) v = fmaxf(v, 0.f);
        C[(long long)gm * ldc + gn] = v;
      }
    }
  }
}

// Deterministic split-K reduction: fixed order over splits.
__global__ void splitk_reduce_kernel(int M, int N, int splits, const float* __restrict__ partial,
                                     float* __restrict__ C, long long ldc, const float* __restrict__ bias, int flags) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  int gm = (int)(i / N), gn = (int)(i % N);
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += partial[(long long)s * M * N + i];
  if (bias) v += bias[gn];
  if (flags & LSTUR_GEMM_ACCUM) v += C[(long long)gm * ldc + gn];
  if (flags & LSTUR_GEMM_RELU) v = fmaxf(v, 0.f);
  C[(long long)gm * ldc + gn] = v;
}

}  // namespace lstur

using namespace lstur;

extern "C" size_t lstur_gemm_f32_workspace_bytes(int M, int N, int K, int* splits_out) {
  // Split K when the output grid alone cannot fill the 148 SMs.
  long long tiles = (long long)cdiv(M, BM) * cdiv(N, BN);
  int splits = 1;
  if (tiles < 148 && K >= 4096) {
    splits = (int)((148 * 2 + tiles - 1) / tiles);
    int maxs = K / 1024;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  if (splits_out) *splits_out = splits;
  return splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0;
}

extern "C" int lstur_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, long long lda,
                              const float* B, long long ldb, float* C, long long ldc, const float* bias, int flags,
                              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  LSTUR_REQUIRE(M >= 0 && N >= 0 && K >= 0, "lstur_gemm_f32");
  if (M == 0 || N == 0) return LSTUR_OK;
  int splits = 1;
  size_t need = lstur_gemm_f32_workspace_bytes(M, N, K, &splits);
  if (need > workspace_bytes || workspace == nullptr) splits = 1;
  int kps = cdiv(cdiv(K, splits), BK) * BK;
  if (kps == 0) kps = BK;
  splits = K > 0 ? cdiv(K, kps) : 1;
  dim3 grid(cdiv(N, BN), cdiv(M, BM), splits);
  float* partial = (float*)workspace;
#define LAUNCH(TA_, TB_) \
  gemm_f32_kernel<TA_, TB_><<<grid, GEMM_THREADS, 0, stream>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, flags, kps, partial)
  if (!transA && !transB) LAUNCH(false, false);
  else if (transA && !transB) LAUNCH(true, false);
  else if (!transA && transB) LAUNCH(false, true);
  else LAUNCH(true, true);
#undef LAUNCH
  LSTUR_CHECK_LAUNCH("lstur_gemm_f32");
  if (splits > 1) {
    long long n = (long long)M * N;
    splitk_reduce_kernel<<<cdiv(n, 256), 256, 0, stream>>>(M, N, splits, partial, C, ldc, bias, flags);
    LSTUR_CHECK_LAUNCH("lstur_gemm_f32(splitk_reduce)");
  }
  return LSTUR_OK;
}
