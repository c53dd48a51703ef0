// Keras-2.2 Adam (dense and row-sparse) and the deterministic segment-sorted
// embedding-gradient scatter-add.
//
// Reference: keras.optimizers.Adam(lr) task/paper.py:656 — t<-t+1;
// lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m<-b1 m+(1-b1)g; v<-b2 v+(1-b2)g^2;
// p<-p-lr_t*m/(sqrt(v)+eps), eps=1e-7 [K].  Embedding
// gradients: the backward of keras Embedding (task/paper.py:132-138, 589-591)
// is a scatter-add of row gradients by index; here indices are sorted once,
// each unique row is owned by one warp and summed in sorted order, so the
// result is bit-reproducible (no atomics).
#include "common.cuh"

namespace lstur {

__global__ void adam_dense_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                  float* __restrict__ m, float* __restrict__ v, float lr_t, float b1, float b2,
                                  float eps, float gscale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gi = g[i] * gscale;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// Row-sparse ("lazy") Adam: only rows listed in rows[0..*n_rows) are touched.
__global__ void adam_rows_kernel(const int* __restrict__ n_rows_ptr, int D, const int* __restrict__ rows,
                                 const float* __restrict__ g_rows, float* __restrict__ p, float* __restrict__ m,
                                 float* __restrict__ v, float lr_t, float b1, float b2, float eps, float gscale) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= *n_rows_ptr) return;
  long long o = (long long)rows[r] * D;
  for (int d = lane; d < D; d += 32) {
    float gi = g_rows[(long long)r * D + d] * gscale;
    float mi = b1 * m[o + d] + (1.f - b1) * gi;
    float vi = b2 * v[o + d] + (1.f - b2) * gi * gi;
    m[o + d] = mi;
    v[o + d] = vi;
    p[o + d] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// Single-CTA bitonic sort of (key<<32 | position) for n <= SORT_MAX keys, then
// unique / segment boundaries / inverse map.  Stable by construction.
constexpr int SORT_MAX = 16384;
__global__ void __launch_bounds__(1024)
sort_unique_small_kernel(int n, const int* __restrict__ keys, int* __restrict__ sorted_pos, int* __restrict__ uniq,
                         int* __restrict__ seg_start, int* __restrict__ inverse, int* __restrict__ n_uniq) {
  extern __shared__ unsigned long long sk[];
  __shared__ int s_count;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x)
    sk[i] = i < n ? (((unsigned long long)(unsigned)keys[i] << 32) | (unsigned)i) : ~0ull;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = sk[i], b = sk[ixj];
          bool up = (i & k) == 0;
          if ((a > b) == up) { sk[i] = b; sk[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  // segment heads, numbered in sorted order with a serial-per-chunk scan (n is small)
  // pass 1: each thread counts heads in its contiguous chunk
  int per = (n + blockDim.x - 1) / blockDim.x;
  int beg = min(n, (int)threadIdx.x * per), end = min(n, beg + per);
  int cnt = 0;
  for (int i = beg; i < end; ++i) cnt += (i == 0) || ((sk[i] >> 32) != (sk[i - 1] >> 32));
  // exclusive scan of cnt over threads (blockDim <= 1024) via shared memory
  __shared__ int s_scan[1024];
  s_scan[threadIdx.x] = cnt;
  __syncthreads();
  for (int off = 1; off < (int)blockDim.x; off <<= 1) {
    int v = threadIdx.x >= (unsigned)off ? s_scan[threadIdx.x - off] : 0;
    __syncthreads();
    s_scan[threadIdx.x] += v;
    __syncthreads();
  }
  int base = s_scan[threadIdx.x] - cnt;
  for (int i = beg; i < end; ++i) {
    bool head = (i == 0) || ((sk[i] >> 32) != (sk[i - 1] >> 32));
    if (head) {
      uniq[base] = (int)(sk[i] >> 32);
      seg_start[base] = i;
      ++base;
    }
    int pos = (int)(sk[i] & 0xffffffffu);
    sorted_pos[i] = pos;
    if (inverse) inverse[pos] = base - 1;
  }
  if (threadIdx.x == blockDim.x - 1) {
    *n_uniq = s_scan[threadIdx.x];
    seg_start[s_scan[threadIdx.x]] = n;
  }
}

// out[s, :] = sum_{i in [seg_start[s], seg_start[s+1])} src[sorted_pos[i], :]  — one warp per segment.
__global__ void segment_sum_rows_kernel(int n_max, int D, const int* __restrict__ n_uniq,
                                        const int* __restrict__ seg_start, const int* __restrict__ sorted_pos,
                                        const float* __restrict__ src, long long lds, float* __restrict__ out) {
  int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= n_max || s >= *n_uniq) return;
  int i0 = seg_start[s], i1 = seg_start[s + 1];
  for (int d = lane; d < D; d += 32) {
    float acc = 0.f;
    for (int i = i0; i < i1; ++i) acc += src[(long long)sorted_pos[i] * lds + d];
    out[(long long)s * D + d] = acc;
  }
}

// table[rows[s], :] += g_rows[s, :]   (dense-gradient mode; rows are unique so no races)
__global__ void rows_add_kernel(const int* __restrict__ n_rows_ptr, int D, const int* __restrict__ rows,
                                const float* __restrict__ g_rows, float* __restrict__ table) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= *n_rows_ptr) return;
  for (int d = lane; d < D; d += 32) table[(long long)rows[r] * D + d] += g_rows[(long long)r * D + d];
}

__global__ void axpby_kernel(long long n, float a, const float* __restrict__ x, float b, float* __restrict__ y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = a * x[i] + (b != 0.f ? b * y[i] : 0.f);
}

}  // namespace lstur

using namespace lstur;

extern "C" int lstur_adam_dense(long long n, float* p, const float* g, float* m, float* v, float lr, int t, float beta1,
                                float beta2, float eps, float grad_scale, cudaStream_t stream) {
  LSTUR_REQUIRE(n >= 0 && t >= 1, "lstur_adam_dense");
  if (n == 0) return LSTUR_OK;
  float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  int grid = cdiv(n, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  adam_dense_kernel<<<grid, 256, 0, stream>>>(n, p, g, m, v, lr_t, beta1, beta2, eps, grad_scale);
  LSTUR_CHECK_LAUNCH("lstur_adam_dense");
  return LSTUR_OK;
}

extern "C" int lstur_adam_rows(int max_rows, const int* n_rows_dev, int D, const int* rows, const float* g_rows,
                               float* p, float* m, float* v, float lr, int t, float beta1, float beta2, float eps,
                               float grad_scale, cudaStream_t stream) {
  LSTUR_REQUIRE(max_rows >= 0 && D > 0 && t >= 1, "lstur_adam_rows");
  if (max_rows == 0) return LSTUR_OK;
  float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  adam_rows_kernel<<<cdiv((long long)max_rows * 32, 256), 256, 0, stream>>>(n_rows_dev, D, rows, g_rows, p, m, v, lr_t,
                                                                            beta1, beta2, eps, grad_scale);
  LSTUR_CHECK_LAUNCH("lstur_adam_rows");
  return LSTUR_OK;
}

// Index dedup (bit-exact integer path): keys[n] -> sorted_pos[n], uniq[<=n] ascending,
// seg_start[<=n+1], inverse[n] (index into uniq for each input position), n_uniq[1].
extern "C" int lstur_sort_unique_i32(int n, const int* keys, int* sorted_pos, int* uniq, int* seg_start, int* inverse,
                                     int* n_uniq, cudaStream_t stream) {
  LSTUR_REQUIRE(n >= 0 && n <= SORT_MAX, "lstur_sort_unique_i32");
  if (n == 0) {
    cudaMemsetAsync(n_uniq, 0, sizeof(int), stream);
    return LSTUR_OK;
  }
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  size_t smem = (size_t)np2 * sizeof(unsigned long long);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(sort_unique_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sort_unique_small_kernel<<<1, 1024, smem, stream>>>(n, keys, sorted_pos, uniq, seg_start, inverse, n_uniq);
  LSTUR_CHECK_LAUNCH("lstur_sort_unique_i32");
  return LSTUR_OK;
}

extern "C" int lstur_segment_sum_rows(int n, int D, const int* n_uniq, const int* seg_start, const int* sorted_pos,
                                      const float* src, long long lds, float* out, cudaStream_t stream) {
  LSTUR_REQUIRE(n >= 0 && D > 0, "lstur_segment_sum_rows");
  if (n == 0) return LSTUR_OK;
  segment_sum_rows_kernel<<<cdiv((long long)n * 32, 256), 256, 0, stream>>>(n, D, n_uniq, seg_start, sorted_pos, src, lds,
                                                                            out);
  LSTUR_CHECK_LAUNCH("lstur_segment_sum_rows");
  return LSTUR_OK;
}

extern "C" int lstur_rows_add(int max_rows, const int* n_rows_dev, int D, const int* rows, const float* g_rows,
                              float* table, cudaStream_t stream) {
  LSTUR_REQUIRE(max_rows >= 0 && D > 0, "lstur_rows_add");
  if (max_rows == 0) return LSTUR_OK;
  rows_add_kernel<<<cdiv((long long)max_rows * 32, 256), 256, 0, stream>>>(n_rows_dev, D, rows, g_rows, table);
  LSTUR_CHECK_LAUNCH("lstur_rows_add");
  return LSTUR_OK;
}

// y = a*x + b*y
extern "C" int lstur_axpby(long long n, float a, const float* x, float b, float* y, cudaStream_t stream) {
  LSTUR_REQUIRE(n >= 0, "lstur_axpby");
  if (n == 0) return LSTUR_OK;
  int grid = cdiv(n, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  axpby_kernel<<<grid, 256, 0, stream>>>(n, a, x, b, y);
  LSTUR_CHECK_LAUNCH("lstur_axpby");
  return LSTUR_OK;
}
