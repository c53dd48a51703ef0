// Word-embedding gradient: deterministic segment-sorted scatter-add of the per-token input gradients of the title
// Conv1D into the rows of the word table.
//
// Reference: keras Embedding(..., trainable=config.textual_embedding_trainable) (task/paper.py:132-138, main.py:36); its
// backward is tf.IndexedSlices -> unsorted_segment_sum over the token ids (Keras / TF internal).
//
//   d word_emb[v, :] = sum over token positions p with tok[p] == v of  dX[p, :] * xdrop[p, :]
//
// No floating-point atomics: the positions are sorted by token with a stable LSD radix sort (so a token's positions stay
// in ascending order), and every token's rows are summed in that order by a fixed tree: chunks of R consecutive rows ->
// partial rows -> chunks of R partial rows -> ... (three levels; the last one sums whatever is left sequentially).  A
// step is therefore bit-reproducible, and the heavy tokens of a Zipf vocabulary (one token can own > 10 % of all
// positions) are spread over many warps instead of serialising on one.
// Positions whose gradient is identically zero are dropped before the sort: a pad position that has no real token
// within the conv window (dPre is exactly 0 under the pad mask, task/paper.py:150-155).
#include "common.cuh"

namespace lstur {
namespace wg {

constexpr int RB = 9, RADIX = 1 << RB;             // radix-sort digit
constexpr int SORT_THREADS = 256, SORT_ITEMS = 8, SORT_CH = SORT_THREADS * SORT_ITEMS;
constexpr int R = 128;                             // rows summed by one work item
constexpr int RED_THREADS = 256;                   // 8 warps = 8 work items per block

// ---- keys: token id of every live position, V for dead ones (sorted to the end and ignored) --------------------------
__global__ void keys_kernel(long long n_pos, int L, int V, const int* __restrict__ tok, int* __restrict__ keys,
                            int* __restrict__ pos, const int* __restrict__ n_titles_dev) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pos) return;
  const int t = (int)(p % L);
  const bool in_range = !n_titles_dev || p < (long long)__ldg(n_titles_dev) * L;   // compacted title list: rows beyond are unused
  int id = in_range ? tok[p] : 0;
  const int prev = (in_range && t > 0) ? tok[p - 1] : 0, next = (in_range && t + 1 < L) ? tok[p + 1] : 0;
  const bool live = (id != 0) || (prev != 0) || (next != 0);
  id = (id < 0 || id >= V) ? 0 : id;          // out-of-range ids read row 0, like the forward gather
  keys[p] = live ? id : V;
  pos[p] = (int)p;
}

// ---- stable LSD radix sort, one 9-bit digit per pass --------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS) radix_hist_kernel(const int* __restrict__ keys, long long n, int shift,
                                                                  int nblk, int* __restrict__ ghist) {
  __shared__ int h[RADIX];
  for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) h[i] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * SORT_CH;
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const long long i = base + k * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & (RADIX - 1)], 1);     // integer counts: order-independent
  }
  __syncthreads();
  for (int b = threadIdx.x; b < RADIX; b += SORT_THREADS) ghist[(long long)b * nblk + blockIdx.x] = h[b];
}

__device__ __forceinline__ int block_inclusive_scan_256(int v, int* sm /* 8 ints */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) sm[warp] = v;
  __syncthreads();
  int add = 0;
  for (int w = 0; w < warp; ++w) add += sm[w];
  __syncthreads();
  return v + add;
}

// row b of ghist (the digit's counts per block) -> exclusive prefix over the blocks; bintot[b] = the digit's total
__global__ void __launch_bounds__(256) radix_rowscan_kernel(int* __restrict__ ghist, int nblk, int* __restrict__ bintot) {
  __shared__ int sm[8];
  int* row = ghist + (long long)blockIdx.x * nblk;
  int carry = 0;
  for (int i0 = 0; i0 < nblk; i0 += 256) {
    const int i = i0 + threadIdx.x;
    const int v = i < nblk ? row[i] : 0;
    const int inc = block_inclusive_scan_256(v, sm);
    if (i < nblk) row[i] = carry + inc - v;
    __shared__ int tot;
    if (threadIdx.x == 255) tot = inc;
    __syncthreads();
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) bintot[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(const int* __restrict__ keys_in, const int* __restrict__ pos_in,
                                                                     int* __restrict__ keys_out, int* __restrict__ pos_out,
                                                                     long long n, int shift, int nblk,
                                                                     const int* __restrict__ ghist, const int* __restrict__ bintot) {
  __shared__ int cnt[SORT_THREADS / 32][RADIX];
  __shared__ int binbase[RADIX];
  __shared__ int sm[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {   // exclusive scan of the 512 digit totals: two per thread
    const int a = bintot[2 * tid], b = bintot[2 * tid + 1];
    const int inc = block_inclusive_scan_256(a + b, sm);
    binbase[2 * tid] = inc - a - b;
    binbase[2 * tid + 1] = inc - b;
  }
  for (int i = tid; i < (SORT_THREADS / 32) * RADIX; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
  __syncthreads();
  // A warp owns 256 consecutive items, 32 per round: the rank of an item inside its (warp, digit) group is the number
  // of earlier items of the group — earlier rounds (cnt) plus lower lanes of this round (match_any) — so equal digits
  // keep their input order (stable).
  const long long base = (long long)blockIdx.x * SORT_CH + warp * (SORT_ITEMS * 32);
  int key[SORT_ITEMS], ps[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const long long i = base + r * 32 + lane;
    const bool valid = i < n;
    key[r] = valid ? keys_in[i] : 0;
    ps[r] = valid ? pos_in[i] : 0;
    const int d = (key[r] >> shift) & (RADIX - 1);
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : -1 - lane);
    const int lrank = __popc(peers & ((1u << lane) - 1u));
    const int old = valid ? cnt[warp][d] : 0;
    __syncwarp();
    if (valid && lrank == 0) cnt[warp][d] = old + __popc(peers);
    __syncwarp();
    rank[r] = old + lrank;
  }
  __syncthreads();
  for (int b = tid; b < RADIX; b += SORT_THREADS) {     // exclusive prefix of the digit's counts over the warps
    int run = 0;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; ++w) {
      const int t = cnt[w][b];
      cnt[w][b] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const long long i = base + r * 32 + lane;
    if (i < n) {
      const int d = (key[r] >> shift) & (RADIX - 1);
      const long long dst = (long long)binbase[d] + ghist[(long long)d * nblk + blockIdx.x] + cnt[warp][d] + rank[r];
      keys_out[dst] = key[r];
      pos_out[dst] = ps[r];
    }
  }
}

// ---- segments of the sorted keys: seg_beg[v] / seg_end[v] (both zero-initialised; absent tokens keep 0, 0) ------------
__global__ void bounds_kernel(long long n, int V, const int* __restrict__ keys, int* __restrict__ seg_beg,
                              int* __restrict__ seg_end) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = keys[i];
  if (k >= V) return;
  if (i == 0 || keys[i - 1] != k) seg_beg[k] = (int)i;
  if (i + 1 == n || keys[i + 1] != k) seg_end[k] = (int)(i + 1);
}

// ---- reduction plan: per token v with c = seg_end - seg_beg rows
//   level 1: np1 = ceil(c / R) work items; the token keeps np1 partial rows iff np1 > 1
//   level 2: over those partial rows, np2 = ceil(np1 / R) work items (0 if np1 <= 1); keeps np2 partial rows iff np2 > 1
//   level 3: one work item iff np2 > 1, sums all np2 rows
// plan[0..5][V+1] = exclusive prefix sums of (np1, np1>1 ? np1 : 0, np2, np2>1 ? np2 : 0, np3) + totals at index V.
constexpr int PLAN_SEQ = 5;
__global__ void __launch_bounds__(1024) plan_kernel(int V, const int* __restrict__ seg_beg, const int* __restrict__ seg_end,
                                                    int* __restrict__ plan) {
  __shared__ int part[PLAN_SEQ][1024];
  const int tid = threadIdx.x;
  const int per = (V + 1023) / 1024;
  const int v0 = tid * per, v1 = min(V, v0 + per);
  auto counts = [&](int v, int* o) {
    const int c = seg_end[v] - seg_beg[v];
    const int np1 = (c + R - 1) / R, np2 = np1 > 1 ? (np1 + R - 1) / R : 0;
    o[0] = np1; o[1] = np1 > 1 ? np1 : 0; o[2] = np2; o[3] = np2 > 1 ? np2 : 0; o[4] = np2 > 1 ? 1 : 0;
  };
  int sum[PLAN_SEQ] = {0, 0, 0, 0, 0};
  for (int v = v0; v < v1; ++v) {
    int o[PLAN_SEQ];
    counts(v, o);
#pragma unroll
    for (int q = 0; q < PLAN_SEQ; ++q) sum[q] += o[q];
  }
#pragma unroll
  for (int q = 0; q < PLAN_SEQ; ++q) part[q][tid] = sum[q];
  __syncthreads();
  if (tid < PLAN_SEQ) {       // 1024 partial sums per sequence: one thread each (cheap next to the passes over V)
    int run = 0;
    for (int i = 0; i < 1024; ++i) {
      const int t = part[tid][i];
      part[tid][i] = run;
      run += t;
    }
    plan[(long long)tid * (V + 1) + V] = run;
  }
  __syncthreads();
  int run[PLAN_SEQ];
#pragma unroll
  for (int q = 0; q < PLAN_SEQ; ++q) run[q] = part[q][tid];
  for (int v = v0; v < v1; ++v) {
    int o[PLAN_SEQ];
    counts(v, o);
#pragma unroll
    for (int q = 0; q < PLAN_SEQ; ++q) {
      plan[(long long)q * (V + 1) + v] = run[q];
      run[q] += o[q];
    }
  }
}

// last index s in [0, V) with a[s] <= w (a is non-decreasing, a[V] > w)
__device__ __forceinline__ int find_segment(const int* __restrict__ a, int V, int w) {
  int lo = 0, hi = V;        // invariant: a[lo] <= w < a[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) <= w) lo = mid; else hi = mid;
  }
  return lo;
}

// Row sources.  A lane owns the 4-column pieces lane, lane+32, ... (NP of them) of a row.
struct Src16 {          // level 1, tensor-core modes: 16-bit dX rows (n_pos, Ep) + keep bytes of the X-dropout
  const uint16_t* dx;
  const uint8_t* xmask;   // or null
  int Ep, fp16;
};
struct Src32 {          // level 1, fp32 mode: fp32 rows of the padded title buffer dXp (n_titles, Lp, E) + hash replay
  const float* dx;
  int L, Lp, pl, E;
  uint32_t drop_thr, seed;
};
struct SrcPart {        // levels 2 and 3: fp32 partial rows (rows, E)
  const float* rows;
  int E;
};

template <int NP>
__device__ __forceinline__ void add_row(const Src16& s, long long pos, int lane, float (&acc)[NP][4]) {
  const uint16_t* row = s.dx + pos * s.Ep;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int piece = lane + 32 * j;
    if (piece * 4 < s.Ep) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(row + piece * 4));
      unsigned m = 0xffu;
      if (s.xmask) m = __ldg(s.xmask + pos * (s.Ep >> 3) + (piece >> 1));
      const int wj = (piece & 1) * 2;      // 32-bit word of the 16-byte piece; bit wj+h / 4+wj+h = low / high half of word wj+h
      float v[4];
      if (s.fp16) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
      } else {
        v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
        v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
      }
      acc[j][0] += ((m >> wj) & 1u) ? v[0] : 0.f;
      acc[j][1] += ((m >> (4 + wj)) & 1u) ? v[1] : 0.f;
      acc[j][2] += ((m >> (wj + 1)) & 1u) ? v[2] : 0.f;
      acc[j][3] += ((m >> (5 + wj)) & 1u) ? v[3] : 0.f;
    }
  }
}
template <int NP>
__device__ __forceinline__ void add_row(const Src32& s, long long pos, int lane, float (&acc)[NP][4]) {
  const long long n = pos / s.L;
  const int t = (int)(pos % s.L);
  const float* row = s.dx + (n * s.Lp + t + s.pl) * s.E;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int e = (lane + 32 * j) * 4;
    if (e < s.E) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row + e));
      bool k0 = true, k1 = true, k2 = true, k3 = true;
      if (s.drop_thr) {
        const uint64_t base = (uint64_t)pos * (uint64_t)s.E + e;
        k0 = (rng_u32(s.seed, base + 0) >> 8) >= s.drop_thr; k1 = (rng_u32(s.seed, base + 1) >> 8) >= s.drop_thr;
        k2 = (rng_u32(s.seed, base + 2) >> 8) >= s.drop_thr; k3 = (rng_u32(s.seed, base + 3) >> 8) >= s.drop_thr;
      }
      acc[j][0] += k0 ? v.x : 0.f; acc[j][1] += k1 ? v.y : 0.f; acc[j][2] += k2 ? v.z : 0.f; acc[j][3] += k3 ? v.w : 0.f;
    }
  }
}
template <int NP>
__device__ __forceinline__ void add_row(const SrcPart& s, long long r, int lane, float (&acc)[NP][4]) {
  const float* row = s.rows + r * s.E;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int e = (lane + 32 * j) * 4;
    if (e < s.E) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row + e));
      acc[j][0] += v.x; acc[j][1] += v.y; acc[j][2] += v.z; acc[j][3] += v.w;
    }
  }
}

// One warp per work item w of level LEVEL (1, 2, 3):
//   token s = find_segment(wstart, w), k = w - wstart[s]
//   rows    = item range [ibeg + k*R, min(iend, ibeg + (k+1)*R))  (level 3: the whole range), taken in order
//   output  = the token's row of d_word_emb (x scale) if this is its only work item at this level, else partial row
//             pstart[s] + k of this level's partial buffer.
// Level 1 rows are reached through the sorted position list; levels 2 / 3 read the previous level's partial rows.
template <typename SRC, int NP, int LEVEL>
__global__ void __launch_bounds__(RED_THREADS) reduce_kernel(SRC src, int V, int E, const int* __restrict__ plan,
                                                             const int* __restrict__ seg_beg, const int* __restrict__ seg_end,
                                                             const int* __restrict__ pos_sorted, float* __restrict__ part_out,
                                                             float* __restrict__ d_word_emb, float scale) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * RED_THREADS + threadIdx.x) >> 5;
  const int* wstart = plan + (long long)(LEVEL == 1 ? 0 : LEVEL == 2 ? 2 : 4) * (V + 1);
  if (w >= wstart[V]) return;
  const int s = find_segment(wstart, V, (int)w);
  const int k = (int)w - wstart[s], nparts = wstart[s + 1] - wstart[s];
  long long ibeg, iend;
  if (LEVEL == 1) { ibeg = seg_beg[s]; iend = seg_end[s]; }
  else {            // this token's partial rows of the previous level
    const int* pprev = plan + (long long)(LEVEL == 2 ? 1 : 3) * (V + 1);
    ibeg = pprev[s]; iend = pprev[s + 1];
  }
  long long i0 = ibeg + (long long)k * R, i1 = LEVEL == 3 ? iend : min(iend, i0 + R);
  float acc[NP][4];
#pragma unroll
  for (int j = 0; j < NP; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  for (long long b = i0; b < i1; b += 32) {
    const int cntb = (int)min((long long)32, i1 - b);
    long long mine = b + lane;
    if (LEVEL == 1) mine = lane < cntb ? pos_sorted[b + lane] : 0;
#pragma unroll 4
    for (int j = 0; j < cntb; ++j) {
      const long long r = __shfl_sync(0xffffffffu, mine, j);
      add_row<NP>(src, r, lane, acc);
    }
  }
  float* dst;
  float sc = 1.f;
  if (nparts == 1) { dst = d_word_emb + (long long)s * E; sc = scale; }
  else {
    const int* pcur = plan + (long long)(LEVEL == 1 ? 1 : 3) * (V + 1);
    dst = part_out + ((long long)pcur[s] + k) * E;
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int e = (lane + 32 * j) * 4;
    if (e < E) *reinterpret_cast<float4*>(dst + e) = make_float4(acc[j][0] * sc, acc[j][1] * sc, acc[j][2] * sc, acc[j][3] * sc);
  }
}

struct Ws {          // carve-up of the caller's workspace
  int *keys_a, *keys_b, *pos_a, *pos_b, *ghist, *bintot, *seg_beg, *seg_end, *plan;
  float *part1, *part2;
  size_t bytes;
};
static long long part1_rows(long long n) { return 2 * (n / R) + 2; }               // tokens with > R rows: sum ceil(c/R) <= n/R + n/R
static long long part2_rows(long long n) { return 2 * (part1_rows(n) / R) + 2; }
static Ws carve(void* base, long long n, int V, int E) {
  Ws w;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = (off + b + 255) & ~(size_t)255; return (char*)base + o; };
  const long long nblk = (n + SORT_CH - 1) / SORT_CH;
  w.keys_a = (int*)take(n * 4); w.keys_b = (int*)take(n * 4); w.pos_a = (int*)take(n * 4); w.pos_b = (int*)take(n * 4);
  w.ghist = (int*)take((size_t)RADIX * nblk * 4); w.bintot = (int*)take(RADIX * 4);
  w.seg_beg = (int*)take((size_t)V * 4 * 2); w.seg_end = w.seg_beg + V;       // contiguous: one memset
  w.plan = (int*)take((size_t)PLAN_SEQ * (V + 1) * 4);
  w.part1 = (float*)take((size_t)part1_rows(n) * E * 4);
  w.part2 = (float*)take((size_t)part2_rows(n) * E * 4);
  w.bytes = off;
  return w;
}

template <typename SRC>
static int run(SRC src, long long n, int L, int V, int E, const int* tokens, const int* n_titles_dev, float scale,
               float* d_word_emb, void* workspace, size_t workspace_bytes, cudaStream_t st, const char* name) {
  LSTUR_REQUIRE(n >= 0 && n < (1LL << 31) && L >= 1 && V >= 1 && E >= 4 && E % 4 == 0 && E <= 512 && tokens && d_word_emb, name);
  Ws w = carve(workspace, n, V, E);
  LSTUR_REQUIRE(workspace != nullptr && workspace_bytes >= w.bytes, name);
  cudaMemsetAsync(d_word_emb, 0, (size_t)V * E * sizeof(float), st);
  if (n == 0) return LSTUR_OK;
  keys_kernel<<<cdiv(n, 256), 256, 0, st>>>(n, L, V, tokens, w.keys_a, w.pos_a, n_titles_dev);
  LSTUR_CHECK_LAUNCH(name);
  int bits = 1;
  while ((1LL << bits) <= V) ++bits;              // keys are 0 .. V inclusive
  const int nblk = (int)((n + SORT_CH - 1) / SORT_CH);
  int *ka = w.keys_a, *kb = w.keys_b, *pa = w.pos_a, *pb = w.pos_b;
  for (int shift = 0; shift < bits; shift += RB) {
    radix_hist_kernel<<<nblk, SORT_THREADS, 0, st>>>(ka, n, shift, nblk, w.ghist);
    LSTUR_CHECK_LAUNCH(name);
    radix_rowscan_kernel<<<RADIX, 256, 0, st>>>(w.ghist, nblk, w.bintot);
    LSTUR_CHECK_LAUNCH(name);
    radix_scatter_kernel<<<nblk, SORT_THREADS, 0, st>>>(ka, pa, kb, pb, n, shift, nblk, w.ghist, w.bintot);
    LSTUR_CHECK_LAUNCH(name);
    int* t = ka; ka = kb; kb = t;
    t = pa; pa = pb; pb = t;
  }
  cudaMemsetAsync(w.seg_beg, 0, (size_t)V * 4 * 2, st);
  bounds_kernel<<<cdiv(n, 256), 256, 0, st>>>(n, V, ka, w.seg_beg, w.seg_end);
  LSTUR_CHECK_LAUNCH(name);
  plan_kernel<<<1, 1024, 0, st>>>(V, w.seg_beg, w.seg_end, w.plan);
  LSTUR_CHECK_LAUNCH(name);
  const int wpb = RED_THREADS / 32;
  const long long nw1 = n / R + (V < n ? V : n) + 1, nw2 = part1_rows(n) / R + n / R + 2, nw3 = part1_rows(n) / R + 2;
  const int np = (E / 4 + 31) / 32;               // 4-column pieces per lane
  const SrcPart s1{w.part1, E}, s2{w.part2, E};
#define WG_LEVELS(NP_)                                                                                                    \
  do {                                                                                                                    \
    reduce_kernel<SRC, NP_, 1><<<cdiv(nw1, wpb), RED_THREADS, 0, st>>>(src, V, E, w.plan, w.seg_beg, w.seg_end, pa, w.part1, d_word_emb, scale); \
    LSTUR_CHECK_LAUNCH(name);                                                                                             \
    reduce_kernel<SrcPart, NP_, 2><<<cdiv(nw2, wpb), RED_THREADS, 0, st>>>(s1, V, E, w.plan, w.seg_beg, w.seg_end, nullptr, w.part2, d_word_emb, scale); \
    LSTUR_CHECK_LAUNCH(name);                                                                                             \
    reduce_kernel<SrcPart, NP_, 3><<<cdiv(nw3, wpb), RED_THREADS, 0, st>>>(s2, V, E, w.plan, w.seg_beg, w.seg_end, nullptr, nullptr, d_word_emb, scale); \
    LSTUR_CHECK_LAUNCH(name);                                                                                             \
  } while (0)
  if (np <= 1) WG_LEVELS(1);
  else if (np == 2) WG_LEVELS(2);
  else if (np == 3) WG_LEVELS(3);
  else WG_LEVELS(4);
#undef WG_LEVELS
  return LSTUR_OK;
}

}  // namespace wg
}  // namespace lstur

using namespace lstur;

extern "C" int lstur_tc_padded_e(int E);

extern "C" size_t lstur_word_grad_workspace_bytes(long long n_pos, int V, int E) {
  if (n_pos < 0 || V < 1 || E < 1) return 0;
  return wg::carve(nullptr, n_pos, V, E).bytes;
}

// tensor-core modes: dx16 (n_titles, L, lstur_tc_padded_e(E)) from lstur_conv_dgrad_tc, keep bytes from the forward
extern "C" int lstur_word_grad_scatter_16(int n_titles, int L, int E, int V, const int* tokens, const void* dx16, int fp16,
                                          float scale, const void* xmask, float* d_word_emb, void* workspace,
                                          size_t workspace_bytes, const int* n_titles_dev, cudaStream_t stream) {
  LSTUR_REQUIRE(n_titles >= 0 && dx16 != nullptr, "lstur_word_grad_scatter_16");
  wg::Src16 s{(const uint16_t*)dx16, (const uint8_t*)xmask, lstur_tc_padded_e(E), fp16};
  return wg::run(s, (long long)n_titles * L, L, V, E, tokens, n_titles_dev, scale, d_word_emb, workspace, workspace_bytes,
                 stream, "lstur_word_grad_scatter_16");
}

// fp32 mode: dXp (n_titles, L+KS-1, E) = gradient of the zero-haloed title buffer of lstur_embed_gather_pad; the X-dropout
// of that kernel (seed, dropout) is replayed
extern "C" int lstur_word_grad_scatter_f32(int n_titles, int L, int KS, int E, int V, const int* tokens, const float* dXp,
                                           float dropout, unsigned seed, float scale, float* d_word_emb, void* workspace,
                                           size_t workspace_bytes, cudaStream_t stream) {
  LSTUR_REQUIRE(n_titles >= 0 && dXp != nullptr && KS >= 1 && dropout >= 0.f && dropout < 1.f, "lstur_word_grad_scatter_f32");
  wg::Src32 s{dXp, L, L + KS - 1, (KS - 1) / 2, E, dropout > 0.f ? dropout_threshold(dropout) : 0u, seed};
  return wg::run(s, (long long)n_titles * L, L, V, E, tokens, nullptr, scale / (1.f - dropout), d_word_emb, workspace,
                 workspace_bytes, stream, "lstur_word_grad_scatter_f32");
}
