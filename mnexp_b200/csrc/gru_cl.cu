// Masked Keras-2.2 GRU recurrence with the recurrent weights resident in shared memory.
//
// Reference call sites: keras.layers.GRU(U)(Masking()(clicked), initial_state=user_vec)
// task/paper.py:612-613 (LSTUR-ini) / :596-611 (LSTUR-con); semantics restated in SURVEY.md §9.4.
//
// Decomposition: a thread-block cluster of CS CTAs owns NB = 4*RB batch rows for all W steps.
// CTA `rank` owns the hidden units j in [rank*UC, rank*UC+UC) and keeps, for the whole launch, the
// (G x 3*UC) fp32 slice of Wh that produces them in shared memory (G=200, CS=4: 120 KB) — the
// weights are read from L2 once per launch instead of once per step.  The state h (and r*h) of all
// G units is mirrored in every CTA ([k][32 row slots]); after each phase the owners push their
// slice to all peers with st.async (distributed shared memory + mbarrier complete_tx), so a step
// costs two data-flow waits and no cluster barrier / fence.
//
// Thread mapping (256 threads): warp w = (ks, wq).  wq picks two row groups (half a warp each) and
// half of the unit pairs; a thread accumulates RB rows x 2 units (all gates of the phase) with
// packed FFMA2 over row pairs.  A 128-bit shared load costs two wavefronts however many lanes
// share the address, so fetching the state rows once per half-warp and reusing them for two units
// and two gates keeps the FMA pipe, not the shared-memory pipe, the limiter.  ks splits the
// reduction range in two so that every SM sub-partition has two warps (FFMA2 issues every other
// cycle; the partner fills the gaps); partner warps swap partial sums through shared memory and
// each finishes (activations, stores, pushes) half of the rows.
// Steps at which no row of the cluster is active (left padding) are skipped cluster-uniformly.
#include "common.cuh"

namespace lstur {
namespace grucl {

constexpr int ROWS = 32;       // row slots per cluster tile (4 row groups x 8)
constexpr int THREADS = 256;
constexpr int HALF = THREADS / 2;

__device__ __forceinline__ float rec_act(float x, int act) {
  return act == LSTUR_ACT_HARD_SIGMOID ? hard_sigmoid_f(x) : 1.f / (1.f + expf(-x));
}
__device__ __forceinline__ float rec_act_grad(float y, int act) {
  return act == LSTUR_ACT_HARD_SIGMOID ? ((y > 0.f && y < 1.f) ? 0.2f : 0.f) : y * (1.f - y);
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_remote(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// Remote store that signals the destination CTA's mbarrier when it lands (no fence, no cluster barrier needed).
__device__ __forceinline__ void st_async(uint32_t addr, uint32_t bar, float v) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(addr), "f"(v), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_async4(uint32_t addr, uint32_t bar, float a, float b, float c, float d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(addr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
template <int CS>
__device__ __forceinline__ void cluster_sync() {
  if (CS > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}
// Exchange point: CS == 1 -> CTA barrier; CS > 1 -> wait until every owner's slice has landed in this CTA's copy.
// Ordering between steps needs no further barrier: a buffer is only overwritten by data that causally depends on every
// CTA having finished reading the previous contents (see the kernels).
template <int CS>
__device__ __forceinline__ void exchange_wait(uint32_t bar, uint32_t& parity) {
  if (CS > 1) {
    mbar_wait(bar, parity);
    parity ^= 1u;
  } else {
    __syncthreads();
  }
}
__device__ __forceinline__ void pair_sync(int wq) {   // the two warps (ks = 0, 1) that share wq
  asm volatile("bar.sync %0, 64;" ::"r"(wq + 1) : "memory");
}
// acc (two rows) += x (two rows) * w
__device__ __forceinline__ void fma2(float2& acc, const float2 x, const float2 w) {
  unsigned long long a = *reinterpret_cast<unsigned long long*>(&acc);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(a)
      : "l"(*reinterpret_cast<const unsigned long long*>(&x)), "l"(*reinterpret_cast<const unsigned long long*>(&w)));
  acc = *reinterpret_cast<float2*>(&a);
}

// RB values of one row group at s[0..RB)
template <int RB>
struct RowVec {
  float v[RB];
  __device__ __forceinline__ void load(const float* s) {
    if (RB == 8) {
      float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
      v[0] = a.x; v[1 % RB] = a.y; v[2 % RB] = a.z; v[3 % RB] = a.w;
      v[4 % RB] = b.x; v[5 % RB] = b.y; v[6 % RB] = b.z; v[7 % RB] = b.w;
    } else if (RB == 4) {
      float4 a = *reinterpret_cast<const float4*>(s);
      v[0] = a.x; v[1 % RB] = a.y; v[2 % RB] = a.z; v[3 % RB] = a.w;
    } else if (RB == 2) {
      float2 a = *reinterpret_cast<const float2*>(s);
      v[0] = a.x; v[1 % RB] = a.y;
    } else {
      v[0] = s[0];
    }
  }
};
// acc[i] += x[i] * w for the RB rows of a group (FFMA2 over row pairs when RB is even)
template <int RB>
__device__ __forceinline__ void fma_rows(float* acc, const float* x, float w) {
#ifdef LSTUR_GRU_SCALAR_FMA
  constexpr bool kPacked = false;
#else
  constexpr bool kPacked = (RB % 2 == 0);
#endif
  if (kPacked) {
    const float2 w2 = make_float2(w, w);
#pragma unroll
    for (int q = 0; q < RB / 2; ++q) {
      float2 a = make_float2(acc[2 * q], acc[2 * q + 1]);
      fma2(a, make_float2(x[2 * q], x[2 * q + 1]), w2);
      acc[2 * q] = a.x; acc[2 * q + 1] = a.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < RB; ++i) acc[i] = fmaf(x[i], w, acc[i]);
  }
}
// push the RH finished rows (slots slot0 ..) of unit j into buffer `buf` of every CTA of the cluster
template <int RH, int CS>
__device__ __forceinline__ void push_rows(float* buf, uint32_t bar, int j, int slot0, const float* v) {
  float* dst = buf + j * ROWS + slot0;
  if (CS == 1) {
#pragma unroll
    for (int i = 0; i < RH; ++i) dst[i] = v[i];
    return;
  }
  const uint32_t la = smem_addr(dst);
#pragma unroll
  for (int c = 0; c < CS; ++c) {
    const uint32_t ra = map_remote(la, c), rb = map_remote(bar, c);
    if (RH == 4) {
      st_async4(ra, rb, v[0], v[1 % RH], v[2 % RH], v[3 % RH]);
    } else {
#pragma unroll
      for (int i = 0; i < RH; ++i) st_async(ra + 4 * i, rb, v[i]);
    }
  }
}

struct Params {
  int B, W, G, UC, NUP, UPW, NB, K0, act;
  const float* gm;       // (B, W)
  const int* row_order;  // optional permutation of the batch rows, or null
  // forward
  const float* XW;       // (B, W, 3G)
  const float* h0; long long ldh0;
  const float* Wh;       // (G, 3G)
  float* hT; long long ldo;
  float *Z, *R, *HH, *HP, *RH;   // (B, W, G) each or all null
  // backward
  const float* WhT;      // (3G, G)
  const float* dhT; long long lddh;
  float* dA;             // (B, W, 3G)
  float* dh0; long long lddh0;
};

// shared-memory carve:  w1 [G][NUP] float4 | w2 [G][NUP] float2 | x0,x1(,x2) [G][32] | partial swap 16 KB |
//                       mask bytes [W][32] | any [W] | 2 mbarriers
constexpr int PART_FLOATS = 2 * 4 * HALF * 4;   // [ks][4 float4 groups][128 threads]
__host__ __device__ inline size_t smem_bytes(int G, int NUP, int W, int nbuf) {
  size_t b = (size_t)G * NUP * 24 + (size_t)nbuf * G * ROWS * 4 + (size_t)PART_FLOATS * 4 + (size_t)W * ROWS +
             (size_t)((W + 15) / 16) * 16;
  return b + 16 + 32;
}

struct TileCtx {
  int rank, tile, ks, wq, t128, rg, up, upc, u0, u1, b0, kb, ke;
  bool act0, act1;
};
template <int CS>
__device__ __forceinline__ TileCtx tile_ctx(const Params& p) {
  TileCtx c;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  c.ks = w >> 2;
  c.wq = w & 3;
  c.t128 = threadIdx.x & (HALF - 1);
  c.rank = CS > 1 ? (int)cta_rank() : 0;
  c.tile = blockIdx.x / CS;
  c.rg = 2 * (c.wq >> 1) + (lane >> 4);
  c.up = (c.wq & 1) * p.UPW + (lane & 15);
  const bool pair_ok = (lane & 15) < p.UPW && c.up < p.NUP;
  c.upc = pair_ok ? c.up : 0;
  c.u0 = c.rank * p.UC + 2 * c.up;
  c.u1 = c.u0 + 1;
  c.act0 = pair_ok && 2 * c.up < p.UC && c.u0 < p.G;
  c.act1 = pair_ok && 2 * c.up + 1 < p.UC && c.u1 < p.G;
  c.b0 = c.tile * p.NB;
  c.kb = c.ks == 0 ? 0 : p.K0;
  c.ke = c.ks == 0 ? p.K0 : p.G;
  return c;
}
// batch row held by slot s (= rg*8 + i) of this tile, or -1
template <int RB>
__device__ __forceinline__ int slot_row(const Params& p, int b0, int s) {
  const int r = (s >> 3) * RB + (s & 7);
  if ((s & 7) >= RB || b0 + r >= p.B) return -1;
  return p.row_order ? p.row_order[b0 + r] : b0 + r;
}
template <int RB>
__device__ __forceinline__ void load_masks(const Params& p, int b0, uint8_t* sGM, uint8_t* sAny) {
  for (int i = threadIdx.x; i < p.W * ROWS; i += blockDim.x) {
    const int t = i / ROWS, b = slot_row<RB>(p, b0, i % ROWS);
    sGM[i] = (b >= 0 && p.gm[(long long)b * p.W + t] != 0.f) ? 1 : 0;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < p.W; t += blockDim.x) {
    int any = 0;
    for (int s = 0; s < ROWS; ++s) any |= sGM[t * ROWS + s];
    sAny[t] = (uint8_t)any;
  }
}

// ---- software-pipelined inner products over k in [kb, ke) (two k per stage; the next stage's operands are loaded first)
// two-gate phase: {a0,b0,a1,b1}[row] += x[k][row] * w[k].{x,y,z,w}   (gate A / gate B of unit 0, unit 1), one x buffer
template <int RB>
struct Stage2 {
  float x[2][RB];
  float4 w[2];
  __device__ __forceinline__ void load(const float* xs, const float4* ws, int k, int NUP) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      w[s] = ws[(k + s) * NUP];
      RowVec<RB> v;
      v.load(xs + (k + s) * ROWS);
#pragma unroll
      for (int i = 0; i < RB; ++i) x[s][i] = v.v[i];
    }
  }
  __device__ __forceinline__ void apply(float* a0, float* b0, float* a1, float* b1) const {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      fma_rows<RB>(a0, x[s], w[s].x); fma_rows<RB>(b0, x[s], w[s].y);
      fma_rows<RB>(a1, x[s], w[s].z); fma_rows<RB>(b1, x[s], w[s].w);
    }
  }
};
template <int RB>
__device__ __forceinline__ void dot2(int kb, int ke, int NUP, const float* xs, const float4* ws, float* a0, float* b0,
                                     float* a1, float* b1) {
  if (kb >= ke) return;
  Stage2<RB> p, q;
  p.load(xs, ws, kb, NUP);
  for (int k = kb; k < ke; k += 4) {
    q.load(xs, ws, k + 2, NUP);
    p.apply(a0, b0, a1, b1);
    if (k + 4 < ke) p.load(xs, ws, k + 4, NUP);
    q.apply(a0, b0, a1, b1);
  }
}
// as dot2, but gate A reads buffer xa and gate B reads buffer xb (backward phase A)
template <int RB>
struct Stage2x {
  float xa[2][RB], xb[2][RB];
  float4 w[2];
  __device__ __forceinline__ void load(const float* xas, const float* xbs, const float4* ws, int k, int NUP) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      w[s] = ws[(k + s) * NUP];
      RowVec<RB> v, u;
      v.load(xas + (k + s) * ROWS);
      u.load(xbs + (k + s) * ROWS);
#pragma unroll
      for (int i = 0; i < RB; ++i) { xa[s][i] = v.v[i]; xb[s][i] = u.v[i]; }
    }
  }
  __device__ __forceinline__ void apply(float* a0, float* b0, float* a1, float* b1) const {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      fma_rows<RB>(a0, xa[s], w[s].x); fma_rows<RB>(b0, xb[s], w[s].y);
      fma_rows<RB>(a1, xa[s], w[s].z); fma_rows<RB>(b1, xb[s], w[s].w);
    }
  }
};
template <int RB>
__device__ __forceinline__ void dot2x(int kb, int ke, int NUP, const float* xas, const float* xbs, const float4* ws,
                                      float* a0, float* b0, float* a1, float* b1) {
  if (kb >= ke) return;
  Stage2x<RB> p, q;
  p.load(xas, xbs, ws, kb, NUP);
  for (int k = kb; k < ke; k += 4) {
    q.load(xas, xbs, ws, k + 2, NUP);
    p.apply(a0, b0, a1, b1);
    if (k + 4 < ke) p.load(xas, xbs, ws, k + 4, NUP);
    q.apply(a0, b0, a1, b1);
  }
}
// one-gate phase: {a0,a1}[row] += x[k][row] * w[k].{x,y}
template <int RB>
struct Stage1 {
  float x[2][RB];
  float2 w[2];
  __device__ __forceinline__ void load(const float* xs, const float2* ws, int k, int NUP) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      w[s] = ws[(k + s) * NUP];
      RowVec<RB> v;
      v.load(xs + (k + s) * ROWS);
#pragma unroll
      for (int i = 0; i < RB; ++i) x[s][i] = v.v[i];
    }
  }
  __device__ __forceinline__ void apply(float* a0, float* a1) const {
#pragma unroll
    for (int s = 0; s < 2; ++s) { fma_rows<RB>(a0, x[s], w[s].x); fma_rows<RB>(a1, x[s], w[s].y); }
  }
};
template <int RB>
__device__ __forceinline__ void dot1(int kb, int ke, int NUP, const float* xs, const float2* ws, float* a0, float* a1) {
  if (kb >= ke) return;
  Stage1<RB> p, q;
  p.load(xs, ws, kb, NUP);
  for (int k = kb; k < ke; k += 4) {
    q.load(xs, ws, k + 2, NUP);
    p.apply(a0, a1);
    if (k + 4 < ke) p.load(xs, ws, k + 4, NUP);
    q.apply(a0, a1);
  }
}

// Partner swap of partial sums.  `acc` holds NV arrays of RB rows; this thread finishes rows [r0, r0+RH) and hands the
// partial sums of the other RH rows (o0..) to its partner, receiving the partner's partials for its own rows.
template <int RB, int RH, int NV>
__device__ __forceinline__ void swap_partials(float* sPart, const TileCtx& c, float (&acc)[NV][RB], int r0, int o0) {
  constexpr int N = NV * RH, NQ = (N + 3) / 4;
  static_assert(NQ <= 4, "partial swap buffer too small");
  float send[NQ * 4];
#pragma unroll
  for (int i = 0; i < NQ * 4; ++i) send[i] = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < RH; ++i) send[v * RH + i] = (o0 + i < RB) ? acc[v][(o0 + i) % RB] : 0.f;
  float4* mine = reinterpret_cast<float4*>(sPart) + (size_t)c.ks * 4 * HALF + c.t128;
  const float4* theirs = reinterpret_cast<const float4*>(sPart) + (size_t)(c.ks ^ 1) * 4 * HALF + c.t128;
#pragma unroll
  for (int q = 0; q < NQ; ++q) mine[q * HALF] = make_float4(send[4 * q], send[4 * q + 1], send[4 * q + 2], send[4 * q + 3]);
  pair_sync(c.wq);
  float recv[NQ * 4];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 g = theirs[q * HALF];
    recv[4 * q] = g.x; recv[4 * q + 1] = g.y; recv[4 * q + 2] = g.z; recv[4 * q + 3] = g.w;
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < RH; ++i)
      if (r0 + i < RB) acc[v][(r0 + i) % RB] += recv[v * RH + i];
}

template <int RB, int CS, int KS>
__device__ __forceinline__ void gru_fwd_body(const Params& p, uint8_t* smraw) {
  constexpr int RH = (RB + 1) / 2;
  const int G = p.G, UC = p.UC, NUP = p.NUP, W = p.W, G3 = 3 * p.G;
  float4* sW1 = reinterpret_cast<float4*>(smraw);                              // (wz_u0, wr_u0, wz_u1, wr_u1)
  float2* sW2 = reinterpret_cast<float2*>(smraw + (size_t)G * NUP * 16);       // (wh_u0, wh_u1)
  float* sH = reinterpret_cast<float*>(smraw + (size_t)G * NUP * 24);
  float* sRH = sH + (size_t)G * ROWS;
  float* sPart = sRH + (size_t)G * ROWS;
  uint8_t* sGM = reinterpret_cast<uint8_t*>(sPart + PART_FLOATS);
  uint8_t* sAny = sGM + (size_t)W * ROWS;
  const uint32_t bar_rh = smem_addr(sAny + (size_t)((W + 15) / 16) * 16), bar_h = bar_rh + 8;
  const int tid = threadIdx.x;
  const TileCtx c = tile_ctx<CS>(p);
  constexpr int r0 = KS * RH, o0 = (KS ^ 1) * RH;    // rows this thread finishes / hands to its partner
  const int slot0 = c.rg * 8 + r0;
  const bool save = p.HP != nullptr;
  const uint32_t xbytes = (uint32_t)G * 4u * RB * 4u;   // bytes every CTA receives per exchange

  // ---- one-time loads: weight slice, masks, initial state
  if (CS > 1 && tid == 0) {
    mbar_init(bar_rh, 1);
    mbar_init(bar_h, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < G * NUP * 2; i += blockDim.x) {
    const int k = i / (NUP * 2), r = i % (NUP * 2), jj = c.rank * UC + r;   // r = 2*up + which unit of the pair
    float wz = 0.f, wr = 0.f, wh = 0.f;
    if (r < UC && jj < G) {
      wz = p.Wh[(long long)k * G3 + jj];
      wr = p.Wh[(long long)k * G3 + G + jj];
      wh = p.Wh[(long long)k * G3 + 2 * G + jj];
    }
    float* w1 = reinterpret_cast<float*>(sW1) + ((size_t)k * NUP + (r >> 1)) * 4 + (r & 1) * 2;
    w1[0] = wz; w1[1] = wr;
    reinterpret_cast<float*>(sW2)[((size_t)k * NUP + (r >> 1)) * 2 + (r & 1)] = wh;
  }
  for (int i = tid; i < G * ROWS; i += blockDim.x) {
    const int k = i / ROWS, b = slot_row<RB>(p, c.b0, i % ROWS);
    sH[i] = (b >= 0 && p.h0) ? p.h0[(long long)b * p.ldh0 + k] : 0.f;
    sRH[i] = 0.f;
  }
  load_masks<RB>(p, c.b0, sGM, sAny);
  __syncthreads();
  // per finished row: element offset of (b, t=0, unit 0) in the (B,W,G) tensors, or -1
  long long rowG[RH];
  float h0v[RH], h1v[RH];
#pragma unroll
  for (int i = 0; i < RH; ++i) {
    const int b = (r0 + i < RB) ? slot_row<RB>(p, c.b0, slot0 + i) : -1;
    rowG[i] = b >= 0 ? (long long)b * W * G : -1;
    h0v[i] = (c.act0 && r0 + i < RB) ? sH[c.u0 * ROWS + slot0 + i] : 0.f;
    h1v[i] = (c.act1 && r0 + i < RB) ? sH[c.u1 * ROWS + slot0 + i] : 0.f;
  }
  const bool vec2 = c.act1 && ((c.u0 & 1) == 0) && ((G & 1) == 0);   // the unit pair is 8-byte aligned in every tensor
  cluster_sync<CS>();   // every CTA of the cluster is resident and initialised before the first remote store

  // pre-activations (XW) of the two own units, gate `gate`, at step t for the finished rows
  auto load_x = [&](int t, int gate, float* x0, float* x1) {
#pragma unroll
    for (int i = 0; i < RH; ++i) {
      x0[i] = 0.f; x1[i] = 0.f;
      if (rowG[i] >= 0 && t < W) {
        const float* x = p.XW + (rowG[i] + (long long)t * G) * 3 + gate * G + c.u0;
        if (vec2) {
          const float2 v = __ldg(reinterpret_cast<const float2*>(x));
          x0[i] = v.x; x1[i] = v.y;
        } else {
          if (c.act0) x0[i] = __ldg(x);
          if (c.act1) x1[i] = __ldg(x + 1);
        }
      }
    }
  };
  auto store2 = [&](float* base, long long o, float v0, float v1) {
    if (vec2) {
      *reinterpret_cast<float2*>(base + o) = make_float2(v0, v1);
    } else {
      if (c.act0) base[o] = v0;
      if (c.act1) base[o + 1] = v1;
    }
  };
  float nz0[RH], nz1[RH], nr0[RH], nr1[RH], nh0[RH], nh1[RH];
  load_x(0, 0, nz0, nz1); load_x(0, 1, nr0, nr1); load_x(0, 2, nh0, nh1);
  uint32_t par_rh = 0, par_h = 0;

  for (int t = 0; t < W; ++t) {
    if (!sAny[t]) {   // whole tile masked at this step: carry state (cluster-uniform decision)
      if (save && (c.act0 || c.act1)) {
#pragma unroll
        for (int i = 0; i < RH; ++i)
          if (rowG[i] >= 0) {
            const long long o = rowG[i] + (long long)t * G + c.u0;
            store2(p.Z, o, 0.f, 0.f); store2(p.R, o, 0.f, 0.f); store2(p.HH, o, 0.f, 0.f); store2(p.RH, o, 0.f, 0.f);
            store2(p.HP, o, h0v[i], h1v[i]);
          }
      }
      load_x(t + 1, 0, nz0, nz1); load_x(t + 1, 1, nr0, nr1); load_x(t + 1, 2, nh0, nh1);
      continue;
    }
    if (CS > 1 && tid == 0) {   // arm this step's two exchanges (their previous phases completed last step)
      mbar_arrive_expect(bar_rh, xbytes);
      mbar_arrive_expect(bar_h, xbytes);
    }
    // ---- phase 1: z, r of the own units (partial over this thread's k range; XW seeds the rows finished here)
    float acc[4][RB];   // z0, r0, z1, r1
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int i = 0; i < RB; ++i) acc[v][i] = 0.f;
    dot2<RB>(c.kb, c.ke, NUP, sH + c.rg * 8, sW1 + c.upc, acc[0], acc[1], acc[2], acc[3]);
    swap_partials<RB, RH, 4>(sPart, c, acc, r0, o0);
    // (low-k partial + high-k partial) + XW: the same expression whichever warp set finishes the row, so the result
    // does not depend on where a batch row sits in the tile
    float z0[RH], z1[RH], rr0[RH], rr1[RH], rh0[RH], rh1[RH];
#pragma unroll
    for (int i = 0; i < RH; ++i) {
      const int ii = (r0 + i) % RB;
      z0[i] = rec_act(acc[0][ii] + nz0[i], p.act); rr0[i] = rec_act(acc[1][ii] + nr0[i], p.act);
      z1[i] = rec_act(acc[2][ii] + nz1[i], p.act); rr1[i] = rec_act(acc[3][ii] + nr1[i], p.act);
      rh0[i] = rr0[i] * h0v[i]; rh1[i] = rr1[i] * h1v[i];
    }
    // sRH may be overwritten now: every peer that reaches this point has received all of h(t-1), which each CTA
    // sends only after it finished reading sRH in the previous step's phase 2.
    if (r0 < RB) {
      if (c.act0) push_rows<RH, CS>(sRH, bar_rh, c.u0, slot0, rh0);
      if (c.act1) push_rows<RH, CS>(sRH, bar_rh, c.u1, slot0, rh1);
    }
    // next step's pre-activations: in flight during the exchange and phase 2
    load_x(t + 1, 0, nz0, nz1); load_x(t + 1, 1, nr0, nr1);
    exchange_wait<CS>(bar_rh, par_rh);
    // ---- phase 2: candidate state
    float ah[2][RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) { ah[0][i] = 0.f; ah[1][i] = 0.f; }
    dot1<RB>(c.kb, c.ke, NUP, sRH + c.rg * 8, sW2 + c.upc, ah[0], ah[1]);
    swap_partials<RB, RH, 2>(sPart, c, ah, r0, o0);
    float hn0[RH], hn1[RH];
#pragma unroll
    for (int i = 0; i < RH; ++i) {
      const int ii = (r0 + i) % RB;
      const bool on = (r0 + i < RB) && sGM[t * ROWS + slot0 + i] != 0;
      const float hh0 = tanhf(ah[0][ii] + nh0[i]), hh1 = tanhf(ah[1][ii] + nh1[i]);
      const float hnew0 = z0[i] * h0v[i] + (1.f - z0[i]) * hh0, hnew1 = z1[i] * h1v[i] + (1.f - z1[i]) * hh1;
      if (save && rowG[i] >= 0 && (c.act0 || c.act1)) {
        const long long o = rowG[i] + (long long)t * G + c.u0;
        store2(p.Z, o, on ? z0[i] : 0.f, on ? z1[i] : 0.f);
        store2(p.R, o, on ? rr0[i] : 0.f, on ? rr1[i] : 0.f);
        store2(p.HH, o, on ? hh0 : 0.f, on ? hh1 : 0.f);
        store2(p.RH, o, on ? rh0[i] : 0.f, on ? rh1[i] : 0.f);
        store2(p.HP, o, h0v[i], h1v[i]);
      }
      if (on) { h0v[i] = hnew0; h1v[i] = hnew1; }
      hn0[i] = h0v[i]; hn1[i] = h1v[i];
    }
    load_x(t + 1, 2, nh0, nh1);
    if (CS == 1) __syncthreads();   // single CTA: everyone is done reading sRH before sH / sRH are rewritten
    // sH may be overwritten: all of rh(t) has arrived here, which each CTA sends only after its phase 1 (last read of sH)
    if (r0 < RB) {
      if (c.act0) push_rows<RH, CS>(sH, bar_h, c.u0, slot0, hn0);
      if (c.act1) push_rows<RH, CS>(sH, bar_h, c.u1, slot0, hn1);
    }
    exchange_wait<CS>(bar_h, par_h);
  }
#pragma unroll
  for (int i = 0; i < RH; ++i)
    if (rowG[i] >= 0) {
      const long long b = rowG[i] / ((long long)W * G);
      if (c.act0) p.hT[b * p.ldo + c.u0] = h0v[i];
      if (c.act1) p.hT[b * p.ldo + c.u1] = h1v[i];
    }
  cluster_sync<CS>();   // no CTA retires while a peer's stores into it could still be in flight
}

// BPTT.  A thread owns d h[:, k] of its two units k for the rows it finishes.  Per step: dah, daz (elementwise) are
// pushed to all CTAs; phase A: drh_k = sum_j dah_j Wh[k][2G+j] and dhp_k += sum_j daz_j Wh[k][j]; dar is pushed;
// phase B: dhp_k += sum_j dar_j Wh[k][G+j].  The weight slice is the k-rows of Wh, staged j-major from WhT.
template <int RB, int CS, int KS>
__device__ __forceinline__ void gru_bwd_body(const Params& p, uint8_t* smraw) {
  constexpr int RH = (RB + 1) / 2;
  const int G = p.G, UC = p.UC, NUP = p.NUP, W = p.W;
  float4* sW1 = reinterpret_cast<float4*>(smraw);                          // (Wh[k0][2G+j], Wh[k0][j], Wh[k1][2G+j], Wh[k1][j])
  float2* sW2 = reinterpret_cast<float2*>(smraw + (size_t)G * NUP * 16);   // (Wh[k0][G+j], Wh[k1][G+j])
  float* sDAH = reinterpret_cast<float*>(smraw + (size_t)G * NUP * 24);
  float* sDAZ = sDAH + (size_t)G * ROWS;
  float* sDAR = sDAZ + (size_t)G * ROWS;
  float* sPart = sDAR + (size_t)G * ROWS;
  uint8_t* sGM = reinterpret_cast<uint8_t*>(sPart + PART_FLOATS);
  uint8_t* sAny = sGM + (size_t)W * ROWS;
  const uint32_t bar_a = smem_addr(sAny + (size_t)((W + 15) / 16) * 16), bar_b = bar_a + 8;
  const int tid = threadIdx.x;
  const TileCtx c = tile_ctx<CS>(p);
  constexpr int r0 = KS * RH, o0 = (KS ^ 1) * RH;
  const int slot0 = c.rg * 8 + r0;
  const uint32_t xbytes = (uint32_t)G * 4u * RB * 4u;

  if (CS > 1 && tid == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < G * NUP * 2; i += blockDim.x) {
    const int jj = i / (NUP * 2), r = i % (NUP * 2), kk = c.rank * UC + r;
    float wh = 0.f, wz = 0.f, wr = 0.f;
    if (r < UC && kk < G) {
      wz = p.WhT[(long long)jj * G + kk];
      wr = p.WhT[(long long)(G + jj) * G + kk];
      wh = p.WhT[(long long)(2 * G + jj) * G + kk];
    }
    float* w1 = reinterpret_cast<float*>(sW1) + ((size_t)jj * NUP + (r >> 1)) * 4 + (r & 1) * 2;
    w1[0] = wh; w1[1] = wz;
    reinterpret_cast<float*>(sW2)[((size_t)jj * NUP + (r >> 1)) * 2 + (r & 1)] = wr;
  }
  // unused row slots of the exchange buffers must hold finite values (they are multiplied, never stored)
  for (int i = tid; i < 3 * G * ROWS; i += blockDim.x) sDAH[i] = 0.f;
  load_masks<RB>(p, c.b0, sGM, sAny);
  long long rowG[RH];
  float dh0v[RH], dh1v[RH];
#pragma unroll
  for (int i = 0; i < RH; ++i) {
    const int b = (r0 + i < RB) ? slot_row<RB>(p, c.b0, slot0 + i) : -1;
    rowG[i] = b >= 0 ? (long long)b * W * G : -1;
    dh0v[i] = (c.act0 && b >= 0) ? p.dhT[(long long)b * p.lddh + c.u0] : 0.f;
    dh1v[i] = (c.act1 && b >= 0) ? p.dhT[(long long)b * p.lddh + c.u1] : 0.f;
  }
  const bool vec2 = c.act1 && ((c.u0 & 1) == 0) && ((G & 1) == 0);
  cluster_sync<CS>();

  // saved forward tensor `src` of the two own units at step t for the finished rows (zeros on masked rows)
  auto load_saved = [&](const float* src, int t, float* v0, float* v1) {
#pragma unroll
    for (int i = 0; i < RH; ++i) {
      v0[i] = 0.f; v1[i] = 0.f;
      if (t >= 0 && rowG[i] >= 0 && sGM[t * ROWS + slot0 + i] != 0) {
        const float* q = src + rowG[i] + (long long)t * G + c.u0;
        if (vec2) {
          const float2 v = __ldg(reinterpret_cast<const float2*>(q));
          v0[i] = v.x; v1[i] = v.y;
        } else {
          if (c.act0) v0[i] = __ldg(q);
          if (c.act1) v1[i] = __ldg(q + 1);
        }
      }
    }
  };
  auto store2 = [&](float* base, long long o, float v0, float v1) {
    if (vec2) {
      *reinterpret_cast<float2*>(base + o) = make_float2(v0, v1);
    } else {
      if (c.act0) base[o] = v0;
      if (c.act1) base[o + 1] = v1;
    }
  };
  float nz0[RH], nz1[RH], nr0[RH], nr1[RH], nhh0[RH], nhh1[RH], nhp0[RH], nhp1[RH];
  load_saved(p.Z, W - 1, nz0, nz1); load_saved(p.R, W - 1, nr0, nr1);
  load_saved(p.HH, W - 1, nhh0, nhh1); load_saved(p.HP, W - 1, nhp0, nhp1);
  uint32_t par_a = 0, par_b = 0;

  for (int t = W - 1; t >= 0; --t) {
    if (!sAny[t]) {
      if (c.act0 || c.act1) {
#pragma unroll
        for (int i = 0; i < RH; ++i)
          if (rowG[i] >= 0) {
            const long long o = (rowG[i] + (long long)t * G) * 3 + c.u0;
            store2(p.dA, o, 0.f, 0.f); store2(p.dA, o + G, 0.f, 0.f); store2(p.dA, o + 2 * G, 0.f, 0.f);
          }
      }
      load_saved(p.Z, t - 1, nz0, nz1); load_saved(p.R, t - 1, nr0, nr1);
      load_saved(p.HH, t - 1, nhh0, nhh1); load_saved(p.HP, t - 1, nhp0, nhp1);
      continue;
    }
    if (CS > 1 && tid == 0) {
      mbar_arrive_expect(bar_a, 2 * xbytes);
      mbar_arrive_expect(bar_b, xbytes);
    }
    float r0v[RH], r1v[RH], hp0[RH], hp1[RH], dah0[RH], dah1[RH], daz0[RH], daz1[RH], dz0[RH], dz1[RH];
    float acc[4][RB];   // drh0, dhp0, drh1, dhp1 (partials)
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int i = 0; i < RB; ++i) acc[v][i] = 0.f;
#pragma unroll
    for (int i = 0; i < RH; ++i) {
      const bool on = (r0 + i < RB) && sGM[t * ROWS + slot0 + i] != 0;
      r0v[i] = nr0[i]; r1v[i] = nr1[i]; hp0[i] = nhp0[i]; hp1[i] = nhp1[i];
      const float za = nz0[i], zb = nz1[i], hha = nhh0[i], hhb = nhh1[i];
      dah0[i] = on ? dh0v[i] * (1.f - za) * (1.f - hha * hha) : 0.f;
      dah1[i] = on ? dh1v[i] * (1.f - zb) * (1.f - hhb * hhb) : 0.f;
      daz0[i] = on ? dh0v[i] * (hp0[i] - hha) * rec_act_grad(za, p.act) : 0.f;
      daz1[i] = on ? dh1v[i] * (hp1[i] - hhb) * rec_act_grad(zb, p.act) : 0.f;
      dz0[i] = dh0v[i] * za; dz1[i] = dh1v[i] * zb;
    }
    // sDAH / sDAZ may be overwritten: all of dar(t+1) has arrived here, which each CTA sends after its phase A
    if (r0 < RB) {
      if (c.act0) { push_rows<RH, CS>(sDAH, bar_a, c.u0, slot0, dah0); push_rows<RH, CS>(sDAZ, bar_a, c.u0, slot0, daz0); }
      if (c.act1) { push_rows<RH, CS>(sDAH, bar_a, c.u1, slot0, dah1); push_rows<RH, CS>(sDAZ, bar_a, c.u1, slot0, daz1); }
    }
    if (c.act0 || c.act1) {
#pragma unroll
      for (int i = 0; i < RH; ++i)
        if (rowG[i] >= 0) {
          const long long o = (rowG[i] + (long long)t * G) * 3 + c.u0;
          store2(p.dA, o, daz0[i], daz1[i]);
          store2(p.dA, o + 2 * G, dah0[i], dah1[i]);
        }
    }
    load_saved(p.Z, t - 1, nz0, nz1); load_saved(p.HH, t - 1, nhh0, nhh1);   // next step, in flight under phase A
    exchange_wait<CS>(bar_a, par_a);
    dot2x<RB>(c.kb, c.ke, NUP, sDAH + c.rg * 8, sDAZ + c.rg * 8, sW1 + c.upc, acc[0], acc[1], acc[2], acc[3]);
    swap_partials<RB, RH, 4>(sPart, c, acc, r0, o0);
    float dar0[RH], dar1[RH], dhpa0[RH], dhpa1[RH];
    float dhp[2][RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) { dhp[0][i] = 0.f; dhp[1][i] = 0.f; }
#pragma unroll
    for (int i = 0; i < RH; ++i) {
      const int ii = (r0 + i) % RB;
      const bool on = (r0 + i < RB) && sGM[t * ROWS + slot0 + i] != 0;
      const float drh0 = acc[0][ii], drh1 = acc[2][ii];
      // placement-independent order: (partial + partial) + seed, as in the forward kernel
      dhpa0[i] = fmaf(drh0, r0v[i], acc[1][ii] + dz0[i]);
      dhpa1[i] = fmaf(drh1, r1v[i], acc[3][ii] + dz1[i]);
      dar0[i] = on ? drh0 * hp0[i] * rec_act_grad(r0v[i], p.act) : 0.f;
      dar1[i] = on ? drh1 * hp1[i] * rec_act_grad(r1v[i], p.act) : 0.f;
    }
    // sDAR may be overwritten: all of dah/daz(t) has arrived here, which each CTA sends after its phase B of step t+1
    if (r0 < RB) {
      if (c.act0) push_rows<RH, CS>(sDAR, bar_b, c.u0, slot0, dar0);
      if (c.act1) push_rows<RH, CS>(sDAR, bar_b, c.u1, slot0, dar1);
    }
    if (c.act0 || c.act1) {
#pragma unroll
      for (int i = 0; i < RH; ++i)
        if (rowG[i] >= 0) store2(p.dA, (rowG[i] + (long long)t * G) * 3 + G + c.u0, dar0[i], dar1[i]);
    }
    load_saved(p.R, t - 1, nr0, nr1); load_saved(p.HP, t - 1, nhp0, nhp1);
    exchange_wait<CS>(bar_b, par_b);
    dot1<RB>(c.kb, c.ke, NUP, sDAR + c.rg * 8, sW2 + c.upc, dhp[0], dhp[1]);
    swap_partials<RB, RH, 2>(sPart, c, dhp, r0, o0);
#pragma unroll
    for (int i = 0; i < RH; ++i)
      if ((r0 + i < RB) && sGM[t * ROWS + slot0 + i] != 0) {
        dh0v[i] = dhp[0][(r0 + i) % RB] + dhpa0[i];
        dh1v[i] = dhp[1][(r0 + i) % RB] + dhpa1[i];
      }
    if (CS == 1) __syncthreads();   // single CTA: phase B reads of sDAR done before the next step rewrites the buffers
  }
#pragma unroll
  for (int i = 0; i < RH; ++i)
    if (rowG[i] >= 0) {
      const long long b = rowG[i] / ((long long)W * G);
      if (c.act0) p.dh0[b * p.lddh0 + c.u0] = dh0v[i];
      if (c.act1) p.dh0[b * p.lddh0 + c.u1] = dh1v[i];
    }
  cluster_sync<CS>();   // no CTA retires while a peer's stores into it could still be in flight
}

// The two k-split warp sets run the same body with their row split fixed at compile time (all register indexing static).
template <int RB, int CS>
__global__ void __launch_bounds__(THREADS, 1) gru_fwd_cl_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smraw[];
  if (threadIdx.x < HALF) gru_fwd_body<RB, CS, 0>(p, smraw);
  else gru_fwd_body<RB, CS, 1>(p, smraw);
}
template <int RB, int CS>
__global__ void __launch_bounds__(THREADS, 1) gru_bwd_cl_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smraw[];
  if (threadIdx.x < HALF) gru_bwd_body<RB, CS, 0>(p, smraw);
  else gru_bwd_body<RB, CS, 1>(p, smraw);
}

struct Geometry {
  int CS, UC, NUP, UPW, RB, NB, K0, tiles;
  size_t smem;
  bool ok;
};

static Geometry pick_geometry(int B, int W, int G, int nbuf) {
  Geometry g = {};
  g.ok = false;
  if (G % 4 != 0) return g;
  for (int cs = 1; cs <= 8; cs *= 2) {
    const int uc = (G + cs - 1) / cs, nup = (uc + 1) / 2;
    if (uc > 64) continue;
    const size_t sm = smem_bytes(G, nup, W, nbuf);
    if (sm > 226 * 1024) continue;
    g.CS = cs; g.UC = uc; g.NUP = nup; g.UPW = (nup + 1) / 2; g.smem = sm; g.ok = true;
    break;
  }
  if (!g.ok) return g;
  g.K0 = ((G / 4 + 1) / 2) * 4;   // reduction range split between the two warp sets, both halves multiples of 4
  // clusters resident at once (B300_MICROARCH: 148 CTAs for cluster size <= 2, 132 for 4); keep one wave when possible
  const int max_tiles = g.CS <= 2 ? 148 / g.CS : (g.CS == 4 ? 33 : 14);
  g.RB = 8;
  for (int rb = 1; rb <= 8; rb *= 2)
    if ((B + 4 * rb - 1) / (4 * rb) <= max_tiles) { g.RB = rb; break; }
  g.NB = 4 * g.RB;
  g.tiles = (B + g.NB - 1) / g.NB;
  return g;
}

template <typename K>
static int launch(K kernel, const Geometry& g, const Params& p, cudaStream_t stream, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) {
    set_error("%s: cannot opt in to %zu B of shared memory: %s", name, g.smem, cudaGetErrorString(e));
    return LSTUR_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.tiles * g.CS);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kernel, p);
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", name, cudaGetErrorString(e));
    return LSTUR_ERR_CUDA;
  }
  return LSTUR_OK;
}

#define GRUCL_DISPATCH(KERNEL, g, p, stream, name)                                          \
  do {                                                                                      \
    int rc_ = LSTUR_ERR_UNSUPPORTED;                                                        \
    switch ((g).CS * 16 + (g).RB) {                                                         \
      case 1 * 16 + 1: rc_ = launch(KERNEL<1, 1>, g, p, stream, name); break;               \
      case 1 * 16 + 2: rc_ = launch(KERNEL<2, 1>, g, p, stream, name); break;               \
      case 1 * 16 + 4: rc_ = launch(KERNEL<4, 1>, g, p, stream, name); break;               \
      case 1 * 16 + 8: rc_ = launch(KERNEL<8, 1>, g, p, stream, name); break;               \
      case 2 * 16 + 1: rc_ = launch(KERNEL<1, 2>, g, p, stream, name); break;               \
      case 2 * 16 + 2: rc_ = launch(KERNEL<2, 2>, g, p, stream, name); break;               \
      case 2 * 16 + 4: rc_ = launch(KERNEL<4, 2>, g, p, stream, name); break;               \
      case 2 * 16 + 8: rc_ = launch(KERNEL<8, 2>, g, p, stream, name); break;               \
      case 4 * 16 + 1: rc_ = launch(KERNEL<1, 4>, g, p, stream, name); break;               \
      case 4 * 16 + 2: rc_ = launch(KERNEL<2, 4>, g, p, stream, name); break;               \
      case 4 * 16 + 4: rc_ = launch(KERNEL<4, 4>, g, p, stream, name); break;               \
      case 4 * 16 + 8: rc_ = launch(KERNEL<8, 4>, g, p, stream, name); break;               \
      case 8 * 16 + 1: rc_ = launch(KERNEL<1, 8>, g, p, stream, name); break;               \
      case 8 * 16 + 2: rc_ = launch(KERNEL<2, 8>, g, p, stream, name); break;               \
      case 8 * 16 + 4: rc_ = launch(KERNEL<4, 8>, g, p, stream, name); break;               \
      case 8 * 16 + 8: rc_ = launch(KERNEL<8, 8>, g, p, stream, name); break;               \
    }                                                                                       \
    if (rc_) return rc_;                                                                    \
  } while (0)

static void fill_params(Params& p, const Geometry& g, int B, int W, int G, int rec_act) {
  p.B = B; p.W = W; p.G = G; p.UC = g.UC; p.NUP = g.NUP; p.UPW = g.UPW; p.NB = g.NB; p.K0 = g.K0; p.act = rec_act;
}

}  // namespace grucl
}  // namespace lstur

using namespace lstur;

// 1 if the shared-memory-resident cluster kernels cover this shape (otherwise the streaming kernels of gru.cu run).
extern "C" int lstur_gru_cluster_supported(int B, int W, int G) {
  return B > 0 && W > 0 && G > 0 && grucl::pick_geometry(B, W, G, 3).ok ? 1 : 0;
}

extern "C" int lstur_gru_fwd_cluster(int B, int W, int G, const float* XW, const float* gm, const float* h0,
                                     long long ldh0, const float* Wh, int rec_act, float* hT, long long ldo, float* Z,
                                     float* R, float* HH, float* HP, float* RH, const int* row_order,
                                     cudaStream_t stream) {
  LSTUR_REQUIRE(B >= 0 && W > 0 && G > 0 && XW && gm && Wh && hT, "lstur_gru_fwd_cluster");
  LSTUR_REQUIRE((Z && R && HH && HP && RH) || (!Z && !R && !HH && !HP && !RH), "lstur_gru_fwd_cluster");
  if (B == 0) return LSTUR_OK;
  grucl::Geometry g = grucl::pick_geometry(B, W, G, 2);
  LSTUR_REQUIRE(g.ok, "lstur_gru_fwd_cluster(shape)");
  grucl::Params p = {};
  grucl::fill_params(p, g, B, W, G, rec_act);
  p.gm = gm; p.row_order = row_order; p.XW = XW; p.h0 = h0; p.ldh0 = ldh0; p.Wh = Wh; p.hT = hT; p.ldo = ldo;
  p.Z = Z; p.R = R; p.HH = HH; p.HP = HP; p.RH = RH;
  GRUCL_DISPATCH(grucl::gru_fwd_cl_kernel, g, p, stream, "lstur_gru_fwd_cluster");
  LSTUR_CHECK_LAUNCH("lstur_gru_fwd_cluster");
  return LSTUR_OK;
}

extern "C" int lstur_gru_bwd_cluster(int B, int W, int G, const float* gm, const float* Z, const float* R,
                                     const float* HH, const float* HP, const float* WhT, int rec_act, const float* dhT,
                                     long long lddh, float* dA, float* dh0, long long lddh0, const int* row_order,
                                     cudaStream_t stream) {
  LSTUR_REQUIRE(B >= 0 && W > 0 && G > 0 && gm && Z && R && HH && HP && WhT && dhT && dA && dh0, "lstur_gru_bwd_cluster");
  if (B == 0) return LSTUR_OK;
  grucl::Geometry g = grucl::pick_geometry(B, W, G, 3);
  LSTUR_REQUIRE(g.ok, "lstur_gru_bwd_cluster(shape)");
  grucl::Params p = {};
  grucl::fill_params(p, g, B, W, G, rec_act);
  p.gm = gm; p.row_order = row_order; p.Z = (float*)Z; p.R = (float*)R; p.HH = (float*)HH; p.HP = (float*)HP;
  p.WhT = WhT; p.dhT = dhT; p.lddh = lddh; p.dA = dA; p.dh0 = dh0; p.lddh0 = lddh0;
  GRUCL_DISPATCH(grucl::gru_bwd_cl_kernel, g, p, stream, "lstur_gru_bwd_cluster");
  LSTUR_CHECK_LAUNCH("lstur_gru_bwd_cluster");
  return LSTUR_OK;
}
