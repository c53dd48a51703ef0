// Masked Keras-2.2 GRU recurrence on the 5th-gen tensor cores (tensor-core precision modes only).
//
// Reference call sites: keras.layers.GRU(U)(Masking()(clicked), initial_state=user_vec), task/paper.py:596-613;
// semantics SURVEY.md §9.4 (same as gru.cu / gru_cl.cu, which stay the fp32 verification path).
//
// The FFMA kernels of gru_cl.cu sit on the FP32 pipe.  Here a step is two small UMMAs per CTA:
//     D[gate column m][batch row n] = sum_k A[m][k] * S[n][k]
//   A : the CTA's slice of Wh^T.  It is static for the whole launch, so it lives in TENSOR MEMORY (the A operand of
//       tcgen05.mma may come from TMEM): 128 lanes x (K/2) 32-bit columns per matrix, fp16 hi and lo parts.  The
//       recurrent weights are read from shared/global memory exactly once per launch.
//   S : the state h (phase 1) or r*h (phase 2) of all G units for the cluster's 32 batch rows, MN-major SWIZZLE_64B in
//       shared memory (one 64-byte row per unit k = the 32 rows of that unit), fp16 hi + lo
//   3-term split  A_hi.S_hi + A_lo.S_hi + A_hi.S_lo  (fp32 accumulate in TMEM): ~2^-21 relative, i.e. fp32-like.
// A cluster of 4 CTAs owns 32 batch rows; CTA `rank` owns UC = G/4 units: its phase-1 matrix holds the z gates in rows
// 0..UC-1 and the r gates in rows 64..64+UC-1 (M = 128), its phase-2 matrix the candidate gates in rows 0..UC-1.
// Eight epilogue warps read the accumulators (warp w: TMEM lane quarter w&3, batch rows 16*(w>>2)..+15): threads of
// quarters 0-1 own z_u, hh_u and the state h_u of unit u in registers, quarters 2-3 own r_u.  After each phase the
// owners write their unit's fp16 hi/lo row into the CTA's own S buffer and one thread bulk-copies the CTA's contiguous
// slice into the three peers (cp.async.bulk shared::cta -> shared::cluster, completing on the peer's mbarrier), so a
// step costs two data-flow waits and no cluster barrier.
// TMEM: 2 accumulators x 32 + 4 x (K/2) columns <= 512  =>  G <= 224.
#include "tc_common.cuh"

namespace lstur {
namespace grutc {

using namespace lstur::tc;

constexpr int NROWS = 32;            // batch rows per cluster (= UMMA N)
constexpr int HROWS = 16;            // batch rows per epilogue thread
constexpr int EPI_THREADS = 256;     // warps 0-7: epilogue (lane quarter = warp & 3, row half = warp >> 2)
constexpr int THREADS = 288;         // + warp 8: MMA issuer / TMEM owner
constexpr int TMEM_COLS = 512;

// exp-based forms on the special-function unit (ex2.approx + rcp.approx: ~1e-6 absolute on outputs in [-1, 1])
__device__ __forceinline__ float rec_act(float x, int act) {
  return act == LSTUR_ACT_HARD_SIGMOID ? hard_sigmoid_f(x) : __fdividef(1.f, 1.f + __expf(-x));
}
__device__ __forceinline__ float rec_act_grad(float y, int act) {
  return act == LSTUR_ACT_HARD_SIGMOID ? ((y > 0.f && y < 1.f) ? 0.2f : 0.f) : y * (1.f - y);
}
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(-2.f * fabsf(x));                // in (0, 1]: no overflow
  return copysignf(__fdividef(1.f - e, 1.f + e), x);
}
// MN-major SWIZZLE_64B (S): one 64-byte row (32 MN elements) per k, 8-k atoms of 512 B; a single MN group
__device__ __forceinline__ uint64_t desc_mn64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(512 >> 4) << 16;       // LBO (stride between MN groups; only one group is used)
  d |= (uint64_t)(512 >> 4) << 32;       // SBO: 8 k rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// shared::cta -> a peer's shared memory, completing on the peer's mbarrier
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]^T
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

#ifdef LSTUR_GRUTC_PROF
#define PROF_DECL long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pc = clock64();
#define PROF(i) { const long long now_ = clock64(); pt[i] += now_ - pc; pc = now_; }
#define PROF_PRINT(tag, cond) if ((cond) && blockIdx.x == 4) printf("%s: %lld %lld %lld %lld %lld %lld %lld %lld\n", tag, pt[0], pt[1], pt[2], pt[3], pt[4], pt[5], pt[6], pt[7]);
#else
#define PROF_DECL
#define PROF(i)
#define PROF_PRINT(tag, cond)
#endif

struct Params {
  int B, W, G, UC, act;
  const float* gm;        // (B, W)
  const int* row_order;   // optional permutation of the batch rows, or null
  const float* XW;        // (B, W, 3G)
  const float* h0; long long ldh0;
  const float* Wh;        // (G, 3G)
  float* hT; long long ldo;
  float *Z, *R, *HH, *HP, *RH;   // (B, W, G) each or all null
};

// shared memory (bytes): S_h hi|lo, S_rh hi|lo [KS16*16][64] | h fp32 [64][32] | masks [W][32] | any [W] | barriers, rows
__host__ __device__ inline size_t smem_bytes(int G, int W) {
  const int KS16 = (G + 15) / 16;
  return (size_t)4 * KS16 * 1024 + (size_t)64 * NROWS * 4 + (size_t)W * NROWS + (size_t)((W + 15) / 16) * 16 + 256 + 1024;
}

__global__ void __launch_bounds__(THREADS, 1) gru_fwd_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  const int G = p.G, UC = p.UC, W = p.W, G3 = 3 * G;
  const int ksteps = (G + 15) / 16, KC = ksteps * 8;           // TMEM columns per A matrix part
  const uint32_t s_bytes = (uint32_t)ksteps * 1024;
  const uint32_t oH = 4 * s_bytes;                              // fp32 state exchange [64 units][32 rows]
  const uint32_t oGM = oH + 64 * NROWS * 4;
  const uint32_t oAny = oGM + (uint32_t)W * NROWS;
  const uint32_t oBar = (oAny + (uint32_t)((W + 15) / 16) * 16 + 15u) & ~15u;
  float* sH = reinterpret_cast<float*>(sm + oH);
  uint8_t* sGM = sm + oGM;
  uint8_t* sAny = sm + oAny;
  const uint32_t bar_h = smem_base + oBar, bar_rh = bar_h + 8, bar_d1 = bar_h + 16, bar_d2 = bar_h + 24;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + oBar + 32);
  int* sRow = reinterpret_cast<int*>(sm + oBar + 64);          // batch index of each of the 32 rows, or -1
  const uint32_t S_h = smem_base, S_rh = S_h + 2 * s_bytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  constexpr int CS = 4;
  const int tile = blockIdx.x / CS, b0 = tile * NROWS;
  const bool save = p.HP != nullptr;
  // exchange: a CTA writes the rows of its own units into its own S buffer, then bulk-copies that contiguous slice
  // (hi and lo) into the three peers; it expects the slices of the other three CTAs
  const int own_units = max(0, min(UC, G - rank * UC));
  const uint32_t own_off = (uint32_t)(rank * UC) * 64u, own_bytes = (uint32_t)own_units * 64u;
  const uint32_t xbytes = (uint32_t)(G - own_units) * 64u * 2u;

  // ---- one-time setup
  if (tid == 0) {
    mbar_init(bar_h, 1);
    mbar_init(bar_rh, 1);
    mbar_init(bar_d1, 1);
    mbar_init(bar_d2, 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the state buffers (k rows >= G must stay zero)
  for (uint32_t i = tid; i < oGM / 16; i += THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < W * NROWS; i += THREADS) {
    const int t = i / NROWS, n = i % NROWS;
    int b = -1;
    if (b0 + n < p.B) b = p.row_order ? p.row_order[b0 + n] : b0 + n;
    sGM[i] = (b >= 0 && p.gm[(long long)b * W + t] != 0.f) ? 1 : 0;
    if (t == 0) sRow[n] = b;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // (independent accumulators per split term were measured: no faster — the MMAs are paced at ~40 cycles each whatever
  // their dependence — so each phase accumulates its 3 x ksteps MMAs into one 32-column accumulator)
  const uint32_t tD1 = tmem_base, tD2 = tmem_base + 32, tA1 = tmem_base + 64, tA2 = tA1 + 2 * KC;   // hi at tA, lo at tA + KC
  for (int t = tid; t < W; t += THREADS) {
    int any = 0;
    for (int n = 0; n < NROWS; ++n) any |= sGM[t * NROWS + n];
    sAny[t] = (uint8_t)any;
  }
  if (warp < 8) {
    // recurrent weights -> tensor memory.  Thread (quarter q, lane l) owns matrix row m = 32q + l; the two warps of a
    // quarter split the k steps.  Row m of a K-major fp16 A operand in TMEM: lane m, column k/2 = (k even | k odd << 16).
    const int m = (warp & 3) * 32 + lane, u = m & 63, j = rank * UC + u;
    const bool rowact = u < UC && j < G;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    for (int ks = warp >> 2; ks < ksteps; ks += 2) {
      uint32_t hi1[8], lo1[8], hi2[8], lo2[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float w1[2] = {0.f, 0.f}, w2[2] = {0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = ks * 16 + 2 * q + e;
          if (rowact && k < G) {
            w1[e] = __ldg(p.Wh + (long long)k * G3 + (m < 64 ? 0 : G) + j);      // z gate (rows 0..) / r gate (rows 64..)
            if (m < 64) w2[e] = __ldg(p.Wh + (long long)k * G3 + 2 * G + j);     // candidate gate
          }
        }
        const __half2 h1 = __floats2half2_rn(w1[0], w1[1]), h2 = __floats2half2_rn(w2[0], w2[1]);
        const float2 f1 = __half22float2(h1), f2 = __half22float2(h2);
        const __half2 l1 = __floats2half2_rn(w1[0] - f1.x, w1[1] - f1.y), l2 = __floats2half2_rn(w2[0] - f2.x, w2[1] - f2.y);
        hi1[q] = *reinterpret_cast<const uint32_t*>(&h1); lo1[q] = *reinterpret_cast<const uint32_t*>(&l1);
        hi2[q] = *reinterpret_cast<const uint32_t*>(&h2); lo2[q] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      tmem_st8(tA1 + lane_addr + 8 * ks, hi1);
      tmem_st8(tA1 + KC + lane_addr + 8 * ks, lo1);
      tmem_st8(tA2 + lane_addr + 8 * ks, hi2);
      tmem_st8(tA2 + KC + lane_addr + 8 * ks, lo2);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  fence_proxy_async();                 // the zeroed S buffers are visible to the tensor core
  cluster_sync_all();                  // every CTA is resident and initialised before the first remote copy

  // one thread, after the writers' barrier: local arrival + the slice to every peer
  auto send_slice = [&](uint32_t S, uint32_t bar) {
    mbar_arrive_expect(bar, xbytes);
    if (own_bytes) {
#pragma unroll
      for (int c = 1; c < CS; ++c) {
        const uint32_t peer = (uint32_t)((rank + c) % CS);
        const uint32_t rbar = map_to_cta(bar, peer);
        bulk_s2peer(map_to_cta(S + own_off, peer), S + own_off, own_bytes, rbar);
        bulk_s2peer(map_to_cta(S + s_bytes + own_off, peer), S + s_bytes + own_off, own_bytes, rbar);
      }
    }
  };

  if (warp == 8) {
    // ===================== MMA issuer (the whole warp runs the loop; one elected lane issues) =====================
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc(128, NROWS, true) | (1u << 16);     // A K-major (TMEM), B (= S) MN-major
      uint32_t ph_h = 0, ph_rh = 0;
      PROF_DECL
      for (int t = 0; t < W; ++t) {
        if (!sAny[t]) continue;
        // ---- phase 1: z, r pre-activations from h
        PROF(4)
        mbar_wait(bar_h, ph_h, 31);
        PROF(0)
        ph_h ^= 1;
        tc_fence_after();
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t s_hi = desc_mn64(S_h + (uint32_t)ks * 1024), s_lo = desc_mn64(S_h + s_bytes + (uint32_t)ks * 1024);
          if (leader) {
            umma_ts(tD1, tA1 + 8 * ks, s_hi, idesc, ks > 0);
            umma_ts(tD1, tA1 + KC + 8 * ks, s_hi, idesc, 1);
            umma_ts(tD1, tA1 + 8 * ks, s_lo, idesc, 1);
          }
        }
        if (leader) umma_commit(bar_d1);
        __syncwarp();
        PROF(1)
        // ---- phase 2: candidate pre-activation from r*h
        mbar_wait(bar_rh, ph_rh, 32);
        PROF(2)
        ph_rh ^= 1;
        tc_fence_after();
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t s_hi = desc_mn64(S_rh + (uint32_t)ks * 1024), s_lo = desc_mn64(S_rh + s_bytes + (uint32_t)ks * 1024);
          if (leader) {
            umma_ts(tD2, tA2 + 8 * ks, s_hi, idesc, ks > 0);
            umma_ts(tD2, tA2 + KC + 8 * ks, s_hi, idesc, 1);
            umma_ts(tD2, tA2 + 8 * ks, s_lo, idesc, 1);
          }
        }
        if (leader) umma_commit(bar_d2);
        __syncwarp();
        PROF(3)
      }
      PROF_PRINT("mma  wait_h issue1 wait_rh issue2 loop", lane == 0)
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3, half = warp >> 2;
    const int m = quarter * 32 + lane;          // accumulator row
    const bool is_z = m < 64;                   // z / candidate / state thread of unit u; else r thread of unit u
    const int u = m & 63;
    const int j = rank * UC + u;
    const bool act = u < UC && j < G;
    const int n0 = half * HROWS;                // first of this thread's 16 batch rows
    const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
    // write this thread's 16 values of unit j into the local S buffer (hi at S, lo at S + s_bytes): two 16-byte chunks
    auto put_unit = [&](uint32_t S, const float* v) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const __half2 h2 = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        const float2 hf = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(v[2 * q] - hf.x, v[2 * q + 1] - hf.y);
        hi[q] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      uint8_t* row = sm + (S - smem_base) + (size_t)j * 64;
      const int sw = (j >> 1) & 3;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t off = (uint32_t)(((2 * half + c) ^ sw) << 4);
        *reinterpret_cast<uint4*>(row + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        *reinterpret_cast<uint4*>(row + s_bytes + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
      }
      fence_proxy_async();      // visible to the bulk copies and to the tensor core
    };
    auto put_state = [&](const float* v) {
      float4* dst = reinterpret_cast<float4*>(sH + u * NROWS + n0);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    };
    long long rowoff[HROWS];                    // element offset of (row, step 0, unit j) in the (B, W, G) tensors, or -1
    float h[HROWS];                             // state of unit j (z threads)
#pragma unroll
    for (int i = 0; i < HROWS; ++i) {
      const int b = sRow[n0 + i];
      rowoff[i] = (b >= 0 && act) ? (long long)b * W * G + j : -1;
      h[i] = (is_z && act && b >= 0 && p.h0) ? p.h0[(long long)b * p.ldh0 + j] : 0.f;
    }
    int t_last = -1;
    for (int t = W - 1; t >= 0; --t) if (sAny[t]) { t_last = t; break; }
    // initial state: exchange h(0) (not when every step of the tile is masked: nobody would wait for the copies)
    if (is_z && act && t_last >= 0) {
      put_state(h);
      put_unit(S_h, h);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");       // state exchange buffer and own S_h slice written
    if (tid == 0 && t_last >= 0) send_slice(S_h, bar_h);
    uint32_t ph_d1 = 0, ph_d2 = 0;
    PROF_DECL
    for (int t = 0; t < W; ++t) {
      if (!sAny[t]) {   // whole tile masked at this step: carry state
        if (save) {
#pragma unroll
          for (int i = 0; i < HROWS; ++i) {
            if (rowoff[i] < 0) continue;
            const long long o = rowoff[i] + (long long)t * G;
            if (is_z) { p.Z[o] = 0.f; p.HH[o] = 0.f; p.HP[o] = h[i]; }
            else { p.R[o] = 0.f; p.RH[o] = 0.f; }
          }
        }
        continue;
      }
      // pre-activations of this thread's gate(s) from the input projection (in flight during the MMAs)
      float x1[HROWS], xh[HROWS];
#pragma unroll
      for (int i = 0; i < HROWS; ++i) {
        const float* x = p.XW + (rowoff[i] - j + (long long)t * G) * 3 + j;
        x1[i] = rowoff[i] >= 0 ? __ldg(x + (is_z ? 0 : G)) : 0.f;
        xh[i] = (rowoff[i] >= 0 && is_z) ? __ldg(x + 2 * G) : 0.f;
      }
      // ---- phase 1 results
      PROF(0)
      mbar_wait(bar_d1, ph_d1, 33);
      PROF(1)
      ph_d1 ^= 1;
      tc_fence_after();
      {
        uint32_t r[HROWS];
        TMEM_LD_16(tD1 + tlane + n0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HROWS; ++i) x1[i] = rec_act(__uint_as_float(r[i]) + x1[i], p.act);    // z or r
      }
      if (!is_z) {
        float rh[HROWS];
        const float4* src = reinterpret_cast<const float4*>(sH + u * NROWS + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 hv = src[q];
          rh[4 * q] = x1[4 * q] * hv.x; rh[4 * q + 1] = x1[4 * q + 1] * hv.y;
          rh[4 * q + 2] = x1[4 * q + 2] * hv.z; rh[4 * q + 3] = x1[4 * q + 3] * hv.w;
        }
        if (act) put_unit(S_rh, rh);
        asm volatile("bar.sync 2, 128;" ::: "memory");     // the r threads' slice of S_rh is complete
        if (tid == 64) send_slice(S_rh, bar_rh);
        PROF(2)
        if (save) {
#pragma unroll
          for (int i = 0; i < HROWS; ++i) {
            if (rowoff[i] < 0) continue;
            const bool on = sGM[t * NROWS + n0 + i] != 0;
            const long long o = rowoff[i] + (long long)t * G;
            p.R[o] = on ? x1[i] : 0.f;
            p.RH[o] = on ? rh[i] : 0.f;
          }
        }
        ph_d2 ^= 1;      // r threads only track the phase-2 barrier's parity
        PROF(3)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        PROF(6)
      } else {
        // ---- phase 2 results (z threads)
        PROF(2)
        mbar_wait(bar_d2, ph_d2, 34);
        PROF(4)
        ph_d2 ^= 1;
        tc_fence_after();
        uint32_t r[HROWS];
        TMEM_LD_16(tD2 + tlane + n0, r);
        tmem_ld_wait();
        tc_fence_before();
        float hp[HROWS];
        uint32_t onmask = 0;
#pragma unroll
        for (int i = 0; i < HROWS; ++i) {
          const bool on = sGM[t * NROWS + n0 + i] != 0;
          const float hh = tanh_fast(__uint_as_float(r[i]) + xh[i]);
          xh[i] = hh;
          hp[i] = h[i];
          if (on) { h[i] = x1[i] * h[i] + (1.f - x1[i]) * hh; onmask |= 1u << i; }
        }
        if (act && t != t_last) {
          put_state(h);
          put_unit(S_h, h);
        }
        PROF(5)
        asm volatile("bar.sync 1, 256;" ::: "memory");     // new state visible to the r threads of the next step
        if (tid == 0 && t != t_last) send_slice(S_h, bar_h);
        PROF(6)
        if (save) {        // after the hand-off: these stores overlap the next step's MMAs
#pragma unroll
          for (int i = 0; i < HROWS; ++i) {
            if (rowoff[i] < 0) continue;
            const bool on = (onmask >> i) & 1;
            const long long o = rowoff[i] + (long long)t * G;
            p.Z[o] = on ? x1[i] : 0.f;
            p.HH[o] = on ? xh[i] : 0.f;
            p.HP[o] = hp[i];
          }
        }
        PROF(7)
      }
    }
    PROF_PRINT("z    xload wait_d1 act1 - wait_d2 upd+put barsync+send save", tid == 0)
    PROF_PRINT("r    xload wait_d1 act+put+send save - - barsync", tid == 64)
    if (is_z && act) {
#pragma unroll
      for (int i = 0; i < HROWS; ++i) {
        const int b = sRow[n0 + i];
        if (b >= 0) p.hT[(long long)b * p.ldo + j] = h[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // no CTA retires while a peer's copies into it could still be in flight
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// BPTT on the tensor cores.  Thread (unit k, 8 batch rows) keeps d h[:, k] in registers.  Per step:
//   dah, daz (elementwise) -> own slice of S_dah / S_daz -> peers;   phase A:  drh_k = sum_j dah_j Wh[k][2G+j]  (critical)
//   and, behind it on the tensor core, dhp_k = sum_j daz_j Wh[k][j];  dar = drh * hp * act'(r) -> S_dar -> peers;
//   phase B: dhp_k += sum_j dar_j Wh[k][G+j].   The three weight slices (rows k of Wh, fp16) stay in tensor memory; the
// exchanged operands are fp16 hi + lo of the value times a power of two that follows the cluster-wide max |d h| (the
// CTAs' maxima ride along with the slices, one step of lag; first step: max|dhT| of the tile), so the fp16 range stays
// centred on the gradients however they decay or grow over the window; conversions saturate.  2 MMAs per k step:
// W_hi.S_hi + W_hi.S_lo.  The weight rounding (2^-12 relative) is that of every other backward GEMM of the tensor-core
// modes.  Matrix rows 64..127 repeat rows 0..63, so all eight epilogue warps read accumulators (8 batch rows each).
// S_daz is double-buffered by step parity: its MMAs run off the critical path and may still be in flight when the
// peers' next slices arrive.
struct BwdParams {
  int B, W, G, UC, act;
  const float* gm;
  const int* row_order;
  const float *Z, *R, *HH, *HP;   // (B, W, G)
  const float* Wh;                // (G, 3G)
  const float* dhT; long long lddh;
  float* dA;                      // (B, W, 3G)
  float* dh0; long long lddh0;
  float* db_partial;              // optional (4 * tiles, 3G): per (tile, row group) sums of dA over rows and steps
};
constexpr int BROWS = 8;            // batch rows per epilogue thread in the backward kernel

__host__ __device__ inline size_t bwd_smem_bytes(int G, int W) {
  const int KS16 = (G + 15) / 16;
  return (size_t)10 * KS16 * 1024 + (size_t)W * NROWS + (size_t)((W + 15) / 16) * 16 + 512 + 1024;
}

__global__ void __launch_bounds__(THREADS, 1) gru_bwd_tc_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  const int G = p.G, UC = p.UC, W = p.W, G3 = 3 * G;
  const int ksteps = (G + 15) / 16, KC = ksteps * 8;
  const uint32_t s_bytes = (uint32_t)ksteps * 1024;
  // S buffers (hi, lo each): dah | daz[0] | daz[1] | dar
  const uint32_t S_dah = smem_base, S_daz = S_dah + 2 * s_bytes, S_dar = S_dah + 8 * s_bytes;
  const uint32_t oGM = 10 * s_bytes;
  const uint32_t oAny = oGM + (uint32_t)W * NROWS;
  const uint32_t oBar = (oAny + (uint32_t)((W + 15) / 16) * 16 + 15u) & ~15u;
  uint8_t* sGM = sm + oGM;
  uint8_t* sAny = sm + oAny;
  const uint32_t bar_a = smem_base + oBar, bar_r = bar_a + 8, bar_dA = bar_a + 16, bar_dB = bar_a + 24;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + oBar + 32);
  float* sMax = reinterpret_cast<float*>(sm + oBar + 40);
  uint32_t* sLmax = reinterpret_cast<uint32_t*>(sm + oBar + 44);   // max |d h| of this CTA's units (float bits)
  int* sRow = reinterpret_cast<int*>(sm + oBar + 64);
  float* sPeer = reinterpret_cast<float*>(sm + oBar + 256);        // [step parity][rank][4]: the CTAs' max |d h|
  const uint32_t sPeer_addr = smem_base + oBar + 256;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  constexpr int CS = 4;
  const int tile = blockIdx.x / CS, b0 = tile * NROWS;
  const int own_units = max(0, min(UC, G - rank * UC));
  const uint32_t own_off = (uint32_t)(rank * UC) * 64u, own_bytes = (uint32_t)own_units * 64u;
  const uint32_t xbytes = (uint32_t)(G - own_units) * 64u * 2u;      // one operand (hi + lo) from the three peers

  if (tid == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_r, 1);
    mbar_init(bar_dA, 1);
    mbar_init(bar_dB, 1);
    *sMax = 0.f;
    *sLmax = 0u;
    for (int i = 0; i < 32; ++i) sPeer[i] = 0.f;
    fence_barrier_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = tid; i < oGM / 16; i += THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < W * NROWS; i += THREADS) {
    const int t = i / NROWS, n = i % NROWS;
    int b = -1;
    if (b0 + n < p.B) b = p.row_order ? p.row_order[b0 + n] : b0 + n;
    sGM[i] = (b >= 0 && p.gm[(long long)b * W + t] != 0.f) ? 1 : 0;
    if (t == 0) sRow[n] = b;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tDrh = tmem_base, tDhp = tmem_base + 32, tAh = tmem_base + 64, tAz = tAh + KC, tAr = tAz + KC;
  for (int t = tid; t < W; t += THREADS) {
    int any = 0;
    for (int n = 0; n < NROWS; ++n) any |= sGM[t * NROWS + n];
    sAny[t] = (uint8_t)any;
  }
  {   // per-tile scale: max |dhT| over the tile's rows (every CTA of the cluster computes the same value)
    float mx = 0.f;
    for (int i = tid; i < NROWS * G; i += THREADS) {
      const int b = sRow[i / G];
      if (b >= 0) mx = fmaxf(mx, fabsf(p.dhT[(long long)b * p.lddh + i % G]));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx > 0.f) atomicMax(reinterpret_cast<int*>(sMax), __float_as_int(mx));
  }
  if (warp < 8) {
    // weight rows -> tensor memory (fp16): lane m holds Wh[rank*UC + (m & 63)][gate*G + j], j = column*2 (+1)
    const int m = (warp & 3) * 32 + lane, u = m & 63, k = rank * UC + u;
    const bool rowact = u < UC && k < G;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    for (int ks = warp >> 2; ks < ksteps; ks += 2) {
      uint32_t wh[8], wz[8], wr[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float a[2] = {0.f, 0.f}, b[2] = {0.f, 0.f}, c[2] = {0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = ks * 16 + 2 * q + e;
          if (rowact && j < G) {
            const float* row = p.Wh + (long long)k * G3;
            b[e] = __ldg(row + j); c[e] = __ldg(row + G + j); a[e] = __ldg(row + 2 * G + j);
          }
        }
        const __half2 ha = __floats2half2_rn(a[0], a[1]), hb = __floats2half2_rn(b[0], b[1]), hc = __floats2half2_rn(c[0], c[1]);
        wh[q] = *reinterpret_cast<const uint32_t*>(&ha);
        wz[q] = *reinterpret_cast<const uint32_t*>(&hb);
        wr[q] = *reinterpret_cast<const uint32_t*>(&hc);
      }
      tmem_st8(tAh + lane_addr + 8 * ks, wh);
      tmem_st8(tAz + lane_addr + 8 * ks, wz);
      tmem_st8(tAr + lane_addr + 8 * ks, wr);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  fence_proxy_async();
  cluster_sync_all();
  // scale = 2^e with max * scale in [2^3, 2^4).  The first step takes it from max |dhT| of the tile; every later step
  // from the cluster-wide max |d h| the CTAs exchanged one step earlier (so vanishing or growing gradients over long
  // windows stay inside the fp16 range; one update of lag is covered by the 2^12 headroom).
  float scale = 1.f, inv_scale = 1.f;
  auto set_scale = [&](float mx, float& sc, float& isc) {
    if (mx > 0.f && mx < 3.0e38f) {
      int e;
      frexpf(mx, &e);                       // mx = f * 2^e, f in [0.5, 1)
      e = max(-120, min(120, 4 - e));
      sc = ldexpf(1.f, e);
      isc = ldexpf(1.f, -e);
    }
  };
  set_scale(*sMax, scale, inv_scale);
  auto send_op = [&](uint32_t S, uint32_t bar) {      // hi and lo slices of one operand to every peer
    if (own_bytes) {
#pragma unroll
      for (int c = 1; c < CS; ++c) {
        const uint32_t peer = (uint32_t)((rank + c) % CS);
        const uint32_t rbar = map_to_cta(bar, peer);
        bulk_s2peer(map_to_cta(S + own_off, peer), S + own_off, own_bytes, rbar);
        bulk_s2peer(map_to_cta(S + s_bytes + own_off, peer), S + s_bytes + own_off, own_bytes, rbar);
      }
    }
  };

  if (warp == 8) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(128, NROWS, true) | (1u << 16);
    uint32_t ph_a = 0, ph_r = 0, par = 0;
    for (int t = W - 1; t >= 0; --t) {
      if (!sAny[t]) continue;
      const uint32_t Sz = S_daz + par * 2 * s_bytes;
      mbar_wait(bar_a, ph_a, 41);
      ph_a ^= 1;
      tc_fence_after();
      if (leader) {
        for (int ks = 0; ks < ksteps; ++ks) {
          umma_ts(tDrh, tAh + 8 * ks, desc_mn64(S_dah + (uint32_t)ks * 1024), idesc, ks > 0);
          umma_ts(tDrh, tAh + 8 * ks, desc_mn64(S_dah + s_bytes + (uint32_t)ks * 1024), idesc, 1);
        }
        umma_commit(bar_dA);
        for (int ks = 0; ks < ksteps; ++ks) {     // off the critical path: overlaps the dar epilogue and exchange
          umma_ts(tDhp, tAz + 8 * ks, desc_mn64(Sz + (uint32_t)ks * 1024), idesc, ks > 0);
          umma_ts(tDhp, tAz + 8 * ks, desc_mn64(Sz + s_bytes + (uint32_t)ks * 1024), idesc, 1);
        }
      }
      __syncwarp();
      mbar_wait(bar_r, ph_r, 42);
      ph_r ^= 1;
      tc_fence_after();
      if (leader) {
        for (int ks = 0; ks < ksteps; ++ks) {
          umma_ts(tDhp, tAr + 8 * ks, desc_mn64(S_dar + (uint32_t)ks * 1024), idesc, 1);
          umma_ts(tDhp, tAr + 8 * ks, desc_mn64(S_dar + s_bytes + (uint32_t)ks * 1024), idesc, 1);
        }
        umma_commit(bar_dB);
      }
      __syncwarp();
      par ^= 1;
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3, half = warp >> 2;
    const int m = quarter * 32 + lane;
    const int u = m & 63, k = rank * UC + u;
    const bool act = u < UC && k < G;
    const int n0 = ((quarter >> 1) * 2 + half) * BROWS;      // this thread's 8 batch rows
    const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
    const int chunk_pos = ((n0 >> 3) ^ ((k >> 1) & 3)) << 4;  // 16-byte chunk of the unit's 64-byte row
    // scaled fp16 hi/lo of 8 values -> own row of the local S buffer
    auto put8 = [&](uint32_t S, const float* v) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = v[2 * q] * scale, b = v[2 * q + 1] * scale;
        uint32_t h;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
        uint32_t l;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(b - hf.y), "f"(a - hf.x));
        hi[q] = h; lo[q] = l;
      }
      uint8_t* row = sm + (S - smem_base) + (size_t)k * 64 + chunk_pos;
      *reinterpret_cast<uint4*>(row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(row + s_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      fence_proxy_async();
    };
    long long rowoff[BROWS];      // element offset of (row, step 0, unit k) in the (B, W, G) tensors, or -1
    float dh[BROWS];
#pragma unroll
    for (int i = 0; i < BROWS; ++i) {
      const int b = sRow[n0 + i];
      rowoff[i] = (b >= 0 && act) ? (long long)b * W * G + k : -1;
      dh[i] = rowoff[i] >= 0 ? p.dhT[(long long)b * p.lddh + k] : 0.f;
    }
    auto publish_local_max = [&]() {      // this thread's max |d h| -> CTA-wide max (completed by the next CTA barrier)
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < BROWS; ++i) mx = fmaxf(mx, fabsf(dh[i]));
#pragma unroll
      for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0 && mx > 0.f && mx < 3.0e38f) atomicMax(sLmax, __float_as_uint(mx));
    };
    publish_local_max();
    auto load_saved = [&](int t, float* z, float* r, float* hh, float* hp) {
#pragma unroll
      for (int i = 0; i < BROWS; ++i) {
        const long long o = rowoff[i] + (long long)t * G;
        const bool ok = rowoff[i] >= 0 && t >= 0;
        z[i] = ok ? __ldg(p.Z + o) : 0.f;
        r[i] = ok ? __ldg(p.R + o) : 0.f;
        hh[i] = ok ? __ldg(p.HH + o) : 0.f;
        hp[i] = ok ? __ldg(p.HP + o) : 0.f;
      }
    };
    auto next_active = [&](int t) {
      --t;
      while (t >= 0 && !sAny[t]) --t;
      return t;
    };
    float nz[BROWS], nr[BROWS], nhh[BROWS], nhp[BROWS];
    int t = next_active(W);
    load_saved(t, nz, nr, nhh, nhp);
    // steps after the last active one (none of the tile's rows is on): dA = 0
    auto zero_steps = [&](int t_hi, int t_lo) {      // steps t_lo < s < t_hi
      for (int s = t_hi - 1; s > t_lo; --s)
#pragma unroll
        for (int i = 0; i < BROWS; ++i)
          if (rowoff[i] >= 0) {
            float* d = p.dA + (rowoff[i] - k + (long long)s * G) * 3 + k;
            d[0] = 0.f; d[G] = 0.f; d[2 * G] = 0.f;
          }
    };
    zero_steps(W, t);
    uint32_t ph_dA = 0, ph_dB = 0, ph_a = 0, par = 0;
    float sum_z = 0.f, sum_r = 0.f, sum_h = 0.f;      // bias gradient: this thread's share of the column sums of dA
    while (t >= 0) {
      float z[BROWS], r[BROWS], hh[BROWS], hp[BROWS], dah[BROWS], daz[BROWS];
      uint32_t onmask = 0;
#pragma unroll
      for (int i = 0; i < BROWS; ++i) {
        z[i] = nz[i]; r[i] = nr[i]; hh[i] = nhh[i]; hp[i] = nhp[i];
        const bool on = rowoff[i] >= 0 && sGM[t * NROWS + n0 + i] != 0;
        if (on) onmask |= 1u << i;
        dah[i] = on ? dh[i] * (1.f - z[i]) * (1.f - hh[i] * hh[i]) : 0.f;
        daz[i] = on ? dh[i] * (hp[i] - hh[i]) * rec_act_grad(z[i], p.act) : 0.f;
      }
      if (act) {
        put8(S_dah, dah);
        put8(S_daz + par * 2 * s_bytes, daz);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) {
        const uint32_t slot = (par * 4 + (uint32_t)rank) * 16;
        sPeer[slot / 4] = __uint_as_float(*sLmax);      // this CTA's max |d h| rides with the slices
        *sLmax = 0u;
        fence_proxy_async();
        mbar_arrive_expect(bar_a, 2 * xbytes + 3 * 16);
        send_op(S_dah, bar_a);
        send_op(S_daz + par * 2 * s_bytes, bar_a);
#pragma unroll
        for (int c = 1; c < CS; ++c) {
          const uint32_t peer = (uint32_t)((rank + c) % CS);
          bulk_s2peer(map_to_cta(sPeer_addr + slot, peer), sPeer_addr + slot, 16, map_to_cta(bar_a, peer));
        }
      }
      // off the critical path: this step's dA (z and candidate parts), the next active step's saved activations
      const int tn = next_active(t);
#pragma unroll
      for (int i = 0; i < BROWS; ++i)
        if (rowoff[i] >= 0) {
          float* d = p.dA + (rowoff[i] - k + (long long)t * G) * 3 + k;
          d[0] = daz[i];
          d[2 * G] = dah[i];
          sum_z += daz[i];
          sum_h += dah[i];
        }
      load_saved(tn, nz, nr, nhh, nhp);
      // ---- phase A result: drh
      mbar_wait(bar_dA, ph_dA, 43);
      ph_dA ^= 1;
      tc_fence_after();
      // the four CTAs' max |d h| arrived with this step's slices (bar_a completed before the MMAs ran): next step's scale
      mbar_wait(bar_a, ph_a, 45);
      ph_a ^= 1;
      float scale_next = scale, inv_next = inv_scale;
      {
        const float* q = sPeer + par * 16;
        set_scale(fmaxf(fmaxf(q[0], q[4]), fmaxf(q[8], q[12])), scale_next, inv_next);
      }
      float dar[BROWS], dhn[BROWS];
      {
        uint32_t d[BROWS];
        TMEM_LD_8(tDrh + tlane + n0, d);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int i = 0; i < BROWS; ++i) {
          const float drh = __uint_as_float(d[i]) * inv_scale;
          const bool on = (onmask >> i) & 1;
          dar[i] = on ? drh * hp[i] * rec_act_grad(r[i], p.act) : 0.f;
          dhn[i] = fmaf(drh, r[i], dh[i] * z[i]);
        }
      }
      if (act) put8(S_dar, dar);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) {
        mbar_arrive_expect(bar_r, xbytes);
        send_op(S_dar, bar_r);
      }
#pragma unroll
      for (int i = 0; i < BROWS; ++i)
        if (rowoff[i] >= 0) {
          p.dA[(rowoff[i] - k + (long long)t * G) * 3 + G + k] = dar[i];
          sum_r += dar[i];
        }
      zero_steps(t, tn);
      // ---- phase B result: dhp
      mbar_wait(bar_dB, ph_dB, 44);
      ph_dB ^= 1;
      tc_fence_after();
      {
        uint32_t d[BROWS];
        TMEM_LD_8(tDhp + tlane + n0, d);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int i = 0; i < BROWS; ++i)
          if ((onmask >> i) & 1) dh[i] = fmaf(__uint_as_float(d[i]), inv_scale, dhn[i]);
      }
      scale = scale_next; inv_scale = inv_next;
      publish_local_max();
      par ^= 1;
      t = tn;
    }
    if (p.dh0) {
#pragma unroll
      for (int i = 0; i < BROWS; ++i)
        if (rowoff[i] >= 0) {
          const int b = sRow[n0 + i];
          p.dh0[(long long)b * p.lddh0 + k] = dh[i];
        }
    }
    if (p.db_partial && act) {      // row (tile, row group); summed by a fixed-order column sum afterwards
      float* row = p.db_partial + ((long long)tile * 4 + (n0 >> 3)) * G3;
      row[k] = sum_z; row[G + k] = sum_r; row[2 * G + k] = sum_h;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace grutc
}  // namespace lstur

using namespace lstur;

// Shapes the tensor-core recurrence covers: clusters of 4, at most 64 units per CTA, weights resident in tensor memory (G <= 224).
extern "C" int lstur_gru_tc_supported(int B, int W, int G) {
  if (!(B > 0 && W > 0 && G > 0 && G % 8 == 0)) return 0;
  const int UC = (G + 3) / 4;
  if (UC > 64) return 0;
  if (64 + 4 * ((G + 15) / 16) * 8 > grutc::TMEM_COLS) return 0;      // 2 accumulators + four weight parts in tensor memory
  return (grutc::smem_bytes(G, W) <= 227 * 1024 && grutc::bwd_smem_bytes(G, W) <= 227 * 1024) ? 1 : 0;
}

extern "C" int lstur_gru_fwd_tc(int B, int W, int G, const float* XW, const float* gm, const float* h0, long long ldh0,
                                const float* Wh, int rec_act, float* hT, long long ldo, float* Z, float* R, float* HH,
                                float* HP, float* RH, const int* row_order, cudaStream_t stream) {
  LSTUR_REQUIRE(B >= 0 && W > 0 && G > 0 && XW && gm && Wh && hT, "lstur_gru_fwd_tc");
  LSTUR_REQUIRE((Z && R && HH && HP && RH) || (!Z && !R && !HH && !HP && !RH), "lstur_gru_fwd_tc");
  LSTUR_REQUIRE(lstur_gru_tc_supported(B > 0 ? B : 1, W, G), "lstur_gru_fwd_tc(shape)");
  if (B == 0) return LSTUR_OK;
  grutc::Params p = {};
  p.B = B; p.W = W; p.G = G; p.UC = (G + 3) / 4; p.act = rec_act;
  p.gm = gm; p.row_order = row_order; p.XW = XW; p.h0 = h0; p.ldh0 = ldh0; p.Wh = Wh; p.hT = hT; p.ldo = ldo;
  p.Z = Z; p.R = R; p.HH = HH; p.HP = HP; p.RH = RH;
  const size_t smem = grutc::smem_bytes(G, W);
  cudaError_t e = cudaFuncSetAttribute(grutc::gru_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("lstur_gru_fwd_tc: cannot opt in to %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    return LSTUR_ERR_CUDA;
  }
  const int tiles = (B + grutc::NROWS - 1) / grutc::NROWS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles * 4);
  cfg.blockDim = dim3(grutc::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, grutc::gru_fwd_tc_kernel, p);
  if (e != cudaSuccess) {
    set_error("lstur_gru_fwd_tc: launch failed: %s", cudaGetErrorString(e));
    return LSTUR_ERR_CUDA;
  }
  LSTUR_CHECK_LAUNCH("lstur_gru_fwd_tc");
  return LSTUR_OK;
}

extern "C" int lstur_gru_tc_db_rows(int B) { return 4 * ((B + grutc::NROWS - 1) / grutc::NROWS); }

extern "C" int lstur_gru_bwd_tc(int B, int W, int G, const float* gm, const float* Z, const float* R, const float* HH,
                                const float* HP, const float* Wh, int rec_act, const float* dhT, long long lddh, float* dA,
                                float* dh0, long long lddh0, const int* row_order, float* db_partial, cudaStream_t stream) {
  LSTUR_REQUIRE(B >= 0 && W > 0 && G > 0 && gm && Z && R && HH && HP && Wh && dhT && dA, "lstur_gru_bwd_tc");
  LSTUR_REQUIRE(lstur_gru_tc_supported(B > 0 ? B : 1, W, G), "lstur_gru_bwd_tc(shape)");
  if (B == 0) return LSTUR_OK;
  grutc::BwdParams p = {};
  p.B = B; p.W = W; p.G = G; p.UC = (G + 3) / 4; p.act = rec_act;
  p.gm = gm; p.row_order = row_order; p.Z = Z; p.R = R; p.HH = HH; p.HP = HP; p.Wh = Wh;
  p.dhT = dhT; p.lddh = lddh; p.dA = dA; p.dh0 = dh0; p.lddh0 = lddh0; p.db_partial = db_partial;
  const size_t smem = grutc::bwd_smem_bytes(G, W);
  cudaError_t e = cudaFuncSetAttribute(grutc::gru_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("lstur_gru_bwd_tc: cannot opt in to %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    return LSTUR_ERR_CUDA;
  }
  const int tiles = (B + grutc::NROWS - 1) / grutc::NROWS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles * 4);
  cfg.blockDim = dim3(grutc::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, grutc::gru_bwd_tc_kernel, p);
  if (e != cudaSuccess) {
    set_error("lstur_gru_bwd_tc: launch failed: %s", cudaGetErrorString(e));
    return LSTUR_ERR_CUDA;
  }
  LSTUR_CHECK_LAUNCH("lstur_gru_bwd_tc");
  return LSTUR_OK;
}
