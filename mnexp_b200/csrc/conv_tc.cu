// News encoder forward on the 5th-gen tensor cores: word-embedding gather fused into the
// Conv1D implicit GEMM (tcgen05.mma, fp32 accumulator in TMEM) with bias + ReLU + pad mask +
// Masking + Dropout + additive-attention pooling fused in the epilogue.
//
// Reference ops replaced (task/paper.py:141-158, models.py:474-489):
//   Embedding(mask_zero=False) -> Dropout -> Conv1D(F,3,'same',relu) -> pad-token mask ->
//   Masking -> Dropout -> SimpleAttentionMaskSupport.
//
// GEMM view: D[m, f] = sum_{j<3} sum_e X[m+j-1, e] * Wc[j, e, f], m = token position.
//   M tile  = 128 rows = 4 title slots of 32 rows (L <= 31 tokens + >=1 zero row, which is both the
//             right halo of its title and the left halo of the next one; rows wrap inside the tile).
//   N       = F (<= 512 TMEM columns), issued as two UMMAs per K step (256 + rest).
//   K       = 3 taps x Ep (E padded to a multiple of 64), K block = 64 bf16 = one 128B swizzle row.
// A operand: producer warps gather each embedding row ONCE per 64-column chunk (16B loads, 8 lanes
//   per row), apply the input dropout mask, and store it into the three tap tiles at row offsets
//   +1/0/-1 in the canonical K-major SWIZZLE_128B layout (16B chunk index XOR row%8).
// B operand: the conv weights are re-packed per call into bf16 K-major SWIZZLE_128B images in
//   consumption order, so one cp.async.bulk (TMA, UBLKCP) per K block lands an MMA-ready tile.
// Warp roles (512 threads): w0 B loader, w1 TMEM alloc + MMA issuer, w4-7 A producers,
//   w8-15 epilogue (TMEM lane quarter = warp%4, column half = (warp-8)/4).
#include <stdlib.h>

#include "tc_common.cuh"

namespace lstur {
namespace tc {

constexpr int SLOT = 32;                      // rows per title slot
constexpr int TPT = TILE_M / SLOT;            // titles per tile
constexpr int EPAD = 64;                      // embedding rows are padded to a multiple of 64 columns
constexpr int KBLK = 32;                      // K elements per pipeline block: 64-byte rows, SWIZZLE_64B
constexpr int ROWB = KBLK * 2;                // bytes per shared-memory row
constexpr int A_TAP_BYTES = TILE_M * ROWB;    // 8 KB
constexpr int TAPS = 3;
constexpr int A_STAGE_BYTES = TAPS * A_TAP_BYTES;   // 24 KB: the three shifted tap tiles of one 32-column chunk
constexpr int NUM_A_STAGES = 3;
constexpr int NUM_B_STAGES = 5;               // deep ring: covers TMA + barrier round-trip latency

// K-major SWIZZLE_64B descriptor: rows of 64 B, 8-row atoms of 512 B (16B chunk index XOR (row>>1)&3)
__device__ __forceinline__ uint64_t make_desc_k64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
constexpr int THREADS = 512;

// ------------------------------------------------------------------ operand packing kernels
// fp32 (V,E) -> bf16 (V,Ep), zero-padded columns.
__global__ void pack_emb_bf16_kernel(long long V, int E, int Ep, const float* __restrict__ src,
                                     uint16_t* __restrict__ dst, bool fp16) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * Ep) return;
  long long v = i / Ep;
  int e = (int)(i % Ep);
  dst[i] = to16(e < E ? src[v * E + e] : 0.f, fp16);
}

// conv_w fp32 (3,E,F) -> per K block i = c*3 + j an [F rows][32 k] 16-bit image, K-major, 64B-swizzled.
__global__ void pack_conv_w_kernel(int E, int F, int EC, const float* __restrict__ Wc, uint16_t* __restrict__ img,
                                   bool fp16) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long per_blk = (long long)F * KBLK;
  if (i >= per_blk * EC * TAPS) return;
  int blk = (int)(i / per_blk);
  int rem = (int)(i % per_blk);
  int f = rem / KBLK, kk = rem % KBLK;
  int c = blk / TAPS, j = blk % TAPS;
  int e = c * KBLK + kk;
  float v = e < E ? Wc[((long long)j * E + e) * F + f] : 0.f;
  long long byte = (long long)f * ROWB + ((((kk >> 3) ^ ((f >> 1) & 3)) << 4) | ((kk & 7) << 1));
  img[blk * per_blk + byte / 2] = to16(v, fp16);
}

struct FwdParams {
  int n_titles, L, F, EC, Ep, V;
  const int* tok;                 // (n_titles, L)
  const uint16_t* emb;            // (V, Ep) fp16 or bf16
  const uint16_t* wimg;           // EC*3 blocks of F*32 (EC = Ep/32)
  const float* conv_b;            // (F)
  const float* att_w;             // (F)
  const float* att_b;             // (1)
  uint16_t* c_out;                // (n_titles, L, F) attention input (fp16/bf16), saved for backward
  float* pooled;                  // (n_titles, F)
  float* att_a;                   // (n_titles, L) or null
  float* att_wt;                  // (n_titles, L) or null
  uint32_t drop_thr16;            // 0 = no dropout; keep iff h16 >= thr
  float inv_keep;
  uint32_t seed_x, seed_c;
  long long* trace;               // optional clock64 trace of CTA 0 (tools/perf_conv.py); null in production
};
#define TRACE(it, slot)                                                                    \
  do {                                                                                     \
    if (p.trace && blockIdx.x == 0 && (it) < 8) p.trace[(it) * 16 + (slot)] = clock64();   \
  } while (0)

struct EpiCtx {
  float xs, inv_keep;
  uint32_t thr, base_lo, inner0, inner1;
  const float* s_bias;
  const float* s_ka;
};

// Epilogue pass 1 for NCOLS accumulator columns of one token row: bias + ReLU (+ dropout) -> 16-bit C (stored),
// attention-logit partial sum z (from the ROUNDED values, so forward and backward see the same C) and the row max.
template <bool FP16, bool DROP, int NCOLS>
__device__ __forceinline__ void epi_pass1_chunk(const EpiCtx& ec, uint32_t taddr, int c0, bool live, bool valid,
                                                uint16_t* crow, float& z, float& vmax) {
  uint32_t r[32];
  if (NCOLS == 32) { TMEM_LD_32(taddr, r); } else { TMEM_LD_16(taddr, r); }
  tmem_ld_wait();
  uint32_t packed[NCOLS / 2];
  if (live) {
#pragma unroll
    for (int g = 0; g < NCOLS / 4; ++g) {
      const float4 b4 = *reinterpret_cast<const float4*>(ec.s_bias + c0 + 4 * g);
      const float4 k4 = *reinterpret_cast<const float4*>(ec.s_ka + c0 + 4 * g);
      float v0 = fmaxf(fmaf(__uint_as_float(r[4 * g + 0]), ec.xs, b4.x), 0.f);
      float v1 = fmaxf(fmaf(__uint_as_float(r[4 * g + 1]), ec.xs, b4.y), 0.f);
      float v2 = fmaxf(fmaf(__uint_as_float(r[4 * g + 2]), ec.xs, b4.z), 0.f);
      float v3 = fmaxf(fmaf(__uint_as_float(r[4 * g + 3]), ec.xs, b4.w), 0.f);
      vmax = fmaxf(vmax, fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)));      // Masking(): any(C != 0) before the dropout
      if (DROP) {
        const uint32_t lo0 = ec.base_lo + (uint32_t)((c0 >> 1) + 2 * g), lo1 = lo0 + 1u;
        const uint32_t h0 = lowbias32(lo0 ^ (lo0 < ec.base_lo ? ec.inner1 : ec.inner0));
        const uint32_t h1 = lowbias32(lo1 ^ (lo1 < ec.base_lo ? ec.inner1 : ec.inner0));
        v0 = (h0 & 0xffffu) >= ec.thr ? v0 * ec.inv_keep : 0.f;
        v1 = (h0 >> 16) >= ec.thr ? v1 * ec.inv_keep : 0.f;
        v2 = (h1 & 0xffffu) >= ec.thr ? v2 * ec.inv_keep : 0.f;
        v3 = (h1 >> 16) >= ec.thr ? v3 * ec.inv_keep : 0.f;
      }
      const uint32_t p0 = pack16x2<FP16>(v0, v1), p1 = pack16x2<FP16>(v2, v3);
      packed[2 * g] = p0;
      packed[2 * g + 1] = p1;
      z = fmaf(lo16<FP16>(p0), k4.x, z);
      z = fmaf(hi16<FP16>(p0), k4.y, z);
      z = fmaf(lo16<FP16>(p1), k4.z, z);
      z = fmaf(hi16<FP16>(p1), k4.w, z);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NCOLS / 2; ++i) packed[i] = 0u;
  }
  if (valid) {
    uint4* dst = reinterpret_cast<uint4*>(crow + c0);
#pragma unroll
    for (int g = 0; g < NCOLS / 8; ++g) dst[g] = make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
  }
}

__device__ __forceinline__ void epi_load_chunk(const uint16_t* crow, int c0, int F, bool live, uint4* v) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    v[g] = make_uint4(0, 0, 0, 0);
    if (live && c0 + 8 * g < F) v[g] = *reinterpret_cast<const uint4*>(crow + c0 + 8 * g);
  }
}

template <bool FP16, bool DROP>
__global__ void __launch_bounds__(THREADS, 1) news_conv_tc_fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int F = p.F, EC = p.EC;
  const uint32_t b_stage_bytes = (uint32_t)F * ROWB;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + NUM_A_STAGES * A_STAGE_BYTES;
  const uint32_t misc_base = b_base + NUM_B_STAGES * b_stage_bytes;
  uint8_t* misc_gen = smem_gen + (misc_base - smem_base);
  // barriers (8 B each): a_full[<=4] a_empty[<=4] b_full[<=6] b_empty[<=6] tmem_full tmem_empty
  const uint32_t bar_a_full = misc_base, bar_a_empty = misc_base + 32, bar_b_full = misc_base + 64,
                 bar_b_empty = misc_base + 112, bar_t_full = misc_base + 160, bar_t_empty = misc_base + 168;
  uint32_t* tmem_ptr_smem = (uint32_t*)(misc_gen + 176);
  float* s_z = (float*)(misc_gen + 256);          // [2 parity][2 halves][128 rows]
  int* s_any = (int*)(misc_gen + 256 + 2048);     // [2][2][128]
  float* s_bias = (float*)(misc_gen + 256 + 4096);  // [F]
  float* s_ka = s_bias + F;                          // [F]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.n_titles + TPT - 1) / TPT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NUM_A_STAGES; ++s) {
      mbar_init(bar_a_full + 8 * s, 4);   // one arrival per producer warp
      mbar_init(bar_a_empty + 8 * s, 1);  // tcgen05.commit
    }
    for (int s = 0; s < NUM_B_STAGES; ++s) {
      mbar_init(bar_b_full + 8 * s, 1);   // expect_tx arrival + bytes
      mbar_init(bar_b_empty + 8 * s, 1);
    }
    mbar_init(bar_t_full, 1);
    mbar_init(bar_t_empty, 8);            // one arrival per epilogue warp
    fence_barrier_init();
  }
  for (int f = threadIdx.x; f < F; f += THREADS) {
    s_bias[f] = p.conv_b[f];
    s_ka[f] = p.att_w[f];
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int n0 = F > 256 ? 256 : F, n1 = F - n0;

  if (warp == 0) {
    // ===================== B loader (TMA bulk copies of pre-swizzled weight blocks) =====================
    if (lane == 0) {
      int sb = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        for (int i = 0; i < EC * TAPS; ++i) {
          mbar_wait(bar_b_empty + 8 * sb, ph ^ 1, 1);
          if (i == 0) TRACE(it, 9);
          if (i == EC * TAPS - 1) TRACE(it, 10);
          mbar_expect_tx(bar_b_full + 8 * sb, b_stage_bytes);
          bulk_g2s(b_base + sb * b_stage_bytes, (const uint8_t*)p.wimg + (size_t)i * b_stage_bytes, b_stage_bytes,
                   bar_b_full + 8 * sb);
          if (++sb == NUM_B_STAGES) { sb = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc0 = make_idesc(TILE_M, n0, FP16), idesc1 = make_idesc(TILE_M, n1 > 0 ? n1 : 16, FP16);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0, pht = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(bar_t_empty, pht ^ 1, 2);
        tc_fence_after();
        TRACE(it, 0);
        uint32_t accum = 0;
        for (int c = 0; c < EC; ++c) {
          mbar_wait(bar_a_full + 8 * sa, pha, 3);
          tc_fence_after();
          if (c == 0) TRACE(it, 1);
          for (int j = 0; j < TAPS; ++j) {
            mbar_wait(bar_b_full + 8 * sb, phb, 4);
            tc_fence_after();
            const uint32_t a_addr = a_base + sa * A_STAGE_BYTES + j * A_TAP_BYTES;
            const uint32_t b_addr = b_base + sb * b_stage_bytes;
#pragma unroll
            for (int kk = 0; kk < KBLK / 16; ++kk) {
              const uint64_t ad = make_desc_k64(a_addr + kk * 32);
              umma_bf16(tmem_base, ad, make_desc_k64(b_addr + kk * 32), idesc0, accum);
              if (n1 > 0) umma_bf16(tmem_base + n0, ad, make_desc_k64(b_addr + 256 * ROWB + kk * 32), idesc1, accum);
              accum = 1;
            }
            umma_commit(bar_b_empty + 8 * sb);
            if (++sb == NUM_B_STAGES) { sb = 0; phb ^= 1; }
          }
          umma_commit(bar_a_empty + 8 * sa);
          if (++sa == NUM_A_STAGES) { sa = 0; pha ^= 1; }
        }
        TRACE(it, 2);
        umma_commit(bar_t_full);
        pht ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== A producers: embedding gather -> three shifted swizzled tap tiles =====================
    // lane -> (row within a group of 8, 16-byte piece of the 64-byte row); loads for chunk c+1 (and the token ids of
    // the next tile) are issued before chunk c is hashed and stored, so L2 latency is off the critical path.
    const int pw = warp - 4;                 // title slot of the tile
    const int rsub = lane >> 2, piece = lane & 3;
    int sa = 0;
    uint32_t pha = 0;
    auto load_ids = [&](int tile, int* ids) {
      const int n = tile * TPT + pw;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = 8 * i + rsub;
        int id = -1;
        if (tile < n_tiles && n < p.n_titles && t < p.L) {
          id = p.tok[(long long)n * p.L + t];
          id = (id < 0 || id >= p.V) ? 0 : id;
        }
        ids[i] = id;
      }
    };
    auto load_rows = [&](const int* ids, int c, uint4* v) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = make_uint4(0, 0, 0, 0);
        if (ids[i] >= 0) v[i] = __ldg((const uint4*)(p.emb + (long long)ids[i] * p.Ep + c * KBLK + piece * 8));
      }
    };
    int ids[4], ids_next[4];
    uint4 v[4], v_next[4];
    load_ids(blockIdx.x, ids);
    load_rows(ids, 0, v_next);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int n = tile * TPT + pw;
      if (warp == 4 && lane == 0) TRACE(it, 7);
      for (int c = 0; c < EC; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = v_next[i];
        if (c + 1 < EC) {
          load_rows(ids, c + 1, v_next);
          if (c + 2 == EC) load_ids(tile + gridDim.x, ids_next);
        } else {
          if (EC == 1) load_ids(tile + gridDim.x, ids_next);
          load_rows(ids_next, 0, v_next);
        }
        if (DROP) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (ids[i] < 0) continue;
            const int t = 8 * i + rsub;
            const uint64_t pair0 = (((uint64_t)n * p.L + t) * (uint64_t)p.Ep + (uint64_t)(c * KBLK + piece * 8)) >> 1;
            uint32_t* w = reinterpret_cast<uint32_t*>(&v[i]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t h = rng_u32(p.seed_x, pair0 + q);
              uint32_t m = ((h & 0xffffu) >= p.drop_thr16 ? 0x0000ffffu : 0u) | ((h >> 16) >= p.drop_thr16 ? 0xffff0000u : 0u);
              w[q] &= m;
            }
          }
        }
        mbar_wait(bar_a_empty + 8 * sa, pha ^ 1, 5);
        const uint32_t stage = a_base + sa * A_STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = pw * SLOT + 8 * i + rsub;
#pragma unroll
          for (int j = 0; j < TAPS; ++j) {
            const int rr = (r + 1 - j) & (TILE_M - 1);
            const uint32_t addr = stage + j * A_TAP_BYTES + rr * ROWB + ((piece ^ ((rr >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[i].x), "r"(v[i].y), "r"(v[i].z),
                         "r"(v[i].w)
                         : "memory");
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a_full + 8 * sa);
        if (++sa == NUM_A_STAGES) { sa = 0; pha ^= 1; }
        if (c + 1 == EC) {
#pragma unroll
          for (int i = 0; i < 4; ++i) ids[i] = ids_next[i];
          if (warp == 4 && lane == 0) TRACE(it, 8);
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: bias/ReLU/masks/dropout/attention pooling =====================
    // Thread (q, lane) owns token row 32q+lane of the tile = token `lane` of title tile*4+q; the two
    // column halves of a row are handled by warps ew and ew+4 and combined through shared memory.
    const int ew = warp - 8, q = ew & 3, half = ew >> 2;
    const int nch = (F + 31) / 32;
    const int ch_split = (nch + 1) / 2;
    const int ch_beg = half == 0 ? 0 : ch_split, ch_end = half == 0 ? ch_split : nch;
    const float att_bias = p.att_b[0];
    EpiCtx ec;
    ec.xs = DROP ? p.inv_keep : 1.f;   // input-dropout scale folded into the epilogue
    ec.inv_keep = p.inv_keep;
    ec.thr = p.drop_thr16;
    ec.s_bias = s_bias;
    ec.s_ka = s_ka;
    uint32_t pht = 0;
    int par = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int n = tile * TPT + q, t = lane;
      const bool valid = n < p.n_titles && t < p.L;
      const long long m = (long long)n * p.L + t;
      const int tk = valid ? p.tok[m] : 0;
      uint16_t* crow = p.c_out + m * F;
      if (DROP) {
        const uint64_t base = ((uint64_t)m * (uint64_t)F) >> 1;     // F is even: pair index of (m, f) = base + f/2
        ec.base_lo = (uint32_t)base;
        const uint32_t hi = (uint32_t)(base >> 32), k = 0x9e3779b9u * (p.seed_c + 1u);
        ec.inner0 = lowbias32(hi + k);
        ec.inner1 = lowbias32(hi + 1u + k);
      }
      mbar_wait(bar_t_full, pht, 6);
      pht ^= 1;
      tc_fence_after();
      if (warp == 8 && lane == 0) TRACE(it, 3);
      float z = 0.f, vmax = 0.f;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int ch = ch_beg; ch < ch_end; ++ch) {
        const int c0 = ch * 32;
        if (F - c0 >= 32) epi_pass1_chunk<FP16, DROP, 32>(ec, trow + c0, c0, tk != 0, valid, crow, z, vmax);
        else epi_pass1_chunk<FP16, DROP, 16>(ec, trow + c0, c0, tk != 0, valid, crow, z, vmax);
      }
      // TMEM drained: let the MMA warp start the next tile while we pool
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_t_empty);
      if (warp == 8 && lane == 0) TRACE(it, 4);
      const int row = q * 32 + lane;
      s_z[(par * 2 + half) * 128 + row] = z;
      s_any[(par * 2 + half) * 128 + row] = vmax > 0.f ? 1 : 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 8 && lane == 0) TRACE(it, 5);
      const float zt = s_z[(par * 2) * 128 + row] + s_z[(par * 2 + 1) * 128 + row];
      const int anyt = s_any[(par * 2) * 128 + row] | s_any[(par * 2 + 1) * 128 + row];
      par ^= 1;
      const float a = tanhf(zt + att_bias);
      const float e = (valid && anyt) ? expf(a) : 0.f;
      const float S = warp_sum(e);
      const float w = e / (S + 1e-7f);
      if (half == 0 && valid) {
        if (p.att_a) p.att_a[m] = a;
        if (p.att_wt) p.att_wt[m] = w;
      }
      // pass 2: pooled[n, f] = sum_t w_t * C[t, f]; butterfly reduce-scatter over the 32 lanes (rows).  The row is
      // re-read from global (own writes, L2) one chunk ahead of the reduction.
      const bool live = valid && w != 0.f;
      uint4 nxt[4];
      epi_load_chunk(crow, ch_beg * 32, F, live, nxt);
      for (int ch = ch_beg; ch < ch_end; ++ch) {
        const int c0 = ch * 32;
        const int ncols = min(32, F - c0);
        uint4 cur[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) cur[g] = nxt[g];
        if (ch + 1 < ch_end) epi_load_chunk(crow, c0 + 32, F, live, nxt);
        float x[32];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          x[g * 8 + 0] = w * lo16<FP16>(cur[g].x); x[g * 8 + 1] = w * hi16<FP16>(cur[g].x);
          x[g * 8 + 2] = w * lo16<FP16>(cur[g].y); x[g * 8 + 3] = w * hi16<FP16>(cur[g].y);
          x[g * 8 + 4] = w * lo16<FP16>(cur[g].z); x[g * 8 + 5] = w * hi16<FP16>(cur[g].z);
          x[g * 8 + 6] = w * lo16<FP16>(cur[g].w); x[g * 8 + 7] = w * hi16<FP16>(cur[g].w);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int k = 0; k < off; ++k) {
            const float send = up ? x[k] : x[k + off];
            const float keep = up ? x[k + off] : x[k];
            x[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        if (n < p.n_titles && lane < ncols) p.pooled[(long long)n * F + c0 + lane] = x[0];
      }
      if (warp == 8 && lane == 0) TRACE(it, 6);
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// =====================================================================================================
// Conv1D weight gradient on tcgen05:  dW[(j,e), f] = sum_m X[m+j-1, e] * dPre[m, f]
//   (backward of keras Conv1D, task/paper.py:146; X = dropout(Embedding(tok)), dPre from attn bwd)
// GEMM view: M = 3*Ep rows (j,e) split in 128-row slices (one per CTA.x), N = F, K = token slot rows.
// Both operands are "MN-major" (the reduction index = token is the slow index of the row-major data):
//   A[K=token][M=e] : gathered embedding rows — the same 128B-swizzled row image as in the forward, but
//                     described to the tensor core as MN-major (64-element groups, LBO between groups).
//   B[K=token][N=f] : dPre, written by the attention-backward kernel directly as pre-swizzled K-block
//                     images (64 tokens x ceil(F/64) groups x 128 B), fetched with one bulk copy each.
// The token range is split over CTA.y; partial sums go to global and are reduced in a fixed order.
constexpr int WG_KTOK = 32;                         // tokens (K rows) per pipeline stage = one title slot
constexpr int WG_GROUP_BYTES = WG_KTOK * 128;       // one 64-element group of one stage: 4 KB
constexpr int WG_A_STAGE_BYTES = 2 * WG_GROUP_BYTES;
constexpr int WG_STAGES = 6;

struct WgradParams {
  int n_titles, L, F, EC, Ep, V;
  int n_kblocks, kb_per_split, n_slices, ngroups;
  int cluster;                // CTAs per cluster along x (= slices sharing one dPre stream), 1 = no multicast
  const int* tok;
  const uint16_t* emb;        // (V, Ep)
  const uint16_t* dpre_img;   // n_kblocks * ngroups * 8 KB
  float* partial;             // [splits][n_slices*128][F]
  uint32_t drop_thr16, seed_x;
  float scale;
  long long* trace;           // optional: accumulated wait cycles of CTA (0,0)'s roles (tools/perf_conv.py)
};

template <bool FP16>
__global__ void __launch_bounds__(384, 1) conv_wgrad_tc_kernel(const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int F = p.F;
  const uint32_t b_stage_bytes = (uint32_t)p.ngroups * WG_GROUP_BYTES;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + WG_STAGES * WG_A_STAGE_BYTES;
  const uint32_t misc_base = b_base + WG_STAGES * b_stage_bytes;
  uint8_t* misc_gen = smem_gen + (misc_base - smem_base);
  const uint32_t bar_full = misc_base, bar_empty = misc_base + 64, bar_t_full = misc_base + 128;
  uint32_t* tmem_ptr_smem = (uint32_t*)(misc_gen + 144);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x, split = blockIdx.y;
  const int CS = p.cluster;
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << CS) - 1u);
  const int kb_beg = split * p.kb_per_split;
  const int kb_end = min(p.n_kblocks, kb_beg + p.kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 5);    // 4 producer warps + the B loader's expect_tx arrival
      mbar_init(bar_empty + 8 * s, CS);  // tcgen05.commit from every CTA of the cluster (stage reused cluster-wide)
    }
    mbar_init(bar_t_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();     // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int n0 = F > 256 ? 256 : F, n1 = F - n0;

  if (warp == 0) {
    if (lane == 0) {
      // each CTA fetches 1/CS of the dPre block and multicasts it to every CTA of the cluster
      const uint32_t chunk = ((b_stage_bytes / CS) + 15u) & ~15u;
      const uint32_t my_off = crank * chunk;
      const uint32_t my_len = my_off < b_stage_bytes ? min(chunk, b_stage_bytes - my_off) : 0u;
      int s = 0;
      uint32_t ph = 0;
      long long twl = 0;
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        long long t0 = clock64();
        mbar_wait(bar_empty + 8 * s, ph ^ 1, 11);
        twl += clock64() - t0;
        if (kb + 1 == kb_end && p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[5] = twl;
        mbar_expect_tx(bar_full + 8 * s, b_stage_bytes);
        const uint8_t* src = (const uint8_t*)p.dpre_img + (size_t)kb * b_stage_bytes + my_off;
        if (CS > 1) {
          if (my_len) bulk_g2s_mc(b_base + s * b_stage_bytes + my_off, src, my_len, bar_full + 8 * s, cmask);
        } else {
          bulk_g2s(b_base + s * b_stage_bytes, src, b_stage_bytes, bar_full + 8 * s);
        }
        if (++s == WG_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc0 = make_idesc_mn(TILE_M, n0, FP16), idesc1 = make_idesc_mn(TILE_M, n1 > 0 ? n1 : 16, FP16);
      int s = 0;
      uint32_t ph = 0, accum = 0;
      long long tw = 0, t_start = clock64();
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        long long t0 = clock64();
        mbar_wait(bar_full + 8 * s, ph, 12);
        tw += clock64() - t0;
        tc_fence_after();
        const uint32_t a_addr = a_base + s * WG_A_STAGE_BYTES, b_addr = b_base + s * b_stage_bytes;
#pragma unroll
        for (int kk = 0; kk < WG_KTOK / 16; ++kk) {
          const uint64_t ad = make_desc_mn128(a_addr + kk * 2048, WG_GROUP_BYTES);
          umma_bf16(tmem_base, ad, make_desc_mn128(b_addr + kk * 2048, WG_GROUP_BYTES), idesc0, accum);
          if (n1 > 0)
            umma_bf16(tmem_base + n0, ad, make_desc_mn128(b_addr + 4 * WG_GROUP_BYTES + kk * 2048, WG_GROUP_BYTES), idesc1,
                      accum);
          accum = 1;
        }
        if (CS > 1) umma_commit_mc(bar_empty + 8 * s, cmask);
        else umma_commit(bar_empty + 8 * s);
        if (++s == WG_STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(bar_t_full);
      if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) { p.trace[0] = tw; p.trace[1] = clock64() - t_start; p.trace[2] = kb_end - kb_beg; }
    }
  } else if (warp >= 4 && warp < 8) {
    // A producers: one K block = the 32 token rows of title n = kb; thread -> (row r, piece pair q): 16-byte pieces
    // 2q, 2q+1 of both 64-column chunks (u0,u1) of this CTA's (tap, e) slice.  Next block's ids / rows are prefetched.
    const int pt = threadIdx.x - 128;          // 0..127
    const int r = pt >> 2, q = pt & 3;
    int uj[2], uc[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int u = 2 * slice + i;
      uj[i] = u < TAPS * p.EC ? u / p.EC : -1;
      uc[i] = u < TAPS * p.EC ? u % p.EC : 0;
    }
    auto load_id = [&](int kb) {
      int id = -1;
      if (kb < kb_end && kb < p.n_titles && r < p.L) {
        id = p.tok[(long long)kb * p.L + r];
        id = (id < 0 || id >= p.V) ? 0 : id;
      }
      return id;
    };
    auto load_rows = [&](int id, uint4* v) {
#pragma unroll
      for (int ui = 0; ui < 2; ++ui)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          v[ui * 2 + h] = make_uint4(0, 0, 0, 0);
          if (id >= 0 && uj[ui] >= 0)
            v[ui * 2 + h] = __ldg((const uint4*)(p.emb + (long long)id * p.Ep + uc[ui] * EPAD + (2 * q + h) * 8));
        }
    };
    int s = 0;
    uint32_t ph = 0;
    long long twait = 0, t_start = clock64();
    int id = load_id(kb_beg), id_next = load_id(kb_beg + 1);
    uint4 v[4], v_next[4];
    load_rows(id, v_next);
    for (int kb = kb_beg; kb < kb_end; ++kb) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = v_next[i];
      const int id_cur = id;
      id = id_next;
      load_rows(id, v_next);
      id_next = load_id(kb + 2);
      if (p.drop_thr16 && id_cur >= 0) {
        const uint64_t mrow = (uint64_t)kb * p.L + r;
#pragma unroll
        for (int ui = 0; ui < 2; ++ui) {
          if (uj[ui] < 0) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t pair0 = (mrow * (uint64_t)p.Ep + (uint64_t)(uc[ui] * EPAD + (2 * q + h) * 8)) >> 1;
            uint32_t* w = reinterpret_cast<uint32_t*>(&v[ui * 2 + h]);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              uint32_t hsh = rng_u32(p.seed_x, pair0 + x);
              uint32_t m = ((hsh & 0xffffu) >= p.drop_thr16 ? 0x0000ffffu : 0u) | ((hsh >> 16) >= p.drop_thr16 ? 0xffff0000u : 0u);
              w[x] &= m;
            }
          }
        }
      }
      long long t0 = clock64();
      mbar_wait(bar_empty + 8 * s, ph ^ 1, 13);
      twait += clock64() - t0;
      const uint32_t stage = a_base + s * WG_A_STAGE_BYTES;
#pragma unroll
      for (int ui = 0; ui < 2; ++ui) {
        const int j = uj[ui] < 0 ? 1 : uj[ui];
        const int rr = (r + 1 - j) & (WG_KTOK - 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t addr = stage + ui * WG_GROUP_BYTES + rr * 128 + (((2 * q + h) ^ (rr & 7)) << 4);
          const uint4 x = v[ui * 2 + h];
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
      if (++s == WG_STAGES) { s = 0; ph ^= 1; }
    }
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && pt == 0) { p.trace[3] = twait; p.trace[4] = clock64() - t_start; }
  } else if (warp >= 8) {
    // epilogue (once): TMEM -> scaled fp32 partial sums in global memory
    const int q = warp & 3;
    mbar_wait(bar_t_full, 0, 14);
    tc_fence_after();
    const int row = q * 32 + lane;
    float* dst = p.partial + ((size_t)split * p.n_slices * TILE_M + (size_t)slice * TILE_M + row) * F;
    const int nch = (F + 31) / 32;
    for (int ch = 0; ch < nch; ++ch) {
      const int c0 = ch * 32, ncols = min(32, F - c0);
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      if (ncols == 32) { TMEM_LD_32(taddr, r); } else { TMEM_LD_16(taddr, r); }
      tmem_ld_wait();
      if (kb_end > kb_beg) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (i < ncols)
            *reinterpret_cast<float4*>(dst + c0 + i) =
                make_float4(__uint_as_float(r[i]) * p.scale, __uint_as_float(r[i + 1]) * p.scale,
                            __uint_as_float(r[i + 2]) * p.scale, __uint_as_float(r[i + 3]) * p.scale);
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (i < ncols) *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();     // no CTA leaves while peers may still multicast into it / arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// d_conv_w[j][e][f] = sum_split partial[split][j*Ep + e][f]   (fixed order, deterministic)
__global__ void wgrad_reduce_kernel(int E, int Ep, int F, int splits, int rows_total, const float* __restrict__ partial,
                                    float* __restrict__ dW) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)TAPS * E * F) return;
  int f = (int)(i % F);
  int je = (int)(i / F);
  int j = je / E, e = je % E;
  long long row = (long long)j * Ep + e;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += partial[((long long)s * rows_total + row) * F + f];
  dW[i] = acc;
}

}  // namespace tc
}  // namespace lstur

// ======================================================================== host side
#include "plan.h"

using namespace lstur;

extern "C" int lstur_conv_tc_available(void) { return 1; }

// Padded embedding width (multiple of the 64-element K block).
extern "C" int lstur_tc_padded_e(int E) { return (E + tc::EPAD - 1) / tc::EPAD * tc::EPAD; }
// Elements (bf16) of the packed conv-weight image.
extern "C" long long lstur_tc_wimg_elems(int E, int F) {
  return (long long)(lstur_tc_padded_e(E) / tc::KBLK) * tc::TAPS * F * tc::KBLK;
}

// keras Embedding weights (task/paper.py:132-138) -> bf16 (V, Ep) table used by the tensor-core gather.
extern "C" int lstur_pack_word_emb_16(long long V, int E, const float* word_emb, void* emb_bf16, int fp16,
                                      cudaStream_t stream) {
  LSTUR_REQUIRE(V > 0 && E > 0, "lstur_pack_word_emb_16");
  int Ep = lstur_tc_padded_e(E);
  long long n = V * Ep;
  tc::pack_emb_bf16_kernel<<<cdiv(n, 256), 256, 0, stream>>>(V, E, Ep, word_emb, (uint16_t*)emb_bf16, fp16 != 0);
  LSTUR_CHECK_LAUNCH("lstur_pack_word_emb_16");
  return LSTUR_OK;
}

// Conv1D kernel (3,E,F) (task/paper.py:146) -> swizzled bf16 K-block images.
extern "C" int lstur_pack_conv_w_tc(int E, int F, const float* conv_w, void* wimg, int fp16, cudaStream_t stream) {
  LSTUR_REQUIRE(E > 0 && F > 0, "lstur_pack_conv_w_tc");
  long long n = lstur_tc_wimg_elems(E, F);
  tc::pack_conv_w_kernel<<<cdiv(n, 256), 256, 0, stream>>>(E, F, lstur_tc_padded_e(E) / tc::KBLK, conv_w,
                                                           (uint16_t*)wimg, fp16 != 0);
  LSTUR_CHECK_LAUNCH("lstur_pack_conv_w_tc");
  return LSTUR_OK;
}

extern "C" int lstur_tc_supported(int L, int E, int F, int KS) {
  return KS == 3 && L >= 1 && L <= tc::SLOT - 1 && E >= 1 && F >= 16 && F % 16 == 0 && F <= tc::TMEM_COLS &&
         (F <= 256 || F - 256 >= 16);
}

static void* g_tc_trace_ptr = nullptr;
// Debug/profiling hook: device buffer of 8*16 int64 receiving clock64() stamps of CTA 0's warp roles (NULL = off).
extern "C" int lstur_tc_set_trace(void* dev_buf) { g_tc_trace_ptr = dev_buf; return LSTUR_OK; }

// Fused news-encoder forward (k1-k6): tokens (n_titles,L) -> C (bf16, saved), pooled (n_titles,F), att a / w.
extern "C" int lstur_news_conv_tc_fwd(int n_titles, int L, int E, int F, int V, const int* tokens, const void* emb_bf16,
                                      const void* wimg, const float* conv_b, const float* att_w, const float* att_b,
                                      void* c_out_bf16, float* pooled, float* att_a, float* att_wt, float dropout,
                                      unsigned seed, int fp16, int max_ctas, cudaStream_t stream) {
  LSTUR_REQUIRE(n_titles >= 0 && lstur_tc_supported(L, E, F, 3), "lstur_news_conv_tc_fwd");
  LSTUR_REQUIRE(dropout >= 0.f && dropout < 1.f && c_out_bf16 && pooled, "lstur_news_conv_tc_fwd");
  if (n_titles == 0) return LSTUR_OK;
  tc::FwdParams p;
  p.n_titles = n_titles; p.L = L; p.F = F; p.Ep = lstur_tc_padded_e(E); p.EC = p.Ep / tc::KBLK; p.V = V;
  p.tok = tokens; p.emb = (const uint16_t*)emb_bf16; p.wimg = (const uint16_t*)wimg;
  p.conv_b = conv_b; p.att_w = att_w; p.att_b = att_b;
  p.c_out = (uint16_t*)c_out_bf16; p.pooled = pooled; p.att_a = att_a; p.att_wt = att_wt;
  p.drop_thr16 = dropout > 0.f ? (uint32_t)(dropout * 65536.0f) : 0u;
  p.inv_keep = 1.f / (1.f - dropout);
  p.seed_x = seed * 2u; p.seed_c = seed * 2u + 1u;
  p.trace = (long long*)g_tc_trace_ptr;
  size_t smem = 1024 + (size_t)tc::NUM_A_STAGES * tc::A_STAGE_BYTES + (size_t)tc::NUM_B_STAGES * F * tc::ROWB + 256 + 4096 +
                (size_t)2 * F * sizeof(float);
  static bool attr_set = false;
  static size_t attr_smem = 0;
  if (!attr_set || smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(tc::news_conv_tc_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::news_conv_tc_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::news_conv_tc_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::news_conv_tc_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("lstur_news_conv_tc_fwd: cannot opt in to %zu B of shared memory: %s", smem, cudaGetErrorString(e));
      return LSTUR_ERR_CUDA;
    }
    attr_set = true;
    attr_smem = smem;
  }
  int n_tiles = (n_titles + tc::TPT - 1) / tc::TPT;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int grid = n_tiles < sms ? n_tiles : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  const bool drop = p.drop_thr16 != 0;
  if (fp16 && drop) tc::news_conv_tc_fwd_kernel<true, true><<<grid, tc::THREADS, smem, stream>>>(p);
  else if (fp16) tc::news_conv_tc_fwd_kernel<true, false><<<grid, tc::THREADS, smem, stream>>>(p);
  else if (drop) tc::news_conv_tc_fwd_kernel<false, true><<<grid, tc::THREADS, smem, stream>>>(p);
  else tc::news_conv_tc_fwd_kernel<false, false><<<grid, tc::THREADS, smem, stream>>>(p);
  LSTUR_CHECK_LAUNCH("lstur_news_conv_tc_fwd");
  return LSTUR_OK;
}

extern "C" int lstur_news_encoder_tc_fwd_internal(const lstur_plan* p, const lstur_weights* w, void* ws, int training,
                                                  unsigned seed, cudaStream_t st) {
  const lstur_config& c = p->c;
  void* emb = W<void>(p, ws, "emb_bf16");
  void* wimg = W<void>(p, ws, "wimg");
  LSTUR_REQUIRE(emb && wimg, "lstur_news_encoder_tc_fwd_internal");
  const int fp16 = c.precision == LSTUR_PREC_FP16_TC;
  RC(lstur_pack_word_emb_16(c.V, c.E, w->word_emb, emb, fp16, st));
  RC(lstur_pack_conv_w_tc(c.E, c.F, DP(p, w->dense, "conv_w"), wimg, fp16, st));
  PROBE_BEGIN(p, LSTUR_PROBE_CONV_FWD, st);
  RC(lstur_news_conv_tc_fwd(p->N, c.L, c.E, c.F, c.V, W<int>(p, ws, "tokens"), emb, wimg, DP(p, w->dense, "conv_b"),
                            DP(p, w->dense, "att_w"), DP(p, w->dense, "att_b"), W<void>(p, ws, "C16"),
                            W<float>(p, ws, "pooled"), W<float>(p, ws, "att_a"), W<float>(p, ws, "att_w"),
                            training ? c.dropout : 0.f, seed, fp16, 0, st));
  PROBE_END(p, LSTUR_PROBE_CONV_FWD, st);
  return LSTUR_OK;
}


// ---- wgrad host side ------------------------------------------------------------------------------
extern "C" int lstur_tc_wgrad_kblocks(int n_titles) { return n_titles; }   // one 32-row slot = one K block
extern "C" int lstur_tc_wgrad_groups(int F) { return (F + 63) / 64; }
// bytes of the dPre image consumed by lstur_conv_wgrad_tc
extern "C" size_t lstur_tc_dpre_img_bytes(int n_titles, int F) {
  return (size_t)lstur_tc_wgrad_kblocks(n_titles) * lstur_tc_wgrad_groups(F) * tc::WG_GROUP_BYTES;
}
extern "C" int lstur_tc_wgrad_splits(int n_titles, int E) {
  int n_slices = (tc::TAPS * lstur_tc_padded_e(E) + tc::TILE_M - 1) / tc::TILE_M;
  int kb = lstur_tc_wgrad_kblocks(n_titles);
  int sms = 148;
  int splits = sms / n_slices;
  if (splits < 1) splits = 1;
  if (splits > kb) splits = kb;
  return splits;
}
extern "C" size_t lstur_tc_wgrad_partial_bytes(int n_titles, int E, int F) {
  int n_slices = (tc::TAPS * lstur_tc_padded_e(E) + tc::TILE_M - 1) / tc::TILE_M;
  return (size_t)lstur_tc_wgrad_splits(n_titles, E) * n_slices * tc::TILE_M * F * sizeof(float);
}

// d_conv_w (3,E,F) = sum over tokens of X[m+j-1,e] * dPre[m,f]; X re-gathered from emb_16 with the forward's dropout
// stream (seed), dPre given as the K-block image written by lstur_attn_pool_bwd_img.
extern "C" int lstur_conv_wgrad_tc(int n_titles, int L, int E, int F, int V, const int* tokens, const void* emb_16,
                                   const void* dpre_img, float* d_conv_w, float dropout, unsigned seed, int fp16,
                                   void* partial_ws, size_t partial_bytes, cudaStream_t stream) {
  LSTUR_REQUIRE(n_titles >= 0 && lstur_tc_supported(L, E, F, 3), "lstur_conv_wgrad_tc");
  if (n_titles == 0) {
    cudaMemsetAsync(d_conv_w, 0, (size_t)3 * E * F * sizeof(float), stream);
    return LSTUR_OK;
  }
  tc::WgradParams p;
  p.n_titles = n_titles; p.L = L; p.F = F; p.Ep = lstur_tc_padded_e(E); p.EC = p.Ep / tc::EPAD; p.V = V;
  p.n_kblocks = lstur_tc_wgrad_kblocks(n_titles);
  p.n_slices = (tc::TAPS * p.Ep + tc::TILE_M - 1) / tc::TILE_M;
  int splits = lstur_tc_wgrad_splits(n_titles, E);
  p.kb_per_split = (p.n_kblocks + splits - 1) / splits;
  p.ngroups = lstur_tc_wgrad_groups(F);
  p.tok = tokens; p.emb = (const uint16_t*)emb_16; p.dpre_img = (const uint16_t*)dpre_img;
  p.partial = (float*)partial_ws;
  LSTUR_REQUIRE(partial_ws != nullptr && partial_bytes >= lstur_tc_wgrad_partial_bytes(n_titles, E, F), "lstur_conv_wgrad_tc");
  p.drop_thr16 = dropout > 0.f ? (uint32_t)(dropout * 65536.0f) : 0u;
  p.seed_x = seed * 2u;
  p.scale = 1.f / (1.f - dropout);
  p.trace = (long long*)g_tc_trace_ptr;
  size_t smem = 1024 + (size_t)tc::WG_STAGES * (tc::WG_A_STAGE_BYTES + (size_t)p.ngroups * tc::WG_GROUP_BYTES) + 256;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(tc::conv_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tc::conv_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("lstur_conv_wgrad_tc: cannot opt in to %zu B of shared memory: %s", smem, cudaGetErrorString(e));
      return LSTUR_ERR_CUDA;
    }
    attr_smem = smem;
  }
  // Cluster multicast of the dPre stream (one L2 read for all slices) is correct but measured SLOWER on B200
  // (only 15 clusters of 8 fit in one wave and the slices run in lockstep): off unless LSTUR_MULTICAST=1.
  p.cluster = (p.n_slices >= 2 && p.n_slices <= 8 && getenv("LSTUR_MULTICAST")) ? p.n_slices : 1;
  if (p.cluster > 1) {
    // all clusters must be co-resident in ONE wave (a cluster needs `cluster` SMs of the same GPC): ask the runtime
    static int max_clusters[9] = {0};
    if (!max_clusters[p.cluster]) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(p.cluster, 1);
      q.blockDim = dim3(384);
      q.dynamicSmemBytes = smem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = p.cluster; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int nc = 0;
      cudaError_t e = fp16 ? cudaOccupancyMaxActiveClusters(&nc, tc::conv_wgrad_tc_kernel<true>, &q)
                           : cudaOccupancyMaxActiveClusters(&nc, tc::conv_wgrad_tc_kernel<false>, &q);
      if (e != cudaSuccess || nc < 1) { cudaGetLastError(); nc = 1; p.cluster = 1; }
      max_clusters[p.cluster] = nc;
    }
    if (p.cluster > 1 && splits > max_clusters[p.cluster]) splits = max_clusters[p.cluster];
    p.kb_per_split = (p.n_kblocks + splits - 1) / splits;
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.n_slices, splits);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = fp16 ? cudaLaunchKernelEx(&cfg, tc::conv_wgrad_tc_kernel<true>, p)
                         : cudaLaunchKernelEx(&cfg, tc::conv_wgrad_tc_kernel<false>, p);
    if (e != cudaSuccess) {
      set_error("lstur_conv_wgrad_tc: launch failed: %s", cudaGetErrorString(e));
      return LSTUR_ERR_CUDA;
    }
  }
  LSTUR_CHECK_LAUNCH("lstur_conv_wgrad_tc");
  long long n = (long long)3 * E * F;
  tc::wgrad_reduce_kernel<<<cdiv(n, 256), 256, 0, stream>>>(E, p.Ep, F, splits, p.n_slices * tc::TILE_M, p.partial, d_conv_w);
  LSTUR_CHECK_LAUNCH("lstur_conv_wgrad_tc(reduce)");
  return LSTUR_OK;
}
