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
#include "tc_common.cuh"

namespace lstur {
namespace tc {

constexpr int SLOT = 32;                      // rows per title slot
constexpr int TPT = TILE_M / SLOT;            // titles per tile
constexpr int KBLK = 64;                      // K elements per block (128 B of bf16)
constexpr int A_TAP_BYTES = TILE_M * 128;     // 16 KB
constexpr int TAPS = 3;
constexpr int A_STAGE_BYTES = TAPS * A_TAP_BYTES;
constexpr int NUM_A_STAGES = 2;
constexpr int NUM_B_STAGES = 2;
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

// conv_w fp32 (3,E,F) -> per K block i = c*3 + j an [F rows][64 k] bf16 image, K-major, 128B-swizzled.
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
  long long byte = (long long)f * 128 + ((((kk >> 3) ^ (f & 7)) << 4) | ((kk & 7) << 1));
  img[blk * per_blk + byte / 2] = to16(v, fp16);
}

struct FwdParams {
  int n_titles, L, F, EC, Ep, V;
  const int* tok;                 // (n_titles, L)
  const uint16_t* emb;            // (V, Ep) fp16 or bf16
  const uint16_t* wimg;           // EC*3 blocks of F*64
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
};

template <bool FP16>
__global__ void __launch_bounds__(THREADS, 1) news_conv_tc_fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int F = p.F, EC = p.EC;
  const uint32_t b_stage_bytes = (uint32_t)F * 128u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + NUM_A_STAGES * A_STAGE_BYTES;
  const uint32_t misc_base = b_base + NUM_B_STAGES * b_stage_bytes;
  uint8_t* misc_gen = smem_gen + (misc_base - smem_base);
  // barriers: a_full[2], a_empty[2], b_full[2], b_empty[2], tmem_full, tmem_empty  (8 B each)
  const uint32_t bar_a_full = misc_base, bar_a_empty = misc_base + 16, bar_b_full = misc_base + 32,
                 bar_b_empty = misc_base + 48, bar_t_full = misc_base + 64, bar_t_empty = misc_base + 72;
  uint32_t* tmem_ptr_smem = (uint32_t*)(misc_gen + 80);
  float* s_z = (float*)(misc_gen + 128);          // [2 parity][2 halves][128 rows]
  int* s_any = (int*)(misc_gen + 128 + 2048);     // [2][2][128]
  float* s_bias = (float*)(misc_gen + 128 + 4096);  // [F]
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
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int i = 0; i < EC * TAPS; ++i) {
          mbar_wait(bar_b_empty + 8 * sb, ph ^ 1, 1);
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
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(bar_t_empty, pht ^ 1, 2);
        tc_fence_after();
        uint32_t accum = 0;
        for (int c = 0; c < EC; ++c) {
          mbar_wait(bar_a_full + 8 * sa, pha, 3);
          tc_fence_after();
          for (int j = 0; j < TAPS; ++j) {
            mbar_wait(bar_b_full + 8 * sb, phb, 4);
            tc_fence_after();
            const uint32_t a_addr = a_base + sa * A_STAGE_BYTES + j * A_TAP_BYTES;
            const uint32_t b_addr = b_base + sb * b_stage_bytes;
#pragma unroll
            for (int kk = 0; kk < KBLK / 16; ++kk) {
              const uint64_t ad = make_desc_k128(a_addr + kk * 32);
              umma_bf16(tmem_base, ad, make_desc_k128(b_addr + kk * 32), idesc0, accum);
              if (n1 > 0) umma_bf16(tmem_base + n0, ad, make_desc_k128(b_addr + 256 * 128 + kk * 32), idesc1, accum);
              accum = 1;
            }
            umma_commit(bar_b_empty + 8 * sb);
            if (++sb == NUM_B_STAGES) { sb = 0; phb ^= 1; }
          }
          umma_commit(bar_a_empty + 8 * sa);
          if (++sa == NUM_A_STAGES) { sa = 0; pha ^= 1; }
        }
        umma_commit(bar_t_full);
        pht ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== A producers: embedding gather -> three shifted swizzled tap tiles =====================
    const int pw = warp - 4;                 // title slot of the tile
    const int rsub = lane >> 3, piece = lane & 7;
    int sa = 0;
    uint32_t pha = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int n = tile * TPT + pw;
      int ids[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int t = 4 * i + rsub;
        int id = -1;
        if (n < p.n_titles && t < p.L) {
          id = p.tok[(long long)n * p.L + t];
          id = (id < 0 || id >= p.V) ? 0 : id;
        }
        ids[i] = id;
      }
      for (int c = 0; c < EC; ++c) {
        mbar_wait(bar_a_empty + 8 * sa, pha ^ 1, 5);
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = make_uint4(0, 0, 0, 0);
          if (ids[i] >= 0) v[i] = __ldg((const uint4*)(p.emb + (long long)ids[i] * p.Ep + c * KBLK + piece * 8));
        }
        if (p.drop_thr16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (ids[i] < 0) continue;
            const int t = 4 * i + rsub;
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
        const uint32_t stage = a_base + sa * A_STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = pw * SLOT + 4 * i + rsub;
#pragma unroll
          for (int j = 0; j < TAPS; ++j) {
            const int rr = (r + 1 - j) & (TILE_M - 1);
            const uint32_t addr = stage + j * A_TAP_BYTES + rr * 128 + ((piece ^ (rr & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[i].x), "r"(v[i].y), "r"(v[i].z),
                         "r"(v[i].w)
                         : "memory");
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a_full + 8 * sa);
        if (++sa == NUM_A_STAGES) { sa = 0; pha ^= 1; }
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
    const float xs = p.drop_thr16 ? p.inv_keep : 1.f;   // input-dropout scale folded into the epilogue
    uint32_t pht = 0;
    int par = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int n = tile * TPT + q, t = lane;
      const bool valid = n < p.n_titles && t < p.L;
      const long long m = (long long)n * p.L + t;
      const int tk = valid ? p.tok[m] : 0;
      uint16_t* crow = p.c_out + m * F;
      mbar_wait(bar_t_full, pht, 6);
      pht ^= 1;
      tc_fence_after();
      float z = 0.f;
      int any = 0;
      for (int ch = ch_beg; ch < ch_end; ++ch) {
        const int c0 = ch * 32;
        const int ncols = min(32, F - c0);
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        if (ncols == 32) {
          TMEM_LD_32(taddr, r);
        } else {
          TMEM_LD_16(taddr, r);
#pragma unroll
          for (int i = 16; i < 32; ++i) r[i] = 0;
        }
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int f = c0 + i;
          float v0 = 0.f, v1 = 0.f;
          if (i < ncols && tk != 0) {
            v0 = fmaxf(fmaf(__uint_as_float(r[i]), xs, s_bias[f]), 0.f);
            v1 = fmaxf(fmaf(__uint_as_float(r[i + 1]), xs, s_bias[f + 1]), 0.f);
            any |= (v0 > 0.f) | (v1 > 0.f);
            if (p.drop_thr16) {
              const uint32_t h = rng_u32(p.seed_c, (uint64_t)(m * F + f) >> 1);
              v0 = (h & 0xffffu) >= p.drop_thr16 ? v0 * p.inv_keep : 0.f;
              v1 = (h >> 16) >= p.drop_thr16 ? v1 * p.inv_keep : 0.f;
            }
          }
          const uint32_t pk = pack16x2<FP16>(v0, v1);
          packed[i >> 1] = pk;
          if (i < ncols) {
            z = fmaf(lo16<FP16>(pk), s_ka[f], z);
            z = fmaf(hi16<FP16>(pk), s_ka[f + 1], z);
          }
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(crow + c0);
          dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          if (ncols == 32) {
            dst[2] = make_uint4(packed[8], packed[9], packed[10], packed[11]);
            dst[3] = make_uint4(packed[12], packed[13], packed[14], packed[15]);
          }
        }
      }
      // TMEM drained: let the MMA warp start the next tile while we pool
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_t_empty);
      const int row = q * 32 + lane;
      s_z[(par * 2 + half) * 128 + row] = z;
      s_any[(par * 2 + half) * 128 + row] = any;
      asm volatile("bar.sync 1, 256;" ::: "memory");
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
      // pass 2: pooled[n, f] = sum_t w_t * C[t, f]; butterfly reduce-scatter over the 32 lanes (rows)
      for (int ch = ch_beg; ch < ch_end; ++ch) {
        const int c0 = ch * 32;
        const int ncols = min(32, F - c0);
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 0.f;
        if (valid && w != 0.f) {
          const uint4* src = reinterpret_cast<const uint4*>(crow + c0);
          const int nv = ncols == 32 ? 4 : 2;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < nv) {
              const uint4 u = src[g];
              x[g * 8 + 0] = w * lo16<FP16>(u.x); x[g * 8 + 1] = w * hi16<FP16>(u.x);
              x[g * 8 + 2] = w * lo16<FP16>(u.y); x[g * 8 + 3] = w * hi16<FP16>(u.y);
              x[g * 8 + 4] = w * lo16<FP16>(u.z); x[g * 8 + 5] = w * hi16<FP16>(u.z);
              x[g * 8 + 6] = w * lo16<FP16>(u.w); x[g * 8 + 7] = w * hi16<FP16>(u.w);
            }
          }
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
constexpr int WG_KTOK = 64;                         // tokens (K rows) per pipeline stage = 2 title slots
constexpr int WG_GROUP_BYTES = WG_KTOK * 128;       // one 64-element group of one stage: 8 KB
constexpr int WG_A_STAGE_BYTES = 2 * WG_GROUP_BYTES;
constexpr int WG_STAGES = 3;

struct WgradParams {
  int n_titles, L, F, EC, Ep, V;
  int n_kblocks, kb_per_split, n_slices, ngroups;
  const int* tok;
  const uint16_t* emb;        // (V, Ep)
  const uint16_t* dpre_img;   // n_kblocks * ngroups * 8 KB
  float* partial;             // [splits][n_slices*128][F]
  uint32_t drop_thr16, seed_x;
  float scale;
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
  const uint32_t bar_full = misc_base, bar_empty = misc_base + 32, bar_t_full = misc_base + 64;
  uint32_t* tmem_ptr_smem = (uint32_t*)(misc_gen + 80);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x, split = blockIdx.y;
  const int kb_beg = split * p.kb_per_split;
  const int kb_end = min(p.n_kblocks, kb_beg + p.kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 5);    // 4 producer warps + the B loader's expect_tx arrival
      mbar_init(bar_empty + 8 * s, 1);   // tcgen05.commit
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int n0 = F > 256 ? 256 : F, n1 = F - n0;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1, 11);
        mbar_expect_tx(bar_full + 8 * s, b_stage_bytes);
        bulk_g2s(b_base + s * b_stage_bytes, (const uint8_t*)p.dpre_img + (size_t)kb * b_stage_bytes, b_stage_bytes,
                 bar_full + 8 * s);
        if (++s == WG_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc0 = make_idesc_mn(TILE_M, n0, FP16), idesc1 = make_idesc_mn(TILE_M, n1 > 0 ? n1 : 16, FP16);
      int s = 0;
      uint32_t ph = 0, accum = 0;
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        mbar_wait(bar_full + 8 * s, ph, 12);
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
        umma_commit(bar_empty + 8 * s);
        if (++s == WG_STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(bar_t_full);
    }
  } else if (warp >= 4 && warp < 8) {
    // A producers: rows = tokens of the K block, two 64-column chunks (u0,u1) of this CTA's (tap, e) slice
    const int pw = warp - 4, rsub = lane >> 3, piece = lane & 7;
    int uj[2], uc[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int u = 2 * slice + i;
      uj[i] = u < TAPS * p.EC ? u / p.EC : -1;
      uc[i] = u < TAPS * p.EC ? u % p.EC : 0;
    }
    int s = 0;
    uint32_t ph = 0;
    for (int kb = kb_beg; kb < kb_end; ++kb) {
      int ids[4];
      long long mrow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = pw * 16 + 4 * i + rsub;
        const long long R = (long long)kb * WG_KTOK + r;
        const int n = (int)(R / SLOT), t = (int)(R % SLOT);
        int id = -1;
        if (n < p.n_titles && t < p.L) {
          id = p.tok[(long long)n * p.L + t];
          id = (id < 0 || id >= p.V) ? 0 : id;
        }
        ids[i] = id;
        mrow[i] = (long long)n * p.L + t;
      }
      mbar_wait(bar_empty + 8 * s, ph ^ 1, 13);
      const uint32_t stage = a_base + s * WG_A_STAGE_BYTES;
#pragma unroll
      for (int ui = 0; ui < 2; ++ui) {
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = make_uint4(0, 0, 0, 0);
          if (ids[i] >= 0 && uj[ui] >= 0)
            v[i] = __ldg((const uint4*)(p.emb + (long long)ids[i] * p.Ep + uc[ui] * KBLK + piece * 8));
        }
        if (p.drop_thr16 && uj[ui] >= 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (ids[i] < 0) continue;
            const uint64_t pair0 = ((uint64_t)mrow[i] * (uint64_t)p.Ep + (uint64_t)(uc[ui] * KBLK + piece * 8)) >> 1;
            uint32_t* w = reinterpret_cast<uint32_t*>(&v[i]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t h = rng_u32(p.seed_x, pair0 + q);
              uint32_t m = ((h & 0xffffu) >= p.drop_thr16 ? 0x0000ffffu : 0u) | ((h >> 16) >= p.drop_thr16 ? 0xffff0000u : 0u);
              w[q] &= m;
            }
          }
        }
        const int j = uj[ui] < 0 ? 1 : uj[ui];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = pw * 16 + 4 * i + rsub;
          const int rr = (r + 1 - j) & (WG_KTOK - 1);
          const uint32_t addr = stage + ui * WG_GROUP_BYTES + rr * 128 + ((piece ^ (rr & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[i].x), "r"(v[i].y), "r"(v[i].z),
                       "r"(v[i].w)
                       : "memory");
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
      if (++s == WG_STAGES) { s = 0; ph ^= 1; }
    }
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
extern "C" int lstur_tc_padded_e(int E) { return (E + tc::KBLK - 1) / tc::KBLK * tc::KBLK; }
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
  size_t smem = 1024 + (size_t)tc::NUM_A_STAGES * tc::A_STAGE_BYTES + (size_t)tc::NUM_B_STAGES * F * 128 + 128 + 4096 +
                (size_t)2 * F * sizeof(float);
  static bool attr_set = false;
  static size_t attr_smem = 0;
  if (!attr_set || smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(tc::news_conv_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tc::news_conv_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  if (fp16) tc::news_conv_tc_fwd_kernel<true><<<grid, tc::THREADS, smem, stream>>>(p);
  else tc::news_conv_tc_fwd_kernel<false><<<grid, tc::THREADS, smem, stream>>>(p);
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
extern "C" int lstur_tc_wgrad_kblocks(int n_titles) { return (n_titles * tc::SLOT + tc::WG_KTOK - 1) / tc::WG_KTOK; }
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
  p.n_titles = n_titles; p.L = L; p.F = F; p.Ep = lstur_tc_padded_e(E); p.EC = p.Ep / tc::KBLK; p.V = V;
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
  dim3 grid(p.n_slices, splits);
  if (fp16) tc::conv_wgrad_tc_kernel<true><<<grid, 384, smem, stream>>>(p);
  else tc::conv_wgrad_tc_kernel<false><<<grid, 384, smem, stream>>>(p);
  LSTUR_CHECK_LAUNCH("lstur_conv_wgrad_tc");
  long long n = (long long)3 * E * F;
  tc::wgrad_reduce_kernel<<<cdiv(n, 256), 256, 0, stream>>>(E, p.Ep, F, splits, p.n_slices * tc::TILE_M, p.partial, d_conv_w);
  LSTUR_CHECK_LAUNCH("lstur_conv_wgrad_tc(reduce)");
  return LSTUR_OK;
}
