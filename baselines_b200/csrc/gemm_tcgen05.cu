// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a
// (fp16 operands, fp32 accumulate in TMEM).
//
//   warp 0      : TMA producer   (tiled or IM2COL-mode cp.async.bulk.tensor -> swizzled smem ring, mbarrier tx)
//   warp 1      : MMA issuer     (one lane issues tcgen05.mma.cta_group::1.kind::f16, commits to mbarriers)
//   warp 2      : TMEM allocator (2 accumulator stages: the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4..7  : epilogue       (tcgen05.ld 32x32b -> registers -> fused bias / activation / mask / atomics /
//                                 pixel-shuffle scatter)
//
// The reduction dimension is processed in stages of 64 elements made of 64/CPT "taps" of CPT elements
// (CPT = 64, 32 or 16 -> 128 B / 64 B / 32 B swizzled smem rows):
//   plain GEMM      : a tap is a 64-wide K chunk                                 (CPT = 64)
//   implicit conv   : a tap is one filter position (r, s) x CPT input channels, fetched straight from the NHWC
//                     activation by TMA im2col mode -- the im2col matrix is never materialised
// Operand layouts:
//   K-major  : A[M,K] rows = output pixels, B[N,K] = weights          C = A * B^T     (forward, dgrad)
//   MN-major : A[K,M], B[K,N] with K = batch pixels                    C = A^T * B     (wgrad, split-K atomics)
//
// Replaces tf.matmul (a2c/utils.py:63), tf.nn.conv2d (a2c/utils.py:56) and their gradients
// (tf.gradients via ppo2/model.py:102) of the reference.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200rl {

static constexpr int BM = 128;
static constexpr int BK = 64;          // elements of the reduction dimension per pipeline stage
static constexpr int UMMA_K = 16;
// TMA, MMA, TMEM alloc, spare + epilogue warps: 2 accumulator stages x GEMM_CG column groups x 4 lane quadrants.  The
// epilogue is a latency-bound dependent chain per warp (tcgen05.ld -> math -> pack -> store), so the columns of a tile
// are split over GEMM_CG warps per quadrant.
static constexpr int GEMM_CG = 2;      // measured: fc1 data gradient 10.2 ms -> 8.3 ms per cfg-2 update, fwd / wgrad unchanged
static constexpr int NUM_THREADS = 128 + 2 * GEMM_CG * 128;


struct ConvCoords {       // im2col traversal of the A operand (all zero for plain GEMMs)
  int OW, OH;             // grid of base pixels per image (GEMM rows = n*OH*OW + p*OW + q)
  int stride_w, stride_h; // traversal strides (input pixels per base-pixel step)
  int lower_w, lower_h;   // coordinate of base pixel 0 (= -padding)
  int S, taps;            // filter width (taps per filter row) and total number of taps R*S
};

struct ShuffleOut {       // MODE_F16_SHUFFLE: GEMM row (n,i,j), col (py,px,c) -> dx[n, s*i+py, s*j+px, c]
  int H, W, C, s;
};

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  void* C;
  long long ldc;
  const float* bias;
  const __half* saved;     // saved activation for MODE_F16_DACT / SHUFFLE masks
  const uint16_t* saved_bits;   // MODE_F16_DACT alternative: 1 bit per element (activation > 0), word (row*ld_saved + col)/16
  long long ld_saved;
  float alpha;
  int mode, act;
  ConvCoords cv;
  ShuffleOut sh;
  bool vec32;               // f16 epilogue rows/columns are 32-byte aligned: 256-bit loads / stores
  int rm_C, rm_OW, rm_Wg;   // rm_C > 0: output column (pix*rm_C + c) is stored at ((pix/rm_OW)*rm_Wg + pix%rm_OW)*rm_C + c
};

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// CPT: channels per tap (64 / 32 / 16); MN_MAJOR: wgrad layout; IM2COL: A operand through TMA im2col mode
// BRES (K-major implicit conv only): the whole weight matrix (taps x [BN x CPT]) is loaded ONCE per CTA and
// stays resident in shared memory; the ring then carries only the A (activation patch) tiles.
static constexpr int BRES_STAGES = 7;
static constexpr int BRES_B_BYTES = 80 * 1024;

// TMODE >= 0 fixes the epilogue mode at compile time (plain GEMMs): the multi-mode epilogue is ~35-40 KB of
// SASS, beyond the 32 KB L1.5 instruction cache, and the epilogue warps then stall on instruction fetch.
template <int BN, int CPT, bool MN_MAJOR, bool IM2COL, bool BRES, int TMODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using C_ = Cfg<BN>;
  static_assert(!BRES || (IM2COL && !MN_MAJOR), "resident weights: K-major implicit conv only");
  constexpr int STAGES = BRES ? BRES_STAGES : C_::STAGES;
  constexpr int STAGE_BYTES = BRES ? C_::A_BYTES : C_::STAGE_BYTES;
  constexpr int BROWB = MN_MAJOR ? (BN >= 64 ? 128 : BN * 2) : 0;     // MN-major B: bytes per pixel row
  constexpr uint32_t LAYOUT_B = (BROWB == 64) ? 4u : 2u;
  constexpr int BCHUNKS = (BN >= 64) ? BN / 64 : 1;
  constexpr int TPS = BK / CPT;                     // taps per stage
  constexpr int ROWB = CPT * 2;                     // bytes per smem row of a K-major / A-MN sub-tile
  constexpr uint32_t LAYOUT_A = (ROWB == 128) ? 2u : (ROWB == 64) ? 4u : 6u;
  constexpr int A_SUB = BM * ROWB;                  // K-major: one tap of 128 rows
  constexpr int B_SUB = BN * ROWB;                  // K-major: one tap of BN weight rows
  constexpr int A_CHUNK = BK * ROWB;                // MN-major: 64 pixel rows x CPT channels (one tap)
  constexpr int MCH = BM / CPT;                     // MN-major: taps (M chunks) per M tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bres = smem + STAGES * STAGE_BYTES;            // resident weights (BRES only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(bres + (BRES ? BRES_B_BYTES : 0));
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint64_t* bres_bar = bars + 2 * STAGES + 4;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_work = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128 * GEMM_CG);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      if (BRES && elect_one()) {
        mbar_arrive_expect_tx(bres_bar, (uint32_t)p.cv.taps * B_SUB);
        for (int g = 0; g < p.cv.taps; ++g) tma_load_2d(bres + g * B_SUB, &tmB, bres_bar, g * CPT, 0);
      }
      __syncwarp();
      int s = 0;
      uint32_t ph = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int n_tile = work % p.n_tiles;
        const int t2 = work / p.n_tiles;
        const int m_tile = t2 % p.m_tiles;
        const int split = t2 / p.m_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        int cw = 0, ch = 0, cn = 0;
        if (IM2COL && !MN_MAJOR) {                 // base pixel of this 128-row tile
          const int m0 = m_tile * BM;
          const int q = m0 % p.cv.OW, t3 = m0 / p.cv.OW;
          cw = q * p.cv.stride_w + p.cv.lower_w;
          ch = (t3 % p.cv.OH) * p.cv.stride_h + p.cv.lower_h;
          cn = t3 / p.cv.OH;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one()) {
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + C_::A_BYTES;
          if (!MN_MAJOR) {
            const int ntap = IM2COL ? min(TPS, p.cv.taps - kb * TPS) : TPS;
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)ntap * (A_SUB + (BRES ? 0 : B_SUB)));
#pragma unroll
            for (int t = 0; t < TPS; ++t) {
              if (t < ntap) {
                const int g = kb * TPS + t;
                if (IM2COL)
                  tma_load_im2col_4d(sa + t * A_SUB, &tmA, &full_bar[s], 0, cw, ch, cn, (uint16_t)(g % p.cv.S),
                                     (uint16_t)(g / p.cv.S));
                else
                  tma_load_2d(sa + t * A_SUB, &tmA, &full_bar[s], g * CPT, m_tile * BM);
                if (!BRES) tma_load_2d(sb + t * B_SUB, &tmB, &full_bar[s], g * CPT, n_tile * BN);
              }
            }
          } else {
            int nch = MCH;
            if (IM2COL) {
              nch = min(MCH, p.cv.taps - m_tile * MCH);
              const int k0 = kb * BK;              // first pixel row of this K block
              const int q = k0 % p.cv.OW, t3 = k0 / p.cv.OW;
              cw = q * p.cv.stride_w + p.cv.lower_w;
              ch = (t3 % p.cv.OH) * p.cv.stride_h + p.cv.lower_h;
              cn = t3 / p.cv.OH;
            }
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)nch * A_CHUNK + C_::B_BYTES);
#pragma unroll
            for (int j = 0; j < MCH; ++j) {
              if (j < nch) {
                if (IM2COL) {
                  const int g = m_tile * MCH + j;
                  tma_load_im2col_4d(sa + j * A_CHUNK, &tmA, &full_bar[s], 0, cw, ch, cn, (uint16_t)(g % p.cv.S),
                                     (uint16_t)(g / p.cv.S));
                } else {
                  tma_load_2d(sa + j * A_CHUNK, &tmA, &full_bar[s], m_tile * BM + j * CPT, kb * BK);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < BCHUNKS; ++j)
              tma_load_2d(sb + j * (BK * BROWB), &tmB, &full_bar[s], n_tile * BN + j * 64, kb * BK);
          }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t IDESC = (1u << 4)                      // D = f32
                               | (0u << 7) | (0u << 10)       // A, B = f16
                               | ((MN_MAJOR ? 1u : 0u) << 15) | ((MN_MAJOR ? 1u : 0u) << 16)
                               | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    {
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int t2 = work / p.n_tiles;
        const int split = t2 / p.m_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(&tempty_bar[as], aph ^ 1);
        if (BRES) mbar_wait(bres_bar, 0);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        uint32_t acc = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = BRES ? smem_u32(bres) + (uint32_t)(kb * TPS) * B_SUB : a_addr + C_::A_BYTES;
          if (!MN_MAJOR) {
            const int ntap = IM2COL ? min(TPS, p.cv.taps - kb * TPS) : TPS;
#pragma unroll
            for (int t = 0; t < TPS; ++t) {
              if (t < ntap) {
#pragma unroll
                for (int k = 0; k < CPT / UMMA_K; ++k) {
                  const uint64_t adesc = make_sdesc(a_addr + t * A_SUB + k * (UMMA_K * 2), 16, 8 * ROWB, LAYOUT_A);
                  const uint64_t bdesc = make_sdesc(b_addr + t * B_SUB + k * (UMMA_K * 2), 16, 8 * ROWB, LAYOUT_A);
                  umma_f16(tmem_d, adesc, bdesc, IDESC, acc);
                  acc = 1;
                }
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = make_sdesc(a_addr + k * (UMMA_K * ROWB), A_CHUNK, 8 * ROWB, LAYOUT_A);
              const uint64_t bdesc = make_sdesc(b_addr + k * (UMMA_K * BROWB), BK * BROWB, 8 * BROWB, LAYOUT_B);
              umma_f16(tmem_d, adesc, bdesc, IDESC, acc);
              acc = 1;
            }
          }
          umma_commit(&empty_bar[s]);   // smem slot reusable once these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull_bar[as]);    // accumulator ready for the epilogue
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    // warp e = warp - 4: accumulator stage as = e / (4*GEMM_CG) (even / odd work items of this CTA), column group
    // cg = (e / 4) % GEMM_CG, TMEM lane quarter ew = warp % 4
    const int ew = warp & 3;
    const int mode = (TMODE >= 0) ? TMODE : p.mode;
    const int as = (warp - 4) / (4 * GEMM_CG);
    const int cg = ((warp - 4) >> 2) % GEMM_CG;
    constexpr int NCG = (BN / GEMM_CG >= 16) ? BN / GEMM_CG : 16;         // columns per group (BN = 16*k)
    constexpr int NCH = NCG / 16;
    constexpr int PF = (NCH > 4) ? 4 : NCH;                               // mask chunks prefetched ahead
    const bool active = cg * NCG < BN;
    uint32_t aph = 0;
    for (int work = blockIdx.x + as * (int)gridDim.x; work < total_work; work += 2 * (int)gridDim.x) {
      const int n_tile = work % p.n_tiles;
      const int t2 = work / p.n_tiles;
      const int m_tile = t2 % p.m_tiles;
      const int row = m_tile * BM + ew * 32 + lane;
      const bool row_ok = row < p.M;
      // the activation mask of a data gradient does not depend on the accumulator: fetch it while the MMAs run
      uint32_t sv[PF][8];
      const bool dact_vec = (mode == MODE_F16_DACT) && p.vec32 && row_ok && active;
      const bool use_bits = (mode == MODE_F16_DACT) && p.saved_bits != nullptr;
      auto prefetch = [&](int k, int slot) {
        const int col0 = n_tile * BN + cg * NCG + 16 * k;
        if (dact_vec && col0 + 16 <= p.N) {
          if (use_bits) sv[slot][0] = __ldg(p.saved_bits + (((long long)row * p.ld_saved + col0) >> 4));
          else ldg256(p.saved + (long long)row * p.ld_saved + col0, sv[slot]);
        }
      };
      if (mode == MODE_F16_DACT) {
#pragma unroll
        for (int k = 0; k < PF; ++k) prefetch(k, k);
      }
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      int sh_n = 0, sh_i = 0, sh_j = 0;
      if (mode == MODE_F16_SHUFFLE) {
        sh_j = row % p.cv.OW;
        const int t3 = row / p.cv.OW;
        sh_i = t3 % p.cv.OH;
        sh_n = t3 / p.cv.OH;
      }
      const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN + cg * NCG);
      if (active) {
#pragma unroll 1
      for (int k0 = 0; k0 < NCH; k0 += PF) {
#pragma unroll
      for (int kk = 0; kk < PF; ++kk) {
        const int k = k0 + kk;
        uint32_t r[16];
        tmem_ld16(taddr0 + 16 * k, r);
        tmem_ld_wait();
        const int col0 = n_tile * BN + cg * NCG + 16 * k;
        if (row_ok && col0 < p.N) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
          const bool full = (col0 + 16 <= p.N);
          const int nvalid = full ? 16 : p.N - col0;
          if (mode == MODE_F32_ATOMIC) {
            float* out = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nvalid) atomicAdd(out + i, v[i]);
          } else if (mode == MODE_F32_STORE) {
            float* out = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nvalid) out[i] = v[i] + (p.bias ? p.bias[col0 + i] : 0.0f);
          } else if (mode == MODE_F16_SHUFFLE) {
            // dgrad of a strided conv: column block (py, px, c0..c0+15) of GEMM row (n, i, j)
            const int cls = col0 / p.sh.C, c0 = col0 % p.sh.C;
            const int y = p.sh.s * sh_i + cls / p.sh.s, x = p.sh.s * sh_j + cls % p.sh.s;
            if (y < p.sh.H && x < p.sh.W) {
              const long long o = (((long long)sh_n * p.sh.H + y) * p.sh.W + x) * p.sh.C + c0;
              if (p.saved) mask16(v, p.saved + o, true, 16, p.act);
              store16_f16(v, reinterpret_cast<__half*>(p.C) + o, true, 16);
            }
          } else {
            if (mode == MODE_F16_ACT) {
              if (p.bias) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += (i < nvalid) ? __ldg(p.bias + col0 + i) : 0.0f;
              }
              if (p.act == ACT_TANH) {
#pragma unroll 1
                for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
              } else {
                const float lo = (p.act == ACT_RELU) ? 0.0f : -INFINITY;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], lo);
              }
            } else {  // MODE_F16_DACT: dX = (dY W^T) * act'(saved activation)
              if (full && p.vec32) {
                if (use_bits) {                    // 1 bit per element: relu'(h) = (h > 0)
                  const uint32_t bw = sv[kk][0];
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = ((bw >> i) & 1u) ? v[i] : 0.0f;
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&sv[kk][i]));
                    v[2 * i] *= act_grad_from_saved(f.x, p.act);
                    v[2 * i + 1] *= act_grad_from_saved(f.y, p.act);
                  }
                }
              } else {
                mask16(v, p.saved + (long long)row * p.ld_saved + col0, false, nvalid, p.act);
              }
            }
            int ocol = col0;
            if (p.rm_C > 0) {
              const int pix = col0 / p.rm_C;
              ocol = ((pix / p.rm_OW) * p.rm_Wg + pix % p.rm_OW) * p.rm_C + col0 % p.rm_C;
            }
            store16_f16(v, reinterpret_cast<__half*>(p.C) + (long long)row * p.ldc + ocol,
                        full && p.vec32, nvalid);
          }
        }
        if (mode == MODE_F16_DACT && k + PF < NCH) prefetch(k + PF, kk);      // refill the slot just consumed
      }
      }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      aph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<C_::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void* driver_fn(const char* name) {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &qres);
  return (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) ? ptr : nullptr;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                          : CU_TENSOR_MAP_SWIZZLE_32B;
}

// 2-D fp16 tensor [rows, cols] with row pitch ld (elements); box = {box_cols (inner), box_rows}
static int make_tmap(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                     int box_rows) {
  static PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(driver_fn("cuTensorMapEncodeTiled"));
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return B200RL_ERR_DRIVER;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_cols * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d ptr=%p", (int)r, rows,
                   cols, ld, box_cols, box_rows, ptr);
    return B200RL_ERR_DRIVER;
  }
  return B200RL_OK;
}

// NHWC fp16 activation [B, H, W, C] read in im2col mode: `pixels` base pixels x C channels per load
static int make_tmap_im2col(CUtensorMap* tm, const void* ptr, long long B, int H, int W, int C, int lower_w,
                            int lower_h, int upper_w, int upper_h, int stride_w, int stride_h, int pixels) {
  static PFN_encodeIm2col enc = reinterpret_cast<PFN_encodeIm2col>(driver_fn("cuTensorMapEncodeIm2col"));
  if (!enc) {
    set_last_error("cuTensorMapEncodeIm2col entry point unavailable");
    return B200RL_ERR_DRIVER;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lo[2] = {lower_w, lower_h};
  int hi[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, lo, hi, (cuuint32_t)C,
                   (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(C * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeIm2col failed (%d): B=%lld H=%d W=%d C=%d lo=(%d,%d) hi=(%d,%d) st=(%d,%d) px=%d",
                   (int)r, B, H, W, C, lower_w, lower_h, upper_w, upper_h, stride_w, stride_h, pixels);
    return B200RL_ERR_DRIVER;
  }
  return B200RL_OK;
}

int make_tmap_2d_f16(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                     int box_rows) {
  return make_tmap(tm, ptr, rows, cols, ld, box_cols, box_rows);
}

static int g_num_sms = 0;
int device_num_sms();
static int num_sms() { return device_num_sms(); }
int device_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int CPT, bool MN, bool IM2COL, bool BRES = false, int TMODE = -1>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int max_ctas,
                  cudaStream_t stream) {
  using C_ = Cfg<BN>;
  constexpr int SMEM = BRES ? (BRES_STAGES * C_::A_BYTES + BRES_B_BYTES + 1024 + 256) : C_::SMEM_BYTES;
  static bool attr_set = false;
  auto kern = gemm_tcgen05_kernel<BN, CPT, MN, IM2COL, BRES, TMODE>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(smem=%d): %s", SMEM, cudaGetErrorString(e));
      return B200RL_ERR_CUDA;
    }
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.splits;
  int grid = total < num_sms() ? total : num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  kern<<<grid, NUM_THREADS, SMEM, stream>>>(tmA, tmB, p);
  return check_launch("gemm_tcgen05_kernel");
}

static void fill_splits(GemmParams& p, int split_k) {
  int splits = split_k < 1 ? 1 : split_k;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = ceil_div(p.kb_total, splits);
  p.splits = ceil_div(p.kb_total, p.kb_per_split);
}

// C-ABI body (declared in include/b200rl.h)
int gemm_f16_impl(const void* A, const void* B, void* C, const float* bias, const void* saved, int M, int N, int K,
                  long long lda, long long ldb, long long ldc, long long ld_saved, int mn_major, int mode, int act,
                  float alpha, int split_k, int max_ctas, int rm_C, int rm_OW, int rm_Wg, const void* saved_bits,
                  cudaStream_t stream) {
  B200RL_REQUIRE(A && B && C, "gemm: null operand");
  B200RL_REQUIRE(rm_C == 0 || (rm_C % 16 == 0 && rm_OW > 0 && rm_Wg >= rm_OW && (mode == MODE_F16_ACT || mode == MODE_F16_DACT)),
                 "gemm: bad column remap");
  B200RL_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  B200RL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 fp16 (16 B): %lld %lld", lda,
                 ldb);
  B200RL_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                 "gemm: operands must be 16-byte aligned");
  B200RL_REQUIRE(mode >= 0 && mode <= 3, "gemm: bad mode %d", mode);
  B200RL_REQUIRE(mode != MODE_F16_DACT || saved != nullptr || saved_bits != nullptr,
                 "gemm: MODE_F16_DACT needs the saved activation (or its bit mask)");
  B200RL_REQUIRE(saved_bits == nullptr || (mode == MODE_F16_DACT && act == ACT_RELU && (ld_saved % 16) == 0 && (N % 16) == 0),
                 "gemm: saved_bits is a ReLU mask for MODE_F16_DACT with ld_saved and N multiples of 16");

  int BN;
  // wgrad form: a 256-wide N tile halves the shared-memory traffic per MMA column (A is staged once per 256 instead of
  // per 128 output columns): fc1's weight gradient is bound by the 128 B/clk smem port, not by the tensor pipe
  if (mn_major) BN = (N > 128 && N % 256 == 0) ? 256 : (N > 64) ? 128 : 64;
  else BN = (N > 128 && (N % 256 == 0)) ? 256 : (N > 64) ? 128 : (N > 32) ? 64 : 32;

  GemmParams p = {};
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = ceil_div(M, BM);
  p.n_tiles = ceil_div(N, BN);
  p.kb_total = ceil_div(K, BK);
  B200RL_REQUIRE(split_k <= 1 || mode == MODE_F32_ATOMIC, "gemm: split_k needs the fp32 atomic epilogue");
  fill_splits(p, split_k);
  p.C = C; p.ldc = ldc; p.bias = bias; p.saved = reinterpret_cast<const __half*>(saved); p.ld_saved = ld_saved;
  p.saved_bits = reinterpret_cast<const uint16_t*>(saved_bits);
  p.alpha = alpha; p.mode = mode; p.act = act;
  p.rm_C = rm_C; p.rm_OW = rm_OW; p.rm_Wg = rm_Wg;
  p.vec32 = ((ldc & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 31) == 0) && (rm_C == 0 || (rm_C & 15) == 0) &&
            (!saved || (((ld_saved & 15) == 0) && ((reinterpret_cast<uintptr_t>(saved) & 31) == 0)));
  B200RL_REQUIRE(saved_bits == nullptr || p.vec32, "gemm: saved_bits needs 32-byte aligned 16-column output chunks");

  CUtensorMap tmA, tmB;
  int rc;
  if (!mn_major) {
    if ((rc = make_tmap(&tmA, A, M, K, lda, BK, BM)) != 0) return rc;
    if ((rc = make_tmap(&tmB, B, N, K, ldb, BK, BN)) != 0) return rc;
  } else {
    if ((rc = make_tmap(&tmA, A, K, M, lda, 64, BK)) != 0) return rc;
    if ((rc = make_tmap(&tmB, B, K, N, ldb, 64, BK)) != 0) return rc;
  }
  if (mn_major) {
    B200RL_REQUIRE(mode == MODE_F32_ATOMIC, "gemm: the MN-major (wgrad) layout uses the fp32 atomic epilogue");
    if (BN == 64) return launch<64, 64, true, false, false, MODE_F32_ATOMIC>(tmA, tmB, p, max_ctas, stream);
    if (BN == 256) return launch<256, 64, true, false, false, MODE_F32_ATOMIC>(tmA, tmB, p, max_ctas, stream);
    return launch<128, 64, true, false, false, MODE_F32_ATOMIC>(tmA, tmB, p, max_ctas, stream);
  }
#define GEMM_KMAJOR(bn)                                                                                         \
  switch (mode) {                                                                                               \
    case MODE_F16_ACT: return launch<bn, 64, false, false, false, MODE_F16_ACT>(tmA, tmB, p, max_ctas, stream);   \
    case MODE_F32_STORE: return launch<bn, 64, false, false, false, MODE_F32_STORE>(tmA, tmB, p, max_ctas, stream); \
    case MODE_F16_DACT: return launch<bn, 64, false, false, false, MODE_F16_DACT>(tmA, tmB, p, max_ctas, stream);  \
    default: return launch<bn, 64, false, false, false, MODE_F32_ATOMIC>(tmA, tmB, p, max_ctas, stream);          \
  }
  switch (BN) {
    case 32: GEMM_KMAJOR(32)
    case 64: GEMM_KMAJOR(64)
    case 128: GEMM_KMAJOR(128)
    default: GEMM_KMAJOR(256)
  }
#undef GEMM_KMAJOR
}

template <int CPT, bool MN>
static int launch_conv(int BN, bool bres, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                       cudaStream_t st) {
  if (MN) {
    switch (BN) {
      case 32: return launch<32, CPT, MN, true>(tmA, tmB, p, 0, st);
      case 64: return launch<64, CPT, MN, true>(tmA, tmB, p, 0, st);
      case 128: return launch<128, CPT, MN, true>(tmA, tmB, p, 0, st);
      default: break;
    }
  } else if (bres) {
    switch (BN) {
      case 32: return launch<32, CPT, false, true, !MN>(tmA, tmB, p, 0, st);
      case 64: return launch<64, CPT, false, true, !MN>(tmA, tmB, p, 0, st);
      case 128: return launch<128, CPT, false, true, !MN>(tmA, tmB, p, 0, st);
      default: break;
    }
  } else {
    switch (BN) {
      case 32: return launch<32, CPT, false, true>(tmA, tmB, p, 0, st);
      case 64: return launch<64, CPT, false, true>(tmA, tmB, p, 0, st);
      case 128: return launch<128, CPT, false, true>(tmA, tmB, p, 0, st);
      default: break;
    }
  }
  set_last_error("conv_gemm: unsupported tile N=%d", BN);
  return B200RL_ERR_UNSUPPORTED;
}

// Implicit-GEMM convolution on an NHWC fp16 tensor x[B, H, W, C] whose filter window is described by
// (R x S taps, stride, lower padding).  kind: 0 = forward / dgrad-style (rows = base pixels, K = taps*C,
// Wt = [N, taps*C] K-major), 1 = wgrad (out[taps*C, N] += x_patches^T * dz, dz = [B*OH*OW, N]).
int conv_gemm_impl(const void* x, long long B, int H, int W, int C, int R, int S, int stride_h, int stride_w,
                   int pad_h, int pad_w, int OH, int OW, const void* Wt_or_dz, long long ldb, void* out, long long ldc,
                   const float* bias, const void* saved, long long ld_saved, int N, int kind, int mode, int act,
                   float alpha, int split_k, int sh_H, int sh_W, int sh_C, int sh_s, cudaStream_t stream) {
  B200RL_REQUIRE(x && Wt_or_dz && out && B > 0, "conv_gemm: null operand");
  B200RL_REQUIRE(C == 16 || C == 32 || C == 64, "conv_gemm: channels per tap must be 16, 32 or 64 (got %d)", C);
  B200RL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldb % 8) == 0, "conv_gemm: alignment");
  B200RL_REQUIRE(kind == 0 || kind == 1, "conv_gemm: bad kind");
  const int taps = R * S;
  const long long rows = B * OH * OW;
  B200RL_REQUIRE(rows < (1LL << 31), "conv_gemm: too many rows");
  // bounding box of base pixels: lower = -pad, upper chosen so that exactly OW x OH base pixels exist:
  //   OW = (W + upper_w - lower_w - 1) / stride_w + 1   (cute::make_im2col_tma_copy_desc convention)
  const int lower_w = -pad_w, lower_h = -pad_h;
  const int upper_w = (OW - 1) * stride_w + 1 + lower_w - W;
  const int upper_h = (OH - 1) * stride_h + 1 + lower_h - H;

  GemmParams p = {};
  p.C = out; p.ldc = ldc; p.bias = bias; p.saved = reinterpret_cast<const __half*>(saved); p.ld_saved = ld_saved;
  p.alpha = alpha; p.mode = mode; p.act = act;
  p.cv.OW = OW; p.cv.OH = OH; p.cv.stride_w = stride_w; p.cv.stride_h = stride_h;
  p.cv.lower_w = lower_w; p.cv.lower_h = lower_h; p.cv.S = S; p.cv.taps = taps;
  p.sh.H = sh_H; p.sh.W = sh_W; p.sh.C = sh_C; p.sh.s = sh_s;
  p.vec32 = ((ldc & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 31) == 0) &&
            (!saved || (((ld_saved & 15) == 0) && ((reinterpret_cast<uintptr_t>(saved) & 31) == 0)));
  if (mode == MODE_F16_SHUFFLE)
    B200RL_REQUIRE((sh_C & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0 &&
                       (!saved || (reinterpret_cast<uintptr_t>(saved) & 31) == 0),
                   "conv_gemm: shuffle epilogue needs C %% 16 == 0 and 32-byte aligned tensors");
  CUtensorMap tmA, tmB;
  int rc;
  if (kind == 0) {
    B200RL_REQUIRE(mode == MODE_F16_ACT || mode == MODE_F16_DACT || mode == MODE_F16_SHUFFLE, "conv_gemm: bad mode");
    B200RL_REQUIRE(mode != MODE_F16_SHUFFLE || (sh_C % 16 == 0 && sh_s >= 1 && N == sh_s * sh_s * sh_C),
                   "conv_gemm: shuffle epilogue needs N == s*s*C and C %% 16 == 0");
    const int BN = (N > 64) ? 128 : (N > 32) ? 64 : 32;
    p.M = (int)rows; p.N = N; p.K = taps * C;
    p.m_tiles = ceil_div(p.M, BM);
    p.n_tiles = ceil_div(N, BN);
    p.kb_total = ceil_div(taps, BK / C);
    fill_splits(p, 1);
    if ((rc = make_tmap_im2col(&tmA, x, B, H, W, C, lower_w, lower_h, upper_w, upper_h, stride_w, stride_h, BM)) != 0)
      return rc;
    if ((rc = make_tmap(&tmB, Wt_or_dz, N, (long long)taps * C, ldb, C, BN)) != 0) return rc;
    // weights resident in smem when the whole [N, taps*C] matrix fits next to the A ring
    const bool bres = (p.n_tiles == 1) && ((long long)taps * BN * C * 2 <= BRES_B_BYTES) && (split_k != -1);
    if (C == 64) return launch_conv<64, false>(BN, bres, tmA, tmB, p, stream);
    if (C == 32) return launch_conv<32, false>(BN, bres, tmA, tmB, p, stream);
    return launch_conv<16, false>(BN, bres, tmA, tmB, p, stream);
  }
  B200RL_REQUIRE(mode == MODE_F32_ATOMIC, "conv_gemm: wgrad needs the fp32 atomic epilogue");
  const int BN = (N > 64) ? 128 : (N > 32) ? 64 : 32;
  p.M = taps * C; p.N = N; p.K = (int)rows;
  p.m_tiles = ceil_div(p.M, BM);
  p.n_tiles = ceil_div(N, BN);
  p.kb_total = ceil_div(p.K, BK);
  fill_splits(p, split_k);
  if ((rc = make_tmap_im2col(&tmA, x, B, H, W, C, lower_w, lower_h, upper_w, upper_h, stride_w, stride_h, BK)) != 0)
    return rc;
  if ((rc = make_tmap(&tmB, Wt_or_dz, rows, N, ldb, BN < 64 ? BN : 64, BK)) != 0) return rc;
  if (C == 64) return launch_conv<64, true>(BN, false, tmA, tmB, p, stream);
  if (C == 32) return launch_conv<32, true>(BN, false, tmA, tmB, p, stream);
  return launch_conv<16, true>(BN, false, tmA, tmB, p, stream);
}

}  // namespace b200rl
