// Persistent warp-specialised tcgen05 GEMM for sm_100a (fp16 operands, fp32 accumulate in TMEM).
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1      : MMA issuer     (one lane issues tcgen05.mma.cta_group::1.kind::f16, commits to mbarriers)
//   warp 2      : TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps the MMAs of i+1)
//   warps 4..7  : epilogue       (tcgen05.ld 32x32b -> registers -> fused bias / activation / mask / atomics)
//
// Two operand layouts:
//   K-major  : A[M,K] row-major, B[N,K] row-major            C = A * B^T          (forward, dgrad)
//   MN-major : A[K,M] row-major, B[K,N] row-major            C = A^T * B          (wgrad; K = batch rows)
// plus split-K over the reduction dimension with an fp32 atomic epilogue (wgrad).
//
// This replaces the TF1 ops tf.matmul (a2c/utils.py:63) and -- after lowering by conv_lowering.cu --
// tf.nn.conv2d (a2c/utils.py:56) and their gradients (ppo2/model.py:102) of the reference.
#include <stdarg.h>
#include <stdio.h>

#include <mutex>

#include "common.cuh"

namespace b200rl {

static constexpr int BM = 128;
static constexpr int BK = 64;          // 64 fp16 = 128 bytes = one swizzle row
static constexpr int UMMA_K = 16;
static constexpr int NUM_THREADS = 256;

enum : int { MODE_F16_ACT = 0, MODE_F32_STORE = 1, MODE_F32_ATOMIC = 2, MODE_F16_DACT = 3 };
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  void* C;
  long long ldc;
  const float* bias;
  const __half* saved;     // saved activation for MODE_F16_DACT
  long long ld_saved;
  float alpha;
  int mode, act;
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

__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_RELU) return fmaxf(x, 0.0f);
  if (act == ACT_TANH) return tanhf(x);
  return x;
}
__device__ __forceinline__ float act_grad_from_saved(float h, int act) {
  if (act == ACT_RELU) return h > 0.0f ? 1.0f : 0.0f;
  if (act == ACT_TANH) return 1.0f - h * h;
  return 1.0f;
}

template <int BN, bool MN_MAJOR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using C_ = Cfg<BN>;
  constexpr int STAGES = C_::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C_::STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

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
      mbar_init(&tempty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int n_tile = work % p.n_tiles;
        const int t2 = work / p.n_tiles;
        const int m_tile = t2 % p.m_tiles;
        const int split = t2 / p.m_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * C_::STAGE_BYTES;
          uint8_t* sb = sa + C_::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], C_::STAGE_BYTES);
          if (!MN_MAJOR) {
            tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m_tile * BM);
            tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n_tile * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sa + j * (64 * BK * 2), &tmA, &full_bar[s], m_tile * BM + j * 64, kb * BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (64 * BK * 2), &tmB, &full_bar[s], n_tile * BN + j * 64, kb * BK);
          }
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
    if (lane == 0) {
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
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * C_::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C_::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            uint64_t adesc, bdesc;
            if (!MN_MAJOR) {
              adesc = make_sdesc(a_addr + k * (UMMA_K * 2), 16, 1024);
              bdesc = make_sdesc(b_addr + k * (UMMA_K * 2), 16, 1024);
            } else {
              adesc = make_sdesc(a_addr + k * (UMMA_K * 128), 64 * BK * 2, 1024);
              bdesc = make_sdesc(b_addr + k * (UMMA_K * 128), 64 * BK * 2, 1024);
            }
            umma_f16(tmem_d, adesc, bdesc, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);   // smem slot reusable once these MMAs retire
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[as]);    // accumulator ready for the epilogue
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;            // == warp % 4: TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aph = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      const int n_tile = work % p.n_tiles;
      const int t2 = work / p.n_tiles;
      const int m_tile = t2 % p.m_tiles;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row = m_tile * BM + ew * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr0 + c, r);
        tmem_ld_wait();
        const int col0 = n_tile * BN + c;
        if (row_ok && col0 < p.N) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
          const bool full = (col0 + 16 <= p.N);
          if (p.mode == MODE_F32_ATOMIC) {
            float* out = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (full || col0 + i < p.N) atomicAdd(out + i, v[i]);
          } else if (p.mode == MODE_F32_STORE) {
            float* out = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (full || col0 + i < p.N) out[i] = v[i] + (p.bias ? p.bias[col0 + i] : 0.0f);
          } else {
            if (p.mode == MODE_F16_ACT) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float b = (p.bias && (full || col0 + i < p.N)) ? __ldg(p.bias + col0 + i) : 0.0f;
                v[i] = apply_act(v[i] + b, p.act);
              }
            } else {  // MODE_F16_DACT: dX = (dY W^T) * act'(saved activation)
              const __half* sv = p.saved + (long long)row * p.ld_saved + col0;
              if (full && ((p.ld_saved & 7) == 0)) {
                uint4 q0 = *reinterpret_cast<const uint4*>(sv);
                uint4 q1 = *reinterpret_cast<const uint4*>(sv + 8);
                const __half2* h0 = reinterpret_cast<const __half2*>(&q0);
                const __half2* h1 = reinterpret_cast<const __half2*>(&q1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float2 f = __half22float2(h0[i]);
                  v[2 * i] *= act_grad_from_saved(f.x, p.act);
                  v[2 * i + 1] *= act_grad_from_saved(f.y, p.act);
                  float2 g = __half22float2(h1[i]);
                  v[8 + 2 * i] *= act_grad_from_saved(g.x, p.act);
                  v[8 + 2 * i + 1] *= act_grad_from_saved(g.y, p.act);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (col0 + i < p.N) v[i] *= act_grad_from_saved(__half2float(sv[i]), p.act);
              }
            }
            __half* out = reinterpret_cast<__half*>(p.C) + (long long)row * p.ldc + col0;
            if (full && ((p.ldc & 7) == 0)) {
              uint4 q0, q1;
              __half2* h0 = reinterpret_cast<__half2*>(&q0);
              __half2* h1 = reinterpret_cast<__half2*>(&q1);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                h0[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                h1[i] = __floats2half2_rn(v[8 + 2 * i], v[8 + 2 * i + 1]);
              }
              *reinterpret_cast<uint4*>(out) = q0;
              *reinterpret_cast<uint4*>(out + 8) = q1;
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col0 + i < p.N) out[i] = __float2half_rn(v[i]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
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

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

// 2-D fp16 tensor [rows, cols] with row pitch ld (elements); box = {box_cols (inner), box_rows}
static int make_tmap(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                     int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return B200RL_ERR_DRIVER;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d ptr=%p", (int)r, rows,
                   cols, ld, box_cols, box_rows, ptr);
    return B200RL_ERR_DRIVER;
  }
  return B200RL_OK;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, bool MN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int max_ctas,
                  cudaStream_t stream) {
  using C_ = Cfg<BN>;
  static bool attr_set = false;
  auto kern = gemm_tcgen05_kernel<BN, MN>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(smem=%d): %s", C_::SMEM_BYTES, cudaGetErrorString(e));
      return B200RL_ERR_CUDA;
    }
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.splits;
  int grid = total < num_sms() ? total : num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  kern<<<grid, NUM_THREADS, C_::SMEM_BYTES, stream>>>(tmA, tmB, p);
  return check_launch("gemm_tcgen05_kernel");
}

// C-ABI body (declared in include/b200rl.h)
int gemm_f16_impl(const void* A, const void* B, void* C, const float* bias, const void* saved, int M, int N, int K,
                  long long lda, long long ldb, long long ldc, long long ld_saved, int mn_major, int mode, int act,
                  float alpha, int split_k, int max_ctas, cudaStream_t stream) {
  B200RL_REQUIRE(A && B && C, "gemm: null operand");
  B200RL_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  B200RL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 fp16 (16 B): %lld %lld", lda,
                 ldb);
  B200RL_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                 "gemm: operands must be 16-byte aligned");
  B200RL_REQUIRE(mode >= 0 && mode <= 3, "gemm: bad mode %d", mode);
  B200RL_REQUIRE(mode != MODE_F16_DACT || saved != nullptr, "gemm: MODE_F16_DACT needs the saved activation");

  int BN;
  if (mn_major) BN = (N > 64) ? 128 : 64;
  else BN = (N > 128 && (N % 256 == 0)) ? 256 : (N > 64) ? 128 : (N > 32) ? 64 : 32;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = ceil_div(M, BM);
  p.n_tiles = ceil_div(N, BN);
  p.kb_total = ceil_div(K, BK);
  int splits = split_k < 1 ? 1 : split_k;
  if (splits > p.kb_total) splits = p.kb_total;
  B200RL_REQUIRE(splits == 1 || mode == MODE_F32_ATOMIC, "gemm: split_k needs the fp32 atomic epilogue");
  p.kb_per_split = ceil_div(p.kb_total, splits);
  p.splits = ceil_div(p.kb_total, p.kb_per_split);
  p.C = C; p.ldc = ldc; p.bias = bias; p.saved = reinterpret_cast<const __half*>(saved); p.ld_saved = ld_saved;
  p.alpha = alpha; p.mode = mode; p.act = act;

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
    if (BN == 64) return launch<64, true>(tmA, tmB, p, max_ctas, stream);
    return launch<128, true>(tmA, tmB, p, max_ctas, stream);
  }
  switch (BN) {
    case 32: return launch<32, false>(tmA, tmB, p, max_ctas, stream);
    case 64: return launch<64, false>(tmA, tmB, p, max_ctas, stream);
    case 128: return launch<128, false>(tmA, tmB, p, max_ctas, stream);
    default: return launch<256, false>(tmA, tmB, p, max_ctas, stream);
  }
}

}  // namespace b200rl
