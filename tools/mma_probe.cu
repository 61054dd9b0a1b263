// Hardware probe (measurement tool, not product code): cycles per tcgen05.mma as a function of the operand source
// and N, with and without concurrent shared-memory LSU traffic, plus TMEM load bandwidth.  The numbers feed the
// "operand fetch" model in DESIGN.md that bounds the N <= 64 convolution kernels.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mma_probe tools/mma_probe.cu
//   tools/_build/mma_probe            (prints one JSON line per experiment)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../baselines_b200/csrc/common.cuh"
#include "../baselines_b200/csrc/tc_common.cuh"

using namespace b200rl;

struct Result {
  long long clk_mma;       // cycles for all MMAs (issue of first .. completion of last)
  long long lsu_bytes;     // bytes moved by the background LSU warps in that window (0 if none)
  long long clk_tmem;      // cycles of the TMEM load loop
};

// A from TMEM ("ts"): D[tmem] += A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// MODE 0: SS K-major (A [M rows x 128 B], B [N rows x 128 B], SWIZZLE_128B)
// MODE 1: TS (A in TMEM), B as above
// MODE 2: SS MN-major both (the wgrad form): A [16 K-rows x 128 M], B [16 K-rows x N]
template <int M, int N, int MODE, bool LSU, int NMMA, int AOFF = 0>
__global__ void __launch_bounds__(384, 1) probe_kernel(Result* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 32 KB
  uint8_t* sB = smem + 32 * 1024;     // 64 KB
  uint8_t* sL = smem + 96 * 1024;     // 64 KB LSU playground
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  __shared__ unsigned long long lsu_total;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    stop_flag = 0;
    lsu_total = 0;
  }
  if (warp == 2) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  long long clk_mma = 0, clk_tmem = 0;

  if (warp == 1) {
    constexpr uint32_t IDESC = (1u << 4) | ((MODE == 2 ? 1u : 0u) << 15) | ((MODE == 2 ? 1u : 0u) << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (elect_one()) {
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      const long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < NMMA; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (MODE == 0) {
            const uint64_t ad = make_sdesc(a_addr + AOFF * 128 + ((i >> 2) & 1) * 8192 + k * 32, 16, 1024, 2u);   // AOFF rows: the shifted-tap start
            const uint64_t bd = make_sdesc(b_addr + ((i >> 2) & 1) * 32768 + k * 32, 16, 1024, 2u);
            umma_f16(tmem_base, ad, bd, IDESC, 1);
          } else if (MODE == 1) {
            const uint64_t bd = make_sdesc(b_addr + ((i >> 2) & 1) * 32768 + k * 32, 16, 1024, 2u);
            umma_f16_ts(tmem_base, tmem_base + 256 + k * 8, bd, IDESC, 1);
          } else {
            const uint64_t ad = make_sdesc(a_addr + k * (16 * 128), 8192, 1024, 2u);
            const uint64_t bd = (N >= 64) ? make_sdesc(b_addr + k * (16 * 128), 8192, 1024, 2u)
                                          : make_sdesc(b_addr + k * (16 * 64), 64 * 64, 8 * 64, 4u);
            umma_f16(tmem_base, ad, bd, IDESC, 1);
          }
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      clk_mma = clock64() - t0;
      stop_flag = 1;
    }
    __syncwarp();
  } else if (LSU && warp >= 4) {
    // background shared-memory traffic: 16-byte loads + stores, conflict-free, until the MMA warp is done
    const uint32_t base = smem_u32(sL) + (warp - 4) * 8192 + lane * 16;
    unsigned long long bytes = 0;
    uint4 v = make_uint4(0, 0, 0, 0);
    while (!stop_flag) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 w;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(base + j * 512));
        v.x ^= w.x;
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + 4096 + j * 512), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      bytes += 8 * 2 * 512;
    }
    if (lane == 0) atomicAdd(&lsu_total, bytes);
  }
  __syncthreads();
  // TMEM load bandwidth: warps 4..11 (two per lane quadrant) each read 32 lanes x 32 columns per instruction
  if (warp >= 4) {
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; ++i) {
      uint32_t r[32], q[32];
      tmem_ld32(taddr + (i & 3) * 64, r);
      tmem_ld32(taddr + (i & 3) * 64 + 32, q);
      tmem_ld_wait();
      acc ^= r[0] ^ r[31] ^ q[0] ^ q[31];
    }
    const long long t1 = clock64();
    if (acc == 0x12345u) printf("x");
    if (warp == 4 && lane == 0) clk_tmem = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 32) {
    out[blockIdx.x].clk_mma = clk_mma;
    out[blockIdx.x].lsu_bytes = (long long)lsu_total;
  }
  if (threadIdx.x == 128) out[blockIdx.x].clk_tmem = clk_tmem;
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int M, int N, int MODE, bool LSU, int AOFF = 0>
static void run(const char* name, int ctas) {
  constexpr int NMMA = 4096;
  constexpr int SMEM = 161 * 1024 + 1024;
  auto kern = probe_kernel<M, N, MODE, LSU, NMMA, AOFF>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  Result* d;
  cudaMalloc(&d, sizeof(Result) * ctas);
  cudaMemset(d, 0, sizeof(Result) * ctas);
  kern<<<ctas, 384, SMEM>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("{\"probe\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e));
    cudaFree(d);
    return;
  }
  Result* h = new Result[ctas];
  cudaMemcpy(h, d, sizeof(Result) * ctas, cudaMemcpyDeviceToHost);
  double clk = 0, lsu = 0, tm = 0;
  for (int i = 0; i < ctas; ++i) { clk += h[i].clk_mma; lsu += h[i].lsu_bytes; tm += h[i].clk_tmem; }
  clk /= ctas; lsu /= ctas; tm /= ctas;
  const double per = clk / NMMA;
  const double a_bytes = (MODE == 1) ? 0.0 : M * 32.0, b_bytes = N * 32.0;
  const double floor_clk = (M < 128 ? 128 : M) * (double)N / 256.0;
  printf("{\"probe\": \"%s\", \"M\": %d, \"N\": %d, \"mode\": \"%s\", \"lsu_background\": %s, \"ctas\": %d, "
         "\"clk_per_mma\": %.2f, \"tensor_floor_clk\": %.1f, \"operand_bytes_per_mma\": %.0f, "
         "\"operand_B_per_clk\": %.1f, \"lsu_B_per_clk\": %.1f, \"tmem_ld_B_per_clk_8warps\": %.1f}\n",
         name, M, N, MODE == 0 ? "SS K-major" : MODE == 1 ? "TS (A in TMEM)" : "SS MN-major", LSU ? "true" : "false",
         ctas, per, floor_clk, a_bytes + b_bytes, (a_bytes + b_bytes) / per, lsu / clk,
         tm > 0 ? 8.0 * 256 * 32 * 32 * 4 / tm : 0.0);
  delete[] h;
  cudaFree(d);
}

int main(int argc, char** argv) {
  int ctas = 148;
  if (argc > 1) ctas = atoi(argv[1]);
  run<128, 32, 0, false>("ss_n32", ctas);
  run<128, 64, 0, false>("ss_n64", ctas);
  run<128, 128, 0, false>("ss_n128", ctas);
  run<128, 256, 0, false>("ss_n256", ctas);
  run<64, 64, 0, false>("ss_m64_n64", ctas);
  run<64, 256, 0, false>("ss_m64_n256", ctas);
  run<128, 32, 0, true>("ss_n32_lsu", ctas);
  run<128, 64, 0, true>("ss_n64_lsu", ctas);
  run<128, 128, 0, true>("ss_n128_lsu", ctas);
  run<128, 64, 1, false>("ts_n64", ctas);
  run<128, 128, 1, false>("ts_n128", ctas);
  run<128, 256, 1, false>("ts_n256", ctas);
  run<128, 256, 1, true>("ts_n256_lsu", ctas);
  run<128, 32, 2, false>("mn_n32", ctas);
  run<128, 64, 2, false>("mn_n64", ctas);
  run<128, 128, 2, false>("mn_n128", ctas);
  run<128, 32, 0, false>("ss_n32_1cta", 1);
  // A operand starting at a row that is not a multiple of 8 (the shifted filter taps of csrc/conv_shift.cu)
  run<128, 32, 0, false, 1>("ss_n32_aoff1", ctas);
  run<128, 32, 0, false, 21>("ss_n32_aoff21", ctas);
  run<128, 64, 0, false, 1>("ss_n64_aoff1", ctas);
  run<128, 64, 0, false, 21>("ss_n64_aoff21", ctas);
  run<128, 64, 0, false, 8>("ss_n64_aoff8", ctas);
  run<128, 128, 0, false, 10>("ss_n128_aoff10", ctas);
  run<128, 192, 0, false, 0>("ss_n192", ctas);
  run<128, 192, 0, false, 9>("ss_n192_aoff9", ctas);
  run<128, 256, 0, false, 5>("ss_n256_aoff5", ctas);
  return 0;
}
