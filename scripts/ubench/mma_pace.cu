// Microbenchmark (B200): how many SM cycles does ONE tcgen05.mma (kind::f16, K = 16) cost, as a function of cta_group,
// N, the commit cadence and -- the point of round 2 -- HOW THE ISSUING THREAD IS SELECTED:
//   STYLE 0: `if (lane == 0) { ... tcgen05.mma ... }`  (round-1 GEMM).  ptxas cannot prove the operands warp-uniform and
//            wraps every UTCHMMA / UTCBAR in a "waterfall" loop (ELECT + 5 x R2UR.BROADCAST + BRA.U.ANY).
//   STYLE 1: the whole warp runs the loop, `if (elect_one_sync()) tcgen05.mma` -- operands live in uniform registers,
//            UTCHMMA is issued straight from the uniform datapath.
// One warp issues R k-blocks of 4 MMAs over a ring of 4 zero-filled shared-memory stages (no TMA, no epilogue),
// commits, waits; cycles / MMA = (clock after the final wait - clock before the first issue) / (4 R).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_pace mma_pace.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../stllm_b200/csrc/common.cuh"

void stllm_set_last_error(const char*, ...) {}
bool stllm_pdl_enabled() { return false; }
bool stllm_carveout_pin_once(const void*) { return false; }

DEVI uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred;
}

struct Cfg {
  int M, N;          // instruction shape (M = rows over the whole CTA group)
  int kblocks;       // k-blocks of 4 MMAs
  int commit_every;  // tcgen05.commit to a dummy barrier every n k-blocks (0 = never)
  int alt_acc;       // 1: alternate between two accumulators every MMA (independent accumulation chains)
};

template <int CG>
DEVI void issue_kblock(uint32_t st, uint32_t d0, uint32_t d1, uint32_t idesc, bool first, int alt) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint64_t da = make_sdesc_k_sw128(st + kk * 32), db = make_sdesc_k_sw128(st + 16384 + kk * 32);
    const uint32_t d = (alt && (kk & 1)) ? d1 : d0;
    const uint32_t acc = (first && kk < (alt ? 2 : 1)) ? 0u : 1u;
    if (CG == 2) umma_f16_ss_2sm(d, da, db, idesc, acc); else umma_f16_ss(d, da, db, idesc, acc);
  }
}

template <int CG, int STYLE>
__global__ void __launch_bounds__(128, 1) pace_kernel(Cfg c, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int STAGE = 32768;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 6 * STAGE);
  uint64_t* dummy = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 6 * STAGE / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(dummy, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2sm(slot, 512); else tmem_alloc(slot, 512);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  const bool leader = (CG == 1) || cluster_ctarank() == 0;
  const uint32_t idesc = make_idesc_f16(c.M, c.N, false);
  const uint32_t s0 = smem_u32(smem);
  if (warp == 0 && leader) {
    long long t0 = 0, t_issue = 0, t1 = 0;
    if (STYLE == 0) {
      if (lane == 0) {
        t0 = clock64();
        for (int i = 0; i < c.kblocks; ++i) {
          issue_kblock<CG>(s0 + (i & 3) * STAGE, tmem_base, tmem_base + 256, idesc, i == 0, c.alt_acc);
          if (c.commit_every == 1 || (c.commit_every == 2 && (i & 1))) {
            if (CG == 2) umma_commit_2sm(dummy, 1); else umma_commit(dummy);
          }
        }
        t_issue = clock64();
        if (CG == 2) umma_commit_2sm(bar, 1); else umma_commit(bar);
        mbar_wait(bar, 0);
        t1 = clock64();
      }
    } else {
      t0 = clock64();
      for (int i = 0; i < c.kblocks; ++i) {
        if (elect_one()) issue_kblock<CG>(s0 + (i & 3) * STAGE, tmem_base, tmem_base + 256, idesc, i == 0, c.alt_acc);
        if (c.commit_every == 1 || (c.commit_every == 2 && (i & 1))) {
          if (elect_one()) { if (CG == 2) umma_commit_2sm(dummy, 1); else umma_commit(dummy); }
        }
        __syncwarp();
      }
      t_issue = clock64();
      if (elect_one()) { if (CG == 2) umma_commit_2sm(bar, 1); else umma_commit(bar); }
      __syncwarp();
      mbar_wait(bar, 0);
      t1 = clock64();
    }
    if (lane == 0) {
      out[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
      out[blockIdx.x * 2 + 1] = (unsigned long long)(t_issue - t0);
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

template <int CG, int STYLE>
void run(const Cfg& c, int ctas, unsigned long long* d_out) {
  const int smem = 6 * 32768 + 2048;
  auto kern = pace_kernel<CG, STYLE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms = 0;
  for (int it = 0; it < 3; ++it) {
    cudaMemset(d_out, 0, 160 * 2 * sizeof(unsigned long long));
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, c, d_out);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      exit(1);
    }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  std::vector<unsigned long long> h(160 * 2);
  cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost);
  unsigned long long mx = 0, iss = 0;
  for (int b = 0; b < ctas; b += CG)
    if (h[b * 2] > mx) { mx = h[b * 2]; iss = h[b * 2 + 1]; }
  const double n_mma = 4.0 * c.kblocks;
  const double per = (double)mx / n_mma;
  const double macs = (double)c.M / CG * c.N * 16;   // per SM and MMA
  printf("%s cg%d M=%3d N=%3d alt=%d commit/%d kblk ctas=%3d : %6.1f cyc/MMA = %6.1f cyc/k-block (issue alone %6.1f)  "
         "%5.0f MAC/cyc/SM (%4.1f %% of 4096)  => %.0f MHz\n",
         STYLE ? "elect" : "lane0", CG, c.M, c.N, c.alt_acc, c.commit_every, ctas, per, per * 4, (double)iss / n_mma,
         macs / per, macs / per / 40.96, (double)mx / (ms * 1e3));
}

template <int CG>
void sweep(int ctas, unsigned long long* d_out) {
  const int R = 2048, M = 128 * CG;
  for (int n : {64, 128, 192, 256}) {
    run<CG, 0>(Cfg{M, n, R, 1, 0}, ctas, d_out);
    run<CG, 1>(Cfg{M, n, R, 1, 0}, ctas, d_out);
  }
  for (int n : {64, 128}) {
    run<CG, 1>(Cfg{M, n, R, 0, 0}, ctas, d_out);
    run<CG, 1>(Cfg{M, n, R, 2, 0}, ctas, d_out);
    run<CG, 1>(Cfg{M, n, R, 1, 1}, ctas, d_out);
    run<CG, 0>(Cfg{M, n, R, 0, 0}, ctas, d_out);
  }
  for (int n : {96, 160, 224}) run<CG, 1>(Cfg{M, n, R, 1, 0}, ctas, d_out);
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 160 * 2 * sizeof(unsigned long long));
  sweep<1>(1, d_out);
  sweep<1>(148, d_out);
  sweep<2>(2, d_out);
  sweep<2>(148, d_out);
  return 0;
}
