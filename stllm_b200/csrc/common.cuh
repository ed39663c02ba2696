// Common device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM inline PTX.
// Everything here is hand-written PTX (no CUTLASS/CuTe dependency).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define DEVI __device__ __forceinline__

// ----------------------------------------------------------------------------------------------
// error codes of the C ABI (include/stllm_b200.h)
// ----------------------------------------------------------------------------------------------
#define STLLM_OK 0
#define STLLM_ERR_BAD_ARG (-1)
#define STLLM_ERR_CUDA (-2)
#define STLLM_ERR_UNSUPPORTED (-3)
#define STLLM_ERR_DRIVER (-4)

#define STLLM_CHECK_ARG(cond)                                                               \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      stllm_set_last_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);           \
      return STLLM_ERR_BAD_ARG;                                                             \
    }                                                                                       \
  } while (0)

#define STLLM_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      stllm_set_last_error("%s:%d: CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e,  \
                           cudaGetErrorString(_e), #expr);                                  \
      return STLLM_ERR_CUDA;                                                                \
    }                                                                                       \
  } while (0)

void stllm_set_last_error(const char* fmt, ...);

// ----------------------------------------------------------------------------------------------
// shared-memory addresses
// ----------------------------------------------------------------------------------------------
DEVI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
DEVI void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DEVI void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DEVI void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

DEVI void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
DEVI void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEVI uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must trap (-> CUDA error on the host) instead of hanging the GPU.
// Returns the number of cycles spent waiting (0 if the phase had already completed).
// hot-loop form: no clock reads, no timeout trap (the MMA issuer shares its scheduler with two epilogue warps; every
// instruction of its k-block loop delays the next tcgen05.mma when those warps are busy)
DEVI void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// returns the cycles spent waiting INCLUDING the first try: mbarrier.try_wait suspends the thread in hardware for a
// while before it reports failure, so a count that starts after the first failed try misses most of a short wait
DEVI long long mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  if (mbar_try_wait(bar, parity)) return clock64() - t0;
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("stllm: mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
  return clock64() - t0;
}

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D tiled loads into swizzled shared memory
// ----------------------------------------------------------------------------------------------
DEVI void tma_prefetch_desc(const CUtensorMap* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
DEVI void tma_load_2d(const CUtensorMap* desc, uint64_t* bar, void* smem_dst, int32_t c_inner, int32_t c_outer,
                      uint64_t cache_policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c_inner),
        "r"(c_outer), "l"(cache_policy)
      : "memory");
}
// L2 cache-policy constants (same encodings CUTLASS uses for TMA::CacheHintSm90)
#define TMA_EVICT_NORMAL 0x1000000000000000ull
#define TMA_EVICT_FIRST 0x12F0000000000000ull
#define TMA_EVICT_LAST 0x14F0000000000000ull

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
DEVI void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DEVI void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *smem_dst
DEVI void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
DEVI void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], fp16/bf16 inputs, fp32 accumulate. One thread issues.
DEVI void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
DEVI void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// SMEM matrix descriptor, K-major operand, 128-byte swizzle, rows of 64 fp16 (=128 B), 8-row groups
// 1024 B apart (the layout a TMA box {64, rows} with CU_TENSOR_MAP_SWIZZLE_128B produces).
// Bit layout: cute::UMMA::SmemDescriptor (start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout_type [61,64) with SWIZZLE_128B = 2).
DEVI uint64_t make_sdesc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO (ignored for swizzled K-major), canonical value 1
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: 8 rows * 128 B
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}

// the same descriptor split into its address-dependent low word and its constant high word: the MMA issuers keep the
// low word of the current pipeline stage in a register and add constants per operand / k-step
DEVI uint32_t sdesc_lo_k_sw128(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t SDESC_HI_K_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
DEVI uint64_t sdesc_join(uint32_t lo) { return ((uint64_t)SDESC_HI_K_SW128 << 32) | (uint64_t)lo; }

// Instruction descriptor for kind::f16: fp16 A/B (K-major both), fp32 D.
// Bit layout: cute::UMMA::InstrDescriptor (c_format [4,6) | a_format [7,10) | b_format [10,13) |
// a_major 15 | b_major 16 | N>>3 [17,23) | M>>4 [24,29)).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, bool bf16) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (base_lane + t), columns [col, col+32)
DEVI void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
DEVI void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One elected lane of a fully converged warp (elect.sync, deterministic: always the same lane for the same mask).
// The single-thread tcgen05.mma / tcgen05.commit / TMA issue sites use `if (elect_one()) ...` inside WARP-UNIFORM control
// flow instead of `if (lane == 0) { whole loop }`: with a divergent lane-0 region ptxas cannot prove the operands
// warp-uniform and wraps every UTCHMMA / UTCBAR / UTMALDG in a "waterfall" loop (ELECT + 5 x R2UR.BROADCAST + BRA.U.ANY)
// that costs ~87 cycles per MMA issue -- measured on B200 (scripts/ubench/mma_pace.cu, profiles/r02_mma_pace.log): a
// K = 64 k-block of four N = 128 MMAs takes 349 cycles issued from a lane-0 region, 256 (= the tensor pipe's time) issued
// through elect_one() from uniform registers.
DEVI uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}"
      : "+r"(pred));
  return pred;
}

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
DEVI uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
DEVI float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DEVI float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: cluster of 2 CTAs on one TPC, leader = cluster rank 0
// ----------------------------------------------------------------------------------------------
DEVI uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DEVI void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `cta` of the cluster
DEVI uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
// Arrivals that only hand a TMEM accumulator stage back to the MMA issuer carry no memory payload: the tcgen05.ld
// results are in registers (tcgen05.wait::ld) and tcgen05.fence::before_thread_sync orders the async proxy.  The default
// .release form compiles to MEMBAR.ALL.GPU + ERRBAR at cluster scope, i.e. it waits until every global store of the
// epilogue has been acknowledged by L2 (6-10 % of the epilogue warps' stall samples, profiles/r02_epilogue.md).
DEVI void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEVI void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
DEVI float4 lds_f4(uint32_t smem_addr) {
  float4 f;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(smem_addr));
  return f;
}
// TMA load issued by either CTA of a pair; the transaction bytes are credited to the LEADER CTA's mbarrier
// (peer bit 24 of the shared::cluster address cleared), as cute::SM100_TMA_2SM_LOAD_2D does.
DEVI void tma_load_2d_2sm(const CUtensorMap* desc, uint64_t* bar, void* smem_dst, int32_t c_inner, int32_t c_outer,
                          uint64_t cache_policy) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(mbar), "r"(c_inner), "r"(c_outer),
        "l"(cache_policy)
      : "memory");
}
DEVI void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
DEVI void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows per CTA) * B (N/2 rows per CTA); issued by ONE thread of the leader CTA
DEVI void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread are done) on the mbarrier at this offset in every CTA of `cta_mask`
DEVI void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor is still running; it must execute pdl_wait() before reading or writing any global
// memory the predecessor touches.  pdl_launch_dependents() lets the NEXT kernel begin its launch/prologue early.
// Both are no-ops for kernels launched without the attribute.
// ----------------------------------------------------------------------------------------------
DEVI void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
DEVI void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool stllm_pdl_enabled();
// STLLM_CARVEOUT=1 (experiment, default off, not yet A/B-ed on the device): every kernel launched through stllm_launch is
// pinned once to the maximum shared-memory carve-out, so that consecutive kernels of a step (GEMM ~200 KB, attention
// 2 x 96 KB, element-wise kernels 0 KB of shared memory) do not make the driver switch the L1 / shared split of the SMs
// between launches.  Returns true the first time it sees `fn` (and only when the experiment is enabled).
bool stllm_carveout_pin_once(const void* fn);

template <typename... KArgs, typename... Args>
cudaError_t stllm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                         Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = stllm_pdl_enabled() ? 1 : 0;
  if (stllm_carveout_pin_once(reinterpret_cast<const void*>(kernel)))
    (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
