// Micro-benchmarks that size the fused rollout kernel's design (tools/, not product): tcgen05.ld / tcgen05.st throughput,
// tcgen05.mma issue-to-retire rate for TS-mode (A in TMEM) and SS-mode (A in smem) at several N, MMA + concurrent epilogue
// loads, and MUFU tanh throughput.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench_tcgen05.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { for (uint32_t it = 0; it < (1u << 24); ++it) if (mbar_try_wait(bar, parity)) return; __trap(); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ inline uint32_t make_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

#define LD32(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
    "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory")
#define ST16(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
  ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")

// mode: 0 = ld only (nwarps loaders), 1 = st only, 2 = TS mma only, 3 = SS mma only, 4 = TS mma + loaders, 5 = SS mma + loaders,
//       6 = mufu tanh bf16x2, 7 = mufu tanh f32, 8 = mufu ex2 f32, 9 = TS mma + loaders + storers(half of the warps store)
__global__ void __launch_bounds__(320, 1) ubench(int mode, int nload_warps, int n, int iters, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 7);
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base;
  const bool do_mma = mode == 2 || mode == 3 || mode == 4 || mode == 5 || mode == 9 || mode == 10 || mode == 11;
  const int rot = (mode == 10 || mode == 11) ? nload_warps : 1;
  const bool ss = mode == 3 || mode == 5 || mode == 11;
  const bool do_ld = mode == 0 || mode == 4 || mode == 5 || mode == 9;
  uint32_t t0 = 0, t1 = 0;
  if (warp == 9 && do_mma) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(n);
      const uint64_t bdesc = make_desc(smem_u32(smem), 128u, 17 * 128u * 2);   // K-major no-swizzle, kp = 272
      const uint64_t adesc = make_desc(smem_u32(smem + 100 * 1024), 128u, 17 * 128u * 2);
      t0 = clock();
      for (int it = 0; it < iters; ++it) {
        const uint32_t koff = (uint32_t)(it % 16);
        const uint32_t dcol = tmem + 256 + (uint32_t)((it % rot) * n);
        if (ss) mma_ss(dcol, adesc + koff * 16, bdesc + koff * 16, idesc, it >= rot);
        else mma_ts(dcol, tmem + koff * 8, bdesc + koff * 16, idesc, it >= rot);
      }
      tc_commit(&bar);
      mbar_wait(&bar, 0);
      t1 = clock();
      out[0] = t1 - t0;
    }
  } else if (warp < 8) {
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    if (mode == 1 || (mode == 9 && warp >= 4)) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = lane + j;
      __syncwarp();
      t0 = clock();
      for (int it = 0; it < iters; ++it) { ST16(lane_base + 128 + (uint32_t)((warp >> 2) * 64) + (uint32_t)((it & 3) * 16), r); }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      t1 = clock();
      if (lane == 0) out[8 + warp] = t1 - t0;
    } else if (do_ld && warp < nload_warps) {
      uint32_t r[32]; uint32_t acc = 0;
      t0 = clock();
      for (int it = 0; it < iters; ++it) {
        LD32(lane_base + (uint32_t)((warp >> 2) * 64) + (uint32_t)((it & 1) * 32), r);
        if ((it & 3) == 3) { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[0] ^ r[31]; }
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      t1 = clock();
      if (lane == 0) { out[8 + warp] = t1 - t0; out[24 + warp] = acc; }
    } else if (mode >= 12 && mode <= 15) {
      uint32_t x[8]; float f[8];
      for (int j = 0; j < 8; ++j) { x[j] = 0x3C003C00u + threadIdx.x + j; f[j] = 0.001f * (threadIdx.x + j); }
      t0 = clock();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (mode == 12) { asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(x[j])); }
          else if (mode == 13) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j])); }
          else if (mode == 14) { asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(x[j])); }
          else { asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[j])); }
        }
      }
      t1 = clock();
      uint32_t acc = 0; for (int j = 0; j < 8; ++j) acc += x[j] + __float_as_uint(f[j]);
      if (lane == 0) { out[8 + warp] = t1 - t0; out[24 + warp] = acc; }
    } else if (mode >= 6 && mode <= 8) {
      uint32_t x = 0x3C003C00u + threadIdx.x; float f = 0.001f * threadIdx.x;
      t0 = clock();
#pragma unroll 8
      for (int it = 0; it < iters; ++it) {
        if (mode == 6) { asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(x)); }
        else if (mode == 7) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f)); }
        else { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); }
      }
      t1 = clock();
      if (lane == 0) { out[8 + warp] = t1 - t0; out[24 + warp] = x + __float_as_uint(f); }
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  uint32_t* d; cudaMalloc(&d, 256); uint32_t h[64];
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  auto run = [&](const char* name, int mode, int nw, int n, int iters) {
    cudaMemset(d, 0, 256);
    ubench<<<1, 320, 200 * 1024>>>(mode, nw, n, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
    printf("%-34s mode=%d nw=%d n=%3d iters=%d | mma cyc/iter %.1f | warp cyc/iter:", name, mode, nw, n, iters, h[0] / (double)iters);
    for (int w = 0; w < 8; ++w) printf(" %.1f", h[8 + w] / (double)iters);
    printf("\n");
  };
  if (getenv("UB2")) {
    for (int n : {16, 32, 64}) for (int r : {1, 2, 4}) run("mma TS rot accumulators (nw=rot)", 10, r, n, 2000);
    for (int n : {96, 112, 128}) for (int r : {1, 2}) run("mma TS rot accumulators (nw=rot)", 10, r, n, 2000);
    for (int n : {16, 64}) for (int r : {1, 4}) run("mma SS rot accumulators (nw=rot)", 11, r, n, 2000);
    run("mufu tanh.bf16x2 x8 indep /iter", 12, 8, 0, 1024);
    run("mufu ex2.f32 x8 indep /iter", 13, 8, 0, 1024);
    run("hfma2.bf16 x8 indep /iter", 14, 8, 0, 1024);
    run("ffma x8 indep /iter", 15, 8, 0, 1024);
    return 0;
  }
  for (int nw : {1, 4, 8}) run("tcgen05.ld 32x32b.x32 (4KB/warp-op)", 0, nw, 0, 2000);
  run("tcgen05.st 32x32b.x16 (2KB/warp-op)", 1, 8, 0, 2000);
  for (int n : {16, 64, 128, 208, 256}) run("mma TS (A in TMEM)", 2, 0, n, 2000);
  for (int n : {16, 64, 128, 208, 256}) run("mma SS (A in smem)", 3, 0, n, 2000);
  for (int n : {64, 128, 256}) run("mma TS + 8 ld warps", 4, 8, n, 2000);
  for (int n : {64, 128, 256}) run("mma TS + 4 ld warps", 4, 4, n, 2000);
  for (int n : {128, 256}) run("mma SS + 8 ld warps", 5, 8, n, 2000);
  for (int n : {128, 256}) run("mma TS + 4 ld + 4 st warps", 9, 4, n, 2000);
  run("mufu tanh.bf16x2 (8 warps)", 6, 8, 0, 4096);
  run("mufu tanh.f32 (8 warps)", 7, 8, 0, 4096);
  run("mufu ex2.f32 (8 warps)", 8, 8, 0, 4096);
  return 0;
}
