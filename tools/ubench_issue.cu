// Micro-benchmark #2 (tools/, not product): how fast can tcgen05.mma be ISSUED?  Variants of the issue loop
// (lane==0 branch, elect.sync branch, whole-warp loop with predicated mma, unrolled), same vs rotating accumulators,
// one vs two issuing warps; plus MUFU / HFMA2 / FFMA throughput with independent chains.
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
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred;
}
// predicated TS-mode MMA: issue only when `go` != 0
__device__ __forceinline__ void mma_ts_p(uint32_t d, uint32_t a, uint32_t dlo, uint32_t dhi, uint32_t idesc, uint32_t acc, uint32_t go) {
  asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 dd;\n\tsetp.ne.u32 p, %5, 0;\n\tsetp.ne.u32 q, %6, 0;\n\tmov.b64 dd, {%2, %3};\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], dd, %4, p;\n\t}\n" ::"r"(d), "r"(a), "r"(dlo), "r"(dhi), "r"(idesc), "r"(acc), "r"(go) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint32_t dlo, uint32_t dhi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 dd;\n\tsetp.ne.u32 p, %5, 0;\n\tmov.b64 dd, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], dd, %4, p;\n\t}\n" ::"r"(d), "r"(a), "r"(dlo), "r"(dhi), "r"(idesc), "r"(acc) : "memory");
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

// variant: 0 lane==0 branch; 1 elect.sync branch; 2 whole-warp loop + predicated mma; 3 elect branch, 16 MMAs fully unrolled per outer iteration
// ROT: number of accumulators rotated through (compile time).  issuers: 1 or 2 warps (warp 8 and 9) issuing concurrently.
// queue-depth probe: issue `reps` x 16 MMAs back to back (unrolled), then stamp: after the issue loop, after commit, after completion
__global__ void __launch_bounds__(320, 1) queue_probe(int n, int reps, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 7);
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 8) {
    const uint32_t idesc = make_idesc(n);
    const uint64_t bdesc = make_desc(smem_u32(smem), 128u, 16 * 128u * 2);
    const uint32_t dhi = (uint32_t)(bdesc >> 32), dlo0 = (uint32_t)bdesc;
    if (elect_one()) {
      const uint32_t t0 = clock();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) mma_ts(tmem + 128, tmem + 8u * k, dlo0 + 16u * k, dhi, idesc, (r > 0 || k > 0) ? 1u : 0u);
      }
      const uint32_t t1 = clock();
      tc_commit(&bar[0]);
      const uint32_t t2 = clock();
      mbar_wait(&bar[0], 0);
      const uint32_t t3 = clock();
      out[0] = t1 - t0; out[1] = t2 - t0; out[2] = t3 - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int VARIANT, int ROT>
__global__ void __launch_bounds__(320, 1) issue_bench(int n, int outer, int issuers, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 7);
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp >= 8 && (warp - 8) < issuers) {
    const int wi = warp - 8;
    const uint32_t idesc = make_idesc(n);
    const uint64_t bdesc = make_desc(smem_u32(smem), 128u, 16 * 128u * 2);    // K-major no-swizzle, kp = 256
    const uint32_t dhi = (uint32_t)(bdesc >> 32), dlo0 = (uint32_t)bdesc;
    const uint32_t dbase = tmem + 128 + (uint32_t)wi * 192;                    // accumulators: [128,320) / [320,512)
    const uint32_t abase = tmem;                                               // A: cols [0,128) = K 256
    uint32_t t0 = 0, t1 = 0;
    if (VARIANT == 0) {
      if (lane == 0) {
        t0 = clock();
        for (int o = 0; o < outer; ++o) {
          uint32_t dlo = dlo0, a = abase;
#pragma unroll 4
          for (int k = 0; k < 16; ++k) { mma_ts(dbase + (uint32_t)((k % ROT) * n), a, dlo, dhi, idesc, (o > 0 || k >= ROT) ? 1u : 0u); dlo += 16u; a += 8u; }
        }
        tc_commit(&bar[wi]); mbar_wait(&bar[wi], 0); t1 = clock();
        out[wi] = t1 - t0;
      }
    } else if (VARIANT == 1 || VARIANT == 3) {
      if (elect_one()) {
        t0 = clock();
        for (int o = 0; o < outer; ++o) {
          uint32_t dlo = dlo0, a = abase;
          if (VARIANT == 3) {
#pragma unroll
            for (int k = 0; k < 16; ++k) mma_ts(dbase + (uint32_t)((k % ROT) * n), abase + 8u * k, dlo0 + 16u * k, dhi, idesc, (o > 0 || k >= ROT) ? 1u : 0u);
          } else {
#pragma unroll 4
            for (int k = 0; k < 16; ++k) { mma_ts(dbase + (uint32_t)((k % ROT) * n), a, dlo, dhi, idesc, (o > 0 || k >= ROT) ? 1u : 0u); dlo += 16u; a += 8u; }
          }
        }
        tc_commit(&bar[wi]); mbar_wait(&bar[wi], 0); t1 = clock();
        out[wi] = t1 - t0;
      }
    } else {
      const uint32_t go = elect_one();
      t0 = clock();
      for (int o = 0; o < outer; ++o) {
        uint32_t dlo = dlo0, a = abase;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) { mma_ts_p(dbase + (uint32_t)((k % ROT) * n), a, dlo, dhi, idesc, (o > 0 || k >= ROT) ? 1u : 0u, go); dlo += 16u; a += 8u; }
      }
      if (go) { tc_commit(&bar[wi]); mbar_wait(&bar[wi], 0); t1 = clock(); out[wi] = t1 - t0; }
      __syncwarp();
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int OP>
__global__ void __launch_bounds__(256) alu_bench(int iters, uint32_t* out) {
  uint32_t x0 = 0x3C003C00u + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  float f0 = 0.001f * threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
  const uint32_t t0 = clock();
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) {
      asm volatile("tanh.approx.bf16x2 %0, %0;\n\ttanh.approx.bf16x2 %1, %1;\n\ttanh.approx.bf16x2 %2, %2;\n\ttanh.approx.bf16x2 %3, %3;\n\t"
                   "tanh.approx.bf16x2 %4, %4;\n\ttanh.approx.bf16x2 %5, %5;\n\ttanh.approx.bf16x2 %6, %6;\n\ttanh.approx.bf16x2 %7, %7;"
                   : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7));
    } else if (OP == 1) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;\n\tex2.approx.ftz.f32 %1, %1;\n\tex2.approx.ftz.f32 %2, %2;\n\tex2.approx.ftz.f32 %3, %3;\n\t"
                   "ex2.approx.ftz.f32 %4, %4;\n\tex2.approx.ftz.f32 %5, %5;\n\tex2.approx.ftz.f32 %6, %6;\n\tex2.approx.ftz.f32 %7, %7;"
                   : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3), "+f"(f4), "+f"(f5), "+f"(f6), "+f"(f7));
    } else if (OP == 2) {
      asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;\n\tfma.rn.bf16x2 %1, %1, %1, %1;\n\tfma.rn.bf16x2 %2, %2, %2, %2;\n\tfma.rn.bf16x2 %3, %3, %3, %3;\n\t"
                   "fma.rn.bf16x2 %4, %4, %4, %4;\n\tfma.rn.bf16x2 %5, %5, %5, %5;\n\tfma.rn.bf16x2 %6, %6, %6, %6;\n\tfma.rn.bf16x2 %7, %7, %7, %7;"
                   : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7));
    } else if (OP == 3) {
      asm volatile("fma.rn.f32 %0, %0, %0, %0;\n\tfma.rn.f32 %1, %1, %1, %1;\n\tfma.rn.f32 %2, %2, %2, %2;\n\tfma.rn.f32 %3, %3, %3, %3;\n\t"
                   "fma.rn.f32 %4, %4, %4, %4;\n\tfma.rn.f32 %5, %5, %5, %5;\n\tfma.rn.f32 %6, %6, %6, %6;\n\tfma.rn.f32 %7, %7, %7, %7;"
                   : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3), "+f"(f4), "+f"(f5), "+f"(f6), "+f"(f7));
    } else if (OP == 4) {
      asm volatile("tanh.approx.f32 %0, %0;\n\ttanh.approx.f32 %1, %1;\n\ttanh.approx.f32 %2, %2;\n\ttanh.approx.f32 %3, %3;\n\t"
                   "tanh.approx.f32 %4, %4;\n\ttanh.approx.f32 %5, %5;\n\ttanh.approx.f32 %6, %6;\n\ttanh.approx.f32 %7, %7;"
                   : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3), "+f"(f4), "+f"(f5), "+f"(f6), "+f"(f7));
    } else {
      asm volatile("fma.rn.f16x2 %0, %0, %0, %0;\n\tfma.rn.f16x2 %1, %1, %1, %1;\n\tfma.rn.f16x2 %2, %2, %2, %2;\n\tfma.rn.f16x2 %3, %3, %3, %3;\n\t"
                   "fma.rn.f16x2 %4, %4, %4, %4;\n\tfma.rn.f16x2 %5, %5, %5, %5;\n\tfma.rn.f16x2 %6, %6, %6, %6;\n\tfma.rn.f16x2 %7, %7, %7, %7;"
                   : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7));
    }
  }
  const uint32_t t1 = clock();
  if ((threadIdx.x & 31) == 0) { out[threadIdx.x >> 5] = t1 - t0; out[16 + (threadIdx.x >> 5)] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + __float_as_uint(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7); }
}

uint32_t* d; uint32_t h[64];
template <int V, int R> void run_issue(int n, int issuers) {
  const int outer = 200;
  cudaFuncSetAttribute(issue_bench<V, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaMemset(d, 0, 256);
  issue_bench<V, R><<<1, 320, 64 * 1024>>>(n, outer, issuers, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
  printf("variant %d rot %d n=%3d issuers=%d : cycles/MMA  w8 %.1f  w9 %.1f\n", V, R, n, issuers, h[0] / (16.0 * outer), h[1] / (16.0 * outer));
}
template <int OP> void run_alu(const char* name) {
  cudaMemset(d, 0, 256);
  alu_bench<OP><<<1, 256>>>(2000, d);
  cudaDeviceSynchronize();
  cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
  printf("%-18s 8 warps x 8 indep ops/iter: %.1f cycles/iter (per SMSP: 2 warps x 8 warp-instr) -> %.2f cycles per warp-instr per SMSP\n", name, h[0] / 2000.0, h[0] / 2000.0 / 16.0);
}

int main() {
  cudaMalloc(&d, 256);
  cudaFuncSetAttribute(queue_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int n : {16, 64}) for (int reps : {1, 2, 4, 8}) {
    cudaMemset(d, 0, 256);
    queue_probe<<<1, 320, 64 * 1024>>>(n, reps, d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("probe failed\n"); return 1; }
    cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
    printf("queue probe n=%d: %d MMAs: issue loop done at %u cycles, commit returned at %u, completion seen at %u (exec floor %d)\n", n, 16 * reps, h[0], h[1], h[2], 16 * reps * n / 2);
  }
  if (getenv("PROBE_ONLY")) return 0;
  for (int n : {16, 64, 128}) { run_issue<0, 1>(n, 1); run_issue<1, 1>(n, 1); run_issue<2, 1>(n, 1); run_issue<3, 1>(n, 1); }
  for (int n : {16, 64}) { run_issue<1, 2>(n, 1); run_issue<3, 2>(n, 1); run_issue<3, 4>(n, 1); }
  for (int n : {16, 64, 128}) { run_issue<1, 1>(n, 2); run_issue<3, 1>(n, 2); }
  run_alu<0>("tanh.bf16x2"); run_alu<4>("tanh.f32"); run_alu<1>("ex2.f32"); run_alu<2>("fma.bf16x2"); run_alu<5>("fma.f16x2"); run_alu<3>("fma.f32");
  return 0;
}
