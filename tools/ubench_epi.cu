// Micro-benchmark (not part of the library): what bounds a hidden-layer epilogue round of the rollout kernel
//   tcgen05.ld (2 x 16 columns) -> wait -> fp32 -> [ReLU | SiLU] -> packed bf16 -> st.shared (canonical K-major rows)
// as a function of the number of warps sharing the SM.  Variants isolate the pieces (load latency, conversion throughput, MUFU, stores).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_epi ubench_epi.cu && ./ubench_epi
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define LD16(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory")
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) { uint32_t r; asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t silu_bf16x2(uint32_t x) {
  uint32_t h, t, r; const uint32_t half2 = 0x3F003F00u;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(h) : "r"(x), "r"(half2));
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(h));
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(h), "r"(t), "r"(h));
  return r;
}
// round-half-up pack without the conversion instruction: add 0x8000 to each fp32 pattern, take the upper halves (PRMT)
__device__ __forceinline__ uint32_t pack_int(uint32_t lo, uint32_t hi) { return __byte_perm(lo + 0x8000u, hi + 0x8000u, 0x7632); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// variant: 0 ld+wait, 1 ld+relu pack, 2 ld+relu pack+store, 3 ld+silu, 4 ld+silu+store, 5 relu pack only (no ld), 6 int pack only, 7 silu only,
//          8 ld + int pack + max.bf16x2 + store, 9 ld + (half F2FP half int) relu + store
template <int variant>
__global__ void __launch_bounds__(512, 1) ubench(int rounds, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t region = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 96;
  const uint32_t xrow = smem_u32(smem) + (uint32_t)(((warp & 3) * 32 + lane) * 16) + (uint32_t)(warp >> 2) * 16384u;
  uint32_t ra[16], rb[16], acc = 0;
  for (int j = 0; j < 16; ++j) { ra[j] = __float_as_uint(0.01f * (threadIdx.x + j)); rb[j] = __float_as_uint(-0.02f * (threadIdx.x + j)); }
  __syncthreads();
  const uint32_t t0 = clock();
  for (int it = 0; it < rounds; ++it) {
    const int c = (it % 3) * 32;
    if (variant <= 4 || variant >= 8) { LD16(region + c, ra); LD16(region + c + 16, rb); ld_wait(); }
    if (variant == 0) { acc += ra[0] ^ rb[15]; continue; }
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (variant == 1 || variant == 2 || variant == 5) { pk[j] = pack_bf16_relu(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1])); pk[8 + j] = pack_bf16_relu(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1])); }
      else if (variant == 3 || variant == 4 || variant == 7) { pk[j] = silu_bf16x2(pack_bf16(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]))); pk[8 + j] = silu_bf16x2(pack_bf16(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]))); }
      else if (variant == 6) { pk[j] = pack_int(ra[2 * j], ra[2 * j + 1]); pk[8 + j] = pack_int(rb[2 * j], rb[2 * j + 1]); }
      else if (variant == 8) {
        uint32_t a = pack_int(ra[2 * j], ra[2 * j + 1]), b = pack_int(rb[2 * j], rb[2 * j + 1]); const uint32_t z = 0u;
        asm("max.bf16x2 %0, %1, %2;" : "=r"(pk[j]) : "r"(a), "r"(z)); asm("max.bf16x2 %0, %1, %2;" : "=r"(pk[8 + j]) : "r"(b), "r"(z));
      } else {
        pk[j] = pack_bf16_relu(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]));
        uint32_t b = pack_int(rb[2 * j], rb[2 * j + 1]); const uint32_t z = 0u;
        asm("max.bf16x2 %0, %1, %2;" : "=r"(pk[8 + j]) : "r"(b), "r"(z));
      }
    }
    if (variant == 2 || variant == 4 || variant == 8 || variant == 9) {
      const uint32_t a = xrow + (uint32_t)(c >> 3) * 2048u;
      st_shared_v4(a, pk[0], pk[1], pk[2], pk[3]); st_shared_v4(a + 2048u, pk[4], pk[5], pk[6], pk[7]);
      st_shared_v4(a + 4096u, pk[8], pk[9], pk[10], pk[11]); st_shared_v4(a + 6144u, pk[12], pk[13], pk[14], pk[15]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += pk[j];
      if (variant >= 5 && variant <= 7) { ra[0] ^= acc & 0x00010000u; rb[3] ^= acc & 0x00010000u; }     // keep the math live and varying
    }
  }
  const uint32_t t1 = clock();
  if (lane == 0) { out[warp] = t1 - t0; out[32 + warp] = acc; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  uint32_t* d; cudaMalloc(&d, 512); uint32_t h[64];
  const char* names[10] = {"ld2x16+wait", "ld + relu F2FP", "ld + relu F2FP + st.shared", "ld + silu", "ld + silu + st.shared", "relu F2FP only",
                           "int pack only", "silu only", "ld + int pack + max + st", "ld + half F2FP half int + st"};
  const int rounds = 3000;
  for (int v = 0; v < 10; ++v)
    for (int nw : {4, 8, 12, 16}) {
      cudaMemset(d, 0, 512);
      switch (v) {
#define CASE(V) case V: cudaFuncSetAttribute(ubench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); ubench<V><<<1, nw * 32, 128 * 1024>>>(rounds, d); break;
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", names[v], cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
      uint32_t mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("%-32s warps %2d (%d per scheduler): %7.1f cycles per 32-column round per warp ; %6.1f cycles per round per scheduler\n", names[v], nw, nw / 4,
             (double)mx / rounds, (double)mx / rounds / (nw / 4));
    }
  return 0;
}
