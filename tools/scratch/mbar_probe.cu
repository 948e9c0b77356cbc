// Micro-probe: what does the 64-bit mbarrier object look like before / after phase completions (count = 1, with and without tx)?
#include <cstdio>
#include <cstdint>
__global__ void k() {
  __shared__ uint64_t bar;
  uint32_t a = (uint32_t)__cvta_generic_to_shared(&bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
  asm volatile("fence.mbarrier_init.release.cluster;");
  for (int i = 0; i < 5; ++i) {
    unsigned long long v = *(volatile unsigned long long*)&bar;
    printf("phase %d: word = %016llx\n", i, v);
    if (i & 1) asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], 0; }" ::"r"(a) : "memory");
    else asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(a) : "memory");
  }
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], 4096; }" ::"r"(a) : "memory");
  printf("pending tx 4096: word = %016llx\n", *(volatile unsigned long long*)&bar);
}
int main() { k<<<1, 1>>>(); cudaDeviceSynchronize(); return 0; }
