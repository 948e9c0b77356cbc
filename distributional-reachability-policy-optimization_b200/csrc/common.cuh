// Shared helpers of libdrpo_sm100.so: error plumbing, launch accounting, torch-compatible math, Philox noise.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/drpo_b200.h"

namespace drpo {

// ---- error plumbing: nothing throws across the C ABI ------------------------------------------------------
void set_error(const char* fmt, ...);
extern int64_t g_launch_count;

// In-kernel watchdog plumbing of the tcgen05 kernels (every wait inside them is bounded; a protocol bug is reported, never a hang).
// Two pinned, device-mapped host words (0 = rollout kernels, 1 = update-step kernels) are STICKY: the device writes a word only
// when a launch sequence flagged a time-out and nothing ever clears it, so a later successful launch cannot hide the failure.
// publish_status() enqueues that write for `err_flag` (device int, non-zero = failing wait's code) and, when `loss_flag` is given,
// stores 0/1 into the update step's losses[DRPO_LOSS_ERR_SLOT], which phase 2 and the gradient all-reduce consume.
int* status_words_host();
int publish_status(const int* err_flag, int which, float* loss_flag, void* stream);
int kernel_status_peek();          // non-blocking: the sticky words as the host sees them now
int kernel_status_sync();          // device-synchronising (drpo_kernel_status)

#define DRPO_CHECK_ARG(cond, ...)                                                  \
  do {                                                                             \
    if (!(cond)) { ::drpo::set_error(__VA_ARGS__); return DRPO_ERR_ARG; }          \
  } while (0)

#define DRPO_CUDA_OK(expr)                                                         \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      ::drpo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DRPO_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

// every kernel launch goes through this so drpo_launch_count() is an honest count
#define DRPO_LAUNCH(kernel, grid, block, smem, stream, ...)                        \
  do {                                                                             \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);      \
    ++::drpo::g_launch_count;                                                      \
    cudaError_t _e = cudaPeekAtLastError();                                        \
    if (_e != cudaSuccess) {                                                       \
      ::drpo::set_error("launch of %s failed: %s (%s:%d)", #kernel, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DRPO_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller-provided workspace
struct Arena {
  char* base; int64_t size; int64_t off;
  Arena(void* p, int64_t n) : base((char*)p), size(n), off(0) {}
  template <typename T> T* take(int64_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base + off);
    off += count * (int64_t)sizeof(T);
    return r;
  }
  bool ok() const { return off <= size; }
};

// ---- activations, exactly as torch evaluates them in fp32 ---------------------------------------------------
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2, ACT_TANH = 3 };

__device__ __forceinline__ float softplus_f(float x) {          // F.softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return x < 0.f ? 0.f : x;              // NaN-propagating, like torch.relu (fmaxf would turn NaN into 0)
    case ACT_SILU: return silu_f(x);
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}
// soft clamp used by the ensemble log-var head (src/dynamics.py:120-121) and Qc log-std head (src/ssac.py:75-76)
__device__ __forceinline__ float soft_clamp(float x, float lo, float hi) {
  x = hi - softplus_f(hi - x);
  return lo + softplus_f(x - lo);
}

// ---- Philox4x32-10, counter = (row, column group, stream tag, step), key = seed ------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0; k.y += W1;
  }
  return c;
}

// 4 standard normals for (row, column group).  Written with explicit-rounding intrinsics so that every kernel
// (and drpo_philox_normal, which exports the stream to the CPU oracle) produces bit-identical values.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t row, uint32_t colgrp, uint32_t tag, uint32_t step) {
  uint4 r = philox4x32_10(make_uint4(row, colgrp, tag, step), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float S = 2.3283064365386963e-10f;  // 2^-32
  float u0 = __fmaf_rn((float)r.x, S, 0.5f * S), u1 = __fmul_rn((float)r.y, S);
  float u2 = __fmaf_rn((float)r.z, S, 0.5f * S), u3 = __fmul_rn((float)r.w, S);
  u0 = fminf(u0, 0.99999994f); u2 = fminf(u2, 0.99999994f);
  float r0 = __fsqrt_rn(__fmul_rn(-2.f, __logf(u0))), r1 = __fsqrt_rn(__fmul_rn(-2.f, __logf(u2)));
  float s0, c0, s1, c1;
  __sincosf(__fmul_rn(6.283185307179586f, u1), &s0, &c0);
  __sincosf(__fmul_rn(6.283185307179586f, u3), &s1, &c1);
  return make_float4(__fmul_rn(r0, c0), __fmul_rn(r0, s0), __fmul_rn(r1, c1), __fmul_rn(r1, s1));
}
__device__ __forceinline__ float philox_normal1(uint64_t seed, uint32_t row, uint32_t col, uint32_t tag, uint32_t step) {
  float4 v = philox_normal4(seed, row, col >> 2, tag, step);
  switch (col & 3) { case 0: return v.x; case 1: return v.y; case 2: return v.z; default: return v.w; }
}

// stream tags (one per distinct draw of the hot path)
enum NoiseTag : uint32_t {
  TAG_ROLLOUT_POLICY = 1, TAG_ROLLOUT_MODEL = 2, TAG_CRITIC_ACTOR = 3, TAG_CRITIC_SAFE = 4, TAG_CRITIC_QC = 5,
  TAG_MULT_ACTOR = 6, TAG_ACTOR_ACTOR = 7, TAG_ACTOR_SAFE = 8, TAG_USER = 16
};

// device view of drpo_noise
struct NoiseView {
  const float* eps; int64_t row_stride; int64_t row_off; uint64_t seed; uint32_t tag, step;
  // injected draws are indexed by the row as given; Philox draws by the GLOBAL row id (row + row_off) so that a
  // sharded run reproduces the single-GPU stream
  __device__ __forceinline__ float get(int64_t row, int col) const {
    return eps ? eps[row * row_stride + col]
               : philox_normal1(seed, (uint32_t)(row + row_off), (uint32_t)col, tag, step);
  }
};
// four consecutive columns [4*colgrp, 4*colgrp+4) of one row (same values as get(); one Philox call instead of four)
__device__ __forceinline__ float4 noise_get4(const NoiseView& n, int64_t row, int colgrp, int ncols) {
  if (n.eps) {
    const float* p = n.eps + row * n.row_stride + 4 * colgrp;
    const int left = ncols - 4 * colgrp;
    return make_float4(left > 0 ? p[0] : 0.f, left > 1 ? p[1] : 0.f, left > 2 ? p[2] : 0.f, left > 3 ? p[3] : 0.f);
  }
  return philox_normal4(n.seed, (uint32_t)(row + n.row_off), (uint32_t)colgrp, n.tag, n.step);
}
static inline NoiseView make_noise(const float* eps, int64_t stride, uint64_t seed, uint32_t tag, uint32_t step,
                                   int64_t row_off = 0) {
  NoiseView n; n.eps = eps; n.row_stride = stride; n.row_off = row_off; n.seed = seed; n.tag = tag; n.step = step; return n;
}

// ---- warp / block reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace drpo
