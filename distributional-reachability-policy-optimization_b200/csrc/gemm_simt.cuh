// fp32 FMA GEMM with fused prologue/epilogue — the DRPO_PREC_FP32 (<=1e-5 parity) path and the workhorse of the
// hand-written backward pass.  C[M,N] = epilogue( sum_k A(m,k) * B(k,n) ), operands addressed through element
// strides so that  Y = X W^T (forward),  dX = dY W (backward-data)  and  dW = dY^T X (backward-weight, split-K
// over the batch) are the same kernel.  64x64x16 tiles, 256 threads, 4x4 outputs per thread.
#pragma once
#include <vector>

#include "common.cuh"

namespace drpo {

struct GemmArgs {
  const float* A; int64_t a_sm, a_sk;    // A(m,k) = A[m*a_sm + k*a_sk]
  const float* B; int64_t b_sk, b_sn;    // B(k,n) = B[k*b_sk + n*b_sn]
  float* C; int64_t ldc;                 // C[m*ldc + n]
  const float* bias;                     // [N], added before the activation (may be null)
  int M, N, K;
  const int* m_dev;                      // optional device-side row count (rollout: alive rows of this step)
  int act;                               // Act applied to (acc + bias)
  const float* mask; int64_t ldmask;     // backward-data: multiply by act'(saved post-activation output)
  int mask_mode;                         // 0 none, 1 relu' (mask>0), 2 tanh' (1-mask^2)
  float beta;                            // C = result + beta*C
  float* bias_out;                       // backward-weight: B gets a virtual all-ones last column; that column of the
                                         // result (= column sums of dY = db) is routed to bias_out[m]
  int splitk; float* partial;            // split-K: raw partial sums [splitk, M, N], reduced by splitk_reduce_kernel
  int batch;                             // > 1 (and splitk == 1): blockIdx.z walks independent problems (ensemble members) whose
  int64_t a_sb, b_sb, c_sb, bias_sb, mask_sb, bo_sb;   // operands are `*_sb` elements apart (0 = shared by all)
};

constexpr int GM = 64, GN = 64, GK = 16, GPAD = 4;

static __global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[GK][GM + GPAD];
  __shared__ __align__(16) float Bs[GK][GN + GPAD];
  int M = g.M;
  if (g.m_dev) M = min(M, *g.m_dev);
  if (g.batch > 1) {                      // per-problem operand bases
    const int64_t bz = blockIdx.z;
    g.A += bz * g.a_sb; g.B += bz * g.b_sb; g.C += bz * g.c_sb;
    if (g.bias) g.bias += bz * g.bias_sb;
    if (g.mask) g.mask += bz * g.mask_sb;
    if (g.bias_out) g.bias_out += bz * g.bo_sb;
  }
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  if (m0 >= M) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int N = g.N, K = g.K;
  const int n_real = g.bias_out ? N - 1 : N;     // columns of B that exist in memory
  int k_begin = 0, k_end = K;
  if (g.splitk > 1) {
    int chunk = ((K + g.splitk - 1) / g.splitk + GK - 1) / GK * GK;
    k_begin = blockIdx.z * chunk;
    k_end = min(K, k_begin + chunk);
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kcontig = (g.a_sk == 1), b_kcontig = (g.b_sk == 1);
  for (int k0 = k_begin; k0 < k_end; k0 += GK) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int m, k;
      if (a_kcontig) { k = tid & 15; m = (tid >> 4) + 16 * p; } else { m = tid & 63; k = (tid >> 6) + 4 * p; }
      float v = 0.f;
      if (m0 + m < M && k0 + k < k_end) v = g.A[(int64_t)(m0 + m) * g.a_sm + (int64_t)(k0 + k) * g.a_sk];
      As[k][m] = v;
      int n;
      if (b_kcontig) { k = tid & 15; n = (tid >> 4) + 16 * p; } else { n = tid & 63; k = (tid >> 6) + 4 * p; }
      v = 0.f;
      if (k0 + k < k_end) {
        if (n0 + n < n_real) v = g.B[(int64_t)(k0 + k) * g.b_sk + (int64_t)(n0 + n) * g.b_sn];
        else if (g.bias_out && n0 + n == N - 1) v = 1.f;
      }
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (g.splitk > 1) { g.partial[((int64_t)blockIdx.z * g.M + m) * N + n] = v; continue; }
      if (g.bias) v += g.bias[n];
      v = apply_act(v, g.act);
      if (g.mask_mode) {
        const float mk = g.mask[(int64_t)m * g.ldmask + n];
        v *= (g.mask_mode == 1) ? (mk > 0.f ? 1.f : 0.f) : (1.f - mk * mk);
      }
      if (g.bias_out && n == N - 1) { g.bias_out[m] = v + (g.beta != 0.f ? g.beta * g.bias_out[m] : 0.f); continue; }
      float* c = g.C + (int64_t)m * g.ldc + n;
      *c = v + (g.beta != 0.f ? g.beta * *c : 0.f);
    }
  }
}

// deterministic second stage of split-K (fixed summation order over the splits)
static __global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splitk, int M, int N, float* C, int64_t ldc,
                                     float* bias_out, float beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float s = 0.f;
  for (int z = 0; z < splitk; ++z) s += partial[((int64_t)z * M + m) * N + n];
  if (bias_out && n == N - 1) { bias_out[m] = s + (beta != 0.f ? beta * bias_out[m] : 0.f); return; }
  float* c = C + (int64_t)m * ldc + n;
  *c = s + (beta != 0.f ? beta * *c : 0.f);
}

static inline GemmArgs gemm_args() { GemmArgs g; memset(&g, 0, sizeof(g)); g.splitk = 1; return g; }

static inline int launch_gemm(const GemmArgs& g, void* stream) {
  if (g.M <= 0 || g.N <= 0) return DRPO_OK;
  if (g.batch > 1 && g.splitk > 1) { set_error("launch_gemm: split-K and batching are exclusive"); return DRPO_ERR_ARG; }
  dim3 grid((g.N + GN - 1) / GN, (g.M + GM - 1) / GM, g.splitk > 1 ? g.splitk : (g.batch > 1 ? g.batch : 1));
  DRPO_LAUNCH(gemm_f32_kernel, grid, 256, 0, stream, g);
  if (g.splitk > 1) {
    int64_t total = (int64_t)g.M * g.N;
    DRPO_LAUNCH(splitk_reduce_kernel, (unsigned)((total + 255) / 256), 256, 0, stream, g.partial, g.splitk, g.M, g.N,
                g.C, g.ldc, g.bias_out, g.beta);
  }
  return DRPO_OK;
}

// ---- tensor-core mode (DRPO_PREC_BF16 for the critic / multiplier steps) ---------------------------------------------
// The plain dense contractions of the critic step go to cuBLAS TF32 tensor-op GEMMs (a library GEMM: the baseline a
// hand-written tcgen05 critic kernel has to beat); bias + activation, activation-gradient masks and
// bias gradients (column sums) run as the fused elementwise kernels below.
}  // namespace drpo
#include <cublas_v2.h>
#include <cublasLt.h>
namespace drpo {

extern thread_local int g_gemm_mode;            // 0 = fp32 FMA kernels, 1 = cuBLAS TF32 (set per C-ABI call)
cublasHandle_t gemm_cublas_handle();            // lazily created, one per process (defined in drpo_api.cu)

#define DRPO_CUBLAS_OK(expr)                                                      \
  do {                                                                             \
    cublasStatus_t _s = (expr);                                                    \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                             \
      ::drpo::set_error("%s failed: cublas status %d (%s:%d)", #expr, (int)_s, __FILE__, __LINE__); \
      return DRPO_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

static __global__ void __launch_bounds__(256) bias_act_kernel(float* __restrict__ Y, int64_t ldy, const float* __restrict__ bias, int M, int N, int act) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N; const int n = (int)(i - m * N);
    float* y = Y + m * ldy + n;
    *y = apply_act(*y + (bias ? bias[n] : 0.f), act);
  }
}
static __global__ void __launch_bounds__(256) act_mask_kernel(float* __restrict__ dX, int64_t ldx, const float* __restrict__ saved, int64_t lds, int M, int N, int mode) {
  if ((N & 3) == 0 && (ldx & 3) == 0 && (lds & 3) == 0) {             // float4 path: rows are 16-byte aligned
    const int n4 = N >> 2;
    const int64_t total = (int64_t)M * n4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t m = i / n4; const int c = (int)(i - m * n4);
      float4* px = reinterpret_cast<float4*>(dX + m * ldx) + c;
      const float4 mk = *(reinterpret_cast<const float4*>(saved + m * lds) + c);
      float4 v = *px;
      if (mode == 1) { v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f; v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f; }
      else { v.x *= 1.f - mk.x * mk.x; v.y *= 1.f - mk.y * mk.y; v.z *= 1.f - mk.z * mk.z; v.w *= 1.f - mk.w * mk.w; }
      *px = v;
    }
    return;
  }
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N; const int n = (int)(i - m * N);
    const float mk = saved[m * lds + n];
    dX[m * ldx + n] *= (mode == 1) ? (mk > 0.f ? 1.f : 0.f) : (1.f - mk * mk);
  }
}
// db[n] = sum_m dY[m,n]: each block sums a slab of rows with float4 loads (a row segment of N floats is one coalesced run),
// fixed summation order, fp32 partials [slab][N]; one small reducer finishes
static __global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ dY, int64_t ldy, int M, int N, int rows_per_block, float* __restrict__ partial) {
  __shared__ float4 sm[256];
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  const bool vec = (N % 4 == 0) && (ldy % 4 == 0) && N <= 1024;
  if (vec) {
    const int n4 = N >> 2, lanes = 256 / n4;                    // row lanes per block (N = 256 -> 4)
    const int c = threadIdx.x % n4, r = threadIdx.x / n4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < lanes)
      for (int m = m0 + r; m < m1; m += lanes) {
        const float4 v = *reinterpret_cast<const float4*>(dY + (int64_t)m * ldy + 4 * c);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    sm[threadIdx.x] = s;
    __syncthreads();
    if (r == 0) {
      for (int k = 1; k < lanes; ++k) { const float4 v = sm[k * n4 + c]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
      *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.y * N + 4 * c) = s;
    }
  } else {
    for (int n = threadIdx.x; n < N; n += 256) {
      float s = 0.f;
      for (int m = m0; m < m1; ++m) s += dY[(int64_t)m * ldy + n];
      partial[(int64_t)blockIdx.y * N + n] = s;
    }
  }
}
// second stage: block = 32 columns x 8 partial-lanes, fixed summation order
static __global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int nblocks, int N, float* __restrict__ db) {
  __shared__ float sm[8][33];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  float s = 0.f;
  if (n < N) for (int b = w; b < nblocks; b += 8) s += partial[(int64_t)b * N + n];
  sm[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
    db[n] = t;
  }
}
static inline unsigned ew_grid(int64_t work) { int64_t g = (work + 255) / 256; if (g < 1) g = 1; if (g > 148 * 16) g = 148 * 16; return (unsigned)g; }

// row-major C[m,n] (+)= op over row-major operands, expressed in cuBLAS' column-major terms
static inline int tc_gemm(cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const float* A, int64_t lda, const float* B, int64_t ldb,
                          float* C, int64_t ldc, float beta, void* stream) {
  cublasHandle_t h = gemm_cublas_handle();
  if (!h) { set_error("cuBLAS handle could not be created"); return DRPO_ERR_CUDA; }
  DRPO_CUBLAS_OK(cublasSetStream(h, (cudaStream_t)stream));
  const float alpha = 1.f;
  DRPO_CUBLAS_OK(cublasGemmEx(h, ta, tb, m, n, k, &alpha, A, CUDA_R_32F, (int)lda, B, CUDA_R_32F, (int)ldb, &beta, C, CUDA_R_32F, (int)ldc,
                              CUBLAS_COMPUTE_32F_FAST_TF32, CUBLAS_GEMM_DEFAULT_TENSOR_OP));
  ++g_launch_count;
  return DRPO_OK;
}

// Forward layer with the bias (+ReLU) fused into the GEMM epilogue through cuBLASLt: Y^T[N,M] = W[K,N]^T X^T[K,M] (+ b, relu).
// One heuristic query per distinct (N, M, K, ld, epilogue) shape, cached for the life of the process.
struct LtPlan { int n, m, k; int64_t ldx, ldy; int epi; uint32_t al; cublasLtMatmulDesc_t desc; cublasLtMatrixLayout_t a, b, c; cublasLtMatmulAlgo_t algo; bool ok; };
extern thread_local void* g_lt_workspace; extern thread_local size_t g_lt_workspace_bytes;   // lent by the caller per C-ABI call
static inline int tc_linear_fwd_lt(const float* X, int64_t ldx, const drpo_linear& L, float* Y, int64_t ldy, int M, int act, void* stream, bool* handled) {
  static std::vector<LtPlan> plans;
  *handled = false;
  if (act != ACT_NONE && act != ACT_RELU) return DRPO_OK;
  const int epi = (act == ACT_RELU) ? (L.b ? 2 : 1) : (L.b ? 3 : 0);   // 0 none, 1 relu, 2 relu+bias, 3 bias
  if (epi == 0) return DRPO_OK;
  cublasLtHandle_t lt = (cublasLtHandle_t)gemm_cublas_handle();
  if (!lt) return DRPO_OK;
  // the parameters live at arbitrary 4-byte offsets of the flat arena: the kernel choice must know the real alignments
  auto align_of = [](const void* p) -> uint32_t { const uintptr_t v = (uintptr_t)p; uint32_t a = 256; while (a > 4 && (v & (a - 1))) a >>= 1; return a; };
  const uint32_t alW = align_of(L.w), alX = align_of(X), alY = align_of(Y), alB = L.b ? align_of(L.b) : 256;
  const uint32_t al = alW | (alX << 8) | (alY << 16) | ((alB >= 16 ? 16u : alB) << 24);
  LtPlan* P = nullptr;
  for (auto& q : plans) if (q.n == L.out_dim && q.m == M && q.k == L.in_dim && q.ldx == ldx && q.ldy == ldy && q.epi == epi && q.al == al) { P = &q; break; }
  if (!P) {
    LtPlan q{}; q.n = L.out_dim; q.m = M; q.k = L.in_dim; q.ldx = ldx; q.ldy = ldy; q.epi = epi; q.al = al; q.ok = false;
    cublasOperation_t ta = CUBLAS_OP_T, tb = CUBLAS_OP_N;
    const cublasLtEpilogue_t e = epi == 1 ? CUBLASLT_EPILOGUE_RELU : (epi == 2 ? CUBLASLT_EPILOGUE_RELU_BIAS : CUBLASLT_EPILOGUE_BIAS);
    bool good = cublasLtMatmulDescCreate(&q.desc, CUBLAS_COMPUTE_32F_FAST_TF32, CUDA_R_32F) == CUBLAS_STATUS_SUCCESS;
    good = good && cublasLtMatmulDescSetAttribute(q.desc, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)) == CUBLAS_STATUS_SUCCESS;
    good = good && cublasLtMatmulDescSetAttribute(q.desc, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)) == CUBLAS_STATUS_SUCCESS;
    good = good && cublasLtMatmulDescSetAttribute(q.desc, CUBLASLT_MATMUL_DESC_EPILOGUE, &e, sizeof(e)) == CUBLAS_STATUS_SUCCESS;
    // A = W stored [K,N] column-major (ld K), B = X^T stored [K,M] (ld ldx), C = Y^T stored [N,M] (ld ldy)
    good = good && cublasLtMatrixLayoutCreate(&q.a, CUDA_R_32F, L.in_dim, L.out_dim, L.in_dim) == CUBLAS_STATUS_SUCCESS;
    good = good && cublasLtMatrixLayoutCreate(&q.b, CUDA_R_32F, L.in_dim, M, ldx) == CUBLAS_STATUS_SUCCESS;
    good = good && cublasLtMatrixLayoutCreate(&q.c, CUDA_R_32F, L.out_dim, M, ldy) == CUBLAS_STATUS_SUCCESS;
    if (good) {
      cublasLtMatmulPreference_t pref; cublasLtMatmulHeuristicResult_t res{}; int found = 0;
      size_t wsb = g_lt_workspace_bytes;
      if (cublasLtMatmulPreferenceCreate(&pref) == CUBLAS_STATUS_SUCCESS) {
        cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &wsb, sizeof(wsb));
        cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_A_BYTES, &alW, sizeof(alW));
        cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_B_BYTES, &alX, sizeof(alX));
        cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_C_BYTES, &alY, sizeof(alY));
        cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_D_BYTES, &alY, sizeof(alY));
        // the bias pointer's alignment takes part in the heuristic: set a representative one
        const void* bias = L.b;
        if (epi >= 2) cublasLtMatmulDescSetAttribute(q.desc, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias));
        if (cublasLtMatmulAlgoGetHeuristic(lt, q.desc, q.a, q.b, q.c, q.c, pref, 1, &res, &found) == CUBLAS_STATUS_SUCCESS && found > 0) { q.algo = res.algo; q.ok = true; }
        cublasLtMatmulPreferenceDestroy(pref);
      }
    }
    plans.push_back(q);
    P = &plans.back();
  }
  if (!P->ok) return DRPO_OK;                              // no fused kernel for this shape: the caller falls back to GEMM + bias_act
  const void* bias = L.b;
  if (epi >= 2) DRPO_CUBLAS_OK(cublasLtMatmulDescSetAttribute(P->desc, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)));
  const float alpha = 1.f, beta = 0.f;
  DRPO_CUBLAS_OK(cublasLtMatmul(lt, P->desc, &alpha, L.w, P->a, X, P->b, &beta, Y, P->c, Y, P->c, &P->algo, g_lt_workspace, g_lt_workspace_bytes, (cudaStream_t)stream));
  ++g_launch_count;
  *handled = true;
  return DRPO_OK;
}

// ---- convenience wrappers -------------------------------------------------------------------------------------
// Y[M,N] = act(X[M,K] W[N,K]^T + b)
static inline int linear_fwd(const float* X, int64_t ldx, const drpo_linear& L, float* Y, int64_t ldy, int M, int act,
                             const int* m_dev, void* stream) {
  if (g_gemm_mode == 1 && !m_dev && M > 0) {
    bool handled = false;
    int rc0 = tc_linear_fwd_lt(X, ldx, L, Y, ldy, M, act, stream, &handled);
    if (rc0) return rc0;
    if (handled) return DRPO_OK;
    // Y^T[N,M] = W[K,N]^T X^T[K,M] in column-major terms
    int rc = tc_gemm(CUBLAS_OP_T, CUBLAS_OP_N, L.out_dim, M, L.in_dim, L.w, L.in_dim, X, ldx, Y, ldy, 0.f, stream);
    if (rc) return rc;
    if (L.b || act != ACT_NONE) DRPO_LAUNCH(bias_act_kernel, ew_grid((int64_t)M * L.out_dim), 256, 0, stream, Y, ldy, L.b, M, L.out_dim, act);
    return DRPO_OK;
  }
  GemmArgs g = gemm_args();
  g.A = X; g.a_sm = ldx; g.a_sk = 1;
  g.B = L.w; g.b_sk = 1; g.b_sn = L.in_dim;
  g.C = Y; g.ldc = ldy; g.bias = L.b; g.M = M; g.N = L.out_dim; g.K = L.in_dim; g.act = act; g.m_dev = m_dev;
  return launch_gemm(g, stream);
}
// ---- the same three contractions for E independent layers in one launch (ensemble members; fp32 FFMA kernel only) --------
// X[e] = X + e*x_sb (x_sb = 0: every member reads the same rows), W[e] = W + e*N*K, b[e] = b + e*N, Y[e] = Y + e*y_sb
static inline int linear_fwd_batched(int E, const float* X, int64_t ldx, int64_t x_sb, const drpo_linear& L0, float* Y, int64_t ldy, int64_t y_sb,
                                     int M, int act, void* stream) {
  GemmArgs g = gemm_args();
  g.A = X; g.a_sm = ldx; g.a_sk = 1; g.a_sb = x_sb;
  g.B = L0.w; g.b_sk = 1; g.b_sn = L0.in_dim; g.b_sb = (int64_t)L0.out_dim * L0.in_dim;
  g.C = Y; g.ldc = ldy; g.c_sb = y_sb; g.bias = L0.b; g.bias_sb = L0.out_dim;
  g.M = M; g.N = L0.out_dim; g.K = L0.in_dim; g.act = act; g.batch = E;
  return launch_gemm(g, stream);
}
static inline int linear_bwd_data_batched(int E, const float* dY, int64_t ldy, int64_t dy_sb, const drpo_linear& L0, float* dX, int64_t lddx,
                                          int64_t dx_sb, int M, float beta, void* stream) {
  GemmArgs g = gemm_args();
  g.A = dY; g.a_sm = ldy; g.a_sk = 1; g.a_sb = dy_sb;
  g.B = L0.w; g.b_sk = L0.in_dim; g.b_sn = 1; g.b_sb = (int64_t)L0.out_dim * L0.in_dim;
  g.C = dX; g.ldc = lddx; g.c_sb = dx_sb; g.M = M; g.N = L0.in_dim; g.K = L0.out_dim; g.beta = beta; g.batch = E;
  return launch_gemm(g, stream);
}
// dW[e] = dY[e]^T X[e], db[e] = column sums of dY[e]; no split-K: meant for a few hundred rows per member
static inline int linear_bwd_weight_batched(int E, const float* dY, int64_t ldy, int64_t dy_sb, const float* X, int64_t ldx, int64_t x_sb, int M,
                                            int n_out, int k_in, float* dW, float* db, void* stream) {
  GemmArgs g = gemm_args();
  g.A = dY; g.a_sm = 1; g.a_sk = ldy; g.a_sb = dy_sb;
  g.B = X; g.b_sk = ldx; g.b_sn = 1; g.b_sb = x_sb;
  g.C = dW; g.ldc = k_in; g.c_sb = (int64_t)n_out * k_in; g.M = n_out; g.N = k_in + 1; g.K = M; g.bias_out = db; g.bo_sb = n_out; g.batch = E;
  return launch_gemm(g, stream);
}
// dX[M,K] = (dY[M,N] W[N,K]) * act'(saved)   (+ beta*dX)
static inline int linear_bwd_data(const float* dY, int64_t ldy, const drpo_linear& L, float* dX, int64_t lddx, int M,
                                  const float* saved, int64_t ldsaved, int mask_mode, float beta, void* stream) {
  if (g_gemm_mode == 1 && M > 0 && (beta == 0.f || mask_mode <= 1)) {
    // dX^T[K,M] = W[K,N] dY^T[N,M]; with beta = 1 the accumulated term already carries the same 0/1 relu mask (idempotent)
    int rc = tc_gemm(CUBLAS_OP_N, CUBLAS_OP_N, L.in_dim, M, L.out_dim, L.w, L.in_dim, dY, ldy, dX, lddx, beta, stream);
    if (rc) return rc;
    if (mask_mode) DRPO_LAUNCH(act_mask_kernel, ew_grid((int64_t)M * L.in_dim), 256, 0, stream, dX, lddx, saved, ldsaved, M, L.in_dim, mask_mode);
    return DRPO_OK;
  }
  GemmArgs g = gemm_args();
  g.A = dY; g.a_sm = ldy; g.a_sk = 1;
  g.B = L.w; g.b_sk = L.in_dim; g.b_sn = 1;
  g.C = dX; g.ldc = lddx; g.M = M; g.N = L.in_dim; g.K = L.out_dim;
  g.mask = saved; g.ldmask = ldsaved; g.mask_mode = mask_mode; g.beta = beta;
  return launch_gemm(g, stream);
}
// dW[N,K] = dY[M,N]^T X[M,K],  db[N] = column sums of dY ; split-K over the batch with a deterministic reduction
static inline int linear_bwd_weight(const float* dY, int64_t ldy, const float* X, int64_t ldx, int M, int n_out, int k_in,
                                    float* dW, float* db, float* partial, int64_t partial_floats, void* stream) {
  if (g_gemm_mode == 1 && M > 0) {
    // dW^T[K,N] = X^T[K,M] dY[M,N]  (contraction over the batch), db = column sums of dY
    int rc = tc_gemm(CUBLAS_OP_N, CUBLAS_OP_T, k_in, n_out, M, X, ldx, dY, ldy, dW, k_in, 0.f, stream);
    if (rc) return rc;
    if (db) {
      int nb = (M + 127) / 128;                                         // ~128 rows per block: 512 blocks at B = 64k
      while ((int64_t)nb * n_out > partial_floats && nb > 1) nb = (nb + 1) / 2;
      const int rows = (M + nb - 1) / nb;
      DRPO_LAUNCH(colsum_partial_kernel, dim3(1, nb), 256, 0, stream, dY, ldy, M, n_out, rows, partial);
      DRPO_LAUNCH(colsum_final_kernel, (n_out + 31) / 32, 256, 0, stream, partial, nb, n_out, db);
    }
    return DRPO_OK;
  }
  GemmArgs g = gemm_args();
  g.A = dY; g.a_sm = 1; g.a_sk = ldy;
  g.B = X; g.b_sk = ldx; g.b_sn = 1;
  g.C = dW; g.ldc = k_in; g.M = n_out; g.N = k_in + 1; g.K = M; g.bias_out = db;
  int tiles = ((g.N + GN - 1) / GN) * ((g.M + GM - 1) / GM);
  int want = (M + 511) / 512;                         // >= 512 batch rows per split
  int cap = (296 + tiles - 1) / tiles;                // about two waves of 148 SMs
  int sk = want < cap ? want : cap;
  while (sk > 1 && (int64_t)sk * g.M * g.N > partial_floats) --sk;
  if (sk > 1) { g.splitk = sk; g.partial = partial; }
  return launch_gemm(g, stream);
}

}  // namespace drpo
