// DRPO_PREC_BF16 multiplier step (SSAC.update_multiplier, src/ssac.py:529-578) and actor / temperature / safe-actor step
// (SSAC.update_actor_and_alpha, src/ssac.py:458-527) on the fused tcgen05 op tables of umma_ops.cuh: per 128-row tile every
// dense layer of the update is one table-driven "op" (tcgen05.mma bf16 x bf16 -> fp32 into four 64-column TMEM accumulators,
// activations kept in TMEM as the next op's A operand), the narrow heads run on the CUDA cores of the epilogue groups, and the
// per-row loss / output-gradient algebra sits in the ops' post steps.  What leaves the SM is what the split-K dW kernel needs
// (bf16 activations / activation gradients of the TRAINED nets in the octet layout), the ReLU masks of the frozen critics the
// action gradient flows through, per-CTA column sums and loss partials.
//
//   multiplier step (15 ops / tile):  actor -> a ; Qc(obs, a) -> penalty ; safe actor (eval) -> a_s ; Qc(obs, a_s) -> safe_Qc ;
//       lambda net (tanh layers) on [obs, safe_Qc] -> loss, d raw -> head backward -> dh1 through W1^T.
//   actor step (27 ops / tile):  safe actor -> a_eval ; Qc(obs, a_eval) -> safe_Qc ; lambda net -> lambda (no gradient) ;
//       actor -> a, log_prob ; Q_k(obs, a) forward + dX down to the action columns ; Qc(obs, a) forward + dX (both heads, trunk) ;
//       squashed-Gaussian rsample backward -> actor head backward -> dh1 ; a_safe' = rsample of the safe actor ;
//       Qc(obs, a_safe') forward + dX ; safe-actor head backward -> dh1.
//   The last layer of the dX chain through a frozen critic (dh1 . W0, of which only the A action columns are wanted) is a CUDA-core
//   dot product inside the backward epilogue, not an N = 256 MMA.
#include "critic_umma_api.h"
#include "umma_ops.cuh"
#include "actor.cuh"

namespace drpo {
namespace cu {

constexpr int S_MAX_OPS = 28;
enum SPost {
  SP_NONE = 0, SP_KEEP,
  // multiplier step
  SP_POLICY, SP_QCUB_PEN, SP_POLICY_EVAL, SP_QCUB_SAFE, SP_MULT,
  // actor step
  SP_SAFE_FWD, SP_LAM, SP_QK, SP_DA_Q, SP_QCUB_GRAD1, SP_DA_ACTOR, SP_PATCH_SAFE, SP_QCUB_GRAD2, SP_DA_SAFE
};
enum SMode { MODE_MULT = 0, MODE_ACTOR = 1 };

struct SolverParams {
  FOp op[S_MAX_OPS];
  EOp eop[S_MAX_OPS];
  int n_ops;
  JobSched sch;
  const uint8_t* wimg;
  const float* ctab; int ctab_floats;
  int hw_lam2;                             // multiplier: head weights of the lambda net (backward)
  int hw_q2, hw_cm, hw_cl, hw_actor2, hw_safe2, zero_off;
  const float* obs;
  NoiseView n_actor, n_safe;
  const float* log_alpha;
  float ratio, thr, pen_lb, pen_ub, ub, lam_eps, inv_bg, target_entropy;
  int64_t B, Bpad; int S, A, C, D, Kx, stages, n_tiles;
  __nv_bfloat16* x_obs;                    // octets of [obs, 0]          (actor step: B operand of the first-layer dW jobs)
  __nv_bfloat16* x_aug;                    // octets of [obs, safe_Qc]    (multiplier step)
  __nv_bfloat16* sv[16];                   // [0] unused
  float* gacc_out; int nv;
  double* loss_part;                       // [grid][4]
  int* err_flag;
  float* dbg;                              // optional [B,16] per-row intermediates (tests)
};

// saved arrays
enum MSave { MS_H1 = 1, MS_DH2, MS_DH1, MS_COUNT = 3 };
enum ASave { AS_SH1 = 1, AS_SH2, AS_PH1, AS_PH2, AS_QH1, AS_CT1, AS_CT2, AS_PDH2, AS_PDH1, AS_SDH2, AS_SDH1, AS_COUNT = 11 };
// column-sum slots.  multiplier: W2, b1, b0, scalars.  actor: per trained net n (0 actor, 1 safe) 2A head rows, b1, b0; then scalars
constexpr int MSLOT_W2 = 0, MSLOT_B1 = 1, MSLOT_B0 = 2, MSLOT_SCAL = 3, M_NV = 4;
__host__ __device__ inline int aslot_w2(int n, int A) { return n * (2 * A + 2); }
__host__ __device__ inline int aslot_b1(int n, int A) { return n * (2 * A + 2) + 2 * A; }
__host__ __device__ inline int aslot_b0(int n, int A) { return n * (2 * A + 2) + 2 * A + 1; }
__host__ __device__ inline int aslot_scal(int A) { return 2 * (2 * A + 2); }
__host__ __device__ inline int a_nv(int A) { return 2 * (2 * A + 2) + 1; }

static float* g_solver_dbg = nullptr;
void solver_set_debug_rows(float* p) { g_solver_dbg = p; }

template <int A>
__device__ __forceinline__ void patch_action(uint8_t* xs0, const Epi& e, int S, const float (&an)[A]) {
  if (e.g == 0) {
#pragma unroll
    for (int j = 0; j < A; ++j) {
      const int k = S + j;
      *reinterpret_cast<__nv_bfloat16*>(xs0 + (k >> 3) * 2048 + e.row * 16 + (k & 7) * 2) = __float2bfloat16_rn(e.valid ? an[j] : 0.f);
    }
    fence_proxy_async();
  }
}

// squashed-Gaussian rsample + log-prob of one row (src/policy.py:89-97, src/squashed_gaussian.py:7-16); keeps what the backward needs
template <int A>
__device__ __forceinline__ float rsample_row(const float (&out)[MAXO], const NoiseView& nv, int64_t gr, bool valid, float (&t)[A], float (&sd)[A],
                                             float (&eps)[A], float (&sg)[A]) {
  float lp = 0.f;
#pragma unroll
  for (int j = 0; j < A; ++j) {
    const float mu = out[j], raw = out[A + j];
    sg[j] = sigmoid_f(raw);
    sd[j] = expf(-6.f + 10.f * sg[j]);
    eps[j] = valid ? nv.get(gr, j) : 0.f;
    const float x = fmaf(eps[j], sd[j], mu);
    t[j] = tanhf(x);
    const float ladj = 2.f * (0.69314718055994531f - x - softplus_f(-2.f * x));
    const float dd = x - mu;
    lp += (0.f - ladj) + (-(dd * dd) / (2.f * (sd[j] * sd[j])) - logf(sd[j]) - 0.91893853320467267f);
  }
  return lp;
}

template <int A, int C, int MODE>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(F_THREADS, 1) solver_fused_kernel(const __grid_constant__ SolverParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stages = p.stages;
  uint8_t* ring = smem;
  uint8_t* xs0 = ring + (size_t)stages * CHUNK_BYTES;        // [obs, action]  (the action columns are patched by the policy post steps)
  uint8_t* xs1 = xs0 + TILE * p.Kx * 2;                      // [obs, safe_Qc] (the lambda net's input)
  uint8_t* ones = xs1 + TILE * p.Kx * 2;
  float* ctab = reinterpret_cast<float*>(ones + TILE * KBIAS * 2);
  float* gacc = ctab + ((p.ctab_floats + 3) & ~3);
  float4* hp = reinterpret_cast<float4*>(gacc + p.nv * HID);
  FusedSmem* sm = reinterpret_cast<FusedSmem*>(hp + NGROUPS * TILE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int PRODUCER = EPI_THREADS / 32, ISSUER = PRODUCER + 1;
  int* err = p.err_flag;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], CLUSTER); }
    for (int g = 0; g < NGROUPS; ++g) { mbar_init(&sm->acc_full[g], 1); mbar_init(&sm->acc_free[g], 4); }
    mbar_init(&sm->act_ready[0], EPI_THREADS / 32); mbar_init(&sm->act_ready[1], EPI_THREADS / 32);
    fence_barrier_init();
  }
  if (warp == ISSUER) tmem_alloc(&sm->tmem_base, 512);
  for (int i = threadIdx.x; i < p.ctab_floats; i += F_THREADS) ctab[i] = p.ctab[i];
  for (int i = threadIdx.x; i < p.nv * HID; i += F_THREADS) gacc[i] = 0.f;
  for (int i = threadIdx.x; i < TILE * KBIAS; i += F_THREADS) {
    const int k = (i >> 10) * 8 + (i & 7);
    reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16_rn(k < 2 ? 1.f : 0.f);
  }
  for (int i = threadIdx.x; i < TILE * p.Kx / 2; i += F_THREADS) reinterpret_cast<uint32_t*>(xs1)[i] = 0u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  const uint32_t crank = cluster_ctarank();
  const int n_clusters = (int)gridDim.x / CLUSTER, cid = (int)blockIdx.x / CLUSTER;
  const int my_jobs = sched_my_jobs(p.sch, p.n_tiles, cid, n_clusters);

  if (warp == PRODUCER) {
    fused_producer(p, sm, ring, my_jobs, cid, n_clusters, crank, err);
  } else if (warp == ISSUER) {
    fused_issuer(p, sm, ring, xs0, xs1, ones, tmem, my_jobs, cid, n_clusters, lane, err, (long long*)nullptr);
  } else {
    Epi e;
    e.sm = sm; e.ctab = ctab; e.gacc = gacc; e.hp = hp; e.g = warp >> 2; e.lane = lane; e.row = (warp & 3) * 32 + lane;
    e.tm = tmem + ((uint32_t)((warp & 3) * 32) << 16); e.it = 0; e.arr = 0; e.err = err; e.Bpad = p.Bpad; e.prof = nullptr;
    const int S = p.S;
    const float alpha = MODE == MODE_ACTOR ? expf(*p.log_alpha) : 0.f;
    // Loss sums and last-layer bias gradients go straight to shared memory (warp sum + one atomic per warp at the few sites that
    // produce them) instead of living in registers across the op loop: at 96 registers per thread every loop-carried value is a
    // spill to L2 around each epilogue.  lacc = three doubles in the unused tail of the scalar slot of the column-sum accumulators.
    const int scal = MODE == MODE_MULT ? MSLOT_SCAL : aslot_scal(A);
    float* scacc = gacc + scal * HID;                 // [0..2A): first trained net's head bias gradient, [4..4+2A): the safe actor's
    double* lacc = reinterpret_cast<double*>(scacc + 16);
    auto add_loss = [&](int k, double v) { v = warp_sum_d(v); if (lane == 0) atomicAdd(&lacc[k], v); };
    auto add_scal = [&](int k, float v) { v = warp_sum(v); if (lane == 0) atomicAdd(&scacc[k], v); };

    for (int t = 0; t < my_jobs; ++t) {
      const int tile = CLUSTER * (cid + t * n_clusters) + (int)crank;      // (no independent chains in these updates: sch.split = 0)
      e.grow = (int64_t)tile * TILE + e.row;
      e.valid = e.grow < p.B;
      const int64_t gr = e.valid ? e.grow : 0;
      // ---- stage xs0 = [obs, 0] (bf16, K-major); the actor step also keeps it as octets for the first-layer dW jobs --------------
      for (int j = e.g; j < (p.Kx >> 3); j += NGROUPS) {
        uint32_t w0[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = j * 8 + q * 2;
          const float v0 = (e.valid && k < S) ? p.obs[gr * S + k] : 0.f;
          const float v1 = (e.valid && k + 1 < S) ? p.obs[gr * S + k + 1] : 0.f;
          w0[q] = pack_bf16(v0, v1);
        }
        const uint4 v = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        *reinterpret_cast<uint4*>(xs0 + j * 2048 + e.row * 16) = v;
        if (MODE == MODE_ACTOR) *reinterpret_cast<uint4*>(p.x_obs + oct_index(e.grow, j, p.Kx >> 3)) = v;
      }
      fence_proxy_async();
      epi_op_done(e);

      float hpart[MAXO] = {0.f, 0.f, 0.f, 0.f}, hkeep[MAXO] = {0.f, 0.f, 0.f, 0.f};
      int hb_keep = 0;
      float pt[A], psd[A], peps[A], psg[A], logp = 0.f;        // rsample state of the policy being differentiated
      float mu_s[A], raw_s[A], da[A];
      float penalty = 0.f, safe_qc = 0.f, lam = 0.f;
#pragma unroll
      for (int j = 0; j < A; ++j) { pt[j] = psd[j] = peps[j] = psg[j] = mu_s[j] = raw_s[j] = da[j] = 0.f; }
#pragma unroll 1
      for (int o = 0; o < p.n_ops; ++o) {
        const EOp d = p.eop[o];
        const uint32_t region = d.out_region == 0 ? 0u : (d.out_region == 1 ? TM_R0 : TM_R1);
        if (!d.backward) epi_forward_halfwise(e, region, p.sv[d.save], d.hw_off, d.no, hpart, d.wait_all != 0, d.act);
        else epi_backward(e, p.sv[d.hsave], d.bias_slot, p.sv[d.save], region, d.wait_all != 0, d.act, d.hw_off, d.no, hpart);
        float* dbg = (p.dbg && e.g == 0 && e.valid) ? p.dbg + gr * 16 : nullptr;

        if (d.post == SP_KEEP) {
#pragma unroll
          for (int j = 0; j < MAXO; ++j) { hkeep[j] = hpart[j]; hpart[j] = 0.f; }
          hb_keep = d.hb_off;
        } else if (d.post == SP_POLICY || d.post == SP_POLICY_EVAL || d.post == SP_SAFE_FWD) {
          float out[MAXO];
          head_combine(e, hpart, d.hb_off, out);
          float an[A];
          if (d.post == SP_POLICY) {
            // action = actor.distr(obs).rsample() (+ log_prob)                               src/ssac.py:459-461, 530-531
            logp = rsample_row<A>(out, p.n_actor, gr, e.valid, pt, psd, peps, psg);
#pragma unroll
            for (int j = 0; j < A; ++j) an[j] = pt[j];
            if (dbg) { if (MODE == MODE_MULT) { dbg[0] = an[0]; dbg[1] = an[A - 1]; } else { dbg[3] = an[0]; dbg[4] = logp; } }
          } else {
            // action_safe = actor_safe.act(obs, eval=True) = tanh(mu)                         src/ssac.py:473, 546
#pragma unroll
            for (int j = 0; j < A; ++j) { an[j] = tanhf(out[j]); mu_s[j] = out[j]; raw_s[j] = out[A + j]; }
            if (dbg) dbg[MODE == MODE_MULT ? 4 : 0] = an[0];
          }
          patch_action<A>(xs0, e, S, an);
#pragma unroll
          for (int j = 0; j < MAXO; ++j) hpart[j] = 0.f;
        } else if (d.post == SP_QCUB_PEN || d.post == SP_QCUB_SAFE || d.post == SP_QCUB_GRAD1 || d.post == SP_QCUB_GRAD2) {
          // Qc_ub = max_c(mean + std_ratio * std) of constraint_critic(obs, a, uncertainty=True)      src/ssac.py:85, 588-600
          const bool grad = d.post == SP_QCUB_GRAD1 || d.post == SP_QCUB_GRAD2;
          if (grad) tmem_st_wait();
          float om[MAXO], ol[MAXO];
          head_combine(e, hkeep, hb_keep, om);
          named_bar_sync(1, EPI_THREADS);
          head_combine(e, hpart, d.hb_off, ol);
          float dmean[MAXO] = {0.f, 0.f, 0.f, 0.f}, dls[MAXO] = {0.f, 0.f, 0.f, 0.f};
          const float w = !e.valid ? 0.f : (d.post == SP_QCUB_GRAD1 ? lam * p.inv_bg : p.inv_bg);
          const float qc = qc_ub_max_grad(om, ol, C, p.ratio, w, dmean, dls);
          if (d.post == SP_QCUB_PEN) {
            // penalty = clamp(actor_Qc - threshold, lb, ub)                                   src/ssac.py:540-543
            penalty = fminf(fmaxf(qc - p.thr, p.pen_lb), p.pen_ub);
            if (dbg) { dbg[2] = qc; dbg[3] = penalty; }
          } else if (d.post == SP_QCUB_SAFE) {
            // safe_Qc -> the lambda net's input [obs, safe_Qc]                                src/ssac.py:107-108, 475-478, 548-549
            safe_qc = qc;
            for (int j = e.g; j < (p.Kx >> 3); j += NGROUPS) {
              uint32_t w0[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float v[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int k = j * 8 + q * 2 + u;
                  v[u] = !e.valid ? 0.f : (k < S ? p.obs[gr * S + k] : (k == S ? qc : 0.f));
                }
                w0[q] = pack_bf16(v[0], v[1]);
              }
              const uint4 vv = make_uint4(w0[0], w0[1], w0[2], w0[3]);
              *reinterpret_cast<uint4*>(xs1 + j * 2048 + e.row * 16) = vv;
              if (MODE == MODE_MULT) *reinterpret_cast<uint4*>(p.x_aug + oct_index(e.grow, j, p.Kx >> 3)) = vv;
            }
            fence_proxy_async();
            if (dbg) dbg[MODE == MODE_MULT ? 5 : 1] = qc;
          } else {
            // backward through both heads of the frozen constraint critic (m1 stashed in R0, l1 in R1)
            if (e.g == 0) add_loss(d.post == SP_QCUB_GRAD1 ? 0 : 1, !e.valid ? 0.0 : (d.post == SP_QCUB_GRAD1 ? (double)(lam * qc) : (double)qc));
            if (dbg) dbg[d.post == SP_QCUB_GRAD1 ? 6 : 12] = qc;
#pragma unroll 1
            for (int hd = 0; hd < 2; ++hd) {
              float dsel[MAXO];
#pragma unroll
              for (int j = 0; j < MAXO; ++j) dsel[j] = hd == 0 ? dmean[j] : dls[j];
              epi_head_backward(e, hd == 0 ? TM_R0 : TM_R1, dsel, C, hd == 0 ? p.hw_cm : p.hw_cl, -1, NO_SLOT, nullptr);
            }
          }
#pragma unroll
          for (int j = 0; j < MAXO; ++j) hpart[j] = 0.f;
        } else if (d.post == SP_MULT) {
          // lams = multiplier(obs, safe_Qc); lam_loss; backward through the head                src/ssac.py:107-111, 549-565
          tmem_st_wait();
          float out[MAXO];
          head_combine(e, hpart, d.hb_off, out);
          const float th = tanhf(out[0] / p.ub * 2.f);
          const float lm = p.ub / 2.f * (1.f + th);
          const bool unsafe = safe_qc > 0.f;
          const float ls = unsafe ? 0.f : lm, lu = unsafe ? lm : 0.f;
          const float tgt = unsafe ? (p.ub - p.lam_eps) : 0.f;
          const float dlam = unsafe ? 2.f * (lu - tgt) * p.inv_bg : -0.5f * penalty * p.inv_bg;
          float dr[MAXO] = {e.valid ? dlam * (1.f - th * th) : 0.f, 0.f, 0.f, 0.f};
          if (e.g == 0) {
            add_loss(0, e.valid ? (double)(ls * penalty) : 0.0); add_loss(1, e.valid ? (double)((lu - tgt) * (lu - tgt)) : 0.0);
            add_scal(0, dr[0]);
            if (dbg) { dbg[6] = out[0]; dbg[7] = lm; dbg[8] = dr[0]; }
          }
          epi_head_backward(e, TM_R1, dr, 1, p.hw_lam2, MSLOT_W2, MSLOT_B1, p.sv[MS_DH2], 1);
          hpart[0] = 0.f;
        } else if (d.post == SP_LAM) {
          float out[MAXO];
          head_combine(e, hpart, d.hb_off, out);
          lam = p.ub / 2.f * (1.f + tanhf(out[0] / p.ub * 2.f));                  // src/ssac.py:109-110 (detached)
          if (dbg) dbg[2] = lam;
          hpart[0] = 0.f;
        } else if (d.post == SP_QK) {
          // actor_Q = critic.random_choice(obs, action); d loss / d Q = -1/B; backward through the frozen head       src/ssac.py:462-464
          tmem_st_wait();
          float out[MAXO];
          head_combine(e, hpart, d.hb_off, out);
          if (e.g == 0) add_loss(0, e.valid ? (double)(alpha * logp - out[0]) : 0.0);
          if (dbg) dbg[5] = out[0];
          float dq[MAXO] = {e.valid ? -p.inv_bg : 0.f, 0.f, 0.f, 0.f};
          epi_head_backward(e, TM_R1, dq, 1, p.hw_q2, -1, NO_SLOT, nullptr);
          hpart[0] = 0.f;
        } else if (d.post == SP_DA_Q) {
          float out[MAXO];
          head_combine(e, hpart, p.zero_off, out);
#pragma unroll
          for (int j = 0; j < A; ++j) da[j] = out[j];
          if (dbg) dbg[7] = da[0];
#pragma unroll
          for (int j = 0; j < MAXO; ++j) hpart[j] = 0.f;
        } else if (d.post == SP_DA_ACTOR || d.post == SP_DA_SAFE) {
          // d loss / d action is complete: rsample / tanh / log-prob backward (closed form), then the policy head backward
          const bool perf = d.post == SP_DA_ACTOR;
          float out[MAXO];
          head_combine(e, hpart, p.zero_off, out);
          const float wl = perf ? alpha * p.inv_bg : 0.f;
          float dpo[MAXO] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < A; ++j) {
            const float dat = perf ? da[j] + out[j] : out[j];
            const float dx = dat * (1.f - pt[j] * pt[j]) + wl * 2.f * pt[j];
            dpo[j] = e.valid ? dx : 0.f;
            dpo[A + j] = e.valid ? (dx * psd[j] * peps[j] - wl) * 10.f * psg[j] * (1.f - psg[j]) : 0.f;
            if (dbg && j == 0) { dbg[perf ? 8 : 13] = dat; dbg[perf ? 9 : 14] = dpo[0]; dbg[perf ? 10 : 15] = dpo[A]; }
          }
          if (e.g == 0) {
            if (perf) add_loss(2, e.valid ? (double)(logp + p.target_entropy) : 0.0);
#pragma unroll
            for (int j = 0; j < 2 * A; ++j) add_scal((perf ? 0 : 4) + j, dpo[j]);
          }
          const int n = perf ? 0 : 1;
          epi_head_backward(e, TM_R0, dpo, 2 * A, perf ? p.hw_actor2 : p.hw_safe2, aslot_w2(n, A), aslot_b1(n, A),
                            p.sv[perf ? AS_PDH2 : AS_SDH2], 0, p.sv[perf ? AS_PH2 : AS_SH2]);
#pragma unroll
          for (int j = 0; j < MAXO; ++j) hpart[j] = 0.f;
        } else if (d.post == SP_PATCH_SAFE) {
          // action_safe' = actor_safe.distr(obs).rsample(): same net output as the eval action                      src/ssac.py:488-490
          float out[MAXO] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < A; ++j) { out[j] = mu_s[j]; out[A + j] = raw_s[j]; }
          (void)rsample_row<A>(out, p.n_safe, gr, e.valid, pt, psd, peps, psg);
          if (dbg) dbg[11] = pt[0];
          patch_action<A>(xs0, e, S, pt);
        }
        if (o != p.n_ops - 1) epi_op_done(e);
      }
    }
    // ---- per-CTA results ---------------------------------------------------------------------------------------------------
    named_bar_sync(1, EPI_THREADS);
    if (threadIdx.x == 0) {
      for (int k = 0; k < 3; ++k) p.loss_part[4 * blockIdx.x + k] = lacc[k];
      p.loss_part[4 * blockIdx.x + 3] = 0.0;
    }
    for (int i = threadIdx.x; i < p.nv * HID; i += EPI_THREADS) p.gacc_out[(int64_t)blockIdx.x * p.nv * HID + i] = gacc[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == ISSUER) tmem_dealloc(tmem, 512);
}

// losses: multiplier  [0] = (-0.5 * sum(lams_safe * penalty) + sum((lams_unsafe - target)^2)) / B          src/ssac.py:561-565
//         actor       [0] actor loss, [1] = [5] = -alpha * mean(log_prob + target_entropy), [2] safe-actor loss   src/ssac.py:464-503
static __global__ void solver_loss_finalize_kernel(const double* partials, int nblocks, int mode, double inv_bg, const float* log_alpha, float* losses) {
  double s[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double v = 0;
    for (int b = threadIdx.x; b < nblocks; b += 32) v += partials[4 * b + k];
    s[k] = warp_sum_d(v);
  }
  if (threadIdx.x == 0) {
    if (mode == MODE_MULT) { losses[0] = (float)((-0.5 * s[0] + s[1]) * inv_bg); }
    else {
      const float alpha = expf(*log_alpha);
      const float M = (float)(s[2] * inv_bg);
      losses[0] = (float)(s[0] * inv_bg); losses[1] = -alpha * M; losses[2] = (float)(s[1] * inv_bg); losses[5] = -alpha * M;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct SPlan {
  int64_t Bpad; int Kx, n_tiles, grid, nv, n_sv, n_big, n_small;
  int64_t img_bytes; int ctab_floats;
};
static SPlan solver_plan(int mode, int64_t B, int S, int A, int C) {
  SPlan pl;
  pl.Bpad = (B + CLUSTER * TILE - 1) / (CLUSTER * TILE) * (CLUSTER * TILE);
  pl.Kx = round_up(S + A, 16);
  pl.n_tiles = (int)(pl.Bpad / TILE);
  pl.grid = std::min(pl.n_tiles, 148) / CLUSTER * CLUSTER;
  pl.nv = mode == MODE_MULT ? M_NV : a_nv(A);
  pl.n_sv = mode == MODE_MULT ? (int)MS_COUNT : (int)AS_COUNT;
  pl.n_big = mode == MODE_MULT ? 1 : 2; pl.n_small = pl.n_big;
  // images (upper bound of both modes): 5 first-layer (kp = Kx) + 7 hidden forward + 6 transposed
  pl.img_bytes = (int64_t)5 * HID * (pl.Kx + KBIAS) * 2 + (int64_t)7 * HID * (HID + KBIAS) * 2 + (int64_t)6 * HID * HID * 2;
  pl.ctab_floats = (6 * A + 2 * C + 2) * HID + 128;
  return pl;
}
static void dw_splits_n(const SPlan& pl, int n_slabs, int& ks_big, int& ks_small) {
  const double wb = 64.0, ws = 32.0 + pl.Kx / 8.0;
  const double unit = 148.0 / (pl.n_big * wb + pl.n_small * ws);
  ks_big = std::max(1, std::min(n_slabs, (int)(unit * wb)));
  ks_small = std::max(1, std::min(n_slabs, (148 - pl.n_big * ks_big) / pl.n_small));
}
int64_t solver_ws_bytes(int mode, int64_t B, int S, int A, int C) {
  SPlan pl = solver_plan(mode, B, S, A, C);
  int ksb, kss; dw_splits_n(pl, (int)(pl.Bpad / DW_ROWS), ksb, kss);
  int64_t b = 0;
  b += align_up(pl.img_bytes, 256) + align_up((int64_t)pl.ctab_floats * 4, 256);
  b += pl.n_sv * align_up(pl.Bpad * HID * 2, 256) + align_up(pl.Bpad * pl.Kx * 2, 256);
  b += align_up((int64_t)148 * pl.nv * HID * 4, 256) + align_up(148 * 4 * 8, 256);
  b += align_up(((int64_t)pl.n_big * ksb * HID * HID + (int64_t)pl.n_small * kss * HID * pl.Kx) * 4, 256);
  return b + 4096;
}

template <int A, int C, int MODE>
static int launch_solver(const SolverParams& fp, int& grid, size_t smem, cudaStream_t st) {
  auto k = solver_fused_kernel<A, C, MODE>;
  static int max_clusters = 0;
  if (!max_clusters) {
    DRPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(F_THREADS); cfg.dynamicSmemBytes = 232448 - 1024;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CLUSTER; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = 148 / CLUSTER; }
    max_clusters = n;
  }
  grid = std::min(grid, max_clusters * CLUSTER);
  DRPO_LAUNCH(k, grid, F_THREADS, smem, st, fp);
  return DRPO_OK;
}
template <int MODE>
static int launch_solver_ac(int A, int C, const SolverParams& fp, int& grid, size_t smem, cudaStream_t st) {
  if (A == 1 && C == 1) return launch_solver<1, 1, MODE>(fp, grid, smem, st);
  if (A == 1 && C == 2) return launch_solver<1, 2, MODE>(fp, grid, smem, st);
  if (A == 1 && C == 4) return launch_solver<1, 4, MODE>(fp, grid, smem, st);
  if (A == 2 && C == 1) return launch_solver<2, 1, MODE>(fp, grid, smem, st);
  if (A == 2 && C == 2) return launch_solver<2, 2, MODE>(fp, grid, smem, st);
  return launch_solver<2, 4, MODE>(fp, grid, smem, st);
}

// builds the op tables, the weight images and the constant table of one call
struct Builder {
  SolverParams& fp; PackTable pt; CopyTable ct; int Kx;
  int64_t img_off = 0; int ctab_off = 0; int n_ops = 0;
  struct Seen { const float* w; int transposed; uint32_t off; };
  std::vector<Seen> images;
  struct SeenC { const float* src; int n; int stride; int off; };
  std::vector<SeenC> consts;
  Builder(SolverParams& f, int kx) : fp(f), Kx(kx) { pt.n = 0; ct.n = 0; }

  uint32_t image(const drpo_linear& l, bool transposed) {          // one image per (weight, orientation): the frozen critics are used 2-3 x
    for (const Seen& s : images) if (s.w == l.w && s.transposed == (transposed ? 1 : 0)) return s.off;
    PackEntry& e = pt.e[pt.n++];
    e.W = l.w; e.transposed = transposed ? 1 : 0; e.bias = transposed ? nullptr : l.b;
    if (!transposed) { e.n_real = l.out_dim; e.k_real = l.in_dim; e.kp = l.in_dim == HID ? HID : Kx; }
    else { e.n_real = l.in_dim; e.k_real = l.out_dim; e.kp = HID; }
    e.dst = img_off / 2;
    const uint32_t off = (uint32_t)img_off;
    img_off += (int64_t)HID * (e.kp + (transposed ? 0 : KBIAS)) * 2;
    images.push_back({l.w, transposed ? 1 : 0, off});
    return off;
  }
  int cst(const float* src, int n, int stride = 1) {
    for (const SeenC& s : consts) if (src && s.src == src && s.n == n && s.stride == stride) return s.off;
    CopyEntry& e = ct.e[ct.n++]; e.src = src; e.n = n; e.dst = ctab_off; e.stride = stride;
    const int off = ctab_off; ctab_off += (n + 3) & ~3;
    consts.push_back({src, n, stride, off});
    return off;
  }
  void fwd(const drpo_linear& l, int a_src, int out_region, int save, bool wait_all = false, int act = 0) {
    FOp& op = fp.op[n_ops];
    op.w_off[0] = image(l, false); op.kp = (uint16_t)(l.in_dim == HID ? HID : Kx); op.a_src[0] = (uint8_t)a_src; op.parts = 1; op.bias = 1; op.early = 0;
    EOp& e = fp.eop[n_ops];
    e.out_region = (uint8_t)out_region; e.save = (uint8_t)save; e.wait_all = wait_all ? 1 : 0; e.act = (uint8_t)act; e.bias_slot = NO_SLOT;
    ++n_ops;
  }
  int head(const drpo_linear& h, int post) {
    EOp& e = fp.eop[n_ops - 1];
    e.no = (uint8_t)h.out_dim; e.hw_off = cst(h.w, h.out_dim * HID); e.hb_off = cst(h.b, h.out_dim); e.post = (uint8_t)post;
    return e.hw_off;
  }
  void bwd(const drpo_linear& l, int a_src, int hsave, int bias_slot, int save, int out_region, bool wait_all, int act = 0) {
    FOp& op = fp.op[n_ops];
    op.w_off[0] = image(l, true); op.kp = HID; op.a_src[0] = (uint8_t)a_src; op.parts = 1;
    EOp& e = fp.eop[n_ops];
    e.backward = 1; e.hsave = (uint8_t)hsave; e.bias_slot = (uint8_t)bias_slot; e.save = (uint8_t)save; e.out_region = (uint8_t)out_region;
    e.wait_all = wait_all ? 1 : 0; e.act = (uint8_t)act;
    ++n_ops;
  }
  // first-layer action columns of a frozen critic as `A` fp32 rows: the op just added also evaluates d loss / d action
  void action_dot(const drpo_linear& l0, int S, int A, int post) {
    EOp& e = fp.eop[n_ops - 1];
    int off0 = -1;
    for (int j = 0; j < A; ++j) { const int off = cst(l0.w + S + j, HID, l0.in_dim); if (j == 0) off0 = off; }
    e.hw_off = off0; e.no = (uint8_t)A; e.post = (uint8_t)post;
  }
  // constraint_critic(obs, a, uncertainty=True): trunk0, trunk1, mean head, log-std head.  stash: keep m1 / l1 in R0 / R1 (backward)
  void qc_forward(const drpo_qc& q, int sv_t1, int sv_t2, bool stash, int post) {
    fwd(q.trunk0, A_XS0, 1, sv_t1); fwd(q.trunk1, A_R0, 2, sv_t2);
    fwd(q.mean0, A_R1, stash ? 1 : 0, 0); fp.hw_cm = head(q.mean1, SP_KEEP);
    fwd(q.lstd0, A_R1, stash ? 2 : 0, 0, stash); fp.hw_cl = head(q.lstd1, post);
  }
  // dX chain of the frozen constraint critic from (dm1 in R0, dl1 in R1) down to the action columns
  void qc_backward(const drpo_qc& q, int sv_t1, int sv_t2, int S, int A, int post) {
    FOp& op = fp.op[n_ops];
    op.w_off[0] = image(q.mean0, true); op.w_off[1] = image(q.lstd0, true);
    op.kp = HID; op.a_src[0] = A_R0; op.a_src[1] = A_R1; op.parts = 2;
    EOp& e = fp.eop[n_ops];
    e.backward = 1; e.hsave = (uint8_t)sv_t2; e.bias_slot = NO_SLOT; e.save = 0; e.out_region = 1; e.wait_all = 1;
    ++n_ops;
    bwd(q.trunk1, A_R0, sv_t1, NO_SLOT, 0, 0, false);
    action_dot(q.trunk0, S, A, post);
  }
};

struct DwBuild {
  DwParams dp; ReduceTable rt[2]; int cta = 0; float* part; const SPlan& pl; int ksb, kss;
  DwBuild(const SPlan& p, float* dw_part, int n_slabs, int* err_flag) : part(dw_part), pl(p) {
    memset(&dp, 0, sizeof(dp)); rt[0].n = rt[1].n = 0;
    dw_splits_n(pl, n_slabs, ksb, kss);
    dp.Bpad = pl.Bpad; dp.n_slabs = n_slabs; dp.err_flag = err_flag; dp.lbo = 128; dp.sbo = DW_PANEL;
  }
  void job(int t, const float* P, const __nv_bfloat16* dH, const __nv_bfloat16* Hm, const drpo_linear& l) {
    const bool big = l.in_dim == HID;
    DwJob& j = dp.job[dp.n_jobs++];
    j.a = dH; j.b = Hm; j.b_octets = big ? 32 : pl.Kx / 8; j.cta0 = cta; j.ksplit = big ? ksb : kss; j.partial = part;
    const int N = j.b_octets * 8;
    ReduceEntry& r = rt[t].e[rt[t].n++];
    r.dst = (int64_t)(l.w - P); r.src = part; r.n_src = j.ksplit; r.stride = (int64_t)HID * N; r.rows = HID; r.cols = l.in_dim; r.ld = N;
    cta += j.ksplit; part += (int64_t)j.ksplit * HID * N;
  }
  void vec(int t, const float* P, const float* dst_param, const float* gacc_out, int slot, int off, int n) {
    ReduceEntry& r = rt[t].e[rt[t].n++];
    r.dst = (int64_t)(dst_param - P); r.src = gacc_out + slot * HID + off; r.n_src = pl.grid; r.stride = (int64_t)pl.nv * HID; r.rows = 1; r.cols = n; r.ld = n;
  }
  int launch(cudaStream_t st, float* const* grads, int n_tables) {
    static bool attr_done = false;
    const size_t dsm = DW_STAGES * DW_STAGE_BYTES + sizeof(DwSmem);
    if (!attr_done) { DRPO_CUDA_OK(cudaFuncSetAttribute(critic_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm)); attr_done = true; }
    DRPO_LAUNCH(critic_dw_kernel, cta, DW_THREADS, dsm, st, dp);
    for (int t = 0; t < n_tables; ++t) {
      dim3 grid(64, rt[t].n);
      DRPO_LAUNCH(critic_grad_reduce_kernel, grid, 256, 0, st, rt[t], grads[t]);
    }
    return DRPO_OK;
  }
};

struct SolverWs {
  uint8_t* img; float* ctab; __nv_bfloat16* sv[16]; __nv_bfloat16* x; float* gacc_out; double* loss_part; float* dw_part;
};
static bool solver_carve(const SPlan& pl, void* ws, int64_t ws_bytes, int n_slabs, SolverWs& w, int64_t& need) {
  int ksb, kss; dw_splits_n(pl, n_slabs, ksb, kss);
  Arena ar(ws, ws_bytes);
  w.img = ar.take<uint8_t>(pl.img_bytes);
  w.ctab = ar.take<float>(pl.ctab_floats);
  for (int i = 0; i < pl.n_sv; ++i) w.sv[i] = ar.take<__nv_bfloat16>(pl.Bpad * HID);
  w.x = ar.take<__nv_bfloat16>(pl.Bpad * pl.Kx);
  w.gacc_out = ar.take<float>((int64_t)148 * pl.nv * HID);
  w.loss_part = ar.take<double>(148 * 4);
  w.dw_part = ar.take<float>((int64_t)pl.n_big * ksb * HID * HID + (int64_t)pl.n_small * kss * HID * pl.Kx);
  need = ar.off;
  return ar.ok();
}
static int solver_dims_ok(const char* who, int S, int A, int C, int hidden_ok) {
  DRPO_CHECK_ARG(hidden_ok, "%s(bf16): the fused kernel needs hidden width 256 in every net", who);
  DRPO_CHECK_ARG(S + A <= 64 && S + 1 <= 64, "%s(bf16): state_dim + action_dim must be <= 64", who);
  DRPO_CHECK_ARG((A == 1 || A == 2) && (C == 1 || C == 2 || C == 4), "%s(bf16): fused kernel is built for action_dim 1-2, con_dim 1/2/4", who);
  return DRPO_OK;
}
static int solver_run(int mode, Builder& b, SolverParams& fp, SPlan& pl, const SolverWs& w, int A, int C, int* err_flag, cudaStream_t st) {
  if (b.n_ops > S_MAX_OPS || b.img_off > pl.img_bytes || b.ctab_off > pl.ctab_floats || b.pt.n > 28 || b.ct.n > 48) {
    set_error("solver step(bf16): internal plan mismatch (%d ops, %lld image bytes, %d consts)", b.n_ops, (long long)b.img_off, b.ctab_off);
    return DRPO_ERR_ARG;
  }
  {
    dim3 grid(64, b.pt.n + b.ct.n);
    DRPO_LAUNCH(pack_gather_kernel, grid, 256, 0, st, b.pt, reinterpret_cast<__nv_bfloat16*>(w.img), b.ct, w.ctab);
  }
  fp.n_ops = b.n_ops; fp.sch.split = 0; fp.sch.n[0] = (uint8_t)b.n_ops;
  for (int i = 0; i < b.n_ops; ++i) fp.sch.order[0][i] = (uint8_t)i;
  fp.wimg = w.img; fp.ctab = w.ctab; fp.ctab_floats = b.ctab_off;
  fp.sv[0] = nullptr;
  for (int i = 0; i < pl.n_sv; ++i) fp.sv[1 + i] = w.sv[i];
  fp.gacc_out = w.gacc_out; fp.nv = pl.nv; fp.loss_part = w.loss_part; fp.err_flag = err_flag; fp.dbg = g_solver_dbg;
  const size_t fixed = (size_t)2 * TILE * pl.Kx * 2 + TILE * KBIAS * 2 + (size_t)((b.ctab_off + 3) & ~3) * 4 + (size_t)pl.nv * HID * 4 + NGROUPS * TILE * 16 + sizeof(FusedSmem);
  int stages = (int)((232448 - 1024 - fixed) / CHUNK_BYTES);
  if (stages > 6) stages = 6;
  if (stages < 2) { set_error("solver step(bf16): shared-memory budget exceeded"); return DRPO_ERR_ARG; }
  fp.stages = stages;
  const size_t smem = fixed + (size_t)stages * CHUNK_BYTES;
  return mode == MODE_MULT ? launch_solver_ac<MODE_MULT>(A, C, fp, pl.grid, smem, st) : launch_solver_ac<MODE_ACTOR>(A, C, fp, pl.grid, smem, st);
}

// phase 1 of drpo_multiplier_step in DRPO_PREC_BF16: fills a.grads and a.losses[0]
int multiplier_phase1(const drpo_multiplier_args& a, int* err_flag) {
  const int64_t B = a.batch_size; const int S = a.state_dim, A = a.action_dim, C = a.con_dim;
  cudaStream_t st = (cudaStream_t)a.stream;
  int rc = solver_dims_ok("drpo_multiplier_step", S, A, C, a.actor->l0.out_dim == HID && a.actor->l1.out_dim == HID && a.actor_safe->l0.out_dim == HID &&
                          a.actor_safe->l1.out_dim == HID && a.qc->trunk0.out_dim == HID && a.lam.l0.out_dim == HID && a.lam.l1.out_dim == HID);
  if (rc) return rc;
  SPlan pl = solver_plan(MODE_MULT, B, S, A, C);
  const int n_slabs = (int)(pl.Bpad / DW_ROWS);
  SolverWs w; int64_t need;
  if (!solver_carve(pl, a.workspace, a.workspace_bytes, n_slabs, w, need)) {
    set_error("drpo_multiplier_step(bf16): workspace too small (%lld needed, %lld given)", (long long)need, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE;
  }
  SolverParams fp; memset(&fp, 0, sizeof(fp));
  Builder b(fp, pl.Kx);
  fp.zero_off = b.cst(nullptr, 4);
  // action = actor.distr(obs).rsample()                                                     src/ssac.py:530-531
  b.fwd(a.actor->l0, A_XS0, 1, 0); b.fwd(a.actor->l1, A_R0, 0, 0); b.head(a.actor->l2, SP_POLICY);
  // penalty from constraint_critic(obs, action, uncertainty)                                src/ssac.py:533-543
  b.qc_forward(*a.qc, 0, 0, false, SP_QCUB_PEN);
  // action_safe = actor_safe.act(obs, eval=True); safe_Qc                                   src/ssac.py:546-548
  b.fwd(a.actor_safe->l0, A_XS0, 1, 0); b.fwd(a.actor_safe->l1, A_R0, 0, 0); b.head(a.actor_safe->l2, SP_POLICY_EVAL);
  b.qc_forward(*a.qc, 0, 0, false, SP_QCUB_SAFE);
  // lams = multiplier(obs, safe_Qc): two tanh layers + head; loss; backward                 src/ssac.py:549-565, 100-111
  b.fwd(a.lam.l0, A_XS1, 1, MS_H1, false, 1);
  b.fwd(a.lam.l1, A_R0, 2, 0, false, 1); fp.hw_lam2 = b.head(a.lam.l2, SP_MULT);              // h2 stashed in R1
  b.bwd(a.lam.l1, A_R1, MS_H1, MSLOT_B0, MS_DH1, 0, false, 1);                                // dh1 = (dh2 W1) * (1 - h1^2)
  fp.obs = a.obs;
  fp.n_actor = make_noise(a.eps_actor, A, a.seed, TAG_MULT_ACTOR, a.noise_step, a.row_id_offset);
  fp.n_safe = make_noise(nullptr, 0, 0, 0, 0);
  fp.ratio = (float)a.std_ratio; fp.thr = (float)a.constraint_threshold; fp.pen_lb = (float)a.penalty_lb; fp.pen_ub = (float)a.penalty_ub;
  fp.ub = (float)a.upper_bound; fp.lam_eps = (float)a.lam_epsilon; fp.inv_bg = (float)(1.0 / (double)a.global_batch_size);
  fp.B = B; fp.Bpad = pl.Bpad; fp.S = S; fp.A = A; fp.C = C; fp.D = S + A; fp.Kx = pl.Kx; fp.n_tiles = pl.n_tiles;
  fp.x_aug = w.x; fp.x_obs = nullptr;
  if ((rc = solver_run(MODE_MULT, b, fp, pl, w, A, C, err_flag, st))) return rc;

  DwBuild dw(pl, w.dw_part, n_slabs, err_flag);
  const float* P = a.params;
  dw.job(0, P, fp.sv[MS_DH2], fp.sv[MS_H1], a.lam.l1);
  dw.job(0, P, fp.sv[MS_DH1], w.x, a.lam.l0);
  dw.vec(0, P, a.lam.l2.w, w.gacc_out, MSLOT_W2, 0, HID); dw.vec(0, P, a.lam.l2.b, w.gacc_out, MSLOT_SCAL, 0, 1);
  dw.vec(0, P, a.lam.l1.b, w.gacc_out, MSLOT_B1, 0, HID); dw.vec(0, P, a.lam.l0.b, w.gacc_out, MSLOT_B0, 0, HID);
  float* grads[1] = {a.grads};
  if ((rc = dw.launch(st, grads, 1))) return rc;
  DRPO_LAUNCH(solver_loss_finalize_kernel, 1, 32, 0, st, w.loss_part, pl.grid, (int)MODE_MULT, 1.0 / (double)a.global_batch_size, (const float*)nullptr, a.losses);
  return DRPO_OK;
}

// phase 1 of drpo_actor_step in DRPO_PREC_BF16: fills grads_actor, grads_safe and losses[0..2], [5]
int actor_phase1(const drpo_actor_args& a, int* err_flag) {
  const int64_t B = a.batch_size; const int S = a.state_dim, A = a.action_dim, C = a.con_dim;
  cudaStream_t st = (cudaStream_t)a.stream;
  int rc = solver_dims_ok("drpo_actor_step", S, A, C, a.actor.l0.out_dim == HID && a.actor.l1.out_dim == HID && a.actor_safe.l0.out_dim == HID &&
                          a.actor_safe.l1.out_dim == HID && a.q->l0.out_dim == HID && a.q->l1.out_dim == HID && a.qc->trunk0.out_dim == HID &&
                          a.lam->l0.out_dim == HID && a.lam->l1.out_dim == HID);
  if (rc) return rc;
  SPlan pl = solver_plan(MODE_ACTOR, B, S, A, C);
  const int n_slabs = (int)(pl.Bpad / DW_ROWS);
  SolverWs w; int64_t need;
  if (!solver_carve(pl, a.workspace, a.workspace_bytes, n_slabs, w, need)) {
    set_error("drpo_actor_step(bf16): workspace too small (%lld needed, %lld given)", (long long)need, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE;
  }
  SolverParams fp; memset(&fp, 0, sizeof(fp));
  Builder b(fp, pl.Kx);
  fp.zero_off = b.cst(nullptr, 4);
  // no grad: action_safe = actor_safe.act(obs, eval=True); safe_Qc; lams = multiplier(obs, safe_Qc)          src/ssac.py:472-478
  b.fwd(a.actor_safe.l0, A_XS0, 1, AS_SH1); b.fwd(a.actor_safe.l1, A_R0, 0, AS_SH2); fp.hw_safe2 = b.head(a.actor_safe.l2, SP_SAFE_FWD);
  b.qc_forward(*a.qc, 0, 0, false, SP_QCUB_SAFE);
  b.fwd(a.lam->l0, A_XS1, 1, 0, false, 1); b.fwd(a.lam->l1, A_R0, 0, 0, false, 1); b.head(a.lam->l2, SP_LAM);
  // action = actor.distr(obs).rsample(), log_prob                                                            src/ssac.py:459-461
  b.fwd(a.actor.l0, A_XS0, 1, AS_PH1); b.fwd(a.actor.l1, A_R0, 0, AS_PH2); fp.hw_actor2 = b.head(a.actor.l2, SP_POLICY);
  // actor_Q = critic.random_choice(obs, action) and its dX chain                                             src/ssac.py:462
  b.fwd(a.q->l0, A_XS0, 1, AS_QH1); b.fwd(a.q->l1, A_R0, 2, 0); fp.hw_q2 = b.head(a.q->l2, SP_QK);
  b.bwd(a.q->l1, A_R1, AS_QH1, NO_SLOT, 0, 0, false); b.action_dot(a.q->l0, S, A, SP_DA_Q);
  // actor_Qc = max_c constraint_critic(obs, action, uncertainty) and its dX chain                            src/ssac.py:468-469
  b.qc_forward(*a.qc, AS_CT1, AS_CT2, true, SP_QCUB_GRAD1);
  b.qc_backward(*a.qc, AS_CT1, AS_CT2, S, A, SP_DA_ACTOR);                                  // + rsample backward, actor head backward -> R0
  b.bwd(a.actor.l1, A_R0, AS_PH1, aslot_b0(0, A), AS_PDH1, 0, false); fp.eop[b.n_ops - 1].post = SP_PATCH_SAFE;
  // safe actor: Qc(obs, action_safe') and its dX chain                                                       src/ssac.py:488-492
  b.qc_forward(*a.qc, AS_CT1, AS_CT2, true, SP_QCUB_GRAD2);
  b.qc_backward(*a.qc, AS_CT1, AS_CT2, S, A, SP_DA_SAFE);
  b.bwd(a.actor_safe.l1, A_R0, AS_SH1, aslot_b0(1, A), AS_SDH1, 0, false);
  fp.obs = a.obs;
  fp.n_actor = make_noise(a.eps_actor, A, a.seed, TAG_ACTOR_ACTOR, a.noise_step, a.row_id_offset);
  fp.n_safe = make_noise(a.eps_safe, A, a.seed, TAG_ACTOR_SAFE, a.noise_step, a.row_id_offset);
  fp.log_alpha = a.log_alpha;
  fp.ratio = (float)a.std_ratio; fp.ub = (float)a.multiplier_ub; fp.inv_bg = (float)(1.0 / (double)a.global_batch_size);
  fp.target_entropy = (float)a.target_entropy;
  fp.B = B; fp.Bpad = pl.Bpad; fp.S = S; fp.A = A; fp.C = C; fp.D = S + A; fp.Kx = pl.Kx; fp.n_tiles = pl.n_tiles;
  fp.x_obs = w.x; fp.x_aug = nullptr;
  if ((rc = solver_run(MODE_ACTOR, b, fp, pl, w, A, C, err_flag, st))) return rc;

  DwBuild dw(pl, w.dw_part, n_slabs, err_flag);
  const int scal = aslot_scal(A);
  for (int n = 0; n < 2; ++n) {
    const drpo_mlp3& net = n == 0 ? a.actor : a.actor_safe;
    const float* P = n == 0 ? a.params_actor : a.params_safe;
    dw.job(n, P, fp.sv[n == 0 ? AS_PDH2 : AS_SDH2], fp.sv[n == 0 ? AS_PH1 : AS_SH1], net.l1);
    dw.job(n, P, fp.sv[n == 0 ? AS_PDH1 : AS_SDH1], w.x, net.l0);
    for (int o = 0; o < 2 * A; ++o) dw.vec(n, P, net.l2.w + o * HID, w.gacc_out, aslot_w2(n, A) + o, 0, HID);
    dw.vec(n, P, net.l2.b, w.gacc_out, scal, 4 * n, 2 * A);
    dw.vec(n, P, net.l1.b, w.gacc_out, aslot_b1(n, A), 0, HID); dw.vec(n, P, net.l0.b, w.gacc_out, aslot_b0(n, A), 0, HID);
  }
  float* grads[2] = {a.grads_actor, a.grads_safe};
  if ((rc = dw.launch(st, grads, 2))) return rc;
  DRPO_LAUNCH(solver_loss_finalize_kernel, 1, 32, 0, st, w.loss_part, pl.grid, (int)MODE_ACTOR, 1.0 / (double)a.global_batch_size, (const float*)a.log_alpha, a.losses);
  return DRPO_OK;
}

}  // namespace cu
}  // namespace drpo
