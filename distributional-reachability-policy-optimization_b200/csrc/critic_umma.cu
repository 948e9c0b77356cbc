// DRPO_PREC_BF16 critic step (SSAC.update_critic, src/ssac.py:437-456): the forward of all nine passes, the targets, the
// losses and the dX half of the backward pass run in ONE persistent tcgen05 kernel per update; the dW half is a second
// tcgen05 kernel (split-K over the batch); a table-driven reduce kernel assembles the flat gradient arena.
//
//   critic_fused_kernel   one CTA per SM, 128 batch rows per tile.  Per tile, 24 dense "ops" (one nn.Linear each, the Qc
//       trunk back-propagation being one op with two K parts) run as tcgen05.mma (bf16 x bf16 -> fp32) into four 64-column
//       TMEM accumulators; four epilogue groups (one per accumulator / 64-column slab, thread = batch row) apply bias + ReLU,
//       keep the activation in TMEM as the next op's A operand (TS-mode MMA), evaluate the narrow heads (policy mean/log-std,
//       Q, Qc mean/log-std) with CUDA-core dot products, and - once the row's targets are known - turn the per-row loss
//       gradients into dH tiles that feed the transposed-weight MMAs.  Weights (forward and transposed images, packed once
//       per update) stream through a shared-memory ring by TMA bulk copies.  What leaves the SM: the activations and
//       activation gradients the dW kernel needs (bf16, "octet" layout below), per-CTA column sums (bias / head-weight
//       gradients, warp-shuffle butterflies + shared-memory accumulators) and the two loss partials.
//   critic_dw_kernel      dW[out,in] = sum_rows dH[row,out] * H[row,in]: both operands are read straight from the octet
//       layout as MN-major UMMA operands (K = batch rows), 256 x N fp32 accumulated in TMEM, one partial per CTA.
//   critic_grad_reduce_kernel   sums the split-K partials / per-CTA column sums into the gradient arena.
//
// Slab-octet layout of a saved [rows, F] bf16 matrix (umma_ops.cuh: oct_index): per 64-row slab one [64 rows][8] panel per 8 features,
// the slab's panels contiguous.  The fused kernel's threads (one per row) write 16-byte vectors that are contiguous across a warp,
// one panel of a slab is a contiguous 1 KB block = 8 no-swizzle MN-major core matrices, and a whole slab (all panels) is what one
// stage of the dW kernel consumes: one bulk copy per operand.
#include "critic_umma_api.h"
#include "umma_ops.cuh"

namespace drpo {
namespace cu {


// ---------------------------------------------------------------------------------------------------------------
// the fused forward / loss / dX kernel
// ---------------------------------------------------------------------------------------------------------------
template <int A, int C>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(F_THREADS, 1) critic_fused_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stages = p.stages;
  uint8_t* ring = smem;
  uint8_t* xs0 = ring + (size_t)stages * CHUNK_BYTES;
  uint8_t* xs1 = xs0 + TILE * p.Kx * 2;
  uint8_t* ones = xs1 + TILE * p.Kx * 2;                    // [128 rows][16 k] K-major, columns 0 and 1 are 1.0: A operand of the bias MMA
  float* ctab = reinterpret_cast<float*>(ones + TILE * KBIAS * 2);
  float* gacc = ctab + ((p.ctab_floats + 3) & ~3);
  float4* hp = reinterpret_cast<float4*>(gacc + p.nv * HID);
  float4* polres = hp + NGROUPS * TILE;                     // [128 rows] sampled action + log-prob of the policy post step
  FusedSmem* sm = reinterpret_cast<FusedSmem*>(polres + TILE);
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
  for (int i = threadIdx.x; i < TILE * KBIAS; i += F_THREADS) {          // element (row, k) at (k/8)*2048 + row*16 + (k%8)*2
    const int k = (i >> 10) * 8 + (i & 7);
    reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16_rn(k < 2 ? 1.f : 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  // the two CTAs of a cluster walk through the same number of tiles (tile pairs are dealt to clusters round-robin) in
  // lock step, coupled by the weight ring: each CTA fetches half of every chunk and multicasts it to both
  const uint32_t crank = cluster_ctarank();
  const int n_clusters = (int)gridDim.x / CLUSTER, cid = (int)blockIdx.x / CLUSTER;
  const int my_jobs = sched_my_jobs(p.sch, p.n_tiles, cid, n_clusters);

  if (warp == PRODUCER) {
    fused_producer(p, sm, ring, my_jobs, cid, n_clusters, crank, err);
  } else if (warp == ISSUER) {
    fused_issuer(p, sm, ring, xs0, xs1, ones, tmem, my_jobs, cid, n_clusters, lane, err, p.prof);
  } else {
    // ---- epilogue groups ----------------------------------------------------------------------------------------------------
    Epi e;
    e.sm = sm; e.ctab = ctab; e.gacc = gacc; e.hp = hp; e.g = warp >> 2; e.lane = lane; e.row = (warp & 3) * 32 + lane;
    e.tm = tmem + ((uint32_t)((warp & 3) * 32) << 16); e.it = 0; e.arr = 0; e.err = err; e.Bpad = p.Bpad;
    e.prof = (blockIdx.x == 0 && (threadIdx.x & 127) == 0) ? p.prof : nullptr;          // first thread of every group
    const int S = p.S, D = p.D;
    const float alpha = expf(*p.log_alpha);
    // Loss sums and head-bias gradients go straight to shared memory (warp sum + one atomic per warp at the two sites that produce
    // them) instead of living in registers across the op loop: at 96 registers per thread a loop-carried value is a spill to L2.
    // lacc = two doubles in the unused tail of the scalar slot of the column-sum accumulators.
    float* scacc = gacc + SLOT_SCAL * HID;
    double* lacc = reinterpret_cast<double*>(scacc + 16);
    auto add_loss = [&](int k, double v) { v = warp_sum_d(v); if (lane == 0) atomicAdd(&lacc[k], v); };
    auto add_scal = [&](int k, float v) { v = warp_sum(v); if (lane == 0) atomicAdd(&scacc[k], v); };

    for (int t = 0; t < my_jobs; ++t) {
      const int job = cid + t * n_clusters, chain = sched_chain(p.sch, job, n_clusters), n_chain = p.sch.n[chain];
      const int tile = CLUSTER * sched_pair(p.sch, job) + (int)crank;
      e.grow = (int64_t)tile * TILE + e.row;
      e.valid = e.grow < p.B;
      const int64_t gr = e.valid ? e.grow : 0;
      // ---- stage the tile's inputs: xs0 = [next_obs, 0], xs1 = [obs, act] (bf16, K-major), x_sa octets to global -------
      for (int j = e.g; j < (p.Kx >> 3); j += NGROUPS) {
        uint32_t w0[4], w1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a0[2], a1[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k = j * 8 + q * 2 + u;
            a0[u] = (e.valid && k < S) ? p.next_obs[gr * S + k] : 0.f;
            a1[u] = !e.valid ? 0.f : (k < S ? p.obs[gr * S + k] : (k < D ? p.act[gr * A + (k - S)] : 0.f));
          }
          w0[q] = pack_bf16(a0[0], a0[1]); w1[q] = pack_bf16(a1[0], a1[1]);
        }
        *reinterpret_cast<uint4*>(xs0 + j * 2048 + e.row * 16) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        const uint4 v1 = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        *reinterpret_cast<uint4*>(xs1 + j * 2048 + e.row * 16) = v1;
        *reinterpret_cast<uint4*>(p.x_sa + oct_index(e.grow, j, p.Kx >> 3)) = v1;
      }
      const float rew = e.valid ? p.rew[gr] : 0.f;
      const float dn = (e.valid && p.done[gr]) ? 1.f : 0.f;
      float cvr[C];
#pragma unroll
      for (int c = 0; c < C; ++c) cvr[c] = e.valid ? p.cv[gr * C + c] : 0.f;
      fence_proxy_async();
      epi_op_done(e);                                                     // op 0 may start

      // ---- the tile's 24 ops: one generic forward / backward epilogue, then the op's post step -----------------------------
      float hpart[MAXO] = {0.f, 0.f, 0.f, 0.f}, hkeep[MAXO] = {0.f, 0.f, 0.f, 0.f};
      int hb_keep = 0;
      float a1[A], a2[A], logp = 0.f, qt0 = 0.f, q_target = 0.f, nqc[C];
#pragma unroll
      for (int j = 0; j < A; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
#pragma unroll
      for (int c = 0; c < C; ++c) nqc[c] = 0.f;
#pragma unroll 1
      for (int oi = 0; oi < n_chain; ++oi) {
        const EOp d = p.eop[p.sch.order[chain][oi]];
        const uint32_t region = d.out_region == 0 ? 0u : (d.out_region == 1 ? TM_R0 : TM_R1);
        if (!d.backward) epi_forward_halfwise(e, region, p.sv[d.save], d.hw_off, d.no, hpart, d.wait_all != 0);
        else epi_backward(e, p.sv[d.hsave], d.bias_slot, p.sv[d.save], region, d.wait_all != 0);
        if (d.post == POST_POLICY0 || d.post == POST_POLICY1) {
          // squashed-Gaussian sample of the next action (+ log-prob)                       src/ssac.py:286-288, 340-341
          const bool first = d.post == POST_POLICY0;
          float out[MAXO];
          head_combine(e, hpart, d.hb_off, out);
          // one warp per group does the transcendental work for 32 of the tile's rows (group g: rows 32g.., four warps on four
          // different schedulers) and shares the result through shared memory; all 512 threads computing their own row
          // redundantly cost ~4x the issue slots on the critical path
          if ((warp & 3) == e.g) {
            float lp = 0.f, an[A];
#pragma unroll
            for (int j = 0; j < A; ++j) {
              const float mu = out[j], raw = out[A + j];
              const float log_std = -6.f + 10.f * sigmoid_f(raw);
              const float sd = expf(log_std);
              const float eps = !e.valid ? 0.f : (first ? p.n_actor.get(gr, j) : p.n_safe.get(gr, j));
              const float x = fmaf(eps, sd, mu);
              an[j] = tanhf(x);
              const float ladj = 2.f * (0.69314718055994531f - x - softplus_f(-2.f * x));
              const float dd = x - mu;
              lp += (0.f - ladj) + (-(dd * dd) / (2.f * (sd * sd)) - logf(sd) - 0.91893853320467267f);
            }
            polres[e.row] = make_float4(an[0], an[A - 1], lp, 0.f);
          }
          named_bar_sync(1, EPI_THREADS);
          const float4 pr = polres[e.row];
          float lp = pr.z, an[A];
          an[0] = pr.x; an[A - 1] = pr.y;
          // patch the sampled action into xs0's action columns: the next ops read [next_obs, a].  Order of the ops: actor,
          // target Q1, target Q2, actor_safe, target Qc - so the columns hold a1 while the Q's read them and a2 afterwards
          if (e.g == 0) {
#pragma unroll
            for (int j = 0; j < A; ++j) {
              const int k = S + j;
              *reinterpret_cast<__nv_bfloat16*>(xs0 + (k >> 3) * 2048 + e.row * 16 + (k & 7) * 2) = __float2bfloat16_rn(e.valid ? an[j] : 0.f);
            }
            fence_proxy_async();
          }
          if (p.dbg && e.g == 0 && e.valid) {
            float* dd = p.dbg + gr * 16;
            if (first) { dd[0] = an[0]; dd[1] = an[A - 1]; dd[2] = lp; } else { dd[3] = an[0]; dd[4] = an[A - 1]; }
          }
          if (first) {
            logp = lp;
#pragma unroll
            for (int j = 0; j < A; ++j) a1[j] = an[j];
          } else {
#pragma unroll
            for (int j = 0; j < A; ++j) a2[j] = an[j];
          }
#pragma unroll
          for (int j = 0; j < MAXO; ++j) hpart[j] = 0.f;
        } else if (d.post == POST_QT0 || d.post == POST_QT1) {
          float out[MAXO];
          head_combine(e, hpart, d.hb_off, out);
          if (d.post == POST_QT0) qt0 = out[0];
          else {
            // compute_target                                                              src/ssac.py:284-294
            q_target = rew + p.gamma * (1.f - dn) * (fminf(qt0, out[0]) - alpha * logp);
            if (p.dbg && e.g == 0 && e.valid) { p.dbg[gr * 16 + 5] = qt0; p.dbg[gr * 16 + 6] = out[0]; }
          }
          hpart[0] = 0.f;
        } else if (d.post == POST_KEEP) {
          // mean head of a constraint critic: its partial waits for the log-std head
#pragma unroll
          for (int j = 0; j < MAXO; ++j) { hkeep[j] = hpart[j]; hpart[j] = 0.f; }
          hb_keep = d.hb_off;
        } else if (d.post == POST_QCT) {
          // constraint_critic_target(next_obs, a2, sample=True)                             src/ssac.py:342-344, 88-90
          float om[MAXO], ol[MAXO];
          head_combine(e, hkeep, hb_keep, om);
          named_bar_sync(1, EPI_THREADS);                                 // hp is reused by the second combine
          head_combine(e, hpart, d.hb_off, ol);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float sd = expf(soft_clamp(ol[c], -4.f, 4.f));
            const float ee = fminf(fmaxf(e.valid ? p.n_qc.get(gr, c) : 0.f, -2.f), 2.f);
            nqc[c] = fmaf(ee, sd, om[c]);
          }
          if (p.dbg && e.g == 0 && e.valid) p.dbg[gr * 16 + 7] = nqc[0];
#pragma unroll
          for (int j = 0; j < MAXO; ++j) hpart[j] = 0.f;
        } else if (d.post == POST_Q0 || d.post == POST_Q1) {
          // Q_i loss and its backward through the head                                      src/ssac.py:296-298, 437-441
          const int i = d.post == POST_Q0 ? 0 : 1;
          tmem_st_wait();
          float oq[MAXO];
          head_combine(e, hpart, d.hb_off, oq);
          const float err_q = oq[0] - q_target;
          float dq[MAXO] = {e.valid ? err_q * p.inv_bg : 0.f, 0.f, 0.f, 0.f};
          if (e.g == 0) {
            add_loss(0, e.valid ? 0.5 * (double)err_q * err_q : 0.0);
            add_scal(i, dq[0]);
            if (p.dbg && e.valid) { p.dbg[gr * 16 + 8 + i] = oq[0]; p.dbg[gr * 16 + 12 + i] = dq[0]; }
          }
          epi_head_backward(e, TM_R1, dq, 1, p.hw_q[i], slot_q_w2(i), slot_q_b1(i), p.sv[SV_Q_DH2 + i]);
          hpart[0] = 0.f;
        } else if (d.post == POST_QC) {
          // distributional Qc loss (reachability targets with the TD bound) and its backward through both heads
          //                                                                                src/ssac.py:345-354, 416-423
          tmem_st_wait();
          float om[MAXO], ol[MAXO];
          head_combine(e, hkeep, hb_keep, om);
          named_bar_sync(1, EPI_THREADS);
          head_combine(e, hpart, d.hb_off, ol);
          float dmean[MAXO] = {0.f, 0.f, 0.f, 0.f}, dls[MAXO] = {0.f, 0.f, 0.f, 0.f};
          double lc = 0.0;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float h = cvr[c], mu = om[c];
            const float nonterm = p.one_minus_gamma * h + p.gamma * fmaxf(h, nqc[c]);
            const float tu = nonterm * (1.f - dn) + h * dn;
            const float tb = fminf(fmaxf(tu - mu, -p.td_bound), p.td_bound) + mu;
            const float x = ol[c];
            const float y1 = 4.f - softplus_f(4.f - x);
            const float ls = -4.f + softplus_f(y1 + 4.f);
            const float sd = expf(ls), var = sd * sd;
            const float du = mu - tu, db = mu - tb;
            if (e.valid) {
              dmean[c] = du / var * p.inv_bgc;
              dls[c] = (1.f - db * db / var) * p.inv_bgc * dsoftplus(y1 + 4.f) * dsoftplus(4.f - x);
              lc += (double)(du * du / (2.f * var) + db * db / (2.f * var) + logf(sd));
            }
          }
          if (e.g == 0) {
            add_loss(1, lc);
#pragma unroll
            for (int c = 0; c < C; ++c) { add_scal(2 + c, dmean[c]); add_scal(2 + C + c, dls[c]); }
          }
          if (p.dbg && e.g == 0 && e.valid) {
            float* dd = p.dbg + gr * 16;
            dd[10] = om[0]; dd[11] = ol[0]; dd[14] = dmean[0]; dd[15] = dls[0];
          }
#pragma unroll 1
          for (int hd = 0; hd < 2; ++hd) {                                 // mean head (m1 stashed in R0), log-std head (l1 in R1)
            float dsel[MAXO];
#pragma unroll
            for (int j = 0; j < MAXO; ++j) dsel[j] = hd == 0 ? dmean[j] : dls[j];
            epi_head_backward(e, hd == 0 ? TM_R0 : TM_R1, dsel, C, hd == 0 ? p.hw_cm : p.hw_cl, hd == 0 ? slot_c_wm(0) : slot_c_wl(0, C),
                              hd == 0 ? SLOT_C_BM0 : SLOT_C_BL0, p.sv[hd == 0 ? SV_C_DM1 : SV_C_DL1]);
          }
#pragma unroll
          for (int j = 0; j < MAXO; ++j) hpart[j] = 0.f;
        }
        if (oi != n_chain - 1) epi_op_done(e);          // the arrival for the next tile's first op follows its staging
      }
    }
    // ---- per-CTA results ---------------------------------------------------------------------------------------------------
    named_bar_sync(1, EPI_THREADS);
    if (threadIdx.x == 0) { p.loss_part[2 * blockIdx.x] = lacc[0]; p.loss_part[2 * blockIdx.x + 1] = lacc[1]; }
    for (int i = threadIdx.x; i < p.nv * HID; i += EPI_THREADS) p.gacc_out[(int64_t)blockIdx.x * p.nv * HID + i] = gacc[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA leaves while its peer may still multicast into it or signal its barriers
  if (warp == ISSUER) tmem_dealloc(tmem, 512);
}


// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static float* g_dbg_rows = nullptr;
static long long* g_prof = nullptr;

struct Plan {
  int64_t Bpad; int Kx, n_tiles, grid, nv;
  int64_t img_bytes; int ctab_floats;
  int dw_ctas; int64_t dw_partial_floats;
};
static Plan make_plan(int64_t B, int S, int A, int C) {
  Plan pl;
  pl.Bpad = (B + CLUSTER * TILE - 1) / (CLUSTER * TILE) * (CLUSTER * TILE);        // whole tile pairs (one per CTA pair)
  pl.Kx = round_up(S + A, 16);
  pl.n_tiles = (int)(pl.Bpad / TILE);
  pl.grid = std::min(pl.n_tiles, 148) / CLUSTER * CLUSTER;
  pl.nv = n_slots(C);
  // images: 8 first-layer (kp = Kx) + 17 hidden (kp = 256)
  pl.img_bytes = (int64_t)8 * HID * (pl.Kx + KBIAS) * 2 + (int64_t)12 * HID * (HID + KBIAS) * 2 + (int64_t)5 * HID * HID * 2;
  pl.ctab_floats = (4 * A + 4 + 4 * C) * HID + 64;
  return pl;
}
// split-K factors of the dW jobs: big jobs (N = 256) and first-layer jobs (N = Kx) share 148 CTAs in proportion to their bytes
static void dw_splits(const Plan& pl, int n_slabs, int& ks_big, int& ks_small) {
  const double wb = 64.0, ws = 32.0 + pl.Kx / 8.0;
  const double unit = 148.0 / (5 * wb + 3 * ws);
  ks_big = std::max(1, std::min(n_slabs, (int)(unit * wb)));
  ks_small = std::max(1, std::min(n_slabs, (int)((148 - 5 * ks_big) / 3)));
}

int64_t critic_ws_bytes(int64_t B, int S, int A, int C) {
  Plan pl = make_plan(B, S, A, C);
  int ksb, kss; dw_splits(pl, (int)(pl.Bpad / DW_ROWS), ksb, kss);
  int64_t b = 0;
  b += align_up(pl.img_bytes, 256) + align_up((int64_t)pl.ctab_floats * 4, 256);
  b += 12 * align_up(pl.Bpad * HID * 2, 256) + align_up(pl.Bpad * pl.Kx * 2, 256);
  b += align_up((int64_t)148 * pl.nv * HID * 4, 256) + align_up(148 * 2 * 8, 256);
  b += align_up(((int64_t)5 * ksb * HID * HID + (int64_t)3 * kss * HID * pl.Kx) * 4, 256);
  return b + 4096;
}

void critic_set_debug_rows(float* p) { g_dbg_rows = p; }
void critic_set_prof(long long* p) { g_prof = p; }

template <int A, int C>
static int launch_fused(const FusedParams& fp, int& grid, size_t smem, cudaStream_t st) {
  auto k = critic_fused_kernel<A, C>;
  static int max_clusters = 0;
  if (!max_clusters) {
    DRPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    // CTA pairs must sit in one GPC: ask how many pairs of this footprint can be resident (74 on a full B200)
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

// phase 1 of drpo_critic_step in DRPO_PREC_BF16: fills a.grads and a.losses[0..1]
int critic_phase1(const drpo_critic_args& a, int* err_flag) {
  const int64_t B = a.batch_size; const int S = a.state_dim, A = a.action_dim, C = a.con_dim, D = S + A;
  cudaStream_t st = (cudaStream_t)a.stream;
  DRPO_CHECK_ARG(a.q[0].l0.out_dim == HID && a.q[0].l1.out_dim == HID && a.actor->l0.out_dim == HID && a.actor->l1.out_dim == HID &&
                     a.qc.trunk0.out_dim == HID, "drpo_critic_step(bf16): the fused kernel needs hidden width 256");
  DRPO_CHECK_ARG(D <= 64, "drpo_critic_step(bf16): state_dim + action_dim must be <= 64 (got %d)", D);
  DRPO_CHECK_ARG((A == 1 || A == 2) && (C == 1 || C == 2 || C == 4), "drpo_critic_step(bf16): fused kernel is built for action_dim 1-2, con_dim 1/2/4");
  Plan pl = make_plan(B, S, A, C);
  const int n_slabs = (int)(pl.Bpad / DW_ROWS);
  int ksb, kss; dw_splits(pl, n_slabs, ksb, kss);
  Arena ar(a.workspace, a.workspace_bytes);
  uint8_t* img = ar.take<uint8_t>(pl.img_bytes);
  float* ctab = ar.take<float>(pl.ctab_floats);
  __nv_bfloat16* sv[12];
  for (int i = 0; i < 12; ++i) sv[i] = ar.take<__nv_bfloat16>(pl.Bpad * HID);
  __nv_bfloat16* x_sa = ar.take<__nv_bfloat16>(pl.Bpad * pl.Kx);
  float* gacc_out = ar.take<float>((int64_t)148 * pl.nv * HID);
  double* loss_part = ar.take<double>(148 * 2);
  float* dw_part = ar.take<float>((int64_t)5 * ksb * HID * HID + (int64_t)3 * kss * HID * pl.Kx);
  if (!ar.ok()) { set_error("drpo_critic_step(bf16): workspace too small (%lld needed, %lld given)", (long long)ar.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }

  // ---- plan: images, constant table, ops ---------------------------------------------------------------------------------
  FusedParams fp; memset(&fp, 0, sizeof(fp));
  PackTable pt; pt.n = 0; CopyTable ct; ct.n = 0;
  int64_t img_off = 0; int ctab_off = 0; int n_ops = 0;
  auto add_image = [&](const drpo_linear& l, bool transposed) -> uint32_t {
    PackEntry& e = pt.e[pt.n++];
    e.W = l.w; e.transposed = transposed ? 1 : 0; e.bias = transposed ? nullptr : l.b;
    if (!transposed) { e.n_real = l.out_dim; e.k_real = l.in_dim; e.kp = l.in_dim == HID ? HID : pl.Kx; }
    else { e.n_real = l.in_dim; e.k_real = l.out_dim; e.kp = HID; }
    e.dst = img_off / 2;
    const uint32_t off = (uint32_t)img_off;
    img_off += (int64_t)HID * (e.kp + (transposed ? 0 : KBIAS)) * 2;
    return off;
  };
  auto add_const = [&](const float* src, int n) -> int {
    CopyEntry& e = ct.e[ct.n++]; e.src = src; e.n = n; e.dst = ctab_off; e.stride = 1;
    const int off = ctab_off; ctab_off += (n + 3) & ~3; return off;
  };
  auto add_fwd = [&](const drpo_linear& l, int a_src, int out_region, int save, bool wait_all = false, bool early = false) {
    FOp& op = fp.op[n_ops];
    op.w_off[0] = add_image(l, false); op.kp = (uint16_t)(l.in_dim == HID ? HID : pl.Kx); op.a_src[0] = (uint8_t)a_src; op.parts = 1;
    op.bias = 1; op.early = early ? 1 : 0;
    EOp& e = fp.eop[n_ops];
    e.out_region = (uint8_t)out_region; e.save = (uint8_t)save; e.wait_all = wait_all ? 1 : 0;
    ++n_ops;
  };
  auto set_head = [&](const drpo_linear& head, int post) -> int {            // CUDA-core head on the op just added
    EOp& e = fp.eop[n_ops - 1];
    e.no = (uint8_t)head.out_dim; e.hw_off = add_const(head.w, head.out_dim * HID); e.hb_off = add_const(head.b, head.out_dim);
    e.post = (uint8_t)post;
    return e.hw_off;
  };
  auto add_bwd = [&](const drpo_linear& l, int a_src, int hsave, int bias_slot, int save, int out_region, bool wait_all) {
    FOp& op = fp.op[n_ops];
    op.w_off[0] = add_image(l, true); op.kp = HID; op.a_src[0] = (uint8_t)a_src; op.parts = 1;
    EOp& e = fp.eop[n_ops];
    e.backward = 1; e.hsave = (uint8_t)hsave; e.bias_slot = (uint8_t)bias_slot; e.save = (uint8_t)save; e.out_region = (uint8_t)out_region;
    e.wait_all = wait_all ? 1 : 0;
    ++n_ops;
  };
  // order: actor, target Q1, target Q2, actor_safe, target Qc (the sampled action in xs0 is a1, then a2), Q1, Q2, Qc
  add_fwd(a.actor->l0, A_XS0, 1, 0); add_fwd(a.actor->l1, A_R0, 0, 0); set_head(a.actor->l2, POST_POLICY0);
  for (int i = 0; i < 2; ++i) {
    add_fwd(a.q_target[i].l0, A_XS0, 1, 0, false, i == 1);      // (target Q1 waits for a1 to be patched into xs0; target Q2 need not)
    add_fwd(a.q_target[i].l1, A_R0, 0, 0); set_head(a.q_target[i].l2, i == 0 ? POST_QT0 : POST_QT1);
  }
  add_fwd(a.actor_safe->l0, A_XS0, 1, 0, false, true); add_fwd(a.actor_safe->l1, A_R0, 0, 0); set_head(a.actor_safe->l2, POST_POLICY1);
  add_fwd(a.qc_target.trunk0, A_XS0, 1, 0); add_fwd(a.qc_target.trunk1, A_R0, 2, 0);
  add_fwd(a.qc_target.mean0, A_R1, 0, 0); set_head(a.qc_target.mean1, POST_KEEP);
  add_fwd(a.qc_target.lstd0, A_R1, 0, 0); set_head(a.qc_target.lstd1, POST_QCT);
  for (int i = 0; i < 2; ++i) {
    add_fwd(a.q[i].l0, A_XS1, 1, SV_Q_H1 + i, false, true);
    add_fwd(a.q[i].l1, A_R0, 2, 0); fp.hw_q[i] = set_head(a.q[i].l2, i == 0 ? POST_Q0 : POST_Q1);           // h2 stashed in R1
    add_bwd(a.q[i].l1, A_R1, SV_Q_H1 + i, slot_q_b0(i), SV_Q_DH1 + i, 0, false);                              // dh1 = (dh2 W1) * (h1 > 0)
  }
  add_fwd(a.qc.trunk0, A_XS1, 1, SV_C_T1, false, true); add_fwd(a.qc.trunk1, A_R0, 2, SV_C_T2);
  add_fwd(a.qc.mean0, A_R1, 1, 0); fp.hw_cm = set_head(a.qc.mean1, POST_KEEP);                              // m1 stashed in R0
  add_fwd(a.qc.lstd0, A_R1, 2, 0, true); fp.hw_cl = set_head(a.qc.lstd1, POST_QC);                          // l1 stashed over t2
  {
    FOp& op = fp.op[n_ops];                                       // dt2 = (dm1 W_m0 + dl1 W_l0) * (t2 > 0): two K parts
    op.w_off[0] = add_image(a.qc.mean0, true); op.w_off[1] = add_image(a.qc.lstd0, true);
    op.kp = HID; op.a_src[0] = A_R0; op.a_src[1] = A_R1; op.parts = 2;
    EOp& e = fp.eop[n_ops];
    e.backward = 1; e.hsave = SV_C_T2; e.bias_slot = SLOT_C_BT1; e.save = SV_C_DT2; e.out_region = 1; e.wait_all = 1;
    ++n_ops;
  }
  add_bwd(a.qc.trunk1, A_R0, SV_C_T1, SLOT_C_BT0, SV_C_DT1, 0, false);                                        // dt1 = (dt2 W_t1) * (t1 > 0)
  if (n_ops != MAX_OPS || img_off > pl.img_bytes || ctab_off > pl.ctab_floats) {
    set_error("drpo_critic_step(bf16): internal plan mismatch (%d ops, %lld image bytes, %d consts)", n_ops, (long long)img_off, ctab_off);
    return DRPO_ERR_ARG;
  }
  {
    dim3 grid(64, pt.n + ct.n);
    DRPO_LAUNCH(pack_gather_kernel, grid, 256, 0, st, pt, reinterpret_cast<__nv_bfloat16*>(img), ct, ctab);
  }
  fp.n_ops = n_ops; fp.wimg = img; fp.ctab = ctab; fp.ctab_floats = ctab_off;
  {
    // The update is two chains that never exchange a value: {actor, target Q1, target Q2, Q1, Q2} (ops 0-5, 12-17) and
    // {safe actor, target Qc, Qc} (ops 6-11, 18-23), 12 ops each.  Dealing (tile pair, chain) jobs to the CTA pairs instead of
    // whole tile pairs halves the step of a small shard (8 192 rows: 32 tile pairs -> 64 jobs on 74 CTA pairs) and shortens the
    // tail of a large one (65 536 rows: 4 rounds of 24 ops -> 7 rounds of 12).  Taken whenever it lowers the number of rounds.
    const int n_pairs = pl.n_tiles / CLUSTER, mc = 148 / CLUSTER;
    bool split = (2 * n_pairs + mc - 1) / mc < 2 * ((n_pairs + mc - 1) / mc);
    if (const char* ev = getenv("DRPO_CRITIC_SPLIT")) split = ev[0] == '1';
    fp.sch.split = split ? 1 : 0;
    if (!split) {
      fp.sch.n[0] = (uint8_t)n_ops;
      for (int i = 0; i < n_ops; ++i) fp.sch.order[0][i] = (uint8_t)i;
    } else {
      fp.sch.n[0] = fp.sch.n[1] = 12;
      for (int i = 0; i < 6; ++i) {
        fp.sch.order[0][i] = (uint8_t)i; fp.sch.order[0][6 + i] = (uint8_t)(12 + i);
        fp.sch.order[1][i] = (uint8_t)(6 + i); fp.sch.order[1][6 + i] = (uint8_t)(18 + i);
      }
      pl.grid = std::min(2 * n_pairs, mc) * CLUSTER;
    }
  }
  const drpo_batch& b = a.batch;
  fp.obs = b.obs; fp.act = b.act; fp.next_obs = b.next_obs; fp.rew = b.rew; fp.cv = b.cv; fp.done = b.done;
  fp.n_actor = make_noise(a.eps_actor, A, a.seed, TAG_CRITIC_ACTOR, a.noise_step, a.row_id_offset);
  fp.n_safe = make_noise(a.eps_safe, A, a.seed, TAG_CRITIC_SAFE, a.noise_step, a.row_id_offset);
  fp.n_qc = make_noise(a.eps_qc, C, a.seed, TAG_CRITIC_QC, a.noise_step, a.row_id_offset);
  fp.log_alpha = a.log_alpha;
  fp.gamma = (float)a.discount; fp.one_minus_gamma = (float)(1.0 - a.discount); fp.td_bound = (float)a.qc_td_bound;
  fp.inv_bg = (float)(1.0 / (double)a.global_batch_size); fp.inv_bgc = (float)(1.0 / ((double)a.global_batch_size * C));
  fp.B = B; fp.Bpad = pl.Bpad; fp.S = S; fp.A = A; fp.C = C; fp.D = D; fp.Kx = pl.Kx; fp.n_tiles = pl.n_tiles;
  fp.x_sa = x_sa;
  fp.sv[0] = nullptr;
  for (int i = 0; i < 12; ++i) fp.sv[1 + i] = sv[i];
  fp.gacc_out = gacc_out; fp.nv = pl.nv; fp.loss_part = loss_part; fp.err_flag = err_flag; fp.dbg = g_dbg_rows; fp.prof = g_prof;
  // shared memory: ring + 2 x-buffers + constants + column sums + head partials + barriers
  const size_t fixed = (size_t)2 * TILE * pl.Kx * 2 + TILE * KBIAS * 2 + (size_t)((ctab_off + 3) & ~3) * 4 + (size_t)pl.nv * HID * 4 + (NGROUPS + 1) * TILE * 16 + sizeof(FusedSmem);
  int stages = (int)((232448 - 1024 - fixed) / CHUNK_BYTES);
  if (stages > 6) stages = 6;
  if (stages < 2) { set_error("drpo_critic_step(bf16): shared-memory budget exceeded"); return DRPO_ERR_ARG; }
  fp.stages = stages;
  const size_t smem = fixed + (size_t)stages * CHUNK_BYTES;
  int rc;
  // (pl.grid may shrink to the number of co-resident CTA pairs; the reductions below read it afterwards)
  if (A == 1 && C == 1) rc = launch_fused<1, 1>(fp, pl.grid, smem, st);
  else if (A == 1 && C == 2) rc = launch_fused<1, 2>(fp, pl.grid, smem, st);
  else if (A == 1 && C == 4) rc = launch_fused<1, 4>(fp, pl.grid, smem, st);
  else if (A == 2 && C == 1) rc = launch_fused<2, 1>(fp, pl.grid, smem, st);
  else if (A == 2 && C == 2) rc = launch_fused<2, 2>(fp, pl.grid, smem, st);
  else rc = launch_fused<2, 4>(fp, pl.grid, smem, st);
  if (rc) return rc;

  // ---- dW ------------------------------------------------------------------------------------------------------------------
  DwParams dp; memset(&dp, 0, sizeof(dp));
  dp.Bpad = pl.Bpad; dp.n_slabs = n_slabs; dp.err_flag = err_flag; dp.lbo = 128; dp.sbo = DW_PANEL;
  ReduceTable rt; rt.n = 0;
  float* G = a.grads; const float* P = a.params;
  auto goff = [&](const float* q) { return (int64_t)(q - P); };
  int cta = 0; float* part = dw_part;
  auto add_job = [&](const __nv_bfloat16* dH, const __nv_bfloat16* Hm, int in_dim, const drpo_linear& l) {
    const bool big = in_dim == HID;
    DwJob& j = dp.job[dp.n_jobs++];
    j.a = dH; j.b = Hm; j.b_octets = big ? 32 : pl.Kx / 8; j.cta0 = cta; j.ksplit = big ? ksb : kss; j.partial = part;
    const int N = j.b_octets * 8;
    ReduceEntry& r = rt.e[rt.n++];
    r.dst = goff(l.w); r.src = part; r.n_src = j.ksplit; r.stride = (int64_t)HID * N; r.rows = HID; r.cols = l.in_dim; r.ld = N;
    cta += j.ksplit; part += (int64_t)j.ksplit * HID * N;
  };
  auto add_vec = [&](const float* dst_param, int slot, int off, int n) {
    ReduceEntry& r = rt.e[rt.n++];
    r.dst = goff(dst_param); r.src = gacc_out + slot * HID + off; r.n_src = pl.grid; r.stride = (int64_t)pl.nv * HID; r.rows = 1; r.cols = n; r.ld = n;
  };
  for (int i = 0; i < 2; ++i) {
    add_job(fp.sv[SV_Q_DH2 + i], fp.sv[SV_Q_H1 + i], HID, a.q[i].l1);
    add_job(fp.sv[SV_Q_DH1 + i], x_sa, D, a.q[i].l0);
    add_vec(a.q[i].l2.w, slot_q_w2(i), 0, HID); add_vec(a.q[i].l2.b, SLOT_SCAL, i, 1);
    add_vec(a.q[i].l1.b, slot_q_b1(i), 0, HID); add_vec(a.q[i].l0.b, slot_q_b0(i), 0, HID);
  }
  add_job(fp.sv[SV_C_DM1], fp.sv[SV_C_T2], HID, a.qc.mean0); add_job(fp.sv[SV_C_DL1], fp.sv[SV_C_T2], HID, a.qc.lstd0);
  add_job(fp.sv[SV_C_DT2], fp.sv[SV_C_T1], HID, a.qc.trunk1); add_job(fp.sv[SV_C_DT1], x_sa, D, a.qc.trunk0);
  for (int c = 0; c < C; ++c) {
    add_vec(a.qc.mean1.w + c * HID, slot_c_wm(c), 0, HID); add_vec(a.qc.lstd1.w + c * HID, slot_c_wl(c, C), 0, HID);
  }
  add_vec(a.qc.mean1.b, SLOT_SCAL, 2, C); add_vec(a.qc.lstd1.b, SLOT_SCAL, 2 + C, C);
  add_vec(a.qc.mean0.b, SLOT_C_BM0, 0, HID); add_vec(a.qc.lstd0.b, SLOT_C_BL0, 0, HID);
  add_vec(a.qc.trunk1.b, SLOT_C_BT1, 0, HID); add_vec(a.qc.trunk0.b, SLOT_C_BT0, 0, HID);
  {
    static bool attr_done = false;
    const size_t dsm = DW_STAGES * DW_STAGE_BYTES + sizeof(DwSmem);
    if (!attr_done) { DRPO_CUDA_OK(cudaFuncSetAttribute(critic_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm)); attr_done = true; }
    DRPO_LAUNCH(critic_dw_kernel, cta, DW_THREADS, dsm, st, dp);
    dim3 grid(64, rt.n);
    DRPO_LAUNCH(critic_grad_reduce_kernel, grid, 256, 0, st, rt, G);
  }
  Scale2 sc; sc.v[0] = 1.0 / (double)a.global_batch_size; sc.v[1] = 1.0 / ((double)a.global_batch_size * C);
  DRPO_LAUNCH(loss_finalize2_kernel, 1, 32, 0, st, loss_part, pl.grid, sc, a.losses);
  return DRPO_OK;
}

// test aid: dW of one (dH, H) pair given in the octet layout -> out[256, 8*b_octets]
int critic_debug_dw(const void* a_oct, const void* b_oct, int b_octets, int64_t rows_padded, int ksplit, float* partial, float* out,
                    int* err_flag, void* stream) {
  DwParams dp; memset(&dp, 0, sizeof(dp));
  dp.Bpad = rows_padded; dp.n_slabs = (int)(rows_padded / DW_ROWS); dp.err_flag = err_flag; dp.n_jobs = 1;
  const bool swap = getenv("DRPO_DW_SWAP") != nullptr;          // experiment switch: LBO/SBO roles of the MN-major descriptor
  dp.lbo = swap ? DW_PANEL : 128; dp.sbo = swap ? 128 : DW_PANEL;
  dp.job[0].a = (const __nv_bfloat16*)a_oct; dp.job[0].b = (const __nv_bfloat16*)b_oct; dp.job[0].b_octets = b_octets;
  dp.job[0].cta0 = 0; dp.job[0].ksplit = ksplit; dp.job[0].partial = partial;
  const size_t dsm = DW_STAGES * DW_STAGE_BYTES + sizeof(DwSmem);
  DRPO_CUDA_OK(cudaFuncSetAttribute(critic_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
  DRPO_LAUNCH(critic_dw_kernel, ksplit, DW_THREADS, dsm, (cudaStream_t)stream, dp);
  ReduceTable rt; rt.n = 1;
  const int N = b_octets * 8;
  rt.e[0].dst = 0; rt.e[0].src = partial; rt.e[0].n_src = ksplit; rt.e[0].stride = (int64_t)HID * N; rt.e[0].rows = HID; rt.e[0].cols = N; rt.e[0].ld = N;
  dim3 grid(32, 1);
  DRPO_LAUNCH(critic_grad_reduce_kernel, grid, 256, 0, (cudaStream_t)stream, rt, out);
  return DRPO_OK;
}

}  // namespace cu
}  // namespace drpo
