// ls2d_general.cuh -- icp_general_kernel: the aligner with EVERY option, for clouds of any size.
//
// Same projection / z-buffer / finder / factor code as icp_stream_kernel (ls2d_icp.cuh: per-point column and rho
// stashed in shared memory, xor-butterfly reduction, thread t owns points t, t + T, ...), plus what the shipped
// configurations leave switched off and BASELINE.json's north_star names:
//   * Levenberg-Marquardt rounds (IterationAlgorithmLM; oracle decisions L1..L8): thread 0 proposes a damped step,
//     the CTA evaluates the robustified chi2 of the round's correspondences at the trial pose (a moving point that
//     made a correspondence carries a flag in bit 31 of its stashed rho), thread 0 takes the gain ratio and accepts
//     or rejects; lambda is carried across the rounds of one alignment.
//   * MultiAligner2D.enable_inlier_only_runs (I1): a second phase of rounds in which kernelized factors weigh 0.
//   * a termination criterion on the relative chi2 decay between rounds (T1).
// The point-to-point factor (D19) needs none of this: linearize_point() carries it in every run-time-shaped kernel.
// Reference: L0.json:9-37 (MultiAligner2D parameters), :83-88,193-215 (solver / algorithm).
#pragma once

#include "ls2d_icp.cuh"

namespace ls2d {

constexpr int STOP_PHASE_DONE = 100;  // bc->stop: the termination criterion ended the phase (not a failure)

// state of one alignment that outlives a barrier (thread 0 writes, everybody reads after __syncthreads)
struct general_shared {
  float v[NSUM];         // totals of the round's linearisation
  int n_in, n_k;
  iso X;                 // accepted estimate (bc holds the pose the CTA currently evaluates: X or an LM trial)
  iso Xt;                // LM trial estimate
  float dx[3];
  double D[3];           // lambda * D_j of the trial
  double lambda, nu;
  int lm_started, lm_trial, lm_pending, lm_done, lm_rejected;
  float chi_prev;
  int n_corr_last;
  int phase_done;        // the termination criterion ended the phase: written by thread 0 between the round's last two
                         // barriers and read after the last one (bc->stop is written before them and read between them)
};

constexpr size_t icp_general_smem_bytes(int cols, int threads, int max_points) {
  return icp_stream_smem_bytes(cols, threads, max_points, false) + (size_t)(threads / 32) * 4 + sizeof(general_shared) + 32;
}

template <int T, bool SENSOR>
__global__ void __launch_bounds__(T, 2) icp_general_kernel(const dev_params P, const align_args A, int max_points) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C          = P.cam.cols;
  float4* fimg         = reinterpret_cast<float4*>(smem_raw);
  float* fdepth        = reinterpret_cast<float*>(fimg + C);
  unsigned* zdepth     = reinterpret_cast<unsigned*>(fdepth + C);
  unsigned* zidx       = zdepth + C;
  unsigned* srho       = zidx + C;                                              // [max_points] rho bits | active << 31
  unsigned short* scol = reinterpret_cast<unsigned short*>(srho + max_points);  // [max_points] column, 0xFFFF = none
  float* red           = reinterpret_cast<float*>(scol + 2 * ((max_points + 1) / 2));
  float* red1          = red + (T / 32) * RED_STRIDE;                           // [T / 32] chi2 partials of an LM trial
  pose_bc* bc          = reinterpret_cast<pose_bc*>(red1 + T / 32);
  general_shared* gs   = reinterpret_cast<general_shared*>((reinterpret_cast<uintptr_t>(bc + 1) + 15) & ~uintptr_t(15));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = blockIdx.x + A.pair_base;
  const int fcl  = A.fixed_const >= 0 ? A.fixed_const : (A.fixed_id ? A.fixed_id[pair] : pair);
  const int mcl  = A.moving_id ? A.moving_id[pair / A.moving_div] : pair / A.moving_div;
  const int f0 = A.fixed_off[fcl], nf = A.fixed_off[fcl + 1] - f0;
  const int m0 = A.moving_off[mcl], nm = A.moving_off[mcl + 1] - m0;
  const float4* fpts = A.fixed_pts + f0;
  const float4* mpts = A.moving_pts + m0;

  for (int k = tid; k < C; k += T) {
    fdepth[k] = -1.f;
    zdepth[k] = Z_EMPTY_DEPTH;
    zidx[k]   = Z_EMPTY_IDX;
  }
  if (tid == 0) {
    gs->X = load_pose(A.init_pose, (size_t) pair, A.pose_stride);
    publish_pose(bc, P, gs->X, SENSOR, 0);
    bc->tie = 0;
    for (int k = 0; k < NSUM; ++k) gs->v[k] = 0.f;
    gs->n_in = gs->n_k = 0;
    gs->phase_done = 0;
    gs->lm_started = 0, gs->lm_rejected = 0;
    gs->lambda = 0.0;
    gs->n_corr_last = 0;
  }
  __syncthreads();

  // ---- fixed range image (identity camera), exact two-pass z-buffer
  for (int i = tid; i < nf; i += T) {
    const float4 p  = ldg4(fpts + i);
    const float rho = fsqrt(fadd(fmul(p.x, p.x), fmul(p.y, p.y)));
    int col         = -1;
    if (!(rho < P.range_min || rho > P.range_max)) col = polar_column(P.cam, p.y, p.x);
    scol[i] = (unsigned short) (col < 0 ? 0xFFFF : col);
    srho[i] = f2u(rho);
    if (col >= 0) atomicMin(&zdepth[col], f2u(rho));
  }
  __syncthreads();
  for (int i = tid; i < nf; i += T) {
    const unsigned c = scol[i];
    if (c != 0xFFFF && zdepth[c] == srho[i]) atomicMin(&zidx[c], (unsigned) i);
  }
  __syncthreads();
  for (int i = tid; i < nf; i += T) {
    const unsigned c = scol[i];
    if (c != 0xFFFF && zidx[c] == (unsigned) i) {
      fimg[c]   = ldg4(fpts + i);
      fdepth[c] = u2f(srho[i]);
    }
  }
  __syncthreads();
  for (int k = tid; k < C; k += T) zdepth[k] = Z_EMPTY_DEPTH, zidx[k] = Z_EMPTY_IDX;
  __syncthreads();

  const bool lm      = P.algorithm == LS2D_ALGORITHM_LM;
  const int n_phases = (P.inlier_only_runs && !A.score_only) ? 2 : 1;
  const int max_it   = A.score_only ? 1 : P.max_iterations;
  int it             = 0;  // rounds executed over both phases
  int status         = -1;
  for (int phase = 0; phase < n_phases && status < 0; ++phase) {
    if (phase == 1 && gs->n_in < P.min_num_inliers) break;  // I1 (uniform: gs is stable here)
    const bool inlier_only = phase == 1;
    bool exact = false;
    for (int k = 0; k < max_it; ++k) {
      if (exact) __syncthreads();
      project_and_stash<T, 4>(P, bc, nm, [&](int i) { return ldg4(mpts + i); }, scol, srho, zdepth);
      __syncthreads();
      if (exact) {
        for (int i = tid; i < nm; i += T) {
          const unsigned c = scol[i];
          if (c != 0xFFFF && zdepth[c] == srho[i]) atomicMin(&zidx[c], (unsigned) i);
        }
        __syncthreads();
      }
      const float Xtx = bc->Xtx, Xty = bc->Xty, Lc = bc->Lc, Ls = bc->Ls;
      float acc[16];
#pragma unroll
      for (int s = 0; s < 16; ++s) acc[s] = 0.f;
      unsigned cnt = 0;
      for (int i = tid; i < nm; i += T) {  // ascending i within a thread: the reduction shape the oracle mirrors
        const unsigned c = scol[i];
        if (c == 0xFFFF || zdepth[c] != srho[i]) continue;
        if (exact) {
          if (zidx[c] != (unsigned) i) continue;
        } else if (atomicCAS(&zidx[c], Z_EMPTY_IDX, (unsigned) i) != Z_EMPTY_IDX) {
          bc->tie = 1;
          continue;
        }
        const float4 M = ldg4(mpts + i);
        if (linearize_point<SENSOR, true>(P, bc, fdepth[c], fimg[c], M, u2f(srho[i]), Xtx, Xty, Lc, Ls, acc, cnt,
                                          inlier_only))
          srho[i] |= 0x80000000u;  // a correspondence: the LM trials re-evaluate it (only its owner reads srho again)
      }
      store_partials(acc, cnt, red, lane, warp);
      __syncthreads();
      for (int kk = tid; kk < C; kk += T) zdepth[kk] = Z_EMPTY_DEPTH, zidx[kk] = Z_EMPTY_IDX;
      if (!exact && bc->tie) {
        __syncthreads();
        if (tid == 0) bc->tie = 0;
        exact = true;
        --k;
        continue;
      }
      exact = false;

      // ---- totals, gate, Gauss-Newton step or first LM trial: warp 0
      if (warp == 0) {
        float tot        = 0.f;
        unsigned tot_cnt = 0;
        if (lane < NSUM) {
          tot = red[lane];
#pragma unroll
          for (int w = 1; w < T / 32; ++w) tot = fadd(tot, red[w * RED_STRIDE + lane]);
        } else if (lane == NSUM) {
#pragma unroll
          for (int w = 0; w < T / 32; ++w) tot_cnt += __float_as_uint(red[w * RED_STRIDE + NSUM]);
        }
        float v[NSUM];
#pragma unroll
        for (int s = 0; s < NSUM; ++s) v[s] = __shfl_sync(0xffffffffu, tot, s);
        const unsigned c2 = __shfl_sync(0xffffffffu, tot_cnt, NSUM);
        if (lane == 0) {
          const int n_in = c2 & 0xffff, n_k = c2 >> 16, n_corr = n_in + n_k;
#pragma unroll
          for (int s = 0; s < NSUM; ++s) gs->v[s] = v[s];
          gs->n_in = n_in, gs->n_k = n_k, gs->n_corr_last = n_corr;
          int stop      = 0;
          gs->lm_done   = 1;
          if (n_corr <= P.min_num_correspondences) {
            stop = 1 + LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES;
          } else if (!A.score_only) {
            if (!lm) {
              float dx[3];
              if (!solve3(v, P.damping, dx))
                stop = 1 + LS2D_STATUS_SINGULAR;
              else
                gs->X = iso_compose(gs->X, iso_v2t(dx[0], dx[1], dx[2]));
            } else {
              if (!gs->lm_started) {  // L2
                double mx = (double) v[0] > (double) v[3] ? (double) v[0] : (double) v[3];
                mx        = mx > (double) v[5] ? mx : (double) v[5];
                gs->lambda = P.lm_user_lambda_init > 0.f ? (double) P.lm_user_lambda_init : dmul((double) P.lm_tau, mx);
                gs->lm_started = 1;
              }
              gs->nu = 2.0, gs->lm_trial = 0, gs->lm_pending = 0, gs->lm_done = 0;
            }
          }
          bc->stop = stop;
        }
      }
      __syncthreads();

      // ---- Levenberg-Marquardt trials (L3..L7); uniform control flow through gs->lm_done
      if (lm && !A.score_only && bc->stop == 0) {
        for (;;) {
          if (tid == 0) {
            const float* v = gs->v;
            if (gs->lm_pending) {  // decide the evaluated trial (L5, L6)
              float chi1 = red1[0];
#pragma unroll
              for (int w = 1; w < T / 32; ++w) chi1 = fadd(chi1, red1[w]);
              const float chi0 = fadd(v[9], v[10]);
              double scale     = 0.0;
#pragma unroll
              for (int j = 0; j < 3; ++j)
                scale = dadd(scale, dmul((double) gs->dx[j], dsub(dmul(gs->D[j], (double) gs->dx[j]), (double) v[6 + j])));
              scale            = dadd(scale, 1e-3);
              const double rho = ddiv(dsub((double) chi0, (double) chi1), scale);
              if (rho > 0.0 && isfinite(chi1)) {
                const double q = dsub(dmul(2.0, rho), 1.0);
                double alpha   = dsub(1.0, dmul(dmul(q, q), q));
                alpha          = alpha < (double) P.lm_step_high ? alpha : (double) P.lm_step_high;
                const double g = alpha > (double) P.lm_step_low ? alpha : (double) P.lm_step_low;
                gs->lambda     = dmul(gs->lambda, g);
                gs->X          = gs->Xt;
                gs->lm_done    = 1;
              } else {
                gs->lambda = dmul(gs->lambda, gs->nu);
                gs->nu     = dmul(gs->nu, 2.0);
                gs->lm_rejected++;
                gs->lm_trial++;
              }
              gs->lm_pending = 0;
            }
            while (!gs->lm_done && gs->lm_trial < P.lm_iterations_max) {  // next solvable trial (L3)
              const double diag[3] = {(double) v[0], (double) v[3], (double) v[5]};
#pragma unroll
              for (int j = 0; j < 3; ++j) gs->D[j] = P.lm_variable_damping ? dmul(gs->lambda, diag[j]) : gs->lambda;
              float dx[3];
              if (solve3d(v, gs->D[0], gs->D[1], gs->D[2], dx)) {
                gs->dx[0] = dx[0], gs->dx[1] = dx[1], gs->dx[2] = dx[2];
                gs->Xt    = iso_compose(gs->X, iso_v2t(dx[0], dx[1], dx[2]));
                publish_trial_pose(bc, P, gs->Xt, SENSOR);  // bc->stop is 0 and stays
                gs->lm_pending = 1;
                break;
              }
              gs->lambda = dmul(gs->lambda, gs->nu);
              gs->nu     = dmul(gs->nu, 2.0);
              gs->lm_rejected++;
              gs->lm_trial++;
            }
            if (!gs->lm_pending) gs->lm_done = 1;  // accepted, or out of trials: X stays (L7)
          }
          __syncthreads();
          if (gs->lm_done) break;
          // L4: robustified chi2 of the round's correspondences at the trial pose
          float chi = 0.f;
          for (int i = tid; i < nm; i += T) {
            if (!(srho[i] >> 31)) continue;
            const unsigned c = scol[i];
            chi = fadd(chi, correspondence_chi<SENSOR>(P, bc, fimg[c], ldg4(mpts + i)));
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) chi = fadd(chi, __shfl_xor_sync(0xffffffffu, chi, off));
          if (lane == 0) red1[warp] = chi;
          __syncthreads();
        }
      }

      // ---- record the round, termination criterion, next pose
      if (tid == 0 && bc->stop == 0 && !A.score_only) {
        const iso X = gs->X;
        if (A.iters) {
          ls2d_iter_stats st;
          st.x = X.tx, st.y = X.ty, st.theta = atan2f_fdlibm(X.s, X.c);
          st.chi_inliers = gs->v[9], st.chi_kernelized = gs->v[10];
          st.n_inliers = gs->n_in, st.n_kernelized = gs->n_k, st.n_corr = gs->n_in + gs->n_k;
          st.c = X.c, st.s = X.s;
          A.iters[(size_t) pair * A.iters_stride + it] = st;
        }
        int stop        = 0;
        const float chi = fadd(gs->v[9], gs->v[10]);
        if (P.termination_epsilon > 0.f && k > 0 && fsub(gs->chi_prev, chi) < fmul(P.termination_epsilon, gs->chi_prev))
          stop = STOP_PHASE_DONE;  // T1
        gs->chi_prev = chi;
        publish_trial_pose(bc, P, X, SENSOR);  // bc->stop stays 0: other threads may still be reading it
        gs->phase_done = stop;
      }
      __syncthreads();
      const int stop = bc->stop ? bc->stop : gs->phase_done;
      if (stop && stop != STOP_PHASE_DONE) {
        status = stop - 1;
        break;
      }
      ++it;
      if (stop == STOP_PHASE_DONE) break;
    }
    __syncthreads();  // everybody has read bc->stop before thread 0 of the next phase clears it
    if (tid == 0) bc->stop = 0, gs->phase_done = 0;
    __syncthreads();
  }

  if (tid == 0) {
    int n_in = gs->n_in, n_k = gs->n_k;
    const int n_corr = gs->n_corr_last;
    float v[NSUM];
#pragma unroll
    for (int s = 0; s < NSUM; ++s) v[s] = gs->v[s];
    if (status == LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES) {  // the oracle reports empty sums here
#pragma unroll
      for (int s = 0; s < NSUM; ++s) v[s] = 0.f;
      n_in = n_k = 0;
    }
    if (status < 0) status = n_in < P.min_num_inliers ? LS2D_STATUS_NOT_ENOUGH_INLIERS : LS2D_STATUS_SUCCESS;
    ls2d_result r;
    r.x = gs->X.tx, r.y = gs->X.ty, r.theta = atan2f_fdlibm(gs->X.s, gs->X.c);
    r.chi_inliers = v[9], r.chi_kernelized = v[10];
    r.n_inliers = n_in, r.n_kernelized = n_k, r.n_corr = n_corr;
    r.status = status, r.iterations = it;
#pragma unroll
    for (int s = 0; s < 6; ++s) r.H[s] = v[s];
    r.c = gs->X.c, r.s = gs->X.s;
    r.lm_rejected = gs->lm_rejected, r.reserved = 0;
    A.out[pair] = r;
  }
}

}  // namespace ls2d
