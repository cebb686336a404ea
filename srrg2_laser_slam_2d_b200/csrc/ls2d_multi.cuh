// ls2d_multi.cuh -- the multi-slice aligner: MultiAligner2D with several laser slice processors and an optional
// odometry prior fused into ONE 3x3 system per iteration (MULTI.json:700-730: al_sl_laser_0 + ad_sl_odom +
// al_sl_laser_1; LASER_0.json:502-506: laser + odom).  One CTA per pair.  Every slice owns a fixed range image in
// shared memory (built once), its projector / finder / robustifier parameters and its sensor_in_robot; the
// z-buffer and the per-point stash are shared by the slices, which run one after the other inside an iteration.
// Reduction shape per slice = icp_stream_kernel's (thread t owns points t, t+T, ...), the slices' totals are then
// added in slice order, the prior last (oracle decisions D14-D17).
#pragma once

#include "ls2d_icp.cuh"

namespace ls2d {

constexpr int MULTI_THREADS = 256;  // threads per pair: the shape of the per-slice reduction tree



struct multi_shared {
  pose_bc bc[MAX_SLICES];
  iso Zinv;
  float tot[NSUM];
  int n_in, n_k, n_corr;
};

__host__ __device__ inline size_t multi_smem_bytes(const int* cols, int n_slices, int max_cols, int max_points,
                                                  int threads) {
  size_t b = 0;
  for (int s = 0; s < n_slices; ++s) b += (size_t) cols[s] * (16 + 4);
  b += (size_t) max_cols * 8;
  b += (size_t) max_points * 4 + (size_t)((max_points + 1) / 2) * 4;
  b += (size_t) MAX_SLICES * (threads / 32) * RED_STRIDE * 4;
  b += sizeof(multi_shared) + 64;
  return b;
}

// SE2PriorErrorFactor (AlignerSliceOdom2DPrior; L0.json:291-310): e = t2v(Z^-1 X), J = blockdiag(R(Z^-1 X), 1);
// operation order of oracle/ls2d_oracle.c prior_contribution() (decision D15).  Returns true if inlier.
__device__ __forceinline__ bool prior_contribution(const multi_args& A, const iso& Zinv, const iso& X, float* v) {
  const iso P      = iso_compose(Zinv, X);
  const float e[3] = {P.tx, P.ty, atan2f_fdlibm(P.s, P.c)};
  const float* O   = A.prior_info;
  const float Om[3][3] = {{O[0], O[1], O[2]}, {O[1], O[3], O[4]}, {O[2], O[4], O[5]}};
  float Oe[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) Oe[i] = fadd(fadd(fmul(Om[i][0], e[0]), fmul(Om[i][1], e[1])), fmul(Om[i][2], e[2]));
  const float chi = fadd(fadd(fmul(e[0], Oe[0]), fmul(e[1], Oe[1])), fmul(e[2], Oe[2]));
  float w = 1.f, chi_in = chi, chi_k = 0.f;
  bool inlier = true;
  if (A.prior_tau > 0.f && !(chi < A.prior_tau)) {
    const float aux = fadd(fmul(chi, A.prior_inv_tau), 1.f);
    chi_k           = fmul(A.prior_tau, __logf(aux));
    w               = frcp(aux);
    chi_in          = 0.f;
    inlier          = false;
  }
  float Aw[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Aw[0][j] = fmul(fadd(fmul(P.c, Om[0][j]), fmul(P.s, Om[1][j])), w);
    Aw[1][j] = fmul(fadd(fmul(-P.s, Om[0][j]), fmul(P.c, Om[1][j])), w);
    Aw[2][j] = fmul(Om[2][j], w);
  }
  float H[3][3], b[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    H[i][0] = fadd(fmul(Aw[i][0], P.c), fmul(Aw[i][1], P.s));
    H[i][1] = fadd(fmul(Aw[i][0], -P.s), fmul(Aw[i][1], P.c));
    H[i][2] = Aw[i][2];
    b[i]    = fadd(fadd(fmul(Aw[i][0], e[0]), fmul(Aw[i][1], e[1])), fmul(Aw[i][2], e[2]));
  }
  v[0] = H[0][0], v[1] = H[0][1], v[2] = H[0][2], v[3] = H[1][1], v[4] = H[1][2], v[5] = H[2][2];
  v[6] = b[0], v[7] = b[1], v[8] = b[2];
  v[9]  = chi_in;
  v[10] = chi_k;
  return inlier;
}

template <int T, int MINB>
__global__ void __launch_bounds__(T, MINB) icp_multi_kernel(const multi_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int S = A.n_slices;
  // ---- shared-memory carve-up: fixed images (float4 part, then depth part), z-buffer, stash, partials, poses
  float4* fimg[MAX_SLICES];
  float* fdepth[MAX_SLICES];
  {
    float4* p = reinterpret_cast<float4*>(smem_raw);
#pragma unroll
    for (int s = 0; s < MAX_SLICES; ++s) {
      fimg[s] = p;
      if (s < S) p += A.sl[s].P.cam.cols;
    }
    float* q = reinterpret_cast<float*>(p);
#pragma unroll
    for (int s = 0; s < MAX_SLICES; ++s) {
      fdepth[s] = q;
      if (s < S) q += A.sl[s].P.cam.cols;
    }
  }
  int img_cols = 0;
  for (int s = 0; s < S; ++s) img_cols += A.sl[s].P.cam.cols;
  unsigned* zdepth     = reinterpret_cast<unsigned*>(smem_raw + (size_t) img_cols * 20);
  unsigned* zidx       = zdepth + A.max_cols;
  unsigned* srho       = zidx + A.max_cols;
  unsigned short* scol = reinterpret_cast<unsigned short*>(srho + A.max_points);
  float* red           = reinterpret_cast<float*>(scol + 2 * ((A.max_points + 1) / 2));  // [MAX_SLICES][T/32][RED_STRIDE]
  multi_shared* sh     = reinterpret_cast<multi_shared*>(
      (reinterpret_cast<uintptr_t>(red + MAX_SLICES * (T / 32) * RED_STRIDE) + 15) & ~uintptr_t(15));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = blockIdx.x;
  const int fcl  = A.fixed_id ? A.fixed_id[pair] : pair;
  const int mcl  = A.moving_id ? A.moving_id[pair] : pair;

  for (int k = tid; k < img_cols; k += T) fdepth[0][k] = -1.f;  // the depth parts are contiguous
  for (int k = tid; k < A.max_cols; k += T) {
    zdepth[k] = Z_EMPTY_DEPTH;
    zidx[k]   = Z_EMPTY_IDX;
  }
  if (tid == 0) {
    const iso X = load_pose(A.init_pose, (size_t) pair, A.pose_stride);
    for (int s = 0; s < S; ++s) {
      publish_pose(&sh->bc[s], A.sl[s].P, X, A.sl[s].P.with_sensor != 0, 0);
      sh->bc[s].tie = 0;
    }
    if (A.prior_z) sh->Zinv = iso_inverse(load_pose(A.prior_z, (size_t) pair, A.pose_stride));
    for (int k = 0; k < NSUM; ++k) sh->tot[k] = 0.f;
    sh->n_in = sh->n_k = sh->n_corr = 0;
  }
  __syncthreads();

  // ---- fixed range images, identity camera (R/registration/correspondence_finder_projective_2d.cpp:37-44)
  for (int s = 0; s < S; ++s) {
    const dev_params& P = A.sl[s].P;
    const int f0 = A.sl[s].fixed_off[fcl], nf = A.sl[s].fixed_off[fcl + 1] - f0;
    const float4* fpts = A.sl[s].fixed_pts + f0;
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
        fimg[s][c]   = ldg4(fpts + i);
        fdepth[s][c] = u2f(srho[i]);
      }
    }
    __syncthreads();
    for (int i = tid; i < nf; i += T) {
      const unsigned c = scol[i];
      if (c != 0xFFFF) zdepth[c] = Z_EMPTY_DEPTH, zidx[c] = Z_EMPTY_IDX;
    }
    __syncthreads();
  }

  const int max_it = A.score_only ? 1 : A.sl[0].P.max_iterations;  // aligner-level (D17)
  int it           = 0;
  int status       = -1;
  for (; it < max_it; ++it) {
    for (int s = 0; s < S; ++s) {
      const dev_params& P = A.sl[s].P;
      const pose_bc* bc   = &sh->bc[s];
      const int m0 = A.sl[s].moving_off[mcl], nm = A.sl[s].moving_off[mcl + 1] - m0;
      const float4* mpts = A.sl[s].moving_pts + m0;
      const float Xtx = bc->Xtx, Xty = bc->Xty, Lc = bc->Lc, Ls = bc->Ls;
      bool exact = false;
      for (;;) {  // optimistic z-buffer pass, redone exactly on a tie (see icp_fused_kernel)
        project_and_stash<T, 1>(P, bc, nm, [&](int i) { return ldg4(mpts + i); }, scol, srho, zdepth);
        __syncthreads();
        if (exact) {
          for (int i = tid; i < nm; i += T) {
            const unsigned c = scol[i];
            if (c != 0xFFFF && zdepth[c] == srho[i]) atomicMin(&zidx[c], (unsigned) i);
          }
          __syncthreads();
        }
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.f;
        unsigned cnt = 0;
        for (int i = tid; i < nm; i += T) {
          const unsigned c = scol[i];
          if (c == 0xFFFF || zdepth[c] != srho[i]) continue;
          if (exact) {
            if (zidx[c] != (unsigned) i) continue;
          } else if (atomicCAS(&zidx[c], Z_EMPTY_IDX, (unsigned) i) != Z_EMPTY_IDX) {
            sh->bc[0].tie = 1;
            continue;
          }
          const float4 M = ldg4(mpts + i);
          if (P.with_sensor)
            linearize_point<true>(P, bc, fdepth[s][c], fimg[s][c], M, u2f(srho[i]), Xtx, Xty, Lc, Ls, acc, cnt);
          else
            linearize_point<false>(P, bc, fdepth[s][c], fimg[s][c], M, u2f(srho[i]), Xtx, Xty, Lc, Ls, acc, cnt);
        }
        store_partials(acc, cnt, red + s * (T / 32) * RED_STRIDE, lane, warp);
        __syncthreads();
        for (int i = tid; i < nm; i += T) {
          const unsigned c = scol[i];
          if (c != 0xFFFF) zdepth[c] = Z_EMPTY_DEPTH, zidx[c] = Z_EMPTY_IDX;
        }
        const bool redo = !exact && sh->bc[0].tie;
        __syncthreads();  // cells handed back (and the tie flag read) before the next pass touches them
        if (!redo) break;
        if (tid == 0) sh->bc[0].tie = 0;
        exact = true;
      }
    }

    // ---- totals, gates, prior, Gauss-Newton step: warp 0
    if (warp == 0) {
      float v[NSUM];
      int n_in = 0, n_k = 0, n_corr = 0, n_found = 0;
      bool contributed = false;
      for (int s = 0; s < S; ++s) {
        const float* r = red + s * (T / 32) * RED_STRIDE;
        float t        = 0.f;
        unsigned c     = 0;
        if (lane < NSUM) {
          t = r[lane];
#pragma unroll
          for (int w = 1; w < T / 32; ++w) t = fadd(t, r[w * RED_STRIDE + lane]);
        } else if (lane == NSUM) {
#pragma unroll
          for (int w = 0; w < T / 32; ++w) c += __float_as_uint(r[w * RED_STRIDE + NSUM]);
        }
        const unsigned c2 = __shfl_sync(0xffffffffu, c, NSUM);
        const int s_in = c2 & 0xffff, s_k = c2 >> 16;
        n_found += s_in + s_k;
        const bool use = s_in + s_k > A.sl[s].P.min_num_correspondences;  // D14 (uniform over the warp)
#pragma unroll
        for (int k = 0; k < NSUM; ++k) {
          const float x = __shfl_sync(0xffffffffu, t, k);
          if (use) v[k] = contributed ? fadd(v[k], x) : x;                // D16
        }
        if (use) n_in += s_in, n_k += s_k, n_corr += s_in + s_k, contributed = true;
      }
      if (lane == 0) {
        iso X;
        X.tx = sh->bc[0].Xtx, X.ty = sh->bc[0].Xty, X.c = sh->bc[0].Xc, X.s = sh->bc[0].Xs;
        int stop = 0;
        if (!contributed) {
          stop = 1 + LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES;
#pragma unroll
          for (int k = 0; k < NSUM; ++k) v[k] = 0.f;
          n_in = n_k = 0;
          n_corr = n_found;
        } else {
          if (A.prior_z) {
            float pv[NSUM];
            const bool inl = prior_contribution(A, sh->Zinv, X, pv);
#pragma unroll
            for (int k = 0; k < NSUM; ++k) v[k] = fadd(v[k], pv[k]);
            n_in += inl ? 1 : 0;
            n_k += inl ? 0 : 1;
          }
          if (!A.score_only) {
            float dx[3];
            if (!solve3(v, A.sl[0].P.damping, dx)) {
              stop = 1 + LS2D_STATUS_SINGULAR;
            } else {
              X = iso_compose(X, iso_v2t(dx[0], dx[1], dx[2]));
              if (A.iters) {
                ls2d_iter_stats st;
                st.x = X.tx, st.y = X.ty, st.theta = atan2f_fdlibm(X.s, X.c);
                st.chi_inliers = v[9], st.chi_kernelized = v[10];
                st.n_inliers = n_in, st.n_kernelized = n_k, st.n_corr = n_corr;
                st.c = X.c, st.s = X.s;
                A.iters[(size_t) pair * A.sl[0].P.max_iterations + it] = st;
              }
            }
          }
        }
#pragma unroll
        for (int k = 0; k < NSUM; ++k) sh->tot[k] = v[k];
        sh->n_in = n_in, sh->n_k = n_k, sh->n_corr = n_corr;
        for (int s = 0; s < S; ++s) publish_pose(&sh->bc[s], A.sl[s].P, X, A.sl[s].P.with_sensor != 0, stop);
      }
    }
    __syncthreads();
    if (sh->bc[0].stop) {
      status = sh->bc[0].stop - 1;
      break;
    }
  }

  if (tid == 0) {
    if (status < 0) status = sh->n_in < A.sl[0].P.min_num_inliers ? LS2D_STATUS_NOT_ENOUGH_INLIERS : LS2D_STATUS_SUCCESS;
    ls2d_result r;
    r.x = sh->bc[0].Xtx, r.y = sh->bc[0].Xty, r.theta = atan2f_fdlibm(sh->bc[0].Xs, sh->bc[0].Xc);
    r.chi_inliers = sh->tot[9], r.chi_kernelized = sh->tot[10];
    r.n_inliers = sh->n_in, r.n_kernelized = sh->n_k, r.n_corr = sh->n_corr;
    r.status = status, r.iterations = it;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.H[k] = sh->tot[k];
    r.c = sh->bc[0].Xc, r.s = sh->bc[0].Xs;
    r.lm_rejected = 0, r.reserved = 0;
    A.out[pair] = r;
  }
}

}  // namespace ls2d
