// ls2d_multi2.cuh -- icp_multi2_kernel: the multi-slice aligner (MultiAligner2D with two laser slice processors and an
// optional odometry prior fused into ONE 3x3 system per iteration -- MULTI.json:700-730: al_sl_laser_0 + ad_sl_odom +
// al_sl_laser_1) in the style of icp_fused2_kernel (ls2d_icp2.cuh), for the shape of the shipped configurations: two
// laser slices with their own fixed clouds ("points_0", "points_1": up to 768 points, canvases below 768 columns) that
// align ONE shared moving cloud (the local map "points": up to 1536 points), plane-to-plane factors.
//
// One CTA per pair.  Every slice owns a fixed range image and a z-buffer in shared memory (28 B per column with the
// compile-time stride CS, explicit shared-space addressing); the shared moving cloud lives in registers for all
// iterations and is projected once per slice (owner-computes), its normals in shared memory.  Per iteration: [project both slices' moving clouds, fight for the
// columns] | [winners of slice 0 linearise, per-warp reduction; the same for slice 1] | [warp 0: totals per slice,
// slice gating (decision D14), sum in slice order + prior (D15, D16), binary64 3x3 solve, both slices' next poses] --
// three barriers, where the stash-in-shared-memory kernel (ls2d_multi.cuh) spends seven per slice.
// Reduction shape per slice: thread t owns moving points t, t + T, ...; per warp two ascending 16-lane halves
// (transposed tile); warps in order (ORC_SUM_TREE with bit 16; ls2d_multi_reduction_shape()).
#pragma once

#include "ls2d_icp2.cuh"
#include "ls2d_multi.cuh"

namespace ls2d {

template <int T, int PPF, int PPM, int CS>
struct multi2_map {
  static constexpr int NW    = T / 32;
  static constexpr int WTILE = NSUM * 36 * 4;                      // per-warp transposed tile: 1584 B
  static constexpr int MNRM  = 0;                                  // float2[T * PPM] normals of the shared moving cloud
  static constexpr int WRED  = MNRM + T * PPM * 8;                 // NW tiles
  static constexpr int RED   = WRED + NW * WTILE;                  // float[2][NW][RED_STRIDE] warp totals per slice
  static constexpr int SH    = RED + 2 * NW * RED_STRIDE * 4;      // multi_shared
  static constexpr int Z     = (SH + (int) sizeof(multi_shared) + 15) & ~15;  // slice 0's columns; slice 1 at + 28 * CS
  static constexpr int ZI = 4 * CS, FD = 8 * CS, FI = 12 * CS;     // from a slice's Z: zidx, fixed rho, fixed point
  static constexpr int SLICE = 28 * CS;
  static constexpr int BYTES = Z + 2 * SLICE;
  static_assert(CS % 4 == 0, "column stride keeps the float4 images 16-byte aligned");
};

template <int T, int PPF, int PPM, int CS, bool FUSED>
__global__ void __launch_bounds__(T, 3) icp_multi2_kernel(const multi_args A) {
  using M = multi2_map<T, PPF, PPM, CS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned sb = sm::addr(smem_raw);
  multi_shared* sh  = reinterpret_cast<multi_shared*>(smem_raw + M::SH);
  float* red        = reinterpret_cast<float*>(smem_raw + M::RED);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = blockIdx.x;
  const int fcl  = A.fixed_id ? A.fixed_id[pair] : pair;
  const int mcl  = A.moving_id ? A.moving_id[pair] : pair;
  static_assert(offsetof(multi_shared, bc) == 0, "the slices' pose blocks lead the shared record");

  // ---- empty z-buffers and fixed images, moving clouds (coordinates -> registers, normals -> shared memory), poses
  for (int k = tid; k < 2 * CS; k += T) {
    const int s = k >= CS, c = k - s * CS;
    const unsigned a = sb + M::Z + s * M::SLICE + 4u * c;
    sm::st_f32<M::FD>(a, -1.f);
    sm::st_u32<0>(a, Z_EMPTY_DEPTH);
    sm::st_u32<M::ZI>(a, Z_EMPTY_IDX);
  }
  float2 mp[PPM];  // the moving cloud both slices share (the launcher checks that they do)
  {
    const int m0 = A.sl[0].moving_off[mcl], nm = A.sl[0].moving_off[mcl + 1] - m0;
#pragma unroll
    for (int j = 0; j < PPM; ++j) {
      const int i = tid + j * T;
      // lanes past the end of the cloud hold a point no pose brings inside the range gates (its squared range overflows)
      const float4 m = i < nm ? ldg4_once(A.sl[0].moving_pts + m0 + i) : make_float4(1e30f, 0.f, 0.f, 0.f);
      mp[j]          = make_float2(m.x, m.y);
      sm::st_f32x2<0>(sb + M::MNRM + 8u * (unsigned) (j * T + tid), m.z, m.w);
    }
  }
  if (tid == 0) {
    const iso X = load_pose(A.init_pose, (size_t) pair, A.pose_stride);
    for (int s = 0; s < 2; ++s) {
      publish_pose(&sh->bc[s], A.sl[s].P, X, A.sl[s].P.with_sensor != 0, 0);
      sh->bc[s].tie = 0;
    }
    if (A.prior_z) sh->Zinv = iso_inverse(load_pose(A.prior_z, (size_t) pair, A.pose_stride));
    for (int k = 0; k < NSUM; ++k) sh->tot[k] = 0.f;
    sh->n_in = sh->n_k = sh->n_corr = 0;
  }
  __syncthreads();

  // ---- fixed range images of both slices: identity camera (correspondence_finder_projective_2d.cpp:37-44), exact
  // two-pass z-buffer
  {
    float4 fp[2][PPF];
    unsigned za[2][PPF], rb[2][PPF];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const dev_params& P = A.sl[s].P;
      const int C  = P.cam.cols;
      const int f0 = A.sl[s].fixed_off[fcl], nf = A.sl[s].fixed_off[fcl + 1] - f0;
      const unsigned zb = sb + M::Z + s * M::SLICE;
#pragma unroll
      for (int j = 0; j < PPF; ++j) {
        const int i = tid + j * T;
        int col     = C;
        rb[s][j]    = 0;
        if (i < nf) {
          fp[s][j]        = ldg4_once(A.sl[s].fixed_pts + f0 + i);
          const float rho = fsqrt(fadd(fmul(fp[s][j].x, fp[s][j].x), fmul(fp[s][j].y, fp[s][j].y)));
          if (!(rho < P.range_min || rho > P.range_max)) {
            const int c = polar_column(P.cam, fp[s][j].y, fp[s][j].x);
            if (c >= 0) col = c, rb[s][j] = f2u(rho);
          }
        }
        za[s][j] = zb + 4u * col;
        if (col != C) sm::atom_min_u32<0>(za[s][j], rb[s][j]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int j = 0; j < PPF; ++j)
        if (sm::ld_u32<0>(za[s][j]) == rb[s][j]) sm::atom_min_u32<M::ZI>(za[s][j], (unsigned) (tid + j * T));
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const unsigned fk = 3u * (sb + M::Z + s * M::SLICE) - M::FI;  // fixed point of the cell za: 4 za - fk
#pragma unroll
      for (int j = 0; j < PPF; ++j)
        if (sm::ld_u32<0>(za[s][j]) == rb[s][j] && sm::ld_u32<M::ZI>(za[s][j]) == (unsigned) (tid + j * T)) {
          sm::st_f32x4<0>(4u * za[s][j] - fk, fp[s][j]);
          sm::st_f32<M::FD>(za[s][j], u2f(rb[s][j]));
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int j = 0; j < PPF; ++j) {
        sm::st_u32<0>(za[s][j], Z_EMPTY_DEPTH);
        sm::st_u32<M::ZI>(za[s][j], Z_EMPTY_IDX);
      }
    __syncthreads();
  }

  // ---- ICP loop (MultiAligner2D::compute; MULTI.json:700-730)
  const int max_it = A.score_only ? 1 : A.sl[0].P.max_iterations;  // aligner-level (D17)
  int it           = 0;
  int status       = -1;
  const unsigned wt = sb + M::WRED + (unsigned) warp * M::WTILE;

  // phase 1 of a slice: project the shared moving cloud with the slice's pose (camera = local_map_in_sensor^-1,
  // .cpp:47-48) and fight for the columns of the slice's z-buffer
  auto project_slice = [&](const int s, unsigned (&za)[PPM], unsigned (&rb)[PPM]) {
    const dev_params& P = A.sl[s].P;
    const int C       = P.cam.cols;
    const unsigned zb = sb + M::Z + s * M::SLICE;
    const pose_bc* bc = &sh->bc[s];
    const unsigned tie = sb + M::SH + (unsigned) (s * sizeof(pose_bc)) + (unsigned) offsetof(pose_bc, tie);
    const float Lc = bc->Lc, Ls = bc->Ls, Wtx = bc->Wtx, Wty = bc->Wty;
    f2 pc[PPM];
    int col[PPM];
    bool near[PPM], up[PPM], in[PPM];
#pragma unroll
    for (int j = 0; j < PPM; ++j) {
      const f2 ra = mul2s(mk2(Lc, Ls), mp[j].x), rb2 = mul2s(mk2(-Ls, Lc), mp[j].y);
      pc[j]       = add2(mk2(fadd(ra.x, rb2.x), fadd(ra.y, rb2.y)), mk2(Wtx, Wty));
      const f2 pq = mul2(pc[j], pc[j]);
      const float a = fadd(pq.x, pq.y);
      in[j]       = a >= P.gate2.lo && a <= P.gate2.hi;  // the range gate on the squared range (ls2d_math.cuh)
      rb[j]       = f2u(fsqrt_gated(a));
      col[j]      = polar_column_fast2(P.cam, pc[j].y, pc[j].x, near[j], up[j]);
      near[j]     = near[j] && in[j];
    }
    bool any_near = false;
#pragma unroll
    for (int j = 0; j < PPM; ++j) any_near |= near[j];
    if (any_near) {  // rare, out of line: side of the rounding edge (binary32), then the exact atan2f
#pragma unroll
      for (int j = 0; j < PPM; ++j)
        if (near[j]) col[j] = polar_column_resolve(P.cam, pc[j].y, pc[j].x, u2f(rb[j]), col[j], up[j]);
    }
#pragma unroll
    for (int j = 0; j < PPM; ++j) {
      const bool ok = in[j] && (unsigned) col[j] < (unsigned) C;
      za[j]         = zb + 4u * (ok ? col[j] : C);
      rb[j]         = ok ? rb[j] : 0u;  // never equals the dummy cell's EMPTY
      if (ok && sm::atom_min_u32<0>(za[j], rb[j]) == rb[j]) sm::st_u32<0>(tie, 1u);  // an equal rho was there
    }
  };
  // phase 2 of a slice (after a barrier): the winners gate against the fixed column (.cpp:61-73), linearise their
  // correspondence and hand the column's cell back themselves (a column has one winner; the losers only ever see its
  // rho or EMPTY); per-warp reduction into the slice's rows
  auto linearise_slice = [&](const int s, const unsigned (&za)[PPM], const unsigned (&rb)[PPM]) {
    const dev_params& P = A.sl[s].P;
    const pose_bc* bc   = &sh->bc[s];
    const unsigned fk   = 3u * (sb + M::Z + s * M::SLICE) - M::FI;
    const bool tied = bc->tie != 0;  // uniform; rare: lowest index among equal rho (decision D3), one more barrier
    if (tied) {
#pragma unroll
      for (int j = 0; j < PPM; ++j)
        if (sm::ld_u32<0>(za[j]) == rb[j]) sm::atom_min_u32<M::ZI>(za[j], (unsigned) (tid + j * T));
      __syncthreads();
    }
    const float Xtx = bc->Xtx, Xty = bc->Xty, Lc = bc->Lc, Ls = bc->Ls, Xc = bc->Xc, Xs = bc->Xs;
    float acc[NSUM];
#pragma unroll
    for (int q = 0; q < NSUM; ++q) acc[q] = 0.f;
    unsigned cnt = 0;  // n_inliers | n_kernelized << 16
#pragma unroll
    for (int j = 0; j < PPM; ++j) {
      bool win = sm::ld_u32<0>(za[j]) == rb[j];
      if (tied) win = win && sm::ld_u32<M::ZI>(za[j]) == (unsigned) (tid + j * T);
      if (win) {
        const float fd  = sm::ld_f32<M::FD>(za[j]);
        const float4 F  = sm::ld_f32x4<0>(4u * za[j] - fk);
        const float2 Mn = sm::ld_f32x2<0>(sb + M::MNRM + 8u * (unsigned) (j * T + tid));
        // FUSED: the accumulation arithmetic of decision D18 (ls2d_icp2.cuh: linearize2f), else single-rounding sums
        if (P.with_sensor) {
          if (FUSED)
            linearize2f<true, false>(P, fd, F, mp[j].x, mp[j].y, Mn, u2f(rb[j]), Xtx, Xty, Xc, Xs, Lc, Ls, acc, cnt);
          else
            linearize2<true, false>(P, fd, F, mp[j].x, mp[j].y, Mn, u2f(rb[j]), Xtx, Xty, Xc, Xs, Lc, Ls, acc, cnt);
        } else {
          if (FUSED)
            linearize2f<false, false>(P, fd, F, mp[j].x, mp[j].y, Mn, u2f(rb[j]), Xtx, Xty, Xc, Xs, Lc, Ls, acc, cnt);
          else
            linearize2<false, false>(P, fd, F, mp[j].x, mp[j].y, Mn, u2f(rb[j]), Xtx, Xty, Xc, Xs, Lc, Ls, acc, cnt);
        }
        sm::st_u32<0>(za[j], Z_EMPTY_DEPTH);
        if (tied) sm::st_u32<M::ZI>(za[j], Z_EMPTY_IDX);
      }
    }
    store_partials2(acc, cnt, wt, sb + M::RED + (unsigned) ((s * M::NW + warp) * RED_STRIDE * 4), lane);
    __syncwarp();  // the tile is reused by the next slice
  };

  for (; it < max_it; ++it) {
    {
      unsigned za[PPM], rb[PPM];
      project_slice(0, za, rb);
      __syncthreads();
      linearise_slice(0, za, rb);
    }
    {
      unsigned za[PPM], rb[PPM];
      project_slice(1, za, rb);
      __syncthreads();
      linearise_slice(1, za, rb);
    }
    __syncthreads();

    // ---- totals, gates, prior, Gauss-Newton step: warp 0 (the logic of icp_multi_kernel, decisions D14-D17)
    if (warp == 0) {
      if (lane < 2) sh->bc[lane].tie = 0;  // everybody read the flags before the barrier above
      float v[NSUM];
      int n_in = 0, n_k = 0, n_corr = 0, n_found = 0;
      bool contributed = false;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float* r = red + s * M::NW * RED_STRIDE;
        float t        = 0.f;
        unsigned c     = 0;
        if (lane < NSUM) {
          t = r[lane];
#pragma unroll
          for (int w = 1; w < M::NW; ++w) t = fadd(t, r[w * RED_STRIDE + lane]);
          t = fadd(t, 0.f);
        } else if (lane == NSUM) {
#pragma unroll
          for (int w = 0; w < M::NW; ++w) c += __float_as_uint(r[w * RED_STRIDE + NSUM]);
        }
        const unsigned c2 = __shfl_sync(0xffffffffu, c, NSUM);
        const int s_in = c2 & 0xffff, s_k = c2 >> 16;
        n_found += s_in + s_k;
        const bool use = s_in + s_k > A.sl[s].P.min_num_correspondences;  // D14 (uniform over the warp)
#pragma unroll
        for (int k = 0; k < NSUM; ++k) {
          const float x = __shfl_sync(0xffffffffu, t, k);
          if (use) v[k] = contributed ? fadd(v[k], x) : x;  // D16
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
        for (int s = 0; s < 2; ++s) publish_pose(&sh->bc[s], A.sl[s].P, X, A.sl[s].P.with_sensor != 0, stop);
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
