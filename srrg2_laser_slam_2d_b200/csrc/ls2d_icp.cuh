// ls2d_icp.cuh -- the run-time-shaped aligner kernels.
//
//  icp_fused_kernel   one CTA per scan pair; the whole MultiAligner2D::compute() loop on chip:
//                     fixed range image built once in shared memory, the moving cloud lives in
//                     REGISTERS for all iterations (owner-computes: the thread that owns a moving point
//                     projects it, fights for its column with two 32-bit shared-memory atomicMin passes,
//                     and -- if it won -- evaluates the correspondence itself), 11 sums + 2 counts reduced
//                     by a recursive-halving warp shuffle + one shared-memory stage, 3x3 solve and SE(2)
//                     update by thread 0.  HBM traffic = the compulsory bytes (each cloud read once,
//                     80 B written).  Cloud sizes the compile-time-stride kernel (ls2d_icp2.cuh) does not cover.
//  icp_stream_kernel  the GENERAL kernel: clouds of any size (per-point state stashed in shared memory) and every
//                     option of the aligner (Levenberg-Marquardt rounds, inlier-only runs, termination criterion).
#pragma once

#include "ls2d_common.cuh"

namespace ls2d {

template <int T, int PPT, bool SENSOR, int MINB>
__global__ void __launch_bounds__(T, MINB) icp_fused_kernel(const dev_params P, const align_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C      = P.cam.cols;
  float4* fimg     = reinterpret_cast<float4*>(smem_raw);           // fixed image: x y nx ny per column
  float* fdepth    = reinterpret_cast<float*>(fimg + C);            // fixed image: rho, < 0 = empty
  unsigned* zdepth = reinterpret_cast<unsigned*>(fdepth + C);       // z-buffer pass 1: min rho bits
  unsigned* zidx   = zdepth + C;                                    // z-buffer pass 2: min index among ties
  float2* mnrm     = reinterpret_cast<float2*>(zidx + C + (C & 1)); // [T * PPT] moving normals (phase 2 only)
  float* red       = reinterpret_cast<float*>(mnrm + T * PPT);      // [T/32][RED_STRIDE]
  pose_bc* bc      = reinterpret_cast<pose_bc*>(red + (T / 32) * RED_STRIDE);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = blockIdx.x + A.pair_base;
  const int fcl  = A.fixed_const >= 0 ? A.fixed_const : (A.fixed_id ? A.fixed_id[pair] : pair);
  const int mcl  = A.moving_id ? A.moving_id[pair / A.moving_div] : pair / A.moving_div;
  const int f0 = A.fixed_off[fcl], nf = A.fixed_off[fcl + 1] - f0;
  const int m0 = A.moving_off[mcl], nm = A.moving_off[mcl + 1] - m0;

  for (int k = tid; k < C; k += T) {
    fdepth[k] = -1.f;
    zdepth[k] = Z_EMPTY_DEPTH;
    zidx[k]   = Z_EMPTY_IDX;
  }
  // moving cloud (issued early; consumed after the fixed image is built): coordinates -> registers for all
  // iterations, normals -> shared memory (only winners read them; the register allocator would spill them to
  // local memory otherwise)
  float2 mp[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i    = tid + j * T;
    const float4 m = i < nm ? ldg4(A.moving_pts + m0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    mp[j]          = make_float2(m.x, m.y);
    mnrm[i]        = make_float2(m.z, m.w);
  }
  if (tid == 0) {
    const iso X = load_pose(A.init_pose, (size_t) pair, A.pose_stride);
    publish_pose(bc, P, X, SENSOR, 0);
    bc->tie = 0;
  }
  __syncthreads();

  // ---- fixed range image: identity camera (R/registration/correspondence_finder_projective_2d.cpp:37-44)
  {
    float4 fp[PPT];
    int col[PPT];
    unsigned rb[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int i = tid + j * T;
      col[j]      = -1;
      rb[j]       = 0;
      if (i < nf) {
        fp[j]           = ldg4(A.fixed_pts + f0 + i);
        const float rho = fsqrt(fadd(fmul(fp[j].x, fp[j].x), fmul(fp[j].y, fp[j].y)));
        if (!(rho < P.range_min || rho > P.range_max)) {
          col[j] = polar_column(P.cam, fp[j].y, fp[j].x);
          rb[j]  = f2u(rho);
          if (col[j] >= 0) atomicMin(&zdepth[col[j]], rb[j]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (col[j] >= 0 && zdepth[col[j]] == rb[j]) atomicMin(&zidx[col[j]], (unsigned) (tid + j * T));
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (col[j] >= 0 && zidx[col[j]] == (unsigned) (tid + j * T)) {
        fimg[col[j]]   = fp[j];
        fdepth[col[j]] = u2f(rb[j]);
      }
    __syncthreads();
    // hand the z-buffer back empty for the moving cloud (a separate pass: the losers of a column were still
    // reading zidx above; every toucher writes the same EMPTY values)
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (col[j] >= 0) {
        zdepth[col[j]] = Z_EMPTY_DEPTH;
        zidx[col[j]]   = Z_EMPTY_IDX;
      }
    __syncthreads();
  }

  // ---- ICP loop (MultiAligner2D::compute; L0.json:487-517)
  const int max_it = A.score_only ? 1 : P.max_iterations;
  int it           = 0;
  int status       = -1;
  float tot        = 0.f;  // lane s of warp 0: total of slot s for the last linearisation
  unsigned tot_cnt = 0;
  // z-buffer winner = lowest index among the points of minimal rho (decision D3).  Equal rho bits in one
  // column are rare, so an iteration first runs OPTIMISTIC: one atomicMin pass, then every minimal-rho point
  // claims its cell with a CAS; a failed claim means a tie, and the whole iteration is redone EXACT with the
  // index tie-break pass (one more barrier).  `exact` is uniform over the CTA.
  bool exact = false;
  for (; it < max_it; ++it) {
    if (exact) __syncthreads();  // redo pass: tie flag cleared and all cells handed back
    int col[PPT];
    unsigned rb[PPT];
    // phase 1: project the moving cloud (camera = local_map_in_sensor^-1, .cpp:47-48), z-buffer pass 1
    {
    const float Lc = bc->Lc, Ls = bc->Ls, Wtx = bc->Wtx, Wty = bc->Wty;  // phase-1 copies die at the barrier
    // the three points of a thread run as independent straight-line chains (transform, rho, fast column); the rare
    // points whose column only the exact atan2 may decide are visited afterwards
    f2 pc[PPT];
    bool near[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const f2 ra = mul2s(mk2(Lc, Ls), mp[j].x), rb2 = mul2s(mk2(-Ls, Lc), mp[j].y);
      pc[j]       = add2(mk2(fadd(ra.x, rb2.x), fadd(ra.y, rb2.y)), mk2(Wtx, Wty));
      const f2 pq = mul2(pc[j], pc[j]);
      const float rho = fsqrt(fadd(pq.x, pq.y));
      rb[j]       = f2u(rho);
      col[j]      = polar_column_fast(P.cam, pc[j].y, pc[j].x, near[j]);
      near[j]     = near[j] && !(rho < P.range_min || rho > P.range_max);
    }
    bool any_near = false;
#pragma unroll
    for (int j = 0; j < PPT; ++j) any_near |= near[j];
    if (any_near) {
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (near[j]) col[j] = polar_column_exact(P.cam, pc[j].y, pc[j].x);
    }
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const float rho = u2f(rb[j]);
      const bool ok   = tid + j * T < nm && !(rho < P.range_min || rho > P.range_max) && col[j] >= 0 && col[j] < C;
      col[j]          = ok ? col[j] : -1;
      if (ok) atomicMin(&zdepth[col[j]], rb[j]);
    }
    }
    __syncthreads();
    const float Xtx = bc->Xtx, Xty = bc->Xty, Lc = bc->Lc, Ls = bc->Ls;  // re-read: shorter live ranges than 6 registers
    if (exact) {  // z-buffer pass 2: lowest index among equal depths
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (col[j] >= 0 && zdepth[col[j]] == rb[j]) atomicMin(&zidx[col[j]], (unsigned) (tid + j * T));
      __syncthreads();
    }
    // phase 2: winners gate against the fixed column (.cpp:61-73) and linearise their correspondence
    float acc[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) acc[s] = 0.f;
    unsigned cnt = 0;  // n_inliers | n_kernelized << 16
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (col[j] < 0 || zdepth[col[j]] != rb[j]) continue;
      const int c = col[j];
      if (exact) {
        if (zidx[c] != (unsigned) (tid + j * T)) continue;
      } else if (atomicCAS(&zidx[c], Z_EMPTY_IDX, (unsigned) (tid + j * T)) != Z_EMPTY_IDX) {
        bc->tie = 1;  // two points of equal minimal rho in one column: redo this iteration exactly
        continue;
      }
      const float2 mn = mnrm[tid + j * T];
      linearize_point<SENSOR>(P, bc, fdepth[c], fimg[c], make_float4(mp[j].x, mp[j].y, mn.x, mn.y), u2f(rb[j]), Xtx,
                              Xty, Lc, Ls, acc, cnt);
    }
    store_partials(acc, cnt, red, lane, warp);
    __syncthreads();
    // hand the touched cells back for the next pass (every toucher writes the same EMPTY values)
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (col[j] >= 0) {
        zdepth[col[j]] = Z_EMPTY_DEPTH;
        zidx[col[j]]   = Z_EMPTY_IDX;
      }
    if (!exact && bc->tie) {  // uniform: bc->tie was written before the barrier above
      __syncthreads();        // every thread has read the flag and handed its cells back
      if (tid == 0) bc->tie = 0;
      exact = true;
      --it;
      continue;               // the next pass starts after the barrier at the loop head
    }
    exact = false;
    if (warp == 0) warp0_update<T, SENSOR>(P, A, bc, red, pair, it, lane, tot, tot_cnt);
    __syncthreads();
    if (bc->stop) {
      status = bc->stop - 1;
      break;
    }
  }

  if (tid < 32) write_result(P, A, bc, pair, it, status, tot, tot_cnt);
}

// Phase 1 of the streaming kernels: every moving point is seen from the camera (W = inverse(inverse(
// local_map_in_sensor)), decision D13), its rho and column are stashed in shared memory and its rho fights for the
// column's z-buffer cell.  U points of a thread at a time run as straight-line chains; the rare points whose
// column only the exact atan2 may decide are visited afterwards.
template <int T, int U, typename Load>
__device__ __forceinline__ void project_and_stash(const dev_params& P, const pose_bc* bc, int nm, Load load,
                                                  unsigned short* scol, unsigned* srho, unsigned* zdepth) {
  const float Lc = bc->Lc, Ls = bc->Ls, Wtx = bc->Wtx, Wty = bc->Wty;
  const int C = P.cam.cols;
  for (int i0 = threadIdx.x; i0 < nm; i0 += U * T) {
    f2 pc[U];
    unsigned rbv[U];
    int colv[U];
    bool near[U], up[U];
    bool any_near = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i    = i0 + u * T;
      const float4 M = i < nm ? load(i) : make_float4(0.f, 0.f, 0.f, 0.f);
      const f2 ra = mul2s(mk2(Lc, Ls), M.x), rb2 = mul2s(mk2(-Ls, Lc), M.y);
      pc[u]       = add2(mk2(fadd(ra.x, rb2.x), fadd(ra.y, rb2.y)), mk2(Wtx, Wty));
      const f2 pq = mul2(pc[u], pc[u]);
      const float rho = fsqrt(fadd(pq.x, pq.y));
      rbv[u]      = f2u(rho);
      colv[u]     = polar_column_fast2(P.cam, pc[u].y, pc[u].x, near[u], up[u]);
      near[u]     = near[u] && i < nm && !(rho < P.range_min || rho > P.range_max);
      any_near |= near[u];
    }
    if (any_near) {  // rare: side of the rounding edge's ray (when the camera carries an edge table), then exact atan2f
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (near[u]) {
          bool undecided;
          const int c2 = polar_column_edge(P.cam, pc[u].y, pc[u].x, u2f(rbv[u]), colv[u] + (up[u] ? 1 : 0), undecided);
          colv[u]      = undecided ? polar_column_exact(P.cam, pc[u].y, pc[u].x) : c2;
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * T;
      if (i < nm) {
        const float rho = u2f(rbv[u]);
        const bool ok   = !(rho < P.range_min || rho > P.range_max) && colv[u] >= 0 && colv[u] < C;
        scol[i]         = (unsigned short) (ok ? colv[u] : 0xFFFF);
        srho[i]         = rbv[u];
        if (ok) atomicMin(&zdepth[colv[u]], rbv[u]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// icp_stream_kernel: same algorithm and the same reduction shape (thread t owns points t, t+T, ...; ascending) for
// clouds of any size.  Per-point state does not live in registers: column and rho of every point are stashed in
// shared memory (6 B/point) by the projection pass and re-read by the later passes; the points themselves are
// either staged in shared memory once (MP_SMEM, 16 B/point) or re-read from global memory / L2 every pass.
constexpr size_t icp_stream_smem_bytes(int cols, int threads, int max_points, bool mp_smem) {
  return (size_t) cols * (16 + 4 + 4 + 4) + (size_t)(threads / 32) * RED_STRIDE * 4 + sizeof(pose_bc) + 16 +
         (size_t) max_points * 4 + (size_t)((max_points + 1) / 2) * 4 + 16 + (mp_smem ? (size_t) max_points * 16 : 0);
}

template <int T, bool SENSOR, bool MP_SMEM, int MINB>
__global__ void __launch_bounds__(T, MINB) icp_stream_kernel(const dev_params P, const align_args A, int max_points) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C          = P.cam.cols;
  float4* fimg         = reinterpret_cast<float4*>(smem_raw);
  float4* smp          = fimg + C;                                            // MP_SMEM: the moving cloud
  float* fdepth        = reinterpret_cast<float*>(smp + (MP_SMEM ? max_points : 0));
  unsigned* zdepth     = reinterpret_cast<unsigned*>(fdepth + C);
  unsigned* zidx       = zdepth + C;
  unsigned* srho       = zidx + C;                                            // [max_points] rho bits
  unsigned short* scol = reinterpret_cast<unsigned short*>(srho + max_points);  // [max_points] column, 0xFFFF = none
  float* red           = reinterpret_cast<float*>(scol + 2 * ((max_points + 1) / 2));
  pose_bc* bc          = reinterpret_cast<pose_bc*>(red + (T / 32) * RED_STRIDE);

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
  if (MP_SMEM)
    for (int i = tid; i < nm; i += T) smp[i] = ldg4(mpts + i);
  if (tid == 0) {
    const iso X = load_pose(A.init_pose, (size_t) pair, A.pose_stride);
    publish_pose(bc, P, X, SENSOR, 0);
    bc->tie = 0;
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
  for (int i = tid; i < nf; i += T) {
    const unsigned c = scol[i];
    if (c != 0xFFFF) zdepth[c] = Z_EMPTY_DEPTH, zidx[c] = Z_EMPTY_IDX;
  }
  __syncthreads();

  const int max_it = A.score_only ? 1 : P.max_iterations;
  int it           = 0;
  int status       = -1;
  float tot        = 0.f;
  unsigned tot_cnt = 0;
  bool exact       = false;
  for (; it < max_it; ++it) {
    if (exact) __syncthreads();
    project_and_stash<T, 4>(P, bc, nm, [&](int i) { return MP_SMEM ? smp[i] : ldg4(mpts + i); }, scol, srho, zdepth);
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
    // phase 2: a thread first finds the z-buffer winners among its next 32 points (cheap scan), then linearises only
    // those, in ascending order: with far more points than columns most points lose, and a warp now runs the heavy
    // path max-winners-per-lane times instead of once per scanned point
    for (int k0 = 0; tid + k0 * T < nm; k0 += 32) {
      unsigned wmask = 0;
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        const int i = tid + (k0 + k) * T;
        if (i >= nm) break;
        const unsigned c = scol[i];
        if (c == 0xFFFF || zdepth[c] != srho[i]) continue;
        if (exact) {
          if (zidx[c] != (unsigned) i) continue;
        } else if (atomicCAS(&zidx[c], Z_EMPTY_IDX, (unsigned) i) != Z_EMPTY_IDX) {
          bc->tie = 1;
          continue;
        }
        wmask |= 1u << k;
      }
      while (wmask) {
        const int k = __ffs(wmask) - 1;
        wmask &= wmask - 1;
        const int i      = tid + (k0 + k) * T;
        const unsigned c = scol[i];
        const float4 M   = MP_SMEM ? smp[i] : ldg4(mpts + i);
        linearize_point<SENSOR>(P, bc, fdepth[c], fimg[c], M, u2f(srho[i]), Xtx, Xty, Lc, Ls, acc, cnt);
      }
    }
    store_partials(acc, cnt, red, lane, warp);
    __syncthreads();
    for (int k = tid; k < C; k += T) zdepth[k] = Z_EMPTY_DEPTH, zidx[k] = Z_EMPTY_IDX;  // wholesale: C <= points
    if (!exact && bc->tie) {
      __syncthreads();
      if (tid == 0) bc->tie = 0;
      exact = true;
      --it;
      continue;
    }
    exact = false;
    if (warp == 0) warp0_update<T, SENSOR>(P, A, bc, red, pair, it, lane, tot, tot_cnt);
    __syncthreads();
    if (bc->stop) {
      status = bc->stop - 1;
      break;
    }
  }
  if (tid < 32) write_result(P, A, bc, pair, it, status, tot, tot_cnt);
}

}  // namespace ls2d
