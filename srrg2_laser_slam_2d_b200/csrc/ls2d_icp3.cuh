// ls2d_icp3.cuh -- icp_duo_kernel: two pairs per CTA, the Gauss-Newton step on a warp of its own.
//
// icp_fused2_kernel spends ~15 % of its warp-time at the barrier behind warp 0's serial section (sum of the warps'
// partials -> 3x3 LDL^T in binary64 -> sincos -> next pose, ~330 dependent instructions): nothing else of THAT pair can
// run until the pose exists.  Here a CTA owns TWO pairs, A and B, and a 33rd..(TC/32+1)-th warp that owns no points:
//
//   compute warps:  [wait pose A] project A | gate+linearise A | [wait pose B] project B | gate+linearise B | ...
//   solver warp:                                                update A ------------------>  update B -------->
//
// so the update of one pair runs under the projection / linearisation of the other and never sits in front of a
// barrier the compute warps wait at.  The compute warps synchronise among themselves with named barrier 1 (z-buffer
// phases); "partials of X ready" (compute arrive, solver sync) and "pose of X published" (solver arrives, compute
// sync) are named barriers of their own.  Same per-point code as icp_fused2_kernel (ls2d_icp2.cuh), two points per
// thread per pair, so 17 compute warps cover the 1081 points of a Hokuyo scan with no all-invalid slot.
// Summation shape: thread t of the TC compute threads owns points t and t + TC; warps combine as in
// icp_fused2_kernel (ls2d_reduction_shape() = TC | 1 << 16).
#pragma once

#include "ls2d_icp2.cuh"

namespace ls2d {

namespace nb {  // named barriers
constexpr int COMPUTE = 1, READY = 2 /* + x */, POSE = 4 /* + x */;
__device__ __forceinline__ void sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void arrive(int id, int n) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
}  // namespace nb

template <int TC, int CS>
struct duo_map {
  static constexpr int NWC   = TC / 32;
  static constexpr int WTILE = NSUM * 36 * 4;
  static constexpr int PAIR_MN = TC * 2 * 8;                       // float2[TC * 2] moving normals of one pair
  static constexpr int WRED  = 2 * PAIR_MN;                        // NWC transposed tiles (a warp reuses its tile)
  static constexpr int RED   = WRED + NWC * WTILE;                 // float[2][NWC][RED_STRIDE]
  static constexpr int PAIR_RED = NWC * RED_STRIDE * 4;
  static constexpr int BC    = RED + 2 * PAIR_RED;                 // pose_bc[2], 48 B apart
  static constexpr int Z     = (BC + 2 * 48 + 15) & ~15;           // per pair: zdepth, zidx, fdepth, fimg (28 CS bytes)
  static constexpr int PAIR_Z = 28 * CS;
  static constexpr int ZI = 4 * CS, FD = 8 * CS, FI = 12 * CS;     // from a pair's z base
  static constexpr int BYTES = Z + 2 * PAIR_Z;
  static_assert(CS % 4 == 0, "column stride keeps the float4 image 16-byte aligned");
};

// the solver warp's step for one pair: totals of the NW warps' partial rows in warp order, gates, 3x3 solve,
// X <- X * v2t(dx), next pose (same arithmetic and order as warp0_update<.., CANON = true>)
template <int NW, bool SENSOR>
__device__ __forceinline__ void duo_update(const dev_params& P, const align_args& A, pose_bc* bc, const float* red,
                                           int pair, int it, int lane, float& tot, unsigned& tot_cnt) {
  warp0_update<NW * 32, SENSOR, true>(P, A, bc, red, pair, it, lane, tot, tot_cnt);
}

template <int TC, bool SENSOR, int CS>
__global__ void __launch_bounds__(TC + 32, 2) icp_duo_kernel(const dev_params P, const align_args A) {
  using M = duo_map<TC, CS>;
  constexpr int NT = TC + 32;  // all threads of the CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned sb = sm::addr(smem_raw);
  const int C       = P.cam.cols;  // < CS: column C is the dummy cell of invalid points
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair0  = 2 * blockIdx.x + A.pair_base;
  const int n_here = (pair0 + 1 < A.pair_base + A.n_pairs) ? 2 : 1;  // the last CTA of an odd batch owns one pair
  const int max_it = A.score_only ? 1 : P.max_iterations;

  if (warp == M::NWC) {
    // ================================================================ solver warp
    float tot[2]        = {0.f, 0.f};
    unsigned tot_cnt[2] = {0u, 0u};
    bool alive[2]       = {true, n_here > 1};
    int its[2]          = {0, 0};
    if (lane == 0) {
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (x < n_here) {
          pose_bc* bc = reinterpret_cast<pose_bc*>(smem_raw + M::BC + x * 48);
          const int p = pair0 + x;
          const iso X = iso_v2t(A.init_xyt[3 * p], A.init_xyt[3 * p + 1], A.init_xyt[3 * p + 2]);
          publish_pose(bc, P, X, SENSOR, 0);
          bc->tie = 0;
        }
      }
    }
    __threadfence_block();
    if (max_it > 0) {
      nb::arrive(nb::POSE + 0, NT);
      if (n_here > 1) nb::arrive(nb::POSE + 1, NT);
    }
    for (int it = 0; it < max_it; ++it) {
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (!alive[x]) continue;
        pose_bc* bc      = reinterpret_cast<pose_bc*>(smem_raw + M::BC + x * 48);
        const float* red = reinterpret_cast<const float*>(smem_raw + M::RED + x * M::PAIR_RED);
        nb::sync(nb::READY + x, NT);  // the compute warps' partial rows of pair x are in shared memory
        duo_update<M::NWC, SENSOR>(P, A, bc, red, pair0 + x, it, lane, tot[x], tot_cnt[x]);
        its[x] = it + 1;
        __threadfence_block();
        const int stop = __shfl_sync(0xffffffffu, lane == 0 ? bc->stop : 0, 0);
        if (stop) alive[x] = false, its[x] = it;  // write_result reports the iteration the stop happened in
        if (it + 1 < max_it) nb::arrive(nb::POSE + x, NT);  // the compute warps read the pose (or the stop flag)
      }
    }
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      if (x < n_here) {
        const pose_bc* bc = reinterpret_cast<const pose_bc*>(smem_raw + M::BC + x * 48);
        const int stop    = __shfl_sync(0xffffffffu, lane == 0 ? bc->stop : 0, 0);
        write_result(P, A, bc, pair0 + x, its[x], stop ? stop - 1 : -1, tot[x], tot_cnt[x], TC);
      }
    }
    return;
  }

  // ================================================================== compute warps
  struct pstate {
    float2 mp[2];
    unsigned za[2], rb[2];
  } st[2];
  bool alive[2] = {true, n_here > 1};

  // ---- per pair: empty z-buffer / fixed image, moving cloud -> registers + normals -> shared memory, fixed image
  static_for<0, 2>([&](auto xc) {
    constexpr int X = decltype(xc)::value;
    if (X >= n_here) return;
    const int pair = pair0 + X;
    const unsigned zb = sb + M::Z + X * M::PAIR_Z;
    const unsigned fk = 3u * zb - M::FI;
    const int fcl  = A.fixed_const >= 0 ? A.fixed_const : (A.fixed_id ? A.fixed_id[pair] : pair);
    const int mcl  = A.moving_id ? A.moving_id[pair / A.moving_div] : pair / A.moving_div;
    const int f0 = A.fixed_off[fcl], nf = A.fixed_off[fcl + 1] - f0;
    const int m0 = A.moving_off[mcl], nm = A.moving_off[mcl + 1] - m0;
    for (int k = tid; k <= C; k += TC) {
      const unsigned a = zb + 4u * k;
      sm::st_f32<M::FD>(a, -1.f);
      sm::st_u32<0>(a, Z_EMPTY_DEPTH);
      sm::st_u32<M::ZI>(a, Z_EMPTY_IDX);
    }
    const unsigned mna = sb + X * M::PAIR_MN + 8u * tid;
    static_for<0, 2>([&](auto jc) {
      constexpr int J = decltype(jc)::value;
      const int i     = tid + J * TC;
      // lanes past the end of the cloud hold a point no pose brings inside the range gates (rho overflows to inf)
      const float4 m  = i < nm ? ldg4(A.moving_pts + m0 + i) : make_float4(1e30f, 0.f, 0.f, 0.f);
      st[X].mp[J]     = make_float2(m.x, m.y);
      sm::st_f32x2<J * TC * 8>(mna, m.z, m.w);
    });
    nb::sync(nb::COMPUTE, TC);
    float4 fp[2];
    unsigned za[2], rb[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int i = tid + j * TC;
      int col     = C;
      rb[j]       = 0;
      if (i < nf) {
        fp[j]           = ldg4(A.fixed_pts + f0 + i);
        const float rho = fsqrt(fadd(fmul(fp[j].x, fp[j].x), fmul(fp[j].y, fp[j].y)));
        if (!(rho < P.range_min || rho > P.range_max)) {
          const int c = polar_column(P.cam, fp[j].y, fp[j].x);
          if (c >= 0) col = c, rb[j] = f2u(rho);
        }
      }
      za[j] = zb + 4u * col;
      if (col != C) sm::atom_min_u32<0>(za[j], rb[j]);
    }
    nb::sync(nb::COMPUTE, TC);
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (sm::ld_u32<0>(za[j]) == rb[j]) sm::atom_min_u32<M::ZI>(za[j], (unsigned) (tid + j * TC));
    nb::sync(nb::COMPUTE, TC);
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (sm::ld_u32<0>(za[j]) == rb[j] && sm::ld_u32<M::ZI>(za[j]) == (unsigned) (tid + j * TC)) {
        sm::st_f32x4<0>(4u * za[j] - fk, fp[j]);
        sm::st_f32<M::FD>(za[j], u2f(rb[j]));
      }
    nb::sync(nb::COMPUTE, TC);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      sm::st_u32<0>(za[j], Z_EMPTY_DEPTH);
      sm::st_u32<M::ZI>(za[j], Z_EMPTY_IDX);
    }
    nb::sync(nb::COMPUTE, TC);
  });

  const unsigned wt = sb + M::WRED + (unsigned) warp * M::WTILE;
  for (int it = 0; it < max_it; ++it) {
    static_for<0, 2>([&](auto xc) {
      constexpr int X = decltype(xc)::value;
      if (!alive[X]) return;  // uniform over the CTA
      const unsigned zb  = sb + M::Z + X * M::PAIR_Z;
      const unsigned fk  = 3u * zb - M::FI;
      const unsigned bca = sb + M::BC + X * 48;
      const unsigned mna = sb + X * M::PAIR_MN + 8u * tid;
      unsigned(&za)[2]   = st[X].za;
      unsigned(&rb)[2]   = st[X].rb;
      float2(&mp)[2]     = st[X].mp;
      nb::sync(nb::POSE + X, NT);  // the solver warp has published this iteration's pose of pair X
      if (sm::ld_u32<BC_STOP>(bca)) {
        alive[X] = false;
        return;
      }
      // phase 1: project the moving cloud (camera = local_map_in_sensor^-1, .cpp:47-48) and fight for the column
      {
        const float Lc = sm::ld_f32<BC_LC>(bca), Ls = sm::ld_f32<BC_LS>(bca);
        const float Wtx = sm::ld_f32<BC_WTX>(bca), Wty = sm::ld_f32<BC_WTY>(bca);
        f2 pc[2];
        int col[2];
        bool near[2], up[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const f2 ra = mul2s(mk2(Lc, Ls), mp[j].x), rb2 = mul2s(mk2(-Ls, Lc), mp[j].y);
          pc[j]       = add2(mk2(fadd(ra.x, rb2.x), fadd(ra.y, rb2.y)), mk2(Wtx, Wty));
          const f2 pq = mul2(pc[j], pc[j]);
          const float rho = fsqrt(fadd(pq.x, pq.y));
          rb[j]       = f2u(rho);
          col[j]      = polar_column_fast2(P.cam, pc[j].y, pc[j].x, near[j], up[j]);
          near[j]     = near[j] && !(rho < P.range_min || rho > P.range_max);
        }
        if (near[0] || near[1]) {  // rare: second tier (side of the rounding edge), then the exact atan2f
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (near[j]) {
              bool undecided;
              const int c2 = polar_column_edge(P.cam, pc[j].y, pc[j].x, u2f(rb[j]), col[j] + (up[j] ? 1 : 0), undecided);
              col[j]       = undecided ? polar_column_exact(P.cam, pc[j].y, pc[j].x) : c2;
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float rho = u2f(rb[j]);
          const bool ok   = !(rho < P.range_min || rho > P.range_max) && (unsigned) col[j] < (unsigned) C;
          za[j]           = zb + 4u * (ok ? col[j] : C);
          rb[j]           = ok ? rb[j] : 0u;  // never equals the dummy cell's EMPTY
          if (ok && sm::atom_min_u32<0>(za[j], rb[j]) == rb[j]) sm::st_u32<BC_TIE>(bca, 1u);  // an equal rho was there
        }
      }
      nb::sync(nb::COMPUTE, TC);
      const bool tie = sm::ld_u32<BC_TIE>(bca) != 0;  // uniform
      if (tie) {  // exact pass: lowest index among the points of minimal rho (decision D3)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (sm::ld_u32<0>(za[j]) == rb[j]) sm::atom_min_u32<M::ZI>(za[j], (unsigned) (tid + j * TC));
        nb::sync(nb::COMPUTE, TC);
      }
      // phase 2: winners gate against the fixed column (.cpp:61-73) and linearise their correspondence
      {
        const float Xtx = sm::ld_f32<BC_XTX>(bca), Xty = sm::ld_f32<BC_XTY>(bca);
        const float Lc = sm::ld_f32<BC_LC>(bca), Ls = sm::ld_f32<BC_LS>(bca);
        float Xc = 0.f, Xs = 0.f;
        if (SENSOR) Xc = sm::ld_f32<BC_XC>(bca), Xs = sm::ld_f32<BC_XS>(bca);
        float acc[NSUM];
#pragma unroll
        for (int s = 0; s < NSUM; ++s) acc[s] = 0.f;
        unsigned cnt = 0;  // n_inliers | n_kernelized << 16
        static_for<0, 2>([&](auto jc) {
          constexpr int J = decltype(jc)::value;
          bool win        = sm::ld_u32<0>(za[J]) == rb[J];
          if (tie) win = win && sm::ld_u32<M::ZI>(za[J]) == (unsigned) (tid + J * TC);
          if (win) {
            const float fd  = sm::ld_f32<M::FD>(za[J]);
            const float4 F  = sm::ld_f32x4<0>(4u * za[J] - fk);
            const float2 Mn = sm::ld_f32x2<J * TC * 8>(mna);
            linearize2<SENSOR, J == 0>(P, fd, F, mp[J].x, mp[J].y, Mn, u2f(rb[J]), Xtx, Xty, Xc, Xs, Lc, Ls, acc, cnt);
          }
        });
        store_partials2(acc, cnt, wt, sb + M::RED + X * M::PAIR_RED + (unsigned) warp * (RED_STRIDE * 4), lane);
      }
      __threadfence_block();
      nb::arrive(nb::READY + X, NT);  // the solver warp may take pair X from here
      nb::sync(nb::COMPUTE, TC);      // every compute warp is through with the z-buffer of pair X
      // hand the touched cells back for the next pass (every toucher writes the same EMPTY values)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sm::st_u32<0>(za[j], Z_EMPTY_DEPTH);
        if (tie) sm::st_u32<M::ZI>(za[j], Z_EMPTY_IDX);
      }
      if (tie && tid == 0) sm::st_u32<BC_TIE>(bca, 0u);  // everybody read it before the barrier above
    });
  }
}

// icp_joint_kernel: two pairs per CTA WITHOUT a solver warp -- the four projection chains of a thread (two points of
// pair A, two of pair B) run interleaved as in icp_fused2_kernel<288, 4>, both pairs share every barrier (3 per
// iteration for two pairs), and the two Gauss-Newton steps run side by side on warps 0 and 1.  TC threads own points
// t and t + TC of either pair (no all-invalid slot at 1081 points); same summation shape as icp_duo_kernel.
template <int TC, bool SENSOR, int CS>
__global__ void __launch_bounds__(TC, 2) icp_joint_kernel(const dev_params P, const align_args A) {
  using M = duo_map<TC, CS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned sb = sm::addr(smem_raw);
  const int C       = P.cam.cols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair0  = 2 * blockIdx.x + A.pair_base;
  const int n_here = (pair0 + 1 < A.pair_base + A.n_pairs) ? 2 : 1;
  const int max_it = A.score_only ? 1 : P.max_iterations;

  float2 mp[4];          // [2 X + J]
  unsigned za[4], rb[4];
  // ---- prologue per pair (as icp_duo_kernel, whole CTA)
  static_for<0, 2>([&](auto xc) {
    constexpr int X = decltype(xc)::value;
    pose_bc* bc = reinterpret_cast<pose_bc*>(smem_raw + M::BC + X * 48);
    if (X >= n_here) {  // absent pair: never alive, its points never valid
      mp[2 * X] = mp[2 * X + 1] = make_float2(1e30f, 0.f);
      if (tid == 0) {
        publish_pose(bc, P, iso_identity(), SENSOR, 1 + LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES);
        bc->tie = 0;
      }
      return;
    }
    const int pair = pair0 + X;
    const unsigned zb = sb + M::Z + X * M::PAIR_Z;
    const unsigned fk = 3u * zb - M::FI;
    const int fcl  = A.fixed_const >= 0 ? A.fixed_const : (A.fixed_id ? A.fixed_id[pair] : pair);
    const int mcl  = A.moving_id ? A.moving_id[pair / A.moving_div] : pair / A.moving_div;
    const int f0 = A.fixed_off[fcl], nf = A.fixed_off[fcl + 1] - f0;
    const int m0 = A.moving_off[mcl], nm = A.moving_off[mcl + 1] - m0;
    for (int k = tid; k <= C; k += TC) {
      const unsigned a = zb + 4u * k;
      sm::st_f32<M::FD>(a, -1.f);
      sm::st_u32<0>(a, Z_EMPTY_DEPTH);
      sm::st_u32<M::ZI>(a, Z_EMPTY_IDX);
    }
    const unsigned mna = sb + X * M::PAIR_MN + 8u * tid;
    static_for<0, 2>([&](auto jc) {
      constexpr int J = decltype(jc)::value;
      const int i     = tid + J * TC;
      const float4 m  = i < nm ? ldg4(A.moving_pts + m0 + i) : make_float4(1e30f, 0.f, 0.f, 0.f);
      mp[2 * X + J]   = make_float2(m.x, m.y);
      sm::st_f32x2<J * TC * 8>(mna, m.z, m.w);
    });
    if (tid == 0) {
      const iso Xi = iso_v2t(A.init_xyt[3 * pair], A.init_xyt[3 * pair + 1], A.init_xyt[3 * pair + 2]);
      publish_pose(bc, P, Xi, SENSOR, 0);
      bc->tie = 0;
    }
    __syncthreads();
    float4 fp[2];
    unsigned fza[2], frb[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int i = tid + j * TC;
      int col     = C;
      frb[j]      = 0;
      if (i < nf) {
        fp[j]           = ldg4(A.fixed_pts + f0 + i);
        const float rho = fsqrt(fadd(fmul(fp[j].x, fp[j].x), fmul(fp[j].y, fp[j].y)));
        if (!(rho < P.range_min || rho > P.range_max)) {
          const int c = polar_column(P.cam, fp[j].y, fp[j].x);
          if (c >= 0) col = c, frb[j] = f2u(rho);
        }
      }
      fza[j] = zb + 4u * col;
      if (col != C) sm::atom_min_u32<0>(fza[j], frb[j]);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (sm::ld_u32<0>(fza[j]) == frb[j]) sm::atom_min_u32<M::ZI>(fza[j], (unsigned) (tid + j * TC));
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (sm::ld_u32<0>(fza[j]) == frb[j] && sm::ld_u32<M::ZI>(fza[j]) == (unsigned) (tid + j * TC)) {
        sm::st_f32x4<0>(4u * fza[j] - fk, fp[j]);
        sm::st_f32<M::FD>(fza[j], u2f(frb[j]));
      }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      sm::st_u32<0>(fza[j], Z_EMPTY_DEPTH);
      sm::st_u32<M::ZI>(fza[j], Z_EMPTY_IDX);
    }
  });
  __syncthreads();

  const unsigned wt = sb + M::WRED + (unsigned) warp * M::WTILE;
  float tot        = 0.f;  // warp X, lane s: total of slot s of pair X
  unsigned tot_cnt = 0;
  int its[2]       = {0, 0};  // iterations run per pair (uniform)
  bool alive[2]    = {true, n_here > 1};
  for (int it = 0; it < max_it && (alive[0] || alive[1]); ++it) {
    // phase 1: the four chains of the thread, interleaved
    {
      f2 pc[4];
      int col[4];
      bool near[4], up[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned bca = sb + M::BC + (q >> 1) * 48;
        const float Lc = sm::ld_f32<BC_LC>(bca), Ls = sm::ld_f32<BC_LS>(bca);
        const float Wtx = sm::ld_f32<BC_WTX>(bca), Wty = sm::ld_f32<BC_WTY>(bca);
        const f2 ra = mul2s(mk2(Lc, Ls), mp[q].x), rb2 = mul2s(mk2(-Ls, Lc), mp[q].y);
        pc[q]       = add2(mk2(fadd(ra.x, rb2.x), fadd(ra.y, rb2.y)), mk2(Wtx, Wty));
        const f2 pq = mul2(pc[q], pc[q]);
        const float rho = fsqrt(fadd(pq.x, pq.y));
        rb[q]       = f2u(rho);
        col[q]      = polar_column_fast2(P.cam, pc[q].y, pc[q].x, near[q], up[q]);
        near[q]     = near[q] && alive[q >> 1] && !(rho < P.range_min || rho > P.range_max);
      }
      if (near[0] || near[1] || near[2] || near[3]) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (near[q]) {
            bool undecided;
            const int c2 = polar_column_edge(P.cam, pc[q].y, pc[q].x, u2f(rb[q]), col[q] + (up[q] ? 1 : 0), undecided);
            col[q]       = undecided ? polar_column_exact(P.cam, pc[q].y, pc[q].x) : c2;
          }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned zb  = sb + M::Z + (q >> 1) * M::PAIR_Z;
        const unsigned bca = sb + M::BC + (q >> 1) * 48;
        const float rho = u2f(rb[q]);
        const bool ok   = alive[q >> 1] && !(rho < P.range_min || rho > P.range_max) && (unsigned) col[q] < (unsigned) C;
        za[q]           = zb + 4u * (ok ? col[q] : C);
        rb[q]           = ok ? rb[q] : 0u;
        if (ok && sm::atom_min_u32<0>(za[q], rb[q]) == rb[q]) sm::st_u32<BC_TIE>(bca, 1u);
      }
    }
    __syncthreads();
    const bool tie0 = sm::ld_u32<BC_TIE>(sb + M::BC) != 0, tie1 = sm::ld_u32<BC_TIE>(sb + M::BC + 48) != 0;
    if (tie0 || tie1) {  // exact pass on the pair(s) that saw equal minimal depths
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (((q >> 1) ? tie1 : tie0) && sm::ld_u32<0>(za[q]) == rb[q])
          sm::atom_min_u32<M::ZI>(za[q], (unsigned) (tid + (q & 1) * TC));
      __syncthreads();
    }
    // phase 2 + per-warp reduction, pair by pair
    static_for<0, 2>([&](auto xc) {
      constexpr int X = decltype(xc)::value;
      if (!alive[X]) return;
      const unsigned zb  = sb + M::Z + X * M::PAIR_Z;
      const unsigned fk  = 3u * zb - M::FI;
      const unsigned bca = sb + M::BC + X * 48;
      const unsigned mna = sb + X * M::PAIR_MN + 8u * tid;
      const bool tie     = X ? tie1 : tie0;
      const float Xtx = sm::ld_f32<BC_XTX>(bca), Xty = sm::ld_f32<BC_XTY>(bca);
      const float Lc = sm::ld_f32<BC_LC>(bca), Ls = sm::ld_f32<BC_LS>(bca);
      float Xc = 0.f, Xs = 0.f;
      if (SENSOR) Xc = sm::ld_f32<BC_XC>(bca), Xs = sm::ld_f32<BC_XS>(bca);
      float acc[NSUM];
#pragma unroll
      for (int s = 0; s < NSUM; ++s) acc[s] = 0.f;
      unsigned cnt = 0;
      static_for<0, 2>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        constexpr int Q = 2 * X + J;
        bool win        = sm::ld_u32<0>(za[Q]) == rb[Q];
        if (tie) win = win && sm::ld_u32<M::ZI>(za[Q]) == (unsigned) (tid + J * TC);
        if (win) {
          const float fd  = sm::ld_f32<M::FD>(za[Q]);
          const float4 F  = sm::ld_f32x4<0>(4u * za[Q] - fk);
          const float2 Mn = sm::ld_f32x2<J * TC * 8>(mna);
          linearize2<SENSOR, J == 0>(P, fd, F, mp[Q].x, mp[Q].y, Mn, u2f(rb[Q]), Xtx, Xty, Xc, Xs, Lc, Ls, acc, cnt);
        }
      });
      store_partials2(acc, cnt, wt, sb + M::RED + X * M::PAIR_RED + (unsigned) warp * (RED_STRIDE * 4), lane);
      __syncwarp();  // the warp's tile is reused by the next pair
    });
    __syncthreads();
    // hand-back, then the two Gauss-Newton steps side by side: warp X updates pair X
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sm::st_u32<0>(za[q], Z_EMPTY_DEPTH);
      if ((q >> 1) ? tie1 : tie0) sm::st_u32<M::ZI>(za[q], Z_EMPTY_IDX);
    }
    if (warp < 2 && alive[warp]) {
      pose_bc* bc      = reinterpret_cast<pose_bc*>(smem_raw + M::BC + warp * 48);
      const float* red = reinterpret_cast<const float*>(smem_raw + M::RED + warp * M::PAIR_RED);
      if (lane == 0) bc->tie = 0;
      warp0_update<TC, SENSOR, true>(P, A, bc, red, pair0 + warp, it, lane, tot, tot_cnt);
    }
    __syncthreads();
#pragma unroll
    for (int x = 0; x < 2; ++x)
      if (alive[x]) {
        its[x] = it + 1;
        if (sm::ld_u32<BC_STOP>(sb + M::BC + x * 48)) alive[x] = false, its[x] = it;
      }
  }
  if (warp < n_here) {
    const pose_bc* bc = reinterpret_cast<const pose_bc*>(smem_raw + M::BC + warp * 48);
    const int stop    = bc->stop;
    write_result(P, A, bc, pair0 + warp, its[warp], stop ? stop - 1 : -1, tot, tot_cnt, warp * 32);
  }
}

}  // namespace ls2d
