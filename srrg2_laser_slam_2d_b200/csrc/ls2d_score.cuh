// ls2d_score.cuh -- score_kernel: the single-linearisation pass (ls2d_score_batch; the loop-closure pre-filter "score every
// guess, align the best") as a PERSISTENT, TMA-fed kernel.  This is the regime of the path that is HBM-bound
// (I = 1 => ~3 flop/B): each pair's two clouds are read once, 80 B are written.
//
//  * grid = SMs x CTAs per SM; CTA b scores pairs b, b + grid, ...  Per pair there is no launch, no shared-memory
//    initialisation and no pose-update tail.
//  * every cloud travels global -> shared memory with ONE bulk-async copy (cp.async.bulk ... mbarrier::complete_tx::bytes;
//    SASS UBLKCP + SYNCS) behind an mbarrier of its own: the fixed cloud of pair k + 2 into the buffer pair k just
//    released, the moving cloud of pair k + 1 into the single moving buffer -- it lands while the fixed cloud of pair
//    k + 1 is projected.  Three buffers of up to 1152 points x 16 B keep the CTA at 70 KB: three CTAs = three pairs
//    in flight per SM, which is what the latency-bound projection chains need.
//  * both clouds are projected (identity camera for the fixed one, correspondence_finder_projective_2d.cpp:37-44; the
//    inverse of local_map_in_sensor for the moving one, .cpp:47-48) and fight for the columns of two z-buffers with
//    32-bit ATOMS.MIN on the rho bits (4 B per column and cloud: the fixed winner then overwrites the rho in its cell
//    with its index).  Equal minimal rho in a column (decision D3: lowest index wins) is detected from the atomic's
//    return value and only then resolved exactly, behind two extra barriers.  The thread that owns a winning moving
//    point gates and linearises its correspondence against the fixed winner of its column, read straight from the
//    staged cloud -- no fixed range image is materialised.
//  * TWO CTA barriers per pair, software-pipelined across pairs: [fixed winners of pair k put their index into their
//    cell + project moving(k)] | [winners of pair k linearise and hand their column's cells back + project fixed(k + 1)
//    into the other z-buffer].  The projections are straight-line code (squared-range gate, branch-free square root,
//    column proposal); the rare proposals only the exact path may decide are visited after the chains.  The partial-
//    sum rows are double-buffered by pair parity: warp 0 writes the result record while the other warps already work on
//    the next pair.
//  * reduction shape: thread t owns moving points t, t + T, ...; xor-butterfly per warp; warps in order -- what the
//    oracle's ORC_SUM_TREE mode mirrors (ls2d_score_reduction_shape()), so the records are bit-identical to the oracle.
#pragma once

#include "ls2d_icp2.cuh"

namespace ls2d {

namespace tma {
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy before the first bulk copy names them
__device__ __forceinline__ void fence_init() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void wait_parity(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
}  // namespace tma

// control blocks written by the producer thread before it arms the respective barrier
struct score_ctl_m {  // with the moving cloud of a pair
  pose_bc bc;         // pose of the pair (publish_pose)
  int nm, pair;
};
struct score_ctl_f {  // with the fixed cloud of a pair
  int nf, pad;
};

// shared-memory map (byte offsets; MP = point capacity of a staged cloud, CS = column stride >= cols + 1).  Three cloud
// buffers: the fixed clouds of pairs k and k + 1 alternate between two, the moving cloud of the current pair has one
// (its successor is fetched while the next pair's fixed cloud is projected).
struct score_map {
  int MP, CS, NW, SLOTS;
  __host__ __device__ score_map(int max_points, int cols, int threads, int ppt)
      : MP((max_points + 3) & ~3), CS((cols + 1 + 3) & ~3), NW(threads / 32), SLOTS(threads * ppt) {}
  __host__ __device__ int fixed(int b) const { return b * MP * 16; }
  __host__ __device__ int moving() const { return 2 * MP * 16; }
  __host__ __device__ int zdf(int b) const { return 3 * MP * 16 + b * 4 * CS; }  // u32[CS] per fixed buffer: rho bits, then
                                                                                 // the winner's index
  __host__ __device__ int zdm() const { return zdf(2); }                         // u32[CS]: rho bits of the moving cloud
  __host__ __device__ int red(int p) const { return zdm() + 4 * CS + p * NW * RED_STRIDE * 4; }
  __host__ __device__ int ctl_m(int p) const { return red(2) + p * (int) sizeof(score_ctl_m); }
  __host__ __device__ int ctl_f(int b) const { return ctl_m(2) + b * (int) sizeof(score_ctl_f); }
  __host__ __device__ int tie() const { return ctl_f(2); }                 // two ints: equal minimal rho among fixed / moving points
  __host__ __device__ int bar_f(int b) const { return tie() + 8 + b * 8; }  // mbarriers: two fixed buffers, one moving
  __host__ __device__ int bar_m() const { return bar_f(2); }
  // every thread reads its PPT point slots of a staged cloud unconditionally (slots past the cloud are masked, not
  // branched around): the allocation covers the furthest such read
  __host__ __device__ int bytes() const {
    const int reach = moving() + SLOTS * 16, end = bar_m() + 8;
    return end > reach ? end : reach;
  }
};
static_assert(sizeof(score_ctl_m) % 8 == 0 && sizeof(score_ctl_f) % 8 == 0, "the mbarriers stay 8-byte aligned");

// projection of one point already in the camera frame: squared-range gate, gated square root, fast column proposal
// (no branches: the rare proposals only the exact path may decide are flagged and visited afterwards)
struct score_eval {
  float x, y, rho;
  int col;
  bool in, near, up;
};
__device__ __forceinline__ score_eval project_fast(const dev_params& P, float x, float y) {
  score_eval e;
  e.x = x, e.y = y;
  const f2 qq   = mul2(mk2(x, y), mk2(x, y));
  const float a = fadd(qq.x, qq.y);
  e.in          = a >= P.gate2.lo && a <= P.gate2.hi;
  e.rho         = fsqrt_gated(a);
  e.col         = polar_column_fast2(P.cam, y, x, e.near, e.up);
  e.near        = e.near && e.in;
  return e;
}
// the rare proposals, out of line (ls2d_math.cuh: polar_column_resolve)
__device__ __forceinline__ void project_slow(const dev_params& P, score_eval& e) {
  e.col = polar_column_resolve(P.cam, e.y, e.x, e.rho, e.col, e.up);
}

template <int T, int PPT, bool SENSOR, bool FUSED, int MINB>
__global__ void __launch_bounds__(T, MINB) score_kernel(const dev_params P, const align_args A, int max_points) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const score_map M(max_points, P.cam.cols, T, PPT);
  const unsigned sb = sm::addr(smem_raw);
  const int C       = P.cam.cols;  // column C is the dummy cell of points that hit no column
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int n_mine = blockIdx.x < A.n_pairs ? (A.n_pairs - 1 - (int) blockIdx.x) / G + 1 : 0;
  const unsigned zdm = sb + M.zdm();
  const unsigned tie = sb + M.tie();
  const unsigned stM = sb + M.moving();

  // ---- one-time set-up: empty z-buffers, barriers, tie flags
  for (int k = tid; k < 3 * M.CS; k += T) sm::st_u32<0>(sb + M.zdf(0) + 4u * k, Z_EMPTY_DEPTH);  // zdf[0] | zdf[1] | zdm
  if (tid == 0) {
    tma::mbar_init(sb + M.bar_f(0), 1);
    tma::mbar_init(sb + M.bar_f(1), 1);
    tma::mbar_init(sb + M.bar_m(), 1);
    tma::fence_init();
    sm::st_u32<0>(tie, 0u), sm::st_u32<4>(tie, 0u);
  }
  __syncthreads();

  // producer (thread T - 1: warp 0 has the result tail to run): one bulk-async copy per cloud behind its barrier
  auto produce_fixed = [&](int k) {
    const int b    = k & 1;
    const int pair = blockIdx.x + k * G + A.pair_base;
    const int fcl  = A.fixed_const >= 0 ? A.fixed_const : (A.fixed_id ? A.fixed_id[pair] : pair);
    const int f0 = A.fixed_off[fcl], nf = min(A.fixed_off[fcl + 1] - f0, M.MP);
    reinterpret_cast<score_ctl_f*>(smem_raw + M.ctl_f(b))->nf = nf;
    const unsigned bar = sb + M.bar_f(b);
    tma::expect_tx(bar, (unsigned) nf * 16u);  // also releases the control block to the waiters
    if (nf) tma::bulk_g2s(sb + M.fixed(b), A.fixed_pts + f0, (unsigned) nf * 16u, bar);
  };
  auto produce_moving = [&](int k) {
    const int pair = blockIdx.x + k * G + A.pair_base;
    const int mcl  = A.moving_id ? A.moving_id[pair / A.moving_div] : pair / A.moving_div;
    const int m0 = A.moving_off[mcl], nm = min(A.moving_off[mcl + 1] - m0, M.MP);
    score_ctl_m* c = reinterpret_cast<score_ctl_m*>(smem_raw + M.ctl_m(k & 1));
    publish_pose(&c->bc, P, load_pose(A.init_pose, (size_t) pair, A.pose_stride), SENSOR, 0);
    c->nm = nm, c->pair = pair;
    const unsigned bar = sb + M.bar_m();
    tma::expect_tx(bar, (unsigned) nm * 16u);
    if (nm) tma::bulk_g2s(stM, A.moving_pts + m0, (unsigned) nm * 16u, bar);
  };
  if (tid == T - 1) {
    if (n_mine > 0) produce_fixed(0), produce_moving(0);
    if (n_mine > 1) produce_fixed(1);
  }

  // projection of the fixed cloud of pair k (identity camera) into its z-buffer: cell offsets and rho bits come back;
  // an equal minimal rho in a column (decision D3) raises the fixed tie flag
  unsigned cf[PPT], rf[PPT];  // cell offset (4 * column; dummy column C for rejected points) and rho bits
  auto project_fixed = [&](int k) {
    tma::wait_parity(sb + M.bar_f(k & 1), (unsigned) (k >> 1) & 1u);
    const unsigned stF = sb + M.fixed(k & 1), zdf = sb + M.zdf(k & 1);
    const int nf = reinterpret_cast<const score_ctl_f*>(smem_raw + M.ctl_f(k & 1))->nf;
    score_eval e[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int i    = tid + j * T;
      const float2 q = sm::ld_f32x2<0>(stF + 16u * i);  // slots past the cloud read stale bytes of the buffer: masked
      e[j]           = project_fast(P, q.x, q.y);
      e[j].in        = e[j].in && i < nf;
      e[j].near      = e[j].near && i < nf;
    }
    bool any_near = false;
#pragma unroll
    for (int j = 0; j < PPT; ++j) any_near |= e[j].near;
    if (any_near) {
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (e[j].near) project_slow(P, e[j]);
    }
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const bool ok = e[j].in && (unsigned) e[j].col < (unsigned) C;
      cf[j]         = 4u * (ok ? e[j].col : C);
      rf[j]         = ok ? f2u(e[j].rho) : 0u;
      if (ok && sm::atom_min_u32<0>(zdf + cf[j], rf[j]) == rf[j]) sm::st_u32<0>(tie, 1u);
    }
  };
  if (n_mine > 0) project_fixed(0);
  __syncthreads();

  for (int k = 0; k < n_mine; ++k) {
    const unsigned stF = sb + M.fixed(k & 1), zdf = sb + M.zdf(k & 1);
    const score_ctl_m* ctl = reinterpret_cast<const score_ctl_m*>(smem_raw + M.ctl_m(k & 1));

    // ---- interval A: the fixed winner of a column replaces the rho in its cell with its own index; the moving cloud
    // is projected (camera = local_map_in_sensor^-1; W = its double inverse, decision D13) and fights for zdm
    const bool tie_f = sm::ld_u32<0>(tie) != 0;  // uniform; rare: lowest index among equal rho, behind two barriers
    if (tie_f) {
      bool cand[PPT];
#pragma unroll
      for (int j = 0; j < PPT; ++j) cand[j] = sm::ld_u32<0>(zdf + cf[j]) == rf[j];
      __syncthreads();
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (cand[j]) sm::atom_min_u32<0>(zdf + cf[j], (unsigned) (tid + j * T));  // an index is below any rho's bits
      if (tid == 0) sm::st_u32<0>(tie, 0u);
      __syncthreads();
    }
    tma::wait_parity(sb + M.bar_m(), (unsigned) k & 1u);
    const int nm = ctl->nm;
    unsigned cm[PPT], rm[PPT];
    {
      const float Lc = ctl->bc.Lc, Ls = ctl->bc.Ls, Wtx = ctl->bc.Wtx, Wty = ctl->bc.Wty;
      score_eval e[PPT];
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const int i    = tid + j * T;
        const float2 m = sm::ld_f32x2<0>(stM + 16u * i);
        const f2 ra = mul2s(mk2(Lc, Ls), m.x), rb = mul2s(mk2(-Ls, Lc), m.y);
        const f2 pc = add2(mk2(fadd(ra.x, rb.x), fadd(ra.y, rb.y)), mk2(Wtx, Wty));
        e[j]        = project_fast(P, pc.x, pc.y);
        e[j].in     = e[j].in && i < nm;
        e[j].near   = e[j].near && i < nm;
        if (!tie_f && sm::ld_u32<0>(zdf + cf[j]) == rf[j]) sm::st_u32<0>(zdf + cf[j], (unsigned) i);  // the one winner
      }
      bool any_near = false;
#pragma unroll
      for (int j = 0; j < PPT; ++j) any_near |= e[j].near;
      if (any_near) {
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (e[j].near) project_slow(P, e[j]);
      }
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const bool ok = e[j].in && (unsigned) e[j].col < (unsigned) C;
        cm[j]         = 4u * (ok ? e[j].col : C);
        rm[j]         = ok ? f2u(e[j].rho) : 0u;
        if (ok && sm::atom_min_u32<0>(zdm + cm[j], rm[j]) == rm[j]) sm::st_u32<4>(tie, 1u);  // an equal rho was there
      }
    }
    __syncthreads();

    // ---- interval B: the owner of a winning moving point gates and linearises its correspondence (.cpp:61-73) and hands
    // the column's cells back (a column has one winner; losers only ever see its rho or EMPTY); the next pair's fixed
    // cloud is projected into the other z-buffer in the same interval
    const bool tie_m = sm::ld_u32<4>(tie) != 0;  // uniform; rare
    bool win[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) win[j] = sm::ld_u32<0>(zdm + cm[j]) == rm[j];
    if (tie_m) {  // lowest index among equal rho: the candidates put their index into the cell (below any rho's bits)
      __syncthreads();
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (win[j]) sm::atom_min_u32<0>(zdm + cm[j], (unsigned) (tid + j * T));
      if (tid == 0) sm::st_u32<4>(tie, 0u);
      __syncthreads();
#pragma unroll
      for (int j = 0; j < PPT; ++j) win[j] = win[j] && sm::ld_u32<0>(zdm + cm[j]) == (unsigned) (tid + j * T);
    }
    const float Xtx = ctl->bc.Xtx, Xty = ctl->bc.Xty, Lc = ctl->bc.Lc, Ls = ctl->bc.Ls;
    float Xc = 0.f, Xs = 0.f;
    if (SENSOR) Xc = ctl->bc.Xc, Xs = ctl->bc.Xs;
    float acc[NSUM];
#pragma unroll
    for (int q = 0; q < NSUM; ++q) acc[q] = 0.f;
    unsigned cnt = 0;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (win[j]) {
        const unsigned fi = sm::ld_u32<0>(zdf + cm[j]);  // index of the column's fixed winner, or EMPTY
        if (fi != Z_EMPTY_DEPTH) {
          const float4 F  = sm::ld_f32x4<0>(stF + 16u * fi);
          const float4 Mv = sm::ld_f32x4<0>(stM + 16u * (unsigned) (tid + j * T));
          const f2 ff     = mul2(mk2(F.x, F.y), mk2(F.x, F.y));
          const float fd  = fsqrt_gated(fadd(ff.x, ff.y));  // the fixed cell's depth: the same operations as its projection
          if (FUSED)
            linearize2f<SENSOR, false>(P, fd, F, Mv.x, Mv.y, make_float2(Mv.z, Mv.w), u2f(rm[j]), Xtx, Xty, Xc, Xs, Lc, Ls,
                                       acc, cnt);
          else
            linearize2<SENSOR, false>(P, fd, F, Mv.x, Mv.y, make_float2(Mv.z, Mv.w), u2f(rm[j]), Xtx, Xty, Xc, Xs, Lc, Ls,
                                      acc, cnt);
          sm::st_u32<0>(zdf + cm[j], Z_EMPTY_DEPTH);
        }
        sm::st_u32<0>(zdm + cm[j], Z_EMPTY_DEPTH);
      }
    }
    // a fixed winner whose column no moving point reached hands its own cell back (nobody reads it); tie losers of
    // the moving cloud left their cell to the winner above
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (sm::ld_u32<0>(zdf + cf[j]) == (unsigned) (tid + j * T) && sm::ld_u32<0>(zdm + cf[j]) == Z_EMPTY_DEPTH)
        sm::st_u32<0>(zdf + cf[j], Z_EMPTY_DEPTH);
    if (k + 1 < n_mine) project_fixed(k + 1);  // overwrites cf / rf: pair k is done with them
    {
      float v16[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v16[q] = q < NSUM ? acc[q] : 0.f;
      store_partials(v16, cnt, reinterpret_cast<float*>(smem_raw + M.red(k & 1)), lane, warp);
    }
    const int pair = ctl->pair;
    const float rx = ctl->bc.Xtx, ry = ctl->bc.Xty, rc = ctl->bc.Xc, rs = ctl->bc.Xs;
    __syncthreads();

    // ---- the moving buffer and this pair's fixed buffer are free: next copies; warp 0 writes the record
    if (tid == T - 1) {
      if (k + 1 < n_mine) produce_moving(k + 1);
      if (k + 2 < n_mine) produce_fixed(k + 2);
    }
    if (warp == 0) {
      const float* red = reinterpret_cast<const float*>(smem_raw + M.red(k & 1));
      float tot        = 0.f;
      unsigned tot_cnt = 0;
      if (lane < NSUM) {
        tot = red[lane];
#pragma unroll
        for (int w = 1; w < T / 32; ++w) tot = fadd(tot, red[w * RED_STRIDE + lane]);
        tot = fadd(tot, 0.f);
      } else if (lane == NSUM) {
#pragma unroll
        for (int w = 0; w < T / 32; ++w) tot_cnt += __float_as_uint(red[w * RED_STRIDE + NSUM]);
      }
      float v[NSUM];
#pragma unroll
      for (int q = 0; q < NSUM; ++q) v[q] = __shfl_sync(0xffffffffu, tot, q);
      const unsigned c2 = __shfl_sync(0xffffffffu, tot_cnt, NSUM);
      if (lane == 0) {
        int n_in = c2 & 0xffff, n_k = c2 >> 16;
        const int n_corr = n_in + n_k;
        int status, it = 1;
        if (n_corr <= P.min_num_correspondences) {  // the oracle reports empty sums and no completed round here
          status = LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES, it = 0;
#pragma unroll
          for (int q = 0; q < NSUM; ++q) v[q] = 0.f;
          n_in = n_k = 0;
        } else {
          status = n_in < P.min_num_inliers ? LS2D_STATUS_NOT_ENOUGH_INLIERS : LS2D_STATUS_SUCCESS;
        }
        ls2d_result r;
        r.x = rx, r.y = ry, r.theta = atan2f_fdlibm(rs, rc);
        r.chi_inliers = v[9], r.chi_kernelized = v[10];
        r.n_inliers = n_in, r.n_kernelized = n_k, r.n_corr = n_corr;
        r.status = status, r.iterations = it;
#pragma unroll
        for (int q = 0; q < 6; ++q) r.H[q] = v[q];
        r.c = rc, r.s = rs;
        r.lm_rejected = 0, r.reserved = 0;
        A.out[pair] = r;
      }
    }
  }
}

}  // namespace ls2d
