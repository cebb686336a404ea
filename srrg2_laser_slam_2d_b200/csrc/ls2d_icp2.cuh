// ls2d_icp2.cuh -- icp_fused2_kernel: the instruction-diet version of icp_fused_kernel (same algorithm, same
// decisions; see ls2d_icp.cuh and ls2d_common.cuh for the algorithm and the reference citations).
//
// What changed, all of it aimed at the issue slots the ncu source view showed being spent on bookkeeping
// (profiles/r01_icp_ncu_summary.md: phase 2 + reduction = 65 % of the executed warp instructions):
//   * shared memory is addressed with explicit 32-bit shared-space addresses (ld/st/atom.shared with immediate
//     offsets).  Every per-column array has the compile-time stride CS, so a point keeps ONE register -- the address
//     of its z-buffer cell -- and reaches the cell's depth / index / fixed rho with immediates and the fixed point
//     with one shift-add; the generic-pointer version recomputed the shared window base and four array bases per
//     point.
//   * invalid points (outside the range gates / the canvas / past the end of the cloud) are parked on a dummy
//     column that nobody ever wins, so neither phase 2 nor the hand-back pass tests validity.
//   * ties (two points of equal minimal rho in one column, decision D3) are detected where they happen: the
//     z-buffer atomicMin returns the previous value, and "previous == mine" means an equal rho was already there.
//     Only then the iteration runs the exact lowest-index pass; the per-winner compare-and-swap claim is gone.
//   * the per-warp reduction of the 11 sums goes through a transposed shared-memory tile (11 conflict-free stores,
//     4 x 128-bit loads and 15 adds on 22 lanes, one shuffle) instead of the 16-shuffle recursive-halving tree with
//     its ~30 selects.  This is a different -- equally fixed -- summation shape: lanes 0..15 and 16..31 of a warp are
//     summed in ascending order, the two halves added, the warps added in ascending order; the oracle offers it as
//     ORC_SUM_TREE with bit 16 of tree_threads set (ls2d_reduction_shape() reports it).
//   * the first of a thread's points assigns its contribution instead of adding it to zero (the totals are
//     canonicalised with +0.0f in warp 0, so even an all-minus-zero sum matches the oracle's 0 + x).
//   * the column of a moving point is decided in three tiers (ls2d_math.cuh): fast proposal, side of the rounding
//     edge's ray in binary64, exact atan2f -- the single-lane exact path in front of the CTA barrier is 3 x rarer.
//   * FUSED (default at the 1152-column stride): the error / Jacobian entries of an accepted correspondence and their
//     accumulation use fused multiply-adds (linearize2f, oracle decision D18, ORC_SUM_TREE bit 17); every gate and
//     everything that decides a pixel index stays single-rounding arithmetic.
//   * lanes past the end of the cloud hold a far point (1e30, 0): its rho overflows to inf, no range gate accepts it,
//     and the iteration loop never compares point indices.
#pragma once

#include <type_traits>

#include "ls2d_icp.cuh"

namespace ls2d {

namespace sm {
__device__ __forceinline__ unsigned addr(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
template <int OFF>
__device__ __forceinline__ unsigned ld_u32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ float ld_f32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ float2 ld_f32x2(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ float4 ld_f32x4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(a), "n"(OFF)
               : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ void st_u32(unsigned a, unsigned v) {
  asm volatile("st.shared.u32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void st_f32(unsigned a, float v) {
  asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "f"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void st_f32x2(unsigned a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "f"(x), "f"(y) : "memory");
}
template <int OFF>
__device__ __forceinline__ void st_f32x4(unsigned a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
template <int OFF>
__device__ __forceinline__ unsigned atom_min_u32(unsigned a, unsigned v) {
  unsigned old;
  asm volatile("atom.shared.min.u32 %0, [%1+%2], %3;" : "=r"(old) : "r"(a), "n"(OFF), "r"(v) : "memory");
  return old;
}
}  // namespace sm

// compile-time loop: f(std::integral_constant<int, J>) for J = 0 .. N-1 (immediate offsets need constant J)
template <int J, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (J < N) {
    f(std::integral_constant<int, J>{});
    static_for<J + 1, N>(f);
  }
}

// shared-memory map of icp_fused2_kernel (byte offsets from the start of dynamic shared memory)
template <int T, int PPT, int CS>
struct icp2_map {
  static constexpr int NW        = T / 32;
  static constexpr int WTILE_ROW = 36 * 4;            // one slot's 32 lane values + 4 floats of padding
  static constexpr int WTILE     = NSUM * WTILE_ROW;  // per-warp transposed tile: 1584 B
  static constexpr int MNRM      = 0;                           // float2[T * PPT] moving normals
  static constexpr int WRED      = MNRM + T * PPT * 8;          // NW tiles
  static constexpr int RED       = WRED + NW * WTILE;           // float[NW][RED_STRIDE] warp totals
  static constexpr int BC        = RED + NW * RED_STRIDE * 4;   // pose_bc
  static constexpr int Z         = (BC + (int) sizeof(pose_bc) + 15) & ~15;  // u32[CS] z-buffer depth (rho bits)
  static constexpr int ZI        = 4 * CS;                      // from Z: u32[CS] z-buffer index (exact passes)
  static constexpr int FD        = 8 * CS;                      // from Z: float[CS] fixed image rho, < 0 = empty
  static constexpr int FI        = 12 * CS;                     // from Z: float4[CS] fixed image point
  static constexpr int BYTES     = Z + 28 * CS;
  static_assert(CS % 4 == 0, "column stride keeps the float4 image 16-byte aligned");
};

constexpr int BC_XTX = 0, BC_XTY = 4, BC_XC = 8, BC_XS = 12, BC_LC = 16, BC_LS = 20, BC_WTX = 24, BC_WTY = 28,
              BC_STOP = 32, BC_TIE = 36;
static_assert(sizeof(pose_bc) == 40, "pose_bc layout is addressed by byte offsets");

// One winner against its fixed cell: gates of CorrespondenceFinderProjective2f (.cpp:61-73), SE2Plane2PlaneErrorFactor,
// Cauchy, H/b terms -- operation for operation the arithmetic of linearize_point() (ls2d_common.cuh).  FIRST: the
// thread's sums are still zero, assign instead of add.
template <bool SENSOR, bool FIRST, bool P2P = false>
__device__ __forceinline__ void linearize2(const dev_params& P, float fd, const float4 F, float Mx, float My,
                                           float2 Mn, float rho, float Xtx, float Xty, float Xc, float Xs, float Lc,
                                           float Ls, float (&acc)[NSUM], unsigned& cnt) {
  if (fd < 0.f || fabsf(fsub(fd, rho)) > P.point_distance) return;
  const f2 rc1 = mk2(Lc, Ls), rc2 = mk2(-Ls, Lc);  // columns of R(local_map_in_sensor)
  const f2 na = mul2s(rc1, Mn.x), nb = mul2s(rc2, Mn.y);
  const float nx = fadd(na.x, nb.x);  // transformed normal
  const float ny = fadd(na.y, nb.y);
  const f2 fn = mk2(F.z, F.w);
  const f2 nd = mul2(mk2(nx, ny), fn);
  if (fadd(nd.x, nd.y) < P.normal_cos) return;
  f2 p;
  if (SENSOR) {
    const f2 qa = mul2s(mk2(Xc, Xs), Mx), qb = mul2s(mk2(-Xs, Xc), My);
    const float qx = fadd(fadd(qa.x, qb.x), Xtx);
    const float qy = fadd(fadd(qa.y, qb.y), Xty);
    iso_apply(P.Sinv, qx, qy, p.x, p.y);
  } else {
    const f2 pa = mul2s(rc1, Mx), pb = mul2s(rc2, My);
    p = add2(mk2(fadd(pa.x, pb.x), fadd(pa.y, pb.y)), mk2(Xtx, Xty));
  }
  const f2 d  = add2(p, mk2(-F.x, -F.y));
  if (P2P) {
    // SE2Point2PointErrorFactor[WithSensor] (oracle decision D19): e = p - p_fixed, J = [R | R (-y, x)^T], Omega = I2;
    // operation for operation the point-to-point arm of linearize_point() (ls2d_common.cuh)
    const float jc0 = fadd(fmul(Lc, -My), fmul(-Ls, Mx));
    const float jc1 = fadd(fmul(Ls, -My), fmul(Lc, Mx));
    const float chi = fadd(fmul(d.x, d.x), fmul(d.y, d.y));
    float w = 1.f, chi_in = chi, chi_k = 0.f;
    if (P.tau > 0.f && !(chi < P.tau)) {  // RobustifierCauchy
      const float aux = fadd(fmul(chi, P.inv_tau), 1.f);
      chi_k           = fmul(P.tau, __logf(aux));  // statistics only (tolerance parity)
      w               = frcp(aux);
      chi_in          = 0.f;
      cnt += 1u << 16;
    } else {
      cnt += 1u;
    }
    const float wc = fmul(Lc, w), ws = fmul(Ls, w), wms = fmul(-Ls, w), wj0 = fmul(jc0, w), wj1 = fmul(jc1, w);
    const float t[9] = {fadd(fmul(wc, Lc), fmul(ws, Ls)),    fadd(fmul(wc, -Ls), fmul(ws, Lc)),
                        fadd(fmul(wc, jc0), fmul(ws, jc1)),  fadd(fmul(wms, -Ls), fmul(wc, Lc)),
                        fadd(fmul(wms, jc0), fmul(wc, jc1)), fadd(fmul(wj0, jc0), fmul(wj1, jc1)),
                        fadd(fmul(wc, d.x), fmul(ws, d.y)),  fadd(fmul(wms, d.x), fmul(wc, d.y)),
                        fadd(fmul(wj0, d.x), fmul(wj1, d.y))};
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = FIRST ? t[q] : fadd(acc[q], t[q]);
    acc[9]  = FIRST ? chi_in : fadd(acc[9], chi_in);
    acc[10] = FIRST ? chi_k : fadd(acc[10], chi_k);
    return;
  }
  const f2 de = mul2(d, fn);
  const float e0 = fadd(de.x, de.y);
  const f2 e12 = add2(mk2(nx, ny), mk2(-F.z, -F.w));  // e1, e2
  const f2 ja = mul2s(mk2(Lc, -Ls), F.z), jb = mul2s(mk2(Ls, Lc), F.w);
  const float Ja = fadd(ja.x, jb.x);
  const float Jb = fadd(ja.y, jb.y);
  const f2 jc = mul2(mk2(Ja, Jb), mk2(-My, Mx));
  const float Jc = fadd(jc.x, jc.y);
  const float d0 = -ny, d1 = nx;  // R * (-n.y, n.x)^T, exact in binary32
  const f2 ee = mul2(e12, e12);
  const float chi = fadd(fadd(fmul(e0, e0), ee.x), ee.y);
  float w = 1.f, chi_in = chi, chi_k = 0.f;
  if (P.tau > 0.f && !(chi < P.tau)) {  // RobustifierCauchy (L0.json:76-81)
    const float aux = fadd(fmul(chi, P.inv_tau), 1.f);
    chi_k           = fmul(P.tau, __logf(aux));  // statistics only (tolerance parity)
    w               = frcp(aux);
    chi_in          = 0.f;
    cnt += 1u << 16;
  } else {
    cnt += 1u;
  }
  const f2 wab = mul2s(mk2(Ja, Jb), w);  // wa, wb
  const float wc = fmul(Jc, w);
  const f2 wd = mul2s(mk2(d0, d1), w);   // wd0, wd1
  const f2 h01 = mul2s(mk2(Ja, Jb), wab.x);            // wa*Ja, wa*Jb
  const f2 h23 = mul2(wab, mk2(Jc, Jb));               // wa*Jc, wb*Jb
  const f2 h4c = mul2(mk2(wab.y, wc), mk2(Jc, Jc));    // wb*Jc, wc*Jc
  const f2 hdd = mul2(wd, mk2(d0, d1));                // wd0*d0, wd1*d1
  const f2 b01 = mul2s(wab, e0);                       // wa*e0, wb*e0
  const f2 bde = mul2(wd, e12);                        // wd0*e1, wd1*e2
  const float t5 = fadd(fadd(h4c.y, hdd.x), hdd.y);
  const float t8 = fadd(fadd(fmul(wc, e0), bde.x), bde.y);
  if (FIRST) {
    acc[0] = h01.x, acc[1] = h01.y, acc[2] = h23.x, acc[3] = h23.y, acc[4] = h4c.x, acc[5] = t5;
    acc[6] = b01.x, acc[7] = b01.y, acc[8] = t8, acc[9] = chi_in, acc[10] = chi_k;
  } else {
    acc[0]  = fadd(acc[0], h01.x);
    acc[1]  = fadd(acc[1], h01.y);
    acc[2]  = fadd(acc[2], h23.x);
    acc[3]  = fadd(acc[3], h23.y);
    acc[4]  = fadd(acc[4], h4c.x);
    acc[5]  = fadd(acc[5], t5);
    acc[6]  = fadd(acc[6], b01.x);
    acc[7]  = fadd(acc[7], b01.y);
    acc[8]  = fadd(acc[8], t8);
    acc[9]  = fadd(acc[9], chi_in);
    acc[10] = fadd(acc[10], chi_k);
  }
}

// The same winner with FUSED accumulation arithmetic (decision D18 of oracle/ls2d_oracle.c, ORC_SUM_TREE bit 17): the
// gates -- everything that decides a correspondence -- are the single-rounding operations of linearize2(); the error /
// Jacobian entries and the sums use fused multiply-adds in exactly the oracle's association.  ~15 instructions fewer
// per winner.
template <bool SENSOR, bool FIRST>
__device__ __forceinline__ void linearize2f(const dev_params& P, float fd, const float4 F, float Mx, float My,
                                            float2 Mn, float rho, float Xtx, float Xty, float Xc, float Xs, float Lc,
                                            float Ls, float (&acc)[NSUM], unsigned& cnt) {
  if (fd < 0.f || fabsf(fsub(fd, rho)) > P.point_distance) return;
  const f2 rc1 = mk2(Lc, Ls), rc2 = mk2(-Ls, Lc);  // columns of R(local_map_in_sensor)
  const f2 na = mul2s(rc1, Mn.x), nb = mul2s(rc2, Mn.y);
  const float nx = fadd(na.x, nb.x);  // transformed normal (unfused: the gate below reads it)
  const float ny = fadd(na.y, nb.y);
  const f2 nd = mul2(mk2(nx, ny), mk2(F.z, F.w));
  if (fadd(nd.x, nd.y) < P.normal_cos) return;
  f2 p;
  if (SENSOR) {
    const f2 qa = mul2s(mk2(Xc, Xs), Mx), qb = mul2s(mk2(-Xs, Xc), My);
    const float qx = fadd(fadd(qa.x, qb.x), Xtx);
    const float qy = fadd(fadd(qa.y, qb.y), Xty);
    iso_apply(P.Sinv, qx, qy, p.x, p.y);
  } else {  // px = fma(Lc, Mx, fma(-Ls, My, Xtx)), py = fma(Ls, Mx, fma(Lc, My, Xty))
    p = fma2(rc1, mk2(Mx, Mx), fma2(rc2, mk2(My, My), mk2(Xtx, Xty)));
  }
  const f2 d     = add2(p, mk2(-F.x, -F.y));
  const float e0 = ffma(d.x, F.z, fmul(d.y, F.w));
  const f2 e12   = add2(mk2(nx, ny), mk2(-F.z, -F.w));  // e1, e2
  // Ja = fma(F.z, Lc, F.w * Ls), Jb = fma(F.w, Lc, F.z * -Ls)
  const f2 Jab   = fma2(mk2(F.z, F.w), mk2(Lc, Lc), mul2(mk2(F.w, F.z), mk2(Ls, -Ls)));
  const float Ja = Jab.x, Jb = Jab.y;
  const float Jc = ffma(Ja, -My, fmul(Jb, Mx));
  const float d0 = -ny, d1 = nx;
  const float chi = ffma(e12.y, e12.y, ffma(e12.x, e12.x, fmul(e0, e0)));
  float w = 1.f, chi_in = chi, chi_k = 0.f;
  if (P.tau > 0.f && !(chi < P.tau)) {  // RobustifierCauchy (L0.json:76-81)
    const float aux = fadd(fmul(chi, P.inv_tau), 1.f);
    chi_k           = fmul(P.tau, __logf(aux));  // statistics only (tolerance parity)
    w               = frcp(aux);
    chi_in          = 0.f;
    cnt += 1u << 16;
  } else {
    cnt += 1u;
  }
  const f2 wab   = mul2s(Jab, w);  // wa, wb
  const float wc = fmul(Jc, w);
  const f2 wd    = mul2s(mk2(d0, d1), w);  // wd0, wd1
  if (FIRST) {
    const f2 a01 = mul2s(Jab, wab.x);                      // wa Ja, wa Jb
    const f2 a23 = mul2(wab, mk2(Jc, Jb));                 // wa Jc, wb Jb
    const f2 a67 = mul2s(wab, e0);                         // wa e0, wb e0
    f2 a58       = mul2s(mk2(Jc, e0), wc);                 // wc Jc, wc e0
    a58          = fma2(mk2(wd.x, wd.x), mk2(d0, e12.x), a58);
    a58          = fma2(mk2(wd.y, wd.y), mk2(d1, e12.y), a58);
    acc[0] = a01.x, acc[1] = a01.y, acc[2] = a23.x, acc[3] = a23.y, acc[4] = fmul(wab.y, Jc), acc[5] = a58.x;
    acc[6] = a67.x, acc[7] = a67.y, acc[8] = a58.y, acc[9] = chi_in, acc[10] = chi_k;
  } else {
    const f2 a01 = fma2(mk2(wab.x, wab.x), Jab, mk2(acc[0], acc[1]));
    const f2 a23 = fma2(wab, mk2(Jc, Jb), mk2(acc[2], acc[3]));
    const f2 a67 = fma2(wab, mk2(e0, e0), mk2(acc[6], acc[7]));
    f2 a58       = fma2(mk2(wc, wc), mk2(Jc, e0), mk2(acc[5], acc[8]));
    a58          = fma2(mk2(wd.x, wd.x), mk2(d0, e12.x), a58);
    a58          = fma2(mk2(wd.y, wd.y), mk2(d1, e12.y), a58);
    acc[0] = a01.x, acc[1] = a01.y, acc[2] = a23.x, acc[3] = a23.y, acc[5] = a58.x, acc[8] = a58.y;
    acc[6] = a67.x, acc[7] = a67.y;
    acc[4]  = ffma(wab.y, Jc, acc[4]);
    acc[9]  = fadd(acc[9], chi_in);
    acc[10] = fadd(acc[10], chi_k);
  }
}

// per-warp reduction through the warp's transposed tile; the total of slot s ends on lane 2s and goes to the warp's
// row of `red` (same row format as store_partials()).  wt = shared address of the warp's tile, rrow = of its row.
__device__ __forceinline__ void store_partials2(const float (&acc)[NSUM], unsigned cnt, unsigned wt, unsigned rrow,
                                                int lane) {
  constexpr int ROW = 36 * 4;
  const unsigned wl = wt + 4u * lane;
  sm::st_f32<0 * ROW>(wl, acc[0]);
  sm::st_f32<1 * ROW>(wl, acc[1]);
  sm::st_f32<2 * ROW>(wl, acc[2]);
  sm::st_f32<3 * ROW>(wl, acc[3]);
  sm::st_f32<4 * ROW>(wl, acc[4]);
  sm::st_f32<5 * ROW>(wl, acc[5]);
  sm::st_f32<6 * ROW>(wl, acc[6]);
  sm::st_f32<7 * ROW>(wl, acc[7]);
  sm::st_f32<8 * ROW>(wl, acc[8]);
  sm::st_f32<9 * ROW>(wl, acc[9]);
  sm::st_f32<10 * ROW>(wl, acc[10]);
  const unsigned wcnt = __reduce_add_sync(0xffffffffu, cnt);
  __syncwarp();
  float v = 0.f;
  if (lane < 2 * NSUM) {
    // lane = 2 * slot + half: 16 consecutive lane values of the slot, summed in ascending lane order
    const unsigned src = wt + (unsigned) (lane >> 1) * ROW + (unsigned) (lane & 1) * 64u;
    const float4 a = sm::ld_f32x4<0>(src), b = sm::ld_f32x4<16>(src), c = sm::ld_f32x4<32>(src),
                 d = sm::ld_f32x4<48>(src);
    v = fadd(a.x, a.y);
    v = fadd(v, a.z), v = fadd(v, a.w);
    v = fadd(v, b.x), v = fadd(v, b.y), v = fadd(v, b.z), v = fadd(v, b.w);
    v = fadd(v, c.x), v = fadd(v, c.y), v = fadd(v, c.z), v = fadd(v, c.w);
    v = fadd(v, d.x), v = fadd(v, d.y), v = fadd(v, d.z), v = fadd(v, d.w);
  }
  const float o = __shfl_xor_sync(0xffffffffu, v, 1);
  if (lane < 2 * NSUM && !(lane & 1)) sm::st_f32<0>(rrow + 2u * lane, fadd(v, o));  // lower half + upper half
  if (lane == 0) sm::st_u32<NSUM * 4>(rrow, wcnt);
}

template <int T, int PPT, bool SENSOR, int MINB, int CS, bool FUSED = true, bool P2P = false>
__global__ void __launch_bounds__(T, MINB) icp_fused2_kernel(const dev_params P, const align_args A) {
  static_assert(!(FUSED && P2P), "the fused accumulation arithmetic (D18) is the plane-to-plane factor's");
  using M = icp2_map<T, PPT, CS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned sb = sm::addr(smem_raw);
  const unsigned zb = sb + M::Z;        // z-buffer base; cell of column c = zb + 4 c
  const unsigned fk = 3u * zb - M::FI;  // fixed point of the column whose cell is za: 4 za - fk = zb + FI + 16 c
  float* red        = reinterpret_cast<float*>(smem_raw + M::RED);
  pose_bc* bc       = reinterpret_cast<pose_bc*>(smem_raw + M::BC);
  const unsigned bca = sb + M::BC;

  const int C   = P.cam.cols;  // < CS: column C is the dummy cell of invalid points
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = blockIdx.x + A.pair_base;
  const int fcl  = A.fixed_const >= 0 ? A.fixed_const : (A.fixed_id ? A.fixed_id[pair] : pair);
  const int mcl  = A.moving_id ? A.moving_id[pair / A.moving_div] : pair / A.moving_div;
  const int f0 = A.fixed_off[fcl], nf = A.fixed_off[fcl + 1] - f0;
  const int m0 = A.moving_off[mcl], nm = A.moving_off[mcl + 1] - m0;

  for (int k = tid; k <= C; k += T) {
    const unsigned a = zb + 4u * k;
    sm::st_f32<M::FD>(a, -1.f);
    sm::st_u32<0>(a, Z_EMPTY_DEPTH);
    sm::st_u32<M::ZI>(a, Z_EMPTY_IDX);
  }
  // moving cloud (issued early; consumed after the fixed image is built): coordinates -> registers for all
  // iterations, normals -> shared memory (only winners read them)
  const unsigned mna = sb + M::MNRM + 8u * tid;
  float2 mp[PPT];
  static_for<0, PPT>([&](auto jc) {
    constexpr int J = decltype(jc)::value;
    const int i     = tid + J * T;
    // lanes past the end of the cloud hold a point no pose brings inside the range gates (rho overflows to inf)
    const float4 m  = i < nm ? ldg4_once(A.moving_pts + m0 + i) : make_float4(1e30f, 0.f, 0.f, 0.f);
    mp[J]           = make_float2(m.x, m.y);
    sm::st_f32x2<J * T * 8>(mna, m.z, m.w);
  });
  if (tid == 0) {
    const iso X = load_pose(A.init_pose, (size_t) pair, A.pose_stride);
    publish_pose(bc, P, X, SENSOR, 0);
    bc->tie = 0;
  }
  __syncthreads();

  // ---- fixed range image: identity camera (R/registration/correspondence_finder_projective_2d.cpp:37-44), exact
  // two-pass z-buffer
  {
    float4 fp[PPT];
    unsigned za[PPT];
    unsigned rb[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int i = tid + j * T;
      int col     = C;
      rb[j]       = 0;
      if (i < nf) {
        fp[j]           = ldg4_once(A.fixed_pts + f0 + i);
        const float rho = fsqrt(fadd(fmul(fp[j].x, fp[j].x), fmul(fp[j].y, fp[j].y)));
        if (!(rho < P.range_min || rho > P.range_max)) {
          const int c = polar_column(P.cam, fp[j].y, fp[j].x);
          if (c >= 0) col = c, rb[j] = f2u(rho);
        }
      }
      za[j] = zb + 4u * col;
      if (col != C) sm::atom_min_u32<0>(za[j], rb[j]);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (sm::ld_u32<0>(za[j]) == rb[j]) sm::atom_min_u32<M::ZI>(za[j], (unsigned) (tid + j * T));
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (sm::ld_u32<0>(za[j]) == rb[j] && sm::ld_u32<M::ZI>(za[j]) == (unsigned) (tid + j * T)) {
        sm::st_f32x4<0>(4u * za[j] - fk, fp[j]);  // = zb + FI + 16 c
        sm::st_f32<M::FD>(za[j], u2f(rb[j]));
      }
    __syncthreads();
    // hand the z-buffer back empty for the moving cloud (the dummy cell included: it only ever holds EMPTY)
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      sm::st_u32<0>(za[j], Z_EMPTY_DEPTH);
      sm::st_u32<M::ZI>(za[j], Z_EMPTY_IDX);
    }
    __syncthreads();
  }

  // ---- ICP loop (MultiAligner2D::compute; L0.json:487-517)
  const int max_it = A.score_only ? 1 : P.max_iterations;
  int it           = 0;
  int status       = -1;
  float tot        = 0.f;  // lane s of warp 0: total of slot s for the last linearisation
  unsigned tot_cnt = 0;
  const unsigned wt   = sb + M::WRED + (unsigned) warp * M::WTILE;
  const unsigned rrow = sb + M::RED + (unsigned) warp * (RED_STRIDE * 4);
  for (; it < max_it; ++it) {
    unsigned za[PPT];
    unsigned rb[PPT];
    // phase 1: project the moving cloud (camera = local_map_in_sensor^-1, .cpp:47-48) and fight for the column
    {
      const float Lc = sm::ld_f32<BC_LC>(bca), Ls = sm::ld_f32<BC_LS>(bca);
      const float Wtx = sm::ld_f32<BC_WTX>(bca), Wty = sm::ld_f32<BC_WTY>(bca);
      f2 pc[PPT];
      int col[PPT];
      bool near[PPT], up[PPT], in[PPT];
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const f2 ra = mul2s(mk2(Lc, Ls), mp[j].x), rb2 = mul2s(mk2(-Ls, Lc), mp[j].y);
        pc[j]       = add2(mk2(fadd(ra.x, rb2.x), fadd(ra.y, rb2.y)), mk2(Wtx, Wty));
        const f2 pq = mul2(pc[j], pc[j]);
        const float a = fadd(pq.x, pq.y);
        in[j]       = a >= P.gate2.lo && a <= P.gate2.hi;  // the range gate, taken on the squared range (ls2d_math.cuh)
        rb[j]       = f2u(fsqrt_gated(a));                 // exact wherever the gate is open
        col[j]      = polar_column_fast2(P.cam, pc[j].y, pc[j].x, near[j], up[j]);
        near[j]     = near[j] && in[j];
      }
      bool any_near = false;
#pragma unroll
      for (int j = 0; j < PPT; ++j) any_near |= near[j];
      if (any_near) {  // rare, out of line: second tier (side of the rounding edge, binary32), then the exact atan2f
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (near[j]) col[j] = polar_column_resolve(P.cam, pc[j].y, pc[j].x, u2f(rb[j]), col[j], up[j]);
      }
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const bool ok   = in[j] && (unsigned) col[j] < (unsigned) C;
        za[j]           = zb + 4u * (ok ? col[j] : C);
        rb[j]           = ok ? rb[j] : 0u;  // never equals the dummy cell's EMPTY
        if (ok && sm::atom_min_u32<0>(za[j], rb[j]) == rb[j]) sm::st_u32<BC_TIE>(bca, 1u);  // an equal rho was there
      }
    }
    __syncthreads();
    const bool tie = sm::ld_u32<BC_TIE>(bca) != 0;  // uniform
    if (tie) {  // exact pass: lowest index among the points of minimal rho (decision D3)
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (sm::ld_u32<0>(za[j]) == rb[j]) sm::atom_min_u32<M::ZI>(za[j], (unsigned) (tid + j * T));
      __syncthreads();
    }
    // phase 2: winners gate against the fixed column (.cpp:61-73) and linearise their correspondence
    const float Xtx = sm::ld_f32<BC_XTX>(bca), Xty = sm::ld_f32<BC_XTY>(bca);
    const float Lc = sm::ld_f32<BC_LC>(bca), Ls = sm::ld_f32<BC_LS>(bca);
    float Xc = 0.f, Xs = 0.f;
    if (SENSOR) Xc = sm::ld_f32<BC_XC>(bca), Xs = sm::ld_f32<BC_XS>(bca);
    float acc[NSUM];
#pragma unroll
    for (int s = 0; s < NSUM; ++s) acc[s] = 0.f;
    unsigned cnt = 0;  // n_inliers | n_kernelized << 16
    static_for<0, PPT>([&](auto jc) {
      constexpr int J = decltype(jc)::value;
      bool win        = sm::ld_u32<0>(za[J]) == rb[J];
      if (tie) win = win && sm::ld_u32<M::ZI>(za[J]) == (unsigned) (tid + J * T);
      if (win) {
        const float fd  = sm::ld_f32<M::FD>(za[J]);
        const float4 F  = sm::ld_f32x4<0>(4u * za[J] - fk);
        const float2 Mn = sm::ld_f32x2<J * T * 8>(mna);
        if (FUSED)
          linearize2f<SENSOR, J == 0>(P, fd, F, mp[J].x, mp[J].y, Mn, u2f(rb[J]), Xtx, Xty, Xc, Xs, Lc, Ls, acc, cnt);
        else
          linearize2<SENSOR, J == 0, P2P>(P, fd, F, mp[J].x, mp[J].y, Mn, u2f(rb[J]), Xtx, Xty, Xc, Xs, Lc, Ls, acc, cnt);
      }
    });
    store_partials2(acc, cnt, wt, rrow, lane);
    __syncthreads();
    // hand the touched cells back for the next pass (every toucher writes the same EMPTY values)
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      sm::st_u32<0>(za[j], Z_EMPTY_DEPTH);
      if (tie) sm::st_u32<M::ZI>(za[j], Z_EMPTY_IDX);
    }
    if (warp == 0) {
      if (tie && tid == 0) sm::st_u32<BC_TIE>(bca, 0u);  // everybody read it before the barrier above
      warp0_update<T, SENSOR, true>(P, A, bc, red, pair, it, lane, tot, tot_cnt);
    }
    __syncthreads();
    const int stop = (int) sm::ld_u32<BC_STOP>(bca);
    if (stop) {
      status = stop - 1;
      break;
    }
  }

  if (tid < 32) write_result(P, A, bc, pair, it, status, tot, tot_cnt);
}

}  // namespace ls2d
