// ls2d_mathcheck.cpp -- host build of ls2d_math.cuh for the CPU test-suite (tests/test_math_host.py):
// lets the tests compare the device math (which is the same source, compiled operation by operation)
// against the host libm without a GPU.  Not part of the product path.
#include <vector>

#include "ls2d_math.cuh"

extern "C" {
float ls2d_host_atan2f(float y, float x) { return ls2d::atan2f_fdlibm(y, x); }
float ls2d_host_sinf(float x) { return ls2d::sinf_glibc(x); }
float ls2d_host_cosf(float x) { return ls2d::cosf_glibc(x); }
float ls2d_host_atan2f_fast(float y, float x) { return ls2d::atan2f_fast(y, x); }
int ls2d_host_polar_column(int cols, float amin, float amax, float y, float x) {
  return ls2d::polar_column(ls2d::make_polar_cam(cols, amin, amax), y, x);
}
int ls2d_host_polar_column_exact(int cols, float amin, float amax, float y, float x) {
  return ls2d::polar_column_exact(ls2d::make_polar_cam(cols, amin, amax), y, x);
}
// bulk drivers (n inputs) so python does not loop
void ls2d_host_atan2f_n(const float* y, const float* x, float* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = ls2d::atan2f_fdlibm(y[i], x[i]);
}
void ls2d_host_sincosf_n(const float* x, float* s, float* c, long n) {
  for (long i = 0; i < n; ++i) s[i] = ls2d::sinf_glibc(x[i]), c[i] = ls2d::cosf_glibc(x[i]);
}
void ls2d_host_logf_n(const float* x, float* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = ls2d::logf_glibc(x[i]);
}
void ls2d_host_polar_column_n(int cols, float amin, float amax, const float* y, const float* x, int* fast,
                              int* exact, long n) {
  const ls2d::polar_cam k = ls2d::make_polar_cam(cols, amin, amax);
  for (long i = 0; i < n; ++i) fast[i] = ls2d::polar_column(k, y[i], x[i]), exact[i] = ls2d::polar_column_exact(k, y[i], x[i]);
}
// the three-tier decision of icp_fused2_kernel (fast proposal -> side of the rounding edge -> exact), with the
// tier that decided each point (1, 2, 3) so the tests can report how rare the exact path is
void ls2d_host_polar_column_tiered_n(int cols, float amin, float amax, const float* y, const float* x, int* col,
                                     int* tier, long n) {
  ls2d::polar_cam k = ls2d::make_polar_cam(cols, amin, amax);
  std::vector<ls2d::polar_edge> edges((size_t) cols + 1);
  ls2d::fill_polar_edges(k, edges.data());
  k.edge = edges.data();
  for (long i = 0; i < n; ++i) {
    bool near, up;
    int c        = ls2d::polar_column_fast2(k, y[i], x[i], near, up);
    const int kb = c + (up ? 1 : 0);
    tier[i] = 1;
    if (near) {
      const float rho = ls2d::fsqrt(ls2d::fadd(ls2d::fmul(x[i], x[i]), ls2d::fmul(y[i], y[i])));
      bool undecided;
      c       = ls2d::polar_column_edge(k, y[i], x[i], rho, kb, undecided);
      tier[i] = 2;
      if (undecided) c = ls2d::polar_column_exact(k, y[i], x[i]), tier[i] = 3;
    }
    col[i] = (c < 0 || c >= cols) ? -1 : c;
  }
}
void ls2d_host_range_gate2(float range_min, float range_max, float* lo, float* hi) {
  const ls2d::range_gate2 g = ls2d::make_range_gate2(range_min, range_max);
  *lo = g.lo, *hi = g.hi;
}
float ls2d_host_edge_tol(int cols, float amin, float amax) { return ls2d::make_polar_cam(cols, amin, amax).edge_tol; }
float ls2d_host_margin(int cols, float amin, float amax) { return ls2d::make_polar_cam(cols, amin, amax).margin; }
}
