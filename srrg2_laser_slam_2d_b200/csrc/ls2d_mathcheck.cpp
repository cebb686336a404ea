// ls2d_mathcheck.cpp -- host build of ls2d_math.cuh for the CPU test-suite (tests/test_math_host.py):
// lets the tests compare the device math (which is the same source, compiled operation by operation)
// against the host libm without a GPU.  Not part of the product path.
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
void ls2d_host_polar_column_n(int cols, float amin, float amax, const float* y, const float* x, int* fast,
                              int* exact, long n) {
  const ls2d::polar_cam k = ls2d::make_polar_cam(cols, amin, amax);
  for (long i = 0; i < n; ++i) fast[i] = ls2d::polar_column(k, y[i], x[i]), exact[i] = ls2d::polar_column_exact(k, y[i], x[i]);
}
float ls2d_host_margin(int cols, float amin, float amax) { return ls2d::make_polar_cam(cols, amin, amax).margin; }
}
