// Intrinsic latency of the Gauss-Newton tail of the ICP kernels (one lane, nothing else on the SM):
// reduce over 9 warps -> binary64 LDL^T -> v2t (sincos) -> compose -> publish.  nvcc -arch=sm_100a, run on the box.
#include <cstdio>
#include "../../srrg2_laser_slam_2d_b200/csrc/ls2d_common.cuh"
using namespace ls2d;

__global__ void k(const float* in, float* out, long long* cyc, dev_params P) {
  __shared__ pose_bc bc;
  __shared__ float red[9 * RED_STRIDE];
  for (int i = threadIdx.x; i < 9 * RED_STRIDE; i += 32) red[i] = in[i];
  if (threadIdx.x == 0) publish_pose(&bc, P, iso_identity(), false, 0);
  __syncwarp();
  long long t[6];
  float acc = 0.f;
  for (int rep = 0; rep < 5; ++rep) {
    const int lane = threadIdx.x;
    t[0] = clock64();
    float tot = 0.f;
    if (lane < NSUM) {
      tot = red[lane];
#pragma unroll
      for (int w = 1; w < 9; ++w) tot = fadd(tot, red[w * RED_STRIDE + lane]);
    }
    float v[NSUM];
#pragma unroll
    for (int s = 0; s < NSUM; ++s) v[s] = __shfl_sync(0xffffffffu, tot, s);
    t[1] = clock64();
    float dx[3] = {0, 0, 0};
    iso X;
    if (lane == 0) {
      const bool ok = solve3(v, P.damping, dx);
      acc += ok;
      t[2] = clock64();
      X.tx = bc.Xtx, X.ty = bc.Xty, X.c = bc.Xc, X.s = bc.Xs;
      const iso D = iso_v2t(dx[0], dx[1], dx[2]);
      t[3] = clock64();
      X = iso_compose(X, D);
      publish_pose(&bc, P, X, false, 0);
      t[4] = clock64();
    }
    __syncwarp();
    red[1] += 1e-3f * bc.Xtx;  // keep the chain alive
    __syncwarp();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) cyc[i] = t[i + 1] - t[i];
    out[0] = acc + bc.Xtx;
  }
}

int main() {
  float h[9 * RED_STRIDE];
  for (int w = 0; w < 9; ++w)
    for (int s = 0; s < RED_STRIDE; ++s) h[w * RED_STRIDE + s] = 0.f;
  // a well-conditioned system spread over the warps: H = diag(50, 60, 400) + small off-diagonals, b small
  const float H[11] = {50.f, 1.f, 2.f, 60.f, 3.f, 400.f, 0.5f, -0.3f, 0.2f, 1.f, 0.f};
  for (int w = 0; w < 9; ++w)
    for (int s = 0; s < 11; ++s) h[w * RED_STRIDE + s] = H[s] / 9.f;
  float *din, *dout;
  long long* dc;
  cudaMalloc(&din, sizeof(h)), cudaMalloc(&dout, 64), cudaMalloc(&dc, 64);
  cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
  dev_params P;
  memset(&P, 0, sizeof(P));
  P.damping = 0.f;
  k<<<1, 32>>>(din, dout, dc, P);
  long long c[4];
  cudaMemcpy(c, dc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("cycles: reduce+gather %lld, solve3 (binary64 LDL^T) %lld, v2t (sincos) %lld, compose+publish %lld  [%s]\n", c[0],
         c[1], c[2], c[3], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
