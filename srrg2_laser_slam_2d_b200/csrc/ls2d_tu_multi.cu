// ls2d_tu_multi.cu -- the multi-slice aligner (ls2d_multi.cuh)
#include "ls2d_internal.h"
#include "ls2d_multi.cuh"

namespace ls2d {

int multi_reduction_threads() { return MULTI_THREADS; }

int launch_multi(ls2d_handle* h, const multi_args& a, const int* cols) {
  if (a.n_pairs <= 0) return LS2D_OK;
  constexpr int T   = MULTI_THREADS;  // 4 CTAs of 8 warps per SM measured best (512 x 2: +28 % time, 384 x 3: +16 %)
  const size_t smem = multi_smem_bytes(cols, a.n_slices, a.max_cols, a.max_points, T);
  if (smem > SMEM_LIMIT) return LS2D_ERR_UNSUPPORTED;
  auto kern = icp_multi_kernel<T, 4>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  kern<<<a.n_pairs, T, smem, h->stream>>>(a);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

}  // namespace ls2d
