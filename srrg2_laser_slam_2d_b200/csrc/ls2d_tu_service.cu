// ls2d_tu_service.cu -- projector / finder / clipper / merger / best-of kernels (ls2d_service.cuh) and the raw-scan
// pre-processor, voxelizing clipper and CSR packing built on the same z-buffer helpers (ls2d_scan.cuh)
#include <algorithm>

#include "ls2d_internal.h"
#include "ls2d_scan.cuh"

namespace ls2d {

int launch_project(ls2d_handle* h, const project_args& a) {
  const size_t smem = sizeof(unsigned) * 2 * (size_t) h->dp.cam.cols;
  if (int rc = configure_kernel(h, project_kernel, (size_t) ((int) smem), false)) return rc;
  project_kernel<<<1, 256, smem, h->stream>>>(h->dp, a);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_correspond(ls2d_handle* h, const correspond_args& a) {
  const size_t smem = sizeof(unsigned) * 4 * (size_t) h->dp.cam.cols;
  if (int rc = configure_kernel(h, correspond_kernel, (size_t) ((int) smem), false)) return rc;
  correspond_kernel<<<1, 256, smem, h->stream>>>(h->dp, a);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_clip(ls2d_handle* h, const clip_args& a, int n) {
  if (n <= 0) return LS2D_OK;
  if (h->dp.cam.cols > 32 * CLIP_T) return LS2D_ERR_UNSUPPORTED;
  const size_t smem = sizeof(unsigned) * 2 * (size_t) h->dp.cam.cols;
  if (int rc = configure_kernel(h, clip_kernel, (size_t) ((int) smem), false)) return rc;
  clip_kernel<<<n, CLIP_T, smem, h->stream>>>(h->dp, a);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_merge(ls2d_handle* h, const merge_args& a) {
  const size_t smem = sizeof(unsigned) * 4 * (size_t) h->dp.cam.cols;
  if (int rc = configure_kernel(h, merge_kernel, (size_t) ((int) smem), false)) return rc;
  merge_kernel<<<1, 256, smem, h->stream>>>(h->dp, a);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_classify(ls2d_handle* h, const classify_args& a, const int* off_f, int cloud_f, const int* off_m, int cloud_m) {
  if (a.n <= 0) return LS2D_OK;
  classify_dev d;
  d.off_f = off_f, d.off_m = off_m, d.cloud_f = cloud_f, d.cloud_m = cloud_m;
  classify_kernel<<<(a.n + 127) / 128, 128, 0, h->stream>>>(h->dp, a, d);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_selftest_sqrt(ls2d_handle* h, unsigned lo_bits, unsigned hi_bits, unsigned long long* n_mismatch_dev) {
  selftest_sqrt_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(lo_bits, hi_bits, n_mismatch_dev);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_best_of(ls2d_handle* h, const ls2d_result* res, int n, int n_guess, const ls2d_gates& g, int candidate_base,
                   ls2d_best* out) {
  best_of_kernel<<<1, 1024, 0, h->stream>>>(res, n, n_guess, g, candidate_base, out);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_best_of_groups(ls2d_handle* h, const ls2d_result* res, const int* group_off, int n_groups,
                          const int* moving_id, const ls2d_gates& g, ls2d_best* out) {
  if (n_groups <= 0) return LS2D_OK;
  best_of_groups_kernel<<<(n_groups + 7) / 8, 256, 0, h->stream>>>(res, group_off, n_groups, moving_id, g, out);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}


int launch_preprocess(ls2d_handle* h, const scan_dev_params& P, const scan_args& a_in, int n_scans) {
  if (n_scans <= 0) return LS2D_OK;
  if (P.n_beams > 32 * SCAN_T || P.n_beams >= (1 << 14)) return LS2D_ERR_UNSUPPORTED;
  const size_t smem = scan_smem_bytes(P.n_beams, P.inv_res != 0.f);
  if (smem > SMEM_LIMIT) return LS2D_ERR_UNSUPPORTED;
  int rc;
  if (!h->d_beam.p || h->beam_n != P.n_beams || memcmp(&h->beam_ifx, &P.ifx, 4) || memcmp(&h->beam_cx, &P.cx, 4)) {
    if ((rc = reserve(h->d_beam, sizeof(float2) * (size_t) P.n_beams))) return rc;
    beam_table_kernel<<<(P.n_beams + 255) / 256, 256, 0, h->stream>>>(P.ifx, P.cx, P.n_beams, (float2*) h->d_beam.p);
    CU(cudaGetLastError());
    h->launches++;
    h->beam_n = P.n_beams, h->beam_ifx = P.ifx, h->beam_cx = P.cx;
  }
  scan_args a = a_in;
  a.beam_cs   = (const float2*) h->d_beam.p;
  if (int rc = configure_kernel(h, preprocess_kernel, (size_t) ((int) smem), false)) return rc;
  if (!h->d_ticket.p) {
    if ((rc = reserve(h->d_ticket, sizeof(int)))) return rc;
    CU(cudaMemsetAsync(h->d_ticket.p, 0, sizeof(int), h->stream));
  }
  a.ticket = (int*) h->d_ticket.p;
  preprocess_kernel<<<n_scans, SCAN_T, smem, h->stream>>>(P, a);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_clip_voxel(ls2d_handle* h, const clip_args& a, int n, float inv_res) {
  if (n <= 0) return LS2D_OK;
  const int C = h->dp.cam.cols;
  if (C > 32 * SCAN_T || C >= (1 << 14)) return LS2D_ERR_UNSUPPORTED;
  const size_t smem = clip_voxel_smem_bytes(C);
  if (smem > SMEM_LIMIT) return LS2D_ERR_UNSUPPORTED;
  if (int rc = configure_kernel(h, clip_voxel_kernel, (size_t) ((int) smem), false)) return rc;
  clip_voxel_kernel<<<n, SCAN_T, smem, h->stream>>>(h->dp, a, inv_res);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_largest_cloud(ls2d_handle* h, const int* off, int n, int* out) {
  CU(cudaMemsetAsync(out, 0, sizeof(int), h->stream));
  largest_cloud_kernel<<<(n + 1023) / 1024 > 1024 ? 1024 : (n + 1023) / 1024, 1024, 0, h->stream>>>(off, n, out);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

int launch_scan_pack(ls2d_handle* h, const float4* strided, const int* off, int stride, int n, float4* packed) {
  if (n <= 0) return LS2D_OK;
  scan_pack_kernel<<<n, 128, 0, h->stream>>>(strided, off, stride, packed);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

}  // namespace ls2d
