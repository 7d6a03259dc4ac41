// oracle/ref_cabi.cu -- extern "C" doorway onto the UNMODIFIED reference CUDA kernels.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference's kernels are compiled from
// where they lie under /root/reference by oracle/Makefile; this file only declares their
// launcher prototypes (exactly as the reference headers do) and forwards to them, plus the
// host-side halves that live in the reference's .cpp wrappers and cannot be compiled without
// libtorch: the cudaMalloc / D2H / greedy scan of iou3d.cpp:73-120 is restated here.
//
// launchers: pointnet2_lib/pointnet2/src/{sampling,ball_query,group_points,interpolate}_gpu.h,
//            lib/utils/iou3d/src/iou3d.cpp:24-28, lib/utils/roipool3d/src/roipool3d.cpp:8-12
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <vector>

void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float *points, const int *idx, float *out, cudaStream_t stream);
void gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void furthest_point_sampling_kernel_launcher(int b, int n, int m, const float *dataset, float *temp, int *idxs, cudaStream_t stream);
void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz, int *idx, cudaStream_t stream);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out, cudaStream_t stream);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, cudaStream_t stream);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out, cudaStream_t stream);
void three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points, cudaStream_t stream);
void boxesoverlapLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_overlap);
void boxesioubevLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_iou);
void nmsLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);
void nmsNormalLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);
void roipool3dLauncher_slow(int batch_size, int pts_num, int boxes_num, int feature_in_len, int sampled_pts_num, const float *xyz, const float *boxes3d, const float *pts_feature, float *pooled_features, int *pooled_empty_flag);
void roipool3dLauncher(int batch_size, int pts_num, int boxes_num, int feature_in_len, int sampled_pts_num, const float *xyz, const float *boxes3d, const float *pts_feature, float *pooled_features, int *pooled_empty_flag);

extern "C" {

void ref_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream) {
    furthest_point_sampling_kernel_launcher(b, n, m, xyz, temp, idx, (cudaStream_t)stream);
}
void ref_gather(int b, int c, int n, int np, const float *pts, const int *idx, float *out, void *stream) {
    gather_points_kernel_launcher_fast(b, c, n, np, pts, idx, out, (cudaStream_t)stream);
}
void ref_gather_grad(int b, int c, int n, int np, const float *go, const int *idx, float *gp, void *stream) {
    gather_points_grad_kernel_launcher_fast(b, c, n, np, go, idx, gp, (cudaStream_t)stream);
}
// positional order of the reference call site: (b,n,m,radius,nsample,new_xyz,xyz,idx), pointnet2_utils.py:220
void ref_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx, void *stream) {
    ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx, (cudaStream_t)stream);
}
void ref_group(int b, int c, int n, int np, int ns, const float *pts, const int *idx, float *out, void *stream) {
    group_points_kernel_launcher_fast(b, c, n, np, ns, pts, idx, out, (cudaStream_t)stream);
}
void ref_group_grad(int b, int c, int n, int np, int ns, const float *go, const int *idx, float *gp, void *stream) {
    group_points_grad_kernel_launcher_fast(b, c, n, np, ns, go, idx, gp, (cudaStream_t)stream);
}
void ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, void *stream) {
    three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx, (cudaStream_t)stream);
}
void ref_three_interpolate(int b, int c, int m, int n, const float *pts, const int *idx, const float *w, float *out, void *stream) {
    three_interpolate_kernel_launcher_fast(b, c, m, n, pts, idx, w, out, (cudaStream_t)stream);
}
void ref_three_interpolate_grad(int b, int c, int n, int m, const float *go, const int *idx, const float *w, float *gp, void *stream) {
    three_interpolate_grad_kernel_launcher_fast(b, c, n, m, go, idx, w, gp, (cudaStream_t)stream);
}
// iou3d / roipool3d launch on the legacy default stream, as the reference does
void ref_boxes_overlap_bev(int na, const float *a, int nb, const float *b, float *out) { boxesoverlapLauncher(na, a, nb, b, out); }
void ref_boxes_iou_bev(int na, const float *a, int nb, const float *b, float *out) { boxesioubevLauncher(na, a, nb, b, out); }

// mask only (device pointer out), for bit-exact mask comparison
void ref_nms_mask(const float *boxes, int n, float thresh, int normal, unsigned long long *mask_dev) {
    if (normal) nmsNormalLauncher(boxes, mask_dev, n, thresh); else nmsLauncher(boxes, mask_dev, n, thresh);
}

// whole nms_gpu / nms_normal_gpu of iou3d.cpp:73-170: malloc, kernel, blocking D2H, free, host scan
int ref_nms(const float *boxes, int n, float thresh, int normal, long long *keep_host) {
    const int col_blocks = (n + 63) / 64;
    unsigned long long *mask_dev = nullptr;
    if (cudaMalloc((void **)&mask_dev, sizeof(unsigned long long) * (size_t)n * col_blocks) != cudaSuccess) return -1;
    if (normal) nmsNormalLauncher(boxes, mask_dev, n, thresh); else nmsLauncher(boxes, mask_dev, n, thresh);
    std::vector<unsigned long long> mask((size_t)n * col_blocks);
    if (cudaMemcpy(mask.data(), mask_dev, sizeof(unsigned long long) * mask.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    cudaFree(mask_dev);
    std::vector<unsigned long long> remv(col_blocks, 0ULL);
    int num_to_keep = 0;
    for (int i = 0; i < n; i++) {
        int nblock = i / 64, inblock = i % 64;
        if (!(remv[nblock] & (1ULL << inblock))) {
            keep_host[num_to_keep++] = i;
            const unsigned long long *p = mask.data() + (size_t)i * col_blocks;
            for (int j = nblock; j < col_blocks; j++) remv[j] |= p[j];
        }
    }
    return num_to_keep;
}

void ref_roipool3d(int B, int N, int M, int C, int S, const float *xyz, const float *boxes3d, const float *feat, float *pooled, int *empty) {
    roipool3dLauncher(B, N, M, C, S, xyz, boxes3d, feat, pooled, empty);
}
void ref_roipool3d_slow(int B, int N, int M, int C, int S, const float *xyz, const float *boxes3d, const float *feat, float *pooled, int *empty) {
    roipool3dLauncher_slow(B, N, M, C, S, xyz, boxes3d, feat, pooled, empty);
}

}  // extern "C"
