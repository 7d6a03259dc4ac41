"""`pointnet2_cuda` -- same exports as pointnet2_lib/pointnet2/src/pointnet2_api.cpp:10-23.

Every function takes the reference's positional arguments (explicit dims + caller-allocated tensors),
sets the device of the first tensor, launches on torch's current stream and raises RuntimeError on
failure (the reference prints and exit(-1)s).
"""
import torch

from .. import _cabi as C


def _guard(t):
    C.require_cuda(t)
    return torch.cuda.device(t.device)


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    C.require_cuda(new_xyz, xyz, idx); C.require_contig(new_xyz, xyz, idx)
    with _guard(xyz):
        C.check(C.lib().prb_ball_query(int(b), int(n), int(m), C.c_float(radius), int(nsample), C.ptr(new_xyz),
                                       C.ptr(xyz), C.ptr(idx), C.stream()), "ball_query")
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    C.require_cuda(points, idx, out); C.require_contig(points, idx, out)
    with _guard(points):
        C.check(C.lib().prb_group_points(int(b), int(c), int(n), int(npoints), int(nsample), C.ptr(points), C.ptr(idx),
                                         C.ptr(out), C.stream()), "group_points")
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    C.require_cuda(grad_out, idx, grad_points); C.require_contig(grad_out, idx, grad_points)
    with _guard(grad_out):
        C.check(C.lib().prb_group_points_grad(int(b), int(c), int(n), int(npoints), int(nsample), C.ptr(grad_out),
                                              C.ptr(idx), C.ptr(grad_points), C.stream()), "group_points_grad")
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    C.require_cuda(points, idx, out); C.require_contig(points, idx, out)
    with _guard(points):
        C.check(C.lib().prb_gather_points(int(b), int(c), int(n), int(npoints), C.ptr(points), C.ptr(idx), C.ptr(out),
                                          C.stream()), "gather_points")
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    C.require_cuda(grad_out, idx, grad_points); C.require_contig(grad_out, idx, grad_points)
    with _guard(grad_out):
        C.check(C.lib().prb_gather_points_grad(int(b), int(c), int(n), int(npoints), C.ptr(grad_out), C.ptr(idx),
                                               C.ptr(grad_points), C.stream()), "gather_points_grad")
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    C.require_cuda(points, temp, idx); C.require_contig(points, temp, idx)
    with _guard(points):
        lib = C.lib()
        wsb = lib.prb_fps_workspace_bytes(int(b), int(n))
        ws = torch.empty(wsb, dtype=torch.uint8, device=points.device) if wsb else None
        C.check(lib.prb_furthest_point_sampling_ws(int(b), int(n), int(m), C.ptr(points), C.ptr(temp), C.ptr(idx), None,
                                                   C.ptr(ws), C.c_size_t(wsb), C.stream()), "furthest_point_sampling")
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    C.require_cuda(unknown, known, dist2, idx); C.require_contig(unknown, known, dist2, idx)
    with _guard(unknown):
        C.check(C.lib().prb_three_nn(int(b), int(n), int(m), C.ptr(unknown), C.ptr(known), C.ptr(dist2), C.ptr(idx), None,
                                     C.stream()), "three_nn")


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    C.require_cuda(points, idx, weight, out); C.require_contig(points, idx, weight, out)
    with _guard(points):
        C.check(C.lib().prb_three_interpolate(int(b), int(c), int(m), int(n), C.ptr(points), C.ptr(idx), C.ptr(weight),
                                              C.ptr(out), C.stream()), "three_interpolate")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    C.require_cuda(grad_out, idx, weight, grad_points); C.require_contig(grad_out, idx, weight, grad_points)
    with _guard(grad_out):
        C.check(C.lib().prb_three_interpolate_grad(int(b), int(c), int(n), int(m), C.ptr(grad_out), C.ptr(idx),
                                                   C.ptr(weight), C.ptr(grad_points), C.stream()), "three_interpolate_grad")
