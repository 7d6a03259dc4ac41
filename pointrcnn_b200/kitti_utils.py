"""The three box helpers of lib/utils/kitti_utils.py that sit ON the operator path (reference :45-63,
:134-160).  Inside the reference tree the wrappers import the reference's own `lib.utils.kitti_utils`;
this module is what they fall back to when that tree is absent (GPU box, tests, bench)."""
import numpy as np
import torch


def boxes3d_to_bev_torch(boxes3d):
    """(N,7) [x,y,z,h,w,l,ry] -> (N,5) [x1,y1,x2,y2,ry] in the x-z plane"""
    boxes_bev = boxes3d.new_empty((boxes3d.shape[0], 5))
    cu, cv = boxes3d[:, 0], boxes3d[:, 2]
    half_l, half_w = boxes3d[:, 5] / 2, boxes3d[:, 4] / 2
    boxes_bev[:, 0], boxes_bev[:, 1] = cu - half_l, cv - half_w
    boxes_bev[:, 2], boxes_bev[:, 3] = cu + half_l, cv + half_w
    boxes_bev[:, 4] = boxes3d[:, 6]
    return boxes_bev


def enlarge_box3d(boxes3d, extra_width):
    """h,w,l += 2*extra_width; y (bottom centre) += extra_width"""
    large = boxes3d.copy() if isinstance(boxes3d, np.ndarray) else boxes3d.clone()
    large[:, 3:6] += extra_width * 2
    large[:, 1] += extra_width
    return large


def rotate_pc_along_y_torch(pc, rot_angle):
    """pc (N,512,3+C), rot_angle (N): rotate the (x,z) columns by rot_angle, in place"""
    cosa = torch.cos(rot_angle).view(-1, 1)
    sina = torch.sin(rot_angle).view(-1, 1)
    R = torch.stack([torch.cat([cosa, -sina], dim=1), torch.cat([sina, cosa], dim=1)], dim=1)   # (N,2,2)
    pc[:, :, [0, 2]] = torch.matmul(pc[:, :, [0, 2]], R.permute(0, 2, 1))
    return pc
