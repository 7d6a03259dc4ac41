// iou3d_dev.cuh -- device functions of the rotated / axis-aligned BEV IoU, shared by iou3d.cu (matrices, bitmask
// NMS) and proposal.cu (top-k greedy NMS inside the RPN proposal kernel).
//
// The overlap arithmetic IS the spec (keep masks are compared bit-exactly with lib/utils/iou3d/src/iou3d_kernel.cu:
// 34-221): corners rotated about the box centre, 4x4 edge intersections (bbox reject, straddle test, line solve with
// an EPS fallback), corners of one box inside the other (MARGIN 1e-5), centroid, bubble sort by atan2f, fan area / 2.0.
#pragma once
#include "common.cuh"

namespace prb {

struct P2 {
    float x, y;
};
__device__ __forceinline__ P2 mk(float x, float y) { P2 p; p.x = x; p.y = y; return p; }

__device__ __forceinline__ float cross2(const P2 &a, const P2 &b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float cross3(const P2 &p1, const P2 &p2, const P2 &p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
__device__ __forceinline__ int rect_cross(const P2 &p1, const P2 &p2, const P2 &q1, const P2 &q2) {
    return min(p1.x, p2.x) <= max(q1.x, q2.x) && min(q1.x, q2.x) <= max(p1.x, p2.x) &&
           min(p1.y, p2.y) <= max(q1.y, q2.y) && min(q1.y, q2.y) <= max(p1.y, p2.y);
}
__device__ __forceinline__ int in_box2d(const float *box, const P2 &p) {
    const float MARGIN = 1e-5;
    float center_x = (box[0] + box[2]) / 2;
    float center_y = (box[1] + box[3]) / 2;
    float angle_cos = cos(-box[4]), angle_sin = sin(-box[4]);
    float rot_x = (p.x - center_x) * angle_cos + (p.y - center_y) * angle_sin + center_x;
    float rot_y = -(p.x - center_x) * angle_sin + (p.y - center_y) * angle_cos + center_y;
    return (rot_x > box[0] - MARGIN && rot_x < box[2] + MARGIN && rot_y > box[1] - MARGIN && rot_y < box[3] + MARGIN);
}
__device__ __forceinline__ int seg_intersection(const P2 &p1, const P2 &p0, const P2 &q1, const P2 &q0, P2 &ans) {
    const float EPS = 1e-8;
    if (rect_cross(p0, p1, q0, q1) == 0) return 0;
    float s1 = cross3(q0, p1, p0);
    float s2 = cross3(p1, q1, p0);
    float s3 = cross3(p0, q1, q0);
    float s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross3(q1, p1, p0);
    if (fabs(s5 - s1) > EPS) {
        ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans.x = (b0 * c1 - b1 * c0) / D;
        ans.y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}
__device__ __forceinline__ void rot_center(const P2 &c, const float angle_cos, const float angle_sin, P2 &p) {
    float new_x = (p.x - c.x) * angle_cos + (p.y - c.y) * angle_sin + c.x;
    float new_y = -(p.x - c.x) * angle_sin + (p.y - c.y) * angle_cos + c.y;
    p.x = new_x;
    p.y = new_y;
}
__device__ __forceinline__ int angle_gt(const P2 &a, const P2 &b, const P2 &c) {
    return atan2(a.y - c.y, a.x - c.x) > atan2(b.y - c.y, b.x - c.x);
}

// rotated-rectangle intersection area of two [x1,y1,x2,y2,ry] boxes
__device__ inline float box_overlap(const float *box_a, const float *box_b) {
    float a_x1 = box_a[0], a_y1 = box_a[1], a_x2 = box_a[2], a_y2 = box_a[3], a_angle = box_a[4];
    float b_x1 = box_b[0], b_y1 = box_b[1], b_x2 = box_b[2], b_y2 = box_b[3], b_angle = box_b[4];
    P2 center_a = mk((a_x1 + a_x2) / 2, (a_y1 + a_y2) / 2);
    P2 center_b = mk((b_x1 + b_x2) / 2, (b_y1 + b_y2) / 2);
    P2 ca[5], cb[5];
    ca[0] = mk(a_x1, a_y1); ca[1] = mk(a_x2, a_y1); ca[2] = mk(a_x2, a_y2); ca[3] = mk(a_x1, a_y2);
    cb[0] = mk(b_x1, b_y1); cb[1] = mk(b_x2, b_y1); cb[2] = mk(b_x2, b_y2); cb[3] = mk(b_x1, b_y2);
    float a_angle_cos = cos(a_angle), a_angle_sin = sin(a_angle);
    float b_angle_cos = cos(b_angle), b_angle_sin = sin(b_angle);
    for (int k = 0; k < 4; k++) {
        rot_center(center_a, a_angle_cos, a_angle_sin, ca[k]);
        rot_center(center_b, b_angle_cos, b_angle_sin, cb[k]);
    }
    ca[4] = ca[0];
    cb[4] = cb[0];

    P2 cp[16];
    P2 poly_center = mk(0, 0);
    int cnt = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], cp[cnt])) {
                poly_center = mk(poly_center.x + cp[cnt].x, poly_center.y + cp[cnt].y);
                cnt++;
            }
    for (int k = 0; k < 4; k++) {
        if (in_box2d(box_a, cb[k])) {
            poly_center = mk(poly_center.x + cb[k].x, poly_center.y + cb[k].y);
            cp[cnt] = cb[k];
            cnt++;
        }
        if (in_box2d(box_b, ca[k])) {
            poly_center = mk(poly_center.x + ca[k].x, poly_center.y + ca[k].y);
            cp[cnt] = ca[k];
            cnt++;
        }
    }
    poly_center.x /= cnt;
    poly_center.y /= cnt;

    P2 t;
    for (int j = 0; j < cnt - 1; j++)
        for (int i = 0; i < cnt - j - 1; i++)
            if (angle_gt(cp[i], cp[i + 1], poly_center)) {
                t = cp[i];
                cp[i] = cp[i + 1];
                cp[i + 1] = t;
            }

    float area = 0;
    for (int k = 0; k < cnt - 1; k++)
        area += cross2(mk(cp[k].x - cp[0].x, cp[k].y - cp[0].y), mk(cp[k + 1].x - cp[0].x, cp[k + 1].y - cp[0].y));
    return fabs(area) / 2.0;
}

__device__ __forceinline__ float iou_bev(const float *box_a, const float *box_b) {
    const float EPS = 1e-8;
    float sa = (box_a[2] - box_a[0]) * (box_a[3] - box_a[1]);
    float sb = (box_b[2] - box_b[0]) * (box_b[3] - box_b[1]);
    float s_overlap = box_overlap(box_a, box_b);
    return s_overlap / fmaxf(sa + sb - s_overlap, EPS);
}

__device__ __forceinline__ float iou_normal(const float *a, const float *b) {
    const float EPS = 1e-8;
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    float interS = width * height;
    float Sa = (a[2] - a[0]) * (a[3] - a[1]);
    float Sb = (b[2] - b[0]) * (b[3] - b[1]);
    return interS / fmaxf(Sa + Sb - interS, EPS);
}

}  // namespace prb
