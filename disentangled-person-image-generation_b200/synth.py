"""Deterministic synthetic stand-ins for a Market-1501 batch: there is no dataset or
network here, so tests and bench.py draw inputs of the reference's shapes and value ranges.

Tensors mirror what `_load_batch_pair_pose` hands to build_model (reference trainer.py:537-564):
  x         fp32 [B,H,W,3] in [-1,1]            (process_image(raw, 127.5, 127.5))
  pose_rcv  fp32 [B,18,3]  (row, col, visible)  ('pose_peaks_0_rcv')
  mask      fp32 [B,H,W,1] in {0,1}             ('pose_mask_r6_0')
  part_bbox int  [B,37,4]  (y1,x1,y2,x2) pixels ('part_bbox_0', built like convert_market.py:640-728)
  part_vis  fp32 [B,37]
"""
import numpy as np

# keypoint groups of the first 7 body parts (reference datasets/convert_market.py:664-670)
PART_GROUPS = [[0, 1, 2, 5, 14, 15, 16, 17], [2, 3, 4, 5, 6, 7, 8, 11], [8, 9, 10, 11, 12, 13], [5, 6, 7],
               [2, 3, 4], [11, 12, 13], [8, 9, 10]]


def part_boxes(rcv, img_h, img_w, radius=7, r_single=10, n_parts=37):
    """get_part_bbox37 rule: min/max of the visible member keypoints +- radius (single keypoint:
    +- 10), clamped to the image; [0,0,1,1] and visibility 0 if no member is visible."""
    groups = list(PART_GROUPS)
    groups += [[2, 5, 8, 11], [5, 6], [6, 7], [2, 3], [3, 4], [11, 12], [12, 13], [8, 9], [9, 10], list(range(18))]
    groups += [[i] for i in range(18)]
    groups += [[2, 3, 4, 8, 9, 10], [5, 6, 7, 11, 12, 13]]
    groups = groups[:n_parts]
    B = rcv.shape[0]
    bbox = np.zeros((B, n_parts, 4), np.int64)
    vis = np.zeros((B, n_parts), np.float32)
    for b in range(B):
        for i, g in enumerate(groups):
            pts = [(rcv[b, k, 0], rcv[b, k, 1]) for k in g if rcv[b, k, 2] > 0]
            if not pts:
                bbox[b, i] = [0, 0, 1, 1]
                continue
            ys = [int(p[0]) for p in pts]
            xs = [int(p[1]) for p in pts]
            r = radius if len(pts) > 1 else r_single
            bbox[b, i] = [max(0, min(ys) - r), max(0, min(xs) - r), min(img_h - 1, max(ys) + r),
                          min(img_w - 1, max(xs) + r)]
            vis[b, i] = 1.0
    return bbox, vis


def make_batch(batch, img_h=128, img_w=64, seed=123, keypoints=18, n_parts=37):
    """One synthetic batch as numpy arrays (see module docstring)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.0, 1.0, size=(batch, img_h, img_w, 3)).astype(np.float32)
    rcv = np.zeros((batch, keypoints, 3), np.float32)
    # a loose standing skeleton jittered per sample, so ROIs have realistic extents
    cy, cx = img_h * 0.5, img_w * 0.5
    rcv[:, :, 0] = np.clip(rng.normal(cy, img_h * 0.22, size=(batch, keypoints)), 0, img_h - 1)
    rcv[:, :, 1] = np.clip(rng.normal(cx, img_w * 0.18, size=(batch, keypoints)), 0, img_w - 1)
    rcv[:, :, 0:2] = np.floor(rcv[:, :, 0:2])
    rcv[:, :, 2] = (rng.uniform(size=(batch, keypoints)) < 0.85).astype(np.float32)
    yy, xx = np.mgrid[0:img_h, 0:img_w]
    mask = np.zeros((batch, img_h, img_w, 1), np.float32)
    for b in range(batch):
        ay = img_h * rng.uniform(0.36, 0.44)
        ax = img_w * rng.uniform(0.24, 0.32)
        mask[b, :, :, 0] = (((yy - cy) / ay) ** 2 + ((xx - cx) / ax) ** 2 <= 1.0).astype(np.float32)
    bbox, vis = part_boxes(rcv, img_h, img_w, n_parts=n_parts)
    return dict(x=x, pose_rcv=rcv, mask=mask, part_bbox=bbox, part_vis=vis)
