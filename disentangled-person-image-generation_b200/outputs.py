"""Result files of the reference's test() / generate() (tester.py:138-202, 213-253; trainer.py:498-526; utils.py:157-182)
which score.py consumes (score.py:33-36): per-sample PNGs under
    <model_dir>/<test_dir_name>/{x, x_target, G, pose, pose_target, G_pose, mask, mask_target}/
with the reference's file-name patterns, plus the 8-per-row sample grids of save_image().

Everything numeric stays on the GPU until the end (denorm -> uint8 by dpig_denorm_u8, SSIM by dpig_ssim_gray_u8); PNG
encoding is host work and is batched over a thread pool (PIL releases the GIL while it deflates), so one test() batch
of 8 x batch_size files does not serialise behind the Python thread the way the reference's per-file loop does."""
import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

DIRS = ("x", "x_target", "G", "pose", "pose_target", "G_pose", "mask", "mask_target")


def make_grid(tensor, nrow=8, padding=2):
    """utils.make_grid (utils.py:157-175): uint8 [n,h,w,c] -> one uint8 [H,W,3] sheet, nrow images per row."""
    tensor = np.asarray(tensor)
    if tensor.ndim == 3:
        tensor = tensor[..., None]
    nmaps = tensor.shape[0]
    xmaps = min(nrow, nmaps)
    ymaps = int(math.ceil(float(nmaps) / xmaps))
    height, width = int(tensor.shape[1] + padding), int(tensor.shape[2] + padding)
    grid = np.zeros([height * ymaps + 1 + padding // 2, width * xmaps + 1 + padding // 2, 3], dtype=np.uint8)
    k = 0
    for y in range(ymaps):
        for x in range(xmaps):
            if k >= nmaps:
                break
            h, hh = y * height + 1 + padding // 2, height - padding
            w, ww = x * width + 1 + padding // 2, width - padding
            grid[h:h + hh, w:w + ww] = tensor[k].astype(np.uint8)     # a 1-channel image broadcasts over RGB
            k += 1
    return grid


def _save_png(arr, path):
    from PIL import Image
    Image.fromarray(arr).save(path)


def save_image(tensor, filename, nrow=8, padding=2):
    """utils.save_image (utils.py:177-182)."""
    _save_png(make_grid(tensor, nrow=nrow, padding=padding), filename)


class ResultWriter:
    """Writes the per-sample files of one test() run; `add_batch` returns immediately, `close()` waits for the pool."""

    def __init__(self, root, workers=8):
        self.root = root
        os.makedirs(root, exist_ok=True)
        for d in DIRS:
            os.makedirs(os.path.join(root, d), exist_ok=True)
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pending = []
        self.files = 0

    def _submit(self, arr, path):
        self.pending.append(self.pool.submit(_save_png, np.ascontiguousarray(arr, dtype=np.uint8), path))
        self.files += 1

    def add_batch(self, i, batch_size, x, x_target, G, pose_max, pose_target_max, G_pose, mask, mask_target, scores):
        """Batch i of test(): all arrays in [0,255]; x / x_target / G / G_pose [B,H,W,3], pose_max / pose_target_max /
        mask / mask_target [B,H,W] (tester.py:177-195)."""
        r = self.root
        for j in range(batch_size):
            idx = i * batch_size + j
            self._submit(x[j], "%s/x/%05d.png" % (r, idx))
            self._submit(x_target[j], "%s/x_target/%05d.png" % (r, idx))
            self._submit(G[j], "%s/G/%04d_c1s1_%06d_%05d_%f.png" % (r, i, j, idx, float(scores[j])))
            self._submit(pose_max[j], "%s/pose/%05d.png" % (r, idx))
            self._submit(pose_target_max[j], "%s/pose_target/%05d.png" % (r, idx))
            self._submit(G_pose[j], "%s/G_pose/%04d_%04d.png" % (r, i, j))
            self._submit(np.squeeze(mask[j]), "%s/mask/%05d.png" % (r, idx))
            self._submit(np.squeeze(mask_target[j]), "%s/mask_target/%05d.png" % (r, idx))

    def add_grid(self, tensor, name):
        self.pending.append(self.pool.submit(save_image, np.asarray(tensor), os.path.join(self.root, name)))
        self.files += 1

    def close(self):
        for f in self.pending:
            f.result()           # re-raises an I/O error of a worker
        self.pending = []
        self.pool.shutdown()
        return self.files
