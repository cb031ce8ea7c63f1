"""Drop-in for the one kornia symbol the reference uses: ``kornia.geometry.HomographyWarper``
(kornia==0.1.4.post2, requirements.txt:59; call sites quick_start/align2images.py:61,65,
evaluation/evalHpatch/evaluation.py:190,218).  ``warp_grid(H)`` = homogeneous multiply of the
normalised base grid by H and division by z (no inversion, no epsilon)."""
import torch

from . import ops


class HomographyWarper:
    def __init__(self, height, width, mode="bilinear", padding_mode="zeros", normalized_coordinates=True):
        self.height, self.width = int(height), int(width)
        self.mode, self.padding_mode = mode, padding_mode

    def warp_grid(self, dst_homo_src):
        """(N,3,3) -> (N,h,w,2) sampling grid in [-1,1] coordinates (x, y)."""
        H = dst_homo_src
        if not H.is_cuda:
            raise ops._lib.RFError("HomographyWarper.warp_grid: CUDA tensors only (no CPU path)")
        return ops.warp_grid(H, self.height, self.width)

    def __call__(self, patch_src, dst_homo_src):
        return ops.grid_sample(patch_src, self.warp_grid(dst_homo_src))

    forward = __call__
