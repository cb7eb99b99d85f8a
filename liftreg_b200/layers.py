"""Drop-in mirror of the hot-path part of the reference's src/liftreg/layers/layers.py: proj_layer (:159-236).
The conv / FC / smoothing blocks of that file are stock cuDNN/cuBLAS modules and stay with the reference.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class proj_layer(nn.Module):
    """Differentiable batched DRR layer: x (B,d,w,h) -> (B,P,*out_shape)  (reference layers.py:159-192).

    Geometry as in the reference: emitter arc at 3.0*w (:170-172), detector int(resolution_scale*d) x
    int(resolution_scale*h) (:175-176), y normalised by w (not w-1, :234), no mm->cm factor.
    The reference pre-computes a (P,rd,rh,w,3) grid in its constructor and repeats it B times per forward
    (:186); here the geometry is recomputed in registers per sample, `.dx` is computed once, and `.grids` is
    materialised only if somebody reads it.  The trailing nearest-neighbour resize (:190) is torch's own.
    """

    def __init__(self, volume_spacing, resolution_scale, scan_range, proj_num, in_shape, out_shape, device):
        super(proj_layer, self).__init__()
        self.spacing = volume_spacing
        self.resolution_scale = resolution_scale
        self.sample_rate = [int(1), int(1), int(1)]
        self.out_shape = out_shape
        self.in_shape = tuple(int(s) for s in in_shape)
        self.device = torch.device(device)

        angle_half = scan_range / 2.
        self.poses_scale = np.ndarray((proj_num, 3), dtype=np.float64)
        self.poses_scale[:, 1] = 3.
        self.poses_scale[:, 0] = np.tan(np.linspace(-angle_half, angle_half, num=proj_num) / 180. * np.pi) * 3.
        self.poses_scale[:, 2] = np.linspace(-0.2, 0.2, num=proj_num)
        self.emi_poses = self.poses_scale * in_shape[1]
        self.proj_resolution = [int(in_shape[0] * self.resolution_scale), int(in_shape[2] * self.resolution_scale)]
        _, self.dx = ops.project_grid(self.emi_poses, self.proj_resolution, self.in_shape, self.spacing, self.device,
                                      ops.YNORM_W, flip=True, want_grid=False)
        self._grids = None

    @property
    def grids(self):
        """(P,rd,rh,w,3) sample grid in grid_sample order, as the reference stores it (:180)."""
        if self._grids is None:
            self._grids, _ = ops.project_grid(self.emi_poses, self.proj_resolution, self.in_shape, self.spacing,
                                              self.device, ops.YNORM_W, flip=True)
        return self._grids

    def forward(self, x):
        x_proj = ops.drr_project(x, self.emi_poses, self.proj_resolution, self.spacing, ops.YNORM_W, out_scale=1.0)
        if tuple(int(v) for v in self.out_shape) == tuple(x_proj.shape[2:]):
            return x_proj                                   # nearest resize to the same size is the identity
        return F.interpolate(x_proj, self.out_shape)        # :190 (nearest): the one stock-torch op left on this layer
