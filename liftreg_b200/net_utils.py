"""Drop-in mirror of the hot-path part of the reference's src/liftreg/utils/net_utils.py:
Bilinear (:9-56), identity_map (:59-87), not_normalized_identity_map (:90-110), gen_identity_map (:113-125).
The checkpoint helpers of that file (:127-235) are plain torch.save/load and stay with the reference.
"""
import numpy as np
import torch
from torch.nn import Module

from . import ops

dim = 3


class Bilinear(Module):
    """Spatial transformer in BCXYZ layout (reference net_utils.py:9-56), one fused sm_100a kernel.

    input1 (B,C,X,Y,Z) image, input2 (B,3,X,Y,Z) map in [-1,1] whose channel c addresses axis c.
    zero_boundary -> zeros padding else border (:21); using_scale rescales intensities [-1,1]->[0,1] before
    and back after sampling (:48-52); mode "bilinear" | "nearest" (:23).  Differentiable wrt both inputs.

    Host (CPU) float32 tensors, which tools/evaluate_dir_lab.py:217-222 passes, are staged through the GPU and
    returned on the host; nothing is ever computed on the CPU.
    """

    def __init__(self, zero_boundary=False, using_scale=True, mode="bilinear"):
        super(Bilinear, self).__init__()
        self.zero_boundary = 'zeros' if zero_boundary else 'border'
        self.using_scale = using_scale
        self.mode = mode

    def _run(self, input1, input2, using_scale):
        on_host = not input1.is_cuda
        if on_host:
            if not torch.cuda.is_available():
                raise RuntimeError("Bilinear needs a CUDA device: liftreg_b200 has no CPU path")
            input1, input2 = input1.cuda(non_blocking=True), input2.cuda(non_blocking=True)
        out = ops.warp(input1, input2, zero_boundary=(self.zero_boundary == 'zeros'), using_scale=using_scale,
                       mode=self.mode)
        return out.cpu() if on_host else out

    def forward_stn(self, input1, input2):
        """Sampling without the intensity rescale (reference :26-38)."""
        return self._run(input1, input2, False)

    def forward(self, input1, input2):
        return self._run(input1, input2, bool(self.using_scale))


def identity_map(sz, dtype=np.float32):
    """Normalised identity map (dim,*sz) on the current CUDA device (reference :59-87).
    3-D maps are generated on the device; 1-D / 2-D maps (unused by the resampling path) are built on the host
    exactly as the reference does and copied over."""
    nd = len(sz)
    if nd == 3:
        return ops.identity_map(sz, torch.device("cuda", torch.cuda.current_device()))
    if nd == 1:
        idm = np.mgrid[0:sz[0]]
    elif nd == 2:
        idm = np.mgrid[0:sz[0], 0:sz[1]]
    else:
        raise ValueError('Only dimensions 1-3 are currently supported for the identity map')
    idm = np.array(idm.astype(dtype))
    if nd == 1:
        idm = idm.reshape(1, sz[0])
    spacing = 1. / (np.array(sz) - 1)
    for d in range(nd):
        idm[d] *= spacing[d]
        idm[d] = idm[d] * 2 - 1
    return torch.from_numpy(idm.astype(np.float32)).cuda()


def not_normalized_identity_map(sz):
    """Voxel-index identity map (reference :90-110)."""
    nd = len(sz)
    if nd == 1:
        idm = np.mgrid[0:sz[0]]
    elif nd == 2:
        idm = np.mgrid[0:sz[0], 0:sz[1]]
    elif nd == 3:
        idm = np.mgrid[0:sz[0], 0:sz[1], 0:sz[2]]
    else:
        raise ValueError('Only dimensions 1-3 are currently supported for the identity map')
    return torch.from_numpy(idm.astype(np.float32)).cuda()


def gen_identity_map(img_sz, resize_factor=1., normalized=True):
    """Identity map for an (optionally resized) image size (reference :113-125)."""
    if isinstance(resize_factor, list):
        img_sz = [int(img_sz[i] * resize_factor[i]) for i in range(dim)]
    else:
        img_sz = [int(img_sz[i] * resize_factor) for i in range(dim)]
    return identity_map(img_sz) if normalized else not_normalized_identity_map(img_sz)
