"""Drop-in mirrors of the reference's loss-side classes on the warp output: the similarity class of
src/liftreg/layers/losses.py and the subspace loss of src/liftreg/losses/SubspaceLoss.py.

`NCCLoss` is what `losses/SubspaceLoss.py:12` instantiates by its dotted name ('layers.losses.NCCLoss') and what
`networks/RegistrationNet.py:210-212` uses as the validation score; both call it on the warp output.  The reference
module itself cannot be imported without `mermaid` (it imports mermaid.finite_differences at the top); this class needs
only the native library."""
import math

import torch
import torch.nn as nn

from . import ops


class NCCLoss(nn.Module):
    """1 - mean_b NCC(input_b, target_b), reference layers/losses.py:14-29; one fused pass instead of ~10 kernels."""

    def forward(self, input, target):
        loss = ops.ncc_loss(input.reshape(input.shape[0], -1), target.reshape(target.shape[0], -1))
        assert not torch.isnan(loss), 'NCC loss is Nan.'          # losses.py:27
        return loss


def sigmoid_decay(ep, static=5, k=5):
    """reference utils/utils.py:93-107: 1 for the first `static` epochs, then k / (k + exp((ep - static) / k))."""
    if ep < static:
        return float(1.)
    ep = ep - static
    return float(k / (k + math.exp(ep / k)))


def _opt(opt, key, default):
    """The reference reads its settings through mermaid's ParameterDict (`opt[(key, default, doc)]`); a plain dict or
    None works here too."""
    if opt is None:
        return default
    try:
        return opt[(key, default, "")]
    except (KeyError, TypeError):
        return opt.get(key, default) if hasattr(opt, "get") else default


class SubspaceLoss(nn.Module):
    """reference losses/SubspaceLoss.py:9-67 (class `loss`): sim_factor * similarity + reg_factor(epoch) * regulariser.

    The similarity is `NCCLoss` (the reference's default sim_class, :12) and the regulariser the fused
    finite-difference kernel (`ops.diffusion_reg`), both differentiable.  `fd_boundary` selects the face rule of
    mermaid's finite differences ("linear" = FD_torch's default)."""

    def __init__(self, opt=None, sim=None, fd_boundary="linear"):
        super().__init__()
        self.sim_factor = 1.
        self.sim = sim if sim is not None else NCCLoss()
        self.initial_reg_factor = _opt(opt, 'initial_reg_factor', 10)
        self.min_reg_factor = _opt(opt, 'min_reg_factor', 1e-3)
        self.reg_factor_decay_from = _opt(opt, 'reg_factor_decay_from', 10)
        self.fd_boundary = fd_boundary

    def forward(self, input):
        sim_loss = self.sim(input["warped"], input["target"])
        reg_loss = self.compute_reg_loss(input["params"])
        total_loss = self.sim_factor * sim_loss + self.get_reg_factor(input["epoch"]) * reg_loss
        return {"total_loss": total_loss, "sim_loss": sim_loss.item(), "reg_loss": reg_loss.item()}

    def get_reg_factor(self, epoch):
        """:40-49"""
        return float(max(sigmoid_decay(epoch, static=self.reg_factor_decay_from, k=2) * self.initial_reg_factor,
                         self.min_reg_factor))

    def compute_reg_loss(self, affine_param):
        """:51-67"""
        return ops.diffusion_reg(affine_param, self.fd_boundary)


loss = SubspaceLoss          # the reference's class name inside losses/SubspaceLoss.py
