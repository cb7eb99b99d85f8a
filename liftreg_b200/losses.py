"""Drop-in mirror of the similarity class of the reference's src/liftreg/layers/losses.py.

`NCCLoss` is what `losses/SubspaceLoss.py:12` instantiates by its dotted name ('layers.losses.NCCLoss') and what
`networks/RegistrationNet.py:210-212` uses as the validation score; both call it on the warp output.  The reference
module itself cannot be imported without `mermaid` (it imports mermaid.finite_differences at the top); this class needs
only the native library."""
import torch
import torch.nn as nn

from . import ops


class NCCLoss(nn.Module):
    """1 - mean_b NCC(input_b, target_b), reference layers/losses.py:14-29; two fused passes instead of ~10 kernels."""

    def forward(self, input, target):
        loss = ops.ncc_loss(input.reshape(input.shape[0], -1), target.reshape(target.shape[0], -1))
        assert not torch.isnan(loss), 'NCC loss is Nan.'          # losses.py:27
        return loss
