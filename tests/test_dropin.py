"""liftreg_b200.dropin.install() rebinds the reference's own modules (only runnable where the reference exists, i.e. in
the build container; the GPU box has no /root/reference and skips)."""
import os
import subprocess
import sys

import pytest

REF = os.environ.get("LIFTREG_REF", "/root/reference/src")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "liftreg")), reason="reference tree not present")
def test_install_rebinds_reference_symbols():
    code = r'''
import sys
sys.dont_write_bytecode = True
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
np.float = float
# the model module binds names with `from ..utils.net_utils import Bilinear` at import time: import it FIRST
import liftreg.models.LiftRegDeformSubspaceBackproj as ref_model
import liftreg.utils.sdct_projection_utils as ref_sdct
import liftreg.utils.net_utils as ref_net
import liftreg.layers.layers as ref_layers
old_bilinear = ref_model.Bilinear
import liftreg_b200.dropin as dropin
import liftreg_b200.sdct_projection_utils as ours_sdct, liftreg_b200.net_utils as ours_net, liftreg_b200.layers as ours_layers
patched = dropin.install()
assert ref_sdct.calculate_projection is ours_sdct.calculate_projection
assert ref_sdct.calculate_projection_wraper is ours_sdct.calculate_projection_wraper
assert ref_sdct.backproj_grids_with_poses is ours_sdct.backproj_grids_with_poses
assert ref_net.Bilinear is ours_net.Bilinear and ref_net.gen_identity_map is ours_net.gen_identity_map
assert ref_layers.proj_layer is ours_layers.proj_layer
# names imported earlier by the model module are rebound too
assert ref_model.Bilinear is ours_net.Bilinear and ref_model.Bilinear is not old_bilinear
assert ref_model.gen_identity_map is ours_net.gen_identity_map
assert ref_model.backproj_grids_with_poses is ours_sdct.backproj_grids_with_poses
assert ref_model.model._estimate_flow is dropin._estimate_flow
assert len(patched) >= 17
assert dropin.install() == patched        # idempotent
print("OK", len(patched))
''' % (REF, ROOT)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout + res.stderr


def test_similarity_class_resolves_by_dotted_name_like_the_reference_does():
    """losses/SubspaceLoss.py:12 instantiates the similarity through utils/general.py:9-15 get_class(dotted name); the
    mirror must be reachable the same way (sim_class = 'liftreg_b200.losses.NCCLoss') and be an nn.Module."""
    import torch.nn as nn

    def get_class(kls):                    # the reference's lookup, restated: __import__ the module, walk the attributes
        parts = kls.split('.')
        m = __import__(".".join(parts[:-1]))
        for comp in parts[1:]:
            m = getattr(m, comp)
        return m

    sys.path.insert(0, ROOT)
    cls = get_class("liftreg_b200.losses.NCCLoss")
    assert isinstance(cls(), nn.Module) and callable(getattr(cls, "forward"))
    from liftreg_b200 import mermaid_utils
    assert callable(mermaid_utils.compute_warped_image_multiNC)
