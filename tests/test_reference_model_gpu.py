"""BASELINE configs[2] (cfg 3) with the REAL reference model: `LiftRegDeformSubspaceBackproj.model` from the reference's
own sources (oracle/_ref/, placed there by oracle/vendor_reference.py; git-ignored, travels to the GPU box) runs its full
forward at 160^3, batch 8, first un-patched on CUDA (stock F.grid_sample / torch.cat / F.linear), then after
`liftreg_b200.dropin.install()` with the same weights and inputs.  Outputs must agree to the north-star tolerance
(1e-5 relative L2 per volume) in both numerics modes."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

SCRIPT = r'''
import json, sys
sys.dont_write_bytecode = True
sys.path.insert(0, %(ref)r); sys.path.insert(0, %(root)r)
import numpy as np
np.float = float                                   # removed alias the reference still uses (sdct:141)
import torch
torch.backends.cudnn.deterministic = True          # same conv algorithm for both runs
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from liftreg_b200 import synthetic, _native
import liftreg_b200.sdct_projection_utils as ours_sdct

shape, det, B, P, K = (160, 160, 160), (256, 256), %(B)d, 4, 56
N = 3 * shape[0] * shape[1] * shape[2]
rs = np.random.RandomState(0)
base = (rs.standard_normal((K, 65536)) * 1e-3).astype(np.float32)
fake = {"pca_vectors.npy": np.tile(base, (1, N // 65536 + 1))[:, :N], "pca_mean.npy": np.zeros(N, np.float32)}
real_load = np.load
np.load = lambda path, *a, **k: fake[path.split("/")[-1]] if path.split("/")[-1] in fake else real_load(path, *a, **k)

import liftreg.models.LiftRegDeformSubspaceBackproj as M      # the reference's own module
dev = torch.device("cuda:0")
opt = {"drr_feature_num": P, "latent_dim": K, "pca_path": "/synthetic"}
torch.manual_seed(2021)
ref = M.model(shape, opt).to(dev).eval()

hu = synthetic.ct_phantom(shape)
poses = synthetic.wrapper_poses(60.0, P, shape[1])
proj = ours_sdct.calculate_projection(synthetic.hu_to_mu(hu), poses, det, [1, 1, 1], (2.2, 2.2, 2.2), dev)
tp0 = synthetic.normalise_projection(proj)
unit = synthetic.hu_to_unit(hu)
moving = np.stack([np.roll(unit, 3 * b, axis=b %% 3) for b in range(B)])[:, None]
target_proj = np.stack([np.roll(tp0, 2 * b, axis=1 + b %% 2) * (1.0 - 0.05 * b) for b in range(B)])
inp = {"source": torch.from_numpy(moving).to(dev), "target": torch.from_numpy(moving[::-1].copy()).to(dev),
       "target_proj": torch.from_numpy(target_proj.astype(np.float32)).to(dev),
       "target_poses": torch.from_numpy(np.repeat(poses[None], B, 0).astype(np.float32))}
def timed_forward(model):
    """ms of one more forward (the first call built caches / picked conv algorithms), CUDA events."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    with torch.no_grad():
        model(inp)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)

with torch.no_grad():
    out_ref = ref(inp)
want = {k: out_ref[k].float().cpu() for k in ("warped", "phi", "params", "pca_coefs")}
ms_reference_cuda = min(timed_forward(ref) for _ in range(3))
state = ref.state_dict()
del ref, out_ref
torch.cuda.empty_cache()

import liftreg_b200.dropin as dropin
patched = dropin.install()
assert "liftreg.models.LiftRegDeformSubspaceBackproj.model._estimate_flow" in patched
res = {}
for mode in ("exact", "fast"):
    _native.set_numerics(mode)
    _native.launch_count_reset()
    m = M.model(shape, opt).to(dev).eval()              # now built from the B200 Bilinear / gen_identity_map
    m.load_state_dict(state)
    with torch.no_grad():
        out = m(inp)
    torch.cuda.synchronize()
    r = {"launches": _native.launch_count(), "basis_contiguous": bool(m.pca_vectors.is_contiguous()),
         "ms_forward": min(timed_forward(m) for _ in range(3)), "ms_forward_reference_cuda": ms_reference_cuda}
    xin = torch.randn((B, 1 + P) + shape, device=dev)      # the conv encoder + FC alone (the reference's own modules, both runs)
    def encoder_ms():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        with torch.no_grad():
            x = xin
            for enc in m.encoders:
                x = enc(x)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    encoder_ms()
    r["ms_encoder_alone"] = min(encoder_ms() for _ in range(3))
    del xin
    for k, w in want.items():
        o = out[k].float().cpu()
        num = (o.double() - w.double()).reshape(B, -1).norm(dim=1)
        den = w.double().reshape(B, -1).norm(dim=1).clamp_min(1e-30)
        r[k] = float((num / den).max())
    res[mode] = r
    del m, out
    torch.cuda.empty_cache()
print("RESULT " + json.dumps(res))
'''


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "liftreg", "models", "LiftRegDeformSubspaceBackproj.py")),
                    reason="oracle/_ref not populated (run oracle/vendor_reference.py where /root/reference exists)")
def test_cfg3_real_reference_model_forward_with_dropin_batch8():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    code = SCRIPT % {"ref": REF, "root": ROOT, "B": 8}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    if os.environ.get("LIFTREG_B200_TIMING_OUT"):        # full-model forward times of this run (stock torch CUDA ops vs drop-in)
        with open(os.environ["LIFTREG_B200_TIMING_OUT"], "w") as f:
            json.dump(res, f, indent=1)
    for mode in ("exact", "fast"):
        r = res[mode]
        assert r["launches"] >= 3, r                     # backprojection, PCA decode, warp ran in the native library
        assert r["basis_contiguous"], r                  # the (N,K) view was made dense once
        for k in ("warped", "phi", "params", "pca_coefs"):
            assert r[k] <= 1e-5, (mode, k, r)
