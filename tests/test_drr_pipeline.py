"""Row f3: the dataset DRR loop of tools/preprocessingDRR.py:123-154 as a pipeline (liftreg_b200/drr_pipeline.py).
CPU tests inject the C oracle as the compute stage; the GPU test compares against the serial mirror calls."""
import os

import numpy as np
import pytest

from liftreg_b200 import drr_pipeline, sdct_projection_utils as sdct
from pipeline_host_stage import host_stage


def _make_dataset(root, n, shape, seed=7, dtype=np.float32):
    rs = np.random.RandomState(seed)
    ids = ["case%03d" % i for i in range(n)]
    for i in ids:
        for kind in ("target", "source"):
            hu = (rs.uniform(-1200, 600, shape)).astype(dtype)       # includes values below -1000 (clamped, sdct:8)
            np.save(os.path.join(root, "%s_%s.npy" % (i, kind)), hu)
    return ids


def _oracle_project(mu, poses, resolution, spacing):
    from oracle import c_oracle
    return c_oracle.drr_forward(mu, poses, resolution, spacing)


@pytest.mark.parametrize("depth", [1, 3])
def test_pipeline_host_stage_matches_serial_loop(tmp_path, depth):
    from oracle import c_oracle
    shape = (10, 12, 9)
    pre, out = str(tmp_path / "pre"), str(tmp_path / "drr")
    os.makedirs(pre)
    ids = _make_dataset(pre, 5, shape)
    poses = drr_pipeline.generate_drr_dataset(pre, ids, out, scan_range=60.0, scan_num=3, receptor_size=(14, 13), depth=depth,
                                              stage_factory=host_stage(_oracle_project))
    # poses.npy: sdct:139-155 (float64, voxel units), written once (:154)
    want_poses = sdct._wrapper_poses_scale(60.0, 3, 3.5) * shape[1]
    assert poses.dtype == np.float64 and np.array_equal(poses, want_poses)
    assert np.array_equal(np.load(os.path.join(out, "poses.npy")), want_poses)
    assert sorted(os.listdir(out)) == sorted(["poses.npy"] + ["%s_%s_proj.npy" % (i, k) for i in ids for k in ("target", "source")])
    for i in ids:
        for kind in ("target", "source"):
            vol = np.flip(np.load(os.path.join(pre, "%s_%s.npy" % (i, kind))), axis=1)             # :134-135
            ref = c_oracle.drr_forward(sdct.calc_relative_atten_coef(vol), want_poses, (14, 13), (2.2, 2.2, 2.2))
            got = np.load(os.path.join(out, "%s_%s_proj.npy" % (i, kind)))
            assert got.dtype == np.float32 and got.shape == (3, 14, 13)
            assert np.array_equal(got, ref)


def test_pipeline_geo_csv_default_detector_and_int16(tmp_path):
    shape = (8, 6, 10)
    pre, out = str(tmp_path / "pre"), str(tmp_path / "drr")
    os.makedirs(pre)
    ids = _make_dataset(pre, 2, shape, dtype=np.int16)
    geo = tmp_path / "geo.csv"
    geo.write_text("x,y,z\n-100.0,46.2,-3.0\n0.0,46.2,0.0\n100.0,46.2,3.0\n")
    seen = []

    def fn(mu, poses, resolution, spacing):
        seen.append((mu.dtype, tuple(resolution)))
        return _oracle_project(mu, poses, resolution, spacing)

    poses = drr_pipeline.generate_drr_dataset(pre, ids, out, geo_path=str(geo), stage_factory=host_stage(fn))
    want = np.array([[-100.0, 46.2, -3.0], [0.0, 46.2, 0.0], [100.0, 46.2, 3.0]]) / (2.2, 2.2, 2.2)   # sdct:162-163
    assert np.array_equal(poses, want)
    assert all(dt == np.float32 and res == (12, 15) for dt, res in seen)      # int(1.5*d) x int(1.5*h), sdct:149-151
    assert np.load(os.path.join(out, "case001_source_proj.npy")).shape == (3, 12, 15)


def test_pipeline_errors_surface(tmp_path):
    pre, out = str(tmp_path / "pre"), str(tmp_path / "drr")
    os.makedirs(pre)
    ids = _make_dataset(pre, 2, (6, 6, 6))
    np.save(os.path.join(pre, "case001_source.npy"), np.zeros((5, 6, 6), np.float32))     # shape mismatch inside the run
    with pytest.raises(ValueError):
        drr_pipeline.generate_drr_dataset(pre, ids, out, scan_range=60.0, scan_num=2, stage_factory=host_stage(_oracle_project))
    with pytest.raises(FileNotFoundError):
        drr_pipeline.generate_drr_dataset(pre, ["missing"], out, scan_range=60.0, scan_num=2, stage_factory=host_stage(_oracle_project))
    with pytest.raises(ValueError):
        drr_pipeline.generate_drr_dataset(pre, ids[:1], out, stage_factory=host_stage(_oracle_project))   # no geometry given
    assert drr_pipeline.generate_drr_dataset(pre, [], out, scan_range=60.0, scan_num=2, stage_factory=host_stage(_oracle_project)) is None


@pytest.mark.gpu
def test_pipeline_cuda_matches_serial_mirror_calls(tmp_path):
    import torch
    assert torch.cuda.is_available()
    shape = (24, 20, 28)
    pre, out = str(tmp_path / "pre"), str(tmp_path / "drr")
    os.makedirs(pre)
    ids = _make_dataset(pre, 7, shape)
    poses = drr_pipeline.generate_drr_dataset(pre, ids, out, scan_range=60.0, scan_num=4, depth=3)
    for i in ids:
        for kind in ("target", "source"):
            vol = np.flip(np.load(os.path.join(pre, "%s_%s.npy" % (i, kind))), axis=1)
            ref, ref_poses = sdct.calculate_projection_wraper(sdct.calc_relative_atten_coef(vol), 60.0, 4, (2.2, 2.2, 2.2))
            got = np.load(os.path.join(out, "%s_%s_proj.npy" % (i, kind)))
            assert np.array_equal(got, ref)              # same kernel, same bits: batching and device-side HU->mu change nothing
            assert np.array_equal(poses, ref_poses)


def test_pipeline_case_sharding_covers_every_case_once(tmp_path):
    """shard=(rank, world): the ranks' case lists partition the ids, rank 0 alone writes poses.npy, and the union of
    the files equals the unsharded run."""
    shape = (8, 9, 7)
    pre, out_a, out_b = str(tmp_path / "pre"), str(tmp_path / "a"), str(tmp_path / "b")
    os.makedirs(pre)
    ids = _make_dataset(pre, 5, shape)
    kw = dict(scan_range=60.0, scan_num=2, receptor_size=(10, 11), stage_factory=host_stage(_oracle_project))
    drr_pipeline.generate_drr_dataset(pre, ids, out_a, **kw)
    drr_pipeline.generate_drr_dataset(pre, ids, out_b, shard=(1, 3), **kw)
    assert sorted(os.listdir(out_b)) == sorted("%s_%s_proj.npy" % (i, k) for i in ids[1::3] for k in ("target", "source"))
    drr_pipeline.generate_drr_dataset(pre, ids, out_b, shard=(0, 3), **kw)
    drr_pipeline.generate_drr_dataset(pre, ids, out_b, shard=(2, 3), **kw)
    assert sorted(os.listdir(out_a)) == sorted(os.listdir(out_b))
    for f in os.listdir(out_a):
        assert np.array_equal(np.load(os.path.join(out_a, f)), np.load(os.path.join(out_b, f)))
    with pytest.raises(ValueError):
        drr_pipeline.generate_drr_dataset(pre, ids, out_b, shard=(3, 3), **kw)
    # more ranks than cases: the surplus rank writes nothing and does not fail
    out_c = str(tmp_path / "c")
    drr_pipeline.generate_drr_dataset(pre, ids[:1], out_c, shard=(1, 2), **kw)
    assert os.listdir(out_c) == []
