#!/usr/bin/env python
"""Benchmark of the LiftReg resampling hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|drr_cfg1|drr_cfg4]

One "step" = one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE.json
configs[1]: backprojection of 4 limited-angle 256^2 DRRs into a 160^3 volume + displacement warp of the moving
160^3 CT, batch 1 per GPU.  Units per step = 4*160^3 voxel-samples (backprojection) + 160^3 voxels (warp).

  value   device-resident throughput: inputs already in HBM, CUDA-graph replay of the two kernels, CUDA events on
          the launching stream, max over ranks.  Consecutive steps use different buffer sets (rotation of R sets,
          R * 148 MB >> 126 MB L2) so every step reads cold inputs.
  e2e     same metric through the host-buffer C-ABI calls (lr_backproject_forward_host + lr_warp_forward_host):
          pinned host inputs -> H2D -> kernels -> D2H of both results, every step, synchronous.
  roofline  dominant kernel, algorithmic bytes / mean launch duration (CUDA events, same rotation) vs measured HBM peak.
  cpu_baseline  the reference's CPU path (oracle/torch_port.py, op-for-op torch restatement) on this box's host cores.

N > 1 (torchrun): every rank owns one batch item (data-parallel sharding of the batch, no data-path collective),
so per-GPU work is fixed ("weak"); value = all ranks' units / max-over-ranks time.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "ray+voxel samples/s for DRR, backproj, warp; HBM GB/s vs peak at 1/2/4/8 B200"
UNIT = "samples/s"
VOL = (160, 160, 160)
DET = (256, 256)
P = 4
ROTATION = 8


def _peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(kernel):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML (what nvidia-smi reads) every 20 ms in a thread."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.util = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append(mhz)
                self.util.append(util)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "sm_mhz_min": int(min(self.samples)), "sm_mhz_max_seen": int(max(self.samples))}


# ----------------------------------------------------------------------------------------------- inputs
def make_inputs():
    """Seeded synthetic cfg2 inputs on the host (numpy fp32)."""
    from liftreg_b200 import synthetic
    hu = synthetic.ct_phantom(VOL)
    moving = synthetic.hu_to_unit(hu)[None, None]                          # (1,1,160,160,160) in [-1,1]
    phi = (synthetic.smooth_displacement(VOL) + synthetic.identity_map_np(VOL))[None]   # (1,3,...) model :68
    poses = synthetic.wrapper_poses(60.0, P, VOL[1])
    return hu, moving, phi.astype(np.float32), poses


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference_step_fn(moving, phi, target_proj, poses32, nz=None):
    """The reference's CPU path for cfg2 (oracle/torch_port.py): backprojection grid cached like the model does
    (:85-87), then Bilinear warp.  nz < 160 restricts both outputs to their first nz axial planes (a bounded sample:
    grid_sample takes an output grid of any extent over the full input)."""
    import torch
    from oracle import torch_port
    nz = VOL[0] if nz is None else int(nz)
    grids = torch_port.backproj_grid(poses32[None], VOL, DET).permute(0, 1, 3, 4, 5, 2)
    t_proj, t_moving, t_phi = torch.from_numpy(target_proj), torch.from_numpy(moving), torch.from_numpy(phi)
    if nz < VOL[0]:
        grids = grids[:, :, :nz].contiguous()
        t_phi = t_phi[:, :, :nz].contiguous()

    def step():
        lifted = torch_port.backproject(t_proj, grids)
        warped = torch_port.warp(t_moving, t_phi, zero_boundary=True, using_scale=True)
        return lifted, warped

    units = (P + 1) * nz * VOL[1] * VOL[2]
    return step, units, nz


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (torch CPU ops, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    torch.set_num_threads(os.cpu_count())
    from liftreg_b200 import synthetic
    from oracle import torch_port
    hu, moving, phi, poses = make_inputs()
    poses32 = poses.astype(np.float32)
    mu = synthetic.hu_to_mu(hu)
    proj = torch_port.drr(mu, poses, DET, (2.2, 2.2, 2.2))                 # reference DRR on CPU makes the inputs
    target_proj = synthetic.normalise_projection(proj)[None]
    step, units, nz = cpu_reference_step_fn(moving, phi, target_proj, poses32)
    step()
    t0 = time.perf_counter(); step(); t_full = time.perf_counter() - t0
    budget = 120.0
    warm = max(1, min(args.warmup, 3))
    sample = "full cfg2 step (B=1)"
    if (args.steps + warm) * t_full > budget:                              # bounded sample: an axial slab of both outputs
        nz = int(min(VOL[0], max(4, VOL[0] * budget / ((args.steps + warm) * t_full))))
        step, units, nz = cpu_reference_step_fn(moving, phi, target_proj, poses32, nz)
        sample = "first %d of %d axial planes of both outputs per step (bounded to ~%.0f s total)" % (nz, VOL[0], budget)
    for _ in range(warm):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter(); step(); times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = units / (ms * 1e-3)
    sample += "; torch CPU backprojection (cached grid) + Bilinear warp via oracle/torch_port.py"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: backprojection 4x256^2 -> 160^3 + warp 160^3 (zeros, using_scale), batch 1 per GPU",
                   "units_per_step": units, "full_step_s": t_full,
                   "host": "CPU only (torch %s, %d threads, %d cpus)" % (torch.__version__, torch.get_num_threads(), os.cpu_count())},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)
    return 0


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from liftreg_b200 import _native, ops, synthetic
    from liftreg_b200 import sdct_projection_utils as sdct

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _native.lib()

    hu, moving, phi, poses = make_inputs()
    poses32 = np.ascontiguousarray(poses.astype(np.float32))
    mu = synthetic.hu_to_mu(hu)
    proj = sdct.calculate_projection(mu, poses, DET, [1, 1, 1], (2.2, 2.2, 2.2), dev)     # our DRR makes the inputs
    target_proj = synthetic.normalise_projection(proj)[None]                             # (1,4,256,256)

    nv = VOL[0] * VOL[1] * VOL[2]
    units_bp, units_warp = P * nv, nv
    units = units_bp + units_warp
    bytes_bp = 4 * units_bp + 4 * P * DET[0] * DET[1]          # 4 B write / voxel-sample + projections once
    bytes_warp = 20 * units_warp                               # 4 out + 12 phi + 4 image per voxel

    # R rotating buffer sets (distinct HBM) -> every step touches cold data: R*148 MB >> 126 MB L2
    R = ROTATION
    sets = []
    for r in range(R):
        sets.append(dict(proj=torch.from_numpy(target_proj).to(dev), moving=torch.from_numpy(moving).to(dev),
                         phi=torch.from_numpy(phi).to(dev),
                         lifted=torch.empty((1, P) + VOL, device=dev), warped=torch.empty((1, 1) + VOL, device=dev)))
    pp = ops._fp(poses32)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    def k_backproject(s, st):
        _native.check(lib.lr_backproject_forward(vp(s["proj"]), pp, 1, P, DET[0], DET[1], VOL[0], VOL[1], VOL[2],
                                                 vp(s["lifted"]), P * nv, nv, st), "lr_backproject_forward")

    def k_warp(s, st):
        _native.check(lib.lr_warp_forward(vp(s["moving"]), vp(s["phi"]), 1, 1, VOL[0], VOL[1], VOL[2], 0, 0, 1, 0,
                                          vp(s["warped"]), st), "lr_warp_forward")

    def step(s, st):
        k_backproject(s, st)
        k_warp(s, st)

    stream = torch.cuda.Stream(device=dev)
    side = torch.cuda.Stream(device=dev)       # second branch of the step graph

    def step_forked(s, st):
        """One step as two parallel graph branches: the backprojection and the warp of a step are independent (in the
        model they are separated by the encoder), so the step forks onto a side stream and joins before the next."""
        fork = torch.cuda.Event(); fork.record(stream); side.wait_event(fork)
        k_backproject(s, ctypes.c_void_p(side.cuda_stream))
        k_warp(s, st)
        join = torch.cuda.Event(); join.record(side); stream.wait_event(join)

    def capture(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            st = ctypes.c_void_p(stream.cuda_stream)
            fn(sets[0], st)                                     # module load / first-launch outside capture
            stream.synchronize(); side.synchronize()
            with torch.cuda.graph(g, stream=stream):
                for r in range(R):
                    fn(sets[r], st)
        return g

    def run_steps(graph, fn, n):
        """exactly n steps on `stream`: graph replays of R steps + eager remainder."""
        with torch.cuda.stream(stream):
            for _ in range(n // R):
                graph.replay()
            st = ctypes.c_void_p(stream.cuda_stream)
            for r in range(n % R):
                fn(sets[r], st)

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(graph, fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            e0.record(stream)
        run_steps(graph, fn, n)
        with torch.cuda.stream(stream):
            e1.record(stream)
        torch.cuda.synchronize(); barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    _native.launch_count_reset()
    g_step = capture(step_forked)
    launches_per_step = (_native.launch_count() - 2) // R       # counted at capture time (replays re-issue them)
    step = step_forked
    g_bp, g_warp = capture(k_backproject), capture(k_warp)

    sampler = ClockSampler(local)
    run_steps(g_step, step, max(args.warmup, 3))                 # W untimed warm-up steps
    torch.cuda.synchronize()
    sampler.start()
    total_ms = timed(g_step, step, args.steps)                   # EXACTLY K steps
    ms_per_step = total_ms / args.steps
    value = world * units / (ms_per_step * 1e-3)

    # per-kernel mean launch duration, same rotation, same stream (enough launches to sample clocks under load)
    n_k = max(args.steps, 4000)
    run_steps(g_bp, k_backproject, 64); us_bp = 1e3 * timed(g_bp, k_backproject, n_k) / n_k
    run_steps(g_warp, k_warp, 64); us_warp = 1e3 * timed(g_warp, k_warp, n_k) / n_k
    n_long = int(min(200000, max(n_k, 1.0e6 / max(us_bp + us_warp, 1.0))))   # ~1 s of back-to-back steps for the clock record
    sustained_ms = timed(g_step, step, n_long) / n_long

    # ---- DRR forward (BASELINE configs[0] geometry: 160^3, 4 views / 60 deg, 240^2 and 256^2 detectors), reported
    # next to the headline because the north star names all three operators.  Nominal ray-samples = P*rd*rh*w.
    drr_extra = {}
    if rank == 0:
        mus = [torch.from_numpy(mu)[None].to(dev) for _ in range(R)]
        sp3 = np.array([2.2, 2.2, 2.2], np.float32)
        poses64 = np.ascontiguousarray(poses, np.float64)
        for det in ((240, 240), (256, 256)):
            outs = [torch.empty((1, P) + det, device=dev) for _ in range(R)]

            def k_drr(sidx, stx, det=det, outs=outs):
                _native.check(lib.lr_drr_forward(vp(mus[sidx]), 1, VOL[0], VOL[1], VOL[2], ops._dp(poses64), 1, P, det[0], det[1],
                                                 ops._fp(sp3), 0, ctypes.c_float(0.1), vp(outs[sidx]), stx), "lr_drr_forward")

            g_drr = torch.cuda.CUDAGraph()
            with torch.cuda.stream(stream):
                stx = ctypes.c_void_p(stream.cuda_stream)
                k_drr(0, stx); stream.synchronize()
                with torch.cuda.graph(g_drr, stream=stream):
                    for r in range(R):
                        k_drr(r, stx)
            n_d = 40 * R
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                for _ in range(4):
                    g_drr.replay()
                e0.record(stream)
                for _ in range(n_d // R):
                    g_drr.replay()
                e1.record(stream)
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / n_d
            nominal = P * det[0] * det[1] * VOL[1]
            comp_bytes = 4 * nv + 4 * P * det[0] * det[1]
            drr_extra["%dx%d" % det] = {
                "us": us, "nominal_ray_samples": nominal, "samples_per_s": nominal / us * 1e6,
                "compulsory_bytes": comp_bytes, "gbps_compulsory": comp_bytes / us * 1e-3,
                "gather_model_gbps": 16.0 * nominal / us * 1e-3}

    # ---- the two streaming kernels at batch 8 (BASELINE configs[2]: the full forward runs them on 8 items per launch):
    # same kernels, same entry points, 2 rotating buffer sets of 1.2 GB each; reported per launch and per item
    batch8 = None
    if rank == 0:
        B8 = 8
        b_proj = [torch.from_numpy(target_proj).to(dev).repeat(B8, 1, 1, 1) for _ in range(2)]
        b_mov = [torch.from_numpy(moving).to(dev).repeat(B8, 1, 1, 1, 1) for _ in range(2)]
        b_phi = [torch.from_numpy(phi).to(dev).repeat(B8, 1, 1, 1, 1) for _ in range(2)]
        b_lift = [torch.empty((B8, P) + VOL, device=dev) for _ in range(2)]
        b_warp = [torch.empty((B8, 1) + VOL, device=dev) for _ in range(2)]

        def k8_bp(i, stx):
            _native.check(lib.lr_backproject_forward(vp(b_proj[i]), pp, B8, P, DET[0], DET[1], VOL[0], VOL[1], VOL[2],
                                                     vp(b_lift[i]), P * nv, nv, stx), "lr_backproject_forward")

        def k8_warp(i, stx):
            _native.check(lib.lr_warp_forward(vp(b_mov[i]), vp(b_phi[i]), B8, 1, VOL[0], VOL[1], VOL[2], 0, 0, 1, 0,
                                              vp(b_warp[i]), stx), "lr_warp_forward")

        batch8 = {}
        for name, fn, nbytes, nunits in (("backproject_forward_kernel", k8_bp, B8 * bytes_bp, B8 * units_bp),
                                         ("warp_forward_kernel", k8_warp, B8 * bytes_warp, B8 * units_warp)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                stx = ctypes.c_void_p(stream.cuda_stream)
                for i in range(4):
                    fn(i % 2, stx)
                e0.record(stream)
                for i in range(200):
                    fn(i % 2, stx)
                e1.record(stream)
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / 200
            batch8[name] = {"us_per_launch": us, "us_per_item": us / B8, "bytes": nbytes, "gbps": nbytes / us * 1e-3,
                            "units_per_s": nunits / us * 1e6}
        del b_proj, b_mov, b_phi, b_lift, b_warp

    # ---- PCA-subspace decode (model :102): streams the 2.75 GB basis once
    pca_extra = None
    if rank == 0:
        K = 56
        basis = torch.empty((3 * nv, K), device=dev).normal_(0, 1e-3)
        pmean = torch.zeros(3 * nv, device=dev)
        coefs = torch.randn(1, K, device=dev)
        pouts = [torch.empty((1, 3 * nv), device=dev) for _ in range(2)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            stx = ctypes.c_void_p(stream.cuda_stream)
            for i in range(3):
                _native.check(lib.lr_pca_decode(vp(coefs), vp(basis), vp(pmean), 1, K, 3 * nv, 1, VOL[0], VOL[1], VOL[2], vp(pouts[i % 2]), stx), "lr_pca_decode")
            e0.record(stream)
            for i in range(20):
                _native.check(lib.lr_pca_decode(vp(coefs), vp(basis), vp(pmean), 1, K, 3 * nv, 1, VOL[0], VOL[1], VOL[2], vp(pouts[i % 2]), stx), "lr_pca_decode")
            e1.record(stream)
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 20
        pbytes = 4 * 3 * nv * K + 8 * 3 * nv
        pca_extra = {"us": us, "bytes": pbytes, "gbps": pbytes / us * 1e-3, "workload": "B=1, K=56, N=3*160^3 (+mean, +identity)"}
        del basis, pmean, pouts

    # ---- DRR end to end: calculate_projection's numpy-in / numpy-out contract (sdct:59-100) as preprocessingDRR.py
    # calls it (H2D of the 160^3 volume, kernel, D2H of the 4 x 240^2 images, synchronous), per call
    drr_e2e_ms = None
    if rank == 0:
        for _ in range(2):
            sdct.calculate_projection(mu, poses, (240, 240), [1, 1, 1], (2.2, 2.2, 2.2), dev)
        t0 = time.perf_counter()
        for _ in range(10):
            sdct.calculate_projection(mu, poses, (240, 240), [1, 1, 1], (2.2, 2.2, 2.2), dev)
        drr_e2e_ms = 1e2 * (time.perf_counter() - t0)

    # ---- e2e through the host-buffer C-ABI (pinned host memory, H2D + kernels + D2H every step)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_proj, h_moving, h_phi = pin(target_proj), pin(moving), pin(phi)
    h_lifted = torch.empty((1, P) + VOL).pin_memory()
    h_warped = torch.empty((1, 1) + VOL).pin_memory()
    ws_bp = torch.empty(lib.lr_backproject_forward_host_workspace_bytes(1, P, DET[0], DET[1], *VOL), dtype=torch.uint8, device=dev)
    ws_w = torch.empty(lib.lr_warp_forward_host_workspace_bytes(1, 1, *VOL), dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(stream.cuda_stream)

    def e2e_step():
        _native.check(lib.lr_backproject_forward_host(vp(h_proj), pp, 1, P, DET[0], DET[1], VOL[0], VOL[1], VOL[2],
                                                      vp(h_lifted), vp(ws_bp), ws_bp.numel(), st), "backproject host")
        _native.check(lib.lr_warp_forward_host(vp(h_moving), vp(h_phi), 1, 1, VOL[0], VOL[1], VOL[2], 0, 0, 1, 0,
                                               vp(h_warped), vp(ws_w), ws_w.numel(), st), "warp host")

    n_e2e = max(3, min(args.steps, 50))
    for _ in range(3):
        e2e_step()
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()                                               # each call synchronises its stream before returning
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    # the two calls are independent: issued from two host threads on two streams they overlap on the full-duplex link
    from concurrent.futures import ThreadPoolExecutor
    stream2 = torch.cuda.Stream(device=dev)
    st2 = ctypes.c_void_p(stream2.cuda_stream)
    pool = ThreadPoolExecutor(2)

    def e2e_bp():
        torch.cuda.set_device(local)      # the current device is per host thread; pool threads start on device 0
        _native.check(lib.lr_backproject_forward_host(vp(h_proj), pp, 1, P, DET[0], DET[1], VOL[0], VOL[1], VOL[2],
                                                      vp(h_lifted), vp(ws_bp), ws_bp.numel(), st), "backproject host")

    def e2e_warp():
        torch.cuda.set_device(local)
        _native.check(lib.lr_warp_forward_host(vp(h_moving), vp(h_phi), 1, 1, VOL[0], VOL[1], VOL[2], 0, 0, 1, 0,
                                               vp(h_warped), vp(ws_w), ws_w.numel(), st2), "warp host")

    def e2e_step_concurrent():
        fa, fb = pool.submit(e2e_bp), pool.submit(e2e_warp)
        fa.result(); fb.result()

    for _ in range(3):
        e2e_step_concurrent()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step_concurrent()
    torch.cuda.synchronize()
    e2e_conc_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    pool.shutdown()
    if world > 1:
        t = torch.tensor([e2e_ms, e2e_conc_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_conc_ms = float(t[0].item()), float(t[1].item())
    clocks = sampler.stop()
    h2d = 4 * (h_proj.numel() + h_moving.numel() + h_phi.numel())
    d2h = 4 * (h_lifted.numel() + h_warped.numel())
    # the e2e outputs are the same bits as the device-resident ones
    assert torch.equal(h_warped, sets[0]["warped"].cpu()) and torch.equal(h_lifted, sets[0]["lifted"].cpu())

    # ---- CPU baseline (rank 0, N=1 only): the reference's torch path on the host cores, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        cstep, cunits, _ = cpu_reference_step_fn(moving, phi, target_proj, poses32)
        cstep()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); lifted_ref, warped_ref = cstep(); ts.append(time.perf_counter() - t0)
        cpu_baseline = {"value": cunits / min(ts), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": "full cfg2 step (B=1) x3 after 1 warm-up, best; oracle/torch_port.py = the reference's "
                                  "torch CPU ops (cached backprojection grid, Bilinear warp)",
                        "ms_per_step": 1e3 * min(ts), "host_cpus": os.cpu_count()}
        # parity of the benchmarked outputs against the CPU path, in the same run
        def rl2(a, b):
            a = a.double(); b = b.double()
            return float(((a - b).norm() / b.norm()).item())
        cpu_baseline["parity_rel_l2"] = {"backproject": rl2(sets[0]["lifted"].cpu(), lifted_ref),
                                         "warp": rl2(sets[0]["warped"].cpu(), warped_ref)}

    if rank == 0:
        peak, peak_src = _peak_hbm()
        kern = {
            "backproject_forward_kernel": {"us": us_bp, "bytes": bytes_bp, "gbps": bytes_bp / us_bp * 1e-3,
                                           "units_per_s": units_bp / us_bp * 1e6},
            "warp_forward_kernel": {"us": us_warp, "bytes": bytes_warp, "gbps": bytes_warp / us_warp * 1e-3,
                                    "units_per_s": units_warp / us_warp * 1e6},
        }
        for k in kern.values():
            k["frac_of_hbm_peak"] = k["gbps"] / peak
        dom = max(kern, key=lambda k: kern[k]["us"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: backprojection 4x256^2 -> 160^3 + warp 160^3 (zeros, using_scale), batch 1 per GPU",
                       "units_per_step_per_gpu": units, "parallelism": "batch-sharded dp%d, no collective" % world,
                       "l2": "rotating %d buffer sets (%.0f MB) > 126 MB L2; no flush kernel" % (R, R * 148.5),
                       "launch": "CUDA graph of %d steps replayed (remainder launched eagerly); inside a step the two "
                                 "independent kernels run as parallel graph branches on two streams and join" % R},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbps"], "peak": peak, "unit": "GB/s",
                         "frac": kern[dom]["gbps"] / peak, "traffic": _traffic(dom), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kern[dom]["bytes"], "us_per_launch": kern[dom]["us"]},
            "kernels": kern,
            "drr_forward_cfg1": {k: dict(v, frac_of_hbm_peak_compulsory=v["gbps_compulsory"] / peak) for k, v in drr_extra.items()},
            "drr_calculate_projection_e2e_ms": drr_e2e_ms,
            "pca_decode": dict(pca_extra, frac_of_hbm_peak=pca_extra["gbps"] / peak) if pca_extra else None,
            "batch8_cfg3": ({k: dict(v, frac_of_hbm_peak=v["gbps"] / peak) for k, v in batch8.items()} if batch8 else None),
            "sustained_ms_per_step": sustained_ms,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": world * units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": n_e2e,
                    "api": "lr_backproject_forward_host then lr_warp_forward_host, one host thread (pinned host buffers, "
                           "synchronous calls)",
                    "two_host_threads": {"value": world * units / (e2e_conc_ms * 1e-3), "ms_per_step": e2e_conc_ms,
                                         "note": "same two calls issued concurrently from two host threads / streams"}},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def _reserve_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version line on the
    first communicator when NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the whole run and
    the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
