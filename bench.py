#!/usr/bin/env python
"""Benchmark of the LiftReg resampling hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-cpu-baseline] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input.  Headline workload = BASELINE.json
configs[1]: backprojection of 4 limited-angle 256^2 DRRs into a 160^3 volume + displacement warp of the moving
160^3 CT, batch 1 per GPU.  Units per step = 4*160^3 voxel-samples (backprojection) + 160^3 voxels (warp).

  value     device-resident throughput: inputs already in HBM, CUDA graphs of EXACTLY `steps` steps (no eagerly launched
            remainder), CUDA events on the launching stream, max over ranks.  Consecutive steps use different buffer
            sets (rotation of R sets, R * 148 MB >> 126 MB L2) so every step reads cold inputs.
  e2e       same metric through the host-buffer C-ABI (lr_backproject_forward_host_async + lr_warp_forward_host_async on
            two streams, one host thread, then lr_stream_synchronize): pinned host inputs -> H2D -> kernels -> D2H of both
            results, every step.  e2e.serial = the blocking calls one after the other (the reference's contract).
  roofline  dominant kernel, algorithmic bytes / mean launch duration (CUDA events, same rotation) vs measured HBM peak.
  cpu_baseline  the reference's CPU path (oracle/torch_port.py, op-for-op torch restatement) on this box's host cores.

N > 1 (torchrun).  `value` stays the batch-sharded form of cfg2 (one item per rank, no data-path collective, "weak") for
round-over-round continuity.  The SAME JSON line also carries north_star's partitioning, timed on every rank with the
sharded result checked against the unsharded one in the run (`sharded_parity`):
  cfg4_drr_view_sharded   BASELINE configs[3]: 512^3 CT, 64 views, 512^2 detector; views split across ranks, each kernel
                          writes into its slot of the gather buffer, NCCL all-gather INSIDE the CUDA-event window (strong).
  cfg5_training_ops       BASELINE configs[4]: backprojection + warp forward + warp d/dphi of a batch of 32 160^3 pairs,
                          batch-sharded (32/N items per rank); and one item z-slab-sharded over the N ranks (B < G).
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
NV = VOL[0] * VOL[1] * VOL[2]
WORKLOAD = "cfg2: backprojection 4x256^2 -> 160^3 + warp 160^3 (zeros, using_scale), batch 1 per GPU"


def workload_config(world, **extra):
    """The `config` object of the JSON line: identical keys in both arms."""
    cfg = {"workload": WORKLOAD, "units_per_step_per_gpu": (P + 1) * NV,
           "parallelism": "batch-sharded dp%d, no collective" % world,
           "l2": None, "launch": None, "numerics": None, "host": None}
    cfg.update(extra)
    return cfg


def _peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


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


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore its first-touch pinned allocations) to the CPUs of the NUMA node the GPU hangs
    off, so that at N > 1 the ranks' host<->device streams do not cross the socket interconnect.  Returns a note."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:              # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa: single node (no binding needed)"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "numa: bound to node %d (%d cpus)" % (node, len(allowed))
        return "numa: node %d has no allowed cpus (unbound)" % node
    except Exception as e:
        return "numa: unbound (%s)" % type(e).__name__


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
def cpu_reference_step_fn(moving, phi, target_proj, poses32, nz=None, items=1):
    """The reference's CPU path for cfg2 (oracle/torch_port.py): backprojection grid cached like the model does
    (:85-87), then Bilinear warp, on `items` batch items.  nz < 160 restricts both outputs to their first nz axial planes
    (a bounded sample: grid_sample takes an output grid of any extent over the full input)."""
    import torch
    from oracle import torch_port
    nz = VOL[0] if nz is None else int(nz)
    grids = torch_port.backproj_grid(poses32[None], VOL, DET).permute(0, 1, 3, 4, 5, 2)
    rep = lambda a, nd: torch.from_numpy(a).repeat(items, *([1] * nd)).contiguous()
    t_proj, t_moving, t_phi = rep(target_proj, 3), rep(moving, 4), rep(phi, 4)
    if nz < VOL[0]:
        grids = grids[:, :, :nz].contiguous()
        t_phi = t_phi[:, :, :nz].contiguous()

    def step():
        lifted = torch_port.backproject(t_proj, grids)
        warped = torch_port.warp(t_moving, t_phi, zero_boundary=True, using_scale=True)
        return lifted, warped

    units = items * (P + 1) * nz * VOL[1] * VOL[2]
    return step, units, nz


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (torch CPU ops, all host threads).  At
    --gpus N a step processes N items (what the N ranks of the other arm process per step), on rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    torch.set_num_threads(os.cpu_count())
    from liftreg_b200 import synthetic
    from oracle import torch_port
    items = max(1, args.gpus)
    hu, moving, phi, poses = make_inputs()
    poses32 = poses.astype(np.float32)
    mu = synthetic.hu_to_mu(hu)
    proj = torch_port.drr(mu, poses, DET, (2.2, 2.2, 2.2))                 # reference DRR on CPU makes the inputs
    target_proj = synthetic.normalise_projection(proj)[None]
    step, units, nz = cpu_reference_step_fn(moving, phi, target_proj, poses32, items=items)
    step()
    t0 = time.perf_counter(); step(); t_full = time.perf_counter() - t0
    budget = 150.0
    warm = max(0, args.warmup)
    sample = "full cfg2 step x %d item(s)" % items
    if (args.steps + warm) * t_full > budget:                              # bounded sample: an axial slab of both outputs
        nz = int(min(VOL[0], max(2, VOL[0] * budget / ((args.steps + warm) * t_full))))
        step, units, nz = cpu_reference_step_fn(moving, phi, target_proj, poses32, nz, items=items)
        sample = ("first %d of %d axial planes of both outputs, %d item(s) per step (bounded to ~%.0f s for %d+%d steps)"
                  % (nz, VOL[0], items, budget, warm, args.steps))
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
        "config": workload_config(items, launch="torch CPU operators, eager", l2="n/a (host)", numerics="ATen (the reference's own order)",
                                  host="CPU only (torch %s, %d threads, %d cpus)" % (torch.__version__, torch.get_num_threads(), os.cpu_count()),
                                  units_per_step=units, full_step_s=t_full),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)
    return 0


# ----------------------------------------------------------------------------------------------- B200 arm
class Bench:
    """Shared state of the B200 arm (device, streams, library, distributed group)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.numa = bind_to_gpu_numa_node(self.local)          # before any pinned allocation / thread pool
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        from liftreg_b200 import _native
        self.native = _native
        self.lib = _native.lib()
        self.stream = torch.cuda.Stream(device=self.dev)
        self.side = torch.cuda.Stream(device=self.dev)
        self.st = ctypes.c_void_p(self.stream.cuda_stream)

    @staticmethod
    def vp(t):
        return ctypes.c_void_p(t.data_ptr())

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return vals if len(vals) > 1 else vals[0]
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        out = [float(x) for x in t.tolist()]
        return out if len(out) > 1 else out[0]

    def all_true(self, flag):
        if self.world == 1:
            return bool(flag)
        t = self.torch.tensor([1.0 if flag else 0.0], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)


class GraphSteps:
    """`fn(set_index, stream_ptr)` captured as CUDA graphs over a rotation of R buffer sets; run(n) launches EXACTLY n
    steps as graph replays only.  Up to CHUNK steps are ONE graph (a single launch: the driver's --steps 20 is one
    graph); beyond that, n // CHUNK replays of a CHUNK-step graph plus one graph of the remainder."""
    CHUNK = 256

    def __init__(self, b, fn, R, begin=None, end=None):
        self.b, self.fn, self.R, self.graphs = b, fn, R, {}
        self.begin, self.end = begin, end                    # optional hooks around the captured steps (fork / join)
        torch = b.torch
        with torch.cuda.stream(b.stream):
            fn(0, b.st)                                      # module load / first-launch outside capture
            b.stream.synchronize(); b.side.synchronize()

    def _graph(self, n):
        if n not in self.graphs:
            torch, b = self.b.torch, self.b
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(b.stream):
                with torch.cuda.graph(g, stream=b.stream):
                    if self.begin:
                        self.begin()
                    for r in range(n):
                        self.fn(r % self.R, b.st)
                    if self.end:
                        self.end()
            self.graphs[n] = g
        return self.graphs[n]

    def prepare(self, n):
        if n >= self.CHUNK:
            self._graph(self.CHUNK)
        if n % self.CHUNK:
            self._graph(n % self.CHUNK)

    def run(self, n):
        self.prepare(n)
        with self.b.torch.cuda.stream(self.b.stream):
            for _ in range(n // self.CHUNK):
                self.graphs[self.CHUNK].replay()
            if n % self.CHUNK:
                self.graphs[n % self.CHUNK].replay()

    def timed_ms(self, n):
        """Total device time of exactly n steps: barrier + synchronize on both sides, CUDA events on the launching
        stream, max over ranks."""
        torch, b = self.b.torch, self.b
        self.prepare(n)
        lead = self._graph(min(self.R, 8))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b.barrier(); torch.cuda.synchronize()
        with torch.cuda.stream(b.stream):
            # untimed lead-in (a few more warm-up steps, enqueued BEFORE the start event): the device is busy while the
            # host submits the timed graph(s), so the event window holds exactly the n timed steps and no submission gap
            lead.replay()
            e0.record(b.stream)
        self.run(n)
        with torch.cuda.stream(b.stream):
            e1.record(b.stream)
        torch.cuda.synchronize(); b.barrier()
        return b.max_over_ranks(e0.elapsed_time(e1))


def section_headline(b, target_proj, moving, phi, poses32):
    """cfg2 step (device-resident) + per-kernel launch durations."""
    torch, lib, native, vp, dev = b.torch, b.lib, b.native, b.vp, b.dev
    from liftreg_b200 import ops
    R = ROTATION
    sets = [dict(proj=torch.from_numpy(target_proj).to(dev), moving=torch.from_numpy(moving).to(dev),
                 phi=torch.from_numpy(phi).to(dev), lifted=torch.empty((1, P) + VOL, device=dev),
                 warped=torch.empty((1, 1) + VOL, device=dev), gphi=torch.empty((1, 3) + VOL, device=dev))
            for _ in range(R)]
    gout = torch.randn((1, 1) + VOL, device=dev)
    pp = ops._fp(poses32)

    def k_backproject(r, st):
        s = sets[r]
        native.check(lib.lr_backproject_forward(vp(s["proj"]), pp, 1, P, DET[0], DET[1], VOL[0], VOL[1], VOL[2],
                                                vp(s["lifted"]), P * NV, NV, st), "lr_backproject_forward")

    def k_warp(r, st):
        s = sets[r]
        native.check(lib.lr_warp_forward(vp(s["moving"]), vp(s["phi"]), 1, 1, VOL[0], VOL[1], VOL[2], 0, 0, 1, 0,
                                         vp(s["warped"]), st), "lr_warp_forward")

    def k_warp_bwd(r, st):      # d/dphi (what a training step needs: model :69 with the moving image as data)
        s = sets[r]
        native.check(lib.lr_warp_backward(vp(gout), vp(s["moving"]), vp(s["phi"]), 1, 1, VOL[0], VOL[1], VOL[2], 0, 0, 1, 0,
                                          None, vp(s["gphi"]), st), "lr_warp_backward")

    def step_forked(r, st):
        """One step as two parallel graph branches: the backprojection and the warp of a step are independent (in the
        model they are separated by the encoder), so the step forks onto a side stream and joins before the next."""
        fork = torch.cuda.Event(); fork.record(b.stream); b.side.wait_event(fork)
        k_backproject(r, ctypes.c_void_p(b.side.cuda_stream))
        k_warp(r, st)
        join = torch.cuda.Event(); join.record(b.side); b.stream.wait_event(join)

    # The backprojection and the warp of a step are independent (in the model the encoder sits between them), and so are
    # consecutive steps (rotating buffer sets).  Default: the K backprojections run back to back on one stream and the K
    # warps on another, forked at the start of the graph and joined at its end, so one kernel's ramp-up and drain overlap
    # the other kernel's steady state (37.0 us per step; with a fork and a join inside EVERY step, LIFTREG_BENCH_CHAINS=0,
    # 39.4 us; more than one chain per kernel changes nothing).
    n_chains = int(os.environ.get("LIFTREG_BENCH_CHAINS", "1"))
    extra = [torch.cuda.Stream(device=dev) for _ in range(max(0, 2 * n_chains - 2))]
    bp_streams = ([b.side] + extra[:n_chains - 1]) if n_chains else []
    warp_streams = ([b.stream] + extra[n_chains - 1:]) if n_chains else []

    def chains_begin():
        fork = torch.cuda.Event(); fork.record(b.stream)
        for s_ in bp_streams + warp_streams[1:]:
            s_.wait_event(fork)

    def step_chained(r, st):
        k_backproject(r, ctypes.c_void_p(bp_streams[r % n_chains].cuda_stream))
        k_warp(r, ctypes.c_void_p(warp_streams[r % n_chains].cuda_stream))

    def chains_end():
        for s_ in bp_streams + warp_streams[1:]:
            join = torch.cuda.Event(); join.record(s_); b.stream.wait_event(join)

    native.launch_count_reset()
    if n_chains:
        g_step = GraphSteps(b, step_chained, R, begin=chains_begin, end=chains_end)
    else:
        g_step = GraphSteps(b, step_forked, R)
    launches_per_step = native.launch_count()                     # the constructor ran exactly one step eagerly
    g_bp, g_warp, g_wb = GraphSteps(b, k_backproject, R), GraphSteps(b, k_warp, R), GraphSteps(b, k_warp_bwd, R)

    steps, warm = b.args.steps, max(b.args.warmup, 3)
    g_step.prepare(steps); g_step.prepare(warm)
    g_step.run(warm)                                              # W untimed warm-up steps
    torch.cuda.synchronize()
    total_ms = g_step.timed_ms(steps)                             # EXACTLY K steps, graph replays only
    n_k = max(steps, 4000)
    us = {}
    for name, g in (("backproject_forward_kernel", g_bp), ("warp_forward_kernel", g_warp), ("warp_backward_phi_kernel", g_wb)):
        g.run(64)
        us[name] = 1e3 * g.timed_ms(n_k) / n_k
    n_long = int(min(200000, max(n_k, 1.0e6 / max(us["backproject_forward_kernel"] + us["warp_forward_kernel"], 1.0))))
    sustained_ms = g_step.timed_ms(n_long) / n_long               # ~1 s of back-to-back steps for the clock record
    return dict(total_ms=total_ms, us=us, sustained_ms=sustained_ms, launches_per_step=launches_per_step, sets=sets)


def event_time_ms(b, fn, n, warm=2):
    """fn() n times on the current stream between CUDA events (eager launches); max over ranks."""
    torch = b.torch
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize(); b.barrier()
    return b.max_over_ranks(e0.elapsed_time(e1)) / n


def section_cfg4(b):
    """BASELINE configs[3]: DRR sweep, 512^3 CT, 64 views over 60 deg, 512^2 detector, view-sharded over the ranks with
    the NCCL all-gather of the detector images inside the timed window (strong scaling: the total work is fixed)."""
    torch, dev = b.torch, b.dev
    from liftreg_b200 import ops, sharding, synthetic
    n, Pn, det = 512, 64, (512, 512)
    z = np.arange(n, dtype=np.float32)
    vol = (0.1 + 0.05 * np.sin(z / 13.0)[:, None, None] * np.cos(z / 17.0)[None, :, None]
           + 0.04 * np.sin(z / 11.0)[None, None, :]).astype(np.float32)
    tv = torch.from_numpy(vol[None]).to(dev)                       # replicated volume (537 MB)
    poses = synthetic.wrapper_poses(60.0, Pn, n)
    sp = (1.0, 1.0, 1.0)
    slot = (Pn + b.world - 1) // b.world
    buf = torch.zeros((b.world, slot) + det, device=dev)           # gather buffer, reused by every sweep
    reps = 3 if b.world == 1 else 10
    ms_nccl = event_time_ms(b, lambda: sharding.drr_project_sharded(tv, poses, det, sp, buf=buf), reps, warm=1)
    ms_nogather = event_time_ms(b, lambda: sharding.drr_project_sharded(tv, poses, det, sp, buf=buf, gather=False), reps, warm=1)
    full = sharding.drr_project_sharded(tv, poses, det, sp, buf=buf)
    # exchange form (a): the DRR kernel stores into every rank's gather buffer over NVLink (CUDA IPC peer memory); only a
    # 4-byte all-reduce remains as barrier.  Falls back to the NCCL all-gather above if the peer mapping cannot be set up.
    pg, peer_err, ms_peer, ok_peer = None, None, None, True
    if b.world > 1:
        try:
            pg = sharding.PeerGather(Pn, det[0], det[1], dev)
        except Exception as e:                                   # noqa: BLE001  (reported in the JSON line)
            peer_err = "%s: %s" % (type(e).__name__, e)
        if not b.all_true(pg is not None):
            if pg is not None:
                pg.close()
            pg = None
    if pg is not None:
        ms_peer = event_time_ms(b, lambda: sharding.drr_project_sharded(tv, poses, det, sp, peers=pg), reps, warm=2)
        ok_peer = bool(torch.equal(sharding.drr_project_sharded(tv, poses, det, sp, peers=pg), full))
    ms = ms_peer if ms_peer is not None else ms_nccl
    # parity in the run: every rank recomputes a strided subset of the views unsharded and compares bit for bit; rank 0
    # compares ALL views (and so also times the 1-GPU sweep)
    sub = list(range((b.rank * 5) % Pn, Pn, max(1, b.world) + 3))[:4]        # views computed by OTHER ranks too
    ok = all(torch.equal(full[0, v], ops.drr_project(tv, poses[v:v + 1], det, sp)[0, 0]) for v in sub)
    ms_one = None
    if b.rank == 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ref = ops.drr_project(tv, poses, det, sp)
        torch.cuda.synchronize()
        e0.record(); ref = ops.drr_project(tv, poses, det, sp); e1.record()
        torch.cuda.synchronize()
        ms_one = e0.elapsed_time(e1)
        ok = ok and bool(torch.equal(full, ref))
        del ref
    ok = b.all_true(ok and ok_peer)
    if pg is not None:
        pg.close()
    nominal = Pn * det[0] * det[1] * n
    gather_bytes = 4 * Pn * det[0] * det[1]
    out = {"workload": "cfg4: DRR sweep 512^3, 64 views / 60 deg, 512^2 detector; views dealt round-robin to %d rank(s), volume "
                       "replicated, exchange of the images (see `exchange`) inside the timed window" % b.world,
           "scaling": "strong", "ms_per_sweep": ms, "ms_per_sweep_without_all_gather": ms_nogather,
           "exchange": ("p2p stores from the DRR kernel into every rank's buffer (CUDA IPC over NVLink) + 4-byte all-reduce"
                        if ms_peer is not None else "NCCL all-gather + de-interleave" + (" (peer setup failed: %s)" % peer_err if peer_err else "")),
           "ms_per_sweep_p2p_stores": ms_peer, "ms_per_sweep_nccl_all_gather": ms_nccl,
           "all_gather_ms": max(0.0, ms_nccl - ms_nogather), "all_gather_bytes": gather_bytes,
           "nominal_ray_samples": nominal, "samples_per_s": nominal / ms * 1e3, "views_per_rank": slot,
           "one_gpu_unsharded_ms_on_rank0": ms_one, "sharded_parity": ok,
           "limiter": "the DRR kernel itself (issue/latency-bound, see drr_forward_cfg1); beyond it the all-gather "
                      "(%.0f MB, latency-bound over NVLink) and the max over ranks of unequal per-view cost" % (gather_bytes / 1e6)}
    del tv, buf, full
    torch.cuda.empty_cache()
    return out


def section_cfg5(b, target_proj, moving, phi, poses32):
    """BASELINE configs[4]: the resampling ops of a training step (backprojection, warp forward, warp d/dphi) on a batch
    of 32 160^3 pairs.  (a) batch-sharded: 32/N items per rank (the outer level);  (b) z-slab-sharded: ONE
    item (B < G) whose output planes are split over the ranks (lr_*_slab entry points; projections and the moving image
    replicated, so no halo and no collective).  CUDA graphs over rotating buffer sets, max over ranks; the z-slab result is
    gathered after the timed region and compared bit for bit with the unsharded kernels."""
    torch, lib, native, vp, dev = b.torch, b.lib, b.native, b.vp, b.dev
    from liftreg_b200 import ops, sharding
    pp = ops._fp(poses32)
    units_item = (P + 2) * NV                 # 4 backprojection voxel-samples + warp voxel + d/dphi voxel, per voxel
    res = {}

    # ---- (a) batch-sharded
    Btot = 32
    lo, hi = sharding.split_range(Btot, b.world, b.rank)
    Bl = hi - lo
    R = 2
    rep = lambda a, nd: torch.from_numpy(a).to(dev).repeat(Bl, *([1] * nd)).contiguous()
    sets = [dict(proj=rep(target_proj, 3), moving=rep(moving, 4), phi=rep(phi, 4),
                 lifted=torch.empty((Bl, P) + VOL, device=dev), warped=torch.empty((Bl, 1) + VOL, device=dev),
                 gphi=torch.empty((Bl, 3) + VOL, device=dev)) for _ in range(R)]
    gout = torch.randn((Bl, 1) + VOL, device=dev)

    def step_batch(r, st):
        s = sets[r]
        native.check(lib.lr_backproject_forward(vp(s["proj"]), pp, Bl, P, DET[0], DET[1], *VOL, vp(s["lifted"]), P * NV, NV, st), "bp")
        native.check(lib.lr_warp_forward(vp(s["moving"]), vp(s["phi"]), Bl, 1, *VOL, 0, 0, 1, 0, vp(s["warped"]), st), "warp")
        native.check(lib.lr_warp_backward(vp(gout), vp(s["moving"]), vp(s["phi"]), Bl, 1, *VOL, 0, 0, 1, 0, None, vp(s["gphi"]), st), "warp_bwd")

    g = GraphSteps(b, step_batch, R)
    n = 8
    g.run(4)
    ms = g.timed_ms(n) / n
    res["batch_sharded"] = {"workload": "cfg5: backprojection + warp + warp d/dphi of 32 items (160^3, 4x256^2), %d item(s) per rank" % Bl,
                            "scaling": "strong", "items": Btot, "items_per_rank": Bl, "ms_per_step": ms,
                            "items_per_s": Btot / ms * 1e3, "samples_per_s": Btot * units_item / ms * 1e3,
                            "l2": "%d rotating buffer sets of %.0f MB per rank" % (R, Bl * 212.0)}
    del sets, gout, g
    torch.cuda.empty_cache()

    # ---- (b) one item, z-slab-sharded (B < G)
    z0, z1 = sharding.split_range(VOL[0], b.world, b.rank)
    nz = z1 - z0
    R = ROTATION
    t_phi = torch.from_numpy(phi).to(dev)
    gout_full = torch.randn((1, 1) + VOL, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    phi_slab = t_phi[:, :, z0:z1].contiguous()
    gout_slab = gout_full[:, :, z0:z1].contiguous()
    sets = [dict(proj=torch.from_numpy(target_proj).to(dev), moving=torch.from_numpy(moving).to(dev), phi=phi_slab.clone(),
                 gout=gout_slab.clone(), lifted=torch.empty((1, P, nz) + VOL[1:], device=dev),
                 warped=torch.empty((1, 1, nz) + VOL[1:], device=dev), gphi=torch.empty((1, 3, nz) + VOL[1:], device=dev))
            for _ in range(R)]
    nvs = nz * VOL[1] * VOL[2]

    def step_slab(r, st):
        s = sets[r]
        if nz == 0:
            return
        native.check(lib.lr_backproject_forward_slab(vp(s["proj"]), pp, 1, P, DET[0], DET[1], *VOL, z0, nz, vp(s["lifted"]), P * nvs, nvs, st), "bp_slab")
        native.check(lib.lr_warp_forward_slab(vp(s["moving"]), vp(s["phi"]), 1, 1, *VOL, z0, nz, 0, 0, 1, 0, vp(s["warped"]), st), "warp_slab")
        native.check(lib.lr_warp_backward_slab(vp(s["gout"]), vp(s["moving"]), vp(s["phi"]), 1, 1, *VOL, z0, nz, 0, 0, 1, 0, None, vp(s["gphi"]), st), "warp_bwd_slab")

    g = GraphSteps(b, step_slab, R)
    n = 200
    g.run(R)
    ms = g.timed_ms(n) / n
    # parity: gather the slabs (outside the timed region) and compare with the unsharded kernels, bit for bit
    s0 = sets[0]
    ok = True
    for key, dim_full in (("lifted", (1, P) + VOL), ("warped", (1, 1) + VOL), ("gphi", (1, 3) + VOL)):
        full = sharding._gather_slabs(s0[key], VOL[0], 2, b.world, None)
        if b.rank == 0:
            ref = torch.empty(dim_full, device=dev)
            if key == "lifted":
                native.check(lib.lr_backproject_forward(vp(s0["proj"]), pp, 1, P, DET[0], DET[1], *VOL, vp(ref), P * NV, NV, b.st), "bp")
            elif key == "warped":
                native.check(lib.lr_warp_forward(vp(s0["moving"]), vp(t_phi), 1, 1, *VOL, 0, 0, 1, 0, vp(ref), b.st), "warp")
            else:
                native.check(lib.lr_warp_backward(vp(gout_full), vp(s0["moving"]), vp(t_phi), 1, 1, *VOL, 0, 0, 1, 0, None, vp(ref), b.st), "warp_bwd")
            b.stream.synchronize(); torch.cuda.synchronize()
            ok = ok and bool(torch.equal(full, ref))
            del ref
        del full
    ok = b.all_true(ok)
    res["zslab_one_item"] = {"workload": "cfg5, B < G: ONE item's backprojection + warp + warp d/dphi, output planes split over %d rank(s) "
                                         "(%d planes on rank 0), inputs replicated, no collective" % (b.world, sharding.split_range(VOL[0], b.world, 0)[1]),
                             "scaling": "strong", "ms_per_step": ms, "items_per_s": 1e3 / ms, "samples_per_s": units_item / ms * 1e3,
                             "sharded_parity": ok, "l2": "%d rotating buffer sets per rank" % R}
    del sets, g
    torch.cuda.empty_cache()
    return res


def section_drr_cfg1(b, mu, poses):
    """DRR forward at BASELINE configs[0]'s geometry (160^3, 4 views / 60 deg, 240^2 and 256^2 detectors) with BOTH byte
    models of SURVEY 8d, plus the kernel's fraction of the two resources that actually bind it (measured here with the
    probe kernels): aggregate L1 gather bandwidth and warp-instruction issue rate."""
    torch, lib, native, vp, dev = b.torch, b.lib, b.native, b.vp, b.dev
    from liftreg_b200 import ops
    R = ROTATION
    mus = [torch.from_numpy(mu)[None].to(dev) for _ in range(R)]
    sp3 = np.array([2.2, 2.2, 2.2], np.float32)
    poses64 = np.ascontiguousarray(poses, np.float64)
    # probes: L1-hit gather bandwidth and issue rate of this GPU
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks, fpb, iters = sm * 8, 4096, 2000
    pbuf = torch.zeros(blocks * fpb, device=dev)
    sink = torch.zeros(4, device=dev)

    def timed(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(b.stream):
            e0.record(b.stream)
            for _ in range(n):
                fn()
            e1.record(b.stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    ms_l1 = timed(lambda: native.check(lib.lr_probe_l1_gather(vp(pbuf), pbuf.numel(), blocks, fpb, iters, vp(sink), b.st), "probe_l1"))
    l1_gbps = blocks * 256 * iters * 32 / ms_l1 * 1e-6
    ms_issue = timed(lambda: native.check(lib.lr_probe_issue(blocks, 4000, vp(sink), b.st), "probe_issue"))
    issue_ginst = blocks * 8 * 4000 * 8 / ms_issue * 1e-6            # warp-instructions (FFMA) per second, in G/s
    inst = _profile_json("inst_counts.json")
    out = {"probes": {"l1_gather_gbps": l1_gbps, "issue_gwarp_inst_per_s": issue_ginst,
                      "how": "lr_probe_l1_gather (coalesced L1-hit LDG.32, 8 in flight per thread, %d blocks x 256) and "
                             "lr_probe_issue (8 independent FFMA chains per thread), CUDA events" % blocks}}
    for det in ((240, 240), (256, 256)):
        outs = [torch.empty((1, P) + det, device=dev) for _ in range(R)]

        def k_drr(r, st, det=det, outs=outs):
            native.check(lib.lr_drr_forward(vp(mus[r]), 1, VOL[0], VOL[1], VOL[2], ops._dp(poses64), 1, P, det[0], det[1],
                                            ops._fp(sp3), 0, ctypes.c_float(0.1), vp(outs[r]), st), "lr_drr_forward")

        g = GraphSteps(b, k_drr, R)
        g.run(4 * R)
        n_d = 40 * R
        us = 1e3 * g.timed_ms(n_d) / n_d if b.world == 1 else None
        if us is None:          # ranks other than 0 do not run this section: time locally without collectives
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(b.stream):
                e0.record(b.stream)
            g.run(n_d)
            with torch.cuda.stream(b.stream):
                e1.record(b.stream)
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / n_d
        nominal = P * det[0] * det[1] * VOL[1]
        comp_bytes = 4 * NV + 4 * P * det[0] * det[1]
        d = {"us": us, "nominal_ray_samples": nominal, "samples_per_s": nominal / us * 1e6,
             "compulsory_bytes": comp_bytes, "gbps_compulsory": comp_bytes / us * 1e-3,
             "gather_model_gbps": 16.0 * nominal / us * 1e-3,
             "frac_of_l1_peak": 16.0 * nominal / us * 1e-3 / l1_gbps}
        key = "drr_forward_kernel_%dx%d" % det
        if key in inst:
            d["warp_instructions"] = inst[key]
            d["frac_of_issue_floor"] = inst[key] / (issue_ginst * 1e3) / us   # (time at the measured issue rate) / time
        out["%dx%d" % det] = d
    del mus, pbuf
    return out


def section_batch8(b, target_proj, moving, phi, poses32):
    """The two streaming kernels at batch 8 (BASELINE configs[2]: the full forward runs them on 8 items per launch)."""
    torch, lib, native, vp, dev = b.torch, b.lib, b.native, b.vp, b.dev
    from liftreg_b200 import ops
    pp = ops._fp(poses32)
    B8 = 8
    b_proj = [torch.from_numpy(target_proj).to(dev).repeat(B8, 1, 1, 1) for _ in range(2)]
    b_mov = [torch.from_numpy(moving).to(dev).repeat(B8, 1, 1, 1, 1) for _ in range(2)]
    b_phi = [torch.from_numpy(phi).to(dev).repeat(B8, 1, 1, 1, 1) for _ in range(2)]
    b_lift = [torch.empty((B8, P) + VOL, device=dev) for _ in range(2)]
    b_warp = [torch.empty((B8, 1) + VOL, device=dev) for _ in range(2)]
    b_gphi = [torch.empty((B8, 3) + VOL, device=dev) for _ in range(2)]
    gout = torch.randn((B8, 1) + VOL, device=dev)

    def k8_bp(i, st):
        native.check(lib.lr_backproject_forward(vp(b_proj[i]), pp, B8, P, DET[0], DET[1], *VOL, vp(b_lift[i]), P * NV, NV, st), "bp")

    def k8_warp(i, st):
        native.check(lib.lr_warp_forward(vp(b_mov[i]), vp(b_phi[i]), B8, 1, *VOL, 0, 0, 1, 0, vp(b_warp[i]), st), "warp")

    def k8_wb(i, st):
        native.check(lib.lr_warp_backward(vp(gout), vp(b_mov[i]), vp(b_phi[i]), B8, 1, *VOL, 0, 0, 1, 0, None, vp(b_gphi[i]), st), "warp_bwd")

    out = {}
    for name, fn, nbytes, nunits in (("backproject_forward_kernel", k8_bp, B8 * (4 * P * NV + 4 * P * DET[0] * DET[1]), B8 * P * NV),
                                     ("warp_forward_kernel", k8_warp, B8 * 20 * NV, B8 * NV),
                                     ("warp_backward_phi_kernel", k8_wb, B8 * 32 * NV, B8 * NV)):
        g = GraphSteps(b, fn, 2)
        g.run(4)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(b.stream):
            e0.record(b.stream)
        g.run(200)
        with torch.cuda.stream(b.stream):
            e1.record(b.stream)
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 200
        out[name] = {"us_per_launch": us, "us_per_item": us / B8, "bytes": nbytes, "gbps": nbytes / us * 1e-3,
                     "units_per_s": nunits / us * 1e6}
    return out


def section_pca(b):
    """PCA-subspace decode (model :102): streams the 2.75 GB basis once."""
    torch, lib, native, vp, dev = b.torch, b.lib, b.native, b.vp, b.dev
    K = 56
    basis = torch.empty((3 * NV, K), device=dev).normal_(0, 1e-3)
    pmean = torch.zeros(3 * NV, device=dev)
    coefs = torch.randn(1, K, device=dev)
    pouts = [torch.empty((1, 3 * NV), device=dev) for _ in range(2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(b.stream):
        for i in range(3):
            native.check(lib.lr_pca_decode(vp(coefs), vp(basis), vp(pmean), 1, K, 3 * NV, 1, *VOL, vp(pouts[i % 2]), b.st), "lr_pca_decode")
        e0.record(b.stream)
        for i in range(20):
            native.check(lib.lr_pca_decode(vp(coefs), vp(basis), vp(pmean), 1, K, 3 * NV, 1, *VOL, vp(pouts[i % 2]), b.st), "lr_pca_decode")
        e1.record(b.stream)
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    pbytes = 4 * 3 * NV * K + 8 * 3 * NV
    # adjoint wrt the coefficients: the training step's second pass over the basis
    gout = torch.randn(1, 3 * NV, device=dev)
    gco = torch.zeros(1, K, device=dev)
    with torch.cuda.stream(b.stream):
        for i in range(3):
            native.check(lib.lr_pca_decode_backward(vp(gout), vp(basis), 1, K, 3 * NV, vp(gco), b.st), "lr_pca_decode_backward")
        e0.record(b.stream)
        for i in range(20):
            native.check(lib.lr_pca_decode_backward(vp(gout), vp(basis), 1, K, 3 * NV, vp(gco), b.st), "lr_pca_decode_backward")
        e1.record(b.stream)
    torch.cuda.synchronize()
    us_b = 1e3 * e0.elapsed_time(e1) / 20
    bbytes = 4 * 3 * NV * K + 4 * 3 * NV
    return {"us": us, "bytes": pbytes, "gbps": pbytes / us * 1e-3, "workload": "B=1, K=56, N=3*160^3 (+mean, +identity)",
            "backward": {"us": us_b, "bytes": bbytes, "gbps": bbytes / us_b * 1e-3}}


def section_e2e(b, target_proj, moving, phi, poses32, check_sets):
    """End to end through the host-buffer C-ABI: pinned host memory, H2D + kernels + D2H every step."""
    torch, lib, native, vp, dev = b.torch, b.lib, b.native, b.vp, b.dev
    from liftreg_b200 import ops
    pp = ops._fp(poses32)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_proj, h_moving, h_phi = pin(target_proj), pin(moving), pin(phi)
    h_lifted = torch.empty((1, P) + VOL).pin_memory()
    h_warped = torch.empty((1, 1) + VOL).pin_memory()
    ws_bp = torch.empty(lib.lr_backproject_forward_host_workspace_bytes(1, P, DET[0], DET[1], *VOL), dtype=torch.uint8, device=dev)
    ws_w = torch.empty(lib.lr_warp_forward_host_workspace_bytes(1, 1, *VOL), dtype=torch.uint8, device=dev)
    st, st2 = b.st, ctypes.c_void_p(b.side.cuda_stream)
    bp_args = (vp(h_proj), pp, 1, P, DET[0], DET[1], VOL[0], VOL[1], VOL[2], vp(h_lifted), vp(ws_bp), ws_bp.numel())
    w_args = (vp(h_moving), vp(h_phi), 1, 1, VOL[0], VOL[1], VOL[2], 0, 0, 1, 0, vp(h_warped), vp(ws_w), ws_w.numel())

    def step_serial():          # the reference's contract: each call blocks until its result is on the host
        native.check(lib.lr_backproject_forward_host(*bp_args, st), "backproject host")
        native.check(lib.lr_warp_forward_host(*w_args, st), "warp host")

    def step_overlapped():      # one host thread, two streams: the transfers share the full-duplex link
        native.check(lib.lr_backproject_forward_host_async(*bp_args, st), "backproject host async")
        native.check(lib.lr_warp_forward_host_async(*w_args, st2), "warp host async")
        native.check(lib.lr_stream_synchronize(st), "sync")
        native.check(lib.lr_stream_synchronize(st2), "sync")

    # Two steps in flight: step k+1 is enqueued (own streams, workspaces and pinned result buffers) before the host waits for
    # step k, so the next step's H2D runs while this step's D2H drains and neither copy engine idles between steps.  Every
    # step still copies all its inputs in and both results out inside the timed region.
    h_lifted2, h_warped2 = torch.empty((1, P) + VOL).pin_memory(), torch.empty((1, 1) + VOL).pin_memory()
    ws_bp2, ws_w2 = torch.empty_like(ws_bp), torch.empty_like(ws_w)
    s3, s4 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    sets2 = [(bp_args, w_args, st, st2),
             ((vp(h_proj), pp, 1, P, DET[0], DET[1], VOL[0], VOL[1], VOL[2], vp(h_lifted2), vp(ws_bp2), ws_bp2.numel()),
              (vp(h_moving), vp(h_phi), 1, 1, VOL[0], VOL[1], VOL[2], 0, 0, 1, 0, vp(h_warped2), vp(ws_w2), ws_w2.numel()),
              ctypes.c_void_p(s3.cuda_stream), ctypes.c_void_p(s4.cuda_stream))]

    def run_pipelined(n):
        for k in range(n):
            a_bp, a_w, sa, sb = sets2[k % 2]
            native.check(lib.lr_backproject_forward_host_async(*a_bp, sa), "backproject host async")
            native.check(lib.lr_warp_forward_host_async(*a_w, sb), "warp host async")
            if k >= 1:                                   # step k-1 is complete: its results are in host memory
                _, _, pa, pb = sets2[(k - 1) % 2]
                native.check(lib.lr_stream_synchronize(pa), "sync")
                native.check(lib.lr_stream_synchronize(pb), "sync")
        _, _, pa, pb = sets2[(n - 1) % 2]
        native.check(lib.lr_stream_synchronize(pa), "sync")
        native.check(lib.lr_stream_synchronize(pb), "sync")

    n_e2e = 50          # fixed (reported as e2e.steps): a two-deep pipeline needs a run long enough to amortise its fill and drain
    res = {}
    for name, fn in (("serial", step_serial), ("overlapped", step_overlapped), ("pipelined", None)):
        run = (lambda n: [fn() for _ in range(n)]) if fn else run_pipelined
        run(3)
        b.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(n_e2e)
        torch.cuda.synchronize()
        res[name] = 1e3 * (time.perf_counter() - t0) / n_e2e
        # the e2e outputs are the same bits as the device-resident ones
        assert torch.equal(h_warped, check_sets[0]["warped"].cpu()) and torch.equal(h_lifted, check_sets[0]["lifted"].cpu())
        if fn is None:
            assert torch.equal(h_warped2, h_warped) and torch.equal(h_lifted2, h_lifted)
        h_warped.zero_(); h_lifted.zero_()
    res["serial"], res["overlapped"], res["pipelined"] = b.max_over_ranks(res["serial"], res["overlapped"], res["pipelined"])
    h2d = 4 * (h_proj.numel() + h_moving.numel() + h_phi.numel())
    d2h = 4 * (h_lifted.numel() + h_warped.numel())
    return res, h2d, d2h, n_e2e


def run_b200(args):
    b = Bench(args)
    torch, dev, rank, world = b.torch, b.dev, b.rank, b.world
    from liftreg_b200 import synthetic
    from liftreg_b200 import sdct_projection_utils as sdct

    hu, moving, phi, poses = make_inputs()
    poses32 = np.ascontiguousarray(poses.astype(np.float32))
    mu = synthetic.hu_to_mu(hu)
    proj = sdct.calculate_projection(mu, poses, DET, [1, 1, 1], (2.2, 2.2, 2.2), dev)     # our DRR makes the inputs
    target_proj = synthetic.normalise_projection(proj)[None]                             # (1,4,256,256)

    units_bp, units_warp = P * NV, NV
    units = units_bp + units_warp
    kbytes = {"backproject_forward_kernel": 4 * units_bp + 4 * P * DET[0] * DET[1],   # 4 B write / voxel-sample + projections once
              "warp_forward_kernel": 20 * units_warp,                                 # 4 out + 12 phi + 4 image per voxel
              "warp_backward_phi_kernel": 32 * units_warp}                            # grad 4 + phi 12 + image 4 read, 12 written
    kunits = {"backproject_forward_kernel": units_bp, "warp_forward_kernel": units_warp, "warp_backward_phi_kernel": units_warp}

    sampler = ClockSampler(b.local)
    sampler.start()
    head = section_headline(b, target_proj, moving, phi, poses32)
    ms_per_step = head["total_ms"] / args.steps
    value = world * units / (ms_per_step * 1e-3)
    clocks = sampler.stop()

    extras = not args.no_extras
    drr_extra = section_drr_cfg1(b, mu, poses) if (rank == 0 and extras) else {}
    batch8 = section_batch8(b, target_proj, moving, phi, poses32) if (rank == 0 and extras) else None
    pca_extra = section_pca(b) if (rank == 0 and extras) else None
    drr_e2e_ms = None
    if rank == 0 and extras:
        # DRR end to end: calculate_projection's numpy-in / numpy-out contract (sdct:59-100) as preprocessingDRR.py calls it
        for _ in range(2):
            sdct.calculate_projection(mu, poses, (240, 240), [1, 1, 1], (2.2, 2.2, 2.2), dev)
        t0 = time.perf_counter()
        for _ in range(10):
            sdct.calculate_projection(mu, poses, (240, 240), [1, 1, 1], (2.2, 2.2, 2.2), dev)
        drr_e2e_ms = 1e2 * (time.perf_counter() - t0)
    torch.cuda.empty_cache()
    b.barrier()

    # ---- north_star's multi-GPU partitioning (all ranks; also run at N = 1 so that the scaling curve has its base point)
    cfg4 = section_cfg4(b) if extras else None
    cfg5 = section_cfg5(b, target_proj, moving, phi, poses32) if extras else None

    e2e, h2d, d2h, n_e2e = section_e2e(b, target_proj, moving, phi, poses32, head["sets"])

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

        def rl2(a, c):      # parity of the benchmarked outputs against the CPU path, in the same run
            a = a.double(); c = c.double()
            return float(((a - c).norm() / c.norm()).item())
        cpu_baseline["parity_rel_l2"] = {"backproject": rl2(head["sets"][0]["lifted"].cpu(), lifted_ref),
                                         "warp": rl2(head["sets"][0]["warped"].cpu(), warped_ref)}

    if rank == 0:
        peak, peak_src = _peak_hbm()
        traffic = _profile_json("traffic.json")
        kern = {}
        for k, us in head["us"].items():
            kern[k] = {"us": us, "bytes": kbytes[k], "gbps": kbytes[k] / us * 1e-3, "units_per_s": kunits[k] / us * 1e6,
                       "frac_of_hbm_peak": kbytes[k] / us * 1e-3 / peak, "traffic": traffic.get(k)}
        dom = max(("backproject_forward_kernel", "warp_forward_kernel"), key=lambda k: kern[k]["us"])   # kernels of the step
        parities = [s.get("sharded_parity") for s in ([cfg4] if cfg4 else []) + ([cfg5["zslab_one_item"]] if cfg5 else [])]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(
                world, l2="rotating %d buffer sets (%.0f MB) > 126 MB L2; no flush kernel" % (ROTATION, ROTATION * 148.5),
                launch="CUDA graphs of exactly `steps` steps (one graph up to %d steps, else that graph replayed + one graph of the "
                       "remainder; nothing launched eagerly); the K backprojections and the K warps are two independent chains of "
                       "graph nodes on two streams (forked at the start of the graph, joined at its end; "
                       "LIFTREG_BENCH_CHAINS=0: fork and join inside every step)" % GraphSteps.CHUNK,
                numerics="%s (lr_set_numerics; indices and weights bit-exact in both modes)" % b.native.get_numerics(),
                host="%d cpus; %s" % (os.cpu_count(), b.numa)),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbps"], "peak": peak, "unit": "GB/s",
                         "frac": kern[dom]["gbps"] / peak, "traffic": traffic.get(dom), "peak_source": peak_src,
                         "traffic_source": traffic.get("_how"),
                         "algorithmic_bytes_per_launch": kern[dom]["bytes"], "us_per_launch": kern[dom]["us"]},
            # the step as it runs (both kernels concurrently as graph branches): algorithmic bytes of the two launches over
            # the step time -- per GPU, so comparable with `peak` at any N
            "step_roofline": (lambda nbytes: {"bound": "hbm", "achieved": nbytes / (ms_per_step * 1e-3) * 1e-9, "peak": peak,
                                              "unit": "GB/s", "frac": nbytes / (ms_per_step * 1e-3) * 1e-9 / peak,
                                              "algorithmic_bytes_per_step": nbytes,
                                              "what": "backprojection + warp launches of one step, overlapped on one GPU"})(
                kern["backproject_forward_kernel"]["bytes"] + kern["warp_forward_kernel"]["bytes"]),
            "kernels": kern,
            "drr_forward_cfg1": {k: (dict(v, frac_of_hbm_peak_compulsory=v["gbps_compulsory"] / peak) if "gbps_compulsory" in v else v)
                                 for k, v in drr_extra.items()},
            "drr_calculate_projection_e2e_ms": drr_e2e_ms,
            "pca_decode": (dict(pca_extra, frac_of_hbm_peak=pca_extra["gbps"] / peak,
                                backward=dict(pca_extra["backward"], frac_of_hbm_peak=pca_extra["backward"]["gbps"] / peak))
                           if pca_extra else None),
            "batch8_cfg3": ({k: dict(v, frac_of_hbm_peak=v["gbps"] / peak) for k, v in batch8.items()} if batch8 else None),
            "cfg4_drr_view_sharded": cfg4,
            "cfg5_training_ops": cfg5,
            "sharded_parity": (all(parities) if parities else None),
            "sustained_ms_per_step": head["sustained_ms"],
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": world * units / (e2e["pipelined"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e["pipelined"], "steps": n_e2e,
                    "api": "lr_backproject_forward_host_async (stream A) + lr_warp_forward_host_async (stream B) per step, one host "
                           "thread, pinned host buffers, two steps in flight (step k+1 is enqueued on its own streams / workspaces "
                           "/ result buffers before lr_stream_synchronize x2 waits for step k); every step copies all inputs in "
                           "and both results out",
                    "aggregate_link_gbps": world * (h2d + d2h) / (e2e["pipelined"] * 1e-3) * 1e-9,
                    "one_step_in_flight": {"value": world * units / (e2e["overlapped"] * 1e-3), "ms_per_step": e2e["overlapped"],
                                           "api": "the same two async calls, synchronised after every step"},
                    "serial": {"value": world * units / (e2e["serial"] * 1e-3), "ms_per_step": e2e["serial"],
                               "api": "lr_backproject_forward_host then lr_warp_forward_host (blocking calls, the reference's "
                                      "contract sdct:70-72,97-99)"}},
            "gpu_launches": head["launches_per_step"] * args.steps,
            "clocks": clocks,
        }
        _emit(line)
    if world > 1:
        b.dist.destroy_process_group()
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
    ap.add_argument("--no-extras", action="store_true", help="headline + e2e only (skip the DRR / batch-8 / PCA / cfg4 / cfg5 sections)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
