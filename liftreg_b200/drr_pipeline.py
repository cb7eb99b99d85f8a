"""DRR generation for a whole preprocessed dataset: the loop of the reference's tools/preprocessingDRR.py:123-154
 as a three-stage pipeline.

The reference does, per case and strictly in sequence: np.load x2 -> flip -> HU->mu on the host -> H2D -> grid build
+ grid_sample + sum -> D2H -> empty_cache -> (same again for the source) -> np.save x2.  Once the projection itself
takes ~0.1 ms the loop is bound by the file system and PCIe, so here

  * a loader thread reads `{id}_target.npy` / `{id}_source.npy`, applies the SAR->SPR flip (:134-135) while copying into
    a pinned staging slot, and hands the slot over;
  * the main thread uploads the pair on a copy stream, converts HU -> attenuation on the device (lr_atten_coef: the same
    fp32 expression as calc_relative_atten_coef, sdct:6-9, bit for bit), projects BOTH volumes in one lr_drr_forward
    launch (batch 2, shared poses) and downloads the two DRR stacks into a pinned output slot;
  * a writer thread waits for the download event and np.save's `{id}_target_proj.npy` / `{id}_source_proj.npy`;
    `poses.npy` is written once at the end (:154).

Slots form a ring of `depth` entries, so disk reads, PCIe copies, the kernel and disk writes of different cases overlap.
File names, array shapes / dtypes and values are those of the reference loop (tests compare against the serial mirror
calls).  The compute stage is the CUDA one; `stage_factory` is the seam through which the CPU tests substitute a stage of
their own (tests/pipeline_host_stage.py) -- the package itself ships no CPU path.
"""
import os
import queue
import threading

import numpy as np

from . import sdct_projection_utils as sdct

SPACING = (2.2, 2.2, 2.2)          # tools/preprocessingDRR.py:139-143


def dataset_poses(shape, scan_range=None, scan_num=None, geo_path=None, spacing=SPACING):
    """Emitter positions in voxels, float64 (P,3): sdct:139-155 (arc) or sdct:162-163 (CSV in mm / spacing)."""
    if geo_path is not None:
        return np.genfromtxt(geo_path, delimiter=',')[1:] / spacing
    if scan_range is None or scan_num is None:
        raise ValueError("either geo_path or (scan_range, scan_num) is required")
    return sdct._wrapper_poses_scale(scan_range, int(scan_num), 3.5) * shape[1]


def _load_case(preprocessed_path, case_id, flip_axis1):
    target = np.load(os.path.join(preprocessed_path, "%s_target.npy" % case_id))
    source = np.load(os.path.join(preprocessed_path, "%s_source.npy" % case_id))
    if target.shape != source.shape or target.ndim != 3:
        raise ValueError("case %s: target %s / source %s must be equal-shaped 3-D volumes" % (case_id, target.shape, source.shape))
    if flip_axis1:                                           # :134-135 "change orientation from SAR to SPR"
        target, source = np.flip(target, axis=1), np.flip(source, axis=1)
    return target, source


def generate_drr_dataset(preprocessed_path, data_ids, drr_folder, scan_range=None, scan_num=None, geo_path=None,
                         receptor_size=None, spacing=SPACING, flip_axis1=True, device="cuda", depth=3, stage_factory=None,
                         shard=None):
    """Writes `{id}_target_proj.npy`, `{id}_source_proj.npy` (float32 (P,rd,rh)) for every id and `poses.npy`
    (float64 (P,3)) into `drr_folder`; returns the poses.  See the module docstring for the pipeline.

    shard=(rank, world): one process per GPU, each takes the cases rank, rank+world, ... (cases are independent: no
    collective); rank 0 writes `poses.npy`."""
    data_ids = [str(d) for d in data_ids]
    os.makedirs(drr_folder, exist_ok=True)
    if not data_ids:
        return None
    rank, world = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
    if not 0 <= rank < world:
        raise ValueError("shard=(rank, world) needs 0 <= rank < world, got %r" % (shard,))
    first_id = data_ids[0]
    data_ids = data_ids[rank::world]
    depth = max(1, int(depth))
    first_t, _ = _load_case(preprocessed_path, first_id, flip_axis1)
    shape = first_t.shape
    poses = dataset_poses(shape, scan_range, scan_num, geo_path, spacing)
    resolution = sdct._default_resolution(shape, receptor_size)
    P, (rd, rh) = poses.shape[0], (int(resolution[0]), int(resolution[1]))

    # stage protocol: input_slot(slot) -> (2,d,w,h) host array; project(slot) -> token; wait_output(slot, token) -> (2,P,rd,rh)
    if stage_factory is not None:
        stage = stage_factory(shape, P, rd, rh, depth, poses, spacing)
    else:
        stage = _CudaStage(shape, P, rd, rh, depth, device, poses, spacing)

    free_in, ready_in, ready_out = queue.Queue(), queue.Queue(), queue.Queue()
    for s in range(depth):
        free_in.put(s)
    errors = []

    def loader():
        try:
            for case_id in data_ids:
                slot = free_in.get()
                if slot is None:
                    return
                target, source = _load_case(preprocessed_path, case_id, flip_axis1)
                if target.shape != shape:
                    raise ValueError("case %s has shape %s, expected %s" % (case_id, target.shape, shape))
                buf = stage.input_slot(slot)
                np.copyto(buf[0], target, casting="same_kind")       # the flip happens inside this copy
                np.copyto(buf[1], source, casting="same_kind")
                ready_in.put((case_id, slot))
        except BaseException as e:                                   # noqa: BLE001 - reported by the main thread
            errors.append(e)
        finally:
            ready_in.put(None)

    def writer():
        try:
            while True:
                item = ready_out.get()
                if item is None:
                    return
                case_id, slot, token = item
                proj = stage.wait_output(slot, token)                # (2,P,rd,rh) host view, valid until the slot is reused
                np.save(os.path.join(drr_folder, "%s_target_proj.npy" % case_id), proj[0])
                np.save(os.path.join(drr_folder, "%s_source_proj.npy" % case_id), proj[1])
                free_in.put(slot)
        except BaseException as e:                                   # noqa: BLE001
            errors.append(e)
            free_in.put(None)

    t_load = threading.Thread(target=loader, name="drr-loader", daemon=True)
    t_save = threading.Thread(target=writer, name="drr-writer", daemon=True)
    t_load.start(); t_save.start()
    try:
        while True:
            item = ready_in.get()
            if item is None or errors:
                break
            case_id, slot = item
            ready_out.put((case_id, slot, stage.project(slot)))
    finally:
        ready_out.put(None)
        t_save.join()
        free_in.put(None)
        t_load.join()
    if errors:
        raise errors[0]
    if rank == 0:
        np.save(os.path.join(drr_folder, "poses.npy"), poses)        # :154
    return poses


class _CudaStage:
    """Pinned input / output rings, device buffers, one copy stream and one compute stream."""

    def __init__(self, shape, P, rd, rh, depth, device, poses, spacing):
        import torch
        from . import ops
        self.torch, self.ops = torch, ops
        self.dev = sdct._cuda_device(device)
        self.poses, self.res, self.spacing = poses, (rd, rh), spacing
        self.pin_in = [torch.empty((2,) + tuple(shape), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.pin_out = [torch.empty((2, P, rd, rh), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.np_in = [t.numpy() for t in self.pin_in]
        self.np_out = [t.numpy() for t in self.pin_out]
        with torch.cuda.device(self.dev):
            self.dev_in = [torch.empty((2,) + tuple(shape), dtype=torch.float32, device=self.dev) for _ in range(2)]
            self.copy_stream, self.compute_stream = torch.cuda.Stream(), torch.cuda.Stream()
            self.in_free = [None, None]          # event: the kernel that read dev_in[k] has finished
        self.n = 0

    def input_slot(self, slot):
        return self.np_in[slot]

    def project(self, slot):
        torch = self.torch
        k = self.n & 1
        self.n += 1
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.copy_stream):
                if self.in_free[k] is not None:
                    self.copy_stream.wait_event(self.in_free[k])
                self.dev_in[k].copy_(self.pin_in[slot], non_blocking=True)
                uploaded = torch.cuda.Event()
                uploaded.record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(uploaded)
                self.ops.atten_coef_(self.dev_in[k])                               # sdct:6-9 on the device
                proj = self.ops.drr_project(self.dev_in[k], self.poses, self.res, self.spacing, self.ops.YNORM_WM1, 0.1)
                self.in_free[k] = torch.cuda.Event()
                self.in_free[k].record(self.compute_stream)
                self.pin_out[slot].copy_(proj, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.compute_stream)
        return done

    def wait_output(self, slot, token):
        token.synchronize()
        return self.np_out[slot]
