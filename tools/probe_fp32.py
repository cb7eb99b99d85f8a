#!/usr/bin/env python
"""Scalar FFMA vs packed fp32x2 FFMA2 instruction throughput on the current GPU (lr_probe_issue / lr_probe_issue_packed).

    python tools/probe_fp32.py

B200: 3.67 scalar FFMA and 1.98 FFMA2 per clock per SM -- a packed instruction holds the FP32 pipe for two cycles."""
import ctypes, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liftreg_b200 import _native
lib = _native.lib()
sink = torch.zeros(4, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
blocks, iters = 148 * 8, 4000
for name, fn in (("scalar FFMA", lib.lr_probe_issue), ("packed FFMA2", lib.lr_probe_issue_packed)):
    for occ in (8, 4):
        b = 148 * occ
        fn(b, 100, ctypes.c_void_p(sink.data_ptr()), st); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(b, iters, ctypes.c_void_p(sink.data_ptr()), st); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        winst = b * 8 * iters * 8          # warps * instructions
        print("%-13s %d blocks/SM: %.3f ms, %.3f T warp-inst/s = %.2f per clock per SM at 1.9 GHz" % (name, occ, ms, winst / ms * 1e-9, winst / (ms * 1e-3) / 148 / 1.9e9))
