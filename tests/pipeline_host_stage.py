"""CPU test double of liftreg_b200.drr_pipeline._CudaStage (test infrastructure: lives under tests/, not in the package).

`host_stage(project_fn)` returns a stage_factory for drr_pipeline.generate_drr_dataset whose compute step is
`project_fn(mu (d,w,h) float32, poses, resolution, spacing) -> (P,rd,rh)` -- the CPU tests pass the C oracle."""
import numpy as np

from liftreg_b200 import sdct_projection_utils as sdct


class HostStage:
    def __init__(self, shape, P, rd, rh, depth, project_fn, poses, spacing):
        self.fn, self.poses, self.res, self.spacing = project_fn, poses, (rd, rh), spacing
        self.np_in = [np.empty((2,) + tuple(shape), np.float32) for _ in range(depth)]
        self.np_out = [np.empty((2, P, rd, rh), np.float32) for _ in range(depth)]

    def input_slot(self, slot):
        return self.np_in[slot]

    def project(self, slot):
        for v in range(2):
            mu = sdct.calc_relative_atten_coef(self.np_in[slot][v])
            self.np_out[slot][v] = self.fn(mu, self.poses, self.res, self.spacing)
        return None

    def wait_output(self, slot, token):
        return self.np_out[slot]


def host_stage(project_fn):
    return lambda shape, P, rd, rh, depth, poses, spacing: HostStage(shape, P, rd, rh, depth, project_fn, poses, spacing)
