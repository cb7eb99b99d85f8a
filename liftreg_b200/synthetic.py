"""Seeded synthetic inputs for the resampling path.

Used by bench.py, the tests and oracle/make_golden.py so that the reference, the
oracle and the CUDA path all see the same tensors.  numpy RandomState only
(bit-stable Mersenne Twister streams); no torch RNG.
"""
import numpy as np


def ct_phantom(shape=(160, 160, 160), seed=2021, sigma=1.5, noise_hu=5.0, nodules=20):
    """Synthetic chest CT in HU, float32 (d,w,h): air -1000, ellipsoid body 0, two lungs -800,
    `nodules` spheres +200..+700 HU, Gaussian-smoothed, plus N(0, noise_hu)."""
    from scipy.ndimage import gaussian_filter
    d, w, h = shape
    rs = np.random.RandomState(seed)
    z, y, x = np.meshgrid(np.linspace(-1, 1, d, dtype=np.float32), np.linspace(-1, 1, w, dtype=np.float32),
                          np.linspace(-1, 1, h, dtype=np.float32), indexing="ij")
    hu = np.full(shape, -1000.0, np.float32)
    body = (z / 0.84) ** 2 + (y / 0.70) ** 2 + (x / 0.84) ** 2 <= 1.0
    hu[body] = 0.0
    for cx in (-0.38, 0.38):
        lung = (z / 0.62) ** 2 + (y / 0.45) ** 2 + ((x - cx) / 0.30) ** 2 <= 1.0
        hu[lung] = -800.0
    for _ in range(nodules):
        c = rs.uniform(-0.55, 0.55, 3)
        r = rs.uniform(0.03, 0.08)
        val = rs.uniform(200.0, 700.0)
        hu[(z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2 <= r * r] = val
    hu = gaussian_filter(hu, sigma=sigma, mode="nearest").astype(np.float32)
    hu += rs.standard_normal(shape).astype(np.float32) * np.float32(noise_hu)
    return hu


def hu_to_mu(hu):
    """mu = (max(HU,-1000)+1000)/1000*0.2 in fp32 (reference sdct:6-9)."""
    out = hu.astype(np.float32).copy()
    out[out < -1000] = -1000
    return ((out + np.float32(1000.)) / np.float32(1000.) * np.float32(0.2)).astype(np.float32)


def hu_to_unit(hu):
    """Intensity normalisation of the registration dataset: clip(HU,-1000,0)/1000*2+1 in [-1,1]
    (reference Registration2D3DDataset.py:186-209)."""
    return (np.clip(hu, -1000.0, 0.0) / np.float32(1000.0) * np.float32(2.0) + np.float32(1.0)).astype(np.float32)


def wrapper_poses(scan_range, proj_num, ref_len, emitter_y=3.5):
    """Emitter poses (P,3) float64 in voxel units (reference sdct:139-144,155)."""
    half = scan_range / 2.0
    ps = np.zeros((proj_num, 3), dtype=np.float64)
    ps[:, 1] = emitter_y
    ps[:, 0] = np.tan(np.linspace(-half, half, num=proj_num) / 180.0 * np.pi) * 3.0
    ps[:, 2] = np.linspace(-0.2, 0.2, num=proj_num)
    return ps * ref_len


def smooth_displacement(shape=(160, 160, 160), seed=2021, max_disp=0.05, coarse=10):
    """Smooth random displacement (3,d,w,h) fp32 in normalised [-1,1] units, max |disp| = max_disp:
    N(0,1) on a coarse^3 lattice, separable linear interpolation (align-corners) to full size."""
    rs = np.random.RandomState(seed)
    lat = rs.standard_normal((3, coarse, coarse, coarse))
    out = lat
    for ax, n in enumerate(shape):
        pos = np.linspace(0, coarse - 1, n)
        i0 = np.minimum(np.floor(pos).astype(np.int64), coarse - 2)
        t = pos - i0
        a = np.take(out, i0, axis=ax + 1)
        b = np.take(out, i0 + 1, axis=ax + 1)
        sh = [1, 1, 1, 1]
        sh[ax + 1] = n
        t = t.reshape(sh)
        out = a * (1 - t) + b * t
    out = out / np.abs(out).max() * max_disp
    return out.astype(np.float32)


def identity_map_np(sz):
    """Normalised identity map (3,*sz) fp32 (reference net_utils.py:59-87), numpy-2 semantics:
    float64 index*spacing rounded to fp32, then *2-1 in fp32."""
    idm = np.mgrid[0:sz[0], 0:sz[1], 0:sz[2]].astype(np.float32)
    sp = 1.0 / (np.array(sz) - 1)
    for c in range(3):
        idm[c] *= sp[c]
        idm[c] = idm[c] * 2 - 1
    return idm


def normalise_projection(proj):
    """DRR -> network input range: clip(p,0,6)/6*2-1 (reference Registration2D3DDataset.py:112)."""
    return (np.clip(proj, 0.0, 6.0) / np.float32(6.0) * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
