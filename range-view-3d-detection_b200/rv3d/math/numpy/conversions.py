"""Host-side mirror of ``torchbox3d/math/numpy/conversions.py``: numpy in / numpy out like the
reference, computed on the GPU (torch CUDA tensors are accepted too and returned as tensors)."""
from __future__ import annotations

import numpy as np
import torch

from ... import _native as N
from ..._util import ptr, scratch, stream_ptr

__all__ = ["cart_to_sph", "build_range_view_coordinates", "z_buffer"]


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("rv3d operators need a CUDA device (no CPU fallback)")
    return torch.device(device or "cuda")


def _to_dev(x, dtype, dev):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), device=dev).to(dtype).contiguous()


def _back(t: torch.Tensor, like):
    return t if isinstance(like, torch.Tensor) else t.cpu().numpy()


def cart_to_sph(cart, device=None):
    """math/numpy/conversions.py:46-73: (N,3) -> (N,3) [azimuth, inclination, radius] in float64."""
    dev = _dev(device)
    c = _to_dev(cart, torch.float64, dev).reshape(-1, 3)
    out = torch.empty_like(c)
    N.check(N.lib().rv3d_cart_to_sph(ptr(c), ptr(out), c.shape[0], stream_ptr(dev)), "rv3d_cart_to_sph")
    res = _back(out.reshape(tuple(np.shape(cart))), cart)
    return res if isinstance(cart, torch.Tensor) else res.astype(np.asarray(cart).dtype, copy=False)


def build_range_view_coordinates(cart, sph, laser_numbers, laser_mapping, n_inclination_bins: int = 64,
                                 n_azimuth_bins: int = 1800, device=None, col_mode: str = "library"):
    """math/numpy/conversions.py:9-43 -> (N,3) [row, col, radius].  Like the reference this MUTATES
    ``sph[:, 0]`` in place (az' = (az + pi) * W / tau) when ``sph`` is a float64 numpy array / CUDA tensor."""
    dev = _dev(device)
    s = _to_dev(sph, torch.float64, dev)
    n = s.shape[0]
    las = _to_dev(laser_numbers, torch.int64, dev)
    mp = _to_dev(laser_mapping, torch.int64, dev)
    hybrid = torch.empty((n, 3), dtype=torch.float64, device=dev)
    # "converter_uniform": converters/av2/utils.py:138-145 (build_uniform_inclination=True)
    mode = {"library": N.COL_LIBRARY, "converter": N.COL_CONVERTER, "converter_uniform": N.COL_CONVERTER_UNIFORM}[col_mode]
    N.check(N.lib().rv3d_range_view_coordinates(ptr(s), ptr(las), ptr(mp), mp.numel(), n, n_inclination_bins,
                                                n_azimuth_bins, mode, ptr(hybrid), stream_ptr(dev)),
            "rv3d_range_view_coordinates")
    if isinstance(sph, np.ndarray):
        sph[:, 0] = s[:, 0].cpu().numpy()              # the in-place side effect callers rely on (H3)
    elif isinstance(sph, torch.Tensor) and sph.data_ptr() != s.data_ptr():
        sph[:, 0] = s[:, 0].to(sph.dtype)
    res = _back(hybrid, cart)
    return res if isinstance(cart, torch.Tensor) else res.astype(np.asarray(cart).dtype, copy=False)


def z_buffer(indices, distances, features, height: int, width: int, min_distance: float = 1.0, device=None,
             return_winner: bool = False):
    """math/numpy/conversions.py:106-128: nearest-return scatter, (2,N) int indices, (N,) distances
    (float64 or float32 -- the float32 depth buffer / float64 distance quirk is reproduced exactly),
    (C,N) features -> (C,H,W) float32."""
    dev = _dev(device)
    d_is64 = (distances.dtype in (np.float64, torch.float64))
    f_is64 = (features.dtype in (np.float64, torch.float64))
    idx = _to_dev(indices, torch.int64, dev)
    dist = _to_dev(distances, torch.float64 if d_is64 else torch.float32, dev)
    feat = _to_dev(features, torch.float64 if f_is64 else torch.float32, dev)
    C, n = feat.shape
    image = torch.empty((C, height, width), dtype=torch.float32, device=dev)
    winner = torch.empty((height, width), dtype=torch.int32, device=dev) if return_winner else None
    lib = N.lib()
    work = scratch(lib.rv3d_zbuffer_scratch_bytes(height, width), dev)
    rows, cols = idx[0].contiguous(), idx[1].contiguous()
    N.check(lib.rv3d_zbuffer(ptr(rows), ptr(cols), ptr(dist), int(d_is64), ptr(feat), int(f_is64), C, n, height, width,
                             float(min_distance), ptr(image), ptr(winner), ptr(work), work.numel(), stream_ptr(dev)),
            "rv3d_zbuffer")
    img = _back(image, features)
    return (img, _back(winner, features)) if return_winner else img
