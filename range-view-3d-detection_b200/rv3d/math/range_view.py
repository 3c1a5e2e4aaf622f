"""Range-image rasterization.  Host-side mirror of ``torchbox3d/math/range_view.py``.

``build_range_view`` keeps the reference signature (math/range_view.py:14-22) and semantics;
``rasterize_sweeps`` is the batched tensor-in / tensor-out entry the hot path uses.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import _native as N
from .._util import ptr, require_cuda, scratch, stream_ptr

__all__ = ["build_range_view", "rasterize_sweeps", "pack_sweeps"]


_RASTER_PARAMS: dict = {}


def raster_params(B: int, nmax: int, height: int, width: int, n_azimuth_bins, num_lasers, mapping_len: int, col_mode: str,
                  lidar_offset: Sequence[float], min_distance: float):
    """-> (rv3d_raster_params, scratch bytes).  The parameter block (and the scratch size it implies) only depends on shapes
    and options: built once per key, so that the launch is not waiting on host work every step."""
    key = (B, nmax, height, width, n_azimuth_bins, num_lasers, mapping_len, col_mode,
           tuple(float(v) for v in lidar_offset), float(min_distance))
    hit = _RASTER_PARAMS.get(key)
    if hit is None:
        p = N.RasterParams()
        p.batch, p.max_points, p.height, p.width = B, nmax, height, width
        p.azimuth_bins = width if n_azimuth_bins is None else n_azimuth_bins
        p.num_lasers = mapping_len if num_lasers is None else num_lasers
        if p.num_lasers > mapping_len:
            raise ValueError("laser_mapping is shorter than num_lasers")
        p.col_mode = {"library": N.COL_LIBRARY, "converter": N.COL_CONVERTER}[col_mode]
        p.lidar_offset[:] = [float(v) for v in lidar_offset]
        p.min_distance = float(min_distance)
        if len(_RASTER_PARAMS) > 64:
            _RASTER_PARAMS.clear()
        hit = _RASTER_PARAMS[key] = (p, N.lib().rv3d_rasterize_scratch_bytes(p))
    return hit


def rasterize_sweeps(points: torch.Tensor, laser: torch.Tensor, n_points: torch.Tensor,
                     laser_mapping: torch.Tensor, lidar_offset: Sequence[float], height: int = 64,
                     width: int = 1800, n_azimuth_bins: Optional[int] = None, num_lasers: Optional[int] = None,
                     col_mode: str = "library", min_distance: float = 1.0, return_winner: bool = False,
                     out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None
                     ) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """Rasterize B raw sweeps in one launch pair.

    points (B,Nmax,4) f32 [x,y,z,intensity] (ego frame), laser (B,Nmax) u8, n_points (B,) i32,
    laser_mapping (num_lasers,) i32 -> image (B,7,H,W) f32 [az, inc, range, x, y, z, intensity]
    (math/range_view.py:33), optionally the winner map (B,H,W) i32 (-1 = empty pixel).
    ``n_azimuth_bins`` defaults to ``width`` (the reference's library wrapper always uses 1800,
    see ``build_range_view``)."""
    dev = require_cuda(points, laser, n_points, laser_mapping)
    if points.dtype != torch.float32 or points.dim() != 3 or points.shape[-1] != 4:
        raise ValueError("points must be (B,Nmax,4) float32")
    if laser.dtype != torch.uint8 or n_points.dtype != torch.int32 or laser_mapping.dtype != torch.int32:
        raise ValueError("laser must be uint8; n_points and laser_mapping int32")
    points, laser = points.contiguous(), laser.contiguous()
    B, nmax, _ = points.shape
    lib = N.lib()
    p, need = raster_params(B, nmax, height, width, n_azimuth_bins, num_lasers, laser_mapping.numel(), col_mode, lidar_offset,
                            min_distance)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = scratch(need, dev)
    image = out if out is not None else torch.empty((B, 7, height, width), dtype=torch.float32, device=dev)
    winner = torch.empty((B, height, width), dtype=torch.int32, device=dev) if return_winner else None
    N.check(lib.rv3d_rasterize(p, ptr(points), ptr(laser), ptr(n_points), ptr(laser_mapping), ptr(image),
                               ptr(winner), ptr(workspace), workspace.numel() * workspace.element_size(),
                               stream_ptr(dev)), "rv3d_rasterize")
    return (image, winner) if return_winner else image


def pack_sweeps(sweeps, device, pin: bool = False):
    """[(xyz (n,3) f32, intensity (n,), laser (n,) u8), ...] -> padded host tensors
    (points (B,Nmax,4) f32, laser (B,Nmax) u8, n_points (B,) i32)."""
    B = len(sweeps)
    nmax = max(len(s[0]) for s in sweeps)
    nmax = (nmax + 63) // 64 * 64
    pts = torch.zeros((B, nmax, 4), dtype=torch.float32, pin_memory=pin)
    las = torch.full((B, nmax), 255, dtype=torch.uint8, pin_memory=pin)
    cnt = torch.zeros((B,), dtype=torch.int32, pin_memory=pin)
    for b, (xyz, inten, laser) in enumerate(sweeps):
        n = len(xyz)
        pts[b, :n, :3] = torch.as_tensor(np.ascontiguousarray(xyz, dtype=np.float32))
        pts[b, :n, 3] = torch.as_tensor(np.ascontiguousarray(inten, dtype=np.float32))
        las[b, :n] = torch.as_tensor(np.ascontiguousarray(laser, dtype=np.uint8))
        cnt[b] = n
    return pts, las, cnt


def _column(sweep, name: str) -> np.ndarray:
    col = sweep[name]
    if hasattr(col, "to_numpy"):
        col = col.to_numpy()
    return np.asarray(col)


def _build_range_view_f64(xyz, intensity, laser, laser_mapping, lidar_offset, num_lasers: int, width: int, device) -> np.ndarray:
    """math/range_view.py:23-44 on float64 columns with rv3d.math.numpy.conversions' device operators."""
    from .numpy.conversions import build_range_view_coordinates, cart_to_sph, z_buffer
    keep = np.asarray(laser) < num_lasers                                              # :23-26
    xyz, intensity, laser = xyz[keep], np.asarray(intensity)[keep], np.asarray(laser)[keep].astype(np.int64)
    cart = xyz - np.asarray(lidar_offset, dtype=np.float64)                            # :29
    sph = cart_to_sph(cart, device)                                                    # :30
    features = np.concatenate([sph, xyz, intensity.reshape(-1, 1).astype(np.float64)], axis=1).transpose(1, 0)   # :33 (before the rescale, H3)
    hybrid = build_range_view_coordinates(cart, sph, laser, np.asarray(laser_mapping), n_inclination_bins=num_lasers,
                                          device=device)                               # :34-40 (1800 azimuth bins, the reference's default)
    indices = np.ascontiguousarray(hybrid[:, :2].transpose(1, 0).astype(int))          # :41
    return z_buffer(indices, hybrid[:, 2], np.ascontiguousarray(features), height=num_lasers, width=width, device=device)   # :42-43


def build_range_view(sweep, laser_mapping: np.ndarray, lidar_offset: np.ndarray, timestamp_ns: Optional[int] = None,
                     max_timestamp_ns: Optional[int] = None, num_lasers: int = 64, width: int = 1800,
                     device: Union[str, torch.device] = "cuda") -> np.ndarray:
    """Drop-in for ``torchbox3d.math.range_view.build_range_view`` (math/range_view.py:14-44).

    ``sweep``: a polars DataFrame (as in the reference) or any mapping with columns
    ``x, y, z, intensity, laser_number``.  Returns a (7, num_lasers, width) float32 numpy array.
    Like the reference, the azimuth column is always computed for 1800 bins because the
    reference never forwards ``width`` to build_range_view_coordinates (math/range_view.py:34-40)."""
    del timestamp_ns, max_timestamp_ns  # unused upstream as well (range_view.py:25)
    xyz = np.stack([_column(sweep, "x"), _column(sweep, "y"), _column(sweep, "z")], axis=1)
    if xyz.dtype == np.float64:
        # Float64 coordinate columns (the reference takes whatever dtype the frame holds, range_view.py:27-29): the fused
        # kernel reads float32 points, so these go through the free-standing operators, statement by statement.
        return _build_range_view_f64(xyz, _column(sweep, "intensity"), _column(sweep, "laser_number"), laser_mapping,
                                     lidar_offset, num_lasers, width, device)
    laser = _column(sweep, "laser_number")
    if laser.max(initial=0) > 255:
        raise ValueError("laser_number must fit uint8")
    pts, las, cnt = pack_sweeps([(xyz.astype(np.float32), _column(sweep, "intensity").astype(np.float32),
                                  laser.astype(np.uint8))], device)
    dev = torch.device(device)
    mapping = torch.as_tensor(np.asarray(laser_mapping)[:num_lasers].astype(np.int32), device=dev)
    if mapping.numel() < num_lasers:
        raise ValueError("laser_mapping is shorter than num_lasers")
    image = rasterize_sweeps(pts.to(dev), las.to(dev), cnt.to(dev), mapping, np.asarray(lidar_offset, dtype=np.float64),
                             height=num_lasers, width=width, n_azimuth_bins=1800, num_lasers=num_lasers)
    return image[0].cpu().numpy()
