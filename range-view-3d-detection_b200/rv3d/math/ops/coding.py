"""Host-side mirror of ``torchbox3d/math/ops/coding.py`` (decode only; the polars
``build_dataframe`` is outside the hot path)."""
from __future__ import annotations

import torch
from torch import Tensor

from ... import _native as N
from ..._pipeline import cart_as, dtype_code
from ..._util import ptr, require_cuda, stream_ptr

__all__ = ["decode_range_view"]


def decode_range_view(regressands: Tensor, cart: Tensor, enable_azimuth_invariant_targets: bool) -> Tensor:
    """Drop-in for math/ops/coding.py:110-144: (B,8,H,W), (B,3,H,W) -> (B,7,H,W) in the regressands'
    dtype; cart is read in its own dtype; fp64 arithmetic inside, one cast at the end."""
    dev = require_cuda(regressands, cart)
    if regressands.dim() != 4 or regressands.shape[1] != 8 or cart.dim() != 4 or cart.shape[1] != 3:
        raise ValueError("decode_range_view expects regressands (B,8,H,W) and cart (B,3,H,W)")
    B, _, H, W = regressands.shape
    reg = regressands.contiguous()
    crt = cart_as(reg.dtype, cart)
    out = torch.empty((B, 7, H, W), dtype=reg.dtype, device=dev)
    N.check(N.lib().rv3d_decode_range_view(ptr(reg), ptr(crt), ptr(out), dtype_code(reg.dtype), dtype_code(crt.dtype), B, H, W,
                                           int(bool(enable_azimuth_invariant_targets)), stream_ptr(dev)),
            "rv3d_decode_range_view")
    return out
