"""Host-side mirror of ``torchbox3d/math/linalg/lie/SO3.py::yaw_to_quat``."""
import torch
from torch import Tensor

from .... import _native as N
from ...._util import ptr, require_cuda, stream_ptr


def yaw_to_quat(yaw_rad: Tensor) -> Tensor:
    """math/linalg/lie/SO3.py:122-134: (...,1) yaw -> (...,4) scalar-first quaternion (wxyz)."""
    dev = require_cuda(yaw_rad)
    y = yaw_rad[:, -1].float().contiguous()
    out = torch.empty((y.shape[0], 4), dtype=torch.float32, device=dev)
    N.check(N.lib().rv3d_yaw_to_quat(ptr(y), ptr(out), y.shape[0], stream_ptr(dev)), "rv3d_yaw_to_quat")
    return out.to(yaw_rad.dtype)
