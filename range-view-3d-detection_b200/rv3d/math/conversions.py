"""Host-side mirror of the one ``torchbox3d/math/conversions.py`` helper on the hot path."""
from torch import Tensor


def BCHW_to_BKC(x: Tensor) -> Tensor:
    """math/conversions.py:174-186: a view, no arithmetic."""
    b, c, _, _ = x.shape
    return x.permute(0, 2, 3, 1).reshape(b, -1, c)
