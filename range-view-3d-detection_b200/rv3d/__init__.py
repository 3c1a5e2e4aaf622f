"""rv3d -- B200-native rasterize -> decode -> NMS path behind torchbox3d's operator signatures.

Module layout mirrors ``torchbox3d`` so a config swaps ``torchbox3d.`` for ``rv3d.``:
    rv3d.math.range_view.build_range_view          rv3d.math.numpy.conversions.{cart_to_sph, ...}
    rv3d.math.ops.coding.decode_range_view         rv3d.math.ops.nms.batched_multiclass_nms
    rv3d.math.ops.iou.iou_3d_axis_aligned          rv3d.nn.decoders.range_decoder.RangeDecoder
"""
from . import _native  # noqa: F401

__version__ = "0.1.0"
