"""CPU oracle for the rasterize -> decode -> NMS path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
timed CPU baseline.  Nothing under ``range-view-3d-detection_b200/`` imports it; the
product path fails loudly when its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * rasterize / decode / sample_by_range / NMS control flow: PINNED against outputs
    of the reference's own functions imported verbatim from ``/root/reference/src``
    (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
  * rotated IoU (detectron2 / mmcv), weighted-NMS IoU + merge (TorchEx), and
    ``quaternion_from_euler`` (kornia): third-party code that is NOT in the
    reference tree and is not installed here -> restated from the published
    algorithms => **parity unpinned** (float32 bits) for those four; values anchored on
    detectron2's and mmcv's own published unit-test vectors (restated in
    tests/test_oracle_iou.py: they pin mmcv's rotation direction, which differs from
    detectron2's), an independent fp64 polygon clip and OpenCV.
"""
from .build import lib, build  # noqa: F401
from .rv_oracle import *  # noqa: F401,F403
