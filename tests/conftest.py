"""pytest configuration: registers the ``gpu`` marker and puts the repo root and the product
package directory (``range-view-3d-detection_b200/``, which holds the ``rv3d`` package) on sys.path."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "range-view-3d-detection_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
