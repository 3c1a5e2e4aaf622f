"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
include/rv3d.h declares; the Python mirror exposes the reference's operator names; operators refuse
CPU tensors instead of falling back."""
import ctypes
import inspect
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "rv3d.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rv3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from rv3d import _native as N
    assert N.LIB_PATH.exists(), "librv3d.so not built: run python range-view-3d-detection_b200/build.py"
    lib = ctypes.CDLL(str(N.LIB_PATH))
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(names) == set(N._SIGNATURES), set(names) ^ set(N._SIGNATURES)
    L = N.lib()
    assert L.rv3d_version() == 100
    assert L.rv3d_strerror(0) == b"ok" and b"scratch" in L.rv3d_strerror(-3)


def test_argument_validation_without_gpu():
    """Entry points validate arguments before touching the device (no compute without a GPU)."""
    from rv3d import _native as N
    L = N.lib()
    p = N.RasterParams()
    assert L.rv3d_rasterize(p, None, None, None, None, None, None, None, 0, None) == -1
    parts = N.make_partitions([0, 15, 30], [15, 30, float("inf")], [8, 2, 1])
    assert L.rv3d_num_candidates(parts, 64, 1800) == 187200
    assert L.rv3d_num_candidates(parts, 64, 2650) == 275648
    assert L.rv3d_num_candidates(N.make_partitions([], [], []), 64, 1800) == 64 * 1800
    q = N.NmsParams()
    assert L.rv3d_nms(q, None, None, None, None, None, None, None, None, None, None, 0, None) == -1


def test_python_mirror_signatures():
    from rv3d.math.numpy.conversions import build_range_view_coordinates, cart_to_sph, z_buffer
    from rv3d.math.ops.coding import decode_range_view
    from rv3d.math.ops.iou import iou_3d_axis_aligned
    from rv3d.math.ops.nms import (batched_multiclass_nms, hard_multiclass_nms, weighted_multiclass_nms,
                                   weighted_nms)
    from rv3d.math.range_view import build_range_view
    from rv3d.nn.decoders.range_decoder import RangeDecoder, sample_by_range

    def names(f):
        return list(inspect.signature(f).parameters)

    assert names(build_range_view)[:7] == ["sweep", "laser_mapping", "lidar_offset", "timestamp_ns",
                                           "max_timestamp_ns", "num_lasers", "width"]       # range_view.py:14-22
    assert names(z_buffer)[:6] == ["indices", "distances", "features", "height", "width", "min_distance"]
    assert names(build_range_view_coordinates)[:6] == ["cart", "sph", "laser_numbers", "laser_mapping",
                                                       "n_inclination_bins", "n_azimuth_bins"]
    assert names(cart_to_sph)[:1] == ["cart"]
    assert names(decode_range_view) == ["regressands", "cart", "enable_azimuth_invariant_targets"]
    assert names(sample_by_range) == ["scores", "categories", "cuboids", "cart", "lower_bounds", "upper_bounds",
                                      "subsampling_rates"]
    assert names(batched_multiclass_nms) == ["cuboids", "scores", "categories", "num_pre_nms", "num_post_nms",
                                             "iou_threshold", "min_confidence", "nms_mode"]
    for f in (hard_multiclass_nms, weighted_multiclass_nms):
        assert names(f) == ["cuboids_i", "scores_i", "categories_i", "iou_threshold", "num_pre_nms", "num_post_nms"]
    assert names(weighted_nms) == ["boxes", "data2merge", "scores", "nms_threshold", "merge_thresh"]
    assert names(iou_3d_axis_aligned) == ["cuboids_a", "cuboids_b"]
    dec = RangeDecoder(True, True, [0, 15, 30], [15, 30, float("inf")], [8, 2, 1])         # range_decoder.py:20-27
    assert names(dec.decode)[:4] == ["multiscale_outputs", "post_processing_config", "task_config", "use_nms"]


def test_no_cpu_fallback():
    from rv3d.math.ops.coding import decode_range_view
    from rv3d.math.ops.nms import batched_multiclass_nms
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        decode_range_view(torch.zeros(1, 8, 2, 2), torch.zeros(1, 3, 2, 2), True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        batched_multiclass_nms(torch.zeros(1, 4, 7), torch.ones(1, 4), torch.zeros(1, 4), 10, 10, 0.3, 0.1, "HARD")


def test_product_does_not_import_oracle():
    pkg = ROOT / "range-view-3d-detection_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
        assert "liboracle" not in text and "oracle.c" not in text.replace("oracle/csrc/oracle.c", ""), f


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The Python mirror's ctypes structures must have the C compiler's layout of include/rv3d.h (a drift would corrupt
    every call silently): sizes and field offsets are printed by a tiny C program built against the header."""
    import shutil
    import subprocess
    from rv3d import _native as N
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = {"rv3d_raster_params": N.RasterParams, "rv3d_inputs_params": N.InputsParams, "rv3d_partitions": N.Partitions,
               "rv3d_decode_params": N.DecodeParams, "rv3d_nms_params": N.NmsParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rv3d.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('  printf("rv3d_detection_record %zu\\n", sizeof(rv3d_detection_record));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    from rv3d.math.ops.coding import RECORD_DTYPE
    assert int(out["rv3d_detection_record"]) == RECORD_DTYPE.itemsize == 64


def test_host_side_helpers():
    """Pure host logic of the mirror (no device): dtype codes, threshold rounding, cart dtype rule, partitions."""
    from rv3d import _native as N
    from rv3d._pipeline import cart_as, dtype_code, threshold_as
    assert [dtype_code(d) for d in (torch.float32, torch.float16, torch.bfloat16)] == [0, 1, 2]
    with pytest.raises(TypeError):
        dtype_code(torch.float64)
    # torch compares `scores >= 0.1` in the tensor's dtype
    assert threshold_as(torch.float32, 0.1) == float(torch.tensor(0.1))
    assert threshold_as(torch.float16, 0.1) == float(torch.tensor(0.1, dtype=torch.float16)) != 0.1
    c32, c16 = torch.zeros(1, 3, 2, 2), torch.zeros(1, 3, 2, 2, dtype=torch.float16)
    assert cart_as(torch.float16, c32).dtype == torch.float32          # autocast: cart keeps float32 next to half heads
    assert cart_as(torch.float16, c16).dtype == torch.float16
    assert cart_as(torch.float32, c16).dtype == torch.float32          # widened (exact)
    with pytest.raises(TypeError):
        cart_as(torch.float32, c32.double())
    parts = N.make_partitions([0, 15, 30], [15, 30, float("inf")], [8, 2, 1])
    assert parts.n_partitions == 3 and list(parts.rate)[:3] == [8, 2, 1] and parts.upper[2] == float("inf")
    with pytest.raises(ValueError):
        N.make_partitions([0] * 9, [1] * 9, [1] * 9)                  # more than RV3D_MAX_PARTITIONS
