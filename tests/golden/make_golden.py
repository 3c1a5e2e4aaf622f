"""Mint golden vectors from the reference's OWN functions, imported verbatim.

Run in the build container only (needs /root/reference; the GPU box has none):

    python tests/golden/make_golden.py

The reference ships no tests / fixtures, so these files are the
pin for the oracle (tests/test_oracle_golden.py) and for the CUDA path
(tests/test_gpu_*.py).  What is verbatim reference code and what is patched:

  raster_*      torchbox3d.math.numpy.conversions.{cart_to_sph,
                build_range_view_coordinates, z_buffer} verbatim, composed exactly as
                math/range_view.py:29-43 composes them (that wrapper itself needs
                polars, which is not installed).
  raster_conv_* converters/av2/utils.py {cart_to_sph, build_range_view_coordinates,
                z_buffer} verbatim (av2 / polars stubbed; they are not on this path).
  decode_*      torchbox3d.math.ops.coding.decode_range_view and
                torchbox3d.nn.decoders.range_decoder.{sample_by_range, RangeDecoder}
                verbatim (polars / omegaconf / kornia / detectron2 / mmcv /
                weighted_nms_ext stubbed with MagicMock; none of them is executed
                except where stated below).
  nms_*         torchbox3d.math.ops.nms.{batched,hard,weighted}_multiclass_nms
                verbatim Python control flow; the two native calls it makes
                (detectron2 nms_rotated, TorchEx weighted_nms wrapper) are replaced by
                the oracle's C restatement -> these files pin the CONTROL FLOW
                (thresholding, per-class loop, top-k, ordering, dtypes), not the
                third-party IoU arithmetic, which stays "parity unpinned".
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import math
import sys
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

_STUB_TOPLEVEL = {"polars", "omegaconf", "weighted_nms_ext", "detectron2", "kornia", "mmcv", "av2",
                  "pytorch_lightning", "lightning", "hydra", "wandb", "filelock", "cv2", "kornia", "pyarrow", "joblib", "tqdm"}


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Any import below one of the absent third-party packages resolves to a MagicMock module."""

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_TOPLEVEL:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock()
        m.__name__, m.__path__, m.__spec__ = spec.name, [], spec
        return m

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _StubFinder())
sys.path.insert(0, str(REF / "src"))

from torchbox3d.math.numpy import conversions as ref_np  # noqa: E402
from torchbox3d.math.ops import coding as ref_coding  # noqa: E402
from torchbox3d.math.ops import nms as ref_nms  # noqa: E402
from torchbox3d.nn.decoders import range_decoder as ref_dec  # noqa: E402

import oracle  # noqa: E402
from tests import synth  # noqa: E402


def load_converter_utils():
    spec = importlib.util.spec_from_file_location("ref_conv_utils", REF / "converters/av2/utils.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules.setdefault("constants", MagicMock())
    spec.loader.exec_module(mod)
    return mod


def ref_rasterize(xyz, intensity, laser, laser_mapping, lidar_offset, H, W):
    """math/range_view.py:23-43 with the polars plumbing replaced by numpy slicing."""
    keep = laser < H
    xyz, intensity, laser = xyz[keep], intensity[keep], laser[keep]
    xyz_object = xyz
    cart = xyz - lidar_offset
    sph = ref_np.cart_to_sph(cart)
    features = np.concatenate([sph, xyz_object, intensity[:, None]], axis=1).transpose(1, 0)
    hybrid = ref_np.build_range_view_coordinates(cart, sph, laser.astype(np.int64), laser_mapping,
                                                 n_inclination_bins=H, n_azimuth_bins=W)
    indices = np.ascontiguousarray(hybrid[:, :2].transpose(1, 0).astype(int))
    return ref_np.z_buffer(indices, hybrid[:, 2], features, height=H, width=W)


def main():
    # ---------------- rasterize ------------------------------------------------------
    off = np.array([1.356, 0.0, 1.726])
    for tag, (n, H, W, mapping, seed) in {
        "a": (30000, 64, 600, oracle.ROW_MAPPING_64, 0),
        "b": (40000, 32, 1800, np.arange(32), 1),
    }.items():
        xyz, inten, laser = synth.make_points(n, H, seed, offset=off)
        img = ref_rasterize(xyz, inten, laser, mapping, off, H, W)
        np.savez_compressed(OUT / f"raster_{tag}.npz", xyz=xyz, intensity=inten, laser=laser,
                            mapping=mapping, offset=off, H=H, W=W, image=img)
    # adversarial: ties / f32-rounding boundary radii / az = +-pi / sub-1.0 radii
    xyz, inten, laser = synth.make_adversarial_points(seed=7)
    img = ref_rasterize(xyz, inten, laser, np.arange(8), np.zeros(3), 8, 64)
    np.savez_compressed(OUT / "raster_adv.npz", xyz=xyz, intensity=inten, laser=laser, mapping=np.arange(8),
                        offset=np.zeros(3), H=8, W=64, image=img)
    # all-f32 path (f32 offset -> f32 distances)
    xyz, inten, laser = synth.make_points(20000, 16, 3, offset=off)
    off32 = off.astype(np.float32)
    img = ref_rasterize(xyz, inten, laser, np.arange(16), off32, 16, 900)
    np.savez_compressed(OUT / "raster_f32.npz", xyz=xyz, intensity=inten, laser=laser, mapping=np.arange(16),
                        offset=off32, H=16, W=900, image=img)

    # converter flavour of the column formula + z_buffer
    cu = load_converter_utils()
    xyz, inten, laser = synth.make_points(20000, 32, 5, offset=np.zeros(3))
    cart = xyz.astype(np.float64)
    sph = cu.cart_to_sph(cart)
    rng = sph[:, 2].copy()
    hyb = cu.build_range_view_coordinates(cart, sph, laser.astype(int), np.arange(32), 32, 1200)
    feats = np.concatenate([cart, inten[:, None].astype(np.float64), laser[:, None].astype(np.float64),
                            rng[:, None]], axis=-1).T
    img = cu.z_buffer(hyb[:, :2].astype(int).T, hyb[:, -1], feats, height=32, width=1200)
    np.savez_compressed(OUT / "raster_conv.npz", xyz=xyz, intensity=inten, laser=laser, H=32, W=1200, image=img)

    # ---------------- decode ---------------------------------------------------------
    for tag, (B, C, H, W, seed) in {"a": (2, 3, 16, 256, 0), "b": (1, 26, 8, 200, 1)}.items():
        head = synth.make_head_outputs(B, C, H, W, seed, n_objects=6)
        out = {}
        for flag in (True, False):
            out[f"cuboids_{int(flag)}"] = ref_coding.decode_range_view(head["regressands"], head["cart"], flag).numpy()
        scores = head["logits"].sigmoid() * head["mask"]
        scores, cats = scores.max(dim=1, keepdim=True)
        cub = torch.from_numpy(out["cuboids_1"])
        s2, c2, b2 = ref_dec.sample_by_range(scores, cats, cub, head["cart"], (0, 15, 30), (15, 30, math.inf), (8, 2, 1))
        np.savez_compressed(OUT / f"decode_{tag}.npz", **{k: v.numpy() for k, v in head.items()}, **out,
                            scores=scores.numpy(), categories=cats.numpy(), sbr_scores=s2.numpy(),
                            sbr_categories=c2.numpy(), sbr_cuboids=b2.numpy())

    # ---------------- NMS control flow + full RangeDecoder.decode ---------------------
    ref_nms.nms_rotated = lambda boxes, scores, iou_threshold: oracle.nms_rotated(boxes, scores, iou_threshold)
    ref_nms.weighted_nms = oracle.weighted_nms
    ref_dec.yaw_to_quat = oracle.yaw_to_quat
    head = synth.make_head_outputs(2, 3, 16, 256, 11, n_objects=8)
    for mode in ("HARD", "WEIGHTED"):
        dec = ref_dec.RangeDecoder(True, True, [0, 15, 30], [15, 30, math.inf], [8, 2, 1])
        ms = {1: {"cart": head["cart"], "mask": head["mask"],
                  0: {"logits": head["logits"], "regressands": head["regressands"]}}}
        pp = {"num_pre_nms": 50000, "num_post_nms": 1000, "nms_threshold": 0.3, "min_confidence": 0.1,
              "nms_mode": mode}
        params, scores, cats, bidx = dec.decode(ms, pp, {0: ["A", "B", "C"]})
        np.savez_compressed(OUT / f"pipeline_{mode.lower()}.npz", **{k: v.numpy() for k, v in head.items()},
                            params=params.numpy(), scores=scores.numpy(), categories=cats.numpy(),
                            batch_index=bidx.numpy())
    # use_nms=False branch
    dec = ref_dec.RangeDecoder(True, False, [0, 15, 30], [15, 30, math.inf], [8, 2, 1])
    params, scores, cats, bidx = dec.decode(ms, pp, {0: ["A", "B", "C"]}, use_nms=False)
    np.savez_compressed(OUT / "pipeline_nonms.npz", **{k: v.numpy() for k, v in head.items()},
                        params=params.numpy(), scores=scores.numpy(), categories=cats.numpy(),
                        batch_index=bidx.numpy())
    # standalone batched NMS on random clustered boxes, tight pre/post caps
    cub, sc, ca = synth.make_nms_candidates(B=2, K=3000, n_classes=4, n_objects=12, seed=5)
    for mode in ("HARD", "WEIGHTED"):
        o = ref_nms.batched_multiclass_nms(cub, sc, ca, num_pre_nms=500, num_post_nms=20, iou_threshold=0.3,
                                           min_confidence=0.1, nms_mode=mode)
        np.savez_compressed(OUT / f"nms_{mode.lower()}.npz", cuboids=cub.numpy(), scores=sc.numpy(),
                            categories=ca.numpy(), out_cuboids=o[0].numpy(), out_scores=o[1].numpy(),
                            out_categories=o[2].numpy(), out_batch_index=o[3].numpy())
    # ---------------- loader-side subsample_range_view -------------------------
    # the module itself cannot be imported (its Lightning base classes are stubs), so the function's own
    # source is extracted with ast and executed verbatim
    import ast
    import types
    src = (REF / "src/torchbox3d/prototype/loader.py").read_text()
    fn_src = next(ast.get_source_segment(src, n) for n in ast.parse(src).body
                  if isinstance(n, ast.FunctionDef) and n.name == "subsample_range_view")
    ref_loader = types.SimpleNamespace()
    ns = {"torch": torch, "Tensor": torch.Tensor, "Tuple": __import__("typing").Tuple}
    exec(fn_src, ns)
    ref_loader.subsample_range_view = ns["subsample_range_view"]
    xyz, inten, laser = synth.make_points(12000, 16, 21, offset=off)
    img = torch.from_numpy(ref_rasterize(xyz, inten, laser, np.arange(16), off, 16, 300))
    feats, cart, mask = img[[6, 2, 3, 4, 5]].clone(), img[3:6].clone(), img[2:3] > 0
    outs = {}
    for ds, stride, mode in (("av2", 1, "circular"), ("av2", 4, "circular"), ("waymo", 4, "constant")):
        f, m, c = ref_loader.subsample_range_view(feats.clone(), mask.clone(), cart.clone(), ds, stride, mode)
        outs[f"{ds}_{stride}_{mode}_f"], outs[f"{ds}_{stride}_{mode}_m"], outs[f"{ds}_{stride}_{mode}_c"] = f.numpy(), m.numpy(), c.numpy()
    np.savez_compressed(OUT / "subsample.npz", features=feats.numpy(), cart=cart.numpy(), mask=mask.numpy(), **outs)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
