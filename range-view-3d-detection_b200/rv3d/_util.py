"""Small torch-side helpers: device pointers, the current stream, scratch allocation.
torch is plumbing here (device memory + streams); all arithmetic is in librv3d.so."""
from __future__ import annotations

import ctypes as C

import torch


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("rv3d operators run on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")
        dev = dev or t.device
        if t.device != dev:
            raise RuntimeError("all tensors must live on the same CUDA device")
    return dev


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def scratch(nbytes: int, device: torch.device) -> torch.Tensor:
    # torch's caching allocator returns >=512 B aligned blocks
    return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=device)
