"""Zero-copy torch views over raw device pointers owned by libo3db200.so
(the counterpart of Open3D's Tensor-from-Blob views, core/Blob.h:44-70)."""
from __future__ import annotations

import numpy as np
import torch

_TYPESTR = {torch.float32: "<f4", torch.int32: "<i4", torch.uint16: "<i2", torch.uint8: "|u1",
            torch.int64: "<i8", torch.float64: "<f8"}


class _Raw:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(int(s) for s in shape),
                                         "typestr": typestr, "version": 3, "strides": None}


def device_view(ptr, shape, dtype) -> torch.Tensor:
    """The returned tensor aliases library memory: valid until the owner grows or is destroyed."""
    if not ptr or int(np.prod(shape)) == 0:
        return torch.empty(tuple(shape), dtype=dtype, device="cuda")
    t = torch.as_tensor(_Raw(ptr, shape, _TYPESTR[dtype]), device="cuda")
    return t.view(torch.uint16) if dtype == torch.uint16 else t
