"""`cppcuda_cffi.roialign` with the reference's cffi calling convention (caller allocates and zeroes `output`,
lib/cppcuda_cffi/src/roi_align_forward_cuda.h:1-7) so that the UNMODIFIED reference lib/model/roi_align.py
(which imports this module on torch != 0.4, roi_align.py:20) runs on the sm_100a kernel."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from detectorch_b200 import _lib  # noqa: E402


def roi_align_forward_cuda(input, rois, output, pooled_height, pooled_width, spatial_scale, sampling_ratio):
    if not (input.is_cuda and rois.is_cuda and output.is_cuda):
        raise TypeError("roi_align_forward_cuda: CUDA tensors required")
    input, rois = input.contiguous(), rois.contiguous()
    ok = _lib.lib().dt_roi_align_forward_nchw(input.data_ptr(), rois.data_ptr(), rois.size(0), rois.size(1), input.size(1), input.size(2),
                                              input.size(3), int(pooled_height), int(pooled_width), float(spatial_scale), int(sampling_ratio),
                                              output.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return 1 if ok == 1 else 0


def roi_align_forward_cpu(*args, **kwargs):
    raise RuntimeError("detectorch_b200 has no CPU RoIAlign (the reference CPU loop is the parity oracle, oracle/_ref)")


def roi_align_backward_cuda(*args, **kwargs):
    raise NotImplementedError("RoIAlign backward is outside the inference hot path of this build")
