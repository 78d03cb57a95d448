"""Mirror of the image half of the reference lib/utils/blob.py (prep_im_for_blob :57-87, im_list_to_blob :27-55) on the device:
the uint8 image is uploaded once (a quarter of the fp32 bytes, and before the up-scaling), mean subtraction, OpenCV-exact
bilinear resize, zero padding to the FPN stride and the HWC->CHW transpose happen in one kernel (dt_prep_image)."""
import ctypes

import numpy as np
import torch

from .. import _lib

PIXEL_MEANS = [122.7717, 115.9465, 102.9801]


def im_scale_for(im_shape, target_size=800, max_size=1333):
    """The scale factor of prep_im_for_blob (blob.py:67-76) for one target size."""
    im_size_min, im_size_max = np.min(im_shape[0:2]), np.max(im_shape[0:2])
    im_scale = float(target_size) / float(im_size_min)
    if np.round(im_scale * im_size_max) > max_size:
        im_scale = float(max_size) / float(im_size_max)
    return im_scale


def image_to_blob(im, pixel_means=PIXEL_MEANS, target_size=800, max_size=1333, fpn_on=False, fpn_coarsest_stride=32, device=None):
    """uint8 BGR image [H,W,3] (numpy or CUDA uint8 tensor) -> (blob CUDA fp32 [1,3,Hp,Wp], im_scale): exactly
    torch.FloatTensor(im_list_to_blob(prep_im_for_blob(im, ...)[0], fpn_on)) of the reference, without the host round trip."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    t = im if torch.is_tensor(im) else torch.from_numpy(np.ascontiguousarray(im))
    if t.dtype != torch.uint8 or t.dim() != 3 or t.size(2) != 3:
        raise TypeError("image_to_blob: expected a uint8 [H,W,3] image")
    t = t.to(dev).contiguous()
    h, w = int(t.size(0)), int(t.size(1))
    s = im_scale_for((h, w), target_size, max_size)
    oh, ow = int(np.rint(h * s)), int(np.rint(w * s))            # cvRound, the size cv2.resize(fx, fy) produces
    bh, bw = oh, ow
    if fpn_on:
        stride = float(fpn_coarsest_stride)
        bh, bw = int(np.ceil(oh / stride) * stride), int(np.ceil(ow / stride) * stride)
    blob = torch.empty((1, 3, bh, bw), dtype=torch.float32, device=dev)
    means = (ctypes.c_double * 3)(*[float(m) for m in pixel_means])
    ok = _lib.lib().dt_prep_image(ctypes.c_void_p(t.data_ptr()), h, w, ctypes.cast(means, ctypes.c_void_p), float(s), oh, ow,
                                  ctypes.c_void_p(blob.data_ptr()), bh, bw, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(ok, "dt_prep_image")
    return blob, s


def prep_im_for_blob(im, pixel_means=PIXEL_MEANS, target_sizes=[800], max_size=1333):
    """Same return structure as blob.py:57-87 (list of HWC float32 images, list of scales) but the images are CUDA tensors."""
    ims, scales = [], []
    for ts in target_sizes:
        blob, s = image_to_blob(im, pixel_means, ts, max_size, fpn_on=False)
        ims.append(blob[0].permute(1, 2, 0).contiguous())
        scales.append(s)
    return ims, scales


def im_list_to_blob(ims, fpn_on=False, fpn_coarsest_stride=32):
    """blob.py:27-55 for CUDA HWC tensors (as returned by prep_im_for_blob above): zero-pad to the common / stride-aligned shape,
    NHWC -> NCHW."""
    mh, mw = max(int(i.shape[0]) for i in ims), max(int(i.shape[1]) for i in ims)
    if fpn_on:
        stride = float(fpn_coarsest_stride)
        mh, mw = int(np.ceil(mh / stride) * stride), int(np.ceil(mw / stride) * stride)
    blob = torch.zeros((len(ims), mh, mw, 3), dtype=torch.float32, device=ims[0].device)
    for i, im in enumerate(ims):
        blob[i, :im.shape[0], :im.shape[1], :] = im
    return blob.permute(0, 3, 1, 2).contiguous()
