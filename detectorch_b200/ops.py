"""Torch-tensor front-ends of the C-ABI operators.  Torch is only plumbing here (device memory,
current stream); all compute happens in detectorch_b200/csrc.  Every function requires CUDA
tensors -- there is no CPU path."""
import ctypes

import torch

from . import _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise TypeError("detectorch_b200 operators run on CUDA tensors only (no CPU fallback)")


def roi_align_forward_nchw(features, rois, pooled_height, pooled_width, spatial_scale, sampling_ratio, out=None):
    """features [B,C,H,W] fp32 contiguous, rois [R,4|5] -> [R,C,ph,pw].  Mirrors
    roialign.roi_align_forward_cuda of the reference (lib/model/roi_align.py:46-63)."""
    _need_cuda(features, rois)
    if features.dim() != 4 or rois.dim() != 2 or rois.size(1) not in (4, 5):
        raise RuntimeError("roi_align: features must be [B,C,H,W] and rois [R,4|5]")
    if not (features.is_contiguous() and rois.is_contiguous()):
        raise RuntimeError("roi_align: features and rois must be contiguous")   # reference AT_CHECK, roi_align_forward_cuda.cu:189
    if features.dtype != torch.float32 or rois.dtype != torch.float32:
        raise RuntimeError("roi_align: fp32 only")
    R, C = rois.size(0), features.size(1)
    if out is None:
        out = torch.empty((R, C, pooled_height, pooled_width), device=features.device, dtype=torch.float32)
    ok = _lib.lib().dt_roi_align_forward_nchw(_p(features), _p(rois), R, rois.size(1), C, features.size(2), features.size(3),
                                              int(pooled_height), int(pooled_width), float(spatial_scale), int(sampling_ratio),
                                              _p(out), _stream())
    _lib.check(ok, "dt_roi_align_forward_nchw")
    return out


def roi_align_backward_nchw(rois, grad_output, features_size, pooled_height, pooled_width, spatial_scale, sampling_ratio):
    """grad_output [R,C,ph,pw], rois [R,4|5] -> gradient w.r.t. the NCHW features [B,C,H,W] (lib/model/roi_align.py:91-147)."""
    _need_cuda(rois, grad_output)
    B, C, H, W = [int(v) for v in features_size]
    R = rois.size(0)
    if grad_output.dim() != 4 or grad_output.size(0) != R or grad_output.size(1) != C:
        raise RuntimeError("roi_align backward: grad_output must be [R,C,ph,pw]")
    grad_input = torch.zeros((B, C, H, W), device=grad_output.device, dtype=torch.float32)
    ok = _lib.lib().launch_roi_align_backward_cuda(int(grad_output.numel() & 0x7fffffff), _p(grad_output), R, float(spatial_scale), C, H, W,
                                                   int(pooled_height), int(pooled_width), int(sampling_ratio), _p(grad_input), _p(rois),
                                                   rois.size(1), _stream())
    _lib.check(ok, "launch_roi_align_backward_cuda")
    return grad_input


def roi_align_backward_nchw_deterministic(rois, grad_output, features_size, pooled_height, pooled_width, spatial_scale, sampling_ratio, grad_input=None):
    """Same contract as roi_align_backward_nchw, without atomics: every feature-map cell sums its contributions in the order of the
    reference's single-threaded CPU backward (lib/cppcuda/roi_align_backward_cpu.cpp:79-186), so the result is bit-identical to that loop
    and bit-reproducible run to run.  One host read (the contribution count: the sampling grid is adaptive when sampling_ratio == 0)."""
    _need_cuda(rois, grad_output)
    B, C, H, W = [int(v) for v in features_size]
    R = rois.size(0)
    if grad_output.dim() != 4 or grad_output.size(0) != R or grad_output.size(1) != C:
        raise RuntimeError("roi_align backward: grad_output must be [R,C,ph,pw]")
    rois, grad_output = rois.contiguous().float(), grad_output.contiguous().float()
    if grad_input is None:
        grad_input = torch.zeros((B, C, H, W), device=grad_output.device, dtype=torch.float32)
    if R == 0:
        return grad_input
    L = _lib.lib()
    dev = grad_output.device
    with torch.cuda.device(dev):
        scratch = torch.empty((L.dt_roi_align_backward_det_workspace_bytes(R, 0),), dtype=torch.uint8, device=dev)
        total = torch.zeros((1,), dtype=torch.int64, device=dev)
        _lib.check(L.dt_roi_align_backward_plan(_p(rois), R, rois.size(1), float(spatial_scale), int(pooled_height), int(pooled_width),
                                                int(sampling_ratio), _p(scratch), _p(total), _stream()), "dt_roi_align_backward_plan")
        m = int(total.item())
        ws = torch.empty((L.dt_roi_align_backward_det_workspace_bytes(R, m),), dtype=torch.uint8, device=dev)
        _lib.check(L.dt_roi_align_backward_deterministic(_p(grad_output), _p(rois), R, rois.size(1), B, C, H, W, int(pooled_height), int(pooled_width),
                                                         float(spatial_scale), int(sampling_ratio), m, _p(grad_input), _p(ws), _stream()),
                   "dt_roi_align_backward_deterministic")
    return grad_input


def roi_align_forward_nchw_fast(features, rois, pooled_height, pooled_width, spatial_scale, sampling_ratio, out=None):
    """Same contract as roi_align_forward_nchw; sampling_ratio == 2 takes the separable / FMA fast path (fp32 re-association
    differences only), anything else is forwarded to the exact kernel."""
    _need_cuda(features, rois)
    if not (features.is_contiguous() and rois.is_contiguous()):
        raise RuntimeError("roi_align: features and rois must be contiguous")
    B, C, H, W = features.shape
    R = rois.size(0)
    if out is None:
        out = torch.empty((R, C, pooled_height, pooled_width), device=features.device, dtype=torch.float32)
    L = _lib.lib()
    ws = torch.empty((L.dt_roi_align_fast_workspace_bytes(B, C, H, W, R, int(pooled_height), int(pooled_width)),), dtype=torch.uint8, device=features.device)
    ok = L.dt_roi_align_forward_nchw_fast(_p(features), B, _p(rois), R, rois.size(1), C, H, W, int(pooled_height), int(pooled_width),
                                          float(spatial_scale), int(sampling_ratio), _p(out), _p(ws), _stream())
    _lib.check(ok, "dt_roi_align_forward_nchw_fast")
    return out


def roi_align_forward_nhwc(feats, scales, rois, level, pooled_height, pooled_width, sampling_ratio, num_rois=None):
    """feats: list of NHWC maps [B,H,W,C]; rois [R,5]; level int32 [R] or None -> out [R,ph,pw,C]."""
    _need_cuda(rois, *feats)
    n = len(feats)
    ptrs = (ctypes.c_void_p * n)(*[f.data_ptr() for f in feats])
    hs = (ctypes.c_int * n)(*[f.size(1) for f in feats])
    ws = (ctypes.c_int * n)(*[f.size(2) for f in feats])
    sc = (ctypes.c_float * n)(*[float(s) for s in scales])
    C = feats[0].size(3)
    R = rois.size(0)
    out = torch.empty((R, pooled_height, pooled_width, C), device=rois.device, dtype=torch.float32)
    ok = _lib.lib().dt_roi_align_forward_nhwc(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(hs, ctypes.c_void_p),
                                              ctypes.cast(ws, ctypes.c_void_p), ctypes.cast(sc, ctypes.c_void_p), n, _p(rois),
                                              _p(level), _p(num_rois), R, C, pooled_height, pooled_width, sampling_ratio, _p(out),
                                              _stream())
    _lib.check(ok, "dt_roi_align_forward_nhwc")
    return out


def nms(dets, thresh):
    """dets [N,5] fp32 CUDA (x1,y1,x2,y2,score) -> int64 CUDA tensor of ascending kept indices
    (the contract of lib/utils/boxes.py:332-336)."""
    _need_cuda(dets)
    n = dets.size(0)
    if n == 0:
        return torch.zeros((0,), dtype=torch.int64, device=dets.device)
    dets = dets.contiguous().float()
    L = _lib.lib()
    ws = torch.empty((L.dt_nms_workspace_bytes(n),), dtype=torch.uint8, device=dets.device)
    keep = torch.empty((n,), dtype=torch.int64, device=dets.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=dets.device)
    _lib.check(L.dt_nms(_p(dets), n, float(thresh), _p(keep), _p(cnt), _p(ws), _stream()), "dt_nms")
    return keep[:int(cnt.item())]


def segm_rle(masks, classes, boxes, im_h, im_w, thresh=0.5, expanded=False, num_dets=None, runs_cap=4096):
    """Mask paste + COCO RLE on the device (lib/utils/result_utils.py:170-228).
    masks [D,K,M,M] (classes int32 [D] selects the plane) or [D,M,M]; boxes [D,4] fp32 reference boxes, or int32 already
    expanded boxes when expanded=True.  Returns (counts list of uint32 arrays, strings list of bytes)."""
    _need_cuda(masks, boxes, classes)
    D = masks.size(0)
    if D == 0:
        return [], []
    masks = masks.contiguous().float()
    M = masks.size(-1)
    K = masks.size(1) if masks.dim() == 4 else 1
    boxes = boxes.contiguous()
    if expanded and boxes.dtype != torch.int32:
        raise TypeError("segm_rle: expanded boxes must be int32")
    if not expanded:
        boxes = boxes.float()
    if classes is not None:
        classes = classes.contiguous().to(torch.int32)
    L = _lib.lib()
    dev = masks.device
    while True:
        ws = torch.empty((L.dt_segm_workspace_bytes(D, runs_cap),), dtype=torch.uint8, device=dev)
        counts = torch.empty((D, runs_cap), dtype=torch.int32, device=dev)
        ncount = torch.empty((D,), dtype=torch.int32, device=dev)
        strings = torch.empty((L.dt_segm_strings_bytes(D, runs_cap),), dtype=torch.uint8, device=dev)
        offs = torch.empty((D + 1,), dtype=torch.int64, device=dev)
        over = torch.zeros((1,), dtype=torch.int32, device=dev)
        ok = L.dt_segm_rle(_p(masks), _p(classes), K, M, _p(None if expanded else boxes), _p(boxes if expanded else None), _p(num_dets), D,
                           int(im_h), int(im_w), float(thresh), _p(counts), _p(ncount), runs_cap, _p(strings), _p(offs), _p(over), _p(ws),
                           _stream())
        _lib.check(ok, "dt_segm_rle")
        need = int(over.item())
        if need == 0:
            break
        runs_cap = max(2 * runs_cap, need + 1)     # a detection needed more runs than provisioned: grow and redo
    offs_h = offs.cpu().numpy()
    str_h = strings[:int(offs_h[-1])].cpu().numpy().tobytes()
    n_h = ncount.cpu().numpy()
    nmax = int(n_h.max()) if D else 0
    cnt_h = counts[:, :max(nmax, 1)].cpu().numpy().view("uint32")
    return [cnt_h[d, :n_h[d]].copy() for d in range(D)], [str_h[offs_h[d]:offs_h[d + 1]] for d in range(D)]


def segm_paste(masks, classes, boxes, im_h, im_w, thresh=0.5, expanded=False, num_dets=None):
    """Pasted binary masks uint8 [D, im_h, im_w] (the im_mask of result_utils.py:203-217 for every detection)."""
    _need_cuda(masks, boxes, classes)
    D = masks.size(0)
    masks = masks.contiguous().float()
    M = masks.size(-1)
    K = masks.size(1) if masks.dim() == 4 else 1
    boxes = boxes.contiguous() if expanded else boxes.contiguous().float()
    if classes is not None:
        classes = classes.contiguous().to(torch.int32)
    out = torch.empty((D, int(im_h), int(im_w)), dtype=torch.uint8, device=masks.device)
    ok = _lib.lib().dt_segm_paste(_p(masks), _p(classes), K, M, _p(None if expanded else boxes), _p(boxes if expanded else None),
                                  _p(num_dets), D, int(im_h), int(im_w), float(thresh), _p(out), _stream())
    _lib.check(ok, "dt_segm_paste")
    return out


def tf32_residual(w):
    lo = torch.empty_like(w)
    _lib.check(_lib.lib().dt_tf32_residual(_p(w), _p(lo), w.numel(), _stream()), "dt_tf32_residual")
    return lo


def fp16_split(w, multiplier=1.0):
    """w fp32 -> (hi, lo) fp16 tensors of w * multiplier (multiplier must be a power of two)."""
    _need_cuda(w)
    hi = torch.empty(w.shape, dtype=torch.float16, device=w.device)
    lo = torch.empty_like(hi)
    _lib.check(_lib.lib().dt_fp16_split(_p(w.contiguous()), w.numel(), float(multiplier), _p(hi), _p(lo), _stream()), "dt_fp16_split")
    return hi, lo


def weight_multiplier(w):
    """Power of two that brings max|w| into [2^13, 2^14): the fp16 low half then stays a normal number for every weight
    within 2^-9 of the largest one."""
    m = float(w.abs().max())
    if not (m > 0.0):
        return 1.0
    import math
    return 2.0 ** (13 - math.floor(math.log2(m)))


def conv2d_nhwc(x, w_kmajor, scale, shift, kh, kw, pad, stride, w_lo=None, residual=None, up_src=None, relu=False, sigmoid_ch=0,
                passes=3, force_block_n=0, out=None, kind="tf32", range_flag=None):
    """x [N,H,W,Cin] fp32 (channels-last memory), w_kmajor [Cout, kh*kw*Cin] -> y [N,Ho,Wo,Cout].
    kind "tf32": 3xTF32 (w_lo = tf32 residual); kind "f16": 3xFP16 on the kind::f16 pipe (weights split here)."""
    _need_cuda(x, w_kmajor)
    N, H, W, Cin = x.shape
    Cout = w_kmajor.size(0)
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    if out is None:
        out = torch.empty((N, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
    res_mode = 1 if residual is not None else (2 if up_src is not None else 0)
    up_h = up_src.size(1) if up_src is not None else 0
    up_w = up_src.size(2) if up_src is not None else 0
    if kind == "f16":
        mult = weight_multiplier(w_kmajor)
        hi, lo = fp16_split(w_kmajor, mult)
        sc = (scale / mult).contiguous()
        ok = _lib.lib().dt_conv2d_nhwc_f16x3(_p(x), N, H, W, Cin, Cin, _p(hi), _p(lo), Cout, kh, kw, pad, stride, _p(sc), _p(shift), _p(residual),
                                             res_mode, _p(up_src), up_h, up_w, int(relu), int(sigmoid_ch), int(passes), int(force_block_n),
                                             _p(range_flag), _p(out), Cout, _stream())
        _lib.check(ok, "dt_conv2d_nhwc_f16x3")
        return out
    if w_lo is None:
        w_lo = tf32_residual(w_kmajor)
    ok = _lib.lib().dt_conv2d_nhwc(_p(x), N, H, W, Cin, Cin, _p(w_kmajor), _p(w_lo), Cout, kh, kw, pad, stride, _p(scale), _p(shift),
                                   _p(residual), res_mode, _p(up_src), up_h, up_w, int(relu), int(sigmoid_ch), int(passes),
                                   int(force_block_n), _p(out), Cout, _stream())
    _lib.check(ok, "dt_conv2d_nhwc")
    return out
