"""Image-side helpers of the DREAM hot path.

`peaks_from_belief_maps` (reference: dream/image_proc.py:914-1018) runs on the device through
dreamb200_peaks; the resolution / keypoint-frame helpers the `DreamNetwork` facade needs
(image_proc.py:18-132, :135-260, :290-350) are small host functions with the reference's names,
argument meaning ((width, height) tuples) and assertions.
"""
import ctypes as C
import os

import numpy as np
import torch

from ._lib import check, lib

KNOWN_IMAGE_PREPROC_TYPES = ["none", "resize", "shrink", "shrink-and-crop"]

_SIGMA = 3.0
_TRUNCATE = 4.0


def gaussian_half_kernel(sigma=_SIGMA, truncate=_TRUNCATE):
    """fp64 taps w[0..radius] of scipy.ndimage.gaussian_filter(sigma) (the filter the reference
    calls at image_proc.py:935), computed on the host with numpy exactly like scipy does:
    radius=int(truncate*sigma+0.5); exp(-0.5/sigma^2 * x^2) normalised over the full kernel."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:], dtype=np.float64), radius


class PeakTable:
    """Raw device results of dreamb200_peaks for n maps (all tensors on the maps' device)."""

    def __init__(self, n_maps, cap, device):
        self.n_maps, self.cap = n_maps, cap
        self.xy = torch.empty((n_maps, cap, 2), dtype=torch.float64, device=device)
        self.score = torch.empty((n_maps, cap), dtype=torch.float32, device=device)
        self.ij = torch.empty((n_maps, cap, 2), dtype=torch.int32, device=device)
        self.counts = torch.empty((n_maps,), dtype=torch.int32, device=device)
        self.summary = torch.empty((n_maps, 4), dtype=torch.float64, device=device)


def find_peaks_device(maps, offset_due_to_upsampling, cap=64):
    """maps: CUDA fp32 tensor [..., h, w] (any leading dims) -> PeakTable (device, no sync)."""
    assert maps.is_cuda and maps.dtype == torch.float32, "belief maps must be a CUDA fp32 tensor"
    maps = maps.contiguous()
    h, w = int(maps.shape[-2]), int(maps.shape[-1])
    n_maps = maps.numel() // (h * w)
    wts, radius = gaussian_half_kernel()
    table = PeakTable(n_maps, cap, maps.device)
    need, mode = C.c_longlong(0), C.c_int(0)
    check(lib().dreamb200_peaks_plan(n_maps, h, w, radius, cap, C.byref(need), C.byref(mode)), "dreamb200_peaks_plan")
    scratch = torch.empty((need.value,), dtype=torch.float32, device=maps.device) if need.value else None
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib().dreamb200_peaks(C.c_void_p(maps.data_ptr()), n_maps, h, w,
                                wts.ctypes.data_as(C.c_void_p), radius, float(offset_due_to_upsampling),
                                C.c_void_p(scratch.data_ptr() if scratch is not None else 0), cap, C.c_void_p(table.xy.data_ptr()),
                                C.c_void_p(table.score.data_ptr()), C.c_void_p(table.ij.data_ptr()),
                                C.c_void_p(table.counts.data_ptr()), C.c_void_p(table.summary.data_ptr()),
                                stream), "dreamb200_peaks")
    return table


def gaussian_smooth_device(maps, sigma=_SIGMA):
    """scipy.ndimage.gaussian_filter(map, sigma) of every [h,w] map of a CUDA fp32 tensor, bit for bit (the smoothing
    stage of `peaks_from_belief_maps`, image_proc.py:935, on its own)."""
    assert maps.is_cuda and maps.dtype == torch.float32, "belief maps must be a CUDA fp32 tensor"
    maps = maps.contiguous()
    h, w = int(maps.shape[-2]), int(maps.shape[-1])
    n_maps = maps.numel() // (h * w)
    wts, radius = gaussian_half_kernel(sigma)
    need, mode = C.c_longlong(0), C.c_int(0)
    check(lib().dreamb200_peaks_plan(n_maps, h, w, radius, 1, C.byref(need), C.byref(mode)), "dreamb200_peaks_plan")
    # the smoothing-only entry runs the whole-map kernel (mode 0) or the two generic passes (scratch n_maps*h*w)
    scratch = None if (mode.value == 0 and not os.environ.get("DREAMB200_PEAKS_BANDED")) else \
        torch.empty((n_maps * h * w,), dtype=torch.float32, device=maps.device)
    out = torch.empty_like(maps)
    check(lib().dreamb200_gaussian_smooth(C.c_void_p(maps.data_ptr()), n_maps, h, w, wts.ctypes.data_as(C.c_void_p),
                                          radius, C.c_void_p(scratch.data_ptr() if scratch is not None else 0),
                                          C.c_void_p(out.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)),
          "dreamb200_gaussian_smooth")
    return out


def peaks_from_belief_maps(belief_map_tensor, offset_due_to_upsampling):
    """Drop-in for dream/image_proc.py:914: [N,h,w] belief maps -> list[N] of list[(x, y, score, id)]
    (x, y Python floats; score np.float32; id running int), peaks in raster order."""
    assert len(belief_map_tensor.shape) == 3, \
        "Expected belief_map_tensor to have shape [N x height x width], but it is {}.".format(
            belief_map_tensor.shape)
    maps = belief_map_tensor.detach()
    if not maps.is_cuda:
        maps = maps.cuda()
    maps = maps.float()
    cap = 64
    while True:
        table = find_peaks_device(maps, offset_due_to_upsampling, cap=cap)
        counts = table.counts.cpu().numpy()
        if counts.size == 0 or int(counts.max()) <= cap:
            break
        cap = int(counts.max())          # rare: a noisy map with more local maxima than the table holds
    xy = table.xy.cpu().numpy()
    score = table.score.cpu().numpy()
    all_peaks, counter = [], 0
    for j in range(maps.shape[0]):
        lst = []
        for i in range(int(counts[j])):
            lst.append((float(xy[j, i, 0]), float(xy[j, i, 1]), score[j, i], counter))
            counter += 1
        all_peaks.append(lst)
    return all_peaks


def select_keypoints_device(table, next_best_score, sentinel=-999.999):
    """Vectorised decision table of DreamNetwork.inference (dream/network.py:548-577) on the
    kernel's summary: exactly one peak -> take it; several -> the best one iff
    score[0]-score[1] >= next_best_score (fp32 arithmetic like the reference's np.float32 scores);
    otherwise the (-999.999, -999.999) sentinel.  Returns a [n_maps, 2] float64 tensor (device)."""
    cnt = table.counts
    s = table.summary
    gap = s[:, 2].float() - s[:, 3].float()
    # (a Python scalar, not torch.tensor(..., device=cuda): that would be a pageable H2D copy, illegal while a CUDA
    # graph is being captured; the comparison still happens in fp32 against float32(next_best_score))
    take = (cnt == 1) | ((cnt > 1) & (gap >= float(np.float32(next_best_score))))
    # torch.where, not boolean-mask assignment: no data-dependent shape, hence no hidden host sync
    return torch.where(take.unsqueeze(1), s[:, :2], torch.full_like(s[:, :2], sentinel))


def belief_targets_device(keypoints, image_resolution, sigma=2):
    """Device form of `create_belief_map` for a whole batch: keypoints fp32 [..., 2]
    (x, y) in the target frame -- the dataset's `keypoint_projections_output` -- -> CUDA fp32 [..., h, w], equal
    bit for bit to `torch.tensor(create_belief_map(res, kps)).float()` (dream/datasets.py:196-200).  The stamp
    values are computed here in fp64 with numpy exactly like the reference and rounded once to fp32."""
    assert len(image_resolution) == 2, \
        'Expected "image_resolution" to have length 2, but it has length {}.'.format(len(image_resolution))
    width, height = int(image_resolution[0]), int(image_resolution[1])
    kp = torch.as_tensor(keypoints)
    assert kp.shape[-1] == 2, "expected keypoints [..., 2]"
    kp = kp.to(device="cuda", dtype=torch.float32).contiguous()
    n_maps = kp.numel() // 2
    w = int(sigma * 2)
    d2 = np.arange(2 * w * w + 1)
    table = np.ascontiguousarray(np.exp(-(d2 / (2 * (sigma ** 2)))).astype(np.float32))
    out = torch.empty(tuple(kp.shape[:-1]) + (height, width), dtype=torch.float32, device=kp.device)
    check(lib().dreamb200_belief_targets(C.c_void_p(kp.data_ptr()), n_maps, height, width, w,
                                         table.ctypes.data_as(C.c_void_p), C.c_void_p(out.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)),
          "dreamb200_belief_targets")
    return out


def normalize_u8_device(frames_u8, mean, std):
    """uint8 [B,H,W,3] CUDA frames -> fp32 NCHW [B,3,H,W], bit-identical to the dataset's
    ToTensor + Normalize(mean, std) (dream/datasets.py:60-75,177-179)."""
    assert frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 \
        and frames_u8.shape[3] == 3, "expected a CUDA uint8 batch [B,H,W,3]"
    x = frames_u8.contiguous()
    B, H, W = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
    m = np.ascontiguousarray(np.broadcast_to(np.asarray(mean, dtype=np.float32).reshape(-1), (3,)))
    s = np.ascontiguousarray(np.broadcast_to(np.asarray(std, dtype=np.float32).reshape(-1), (3,)))
    y = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
    check(lib().dreamb200_normalize_u8(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), B, H, W,
                                       m.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)),
          "dreamb200_normalize_u8")
    return y


def create_belief_map(image_resolution, pointsBelief, sigma=2):
    """Training targets (dream/image_proc.py:866-910): one (2*2sigma+1)^2 Gaussian stamp per point
    whose window lies fully inside the frame; returns fp64 [n_points, height, width]."""
    assert len(image_resolution) == 2, \
        'Expected "image_resolution" to have length 2, but it has length {}.'.format(len(image_resolution))
    width, height = image_resolution
    out = np.zeros((len(pointsBelief), height, width))
    w = int(sigma * 2)
    for n, point in enumerate(pointsBelief):
        u, v = int(point[0]), int(point[1])
        if u - w >= 0 and u + w + 1 < width and v - w >= 0 and v + w + 1 < height:
            ii = np.arange(u - w, u + w + 1)
            jj = np.arange(v - w, v + w + 1)
            d2 = (ii[None, :] - u) ** 2 + (jj[:, None] - v) ** 2
            out[n, v - w:v + w + 1, u - w:u + w + 1] = np.exp(-(d2 / (2 * (sigma ** 2))))
    return out


# ----------------------------------------------------------------------------------------------
# resolution arithmetic and keypoint frame conversions ((width, height) everywhere)
# ----------------------------------------------------------------------------------------------
def _check_preproc(image_preprocessing):
    assert image_preprocessing in KNOWN_IMAGE_PREPROC_TYPES, \
        'Image preprocessing type "{}" is not recognized.'.format(image_preprocessing)


def shrink_resolution(image_input_resolution, image_ref_resolution):
    factor = float(image_ref_resolution[1]) / float(image_input_resolution[1])
    return (int(image_input_resolution[0] * factor), image_ref_resolution[1])


def shrink_and_crop_resolution(image_input_resolution, image_ref_resolution):
    in_w, in_h = image_input_resolution
    ref_w, ref_h = image_ref_resolution
    ref_h_from_w = int(float(in_w) / float(ref_w) * ref_h)
    ref_w_from_h = int(float(in_h) / float(ref_h) * ref_w)
    if in_w >= ref_w_from_h:
        cropped = (ref_w_from_h, in_h)
    else:
        assert in_h >= ref_h_from_w
        cropped = (in_w, ref_h_from_w)
    coords = ((in_w - cropped[0]) // 2, (in_h - cropped[1]) // 2)
    return cropped, coords


def resolution_after_preprocessing(image_input_resolution, image_ref_resolution, image_preprocessing):
    assert len(image_input_resolution) == 2, \
        'Expected "image_input_resolution" to have length 2, but it has length {}.'.format(
            len(image_input_resolution))
    assert len(image_ref_resolution) == 2, \
        'Expected "image_ref_resolution" to have length 2, but it has length {}.'.format(
            len(image_ref_resolution))
    _check_preproc(image_preprocessing)
    if image_preprocessing == "none":
        return image_input_resolution
    if image_preprocessing == "shrink":
        return shrink_resolution(image_input_resolution, image_ref_resolution)
    return image_ref_resolution          # "resize", "shrink-and-crop"


def preprocess_image(input_image, image_ref_resolution, image_preprocessing):
    from PIL import Image as PILImage
    assert isinstance(input_image, PILImage.Image), \
        'Expected "input_image" to be a PIL Image, but it is "{}".'.format(type(input_image))
    _check_preproc(image_preprocessing)
    if image_preprocessing == "none":
        return input_image
    if image_preprocessing == "resize":
        return input_image.resize(tuple(image_ref_resolution), resample=PILImage.BILINEAR)
    if image_preprocessing == "shrink":
        new_res = shrink_resolution(input_image.size, image_ref_resolution)
        return input_image.resize(new_res, resample=PILImage.BILINEAR)
    cropped, coords = shrink_and_crop_resolution(input_image.size, image_ref_resolution)
    box = (coords[0], coords[1], coords[0] + cropped[0], coords[1] + cropped[1])
    return input_image.crop(box).resize(tuple(image_ref_resolution), resample=PILImage.BILINEAR)


def convert_keypoints_to_netin_from_netout(keypoints_netout, net_output_resolution, net_input_resolution):
    kp = np.asarray(keypoints_netout, dtype=float).reshape(-1, 2)
    return kp / np.asarray(net_output_resolution, dtype=float) * np.asarray(net_input_resolution, dtype=float)


def convert_keypoints_to_netout_from_netin(keypoints_netin, net_input_resolution, net_output_resolution):
    kp = np.asarray(keypoints_netin, dtype=float).reshape(-1, 2)
    return kp / np.asarray(net_input_resolution, dtype=float) * np.asarray(net_output_resolution, dtype=float)


def convert_keypoints_to_raw_from_netin(keypoints_netin, net_input_resolution, image_raw_resolution,
                                        image_preprocessing):
    _check_preproc(image_preprocessing)
    kp = np.asarray(keypoints_netin, dtype=float).reshape(-1, 2)
    if image_preprocessing == "none":
        return kp
    nin = np.asarray(net_input_resolution, dtype=float)
    if image_preprocessing in ("resize", "shrink"):
        return kp / nin * np.asarray(image_raw_resolution, dtype=float)
    cropped, coords = shrink_and_crop_resolution(image_raw_resolution, net_input_resolution)
    return kp / nin * np.asarray(cropped, dtype=float) + np.asarray(coords, dtype=float)


def convert_keypoints_to_netin_from_raw(keypoints_raw, image_raw_resolution, net_input_resolution,
                                        image_preprocessing):
    _check_preproc(image_preprocessing)
    kp = np.asarray(keypoints_raw, dtype=float).reshape(-1, 2)
    if image_preprocessing == "none":
        return kp
    raw = np.asarray(image_raw_resolution, dtype=float)
    if image_preprocessing == "resize":
        return kp / raw * np.asarray(net_input_resolution, dtype=float)
    if image_preprocessing == "shrink":
        return kp / raw * np.asarray(shrink_resolution(image_raw_resolution, net_input_resolution), dtype=float)
    cropped, coords = shrink_and_crop_resolution(image_raw_resolution, net_input_resolution)
    return (kp - np.asarray(coords, dtype=float)) / np.asarray(cropped, dtype=float) \
        * np.asarray(net_input_resolution, dtype=float)
