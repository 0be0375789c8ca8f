"""Batched form of the per-sample post-loop of `analyze_ndds_dataset` (dream/analysis.py:214-277).

The reference converts and scores one frame at a time in Python ("we don't have on-batch keypoint
convertion right now", analysis.py:218-219); at several thousand frames/s per GPU that loop is the
next bottleneck after the device-side peak extraction, so it is restated here on whole [B,K,2]
arrays (host numpy, same float64 arithmetic, same skip rules and the 999.999 "no valid keypoint"
value).  SURVEY.md 8(f) row f1.
"""
import numpy as np

from . import image_proc

NO_METRIC = 999.999


def detected_keypoints_raw(keypoints_netout, net_output_resolution, net_input_resolution,
                           image_raw_resolution, image_preprocessing):
    """[B,K,2] network-output-frame keypoints (the second result of `DreamNetwork.inference`) ->
    [B,K,2] float64 raw-image-frame keypoints (analysis.py:220-233, both conversions at once)."""
    kp = np.asarray(keypoints_netout, dtype=float)
    shape = kp.shape
    netin = image_proc.convert_keypoints_to_netin_from_netout(kp.reshape(-1, 2), net_output_resolution,
                                                              net_input_resolution)
    raw = image_proc.convert_keypoints_to_raw_from_netin(netin, net_input_resolution, image_raw_resolution,
                                                         image_preprocessing)
    return raw.reshape(shape)


def sample_l2_metrics(detected_raw, gt_raw, image_raw_resolution):
    """Per-sample mean L2 error over keypoints that were detected (not the -999.999 sentinel) and whose
    ground truth lies inside the raw frame, bounds inclusive (analysis.py:241-262) -> [B] float64;
    999.999 where no keypoint qualifies."""
    det = np.asarray(detected_raw, dtype=float)
    gt = np.asarray(gt_raw, dtype=float)
    assert det.shape == gt.shape and det.ndim == 3 and det.shape[-1] == 2
    w, h = image_raw_resolution
    skip = ((det[..., 0] < -999.0) & (det[..., 1] < -999.0)) | (gt[..., 0] < 0.0) | (gt[..., 0] > w) \
        | (gt[..., 1] < 0.0) | (gt[..., 1] > h)
    out = np.full(det.shape[0], NO_METRIC)
    for b in np.nonzero((~skip).any(axis=1))[0]:
        # np.mean over the python list of per-keypoint norms, like the reference: same pairwise order
        d = det[b][~skip[b]] - gt[b][~skip[b]]
        out[b] = np.mean([np.linalg.norm(v) for v in d])
    return out


def analyze_batch(keypoints_netout, gt_keypoints_raw, net_output_resolution, net_input_resolution,
                  image_raw_resolution, image_preprocessing):
    """One call per inference batch: returns (detected_raw [B,K,2], metric [B])."""
    det = detected_keypoints_raw(keypoints_netout, net_output_resolution, net_input_resolution,
                                 image_raw_resolution, image_preprocessing)
    return det, sample_l2_metrics(det, gt_keypoints_raw, image_raw_resolution)
