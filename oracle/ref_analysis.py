"""ORACLE (test infrastructure only -- never imported by the product path).

Per-sample restatement of the post-loop of dream/analysis.py:214-262 (frame conversion through
dream/image_proc.py:135-260 and the in-frame / detected L2 metric), one frame at a time exactly as
the reference iterates, for checking dream_b200.analysis' batched form.
"""
import numpy as np


def _shrink_and_crop(in_res, ref_res):
    # dream/image_proc.py:91-132
    in_w, in_h = in_res
    ref_w, ref_h = ref_res
    ref_h_from_w = int(float(in_w) / float(ref_w) * ref_h)
    ref_w_from_h = int(float(in_h) / float(ref_h) * ref_w)
    if in_w >= ref_w_from_h:
        cropped = (ref_w_from_h, in_h)
    else:
        cropped = (in_w, ref_h_from_w)
    return cropped, ((in_w - cropped[0]) // 2, (in_h - cropped[1]) // 2)


def sample_loop(kps_netout_batch, gt_raw_batch, net_out_res, net_in_res, raw_res, preproc):
    dets, metrics = [], []
    for b in range(len(kps_netout_batch)):
        netout = np.array(kps_netout_batch[b], dtype=float)
        netin = []
        for kp in netout:                                        # image_proc.py:135-160
            netin.append([kp[0] / net_out_res[0] * net_in_res[0], kp[1] / net_out_res[1] * net_in_res[1]])
        raw = []
        for kp in netin:                                         # image_proc.py:205-260
            if preproc == "none":
                raw.append([kp[0], kp[1]])
            elif preproc in ("resize", "shrink"):
                raw.append([kp[0] / net_in_res[0] * raw_res[0], kp[1] / net_in_res[1] * raw_res[1]])
            else:
                cropped, coords = _shrink_and_crop(raw_res, net_in_res)
                raw.append([kp[0] / net_in_res[0] * cropped[0] + coords[0],
                            kp[1] / net_in_res[1] * cropped[1] + coords[1]])
        raw = np.array(raw)
        gt = np.array(gt_raw_batch[b], dtype=float)
        errs = []
        for d, g in zip(raw, gt):                                # analysis.py:241-257
            if (d[0] < -999.0 and d[1] < -999.0) or g[0] < 0.0 or g[0] > raw_res[0] or g[1] < 0.0 \
                    or g[1] > raw_res[1]:
                continue
            errs.append(np.linalg.norm(d - g))
        metrics.append(np.mean(errs) if errs else 999.999)
        dets.append(raw.tolist())
    return np.array(dets), np.array(metrics)
