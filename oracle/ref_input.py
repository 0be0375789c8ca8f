"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the dataset's per-sample tensor work, pinned against
tests/golden/{normalize,targets}.npz which oracle/make_golden_host.py generated from the reference's
own code:
  * normalize_u8     : torchvision ToTensor + Normalize as built at dream/datasets.py:60-75
  * create_belief_map: dream/image_proc.py:866-910, the per-pixel loop as written there
"""
import numpy as np


def normalize_u8(img_hwc_u8, mean, std):
    """uint8 [H,W,3] -> fp32 [3,H,W]: x/255 then (x-mean)/std, each a correctly rounded fp32 operation."""
    x = img_hwc_u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1).astype(np.float32))


def create_belief_map(image_resolution, points, sigma=2):
    """image_proc.py:866-910 -> float64 [n, height, width]."""
    width, height = image_resolution
    out = np.zeros((len(points), height, width))
    w = int(sigma * 2)
    for n, point in enumerate(points):
        u, v = int(point[0]), int(point[1])                      # truncation toward zero (:888-889)
        if u - w >= 0 and u + w + 1 < width and v - w >= 0 and v + w + 1 < height:     # :893-898
            for i in range(u - w, u + w + 1):
                for j in range(v - w, v + w + 1):
                    out[n, j, i] = np.exp(-(((i - u) ** 2 + (j - v) ** 2) / (2 * (sigma ** 2))))
    return out
