"""CPU restatement of the keypoint extraction that follows the network (TEST INFRASTRUCTURE).

  peaks_from_belief_maps  -- dream/image_proc.py:914-1018
  gaussian_filter_f32     -- scipy.ndimage.gaussian_filter(sigma=3) as called at image_proc.py:935
                             (scipy is the reference's third-party dependency, unpinned in
                             requirements.txt:13; 1.18.1 in this image).  Restated from scipy's
                             published algorithm: _gaussian_kernel1d (radius = int(4*sigma+0.5) = 12,
                             w = exp(-x^2/(2 sigma^2)) normalised in fp64) + correlate1d (symmetric
                             branch: out = x[0]*w[0]; for j=-r..-1: out += (x[j]+x[-j])*w[j], fp64),
                             mode="reflect" (d c b a | a b c d | d c b a), axis 0 then axis 1, each
                             pass rounded to fp32 on store.  tests/ check it bit-for-bit against scipy.
  select_keypoints        -- the decision table of DreamNetwork.inference, dream/network.py:548-577
  create_belief_map       -- dream/image_proc.py:866-910 (used to synthesise test maps)
"""
import numpy as np

SIGMA = 3
TRUNCATE = 4.0
THRESH = 0.01            # thresh_map_after_gaussian_filter, image_proc.py:925
NEXT_BEST = 0.25         # belief_peak_next_best_score, network.py:191
SENTINEL = -999.999      # network.py:572


def gaussian_weights(sigma=SIGMA, truncate=TRUNCATE):
    """Half kernel w[0..r] (w[d] multiplies x[i-d]+x[i+d]); same arithmetic as scipy's
    _gaussian_kernel1d: exp(-0.5/sigma^2 * x^2) / sum, all fp64."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    phi = phi / phi.sum()
    return phi[radius:].copy(), radius


def _reflect_index(i, n):
    """scipy 'reflect' (half-sample symmetric) index map for any out-of-range i."""
    period = 2 * n
    i = np.mod(i, period)
    return np.where(i >= n, period - 1 - i, i)


def _correlate1d_sym(a32, w, radius, axis):
    a = np.moveaxis(a32.astype(np.float64), axis, -1)
    n = a.shape[-1]
    idx = np.arange(n)
    out = a * w[0]
    for d in range(radius, 0, -1):            # j = -radius .. -1  (farthest pair first)
        lo = a[..., _reflect_index(idx - d, n)]
        hi = a[..., _reflect_index(idx + d, n)]
        out = out + (lo + hi) * w[d]
    return np.moveaxis(out.astype(np.float32), -1, axis)


def gaussian_filter_f32(m):
    w, r = gaussian_weights()
    t = _correlate1d_sym(np.asarray(m, dtype=np.float32), w, r, 0)
    return _correlate1d_sym(t, w, r, 1)


def peak_mask(sm):
    """image_proc.py:936-954: >= the four zero-padded shifted copies and > 0.01 (compared in fp32)."""
    z = np.zeros_like(sm)
    up = z.copy(); up[1:, :] = sm[:-1, :]
    dn = z.copy(); dn[:-1, :] = sm[1:, :]
    lf = z.copy(); lf[:, 1:] = sm[:, :-1]
    rt = z.copy(); rt[:, :-1] = sm[:, 1:]
    return (sm >= up) & (sm >= dn) & (sm >= lf) & (sm >= rt) & (sm > np.float32(THRESH))


def _pairwise_sum25(v):
    """numpy's pairwise summation for a contiguous 25-element fp64 vector (n < 128 block:
    8 running partials over the first 24, tree-combined, then the tail)."""
    r = [v[j] for j in range(8)]
    for i in (8, 16):
        for j in range(8):
            r[j] = r[j] + v[i + j]
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    return res + v[24]


def refine_peak(map_ori, px, py, offset):
    """5x5 weighted centroid on the UNsmoothed map (image_proc.py:961-998).  Returns (x, y)."""
    H, W = map_ori.shape
    wts = np.zeros(25); xv = np.zeros(25); yv = np.zeros(25)
    for i in range(-2, 3):          # row offset
        for j in range(-2, 3):      # col offset
            yy, xx = py + i, px + j
            if yy < 0 or yy >= H or xx < 0 or xx >= W:
                continue
            k = (j + 2) * 5 + (i + 2)      # weights[j+ran, i+ran]
            wts[k] = float(map_ori[yy, xx]); xv[k] = xx; yv[k] = yy
    scl = _pairwise_sum25(wts)
    if scl == 0.0:                  # np.average raises ZeroDivisionError -> integer peak
        return float(px + offset), float(py + offset)
    with np.errstate(all="ignore"):
        x = _pairwise_sum25(xv * wts) / scl + offset
        y = _pairwise_sum25(yv * wts) / scl + offset
    return float(x), float(y)


def peaks_from_belief_maps(maps, offset_due_to_upsampling):
    """maps: [N,H,W] float32 array-like -> list[N] of list[(x, y, score, id)] in raster order."""
    maps = np.asarray(maps, dtype=np.float32)
    assert maps.ndim == 3
    all_peaks, counter = [], 0
    for j in range(maps.shape[0]):
        ori = maps[j]
        sm = gaussian_filter_f32(ori)
        ys, xs = np.nonzero(peak_mask(sm))
        lst = []
        for px, py in zip(xs.tolist(), ys.tolist()):
            x, y = refine_peak(ori, px, py, offset_due_to_upsampling)
            lst.append((x, y, ori[py, px], counter))
            counter += 1
        all_peaks.append(lst)
    return all_peaks


def select_keypoints(peaks, next_best=NEXT_BEST):
    """network.py:548-577: one peak -> take it; several -> best iff score gap >= 0.25; else sentinel."""
    out = []
    for peak in peaks:
        if len(peak) == 1:
            out.append([peak[0][0], peak[0][1]])
        elif len(peak) > 1:
            srt = sorted(peak, key=lambda q: q[2], reverse=True)
            if np.float32(srt[0][2]) - np.float32(srt[1][2]) >= np.float32(next_best):
                out.append([srt[0][0], srt[0][1]])
            else:
                out.append([SENTINEL, SENTINEL])
        else:
            out.append([SENTINEL, SENTINEL])
    return out


def create_belief_map(image_resolution, points, sigma=2):
    """image_proc.py:866-910: one (2w+1)^2 Gaussian stamp per in-frame point, fp64 [n,H,W]."""
    width, height = image_resolution
    out = np.zeros((len(points), height, width))
    w = int(sigma * 2)
    for n, pt in enumerate(points):
        u, v = int(pt[0]), int(pt[1])
        if u - w >= 0 and u + w + 1 < width and v - w >= 0 and v + w + 1 < height:
            ii = np.arange(u - w, u + w + 1)
            jj = np.arange(v - w, v + w + 1)
            d2 = (ii[None, :] - u) ** 2 + (jj[:, None] - v) ** 2
            out[n, v - w:v + w + 1, u - w:u + w + 1] = np.exp(-(d2 / (2 * (sigma ** 2))))
    return out


def soft_argmax(heatmaps, beta, size_mult=1.0):
    """SoftArgmaxPavlo.forward, dream/spatial_softmax.py:24-95 (torch CPU fp32)."""
    import torch
    import torch.nn.functional as F
    h = torch.as_tensor(heatmaps, dtype=torch.float32)
    B, K, R, Cn = h.shape
    a = F.avg_pool2d(h, 7, stride=1, padding=3).reshape(B, K, -1)
    a = a - a.max(dim=2, keepdim=True)[0]
    e = torch.exp(torch.as_tensor(beta, dtype=torch.float32).view(1, K, 1) * a)
    nrm = (e / (e.sum(dim=2, keepdim=True) + 1e-8)).view(B, K, R, Cn)
    cols = (torch.arange(0, Cn) * size_mult).float().view(1, 1, 1, Cn)
    rows = (torch.arange(0, R) * size_mult).float().view(1, 1, R, 1)
    xs = (nrm * cols).reshape(B, K, -1).sum(dim=2)
    ys = (nrm * rows).reshape(B, K, -1).sum(dim=2)
    return torch.stack((xs, ys), dim=2)
