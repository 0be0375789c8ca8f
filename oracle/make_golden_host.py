"""Generate the small host-side fixtures tests/golden/{frames,targets}.npz from the REFERENCE's own
dream/image_proc.py (build container only; /root/reference does not travel):
  * keypoint frame conversions (image_proc.py:135-260) for every preprocessing type,
  * create_belief_map training targets (image_proc.py:866-910).
Run:  python oracle/make_golden_host.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import GOLD, load_reference  # noqa: E402


def main():
    _, _, image_proc = load_reference()
    rng = np.random.default_rng(11)
    out = {}
    cases = [("shrink-and-crop", (640, 480), (400, 400)), ("resize", (640, 480), (400, 400)),
             ("none", (400, 400), (400, 400)), ("shrink-and-crop", (480, 640), (400, 400)),
             ("shrink", (640, 480), (400, 400))]
    for i, (preproc, raw, net_in) in enumerate(cases):
        if preproc == "shrink":
            net_in = image_proc.resolution_after_preprocessing(raw, net_in, preproc)
        net_out = (net_in[0] // 4, net_in[1] // 4)
        kps = rng.uniform(0, net_out[0], size=(16, 2))
        kps[3] = -999.999
        netin = np.array(image_proc.convert_keypoints_to_netin_from_netout(kps, net_out, net_in))
        raw_kp = np.array(image_proc.convert_keypoints_to_raw_from_netin(netin, net_in, raw, preproc))
        back = np.array(image_proc.convert_keypoints_to_netin_from_raw(raw_kp, raw, net_in, preproc))
        out["case%d::meta" % i] = np.array([preproc, raw[0], raw[1], net_in[0], net_in[1]], dtype=object).astype(str)
        out["case%d::kps" % i] = kps
        out["case%d::netin" % i] = netin
        out["case%d::raw" % i] = raw_kp
        out["case%d::back" % i] = back
    np.savez_compressed(os.path.join(GOLD, "frames.npz"), **out)
    print("frames", len(cases))

    pts = rng.uniform(-3, 103, size=(40, 2))
    pts[0] = (4.0, 4.0); pts[1] = (3.99, 50.0); pts[2] = (94.0, 94.0); pts[3] = (95.0, 50.0)   # window edge cases
    pts[4] = (50.5, 94.999); pts[5] = (-999.999, -999.999)
    tgt = image_proc.create_belief_map((100, 100), pts, sigma=2)
    pts2 = rng.uniform(0, 208, size=(7, 2))
    tgt2 = image_proc.create_belief_map((208, 160), pts2 * np.array([1.0, 160 / 208.0]), sigma=2)
    np.savez_compressed(os.path.join(GOLD, "targets.npz"), pts=pts, tgt=tgt.astype(np.float32), tgt64_sum=tgt.sum(),
                        pts2=pts2 * np.array([1.0, 160 / 208.0]), tgt2=tgt2.astype(np.float32))
    print("targets", tgt.shape, tgt2.shape, float(tgt.sum()))

    # the dataset's image tensor transform (dream/datasets.py:60-75: Compose([ToTensor(), Normalize(mean, std)]))
    import torchvision.transforms as TVTransforms
    from PIL import Image as PILImage
    img = rng.integers(0, 256, size=(24, 36, 3), dtype=np.uint8)
    img.reshape(-1)[:768] = np.repeat(np.arange(256, dtype=np.uint8), 3)      # every byte value in every channel
    out = {"img": img}
    for tag, mean, std in (("half", (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
                           ("imagenet", (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        tform = TVTransforms.Compose([TVTransforms.ToTensor(), TVTransforms.Normalize(mean, std)])
        out[tag + "::mean"] = np.array(mean, dtype=np.float32)
        out[tag + "::std"] = np.array(std, dtype=np.float32)
        out[tag + "::x"] = tform(PILImage.fromarray(img)).numpy()
    np.savez_compressed(os.path.join(GOLD, "normalize.npz"), **out)
    print("normalize", out["half::x"].shape)


if __name__ == "__main__":
    main()
