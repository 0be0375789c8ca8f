"""Generate tests/golden/net_ms*.npz from the REFERENCE's DreamHourglassMultiStage (dream/models.py:350-553),
CPU fp32, with oracle.ref_models.multistage_state_dict weights (regenerable from names; only the per-stage head
gains are stored).  Build container only.   Run:  python oracle/make_golden_multistage.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_models  # noqa: E402
from oracle.make_golden import GOLD, load_reference  # noqa: E402


def main():
    _, models, _ = load_reference()
    torch.set_num_threads(8)
    cases = {
        "ms2": (dict(n_stages=2), (2, 3, 64, 80)),
        "ms3_full": (dict(n_stages=3, full_output=True), (1, 3, 32, 48)),
    }
    for name, (kw, xshape) in cases.items():
        S = kw["n_stages"]
        net = models.DreamHourglassMultiStage(7, internalize_spatial_softmax=False, **kw).eval()
        g = torch.Generator().manual_seed(43)
        x = torch.rand(xshape, generator=g) * 2 - 1
        mk = dict(n_keypoints=7, n_stages=S, full_output=kw.get("full_output", False), prefix="")
        gains = [1.0] * S
        for s in range(S):            # stage by stage: scale each head so its belief maps peak at ~1
            net.load_state_dict(ref_models.multistage_state_dict(gains=gains, **mk))
            with torch.no_grad():
                gains[s] = float(np.float32(1.0 / net(x)[s].abs().max().item()))
        sd = ref_models.multistage_state_dict(gains=gains, **mk)
        assert list(net.state_dict().keys()) == list(sd.keys()), "key order differs"
        net.load_state_dict(sd)
        with torch.no_grad():
            ys = net(x)
        out = {"x": x.numpy(), "gains": np.array(gains, dtype=np.float64)}
        for s, y in enumerate(ys):
            out["y%d" % (s + 1)] = y.numpy()
        # the multi-stage loss of DreamNetwork.loss (network.py:345-352) and a few gradients, including stage-1
        # parameters that the later stages' losses reach through the concatenated belief maps
        net.train()
        net.zero_grad()
        tg = torch.rand(ys[0].shape, generator=g)
        outs = net(x)
        loss = torch.nn.MSELoss()(torch.stack(outs), tg.unsqueeze(0).expand([S] + [-1] * tg.dim()))
        loss.backward()
        out["target"] = tg.numpy()
        out["loss"] = np.float64(loss.item())
        params = dict(net.named_parameters())
        for k in ("stage1.layer_0_1_down.0.weight", "stage1.heads_0.4.weight", "stage1.layer_0_3_down.12.weight",
                  "stage2.layer_0_1_down.0.weight", "stage2.heads_0.2.weight", "stage%d.heads_0.4.weight" % S):
            gk = params[k].grad.numpy()
            out["grad::" + k] = gk if gk.size <= 100000 else gk[:4]
        np.savez_compressed(os.path.join(GOLD, "net_%s.npz" % name), **out)
        print(name, [tuple(y.shape) for y in ys], gains, float(loss))


if __name__ == "__main__":
    main()
