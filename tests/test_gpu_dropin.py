"""GPU part of the drop-in evidence (the reference checkout does not exist on the GPU box, so its scripts cannot run
there; tests/test_dropin_shim.py runs them, unmodified, up to the facade in the build container).  Here the SAME call
sequences are driven through the `dream` overlay name:
  * scripts/network_inference_dataset.py -> dream/analysis.py:139-277: config from YAML, `create_network_from_config_data`,
    `model.load_state_dict(torch.load(path))`, `enable_evaluation`, per batch `.cuda()` -> `inference` -> keypoint frame
    conversions;
  * scripts/train_network.py:403-507, 612-665: `enable_training`, per batch `train([x], target)` + `loss.item()`,
    validation `loss(...)` under no_grad, per-epoch `save_network`, then a resume from the files just written
    (dream_b200.checkpoint.find_resume_point / restore) that continues training."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import ref_models, ref_peaks  # noqa: E402  (checker only)


@pytest.fixture()
def dream_overlay(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(ROOT, "shim"))
    monkeypatch.delenv("DREAM_REFERENCE", raising=False)
    for k in [k for k in sys.modules if k == "dream" or k.startswith("dream.")]:
        monkeypatch.delitem(sys.modules, k)
    import dream
    yield dream
    for k in [k for k in sys.modules if k == "dream" or k.startswith("dream.")]:
        sys.modules.pop(k, None)


def _config(res=(128, 96)):
    from conftest import panda_config
    cfg = panda_config("vgg")
    cfg["training"]["config"]["net_input_resolution"] = list(res)
    cfg["training"]["config"]["epochs"] = 2
    cfg["training"]["config"]["batch_size"] = 2
    return cfg


def test_inference_script_call_sequence_through_the_overlay(dream_overlay, tmp_path, built_lib):
    dream = dream_overlay
    import dream_b200.network
    assert dream.DreamNetwork is dream_b200.network.DreamNetwork
    cfg_path, params_path = str(tmp_path / "net.yaml"), str(tmp_path / "net.pth")
    dream_b200.network.dump_yaml_config(_config(), cfg_path)
    sd = ref_models.synth_state_dict(ref_models.vgg_state_shapes(7), seed=8, out_gain=13.0, mode="default")
    torch.save(sd, params_path)
    # analysis.py:139-155
    network_config = dream_b200.network.load_yaml_config(cfg_path)
    network_config["training"]["platform"]["gpu_ids"] = [0]
    net = dream.create_network_from_config_data(network_config)
    net.model.load_state_dict(torch.load(params_path))
    net.enable_evaluation()
    in_res, out_res = net.net_resolutions_from_image_raw_resolution((640, 480))
    assert tuple(in_res) == (128, 96) and tuple(out_res) == (32, 24)
    g = torch.Generator().manual_seed(3)
    frames = torch.rand((5, 3, 96, 128), generator=g) * 2 - 1
    rows, maps = [], []
    with torch.no_grad():
        for i in range(0, 5, 2):                                  # analysis.py:200-262, batch_size 2, ragged last batch
            batch = frames[i:i + 2]
            belief, kps = net.inference(batch.cuda())
            assert kps.shape == (batch.shape[0], 7, 2) and not kps.is_cuda and belief.is_cuda
            maps.append(belief.cpu().numpy())
            for b in range(batch.shape[0]):
                kp_in = dream.image_proc.convert_keypoints_to_netin_from_netout(kps[b].numpy(), out_res, in_res)
                rows.append(dream.image_proc.convert_keypoints_to_raw_from_netin(kp_in, in_res, (640, 480),
                                                                                 net.image_preprocessing()))
    maps = np.concatenate(maps)
    # belief maps vs the oracle model; keypoints vs the oracle's peak extraction + decision table on the same maps
    want_maps = ref_models.vgg_forward(sd, frames).numpy()
    assert np.abs(maps - want_maps).max() <= 1e-3 * max(1.0, np.abs(want_maps).max())
    for b in range(5):
        ref_k = np.array(ref_peaks.select_keypoints(ref_peaks.peaks_from_belief_maps(maps[b], 0.4395)), dtype=np.float32)
        kp_in = dream.image_proc.convert_keypoints_to_netin_from_netout(ref_k, out_res, in_res)
        want = dream.image_proc.convert_keypoints_to_raw_from_netin(kp_in, in_res, (640, 480), net.image_preprocessing())
        assert np.array_equal(rows[b], want), b


def test_training_script_call_sequence_and_resume(dream_overlay, tmp_path, built_lib):
    dream = dream_overlay
    from dream_b200 import checkpoint
    out_dir = str(tmp_path / "run")
    cfg = _config()
    net = dream.create_network_from_config_data(cfg)             # train_network.py:403
    net.enable_training()                                        # :407
    g = torch.Generator().manual_seed(4)
    data = [(torch.rand((2, 3, 96, 128), generator=g) * 2 - 1, torch.rand((2, 7, 24, 32), generator=g)) for _ in range(3)]
    log = {"random_seed": 7, "start_time": 0.0, "epochs": []}
    best = float("inf")
    losses = []
    for epoch in (1, 2):
        net.enable_training()
        for x, t in data[:2]:                                     # :478-507
            loss = net.train([x.cuda()], t.cuda())
            losses.append(loss.item())
        net.enable_evaluation()
        with torch.no_grad():                                     # :520-560 validation
            val = float(np.mean([net.loss([x.cuda()], t.cuda()).item() for x, t in data[2:]]))
        net.network_config["training"]["results"] = {"epochs_trained": epoch, "validation_loss": {"mean": val, "stdev": 0.0}}
        log["epochs"].append(epoch)
        checkpoint.save_epoch(net, out_dir, epoch, train_log=log, previous_epoch=epoch - 1, is_best=val < best)
        best = min(best, val)
    assert losses[-1] < losses[0]
    assert {"epoch_2.pth", "epoch_2.yaml", "best_network.pth", "best_network.yaml", "optim_epoch_2.pt"} <= set(os.listdir(out_dir))
    # resume (train_network.py:66-147, 326-407) from the files just written
    rp = checkpoint.find_resume_point(out_dir, new_network_config=_config(), total_epochs=4)
    assert rp.start_epoch == 2 and rp.random_seed == 7 and rp.best_valid_loss == pytest.approx(best)
    fresh = dream.create_network_from_config_data(rp.network_config)
    checkpoint.restore(fresh, rp, map_location="cuda")
    for (ka, a), (kb, b) in zip(net.model.state_dict().items(), fresh.model.state_dict().items()):
        assert ka == kb and torch.equal(a, b)
    x, t = data[0]
    net.enable_training()                                         # (the epoch loop left it in evaluation mode)
    la, lb = net.train([x.cuda()], t.cuda()).item(), fresh.train([x.cuda()], t.cuda()).item()
    assert abs(la - lb) <= 1e-6 * abs(la)                         # same weights AND same Adam moments: same step
    for a, b in zip(net.model.parameters(), fresh.model.parameters()):
        assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(a.abs().max()))
