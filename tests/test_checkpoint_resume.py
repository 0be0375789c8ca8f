"""CPU: checkpoint / resume bookkeeping (dream_b200/checkpoint.py) against the on-disk contract of
the reference's scripts/train_network.py (:66-147 resume discovery, :326-407 consistency checks, :612-665 epoch files)."""
import copy
import os
import pickle

import pytest
import torch

from dream_b200 import checkpoint, models, network


class _StubNetwork:
    """The slice of DreamNetwork the bookkeeping touches, on the CPU: the facade's own save / enable_training methods
    over a small module wrapped like the real `.model` (DataParallelShim -> `module.` keys)."""
    save_network_config = network.DreamNetwork.save_network_config
    save_network_params = network.DreamNetwork.save_network_params
    save_network = network.DreamNetwork.save_network
    enable_training = network.DreamNetwork.enable_training

    def __init__(self, cfg, seed=0):
        torch.manual_seed(seed)
        self.model = models.DataParallelShim(torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Conv2d(4, 2, 1)))
        self.network_config = cfg
        self.optimizer = None


def _config():
    return {
        "data_path": "/data/panda", "manipulator": {"name": "panda"}, "architecture": {"type": "vgg"},
        "training": {"config": {"training_data_fraction": 0.8, "validation_data_fraction": 0.2, "batch_size": 4,
                                "data_augmentation": {"image_rgb": True}, "worker_size": 2,
                                "optimizer": {"type": "adam", "learning_rate": 1e-3},
                                "image_preprocessing": "shrink-and-crop", "image_raw_resolution": [640, 480],
                                "net_input_resolution": [400, 400]},
                     "results": {"epochs_trained": 0, "validation_loss": {"mean": 1.0, "stdev": 0.0}}},
    }


def _train_one_epoch(net):
    net.enable_training()
    x = torch.randn(2, 3, 8, 8)
    net.optimizer.zero_grad()
    net.model(x).square().mean().backward()
    net.optimizer.step()


def test_epoch_files_follow_the_reference_contract_and_resume_restores_everything(tmp_path):
    out = str(tmp_path)
    net = _StubNetwork(_config())
    log = {"random_seed": 1234, "start_time": 10.0, "epochs": []}
    for epoch in (1, 2, 3):
        _train_one_epoch(net)
        net.network_config["training"]["results"]["epochs_trained"] = epoch
        net.network_config["training"]["results"]["validation_loss"] = {"mean": 1.0 / epoch, "stdev": 0.01}
        log["epochs"].append(epoch)
        checkpoint.save_epoch(net, out, epoch, train_log=log, previous_epoch=epoch - 1, is_best=True)
    files = set(os.listdir(out))
    assert {"epoch_1.pth", "epoch_1.yaml", "epoch_3.pth", "epoch_3.yaml", "best_network.pth", "best_network.yaml",
            "optim_epoch_3.pt", "training_log_e3.pkl"} <= files
    assert "training_log_e2.pkl" not in files                       # rolling log (train_network.py:647-653)
    assert all(k.startswith("module.") for k in torch.load(os.path.join(out, "epoch_3.pth")))
    checkpoint.finish_training(out, 3)
    assert os.path.exists(os.path.join(out, "training_log.pkl"))

    assert [c[0] for c in checkpoint.list_epoch_checkpoints(out)] == [3, 2, 1]
    # the REFERENCE's resume scan over the same directory (scripts/train_network.py:70-85, restated): every entry it
    # takes for a weights file must be a model state dict -- the optimizer file must be invisible to it
    ref_src = "/root/reference/scripts/train_network.py"
    if os.path.exists(ref_src):                                     # build container only: the rule is still the one restated
        assert 'x.startswith("epoch") and x.endswith(".pth")' in open(ref_src).read()
    picked = [x for x in os.listdir(out) if x.startswith("epoch") and x.endswith(".pth")]
    numbers = [int(x.split("_")[1].split(".")[0]) for x in picked]
    assert sorted(numbers) == [1, 2, 3]                             # one candidate per epoch: no tie for the newest
    newest = sorted(zip(picked, numbers), key=lambda pair: pair[1], reverse=True)[0][0]
    assert newest == "epoch_3.pth"
    for name in picked:
        assert all(k.startswith("module.") for k in torch.load(os.path.join(out, name)))
    rp = checkpoint.find_resume_point(out, new_network_config=_config(), total_epochs=5)
    assert rp.start_epoch == 3 and rp.random_seed == 1234
    assert rp.best_valid_loss == pytest.approx(1.0 / 3)
    assert rp.train_log["start_time"][0] == 10.0 and len(rp.train_log["start_time"]) == 2
    assert rp.train_log["epochs_resumed"] == [4]
    assert os.path.exists(os.path.join(out, "training_log_e3.pkl")) and not os.path.exists(os.path.join(out, "training_log.pkl"))
    assert rp.network_config["training"]["results"]["epochs_trained"] == 3   # "use this one instead"

    fresh = checkpoint.restore(_StubNetwork(copy.deepcopy(rp.network_config), seed=99), rp)
    for (k, a), (_, b) in zip(net.model.state_dict().items(), fresh.model.state_dict().items()):
        assert torch.equal(a, b), k
    sa, sb = net.optimizer.state_dict()["state"], fresh.optimizer.state_dict()["state"]
    assert sa.keys() == sb.keys() and all(torch.equal(sa[k]["exp_avg"], sb[k]["exp_avg"]) for k in sa)
    # continuing from the restored state == continuing the original run (Adam moments included)
    torch.manual_seed(5); _train_one_epoch(net)
    torch.manual_seed(5); _train_one_epoch(fresh)
    for a, b in zip(net.model.parameters(), fresh.model.parameters()):
        assert torch.equal(a, b)


def test_resume_refuses_finished_runs_changed_configs_and_missing_files(tmp_path):
    out = str(tmp_path)
    with pytest.raises(AssertionError):
        checkpoint.find_resume_point(out)                            # nothing there
    net = _StubNetwork(_config())
    _train_one_epoch(net)
    checkpoint.save_epoch(net, out, 1, train_log={"random_seed": 1, "start_time": 0.0}, is_best=False, save_optimizer=False)
    with pytest.raises(AssertionError, match="best validation loss"):
        checkpoint.find_resume_point(out)                            # train_network.py:98-100
    checkpoint.save_epoch(net, out, 1, is_best=True, save_optimizer=False)
    with pytest.raises(AssertionError, match="already trained"):
        checkpoint.find_resume_point(out, total_epochs=1)            # train_network.py:91-93
    changed = _config()
    changed["training"]["config"]["batch_size"] = 8
    with pytest.raises(AssertionError, match="batch_size"):
        checkpoint.find_resume_point(out, new_network_config=changed, load_log=False)
    rp = checkpoint.find_resume_point(out, new_network_config=_config(), total_epochs=2)
    assert rp.optimizer_path is None                                 # reference-style checkpoint: moments restart
    restored = checkpoint.restore(_StubNetwork(_config(), seed=3), rp)
    assert restored.optimizer is not None and len(restored.optimizer.state_dict()["state"]) == 0
    os.remove(os.path.join(out, "training_log_e1.pkl"))
    with pytest.raises(AssertionError, match="training log"):
        checkpoint.find_resume_point(out)                            # train_network.py:128-129


def test_async_writer_writes_the_same_files(tmp_path):
    a_dir, b_dir = str(tmp_path / "sync"), str(tmp_path / "async")
    net = _StubNetwork(_config())
    _train_one_epoch(net)
    checkpoint.save_epoch(net, a_dir, 7, is_best=True)
    writer = checkpoint.AsyncCheckpointWriter()
    checkpoint.save_epoch(net, b_dir, 7, is_best=True, writer=writer)
    # the snapshot is taken at submit time: later updates must not leak into the files
    with torch.no_grad():
        for p in net.model.parameters():
            p.add_(1.0)
    net.network_config["training"]["results"]["epochs_trained"] = 99
    writer.close()
    assert sorted(os.listdir(a_dir)) == sorted(os.listdir(b_dir))
    for name in ("epoch_7.pth", "best_network.pth"):
        sa, sb = torch.load(os.path.join(a_dir, name)), torch.load(os.path.join(b_dir, name))
        assert sa.keys() == sb.keys() and all(torch.equal(sa[k], sb[k]) for k in sa)
    assert network.load_yaml_config(os.path.join(b_dir, "epoch_7.yaml"))["training"]["results"]["epochs_trained"] == 0
    oa, ob = torch.load(os.path.join(a_dir, "optim_epoch_7.pt")), torch.load(os.path.join(b_dir, "optim_epoch_7.pt"))
    assert all(torch.equal(oa["state"][k]["exp_avg_sq"], ob["state"][k]["exp_avg_sq"]) for k in oa["state"])
