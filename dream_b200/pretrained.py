"""ImageNet trunk weights for freshly constructed networks, WITHOUT downloading.

The reference builds its trunks from `torchvision.models.vgg19(pretrained=True)` / `resnet101(pretrained=...)`
(dream/models.py:22, 587), which downloads `vgg19-dcbb9e9d.pth` / `resnet101-*.pth` on first use.  This package never
touches the network: it looks for the same files locally --

    $DREAMB200_PRETRAINED_DIR,  then  torch.hub.get_dir()/checkpoints  (= $TORCH_HOME/hub/checkpoints)

-- and copies the matching tensors into the parameter tree.  When no file is found the trunk keeps its random
initialisation and a warning says so: a run of scripts/train_network.py started that way does not reproduce the
reference's convergence (loading a released DREAM `.pth` afterwards overrides everything, so inference is unaffected).
"""
import glob
import os
import warnings

import torch

_WARNED = set()


def _find(patterns):
    dirs = []
    if os.environ.get("DREAMB200_PRETRAINED_DIR"):
        dirs.append(os.environ["DREAMB200_PRETRAINED_DIR"])
    try:
        dirs.append(os.path.join(torch.hub.get_dir(), "checkpoints"))
    except Exception:
        pass
    for d in dirs:
        for pat in patterns:
            hits = sorted(glob.glob(os.path.join(d, pat)))
            if hits:
                return hits[0]
    return None


def _warn(kind, patterns):
    if kind in _WARNED:
        return
    _WARNED.add(kind)
    warnings.warn(
        "dream_b200: no local torchvision {} checkpoint ({}); the trunk keeps its RANDOM initialisation, unlike the "
        "reference (pretrained ImageNet weights, dream/models.py:22,587).  Put the file into $DREAMB200_PRETRAINED_DIR "
        "or torch's hub cache to reproduce the reference's training start; loading a DREAM .pth is unaffected."
        .format(kind, " / ".join(patterns)), stacklevel=3)


def load_vgg19_trunk(model):
    """layer_0_k_down.{j} <- vgg19.features.{j} for every trunk conv except the fresh first one (models.py:591-615).
    Returns the path used or None."""
    patterns = ("vgg19-*.pth",)
    path = _find(patterns)
    if path is None:
        _warn("vgg19", patterns)
        return None
    sd = torch.load(path, map_location="cpu")
    own = dict(model.named_parameters())
    with torch.no_grad():
        for name, p in own.items():
            if not name.startswith("layer_0_") or name.startswith("layer_0_1_down.0."):
                continue
            idx, leaf = name.split(".")[1], name.split(".")[2]
            src = sd["features.%s.%s" % (idx, leaf)]
            assert tuple(src.shape) == tuple(p.shape), (name, tuple(src.shape), tuple(p.shape))
            p.copy_(src)
    return path


def load_resnet101_trunk(model):
    """conv1 / bn1 / layer1..4 <- torchvision resnet101 (same key names, models.py:22-32).  Returns the path or None."""
    patterns = ("resnet101-*.pth",)
    path = _find(patterns)
    if path is None:
        _warn("resnet101", patterns)
        return None
    sd = torch.load(path, map_location="cpu")
    own = dict(model.named_parameters())
    own.update(dict(model.named_buffers()))
    with torch.no_grad():
        for name, t in own.items():
            if name.split(".")[0] in ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4") and name in sd:
                assert tuple(sd[name].shape) == tuple(t.shape), name
                t.copy_(sd[name])
    return path
