import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the B200 box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """Path of libdreamb200.so (built on demand when nvcc is around)."""
    from dream_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.LIB_PATH


def panda_config(arch="vgg", **arch_extra):
    """Minimal network_config in the shape scripts/train_network.py assembles (train_network.py:259-323)."""
    names = ["panda_link0", "panda_link2", "panda_link3", "panda_link4", "panda_link6", "panda_link7", "panda_hand"]
    cfg = {
        "architecture": dict({
            "type": arch, "target": "belief_maps", "input_heads": ["image_rgb"], "output_heads": ["belief_maps"],
            "image_normalization": {"mean": [0.5, 0.5, 0.5], "stdev": [0.5, 0.5, 0.5]},
            "loss": {"type": "mse"}, "image_preprocessing": "shrink-and-crop",
        }, **arch_extra),
        "manipulator": {"name": "panda", "keypoints": [{"name": n, "friendly_name": n, "ros_frame": n} for n in names]},
        "training": {"config": {"net_input_resolution": [400, 400],
                                "optimizer": {"type": "adam", "learning_rate": 1.5e-4}},
                     "platform": {"gpu_ids": [0]}},
    }
    return cfg
