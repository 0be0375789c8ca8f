"""`dream` overlay: the reference's package name with the hot path swapped for dream_b200.

Put THIS directory's parent in front of the reference checkout on PYTHONPATH (or set DREAM_REFERENCE to the checkout):

    PYTHONPATH=/path/to/dream_b200_repo/shim:/path/to/dream_b200_repo:/path/to/DREAM  python scripts/network_inference_dataset.py ...

`import dream` then resolves
    dream.network, dream.models, dream.spatial_softmax      -> dream_b200 (libdreamb200.so kernels)
    dream.image_proc                                        -> the reference's module with peaks_from_belief_maps
                                                               replaced by the device kernel
    dream.analysis, dream.datasets, dream.utilities, dream.geometric_vision, ...  -> the reference's own files
so `scripts/network_inference_dataset.py`, `scripts/train_network.py`, `launch_dream_ros.py` run unmodified
(SURVEY.md 8b).  Nothing of the reference is copied: its modules are loaded from where they lie.
"""
import importlib
import os
import sys

__version__ = "1.3.0+dream_b200"

_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_package_dir():
    cand = []
    if os.environ.get("DREAM_REFERENCE"):
        cand.append(os.path.join(os.environ["DREAM_REFERENCE"], "dream"))
    for entry in sys.path:
        d = os.path.join(entry or ".", "dream")
        if os.path.abspath(d) != _HERE:
            cand.append(d)
    for d in cand:
        if os.path.isfile(os.path.join(d, "analysis.py")) and os.path.isfile(os.path.join(d, "datasets.py")):
            return os.path.abspath(d)
    return None


_REF = _reference_package_dir()
# submodule search order: the overlay's own modules first, everything else from the reference checkout
__path__ = [_HERE] + ([_REF] if _REF else [])

# the hot path (same public names as dream/network.py, dream/models.py, dream/spatial_softmax.py)
from . import spatial_softmax, models, network          # noqa: E402,F401
from .network import *                                   # noqa: E402,F401,F403
from .models import *                                    # noqa: E402,F401,F403

if _REF is not None:
    # host-side modules of the reference, in its own import order (dream/__init__.py:3-9); `image_proc` is the
    # overlay's wrapper around the reference's module
    from . import image_proc                             # noqa: E402,F401
    from .image_proc import *                            # noqa: E402,F401,F403
    for _name in ("utilities", "geometric_vision", "datasets", "analysis"):
        _mod = importlib.import_module("dream." + _name)
        globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("_")})
else:
    from dream_b200 import image_proc                    # noqa: E402,F401  (the subset the facade needs)
    sys.modules[__name__ + ".image_proc"] = image_proc
