"""dream.image_proc: the reference's module (loaded from its checkout, untouched) with `peaks_from_belief_maps`
(dream/image_proc.py:914-1018) replaced by the device kernel behind dream_b200.image_proc."""
import importlib.util
import os

import dream as _pkg
from dream_b200 import image_proc as _ours

_ref_file = os.path.join(_pkg._REF, "image_proc.py")
_spec = importlib.util.spec_from_file_location("dream._reference_image_proc", _ref_file)
_ref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_ref)
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
reference_peaks_from_belief_maps = _ref.peaks_from_belief_maps
peaks_from_belief_maps = _ours.peaks_from_belief_maps
