"""Drop-in check of the `dream` overlay (shim/dream): with it in front of a reference checkout on PYTHONPATH the
reference's UNMODIFIED host code -- dream/analysis.py:93-277 behind scripts/network_inference_dataset.py -- imports,
walks its argument checks on a synthetic NDDS folder and constructs the network through OUR facade.  In the build
container (no GPU) that ends at the facade's loud "needs a CUDA device" error, which is exactly the evidence wanted
here: the reference's analysis loop reached dream_b200.network.DreamNetwork with its own config, and there is no CPU
fallback behind it.  Skipped where /root/reference does not exist (the GPU box); on a GPU the same script goes on
into inference (tests/test_gpu_dropin.py drives that part without the reference's third-party dependencies).

The reference's third-party imports that this image lacks (matplotlib, ruamel.yaml, albumentations, pyrr, webcolors,
seaborn, tensorboardX) are stubbed in the child process; ruamel's YAML is backed by PyYAML."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DREAM_REFERENCE", "/root/reference")

CHILD = textwrap.dedent('''
    import json, os, sys, types
    # ---- stubs for the reference's third-party imports missing from this image ----
    import yaml as _pyyaml
    for name in ("matplotlib", "matplotlib.pyplot", "webcolors", "albumentations", "pyrr", "seaborn", "tensorboardX",
                 "ruamel", "ruamel.yaml"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pyrr"].Quaternion = type("Quaternion", (), {})
    class _YAML:
        def __init__(self, typ=None): pass
        def load(self, f): return _pyyaml.safe_load(f)
        def dump(self, data, f): _pyyaml.safe_dump(data, f)
    sys.modules["ruamel.yaml"].YAML = _YAML
    sys.modules["ruamel"].yaml = sys.modules["ruamel.yaml"]

    import dream, dream_b200.network, dream_b200.models, dream_b200.image_proc
    out = {
        "shim_file": dream.__file__,
        "network_is_ours": dream.DreamNetwork is dream_b200.network.DreamNetwork
                           and dream.network.DreamNetwork is dream_b200.network.DreamNetwork
                           and dream.create_network_from_config_data is dream_b200.network.create_network_from_config_data,
        "models_are_ours": dream.models.DreamHourglass is dream_b200.models.DreamHourglass
                           and dream.ResnetSimple is dream_b200.models.ResnetSimple,
        "peaks_are_ours": dream.image_proc.peaks_from_belief_maps is dream_b200.image_proc.peaks_from_belief_maps
                          and dream.peaks_from_belief_maps is dream_b200.image_proc.peaks_from_belief_maps,
        "analysis_file": dream.analysis.__file__, "datasets_file": dream.datasets.__file__,
        "utilities_file": dream.utilities.__file__,
        "reference_helpers_present": hasattr(dream.image_proc, "convert_keypoints_to_raw_from_netin")
                                     and hasattr(dream, "ManipulatorNDDSDataset") and hasattr(dream, "solve_pnp"),
    }
    # ---- the reference's analysis entry point on a synthetic NDDS folder ----
    work = sys.argv[1]
    try:
        dream.analysis.analyze_ndds_dataset(os.path.join(work, "net.pth"), os.path.join(work, "net.yaml"),
                                            os.path.join(work, "ndds"), os.path.join(work, "out"),
                                            visualize_belief_maps=False, pnp_analysis=False, force_overwrite=True,
                                            batch_size=4, num_workers=0, gpu_ids=[0])
        out["analysis"] = "completed"
    except Exception as e:
        import traceback
        tb = traceback.extract_tb(e.__traceback__)
        out["analysis"] = "%s: %s" % (type(e).__name__, e)
        out["raised_in"] = [os.path.relpath(f.filename, "/") + ":" + f.name for f in tb][-3:]
    print("RESULT " + json.dumps(out))
''')


def _write_inputs(work):
    import torch
    import yaml
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import panda_config
    os.makedirs(os.path.join(work, "ndds"), exist_ok=True)
    with open(os.path.join(work, "net.yaml"), "w") as f:
        yaml.safe_dump(panda_config("vgg"), f)
    torch.save({}, os.path.join(work, "net.pth"))
    from PIL import Image
    json.dump({"camera_settings": [{"name": "camera", "captured_image_size": {"width": 640, "height": 480},
                                    "intrinsic_settings": {"fx": 320.0, "fy": 320.0, "cx": 320.0, "cy": 240.0, "s": 0}}]},
              open(os.path.join(work, "ndds", "_camera_settings.json"), "w"))
    json.dump({"exported_objects": []}, open(os.path.join(work, "ndds", "_object_settings.json"), "w"))
    for i in range(3):
        Image.new("RGB", (640, 480), (i * 40, 10, 10)).save(os.path.join(work, "ndds", "%06d.rgb.jpg" % i))
        json.dump({"objects": []}, open(os.path.join(work, "ndds", "%06d.json" % i), "w"))


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "dream", "analysis.py")),
                    reason="needs the reference checkout (build container only)")
def test_reference_analysis_runs_unmodified_over_the_overlay(tmp_path):
    _write_inputs(str(tmp_path))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "shim"), ROOT, REF]), CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, "-c", CHILD, str(tmp_path)], env=env, capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, (p.stdout[-2000:], p.stderr[-4000:])
    out = json.loads(lines[-1][7:])
    assert out["shim_file"].startswith(os.path.join(ROOT, "shim"))
    assert out["network_is_ours"] and out["models_are_ours"] and out["peaks_are_ours"], out
    for k in ("analysis_file", "datasets_file", "utilities_file"):                  # the reference's own files, untouched
        assert out[k].startswith(os.path.join(REF, "dream")), out
    assert out["reference_helpers_present"], out
    # the reference's loop got as far as building the network through our facade, which refuses to run without CUDA
    assert out["analysis"].startswith("RuntimeError") and "CUDA" in out["analysis"], out
    assert any("dream_b200/network.py" in f for f in out["raised_in"]), out
    assert any("dream/analysis.py:analyze_ndds_dataset" in f for f in out["raised_in"]), out
