import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_training as T
for key, kind, m in T.resnet_unit_report(every=5):
    print("%-28s %-26s %s" % (key, kind, "  ".join("%s rel %.4f cos %.6f" % (n, r, c) for n, (r, c) in m.items())))
