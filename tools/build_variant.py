"""Build a variant of libdreamb200.so from a source tree (A/B measurements): python tools/build_variant.py NAME SRC_ROOT
SRC_ROOT holds dream_b200/csrc and include/ (e.g. a `git archive` of another revision).  Output: variants/NAME.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dream_b200 import build as B
name, src = sys.argv[1], os.path.abspath(sys.argv[2])
extra = sys.argv[3:]
csrc = os.path.join(src, "dream_b200", "csrc")
out_dir = os.path.join(ROOT, "variants"); os.makedirs(out_dir, exist_ok=True)
obj_dir = os.path.join("/tmp", "variant_" + name); os.makedirs(obj_dir, exist_ok=True)
procs, objs = [], []
for s in B.SOURCES:
    o = os.path.join(obj_dir, s + ".o"); objs.append(o)
    procs.append(subprocess.Popen([B._nvcc()] + B.NVCC_FLAGS + extra + ["-I", os.path.join(src, "include"), "-I", csrc, "-c",
                                   os.path.join(csrc, s), "-o", o]))
assert all(p.wait() == 0 for p in procs)
lib = os.path.join(out_dir, name + ".so")
subprocess.check_call([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
print(lib)
