"""Developer tool: build a variant of the library into experiments/_libs/<name>.so with extra nvcc flags
(usage: build_variant.py name [-DRR_X=1 ...]); tools/ab_perf.py then times the variants side by side on one box."""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rustyread_b200 import _build
name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(os.path.dirname(_build.HERE), "experiments", "_libs", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cus = [s for s in _build.sources() if s.endswith((".cu", ".cpp"))]
subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + _build.NVCC_FLAGS + extra + ["-o", out] + cus)
print(out)
