"""Build a variant of the CUDA library with extra nvcc flags: python tools/build_variant.py NAME -DHR_SEARCH_MAXNREG=88 ..."""
import sys, pathlib, subprocess
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import hr_pkg
hr_pkg.load()
from hopperrender_b200 import build as b
name, extra = sys.argv[1], sys.argv[2:]
out = ROOT / "tools" / "variants" / ("libhr_%s.so" % name)
out.parent.mkdir(exist_ok=True)
cmd = [b.nvcc_path(), *b.NVCC_FLAGS, *extra, "-o", str(out), str(b.CSRC / "hr_cuda.cu")]
r = subprocess.run(cmd, capture_output=True, text=True)
print(r.stdout, r.stderr)
print(out if r.returncode == 0 else "FAILED")
