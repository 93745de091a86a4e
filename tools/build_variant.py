"""A/B builds of libanimerec.so (developer tool): python tools/build_variant.py NAME -DAR_STEP_WARPS=12 ...
writes anime_recommendations_b200/lib/variants/libanimerec_NAME.so; select it with ANIMEREC_LIB=<path>."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anime_recommendations_b200 import build as b  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
b.build_lib()
out_dir = os.path.join(b.LIBDIR, "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, "train_%s.o" % name)
cmd = [b._nvcc()] + b.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-c", os.path.join(b.CSRC, "train.cu"), "-o", obj]
r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
if r.returncode:
    sys.exit(r.stdout)
lines = r.stdout.splitlines()
for i, ln in enumerate(lines):
    if "chunk_kernelILi1E" in ln and "Compiling" in ln:
        print(name, lines[i + 2].strip(), "|", lines[i + 3].strip() if i + 3 < len(lines) else "")
objs = [os.path.join(b.LIBDIR, s.replace(".cu", ".o")) for s in b.SOURCES if s != "train.cu"] + [obj]
lib = os.path.join(out_dir, "libanimerec_%s.so" % name)
subprocess.check_call([b._nvcc(), "-shared", "-o", lib] + objs)
print(lib)
