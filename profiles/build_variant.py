"""Builds an experimental variant of the library next to the product one (scratch, git-ignored):
    python profiles/build_variant.py NAME [-DSCAN_WARPS_N=20 ...]   ->  variants/libigd_NAME.so
Only one source (rss_scan.cu unless named) is recompiled with the extra flags; run it with IGD_B200_LIB=variants/libigd_NAME.so."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from igdetective_b200 import build  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
SRC = "rss_scan.cu"
if extra and extra[0].endswith(".cu"):          # python profiles/build_variant.py NAME gotoh.cu -D...
    SRC, extra = extra[0], extra[1:]
build.build()
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
obj = os.path.join(ROOT, "variants", "%s_%s.o" % (SRC[:-3], name))
nvcc = build._nvcc()
r = subprocess.run([nvcc] + build.NVCC_FLAGS + extra + ["-c", os.path.join(build.CSRC, SRC), "-o", obj],
                   capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stderr)
for line in r.stderr.splitlines():
    if "rss_scan_kernelILb1" in line or "Used" in line or "spill" in line or "gotoh_packed_kernelILi32ELi11ELi0ELb0ELb1" in line:
        print(line.strip()[:160])
objs = [os.path.join(build.OBJ, s[:-3] + ".o") for s in build.SOURCES if s != SRC] + [obj]
lib = os.path.join(ROOT, "variants", "libigd_%s.so" % name)
subprocess.run([nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], check=True)
print(lib)
