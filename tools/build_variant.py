"""
Developer tool: build a variant of libgcs_b200.so with extra -D flags on chosen sources, for A/B timing on the GPU box.

    python tools/build_variant.py NAME "-DGCS_TC_FETCH_AT=1 -DGCS_TC_SELF_ISSUE=0" [source.cu[=/other/version.cu] ...]
                                                                                     (default: gcs_bins_tc.cu)

Writes gc-slam_b200/lib/variants/libgcs_b200.NAME.so (git-ignored, travels with the gpurun snapshot); select it with
GCS_B200_LIB=<path>.  The other objects are taken from the last regular build (gc-slam_b200/lib/*.o).
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gc_slam_b200 import build as B  # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2].split()
    srcs = sys.argv[3:] or ["gcs_bins_tc.cu"]
    B.build()
    vdir = os.path.join(B.LIBDIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = {os.path.basename(p)[:-3]: os.path.join(B.LIBDIR, os.path.basename(p)[:-3] + ".o") for p in B._sources()}
    for src in srcs:
        src, _, alt = src.partition("=")       # name.cu=/path/to/another/version.cu (e.g. `git show HEAD:...` output)
        stem = src[:-3]
        obj = os.path.join(vdir, f"{stem}.{name}.o")
        cmd = [nvcc, *B.NVCC_FLAGS, *defs, "-I", B.INCLUDE, "-I", B.CSRC, "-c", alt or os.path.join(B.CSRC, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise SystemExit(1)
        for line in r.stdout.splitlines():
            if "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
                print("  ", line.strip())
        objs[stem] = obj
    lib = os.path.join(vdir, f"libgcs_b200.{name}.so")
    r = subprocess.run([nvcc, "-shared", "-o", lib, *objs.values(), "-Xcompiler", "-fPIC", "-cudart", "static"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise SystemExit(1)
    print(lib)


if __name__ == "__main__":
    main()
