"""Builds csrc/*.cu into libcy4.so (in-tree, next to this file) for sm_100a only.

    python complex-yolov4-pytorch_b200/csrc/build.py [--force] [-v]

Per-file flags: the geometry / loss-head translation units are compiled with --fmad=false so that
every fp32 product and sum rounds separately, as the reference's torch CPU ops do.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
# CY4_LIB_NAME / CY4_EXTRA_NVCC_FLAGS: side builds for experiments (e.g. libcy4_probe.so with -DCY4_PROBE);
# the product is always libcy4.so with the default flags.
OUT = os.path.join(HERE, os.environ.get("CY4_LIB_NAME", "libcy4.so"))
OBJ = os.path.join(HERE, "build" if OUT.endswith("libcy4.so") else "build_" + os.path.basename(OUT).split(".")[0])
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "--expt-extended-lambda", "--expt-relaxed-constexpr",
          "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-DCY4_BUILD"]
COMMON += os.environ.get("CY4_EXTRA_NVCC_FLAGS", "").split()      # e.g. -DCY4_PROBE for tools/probe_pipeline.py
PER_FILE = {
    "rgiou.cu": ["--fmad=false"],
    "yolo_head.cu": ["--fmad=false"],
    "nms.cu": ["--fmad=false"],
}


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _digest(path, flags):
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    for f in sorted(os.listdir(HERE)):
        if f.endswith((".cuh", ".h")) or f == os.path.basename(path):
            h.update(open(os.path.join(HERE, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "..", "include", "cy4.h"), "rb").read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    objs = []
    for src in sources():
        flags = ARCH + COMMON + PER_FILE.get(src, [])
        if verbose:
            flags = flags + ["-Xptxas", "-v"]
        obj = os.path.join(OBJ, src[:-3] + ".o")
        stamp = obj + ".sha"
        dg = _digest(os.path.join(HERE, src), flags)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dg:
            continue
        jobs.append((src, obj, stamp, dg, flags))

    def run(job):
        src, obj, stamp, dg, flags = job
        cmd = [NVCC] + flags + ["-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        open(stamp, "w").write(dg)
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s in ex.map(run, jobs):
                print("compiled", s)
    if jobs or force or not os.path.exists(OUT):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        print("linked", OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
