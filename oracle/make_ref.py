"""Recipe for oracle/_ref/: a byte-for-byte copy of the reference's Python sources, made at build() time.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference (maudzung/Complex-YOLOv4-Pytorch) is pure Python, so
there is nothing to compile: "building" it means copying /root/reference/src (read-only, build container
only) into the git-ignored oracle/_ref/src so that it travels to the GPU box with the snapshot, exactly
like an in-tree .so does.  Nothing under oracle/_ref/ is ever committed, imported by the product, or
edited; `bench.py --impl reference`, the GIoU micro-benchmark's CPU leg and tests/test_reference_* then run
the UNMODIFIED reference modules (Darknet, YoloLayer, iou_pred_vs_target_boxes, train.py) on the box's host
cores.  shapely / easydict / matplotlib are not installed in the image: oracle/ref_stubs.py provides
labelled stand-ins (the shapely one is oracle/shapely_standin.py, fp64 convex clipping).

    python -m oracle.make_ref          # from the repo root; no-op when /root/reference is absent
"""
import filecmp
import os
import shutil

SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "src")


def ref_src():
    """Directory holding the reference's src/ tree: the live checkout in the build container, else the copy."""
    if os.path.isdir(SRC):
        return SRC
    if os.path.isdir(DST):
        return DST
    return None


def build(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/_ref: %s not present (GPU box): using the prebuilt copy" % SRC if os.path.isdir(DST)
                  else "oracle/_ref: no reference tree available")
        return DST if os.path.isdir(DST) else None
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, SRC)
        os.makedirs(os.path.join(DST, rel), exist_ok=True)
        for f in files:
            if f.endswith((".pyc", ".pyo")):
                continue
            s, d = os.path.join(root, f), os.path.join(DST, rel, f)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
                n += 1
    if verbose:
        print("oracle/_ref: %d file(s) copied from %s" % (n, SRC))
    return DST


if __name__ == "__main__":
    build()
