#!/usr/bin/env python
"""Overlay launcher: run one of the reference's own scripts (src/train.py, src/evaluate.py, src/test.py) UNMODIFIED with this
repository's drop-in packages first on sys.path.

    python tools/run_reference_script.py train.py [--synthetic-batches N] -- <the script's own arguments>

What it does, and nothing else:
  * puts complex-yolov4-pytorch_b200/ (models/, utils/, data_process/, cy4/) FIRST on sys.path and the reference's src/ LAST:
    `models.model_utils.create_model` -> our Darknet on the sm_100a engine; every module this repository does not replace
    (utils.train_utils, utils.misc, utils.logger, config.*, evaluate, ...) resolves to the reference's own file;
  * registers import stand-ins for packages the image lacks (shapely, easydict, matplotlib: oracle/ref_stubs.py);
  * with --synthetic-batches N (there is no KITTI data on the box): provides `data_process.kitti_dataloader` with a
    `create_train_dataloader(configs)` that yields N batches `(paths, imgs[B,3,608,608] fp32, targets[nT,8])` exactly as
    KittiDataset.collate_fn emits them (reference src/data_process/kitti_dataset.py:216-233);
  * runs the script with runpy.run_path(..., run_name="__main__") from the reference's src/ directory (run_path does not
    put the script's directory in front of sys.path, so the overlay stays in effect).
The reference tree is /root/reference/src in the build container, else the byte-identical copy oracle/_ref/src.
"""
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "complex-yolov4-pytorch_b200")


def synthetic_dataloader_module(n_batches):
    import torch
    from cy4 import synth
    mod = types.ModuleType("data_process.kitti_dataloader")

    class _Loader:
        def __init__(self, configs):
            self.B = int(configs.batch_size)
            self.size = int(getattr(configs, "img_size", 608))
            self.strides = (16, 32) if "tiny" in str(configs.cfgfile) else (8, 16, 32)

        def __len__(self):
            return n_batches

        def __iter__(self):
            for i in range(n_batches):
                imgs = synth.make_bev(self.B, img_size=self.size, seed=1234 + i)
                tg = torch.tensor(synth.make_targets(self.B, per_image=5, seed=4321 + i, img_size=self.size, strides=self.strides))
                yield ["synthetic_%06d" % (i * self.B + j) for j in range(self.B)], imgs, tg

    def create_train_dataloader(configs):
        return _Loader(configs), None

    def _no_data(configs):
        raise RuntimeError("--synthetic-batches only provides the training loader (run train.py with --no-val)")

    mod.create_train_dataloader = create_train_dataloader
    mod.create_val_dataloader = _no_data
    mod.create_test_dataloader = _no_data
    return mod


def main():
    argv = sys.argv[1:]
    if not argv:
        sys.exit(__doc__)
    script = argv.pop(0)
    n_syn = 0
    if argv and argv[0] == "--synthetic-batches":
        n_syn = int(argv[1]); argv = argv[2:]
    if argv and argv[0] == "--":
        argv = argv[1:]
    sys.path.insert(0, ROOT)
    from oracle import make_ref, ref_stubs          # launcher = integration tooling, not the product path
    src = make_ref.ref_src()
    if src is None:
        sys.exit("no reference tree: neither /root/reference/src nor oracle/_ref/src exists (run `python -m oracle.make_ref`)")
    kinds = ref_stubs.install()
    sys.path.insert(0, PKG)
    sys.path.append(src)
    if n_syn:
        import data_process                              # our overlay package (extends __path__ with the reference's directory)
        m = synthetic_dataloader_module(n_syn)
        sys.modules["data_process.kitti_dataloader"] = m
        data_process.kitti_dataloader = m
    path = script if os.path.isabs(script) else os.path.join(src, script)
    os.chdir(src)                                        # the scripts use paths relative to src/ (config/cfg/..., ../dataset)
    sys.argv = [os.path.basename(path)] + argv
    import models.darknet2pytorch as _d
    print("[overlay] reference tree: %s ; stand-ins: %s ; models.darknet2pytorch.Darknet -> %s.%s" %
          (src, kinds, _d.Darknet.__module__, _d.Darknet.__name__), flush=True)
    runpy.run_path(path, run_name="__main__")


if __name__ == "__main__":
    main()
