"""Same-box library comparator for the conv stack (SURVEY section 8d, item iv): the complex_yolov4 graph as plain
PyTorch modules on the B200 (cuDNN / cuBLAS kernels, channels_last, fp16 autocast, Adam), forward + backward + step at
bs=32 with a dummy loss on the three head tensors (the reference's loss head runs shapely on the CPU and is not a GPU
comparison).  NOT part of the product and not used by bench.py; it answers "what does the stock library stack do on
this box" next to our 41 ms step.
    python tools/cudnn_comparator.py [batch] [fp16|tf32]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
import torch.nn as nn
import torch.nn.functional as F
from cy4 import netdefs, synth


class Mish(nn.Module):
    def forward(self, x):
        return F.mish(x)


class Net(nn.Module):
    def __init__(self, blocks):
        super().__init__()
        self.blocks = [b for b in blocks if b["type"] != "net"]
        self.mods = nn.ModuleList()
        chans, c = [], 3
        for i, b in enumerate(self.blocks):
            t = b["type"]
            m = nn.Identity()
            if t == "convolutional":
                k, s = int(b["size"]), int(b["stride"])
                bn = int(b["batch_normalize"])
                layers = [nn.Conv2d(c, int(b["filters"]), k, s, (k - 1) // 2 if int(b["pad"]) else 0, bias=not bn)]
                if bn:
                    layers.append(nn.BatchNorm2d(int(b["filters"])))
                if b["activation"] == "leaky":
                    layers.append(nn.LeakyReLU(0.1, inplace=True))
                elif b["activation"] == "mish":
                    layers.append(Mish())
                m = nn.Sequential(*layers)
                c = int(b["filters"])
            elif t == "maxpool":
                m = nn.MaxPool2d(int(b["size"]), int(b["stride"]), int(b["size"]) // 2)
            elif t == "upsample":
                m = nn.Upsample(scale_factor=2, mode="nearest")
            elif t == "route":
                ls = [int(j) if int(j) > 0 else int(j) + i for j in b["layers"].split(",")]
                c = sum(chans[l] for l in ls)
            elif t == "shortcut":
                c = chans[i - 1]
            self.mods.append(m)
            chans.append(c)

    def forward(self, x):
        outs, heads = [], []
        for i, (b, m) in enumerate(zip(self.blocks, self.mods)):
            t = b["type"]
            if t in ("convolutional", "maxpool", "upsample"):
                x = m(x)
            elif t == "route":
                ls = [int(j) if int(j) > 0 else int(j) + i for j in b["layers"].split(",")]
                x = outs[ls[0]] if len(ls) == 1 else torch.cat([outs[l] for l in ls], 1)
            elif t == "shortcut":
                f = int(b["from"])
                x = outs[f if f > 0 else f + i] + outs[i - 1]
            elif t == "yolo":
                heads.append(x)
            outs.append(x)
        return heads


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    mode = sys.argv[2] if len(sys.argv) > 2 else "fp16"
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)
    net = Net(netdefs.NETS["complex_yolov4"]()).cuda().train().to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    x = synth.make_bev(B).cuda().to(memory_format=torch.channels_last)
    scaler = torch.amp.GradScaler("cuda", enabled=mode == "fp16")

    def step():
        with torch.autocast("cuda", dtype=torch.float16, enabled=mode == "fp16"):
            loss = sum(h.float().pow(2).mean() for h in net(x))
        scaler.scale(loss).backward()
        scaler.step(opt); scaler.update()
        opt.zero_grad(set_to_none=True)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    a.record()
    for _ in range(n):
        step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(json.dumps({"comparator": "PyTorch %s + cuDNN %s, channels_last, %s, dummy loss on the 3 heads" % (torch.__version__, torch.backends.cudnn.version(), mode),
                      "batch": B, "ms_per_step": round(ms, 3), "img_per_s": round(B / ms * 1e3, 1),
                      "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))


if __name__ == "__main__":
    main()
