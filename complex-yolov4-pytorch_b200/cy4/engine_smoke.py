"""smoke(): one tiny training step of complex_yolov4_tiny through the B200 engine."""
import torch


def run():
    from . import netdefs, synth
    from .darknet import Darknet
    torch.manual_seed(0)
    model = Darknet(netdefs.cfg_path("complex_yolov4_tiny"), use_giou_loss=True).cuda()
    model.train()
    x = synth.make_bev(2, img_size=256).cuda()
    tg = torch.tensor(synth.make_targets(2, per_image=3, seed=1, img_size=256, strides=(16, 32))).cuda()
    loss, out = model(x, tg)
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss).all(), loss
    assert out.shape == (2, 3 * (8 * 8 + 16 * 16), 10), out.shape
    n = sum(1 for p in model.parameters() if p.grad is not None and torch.isfinite(p.grad).all())
    assert n == len(list(model.parameters())), (n, len(list(model.parameters())))
    print("smoke: complex_yolov4_tiny fwd+loss+bwd on cuda:0, loss %.4f, %d parameter gradients finite" % (loss.item(), n))
