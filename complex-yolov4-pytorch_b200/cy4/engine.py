"""Step engine: executes the Darknet graph (reference src/models/darknet2pytorch.py:162-230 forward
interpreter + autograd backward) as a static plan of hand-written sm_100a kernels.

Layout in HBM (per plan = per input shape):
  * every activation is NHWC fp16 in a [B*H*W, ld] buffer, ld = channels rounded up to 32 (tensors
    consumed by a multi-input route are produced directly inside the route's buffer);
  * each conv keeps its raw output Y (pre-BatchNorm) and the activated output A; backward
    recomputes BN/activation derivatives from Y instead of storing them;
  * per-channel BN quantities (batch sums, scale/shift, mean/rstd, d gamma, d beta) of ALL layers
    live in a few flat fp32 buffers, zeroed with one memset per step;
  * fp32 master weights stay in the nn.Parameters; K-major fp16 packs for fprop / dgrad are rebuilt
    when the parameter version changes; weight gradients are accumulated in fp32 ([Cout][tap][Cin])
    and unpacked into one flat OIHW gradient buffer whose slices are returned as the .grad tensors;
  * gradient tensors are fp16 under a per-step power-of-two loss scale chosen on the device from
    max |d loss / d head| (target `model.grad_scale_target`), undone in fp32 -- no host sync.

The whole network + loss is ONE autograd node (`_NetFn`): train.py's `loss.backward()` runs the
backward plan, DDP sees ordinary parameter gradients.
"""
import ctypes

import torch

from . import _lib
from . import convops as co
from .yolo import LazyMetrics, make_desc, check_status

ACT = {"linear": 0, "leaky": 1, "mish": 2}
_PROFILED = ("cy4_conv_fwd", "cy4_conv_fwd_stats", "cy4_conv_dgrad", "cy4_conv_dgrad_fused", "cy4_conv_wgrad")


def rup(x, m):
    return (x + m - 1) // m * m


class Storage:
    """An NHWC fp16 tensor [B,H,W,ld] (+ its gradient buffer)."""

    def __init__(self, B, H, W, C, device, dtype=torch.float16, ld=None):
        self.B, self.H, self.W, self.C = B, H, W, C
        self.ld = ld or rup(C, 32)
        self.M = B * H * W
        self.buf = torch.zeros(B, H, W, self.ld, device=device, dtype=dtype)
        self.grad = None
        self.gwritten = []       # channel intervals of .grad written so far in this backward pass

    def ensure_grad(self):
        if self.grad is None:
            self.grad = torch.zeros_like(self.buf)
        return self.grad


class View:
    """Channels [off, off+C) of a storage."""

    def __init__(self, st, off=0, C=None):
        self.st, self.off, self.C = st, off, (st.C if C is None else C)

    @property
    def ptr(self):
        return self.st.buf.data_ptr() + self.off * self.st.buf.element_size()

    @property
    def gptr(self):
        return self.st.ensure_grad().data_ptr() + self.off * 2

    @property
    def ld(self):
        return self.st.ld

    def grad_has(self):
        """True if any part of this view's gradient has been written in the current backward pass."""
        lo, hi = self.off, self.off + self.C
        return any(a < hi and lo < b for a, b in self.st.gwritten)

    def zero_unwritten(self):
        """Zero-fills the channel intervals of this view's gradient that no consumer has written in this backward
        pass (the .grad buffers persist across steps, so an uncovered interval would otherwise hold stale data)."""
        lo, hi = self.off, self.off + self.C
        st = self.st
        pos = lo
        for a, b in sorted((max(a, lo), min(b, hi)) for a, b in st.gwritten if a < hi and lo < b):
            if a > pos:
                st.grad[..., pos:a].zero_()
            pos = max(pos, b)
        if pos < hi:
            st.grad[..., pos:hi].zero_()
            st.gwritten.append((lo, hi))

    def grad_mode(self):
        """Call before writing this view's gradient.  Returns 1 if the kernel must accumulate, 0 if it
        may overwrite; zero-fills the not-yet-written part when the view is only partly covered."""
        lo, hi = self.off, self.off + self.C
        st = self.st
        st.ensure_grad()
        inter = sorted((max(a, lo), min(b, hi)) for a, b in st.gwritten if a < hi and lo < b)
        if not inter:
            st.gwritten.append((lo, hi))
            return 0
        pos = lo
        for a, b in inter:
            if a > pos:
                st.grad[..., pos:a].zero_()
            pos = max(pos, b)
        if pos < hi:
            st.grad[..., pos:hi].zero_()
        st.gwritten.append((lo, hi))
        return 1


class _NetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, x, targets, *params):
        loss = eng._forward(x, targets)
        ctx.eng = eng
        return loss

    @staticmethod
    def backward(ctx, gloss):
        grads = ctx.eng._backward(gloss)
        return (None, None, None) + tuple(grads)


class StepEngine:
    def __init__(self, model):
        self.model = model
        self.plan = None
        self.key = None
        self.L = None

    # ------------------------------------------------------------------------------------------
    def run(self, x, targets):
        _lib.require_device()
        self.L = _lib.lib()
        model = self.model
        dev_in = x.device
        if not x.is_cuda:
            x = x.cuda()
            if targets is not None:
                targets = targets.cuda()
        params = [p for p in model.parameters()]
        if params and params[0].device != x.device:
            raise RuntimeError("Darknet parameters are on %s but the input is on %s" % (params[0].device, x.device))
        training_graph = targets is not None and torch.is_grad_enabled() and any(p.requires_grad for p in params)
        # eval-mode forward without a backward pass: BatchNorm folded into the packed weights, activation (+ shortcut) in the
        # conv epilogue, no raw conv outputs kept
        infer = (not model.training) and not training_graph and bool(getattr(model, "fuse_eval", True))
        key = (tuple(x.shape), x.device, model.training, targets is not None,
               tuple(p.data_ptr() for p in params[:4]), len(params), infer, int(getattr(model, "fuse_bn_backward", 0)))
        with torch.cuda.device(x.device):
            if self.key != key:
                self.plan = None                          # release the old plan's buffers before allocating the new ones
                self.plan = Plan(model, x.shape, x.device, self.L, infer=infer)
                self.key = key
            self.plan.params = params
            if targets is None:
                with torch.no_grad():
                    self._forward(x, None)
                out = self.plan.outputs()
                if getattr(model, "outputs_on_device", False):
                    return out                            # device-resident: feed utils.evaluation_utils.post_processing_v2 directly
                # reference: to_cpu(torch.cat(yolo_outputs, 1)).  Through a pinned staging buffer (a pageable copy of the 29 MB
                # tensor costs ~8 ms); a fresh CPU tensor is returned, the staging buffer is reused by the next call.
                stage = self.plan.pinned_out(out)
                stage.copy_(out, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return stage.clone()
            if training_graph:
                loss = _NetFn.apply(self, x, targets, *params)
            else:
                with torch.no_grad():
                    loss = self._forward(x, targets)
            out = self.plan.outputs()
            if model.sync_outputs or not model.training:
                out_cpu = out.to("cpu")
            else:
                # asynchronous copy into pinned memory: complete at the next stream sync
                # (train.py reads loss.item() every step, which is such a sync)
                # The 29 MB detection tensor goes to pinned memory on a side stream, overlapped with the backward pass; the
                # compute stream waits for that copy at the end of backward (or at the start of the next forward), so the
                # host sees it complete at train.py's per-step loss read like before.
                out_cpu = self.plan.pinned_out(out)
                plan = self.plan
                if plan.d2h_stream is None:
                    plan.d2h_stream = torch.cuda.Stream(device=x.device)
                    plan.d2h_done = torch.cuda.Event()
                ready = torch.cuda.Event()
                ready.record()
                with torch.cuda.stream(plan.d2h_stream):
                    plan.d2h_stream.wait_event(ready)
                    out_cpu.copy_(out, non_blocking=True)
                    plan.d2h_done.record()
                out.record_stream(plan.d2h_stream)
                plan.d2h_pending = True
            return loss, out_cpu

    # ---- optional CUDA-graph replay of the two launch sequences (model.use_cuda_graph) --------------
    # The plan is static per (input shape, number of targets): after two eager steps the forward and
    # the backward launch sequences (~430 + ~760 kernels) are captured once and replayed, which removes
    # the per-launch CPU cost and the inter-kernel gaps.  Inputs / upstream gradient are copied into
    # static buffers; outputs (loss, metrics, detections, gradients) live in static buffers as well.
    def _forward(self, x, targets):
        plan, model = self.plan, self.model
        plan.join_d2h()                    # (outside any graph capture) the previous step's detections copy has left the device buffers
        use_graph = bool(getattr(model, "use_cuda_graph", False)) and targets is not None and model.training
        if not use_graph:
            plan.graph_state = None
            return plan.forward(x, targets, model.use_giou_loss)
        gs = plan.graph_state
        key = (int(targets.shape[0]), bool(model.use_giou_loss), int(getattr(model, "wgrad_overlap", 2)), int(getattr(model, "dy_ring", 4)),
               int(getattr(model, "wgrad_priority", 0)))
        if gs is None or gs["key"] != key:
            gs = plan.graph_state = dict(key=key, eager=0, fwd=None, bwd=None, x=torch.empty_like(x, dtype=torch.float32),
                                         tg=torch.empty(targets.shape, device=x.device, dtype=torch.float32),
                                         g=torch.ones(1, device=x.device, dtype=torch.float32))
        if gs["fwd"] is None and gs["eager"] < 2:
            gs["eager"] += 1
            return plan.forward(x, targets, model.use_giou_loss)
        gs["x"].copy_(x)
        gs["tg"].copy_(targets)
        if gs["fwd"] is None:
            plan.force_pack = True
            n0 = int(self.L.cy4_kernel_launches(0))
            gs["fwd"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gs["fwd"]):
                gs["loss"] = plan.forward(gs["x"], gs["tg"], model.use_giou_loss)
            gs["fwd_launches"] = int(self.L.cy4_kernel_launches(0)) - n0      # kernel nodes of this library in the graph
            plan.force_pack = False
        gs["fwd"].replay()
        self.L.cy4_note_graph_replay(gs["fwd_launches"])
        for y in plan.yolos:
            y["layer"].metrics = LazyMetrics(y["metrics_out"])
        return gs["loss"].clone()

    def _backward(self, gloss):
        plan = self.plan
        plan.join_d2h()                    # the host sees the detections complete at the synchronisation that follows backward
        gs = plan.graph_state
        if gs is None or gs["fwd"] is None:
            return plan.backward(gloss)
        gs["g"].copy_(gloss.reshape(-1)[:1])
        if gs["bwd"] is None:
            n0 = int(self.L.cy4_kernel_launches(0))
            gs["bwd"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gs["bwd"]):
                plan.backward(gs["g"])                   # (its return value -- copies made inside the capture -- is not used)
            gs["bwd_launches"] = int(self.L.cy4_kernel_launches(0)) - n0
        gs["bwd"].replay()
        self.L.cy4_note_graph_replay(gs["bwd_launches"])
        # The gradient base tensors are static graph memory.  With the engine-side exchange (models.model_utils
        # overlap_gradient_exchange: DDP carries a no-op hook) they are averaged over the ranks here, after the replay --
        # not overlapped in this mode, but the ~850 launches of the step cost the host nothing, which is what bounds
        # multi-rank steps (bench.py host_enqueue_ms_per_step).
        if getattr(self.model, "engine_allreduce", False):
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                for b in plan._grad_bases:
                    dist.all_reduce(b, op=dist.ReduceOp.AVG)
        return plan.materialise()


class Plan:
    """Buffers + op lists for one (batch, height, width) on one device."""

    def __init__(self, model, xshape, device, L, infer=False):
        self.L = L
        self.model = model
        self.device = device
        self.infer = infer
        self.B, cin, self.H, self.W = xshape
        assert cin == 3, "the BEV input has 3 channels"
        self.scale_target = float(model.grad_scale_target)
        self.fwd_ops, self.bwd_ops = [], []
        self.params = None
        self._pinned = None
        self.prof = None             # list -> CUDA-event timing of every conv launch (bench.py roofline pass)
        self.graph_state = None
        self.force_pack = False
        self.d2h_stream, self.d2h_done, self.d2h_pending = None, None, False
        self._build()

    # ---- helpers -----------------------------------------------------------------------------
    def _call(self, fn, *args):
        prof = self.prof
        if prof is not None and fn.__name__ in _PROFILED:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            d = args[0]._obj
            prof.append((fn.__name__, e0, e1, 2.0 * d.B * d.Ho * d.Wo * d.Cout * d.ksize * d.ksize * d.Cin,
                         (d.Cin, d.Cout, d.ksize, d.stride, d.Ho)))
        else:
            rc = fn(*args)
        if rc < 0:
            _lib.check(rc, fn.__name__)

    def _analyse(self):
        """Shape / consumer pre-pass over the cfg blocks.  Decides
          * shortcut fusion: a conv+BN block whose only consumer is the following [shortcut] adds the
            residual in its BN/activation pass and writes the shortcut's output directly;
          * concat placement: a tensor consumed by a multi-layer [route] is produced directly inside
            that route's buffer (at most one concat per tensor; further concats copy)."""
        blocks = self.model.blocks
        info = {}                       # ind -> dict(type, C, H, W, origin, srcs)
        consumers = {}
        H, W, C = self.H, self.W, 3
        ind = -2
        prev = None
        for block in blocks:
            ind += 1
            t = block["type"]
            if t == "net":
                continue
            srcs = []
            origin = ind
            if t == "convolutional":
                k, s = int(block["size"]), int(block["stride"])
                pad = (k - 1) // 2 if int(block["pad"]) else 0
                H, W, C = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1, int(block["filters"])
                srcs = [prev] if prev is not None else []
            elif t == "route":
                ls = [int(i) if int(i) > 0 else int(i) + ind for i in block["layers"].split(",")]
                srcs = ls
                H, W = info[ls[0]]["H"], info[ls[0]]["W"]
                if len(ls) == 1:
                    g = int(block.get("groups", 1))
                    C = info[ls[0]]["C"] // g
                    origin = info[ls[0]]["origin"] if g == 1 else None      # a channel slice is not placeable
                else:
                    C = sum(info[l]["C"] for l in ls)
            elif t == "shortcut":
                f = int(block["from"])
                srcs = [f if f > 0 else f + ind, ind - 1]
            elif t == "maxpool":
                k, s = int(block["size"]), int(block["stride"])
                pad = k // 2 if (s == 1 and k % 2) else 0
                H, W = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
                srcs = [prev]
            elif t == "upsample":
                H, W = 2 * H, 2 * W
                srcs = [prev]
            elif t == "yolo":
                srcs = [prev]
            info[ind] = dict(type=t, C=C, H=H, W=W, origin=origin, srcs=srcs, block=block)
            for sidx in srcs:
                consumers.setdefault(sidx, []).append(ind)
            prev = ind
        fused = {}                      # shortcut ind -> conv ind (ind - 1)
        for i, inf in info.items():
            if inf["type"] == "shortcut":
                c = i - 1
                if info[c]["type"] == "convolutional" and int(info[c]["block"]["batch_normalize"]) and consumers.get(c) == [i]:
                    fused[i] = c
        placement = {}                  # producing layer -> (route ind, channel offset)
        for i, inf in info.items():
            if inf["type"] == "route" and len(inf["srcs"]) > 1:
                off = 0
                for sidx in inf["srcs"]:
                    o = info[sidx]["origin"]
                    if o is not None and info[o]["type"] == "shortcut" and o in fused:
                        pass                                    # the fused conv writes the shortcut output
                    ok = o is not None and o not in placement and info[o]["type"] in ("convolutional", "shortcut", "maxpool", "upsample")
                    if ok and info[o]["type"] == "convolutional" and not int(info[o]["block"]["batch_normalize"]):
                        ok = False                              # head convs keep their fp32 buffer
                    if ok and o in fused.values():
                        ok = False                              # its output is never materialised
                    if ok:
                        placement[o] = (i, off)
                    off += info[sidx]["C"]
        return info, consumers, fused, placement

    def _build(self):
        model, L, dev, B = self.model, self.L, self.device, self.B
        info, consumers, fused, placement = self._analyse()
        fused_convs = {c: s for s, c in fused.items()}          # conv ind -> shortcut ind
        outs = {}                       # layer index -> View
        self.convs, self.yolos = [], []
        self.routes, self.shorts, self.pools, self.ups = [], [], [], []
        cat_storage = {}                # route ind -> Storage

        def cat_of(route_ind):
            if route_ind not in cat_storage:
                inf = info[route_ind]
                cat_storage[route_ind] = Storage(B, inf["H"], inf["W"], inf["C"], dev)
            return cat_storage[route_ind]

        def out_view(i):
            """Where layer i's output lives: a slice of a concat buffer or its own storage."""
            inf = info[i]
            if i in placement:
                r, off = placement[i]
                return View(cat_of(r), off, inf["C"])
            return View(Storage(B, inf["H"], inf["W"], inf["C"], dev))

        cur = None
        totalC = 0
        for ind in sorted(info):
            inf = info[ind]
            block, t = inf["block"], inf["type"]
            if t == "convolutional":
                seq = model.models[ind]
                conv = seq[0]
                bn = seq[1] if int(block["batch_normalize"]) else None
                k, stride = int(block["size"]), int(block["stride"])
                pad = (k - 1) // 2 if int(block["pad"]) else 0
                Cout = int(block["filters"])
                if block["activation"] not in ACT:
                    raise NotImplementedError("activation %s" % block["activation"])
                src = outs[inf["srcs"][0]] if inf["srcs"] else None
                Hi, Wi = (info[inf["srcs"][0]]["H"], info[inf["srcs"][0]]["W"]) if inf["srcs"] else (self.H, self.W)
                rec = dict(ind=ind, conv=conv, bn=bn, k=k, stride=stride, pad=pad, Cout=Cout, Hi=Hi, Wi=Wi, Ho=inf["H"], Wo=inf["W"],
                           act=ACT[block["activation"]], src=src, stem=src is None, Cin=conv.in_channels, res=None)
                if rec["stem"]:
                    assert conv.in_channels * k * k <= 32, "stem conv must have C*k*k <= 32"
                    rec["cols"] = Storage(B, inf["H"], inf["W"], 32, dev, ld=32)
                if bn is not None:
                    rec["Y"] = None if self.infer else Storage(B, inf["H"], inf["W"], Cout, dev)
                    rec["M"] = B * inf["H"] * inf["W"]
                    if ind in fused_convs:                      # conv + BN + act + residual -> shortcut output
                        sc = fused_convs[ind]
                        rec["res"] = outs[info[sc]["srcs"][0]]
                        rec["A"] = out_view(sc)
                    else:
                        rec["A"] = out_view(ind)
                    rec["coff"] = totalC
                    totalC += rup(Cout, 8)
                    cur = rec["A"]
                else:
                    rec["P"] = Storage(B, inf["H"], inf["W"], rup(Cout, 32), dev, dtype=torch.float32, ld=rup(Cout, 32))
                    rec["dP"] = torch.zeros_like(rec["P"].buf)
                    cur = View(rec["P"])
                self.convs.append(rec)
                self.fwd_ops.append(("conv", rec))
            elif t == "route":
                ls = inf["srcs"]
                if len(ls) == 1:
                    src = outs[ls[0]]
                    g = int(block.get("groups", 1))
                    cur = src if g == 1 else View(src.st, src.off + (src.C // g) * int(block["group_id"]), src.C // g)
                else:
                    cat = cat_of(ind)
                    off, copies = 0, []
                    for sidx in ls:
                        s = outs[sidx]
                        if not (s.st is cat and s.off == off):   # not produced in place: copy
                            self.fwd_ops.append((L.cy4_add_copy, (s.ptr, s.ld, None, 0, cat.buf.data_ptr() + off * 2, cat.ld, cat.M, s.C)))
                            copies.append((s, off))
                        off += s.C
                    cur = View(cat)
                    if copies:
                        self.routes.append((ind, cat, copies))
            elif t == "shortcut":
                assert block["activation"] == "linear", "shortcut activation %s" % block["activation"]
                a = outs[inf["srcs"][0]]
                if ind in fused:
                    cur = outs[ind - 1]                          # already the sum (written by the fused conv)
                    self.shorts.append((ind, cur, a, None))
                else:
                    b = outs[ind - 1]
                    cur = out_view(ind)
                    self.fwd_ops.append((L.cy4_add_copy, (a.ptr, a.ld, b.ptr, b.ld, cur.ptr, cur.ld, cur.st.M, a.C)))
                    self.shorts.append((ind, cur, a, b))
            elif t == "maxpool":
                k, stride = int(block["size"]), int(block["stride"])
                if stride == 1 and k % 2:
                    pad = k // 2
                elif stride == k:
                    pad = 0
                else:
                    raise NotImplementedError("MaxPoolDark (size %d stride %d) is outside the complex-yolov4 cfgs" % (k, stride))
                src = outs[inf["srcs"][0]]
                Hi, Wi = info[inf["srcs"][0]]["H"], info[inf["srcs"][0]]["W"]
                cur = out_view(ind)
                Hp, Wp = (Hi + 2 * pad - k) // stride + 1, (Wi + 2 * pad - k) // stride + 1
                if self.infer or k * k > 255:
                    amax = None
                    self.fwd_ops.append((L.cy4_maxpool_fwd, (src.ptr, src.ld, cur.ptr, cur.ld, B, Hi, Wi, src.C, k, stride, pad)))
                else:       # training plan: keep the argmax (one byte per output element) for the backward routing
                    amax = torch.empty(B * Hp * Wp * src.C, device=self.device, dtype=torch.uint8)
                    if getattr(self, "pool_ws", None) is None or self.pool_ws.numel() < 3 * B * Hi * Wi * src.C:
                        self.pool_ws = torch.empty(3 * B * Hi * Wi * src.C, device=self.device, dtype=torch.uint8)   # separable-pass workspace
                    self.fwd_ops.append((L.cy4_maxpool_fwd_idx, (src.ptr, src.ld, cur.ptr, cur.ld, amax.data_ptr(), self.pool_ws.data_ptr(),
                                                                 B, Hi, Wi, src.C, k, stride, pad)))
                self.pools.append((ind, cur, src, k, stride, pad, Hi, Wi, amax))
            elif t == "upsample":
                assert int(block["stride"]) == 2
                src = outs[inf["srcs"][0]]
                Hi, Wi = info[inf["srcs"][0]]["H"], info[inf["srcs"][0]]["W"]
                cur = out_view(ind)
                self.fwd_ops.append((L.cy4_upsample2x_fwd, (src.ptr, src.ld, cur.ptr, cur.ld, B, Hi, Wi, src.C)))
                self.ups.append((ind, cur, src, Hi, Wi))
            elif t == "yolo":
                layer = model.models[ind]
                head = self.convs[-1]
                assert "P" in head and head["ind"] == ind - 1, "a [yolo] block must follow a linear conv without batch norm"
                G = head["Ho"]
                nA, nC = layer.num_anchors, layer.num_classes
                stride = self.H / G
                anchors4 = torch.tensor([(aw / stride, ah / stride, im, re) for aw, ah, im, re in layer.anchors], device=dev,
                                        dtype=torch.float32)
                yrec = dict(ind=ind, layer=layer, head=head, G=G, nA=nA, nC=nC, anchors4=anchors4,
                            out=torch.empty(B, nA * G * G, 7 + nC, device=dev, dtype=torch.float32),
                            loss=torch.zeros(1, device=dev), metrics=torch.zeros(18, device=dev),
                            status=torch.zeros(1, device=dev, dtype=torch.int32), ws=None, nT=-1)
                self.yolos.append(yrec)
                self.fwd_ops.append(("yolo", yrec))
            else:
                raise NotImplementedError("block type %s" % t)
            outs[ind] = cur
            if t == "convolutional" and ind in fused_convs:
                outs[ind] = cur                                   # (= the shortcut's tensor; nobody else reads it)

        # flat per-channel buffers
        tc = max(totalC, 8)
        f32 = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        self.stats = f32(2, tc)            # sum, sum of squares (zeroed every forward)
        self.bnq = f32(4, tc)              # scale, shift, mean, rstd
        self.dbn = f32(2, tc)              # d beta, d gamma (under the loss scale), zeroed every backward
        # BatchNorm statistics are summed about a per-channel shift c = the previous step's batch mean (cy4_conv_fwd_stats): no
        # cancellation in E[y^2] - E[y]^2 however large |mean| / sigma grows during training.  The apply pass publishes the new
        # mean into shift_next; it becomes shift_cur at the start of the next forward (never while a kernel reads it).
        self.shift_cur, self.shift_next = f32(tc), f32(tc)
        self.shifted_stats = bool(getattr(model, "bn_shifted_stats", True)) and not self.infer
        # weight packs / gradient accumulators ([Cout_pad][taps][Cin] fp32, one flat buffer, one memset)
        wtot, atot = 0, 0
        for rec in self.convs:
            conv = rec["conv"]
            Cout, Cin, k = rec["Cout"], rec["Cin"], rec["k"]
            rec["woff"] = wtot
            wtot += conv.weight.numel()
            rec["aoff"] = atot
            rec["ashape"] = (rup(Cout, 32), 1, 32) if rec["stem"] else (rup(Cout, 32), k * k, Cin)
            atot += rec["ashape"][0] * rec["ashape"][1] * rec["ashape"][2]
            if rec["stem"]:
                rec["wf"] = torch.zeros(rup(Cout, 32), 32, device=dev, dtype=torch.float16)
            else:
                rec["wf"] = torch.empty(rup(Cout, 32), k * k * Cin, device=dev, dtype=torch.float16)
                if not self.infer:
                    rec["wd"] = torch.empty(rup(Cin, 32), k * k * rup(Cout, 32), device=dev, dtype=torch.float16)
            rec["wver"] = -1
        self.acc_flat = f32(1 if self.infer else max(atot, 1))          # (an inference plan has no backward buffers)
        for rec in self.convs:
            n = rec["ashape"][0] * rec["ashape"][1] * rec["ashape"][2]
            rec["acc"] = None if self.infer else self.acc_flat[rec["aoff"]:rec["aoff"] + n].view(rec["ashape"])
        self.gw_numel = max(wtot, 1)
        self.gw_flat = None
        self._unpack_dev = None
        self.dy_ring = None                # dY buffers (BN-backward output / head gradients in fp16), rotated over the layers
        self._wg = None                    # side stream + events of the overlapped weight-gradient launches
        self.pool_scratch = None
        max_dy = 0
        for rec in self.convs:
            max_dy = max(max_dy, rec["M"] * rup(rec["Cout"], 64) if rec["bn"] is not None else rec["P"].M * 64)
        self._max_dy = max_dy
        self._plan_bwd_fusion()
        self._storages = set()
        for rec in self.convs:
            if "A" in rec:
                self._storages.add(rec["A"].st)
        for lst in (self.shorts, self.pools, self.ups):
            for r in lst:
                self._storages.add(r[1].st)
        for st in cat_storage.values():
            self._storages.add(st)

    def _plan_bwd_fusion(self, mode=None):
        """Static pass over the backward order: for every BatchNorm conv P whose activated output A receives its LAST
        gradient contribution from the input-gradient kernel of a conv C reading exactly that view, C's epilogue takes over
        the first pass of P's BN/activation backward (cy4_conv_dgrad_fused: dz = dA_total * act'(z) stored instead of dA,
        sum dz / sum dz*Y accumulated) and P skips cy4_bn_act_bwd_reduce.  Tensors whose gradient is completed by a route
        copy, a shortcut, a pooling / upsampling backward or by the consumer of a concat buffer keep the separate pass."""
        for rec in self.convs:
            rec["fuse_bwd"] = None          # consumer side: the producer rec whose reduce pass this conv's dgrad performs
            rec["reduce_fused"] = False     # producer side
        if mode is None:
            mode = int(getattr(self.model, "fuse_bn_backward", 0))
        if self.infer or not mode:
            return
        writers = {}                        # id(storage) -> [(lo, hi, kind, rec)] in backward order
        events = {}
        for rec in self.convs:
            events[rec["ind"]] = ("conv", rec)
        for r in self.routes:
            events[r[0]] = ("route", r)
        for r in self.shorts:
            events[r[0]] = ("short", r)
        for r in self.pools:
            events[r[0]] = ("pool", r)
        for r in self.ups:
            events[r[0]] = ("up", r)

        def note(view, kind, rec=None):
            writers.setdefault(id(view.st), []).append((view.off, view.off + view.C, kind, rec))

        for ind in sorted(events, reverse=True):
            kind, r = events[ind]
            if kind == "conv":
                if not r["stem"]:
                    note(r["src"], "conv", r)
            elif kind == "route":
                for s_, _off in r[2]:
                    note(s_, "copy")
            elif kind == "short":
                note(r[2], "copy")
                if r[3] is not None:
                    note(r[3], "copy")
            elif kind == "pool":
                note(r[2], "copy")
            elif kind == "up":
                note(r[2], "copy")
        for P in self.convs:
            if P["bn"] is None:
                continue
            A = P["A"]
            lo, hi = A.off, A.off + A.C
            ws = [w for w in writers.get(id(A.st), []) if w[0] < hi and lo < w[1]]
            if not ws or any((w[0], w[1]) != (lo, hi) for w in ws):
                continue                    # no gradient, or partially overlapping writers (concat consumers, channel slices)
            last = ws[-1]
            if last[2] != "conv" or last[3]["fuse_bwd"] is not None:
                continue
            if P["res"] is not None:
                continue                    # A is a fused shortcut output: its RAW gradient is still needed by the residual branch
                                            # (the [shortcut] event copies it after the last writer) -- dz must not replace it early
            # Mish' costs two MUFU operations per element and the eight epilogue warps are all the SM has for it while the
            # tile's MMAs run: 16 * N cycles per 128 x N tile.  The tensor work of that tile is k^2 * Cout / 64 k-blocks of
            # 2 * N cycles, so only input-gradient GEMMs with a long reduction (3 x 3, or very wide 1 x 1) hide the extra pass;
            # on the short-K 1 x 1 layers the fused epilogue was measured slower than the separate bandwidth-bound pass
            # (profiles/r2_bn_backward_fusion.md).  LeakyReLU / linear producers have no MUFU work and always fuse.
            cons = last[3]
            if mode == 1 and P["act"] == ACT["mish"] and cons["k"] * cons["k"] * cons["Cout"] < 1024:
                continue
            last[3]["fuse_bwd"] = P
            P["reduce_fused"] = True

    # ---- forward -----------------------------------------------------------------------------
    def _pack_weights(self, st):
        """fp32 OIHW parameters -> K-major fp16 packs, one batched launch, only when a parameter changed."""
        from ._sigs_engine import PackItem
        L = self.L
        ptrs = tuple(rec["conv"].weight.data_ptr() for rec in self.convs)
        sig = tuple(rec["conv"].weight._version for rec in self.convs) + ptrs
        if self.infer:      # the folded packs also depend on the BatchNorm parameters and running statistics
            sig += tuple(t._version for rec in self.convs if rec["bn"] is not None
                         for t in (rec["bn"].weight, rec["bn"].bias, rec["bn"].running_mean, rec["bn"].running_var))
        if sig == getattr(self, "_pack_sig", None) and not self.force_pack:
            return
        if self.infer:      # scale = gamma * rsqrt(running_var + eps), shift = beta - running_mean * scale  (per layer, once)
            for rec in self.convs:
                bn = rec["bn"]
                if bn is None:
                    continue
                q = [self.bnq[i, rec["coff"]:].data_ptr() for i in range(4)]
                self._call(L.cy4_bn_finalize, None, None, 1.0, bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                           bn.running_var.data_ptr(), None, float(bn.momentum), float(bn.eps), 0, rec["Cout"], q[0], q[1], q[2], q[3], st)
        if ptrs != getattr(self, "_pack_ptrs", None):
            items, tiles = [], 0
            for rec in self.convs:
                if rec["stem"]:
                    continue
                it = PackItem()
                it.tile_begin = tiles
                tiles += (rup(rec["Cout"], 32) // 32) * (rup(rec["Cin"], 32) // 32)
                w = rec["conv"].weight
                it.w_oihw, it.w_fprop, it.w_dgrad = w.data_ptr(), rec["wf"].data_ptr(), (rec["wd"].data_ptr() if "wd" in rec else None)
                it.Cout, it.Cin, it.ksize = rec["Cout"], rec["Cin"], rec["k"]
                it.cout_pad, it.cin_pad = rup(rec["Cout"], 32), rup(rec["Cin"], 32)
                if self.infer:
                    it.fold_scale = self.bnq[0, rec["coff"]:].data_ptr() if rec["bn"] is not None else None
                items.append(it)
            arr = (PackItem * len(items))(*items)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
            self._pack_table = host.to(self.device)
            self._pack_n = len(items)
            self._pack_ptrs = ptrs
        self._pack_sig = sig
        for rec in self.convs:
            if rec["stem"]:       # (r, s, c) column order of cy4_stem_im2col, padded to 32
                w = rec["conv"].weight.detach().permute(0, 2, 3, 1).reshape(rec["Cout"], -1)
                if self.infer and rec["bn"] is not None:
                    w = w * self.bnq[0, rec["coff"]:rec["coff"] + rec["Cout"], None]
                rec["wf"][:rec["Cout"], :rec["Cin"] * rec["k"] ** 2] = w.to(torch.float16)
        if self._pack_n:
            self._call(L.cy4_pack_weights_batched, self._pack_table.data_ptr(), self._pack_n, st)

    def forward(self, x, targets, use_giou):
        L = self.L
        st = _lib.stream()
        model = self.model
        training = model.training
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        self._x = x
        self._pack_weights(st)
        self.stats.zero_()
        if training and self.shifted_stats:
            self.shift_cur.copy_(self.shift_next)
        total = None
        if targets is not None:
            tg = targets.detach().to(self.device, torch.float32).contiguous()
            self._tg = tg
        B = self.B
        for op in self.fwd_ops:
            if op[0] == "conv":
                self._conv_forward(op[1], x, training, st)
            elif op[0] == "yolo":
                y = op[1]
                head = y["head"]
                G, nA, nC = y["G"], y["nA"], y["nC"]
                ldp = head["P"].ld
                d = make_desc(B, G, nA, nC, (G * G * ldp, 1, G * ldp, ldp), self.H, y["layer"].ignore_thresh, use_giou)
                y["desc"] = d
                layer = y["layer"]
                layer.img_size, layer.use_giou_loss, layer.grid_size = self.H, use_giou, G
                if targets is None:
                    self._call(L.cy4_yolo_decode, ctypes.byref(d), head["P"].buf.data_ptr(), y["anchors4"].data_ptr(),
                               y["out"].data_ptr(), st)
                else:
                    nT = tg.shape[0]
                    if y["nT"] != nT or y["ws"] is None:
                        y["ws"] = torch.empty(L.cy4_yolo_workspace_bytes(ctypes.byref(d), nT), device=self.device, dtype=torch.uint8)
                        y["nT"] = nT
                    self._call(L.cy4_yolo_loss_fwd, ctypes.byref(d), head["P"].buf.data_ptr(), y["anchors4"].data_ptr(),
                               tg.data_ptr() if nT else None, nT, y["out"].data_ptr(), y["loss"].data_ptr(),
                               y["metrics"].data_ptr(), y["status"].data_ptr(), y["ws"].data_ptr(), st)
                    y["metrics_out"] = y["metrics"].clone()
                    layer.metrics = LazyMetrics(y["metrics_out"])
                    layer._status = y["status"]
                    if layer.check_targets:
                        check_status(y["status"], "YoloLayer")
                    total = y["loss"].clone() if total is None else total + y["loss"]
            else:
                fn, args = op
                self._call(fn, *args, st)
        if targets is None:
            return None
        # reference: `loss = 0.; loss += layer_loss` -> shape [1] with GIoU, 0-dim without
        return total if use_giou else total.reshape(())

    def _conv_forward(self, rec, x, training, st):
        L = self.L
        B = self.B
        k, stride, pad, Cout, Cin = rec["k"], rec["stride"], rec["pad"], rec["Cout"], rec["Cin"]
        if rec["stem"]:
            self._call(L.cy4_stem_im2col, x.data_ptr(), B, Cin, self.H, self.W, k, stride, pad, rec["cols"].buf.data_ptr(), st)
            d = co.conv_desc(B, rec["Ho"], rec["Wo"], 32, Cout, 1, 1, 0, 32, 0, 0)
            src_ptr = rec["cols"].buf.data_ptr()
            amat = co.CONV_A_MATRIX
        else:
            src = rec["src"]
            d = co.conv_desc(B, rec["Hi"], rec["Wi"], Cin, Cout, k, stride, pad, src.ld, 0, 0)
            src_ptr = src.ptr
            amat = 0
        if rec["bn"] is not None and self.infer:
            A, res = rec["A"], rec["res"]
            d.ldy = A.ld
            d.flags = amat
            self._call(L.cy4_conv_fwd_fused, ctypes.byref(d), src_ptr, rec["wf"].data_ptr(), A.ptr, self.bnq[1, rec["coff"]:].data_ptr(),
                       rec["act"], res.ptr if res is not None else None, res.ld if res is not None else 0, st)
        elif rec["bn"] is not None:
            bn = rec["bn"]
            Y, A = rec["Y"], rec["A"]
            res = rec["res"]
            c0 = rec["coff"]
            d.ldy = Y.ld
            d.flags = amat | (co.CONV_STATS if training else 0)
            s1 = self.stats[0, c0:].data_ptr(); s2 = self.stats[1, c0:].data_ptr()
            shifted = training and self.shifted_stats
            if shifted:
                self._call(L.cy4_conv_fwd_stats, ctypes.byref(d), src_ptr, rec["wf"].data_ptr(), Y.buf.data_ptr(), s1, s2,
                           self.shift_cur[c0:].data_ptr(), st)
            else:
                self._call(L.cy4_conv_fwd, ctypes.byref(d), src_ptr, rec["wf"].data_ptr(), Y.buf.data_ptr(), None, s1, s2, st)
            q = [self.bnq[i, c0:].data_ptr() for i in range(4)]
            if training:      # batch statistics -> scale / shift inside the apply pass (one launch)
                self._call(L.cy4_bn_train_act_fwd, Y.buf.data_ptr(), Y.ld, s1, s2, float(Y.M), bn.weight.data_ptr(), bn.bias.data_ptr(),
                           bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), float(bn.momentum),
                           float(bn.eps), q[0], q[1], q[2], q[3], rec["act"], res.ptr if res is not None else None,
                           res.ld if res is not None else 0, A.ptr, A.ld, Y.M, Cout,
                           self.shift_cur[c0:].data_ptr() if shifted else None, self.shift_next[c0:].data_ptr() if shifted else None, st)
            else:
                self._call(L.cy4_bn_finalize, s1, s2, float(Y.M), bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                           bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), float(bn.momentum), float(bn.eps),
                           0, Cout, q[0], q[1], q[2], q[3], st)
                self._call(L.cy4_bn_act_fwd, Y.buf.data_ptr(), Y.ld, q[0], q[1], rec["act"], res.ptr if res is not None else None,
                           res.ld if res is not None else 0, A.ptr, A.ld, Y.M, Cout, st)
        else:
            P = rec["P"]
            d.ldy = P.ld
            d.flags = amat | co.CONV_OUT_F32
            bias = rec["conv"].bias
            if bias is not None:
                if "bias32" not in rec:
                    rec["bias32"] = torch.zeros(rup(Cout, 32), device=self.device, dtype=torch.float32)
                rec["bias32"][:Cout] = bias.detach()
            self._call(L.cy4_conv_fwd, ctypes.byref(d), src_ptr, rec["wf"].data_ptr(), P.buf.data_ptr(),
                       rec["bias32"].data_ptr() if bias is not None else None, None, None, st)

    def join_d2h(self):
        """Make the compute stream wait for the asynchronous detections copy of this plan (if one is in flight)."""
        if self.d2h_pending:
            torch.cuda.current_stream().wait_event(self.d2h_done)
            self.d2h_pending = False

    def outputs(self):
        return torch.cat([y["out"] for y in self.yolos], 1)

    def pinned_out(self, out):
        if self._pinned is None or self._pinned.shape != out.shape:
            self._pinned = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        return self._pinned

    # ---- backward ----------------------------------------------------------------------------
    def backward(self, gloss):
        L = self.L
        st = _lib.stream()
        B = self.B
        model = self.model
        training = model.training
        dev = self.device
        # Weight gradients on a second stream (model.wgrad_overlap): the tensor-bound cy4_conv_wgrad of layer L runs while the
        # compute stream is in the HBM-bound BatchNorm / activation backward passes of layer L-1 (they need only dgrad_L).
        #   0: everything on one stream;  1: wgrad_L may start as soon as dY_L exists (next to dgrad_L);
        #   2: wgrad_L starts after dgrad_L (next to the BN passes of L-1).
        # dY lives in a ring of model.dy_ring buffers; a slot is rewritten only after the wgrad that read it has finished.
        mode = int(getattr(model, "wgrad_overlap", 2))
        if self.prof is not None:
            mode = 0                                     # per-launch event timing assumes one stream
        ring = max(2, int(getattr(model, "dy_ring", 4))) if mode else 1
        if self.dy_ring is None or len(self.dy_ring) != ring:
            self.dy_ring = None
            self.dy_ring = [torch.zeros(self._max_dy, device=dev, dtype=torch.float16) for _ in range(ring)]
        if getattr(self, "gscale", None) is None:
            self.gscale = torch.zeros(3, device=dev, dtype=torch.float32)      # [S, 1/S, amax]
        prio = int(getattr(model, "wgrad_priority", 0))
        if mode and (self._wg is None or self._wg["prio"] != prio):
            # priority -1: when a weight-gradient kernel and a BN pass become ready together, the wgrad CTAs (one per SM, ~194 KB of
            # shared memory) are placed first and the pass's blocks fill in beside them; with equal priorities whichever kernel was
            # launched first takes every SM (3 resident BN blocks leave no registers for a wgrad CTA) and the two serialise.
            self._wg = dict(prio=prio, stream=torch.cuda.Stream(device=dev, priority=prio), fork=[torch.cuda.Event() for _ in self.convs],
                            slot=[torch.cuda.Event() for _ in range(64)], done=torch.cuda.Event())
        for i, rec in enumerate(self.convs):
            rec["cidx"] = i
        self._wg_mode = mode
        self._dy_next = 0
        self._slot_busy = [False] * ring
        self._wg_open = False                            # side-stream work not yet joined by the compute stream
        g = gloss.reshape(-1)[:1].to(torch.float32).contiguous()
        self.dbn.zero_()
        for s_ in self._storages:
            s_.gwritten = []
        # head gradients first (fp32, unscaled), then the loss scale of everything below them
        self.gscale[2:].zero_()
        for y in self.yolos:
            head = y["head"]
            P = head["P"]
            ldp, G = P.ld, y["G"]
            self._call(L.cy4_yolo_loss_bwd, ctypes.byref(y["desc"]), P.buf.data_ptr(), y["anchors4"].data_ptr(),
                       self._tg.data_ptr() if y["nT"] else None, y["nT"], g.data_ptr(), y["ws"].data_ptr(),
                       head["dP"].data_ptr(), G * G * ldp, 1, G * ldp, ldp, st)
            self._call(L.cy4_absmax_f32, head["dP"].data_ptr(), head["dP"].numel(), self.gscale[2:].data_ptr(), st)
            head["has_grad"] = True
        self._call(L.cy4_make_scale, self.gscale[2:].data_ptr(), self.scale_target, self.gscale.data_ptr(), st)
        # The unpack launch writes every conv gradient into the persistent flat buffer (its item table holds the
        # addresses); what autograd receives are slices of a FRESH copy made after the unpack: AccumulateGrad keeps
        # the tensors it is handed as .grad (it looks at the view's own use count, not at the base buffer's), so
        # returning views of a persistent buffer would alias .grad with the next backward's output and break
        # gradient accumulation (train.py accumulates over `subdivisions` backward calls per optimizer step).
        if getattr(self, "gw_flat", None) is None:
            self.gw_flat = torch.zeros(self.gw_numel, device=dev, dtype=torch.float32)
        gw_flat = self.gw_flat
        grads = {}
        self._wslices = []

        events = {}
        for rec in self.convs:
            events[rec["ind"]] = ("conv", rec)
        for r in self.routes:
            events[r[0]] = ("route", r)
        for r in self.shorts:
            events[r[0]] = ("short", r)
        for r in self.pools:
            events[r[0]] = ("pool", r)
        for r in self.ups:
            events[r[0]] = ("up", r)

        def accumulate_into(view, src_ptr, src_ld, M):
            """view.grad (+)= src"""
            if view.grad_mode():
                self._call(L.cy4_add_copy, view.gptr, view.ld, src_ptr, src_ld, view.gptr, view.ld, M, view.C, st)
            else:
                self._call(L.cy4_add_copy, src_ptr, src_ld, None, 0, view.gptr, view.ld, M, view.C, st)

        # gradient exchange done by the engine itself (model.engine_allreduce, set by models.model_utils.make_data_parallel
        # together with a no-op DDP communication hook): grouped, asynchronous, overlapped with the rest of backward
        ar_groups, works = None, []
        if getattr(model, "engine_allreduce", False) and (self.graph_state is None or self.graph_state["fwd"] is None):   # (eager steps)
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                ar_groups = self._allreduce_groups()

        for ind in sorted(events, reverse=True):
            kind, r = events[ind]
            if kind == "conv":
                self._conv_backward(r, training, st, gw_flat, grads)
                if ar_groups is not None and ind in ar_groups:
                    first, n, lo, hi = ar_groups[ind]
                    self._join_wgrad()
                    self._unpack_range(st, first, n)
                    works.append(dist.all_reduce(gw_flat[lo:hi], op=dist.ReduceOp.AVG, async_op=True))
            elif kind == "route":
                _, cat, copies = r
                for s_, off in copies:                   # sources that were copied (not produced in place)
                    cv = View(cat, off, s_.C)
                    if cv.grad_has():
                        accumulate_into(s_, cv.gptr, cat.ld, cat.M)
            elif kind == "short":
                _, out, a, b = r
                if not out.grad_has():
                    continue
                accumulate_into(a, out.gptr, out.ld, out.st.M)
                if b is not None:                        # unfused: the conv branch receives the same gradient
                    accumulate_into(b, out.gptr, out.ld, out.st.M)
            elif kind == "pool":
                _, out, src, k, stride, pad, Hi, Wi, amax = r
                if not out.grad_has():
                    continue
                need = B * Hi * Wi * src.C
                if self.pool_scratch is None or self.pool_scratch.numel() < need:
                    self.pool_scratch = torch.zeros(need, device=dev, dtype=torch.float32)
                self.pool_scratch[:need].zero_()
                if amax is not None:
                    self._call(L.cy4_maxpool_bwd_idx, amax.data_ptr(), out.gptr, out.ld, self.pool_scratch.data_ptr(), B, Hi, Wi, src.C,
                               k, stride, pad, st)
                else:
                    self._call(L.cy4_maxpool_bwd, src.ptr, src.ld, out.gptr, out.ld, self.pool_scratch.data_ptr(), B, Hi, Wi, src.C,
                               k, stride, pad, st)
                acc = src.grad_mode()
                self._call(L.cy4_f32_to_f16, self.pool_scratch.data_ptr(), src.C, 1.0, None, src.gptr, src.ld, B * Hi * Wi, src.C, acc, st)
            elif kind == "up":
                _, out, src, Hi, Wi = r
                if not out.grad_has():
                    continue
                acc = src.grad_mode()
                self._call(L.cy4_upsample2x_bwd, out.gptr, out.ld, src.gptr, src.ld, B, Hi, Wi, src.C, acc, st)

        self._join_wgrad()
        if ar_groups is None:
            _t, n_items = self._unpack_all(st)
            self._unpack_range(st, 0, n_items)
        # BN parameter gradients: d beta = sum dz, d gamma = sum dz*xhat (both carry the loss scale)
        gbn = self.dbn * self.gscale[1]
        if ar_groups is not None:
            works.append(dist.all_reduce(gbn, op=dist.ReduceOp.AVG, async_op=True))
            for rec in self.convs:
                gb = grads.get(id(rec["conv"].bias)) if rec["conv"].bias is not None else None
                if gb is not None:
                    works.append(dist.all_reduce(gb, op=dist.ReduceOp.AVG, async_op=True))
            for w_ in works:
                w_.wait()                                # the current stream waits for the exchange; the host does not block
        # What autograd receives: views of a few BASE tensors (the flat conv-gradient buffer, the BN-gradient pair, the head
        # biases).  `materialise` hands out views of private copies of the bases: AccumulateGrad keeps the tensors it is given
        # as .grad, and the persistent / graph-static buffers are overwritten by the next backward.
        bases, layout = [gw_flat, gbn], {}
        for w, off in self._wslices:
            layout[id(w)] = (0, off, w.numel(), tuple(w.shape))
        tc = gbn.shape[1]
        for rec in self.convs:
            if rec["bn"] is not None and rec.get("bn_bwd_done"):
                c0, C = rec["coff"], rec["Cout"]
                layout[id(rec["bn"].bias)] = (1, c0, C, (C,))
                layout[id(rec["bn"].weight)] = (1, tc + c0, C, (C,))
            b = rec["conv"].bias
            if b is not None and id(b) in grads:
                layout[id(b)] = (len(bases), 0, b.numel(), tuple(b.shape))
                bases.append(grads[id(b)])
        self._grad_bases = bases
        self._grad_layout = [layout.get(id(p)) if p.requires_grad else None for p in self.params]
        return self.materialise()

    def materialise(self):
        """Private copies of the gradient base tensors, re-sliced into one tensor per parameter (None where no gradient)."""
        flats = [b.clone().reshape(-1) for b in self._grad_bases]
        return [None if v is None else flats[v[0]][v[1]:v[1] + v[2]].view(v[3]) for v in self._grad_layout]

    def _join_wgrad(self):
        """The compute stream waits for every weight-gradient launch issued so far on the side stream."""
        if self._wg_mode and self._wg_open:
            wg = self._wg
            wg["done"].record(wg["stream"])
            torch.cuda.current_stream().wait_event(wg["done"])
            self._wg_open = False
            self._slot_busy = [False] * len(self._slot_busy)

    def _allreduce_groups(self):
        """Gradient exchange plan: conv layers in backward order, cut into groups of roughly equal parameter
        bytes.  As soon as the weight gradients of a group are complete they are unpacked into their (contiguous) range of
        the flat OIHW buffer and averaged over the ranks with an asynchronous NCCL all-reduce that overlaps the rest of the
        backward pass -- the heavy parameters (1024-channel 3x3 layers) come first in backward, the large-activation
        layers that hide the exchange come last."""
        if getattr(self, "_ar_groups", None) is None:
            convs = [r for r in self.convs if not r["stem"]]
            total = sum(r["conv"].weight.numel() for r in convs)
            ngroups = max(1, int(getattr(self.model, "allreduce_groups", 6)))
            groups, cur, acc = [], [], 0
            for item, r in reversed(list(enumerate(convs))):
                cur.append((item, r))
                acc += r["conv"].weight.numel()
                if acc >= total / ngroups and len(groups) < ngroups - 1:
                    groups.append(cur); cur, acc = [], 0
            if cur:
                groups.append(cur)
            out = {}
            for g in groups:
                items = [i for i, _ in g]
                lo_rec, hi_rec = g[-1][1], g[0][1]              # lowest / highest layer index of the group
                flat_lo = lo_rec["woff"]
                flat_hi = hi_rec["woff"] + hi_rec["conv"].weight.numel()
                out[lo_rec["ind"]] = (min(items), len(items), flat_lo, flat_hi)
            if self.convs and self.convs[0]["stem"]:            # the stem's gradient is written by torch ops at the very end
                first = min(out)
                it, n, lo, hi = out.pop(first)
                out[self.convs[0]["ind"]] = (it, n, self.convs[0]["woff"], hi)
            self._ar_groups = out                                 # trigger layer index -> (first item, n items, flat range)
        return self._ar_groups

    def _unpack_all(self, st):
        """[Cout][tap][Cin] fp32 accumulators -> OIHW gradients of every conv, one launch.  The item
        table is built once per plan: accumulators and the flat gradient buffer are persistent."""
        from ._sigs_engine import UnpackItem
        if getattr(self, "_unpack_dev", None) is None:
            items, tiles = [], 0
            for rec in self.convs:
                if rec["stem"]:
                    continue
                it = UnpackItem()
                it.tile_begin = tiles
                tiles += (rup(rec["Cout"], 32) // 32) * (rup(rec["Cin"], 32) // 32)
                n = rec["conv"].weight.numel()
                it.dw_acc = rec["acc"].data_ptr()
                it.gw_oihw = self.gw_flat.data_ptr() + 4 * rec["woff"]
                it.Cout, it.Cin, it.ksize = rec["Cout"], rec["Cin"], rec["k"]
                items.append(it)
            self._unpack_n = len(items)
            if items:
                arr = (UnpackItem * len(items))(*items)
                self._unpack_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().to(self.device)
        return self._unpack_dev, self._unpack_n

    def _unpack_range(self, st, first, n):
        """One launch for items [first, first+n) of the unpack table."""
        from ._sigs_engine import UnpackItem
        table, total = self._unpack_all(st)
        if n > 0 and table is not None:
            self._call(self.L.cy4_unpack_wgrad_batched, table.data_ptr() + first * ctypes.sizeof(UnpackItem), n,
                       self.gscale[1:2].data_ptr(), st)

    def _conv_backward(self, rec, training, st, gw_flat, grads):
        L = self.L
        B = self.B
        inv_s = self.gscale[1:2]
        k, stride, pad, Cout, Cin = rec["k"], rec["stride"], rec["pad"], rec["Cout"], rec["Cin"]
        conv = rec["conv"]
        rec["bn_bwd_done"] = False
        if not (rec["A"].grad_has() if rec["bn"] is not None else rec.get("has_grad")):
            return                                       # no gradient reaches this layer
        mode, wg = self._wg_mode, self._wg
        slot = self._dy_next % len(self.dy_ring)
        self._dy_next += 1
        dy = self.dy_ring[slot]
        if self._slot_busy[slot]:                        # the wgrad that read this buffer last must be done
            torch.cuda.current_stream().wait_event(wg["slot"][slot])
            self._slot_busy[slot] = False
        if rec["bn"] is not None:
            A, Y = rec["A"], rec["Y"]
            A.zero_unwritten()                           # e.g. only a `groups` slice of A was consumed
            c0 = rec["coff"]
            q = [self.bnq[i, c0:].data_ptr() for i in range(4)]
            sdz, sdzx = self.dbn[0, c0:].data_ptr(), self.dbn[1, c0:].data_ptr()
            if rec["reduce_fused"] and rec.pop("_dz_ready", False):
                # the last writer of A's gradient (a consumer's cy4_conv_dgrad_fused) left dz in place of dA and the raw
                # sums (sum dz, sum dz*Y): only the per-channel conversion to sum dz*xhat is left
                self._call(L.cy4_bn_bwd_fixup, sdz, sdzx, q[2], q[3], Cout, st)
            else:
                self._call(L.cy4_bn_act_bwd_reduce, Y.buf.data_ptr(), Y.ld, A.gptr, A.ld, q[0], q[1], q[2], q[3], rec["act"],
                           Y.M, Cout, sdz, sdzx, st)
            ldy, M = rup(Cout, 64), Y.M                    # the dY scratch keeps 64-channel rows (weight-gradient TMA boxes)
            self._call(L.cy4_bn_act_bwd_apply, Y.buf.data_ptr(), Y.ld, A.gptr, A.ld, q[0], q[1], q[2], q[3], sdz, sdzx,
                       1.0 / Y.M, 1 if training else 0, rec["act"], 1, dy.data_ptr(), ldy, Y.M, Cout, st)
            rec["bn_bwd_done"] = True
            cpad = Cout
        else:
            rec["has_grad"] = False
            P = rec["P"]
            M = P.M
            ldy = 64
            cpad = rup(Cout, 32)
            # dP (fp32, true gradient) -> fp16 dY [M, 64] under the loss scale
            self._call(L.cy4_f32_to_f16, rec["dP"].data_ptr(), P.ld, 1.0, self.gscale.data_ptr(), dy.data_ptr(), ldy, M, cpad, 0, st)
            if conv.bias is not None:
                gb = torch.empty(Cout, device=self.device, dtype=torch.float32)
                self._call(L.cy4_colsum_f32, rec["dP"].data_ptr(), P.ld, M, Cout, 1.0, gb.data_ptr(), 0, st)
                grads[id(conv.bias)] = gb
        side = bool(mode) and not rec["stem"]
        if side and mode == 1:
            wg["fork"][rec["cidx"]].record()
        # input gradient
        if not rec["stem"]:
            src = rec["src"]
            acc = src.grad_mode()
            d = co.conv_desc(B, rec["Hi"], rec["Wi"], Cin, cpad, k, stride, pad, src.ld, ldy, co.CONV_ACCUM if acc else 0)
            P = rec["fuse_bwd"]
            if P is not None:
                pc = P["coff"]
                self._call(L.cy4_conv_dgrad_fused, ctypes.byref(d), dy.data_ptr(), rec["wd"].data_ptr(), src.gptr, P["Y"].buf.data_ptr(),
                           P["Y"].ld, self.bnq[0, pc:].data_ptr(), self.bnq[1, pc:].data_ptr(), P["act"],
                           self.dbn[0, pc:].data_ptr(), self.dbn[1, pc:].data_ptr(), st)
                P["_dz_ready"] = True
            else:
                self._call(L.cy4_conv_dgrad, ctypes.byref(d), dy.data_ptr(), rec["wd"].data_ptr(), src.gptr, st)
        # weight gradient (accumulator zeroed once per backward with the flat buffer)
        gw = gw_flat[rec["woff"]:rec["woff"] + conv.weight.numel()].view_as(conv.weight)
        if rec["stem"]:
            d = co.conv_desc(B, rec["Ho"], rec["Wo"], 32, Cout, 1, 1, 0, 32, ldy, co.CONV_A_MATRIX | co.CONV_ZERO_ACC)
            self._call(L.cy4_conv_wgrad, ctypes.byref(d), rec["cols"].buf.data_ptr(), dy.data_ptr(), rec["acc"].data_ptr(), st)
            gw.copy_((rec["acc"][:Cout, 0, :Cin * k * k] * inv_s).view(Cout, k, k, Cin).permute(0, 3, 1, 2))
        else:
            src = rec["src"]
            d = co.conv_desc(B, rec["Hi"], rec["Wi"], Cin, Cout, k, stride, pad, src.ld, ldy, co.CONV_ZERO_ACC)
            if side:
                if mode != 1:
                    wg["fork"][rec["cidx"]].record()
                ws = wg["stream"]
                ws.wait_event(wg["fork"][rec["cidx"]])
                self._call(L.cy4_conv_wgrad, ctypes.byref(d), src.ptr, dy.data_ptr(), rec["acc"].data_ptr(), ws.cuda_stream)
                wg["slot"][slot].record(ws)
                self._slot_busy[slot] = True
                self._wg_open = True
            else:
                self._call(L.cy4_conv_wgrad, ctypes.byref(d), src.ptr, dy.data_ptr(), rec["acc"].data_ptr(), st)
            # (unpacked for all layers by one launch at the end of backward)
        self._wslices.append((conv.weight, rec["woff"]))
