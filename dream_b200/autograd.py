"""Training forward/backward of DreamHourglass (vgg-Q family) on the hand-written kernels.

Reference: `DreamNetwork.train` -> `loss.backward()` (dream/network.py:328-338) differentiates
DreamHourglass.forward (dream/models.py:761-827) with torch autograd over cuDNN ops.  Here the whole
network is ONE `torch.autograd.Function`: forward runs the same kernel plan as inference while keeping
every layer's fp16 NHWC activation, backward walks the tape in reverse:

    dY --(ReLU mask from the saved output)--> bias grad (column sums)
       --> weight grad : dreamb200_wgrad   (tcgen05 GEMM over the pixel axis, split-K, fp32 atomics)
       --> data grad   : dreamb200_conv2d_fwd with the 180-degree-rotated, Cin<->Cout swapped weights
    2x2 max-pool / nearest-upsample backward in between.

The MSE / Huber loss and the optimizer stay PyTorch (north_star): the loss gradient arrives as
`grad_output` (fp32 NCHW) and parameter gradients leave as fp32 tensors in PyTorch OIHW layout.
fp16 has a narrow exponent range, so `grad_output` is multiplied by a power-of-two loss scale computed
on the device (no host sync) and parameter gradients are un-scaled in fp32.
"""
import torch

from . import models, ops

# Test hook: when set to a list, backward appends one record per conv layer:
# (key, dY as fed to the layer's kernels [fp16, already scaled+masked], cumulative scale [1-elem tensor],
#  layer input [fp16 NHWC], dX produced [fp16 NHWC or None]).  None = no capture, no cost.
DEBUG_CAPTURE = None
# Fold ReLU masks / re-scaling into the gradient-producing kernels (see backward).  DREAMB200_FUSE_GATES=0 restores
# the separate pass over every dY (kept for A/B timing).
import os as _os
_FUSE_GATES = _os.environ.get("DREAMB200_FUSE_GATES", "1") != "0"


def _dgrad_pack(weight, cin_pad, cout_pad):
    """Weights for the data gradient of a 3x3 'same' conv: dX[q] = sum_rs dY[q-(r-1,s-1)] W[:,:,r,s]^T."""
    wt = weight.detach().permute(1, 0, 2, 3)          # [Cin, Cout, 3, 3]: "output" channels are Cin now
    # kernel taps visited in reverse so the input offsets (1-r, 1-s) come out in the standard 3x3 order
    # (-1,-1)..(1,1): the data gradient then qualifies for the same kernels as a forward 3x3 conv
    rs = [(2 - r, 2 - s) for r in range(3) for s in range(3)]
    w = ops.pack_conv_weight(wt, rs, cin_pad=cout_pad, cout_pad=cin_pad)
    taps = [(1 - r, 1 - s) for r, s in rs]
    assert taps == ops.TAPS_3x3
    return w, taps


class _HourglassTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        P = model.plan()
        tape = []          # (kind, key, input, output)
        if model.n_image_input_channels == 3 and ops.wgrad_first_supported(x):
            # fused gather + pack + conv (first_conv.cu); its weight gradient is taken straight from x as well
            y = ops.first_conv3x3(x, P["first"].w, P["first"].b)
            tape.append(("first", "layer_0_1_down.0", x, y))
        elif model.n_image_input_channels == 3:
            t = ops.im2col_first(x, 3, 3, 1, 1, 64)
            y = models._run_conv(P["first"], t)
            tape.append(("first", "layer_0_1_down.0", t, y))
        else:
            # stage >= 2 of DreamHourglassMultiStage: image + previous belief maps, packed to 64 channels; an
            # ordinary conv layer whose data gradient is handed back to autograd ("input" entry below)
            t = ops.nchw_to_nhwc_f16(x, 64)
            y = models._run_conv(P["first"], t)
            tape.append(("input", None, None, None))
            tape.append(("conv", "layer_0_1_down.0", t, y))
        t = y
        sk = model.skip_connections
        skips = {}

        def add_skip(t, name):
            """hourglass skip connection (models.py:775-807): out = t + skip; both addends get the gradient"""
            out = ops.add_(t.clone(), skips[name])
            tape.append(("add", name, None, None))
            return out

        pooled = None
        for bi, (block, idxs, _) in enumerate(models.VGG_TRUNK):
            if bi > 0:
                y = pooled if pooled is not None else ops.maxpool(t, 2, 2, 0)
                pooled = None
                tape.append(("pool", None, t, y))
                t = y
                if sk:
                    skips["pool%d" % bi] = t
                    tape.append(("mark", "pool%d" % bi, None, None))
            for j in idxs:
                if block == "layer_0_1_down" and j == 0:
                    continue
                key = "%s.%d" % (block, j)
                pc = P[key]
                if j == idxs[-1] and bi < len(models.VGG_TRUNK) - 1 and ops.pool_fusion_pays(t.shape[1], t.shape[2]):
                    # nn.MaxPool2d(2) fused into this conv's epilogue; the tape needs the un-pooled tensor too
                    y, pooled = ops.conv_taps(t, pc.w, pc.b, pc.taps, t.shape[1], t.shape[2], relu=pc.relu,
                                              pool="both")
                else:
                    y = models._run_conv(pc, t)
                tape.append(("conv", key, t, y))
                t = y
            if sk and block == "layer_0_1_down":
                skips["x_0_1"] = t
                tape.append(("mark", "x_0_1", None, None))
        if sk:
            t = add_skip(t, "pool4")
        if model.deconv_decoder:
            # ConvTranspose2d(3, s2, p1, op1) + ReLU (+ conv3x3 + ReLU), models.py:618-686
            after = {"deconv_0_4": "pool3", "deconv_0_3": "pool2", "deconv_0_2": "pool1", "deconv_0_1": "x_0_1"}
            for name in ("deconv_0_4", "deconv_0_3", "deconv_0_2", "deconv_0_1"):
                y = models._run_deconv(P[name + ".0"], t)
                tape.append(("deconv", name + ".0", t, y))
                t = y
                if name != "deconv_0_1":
                    y = models._run_conv(P[name + ".2"], t)
                    tape.append(("conv", name + ".2", t, y))
                    t = y
                if sk:
                    t = add_skip(t, after[name])
        else:
            stages = [("upsample_0_4", ".4", ".6"), ("upsample_0_3", ".4", ".6")]
            if model.full_output:
                stages += [("upsample_0_2", ".2", ".4"), ("upsample_0_1", ".2", ".4")]
            for name, a, b in stages:
                y = ops.upsample2(t)
                tape.append(("up", None, t, y))
                t = y
                for suffix in (a, b):
                    y = models._run_conv(P[name + suffix], t)
                    tape.append(("conv", name + suffix, t, y))
                    t = y
                if sk and name == "upsample_0_4":
                    t = add_skip(t, "pool3")
        for key in ("heads_0.0", "heads_0.2"):
            y = models._run_conv(P[key], t)
            tape.append(("conv", key, t, y))
            t = y
        out = models._run_conv(P["heads_0.4"], t, head_cout=model.n_keypoints)
        tape.append(("head", "heads_0.4", t, None))
        ctx.tape = tape
        ctx.model = model
        ctx.param_names = [n for n, _ in model.named_parameters()]
        ctx.n_in = model.n_image_input_channels
        return out

    @staticmethod
    def backward(ctx, grad_out):
        model, tape = ctx.model, ctx.tape
        P = model.plan()
        grads = {}
        # multi-GPU training: finished parameter gradients go straight into the GradReducer's flat bucket buffer
        # (dream_b200.distributed), whose all-reduce starts while the layers below are still being differentiated
        sink = getattr(model, "_grad_sink", None)

        def emit(name, param, value):
            if sink is not None and sink.accepts(param):
                sink.deposit(param, value)
            else:
                grads[name] = value.contiguous()

        go = grad_out.contiguous().float()
        # fp16 gradients: every layer's dY is re-scaled by a power of two (computed on the device, no host
        # sync) so that max|dY| stays near 2^8; `cum` is the product of all factors applied so far and the
        # fp32 parameter gradients are divided by it.  Without this the trunk's dY drift into fp16
        # subnormals (gradient norms shrink ~400x from the head to the first layer).
        #
        # Where a layer's input is directly the ReLU output of the previous conv (or its 2x2 max pool), that ReLU's
        # mask and the re-scaling are applied by the kernel that PRODUCES the gradient -- the data-gradient conv's
        # epilogue (`gate`, `out_scale`) or the pool backward (`relu_gate`) -- instead of a separate pass over dY;
        # `ready` says the current g already went through them.  The scale then lags one layer behind (it is
        # derived from max|dY| of the layer above): adjacent layers' gradient magnitudes differ by far less than
        # the 2^7 of headroom on either side.
        # (`cum` is updated IN PLACE by ops.loss_scale_step; everything that needs an earlier value clones it)
        amax = go.abs().amax().reshape(1)
        cum = torch.ones((1,), dtype=torch.float32, device=go.device)
        f, inv = ops.loss_scale_step(amax, cum)
        g = ops.nchw_to_nhwc_f16((go * f).contiguous(), 64)             # [B,h,w,64], channels >= K are zero
        amax = ops.absmax(g)            # later layers get max|dY| for free from the producing data-gradient kernel
        ready = False
        db_ready = None                 # bias gradient of the next conv layer, when its producer already summed it
        stash = {}                      # skip connections: gradient of the skip addend, with the scale it carries
        gx = None
        fuse = _FUSE_GATES

        def producer_of(i):
            """(gate tensor or None, fusable) for the gradient flowing INTO tape entry i's output."""
            kind, key, _xin, yout = tape[i]
            if kind in ("conv", "deconv", "first"):
                pc = P["first"] if kind == "first" else P[key]
                return (yout if pc.relu else None), True
            return None, False

        for i in range(len(tape) - 1, -1, -1):
            kind, key, xin, yout = tape[i]
            if kind == "input":
                if ctx.needs_input_grad[1]:
                    gx = ops.nhwc_to_nchw_f32(g, ctx.n_in) * inv
                continue
            if kind == "add":
                stash[key] = (g.clone(), cum.clone())
                ready = False
                db_ready = None
                continue
            if kind == "mark":
                if key not in stash:            # this tensor was not used as a skip by the configured decoder
                    continue
                sg, scum = stash.pop(key)
                ops.scale_mask_(sg, None, cum / scum)           # bring it to the current loss scale, then accumulate
                ops.add_(g, sg)
                amax = ops.absmax(g)
                ready = False
                db_ready = None
                continue
            if kind in ("conv", "head", "first", "deconv"):
                node = models._node_for(model, key)
                pc = P["first"] if kind == "first" else P[key]
                deconv = kind == "deconv"
                if deconv:
                    cin, cout = node.weight.shape[0], node.weight.shape[1]      # ConvTranspose2d: [Cin, Cout, 3, 3]
                else:
                    cout, cin = node.weight.shape[0], node.weight.shape[1]
                if ready:                                   # mask + scale already applied by the producer of g
                    db = db_ready if db_ready is not None else (ops.bias_grad(g) if node.bias is not None else None)
                else:
                    # ReLU mask + re-scaling + bias gradient in one pass over dY
                    f, inv = ops.loss_scale_step(amax, cum)
                    db = ops.scale_mask_bias_(g, yout if pc.relu else None, f)
                if node.bias is not None:
                    emit(key + ".bias", node.bias, db[:cout] * inv)
                if kind == "first":
                    if xin.dtype == torch.float32:                                     # the raw input (fused path)
                        dw = ops.wgrad_first(g, xin)[:cout]                            # [co, (r,s,c)]
                    else:                                                              # the im2col'ed input
                        dw = ops.wgrad(g, xin, [(0, 0)])[0, :cout, :27]
                    emit(key + ".weight", node.weight, (dw * inv).view(cout, 3, 3, 3).permute(0, 3, 1, 2))
                    if DEBUG_CAPTURE is not None:
                        DEBUG_CAPTURE.append((key, None, cum.clone(), xin, None, None, None))
                    g = None                                                           # the image needs no grad
                    continue
                # what sits below this layer's input decides how its data gradient leaves the kernel
                gate, below_is_conv = producer_of(i - 1) if (fuse and i > 0) else (None, False)
                below_is_pool = fuse and i > 0 and tape[i - 1][0] == "pool"
                gate_t = gate if below_is_conv else None
                if below_is_pool and i > 1:
                    # the pooled tensor is a max over ReLU outputs: gating by it here equals gating inside the pool
                    # backward, and lets this kernel also sum the bias gradient of the conv below the pool
                    g2, conv2 = producer_of(i - 2)
                    if conv2 and g2 is not None:
                        gate_t = xin
                db_ready = None
                if gate_t is not None or below_is_conv:
                    db_ready = torch.zeros((xin.shape[3],), dtype=torch.float32, device=g.device)
                g_in = g
                cum_in = cum.clone() if DEBUG_CAPTURE is not None else None
                f_out = None
                inv_here = inv
                if ready and (below_is_conv or below_is_pool):
                    # `amax` is the measured max of this (already gated) dY: choose the factor its data gradient
                    # leaves with.  (After the unfused pass above dY was just normalised: factor 1.)
                    f_out, inv = ops.loss_scale_step(amax, cum)
                amax = torch.zeros((1,), dtype=torch.float32, device=g.device)
                B, H, W, _ = xin.shape
                if deconv:
                    # y[2p + (ky-1, kx-1)] += x[p] W[:, :, ky, kx]  =>  dW[tap] = sum_p dY[2p + tap-1] (x) X[p]
                    dw = ops.wgrad(g, xin, ops.TAPS_3x3, deconv=True)[:, :cout, :cin]      # [9, co, ci]
                    emit(key + ".weight", node.weight, (dw * inv_here).permute(2, 1, 0).reshape(cin, cout, 3, 3))
                    # dX[p] = sum_taps W[:, :, tap] dY[2p + tap-1]: a stride-2 3x3 conv over dY
                    rs = [(r, s_) for r in range(3) for s_ in range(3)]
                    wd = ops.pack_conv_weight(node.weight.detach(), rs, cin_pad=g.shape[3], cout_pad=xin.shape[3])
                    g = ops.conv_taps(g, wd, None, ops.TAPS_3x3, H, W, stride=2, absmax=amax,
                                      gate=gate_t, out_scale=f_out, colsum=db_ready)
                else:
                    dw = ops.wgrad(g, xin, ops.TAPS_3x3)[:, :cout, :cin]               # [9, co, ci]
                    emit(key + ".weight", node.weight, (dw * inv_here).permute(1, 2, 0).reshape(cout, cin, 3, 3))
                    wd, taps = _dgrad_pack(node.weight, xin.shape[3], g.shape[3])
                    g = ops.conv_taps(g, wd, None, taps, H, W, absmax=amax,
                                      gate=gate_t, out_scale=f_out, colsum=db_ready)
                ready = below_is_conv            # (below a pool the ReLU gate is applied by the pool backward)
                if DEBUG_CAPTURE is not None:
                    # (key, dY fed to the kernels, its scale, layer input, dX produced, ReLU gate folded into dX, dX's scale)
                    DEBUG_CAPTURE.append((key, g_in, cum_in, xin, g.clone(), gate_t, cum.clone()))
            elif kind == "pool":
                gate, below_is_conv = producer_of(i - 1) if (fuse and i > 0) else (None, False)
                g = ops.maxpool2_bwd(xin, g, relu_gate=below_is_conv and gate is not None)
                # the re-scaling for the layer below was folded into the data-gradient conv above the pool (`f_out`),
                # the ReLU gate of that layer is applied here (and, for the bias sum, already by the conv above);
                # max|g| does not grow, so `amax` stays a bound and a column sum taken above stays the bias gradient
                ready = below_is_conv
                if not ready:
                    db_ready = None
            elif kind == "up":
                g = ops.upsample2_bwd(g)        # sums 4 values: the stale `amax` can under-estimate by <= 4x (256x headroom)
                ready = False
                db_ready = None
        ctx.tape = None
        return (None, gx) + tuple(grads.get(n) for n in ctx.param_names)


def hourglass_train_forward(model, x):
    x = model._check_input(x)
    params = [p for _, p in model.named_parameters()]
    return _HourglassTrainFn.apply(model, x, *params)
