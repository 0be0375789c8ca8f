"""Training forward/backward of ResnetSimple (resnet-H / resnet-F) on the hand-written kernels.

Reference: `loss.backward()` (dream/network.py:328-338) through ResnetSimple.forward (dream/models.py:140-155):
torchvision ResNet-101 trunk (Bottleneck: 1x1 -> 3x3 (stride) -> 1x1, BatchNorm after each, residual add + ReLU),
4-5 x [ConvTranspose2d(4,2,1) + BatchNorm + ReLU], 1x1 head.  BatchNorm runs in training mode: batch
statistics per process (the reference's DataParallel replicas also normalise per replica), running statistics
updated with momentum 0.1 and the unbiased variance like nn.BatchNorm2d.

One `torch.autograd.Function` for the whole network.  Forward keeps, per conv+BN unit, the input, the raw conv
output z (fp16), the batch mean / inverse std and the activated output; backward walks the units in reverse
with a gradient accumulator per tensor (a block's input feeds both the 1x1 and the identity / downsample path):

  dY --ReLU mask--> (copy to the residual's accumulator) --> BatchNorm backward (two per-channel reductions +
  one elementwise pass) --> weight gradient (dreamb200_wgrad / _wgrad_strided / _wgrad_deconv) and data
  gradient (forward conv kernels with swapped weights; stride-2 layers as sub-pixel phases) --> accumulate.

fp16 gradients carry one power-of-two loss scale that is refreshed (on the device) at every block boundary, where
a single gradient tensor is live.
"""
import torch

from . import models, ops

BN_MOMENTUM = 0.1
EPS = 1e-5

# Test hook: when a list, backward appends per conv+BN unit a dict with the unit, the incoming dL/dy (fp16 clone,
# before the ReLU mask), the scale it carries, and the produced dL/dx (fp16 clone) with its scale.
DEBUG_CAPTURE = None


def _pad_vec(v, n):
    out = torch.zeros((n,), dtype=torch.float32, device=v.device)
    out[: v.numel()] = v.detach().float()
    return out


def _train_plan(model):
    """Raw (un-folded) packed weights for training, cached on the model next to the inference plan."""
    key = tuple((p.data_ptr(), p._version) for p in model.parameters())
    cached = getattr(model, "_train_plan_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    P = {}
    with torch.no_grad():
        c1 = models._node_for(model, "conv1")
        P["conv1"] = models._PackedConv(ops.pack_first_weight(c1.weight, 192), None, [(0, 0)], cout=64)
        for li, nblocks in enumerate(models.RESNET101_BLOCKS, start=1):
            for bi in range(nblocks):
                k = "layer%d.%d" % (li, bi)
                stride = 2 if (li > 1 and bi == 0) else 1
                P[k + ".conv1"] = models._pack3x3(models._node_for(model, k + ".conv1"), relu=False)
                P[k + ".conv2"] = models._pack3x3(models._node_for(model, k + ".conv2"), relu=False, stride=stride)
                P[k + ".conv3"] = models._pack3x3(models._node_for(model, k + ".conv3"), relu=False)
                if bi == 0:
                    P[k + ".downsample.0"] = models._pack3x3(models._node_for(model, k + ".downsample.0"), relu=False,
                                                             stride=stride)
        dec = ["upsample.%d" % (3 * i) for i in range(4)] + (["upsample2.0"] if model.full else [])
        for name in dec:
            P[name] = models._pack_deconv(models._node_for(model, name), 4, relu=False)
        head = "upsample2.3" if model.full else "upsample.12"
        P[head] = models._pack3x3(models._node_for(model, head), relu=False, cout_pad=16)
    model._train_plan_cache = (key, P)
    return P


class _Unit:
    """One conv (+bias) + BatchNorm (+residual) (+ReLU) unit of the tape."""
    __slots__ = ("kind", "conv_key", "bn_key", "x", "z", "y", "mean", "invstd", "rows", "residual", "relu",
                 "stride", "block_end")


def _bn_forward(model, bn_key, z, residual, relu):
    bn = models._node_for(model, bn_key)
    Cp = z.shape[3]
    rows = z.numel() // Cp
    s, ss = ops.bn_stats(z)
    mean = s / rows
    var = (ss / rows - mean * mean).clamp_min(0.0)
    invstd = torch.rsqrt(var + EPS)
    gamma, beta = _pad_vec(bn.weight, Cp), _pad_vec(bn.bias, Cp)
    scale = gamma * invstd
    shift = beta - mean * scale
    y = ops.bn_apply(z, scale, shift, residual=residual, relu=relu)
    with torch.no_grad():                                       # nn.BatchNorm2d running statistics (momentum 0.1)
        c = bn.weight.numel()
        bn.running_mean.mul_(1 - BN_MOMENTUM).add_(mean[:c], alpha=BN_MOMENTUM)
        bn.running_var.mul_(1 - BN_MOMENTUM).add_(var[:c] * (rows / max(rows - 1, 1)), alpha=BN_MOMENTUM)
        bn.num_batches_tracked.add_(1)
    return y, mean, invstd, rows


class _ResnetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        P = _train_plan(model)
        tape = []

        def unit(kind, conv_key, bn_key, xin, relu, residual=None, stride=1):
            pc = P[conv_key]
            if kind == "deconv":
                z = models._run_deconv(pc, xin)                 # ConvTranspose2d(4,2,1) + bias (no ReLU here)
            else:
                z = models._run_conv(pc, xin)
            y, mean, invstd, rows = _bn_forward(model, bn_key, z, residual, relu)
            u = _Unit()
            u.kind, u.conv_key, u.bn_key, u.x, u.z, u.y = kind, conv_key, bn_key, xin, z, y
            u.mean, u.invstd, u.rows, u.residual, u.relu, u.stride, u.block_end = mean, invstd, rows, residual, relu, stride, False
            tape.append(u)
            return y

        patches = ops.im2col_first(x, 7, 7, 2, 3, 192)
        t = unit("first", "conv1", "bn1", patches, relu=True)
        pooled = ops.maxpool(t, 3, 2, 1)
        tape.append(("maxpool3", t, pooled))
        t = pooled
        for li, nblocks in enumerate(models.RESNET101_BLOCKS, start=1):
            for bi in range(nblocks):
                k = "layer%d.%d" % (li, bi)
                stride = 2 if (li > 1 and bi == 0) else 1
                o = unit("conv", k + ".conv1", k + ".bn1", t, relu=True)
                o = unit("conv", k + ".conv2", k + ".bn2", o, relu=True, stride=stride)
                idn = unit("conv", k + ".downsample.0", k + ".downsample.1", t, relu=False, stride=stride) \
                    if bi == 0 else t
                t = unit("conv", k + ".conv3", k + ".bn3", o, relu=True, residual=idn)
                tape[-1].block_end = True
        dec = [("upsample.%d" % (3 * i), "upsample.%d" % (3 * i + 1)) for i in range(4)]
        if model.full:
            dec.append(("upsample2.0", "upsample2.1"))
        for ck, bk in dec:
            t = unit("deconv", ck, bk, t, relu=True)
            tape[-1].block_end = True
        head = "upsample2.3" if model.full else "upsample.12"
        out = models._run_conv(P[head], t, head_cout=model.n_keypoints)
        ctx.tape, ctx.model, ctx.head, ctx.head_in = tape, model, head, t
        ctx.param_names = [n for n, _ in model.named_parameters()]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        model, tape = ctx.model, ctx.tape
        grads = {}
        sink = getattr(model, "_grad_sink", None)                # multi-GPU: dream_b200.distributed.GradReducer

        def emit(name, param, value):
            if sink is not None and sink.accepts(param):
                sink.deposit(param, value)                       # its bucket's all-reduce may start right here
            else:
                grads[name] = value.contiguous()

        G = {}                                                   # id(tensor) -> accumulated fp16 gradient

        def acc(t, g):
            if id(t) in G:
                ops.add_(G[id(t)], g)
            else:
                G[id(t)] = g

        go = grad_out.contiguous().float()
        amax = go.abs().amax().clamp_min(1e-30)
        cum = torch.exp2(torch.floor(torch.log2(256.0 / amax))).reshape(1)
        g = ops.nchw_to_nhwc_f16((go * cum).contiguous(), 64)
        # ---- head: 1x1 conv + bias ----
        hn = models._node_for(model, ctx.head)
        K, cin = hn.weight.shape[0], hn.weight.shape[1]
        inv = 1.0 / cum
        emit(ctx.head + ".bias", hn.bias, ops.bias_grad(g)[:K] * inv)
        dw = ops.wgrad(g, ctx.head_in, [(0, 0)])[0, :K, :cin]
        emit(ctx.head + ".weight", hn.weight, (dw * inv).reshape(K, cin, 1, 1))
        wd = ops.pack_conv_weight(hn.weight.detach().permute(1, 0, 2, 3), [(0, 0)], cin_pad=64,
                                  cout_pad=ctx.head_in.shape[3])
        B, H, W, _ = ctx.head_in.shape
        acc(ctx.head_in, ops.conv_taps(g, wd, None, [(0, 0)], H, W))

        for u in reversed(tape):
            if isinstance(u, tuple):                             # ("maxpool3", x, y)
                _, xin, yout = u
                acc(xin, ops.maxpool3_bwd(xin, G.pop(id(yout))))
                continue
            g = G.pop(id(u.y))
            cap = None
            if DEBUG_CAPTURE is not None:
                cap = {"unit": u, "g_in": g.clone(), "cum_in": cum.clone()}
                DEBUG_CAPTURE.append(cap)
            if u.block_end:
                # single live gradient here: refresh the loss scale (power of two, computed on the device)
                f = torch.exp2(torch.floor(torch.log2(256.0 / ops.absmax(g).clamp_min(1e-30)))).clamp(2.0 ** -12, 2.0 ** 12)
                ops.scale_mask_(g, u.y if u.relu else None, f)
                cum = cum * f
                # every other pending accumulator belongs to an earlier tensor that nothing has touched yet
                assert len(G) == 0
            elif u.relu:
                ops.scale_mask_(g, u.y)
            inv = 1.0 / cum
            if u.residual is not None:
                acc(u.residual, g.clone())
            # ---- BatchNorm backward ----
            bn = models._node_for(model, u.bn_key)
            c = bn.weight.numel()
            Cp = u.z.shape[3]
            sum_dy, sum_dyz = ops.bn_bwd_reduce(g, u.z)
            gamma = _pad_vec(bn.weight, Cp)
            dgamma = u.invstd * (sum_dyz - u.mean * sum_dy)
            emit(u.bn_key + ".weight", bn.weight, dgamma[:c] * inv)
            emit(u.bn_key + ".bias", bn.bias, sum_dy[:c] * inv)
            A = gamma * u.invstd
            Bc = -A * u.invstd * (dgamma / u.rows)
            Cc = -A * (sum_dy / u.rows) - Bc * u.mean
            ops.bn_bwd_apply_(g, u.z, A, Bc, Cc)                 # g is now dL/dz
            # ---- conv / deconv backward ----
            node = models._node_for(model, u.conv_key)
            if u.kind == "deconv":
                ci, co = node.weight.shape[0], node.weight.shape[1]
                taps = [(ky - 1, kx - 1) for ky in range(4) for kx in range(4)]
                rs = [(ky, kx) for ky in range(4) for kx in range(4)]
                emit(u.conv_key + ".bias", node.bias, ops.bias_grad(g)[:co] * inv)
                dw = ops.wgrad(g, u.x, taps, deconv=True)[:, :co, :ci]
                emit(u.conv_key + ".weight", node.weight, (dw * inv).permute(2, 1, 0).reshape(ci, co, 4, 4))
                wd = ops.pack_conv_weight(node.weight.detach(), rs, cin_pad=g.shape[3], cout_pad=u.x.shape[3])
                B, H, W, _ = u.x.shape
                dx = ops.conv_taps(g, wd, None, taps, H, W, stride=2)
                if cap is not None:
                    cap["dx"], cap["cum_out"] = dx.clone(), cum.clone()
                acc(u.x, dx)
            elif u.kind == "first":
                co = node.weight.shape[0]
                dw = ops.wgrad(g, u.x, [(0, 0)])[0, :co, :147]                        # [co, (r,s,c)]
                emit(u.conv_key + ".weight", node.weight, (dw * inv).view(co, 7, 7, 3).permute(0, 3, 1, 2))
                if cap is not None:
                    cap["cum_out"] = cum.clone()
            else:
                co, ci, ksz = node.weight.shape[0], node.weight.shape[1], node.weight.shape[2]
                pad = ksz // 2
                rs = [(r, s_) for r in range(ksz) for s_ in range(ksz)]
                taps = [(r - pad, s_ - pad) for r, s_ in rs]
                B, Hx, Wx, _ = u.x.shape
                if u.stride == 1:
                    dw = ops.wgrad(g, u.x, taps)[:, :co, :ci]
                    rs_rev = [(ksz - 1 - r, ksz - 1 - s_) for r, s_ in rs]
                    wd = ops.pack_conv_weight(node.weight.detach().permute(1, 0, 2, 3), rs_rev, cin_pad=g.shape[3],
                                              cout_pad=u.x.shape[3])
                    dx = ops.conv_taps(g, wd, None, taps, Hx, Wx)
                else:
                    dw = ops.wgrad_strided(g, u.x, taps)[:, :co, :ci]
                    dx = _dgrad_stride2(node.weight.detach(), g, u.x.shape)
                emit(u.conv_key + ".weight", node.weight, (dw * inv).permute(1, 2, 0).reshape(co, ci, ksz, ksz))
                if cap is not None:
                    cap["dx"], cap["cum_out"] = dx.clone(), cum.clone()
                acc(u.x, dx)
        ctx.tape = None
        return (None, None) + tuple(grads.get(n) for n in ctx.param_names)


def _dgrad_stride2(weight, g, x_shape):
    """Data gradient of a stride-2 conv (k3 p1 or k1 p0) as sub-pixel phases writing an interleaved view of dX
    (the adjoint of the forward's ConvTranspose decomposition, models._deconv_phase_taps)."""
    B, Hx, Wx, Cx = x_shape
    co, ci, ksz = weight.shape[0], weight.shape[1], weight.shape[2]
    dx = torch.zeros((B, Hx, Wx, Cx), dtype=torch.float16, device=g.device)
    wt = weight                                                   # as a ConvTranspose weight: [in=co, out=ci, k, k]
    strides = (2 * Cx, 2 * Wx * Cx, Hx * Wx * Cx)
    for py in range(2):
        for px in range(2):
            if ksz == 1:
                if py or px:
                    continue
                ty, tx = [(0, 0)], [(0, 0)]
            else:
                ty, tx = models._deconv_phase_taps(ksz, py), models._deconv_phase_taps(ksz, px)
            Hp, Wp = (Hx - py + 1) // 2, (Wx - px + 1) // 2
            if Hp <= 0 or Wp <= 0:
                continue
            rs = [(ky, kx) for (_, ky) in ty for (_, kx) in tx]
            taps = [(dy, dx_) for (dy, _) in ty for (dx_, _) in tx]
            w = ops.pack_conv_weight(wt.permute(1, 0, 2, 3), rs, cin_pad=g.shape[3], cout_pad=Cx)
            ops.conv_taps(g, w, None, taps, Hp, Wp, y=dx, y_strides=strides, y_offset=(py * Wx + px) * Cx)
    return dx


def resnet_train_forward(model, x):
    x = model._check_input(x)
    params = [p for _, p in model.named_parameters()]
    return _ResnetTrainFn.apply(model, x, *params)
