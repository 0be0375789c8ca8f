"""B200-native DREAM networks: reference-compatible parameter trees + a layer plan executed by
libdreamb200's tensor-core kernels.

Reference: dream/models.py -- DreamHourglass (:557-827, vgg-Q / vgg-F), ResnetSimple (:17-155,
resnet-H / resnet-F).  The modules here are NOT nn.Sequential stacks of torch ops: they hold fp32
`nn.Parameter`s / BN buffers under exactly the reference's state-dict names (so released `.pth`
files load, `torch.save` round-trips and torch optimizers update them in place), and
`forward` runs a plan of C-ABI calls on NHWC fp16 activations:

  first layer : dreamb200_im2col_first (patch gather fused with the fp32 NCHW -> fp16 NHWC pack)
                + one 1-tap GEMM,
  conv 3x3    : 9-tap implicit GEMM, bias+ReLU in the epilogue,
  1x1 / s2    : 1-tap GEMM, TMA element stride 2 for strided layers,
  BN (eval)   : folded into the packed fp16 weights + fp32 bias at pack time,
  bottleneck  : residual add + ReLU fused in the epilogue of the block's last 1x1,
  ConvTranspose k3s2p1op1 / k4s2p1 : 4 sub-pixel phases, each a small-tap conv writing an
                interleaved (strided TMA store) view of the output -- no zero insertion,
  head        : last conv writes fp32 NCHW belief maps directly.

Packed weights are cached and re-packed only when a parameter's version counter changes.
"""
import math

import torch
import torch.nn as nn

from . import ops

VGG_TRUNK = [
    ("layer_0_1_down", (0, 2), 64),
    ("layer_0_2_down", (5, 7), 128),
    ("layer_0_3_down", (10, 12, 14, 16), 256),
    ("layer_0_4_down", (19, 21, 23, 25), 512),
    ("layer_0_5_down", (28, 30, 32, 34), 512),
]
RESNET101_BLOCKS = (3, 4, 23, 3)
BN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------
# parameter tree helpers
# ----------------------------------------------------------------------------------------------
class _Node(nn.Module):
    """Container with no forward: only gives parameters their reference state-dict path."""


def _node_for(root, dotted):
    mod = root
    parts = dotted.split(".")
    for p in parts:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    return mod


def _add_conv(root, key, cin, cout, k, bias=True, transposed=False):
    n = _node_for(root, key)
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    w = torch.empty(shape)
    nn.init.kaiming_uniform_(w, a=math.sqrt(5))          # nn.Conv2d / nn.ConvTranspose2d default
    n.weight = nn.Parameter(w)
    if bias:
        fan_in = shape[1] * k * k
        bound = 1.0 / math.sqrt(fan_in)
        n.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
    else:
        n.register_parameter("bias", None)


def _add_bn(root, key, c):
    n = _node_for(root, key)
    n.weight = nn.Parameter(torch.ones(c))
    n.bias = nn.Parameter(torch.zeros(c))
    n.register_buffer("running_mean", torch.zeros(c))
    n.register_buffer("running_var", torch.ones(c))
    n.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


def _deconv_phase_taps(k, phase):
    """ConvTranspose (stride 2, padding 1): output row 2*y'+phase receives kernel rows ky with
    (phase + 1 - ky) even, from input row y' + (phase + 1 - ky) / 2."""
    return [((phase + 1 - ky) // 2, ky) for ky in range(k) if (phase + 1 - ky) % 2 == 0]


class _PackedConv:
    """One plan step: packed fp16 weights + fp32 bias + tap list for dreamb200_conv2d_fwd."""
    __slots__ = ("w", "b", "taps", "stride", "relu", "cout", "phases")

    def __init__(self, w, b, taps, stride=1, relu=False, cout=None, phases=None):
        self.w, self.b, self.taps, self.stride, self.relu, self.cout, self.phases = \
            w, b, taps, stride, relu, cout, phases


def _bn_scale_shift(bn, conv_bias):
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + BN_EPS)
    base = conv_bias.detach().float() if conv_bias is not None else torch.zeros_like(scale)
    shift = (base - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return scale, shift


def _pack3x3(node, relu, bn=None, stride=1, cout_pad=None, cin_pad=None):
    scale, bias = (None, node.bias)
    if bn is not None:
        scale, bias = _bn_scale_shift(bn, node.bias)
    ksz = node.weight.shape[2]
    rs = [(r, s) for r in range(ksz) for s in range(ksz)]
    w = ops.pack_conv_weight(node.weight, rs, cin_pad=cin_pad, cout_pad=cout_pad, scale=scale)
    taps = [(r - ksz // 2, s - ksz // 2) for r, s in rs]
    return _PackedConv(w, ops.pad_bias(bias, w.shape[1], w.device), taps, stride=stride, relu=relu,
                       cout=node.weight.shape[0])


def _pack_deconv(node, k, relu, bn=None):
    """ConvTranspose2d weight [Cin,Cout,k,k] -> 4 phase convs (py,px) with their own tap lists."""
    scale, bias = (None, node.bias)
    if bn is not None:
        scale, bias = _bn_scale_shift(bn, node.bias)
    w_oihw = node.weight.detach().permute(1, 0, 2, 3)            # -> [Cout,Cin,k,k]
    cout = w_oihw.shape[0]
    phases = []
    for py in range(2):
        for px in range(2):
            ty, tx = _deconv_phase_taps(k, py), _deconv_phase_taps(k, px)
            rs = [(ky, kx) for (_, ky) in ty for (_, kx) in tx]
            taps = [(dy, dx) for (dy, _) in ty for (dx, _) in tx]
            w = ops.pack_conv_weight(w_oihw, rs, scale=scale)
            phases.append((py, px, w, taps))
    b = ops.pad_bias(bias, ops.round_up(cout, 64), node.weight.device)
    return _PackedConv(None, b, None, relu=relu, cout=cout, phases=phases)


def _pack_upsampled_conv(node, relu):
    """nn.Upsample(scale_factor=2) (nearest) followed by a 3x3 'same' conv (models.py:691-697,703-709),
    folded: output pixel (2y'+py, 2x'+px) reads up[2y'+py+r-1] = in[(2y'+py+r-1)//2], so the three kernel
    rows collapse onto two input rows per phase ((r) -> dy: py=0: {0:-1, 1:0, 2:0}; py=1: {0:0, 1:0, 2:+1})
    and likewise for columns.  Each of the 4 output phases is a 2x2-tap conv on the LOW-resolution input with
    weights pre-summed in fp32 -- 4/9 of the MACs and no upsampled tensor in HBM."""
    w = node.weight.detach().float()                      # [Cout, Cin, 3, 3]
    cout = w.shape[0]
    rows = {0: {-1: [0], 0: [1, 2]}, 1: {0: [0, 1], 1: [2]}}
    phases = []
    for py in range(2):
        for px in range(2):
            mats, taps = [], []
            for dy, rr in rows[py].items():
                for dx, ss in rows[px].items():
                    acc = torch.zeros_like(w[:, :, 0, 0])
                    for r in rr:
                        for s_ in ss:
                            acc = acc + w[:, :, r, s_]
                    mats.append(acc)
                    taps.append((dy, dx))
            w4 = torch.stack(mats, dim=-1).unsqueeze(-1)                 # [Cout, Cin, T, 1] -> taps along dim 2
            packed = ops.pack_conv_weight(w4, [(t, 0) for t in range(len(taps))])
            phases.append((py, px, packed, taps))
    b = ops.pad_bias(node.bias, ops.round_up(cout, 64), node.weight.device)
    return _PackedConv(None, b, None, relu=relu, cout=cout, phases=phases)


def _run_conv(pc, x, residual=None, head_cout=None, residual_f32=None, want_f32=False):
    B, H, W, _ = x.shape
    if pc.stride == 1:
        Ho, Wo = H, W
    else:
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1      # k=1/p=0 and k=3/p=1 give the same size
    y_f32 = None
    if want_f32:
        y_f32 = torch.empty((B, Ho, Wo, pc.w.shape[1]), dtype=torch.float32, device=x.device)
    y = ops.conv_taps(x, pc.w, pc.b, pc.taps, Ho, Wo, stride=pc.stride, relu=pc.relu, residual=residual,
                      head_cout=head_cout, residual_f32=residual_f32, y_f32=y_f32)
    return (y, y_f32) if want_f32 else y


def _run_deconv(pc, x):
    B, H, W, _ = x.shape
    cpad = pc.b.numel()
    Ho, Wo = 2 * H, 2 * W
    y = torch.empty((B, Ho, Wo, cpad), dtype=torch.float16, device=x.device)
    strides = (2 * cpad, 2 * Wo * cpad, Ho * Wo * cpad)
    # the four phases go to the library as one group: one launch when they are uniform (dreamb200_conv2d_fwd_phases)
    ops.conv_phases(x, [((py * Wo + px) * cpad, w, taps) for py, px, w, taps in pc.phases], pc.b, H, W, y, strides,
                    relu=pc.relu)
    return y


class _LazyPlan(dict):
    """Plan entries are packed on first use: a training step only touches the training-tape entries, inference only
    the fused ones (e.g. the upsample-folded phase convs), and the packs are redone after every optimizer step."""

    def __init__(self, builders):
        super().__init__()
        self._builders = builders

    def __missing__(self, key):
        with torch.no_grad():
            value = self._builders[key]()
        self[key] = value
        return value

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._builders


class _PlanModule(nn.Module):
    """Shared machinery: lazy weight packing keyed on parameter versions; eval-only in this round's
    forward (training goes through dream_b200.autograd)."""

    def __init__(self):
        super().__init__()
        self._plan = None
        self._plan_key = None

    def _version_key(self):
        key = [self.training]
        for t in list(self.parameters()) + list(self.buffers()):
            key.append((t.data_ptr(), t._version))
        return tuple(key)

    def invalidate_packed_weights(self):
        """Forget the packed fp16 weights.  They are rebuilt whenever a parameter's or buffer's version counter (or
        storage) changes -- every in-place torch op bumps it.  Updates that bypass the counters do not: torch's
        single-kernel optimizers (`torch.optim.Adam(..., fused=True)`) and writes through raw pointers; call this
        after such an update (DreamNetwork.train uses the default multi-tensor optimizers and needs nothing)."""
        self._plan = self._plan_key = None
        if hasattr(self, "_pp"):
            self._pp = None

    def plan(self):
        key = self._version_key()
        if self._plan is None or key != self._plan_key:
            with torch.no_grad():
                self._plan = self._build_plan()
            self._plan_key = key
        return self._plan

    # (mean[3], stdev[3]) of the config's image_normalization; DreamNetwork sets it.  Only used when the
    # input is a raw uint8 [B,H,W,3] batch (the dataset's ToTensor + Normalize on device).
    input_normalization = ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0))

    @staticmethod
    def _is_raw_frames(x):
        return x.dtype == torch.uint8

    def _check_input(self, x):
        if not x.is_cuda:
            raise RuntimeError("dream_b200 runs on a CUDA device only (no CPU fallback); got a CPU tensor")
        if self._is_raw_frames(x):
            assert x.dim() == 4 and x.shape[3] == 3, \
                "expected raw uint8 frames [B,H,W,3], got {}".format(tuple(x.shape))
            from .image_proc import normalize_u8_device
            return normalize_u8_device(x, *self.input_normalization)
        n_in = getattr(self, "n_image_input_channels", 3)
        assert x.dim() == 4 and x.shape[1] == n_in, \
            "expected an input batch [B,{},H,W], got {}".format(n_in, tuple(x.shape))
        return x.contiguous().float()


# ----------------------------------------------------------------------------------------------
# DreamHourglass (vgg-Q / vgg-F)
# ----------------------------------------------------------------------------------------------
class DreamHourglass(_PlanModule):
    """dream/models.py:557-827.  Same constructor keywords; `forward(x)` returns `[belief_maps]`
    (plus the soft-argmax keypoints when `internalize_spatial_softmax`)."""

    def __init__(self, n_keypoints, n_image_input_channels=3, internalize_spatial_softmax=True,
                 learned_beta=True, initial_beta=1.0, skip_connections=False, deconv_decoder=False,
                 full_output=False):
        super().__init__()
        assert 1 <= n_image_input_channels <= 64, "the first layer takes 1..64 input channels"
        self.n_keypoints = n_keypoints
        self.n_image_input_channels = n_image_input_channels
        self.internalize_spatial_softmax = internalize_spatial_softmax
        self.skip_connections = skip_connections
        self.deconv_decoder = deconv_decoder
        self.full_output = full_output
        if internalize_spatial_softmax:
            self.n_output_heads = 2
            self.learned_beta = learned_beta
            self.initial_beta = initial_beta
        else:
            self.n_output_heads = 1
            self.learned_beta = False
        cin = n_image_input_channels
        for block, idxs, ch in VGG_TRUNK:
            for j in idxs:
                _add_conv(self, "%s.%d" % (block, j), cin, ch, 3)
                cin = ch
        if deconv_decoder:
            for name, ci, co, with_conv in (("deconv_0_4", 512, 256, True), ("deconv_0_3", 256, 128, True),
                                            ("deconv_0_2", 128, 64, True), ("deconv_0_1", 64, 64, False)):
                _add_conv(self, name + ".0", ci, co, 3, transposed=True)
                if with_conv:
                    _add_conv(self, name + ".2", co, co, 3)
        else:
            _add_conv(self, "upsample_0_4.4", 512, 256, 3); _add_conv(self, "upsample_0_4.6", 256, 256, 3)
            _add_conv(self, "upsample_0_3.4", 256, 128, 3); _add_conv(self, "upsample_0_3.6", 128, 64, 3)
            if full_output:
                for name in ("upsample_0_2", "upsample_0_1"):
                    _add_conv(self, name + ".2", 64, 64, 3); _add_conv(self, name + ".4", 64, 64, 3)
        _add_conv(self, "heads_0.0", 64, 64, 3)
        _add_conv(self, "heads_0.2", 64, 32, 3)
        _add_conv(self, "heads_0.4", 32, n_keypoints, 3)
        # the reference's trunk is torchvision's pretrained VGG-19 (models.py:587); taken from a LOCAL checkpoint when
        # there is one, otherwise random with a warning (dream_b200/pretrained.py) -- never downloaded
        from . import pretrained as _pretrained
        _pretrained.load_vgg19_trunk(self)
        if internalize_spatial_softmax:
            from .spatial_softmax import SoftArgmaxPavlo
            self.softmax = nn.Sequential()
            self.softmax.add_module("0", SoftArgmaxPavlo(n_keypoints=n_keypoints, learned_beta=self.learned_beta,
                                                         initial_beta=self.initial_beta))

    def _n(self, key):
        return _node_for(self, key)

    def _build_plan(self):
        n = self._n
        Bd = {}
        first = n("layer_0_1_down.0")
        if self.n_image_input_channels == 3:
            def pack_first():
                wf = ops.pack_first_weight(first.weight, 64)
                return _PackedConv(wf, ops.pad_bias(first.bias, 64, wf.device), [(0, 0)], relu=True, cout=64)
            Bd["first"] = pack_first
        else:
            # later stages of DreamHourglassMultiStage see image + previous belief maps (models.py:404-407):
            # an ordinary 9-tap layer on the input packed to 64 zero-padded channels
            Bd["first"] = lambda: _pack3x3(first, relu=True, cin_pad=64)
        for block, idxs, _ in VGG_TRUNK:
            for j in idxs:
                if block == "layer_0_1_down" and j == 0:
                    continue
                Bd["%s.%d" % (block, j)] = lambda k="%s.%d" % (block, j): _pack3x3(n(k), relu=True)
        if self.deconv_decoder:
            for name in ("deconv_0_4", "deconv_0_3", "deconv_0_2", "deconv_0_1"):
                Bd[name + ".0"] = lambda k=name + ".0": _pack_deconv(n(k), 3, relu=True)
                if name != "deconv_0_1":
                    Bd[name + ".2"] = lambda k=name + ".2": _pack3x3(n(k), relu=True)
        else:
            for name in ("upsample_0_4", "upsample_0_3"):
                Bd[name + ".4"] = lambda k=name + ".4": _pack3x3(n(k), relu=True)              # training tape
                Bd[name + ".4/up"] = lambda k=name + ".4": _pack_upsampled_conv(n(k), relu=True)   # inference
                Bd[name + ".6"] = lambda k=name + ".6": _pack3x3(n(k), relu=False)
            if self.full_output:
                for name in ("upsample_0_2", "upsample_0_1"):
                    Bd[name + ".2"] = lambda k=name + ".2": _pack3x3(n(k), relu=True)
                    Bd[name + ".2/up"] = lambda k=name + ".2": _pack_upsampled_conv(n(k), relu=True)
                    Bd[name + ".4"] = lambda k=name + ".4": _pack3x3(n(k), relu=True)
        Bd["heads_0.0"] = lambda: _pack3x3(n("heads_0.0"), relu=True)
        Bd["heads_0.2"] = lambda: _pack3x3(n("heads_0.2"), relu=True)            # 32 real + 32 zero channels
        Bd["heads_0.4"] = lambda: _pack3x3(n("heads_0.4"), relu=False, cout_pad=16, cin_pad=64)
        P = _LazyPlan(Bd)
        if self.n_image_input_channels != 3:
            Bd["layer_0_1_down.0"] = lambda: P["first"]
        return P

    def belief_maps(self, x):
        """Inference forward: fp32 NCHW [B,3,H,W] (cuda) -> fp32 NCHW belief maps [B,K,h,w].
        A uint8 [B,H,W,3] batch is taken as raw frames and normalised inside the first layer's gather."""
        P = self.plan()
        if self.n_image_input_channels != 3:
            t = _run_conv(P["first"], ops.nchw_to_nhwc_f16(self._check_input(x), 64))
        elif x.is_cuda and self._is_raw_frames(x) and x.dim() == 4 and x.shape[3] == 3:
            t = ops.first_conv3x3(x.contiguous(), P["first"].w, P["first"].b, u8_norm=self.input_normalization)
        else:
            x = self._check_input(x)
            t = ops.first_conv3x3(x, P["first"].w, P["first"].b)  # gather + pack + conv + bias + ReLU fused
        skips = {}
        sk = self.skip_connections
        pooled = None
        for bi, (block, idxs, _) in enumerate(VGG_TRUNK):
            if bi > 0:
                t = pooled if pooled is not None else ops.maxpool(t, 2, 2, 0)
                pooled = None
                skips["pool%d" % bi] = t
            for j in idxs:
                if block == "layer_0_1_down" and j == 0:
                    continue
                pc = P["%s.%d" % (block, j)]
                last = j == idxs[-1] and bi < len(VGG_TRUNK) - 1
                if last and ops.pool_fusion_pays(t.shape[1], t.shape[2]):
                    # nn.MaxPool2d(2) fused into this conv's epilogue; the un-pooled map is only kept
                    # when a skip connection needs it (layer_0_1_down output, models.py:807)
                    keep = sk and block == "layer_0_1_down"
                    B, H, W, _c = t.shape
                    full, pooled = ops.conv_taps(t, pc.w, pc.b, pc.taps, H, W, relu=pc.relu,
                                                 pool="both" if keep else "only")
                    t = full
                else:
                    t = _run_conv(pc, t)
            skips[block] = t
        if sk:
            t = ops.add_(t.clone(), skips["pool4"])
        if self.deconv_decoder:
            t = _run_conv(P["deconv_0_4.2"], _run_deconv(P["deconv_0_4.0"], t))
            if sk:
                t = ops.add_(t, skips["pool3"])
            t = _run_conv(P["deconv_0_3.2"], _run_deconv(P["deconv_0_3.0"], t))
            if sk:
                t = ops.add_(t, skips["pool2"])
            t = _run_conv(P["deconv_0_2.2"], _run_deconv(P["deconv_0_2.0"], t))
            if sk:
                t = ops.add_(t, skips["pool1"])
            t = _run_deconv(P["deconv_0_1.0"], t)
            if sk:
                t = ops.add_(t, skips["layer_0_1_down"])
        else:
            # upsample x2 + conv folded into 4 phase convs on the low-res tensor (_pack_upsampled_conv)
            t = _run_conv(P["upsample_0_4.6"], _run_deconv(P["upsample_0_4.4/up"], t))
            if sk:
                t = ops.add_(t, skips["pool3"])
            t = _run_conv(P["upsample_0_3.6"], _run_deconv(P["upsample_0_3.4/up"], t))
            if self.full_output:
                for name in ("upsample_0_2", "upsample_0_1"):
                    t = _run_conv(P[name + ".4"], _run_deconv(P[name + ".2/up"], t))
        t = _run_conv(P["heads_0.0"], t)
        t = _run_conv(P["heads_0.2"], t)
        return _run_conv(P["heads_0.4"], t, head_cout=self.n_keypoints)

    def forward(self, x):
        # like a torch module the network is differentiable whenever autograd is recording, in train() AND eval()
        # mode (a vgg network has no mode-dependent layer; the reference's loops call `loss.backward()` after
        # `enable_training()`, but nothing in torch would stop them in eval mode either)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd import hourglass_train_forward
            out = hourglass_train_forward(self, x)
        else:
            out = self.belief_maps(x)
        outputs = [out]
        if self.internalize_spatial_softmax:
            outputs.append(self.softmax(out))
        return outputs


class DreamHourglassMultiStage(nn.Module):
    """dream/models.py:350-553: up to 6 DreamHourglass stages; stage s>1 sees the image concatenated with the
    previous stage's belief maps (nearest x4 unless the stage already outputs full resolution, :487-493) and
    `forward` returns the list of every stage's belief maps.  Same constructor keywords and `stageN.*` state-dict
    names.  Each stage runs the DreamHourglass kernel plan; the concat / x4 replication between stages are
    plain tensor ops, and in training each stage's autograd node also returns the gradient of its input so the
    loss on stage s reaches stages < s."""

    def __init__(self, n_keypoints, n_image_input_channels=3, internalize_spatial_softmax=True, learned_beta=True,
                 initial_beta=1.0, n_stages=2, skip_connections=False, deconv_decoder=False, full_output=False):
        super().__init__()
        self.n_keypoints = n_keypoints
        self.n_image_input_channels = n_image_input_channels
        self.internalize_spatial_softmax = internalize_spatial_softmax
        self.skip_connections = skip_connections
        self.deconv_decoder = deconv_decoder
        self.full_output = full_output
        if internalize_spatial_softmax:
            print("WARNING: Keypoint softmax output head is currently unused. Prefer training new models of this "
                  "type with internalize_spatial_softmax = False.")
            self.n_output_heads = 2
            self.learned_beta = learned_beta
            self.initial_beta = initial_beta
        else:
            self.n_output_heads = 1
            self.learned_beta = False
        assert isinstance(n_stages, int), \
            'Expected "n_stages" to be an integer, but it is {}.'.format(type(n_stages))
        assert 0 < n_stages and n_stages <= 6, \
            "DreamHourglassMultiStage can only be constructed with 1 to 6 stages at this time."
        self.num_stages = n_stages
        for s in range(1, n_stages + 1):
            n_in = n_image_input_channels if s == 1 else n_image_input_channels + n_keypoints
            setattr(self, "stage%d" % s,
                    DreamHourglass(n_keypoints, n_in, internalize_spatial_softmax, learned_beta, initial_beta,
                                   skip_connections=skip_connections, deconv_decoder=deconv_decoder,
                                   full_output=full_output))

    input_normalization = _PlanModule.input_normalization

    def forward(self, x, verbose=False):
        stage1 = self.stage1
        if x.is_cuda and x.dtype == torch.uint8:
            stage1.input_normalization = self.input_normalization
            x = stage1._check_input(x)                       # raw frames -> normalised fp32 NCHW, once for all stages
        outputs = []
        y = stage1(x)[0]                                     # "just keeping belief maps for now" (:477)
        outputs.append(y)
        for s in range(2, self.num_stages + 1):
            if self.deconv_decoder or self.full_output:
                y_up = y
            else:
                y_up = nn.functional.interpolate(y, scale_factor=4)
            y = getattr(self, "stage%d" % s)(torch.cat([x, y_up], dim=1))[0]
            outputs.append(y)
        return outputs


# ----------------------------------------------------------------------------------------------
# ResnetSimple (resnet-H / resnet-F)
# ----------------------------------------------------------------------------------------------
class ResnetSimple(_PlanModule):
    """dream/models.py:17-155: ResNet-101 trunk + 4 (H) or 5 (F) ConvTranspose(4,2,1)+BN+ReLU + 1x1 head."""

    def __init__(self, n_keypoints=7, freeze=False, pretrained=True, full=False):
        super().__init__()
        self.full = full
        self.n_keypoints = n_keypoints
        _add_conv(self, "conv1", 3, 64, 7, bias=False)
        _add_bn(self, "bn1", 64)
        inplanes = 64
        for li, nblocks in enumerate(RESNET101_BLOCKS, start=1):
            planes = 64 * 2 ** (li - 1)
            for bi in range(nblocks):
                k = "layer%d.%d" % (li, bi)
                _add_conv(self, k + ".conv1", inplanes, planes, 1, bias=False); _add_bn(self, k + ".bn1", planes)
                _add_conv(self, k + ".conv2", planes, planes, 3, bias=False); _add_bn(self, k + ".bn2", planes)
                _add_conv(self, k + ".conv3", planes, planes * 4, 1, bias=False); _add_bn(self, k + ".bn3", planes * 4)
                if bi == 0:
                    _add_conv(self, k + ".downsample.0", inplanes, planes * 4, 1, bias=False)
                    _add_bn(self, k + ".downsample.1", planes * 4)
                inplanes = planes * 4
        cin = 2048
        for i in range(4):
            _add_conv(self, "upsample.%d" % (3 * i), cin, 256, 4, transposed=True)
            _add_bn(self, "upsample.%d" % (3 * i + 1), 256)
            cin = 256
        if full:
            _add_conv(self, "upsample2.0", 256, 256, 4, transposed=True)
            _add_bn(self, "upsample2.1", 256)
            _add_conv(self, "upsample2.3", 256, n_keypoints, 1)
        else:
            _add_conv(self, "upsample.12", 256, n_keypoints, 1)
        # `freeze` is accepted and ignored exactly like the reference's (models.py:18-21 never reads it);
        # `pretrained=True` takes torchvision's resnet101 weights from a LOCAL checkpoint (dream_b200/pretrained.py)
        if pretrained:
            from . import pretrained as _pretrained
            _pretrained.load_resnet101_trunk(self)

    def _n(self, key):
        return _node_for(self, key)

    def _build_plan(self):
        P = {}
        c1, b1 = self._n("conv1"), self._n("bn1")
        scale, shift = _bn_scale_shift(b1, None)
        wf = ops.pack_first_weight(c1.weight, 192, scale=scale)
        P["conv1"] = _PackedConv(wf, ops.pad_bias(shift, 64, wf.device), [(0, 0)], relu=True, cout=64)
        for li, nblocks in enumerate(RESNET101_BLOCKS, start=1):
            for bi in range(nblocks):
                k = "layer%d.%d" % (li, bi)
                stride = 2 if (li > 1 and bi == 0) else 1
                P[k + ".conv1"] = _pack3x3(self._n(k + ".conv1"), relu=True, bn=self._n(k + ".bn1"))
                P[k + ".conv2"] = _pack3x3(self._n(k + ".conv2"), relu=True, bn=self._n(k + ".bn2"), stride=stride)
                P[k + ".conv3"] = _pack3x3(self._n(k + ".conv3"), relu=True, bn=self._n(k + ".bn3"))
                if bi == 0:
                    P[k + ".down"] = _pack3x3(self._n(k + ".downsample.0"), relu=False,
                                              bn=self._n(k + ".downsample.1"), stride=stride)
        for i in range(4):
            P["up%d" % i] = _pack_deconv(self._n("upsample.%d" % (3 * i)), 4, relu=True,
                                         bn=self._n("upsample.%d" % (3 * i + 1)))
        if self.full:
            P["up4"] = _pack_deconv(self._n("upsample2.0"), 4, relu=True, bn=self._n("upsample2.1"))
            P["head"] = _pack3x3(self._n("upsample2.3"), relu=False, cout_pad=16)
        else:
            P["head"] = _pack3x3(self._n("upsample.12"), relu=False, cout_pad=16)
        return P

    def belief_maps(self, x):
        if self.precise:
            return self.belief_maps_precise(x)
        x = self._check_input(x)
        P = self.plan()
        t = ops.im2col_first(x, 7, 7, 2, 3, 192)
        t = _run_conv(P["conv1"], t)
        t = ops.maxpool(t, 3, 2, 1)
        # The identity stream of the 33 bottlenecks stays fp32 (t32) next to the fp16 copy (t) that feeds
        # the tensor cores: rounding the trunk to fp16 at every block random-walks past the 1e-3 gate.
        # (Tried in round 2 and dropped: running each stage on batch chunks sized for the 126 MB L2 so that the fp32
        #  stream stays on chip.  Same-box: resnet-H 12.9 -> 16.6 ms (120 MB chunks) / 22.4 ms (40 MB), resnet-F
        #  9.9 -> 14.8 ms -- 3-8x more launches with too few tiles each cost far more than the saved HBM traffic;
        #  profiles/r02_resnet_l2_blocking.txt.)
        t32 = None
        for li, nblocks in enumerate(RESNET101_BLOCKS, start=1):
            for bi in range(nblocks):
                k = "layer%d.%d" % (li, bi)
                idn32 = _run_conv(P[k + ".down"], t, want_f32=True)[1] if bi == 0 else t32
                o = _run_conv(P[k + ".conv1"], t)
                o = _run_conv(P[k + ".conv2"], o)
                if bi == nblocks - 1:
                    # last block of a stage: only the fp16 tensor is read again (the next stage's identity comes from
                    # its own downsample conv), so no fp32 copy is written
                    t = _run_conv(P[k + ".conv3"], o, residual_f32=idn32)
                else:
                    t, t32 = _run_conv(P[k + ".conv3"], o, residual_f32=idn32, want_f32=True)  # relu(bn3+identity)
        for i in range(4):
            t = _run_deconv(P["up%d" % i], t)
        if self.full:
            t = _run_deconv(P["up4"], t)
        return _run_conv(P["head"], t, head_cout=self.n_keypoints)

    # ------------------------------------------------------------------------------------------
    # precise mode: split-fp16 operands (x = hi + lo, w = hi + lo; three of the four products)
    # ------------------------------------------------------------------------------------------
    # Tensor-core operands carry 11 significant bits (fp16, and TF32 alike).  On weights that keep every one of the
    # 104 layers O(1) that format alone puts the belief maps 1.1e-3 .. 1.7e-3 from the fp32 reference (the oracle run
    # with nothing but fp16-rounded operands is that far off, tests/test_gpu_parity.py), i.e. the default path sits at
    # the format's floor but outside BASELINE's 1e-3.  `precise = True` (or DREAMB200_PRECISE=1) buys the bits back
    # with the SAME kernels: every activation is kept in fp32 and split into hi = fp16(a), lo = fp16(a - hi); the conv
    # runs over 3*Cin channels [a_hi | a_lo | a_hi] against [w_hi | w_hi | w_lo], i.e. a_hi w_hi + a_lo w_hi + a_hi w_lo
    # (the dropped lo*lo term is 2^-22), accumulating in fp32 in TMEM and leaving through the fp32 epilogue output.
    # 3x the MMA work and fp32 activation traffic: a parity / validation mode, not the benchmarked path.
    precise = __import__("os").environ.get("DREAMB200_PRECISE", "0") == "1"

    @staticmethod
    def _split3(t32):
        """fp32 NHWC [..., C] -> fp16 [..., 3C] = [hi | lo | hi]"""
        hi = t32.half()
        lo = (t32 - hi.float()).half()
        return torch.cat([hi, lo, hi], dim=-1).contiguous()

    @staticmethod
    def _precise_pack(w_oihw, rs, scale, cin_pad=None, cout_pad=None, first=False):
        wf = w_oihw.detach().float()
        if scale is not None:
            wf = wf * scale.view(-1, 1, 1, 1)
        hi = wf.half().float()
        lo = wf - hi
        if first:                                         # 7x7 stem: im2col'ed patches, k = (r*S+s)*3+c padded to 192
            ph, pl = ops.pack_first_weight(hi, 192, cout_pad=cout_pad), ops.pack_first_weight(lo, 192, cout_pad=cout_pad)
        else:
            ph = ops.pack_conv_weight(hi, rs, cin_pad=cin_pad, cout_pad=cout_pad)
            pl = ops.pack_conv_weight(lo, rs, cin_pad=cin_pad, cout_pad=cout_pad)
        return torch.cat([ph, ph, pl], dim=2).contiguous()

    def _precise_plan(self):
        key = self._version_key()
        if getattr(self, "_pp", None) is not None and self._pp_key == key:
            return self._pp
        n = self._n
        P = {}
        with torch.no_grad():
            def conv(key_, bn_key, relu, stride=1, cout_pad=None):
                node, bn = n(key_), (n(bn_key) if bn_key else None)
                scale, bias = (None, node.bias) if bn is None else _bn_scale_shift(bn, node.bias)
                ksz = node.weight.shape[2]
                rs = [(r, s_) for r in range(ksz) for s_ in range(ksz)]
                w3 = self._precise_pack(node.weight, rs, scale, cout_pad=cout_pad)
                taps = [(r - ksz // 2, s_ - ksz // 2) for r, s_ in rs]
                return _PackedConv(w3, ops.pad_bias(bias, w3.shape[1], w3.device), taps, stride=stride, relu=relu,
                                   cout=node.weight.shape[0])

            def deconv(key_, bn_key):
                node, bn = n(key_), n(bn_key)
                scale, bias = _bn_scale_shift(bn, node.bias)
                w_oihw = node.weight.detach().permute(1, 0, 2, 3)
                phases = []
                for py in range(2):
                    for px in range(2):
                        ty, tx = _deconv_phase_taps(4, py), _deconv_phase_taps(4, px)
                        rs = [(ky, kx) for (_, ky) in ty for (_, kx) in tx]
                        taps = [(dy, dx) for (dy, _) in ty for (dx, _) in tx]
                        phases.append((py, px, self._precise_pack(w_oihw, rs, scale), taps))
                cout = w_oihw.shape[0]
                return _PackedConv(None, ops.pad_bias(bias, ops.round_up(cout, 64), node.weight.device), None, relu=True,
                                   cout=cout, phases=phases)
            scale, shift = _bn_scale_shift(n("bn1"), None)
            w3 = self._precise_pack(n("conv1").weight, None, scale, first=True)
            P["conv1"] = _PackedConv(w3, ops.pad_bias(shift, 64, w3.device), [(0, 0)], relu=True, cout=64)
            for li, nblocks in enumerate(RESNET101_BLOCKS, start=1):
                for bi in range(nblocks):
                    k = "layer%d.%d" % (li, bi)
                    stride = 2 if (li > 1 and bi == 0) else 1
                    P[k + ".conv1"] = conv(k + ".conv1", k + ".bn1", True)
                    P[k + ".conv2"] = conv(k + ".conv2", k + ".bn2", True, stride=stride)
                    P[k + ".conv3"] = conv(k + ".conv3", k + ".bn3", True)
                    if bi == 0:
                        P[k + ".down"] = conv(k + ".downsample.0", k + ".downsample.1", False, stride=stride)
            for i in range(4):
                P["up%d" % i] = deconv("upsample.%d" % (3 * i), "upsample.%d" % (3 * i + 1))
            if self.full:
                P["up4"] = deconv("upsample2.0", "upsample2.1")
                P["head"] = conv("upsample2.3", None, False, cout_pad=16)
            else:
                P["head"] = conv("upsample.12", None, False, cout_pad=16)
        self._pp, self._pp_key = P, key
        return P

    def belief_maps_precise(self, x):
        x = self._check_input(x)
        P = self._precise_plan()

        def run(pc, t32, residual_f32=None):
            B, H, W, _ = t32.shape
            Ho, Wo = (H, W) if pc.stride == 1 else ((H - 1) // 2 + 1, (W - 1) // 2 + 1)
            y32 = torch.empty((B, Ho, Wo, pc.w.shape[1]), dtype=torch.float32, device=t32.device)
            ops.conv_taps(self._split3(t32), pc.w, pc.b, pc.taps, Ho, Wo, stride=pc.stride, relu=pc.relu,
                          residual_f32=residual_f32, y_f32=y32)
            return y32

        def run_deconv(pc, t32):
            B, H, W, _ = t32.shape
            cpad = pc.b.numel()
            x3 = self._split3(t32)
            out = torch.empty((B, 2 * H, 2 * W, cpad), dtype=torch.float32, device=t32.device)
            for py, px, w3, taps in pc.phases:
                y32 = torch.empty((B, H, W, cpad), dtype=torch.float32, device=t32.device)
                ops.conv_taps(x3, w3, pc.b, taps, H, W, relu=pc.relu, y_f32=y32)
                out[:, py::2, px::2] = y32
            return out
        # stem: fp16 patches of x_hi and x_lo, 3 x 192 channels
        x_hi = x.half().float()
        p_hi, p_lo = ops.im2col_first(x_hi, 7, 7, 2, 3, 192), ops.im2col_first(x - x_hi, 7, 7, 2, 3, 192)
        pc = P["conv1"]
        B, Ho, Wo, _ = p_hi.shape
        t32 = torch.empty((B, Ho, Wo, 64), dtype=torch.float32, device=x.device)
        ops.conv_taps(torch.cat([p_hi, p_lo, p_hi], dim=-1).contiguous(), pc.w, pc.b, pc.taps, Ho, Wo, relu=True, y_f32=t32)
        t32 = nn.functional.max_pool2d(t32.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()
        for li, nblocks in enumerate(RESNET101_BLOCKS, start=1):
            for bi in range(nblocks):
                k = "layer%d.%d" % (li, bi)
                idn = run(P[k + ".down"], t32) if bi == 0 else t32
                o = run(P[k + ".conv2"], run(P[k + ".conv1"], t32))
                t32 = run(P[k + ".conv3"], o, residual_f32=idn)
        for i in range(4):
            t32 = run_deconv(P["up%d" % i], t32)
        if self.full:
            t32 = run_deconv(P["up4"], t32)
        pc = P["head"]
        return ops.conv_taps(self._split3(t32), pc.w, pc.b, pc.taps, t32.shape[1], t32.shape[2],
                             head_cout=self.n_keypoints)

    def forward(self, x):
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            from .autograd_resnet import resnet_train_forward
            return [resnet_train_forward(self, x)]
        # (train() mode under torch.no_grad(), e.g. the constructor's shape probe, uses the running statistics)
        if torch.is_grad_enabled() and not self.training and any(p.requires_grad for p in self.parameters()) \
                and getattr(self, "strict_eval_autograd", True):
            raise NotImplementedError(
                "dream_b200.ResnetSimple: autograd through eval-mode BatchNorm (running statistics) is not implemented; "
                "call enable_training() / model.train() before differentiating, or wrap inference in torch.no_grad()")
        return [self.belief_maps(x)]


class DataParallelShim(nn.Module):
    """Gives parameters the `module.` prefix torch.nn.DataParallel puts into the reference's
    state dicts (dream/network.py:244-256,616) without DataParallel's per-forward replicate /
    scatter / gather: multi-GPU is process-per-GPU in dream_b200 (see dream_b200/distributed.py)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
