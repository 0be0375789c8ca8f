"""CPU fp32 restatement of the reference CNNs, driven by a reference-format state dict.

Follows (does not copy) dream/models.py:
  * vgg_forward     -- DreamHourglass.forward, models.py:761-827, with the layer lists built in
                       DreamHourglass.__init__, models.py:587-747 (vgg19.features indices 1-3, 5-8,
                       10-17, 19-26, 28-35 => conv keys 0/2, 5/7, 10..16, 19..25, 28..34).
  * resnet_forward  -- ResnetSimple.forward, models.py:140-155, trunk = torchvision resnet101
                       (Bottleneck, stride on the 3x3 conv, BN eps 1e-5), decoder models.py:37-136.
The arithmetic is ATen's (the reference's own third-party dependency, torch unpinned in
requirements.txt:15-16; here torch 2.11.0): conv2d / conv_transpose2d / batch_norm / max_pool2d /
interpolate(nearest).  Everything is functional so autograd gives the reference gradients too.
"""
import zlib

import torch
import torch.nn.functional as F

VGG_TRUNK = [
    ("layer_0_1_down", (0, 2)),
    ("layer_0_2_down", (5, 7)),
    ("layer_0_3_down", (10, 12, 14, 16)),
    ("layer_0_4_down", (19, 21, 23, 25)),
    ("layer_0_5_down", (28, 30, 32, 34)),
]
VGG_TRUNK_CH = {"layer_0_1_down": 64, "layer_0_2_down": 128, "layer_0_3_down": 256,
                "layer_0_4_down": 512, "layer_0_5_down": 512}
RESNET101_BLOCKS = (3, 4, 23, 3)


def _conv(sd, key, x, relu, padding=1, stride=1):
    y = F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=padding)
    return F.relu(y) if relu else y


def _deconv(sd, key, x, stride, padding, output_padding):
    return F.conv_transpose2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride,
                              padding=padding, output_padding=output_padding)


def vgg_forward(sd, x, deconv_decoder=False, full_output=False, skip_connections=False, prefix="module.",
                collect=None):
    """DreamHourglass.forward (models.py:761-827) without the optional soft-argmax head.
    Returns the belief-map tensor [B,K,h,w].  `collect` (dict) receives named intermediates."""
    p = prefix
    feats = []
    t = x
    for i, (block, idxs) in enumerate(VGG_TRUNK):
        if i > 0:
            t = F.max_pool2d(t, 2)                       # self.down_sample, models.py:589
            feats.append(("pool%d" % i, t))
        for j in idxs:
            t = _conv(sd, "%s%s.%d" % (p, block, j), t, relu=True)
        feats.append((block, t))
    f = dict(feats)
    x_0_1, x_0_1_d = f["layer_0_1_down"], f["pool1"]
    x_0_2_d, x_0_3_d, x_0_4_d = f["pool2"], f["pool3"], f["pool4"]
    x_0_5 = f["layer_0_5_down"]
    dec_in = x_0_5 + x_0_4_d if skip_connections else x_0_5

    if deconv_decoder:                                   # models.py:780-801
        def stage(name, t, with_conv=True):
            t = F.relu(_deconv(sd, p + name + ".0", t, 2, 1, 1))
            if with_conv:
                t = _conv(sd, p + name + ".2", t, relu=True)
            return t
        y = stage("deconv_0_4", dec_in)
        y = stage("deconv_0_3", y + x_0_3_d if skip_connections else y)
        y = stage("deconv_0_2", y + x_0_2_d if skip_connections else y)
        y = stage("deconv_0_1", y + x_0_1_d if skip_connections else y, with_conv=False)
        head_in = y + x_0_1 if skip_connections else y
    else:                                                # models.py:803-815
        def up_stage(name, t):
            t = F.interpolate(t, scale_factor=2)         # nn.Upsample default = nearest
            t = _conv(sd, p + name + ".4", t, relu=True)
            return _conv(sd, p + name + ".6", t, relu=False)
        y = up_stage("upsample_0_4", dec_in)
        y = up_stage("upsample_0_3", y + x_0_3_d if skip_connections else y)
        if full_output:
            for name in ("upsample_0_2", "upsample_0_1"):
                y = F.interpolate(y, scale_factor=2)
                y = _conv(sd, p + name + ".2", y, relu=True)
                y = _conv(sd, p + name + ".4", y, relu=True)
        head_in = y
    h = _conv(sd, p + "heads_0.0", head_in, relu=True)
    h = _conv(sd, p + "heads_0.2", h, relu=True)
    out = _conv(sd, p + "heads_0.4", h, relu=False)
    if collect is not None:
        collect.update(f)
        collect["head_in"] = head_in
    return out


def multistage_forward(sd, x, n_stages=2, deconv_decoder=False, full_output=False, skip_connections=False,
                       prefix="module."):
    """DreamHourglassMultiStage.forward (models.py:473-553): returns [y1, ..., yS]; stage s>1 sees
    cat(x, previous belief maps), the latter nearest-upsampled x4 unless the stages already output at the input
    resolution (:483-491)."""
    outs = []
    y = None
    for s in range(1, n_stages + 1):
        if s == 1:
            xin = x
        else:
            y_up = y if (deconv_decoder or full_output) else F.interpolate(y, scale_factor=4)
            xin = torch.cat([x, y_up], dim=1)
        y = vgg_forward(sd, xin, deconv_decoder=deconv_decoder, full_output=full_output,
                        skip_connections=skip_connections, prefix="%sstage%d." % (prefix, s))
        outs.append(y)
    return outs


def multistage_state_dict(n_keypoints, n_stages, gains, deconv_decoder=False, full_output=False, seed=0,
                          mode="default", prefix="module."):
    """Synthetic weights for DreamHourglassMultiStage: stage s is synth_state_dict of a DreamHourglass with
    3 (s=1) or 3+K input channels, seed+s, head gain gains[s-1], re-keyed to `stage<s>.*` (models.py:396-470)."""
    sd = {}
    for s in range(1, n_stages + 1):
        shapes = vgg_state_shapes(n_keypoints, deconv_decoder=deconv_decoder, full_output=full_output,
                                  prefix="module.", n_in=3 if s == 1 else 3 + n_keypoints)
        one = synth_state_dict(shapes, seed=seed + s, out_gain=float(gains[s - 1]), mode=mode)
        for k, v in one.items():
            sd["%sstage%d.%s" % (prefix, s, k[len("module."):])] = v
    return sd


def _bn(sd, key, x, training, momentum=0.1):
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"],
                        sd[key + ".bias"], training=training, momentum=momentum, eps=1e-5)


def resnet_forward(sd, x, full=False, training=False, prefix="module."):
    """ResnetSimple.forward (models.py:140-155).  BN in eval mode unless training=True
    (then running stats in `sd` are updated in place like nn.BatchNorm2d)."""
    p = prefix
    t = F.conv2d(x, sd[p + "conv1.weight"], None, stride=2, padding=3)
    t = F.relu(_bn(sd, p + "bn1", t, training))
    t = F.max_pool2d(t, 3, 2, 1)
    for li, nblocks in enumerate(RESNET101_BLOCKS, start=1):
        for bi in range(nblocks):
            k = "%slayer%d.%d" % (p, li, bi)
            stride = 2 if (li > 1 and bi == 0) else 1
            idn = t
            o = F.relu(_bn(sd, k + ".bn1", F.conv2d(t, sd[k + ".conv1.weight"]), training))
            o = F.relu(_bn(sd, k + ".bn2", F.conv2d(o, sd[k + ".conv2.weight"], stride=stride, padding=1),
                           training))
            o = _bn(sd, k + ".bn3", F.conv2d(o, sd[k + ".conv3.weight"]), training)
            if (k + ".downsample.0.weight") in sd:
                idn = _bn(sd, k + ".downsample.1",
                          F.conv2d(t, sd[k + ".downsample.0.weight"], stride=stride), training)
            t = F.relu(o + idn)
    # decoder: 4 x [ConvT(4,2,1) + BN + ReLU] (+1 in upsample2 when full) + 1x1 conv
    for i in range(4):
        t = _deconv(sd, "%supsample.%d" % (p, 3 * i), t, 2, 1, 0)
        t = F.relu(_bn(sd, "%supsample.%d" % (p, 3 * i + 1), t, training))
    if full:
        t = _deconv(sd, p + "upsample2.0", t, 2, 1, 0)
        t = F.relu(_bn(sd, p + "upsample2.1", t, training))
        return F.conv2d(t, sd[p + "upsample2.3.weight"], sd[p + "upsample2.3.bias"])
    return F.conv2d(t, sd[p + "upsample.12.weight"], sd[p + "upsample.12.bias"])


# ----------------------------------------------------------------------------------------------
# state-dict shapes (reference key names) and deterministic synthetic weights
# ----------------------------------------------------------------------------------------------
def vgg_state_shapes(n_keypoints, deconv_decoder=False, full_output=False, prefix="module.", n_in=3):
    shapes = {}

    def conv(key, cin, cout, k=3):
        shapes[prefix + key + ".weight"] = (cout, cin, k, k)
        shapes[prefix + key + ".bias"] = (cout,)

    def deconv(key, cin, cout, k=3):
        shapes[prefix + key + ".weight"] = (cin, cout, k, k)
        shapes[prefix + key + ".bias"] = (cout,)
    cin = n_in
    for block, idxs in VGG_TRUNK:
        for j in idxs:
            conv("%s.%d" % (block, j), cin, VGG_TRUNK_CH[block])
            cin = VGG_TRUNK_CH[block]
    if deconv_decoder:
        deconv("deconv_0_4.0", 512, 256); conv("deconv_0_4.2", 256, 256)
        deconv("deconv_0_3.0", 256, 128); conv("deconv_0_3.2", 128, 128)
        deconv("deconv_0_2.0", 128, 64); conv("deconv_0_2.2", 64, 64)
        deconv("deconv_0_1.0", 64, 64)
    else:
        conv("upsample_0_4.4", 512, 256); conv("upsample_0_4.6", 256, 256)
        conv("upsample_0_3.4", 256, 128); conv("upsample_0_3.6", 128, 64)
        if full_output:
            for name in ("upsample_0_2", "upsample_0_1"):
                conv(name + ".2", 64, 64); conv(name + ".4", 64, 64)
    conv("heads_0.0", 64, 64); conv("heads_0.2", 64, 32); conv("heads_0.4", 32, n_keypoints)
    return shapes


def resnet_state_shapes(n_keypoints, full=False, prefix="module."):
    shapes = {}

    def bn(key, c):
        shapes[prefix + key + ".weight"] = (c,)
        shapes[prefix + key + ".bias"] = (c,)
        shapes[prefix + key + ".running_mean"] = (c,)
        shapes[prefix + key + ".running_var"] = (c,)
        shapes[prefix + key + ".num_batches_tracked"] = ()
    shapes[prefix + "conv1.weight"] = (64, 3, 7, 7)
    bn("bn1", 64)
    inplanes = 64
    for li, nblocks in enumerate(RESNET101_BLOCKS, start=1):
        planes = 64 * 2 ** (li - 1)
        for bi in range(nblocks):
            k = "layer%d.%d" % (li, bi)
            shapes[prefix + k + ".conv1.weight"] = (planes, inplanes, 1, 1); bn(k + ".bn1", planes)
            shapes[prefix + k + ".conv2.weight"] = (planes, planes, 3, 3); bn(k + ".bn2", planes)
            shapes[prefix + k + ".conv3.weight"] = (planes * 4, planes, 1, 1); bn(k + ".bn3", planes * 4)
            if bi == 0:
                shapes[prefix + k + ".downsample.0.weight"] = (planes * 4, inplanes, 1, 1)
                bn(k + ".downsample.1", planes * 4)
            inplanes = planes * 4
    cin = 2048
    for i in range(4):
        shapes["%supsample.%d.weight" % (prefix, 3 * i)] = (cin, 256, 4, 4)
        shapes["%supsample.%d.bias" % (prefix, 3 * i)] = (256,)
        bn("upsample.%d" % (3 * i + 1), 256)
        cin = 256
    if full:
        shapes[prefix + "upsample2.0.weight"] = (256, 256, 4, 4)
        shapes[prefix + "upsample2.0.bias"] = (256,)
        bn("upsample2.1", 256)
        shapes[prefix + "upsample2.3.weight"] = (n_keypoints, 256, 1, 1)
        shapes[prefix + "upsample2.3.bias"] = (n_keypoints,)
    else:
        shapes[prefix + "upsample.12.weight"] = (n_keypoints, 256, 1, 1)
        shapes[prefix + "upsample.12.bias"] = (n_keypoints,)
    return shapes


def _is_deconv_key(key):
    k = key.split("module.")[-1]
    if k.startswith("deconv_0_") and k.endswith(".0.weight"):
        return True                                   # vgg-F ConvTranspose2d, models.py:620-686
    if k in ("upsample.0.weight", "upsample.3.weight", "upsample.6.weight", "upsample.9.weight",
             "upsample2.0.weight"):
        return True                                   # resnet ConvTranspose2d, models.py:38-132
    return False


def synth_state_dict(shapes, seed=0, out_gain=1.0, mode="he"):
    """Deterministic synthetic weights; each tensor is drawn from its own generator seeded by
    crc32(key)+seed, so both sides of every parity test can regenerate them from key names alone
    (weights are never stored).  `out_gain` scales the last conv (weight and bias) so belief maps reach
    O(1) like a trained network's.

    mode="default": the statistics the reference's own constructors give a vgg network when nothing is
        downloaded: VGG-19 trunk convs kaiming-normal(fan_out, relu) with zero bias
        (torchvision vgg init), every freshly added conv / deconv (models.py:591-599, 618-747) PyTorch's
        default kaiming-uniform(a=sqrt 5) weight and U(+-1/sqrt(fan_in)) bias.
    mode="he": stress weights -- every conv He-scaled (std sqrt(2/fan_in)) so activations stay O(1)
        through all layers and no bias path dominates; the worst case for 11-bit MMA operands.
        BatchNorm: gamma U(0.8,1.2) (x0.25 on bn3 to keep the residual trunk O(1)), beta/mean N(0,0.05^2),
        running_var U(0.75,1.25).
    """
    import math
    sd = {}
    last = [k for k in shapes if k.endswith(".weight") and len(shapes[k]) == 4][-1]
    last_bias = last[:-len("weight")] + "bias"
    for key, shape in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
        k = key.split("module.")[-1]
        vgg_trunk = k.startswith("layer_0_") and not k.startswith("layer_0_1_down.0.")
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.long)
        elif key.endswith("running_var"):
            sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif key.endswith("running_mean"):
            sd[key] = torch.randn(shape, generator=g) * 0.05
        elif len(shape) == 1 and key.endswith(".weight"):      # BN gamma
            gain = 0.25 if key.endswith("bn3.weight") else 1.0
            sd[key] = (torch.rand(shape, generator=g) * 0.4 + 0.8) * gain
        elif len(shape) == 1:                                   # conv bias / BN beta
            wkey = key[:-len("bias")] + "weight"
            if mode == "default" and len(shapes.get(wkey, ())) == 4:
                if vgg_trunk:
                    sd[key] = torch.zeros(shape)
                else:
                    ws = shapes[wkey]
                    bound = 1.0 / math.sqrt(ws[1] * ws[2] * ws[3])
                    sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
            else:
                sd[key] = torch.randn(shape, generator=g) * 0.05
        else:
            if mode == "default":
                if vgg_trunk:
                    std = math.sqrt(2.0 / (shape[0] * shape[2] * shape[3]))          # fan_out
                    sd[key] = torch.randn(shape, generator=g) * std
                else:
                    bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])           # torch's fan_in
                    sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
            else:
                if _is_deconv_key(key):   # [Cin,Cout,k,k]; an output pixel sees k*k/4 taps per input channel
                    fan_in = shape[0] * (shape[2] * shape[3]) / 4.0
                else:
                    fan_in = shape[1] * shape[2] * shape[3]
                sd[key] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    sd[last] = sd[last] * out_gain
    if last_bias in sd:
        sd[last_bias] = sd[last_bias] * out_gain
    return sd


class fp16_operands:
    """Context manager (test infrastructure): evaluates the oracle with every conv / deconv operand
    (activation and weight) rounded to fp16 and fp32 accumulation -- the reference algorithm as seen
    through the tensor cores' 11-bit-significand input format.  Differences between the CUDA path and
    THIS are implementation error; differences between this and the fp32 oracle are the operand format."""

    def __enter__(self):
        self._c, self._d = F.conv2d, F.conv_transpose2d
        c, d = self._c, self._d

        def conv2d(x, w, b=None, stride=1, padding=0):
            return c(x.half().float(), w.half().float(), b, stride=stride, padding=padding)

        def conv_transpose2d(x, w, b=None, stride=1, padding=0, output_padding=0):
            return d(x.half().float(), w.half().float(), b, stride=stride, padding=padding,
                     output_padding=output_padding)
        F.conv2d, F.conv_transpose2d = conv2d, conv_transpose2d
        return self

    def __exit__(self, *exc):
        F.conv2d, F.conv_transpose2d = self._c, self._d
        return False
