"""Torch-tensor wrappers over the C-ABI (include/dreamb200.h).

Tensors are only carriers of device memory here: every op hands raw pointers and the
current CUDA stream to libdreamb200.so.  Activations are NHWC fp16 (`[B,H,W,C]`, C % 64 == 0).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import ConvDesc, check, lib


# Optional per-launch timing (bench.py / tools): when PROFILE is a list, every tensor-core conv
# launch is bracketed by CUDA events on the launching stream and appended as
# (tag, algorithmic_flops, start_event, end_event).  None (default) = zero overhead.
PROFILE = None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def round_up(v, m):
    return (v + m - 1) // m * m


def conv_taps(x, w, bias, taps, Ho, Wo, stride=1, relu=False, residual=None, y=None,
              y_strides=None, head_cout=None, y_offset=0, residual_f32=None, y_f32=None, pool=None, absmax=None,
              gate=None, out_scale=None, colsum=None):
    """Sum-of-shifted-GEMMs convolution on the tensor cores (dreamb200_conv2d_fwd).

    x: [B,H,W,Cin] fp16 contiguous; w: [T,Cout_pad,Cin] fp16; bias fp32 [Cout_pad] or None;
    taps: list of (dy,dx) input offsets.  Returns y [B,Ho,Wo,Cout_pad] fp16, or when
    `head_cout` is given a fp32 NCHW [B,head_cout,Ho,Wo] tensor (Cout_pad must be 16).
    `y`/`y_strides` (w,h,b element strides) let a deconv phase write an interleaved view.
    """
    assert x.is_cuda and x.dtype == torch.float16 and x.is_contiguous() and x.dim() == 4
    assert w.dtype == torch.float16 and w.is_contiguous() and w.dim() == 3
    B, H, W_, Cin = x.shape
    T, Cout_pad, Cin_w = w.shape
    assert Cin_w == Cin and T == len(taps), (w.shape, x.shape, len(taps))
    d = ConvDesc()
    d.x = x.data_ptr(); d.B = B; d.H = H; d.W = W_; d.Cin = Cin; d.in_stride = stride
    d.w = w.data_ptr(); d.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= Cout_pad and bias.is_cuda
    d.taps = T; d.Cout_pad = Cout_pad
    for i, (dy, dx) in enumerate(taps):
        d.tap_dy[i] = dy; d.tap_dx[i] = dx
    d.Ho = Ho; d.Wo = Wo
    if head_cout is not None:
        if y is None:
            y = torch.empty((B, head_cout, Ho, Wo), dtype=torch.float32, device=x.device)
        assert y.dtype == torch.float32 and y.is_contiguous()
        d.out_mode = _lib.OUT_NCHW_F32; d.cout_real = head_cout
        d.y_stride_w = 1; d.y_stride_h = Wo; d.y_stride_b = head_cout * Ho * Wo
    else:
        y_pool = None
        if pool is not None:       # "only": pooled tensor only; "both": un-pooled tensor too (training tape)
            y_pool = torch.empty((B, Ho // 2, Wo // 2, Cout_pad), dtype=torch.float16, device=x.device)
            d.y_pool = y_pool.data_ptr()
        if pool == "only":
            y_strides = (Cout_pad, Wo * Cout_pad, Ho * Wo * Cout_pad)
        elif y is None:
            y = torch.empty((B, Ho, Wo, Cout_pad), dtype=torch.float16, device=x.device)
            y_strides = (Cout_pad, Wo * Cout_pad, Ho * Wo * Cout_pad)
        elif y_strides is None:
            assert y.is_contiguous() and tuple(y.shape) == (B, Ho, Wo, Cout_pad)
            y_strides = (Cout_pad, Wo * Cout_pad, Ho * Wo * Cout_pad)
        d.out_mode = _lib.OUT_NHWC_F16; d.cout_real = Cout_pad
        d.y_stride_w, d.y_stride_h, d.y_stride_b = y_strides
    d.y = (y.data_ptr() + y_offset * y.element_size()) if y is not None else None
    if residual is not None:
        assert residual.dtype == torch.float16 and residual.is_contiguous()
        assert tuple(residual.shape) == (B, Ho, Wo, Cout_pad)
        d.residual = residual.data_ptr()
    d.relu = 1 if relu else 0
    if absmax is not None:          # 1-element fp32 cuda tensor (zeroed): receives max |y| of this launch
        assert absmax.dtype == torch.float32 and absmax.numel() == 1 and head_cout is None
        d.absmax = absmax.data_ptr()
    if gate is not None:            # backward pass: zero the outputs where this fp16 tensor (a ReLU output) is <= 0
        assert gate.dtype == torch.float16 and gate.is_contiguous() and tuple(gate.shape) == (B, Ho, Wo, Cout_pad)
        assert head_cout is None and pool is None
        d.gate = gate.data_ptr()
    if colsum is not None:          # fp32 [Cout_pad] (zeroed): += per-channel sum of the outputs over all pixels
        assert colsum.dtype == torch.float32 and colsum.numel() == Cout_pad and head_cout is None and pool is None
        d.colsum = colsum.data_ptr()
    if out_scale is not None:       # 1-element fp32 cuda tensor multiplied into every output
        assert out_scale.dtype == torch.float32 and out_scale.numel() == 1 and head_cout is None
        d.out_scale = out_scale.data_ptr()
    for name, t in (("residual_f32", residual_f32), ("y_f32", y_f32)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (B, Ho, Wo, Cout_pad)
            setattr(d, name, t.data_ptr())
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().dreamb200_conv2d_fwd(C.byref(d), _stream()), "dreamb200_conv2d_fwd")
        e1.record()
        block_n = 16 if head_cout is not None else (256 if Cout_pad % 256 == 0 else 128 if Cout_pad % 128 == 0 else 64)
        rs = (T == 9 and stride == 1 and block_n in (64, 128) and tuple(taps) == tuple(TAPS_3x3) and
              Ho * Wo / (((Wo + 7) // 8) * ((Ho + 15) // 16) * 128.0) >= float(os.environ.get("DREAMB200_RS_MIN_UTIL", 0.85)))
        if head_cout is not None:      # the head takes the slab kernel whenever it is a plain 3x3 over 64 padded channels
            rs = (T == 9 and stride == 1 and Cin == 64 and tuple(taps) == tuple(TAPS_3x3) and
                  os.environ.get("DREAMB200_RS_HEAD", "1") != "0")
        fam = "conv_rs" if rs else "conv_tc"
        if (not rs and head_cout is None and Cout_pad % 256 == 0 and B * Ho * Wo >= 256 and
                os.environ.get("DREAMB200_TC2", "1")[:1] != "0"):      # mirror of try_conv_tc2 (conv_tc2.cu)
            fam = "conv_tc2"
        if rs and head_cout is None and Ho >= 32:      # mirror of try_conv_rs2 (conv_rs2.cu): the CTA-pair slab kernel
            mode = int(os.environ.get("DREAMB200_RS2", "7"))
            pair_util = Ho * Wo / (((Wo + 7) // 8) * ((Ho + 31) // 32) * 256.0)
            want = (mode & 1) if (Cout_pad == 64 and Cin == 64) else (mode & 4) if Cout_pad == 64 else \
                   ((mode & 8) if Cin == 64 else (mode & 2)) if Cout_pad == 128 else 0
            if want and pair_util >= float(os.environ.get("DREAMB200_RS2_MIN_UTIL", 0.7)):
                fam = "conv_rs2"
        plain = (residual is None and residual_f32 is None and y_f32 is None and gate is None and out_scale is None and
                 colsum is None and absmax is None)
        if (head_cout is None and T == 9 and stride == 1 and tuple(taps) == tuple(TAPS_3x3) and Cout_pad == 64 and
                Cin == 64 and Ho >= 32 and plain and pool != "both"):      # mirror of try_conv_rs3 (conv_rs3.cu)
            mode3 = int(os.environ.get("DREAMB200_RS3", "3"))
            util3 = Ho * Wo / (((Wo + 15) // 16) * ((Ho + 31) // 32) * 512.0)
            if (mode3 & (1 if pool else 2)) and util3 >= 0.65:
                fam = "conv_rs3"
        tag = "%s<%d> T%d Cin%d Cout%d %dx%d s%d" % (fam, block_n, T, Cin, Cout_pad, Ho, Wo, stride)
        PROFILE.append((tag + (" +pool" if pool else ""), 2.0 * B * Ho * Wo * Cout_pad * Cin * T, e0, e1))
    else:
        check(lib().dreamb200_conv2d_fwd(C.byref(d), _stream()), "dreamb200_conv2d_fwd")
    if pool is not None and head_cout is None:
        return (y, y_pool)
    return y


def conv_phases(x, phases, bias, Ho, Wo, y, y_strides, relu=False, stride=1):
    """The sub-pixel phases of a stride-2 ConvTranspose / folded upsample + conv as one call
    (dreamb200_conv2d_fwd_phases): `phases` = [(element offset into y, w [T,Cout_pad,Cin] fp16, taps [(dy,dx)])], all
    writing the same interleaved view (`y_strides`) of `y` from the same input.  One launch when the group is uniform
    (CTA-pair kernel, Cout_pad % 128 == 0), otherwise one launch per phase -- identical results."""
    assert x.is_cuda and x.dtype == torch.float16 and x.is_contiguous() and x.dim() == 4
    B, H, W_, Cin = x.shape
    n = len(phases)
    assert 1 <= n <= 4
    descs = (ConvDesc * n)()
    for i, (y_offset, w, taps) in enumerate(phases):
        assert w.dtype == torch.float16 and w.is_contiguous() and w.dim() == 3 and w.shape[2] == Cin and w.shape[0] == len(taps)
        d = descs[i]
        d.x = x.data_ptr(); d.B = B; d.H = H; d.W = W_; d.Cin = Cin; d.in_stride = stride
        d.w = w.data_ptr(); d.bias = bias.data_ptr() if bias is not None else None
        d.taps = len(taps); d.Cout_pad = w.shape[1]
        for t, (dy, dx) in enumerate(taps):
            d.tap_dy[t] = dy; d.tap_dx[t] = dx
        d.Ho = Ho; d.Wo = Wo
        d.out_mode = _lib.OUT_NHWC_F16; d.cout_real = w.shape[1]
        d.y_stride_w, d.y_stride_h, d.y_stride_b = y_strides
        d.y = y.data_ptr() + y_offset * y.element_size()
        d.relu = 1 if relu else 0
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().dreamb200_conv2d_fwd_phases(descs, n, _stream()), "dreamb200_conv2d_fwd_phases")
    if PROFILE is not None:
        e1.record()
        Cout_pad = phases[0][1].shape[1]
        T = sum(len(p[2]) for p in phases)
        uniform = len({len(p[2]) for p in phases}) == 1 and Cout_pad % 128 == 0 and n > 1 and \
            os.environ.get("DREAMB200_TC2", "1")[:1] != "0" and os.environ.get("DREAMB200_PHASE_GROUPS", "1")[:1] != "0"
        tag = "%s<%d> %dxT%d Cin%d Cout%d %dx%d s%d" % ("conv_tc2" if uniform else "conv_phases", 256 if Cout_pad % 256 == 0 else 128,
                                                        n, T // n, Cin, Cout_pad, Ho, Wo, stride)
        PROFILE.append((tag, 2.0 * B * Ho * Wo * Cout_pad * Cin * T, e0, e1))
    return y


def _norm3(v):
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float32).reshape(-1), (3,)))
    return a


def first_conv3x3(x, w, bias, u8_norm=None):
    """Fused first layer: fp32 NCHW [B,3,H,W] -> relu(conv3x3(x)+bias) as fp16 NHWC [B,H,W,64].
    With `u8_norm=(mean, std)` x is a raw uint8 [B,H,W,3] batch, normalised while it is gathered."""
    assert tuple(w.shape) == (1, 64, 64) and w.dtype == torch.float16 and bias.numel() >= 64
    if u8_norm is not None:
        assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous() and x.shape[3] == 3
        B, H, W_, _ = x.shape
        mean, std = _norm3(u8_norm[0]), _norm3(u8_norm[1])
        y = torch.empty((B, H, W_, 64), dtype=torch.float16, device=x.device)
        check(lib().dreamb200_first_conv3x3_u8(_ptr(x), mean.ctypes.data_as(C.c_void_p),
                                               std.ctypes.data_as(C.c_void_p), _ptr(w), _ptr(bias), _ptr(y),
                                               B, H, W_, _stream()), "dreamb200_first_conv3x3_u8")
        return y
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3
    B, _, H, W_ = x.shape
    y = torch.empty((B, H, W_, 64), dtype=torch.float16, device=x.device)
    e0 = e1 = None
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().dreamb200_first_conv3x3(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), B, H, W_, _stream()),
          "dreamb200_first_conv3x3")
    if PROFILE is not None:
        e1.record()
        PROFILE.append(("first_conv3x3 %dx%d" % (H, W_), 2.0 * B * H * W_ * 64 * 27, e0, e1))
    return y


def im2col_first(x, R, S, stride, pad, Kpad):
    """fp32 NCHW [B,3,H,W] -> fp16 NHWC patches [B,Ho,Wo,Kpad] (k = (r*S+s)*3+c)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3
    B, _, H, W_ = x.shape
    Ho = (H + 2 * pad - R) // stride + 1
    Wo = (W_ + 2 * pad - S) // stride + 1
    out = torch.empty((B, Ho, Wo, Kpad), dtype=torch.float16, device=x.device)
    check(lib().dreamb200_im2col_first(_ptr(x), _ptr(out), B, H, W_, R, S, stride, pad, Ho, Wo, Kpad,
                                       _stream()), "dreamb200_im2col_first")
    return out


def pool_fusion_pays(Ho, Wo, stride=1):
    """Fusing the 2x2 pool into the conv epilogue forces even tile sides; worth it unless that wastes
    noticeably more accumulator rows than the free tile choice (e.g. 50x50 maps)."""
    l = lib()
    return l.dreamb200_conv_tile_utilization(Wo, Ho, stride, 1) >= 0.9 * l.dreamb200_conv_tile_utilization(Wo, Ho, stride, 0)


def maxpool(x, k, s, p):
    B, H, W_, Cc = x.shape
    Ho = (H + 2 * p - k) // s + 1
    Wo = (W_ + 2 * p - k) // s + 1
    y = torch.empty((B, Ho, Wo, Cc), dtype=torch.float16, device=x.device)
    check(lib().dreamb200_maxpool_nhwc(_ptr(x), _ptr(y), B, H, W_, Cc, k, s, p, Ho, Wo, _stream()),
          "dreamb200_maxpool_nhwc")
    return y


def add_(y, x):
    """y += x in place (fp16, same shape)."""
    assert y.dtype == torch.float16 and x.dtype == torch.float16 and y.shape == x.shape
    assert y.is_contiguous() and x.is_contiguous()
    check(lib().dreamb200_add_f16(_ptr(y), _ptr(x), y.numel(), _stream()), "dreamb200_add_f16")
    return y


def upsample2(x):
    B, H, W_, Cc = x.shape
    y = torch.empty((B, 2 * H, 2 * W_, Cc), dtype=torch.float16, device=x.device)
    check(lib().dreamb200_upsample2_nhwc(_ptr(x), _ptr(y), B, H, W_, Cc, _stream()), "dreamb200_upsample2_nhwc")
    return y


def nhwc_to_nchw_f32(x, C_real):
    B, H, W_, Cpad = x.shape
    y = torch.empty((B, C_real, H, W_), dtype=torch.float32, device=x.device)
    check(lib().dreamb200_nhwc_f16_to_nchw_f32(_ptr(x), _ptr(y), B, H, W_, Cpad, C_real, _stream()),
          "dreamb200_nhwc_f16_to_nchw_f32")
    return y


def nchw_to_nhwc_f16(x, Cpad):
    B, Cc, H, W_ = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty((B, H, W_, Cpad), dtype=torch.float16, device=x.device)
    check(lib().dreamb200_nchw_f32_to_nhwc_f16(_ptr(x), _ptr(y), B, H, W_, Cc, Cpad, _stream()),
          "dreamb200_nchw_f32_to_nhwc_f16")
    return y


# ----------------------------------------------------------------------------------------------
# weight packing (load time / once per optimizer step; plain torch ops, not on the hot path)
# ----------------------------------------------------------------------------------------------
TAPS_3x3 = [(r - 1, s - 1) for r in range(3) for s in range(3)]


def pack_conv_weight(w, rs_list, cin_pad=None, cout_pad=None, scale=None):
    """OIHW fp32 conv weight -> fp16 [taps, Cout_pad, Cin_pad]; rs_list selects (r,s) per tap.
    `scale` (per-Cout fp32) folds an eval-mode BatchNorm into the weights."""
    Cout, Cin, R, S = w.shape
    cin_pad = cin_pad or round_up(Cin, 64)
    cout_pad = cout_pad or round_up(Cout, 64)
    wf = w.detach().float()
    if scale is not None:
        wf = wf * scale.view(-1, 1, 1, 1)
    order = [r * S + s for r, s in rs_list]
    flat = wf.reshape(Cout, Cin, R * S)
    if order == list(range(R * S)):                      # all taps in raster order: one permuted copy
        taps = flat.permute(2, 0, 1)
    elif order == list(range(R * S - 1, -1, -1)):        # ... or reversed (data gradient of a 'same' conv)
        taps = flat.flip(2).permute(2, 0, 1)
    else:
        taps = torch.stack([wf[:, :, r, s] for r, s in rs_list], dim=0)
    if Cout == cout_pad and Cin == cin_pad:
        out = torch.empty((len(rs_list), Cout, Cin), dtype=torch.float16, device=w.device)
        out.copy_(taps)                                  # permute + fp32 -> fp16 in one kernel
        return out
    out = torch.zeros((len(rs_list), cout_pad, cin_pad), dtype=torch.float16, device=w.device)
    out[:, :Cout, :Cin] = taps
    return out


def pack_first_weight(w, Kpad, cout_pad=None, scale=None):
    """First-layer weight [Cout,3,R,S] -> [1, Cout_pad, Kpad] matching im2col_first's k order."""
    Cout, Cin, R, S = w.shape
    assert Cin == 3
    cout_pad = cout_pad or round_up(Cout, 64)
    wf = w.detach().float()
    if scale is not None:
        wf = wf * scale.view(-1, 1, 1, 1)
    flat = wf.permute(0, 2, 3, 1).reshape(Cout, R * S * 3)   # k = (r*S+s)*3 + c
    out = torch.zeros((1, cout_pad, Kpad), dtype=torch.float16, device=w.device)
    out[0, :Cout, :R * S * 3] = flat.to(torch.float16)
    return out.contiguous()


def pad_bias(b, cout_pad, device):
    out = torch.zeros((cout_pad,), dtype=torch.float32, device=device)
    if b is not None:
        out[: b.numel()] = b.detach().float()
    return out


# ----------------------------------------------------------------------------------------------
# backward (training) wrappers
# ----------------------------------------------------------------------------------------------
def wgrad(dy, x, taps, deconv=False):
    """dW[tap][co][ci] = sum_p dy[p][co] * x[p + tap][ci]; dy/x NHWC fp16 of equal H, W -> fp32 [T,Co,Ci].
    deconv=True: stride-2 ConvTranspose weight gradient, dy on the 2x finer grid: sum_p dy[2p + tap][co] * x[p][ci]."""
    B, H, W_, Ci = x.shape
    Co = dy.shape[3]
    assert tuple(dy.shape[:3]) == ((B, 2 * H, 2 * W_) if deconv else (B, H, W_))
    assert dy.dtype == torch.float16 and x.dtype == torch.float16 and dy.is_contiguous() and x.is_contiguous()
    dw = torch.zeros((len(taps), Co, Ci), dtype=torch.float32, device=dy.device)
    tdy = (C.c_int8 * len(taps))(*[t[0] for t in taps])
    tdx = (C.c_int8 * len(taps))(*[t[1] for t in taps])
    e0 = e1 = None
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    fn = lib().dreamb200_wgrad_deconv if deconv else lib().dreamb200_wgrad
    check(fn(_ptr(dy), _ptr(x), _ptr(dw), B, H, W_, Co, Ci, len(taps),
             C.cast(tdy, C.c_void_p), C.cast(tdx, C.c_void_p), _stream()), "dreamb200_wgrad")
    if PROFILE is not None:
        e1.record()
        PROFILE.append(("wgrad_tc T%d Cin%d Cout%d %dx%d" % (len(taps), Ci, Co, H, W_),
                        2.0 * B * H * W_ * Co * Ci * len(taps), e0, e1))
    return dw


def scale_mask_(dy, y=None, scale=None):
    """dy = dy * scale * (y > 0) in place; y (fp16, same shape) and scale (1-element fp32 cuda tensor) optional."""
    assert dy.dtype == torch.float16 and dy.is_contiguous()
    if y is not None:
        assert y.shape == dy.shape and y.dtype == torch.float16 and y.is_contiguous()
    check(lib().dreamb200_scale_mask_f16(_ptr(dy), _ptr(y), _ptr(scale), dy.numel(), _stream()),
          "dreamb200_scale_mask_f16")
    return dy


def relu_mask_(dy, y):
    return scale_mask_(dy, y)


def scale_mask_bias_(dy, y=None, scale=None):
    """scale_mask_ fused with the bias gradient: returns db[c] = sum over pixels of the updated dy (fp32 [C])."""
    assert dy.dtype == torch.float16 and dy.is_contiguous()
    Cc = dy.shape[-1]
    db = torch.zeros((Cc,), dtype=torch.float32, device=dy.device)
    check(lib().dreamb200_scale_mask_bias_f16(_ptr(dy), _ptr(y), _ptr(scale), _ptr(db), dy.numel() // Cc, Cc,
                                              _stream()), "dreamb200_scale_mask_bias_f16")
    return db


def wgrad_first_supported(x):
    return x.shape[3] % 4 == 0 and x.data_ptr() % 16 == 0


def wgrad_first(dy, x):
    """Weight gradient of the first 3x3 conv straight from the fp32 NCHW input (no im2col tensor):
    dy fp16 [B,H,W,64], x fp32 [B,3,H,W] -> fp32 [64, 27] with k = (r*3+s)*3+c."""
    B, H, W_, Co = dy.shape
    assert Co == 64 and dy.dtype == torch.float16 and dy.is_contiguous()
    assert x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (B, 3, H, W_)
    dw = torch.zeros((64, 64), dtype=torch.float32, device=dy.device)
    e0 = e1 = None
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().dreamb200_wgrad_first3x3(_ptr(dy), _ptr(x), _ptr(dw), B, H, W_, _stream()), "dreamb200_wgrad_first3x3")
    if PROFILE is not None:
        e1.record()
        PROFILE.append(("wgrad_first3x3 %dx%d" % (H, W_), 2.0 * B * H * W_ * 64 * 27, e0, e1))
    return dw[:, :27]


def loss_scale_step(amax, cum, target=256.0):
    """f = 2^floor(log2(target / amax)) (clamped); cum *= f IN PLACE; returns (f, 1/cum) as 1-element cuda tensors."""
    out = torch.empty((2,), dtype=torch.float32, device=cum.device)
    f, inv = out[0:1], out[1:2]
    check(lib().dreamb200_loss_scale_step(_ptr(amax), _ptr(cum), _ptr(f), _ptr(inv), float(target), _stream()),
          "dreamb200_loss_scale_step")
    return f, inv


def absmax(x):
    """max |x| of an fp16 tensor as a 1-element fp32 cuda tensor (no host sync)."""
    out = torch.zeros((1,), dtype=torch.float32, device=x.device)
    check(lib().dreamb200_absmax_f16(_ptr(x), x.numel(), _ptr(out), _stream()), "dreamb200_absmax_f16")
    return out


def maxpool2_bwd(x, dy, relu_gate=False):
    """Gradient of nn.MaxPool2d(2); with `relu_gate` x is a ReLU output and the gradient also passes that ReLU."""
    B, H, W_, Cc = x.shape
    dx = torch.empty_like(x)
    check(lib().dreamb200_maxpool2_bwd_nhwc(_ptr(x), _ptr(dy), _ptr(dx), B, H, W_, Cc, 1 if relu_gate else 0,
                                            _stream()), "dreamb200_maxpool2_bwd_nhwc")
    return dx


def upsample2_bwd(dy):
    B, H2, W2, Cc = dy.shape
    dx = torch.empty((B, H2 // 2, W2 // 2, Cc), dtype=torch.float16, device=dy.device)
    check(lib().dreamb200_upsample2_bwd_nhwc(_ptr(dy), _ptr(dx), B, H2 // 2, W2 // 2, Cc, _stream()),
          "dreamb200_upsample2_bwd_nhwc")
    return dx


def bias_grad(dy):
    Cc = dy.shape[-1]
    db = torch.zeros((Cc,), dtype=torch.float32, device=dy.device)
    check(lib().dreamb200_bias_grad(_ptr(dy), _ptr(db), dy.numel() // Cc, Cc, _stream()), "dreamb200_bias_grad")
    return db


# ----------------------------------------------------------------------------------------------
# BatchNorm (training mode) and strided backward wrappers
# ----------------------------------------------------------------------------------------------
def wgrad_strided(dy, x, taps):
    """Weight gradient of a stride-2 conv: dW[tap][co][ci] = sum_p dy[p][co] * x[2p + tap][ci]."""
    B, Ho, Wo, Co = dy.shape
    _, Hx, Wx, Ci = x.shape
    assert dy.is_contiguous() and x.is_contiguous() and dy.dtype == torch.float16 and x.dtype == torch.float16
    dw = torch.zeros((len(taps), Co, Ci), dtype=torch.float32, device=dy.device)
    tdy = (C.c_int8 * len(taps))(*[t[0] for t in taps])
    tdx = (C.c_int8 * len(taps))(*[t[1] for t in taps])
    check(lib().dreamb200_wgrad_strided(_ptr(dy), _ptr(x), _ptr(dw), B, Ho, Wo, Hx, Wx, Co, Ci, len(taps),
                                        C.cast(tdy, C.c_void_p), C.cast(tdx, C.c_void_p), _stream()),
          "dreamb200_wgrad_strided")
    return dw


_BN_TICKETS = {}


def _bn_workspace(rows, Cc, device):
    """(partials, tickets) for the deterministic two-stage reductions: scratch partial sums and the persistent,
    always-zero-between-launches ticket counters of this device (dreamb200_bn_reduce_workspace)."""
    nf, nt = C.c_longlong(0), C.c_int(0)
    check(lib().dreamb200_bn_reduce_workspace(rows, Cc, C.byref(nf), C.byref(nt)), "dreamb200_bn_reduce_workspace")
    tickets = _BN_TICKETS.get(device)
    if tickets is None or tickets.numel() < nt.value:
        tickets = _BN_TICKETS[device] = torch.zeros((max(64, nt.value),), dtype=torch.int32, device=device)
    return torch.empty((nf.value,), dtype=torch.float32, device=device), tickets


def bn_stats(z):
    """Per-channel (sum, sum of squares) of an fp16 NHWC tensor, fp32 [C] each (fixed summation order)."""
    Cc = z.shape[-1]
    s = torch.empty((2, Cc), dtype=torch.float32, device=z.device)
    partials, tickets = _bn_workspace(z.numel() // Cc, Cc, z.device)
    check(lib().dreamb200_bn_stats_f16(_ptr(z), _ptr(s[0]), _ptr(s[1]), z.numel() // Cc, Cc, _ptr(partials),
                                       _ptr(tickets), _stream()), "dreamb200_bn_stats_f16")
    return s[0], s[1]


def bn_apply(z, scale, shift, residual=None, relu=False):
    y = torch.empty_like(z)
    Cc = z.shape[-1]
    check(lib().dreamb200_bn_apply_f16(_ptr(z), _ptr(scale), _ptr(shift), _ptr(residual), _ptr(y),
                                       z.numel() // Cc, Cc, 1 if relu else 0, _stream()), "dreamb200_bn_apply_f16")
    return y


def bn_bwd_reduce(dy, z):
    Cc = z.shape[-1]
    s = torch.empty((2, Cc), dtype=torch.float32, device=z.device)
    partials, tickets = _bn_workspace(z.numel() // Cc, Cc, z.device)
    check(lib().dreamb200_bn_bwd_reduce_f16(_ptr(dy), _ptr(z), _ptr(s[0]), _ptr(s[1]), z.numel() // Cc, Cc,
                                            _ptr(partials), _ptr(tickets), _stream()), "dreamb200_bn_bwd_reduce_f16")
    return s[0], s[1]


def bn_bwd_apply_(dy, z, a, b, c0):
    Cc = z.shape[-1]
    check(lib().dreamb200_bn_bwd_apply_f16(_ptr(dy), _ptr(z), _ptr(a), _ptr(b), _ptr(c0), z.numel() // Cc, Cc,
                                           _stream()), "dreamb200_bn_bwd_apply_f16")
    return dy


def maxpool3_bwd(x, dy):
    B, H, W_, Cc = x.shape
    dx = torch.empty_like(x)
    check(lib().dreamb200_maxpool3_bwd_nhwc(_ptr(x), _ptr(dy), _ptr(dx), B, H, W_, Cc, _stream()),
          "dreamb200_maxpool3_bwd_nhwc")
    return dx
