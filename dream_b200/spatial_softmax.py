"""SoftArgmaxPavlo on the device (reference: dream/spatial_softmax.py:15-95).

Same constructor and `beta` parameter / attribute as the reference module; `forward` is one
launch of dreamb200_softargmax (7x7 average pool, max-shifted exp(beta*x), expected (x, y)).
Inference only: like the reference, `DreamNetwork.loss` refuses configs that train through it
(network.py:361-362).
"""
import ctypes as C

import torch

from ._lib import check, lib


class SoftArgmaxPavlo(torch.nn.Module):
    def __init__(self, n_keypoints=5, learned_beta=False, initial_beta=25.0):
        super().__init__()
        if learned_beta:
            self.beta = torch.nn.Parameter(torch.ones(n_keypoints) * initial_beta)
        else:
            self.register_buffer("_beta_const", torch.ones(n_keypoints) * initial_beta, persistent=False)
            self.beta = self._beta_const

    def forward(self, heatmaps, size_mult=1.0):
        assert heatmaps.is_cuda, "dream_b200 SoftArgmaxPavlo needs a CUDA tensor (no CPU fallback)"
        hm = heatmaps.detach().contiguous().float()
        B, K, H, W = hm.shape
        beta = (self._beta_const if not isinstance(self.beta, torch.nn.Parameter) else self.beta.detach())
        beta = beta.to(hm.device).float().contiguous()
        out = torch.empty((B, K, 2), dtype=torch.float32, device=hm.device)
        scratch = torch.empty_like(hm)
        check(lib().dreamb200_softargmax(C.c_void_p(hm.data_ptr()), C.c_void_p(beta.data_ptr()),
                                         C.c_void_p(out.data_ptr()), B, K, H, W,
                                         C.c_void_p(scratch.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "dreamb200_softargmax")
        return out * size_mult if size_mult != 1.0 else out
