"""Shared helpers for the GPU parity tests (layout conversion, error metrics)."""
import ctypes as C

import numpy as np
import torch

from curla_b200 import _lib

DEV = 'cuda'


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def max_rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


class Geom:
    """Mirror of the engine's geometry for an encoder input of H x W."""

    def __init__(self, H, W, B, C=9):
        self.H, self.W, self.B, self.C = H, W, B, C
        self.Hs, self.pitch = (H + 1) // 2, (W + 1) // 2
        self.S = self.Hs * self.pitch
        self.Ho = [(H - 3) // 2 + 1]
        self.Wo = [(W - 3) // 2 + 1]
        for _ in range(3):
            self.Ho.append(self.Ho[-1] - 2)
            self.Wo.append(self.Wo[-1] - 2)
        self.PAD = _lib.load().curla_conv_pad_rows(self.pitch)
        self.CP1 = 48

    def alloc(self, ch, dtype=torch.bfloat16):
        """padded [PAD + B*S + PAD][ch] buffer; returns (full, logical view)."""
        full = torch.zeros((2 * self.PAD + self.B * self.S, ch), dtype=dtype, device=DEV)
        return full, full[self.PAD:self.PAD + self.B * self.S]

    def nhwc(self, view):
        """kernel layout (channel planes [ch/8][S][8] per sample) -> [B][Hs][pitch][ch] view."""
        ch = view.shape[1]
        v = view.reshape(self.B, ch // 8, self.Hs, self.pitch, 8)
        return v.permute(0, 2, 3, 1, 4).reshape(self.B, self.Hs, self.pitch, ch)

    def to_pitch(self, x_nchw, layer):
        """NCHW fp32 (B, 32, Ho, Wo) -> kernel layout bf16 (zeros elsewhere)."""
        B, ch, h, w = x_nchw.shape
        full, view = self.alloc(ch)
        v = view.view(B, ch // 8, self.Hs, self.pitch, 8)
        v[:, :, :h, :w, :] = x_nchw.reshape(B, ch // 8, 8, h, w).permute(0, 1, 3, 4, 2).to(torch.bfloat16)
        return full, view

    def from_pitch(self, view, h, w):
        return self.nhwc(view)[:, :h, :w, :].permute(0, 3, 1, 2).float()

    def s2d_ref(self, x_nchw):
        """reference space-to-depth: (B, C, H, W) -> [B][Hs][Ws][CP1] float."""
        B, Cc, H, W = x_nchw.shape
        xp = torch.zeros((B, Cc, 2 * self.Hs, 2 * self.pitch), dtype=torch.float32, device=x_nchw.device)
        xp[:, :, :H, :W] = x_nchw.float()
        xp = xp.view(B, Cc, self.Hs, 2, self.pitch, 2).permute(0, 2, 4, 1, 3, 5).reshape(
            B, self.Hs, self.pitch, Cc * 4)
        out = torch.zeros((B, self.Hs, self.pitch, self.CP1), dtype=torch.float32, device=x_nchw.device)
        out[..., :Cc * 4] = xp
        return out


def pack_conv_w(w, first, CP1=48):
    """OIHW fp32 -> kernel shadow layout bf16 via the library's own packer."""
    F_, Cin = w.shape[0], w.shape[1]
    src = w.contiguous().float().to(DEV)
    if first:
        dst = torch.zeros((4, F_, CP1), dtype=torch.bfloat16, device=DEV)
        seg = np.array([0, 0, 2, F_, Cin, F_, CP1], dtype=np.int64)
    else:
        dst = torch.zeros((9, F_, Cin), dtype=torch.bfloat16, device=DEV)
        seg = np.array([0, 0, 1, F_, Cin, F_, Cin], dtype=np.int64)
    _lib.call('curla_pack_shadows', _lib.ptr(src), _lib.ptr(dst), seg.ctypes.data_as(C.c_void_p), 1, stream())
    return dst


def bf16r(x):
    return x.to(torch.bfloat16).float()
