"""Timing experiment (not a test): the engine's GEMM call patterns at the bench geometry (B = 512,
crop 76x135, hidden 1024) under the tcgen05 kernel (gemm_tc.cu) and the mma.sync kernel (gemm.cu).
usage: python tests/bench_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from curla_b200 import _lib
from helpers import stream

B, hid = 512, 1024
S, Ho, pitch = 38 * 68, 31, 68
Kfc, seg_len, seg_stride = Ho * pitch * 32, Ho * pitch * 8, S * 8
dev = 'cuda'
rnd = lambda *sh: torch.randn(*sh, device=dev).to(torch.bfloat16)
H = rnd(2, B, hid); W = rnd(2, hid, hid); bias = torch.randn(2, hid, device=dev)
out_bf = torch.zeros(2, B, hid, device=dev, dtype=torch.bfloat16)
out_f = torch.zeros(2, hid, hid, device=dev)
X = rnd(B, 64); dX = torch.zeros(2, B, 64, device=dev); W0 = rnd(2, hid, 64); dW0 = torch.zeros(2, hid, 52, device=dev)
act = rnd(B, S * 32); Wfc = rnd(64, Kfc); dfc = rnd(B, 64); dact = torch.zeros(B, S * 32, device=dev, dtype=torch.bfloat16)
sp = _lib.load().curla_gemm_effective_splits(Kfc, 74)
part = torch.zeros(sp, B, 64, device=dev); dWfc = torch.zeros(64, Kfc, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
P = _lib.ptr
cases = {
    'trunk fwd  512x1024x1024 x2 (bias, relu, bf16)': lambda: _lib.call(
        'curla_gemm_bf16_batched', P(H), hid, P(W), hid, P(out_bf), hid, B, hid, hid, 3, hid, 1, P(bias), 1, None, 0, 1.0, 2,
        B * hid, hid * hid, B * hid, hid, 0, stream()),
    'trunk dgrad 512x1024x1024 x2 (mask, bf16)': lambda: _lib.call(
        'curla_gemm_bf16_batched', P(H), hid, P(W), hid, P(out_bf), hid, B, hid, hid, 1, hid, 1, None, 0, P(H), hid, 1.0, 2,
        B * hid, hid * hid, B * hid, 0, B * hid, stream()),
    'trunk wgrad 1024x1024x512 x2 (fp32)': lambda: _lib.call(
        'curla_gemm_bf16_batched', P(H), hid, P(H), hid, P(out_f), hid, hid, hid, B, 0, hid, 0, None, 0, None, 0, 1.0, 2,
        B * hid, B * hid, hid * hid, 0, 0, stream()),
    'trunk wgrad-in 1024x64x512 x2': lambda: _lib.call(
        'curla_gemm_bf16_batched', P(H), hid, P(X), 64, P(dW0), 52, hid, 64, B, 0, 52, 0, None, 0, None, 0, 1.0, 2,
        B * hid, 0, hid * 52, 0, 0, stream()),
    'trunk dX 512x64x1024 x2': lambda: _lib.call(
        'curla_gemm_bf16_batched', P(H), hid, P(W0), 64, P(dX), 64, B, 64, hid, 1, 64, 0, None, 0, None, 0, 1.0, 2,
        B * hid, hid * 64, B * 64, 0, 0, stream()),
    'fc fwd   512x64x67456 split-K': lambda: _lib.call(
        'curla_gemm_bf16_seg', P(act), S * 32, P(Wfc), Kfc, P(part), 64, B, 64, Kfc, 3, 64, 0, None, 0, None, 0, sp, B * 64, 1.0,
        seg_len, seg_stride, 1, stream()),
    'fc dgrad 512x67456x64 (mask, bf16)': lambda: _lib.call(
        'curla_gemm_bf16_seg', P(dfc), 64, P(Wfc), Kfc, P(dact), S * 32, B, Kfc, 64, 1, Kfc, 1, None, 0, P(act), S * 32, 1, 0, 1.0,
        seg_len, seg_stride, 4, stream()),
    'fc wgrad 50x67456x512 (fp32)': lambda: _lib.call(
        'curla_gemm_bf16_seg', P(dfc), 64, P(act), S * 32, P(dWfc), Kfc, 50, Kfc, B, 0, Kfc, 0, None, 0, None, 0, 1, 0, 1.0,
        seg_len, seg_stride, 2, stream()),
}
for name, fn in cases.items():
    res = []
    for tc in ('2', '0'):
        os.environ['CURLA_GEMM_TC'] = tc
        for _ in range(3):
            fn()
        ts = []
        for _ in range(7):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 4)
        ts.sort()
        res.append(ts[len(ts) // 2])
        if tc == '2':
            import ctypes as C
            buf = (C.c_longlong * 8)()
            _lib.call('curla_gemm_tc_debug_read', buf)
            dbg = list(buf)
    print('%-48s tcgen05 %7.1f us   mma.sync %7.1f us   | MMA thread of CTA 0, clk: wait full %d, issue %d, K loop %d (%d steps), kernel %d'
          % (name, res[0], res[1], dbg[0], dbg[1], dbg[4], dbg[5], dbg[6]))
