"""Timing experiment (not a test): the split-K fc forward GEMM (512 x 64 x 67,456, segmented A) alone, operands evicted
from L2 by a read-only sweep before every launch, with the clock counters of the MMA thread of CTA (0,0,0).
usage: python tests/manual/fc_fwd_roles.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from curla_b200 import _lib
from helpers import stream

B = 512
S, Ho, pitch = 38 * 68, 31, 68
Kfc, seg_len, seg_stride = Ho * pitch * 32, Ho * pitch * 8, S * 8
dev = 'cuda'
act = torch.randn(B, S * 32, device=dev).to(torch.bfloat16)
Wfc = torch.randn(64, Kfc, device=dev).to(torch.bfloat16)
other = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
P = _lib.ptr
lib = _lib.load()
os.environ['CURLA_GEMM_STAMPS'] = '1'
for env in ({}, {'CURLA_FC_SPLITS_X': '37'}):
    for k in ('CURLA_FC_KSUB', 'CURLA_FC_STAGES'):
        os.environ.pop(k, None)
    os.environ['CURLA_GEMM_STAMPS'] = '1'
    os.environ.update({k: v for k, v in env.items() if not k.endswith('_X')})
    sp = lib.curla_gemm_effective_splits(Kfc, int(env.get('CURLA_FC_SPLITS_X', 74)))
    part = torch.zeros(sp, B, 64, device=dev)

    def run():
        _lib.call('curla_gemm_bf16_seg', P(act), S * 32, P(Wfc), Kfc, P(part), 64, B, 64, Kfc, 3, 64, 0, None, 0, None, 0, sp, B * 64, 1.0,
                  seg_len, seg_stride, 1, stream())
    run()
    ts = []
    for it in range(6):
        other.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    buf = (C.c_longlong * 8)()
    _lib.call('curla_gemm_tc_debug_read', buf)
    d = list(buf)
    print('%-52s splits %3d: %5.1f us (min %5.1f) | CTA 0: kernel %6d clk, K loop %6d clk over %2d steps (wait full %6d, issue %5d)' % (
        env or 'default', sp, sorted(ts)[len(ts) // 2], min(ts), d[6], d[4], d[5], d[0], d[1]), flush=True)
    n = min(4 * sp, 1024)
    sb = (C.c_longlong * (8 * n))()
    _lib.call('curla_gemm_tc_stamps_read', sb, n)
    st = np.array(list(sb), dtype=np.int64).reshape(n, 8)
    t0 = st[:, 0].min()
    names = ['kernel entry', 'after dependency wait', 'first stage full', 'K loop issued', 'accumulator complete', 'epilogue stores issued', 'epilogue: tile in smem']
    for i, nm in enumerate(names):
        v = (st[:, i] - t0) / 1e3
        print('      %-24s min %6.2f  median %6.2f  max %6.2f us after the first CTA entry' % (nm, v.min(), np.median(v), v.max()))
    os.environ['CURLA_GEMM_STAMPS_CLK'] = '1'
    _lib.call('curla_gemm_tc_stamps_read', sb, n)
    os.environ.pop('CURLA_GEMM_STAMPS_CLK')
    ck = np.array(list(sb), dtype=np.int64).reshape(n, 8)
    print('      epilogue warp 2, SM clocks: phase 1 median %d (max %d), phase 2 median %d (max %d)' % (
        np.median(ck[:, 2]), ck[:, 2].max(), np.median(ck[:, 3]), ck[:, 3].max()))
    dur = (st[:, 5] - st[:, 0]) / 1e3
    print('      CTA lifetime median %.2f us; CTAs per SM: max %d; last epilogue - first entry = %.2f us' % (
        np.median(dur), np.bincount(st[:, 7].astype(int)).max(), (st[:, 5].max() - t0) / 1e3))
