"""Timing experiment (not a test): one conv layer in isolation at the bench geometry.
usage: python tests/bench_conv.py [layer(0..3)] [dgrad|wgrad]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from curla_b200 import _lib
from helpers import Geom, pack_conv_w, stream

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dgrad = len(sys.argv) > 2 and sys.argv[2] == 'dgrad'
wgrad = len(sys.argv) > 2 and sys.argv[2] == 'wgrad'
B = 512
g = Geom(76, 135, B)
torch.manual_seed(0)
cin = g.CP1 if layer == 0 else 32
fin, vin = g.alloc(cin)
vin.copy_((torch.rand(vin.shape, device='cuda') * 2).to(torch.bfloat16))
fout, vout = g.alloc(32)
if wgrad:
    vout.copy_((torch.rand(vout.shape, device='cuda') - 0.5).to(torch.bfloat16))
w = torch.randn(32, 9 if layer == 0 else 32, 3, 3) * 0.05
wsh = pack_conv_w(w, layer == 0)
bias = torch.zeros(32, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


ws_buf = torch.zeros(int(max(_lib.load().curla_conv_wgrad_workspace_floats(0), _lib.load().curla_conv_wgrad_workspace_floats(1))),
                     device='cuda')
dW = torch.zeros(32, 9 if layer == 0 else 32, 3, 3, device='cuda')
db = torch.zeros(32, device='cuda')


def run():
    if wgrad:
        _lib.call('curla_conv_wgrad', _lib.ptr(vin), g.S * cin, _lib.ptr(vout), g.S * 32, _lib.ptr(ws_buf), _lib.ptr(dW), _lib.ptr(db),
                  1.0, B, g.pitch, g.S, g.Ho[layer], g.Wo[layer], 9 if layer == 0 else 32, 1 if layer == 0 else 0, stream())
    elif dgrad:
        _lib.call('curla_conv_dgrad', _lib.ptr(vin), g.S * 32, _lib.ptr(wsh), _lib.ptr(vout), _lib.ptr(vout), g.S * 32,
                  B, g.pitch, g.S, g.Ho[layer - 1], g.Wo[layer - 1], stream())
    else:
        _lib.call('curla_conv_fwd', _lib.ptr(vin), g.S * cin, _lib.ptr(wsh), _lib.ptr(bias), 1.0, _lib.ptr(vout),
                  g.S * 32, B, g.pitch, g.S, g.Ho[layer], g.Wo[layer], 1 if layer == 0 else 0, stream())


for _ in range(3):
    run()
ts = []
for _ in range(10):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
if wgrad:
    print('layer %d wgrad (+ reduce) copies=%s stages=%s rows=%s: median %.1f us  min %.1f us' % (
        layer, os.environ.get('CURLA_WG_COPIES', 'default'), os.environ.get('CURLA_WG_STAGES', '2'), os.environ.get('CURLA_WG_ROWS', 'max'),
        ts[len(ts) // 2], ts[0]))
    sys.exit(0)
print('layer %d %s debug=%s stages=%s: median %.1f us  min %.1f us' % (
    layer, 'dgrad' if dgrad else 'fwd', os.environ.get('CURLA_TC_DEBUG', '0'), os.environ.get('CURLA_TC_STAGES', 'max'),
    ts[len(ts) // 2], ts[0]))

if int(os.environ.get('CURLA_TC_DEBUG', '0')) & 64:
    import ctypes as C
    import numpy as np
    buf = (C.c_longlong * (148 * 20))()
    _lib.call('curla_conv_debug_read', buf, 148)
    a = np.array(list(buf), dtype=np.int64).reshape(148, 20)
    if os.environ.get('CURLA_TC_DUMP'):
        np.save(os.environ['CURLA_TC_DUMP'], a)
    names = ['mma wait tempty', 'mma wait full', 'mma issue+sync', 'producer wait empty', 'mma loop total', 'tiles',
             'epi w0 wait tfull', 'epi w0 busy']
    for i, nme in enumerate(names):
        print('   %-22s mean %9.0f  min %9d  max %9d clk' % (nme, a[:, i].mean(), a[:, i].min(), a[:, i].max()))
    if os.environ.get('CURLA_CONV_N96') == '1' and layer > 0:
        for i, nme in ((15, 'epi w0 TMEM phase'), (16, 'epi w0 bar.sync'), (17, 'epi w0 exchange+store')):
            print('   %-22s mean %9.0f  min %9d  max %9d clk' % (nme, a[:, i].mean(), a[:, i].min(), a[:, i].max()))
        sys.exit(0)
    t0 = a[:, 8].min()
    for i, nme in ((8, 'CTA entry'), (9, 'MMA loop start'), (10, 'MMA loop end'), (11, 'CTA exit')):
        v = (a[:, i] - t0) / 1e3
        print('   %-22s mean %8.2f  min %8.2f  max %8.2f us after the first CTA entry' % (nme, v.mean(), v.min(), v.max()))
