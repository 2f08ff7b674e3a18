"""Timing experiment (not a test): the conv weight-gradient kernel at the bench geometry (B = 512, 76 x 135 crop) over
its staging knobs -- CURLA_WG_COPIES (dy copies staged per image row: horizontal taps in N vs on the A side),
CURLA_WG_STAGES (ring depth), CURLA_WG_ROWS (image rows per stage).  L2 flushed between launches.
usage: python tests/manual/wgrad_sweep.py [layers, e.g. 013]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from curla_b200 import _lib
from helpers import Geom, stream

layers = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else '013')]
B = 512
g = Geom(76, 135, B)
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ws_buf = torch.zeros(int(max(_lib.load().curla_conv_wgrad_workspace_floats(0), _lib.load().curla_conv_wgrad_workspace_floats(1))),
                     device='cuda')
db = torch.zeros(32, device='cuda')
nparts = (__import__('ctypes').c_int * 1)()
for layer in layers:
    cin = g.CP1 if layer == 0 else 32
    fin, vin = g.alloc(cin)
    vin.copy_((torch.rand(vin.shape, device='cuda') * 2).to(torch.bfloat16))
    fout, vout = g.alloc(32)
    vout.copy_((torch.rand(vout.shape, device='cuda') - 0.5).to(torch.bfloat16))
    dW = torch.zeros(32, 9 if layer == 0 else 32, 3, 3, device='cuda')
    for copies in ('3', '2', '1'):
        if layer == 0 and copies == '3':
            continue
        for stages in ('2', '3', '4'):
            for rows in ('max',) + (('3', '2') if stages == '4' else ()):
                os.environ['CURLA_WG_COPIES'] = copies
                os.environ['CURLA_WG_STAGES'] = stages
                if rows == 'max':
                    os.environ.pop('CURLA_WG_ROWS', None)
                else:
                    os.environ['CURLA_WG_ROWS'] = rows

                def run():
                    # first stage only (the per-CTA partials): the second stage is one launch for all four layers in the update
                    _lib.call('curla_conv_wgrad_partial', _lib.ptr(vin), g.S * cin, _lib.ptr(vout), g.S * 32, _lib.ptr(ws_buf), B, g.pitch,
                              g.S, g.Ho[layer], g.Wo[layer], 1 if layer == 0 else 0, nparts, stream())
                for _ in range(2):
                    run()
                ts = []
                for _ in range(7):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    run()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                print('layer %d wgrad copies=%s stages=%s rows=%s: median %.1f us  min %.1f us' % (layer, copies, stages, rows, ts[len(ts) // 2], ts[0]),
                      flush=True)
