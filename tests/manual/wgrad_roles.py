"""Timing experiment (not a test): per-role clock counters of the row-ring conv weight-gradient kernel (CTA 0) at the
bench geometry, B = 512, one 3x3 layer.  usage: python tests/manual/wgrad_roles.py [layer]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from curla_b200 import _lib
from helpers import Geom, stream

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = 512
g = Geom(76, 135, B)
torch.manual_seed(0)
ws_buf = torch.zeros(int(max(_lib.load().curla_conv_wgrad_workspace_floats(0), _lib.load().curla_conv_wgrad_workspace_floats(1))), device='cuda')
nparts = (C.c_int * 1)()
cin = g.CP1 if layer == 0 else 32
fin, vin = g.alloc(cin)
vin.copy_((torch.rand(vin.shape, device='cuda') * 2).to(torch.bfloat16))
fout, vout = g.alloc(32)
vout.copy_((torch.rand(vout.shape, device='cuda') - 0.5).to(torch.bfloat16))
# stands in for the kernels between two weight-gradient launches of the update: clean (read-only) lines in L2
other = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
os.environ['CURLA_WG_DEBUG'] = '1'
for cfg in ({}, {'CURLA_WG_RS': '1'}, {'CURLA_WG_RS': '3'}, {'CURLA_WG_RS': '4'}, {'CURLA_WG_PRODUCERS': '1'}, {'CURLA_WG_PRODUCERS': '3'},
            {'CURLA_WG_NSD': '6'}, {'CURLA_WG_TMAP': '1'}, {'CURLA_WG_TMAP': '1', 'CURLA_WG_RS': '3'}, {'CURLA_WG_COPIES': '2'},
            {'CURLA_WG_COPIES': '2', 'CURLA_WG_RS': '3'}, {'CURLA_WG_RING': '0'}):
    for k in ('CURLA_WG_PRODUCERS', 'CURLA_WG_NSD', 'CURLA_WG_TMAP', 'CURLA_WG_COPIES', 'CURLA_WG_RING', 'CURLA_WG_RS'):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    ts = []
    for it in range(5):
        other.sum()                   # evicts the operands with clean lines
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call('curla_conv_wgrad_partial', _lib.ptr(vin), g.S * cin, _lib.ptr(vout), g.S * 32, _lib.ptr(ws_buf), B, g.pitch,
                  g.S, g.Ho[layer], g.Wo[layer], 1 if layer == 0 else 0, nparts, stream())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    buf = (C.c_longlong * 16)()
    _lib.call('curla_conv_wgrad_debug_read', buf)
    d = list(buf)
    rows = max(d[12], 1)
    print('%-52s %6.1f us | per image row, clk: MMA warp wait %5.0f issue %5.0f total %5.0f | producer wait %5.0f issue %5.0f '
          'total %5.0f | helper0 wait %5.0f issue %5.0f total %5.0f' % (
              cfg or 'ring default', sorted(ts)[len(ts) // 2], d[0] / rows, d[2] / rows, d[3] / rows, d[4] / rows, d[5] / rows,
              d[6] / rows, d[8] / rows, d[9] / rows, d[10] / rows), flush=True)
