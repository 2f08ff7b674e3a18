"""Timing experiment (not a test): where the MAIN stream's time goes in one update with every side stream running.
curla_profile_enable(2) records a CUDA event on the main stream at each phase boundary (waits for joins included) of
eagerly launched updates; prints the mean per update, even (critic + actor + EMA + CPC) and odd (critic + CPC) steps apart.
usage: python tests/manual/phase_times.py [updates per parity, default 10]"""
import contextlib
import ctypes as C
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from curla_b200 import _lib, augmentations, curl_sac, utils

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
device = torch.device('cuda', 0)
torch.cuda.set_device(device)
np.random.seed(12345)
torch.manual_seed(0)
wl = bench.WORKLOADS['curl_crop']
with contextlib.redirect_stdout(io.StringIO()):
    aug = augmentations.make_augmentor(wl['aug'], bench.FRAME[1:])
    rb = utils.ReplayBuffer(bench.FRAME, bench.ACTION, bench.CAPACITY, bench.BATCH, device, aug)
bench.fill_replay(rb)
agent = curl_sac.CurlSacAgent((9, *aug.output_shape), bench.ACTION, device, aug, log_interval=10 ** 9, pixel_sac=False, **bench.HP)
L = bench.NullLogger()
lib = _lib.load()
for i in range(8):
    agent.update(rb, L, i)
torch.cuda.synchronize()
step0 = 8
for parity in (0, 1):
    acc = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall = 0.0
    for i in range(n):
        step = step0 + 2 * i + parity
        lib.curla_profile_enable(2)
        ev0.record()
        agent.update(rb, L, step)
        ev1.record()
        torch.cuda.synchronize()
        wall += ev0.elapsed_time(ev1)
        buf = C.create_string_buffer(1 << 16)
        m = lib.curla_profile_read(buf, len(buf))
        lib.curla_profile_enable(0)
        for line in buf.raw[:max(m, 0)].decode().splitlines():
            nm, cnt, tot = line.split()
            acc[nm] = acc.get(nm, 0.0) + float(tot)
    tot = sum(acc.values())
    print('%s steps: %.3f ms per update on the main stream (eager launches, events around the call: %.3f ms)' % (
        'even (critic, actor, EMA, CPC)' if parity == 0 else 'odd (critic, CPC)', tot / n, wall / n))
    for k in sorted(acc):
        print('   %-52s %8.1f us' % (k, acc[k] / n * 1e3))
