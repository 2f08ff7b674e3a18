"""Where does the host time of one e2e step go?  (ReplayBuffer.add + CurlSacAgent.update + logged scalars)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from curla_b200 import augmentations, curl_sac, utils, engine as E
import contextlib, io
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
np.random.seed(1); torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    aug = augmentations.make_augmentor('random_crop', bench.FRAME[1:])
    rb = utils.ReplayBuffer(bench.FRAME, bench.ACTION, bench.CAPACITY, 512, dev, aug)
bench.fill_replay(rb)
agent = curl_sac.CurlSacAgent((9, *aug.output_shape), bench.ACTION, dev, aug, log_interval=1, **bench.HP)
L = bench.NullLogger()
T = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); T[key] = T.get(key, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, name, g)
wrap(rb, 'add', 'rb.add'); wrap(rb, 'draw_indices', 'draw_indices'); wrap(E.Engine, 'update', 'engine.update (C call)')
sync = torch.cuda.Stream.synchronize
def s2(self):
    t0 = time.perf_counter(); sync(self); T['stream.synchronize (GPU wait)'] = T.get('stream.synchronize (GPU wait)', 0.0) + time.perf_counter() - t0
torch.cuda.Stream.synchronize = s2
host = np.random.RandomState(7)
ob = host.randint(0, 256, size=bench.FRAME, dtype=np.uint8); nb = host.randint(0, 256, size=bench.FRAME, dtype=np.uint8)
act = host.uniform(-1, 1, size=2).astype(np.float32)
for i in range(12):
    rb.add(ob, act, 0.5, nb, False); agent.update(rb, L, i)
torch.cuda.synchronize(); T.clear()
N = 200
t0 = time.perf_counter()
for i in range(N):
    rb.add(ob, act, 0.5, nb, False); agent.update(rb, L, 12 + i)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print('e2e %.1f updates/s, %.1f us per step' % (N / tot, tot / N * 1e6))
for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
    print('  %-34s %8.1f us per step' % (k, v / N * 1e6))
print('  %-34s %8.1f us per step' % ('everything else (python in update)', (tot - sum(T.values())) / N * 1e6))
