"""Manual GPU check (not collected by pytest): unusual shapes and flags through the public API --
batch 1 / 7 / 4096, capacity 100,000 (train.py default, 26 GB in HBM), pixel_sac, detach_encoder,
color_jiggle, noisy_cover.  python tests/manual/stress_configs.py"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from curla_b200 import augmentations, curl_sac, utils
dev = torch.device('cuda')
class L:
    def __init__(s): s.r = {}
    def log(s, k, v, st): s.r[k] = float(v)
def run(B, cap, aug_name, steps=3, **kw):
    np.random.seed(1); torch.manual_seed(0)
    aug = augmentations.make_augmentor(aug_name, (90, 160))
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        rb = utils.ReplayBuffer((9, 90, 160), (2,), cap, B, dev, aug)
    n = min(cap, 2048)
    rb.obses[:n] = torch.randint(0, 256, (n, 9, 90, 160), device=dev, dtype=torch.uint8)
    rb.next_obses[:n] = torch.randint(0, 256, (n, 9, 90, 160), device=dev, dtype=torch.uint8)
    rb.actions[:n] = torch.rand(n, 2, device=dev) * 2 - 1; rb.rewards[:n] = torch.randn(n, 1, device=dev); rb.not_dones[:n] = 1
    rb.idx, rb.full = n % cap, n == cap
    agent = curl_sac.CurlSacAgent((9, *aug.output_shape), (2,), dev, aug, hidden_dim=1024, log_interval=1, init_temperature=0.1, **kw)
    l = L()
    for s in range(steps):
        agent.update(rb, l, s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps, steps + 4):
        agent.update(rb, l, s)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 4 * 1e3
    ok = all(np.isfinite(v) for v in l.r.values())
    print('B=%d cap=%d aug=%s %s: %.2f ms/update finite=%s critic=%.4f curl=%s mem=%.1f GB' % (
        B, cap, aug_name, kw, ms, ok, l.r.get('train_critic/loss', float('nan')), l.r.get('train/curl_loss'), torch.cuda.max_memory_allocated() / 2**30))
    assert ok
    del agent, rb
    torch.cuda.empty_cache()
run(1, 64, 'random_crop')
run(7, 64, 'random_crop')
run(4096, 4096, 'random_crop')
run(512, 100000, 'random_crop')
run(256, 1024, 'identity', pixel_sac=True)
run(128, 1024, 'color_jiggle')
run(128, 1024, 'noisy_cover')
run(512, 1024, 'random_crop', detach_encoder=True)
print('STRESS OK')
