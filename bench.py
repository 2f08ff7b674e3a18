#!/usr/bin/env python
"""Benchmark of the CurlSacAgent.update hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload: BASELINE.json configs[1] -- CURL random-crop SAC update at train.py defaults
(batch 512 per GPU, 9x90x160 stored frames cropped to 9x76x135, hidden 1024, feature 50,
4x32 conv filters), synthetic replay buffer of 16,384 transitions (4.25 GB of uint8
frames, far larger than the 126 MB L2) resident in HBM, random-init weights.  With N GPUs
the update is data parallel: per-GPU batch stays 512 (weak scaling), the global batch
512*N is what the CURL head sees (all-gathered keys), so N=8 is configs[3]'s batch 4096.
`value` is observations/s through the update over all ranks (= global batch * updates/s;
`updates_per_s` is printed beside it).

One JSON line on stdout (rank 0).  --impl reference times the reference's OWN update (the
unmodified modules shipped under baseline/_ref by build(), imported through oracle/shims) on
the host cores at the same batch / replay capacity for the real step count (kind "reference";
the oracle port only if those files are absent).  --impl reference_cuda is the informational
"PyTorch eager on the B200" baseline of the same unmodified code.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

FRAME = (9, 90, 160)
ACTION = (2,)
CAPACITY = 16384
BATCH = 512
# train.py defaults (train.py:62-104)
HP = dict(hidden_dim=1024, discount=0.99, init_temperature=0.1, alpha_lr=1e-4, alpha_beta=0.5, actor_lr=1e-3,
          actor_beta=0.9, actor_log_std_min=-10, actor_log_std_max=2, actor_update_freq=2, critic_lr=1e-3,
          critic_beta=0.9, critic_tau=0.01, critic_target_update_freq=2, encoder_feature_dim=50,
          encoder_lr=1e-3, encoder_tau=0.05, num_layers=4, num_filters=32, cpc_update_freq=1)


class NullLogger:
    def __init__(self):
        self.last = {}

    def log(self, key, value, step):
        self.last[key] = float(value)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d['hbm_gbs'], d.get('bf16_tflops_sustained', d['bf16_tflops']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1400.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([x.strip() for x in line.split(',')])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v == 'Active':
                    reasons.add(n)
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def fill_replay(rb, seed=1):
    """SURVEY 8(d) synthetic fill, generated on the device in chunks (4.25 GB)."""
    g = torch.Generator(device=rb.device).manual_seed(seed)
    n = rb.capacity
    for s in range(0, n, 1024):
        e = min(n, s + 1024)
        rb.obses[s:e] = torch.randint(0, 256, (e - s, *FRAME), device=rb.device, dtype=torch.uint8, generator=g)
        rb.next_obses[s:e] = torch.randint(0, 256, (e - s, *FRAME), device=rb.device, dtype=torch.uint8, generator=g)
    rb.actions.copy_(torch.rand((n, 2), device=rb.device, generator=g) * 2 - 1)
    rb.rewards.copy_(torch.randn((n, 1), device=rb.device, generator=g))
    rb.not_dones.copy_((torch.rand((n, 1), device=rb.device, generator=g) > 0.01).float())
    rb.idx, rb.full = 0, True


# ------------------------------------------------------------------ algorithmic work
WORKLOADS = {
    # BASELINE.json configs[1] (the metric's configuration) / configs[3] with --gpus 8
    'curl_crop': dict(aug='random_crop', pixel_sac=False,
                      what='CURL random-crop SAC update, train.py defaults'),
    # configs[2]
    'pixel_sac': dict(aug='identity', pixel_sac=True, what='Pixel SAC (--pixel_sac: identity 90x160, no contrastive head)'),
    # configs[0]'s augmentation at the default batch (the reference times it on CPU at a small batch)
    'curl_identity': dict(aug='identity', pixel_sac=False, what='CURL identity-augmentation update (90x160 encoder input)'),
}


def geometry(hw):
    ho, wo = [(hw[0] - 3) // 2 + 1], [(hw[1] - 3) // 2 + 1]
    for _ in range(3):
        ho.append(ho[-1] - 2)
        wo.append(wo[-1] - 2)
    return ho, wo


def run_ours(args, rank, world, device):
    from curla_b200 import _lib, augmentations, curl_sac, utils
    torch.cuda.set_device(device)
    np.random.seed(12345)                    # identical global index stream on every rank
    torch.manual_seed(0)
    wl = WORKLOADS[args.workload]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):           # the reference's make_augmentor prints its choice
        aug = augmentations.make_augmentor(wl['aug'], FRAME[1:])
    Bg = args.batch * world
    with contextlib.redirect_stdout(io.StringIO()):
        rb = utils.ReplayBuffer(FRAME, ACTION, CAPACITY, Bg, device, aug)
    fill_replay(rb)
    agent = curl_sac.CurlSacAgent((9, *aug.output_shape), ACTION, device, aug, log_interval=10 ** 9,
                                  pixel_sac=wl['pixel_sac'], **HP)
    L = NullLogger()
    step0 = 0

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    # set-up, not warm-up: every update variant (step parity x index staging slot) runs eagerly once and is then
    # captured as a CUDA graph (engine.cu); the W warm-up steps and the timed steps below only replay
    for i in range(8):
        agent.update(rb, L, step0 + i)
    step0 += 8
    for i in range(args.warmup):
        agent.update(rb, L, step0 + i)
    step0 += args.warmup
    if step0 % 2:                             # start timed regions on an even step
        agent.update(rb, L, step0)
        step0 += 1

    # ---- device-timed region: K updates, inputs resident in HBM
    sampler = ClockSampler(device.index or 0) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record()
    for i in range(args.steps):
        agent.update(rb, L, step0 + i)
        launches += agent.engine.last_launches()
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.finish() if sampler else None
    step0 += args.steps
    t = torch.tensor([ms], device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(t) / args.steps

    # ---- end to end through the public API with HOST buffers: every step ingests one
    # transition from host numpy (ReplayBuffer.add: pinned staging + H2D), uploads the
    # sampled indices, runs the update and reads the logged scalars back (D2H).
    agent.log_interval = 1
    host = np.random.RandomState(7)
    h_obs = host.randint(0, 256, size=FRAME, dtype=np.uint8)
    h_next = host.randint(0, 256, size=FRAME, dtype=np.uint8)
    h_act = host.uniform(-1, 1, size=2).astype(np.float32)
    for i in range(2):
        rb.add(h_obs, h_act, 0.5, h_next, False)
        agent.update(rb, L, step0 + i)
    step0 += 2
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        rb.add(h_obs, h_act, 0.5, h_next, False)
        agent.update(rb, L, step0 + i)       # logs -> metrics D2H + sync every step
    sync_all()
    e2e_s = time.perf_counter() - t0
    step0 += args.steps
    agent.log_interval = 10 ** 9
    t = torch.tensor([e2e_s], device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_s = float(t)
    h2d = 2 * int(np.prod(FRAME)) + 4 * 4 + 7 * Bg * 8
    d2h = 16 * 4

    # ---- live per-kernel times (CUDA events after every launch) for the roofline entry
    prof = {}
    if rank == 0 or world > 1:
        lib = _lib.load()
        lib.curla_profile_enable(1)
        for i in range(args.prof_steps):
            agent.update(rb, L, step0 + i)
        torch.cuda.synchronize()
        import ctypes as C
        buf = C.create_string_buffer(1 << 16)
        n = lib.curla_profile_read(buf, len(buf))
        lib.curla_profile_enable(0)
        for line in buf.raw[:max(n, 0)].decode().splitlines():
            nm, cnt, tot = line.split()
            prof[nm] = (int(cnt), float(tot))
        step0 += args.prof_steps
    nparams = {k: int(agent.engine.t['grad.' + k].numel()) for k in ('critic', 'actor', 'cpc')}
    return dict(ms_per_step=ms_per_step, launches=launches, clocks=clocks, e2e_s=e2e_s, h2d=h2d, d2h=d2h,
                prof=prof, prof_steps=args.prof_steps, Bg=Bg, last=L.last, obs_hw=tuple(aug.output_shape), nparams=nparams)


def fill_host_replay(rb, seed=1):
    """SURVEY 8(d) synthetic fill of a host-numpy replay (the reference's ReplayBuffer): uniform uint8
    frames (torch.randint per chunk: the values are not the measured thing, the 4.25 GB footprint is)."""
    g = torch.Generator().manual_seed(seed)
    n = rb.capacity
    for s0 in range(0, n, 512):
        e = min(n, s0 + 512)
        rb.obses[s0:e] = torch.randint(0, 256, (e - s0, *FRAME), dtype=torch.uint8, generator=g).numpy()
        rb.next_obses[s0:e] = torch.randint(0, 256, (e - s0, *FRAME), dtype=torch.uint8, generator=g).numpy()
    rs = np.random.RandomState(seed)
    rb.actions[:] = rs.uniform(-1, 1, size=rb.actions.shape).astype(np.float32)
    rb.rewards[:] = rs.standard_normal(size=rb.rewards.shape).astype(np.float32)
    rb.not_dones[:] = (rs.uniform(size=rb.not_dones.shape) > 0.01).astype(np.float32)
    rb.idx, rb.full = 0, True


def reference_updates(batch, steps, warmup, threads, workload='curl_crop', device='cpu', tf32=False):
    """Seconds per update of the reference's OWN CurlSacAgent.update (curl_sac.py:426-451), sample_cpc
    included, at the benchmarked configuration: batch `batch`, host replay of CAPACITY transitions,
    train.py defaults.  kind 'reference' = the unmodified reference modules (/root/reference, or their
    verbatim copies under baseline/_ref/ on the GPU box) imported through the three import shims of
    oracle/shims; kind 'port' = the oracle restatement, only when neither is present.
    Returns (seconds per update, kind)."""
    from oracle import make_golden as MG
    wl = WORKLOADS[workload]
    torch.set_num_threads(threads)
    ref_dir = MG.reference_dir()
    if ref_dir is None:
        if device != 'cpu':
            raise SystemExit('reference modules not found (baseline/_ref): the CUDA baseline needs the reference itself')
        return port_updates(batch, steps, warmup, workload), 'port'
    import contextlib
    import io
    ref_utils, ref_aug, ref_sac, _ = MG.import_reference(ref_dir)
    dev = torch.device(device)
    if dev.type == 'cuda':
        torch.backends.cudnn.allow_tf32 = bool(tf32)
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    with contextlib.redirect_stdout(io.StringIO()):
        aug = ref_aug.make_augmentor(wl['aug'], FRAME[1:])
        rb = ref_utils.ReplayBuffer(FRAME, ACTION, CAPACITY, batch, dev, aug)
    fill_host_replay(rb)
    ref_utils.set_seed_everywhere(0)
    agent = ref_sac.CurlSacAgent((9, *aug.output_shape), ACTION, dev, aug, log_interval=10 ** 9,
                                 pixel_sac=wl['pixel_sac'], **HP)
    L = NullLogger()
    sync = (lambda: torch.cuda.synchronize()) if dev.type == 'cuda' else (lambda: None)
    step = 2                                  # even start; never a multiple of log_interval
    for _ in range(warmup):
        agent.update(rb, L, step)
        step += 1
    if step % 2:
        agent.update(rb, L, step)
        step += 1
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        agent.update(rb, L, step)
        step += 1
    sync()
    return (time.perf_counter() - t0) / steps, 'reference'


def port_updates(batch, steps, warmup, workload='curl_crop'):
    """Fallback when the reference modules are absent: the oracle port, same batch / capacity."""
    from oracle import curla_oracle as O
    wl = WORKLOADS[workload]
    crop = wl['aug'] == 'random_crop'
    ohw = (76, 135) if crop else FRAME[1:]

    class _RB:
        capacity = CAPACITY
    rb = _RB()
    rb.obses = np.empty((CAPACITY, *FRAME), dtype=np.uint8)
    rb.next_obses = np.empty((CAPACITY, *FRAME), dtype=np.uint8)
    rb.actions = np.empty((CAPACITY, 2), dtype=np.float32)
    rb.rewards = np.empty((CAPACITY, 1), dtype=np.float32)
    rb.not_dones = np.empty((CAPACITY, 1), dtype=np.float32)
    fill_host_replay(rb)
    agent = O.OracleAgent((9, *ohw), 2, hidden_dim=HP['hidden_dim'], pixel_sac=wl['pixel_sac'])
    agent.init_random(0)
    np.random.seed(0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        d = O.draw_sample_indices(CAPACITY, 0, True, batch, wl['aug'], FRAME[1:], ohw)
        f = lambda a: torch.from_numpy(a).float()
        if crop:
            obs = f(O.gather_crop(rb.obses, d['idxs'], d['h1_obs'], d['w1_obs'], ohw))
            nxt = f(O.gather_crop(rb.next_obses, d['idxs'], d['h1_next'], d['w1_next'], ohw))
            pos = f(O.gather_crop(rb.obses, d['idxs'], d['h1_pos'], d['w1_pos'], ohw))
        else:
            obs, nxt = f(O.gather(rb.obses, d['idxs'])), f(O.gather(rb.next_obses, d['idxs']))
            pos = obs.clone()
        noise = torch.randn(2, batch, 2)
        agent.update(obs, torch.from_numpy(rb.actions[d['idxs']]), torch.from_numpy(rb.rewards[d['idxs']]), nxt,
                     torch.from_numpy(rb.not_dones[d['idxs']]), pos, 2 + i, noise[0], noise[1])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def latent_sweep(args, rank, local):
    """BASELINE.json configs[4]: encoder-only latent extraction over a synthetic episode set
    (plot_tsne/latent_data.py:63-104), batch sweep.  One JSON line; `value` is the best batch."""
    if rank != 0:
        return
    from curla_b200 import augmentations, curl_sac, latent
    from oracle import curla_oracle as O
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    torch.manual_seed(0)
    N = 20000                                    # latent_episodes.py:189 collects 20,000 observations
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        aug = augmentations.make_augmentor('random_crop', FRAME[1:])
    agent = curl_sac.CurlSacAgent((9, *aug.output_shape), ACTION, device, aug, **HP)
    g = torch.Generator(device=device).manual_seed(1)
    frames = torch.empty((N, *FRAME), dtype=torch.uint8, device=device)          # 2.6 GB > L2
    for s0 in range(0, N, 2000):
        frames[s0:s0 + 2000] = torch.randint(0, 256, (min(2000, N - s0), *FRAME), device=device, dtype=torch.uint8, generator=g)
    top, left = latent.center_window(FRAME[1:], tuple(aug.output_shape))
    sweep = {}
    for b in (1, 8, 64, 512):
        agent.INFER_BATCH = b
        agent._make_engine(batch=b, frame_hw=agent.image_shape)
        n = min(N, max(b * 8, 256)) if b < 512 else N
        for _ in range(max(args.warmup, 3)):
            agent.encode_frames(0, frames[:min(n, 4 * b)], top, left)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        z = agent.encode_frames(0, frames[:n], top, left)
        e1.record()
        torch.cuda.synchronize()
        sweep[str(b)] = {'obs_per_s': n / (e0.elapsed_time(e1) * 1e-3), 'observations': n}
    # end to end through the public call with HOST frames: H2D of the uint8 frames + encoders + policy + Q + D2H
    agent.INFER_BATCH = 512
    agent._make_engine(batch=512, frame_hw=agent.image_shape)
    host = frames[:4096].cpu().numpy()
    latent.extract(agent, host[:1024])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = latent.extract(agent, host)
    e2e_s = time.perf_counter() - t0
    # CPU: the oracle port of agent.actor.encoder at batch 64 on the host cores (bounded sample)
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    o = O.OracleAgent((9, *aug.output_shape), 2, hidden_dim=HP['hidden_dim'])
    o.init_random(0)
    x = torch.from_numpy(host[:64, :, top:top + 76, left:left + 135].astype(np.float32))
    with torch.no_grad():
        O.encoder_forward(o.actor, 'encoder.', x)
        t0 = time.perf_counter()
        for _ in range(4):
            O.encoder_forward(o.actor, 'encoder.', x)
        cpu_s = (time.perf_counter() - t0) / 4
    best = max(sweep, key=lambda k: sweep[k]['obs_per_s'])
    hbm, tf, how = peaks()
    flops_obs = 134.01e6
    line = {'metric': 'encoder_latent_obs_per_sec', 'value': sweep[best]['obs_per_s'], 'unit': 'obs/s', 'n_gpus': 1,
            'steps': 1, 'warmup': max(args.warmup, 3), 'ms_per_step': 1e3 * sweep[best]['observations'] / sweep[best]['obs_per_s'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': 'encoder-only latent extraction (actor.encoder forward, center crop 90x160 -> 76x135) over %d '
                                   'synthetic uint8 frames resident in HBM (2.6 GB, > L2); batch sweep' % N,
                       'best_batch': int(best)},
            'sweep': sweep,
            'roofline': {'bound': 'tensor', 'achieved': sweep[best]['obs_per_s'] * flops_obs / 1e12, 'peak': tf, 'unit': 'TFLOP/s',
                         'frac': sweep[best]['obs_per_s'] * flops_obs / 1e12 / tf, 'traffic': None, 'peak_source': how,
                         'note': '134.0 MFLOP per observation (SURVEY 8d); N=32 conv taps cap the tensor pipe at 40 % (profiles/r01c_microbench.txt)'},
            'e2e': {'value': len(host) / e2e_s, 'unit': 'obs/s', 'h2d_bytes_per_step': int(host.nbytes),
                    'd2h_bytes_per_step': int(sum(v.nbytes for v in out.values())),
                    'what': 'curla_b200.latent.extract(agent, host uint8 frames): latents + sampled actions + min-Q for 4096 observations'},
            'cpu_baseline': {'value': 64 / cpu_s, 'unit': 'obs/s', 'cores': cores, 'kind': 'port',
                             'sample': 'oracle port of actor.encoder forward (torch fp32 CPU), batch 64 x 4 passes'},
            'gpu_launches': None}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=4)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'reference_cuda'])
    ap.add_argument('--tf32', type=int, default=0, help='--impl reference_cuda: allow TF32 in cuDNN/cuBLAS (informational baseline)')
    ap.add_argument('--batch', type=int, default=BATCH, help='per-GPU batch (default: train.py default 512)')
    ap.add_argument('--prof-steps', type=int, default=4)
    ap.add_argument('--cpu-baseline-steps', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='curl_crop', choices=sorted(WORKLOADS) + ['latent'],
                    help='curl_crop = BASELINE.json configs[1] (default, the metric); pixel_sac = configs[2]; '
                         'curl_identity = configs[0] at the default batch; latent = configs[4] (encoder-only sweep)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    cores = os.cpu_count()
    if args.workload == 'latent':
        return latent_sweep(args, rank, local)
    wl = WORKLOADS[args.workload]
    config = {'workload': '%s: per-GPU batch %d, frames 9x90x160%s, hidden 1024, feature 50, 4x32 filters; synthetic replay '
                          'of %d transitions (4.25 GB, > L2) resident in HBM'
                          % (wl['what'], args.batch, ' -> crop 9x76x135' if wl['aug'] == 'random_crop' else '', CAPACITY),
              'per_gpu_batch': args.batch, 'global_batch': args.batch * world,
              'parallelism': 'dp%d (grad all-reduce + all-gathered CURL keys)' % world,
              'l2': 'inputs larger than L2 (random replay rows from 4.25 GB; ~1.4 GB of activations per update)',
              'precision': 'bf16 operands, fp32 accumulate, fp32 master weights/Adam/EMA/LN/losses',
              'launch': ('eager kernel launches (CURLA_GRAPH=0)' if os.environ.get('CURLA_GRAPH', '1')[:1] == '0' else
                         'one captured CUDA graph per update variant, replayed (Adam step counters and Philox offset in device memory)')}

    if args.impl in ('reference', 'reference_cuda'):
        if rank != 0:
            return
        # the reference's own update at the SAME configuration (batch, capacity, hyper-parameters) for the
        # REAL step count: K full updates after W warm-ups (about 1.4 s each on 16 host cores)
        cuda = args.impl == 'reference_cuda'
        if cuda and not torch.cuda.is_available():
            raise SystemExit('--impl reference_cuda needs a CUDA device')
        sec, kind = reference_updates(args.batch, args.steps, args.warmup, cores, args.workload,
                                      device='cuda' if cuda else 'cpu', tf32=bool(args.tf32))
        val = args.batch / sec
        what = ('unmodified reference CurlSacAgent.update + ReplayBuffer.sample_cpc (baseline/_ref via oracle/shims)'
                if kind == 'reference' else 'oracle port of the reference update (reference modules absent)')
        where = ('PyTorch %s eager on cuda:0 (cuDNN/cuBLAS), allow_tf32=%s, host replay + H2D per update as the reference does'
                 % (torch.__version__, bool(args.tf32))) if cuda else 'torch fp32 CPU, %d threads' % cores
        line = {'impl': args.impl, 'metric': 'sac_curl_update_obs_per_sec', 'value': val, 'unit': 'obs/s',
                'updates_per_s': 1.0 / sec, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': val, 'unit': 'obs/s', 'cores': cores, 'kind': kind,
                                 'sample': '%s; %s; batch %d, replay capacity %d, %d timed updates after %d warm-ups'
                                           % (what, where, args.batch, CAPACITY, args.steps, args.warmup)},
                'e2e': {'value': val, 'unit': 'obs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device: the hot path has no CPU fallback')
    device = torch.device('cuda', local)
    if world > 1:
        torch.cuda.set_device(device)
        torch.distributed.init_process_group('nccl', device_id=device)
    r = run_ours(args, rank, world, device)
    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    Bg = r['Bg']
    ups = 1e3 / r['ms_per_step']
    value = Bg * ups
    hbm, tf, how = peaks()
    # per-kernel rooflines from the live CUDA-event profile (algorithmic bytes: DESIGN.md section 4);
    # the headline `roofline` entry is the kernel with the largest share of the update
    H_, W_ = r['obs_hw']
    ho, wo = geometry((H_, W_))
    Bb = args.batch
    act = lambda l: Bb * ho[l] * wo[l] * 64                      # valid bf16 positions x 32 channels
    # space-to-depth input: 4*9 = 36 real channels live in five 8-channel planes (the sixth plane of the 48-channel
    # operand is all zero: never written by the gather, never read by conv-1)
    s2d_bytes = Bb * ((H_ + 1) // 2) * ((W_ + 1) // 2) * 40 * 2
    kfc = ho[3] * ((W_ + 1) // 2) * 32                             # fc input in the pitch layout (67,456 for the crop)
    conv_flops = lambda l: 2.0 * Bb * ho[l] * wo[l] * 32 * (81 if l == 0 else 288)
    table = {
        'conv_fwd': dict(kernel='k_conv_tc<fwd> (tcgen05 conv forward, 4 layers)', launches=4,
                         bytes=s2d_bytes + act(0) + sum(act(l - 1) + act(l) for l in (1, 2, 3)),
                         flops=sum(conv_flops(l) for l in range(4))),
        'conv_dgrad': dict(kernel='k_conv_tc<dgrad> (tcgen05 conv dgrad, layers 4..2)', launches=3,
                           bytes=sum(act(l) + 2 * act(l - 1) for l in (1, 2, 3)),
                           flops=sum(conv_flops(l) for l in (1, 2, 3))),
        'conv_wgrad': dict(kernel='k_conv_wgrad (conv wgrad, 4 layers)', launches=4,
                           bytes=s2d_bytes + act(0) + sum(act(l - 1) + act(l) for l in (1, 2, 3)),
                           flops=sum(conv_flops(l) for l in range(4))),
        # ONE launch gathers every stream of the update (obs, next_obs, and for the random crop an independent pos
        # window: utils.py:151-158): algorithmic bytes = streams x (uint8 window read + bf16 s2d planes written)
        'gather_s2d': dict(kernel='k_gather_s2d_u8 (replay gather + crop + u8->bf16 s2d, cp.async.bulk staged, all streams in one launch)',
                           launches=1, bytes=(3 if (wl['aug'] == 'random_crop' and not wl['pixel_sac']) else 2) * (Bb * 9 * H_ * W_ + s2d_bytes),
                           flops=0.0),
        'gemm_fc_fwd': dict(kernel='k_gemm_tc (tcgen05 + TMA, encoder fc forward, split-K)', launches=1,
                            bytes=Bb * kfc * 2 + 64 * kfc * 2, flops=2.0 * Bb * 64 * kfc),
        'gemm_fc_dgrad': dict(kernel='k_gemm_tc_nloop (tcgen05 + TMA, encoder fc dgrad + ReLU mask)', launches=1,
                              bytes=2 * Bb * kfc * 2 + 64 * kfc * 2, flops=2.0 * Bb * 64 * kfc),
        'gemm_fc_wgrad': dict(kernel='k_gemm_tc (tcgen05 + TMA, encoder fc wgrad, swapped operands)', launches=1,
                              bytes=Bb * kfc * 2 + 50 * kfc * 4, flops=2.0 * Bb * 64 * kfc),
    }
    # elementwise optimizer kernels: 28 B per parameter stepped (p, g, m, v read; p, m, v written), 12 B per parameter of
    # the EMA (target, online read; target written).  The profiled steps start on an even step: the critic and the
    # encoder+CURL optimizers step every update, the actor's and the EMA every second one (train.py defaults).
    ps_ = r['prof_steps']
    np_ = r['nparams']
    even = (ps_ + 1) // 2
    if not wl['pixel_sac']:
        adam_params = ps_ * (np_['critic'] + np_['cpc']) + even * np_['actor']
    else:
        adam_params = ps_ * np_['critic'] + even * np_['actor']
    table['adam_f32'] = dict(kernel='k_adam (fused Adam, one launch per optimizer)', total_bytes=28.0 * adam_params, flops=0.0)
    table['ema_f32'] = dict(kernel='k_ema (soft target update, encoder + Q1 + Q2 in one launch)', total_bytes=12.0 * even * np_['critic'], flops=0.0)
    traffic_db = {}
    tp = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')       # per-launch dram bytes from ncu --set full
    if os.path.exists(tp):
        traffic_db = json.load(open(tp))
    roofs = {}
    for name, t_ in table.items():
        if name not in r['prof']:
            continue
        cnt, tot = r['prof'][name]
        if 'total_bytes' in t_:                                   # bytes of all profiled launches given directly
            t_ = dict(t_, bytes=t_['total_bytes'] / cnt, launches=1)
        passes = cnt / float(t_['launches'])                      # conv-stack passes in the profiled steps
        if name == 'conv_fwd':
            # independent passes share launches (curla_conv_fwd_multi), so count passes from the update
            # schedule instead: CURL 5 per update; pixel SAC 4 with / 3 without the actor step
            # (the profiled steps start on an even step and alternate)
            ps = r['prof_steps']
            passes = 5.0 * ps if not wl['pixel_sac'] else 4.0 * ((ps + 1) // 2) + 3.0 * (ps // 2)
        sec = tot / 1e3
        ach = t_['bytes'] * passes / sec / 1e9
        roofs[name] = {'kernel': t_['kernel'], 'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s',
                       'frac': ach / hbm, 'traffic': traffic_db.get(name), 'peak_source': how,
                       'algorithmic_bytes_per_launch': t_['bytes'] * passes / cnt,
                       'avg_launch_us': tot * 1e3 / cnt, 'tflops': t_['flops'] * passes / sec / 1e12,
                       'launches_per_update': cnt / r['prof_steps'], 'ms_per_update': tot / r['prof_steps']}
    roof = None
    if roofs:
        top = max(roofs, key=lambda k: roofs[k]['ms_per_update'])
        roof = dict(roofs[top])
        roof['other_kernels'] = {k: {kk: v[kk] for kk in ('achieved', 'frac', 'avg_launch_us', 'ms_per_update', 'tflops')}
                                 for k, v in roofs.items() if k != top}
    breakdown = {k: round(v[1] / r['prof_steps'], 4) for k, v in sorted(r['prof'].items(), key=lambda kv: -kv[1][1])}

    cpu = None
    if not args.no_cpu_baseline:
        # bounded sample of the SAME workload: the reference's own update at the full batch and replay
        # capacity, a few updates (about 1.4 s each on 16 cores)
        sec, kind = reference_updates(args.batch, args.cpu_baseline_steps, 1, cores, args.workload)
        cpu = {'value': args.batch / sec, 'unit': 'obs/s', 'cores': cores, 'kind': kind,
               'sample': '%s, torch fp32 CPU %d threads: batch %d, replay capacity %d, %d updates after 1 warm-up'
                         % ('unmodified reference update (baseline/_ref)' if kind == 'reference' else 'oracle port of the reference update',
                            cores, args.batch, CAPACITY, args.cpu_baseline_steps),
               'updates_per_s': 1.0 / sec}
    line = {'metric': 'sac_curl_update_obs_per_sec', 'value': value, 'unit': 'obs/s', 'updates_per_s': ups,
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': config, 'clocks': r['clocks'],
            'e2e': {'value': Bg * args.steps / r['e2e_s'], 'unit': 'obs/s', 'updates_per_s': args.steps / r['e2e_s'],
                    'h2d_bytes_per_step': r['h2d'], 'd2h_bytes_per_step': r['d2h'],
                    'what': 'ReplayBuffer.add(host frame) + CurlSacAgent.update(...) + logged scalars read back, per step'},
            'gpu_launches': r['launches'], 'roofline': roof, 'cpu_baseline': cpu,
            'kernel_ms_per_update': breakdown, 'losses': r['last']}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
