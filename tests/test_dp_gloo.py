"""World-size-2 (gloo, CPU) tests of the data-parallel plan of CurlSacAgent.update.

What the engine does with NCCL on the GPUs (csrc/engine.cu) is restated here with
torch.distributed/gloo on the CPU oracle: each rank takes its slice of the IDENTICAL global
index block, computes its shard's gradient sums scaled by 1/B_global, all-gathers the CURL
keys, offsets its labels by rank*B and SUM-all-reduces the gradients.  The result must equal
the single-process full-batch oracle update (the reference's semantics: full-batch negatives
and means over the global batch, curl_sac.py:359,379,411-413).
"""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from curla_b200 import dp
from oracle import curla_oracle as O

OBS = (9, 64, 64)
BG, WORLD, HID = 8, 2, 32


def _make_agent(seed=0):
    a = O.OracleAgent(OBS, 2, hidden_dim=HID)
    a.init_random(seed)
    return a


def _batch(seed=1):
    rs = np.random.RandomState(seed)
    f = lambda *s: torch.from_numpy(rs.randint(0, 256, size=s).astype(np.float32))
    obs, nxt, pos = f(BG, *OBS), f(BG, *OBS), f(BG, *OBS)
    act = torch.from_numpy(rs.uniform(-1, 1, size=(BG, 2)).astype(np.float32))
    rew = torch.from_numpy(rs.standard_normal(size=(BG, 1)).astype(np.float32))
    nd = torch.from_numpy((rs.uniform(size=(BG, 1)) > 0.1).astype(np.float32))
    noise = torch.from_numpy(rs.standard_normal(size=(2, BG, 2)).astype(np.float32))
    return obs, act, rew, nxt, nd, pos, noise


def _sharded_grads(agent, batch, rank, world):
    """One rank's share of the critic and CURL gradients, exactly as the engine forms them."""
    obs, act, rew, nxt, nd, pos, noise = batch
    sl = dp.shard_slice(rank, world, BG)
    gs = dp.grad_scale(BG)
    B = dp.local_batch(BG, world)
    # ---- critic loss on the shard (sum of squared errors * 1/B_global == share of the global mean)
    with torch.no_grad():
        _, pa, lp, _ = O.actor_forward(agent.actor, nxt[sl], noise[0][sl], agent.log_std_min, agent.log_std_max)
        tq1, tq2 = O.critic_forward(agent.target, nxt[sl], pa)
        tq = rew[sl] + nd[sl] * agent.discount * (torch.min(tq1, tq2) - agent.alpha.detach() * lp)
    q1, q2 = O.critic_forward(agent.critic, obs[sl], act[sl])
    loss_c = (((q1 - tq) ** 2).sum() + ((q2 - tq) ** 2).sum()) * gs
    for v in agent.critic.values():
        v.grad = None
    loss_c.backward()
    g_critic = {k: v.grad.clone() for k, v in agent.critic.items()}
    # ---- CURL: local anchors, ALL-GATHERED keys, labels offset by rank*B
    for v in agent.critic.values():
        v.grad = None
    agent.W.grad = None
    z_a = O.encoder_forward(agent.critic, 'encoder.', obs[sl])
    with torch.no_grad():
        z_pos_local = O.encoder_forward(agent.target, 'encoder.', pos[sl])
    parts = [torch.zeros_like(z_pos_local) for _ in range(world)]
    dist.all_gather(parts, z_pos_local)
    z_pos = torch.cat(parts, 0)
    logits = O.curl_logits(agent.W, z_a, z_pos)                       # (B, B_global)
    labels = dp.label_offset(rank, world, BG) + torch.arange(B)
    loss_k = F.cross_entropy(logits, labels, reduction='sum') * gs
    loss_k.backward()
    g_cpc = {k: v.grad.clone() for k, v in agent.critic.items() if k.startswith('encoder.')}
    g_cpc['W'] = agent.W.grad.clone()
    # ---- SUM all-reduce of the flat buckets
    for g in (g_critic, g_cpc):
        flat = torch.cat([g[k].flatten() for k in sorted(g)])
        dist.all_reduce(flat)
        o = 0
        for k in sorted(g):
            n = g[k].numel()
            g[k] = flat[o:o + n].view_as(g[k]).clone()
            o += n
    losses = torch.tensor([float(loss_c.detach()), float(loss_k.detach())], dtype=torch.float64)
    dist.all_reduce(losses)
    return g_critic, g_cpc, losses


def _worker(rank, world, init_file, out_dir):
    dist.init_process_group('gloo', init_method='file://' + init_file, rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        assert dp.world_info() == (rank, world)
        # unique-id style broadcast
        payload = bytes(range(128)) if rank == 0 else b''
        got = dp.broadcast_bytes(payload, 128)
        assert got == bytes(range(128))
        # identical global draws on every rank (same seed), disjoint contiguous slices
        np.random.seed(1234)
        d = O.draw_sample_indices(64, 0, True, BG, 'random_crop', (90, 160), (76, 135))
        mine = d['idxs'][dp.shard_slice(rank, world, BG)]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine.tolist())
        assert sum(gathered, []) == d['idxs'].tolist()
        agent = _make_agent()
        g_critic, g_cpc, losses = _sharded_grads(agent, _batch(), rank, world)
        if rank == 0:
            torch.save(dict(g_critic=g_critic, g_cpc=g_cpc, losses=losses), os.path.join(out_dir, 'dp.pt'))
    finally:
        dist.destroy_process_group()


def test_dp_host_helpers():
    assert dp.local_batch(512, 8) == 64
    assert dp.shard_slice(3, 8, 512) == slice(192, 256)
    assert dp.label_offset(3, 8, 512) == 192
    assert dp.grad_scale(4096) == 1.0 / 4096
    with pytest.raises(ValueError):
        dp.local_batch(10, 4)
    assert dp.world_info() == (0, 1)


def test_dp_two_ranks_equal_full_batch():
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, 'rdzv')
        mp.spawn(_worker, args=(WORLD, init_file, tmp), nprocs=WORLD, join=True)
        got = torch.load(os.path.join(tmp, 'dp.pt'), weights_only=False)
    # single-process full-batch oracle (the reference semantics)
    agent = _make_agent()
    obs, act, rew, nxt, nd, pos, noise = _batch()
    agent.update_critic(obs, act, rew, nxt, nd, noise[0])
    ref_critic = agent.dbg['critic_grads']
    ref_closs = agent.metrics['critic_loss']
    agent2 = _make_agent()
    agent2.dbg = {}
    agent2.update_cpc(obs, pos)
    # fp32 summation order differs (2 shards + all-reduce vs one pass): 2e-5 of the tensor's max
    same = lambda a, b: float((a - b).abs().max()) <= 2e-5 * max(float(b.abs().max()), 1e-12)
    for k, g in ref_critic.items():
        assert same(got['g_critic'][k], g), k
    for k, g in agent2.dbg['cpc_grads'].items():
        assert same(got['g_cpc'][k], g), k
    assert same(got['g_cpc']['W'], agent2.dbg['W_grad'])
    assert abs(float(got['losses'][0]) - ref_closs) < 1e-4 * max(1.0, abs(ref_closs))
    assert abs(float(got['losses'][1]) - agent2.metrics['curl_loss']) < 1e-4
