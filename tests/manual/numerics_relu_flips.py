"""CPU experiment (not a test): how many ReLU units of the MLP trunks flip under bf16 operands, and what that does to
the ACTOR gradient layer by layer.

    python tests/manual/numerics_relu_flips.py        (output: profiles/r05_numerics_relu_flips.txt)

tests/test_update_parity_gpu.py bounds the actor bucket of the eight-step B = 64 / hidden-256 scenario by 8e-2 because
update 6 measures 6.9e-2 on the GPU: the policy trunk's output layer agrees to 0.6e-2, its hidden layers and everything
below them to 1.1e-1 each.  This script reproduces the mechanism on the oracle alone: it runs the scenario's first N
updates in fp32 (the oracle's own trajectory), then evaluates the actor loss of the next update twice on the SAME
state -- in fp32, and with the trunk GEMM operands (weights, the stored input row and the stored hidden activations)
rounded to bf16 as the CUDA path stores them (fp32 accumulation, fp32 everything else; the encoder latents are
computed in fp32 both times so that only the MLPs differ) -- counts the hidden units whose ReLU mask differs, and
reports the relative L2 error of every actor gradient tensor."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import curla_oracle as O
from oracle import scenario as S

torch.set_num_threads(8)
cfg = dict(aug='random_crop', frame_hw=(90, 160), B=64, capacity=128, hidden=256, steps=list(range(8)),
           only_cpc=[False, False, False, True, False, False, False, False], pixel_sac=False, detach_encoder=False)
bf = lambda t: t + (t.to(torch.bfloat16).float() - t).detach()          # bf16 value, fp32 master (straight-through)


def mlp3(p, prefix, x, rounded, masks):
    q = bf if rounded else (lambda t: t)
    h1 = torch.relu(F.linear(q(x), q(p[prefix + '0.weight']), p[prefix + '0.bias']))
    masks.append((h1 > 0).detach())
    h2 = torch.relu(F.linear(q(h1), q(p[prefix + '2.weight']), p[prefix + '2.bias']))
    masks.append((h2 > 0).detach())
    return F.linear(q(h2), p[prefix + '4.weight'], p[prefix + '4.bias'])           # (the head reads fp32 weights on the GPU too)


def actor_grads(ag, obs, noise, rounded):
    masks = []
    with torch.no_grad():
        z_a = O.encoder_forward(ag.actor, 'encoder.', obs)                          # fp32 latents both times
        z_c = O.encoder_forward(ag.critic, 'encoder.', obs)
    params = {k: v for k, v in ag.actor.items() if k.startswith('trunk.')}
    for v in params.values():
        v.grad = None
    mu, log_std = mlp3(ag.actor, 'trunk.', z_a, rounded, masks).chunk(2, dim=-1)
    log_std = torch.tanh(log_std)
    log_std = ag.log_std_min + 0.5 * (ag.log_std_max - ag.log_std_min) * (log_std + 1)
    pi = mu + noise * log_std.exp()
    log_pi = O.gaussian_logprob(noise, log_std)
    pi = torch.tanh(pi)
    log_pi = log_pi - torch.log(F.relu(1 - pi.pow(2)) + 1e-6).sum(-1, keepdim=True)
    za = torch.cat([z_c, pi], dim=1)
    q1 = mlp3(ag.critic, 'Q1.trunk.', za, rounded, masks)
    q2 = mlp3(ag.critic, 'Q2.trunk.', za, rounded, masks)
    loss = (ag.alpha.detach() * log_pi - torch.min(q1, q2)).mean()
    gs = torch.autograd.grad(loss, list(params.values()))
    return float(loss), dict(zip(params.keys(), gs)), masks


lines = []
for n_updates in (0, 2, 4, 6):
    run = S.OracleRun(cfg)
    for _ in range(n_updates):
        run.step()
    d, b = run.sample()
    obs = torch.from_numpy(b['obs']).float()
    noise = run.noise[n_updates, 1]
    l0, g0, m0 = actor_grads(run.agent, obs, noise, False)
    l1, g1, m1 = actor_grads(run.agent, obs, noise, True)
    names = ['actor H1', 'actor H2', 'Q1 H1', 'Q1 H2', 'Q2 H1', 'Q2 H2']
    flips = ', '.join('%s %d of %d active' % (nm, int((a != b_).sum()), int(a.sum())) for nm, a, b_ in zip(names, m0, m1))
    num = sum(float((g1[k] - g0[k]).pow(2).sum()) for k in g0)
    den = sum(float(g0[k].pow(2).sum()) for k in g0)
    lines.append('after %d fp32 updates: actor loss fp32 %.5f / bf16 operands %.5f; flipped ReLU units: %s' % (n_updates, l0, l1, flips))
    lines.append('    trunk gradient bucket rel-L2 error %.2e; per tensor: %s' % (
        (num / den) ** 0.5, ', '.join('%s %.2e' % (k[len('trunk.'):], float((g1[k] - g0[k]).norm() / g0[k].norm())) for k in g0)))
print('\n'.join(lines))
