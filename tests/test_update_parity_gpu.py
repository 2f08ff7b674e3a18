"""Whole-update parity: CUDA CurlSacAgent.update vs the CPU oracle (and the reference's
golden fixtures) on identical replay contents, weights, sampled indices, crop offsets and
policy noise.  Run on the B200 box: pytest -m gpu.

Tolerances (fp32 reference vs bf16-operand / fp32-accumulate tensor-core kernels):
  sampling, crop, gather .......... bit exact
  latents z, Q values ............. 2e-2 relative L2
  losses .......................... 2e-2 relative (curl loss: 2e-2 absolute on O(1..5) values)
  gradients (per tensor) .......... 5e-2 relative L2 (tensors with non-negligible norm)
  parameters after the step ....... |dp| within 2.5*lr per element of the oracle's and
                                     mean |dp error| < 0.35*lr  (Adam is sign-like for the
                                     first steps, so tiny-gradient elements may flip)
"""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import scenario as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DEV = 'cuda'


class NullLogger:
    def __init__(self):
        self.rows = {}

    def log(self, key, value, step):
        self.rows[(step, key)] = float(value)


def rel_l2(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def build_cuda_agent(cfg, run):
    from curla_b200 import augmentations, curl_sac, utils
    hw = tuple(cfg['frame_hw'])
    augmentor = augmentations.make_augmentor(cfg['aug'], hw)
    assert tuple(augmentor.output_shape) == tuple(run.ohw)
    rb = utils.ReplayBuffer((S.FRAME_C, *hw), (S.ACTION_DIM,), cfg['capacity'], cfg['B'], DEV, augmentor)
    for dst, src in zip((rb.obses, rb.next_obses, rb.actions, rb.rewards, rb.not_dones), run.arrays):
        dst.copy_(torch.from_numpy(src))
    rb.idx, rb.full = 0, True
    agent = curl_sac.CurlSacAgent(run.obs_shape, (S.ACTION_DIM,), DEV, augmentor, hidden_dim=cfg['hidden'],
                                  log_interval=1, detach_encoder=cfg['detach_encoder'],
                                  pixel_sac=cfg['pixel_sac'], **S.HP)
    actor_sd, critic_sd, W = run.state_dicts
    agent.critic.load_state_dict(critic_sd)
    agent.actor.load_state_dict(actor_sd)
    agent.critic_target.load_state_dict(agent.critic.state_dict())
    agent.CURL.W.copy_(W)
    return agent, rb


def grad_views(agent):
    """name -> gradient tensor, decoded from the flat grad arenas."""
    eng = agent.engine
    out = {}
    base = {}
    p_off = lambda k: eng.info[k][1] // 4
    crit0 = p_off('critic.encoder.convs.0.weight')
    act0 = p_off('actor.encoder.fc.weight_canon')
    w0 = p_off('CURL.W')
    for k, (arena, off, shape, dt) in eng.info.items():
        if arena != 0:
            continue
        n = int(np.prod(shape))
        o = off // 4
        if k.startswith('critic.'):
            out['critic_opt/' + k] = eng.t['grad.critic'][o - crit0:o - crit0 + n].view(shape)
            if k.startswith('critic.encoder.'):
                out['cpc_opt/' + k] = eng.t['grad.cpc'][o - w0:o - w0 + n].view(shape)
        elif k.startswith('actor.'):
            out['actor_opt/' + k] = eng.t['grad.actor'][o - act0:o - act0 + n].view(shape)
        elif k == 'CURL.W':
            out['cpc_opt/W'] = eng.t['grad.cpc'][0:n].view(shape)
    return out


def to_torch_layout(eng, key, t):
    return eng.fc_to_torch(t) if key.endswith('fc.weight_canon') else t


@pytest.mark.parametrize('name', list(S.SCENARIOS))
def test_update_matches_oracle(name):
    torch.set_num_threads(max(1, os.cpu_count() // 2))
    cfg = S.SCENARIOS[name]
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    run = S.OracleRun(cfg)
    agent, rb = build_cuda_agent(cfg, run)
    np_state = np.random.get_state()          # OracleRun seeded the global stream
    L = NullLogger()
    lr = 1e-3
    report = []
    for u, (step, only_cpc) in enumerate(zip(cfg['steps'], cfg['only_cpc'])):
        # oracle step (consumes the numpy stream), then rewind for the CUDA agent
        np.random.set_state(np_state)
        d, b, om = run.step()
        after = np.random.get_state()
        np.random.set_state(np_state)
        agent._noise_override = (run.noise[u, 0], run.noise[u, 1])
        before = {k: v.clone() for k, v in agent.engine.t.items() if agent.engine.info[k][0] == 0}
        agent.update(rb, L, step, only_cpc=only_cpc)
        torch.cuda.synchronize()
        assert np.array_equal(np.random.get_state()[1], after[1]), 'RNG consumption differs'
        np_state = after
        eng, o = agent.engine, run.agent
        p = 'u%d/' % u
        # ---- sampled frames: bit exact (decode the s2d staging buffers)
        B, Cc = cfg['B'], S.FRAME_C
        H, W = run.ohw
        Hs, Ws = (H + 1) // 2, (W + 1) // 2
        def unpack(key):
            v = eng.t[key].view(B, Hs, Ws, -1)[..., :Cc * 4].float().view(B, Hs, Ws, Cc, 2, 2)
            return v.permute(0, 3, 1, 4, 2, 5).reshape(B, Cc, 2 * Hs, 2 * Ws)[:, :, :H, :W].to(torch.uint8).cpu().numpy()
        assert np.array_equal(unpack('s2d.obs'), b['obs'])
        assert zlib.crc32(unpack('s2d.obs').tobytes()) == int(gold[p + 'crc/obs'][0])
        if not only_cpc:
            assert np.array_equal(unpack('s2d.next'), b['next'])
            assert np.array_equal(eng.t['batch.action'].cpu().numpy(), b['action'])
            assert np.array_equal(eng.t['batch.reward'].cpu().numpy(), b['reward'][:, 0])
        if cfg['aug'] == 'random_crop' and not cfg['pixel_sac']:
            assert np.array_equal(unpack('s2d.pos'), b['pos'])
        # ---- scalars
        keymap = {'train/batch_reward': 'batch_reward', 'train_critic/loss': 'critic_loss',
                  'train_actor/loss': 'actor_loss', 'train_actor/entropy': 'entropy',
                  'train_alpha/loss': 'alpha_loss', 'train/curl_loss': 'curl_loss'}
        for rk, ok in keymap.items():
            if ok in om:
                got = L.rows[(step, rk)]
                tol = 2e-2 * max(abs(om[ok]), 1.0 if ok in ('curl_loss', 'actor_loss', 'alpha_loss') else 1e-3)
                assert abs(got - om[ok]) <= tol, (name, u, rk, got, om[ok])
                assert abs(got - float(gold[p + 'metric/' + rk][0])) <= 1.5 * tol, (name, u, rk, 'golden')
                report.append((u, rk, got, om[ok]))
        # ---- latents / Q
        if not only_cpc:
            assert rel_l2(eng.t['p3.z'][:, :S.FEATURE_DIM], o.dbg['z_critic']) < 2e-2
            assert rel_l2(eng.t['p3.q1.out'], o.dbg['q1']) < 2e-2
            assert rel_l2(eng.t['target_q'], o.dbg['target_q'][:, 0]) < 2e-2
        if 'z_a' in o.dbg:
            assert rel_l2(eng.t['p5.z'][:, :S.FEATURE_DIM], o.dbg['z_a']) < 2e-2
            assert rel_l2(eng.t['p7.z'][:, :S.FEATURE_DIM], o.dbg['z_pos']) < 2e-2
            assert rel_l2(eng.t['p5.z'][:, :S.FEATURE_DIM], torch.from_numpy(gold[p + 'out/critic_z'])) < 2e-2
        # ---- gradients
        gv = grad_views(agent)
        for tag, og in (('critic_opt', o.dbg.get('critic_grads')), ('actor_opt', o.dbg.get('actor_grads')),
                        ('cpc_opt', o.dbg.get('cpc_grads'))):
            if og is None:
                continue
            net = 'actor.' if tag == 'actor_opt' else 'critic.'
            tot_n, tot_d = 0.0, 0.0
            for k, g_ref in og.items():
                ek = net + k.replace('fc.weight', 'fc.weight_canon') if k.endswith('encoder.fc.weight') else net + k
                ours = to_torch_layout(eng, ek, gv[tag + '/' + ek])
                dn = float((ours.cpu().double() - g_ref.double()).norm())
                rn = float(g_ref.double().norm())
                tot_n += dn ** 2; tot_d += rn ** 2
                if rn > 1e-6:
                    assert dn / rn < 5e-2, (name, u, tag, k, dn / rn)
            assert (tot_n / max(tot_d, 1e-30)) ** 0.5 < 3e-2, (name, u, tag)
        if 'W_grad' in o.dbg:
            assert rel_l2(gv['cpc_opt/W'], o.dbg['W_grad']) < 5e-2
        # ---- parameters after the update
        for net, osd in (('actor', o.actor), ('critic', o.critic), ('target', o.target)):
            for k, v in osd.items():
                if net == 'actor' and k.startswith('encoder.convs.'):
                    continue
                ek = net + '.' + (k.replace('fc.weight', 'fc.weight_canon') if k == 'encoder.fc.weight' else k)
                ours = to_torch_layout(eng, ek, eng.t[ek]).cpu()
                err = (ours - v.detach()).abs()
                assert float(err.max()) <= 2.5 * lr * (u + 1), (name, u, ek, float(err.max()))
                assert float(err.mean()) <= 0.35 * lr * (u + 1), (name, u, ek, float(err.mean()))
        assert float((eng.t['CURL.W'].cpu() - o.W.detach()).abs().max()) <= 2.5 * lr * (u + 1)
        assert abs(float(agent.log_alpha) - float(o.log_alpha.detach())) < 2e-6 * (u + 1)
        assert abs(float(agent.log_alpha) - float(gold[p + 'param/log_alpha'][0])) < 2e-6 * (u + 1)
    print('\n'.join('%s u%d %-24s cuda % .6f oracle % .6f' % (name, *r) for r in report))
