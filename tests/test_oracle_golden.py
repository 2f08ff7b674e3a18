"""The oracle (oracle/curla_oracle.py) against outputs of the UNMODIFIED reference.

tests/golden/*.npz were produced by oracle/make_golden.py in the build container
(the reference ships no tests or golden vectors of its own: SURVEY.md section 4).
Sampling / crop indexing must be bit-exact; fp32-vs-fp32 values within 1e-4."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import scenario as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

METRIC_KEYS = {  # reference log key -> oracle metric name (curl_sac.py:361-432)
    'train/batch_reward': 'batch_reward', 'train_critic/loss': 'critic_loss',
    'train_actor/loss': 'actor_loss', 'train_actor/entropy': 'entropy',
    'train_alpha/loss': 'alpha_loss', 'train_alpha/value': 'alpha', 'train/curl_loss': 'curl_loss'}


def close(a, b, rtol=2e-4, atol=2e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return np.all(np.abs(a - b) <= atol + rtol * np.maximum(np.abs(b), scale * 1e-3))


def fp_close(a, b, rtol):
    """Fingerprints (S.summarize): [sum, abs-sum, samples...].  The plain sum
    cancels heavily, so it is compared relative to the abs-sum; samples relative
    to the largest sample."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ok = abs(a[0] - b[0]) <= rtol * max(b[1], 1e-30) and abs(a[1] - b[1]) <= rtol * max(b[1], 1e-30)
    return ok and np.abs(a[2:] - b[2:]).max() <= rtol * max(np.abs(b[2:]).max(), 1e-30)


@pytest.mark.parametrize('name', list(S.SCENARIOS))
def test_oracle_matches_reference(name):
    torch.set_num_threads(1)
    cfg = S.SCENARIOS[name]
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    run = S.OracleRun(cfg)
    for u in range(len(cfg['steps'])):
        d, b, metrics = run.step()
        p = 'u%d/' % u
        # ---- integer / byte work: bit exact
        for k, v in d.items():
            assert np.array_equal(v, gold[p + 'idx/' + k]), (name, u, k)
        for k in ('obs', 'next', 'pos'):
            assert zlib.crc32(b[k].tobytes()) == int(gold[p + 'crc/' + k][0]), (name, u, k)
        for k in ('action', 'reward', 'not_done'):
            assert np.array_equal(b[k], gold[p + 'batch/' + k])
        # ---- scalars
        for rk, ok in METRIC_KEYS.items():
            if p + 'metric/' + rk in gold.files:
                assert close(metrics[ok], gold[p + 'metric/' + rk][0]), (name, u, rk)
            else:
                assert ok not in metrics, (name, u, rk)
        a = run.agent
        # ---- gradients (fingerprints) right before each optimizer step
        for tag, grads in (('critic_opt', a.dbg.get('critic_grads')),
                           ('actor_opt', a.dbg.get('actor_grads')),
                           ('cpc_opt', a.dbg.get('cpc_grads'))):
            keys = [k for k in gold.files if k.startswith(p + 'grad/' + tag + '/')]
            if grads is None:
                assert not keys
                continue
            n = 0
            for k in keys:
                pk = k.split('/', 3)[3]
                if pk in ('W', 'log_alpha'):
                    continue
                net, sub = pk.split('.', 1)
                assert fp_close(S.summarize(grads[sub]), gold[k], rtol=1e-3), (name, u, k)
                n += 1
            assert n == len(grads), (name, u, tag, n, len(grads))
        if p + 'grad/cpc_opt/W' in gold.files:
            assert fp_close(S.summarize(a.dbg['W_grad']), gold[p + 'grad/cpc_opt/W'], rtol=1e-3)
        # ---- parameters after the update
        for net, sd in (('actor', a.actor), ('critic', a.critic), ('target', a.target)):
            for k, v in sd.items():
                # Adam at t<=4 is sign-like: allow a couple of lr-sized flips on ~0 grads
                g, o = gold[p + 'param/' + net + '.' + k], S.summarize(v)
                assert np.abs(o[2:] - g[2:]).max() <= 5e-5, (name, u, net, k)
                assert abs(o[1] - g[1]) <= 1e-4 * max(g[1], 1.0), (name, u, net, k)
        assert fp_close(S.summarize(a.W), gold[p + 'param/W'], rtol=1e-4)
        assert abs(float(a.log_alpha.detach()) - gold[p + 'param/log_alpha'][0]) < 1e-9
        # ---- latents
        if 'z_a' in a.dbg:
            assert close(a.dbg['z_a'].numpy(), gold[p + 'out/critic_z'])
            assert close(a.dbg['z_pos'].numpy(), gold[p + 'out/target_z'])
