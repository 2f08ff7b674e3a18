"""Deterministic synthetic scenarios -- TEST INFRASTRUCTURE ONLY (see curla_oracle.py).

Everything is derived from numpy RandomState streams (stable across numpy
versions and platforms), so the golden generator (which drives the real
reference in the build container) and the tests (which drive the oracle and the
CUDA path, possibly on another machine) see bit-identical inputs without
shipping them.  Shapes follow SURVEY.md section 8(d).
"""
import math

import numpy as np
import torch

from . import curla_oracle as O

# name -> config.  'frame_hw' is the stored frame size, 'obs_hw' what the encoder sees.
SCENARIOS = {
    # NB: the reference encoder only works for 76x135 and 90x160 inputs (its
    # square OUT_DIM tables hold ints and fail at encoder.py:66), so every
    # scenario uses the 90x160 stored frame of train.py:45-46.
    # 4 updates cover both step parities (actor/alpha/EMA every 2nd step)
    'identity90x160': dict(aug='identity', frame_hw=(90, 160), B=4, capacity=16, hidden=64,
                           steps=[0, 1, 2, 3], only_cpc=[False] * 4, pixel_sac=False,
                           detach_encoder=False),
    # the headline geometry (90x160 -> random crop 76x135)
    'crop90x160': dict(aug='random_crop', frame_hw=(90, 160), B=4, capacity=16, hidden=64,
                       steps=[0, 1], only_cpc=[False, False], pixel_sac=False,
                       detach_encoder=False),
    'pixelsac90x160': dict(aug='identity', frame_hw=(90, 160), B=4, capacity=16, hidden=64,
                           steps=[0, 1], only_cpc=[False, False], pixel_sac=True,
                           detach_encoder=False),
    'detach_onlycpc_crop': dict(aug='random_crop', frame_hw=(90, 160), B=6, capacity=16,
                                hidden=32, steps=[0, 1, 2], only_cpc=[False, True, False],
                                pixel_sac=False, detach_encoder=True),
}

ACTION_DIM = 2
FRAME_C = 9
FEATURE_DIM = 50
# train.py defaults (train.py:62-104)
HP = dict(discount=0.99, init_temperature=0.1, alpha_lr=1e-4, alpha_beta=0.5, actor_lr=1e-3,
          actor_beta=0.9, actor_log_std_min=-10, actor_log_std_max=2, actor_update_freq=2,
          critic_lr=1e-3, critic_beta=0.9, critic_tau=0.01, critic_target_update_freq=2,
          encoder_feature_dim=FEATURE_DIM, encoder_lr=1e-3, encoder_tau=0.05, num_layers=4,
          num_filters=32, cpc_update_freq=1)


def obs_hw(cfg):
    if cfg['aug'] == 'random_crop':
        return O.crop_output_shape(cfg['frame_hw'])
    return tuple(cfg['frame_hw'])


def make_replay_arrays(capacity, frame_hw, seed=1):
    """RandomState(seed) fill of a FULL buffer."""
    rs = np.random.RandomState(seed)
    h, w = frame_hw
    obses = rs.randint(0, 256, size=(capacity, FRAME_C, h, w), dtype=np.uint8)
    next_obses = rs.randint(0, 256, size=(capacity, FRAME_C, h, w), dtype=np.uint8)
    actions = rs.uniform(-1, 1, size=(capacity, ACTION_DIM)).astype(np.float32)
    rewards = rs.standard_normal(size=(capacity, 1)).astype(np.float32)
    not_dones = (rs.uniform(size=(capacity, 1)) > 0.01).astype(np.float32)
    return obses, next_obses, actions, rewards, not_dones


def _normal(rs, shape, std):
    return torch.from_numpy((rs.standard_normal(size=shape) * std).astype(np.float32))


def make_state_dicts(obs_shape, hidden, seed=2, num_filters=32, num_layers=4,
                     feature_dim=FEATURE_DIM, action_dim=ACTION_DIM):
    """Dense random weights (He-style scale) so every conv tap matters.

    Returns (actor_sd, critic_sd, W) with the reference's key names.  The
    actor's conv entries equal the critic's (they are tied in the agent)."""
    rs = np.random.RandomState(seed)
    oh, ow = O.OUT_DIMS[tuple(obs_shape[1:])][num_layers]

    def enc():
        d = {}
        cin = obs_shape[0]
        for i in range(num_layers):
            d['encoder.convs.%d.weight' % i] = _normal(rs, (num_filters, cin, 3, 3),
                                                       math.sqrt(2.0 / (cin * 9)))
            d['encoder.convs.%d.bias' % i] = _normal(rs, (num_filters,), 0.05)
            cin = num_filters
        fin = num_filters * oh * ow
        d['encoder.fc.weight'] = _normal(rs, (feature_dim, fin), math.sqrt(1.0 / fin))
        d['encoder.fc.bias'] = _normal(rs, (feature_dim,), 0.05)
        d['encoder.ln.weight'] = 1.0 + _normal(rs, (feature_dim,), 0.1)
        d['encoder.ln.bias'] = _normal(rs, (feature_dim,), 0.1)
        return d

    def mlp(prefix, i, o):
        return {prefix + '0.weight': _normal(rs, (hidden, i), math.sqrt(2.0 / i)),
                prefix + '0.bias': _normal(rs, (hidden,), 0.05),
                prefix + '2.weight': _normal(rs, (hidden, hidden), math.sqrt(2.0 / hidden)),
                prefix + '2.bias': _normal(rs, (hidden,), 0.05),
                prefix + '4.weight': _normal(rs, (o, hidden), math.sqrt(1.0 / hidden)),
                prefix + '4.bias': _normal(rs, (o,), 0.05)}

    critic = enc()
    critic.update(mlp('Q1.trunk.', feature_dim + action_dim, 1))
    critic.update(mlp('Q2.trunk.', feature_dim + action_dim, 1))
    actor = enc()
    for k in list(actor):
        if k.startswith('encoder.convs.'):
            actor[k] = critic[k].clone()
    actor.update(mlp('trunk.', feature_dim, 2 * action_dim))
    W = torch.from_numpy(rs.uniform(size=(feature_dim, feature_dim)).astype(np.float32))
    return actor, critic, W


def make_noise(n_updates, B, seed=3, action_dim=ACTION_DIM):
    """Policy noise per update: [u][0] feeds actor(next_obs) in update_critic
    (curl_sac.py:351), [u][1] feeds actor(obs) in update_actor_and_alpha (:375)."""
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.standard_normal(size=(n_updates, 2, B, action_dim))
                            .astype(np.float32))


SAMPLING_SEED = 7   # np.random.seed() before the first sample_cpc of a scenario


def summarize(t, n=64):
    """Compact fingerprint of a tensor: (sum, abs-sum, strided sample)."""
    t = t.detach().double().reshape(-1)
    step = max(1, t.numel() // n)
    return np.concatenate([[float(t.sum()), float(t.abs().sum())], t[::step][:n].numpy()])


class OracleRun:
    """Drives OracleAgent through a scenario exactly like make_golden.py drives
    the reference: same replay arrays, state-dicts, numpy sampling stream, noise."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.hw = tuple(cfg['frame_hw'])
        self.ohw = obs_hw(cfg)
        self.arrays = make_replay_arrays(cfg['capacity'], self.hw)
        self.obs_shape = (FRAME_C, *self.ohw)
        self.agent = O.OracleAgent(self.obs_shape, ACTION_DIM, hidden_dim=cfg['hidden'],
                                   detach_encoder=cfg['detach_encoder'],
                                   pixel_sac=cfg['pixel_sac'], **HP)
        self.state_dicts = make_state_dicts(self.obs_shape, cfg['hidden'])
        self.agent.load_state(*self.state_dicts)
        self.noise = make_noise(len(cfg['steps']), cfg['B'])
        self.u = 0
        np.random.seed(SAMPLING_SEED)

    def sample(self):
        """Oracle restatement of ReplayBuffer.sample_cpc (utils.py:144-187)."""
        cfg = self.cfg
        obses, next_obses, actions, rewards, not_dones = self.arrays
        d = O.draw_sample_indices(cfg['capacity'], 0, True, cfg['B'], cfg['aug'], self.hw,
                                  self.ohw)
        idxs = d['idxs']
        if cfg['aug'] == 'random_crop':
            obs = O.gather_crop(obses, idxs, d['h1_obs'], d['w1_obs'], self.ohw)
            nxt = O.gather_crop(next_obses, idxs, d['h1_next'], d['w1_next'], self.ohw)
            pos = O.gather_crop(obses, idxs, d['h1_pos'], d['w1_pos'], self.ohw)
        else:
            obs, nxt = O.gather(obses, idxs), O.gather(next_obses, idxs)
            pos = obs.copy()
        batch = dict(obs=obs, next=nxt, pos=pos, action=actions[idxs], reward=rewards[idxs],
                     not_done=not_dones[idxs])
        return d, batch

    def step(self):
        """One update; returns (draws, uint8 batch, metrics)."""
        cfg, u = self.cfg, self.u
        d, b = self.sample()
        f = lambda a: torch.from_numpy(a).float()
        m = self.agent.update(f(b['obs']), torch.from_numpy(b['action']),
                              torch.from_numpy(b['reward']), f(b['next']),
                              torch.from_numpy(b['not_done']), f(b['pos']), cfg['steps'][u],
                              self.noise[u, 0], self.noise[u, 1], only_cpc=cfg['only_cpc'][u])
        self.u += 1
        return d, b, dict(m)
