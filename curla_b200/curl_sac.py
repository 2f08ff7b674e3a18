"""Drop-in for the reference's curl_sac.py (curl_sac.py:20-465).

Same public classes and signatures: CurlSacAgent (constructor kwargs, update,
sample_action, select_action, train, save, load, alpha, .actor/.critic/.critic_target/
.CURL/.log_alpha), Actor, Critic, QFunction, CURL, gaussian_logprob, squash, weight_init.
The maths runs in hand-written sm_100a kernels behind the C ABI (include/curla_b200.h);
this file is host-side plumbing: it draws the replay indices exactly like the reference,
hands pointers to the engine and logs the eight training scalars.  There is no CPU or
PyTorch-eager fallback.
"""
import ctypes as C
import math
import os
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import augmentations
from . import dp
from . import encoder
from . import utils
from .engine import Engine

LOG_FREQ = 25_000


def gaussian_logprob(noise, log_std):
    """Compute Gaussian log probability (curl_sac.py:20-23)."""
    residual = (-0.5 * noise.pow(2) - log_std).sum(-1, keepdim=True)
    return residual - 0.5 * np.log(2 * np.pi) * noise.size(-1)


def squash(mu, pi, log_pi):
    """curl_sac.py:26-35 (tensor helper kept for API parity; the agent uses K10)."""
    mu = torch.tanh(mu)
    if pi is not None:
        pi = torch.tanh(pi)
    if log_pi is not None:
        log_pi = log_pi - torch.log(F.relu(1 - pi.pow(2)) + 1e-6).sum(-1, keepdim=True)
    return mu, pi, log_pi


def weight_init(m):
    """Custom weight init for Conv2D and Linear layers (curl_sac.py:38-54)."""
    if isinstance(m, nn.Linear):
        nn.init.orthogonal_(m.weight.data)
        if m.bias is not None:
            m.bias.data.fill_(0.0)
    elif isinstance(m, nn.Conv2d) or isinstance(m, nn.ConvTranspose2d):
        assert m.weight.size(2) == m.weight.size(3)
        m.weight.data.fill_(0.0)
        if m.bias is not None:
            m.bias.data.fill_(0.0)
        mid = m.weight.size(2) // 2
        gain = nn.init.calculate_gain('relu')
        nn.init.orthogonal_(m.weight.data[:, :, mid, mid], gain)


# ---------------------------------------------------------------------------------
# construction-time initial weights: the same nn.Module constructors in the same order as
# the reference, on CPU, so the torch RNG stream (and therefore every initial weight) is
# identical for a given seed.  Not on the hot path.
# ---------------------------------------------------------------------------------
class _InitEncoder(nn.Module):
    def __init__(self, obs_shape, feature_dim, num_layers, num_filters):
        super().__init__()
        out_dim = encoder.out_dim_for(obs_shape, num_layers)
        self.convs = nn.ModuleList([nn.Conv2d(obs_shape[0], num_filters, 3, stride=2)])
        for _ in range(num_layers - 1):
            self.convs.append(nn.Conv2d(num_filters, num_filters, 3, stride=1))
        self.fc = nn.Linear(num_filters * out_dim[0] * out_dim[1], feature_dim)
        self.ln = nn.LayerNorm(feature_dim)


class _InitActor(nn.Module):
    def __init__(self, obs_shape, action_shape, hidden_dim, feature_dim, num_layers, num_filters):
        super().__init__()
        self.encoder = _InitEncoder(obs_shape, feature_dim, num_layers, num_filters)
        self.trunk = nn.Sequential(nn.Linear(feature_dim, hidden_dim), nn.ReLU(),
                                   nn.Linear(hidden_dim, hidden_dim), nn.ReLU(),
                                   nn.Linear(hidden_dim, 2 * action_shape[0]))
        self.apply(weight_init)


class _InitQ(nn.Module):
    def __init__(self, obs_dim, action_dim, hidden_dim):
        super().__init__()
        self.trunk = nn.Sequential(nn.Linear(obs_dim + action_dim, hidden_dim), nn.ReLU(),
                                   nn.Linear(hidden_dim, hidden_dim), nn.ReLU(),
                                   nn.Linear(hidden_dim, 1))


class _InitCritic(nn.Module):
    def __init__(self, obs_shape, action_shape, hidden_dim, feature_dim, num_layers, num_filters):
        super().__init__()
        self.encoder = _InitEncoder(obs_shape, feature_dim, num_layers, num_filters)
        self.Q1 = _InitQ(feature_dim, action_shape[0], hidden_dim)
        self.Q2 = _InitQ(feature_dim, action_shape[0], hidden_dim)
        self.apply(weight_init)


def initial_state(obs_shape, action_shape, hidden_dim, feature_dim, num_layers=4, num_filters=32):
    """(actor_sd, critic_sd, W) drawn from the torch global RNG in the reference constructor's
    order (curl_sac.py:275-317): Actor, Critic, Critic (target: its draws are discarded by the
    copy at :287 but advance the stream), then CURL.W = torch.rand (:192).  The actor's conv
    entries are the critic's (tie at :290).  CPU only; tests/test_capi_cpu.py pins it against
    fingerprints of the reference's own freshly constructed agent (tests/golden/init_seed0.npz)."""
    actor0 = _InitActor(obs_shape, action_shape, hidden_dim, feature_dim, num_layers, num_filters)
    critic0 = _InitCritic(obs_shape, action_shape, hidden_dim, feature_dim, num_layers, num_filters)
    _InitCritic(obs_shape, action_shape, hidden_dim, feature_dim, num_layers, num_filters)
    W = torch.rand(feature_dim, feature_dim)
    critic_sd = {k: v.detach().clone() for k, v in critic0.state_dict().items()}
    actor_sd = {k: v.detach().clone() for k, v in actor0.state_dict().items()}
    actor_sd.update({k: v.clone() for k, v in critic_sd.items() if k.startswith('encoder.convs.')})
    return actor_sd, critic_sd, W


# ---------------------------------------------------------------------------------
# module-like views on the engine's tensors
# ---------------------------------------------------------------------------------
class _Linear(object):
    """nn.Linear look-alike on engine tensors.  `.weight.grad` / `.bias.grad` (what
    Logger.log_param reads, logger.py:155-162) are views of the gradient bucket of the last update."""

    def __init__(self, host, prefix):
        self._host, self._prefix = host, prefix

    def _param(self, name):
        eng = self._host.engine
        t = eng.t[self._prefix + name]
        g = eng.grad_view(self._prefix + name)
        if g is not None:
            t.grad = g
        return t

    @property
    def weight(self):
        return self._param('weight')

    @property
    def bias(self):
        return self._param('bias')


class _Trunk(object):
    """nn.Sequential(Linear, ReLU, Linear, ReLU, Linear) look-alike: trunk[0], [2], [4]."""

    def __init__(self, host, prefix):
        self._host, self._prefix = host, prefix
        self._layers = {i: _Linear(host, '%s%d.' % (prefix, i)) for i in (0, 2, 4)}

    def __getitem__(self, i):
        return self._layers[i]

    def keys(self):
        return ['%d.%s' % (i, k) for i in (0, 2, 4) for k in ('weight', 'bias')]


class _NetBase(object):
    training = True

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def to(self, device):
        return self

    def _items(self):
        raise NotImplementedError

    def parameters(self):
        return [t for _, t in self._items()]

    def _after_param_write(self):
        self._host.engine.refresh_shadows()


def _mlp_state(host, prefix, out_prefix, sd):
    for k in _Trunk(host, prefix).keys():
        sd[out_prefix + k] = host.engine.t[prefix + k].clone()


def _mlp_load(host, prefix, in_prefix, sd):
    for k in _Trunk(host, prefix).keys():
        host.engine.t[prefix + k].copy_(sd[in_prefix + k])


class Actor(_NetBase):
    """MLP actor network (curl_sac.py:57-121)."""

    def __init__(self, host, obs_shape, action_shape, hidden_dim, encoder_feature_dim, log_std_min,
                 log_std_max, num_layers, num_filters):
        self._host = host
        self.encoder = encoder.CNNEncoder(obs_shape, encoder_feature_dim, num_layers, num_filters,
                                          output_logits=True, _host=host, _net=0, _prefix='actor.encoder.')
        self.log_std_min, self.log_std_max = log_std_min, log_std_max
        self.trunk = _Trunk(host, 'actor.trunk.')
        self.outputs = dict()

    def forward(self, obs, compute_pi=True, compute_log_pi=True, detach_encoder=False):
        z = self.encoder(obs, detach=detach_encoder)
        mu, pi, log_pi, log_std = self._host.actor_head(z, compute_pi, compute_log_pi)
        self.outputs['mu'] = mu
        self.outputs['std'] = log_std.exp()
        return mu, pi, log_pi, log_std

    __call__ = forward

    def _items(self):
        t = self._host.engine.t
        items = [('encoder.' + k, v) for k, v in zip(
            ['convs.%d.%s' % (i, s) for i in range(4) for s in ('weight', 'bias')] +
            ['fc.weight', 'fc.bias', 'ln.weight', 'ln.bias'], self.encoder.parameters())]
        return items + [('trunk.' + k, t['actor.trunk.' + k]) for k in self.trunk.keys()]

    def state_dict(self):
        sd = self.encoder.state_dict('encoder.')
        _mlp_state(self._host, 'actor.trunk.', 'trunk.', sd)
        return sd

    def load_state_dict(self, sd):
        self.encoder.load_state_dict(sd, 'encoder.', refresh=False)
        _mlp_load(self._host, 'actor.trunk.', 'trunk.', sd)
        self._after_param_write()

    def log(self, L, step, log_freq=LOG_FREQ):
        if step % log_freq != 0:
            return
        for k, v in self.outputs.items():
            L.log_histogram('train_actor/%s_hist' % k, v, step)
        L.log_param('train_actor/fc1', self.trunk[0], step)
        L.log_param('train_actor/fc2', self.trunk[2], step)
        L.log_param('train_actor/fc3', self.trunk[4], step)


class QFunction(_NetBase):
    """MLP for q-function (curl_sac.py:124-139)."""

    def __init__(self, host, prefix):
        self._host, self._prefix = host, prefix
        self.trunk = _Trunk(host, prefix + 'trunk.')

    def _items(self):
        return [('trunk.' + k, self._host.engine.t[self._prefix + 'trunk.' + k]) for k in self.trunk.keys()]


class Critic(_NetBase):
    """Critic network, employs two Q-functions (curl_sac.py:142-180)."""

    def __init__(self, host, name, net, obs_shape, action_shape, hidden_dim, encoder_feature_dim,
                 num_layers, num_filters):
        self._host, self._name, self._net = host, name, net
        self.encoder = encoder.CNNEncoder(obs_shape, encoder_feature_dim, num_layers, num_filters,
                                          output_logits=True, _host=host, _net=net,
                                          _prefix=name + '.encoder.')
        self.Q1 = QFunction(host, name + '.Q1.')
        self.Q2 = QFunction(host, name + '.Q2.')
        self.outputs = dict()

    def forward(self, obs, action, detach_encoder=False):
        z = self.encoder(obs, detach=detach_encoder)
        assert z.size(0) == action.size(0)
        q1, q2 = self._host.q_heads(self._net, z, action)
        self.outputs['q1'] = q1
        self.outputs['q2'] = q2
        return q1, q2

    __call__ = forward

    def _items(self):
        items = [('encoder.' + k, v) for k, v in zip(
            ['convs.%d.%s' % (i, s) for i in range(4) for s in ('weight', 'bias')] +
            ['fc.weight', 'fc.bias', 'ln.weight', 'ln.bias'], self.encoder.parameters())]
        return items + [('Q1.' + k, v) for k, v in self.Q1._items()] + \
            [('Q2.' + k, v) for k, v in self.Q2._items()]

    def state_dict(self):
        sd = self.encoder.state_dict('encoder.')
        _mlp_state(self._host, self._name + '.Q1.trunk.', 'Q1.trunk.', sd)
        _mlp_state(self._host, self._name + '.Q2.trunk.', 'Q2.trunk.', sd)
        return sd

    def load_state_dict(self, sd):
        self.encoder.load_state_dict(sd, 'encoder.', refresh=False)
        _mlp_load(self._host, self._name + '.Q1.trunk.', 'Q1.trunk.', sd)
        _mlp_load(self._host, self._name + '.Q2.trunk.', 'Q2.trunk.', sd)
        self._after_param_write()

    def log(self, L, step, log_freq=LOG_FREQ):
        if step % log_freq != 0:
            return
        for k, v in self.outputs.items():
            L.log_histogram('train_critic/%s_hist' % k, v, step)
        for i in range(3):
            L.log_param('train_critic/q1_fc%d' % i, self.Q1.trunk[i * 2], step)
            L.log_param('train_critic/q2_fc%d' % i, self.Q2.trunk[i * 2], step)


class CURL(_NetBase):
    """CURL head (curl_sac.py:183-222): shares critic.encoder / critic_target.encoder, owns W."""

    def __init__(self, host, obs_shape, z_dim, critic, critic_target, output_type="continuous"):
        self._host = host
        self.encoder = critic.encoder
        self.encoder_target = critic_target.encoder
        self.output_type = output_type

    @property
    def W(self):
        return self._host.engine.t['CURL.W']

    def encode(self, x, detach=False, ema=False):
        return self.encoder_target(x) if ema else self.encoder(x)

    def compute_logits(self, z_a, z_pos):
        """(B,B) logits z_a (W z_pos^T) - rowmax (curl_sac.py:211-222); tensor helper for
        callers outside the update (the update itself uses the fused K12 kernels)."""
        Wz = torch.matmul(self.W, z_pos.T)
        logits = torch.matmul(z_a, Wz)
        return logits - torch.max(logits, 1)[0][:, None]

    def _items(self):
        return [('W', self.W)] + [('encoder.' + k, v) for k, v in zip(range(12), self.encoder.parameters())]

    def state_dict(self):
        sd = {'W': self.W.clone()}
        sd.update(self.encoder.state_dict('encoder.'))
        sd.update(self.encoder_target.state_dict('encoder_target.'))
        return sd

    def load_state_dict(self, sd):
        self.W.copy_(sd['W'])
        self.encoder.load_state_dict(sd, 'encoder.', refresh=False)
        self.encoder_target.load_state_dict(sd, 'encoder_target.', refresh=False)
        self._after_param_write()


# ---------------------------------------------------------------------------------
class _Host(object):
    """Owns the Engine and the small inference helpers shared by the module views."""

    engine = None

    def _stage(self, obs, b0, b1):
        eng = self.engine
        c = eng.cfg
        x = obs[b0:b1].contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        s2d = eng.t['s2d.next']
        with torch.cuda.device(eng.device):
            _lib.call('curla_f32_to_s2d', _lib.ptr(x), c.C, c.H, c.W, b1 - b0, s2d.shape[1],
                      (s2d.shape[0] // c.batch) * s2d.shape[1], _lib.ptr(s2d), eng.stream())
        return s2d

    INFER_BATCH = 512     # engine batch used for inference-only agents (eval.py, latent tools)

    def _grow_for_inference(self, n):
        """An agent that has not trained yet (eval.py / plot_tsne/*.py never call update) gets an
        engine large enough to run batched inference in few launches; once training started the
        engine batch is the replay batch and larger inputs are processed in chunks of it."""
        c = self.engine.cfg
        target = min(int(n), self.INFER_BATCH)
        if target > c.batch and getattr(self, '_update_count', 1) == 0 and hasattr(self, '_make_engine'):
            self._make_engine(batch=target, frame_hw=(c.Hf, c.Wf))

    def encode(self, net, obs, apply_tanh=False):
        obs = torch.as_tensor(obs, device=self.engine.device)
        if obs.dim() == 3:
            obs = obs.unsqueeze(0)
        self._grow_for_inference(obs.shape[0])
        eng = self.engine
        c = eng.cfg
        assert tuple(obs.shape[1:]) == (c.C, c.H, c.W), 'encoder input %s != %s' % (
            tuple(obs.shape[1:]), (c.C, c.H, c.W))
        n = obs.shape[0]
        out = torch.zeros((n, 64), dtype=torch.float32, device=eng.device)
        for b0 in range(0, n, c.batch):
            b1 = min(n, b0 + c.batch)
            s2d = self._stage(obs, b0, b1)
            with torch.cuda.device(eng.device):
                _lib.check(eng.lib.curla_agent_encode(eng.h, net, _lib.ptr(s2d), b1 - b0, int(apply_tanh),
                                                      _lib.ptr(out[b0:b1]), eng.stream()), 'encode')
        return out[:, :c.feature_dim]

    def encode_frames(self, net, frames, top=0, left=0, apply_tanh=False):
        """Encoder forward straight from uint8 frames (N, C, Hf, Wf) resident on the device: the
        fused gather+crop kernel takes the (top, left) window of every frame (the center crop of
        augmentations.py:37-43 when the stored frames are larger than the encoder input), so no
        float copy of the observations is ever made.  Batched form of latent_data.py:74-83."""
        frames = torch.as_tensor(frames)
        assert frames.dtype == torch.uint8 and frames.dim() == 4
        self._grow_for_inference(frames.shape[0])
        eng = self.engine
        c = eng.cfg
        frames = frames.to(eng.device).contiguous()
        n, C_, Hf, Wf = frames.shape
        assert C_ == c.C and top + c.H <= Hf and left + c.W <= Wf, 'frames %s do not contain a %dx%d window at (%d, %d)' % (
            tuple(frames.shape), c.H, c.W, top, left)
        out = torch.zeros((n, 64), dtype=torch.float32, device=eng.device)
        s2d = eng.t['s2d.next']
        stride = (s2d.shape[0] // c.batch) * s2d.shape[1]
        idx = torch.arange(n, dtype=torch.int64, device=eng.device)
        h1 = torch.full((c.batch,), int(top), dtype=torch.int64, device=eng.device)
        w1 = torch.full((c.batch,), int(left), dtype=torch.int64, device=eng.device)
        for b0 in range(0, n, c.batch):
            b1 = min(n, b0 + c.batch)
            with torch.cuda.device(eng.device):
                _lib.call('curla_gather_crop_s2d', _lib.ptr(frames), c.C, Hf, Wf, _lib.ptr(idx[b0:b1]), _lib.ptr(h1),
                          _lib.ptr(w1), b1 - b0, c.H, c.W, s2d.shape[1], stride, _lib.ptr(s2d), eng.stream())
                _lib.check(eng.lib.curla_agent_encode(eng.h, net, _lib.ptr(s2d), b1 - b0, int(apply_tanh),
                                                      _lib.ptr(out[b0:b1]), eng.stream()), 'encode')
        return out[:, :c.feature_dim]

    def _pad64(self, z):
        eng = self.engine
        zp = torch.zeros((z.shape[0], 64), dtype=torch.float32, device=eng.device)
        zp[:, :z.shape[1]] = z
        return zp

    def actor_head(self, z, compute_pi=True, compute_log_pi=True, noise=None):
        eng = self.engine
        c = eng.cfg
        n, A = z.shape[0], c.action_dim
        zp = self._pad64(z)
        mu = torch.empty((n, A), dtype=torch.float32, device=eng.device)
        ls = torch.empty_like(mu)
        pi = torch.empty_like(mu) if compute_pi else None
        log_pi = torch.empty((n, 1), dtype=torch.float32, device=eng.device) if (compute_pi and compute_log_pi) else None
        self._rng_offset = getattr(self, '_rng_offset', 0) + 1
        for b0 in range(0, n, c.batch):
            b1 = min(n, b0 + c.batch)
            with torch.cuda.device(eng.device):
                _lib.check(eng.lib.curla_agent_actor_head(
                    eng.h, _lib.ptr(zp[b0:b1]), b1 - b0, _lib.ptr(noise[b0:b1].contiguous()) if noise is not None else None,
                    self._seed(), (1 << 40) + self._rng_offset * 65536 + b0, int(compute_pi), int(compute_log_pi),
                    _lib.ptr(mu[b0:b1]), _lib.ptr(pi[b0:b1]) if pi is not None else None,
                    _lib.ptr(log_pi[b0:b1]) if log_pi is not None else None, _lib.ptr(ls[b0:b1]), eng.stream()),
                    'actor_head')
        return mu, pi, log_pi, ls

    def q_heads(self, net, z, action):
        eng = self.engine
        c = eng.cfg
        n = z.shape[0]
        zp = self._pad64(z)
        action = action.to(eng.device, torch.float32).contiguous()
        q1 = torch.empty((n, 1), dtype=torch.float32, device=eng.device)
        q2 = torch.empty_like(q1)
        for b0 in range(0, n, c.batch):
            b1 = min(n, b0 + c.batch)
            with torch.cuda.device(eng.device):
                _lib.check(eng.lib.curla_agent_q_heads(eng.h, net, _lib.ptr(zp[b0:b1]), _lib.ptr(action[b0:b1]),
                                                       b1 - b0, _lib.ptr(q1[b0:b1]), _lib.ptr(q2[b0:b1]),
                                                       eng.stream()), 'q_heads')
        return q1, q2

    def _seed(self):
        return getattr(self, '_noise_seed', 0x5eed)


class _ActionGraph(object):
    """The B=1 action path of train.py:418 / eval.py:78 (curl_sac.py:330-347) as ONE captured CUDA
    graph with pinned-host input and output: H2D of the observation (+ policy noise), the fused
    gather/center-crop -> space-to-depth kernel, the actor encoder (4 tcgen05 conv launches, fc
    split-K GEMM, LayerNorm), the trunk + tanh-Gaussian head, D2H of (mu, pi).  One graph launch
    and one stream sync per environment step instead of ~15 launches and three blocking copies.
    The graph replays on the caller's current stream, so it is ordered after any update in flight
    (it shares the engine's inference scratch buffers)."""

    def __init__(self, agent, in_shape, in_dtype, top, left, sample):
        eng = agent.engine
        c = eng.cfg
        dev = eng.device
        A = c.action_dim
        self.engine, self.A, self.sample = eng, A, sample
        self.h_in = torch.empty(in_shape, dtype=in_dtype).pin_memory()
        self.h_noise = torch.zeros((1, A), dtype=torch.float32).pin_memory()
        self.h_out = torch.zeros((2, A), dtype=torch.float32).pin_memory()
        self.d_in = torch.zeros(in_shape, dtype=in_dtype, device=dev)
        self.d_noise = torch.zeros((1, A), dtype=torch.float32, device=dev)
        self.d_out = torch.zeros((2, A), dtype=torch.float32, device=dev)     # row 0 = mu, row 1 = pi
        self.d_ls = torch.zeros((1, A), dtype=torch.float32, device=dev)
        self.d_z = torch.zeros((1, 64), dtype=torch.float32, device=dev)
        self.idx = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.h1 = torch.full((1,), int(top), dtype=torch.int64, device=dev)
        self.w1 = torch.full((1,), int(left), dtype=torch.int64, device=dev)
        s2d = eng.t['s2d.next']
        self.s2d, self.stride = s2d, (s2d.shape[0] // c.batch) * s2d.shape[1]
        self._run()                                   # eager warm-up (function attributes, lazy loading)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()

    def _run(self):
        eng = self.engine
        c = eng.cfg
        self.d_in.copy_(self.h_in, non_blocking=True)
        self.d_noise.copy_(self.h_noise, non_blocking=True)
        with torch.cuda.device(eng.device):
            if self.d_in.dtype == torch.uint8:
                _lib.call('curla_gather_crop_s2d', _lib.ptr(self.d_in), c.C, self.d_in.shape[1], self.d_in.shape[2],
                          _lib.ptr(self.idx), _lib.ptr(self.h1), _lib.ptr(self.w1), 1, c.H, c.W, self.s2d.shape[1],
                          self.stride, _lib.ptr(self.s2d), eng.stream())
            else:
                _lib.call('curla_f32_to_s2d', _lib.ptr(self.d_in), c.C, c.H, c.W, 1, self.s2d.shape[1], self.stride,
                          _lib.ptr(self.s2d), eng.stream())
            _lib.check(eng.lib.curla_agent_encode(eng.h, 0, _lib.ptr(self.s2d), 1, 0, _lib.ptr(self.d_z), eng.stream()),
                       'encode')
            _lib.check(eng.lib.curla_agent_actor_head(
                eng.h, _lib.ptr(self.d_z), 1, _lib.ptr(self.d_noise), 0, 0, int(self.sample), 0, _lib.ptr(self.d_out[0]),
                _lib.ptr(self.d_out[1]) if self.sample else None, None, _lib.ptr(self.d_ls), eng.stream()), 'actor_head')
        self.h_out.copy_(self.d_out, non_blocking=True)

    def __call__(self, obs):
        np.copyto(self.h_in.numpy(), obs, casting='unsafe')
        if self.sample:
            torch.randn(self.h_noise.shape, out=self.h_noise)       # fresh policy noise (curl_sac.py:97)
        self.graph.replay()
        torch.cuda.current_stream(self.engine.device).synchronize()
        return self.h_out[1 if self.sample else 0].numpy().copy()


class _StandaloneEncoderHost(_Host):
    """Private engine behind a bare CNNEncoder(...) (encoder.py:132-169 style use)."""

    def __init__(self, enc):
        dev = torch.device('cuda')
        c, h, w = enc.obs_shape
        self.engine = Engine(dev, C=c, H=h, W=w, Hf=h, Wf=w, feature_dim=enc.feature_dim, hidden_dim=8,
                             action_dim=2, num_filters=enc.num_filters, num_layers=enc.num_layers, batch=64,
                             global_batch=64, rank=0, world=1, actor_update_freq=1,
                             critic_target_update_freq=1, cpc_update_freq=1)
        init = _InitEncoder(enc.obs_shape, enc.feature_dim, enc.num_layers, enc.num_filters)
        sd = init.state_dict()
        t = self.engine.t
        for i in range(enc.num_layers):
            t['critic.encoder.convs.%d.weight' % i].copy_(sd['convs.%d.weight' % i])
            t['critic.encoder.convs.%d.bias' % i].copy_(sd['convs.%d.bias' % i])
        self.engine.fc_from_torch(sd['fc.weight'].to(dev), t['critic.encoder.fc.weight_canon'])
        for k in ('fc.bias', 'ln.weight', 'ln.bias'):
            t['critic.encoder.' + k].copy_(sd[k])
        self.engine.refresh_shadows()


class CurlSacAgent(_Host):
    """CURL representation learning with SAC (curl_sac.py:224-465)."""

    def __init__(
        self,
        obs_shape,
        action_shape,
        device,
        augmentor,
        hidden_dim=256,
        discount=0.99,
        init_temperature=0.01,
        alpha_lr=1e-3,
        alpha_beta=0.9,
        actor_lr=1e-3,
        actor_beta=0.9,
        actor_log_std_min=-10,
        actor_log_std_max=2,
        actor_update_freq=2,
        critic_lr=1e-3,
        critic_beta=0.9,
        critic_tau=0.005,
        critic_target_update_freq=2,
        encoder_feature_dim=50,
        encoder_lr=1e-3,
        encoder_tau=0.005,
        num_layers=4,
        num_filters=32,
        cpc_update_freq=1,
        log_interval=100,
        log_param_hist_imgs=False,
        detach_encoder=False,
        pixel_sac=False
    ):
        self.augmentor = augmentor
        self.device = torch.device(device)
        self.discount = discount
        self.critic_tau = critic_tau
        self.encoder_tau = encoder_tau
        self.actor_update_freq = actor_update_freq
        self.critic_target_update_freq = critic_target_update_freq
        self.cpc_update_freq = cpc_update_freq
        self.log_interval = log_interval
        self.log_param_hist_imgs = log_param_hist_imgs
        self.image_shape = tuple(obs_shape[-2:])
        self.detach_encoder = detach_encoder
        self.pixel_sac = pixel_sac
        self.obs_shape = tuple(obs_shape)
        self.action_shape = tuple(action_shape)
        self.target_entropy = -np.prod(action_shape)        # curl_sac.py:296

        # data parallel: one process per GPU; the engine all-reduces grads / all-gathers keys
        self.rank, self.world = dp.world_info()

        self._cfg = dict(
            C=obs_shape[0], H=obs_shape[1], W=obs_shape[2], feature_dim=encoder_feature_dim,
            hidden_dim=hidden_dim, action_dim=action_shape[0], num_filters=num_filters,
            num_layers=num_layers, rank=self.rank, world=self.world,
            detach_encoder=int(detach_encoder), pixel_sac=int(pixel_sac),
            actor_update_freq=actor_update_freq, critic_target_update_freq=critic_target_update_freq,
            cpc_update_freq=cpc_update_freq, discount=discount, critic_tau=critic_tau,
            encoder_tau=encoder_tau, actor_lr=actor_lr, actor_beta=actor_beta, critic_lr=critic_lr,
            critic_beta=critic_beta, alpha_lr=alpha_lr, alpha_beta=alpha_beta, encoder_lr=encoder_lr,
            log_std_min=actor_log_std_min, log_std_max=actor_log_std_max,
            target_entropy=float(self.target_entropy))
        if num_layers != 4 or num_filters != 32:
            raise NotImplementedError('the sm_100a kernels are built for num_layers=4, num_filters=32')
        encoder.out_dim_for(obs_shape, num_layers)     # raises NotImplementedError like the reference

        # initial weights: same constructors / RNG order as the reference (CPU, one-off)
        actor_sd0, critic_sd0, W0 = initial_state(obs_shape, action_shape, hidden_dim, encoder_feature_dim,
                                                  num_layers, num_filters)
        self.engine = None
        self._act_graphs = {}
        self._make_engine(batch=1, frame_hw=self.image_shape)

        self.actor = Actor(self, obs_shape, action_shape, hidden_dim, encoder_feature_dim, actor_log_std_min,
                           actor_log_std_max, num_layers, num_filters)
        self.critic = Critic(self, 'critic', 1, obs_shape, action_shape, hidden_dim, encoder_feature_dim,
                             num_layers, num_filters)
        self.critic_target = Critic(self, 'target', 2, obs_shape, action_shape, hidden_dim,
                                    encoder_feature_dim, num_layers, num_filters)
        self.critic.load_state_dict(critic_sd0)
        self.actor.load_state_dict(actor_sd0)                             # tied convs: curl_sac.py:290
        self.critic_target.load_state_dict(self.critic.state_dict())     # curl_sac.py:287
        self.engine.t['log_alpha'].fill_(float(np.log(init_temperature)))  # curl_sac.py:292
        self.CURL = CURL(self, obs_shape, encoder_feature_dim, self.critic, self.critic_target,
                         output_type='continuous')
        self.engine.t['CURL.W'].copy_(W0)                                 # curl_sac.py:192
        self.engine.refresh_shadows()

        self._update_count = 0
        self._ptr_cache = {}
        self._metrics_host = None
        # Philox key of the policy noise: derived from the torch seed WITHOUT drawing from the global
        # stream, so the stream stands where the reference's does after construction
        self._noise_seed = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + 0x5EED) & ((1 << 63) - 1)
        self._noise_override = None      # tests inject (noise_next, noise_cur)
        self._arange = None
        self.train()
        self.critic_target.train()

    # -- engine lifetime -------------------------------------------------------------
    def _make_engine(self, batch, frame_hw):
        old = self.engine
        cfg = dict(self._cfg)
        cfg.update(batch=batch, global_batch=batch * self.world, Hf=frame_hw[0], Wf=frame_hw[1])
        eng = Engine(self.device, **cfg)
        if old is not None:
            for i in range(4):           # params, shadows, grads, adam: batch-independent layout
                eng.arenas[i].copy_(old.arenas[i])
            for k in ('log_alpha', 'adam.log_alpha'):
                eng.t[k].copy_(old.t[k])
            # the Adam step counters live on the host side of the engine: without them the new engine
            # would restart bias correction at t = 1 on warm moments
            steps = (C.c_int * 4)()
            _lib.check(eng.lib.curla_agent_get_opt_steps(old.h, steps), 'get_opt_steps')
            _lib.check(eng.lib.curla_agent_set_opt_steps(eng.h, steps[0], steps[1], steps[2], steps[3]), 'set_opt_steps')
        self.engine = eng
        self._engine_gen = getattr(self, '_engine_gen', 0) + 1      # invalidates captured action graphs
        if os.environ.get('CURLA_MAILBOX', '1')[:1] != '0':
            # the logged scalars arrive in 64 bytes of pinned host memory as soon as the update has computed them
            # (engine.cu: curla_publish_metrics; data parallel: their mean over the ranks, one 64-byte ncclAvg inside the
            # update), so logging every step does not wait for the update's tail
            if getattr(self, '_mailbox', None) is None:
                self._mailbox = torch.zeros(16, dtype=torch.float32).pin_memory()
                self._mailbox_f = self._mailbox.numpy()
                self._mailbox_u = self._mailbox_f.view(np.uint32)
            _lib.check(eng.lib.curla_agent_set_mailbox(eng.h, C.c_void_p(self._mailbox.data_ptr())), 'set_mailbox')
        if self.world > 1:
            if old is not None and getattr(self, '_comm_ready', False):
                # same rank, same world: the communicator moves to the new engine (no collective, so a
                # single rank may grow its engine for inference on its own)
                _lib.check(eng.lib.curla_agent_take_comm(eng.h, old.h), 'take_comm')
            else:
                self._init_comm()
                self._comm_ready = True

    def _init_comm(self):
        uid = dp.nccl_unique_id(self.engine.lib) if self.rank == 0 else b''
        raw = dp.broadcast_bytes(uid, 128, self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.engine.lib.curla_agent_init_comm(self.engine.h, C.c_char_p(raw)), 'init_comm')

    def _ensure_engine(self, batch, frame_hw):
        c = self.engine.cfg
        if c.batch != batch or (c.Hf, c.Wf) != tuple(frame_hw):
            if self._update_count > 0 and c.batch != batch:
                raise _lib.CurlaError('batch size changed after training started (%d -> %d)' % (c.batch, batch))
            self._make_engine(batch, frame_hw)

    # -- reference API -----------------------------------------------------------------
    def train(self, training=True):
        self.training = training
        self.actor.train(training)
        self.critic.train(training)
        self.CURL.train(training)

    @property
    def log_alpha(self):
        return self.engine.t['log_alpha'].view(())

    @property
    def alpha(self):
        return self.log_alpha.exp()

    def _act(self, obs, sample, crop):
        """Graph-captured B=1 action path.  uint8 observations (what FrameStack yields) go to the device
        as bytes and are center-cropped by the gather kernel; other dtypes take the float route."""
        obs = np.asarray(obs)
        if crop and tuple(obs.shape[-2:]) != self.image_shape:
            if obs.dtype == np.uint8:
                top, left = (obs.shape[-2] - self.image_shape[0]) // 2, (obs.shape[-1] - self.image_shape[1]) // 2
            else:
                obs, top, left = self.augmentor.evaluation_augmentation(obs), 0, 0      # augmentations.py:37-43
        else:
            top, left = 0, 0
        if obs.dtype != np.uint8:
            obs = obs.astype(np.float32, copy=False)
        assert obs.shape[0] == self.obs_shape[0] and obs.shape[-2] - top >= self.image_shape[0] and \
            obs.shape[-1] - left >= self.image_shape[1], 'observation %s does not fit the encoder input %s' % (
                obs.shape, self.obs_shape)
        if not crop and tuple(obs.shape) != self.obs_shape:
            # select_action does not crop (curl_sac.py:330-336): the reference fails at the fc layer on
            # any other shape; never encode a silent top-left window
            raise ValueError('select_action: observation %s is not the encoder input %s (center-crop it with '
                             'augmentor.evaluation_augmentation first)' % (tuple(obs.shape), self.obs_shape))
        key = (self._engine_gen, obs.shape, obs.dtype.str, top, left, bool(sample))
        g = self._act_graphs.get(key)
        if g is None:
            for k in [k for k in self._act_graphs if k[0] != self._engine_gen]:
                del self._act_graphs[k]                  # engine re-created: its buffers are gone
            in_dtype = torch.uint8 if obs.dtype == np.uint8 else torch.float32
            g = self._act_graphs[key] = _ActionGraph(self, tuple(obs.shape), in_dtype, top, left, sample)
        return g(obs)

    def select_action(self, obs):
        """curl_sac.py:330-336: tanh(mu) of one observation (already encoder-sized)."""
        if os.environ.get('CURLA_NO_GRAPH') == '1':
            with torch.no_grad():
                obs = torch.FloatTensor(np.ascontiguousarray(obs)).to(self.device).unsqueeze(0)
                mu, _, _, _ = self.actor(obs, compute_pi=False, compute_log_pi=False)
                return mu.cpu().data.numpy().flatten()
        return self._act(obs, sample=False, crop=False)

    def sample_action(self, obs):
        """curl_sac.py:338-347: center-crop if needed, then a sample of the tanh-Gaussian policy."""
        if os.environ.get('CURLA_NO_GRAPH') == '1':
            if obs.shape[-2:] != self.image_shape:
                obs = self.augmentor.evaluation_augmentation(obs)
            with torch.no_grad():
                obs = torch.FloatTensor(np.ascontiguousarray(obs)).to(self.device).unsqueeze(0)
                mu, pi, _, _ = self.actor(obs, compute_log_pi=False)
                return pi.cpu().data.numpy().flatten()
        return self._act(obs, sample=True, crop=True)

    def update(self, replay_buffer, L, step, only_cpc=False, _phases=0):
        """One whole SAC+CURL update (curl_sac.py:426-451) as a single engine call.

        `_phases` (tests only) is a CURLA_PHASE_* mask: the parity tests run sample / critic /
        actor / EMA / CPC as separate calls on one staged minibatch (include/curla_b200.h)."""
        if _phases and not (_phases & 1):
            a = self._last_args                  # same minibatch, noise and step as the SAMPLE call
            a.phases = int(_phases)
            self.engine.update(a)
            return
        Bg = replay_buffer.batch_size
        B = dp.local_batch(Bg, self.world)
        a = _lib.UpdateArgs()
        keep = []
        fused = isinstance(replay_buffer, utils.ReplayBuffer) and \
            type(replay_buffer.augmentor) in (augmentations.RandomCrop, augmentations.IdentityAugmentation)
        if fused:
            # sample_cpc's RNG draws (utils.py:147,156-158), gather fused into the update
            self._ensure_engine(B, replay_buffer.obs_shape[1:])
            d, dev = replay_buffer.draw_indices()
            sl = dp.shard_slice(self.rank, self.world, Bg)
            # device pointers of this rank's slices of the [7][Bg] index block (two staging slots),
            # computed once per (buffer, slot): 8 bytes per int64, row stride Bg
            ck = (id(replay_buffer), dev.data_ptr(), sl.start)
            ptrs = self._ptr_cache.get(ck)
            if ptrs is None:
                if len(self._ptr_cache) > 8:
                    self._ptr_cache.clear()
                base = dev.data_ptr() + 8 * sl.start
                ptrs = self._ptr_cache[ck] = (
                    replay_buffer.obses.data_ptr(), replay_buffer.next_obses.data_ptr(), replay_buffer.actions.data_ptr(),
                    replay_buffer.rewards.data_ptr(), replay_buffer.not_dones.data_ptr(),
                    [base + 8 * Bg * r for r in range(7)])
            a.obses, a.next_obses, a.actions, a.rewards, a.not_dones = ptrs[:5]
            rows = ptrs[5]
            a.idxs = rows[0]
            if isinstance(replay_buffer.augmentor, augmentations.RandomCrop):
                a.h1_obs, a.w1_obs, a.h1_next, a.w1_next, a.h1_pos, a.w1_pos = rows[1:7]
                a.pos_is_obs = 0
            else:
                a.pos_is_obs = 1
            reward_for_log = lambda: replay_buffer.rewards[dev[0, sl]].mean()
        else:
            # any other buffer / augmentation: its own sample_cpc() output (float tensors)
            obs, action, reward, next_obs, not_done, kw = replay_buffer.sample_cpc()
            self._ensure_engine(B, self.image_shape)
            sl = dp.shard_slice(self.rank, self.world, Bg)
            f = lambda t: t[sl].to(self.device, torch.float32).contiguous()
            obs_l, next_l, pos_l = f(obs), f(next_obs), f(kw['obs_pos'])
            act_l, rew_l, nd_l = f(action), f(reward), f(not_done)
            if self._arange is None or self._arange.numel() != B:
                self._arange = torch.arange(B, dtype=torch.int64, device=self.device)
            keep += [obs_l, next_l, pos_l, act_l, rew_l, nd_l]
            a.obs_f32, a.next_f32, a.pos_f32 = obs_l.data_ptr(), next_l.data_ptr(), pos_l.data_ptr()
            a.actions, a.rewards, a.not_dones = act_l.data_ptr(), rew_l.data_ptr(), nd_l.data_ptr()
            a.idxs = self._arange.data_ptr()
            a.pos_is_obs = 0
            reward_for_log = lambda: rew_l.mean()
        if self._noise_override is not None:
            n1, n2 = self._noise_override
            n1 = n1.to(self.device, torch.float32).contiguous()
            n2 = n2.to(self.device, torch.float32).contiguous()
            keep += [n1, n2]
            a.noise_next, a.noise_cur = n1.data_ptr(), n2.data_ptr()
        a.seed, a.offset = self._noise_seed, self._update_count
        a.step, a.only_cpc = int(step), int(bool(only_cpc))
        a.phases = int(_phases)
        self._last_args, self._last_keep = a, keep
        self.engine.update(a)
        self._update_count += 1

        if step % self.log_interval == 0:
            published = (not only_cpc) or (not self.pixel_sac and step % self.cpc_update_freq == 0)      # some phase wrote a scalar
            if getattr(self, '_mailbox', None) is not None and not _phases and published and self._noise_override is None:
                # wait for THIS update's sequence word (update count + 1), then read the scalars published with it
                want, mu, t0 = np.uint32((self._update_count) & 0xFFFFFFFF), self._mailbox_u, None
                spins = 0
                while mu[15] != want:
                    spins += 1
                    if spins & 0xFFF == 0:
                        t0 = t0 or time.perf_counter()
                        if time.perf_counter() - t0 > 30.0:
                            torch.cuda.current_stream(self.device).synchronize()      # surfaces a launch error, if any
                            raise _lib.CurlaError('update: the logged scalars of update %d never arrived' % self._update_count)
                m = self._mailbox_f.copy()
            else:
                m = self.engine.t['metrics']
                if self.world > 1 and torch.distributed.get_backend() == 'nccl':
                    m = m.clone()
                    torch.distributed.all_reduce(m)
                    m /= self.world
                if self._metrics_host is None:
                    self._metrics_host = torch.empty(m.shape, dtype=m.dtype).pin_memory()
                self._metrics_host.copy_(m, non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
                m = self._metrics_host.numpy()
            L.log('train/batch_reward', float(m[0]) if not only_cpc else float(reward_for_log()), step)
            if not only_cpc:
                L.log('train_critic/loss', float(m[1]), step)
                if step % self.actor_update_freq == 0:
                    L.log('train_actor/loss', float(m[2]), step)
                    L.log('train_actor/target_entropy', self.target_entropy, step)
                    L.log('train_actor/entropy', float(m[3]), step)
                    L.log('train_alpha/loss', float(m[4]), step)
                    L.log('train_alpha/value', float(m[5]), step)
            if not self.pixel_sac and step % self.cpc_update_freq == 0:
                L.log('train/curl_loss', float(m[6]), step)
        if self.log_param_hist_imgs:
            self._log_taps(L, step, only_cpc)

    def _log_taps(self, L, step, only_cpc):
        """--log_param_hist_imgs (curl_sac.py:370-371, 394-395): critic.log right after the critic step,
        actor.log right after the actor step.  The whole update is ONE engine call here, so both run at
        its end -- with the same content: the Q trunks and their gradients are not touched by the actor
        / CURL phases, nor the actor trunk by the CURL phase; `outputs` holds what the reference's
        modules hold at their log call (Q(obs, action) of update_critic, pre-tanh mu and std of
        actor(obs)).  Afterwards the dicts are left as the reference leaves them at the end of update()."""
        if only_cpc and not (not self.pixel_sac and step % self.cpc_update_freq == 0):
            return
        t = self.engine.t
        A = self.action_shape[0]
        did_actor = (not only_cpc) and step % self.actor_update_freq == 0
        if not only_cpc:
            self.critic.outputs['q1'] = t['p3.q1.out'].clone()
            self.critic.outputs['q2'] = t['p3.q2.out'].clone()
            self.critic.log(L, step)
            if did_actor:
                self.actor.outputs['mu'] = t['p4.trunk.out'][:, :A].clone()
                self.actor.outputs['std'] = t['log_std'].exp()
                self.actor.log(L, step)
                self.critic.outputs['q1'] = t['p5.q1.out'].clone()      # critic(obs, pi): the last critic forward
                self.critic.outputs['q2'] = t['p5.q2.out'].clone()
        self._fill_encoder_outputs(only_cpc, did_actor)

    def _fill_encoder_outputs(self, only_cpc, did_actor):
        """encoder.outputs as the reference's modules hold them after update() (encoder.py:79-108):
        obs / conv1..4 / fc / ln of each encoder's LAST forward.  Decoded lazily from the engine's
        activation buffers; only under --log_param_hist_imgs."""
        eng = self.engine
        t = eng.t
        fd = eng.cfg.feature_dim
        do_cpc = not self.pixel_sac
        enc = self.critic.encoder
        if do_cpc or did_actor or not only_cpc:
            # the critic encoder's last pass ran on obs (F3, or F4 = F5 = F6) into actA
            for l in range(4):
                enc.outputs['conv%d' % (l + 1)] = eng.act_nchw('actA', l)
            tail = 'p5' if (do_cpc or did_actor) else 'p3'
            enc.outputs['fc'] = t[tail + '.fc_out'][:, :fd].clone()
            enc.outputs['ln'] = t[tail + '.z'][:, :fd].clone()
        if do_cpc:
            tgt = self.critic_target.encoder
            for l in range(4):
                tgt.outputs['conv%d' % (l + 1)] = eng.act_nchw('actB', l)
            tgt.outputs['fc'] = t['p7.fc_out'][:, :fd].clone()
            tgt.outputs['ln'] = t['p7.z'][:, :fd].clone()
        if did_actor:
            aenc = self.actor.encoder
            for l in range(4):
                aenc.outputs['conv%d' % (l + 1)] = enc.outputs['conv%d' % (l + 1)]      # tied convs, same obs
            aenc.outputs['fc'] = t['p4.fc_out'][:, :fd].clone()
            aenc.outputs['ln'] = t['p4.z'][:, :fd].clone()

    def save(self, model_dir, augmentation, step):
        cpu = lambda sd: {k: v.detach().cpu() for k, v in sd.items()}
        torch.save(cpu(self.CURL.state_dict()), '%s/%s_curl_%s.pt' % (model_dir, augmentation, step))
        torch.save(cpu(self.actor.state_dict()), '%s/%s_actor_%s.pt' % (model_dir, augmentation, step))
        torch.save(cpu(self.critic.state_dict()), '%s/%s_critic_%s.pt' % (model_dir, augmentation, step))

    def load(self, model_dir, augmentation, step):
        ld = lambda p: torch.load(p, map_location='cpu')
        self.CURL.load_state_dict(ld('%s/%s_curl_%s.pt' % (model_dir, augmentation, step)))
        print('Loaded model %s/%s_curl_%s.pt' % (model_dir, augmentation, step))
        self.actor.load_state_dict(ld('%s/%s_actor_%s.pt' % (model_dir, augmentation, step)))
        print('Loaded model %s/%s_actor_%s.pt' % (model_dir, augmentation, step))
        self.critic.load_state_dict(ld('%s/%s_critic_%s.pt' % (model_dir, augmentation, step)))
        self.critic_target.load_state_dict(self.critic.state_dict())
        print('Loaded model %s/%s_critic_%s.pt' % (model_dir, augmentation, step))
