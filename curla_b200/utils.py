"""Drop-in for the reference's utils.py (utils.py:21-268).

ReplayBuffer keeps the reference's constructor, attributes and methods, but the storage
lives in GPU HBM (uint8 frames stay uint8) and sample_cpc() is the fused gather+crop
kernel instead of a host fancy-index + PCIe upload.  The numpy global RNG is consumed in
the reference's order (utils.py:147, augmentations.py:66-67,192-194).
"""
import ctypes as C
import os
import random
from collections import deque

import numpy as np
import torch

from . import _lib
from . import augmentations


class eval_mode(object):
    """utils.py:21-34"""

    def __init__(self, *models):
        self.models = models

    def __enter__(self):
        self.prev_states = []
        for model in self.models:
            self.prev_states.append(model.training)
            model.train(False)

    def __exit__(self, *args):
        for model, state in zip(self.models, self.prev_states):
            model.train(state)
        return False


def soft_update_params(net, target_net, tau):
    """utils.py:37-41 for any two objects exposing parameters(); the agent's own target
    update goes through the fused EMA kernel over the flat arenas instead."""
    pairs = list(zip(net.parameters(), target_net.parameters()))
    for param, target_param in pairs:
        n = param.numel()
        if param.is_cuda and param.dtype == torch.float32 and param.is_contiguous() \
                and target_param.is_contiguous():
            with torch.cuda.device(param.device):
                _lib.call('curla_ema_f32', _lib.ptr(target_param.data), _lib.ptr(param.data), n, n,
                          float(tau), float(tau),
                          C.c_void_p(torch.cuda.current_stream(param.device).cuda_stream))
        else:
            raise _lib.CurlaError('soft_update_params: contiguous fp32 CUDA parameters required')
    for net_obj in (target_net,):
        if hasattr(net_obj, '_after_param_write'):
            net_obj._after_param_write()


def set_seed_everywhere(seed):
    """utils.py:44-49"""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)


def make_dir(dir_path):
    """utils.py:59-65"""
    try:
        os.mkdir(dir_path)
    except OSError:
        print('Unable to create directory ' + dir_path)
    return dir_path


class ReplayBuffer(object):
    """Buffer to store environment transitions (utils.py:80-216), resident in HBM."""

    def __init__(self, obs_shape, action_shape, capacity, batch_size, device, augmentor, transform=None):
        self.capacity = capacity
        self.batch_size = batch_size
        self.device = torch.device(device)
        self.augmentor = augmentor
        self.transform = transform
        if self.device.type != 'cuda':
            raise _lib.CurlaError('ReplayBuffer: the B200 path keeps the buffer in GPU memory; '
                                  'device must be a CUDA device (no CPU fallback)')
        if len(obs_shape) != 3:
            raise NotImplementedError('only pixel observations (C, H, W) are on the accelerated path')
        self.obs_shape = tuple(obs_shape)
        self.action_shape = tuple(action_shape)

        frame_bytes = int(np.prod(obs_shape))
        total_bytes = 2 * capacity * frame_bytes + capacity * 4 * (int(np.prod(action_shape)) + 2)
        free, total = torch.cuda.mem_get_info(self.device)
        print('-' * 50)
        unit, name = (1024 ** 3, 'GB') if total_bytes > 1024 ** 3 else (1024 ** 2, 'MB')
        print('Replay buffer size: %.2f %s' % (total_bytes / unit, name))
        print('Total memory: %.2f %s' % (total / unit, name))
        print('Total available memory: %.2f %s' % (free / unit, name))
        print('-' * 50)
        if total_bytes > free:
            raise ValueError('Replay buffer size exceeds available memory')   # utils.py:112-113

        self.obses = torch.empty((capacity, *obs_shape), dtype=torch.uint8, device=self.device)
        self.next_obses = torch.empty((capacity, *obs_shape), dtype=torch.uint8, device=self.device)
        self.actions = torch.empty((capacity, *action_shape), dtype=torch.float32, device=self.device)
        self.rewards = torch.empty((capacity, 1), dtype=torch.float32, device=self.device)
        self.not_dones = torch.empty((capacity, 1), dtype=torch.float32, device=self.device)

        # pinned staging for add() (one transition) and for the sampled index block
        self._stage_obs = torch.empty((2, *obs_shape), dtype=torch.uint8).pin_memory()
        self._stage_vec = torch.empty((int(np.prod(action_shape)) + 2,), dtype=torch.float32).pin_memory()
        self._stage_idx = torch.empty((2, 7, batch_size), dtype=torch.int64).pin_memory()
        self._dev_idx = torch.empty((2, 7, batch_size), dtype=torch.int64, device=self.device)
        self._idx_slot = 0
        self._idx_events = [None, None]
        self._add_event = None
        # numpy views of the pinned staging buffers (host-side writes without tensor indexing overhead)
        self._np_obs = self._stage_obs.numpy()
        self._np_vec = self._stage_vec.numpy()
        self._np_idx = self._stage_idx.numpy()
        self._dev_vec = torch.zeros_like(self._stage_vec, device=self.device)
        self._add_ptrs = None

        self.idx = 0
        self.last_save = 0
        self.full = False

    # -- ingest (utils.py:120-128) -----------------------------------------------
    def add(self, obs, action, reward, next_obs, done):
        """One transition from host numpy into the HBM ring: two pinned-staged async H2D copies for
        the frame stacks and one 16-byte copy + one scatter kernel for action / reward / not_done."""
        if self._add_event is not None:
            self._add_event.synchronize()        # staging buffers are free again
        na = self.actions.shape[1]
        np.copyto(self._np_obs[0], obs, casting='unsafe')
        np.copyto(self._np_obs[1], next_obs, casting='unsafe')
        v = self._np_vec
        v[:na] = np.asarray(action, dtype=np.float32).reshape(-1)
        v[na] = reward
        v[na + 1] = float(not done)
        i = self.idx
        stream = torch.cuda.current_stream(self.device)
        if self._add_ptrs is None:
            fb = int(np.prod(self.obs_shape))
            self._add_ptrs = (self._stage_obs[0].data_ptr(), self._stage_obs[1].data_ptr(), fb, self._stage_vec.data_ptr(),
                              self._dev_vec.data_ptr(), self.obses.data_ptr(), self.next_obses.data_ptr(),
                              self.actions.data_ptr(), self.rewards.data_ptr(), self.not_dones.data_ptr())
        p = self._add_ptrs
        with torch.cuda.device(self.device):
            _lib.call('curla_replay_add', p[0], p[1], p[2], p[3], p[4], na, i, p[5], p[6], p[7], p[8], p[9],
                      C.c_void_p(stream.cuda_stream))
        if self._add_event is None:
            self._add_event = torch.cuda.Event()
        self._add_event.record(stream)

        self.idx = (self.idx + 1) % self.capacity
        self.full = self.full or self.idx == 0

    # -- sampling ----------------------------------------------------------------
    def draw_indices(self):
        """Consume the numpy global RNG exactly like sample_cpc (utils.py:147,156-158) and
        upload the draws as one int64 block.  Returns (dict of host arrays, device [7,B])
        with rows idxs, h1_obs, w1_obs, h1_next, w1_next, h1_pos, w1_pos."""
        B = self.batch_size
        ev = self._idx_events[self._idx_slot]
        if ev is not None:
            ev.synchronize()                                  # pinned slot no longer in flight
        host = self._np_idx[self._idx_slot]
        d = {}
        d['idxs'] = np.random.randint(0, self.capacity if self.full else self.idx, size=B)
        host[0] = d['idxs']
        if isinstance(self.augmentor, augmentations.RandomCrop):
            for r, name in enumerate(('obs', 'next', 'pos')):
                h1, w1 = self.augmentor.draw_offsets(B, self.obs_shape[1:])
                d['h1_' + name], d['w1_' + name] = h1, w1
                host[1 + 2 * r] = h1
                host[2 + 2 * r] = w1
        dev = self._dev_idx[self._idx_slot]
        dev.copy_(self._stage_idx[self._idx_slot], non_blocking=True)
        if ev is None:
            ev = self._idx_events[self._idx_slot] = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._idx_slot ^= 1
        return d, dev

    def _gather_f32(self, frames, idxs, h1, w1, out_hw):
        B = self.batch_size
        c, hf, wf = self.obs_shape
        out = torch.empty((B, c, out_hw[0], out_hw[1]), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call('curla_gather_crop_f32', _lib.ptr(frames), c, hf, wf, _lib.ptr(idxs), _lib.ptr(h1),
                      _lib.ptr(w1), B, out_hw[0], out_hw[1], _lib.ptr(out),
                      C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def _gather_rows(self, src, idxs):
        B = self.batch_size
        k = src.shape[1]
        out = torch.empty((B, k), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call('curla_gather_rows_f32', _lib.ptr(src), _lib.ptr(idxs), B, k, _lib.ptr(out),
                      C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def sample_cpc(self):
        """utils.py:144-187: (obs, action, reward, next_obs, not_done, cpc_kwargs) as float32
        tensors on `device`; obs_anchor IS obs (same tensor object)."""
        d, dev = self.draw_indices()
        idxs = dev[0]
        if isinstance(self.augmentor, augmentations.RandomCrop):
            ohw = self.augmentor.output_shape
            obses = self._gather_f32(self.obses, idxs, dev[1], dev[2], ohw)
            next_obses = self._gather_f32(self.next_obses, idxs, dev[3], dev[4], ohw)
            pos = self._gather_f32(self.obses, idxs, dev[5], dev[6], ohw)
        else:
            hw = self.obs_shape[1:]
            obses = self._gather_f32(self.obses, idxs, None, None, hw)
            next_obses = self._gather_f32(self.next_obses, idxs, None, None, hw)
            pos = obses.detach().clone()
            obses = self.augmentor.training_augmentation(obses)
            next_obses = self.augmentor.training_augmentation(next_obses)
            pos = self.augmentor.training_augmentation(pos)
        actions = self._gather_rows(self.actions, idxs)
        rewards = self._gather_rows(self.rewards, idxs)
        not_dones = self._gather_rows(self.not_dones, idxs)
        cpc_kwargs = dict(obs_anchor=obses, obs_pos=pos, time_anchor=None, time_pos=None)
        return obses, actions, rewards, next_obses, not_dones, cpc_kwargs

    # -- persistence (utils.py:189-216): same chunk files, numpy payload -------------
    def save(self, save_dir):
        if self.idx == self.last_save:
            return
        path = os.path.join(save_dir, '%d_%d.pt' % (self.last_save, self.idx))
        sl = slice(self.last_save, self.idx)
        payload = [self.obses[sl].cpu().numpy(), self.next_obses[sl].cpu().numpy(),
                   self.actions[sl].cpu().numpy(), self.rewards[sl].cpu().numpy(),
                   self.not_dones[sl].cpu().numpy()]
        self.last_save = self.idx
        torch.save(payload, path)

    def load(self, save_dir):
        chunks = os.listdir(save_dir)
        chucks = sorted(chunks, key=lambda x: int(x.split('_')[0]))
        for chunk in chucks:
            start, end = [int(x) for x in chunk.split('.')[0].split('_')]
            path = os.path.join(save_dir, chunk)
            payload = torch.load(path, weights_only=False)
            assert self.idx == start
            for dst, src in zip((self.obses, self.next_obses, self.actions, self.rewards, self.not_dones),
                                payload):
                dst[start:end].copy_(torch.as_tensor(src))
            self.idx = end

    def sample_proprio(self):
        """utils.py:130-142: un-augmented (obs, action, reward, next_obs, not_done); dead code in the
        reference (no caller), kept for API completeness."""
        B = self.batch_size
        idxs = np.random.randint(0, self.capacity if self.full else self.idx, size=B)
        dev = torch.from_numpy(idxs).to(self.device)
        hw = self.obs_shape[1:]
        obses = self._gather_f32(self.obses, dev, None, None, hw)
        next_obses = self._gather_f32(self.next_obses, dev, None, None, hw)
        return obses, self._gather_rows(self.actions, dev), self._gather_rows(self.rewards, dev), next_obses, \
            self._gather_rows(self.not_dones, dev)

    def __getitem__(self, idx):
        """utils.py:218-233: ONE random transition as host numpy (the argument is ignored, as in the
        reference); dead code there, kept for API completeness."""
        i = int(np.random.randint(0, self.capacity if self.full else self.idx, size=1)[0])
        obs, next_obs = self.obses[i].cpu().numpy(), self.next_obses[i].cpu().numpy()
        if self.transform:
            obs, next_obs = self.transform(obs), self.transform(next_obs)
        return obs, self.actions[i].cpu().numpy(), self.rewards[i].cpu().numpy(), next_obs, self.not_dones[i].cpu().numpy()

    def __len__(self):
        return self.capacity


class FrameStack(object):
    """utils.py:238-268 without the gymnasium dependency (duck-typed wrapper: same
    attributes and reset/step/_get_obs behaviour)."""

    class _Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    def __init__(self, env, k):
        self.env = env
        self._k = k
        self._frames = deque([], maxlen=k)
        shp = env.observation_space.shape
        self.observation_space = FrameStack._Box(low=0, high=1, shape=((shp[0] * k,) + shp[1:]),
                                                 dtype=env.observation_space.dtype)
        self._max_episode_steps = env._max_episode_steps
        self.curl_driving = False

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self):
        obs = self.env.reset()
        self.curl_driving = self.env.curl_driving
        for _ in range(self._k):
            self._frames.append(obs)
        return self._get_obs()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self.env.curl_driving = self.curl_driving
        self._frames.append(obs)
        return self._get_obs(), reward, done, info

    def _get_obs(self):
        assert len(self._frames) == self._k
        return np.concatenate(list(self._frames), axis=0)
