"""Drop-in API tests (GPU): the calls train.py / eval.py / plot_tsne/latent_data.py make on the
reference modules, made on curla_b200's mirrors and checked against the CPU oracle.

  ReplayBuffer.add / sample_cpc / save / load          utils.py:120-216
  CurlSacAgent.sample_action / select_action            curl_sac.py:330-347
  agent.actor.encoder(obs), agent.critic(obs, action)    plot_tsne/latent_data.py:83-93
  CurlSacAgent.save / load (reference file names, keys)  curl_sac.py:453-465
  the learner portion of train.py's main loop            train.py:346-443
"""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import curla_oracle as O
from oracle import scenario as S

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0) if torch.cuda.is_available() else None

import test_update_parity_gpu as T   # build_cuda_agent, NullLogger, rel_l2


def _agent(name='crop90x160'):
    cfg = dict(S.SCENARIOS[name])
    run = S.OracleRun(cfg)
    agent, rb = T.build_cuda_agent(cfg, run)
    return cfg, run, agent, rb


# ------------------------------------------------------------------ replay buffer
def test_replay_add_sample_save_load():
    from curla_b200 import augmentations, utils
    hw, cap, B = (90, 160), 12, 5
    aug = augmentations.make_augmentor('random_crop', hw)
    rb = utils.ReplayBuffer((9, *hw), (2,), cap, B, DEV, aug)
    rs = np.random.RandomState(11)
    host = []
    for i in range(cap + 3):                         # wraps around: idx 3, full
        tr = (rs.randint(0, 256, size=(9, *hw), dtype=np.uint8), rs.uniform(-1, 1, size=2).astype(np.float32),
              float(rs.standard_normal()), rs.randint(0, 256, size=(9, *hw), dtype=np.uint8), bool(i % 4 == 0))
        host.append(tr)
        rb.add(*tr)
        assert rb.idx == (i + 1) % cap and rb.full == (i + 1 >= cap)
    torch.cuda.synchronize()
    # ring contents: slot j holds the last transition written to it
    last = {i % cap: tr for i, tr in enumerate(host)}
    for j, tr in last.items():
        assert np.array_equal(rb.obses[j].cpu().numpy(), tr[0])
        assert np.array_equal(rb.next_obses[j].cpu().numpy(), tr[3])
        assert np.allclose(rb.actions[j].cpu().numpy(), tr[1])
        assert float(rb.rewards[j]) == pytest.approx(tr[2])
        assert float(rb.not_dones[j]) == float(not tr[4])              # stores NOT done (utils.py:125)

    # sample_cpc: same numpy draws, bit-exact frames, reference return contract
    np.random.seed(5)
    st = np.random.get_state()
    obs, action, reward, next_obs, not_done, kw = rb.sample_cpc()
    np.random.set_state(st)
    d = O.draw_sample_indices(cap, rb.idx, rb.full, B, 'random_crop', hw, tuple(aug.output_shape))
    assert kw['obs_anchor'] is obs and sorted(kw) == ['obs_anchor', 'obs_pos', 'time_anchor', 'time_pos']
    assert kw['time_anchor'] is None and kw['time_pos'] is None
    for t in (obs, action, reward, next_obs, not_done, kw['obs_pos']):
        assert t.dtype == torch.float32 and t.device == DEV
    assert tuple(obs.shape) == (B, 9, 76, 135) and tuple(action.shape) == (B, 2) and tuple(reward.shape) == (B, 1)
    ob = rb.obses.cpu().numpy(); nb = rb.next_obses.cpu().numpy()
    assert np.array_equal(obs.cpu().numpy(), O.gather_crop(ob, d['idxs'], d['h1_obs'], d['w1_obs'], (76, 135)).astype(np.float32))
    assert np.array_equal(next_obs.cpu().numpy(), O.gather_crop(nb, d['idxs'], d['h1_next'], d['w1_next'], (76, 135)).astype(np.float32))
    assert np.array_equal(kw['obs_pos'].cpu().numpy(), O.gather_crop(ob, d['idxs'], d['h1_pos'], d['w1_pos'], (76, 135)).astype(np.float32))
    assert np.array_equal(action.cpu().numpy(), rb.actions.cpu().numpy()[d['idxs']])

    # save / load: the reference's chunk files "<start>_<end>.pt" holding five numpy slices
    rb2 = utils.ReplayBuffer((9, *hw), (2,), cap, B, DEV, aug)
    for tr in host[:7]:
        rb2.add(*tr)
    with tempfile.TemporaryDirectory() as tmp:
        rb2.save(tmp)
        assert os.listdir(tmp) == ['0_7.pt']
        payload = torch.load(os.path.join(tmp, '0_7.pt'), weights_only=False)
        assert len(payload) == 5 and isinstance(payload[0], np.ndarray) and payload[0].shape == (7, 9, *hw)
        rb2.save(tmp)                                  # nothing new: no second file (utils.py:190-191)
        assert os.listdir(tmp) == ['0_7.pt']
        rb3 = utils.ReplayBuffer((9, *hw), (2,), cap, B, DEV, aug)
        rb3.load(tmp)
        assert rb3.idx == 7
        assert torch.equal(rb3.obses[:7], rb2.obses[:7]) and torch.equal(rb3.rewards[:7], rb2.rewards[:7])


def test_replay_rejects_cpu_device():
    from curla_b200 import _lib, augmentations, utils
    aug = augmentations.make_augmentor('identity', (90, 160))
    with pytest.raises(_lib.CurlaError):
        utils.ReplayBuffer((9, 90, 160), (2,), 4, 2, torch.device('cpu'), aug)


# ------------------------------------------------------------------ inference entry points
@pytest.mark.parametrize('name', ['crop90x160', 'identity90x160'])
def test_encoder_critic_actor_calls_match_oracle(name):
    """latent_data.py:83-93: agent.actor.encoder(obs), agent.critic(obs, action), and the B=1
    action path; 70 observations exercise the chunking over the engine's batch."""
    cfg, run, agent, rb = _agent(name)
    o = run.agent
    rs = np.random.RandomState(3)
    n = 70
    obs = rs.randint(0, 256, size=(n, *run.obs_shape)).astype(np.float32)
    act = rs.uniform(-1, 1, size=(n, 2)).astype(np.float32)
    to = torch.from_numpy(obs)
    with torch.no_grad():
        z_ref = O.encoder_forward(o.actor, 'encoder.', to)
        zc_ref = O.encoder_forward(o.critic, 'encoder.', to)
        q1_ref, q2_ref = O.critic_forward(o.critic, to, torch.from_numpy(act))
        mu_ref = O.actor_forward(o.actor, to, None, o.log_std_min, o.log_std_max, compute_pi=False, compute_log_pi=False)[0]
    z = agent.actor.encoder(to.to(DEV))
    assert tuple(z.shape) == (n, S.FEATURE_DIM)
    assert T.rel_l2(z, z_ref) < 2e-2
    assert T.rel_l2(agent.critic.encoder(to.to(DEV)), zc_ref) < 2e-2
    q1, q2 = agent.critic(to.to(DEV), torch.from_numpy(act).to(DEV))
    assert tuple(q1.shape) == (n, 1)
    assert T.rel_l2(q1, q1_ref) < 3e-2 and T.rel_l2(q2, q2_ref) < 3e-2
    # select_action: tanh(mu) of one observation, 1-D numpy of length 2 (curl_sac.py:330-336)
    a = agent.select_action(obs[0])
    assert isinstance(a, np.ndarray) and a.shape == (2,)
    assert np.allclose(a, mu_ref[0].numpy(), atol=3e-2)
    # sample_action center-crops stored-size frames itself (curl_sac.py:338-347)
    frame = rs.randint(0, 256, size=(9, *cfg['frame_hw'])).astype(np.uint8)
    s = agent.sample_action(frame)
    assert s.shape == (2,) and np.all(np.abs(s) <= 1.0) and np.all(np.isfinite(s))


def test_batched_latent_extraction_matches_per_observation_oracle():
    """curla_b200.latent.extract == the loop of plot_tsne/latent_data.py:63-104 (center crop,
    actor.encoder latent, sampled action, min(Q1, Q2)) run by the oracle one observation at a time."""
    from curla_b200 import latent
    cfg, run, agent, rb = _agent('crop90x160')
    o = run.agent
    rs = np.random.RandomState(9)
    n = 70
    frames = rs.randint(0, 256, size=(n, 9, 90, 160), dtype=np.uint8)
    noise = rs.standard_normal(size=(n, 2)).astype(np.float32)
    got = latent.extract(agent, frames, sample=True, noise=noise)
    top, left = O.center_crop_offsets((90, 160), (76, 135))
    assert (top, left) == latent.center_window((90, 160), (76, 135)) == (7, 12)
    reps, acts, qs = [], [], []
    with torch.no_grad():
        for i in range(n):
            ob = torch.from_numpy(frames[i:i + 1, :, top:top + 76, left:left + 135].astype(np.float32))
            reps.append(O.encoder_forward(o.actor, 'encoder.', ob))
            pi = O.actor_forward(o.actor, ob, torch.from_numpy(noise[i:i + 1]), o.log_std_min, o.log_std_max,
                                 compute_log_pi=False)[1]
            acts.append(pi)
            q1, q2 = O.critic_forward(o.critic, ob, pi)
            qs.append(torch.minimum(q1, q2).reshape(1))
    assert got['representations'].shape == (n, 50) and got['q_values'].shape == (n,)
    assert T.rel_l2(torch.from_numpy(got['representations']), torch.cat(reps)) < 2e-2
    assert np.abs(got['actions'] - torch.cat(acts).numpy()).max() < 5e-2
    assert T.rel_l2(torch.from_numpy(got['q_values']), torch.cat(qs)) < 5e-2
    # an inference-only agent grows its engine to run this in one chunk
    assert agent.engine.cfg.batch == n


def test_graph_action_path_equals_eager_path(monkeypatch):
    """The captured-graph B=1 path (uint8 frame in, center crop on the device) returns what the
    eager module-call path returns for the same observation."""
    cfg, run, agent, rb = _agent('crop90x160')
    rs = np.random.RandomState(4)
    for _ in range(3):
        frame = rs.randint(0, 256, size=(9, 90, 160)).astype(np.uint8)
        a_graph = agent.select_action(frame[:, 7:83, 12:147])                 # uint8, encoder-sized
        a_f32 = agent.select_action(frame[:, 7:83, 12:147].astype(np.float32))
        monkeypatch.setenv('CURLA_NO_GRAPH', '1')
        a_eager = agent.select_action(frame[:, 7:83, 12:147].astype(np.float32))
        monkeypatch.delenv('CURLA_NO_GRAPH')
        assert np.array_equal(a_graph, a_f32) and np.array_equal(a_graph, a_eager)
    # the cropping route: sample_action on a stored-size frame with zero noise == select_action on its center crop
    frame = rs.randint(0, 256, size=(9, 90, 160)).astype(np.uint8)
    import torch as _t
    real_randn = _t.randn
    monkeypatch.setattr(_t, 'randn', lambda *a, out=None, **k: out.zero_() if out is not None else real_randn(*a, **k))
    s0 = agent.sample_action(frame)
    monkeypatch.setattr(_t, 'randn', real_randn)
    assert np.allclose(s0, agent.select_action(frame[:, 7:83, 12:147]), atol=1e-6)
    # graphs survive an engine re-creation (first update changes the engine batch)
    agent.update(rb, T.NullLogger(), 0)
    a1 = agent.select_action(frame[:, 7:83, 12:147])
    monkeypatch.setenv('CURLA_NO_GRAPH', '1')
    a2 = agent.select_action(frame[:, 7:83, 12:147].astype(np.float32))
    assert np.array_equal(a1, a2)


def test_sample_action_is_stochastic_select_is_not():
    cfg, run, agent, rb = _agent('crop90x160')
    frame = np.random.RandomState(0).randint(0, 256, size=(9, 90, 160)).astype(np.uint8)
    assert np.array_equal(agent.select_action(frame[:, 7:83, 12:147]), agent.select_action(frame[:, 7:83, 12:147]))
    draws = np.stack([agent.sample_action(frame) for _ in range(6)])
    assert np.unique(draws.round(6), axis=0).shape[0] > 1


# ------------------------------------------------------------------ checkpoints
def test_save_load_reference_files_and_keys():
    cfg, run, agent, rb = _agent('crop90x160')
    with tempfile.TemporaryDirectory() as tmp:
        agent.save(tmp, 'random_crop', 1234)
        assert sorted(os.listdir(tmp)) == ['random_crop_actor_1234.pt', 'random_crop_critic_1234.pt', 'random_crop_curl_1234.pt']
        curl = torch.load(os.path.join(tmp, 'random_crop_curl_1234.pt'))
        critic = torch.load(os.path.join(tmp, 'random_crop_critic_1234.pt'))
        actor = torch.load(os.path.join(tmp, 'random_crop_actor_1234.pt'))
        enc = ['encoder.convs.%d.%s' % (i, p) for i in range(4) for p in ('weight', 'bias')] + \
              ['encoder.fc.weight', 'encoder.fc.bias', 'encoder.ln.weight', 'encoder.ln.bias']
        mlp = lambda pre: [pre + '%d.%s' % (i, p) for i in (0, 2, 4) for p in ('weight', 'bias')]
        assert sorted(critic) == sorted(enc + mlp('Q1.trunk.') + mlp('Q2.trunk.'))
        assert sorted(actor) == sorted(enc + mlp('trunk.'))
        assert sorted(curl) == sorted(['W'] + enc + [k.replace('encoder.', 'encoder_target.', 1) for k in enc])
        # PyTorch layouts: conv OIHW, linear [out][in] with the reference's NCHW-flatten fc order
        assert tuple(critic['encoder.convs.0.weight'].shape) == (32, 9, 3, 3)
        assert tuple(critic['encoder.fc.weight'].shape) == (50, 32 * 31 * 61)
        actor_sd, critic_sd, W = run.state_dicts
        for k, v in critic_sd.items():
            assert torch.equal(critic[k], v), k                         # exact round trip of fp32 masters
        assert torch.equal(curl['W'], W)

        # load into a fresh agent: parameters equal, target re-copied from the critic (curl_sac.py:458-465)
        from curla_b200 import curl_sac
        fresh = curl_sac.CurlSacAgent(run.obs_shape, (2,), DEV, agent.augmentor, hidden_dim=cfg['hidden'], **S.HP)
        fresh.load(tmp, 'random_crop', 1234)
        for k, v in agent.critic.state_dict().items():
            assert torch.equal(fresh.critic.state_dict()[k], v), k
            assert torch.equal(fresh.critic_target.state_dict()[k], v), k
        for k, v in agent.actor.state_dict().items():
            assert torch.equal(fresh.actor.state_dict()[k], v), k
        obs = torch.from_numpy(np.random.RandomState(1).randint(0, 256, size=(3, *run.obs_shape)).astype(np.float32)).to(DEV)
        assert torch.equal(fresh.actor.encoder(obs), agent.actor.encoder(obs))


def test_unsupported_shapes_raise_like_the_reference():
    from curla_b200 import augmentations, curl_sac, encoder
    with pytest.raises(NotImplementedError):
        encoder.out_dim_for((9, 50, 50), 4)                               # encoder.py:57-66
    with pytest.raises(ValueError):
        augmentations.make_augmentor('nope', (90, 160))                    # augmentations.py:220-221


# ------------------------------------------------------------------ train.py's learner loop
class _StubEnv:
    """Stands in for utils.FrameStack(CarlaEnv): (9, 90, 160) uint8 observations, Box(2,) actions."""

    def __init__(self, seed=0):
        self.rs = np.random.RandomState(seed)
        self.t = 0

    def _obs(self):
        return self.rs.randint(0, 256, size=(9, 90, 160), dtype=np.uint8)

    def reset(self):
        self.t = 0
        return self._obs()

    def step(self, action):
        self.t += 1
        return self._obs(), float(-np.square(action).sum()), self.t >= 12, {}


def test_train_py_learner_loop_with_stub_env():
    """The body of train.py:346-443 (minus CARLA, logger files and eval): random actions for
    init_steps, then agent.update once per env step, only_cpc during the first steps of an
    episode, sample_action inside utils.eval_mode."""
    from curla_b200 import augmentations, curl_sac, utils
    utils.set_seed_everywhere(3)
    aug = augmentations.make_augmentor('random_crop', (90, 160))
    rb = utils.ReplayBuffer((9, 90, 160), (2,), 64, 8, DEV, aug)
    agent = curl_sac.CurlSacAgent((9, *aug.output_shape), (2,), DEV, aug, hidden_dim=64, log_interval=1, **S.HP)
    env, L = _StubEnv(), T.NullLogger()
    init_steps, acc_steps = 10, 3
    obs, episode_step, done = env.reset(), 0, False
    for step in range(40):
        if done:
            obs, episode_step, done = env.reset(), 0, False
        if step < init_steps:
            action = np.random.uniform(-1, 1, size=2).astype(np.float32)
        elif episode_step < acc_steps:
            action = np.array([0.5, 0.0], dtype=np.float32)
        else:
            with utils.eval_mode(agent):
                action = agent.sample_action(obs)
        if step >= init_steps:
            agent.update(rb, L, step, only_cpc=episode_step < acc_steps)
        next_obs, reward, done, _ = env.step(action)
        rb.add(obs, action, reward, next_obs, float(done))
        obs = next_obs
        episode_step += 1
    torch.cuda.synchronize()
    assert agent.training is True
    keys = {k for (_, k) in L.rows}
    assert {'train/batch_reward', 'train_critic/loss', 'train_actor/loss', 'train_actor/entropy', 'train_alpha/loss',
            'train_alpha/value', 'train/curl_loss'} <= keys
    assert all(np.isfinite(v) for v in L.rows.values())
    only = [s for (s, k) in L.rows if k == 'train/curl_loss']
    crit = [s for (s, k) in L.rows if k == 'train_critic/loss']
    assert len(only) == 30 and 0 < len(crit) < 30             # CPC every update, SAC only outside the warm-up steps


# ------------------------------------------------------------------ logging taps (--log_param_hist_imgs)
class _RecordingLogger(T.NullLogger):
    """The part of logger.Logger the update touches under --log_param_hist_imgs: log_histogram and
    log_param exactly as logger.py:155-173 reads its arguments (.weight.data, .weight.grad.data, ...)."""

    def __init__(self):
        super().__init__()
        self.hist = {}

    def log_histogram(self, key, histogram, step):
        assert key.startswith('train') or key.startswith('eval')
        self.hist[(step, key)] = histogram.detach().float().cpu().clone()

    def log_param(self, key, param, step):
        self.log_histogram(key + '_w', param.weight.data, step)
        if hasattr(param.weight, 'grad') and param.weight.grad is not None:
            self.log_histogram(key + '_w_g', param.weight.grad.data, step)
        if hasattr(param, 'bias'):
            self.log_histogram(key + '_b', param.bias.data, step)
            if hasattr(param.bias, 'grad') and param.bias.grad is not None:
                self.log_histogram(key + '_b_g', param.bias.grad.data, step)


def test_log_param_hist_imgs_taps():
    """curl_sac.py:370-371,394-395 + curl_sac.py:112-121,171-180: with log_param_hist_imgs the update logs
    the critic's q1/q2 outputs, the actor's mu/std and weight / bias / GRADIENT histograms of every trunk
    layer, on steps that are multiples of LOG_FREQ.  Values are checked against the oracle's teacher-free
    first update (identical weights before the first optimizer step)."""
    from curla_b200 import curl_sac
    cfg = dict(S.SCENARIOS['crop90x160'])
    run = S.OracleRun(cfg)
    agent, rb = T.build_cuda_agent(cfg, run)
    agent.log_param_hist_imgs = True
    L = _RecordingLogger()
    st = np.random.get_state()
    d, b, om = run.step()
    np.random.set_state(st)
    agent._noise_override = (run.noise[0, 0], run.noise[0, 1])
    agent.update(rb, L, 0)
    torch.cuda.synchronize()
    keys = {k for (s_, k) in L.hist if s_ == 0}
    expect = {'train_critic/q1_hist', 'train_critic/q2_hist', 'train_actor/mu_hist', 'train_actor/std_hist'}
    for i in range(3):
        for q in ('q1', 'q2'):
            expect |= {'train_critic/%s_fc%d_%s' % (q, i, sfx) for sfx in ('w', 'w_g', 'b', 'b_g')}
        expect |= {'train_actor/fc%d_%s' % (i + 1, sfx) for sfx in ('w', 'w_g', 'b', 'b_g')}
    assert keys == expect, (sorted(keys - expect), sorted(expect - keys))
    o = run.agent
    h = lambda k: L.hist[(0, k)]
    assert T.rel_l2(h('train_critic/q1_hist'), o.dbg['q1']) < 2e-2               # Q(obs, action) of update_critic
    g_ref = o.dbg['critic_grads']
    assert h('train_critic/q1_fc1_w_g').shape == g_ref['Q1.trunk.2.weight'].shape
    assert T.rel_l2(h('train_critic/q1_fc1_w_g'), g_ref['Q1.trunk.2.weight']) < 4e-1     # TOL_SMALL: B=4 scenario
    assert T.rel_l2(h('train_critic/q2_fc2_b_g'), g_ref['Q2.trunk.4.bias']) < 4e-1
    assert torch.equal(h('train_critic/q1_fc0_w'), agent.critic.Q1.trunk[0].weight.cpu())  # weights AFTER the step
    assert T.rel_l2(h('train_actor/fc2_w_g'), o.dbg['actor_grads']['trunk.2.weight']) < 4e-1
    assert h('train_actor/std_hist').shape == (cfg['B'], 2) and float(h('train_actor/std_hist').min()) > 0
    # outputs dicts as the reference's modules hold them after update(): critic(obs, pi) ran last
    assert torch.equal(agent.critic.outputs['q1'], agent.engine.t['p5.q1.out'])
    enc_out = agent.critic.encoder.outputs
    assert tuple(enc_out['conv4'].shape) == (cfg['B'], 32, 31, 61) and tuple(enc_out['ln'].shape) == (cfg['B'], 50)
    # (the CURL anchor latent is computed AFTER the critic's first, sign-like Adam step on a 4-sample batch: drift band,
    # as in __graft_entry__.smoke; measured 5.8e-2)
    assert T.rel_l2(enc_out['ln'], o.dbg['z_a']) < 1e-1
    # an odd step that is a multiple of nothing logs nothing; LOG_FREQ gates the histograms
    n = len(L.hist)
    agent.update(rb, L, 1)
    assert len(L.hist) == n and curl_sac.LOG_FREQ == 25_000


def test_engine_recreation_keeps_optimizer_state():
    """Re-creating the engine mid-run (the stored frame size changes: fused ReplayBuffer path -> float
    path) must carry the Adam step counters with the moments: the run continues bit-identically."""
    cfg = dict(S.SCENARIOS['crop90x160'])
    run = S.OracleRun(cfg)
    a1, rb1 = T.build_cuda_agent(cfg, run)
    a2, rb2 = T.build_cuda_agent(cfg, run)
    L = T.NullLogger()
    st = np.random.get_state()
    for u in range(4):
        for agent, rb in ((a1, rb1), (a2, rb2)):
            np.random.set_state(st)
            agent._noise_override = (run.noise[u % 2, 0], run.noise[u % 2, 1])
            if agent is a2 and u == 2:
                c = agent.engine.cfg
                agent._make_engine(c.batch, (c.Hf + 2, c.Wf + 2))        # any other geometry: new engine
                agent._make_engine(c.batch, (c.Hf - 2, c.Wf - 2))        # ... and back (2 hand-overs)
            agent.update(rb, L, u)
            after = np.random.get_state()
        st = after
    torch.cuda.synchronize()
    steps1, steps2 = (torch.zeros(4, dtype=torch.int32).numpy() for _ in range(2))
    import ctypes as C
    a1.engine.lib.curla_agent_get_opt_steps(a1.engine.h, steps1.ctypes.data_as(C.POINTER(C.c_int)))
    a2.engine.lib.curla_agent_get_opt_steps(a2.engine.h, steps2.ctypes.data_as(C.POINTER(C.c_int)))
    assert list(steps1) == list(steps2) == [4, 2, 2, 4]
    p1, p2 = a1.engine.arenas[0].view(torch.float32), a2.engine.arenas[0].view(torch.float32)
    assert torch.equal(p1, p2), float((p1 - p2).abs().max())


def test_soft_update_params_on_encoders_refreshes_kernel_weights():
    """utils.soft_update_params(agent.critic.encoder, agent.critic_target.encoder, tau) -- the reference
    idiom (curl_sac.py:443-445) -- must reach the bf16 kernel-layout copies the forward reads."""
    from curla_b200 import utils
    cfg, run, agent, rb = _agent('crop90x160')
    sd = agent.critic.state_dict()
    agent.critic.load_state_dict({k: (v * 1.25 if k.startswith('encoder.') else v) for k, v in sd.items()})
    obs = torch.from_numpy(np.random.RandomState(2).randint(0, 256, size=(3, *run.obs_shape)).astype(np.float32)).to(DEV)
    before = agent.critic_target.encoder(obs).clone()
    utils.soft_update_params(agent.critic.encoder, agent.critic_target.encoder, 0.5)
    after = agent.critic_target.encoder(obs)
    assert not torch.equal(before, after)
    tgt = {k: v.cpu() for k, v in agent.critic_target.state_dict().items()}
    for k, v in sd.items():
        if k.startswith('encoder.'):
            assert torch.allclose(tgt[k], (0.5 * 1.25 + 0.5) * v.cpu(), rtol=1e-6, atol=1e-7), k
    with torch.no_grad():
        z_ref = O.encoder_forward(tgt, 'encoder.', obs.cpu())
    assert T.rel_l2(after, z_ref) < 2e-2


def test_select_action_rejects_uncropped_observation():
    cfg, run, agent, rb = _agent('crop90x160')
    frame = np.zeros((9, 90, 160), dtype=np.uint8)
    with pytest.raises(ValueError):
        agent.select_action(frame)                       # the reference fails at the fc layer (curl_sac.py:330-336)
    assert agent.sample_action(frame).shape == (2,)      # sample_action center-crops (curl_sac.py:340-341)


def test_replay_sample_proprio_and_getitem():
    from curla_b200 import augmentations, utils
    hw, cap, B = (90, 160), 8, 3
    aug = augmentations.make_augmentor('identity', hw)
    rb = utils.ReplayBuffer((9, *hw), (2,), cap, B, DEV, aug)
    arrays = S.make_replay_arrays(cap, hw)
    for dst, src in zip((rb.obses, rb.next_obses, rb.actions, rb.rewards, rb.not_dones), arrays):
        dst.copy_(torch.from_numpy(src))
    rb.idx, rb.full = 0, True
    np.random.seed(3)
    obs, act, rew, nxt, nd = rb.sample_proprio()
    idxs = np.random.RandomState(3).randint(0, cap, size=B)
    assert np.array_equal(obs.cpu().numpy(), arrays[0][idxs].astype(np.float32))
    assert np.array_equal(nxt.cpu().numpy(), arrays[1][idxs].astype(np.float32))
    assert np.array_equal(act.cpu().numpy(), arrays[2][idxs]) and np.array_equal(nd.cpu().numpy(), arrays[4][idxs])
    np.random.seed(4)
    o1, a1, r1, n1, d1 = rb[123]
    i = np.random.RandomState(4).randint(0, cap, size=1)[0]
    assert np.array_equal(o1, arrays[0][i]) and np.array_equal(n1, arrays[1][i]) and np.array_equal(a1, arrays[2][i])


# ------------------------------------------------------------------ whole-update CUDA graph
@pytest.mark.parametrize('name', ['crop90x160', 'pixelsac90x160'])
def test_graph_replayed_updates_equal_eager_updates(name, monkeypatch):
    """curla_agent_update replays one captured CUDA graph per update variant (engine.cu); the Adam step
    counters and the Philox offset of the policy noise then live in device memory.  Ten updates (both step
    parities, one only_cpc stretch) replayed must leave parameters, Adam moments and optimizer step
    counters bit-identical to the same ten updates launched eagerly (CURLA_GRAPH=0)."""
    import ctypes as C
    cfg = dict(S.SCENARIOS[name])
    run = S.OracleRun(cfg)
    agents = []
    for flag in ('0', '1'):
        monkeypatch.setenv('CURLA_GRAPH', flag)
        agent, rb = T.build_cuda_agent(cfg, run)
        agent._noise_seed = 1234
        agents.append((agent, rb))
    L = T.NullLogger()
    st = np.random.get_state()
    launches = {}
    for u in range(10):
        only = (not cfg.get('pixel_sac', False)) and u in (6, 7)
        for flag, (agent, rb) in zip(('0', '1'), agents):
            monkeypatch.setenv('CURLA_GRAPH', flag)       # read once per engine, at its first update
            np.random.set_state(st)
            agent.update(rb, L, u, only_cpc=only)
            after = np.random.get_state()
            launches.setdefault(flag, []).append(agent.engine.last_launches())
        st = after
    torch.cuda.synchronize()
    (a0, _), (a1, _) = agents
    assert len(a1.engine.t) == len(a0.engine.t)
    # replays count the captured kernel nodes + the one-thread state launch
    assert launches['1'][-1] == launches['0'][-1] + 1, (launches['0'][-1], launches['1'][-1])
    s0, s1 = (np.zeros(4, dtype=np.int32) for _ in range(2))
    a0.engine.lib.curla_agent_get_opt_steps(a0.engine.h, s0.ctypes.data_as(C.POINTER(C.c_int)))
    a1.engine.lib.curla_agent_get_opt_steps(a1.engine.h, s1.ctypes.data_as(C.POINTER(C.c_int)))
    assert list(s0) == list(s1)
    for arena in (0, 3):                                  # fp32 parameters, Adam moments
        p0, p1 = a0.engine.arenas[arena].view(torch.float32), a1.engine.arenas[arena].view(torch.float32)
        assert torch.equal(p0, p1), (arena, float((p0 - p1).abs().max()))
    assert torch.equal(a0.engine.t['log_alpha'], a1.engine.t['log_alpha'])
    assert torch.equal(a0.engine.t['metrics'], a1.engine.t['metrics'])


def test_logged_scalars_through_the_pinned_mailbox_equal_the_synchronised_read(monkeypatch):
    """With a single GPU the update publishes its logged scalars into pinned host memory right after the last kernel
    that writes one (curla_publish_metrics) and update() returns once THEY have arrived, the CURL backward and the
    optimizer steps still in flight.  Every logged value of every step must equal the value read the slow way
    (metrics D2H after a full stream synchronisation, CURLA_MAILBOX=0), eager and graph-replayed alike."""
    cfg = dict(S.SCENARIOS['crop90x160'])
    run = S.OracleRun(cfg)
    rows = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('CURLA_MAILBOX', mode)
        agent, rb = T.build_cuda_agent(cfg, run)
        assert (getattr(agent, '_mailbox', None) is not None) == (mode == '1')
        agent._noise_seed = 77
        L = T.NullLogger()
        st = np.random.get_state()
        for u in range(9):
            agent.update(rb, L, u, only_cpc=(u == 5))
        np.random.set_state(st)
        torch.cuda.synchronize()
        rows[mode] = dict(L.rows)
        if mode == '1':      # the mailbox holds the last update's scalars, bit for bit what the device buffer holds
            m = agent.engine.t['metrics'].cpu().numpy()
            assert np.array_equal(agent._mailbox_f[:15], m[:15])
            assert int(agent._mailbox_u[15]) == 9
    assert rows['0'].keys() == rows['1'].keys() and len(rows['0']) > 40
    assert rows['0'] == rows['1']
