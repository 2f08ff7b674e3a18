"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # writes tests/golden/<scenario>.npz

The reference modules (curl_sac, encoder, utils, augmentations) are imported
through the three import shims under oracle/shims (gymnasium, skimage, kornia:
SURVEY.md section 8c).  Inputs come from oracle/scenario.py so the tests can rebuild
them; only the reference's OUTPUTS are stored (fingerprints: sum, |sum|, 64
strided samples per tensor; metrics; sampled indices; CRCs of sampled frames).
Policy noise is injected by replacing torch.randn_like while the reference runs.
"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('CURLA_REFERENCE', '/root/reference')


REF_FILES = ('curl_sac.py', 'encoder.py', 'utils.py', 'augmentations.py')     # the hot path's four modules
REF_SHIP = os.path.join(os.path.dirname(HERE), 'baseline', '_ref')            # git-ignored; travels to the GPU box


def reference_dir():
    """Where the UNMODIFIED reference modules can be imported from: /root/reference in the build
    container, else the verbatim copies `ship_reference()` left under baseline/_ref/ (the GPU box
    has no /root/reference).  None when neither exists."""
    for d in (REF, REF_SHIP):
        if all(os.path.isfile(os.path.join(d, f)) for f in REF_FILES):
            return d
    return None


def ship_reference():
    """Build-container step (called by __graft_entry__.build): copy the four reference modules,
    byte for byte, to baseline/_ref/ so that `bench.py --impl reference` can time the reference's
    OWN update on the GPU box's host cores.  The directory is git-ignored: reference sources never
    enter the repository history."""
    import shutil
    if not all(os.path.isfile(os.path.join(REF, f)) for f in REF_FILES):
        return None
    os.makedirs(REF_SHIP, exist_ok=True)
    for f in REF_FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(REF_SHIP, f))
    return REF_SHIP


def import_reference(ref_dir=None):
    ref_dir = ref_dir or REF
    sys.path.insert(0, os.path.join(HERE, 'shims'))
    sys.path.insert(0, ref_dir)
    import utils as ref_utils            # noqa: E402
    import augmentations as ref_aug      # noqa: E402
    import curl_sac as ref_sac           # noqa: E402
    import encoder as ref_enc            # noqa: E402
    assert os.path.realpath(ref_sac.__file__).startswith(os.path.realpath(ref_dir)), ref_sac.__file__
    return ref_utils, ref_aug, ref_sac, ref_enc


class RecordingLogger:
    def __init__(self):
        self.rows = {}

    def log(self, key, value, step):
        if isinstance(value, torch.Tensor):
            value = value.item()
        self.rows[(step, key)] = float(value)


def run_scenario(name, cfg, ref_utils, ref_aug, ref_sac):
    from . import scenario as S
    torch.set_num_threads(1)            # deterministic reduction order
    hw = tuple(cfg['frame_hw'])
    ohw = S.obs_hw(cfg)
    augmentor = ref_aug.make_augmentor(cfg['aug'], hw)
    assert tuple(augmentor.output_shape) == tuple(ohw)
    dev = torch.device('cpu')
    rb = ref_utils.ReplayBuffer((S.FRAME_C, *hw), (S.ACTION_DIM,), cfg['capacity'], cfg['B'],
                                dev, augmentor)
    arrays = S.make_replay_arrays(cfg['capacity'], hw)
    rb.obses[:], rb.next_obses[:], rb.actions[:], rb.rewards[:], rb.not_dones[:] = arrays
    rb.idx, rb.full = 0, True
    obs_shape = (S.FRAME_C, *ohw)
    agent = ref_sac.CurlSacAgent(obs_shape, (S.ACTION_DIM,), dev, augmentor,
                                 hidden_dim=cfg['hidden'], log_interval=1,
                                 detach_encoder=cfg['detach_encoder'],
                                 pixel_sac=cfg['pixel_sac'], **S.HP)
    actor_sd, critic_sd, W = S.make_state_dicts(obs_shape, cfg['hidden'])
    agent.critic.load_state_dict(critic_sd)
    agent.actor.load_state_dict(actor_sd)      # convs are tied: same values as critic's
    agent.critic_target.load_state_dict(agent.critic.state_dict())
    agent.CURL.W.data.copy_(W)
    assert agent.actor.encoder.convs[0].weight is agent.critic.encoder.convs[0].weight

    noise = S.make_noise(len(cfg['steps']), cfg['B'])
    out = {}
    L = RecordingLogger()

    # record grads right before every optimizer step
    grads = {}

    def wrap(opt, tag, named):
        orig = opt.step

        def step(*a, **k):
            for key, prm in named:
                if prm.grad is not None:
                    grads[(tag, key)] = S.summarize(prm.grad)
            return orig(*a, **k)
        opt.step = step

    crit_named = [('critic.' + k, v) for k, v in agent.critic.named_parameters()]
    act_named = [('actor.' + k, v) for k, v in agent.actor.named_parameters()
                 if not k.startswith('encoder.convs.')]
    enc_named = [('critic.encoder.' + k, v) for k, v in agent.critic.encoder.named_parameters()]
    wrap(agent.critic_optimizer, 'critic_opt', crit_named)
    wrap(agent.actor_optimizer, 'actor_opt', act_named)
    wrap(agent.log_alpha_optimizer, 'alpha_opt', [('log_alpha', agent.log_alpha)])
    wrap(agent.cpc_optimizer, 'cpc_opt', enc_named + [('W', agent.CURL.W)])

    real_randn_like = torch.randn_like
    np.random.seed(S.SAMPLING_SEED)
    for u, (step, only_cpc) in enumerate(zip(cfg['steps'], cfg['only_cpc'])):
        # what sample_cpc will draw (verified below through the frame CRCs)
        st = np.random.get_state()
        from . import curla_oracle as O
        draws = O.draw_sample_indices(cfg['capacity'], rb.idx, rb.full, cfg['B'], cfg['aug'],
                                      hw, ohw)
        np.random.set_state(st)
        queue = [noise[u, 0], noise[u, 1]]

        def fake_randn_like(t, *a, **k):
            n = queue.pop(0)
            assert n.shape == t.shape
            return n.clone()

        # capture the sampled batch by wrapping sample_cpc once
        captured = {}
        orig_sample = rb.sample_cpc

        def sample_once():
            r = orig_sample()
            captured['obs'], captured['next'], captured['pos'] = r[0], r[3], r[5]['obs_pos']
            captured['action'], captured['reward'], captured['not_done'] = r[1], r[2], r[4]
            return r
        rb.sample_cpc = sample_once
        grads.clear()
        torch.randn_like = fake_randn_like
        try:
            agent.update(rb, L, step, only_cpc=only_cpc)
        finally:
            torch.randn_like = real_randn_like
            rb.sample_cpc = orig_sample
        p = 'u%d/' % u
        for k, v in draws.items():
            out[p + 'idx/' + k] = v
        for k in ('obs', 'next', 'pos'):
            a = captured[k].numpy()
            assert np.all(a == np.round(a)) and a.min() >= 0 and a.max() <= 255
            out[p + 'crc/' + k] = np.array([zlib.crc32(a.astype(np.uint8).tobytes())],
                                           dtype=np.int64)
        for k in ('action', 'reward', 'not_done'):
            out[p + 'batch/' + k] = captured[k].numpy()
        for (s, key), val in L.rows.items():
            if s == step:
                out[p + 'metric/' + key] = np.array([val])
        L.rows.clear()
        for (tag, key), val in grads.items():
            out[p + 'grad/' + tag + '/' + key] = val
        for net, mod in (('actor', agent.actor), ('critic', agent.critic),
                         ('target', agent.critic_target)):
            for k, v in mod.state_dict().items():
                out[p + 'param/' + net + '.' + k] = S.summarize(v)
        out[p + 'param/W'] = S.summarize(agent.CURL.W)
        out[p + 'param/log_alpha'] = np.array([float(agent.log_alpha)])
        out[p + 'out/critic_z'] = agent.critic.encoder.outputs['ln'].detach().numpy()
        out[p + 'out/target_z'] = agent.critic_target.encoder.outputs['ln'].detach().numpy()
        out[p + 'out/actor_z'] = agent.actor.encoder.outputs['ln'].detach().numpy()
        out[p + 'out/q1'] = agent.critic.outputs['q1'].detach().numpy()
        out[p + 'out/q2'] = agent.critic.outputs['q2'].detach().numpy()
    return out


INIT_CASES = {
    # name -> (seed, obs_shape, hidden): fingerprints of a FRESHLY CONSTRUCTED reference agent
    # (curl_sac.py:271-317) -- pins the product's init-RNG order (curla_b200.curl_sac.initial_state)
    'init_seed0': (0, (9, 76, 135), 64),
    'init_seed3_identity': (3, (9, 90, 160), 32),
}


def run_init(seed, obs_shape, hidden, ref_utils, ref_aug, ref_sac):
    from . import scenario as S
    ref_utils.set_seed_everywhere(seed)
    augmentor = ref_aug.make_augmentor('identity', obs_shape[1:])
    agent = ref_sac.CurlSacAgent(obs_shape, (S.ACTION_DIM,), torch.device('cpu'), augmentor,
                                 hidden_dim=hidden, **S.HP)
    out = {}
    for net, mod in (('actor', agent.actor), ('critic', agent.critic), ('target', agent.critic_target)):
        for k, v in mod.state_dict().items():
            out['param/' + net + '.' + k] = S.summarize(v)
    out['param/W'] = S.summarize(agent.CURL.W)
    out['param/log_alpha'] = np.array([float(agent.log_alpha)])
    out['next_rand'] = torch.rand(4).numpy()        # where the torch stream stands afterwards
    return out


def main():
    ref_utils, ref_aug, ref_sac, _ = import_reference()
    from . import scenario as S
    gold_dir = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
    os.makedirs(gold_dir, exist_ok=True)
    names = sys.argv[1:] or list(S.SCENARIOS) + list(INIT_CASES)
    for name in names:
        if name in INIT_CASES:
            out = run_init(*INIT_CASES[name], ref_utils, ref_aug, ref_sac)
        else:
            out = run_scenario(name, S.SCENARIOS[name], ref_utils, ref_aug, ref_sac)
        path = os.path.join(gold_dir, name + '.npz')
        np.savez_compressed(path, **out)
        print('wrote', path, '%d arrays' % len(out), '%.1f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
