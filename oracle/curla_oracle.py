"""CPU oracle for the CurlSacAgent.update hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain restatement (numpy for byte/index work, torch fp32 on CPU
for the floating-point maths) of the reference algorithm.  It is *not* the
product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import it.  The product path (curla_b200/) never
imports anything from oracle/.

Pinning: the reference ships no tests / golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by oracle/make_golden.py (imports /root/reference through three
import shims) and committed under tests/golden/.  tests/test_oracle_golden.py
re-checks the oracle against those fixtures on every run.  The kornia-based
augmentations (color_jiggle / noisy_cover) are "parity unpinned": kornia is not
available anywhere in this environment, see DESIGN.md.

Every function cites the reference file:line it restates (paths relative to
the reference repo root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# sampling / gather / crop  (integer + byte work: numpy, bit-exact)
# --------------------------------------------------------------------------


def crop_output_shape(input_hw, factor=0.84):
    """augmentations.py:21-24 -- output_shape = ceil(0.84 * dim)."""
    return tuple(int(np.ceil(x * factor)) for x in input_hw)


def center_crop_offsets(input_hw, output_hw):
    """augmentations.py:37-40 -- top/left of the evaluation centre crop."""
    return (input_hw[0] - output_hw[0]) // 2, (input_hw[1] - output_hw[1]) // 2


def draw_sample_indices(capacity, idx, full, batch_size, augmentation, input_hw,
                        output_hw):
    """Consume the numpy GLOBAL RNG exactly as sample_cpc does.

    utils.py:147 draws idxs; for RandomCrop utils.py:156-158 calls
    training_augmentation three times (obs, next_obs, pos), each drawing
    h1 then w1 with an EXCLUSIVE upper bound of (in - out)
    (augmentations.py:63-67).  NoisyCover draws three randint(0,255) scalars
    per call (augmentations.py:192-194), again obs, next_obs, pos.
    Returns a dict of int64 arrays.
    """
    out = {}
    out['idxs'] = np.random.randint(0, capacity if full else idx, size=batch_size)
    if augmentation == 'random_crop':
        ch = input_hw[0] - output_hw[0]
        cw = input_hw[1] - output_hw[1]
        for name in ('obs', 'next', 'pos'):
            out['h1_' + name] = np.random.randint(0, ch, batch_size)
            out['w1_' + name] = np.random.randint(0, cw, batch_size)
    elif augmentation == 'noisy_cover':
        for name in ('obs', 'next', 'pos'):
            out['cover_' + name] = np.array(
                [np.random.randint(0, 255) for _ in range(3)], dtype=np.int64)
    return out


def gather_crop(frames, idxs, h1, w1, out_hw):
    """utils.py:151-152 + augmentations.py:47-75.

    frames (cap, C, H, W) uint8; returns (B, C, oh, ow) uint8 with
    out[i] = frames[idxs[i], :, h1[i]:h1[i]+oh, w1[i]:w1[i]+ow].
    (view_as_windows(...)[arange, h1, w1] is exactly this slice)
    """
    oh, ow = out_hw
    b = len(idxs)
    out = np.empty((b, frames.shape[1], oh, ow), dtype=frames.dtype)
    for i in range(b):
        out[i] = frames[idxs[i], :, h1[i]:h1[i] + oh, w1[i]:w1[i] + ow]
    return out


def gather(frames, idxs):
    """utils.py:171-172 -- plain fancy-index gather (identity branch)."""
    return frames[idxs]


def noisy_cover(batch_f32, cover_values, noise, top, bottom):
    """augmentations.py:172-205 with the Gaussian noise tensor passed in.

    batch (B, 3*fs, H, W) float32.  rows [0,top) and [H-bottom,H) of every
    R/G/B plane are filled with cover_values[0..2]; + noise (std already
    applied); clamp to [0,255].
    """
    b, c, h, w = batch_f32.shape
    x = batch_f32.reshape(-1, 3, h, w).clone()
    rows = np.concatenate((np.arange(0, top), np.arange(h - bottom, h)))
    for ch in range(3):
        x[:, ch, rows, :] = float(cover_values[ch])
    x = x + noise.reshape(x.shape)
    return torch.clamp(x.reshape(b, c, h, w), 0, 255)


def noisy_cover_rows(h):
    """augmentations.py:143-147 -- top = ceil(0.31 h), bottom = ceil(0.20 h)."""
    return int(np.ceil(h * 0.31)), int(np.ceil(h * 0.20))


# ---- colour jiggle: restated from kornia's documented behaviour (unpinned) ----

def rgb_to_hsv(rgb):
    """kornia.color.rgb_to_hsv semantics: h in [0, 2pi), s,v in [0,1]. rgb (...,3,H,W)."""
    r, g, b = rgb[..., 0, :, :], rgb[..., 1, :, :], rgb[..., 2, :, :]
    maxc, _ = rgb.max(-3)
    minc, _ = rgb.min(-3)
    v = maxc
    delta = maxc - minc
    s = delta / (maxc + 1e-8)
    d = torch.where(delta == 0, torch.ones_like(delta), delta)
    rc, gc, bc = (maxc - r), (maxc - g), (maxc - b)
    h = torch.where(maxc == r, bc - gc,
                    torch.where(maxc == g, 2.0 * d + rc - bc, 4.0 * d + gc - rc))
    h = (h / d / 6.0) % 1.0
    return torch.stack([h * 2.0 * math.pi, s, v], dim=-3)


def hsv_to_rgb(hsv):
    """kornia.color.hsv_to_rgb semantics."""
    h = hsv[..., 0, :, :] / (2.0 * math.pi)
    s = hsv[..., 1, :, :]
    v = hsv[..., 2, :, :]
    hi = torch.floor(h * 6.0) % 6
    f = (h * 6.0) % 6 - hi
    p = v * (1 - s)
    q = v * (1 - f * s)
    t = v * (1 - (1 - f) * s)
    hi = hi.long()
    r = torch.where(hi == 0, v, torch.where(hi == 1, q, torch.where(hi == 2, p,
        torch.where(hi == 3, p, torch.where(hi == 4, t, v)))))
    g = torch.where(hi == 0, t, torch.where(hi == 1, v, torch.where(hi == 2, v,
        torch.where(hi == 3, q, torch.where(hi == 4, p, p)))))
    b = torch.where(hi == 0, p, torch.where(hi == 1, p, torch.where(hi == 2, t,
        torch.where(hi == 3, v, torch.where(hi == 4, v, q)))))
    return torch.stack([r, g, b], dim=-3)


def color_jiggle(batch_f32, contrast, saturation, hue, apply_mask, order):
    """augmentations.py:106-136 with kornia's sampled parameters passed in.

    batch (B, 3*fs, H, W) float32 in [0,255]; contrast/saturation/hue/apply_mask
    are per-IMAGE (B*fs) vectors; `order` is a permutation of (0:brightness,
    1:contrast, 2:saturation, 3:hue) shared by the call (kornia samples one
    order per forward).  brightness factor is 0 (additive) => no-op.
    PARITY UNPINNED (kornia source/installation unavailable): follows the
    documented kornia 0.6/0.7 ColorJiggle: contrast = clamp(x*c,0,1),
    saturation = HSV s*=f (clamp), hue = HSV h += f*2pi (mod 2pi).
    """
    b, c, h, w = batch_f32.shape
    x = (batch_f32 / 255.0).reshape(-1, 3, h, w)
    y = x.clone()
    cf = contrast.view(-1, 1, 1, 1)
    for op in order:
        if op == 1:
            y = torch.clamp(y * cf, 0.0, 1.0)
        elif op == 2:
            hsv = rgb_to_hsv(y)
            s = torch.clamp(hsv[:, 1] * saturation.view(-1, 1, 1), 0.0, 1.0)
            y = hsv_to_rgb(torch.stack([hsv[:, 0], s, hsv[:, 2]], 1))
        elif op == 3:
            hsv = rgb_to_hsv(y)
            hh = torch.fmod(hsv[:, 0] + hue.view(-1, 1, 1) * 2.0 * math.pi, 2.0 * math.pi)
            hh = torch.where(hh < 0, hh + 2.0 * math.pi, hh)
            y = hsv_to_rgb(torch.stack([hh, hsv[:, 1], hsv[:, 2]], 1))
    m = apply_mask.view(-1, 1, 1, 1).to(torch.bool)
    y = torch.where(m, y, x)
    return (y.reshape(b, c, h, w) * 255.0)


# --------------------------------------------------------------------------
# networks (functional; parameters are dicts with the reference's
# state-dict key names)
# --------------------------------------------------------------------------

OUT_DIMS = {  # encoder.py:20-29
    (84, 84): {2: (39, 39), 4: (35, 35), 6: (31, 31)},
    (64, 64): {2: (29, 29), 4: (25, 25), 6: (21, 21)},
    (76, 135): {4: (31, 61)},
    (90, 160): {4: (38, 73)},
}


def encoder_forward(p, prefix, obs, num_layers=4, detach=False, output_logits=True,
                    keep=None):
    """encoder.py:77-110.  p[prefix+'convs.i.weight'] etc.  obs float32 (B,C,H,W)."""
    x = obs / 255.0                                                   # encoder.py:78
    x = torch.relu(F.conv2d(x, p[prefix + 'convs.0.weight'], p[prefix + 'convs.0.bias'],
                            stride=2))                                # encoder.py:81
    if keep is not None:
        keep['conv1'] = x
    for i in range(1, num_layers):                                    # encoder.py:84-87
        x = torch.relu(F.conv2d(x, p[prefix + 'convs.%d.weight' % i],
                                p[prefix + 'convs.%d.bias' % i], stride=1))
        if keep is not None:
            keep['conv%d' % (i + 1)] = x
    h = x.reshape(x.size(0), -1)                                      # encoder.py:89 (NCHW flatten)
    if detach:
        h = h.detach()                                                # encoder.py:95-96
    h_fc = F.linear(h, p[prefix + 'fc.weight'], p[prefix + 'fc.bias'])  # encoder.py:98
    if keep is not None:
        keep['fc'] = h_fc
    z = F.layer_norm(h_fc, (h_fc.size(-1),), p[prefix + 'ln.weight'], p[prefix + 'ln.bias'],
                     eps=1e-5)                                        # encoder.py:101
    if not output_logits:
        z = torch.tanh(z)                                             # encoder.py:106
    return z


def mlp3(p, prefix, x):
    """nn.Sequential(Linear, ReLU, Linear, ReLU, Linear): curl_sac.py:70-74,129-133."""
    x = torch.relu(F.linear(x, p[prefix + '0.weight'], p[prefix + '0.bias']))
    x = torch.relu(F.linear(x, p[prefix + '2.weight'], p[prefix + '2.bias']))
    return F.linear(x, p[prefix + '4.weight'], p[prefix + '4.bias'])


def gaussian_logprob(noise, log_std):
    """curl_sac.py:20-23."""
    residual = (-0.5 * noise.pow(2) - log_std).sum(-1, keepdim=True)
    return residual - 0.5 * np.log(2 * np.pi) * noise.size(-1)


def actor_forward(p, obs, noise, log_std_min, log_std_max, detach_encoder=False,
                  compute_pi=True, compute_log_pi=True):
    """curl_sac.py:79-110 with the policy noise INJECTED (reference: randn_like, :97).

    p holds actor keys 'encoder.*' and 'trunk.*'."""
    z = encoder_forward(p, 'encoder.', obs, detach=detach_encoder)
    mu, log_std = mlp3(p, 'trunk.', z).chunk(2, dim=-1)
    log_std = torch.tanh(log_std)
    log_std = log_std_min + 0.5 * (log_std_max - log_std_min) * (log_std + 1)
    pi = log_pi = None
    if compute_pi:
        pi = mu + noise * log_std.exp()
    if compute_log_pi:
        log_pi = gaussian_logprob(noise, log_std)
    mu = torch.tanh(mu)                                               # squash, curl_sac.py:26-35
    if pi is not None:
        pi = torch.tanh(pi)
    if log_pi is not None:
        log_pi = log_pi - torch.log(F.relu(1 - pi.pow(2)) + 1e-6).sum(-1, keepdim=True)
    return mu, pi, log_pi, log_std


def critic_forward(p, obs, action, detach_encoder=False, keep=None):
    """curl_sac.py:158-169 / 135-139.  p holds 'encoder.*', 'Q1.trunk.*', 'Q2.trunk.*'."""
    z = encoder_forward(p, 'encoder.', obs, detach=detach_encoder, keep=keep)
    if keep is not None:
        keep['z'] = z
    za = torch.cat([z, action], dim=1)
    return mlp3(p, 'Q1.trunk.', za), mlp3(p, 'Q2.trunk.', za)


def curl_logits(W, z_a, z_pos):
    """curl_sac.py:211-222."""
    Wz = torch.matmul(W, z_pos.T)
    logits = torch.matmul(z_a, Wz)
    return logits - torch.max(logits, 1)[0][:, None]


# --------------------------------------------------------------------------
# Adam / EMA
# --------------------------------------------------------------------------

class Adam:
    """torch.optim.Adam (eps 1e-8, no weight decay, no amsgrad) restated.

    curl_sac.py:299-313 creates five of these.  Parameters whose grad is None
    are skipped (no state created), exactly like torch.optim.Adam.
    """

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.state = {}

    def zero_grad(self):
        for q in self.params:
            q.grad = None

    @torch.no_grad()
    def step(self):
        for q in self.params:
            if q.grad is None:
                continue
            st = self.state.setdefault(id(q), dict(
                t=0, m=torch.zeros_like(q), v=torch.zeros_like(q)))
            st['t'] += 1
            g = q.grad
            st['m'].mul_(self.b1).add_(g, alpha=1 - self.b1)
            st['v'].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1 = 1 - self.b1 ** st['t']
            bc2 = 1 - self.b2 ** st['t']
            denom = (st['v'].sqrt() / math.sqrt(bc2)).add_(self.eps)
            q.addcdiv_(st['m'], denom, value=-(self.lr / bc1))


def soft_update(params, src_prefix, dst_params, dst_prefix, tau):
    """utils.py:37-41 over every key under src_prefix."""
    with torch.no_grad():
        for k, v in params.items():
            if k.startswith(src_prefix):
                t = dst_params[dst_prefix + k[len(src_prefix):]]
                t.copy_(tau * v + (1 - tau) * t)


# --------------------------------------------------------------------------
# the agent
# --------------------------------------------------------------------------

def _linear_init(out_f, in_f, gen):
    w = torch.empty(out_f, in_f)
    torch.nn.init.orthogonal_(w, generator=gen)
    return w


class OracleAgent:
    """Functional restatement of curl_sac.CurlSacAgent (curl_sac.py:224-465).

    State lives in three dicts with the reference's state-dict key names:
      actor   : encoder.{convs.i,fc,ln}.{weight,bias}, trunk.{0,2,4}.{weight,bias}
      critic  : encoder.*, Q1.trunk.*, Q2.trunk.*
      target  : same keys as critic
    plus W (CURL.W) and log_alpha (float64 0-dim).  The actor's conv tensors ARE
    the critic's (curl_sac.py:290; encoder.py:112-116).
    Initial weights are normally loaded from a state-dict (never re-derived
    from the init RNG); `init_random` exists for benchmarking only.
    """

    def __init__(self, obs_shape, action_dim, hidden_dim=1024, discount=0.99,
                 init_temperature=0.1, alpha_lr=1e-4, alpha_beta=0.5, actor_lr=1e-3,
                 actor_beta=0.9, actor_log_std_min=-10, actor_log_std_max=2,
                 actor_update_freq=2, critic_lr=1e-3, critic_beta=0.9, critic_tau=0.01,
                 critic_target_update_freq=2, encoder_feature_dim=50, encoder_lr=1e-3,
                 encoder_tau=0.05, num_layers=4, num_filters=32, cpc_update_freq=1,
                 detach_encoder=False, pixel_sac=False):
        self.obs_shape = tuple(obs_shape)
        self.action_dim = action_dim
        self.hidden_dim = hidden_dim
        self.discount = discount
        self.log_std_min, self.log_std_max = actor_log_std_min, actor_log_std_max
        self.actor_update_freq = actor_update_freq
        self.critic_tau, self.encoder_tau = critic_tau, encoder_tau
        self.critic_target_update_freq = critic_target_update_freq
        self.cpc_update_freq = cpc_update_freq
        self.feature_dim = encoder_feature_dim
        self.num_layers, self.num_filters = num_layers, num_filters
        self.detach_encoder, self.pixel_sac = detach_encoder, pixel_sac
        self.target_entropy = -float(action_dim)                      # curl_sac.py:296
        self.hp = dict(alpha_lr=alpha_lr, alpha_beta=alpha_beta, actor_lr=actor_lr,
                       actor_beta=actor_beta, critic_lr=critic_lr, critic_beta=critic_beta,
                       encoder_lr=encoder_lr)
        self.log_alpha = torch.tensor(np.log(init_temperature), dtype=torch.float64,
                                      requires_grad=True)             # curl_sac.py:292
        self.actor, self.critic, self.target, self.W = {}, {}, {}, None
        self.metrics = {}

    # -- construction -------------------------------------------------------
    def out_dim(self):
        return OUT_DIMS[self.obs_shape[1:]][self.num_layers]

    def init_random(self, seed=0):
        """weight_init (curl_sac.py:38-54) + CURL.W = rand (curl_sac.py:192)."""
        g = torch.Generator().manual_seed(seed)
        nf, fd, hd, ad = self.num_filters, self.feature_dim, self.hidden_dim, self.action_dim
        oh, ow = self.out_dim()

        def enc():
            d = {}
            cin = self.obs_shape[0]
            for i in range(self.num_layers):
                w = torch.zeros(nf, cin, 3, 3)
                c = torch.empty(nf, cin)
                torch.nn.init.orthogonal_(c, gain=math.sqrt(2.0), generator=g)
                w[:, :, 1, 1] = c
                d['encoder.convs.%d.weight' % i] = w
                d['encoder.convs.%d.bias' % i] = torch.zeros(nf)
                cin = nf
            d['encoder.fc.weight'] = _linear_init(fd, nf * oh * ow, g)
            d['encoder.fc.bias'] = torch.zeros(fd)
            d['encoder.ln.weight'] = torch.ones(fd)
            d['encoder.ln.bias'] = torch.zeros(fd)
            return d

        def mlp(prefix, i, o):
            return {prefix + '0.weight': _linear_init(hd, i, g), prefix + '0.bias': torch.zeros(hd),
                    prefix + '2.weight': _linear_init(hd, hd, g), prefix + '2.bias': torch.zeros(hd),
                    prefix + '4.weight': _linear_init(o, hd, g), prefix + '4.bias': torch.zeros(o)}

        actor = enc()
        actor.update(mlp('trunk.', fd, 2 * ad))
        critic = enc()
        critic.update(mlp('Q1.trunk.', fd + ad, 1))
        critic.update(mlp('Q2.trunk.', fd + ad, 1))
        W = torch.rand(fd, fd, generator=g)
        self.load_state(actor, critic, W)

    def load_state(self, actor_sd, critic_sd, W, target_sd=None, log_alpha=None):
        """Adopt reference state-dicts; ties conv layers; target defaults to a
        copy of the critic (curl_sac.py:287,290)."""
        self.critic = {k: v.detach().clone().float().requires_grad_(True)
                       for k, v in critic_sd.items()}
        self.actor = {}
        for k, v in actor_sd.items():
            if k.startswith('encoder.convs.'):
                self.actor[k] = self.critic[k]                        # tied: same tensor
            else:
                self.actor[k] = v.detach().clone().float().requires_grad_(True)
        src = critic_sd if target_sd is None else target_sd
        self.target = {k: v.detach().clone().float() for k, v in src.items()}
        self.W = W.detach().clone().float().requires_grad_(True)
        if log_alpha is not None:
            self.log_alpha = torch.tensor(float(log_alpha), dtype=torch.float64,
                                          requires_grad=True)
        hp = self.hp
        enc_keys = [k for k in self.critic if k.startswith('encoder.')]
        # curl_sac.py:299-313.  Parameter ORDER inside an optimizer is irrelevant
        # to the maths; membership is what matters.
        self.actor_opt = Adam(self.actor.values(), hp['actor_lr'], (hp['actor_beta'], 0.999))
        self.critic_opt = Adam(self.critic.values(), hp['critic_lr'], (hp['critic_beta'], 0.999))
        self.alpha_opt = Adam([self.log_alpha], hp['alpha_lr'], (hp['alpha_beta'], 0.999))
        self.encoder_opt = Adam([self.critic[k] for k in enc_keys], hp['encoder_lr'])
        # CURL.parameters() = W + critic.encoder + critic_target.encoder; target
        # tensors never get a grad so Adam skips them (SURVEY 3.3-4 ii).
        self.cpc_opt = Adam([self.W] + [self.critic[k] for k in enc_keys], hp['encoder_lr'])

    @property
    def alpha(self):
        return self.log_alpha.exp()

    # -- the update (curl_sac.py:349-451) ------------------------------------
    def update_critic(self, obs, action, reward, next_obs, not_done, noise_next):
        with torch.no_grad():                                         # curl_sac.py:350-355
            _, pol_a, log_pi, _ = actor_forward(self.actor, next_obs, noise_next,
                                                self.log_std_min, self.log_std_max)
            tq1, tq2 = critic_forward(self.target, next_obs, pol_a)
            target_v = torch.min(tq1, tq2) - self.alpha.detach() * log_pi
            target_q = reward + (not_done * self.discount * target_v)
        keep = {}
        q1, q2 = critic_forward(self.critic, obs, action,
                                detach_encoder=self.detach_encoder, keep=keep)
        loss = F.mse_loss(q1, target_q) + F.mse_loss(q2, target_q)   # curl_sac.py:359
        self.critic_opt.zero_grad()
        loss.backward()
        self.dbg = dict(target_q=target_q, q1=q1.detach(), q2=q2.detach(),
                        z_critic=keep['z'].detach(), next_action=pol_a, next_log_pi=log_pi,
                        critic_grads={k: v.grad.clone() for k, v in self.critic.items()
                                      if v.grad is not None})
        self.critic_opt.step()
        self.metrics['critic_loss'] = float(loss.detach())
        return loss

    def update_actor_and_alpha(self, obs, noise_cur):
        _, pi, log_pi, log_std = actor_forward(self.actor, obs, noise_cur, self.log_std_min,
                                               self.log_std_max, detach_encoder=True)
        aq1, aq2 = critic_forward(self.critic, obs, pi, detach_encoder=True)
        actor_q = torch.min(aq1, aq2)
        actor_loss = (self.alpha.detach() * log_pi - actor_q).mean()  # curl_sac.py:379
        entropy = 0.5 * log_std.shape[1] * (1.0 + np.log(2 * np.pi)) + log_std.sum(dim=-1)
        self.actor_opt.zero_grad()
        for v in self.critic.values():       # critic params collect (unused) grads here; keep
            v.grad = None                    # them out of later optimizer steps like zero_grad does
        actor_loss.backward()
        self.dbg.update(pi=pi.detach(), log_pi=log_pi.detach(),
                        actor_grads={k: v.grad.clone() for k, v in self.actor.items()
                                     if v.grad is not None and not k.startswith('encoder.convs.')})
        self.actor_opt.step()
        self.alpha_opt.zero_grad()
        alpha_loss = (self.alpha * (-log_pi - self.target_entropy).detach()).mean()
        self.metrics['alpha'] = float(self.alpha.detach())            # logged before the step: curl_sac.py:400-402
        alpha_loss.backward()
        self.alpha_opt.step()
        self.metrics.update(actor_loss=float(actor_loss.detach()),
                            entropy=float(entropy.mean().detach()),
                            alpha_loss=float(alpha_loss.detach()))

    def update_cpc(self, obs_anchor, obs_pos):
        z_a = encoder_forward(self.critic, 'encoder.', obs_anchor)    # curl_sac.py:408
        with torch.no_grad():
            z_pos = encoder_forward(self.target, 'encoder.', obs_pos)  # curl_sac.py:409
        logits = curl_logits(self.W, z_a, z_pos)
        labels = torch.arange(logits.shape[0]).long()
        loss = F.cross_entropy(logits, labels)                        # curl_sac.py:413
        self.encoder_opt.zero_grad()
        self.cpc_opt.zero_grad()
        loss.backward()
        self.dbg.update(z_a=z_a.detach(), z_pos=z_pos, W_grad=self.W.grad.clone(),
                        cpc_grads={k: v.grad.clone() for k, v in self.critic.items()
                                   if v.grad is not None and k.startswith('encoder.')})
        self.encoder_opt.step()                                       # curl_sac.py:419
        self.cpc_opt.step()                                           # curl_sac.py:420
        self.metrics['curl_loss'] = float(loss.detach())

    def update(self, obs, action, reward, next_obs, not_done, pos, step, noise_next,
               noise_cur, only_cpc=False):
        """curl_sac.py:426-451 after sample_cpc; tensors are float32 CPU."""
        self.dbg = {}
        self.metrics = {'batch_reward': float(reward.mean())}
        if not only_cpc:
            self.update_critic(obs, action, reward, next_obs, not_done, noise_next)
            if step % self.actor_update_freq == 0:
                self.update_actor_and_alpha(obs, noise_cur)
            if step % self.critic_target_update_freq == 0:
                soft_update(self.critic, 'Q1.', self.target, 'Q1.', self.critic_tau)
                soft_update(self.critic, 'Q2.', self.target, 'Q2.', self.critic_tau)
                soft_update(self.critic, 'encoder.', self.target, 'encoder.', self.encoder_tau)
        if not self.pixel_sac and step % self.cpc_update_freq == 0:
            self.update_cpc(obs, pos)
        return self.metrics
