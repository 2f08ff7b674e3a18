"""Drop-in for the reference's encoder.py (encoder.py:12-130).

CNNEncoder keeps the reference's constructor, attributes and call signature.  Its
parameters live in an agent engine's arenas (fp32 masters + bf16 kernel-layout shadows);
forward() runs the CUDA conv stack + fc + LayerNorm kernels.  Outside CurlSacAgent.update
the module is inference-only (that is how the reference's scripts use it: under no_grad
in sample_action / eval / latent extraction).
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib

# encoder.py:20-29.  (The reference's square tables hold ints and crash at encoder.py:66;
# here they are usable.)
OUT_DIM = {2: 39, 4: 35, 6: 31}
OUT_DIM_64 = {2: 29, 4: 25, 6: 21}
OUT_DIM_RECT_76_135 = {4: [31, 61], }
OUT_DIM_RECT_90_160 = {4: [38, 73], }


def tie_weights(src, trg):
    """encoder.py:12-15 (for nn.Module layers)."""
    assert type(src) == type(trg)
    trg.weight = src.weight
    trg.bias = src.bias


def out_dim_for(obs_shape, num_layers):
    hw = tuple(obs_shape[1:])
    if hw == (84, 84):
        d = OUT_DIM[num_layers]
        return [d, d]
    if hw == (64, 64):
        d = OUT_DIM_64[num_layers]
        return [d, d]
    if hw == (76, 135) and num_layers == 4:
        return OUT_DIM_RECT_76_135[num_layers]
    if hw == (90, 160) and num_layers == 4:
        return OUT_DIM_RECT_90_160[num_layers]
    raise NotImplementedError("Encoder does not support input shape")


class _ParamView(object):
    """A layer-like handle (``.weight`` / ``.bias``) on tensors stored in the engine."""

    def __init__(self, owner, prefix, is_fc=False):
        self._owner, self._prefix, self._is_fc = owner, prefix, is_fc

    def _key(self, name):
        k = self._prefix + name
        if k.startswith('actor.encoder.convs.'):           # tied to the critic's (curl_sac.py:290)
            k = 'critic' + k[len('actor'):]
        return k

    def _with_grad(self, t, key, conv=lambda x: x):
        g = self._owner._engine().grad_view(key)           # what Logger.log_param reads (logger.py:155-162)
        if g is not None:
            t.grad = conv(g)
        return t

    @property
    def weight(self):
        eng = self._owner._engine()
        if self._is_fc:
            k = self._key('weight_canon')
            return self._with_grad(eng.fc_to_torch(eng.t[k]), k, eng.fc_to_torch)
        return self._with_grad(eng.t[self._key('weight')], self._key('weight'))

    @property
    def bias(self):
        return self._with_grad(self._owner._engine().t[self._key('bias')], self._key('bias'))


class CNNEncoder(object):
    """Convolutional encoder of pixels observations (encoder.py:32-130)."""

    def __init__(self, obs_shape, feature_dim, num_layers=4, num_filters=32, output_logits=False,
                 _host=None, _net=None, _prefix=None):
        assert len(obs_shape) == 3
        self.out_dim = out_dim_for(obs_shape, num_layers)
        self.obs_shape = tuple(obs_shape)
        self.feature_dim = feature_dim
        self.num_layers = num_layers
        self.num_filters = num_filters
        self.output_logits = output_logits
        self.outputs = dict()
        self.training = True
        self._host, self._net, self._prefix = _host, _net, _prefix
        if _host is None:
            # stand-alone encoder: a private engine holds its weights (default nn init,
            # like a bare reference CNNEncoder)
            from . import curl_sac
            self._host = curl_sac._StandaloneEncoderHost(self)
            self._net, self._prefix = 1, 'critic.encoder.'
        self.convs = [_ParamView(self, '%sconvs.%d.' % (self._prefix, i)) for i in range(num_layers)]
        self.fc = _ParamView(self, self._prefix + 'fc.', is_fc=True)
        self.ln = _ParamView(self, self._prefix + 'ln.')

    # -- module-ish API ------------------------------------------------------------
    def _engine(self):
        return self._host.engine

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def to(self, device):
        return self

    def parameters(self):
        eng = self._engine()
        keys = ['convs.%d.%s' % (i, k) for i in range(self.num_layers) for k in ('weight', 'bias')]
        keys += ['fc.weight_canon', 'fc.bias', 'ln.weight', 'ln.bias']
        return [eng.t[self._param_key(k)] for k in keys]

    def _after_param_write(self):
        """fp32 masters were written from outside the engine (utils.soft_update_params on two
        encoders, utils.py:37-41): rebuild the bf16 kernel-layout shadows."""
        self._engine().refresh_shadows()

    def _param_key(self, k):
        # the actor's conv layers ARE the critic's (tied, curl_sac.py:290)
        if self._prefix.startswith('actor.') and k.startswith('convs.'):
            return 'critic.encoder.' + k
        return self._prefix + k

    def state_dict(self, prefix=''):
        eng = self._engine()
        sd = {}
        for i in range(self.num_layers):
            for k in ('weight', 'bias'):
                sd['%sconvs.%d.%s' % (prefix, i, k)] = eng.t[self._param_key('convs.%d.%s' % (i, k))].clone()
        sd[prefix + 'fc.weight'] = eng.fc_to_torch(eng.t[self._param_key('fc.weight_canon')])
        sd[prefix + 'fc.bias'] = eng.t[self._param_key('fc.bias')].clone()
        sd[prefix + 'ln.weight'] = eng.t[self._param_key('ln.weight')].clone()
        sd[prefix + 'ln.bias'] = eng.t[self._param_key('ln.bias')].clone()
        return sd

    def load_state_dict(self, sd, prefix='', refresh=True):
        eng = self._engine()
        for i in range(self.num_layers):
            for k in ('weight', 'bias'):
                eng.t[self._param_key('convs.%d.%s' % (i, k))].copy_(sd['%sconvs.%d.%s' % (prefix, i, k)])
        eng.fc_from_torch(sd[prefix + 'fc.weight'].to(eng.device), eng.t[self._param_key('fc.weight_canon')])
        for k in ('fc.bias', 'ln.weight', 'ln.bias'):
            eng.t[self._param_key(k)].copy_(sd[prefix + k])
        if refresh:
            eng.refresh_shadows()

    def copy_conv_weights_from(self, source):
        """Tie convolutional layers (encoder.py:112-116).  Inside an agent the actor's conv
        tensors are the critic's by construction; this copies values for any other pair."""
        if self._param_key('convs.0.weight') == source._param_key('convs.0.weight') and \
                self._host is source._host:
            return
        for i in range(self.num_layers):
            for k in ('weight', 'bias'):
                self._engine().t[self._param_key('convs.%d.%s' % (i, k))].copy_(
                    source._engine().t[source._param_key('convs.%d.%s' % (i, k))])
        self._engine().refresh_shadows()

    # -- forward -------------------------------------------------------------------
    def forward_conv(self, obs):
        raise NotImplementedError('forward_conv alone is not exposed; call forward()')

    def forward(self, obs, detach=False):
        """obs: float tensor (B, C, H, W) with values in [0, 255] (encoder.py:77-110)."""
        z = self._host.encode(self._net, obs, apply_tanh=not self.output_logits)
        self.outputs['ln' if self.output_logits else 'tanh'] = z
        return z

    __call__ = forward

    def log(self, L, step, log_freq):
        if step % log_freq != 0:
            return
        for k, v in self.outputs.items():
            L.log_histogram('train_encoder/%s_hist' % k, v, step)
        for i in range(self.num_layers):
            L.log_param('train_encoder/conv%s' % (i + 1), self.convs[i], step)
        L.log_param('train_encoder/fc', self.fc, step)
        L.log_param('train_encoder/ln', self.ln, step)


PixelEncoder = CNNEncoder   # upstream MishaLaskin/curl name used in BASELINE.json
