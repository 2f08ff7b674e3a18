"""Python handle on the C++ agent engine (curla_b200/csrc/engine.cu).

Owns the five device arenas (plain torch allocations) and exposes every tensor of the
engine's memory plan as a torch view, keyed by name.  All compute happens in the CUDA
library; this file only moves pointers around.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_DTYPES = {0: torch.float32, 1: torch.bfloat16, 2: torch.float64, 3: torch.int32, 4: torch.uint8}
ARENA_NAMES = ('params', 'shadow', 'grads', 'adam', 'work')


class Engine:
    def __init__(self, device, **cfg):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.CurlaError('curla_b200 runs on CUDA devices only (no CPU fallback); got %s'
                                  % self.device)
        c = _lib.AgentConfig()
        for k, v in cfg.items():
            if not hasattr(c, k):
                raise KeyError(k)
            setattr(c, k, v)
        self.cfg = c
        with torch.cuda.device(self.device):
            self.h = self.lib.curla_agent_create(C.byref(c))
            if not self.h:
                raise _lib.CurlaError('curla_agent_create: %s' % self.lib.curla_last_error().decode())
            self.arenas = []
            for i in range(5):
                n = self.lib.curla_agent_arena_bytes(self.h, i)
                self.arenas.append(torch.zeros(max(int(n), 256), dtype=torch.uint8, device=self.device))
            ptrs = (C.c_void_p * 5)(*[a.data_ptr() for a in self.arenas])
            _lib.check(self.lib.curla_agent_bind(self.h, ptrs), 'curla_agent_bind')
        self.t = {}          # name -> torch view
        self.info = {}       # name -> (arena, byte_offset, shape, dtype)
        name = C.create_string_buffer(256)
        arena, off, ndim, dt = C.c_int(), C.c_longlong(), C.c_int(), C.c_int()
        dims = (C.c_longlong * 4)()
        for i in range(self.lib.curla_agent_num_tensors(self.h)):
            _lib.check(self.lib.curla_agent_tensor_info(self.h, i, name, 256, C.byref(arena), C.byref(off),
                                                        C.byref(ndim), dims, C.byref(dt)), 'tensor_info')
            shape = tuple(int(dims[k]) for k in range(ndim.value))
            dtype = _DTYPES[dt.value]
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            view = self.arenas[arena.value][off.value:off.value + nbytes].view(dtype).view(shape)
            key = name.value.decode()
            self.t[key] = view
            self.info[key] = (arena.value, off.value, shape, dtype)

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.curla_agent_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------------
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def refresh_shadows(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.curla_agent_refresh_shadows(self.h, self.stream()), 'refresh_shadows')

    def update(self, args):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.curla_agent_update(self.h, C.byref(args), self.stream()), 'curla_agent_update')

    def last_launches(self):
        return self.lib.curla_agent_last_launches(self.h)

    # -- gradients of the last update, as views of the flat gradient buckets -------------
    def grad_view(self, key, bucket=None):
        """View of the gradient bucket that mirrors parameter `key` (same shape as the parameter):
        critic.* -> grad.critic (or grad.cpc with bucket='cpc': the CURL step's encoder gradient),
        actor.* -> grad.actor, CURL.W -> grad.cpc.  None for parameters that never receive one
        (target.*).  Valid until the next update overwrites the bucket."""
        arena, off, shape, _ = self.info[key]
        if arena != 0:
            return None
        n, o = int(np.prod(shape)), off // 4
        base = lambda k: self.info[k][1] // 4
        if key == 'CURL.W':
            return self.t['grad.cpc'][0:n].view(shape)
        if key.startswith('critic.'):
            if bucket == 'cpc':
                if not key.startswith('critic.encoder.'):
                    return None
                b0 = base('CURL.W')
                return self.t['grad.cpc'][o - b0:o - b0 + n].view(shape)
            b0 = base('critic.encoder.convs.0.weight')
            return self.t['grad.critic'][o - b0:o - b0 + n].view(shape)
        if key.startswith('actor.'):
            b0 = base('actor.encoder.fc.weight_canon')
            return self.t['grad.actor'][o - b0:o - b0 + n].view(shape)
        return None

    def act_nchw(self, buf, layer, n=None):
        """Decode a conv-stack activation buffer ('actA.<l>' ...: channel planes [B][4][S][8] bf16) into
        the reference's NCHW float tensor of layer `layer`'s valid output (encoder.py:81-87)."""
        c = self.cfg
        Hs, pitch = (c.H + 1) // 2, (c.W + 1) // 2
        ho, wo = (c.H - 3) // 2 + 1 - 2 * layer, (c.W - 3) // 2 + 1 - 2 * layer
        B = c.batch if n is None else n
        v = self.t['%s.%d' % (buf, layer)][:B * Hs * pitch].view(B, 4, Hs, pitch, 8)[:, :, :ho, :wo, :]
        return v.permute(0, 1, 4, 2, 3).reshape(B, 32, ho, wo).float()

    # -- canonical <-> PyTorch layout of the encoder fc weight ---------------------
    def fc_geometry(self):
        """(feat, Ho, Wo, pitch, planes): canon is [feat][planes][Ho*pitch][8] -- the K order in
        which the conv stack's channel-plane activations are read by the fc GEMM."""
        feat, planes, hp, _ = self.t['critic.encoder.fc.weight_canon'].shape
        pitch = (self.cfg.W + 1) // 2
        return feat, hp // pitch, self._wo4(), pitch, planes

    def _wo4(self):
        return (self.cfg.W - 3) // 2 + 1 - 6

    def fc_to_torch(self, canon):
        """canon -> PyTorch nn.Linear weight [feat][F*Ho*Wo] (NCHW flatten, encoder.py:89)."""
        feat, ho, wo, pitch, planes = self.fc_geometry()
        v = canon.view(feat, planes, ho, pitch, 8)[:, :, :, :wo, :]          # f, j, y, x, e
        return v.permute(0, 1, 4, 2, 3).reshape(feat, planes * 8 * ho * wo).contiguous()

    def fc_from_torch(self, weight, canon_out):
        feat, ho, wo, pitch, planes = self.fc_geometry()
        canon_out.zero_()
        w = weight.reshape(feat, planes, 8, ho, wo).permute(0, 1, 3, 4, 2)   # f, j, y, x, e
        canon_out.view(feat, planes, ho, pitch, 8)[:, :, :, :wo, :] = w.to(canon_out.dtype)
