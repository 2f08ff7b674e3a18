"""Per-kernel parity tests through the C ABI (run on the B200 box: pytest -m gpu).

Integer / byte / index work must be bit exact.  Tensor-core kernels compute with bf16
operands and fp32 accumulation: compared against fp32 torch references evaluated on the
SAME bf16-rounded operands (tolerance 2e-3 relative L2: accumulation-order noise only),
and against the un-rounded fp32 reference at 2e-2 (bf16 operand rounding)."""
import ctypes as C
import ctypes as C_

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from curla_b200 import _lib
from oracle import curla_oracle as O

from helpers import DEV, Geom, bf16r, max_rel, pack_conv_w, rel_l2, stream

pytestmark = pytest.mark.gpu


def dv(a, dtype=None):
    t = torch.as_tensor(a)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV).contiguous()


# ------------------------------------------------------------------ K1 gather / crop
@pytest.mark.parametrize('B,crop', [(5, True), (3, False), (33, True)])
def test_gather_crop_bit_exact(B, crop):
    rs = np.random.RandomState(0)
    cap, Cc, Hf, Wf = 11, 9, 90, 160
    frames = rs.randint(0, 256, size=(cap, Cc, Hf, Wf), dtype=np.uint8)
    idxs = rs.randint(0, cap, size=B)
    oh, ow = (76, 135) if crop else (Hf, Wf)
    h1 = rs.randint(0, Hf - oh + 1, size=B)
    w1 = rs.randint(0, Wf - ow + 1, size=B)
    ref = O.gather_crop(frames, idxs, h1, w1, (oh, ow))
    fr, di, dh, dw = dv(frames), dv(idxs), dv(h1), dv(w1)     # keep the device copies alive
    out = torch.empty((B, Cc, oh, ow), dtype=torch.float32, device=DEV)
    _lib.call('curla_gather_crop_f32', _lib.ptr(fr), Cc, Hf, Wf, _lib.ptr(di), _lib.ptr(dh),
              _lib.ptr(dw), B, oh, ow, _lib.ptr(out), stream())
    assert torch.equal(out.cpu(), torch.from_numpy(ref).float())
    # space-to-depth bf16 variant: exact as well (uint8 is exact in bf16)
    g = Geom(oh, ow, B)
    full, view = g.alloc(g.CP1)
    _lib.call('curla_gather_crop_s2d', _lib.ptr(fr), Cc, Hf, Wf, _lib.ptr(di), _lib.ptr(dh),
              _lib.ptr(dw), B, oh, ow, g.CP1, g.S * g.CP1, _lib.ptr(view), stream())
    want = g.s2d_ref(torch.from_numpy(ref).to(DEV))
    assert torch.equal(g.nhwc(view).float(), want)
    assert float(full[:g.PAD].abs().sum()) == 0 and float(full[g.PAD + B * g.S:].abs().sum()) == 0
    # float-input variant
    full2, view2 = g.alloc(g.CP1)
    _lib.call('curla_f32_to_s2d', _lib.ptr(out), Cc, oh, ow, B, g.CP1, g.S * g.CP1, _lib.ptr(view2), stream())
    assert torch.equal(view2, view)


@pytest.mark.parametrize('bulk', ['1', '0'])
@pytest.mark.parametrize('hw,out_hw,C', [((90, 160), (76, 135), 9), ((90, 160), (90, 160), 9), ((48, 64), (41, 33), 3),
                                         ((90, 160), (75, 134), 12)])
def test_gather_multi_stream_bit_exact(hw, out_hw, C, bulk, monkeypatch):
    """obs / pos / next_obs windows + the batch's action / reward / not_done rows in ONE launch (frames staged
    with cp.async.bulk, magic-number uint8 -> bf16) == the oracle's gather + crop, bit for bit; odd sizes
    (odd H and W, a lone last channel, right-edge groups) included.  bulk = '0': the per-stream fallback."""
    monkeypatch.setenv('CURLA_GATHER_BULK', bulk)
    rs = np.random.RandomState(3)
    cap, B = 13, 7
    Hf, Wf = hw
    oh, ow = out_hw
    f_obs = rs.randint(0, 256, size=(cap, C, Hf, Wf), dtype=np.uint8)
    f_next = rs.randint(0, 256, size=(cap, C, Hf, Wf), dtype=np.uint8)
    acts = rs.uniform(-1, 1, size=(cap, 2)).astype(np.float32)
    rews = rs.standard_normal((cap, 1)).astype(np.float32)
    nds = (rs.uniform(size=(cap, 1)) > 0.3).astype(np.float32)
    idxs = rs.randint(0, cap, size=B)
    offs = [(rs.randint(0, Hf - oh + 1, size=B), rs.randint(0, Wf - ow + 1, size=B)) for _ in range(3)]
    d_obs, d_next, d_idx = dv(f_obs), dv(f_next), dv(idxs)
    d_offs = [(dv(h), dv(w)) for h, w in offs]
    g = Geom(oh, ow, B, C)
    bufs = [g.alloc(g.CP1) for _ in range(3)]
    segs = (_lib.GatherSeg * 3)()
    for k, src in enumerate((d_obs, d_obs, d_next)):
        segs[k].frames = src.data_ptr(); segs[k].h1 = d_offs[k][0].data_ptr(); segs[k].w1 = d_offs[k][1].data_ptr()
        segs[k].out = bufs[k][1].data_ptr()
    rows = _lib.GatherRows()
    d_a, d_r, d_n = dv(acts), dv(rews), dv(nds)
    o_a = torch.zeros((B, 2), device=DEV); o_r = torch.zeros(B, device=DEV); o_n = torch.zeros(B, device=DEV)
    rows.actions, rows.rewards, rows.not_dones = d_a.data_ptr(), d_r.data_ptr(), d_n.data_ptr()
    rows.out_actions, rows.out_rewards, rows.out_not_dones, rows.action_dim = o_a.data_ptr(), o_r.data_ptr(), o_n.data_ptr(), 2
    _lib.call('curla_gather_crop_s2d_multi', C_.byref(segs), 3, _lib.ptr(d_idx), C, Hf, Wf, B, oh, ow, g.CP1, g.S * g.CP1,
              C_.byref(rows), stream())
    torch.cuda.synchronize()
    for k, frames in enumerate((f_obs, f_obs, f_next)):
        ref = O.gather_crop(frames, idxs, offs[k][0], offs[k][1], (oh, ow))
        want = g.s2d_ref(torch.from_numpy(ref).to(DEV))
        assert torch.equal(g.nhwc(bufs[k][1]).float(), want), k
        full = bufs[k][0]
        assert float(full[:g.PAD].abs().sum()) == 0 and float(full[g.PAD + B * g.S:].abs().sum()) == 0
    assert np.array_equal(o_a.cpu().numpy(), acts[idxs]) and np.array_equal(o_r.cpu().numpy(), rews[idxs, 0])
    assert np.array_equal(o_n.cpu().numpy(), nds[idxs, 0])


def test_gather_rows():
    rs = np.random.RandomState(1)
    src = rs.standard_normal((17, 2)).astype(np.float32)
    idxs = rs.randint(0, 17, size=9)
    out = torch.empty((9, 2), dtype=torch.float32, device=DEV)
    ds, di = dv(src), dv(idxs)
    _lib.call('curla_gather_rows_f32', _lib.ptr(ds), _lib.ptr(di), 9, 2, _lib.ptr(out), stream())
    assert np.array_equal(out.cpu().numpy(), src[idxs])


# ------------------------------------------------------------------ conv stack
def _conv_case(H, W, B, seed=0):
    torch.manual_seed(seed)
    g = Geom(H, W, B)
    x = torch.randint(0, 256, (B, 9, H, W)).float()
    ws = [torch.randn(32, 9, 3, 3) * (2.0 / 81) ** 0.5] + [torch.randn(32, 32, 3, 3) * (2.0 / 288) ** 0.5
                                                            for _ in range(3)]
    bs = [torch.randn(32) * 0.05 for _ in range(4)]
    return g, x, ws, bs


def _run_conv_stack(g, x, ws, bs):
    """our kernels: returns per-layer logical views (bf16 pitch layout) + s2d buffer."""
    full_s, s2d = g.alloc(g.CP1)
    xd = x.to(DEV)
    _lib.call('curla_f32_to_s2d', _lib.ptr(xd), 9, g.H, g.W, g.B, g.CP1, g.S * g.CP1, _lib.ptr(s2d), stream())
    acts, keep = [], [full_s]
    inp, stride = s2d, g.S * g.CP1
    for l in range(4):
        wsh = pack_conv_w(ws[l], l == 0)
        full, out = g.alloc(32)
        bd = bs[l].to(DEV)
        keep.append(bd)
        _lib.call('curla_conv_fwd', _lib.ptr(inp), stride, _lib.ptr(wsh), _lib.ptr(bd),
                  1.0 / 255.0 if l == 0 else 1.0, _lib.ptr(out), g.S * 32, g.B, g.pitch, g.S, g.Ho[l], g.Wo[l],
                  1 if l == 0 else 0, stream())
        acts.append(out)
        keep += [full, wsh]
        inp, stride = out, g.S * 32
    return s2d, acts, keep


@pytest.mark.parametrize('H,W,Bs', [(76, 135, (3, 2, 4)), (76, 135, (330,)), (64, 64, (5, 3)), (84, 84, (7,))])
def test_conv_stack_fused_equals_layer_by_layer(H, W, Bs, monkeypatch):
    """curla_conv_stack_fwd (conv-2..4 in ONE launch, each sample resident in shared memory, layers in place, weights
    packed by curla_pack_shadows kind 3) against the three launches of the N = 96 kernel it shares its arithmetic with:
    bit-identical outputs; passes that do not keep conv-2 / conv-3 leave those buffers untouched; up to three passes
    with two weight sets per launch; more samples than CTAs (the buffer is refilled slot by slot behind conv-4)."""
    monkeypatch.setenv('CURLA_CONV_N96', '1')
    monkeypatch.setenv('CURLA_CONV_N64', '0')
    torch.manual_seed(11)
    npass = len(Bs)
    g0 = Geom(H, W, max(Bs))
    Hv = (C.c_int * 3)(*g0.Ho[1:4])
    Wv = (C.c_int * 3)(*g0.Wo[1:4])
    monkeypatch.setenv('CURLA_CONV_FUSED', '1')
    assert _lib.load().curla_conv_stack_fits(g0.pitch, g0.S, Hv, Wv) == 1
    wsets = [[torch.randn(32, 32, 3, 3) * (2.0 / 288) ** 0.5 for _ in range(3)] for _ in range(2)]
    bsets = [[(torch.randn(32) * 0.05).to(DEV) for _ in range(3)] for _ in range(2)]
    # packed N = 96 images of both weight sets through the product's packer
    w96 = []
    for k in range(2):
        src = torch.cat([w.reshape(-1) for w in wsets[k]]).to(DEV)
        dst = torch.zeros(3 * 9216, dtype=torch.bfloat16, device=DEV)
        segs = torch.tensor([[l * 9216, l * 9216, 3, 32, 32, 32, 32] for l in range(3)], dtype=torch.int64)
        _lib.call('curla_pack_shadows', _lib.ptr(src), _lib.ptr(dst), C.c_void_p(segs.data_ptr()), 3, stream())
        w96.append(dst)
    keep_alive, ins, refs = [], [], []
    for k, B in enumerate(Bs):
        g = Geom(H, W, B)
        full, a0 = g.to_pitch(torch.rand(B, 32, g.Ho[0], g.Wo[0], device=DEV), 0)
        keep_alive.append(full)
        ins.append((g, a0))
        # reference: three launches of the per-layer kernel
        cur, outs = a0, []
        for l in range(3):
            wsh = pack_conv_w(wsets[k % 2][l], False)
            fo, o = g.alloc(32)
            _lib.call('curla_conv_fwd', _lib.ptr(cur), g.S * 32, _lib.ptr(wsh), _lib.ptr(bsets[k % 2][l]), 1.0, _lib.ptr(o),
                      g.S * 32, B, g.pitch, g.S, g.Ho[l + 1], g.Wo[l + 1], 0, stream())
            keep_alive += [fo, wsh]
            outs.append(o)
            cur = o
        refs.append(outs)
    segs = (_lib.ConvStackSeg * 3)()
    fused = []
    for k, B in enumerate(Bs):
        g, a0 = ins[k]
        keep = (k % 2 == 0)                                        # passes 0 and 2 keep conv-2 / conv-3
        outs = []
        for l in range(3):
            fo, o = g.alloc(32)
            o.fill_(7.0)
            keep_alive.append(fo)
            outs.append(o)
        fused.append((keep, outs))
        segs[k].inp = a0.data_ptr(); segs[k].w96 = w96[k % 2].data_ptr(); segs[k].B = B
        for l in range(3):
            segs[k].bias[l] = bsets[k % 2][l].data_ptr()
            segs[k].out[l] = outs[l].data_ptr() if (keep or l == 2) else None
    _lib.call('curla_conv_stack_fwd', C.byref(segs), npass, g0.S * 32, g0.pitch, g0.S, Hv, Wv, stream())
    torch.cuda.synchronize()
    for k, B in enumerate(Bs):
        g, _ = ins[k]
        keep, outs = fused[k]
        for l in range(3):
            n = g.Ho[l + 1] * g.pitch
            a = outs[l].view(torch.int16).view(B, 4, g.S, 8)
            b = refs[k][l].view(torch.int16).view(B, 4, g.S, 8)
            if keep or l == 2:
                assert torch.equal(a[:, :, :n], b[:, :, :n]), (k, l, int((a[:, :, :n] != b[:, :, :n]).sum()))
            else:
                assert bool((outs[l] == 7.0).all()), (k, l)


def _conv_mode(monkeypatch, mode):
    """layers 2..4: '32' = one tap per MMA (default), '64' = two horizontal taps per MMA + the third by K accumulation
    (one shifted add in the epilogue), '96' = three horizontal taps per MMA (two shifted adds)."""
    monkeypatch.setenv('CURLA_CONV_N96', '1' if mode == '96' else '0')
    monkeypatch.setenv('CURLA_CONV_N64', '1' if mode == '64' else '0')


@pytest.mark.parametrize('n96', ['96', '64', '32'])
@pytest.mark.parametrize('H,W,B', [(76, 135, 3), (90, 160, 2), (64, 64, 5)])
def test_conv_forward(H, W, B, n96, monkeypatch):
    """The three formulations of the 32 -> 32 channel layers (see _conv_mode): same tolerances for all."""
    _conv_mode(monkeypatch, n96)
    g, x, ws, bs = _conv_case(H, W, B)
    s2d, acts, keep = _run_conv_stack(g, x, ws, bs)
    # reference on bf16-rounded operands, layer by layer (our activations are stored bf16)
    cur = x.to(DEV) / 255.0
    cur_exact = x.to(DEV) / 255.0
    for l in range(4):
        w, b = ws[l].to(DEV), bs[l].to(DEV)
        st = 2 if l == 0 else 1
        if l == 0:   # ours: integer input exact, weights bf16, scale applied in fp32
            ref = torch.relu(F.conv2d(x.to(DEV), bf16r(w), None, stride=2) / 255.0 + b.view(1, -1, 1, 1))
        else:
            ref = torch.relu(F.conv2d(cur, bf16r(w), b, stride=st))
        cur_exact = torch.relu(F.conv2d(cur_exact, w, b, stride=st))
        ours = g.from_pitch(acts[l], g.Ho[l], g.Wo[l])
        assert ref.shape == ours.shape
        e = rel_l2(ours, ref)
        assert e < 6e-3, (l, e)          # bf16 storage rounding of the output + accumulation order
        assert rel_l2(ours, cur_exact) < 3e-2, l
        # invalid positions are exact zeros
        v = g.nhwc(acts[l]).float()
        assert float(v[:, g.Ho[l]:].abs().sum()) == 0 and float(v[:, :, g.Wo[l]:].abs().sum()) == 0
        cur = bf16r(ours)                 # feed OUR stored activations to the next reference layer


def test_conv_forward_dynamic_tiles_match_static(monkeypatch):
    """More than two tiles per SM: tiles are handed out by an atomic counter.  Same bits as the
    round-robin split, launch after launch (the counter pair re-arms itself)."""
    g, x, ws, bs = _conv_case(76, 135, 48, seed=5)
    assert 48 * -(-g.Ho[0] * g.pitch // 256) > 2 * torch.cuda.get_device_properties(0).multi_processor_count
    monkeypatch.setenv('CURLA_TC_STATIC', '1')
    _, ref, keep0 = _run_conv_stack(g, x, ws, bs)
    torch.cuda.synchronize()
    monkeypatch.setenv('CURLA_TC_STATIC', '0')
    for _ in range(3):
        _, acts, keep1 = _run_conv_stack(g, x, ws, bs)
        torch.cuda.synchronize()
        for l in range(4):
            assert torch.equal(acts[l].view(torch.int16), ref[l].view(torch.int16)), l


@pytest.mark.parametrize('H,W', [(76, 135), (90, 160)])
def test_conv_forward_multi_segment(H, W):
    """One launch over three passes (two weight sets, different batch sizes) == three launches, bit for bit."""
    import ctypes as C
    Bs = [3, 2, 4]
    cases = [_conv_case(H, W, b, seed=10 + i) for i, b in enumerate(Bs)]
    wsets = [cases[0][2:], cases[1][2:], cases[0][2:]]          # passes 0 and 2 share weights (F1 / F3)
    singles, s2ds, keep = [], [], []
    for (g, x, _, _), (ws, bs) in zip(cases, wsets):
        s2d, acts, kp = _run_conv_stack(g, x, ws, bs)
        singles.append(acts); s2ds.append(s2d); keep.append(kp)
    wsh = [[pack_conv_w(ws[l], l == 0) for l in range(4)] for ws, _ in wsets[:2]]
    bd = [[bs[l].to(DEV) for l in range(4)] for _, bs in wsets[:2]]
    wsel = [0, 1, 0]
    g0 = cases[0][0]
    ins = s2ds
    for l in range(4):
        outs = []
        segs = (_lib.ConvSeg * 3)()
        for k in range(3):
            full, out = cases[k][0].alloc(32)
            out.fill_(-7.0)                                        # every position must be (re)written
            keep.append(full); outs.append(out)
            segs[k].inp = ins[k].data_ptr(); segs[k].wts = wsh[wsel[k]][l].data_ptr()
            segs[k].bias = bd[wsel[k]][l].data_ptr(); segs[k].out = out.data_ptr(); segs[k].B = Bs[k]
        _lib.call('curla_conv_fwd_multi', C.byref(segs), 3, g0.S * (g0.CP1 if l == 0 else 32),
                  1.0 / 255.0 if l == 0 else 1.0, g0.S * 32, g0.pitch, g0.S, g0.Ho[l], g0.Wo[l], 36 if l == 0 else 0,
                  stream())                                        # 36 real s2d channels: the all-zero plane is not read
        torch.cuda.synchronize()
        tile_out = 254 if (l > 0 and g0.Wo[l] <= g0.pitch - 2) else 256   # N = 96 kernel: tiles overlap by two rows
        n = min(g0.S, -(-g0.Ho[l] * g0.pitch // tile_out) * tile_out)   # positions a launch writes per sample and plane
        for k in range(3):
            a = outs[k].view(torch.int16).view(Bs[k], 4, g0.S, 8)
            b = singles[k][l].view(torch.int16).view(Bs[k], 4, g0.S, 8)
            assert torch.equal(a[:, :, :n], b[:, :, :n]), (l, k)
            a[:, :, n:] = 0                                        # never written: zero like the engine's arenas
        ins = outs
    # more than two distinct weight sets is refused
    segs = (_lib.ConvSeg * 3)()
    for k in range(3):
        segs[k].inp = s2ds[k].data_ptr(); segs[k].wts = wsh[k % 2][0].data_ptr(); segs[k].bias = bd[k % 2][0].data_ptr() + 4 * (k == 2)
        segs[k].out = outs[k].data_ptr(); segs[k].B = Bs[k]
    with pytest.raises(_lib.CurlaError):
        _lib.call('curla_conv_fwd_multi', C.byref(segs), 3, g0.S * g0.CP1, 1.0, g0.S * 32, g0.pitch, g0.S, g0.Ho[0],
                  g0.Wo[0], 1, stream())


@pytest.mark.parametrize('n96', ['96', '64', '32'])
@pytest.mark.parametrize('H,W,B', [(76, 135, 3), (90, 160, 2), (76, 135, 40)])
def test_conv_backward(H, W, B, n96, monkeypatch):
    _conv_mode(monkeypatch, n96)
    g, x, ws, bs = _conv_case(H, W, B, seed=1)
    s2d, acts, keep = _run_conv_stack(g, x, ws, bs)
    torch.manual_seed(2)
    # upstream gradient on the last layer's valid region
    dy4 = torch.randn(B, 32, g.Ho[3], g.Wo[3], device=DEV) * (g.from_pitch(acts[3], g.Ho[3], g.Wo[3]) > 0)
    # torch reference: autograd through convs evaluated on OUR stored activations
    wref = [bf16r(w.to(DEV)).requires_grad_(True) for w in ws]
    bref = [b.to(DEV).clone().requires_grad_(True) for b in bs]
    a_in = [x.to(DEV) / 255.0] + [bf16r(g.from_pitch(acts[l], g.Ho[l], g.Wo[l])) for l in range(3)]
    dfull, dview = g.to_pitch(bf16r(dy4), 3)
    dcur = dview
    ws_buf = torch.zeros(int(max(_lib.load().curla_conv_wgrad_workspace_floats(0),
                                 _lib.load().curla_conv_wgrad_workspace_floats(1))), device=DEV)
    dy_ref = bf16r(dy4)
    for l in (3, 2, 1, 0):
        # reference grads for layer l given dy_ref on its (post-ReLU-masked) output
        xin = a_in[l].detach().requires_grad_(l > 0)
        y = F.conv2d(xin, wref[l], bref[l], stride=2 if l == 0 else 1)
        grads = torch.autograd.grad(y, ([xin] if l > 0 else []) + [wref[l], bref[l]], dy_ref)
        dW = torch.zeros(ws[l].shape, device=DEV)
        db = torch.zeros(32, device=DEV)
        inp = acts[l - 1] if l > 0 else s2d
        _lib.call('curla_conv_wgrad', _lib.ptr(inp), g.S * (32 if l > 0 else g.CP1), _lib.ptr(dcur), g.S * 32,
                  _lib.ptr(ws_buf), _lib.ptr(dW), _lib.ptr(db), 1.0 / 255.0 if l == 0 else 1.0, B, g.pitch, g.S,
                  g.Ho[l], g.Wo[l], ws[l].shape[1], 1 if l == 0 else 0, stream())
        assert rel_l2(dW, grads[-2]) < 5e-3, ('dW', l, rel_l2(dW, grads[-2]))
        assert rel_l2(db, grads[-1]) < 5e-3, ('db', l)
        if l > 0:
            full, dx = g.alloc(32)
            wsh = pack_conv_w(ws[l], False)
            keep.append(wsh)
            _lib.call('curla_conv_dgrad', _lib.ptr(dcur), g.S * 32, _lib.ptr(wsh),
                      _lib.ptr(acts[l - 1]), _lib.ptr(dx), g.S * 32, B, g.pitch, g.S, g.Ho[l - 1], g.Wo[l - 1],
                      stream())
            mask = (g.from_pitch(acts[l - 1], g.Ho[l - 1], g.Wo[l - 1]) > 0)
            dx_ref = grads[0] * mask
            ours = g.from_pitch(dx, g.Ho[l - 1], g.Wo[l - 1])
            assert rel_l2(ours, dx_ref) < 6e-3, ('dx', l, rel_l2(ours, dx_ref))
            v = g.nhwc(dx).float()
            assert float(v[:, g.Ho[l - 1]:].abs().sum()) == 0 and float(v[:, :, g.Wo[l - 1]:].abs().sum()) == 0
            keep.append(full)
            dcur = dx
            dy_ref = bf16r(ours)


@pytest.mark.parametrize('H,W,B', [(76, 135, 5), (90, 160, 3)])
def test_conv_wgrad_staging_modes_agree(H, W, B, monkeypatch):
    """conv_wgrad_tc.cu can stage its operands with tensor-map TMA boxes (CURLA_WG_TMAP=1: all channel planes of an
    image row per request, the shifted dy copies zero-filled by the TMA unit) instead of one linear bulk copy per plane
    per row, and can fetch dy from L2 once, making the shifted copies inside shared memory (CURLA_WG_DY1=1).  Same products, fp32 accumulation: the weight / bias gradients agree to 1e-5, conv-1 (2x2 taps on the
    space-to-depth input, 6 planes) and a 3x3 layer."""
    g, x, ws, bs = _conv_case(H, W, B, seed=4)
    s2d, acts, keep = _run_conv_stack(g, x, ws, bs)
    torch.manual_seed(5)
    ws_buf = torch.zeros(int(max(_lib.load().curla_conv_wgrad_workspace_floats(0),
                                 _lib.load().curla_conv_wgrad_workspace_floats(1))), device=DEV)
    for l in (0, 2):
        dy = torch.randn(B, 32, g.Ho[l], g.Wo[l], device=DEV)
        dfull, dview = g.to_pitch(bf16r(dy), l)
        inp = acts[l - 1] if l > 0 else s2d
        got = {}
        for mode in ('00', '10', '01', '11'):          # (tensor-map staging, dy staged once + replicated in shared memory)
            monkeypatch.setenv('CURLA_WG_TMAP', mode[0])
            monkeypatch.setenv('CURLA_WG_DY1', mode[1])
            dW = torch.zeros(ws[l].shape, device=DEV)
            db = torch.zeros(32, device=DEV)
            _lib.call('curla_conv_wgrad', _lib.ptr(inp), g.S * (32 if l > 0 else g.CP1), _lib.ptr(dview), g.S * 32,
                      _lib.ptr(ws_buf), _lib.ptr(dW), _lib.ptr(db), 1.0, B, g.pitch, g.S,
                      g.Ho[l], g.Wo[l], ws[l].shape[1], 1 if l == 0 else 0, stream())
            torch.cuda.synchronize()
            got[mode] = (dW, db)
        # (the modes may split the batch into different runs per CTA: same products, another fp32 summation order)
        for mode in ('10', '01', '11'):
            assert rel_l2(got[mode][0], got['00'][0]) < 1e-5, (l, mode, rel_l2(got[mode][0], got['00'][0]))
            assert rel_l2(got[mode][1], got['00'][1]) < 1e-5, (l, mode)
        assert torch.equal(got['01'][0], got['00'][0]) and torch.equal(got['01'][1], got['00'][1])      # same runs, same MMAs
        assert float(got['11'][0].abs().sum()) > 0


@pytest.mark.parametrize('H,W,B', [(76, 135, 5), (90, 160, 3), (76, 135, 1), (76, 135, 37)])
def test_conv_wgrad_variants_agree(H, W, B, monkeypatch):
    """conv_wgrad_tc.cu: the horizontal taps ride in N (one shifted copy of dy per tap, CURLA_WG_COPIES=3, the default)
    or on the A side (the input's K window starts 1, 2 rows later against ONE staged dy: CURLA_WG_COPIES=1; 2 = one
    N = 64 and one N = 32 MMA); the dy copies are issued by the producer warp or by up to four helper warps
    (CURLA_WG_PRODUCERS, default 5 threads issuing); ring depth 2..4; operands staged by linear bulk copies or
    tensor-map boxes (CURLA_WG_TMAP).  Same products, fp32 accumulation: every variant agrees with the torch gradient
    and with the others to 1e-5, for conv-1 (2x2 taps) and 3x3 layers; B = 1 leaves most CTAs without work."""
    g, x, ws, bs = _conv_case(H, W, B, seed=6)
    s2d, acts, keep = _run_conv_stack(g, x, ws, bs)
    torch.manual_seed(7)
    ws_buf = torch.zeros(int(max(_lib.load().curla_conv_wgrad_workspace_floats(0),
                                 _lib.load().curla_conv_wgrad_workspace_floats(1))), device=DEV)
    # (copies, stages, producers, tensor-map staging)
    variants = [('3', '2', '1', '0'), ('3', '2', '5', '0'), ('2', '2', '1', '0'), ('1', '3', '1', '0'), ('3', '2', '3', '0'),
                ('1', '4', '4', '0'), ('3', '3', '1', '1'), ('2', '2', '5', '0'), ('1', '2', '2', '1')]
    for l in (0, 1, 3):
        dy = torch.randn(B, 32, g.Ho[l], g.Wo[l], device=DEV)
        dfull, dview = g.to_pitch(bf16r(dy), l)
        inp = acts[l - 1] if l > 0 else s2d
        xin = (x.to(DEV) / 255.0) if l == 0 else bf16r(g.from_pitch(acts[l - 1], g.Ho[l - 1], g.Wo[l - 1]))
        wref = bf16r(ws[l].to(DEV)).requires_grad_(True)
        bref = bs[l].to(DEV).clone().requires_grad_(True)
        y = F.conv2d(xin, wref, bref, stride=2 if l == 0 else 1)
        gw, gb = torch.autograd.grad(y, [wref, bref], bf16r(dy))
        got = {}
        for v in variants:
            copies, depth, prod, tmap = v
            monkeypatch.setenv('CURLA_WG_COPIES', copies)
            monkeypatch.setenv('CURLA_WG_STAGES', depth)
            monkeypatch.setenv('CURLA_WG_PRODUCERS', prod)
            monkeypatch.setenv('CURLA_WG_TMAP', tmap)
            dW = torch.full(ws[l].shape, float('nan'), device=DEV)
            db = torch.full((32,), float('nan'), device=DEV)
            _lib.call('curla_conv_wgrad', _lib.ptr(inp), g.S * (32 if l > 0 else g.CP1), _lib.ptr(dview), g.S * 32,
                      _lib.ptr(ws_buf), _lib.ptr(dW), _lib.ptr(db), 1.0 / 255.0 if l == 0 else 1.0, B, g.pitch, g.S,
                      g.Ho[l], g.Wo[l], ws[l].shape[1], 1 if l == 0 else 0, stream())
            torch.cuda.synchronize()
            got[v] = (dW, db)
            assert rel_l2(dW, gw) < 5e-3, ('dW', l, v, rel_l2(dW, gw))
            assert rel_l2(db, gb) < 5e-3, ('db', l, v)
        ref = got[variants[0]]
        for k, v in got.items():
            assert rel_l2(v[0], ref[0]) < 1e-5 and rel_l2(v[1], ref[1]) < 1e-5, (l, k)


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize('layout', [3, 1, 2, 0])
@pytest.mark.parametrize('M,N,K', [(70, 72, 96), (5, 64, 64), (130, 200, 40), (64, 64, 2048), (128, 64, 64), (300, 136, 200),
                                   (512, 1024, 1024), (50, 4104, 136)])
def test_gemm_layouts(layout, M, N, K):
    torch.manual_seed(layout * 100 + M)
    A = torch.randn(M, K, device=DEV)
    Bm = torch.randn(K, N, device=DEV)
    ref = bf16r(A) @ bf16r(Bm)
    # contiguous extents must be multiples of 8: pad the leading dims
    r8 = lambda n: (n + 7) // 8 * 8
    if layout & 1:
        a = torch.zeros(M, r8(K), device=DEV, dtype=torch.bfloat16); a[:, :K] = A
    else:
        a = torch.zeros(K, r8(M), device=DEV, dtype=torch.bfloat16); a[:, :M] = A.t()
    if layout & 2:
        b = torch.zeros(N, r8(K), device=DEV, dtype=torch.bfloat16); b[:, :K] = Bm.t()
    else:
        b = torch.zeros(K, r8(N), device=DEV, dtype=torch.bfloat16); b[:, :N] = Bm
    out = torch.full((M, N + (N & 1)), 7.0, device=DEV)
    _lib.call('curla_gemm_bf16', _lib.ptr(a), a.shape[1], _lib.ptr(b), b.shape[1], _lib.ptr(out), out.shape[1],
              M, N, K, layout, N, 0, None, 0, None, 0, 1, 0, 1.0, stream())
    assert rel_l2(out[:, :N], ref) < 2e-3, rel_l2(out[:, :N], ref)


@pytest.mark.parametrize('case', ['mlp_fwd', 'mlp_dgrad', 'mlp_wgrad', 'mlp_wgrad_in', 'fc_fwd', 'fc_dgrad', 'fc_wgrad'])
def test_gemm_tcgen05_matches_mma_sync(case, monkeypatch):
    """The engine's GEMM call patterns at a batch that takes the tcgen05 kernel (gemm_tc.cu): same
    call with CURLA_GEMM_TC=0 (mma.sync kernel) and =1, and against torch."""
    torch.manual_seed(hash(case) % 1000)
    Bn, hid = 200, 256
    rnd = lambda *sh: torch.randn(*sh, device=DEV).to(torch.bfloat16)
    seg_len, seg_stride, nseg = 1032, 1048, 4          # Kfc >= 4096: the few-row fc wgrad also takes the tcgen05 kernel
    Kfc = seg_len * nseg

    def scatter(x):
        st = torch.full((x.shape[0], nseg * seg_stride), 3.0, device=DEV, dtype=x.dtype)
        for s_ in range(nseg):
            st[:, s_ * seg_stride:s_ * seg_stride + seg_len] = x[:, s_ * seg_len:(s_ + 1) * seg_len]
        return st

    def run(tc):
        monkeypatch.setenv('CURLA_GEMM_TC', '2' if tc else '0')        # 2 = every supported shape
        if case == 'mlp_fwd':          # H2 = relu(H1 . W^T + b), Q1 || Q2 batched, bf16 out
            H1, W, b = rnd(2, Bn, hid), rnd(2, hid, hid) * 0.1, torch.randn(2, hid, device=DEV)
            out = torch.zeros(2, Bn, hid, device=DEV, dtype=torch.bfloat16)
            _lib.call('curla_gemm_bf16_batched', _lib.ptr(H1), hid, _lib.ptr(W), hid, _lib.ptr(out), hid, Bn, hid, hid, 3, hid, 1,
                      _lib.ptr(b), 1, None, 0, 1.0, 2, Bn * hid, hid * hid, Bn * hid, hid, 0, stream())
            ref = torch.relu(torch.einsum('bmk,bnk->bmn', H1.float(), W.float()) + b[:, None, :])
            return out.float(), ref, (H1, W, b)
        if case == 'mlp_dgrad':        # dH1 = (H1 > 0) * dH2 . W, batched, bf16 out
            dH2, W, H1 = rnd(2, Bn, hid), rnd(2, hid, hid) * 0.1, rnd(2, Bn, hid)
            out = torch.zeros(2, Bn, hid, device=DEV, dtype=torch.bfloat16)
            _lib.call('curla_gemm_bf16_batched', _lib.ptr(dH2), hid, _lib.ptr(W), hid, _lib.ptr(out), hid, Bn, hid, hid, 1, hid, 1,
                      None, 0, _lib.ptr(H1), hid, 1.0, 2, Bn * hid, hid * hid, Bn * hid, 0, Bn * hid, stream())
            ref = torch.einsum('bmk,bkn->bmn', dH2.float(), W.float()) * (H1.float() > 0)
            return out.float(), ref, (dH2, W, H1)
        if case == 'mlp_wgrad':        # dW[hid][hid] = dH2^T . H1, batched, fp32 out
            dH2, H1 = rnd(2, Bn, hid), rnd(2, Bn, hid)
            out = torch.zeros(2, hid, hid, device=DEV)
            _lib.call('curla_gemm_bf16_batched', _lib.ptr(dH2), hid, _lib.ptr(H1), hid, _lib.ptr(out), hid, hid, hid, Bn, 0, hid, 0,
                      None, 0, None, 0, 1.0, 2, Bn * hid, Bn * hid, hid * hid, 0, 0, stream())
            return out, torch.einsum('bkm,bkn->bmn', dH2.float(), H1.float()), (dH2, H1)
        if case == 'mlp_wgrad_in':     # dW0[hid][52] = dH1^T . X (X stored [B][64]), n_store 52, ldc 52
            dH1, X = rnd(Bn, hid), rnd(Bn, 64)
            out = torch.full((hid, 52), 9.0, device=DEV)
            _lib.call('curla_gemm_bf16_batched', _lib.ptr(dH1), hid, _lib.ptr(X), 64, _lib.ptr(out), 52, hid, 64, Bn, 0, 52, 0,
                      None, 0, None, 0, 1.0, 1, 0, 0, 0, 0, 0, stream())
            return out, (dH1.float().t() @ X.float())[:, :52], (dH1, X)
        act = rnd(Bn, Kfc)
        act_s = scatter(act)
        lds = act_s.shape[1]
        W = rnd(64, Kfc) * 0.1
        W[48:] = 0
        if case == 'fc_fwd':           # split-K partials of z = act . W^T, A segmented
            sp = _lib.load().curla_gemm_effective_splits(Kfc, 5)
            part = torch.zeros(sp, Bn, 64, device=DEV)
            _lib.call('curla_gemm_bf16_seg', _lib.ptr(act_s), lds, _lib.ptr(W), Kfc, _lib.ptr(part), 64, Bn, 64, Kfc, 3, 64, 0, None, 0,
                      None, 0, sp, Bn * 64, 1.0, seg_len, seg_stride, 1, stream())
            return part.sum(0), act.float() @ W.float().t(), (act_s, W)
        dz = rnd(Bn, 64)
        if case == 'fc_dgrad':         # dact = (act > 0) * dz . W, C and mask segmented, bf16 out
            out = torch.full((Bn, lds), 5.0, device=DEV, dtype=torch.bfloat16)
            _lib.call('curla_gemm_bf16_seg', _lib.ptr(dz), 64, _lib.ptr(W), Kfc, _lib.ptr(out), lds, Bn, Kfc, 64, 1, Kfc, 1, None, 0,
                      _lib.ptr(act_s), lds, 1, 0, 1.0, seg_len, seg_stride, 4, stream())
            want = scatter(((dz.float() @ W.float()) * (act.float() > 0)).to(torch.bfloat16)).float()
            want[:, [c for c in range(lds) if c % seg_stride >= seg_len]] = 5.0      # gaps untouched
            return out.float(), want, (dz, W, act_s)
        # fc_wgrad: dW[48 (64)][Kfc] = dz^T . act, B segmented, fp32 out, M = 48 rows
        out = torch.zeros(64, Kfc, device=DEV)
        _lib.call('curla_gemm_bf16_seg', _lib.ptr(dz), 64, _lib.ptr(act_s), lds, _lib.ptr(out), Kfc, 48, Kfc, Bn, 0, Kfc, 0, None, 0,
                  None, 0, 1, 0, 1.0, seg_len, seg_stride, 2, stream())
        return out[:48], (dz.float().t() @ act.float())[:48], (dz, act_s)

    st = torch.get_rng_state(), torch.cuda.get_rng_state()
    a, ref, keep_a = run(False)
    torch.set_rng_state(st[0]); torch.cuda.set_rng_state(st[1])
    b, ref_b, keep_b = run(True)
    torch.cuda.synchronize()
    assert torch.equal(ref, ref_b)
    assert rel_l2(a, ref) < 6e-3 and rel_l2(b, ref) < 6e-3, (rel_l2(a, ref), rel_l2(b, ref))
    assert rel_l2(b, a) < 3e-3, rel_l2(b, a)


def test_gemm_epilogues_and_splitk():
    torch.manual_seed(5)
    M, N, K = 37, 64, 4096
    A = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    W = torch.randn(N, K, device=DEV).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    ref = A.float() @ W.float().t()
    # split-K partials
    splits = _lib.load().curla_gemm_effective_splits(K, 7)
    part = torch.zeros((splits, M, N), device=DEV)
    _lib.call('curla_gemm_bf16', _lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(part), N, M, N, K, 3, N, 0, None, 0,
              None, 0, splits, M * N, 1.0, stream())
    assert rel_l2(part.sum(0), ref) < 2e-3
    # bias + relu + bf16 out + mask + n_store
    mask = (torch.randn(M, N, device=DEV) > 0).to(torch.bfloat16)
    out = torch.zeros((M, N), device=DEV, dtype=torch.bfloat16)
    _lib.call('curla_gemm_bf16', _lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(out), N, M, N, K, 3, 50, 1, _lib.ptr(bias),
              1, _lib.ptr(mask), N, 1, 0, 0.5, stream())
    want = torch.relu(ref * 0.5 + bias) * mask.float()
    assert rel_l2(out[:, :50].float(), want[:, :50]) < 6e-3
    assert float(out[:, 50:].abs().sum()) == 0


def test_gemm_segmented_operands():
    """the channel-plane activation layout: the contiguous index of A (fc fwd), B (fc wgrad) or
    C + mask (fc dgrad) is 4 segments of seg_len elements, seg_stride apart."""
    torch.manual_seed(11)
    Bn, feat, seg_len, seg_stride, nseg = 9, 48, 40, 56, 4
    K = seg_len * nseg
    act = torch.randn(Bn, K, device=DEV).to(torch.bfloat16)                 # logical [B][K]
    W = torch.randn(64, K, device=DEV).to(torch.bfloat16)
    W[feat:] = 0

    def scatter(x):   # logical [rows][K] -> segmented storage [rows][nseg*seg_stride] (junk in the gaps)
        st = torch.full((x.shape[0], nseg * seg_stride), 3.0, device=DEV, dtype=x.dtype)
        for s_ in range(nseg):
            st[:, s_ * seg_stride:s_ * seg_stride + seg_len] = x[:, s_ * seg_len:(s_ + 1) * seg_len]
        return st

    act_s = scatter(act)
    lds = act_s.shape[1]
    # fwd: z[B][64] = act . W^T, A segmented (mask 1)
    z = torch.zeros(Bn, 64, device=DEV)
    _lib.call('curla_gemm_bf16_seg', _lib.ptr(act_s), lds, _lib.ptr(W), K, _lib.ptr(z), 64, Bn, 64, K, 3, 64, 0, None, 0,
              None, 0, 1, 0, 1.0, seg_len, seg_stride, 1, stream())
    assert rel_l2(z, act.float() @ W.float().t()) < 2e-3
    # wgrad: dW[64][K] = dz^T . act, B segmented (mask 2); dz stored [B][64] = A MN-major
    dz = torch.randn(Bn, 64, device=DEV).to(torch.bfloat16)
    dW = torch.zeros(64, K, device=DEV)
    _lib.call('curla_gemm_bf16_seg', _lib.ptr(dz), 64, _lib.ptr(act_s), lds, _lib.ptr(dW), K, 64, K, Bn, 0, K, 0, None, 0,
              None, 0, 1, 0, 1.0, seg_len, seg_stride, 2, stream())
    assert rel_l2(dW, dz.float().t() @ act.float()) < 2e-3
    # dgrad: dact[B][K] = (act > 0) * dz . W, C and mask segmented (mask 4), bf16 out
    dact_s = torch.full((Bn, lds), 5.0, device=DEV, dtype=torch.bfloat16)
    _lib.call('curla_gemm_bf16_seg', _lib.ptr(dz), 64, _lib.ptr(W), K, _lib.ptr(dact_s), lds, Bn, K, 64, 1, K, 1, None, 0,
              _lib.ptr(act_s), lds, 1, 0, 1.0, seg_len, seg_stride, 4, stream())
    want = scatter(((dz.float() @ W.float()) * (act.float() > 0)).to(torch.bfloat16))
    gaps = torch.ones(lds, dtype=torch.bool, device=DEV)
    for s_ in range(nseg):
        gaps[s_ * seg_stride:s_ * seg_stride + seg_len] = False
    assert rel_l2(dact_s[:, ~gaps].float(), want[:, ~gaps].float()) < 6e-3
    assert bool((dact_s[:, gaps] == 5.0).all())          # gaps between segments untouched


# ------------------------------------------------------------------ LayerNorm / heads / policy / losses
def test_layernorm_fwd_bwd():
    torch.manual_seed(3)
    B, feat = 19, 50
    part = torch.zeros((3, B, 64), device=DEV); part[:, :, :feat] = torch.randn(3, B, feat, device=DEV)
    bias, gamma, beta = [torch.randn(feat, device=DEV) for _ in range(3)]
    x_out = torch.zeros((B, 64), device=DEV); z = torch.zeros((B, 64), device=DEV)
    _lib.call('curla_ln_fwd', _lib.ptr(part), 3, B * 64, _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta), B, feat,
              0, _lib.ptr(x_out), _lib.ptr(z), stream())
    x = (part.sum(0)[:, :feat] + bias).requires_grad_(True)
    g = gamma.clone().requires_grad_(True); bt = beta.clone().requires_grad_(True)
    zr = F.layer_norm(x, (feat,), g, bt, 1e-5)
    assert max_rel(z[:, :feat], zr) < 1e-5 and float(z[:, feat:].abs().sum()) == 0
    dz = torch.zeros((B, 64), device=DEV); dz[:, :feat] = torch.randn(B, feat, device=DEV)
    dz2 = torch.zeros((B, 64), device=DEV); dz2[:, :feat] = torch.randn(B, feat, device=DEV)
    zr.backward(dz[:, :feat] + dz2[:, :feat])
    dx = torch.zeros((B, 64), device=DEV); dxb = torch.zeros((B, 64), device=DEV, dtype=torch.bfloat16)
    scratch = torch.zeros((2, B, 64), device=DEV)
    dg, dbt, dbias = [torch.zeros(feat, device=DEV) for _ in range(3)]
    _lib.call('curla_ln_bwd', _lib.ptr(dz), _lib.ptr(dz2), _lib.ptr(x_out), _lib.ptr(gamma), B, feat, _lib.ptr(dx),
              _lib.ptr(dxb), _lib.ptr(scratch), _lib.ptr(dg), _lib.ptr(dbt), _lib.ptr(dbias), stream())
    assert max_rel(dx[:, :feat], x.grad) < 1e-4
    assert max_rel(dg, g.grad) < 1e-4 and max_rel(dbt, bt.grad) < 1e-4
    assert max_rel(dbias, x.grad.sum(0)) < 1e-4
    assert rel_l2(dxb.float(), dx) < 4e-3


def test_policy_head_fwd_bwd():
    torch.manual_seed(4)
    B, A = 21, 2
    t = (torch.randn(B, 2 * A, device=DEV) * 1.5).requires_grad_(True)
    noise = torch.randn(B, A, device=DEV)
    mu, ls_raw = t.chunk(2, dim=-1)
    ls = torch.tanh(ls_raw); ls = -10 + 0.5 * 12 * (ls + 1)
    pi_pre = mu + noise * ls.exp()
    lp = O.gaussian_logprob(noise, ls)
    pi = torch.tanh(pi_pre)
    lp = lp - torch.log(F.relu(1 - pi.pow(2)) + 1e-6).sum(-1, keepdim=True)
    o = [torch.zeros((B, A), device=DEV) for _ in range(2)] + [torch.zeros(B, device=DEV)] + \
        [torch.zeros((B, A), device=DEV) for _ in range(2)]
    _lib.call('curla_policy_fwd', _lib.ptr(t.detach()), _lib.ptr(noise), 0, 0, B, A, -10.0, 2.0, 1, 1, _lib.ptr(o[0]),
              _lib.ptr(o[1]), _lib.ptr(o[2]), _lib.ptr(o[3]), _lib.ptr(o[4]), stream())
    assert max_rel(o[0], torch.tanh(mu)) < 1e-5 and max_rel(o[1], pi) < 1e-5
    assert (o[2] - lp[:, 0]).abs().max() < 2e-3 * max(1.0, float(lp.abs().max()))
    assert max_rel(o[3], ls) < 1e-5 and torch.equal(o[4], noise)
    # backward: L = sum(gpi * pi) + glp * sum(log_pi)
    gpi = torch.randn(B, A, device=DEV)
    glp = 0.37
    (gpi * pi).sum().add(glp * lp.sum()).backward()
    dx1 = torch.zeros((B, 64), device=DEV); dx1[:, 50:52] = gpi * 0.25
    dx2 = torch.zeros((B, 64), device=DEV); dx2[:, 50:52] = gpi * 0.75
    dt = torch.zeros((B, 2 * A), device=DEV)
    glp_d = torch.tensor([glp], device=DEV)
    _lib.call('curla_policy_bwd', _lib.ptr(dx1), _lib.ptr(dx2), 50, _lib.ptr(glp_d),
              _lib.ptr(t.detach()), _lib.ptr(noise), _lib.ptr(o[1]), _lib.ptr(o[3]), B, A, -10.0, 2.0, _lib.ptr(dt),
              stream())
    assert rel_l2(dt, t.grad) < 1e-3, rel_l2(dt, t.grad)
    # built-in Philox noise: N(0,1) moments
    Bn = 1 << 16
    tt = torch.zeros((Bn, 2 * A), device=DEV)
    nz = torch.zeros((Bn, A), device=DEV)
    pi_o = torch.zeros((Bn, A), device=DEV); ls_o = torch.zeros((Bn, A), device=DEV)
    _lib.call('curla_policy_fwd', _lib.ptr(tt), None, 1234, 5, Bn, A, -10.0, 2.0, 1, 0, None, _lib.ptr(pi_o), None,
              _lib.ptr(ls_o), _lib.ptr(nz), stream())
    assert abs(float(nz.mean())) < 0.02 and abs(float(nz.std()) - 1.0) < 0.02
    assert abs(float((nz[:, 0] * nz[:, 1]).mean())) < 0.02


def test_mlp_heads_and_losses():
    torch.manual_seed(6)
    B, hid, No = 13, 64, 4
    H = torch.randn(B, hid, device=DEV).clamp_min(0).to(torch.bfloat16)
    W = torch.randn(No, hid, device=DEV); b = torch.randn(No, device=DEV)
    out = torch.zeros((B, No), device=DEV)
    _lib.call('curla_head_fwd', _lib.ptr(H), hid, _lib.ptr(W), _lib.ptr(b), B, hid, No, _lib.ptr(out), stream())
    assert max_rel(out, H.float() @ W.t() + b) < 1e-5
    dO = torch.randn(B, No, device=DEV)
    dH = torch.zeros((B, hid), device=DEV, dtype=torch.bfloat16)
    _lib.call('curla_head_bwd', _lib.ptr(dO), _lib.ptr(W), _lib.ptr(H), B, hid, No, _lib.ptr(dH), stream())
    assert rel_l2(dH.float(), (dO @ W) * (H.float() > 0)) < 4e-3
    dW = torch.zeros((No, hid), device=DEV); db = torch.zeros(No, device=DEV)
    _lib.call('curla_head_wgrad', _lib.ptr(dO), _lib.ptr(H), B, hid, No, _lib.ptr(dW), _lib.ptr(db), stream())
    assert max_rel(dW, dO.t() @ H.float()) < 1e-5 and max_rel(db, dO.sum(0)) < 1e-5
    cs = torch.zeros(hid, device=DEV)
    _lib.call('curla_colsum_bf16', _lib.ptr(dH), B, hid, _lib.ptr(cs), stream())
    assert max_rel(cs, dH.float().sum(0)) < 1e-5
    # losses
    tq1, tq2, lpn, rew, q1, q2 = [torch.randn(B, device=DEV) for _ in range(6)]
    nd = (torch.rand(B, device=DEV) > 0.2).float()
    la = torch.tensor([np.log(0.1)], device=DEV, dtype=torch.float64)
    tq = torch.zeros(B, device=DEV); d1 = torch.zeros(B, device=DEV); d2 = torch.zeros(B, device=DEV)
    m = torch.zeros(16, device=DEV)
    _lib.call('curla_critic_loss', _lib.ptr(tq1), _lib.ptr(tq2), _lib.ptr(lpn), _lib.ptr(rew), _lib.ptr(nd),
              _lib.ptr(la), 0.99, _lib.ptr(q1), _lib.ptr(q2), B, 1.0 / B, _lib.ptr(tq), _lib.ptr(d1), _lib.ptr(d2),
              _lib.ptr(m), stream())
    tref = rew + nd * 0.99 * (torch.min(tq1, tq2) - 0.1 * lpn)
    assert max_rel(tq, tref) < 1e-5
    assert abs(float(m[1]) - float(F.mse_loss(q1, tref) + F.mse_loss(q2, tref))) < 1e-4
    assert max_rel(d1, 2 * (q1 - tref) / B) < 1e-5 and abs(float(m[0]) - float(rew.mean())) < 1e-5
    ls = torch.randn(B, 2, device=DEV); glp = torch.zeros(4, device=DEV)
    gla = torch.zeros(1, device=DEV, dtype=torch.float64)
    _lib.call('curla_actor_loss', _lib.ptr(lpn), _lib.ptr(q1), _lib.ptr(q2), _lib.ptr(ls), B, 2, _lib.ptr(la), -2.0,
              1.0 / B, _lib.ptr(d1), _lib.ptr(d2), _lib.ptr(glp), _lib.ptr(gla), _lib.ptr(m), stream())
    assert abs(float(m[2]) - float((0.1 * lpn - torch.min(q1, q2)).mean())) < 1e-5
    assert abs(float(m[4]) - float((0.1 * (-lpn + 2.0)).mean())) < 1e-5
    assert abs(float(gla) - 0.1 * float((-lpn + 2.0).mean())) < 1e-6
    assert abs(float(glp[0]) - 0.1 / B) < 1e-8
    assert torch.equal(d1 != 0, q1 <= q2)


# ------------------------------------------------------------------ CURL
@pytest.mark.parametrize('ntw', ['', '4', '8'])
@pytest.mark.parametrize('B,Bg,label0', [(12, 12, 0), (8, 24, 8), (70, 70, 0), (512, 512, 0), (128, 1024, 256), (64, 4096, 1024),
                                         (100, 300, 150)])
def test_curl_fwd_bwd(B, Bg, label0, ntw, monkeypatch):
    """(ntw: key columns per warp of the tensor-core kernels / 8 -- 4 and 8 give other column-block counts, so the
    single-block path without the reduce kernel and the multi-block path are both covered at every size)"""
    monkeypatch.setenv('CURLA_CURL_NTW', ntw)
    torch.manual_seed(7)
    feat = 50
    za = torch.zeros((B, 64), device=DEV); za[:, :feat] = torch.randn(B, feat, device=DEV)
    zp = torch.zeros((Bg, 64), device=DEV); zp[:, :feat] = torch.randn(Bg, feat, device=DEV)
    W = torch.rand(feat, feat, device=DEV)
    zar = za[:, :feat].clone().requires_grad_(True); Wr = W.clone().requires_grad_(True)
    logits = O.curl_logits(Wr, zar, zp[:, :feat])
    labels = torch.arange(B, device=DEV) + label0
    loss = F.cross_entropy(logits, labels, reduction='sum') / Bg
    loss.backward()
    ws = torch.zeros(int(_lib.load().curla_curl_workspace_floats(B, Bg)), device=DEV)
    lo = torch.zeros(1, device=DEV); dza = torch.zeros((B, 64), device=DEV); dW = torch.zeros((feat, feat), device=DEV)
    _lib.call('curla_curl_fwd_bwd', _lib.ptr(za), _lib.ptr(zp), _lib.ptr(W), B, Bg, feat, label0, 1.0 / Bg, _lib.ptr(ws),
              _lib.ptr(lo), _lib.ptr(dza), _lib.ptr(dW), None, stream())
    assert abs(float(lo) - float(F.cross_entropy(logits, labels))) < 1e-4 * max(1.0, abs(float(lo)))
    assert rel_l2(dza[:, :feat], zar.grad) < 1e-4
    assert rel_l2(dW, Wr.grad) < 1e-4


# ------------------------------------------------------------------ Adam / EMA / packers
def test_adam_matches_oracle_adam():
    torch.manual_seed(8)
    n = 10007
    p0 = torch.randn(n); lr, b1, b2 = 1e-3, 0.9, 0.999
    q = p0.clone().requires_grad_(True)
    opt = O.Adam([q], lr, (b1, b2))
    p = p0.clone().to(DEV); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    pd = p0.clone().to(DEV); md = torch.zeros(n, device=DEV); vd = torch.zeros(n, device=DEV)
    for t in range(1, 6):
        g = torch.randn(n) * (10.0 ** float(torch.randint(-4, 1, (1,))))
        q.grad = g.clone(); opt.step()
        gd = g.to(DEV)
        _lib.call('curla_adam_f32', _lib.ptr(p), _lib.ptr(gd), _lib.ptr(m), _lib.ptr(v), n, n, lr, b1, b2, 1e-8, t,
                  None, stream())
        # double-step variant on the upper half
        _lib.call('curla_adam_f32', _lib.ptr(pd), _lib.ptr(gd), _lib.ptr(md), _lib.ptr(vd), n, n // 2, lr, b1, b2, 1e-8,
                  t, None, stream())
        assert (p.cpu() - q.detach()).abs().max() < 2e-6, t
    st = opt.state[id(q)]
    assert max_rel(m.cpu(), st['m']) < 1e-5 and max_rel(v.cpu(), st['v']) < 1e-5
    # doubled elements moved exactly twice as far in total (same m, v)
    d1 = (p - p0.to(DEV)); d2 = (pd - p0.to(DEV))
    assert torch.equal(d1[:n // 2], d2[:n // 2])
    assert rel_l2(d2[n // 2:], 2 * d1[n // 2:]) < 1e-4
    # float64 scalar
    la = torch.tensor([np.log(0.1)], dtype=torch.float64, requires_grad=True)
    o2 = O.Adam([la], 1e-4, (0.5, 0.999))
    pa = la.detach().clone().to(DEV); sa = torch.zeros(2, dtype=torch.float64, device=DEV)
    for t in range(1, 4):
        gg = torch.tensor([0.3 * t - 0.5], dtype=torch.float64)
        la.grad = gg.clone(); o2.step()
        ggd = gg.to(DEV)
        _lib.call('curla_adam_f64_scalar', _lib.ptr(pa), _lib.ptr(ggd), _lib.ptr(sa), 1e-4, 0.5, 0.999, 1e-8, t,
                  None, stream())
        torch.cuda.synchronize()
    assert abs(float(pa) - float(la)) < 1e-12


def test_ema_bit_exact():
    torch.manual_seed(9)
    n = 5003
    p, t = torch.randn(n), torch.randn(n)
    want = t.clone()
    want[:2000] = 0.05 * p[:2000] + (1 - 0.05) * t[:2000]
    want[2000:] = 0.01 * p[2000:] + (1 - 0.01) * t[2000:]
    td, pd_ = t.to(DEV), p.to(DEV)
    _lib.call('curla_ema_f32', _lib.ptr(td), _lib.ptr(pd_), n, 2000, 0.05, 0.01, stream())
    assert torch.equal(td.cpu(), want)


def test_pack_rows():
    src = torch.randn(50 * 52 + 3, device=DEV)
    dst = torch.ones((64, 64), device=DEV, dtype=torch.bfloat16)
    seg = np.array([3, 0, 0, 50, 52, 64, 64], dtype=np.int64)
    _lib.call('curla_pack_shadows', _lib.ptr(src), _lib.ptr(dst), seg.ctypes.data_as(C.c_void_p), 1, stream())
    want = torch.zeros((64, 64), device=DEV); want[:50, :52] = src[3:].view(50, 52)
    assert torch.equal(dst, want.to(torch.bfloat16))


# ------------------------------------------------------------------ augmentations (K2/K3)
def test_noisy_cover_exact_with_injected_noise():
    """augmentations.py:172-205: cover rows exact, + injected noise, clamp -- equals the oracle."""
    from curla_b200 import augmentations
    from oracle import curla_oracle as O
    torch.manual_seed(0)
    aug = augmentations.make_augmentor('noisy_cover', (90, 160))
    assert (aug.top, aug.bottom) == O.noisy_cover_rows(90) == (28, 18)
    x = torch.randint(0, 256, (5, 9, 90, 160)).float()
    noise = torch.randn(5, 9, 90, 160) * 10.0
    np.random.seed(4)
    cover = [np.random.randint(0, 255) for _ in range(3)]
    want = O.noisy_cover(x, cover, noise, aug.top, aug.bottom)
    np.random.seed(4)                       # the augmentor draws the same three values itself
    from curla_b200 import augment
    got = augment.noisy_cover(x.to(DEV), aug, noise=noise.to(DEV))
    assert torch.equal(got.cpu(), want)
    assert np.random.randint(0, 255) == np.random.RandomState(4).randint(0, 255, size=4)[3]   # 3 draws consumed


def test_noisy_cover_builtin_noise_is_gaussian():
    from curla_b200 import augmentations
    aug = augmentations.make_augmentor('noisy_cover', (90, 160))
    x = torch.full((8, 9, 90, 160), 128.0, device=DEV)
    y = aug.training_augmentation(x.clone())
    mid = y.view(-1, 3, 90, 160)[:, :, aug.top:90 - aug.bottom, :] - 128.0        # uncovered rows
    assert abs(float(mid.mean())) < 0.05 and abs(float(mid.std()) - 10.0) < 0.05
    k = float(((mid / 10.0) ** 4).mean())
    assert abs(k - 3.0) < 0.1                                                      # Gaussian kurtosis
    cov = y.view(-1, 3, 90, 160)[:, :, :aug.top, :]
    per_ch = cov.mean(dim=(0, 2, 3))
    assert float(cov.std(dim=(0, 2, 3)).max()) < 10.5                               # constant + noise
    assert all(0 <= float(v) <= 255 for v in per_ch)
    y2 = aug.training_augmentation(x.clone())
    assert not torch.equal(y, y2)                                                   # fresh noise per call


def test_color_jiggle_matches_oracle_with_injected_params():
    """augmentations.py:106-136 with kornia's per-image parameters injected."""
    from curla_b200 import augment, augmentations
    from oracle import curla_oracle as O
    torch.manual_seed(1)
    aug = augmentations.make_augmentor('color_jiggle', (90, 160))
    B, fs = 4, 3
    x = torch.randint(0, 256, (B, 3 * fs, 90, 160)).float()
    n = B * fs
    contrast = 1.0 + 0.2 * (2 * torch.rand(n) - 1)
    saturation = 1.0 + 0.5 * (2 * torch.rand(n) - 1)
    hue = 0.5 * (2 * torch.rand(n) - 1)
    apply = (torch.rand(n) < 0.85).float()
    for order in ([0, 1, 2, 3], [3, 2, 1, 0], [2, 0, 3, 1]):
        want = O.color_jiggle(x, contrast, saturation, hue, apply, order)
        params = torch.stack([contrast, saturation, hue, apply]).to(DEV)
        got = augment.color_jiggle(x.clone().to(DEV), aug, params=params, order=order).cpu()
        # fp32 on both sides, different op order inside the HSV maths: 2e-3 of the 0..255 range
        assert float((got - want).abs().max()) < 0.5, (order, float((got - want).abs().max()))
        assert float((got - want).abs().mean()) < 2e-3
        untouched = apply.view(B, fs) == 0
        for b in range(B):
            for f in range(fs):
                if untouched[b, f]:
                    assert torch.equal(got[b, 3 * f:3 * f + 3], x[b, 3 * f:3 * f + 3])


def test_color_jiggle_builtin_parameter_distribution():
    from curla_b200 import augment, augmentations
    aug = augmentations.make_augmentor('color_jiggle', (90, 160))
    n = 3 * 1024
    x = torch.randint(0, 256, (n // 3, 9, 90, 160), device=DEV).float()[:, :, :8, :16].contiguous()
    aug8 = augmentations.ColorJiggle((8, 16))
    pout = torch.zeros(4, n, device=DEV)
    y = augment.color_jiggle(x.clone(), aug8, params_out=pout)
    c, s_, h, a = pout.cpu()
    assert 0.8 <= float(c.min()) and float(c.max()) <= 1.2 and abs(float(c.mean()) - 1.0) < 0.01
    assert 0.5 <= float(s_.min()) and float(s_.max()) <= 1.5 and abs(float(s_.mean()) - 1.0) < 0.03
    assert -0.5 <= float(h.min()) and float(h.max()) <= 0.5 and abs(float(h.mean())) < 0.03
    assert abs(float(a.mean()) - 0.85) < 0.03
    assert float(y.min()) >= 0.0 and float(y.max()) <= 255.0 + 1e-3
    same = (y.view(n, -1) == x.view(n, -1)).all(dim=1).cpu()
    assert bool((same == (a == 0)).all())           # exactly the non-selected frames are untouched
