"""CPU-side checks: the C-ABI library builds, loads and exports every symbol declared in
include/curla_b200.h; host-side layout logic of the engine; no compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from curla_b200 import build
    build.build()
    from curla_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    from curla_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'curla_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = set(re.findall(r'\b(curla_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) > 40
    for n in sorted(names):
        assert hasattr(lib, n), 'missing export: ' + n
        assert n in _lib.SIGNATURES, 'no ctypes signature for ' + n
    assert set(_lib.SIGNATURES) <= names
    assert lib.curla_version() >= 100


def test_header_is_plain_c_and_structs_match_ctypes(tmp_path):
    """include/curla_b200.h compiles as C99 (the boundary is a C ABI) and the structs the Python side
    mirrors with ctypes have the same size and field offsets."""
    import subprocess
    from curla_b200 import _lib
    src = tmp_path / 'probe.c'
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "curla_b200.h"\n'
        'int main(void) {\n'
        '  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(curla_conv_seg), offsetof(curla_conv_seg, in), offsetof(curla_conv_seg, wts),\n'
        '         offsetof(curla_conv_seg, bias), offsetof(curla_conv_seg, out), offsetof(curla_conv_seg, B));\n'
        '  printf("%zu %zu\\n", sizeof(curla_agent_config), sizeof(curla_update_args));\n'
        '  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(curla_conv_stack_seg), offsetof(curla_conv_stack_seg, in),\n'
        '         offsetof(curla_conv_stack_seg, w96), offsetof(curla_conv_stack_seg, bias), offsetof(curla_conv_stack_seg, out),\n'
        '         offsetof(curla_conv_stack_seg, B));\n'
        '  return 0;\n}\n')
    exe = tmp_path / 'probe'
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    vals = [int(v) for v in out]
    cs = _lib.ConvSeg
    assert vals[:6] == [C.sizeof(cs), cs.inp.offset, cs.wts.offset, cs.bias.offset, cs.out.offset, cs.B.offset]
    assert vals[6] == C.sizeof(_lib.AgentConfig) and vals[7] == C.sizeof(_lib.UpdateArgs)
    ss = _lib.ConvStackSeg
    assert vals[8:14] == [C.sizeof(ss), ss.inp.offset, ss.w96.offset, ss.bias.offset, ss.out.offset, ss.B.offset]


def test_host_only_helpers(lib, monkeypatch):
    """Pure host functions of the C ABI (no GPU): split-K ranges are multiples of 64 -- the K step of the
    tcgen05 GEMM -- so the number of partial slices the engine allocates does not depend on which GEMM
    kernel runs; the conv padding covers a slab's halo."""
    for K, sp in ((67456, 74), (4096, 7), (64, 3), (88768, 74), (1000, 1), (8, 5)):
        z = lib.curla_gemm_effective_splits(K, sp)
        steps = -(-K // 64)
        per = -(-steps // sp)
        assert z == -(-steps // per) and 1 <= z <= sp, (K, sp, z)
        assert (z - 1) * per * 64 < K <= z * per * 64            # every slice non-empty, K covered
    assert lib.curla_gemm_effective_splits(67456, 74) == 71      # the encoder fc forward at train.py defaults
    for pitch in (32, 42, 68, 80):
        assert lib.curla_conv_pad_rows(pitch) >= 256 + 2 * pitch + 2
    # conv-2..4 fused in shared memory: the default 76 x 135 crop fits one SM (2552 positions x 64 B + weights),
    # the 90 x 160 identity / pixel-SAC input does not and runs layer by layer
    I3 = C.c_int * 3
    monkeypatch.setenv('CURLA_CONV_FUSED', '1')
    assert lib.curla_conv_stack_fits(68, 38 * 68, I3(35, 33, 31), I3(65, 63, 61)) == 1
    assert lib.curla_conv_stack_fits(80, 45 * 80, I3(42, 40, 38), I3(77, 75, 73)) == 0
    assert lib.curla_conv_stack_fits(42, 42 * 42, I3(39, 37, 35), I3(39, 37, 35)) == 1          # 84 x 84


def test_engine_layout_host_side(lib):
    """curla_agent_create is pure host code: check the memory plan without a GPU."""
    from curla_b200 import _lib
    c = _lib.AgentConfig()
    vals = dict(C=9, H=76, W=135, Hf=90, Wf=160, feature_dim=50, hidden_dim=1024, action_dim=2, num_filters=32,
                num_layers=4, batch=512, global_batch=512, rank=0, world=1, actor_update_freq=2,
                critic_target_update_freq=2, cpc_update_freq=1)
    for k, v in vals.items():
        setattr(c, k, v)
    h = lib.curla_agent_create(C.byref(c))
    assert h
    info = {}
    name = C.create_string_buffer(256)
    arena, off, ndim, dt = C.c_int(), C.c_longlong(), C.c_int(), C.c_int()
    dims = (C.c_longlong * 4)()
    for i in range(lib.curla_agent_num_tensors(h)):
        assert lib.curla_agent_tensor_info(h, i, name, 256, C.byref(arena), C.byref(off), C.byref(ndim), dims,
                                           C.byref(dt)) == 0
        info[name.value.decode()] = (arena.value, off.value, [dims[k] for k in range(ndim.value)], dt.value)
    # fc_in 60,512 = 32*31*61 lives in a [50][31][68][32] canonical tensor
    assert info['critic.encoder.fc.weight_canon'][2] == [50, 4, 31 * 68, 8]     # [feat][planes][Ho*pitch][8]
    assert info['critic.Q1.trunk.0.weight'][2] == [1024, 52]
    assert info['actor.trunk.4.weight'][2] == [4, 1024]
    # critic and target segments have identical structure (EMA runs flat over them)
    c0, t0 = info['critic.encoder.convs.0.weight'][1], info['target.encoder.convs.0.weight'][1]
    for k, v in info.items():
        if k.startswith('critic.') and v[0] == 0:
            assert info['target.' + k[len('critic.'):]][1] - t0 == v[1] - c0, k
    # [W | critic.encoder] is contiguous: the CURL Adam covers it in one launch
    assert info['CURL.W'][1] == 0 and c0 == 2500 * 4
    n_crit = info['grad.critic'][2][0]
    assert n_crit >= 5_265_912 and info['grad.actor'][2][0] >= 4_162_042 - 9248 * 3 - 2624
    assert lib.curla_agent_arena_bytes(h, 4) > 512 * 2584 * 64 * 12
    # bad configs fail loudly
    c.num_filters = 16
    assert not lib.curla_agent_create(C.byref(c))
    assert b'num_filters' in lib.curla_last_error()
    lib.curla_agent_destroy(h)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'curla_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'oracle' not in src.replace('CPU oracle', ''), f


def test_no_cpu_fallback():
    import torch
    from curla_b200 import _lib, utils, augmentations
    with pytest.raises(_lib.CurlaError):
        utils.ReplayBuffer((9, 90, 160), (2,), 8, 4, torch.device('cpu'),
                           augmentations.IdentityAugmentation((90, 160)))


@pytest.mark.parametrize('case', ['init_seed0', 'init_seed3_identity'])
def test_init_rng_matches_reference(case):
    """The product draws its initial weights from the torch global RNG in the reference
    constructor's order (curl_sac.py:271-317: Actor, Critic, target Critic, CURL.W): for a given
    seed every initial tensor equals the reference's freshly constructed agent
    (tests/golden/init_*.npz, written by oracle/make_golden.py from the unmodified reference),
    and the stream stands at the same place afterwards."""
    import numpy as np
    import torch
    from curla_b200 import curl_sac, utils
    from oracle import scenario as S
    from oracle.make_golden import INIT_CASES
    seed, obs_shape, hidden = INIT_CASES[case]
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', case + '.npz'))
    utils.set_seed_everywhere(seed)
    actor_sd, critic_sd, W = curl_sac.initial_state(obs_shape, (S.ACTION_DIM,), hidden, S.FEATURE_DIM)
    after = torch.rand(4).numpy()
    n = 0
    for net, sd in (('actor', actor_sd), ('critic', critic_sd), ('target', critic_sd)):
        for k, v in sd.items():
            assert np.array_equal(S.summarize(v), gold['param/%s.%s' % (net, k)]), (net, k)
            n += 1
    assert n == 3 * 12 + 6 + 12 + 12                        # every tensor of the three state dicts
    assert {k for k in gold.files if k.startswith('param/actor.')} == {'param/actor.' + k for k in actor_sd}
    assert np.array_equal(S.summarize(W), gold['param/W'])
    assert np.array_equal(after, gold['next_rand'])
    assert float(gold['param/log_alpha'][0]) == float(np.log(S.HP['init_temperature']))


def test_one_shot_gemm_kernels_stay_small(lib):
    """Code size is a performance property of the one-shot tcgen05 GEMM kernels ("Instruction
    fetch"): a CTA executes its epilogue exactly once, every instruction a cold fetch behind the operand stream of all
    SMs -- with the epilogue's row loops unrolled (6,300 instructions per instantiation, 6,160 for the N-loop kernel)
    the fc forward spent 7 of its 29 us there.  Guard the plain-loop form: instruction counts from cuobjdump."""
    import shutil
    import subprocess
    from curla_b200 import build
    exe = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(exe):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([exe, '-sass', build.OUT], capture_output=True, text=True).stdout
    counts, name = {}, None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            counts[name] = 0
        elif name and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
            counts[name] += 1
    gemm = {k: v for k, v in counts.items() if 'k_gemm_tcIL' in k}
    nloop = {k: v for k, v in counts.items() if 'k_gemm_tc_nloop' in k}
    assert len(gemm) == 8 and len(nloop) == 2, (sorted(gemm), sorted(nloop))
    assert max(gemm.values()) < 4500, gemm
    assert max(nloop.values()) < 3000, nloop


def test_scripts_compile():
    """bench.py, __graft_entry__.py and the manual experiment / profiling scripts at least byte-compile (they only run
    on the GPU box, where a syntax error costs a GPU call)."""
    import glob
    import py_compile
    files = [os.path.join(ROOT, 'bench.py'), os.path.join(ROOT, '__graft_entry__.py')]
    files += sorted(glob.glob(os.path.join(ROOT, 'tests', 'manual', '*.py')))
    files += sorted(glob.glob(os.path.join(ROOT, 'tests', 'bench_*.py')))
    files += sorted(glob.glob(os.path.join(ROOT, 'profiles', 'tools', '*.py')))
    assert len(files) >= 8
    for f in files:
        py_compile.compile(f, doraise=True)
