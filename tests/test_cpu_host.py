"""CPU-side checks (no GPU): the C-ABI library builds, loads and exports every symbol the header declares,
the ctypes mirrors match the C structs, the drop-in modules keep the reference checkpoint contract, the
host-side segmentation matches the reference's encode() replay, and the product path refuses to run on CPU."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import zs_b200  # noqa: F401
from zs_b200 import _lib, frontend, synthetic as syn
from zs_b200.model import Decoder, Encoder
from conftest import GOLDEN, ROOT


@pytest.fixture(scope='module')
def built():
    return _lib.build()


def test_library_builds_and_exports_every_header_symbol(built):
    header = open(_lib.HEADER).read()
    declared = set(re.findall(r'\b(zs_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found in include/zs_ae.h'
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in zs_ae.h but not exported by libzsae.so'
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    assert lib.zs_version() >= 100


def test_ctypes_structs_match_c_layout(built, tmp_path):
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "zs_ae.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(zs_conv_desc),sizeof(zs_encoder_weights),sizeof(zs_decoder_weights),sizeof(zs_encoder_cfg),'
                   'sizeof(zs_decoder_cfg),offsetof(zs_conv_desc,out),offsetof(zs_conv_desc,bias));return 0;}')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    got = list(map(int, subprocess.check_output([str(exe)]).split()))
    want = [ctypes.sizeof(_lib.ConvDesc), ctypes.sizeof(_lib.EncoderWeights), ctypes.sizeof(_lib.DecoderWeights),
            ctypes.sizeof(_lib.EncoderCfg), ctypes.sizeof(_lib.DecoderCfg), _lib.ConvDesc.out.offset,
            _lib.ConvDesc.bias.offset]
    assert got == want


def test_sass_is_blackwell_native(built):
    sass = subprocess.run(['cuobjdump', '-sass', built], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM'):     # tcgen05.mma, TMA, tcgen05.ld
        assert mnemonic in sass, mnemonic
    assert 'HMMA.16816' not in sass                      # no legacy mma.sync path


@pytest.mark.parametrize('enc_mode', ['one_hot', 'multilabel_binary', 'continues', 'gumbel_t'])
def test_state_dict_contract(enc_mode):
    enc = Encoder(ns=0.01, dp=0.5, enc_size=1024, seg_len=128, enc_mode=enc_mode)
    want = syn.encoder_shapes(enc_size=1024, enc_mode=enc_mode)
    got = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    assert list(got) == list(want)
    assert got == {k: s for k, (s, _) in want.items()}
    enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode=enc_mode), strict=True)
    dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
    wantd = syn.decoder_shapes(c_in=1024, c_h=1024, c_a=102)
    gotd = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    assert list(gotd) == list(wantd)
    assert gotd == {k: s for k, (s, _) in wantd.items()}
    assert sum(p.numel() for p in enc.parameters()) == (12759936 if enc_mode != 'multilabel_binary' else 12759936 + 1024 * 769)
    assert sum(p.numel() for p in dec.parameters()) == 42488321


def test_binary_mode_contract():
    """enc_mode 'binary' (model/model.py:391-392, 466-472): an enc_size^2 projection; supported up to enc_size 128."""
    enc = Encoder(enc_size=16, enc_mode='binary')
    assert tuple(enc.linear.weight.shape) == (256, 768) and enc.noise_shape(2, 128) == (2, 16, 16, 16)
    want = syn.encoder_shapes(enc_size=16, enc_mode='binary')
    assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == {k: s for k, (s, _) in want.items()}
    with pytest.raises(NotImplementedError):
        Encoder(enc_size=1024, enc_mode='binary')          # 1 M output channels: rejected loudly
    with pytest.raises(NotImplementedError):
        Encoder(enc_mode='nonsense')


def test_no_cpu_fallback(built):
    enc = Encoder(ns=0.01, enc_size=32, seg_len=128, enc_mode='one_hot', c_in=33, c_h1=16, c_h2=64, c_h3=16)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        enc(torch.rand(1, 33, 40))
    dec = Decoder(c_in=32, c_out=33, c_h=64, c_a=5, ns=0.01, seg_len=128)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        dec(torch.rand(1, 32, 5), torch.zeros(1, dtype=torch.long))
    if not torch.cuda.is_available():
        assert _lib.lib().zs_device_check() != 0          # compute entry points fail loudly without a device
        assert _lib.lib().zs_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'zerospeech-tts-without-t_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in text.replace('oracle/_ref', ''), f


def test_segment_plan_matches_reference_replay():
    plans = json.load(open(os.path.join(GOLDEN, 'chunk_plans.json')))
    for key, ref in plans.items():
        seg_len, L = map(int, key.split(':'))
        if 'error' in ref:
            with pytest.raises(RuntimeError):
                frontend.segment_plan(L, seg_len)
            continue
        padded, plan, keep = frontend.segment_plan(L, seg_len)
        assert [[s, e - s] for s, e in plan] == ref['calls'], key
        units = sum(Encoder.t8(e - s) for s, e in plan)
        assert (min(units, keep) if keep is not None else units) == ref['n_units'], key
        assert all(9 <= e - s <= 2 * seg_len - 1 for s, e in plan)


def test_write_encodings_format(tmp_path):
    p = tmp_path / 'e.txt'
    frontend.write_encodings(str(p), np.array([[0., 1., 0.], [1., 0., 0.]], dtype=np.float32))
    assert p.read_text() == '0 1 0\n1 0 0\n'


def test_write_unit_ids_equals_write_encodings(tmp_path):
    """convert.py:120-126: the fast id writer produces byte-identical files to the reference format written from one-hot rows."""
    rng = np.random.default_rng(0)
    for enc_size, n in ((1024, 37), (6, 5), (512, 1), (32, 0)):
        ids = rng.integers(0, enc_size, size=n)
        a, b = tmp_path / 'a.txt', tmp_path / 'b.txt'
        frontend.write_unit_ids(str(a), ids, enc_size)
        frontend.write_encodings(str(b), frontend.one_hot_rows(ids, enc_size))
        assert a.read_bytes() == b.read_bytes()
    with pytest.raises(RuntimeError):
        frontend.write_unit_ids(str(tmp_path / 'c.txt'), [7], 6)


def test_synthetic_is_deterministic():
    a, b = syn.encoder_state_dict(0, enc_size=32, c_in=33, c_h1=16, c_h2=64, c_h3=16), \
        syn.encoder_state_dict(0, enc_size=32, c_in=33, c_h1=16, c_h2=64, c_h3=16)
    assert all(torch.equal(a[k], b[k]) for k in a)
    x = syn.spectrogram_batch(2, 16, 0)
    assert x.shape == (2, 513, 16) and float(x.min()) >= 1e-8 and float(x.max()) <= 1.0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): ONE JSON line on stdout with the contract's
    keys, on a bounded sample, without touching a GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1', '--cpu-sample', '2'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['metric'] == 'spectrogram frames/s, AE encode+decode' and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']
    # the record says what ran: --steps / --warmup honoured, the bounded sample and the port named
    assert d['steps'] == 2 and d['warmup'] == 1
    assert '2 segments' in d['config']['reference_sample'] and 'port' in d['config']['reference_sample']


def test_bench_has_single_definitions():
    """a botched paste once left every helper of bench.py defined twice (the later, older copy winning)."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tree = ast.parse(open(os.path.join(root, 'bench.py')).read())
    names = [n.name for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))]
    assert len(names) == len(set(names)), sorted(n for n in names if names.count(n) > 1)
    src = open(os.path.join(root, 'bench.py')).read()
    assert 'environ.pop' not in src and "environ['NCCL_DEBUG']" not in src       # the driver reads NCCL's own log: leave it on


def test_front_end_copy_pool_covers_every_index_once():
    """The batched front-end splits its segment gather / result scatter over a small thread pool (frontend._parallel_ranges):
    contiguous slices, every index exactly once, inline below two tasks' worth of work, exceptions of a worker surface."""
    import numpy as np
    from zs_b200 import frontend as fe
    pool = fe._copy_pool()
    assert pool is fe._copy_pool() and pool._max_workers >= 1
    for n in (0, 1, 31, 32, 100, 960, 1001):
        hits = np.zeros(n, np.int64)
        slices = []

        def fn(lo, hi):
            slices.append((lo, hi))
            hits[lo:hi] += 1
        fe._parallel_ranges(pool, n, fn)
        assert (hits == 1).all(), n
        assert sorted(slices) == sorted(set(slices)) and all(lo < hi for lo, hi in slices if n)
        if n < 32:
            assert slices in ([(0, n)], [(0, 0)]) or n == 0

    def boom(lo, hi):
        raise ValueError('worker failed')
    with pytest.raises(ValueError, match='worker failed'):
        fe._parallel_ranges(pool, 960, boom)
