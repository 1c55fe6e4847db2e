"""World-size-2 `gloo` tests of the N > 1 host logic: utterance sharding with no data-path
collective, reference-order Gumbel draws under sharding, the ordered gather, and the gradient all-reduce contract of
the pretrain_AE step (sum over ranks / world == gradient of the concatenated batch).  The compute stand-in is the
oracle (CPU fp32) on a narrow configuration - the CUDA path itself is exercised by the `-m gpu` tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200  # noqa: E402,F401
from zs_b200 import shard, synthetic as syn  # noqa: E402
from zs_b200.frontend import segment_plan  # noqa: E402

SMALL = dict(c_in=24, c_h1=8, c_h2=64, c_h3=64, enc_size=16)       # c_h3 = 64: the GRU slice unit of the CUDA path
SMALL_DEC = dict(c_in=16, c_out=24, c_h=64, c_a=4)
SEG = 64


def _spawn(fn, world=2, *args):
    port = 29600 + os.getpid() % 300
    mp.spawn(_entry, args=(fn, world, port, args), nprocs=world, join=True)


def _entry(rank, fn, world, port, args):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def test_shards_partition_and_balance():
    rng = np.random.default_rng(0)
    lengths = [int(n) for n in rng.integers(3, 2000, size=97)]
    for world in (1, 2, 4, 8):
        shards = shard.shard_utterances(lengths, world, 128)
        assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
        loads = [sum(sum(e - s for s, e in segment_plan(lengths[i], 128)[1]) for i in sh) for sh in shards]
        assert max(loads) - min(loads) <= max(lengths)          # longest-first greedy bound
    assert shard.shard_utterances([], 2) == [[], []]


def _oracle_units(specs, only, enc_sd):
    """encode() of the reference driver (convert.py:183-221) with the oracle as the model, noise in call order."""
    from oracle import ae_oracle as orc
    lengths = [len(s) for s in specs]
    noise = shard.reference_order_noise(lengths, SEG, lambda b, t: (b, (t + 7) // 8, SMALL['enc_size']),
                                        lambda shape: torch.rand(shape), only)
    out = []
    for u in only:
        spec = np.asarray(specs[u], np.float32)
        padded, plan, keep = segment_plan(len(spec), SEG)
        if padded > len(spec):
            spec = np.concatenate([spec, np.zeros((padded - len(spec), spec.shape[1]), np.float32)])
        parts = []
        for (s, e), uni in zip(plan, noise[u]):
            x = torch.from_numpy(spec[s:e].T.copy())[None]
            with torch.no_grad():
                act, _, _ = orc.encoder_forward(enc_sd, x, uni, seg_len=SEG, enc_size=SMALL['enc_size'])
            parts.append(act[0].T.numpy())
        e = np.concatenate(parts)
        out.append(e[:keep] if keep is not None else e)
    return out


def _sharded_encode(rank, world, tmp):
    rng = np.random.default_rng(1)
    specs = [rng.random((int(n), SMALL['c_in']), dtype=np.float32) for n in (5, 70, 64, 200, 131, 9, 330)]
    enc_sd = syn.encoder_state_dict(0, **SMALL)
    mine = shard.shard_utterances([len(s) for s in specs], world, SEG)[rank]
    torch.manual_seed(7)                                       # the reference's RNG contract: one CPU stream
    local = _oracle_units(specs, mine, enc_sd)
    everything = shard.gather_in_order(local, mine, len(specs))
    # the tensor-collective gather the product uses: int32 unit ids in one flat buffer per rank, shapes known from the lengths
    shards = shard.shard_utterances([len(s) for s in specs], world, SEG)
    n_units = [shard.output_lengths(len(s), SEG)[0] for s in specs]
    ids = shard.gather_arrays([a.argmax(1).astype(np.int32) for a in local], shards, [(n,) for n in n_units], np.int32)
    for a, i in zip(everything, ids):
        assert np.array_equal(a.argmax(1), i) and a.shape[0] == i.shape[0]
    if rank == 0:
        torch.manual_seed(7)
        single = _oracle_units(specs, list(range(len(specs))), enc_sd)
        assert len(everything) == len(single)
        for a, b in zip(everything, single):
            assert a.shape == b.shape and np.array_equal(a, b)  # same draws, same segments: identical one-hot units
        open(os.path.join(tmp, 'ok_encode'), 'w').write('1')


def test_sharded_encode_equals_single_process(tmp_path):
    _spawn(_sharded_encode, 2, str(tmp_path))
    assert os.path.exists(tmp_path / 'ok_encode')


def _dp_step(rank, world, tmp):
    from oracle import ae_oracle as orc
    from zs_b200.train import reduce_gradients
    B, T = 4, 64
    enc_sd, dec_sd = syn.encoder_state_dict(0, **SMALL), syn.decoder_state_dict(0, **SMALL_DEC)
    x = syn.spectrogram_batch(B, T, 0, c_in=SMALL['c_in'])
    c = syn.speaker_ids(B, SMALL_DEC['c_a'], 0)
    u = syn.gumbel_uniform((B, T // 8, SMALL['enc_size']), 0)
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    _, g_enc, g_dec, _, _ = orc.ae_loss_and_grads(enc_sd, dec_sd, x[sl], c[sl], u[sl], dp=0.0, seg_len=SEG)
    names_e, names_d = list(enc_sd), list(dec_sd)
    flat = torch.cat([g_enc[k].reshape(-1) for k in names_e] + [g_dec[k].reshape(-1) for k in names_d])
    reduce_gradients(flat)
    flat /= world
    # every rank now holds the same gradient ...
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    assert all(torch.equal(both[0], b) for b in both)
    if rank == 0:   # ... and it is the gradient of the reference step on the whole batch (trainer.py:325-329)
        _, G_enc, G_dec, _, _ = orc.ae_loss_and_grads(enc_sd, dec_sd, x, c, u, dp=0.0, seg_len=SEG)
        full = torch.cat([G_enc[k].reshape(-1) for k in names_e] + [G_dec[k].reshape(-1) for k in names_d])
        assert full.abs().max() > 0
        assert (flat - full).abs().max() <= 1e-5 * full.abs().max() + 1e-9
        open(os.path.join(tmp, 'ok_dp'), 'w').write('1')


def test_gradient_allreduce_equals_full_batch_step(tmp_path):
    _spawn(_dp_step, 2, str(tmp_path))
    assert os.path.exists(tmp_path / 'ok_dp')


def test_output_lengths_follow_the_chunk_plan():
    # convert.py:139-168: L = 2000, seg_len 128 -> 14 chunks of 128 + one of 207 (its last frame dropped) -> 208 frames out
    assert shard.output_lengths(2000, 128) == (14 * 16 + 26, 14 * 128 + 208)
    assert shard.output_lengths(5, 128) == (1, 16)            # padded to MIN_LEN = 9: 2 unit frames, 1 kept
    assert shard.output_lengths(128, 128) == (16, 128)


def test_gather_arrays_single_process_and_size_check():
    out = shard.gather_arrays([np.arange(3, dtype=np.int32), np.arange(2, dtype=np.int32)], [[1, 0]], [(2,), (3,)], np.int32)
    assert np.array_equal(out[1], [0, 1, 2]) and np.array_equal(out[0], [0, 1])
    with pytest.raises(RuntimeError):
        shard.gather_arrays([np.arange(3, dtype=np.int32)], [[0]], [(2,)], np.int32)


def test_gather_detects_missing_and_duplicate():
    with pytest.raises(RuntimeError):
        shard.gather_in_order(['a'], [0], 2)
    with pytest.raises(ValueError):
        shard.gather_in_order(['a', 'b'], [0], 2)
