"""Data-parallel pretrain_AE check, run under torchrun on N >= 2 GPUs of one box:
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
(1) the all-reduced gradient of N ranks x B/N segments equals the single-rank gradient of the B-segment batch
    (trainer.py:325-329 on the concatenated batch; dropout off, shared Gumbel noise);
(2) after real steps every rank holds bit-identical parameters and the loss went down;
(3) ms/step with the decoder all-reduce overlapped with the encoder backward vs the same step without any exchange."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200  # noqa: E402,F401
from zs_b200 import synthetic as syn, train as zt  # noqa: E402
from zs_b200.model import Decoder, Encoder, gumbel_from_uniform  # noqa: E402


def nets(dp):
    enc = Encoder(ns=0.01, dp=dp, enc_size=1024, seg_len=128, enc_mode='one_hot')
    dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
    enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode='one_hot'))
    dec.load_state_dict(syn.decoder_state_dict(0, c_in=1024, c_h=1024, c_a=102))
    return enc.cuda().train(), dec.cuda().train()


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    per = 32
    B = per * world
    x, c = syn.spectrogram_batch(B, 128, 0).cuda(), syn.speaker_ids(B, 102, 0).cuda()
    noise = gumbel_from_uniform(syn.gumbel_uniform((B, 16, 1024), 0)).cuda()
    sl = slice(rank * per, (rank + 1) * per)

    # (1) gradient identity
    step = zt.PretrainAE(*nets(0.0), loss_scale=2.0 ** 15 * B / world)
    step.forward_backward(x[sl], c[sl], noise[sl], 0, None)
    torch.cuda.current_stream().wait_stream(step.side)
    zt.reduce_gradients(step.enc.grad)
    g_dp = [step.enc.grad / world, step.dec.grad / world]
    solo = [dist.new_group([r]) for r in range(world)][rank]     # a world of one: the reference step on the whole batch
    ref = zt.PretrainAE(*nets(0.0), loss_scale=2.0 ** 15 * B / world, process_group=solo)
    assert ref.world == 1
    ref.forward_backward(x, c, noise, 0, None)
    torch.cuda.synchronize()
    # both runs carry the same loss scale; the local loss is a mean over B/N segments, so sum/N == full-batch mean
    for name, a, b in zip(('encoder', 'decoder'), g_dp, [ref.enc.grad, ref.dec.grad]):
        rel = ((a - b).norm() / b.norm()).item()
        if rank == 0:
            print(f'(1) {name}: |g_dp - g_full| / |g_full| = {rel:.3e}  (|g_full| = {b.norm().item() / ref.loss_scale:.3e})')
        assert rel < 2e-2, rel
    del step, ref

    # (2) real steps stay in lock-step
    step = zt.PretrainAE(*nets(0.5))
    xs = [syn.spectrogram_batch(per, 128, 10 * rank + i).cuda() for i in range(4)]
    cs = [syn.speaker_ids(per, 102, 10 * rank + i).cuda() for i in range(4)]
    losses = []
    for i in range(12):
        losses.append(step.step(xs[i % 4], cs[i % 4]).item())
    flat = torch.cat([step.enc.flat, step.dec.flat])
    mine = torch.stack([flat.double().sum(), flat.double().abs().sum()])
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    assert all(torch.equal(every[0], e) for e in every), 'ranks diverged'
    assert losses[-1] < losses[0] and step.n_skipped == 0
    if rank == 0:
        print(f'(2) {world} ranks bit-identical after 12 steps; local loss {losses[0]:.4f} -> {losses[-1]:.4f}')

    # (3) timing
    def timed(s, n=20):
        for i in range(3):
            s.step(xs[i % 4], cs[i % 4])
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            s.step(xs[i % 4], cs[i % 4])
        e1.record(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    t_dp = timed(step)
    alone = zt.PretrainAE(*nets(0.5), process_group=solo, use_graph=False)
    t_alone = timed(alone)
    if rank == 0:
        print(f'(3) B={per}/rank x {world}: {t_dp:.3f} ms/step with the NCCL all-reduce (221 MB fp32), '
              f'{t_alone:.3f} ms/step eager without exchange -> {per * 128 * world / t_dp * 1e3:.0f} frames/s')
    # (4) sharded inference == one process: utterance shards, reference-order noise, ordered gather
    import numpy as np
    from zs_b200.frontend import AutoencoderPath
    from zs_b200.shard import ShardedPath
    enc_i, dec_i = nets(0.5)
    enc_i.eval(); dec_i.eval()
    path = AutoencoderPath(enc_i, dec_i, seg_len=128, max_batch=64)
    rng = np.random.Generator(np.random.PCG64(3))
    specs = [np.clip(rng.random((int(n), 513), dtype=np.float32), 1e-8, 1) for n in (5, 300, 128, 1000, 77, 640, 2000, 131)]
    spk = [int(i) % 102 for i in range(len(specs))]
    torch.manual_seed(11)
    out_s, units_s = ShardedPath(path).convert_utterances(specs, spk)
    torch.manual_seed(11)
    out_1, units_1 = path.convert_utterances(specs, spk)
    same = all(np.array_equal(a, b) for a, b in zip(units_s, units_1)) and all(np.array_equal(a, b) for a, b in zip(out_s, out_1))
    assert same, 'sharded inference differs from the single-process run'
    if rank == 0:
        print(f'(4) sharded convert_utterances over {world} ranks == single process, bit for bit ({len(specs)} utterances)')
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
