"""OPEN ITEM probe (not run in round 2 for lack of GPU minutes): weight-gradient side streams (zs_wgrad_async) in a
data-parallel pretrain_AE step.  `tools/dp_check.py` with them enabled at world 2 did not finish within 300 s; PretrainAE
therefore enables them for one rank only.  This script forces them on and walks the step stage by stage so that the stage
that blocks is the last line printed:

  timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      tools/wgrad_async_dp_probe.py
"""
import faulthandler
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200  # noqa: E402,F401
from zs_b200 import synthetic as syn, train as zt  # noqa: E402
from zs_b200.model import Decoder, Encoder  # noqa: E402


def nets():
    enc = Encoder(ns=0.01, dp=0.5, enc_size=1024, seg_len=128, enc_mode='one_hot')
    dec = Decoder(ns=0.01, c_in=1024, c_h=1024, c_a=102, seg_len=128)
    enc.load_state_dict(syn.encoder_state_dict(0, enc_size=1024, enc_mode='one_hot'))
    dec.load_state_dict(syn.decoder_state_dict(0, c_in=1024, c_h=1024, c_a=102))
    return enc.cuda().train(), dec.cuda().train()


def say(rank, msg):
    print(f'[rank {rank}] {msg}', flush=True)


def main():
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    faulthandler.dump_traceback_later(40, exit=True)        # a blocked stage prints where the host thread waits
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    x = [syn.spectrogram_batch(32, 128, 10 * rank + i).cuda() for i in range(4)]
    c = [syn.speaker_ids(32, 102, 10 * rank + i).cuda() for i in range(4)]
    for graph in (False, True):
        step = zt.PretrainAE(*nets(), async_wgrad='force', use_graph=graph)
        say(rank, f'use_graph={graph}: built, async_wgrad={step.async_wgrad}')
        for i in range(6):
            loss = step.step(x[i % 4], c[i % 4])
            torch.cuda.synchronize()
            say(rank, f'  step {i} done (graph captured: {step._graph is not None}), loss {loss.item():.4f}')
        dist.barrier()
    say(rank, 'all stages finished: the side streams coexist with the captured all-reduces here')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
