"""Griffin-Lim on a fixed batch: the command ncu wraps for the gl_iter_kernel capture."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zs_b200  # noqa: E402,F401
from zs_b200 import dsp  # noqa: E402


def main():
    n_utt, frames, n_iter = 64, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = np.random.Generator(np.random.PCG64(5))
    rows = torch.from_numpy(np.clip(rng.random((n_utt * frames, 513), dtype=np.float32), 1e-8, 1)).cuda()
    gl = dsp.GriffinLim('cuda', n_iter=n_iter)
    for _ in range(2):
        out = gl.synthesize(rows, [frames] * n_utt, trim=False, to_host=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = gl.synthesize(rows, [frames] * n_utt, trim=False, to_host=False)
    e1.record()
    torch.cuda.synchronize()
    print(f'{n_iter} iterations on {n_utt * frames} frames: {e0.elapsed_time(e1):.2f} ms, {e0.elapsed_time(e1) / max(n_iter, 1) * 1e3:.1f} us per iteration')


if __name__ == '__main__':
    main()
