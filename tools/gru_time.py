"""Times the bi-GRU recurrence alone (zs_gru_recurrence, H = 512, T = 128) for the given batch sizes:
python tools/gru_time.py 300 470 960   ->  us per call and per step; the kernel the launch rule picks is named."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import zs_b200  # noqa: E402,F401
from zs_b200 import _lib  # noqa: E402
import gpu_helpers as gh  # noqa: E402


def main():
    H, T = 512, 128
    lib = _lib.lib()
    for B in [int(a) for a in sys.argv[1:]] or [960]:
        w = (torch.rand(2, 3 * H, H, device='cuda') * 2 - 1) / H ** 0.5
        b = torch.zeros(2, 3 * H, device='cuda')
        gx = torch.randn(B, T, 2, 3 * H, device='cuda')
        out = torch.zeros(B, T, 2 * H, dtype=torch.float16, device='cuda')

        def run():
            _lib.check(lib.zs_gru_recurrence(gh.ptr(gx), gh.ptr(w), gh.ptr(b), B, T, H, gh.ptr(out), T, 2 * H, 0, 0, 0, 0, gh.stream()))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'B={B}: {ms * 1000:.1f} us per call (incl. the W_hh pack + gx cast), {ms / T * 1000:.2f} us per step, finite={bool(torch.isfinite(out.float()).all())}')


if __name__ == '__main__':
    main()
