"""Times the GRU cluster kernel alone under the ZS_GRU_DEBUG experiment flags (timing only)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    import zs_b200
    from zs_b200 import _lib
    import gpu_helpers as gh
    H, B, T = 512, int(sys.argv[2]), 128
    w = (torch.rand(2, 3 * H, H, device='cuda') * 2 - 1) / H ** 0.5
    b = torch.zeros(2, 3 * H, device='cuda')
    gx = torch.randn(B, T, 2, 3 * H, device='cuda')
    out = torch.zeros(B, T, 2 * H, dtype=torch.float16, device='cuda')
    lib = _lib.lib()
    def run():
        _lib.check(lib.zs_gru_recurrence(gh.ptr(gx), gh.ptr(w), gh.ptr(b), B, T, H, gh.ptr(out), T, 2 * H, 0, 0, 0, 2, gh.stream()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(f'B={B} debug={os.environ.get("ZS_GRU_DEBUG","0")}: {e0.elapsed_time(e1) / 10 * 1000:.1f} us per call (incl. W pack), {e0.elapsed_time(e1) / 10 / T * 1000:.2f} us/step')
else:
    for B in (16, 64):
        for dbg in (0, 1, 2, 4, 7):
            env = dict(os.environ, ZS_GRU_DEBUG=str(dbg))
            subprocess.run([sys.executable, __file__, 'child', str(B)], env=env)
