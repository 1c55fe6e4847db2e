"""Host<->device copy rates from pinned memory, alone or on all ranks at once (torchrun): the ceiling of bench.py's e2e."""
import os, time
import torch
import torch.distributed as dist
world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', 1), ('RANK', 0), ('LOCAL_RANK', 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
x = torch.empty(256 * 1024 * 1024 // 4).pin_memory()
d = torch.empty_like(x, device='cuda')
y = torch.empty_like(x).pin_memory(); e = torch.empty_like(d)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(e, non_blocking=True)
res = []
for name, fn, gib in (('H2D', lambda: d.copy_(x, non_blocking=True), 0.25), ('D2H', lambda: x.copy_(d, non_blocking=True), 0.25),
                      ('both directions', both, 0.5)):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): fn()
    torch.cuda.synchronize()
    res.append(f'{name} {8 * gib / (time.perf_counter() - t0):.1f} GiB/s')
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        for r, o in enumerate(out): print(f'rank {r}:', ' | '.join(o))
    dist.destroy_process_group()
else:
    print(' | '.join(res))
