"""Benchmark of the autoencoder hot path: spectrogram frames/s, encode + decode (resynthesis).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--segments S] [--impl ours|reference]

One "step" = one pass of Encoder -> one-hot bottleneck -> speaker-conditioned Decoder over S segments of
128 frames per GPU (enc_size 1024, emb_size 1024, 102 speakers, random-init weights, synthetic spectrograms).
Segments are independent, so ranks shard them with no collective ("weak" scaling: S per GPU is fixed).

  value  : frames/s with inputs already resident in HBM (device-timed with CUDA events, max over ranks)
  e2e    : the same work through the public API from pinned HOST buffers: per step the spectrograms, speaker
           ids and Gumbel noise are copied H2D and the decoded spectrograms + unit ids are copied D2H inside
           the timed region
  roofline: tensor-core roofline of the dominant kernel (the tcgen05 implicit-GEMM conv/linear kernel):
           algorithmic FLOPs of its launches / their CUDA-event time (per-launch events on the launch stream)
  cpu_baseline: the oracle (CPU fp32 restatement of the reference, all host cores) on a bounded sample
  --impl reference: times that CPU restatement as the reference arm (the reference itself is pure PyTorch and
           /root/reference does not exist on the GPU box; the oracle is pinned to it by tests/golden)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

# NCCL prints its version banner to STDOUT at any NCCL_DEBUG level from VERSION up: keep the JSON line alone on stdout
if os.environ.get('ZS_NCCL_DEBUG'):
    os.environ['NCCL_DEBUG'] = os.environ['ZS_NCCL_DEBUG']
else:
    os.environ.pop('NCCL_DEBUG', None)
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES = 128
ENC_SIZE, EMB_SIZE, N_SPK = 1024, 1024, 102
METRIC = 'spectrogram frames/s, AE encode+decode'
# algorithmic forward FLOPs per input frame, measured on the reference modules
MFLOP_PER_FRAME = 60.033


class stdout_to_stderr:
    """fd-level redirect: native libraries (NCCL's version banner) must not write into the JSON-only stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def init_distributed(dev):
    import torch.distributed as dist
    with stdout_to_stderr():
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()                       # communicator creation happens here at the latest


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--segments', type=int, default=960, help='128-frame segments per GPU per step')
    ap.add_argument('--micro-batch', type=int, default=960,
                    help='segments per library call: 960 x 128 frames = 480 column tiles, x 8 row tiles = 3840 tiles = 25.95 waves of 148 '
                         'CTAs on the 1024-channel layers (same on the 2048-channel up-convs), and 30 decoder-GRU clusters of 64 sequences = '
                         'exactly 2 waves of the 15 eight-CTA clusters that fit a B200')
    ap.add_argument('--e2e-micro-batch', type=int, default=960, help='segments per pipelined copy/compute stage (e2e)')
    ap.add_argument('--e2e-buffers', type=int, default=4, help='device buffer sets of the host-to-host pipeline')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--operand', default='fp16', choices=['fp16', 'bf16'])
    ap.add_argument('--cpu-sample', type=int, default=32, help='segments in the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='resynthesis', choices=['resynthesis', 'train'],
                    help="'train' = BASELINE config 4: pretrain_AE step, 32 segments x 128 frames per rank, NCCL gradient all-reduce")
    ap.add_argument('--train-batch', type=int, default=32)
    return ap.parse_args()


def config(args, n):
    return {'workload': f'encode->decode resynthesis, {args.segments} segments x {FRAMES} frames per GPU per step '
                        f'(micro-batches of {args.micro_batch}), enc_size {ENC_SIZE} one_hot, emb_size {EMB_SIZE}, '
                        f'{N_SPK} speakers',
            'segments_per_gpu': args.segments, 'frames_per_segment': FRAMES, 'micro_batch': args.micro_batch,
            'enc_size': ENC_SIZE, 'emb_size': EMB_SIZE, 'enc_mode': 'one_hot', 'parallelism': f'segment-sharded x{n}',
            'l2': 'inputs rotate over distinct batches totalling > 126 MB so no step re-reads its inputs from L2; '
                  'weights stay resident as in steady-state serving'}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle on host cores
# ------------------------------------------------------------------------------------------------
def cpu_run(n_seg, steps, warmup, check=None):
    """Times the oracle; with `check=(x, c, uniform, ids, spec)` from the CUDA path also returns parity numbers."""
    from zs_b200 import synthetic as syn
    from oracle import ae_oracle as orc
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    enc_sd = syn.encoder_state_dict(0, enc_size=ENC_SIZE, enc_mode='one_hot')
    dec_sd = syn.decoder_state_dict(0, c_in=ENC_SIZE, c_h=EMB_SIZE, c_a=N_SPK)
    x = syn.spectrogram_batch(n_seg, FRAMES, 0)
    c = syn.speaker_ids(n_seg, N_SPK, 0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            u = torch.rand(n_seg, 16, ENC_SIZE)            # the reference draws its noise inside forward
            act, _, _ = orc.encoder_forward(enc_sd, x, u)
            orc.decoder_forward(dec_sd, act, c)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        parity = None
        if check is not None:
            cx, cc, cu, ids, spec = check
            o_act, o_logits, o_ids = orc.encoder_forward(enc_sd, cx, cu)
            ours_act = torch.zeros_like(o_act).scatter_(1, ids.long().unsqueeze(1), 1.0)
            o_spec = orc.decoder_forward(dec_sd, ours_act, cc)     # same units -> decoder error only
            parity = {'unit_id_agreement_pct': 100.0 * (ids.long() == o_ids).float().mean().item(),
                      'spectrogram_rel_rms': ((spec - o_spec).norm() / o_spec.norm()).item(),
                      'spectrogram_max_abs': (spec - o_spec).abs().max().item(), 'segments_checked': int(cx.shape[0]),
                      'against': 'oracle/ae_oracle.py (CPU fp32), same weights, same Gumbel noise'}
    t = sum(times) / len(times)
    return n_seg * FRAMES / t, t, cores, parity


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import zs_b200  # noqa: F401
    n_seg = min(args.segments, args.cpu_sample)
    steps = max(1, min(args.steps, 3))
    warm = 1
    fps, t, cores, _ = cpu_run(n_seg, steps, warm)
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warm, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config(args, args.gpus),
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{n_seg} segments x {FRAMES} frames per step, {steps} timed steps, '
                                       'oracle/ae_oracle.py (torch fp32 CPU restatement pinned to the live reference)'},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# host placement: pinned buffers on the GPU's own NUMA node
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(index):
    """Restricts this process to the CPUs local to GPU `index` BEFORE the pinned host buffers are allocated, so that
    first-touch puts them on the GPU's NUMA node (with 8 ranks streaming ~55 GB/s each, remote-socket buffers halve the
    host-to-host rate).  Returns a short description, or None when the topology is not exposed."""
    try:
        bdf = subprocess.run(['nvidia-smi', '--query-gpu=pci.bus_id', '--format=csv,noheader', '-i', str(index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bdf.startswith('0000'):
            bdf = bdf[4:]                      # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        base = f'/sys/bus/pci/devices/{bdf}'
        node = int(open(base + '/numa_node').read())
        cpus = open(base + '/local_cpulist').read().strip()
        if node < 0 or not cpus:
            return None
        ids = set()
        for part in cpus.split(','):
            lo, _, hi = part.partition('-')
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        if not ids:
            return None
        os.sched_setaffinity(0, ids)
        return f'numa node {node}, {len(ids)} cpus'
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _nvml_loop(self):
        """Fast path: NVML queries take microseconds, so a 0.4 s timed region still gets dozens of samples."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))
        while not self._stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = get_reasons(h)
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            self.samples.append([str(sm), str(mx), f'{pw:.1f}'] + ['Active' if r & b else 'Not Active' for _, b in bits])
            self._stop.wait(0.01)

    def _loop(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            pass                                  # no NVML binding: fall back to polling nvidia-smi
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i',
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def init_distributed(dev):
    import torch.distributed as dist
    with stdout_to_stderr():
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()                       # communicator creation happens here at the latest


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--segments', type=int, default=960, help='128-frame segments per GPU per step')
    ap.add_argument('--micro-batch', type=int, default=960,
                    help='segments per library call: 960 x 128 frames = 480 column tiles, x 8 row tiles = 3840 tiles = 25.95 waves of 148 '
                         'CTAs on the 1024-channel layers (same on the 2048-channel up-convs), and 30 decoder-GRU clusters of 64 sequences = '
                         'exactly 2 waves of the 15 eight-CTA clusters that fit a B200')
    ap.add_argument('--e2e-micro-batch', type=int, default=960, help='segments per pipelined copy/compute stage (e2e)')
    ap.add_argument('--e2e-buffers', type=int, default=4, help='device buffer sets of the host-to-host pipeline')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--operand', default='fp16', choices=['fp16', 'bf16'])
    ap.add_argument('--cpu-sample', type=int, default=32, help='segments in the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='resynthesis', choices=['resynthesis', 'train'],
                    help="'train' = BASELINE config 4: pretrain_AE step, 32 segments x 128 frames per rank, NCCL gradient all-reduce")
    ap.add_argument('--train-batch', type=int, default=32)
    return ap.parse_args()


def config(args, n):
    return {'workload': f'encode->decode resynthesis, {args.segments} segments x {FRAMES} frames per GPU per step '
                        f'(micro-batches of {args.micro_batch}), enc_size {ENC_SIZE} one_hot, emb_size {EMB_SIZE}, '
                        f'{N_SPK} speakers',
            'segments_per_gpu': args.segments, 'frames_per_segment': FRAMES, 'micro_batch': args.micro_batch,
            'enc_size': ENC_SIZE, 'emb_size': EMB_SIZE, 'enc_mode': 'one_hot', 'parallelism': f'segment-sharded x{n}',
            'l2': 'inputs rotate over distinct batches totalling > 126 MB so no step re-reads its inputs from L2; '
                  'weights stay resident as in steady-state serving'}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle on host cores
# ------------------------------------------------------------------------------------------------
def cpu_run(n_seg, steps, warmup, check=None):
    """Times the oracle; with `check=(x, c, uniform, ids, spec)` from the CUDA path also returns parity numbers."""
    from zs_b200 import synthetic as syn
    from oracle import ae_oracle as orc
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    enc_sd = syn.encoder_state_dict(0, enc_size=ENC_SIZE, enc_mode='one_hot')
    dec_sd = syn.decoder_state_dict(0, c_in=ENC_SIZE, c_h=EMB_SIZE, c_a=N_SPK)
    x = syn.spectrogram_batch(n_seg, FRAMES, 0)
    c = syn.speaker_ids(n_seg, N_SPK, 0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            u = torch.rand(n_seg, 16, ENC_SIZE)            # the reference draws its noise inside forward
            act, _, _ = orc.encoder_forward(enc_sd, x, u)
            orc.decoder_forward(dec_sd, act, c)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        parity = None
        if check is not None:
            cx, cc, cu, ids, spec = check
            o_act, o_logits, o_ids = orc.encoder_forward(enc_sd, cx, cu)
            ours_act = torch.zeros_like(o_act).scatter_(1, ids.long().unsqueeze(1), 1.0)
            o_spec = orc.decoder_forward(dec_sd, ours_act, cc)     # same units -> decoder error only
            parity = {'unit_id_agreement_pct': 100.0 * (ids.long() == o_ids).float().mean().item(),
                      'spectrogram_rel_rms': ((spec - o_spec).norm() / o_spec.norm()).item(),
                      'spectrogram_max_abs': (spec - o_spec).abs().max().item(), 'segments_checked': int(cx.shape[0]),
                      'against': 'oracle/ae_oracle.py (CPU fp32), same weights, same Gumbel noise'}
    t = sum(times) / len(times)
    return n_seg * FRAMES / t, t, cores, parity


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import zs_b200  # noqa: F401
    n_seg = min(args.segments, args.cpu_sample)
    steps = max(1, min(args.steps, 3))
    warm = 1
    fps, t, cores, _ = cpu_run(n_seg, steps, warm)
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warm, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config(args, args.gpus),
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{n_seg} segments x {FRAMES} frames per step, {steps} timed steps, '
                                       'oracle/ae_oracle.py (torch fp32 CPU restatement pinned to the live reference)'},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# host placement: pinned buffers on the GPU's own NUMA node
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(index):
    """Restricts this process to the CPUs local to GPU `index` BEFORE the pinned host buffers are allocated, so that
    first-touch puts them on the GPU's NUMA node (with 8 ranks streaming ~55 GB/s each, remote-socket buffers halve the
    host-to-host rate).  Returns a short description, or None when the topology is not exposed."""
    try:
        bdf = subprocess.run(['nvidia-smi', '--query-gpu=pci.bus_id', '--format=csv,noheader', '-i', str(index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bdf.startswith('0000'):
            bdf = bdf[4:]                      # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        base = f'/sys/bus/pci/devices/{bdf}'
        node = int(open(base + '/numa_node').read())
        cpus = open(base + '/local_cpulist').read().strip()
        if node < 0 or not cpus:
            return None
        ids = set()
        for part in cpus.split(','):
            lo, _, hi = part.partition('-')
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        if not ids:
            return None
        os.sched_setaffinity(0, ids)
        return f'numa node {node}, {len(ids)} cpus'
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i',
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace('.', '').isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace('.', '').isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import zs_b200  # noqa: F401
    from zs_b200 import _lib, synthetic as syn
    from zs_b200.model import Decoder, Encoder
    from zs_b200.frontend import StreamingResynthesizer

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        init_distributed(dev)

    lib = _lib.lib()
    enc = Encoder(ns=0.01, dp=0.5, enc_size=ENC_SIZE, seg_len=128, enc_mode='one_hot')
    dec = Decoder(ns=0.01, c_in=ENC_SIZE, c_h=EMB_SIZE, c_a=N_SPK, seg_len=128)
    enc.load_state_dict(syn.encoder_state_dict(0, enc_size=ENC_SIZE, enc_mode='one_hot'))
    dec.load_state_dict(syn.decoder_state_dict(0, c_in=ENC_SIZE, c_h=EMB_SIZE, c_a=N_SPK))
    enc.operand = dec.operand = args.operand
    enc.to(dev).eval()
    dec.to(dev).eval()

    S, MB = args.segments, args.micro_batch
    # distinct input sets, rotated so a step never finds its inputs in L2 (126 MB)
    bytes_per_set = S * 513 * FRAMES * 4
    n_sets = max(2, min(8, -(-140_000_000 // bytes_per_set)))
    xs_host = [syn.spectrogram_batch(S, FRAMES, 100 * rank + i).pin_memory() for i in range(n_sets)]
    cs_host = [syn.speaker_ids(S, N_SPK, 100 * rank + i).pin_memory() for i in range(n_sets)]
    # Gumbel noise value (model/model.py:95-98); generated once per set on the host, an INPUT of the path
    nz_host = [(-torch.log(-torch.log(syn.gumbel_uniform((S, 16, ENC_SIZE), 100 * rank + i) + 1e-20) + 1e-20)).pin_memory()
               for i in range(n_sets)]
    xs = [t.to(dev) for t in xs_host]
    cs = [t.to(dev) for t in cs_host]
    nz = [t.to(dev) for t in nz_host]
    spec_out = torch.empty(S, 513, FRAMES, device=dev)
    ids_out = torch.empty(S, 16, dtype=torch.int32, device=dev)
    spec_host = torch.empty(S, 513, FRAMES).pin_memory()
    ids_host = torch.empty(S, 16, dtype=torch.int32).pin_memory()

    def step_device(i):
        k = i % n_sets
        for s0 in range(0, S, MB):
            s1 = min(S, s0 + MB)
            act, _, ids = enc.encode(xs[k][s0:s1], nz[k][s0:s1])
            dec.decode(None, cs[k][s0:s1], unit_ids=ids, out=spec_out[s0:s1])
            ids_out[s0:s1] = ids

    streamer = StreamingResynthesizer(enc, dec, micro_batch=min(args.e2e_micro_batch, S), n_buffers=args.e2e_buffers, device=dev)

    # public API: pinned host spectrograms/speakers/noise in, pinned host spectrograms/units out.  Steps are issued
    # back to back as a streaming server would (run_async): the upload of step i+1 overlaps the compute and download
    # of step i; every step's results land in host memory inside the timed region (two alternating output buffers).
    out_hosts = [(spec_host, ids_host), (torch.empty(S, 513, FRAMES).pin_memory(), torch.empty(S, 16, dtype=torch.int32).pin_memory())]
    pending = []

    def step_e2e(i):
        k = i % n_sets
        sh, ih = out_hosts[i % 2]
        if len(pending) >= 2:
            pending.pop(0).synchronize()        # this output buffer's previous results are complete (and consumable)
        pending.append(streamer.run_async(xs_host[k], cs_host[k], sh, ih, nz_host[k]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, finish=None):
        for i in range(warmup):
            fn(i)
        if finish:
            finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        if finish:
            finish()            # the timing stream waits for every step's download before the closing event
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / 1e3, wall

    W = max(args.warmup, 3)
    with ClockSampler(local) as clk:
        t_dev, _ = timed(step_device, args.steps, W)
    clocks = clk.summary()
    # launch count of one step
    cnt = (C.c_longlong * 3)()
    lib.zs_profile_begin()
    ms3, fl3 = (C.c_double * 3)(), (C.c_double * 3)()
    step_device(0)
    _lib.check(lib.zs_profile_end(ms3, fl3, cnt))
    launches_per_step = int(sum(cnt))
    gemm_ms, gemm_flops, gemm_launches = ms3[0], fl3[0], int(cnt[0])
    gru_ms, other_ms = ms3[1], ms3[2]
    # per-launch-event profile over a few more steps for a stable roofline number
    lib.zs_profile_begin()
    for i in range(3):
        step_device(i + 1)
    _lib.check(lib.zs_profile_end(ms3, fl3, cnt))
    gemm_ms, gemm_flops, gemm_launches = ms3[0] / 3, fl3[0] / 3, int(cnt[0]) // 3
    gru_ms, other_ms = ms3[1] / 3, ms3[2] / 3

    def finish_e2e():
        while pending:
            torch.cuda.current_stream().wait_event(pending.pop(0))

    t_e2e, _ = timed(step_e2e, args.steps, W, finish_e2e)

    frames_per_step = S * FRAMES * world
    value = frames_per_step * args.steps / t_dev
    e2e = frames_per_step * args.steps / t_e2e

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = peaks.get('bf16_tflops_sustained', 1400.0)   # kernel timed inside a long step -> sustained figure
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (fp16 and bf16 share the kind::f16 rate)' if peaks \
            else 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        # DRAM traffic of the GEMM launches from the committed ncu --set full capture (same micro-batch size only)
        traffic, traffic_note = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'r01_gemm_traffic.json')))
            if MB >= 444:     # saturating calls: the traffic is activation traffic and scales with the segments
                n_cap = S / tr['micro_batch']
                traffic = (tr['dram_read_bytes'] + tr['dram_write_bytes']) * n_cap
                traffic_note = (f"bytes per step over all {gemm_launches} GEMM launches = {n_cap:.3f} x "
                                f"{(tr['dram_read_bytes'] + tr['dram_write_bytes']) / 1e9:.2f} GB (ncu --set full capture of one {tr['micro_batch']}-segment call, profiles/r01_ncu_full_conv_gemm_mb960.csv)")
        except Exception:
            pass
        line = {
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': W,
            'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': f'{args.operand} operands, f32 accumulate', 'data': 'synthetic', 'config': config(args, world),
            'clocks': clocks, 'gpu_launches': launches_per_step * args.steps,
            'e2e': {'value': e2e, 'unit': 'frames/s',
                    'h2d_bytes_per_step': S * (513 * FRAMES * 4 + 8 + 16 * ENC_SIZE * 4),
                    'd2h_bytes_per_step': S * (513 * FRAMES * 4 + 16 * 4), 'ms_per_step': t_e2e / args.steps * 1e3,
                    'host_placement': numa,
                    'api': 'StreamingResynthesizer.run_async: pinned host in -> pinned host out, steps issued back to back '
                           '(upload of step i+1 under compute/download of step i)'},
            'roofline': {'bound': 'tensor', 'kernel': 'conv_gemm_kernel (tcgen05 implicit GEMM, all conv/linear layers)',
                         'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak if peak else None,
                         'traffic': traffic, 'traffic_note': traffic_note, 'peak_source': peak_src, 'launches_per_step': gemm_launches,
                         'algorithmic_gflop_per_step': gemm_flops / 1e9, 'kernel_ms_per_step': gemm_ms,
                         'share_of_step': gemm_ms / (gemm_ms + gru_ms + other_ms) if gemm_ms else None,
                         'gru_ms_per_step': gru_ms, 'other_ms_per_step': other_ms},
            'whole_path_tflops': value / world * MFLOP_PER_FRAME * 1e6 / 1e12,
        }
        # large-batch result must equal the 32-segment-batch result bit for bit (segments are independent)
        step_device(0)
        a32, _, i32 = enc.encode(xs[0][:32], nz[0][:32])
        s32 = dec.decode(None, cs[0][:32], unit_ids=i32)
        line['self_check'] = {'batch_invariant': bool(torch.equal(s32, spec_out[:32]) and torch.equal(i32, ids_out[:32]))}
        if not args.no_cpu_baseline and world == 1:
            n_chk = min(S, 8)
            u_chk = syn.gumbel_uniform((S, 16, ENC_SIZE), 100 * rank)[:n_chk]
            check = (xs_host[0][:n_chk].clone(), cs_host[0][:n_chk].clone(), u_chk, ids_out[:n_chk].cpu(), spec_out[:n_chk].cpu())
            fps, t, cores, parity = cpu_run(min(S, args.cpu_sample), 2, 1, check)
            line['parity'] = parity
            line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'{min(S, args.cpu_sample)} segments x {FRAMES} frames, 2 timed passes of '
                                              'oracle/ae_oracle.py (torch fp32, all host cores)'}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """BASELINE config 4 (not the headline line): one pretrain_AE iteration per step (trainer.py:321-332), B segments
    per rank, gradients all-reduced over NCCL when N > 1.  value: batch resident in HBM; e2e: batch from pinned host
    memory every step and the loss read back on the host every step (trainer.py:336 does `.item()` every iteration)."""
    import torch.distributed as dist
    import zs_b200  # noqa: F401
    from zs_b200 import _lib, synthetic as syn, train as zt
    from zs_b200.model import Decoder, Encoder
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', 1), ('RANK', 0), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        init_distributed(dev)
    B = args.train_batch
    enc = Encoder(ns=0.01, dp=0.5, enc_size=ENC_SIZE, seg_len=128, enc_mode='one_hot')
    dec = Decoder(ns=0.01, c_in=ENC_SIZE, c_h=EMB_SIZE, c_a=N_SPK, seg_len=128)
    enc.load_state_dict(syn.encoder_state_dict(0, enc_size=ENC_SIZE, enc_mode='one_hot'))
    dec.load_state_dict(syn.decoder_state_dict(0, c_in=ENC_SIZE, c_h=EMB_SIZE, c_a=N_SPK))
    enc.to(dev).train()
    dec.to(dev).train()
    step = zt.PretrainAE(enc, dec)
    n_sets = 8
    xs_host = [syn.spectrogram_batch(B, FRAMES, 100 * rank + i).pin_memory() for i in range(n_sets)]
    cs_host = [syn.speaker_ids(B, N_SPK, 100 * rank + i).pin_memory() for i in range(n_sets)]
    xs, cs = [t.to(dev) for t in xs_host], [t.to(dev) for t in cs_host]
    x_dev, c_dev = torch.empty_like(xs[0]), torch.empty_like(cs[0])
    loss_host = torch.zeros(1).pin_memory()

    def step_device(i):
        step.step(xs[i % n_sets], cs[i % n_sets])

    def step_e2e(i):
        x_dev.copy_(xs_host[i % n_sets], non_blocking=True)
        c_dev.copy_(cs_host[i % n_sets], non_blocking=True)
        loss_host.copy_(step.step(x_dev, c_dev), non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the reference reads loss.item() every iteration

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / 1e3

    W = max(args.warmup, 5)        # two eager steps + the graph capture come first
    with ClockSampler(local) as clk:
        t_dev = timed(step_device, args.steps, W)
    clocks = clk.summary()
    t_e2e = timed(step_e2e, args.steps, W)
    # kernel classes of one eager iteration (a graph replay launches the same kernels without passing through the
    # library's launch accounting)
    lib = _lib.lib()
    eager = zt.PretrainAE(enc, dec, process_group=None, use_graph=False) if world == 1 else None
    ms3, fl3, cnt = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
    if eager is not None:
        eager.step(xs[0], cs[0])
        lib.zs_profile_begin()
        eager.step(xs[1], cs[1])
        _lib.check(lib.zs_profile_end(ms3, fl3, cnt))
    frames = B * FRAMES * world
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = peaks.get('bf16_tflops_sustained', 1400.0)
        ach = fl3[0] / (ms3[0] * 1e-3) / 1e12 if ms3[0] > 0 else None
        line = {'metric': 'spectrogram frames/s, pretrain_AE step (fwd + bwd + clip + Adam)', 'value': frames * args.steps / t_dev,
                'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': W, 'ms_per_step': t_dev / args.steps * 1e3,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp16 operands, f32 accumulate / master weights',
                'data': 'synthetic',
                'config': {'workload': f'train_ae step, {B} segments x {FRAMES} frames per rank, enc_size {ENC_SIZE} one_hot, dropout 0.5, '
                                       f'Adam(1e-4, (0.5, 0.9)), per-net clip 5; data-parallel x{world}, one 221 MB fp32 gradient all-reduce per step',
                           'l2': 'inputs rotate over 8 distinct batches; weights (110 MB fp16 + 221 MB fp32) exceed what stays in L2 with the activations'},
                'clocks': clocks, 'gpu_launches': int(sum(cnt)) * args.steps if eager is not None else None,
                'e2e': {'value': frames * args.steps / t_e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': B * (513 * FRAMES * 4 + 8),
                        'd2h_bytes_per_step': 4, 'ms_per_step': t_e2e / args.steps * 1e3},
                'roofline': {'bound': 'tensor', 'kernel': 'conv_gemm_kernel + wgrad_gemm_kernel (tcgen05), one eager iteration',
                             'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak if ach else None, 'traffic': None,
                             'kernel_ms_per_step': ms3[0], 'gru_ms_per_step': ms3[1], 'other_ms_per_step': ms3[2],
                             'launches_per_step': int(sum(cnt))},
                'loss': float(step.loss.item()), 'skipped_steps': step.n_skipped}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == 'train' and args.impl != 'reference':
        return run_train(args)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
