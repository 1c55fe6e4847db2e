"""Benchmark of the autoencoder hot path: spectrogram frames/s, encode + decode (resynthesis).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload resynthesis|train]

One "step" = one pass of Encoder -> one-hot bottleneck -> speaker-conditioned Decoder over `--calls-per-step` x
`--segments` segments of 128 frames per GPU (default 24 x 960 = 23 040 segments = 2.95 M frames; enc_size 1024,
emb_size 1024, 102 speakers, random-init weights, synthetic spectrograms), issued as library calls of 960 segments.  A step
is that large so that the default 20 timed steps keep the GPU under load for > 3 s: the clocks settle in the sustained
regime MEASURED_PEAKS.json's `bf16_tflops_sustained` was taken in.  Segments are independent, so ranks shard them with
no collective ("weak" scaling: segments per GPU are fixed).

  value   : frames/s with inputs already resident in HBM (CUDA events, max over ranks)
  e2e     : the same work through the public API (`StreamingResynthesizer.run_async`) from pinned HOST buffers: every call's
            spectrograms and speaker ids are copied H2D and the decoded spectrograms + unit ids D2H inside the timed region
            (headline mode: fp16 features in the (T, 513) layout of the HDF5 files + Gumbel noise drawn on the device;
            `e2e_modes` also carries the reference-exact mode: fp32 (513, T) features + CPU-generator noise uploaded)
  roofline: tensor-core roofline of the dominant kernel (the tcgen05 implicit-GEMM conv/linear kernel): algorithmic FLOPs of
            its launches / their CUDA-event time, against BOTH measured peaks (burst and sustained)
  parity  : unit-id agreement with the oracle over >= 10 000 unit frames; decoded-spectrogram error on a sample
  configs : BASELINE.json configs 2 and 5 and the B = 32 resynthesis, each timed on its own
  cuda_eager_baseline: the same model as stock torch modules (cuDNN TF32 convs, cuBLAS, cuDNN GRU) on this GPU
  cpu_baseline / cpu_baseline_b1: the oracle (CPU fp32 restatement of the reference) on the host cores, batched and in the
            reference's own one-chunk-per-call pattern
  train   : BASELINE config 4, one pretrain_AE iteration per step with the NCCL gradient all-reduce at N > 1
  --impl reference: times the CPU restatement as the reference arm (the reference itself is pure PyTorch and
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES = 128
ENC_SIZE, EMB_SIZE, N_SPK = 1024, 1024, 102
METRIC = 'spectrogram frames/s, AE encode+decode'
# algorithmic forward FLOPs per input frame, measured on the reference modules
MFLOP_PER_FRAME = 60.033


class stdout_to_stderr:
    """fd-level redirect: native libraries (NCCL's version banner, NCCL_DEBUG=INFO lines) must not write into the
    JSON-only stdout - they go to stderr, where the driver reads the communicator log."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def emit(line):
    """The ONE JSON line, written to the real stdout (fd kept aside while fd 1 points at stderr for native code)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


_REAL_STDOUT = os.dup(1)


def init_distributed(dev):
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)
    dist.barrier()                       # communicator creation happens here at the latest


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--segments', type=int, default=960,
                    help='128-frame segments per library call: 960 x 128 frames = 480 column tiles, x 8 row tiles = 3840 tiles = 25.95 '
                         'waves of 148 CTAs on the 1024-channel layers (same on the 2048-channel up-convs), and 30 decoder-GRU clusters '
                         'of 64 sequences = exactly 2 waves of the 15 eight-CTA clusters that fit a B200')
    ap.add_argument('--calls-per-step', type=int, default=24, help='library calls per step (per GPU); 24 keeps 20 timed steps above 3 s at 16-17 M frames/s')
    ap.add_argument('--e2e-buffers', type=int, default=4, help='device buffer sets (= calls in flight) of the host-to-host pipeline')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--operand', default='fp16', choices=['fp16', 'bf16'])
    ap.add_argument('--cpu-sample', type=int, default=32, help='segments in the CPU baseline sample')
    ap.add_argument('--parity-segments', type=int, default=960, help='segments of the unit-id agreement check (16 unit frames each)')
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip every CPU leg (parity, cpu baselines)')
    ap.add_argument('--no-extras', action='store_true', help='headline numbers only: skip configs / eager baseline / train / sharded sub-records')
    ap.add_argument('--workload', default='resynthesis', choices=['resynthesis', 'train'],
                    help="'train' = BASELINE config 4 as the headline line: pretrain_AE step, 32 segments x 128 frames per rank")
    ap.add_argument('--train-batch', type=int, default=32)
    return ap.parse_args()


def config(args, n):
    S = args.segments * args.calls_per_step
    return {'workload': f'encode->decode resynthesis, {S} segments x {FRAMES} frames per GPU per step in {args.calls_per_step} library calls '
                        f'of {args.segments} segments, enc_size {ENC_SIZE} one_hot, emb_size {EMB_SIZE}, {N_SPK} speakers '
                        '(BASELINE.json configs[2]; the saturating per-GPU batch SURVEY 8d names)',
            'segments_per_gpu_per_step': S, 'segments_per_call': args.segments, 'calls_per_step': args.calls_per_step,
            'frames_per_segment': FRAMES, 'enc_size': ENC_SIZE, 'emb_size': EMB_SIZE, 'enc_mode': 'one_hot',
            'parallelism': f'segment-sharded x{n}, no collective',
            'l2': 'calls rotate over 4 distinct input sets of 252 MB each (> 126 MB L2), so no call finds its inputs in L2; '
                  'weights stay resident as in steady-state serving'}


# ------------------------------------------------------------------------------------------------
# CPU legs: the oracle on host cores (reference arm, cpu_baseline, parity)
# ------------------------------------------------------------------------------------------------
def _weights():
    from zs_b200 import synthetic as syn
    return (syn.encoder_state_dict(0, enc_size=ENC_SIZE, enc_mode='one_hot'),
            syn.decoder_state_dict(0, c_in=ENC_SIZE, c_h=EMB_SIZE, c_a=N_SPK))


def cpu_time(n_seg, steps, warmup, batch1=False):
    """Times the oracle's encode -> decode over `n_seg` segments per step on all host cores.  `batch1`: the reference's
    own call pattern - one model call per 128-frame chunk (convert.py:70-83, 154-165)."""
    from zs_b200 import synthetic as syn
    from oracle import ae_oracle as orc
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    enc_sd, dec_sd = _weights()
    x = syn.spectrogram_batch(n_seg, FRAMES, 0)
    c = syn.speaker_ids(n_seg, N_SPK, 0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            if batch1:
                for s in range(n_seg):
                    u = torch.rand(1, 16, ENC_SIZE)            # the reference draws its noise inside forward
                    act, _, _ = orc.encoder_forward(enc_sd, x[s:s + 1], u)
                    orc.decoder_forward(dec_sd, act, c[s:s + 1])
            else:
                u = torch.rand(n_seg, 16, ENC_SIZE)
                act, _, _ = orc.encoder_forward(enc_sd, x, u)
                orc.decoder_forward(dec_sd, act, c)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return n_seg * FRAMES / t, t, cores


def cpu_parity(x, c, uniform, ids, spec_x, spec_c, spec_ids, spec):
    """unit-id agreement of the CUDA path with the oracle on (x, uniform) -> ids [all given segments], and the decoder's
    spectrogram error on (spec_ids, spec_c) -> spec [a sample].  Same weights, same Gumbel noise."""
    from oracle import ae_oracle as orc
    torch.set_num_threads(os.cpu_count())
    enc_sd, dec_sd = _weights()
    agree = n = 0
    rel_l = []
    with torch.no_grad():
        for s0 in range(0, x.shape[0], 96):
            s1 = min(x.shape[0], s0 + 96)
            _, o_logits, o_ids = orc.encoder_forward(enc_sd, x[s0:s1].float(), uniform[s0:s1])
            agree += int((ids[s0:s1].long() == o_ids).sum())
            n += o_ids.numel()
        ours_act = torch.zeros(spec_ids.shape[0], ENC_SIZE, spec_ids.shape[1]).scatter_(1, spec_ids.long().unsqueeze(1), 1.0)
        o_spec = orc.decoder_forward(dec_sd, ours_act, spec_c)     # same units -> decoder error only
    return {'unit_id_agreement_pct': 100.0 * agree / n, 'unit_frames_checked': n, 'segments_checked': int(x.shape[0]),
            'spectrogram_rel_rms': ((spec - o_spec).norm() / o_spec.norm()).item(),
            'spectrogram_max_abs': (spec - o_spec).abs().max().item(), 'spectrogram_segments_checked': int(spec.shape[0]),
            'tolerance': 'unit ids >= 95 % agreement; spectrogram rel-RMS <= 1e-2 (fp16 operands, fp32 accumulate)',
            'against': 'oracle/ae_oracle.py (CPU fp32, pinned to the live reference by tests/golden), same weights, same Gumbel noise'}


def run_reference(args):
    """Reference arm: the CPU restatement of the reference (`kind: port`; the reference is pure PyTorch, cannot be
    pip-installed and /root/reference does not travel to the GPU box) on all host cores, honouring --steps/--warmup.  A step
    here is a BOUNDED SAMPLE of the other arm's step (32 of its segments) - the metric is a rate, so the ratio stands."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import zs_b200  # noqa: F401
    n_seg = min(args.segments, args.cpu_sample)
    steps, warm = max(1, args.steps), max(1, args.warmup)
    fps, t, cores = cpu_time(n_seg, steps, warm)
    cfg = config(args, args.gpus)
    cfg['reference_sample'] = (f'each of the {steps} timed steps (after {warm} warm-up steps) ran {n_seg} segments x {FRAMES} frames - a bounded '
                               f'sample of the {args.segments * args.calls_per_step}-segment step - through oracle/ae_oracle.py (torch fp32 CPU port of '
                               f'the reference, pinned to the live reference by tests/golden) on {cores} host threads')
    emit({'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': steps,
          'warmup': warm, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
          'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
          'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                           'sample': f'{n_seg} segments x {FRAMES} frames per step, {steps} timed steps, oracle/ae_oracle.py'},
          'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})


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
# clocks: NVML polled every 10 ms during the timed region (nvidia-smi polling as the fallback)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    BITS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t, self.source = index, [], threading.Event(), None, None

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self.source = 'nvml'
        while not self._stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = get_reasons(h)
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            self.samples.append((float(sm), float(mx), pw, [name for name, b in self.BITS if r & b]))
            self._stop.wait(0.01)

    def _smi_loop(self):
        self.source = 'nvidia-smi'
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [v.strip() for v in out.split(',')]
                if len(f) >= 7:
                    self.samples.append((float(f[0]), float(f[1]), float(f[2]) if f[2].replace('.', '').isdigit() else 0.0,
                                         [n for (n, _), v in zip(self.BITS, f[3:7]) if v.lower().startswith('active')]))
            except Exception:
                pass
            self._stop.wait(0.1)

    def _loop(self):
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted({r for s in self.samples for r in s[3]})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_min_mhz': sm[0] if sm else None,
                'sm_max_mhz': max((s[1] for s in self.samples), default=None), 'reasons': reasons,
                'power_w_max': max((s[2] for s in self.samples), default=None), 'samples': len(self.samples), 'source': self.source}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


def event_time(fn, n, warmup, dev):
    """ms per call of `fn(i)` over n calls after `warmup` calls, CUDA events on the current stream."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / n


def build_models(dev, operand, enc_size=ENC_SIZE, generator=False):
    from zs_b200 import synthetic as syn
    from zs_b200.model import Decoder, Encoder
    enc = Encoder(ns=0.01, dp=0.5, enc_size=enc_size, seg_len=128, enc_mode='one_hot')
    dec = Decoder(ns=0.01, c_in=enc_size, c_h=EMB_SIZE, c_a=N_SPK, seg_len=128)
    enc.load_state_dict(syn.encoder_state_dict(0, enc_size=enc_size, enc_mode='one_hot'))
    dec.load_state_dict(syn.decoder_state_dict(0, c_in=enc_size, c_h=EMB_SIZE, c_a=N_SPK))
    mods = [enc, dec]
    if generator:
        gen = Decoder(ns=0.01, c_in=enc_size, c_h=EMB_SIZE, c_a=2, seg_len=128)
        gen.load_state_dict(syn.decoder_state_dict(1, c_in=enc_size, c_h=EMB_SIZE, c_a=2))
        mods.append(gen)
    for m in mods:
        m.operand = operand
        m.to(dev).eval()
    return mods


# ------------------------------------------------------------------------------------------------
# sub-records (N = 1): BASELINE configs 2 / 5, B = 32 resynthesis, eager-CUDA bar, HBM-bound kernels
# ------------------------------------------------------------------------------------------------
def sub_small_batches(enc, dec, dev):
    """BASELINE configs[1] (32 segments, encode only = --test_encode) and the same batch through encode -> decode.
    Device-resident: 16 rotating input sets (134 MB > L2); host-to-host: pinned x / speakers / noise up, units (+ spectrograms)
    down, one synchronising call at a time - the latency a caller of the module API sees."""
    from zs_b200 import synthetic as syn
    from zs_b200.model import gumbel_from_uniform
    B, n_sets = 32, 16
    xs_h = [syn.spectrogram_batch(B, FRAMES, 500 + i).pin_memory() for i in range(n_sets)]
    cs_h = [syn.speaker_ids(B, N_SPK, 500 + i).pin_memory() for i in range(n_sets)]
    nz_h = [gumbel_from_uniform(syn.gumbel_uniform((B, 16, ENC_SIZE), 500 + i)).pin_memory() for i in range(n_sets)]
    xs, cs, nz = ([t.to(dev) for t in l] for l in (xs_h, cs_h, nz_h))
    ids_h = torch.empty(B, 16, dtype=torch.int32).pin_memory()
    spec_h = torch.empty(B, 513, FRAMES).pin_memory()
    out = {}

    def enc_dev(i):
        enc.encode(xs[i % n_sets], nz[i % n_sets], want_act=False, want_logits=False)

    def both_dev(i):
        _, _, ids = enc.encode(xs[i % n_sets], nz[i % n_sets], want_act=False, want_logits=False)
        dec.decode(None, cs[i % n_sets], unit_ids=ids)

    def enc_h2h(i):
        k = i % n_sets
        _, _, ids = enc.encode(xs_h[k].to(dev, non_blocking=True), nz_h[k].to(dev, non_blocking=True), want_act=False, want_logits=False)
        ids_h.copy_(ids, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def both_h2h(i):
        k = i % n_sets
        _, _, ids = enc.encode(xs_h[k].to(dev, non_blocking=True), nz_h[k].to(dev, non_blocking=True), want_act=False, want_logits=False)
        spec = dec.decode(None, cs_h[k].to(dev, non_blocking=True), unit_ids=ids)
        ids_h.copy_(ids, non_blocking=True)
        spec_h.copy_(spec, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for name, f_dev, f_h2h, what in (('config2_encode_b32', enc_dev, enc_h2h, 'Encoder -> unit ids (--test_encode), BASELINE configs[1]'),
                                      ('resynthesis_b32', both_dev, both_h2h, 'Encoder -> Decoder, 32 segments')):
        ms = event_time(f_dev, 200, 10, dev)
        ms_h = event_time(f_h2h, 100, 5, dev)
        out[name] = {'workload': f'{what}: {B} segments x {FRAMES} frames per call, enc_size {ENC_SIZE}',
                     'ms_per_call': ms, 'frames_per_s': B * FRAMES / ms * 1e3,
                     'host_to_host_ms_per_call': ms_h, 'host_to_host_frames_per_s': B * FRAMES / ms_h * 1e3,
                     'note': 'one call at a time (latency); 16 rotating input sets'}
    return out


def sub_config5(dev, operand):
    """BASELINE configs[4]: AE + TTS patcher on 2000-frame utterances, enc_size 512 vs 1024.  Through the public driver
    API (`AutoencoderPath.convert_utterances`, enc_only=False, g_mode targeted: Encoder + Decoder + Generator(c_a = 2),
    convert.py:128-168 chunking: 14 chunks of 128 + one of 207 per utterance), numpy in -> numpy out."""
    import numpy as np
    from zs_b200 import _lib
    from zs_b200.frontend import AutoencoderPath
    lib = _lib.lib()
    out = {}
    n_utt, L = 64, 2000
    rng = np.random.Generator(np.random.PCG64(77))
    specs = [np.clip(rng.random((L, 513), dtype=np.float32), 1e-8, 1.0) for _ in range(n_utt)]
    spk = [100 + (i & 1) for i in range(n_utt)]
    for enc_size in (512, 1024):
        enc, dec, gen = build_models(dev, operand, enc_size, generator=True)
        path = AutoencoderPath(enc, dec, gen, g_mode='targeted', device=dev)
        path.convert_utterances(specs, spk, enc_only=False, noise_seed=1, as_ids=True)      # warm-up: packing, workspaces, pinned staging
        torch.cuda.synchronize(dev)
        ms3, fl3, cnt = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
        lib.zs_profile_begin()
        t0 = time.perf_counter()
        res, units = path.convert_utterances(specs, spk, enc_only=False, noise_seed=1, as_ids=True)
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        _lib.check(lib.zs_profile_end(ms3, fl3, cnt))
        frames_in = n_utt * L
        assert res[0].shape == (14 * 128 + 208, 513) and units[0].shape == (14 * 16 + 26,)
        out[f'enc_size_{enc_size}'] = {
            'api_frames_per_s': frames_in / wall, 'api_ms': wall * 1e3, 'kernel_ms': sum(ms3), 'kernel_frames_per_s': frames_in / sum(ms3) * 1e3,
            'gemm_tflops': fl3[0] / (ms3[0] * 1e-3) / 1e12 if ms3[0] else None, 'launches': int(sum(cnt))}
        del enc, dec, gen, path
    out['workload'] = (f'{n_utt} utterances x {L} frames -> {n_utt * 15} chunks (14 x 128 + 1 x 207 frames each), Encoder + Decoder + '
                       "Generator 'targeted' (x_dec += G(enc, c - 100)), AutoencoderPath.convert_utterances numpy -> numpy; "
                       'api_* = wall clock incl. host chunking/stacking/concatenation, kernel_* = CUDA-event sum of the launches')
    return out


def sub_eager(dev):
    """The 'existing Blackwell path': the same model as stock torch modules on this GPU (oracle/torch_modules.py - cuDNN
    convolutions in TF32 as torch defaults to, cuBLAS fp32 linears, cuDNN GRU, one ATen launch per elementwise op; Gumbel
    noise resident on the device).  Same weights; its unit ids are checked against ours."""
    from zs_b200 import synthetic as syn
    from zs_b200.model import gumbel_from_uniform
    from oracle.torch_modules import TorchDecoder, TorchEncoder
    enc_sd, dec_sd = _weights()
    te = TorchEncoder(enc_size=ENC_SIZE).to(dev).eval()
    td = TorchDecoder(c_in=ENC_SIZE, c_h=EMB_SIZE, c_a=N_SPK).to(dev).eval()
    te.load_state_dict(enc_sd)
    td.load_state_dict(dec_sd)
    out = {'what': 'torch eager fp32 modules, cuDNN TF32 convs (torch default), device-resident inputs, CUDA events',
           'allow_tf32': {'cudnn': bool(torch.backends.cudnn.allow_tf32), 'matmul': bool(torch.backends.cuda.matmul.allow_tf32)}}
    with torch.no_grad():
        for B, n_sets, n in ((32, 16, 30), (960, 2, 4)):
            xs = [syn.spectrogram_batch(B, FRAMES, 700 + i).to(dev) for i in range(n_sets)]
            cs = [syn.speaker_ids(B, N_SPK, 700 + i).to(dev) for i in range(n_sets)]
            nz = [gumbel_from_uniform(syn.gumbel_uniform((B, 16, ENC_SIZE), 700 + i)).to(dev) for i in range(n_sets)]

            def f(i):
                act, _ = te(xs[i % n_sets], nz[i % n_sets])
                td(act, cs[i % n_sets])
            ms = event_time(f, n, 3, dev)
            out[f'b{B}'] = {'ms_per_call': ms, 'frames_per_s': B * FRAMES / ms * 1e3}
            del xs, cs, nz
            torch.cuda.empty_cache()
    return out


def sub_dsp(dev, enc, dec, peaks):
    """SURVEY 8f rows 1 and 3 next to the path: featurisation (preprocess.py:231-256) and the Griffin-Lim vocoder
    (convert.py:39-62, 300 iterations), and the wav -> wav variant where only waveforms cross PCIe."""
    import numpy as np
    from zs_b200 import _lib, dsp
    lib = _lib.lib()
    n_utt, frames = 64, 512
    L = 200 * (frames - 1) + 100
    rng = np.random.Generator(np.random.PCG64(5))
    wavs = [(0.1 * rng.standard_normal(L)).astype(np.float32) for _ in range(n_utt)]
    wav_dev = [torch.from_numpy(w).to(dev) for w in wavs]
    hbm = peaks.get('hbm_gbs', 6550.0)
    out = {'workload': f'{n_utt} utterances x {frames} frames ({n_utt * frames} frames, {L} samples each)'}
    dsp.get_spectrograms(wav_dev[:4], device=dev, to_host=False)

    def prof(fn):
        ms3, fl3, cnt = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
        torch.cuda.synchronize(dev)
        lib.zs_profile_begin()
        r = fn()
        _lib.check(lib.zs_profile_end(ms3, fl3, cnt))
        return r, ms3[2], fl3[2], int(cnt[2])
    specs, ms, fl, _ = prof(lambda: dsp.get_spectrograms(wav_dev, device=dev, dtype=torch.float16, to_host=False))
    nf = n_utt * frames
    by = n_utt * L * 4 + nf * 513 * 2
    out['featurisation'] = {'kernel_ms': ms, 'frames_per_s': nf / ms * 1e3, 'algorithmic_gb_per_s': by / (ms * 1e-3) / 1e9,
                            'frac_of_hbm_peak': by / (ms * 1e-3) / 1e9 / hbm, 'gflop_per_s': fl / (ms * 1e-3) / 1e9,
                            'what': 'wav fp32 in -> fp16 (T, 513) encoder-input rows out, one kernel'}
    rows = torch.cat(specs).float()
    gl = dsp.GriffinLim(dev, n_iter=300)
    gl.synthesize(rows[:4 * frames], [frames] * 4, trim=False, to_host=False)
    _, ms, fl, launches = prof(lambda: gl.synthesize(rows, [frames] * n_utt, trim=False, to_host=False))
    by = 300 * nf * 513 * (8 + 8 + 4) + nf * 513 * 8
    out['griffin_lim_300'] = {'kernel_ms': ms, 'frames_per_s': nf / ms * 1e3, 'launches': launches,
                              'algorithmic_gb_per_s': by / (ms * 1e-3) / 1e9, 'frac_of_hbm_peak': by / (ms * 1e-3) / 1e9 / hbm,
                              'fft_gflop_per_s': fl / (ms * 1e-3) / 1e9,
                              'what': 'per iteration and frame: 513 complex64 in + 513 fp32 magnitudes in + 513 complex64 out (20.5 KB); '
                                      '2.23 real 1024-point FFTs (shared-memory radix-8, fp32)'}
    # wav -> wav: pinned waveforms up, waveforms down; spectrograms never leave the device
    wav_host = torch.from_numpy(np.stack(wavs)).pin_memory()
    segs_per_utt = frames // FRAMES
    seeds = torch.arange(n_utt * segs_per_utt, dtype=torch.int64, device=dev) * 7919 + 1
    c = torch.zeros(n_utt * segs_per_utt, dtype=torch.int64, device=dev)

    def wav2wav():
        w = wav_host.to(dev, non_blocking=True)
        sp = dsp.get_spectrograms([w[i] for i in range(n_utt)], device=dev, dtype=torch.float16, to_host=False)
        x = torch.stack(sp).view(n_utt * segs_per_utt, FRAMES, 513)
        _, _, ids = enc.encode(x, None, layout='ntc', noise_seeds=seeds, want_act=False, want_logits=False)
        y = dec.decode(None, c, unit_ids=ids)                                   # (n, 513, 128)
        rows_ = y.permute(0, 2, 1).reshape(n_utt * frames, 513)
        return gl.synthesize(rows_, [frames] * n_utt, trim=False, to_host=True)
    wav2wav()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    res = wav2wav()
    torch.cuda.synchronize(dev)
    t = time.perf_counter() - t0
    out['wav_to_wav'] = {'frames_per_s': nf / t, 'ms': t * 1e3, 'h2d_bytes': n_utt * L * 4, 'd2h_bytes': int(sum(len(r) for r in res)) * 4,
                         'spectrogram_path_bytes_for_comparison': {'h2d': nf * 513 * 2, 'd2h': nf * 513 * 4},
                         'what': 'featurise -> Encoder -> Decoder -> Griffin-Lim x 300 -> de-emphasis; wall clock, one batch'}
    return out


def hbm_kernel_table(lib, names, ms, S, peaks):
    """Achieved HBM GB/s of the memory-bound kernels of one 960-segment call (algorithmic bytes / CUDA-event time)."""
    T8 = FRAMES // 8
    algo = {   # bytes per segment
        'pack_x_dual_kernel': 513 * FRAMES * 4 + (FRAMES + 6) * 576 * 2 + FRAMES * 513 * 2,     # x fp32 in; bank (halo 3, 576 ch) + cat slice fp16 out
        'bottleneck_onehot_kernel': T8 * ENC_SIZE * (4 + 4) + T8 * 4,                        # logits + noise in, ids out (no dense one-hot)
        'unit_gather_kernel': T8 * EMB_SIZE * 2 * 2,                                         # table rows in, x0 rows out
    }
    peak = peaks.get('hbm_gbs', 6550.0)
    out = {}
    for k, b in algo.items():
        t = [m for n, m in zip(names, ms) if n == k]
        if t:
            mean = sum(t) / len(t)
            gbs = b * S / (mean * 1e-3) / 1e9
            out[k] = {'us': mean * 1e3, 'algorithmic_mb': b * S / 1e6, 'gb_per_s': gbs, 'frac_of_hbm_peak': gbs / peak}
    return out


# ------------------------------------------------------------------------------------------------
# training sub-record / headline (BASELINE configs[3])
# ------------------------------------------------------------------------------------------------
def measure_train(args, dev, world, rank, steps, warmup, clocks=False):
    import torch.distributed as dist
    from zs_b200 import _lib, synthetic as syn, train as zt
    B = args.train_batch
    enc, dec = build_models(dev, 'fp16')
    enc.train()
    dec.train()
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
        torch.cuda.current_stream().synchronize()            # the reference reads loss.item() every iteration (trainer.py:336)

    def timed(fn):
        for i in range(warmup):
            fn(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / 1e3

    clk = None
    if clocks:
        with ClockSampler(dev.index) as clk:
            t_dev = timed(step_device)
    else:
        t_dev = timed(step_device)
    t_e2e = timed(step_e2e)
    # the all-reduce alone (same buffers, same stream pairing), for the record
    ar_ms = None
    if world > 1:
        def ar(i):
            zt.reduce_gradients(step.dec.grad)
            zt.reduce_gradients(step.enc.grad)
        ar_ms = event_time(ar, 10, 3, dev)
    lib = _lib.lib()
    ms3, fl3, cnt = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
    eager = zt.PretrainAE(enc, dec, use_graph=False) if world == 1 else None
    if eager is not None:      # kernel classes of one eager iteration (a graph replay bypasses the library's launch accounting)
        eager.step(xs[0], cs[0])
        lib.zs_profile_begin()
        eager.step(xs[1], cs[1])
        _lib.check(lib.zs_profile_end(ms3, fl3, cnt))
    frames = B * FRAMES * world
    n_params = step.enc.flat.numel() + step.dec.flat.numel()
    rec = {'workload': f'train_ae step (trainer.py:321-332), {B} segments x {FRAMES} frames per rank, enc_size {ENC_SIZE} one_hot, dropout 0.5, '
                       f'Adam(1e-4, (0.5, 0.9)), per-net clip 5; data-parallel x{world}',
           'ms_per_step': t_dev / steps * 1e3, 'frames_per_s': frames * steps / t_dev, 'steps': steps, 'warmup': warmup,
           'e2e_ms_per_step': t_e2e / steps * 1e3, 'e2e_frames_per_s': frames * steps / t_e2e,
           'h2d_bytes_per_step': B * (513 * FRAMES * 4 + 8), 'd2h_bytes_per_step': 4,
           'cuda_graph': bool(step.use_graph and step._graph is not None), 'n_ranks': world,
           'wgrad_side_streams': bool(step.async_wgrad),     # zs_wgrad_async: weight-gradient GEMMs overlap the data-gradient chain
           'kernel_ms_note': 'per-class times of one eager iteration with in-order launches (they add up); the timed graph replays overlap '
                             'the weight-gradient GEMMs with the rest, so ms_per_step is below their sum' if world == 1 else None,
           'allreduce': None if world == 1 else {'backend': 'nccl', 'bytes_per_step': n_params * 4, 'tensors': 2,
                                                 'ms_alone': ar_ms, 'algbw_gb_per_s': n_params * 4 / (ar_ms * 1e-3) / 1e9,
                                                 'overlap': 'decoder gradients (170 MB) reduce on a side stream under the encoder backward; '
                                                            'both all-reduces are nodes of the iteration\'s CUDA graph'},
           'loss': float(step.loss.item()), 'skipped_steps': step.n_skipped, 'applied_steps': step.applied_steps(),
           'kernel_ms': {'gemm': ms3[0], 'gru': ms3[1], 'other': ms3[2]} if eager is not None else None,
           'gemm_tflops': fl3[0] / (ms3[0] * 1e-3) / 1e12 if ms3[0] > 0 else None,
           'launches_per_step': int(sum(cnt)) if eager is not None else None}
    return rec, (clk.summary() if clk else None)


def run_train(args):
    """--workload train: BASELINE config 4 as the headline line."""
    import torch.distributed as dist
    import zs_b200  # noqa: F401
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', 1), ('RANK', 0), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        init_distributed(dev)
    W = max(args.warmup, 5)        # two eager steps + the graph capture come first
    rec, clocks = measure_train(args, dev, world, rank, args.steps, W, clocks=True)
    if rank == 0:
        peaks = load_peaks()
        peak = peaks.get('bf16_tflops', 1655.0)
        emit({'metric': 'spectrogram frames/s, pretrain_AE step (fwd + bwd + clip + Adam)', 'value': rec['frames_per_s'], 'unit': 'frames/s',
              'n_gpus': world, 'steps': args.steps, 'warmup': W, 'ms_per_step': rec['ms_per_step'], 'higher_is_better': True,
              'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp16 operands, f32 accumulate / master weights', 'data': 'synthetic',
              'config': {'workload': rec['workload'], 'l2': 'inputs rotate over 8 distinct batches; weights (110 MB fp16 + 221 MB fp32) exceed '
                                                            'what stays in L2 with the activations'},
              'clocks': clocks, 'gpu_launches': (rec['launches_per_step'] or 0) * args.steps,
              'e2e': {'value': rec['e2e_frames_per_s'], 'unit': 'frames/s', 'h2d_bytes_per_step': rec['h2d_bytes_per_step'],
                      'd2h_bytes_per_step': rec['d2h_bytes_per_step'], 'ms_per_step': rec['e2e_ms_per_step']},
              'roofline': {'bound': 'tensor', 'kernel': 'conv_gemm_kernel + wgrad_gemm_kernel (tcgen05), one eager iteration',
                           'achieved': rec['gemm_tflops'], 'peak': peak, 'unit': 'TFLOP/s',
                           'frac': rec['gemm_tflops'] / peak if rec['gemm_tflops'] else None, 'traffic': None,
                           'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst: the timed region is short)'},
              'train': rec})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# sharded product API (every N): ShardedPath.convert_utterances over the ranks
# ------------------------------------------------------------------------------------------------
def sub_critic(dev, check=True):
    """SURVEY 8 f4, forward part: PatchDiscriminator (model/model.py:113-166) value + auxiliary logits on 32 and 256 segments of
    128 frames (trainer.py:257-259 calls it on a 32-segment batch twice per step), checked against the oracle on two samples."""
    from zs_b200 import critic as zc, synthetic as syn
    sd = syn.critic_state_dict(3, n_class=33, seg_len=FRAMES)
    net = zc.PatchDiscriminator(n_class=33, seg_len=FRAMES)
    net.load_state_dict(sd, strict=True)
    net.to(dev).eval()
    flop_per_segment = 0.0
    H, W = 513, FRAMES
    for ci, co in ((1, 64), (64, 128), (128, 256), (256, 512), (512, 512)):
        H, W = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        flop_per_segment += 2.0 * H * W * co * ci * 25
    flop_per_segment += 2.0 * H * W * 32 * 512 + 2.0 * 34 * 32 * H * W
    rec = {'workload': f'PatchDiscriminator.forward(x, classify=True), (B, 513, {FRAMES}) fp32 in HBM -> value + 33 logits, eval mode',
           'gflop_per_segment': flop_per_segment / 1e9}
    for B in (32, 256):
        x = syn.spectrogram_batch(B, FRAMES, 970).to(dev)
        ms = event_time(lambda i: net(x, classify=True), 10, 3, dev)
        rec[f'b{B}'] = {'ms': ms, 'frames_per_s': B * FRAMES / ms * 1e3, 'tflops': B * flop_per_segment / (ms * 1e-3) / 1e12}
    if check:       # the oracle as the checker (CPU leg)
        from oracle import critic_oracle as corc
        x2 = syn.spectrogram_batch(2, FRAMES, 971)
        v, lg = net(x2.to(dev), classify=True)
        v_o, lg_o = corc.patch_discriminator(sd, x2)
        rec['max_abs_vs_oracle'] = max((v.cpu() - v_o).abs().max().item(), (lg.cpu() - lg_o).abs().max().item())
    return rec


def measure_sharded(dev, world, rank, enc, dec):
    import numpy as np
    import torch.distributed as dist
    from zs_b200.frontend import AutoencoderPath
    from zs_b200.shard import ShardedPath
    n_utt = 48 * world
    rng = np.random.Generator(np.random.PCG64(99))
    lengths = [int(v) for v in rng.integers(600, 2001, size=n_utt)]
    specs = [np.clip(np.random.Generator(np.random.PCG64(1000 + i)).random((L, 513), dtype=np.float32), 1e-8, 1.0) for i, L in enumerate(lengths)]
    spk = [int(v) for v in rng.integers(0, N_SPK, size=n_utt)]
    sp = ShardedPath(AutoencoderPath(enc, dec, device=dev))
    sp.convert_utterances(specs, spk, noise_seed=3, as_ids=True, gather=False)          # warm-up: workspaces, pinned staging
    times = {}
    for gather in (False, True):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res = sp.convert_utterances(specs, spk, noise_seed=3, as_ids=True, gather=gather)
        torch.cuda.synchronize(dev)
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[gather] = t.item()
    frames = sum(lengths)
    return {'workload': f'{n_utt} utterances of 600-2000 frames ({frames} frames) sharded whole over {world} ranks, '
                        'ShardedPath.convert_utterances (numpy in -> numpy spectrograms + int32 unit ids out, device-drawn noise)',
            'frames_per_s_rank_local_results': frames / times[False], 'frames_per_s_gathered_on_every_rank': frames / times[True],
            'seconds': {'local': times[False], 'gathered': times[True]},
            'note': 'wall clock of the host API, max over ranks; dominated by host-side numpy chunking and pageable D2H of fp32 spectrograms'}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import zs_b200  # noqa: F401
    from zs_b200 import _lib, synthetic as syn
    from zs_b200.frontend import StreamingResynthesizer
    from zs_b200.model import check_range, gumbel_from_uniform

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        init_distributed(dev)

    lib = _lib.lib()
    enc, dec = build_models(dev, args.operand)
    S, CALLS = args.segments, args.calls_per_step
    n_sets = 4
    # distinct input sets, rotated so a call never finds its inputs in L2 (126 MB)
    xs_host = [syn.spectrogram_batch(S, FRAMES, 100 * rank + i).pin_memory() for i in range(n_sets)]           # (S, 513, T) fp32
    cs_host = [syn.speaker_ids(S, N_SPK, 100 * rank + i).pin_memory() for i in range(n_sets)]
    # Gumbel noise value (model/model.py:95-98) of the reference-exact mode; an INPUT of the path in that mode
    us = [syn.gumbel_uniform((S, 16, ENC_SIZE), 100 * rank + i) for i in range(n_sets)]
    nz_host = [gumbel_from_uniform(u).pin_memory() for u in us]
    # the byte-saving input format: fp16 features in the (T, 513) layout of the HDF5 files (dataloader.py:74)
    xh_host = [x.permute(0, 2, 1).contiguous().half().pin_memory() for x in xs_host]
    xs = [t.to(dev) for t in xs_host]
    cs = [t.to(dev) for t in cs_host]
    nz = [t.to(dev) for t in nz_host]
    spec_out = torch.empty(S, 513, FRAMES, device=dev)
    ids_out = torch.empty(S, 16, dtype=torch.int32, device=dev)

    def call_device(k):
        _, _, ids = enc.encode(xs[k], nz[k], want_act=False, want_logits=False)
        dec.decode(None, cs[k], unit_ids=ids, out=spec_out)
        ids_out.copy_(ids)

    def step_device(i):
        for j in range(CALLS):
            call_device((i * CALLS + j) % n_sets)

    depth = max(2, args.e2e_buffers)
    streamer = StreamingResynthesizer(enc, dec, micro_batch=S, n_buffers=depth, device=dev)
    # host output buffers: one per call in flight (+1 being filled); a buffer is reused only after its call completed
    out_hosts = [(torch.empty(S, 513, FRAMES).pin_memory(), torch.empty(S, 16, dtype=torch.int32).pin_memory()) for _ in range(depth + 1)]
    out_hosts16 = [torch.empty(S, 513, FRAMES, dtype=torch.float16).pin_memory() for _ in range(depth + 1)]
    pending = []
    mode = {'name': None}
    n_issued = [0]

    def step_e2e(i):
        for j in range(CALLS):
            q = n_issued[0]
            n_issued[0] += 1
            k = q % n_sets
            sh, ih = out_hosts[q % len(out_hosts)]
            if len(pending) >= depth:
                pending.pop(0).synchronize()        # the oldest call's results are complete (and its host buffers consumable)
            if mode['name'] == 'fp32_reference_exact':
                pending.append(streamer.run_async(xs_host[k], cs_host[k], sh, ih, nz_host[k]))
            elif mode['name'] == 'fp16_in_fp16_out':
                pending.append(streamer.run_async(xh_host[k], cs_host[k], out_hosts16[q % len(out_hosts16)], ih, None, layout='ntc',
                                                  noise_seed=1234, segment0=q * S))
            else:
                pending.append(streamer.run_async(xh_host[k], cs_host[k], sh, ih, None, layout='ntc', noise_seed=1234, segment0=q * S))

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
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        if finish:
            finish()            # the timing stream waits for every call's download before the closing event
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / 1e3

    W = max(args.warmup, 3)
    with ClockSampler(local) as clk:
        t_dev = timed(step_device, args.steps, W)
        # per-launch CUDA events over one more full step, straight after the timed region (same clocks / thermal state)
        ms3, fl3, cnt = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
        lib.zs_profile_begin()
        step_device(0)
        n_rec = lib.zs_profile_detail(None, None, None, 0)
        d_ms, d_fl, d_cls = (C.c_double * n_rec)(), (C.c_double * n_rec)(), (C.c_int * n_rec)()
        lib.zs_profile_detail(d_ms, d_fl, d_cls, n_rec)
        d_names = [lib.zs_profile_name(i).decode() for i in range(n_rec)]
        _lib.check(lib.zs_profile_end(ms3, fl3, cnt))
    clocks = clk.summary()
    launches_per_step = int(sum(cnt))
    gemm_ms, gemm_flops, gemm_launches = ms3[0], fl3[0], int(cnt[0])
    gru_ms, other_ms = ms3[1], ms3[2]

    def finish_e2e():
        while pending:
            torch.cuda.current_stream().wait_event(pending.pop(0))

    e2e = {}
    for name in ('fp16_ntc_device_noise', 'fp32_reference_exact', 'fp16_in_fp16_out'):
        mode['name'] = name
        t = timed(step_e2e, args.steps, W, finish_e2e)
        h2d = S * CALLS * ((513 * FRAMES * 2 + 8) if name.startswith('fp16') else (513 * FRAMES * 4 + 8 + 16 * ENC_SIZE * 4))
        d2h = S * CALLS * (513 * FRAMES * (2 if name.endswith('fp16_out') else 4) + 16 * 4)
        e2e[name] = {'value': S * CALLS * FRAMES * world * args.steps / t, 'unit': 'frames/s', 'ms_per_step': t / args.steps * 1e3,
                     'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                     'dma_gb_per_s_per_gpu': {'h2d': h2d * args.steps / t / 1e9, 'd2h': d2h * args.steps / t / 1e9}}
    saturations = streamer.check_range()

    frames_per_step = S * CALLS * FRAMES * world
    value = frames_per_step * args.steps / t_dev
    extras = not args.no_extras

    # ---- sub-records that run on every rank ---------------------------------------------------------
    train_rec = sharded_rec = None
    if extras:
        sharded_rec = measure_sharded(dev, world, rank, enc, dec)
        torch.cuda.empty_cache()
        train_rec, _ = measure_train(args, dev, world, rank, 30, 8)
        torch.cuda.empty_cache()

    if rank == 0:
        peaks = load_peaks()
        burst, sustained = peaks.get('bf16_tflops', 1655.0), peaks.get('bf16_tflops_sustained', 1400.0)
        timed_s = t_dev
        regime = 'sustained' if timed_s >= 3.0 else 'burst'
        peak = sustained if regime == 'sustained' else burst
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        # DRAM traffic of the GEMM launches from the committed ncu --set full capture (same call size only)
        traffic, traffic_note = None, None
        for fn in ('r02_gemm_traffic.json', 'r01_gemm_traffic.json'):
            try:
                tr = json.load(open(os.path.join(ROOT, 'profiles', fn)))
                if S == tr['micro_batch']:
                    traffic = (tr['dram_read_bytes'] + tr['dram_write_bytes']) * CALLS
                    traffic_note = (f"bytes per step over all {gemm_launches} GEMM launches = {CALLS} calls x "
                                    f"{(tr['dram_read_bytes'] + tr['dram_write_bytes']) / 1e9:.2f} GB (ncu --set full capture of one {tr['micro_batch']}-segment "
                                    f"call, profiles/{fn})")
                    break
            except Exception:
                pass
        head = e2e['fp16_ntc_device_noise']
        line = {
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': W,
            'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': f'{args.operand} operands, f32 accumulate', 'data': 'synthetic', 'config': config(args, world),
            'timed_region_s': timed_s, 'clocks': clocks, 'gpu_launches': launches_per_step * args.steps,
            'e2e': {'value': head['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': head['h2d_bytes_per_step'],
                    'd2h_bytes_per_step': head['d2h_bytes_per_step'], 'ms_per_step': head['ms_per_step'],
                    'dma_gb_per_s_per_gpu': head['dma_gb_per_s_per_gpu'], 'host_placement': numa, 'calls_in_flight': depth,
                    'mode': 'fp16 features in the (T, 513) HDF5 layout up (bit-identical results: the path rounds its input to fp16 first thing), '
                            'Gumbel noise drawn on the device, fp32 spectrograms + int32 unit ids down',
                    'api': 'StreamingResynthesizer.run_async: pinned host in -> pinned host out, calls issued back to back '
                           '(upload of call i+1 under compute/download of call i)',
                    'pcie_ceiling': 'tools/pcie_probe.py: 51.7 GiB/s H2D / 51.9 D2H alone, 14.8-20.8 GiB/s per GPU with all eight copying both ways (r01 box)'},
            'e2e_modes': dict(e2e, note='fp16_ntc_device_noise (headline) and fp32_reference_exact return bit-identical spectrograms for the same units; '
                                        'fp16_in_fp16_out also rounds the OUTPUT once to fp16 (<= 2.5e-4 absolute on the (0, 1) sigmoid output): the fewest bytes'),
            'roofline': {'bound': 'tensor', 'kernel': 'conv_gemm_kernel (tcgen05 implicit GEMM, all conv/linear layers)',
                         'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak if peak else None,
                         'frac_of_burst_peak': achieved / burst, 'frac_of_sustained_peak': achieved / sustained,
                         'peak_burst': burst, 'peak_sustained': sustained, 'regime': regime,
                         'traffic': traffic, 'traffic_note': traffic_note,
                         'peak_source': f'MEASURED_PEAKS.json bf16_tflops{"_sustained" if regime == "sustained" else ""} (fp16 and bf16 share the kind::f16 '
                                        f'rate); the timed region kept the GPU under load for {timed_s:.1f} s and the profiled step follows it directly'
                         if peaks else 'fallback peaks (B200_PROFILING.md)', 'launches_per_step': gemm_launches,
                         'algorithmic_gflop_per_step': gemm_flops / 1e9, 'kernel_ms_per_step': gemm_ms,
                         'kernel_ms_per_960_segment_call': gemm_ms / CALLS,
                         'share_of_step': gemm_ms / (gemm_ms + gru_ms + other_ms) if gemm_ms else None,
                         'gru_ms_per_step': gru_ms, 'other_ms_per_step': other_ms},
            'hbm_kernels': hbm_kernel_table(lib, d_names, d_ms, S, peaks),
            'whole_path_tflops': value / world * MFLOP_PER_FRAME * 1e6 / 1e12,
            'fp16_saturations': saturations,
        }
        # large-batch result must equal the 32-segment-batch result bit for bit (segments are independent)
        call_device(0)
        _, _, i32 = enc.encode(xs[0][:32], nz[0][:32])
        s32 = dec.decode(None, cs[0][:32], unit_ids=i32)
        line['self_check'] = {'batch_invariant': bool(torch.equal(s32, spec_out[:32]) and torch.equal(i32, ids_out[:32]))}
        # the fp16 (T, 513) upload gives the same spectrograms as the fp32 (513, T) one for the same units
        _, _, i16 = enc.encode(xh_host[0][:32].to(dev), nz[0][:32], layout='ntc')
        line['self_check']['fp16_ntc_input_identical'] = bool(torch.equal(i16, i32))
        if not args.no_cpu_baseline and world == 1:
            n_par = min(S, args.parity_segments)
            par = cpu_parity(xs_host[0][:n_par], cs_host[0][:n_par], us[0][:n_par], ids_out[:n_par].cpu(),
                             xs_host[0][:32], cs_host[0][:32], ids_out[:32].cpu(), spec_out[:32].cpu())
            line['parity'] = par
            n_cpu = min(S, args.cpu_sample)
            fps, t, cores = cpu_time(n_cpu, 2, 1)
            line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'{n_cpu} segments x {FRAMES} frames in one batch, 2 timed passes of oracle/ae_oracle.py (torch fp32, all host cores)'}
            fps1, t1, _ = cpu_time(8, 1, 1, batch1=True)
            line['cpu_baseline_b1'] = {'value': fps1, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                       'sample': '8 chunks of 128 frames, ONE model call per chunk with its own torch.rand draw - the reference\'s call pattern '
                                                 '(convert.py:70-83, 154-165)'}
        if extras and world == 1:
            line['configs'] = sub_small_batches(enc, dec, dev)
            line['configs']['config5_patcher_2000_frames'] = sub_config5(dev, args.operand)
            try:
                line['dsp'] = sub_dsp(dev, enc, dec, peaks)
            except Exception as exc:
                line['dsp'] = {'unavailable': f'{type(exc).__name__}: {exc}'[:300]}
            try:
                line['critic'] = sub_critic(dev, check=not args.no_cpu_baseline)
            except Exception as exc:
                line['critic'] = {'unavailable': f'{type(exc).__name__}: {exc}'[:300]}
            try:
                line['cuda_eager_baseline'] = sub_eager(dev)
                line['cuda_eager_baseline']['ours_vs_eager_b960'] = (value / world) / line['cuda_eager_baseline']['b960']['frames_per_s']
            except Exception as exc:          # a baseline must not take the headline down with it
                line['cuda_eager_baseline'] = {'unavailable': f'{type(exc).__name__}: {exc}'[:300]}
        if train_rec is not None:
            line['train'] = train_rec
        if sharded_rec is not None:
            line['sharded_api'] = sharded_rec
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    with stdout_to_stderr():        # everything native (NCCL banners / NCCL_DEBUG lines) goes to stderr; the JSON line to the real stdout
        if args.impl == 'reference':
            run_reference(args)
        elif args.workload == 'train':
            run_train(args)
        else:
            run_ours(args)


if __name__ == '__main__':
    main()
