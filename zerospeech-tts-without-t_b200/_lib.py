"""ctypes binding of libzsae.so (include/zs_ae.h) and its in-tree build recipe.

The library is built IN-TREE (next to this file) with nvcc for sm_100a so that it
travels with the repository snapshot; there is no JIT cache and no CPU fallback -
`lib()` raises if the shared object is missing.
"""
import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
SO_PATH = os.path.join(_HERE, 'libzsae.so')
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'zs_ae.h')
import glob


NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']

ENC_MODES = {'continues': 0, 'one_hot': 1, 'multilabel_binary': 2, 'gumbel_t': 3, 'binary': 4}
OPERANDS = {'fp16': 0, 'bf16': 1}


def _stale():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = glob.glob(os.path.join(_CSRC, '*.cu')) + glob.glob(os.path.join(_CSRC, '*.cuh')) + [HEADER]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/zs_ae.cu -> libzsae.so for sm_100a (cross-compiles without a GPU)."""
    if not force and not _stale():
        return SO_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: libzsae.so cannot be built')
    # ZS_BUILD_EXPERIMENTS=1 compiles the timing-experiment knobs in (tools/gemm_dbg.sh, tools/gru_probe.py); the default
    # library reads no environment variables and carries no result-changing debug paths
    extra = ['-DZS_EXPERIMENTS'] if os.environ.get('ZS_BUILD_EXPERIMENTS') == '1' else []
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', SO_PATH, os.path.join(_CSRC, 'zs_ae.cu')]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO_PATH


class EncoderCfg(C.Structure):
    _fields_ = [('c_in', C.c_int32), ('c_h1', C.c_int32), ('c_h2', C.c_int32), ('c_h3', C.c_int32),
                ('enc_size', C.c_int32), ('enc_mode', C.c_int32), ('seg_len', C.c_int32), ('operand', C.c_int32),
                ('ns', C.c_float), ('train', C.c_int32)]


class EncoderWeights(C.Structure):
    _fields_ = [('conv1s_w', C.c_void_p * 7), ('conv1s_b', C.c_void_p * 7), ('conv_w', C.c_void_p * 7),
                ('conv_b', C.c_void_p * 7), ('dense_w', C.c_void_p * 4), ('dense_b', C.c_void_p * 4),
                ('gru_w_ih', C.c_void_p * 2), ('gru_w_hh', C.c_void_p * 2), ('gru_b_ih', C.c_void_p * 2),
                ('gru_b_hh', C.c_void_p * 2), ('linear_w', C.c_void_p), ('linear_b', C.c_void_p)]


class DecoderCfg(C.Structure):
    _fields_ = [('c_in', C.c_int32), ('c_out', C.c_int32), ('c_h', C.c_int32), ('c_a', C.c_int32),
                ('seg_len', C.c_int32), ('output_mask', C.c_int32), ('operand', C.c_int32), ('ns', C.c_float),
                ('train', C.c_int32)]


class DecoderWeights(C.Structure):
    _fields_ = [('conv_w', C.c_void_p * 6), ('conv_b', C.c_void_p * 6), ('dense_w', C.c_void_p * 4),
                ('dense_b', C.c_void_p * 4), ('gru_w_ih', C.c_void_p * 2), ('gru_w_hh', C.c_void_p * 2),
                ('gru_b_ih', C.c_void_p * 2), ('gru_b_hh', C.c_void_p * 2), ('dense5_w', C.c_void_p),
                ('dense5_b', C.c_void_p), ('linear_w', C.c_void_p), ('linear_b', C.c_void_p),
                ('input_emb_w', C.c_void_p), ('input_emb_b', C.c_void_p), ('emb', C.c_void_p * 5)]


class PatcherCfg(C.Structure):
    _fields_ = [('c_in', C.c_int32), ('c_out', C.c_int32), ('c_h', C.c_int32), ('c_a', C.c_int32), ('operand', C.c_int32),
                ('ns', C.c_float)]


class PatcherWeights(C.Structure):
    _fields_ = [('input_w', C.c_void_p), ('input_b', C.c_void_p), ('dense_w', C.c_void_p * 4), ('dense_b', C.c_void_p * 4),
                ('gru_w_ih', C.c_void_p * 2), ('gru_w_hh', C.c_void_p * 2), ('gru_b_ih', C.c_void_p * 2),
                ('gru_b_hh', C.c_void_p * 2), ('dense5_w', C.c_void_p), ('dense5_b', C.c_void_p), ('linear_w', C.c_void_p),
                ('linear_b', C.c_void_p), ('emb', C.c_void_p * 2)]


class ConvDesc(C.Structure):
    _fields_ = [('w', C.c_void_p), ('m_rows', C.c_int32), ('m_valid', C.c_int32), ('taps', C.c_int32),
                ('c_in_pad', C.c_int32), ('w_taps', C.c_int32), ('bank', C.c_int32),
                ('in_', C.c_void_p), ('in_rows', C.c_int32), ('in_pitch', C.c_int32), ('in_row0', C.c_int32),
                ('c_in_valid', C.c_int32), ('stride', C.c_int32), ('B', C.c_int32), ('T_out', C.c_int32),
                ('bias', C.c_void_p), ('spk', C.c_void_p), ('n_spk', C.c_int32), ('lrelu', C.c_int32), ('ns', C.c_float),
                ('inorm', C.c_int32), ('res_mode', C.c_int32), ('res', C.c_void_p), ('res_rows', C.c_int32),
                ('res_pitch', C.c_int32), ('res_halo', C.c_int32), ('act', C.c_int32), ('out_mode', C.c_int32),
                ('out', C.c_void_p), ('out_rows', C.c_int32), ('out_pitch', C.c_int32), ('out_halo', C.c_int32),
                ('out_choff', C.c_int32), ('accumulate', C.c_int32), ('operand', C.c_int32), ('nb_hint', C.c_int32), ('out_f16', C.c_int32)]


class WgradDesc(C.Structure):
    _fields_ = [('dy', C.c_void_p), ('dy_rows', C.c_int32), ('dy_pitch', C.c_int32), ('dy_channels', C.c_int32),
                ('dy_ch0', C.c_int32), ('dy_row0', C.c_int32), ('c_out', C.c_int32),
                ('x', C.c_void_p), ('x_rows', C.c_int32), ('x_pitch', C.c_int32), ('x_channels', C.c_int32),
                ('x_ch0', C.c_int32), ('x_row0', C.c_int32), ('c_in', C.c_int32), ('stride', C.c_int32),
                ('B', C.c_int32), ('T', C.c_int32), ('taps', C.c_int32),
                ('grad', C.c_void_p), ('c_in_total', C.c_int32), ('ci_off', C.c_int32), ('k', C.c_int32),
                ('tap0', C.c_int32), ('ps_c', C.c_int32), ('scale', C.c_float), ('grad_is_zero', C.c_int32)]


# every symbol include/zs_ae.h declares: name -> (restype, argtypes)
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
SYMBOLS = {
    'zs_last_error': (C.c_char_p, []),
    'zs_version': (_i, []),
    'zs_device_check': (_i, []),
    'zs_saturation_count': (_i, [_vp, C.POINTER(C.c_ulonglong), _i]),
    'zs_encoder_pack': (_i, [C.POINTER(EncoderCfg), C.POINTER(EncoderWeights), _vp, C.POINTER(_vp)]),
    'zs_decoder_pack': (_i, [C.POINTER(DecoderCfg), C.POINTER(DecoderWeights), _vp, C.POINTER(_vp)]),
    'zs_encoder_free': (None, [_vp]),
    'zs_decoder_free': (None, [_vp]),
    'zs_encoder_workspace_bytes': (_sz, [_vp, _i, _i]),
    'zs_decoder_workspace_bytes': (_sz, [_vp, _i, _i]),
    'zs_encoder_forward': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'zs_encoder_forward_x': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'zs_decoder_forward': (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _sz, _vp]),
    'zs_decoder_forward_x': (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _vp, _sz, _vp]),
    'zs_patcher_pack': (_i, [C.POINTER(PatcherCfg), C.POINTER(PatcherWeights), _vp, C.POINTER(_vp)]),
    'zs_patcher_free': (None, [_vp]),
    'zs_patcher_workspace_bytes': (_sz, [_vp, _i, _i]),
    'zs_patcher_forward': (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _sz, _vp]),
    'zs_stft_tile_frames': (_i, []),
    'zs_griffin_lim_workspace_bytes': (_sz, [C.c_longlong, C.c_longlong]),
    'zs_griffin_lim': (_i, [_vp, _vp, _i, C.c_longlong, C.c_longlong, _i, _i, C.c_float, _vp, _vp, _sz, _vp]),
    'zs_frame_power': (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    'zs_spectrogram': (_i, [_vp, _vp, _i, _i, C.c_float, _vp, _vp, _vp]),
    'zs_profile_begin': (None, []),
    'zs_profile_end': (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    'zs_profile_detail': (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), _i]),
    'zs_profile_name': (C.c_char_p, [_i]),
    'zs_launch_counts': (None, [C.POINTER(C.c_longlong)]),
    'zs_bottleneck_one_hot': (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    'zs_set_gemm_pair_mode': (None, [_i]),
    'zs_conv1d_cl': (_i, [C.POINTER(ConvDesc), _vp]),
    'zs_pack_nct': (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, C.c_float, _i, _i, _vp]),
    'zs_gru_recurrence': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    # pretrain_AE step
    'zs_encoder_repack': (_i, [_vp, C.POINTER(EncoderWeights), _vp]),
    'zs_decoder_repack': (_i, [_vp, C.POINTER(DecoderWeights), _vp]),
    'zs_encoder_train_workspace_bytes': (_sz, [_vp, _i, _i]),
    'zs_decoder_train_workspace_bytes': (_sz, [_vp, _i, _i]),
    'zs_encoder_forward_train': (_i, [_vp, _vp, _i, _i, _vp, C.c_float, C.c_uint64, _vp, C.POINTER(_vp), _vp, _vp, _vp,
                                      _vp, _sz, _vp]),
    'zs_encoder_backward': (_i, [_vp, _vp, C.c_float, _vp, _vp, _i, _i, C.c_float, C.c_uint64, _vp, C.POINTER(_vp),
                                 C.c_float, C.POINTER(EncoderWeights), _vp, _sz, _vp]),
    'zs_decoder_forward_train': (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    'zs_decoder_backward': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, C.c_float, _vp, C.POINTER(DecoderWeights), _vp,
                                 _vp, _sz, _vp]),
    'zs_conv2d_gather': (_i, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              _vp, C.c_float, _vp, C.c_int, _vp]),
    'zs_instnorm2d_stats': (_i, [_vp, C.c_int, C.c_longlong, C.c_int, C.c_int, _vp, _vp]),
    'zs_critic_head': (_i, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp]),
    'zs_wgrad_async': (_i, [C.c_int]),
    'zs_wgrad_join': (_i, [_vp]),
    'zs_grad_sqnorm': (_i, [_vp, _sz, _vp, _vp]),
    'zs_adam_step': (_i, [_vp, _vp, _vp, _vp, _sz, _vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                          C.c_float, _i, _vp, _vp, _vp]),
    'zs_wgrad_cl': (_i, [C.POINTER(WgradDesc), _vp]),
    'zs_train_meta_begin': (_i, [_vp, C.c_uint64, _vp]),
    'zs_train_meta_commit': (_i, [_vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp]),
}

_LIB = None


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f'{SO_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(the autoencoder path has no CPU fallback)')
        handle = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def saturation_count(stream=None, reset=True):
    """fp16 values the GEMM epilogues had to clamp to +-65504 on the current device since the last reset (synchronises)."""
    n = C.c_ulonglong(0)
    check(lib().zs_saturation_count(C.c_void_p(stream or 0), C.byref(n), int(reset)))
    return int(n.value)


def check(rc):
    if rc != 0:
        raise RuntimeError('libzsae: ' + lib().zs_last_error().decode())
