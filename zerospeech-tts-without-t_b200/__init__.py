"""B200-native hot path of the ZeroSpeech "TTS without T" ASR-TTS autoencoder.

Drop-in `Encoder` / `Decoder` modules (reference: model/model.py:283-489) whose
forward passes run hand-written sm_100a CUDA kernels through the C-ABI library
`libzsae.so` (include/zs_ae.h).  There is no CPU fallback: calling a forward
without the built library or without a CUDA device raises.
"""
__all__ = ['synthetic']
