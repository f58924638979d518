"""vox_serve_b200 -- B200-native (sm_100a) streaming SpeechLM decode + vocoder path for VoxServe.

Host side in Python/PyTorch (device memory, streams, NCCL plumbing); all arithmetic on the hot path runs
in hand-written CUDA behind the C ABI of include/vb_api.h (vox_serve_b200/lib/libvoxb200.so).
Module names mirror the reference package (flashinfer_utils, sampling, tokenizer.snac, model, worker) so
the reference's adapters and schedulers can import them unchanged.
"""
__version__ = "0.1.0"
