"""tcgen05 tf32 hi/lo weight packing shared by the codec decoders (Mimi, Qwen3 codec): a convolution whose input width is a
multiple of 32 channels and at least ``min_cin`` (Mimi 128, Qwen3 codec 96: its 96-channel stage runs at 19 200 positions per
chunk; ``VB_CODEC_TC_MIN_CIN`` overrides both) runs on ``snac_gemm_tf32x3_kernel`` (csrc/snac_mma.cu);
narrower layers (test-sized configurations, the last SEANet / decoder blocks) stay on the fp32 SIMT kernels with the same
arguments.  ``VB_CODEC_FP32=1`` keeps everything on the SIMT kernels."""
from __future__ import annotations

import os
from typing import Optional

import torch

from .. import _lib, ops
from .._lib import call


def tc_enabled() -> bool:
    return os.environ.get("VB_CODEC_FP32", "0") != "1"


def _min_cin(default: int) -> int:
    return int(os.environ.get("VB_CODEC_TC_MIN_CIN", default))


def pack_conv(w: torch.Tensor, cin: int, ksize: int, min_cin: int = 128) -> Optional[torch.Tensor]:
    """w [Cout][Cin * ksize] (nn.Conv1d weight flattened: channel-major, tap-minor), on the device -> packed tiles, or None
    when the layer stays on the SIMT kernel."""
    if not (tc_enabled() and w.is_cuda and cin % 32 == 0 and cin >= _min_cin(min_cin)):
        return None
    cout = w.shape[0]
    tap_major = w.view(cout, cin, ksize).permute(0, 2, 1).reshape(cout, ksize * cin).contiguous()
    return _pack(tap_major, 1, cout, ksize * cin)


def pack_convtr(wp: torch.Tensor, cin: int, min_cin: int = 128) -> Optional[torch.Tensor]:
    """wp [stride][Cout][2 Cin] (per output phase: tap 0 | tap 1) -> packed tiles or None."""
    if not (tc_enabled() and wp.is_cuda and cin % 32 == 0 and cin >= _min_cin(min_cin)):
        return None
    s, cout, k = wp.shape
    return _pack(wp.contiguous(), s, cout, k)


def _pack(w: torch.Tensor, phases: int, M: int, K: int) -> Optional[torch.Tensor]:
    n = _lib.load().vb_snac_tf32x3_bytes(phases, M, K)
    if n <= 0:
        return None
    dst = torch.empty(n, dtype=torch.uint8, device=w.device)
    call("vb_snac_pack_tf32x3", dst.data_ptr(), w.data_ptr(), phases, M, K, ops._stream())
    return dst
