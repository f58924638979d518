"""Model-adapter contract (same member names and meanings as ``vox_serve/model/base.py:13-277``) so the
worker and the reference schedulers can treat a B200 adapter exactly like a reference adapter."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Coroutine, List, Optional, Tuple

import torch

from ..requests import Request
from ..sampling import SamplingConfig


@dataclass
class PreprocessOutput:
    """What ``preprocess`` hands to ``ModelWorker.prepare_lm_inputs`` (model/base.py:13-26)."""
    input_tokens: Any
    repetition_cache: Optional[torch.Tensor] = None
    input_masks: Optional[torch.Tensor] = None
    input_features: Optional[torch.Tensor] = None
    decoder_cache: Any = None


class BaseLM(ABC):
    def __init__(self, model_name: str, device: str = "cuda", dtype: torch.dtype = torch.bfloat16,
                 enable_torch_compile: bool = False, audio_decoder_device: str = None):
        self.model_name, self.device, self.dtype = model_name, device, dtype
        self.enable_torch_compile = enable_torch_compile      # accepted for CLI compatibility; never used
        self.audio_decoder_device = audio_decoder_device or device

    # ---- static model facts ---------------------------------------------------------------------
    @property
    @abstractmethod
    def n_codebooks(self) -> int: ...

    @property
    @abstractmethod
    def num_attention_heads(self) -> int: ...

    @property
    @abstractmethod
    def num_key_value_heads(self) -> int: ...

    @property
    @abstractmethod
    def num_hidden_layers(self) -> int: ...

    @property
    @abstractmethod
    def hidden_size(self) -> int: ...

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    @abstractmethod
    def detokenize_interval(self) -> int: ...

    @property
    @abstractmethod
    def detokenize_overlap(self) -> int: ...

    @property
    @abstractmethod
    def max_tokens(self) -> int: ...

    @property
    @abstractmethod
    def vocab_size(self) -> int: ...

    @property
    @abstractmethod
    def n_channels(self) -> int: ...

    @property
    @abstractmethod
    def output_audio_length(self) -> int: ...

    # ---- capability flags (defaults of model/base.py:85-132) ------------------------------------------
    has_depth_transformer = False
    supports_audio_input = False
    needs_watermarking = False
    watermarker_type = None
    needs_input_features = False
    needs_input_masks = False
    supports_input_streaming = False

    @property
    def use_repetition_penalty(self) -> bool:
        c = getattr(self, "default_sampling_config", None)
        return c is not None and c.repetition_penalty is not None and c.repetition_penalty != 1.0

    def audio_decoder_initial_cache(self, batch_size: int):
        return None

    # ---- per-step work ------------------------------------------------------------------------
    @abstractmethod
    def is_stop_id(self, token_ids: List[int]) -> bool: ...

    @abstractmethod
    def preprocess(self, prompt: str = None, audio_path: str = None, **kwargs) -> PreprocessOutput: ...

    @abstractmethod
    def forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, attn_wrapper, kv_cache: torch.Tensor,
                **kwargs) -> torch.Tensor: ...

    @abstractmethod
    def sampling(self, logits: torch.Tensor, requests: List[Request], sampling_params: Optional[SamplingConfig] = None,
                 repetition_cache: Optional[torch.Tensor] = None, cfg_scale: Optional[float] = None,
                 **kwargs) -> Tuple[torch.Tensor, Coroutine]: ...

    @abstractmethod
    def postprocess(self, token_ids: torch.Tensor, **kwargs) -> torch.Tensor: ...


class BaseLMWithDepth(BaseLM):
    """Adapters with a depth transformer (``vox_serve/model/base.py:280-447``): the backbone's ``forward`` also
    returns the hidden state the depth decoder starts from, ``sampling`` returns the depth decoder's first input,
    and ``depth_forward`` / ``depth_sampling`` run one codebook step.  The B200 adapters additionally expose
    ``frame_device``: the whole frame (backbone sample + every depth step) as one capturable launch sequence."""
    has_depth_transformer = True

    @property
    @abstractmethod
    def depth_n_codebooks(self) -> int: ...

    @property
    @abstractmethod
    def depth_num_attention_heads(self) -> int: ...

    @property
    @abstractmethod
    def depth_num_key_value_heads(self) -> int: ...

    @property
    @abstractmethod
    def depth_num_hidden_layers(self) -> int: ...

    @property
    @abstractmethod
    def depth_hidden_size(self) -> int: ...

    @property
    @abstractmethod
    def depth_head_dim(self) -> int: ...

    @property
    @abstractmethod
    def depth_vocab_size(self) -> int: ...

    @abstractmethod
    def depth_forward(self, hidden_states: torch.Tensor, position_ids: torch.Tensor, attn_wrapper, kv_cache: torch.Tensor,
                      **kwargs) -> torch.Tensor: ...

    @abstractmethod
    def depth_sampling(self, logits: torch.Tensor, i_iteration: int, requests: List[Request],
                       sampling_params: Optional[SamplingConfig] = None, cfg_scale: Optional[float] = None, **kwargs): ...
