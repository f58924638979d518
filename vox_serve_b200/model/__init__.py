"""Model-adapter registry with the reference's public surface (``vox_serve/model/__init__.py:17-179``):
``MODEL_REGISTRY``, ``get_model_class``, ``load_model``, ``register_model``, ``list_supported_models``.
Only adapters whose whole decode + vocoder path runs on the sm_100a kernels are registered: Orpheus (+ SNAC) and CSM
(+ Mimi) and Qwen3-TTS (+ its streaming 12 Hz codec decoder)."""
from __future__ import annotations

from typing import Any, Dict, Type

import torch

from ..sampling import SamplingConfig
from .base import BaseLM, BaseLMWithDepth, PreprocessOutput
from .csm import CSMModel
from .orpheus import OrpheusModel
from .qwen3_tts import Qwen3TTSModel

MODEL_REGISTRY: Dict[str, Type[BaseLM]] = {
    "orpheus": OrpheusModel,
    "canopylabs/orpheus-3b-0.1-ft": OrpheusModel,
    "csm": CSMModel,
    "sesame/csm-1b": CSMModel,
    "qwen3-tts": Qwen3TTSModel,
    "qwen/qwen3-tts": Qwen3TTSModel,
}


def get_model_class(model_name: str) -> Type[BaseLM]:
    key = model_name.lower()
    if key in MODEL_REGISTRY:
        return MODEL_REGISTRY[key]
    for pattern, cls in MODEL_REGISTRY.items():
        if pattern in key:
            return cls
    raise ValueError(f"No model class found for '{model_name}'. Available model patterns: {list(MODEL_REGISTRY)}")


_OVERRIDES = ("top_p", "top_k", "min_p", "temperature", "max_tokens", "repetition_penalty", "repetition_window",
              "cfg_scale")


def load_model(model_name: str, device: str = "cuda", dtype: torch.dtype = torch.bfloat16, top_p: float = None,
               top_k: int = None, min_p: float = None, temperature: float = None, max_tokens: int = None,
               repetition_penalty: float = None, repetition_window: int = None, cfg_scale: float = None,
               greedy: bool = False, enable_torch_compile: bool = False, **kwargs: Any) -> BaseLM:
    """CLI sampling overrides replace fields of the adapter's default config; ``greedy`` keeps the
    repetition penalty active (model/__init__.py:132-156)."""
    cls = get_model_class(model_name)
    if kwargs.pop("detokenize_interval", None) is not None:
        raise ValueError(f"Detokenize interval is only supported for Qwen3TTS models, got {model_name}")
    model = cls(model_name=model_name, device=device, dtype=dtype, enable_torch_compile=enable_torch_compile, **kwargs)
    given = dict(top_p=top_p, top_k=top_k, min_p=min_p, temperature=temperature, max_tokens=max_tokens,
                 repetition_penalty=repetition_penalty, repetition_window=repetition_window, cfg_scale=cfg_scale)
    if greedy or any(v is not None for v in given.values()):
        cur = model.default_sampling_config
        merged = {k: (given[k] if given[k] is not None else getattr(cur, k)) for k in _OVERRIDES}
        model.default_sampling_config = SamplingConfig(greedy=greedy, **merged)
    return model


def register_model(pattern: str, model_class: Type[BaseLM]) -> None:
    MODEL_REGISTRY[pattern.lower()] = model_class


def list_supported_models() -> Dict[str, Type[BaseLM]]:
    return MODEL_REGISTRY.copy()
