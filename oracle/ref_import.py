"""Import recipe for running the *reference's own code* on CPU in the authoring container.

Used ONLY by ``oracle/gen_golden.py`` (``/root/reference`` does not exist on the GPU box; nothing in
``tests/``, ``bench.py`` or ``smoke()`` imports this module).  Follows SURVEY.md §8c:

* work-around B: stub the five adapters whose imports need librosa / onnxruntime / inflect so the
  real ``vox_serve.model`` registry, ``vox_serve.worker`` and ``vox_serve.scheduler`` import unmodified;
* transformers 5.x dropped ``config.rope_theta`` -> shim from ``rope_parameters``;
* ``torch.compile`` on ``Sampler`` methods (sampling.py:121,149) is disabled so they run eagerly;
* the three FlashInfer ops are CUDA-only: ``orpheus.rms_norm`` / ``orpheus.apply_rope_pos_ids`` are
  re-pointed at the oracle restatements and a ``PagedWrapperCPU`` is passed as ``attn_wrapper``.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VOX_REFERENCE_ROOT", "/root/reference")


def import_reference():
    os.environ.setdefault("TORCH_COMPILE_DISABLE", "1")
    import torch
    import torch._dynamo  # noqa: F401

    torch._dynamo.config.disable = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    stubs = {
        "qwen3_tts": "Qwen3TTSModel", "cosyvoice2": "CosyVoice2Model", "chatterbox": "ChatterboxModel",
        "step_audio_2": "StepAudio2Model", "zonos": "ZonosModel",
    }
    for mod, cls in stubs.items():
        name = f"vox_serve.model.{mod}"
        if name not in sys.modules:
            m = types.ModuleType(name)
            setattr(m, cls, type(cls, (), {}))
            sys.modules[name] = m
    if "torchaudio" not in sys.modules:
        try:
            import torchaudio  # noqa: F401
        except Exception:
            sys.modules["torchaudio"] = types.ModuleType("torchaudio")
    import vox_serve.model.orpheus as ref_orpheus
    import vox_serve.sampling as ref_sampling
    import vox_serve.tokenizer.snac as ref_snac
    import vox_serve.requests as ref_requests
    import vox_serve.worker.base as ref_worker_base
    import vox_serve.worker.cuda_graph_worker as ref_graph_worker
    import vox_serve.scheduler.base as ref_sched_base

    from . import lm_ops

    ref_orpheus.rms_norm = lambda hidden_states, weight, eps: lm_ops.rms_norm(hidden_states, weight, eps)
    ref_orpheus.apply_rope_pos_ids = (
        lambda query_states, key_states, position_ids, **kw: lm_ops.apply_rope_pos_ids(
            query_states, key_states, position_ids, **kw))
    return types.SimpleNamespace(
        orpheus=ref_orpheus, sampling=ref_sampling, snac=ref_snac, requests=ref_requests,
        worker_base=ref_worker_base, graph_worker=ref_graph_worker, sched_base=ref_sched_base)


def import_reference_cosyvoice2():
    """The reference's ``vox_serve/model/cosyvoice2.py`` itself (``import_reference`` only stubs it): its LM classes
    need nothing beyond torch, but the module imports librosa / onnxruntime / onnx at the top for the audio front end,
    so those three are stubbed with empty modules.  ``rms_norm`` / ``apply_rope_pos_ids`` are re-pointed at the oracle
    restatements as for Orpheus."""
    import importlib

    import_reference()

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {})

    for name in ("librosa", "librosa.filters", "onnxruntime", "onnx"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Any(name)
    sys.modules.pop("vox_serve.model.cosyvoice2", None)
    mod = importlib.import_module("vox_serve.model.cosyvoice2")
    from . import lm_ops

    mod.rms_norm = lambda hidden_states, weight, eps: lm_ops.rms_norm(hidden_states, weight, eps)
    mod.apply_rope_pos_ids = (
        lambda query_states, key_states, position_ids, **kw: lm_ops.apply_rope_pos_ids(
            query_states, key_states, position_ids, **kw))
    return mod


def import_reference_qwen3_tts():
    """The reference's ``vox_serve/model/qwen3_tts.py`` itself (``import_reference`` only stubs it): only librosa is
    missing here, and only its audio front end uses it."""
    import importlib

    import_reference()

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {})

    for name in ("librosa", "librosa.filters"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Any(name)
    sys.modules.pop("vox_serve.model.qwen3_tts", None)
    mod = importlib.import_module("vox_serve.model.qwen3_tts")
    from . import lm_ops

    mod.rms_norm = lambda hidden_states, weight, eps: lm_ops.rms_norm(hidden_states, weight, eps)
    mod.apply_rope_pos_ids = (
        lambda query_states, key_states, position_ids, **kw: lm_ops.apply_rope_pos_ids(
            query_states, key_states, position_ids, **kw))
    return mod
