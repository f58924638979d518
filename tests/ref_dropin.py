"""Import the UNMODIFIED reference package (pip-installed copy under baseline/_ref, git-ignored) with
INTEGRATION.md's import redirection applied: the reference's adapters / schedulers then run on vox_serve_b200's
operator shims (and, for the scheduler tests, on its worker).  Test infrastructure only."""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "vox_serve", "model", "orpheus.py"))


def install_redirects(worker: bool = False):
    """INTEGRATION.md §1 (operators) and, with ``worker``, §2 (worker + model registry)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import vox_serve_b200.flashinfer_utils
    import vox_serve_b200.sampling

    sys.modules["vox_serve.flashinfer_utils"] = vox_serve_b200.flashinfer_utils
    sys.modules["vox_serve.sampling"] = vox_serve_b200.sampling
    if worker:
        import vox_serve_b200.model
        import vox_serve_b200.worker

        sys.modules["vox_serve.worker"] = vox_serve_b200.worker
        sys.modules["vox_serve.model"] = vox_serve_b200.model
    else:
        # SURVEY.md §8c work-around B: the five adapters whose imports need librosa / onnxruntime / inflect are
        # stubbed so that the real vox_serve.model registry imports; Orpheus itself is untouched
        for mod, cls in {"qwen3_tts": "Qwen3TTSModel", "cosyvoice2": "CosyVoice2Model", "chatterbox": "ChatterboxModel",
                         "step_audio_2": "StepAudio2Model", "zonos": "ZonosModel"}.items():
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


def llama_config(dims):
    """transformers LlamaConfig for the reference's OrpheusForCausalLM, with the rope_theta attribute transformers 5.x
    moved into rope_parameters (orpheus.py:62-66 reads config.rope_theta / rope_scaling)."""
    from transformers import LlamaConfig

    hf = LlamaConfig(vocab_size=dims.vocab_size, hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
                     num_hidden_layers=dims.num_hidden_layers, num_attention_heads=dims.num_attention_heads,
                     num_key_value_heads=dims.num_key_value_heads, head_dim=dims.head_dim, rms_norm_eps=dims.rms_norm_eps,
                     hidden_act="silu", attention_bias=False, mlp_bias=False, pad_token_id=None, tie_word_embeddings=False)
    hf.rope_theta = dims.rope_theta
    hf.rope_scaling = {"factor": dims.rope_factor, "low_freq_factor": dims.low_freq_factor,
                       "high_freq_factor": dims.high_freq_factor,
                       "original_max_position_embeddings": dims.old_context_len, "rope_type": "llama3"}
    return hf
