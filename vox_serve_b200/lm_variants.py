"""The decoder stacks of the reference's other adapters on ``LlamaEngine``: the same packed-weight projections, paged
attention, fused reduce / norm / RoPE kernels as Orpheus, configured through ``LlamaDims``:

* **CosyVoice2 speech LM** (BASELINE.json configs[0]; ``vox_serve/model/cosyvoice2.py:26-38, 122-315``): Qwen2-0.5B-shaped
  stack with q / k / v bias, plain rotate-half RoPE at theta 1e6, driven with ``inputs_embeds`` (text-token, task and
  speech-token embeddings), output head ``llm_decoder`` with bias over 6561 + 3 speech ids;
* **GLM-4-Voice decoder** (configs[4]; ``vox_serve/model/glm_voice.py:22-304``): fused biased ``query_key_value``,
  rotation of the first half of every head in (even, odd) pairs, fused ``dense_h_to_4h`` (gate | up), 168 960-row head.

Only the LM side: neither adapter's vocoder (flow + HiFT) is on the CUDA path, so these are engine-level classes, not
registered models.  Parity: tests/test_gpu_lm_variants.py against the golden files produced by the reference's own
``CosyVoice2ForCausalLM`` / ``GLMVoiceForCausalLM`` on CPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from .engine import BF16, LlamaDims, LlamaEngine, LlamaWeights


def cosyvoice2_dims(hidden_size=896, num_hidden_layers=24, num_attention_heads=14, num_key_value_heads=2,
                    intermediate_size=4864, speech_token_size=6561, rms_norm_eps=1e-6, rope_theta=1e6) -> LlamaDims:
    """``CosyVoice2Config`` (cosyvoice2.py:26-38) as engine dims; vocab = the ``llm_decoder`` rows (speech ids + 3 stops)."""
    return LlamaDims(hidden_size, num_hidden_layers, num_attention_heads, num_key_value_heads,
                     hidden_size // num_attention_heads, intermediate_size, speech_token_size + 3, rms_norm_eps, rope_theta,
                     1.0, None, None, None, qkv_bias=True, head_bias=True)


class CosyVoice2LM:
    """prefill with embeddings, decode by feeding ``speech_embedding(id)`` (cosyvoice2.py:934-1024)."""
    PREFIX = "llm.model.model."

    def __init__(self, state_dict: Dict[str, torch.Tensor], dims: LlamaDims, kv_cache: torch.Tensor, page_size: int,
                 max_rows: int = 1024 + 64):
        dev = kv_cache.device
        self.dims = dims
        self.weights = LlamaWeights.from_state_dict(state_dict, dims, dev, prefix=self.PREFIX, embed_key=None,
                                                    head_key="llm_decoder.weight", head_bias_key="llm_decoder.bias")

        def put(k):
            return state_dict[k].to(device=dev, dtype=BF16).contiguous()

        self.speech_embedding = put("speech_embedding.weight")      # [speech ids + 3, H]
        self.llm_embedding = put("llm_embedding.weight")            # [2, H]: sos / task_id
        self.text_embedding = put(self.PREFIX + "embed_tokens.weight")
        self.engine = LlamaEngine(self.weights, kv_cache, page_size, max_rows=max_rows)

    def forward_embeds(self, embeds: torch.Tensor, position_ids: torch.Tensor, last_rows: Optional[torch.Tensor] = None,
                       plan: Optional[ops.RowPlan] = None) -> torch.Tensor:
        """embeds [T, H] bf16 -> logits [T or len(last_rows), speech vocab]."""
        T = embeds.shape[0]
        self.engine.hidden[:T].copy_(embeds)
        return self.engine.forward(None, position_ids, T, last_rows=last_rows, plan=plan)

    def forward_speech_ids(self, ids: torch.Tensor, position_ids: torch.Tensor, plan: Optional[ops.RowPlan] = None):
        """decode rows: ids int32 [B] of the speech tokens sampled by the previous step."""
        B = ids.numel()
        ops.embedding(self.speech_embedding, ids, out=self.engine.hidden[:B])
        return self.engine.forward(None, position_ids, B, plan=plan)


def glm_voice_dims(hidden_size=4096, num_layers=40, num_attention_heads=32, multi_query_group_num=2, ffn_hidden_size=13696,
                   padded_vocab_size=168960, layernorm_epsilon=3.90625e-08, rope_ratio=1.0, rope_theta=10000.0) -> LlamaDims:
    """``GLMVoiceConfig`` (glm_voice.py:22-54) as engine dims."""
    D = hidden_size // num_attention_heads
    return LlamaDims(hidden_size, num_layers, num_attention_heads, multi_query_group_num, D, ffn_hidden_size,
                     padded_vocab_size, layernorm_epsilon, rope_theta, rope_ratio, None, None, None, qkv_bias=True,
                     rotary_dim=D // 2, rope_interleave=True)


def glm_voice_to_llama_names(sd: Dict[str, torch.Tensor], dims: LlamaDims) -> Dict[str, torch.Tensor]:
    """The checkpoint's fused tensors as views under HF-Llama names: ``query_key_value`` rows are [q | k | v]
    (glm_voice.py:123-140), ``dense_h_to_4h`` rows [gate | up] (``torch.chunk(..., 2, dim=-1)``, :196-199)."""
    D, I = dims.head_dim, dims.intermediate_size
    nq, nkv = dims.num_attention_heads * D, dims.num_key_value_heads * D
    out = {"model.embed_tokens.weight": sd["transformer.embedding.word_embeddings.weight"],
           "model.norm.weight": sd["transformer.encoder.final_layernorm.weight"],
           "lm_head.weight": sd["transformer.output_layer.weight"]}
    for i in range(dims.num_hidden_layers):
        s, d = f"transformer.encoder.layers.{i}.", f"model.layers.{i}."
        qkv, qkv_b, gu = sd[s + "self_attention.query_key_value.weight"], sd[s + "self_attention.query_key_value.bias"], \
            sd[s + "mlp.dense_h_to_4h.weight"]
        out[d + "input_layernorm.weight"] = sd[s + "input_layernorm.weight"]
        out[d + "post_attention_layernorm.weight"] = sd[s + "post_attention_layernorm.weight"]
        for name, lo, hi in (("q", 0, nq), ("k", nq, nq + nkv), ("v", nq + nkv, nq + 2 * nkv)):
            out[d + f"self_attn.{name}_proj.weight"], out[d + f"self_attn.{name}_proj.bias"] = qkv[lo:hi], qkv_b[lo:hi]
        out[d + "self_attn.o_proj.weight"] = sd[s + "self_attention.dense.weight"]
        out[d + "mlp.gate_proj.weight"], out[d + "mlp.up_proj.weight"] = gu[:I], gu[I:]
        out[d + "mlp.down_proj.weight"] = sd[s + "mlp.dense_4h_to_h.weight"]
    return out


class GLMVoiceLM:
    def __init__(self, state_dict: Dict[str, torch.Tensor], dims: LlamaDims, kv_cache: torch.Tensor, page_size: int,
                 max_rows: int = 1024 + 64):
        self.dims = dims
        self.weights = LlamaWeights.from_state_dict(glm_voice_to_llama_names(state_dict, dims), dims, kv_cache.device)
        self.engine = LlamaEngine(self.weights, kv_cache, page_size, max_rows=max_rows)

    def forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, last_rows: Optional[torch.Tensor] = None,
                plan: Optional[ops.RowPlan] = None) -> torch.Tensor:
        return self.engine.forward(input_ids, position_ids, input_ids.numel(), last_rows=last_rows, plan=plan)
