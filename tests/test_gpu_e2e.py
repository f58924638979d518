"""Worker-level end-to-end parity on the GPU: prefill + decode + SNAC + PCM through the reference-facing worker
API (prepare_lm_inputs / run_lm_prefill / run_lm_decode / run_detokenize / free_kv_cache) against the CPU oracle,
teacher-forced (see tests/e2e_harness.py)."""
import pytest

from oracle import orpheus as oorph

pytestmark = pytest.mark.gpu


def _check(st):
    print(st)
    # every disagreement must be a bf16 near-tie, and there must be few of them
    assert st["id_mismatch"] == st["low_margin"], st
    assert st["id_mismatch"] <= max(2, st["rows"] // 40), st
    # identical tokens + identical noise -> PCM within a few int16 steps (waveform tolerance 1e-3 relative
    # = 32 steps at full scale; fp32 reassociation in the conv stack stays far below that)
    assert st["chunks"] > 0 and st["pcm_max_lsb"] <= 4, st
    assert st["gpu_launches"] > 0


def test_e2e_tiny_ragged_batch():
    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    _check(__import__("tests.e2e_harness", fromlist=["x"]).run_e2e_parity(
        prompt_lens=(9, 16, 30, 33), n_tokens=45, seed=3, dims=dims, page_size=16, max_pages=128))


def test_e2e_page128_gqa3():
    # Orpheus attention geometry (head_dim 128, 3 q heads per kv head, page 128) with a prompt crossing a page
    dims = oorph.OrpheusDims.tiny(hidden_size=1536, num_hidden_layers=3, num_attention_heads=12,
                                  num_key_value_heads=4, intermediate_size=2048, vocab_size=156940,
                                  stop_token_id=128258, audio_id_base=128266)
    _check(__import__("tests.e2e_harness", fromlist=["x"]).run_e2e_parity(
        prompt_lens=(133, 120, 12), n_tokens=36, seed=4, dims=dims, page_size=128, max_pages=16))
