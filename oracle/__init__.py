"""CPU oracle for the VoxServe streaming SpeechLM decode + vocoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vox_serve_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may use it, and only as the checker / the reported CPU baseline.

What it is: a plain torch-CPU / numpy restatement of the arithmetic the reference performs on the
path SURVEY.md §8(a) lists (rows a1-a21), each function citing the reference ``file:line`` it
follows.  The reference (``/root/reference``, vox-serve @ 6f6b469) is pure Python and has NO tests
or golden vectors of its own ("parity unpinned" by the reference).  Pinning is therefore done by
``oracle/gen_golden.py``: it imports the reference's *own* modules in the authoring container
(``OrpheusForCausalLM``, ``Sampler``, ``SNAC``, ``ModelWorker.prepare_lm_inputs/run_detokenize``
rules), runs them on CPU on seeded inputs, and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` holds this restatement to those files.

Pinning status per piece
  * model-level (Orpheus LM forward, greedy ids, repetition penalty / cache update, SNAC decode,
    de-interleave, window / trim / PCM rules): PINNED against the reference's own code run here.
  * op-level third-party arithmetic (FlashInfer ``rmsnorm``, ``apply_llama31_rope_pos_ids``,
    paged decode/prefill attention; flashinfer-python==0.2.11.post1 in the reference's
    ``pyproject.toml:36``, 0.6.11.post2 installed here): CUDA-only, cannot execute in the authoring
    container.  Restated from the installed headers (``include/flashinfer/norm.cuh:64-101``,
    ``pos_enc.cuh:594-617,1538-1539``) and cross-checked against FlashInfer itself on the GPU box
    by ``tests/test_gpu_flashinfer_xcheck.py`` (skipped when FlashInfer's JIT is unavailable): green on the B200
    (5 tests), so these three ops are pinned there.
  * the other BASELINE.json configs' language models (groundwork for the rows SURVEY.md 8f marks "next"; no CUDA
    path serves them yet): ``cosyvoice2.py`` (configs[0]), ``glm_voice.py`` (configs[4]), ``csm.py`` (configs[3],
    backbone + depth-transformer loop), ``qwen3_tts.py`` (configs[2], talker + code predictor) -- each PINNED bit for
    bit to a golden file produced by the reference's own modules on CPU (``gen_golden.py``).
"""
