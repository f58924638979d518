"""Worker for the multi-codebook LMs with a depth transformer (CSM; ``vox_serve/worker/base.py:430-452, 510-527,
546-614`` + ``cuda_graph_worker.py:1058-1160``): the same five methods and request bookkeeping as ``ModelWorker``, with
a frame (one id per codebook) where the single-codebook worker has a token.

What a step does on the device (``_device_step``; decode steps are ONE CUDA-graph replay per batch size):
  plan -> last frame of every slot -> ``model.frame_device`` (backbone step, codebook 0, the N - 1 depth steps and their
  sampling, all on the device: the reference alternates 32 graph replays with host glue per frame) -> frame fed back
  into the slot state and appended to the slot's frame history -> D2H of the [B, n_codebooks] ids through the ring.
The vocoder pass gathers ``[chunks, interval, n_codebooks]`` windows from the frame history on the device and runs
``model.postprocess`` (Mimi) + PCM16 as one graph replay on the side stream.

Deviations from the reference, both deliberate:
  * a request also ends with ``max_tokens_reached`` when its position passes ``max_tokens``; the reference's CSM adapter
    only ever tests that inside the stop branch (csm.py:716-722), so a request that never samples the stop frame runs
    until the KV pages are gone;
  * watermarking (silentcipher, worker/base.py:683-720) is not applied (third-party post-filter, absent offline).
"""
from __future__ import annotations

from typing import Coroutine, List, Optional

import torch

from .. import ops
from .._lib import VoxB200Error
from ..requests import Request
from .base import CudaGraphWorker, I32

I64 = torch.int64


class DepthModelWorker(CudaGraphWorker):
    _serves_depth = True

    def _prepare_attention_wrappers(self):
        super()._prepare_attention_wrappers()
        m, dev, B = self.model, self.device, self.max_batch_size
        C = m.n_codebooks
        if not hasattr(m, "frame_device"):
            raise VoxB200Error(f"{type(m).__name__} has no frame_device(): the depth worker runs whole frames on the device")
        if m.use_repetition_penalty:
            raise VoxB200Error("repetition penalty over frames is not implemented (no depth model of the reference uses it)")
        # slot-resident frame state (replaces the single-codebook next_input / history of the base class)
        self.frames = torch.zeros(B, C, dtype=I64, device=dev)
        self.history = torch.zeros(B, self.history_cap, C, dtype=I64, device=dev)
        self.n_out = torch.zeros(B, dtype=I64, device=dev)
        # prompt rows of a prefill step: ids + masks [rows, C], staged in pinned memory
        self.prompt_ids_host = torch.zeros(self.max_rows, C, dtype=I64, pin_memory=True)
        self.prompt_mask_host = torch.zeros(self.max_rows, C, dtype=torch.bool, pin_memory=True)
        self.ids_dev = torch.zeros(self.max_rows, C, dtype=I64, device=dev)
        self.mask_dev = torch.zeros(self.max_rows, C, dtype=torch.bool, device=dev)
        self.decode_mask = torch.ones(C, dtype=torch.bool)
        # mask of a decode row: CSM feeds back only the audio streams (csm.py:711-712); Qwen3-TTS rows keep the last column
        # set = "this row carries a codec embedding" (qwen3_tts.py:1941)
        self.decode_mask[-1] = bool(getattr(m, "decode_text_column_mask", False))
        self._win_j = torch.arange(m.detokenize_interval, dtype=I64, device=dev)
        self.voc_win = torch.zeros(self.max_chunks, m.detokenize_interval, C, dtype=I64, device=dev)
        # models whose rows carry input_features (Qwen3-TTS: speaker / ICL sums in the prompt, the predictor-embedding sum of
        # the previous frame in decode, qwen3_tts.py:1835-1853, 2002): per-slot feature row + prompt staging
        self.feats = None
        if m.needs_input_features:
            H = m.hidden_size
            self.feats = torch.zeros(B, H, dtype=torch.bfloat16, device=dev)
            self.prompt_feat_host = torch.zeros(self.max_rows, H, dtype=torch.bfloat16, pin_memory=True)
            self.feat_dev = torch.zeros(self.max_rows, H, dtype=torch.bfloat16, device=dev)
        # codecs with streaming state (Qwen3-TTS: Qwen3TTSDecoderCache): ONE cache over all batch slots; a vocoder call moves
        # the rows of the streams it serves out and back in (DecoderCache.__getitem__ / index_copy_), a stream that starts on
        # a recycled slot gets its rows zeroed.  (The reference keeps one cache per request and cats them per call,
        # cuda_graph_worker.py:1195-1240.)
        init = getattr(m, "audio_decoder_initial_cache", None)
        self.voc_cache = init(B) if init is not None else None

    # ---- token rows ---------------------------------------------------------------------------------------
    def _acquire_slot(self, req: Request) -> int:
        new = req.request_id not in self.slot_of
        s = super()._acquire_slot(req)
        if new and self.voc_cache is not None:
            self.voc_cache.zero_rows_(s)
        return s

    def _stage_prompt_rows(self, req: Request, out, t: int, n: int) -> None:
        if out.input_masks is not None:
            req.input_masks = out.input_masks
        self.prompt_ids_host[t:t + n] = req.input_tokens
        self.prompt_mask_host[t:t + n] = req.input_masks
        if self.feats is not None:
            req.input_features = out.input_features
            self.prompt_feat_host[t:t + n] = out.input_features.to(torch.bfloat16)

    def _stage_decode_row(self, t: int) -> None:
        self.prompt_ids_host[t] = 0              # replaced on the device by the slot's last frame (row_slot >= 0)
        self.prompt_mask_host[t] = self.decode_mask

    def _step_input_ids(self, t: int) -> torch.Tensor:
        return self.ids_dev[:t]

    def _step_input_masks(self, t: int) -> Optional[torch.Tensor]:
        return self.mask_dev[:t]

    # ---- the step -----------------------------------------------------------------------------------------
    def _device_step(self, B: int, T: int, is_prefill: bool):
        st, m = self.staging, self.model
        wrapper = self.prefill_wrapper if is_prefill else self.decode_wrapper
        wrapper.plan_device(st.d("qo", B + 1) if is_prefill else None, st.d("indptr", B + 1), st.d("indices"),
                            st.d("last", B), B, T)
        slots = st.d("slots", B).long()
        if is_prefill:
            self.ids_dev[:T].copy_(self.prompt_ids_host[:T], non_blocking=True)
            self.mask_dev[:T].copy_(self.prompt_mask_host[:T], non_blocking=True)
            st.consumed.record()                 # the pinned prompt rows may be refilled only after these copies
            # rows of requests that are already decoding carry their slot: their ids are the slot's last frame
            rs = st.d("row_slot", T)
            is_dec, rs64 = (rs >= 0)[:, None], rs.clamp(min=0).long()
            ids = torch.where(is_dec, self.frames[rs64], self.ids_dev[:T])
            last_rows = (st.d("qo", B + 1)[1:] - 1).contiguous()
            kw = {}
            if self.feats is not None:
                self.feat_dev[:T].copy_(self.prompt_feat_host[:T], non_blocking=True)
                st.consumed.record()
                kw["input_features"] = torch.where(is_dec, self.feats[rs64], self.feat_dev[:T])
            out = m.frame_device(self.kv_cache, wrapper, st.d("pos", T), B, ids, self.mask_dev[:T], last_rows=last_rows, **kw)
        else:
            kw = {"input_features": self.feats[slots]} if self.feats is not None else {}
            out = m.frame_device(self.kv_cache, wrapper, st.d("pos", T), B, self.frames[slots], **kw)
        if self.feats is not None:
            self.feats.index_copy_(0, slots, m.slot_features(self.kv_cache, B))
        self.out_ids[:B].copy_(out)
        # feedback: the frame becomes the slot's next input and is appended to the slot's history unless it is the stop
        # frame (the host does not append that one to lm_output_audio_tokens either, csm.py:713-722)
        self.frames.index_copy_(0, slots, out)
        at = self.n_out[slots]
        self.history.index_put_((slots, at % self.history_cap), out)
        self.n_out.index_add_(0, slots, (out[:, 0] != m.stop_token_id).to(I64))

    def _capture_decode(self, B: int):
        # the depth decoder's static row plans are built (and synchronised) outside the capture
        self.model.engine_for(self.kv_cache, self.page_size).depth_plans(B)
        return super()._capture_decode(B)

    def _run_step(self, requests: List[Request], lm_inputs) -> Optional[Coroutine]:
        if len(requests) == 0:
            return None
        B, T, is_prefill = len(requests), lm_inputs["n_rows"], lm_inputs["is_prefill"]
        launched = lm_inputs.pop("_launched", None)
        ids_host, ready, _ = launched if launched is not None else self._launch_step(B, T, is_prefill)
        return self._update_req_states(requests, ids_host, ready)

    def _update_req_states(self, requests: List[Request], ids_host: torch.Tensor, ready: torch.cuda.Event):
        """The request bookkeeping of ``sampling`` + ``depth_sampling`` (csm.py:705-724, 764-767) for whole frames, as
        the coroutine the schedulers drive after the step is in flight."""
        m = self.model
        stop, max_tokens, mask = m.stop_token_id, m.max_tokens, self.decode_mask[None, :]

        async def update_req_states():
            ready.synchronize()
            host = ids_host[:len(requests)].clone()
            cb0 = host[:, 0].tolist()
            for i, req in enumerate(requests):
                row = host[i:i + 1]
                req.input_tokens = row.clone()
                if self.feats is None:
                    req.input_tokens[0, -1] = 0            # CSM: the text column of a decode row is unused (csm.py:707-712)
                else:
                    req.input_features = None              # the next row's features live in the slot state on the device
                req.input_masks = mask
                req.lm_output_tokens.append(row)
                if cb0[i] != stop:
                    req.lm_output_audio_tokens.append(row)
                    if req.next_position_id > max_tokens:
                        req.done_lm_generation, req.finish_reason = True, "max_tokens_reached"
                else:
                    req.done_lm_generation, req.finish_reason = True, "stop_id_encountered"

        return update_req_states()

    # ---- vocoder ------------------------------------------------------------------------------------------
    def _vocoder_body(self, n: int):
        interval = self.detokenize_interval
        slot, first, n_valid = (self.win_dev[i, :n].long() for i in range(3))
        # window j of chunk c = frame first + min(j, n_valid - 1): short final windows repeat their last frame
        # (worker/base.py:629-632)
        idx = first[:, None] + torch.minimum(self._win_j[None, :], n_valid[:, None] - 1)
        self.voc_win[:n].copy_(self.history[slot[:, None], idx % self.history_cap])
        if self.voc_cache is not None:
            sub = self.voc_cache[slot]
            audio = self.model.postprocess(self.voc_win[:n], decoder_cache=sub)
            self.voc_cache.index_copy_(slot, sub)
        else:
            audio = self.model.postprocess(self.voc_win[:n])
        ops.pcm16(audio, out=self.voc_pcm[:n])

    def run_lm_decode_resident(self, *a, **kw):
        raise VoxB200Error("the device-resident multi-step loop exists for single-codebook LMs only")
