/*
 * vb_api.h -- C ABI of libvoxb200.so: the B200 (sm_100a) kernels behind VoxServe's operator boundary
 * (level b4).
 *
 * Conventions
 *   - every pointer named d_* / *_dev is a DEVICE pointer; tensors are dense row-major unless a stride
 *     is passed; bf16 = 2-byte bfloat16; ids / page tables are int32 unless stated.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued, nothing synchronises, nothing
 *     allocates: scratch comes from the caller (`*_workspace_bytes` queries), so every entry point is
 *     CUDA-graph capturable and re-entrant (no global mutable state except the thread-local error text).
 *   - return 0 on success, <0 on error; vb_last_error() returns the thread-local message.
 *   - tensor maps: TMA descriptors are 128-byte opaque blobs encoded on the host by vb_tensor_map_*
 *     into caller memory (64-byte aligned) and passed back by pointer; the library keeps no copy.
 *
 * Each group cites the reference interface it replaces (paths relative to the vox-serve tree).
 */
#ifndef VB_API_H_
#define VB_API_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_TENSOR_MAP_BYTES 128

const char* vb_last_error(void);
int vb_version(void);
/* Programmatic dependent launch between the library's kernels (default on; env VB_PDL=0 turns it off).  Returns the
 * previous setting.  With it on, consecutive kernels overlap: a kernel streams its weights / old KV while its
 * predecessor is still running (see csrc/common.cuh). */
int vb_set_pdl(int enabled);
/* SM count / max dynamic smem of the current device (host query used to size persistent grids). */
int vb_device_info(int* sm_count, int* max_smem_optin);
/* dev tool: install (or with NULL remove) a device buffer of u64 words -- [0] = records used (zero it), [1] = capacity,
 * records from word 4 on: {t0, t1 (globaltimer ns), kernel/mark id, aux} -- in which block 0 of the decode-step
 * kernels logs its start / end and a few internal marks: the in-situ timeline of a CUDA-graph replay. */
int vb_set_trace(void* d_buffer);

/* ---- TMA descriptors ------------------------------------------------------------------------ */
/* Whole paged KV cache  [n_slabs = layers*pages][2][page_size][n_kv_heads][head_dim] bf16
 * (layout of vox_serve/worker/base.py:170-179).  Box = [64 dims x box_tokens x all kv heads], 128B swizzle, with the
 * dimensions ordered (dim, token, head, k|v, page) so that it lands in shared memory as [head][token][64]. */
int vb_tensor_map_kv(void* out_map, const void* d_kv, int64_t n_slabs, int page_size, int n_kv_heads,
                     int head_dim, int box_tokens);
/* Row-major [rows][cols] bf16 matrix with leading dimension ld (elements); box = [box_rows x 64], 128B swizzle.
 * Used for GEMM weights (nn.Linear.weight, [out_features][in_features]) and activations. */
int vb_tensor_map_2d_bf16(void* out_map, const void* d_base, int64_t rows, int64_t cols, int64_t ld,
                          int box_rows);

/* ---- RMSNorm: vox_serve/flashinfer_utils.py:251-267 ------------------------------------------ */
/* "Tiled" activation layout XT(t_tile) -- the image of a GEMM activation stage in shared memory, kept in global memory:
 * [token block][k-block of 64][t_tile rows][64] bf16, 16-byte chunk c of row r stored at chunk c ^ (r & 7).  A tile is one
 * contiguous t_tile * 128-byte run, so vb_gemm_bf16 fetches it with ONE linear bulk copy instead of a tensor-map box of
 * t_tile row requests (the TMA unit's issue rate is what bounds the weight stream).  Producers that feed a projection
 * take an `xt_tile` argument: 0 = plain row-major output, else write XT(xt_tile) with xt_tile = vb_gemm_t_tile(T). */
int vb_rmsnorm(void* d_out, const void* d_x, const void* d_weight, int rows, int dim, float eps, int xt_tile,
               void* stream);

/* ---- RoPE at position ids: vox_serve/flashinfer_utils.py:270-324 ------------------------------
 * q [T][n_q][D], k [T][n_kv][D] bf16 -> q_out/k_out (may alias inputs).  d_freq[rotary_dim] is the
 * per-element frequency table; vb_rope_freqs fills it (llama31 != 0 selects the Llama-3.1 smoothing with
 * low/high_freq_factor, old_context_len; interleave selects pairwise (2j,2j+1) rotation). */
int vb_rope_freqs(float* d_freq, int rotary_dim, int interleave, float rope_scale, float rope_theta, int llama31,
                  float low_freq_factor, float high_freq_factor, float old_context_len, void* stream);
int vb_rope(void* d_q_out, void* d_k_out, const void* d_q, const void* d_k, const int32_t* d_pos,
            const float* d_freq, int T, int n_q, int n_kv, int head_dim, int rotary_dim, int interleave,
            void* stream);

/* ---- paged KV bookkeeping: flashinfer_utils.py:86-124 (prefill), :217-225 (decode) -----------
 * Device-side "plan": from the page table (indptr [B+1], indices, last_page_len [B]) and, for prefill,
 * qo_indptr [B+1], derive per query row: owning request, visible kv length, (page, slot) of its new
 * K/V entry; plus the exclusive prefix of chunk_tokens-sized attention tiles per row.
 * qo_indptr == NULL means decode (one row per request).  d_kv_len (optional, [B]) gives each request's kv
 * length explicitly -- then the page table may hold pre-allocated pages beyond it and last_page_len is
 * ignored (device-resident decode loop; the reference derives the length from the table, :96-97).  Rows >= n_rows_valid up to n_rows_padded get
 * page = -1 (their K/V append is skipped; the reference scatters them into page -1).
 * Outputs (device int32): row_req[R], row_kvlen[R], row_page[R], row_slot[R], row_chunk_start[R+1] (exclusive
 * prefix of ceil(kvlen / chunk_tokens): the attention kernel's work list), row_pagebase[R] (= kv_indptr of the
 * row's request, so the attention kernel resolves a tile's page with one load) and row_old[R] (tokens of the
 * request that were already in the cache before this step: tiles below it may be fetched before the kernel that
 * appends the new K/V has finished). */
int vb_plan_rows(const int32_t* d_qo_indptr, const int32_t* d_kv_indptr, const int32_t* d_kv_indices,
                 const int32_t* d_last_page_len, const int32_t* d_kv_len, int n_req, int n_rows_padded, int page_size,
                 int chunk_tokens,
                 int32_t* d_row_req, int32_t* d_row_kvlen, int32_t* d_row_page, int32_t* d_row_slot,
                 int32_t* d_row_chunk_start, int32_t* d_row_pagebase, int32_t* d_row_old, void* stream);

/* kv[page][0][slot] = k ; kv[page][1][slot] = v : flashinfer_utils.py:144-145, 243-244.
 * d_layer_kv points at one layer's [pages][2][page_size][n_kv][D]. */
int vb_kv_append(void* d_layer_kv, const void* d_k, const void* d_v, const int32_t* d_row_page,
                 const int32_t* d_row_slot, int T, int page_size, int n_kv, int head_dim, void* stream);

/* ---- GLM-4-Voice speech tokenizer (Whisper-style VQ encoder, vox_serve/encoder/glm.py:84-323): the stages that are not
 * a GEMM or the block-causal attention.  Token-major [T][C] bf16 activations; every op rounds where the reference's
 * bf16 modules round.
 * vb_add_layernorm: h = bf16(h + delta) in place (delta may be NULL), y = LayerNorm(h) * w + b (y may be NULL: only
 *   the residual add) -- glm.py:195-214.  y_xt_tile > 0: y in the tiled activation layout XT(y_xt_tile).
 * vb_gelu_add: y = bf16(gelu_erf(x)), then y = bf16(y + add) when add != NULL (positions, glm.py:288-295); n elements
 *   in rows of dim; y_xt_tile > 0: y in the tiled layout (y must not alias x).
 * vb_chw_to_rows: out[pad + t][c] = in[c][t], rows [0, pad) zero: the causal convolutions read their taps as
 *   overlapping rows of this buffer (glm.py:84-107) -- no im2col copy.
 * vb_avgpool_rows: out[t] = mean of rows [t k, t k + k) with rows >= T counting as zeros (glm.py:303-313).
 * vb_vq_argmin: ids[t] = first arg-min_n bf16(bf16(c2[n] + |x_t|^2) - 2 acc[t][n]) with acc = x c^T in fp32 (a mode-1
 *   vb_gemm_bf16 over the codebook) -- vector_quantize, glm.py:247-258. */
int vb_add_layernorm(void* d_y, void* d_h, const void* d_delta, const void* d_w, const void* d_b, int rows, int dim,
                     float eps, int y_xt_tile, void* stream);
int vb_gelu_add(void* d_y, const void* d_x, const void* d_add, int64_t n, int dim, int y_xt_tile, void* stream);
int vb_chw_to_rows(void* d_out, const void* d_in, int C, int T, int pad, void* stream);
int vb_avgpool_rows(void* d_out, const void* d_in, int T, int D, int k, void* stream);
int vb_vq_argmin(int64_t* d_ids, const float* d_acc, const void* d_x, const void* d_c2, int T, int N, int D,
                 void* stream);

/* ---- whole-page gather / scatter (prefill-KV hand-off between replicas; the reference pins a
 * request to one replica, launch.py:471-474, and has no such transfer).  d_cache: the whole cache
 * [n_layers][pages_per_layer][page_bytes]; d_staging: [n_layers][n_pages][page_bytes] contiguous (what goes on the
 * wire); d_page_ids [n_pages] int32 (device).  to_cache = 0: staging <- cache (sender), 1: cache <- staging (receiver). */
int vb_copy_pages(void* d_cache, void* d_staging, const int32_t* d_page_ids, int n_pages, int n_layers,
                  int64_t pages_per_layer, int64_t page_bytes, int to_cache, void* stream);

/* ---- paged attention: FlashInferDecodeWrapper.run / FlashInferPrefillWrapper.run
 * (flashinfer_utils.py:132, 228-230).  One query row per entry of the plan; causal by construction
 * (row kv length).  q/out [R][n_q][D] bf16.  d_kv = base of the WHOLE cache [slabs][2][page][n_kv][D];
 * slab_base = layer * pages_per_layer.  d_row_kvlen / d_row_chunk_start / d_row_pagebase / d_row_old come from vb_plan_rows
 * with chunk_tokens = vb_attn_tile_tokens(page_size, n_kv); d_kv_indices is the page table the plan was made from.
 * The linearised (row, tile) list is cut into grid_ctas equal ranges (one CTA per SM is the default).
 * d_workspace: vb_paged_attn_workspace_bytes(max_rows, ws_grid_ctas, ...) bytes, zero-filled once before first
 * use (arrival counters; the kernel restores them to zero); grid_ctas <= ws_grid_ctas. */
/* tile size (tokens) the attention kernel and the plan must use for this geometry: 16 * (8 / n_kv), reduced
 * until it divides page_size */
int vb_attn_tile_tokens(int page_size, int n_kv);
size_t vb_paged_attn_workspace_bytes(int max_rows, int max_grid_ctas, int n_q, int n_kv, int head_dim);
int vb_paged_attn(void* d_out, const void* d_q, const void* d_kv, int64_t slab_base,
                  const int32_t* d_row_kvlen, const int32_t* d_row_chunk_start, const int32_t* d_row_pagebase,
                  const int32_t* d_row_old, const int32_t* d_kv_indices, int n_rows, int n_q, int n_kv, int head_dim, int page_size,
                  int chunk_tokens, float sm_scale, void* d_workspace, size_t workspace_bytes, int grid_ctas,
                  int ws_grid_ctas, int out_xt_tile, void* stream);

/* ---- tiled paged attention for prefill-shaped steps: FlashInferPrefillWrapper.run (flashinfer_utils.py:68-80,
 * 132; causal=True).  Same inputs and result as vb_paged_attn on a prefill plan, but a Q tile of up to
 * vb_prefill_attn_tile_rows(n_q, n_kv) consecutive rows of ONE request shares every K/V tile it reads (mma.sync
 * tensor-core tiles, FlashAttention-2 schedule), so a request's K/V is read once per tile instead of once per row.
 * d_qo_indptr / d_kv_indptr [n_req + 1] and d_kv_indices: the step's page table as given to vb_plan_rows;
 * d_row_kvlen [n_rows]: keys visible to each row (vb_plan_rows' output: the causal prefix; any per-row bound that
 * does not decrease inside a request is honoured as is).  Rows >= qo_indptr[n_req] (graph padding) are zeroed.
 * No workspace, no split-KV: a CTA owns (Q tile, kv head, <= 8 grouped query heads). */
int vb_prefill_attn_tile_rows(int n_q, int n_kv);
int vb_paged_prefill_attn(void* d_out, const void* d_q, const void* d_kv, int64_t slab_base,
                          const int32_t* d_qo_indptr, const int32_t* d_kv_indptr, const int32_t* d_kv_indices,
                          const int32_t* d_row_kvlen, int n_req, int n_rows, int n_q, int n_kv, int head_dim,
                          int page_size, float sm_scale, int out_xt_tile, void* stream);

/* The same operator on the 5th-generation tensor cores: both contractions as tcgen05.mma from swizzled shared-memory
 * operand tiles, S and the per-tile P V product in TMEM, the GQA group folded into the MMA's 128 accumulator rows (a Q
 * tile = 128 / group prompt rows of one request), K/V tiles of 128 tokens (any page size).  Arguments as
 * vb_paged_prefill_attn. */
int vb_prefill_attn_tc_tile_rows(int n_q, int n_kv);
int vb_paged_prefill_attn_tc(void* d_out, const void* d_q, const void* d_kv, int64_t slab_base,
                             const int32_t* d_qo_indptr, const int32_t* d_kv_indptr, const int32_t* d_kv_indices,
                             const int32_t* d_row_kvlen, int n_req, int n_rows, int n_q, int n_kv, int head_dim,
                             int page_size, float sm_scale, int out_xt_tile, void* stream);

/* ---- dense projections (nn.Linear, bias-free): model/orpheus.py:41-47, 68-79, 197 -------------
 * Y[T][N] = X[T][K] * W[N][K]^T on tcgen05: a tile of tile_rows (<= 128, multiple of 8; 0 = 128) weight rows is
 * the M side of the MMA, the tokens are the N side.  tile_rows is free so that a projection can be cut into
 * ~one CTA per SM whatever its N.
 * d_w_tiles: the weight matrix re-tiled once at load time by vb_pack_weight_tiles(W, N, K, ldw, tile_rows) into
 * [n_tile][k_block][tile_rows][64] bf16 with the UMMA 128-byte swizzle applied, so that the (tile, k-block) operand
 * a CTA needs is ONE contiguous run in HBM (a single linear bulk copy per pipeline stage, sequential DRAM bursts)
 * and a CTA's k-blocks follow each other; vb_weight_tiles_bytes gives the size (rows padded to the tile, K to 64).
 * x_map: vb_tensor_map_2d_bf16(X, T, K, ldx, t_tile) with t_tile = vb_gemm_t_tile(T).
 * mode 0: Y bf16 [T][ldy]           (split_k must be 1)
 * mode 1: Y fp32 partials [split_k][T][ldy]   (vb_reduce_residual_rmsnorm / vb_qkv_rope_append sum them in split order)
 * mode 2: W rows packed per tile (tile_rows = 32, 64, 96 or 128) as [16 gate rows][the 16 matching up rows] per
 *         epilogue warp -- gate and up of one output sit 16 lanes apart in the same accumulator quarter, so the
 *         product is one warp shuffle, no shared-memory exchange -- (N = packed rows, zero-padded to a multiple of
 *         tile_rows); Y bf16 [T][ldy] holds silu(gate)*up with the reference's bf16 rounding points
 *         (orpheus.py:46-48) for the first n_out (0 = N/2) outputs. */
size_t vb_weight_tiles_bytes(int N, int K, int tile_rows);
int vb_pack_weight_tiles(void* d_dst, const void* d_w, int N, int K, int64_t ldw, int tile_rows, void* stream);
int vb_gemm_t_tile(int T);
/* shared-memory budget of the projection kernel's operand ring in KiB (0 = keep): ring_kb for every projection (default
 * 104: two CTAs per SM, so a programmatic dependent sits beside its predecessor), gate_up_ring_kb for the gate/up
 * projection (default 200: the long stream gets the whole SM).  Process-wide tuning knobs (env VB_GEMM_SMEM_KB[_GU]). */
int vb_set_gemm_smem_kb(int ring_kb, int gate_up_ring_kb);
/* d_x_tiles (optional): X in the XT(vb_gemm_t_tile(T)) layout -- then x_map may be NULL; y_tiled (mode 2 only): write
 * Y in the XT(vb_gemm_t_tile(T)) layout over n_out columns (it is the down projection's activation). */
int vb_gemm_bf16(void* d_y, const void* d_w_tiles, const void* x_map, const void* d_x_tiles, int T, int N, int K,
                 int ldy, int mode, int split_k, int tile_rows, int n_out, int y_tiled, const void* d_bias, void* stream);
/* (d_bias: optional bf16 [N], mode 0 only: y = bf16(x w^T + bias), one rounding -- nn.Linear with bias, e.g.
 * Qwen3-TTS small_to_mtp_projection, vox_serve/model/qwen3_tts.py:923-930) */

/* ---- fused decode projections (T <= 64): one launch each for what orpheus.py:81-151 does between Linears ----
 * x_map: vb_tensor_map_2d_bf16(X, T, K, ldx, t_tile) with t_tile = 16 / 32 / 64 (smallest >= T).
 * Split-K sums finish inside the kernel: the split_k (<= 8) CTAs of a tile form a thread-block cluster and add
 * their partial tiles through distributed shared memory in split order (deterministic); no workspace.
 *
 * vb_proj_residual: hidden_out[T][N] = bf16(residual + bf16(X W^T))   (o_proj / down_proj + residual add,
 *   orpheus.py:139-150; d_residual may alias d_hidden_out, may be NULL) and, if d_ssq_out, the per-tile sums of
 *   squares of the new hidden rows: ssq_out[tile][T] -- the RMSNorm statistics of the NEXT projection.
 * vb_proj_norm_gateup_silu: act[T][n_out] = silu(gate(xn)) * up(xn), xn = rmsnorm(hidden) * norm_weight formed
 *   on the fly: TMA delivers the raw hidden tile (x_map over hidden [T][K]), the kernel scales it in place with the
 *   row statistics d_ssq [n_ssq_parts][T] (flashinfer norm.cuh rounding); weights packed as for mode 2.
 * vb_proj_norm_qkv_rope_append: q|k|v = Wqkv xn (rows: q heads, k heads, v heads; one head per tile), RoPE on
 *   q and k with the step's cos/sin table (vb_rope_table, rotate-half pairs, full head_dim), q -> q_out
 *   [T][n_q][D], k,v -> the page/slot of each row (row_page < 0: skipped).  orpheus.py:91-106,
 *   flashinfer_utils.py:243-244. */
/* Activations: x_map (row-major, tensor-map boxes) or d_x_tiles (the XT(t_tile) layout, bulk copies; preferred) --
 * one of the two.  d_hidden_tiles_out (optional): a second copy of the new hidden rows in the XT layout, i.e. the
 * next norm-fused projection's d_x_tiles; y_tiled: write the gate/up product in the XT layout. */
int vb_proj_residual(void* d_hidden_out, void* d_hidden_tiles_out, float* d_ssq_out, const void* d_w_tiles,
                     const void* x_map, const void* d_x_tiles, const void* d_residual, int T, int N, int K, int split_k,
                     int tile_rows, void* stream);
int vb_proj_norm_gateup_silu(void* d_act_out, const void* d_w_tiles, const void* x_map, const void* d_x_tiles,
                             const float* d_ssq, int n_ssq_parts, const void* d_norm_weight, float eps, int T,
                             int N_packed, int K, int tile_rows, int n_out, int y_tiled, void* stream);
int vb_proj_norm_qkv_rope_append(void* d_q_out, void* d_layer_kv, const void* d_w_tiles, const void* x_map,
                                 const void* d_x_tiles, const float* d_ssq, int n_ssq_parts, const void* d_norm_weight, float eps,
                                 const float* d_rope_cs, const int32_t* d_row_page, const int32_t* d_row_slot, int T,
                                 int K, int n_q, int n_kv, int head_dim, int page_size, int split_k, void* stream);

/* final norm -> lm_head as one operator (orpheus.py:193-197, 219-221: `self.norm(hidden)` then `self.lm_head`):
 * logits[T][ldy] = bf16(rmsnorm(hidden[T][K]) * norm_weight . W^T (+ bias)).  d_workspace holds the normed rows in the
 * tiled activation layout (vb_norm_lmhead_workspace_bytes(T, K) bytes); two launches, the projection a programmatic
 * dependent of the norm.  (The decode engine gets the final norm for free from the last layer's
 * vb_reduce_residual_rmsnorm; this entry point is the operator for callers that hold a plain hidden state.) */
size_t vb_norm_lmhead_workspace_bytes(int T, int K);
int vb_norm_lmhead(void* d_logits, const void* d_hidden, const void* d_norm_weight, float eps, const void* d_w_tiles,
                   const void* d_bias, int T, int N, int K, int ldy, int tile_rows, void* d_workspace,
                   size_t workspace_bytes, void* stream);

/* cs[T][2][head_dim] = cos | sin of pos[t] * freq[e]: computed once per step, shared by all layers */
int vb_rope_table(float* d_cs, const int32_t* d_pos, const float* d_freq, int T, int head_dim, void* stream);
/* ssq[rows] = sum of squares of each bf16 row (RMSNorm statistics of a hidden state that did not come out of
 * vb_proj_residual: the embedding output) */
int vb_row_ssq(float* d_ssq, const void* d_x, int rows, int dim, void* stream);

/* sum split-K partials -> bf16 Linear output; + residual; then RMSNorm of the new hidden state:
 * orpheus.py:125-151 (residual adds, next layer's input_layernorm / post_attention_layernorm).
 * hidden_out = bf16(residual + bf16(sum_s partial[s])) ; normed_out = rmsnorm(hidden_out) * weight.
 * d_residual may be NULL (no add); d_norm_weight may be NULL (skip the norm output). */
int vb_reduce_residual_rmsnorm(void* d_hidden_out, void* d_normed_out, const float* d_partials, int split_k,
                               const void* d_residual, const void* d_norm_weight, int T, int N, float eps,
                               int normed_xt_tile, void* stream);
/* fused tail of the QKV projection: sum partials -> bf16 q|k|v, RoPE(q,k), write q, scatter k,v into the
 * layer cache (orpheus.py:91-106 + flashinfer_utils.py:243-244).  partials [split_k][T][(n_q+2 n_kv) D]. */
int vb_qkv_rope_append(void* d_q_out, void* d_layer_kv, const float* d_partials, int split_k, const int32_t* d_pos,
                       const float* d_freq, const int32_t* d_row_page, const int32_t* d_row_slot, int T, int n_q,
                       int n_kv, int head_dim, int page_size, int rotary_dim, int interleave, const void* d_q_norm,
                       const void* d_k_norm, float norm_eps, const void* d_qkv_bias, void* stream);
/* (d_q_norm / d_k_norm: optional bf16 [head_dim] weights of a per-head RMSNorm applied to every q and k head between the
 * bf16 rounding of the projection and the rotation -- Qwen3's q_norm / k_norm, vox_serve/model/qwen3_tts.py:603-625;
 * d_qkv_bias: optional bf16 [(n_q + 2 n_kv) head_dim] bias of the q | k | v projections, added in fp32 before the bf16
 * rounding -- CosyVoice2's q/k/v_proj (model/cosyvoice2.py:139-143) and GLM-4-Voice's fused query_key_value
 * (model/glm_voice.py:123-140); rotary_dim < head_dim / interleave = 1: GLM's rotation of the first half of every head in
 * (even, odd) pairs, model/glm_voice.py:148-156) */
/* slot-resident decode state (no host work between CUDA-graph replays; the reference does this bookkeeping
 * in Python every step, worker/base.py:312-325, orpheus.py:447-458):
 *   vb_decode_advance:  kv_len[b] += 1, position[b] += 1 for active rows (d_active NULL = all);
 *   vb_token_feedback:  sampled id of batch row b (int64) -> slot s = d_slots[b] (NULL = b):
 *                       next_input[s] = id; unless id == skip_token (the stop id, -1 = none: it is fed back but is
 *                       not an audio token, orpheus.py:461-463): history[s][n_out[s] % cap] = id; ++n_out[s];
 *   vb_gather_i32:      out[i] = src[idx[i]] (next step's input ids by slot);
 *   vb_build_input_ids: out[i] = row_slot[i] >= 0 ? next_input[row_slot[i]] : host_ids[i]  (decode rows feed
 *                       back the id sampled last step, prefill rows use the uploaded prompt ids; replaces the
 *                       torch.cat of worker/base.py:329);
 *   vb_latest_window:   first[i] = max(0, n_out[slot[i]] - window)  (start of the newest full window);
 *   vb_gather_windows:  windows[i][j] = history[slot[i]][(first[i] + min(j, n_valid[i]-1)) % cap], j < window
 *                       (the detokenize window of cuda_graph_worker.py:1176-1190 incl. last-token padding). */
int vb_decode_advance(int32_t* d_kv_len, int32_t* d_pos, const int32_t* d_active, int B, void* stream);
int vb_token_feedback(const int64_t* d_ids, const int32_t* d_slots, int32_t* d_next_input, int32_t* d_history,
                      int32_t* d_n_out, int B, int history_cap, int skip_token, void* stream);
int vb_gather_i32(int32_t* d_out, const int32_t* d_src, const int32_t* d_idx, int n, void* stream);
int vb_build_input_ids(int32_t* d_out, const int32_t* d_host_ids, const int32_t* d_next_input,
                       const int32_t* d_row_slot, int n, void* stream);
int vb_latest_window(int32_t* d_first, const int32_t* d_n_out, const int32_t* d_slot, int n, int window,
                     void* stream);
int vb_gather_windows(int64_t* d_windows, const int32_t* d_history, const int32_t* d_slot, const int32_t* d_first,
                      const int32_t* d_n_valid, int n, int history_cap, int window, void* stream);
/* embedding gather: orpheus.py:408 */
int vb_embedding(void* d_out, const void* d_table, const int32_t* d_ids, int T, int dim, int vocab, void* stream);
/* Multi-codebook glue of the depth-transformer adapters (vox_serve/model/csm.py:158-168, 637-663, 700-701;
 * cuda_graph_worker.py:1058-1160 does these with torch ops and host indices between graph replays):
 *   vb_multi_embed_sum: out[t] = bf16(sum over columns c < C with mask[t][c] != 0 (NULL = all) of
 *                       c < n_cols_a ? table_a[ids(t,c) + (col0 + c) * col_offset] : table_b[ids(t,c)]), fp32
 *                       accumulation in column order; ids int64 at d_ids[t * ld_t + c * ld_c];
 *   vb_interleave_rows: out[2r] = a[r], out[2r+1] = b[r] (the depth decoder's [hidden, embed(cb0)] 2-row prefill);
 *   vb_transpose_i64:   dst[r][c] = src[c][r] (codebook-major device frame buffer -> request-major ids). */
int vb_multi_embed_sum(void* d_out, int ld_out, const int64_t* d_ids, int64_t ld_t, int64_t ld_c, const uint8_t* d_mask,
                       const void* d_table_a, int64_t rows_a, int64_t col_offset, int col0, int n_cols_a,
                       const void* d_table_b, int64_t rows_b, int T, int C, int dim, int round_each, void* stream);
/* (round_each != 0: the running sum is rounded to bf16 after every column -- a chain of bf16 `+=`, qwen3_tts.py:2002)
 * vb_talker_embed: Qwen3-TTS talker input (qwen3_tts.py:1835-1853):
 *   out[t] = bf16((needs_codec[t] ? bf16(text[t] + codec[cb0[t]]) : text[t]) + features[t]); text row stride ld_text
 *   (0 = broadcast one row), cb0 int64 with stride ld_id, needs_codec NULL = all, features NULL = none. */
int vb_talker_embed(void* d_out, int ld_out, const void* d_text, int64_t ld_text, const void* d_codec, int64_t codec_rows,
                    const int64_t* d_cb0, int64_t ld_id, const uint8_t* d_needs_codec, const void* d_features,
                    int64_t ld_feat, int T, int dim, void* stream);
int vb_interleave_rows(void* d_out, const void* d_a, const void* d_b, int n, int row_bytes, void* stream);
int vb_transpose_i64(int64_t* d_dst, const int64_t* d_src, int B, int C, int ld_dst, int ld_src, void* stream);
/* rows out[i] = in[idx[i] + idx_offset] (last-token gather qo_indptr[1:] - 1, cuda_graph_worker.py:900-902) */
int vb_gather_rows(void* d_out, const void* d_in, const int32_t* d_idx, int n, int row_bytes, int idx_offset,
                   void* stream);

/* ---- sampler: vox_serve/sampling.py (whole file) ---------------------------------------------
 * logits [rows][vocab] bf16 with leading dimension ld_logits, rows = batch * logit_codebooks.
 * d_rep_cache: uint8 (torch.bool) repetition cache [batch][rep_window_slots][rep_codebooks][vocab] or NULL;
 * d_cache_rows (optional int32 [batch]): batch row b uses cache row d_cache_rows[b] (caches kept resident per
 * batch slot instead of being re-stacked every step as worker/base.py:345 does);
 * a token is "seen" if any window slot has it (sampling.py:137); if logit_codebooks == 1 and
 * rep_codebooks != 1 codebook 0 is used (sampling.py:140-141).  Penalty: sampling.py:143-144.
 * strategy: 0 greedy (argmax, first index on ties), 1 top-k, 2 top-p, 3 top-k then top-p, 4 min-p
 * (dispatch order of sampling.py:97-118 is applied by the host wrapper).  out_ids int64.
* Draws use Philox4x32-10(seed, offset, row, round); if d_rng_state (device uint64[3] = {seed, offset, 0}; the
 * third word is an arrival counter the kernel returns to 0) is given it overrides the immediates and its offset is
 * advanced by one per call (CUDA-graph replays stay random).  One launch: a thread-block cluster per row keeps the
 * row on chip (vocab <= 327680).  mask_token >= 0 forces that token's logit to -inf
 * (benchmark hook to pin sequence lengths; -1 = off).  Workspace: vb_sample_workspace_bytes. */
size_t vb_sample_workspace_bytes(int rows, int vocab);
int vb_sample(int64_t* d_out_ids, const void* d_logits, int rows, int vocab, int ld_logits,
              const uint8_t* d_rep_cache, const int32_t* d_cache_rows, int rep_window_slots, int rep_codebooks,
              int logit_codebooks,
              float penalty, int strategy, int top_k, float top_p, float min_p, float temperature, uint64_t seed,
              uint64_t offset, uint64_t* d_rng_state, int mask_token, void* d_workspace, size_t workspace_bytes,
              void* stream);
/* penalised logits only (Sampler.apply_repetition_penalty, sampling.py:120-146), bf16 in/out, dense rows */
int vb_apply_repetition_penalty(void* d_out, const void* d_logits, const uint8_t* d_rep_cache,
                                int rep_window_slots, int rep_codebooks, int logit_codebooks, float penalty,
                                int rows, int vocab, void* stream);
/* cache[b][w][c][ids] = 1 with the reference's batch-union semantics (sampling.py:148-178).
 * cache uint8 [B][W][C][V]; ids int64 [B][C_ids]; window > 1 shifts the window first. */
int vb_update_repetition_cache(uint8_t* d_cache, const int32_t* d_cache_rows, const int64_t* d_ids, int B, int W,
                               int C, int V, int C_ids, int window, void* stream);

/* ---- SNAC decoder: vox_serve/tokenizer/snac.py:119-267, 297-357, 438-441 ----------------------
 * fp32, activations [B][C][T] (T contiguous).  Weights arrive weight-norm-folded (snac.py:244-249) --
 * see vox_serve_b200/tokenizer/snac.py for the fold / repack done once at load.  Every Snake (snac.py:252-258)
 * is fused into a neighbouring stage through the optional alpha_in / alpha_out [C] pointers. */
int vb_snac_from_codes(float* d_z, const int32_t* d_codes0, const int32_t* d_codes1, const int32_t* d_codes2,
                       const float* d_codebooks /*[3][cb_size][cb_dim]*/, const float* d_proj_w /*[3][C][cb_dim]*/,
                       const float* d_proj_b /*[3][C]*/, int B, int C, int T, int cb_size, int cb_dim, int stride0,
                       int stride1, int stride2, void* stream);
/* Every stage takes the half-open range of OUTPUT positions to compute (buffers keep their full [.][T] shape):
 * OrpheusModel.postprocess keeps samples [2048, 4096) of 8192 (orpheus.py:506), so each stage only needs the
 * receptive field of that slice -- the same values as the full computation at about a third of the work. */
/* y = snake_out?( dwconv_k7_dilated( snake_in?(x) ) + bias ) on t in [t_lo, t_hi);  w [C][7] */
int vb_snac_dwconv7(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_alpha_in,
                    const float* d_alpha_out, int B, int C, int T, int dilation, int t_lo, int t_hi, void* stream);
/* pointwise conv  v = sum_ci W[co][ci] x[b][ci][t] (+ bias[co]);
 * epilogue 0: v ; 1: v + resid[b][co][t] (ResidualUnit, snac.py:170-176) ; 2: x[b][co][t] + noise[b][t] * v
 * (NoiseBlock, snac.py:206-212, noise supplied by the caller) ; then snake_out?(.) */
int vb_snac_pwconv(float* d_y, const float* d_x, const float* d_w /*[Cout][Cin]*/, const float* d_bias,
                   const float* d_resid, const float* d_noise, const float* d_alpha_out, int epilogue, int B,
                   int Cin, int Cout, int T, int t_lo, int t_hi, void* stream);
/* y = snake_out?( conv_transpose1d(x) + bias ), kernel 2*stride, padding ceil(stride/2), output_padding
 * stride%2 (snac.py:222-231); d_w_packed [stride][Cout][2*Cin] with
 * w_packed[r][co][tap*Cin + ci] = W_torch[ci][co][r + tap*stride];  y [B][Cout][T*stride]; outputs at least
 * [o_lo, o_hi) are written */
int vb_snac_convtr(float* d_y, const float* d_x, const float* d_w_packed, const float* d_bias,
                   const float* d_alpha_out, int B, int Cin, int Cout, int T, int stride, int o_lo, int o_hi,
                   void* stream);
/* Tensor-core variants of the two GEMM-shaped stages (tcgen05 kind::tf32, every fp32 operand split hi + lo and
 * the product summed as lo*hi + hi*lo + hi*hi into fp32 TMEM accumulators: relative error ~2^-21 per product, fp32
 * storage everywhere).  Same arguments and results as vb_snac_pwconv / vb_snac_convtr except that the weights come
 * packed by vb_snac_pack_tf32x3: d_w [phases][M][K] fp32 (pointwise: phases = 1, M = Cout, K = Cin; transposed:
 * the [stride][Cout][2*Cin] array of vb_snac_convtr) -> [phase][m_tile][k_block][hi|lo][128][32] fp32, rows of 128
 * bytes with the 128-byte UMMA swizzle.  K (Cin) must be a multiple of 32; the caller keeps the SIMT entry points
 * for narrower layers.  vb_snac_tf32x3_bytes returns the packed size (-1 for an unsupported shape). */
long long vb_snac_tf32x3_bytes(int phases, int M, int K);
int vb_snac_pack_tf32x3(void* d_dst, const float* d_w, int phases, int M, int K, void* stream);
int vb_snac_pwconv_tc(float* d_y, const float* d_x, const void* d_w_tiles, const float* d_bias, const float* d_resid,
                      const float* d_noise, const float* d_alpha_out, int epilogue, int B, int Cin, int Cout, int T,
                      int t_lo, int t_hi, void* stream);
int vb_snac_convtr_tc(float* d_y, const float* d_x, const void* d_w_tiles, const float* d_bias,
                      const float* d_alpha_out, int B, int Cin, int Cout, int T, int stride, int o_lo, int o_hi,
                      void* stream);
/* y[b][t - t0] = tanh(conv_k7(snake_in?(x))[t] + bias), Cout = 1 (snac.py:152-156), t in [t0, t1) */
int vb_snac_final(float* d_y, const float* d_x, const float* d_w /*[C][7]*/, const float* d_bias,
                  const float* d_alpha_in, int B, int C, int T, int t0, int t1, void* stream);
/* (audio * 32767) truncated toward zero to int16, no clipping: cuda_graph_worker.py:1252-1253 */
int vb_pcm16(int16_t* d_out, const float* d_audio, int64_t n, void* stream);

/* ---- Mimi decode stages (CSM's vocoder: vox_serve/tokenizer/mimi.py:2993-3018), fp32, activations [B][C][T] ----------
 * Each chunk is decoded with zero left context, as the reference's stateless StreamingConv1d / ConvTranspose1d do
 * (mimi.py:2116-2148, 2192-2215).
 *   vb_mimi_codes_sum:  z[b][d][t] = sum_{k0 <= k < k1} emb[k][codes[b][k][t]][d]  (ResidualVectorQuantization.decode
 *                       :482-490; emb [K_total][bins][D] = embedding_sum / clamp(cluster_usage, eps), codes int64
 *                       [B][K_total][T]);
 *   vb_mimi_conv:       causal Conv1d with kernel ksize / dilation (left zero pad (ksize-1) * dilation), weight
 *                       [Cout][Cin][ksize]; ksize 1 = the transformer's Linear layers and the 1x1 projections.
 *                       elu_in: ELU on the input while it is fetched (SEANet puts ELU in front of every conv :2343-2400).
 *                       epilogue 0: y = conv (+ bias); 1: y = resid + conv (+ bias) (SEANetResnetBlock true skip; the
 *                       sum of the two RVQ branches :830-836); 2: y = resid + scale[co] * conv (LayerScale :1097-1129);
 *                       3: y = gelu(conv) (exact erf GELU of the transformer FFN :1687-1703);
 *   vb_mimi_convtr:     causal ConvTranspose1d kernel 2 s / stride s, rightmost K - S outputs dropped (:2192-2215), weight
 *                       re-packed per output phase [s][Cout][2 Cin] (w[r][co][tap Cin + ci] = W[ci][co][r + tap s]);
 *   vb_mimi_upsample:   the learnt channel-wise x s ConvTrUpsample1d (:2272-2323), weight [C][2 s];
 *   vb_mimi_layernorm:  nn.LayerNorm over the channel axis of [B][C][T] (the [B, T, C] view of :1705-1712);
 *   vb_mimi_attention:  causal self-attention of a chunk (T <= 64) with interleaved-pair RoPE at offset 0 (:874-930) on
 *                       qkv [B][3 C][T] packed "(p h d)" (:1519-1523) -> [B][C][T]. */
int vb_mimi_codes_sum(float* d_z, const int64_t* d_codes, const float* d_emb, int B, int K_total, int k0, int k1, int bins,
                      int D, int T, void* stream);
int vb_mimi_conv(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_resid,
                 const float* d_scale, int epilogue, int elu_in, int B, int Cin, int Cout, int T, int ksize, int dilation,
                 void* stream);
int vb_mimi_convtr(float* d_y, const float* d_x, const float* d_w_packed, const float* d_bias, int elu_in, int B, int Cin,
                   int Cout, int T, int stride, void* stream);
int vb_mimi_upsample(float* d_y, const float* d_x, const float* d_w, int B, int C, int T, int stride, void* stream);
int vb_mimi_layernorm(float* d_y, const float* d_x, const float* d_w, const float* d_bias, int B, int C, int T, float eps,
                      void* stream);
int vb_mimi_attention(float* d_out, const float* d_qkv, int B, int C, int H, int T, float max_period, void* stream);
/* n standard-normal fp32 draws: the NoiseBlock input the reference takes from torch.randn
 * (vox_serve/tokenizer/snac.py:206-212).  Philox4x32-10 + Box-Muller, counter = (offset, element / 4).
 * d_rng_state (optional, device u64 {seed, offset, arrivals}) replaces seed / offset and its offset advances by one
 * per call on the device, so a CUDA-graph replay draws fresh noise without host involvement. */
int vb_randn(float* d_out, int64_t n, uint64_t seed, uint64_t offset, uint64_t* d_rng_state, void* stream);
/* LM ids [B][28] -> SNAC codes (orpheus.py:479-500): codes0 [B][4], codes1 [B][8], codes2 [B][16] int32 */
int vb_orpheus_window_codes(int32_t* d_c0, int32_t* d_c1, int32_t* d_c2, const int64_t* d_ids, int B,
                            int audio_id_base, void* stream);

/* ---- streaming codec decoder stages with per-request caches: the Qwen3-TTS 12 Hz decoder ------------------------------
 * vox_serve/tokenizer/qwen3_codec.py:1541-1667 (forward_chunk).  fp32, activations [B][C][T].  act_in: 0 none, 1 ELU,
 * 2 SnakeBeta with d_act_a = exp(alpha), d_act_ib = 1 / (exp(beta) + 1e-9) per input channel (:1004-1018).  The caches hold
 * ACTIVATED inputs (as the reference's do), context reads are not activated again.
 *   vb_codec_conv:   causal Conv1d / Linear (ksize 1); d_ctx [B][Cin][(ksize-1) dilation] = left context (NULL: zeros,
 *                    :274-340); epilogue 0 plain, 1 resid + v, 2 resid + scale[m] v (LayerScale / ConvNeXt gamma), 3 GELU,
 *                    4 SiLU, 5 resid * v (the gated MLP's product), 6 clamp(-1, 1) (:1667); bias added before the epilogue.
 *   vb_codec_convtr: causal ConvTranspose1d, kernel 2 s, stride s, weights packed [s][Cout][2 Cin] (tap 0 | tap 1);
 *                    d_ctx [B][Cin][1] = the previous chunk's last activated input (NULL: zero, :359-397).  A transposed
 *                    convolution with kernel == stride is the same call with a zero tap-1 half.
 *   vb_codec_cache_update: cache [B][C][pad] <- last pad activated inputs after appending x [B][C][L] (:318-325, 391); run
 *                    it AFTER the convolution that reads the old cache.
 *   vb_codec_dwconv: depthwise causal Conv1d with left context (ConvNeXt dwconv, :434-451); w [C][ksize].
 *   vb_codec_rmsnorm: RMSNorm over the channel axis (:713-718).
 *   vb_codec_attn_chunk: one chunk of the sliding-window attention (:573-655): qkv [B][(H + 2 Hkv) D][T], rotate-half RoPE at
 *                    positions d_pos0[b] + t, cache [Hkv][W][2 D] per item (items cache_batch_stride floats apart: one layer of
 *                    the reference's [B][layers][Hkv][W][2 D] tensor) shifted left by T with the new K | V appended, every
 *                    query attends to all W slots (never-written slots hold zeros, as in the reference) under the mask
 *                    j <= W - T + t; out [B][H D][T]; T < W. */
int vb_codec_conv(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_resid,
                  const float* d_scale, const float* d_ctx, const float* d_act_a, const float* d_act_ib, int epilogue,
                  int act_in, int B, int Cin, int Cout, int T, int ksize, int dilation, void* stream);
int vb_codec_convtr(float* d_y, const float* d_x, const float* d_w_packed, const float* d_bias, const float* d_ctx,
                    const float* d_act_a, const float* d_act_ib, int act_in, int B, int Cin, int Cout, int T, int stride,
                    void* stream);
/* the same two calls on the tcgen05 tf32 hi/lo kernel of the SNAC stages (fp32-grade results): weights packed by
 * vb_snac_pack_tf32x3(phases = 1, M = Cout, K = ksize * Cin) from the TAP-MAJOR matrix [Cout][ksize][Cin] (conv) or
 * (phases = stride, M = Cout, K = 2 Cin) from the per-phase matrix above (transposed conv); Cin must be a multiple of 32 */
int vb_codec_conv_tc(float* d_y, const float* d_x, const void* d_w_tiles, const float* d_bias, const float* d_resid,
                     const float* d_scale, const float* d_ctx, const float* d_act_a, const float* d_act_ib, int epilogue,
                     int act_in, int B, int Cin, int Cout, int T, int ksize, int dilation, void* stream);
int vb_codec_convtr_tc(float* d_y, const float* d_x, const void* d_w_tiles, const float* d_bias, const float* d_ctx,
                       const float* d_act_a, const float* d_act_ib, int act_in, int B, int Cin, int Cout, int T, int stride,
                       void* stream);
/* y = act(x) over [B][C][T] (what the tensor-core calls take as input when the layer has an input activation) */
int vb_codec_activate(float* d_y, const float* d_x, const float* d_act_a, const float* d_act_ib, int act_in, int B, int C, int T,
                      void* stream);
int vb_codec_cache_update(float* d_cache, const float* d_x, const float* d_act_a, const float* d_act_ib, int act_in, int B,
                          int C, int L, int pad, void* stream);
int vb_codec_dwconv(float* d_y, const float* d_x, const float* d_w, const float* d_bias, const float* d_ctx, int B, int C,
                    int T, int ksize, void* stream);
int vb_codec_rmsnorm(float* d_y, const float* d_x, const float* d_w, int B, int C, int T, float eps, void* stream);
int vb_codec_attn_chunk(float* d_out, const float* d_qkv, float* d_cache, int64_t cache_batch_stride, const int64_t* d_pos0,
                        int B, int H, int Hkv, int D, int T, int W, float theta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VB_API_H_ */
