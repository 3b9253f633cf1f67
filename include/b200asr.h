/* b200asr -- C ABI of the B200-native Conformer-CTC encode + decode path.
 *
 * Drop-in boundary.  The reference (Z-yq/TensorflowASR) has no plugin/FFI layer of its own: its
 * deployment code drives three onnxruntime sessions.  Each entry point below stands in for one of those calls (or
 * for one of the reference's native decoders) and keeps its tensor conventions: row-major, contiguous, float32
 * activations, int32 token ids, blank = last class unless told otherwise.
 *
 *   b200asr_encode        <-> encoder session Run: "inputs" f32 [B,L,1] -> "Identity:0" f32 [B,T',D]
 *                             Inference/CppInference/onnx/src/core/asr_session.cpp:77-97 (EncoderInference),
 *                             Inference/PythonInference/asr/src/asr.py:34-39 (extract_feature), test_asr.py:191
 *   b200asr_ctc_logits    <-> ctc_model session Run: "inputs" f32 [B,T',D] -> "Identity:0" f32 [B,T',V]
 *                             asr_session.cpp:100-122 (CTCInference), asr.py:66-70, test_asr.py:193
 *   b200asr_ctc_greedy    <-> std::vector<int> ctc_greedy_decoder(probs, blank_id, vocab_size)
 *                             Inference/CppInference/onnx/src/core/ctc_greedy_decoder.h:4-43; asr.py:56-61;
 *                             externals/ctc_decoders/ctc_greedy_decoder.cpp:4-45; tf.keras.backend.ctc_decode at
 *                             test_asr.py:198 (pads with -1, honours input_length)
 *   b200asr_ctc_beam      <-> ctc_beam_search_decoder / _batch(probs_seq, vocabulary, beam_size, cutoff_prob,
 *                             cutoff_top_n, ext_scorer = nullptr)   externals/ctc_decoders/ctc_beam_search_decoder.h:26-60
 *   b200asr_recognize*    <-> the whole offline_stt chain test_asr.py:186-200 (encoder -> ctc_model -> greedy ids)
 *
 * Conventions: every function returns 0 on success, non-zero on error (then b200asr_last_error() describes it);
 * no exception crosses the ABI.  `*_dev` pointers are CUDA device pointers owned by the caller, `*_host` pointers
 * are host memory (pinned memory makes the copies asynchronous).  `stream` is a cudaStream_t passed as void*
 * (NULL = default stream); device-pointer entry points only enqueue work and never synchronise.  The library owns
 * the weight copy and a workspace that grows on demand.  Every entry point takes the handle's mutex, so a handle may be shared between host threads like an onnxruntime session
 * (calls serialise; use one handle per thread / GPU for concurrency).
 * There is no CPU fallback: creating a handle without a usable sm_100 GPU fails.
 */
#ifndef B200ASR_H_
#define B200ASR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ASR_ABI_VERSION 1

#if defined(__GNUC__)
#define B200ASR_API __attribute__((visibility("default")))
#else
#define B200ASR_API
#endif

typedef struct b200asr_engine* b200asr_handle;

/* arithmetic used by the GEMM-shaped ops (everything else is always fp32) */
enum { B200ASR_PRECISION_TF32 = 0, /* tcgen05 tensor cores, tf32 inputs, fp32 accumulate */
       B200ASR_PRECISION_FP32 = 1  /* CUDA-core fp32 (exact mode, used as on-device checker) */ };

/* Geometry = the reference's model_config / speech_config keys (asr/configs/conformerS.yml, am_data.yml). */
typedef struct {
  int32_t abi_version;        /* B200ASR_ABI_VERSION */
  /* encoder: model_config.{dmodel,num_blocks,num_heads,head_size,kernel_size}, ff = 4*dmodel */
  int32_t dmodel, num_blocks, num_heads, head_size, kernel_size, ff_dim;
  /* CTC decoder: model_config.ctcdecoder_{num_blocks,kernel_size}; vocab = classes incl. blank */
  int32_t ctc_blocks, ctc_kernel_size, vocab;
  /* frontend: speech_config.{num_feature_bins, sample_rate*stride_ms/1000}; n_dft fixed to 1024 by the reference */
  int32_t n_mels, n_dft, hop;
  float ln_eps;               /* Keras LayerNormalization default 1e-3 */
  /* block streaming (StreamingConformerEncoder, conformer_blocks.py:567-594): >0 = encode independent chunks of this
     many samples (speech_config.streaming_bucket * sample_rate); 0 = offline */
  int32_t chunk_samples;
  int32_t precision;          /* B200ASR_PRECISION_* */
  int32_t use_cuda_graph;     /* 1 = capture each (shape, pointer set) once and replay */
  /* translator (pinyin ids + encoder states -> characters, conformer_blocks.py:504-552); tr_blocks = 0: no translator in the blob */
  int32_t tr_blocks, tr_kernel_size, tr_inp_classes, tr_vocab;
  int32_t reserved[4];
} b200asr_config;

/* Weight blob: produced by tensorflowasr_b200.weights.pack_for_device() (see weights.py for the tensor list).
 *   char magic[8] = "B2ASRW01"; uint32 n_entries; uint32 pad;
 *   n_entries x { char name[48]; uint64 byte_offset; uint64 numel; }   (float32 tensors, offsets 128-byte aligned)
 */
B200ASR_API int b200asr_create(const void* weight_blob, size_t blob_bytes, const b200asr_config* cfg, int device,
                   b200asr_handle* out);
B200ASR_API int b200asr_destroy(b200asr_handle h);
B200ASR_API const char* b200asr_last_error(b200asr_handle h /* may be NULL: error of the failed create on this thread */);
B200ASR_API int b200asr_abi_version(void);

/* T' (encoder frames) produced for L samples: two 'same' stride-2 convs over ceil(L/hop) mel frames. */
B200ASR_API int b200asr_out_frames(b200asr_handle h, int num_samples);
/* Pre-size the workspace for batches up to (B, L) so later calls never allocate. */
B200ASR_API int b200asr_reserve(b200asr_handle h, int B, int L);

B200ASR_API int b200asr_encode(b200asr_handle h, const float* wav_dev /*[B,L]*/, int B, int L, float* enc_dev /*[B,T',D]*/,
                   void* stream);
/* mel features only ([B, ceil(L/hop), n_mels]) -- the Melspectrogram layer output (time_frequency.py:173-189) */
B200ASR_API int b200asr_mel(b200asr_handle h, const float* wav_dev, int B, int L, float* mel_dev, void* stream);
B200ASR_API int b200asr_ctc_logits(b200asr_handle h, const float* enc_dev /*[B,T',D]*/, int B, int Tp, float* logits_dev /*[B,T',V]*/,
                       void* stream);
/* lengths_dev may be NULL (= all Tp).  ids_dev [B,Tp] is padded with -1, out_len_dev [B]. */
B200ASR_API int b200asr_ctc_greedy(b200asr_handle h, const float* logits_dev, const int32_t* lengths_dev, int B, int Tp, int V,
                       int blank, int32_t* ids_dev, int32_t* out_len_dev, void* stream);
/* Prefix beam search without external scorer.  ids_dev [B,beam,Tp] (-1 padded), out_len_dev [B,beam],
 * scores_dev [B,beam] (log prob, descending).  logits are softmax-normalised inside. */
B200ASR_API int b200asr_ctc_beam(b200asr_handle h, const float* logits_dev, const int32_t* lengths_dev, int B, int Tp, int V,
                     int blank, int beam, int cutoff_top_n, float cutoff_prob, int32_t* ids_dev, int32_t* out_len_dev,
                     float* scores_dev, void* stream);

/* Translator session Run: "inputs" int32 [B,U] (greedy phone ids, zero padded: the deployment appends ten zeros), "enc" f32 [B,T',D]
 * -> "Identity:0" f32 [B,U,tr_vocab]   (Inference/CppInference/onnx/src/core/asr_session.cpp:125-150, PythonInference asr.py:77-83).
 * Embedding -> RBlocks (FFModule, cross attention of LN(x + sinusoidal positions) over the encoder states, ConvModule, FFModule, LN)
 * -> Dense. */
B200ASR_API int b200asr_translate(b200asr_handle h, const int32_t* ids_dev /*[B,U]*/, const float* enc_dev /*[B,T',D]*/, int B, int U, int Tp,
                                  float* logits_dev /*[B,U,tr_vocab]*/, void* stream);

/* The same decoder on PROBABILITIES (float32 [B,Tp,V], rows summing to 1) -- the reference decoder's own input convention
 * (probs_seq, ctc_beam_search_decoder.h:26-45).  With identical probabilities the scores equal the reference C++ decoder's bit for
 * bit (float log_sum_exp evaluated with correctly rounded exp / log) and so does the hypothesis order. */
B200ASR_API int b200asr_ctc_beam_probs(b200asr_handle h, const float* probs_dev, const int32_t* lengths_dev, int B, int Tp, int V,
                           int blank, int beam, int cutoff_top_n, float cutoff_prob, int32_t* ids_dev, int32_t* out_len_dev,
                           float* scores_dev, void* stream);

/* wav -> greedy token ids in one call, device buffers. */
B200ASR_API int b200asr_recognize(b200asr_handle h, const float* wav_dev, int B, int L, int32_t* ids_dev /*[B,T']*/,
                      int32_t* out_len_dev /*[B]*/, void* stream);
/* Same, honouring per-utterance lengths like tf.keras.backend.ctc_decode(ctc_output, input_length) in the reference's batched
 * evaluation (asr/tester/am_tester.py:34-40): frame_lengths_dev [B] = number of ENCODER frames of each zero-padded utterance that
 * are decoded (NULL = all T').  The encoder itself has no padding mask, exactly like the reference's. */
B200ASR_API int b200asr_recognize_lengths(b200asr_handle h, const float* wav_dev, const int32_t* frame_lengths_dev, int B, int L,
                                          int32_t* ids_dev /*[B,T']*/, int32_t* out_len_dev /*[B]*/, void* stream);
/* Same with HOST buffers: H2D of the waveform, compute, D2H of ids + lengths, stream-synchronised on return. */
B200ASR_API int b200asr_recognize_host(b200asr_handle h, const float* wav_host, int B, int L, int32_t* ids_host,
                           int32_t* out_len_host, void* stream);

/* Two-deep pipeline over the host-buffer call, for serving loops that keep the GPU busy (the reference's batch decoders use a
 * thread pool for the same purpose, ctc_beam_search_decoder.cpp:426-459): submit() enqueues H2D of the waveform (pinned host
 * memory) on a copy stream, the recognise graph and the D2H of ids + lengths on the library's own compute stream and returns at
 * once; collect() blocks until that slot's results are in the host buffers.  slot is 0 or 1; a slot must be collected before
 * it is submitted again; the H2D of one slot overlaps the compute of the other.  Host buffers must stay valid until collect(). */
B200ASR_API int b200asr_recognize_host_submit(b200asr_handle h, int slot, const float* wav_host, int B, int L, int32_t* ids_host,
                                              int32_t* out_len_host);
B200ASR_API int b200asr_recognize_host_collect(b200asr_handle h, int slot);

/* ------------------------------------------------------------------------------------------------------------------------------
 * ChunkConformer: causal chunk streaming with state caches (asr/models/chunk_conformer_blocks.py, driven by test_chunk_asr.py:47-139).
 * A chunk engine is created from its own weight blob (tensorflowasr_b200.chunk_model.pack_chunk_blob) and uses the tcgen05 path
 * only.  A stream state holds the caches of B streams that advance in lockstep (one state per group of concurrent streams; a
 * state belongs to one handle, i.e. one GPU, for its lifetime -- SURVEY 8e).
 *
 *   b200asr_stream_step           <-> ChunkConformer.picker_stream_predict (:807-824): front end (wav cache 2560 samples, mel cache
 *                                     4 frames) -> 15 causal encoder blocks with K|V / conv caches -> picker block -> phone logits
 *   b200asr_stream_feature_pick   <-> ChunkConformer.feature_pick (:913-999): keep the frames whose phone argmax is not blank
 *   b200asr_stream_decoder_step   <-> ChunkConformer.decoder_stream_predict (:826-837): helper blocks -> decoder block with
 *                                     `dec_back` frames of look-ahead (the frames still inside it are carried to the next call)
 */
typedef struct b200asr_stream_state* b200asr_stream;

typedef struct {
  int32_t abi_version;                                   /* B200ASR_ABI_VERSION */
  int32_t dmodel, num_heads, head_size, kernel_size, ff_dim;   /* chunk_conformerS.yml model_config */
  int32_t enc_blocks, picker_blocks, helper_blocks, dec_blocks;
  int32_t win_front, picker_back, dec_back;              /* attention band: frames back / look-ahead of picker and decoder */
  int32_t phone_classes, txt_classes;                    /* real class counts incl. blank (= last); the blob pads both heads to a multiple of 4 */
  int32_t n_mels, n_dft, hop, chunk_num, reduction;      /* chunk_num mel frames per step (16 = 160 ms, 32 = 320 ms), reduction 4 */
  float ln_eps;
  int32_t use_cuda_graph;
  int32_t reserved[8];
} b200asr_chunk_config;

B200ASR_API int b200asr_chunk_create(const void* weight_blob, size_t blob_bytes, const b200asr_chunk_config* cfg, int device, b200asr_handle* out);
B200ASR_API int b200asr_stream_state_create(b200asr_handle h, int B, b200asr_stream* out);
B200ASR_API int b200asr_stream_state_reset(b200asr_handle h, b200asr_stream st);     /* init_picker_caches / init_decoder_caches (:777-792) */
B200ASR_API int b200asr_stream_state_destroy(b200asr_handle h, b200asr_stream st);
/* One picker step: wav_chunk_dev [B, chunk_num*hop] -> phone_logits_dev [B, T, phone_classes], hidden_dev [B, T, dmodel], T = chunk_num / reduction. */
B200ASR_API int b200asr_stream_step(b200asr_handle h, b200asr_stream st, const float* wav_chunk_dev, float* phone_logits_dev, float* hidden_dev,
                                    void* stream);
/* feats_dev [B, T, dmodel] (kept rows first, zero rows behind), picked_logits_dev [B, T, V] (nullable), counts_dev [B]; when
 * n_max_host != NULL the call synchronises `stream` and stores the largest count there (the caller needs it to shape the decoder input). */
B200ASR_API int b200asr_stream_feature_pick(b200asr_handle h, const float* hidden_dev, const float* phone_logits_dev, int B, int T, int V, int blank,
                                            float* feats_dev, float* picked_logits_dev, int32_t* counts_dev, int32_t* n_max_host, void* stream);
/* One decoder step on n picked frames per stream (feats_dev [B, n, dmodel] contiguous, n >= 1).  Writes the text logits of the carried +
 * new frames, txt_logits_dev [B, n_rows, txt_classes] with n_rows = carried + n <= dec_back + n; the first *n_valid_host rows of each
 * stream are final ("valid"), the remaining n_rows - n_valid are still inside the look-ahead ("unvalid") and are re-fed by the next call. */
B200ASR_API int b200asr_stream_decoder_step(b200asr_handle h, b200asr_stream st, const float* feats_dev, int n, float* txt_logits_dev,
                                            int32_t* n_rows_host, int32_t* n_valid_host, void* stream);
/* rows the next decoder step will produce for n new frames (= carried frames + n) */
B200ASR_API int b200asr_stream_decoder_rows(b200asr_handle h, b200asr_stream st, int n);

/* ---------------------------------------------------------------------------------------------------------------------------
 * Session layer: the voice-activity model.  Replaces the onnxruntime session of
 * Inference/PythonInference/vad/src/vad.py:22-28 (`VAD.compile` / `VAD.inference`: vad/models/vad.onnx, input "inputs" f32 [1, N, 80]
 * = N frames of 80 samples of the 8 kHz signal, output f32 [1, N, 1] logits that the callers threshold at 0:
 * offline_asr_session.py:79-90, stream_asr_session.py:333-341).
 *   weight blob  same container as b200asr_create; tensors d0..d3 .w [80, 80] / .b [80] (Dense, [out, in]), c0, c1 .w [80, 5 * 80]
 *                (causal Conv1D, k = tap * 80 + in) / .b [80], ln.g / ln.b [80], d4.w [4, 80] / d4.b [4] (the single output unit in row 0)
 *   wav_dev      [B, N * 80 * stride] f32; stride 2 reads every second sample (16 kHz audio in, the callers' `wav[::2]`), stride 1 = 8 kHz
 *   logits_dev   [B, N] f32.  Exact fp32 (CUDA cores).  The handle is released with b200asr_destroy.
 */
B200ASR_API int b200asr_vad_create(const void* weight_blob, size_t blob_bytes, float ln_eps, int device, b200asr_handle* out);
B200ASR_API int b200asr_vad_infer(b200asr_handle h, const float* wav_dev, int B, int N, int stride, float* logits_dev, void* stream);

/* Session layer: the punctuation model.  Replaces the onnxruntime session of
 * Inference/PythonInference/punc_recover/src/punc_recover.py:40-62 (`Punc.compile` / `Punc.punc_recover`: punc_recover/models/punc.onnx,
 * inputs token ids i32 [1, U] (<S> characters </S>), padding mask, sinusoidal table; output f32 [1, U, 32] class probabilities).
 *   weight blob  tensors emb [V, 64] (pre-scaled by sqrt(64)), pe [rows, 64], in / up / down / out .w [N, K] .b, l0..l4 .qkv.w [192, 64]
 *                (1/sqrt(8) folded into the q rows) .qkv.b .o .f1 .f2 .ln1 .ln2, c0..c2 .w [64, 3 * 64] .b  (punc_model.punc_device_tensors)
 *   ids_dev      [U] i32, one sentence without padding (the mask input of the graph is empty for a single sentence), U <= rows of pe
 *   probs_dev    [U, 32] f32.  Exact fp32.  Synchronises `stream` (an id outside the table is an error).  Released with b200asr_destroy.
 */
B200ASR_API int b200asr_punc_create(const void* weight_blob, size_t blob_bytes, float ln_eps, int device, b200asr_handle* out);
B200ASR_API int b200asr_punc_infer(b200asr_handle h, const int32_t* ids_dev, int U, float* probs_dev, void* stream);

/* Roofline instrumentation (bench.py): time one stage of the schedule alone, `iters` launches bracketed by CUDA events on
 * `stream`; also returns that launch's algorithmic FLOPs and HBM bytes.  Run b200asr_recognize with the same (B, L) first. */
enum { B200ASR_STAGE_CONV2 = 0, B200ASR_STAGE_FFN_W1 = 1, B200ASR_STAGE_FFN_W2 = 2, B200ASR_STAGE_STFT = 3,
       B200ASR_STAGE_SUBLIN = 4, B200ASR_STAGE_ATTENTION = 5, B200ASR_STAGE_CTC_FC = 6,
       B200ASR_STAGE_FFN_CHAIN = 7 /* whole FFModule as one chained kernel: both GEMMs + swish + residual + LayerNorm */,
       B200ASR_STAGE_CONV1 = 8, B200ASR_STAGE_DWCONV = 9, B200ASR_STAGE_QKV = 10 };
B200ASR_API int b200asr_time_stage(b200asr_handle h, int stage, int B, int L, int iters, void* stream, float* ms_per_launch,
                                   double* flops, double* bytes);

/* Test hook: C[M,ldc] = epilogue(A[M,lda(K)] . W[N,K]^T) through the tcgen05 tf32 kernel (use_tensor_cores = 1) or the
 * fp32 CUDA-core kernel (0).  epilogue: 0 bias, 1 bias+relu, 2 bias+swish, 3 GLU (pairs), 4 resid + alpha*(acc+bias), 5 none. */
B200ASR_API int b200asr_debug_gemm(b200asr_handle h, const float* A, const float* W, const float* bias, const float* resid,
                                   float* C, int M, int N, int K, int lda, int ldc, float alpha, int epilogue,
                                   int use_tensor_cores, void* stream);

/* Test hook for the fused LayerNorm epilogues (6: C = resid+alpha*(acc+bias), C2 = LN1(C); 7: C = LN1(resid+alpha*(acc+bias)),
 * C2 = LN2(C) unless ln2_g is NULL; 8: C = acc+bias, C2 = LN1(C)).  tcgen05 path only; N in {64,128,144,192,256}. */
B200ASR_API int b200asr_debug_gemm_ln(b200asr_handle h, const float* A, const float* W, const float* bias, const float* resid,
                                      float* C, float* C2, int M, int N, int K, float alpha, int epilogue, const float* ln1_g,
                                      const float* ln1_b, const float* ln2_g, const float* ln2_b, float eps, void* stream);

/* Test hook: softmax(Q K^T) V per head on qkv [B*T, 3*H*dh] (q | k | v, head-major, q pre-scaled) -> out [B*T, H*dh];
 * win_front < 0 = full attention, else the ChunkConformer band.  tcgen05 kernel (1) or fp32 CUDA-core kernel (0). */
B200ASR_API int b200asr_debug_attention(b200asr_handle h, const float* qkv, float* out, int B, int T, int H, int dh,
                                        int win_front, int win_back, int use_tensor_cores, void* stream);

/* Test hook: C/C2 = LN-epilogue(resid + alpha * (swish(X.W1^T + b1).W2^T + b2)) through the chained tcgen05 kernel
 * (hidden activations stay in TMEM).  X [M,K1], W1 [N1,K1], W2 [N2,N1]; epilogue 6 or 7 (see b200asr_debug_gemm_ln). */
B200ASR_API int b200asr_debug_chain(b200asr_handle h, const float* X, const float* W1, const float* b1, const float* W2,
                                    const float* b2, const float* resid, float* C, float* C2, int M, int K1, int N1, int N2,
                                    float alpha, int epilogue, const float* ln1_g, const float* ln1_b, const float* ln2_g,
                                    const float* ln2_b, float eps, void* stream);

/* Same contract as b200asr_debug_chain through the cluster-pair variant (hidden dimension split across two CTAs, partial
 * accumulators exchanged through distributed shared memory); N2 = 144, N1 = 288 or 576 only.  N1 = 0 (W1 = b1 = NULL, K1 = 144)
 * selects the DIRECT mode: C/C2 = LN-epilogue(resid + alpha * (X . W2^T + b2)) with K split across the pair. */
B200ASR_API int b200asr_debug_chain_pair(b200asr_handle h, const float* X, const float* W1, const float* b1, const float* W2,
                                         const float* b2, const float* resid, float* C, float* C2, int M, int K1, int N1, int N2,
                                         float alpha, int epilogue, const float* ln1_g, const float* ln1_b, const float* ln2_g,
                                         const float* ln2_b, float eps, void* stream);

/* Test hooks for the kernels that only the end-to-end path exercised: the conv module's depthwise convolution, and the two
 * subsampling convolutions (mel [B,T,n_mels] -> [B,T2,F2,D] NHWC) through the engine's precision path. */
B200ASR_API int b200asr_debug_dwconv(b200asr_handle h, const float* x, const float* w, float* y, int B, int T, int D, int K, int pad_left,
                                     int round_tf32, void* stream);
B200ASR_API int b200asr_debug_subsample_convs(b200asr_handle h, const float* mel_dev, int B, int T, float* out_dev, void* stream);

/* Test hook: b200asr_encode without CUDA graph, keeping the residual stream after the subsampler (tap 0) and after every encoder
 * block (taps 1..num_blocks) in taps_dev [n_taps][B*T', D] -- the tensors the reference's ONNX graph exposes as
 * conv_subsampling/dense/BiasAdd:0 and conformer_block_<i>/layer_normalization_<5i+4>/add:0 (profiles/r02_stage_errors.md). */
B200ASR_API int b200asr_debug_encode_taps(b200asr_handle h, const float* wav_dev, int B, int L, float* taps_dev, int n_taps, void* stream);

/* number of kernel launches the library has issued on this handle (bench.py "gpu_launches") */
B200ASR_API int64_t b200asr_launch_count(b200asr_handle h);

#ifdef __cplusplus
}
#endif
#endif /* B200ASR_H_ */
