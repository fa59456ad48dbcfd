/*
 * tsim_b200 -- C ABI of the B200 sampling backend for tsim's compiled sampler.
 *
 * Every entry point replaces one piece of the reference's Python/JAX hot path
 * (paths relative to the tsim repository, v0.1.5):
 *
 *   tsb_program_create   <- CompiledProgram / CompiledComponent / CompiledScalarGraphs pytrees handed to
 *                           jitted code (src/tsim/core/types.py:55-107, src/tsim/compile/compile.py:21-37);
 *                           here: one flat bit-packed blob (tsim_b200/pack.py) uploaded to HBM once.
 *   tsb_sample_host      <- sample_program(program, f_params, key) (src/tsim/sampler.py:117-167) including
 *                           _sample_component (:28-81), evaluate (src/tsim/compile/evaluate.py:15-59), the
 *                           jnp.asarray H2D (:398) and copy_d2h (src/tsim/utils/cuda_helpers.py:105-141).
 *   tsb_sample_device    <- same, for callers that already hold device buffers (multi-GPU shards).
 *   tsb_evaluate_host    <- evaluate(circuit, param_vals) as used by CompiledStateProbs.probability_of
 *                           (src/tsim/sampler.py:906-953).
 *   tsb_host_alloc/free  <- alloc_pinned_numpy / _PinnedBuf (src/tsim/utils/cuda_helpers.py:48-102).
 *   tsb_split_key        <- jax.random.split(key) (call sites src/tsim/sampler.py:74,272,399,482).
 *
 * Conventions: plain pointers and sizes only; all functions return 0 on success and a negative
 * tsb_status otherwise (tsb_last_error() gives the message, thread-local).  A handle owns its device
 * buffers and streams; calls on one handle are not re-entrant.  There is no CPU fallback: without a
 * CUDA device every compute entry point fails with TSB_ERR_CUDA.
 */
#ifndef TSIM_B200_H
#define TSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tsb_program tsb_program;

typedef enum {
  TSB_OK = 0,
  TSB_ERR_INVALID = -1, /* bad argument / malformed blob */
  TSB_ERR_CUDA = -2,    /* CUDA runtime failure (message has the CUDA error string) */
  TSB_ERR_UNSUPPORTED = -3
} tsb_status;

/* f_format */
#define TSB_F_BYTES 0  /* uint8  [B, num_f] row-major, values 0/1 (the reference's f_params) */
#define TSB_F_PACKED 1 /* uint64 [B, ceil(num_f/64)], bit i of a row = f_i */
/* out_format */
#define TSB_OUT_BYTES 0  /* uint8 (numpy bool) [B, num_outputs] -- what sample_program returns */
#define TSB_OUT_PACKED 1 /* uint64 [B, ceil(num_outputs/64)], bit j = output column j; viewed as bytes this
                            is np.packbits(..., axis=1, bitorder="little") padded to 8 bytes per row */

typedef struct {
  int32_t mode;         /* 0 faithful order, 1 reordered-exact per row ("fast"), 2 reordered-exact bit-sliced ("sliced") */
  int32_t words;        /* 32-bit parameter words per shot */
  int32_t num_f;
  int32_t num_outputs;
  int32_t n_direct;
  int32_t n_components;
  int32_t n_draws;      /* compiled (non-direct) outputs = Bernoulli draws per shot */
  int32_t words_f64;    /* uint64 words per packed f row */
  int32_t words_out64;  /* uint64 words per packed output row */
  int32_t resident;     /* 1: all of g_{tki} stays in shared memory; 0: streamed chunk by chunk */
  int32_t n_chunks;
  int32_t smem_bytes;   /* dynamic shared memory per CTA */
  int32_t threads;      /* threads per CTA */
  int32_t grid;         /* CTAs per launch (persistent) */
  int64_t data_bytes;   /* packed g_{tki} in HBM */
} tsb_info;

const char* tsb_last_error(void);
int tsb_device_count(void);

int tsb_program_create(const uint32_t* blob, size_t n_words, int device, tsb_program** out);
int tsb_program_destroy(tsb_program* p);
int tsb_program_info(const tsb_program* p, tsb_info* info);

/* A bit-sliced (mode 2) program works on words of 32 shots and cannot evaluate single rows; the normalisation
 * check of shot 0 (sampler.py:66-72) and tsb_evaluate_host run on a companion per-row program of the same
 * CompiledProgram.  The companion is not owned: destroy it after the sliced program. */
int tsb_program_set_aux(tsb_program* p, tsb_program* aux);

/* Optional pattern cache (no reference counterpart): tabulate |E_k(pattern, prefix)| for every
 * selected-f pattern of weight <= max_weight (0..3; -1 switches the cache off) with the sampling kernel's own
 * evaluator, so that shots with such a pattern only walk the table.  Bits are identical with and without the cache.
 * A bit-sliced program tabulates with its companion's evaluator and runs K0t / K1s / K2a over the remaining rows.
 * max_entries <= 0: default budget (2^22 floats, L2-sized).  entries_out: table size actually built. */
int tsb_program_set_pattern_cache(tsb_program* p, int max_weight, int64_t max_entries, int64_t* entries_out);

/* (carry, sub) = jax.random.split(key): carry = out[0..1], sub = out[2..3] */
void tsb_split_key(uint32_t k0, uint32_t k1, uint32_t out[4]);

/* One batch (or a shard [shot_offset, shot_offset+B) of a batch) with device-resident buffers.
 * d_f: packed rows; d_out: packed rows; d_norm_dev: float[n_components] (written only by the shard that
 * holds shot 0 of the batch, i.e. shot_offset == 0; NULL = keep it inside the handle).
 * stream: cudaStream_t, NULL = the default stream.  Asynchronous: returns after enqueueing. */
int tsb_sample_device(tsb_program* p, const uint64_t* d_f, int64_t B, int64_t shot_offset, uint32_t k0,
                      uint32_t k1, uint64_t* d_out, float* d_norm_dev, void* stream);

/* Host buffers in, host buffers out; H2D, kernels and D2H are pipelined over slices of the batch.
 * norm_dev: float[n_components] or NULL. */
int tsb_sample_host(tsb_program* p, const void* f, int f_format, int64_t B, int64_t shot_offset, uint32_t k0,
                    uint32_t k1, void* out, int out_format, float* norm_dev);

/* amp[2*b], amp[2*b+1] = (re, im) of evaluate(component.compiled_scalar_graphs[level], params)[b];
 * params: uint8 [B, n_params(level)] 0/1. */
int tsb_evaluate_host(tsb_program* p, int component, int level, const uint8_t* params, int64_t B, float* amp);

/* helpers on device buffers (stream NULL = default stream) */
int tsb_pack_f_device(tsb_program* p, const uint8_t* d_bytes, int64_t B, uint64_t* d_packed, void* stream);
int tsb_unpack_out_device(tsb_program* p, const uint64_t* d_packed, int64_t B, uint8_t* d_bytes, void* stream);

/* device time (ms, CUDA events on the launch stream) of the sampling kernel launches of the last
 * tsb_sample_device / tsb_sample_host call, and how many launches that was.  After tsb_sample_device on a bit-sliced
 * program the interval is the sampling kernel alone (sample_sliced_kernel, without the transpose / assemble helpers). */
float tsb_last_kernel_ms(tsb_program* p, int* n_launches);

/* ---- K5: error-mechanism sampler on the device (statistical parity with ChannelSampler.sample,
 * src/tsim/noise/channels.py:624-658; tables as produced by _precompute_sparse, :578-622) ----
 * Channel c has n_outcomes[c] non-identity outcomes; thresholds (concatenated over channels) are the cumulative
 * outcome probabilities scaled to 2^64 (p_fire = last threshold of the channel / 2^64); patterns are the outcomes' packed
 * f rows, [sum n_outcomes][words_f64].  Like the reference (:638-656) the sampler walks from fire to fire with geometric
 * gaps -- per channel and block of 1024 in-batch shots, counter-based Philox4x32-10 -- so the work is O(fires). */
typedef struct tsb_noise tsb_noise;
int tsb_noise_create(int n_channels, const int32_t* n_outcomes, const uint64_t* thresholds, const uint64_t* patterns,
                     int words_f64, int device, tsb_noise** out);
int tsb_noise_destroy(tsb_noise* n);
/* f rows for in-batch shots [shot_offset, shot_offset + B): a pure function of (seed, call, shot index, channel),
 * so any partition of a batch over calls or GPUs gives the same rows.  skip_shot0: leave shot 0 noiseless. */
int tsb_noise_sample_device(tsb_noise* n, int64_t B, int64_t shot_offset, uint64_t seed, uint64_t call, int skip_shot0,
                            uint64_t* d_f, void* stream);
int tsb_noise_sample_host(tsb_noise* n, int64_t B, int64_t shot_offset, uint64_t seed, uint64_t call, int skip_shot0,
                          uint64_t* f_host);
/* noise -> sample -> D2H without the f rows ever leaving the GPU (f_out: optional copy of the packed rows). */
int tsb_sample_noisy_host(tsb_program* p, tsb_noise* n, int64_t B, int64_t shot_offset, uint32_t k0, uint32_t k1,
                          uint64_t noise_seed, uint64_t noise_call, int skip_shot0, void* out, int out_format,
                          float* norm_dev, uint64_t* f_out);

/* The same pipeline with the result in the caller's column layout -- the flag ladder of CompiledDetectorSampler.sample
 * (src/tsim/sampler.py:791-868: detector / observable split, prepend / append observables, reference-sample XOR) and
 * _maybe_bit_pack (:263-276) applied on the device, so that only the bytes the caller asked for cross PCIe.
 * layout: up to four column ranges [lo, lo + n) concatenated along the bit axis; bit_packed = 1 gives
 * np.packbits(..., bitorder="little") bytes per row, bit_packed = 0 one bool byte per column.  split > 0 sends the first
 * `split` ranges to `out` and the rest to a second array `out2` (separate_observables); row sizes: tsb_layout_row_bytes
 * (which = 0 / 1).  xor_row: uint64[words_out64] XORed into every row first (NULL = none).  ref_mask (NULL = none): shot 0
 * of the batch is the reference sample (:404-409): (row 0 & ref_mask) is XORed into every row as well, row 0 itself goes
 * to row0_out (uint64[words_out64]) and is left out of the result, which then has B - 1 rows. */
typedef struct {
  int32_t n_segments;
  int32_t lo[4];
  int32_t n[4];
  int32_t bit_packed;
  int32_t split;
} tsb_layout;
int64_t tsb_layout_row_bytes(const tsb_layout* layout, int which);
int tsb_sample_noisy_host_layout(tsb_program* p, tsb_noise* n, int64_t B, int64_t shot_offset, uint32_t k0, uint32_t k1,
                                 uint64_t noise_seed, uint64_t noise_call, int skip_shot0, const tsb_layout* layout,
                                 const uint64_t* xor_row, const uint64_t* ref_mask, uint8_t* out, uint8_t* out2, uint64_t* row0_out,
                                 float* norm_dev);

/* ---- post-selection session (reference src/tsim/sampler.py:422-545, _sample_batches_with_postselection) ----
 * The host keeps the reference's control flow and key schedule; the data path stays on the GPU.  Per chunk of at most
 * batch_size shots (push_host: packed f rows from the host; push_noise: rows generated by K5): the direct detector bits
 * are computed, shots with (direct ^ ref_row) & mask_row != 0 are discarded (sampler.py:517-520), surviving f rows and
 * their shot indices are appended IN SHOT ORDER to a pending buffer; *pending_out = survivors waiting.  dispatch samples
 * exactly batch_size pending rows under the given batch key (RNG counter = position in that batch, as in the
 * reference's _dispatch, :486-492) and scatters the rows to their shots; final_batch = 1 pads a partial batch with its
 * first row (:499-505).  finish XORs xor_kept / xor_discarded into kept / discarded rows (reference-sample handling,
 * :531-537; NULL = none) and copies out bool bytes or packed rows plus the discard flags.
 * mask_row / ref_row / xor rows: uint64[words_out64] over output columns; num_detectors = columns [0, nd). */
typedef struct tsb_postselect tsb_postselect;
int tsb_postselect_create(tsb_program* p, int64_t shots, int64_t batch_size, const uint64_t* mask_row, const uint64_t* ref_row,
                          int num_detectors, tsb_postselect** out);
int tsb_postselect_push_host(tsb_postselect* s, const uint64_t* f_packed, int64_t n, int64_t* pending_out);
int tsb_postselect_push_noise(tsb_postselect* s, tsb_noise* noise, int64_t n, uint64_t seed, uint64_t call, int64_t* pending_out);
int tsb_postselect_dispatch(tsb_postselect* s, uint32_t k0, uint32_t k1, int final_batch, float* norm_dev, int64_t* pending_out);
int tsb_postselect_finish(tsb_postselect* s, const uint64_t* xor_kept, const uint64_t* xor_discarded, void* out, int out_format,
                          uint8_t* discarded_out);
int tsb_postselect_destroy(tsb_postselect* s);

/* Multi-GPU gather of the packed output rows (the one collective of the path) over peer memory: a
 * rank pushes its rows into every peer's receive buffer with copy-engine transfers -- no SM, so the push overlaps the next
 * batch's sampling kernel.  dst is a peer-mapped (or local) device address, e.g. from CUDA IPC / symmetric memory. */
int tsb_memcpy_peer_async(void* dst, const void* src, size_t nbytes, void* stream);

/* free / total device memory in bytes: what the reference's automatic batch size is derived from
 * (src/tsim/sampler.py:294-320, _estimate_batch_size). */
int tsb_device_mem_info(int device, int64_t* free_bytes, int64_t* total_bytes);

void* tsb_host_alloc(size_t nbytes); /* page-locked host memory, NULL on failure */
void tsb_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* TSIM_B200_H */
