/* libqkdldpc_cuda -- C ABI of the B200 batched LDPC syndrome decoder for QKD information reconciliation.
 *
 * This is the drop-in boundary for the hot path of ColdCloudd/QKD_LDPC_V. The reference has no FFI of its own
 * (its decoders are plain C++ functions, qkd_ldpc_algorithm.hpp:28-109); each entry point below names the reference
 * call site(s) it replaces. Conventions:
 *   - plain C, opaque handles, status-code returns (0 = ok, negative = error) + qkdldpc_last_error();
 *     no exception crosses the boundary; nothing is read from globals (the reference's `CFG` fields the decoders
 *     read are explicit fields of qkdldpc_params);
 *   - the caller owns every host buffer; the library owns device memory and its stream inside the handle;
 *   - calls on one handle are serialised by the caller; different handles may be used from different threads;
 *   - there is NO CPU implementation behind this ABI: without a CUDA device every compute entry point fails
 *     with QKDLDPC_ERR_CUDA.
 *
 * Bit packing ("packed frames"): frame f, bit i lives in word  bits[f * words_per_frame + (i >> 5)]  at bit
 * position (i & 31); words_per_frame = (n + 31) / 32; padding bits are zero. (numpy: packbits(bitorder="little").)
 */
#ifndef QKDLDPC_H_
#define QKDLDPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QKDLDPC_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define QKDLDPC_API __attribute__((visibility("default")))
#else
#define QKDLDPC_API
#endif

/* Status codes */
#define QKDLDPC_OK 0
#define QKDLDPC_ERR_INVALID (-1) /* bad argument / malformed graph */
#define QKDLDPC_ERR_CUDA (-2)    /* CUDA runtime error or no device */
#define QKDLDPC_ERR_NOMEM (-3)
#define QKDLDPC_ERR_STATE (-4)

/* Decoding algorithm selector: same numbering as DEC_* in the reference (config.hpp:201, config.hpp:126-133). */
#define QKDLDPC_ALG_SPA 0        /* sum_product_decoding,                  qkd_ldpc_algorithm.cpp:3-144    */
#define QKDLDPC_ALG_SPA_APPROX 1 /* sum_product_linear_approx_decoding,    qkd_ldpc_algorithm.cpp:174-315  */
#define QKDLDPC_ALG_NMSA 2       /* min_sum_normalized_decoding,           qkd_ldpc_algorithm.cpp:317-482  */
#define QKDLDPC_ALG_OMSA 3       /* min_sum_offset_decoding,               qkd_ldpc_algorithm.cpp:484-650  */
#define QKDLDPC_ALG_ANMSA 4      /* adaptive_min_sum_normalized_decoding,  qkd_ldpc_algorithm.cpp:652-839  */
#define QKDLDPC_ALG_AOMSA 5      /* adaptive_min_sum_offset_decoding,      qkd_ldpc_algorithm.cpp:841-1029 */

/* Output flag bits per frame (decoding_result / LDPC_result, qkd_ldpc_algorithm.hpp:16-26). */
#define QKDLDPC_FLAG_SYNDROMES_MATCH 1u
#define QKDLDPC_FLAG_KEYS_MATCH 2u

/* Tally vector layout (uint64 each; what process_trials_results consumes, simulation.cpp:580-624,683-689):
 *   [0] frames decoded            [1] frames with syndromes_match
 *   [2] frames with syndromes_match && keys_match        [3] sum of executed decoder iterations over all frames
 *   [4 + k], k = 0..max_iterations : number of syndrome-matched frames with iterations_num == k
 * Length = qkdldpc_tally_len(max_iterations) = max_iterations + 5. Integer sums => exact under any sharding
 * of frames over GPUs; a multi-GPU job all-reduces (sum) this vector once per batch. */
#define QKDLDPC_TALLY_FRAMES 0
#define QKDLDPC_TALLY_SYNDROMES_MATCH 1
#define QKDLDPC_TALLY_KEYS_MATCH 2
#define QKDLDPC_TALLY_ITERATIONS 3
#define QKDLDPC_TALLY_HIST 4

typedef struct qkdldpc_code qkdldpc_code; /* one parity-check matrix resident on one GPU + its slot pool */

/* Everything the reference decoders read from the global CFG (qkd_ldpc_algorithm.cpp:73,1056-1085) and from
 * decoding_scaling_factors (config.hpp:50-54). */
typedef struct qkdldpc_params {
    int32_t algorithm;         /* QKDLDPC_ALG_*            = CFG.DECODING_ALGORITHM                         */
    int32_t max_iterations;    /* >= 1                     = CFG.DECODING_ALG_MAX_ITERATIONS                */
    double primary;            /* alpha (NMSA/ANMSA) or beta (OMSA/AOMSA)   = scaling_factors.primary        */
    double secondary;          /* nu (ANMSA) or sigma (AOMSA)               = scaling_factors.secondary      */
    int32_t enable_threshold;  /* = CFG.ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD                               */
    double threshold;          /* = CFG.DECODING_ALG_MSG_LLR_THRESHOLD (> 0)                                */
    int32_t message_precision; /* 0: automatic (the precision policy below); 32: float32 messages; 64: float64
                                  messages (bit-identical to the reference for the min-sum family and SPA-lin)  */
} qkdldpc_params;

/* Precision policy of message_precision == 0 (qkdldpc_effective_precision): the narrowest state that meets the parity
 * bar against the reference's double arithmetic (>= 99 % equal iteration counts, FER inside its 95 % interval):
 *   OMSA, ANMSA, AOMSA            -> 64  (offset / adaptive min-sum is chaotic near threshold: float32 agrees on 95.7 ..
 *                                         99.2 % of the iteration counts only; the float64 on-chip kernel is exact)
 *   SPA, SPA-lin-approx, n > 65536 -> 64  (float32 has an error floor on the n = 102400 codes near capacity)
 *   everything else               -> 32  (NMSA: 99.9 - 100 %; SPA / SPA-lin on n <= 65536: 99.9 %)                   */
QKDLDPC_API int32_t qkdldpc_effective_precision(int32_t algorithm, int32_t n, int32_t message_precision);

/* Tuning knobs of a handle (all optional; 0 = library default). */
typedef struct qkdldpc_options {
    int64_t pool_bytes;     /* upper bound for the message pool in HBM (default 2 GiB)                      */
    int32_t pool_slots;     /* resident frame slots (rounded to whole tiles); 0 = derived from pool_bytes    */
    int32_t steps_per_poll; /* decoder iterations launched between two host checks of the done counter       */
    int32_t frames_per_lane_f32; /* 1, 2 or 4 (tile = 32 x this many frames); 0 = auto: 4, SPA variants 2   */
    int32_t use_graph;      /* 1 (default): replay one captured CUDA graph per poll interval; -1: plain launches */
    int32_t decoder_path;   /* 0 auto; 1 streaming kernels (messages in HBM); 2 on-chip kernels (frame state in shared
                               memory; float32 messages, check degrees <= 64, n < 65535; SPA variants: 4 E + 4 n bytes
                               must fit 227 KB) or QKDLDPC_ERR_INVALID when the code / parameters are not eligible */
    int32_t onchip_threads; /* CTA size of the on-chip kernels (multiple of 32; min-sum <= 768, sum-product <= 1024); 0 = auto */
    int32_t tail_compaction; /* streaming path: 0 (default) move the stragglers of a draining batch into few tiles; -1 never */
    int32_t compaction_max_ctas; /* tail compaction: cap of the copy kernels' grid (0 = 65535); they stride over the moves */
    int32_t copy_chunks;    /* qkdldpc_decode_batch (host buffers): number of chunks the batch is cut into so that the
                               host-to-device copy of chunk k+1 and the device-to-host copy of chunk k-1 overlap the
                               decoding of chunk k; 0 = auto, 1 = no overlap (copy, decode, copy back)                  */
    int32_t onchip_record_bytes; /* float32 on-chip min-sum kernel: 16 = one 16-byte record per row {c1, c2, signs, argmin}
                               (rows of 33..64 edges: two), 8 = 8-byte records {c1, signs | argmin} + a c2 array (rows of
                               28..51 edges: two; wider rows: the 16-byte format), 0 = auto: 8 when every row has at most
                               27 edges; results are identical                                                           */
    int32_t vn_items_per_warp; /* streaming path, narrow variable-node buckets (dv <= 8): items one warp walks per launch,
                               the next item's index records prefetched into L1 behind the current item's messages;
                               1 = one item per warp (vn_kernel_ell), 0 = auto: a walk sized to the grid (float32 with 2
                               or 4 frames per lane, float64), 1 for float32 with 1 frame per lane; results are identical */
    int32_t vn_ctas_per_sm; /* resident CTAs per SM the dv <= 4 walking kernel is compiled for; float32: 3..6 (72 / 64 / 48 /
                               40 registers; the dv <= 8 kernel: 2 for 3, else 3), float64: 3 (77 registers) or 1 (88);
                               0 = auto (float32: 4, with 2 frames per lane 5; float64: 3)                               */
    int32_t compaction_fill_pct; /* tail compaction starts once the queue is empty and at most this percentage of the resident
                               slots is still occupied (1..99); 0 = auto (75)                                            */
} qkdldpc_options;

QKDLDPC_API int qkdldpc_version(void);
/* Message of the last failing call made by the calling thread (thread-local, never NULL). */
QKDLDPC_API const char *qkdldpc_last_error(void);
QKDLDPC_API int64_t qkdldpc_tally_len(int32_t max_iterations);
/* Number of CUDA devices visible (0 if none / no driver). */
QKDLDPC_API int qkdldpc_device_count(void);

/* Builds the device-resident Tanner graph of one parity-check matrix. Replaces the per-call use of
 * `H_matrix` (array_and_matrix_operations.hpp:60-77) by the decoders: the host converts
 * H_matrix.check_nodes to CSR (row_ptr[m+1], col_idx[nnz]); the library derives the column view
 * (H_matrix.bit_nodes) itself, re-lays both out as edge-indexed CSR/CSC with degree-bucketed rows and
 * columns, and keeps them on `device`. col_idx must be ascending and duplicate-free inside every row
 * (true for every matrix the reference ships, quirk Q1); otherwise QKDLDPC_ERR_INVALID.
 * One handle per matrix, reused for every (QBER, scaling factor, rate adaptation) combination, like
 * sim_input.matrix (simulation.hpp:29-34). */
QKDLDPC_API int qkdldpc_code_create(qkdldpc_code **out, int32_t n, int32_t m, int64_t nnz, const int32_t *row_ptr,
                        const int32_t *col_idx, int32_t device, const qkdldpc_options *options /* may be NULL */);
QKDLDPC_API void qkdldpc_code_destroy(qkdldpc_code *code);

/* Optional: run all work of this handle on a caller-owned CUDA stream (a cudaStream_t passed as void*),
 * e.g. torch's current stream, so that the caller's CUDA events bracket the decoder's kernels. */
QKDLDPC_API int qkdldpc_code_set_stream(qkdldpc_code *code, void *cuda_stream);

/* Decodes a batch of independent frames: the drop-in for the trial loop
 *     pool.detach_loop(0, TRIALS, n -> run_trial(...))              simulation.cpp:740-746
 * and, per frame, for QKD_LDPC (qkd_ldpc_algorithm.cpp:1031-1119) or QKD_LDPC_RATE_ADAPT (:1121-1258):
 * a-priori LLRs from Bob's bits and the QBER, Alice's syndrome, the selected decoder, keys compare.
 *
 * Inputs (HOST pointers, packed frames, see top of file):
 *   alice_bits, bob_bits   n_frames x words_per_frame. With rate adaptation these are the EXTENDED frames
 *                          (alice_bit_array_extended / bob_bit_array_extended, :1136-1174): punctured positions
 *                          carry each party's random bit, shortened positions carry 0.
 *   qber                   per-frame QBER used for the LLR magnitude log((1-q)/q) (the "accurate QBER",
 *                          simulation.cpp:555,565); if qber_is_scalar != 0 only qber[0] is read.
 *   punct_pos / short_pos  ascending bit positions shared by the whole batch (H_matrix_params.punctured_bits /
 *                          shortened_bits, array_and_matrix_operations.hpp:44-48); may be NULL with count 0.
 *                          Punctured LLR = 1e-4 (ALMOST_ZERO, qkd_ldpc_algorithm.hpp:13), shortened LLR = the
 *                          largest finite value of the message type (DBL_MAX in the reference, :1164).
 * Outputs (HOST pointers, each may be NULL):
 *   out_bits   n_frames x words_per_frame, bob_solution = last hard decision (packed)
 *   out_iters  decoding_result.iterations_num
 *   out_flags  QKDLDPC_FLAG_* bits
 *   tally      qkdldpc_tally_len(max_iterations) uint64, overwritten with this batch's tallies
 */
QKDLDPC_API int qkdldpc_decode_batch(qkdldpc_code *code, const qkdldpc_params *params, int64_t n_frames,
                         const uint32_t *alice_bits, const uint32_t *bob_bits, const double *qber,
                         int32_t qber_is_scalar, const int32_t *punct_pos, int32_t n_punct, const int32_t *short_pos,
                         int32_t n_short, uint32_t *out_bits, int32_t *out_iters, uint8_t *out_flags,
                         uint64_t *tally);

/* Same call with every frame buffer already resident in device memory (DEVICE pointers for alice_bits, bob_bits,
 * qber, out_bits, out_iters, out_flags, tally; punct_pos / short_pos stay HOST pointers: they are per-batch
 * metadata). Work is enqueued on the handle's stream and the call returns after the batch completed. */
QKDLDPC_API int qkdldpc_decode_batch_device(qkdldpc_code *code, const qkdldpc_params *params, int64_t n_frames,
                                const uint32_t *d_alice_bits, const uint32_t *d_bob_bits, const double *d_qber,
                                int32_t qber_is_scalar, const int32_t *punct_pos, int32_t n_punct,
                                const int32_t *short_pos, int32_t n_short, uint32_t *d_out_bits, int32_t *d_out_iters,
                                uint8_t *d_out_flags, uint64_t *d_tally);

/* Synthetic keys generated on the device with the distribution of run_trial (simulation.cpp:549-555):
 * Alice iid uniform bits, Bob = Alice with EXACTLY floor(n * qber) flips at uniformly random distinct positions
 * (array_and_matrix_operations.cpp:889-933). Device RNG (counter-based), NOT the reference's xoshiro stream:
 * for throughput runs only. Writes packed frames to DEVICE buffers and returns the accurate QBER. */
QKDLDPC_API int qkdldpc_generate_keys_device(qkdldpc_code *code, int64_t n_frames, double qber, uint64_t seed,
                                 uint32_t *d_alice_bits, uint32_t *d_bob_bits, double *accurate_qber_out);

/* Throughput entry point (north_star: "throughput is measured on synthetic keys"): generates n_frames synthetic key pairs
 * on the device as qkdldpc_generate_keys_device does, decodes them with `params`, and returns the tally vector (HOST,
 * qkdldpc_tally_len() entries, may be NULL) and the decode time in seconds (CUDA events around the decode of the
 * device-resident frames; key generation excluded). No reference equivalent: the CPU program times each run_trial
 * (simulation.cpp:559-568). */
QKDLDPC_API int qkdldpc_bench_synthetic(qkdldpc_code *code, const qkdldpc_params *params, int64_t n_frames, double qber,
                            uint64_t seed, uint64_t *tally, double *seconds_out);

/* Trial inputs generated ON THE DEVICE, bit-identical to what the reference's run_trial builds on the CPU for the
 * same per-trial seed (simulation.cpp:549-555): prng = Xoshiro256++(trial_seeds[f] + seed_offset) -- seed_offset is the
 * running combination index the reference adds (simulation.cpp:743) --, fill_random_bits, inject_errors (libstdc++'s
 * std::shuffle reproduced draw for draw), and, when position lists are given, the extended frames of
 * QKD_LDPC_RATE_ADAPT (qkd_ldpc_algorithm.cpp:1148-1174: one more draw per party and punctured bit).
 * trial_seeds is a HOST array (simulation.cpp:713-719); the frames go to DEVICE buffers of n_frames x words_per_frame. */
QKDLDPC_API int qkdldpc_generate_trial_inputs_device(qkdldpc_code *code, int64_t n_frames, const uint64_t *trial_seeds,
                                         uint64_t seed_offset, double qber, const int32_t *punct_pos, int32_t n_punct,
                                         const int32_t *short_pos, int32_t n_short, uint32_t *d_alice_bits,
                                         uint32_t *d_bob_bits, double *accurate_qber_out);

/* The batched run_trial (simulation.cpp:540-577): generate the inputs of n_trials trials on the device as above and decode
 * them; nothing but the seeds goes in and the per-trial results / tallies come out (HOST pointers, each may be NULL).
 * This is the call that replaces  pool.detach_loop(0, TRIALS, n -> run_trial(matrix, QBER, seeds[n] + curr_sim, ...)).
 * Fails with QKDLDPC_ERR_INVALID ("Key size ... is too small for QBER.") where run_trial throws (floor(n*QBER) == 0). */
QKDLDPC_API int qkdldpc_run_trials(qkdldpc_code *code, const qkdldpc_params *params, int64_t n_trials, const uint64_t *trial_seeds,
                       uint64_t seed_offset, double qber, const int32_t *punct_pos, int32_t n_punct,
                       const int32_t *short_pos, int32_t n_short, uint32_t *out_bits, int32_t *out_iters,
                       uint8_t *out_flags, uint64_t *tally, double *accurate_qber_out);

/* One parameter combination of a sweep (sim_combination, simulation.hpp:22-27, plus its running index). */
typedef struct qkdldpc_combination {
    double qber;               /* configured QBER (config_QBER); the accurate one is floor(n * qber) / n          */
    double primary, secondary; /* decoding_scaling_factors of the combination                                    */
    const int32_t *punct_pos;  /* H_matrix_params.punctured_bits (may be NULL with n_punct == 0)                  */
    int32_t n_punct;
    const int32_t *short_pos;  /* H_matrix_params.shortened_bits                                                  */
    int32_t n_short;
    uint64_t seed_offset;      /* curr_sim: added to every trial seed (simulation.cpp:743)                        */
    const int32_t *remove_pos; /* H_matrix_params.bits_to_remove (strictly ascending; may be NULL with n_remove == 0):  */
    int32_t n_remove;          /* when given, remove_bits runs on the device after the decoder, as the last step of     */
    int32_t reserved;          /* QKD_LDPC / QKD_LDPC_RATE_ADAPT (qkd_ldpc_algorithm.cpp:1089-1092, 1218-1220)          */
} qkdldpc_combination;

/* The batched run_trial for SEVERAL combinations of one matrix in one call: the rate-adaptation
 * sweeps of the reference run thousands of combinations x ~100 trials, and 100 frames do not fill a B200. All
 * n_combinations x n_trials frames are generated and decoded by ONE launch each (per-frame QBER, scaling factors and
 * punctured / shortened masks); when no on-chip kernel can be used (float64, long codes, SPA on a graph too large for shared memory) the combinations are
 * processed one after the other. params->primary / secondary are ignored. Outputs (HOST, each may be NULL):
 * out_iters / out_flags: n_combinations x n_trials; tallies: n_combinations x qkdldpc_tally_len(); accurate_qber_out:
 * n_combinations. Results are identical to n_combinations calls of qkdldpc_run_trials. */
QKDLDPC_API int qkdldpc_run_trials_multi(qkdldpc_code *code, const qkdldpc_params *params, int32_t n_combinations,
                             const qkdldpc_combination *combinations, int64_t n_trials, const uint64_t *trial_seeds,
                             int32_t *out_iters, uint8_t *out_flags, uint64_t *tallies, double *accurate_qber_out);

/* The same call with the protocol's last step included: every combination that carries a removal list (remove_pos /
 * n_remove: the privacy-maintenance positions, or the punctured + shortened positions of rate adaptation) has its frames'
 * final keys built on the device by remove_bits -- Alice's (extended) key and bob_solution without the listed positions
 * (alice_bit_array_pm / bob_bit_array_pm, qkd_ldpc_algorithm.cpp:1089-1092; ..._rb, :1216-1220) -- inside the call, as the
 * reference's timed region around QKD_LDPC* includes it (simulation.cpp:559-568). out_alice_keys / out_bob_keys: NULL, or
 * n_combinations HOST pointers (entries may be NULL) receiving n_trials packed frames of n - n_remove bits each.
 * qkdldpc_run_trials_multi is this call with both NULL. */
QKDLDPC_API int qkdldpc_run_trials_multi_keys(qkdldpc_code *code, const qkdldpc_params *params, int32_t n_combinations,
                                  const qkdldpc_combination *combinations, int64_t n_trials, const uint64_t *trial_seeds,
                                  int32_t *out_iters, uint8_t *out_flags, uint64_t *tallies, double *accurate_qber_out,
                                  uint32_t *const *out_alice_keys, uint32_t *const *out_bob_keys);

/* remove_bits (array_and_matrix_operations.cpp:259-287; called by QKD_LDPC / QKD_LDPC_RATE_ADAPT after the decoder,
 * qkd_ldpc_algorithm.cpp:1092,1220): deletes the positions `bits_to_remove` (strictly ascending: the privacy-maintenance
 * list, or punctured + shortened positions, H_matrix_params.bits_to_remove) from every frame. keys: HOST, n_frames packed
 * frames of n bits; out_keys: HOST, n_frames packed frames of n - n_remove bits ((n - n_remove + 31) / 32 words each). */
QKDLDPC_API int qkdldpc_remove_bits(qkdldpc_code *code, int64_t n_frames, const uint32_t *keys, const int32_t *bits_to_remove,
                        int32_t n_remove, uint32_t *out_keys);

/* ---- Multi-GPU: frames are independent, so a batch is sharded by frame (contiguous trial ranges per device, no
 * per-iteration traffic) and the ONLY collective is the sum of the tally vectors, the vector process_trials_results
 * consumes (simulation.cpp:580-624). The handle owns the NCCL communicator. NCCL is bound at run time (libnccl.so.2);
 * without it these calls return QKDLDPC_ERR_STATE and everything else keeps working.
 *   one process per GPU (torchrun / MPI): rank 0 calls qkdldpc_comm_get_unique_id, the caller broadcasts the 128 bytes,
 *       every rank calls qkdldpc_comm_init_rank on its handle;
 *   one process, several devices (qkdldpc_sim --gpus N): qkdldpc_comm_init_all on one handle per device; the all-reduce
 *       is then called from one host thread per device at the same time.
 * Every rank of the communicator must make the same all-reduce calls with the same count. */
#define QKDLDPC_COMM_ID_BYTES 128
QKDLDPC_API int qkdldpc_comm_nccl_version(void); /* e.g. 22809; 0 when NCCL cannot be loaded */
QKDLDPC_API int qkdldpc_comm_get_unique_id(uint8_t *id_out /* QKDLDPC_COMM_ID_BYTES */);
QKDLDPC_API int qkdldpc_comm_init_rank(qkdldpc_code *code, const uint8_t *id, int32_t n_ranks, int32_t rank);
QKDLDPC_API int qkdldpc_comm_init_all(qkdldpc_code *const *codes, int32_t n_codes);
QKDLDPC_API int qkdldpc_comm_size(const qkdldpc_code *code); /* ranks of the handle's communicator, 0 without one */
/* Sum over all ranks, in place. HOST vector (copied to the device, reduced over NVLink, copied back; returns when done): */
QKDLDPC_API int qkdldpc_tally_allreduce(qkdldpc_code *code, uint64_t *tally, int64_t count);
/* DEVICE vector, enqueued on the handle's stream right behind the decode that filled it (no host synchronisation): */
QKDLDPC_API int qkdldpc_tally_allreduce_device(qkdldpc_code *code, uint64_t *d_tally, int64_t count);

/* Introspection used by benchmarks and tests. */
typedef struct qkdldpc_info {
    int32_t n, m;
    int64_t nnz;
    int32_t device;
    int32_t frames_per_tile;   /* of the last batch */
    int32_t pool_tiles;        /* of the last batch */
    int64_t pool_bytes;        /* message pool bytes of the last batch */
    int64_t kernel_launches;   /* kernels launched by this handle so far (counted at launch sites) */
    int64_t decoder_steps;     /* check-node/variable-node step pairs launched so far */
    double last_batch_ms;      /* device time of the last batch (CUDA events on the handle's stream) */
    double last_cn_ms, last_vn_ms, last_sched_ms; /* only filled when profiling is enabled */
    int32_t last_path;         /* decoder path of the last batch: 1 streaming, 2 on-chip */
    int32_t onchip_threads;    /* CTA size of the last on-chip launch */
    int32_t last_precision;    /* message precision the last batch ran in (32 / 64), after the precision policy */
    int32_t onchip_record_bytes; /* record format of the last float32 on-chip min-sum launch (16 / 8; 0 = none yet) */
    int64_t tail_compactions;  /* streaming path: tail compactions run by this handle so far */
    int32_t last_steps_per_poll; /* streaming path: decoder steps between two host polls in the last batch */
    int32_t last_vn_items_per_warp; /* streaming path: items per warp of the dv <= 4 variable-node kernel in the last batch (0: no such bucket) */
} qkdldpc_info;
QKDLDPC_API int qkdldpc_code_info(const qkdldpc_code *code, qkdldpc_info *info);
/* Host only (no device needed): builds the storage layouts of the on-chip min-sum kernels for a graph, checks every table
 * entry against the graph and returns the shared-memory bank model of the result, in wavefronts per decoder iteration:
 * out[0] = 1 when the code is eligible (else the rest is 0), out[1] / out[2] = the check phase's 4-byte gathers and their
 * conflict-free minimum, out[3] / out[4] = the variable phase's record gathers and their minimum (one wavefront per 128
 * bytes), out[5] = 1 when the 16-bit variable-phase table applies (at most 2048 records) -- for the 16-byte record format;
 * out[6..11] the same for the float32 kernel's 8-byte records. QKDLDPC_ERR_STATE when a table fails its self-check. */
QKDLDPC_API int qkdldpc_onchip_layout_model(int32_t n, int32_t m, int64_t nnz, const int32_t *row_ptr, const int32_t *col_idx,
                                            int64_t *out /* [12] */);
/* When enabled, every kernel of the step loop is bracketed by CUDA events (slow; for bench roofline numbers). */
QKDLDPC_API int qkdldpc_code_set_profiling(qkdldpc_code *code, int32_t enabled);

#ifdef __cplusplus
}
#endif
#endif /* QKDLDPC_H_ */
