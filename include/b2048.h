/*
 * b2048.h -- C ABI of lib2048nn_b200.so: the B200 (sm_100a) playout engine for fqjin/2048NN.
 *
 * The reference has no FFI of its own (it is pure Python); this header is the boundary a
 * maintainer binds instead of the Python hot loops.  Each entry point names the reference
 * code it replaces (file:line into the reference repo).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all work is
 *     enqueued on it, nothing synchronises the host and nothing allocates after b2048_create
 *     (b2048_take_error_flag is the one call that synchronises);
 *   - a context belongs to one device; every call makes that device current for its own duration and restores
 *     the caller's current device, so contexts of several GPUs can be driven from one thread;
 *   - calls on one context may be issued from several host threads and on several streams concurrently: the
 *     per-launch work cursors come from an atomic rotating pool and no launch shares scratch memory with another
 *     (the error flag is per context: it says that SOME launch since the last b2048_take_error_flag hit the
 *     reference's ValueError case);
 *   - return value: 0 = ok, non-zero = error (b2048_last_error() gives the text);
 *   - boards are uint64: nibble k (bits 4k..4k+3) is the tile exponent of cell
 *     (row k/4, col k%4), nibble 0 = top-left (board.py:23-33, 228-233);
 *   - directions: 0 Left, 1 Up, 2 Right, 3 Down, taken mod 4 (board.py:151-189);
 *   - scores are uint32 (a move scores at most 524,288; real games stay far below 2^32).
 *
 * RNG protocol (replaces CPython's MT19937 `randrange`, board.py:6,70-98).
 *   Every game owns a counter-based stream:
 *       mix64(z): z=(z^(z>>30))*0xBF58476D1CE4E5B9; z=(z^(z>>27))*0x94D049BB133111EB; z^(z>>31)
 *       key       = mix64(mix64(seed + G) ^ (game_id*0xD1342543DE82EF95 + 1)),  G = 0x9E3779B97F4A7C15
 *       draw(i)   = mix64(key + (i+1)*G)                      (64 bits = two 32-bit words: hi, lo)
 *       randrange(n) := (word * n) >> 32
 *   generate_init_tiles (board.py:68-82) consumes words 0.. of draws 0..7 in the reference's
 *   call order (pos1, pos2 [redrawn while equal], tile1, tile2).  The k-th generate_tile of a
 *   game (board.py:85-101) uses draw 8+k: hi word -> position, lo word -> randrange(10).
 *   Root-search line j of root move i of evaluation e (mcts.py:23, mcts_nn.py:22) is game
 *   game_id = (e*4 + i)*number + j; its first spawn is k = 0.
 */
#ifndef B2048_H
#define B2048_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2048_ctx b2048_ctx;

#define B2048_ABI_VERSION 1

/* reduction modes of b2048_mcts_fixed / b2048_mcts_nn */
#define B2048_MODE_MEANLOG 0 /* mcts.py:4-34   mcts_fixed (README: mcts_fixed_batch) */
#define B2048_MODE_MEDIAN 1  /* mcts.py:37-70  mcts_fixed_moves */
#define B2048_MODE_MIN 2     /* mcts.py:73-103 mcts_fixed_min; mcts_nn.py:4-43 mcts_nn */

/* fixed-point scale of the fused log10 sums (order-independent, shard-invariant) */
#define B2048_LOGSUM_FRAC_BITS 40

int b2048_abi_version(void);
const char *b2048_last_error(void);

/* Per-device context: row tables in device memory (replaces the merge_table /
 * merge_table_rev build + pickle cache, board.py:104-148) and launch scratch. */
int b2048_create(int device, b2048_ctx **out);
int b2048_destroy(b2048_ctx *ctx);
int b2048_num_sms(const b2048_ctx *ctx);
/* kernels launched through this context since b2048_create (launch accounting of bench.py's gpu_launches) */
int64_t b2048_launch_count(const b2048_ctx *ctx);

/* Measurement aid for the K2 roofline (SURVEY 8d bounds the fixed-order playout by SM integer issue): one launch of a
 * pipe-rate probe on `stream`, #SMs CTAs x 1024 threads x iters x 64 probe instructions.  kind 0 = LOP3 only (the ALU pipe,
 * which ncu shows as K2's first binder), kind 1 = LOP3 and IMAD alternating (ALU + FMA pipes: the issue limit).  `sink` is a
 * 4-byte device scratch word; *out_thread_instr (host) receives the number of probe thread-instructions of the launch.  The
 * caller times the launch with events: thread-instructions / second is the measured denominator bench.py reports. */
int b2048_probe_pipe_rate(b2048_ctx *ctx, int kind, uint32_t iters, uint32_t *sink, uint64_t *out_thread_instr,
                          void *stream);

/* Fixed-order playout kernel variant: 2 (default) = both slides of a row in one shared-memory word, the four
 * candidate boards from 8 lookups; 1 = legality table first, then one uniform table slide per move; 0 = the
 * L,U,D,R attempt cascade.  All are bit-exact; the switch exists for A/B measurements. */
int b2048_set_playout_variant(b2048_ctx *ctx, int variant);

/* Latency variants of the tensor-core kernels for small launches (results are identical in every mode):
 *   mode 1  launches with at most one 32-board tile per SM (forward of <= 32 x #SMs boards, root evaluations / main games
 *           with <= 32 x #SMs lines; main games up to 2.5x that) run one tile-set per CTA with all 16 worker warps on its
 *           epilogue;
 *   mode 2  additionally, playouts with at most 6 x #SMs lines in flight run with two-ply speculation: a line's current
 *           board and its four successors are evaluated in one forward, so every forward advances the line by two policy
 *           moves (half the dependent forwards per mcts_nn call); and in the TAIL of any larger playout launch (work queue
 *           empty, at most 6 / 2 lines left in a 32-slot tile) the survivors are regrouped into the two- / three-ply layout;
 *   mode 3  (default) additionally, playouts with at most 2 x #SMs lines in flight (BASELINE config 3: 200 lines) run with
 *           three-ply speculation: the line's board, its four successors and up to 11 of their successors in one forward,
 *           up to three policy moves per forward;
 *   mode 0  all off (A/B measurements). */
int b2048_set_latency_mode(b2048_ctx *ctx, int mode);

/* Device error flag: set when a spawn met countzero == 0 (the reference raises ValueError from
 * randrange(0), board.py:87).  Reads and clears it into *out_host; synchronises `stream`. */
int b2048_take_error_flag(b2048_ctx *ctx, void *stream, int *out_host);

/* Row tables as built on the device: fin_fwd/fin_rev u16[65536], score u32[65536] (identical
 * for both tables), moved_fwd/moved_rev u8[65536].  Any pointer may be NULL.  (board.py:138-148) */
int b2048_tables(b2048_ctx *ctx, uint16_t *fin_fwd, uint16_t *fin_rev, uint32_t *score, uint8_t *moved_fwd,
                 uint8_t *moved_rev, void *stream);

/* K1: move() for N boards (board.py:151-189).  dirs == NULL -> every board uses dir_scalar. */
int b2048_move_batch(b2048_ctx *ctx, const uint64_t *boards, const uint8_t *dirs, int dir_scalar,
                     uint64_t *out_boards, uint32_t *out_score, uint8_t *out_moved, int64_t n, void *stream);

/* transpose (board.py:46-54), countzero (board.py:57-65; 16 empties -> 0) and the legal-move
 * mask (bit d set when move(board,d) moves) for N boards.  Output pointers may be NULL. */
int b2048_board_ops(b2048_ctx *ctx, const uint64_t *boards, uint64_t *out_transpose, uint8_t *out_countzero,
                    uint8_t *out_legal_mask, int64_t n, void *stream);

/* generate_init_tiles (board.py:68-82) for games gid_base..gid_base+n-1 of `seed`. */
int b2048_init_boards(b2048_ctx *ctx, uint64_t seed, uint64_t gid_base, uint64_t *out_boards, int64_t n,
                      void *stream);

/* generate_tile (board.py:85-101): out[g] = spawn number k[g] (k == NULL -> k_scalar) of stream
 * (seed, gids[g]) (gids == NULL -> gid_base+g) placed on boards[g].  A board with
 * countzero == 0 (randrange(0) -> ValueError in the reference) yields out = 0 and sets
 * err_flag[0] = 1 when err_flag != NULL. */
int b2048_spawn_batch(b2048_ctx *ctx, const uint64_t *boards, uint64_t seed, const uint64_t *gids,
                      uint64_t gid_base, const uint32_t *k, uint32_t k_scalar, uint64_t *out_boards,
                      int32_t *err_flag, int64_t n, void *stream);

/* One BoardArray.move_batch / fast_move_batch step (board.py:246-303) for N games: game g plays
 * the first legal direction of its priority row prio[4g..4g+3] (uint8, 4-byte aligned rows),
 * then spawn number k[g] of stream (seed, gids[g] or gid_base+g).  out_moved[g] = 0 means the
 * game is dead (board unchanged, gain 0).  Compaction of dead games is the caller's. */
int b2048_policy_step(b2048_ctx *ctx, const uint64_t *boards, const uint8_t *prio, uint64_t seed,
                      const uint64_t *gids, uint64_t gid_base, const uint32_t *k, uint64_t *out_boards,
                      uint32_t *out_gain, uint8_t *out_moved, int64_t n, void *stream);

/* Training-record encodings (game_dataset.py:6-48).
 *   mode 0: tile exponents as float32[N*16] (position-major, game_dataset.py:15-21)
 *   mode 1: one-hot planes float32[N*16*16] = (N,16,4,4), channel c is 1 where the cell's exponent is c
 *           (game_dataset.py:46-48; the same tensor mcts_nn.py:25-35 feeds the network) */
int b2048_encode_boards(b2048_ctx *ctx, const uint64_t *boards, int mode, float *out, int64_t n, void *stream);

/* Soft targets of a self-play record: softmax(results / rowmax(results) * soft) over the 4 root-move values
 * (game_dataset.py:25-28).  results, out: float32[N*4]. */
int b2048_soft_targets(b2048_ctx *ctx, const float *results, float soft, float *out, int64_t n, void *stream);

/* Distribution summary of n finished games (notebooks/distribution.py:11-14 rows [max tile exponent, score, moves]
 * and the notebooks' histograms / mean log score), ACCUMULATED into acc20 (caller zeroes it):
 *   acc20[0..15] games by max tile exponent, [16] sum of round(log10(score+1) * 2^32) (no overflow below 2^29 games
 *   per accumulator), [17] sum of moves,
 *   [18] max score, [19] games.  Integer atomics only: independent of scheduling and of how games are sharded
 *   (ranks add their acc20 vectors, MAX for [18]).  out_maxtile u8[n] and moves may be NULL. */
int b2048_game_summary(b2048_ctx *ctx, const uint64_t *finals, const uint32_t *score, const uint32_t *moves,
                       uint8_t *out_maxtile, uint64_t *acc20, int64_t n, void *stream);

/* raw stream words for parity tests: out[g*count + j] = draw(first + j) of stream (seed, gid_base+g) */
int b2048_rng_draws(b2048_ctx *ctx, uint64_t seed, uint64_t gid_base, uint64_t first, int count,
                    uint64_t *out, int64_t n, void *stream);

/* K2: N fixed-order games played to the end inside one persistent kernel
 * (board.py:192-225 play_fixed; board.py:246-277 + 306-315 BoardArray.move_batch /
 * play_fixed_batch iterated to termination).
 *   start      N boards, or NULL for fresh boards (generate_init_tiles); a 0 entry also means
 *              "fresh board" (board.py:205 `if not board`)
 *   order4     HOST pointer to 4 direction indices, NULL = (0,1,3,2) = L,U,D,R (board.py:214)
 *   spawn0     index of the first spawn draw (0 for whole games)
 * outputs (each may be NULL), indexed by game: final board, score gained, accepted moves. */
int b2048_playout_fixed(b2048_ctx *ctx, const uint64_t *start, const uint8_t *order4_host, uint64_t seed,
                        uint64_t gid_base, uint32_t spawn0, uint64_t *out_final, uint32_t *out_score,
                        uint32_t *out_moves, int64_t n, void *stream);

/* Root search with fixed-order playouts for G origins x 4 root moves x `number` lines
 * (mcts.py:4-103).  Lines [line_begin, line_end) of every root move are played by this call
 * (a rank plays its shard; the union over ranks is [0, number)).
 *   out_legal   u8[G*4]   move(origin, i) moved
 *   out_rscore  u32[G*4]  root-move merge score s (mcts.py:21,31)
 *   out_lscore  u32[G*4*number]  per-line playout score (BoardArray scores start at 0), or NULL
 *   out_lmoves  u32[G*4*number]  per-line accepted moves L_j after the first spawn, or NULL
 *   out_logsum  i64[G*4]  ACCUMULATED: sum over the lines played of
 *                         round(log10(score + s + 1) * 2^B2048_LOGSUM_FRAC_BITS), or NULL
 *   out_count   i32[G*4]  ACCUMULATED: lines played, or NULL
 *   out_minmov  i32[G*4]  ACCUMULATED (min): min_j L_j (caller initialises to INT32_MAX), or NULL
 * Accumulated outputs are what an allreduce (SUM / SUM / MIN) combines across ranks. */
int b2048_mcts_fixed(b2048_ctx *ctx, const uint64_t *origins, int64_t G, int number, int line_begin,
                     int line_end, uint64_t seed, uint64_t eval_id_base, uint8_t *out_legal,
                     uint32_t *out_rscore, uint32_t *out_lscore, uint32_t *out_lmoves, int64_t *out_logsum,
                     int32_t *out_count, int32_t *out_minmov, void *stream);

/* ------------------------------------------------------------------------------------------
 * Policy network c64b3 = ConvNet(channels=64, blocks=3) (network.py:43-87) and NN-policy
 * playouts.  `blob` is the folded-weight image made by 2048nn_b200/convnet.py from a reference
 * state_dict (BatchNorm folded in float64, eval mode, eps 1e-5).
 *   B2048_PREC_FP32: CUDA-core fp32 forward, bit-close to PyTorch fp32 (exact mode)
 *   B2048_PREC_FP16_TC: tcgen05 tensor-core forward, fp16 operands, fp32 accumulation (throughput mode: argmax agreement
 *                       99.96 %, 99.985 % of boards within 1e-2 of the fp32 reference, max 0.024 on the shipped weights)
 *   B2048_PREC_FP16X2_TC: the same kernels with split operands on the sensitive layers: max |dlogp| < 1e-2 on every board
 * ------------------------------------------------------------------------------------------ */
#define B2048_PREC_FP32 0
#define B2048_PREC_FP16_TC 1
#define B2048_PREC_BF16_TC 2 /* same kernels with bf16 operands: fails the parity bar on the shipped weights, for comparison */
#define B2048_PREC_FP16X2_TC 3 /* precise tensor-core mode: layers L1..L3 take their weights, L1..L2 their activations as fp16
                                 hi+lo pairs on the 32 input channels that carry the rounding error (the blob's channels are
                                 ordered for it, 2048nn_b200/convnet.py): every board of the 150k reference fixture within
                                 1e-2 (max 0.0077); one tile-set per CTA; its own (larger) weight blob */

/* size in bytes of the weight blob for a precision (-1: unsupported) */
int64_t b2048_convnet_blob_bytes(int precision);

/* ConvNet.forward fused with the inline one-hot encoder (mcts_nn.py:25-35, network.py:78-87):
 * boards u64[N] -> log-probabilities f32[N*4] (order Left, Up, Right, Down). */
int b2048_convnet_forward(b2048_ctx *ctx, int precision, const void *blob, const uint64_t *boards,
                          float *out_logp, int64_t n, void *stream);

/* Bring-up/diagnostic entry of the tensor-core forward: pre-softmax outputs f32[N*4] and, when
 * dbg != NULL, the fp32 post-activation output of conv layer dbg_layer (0 = in_block, 1..6 = the
 * block convs) as [board][64 channels][16 positions]. */
int b2048_convnet_tc_debug(b2048_ctx *ctx, const void *blob, const uint64_t *boards, float *out_logits, float *dbg,
                           int dbg_layer, int64_t n, void *stream);

/* N policy games inside one persistent kernel (eval_nn.py:4-58 eval_nn, :61-95 eval_nn_min,
 * mcts_nn.py:46-94 play_nn): each move is the first legal direction in descending order of the
 * network's outputs (mcts_nn.py:36-37 + board.py:264-269; ties -> lower index).
 *   start       N start boards or NULL (fresh); 0 entry -> fresh board
 *   group_size  > 0 with group_min != NULL: games [k*group_size, (k+1)*group_size) form a group
 *               whose min accepted-move count is accumulated (atomic min) into group_min[k]
 *               (caller initialises to INT32_MAX); games stop early once they cannot lower it
 *               (eval_nn_min's first-death semantics).  Otherwise every game runs to its end.
 *   move_counter  optional u64[1], += accepted moves played (throughput accounting)
 * outputs per game (may be NULL): final board, score, accepted moves. */
int b2048_playout_nn(b2048_ctx *ctx, int precision, const void *blob, const uint64_t *start, uint64_t seed,
                     uint64_t gid_base, uint32_t spawn0, int group_size, int32_t *group_min, uint64_t *out_final,
                     uint32_t *out_score, uint32_t *out_moves, uint64_t *move_counter, int64_t n, void *stream);

/* The same policy-game evaluation for every network of a population in ONE launch per <= num_sms models
 * (evolve.py:30-36 NetworkPopulation.eval calls eval_nn(model, number=eval_num) model after model; 100 games
 * fill 1 % of a B200, 50 networks side by side fill it).  Every CTA is bound to one model and streams that
 * model's weight image; work items, spawn streams and results are those of b2048_playout_nn run per model:
 *   blobs             n_models weight images of `precision`, image m at blobs + m*blob_stride (16-byte multiple)
 *   gid_model_stride  game id of game j of model m = gid_base + m*gid_model_stride + j (0: common random
 *                     numbers -- every model plays the same spawn streams)
 *   start             n_per_model start boards shared by all models, or NULL (fresh boards)
 *   group_size/group_min, outputs: as b2048_playout_nn, indexed m*n_per_model + j (group_min by
 *                     (m*n_per_model + j) / group_size; n_per_model must be a multiple of group_size). */
int b2048_playout_nn_population(b2048_ctx *ctx, int precision, const void *blobs, int64_t blob_stride, int n_models,
                                const uint64_t *start, uint64_t seed, uint64_t gid_base, uint64_t gid_model_stride,
                                int group_size, int32_t *group_min, uint64_t *out_final, uint32_t *out_score,
                                uint32_t *out_moves, uint64_t *move_counter, int64_t n_per_model, void *stream);

/* mcts_nn (mcts_nn.py:4-43) for G origins: lines [line_begin,line_end) of `number` per root move.
 *   out_legal u8[G*4], out_rscore u32[G*4] (root merge score; the reference discards it),
 *   out_lmoves u32[G*4*number] per-line moves at stop (may be NULL; lines stopped early report
 *   the count at which they stopped), out_minmov i32[G*4] ACCUMULATED min_j L_j (caller
 *   initialises to INT32_MAX).  mcts_nn's value is 1 + min (0 when illegal).  Ranks allreduce
 *   out_minmov with MIN. */
int b2048_mcts_nn(b2048_ctx *ctx, int precision, const void *blob, const uint64_t *origins, int64_t G, int number,
                  int line_begin, int line_end, uint64_t seed, uint64_t eval_id_base, uint8_t *out_legal,
                  uint32_t *out_rscore, uint32_t *out_lmoves, int32_t *out_minmov, uint64_t *move_counter,
                  void *stream);

/* Whole main games of the fixed-order root search on the device: play_mcts_fixed (mcts.py:106-148; modes
 * B2048_MODE_MEANLOG / MEDIAN / MIN = mcts_fixed / mcts_fixed_moves / mcts_fixed_min) and selfplay_fixed
 * (selfplay.py:9-59: MIN mode averaged over `times` evaluations).  One CTA per main game, no host round trip per
 * main move; the streams are those of b2048_mcts_fixed + b2048_spawn_batch driven by the host loop, so the games are
 * identical:  main game m = game_id_base + g: start board / main spawns from stream (seed, m); evaluation t at main
 * step s has eval_id = (s*times + t + 1)*games_total + m.  games_total = games of the whole session (ranks play
 * disjoint [game_id_base, game_id_base+G) ranges of it).
 *   starts        G start boards or NULL; 0 entry -> fresh board
 *   out_boards    u64[G*max_steps] board before each main move (may be NULL), out_results f32[G*max_steps*4] the
 *                 root values it was chosen from (may be NULL): the npz record of selfplay.py:55-58
 *   out_final/out_score/out_steps [G]; out_status[G]: 0 game over, 1 stopped because max_steps was reached
 *   move_counter  optional u64[1], += playout moves played */
int b2048_selfplay_fixed(b2048_ctx *ctx, const uint64_t *starts, int64_t G, uint64_t game_id_base, uint64_t games_total,
                         int number, int times, int mode, uint64_t seed, int max_steps, uint64_t *out_boards,
                         float *out_results, uint64_t *out_final, int64_t *out_score, int32_t *out_steps,
                         int32_t *out_status, uint64_t *move_counter, void *stream);

/* Whole main games of the NN-policy root search on the device: selfplay() (selfplay.py:62-115) = per main move
 * mcts_nn(model, board, number) (mcts_nn.py:4-43), np.argmax (first maximum), move, generate_tile, until the chosen
 * move does not move.  One persistent launch plays G main games to the end (or to max_steps main moves) with NO
 * host round trip and NO grid-wide barrier per main move: the lines of every root evaluation are work items of a
 * device-side queue, a free game slot takes the next queued line of any main game, and the line that retires last
 * advances its main game and queues the next evaluation.  Streams are those of the step-by-step host loop
 * (b2048_mcts_nn + b2048_move_batch + b2048_spawn_batch), so the games are identical: main game
 * m = game_id_base + g takes its start board / main spawns from stream (seed, m); the evaluation at main step s has
 * eval_id = (s + 1)*games_total + m (line j of root move i: game (eval_id*4 + i)*number + j, first spawn k = 0).
 *   workspace     device memory of b2048_selfplay_nn_workspace_bytes(G, number) bytes, 256-byte aligned (queue +
 *                 game states; caller-allocated, nothing is allocated here)
 *   out_boards    u64[G*max_steps] board before each main move, out_results f32[G*max_steps*4] = mcts_nn's list
 *                 (1 + fewest moves survived; 0 illegal) it was chosen from -- the npz record of selfplay.py:111-114
 *                 (each may be NULL)
 *   out_final/out_score (i64)/out_steps [G]; out_status[G]: 0 game over, 1 stopped at max_steps
 *   move_counter  optional u64[1], += policy moves played by the lines */
int64_t b2048_selfplay_nn_workspace_bytes(int64_t G, int number);
int b2048_selfplay_nn(b2048_ctx *ctx, int precision, const void *blob, const uint64_t *starts, int64_t G,
                      uint64_t game_id_base, uint64_t games_total, int number, uint64_t seed, int max_steps,
                      uint64_t *out_boards, float *out_results, uint64_t *out_final, int64_t *out_score,
                      int32_t *out_steps, int32_t *out_status, uint64_t *move_counter, void *workspace, void *stream);

/* ---- c64b3 training step (trainer.py:14-97) -------------------------------------------------
 * One optimisation step of trainer.py:55-62 -- m(x) in train mode (BatchNorm batch statistics, running-stat
 * update), KLDivLoss(reduction='batchmean') against soft targets, backward, Adam -- in fp32, hand-written
 * (no autograd / cuDNN).  Flat fp32 vectors, `parameters()` order of network.py:43-76:
 *   params[231820]   in_block conv [64][16][3][3], BN weight, BN bias; 3 x (conv, BN w, BN b, conv, BN w, BN b);
 *                    out_block conv [4][64], BN w[4], BN b[4]; policy weight [4][64], bias [4]
 *   bn_buffers[1024] running_mean [8 BN layers][64] then running_var [8][64] (out_block BN: first 4 of its 64)
 *   adam_m, adam_v   [231820] first / second moment
 * workspace: caller-allocated device memory of b2048_train_workspace_bytes(max_batch) bytes (no hidden
 * allocation); after forward_backward it holds the gradient (float[231820] at b2048_train_grad_offset) and the
 * batch's log-probabilities (float[B*4] at b2048_train_logp_offset).
 * forward_backward: boards u64[B] (the one-hot encoder of game_dataset.py:46-48 is fused), targets f32[B*4];
 *   loss_out device float[1].  Deterministic: batch reductions are summed in a fixed order.
 *   mixed = 0: every convolution in fp32 on the FFMA2 pipe (the reference's arithmetic; parity mode).
 *   mixed = 1: the forward and data-gradient convolutions on the tcgen05 tensor cores (fp16 / bf16 operands, fp32
 *              accumulation), weight gradient, BatchNorm, head and Adam in fp32 -- PyTorch-autocast-like numerics.
 * adam: torch.optim.Adam's update (L2 weight decay added to the gradient, bias correction, eps outside the
 *   square root) for 1-based `step`.  Kept separate so that data-parallel ranks can allreduce the gradient
 *   between the two calls. */
int64_t b2048_train_workspace_bytes(int64_t max_batch);
int64_t b2048_train_grad_offset(int64_t max_batch);
int64_t b2048_train_logp_offset(int64_t max_batch);
int b2048_train_forward_backward(b2048_ctx *ctx, const float *params, float *bn_buffers, void *workspace,
                                 int64_t max_batch, const uint64_t *boards, const float *targets, int64_t B,
                                 float bn_momentum, int mixed, float *loss_out, void *stream);
int b2048_train_adam(b2048_ctx *ctx, float *params, const float *grad, float *adam_m, float *adam_v, double lr,
                     double beta1, double beta2, double eps, double weight_decay, int64_t step, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2048_H */
