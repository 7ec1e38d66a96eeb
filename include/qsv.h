/* qsv.h — C ABI of libqsv.so, the sm_100a state-vector gate-application library.
 *
 * This is the drop-in boundary for the reference's hot path.  The
 * reference (onofreiandrea/quantum_simulations, package wenbo_engine) has no FFI layer;
 * its seam is the pair of Python callables the runner dispatches to
 *     a1(chunk, qubit, U)            wenbo_engine/kernel/cpu_scalar.py:21  (cpu_batched.py:12)
 *     a2(chunk, qa, qb, U)           wenbo_engine/kernel/cpu_scalar.py:35  (cpu_batched.py:28)
 * chosen by string in wenbo_engine/runner/single_node.py:103-106, plus the four butterfly
 * functions of wenbo_engine/kernel/cpu_nonlocal.py:22-67 and the chunk store
 * (wenbo_engine/storage/block_store.py:18-65).  Every entry point below names the
 * reference interface it replaces.  INTEGRATION.md shows the ctypes stub a maintainer
 * adds to wenbo_engine to bind them (kernel="cuda").
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success and a negative
 *     QSV_E* code on failure; qsv_last_error() gives the message.
 *   - amplitudes are interleaved (re, im): float[2] (QSV_C64) or double[2] (QSV_C128),
 *     exactly numpy complex64 / complex128 memory.
 *   - LITTLE-ENDIAN qubits (wenbo_engine/circuit/io.py:3-6): qubit q is bit q of the
 *     amplitude index.
 *   - matrices are row-major complex, interleaved doubles: U[2*(r*dim+c)] = Re U[r][c].
 *     2-qubit matrices use the reference's sub-space order row = 2*bit(qa) + bit(qb)
 *     (wenbo_engine/kernel/gates.py:3-11).
 *   - a handle owns ONE shard: the 2^(n_qubits - log2(world)) amplitudes whose top
 *     log2(world) index bits equal `rank` (the reference's "chunk" with
 *     chunk_size = 2^n/world, wenbo_engine/docs/architecture.md:153-154).  Qubit numbers
 *     passed to the apply functions are PHYSICAL bit positions of the full index;
 *     positions >= n_local are rank bits.
 *   - all work is enqueued on the handle's stream; qsv_sync() waits for it.
 */
#ifndef QSV_H
#define QSV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSV_ABI_VERSION 8

/* dtypes (wenbo_engine/storage/block_store.py:11 fixes complex64; the oracle is complex128) */
#define QSV_C64  0
#define QSV_C128 1

/* error codes */
#define QSV_OK            0
#define QSV_EINVAL       -1   /* bad argument */
#define QSV_ENONLOCAL    -2   /* gate mixes a qubit that is a rank bit: remap first
                                 (NotImplementedError "non-local" of cpu_scalar.py:13-18) */
#define QSV_ECUDA        -3   /* CUDA runtime error (message has the cudaError string) */
#define QSV_ENOMEM       -4
#define QSV_ECOMM        -5   /* NCCL / peer-exchange error */
#define QSV_EIO          -6

typedef struct qsv_handle qsv_handle;
typedef struct qsv_program qsv_program;

/* ---------------------------------------------------------------- lifecycle ---- */
int qsv_abi_version(void);
/* Number of visible CUDA devices (0 => the library cannot run; there is no CPU path). */
int qsv_device_count(void);
/* Allocate the shard `rank` of `world` (power of two) of an n_qubits state on `device`. */
int qsv_create(qsv_handle **out, int n_qubits, int dtype, int device, int rank, int world);
int qsv_destroy(qsv_handle *h);
/* qsv_destroy parks a shard buffer >= 64 MiB (one per device) for the next qsv_create of the same
 * size: cudaFree + cudaMalloc of a 16 GiB buffer cost ~150 ms.  This frees the parked buffers. */
int qsv_release_cached(void);
/* Message of the last failure on h (h == NULL: last failure of qsv_create). */
const char *qsv_last_error(const qsv_handle *h);
int qsv_sync(qsv_handle *h);
/* Raw device pointer + CUDA stream of the shard (for zero-copy hand-off to NCCL/torch plumbing). */
int qsv_device_ptr(qsv_handle *h, void **ptr, size_t *n_amps_local, void **stream);

/* ---------------------------------------------------------------- state I/O ----
 * replaces init_zero_state / read_chunk / write_chunk_atomic (block_store.py:18-65) and
 * collect_state (single_node.py:326-346): offsets and counts are in amplitudes of the
 * LOCAL shard. */
int qsv_init_zero(qsv_handle *h);                      /* |0...0>: amp[0]=1 on rank 0 */
int qsv_init_basis(qsv_handle *h, uint64_t index);     /* |index> (global index)      */
int qsv_upload(qsv_handle *h, const void *host, size_t off_amps, size_t n_amps);
int qsv_download(qsv_handle *h, void *host, size_t off_amps, size_t n_amps);
/* Async variants for pinned host buffers (checkpoint path); complete at qsv_sync(). */
int qsv_upload_async(qsv_handle *h, const void *pinned_host, size_t off_amps, size_t n_amps);
int qsv_download_async(qsv_handle *h, void *pinned_host, size_t off_amps, size_t n_amps);
/* ASYNCHRONOUS CHECKPOINTS — the GPU form of the reference's reader / worker / writer pipeline
 * (wenbo_engine/runner/pipeline.py:50-82, 162-171): qsv_snapshot copies the shard into a second device buffer
 * (stream ordered, HBM speed); qsv_snapshot_download_async pulls ranges of that SNAPSHOT to pinned host memory
 * on a second stream (callable from a writer thread) while the handle's stream already runs the next passes;
 * qsv_snapshot_sync waits for those copies.  A new snapshot waits for the previous one to be drained. */
int qsv_snapshot(qsv_handle *h);
int qsv_snapshot_download_async(qsv_handle *h, void *host, size_t offset_amps, size_t n_amps);
int qsv_snapshot_sync(qsv_handle *h);
int qsv_host_alloc(void **ptr, size_t bytes);          /* cudaHostAlloc: pinned staging */
int qsv_host_free(void *ptr);

/* ------------------------------------------------------ per-gate operators ----
 * qsv_apply_1q  replaces a1 = cpu_scalar.apply_1q / cpu_batched.apply_1q and
 *               cpu_nonlocal.apply_1q_pair (any q < n_local is just a stride on a GPU).
 * qsv_apply_2q  replaces a2 = cpu_scalar.apply_2q / cpu_batched.apply_2q and
 *               cpu_nonlocal.apply_2q_pair_qa_local / _qb_local / apply_2q_quad.
 * Both return QSV_ENONLOCAL if a qubit they must mix is >= n_local. */
int qsv_apply_1q(qsv_handle *h, int q, const double U[8]);
int qsv_apply_2q(qsv_handle *h, int qa, int qb, const double U[32]);
/* Diagonal gate on nq <= 6 qubits: amp *= phases[sum_i bit(qs[i]) << (nq-1-i)] (qs[0] is the
 * most significant, like the 2-qubit sub-space order).  Z,S,T,R,CZ,CR of gates.py:33-40,47,64,72.
 * Rank bits are allowed (a per-shard constant: Atlas "insular" qubits, staging.py:65-98). */
int qsv_apply_diag(qsv_handle *h, int nq, const int *qs, const double *phases);
/* Controlled 1-qubit gate |0><0| (x) I + |1><1| (x) U: CNOT, CY, CU (gates.py:58-82).
 * ctrl may be a rank bit; tgt must be local. */
int qsv_apply_ctrl_1q(qsv_handle *h, int ctrl, int tgt, const double U[8]);
/* Dense k-qubit unitary, k <= 5, on local qubits qs[0..k) (qs[0] most significant row bit);
 * U is 2^k x 2^k.  Used for fused gate blocks (batch_levels/fuse_1q_ops, fusion.py:41-142). */
int qsv_apply_kq(qsv_handle *h, int k, const int *qs, const double *U);

/* ------------------------------------------------------------- fused passes ----
 * One pass = ONE read and ONE write of the shard applying a whole list of gates
 * (batch_levels, fusion.py:86-142, moved on chip).  A pass owns a TILE: 2^n_tile
 * amplitudes addressed by n_tile physical bit positions (tile_bits, ascending; the low
 * ones contiguous for coalescing).  It runs as a sequence of ROUNDS; in each round every
 * thread holds 2^QSV_REG_BITS amplitudes in registers (the round's reg_pos bits vary inside
 * a thread, the other tile bits are fixed per thread) and applies the round's ops; between
 * rounds amplitudes are exchanged through a swizzled shared-memory tile. */
#define QSV_MAX_TILE_BITS 14
#define QSV_REG_BITS       4
#define QSV_MAX_ROUNDS    16
#define QSV_MAX_ACTIVE_BITS 52

/* op kinds.  Every op is an IN-PLACE update of register-resident amplitudes (each arithmetic
 * statement overwrites one of its own operands), which is what lets the kernel interpret a gate
 * list without shuffling its 64 data registers.  The pass compiler lowers any 1-qubit unitary
 * to these primitives (U = e^{ia} diag(1,e^{ip}) Ry diag(1,e^{il}), ZYZ).
 * An op with a TARGET mixes the pairs (a, b) = (target bit 0, target bit 1); CONTROLS
 * (reg_ctrl / tile_ctrl / glob_ctrl) restrict it to amplitudes whose named bits are all 1. */
#define QSV_OP_HAD    0  /* (a, b) -> (a + b, a - b), UNNORMALISED (b <- a-b ; a <- 2a-b); the
                            compiler emits the 1/sqrt2 factors as one SCALE per pass; no controls */
#define QSV_OP_ROT    1  /* (a, b) -> (c a - s b, s a + c b), |theta| <= pi/2, as three shears
                            a -= t b ; b += s a ; a -= t b with m[0] = t = tan(theta/2), m[1] = s
                            (m[2], m[3] belong to the PREPHASE pre-op, see QSV_OPF_*)             */
#define QSV_OP_XSWAP  2  /* X: (a, b) -> (b, a), no arithmetic                                */
#define QSV_OP_YSWAP  3  /* Y: (a, b) -> (-i b, i a), no arithmetic                           */
#define QSV_OP_PHASE  4  /* no target: amp *= e^{i phi}, |phi| <= pi/2, where all controls are 1;
                            m[0] = tan(phi/2), m[1] = sin(phi), m[2] = cos(phi), m[3] = sin(phi) */
#define QSV_OP_SIGN   5  /* no target: amp = -amp where all controls are 1 (Z, CZ)            */
#define QSV_OP_SCALE  6  /* no target, no controls: amp *= m[0] (real)                        */
#define QSV_OP_TPHASE 7  /* TABLE PHASE on the b half of a register slot (target): every amplitude whose
                            target bit is 1 is multiplied by the unit complex number
                                T_thr[thread] * prod_k T_run_k[(global_index >> 8 r_k) & 255]
                            read from the pass's table array: m[0] = offset of the 2^(n_tile-4)-entry
                            per-thread table (-1: none), m[1] = offset of the first 256-entry table of
                            the 8-bit index runs r_k named by the bit mask m[2] (ascending; -1: none).
                            One op replaces ALL controlled-phase gates (CR, CZ) between the target and
                            qubits that are not register-resident in the round (the QFT pattern).     */
#define QSV_OP_KINDS  8

/* PRE-OPS of an UNCONTROLLED HAD / ROT (qsv_op.flags).  Diagonal gates that sit right before a
 * mixing gate on its target are folded into it by the compiler, so they cost no dispatch:
 * before (a, b) is mixed, the b half (target bit = 1) is multiplied by
 *     (-1)^( parity(tile_index & tile_ctrl) ^ parity(global_index & glob_ctrl) ^ PRENEG )   and then by
 *     e^{i phi},  m[2] = tan(phi/2), m[3] = sin(phi), |phi| <= pi/2                          (PREPHASE).
 * With any of these flags set reg_ctrl must be 0 and tile_ctrl / glob_ctrl are PARITY masks (the
 * partners of the CZ gates pending on the target), not AND-controls. */
#define QSV_OPF_PRESIGN   1  /* tile_ctrl / glob_ctrl are parity masks of a pre-sign              */
#define QSV_OPF_PRENEG    2  /* b = -b (a pending Z on the target)                                */
#define QSV_OPF_PREPHASE  4  /* b *= e^{i phi} with m[2], m[3]                                     */

typedef struct {
    uint8_t  kind;         /* QSV_OP_*                                                   */
    uint8_t  target;       /* register-slot index 0..QSV_REG_BITS-1 (kinds with a target) */
    uint8_t  reg_ctrl;     /* controls among register slots (bit b = slot b)             */
    uint8_t  flags;        /* QSV_OPF_* pre-ops (HAD / ROT without controls only), else 0 */
    uint32_t tile_ctrl;    /* controls among thread-fixed tile positions (bit i = pos i) */
    uint64_t glob_ctrl;    /* controls among physical bits outside the tile (rank bits ok)*/
    double   m[4];         /* coefficients, meaning depends on kind                      */
} qsv_op;                  /* 48 bytes, 16-byte aligned records                          */

typedef struct {
    uint8_t  reg_pos[QSV_REG_BITS];        /* tile positions held in registers, ascending */
    uint8_t  thr_pos[QSV_MAX_TILE_BITS];   /* tile position of each thread-index bit      */
    int32_t  op_begin, op_end;             /* slice of the pass's op array                */
    int32_t  fold_off;                     /* -1, or first entry of this round's FOLD TABLE in the
                                              pass's table array: 2^(n_tile-4) unit complex numbers
                                              (re, im), entry k = product of all diagonal ops of the
                                              round whose controls are thread-fixed tile bits, for
                                              the thread with index k.  One multiply per thread
                                              replaces the whole CZ / CR layer of the round.      */
} qsv_round;

typedef struct {
    int32_t   n_tile;                          /* tile bits (c128: 12, c64: 13 by default) */
    int32_t   load_bits[QSV_MAX_TILE_BITS];    /* physical bit of tile position i on load  */
    int32_t   store_bits[QSV_MAX_TILE_BITS];   /* ... on store (a permutation of load_bits:
                                                  free qubit relabelling inside the tile)  */
    int32_t   n_rounds;
    qsv_round rounds[QSV_MAX_ROUNDS];
    int32_t   n_ops;
    int32_t   n_fold;                          /* complex entries in this pass's fold-table array */
    int32_t   n_active;                        /* -1: the pass visits every tile.  >= 0: ZERO-SUPPORT SKIPPING — the
                                                  caller guarantees that every amplitude with a 1 at a non-tile
                                                  position outside active_bits[0..n_active) is exactly zero (a
                                                  run from |0...0> before those qubits were touched); the pass
                                                  then visits only the 2^n_active tiles that can hold data.
                                                  Zero tiles stay zero under any gate list, so the result is
                                                  identical.  Honoured by the specialised kernels; the
                                                  interpreting kernels visit every tile.                        */
    int32_t   active_bits[QSV_MAX_ACTIVE_BITS];/* ascending non-tile local positions                          */
    int32_t   zero_input;                      /* 1: FUSED |0...0> INITIALISATION — the pass does not read the
                                                  shard: it behaves as if it held |0...0> (amp[0] = 1 on rank 0,
                                                  zeros elsewhere), so a run from the zero state needs neither
                                                  the memset nor the read half of its first pass.  The
                                                  interpreting kernels initialise the shard themselves first.   */
    uint64_t  store_flip;                      /* physical tile bits XOR-ed into every store address:
                                                  pending X gates are never executed on data — the
                                                  compiler carries them as a Pauli frame and the last
                                                  pass that holds the qubit flips its index bit here */
} qsv_pass;

/* One-shot: run one pass now.  `tables` (may be NULL) holds pass->n_fold complex entries. */
int qsv_apply_pass(qsv_handle *h, const qsv_pass *pass, const qsv_op *ops, const double *tables);
/* Compiled circuit: upload all passes once, replay with one call. */
int qsv_program_create(qsv_handle *h, const qsv_pass *passes, int n_passes,
                       const qsv_op *ops /* concatenated */, const double *tables /* concatenated */,
                       qsv_program **out);
int qsv_program_run(qsv_handle *h, qsv_program *p);
int qsv_program_destroy(qsv_handle *h, qsv_program *p);
int qsv_program_run_range(qsv_handle *h, qsv_program *p, int first_pass, int n_passes);
/* Which passes of a program run on a specialised kernel (flags[i] = 1) — the host plumbing reduces this
 * over all ranks before it chooses a collective execution path (pipelined transitions need every rank to
 * take the same one). */
int qsv_program_specialised(qsv_handle *h, qsv_program *p, int *flags, int n_flags);
/* PIPELINED STAGE TRANSITION (replaces the reference's read / compute / write overlap,
 * wenbo_engine/runner/pipeline.py:50-82, and HiSVSIM's gather_qubits between parts,
 * hisvsim_repo/execute.hpp:665-685): the swap of rank bits global_bits[] with local bits local_bits[],
 * executed chunk by chunk together with the passes next to it.  chunk_bits[] (1..4 ascending local
 * positions, not swapped, in no tile of the named passes) cut the shard into 2^n_chunk chunks; for each
 * chunk: passes [a_first, a_first + a_count) of `pa`, then the exchange, then passes [0, b_count) of `pb`.
 * The pass launches use sm_count - xchg_sms SMs, the exchange kernels (TMA bulk copies through shared
 * memory, csrc/xchg.cuh) xchg_sms SMs on a second stream; cross-GPU ordering is by flag words in peer
 * memory (no NCCL, no host synchronisation).  Collective over the ranks of the swap group; every
 * condition that does not hold is an error (nothing has run then), never a per-rank fallback. */
int qsv_swap_pipelined(qsv_handle *h, qsv_program *pa, int a_first, int a_count, qsv_program *pb, int b_count,
                       int n_swap, const int *global_bits, const int *local_bits, int n_chunk, const int *chunk_bits,
                       int xchg_sms);
/* SCATTER PASS: pass `pass_index` FUSED with the exchange that follows it — one kernel computes the pass
 * and stores every amplitude where it lives AFTER the swap of local bits local_bits[i] with rank bits
 * global_bits[i]: into the second buffer ("shadow") of this GPU or, by plain stores through NVLink,
 * of a peer.  No separate swap kernel, no additional HBM sweep; the shard and its shadow exchange
 * roles afterwards (qsv_device_ptr changes).  Needs 2x the shard in HBM and the specialised kernels.
 *   qsv_shadow_ptr               allocate the shadow (first call) and return its device pointer
 *   qsv_comm_shadow_ipc_handle   64-byte CUDA IPC handle of the shadow (one process per GPU)
 *   qsv_comm_set_shadow_peers    world x 64 bytes of those handles; after qsv_comm_set_peers
 *   qsv_scatter_set_targets      the same wiring from raw device pointers valid in THIS process
 *                                (current[r], shadow[r] for every rank r; several shards of one process)
 *   qsv_pass_scatter_prepare     build / load the kernel ahead of time
 *   qsv_pass_scatter             run it; collective over the 2^n_swap ranks.  With an NCCL communicator
 *                                a stream-ordered barrier follows the kernel; without one (same-process
 *                                shards) the caller synchronises all handles before the next step.
 *                                *fused = 0: conditions not met, "pass, then qsv_swap_global_local" ran. */
int qsv_shadow_ptr(qsv_handle *h, void **ptr);
int qsv_comm_shadow_ipc_handle(qsv_handle *h, void *out64);
int qsv_comm_set_shadow_peers(qsv_handle *h, const void *handles);
int qsv_scatter_set_targets(qsv_handle *h, void *const *current, void *const *shadow);
int qsv_pass_scatter_prepare(qsv_handle *h, qsv_program *p, int pass_index, int n_swap, const int *local_bits);
int qsv_pass_scatter(qsv_handle *h, qsv_program *p, int pass_index, int n_swap, const int *global_bits,
                     const int *local_bits, int *fused);
/* qsv_program_create SPECIALISES each 2^11-amplitude pass (complex128 and complex64) at run time (NVRTC, sm_100a): the same ring
 * kernel with the pass's bit positions and op sequence as straight-line code and the coefficients
 * in the kernel-parameter bank; cubins are cached by pass STRUCTURE (in memory and under
 * <libqsv dir>/jit_cache, env QSV_JIT_CACHE=<dir>|0).  QSV_JIT=0 or QSV_OPT_JIT=0 keeps the
 * interpreting kernels; so does a missing libnvrtc (qsv_last_error says why). */
#define QSV_OPT_JIT          1   /* 1 (default) / 0                                        */
#define QSV_OPT_SIMPLE_PASS  2   /* 1: run passes on the one-CTA-per-tile kernel (tests)    */
#define QSV_OPT_PEER_SWAP    3   /* 1 (default): swaps go through mapped peer memory when available */
int qsv_set_option(qsv_handle *h, int option, long long value);
int qsv_jit_stats(int *compiled, int *disk_hits, int *mem_hits, int *failed, double *compile_seconds);
/* Generate + NVRTC-compile the specialised kernel of one pass WITHOUT a device (build check /
 * cache warm-up; nvrtc cross-compiles sm_100a on a CPU-only host).  log receives the error text. */
/* The CUDA source qsv_program_create would hand to NVRTC for this pass (inspection / profiling). */
int qsv_jit_source(const qsv_pass *pass, const qsv_op *ops, int dtype, char *out, size_t cap, size_t *needed);
int qsv_jit_build_pass(const qsv_pass *pass, const qsv_op *ops, int dtype, size_t *cubin_bytes, char *log, size_t log_cap);
/* The coefficient bank (C.c[i], in op order) the specialised kernel of this pass is launched with. */
int qsv_jit_coefs(const qsv_pass *pass, const qsv_op *ops, int dtype, double *out, size_t cap, size_t *needed);
/* The same two for the SCATTER variant of the pass (stores redirected by the swapped local bits). */
int qsv_jit_source_scatter(const qsv_pass *pass, const qsv_op *ops, int dtype, int n_swap, const int *local_bits,
                           char *out, size_t cap, size_t *needed);
int qsv_jit_build_scatter(const qsv_pass *pass, const qsv_op *ops, int dtype, int n_swap, const int *local_bits,
                          size_t *cubin_bytes, char *log, size_t log_cap);

/* --------------------------------------------------------------- reductions ---- */
int qsv_norm2(qsv_handle *h, double *out);   /* sum |amp|^2 of the LOCAL shard */
/* Deterministic sampling on the local shard (parity unpinned by the reference: it has no
 * sampler — SURVEY.md §2.4-6; definition frozen in oracle/ref_dense.py sample_indices). */
int qsv_sample(qsv_handle *h, uint64_t seed_unused, int shots, const double *sorted_u,
               uint64_t *out_indices);
/* Building blocks of the same definition for a sharded state (the exclusive scan over leaf sums is
 * ONE sequential chain over all shards in rank order, run by the host plumbing):
 *   qsv_leaf_sums          sums of the shard's blocks of 2^10 amplitudes -> out_host[n_local_amps >> 10]
 *   qsv_sample_in_leaves   for each shot: walk local leaf leaf_idx[s] from its exclusive prefix leaf_off[s]
 *                          and return the first LOCAL index whose running sum exceeds x[s]           */
int qsv_leaf_sums(qsv_handle *h, double *out_host);
int qsv_sample_in_leaves(qsv_handle *h, int shots, const uint64_t *leaf_idx, const double *leaf_off,
                         const double *x, uint64_t *out_local_index);

/* Observables after the path (new; definitions in oracle/ref_dense.py): contributions of the LOCAL
 * shard, the caller sums over shards.  qubits are physical positions (rank bits allowed).
 *   qsv_probabilities  marginal distribution over nq <= 20 qubits, out[o] with bit k of o = qubit qubits[k]
 *   qsv_expect_z       <prod_{q in mask} Z_q> = sum_i |amp_i|^2 (-1)^parity(i & mask)               */
int qsv_probabilities(qsv_handle *h, int nq, const int *qubits, double *out_host);
int qsv_expect_z(qsv_handle *h, uint64_t mask, double *out);

/* ---------------------------------------------------- multi-GPU qubit remap ----
 * HiSVSIM-style redistribution (hisvsim_repo/mpi_redistributer.hpp:265-344,
 * svsim-mpi.hpp:123-173): swap `n_swap` (<= 3) rank bits with local bits.  local_bits may be any
 * distinct local positions on the peer-memory path; the ncclSend/ncclRecv path needs them to be the
 * top n_swap local positions in order (contiguous blocks). */
int qsv_comm_unique_id(void *id128);                       /* ncclGetUniqueId, 128 bytes   */
int qsv_comm_init(qsv_handle *h, const void *id128);       /* ncclCommInitRank(world,rank) */
int qsv_swap_global_local(qsv_handle *h, int n_swap, const int *global_bits, const int *local_bits);
/* Peer-memory path of the swap (one box, one process per GPU): every rank exports its shard with
 * CUDA IPC (64-byte handle), the host plumbing all-gathers the handles, qsv_comm_set_peers maps them.
 * The swap is then ONE kernel that exchanges the block pairs in place with loads/stores through
 * NVLink (no bounce buffer, no NCCL data path; NCCL only provides the two stream-ordered barriers).
 * Without it (or QSV_OPT_PEER_SWAP=0) the chunked ncclSend/ncclRecv exchange is used. */
int qsv_comm_ipc_handle(qsv_handle *h, void *out64);
int qsv_comm_set_peers(qsv_handle *h, const void *handles /* world x 64 bytes */);
/* The same wiring for SHARDS OF ONE PROCESS (handles of all ranks created by this process, on one or
 * several devices it can address): no NCCL communicator, no IPC — shards[r] = qsv_device_ptr of rank r.
 * Swaps then run on the exchange kernels of csrc/xchg.cuh, which order themselves through flag words. */
int qsv_comm_init_local(qsv_handle *h);
int qsv_comm_set_peers_local(qsv_handle *h, void *const *shards /* world pointers */);
/* Sum one double across all shards (norm, sampling offsets). world==1: no-op. */
int qsv_allreduce_sum(qsv_handle *h, double *value);

/* ------------------------------------------------------------------- timing ---- */
typedef struct { float ms; int32_t kind; int32_t pass_index; } qsv_timing;
int qsv_timing_enable(qsv_handle *h, int on);
int qsv_get_timings(qsv_handle *h, qsv_timing *out, int max, int *n_out);
/* One CUDA-event stopwatch on the handle's stream (bench.py's timed region). */
int qsv_timer_start(qsv_handle *h);
int qsv_timer_stop(qsv_handle *h, float *elapsed_ms);   /* records, synchronises, reads */

#ifdef __cplusplus
}
#endif
#endif /* QSV_H */
