/* dqc_engine.h -- C-ABI of the B200 quantum-state engine (libdqcengine.so).
 *
 * This is the drop-in boundary for the one hot path under km-campbell/dqc_simulator:
 * evolving the joint ket / density-matrix state that NetSquid's qubit formalism maintains
 * while a partitioned circuit runs.  The reference itself contains no arithmetic; it calls
 * the closed NetSquid wheel through
 *
 *   QuantumProcessor.execute_program(QuantumProgram)   software/dqc_control.py:269,300,317,
 *                                                      367,391,409,460,474,556,608,778,786,945
 *                                                      software/physical_layer.py:719,740
 *   qubitapi.stochastic_operate / apply_pauli_noise /
 *            assign_qstate                             hardware/noise_models.py:130,305,364
 *   QSource + StateSampler (2-qubit state injection)   hardware/connections.py:145-150
 *   qapi.reduced_dm / qapi.fidelity                    util/helper.py:151,209-214
 *
 * Each entry point below names the reference interface it replaces.  Plain pointers and
 * sizes only; no torch / C++ types cross the boundary.  Complex numbers are interleaved
 * (re, im) doubles.  Matrices are row-major; for multi-qubit operators the FIRST listed
 * axis is the MOST significant index (NetSquid's big-endian convention, pinned by
 * tests/unit_tests/qlib/test_circuit_identities.py:93-96).
 *
 * Model
 * -----
 * A handle owns a batch of `batch` independent registers ("shots" / noisy trajectories) of
 * `max_qubits` axes each, in one of two formalisms:
 *   ket : batch x 2^max_qubits complex128, axis a = bit a of the flat index
 *   dm  : batch x 4^max_qubits complex128, flat index = row << max_qubits | col,
 *         axis a = bit (a + max_qubits) of the row part and bit a of the column part
 * Axes are made live / freed dynamically (communication qubits come and go with every remote
 * gate).  A free axis is in |0> and costs no memory traffic: kernels enumerate live bits only.
 * Every shot has a 64-bit classical register; measurement outcomes are written there on the
 * device and classically-controlled corrections are *predicated* per shot, so a circuit runs
 * without any host round trip.
 *
 * enqueue-style calls only append to the handle's gate queue.  dqe_flush() (or any read)
 * cuts the queue into passes -- one pass = one read + one write of the live state, with every
 * queued op whose target bits fit the pass's shared-memory tile applied while the tile is
 * on chip -- and launches them on the handle's stream.
 *
 * Errors: every function returns 0 on success or a negative DQE_E* code; the message is
 * available from dqe_last_error().  No exception crosses the ABI.  CUDA errors are sticky.
 * Threading: a handle is bound to one device + one stream and is not thread-safe.
 */
#ifndef DQC_ENGINE_H
#define DQC_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dqe_engine dqe_t;

#define DQE_KET 0
#define DQE_DM  1

#define DQE_OK          0
#define DQE_EINVAL     -1   /* bad argument / axis state */
#define DQE_ECUDA      -2   /* CUDA runtime error (sticky) */
#define DQE_ENOMEM     -3
#define DQE_EUNSUPPORTED -4

/* ---- lifetime ------------------------------------------------------------------------- */

/* Replaces: ns.set_qstate_formalism(QFormalism.KET|DM) + the QState NetSquid creates per
 * entangled group (SURVEY.md Appendix A last row).  `ext_state` may be NULL (the engine
 * allocates batch * 2^max_qubits (ket) or batch * 4^max_qubits (dm) complex128) or a device
 * pointer to a caller-owned buffer of that size (e.g. a torch tensor's data_ptr()).
 * `stream` is a cudaStream_t cast to void* (NULL = a new non-blocking stream owned by the
 * handle).  State starts as |0..0> with no live axis.
 * device < 0 creates a PLAN-ONLY handle: no CUDA call is made, enqueue / flush run the queue,
 * peephole and pass planner and update the statistics, every read fails with DQE_EUNSUPPORTED
 * (used to test the planner on a machine without a GPU; it computes nothing). */
int dqe_create(dqe_t** out, int formalism, int max_qubits, int64_t batch, int device,
               void* ext_state, void* stream);
int dqe_destroy(dqe_t* h);
/* Back to |0..0> with no live axis, creg = 0, draw index = 0, queue dropped: start the next batch
 * of shots on the same buffers (the reference instead rebuilds the network per shot,
 * docs/source/getting_started.rst:683-770).  The global shot id of shot 0 advances by
 * batch * shot_stride (option "shot_stride", default 1; set it to the number of ranks that share one shot
 * numbering), so consecutive batches draw fresh random numbers; call dqe_seed to replay a batch. */
int dqe_reinit(dqe_t* h);
const char* dqe_last_error(const dqe_t* h);   /* h may be NULL: error of a failed dqe_create */
const char* dqe_version(void);

/* Philox4x32-10 stream: key = seed, counter = (global shot id, draw index).  `shot_offset`
 * is the global id of this handle's shot 0 so sampling is invariant under shot sharding.
 * Resets the draw index to 0. */
int dqe_seed(dqe_t* h, uint64_t seed, uint64_t shot_offset);

/* ---- register bookkeeping (INSTR_INIT / discard; dqc_control.py:473,
 *      compiler_preprocessing.py:46-53, quantum_processors.py:414-422) ------------------- */
int dqe_alloc_axis(dqe_t* h, int axis);   /* make a free axis live in |0> */
int dqe_free_axis(dqe_t* h, int axis);    /* dm: partial trace; ket: unrecorded sampled measurement */
int dqe_reset(dqe_t* h, int axis);        /* INSTR_INIT on an occupied position */
int dqe_live_mask(dqe_t* h, uint64_t* mask);
/* Declare which axes hold qubits after the amplitudes were moved by something the gate queue does not see (a
 * peer exchange, dqe_set_state through dqe_state_ptr): the caller asserts that every axis outside `mask` is in |0>,
 * i.e. that the halves of the state with such a bit set are exactly zero (the free-axis invariant, section
 * "state layout" above).  Flushes first. */
int dqe_force_live_mask(dqe_t* h, uint64_t mask);

/* ---- unitaries (qubitapi.operate via QuantumProgram.apply, dqc_control.py:191-196;
 *      operator matrices qlib/gates.py:29-220).  Matrices need not be unitary
 *      (make_state_gen_op, gates.py:45-48). ------------------------------------------------- */
int dqe_enqueue_1q(dqe_t* h, int axis, const double* m /*2x2 complex*/);
int dqe_enqueue_2q(dqe_t* h, int a, int b, const double* m /*4x4 complex, a = MSB*/);
int dqe_enqueue_ctrl_1q(dqe_t* h, int c, int t, const double* m /*2x2*/);  /* Operator.ctrl */
int dqe_enqueue_diag(dqe_t* h, int k, const int* axes, const double* d /*2^k complex*/);
/* Dense k-qubit gate, k <= 4 (first axis = most significant index): k = 3, 4 run as ONE real (2 * 2^k)-square
 * GEMM against the tile on the FP64 tensor cores (mma.sync m8n8k4.f64, SASS DMMA) -- the only tensor-core op of
 * the engine: every other gate is <= 4 x 4 per amplitude group and bound by memory, not by math.  The planner
 * also builds such blocks itself from runs of queued ket gates that stay inside <= 4 tile bits (option
 * "kblock", default 1).  Replaces qubitapi.operate(qubits, operator) for k-qubit operators
 * (INSTR_*_UNITARY of qlib/gates.py:107-115 generalised; the reference's own programs stop at k = 2). */
int dqe_enqueue_kq(dqe_t* h, int k, const int* axes, const double* m /*2^k x 2^k complex*/);

/* Ops enqueued until dqe_set_predicate(h, -1) act only on shots whose creg bit is 1
 * (classically controlled X / Z corrections, dqc_control.py:347-348, 387-390, 497-503). */
int dqe_set_predicate(dqe_t* h, int creg_bit);
int dqe_enqueue_cond_1q(dqe_t* h, int axis, const double* m, int creg_bit);

/* ---- channels (hardware/noise_models.py:33-131, 226-253, 295-307, 343-366;
 *      NetSquid DepolarNoiseModel wired at quantum_processors.py:868-877).
 *      dm: applied analytically in one pass; ket: one branch sampled per shot. ------------- */
int dqe_enqueue_pauli_channel(dqe_t* h, int axis, const double* p /*pI,pX,pY,pZ*/);
/* The k-qubit form of SURVEY 8b (k = 1, 2): p[4^k], index = 4 * (Pauli on axes[0]) + (Pauli on axes[1]) in the order
 * I, X, Y, Z -- NetSquid's qubitapi.apply_pauli_noise on each qubit is the product case, NDimDepolarNoiseModel
 * (hardware/noise_models.py:295-366) the uniform one; any correlated two-qubit Pauli noise fits.  dm: ONE dense
 * 16 x 16 superoperator block on (row a, row b, col a, col b); ket: one Pauli pair drawn per shot. */
int dqe_enqueue_pauli_channel_k(dqe_t* h, int k, const int* axes, const double* p /*4^k*/);
int dqe_enqueue_depolarize(dqe_t* h, int k, const int* axes, double p);  /* k = 1, 2 */
int dqe_enqueue_kraus_1q(dqe_t* h, int axis, int nk, const double* K /*nk x 2x2 complex*/);

/* ---- 2-qubit state injection (BlackBoxEntanglingQsourceConnection,
 *      hardware/connections.py:145-150; physical_layer.py:451-479).  `state` is a ket
 *      (4 complex) or a density matrix (16 complex, is_dm = 1); a = MSB.  Both axes must be
 *      free.  ket formalism + mixed state: an eigen-component is sampled per shot. --------- */
int dqe_inject_2q(dqe_t* h, int a, int b, const double* state, int is_dm);
/* ket formalism, mixed source state: per shot, component i (a ket of 4 complex) is injected
 * with probability probs[i]; ncomp <= 4.  The decomposition (eigenvectors of the 4x4 state,
 * e.g. werner_state(F), qlib/states.py:63-84) is the caller's so that it is reproducible. */
int dqe_inject_2q_mixture(dqe_t* h, int a, int b, int ncomp, const double* probs,
                          const double* kets /*ncomp x 4 complex*/);

/* ---- measurement (INSTR_MEASURE, dqc_control.py:298,366,446-459,783-787; physical
 *      instruction quantum_processors.py:414-422: noise BEFORE, optional discard).
 *      Z basis.  creg_bit < 0: outcome not recorded.  forced = 0/1 overrides sampling
 *      (-1 = sample).  flip_prob = MeasurementNoiseModel X-flip probability
 *      (noise_models.py:295-307). ---------------------------------------------------------- */
int dqe_measure(dqe_t* h, int axis, int creg_bit, int forced, int discard, double flip_prob);

/* ---- per-shot classical machine: the simulated timeline of every shot --------------------------
 * The reference runs one trajectory per ns.sim_run(): a classically controlled X / Z correction exists
 * (135 us + its gate noise) only when the outcome was 1 (dqc_control.py:347-348, 387-390, 497-503),
 * _cat_disentangle_end executes the pending program only then (:382-403), the link layer re-sends its
 * handshake on a 600 us grid (physical_layer.py:269-335), and every later memory-noise exponent
 * p = 1 - exp(-dt * 1e-9 * rate) (noise_models.py:248-253; NetSquid DepolarNoiseModel wired at
 * quantum_processors.py:868-877) inherits the difference.  A batch of shots therefore has one timeline
 * per outcome history.  Each shot owns dqe_n_time_regs() double registers next to its creg word;
 * dqe_classical_op appends one instruction (d = dst, a, b registers -- or creg bits for the B* ops):
 *    1 CONST d<-imm        2 ADDI d<-a+imm       3 ADD d<-a+b+imm      4 SUB d<-a-b+imm
 *    5 RSUBI d<-imm-a      6 MAX d<-max(a,b+imm) 7 MAXI d<-max(a,imm)  8 SEL d<-(creg bit cbit ? a+imm : b)
 *    9 CEILQ d<-max(0, ceil(a/imm - 1e-12))      10 FMAI d<-a*imm+b    11 QEXP d<-1-exp(-max(0,a)*imm)
 *   12 ADDIF d<-a+(creg bit cbit ? imm : 0)
 *   16 BAND  17 BANDN (a & ~b)  18 BOR  19 BNOT  20 BCOPY      (creg bit d <- f(bit a, bit b))
 * Instructions run per shot in queue order relative to each other and AFTER every measurement queued
 * before them, but the engine may run them EARLIER than quantum ops queued between that measurement and
 * them (it hoists them to just behind the measurement so that the quantum ops still fuse).  Contract: a
 * register / bit must not be overwritten while an op queued earlier since the last measurement (or
 * dqe_classical_fence) still reads it, and the result of a QEXP is read by channels only
 * (dqe_enqueue_dyn_channel), never by another classical instruction: the engine evaluates the QEXPs of
 * a span after its other instructions, spread over the threads of the CTA.  Registers are undefined until
 * written; dqe_reinit keeps them.
 * dqe_enqueue_dyn_channel: one-qubit channel whose strength q is the shot's register `reg`:
 *   kind 0 depolarising (1-q) rho + q I/2, 1 dephasing (off-diagonals x (1-q)), 2 amplitude damping
 *   (gamma = q; dm only).  dm: analytic; ket: one branch sampled per shot. */
int dqe_n_time_regs(void);
int dqe_classical_op(dqe_t* h, int opcode, int dst, int a, int b, int cbit, double imm);
/* d <- 1 - exp(-max(0, a - b + offset) * rate): the strength of an idle-time channel straight from the two
 * times it lies between (a / b = 255 stand for 0).  Like QEXP, its result is read by channels only. */
int dqe_classical_qexp2(dqe_t* h, int dst, int a, int b, double offset, double rate);
int dqe_classical_fence(dqe_t* h);
/* Per-shot random numbers for the classical machine, drawn from the shot's Philox stream (next draw index, like a
 * measurement): kind 0: creg bit dst <- (u < p); kind 1: register dst <- min(kmax, 1 + floor(ln u / ln(1 - p))) =
 * attempts until the first success of probability p.  The probabilistic physical layer needs them: which Bell
 * state the midpoint BSM heralded and how many emit attempts it took (physical_layer.py:685-753,
 * connections.py:399-499) differ from shot to shot and shift each shot's simulated clock. */
int dqe_classical_rand(dqe_t* h, int kind, int dst, double p, double kmax);
int dqe_enqueue_dyn_channel(dqe_t* h, int axis, int kind, int reg);
int dqe_read_time_reg(dqe_t* h, int reg, double* out /*batch*/);
/* Every measurement outcome of every shot, in draw order (the reference's mid-simulation result
 * collector, util/helper.py:226-263, sees them one signal at a time): enable with
 * dqe_set_option(h, "outcome_log", max_draws); rows [first, first + count) of the [draw][batch] int8 log
 * (-1 = not measured).  dqe_draw_count = random draws consumed since dqe_seed / dqe_reinit. */
int dqe_read_outcomes(dqe_t* h, int64_t first, int64_t count, int8_t* out);
int64_t dqe_draw_count(dqe_t* h);

/* ---- execution + read-out ------------------------------------------------------------------ */
int dqe_flush(dqe_t* h);                 /* plan + launch everything queued (asynchronous) */
/* Run the SAME program on the next batch of shots without walking it again.  With the option plan_cache = 1 a flush
 * that executed a whole program on a fresh register (dqe_reinit ... dqe_flush) keeps its planned passes; after the next
 * dqe_reinit, dqe_replay re-uploads the micro-program, launches those passes for the new block of global shot ids and
 * leaves the handle (live axes, draw counter) as the original walk left it.  The reference runs one subprocess per
 * shot and rebuilds everything each time (docs/source/getting_started.rst:683-770); a shot batch re-plans nothing.
 * Errors: no cached plan, queue not empty, register not fresh.  Any dqe_set_option drops the cached plan. */
int dqe_replay(dqe_t* h);
int dqe_sync(dqe_t* h);                  /* flush + wait for the stream */
int dqe_read_creg(dqe_t* h, int bit, int8_t* out /*batch*/);
int dqe_read_creg_words(dqe_t* h, uint64_t* out /*batch*/);
int dqe_histogram(dqe_t* h, int k, const int* cbits, int64_t* out /*2^k, first bit = MSB*/);
/* qapi.reduced_dm(qubits) (util/helper.py:209-214; examples/run_experiment.py:74-75):
 * list order = big-endian.  out: 2^k x 2^k complex, row-major. */
int dqe_reduced_dm(dqe_t* h, int k, const int* axes, int64_t shot, double* out);
/* qapi.fidelity(qubits, ket, squared=True) (util/helper.py:151) for every shot. */
int dqe_fidelity_ket(dqe_t* h, int k, const int* axes, const double* ket, double* out /*batch*/);
/* Raw state of one shot in the fixed layout (2^max_qubits or 4^max_qubits complex). */
int dqe_get_state(dqe_t* h, int64_t shot, double* out);
/* Overwrite one shot's raw state and declare which axes are live (test / restart hook). */
int dqe_set_state(dqe_t* h, int64_t shot, const double* in, uint64_t live_mask);
/* Device pointer + size of the whole state buffer (wrap zero-copy as a torch tensor). */
int dqe_state_ptr(dqe_t* h, void** dev, int64_t* nbytes);

/* ---- peer-memory path for state vectors sharded over GPUs (one handle per rank) ------
 * A gate that mixes a GLOBAL axis (a bit of the rank id) pairs every local amplitude with the
 * amplitude at the same local index on the partner rank.  dqe_peer_1q fuses the gate with the
 * transfer: this rank updates BOTH ranks' amplitudes for one half of the local index range
 * (`half` = 0 or 1; the partner takes the other half), reading and writing the partner's shard
 * directly over NVLink through an IPC-mapped pointer -- no staging buffer, no separate exchange, and
 * race-free in place because the two ranks touch disjoint index ranges.  The caller synchronises the
 * two ranks before and after (both shards quiescent / both halves done).
 *   dqe_ipc_export : 64-byte cudaIpcMemHandle_t of this handle's state buffer (engine-owned only)
 *   dqe_ipc_open   : map a partner's handle, returns a device pointer valid in this process
 *   dqe_peer_1q    : my_bit = this rank's value of the global axis; m = 2x2 complex; ctrl_axis = a
 *                    LOCAL axis that must be 1 (-1: none).  Requires every local axis live. */
int dqe_ipc_export(dqe_t* h, void* handle64);
int dqe_ipc_open(dqe_t* h, const void* handle64, void** peer_state);
int dqe_ipc_close(dqe_t* h, void* peer_state);
int dqe_peer_1q(dqe_t* h, void* peer_state, int my_bit, int half, const double* m, int ctrl_axis);
/* Swap the global axis with LOCAL axis `local_axis` (the data movement of a global<->local qubit swap):
 * element (global bit 0, local bit 1) <-> (global bit 1, local bit 0) at equal remaining bits, done in
 * place over peer memory, this rank handling one half of the pairs.  No staging buffer, both NVLink
 * directions busy at once. */
int dqe_peer_swap(dqe_t* h, void* peer_state, int my_bit, int half, int local_axis);
/* Multi-axis exchange, synchronised on the device (kernel k_peer_exchange): the `s` rank bits `rank_bits`
 * (ascending) trade places with the local axes `local_axes` -- new[R = x, L = y] = old[R = y, L = x] -- in ONE
 * launch that talks to all 2^s - 1 partners over peer memory: (1 - 2^-s) of the shard crosses NVLink once, where s
 * pairwise swaps move s / 2 of it.  No host barrier on either side: the ranks meet through epoch counters in a flag
 * block behind each shard (ready / done per partner), so the call only ENQUEUES and the host goes on planning.
 * The whole buffer moves, free axes included (their upper halves are zero): which axes are live afterwards is the
 * caller's business (dqe_force_live_mask).
 * Every rank of the group must call it with the same rank_bits / local_axes, in the same program order;
 * peer_states[p] = dqe_ipc_open of rank p's handle (entry `rank` ignored).  The NCCL all-to-all of
 * SURVEY.md 8e, as a kernel.  dqe_peer_stats: device time, count and bytes sent of the exchanges so far, and the
 * error word (1 = a partner never arrived within ~4 s; the state is then undefined). */
int dqe_peer_exchange(dqe_t* h, void* const* peer_states, int rank, int world, int s, const int* rank_bits,
                      const int* local_axes);
int dqe_peer_stats(dqe_t* h, double* ms, int64_t* count, int64_t* bytes_sent, int64_t* error);
/* Local part of a measurement on a sharded state vector (single-shot ket): out2 = (sum |amp|^2 over axis bit 0,
 * over axis bit 1) of this shard.  The caller all-reduces, draws, and projects with dqe_measure(forced = m). */
int dqe_probs(dqe_t* h, int axis, double* out2);
/* The same for a density matrix held as a ket on 2n bits (rows | columns) and sharded by row bits
 * (sharded.ShardedDM): out2 = local part of (Tr[(1 - Z_t)/2 ... ]) -- the sum of the DIAGONAL elements of rho on this
 * shard, split by the value of local bit `target` (-1: everything in out2[0]).  An element is diagonal when, for
 * every pair k, local bit a[k] equals local bit b[k] (fixed[k] < 0) or the rank-bit value fixed[k]. */
int dqe_probs_diag(dqe_t* h, int npairs, const int* a, const int* b, const int* fixed, int target, double* out2);

/* ---- tuning + instrumentation --------------------------------------------------------------- */
/* key: "tile_bits" (log2 amplitudes staged per CTA; default 11 = 32 KiB with 128 threads, or 10 = 16 KiB
 * with 64 threads for registers of <= 14 bits, where the <= 16 tiles of a shot form one thread-block
 * cluster), "threads" (per CTA: 64, 128, 256), "cluster" (0: measurements always cut a pass), "min_run_bits" (log2 of the shortest contiguous run a tile may be built
 * from, default 5), "fuse" (0: one op per pass), "merge" (0: no host-side folding of axis-local
 * ops), "group" (0: no register-resident qubit-pair sweeps), "tma" (0: per-thread loads instead
 * of cp.async.bulk), "persistent" (1: persistent CTAs with a double-buffered tile ring instead of
 * one tile per CTA), "ctas_per_sm" (cap for the persistent grid), "pin_mask" (axes that passes always
 * enumerate, live or free: pinning the lowest axes keeps the low bits of every tile contiguous in HBM
 * when those axes hold qubits that come and go, e.g. communication qubits), "cxperm" (0: X / CX inside a
 * qubit-pair group is arithmetic instead of a permutation of the sweep's load / store addresses), "factor"
 * (0: composed axis-local superoperators stay dense 4x4 stand-alone sweeps instead of running as their
 * Pauli-channel / U rho U^dagger factors inside group sweeps), "tmap" (0: a tile is fetched run by run with
 * cp.async.bulk instead of one cp.async.bulk.tensor box per direction), "prefetch" (tiles ahead whose box is
 * prefetched into L2 at CTA start; -1 = one wave of resident CTAs, 0 = off), "shot_stride" (see dqe_reinit),
 * "outcome_log" (see dqe_read_outcomes), "chain" (default 0; 1: consecutive passes that keep one tile per shot in
 * the same layout run as ONE launch -- the tile stays in shared memory, only the micro-program is re-staged per link,
 * a whole Grover-5 step reads and writes HBM once; bit-identical results, measured SLOWER on B200: the
 * long-lived CTAs of an SM drift apart and lose their shared instruction fetch), "chain_waves" (chunks of so many
 * waves of resident CTAs per chained launch, default 1; 0 = one grid), "chain_max" (longest chain, 0 = no limit),
 * "catfuse" (dm, small registers: 3 (default) = additionally fold WHOLE cat-comm remote gates into two ops on the data
 * pair, no comm qubit ever goes live; 2 = fold inject -> interact -> measure-and-discard
 * AND measure-out runs of a remote gate into one fused measurement each at flush -- the first comm qubit of the pair never
 * goes live --, 1 = the injecting runs only, 0 = execute every op as queued; software/dqc_control.py:316-403, 508-613),
 * "smem_pad" (A/B only: extra bytes of dynamic shared memory per CTA, i.e. fewer resident CTAs per SM),
 * "phase_tables" (ket, default 0: leftover one-bit phases share diagonal tables at flush; measured neutral),
 * "pass_ops" (micro-op budget of a pass, <= 48), "plan_cache" (see dqe_replay).  Setting "tile_bits" or "threads" also
 * switches off the per-pass choice of tile / CTA size that density matrices of <= 7 qubits get by default
 * (<= 12 live bits: 2^11 tiles on 128 threads, wider: 2^10 tiles on 64). */
int dqe_set_option(dqe_t* h, const char* key, int64_t value);
/* Counters since creation or dqe_reset_stats: [0] passes (links of a pass chain count one each)
 * [1] micro-ops executed in passes  [2] all kernel launches (a chain is one)  [3] amplitude-updates
 * (elements read+written once per pass)  [4] queued API ops  [5] measure-finish launches
 * [6] bytes of micro-ops + constants copied host->device  [7] ops folded by the peephole. */
int dqe_get_stats(dqe_t* h, int64_t* out /*8*/);
int dqe_reset_stats(dqe_t* h);
/* When enabled, every tile pass is bracketed by CUDA events on the handle's stream;
 * dqe_pass_time_ms returns the accumulated device time and pass count since enabling. */
int dqe_time_passes(dqe_t* h, int enable);
int dqe_pass_time_ms(dqe_t* h, double* total_ms, int64_t* n_passes, int64_t* amp_updates);

#ifdef __cplusplus
}
#endif
#endif /* DQC_ENGINE_H */
