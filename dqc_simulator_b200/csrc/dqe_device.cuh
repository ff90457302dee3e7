// dqe_device.cuh -- device side of the B200 quantum-state engine (sm_100a).
//
// One kernel does the heavy lifting: k_tile_pass.  A *pass* reads a tile of 2^t complex128
// amplitudes of one shot into shared memory (the lowest live bits of the register -- long
// contiguous runs in HBM -- plus the few high bits the queued ops mix), runs a micro-program
// of gate / channel / measurement ops on the tile while it is on chip, and writes it back:
// one HBM read + one HBM write per pass, however many ops were queued.
//
// State layout (see include/dqc_engine.h): ket = 2^P amplitudes per shot, P = max_qubits;
// dm = 4^max_qubits, P = 2*max_qubits, row bits above column bits, so a density matrix is
// handled as a "ket" on 2n bits: U rho U^dagger is U on the row bit and conj(U) on the column
// bit, and a channel is a 4x4 superoperator on (row bit, column bit) -- one pass over rho.
#pragma once
#include <cooperative_groups.h>
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

namespace dqe {

enum OpType : int32_t {
  OP_MAT1 = 1,        // 2x2 on one tile bit, optional control mask
  OP_MAT2 = 2,        // 4x4 on two tile bits (ket 2q gate, dm one-sided 2q, dm 1q superoperator)
  OP_DIAG = 3,        // phase table over <= 4 bits (tile or outer)
  OP_PAULI1 = 4,      // dm 1q Pauli channel, closed form on (row, col) bits
  OP_DEPOL2 = 5,      // dm 2q depolarising, closed form on (ra, rb, ca, cb)
  OP_PROBS = 6,       // partial outcome probabilities of one axis -> partials[]
  OP_COLLAPSE = 7,    // project on recorded outcome, renormalise, optional move to |0>
  OP_TRACE = 8,       // dm partial trace of one axis (result left in the |0><0| block)
  OP_INJECT = 9,      // write a 2-qubit state onto two fresh axes
  OP_PAULI_SAMPLED = 10,  // ket: apply a Pauli string drawn per shot from Philox
  OP_MAT1X2 = 11,     // dm controlled/one-qubit gate, row side + column side in one sweep
  OP_XPAT = 14,       // dm 1q channel with the sparse 'X pattern' (diagonal gate composed with Pauli noise):
                      // e00,e11 mix among themselves and e01,e10 among themselves -- 8 complex coefficients
  OP_DYN1 = 15,       // dm 1q channel whose strength q is read per shot from classical register `aux`
                      // (flags = DynKind): idle-time noise of a shot that follows its own simulated timeline
  OP_KBLOCK = 17,     // dense 2^k x 2^k complex block on k tile bits (k = 3, 4), b[0] = most significant bit of the
                      // matrix index: a real (2 * 2^k)-square GEMM against the tile on the FP64 tensor cores
                      // (mma.sync m8n8k4.f64 -> SASS DMMA); the host fuses runs of gates that stay inside k bits
  OP_LADDER = 18,     // ket: controlled-phase star  phase(i) = prod_j w_j^(b_q b_j)  (centre bit q = b[0], up to
                      // kMaxLadderLeaves leaf bits j in the payload bytes, w_j in the constant pool): the whole
                      // controlled-phase ladder of one QFT stage costs about two complex multiplies per amplitude,
                      // however many controls it has.  flags = tile-local leaves | outer leaves << 8
  OP_SWAP = 20,       // ket: exchange two tile bits (a SWAP gate is a permutation: two loads, two stores, no arithmetic)
  OP_KGROUP = 19,     // ket block sweep: every thread holds the 16 amplitudes spanned by FOUR tile bits (b[0..3],
                      // ascending, all >= 3 so that a warp's accesses stay conflict-free) in registers and runs the
                      // following `aux` sub-ops on them -- one-qubit matrices on those bits (controls anywhere),
                      // diagonal tables and controlled-phase stars on any bits -- in ONE shared-memory round trip:
                      // the H / phase-ladder alternation of a QFT stage costs one sweep per four stages.
                      // Sub-ops keep their type (OP_MAT1, OP_DIAG, OP_LADDER) and carry block ranks:
                      //   MAT1: b[4] = rank of the target, lctrl = thread-side control mask, lctrl2 = element mask
                      //   DIAG: b[4 + i] = rank + 1 of table bit i when it is a block bit, else 0
                      //   LADDER: b[1] = rank + 1 of the centre or 0; leaves ordered thread | element (rank) | outer,
                      //           flags = nt | ne << 8 | no << 16
  OP_CLASSICAL = 16,  // per-shot classical instructions (ClInstr x flags, packed behind the 16-byte header):
                      // run redundantly by every thread of the CTA (identical values: no barrier needed)
  OP_MEASURE = 13,    // fused measurement (cluster launch): tile partials -> DSMEM sum over the shot's CTAs ->
                      // Philox draw -> outcome + renormalisation kept on chip; aux = axis bit, flags = creg bit + 1,
                      // lctrl = forced + 1, p[0] = flip probability, octrl = draw index
  // dm group sweep: every thread holds the 16 elements (ra, rb, ca, cb) of one qubit PAIR (A, B) in
  // registers and runs the following `aux` sub-ops on them -- one shared-memory round trip for a
  // whole run of gates / channels on those two qubits.
  OP_GROUP = 12,
  G_PAULI1 = 101,     // flags bit0 = role of the target axis (0 = A, 1 = B)
  G_MAT1X2 = 103,     // U on the row side / conj(U) on the column side; flags bit1/bit2: row/col control
                      // is the OTHER axis of the group (else octrl/octrl2 on the tile base)
  G_DEPOL2 = 104,
  G_COLLAPSE = 107,
  G_TRACE = 108,
  G_INJECT = 109,
  G_XPAT = 111,
  G_DYN1 = 115,       // flags bit0 = role, bits 1-2 = DynKind, aux = classical register holding q
};

// Per-shot classical register machine (include/dqc_engine.h: dqe_classical_op).  It carries the simulated
// timelines of batched shots: a time that depends on a measurement outcome is a double in the shot's
// register file, the executor's clock arithmetic becomes these instructions, and idle-time noise channels
// read their strength from a register.  Opcodes mirror dqc_simulator_b200/timeline.py.
enum ClOp : uint8_t {
  CL_CONST = 1,   // d <- imm
  CL_ADDI = 2,    // d <- a + imm
  CL_ADD = 3,     // d <- a + b + imm
  CL_SUB = 4,     // d <- a - b + imm
  CL_RSUBI = 5,   // d <- imm - a
  CL_MAX = 6,     // d <- max(a, b + imm)
  CL_MAXI = 7,    // d <- max(a, imm)
  CL_SEL = 8,     // d <- creg bit `cbit` ? a + imm : b
  CL_CEILQ = 9,   // d <- max(0, ceil(a / imm - 1e-12))   (handshake re-sends, physical_layer.py:269-335)
  CL_FMAI = 10,   // d <- a * imm + b
  CL_QEXP = 11,   // d <- 1 - exp(-max(0, a) * imm)       (noise_models.py:248-253)
  CL_ADDIF = 12,  // d <- a + (creg bit `cbit` ? imm : 0)
  CL_QEXP2 = 13,  // d <- 1 - exp(-max(0, a - b + imm) * imm2): idle time and strength in one instruction; a / b = 255
                  // stand for 0, imm2 is a second constant-pool index carried in the (cbit, pad) bytes
  CL_GEOM = 14,   // d <- min(imm2, 1 + floor(ln(u) / imm)), u = the shot's Philox uniform of draw (a | b << 8), imm = ln(1 - p):
                  // number of attempts of a heralded entanglement link until the first success (geometric law)
  CL_BAND = 16,   // creg bit d <- bit a & bit b
  CL_BANDN = 17,  // d <- a & ~b
  CL_BOR = 18,
  CL_BNOT = 19,
  CL_BCOPY = 20,
  CL_RANDBIT = 21,  // creg bit d <- (u < imm), u = the shot's Philox uniform of draw (a | b << 8)
};
enum DynKind : int32_t { DYN_DEPOL = 0, DYN_DEPHASE = 1, DYN_AMPDAMP = 2, DYN_ZFLIP = 3 };   // ZFLIP: Z with probability q
struct ClInstr {   // 8 bytes; ten ride in one MicroOp
  uint8_t op, dst, a, b;
  int8_t cbit;
  uint8_t pad;
  uint16_t imm;    // index (in doubles) into the pass constant pool
};
static_assert(sizeof(ClInstr) == 8, "ClInstr is packed into MicroOp payloads");
constexpr int kTimeRegs = 128;        // doubles per shot
constexpr int kClPerOp = 10;          // ClInstr per OP_CLASSICAL micro-op
constexpr uint32_t kMeasRunsClassical = 1u << 30;   // OP_MEASURE.lctrl2: the classical block that follows runs inside the measurement

constexpr int kLadderLeafOff = 56;   // leaf bit positions of an OP_LADDER live in bytes [56, 96) of the MicroOp
constexpr int kMaxLadderLeaves = 40;
constexpr int kOuter = 64;       // MicroOp.b[] >= kOuter: bit lies outside the tile (physical pos = b - 64)
constexpr int kMaxOpsPerPass = 48;
// double2 constants staged in shared memory per pass.  The pool lives in DYNAMIC shared memory behind the tile and a
// pass asks for what it uses (PassParams::nconsts), so the cap costs nothing: an ordinary dm pass stages ~1 KiB (eight
// 16 KiB-tile CTAs per SM), one that carries fused-measurement tables ~5 KiB
constexpr int kMaxConstsKet = 512;
constexpr int kMaxConstsDm = 1024;
__host__ __device__ constexpr int max_consts(int form) { return form == 0 ? kMaxConstsKet : kMaxConstsDm; }
constexpr int kThreads = 256;          // maximum threads per CTA (launch bound); actual count is blockDim.x

struct __align__(16) MicroOp {
  // dispatch header: the first 16 bytes are everything the interpreter needs to decide what to run next,
  // fetched with ONE shared-memory load one op ahead of its use
  int32_t type;
  int32_t pred_bit;     // creg bit that must be 1 for the shot, -1 = unconditional
  int32_t flags;
  int32_t c_off;        // offset of this op's constants in the pass constant pool (double2 units);
                        // OP_GROUP header: number of sub-ops (= aux), so that skipping a group needs the header only
  uint32_t lctrl;       // control mask, tile-local bits
  uint32_t nb;          // number of entries of b[] in use
  int32_t aux;          // PROBS: physical bit of the axis; sampled ops: draw index; OP_GROUP: number of sub-ops
  uint32_t lctrl2;      // MAT1X2: control mask of the column side
  uint64_t octrl;       // control mask, physical bits outside the tile (tested on the tile base)
  uint8_t b[8];         // bit positions (tile-local, or kOuter + physical)
  uint64_t octrl2;
  double p[4];
};
static_assert(sizeof(MicroOp) == 96, "MicroOp is staged by 16-byte bulk copies");
__device__ __forceinline__ int4 op_header(const MicroOp* op) { return *reinterpret_cast<const int4*>(op); }

struct Seg {  // bits [src, src+len) of a dense index go to physical bits [dst, dst+len)
  uint8_t src, len, dst, pad;
};

struct MeasRec {
  double scale;
  int32_t outcome;
  int32_t pad;
};

struct PassParams {
  double2* state;
  const uint64_t* creg;     // per-shot classical word, read at CTA start
  double* partials;
  const MeasRec* rec;
  // classical state is double-buffered: a pass that changes it (fused measurement, classical instructions)
  // reads the `in` copy and the CTA holding a shot's tile 0 writes the `out` copy when it is done, so that
  // CTAs of the same shot that start later still see the state from before the pass
  uint64_t* creg_out;
  MeasRec* rec_out;
  const double* treg;       // kTimeRegs doubles per shot (null: the pass touches no register)
  double* treg_out;         // non-null: registers are written back
  int8_t* olog;             // outcome log [draw][batch] or null
  int64_t batch;
  const MicroOp* ops;
  const double2* consts;
  // pass chain (nlinks > 1): consecutive passes over the SAME tile layout with one tile per shot run as ONE launch --
  // the tile stays in shared memory, only the micro-program changes: link k = {first micro-op, micro-ops, first
  // constant, constants} relative to ops_all / consts_all; ops / nops / consts / nconsts above describe link 0
  const uint4* links;
  const MicroOp* ops_all;
  const double2* consts_all;
  uint64_t seed, shot_offset;
  int32_t nops, t, P, nmax;
  int32_t formalism, outer_bits, n_tile_segs, n_outer_segs;
  int32_t store_back, nconsts;
  int32_t use_tma, nthreads;
  int32_t nbuf, has_measure;
  int32_t nfree_cols, nlinks;
  int32_t cap_consts, cap_ops;        // constants / micro-ops the dynamic shared-memory area holds (>= nconsts / nops;
                                      // chains: the largest link): constants, then micro-ops, then the nregs classical
                                      // registers of the shot
  uint64_t tile0;                     // first tile of this launch (chains run the batch as chunks of whole waves)
  int32_t uses_treg, writes_creg, writes_treg, nregs;  // planner: what the pass does to the classical state;
                                                       // nregs (even) = registers [0, nregs) are in use
  uint64_t ntiles;
  int8_t ax_col[24], ax_row[24];   // dm: tile-local position of the column / row bit of each axis, -1 = outside
  // dm diagonal enumeration (probabilities), host-precomputed.  Free variable k = k-th axis (ascending)
  // whose column bit lies in the tile.
  int8_t ax_ord[24];       // free-variable index of an axis, -1 = column bit outside the tile
  uint32_t dg_mask[16];    // tile-index contribution of free variable k: column bit | row bit (if in the tile)
  int8_t dg_req[16];       // physical row bit (outside the tile) whose base value variable k must equal, -1 = none
  int8_t dg_fix_src[16], dg_fix_dst[16];   // column bit outside, row bit inside: tile bit dst = base bit src
  int32_t dg_nfix;
  uint32_t dg_req_mask;    // free variables with a dg_req entry
  uint32_t dg_eq_mask;     // axes with both bits outside: base bit a must equal base bit a + nmax
  // TMA tensor map of the pass (tm_rank != 0): the tile is ONE box -- dimension 0 = the contiguous low run,
  // one dimension per higher tile segment, plus a box-1 "base" dimension of stride 2^tm_shift amplitudes
  // whose coordinate carries the shot and the outer bits.  One instruction loads / stores the whole tile.
  alignas(64) CUtensorMap tmap;
  int32_t tm_rank, tm_base_dim, tm_shift, prefetch_dist;
  uint32_t run_hi[128];   // TMA: (physical offset of run r) >> c0, for the 2^(t-c0) <= 128 runs of a tile
  Seg tile_segs[16];
  Seg outer_segs[24];
};

// ------------------------------------------------------------------------------------------
// complex helpers
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 acc) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
  return acc;
}
// conj(a) * b and acc + conj(a) * b without materialising the conjugate (operand negation is free)
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ double2 cfmac(double2 a, double2 b, double2 acc) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
  return acc;
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ double2 conj2(double2 a) { return make_double2(a.x, -a.y); }

// insert a zero bit at position j
__device__ __forceinline__ uint32_t ins0(uint32_t p, int j) {
  const uint32_t lo = (1u << j) - 1u;
  return ((p & ~lo) << 1) | (p & lo);
}

__device__ __forceinline__ uint32_t ins0_4(uint32_t p, const int* sorted) {
  p = ins0(p, sorted[0]); p = ins0(p, sorted[1]); p = ins0(p, sorted[2]); p = ins0(p, sorted[3]);
  return p;
}

template <typename SegT>
__device__ __forceinline__ uint64_t deposit(uint64_t x, const SegT* segs, int n) {
  uint64_t r = 0;
  // the first segments with constant indices: the descriptors are then immediate param-space operands
#pragma unroll
  for (int s = 0; s < 3; ++s)
    if (s < n) r |= ((x >> segs[s].src) & ((1ull << segs[s].len) - 1ull)) << segs[s].dst;
  for (int s = 3; s < n; ++s) {
    r |= ((x >> segs[s].src) & ((1ull << segs[s].len) - 1ull)) << segs[s].dst;
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  counter = (shot_lo, shot_hi, draw_lo, draw_hi), key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ double uniform53(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ void shot_uniforms(uint64_t seed, uint64_t shot, uint64_t draw, double& u, double& u2) {
  uint32_t c[4] = {(uint32_t)shot, (uint32_t)(shot >> 32), (uint32_t)draw, (uint32_t)(draw >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  u = uniform53(c[0], c[1]);
  u2 = uniform53(c[2], c[3]);
}

// ------------------------------------------------------------------------------------------
// micro-ops on the shared-memory tile
__device__ __forceinline__ void apply_mat1(double2* tile, uint32_t T, int j, uint32_t lc,
                                           double2 m00, double2 m01, double2 m10, double2 m11) {
  for (uint32_t p = threadIdx.x; p < (T >> 1); p += blockDim.x) {
    const uint32_t i0 = ins0(p, j);
    if ((i0 & lc) != lc) continue;
    const uint32_t i1 = i0 | (1u << j);
    const double2 a = tile[i0], b = tile[i1];
    tile[i0] = cfma(m01, b, cmul(m00, a));
    tile[i1] = cfma(m11, b, cmul(m10, a));
  }
}

__device__ __forceinline__ void op_mat1(double2* tile, uint32_t T, const MicroOp& op, const double2* cs) {
  apply_mat1(tile, T, op.b[0], op.lctrl, cs[0], cs[1], cs[2], cs[3]);
}

// dm: U on the row bit (control mask lctrl) then conj(U) on the column bit (control lctrl2),
// on the 4 elements (r, c) in registers -- one shared-memory round trip for U rho U^dagger.
__device__ __forceinline__ void op_mat1x2(double2* tile, uint32_t T, const MicroOp& op, const double2* cs,
                                          bool row_on, bool col_on) {
  const int jr = op.b[0], jc = op.b[1];
  const int jlo = jr < jc ? jr : jc, jhi = jr < jc ? jc : jr;
  const double2 m00 = cs[0], m01 = cs[1], m10 = cs[2], m11 = cs[3];
  const double2 n00 = conj2(m00), n01 = conj2(m01), n10 = conj2(m10), n11 = conj2(m11);
  const uint32_t lr = op.lctrl, lc = op.lctrl2;
  for (uint32_t p = threadIdx.x; p < (T >> 2); p += blockDim.x) {
    const uint32_t i00 = ins0(ins0(p, jlo), jhi);
    const bool do_r = row_on && ((i00 & lr) == lr);   // row-side control lives on row bits
    const bool do_c = col_on && ((i00 & lc) == lc);
    if (!do_r && !do_c) continue;
    const uint32_t i01 = i00 | (1u << jc), i10 = i00 | (1u << jr), i11 = i01 | (1u << jr);
    double2 e00 = tile[i00], e01 = tile[i01], e10 = tile[i10], e11 = tile[i11];
    if (do_r) {  // mixes r for fixed c
      double2 a = e00, b = e10;
      e00 = cfma(m01, b, cmul(m00, a)); e10 = cfma(m11, b, cmul(m10, a));
      a = e01; b = e11;
      e01 = cfma(m01, b, cmul(m00, a)); e11 = cfma(m11, b, cmul(m10, a));
    }
    if (do_c) {  // mixes c for fixed r
      double2 a = e00, b = e01;
      e00 = cfma(n01, b, cmul(n00, a)); e01 = cfma(n11, b, cmul(n10, a));
      a = e10; b = e11;
      e10 = cfma(n01, b, cmul(n00, a)); e11 = cfma(n11, b, cmul(n10, a));
    }
    tile[i00] = e00; tile[i01] = e01; tile[i10] = e10; tile[i11] = e11;
  }
}

__device__ __forceinline__ void op_mat2(double2* tile, uint32_t T, const MicroOp& op, const double2* cs) {
  const int jh = op.b[0], jl = op.b[1];
  const int jlo = jh < jl ? jh : jl, jhi = jh < jl ? jl : jh;
  const uint32_t lc = op.lctrl;
  const double2* __restrict__ m = cs;
  for (uint32_t p = threadIdx.x; p < (T >> 2); p += blockDim.x) {
    const uint32_t i0 = ins0(ins0(p, jlo), jhi);
    if ((i0 & lc) != lc) continue;
    const uint32_t i1 = i0 | (1u << jl), i2 = i0 | (1u << jh), i3 = i1 | (1u << jh);
    const double2 v0 = tile[i0], v1 = tile[i1], v2 = tile[i2], v3 = tile[i3];
    tile[i0] = cfma(m[3], v3, cfma(m[2], v2, cfma(m[1], v1, cmul(m[0], v0))));
    tile[i1] = cfma(m[7], v3, cfma(m[6], v2, cfma(m[5], v1, cmul(m[4], v0))));
    tile[i2] = cfma(m[11], v3, cfma(m[10], v2, cfma(m[9], v1, cmul(m[8], v0))));
    tile[i3] = cfma(m[15], v3, cfma(m[14], v2, cfma(m[13], v1, cmul(m[12], v0))));
  }
}

__device__ __forceinline__ void op_swap(double2* tile, uint32_t T, const MicroOp& op) {
  const int ja = op.b[0], jb = op.b[1];
  const int jlo = ja < jb ? ja : jb, jhi = ja < jb ? jb : ja;
  const uint32_t lc = op.lctrl;
  for (uint32_t p = threadIdx.x; p < (T >> 2); p += blockDim.x) {
    const uint32_t i0 = ins0(ins0(p, jlo), jhi);
    if ((i0 & lc) != lc) continue;
    const uint32_t i1 = i0 | (1u << ja), i2 = i0 | (1u << jb);
    const double2 x = tile[i1], y = tile[i2];
    tile[i1] = y; tile[i2] = x;
  }
}

__device__ __forceinline__ void op_diag(double2* tile, uint32_t T, const MicroOp& op, const double2* cs,
                                        uint64_t base) {
  const int nb = op.nb;
  uint32_t fixed = 0;
  int lb[4], ls[4], nl = 0;
  for (int s = 0; s < nb; ++s) {
    const int sh = nb - 1 - s;
    if (op.b[s] >= kOuter) fixed |= (uint32_t)((base >> (op.b[s] - kOuter)) & 1ull) << sh;
    else { lb[nl] = op.b[s]; ls[nl] = sh; ++nl; }
  }
  if (nl == 0) {
    const double2 d = cs[fixed];
    for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) tile[i] = cmul(tile[i], d);
    return;
  }
  for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {
    uint32_t idx = fixed;
    for (int s = 0; s < nl; ++s) idx |= ((i >> lb[s]) & 1u) << ls[s];
    tile[i] = cmul(tile[i], cs[idx]);
  }
}


// a run of consecutive diagonal ops applied in ONE sweep: per amplitude one table look-up and one
// complex multiply per op, one shared-memory round trip for the whole run (the controlled-phase
// ladders of a QFT collapse into a few such sweeps).  Kept out of line so that its registers do not
// weigh on the group-sweep path.
struct DiagPrep {
  int32_t c_off, nl, on;
  uint32_t fixed;
  uint32_t m0, m1, m2, m3;     // single-bit masks of the tile-local bits (0 when unused)
  uint32_t s0, s1, s2, s3;     // their weights in the table index
  uint32_t h0, h1, h2, h3;     // block sweeps: element-index masks of the table bits that are block bits
};
constexpr int kMaxDiagRun = 16;
struct LadderPrep {
  double2 G;         // product of the phases of the outer leaves whose base bit is 1
  uint32_t qmask;    // tile-local mask of the centre bit (0: centre outside the tile, `on` holds its base bit)
  int32_t on;
};
__device__ __forceinline__ const uint8_t* ladder_leaves(const MicroOp& op) {
  return reinterpret_cast<const uint8_t*>(&op) + kLadderLeafOff;
}

__device__ __forceinline__ void op_diag_run(double2* tile, uint32_t T, const MicroOp* ops, int count,
                                         const double2* s_consts, uint64_t base, uint64_t cw, DiagPrep* prep,
                                         LadderPrep* lprep) {
  // ladders: the factor of the OUTER leaves is the same for the whole tile -- eight lanes per op multiply their
  // share of the selected phases and combine through shuffles (uniform trip count: every lane shuffles)
  {
    const int grp = threadIdx.x >> 3, part = threadIdx.x & 7, ngrp = blockDim.x >> 3;
    for (int k0 = 0; k0 < count; k0 += ngrp) {
      const int k = k0 + grp;
      const bool act = k < count && ops[k < count ? k : 0].type == OP_LADDER;
      double2 g = make_double2(1.0, 0.0);
      if (act) {
        const MicroOp& op = ops[k];
        const int nl = op.flags & 0xff, no = (op.flags >> 8) & 0xff;
        const uint8_t* lf = ladder_leaves(op);
        for (int j = part; j < no; j += 8)
          if ((base >> (lf[nl + j] - kOuter)) & 1ull) g = cmul(g, s_consts[op.c_off + nl + j]);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        double2 h;
        h.x = __shfl_xor_sync(0xffffffffu, g.x, o);
        h.y = __shfl_xor_sync(0xffffffffu, g.y, o);
        g = cmul(g, h);
      }
      if (act && part == 0) {
        const MicroOp& op = ops[k];
        LadderPrep L;
        L.G = g;
        L.on = !(op.pred_bit >= 0 && !((cw >> op.pred_bit) & 1ull));
        if (op.b[0] >= kOuter) { L.qmask = 0u; L.on = L.on && ((base >> (op.b[0] - kOuter)) & 1ull); }
        else L.qmask = 1u << op.b[0];
        lprep[k] = L;
      }
    }
  }
  if ((int)threadIdx.x < count && ops[threadIdx.x].type == OP_DIAG) {
    const MicroOp& op = ops[threadIdx.x];
    DiagPrep d;
    d.c_off = op.c_off; d.nl = 0; d.fixed = 0;
    d.on = !(op.pred_bit >= 0 && !((cw >> op.pred_bit) & 1ull));
    uint32_t mk[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0};
    const int nb = op.nb;
    for (int s = 0; s < nb; ++s) {
      const uint32_t w = 1u << (nb - 1 - s);
      if (op.b[s] >= kOuter) { if ((base >> (op.b[s] - kOuter)) & 1ull) d.fixed |= w; }
      else { mk[d.nl] = 1u << op.b[s]; sh[d.nl] = w; ++d.nl; }
    }
    d.m0 = mk[0]; d.m1 = mk[1]; d.m2 = mk[2]; d.m3 = mk[3];
    d.s0 = sh[0]; d.s1 = sh[1]; d.s2 = sh[2]; d.s3 = sh[3];
    prep[threadIdx.x] = d;
  }
  __syncthreads();
  if (T == 16u * blockDim.x) {
    // register-blocked form: a thread owns the 16 amplitudes i = tid + e * blockDim.  For every table
    // the index splits into a part fixed by the thread id (tile bits below log2(blockDim), computed once
    // per op) and a part that depends on e only (bits above; identical for all threads, compile-time e).
    const int lb = 31 - __clz(blockDim.x);
    double2 v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = tile[threadIdx.x + (uint32_t)e * blockDim.x];
    for (int k = 0; k < count; ++k) {
      if (ops[k].type == OP_LADDER) {
        const LadderPrep L = lprep[k];
        if (!L.on) continue;
        const MicroOp& op = ops[k];
        const int nl = op.flags & 0xff;
        const uint8_t* lf = ladder_leaves(op);
        const double2* w = s_consts + op.c_off;
        const uint32_t tid = threadIdx.x;
        // thread part: centre / leaves on thread-id bits (tile-local leaves are sorted ascending)
        if ((L.qmask & (blockDim.x - 1u)) && !(tid & L.qmask)) continue;
        double2 sc = L.G;
        int j = 0;
        for (; j < nl && lf[j] < lb; ++j)
          if ((tid >> lf[j]) & 1u) sc = cmul(sc, w[j]);
        // element part: e = tile bits [lb, lb + 4)
        const uint32_t qe = L.qmask >> lb;   // 0 when the centre is a thread bit or outside the tile
        for (; j < nl; ++j) {
          const uint32_t eb = 1u << (lf[j] - lb);
          const double2 wj = w[j];
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if ((e & eb) && (e & qe) == qe) v[e] = cmul(v[e], wj);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if ((e & qe) == qe) v[e] = cmul(v[e], sc);
        continue;
      }
      const DiagPrep d = prep[k];
      if (!d.on) continue;
      const uint32_t tid = threadIdx.x;
      const uint32_t lowmask = blockDim.x - 1u;
      // thread-fixed contribution: masks that live in the low (thread id) bits
      const uint32_t tf = d.fixed + (((tid & d.m0 & lowmask)) ? d.s0 : 0u) + (((tid & d.m1 & lowmask)) ? d.s1 : 0u) +
                          (((tid & d.m2 & lowmask)) ? d.s2 : 0u) + (((tid & d.m3 & lowmask)) ? d.s3 : 0u);
      // e-dependent masks (shifted down to bit positions of e)
      const uint32_t h0 = d.m0 >> lb, h1 = d.m1 >> lb, h2 = d.m2 >> lb, h3 = d.m3 >> lb;
      const double2* tb = s_consts + d.c_off + tf;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const uint32_t idx = ((e & h0) ? d.s0 : 0u) + ((e & h1) ? d.s1 : 0u) + ((e & h2) ? d.s2 : 0u) + ((e & h3) ? d.s3 : 0u);
        v[e] = cmul(v[e], tb[idx]);
      }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) tile[threadIdx.x + (uint32_t)e * blockDim.x] = v[e];
    return;
  }
  for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {
    double2 v = tile[i];
    for (int k = 0; k < count; ++k) {
      if (ops[k].type == OP_LADDER) {
        const LadderPrep& L = lprep[k];
        if (!L.on || (L.qmask && !(i & L.qmask))) continue;
        const MicroOp& op = ops[k];
        const int nl = op.flags & 0xff;
        const uint8_t* lf = ladder_leaves(op);
        double2 f = L.G;
        for (int j = 0; j < nl; ++j)
          if ((i >> lf[j]) & 1u) f = cmul(f, s_consts[op.c_off + j]);
        v = cmul(v, f);
        continue;
      }
      const DiagPrep& d = prep[k];
      if (!d.on) continue;
      const uint32_t idx = d.fixed + ((i & d.m0) ? d.s0 : 0u) + ((i & d.m1) ? d.s1 : 0u) +
                           ((i & d.m2) ? d.s2 : 0u) + ((i & d.m3) ? d.s3 : 0u);
      v = cmul(v, s_consts[d.c_off + idx]);
    }
    tile[i] = v;
  }
}


// ------------------------------------------------------------------------------------------
// ket block sweep (OP_KGROUP).  v[e]: bit k of e <-> block bit s4[k].
template <int R>
__device__ __forceinline__ void kg_mat1(double2 (&v)[16], const double2* __restrict__ cs, uint32_t em) {
  const double2 m00 = cs[0], m01 = cs[1], m10 = cs[2], m11 = cs[3];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    if (e & (1 << R)) continue;
    if ((e & em) == em) {     // element-side controls (uniform)
      double2& x0 = v[e];
      double2& x1 = v[e | (1 << R)];
      const double2 p = cmul(m10, x0);
      x0 = cfma(m01, x1, cmul(m00, x0));
      x1 = cfma(m11, x1, p);
    }
  }
}

__device__ __noinline__ void run_kgroup(double2* tile, uint32_t T, const MicroOp& g, const MicroOp* sub,
                                        const double2* s_consts, uint64_t base, uint64_t cw, DiagPrep* prep,
                                        LadderPrep* lprep, double2* letab) {
  const int nsub = g.aux;
  // ---- prep: the tile-uniform part of every sub-op
  {
    const int grp = threadIdx.x >> 3, part = threadIdx.x & 7, ngrp = blockDim.x >> 3;
    for (int k0 = 0; k0 < nsub; k0 += ngrp) {
      const int k = k0 + grp;
      const bool act = k < nsub && sub[k < nsub ? k : 0].type == OP_LADDER;
      double2 gg = make_double2(1.0, 0.0);
      if (act) {
        const MicroOp& op = sub[k];
        const int nt = op.flags & 0xff, ne = (op.flags >> 8) & 0xff, no = (op.flags >> 16) & 0xff;
        const uint8_t* lf = ladder_leaves(op);
        for (int j = part; j < no; j += 8)
          if ((base >> (lf[nt + ne + j] - kOuter)) & 1ull) gg = cmul(gg, s_consts[op.c_off + nt + ne + j]);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        double2 hh;
        hh.x = __shfl_xor_sync(0xffffffffu, gg.x, o);
        hh.y = __shfl_xor_sync(0xffffffffu, gg.y, o);
        gg = cmul(gg, hh);
      }
      if (act && part == 0) {
        const MicroOp& op = sub[k];
        LadderPrep L;
        L.G = gg;
        L.on = !(op.pred_bit >= 0 && !((cw >> op.pred_bit) & 1ull));
        L.qmask = 0u;
        if (op.b[0] >= kOuter) L.on = L.on && ((base >> (op.b[0] - kOuter)) & 1ull);
        else if (op.b[1] == 0) L.qmask = 1u << op.b[0];     // centre on a thread-side tile bit
        lprep[k] = L;
      }
      if (act) {
        // stars with two or more leaves on register-resident bits: the phase of element e from those leaves is the
        // same for every thread -- a 16-entry table per star, built here once per CTA (two entries per lane), turns
        // "one predicated multiplication per element per leaf" into one per element
        const MicroOp& op = sub[k];
        const int nt = op.flags & 0xff, ne = (op.flags >> 8) & 0xff;
        if (ne >= 2) {
          const uint8_t* lf = ladder_leaves(op);
          for (int e = part; e < 16; e += 8) {
            double2 ph = make_double2(1.0, 0.0);
            for (int j = 0; j < ne; ++j)
              if ((e >> lf[nt + j]) & 1) ph = cmul(ph, s_consts[op.c_off + nt + j]);
            letab[k * 16 + e] = ph;
          }
        }
      }
    }
  }
  if ((int)threadIdx.x < nsub && sub[threadIdx.x].type == OP_DIAG) {
    const MicroOp& op = sub[threadIdx.x];
    DiagPrep d;
    d.c_off = op.c_off; d.nl = 0; d.fixed = 0;
    d.on = !(op.pred_bit >= 0 && !((cw >> op.pred_bit) & 1ull));
    uint32_t mk[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0}, hk[4] = {0, 0, 0, 0};
    const int nb = op.nb;
    for (int s = 0; s < nb; ++s) {
      const uint32_t w = 1u << (nb - 1 - s);
      if (op.b[s] >= kOuter) { if ((base >> (op.b[s] - kOuter)) & 1ull) d.fixed |= w; }
      else {
        if (op.b[4 + s]) hk[d.nl] = 1u << (op.b[4 + s] - 1); else mk[d.nl] = 1u << op.b[s];
        sh[d.nl] = w; ++d.nl;
      }
    }
    d.m0 = mk[0]; d.m1 = mk[1]; d.m2 = mk[2]; d.m3 = mk[3];
    d.s0 = sh[0]; d.s1 = sh[1]; d.s2 = sh[2]; d.s3 = sh[3];
    d.h0 = hk[0]; d.h1 = hk[1]; d.h2 = hk[2]; d.h3 = hk[3];
    prep[threadIdx.x] = d;
  }
  __syncthreads();
  // ---- sweep
  const int s4[4] = {g.b[0], g.b[1], g.b[2], g.b[3]};
  const uint32_t b0 = __shfl_sync(0xffffffffu, 16u << g.b[0], 0), b1 = __shfl_sync(0xffffffffu, 16u << g.b[1], 0);
  const uint32_t b2 = __shfl_sync(0xffffffffu, 16u << g.b[2], 0), b3 = __shfl_sync(0xffffffffu, 16u << g.b[3], 0);
#define DQE_KOFF(e) ((((e) & 1) ? b0 : 0u) | (((e) & 2) ? b1 : 0u) | (((e) & 4) ? b2 : 0u) | (((e) & 8) ? b3 : 0u))
  for (uint32_t q = threadIdx.x; q < (T >> 4); q += blockDim.x) {
    const uint32_t i0 = ins0_4(q, s4);
    char* __restrict__ tpb = reinterpret_cast<char*>(tile + i0);
    double2 v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = *reinterpret_cast<double2*>(tpb + DQE_KOFF(e));
    for (int s = 0; s < nsub; ++s) {
      const MicroOp& op = sub[s];
      const int4 hdr = op_header(&op);
      if (hdr.x == OP_MAT1) {
        if (hdr.y >= 0 && !((cw >> hdr.y) & 1ull)) continue;
        if ((base & op.octrl) != op.octrl) continue;
        if ((i0 & op.lctrl) != op.lctrl) continue;
        const double2* cs = s_consts + hdr.w;
        const uint32_t em = op.lctrl2;
        switch (op.b[4]) {
          case 0: kg_mat1<0>(v, cs, em); break;
          case 1: kg_mat1<1>(v, cs, em); break;
          case 2: kg_mat1<2>(v, cs, em); break;
          default: kg_mat1<3>(v, cs, em); break;
        }
      } else if (hdr.x == OP_DIAG) {
        const DiagPrep d = prep[s];
        if (!d.on) continue;
        const uint32_t tf = d.fixed + ((i0 & d.m0) ? d.s0 : 0u) + ((i0 & d.m1) ? d.s1 : 0u) + ((i0 & d.m2) ? d.s2 : 0u) +
                            ((i0 & d.m3) ? d.s3 : 0u);
        const double2* tb = s_consts + d.c_off + tf;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const uint32_t idx = ((e & d.h0) ? d.s0 : 0u) + ((e & d.h1) ? d.s1 : 0u) + ((e & d.h2) ? d.s2 : 0u) + ((e & d.h3) ? d.s3 : 0u);
          v[e] = cmul(v[e], tb[idx]);
        }
      } else {   // OP_LADDER
        const LadderPrep L = lprep[s];
        if (!L.on) continue;
        if (L.qmask && !(i0 & L.qmask)) continue;
        const int nt = hdr.z & 0xff, ne = (hdr.z >> 8) & 0xff;
        const uint8_t* lf = ladder_leaves(op);
        const double2* w = s_consts + hdr.w;
        // the product over the thread-index leaves as two independent chains (half the dependent latency)
        double2 sc = L.G, sc2 = make_double2(1.0, 0.0);
        int j = 0;
        for (; j + 1 < nt; j += 2) {
          if ((i0 >> lf[j]) & 1u) sc = cmul(sc, w[j]);
          if ((i0 >> lf[j + 1]) & 1u) sc2 = cmul(sc2, w[j + 1]);
        }
        if (j < nt && ((i0 >> lf[j]) & 1u)) sc = cmul(sc, w[j]);
        sc = cmul(sc, sc2);
        const uint32_t qe = op.b[1] ? (1u << (op.b[1] - 1)) : 0u;
        if (ne >= 2) {       // element-phase table of the star (prep phase): one multiplication per element
          const double2* et = letab + s * 16;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if ((e & qe) == qe) v[e] = cmul(v[e], cmul(sc, et[e]));
        } else {
          if (ne == 1) {
            const uint32_t eb = 1u << lf[nt];
            const double2 wj = w[nt];
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if ((e & eb) && (e & qe) == qe) v[e] = cmul(v[e], wj);
          }
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if ((e & qe) == qe) v[e] = cmul(v[e], sc);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) *reinterpret_cast<double2*>(tpb + DQE_KOFF(e)) = v[e];
  }
#undef DQE_KOFF
}

// ------------------------------------------------------------------------------------------
// Dense k-qubit block (k = 3, 4) on the FP64 tensor cores.
// The complex map y = U x on the 2^k amplitudes of a column is the real map
//     [Re y]   [Re U  -Im U] [Re x]
//     [Im y] = [Im U   Re U] [Im x]
// i.e. a (2K x 2K) real matrix (K = 2^k) against a (2K x ncols) real matrix whose columns are the tile's
// T / K amplitude groups.  One warp takes 8 columns at a time: 2K/8 row tiles x 2K/4 k-steps of
// mma.sync.aligned.m8n8k4 (fp64).  The A fragments (the matrix) are loaded once per op and stay in registers.
// Fragment layout of m8n8k4.f64: A[8x4]: lane holds A[lane / 4][lane % 4]; B[4x8]: lane holds B[lane % 4][lane / 4];
// C / D[8x8]: lane holds rows lane / 4, columns 2 * (lane % 4) + {0, 1}.
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int KQ>
__device__ __noinline__ void op_kblock(double2* tile, uint32_t T, const MicroOp& op, const double2* cs) {
  constexpr int K = 1 << KQ, R = 2 * K;        // complex dimension, real dimension
  constexpr int MT = R / 8, KS = R / 4;        // row tiles, k-steps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  // element e of a column (KQ-bit index, b[0] = MSB) sits at tile offset eoff(e)
  int sb[4];                                   // the block's bit positions, ascending (for the zero insertion)
#pragma unroll
  for (int i = 0; i < KQ; ++i) sb[i] = op.b[i];
#pragma unroll
  for (int i = 0; i < KQ - 1; ++i)
#pragma unroll
    for (int j = 0; j < KQ - 1 - i; ++j)
      if (sb[j] > sb[j + 1]) { const int t = sb[j]; sb[j] = sb[j + 1]; sb[j + 1] = t; }
  auto eoff = [&](int e) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < KQ; ++i) o |= (uint32_t)((e >> (KQ - 1 - i)) & 1) << op.b[i];
    return o;
  };
  // A fragments: real matrix Ar[row][col], row / col < R; (row, col) -> U[row % K][col % K] with the 2 x 2 real
  // embedding by (row / K, col / K): (0,0) Re, (0,1) -Im, (1,0) Im, (1,1) Re
  double afrag[MT][KS];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int row = 8 * mt + lr, col = 4 * ks + lc;
      const double2 u = cs[(row % K) * K + (col % K)];
      const int qr = row / K, qc = col / K;
      afrag[mt][ks] = qr == qc ? u.x : (qr == 0 ? -u.y : u.y);
    }
  // this lane's B rows (one per k-step) and D rows (one per row tile), as byte offsets into a column
  uint32_t boff[KS], doff[MT];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int row = 4 * ks + lc;
    boff[ks] = eoff(row % K) * 16u + (row / K) * 8u;
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int row = 8 * mt + lr;
    doff[mt] = eoff(row % K) * 16u + (row / K) * 8u;
  }
  const uint32_t ncols = T >> KQ;
  char* tb = reinterpret_cast<char*>(tile);
  if (ncols < 8u) {
    // a register too small for one 8-column MMA tile (< 2^(k+3) amplitudes): plain FMAs, one column per thread
    for (uint32_t c = threadIdx.x; c < ncols; c += blockDim.x) {
      uint32_t i0 = c;
#pragma unroll
      for (int j = 0; j < KQ; ++j) i0 = ins0(i0, sb[j]);
      double2 x[K], y[K];
#pragma unroll
      for (int e = 0; e < K; ++e) x[e] = tile[i0 | eoff(e)];
#pragma unroll
      for (int r = 0; r < K; ++r) {
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int e = 0; e < K; ++e) acc = cfma(cs[r * K + e], x[e], acc);
        y[r] = acc;
      }
#pragma unroll
      for (int e = 0; e < K; ++e) tile[i0 | eoff(e)] = y[e];
    }
    return;
  }
  for (uint32_t c0 = warp * 8u; c0 < ncols; c0 += nwarps * 8u) {
    // column index -> tile index with zeros at the block's bits
    auto colbase = [&](uint32_t c) {
      uint32_t i = c;
#pragma unroll
      for (int j = 0; j < KQ; ++j) i = ins0(i, sb[j]);
      return i * 16u;
    };
    const uint32_t bcol = colbase(c0 + lr);                       // B: column lane / 4
    const uint32_t dcol0 = colbase(c0 + 2 * lc), dcol1 = colbase(c0 + 2 * lc + 1);   // D: columns 2 (lane % 4) + {0, 1}
    double bfrag[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) bfrag[ks] = *reinterpret_cast<const double*>(tb + bcol + boff[ks]);
    __syncwarp();   // every lane has read its part of the 8 columns before anyone overwrites them
    double d0[MT], d1[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { d0[mt] = 0.0; d1[mt] = 0.0; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) dmma_m8n8k4(d0[mt], d1[mt], afrag[mt][ks], bfrag[ks]);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      *reinterpret_cast<double*>(tb + dcol0 + doff[mt]) = d0[mt];
      *reinterpret_cast<double*>(tb + dcol1 + doff[mt]) = d1[mt];
    }
  }
}

// cs = {a, b, c, d, f, g, h, k}: e00' = a e00 + b e11, e11' = c e00 + d e11, e01' = f e01 + g e10, e10' = h e01 + k e10
__device__ __forceinline__ void op_xpat(double2* tile, uint32_t T, const MicroOp& op, const double2* cs) {
  const int jr = op.b[0], jc = op.b[1];
  const int jlo = jr < jc ? jr : jc, jhi = jr < jc ? jc : jr;
  const double2 a = cs[0], b = cs[1], c = cs[2], d = cs[3], f = cs[4], g = cs[5], h = cs[6], k = cs[7];
  for (uint32_t p = threadIdx.x; p < (T >> 2); p += blockDim.x) {
    const uint32_t i00 = ins0(ins0(p, jlo), jhi);
    const uint32_t i01 = i00 | (1u << jc), i10 = i00 | (1u << jr), i11 = i01 | (1u << jr);
    const double2 e00 = tile[i00], e01 = tile[i01], e10 = tile[i10], e11 = tile[i11];
    tile[i00] = cfma(b, e11, cmul(a, e00));
    tile[i11] = cfma(d, e11, cmul(c, e00));
    tile[i01] = cfma(g, e10, cmul(f, e01));
    tile[i10] = cfma(k, e10, cmul(h, e01));
  }
}

// ------------------------------------------------------------------------------------------
// per-shot classical machine
__device__ __forceinline__ const ClInstr* cl_payload(const MicroOp& op) {
  return reinterpret_cast<const ClInstr*>(reinterpret_cast<const char*>(&op) + 16);
}
// 1 - exp(-x) for x >= 0.  Idle-time noise lives at x ~ 1e-6 .. 1e-2, where nine Taylor terms are exact to the
// last bit of 1 - exp(-x) evaluated in double (what the reference's p = 1 - exp(-dt * 1e-9 * rate) is,
// noise_models.py:248-253) and cost a quarter of exp()'s latency on the one lane that computes it.
__device__ __forceinline__ double one_minus_exp_neg(double x) {
  if (x < 0.0009765625) {   // 2^-10: the terms beyond x^6 / 720 are below 2^-64 x -- five terms are exact
    double p = 1.0 / 720.0;
    p = fma(p, x, -1.0 / 120.0); p = fma(p, x, 1.0 / 24.0); p = fma(p, x, -1.0 / 6.0);
    p = fma(p, x, 0.5);
    return x - x * x * p;
  }
  if (x < 0.03125) {
    double p = -1.0 / 362880.0;
    p = fma(p, x, 1.0 / 40320.0); p = fma(p, x, -1.0 / 5040.0); p = fma(p, x, 1.0 / 720.0);
    p = fma(p, x, -1.0 / 120.0); p = fma(p, x, 1.0 / 24.0); p = fma(p, x, -1.0 / 6.0);
    p = fma(p, x, 0.5);
    return x - x * x * p;   // x - x^2/2 + x^3/6 - ...
  }
  return 1.0 - exp(-x);
}

// one instruction on the register file `R` (shared or global) and the classical word `cw`; c = its immediate
__device__ __forceinline__ void cl_exec(const ClInstr in, const double c, double* R, uint64_t& cw, const double* imm,
                                        uint64_t seed, uint64_t gshot) {
  switch (in.op) {   // (register operands are read inside the cases: a / b are creg bits for the B* ops)
    case CL_CONST: R[in.dst] = c; break;
    case CL_ADDI: R[in.dst] = R[in.a] + c; break;
    case CL_ADD: R[in.dst] = R[in.a] + R[in.b] + c; break;
    case CL_SUB: R[in.dst] = R[in.a] - R[in.b] + c; break;
    case CL_RSUBI: R[in.dst] = c - R[in.a]; break;
    case CL_MAX: R[in.dst] = fmax(R[in.a], R[in.b] + c); break;
    case CL_MAXI: R[in.dst] = fmax(R[in.a], c); break;
    case CL_SEL: R[in.dst] = ((cw >> in.cbit) & 1ull) ? R[in.a] + c : R[in.b]; break;
    case CL_ADDIF: R[in.dst] = R[in.a] + (((cw >> in.cbit) & 1ull) ? c : 0.0); break;
    case CL_CEILQ: R[in.dst] = fmax(0.0, ceil(R[in.a] / c - 1e-12)); break;
    case CL_FMAI: R[in.dst] = __dadd_rn(__dmul_rn(R[in.a], c), R[in.b]); break;   // like the host: no contraction
    case CL_QEXP: R[in.dst] = one_minus_exp_neg(fmax(0.0, R[in.a]) * c); break;
    case CL_QEXP2: {
      const double ra = in.a == 255 ? 0.0 : R[in.a], rb = in.b == 255 ? 0.0 : R[in.b];
      const double rate = imm[(uint32_t)(uint8_t)in.cbit | ((uint32_t)in.pad << 8)];
      R[in.dst] = one_minus_exp_neg(fmax(0.0, ra - rb + c) * rate);
      break;
    }
    case CL_GEOM: case CL_RANDBIT: {
      double u, u2;
      shot_uniforms(seed, gshot, (uint64_t)in.a | ((uint64_t)in.b << 8), u, u2);
      if (in.op == CL_RANDBIT) {
        cw = (cw & ~(1ull << in.dst)) | ((uint64_t)(u < c) << in.dst);
      } else {
        const double kmax = imm[(uint32_t)(uint8_t)in.cbit | ((uint32_t)in.pad << 8)];
        const double k = u > 0.0 ? 1.0 + floor(log(u) / c) : kmax;
        R[in.dst] = fmin(kmax, k);
      }
      break;
    }
    case CL_BAND: case CL_BANDN: case CL_BOR: case CL_BNOT: case CL_BCOPY: {
      const uint64_t x = (cw >> in.a) & 1ull, y = (cw >> in.b) & 1ull;
      const uint64_t v = in.op == CL_BAND ? (x & y) : in.op == CL_BANDN ? (x & (y ^ 1ull)) : in.op == CL_BOR ? (x | y)
                       : in.op == CL_BNOT ? (x ^ 1ull) : x;
      cw = (cw & ~(1ull << in.dst)) | (v << in.dst);
      break;
    }
    default: break;
  }
}
// a run of instructions in order on one thread; the fetch of instruction i + 1 (and of its immediate) is issued
// before instruction i executes, so the chain pays for the register-file round trips only
// `cnt_flags` = MicroOp.flags of an OP_CLASSICAL: instruction count | (leading serial instructions << 8).  The serial
// part (the clock chain) runs on every calling thread; the rest are exp() evaluations that feed channels only and are
// independent of each other: with WARP = true (the callers inside k_tile_pass, whole warps on one shot) lane i takes the
// i-th, so three idle-time strengths cost one exp() latency instead of three.
template <bool WARP>
__device__ __forceinline__ void cl_run(const ClInstr* ci, int cnt_flags, double* R, uint64_t& cw, const double* imm,
                                       uint64_t seed, uint64_t gshot) {
  const int cnt = cnt_flags & 0xff;
  const int ns = WARP ? min((cnt_flags >> 8) & 0xff, cnt) : cnt;
  if (cnt <= 0) return;
  if (ns > 0) {
    ClInstr nxt = ci[0];
    double cn = imm[nxt.imm];
    for (int i = 0; i < ns; ++i) {
      const ClInstr in = nxt;
      const double c = cn;
      if (i + 1 < ns) { nxt = ci[i + 1]; cn = imm[nxt.imm]; }
      cl_exec(in, c, R, cw, imm, seed, gshot);
    }
  }
  if (WARP && cnt > ns) {
    __syncwarp();
    for (int k = ns + (int)(threadIdx.x & 31u); k < cnt; k += 32) {
      const ClInstr in = ci[k];
      cl_exec(in, imm[in.imm], R, cw, imm, seed, gshot);
    }
    __syncwarp();
  }
}

// real coefficients of a dynamic one-qubit channel of strength q:
// e00' = a e00 + b e11, e11' = c e00 + d e11, e01' = f e01, e10' = f e10
struct DynCoef { double a, b, c, d, f; };
__device__ __forceinline__ DynCoef dyn_coef(int kind, double q) {
  DynCoef k;
  if (kind == DYN_DEPOL) { k.a = k.d = 1.0 - 0.5 * q; k.b = k.c = 0.5 * q; k.f = 1.0 - q; }
  else if (kind == DYN_DEPHASE) { k.a = k.d = 1.0; k.b = k.c = 0.0; k.f = 1.0 - q; }
  else if (kind == DYN_ZFLIP) { k.a = k.d = 1.0; k.b = k.c = 0.0; k.f = 1.0 - 2.0 * q; }
  else { k.a = 1.0; k.b = q; k.c = 0.0; k.d = 1.0 - q; k.f = sqrt(1.0 - q); }
  return k;
}

__device__ __forceinline__ void op_dyn1(double2* tile, uint32_t T, const MicroOp& op, const double* treg) {
  const int jr = op.b[0], jc = op.b[1];
  const int jlo = jr < jc ? jr : jc, jhi = jr < jc ? jc : jr;
  const DynCoef k = dyn_coef(op.flags, treg[op.aux]);
  for (uint32_t p = threadIdx.x; p < (T >> 2); p += blockDim.x) {
    const uint32_t i00 = ins0(ins0(p, jlo), jhi);
    const uint32_t i01 = i00 | (1u << jc), i10 = i00 | (1u << jr), i11 = i01 | (1u << jr);
    const double2 e00 = tile[i00], e01 = tile[i01], e10 = tile[i10], e11 = tile[i11];
    tile[i00] = make_double2(fma(k.a, e00.x, k.b * e11.x), fma(k.a, e00.y, k.b * e11.y));
    tile[i11] = make_double2(fma(k.d, e11.x, k.c * e00.x), fma(k.d, e11.y, k.c * e00.y));
    tile[i01] = make_double2(k.f * e01.x, k.f * e01.y);
    tile[i10] = make_double2(k.f * e10.x, k.f * e10.y);
  }
}

__device__ __forceinline__ void op_pauli1(double2* tile, uint32_t T, const MicroOp& op) {
  const int jr = op.b[0], jc = op.b[1];
  const int jlo = jr < jc ? jr : jc, jhi = jr < jc ? jc : jr;
  const double a = op.p[0], b = op.p[1], c = op.p[2], d = op.p[3];
  for (uint32_t p = threadIdx.x; p < (T >> 2); p += blockDim.x) {
    const uint32_t i00 = ins0(ins0(p, jlo), jhi);
    const uint32_t i01 = i00 | (1u << jc), i10 = i00 | (1u << jr), i11 = i01 | (1u << jr);
    const double2 e00 = tile[i00], e01 = tile[i01], e10 = tile[i10], e11 = tile[i11];
    tile[i00] = make_double2(fma(a, e00.x, b * e11.x), fma(a, e00.y, b * e11.y));
    tile[i11] = make_double2(fma(a, e11.x, b * e00.x), fma(a, e11.y, b * e00.y));
    tile[i01] = make_double2(fma(c, e01.x, d * e10.x), fma(c, e01.y, d * e10.y));
    tile[i10] = make_double2(fma(c, e10.x, d * e01.x), fma(c, e10.y, d * e01.y));
  }
}

__device__ __forceinline__ void sort4(const uint8_t* b, int* s) {
  s[0] = b[0]; s[1] = b[1]; s[2] = b[2]; s[3] = b[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3 - i; ++j)
      if (s[j] > s[j + 1]) { int t = s[j]; s[j] = s[j + 1]; s[j + 1] = t; }
}

// (1-p) rho + p * (I/4 (x) Tr_ab rho) on bits (ra, rb, ca, cb)
__device__ __forceinline__ void op_depol2(double2* tile, uint32_t T, const MicroOp& op) {
  int s[4];
  sort4(op.b, s);
  const uint32_t mra = 1u << op.b[0], mrb = 1u << op.b[1], mca = 1u << op.b[2], mcb = 1u << op.b[3];
  const double pp = op.p[0], keep = 1.0 - pp, mix = 0.25 * pp;
  for (uint32_t p = threadIdx.x; p < (T >> 4); p += blockDim.x) {
    const uint32_t i0 = ins0_4(p, s);
    const uint32_t d0 = i0, d1 = i0 | mrb | mcb, d2 = i0 | mra | mca, d3 = d1 | mra | mca;
    const double2 v0 = tile[d0], v1 = tile[d1], v2 = tile[d2], v3 = tile[d3];
    const double sx = mix * (v0.x + v1.x + v2.x + v3.x), sy = mix * (v0.y + v1.y + v2.y + v3.y);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t ro = ((r & 2) ? mra : 0u) | ((r & 1) ? mrb : 0u);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t idx = i0 | ro | ((c & 2) ? mca : 0u) | ((c & 1) ? mcb : 0u);
        double2 v = tile[idx];
        v.x *= keep; v.y *= keep;
        if (r == c) { v.x += sx; v.y += sy; }
        tile[idx] = v;
      }
    }
  }
}

// dm: partial outcome probabilities of axis `abit` over the diagonal elements of rho held by this tile.
// The enumeration (which tile elements are diagonal, which base bits must agree) is host-precomputed in
// PassParams; the tile-uniform part costs a few uniform instructions, the per-element part one OR per
// free variable.
__device__ __forceinline__ void dm_diag_partials(const double2* tile, const PassParams& pp, uint64_t base, int abit,
                                                 int op_ord, double& a0, double& a1) {
  const int n = pp.nmax, nfree = pp.nfree_cols;
  if ((((base >> n) ^ base) & (uint64_t)pp.dg_eq_mask) != 0ull) return;   // no diagonal element in this tile
  uint32_t ifix = 0, req_val = 0;
  for (int f = 0; f < pp.dg_nfix; ++f) ifix |= (uint32_t)((base >> pp.dg_fix_src[f]) & 1ull) << pp.dg_fix_dst[f];
  const uint32_t req_mask = pp.dg_req_mask;
  // constant loop bounds + constant indices: dg_req / dg_mask become immediate param-space operands
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if ((req_mask >> k) & 1u) req_val |= (uint32_t)((base >> pp.dg_req[k]) & 1ull) << k;
  for (int k = 8; k < nfree; ++k)
    if ((req_mask >> k) & 1u) req_val |= (uint32_t)((base >> pp.dg_req[k]) & 1ull) << k;
  const int kidx = op_ord;
  const uint32_t vfix = (uint32_t)((base >> abit) & 1ull);
  for (uint32_t d = threadIdx.x; d < (1u << nfree); d += blockDim.x) {
    if ((d ^ req_val) & req_mask) continue;
    uint32_t i = ifix;
#pragma unroll
    for (int k = 0; k < 8; ++k) i |= (0u - ((d >> k) & 1u)) & pp.dg_mask[k];   // unused masks are 0
    for (int k = 8; k < nfree; ++k) i |= ((d >> k) & 1u) ? pp.dg_mask[k] : 0u;
    const double w = tile[i].x;
    const uint32_t va = kidx >= 0 ? ((d >> kidx) & 1u) : vfix;
    if (va) a1 += w; else a0 += w;
  }
}

template <int FORM>
__device__ __forceinline__ void op_probs(const double2* tile, uint32_t T, const MicroOp& op, const PassParams& pp,
                                         uint64_t base, uint64_t blk, double* red) {
  double a0 = 0.0, a1 = 0.0;
  const int abit = op.aux;   // physical (column) bit of the measured axis
  if constexpr (FORM == 0) {
    // b[0] = tile-local position of the axis bit, or kOuter + physical when it is fixed for the tile
    const int jb = op.b[0];
    if (jb >= kOuter) {
      double a = 0.0;
      for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) { const double2 v = tile[i]; a += fma(v.x, v.x, v.y * v.y); }
      if ((base >> (jb - kOuter)) & 1ull) a1 = a; else a0 = a;
    } else {
      for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {
        const double2 v = tile[i];
        const double w = fma(v.x, v.x, v.y * v.y);
        if ((i >> jb) & 1u) a1 += w; else a0 += w;
      }
    }
  } else {
    // only the diagonal of rho matters: enumerate it directly
    dm_diag_partials(tile, pp, base, abit, (int)(int8_t)op.b[1], a0, a1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[2 * w] = a0; red[2 * w + 1] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s0 += red[2 * i]; s1 += red[2 * i + 1]; }
    pp.partials[2 * blk] = s0;
    pp.partials[2 * blk + 1] = s1;
  }
}


// shared by the fused-measurement forms: is the group at tile index i0 diagonal in every axis but axA / axY?
__device__ __forceinline__ bool catm_tile_diagonal(const PassParams& pp, uint64_t base, int axA, int axY) {
  const int n = pp.nmax;
  for (int a = 0; a < n; ++a) {   // axes with both bits outside the tile: one test for the whole tile
    if (a == axA || a == axY || pp.ax_col[a] >= 0 || pp.ax_row[a] >= 0) continue;
    if (((base >> a) ^ (base >> (a + n))) & 1ull) return false;
  }
  return true;
}
__device__ __forceinline__ bool catm_group_diagonal(const PassParams& pp, uint64_t base, uint32_t i0, int axA, int axY) {
  const int n = pp.nmax;
  bool ok = true;
  for (int a = 0; a < n; ++a) {
    const int cl = pp.ax_col[a], rl = pp.ax_row[a];
    if (a == axA || a == axY || (cl < 0 && rl < 0)) continue;
    const uint32_t cb = cl >= 0 ? (i0 >> cl) & 1u : (uint32_t)((base >> a) & 1ull);
    const uint32_t rb = rl >= 0 ? (i0 >> rl) & 1u : (uint32_t)((base >> (a + n)) & 1ull);
    ok = ok && cb == rb;
  }
  return ok;
}
// The same two tests from what the planner resolved per op (Planner::localise): lctrl2 = pairs | mixed << 8 (bit 31:
// too many, use the generic loops above), p[1..2] = (column, row) tile positions of the axes inside the tile,
// p[3] = (tile position, base bit) of the axes with one bit outside, octrl2 = axes that live in the base only.
struct CatDiag {
  uint32_t np, mixmask, mixwant;
  uint64_t pw0, pw1;
  bool generic, any;
};
__device__ __forceinline__ CatDiag catm_diag_prepare(const MicroOp& op, const PassParams& pp, uint64_t base) {
  CatDiag d;
  d.generic = (op.lctrl2 >> 31) != 0u;
  d.np = op.lctrl2 & 0xffu; d.mixmask = 0u; d.mixwant = 0u; d.pw0 = 0ull; d.pw1 = 0ull;
  if (d.generic) { d.any = catm_tile_diagonal(pp, base, op.b[4], op.b[5]); return d; }
  d.any = ((((base >> pp.nmax) ^ base) & op.octrl2) == 0ull);
  memcpy(&d.pw0, &op.p[1], 8); memcpy(&d.pw1, &op.p[2], 8);
  uint64_t mw;
  memcpy(&mw, &op.p[3], 8);
  const uint32_t nm = (op.lctrl2 >> 8) & 0xffu;
  for (uint32_t k = 0; k < nm; ++k) {
    const uint32_t tp = (uint32_t)(mw >> (16 * k)) & 0xffu, pb = (uint32_t)(mw >> (16 * k + 8)) & 0xffu;
    d.mixmask |= 1u << tp;
    d.mixwant |= (uint32_t)((base >> pb) & 1ull) << tp;
  }
  return d;
}
// tile index (zeros at the group's four bits) of the d-th diagonal group of the tile: bit k of d sets both bits of
// the k-th in-tile axis, the one-bit-outside axes follow the tile base -- a shot has 2^(other axes) such groups out of
// 4^(other axes), so the probability sums enumerate them directly and most warps skip the op
__device__ __forceinline__ uint32_t catm_diag_group(const CatDiag& d, uint32_t k_bits) {
  uint32_t i0 = d.mixwant;
  for (uint32_t k = 0; k < d.np; ++k) {
    const uint64_t w = k < 4u ? d.pw0 >> (16 * k) : d.pw1 >> (16 * (k - 4u));
    if ((k_bits >> k) & 1u) i0 |= (1u << ((uint32_t)w & 0xffu)) | (1u << ((uint32_t)(w >> 8) & 0xffu));
  }
  return i0;
}

// Fused inject -> interact -> measure-and-discard (host: fuse_cat_measure in dqe_engine.cu).  The op carries the bits
// of a data axis A and of a FRESH axis y (b[0..3] = tile positions of row A, row y, col A, col y; b[4], b[5] = the
// axes).  A thread owns the 16 elements (rA ry cA cy) of a group; on entry only the four with ry = cy = 0 are
// non-zero: in[j], j = rA * 2 + cA.  Constants: t_0[4], t_1[4], T_0[16][4], T_1[16][4]; b[7] = 1: the tables are real
// and packed as doubles.
//   probability of outcome m (flip noise already folded in) = sum over groups that are diagonal in every other
//   axis of Re(sum_j t_m[j] in[j]);   new elements = scale * sum_j T_m[e][j] in[j].
__device__ __noinline__ void catm_partials(const double2* tile, uint32_t T, const MicroOp& op, const PassParams& pp,
                                              const double2* __restrict__ cs, uint64_t base, double& a0, double& a1) {
  const CatDiag dg = catm_diag_prepare(op, pp, base);
  if (!dg.any) return;
  int s4[4];
  sort4(op.b, s4);
  const uint32_t mrA = 1u << op.b[0], mcA = 1u << op.b[2];
  const double2 t00 = cs[0], t01 = cs[1], t02 = cs[2], t03 = cs[3];
  const double2 t10 = cs[4], t11 = cs[5], t12 = cs[6], t13 = cs[7];
  const uint32_t ngrp = dg.generic ? (T >> 4) : (1u << dg.np);
  for (uint32_t q = threadIdx.x; q < ngrp; q += blockDim.x) {
    uint32_t i0;
    if (dg.generic) {
      i0 = ins0_4(q, s4);
      if (!catm_group_diagonal(pp, base, i0, op.b[4], op.b[5])) continue;
    } else {
      i0 = catm_diag_group(dg, q);
    }
    const double2 v0 = tile[i0], v1 = tile[i0 | mcA], v2 = tile[i0 | mrA], v3 = tile[i0 | mrA | mcA];
    a0 += fma(t00.x, v0.x, -t00.y * v0.y) + fma(t01.x, v1.x, -t01.y * v1.y) + fma(t02.x, v2.x, -t02.y * v2.y) +
          fma(t03.x, v3.x, -t03.y * v3.y);
    a1 += fma(t10.x, v0.x, -t10.y * v0.y) + fma(t11.x, v1.x, -t11.y * v1.y) + fma(t12.x, v2.x, -t12.y * v2.y) +
          fma(t13.x, v3.x, -t13.y * v3.y);
  }
}
template <bool REAL>
__device__ __noinline__ void catm_apply(double2* tile, uint32_t T, const MicroOp& op, const double2* cs, const MeasRec& r) {
  int s4[4];
  sort4(op.b, s4);
  const uint32_t mrA = 1u << op.b[0], mrY = 1u << op.b[1], mcA = 1u << op.b[2], mcY = 1u << op.b[3];
  // table of the drawn outcome: REAL: 64 packed doubles (T_m[e][j] at double index 4 e + j), else 64 complex entries
  const double2* __restrict__ Tm = cs + 8 + (REAL ? 32 : 64) * r.outcome;
  const double sc = r.scale;
  for (uint32_t q = threadIdx.x; q < (T >> 4); q += blockDim.x) {
    const uint32_t i0 = ins0_4(q, s4);
    const double2 v0 = cscale(tile[i0], sc), v1 = cscale(tile[i0 | mcA], sc), v2 = cscale(tile[i0 | mrA], sc),
                  v3 = cscale(tile[i0 | mrA | mcA], sc);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const uint32_t idx = i0 | ((e & 8) ? mrA : 0u) | ((e & 4) ? mrY : 0u) | ((e & 2) ? mcA : 0u) | ((e & 1) ? mcY : 0u);
      if constexpr (REAL) {
        const double2 c01 = Tm[2 * e], c23 = Tm[2 * e + 1];
        double2 o;
        o.x = fma(c23.y, v3.x, fma(c23.x, v2.x, fma(c01.y, v1.x, c01.x * v0.x)));
        o.y = fma(c23.y, v3.y, fma(c23.x, v2.y, fma(c01.y, v1.y, c01.x * v0.y)));
        tile[idx] = o;
      } else {
        tile[idx] = cfma(Tm[4 * e + 3], v3, cfma(Tm[4 * e + 2], v2, cfma(Tm[4 * e + 1], v1, cmul(Tm[4 * e], v0))));
      }
    }
  }
}

// Contraction form (b[6] = 1): x = an existing axis, measured and discarded behind a run of constant ops on (A, x).
// b[0..3] = tile positions of (row A, row x, col A, col x); constants t_0[16], t_1[16], T_0[4][16], T_1[4][16] (b[7] = 1: real tables, packed doubles)
// (element index i = rA * 8 + rx * 4 + cA * 2 + cx).  The four surviving elements land in the |0><0| block of x.
__device__ __noinline__ void catm2_partials(const double2* tile, uint32_t T, const MicroOp& op, const PassParams& pp,
                                               const double2* __restrict__ cs, uint64_t base, double& a0, double& a1) {
  const CatDiag dg = catm_diag_prepare(op, pp, base);
  if (!dg.any) return;
  int s4[4];
  sort4(op.b, s4);
  const uint32_t mrA = 1u << op.b[0], mrX = 1u << op.b[1], mcA = 1u << op.b[2], mcX = 1u << op.b[3];
  const uint32_t ngrp = dg.generic ? (T >> 4) : (1u << dg.np);
  for (uint32_t q = threadIdx.x; q < ngrp; q += blockDim.x) {
    uint32_t i0;
    if (dg.generic) {
      i0 = ins0_4(q, s4);
      if (!catm_group_diagonal(pp, base, i0, op.b[4], op.b[5])) continue;
    } else {
      i0 = catm_diag_group(dg, q);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const double2 v = tile[i0 | ((i & 8) ? mrA : 0u) | ((i & 4) ? mrX : 0u) | ((i & 2) ? mcA : 0u) | ((i & 1) ? mcX : 0u)];
      const double2 u0 = cs[i], u1 = cs[16 + i];
      a0 += fma(u0.x, v.x, -u0.y * v.y);
      a1 += fma(u1.x, v.x, -u1.y * v.y);
    }
  }
}
template <bool REAL>
__device__ __noinline__ void catm2_apply(double2* tile, uint32_t T, const MicroOp& op, const double2* __restrict__ cs,
                                         const MeasRec& r) {
  int s4[4];
  sort4(op.b, s4);
  const uint32_t mrA = 1u << op.b[0], mrX = 1u << op.b[1], mcA = 1u << op.b[2], mcX = 1u << op.b[3];
  const double2* __restrict__ Tm = cs + 32 + (REAL ? 32 : 64) * r.outcome;   // REAL: packed doubles, index 16 o + i
  const double sc = r.scale;
  const double2 zero = make_double2(0.0, 0.0);
  for (uint32_t q = threadIdx.x; q < (T >> 4); q += blockDim.x) {
    const uint32_t i0 = ins0_4(q, s4);
    double2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t idx = i0 | ((i & 8) ? mrA : 0u) | ((i & 4) ? mrX : 0u) | ((i & 2) ? mcA : 0u) | ((i & 1) ? mcX : 0u);
      v[i] = tile[idx];
      tile[idx] = zero;
    }
#pragma unroll 1
    for (int o = 0; o < 4; ++o) {     // rolled: one row of T_m at a time keeps the live registers at v[] + a few
      double2 acc = zero;
      if constexpr (REAL) {
        const double2* __restrict__ row = Tm + 8 * o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double2 c = row[i];
          acc.x = fma(c.x, v[2 * i].x, acc.x); acc.y = fma(c.x, v[2 * i].y, acc.y);
          acc.x = fma(c.y, v[2 * i + 1].x, acc.x); acc.y = fma(c.y, v[2 * i + 1].y, acc.y);
        }
      } else {
        const double2* __restrict__ row = Tm + 16 * o;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc = cfma(row[i], v[i], acc);
      }
      tile[i0 | ((o & 2) ? mrA : 0u) | ((o & 1) ? mcA : 0u)] = cscale(acc, sc);
    }
  }
}

// ---- whole remote gate on the data pair (A, B) (host: try_fuse_whole_at).  Constants (double2 units; all tables real,
// packed two doubles to an entry): K[m1][pred][order][64 doubles] (2 orders: 1, q; b[7] & 3 = 2: 4 orders 1, q1, q2, q1 q2;
// b[7] & 4: compact, K[m1][order] -- the predicate is the first outcome or absent),
// then T2[m2][64], t2[m2][16], then scratch the kernel fills per CTA: u[2][16] (as double2 with y = 0), C[256 doubles].
// aux = creg bit of m1 | (predicate bit + 1) << 8 | (register of q1 + 1) << 16 | (register of q2 + 1) << 24.
struct WholeSel { const double* K; double q1, q2; int nord, oT2, ot2, oU, oC; };
__device__ __forceinline__ WholeSel whole_sel(const MicroOp& op, const double2* cs, uint64_t cw, const double* treg) {
  WholeSel w;
  const uint32_t aux = (uint32_t)op.aux;
  w.nord = (op.b[7] & 3) == 2 ? 4 : 2;
  const bool compact = (op.b[7] & 4) != 0;   // K[m1][order] only: the predicate is the first outcome (or absent)
  const int nK = (compact ? 64 : 128) * w.nord;   // double2 entries
  w.oT2 = nK; w.ot2 = nK + 64; w.oU = nK + 80; w.oC = nK + 112;
  const int m1 = (int)((cw >> (aux & 0xffu)) & 1ull);
  const int pb = (int)((aux >> 8) & 0xffu) - 1, r1 = (int)((aux >> 16) & 0xffu) - 1, r2 = (int)((aux >> 24) & 0xffu) - 1;
  const int pv = pb >= 0 ? (int)((cw >> pb) & 1ull) : 0;
  w.q1 = r1 >= 0 ? treg[r1] : 0.0;
  w.q2 = r2 >= 0 ? treg[r2] : 0.0;
  w.K = reinterpret_cast<const double*>(cs) + ((compact ? m1 : m1 * 2 + pv) * w.nord) * 64;
  return w;
}
__device__ __forceinline__ double whole_keff(const WholeSel& w, int k) {
  double v = fma(w.q1, w.K[64 + k], w.K[k]);
  if (w.nord == 4) v += w.q2 * fma(w.q1, w.K[192 + k], w.K[128 + k]);
  return v;
}
__device__ __forceinline__ int whole_scratch_u(const MicroOp& op) {
  return ((op.b[7] & 4) ? 64 : 128) * ((op.b[7] & 3) == 2 ? 4 : 2) + 80;
}
// u_m2[(rA' cA' rB' cB')] = sum_{ry cy} t2_m2[(rB' ry cB' cy)] sum_a K[(a a ry cy), (rA' cA')]: the functional that gives
// p(m2 | m1) on the group's 16 elements; threads 0-31 write one entry each
__device__ __noinline__ void whole_build_u(const MicroOp& op, double2* cs, uint64_t cw, const double* treg) {
  if (threadIdx.x >= 32u) return;
  const WholeSel w = whole_sel(op, cs, cw, treg);
  const int m2 = threadIdx.x >> 4, c = threadIdx.x & 15;
  const int rA = (c >> 3) & 1, rB = (c >> 2) & 1, cA = (c >> 1) & 1, cB = c & 1;
  const double* t2 = reinterpret_cast<const double*>(cs + w.ot2) + 16 * m2;
  double acc = 0.0;
#pragma unroll
  for (int yy = 0; yy < 4; ++yy) {
    double kd = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a) kd += whole_keff(w, ((a * 2 + a) * 4 + yy) * 4 + (rA * 2 + cA));
    acc = fma(t2[rB * 8 + (yy >> 1) * 4 + cB * 2 + (yy & 1)], kd, acc);
  }
  cs[w.oU + threadIdx.x] = make_double2(acc, 0.0);
}
// C[(rA cA rB cB), (rA' cA' rB' cB')] = sum_{ry cy} T2_m2[(rB cB), (rB' ry cB' cy)] K[(rA cA ry cy), (rA' cA')], in the
// element order of the (A, B) group: e = rA * 8 + rB * 4 + cA * 2 + cB
__device__ __noinline__ void whole_build_C(const MicroOp& op, double2* cs, uint64_t cw_in, const double* treg, int m2) {
  const WholeSel w = whole_sel(op, cs, cw_in, treg);
  const double* T2 = reinterpret_cast<const double*>(cs + w.oT2) + 64 * m2;
  double* C = reinterpret_cast<double*>(cs + w.oC);
  for (uint32_t t = threadIdx.x; t < 256u; t += blockDim.x) {
    const int eo = t >> 4, ei = t & 15;
    const int rA = (eo >> 3) & 1, rB = (eo >> 2) & 1, cA = (eo >> 1) & 1, cB = eo & 1;
    const int rA2 = (ei >> 3) & 1, rB2 = (ei >> 2) & 1, cA2 = (ei >> 1) & 1, cB2 = ei & 1;
    double acc = 0.0;
#pragma unroll
    for (int yy = 0; yy < 4; ++yy)
      acc = fma(T2[(rB * 2 + cB) * 16 + rB2 * 8 + (yy >> 1) * 4 + cB2 * 2 + (yy & 1)],
                whole_keff(w, ((rA * 2 + cA) * 4 + yy) * 4 + (rA2 * 2 + cA2)), acc);
    C[t] = acc;
  }
}
// every group of the tile: out = scale * C in (a dense real 16 x 16 matrix against 16 complex elements)
__device__ __noinline__ void whole_apply(double2* tile, uint32_t T, const MicroOp& op, const double2* cs, const MeasRec& r) {
  int s4[4];
  sort4(op.b, s4);
  const uint32_t mrA = 1u << op.b[0], mrB = 1u << op.b[1], mcA = 1u << op.b[2], mcB = 1u << op.b[3];
  const double2* __restrict__ C = cs + whole_scratch_u(op) + 32;   // row e: 16 doubles = 8 double2
  const double sc = r.scale;
  for (uint32_t q = threadIdx.x; q < (T >> 4); q += blockDim.x) {
    const uint32_t i0 = ins0_4(q, s4);
    double2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
      v[i] = cscale(tile[i0 | ((i & 8) ? mrA : 0u) | ((i & 4) ? mrB : 0u) | ((i & 2) ? mcA : 0u) | ((i & 1) ? mcB : 0u)], sc);
#pragma unroll 2
    for (int e = 0; e < 16; ++e) {
      const double2* __restrict__ row = C + 8 * e;
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double2 c = row[i];
        acc.x = fma(c.x, v[2 * i].x, acc.x); acc.y = fma(c.x, v[2 * i].y, acc.y);
        acc.x = fma(c.y, v[2 * i + 1].x, acc.x); acc.y = fma(c.y, v[2 * i + 1].y, acc.y);
      }
      tile[i0 | ((e & 8) ? mrA : 0u) | ((e & 4) ? mrB : 0u) | ((e & 2) ? mcA : 0u) | ((e & 1) ? mcB : 0u)] = acc;
    }
  }
}
// form 3: the first measurement of a whole remote gate only renormalises (its map T1_m1 is applied by form 2 later)
__device__ __noinline__ void whole_rescale(double2* tile, uint32_t T, double sc) {
  for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) tile[i] = cscale(tile[i], sc);
}

// split-phase cluster barrier (arrive at kernel start, wait before the first DSMEM write: every CTA of
// the cluster is then known to be running, and nobody stalls for it)
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Fused measurement for registers whose tiles fit one thread-block cluster (<= 16 CTAs per shot):
// every CTA PUSHES its tile partials into the shared memory of all CTAs of the shot (DSMEM stores, fire
// and forget), one cluster barrier, then each CTA sums the 16 local slots in rank order and draws the same
// Philox number -- the pass simply goes on, no kernel boundary, no remote loads, no exit barrier.
template <int FORM>
__device__ __forceinline__ uint64_t op_measure(double2* tile, uint32_t T, const MicroOp& op, const PassParams& pp,
                                               uint64_t base, uint64_t shot, uint64_t cw, double* red,
                                               double* part, MeasRec* mrec, int mcount, double* s_draw,
                                               const double2* cs, const MicroOp* next_op, double* s_treg,
                                               const double2* s_consts, unsigned long long* s_cw) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int par = mcount & 1;
  // the Philox draw does not depend on the probabilities: lane 0 of warp 1 (idle during the diagonal sum of
  // a small tile) computes it now, off the critical path; thread 0 reads it after the cluster barrier
  if (threadIdx.x == 32) {
    double u, u2;
    shot_uniforms(pp.seed, pp.shot_offset + shot, op.octrl, u, u2);
    *s_draw = u;
  }
  // 1. this tile's partials (same arithmetic as op_probs)
  double a0 = 0.0, a1 = 0.0;
  const int abit = op.aux;
  // fused runs whose tile holds at most 32 diagonal groups: warp 0 alone evaluates and reduces them, the other warps
  // go straight to the cluster barrier (no block-wide reduction, no barrier in front of the DSMEM push)
  bool solo = false;
  if constexpr (FORM == 1) solo = op.nb == 4 && !(op.lctrl2 >> 31) && (op.lctrl2 & 0xffu) <= 5u;
  if constexpr (FORM == 0) {
    const int jb = op.b[0];
    if (jb >= kOuter) {
      double a = 0.0;
      for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) { const double2 v = tile[i]; a += fma(v.x, v.x, v.y * v.y); }
      if ((base >> (jb - kOuter)) & 1ull) a1 = a; else a0 = a;
    } else {
      for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {
        const double2 v = tile[i];
        const double w = fma(v.x, v.x, v.y * v.y);
        if ((i >> jb) & 1u) a1 += w; else a0 += w;
      }
    }
  } else {
    if (op.nb == 4) {   // fused runs
      if (op.b[6] == 2) {   // whole remote gate: the functional of p(m2 | m1) is built per CTA first
        whole_build_u(op, const_cast<double2*>(cs), cw, s_treg);
        if (solo) __syncwarp(); else __syncthreads();
      }
      if (!solo || threadIdx.x < 32u) {
        if (op.b[6] == 2) catm2_partials(tile, T, op, pp, cs + whole_scratch_u(op), base, a0, a1);
        else if (op.b[6]) catm2_partials(tile, T, op, pp, cs, base, a0, a1);
        else catm_partials(tile, T, op, pp, cs, base, a0, a1);
      }
    }
    else dm_diag_partials(tile, pp, base, abit, (int)(int8_t)op.b[1], a0, a1);
  }
  if (!solo || threadIdx.x < 32u) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
  }
  if (!solo) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[2 * w] = a0; red[2 * w + 1] = a1; }
  }
  if (mcount == 0 && pp.outer_bits) cluster_wait();   // completes the arrive of the kernel prologue
  if (!solo) __syncthreads();
  double s0 = a0, s1 = a1;   // (thread 0's copy is the one that counts)
  if (threadIdx.x == 0 && !solo) {
    s0 = 0.0; s1 = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s0 += red[2 * i]; s1 += red[2 * i + 1]; }
  }
  if (pp.outer_bits) {
    // 2. push to every CTA of the shot (own slot included)
    const unsigned nr = cluster.num_blocks();
    if (threadIdx.x == 0) {
      const unsigned me = cluster.block_rank();
      for (unsigned r = 0; r < nr; ++r) {
        double* rp = cluster.map_shared_rank(part + 2u * ((unsigned)par * nr + me), r);   // [parity][rank][outcome]
        rp[0] = s0; rp[1] = s1;
      }
    }
    cluster.sync();
    // 3. every CTA sums the slots in rank order (identical result everywhere)
    if (threadIdx.x == 0) {
      s0 = 0.0; s1 = 0.0;
      for (unsigned r = 0; r < nr; ++r) { s0 += part[2u * ((unsigned)par * nr + r)]; s1 += part[2u * ((unsigned)par * nr + r) + 1u]; }
    }
  } else {
    __syncthreads();   // one tile per shot: the CTA's own sums are the shot's (the draw of thread 32 becomes visible here)
  }
  // ... and draws
  if (threadIdx.x == 0) {
    const double p0 = s0, p1 = s1;
    const double e = op.p[0];
    const double q0 = (1.0 - e) * p0 + e * p1, q1 = (1.0 - e) * p1 + e * p0;
    const double tot = q0 + q1;
    const double u = *s_draw;
    const int forced = (int)op.lctrl - 1;
    const int m = forced >= 0 ? forced : ((u * tot < q1) ? 1 : 0);
    const double pm = m ? q1 : q0;
    MeasRec r;
    r.outcome = m; r.pad = 0;
    r.scale = pm > 0.0 ? (FORM == 1 ? 1.0 / pm : 1.0 / sqrt(pm)) : 0.0;
    *mrec = r;
    // the record and the classical word reach global memory when the CTA holding the shot's tile 0 is done
    // (k_tile_pass epilogue, double-buffered copy): a collapse may land in the next pass
    if (pp.olog && cluster.block_rank() == 0) pp.olog[op.octrl * (uint64_t)pp.batch + shot] = (int8_t)m;
  }
  __syncthreads();
  if constexpr (FORM == 1) {
    if (op.nb == 4 && op.b[6] == 2) {   // the 16 x 16 matrix of the drawn outcome (selectors: the word from before this bit)
      whole_build_C(op, const_cast<double2*>(cs), cw, s_treg, mrec->outcome);
      __syncthreads();
    }
  }
  const int cbit = op.flags - 1;
  if (cbit >= 0) cw = (cw & ~(1ull << cbit)) | ((uint64_t)mrec->outcome << cbit);
  if (op.lctrl2 & kMeasRunsClassical) {
    // the classical block hoisted behind this measurement runs right here (planner flag): no dispatch and no
    // barrier of its own; CTAs of four warps let warp 0 run it while the other warps already write the output block
    if (blockDim.x <= 64u || threadIdx.x < 32u) {
      cl_run<true>(cl_payload(*next_op), next_op->flags, s_treg, cw, reinterpret_cast<const double*>(s_consts), pp.seed,
             pp.shot_offset + shot);
      if (threadIdx.x == 0) *s_cw = cw;
    }
  }
  if constexpr (FORM == 1) {
    if (op.nb == 4) {   // the output block of the drawn outcome, renormalised
      if (op.b[6] == 3) whole_rescale(tile, T, mrec->scale);
      else if (op.b[6] == 2) whole_apply(tile, T, op, cs, *mrec);
      else if (op.b[6]) { if (op.b[7]) catm2_apply<true>(tile, T, op, cs, *mrec); else catm2_apply<false>(tile, T, op, cs, *mrec); }
      else { if (op.b[7]) catm_apply<true>(tile, T, op, cs, *mrec); else catm_apply<false>(tile, T, op, cs, *mrec); }
    }
  }
  return cw;
}

template <int FORM>
__device__ __forceinline__ void op_collapse(double2* tile, uint32_t T, const MicroOp& op, const PassParams& pp,
                                            const MeasRec& r) {
  const int m = r.outcome;
  const double sc = r.scale;
  const bool discard = op.flags & 1;
  const double2 zero = make_double2(0.0, 0.0);
  if constexpr (FORM == 0) {
    const int j = op.b[0];
    for (uint32_t p = threadIdx.x; p < (T >> 1); p += blockDim.x) {
      const uint32_t i0 = ins0(p, j), i1 = i0 | (1u << j);
      const double2 kept = cscale(tile[m ? i1 : i0], sc);
      if (discard || m == 0) { tile[i0] = kept; tile[i1] = zero; }
      else { tile[i1] = kept; tile[i0] = zero; }
    }
  } else {
    const int jr = op.b[0], jc = op.b[1];
    const int jlo = jr < jc ? jr : jc, jhi = jr < jc ? jc : jr;
    const double e = op.p[0];
    for (uint32_t p = threadIdx.x; p < (T >> 2); p += blockDim.x) {
      const uint32_t i00 = ins0(ins0(p, jlo), jhi);
      const uint32_t i01 = i00 | (1u << jc), i10 = i00 | (1u << jr), i11 = i01 | (1u << jr);
      const double2 same = tile[m ? i11 : i00], flip = tile[m ? i00 : i11];
      double2 kept;
      kept.x = ((1.0 - e) * same.x + e * flip.x) * sc;
      kept.y = ((1.0 - e) * same.y + e * flip.y) * sc;
      tile[i01] = zero; tile[i10] = zero;
      if (discard || m == 0) { tile[i00] = kept; tile[i11] = zero; }
      else { tile[i11] = kept; tile[i00] = zero; }
    }
  }
}

__device__ __forceinline__ void op_trace(double2* tile, uint32_t T, const MicroOp& op) {
  const int jr = op.b[0], jc = op.b[1];
  const int jlo = jr < jc ? jr : jc, jhi = jr < jc ? jc : jr;
  const double2 zero = make_double2(0.0, 0.0);
  for (uint32_t p = threadIdx.x; p < (T >> 2); p += blockDim.x) {
    const uint32_t i00 = ins0(ins0(p, jlo), jhi);
    const uint32_t i01 = i00 | (1u << jc), i10 = i00 | (1u << jr), i11 = i01 | (1u << jr);
    tile[i00] = cadd(tile[i00], tile[i11]);
    tile[i01] = zero; tile[i10] = zero; tile[i11] = zero;
  }
}

template <int FORM>
__device__ __forceinline__ void op_inject(double2* tile, uint32_t T, const MicroOp& op, const PassParams& pp,
                                          const double2* cs, uint64_t shot) {
  if constexpr (FORM == 0) {
    const int ja = op.b[0], jb = op.b[1];
    const int jlo = ja < jb ? ja : jb, jhi = ja < jb ? jb : ja;
    int choice = 0;
    if (op.flags & 2) {  // mixture: component drawn per shot
      double u, u2;
      shot_uniforms(pp.seed, pp.shot_offset + shot, (uint64_t)(uint32_t)op.aux, u, u2);
      choice = (u >= op.p[0]) + (u >= op.p[1]) + (u >= op.p[2]);
    }
    const double2 k0 = cs[4 * choice], k1 = cs[4 * choice + 1],
                  k2 = cs[4 * choice + 2], k3 = cs[4 * choice + 3];
    for (uint32_t p = threadIdx.x; p < (T >> 2); p += blockDim.x) {
      const uint32_t i0 = ins0(ins0(p, jlo), jhi);
      const double2 x = tile[i0];
      tile[i0] = cmul(x, k0);
      tile[i0 | (1u << jb)] = cmul(x, k1);
      tile[i0 | (1u << ja)] = cmul(x, k2);
      tile[i0 | (1u << ja) | (1u << jb)] = cmul(x, k3);
    }
  } else {
    int s[4];
    sort4(op.b, s);
    const uint32_t mra = 1u << op.b[0], mrb = 1u << op.b[1], mca = 1u << op.b[2], mcb = 1u << op.b[3];
    for (uint32_t p = threadIdx.x; p < (T >> 4); p += blockDim.x) {
      const uint32_t i0 = ins0_4(p, s);
      const double2 x = tile[i0];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t ro = ((r & 2) ? mra : 0u) | ((r & 1) ? mrb : 0u);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t idx = i0 | ro | ((c & 2) ? mca : 0u) | ((c & 1) ? mcb : 0u);
          tile[idx] = cmul(x, cs[4 * r + c]);
        }
      }
    }
  }
}

// ket: Pauli string drawn per shot.  flags: 0 = 1q channel with cumulative thresholds p[0..2] on u,
// 1 = k-qubit depolarising (w0 = p[0], w = p[1]) on u, 2 = X if u2 < p[0] (measurement flip).
__device__ __forceinline__ void op_pauli_sampled(double2* tile, uint32_t T, const MicroOp& op,
                                                 const PassParams& pp, uint64_t shot, const double* treg,
                                                 const double2* cs) {
  double u, u2;
  shot_uniforms(pp.seed, pp.shot_offset + shot, (uint64_t)(uint32_t)op.aux, u, u2);
  const int k = op.nb;
  int digits[2] = {0, 0};
  if (op.flags == 3) {          // depolarising of per-shot strength q = register p[3] (Pauli weights q / 4)
    const double w = 0.25 * treg[(int)op.p[3]], w0 = 1.0 - 3.0 * w;
    long long s = 0;
    if (!(u < w0) && w > 0.0) { s = 1 + (long long)floor((u - w0) / w); if (s > 3) s = 3; }
    digits[0] = (int)s;
  } else if (op.flags == 4) {   // dephasing of per-shot strength q: Z with probability q / 2
    digits[0] = (u >= 1.0 - 0.5 * treg[(int)op.p[3]]) ? 3 : 0;
  } else if (op.flags == 5) {   // Z with per-shot probability q
    digits[0] = (u >= 1.0 - treg[(int)op.p[3]]) ? 3 : 0;
  } else if (op.flags == 6) {   // general two-qubit Pauli channel: 15 cumulative weights in the constant pool
    const double* cum = reinterpret_cast<const double*>(cs);
    int s = 0;
    for (int i = 0; i < 15; ++i) s += (u >= cum[i]) ? 1 : 0;
    digits[0] = (s >> 2) & 3; digits[1] = s & 3;
  } else if (op.flags == 0) {
    digits[0] = (u >= op.p[0]) + (u >= op.p[1]) + (u >= op.p[2]);
  } else if (op.flags == 1) {
    const double w0 = op.p[0], w = op.p[1];
    const int n = (k == 1) ? 4 : 16;
    long long s = 0;
    if (!(u < w0) && w > 0.0) {
      s = 1 + (long long)floor((u - w0) / w);
      if (s > n - 1) s = n - 1;
    }
    if (k == 1) digits[0] = (int)s;
    else { digits[0] = (int)(s >> 2) & 3; digits[1] = (int)s & 3; }
  } else {
    digits[0] = (u2 < op.p[0]) ? 1 : 0;
  }
  const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0), mone = make_double2(-1.0, 0.0);
  const double2 pi = make_double2(0.0, 1.0), mi = make_double2(0.0, -1.0);
  for (int q = 0; q < k; ++q) {
    const int d = digits[q];
    if (d == 0) continue;
    if (q > 0) __syncthreads();
    if (d == 1) apply_mat1(tile, T, op.b[q], 0u, zero, one, one, zero);
    else if (d == 2) apply_mat1(tile, T, op.b[q], 0u, zero, mi, pi, zero);
    else apply_mat1(tile, T, op.b[q], 0u, one, zero, zero, mone);
  }
}


// ------------------------------------------------------------------------------------------
// dm group sweep.  Register file layout: v[e], e = ra*8 + rb*4 + ca*2 + cb  (A = first axis).
template <int ROLE>
__device__ __forceinline__ constexpr int EI(int r, int c, int ro, int co) {
  return ROLE == 0 ? (r * 8 + ro * 4 + c * 2 + co) : (ro * 8 + r * 4 + co * 2 + c);
}

// The group sub-ops below are written "in place": an output is assigned at the instruction where the input
// it replaces dies, and the only temporaries are partial products computed BEFORE that point.  The sub-op
// loop of run_group is a switch inside a loop, so v[] must sit in the same registers at every merge; with
// read-all-then-write-all bodies ptxas could not coalesce and copied the register file (64+ IMAD.MOV per
// sub-op, 15 % of all executed instructions in the r01k profile).  Arithmetic and rounding are unchanged.
__device__ __forceinline__ void mix2(double2& x0, double2& x1, double2 m00, double2 m01, double2 m10, double2 m11) {
  const double2 p = cmul(m10, x0);
  x0 = cfma(m01, x1, cmul(m00, x0));
  x1 = cfma(m11, x1, p);
}
__device__ __forceinline__ void mix2c(double2& x0, double2& x1, double2 m00, double2 m01, double2 m10, double2 m11) {
  const double2 p = cmulc(m10, x0);   // conjugated coefficients (column side of U rho U^dagger)
  x0 = cfmac(m01, x1, cmulc(m00, x0));
  x1 = cfmac(m11, x1, p);
}

template <int ROLE>
__device__ __forceinline__ void g_pauli1(double2 (&v)[16], const MicroOp& op) {
  const double a = op.p[0], b = op.p[1], c = op.p[2], d = op.p[3];
#pragma unroll
  for (int ro = 0; ro < 2; ++ro)
#pragma unroll
    for (int co = 0; co < 2; ++co) {
      double2& e00 = v[EI<ROLE>(0, 0, ro, co)]; double2& e01 = v[EI<ROLE>(0, 1, ro, co)];
      double2& e10 = v[EI<ROLE>(1, 0, ro, co)]; double2& e11 = v[EI<ROLE>(1, 1, ro, co)];
      const double t0x = b * e00.x, t0y = b * e00.y, t1x = d * e01.x, t1y = d * e01.y;
      e00.x = fma(a, e00.x, b * e11.x); e00.y = fma(a, e00.y, b * e11.y);
      e11.x = fma(a, e11.x, t0x); e11.y = fma(a, e11.y, t0y);
      e01.x = fma(c, e01.x, d * e10.x); e01.y = fma(c, e01.y, d * e10.y);
      e10.x = fma(c, e10.x, t1x); e10.y = fma(c, e10.y, t1y);
    }
}

template <int ROLE>
__device__ __forceinline__ void g_dyn1(double2 (&v)[16], const DynCoef k) {
#pragma unroll
  for (int ro = 0; ro < 2; ++ro)
#pragma unroll
    for (int co = 0; co < 2; ++co) {
      double2& e00 = v[EI<ROLE>(0, 0, ro, co)]; double2& e01 = v[EI<ROLE>(0, 1, ro, co)];
      double2& e10 = v[EI<ROLE>(1, 0, ro, co)]; double2& e11 = v[EI<ROLE>(1, 1, ro, co)];
      const double t0x = k.c * e00.x, t0y = k.c * e00.y;
      e00.x = fma(k.a, e00.x, k.b * e11.x); e00.y = fma(k.a, e00.y, k.b * e11.y);
      e11.x = fma(k.d, e11.x, t0x); e11.y = fma(k.d, e11.y, t0y);
      e01.x *= k.f; e01.y *= k.f;
      e10.x *= k.f; e10.y *= k.f;
    }
}

template <int ROLE>
__device__ __forceinline__ void g_xpat(double2 (&v)[16], const double2* __restrict__ cs) {
  const double2 a = cs[0], b = cs[1], c = cs[2], d = cs[3], f = cs[4], g = cs[5], h = cs[6], k = cs[7];
#pragma unroll
  for (int ro = 0; ro < 2; ++ro)
#pragma unroll
    for (int co = 0; co < 2; ++co) {
      mix2(v[EI<ROLE>(0, 0, ro, co)], v[EI<ROLE>(1, 1, ro, co)], a, b, c, d);
      mix2(v[EI<ROLE>(0, 1, ro, co)], v[EI<ROLE>(1, 0, ro, co)], f, g, h, k);
    }
}

// Controlled-X / X inside a group is a permutation of the thread's 16 elements -- done in the ADDRESSES of
// the group's shared-memory loads / stores, never in registers: element e = (ra, rb, ca, cb) trades places
// with the element whose target row (column) bit is flipped when the row (column) control is on.
// r[(ra << 1) | rb] / c[(ca << 1) | cb] are the XOR masks on the element's tile index (all warp-uniform).
struct GPerm { uint32_t r[4], c[4]; };
// flags: bit0 = role of the target axis, bit1 / bit2 = row / column control is the other axis of the group
__device__ __forceinline__ GPerm make_perm(bool on, int flags, uint64_t octrl, uint64_t octrl2, uint64_t base,
                                           uint32_t mra, uint32_t mrb, uint32_t mca, uint32_t mcb) {
  const int role = flags & 1;
  const bool in_r = flags & 2, in_c = flags & 4;
  const bool r_on = on && (base & octrl) == octrl, c_on = on && (base & octrl2) == octrl2;
  const uint32_t tr = role ? mrb : mra, tc = role ? mcb : mca;
  const uint32_t r1 = r_on ? tr : 0u, r0 = (r_on && !in_r) ? tr : 0u;
  const uint32_t c1 = c_on ? tc : 0u, c0 = (c_on && !in_c) ? tc : 0u;
  GPerm p;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ctl = role ? (i >> 1) : (i & 1);   // value of the control-axis bit of the element
    p.r[i] = ctl ? r1 : r0;
    p.c[i] = ctl ? c1 : c0;
  }
  return p;
}

// target axis = ROLE; the control (if any) is either the other axis of the group (in_r / in_c)
// or outside the tile (r_on / c_on already evaluated on the tile base)
template <int ROLE>
__device__ __forceinline__ void g_mat1x2(double2 (&v)[16], const double2* __restrict__ cs, bool r_on, bool c_on,
                                         bool in_r, bool in_c) {
  const double2 m00 = cs[0], m01 = cs[1], m10 = cs[2], m11 = cs[3];
  if (r_on) {
#pragma unroll
    for (int ro = 0; ro < 2; ++ro) {
      if (in_r && ro == 0) continue;
#pragma unroll
      for (int co = 0; co < 2; ++co)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          mix2(v[EI<ROLE>(0, c, ro, co)], v[EI<ROLE>(1, c, ro, co)], m00, m01, m10, m11);
        }
    }
  }
  if (c_on) {
#pragma unroll
    for (int co = 0; co < 2; ++co) {
      if (in_c && co == 0) continue;
#pragma unroll
      for (int ro = 0; ro < 2; ++ro)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          mix2c(v[EI<ROLE>(r, 0, ro, co)], v[EI<ROLE>(r, 1, ro, co)], m00, m01, m10, m11);
        }
    }
  }
}

__device__ __forceinline__ void g_depol2(double2 (&v)[16], double pp) {
  const double keep = 1.0 - pp, mix = 0.25 * pp;
  const double sx = mix * (v[0].x + v[5].x + v[10].x + v[15].x);
  const double sy = mix * (v[0].y + v[5].y + v[10].y + v[15].y);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    v[e].x *= keep; v[e].y *= keep;
    if ((e >> 2) == (e & 3)) { v[e].x += sx; v[e].y += sy; }
  }
}

template <int ROLE>
__device__ __forceinline__ void g_collapse(double2 (&v)[16], int m, double sc, double e, bool discard) {
  const double2 zero = make_double2(0.0, 0.0);
#pragma unroll
  for (int ro = 0; ro < 2; ++ro)
#pragma unroll
    for (int co = 0; co < 2; ++co) {
      const double2 e00 = v[EI<ROLE>(0, 0, ro, co)], e11 = v[EI<ROLE>(1, 1, ro, co)];
      const double2 same = m ? e11 : e00, flip = m ? e00 : e11;
      double2 kept;
      kept.x = ((1.0 - e) * same.x + e * flip.x) * sc;
      kept.y = ((1.0 - e) * same.y + e * flip.y) * sc;
      const bool to0 = discard || m == 0;
      v[EI<ROLE>(0, 0, ro, co)] = to0 ? kept : zero;
      v[EI<ROLE>(1, 1, ro, co)] = to0 ? zero : kept;
      v[EI<ROLE>(0, 1, ro, co)] = zero;
      v[EI<ROLE>(1, 0, ro, co)] = zero;
    }
}

template <int ROLE>
__device__ __forceinline__ void g_trace(double2 (&v)[16]) {
  const double2 zero = make_double2(0.0, 0.0);
#pragma unroll
  for (int ro = 0; ro < 2; ++ro)
#pragma unroll
    for (int co = 0; co < 2; ++co) {
      v[EI<ROLE>(0, 0, ro, co)] = cadd(v[EI<ROLE>(0, 0, ro, co)], v[EI<ROLE>(1, 1, ro, co)]);
      v[EI<ROLE>(0, 1, ro, co)] = zero; v[EI<ROLE>(1, 0, ro, co)] = zero; v[EI<ROLE>(1, 1, ro, co)] = zero;
    }
}

__device__ __forceinline__ void run_group(double2* tile, uint32_t T, const MicroOp& g, const MicroOp* sub,
                                          const PassParams& pp, const double2* cs_base, uint64_t base,
                                          uint64_t cw, const MeasRec* mrec, const double* treg) {
  const int s4[4] = {g.b[4], g.b[5], g.b[6], g.b[7]};   // the four bit positions, ascending (host-sorted)
  // through a shuffle: the compiler then knows the masks (and every combination of them) are warp-uniform
  const uint32_t mra = __shfl_sync(0xffffffffu, 1u << g.b[0], 0), mrb = __shfl_sync(0xffffffffu, 1u << g.b[1], 0);
  const uint32_t mca = __shfl_sync(0xffffffffu, 1u << g.b[2], 0), mcb = __shfl_sync(0xffffffffu, 1u << g.b[3], 0);
  const uint32_t bra = mra << 4, brb = mrb << 4, bca = mca << 4, bcb = mcb << 4;   // as byte offsets
  const int nsub = g.aux;
  // X / CX hoisted by the planner to the head (flags bits 0-3: on, role, in_r, in_c; controls octrl / octrl2)
  // or the tail (bits 4-7; controls in p[0], p[1]) of the group: applied by the load / store addresses
  const uint64_t* tail_ctrl = reinterpret_cast<const uint64_t*>(&g.p[0]);
  const bool has_pre = g.flags & 1, has_post = g.flags & 16;   // most groups carry neither: plain addresses
  GPerm pre, post;
  if (has_pre) pre = make_perm(true, (g.flags >> 1) & 7, g.octrl, g.octrl2, base, bra, brb, bca, bcb);
  if (has_post) post = make_perm(true, (g.flags >> 5) & 7, tail_ctrl[0], tail_ctrl[1], base, bra, brb, bca, bcb);
  // element e of the thread's group sits at byte offset DQE_GOFF(e) from &tile[i0]: i0 has zeros at the four
  // group bits, so the offsets are warp-uniform BYTE offsets (uniform datapath) and each access is one
  // LDS / STS [R + UR] with no per-thread index arithmetic
#define DQE_GOFF(e) ((((e) & 8) ? bra : 0u) | (((e) & 4) ? brb : 0u) | (((e) & 2) ? bca : 0u) | (((e) & 1) ? bcb : 0u))
#define DQE_GEL(off) (*reinterpret_cast<double2*>(tpb + (off)))
  for (uint32_t q = threadIdx.x; q < (T >> 4); q += blockDim.x) {
    const uint32_t i0 = ins0_4(q, s4);
    char* __restrict__ tpb = reinterpret_cast<char*>(tile + i0);
    double2 v[16];
    if (has_pre) {
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = DQE_GEL(DQE_GOFF(e) ^ pre.r[(e >> 2) & 3] ^ pre.c[e & 3]);
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = DQE_GEL(DQE_GOFF(e));
    }
    int4 hdr = nsub > 0 ? op_header(sub) : make_int4(0, -1, 0, 0);
    for (int s = 0; s < nsub; ++s) {
      const MicroOp& op = sub[s];
      const int op_type = hdr.x, op_pred = hdr.y, op_flags = hdr.z, op_coff = hdr.w;
      if (s + 1 < nsub) hdr = op_header(sub + s + 1);   // next sub-op's header: its latency hides behind this one
      if (op_pred >= 0 && !((cw >> op_pred) & 1ull)) continue;
      const double2* cs = cs_base + op_coff;
      const int role = op_flags & 1;
      switch (op_type) {
        case G_PAULI1:
          if (role) g_pauli1<1>(v, op); else g_pauli1<0>(v, op);
          break;
        case G_XPAT:
          if (role) g_xpat<1>(v, cs); else g_xpat<0>(v, cs);
          break;
        case G_DYN1: {
          const DynCoef k = dyn_coef((op_flags >> 1) & 3, treg[op.aux]);
          if (role) g_dyn1<1>(v, k); else g_dyn1<0>(v, k);
          break;
        }
        case G_MAT1X2: {
          if (op_flags & 16) {
            // X / CX in the middle of a group: a round trip through the thread's own 16 tile slots with
            // permuted store addresses (no barrier: nobody else touches them during the sweep)
            const GPerm mid = make_perm(true, op_flags & 7, op.octrl, op.octrl2, base, bra, brb, bca, bcb);
#pragma unroll
            for (int e = 0; e < 16; ++e) DQE_GEL(DQE_GOFF(e) ^ mid.r[(e >> 2) & 3] ^ mid.c[e & 3]) = v[e];
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = DQE_GEL(DQE_GOFF(e));
            break;
          }
          const bool in_r = op_flags & 2, in_c = op_flags & 4;
          const bool r_on = (base & op.octrl) == op.octrl, c_on = (base & op.octrl2) == op.octrl2;
          if (role) g_mat1x2<1>(v, cs, r_on, c_on, in_r, in_c);
          else g_mat1x2<0>(v, cs, r_on, c_on, in_r, in_c);
          break;
        }
        case G_DEPOL2: g_depol2(v, op.p[0]); break;
        case G_COLLAPSE: {
          const MeasRec r = *mrec;
          if (role) g_collapse<1>(v, r.outcome, r.scale, op.p[0], op_flags & 8);
          else g_collapse<0>(v, r.outcome, r.scale, op.p[0], op_flags & 8);
          break;
        }
        case G_TRACE:
          if (role) g_trace<1>(v); else g_trace<0>(v);
          break;
        case G_INJECT: {
          const double2 x = v[0];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = cmul(x, cs[e]);
          break;
        }
        default: break;
      }
    }
    if (has_post) {
#pragma unroll
      for (int e = 0; e < 16; ++e) DQE_GEL(DQE_GOFF(e) ^ post.r[(e >> 2) & 3] ^ post.c[e & 3]) = v[e];
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) DQE_GEL(DQE_GOFF(e)) = v[e];
    }
  }
#undef DQE_GOFF
#undef DQE_GEL
}

// every stand-alone density-matrix sweep (ops that ride neither in a group nor in a fused measurement)
__device__ __noinline__ void dm_standalone(double2* tile, uint32_t T, const MicroOp& op, int op_type, const PassParams& pp,
                                           const double2* cs, uint64_t base, uint64_t blk, uint64_t shot, double* s_red,
                                           const MeasRec* s_meas, const double* s_treg) {
  switch (op_type) {
    case OP_MAT1X2: {
      const bool r_on = (base & op.octrl) == op.octrl, c_on = (base & op.octrl2) == op.octrl2;
      if (r_on || c_on) op_mat1x2(tile, T, op, cs, r_on, c_on);
      break;
    }
    case OP_MAT2:
      if ((base & op.octrl) == op.octrl) op_mat2(tile, T, op, cs);
      break;
    case OP_DIAG: op_diag(tile, T, op, cs, base); break;
    case OP_PAULI1: op_pauli1(tile, T, op); break;
    case OP_XPAT: op_xpat(tile, T, op, cs); break;
    case OP_DYN1: op_dyn1(tile, T, op, s_treg); break;
    case OP_DEPOL2: op_depol2(tile, T, op); break;
    case OP_KBLOCK:
      if (op.nb == 4) op_kblock<4>(tile, T, op, cs); else op_kblock<3>(tile, T, op, cs);
      break;
    case OP_PROBS: op_probs<1>(tile, T, op, pp, base, blk, s_red); break;
    case OP_COLLAPSE: op_collapse<1>(tile, T, op, pp, *s_meas); break;
    case OP_TRACE: op_trace(tile, T, op); break;
    case OP_INJECT: op_inject<1>(tile, T, op, pp, cs, shot); break;
    default: break;
  }
}

// ------------------------------------------------------------------------------------------
// issue the bulk loads r0, r0+rstep, ... of one tile (a tile = 2^(t-c0) contiguous runs); the
// issuing lanes of different warps work in parallel -- a single thread needs ~100 cycles per copy
__device__ __forceinline__ void tma_expect(uint32_t bar_s, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_bytes(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar_s) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s),
               "l"(src), "r"(bytes), "r"(bar_s)
               : "memory");
}
// whole-tile tensor copies (rank 3..5); c[] = box origin, all zero but the base dimension
__device__ __forceinline__ void tma_tensor_load(const CUtensorMap* tm, int rank, uint32_t dst_s, uint32_t bar_s, const int32_t* c) {
  const uint64_t tp = reinterpret_cast<uint64_t>(tm);
  if (rank == 3)
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst_s), "l"(tp), "r"(bar_s), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
  else if (rank == 4)
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst_s), "l"(tp), "r"(bar_s), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst_s), "l"(tp), "r"(bar_s), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
__device__ __forceinline__ void tma_tensor_store(const CUtensorMap* tm, int rank, uint32_t src_s, const int32_t* c) {
  const uint64_t tp = reinterpret_cast<uint64_t>(tm);
  if (rank == 3)
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tp), "r"(src_s), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
  else if (rank == 4)
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(tp), "r"(src_s), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(tp), "r"(src_s), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
__device__ __forceinline__ void tma_tensor_prefetch(const CUtensorMap* tm, int rank, const int32_t* c) {
  const uint64_t tp = reinterpret_cast<uint64_t>(tm);
  if (rank == 3)
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tp), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
  else if (rank == 4)
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tp), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
  else
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(tp), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
__device__ __forceinline__ void tm_coords(const PassParams& pp, uint64_t blk, int32_t* c) {
  const uint64_t base = deposit(blk & ((1ull << pp.outer_bits) - 1ull), pp.outer_segs, pp.n_outer_segs);
  const uint64_t off = ((blk >> pp.outer_bits) << pp.P) + base;
#pragma unroll
  for (int i = 0; i < 5; ++i) c[i] = 0;
  const int32_t cb = (int32_t)(off >> pp.tm_shift);
#pragma unroll
  for (int i = 1; i < 5; ++i) if (i == pp.tm_base_dim) c[i] = cb;
}

// r0 / rstep must be warp-uniform (every lane runs the loop with the same operands so that the
// address arithmetic stays on the uniform datapath); only the lane with `leader` set issues.
__device__ __forceinline__ void tma_load_tile(const PassParams& pp, const double2* sp, uint32_t dst_s, uint32_t bar_s,
                                              uint32_t T, int c0, uint32_t r0, uint32_t rstep, bool leader) {
  const uint32_t run_bytes = 16u << c0, nruns = T >> c0;
  for (uint32_t r = r0; r < nruns; r += rstep) {
    const double2* src = sp + ((uint64_t)pp.run_hi[r] << c0);
    const uint32_t dst = dst_s + r * run_bytes;
    if (leader)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(src), "r"(run_bytes), "r"(bar_s)
                   : "memory");
  }
}

// %tid.x / %ctaid.x through volatile asm: the store phase of the one-tile kernel RE-derives its addressing
// state instead of keeping a dozen registers alive across the micro-program (the group sweeps need them)
__device__ __forceinline__ uint32_t fresh_tid() { uint32_t t; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t)); return t; }
__device__ __forceinline__ uint32_t fresh_ctaid() { uint32_t t; asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(t)); return t; }

// Persistent, double-buffered pass kernel: each CTA walks tiles blockIdx.x, +gridDim.x, ... ; while
// the micro-program runs on the tile in buffer b, the TMA load of the CTA's next tile streams into
// buffer b^1 and the bulk store of the previous tile drains -- the HBM pipe never idles behind the
// shared-memory / FP64 work of a fused pass.
// PERSIST = false: one tile per CTA (grid = number of tiles) -- the loop state of the persistent form
// (tile counter, buffer parity, prefetch) disappears at compile time.
template <int MAXT, int MINB, int FORM, bool PERSIST>
__global__ void __launch_bounds__(MAXT, MINB) k_tile_pass(const __grid_constant__ PassParams pp) {
  extern __shared__ __align__(128) double2 dyn_tiles[];
  __shared__ double s_red[2 * (kThreads / 32)];
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ DiagPrep s_diag[FORM == 0 ? kMaxDiagRun : 1];   // ket only (multi-diagonal sweeps)
  __shared__ LadderPrep s_lad[FORM == 0 ? kMaxDiagRun : 1];
  __shared__ double2 s_let[FORM == 0 ? kMaxDiagRun * 16 : 1];   // ket block sweeps: element-phase tables of the stars
  __shared__ MeasRec s_meas;
  __shared__ double s_draw;
  __shared__ unsigned long long s_cw;   // classical word handed back by the warp that ran a classical block

  const uint32_t T = 1u << pp.t;
  const int c0 = pp.tile_segs[0].len;
  const bool tma = pp.use_tma != 0;   // planner: low segment contiguous, <= 128 runs per tile
  const int nbuf = PERSIST ? pp.nbuf : 1;   // 2 with TMA (prefetch distance 1), 1 otherwise
  double2* const s_consts = dyn_tiles + ((size_t)nbuf << pp.t);   // the pass's constants, behind the tile buffer(s)
  // ... and its micro-ops behind the constants, sized by the pass (a fixed 48-op array cost the whole-gate passes of
  // small density matrices their eighth resident CTA)
  MicroOp* const s_ops = reinterpret_cast<MicroOp*>(s_consts + pp.cap_consts);
  // the shot's classical registers (simulated times, channel strengths): the nregs in use, behind the micro-ops
  double* const s_treg = reinterpret_cast<double*>(s_ops + pp.cap_ops);
  // [measurement parity][cluster rank][outcome] partials pushed by the shot's CTAs: 32 B for a one-tile shot, behind
  // the registers (a fixed [2][16][2] cost the two-channel whole-gate passes their eighth resident CTA)
  double* const s_part = s_treg + pp.nregs;
  const uint32_t run_bytes = 16u << c0, nruns = T >> c0;
  const uint64_t outer_mask = (1ull << pp.outer_bits) - 1ull;

  const bool tmx = !PERSIST && tma && pp.tm_rank != 0;   // tensor-map path (one copy per tile and direction)
  const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(dyn_tiles);
  const uint32_t bar0_s = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
  uint64_t blk = pp.tile0 + blockIdx.x;
  // per-shot classical state of the first tile: requested before anything else, so that its DRAM latency
  // hides behind the barrier set-up and the tile fetch
  uint64_t cw_first = 0;
  MeasRec mr_first;
  mr_first.scale = 0.0; mr_first.outcome = 0; mr_first.pad = 0;
  if (blk < pp.ntiles) {
    const uint64_t shot0 = blk >> pp.outer_bits;
    cw_first = pp.creg[shot0];
    if (threadIdx.x == 0) mr_first = pp.rec[shot0];   // outcome recorded by a preceding pass / k_finish_measure
  }
  if (pp.treg && !tma && blk < pp.ntiles) {   // (with TMA the registers ride on the tile's barrier, below)
    const double* tr = pp.treg + (blk >> pp.outer_bits) * (uint64_t)kTimeRegs;
    for (int i = threadIdx.x; i < pp.nregs; i += blockDim.x) s_treg[i] = tr[i];
  }
  if (pp.has_measure && pp.outer_bits) cluster_arrive_relaxed();   // waited for in the first op_measure
  if (tma && threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0_s));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0_s + 8u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (blk < pp.ntiles) {
      // the micro-program and its constants ride on the first tile's barrier as two more bulk copies
      const uint32_t ops_bytes = (uint32_t)pp.nops * (uint32_t)sizeof(MicroOp), c_bytes = (uint32_t)pp.nconsts * 16u;
      const uint32_t r_bytes = pp.treg ? (uint32_t)pp.nregs * 8u : 0u;   // the shot's classical registers
      tma_expect(bar0_s, T * 16u + ops_bytes + c_bytes + r_bytes);
      if (r_bytes)
        tma_load_bytes((uint32_t)__cvta_generic_to_shared(s_treg), pp.treg + (blk >> pp.outer_bits) * (uint64_t)kTimeRegs,
                       r_bytes, bar0_s);
      if (tmx) {   // the whole tile in one tensor copy
        int32_t c[5];
        tm_coords(pp, blk, c);
        tma_tensor_load(&pp.tmap, pp.tm_rank, tiles_s, bar0_s, c);
        // ... and the tile of the CTA that will take this slot next (one wave of resident CTAs ahead) into L2
        const uint64_t ahead = blk + (uint64_t)pp.prefetch_dist;
        if (pp.prefetch_dist > 0 && ahead < pp.ntiles) {
          tm_coords(pp, ahead, c);
          tma_tensor_prefetch(&pp.tmap, pp.tm_rank, c);
        }
      }
      tma_load_bytes((uint32_t)__cvta_generic_to_shared(s_ops), pp.ops, ops_bytes, bar0_s);
      if (c_bytes) tma_load_bytes((uint32_t)__cvta_generic_to_shared(s_consts), pp.consts, c_bytes, bar0_s);
    }
  }
  __syncthreads();
  // warp index through a shuffle: the compiler then knows it is warp-uniform (CUTLASS idiom)
  const uint32_t wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarps = blockDim.x >> 5;
  const bool issuer = (threadIdx.x & 31u) == 0u;   // lane 0 of every warp issues its share of the copies
  if (tma && !tmx && blk < pp.ntiles) {
    const uint64_t base0 = deposit(blk & outer_mask, pp.outer_segs, pp.n_outer_segs);
    tma_load_tile(pp, pp.state + ((blk >> pp.outer_bits) << pp.P) + base0, tiles_s, bar0_s, T, c0, wid, nwarps, issuer);
  }
  if (!tma) {  // stage the micro-program with ordinary loads
    const int nw = pp.nops * (int)(sizeof(MicroOp) / 16);
    const uint4* src = reinterpret_cast<const uint4*>(pp.ops);
    uint4* dst = reinterpret_cast<uint4*>(s_ops);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < pp.nconsts; i += blockDim.x) s_consts[i] = __ldg(pp.consts + i);
    __syncthreads();
  }

  uint32_t parity[2] = {0u, 0u};
  for (uint32_t it = 0; blk < pp.ntiles; blk += gridDim.x, ++it) {
    const int b = (nbuf == 2) ? (int)(it & 1u) : 0;
    double2* tile = dyn_tiles + (size_t)b * T;
    const uint32_t tile_s = tiles_s + (uint32_t)b * T * 16u;
    const uint64_t shot = blk >> pp.outer_bits;
    const uint64_t base = deposit(blk & outer_mask, pp.outer_segs, pp.n_outer_segs);
    double2* __restrict__ sp = pp.state + (shot << pp.P) + base;
    uint64_t cw = cw_first;
    MeasRec mr0 = mr_first;
    if (PERSIST && it > 0) {
      cw = pp.creg[shot];
      if (threadIdx.x == 0) mr0 = pp.rec[shot];
      if (pp.treg) {
        const double* tr = pp.treg + shot * (uint64_t)kTimeRegs;
        for (int i = threadIdx.x; i < pp.nregs; i += blockDim.x) s_treg[i] = tr[i];
      }
    }

    if (tma) {
      const uint64_t nxt = blk + gridDim.x;
      if (PERSIST && threadIdx.x == 0 && nxt < pp.ntiles) {
        // buffer b^1 was stored from at the end of the previous iteration: its reads must be done
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        const uint64_t nbase = deposit(nxt & outer_mask, pp.outer_segs, pp.n_outer_segs);
        tma_expect(bar0_s + 8u * (uint32_t)(b ^ 1), T * 16u);
        tma_load_tile(pp, pp.state + ((nxt >> pp.outer_bits) << pp.P) + nbase, tiles_s + (uint32_t)(b ^ 1) * T * 16u,
                      bar0_s + 8u * (uint32_t)(b ^ 1), T, c0, 0u, 1u, true);
      }
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar0_s + 8u * (uint32_t)b), "r"(parity[b])
            : "memory");
      }
      parity[b] ^= 1u;
    } else {
      if (pp.tile_segs[0].dst == 0 && c0 >= 8) {
        const int nhs = pp.n_tile_segs - 1;
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {
          const uint64_t off = (uint64_t)(i & ((1u << c0) - 1u)) | deposit((uint64_t)i, pp.tile_segs + 1, nhs);
          tile[i] = sp[off];
        }
      } else {
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < T; i += blockDim.x)
          tile[i] = sp[deposit((uint64_t)i, pp.tile_segs, pp.n_tile_segs)];
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) s_meas = mr0;
    __syncthreads();
    int mcount = 0;
    bool take_cw = false;

    // the dispatch header of the NEXT op is fetched before the current op runs (one 16-byte load, its
    // latency hidden behind the op); a group header carries its sub-op count in c_off, so the skip over
    // the sub-ops needs nothing else
    int nops = pp.nops;
    for (int link = 0;;) {
    int4 hdr = op_header(&s_ops[0]);
    for (int o = 0; o < nops;) {
      const int cur = o;
      const MicroOp& op = s_ops[cur];
      const int op_type = hdr.x, op_pred = hdr.y, op_flags = hdr.z, hdr_coff = hdr.w;
      const double2* cs = s_consts + hdr.w;
      o = cur + 1 + ((op_type == OP_GROUP || op_type == OP_KGROUP) ? hdr.w : 0);
      if (o < nops) hdr = op_header(&s_ops[o]);
      if (op_pred >= 0 && !((cw >> op_pred) & 1ull)) continue;
      if (op_type == OP_CLASSICAL) {
        if (blockDim.x <= 64u) {
          // two warps: EVERY thread runs the instructions itself, on its own copy of the classical word and on the
          // shared register file: all threads store identical values, and a thread only ever depends on what it
          // has stored itself -- so the block needs no barrier at all (the barrier that ended the previous
          // micro-op already ordered it behind every earlier reader of the registers it overwrites).
          if (hdr_coff & 2) __syncthreads();   // block right behind another span's block (fence only in between)
          cl_run<true>(cl_payload(op), op_flags, s_treg, cw, reinterpret_cast<const double*>(s_consts), pp.seed,
                       pp.shot_offset + shot);
          continue;
        }
        // four or more warps: warp 0 runs the block for the CTA (the chain is latency, not throughput: the other
        // warps would only burn issue slots on identical work), the classical word comes back through shared memory
        if (threadIdx.x < 32u) {
          cl_run<true>(cl_payload(op), op_flags, s_treg, cw, reinterpret_cast<const double*>(s_consts), pp.seed,
                       pp.shot_offset + shot);
          if (threadIdx.x == 0) s_cw = cw;
        }
        __syncthreads();
        cw = s_cw;
        continue;
      }
      if constexpr (FORM == 0) {   // ---- ket micro-ops
        switch (op_type) {
          case OP_MAT1:
            if ((base & op.octrl) == op.octrl) op_mat1(tile, T, op, cs);
            break;
          case OP_MAT2:
            if ((base & op.octrl) == op.octrl) op_mat2(tile, T, op, cs);
            break;
          case OP_DIAG: case OP_LADDER: {
            int run = 1;
            while (cur + run < nops && run < kMaxDiagRun &&
                   (s_ops[cur + run].type == OP_DIAG || s_ops[cur + run].type == OP_LADDER)) ++run;
            if (run == 1 && op_type == OP_DIAG) op_diag(tile, T, op, cs, base);
            else {
              op_diag_run(tile, T, &s_ops[cur], run, s_consts, base, cw, s_diag, s_lad);
              o = cur + run;
              if (o < nops) hdr = op_header(&s_ops[o]);
            }
            break;
          }
          case OP_KBLOCK:
            if (op.nb == 4) op_kblock<4>(tile, T, op, cs); else op_kblock<3>(tile, T, op, cs);
            break;
          case OP_SWAP:
            if ((base & op.octrl) == op.octrl) op_swap(tile, T, op);
            break;
          case OP_KGROUP:
            run_kgroup(tile, T, op, &s_ops[cur + 1], s_consts, base, cw, s_diag, s_lad, s_let);
            break;
          case OP_PROBS: op_probs<0>(tile, T, op, pp, base, blk, s_red); break;
          case OP_COLLAPSE: op_collapse<0>(tile, T, op, pp, s_meas); break;
          case OP_MEASURE:
            cw = op_measure<0>(tile, T, op, pp, base, shot, cw, s_red, s_part, &s_meas, mcount++, &s_draw, cs, &s_ops[cur + 1],
                               s_treg, s_consts, &s_cw);
            if (op.lctrl2 & kMeasRunsClassical) {   // the classical block rode inside: skip it
              o = cur + 2;
              if (o < nops) hdr = op_header(&s_ops[o]);
              take_cw = blockDim.x > 64u;
            }
            break;
          case OP_INJECT: op_inject<0>(tile, T, op, pp, cs, shot); break;
          case OP_PAULI_SAMPLED: op_pauli_sampled(tile, T, op, pp, shot, s_treg, cs); break;
          default: break;
        }
      } else {                     // ---- density-matrix micro-ops
        // the hot ops of the remote-gate passes inline; every stand-alone sweep behind ONE call, so that their loop
        // set-up is not hoisted into the dispatch of every micro-op
        if (op_type == OP_GROUP) run_group(tile, T, op, &s_ops[cur + 1], pp, s_consts, base, cw, &s_meas, s_treg);
        else if (op_type == OP_MEASURE) {
          cw = op_measure<1>(tile, T, op, pp, base, shot, cw, s_red, s_part, &s_meas, mcount++, &s_draw, cs, &s_ops[cur + 1],
                             s_treg, s_consts, &s_cw);
          if (op.lctrl2 & kMeasRunsClassical) {   // the classical block rode inside: skip it
            o = cur + 2;
            if (o < nops) hdr = op_header(&s_ops[o]);
            take_cw = blockDim.x > 64u;
          }
        } else dm_standalone(tile, T, op, op_type, pp, cs, base, blk, shot, s_red, &s_meas, s_treg);
      }
      __syncthreads();
      if (take_cw) { cw = s_cw; take_cw = false; }   // classical word of a block that warp 0 ran inside a measurement
    }
    // chained passes: the tile stays where it is, the next link's micro-program and constants replace the current
    // ones (two bulk copies out of L2 on the tile's mbarrier, or plain loads without TMA)
    if (PERSIST || ++link >= pp.nlinks) break;   // (persistent CTAs: the host never chains)
    __syncthreads();   // (a predicated-off or classical op leaves the loop without the closing barrier)
    {
      const uint4 L = __ldg(pp.links + link);
      nops = (int)L.y;
      if (tma) {
        if (threadIdx.x == 0) {
          // generic-proxy writes into the constant area (per-CTA tables of a whole gate) before the async-proxy copy
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          const uint32_t ob = L.y * (uint32_t)sizeof(MicroOp), cb = L.w * 16u;
          tma_expect(bar0_s, ob + cb);
          if (ob) tma_load_bytes((uint32_t)__cvta_generic_to_shared(s_ops), pp.ops_all + L.x, ob, bar0_s);
          if (cb) tma_load_bytes((uint32_t)__cvta_generic_to_shared(s_consts), pp.consts_all + L.z, cb, bar0_s);
        }
        uint32_t done = 0;
        while (!done) {
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
              "selp.u32 %0, 1, 0, p;\n\t}"
              : "=r"(done)
              : "r"(bar0_s), "r"(parity[0])
              : "memory");
        }
        parity[0] ^= 1u;
      } else {
        const uint4* src = reinterpret_cast<const uint4*>(pp.ops_all + L.x);
        uint4* dst = reinterpret_cast<uint4*>(s_ops);
        for (int i = threadIdx.x; i < nops * (int)(sizeof(MicroOp) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
        for (int i = threadIdx.x; i < (int)L.w; i += blockDim.x) s_consts[i] = __ldg(pp.consts_all + L.z + i);
        __syncthreads();
      }
    }
    }

    // classical state of the shot -> the `out` copies, by the CTA that holds the shot's tile 0 (every CTA of
    // the shot computed the same values)
    if ((pp.creg_out || pp.treg_out) && (blk & outer_mask) == 0) {
      if (pp.creg_out && threadIdx.x == 0) { pp.creg_out[shot] = cw; pp.rec_out[shot] = s_meas; }
      if (pp.treg_out) {
        double* tr = pp.treg_out + shot * (uint64_t)kTimeRegs;
        for (int i = threadIdx.x; i < pp.nregs; i += blockDim.x) tr[i] = s_treg[i];
      }
    }
    if (pp.store_back) {
      if (tma) {
        // generic-proxy writes to the tile -> visible to the async proxy, then bulk stores
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (!PERSIST && pp.tm_rank != 0) {
          if (fresh_tid() == 0) {
            int32_t c[5];
            tm_coords(pp, pp.tile0 + (uint64_t)fresh_ctaid(), c);
            tma_tensor_store(&pp.tmap, pp.tm_rank, (uint32_t)__cvta_generic_to_shared(dyn_tiles), c);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem stays valid until read
          }
        } else if constexpr (!PERSIST) {   // one tile per CTA: addressing state re-derived here (see fresh_tid)
          const uint32_t tid2 = fresh_tid();
          const uint32_t wid2 = __shfl_sync(0xffffffffu, tid2 >> 5, 0), nwarps2 = blockDim.x >> 5;
          const bool issuer2 = (tid2 & 31u) == 0u;
          const uint64_t blk2 = pp.tile0 + fresh_ctaid();
          const int c02 = pp.tile_segs[0].len;
          const uint32_t run_bytes2 = 16u << c02, nruns2 = (1u << pp.t) >> c02;
          const uint64_t base2 = deposit(blk2 & ((1ull << pp.outer_bits) - 1ull), pp.outer_segs, pp.n_outer_segs);
          double2* __restrict__ sp2 = pp.state + ((blk2 >> pp.outer_bits) << pp.P) + base2;
          const uint32_t tile_s2 = (uint32_t)__cvta_generic_to_shared(dyn_tiles);
          for (uint32_t r = wid2; r < nruns2; r += nwarps2) {
            double2* dst = sp2 + ((uint64_t)pp.run_hi[r] << c02);
            const uint32_t srcs = tile_s2 + r * run_bytes2;
            if (issuer2)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(srcs),
                           "r"(run_bytes2)
                           : "memory");
          }
          if (issuer2) {
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem stays valid until read
          }
        } else if (nbuf == 1) {   // every warp issues its share (uniform loop, leader lane issues + commits)
          for (uint32_t r = wid; r < nruns; r += nwarps) {
            double2* dst = sp + ((uint64_t)pp.run_hi[r] << c0);
            const uint32_t srcs = tile_s + r * run_bytes;
            if (issuer)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(srcs),
                           "r"(run_bytes)
                           : "memory");
          }
          if (issuer) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        } else if (threadIdx.x == 0) {   // persistent mode keeps every bulk group on thread 0
          for (uint32_t r = 0; r < nruns; ++r) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                             sp + ((uint64_t)pp.run_hi[r] << c0)),
                         "r"(tile_s + r * run_bytes), "r"(run_bytes)
                         : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        if (pp.tile_segs[0].dst == 0 && c0 >= 8) {
          const int nhs = pp.n_tile_segs - 1;
#pragma unroll 4
          for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {
            const uint64_t off = (uint64_t)(i & ((1u << c0) - 1u)) | deposit((uint64_t)i, pp.tile_segs + 1, nhs);
            sp[off] = tile[i];
          }
        } else {
#pragma unroll 4
          for (uint32_t i = threadIdx.x; i < T; i += blockDim.x)
            sp[deposit((uint64_t)i, pp.tile_segs, pp.n_tile_segs)] = tile[i];
        }
        __syncthreads();   // single buffer: the next iteration's load overwrites the tile
      }
    } else if (!tma) {
      __syncthreads();
    }
    if constexpr (!PERSIST) break;
  }
  // shared memory must stay valid until the last bulk store has read it
  if (PERSIST && tma && issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  // (fused measurements push their partials and end on a cluster barrier: nothing remote is pending here)
}

// ------------------------------------------------------------------------------------------
// measurement finish: one warp per shot sums the tile partials in a fixed order, adds the
// X-flip noise, draws the outcome from Philox and records (outcome, renormalisation) + creg bit.
struct FinishParams {
  const double* partials;
  MeasRec* rec;
  uint64_t* creg;
  uint64_t seed, shot_offset, draw;
  int64_t batch;
  uint64_t tiles_per_shot;
  double flip;   // dm only (ket applies the flip as a sampled X before the probabilities)
  int32_t forced, creg_bit, is_dm, pad;
  int8_t* olog;  // outcome log [draw][batch] or null
};

__global__ void __launch_bounds__(128) k_finish_measure(const FinishParams fp) {
  const int64_t shot = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (shot >= fp.batch) return;
  const double* ps = fp.partials + 2ull * (uint64_t)shot * fp.tiles_per_shot;
  double p0 = 0.0, p1 = 0.0;
  for (uint64_t i = lane; i < fp.tiles_per_shot; i += 32) { p0 += ps[2 * i]; p1 += ps[2 * i + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    p0 += __shfl_xor_sync(0xffffffffu, p0, o);
    p1 += __shfl_xor_sync(0xffffffffu, p1, o);
  }
  if (lane != 0) return;
  const double e = fp.flip;
  const double q0 = (1.0 - e) * p0 + e * p1, q1 = (1.0 - e) * p1 + e * p0;
  const double tot = q0 + q1;
  double u, u2;
  shot_uniforms(fp.seed, fp.shot_offset + (uint64_t)shot, fp.draw, u, u2);
  int m = fp.forced >= 0 ? fp.forced : ((u * tot < q1) ? 1 : 0);
  const double pm = m ? q1 : q0;
  MeasRec r;
  r.outcome = m;
  r.pad = 0;
  r.scale = pm > 0.0 ? (fp.is_dm ? 1.0 / pm : 1.0 / sqrt(pm)) : 0.0;
  fp.rec[shot] = r;
  if (fp.olog) fp.olog[fp.draw * (uint64_t)fp.batch + (uint64_t)shot] = (int8_t)m;
  if (fp.creg_bit >= 0) {
    const uint64_t bit = 1ull << fp.creg_bit;
    fp.creg[shot] = (fp.creg[shot] & ~bit) | ((uint64_t)m << fp.creg_bit);
  }
}

// ------------------------------------------------------------------------------------------
// classical instructions with no quantum op around them (e.g. the clock arithmetic that ends a run):
// one thread per shot, registers updated in place in global memory
struct ClassicalParams {
  const MicroOp* ops;
  const double* imm;
  uint64_t* creg;
  double* treg;
  int64_t batch;
  int32_t nops, pad;
  uint64_t seed, shot_offset;
};
__global__ void __launch_bounds__(128) k_classical(const ClassicalParams cp) {
  const int64_t shot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (shot >= cp.batch) return;
  uint64_t cw = cp.creg[shot];
  const uint64_t cw0 = cw;
  double* R = cp.treg + (uint64_t)shot * kTimeRegs;
  for (int k = 0; k < cp.nops; ++k) {
    const MicroOp& c = cp.ops[k];
    cl_run<false>(cl_payload(c), c.flags, R, cw, cp.imm, cp.seed, cp.shot_offset + (uint64_t)shot);
  }
  if (cw != cw0) cp.creg[shot] = cw;
}

// dqe_reinit: zero the sub-block of every shot that the live axes span (everything else is zero by the
// free-axis invariant), then |0..0>
struct ResetParams {
  double2* state;
  int64_t batch;
  int32_t P, live_bits, n_segs, pad;
  Seg segs[24];
};
__global__ void __launch_bounds__(256) k_reset_live(const ResetParams rp) {
  const uint64_t per = 1ull << rp.live_bits, total = (uint64_t)rp.batch << rp.live_bits;
  const double2 zero = make_double2(0.0, 0.0), one = make_double2(1.0, 0.0);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t shot = i >> rp.live_bits, d = i & (per - 1ull);
    rp.state[(shot << rp.P) + deposit(d, rp.segs, rp.n_segs)] = d == 0 ? one : zero;
  }
}

__global__ void k_init_state(double2* state, int64_t batch, int P) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < batch) state[(uint64_t)s << P] = make_double2(1.0, 0.0);
}

__global__ void k_extract_bit(const uint64_t* creg, int bit, int8_t* out, int64_t batch) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < batch) out[s] = (int8_t)((creg[s] >> bit) & 1ull);
}

struct HistParams {
  const uint64_t* creg;
  unsigned long long* bins;
  int64_t batch;
  int32_t k;
  int32_t bits[24];
};
__global__ void k_histogram(const HistParams hp) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= hp.batch) return;
  const uint64_t w = hp.creg[s];
  uint32_t idx = 0;
  for (int i = 0; i < hp.k; ++i) idx = (idx << 1) | (uint32_t)((w >> hp.bits[i]) & 1ull);
  atomicAdd(hp.bins + idx, 1ull);
}

// ------------------------------------------------------------------------------------------
// fused gate + NVLink transfer for a state vector sharded over GPUs: 2x2 gate on a GLOBAL axis.
// lo = shard holding the axis-bit-0 amplitudes, hi = the bit-1 shard (one of them is the partner's
// memory, mapped through CUDA IPC).  This rank handles indices [begin, end) of BOTH shards.
__global__ void __launch_bounds__(256) k_peer_1q(double2* __restrict__ lo, double2* __restrict__ hi, uint64_t begin,
                                                 uint64_t end, uint64_t ctrl_mask, double2 m00, double2 m01,
                                                 double2 m10, double2 m11) {
  // 4 independent 16-byte peer loads in flight per thread: NVLink latency (~2 us) needs the depth
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end; i0 += 4 * stride) {
    double2 a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t i = i0 + (uint64_t)k * stride;
      if (i < end && (i & ctrl_mask) == ctrl_mask) { a[k] = lo[i]; b[k] = hi[i]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t i = i0 + (uint64_t)k * stride;
      if (i < end && (i & ctrl_mask) == ctrl_mask) {
        lo[i] = cfma(m01, b[k], cmul(m00, a[k]));
        hi[i] = cfma(m11, b[k], cmul(m10, a[k]));
      }
    }
  }
}

// global <-> local axis swap over peer memory: lo[j with bit L = 1] <-> hi[j with bit L = 0]
__global__ void __launch_bounds__(256) k_peer_swap(double2* __restrict__ lo, double2* __restrict__ hi, uint64_t begin,
                                                   uint64_t end, int L) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t lowmask = (1ull << L) - 1ull;
  for (uint64_t j0 = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j0 < end; j0 += 4 * stride) {
    double2 a[4], b[4];
    uint64_t il[4], ih[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t j = j0 + (uint64_t)k * stride;
      ih[k] = ((j & ~lowmask) << 1) | (j & lowmask);   // bit L = 0
      il[k] = ih[k] | (1ull << L);                      // bit L = 1
      if (j < end) { a[k] = lo[il[k]]; b[k] = hi[ih[k]]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t j = j0 + (uint64_t)k * stride;
      if (j < end) { lo[il[k]] = b[k]; hi[ih[k]] = a[k]; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Multi-axis global <-> local exchange over peer memory, synchronised ON THE DEVICE.
// s rank bits (mask J of the rank id) trade places with s local bits L[0..s): new[R_J = x, L = y] = old[R_J = y, L = x].
// Rank r swaps, with every partner p that differs from it in the J bits only, its elements whose L value is x_p
// against the partner's elements whose L value is x_r -- 2^(m-s) pairs per partner, the lower rank of a pair doing
// the first half of them, the higher rank the second half (both NVLink directions busy, no staging buffer).
// No host barrier: a flag block behind each shard carries per-partner epoch counters.
//   start : block 0 stores epoch -> partner.ready_from[r]; every CTA polls its OWN ready_from[p] before it touches p
//   end   : the last CTA to finish (local atomic counter) stores epoch -> partner.done_from[r] and waits for its own
//           done_from[p] -- the kernel, hence everything queued behind it on the stream, ends only when every
//           partner has finished writing into this shard.
// Polls give up after ~4 s of spinning and raise the error word instead of hanging the GPU.
constexpr int kPeerFlagWords = 512;          // uint64 words behind the state buffer
constexpr int kPeerReady = 0, kPeerDone = 64, kPeerCounter = 128, kPeerError = 129;
struct ExchangeParams {
  double2* mine;
  double2* peer[64];        // shard base of every rank (entry `rank` = mine)
  uint64_t epoch[64];       // per partner: number of this exchange between the two ranks (both count alike)
  uint64_t flag_off;        // offset (in double2) of the flag block behind a shard
  int32_t rank, s, m, nparts;
  int32_t jbits[6];         // the rank bits, ascending
  int32_t lbits[6];         // the local bits paired with them
  int32_t lsorted[6];       // the local bits, ascending (zero insertion)
};
__device__ __forceinline__ volatile unsigned long long* peer_flags(double2* base, uint64_t off) {
  return reinterpret_cast<volatile unsigned long long*>(base + off);
}
__device__ __forceinline__ bool poll_ge(volatile unsigned long long* w, unsigned long long want, volatile unsigned long long* err) {
  const long long t0 = clock64();
  while (*w < want) {
    if (clock64() - t0 > 8000000000ll) { *err = 1ull; return false; }
    __nanosleep(64);
  }
  return true;
}
__global__ void __launch_bounds__(256) k_peer_exchange(const __grid_constant__ ExchangeParams ep) {
  volatile unsigned long long* myf = peer_flags(ep.mine, ep.flag_off);
  const int r = ep.rank, s = ep.s;
  int xr = 0;
  for (int k = 0; k < s; ++k) xr |= ((r >> ep.jbits[k]) & 1) << k;
  // ---- start handshake
  if (blockIdx.x == 0 && (int)threadIdx.x >= 1 && (int)threadIdx.x < (1 << s)) {
    const int d = threadIdx.x;
    int p = r;
    for (int k = 0; k < s; ++k) if ((d >> k) & 1) p ^= 1 << ep.jbits[k];
    __threadfence_system();
    peer_flags(ep.peer[p], ep.flag_off)[kPeerReady + r] = ep.epoch[p];
  }
  if ((int)threadIdx.x >= 1 && (int)threadIdx.x < (1 << s)) {
    const int d = threadIdx.x;
    int p = r;
    for (int k = 0; k < s; ++k) if ((d >> k) & 1) p ^= 1 << ep.jbits[k];
    poll_ge(myf + kPeerReady + p, ep.epoch[p], myf + kPeerError);
  }
  __syncthreads();
  __threadfence_system();
  // ---- the swaps: item = (partner d, pair index q)
  const uint64_t per = 1ull << (ep.m - s), half = per >> 1;
  const uint64_t items = (uint64_t)((1 << s) - 1) * half;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t it0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it0 < items; it0 += 4 * stride) {
    double2 a[4], b[4];
    double2* pa[4];
    double2* pb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t it = it0 + (uint64_t)u * stride;
      pa[u] = nullptr;
      if (it < items) {
        const int d = 1 + (int)(it / half);
        uint64_t q = it % half;
        int p = r;
        for (int k = 0; k < s; ++k) if ((d >> k) & 1) p ^= 1 << ep.jbits[k];
        if (r > p) q += half;
        const int xp = xr ^ d;
        uint64_t base = q;
        for (int k = 0; k < s; ++k) {
          const uint64_t lo = (1ull << ep.lsorted[k]) - 1ull;
          base = ((base & ~lo) << 1) | (base & lo);
        }
        uint64_t im = base, ip = base;
        for (int k = 0; k < s; ++k) {
          im |= (uint64_t)((xp >> k) & 1) << ep.lbits[k];
          ip |= (uint64_t)((xr >> k) & 1) << ep.lbits[k];
        }
        pa[u] = ep.mine + im; pb[u] = ep.peer[p] + ip;
        a[u] = *pa[u]; b[u] = *pb[u];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (pa[u]) { *pa[u] = b[u]; *pb[u] = a[u]; }
  }
  // ---- end handshake
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    const unsigned long long old = atomicAdd(const_cast<unsigned long long*>(myf + kPeerCounter), 1ull);
    s_last = old == (unsigned long long)gridDim.x - 1ull;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) { myf[kPeerCounter] = 0ull; __threadfence_system(); }
  if ((int)threadIdx.x >= 1 && (int)threadIdx.x < (1 << s)) {
    const int d = threadIdx.x;
    int p = r;
    for (int k = 0; k < s; ++k) if ((d >> k) & 1) p ^= 1 << ep.jbits[k];
    __threadfence_system();
    peer_flags(ep.peer[p], ep.flag_off)[kPeerDone + r] = ep.epoch[p];
    poll_ge(myf + kPeerDone + p, ep.epoch[p], myf + kPeerError);
  }
}

// (sum |amp|^2 over axis bit = 0, over bit = 1) of a single-shot ket: the local part of a sharded measurement
__global__ void __launch_bounds__(256) k_probs_bit(const double2* __restrict__ st, uint64_t n, int bit, double* out) {
  double a0 = 0.0, a1 = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double2 v = st[i];
    const double w = fma(v.x, v.x, v.y * v.y);
    if ((i >> bit) & 1ull) a1 += w; else a0 += w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  __shared__ double sa[8][2];
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5][0] = a0; sa[threadIdx.x >> 5][1] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) { t0 += sa[w][0]; t1 += sa[w][1]; }
    out[2 * blockIdx.x] = t0; out[2 * blockIdx.x + 1] = t1;   // per-block partials: summed on the host in block order
  }
}

// Density matrix held as a KET on 2n bits (rows | columns) and sharded over GPUs by row bits: the local part of
// Tr(P_axis rho).  A diagonal element has, for every qubit, equal row and column bit; pair k = (a[k], b[k]) local
// bit positions that must agree, fixed[k] >= 0: the partner bit is a rank bit of that value (b[k] unused).
struct DiagProbParams {
  const double2* state;
  uint64_t n;
  int32_t npairs, target;     // target: local bit whose value labels the outcome, or -1 (outcome fixed by the rank)
  int8_t a[32], b[32], fixed[32];
  double* out;
};
__global__ void __launch_bounds__(256) k_probs_diag(const DiagProbParams dp) {
  double a0 = 0.0, a1 = 0.0;
  uint64_t eqmask_a = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dp.n; i += (uint64_t)gridDim.x * blockDim.x) {
    bool diag = true;
    for (int k = 0; k < dp.npairs; ++k) {
      const uint64_t va = (i >> dp.a[k]) & 1ull;
      const uint64_t vb = dp.fixed[k] >= 0 ? (uint64_t)dp.fixed[k] : ((i >> dp.b[k]) & 1ull);
      if (va != vb) { diag = false; break; }
    }
    if (!diag) continue;
    const double w = dp.state[i].x;
    if (dp.target >= 0 && ((i >> dp.target) & 1ull)) a1 += w; else a0 += w;
  }
  (void)eqmask_a;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  __shared__ double sa[8][2];
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5][0] = a0; sa[threadIdx.x >> 5][1] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) { t0 += sa[w][0]; t1 += sa[w][1]; }
    dp.out[2 * blockIdx.x] = t0; dp.out[2 * blockIdx.x + 1] = t1;
  }
}

// reduced density matrix / fidelity of <= 5 axes.  `rest` enumerates the live bits that are
// traced over (dense index -> physical offset through segs).
struct ReadoutParams {
  const double2* state;   // shot 0
  const double2* ket;     // fidelity only
  double2* out_rdm;       // 4^k
  double* out_fid;        // batch
  int64_t batch;
  int32_t P, nmax, formalism, k;
  int32_t axes[8];        // physical column/ket bit of each listed axis (first = MSB)
  int32_t rest_bits, n_rest_segs;
  Seg rest_segs[24];
};

__device__ __forceinline__ uint64_t sel_offset(const ReadoutParams& rp, uint32_t i, int shift) {
  uint64_t off = 0;
  for (int s = 0; s < rp.k; ++s)
    off |= (uint64_t)((i >> (rp.k - 1 - s)) & 1u) << (rp.axes[s] + shift);
  return off;
}

// grid = (4^k, 1): block (i, j) of one shot
__global__ void __launch_bounds__(256) k_reduced_dm(const ReadoutParams rp, int64_t shot) {
  const uint32_t dim = 1u << rp.k;
  const uint32_t i = blockIdx.x / dim, j = blockIdx.x % dim;
  const double2* st = rp.state + ((uint64_t)shot << rp.P);
  const uint64_t nrest = 1ull << rp.rest_bits;
  double ax = 0.0, ay = 0.0;
  if (rp.formalism == 0) {
    const uint64_t oi = sel_offset(rp, i, 0), oj = sel_offset(rp, j, 0);
    for (uint64_t r = threadIdx.x; r < nrest; r += blockDim.x) {
      const uint64_t ro = deposit(r, rp.rest_segs, rp.n_rest_segs);
      const double2 a = st[oi | ro], b = st[oj | ro];
      ax += a.x * b.x + a.y * b.y;   // a * conj(b)
      ay += a.y * b.x - a.x * b.y;
    }
  } else {
    const uint64_t oi = sel_offset(rp, i, rp.nmax), oj = sel_offset(rp, j, 0);
    for (uint64_t r = threadIdx.x; r < nrest; r += blockDim.x) {
      const uint64_t ro = deposit(r, rp.rest_segs, rp.n_rest_segs);
      const double2 a = st[oi | oj | (ro << rp.nmax) | ro];
      ax += a.x; ay += a.y;
    }
  }
  __shared__ double sx[8], sy[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ax += __shfl_xor_sync(0xffffffffu, ax, o);
    ay += __shfl_xor_sync(0xffffffffu, ay, o);
  }
  if ((threadIdx.x & 31) == 0) { sx[threadIdx.x >> 5] = ax; sy[threadIdx.x >> 5] = ay; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tx = 0.0, ty = 0.0;
    for (int w = 0; w < 8; ++w) { tx += sx[w]; ty += sy[w]; }
    rp.out_rdm[blockIdx.x] = make_double2(tx, ty);
  }
}

// one block per shot: F = <phi| rho_S |phi>
__global__ void __launch_bounds__(128) k_fidelity_ket(const ReadoutParams rp) {
  const int64_t shot = blockIdx.x;
  const uint32_t dim = 1u << rp.k;
  const double2* st = rp.state + ((uint64_t)shot << rp.P);
  const uint64_t nrest = 1ull << rp.rest_bits;
  double acc = 0.0;
  if (rp.formalism == 0) {
    for (uint64_t r = threadIdx.x; r < nrest; r += blockDim.x) {
      const uint64_t ro = deposit(r, rp.rest_segs, rp.n_rest_segs);
      double2 s = make_double2(0.0, 0.0);
      for (uint32_t i = 0; i < dim; ++i) s = cfma(conj2(rp.ket[i]), st[sel_offset(rp, i, 0) | ro], s);
      acc += s.x * s.x + s.y * s.y;
    }
  } else {
    // one term conj(phi_i) rho_S[i][j] phi_j per step, threads spread over (rest, i, j) jointly: a register whose
    // live axes are all listed (rest_bits = 0, the shot batches of config 2) still keeps every lane busy
    const uint64_t nterms = nrest << (2 * rp.k);
    for (uint64_t e = threadIdx.x; e < nterms; e += blockDim.x) {
      const uint32_t j = (uint32_t)e & (dim - 1u), i = (uint32_t)(e >> rp.k) & (dim - 1u);
      const uint64_t ro = deposit(e >> (2 * rp.k), rp.rest_segs, rp.n_rest_segs);
      const uint64_t rr = (ro << rp.nmax) | ro;
      const double2 v = cmul(cmul(conj2(rp.ket[i]), st[sel_offset(rp, i, rp.nmax) | sel_offset(rp, j, 0) | rr]), rp.ket[j]);
      acc += v.x;
    }
  }
  __shared__ double sa[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sa[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) rp.out_fid[shot] = sa[0] + sa[1] + sa[2] + sa[3];
}

}  // namespace dqe
