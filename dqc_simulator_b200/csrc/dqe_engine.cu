// dqe_engine.cu -- host side of libdqcengine.so: handle, gate queue, pass planner ("gate-queue
// flusher"), launches and the extern "C" ABI declared in include/dqc_engine.h.
//
// The reference executes one QuantumProgram instruction at a time through NetSquid
// (software/dqc_control.py:899-949 -> QuantumProcessor.execute_program); here the same
// instruction stream is only *queued*, and dqe_flush() cuts the queue into passes.  A pass is one
// launch of k_tile_pass: every queued op whose target bits fit the shared-memory tile is applied
// in the same read-modify-write sweep of the state.  Measurement needs a per-shot global sum, so
// it splits a pass in two (probabilities at the end of one, collapse at the start of the next)
// with a tiny finish kernel in between; outcomes stay on the device.
#include "dqe_device.cuh"
#include "../../include/dqc_engine.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

using namespace dqe;
typedef std::complex<double> cplx;

namespace {

// Factor of an axis-local density-matrix op: the peephole composes memory noise -> gate -> gate noise
// into one 4x4 superoperator (one stand-alone sweep); the same op inside a register-resident group sweep
// runs as its light, in-place factors instead (a dense 4x4 there would cost every sub-op its registers).
struct Factor {
  int kind;            // 1 = Pauli channel (a, b, c, d of OP_PAULI1), 2 = unitary U: rho -> U rho U^dagger
  double d[8];         // kind 1: d[0..3]; kind 2: U row-major, interleaved re / im
};
constexpr int kMaxFactors = 4;

struct HostOp {
  int nf;              // valid factorisation of this (axis-local dm) op: f[0] applied first; 0 = none
  Factor f[kMaxFactors];
  int kind;            // 0 = micro-op, 1 = measurement finish (pass barrier), 2 = classical instruction (ci, imm),
                       // 3 = classical fence (later classical instructions are not hoisted over it)
  ClInstr ci;
  double imm, imm2;    // imm2: rate of a QEXP2
  MicroOp m;           // bit fields hold PHYSICAL positions until the planner localises them
  uint64_t pctrl, pctrl2;  // physical control masks
  uint64_t need;       // physical bits that must lie inside the tile
  uint64_t live_after; // physical live mask after this op
  int nconst;          // number of double2 constants (stored at m.c_off in the pool)
  // finish parameters
  double flip;
  int forced, creg_bit;
  uint64_t draw;
};

struct PlannedPass {
  PassParams pp;
  size_t op_begin;     // offset into the device op array
  size_t const_begin;  // offset of this pass's constants in the device constant array
  uint64_t grid;
  size_t smem;
  uint64_t amps;       // amplitudes touched per shot
  bool is_finish;
  size_t first_op;       // index in the host queue of the first op of the run of ops this pass was cut from
  uint64_t live_before;  // live mask in front of that op (streaming flush: where planning resumes)
  bool classical_only;   // nothing but classical instructions: runs as k_classical (one thread per shot)
  FinishParams fp;
};

thread_local std::string g_create_error;

// the last whole-program flush, kept for dqe_replay: the same program on the next batch of shots needs neither
// the host walk nor the planner again
struct CachedPlan {
  std::vector<PlannedPass> passes;
  std::vector<MicroOp> dev_ops;
  std::vector<double2> dev_consts;
  uint64_t live_axes, planned_live, draw;
  int64_t api_ops;
};

}  // namespace

struct dqe_engine {
  int formalism, nmax, P, device;
  int64_t batch;
  cudaStream_t stream;
  bool own_stream, own_state;
  double2* state;
  // classical state, double-buffered (see PassParams): creg_buf[cpar] / rec_buf[cpar] / treg_buf[tpar] are current
  uint64_t* creg_buf[2];
  MeasRec* rec_buf[2];
  double* treg_buf[2];      // allocated on the first classical instruction
  int cpar, tpar;
  int max_reg;              // highest register index ever written: passes move registers [0, max_reg] only
  int8_t* olog;             // outcome log [draw][batch]
  int64_t olog_draws;
  bool state_dirty;         // the free-axis invariant may not hold (dqe_set_state): next reinit clears everything
  double* partials;
  size_t partials_cap;
  void* scratch;            // read-out staging (creg bits, histograms, reduced matrices, fidelities): reused
  size_t scratch_cap;
  MicroOp* d_ops;
  size_t d_ops_cap;
  uint4* d_links;           // pass chains: link tables of the launches in flight (launch_passes)
  size_t d_links_cap;
  std::vector<uint4> links_host;
  int chain_waves;          // chains launch the batch in chunks of this many waves of resident CTAs (0 = one grid)
  int smem_pad;             // A/B: extra dynamic shared memory per CTA (bytes), i.e. fewer resident CTAs per SM
  int chain_max;            // longest chain (0 = no limit)
  int chain;                // option: consecutive one-tile-per-shot passes over the same tile layout run as one launch
  int64_t chained_links;    // statistics: passes that rode in a chain behind its head
  double2* d_consts;
  size_t d_consts_cap;
  uint64_t live_axes;       // axis mask
  uint64_t pin_axes;        // axes always enumerated by passes, live or not (keeps low bits contiguous)
  uint64_t planned_live;    // physical live mask at the head of the queue
  uint64_t seed, shot_offset, draw;
  int64_t shot_stride;      // handles that share one global shot numbering (ranks): reinit advances by batch * stride
  int pred;
  std::vector<HostOp> queue;
  std::vector<double2> cpool;
  std::string err;
  int cuda_failed;
  int tile_bits, min_run_bits, fuse, merge, group, tma, threads, ctas_per_sm, num_sms, persistent, cluster, cxperm, factor, tmap, prefetch;
  int stream_ops;           // streaming flush: closed passes are launched every this many queued ops (0 = off)
  size_t stream_next;
  int stream_min_bits;      // ... for registers of at least this many bits (default 22: passes of >= 64 MiB)
  int64_t stream_flushes;
  int kgroup;               // ket: block sweeps (OP_KGROUP)
  int64_t kgroups_built;
  uint64_t peer_epoch[64];  // exchanges done with each partner rank (k_peer_exchange handshake)
  cudaEvent_t xev0, xev1;   // device time of the peer exchanges
  bool xev_ok, xev_pending;
  double x_ms;
  int64_t x_count, x_bytes;
  int phase_tables;         // ket: leftover one-bit phases share diagonal tables (ladderize post-pass)
  int64_t phases_merged;
  int ladder;               // ket: controlled-phase stars (OP_LADDER) built from the queued diagonal tables at flush
  int64_t ladders_built;
  int pass_ops;             // planner: micro-op budget of a pass (<= kMaxOpsPerPass)
  int plan_cache;           // keep the plan of a flush that ran a whole program from a fresh register (dqe_replay)
  CachedPlan* cached;
  int64_t api_mark;         // stats[4] at the last reinit
  bool fresh;               // nothing has been launched since the last reinit
  int tile_auto;            // small dm registers (P <= 14): the planner picks tile and CTA size per pass (tile_bits_for)
  int catfuse;              // dm: inject -> interact -> measure-and-discard runs fused at flush (fuse_cat_measure)
  int64_t cat_fused;
  int kblock;               // ket: runs of gates inside <= 4 tile bits become one dense block on the FP64 tensor cores
  int64_t kblocks_fused;
  int64_t tmap_used, tmap_refused;
  int64_t stats[8];
  bool timing;
  cudaEvent_t ev0, ev1;
  double t_ms;
  int64_t t_passes, t_amps;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  std::vector<uint64_t> ev_amps;
};

namespace {

int fail(dqe_t* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      h->cuda_failed = 1;                                                               \
      return fail(h, DQE_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    }                                                                                   \
  } while (0)

inline uint64_t* cur_creg(dqe_t* h) { return h->creg_buf[h->cpar]; }
inline MeasRec* cur_rec(dqe_t* h) { return h->rec_buf[h->cpar]; }

inline uint64_t phys_live(const dqe_t* h, uint64_t axes) {
  return h->formalism == DQE_KET ? axes : (axes | (axes << h->nmax));
}
inline int rowbit(const dqe_t* h, int a) { return a + h->nmax; }

int check_ready(dqe_t* h) {
  if (!h) return DQE_EINVAL;
  if (h->cuda_failed) return DQE_ECUDA;
  return 0;
}
int check_live(dqe_t* h, int axis, const char* what) {
  if (axis < 0 || axis >= h->nmax) return fail(h, DQE_EINVAL, std::string(what) + ": axis out of range");
  if (!((h->live_axes >> axis) & 1ull))
    return fail(h, DQE_EINVAL, std::string(what) + ": axis " + std::to_string(axis) + " holds no qubit");
  return 0;
}

int add_consts(dqe_t* h, const cplx* c, int n) {
  int off = (int)h->cpool.size();
  for (int i = 0; i < n; ++i) h->cpool.push_back(make_double2(c[i].real(), c[i].imag()));
  return off;
}

HostOp blank(dqe_t* h, int type) {
  HostOp o;
  memset(&o, 0, sizeof(o));
  o.kind = 0;
  o.m.type = type;
  o.m.pred_bit = h->pred;
  o.live_after = phys_live(h, h->live_axes);
  return o;
}
Factor pauli_factor(double a, double b, double c, double d) {
  Factor f;
  memset(&f, 0, sizeof(f));
  f.kind = 1; f.d[0] = a; f.d[1] = b; f.d[2] = c; f.d[3] = d;
  return f;
}
Factor unitary_factor(const cplx* u) {
  Factor f;
  memset(&f, 0, sizeof(f));
  f.kind = 2;
  for (int i = 0; i < 4; ++i) { f.d[2 * i] = u[i].real(); f.d[2 * i + 1] = u[i].imag(); }
  return f;
}
// factors of `first` followed by factors of `then`, neighbours of one kind coalesced; false = no
// (short enough) factorisation
bool compose_factors(const HostOp& first, const HostOp& then, int* nf, Factor* out) {
  *nf = 0;
  if (first.nf <= 0 || then.nf <= 0) return false;
  Factor tmp[2 * kMaxFactors];
  memset(tmp, 0, sizeof(tmp));
  int n = 0;
  auto add = [&](const Factor& f) {
    if (n > 0 && tmp[n - 1].kind == f.kind) {
      Factor& p = tmp[n - 1];
      if (f.kind == 1) {   // [[a, b], [b, a]] on (e00, e11), [[c, d], [d, c]] on (e01, e10)
        const double a = f.d[0] * p.d[0] + f.d[1] * p.d[1], b = f.d[0] * p.d[1] + f.d[1] * p.d[0];
        const double c = f.d[2] * p.d[2] + f.d[3] * p.d[3], d = f.d[2] * p.d[3] + f.d[3] * p.d[2];
        p.d[0] = a; p.d[1] = b; p.d[2] = c; p.d[3] = d;
      } else {             // U = U_f . U_p
        cplx A[4], B[4], R[4];
        for (int i = 0; i < 4; ++i) { A[i] = cplx(f.d[2 * i], f.d[2 * i + 1]); B[i] = cplx(p.d[2 * i], p.d[2 * i + 1]); }
        R[0] = A[0] * B[0] + A[1] * B[2]; R[1] = A[0] * B[1] + A[1] * B[3];
        R[2] = A[2] * B[0] + A[3] * B[2]; R[3] = A[2] * B[1] + A[3] * B[3];
        p = unitary_factor(R);
      }
      return;
    }
    tmp[n++] = f;
  };
  for (int i = 0; i < first.nf; ++i) add(first.f[i]);
  for (int i = 0; i < then.nf; ++i) add(then.f[i]);
  if (n > kMaxFactors) return false;
  for (int i = 0; i < n; ++i) out[i] = tmp[i];
  *nf = n;
  return true;
}

bool try_merge(dqe_t* h, const HostOp& o, int axis);
bool try_merge_pair(dqe_t* h, const HostOp& o);
void diag_merge_at(dqe_t* h, size_t k);
void push(dqe_t* h, HostOp& o) {
  // axis-local candidates: ket MAT1 / 1-bit DIAG, dm MAT2 / PAULI1 / 2-bit DIAG on (row, col) of one axis
  int axis = -1;
  if (o.kind == 0 && !o.pctrl && !o.pctrl2) {
    if (h->formalism == DQE_KET) {
      if ((o.m.type == OP_MAT1 || o.m.type == OP_DIAG) && o.m.nb == 1) axis = o.m.b[0];
    } else if ((o.m.type == OP_MAT2 || o.m.type == OP_PAULI1 || o.m.type == OP_DIAG || o.m.type == OP_XPAT) && o.m.nb == 2 &&
               o.m.b[0] == o.m.b[1] + h->nmax) {
      axis = o.m.b[1];
    }
  }
  if (axis >= 0 && try_merge(h, o, axis)) return;
  if (h->formalism == DQE_KET && try_merge_pair(h, o)) return;
  h->queue.push_back(o);
  if (h->formalism == DQE_KET && o.kind == 0 && o.m.type == OP_DIAG) diag_merge_at(h, h->queue.size() - 1);
}

// -------- lowering helpers (physical bits) --------------------------------------------------
void q_mat1(dqe_t* h, int bit, const cplx* m, uint64_t pctrl) {
  HostOp o = blank(h, OP_MAT1);
  o.m.b[0] = (uint8_t)bit; o.m.nb = 1;
  o.pctrl = pctrl;
  o.need = 1ull << bit;
  o.m.c_off = add_consts(h, m, 4); o.nconst = 4;
  push(h, o);
}
void q_mat1x2(dqe_t* h, int rbit, int cbit, const cplx* m, uint64_t pctrl_r, uint64_t pctrl_c) {
  HostOp o = blank(h, OP_MAT1X2);
  o.m.b[0] = (uint8_t)rbit; o.m.b[1] = (uint8_t)cbit; o.m.nb = 2;
  o.pctrl = pctrl_r; o.pctrl2 = pctrl_c;
  o.need = (1ull << rbit) | (1ull << cbit);
  o.m.c_off = add_consts(h, m, 4); o.nconst = 4;
  push(h, o);
}
void q_mat2(dqe_t* h, int bh, int bl, const cplx* m, uint64_t pctrl, const Factor* f = nullptr) {
  HostOp o = blank(h, OP_MAT2);
  if (f) { o.nf = 1; o.f[0] = *f; }
  o.m.b[0] = (uint8_t)bh; o.m.b[1] = (uint8_t)bl; o.m.nb = 2;
  o.pctrl = pctrl;
  o.need = (1ull << bh) | (1ull << bl);
  o.m.c_off = add_consts(h, m, 16); o.nconst = 16;
  push(h, o);
}
void q_diag(dqe_t* h, int nb, const int* bits, const cplx* table, const Factor* f = nullptr) {
  HostOp o = blank(h, OP_DIAG);
  if (f) { o.nf = 1; o.f[0] = *f; }
  o.m.nb = nb;
  for (int i = 0; i < nb; ++i) o.m.b[i] = (uint8_t)bits[i];
  o.need = 0;
  o.m.c_off = add_consts(h, table, 1 << nb); o.nconst = 1 << nb;
  push(h, o);
}
void q_pauli1(dqe_t* h, int axis, double pI, double pX, double pY, double pZ) {
  HostOp o = blank(h, OP_PAULI1);
  o.m.b[0] = (uint8_t)rowbit(h, axis); o.m.b[1] = (uint8_t)axis; o.m.nb = 2;
  o.need = (1ull << rowbit(h, axis)) | (1ull << axis);
  o.m.p[0] = pI + pZ; o.m.p[1] = pX + pY; o.m.p[2] = pI - pZ; o.m.p[3] = pX - pY;
  o.nf = 1; o.f[0] = pauli_factor(o.m.p[0], o.m.p[1], o.m.p[2], o.m.p[3]);
  push(h, o);
}

bool is_zero(cplx z) { return z.real() == 0.0 && z.imag() == 0.0; }
bool is_one(cplx z) { return z.real() == 1.0 && z.imag() == 0.0; }


// -------- peephole: compose ops local to one axis ---------------------------------------------
// dm: a 1q gate, Kraus/Pauli/depolarising channel or 1q diagonal is a 4x4 superoperator on the
// (row, col) bits of its axis; ket: a 2x2 matrix on the axis bit.  A new axis-local op is folded
// into the latest queued axis-local op of the same axis and predicate when everything queued in
// between commutes with it (touches none of the axis' bits): memory noise -> gate -> gate noise
// becomes ONE micro-op instead of three.
uint64_t touched_bits(const HostOp& o) {
  uint64_t t = o.need | o.pctrl | o.pctrl2;
  if (o.m.type == OP_DIAG) for (uint32_t i = 0; i < o.m.nb; ++i) t |= 1ull << o.m.b[i];
  if (o.m.type == OP_LADDER) {
    t |= 1ull << o.m.b[0];
    const uint8_t* lf = reinterpret_cast<const uint8_t*>(&o.m) + kLadderLeafOff;
    for (int j = 0; j < o.m.aux; ++j) t |= 1ull << lf[j];
  }
  return t;
}

// 4x4 (dm) / 2x2 (ket) matrix of a queued axis-local op, or false
bool axis_local_matrix(const dqe_t* h, const HostOp& o, int axis, cplx* M) {
  const int d = h->formalism == DQE_KET ? 2 : 4;
  for (int i = 0; i < d * d; ++i) M[i] = 0.0;
  if (o.kind != 0 || o.pctrl || o.pctrl2) return false;
  if (h->formalism == DQE_KET) {
    if (o.m.type == OP_MAT1 && o.m.b[0] == axis) {
      for (int i = 0; i < 4; ++i) M[i] = cplx(h->cpool[o.m.c_off + i].x, h->cpool[o.m.c_off + i].y);
      return true;
    }
    if (o.m.type == OP_DIAG && o.m.nb == 1 && o.m.b[0] == axis) {
      M[0] = cplx(h->cpool[o.m.c_off].x, h->cpool[o.m.c_off].y);
      M[3] = cplx(h->cpool[o.m.c_off + 1].x, h->cpool[o.m.c_off + 1].y);
      return true;
    }
    return false;
  }
  const int rb = axis + h->nmax;
  if (o.m.type == OP_MAT2 && o.m.b[0] == rb && o.m.b[1] == axis) {
    for (int i = 0; i < 16; ++i) M[i] = cplx(h->cpool[o.m.c_off + i].x, h->cpool[o.m.c_off + i].y);
    return true;
  }
  if (o.m.type == OP_PAULI1 && o.m.b[0] == rb && o.m.b[1] == axis) {
    const double a = o.m.p[0], b = o.m.p[1], c = o.m.p[2], dd = o.m.p[3];
    M[0] = a; M[3] = b; M[5] = c; M[6] = dd; M[9] = dd; M[10] = c; M[12] = b; M[15] = a;
    return true;
  }
  if (o.m.type == OP_DIAG && o.m.nb == 2 && o.m.b[0] == rb && o.m.b[1] == axis) {
    for (int i = 0; i < 4; ++i) M[i * 5] = cplx(h->cpool[o.m.c_off + i].x, h->cpool[o.m.c_off + i].y);
    return true;
  }
  if (o.m.type == OP_XPAT && o.m.b[0] == rb && o.m.b[1] == axis) {
    static const int pos[8] = {0, 3, 12, 15, 5, 6, 9, 10};   // a b c d f g h k
    for (int i = 0; i < 8; ++i) M[pos[i]] = cplx(h->cpool[o.m.c_off + i].x, h->cpool[o.m.c_off + i].y);
    return true;
  }
  return false;
}

// rewrite op `o` (kept in place) as the cheapest micro-op that applies matrix M on `axis`
void set_axis_local(dqe_t* h, HostOp& o, int axis, const cplx* M) {
  const int pred = o.m.pred_bit;
  const uint64_t live_after = o.live_after;
  memset(&o.m, 0, sizeof(o.m));
  o.pctrl = o.pctrl2 = 0;
  o.m.pred_bit = pred;
  o.live_after = live_after;
  if (h->formalism == DQE_KET) {
    if (is_zero(M[1]) && is_zero(M[2])) {
      const cplx d[2] = {M[0], M[3]};
      o.m.type = OP_DIAG; o.m.nb = 1; o.m.b[0] = (uint8_t)axis; o.need = 0;
      o.m.c_off = add_consts(h, d, 2); o.nconst = 2;
    } else {
      o.m.type = OP_MAT1; o.m.nb = 1; o.m.b[0] = (uint8_t)axis; o.need = 1ull << axis;
      o.m.c_off = add_consts(h, M, 4); o.nconst = 4;
    }
    return;
  }
  const int rb = axis + h->nmax;
  o.m.b[0] = (uint8_t)rb; o.m.b[1] = (uint8_t)axis; o.m.nb = 2;
  bool diag = true, pauli = true, xpat = true;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const cplx v = M[i * 4 + j];
      if (i != j && !is_zero(v)) diag = false;
      const bool in_pat = (i == j) || (i + j == 3);
      if (!in_pat && !is_zero(v)) { pauli = false; xpat = false; }
      if (in_pat && v.imag() != 0.0) pauli = false;
    }
  if (pauli) pauli = M[0] == M[15] && M[3] == M[12] && M[5] == M[10] && M[6] == M[9];
  if (diag) {
    const cplx d[4] = {M[0], M[5], M[10], M[15]};
    o.m.type = OP_DIAG; o.need = 0;
    o.m.c_off = add_consts(h, d, 4); o.nconst = 4;
  } else if (pauli) {
    o.m.type = OP_PAULI1; o.need = (1ull << rb) | (1ull << axis);
    o.m.p[0] = M[0].real(); o.m.p[1] = M[3].real(); o.m.p[2] = M[5].real(); o.m.p[3] = M[6].real();
  } else if (xpat) {
    const cplx c8[8] = {M[0], M[3], M[12], M[15], M[5], M[6], M[9], M[10]};
    o.m.type = OP_XPAT; o.need = (1ull << rb) | (1ull << axis);
    o.m.c_off = add_consts(h, c8, 8); o.nconst = 8;
  } else {
    o.m.type = OP_MAT2; o.need = (1ull << rb) | (1ull << axis);
    o.m.c_off = add_consts(h, M, 16); o.nconst = 16;
  }
}

// try to fold the axis-local op `o` (about to be queued) into an earlier one; true = absorbed
bool try_merge(dqe_t* h, const HostOp& o, int axis) {
  if (!h->merge) return false;
  const int d = h->formalism == DQE_KET ? 2 : 4;
  cplx Mn[16], Mo[16], R[16];
  if (!axis_local_matrix(h, o, axis, Mn)) return false;
  const uint64_t mine = h->formalism == DQE_KET ? (1ull << axis) : ((1ull << axis) | (1ull << (axis + h->nmax)));
  int depth = 0;
  for (size_t k = h->queue.size(); k-- > 0 && depth < 96; ++depth) {
    HostOp& e = h->queue[k];
    if (e.kind >= 2) continue;   // classical instructions touch no amplitude
    if (e.kind != 0 || e.m.type == OP_PROBS) return false;
    if (e.m.pred_bit == o.m.pred_bit && axis_local_matrix(h, e, axis, Mo)) {
      for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
          cplx acc = 0.0;
          for (int l = 0; l < d; ++l) acc += Mn[i * d + l] * Mo[l * d + j];
          R[i * d + j] = acc;
        }
      int nf = 0;
      Factor fs[kMaxFactors];
      const bool fact = h->formalism == DQE_DM && compose_factors(e, o, &nf, fs);
      set_axis_local(h, e, axis, R);
      e.nf = fact ? nf : 0;
      for (int i = 0; i < e.nf; ++i) e.f[i] = fs[i];
      h->stats[7]++;
      return true;
    }
    if (touched_bits(e) & mine) return false;
  }
  return false;
}


// -------- ket peephole on qubit PAIRS: fused diagonal / controlled-phase runs ----------------------
// The reference decomposes cp(t) a,b into p . CX . p . CX . p (qlib/macros4parsing.py:386-463), which
// is a diagonal two-qubit gate.  Ops whose axes lie inside one pair {a, b} are composed into a 4x4
// and re-emitted in the cheapest form (diagonal table / controlled 2x2 / dense); diagonal tables on
// different bit sets are then merged up to 4 bits -- a whole run of controlled phases on one target
// becomes ONE table look-up sweep that needs no tile bits at all.
int op_axes(const HostOp& o, int* ax) {   // ket ops expressible on <= 2 axes; -1 = not mergeable
  if (o.kind != 0 || o.pctrl2) return -1;
  int n = 0;
  auto add = [&](int a) { for (int i = 0; i < n; ++i) if (ax[i] == a) return; if (n < 3) ax[n] = a; ++n; };
  if (o.m.type == OP_MAT1) {
    if (__builtin_popcountll(o.pctrl) > 1) return -1;
    add(o.m.b[0]);
    if (o.pctrl) add(__builtin_ctzll(o.pctrl));
  } else if (o.m.type == OP_MAT2) {
    if (o.pctrl) return -1;
    add(o.m.b[0]); add(o.m.b[1]);
  } else if (o.m.type == OP_DIAG) {
    if (o.m.nb > 2) return -1;
    for (uint32_t i = 0; i < o.m.nb; ++i) add(o.m.b[i]);
  } else return -1;
  return n <= 2 ? n : -1;
}

// 4x4 of a queued ket op on the ordered pair (a, b), a = MSB
bool pair_matrix(const dqe_t* h, const HostOp& o, int a, int b, cplx* M) {
  for (int i = 0; i < 16; ++i) M[i] = 0.0;
  auto cst = [&](int i) { return cplx(h->cpool[o.m.c_off + i].x, h->cpool[o.m.c_off + i].y); };
  auto idx = [&](int va, int vb) { return va * 2 + vb; };
  if (o.m.type == OP_MAT1) {
    const int t = o.m.b[0];
    const int c = o.pctrl ? __builtin_ctzll(o.pctrl) : -1;
    if (t != a && t != b) return false;
    if (c >= 0 && c != (t == a ? b : a)) return false;
    for (int vo = 0; vo < 2; ++vo)          // value of the other axis
      for (int r = 0; r < 2; ++r)
        for (int q = 0; q < 2; ++q) {
          cplx u = (c >= 0 && vo == 0) ? cplx(r == q ? 1.0 : 0.0) : cst(r * 2 + q);
          const int row = t == a ? idx(r, vo) : idx(vo, r), col = t == a ? idx(q, vo) : idx(vo, q);
          M[row * 4 + col] = u;
        }
    return true;
  }
  if (o.m.type == OP_MAT2) {
    const int x = o.m.b[0], y = o.m.b[1];
    static const int sw[4] = {0, 2, 1, 3};
    if (x == a && y == b) { for (int i = 0; i < 16; ++i) M[i] = cst(i); return true; }
    if (x == b && y == a) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M[i * 4 + j] = cst(sw[i] * 4 + sw[j]); return true; }
    return false;
  }
  if (o.m.type == OP_DIAG) {
    for (int va = 0; va < 2; ++va)
      for (int vb = 0; vb < 2; ++vb) {
        int t = 0;
        for (uint32_t i = 0; i < o.m.nb; ++i) {
          const int bit = o.m.b[i];
          if (bit != a && bit != b) return false;
          t |= (bit == a ? va : vb) << (o.m.nb - 1 - i);
        }
        M[idx(va, vb) * 5] = cst(t);
      }
    return true;
  }
  return false;
}

// rewrite `o` in place as the cheapest ket micro-op applying M on (a, b)
void set_pair(dqe_t* h, HostOp& o, int a, int b, const cplx* M) {
  const int pred = o.m.pred_bit;
  const uint64_t live_after = o.live_after;
  memset(&o.m, 0, sizeof(o.m));
  o.pctrl = o.pctrl2 = 0;
  o.m.pred_bit = pred;
  o.live_after = live_after;
  bool diag = true;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) if (i != j && !is_zero(M[i * 4 + j])) diag = false;
  if (diag) {
    const cplx d[4] = {M[0], M[5], M[10], M[15]};
    o.m.type = OP_DIAG; o.m.nb = 2; o.m.b[0] = (uint8_t)a; o.m.b[1] = (uint8_t)b; o.need = 0;
    o.m.c_off = add_consts(h, d, 4); o.nconst = 4;
    return;
  }
  // controlled on a: rows/cols with a = 0 are the identity and do not mix with a = 1
  bool ca = true, cb = true;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const cplx v = M[i * 4 + j];
      const int ia = i >> 1, ja = j >> 1, ib = i & 1, jb = j & 1;
      if (ia == 0 || ja == 0) { if (i == j ? !is_one(v) : !is_zero(v)) ca = false; }
      if (ib == 0 || jb == 0) { if (i == j ? !is_one(v) : !is_zero(v)) cb = false; }
    }
  if (ca) {
    const cplx u[4] = {M[10], M[11], M[14], M[15]};
    o.m.type = OP_MAT1; o.m.nb = 1; o.m.b[0] = (uint8_t)b; o.pctrl = 1ull << a; o.need = 1ull << b;
    o.m.c_off = add_consts(h, u, 4); o.nconst = 4;
  } else if (cb) {
    const cplx u[4] = {M[5], M[7], M[13], M[15]};
    o.m.type = OP_MAT1; o.m.nb = 1; o.m.b[0] = (uint8_t)a; o.pctrl = 1ull << b; o.need = 1ull << a;
    o.m.c_off = add_consts(h, u, 4); o.nconst = 4;
  } else {
    o.m.type = OP_MAT2; o.m.nb = 2; o.m.b[0] = (uint8_t)a; o.m.b[1] = (uint8_t)b;
    o.need = (1ull << a) | (1ull << b);
    o.m.c_off = add_consts(h, M, 16); o.nconst = 16;
  }
}

bool try_merge_pair(dqe_t* h, const HostOp& o) {
  if (!h->merge) return false;
  int ax[3], ex[3];
  const int n = op_axes(o, ax);
  if (n < 1) return false;
  uint64_t mine = 0;
  for (int i = 0; i < n; ++i) mine |= 1ull << ax[i];
  int depth = 0;
  for (size_t k = h->queue.size(); k-- > 0 && depth < 96; ++depth) {
    HostOp& e = h->queue[k];
    if (e.kind >= 2) continue;
    if (e.kind != 0 || e.m.type == OP_PROBS) return false;
    if (!(touched_bits(e) & mine)) continue;
    if (e.m.pred_bit != o.m.pred_bit) return false;
    const int m = op_axes(e, ex);
    if (m < 1) return false;
    int pr[4], np = 0;
    for (int i = 0; i < m; ++i) pr[np++] = ex[i];
    for (int i = 0; i < n; ++i) { bool seen = false; for (int j = 0; j < np; ++j) seen |= pr[j] == ax[i]; if (!seen) pr[np++] = ax[i]; }
    if (np != 2) return false;   // np == 1 is the axis-local merge's job
    cplx Mn[16], Mo[16], R[16];
    if (!pair_matrix(h, o, pr[0], pr[1], Mn) || !pair_matrix(h, e, pr[0], pr[1], Mo)) return false;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        cplx acc = 0.0;
        for (int l = 0; l < 4; ++l) acc += Mn[i * 4 + l] * Mo[l * 4 + j];
        R[i * 4 + j] = acc;
      }
    set_pair(h, e, pr[0], pr[1], R);
    h->stats[7]++;
    if (e.m.type == OP_DIAG) diag_merge_at(h, k);
    return true;
  }
  return false;
}

// queue[k] is a ket DIAG: fold it into an earlier DIAG when the union of their bits is <= 4 and
// nothing in between mixes (targets) any of its bits -- diagonals commute with diagonals and controls
void diag_merge_at(dqe_t* h, size_t k) {
  if (!h->merge) return;
  HostOp& d = h->queue[k];
  uint64_t mine = 0;
  for (uint32_t i = 0; i < d.m.nb; ++i) mine |= 1ull << d.m.b[i];
  int depth = 0;
  for (size_t j = k; j-- > 0 && depth < 96; ++depth) {
    HostOp& e = h->queue[j];
    if (e.kind >= 2) continue;
    if (e.kind != 0 || e.m.type == OP_PROBS) return;
    if (e.m.type == OP_DIAG && e.m.pred_bit == d.m.pred_bit) {
      uint8_t bits[8];
      int nb = 0;
      for (uint32_t i = 0; i < e.m.nb; ++i) bits[nb++] = e.m.b[i];
      for (uint32_t i = 0; i < d.m.nb; ++i) { bool seen = false; for (int q = 0; q < nb; ++q) seen |= bits[q] == d.m.b[i]; if (!seen) bits[nb++] = d.m.b[i]; }
      if (nb <= 4) {
        cplx table[16];
        for (int t = 0; t < (1 << nb); ++t) {
          auto sub = [&](const HostOp& x) {
            int idx = 0;
            for (uint32_t i = 0; i < x.m.nb; ++i) {
              int pos = 0;
              while (bits[pos] != x.m.b[i]) ++pos;
              idx |= ((t >> (nb - 1 - pos)) & 1) << (x.m.nb - 1 - i);
            }
            return cplx(h->cpool[x.m.c_off + idx].x, h->cpool[x.m.c_off + idx].y);
          };
          table[t] = sub(e) * sub(d);
        }
        e.m.nb = nb;
        for (int i = 0; i < nb; ++i) e.m.b[i] = bits[i];
        e.m.c_off = add_consts(h, table, 1 << nb); e.nconst = 1 << nb;
        h->queue.erase(h->queue.begin() + k);
        h->stats[7]++;
        return;
      }
    }
    if (e.need & mine) return;   // something mixes one of my bits: cannot move past it
    if (e.m.type == OP_PAULI_SAMPLED || e.m.type == OP_COLLAPSE || e.m.type == OP_INJECT) {
      if (touched_bits(e) & mine) return;
    }
  }
}

// ket-or-dm diagonal gate on k axes (first = MSB); dm table = d[r] conj(d[c]) over (rows, cols)
int lower_diag(dqe_t* h, int k, const int* axes, const cplx* d) {
  if (h->formalism == DQE_KET) {
    if (k > 4) return fail(h, DQE_EUNSUPPORTED, "diag: k <= 4 (ket)");
    q_diag(h, k, axes, d);
    return 0;
  }
  if (k > 2) return fail(h, DQE_EUNSUPPORTED, "diag: k <= 2 (dm)");
  int bits[4];
  cplx table[16];
  for (int i = 0; i < k; ++i) { bits[i] = rowbit(h, axes[i]); bits[k + i] = axes[i]; }
  const int n = 1 << k;
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) table[r * n + c] = d[r] * std::conj(d[c]);
  if (k == 1) {   // D rho D^dagger on one axis: a (diagonal) unitary factor for the group sweeps
    const cplx u[4] = {d[0], 0.0, 0.0, d[1]};
    const Factor f = unitary_factor(u);
    q_diag(h, 2, bits, table, &f);
  } else {
    q_diag(h, 2 * k, bits, table);
  }
  return 0;
}

int lower_1q(dqe_t* h, int axis, const cplx* m) {
  if (is_zero(m[1]) && is_zero(m[2])) {
    const cplx d[2] = {m[0], m[3]};
    return lower_diag(h, 1, &axis, d);
  }
  if (h->formalism == DQE_KET) {
    q_mat1(h, axis, m, 0);
  } else {
    cplx S[16];  // S[(r,c),(r',c')] = U[r,r'] conj(U[c,c'])
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 2; ++c)
        for (int r2 = 0; r2 < 2; ++r2)
          for (int c2 = 0; c2 < 2; ++c2)
            S[(r * 2 + c) * 4 + (r2 * 2 + c2)] = m[r * 2 + r2] * std::conj(m[c * 2 + c2]);
    const Factor f = unitary_factor(m);
    q_mat2(h, rowbit(h, axis), axis, S, 0, &f);
  }
  return 0;
}

int lower_ctrl_1q(dqe_t* h, int c, int t, const cplx* m) {
  if (is_zero(m[1]) && is_zero(m[2])) {
    const int axes[2] = {c, t};
    const cplx d[4] = {1.0, 1.0, m[0], m[3]};
    return lower_diag(h, 2, axes, d);
  }
  if (h->formalism == DQE_KET) q_mat1(h, t, m, 1ull << c);
  else q_mat1x2(h, rowbit(h, t), t, m, 1ull << rowbit(h, c), 1ull << c);
  return 0;
}

int lower_2q(dqe_t* h, int a, int b, const cplx* m) {
  bool diag = true, ctrl = true;
  for (int i = 0; i < 4 && (diag || ctrl); ++i)
    for (int j = 0; j < 4; ++j) {
      if (i != j && !is_zero(m[i * 4 + j])) diag = false;
      if ((i < 2 || j < 2) && !(i >= 2 && j >= 2)) {
        if (i == j ? !is_one(m[i * 4 + j]) : !is_zero(m[i * 4 + j])) ctrl = false;
      }
    }
  if (diag) {
    const int axes[2] = {a, b};
    const cplx d[4] = {m[0], m[5], m[10], m[15]};
    return lower_diag(h, 2, axes, d);
  }
  if (ctrl) {
    const cplx u[4] = {m[10], m[11], m[14], m[15]};
    return lower_ctrl_1q(h, a, b, u);
  }
  if (h->formalism == DQE_KET) {
    q_mat2(h, a, b, m, 0);
  } else {
    cplx mc[16];
    for (int i = 0; i < 16; ++i) mc[i] = std::conj(m[i]);
    q_mat2(h, rowbit(h, a), rowbit(h, b), m, 0);
    q_mat2(h, a, b, mc, 0);
  }
  return 0;
}

void q_sampled(dqe_t* h, int k, const int* axes, int mode, uint64_t draw, double p0, double p1, double p2) {
  HostOp o = blank(h, OP_PAULI_SAMPLED);
  o.m.nb = k;
  for (int i = 0; i < k; ++i) { o.m.b[i] = (uint8_t)axes[i]; o.need |= 1ull << axes[i]; }
  o.m.flags = mode;
  o.m.aux = (int32_t)draw;
  o.m.p[0] = p0; o.m.p[1] = p1; o.m.p[2] = p2;
  push(h, o);
}

int lower_measure(dqe_t* h, int axis, int creg_bit, int forced, int discard, double e) {
  if (h->olog_draws > 0 && (int64_t)h->draw >= h->olog_draws)
    return fail(h, DQE_EINVAL, "outcome log is full: raise the outcome_log option");
  const uint64_t draw = h->draw++;
  if (h->formalism == DQE_KET && e > 0.0) q_sampled(h, 1, &axis, 2, draw, e, 0, 0);
  {
    HostOp o = blank(h, OP_PROBS);
    o.m.pred_bit = -1;
    o.m.aux = axis;
    push(h, o);
  }
  {
    HostOp f;
    memset(&f, 0, sizeof(f));
    f.kind = 1;
    f.flip = h->formalism == DQE_DM ? e : 0.0;
    f.forced = forced; f.creg_bit = creg_bit; f.draw = draw;
    f.live_after = phys_live(h, h->live_axes);
    h->queue.push_back(f);
  }
  if (discard) h->live_axes &= ~(1ull << axis);
  HostOp o = blank(h, OP_COLLAPSE);   // live_after already reflects a discard
  o.m.pred_bit = -1;
  o.m.flags = discard ? 1 : 0;
  o.m.p[0] = h->formalism == DQE_DM ? e : 0.0;
  if (h->formalism == DQE_KET) { o.m.b[0] = (uint8_t)axis; o.m.nb = 1; o.need = 1ull << axis; }
  else {
    o.m.b[0] = (uint8_t)rowbit(h, axis); o.m.b[1] = (uint8_t)axis; o.m.nb = 2;
    o.need = (1ull << rowbit(h, axis)) | (1ull << axis);
  }
  push(h, o);
  return 0;
}

// -------- ket: controlled-phase stars (OP_LADDER), built at flush ---------------------------------
// Every product of <= 2-bit diagonal gates is a quadratic phase form
//     phase(i) = g . prod_a u_a^(b_a) . prod_(a<b) w_ab^(b_a b_b).
// The queued diagonal tables (already merged up to 4 bits by the push-time peepholes) are taken apart into
// these terms: g and the one-bit phases u_a fold into a neighbouring one-qubit matrix of their axis (the H of
// the QFT stage), and the pair terms that share a centre bit q collect in ONE star op -- the whole
// controlled-phase ladder of a QFT stage (qlib/macros4parsing.py:386-463, one cp per lower qubit).  On the
// device a star costs about two complex multiplies per amplitude whatever its length; the <= 4-bit tables
// it replaces cost one look-up + multiply per three controls.  Stars that stay short go back to tables.
struct QuadForm { cplx g, u[4], w[4][4]; };
bool quad_form(const dqe_t* h, const HostOp& o, QuadForm* q) {
  const int nb = (int)o.m.nb;
  if (nb < 1 || nb > 4) return false;
  auto T = [&](int t) { return cplx(h->cpool[o.m.c_off + t].x, h->cpool[o.m.c_off + t].y); };
  auto e = [&](int i) { return 1 << (nb - 1 - i); };   // table index of bit b[i] (b[0] = MSB)
  q->g = T(0);
  if (std::abs(q->g) == 0.0) return false;
  for (int i = 0; i < nb; ++i) {
    if (std::abs(T(e(i))) == 0.0) return false;
    q->u[i] = T(e(i)) / q->g;
  }
  for (int i = 0; i < nb; ++i)
    for (int j = i + 1; j < nb; ++j) q->w[i][j] = T(e(i) | e(j)) * q->g / (T(e(i)) * T(e(j)));
  for (int t = 0; t < (1 << nb); ++t) {
    cplx v = q->g;
    for (int i = 0; i < nb; ++i) if (t & e(i)) v *= q->u[i];
    for (int i = 0; i < nb; ++i)
      for (int j = i + 1; j < nb; ++j) if ((t & e(i)) && (t & e(j))) v *= q->w[i][j];
    if (std::abs(v - T(t)) > 1e-14 * std::max(1.0, std::abs(T(t)))) return false;
  }
  return true;
}
inline uint8_t* leaves_of(HostOp& o) { return reinterpret_cast<uint8_t*>(&o.m) + kLadderLeafOff; }

// does `e` stop a diagonal term on bits `mine` from moving in front of it?
bool blocks_diagonal(const HostOp& e, uint64_t mine) {
  if (e.kind != 0 || e.m.type == OP_PROBS || e.m.type == OP_MEASURE) return true;
  if (e.need & mine) return true;
  if (e.m.type == OP_PAULI_SAMPLED || e.m.type == OP_COLLAPSE || e.m.type == OP_INJECT) return (touched_bits(e) & mine) != 0;
  return false;
}

bool host_is_swap(const dqe_t* h, const HostOp& o) {
  if (o.m.type != OP_MAT2) return false;
  const double2* u = &h->cpool[o.m.c_off];
  for (int i = 0; i < 16; ++i) {
    const bool is1 = i == 0 || i == 6 || i == 9 || i == 15;
    if (u[i].y != 0.0 || u[i].x != (is1 ? 1.0 : 0.0)) return false;
  }
  return true;
}

void ladderize(dqe_t* h) {
  if (h->formalism != DQE_KET || !h->ladder || !h->merge || !h->fuse) return;
  int ndiag = 0;
  for (const HostOp& o : h->queue) ndiag += o.kind == 0 && o.m.type == OP_DIAG && o.m.nb >= 2;
  if (ndiag < 4) return;
  std::vector<HostOp> out;
  out.reserve(h->queue.size());
  const double eps = 1e-15;
  auto is_id = [&](cplx z) { return std::abs(z - 1.0) < eps; };
  const int kDepth = 256;

  // one-bit phase diag(d0, d1) on `bit`
  auto place_phase = [&](const HostOp& src, int bit, cplx d0, cplx d1) {
    if (is_id(d0) && is_id(d1)) return;
    const uint64_t mine = 1ull << bit;
    int depth = 0;
    for (size_t k = out.size(); k-- > 0 && depth < kDepth; ++depth) {
      HostOp& e = out[k];
      if (e.kind >= 2) continue;
      if (e.kind == 0 && e.m.pred_bit == src.m.pred_bit && !e.pctrl && !e.pctrl2 && e.m.b[0] == bit && e.m.nb == 1) {
        if (e.m.type == OP_MAT1) {
          cplx M[4];
          for (int i = 0; i < 4; ++i) M[i] = cplx(h->cpool[e.m.c_off + i].x, h->cpool[e.m.c_off + i].y);
          M[0] *= d0; M[1] *= d0; M[2] *= d1; M[3] *= d1;   // D . M
          e.m.c_off = add_consts(h, M, 4);
          return;
        }
        if (e.m.type == OP_DIAG) {
          const cplx d[2] = {cplx(h->cpool[e.m.c_off].x, h->cpool[e.m.c_off].y) * d0,
                             cplx(h->cpool[e.m.c_off + 1].x, h->cpool[e.m.c_off + 1].y) * d1};
          e.m.c_off = add_consts(h, d, 2);
          return;
        }
      }
      if (e.kind == 0 && e.m.type == OP_MAT2 && e.m.pred_bit == src.m.pred_bit && !e.pctrl && !e.pctrl2 &&
          (e.m.b[0] == bit || e.m.b[1] == bit) && !host_is_swap(h, e)) {   // (a SWAP stays a pure permutation)
        cplx M[16];
        for (int i = 0; i < 16; ++i) M[i] = cplx(h->cpool[e.m.c_off + i].x, h->cpool[e.m.c_off + i].y);
        const int sh = e.m.b[0] == bit ? 1 : 0;   // row index = b[0] value * 2 + b[1] value
        for (int r = 0; r < 4; ++r)
          for (int c = 0; c < 4; ++c) M[r * 4 + c] *= ((r >> sh) & 1) ? d1 : d0;   // D . M
        e.m.c_off = add_consts(h, M, 16);
        return;
      }
      if (blocks_diagonal(e, mine)) break;
    }
    HostOp o = src;
    memset(&o.m, 0, sizeof(o.m));
    o.nf = 0; o.pctrl = o.pctrl2 = 0; o.need = 0;
    o.m.type = OP_DIAG; o.m.pred_bit = src.m.pred_bit; o.m.nb = 1; o.m.b[0] = (uint8_t)bit;
    const cplx d[2] = {d0, d1};
    o.m.c_off = add_consts(h, d, 2); o.nconst = 2;
    out.push_back(o);
  };

  // pair term w^(b_a b_b)
  auto place_cp = [&](const HostOp& src, int a, int b, cplx w) {
    if (is_id(w)) return;
    const uint64_t mine = (1ull << a) | (1ull << b);
    int depth = 0;
    for (size_t k = out.size(); k-- > 0 && depth < kDepth; ++depth) {
      HostOp& e = out[k];
      if (e.kind >= 2) continue;
      if (e.kind == 0 && e.m.type == OP_LADDER && e.m.pred_bit == src.m.pred_bit) {
        uint8_t* lf = leaves_of(e);
        int n = e.m.aux;
        int q = e.m.b[0];
        if (n == 1 && q != a && q != b && (lf[0] == a || lf[0] == b)) {   // a single pair term: its centre is free
          const uint8_t t = lf[0]; lf[0] = (uint8_t)q; e.m.b[0] = t; q = t;
        }
        if (q == a || q == b) {
          const int leaf = q == a ? b : a;
          int j = 0;
          while (j < n && lf[j] != leaf) ++j;
          std::vector<cplx> ws(n + 1);
          for (int i = 0; i < n; ++i) ws[i] = cplx(h->cpool[e.m.c_off + i].x, h->cpool[e.m.c_off + i].y);
          if (j < n) ws[j] *= w;
          else if (n < kMaxLadderLeaves) { lf[n] = (uint8_t)leaf; ws[n] = w; ++n; }
          else j = -1;
          if (j >= 0) {
            e.m.aux = n;
            e.m.c_off = add_consts(h, ws.data(), n); e.nconst = n;
            return;
          }
        }
      }
      if (blocks_diagonal(e, mine)) break;
    }
    HostOp o = src;
    memset(&o.m, 0, sizeof(o.m));
    o.nf = 0; o.pctrl = o.pctrl2 = 0; o.need = 0;
    o.m.type = OP_LADDER; o.m.pred_bit = src.m.pred_bit; o.m.nb = 1; o.m.b[0] = (uint8_t)a; o.m.aux = 1;
    leaves_of(o)[0] = (uint8_t)b;
    o.m.c_off = add_consts(h, &w, 1); o.nconst = 1;
    out.push_back(o);
  };

  // a dense 4x4 that is really "diagonal, then a one-qubit matrix" or the reverse -- what the push-time pair
  // peephole makes of H(q) meeting the controlled phase next to it -- is taken apart again: the matrix costs
  // half of the dense block and needs one tile bit, the diagonal joins its star
  auto split_mat2 = [&](const HostOp& o, HostOp* first, HostOp* second) {
    cplx M[16];
    for (int i = 0; i < 16; ++i) M[i] = cplx(h->cpool[o.m.c_off + i].x, h->cpool[o.m.c_off + i].y);
    const double tol = 1e-14;
    for (int tsel = 0; tsel < 2; ++tsel) {          // target = b[tsel], the other bit only carries phases
      const int ts = tsel == 0 ? 1 : 0, os = 1 - ts;   // shifts of target / other in the 2-bit index
      bool ok = true;
      for (int r = 0; r < 4 && ok; ++r)
        for (int c = 0; c < 4; ++c)
          if (((r >> os) & 1) != ((c >> os) & 1) && std::abs(M[r * 4 + c]) > tol) { ok = false; break; }
      if (!ok) continue;
      auto B = [&](int s_, int r, int c) { return M[((r << ts) | (s_ << os)) * 4 + ((c << ts) | (s_ << os))]; };
      for (int form = 0; form < 2; ++form) {        // 0: (U (x) I) . D  (ratio depends on the column), 1: D . (U (x) I)
        cplx ratio[2];
        bool good = true;
        for (int x = 0; x < 2 && good; ++x) {
          bool have = false;
          for (int y = 0; y < 2; ++y) {
            const cplx b0 = form == 0 ? B(0, y, x) : B(0, x, y), b1 = form == 0 ? B(1, y, x) : B(1, x, y);
            if (std::abs(b0) > 1e-3) {
              if (!have) { ratio[x] = b1 / b0; have = true; }
              else if (std::abs(b1 - ratio[x] * b0) > tol) good = false;
            } else if (std::abs(b1) > tol && std::abs(b0) <= tol) good = false;
          }
          if (!have) good = false;
        }
        if (!good) continue;
        bool dense_u = std::abs(B(0, 0, 1)) > tol || std::abs(B(0, 1, 0)) > tol;
        if (!dense_u) continue;                     // plain diagonal: nothing to split
        const cplx U[4] = {B(0, 0, 0), B(0, 0, 1), B(0, 1, 0), B(0, 1, 1)};
        for (int r = 0; r < 2 && good; ++r)          // the product reproduces every entry
          for (int c = 0; c < 2; ++c)
            if (std::abs(B(1, r, c) - U[r * 2 + c] * ratio[form == 0 ? c : r]) > 1e-13) good = false;
        if (!good) continue;
        HostOp mat = o, dg = o;
        memset(&mat.m, 0, sizeof(mat.m)); memset(&dg.m, 0, sizeof(dg.m));
        mat.nf = dg.nf = 0;
        const int tb = o.m.b[tsel], ob = o.m.b[1 - tsel];
        mat.m.type = OP_MAT1; mat.m.pred_bit = o.m.pred_bit; mat.m.nb = 1; mat.m.b[0] = (uint8_t)tb;
        mat.need = 1ull << tb; mat.m.c_off = add_consts(h, U, 4); mat.nconst = 4;
        // table over (target, other): phase 1 for other = 0, ratio[target value] for other = 1
        const cplx tab[4] = {1.0, ratio[0], 1.0, ratio[1]};
        dg.m.type = OP_DIAG; dg.m.pred_bit = o.m.pred_bit; dg.m.nb = 2; dg.m.b[0] = (uint8_t)tb; dg.m.b[1] = (uint8_t)ob;
        dg.need = 0; dg.m.c_off = add_consts(h, tab, 4); dg.nconst = 4;
        if (form == 0) { *first = dg; *second = mat; } else { *first = mat; *second = dg; }
        return true;
      }
    }
    return false;
  };

  std::vector<HostOp> work;   // ops still to process, in order (a split MAT2 puts its two halves back in front)
  for (size_t qi = 0; qi < h->queue.size() || !work.empty();) {
    HostOp o;
    if (!work.empty()) { o = work.front(); work.erase(work.begin()); }
    else o = h->queue[qi++];
    QuadForm q;
    if (o.kind == 0 && o.m.type == OP_MAT2 && !o.pctrl && !o.pctrl2) {
      HostOp a, b;
      if (split_mat2(o, &a, &b)) { work.insert(work.begin(), b); work.insert(work.begin(), a); continue; }
    }
    if (o.kind == 0 && o.m.type == OP_DIAG && quad_form(h, o, &q)) {
      const int nb = (int)o.m.nb;
      for (int i = 0; i < nb; ++i) place_phase(o, o.m.b[i], i == 0 ? q.g : cplx(1.0), (i == 0 ? q.g : cplx(1.0)) * q.u[i]);
      for (int i = 0; i < nb; ++i)
        for (int j = i + 1; j < nb; ++j) place_cp(o, o.m.b[i], o.m.b[j], q.w[i][j]);
      continue;
    }
    if (o.kind == 0 && o.m.type == OP_MAT1 && !o.pctrl && !o.pctrl2) {
      // a pending one-bit phase of this axis moves forward into the matrix: M . D
      const int bit = o.m.b[0];
      const uint64_t mine = 1ull << bit;
      int depth = 0;
      HostOp m = o;
      for (size_t k = out.size(); k-- > 0 && depth < kDepth; ++depth) {
        HostOp& e = out[k];
        if (e.kind >= 2) continue;
        if (e.kind == 0 && e.m.type == OP_DIAG && e.m.nb == 1 && e.m.b[0] == bit && e.m.pred_bit == o.m.pred_bit) {
          cplx M[4];
          for (int i = 0; i < 4; ++i) M[i] = cplx(h->cpool[o.m.c_off + i].x, h->cpool[o.m.c_off + i].y);
          const cplx d0(h->cpool[e.m.c_off].x, h->cpool[e.m.c_off].y), d1(h->cpool[e.m.c_off + 1].x, h->cpool[e.m.c_off + 1].y);
          M[0] *= d0; M[2] *= d0; M[1] *= d1; M[3] *= d1;
          m.m.c_off = add_consts(h, M, 4);
          out.erase(out.begin() + k);
          break;
        }
        if (blocks_diagonal(e, mine)) break;
      }
      out.push_back(m);
      continue;
    }
    if (o.kind == 0 && o.m.type == OP_MAT2 && !o.pctrl && !o.pctrl2 && !host_is_swap(h, o)) {
      HostOp m = o;
      for (int side = 0; side < 2; ++side) {
        const int bit = o.m.b[side];
        const uint64_t mine = 1ull << bit;
        int depth = 0;
        for (size_t k = out.size(); k-- > 0 && depth < kDepth; ++depth) {
          HostOp& e = out[k];
          if (e.kind >= 2) continue;
          if (e.kind == 0 && e.m.type == OP_DIAG && e.m.nb == 1 && e.m.b[0] == bit && e.m.pred_bit == o.m.pred_bit) {
            cplx M[16];
            for (int i = 0; i < 16; ++i) M[i] = cplx(h->cpool[m.m.c_off + i].x, h->cpool[m.m.c_off + i].y);
            const cplx d0(h->cpool[e.m.c_off].x, h->cpool[e.m.c_off].y), d1(h->cpool[e.m.c_off + 1].x, h->cpool[e.m.c_off + 1].y);
            const int sh = side == 0 ? 1 : 0;
            for (int r = 0; r < 4; ++r)
              for (int c = 0; c < 4; ++c) M[r * 4 + c] *= ((c >> sh) & 1) ? d1 : d0;   // M . D
            m.m.c_off = add_consts(h, M, 16);
            out.erase(out.begin() + k);
            break;
          }
          if (blocks_diagonal(e, mine)) break;
        }
      }
      out.push_back(m);
      continue;
    }
    out.push_back(o);
  }
  // One-bit phases that found no matrix to fold into (an axis that is never mixed again: a QFT qubit behind its H only
  // controls, and collects the one-bit halves of those controlled phases) share tables: each moves forward past the ops
  // that commute with it and joins the next small unpredicated diagonal -- up to four bits per op, one multiplication
  // per amplitude for the lot instead of one each (QFT-30: 14 of the 33 micro-ops of the largest pass were such phases).
  // Option phase_tables, OFF by default: measured neutral (QFT-30 pass 3: 18.95 -> 18.98 ms with 21 instead of 33 micro-ops;
  // the block sweep's cost is not in its diagonal sub-ops).
  if (h->phase_tables) {
    std::vector<char> dead(out.size(), 0);
    auto plain_diag = [&](const HostOp& e) {
      return e.kind == 0 && e.m.type == OP_DIAG && e.m.pred_bit < 0 && !e.pctrl && !e.pctrl2 && e.nf == 0 && e.m.nb >= 1;
    };
    for (size_t i = 0; i < out.size(); ++i) {
      const HostOp a = out[i];
      if (!plain_diag(a) || a.m.nb > 3) continue;
      uint64_t mine = 0;
      for (uint32_t q = 0; q < a.m.nb; ++q) mine |= 1ull << a.m.b[q];
      const int na = 1 << a.m.nb;
      int depth = 0;
      for (size_t j = i + 1; j < out.size() && depth < kDepth; ++j, ++depth) {
        HostOp& e = out[j];
        if (dead[j] || e.kind >= 2) continue;
        if (plain_diag(e) && e.m.nb + a.m.nb <= 4) {
          bool disjoint = true;
          for (uint32_t q = 0; q < e.m.nb; ++q) disjoint = disjoint && !((mine >> e.m.b[q]) & 1ull);
          if (disjoint) {
            const int ne = 1 << e.m.nb;
            cplx t[16];
            for (int x = 0; x < ne; ++x) {
              const cplx v(h->cpool[e.m.c_off + x].x, h->cpool[e.m.c_off + x].y);
              for (int y = 0; y < na; ++y)     // a's bits become the least significant index bits
                t[x * na + y] = v * cplx(h->cpool[a.m.c_off + y].x, h->cpool[a.m.c_off + y].y);
            }
            for (uint32_t q = 0; q < a.m.nb; ++q) e.m.b[e.m.nb + q] = a.m.b[q];
            e.m.nb += a.m.nb;
            e.m.c_off = add_consts(h, t, ne * na); e.nconst = ne * na;
            dead[i] = 1;
            h->phases_merged++;
            break;
          }
        }
        if (blocks_diagonal(e, mine)) break;
      }
    }
    size_t w = 0;
    for (size_t i = 0; i < out.size(); ++i) if (!dead[i]) { if (w != i) out[w] = out[i]; ++w; }
    out.resize(w);
  }
  // short stars are cheaper as one table look-up
  for (HostOp& o : out) {
    if (o.kind != 0 || o.m.type != OP_LADDER) continue;
    const int n = o.m.aux;
    if (n > 3) { h->ladders_built++; continue; }
    uint8_t lf[4];
    cplx w[4];
    for (int j = 0; j < n; ++j) { lf[j] = leaves_of(o)[j]; w[j] = cplx(h->cpool[o.m.c_off + j].x, h->cpool[o.m.c_off + j].y); }
    const int q = o.m.b[0], pred = o.m.pred_bit;
    cplx table[16];
    for (int t = 0; t < (2 << n); ++t) {
      cplx v = 1.0;
      if (t >> n) for (int j = 0; j < n; ++j) if ((t >> (n - 1 - j)) & 1) v *= w[j];
      table[t] = v;
    }
    memset(&o.m, 0, sizeof(o.m));
    o.m.type = OP_DIAG; o.m.pred_bit = pred; o.m.nb = (uint32_t)(n + 1);
    o.m.b[0] = (uint8_t)q;
    for (int j = 0; j < n; ++j) o.m.b[1 + j] = lf[j];
    o.m.c_off = add_consts(h, table, 2 << n); o.nconst = 2 << n;
  }
  h->queue.swap(out);
}

// -------- planner --------------------------------------------------------------------------
// -------- dm: inject -> interact -> measure-and-discard, fused at flush ------------------------------------
// The remote-gate sub-circuits of the reference (software/dqc_control.py:316-403 cat-entangle, :508-613
// teleportation) all open the same way: a fresh entangled pair (x, y) lands on two comm qubits
// (hardware/connections.py:145-150), ONE half x interacts with a data qubit A (CNOT + its gate noise) and is measured
// and discarded at once.  Executed literally, the register grows by two qubits for those few ops (16x the
// elements of rho) although x never outlives the pass.  Everything between the injection and the measurement that
// touches x is a fixed superoperator S on (A, x), so the whole run is ONE linear map from the 4 elements
// rho[rA', cA'] of the input to the 16 elements rho'[rA ry, cA cy] of the output, per outcome m:
//     T_m[(rA ry cA cy), (rA' cA')] = sum_{rx' cx'} sum_k w_m(k) S[(rA k cA k), (rA' rx' cA' cx')] W[(rx' ry), (cx' cy)]
// (W = the injected pair, w_m(k) = 1 - e for k = m, e for k = 1 - m: the measurement flip noise of
// noise_models.py MeasurementNoiseModel folded in), and the outcome probabilities are the linear functionals
// t_m[(rA' cA')] = sum_{a b} T_m[(a b a b), (rA' cA')].  The queue is rewritten accordingly: the INJECT, the ops on
// (A, x) and the COLLAPSE disappear, the PROBS op becomes the fused measurement (nb = 4: bits of A and y; constants
// t_0, t_1, T_0, T_1), x never goes live -- a cat-entangle pass runs over 4x fewer elements.
// Index convention of the 16-dimensional operator space of (A, x): i = rA * 8 + rx * 4 + cA * 2 + cx.
struct Sup16 { cplx m[256]; };
void sup_identity(Sup16& S) { for (int i = 0; i < 256; ++i) S.m[i] = (i / 16 == i % 16) ? 1.0 : 0.0; }
void sup_lmul(Sup16& S, const Sup16& L) {   // S <- L S
  Sup16 R;
  for (int i = 0; i < 16; ++i)
    for (int j = 0; j < 16; ++j) {
      cplx acc = 0.0;
      for (int k = 0; k < 16; ++k) acc += L.m[i * 16 + k] * S.m[k * 16 + j];
      R.m[i * 16 + j] = acc;
    }
  S = R;
}
// role 0 = A (row bit 3, column bit 1 of the index), role 1 = x (row bit 2, column bit 0)
inline int sup_rbit(int role) { return role == 0 ? 3 : 2; }
inline int sup_cbit(int role) { return role == 0 ? 1 : 0; }
void sup_axis_local(Sup16& L, int role, const cplx* M) {   // M: 4x4 on (r, c) of the axis, index r * 2 + c
  const int rb = sup_rbit(role), cb = sup_cbit(role);
  const int mine = (1 << rb) | (1 << cb);
  for (int i = 0; i < 16; ++i)
    for (int j = 0; j < 16; ++j) {
      if ((i & ~mine) != (j & ~mine)) { L.m[i * 16 + j] = 0.0; continue; }
      const int a = ((i >> rb) & 1) * 2 + ((i >> cb) & 1), b = ((j >> rb) & 1) * 2 + ((j >> cb) & 1);
      L.m[i * 16 + j] = M[a * 4 + b];
    }
}
// U on the row bit of `role` (only where the other axis' row bit is 1 when ctrl_r), conj(U) on its column bit
void sup_mat1x2(Sup16& L, int role, const cplx* U, bool ctrl_r, bool ctrl_c) {
  Sup16 Lr, Lc;
  const int rb = sup_rbit(role), cb = sup_cbit(role), orb = sup_rbit(role ^ 1), ocb = sup_cbit(role ^ 1);
  for (int i = 0; i < 16; ++i)
    for (int j = 0; j < 16; ++j) {
      const bool same_r = (i & ~(1 << rb)) == (j & ~(1 << rb)), same_c = (i & ~(1 << cb)) == (j & ~(1 << cb));
      const int ri = (i >> rb) & 1, rj = (j >> rb) & 1, ci = (i >> cb) & 1, cj = (j >> cb) & 1;
      const bool on_r = !ctrl_r || ((i >> orb) & 1), on_c = !ctrl_c || ((i >> ocb) & 1);
      Lr.m[i * 16 + j] = !same_r ? cplx(0.0) : on_r ? U[ri * 2 + rj] : cplx(ri == rj ? 1.0 : 0.0);
      Lc.m[i * 16 + j] = !same_c ? cplx(0.0) : on_c ? std::conj(U[ci * 2 + cj]) : cplx(ci == cj ? 1.0 : 0.0);
    }
  L = Lr;
  sup_lmul(L, Lc);
}
void sup_depol2(Sup16& L, double p) {
  for (int i = 0; i < 16; ++i)
    for (int j = 0; j < 16; ++j) {
      const bool di = (i >> 2) == (i & 3), dj = (j >> 2) == (j & 3);
      L.m[i * 16 + j] = (i == j ? 1.0 - p : 0.0) + (di && dj ? 0.25 * p : 0.0);
    }
}

bool cat_cluster_mode(const dqe_t* h) { return h->cluster && !h->persistent && h->fuse && h->P - (h->tile_auto ? 10 : h->tile_bits) <= 4; }

// INJECT at queue index i created axes (x, y); x_first: x is the first axis of the injected pair.  true = rewritten.
bool try_fuse_cat_at(dqe_t* h, size_t i, int x, int y, bool x_first) {
  std::vector<HostOp>& Q = h->queue;
  const int n = h->nmax;
  const uint64_t xb = (1ull << x) | (1ull << (x + n)), yb = (1ull << y) | (1ull << (y + n));
  int A = -1;
  uint64_t Ab = 0;
  Sup16 S, L;
  sup_identity(S);
  std::vector<size_t> absorbed;
  size_t jp = Q.size();
  auto set_A = [&](int a) {
    if (a == x || a == y || a < 0 || a >= n) return false;
    if (A >= 0) return A == a;
    A = a; Ab = (1ull << a) | (1ull << (a + n));
    return true;
  };
  cplx M[16];
  bool unit_trace;
  {
    const double2* w = &h->cpool[Q[i].m.c_off];
    unit_trace = std::abs(w[0].x + w[5].x + w[10].x + w[15].x - 1.0) < 1e-12 &&
                 std::abs(w[0].y + w[5].y + w[10].y + w[15].y) < 1e-12;
  }
  for (size_t j = i + 1; j < Q.size(); ++j) {
    const HostOp& e = Q[j];
    if (e.kind != 0) continue;   // classical instructions / fences / finish records of other measurements
    if (e.m.type == OP_PROBS) {
      if (e.m.aux == x) { jp = j; break; }
      // another measurement in between: the injection may move behind it (a normalised pair changes no probability),
      // folded ops may not (they need not preserve the trace)
      if (e.m.aux == y || e.m.aux == A || e.m.nb == 4 || !absorbed.empty() || !unit_trace) return false;
      continue;
    }
    const uint64_t tb = touched_bits(e);
    if (!(tb & (xb | yb | Ab))) continue;   // unrelated: commutes with everything folded, stays where it is
    if (tb & yb) return false;
    if (e.m.pred_bit >= 0) return false;
    if (!(tb & xb)) {            // on A alone, behind the first (A, x) op: constant axis-local channels fold in
      if ((tb & ~Ab) || !axis_local_matrix(h, e, A, M)) return false;
      sup_axis_local(L, 0, M);
    } else if (!(tb & ~xb) && axis_local_matrix(h, e, x, M)) {
      sup_axis_local(L, 1, M);
    } else if (e.m.type == OP_MAT1X2) {
      const int T = e.m.b[1];
      if (e.m.b[0] != T + n) return false;
      if (__builtin_popcountll(e.pctrl) > 1 || __builtin_popcountll(e.pctrl2) > 1) return false;
      const int cr = e.pctrl ? __builtin_ctzll(e.pctrl) - n : -1, cc = e.pctrl2 ? __builtin_ctzll(e.pctrl2) : -1;
      if (cr >= 0 && cc >= 0 && cr != cc) return false;
      const int O = cr >= 0 ? cr : cc;
      if (O == T) return false;
      if (T != x && !set_A(T)) return false;
      if (O >= 0 && O != x && !set_A(O)) return false;
      if (T != x && O != x) return false;
      cplx U[4];
      for (int k = 0; k < 4; ++k) U[k] = cplx(h->cpool[e.m.c_off + k].x, h->cpool[e.m.c_off + k].y);
      sup_mat1x2(L, T == x ? 1 : 0, U, cr >= 0, cc >= 0);
    } else if (e.m.type == OP_DEPOL2) {
      const int a = e.m.b[2], b = e.m.b[3];
      if (a != x && b != x) return false;
      if (!set_A(a == x ? b : a)) return false;
      sup_depol2(L, e.m.p[0]);
    } else {
      return false;
    }
    sup_lmul(S, L);
    absorbed.push_back(j);
  }
  if (jp + 2 >= Q.size() || A < 0) return false;
  const HostOp& fin = Q[jp + 1];
  const HostOp& col = Q[jp + 2];
  if (fin.kind != 1 || col.kind != 0 || col.m.type != OP_COLLAPSE || col.m.b[1] != x || !(col.m.flags & 1)) return false;
  const double e = fin.flip;
  // W[(rx ry), (cx cy)] of the injected pair
  const HostOp& inj = Q[i];
  auto W = [&](int rx, int ry, int cx, int cy) {
    const int r = x_first ? rx * 2 + ry : ry * 2 + rx, c = x_first ? cx * 2 + cy : cy * 2 + cx;
    const double2 v = h->cpool[inj.m.c_off + r * 4 + c];
    return cplx(v.x, v.y);
  };
  cplx K[8 + 128];
  for (int m = 0; m < 2; ++m) {
    cplx* T = K + 8 + 64 * m;
    for (int o = 0; o < 16; ++o) {
      const int rA = (o >> 3) & 1, ry = (o >> 2) & 1, cA = (o >> 1) & 1, cy = o & 1;
      for (int jn = 0; jn < 4; ++jn) {
        const int rA2 = jn >> 1, cA2 = jn & 1;
        cplx acc = 0.0;
        for (int k = 0; k < 2; ++k) {
          const double w = k == m ? 1.0 - e : e;
          if (w == 0.0) continue;
          const int row = rA * 8 + k * 4 + cA * 2 + k;
          for (int rx2 = 0; rx2 < 2; ++rx2)
            for (int cx2 = 0; cx2 < 2; ++cx2)
              acc += w * S.m[row * 16 + (rA2 * 8 + rx2 * 4 + cA2 * 2 + cx2)] * W(rx2, ry, cx2, cy);
        }
        T[o * 4 + jn] = acc;
      }
    }
    for (int jn = 0; jn < 4; ++jn) {
      cplx acc = 0.0;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) acc += T[(a * 8 + b * 4 + a * 2 + b) * 4 + jn];
      K[4 * m + jn] = acc;
    }
  }
  // Werner pairs, CNOT / H and depolarising noise give REAL tables: they are staged as packed doubles (half the shared
  // memory, half the multiply-adds); anything complex (a phase gate folded in) keeps the complex form
  bool real_tab = true;
  for (int k = 0; k < 8 + 128; ++k) real_tab = real_tab && K[k].imag() == 0.0;
  int nk = 8 + 128;
  if (real_tab) {
    for (int k = 0; k < 64; ++k) K[8 + k] = cplx(K[8 + 2 * k].real(), K[8 + 2 * k + 1].real());
    nk = 8 + 64;
  }
  HostOp f = Q[jp];   // the PROBS op, reshaped: nb = 4 marks the fused form (Planner::localise, op_measure)
  f.m.nb = 4;
  f.m.b[7] = real_tab ? 1 : 0;
  f.m.b[0] = (uint8_t)(A + n); f.m.b[1] = (uint8_t)(y + n); f.m.b[2] = (uint8_t)A; f.m.b[3] = (uint8_t)y;
  f.m.b[4] = (uint8_t)A; f.m.b[5] = (uint8_t)y;
  f.need = Ab | yb;
  f.live_after = col.live_after;
  f.m.c_off = add_consts(h, K, nk); f.nconst = nk;
  HostOp g = fin;
  g.flip = 0.0;   // folded into T_m
  g.live_after = col.live_after;
  std::vector<HostOp> out;
  out.reserve(Q.size());
  size_t na = 0;
  for (size_t j = 0; j < Q.size(); ++j) {
    if (j == i || j == jp + 2) continue;
    if (na < absorbed.size() && absorbed[na] == j) { ++na; continue; }
    if (j == jp) { out.push_back(f); continue; }
    if (j == jp + 1) { out.push_back(g); continue; }
    out.push_back(Q[j]);
    if (j > i && j < jp) out.back().live_after &= ~(xb | yb);
  }
  Q.swap(out);
  h->stats[7] += (int64_t)absorbed.size() + 2;
  h->cat_fused++;
  return true;
}

// Second form: x is an EXISTING axis (typically the other half of the pair, in the disentangling half of the remote
// gate: CNOT onto the target, H, measure and discard, dqc_control.py:382-403).  The constant, unpredicated ops on
// (A, x) right in front of the measurement fold into the measurement the same way, now as a contraction from the 16
// elements (rA rx cA cx) to the 4 elements (rA cA) left in the |0><0| block of x:
//     T_m[(rA cA), i] = sum_k w_m(k) S[(rA k cA k), i],      t_m[i] = sum_a T_m[(a a), i].
// The scan runs backwards from the PROBS op at queue index jp and stops at the first op on x or A it cannot fold.
bool try_fuse_out_at(dqe_t* h, size_t jp) {
  std::vector<HostOp>& Q = h->queue;
  const int n = h->nmax;
  if (jp + 2 >= Q.size()) return false;
  const HostOp& pr = Q[jp];
  const HostOp& fin = Q[jp + 1];
  const HostOp& col = Q[jp + 2];
  const int x = pr.m.aux;
  if (pr.m.nb == 4 || fin.kind != 1 || col.kind != 0 || col.m.type != OP_COLLAPSE || col.m.b[1] != x || !(col.m.flags & 1))
    return false;
  const uint64_t xb = (1ull << x) | (1ull << (x + n));
  int A = -1;
  uint64_t Ab = 0, skipped = 0;
  std::vector<size_t> absorbed;      // queue indices, descending
  std::vector<Sup16> maps;           // their superoperators, same order
  Sup16 L;
  cplx M[16];
  for (size_t j = jp; j-- > 0;) {
    const HostOp& e = Q[j];
    if (e.kind != 0) continue;
    if (e.m.type == OP_PROBS) break;   // folded ops need not preserve the trace: they stay behind earlier measurements
    const uint64_t tb = touched_bits(e);
    if (!(tb & (xb | Ab))) { skipped |= tb; continue; }
    if (e.m.pred_bit >= 0) break;
    int newA = -1;
    auto want_A = [&](int a) {      // a second axis joins: it must be THE partner, and nothing skipped may touch it
      if (a == x || a < 0 || a >= n) return false;
      if (A >= 0) return A == a;
      if (skipped & ((1ull << a) | (1ull << (a + n)))) return false;
      newA = a;
      return true;
    };
    bool ok = true;
    if (!(tb & xb)) {
      ok = !(tb & ~Ab) && axis_local_matrix(h, e, A, M);
      if (ok) sup_axis_local(L, 0, M);
    } else if (!(tb & ~xb) && axis_local_matrix(h, e, x, M)) {
      sup_axis_local(L, 1, M);
    } else if (e.m.type == OP_MAT1X2) {
      const int T = e.m.b[1];
      const int cr = e.pctrl ? __builtin_ctzll(e.pctrl) - n : -1, cc = e.pctrl2 ? __builtin_ctzll(e.pctrl2) : -1;
      const int O = cr >= 0 ? cr : cc;
      ok = e.m.b[0] == T + n && __builtin_popcountll(e.pctrl) <= 1 && __builtin_popcountll(e.pctrl2) <= 1 &&
           !(cr >= 0 && cc >= 0 && cr != cc) && O != T && (T == x || O == x) && (T == x || want_A(T)) &&
           (O < 0 || O == x || want_A(O));
      if (ok) {
        cplx U[4];
        for (int k = 0; k < 4; ++k) U[k] = cplx(h->cpool[e.m.c_off + k].x, h->cpool[e.m.c_off + k].y);
        sup_mat1x2(L, T == x ? 1 : 0, U, cr >= 0, cc >= 0);
      }
    } else if (e.m.type == OP_DEPOL2) {
      const int a = e.m.b[2], b = e.m.b[3];
      ok = (a == x || b == x) && want_A(a == x ? b : a);
      if (ok) sup_depol2(L, e.m.p[0]);
    } else {
      ok = false;
    }
    if (!ok) break;
    if (newA >= 0) { A = newA; Ab = (1ull << A) | (1ull << (A + n)); }
    absorbed.push_back(j);
    maps.push_back(L);
  }
  if (A < 0 || absorbed.empty()) return false;
  if (getenv("DQE_CAT2_MAX") && h->cat_fused >= atoi(getenv("DQE_CAT2_MAX"))) return false;
  if (getenv("DQE_DUMP_PLAN")) {
    fprintf(stderr, "fuse-out x=%d A=%d jp=%zu absorbed:", x, A, jp);
    for (size_t k : absorbed) fprintf(stderr, " %zu(type %d b=%d,%d,%d,%d pc=%llx pc2=%llx)", k, Q[k].m.type, Q[k].m.b[0], Q[k].m.b[1], Q[k].m.b[2], Q[k].m.b[3], (unsigned long long)Q[k].pctrl, (unsigned long long)Q[k].pctrl2);
    fprintf(stderr, "\n   window:");
    for (size_t k = absorbed.back(); k <= jp + 2; ++k) fprintf(stderr, " [%zu k%d t%d p%d b=%d,%d,%d,%d nb%u]", k, Q[k].kind, Q[k].m.type, Q[k].m.pred_bit, Q[k].m.b[0], Q[k].m.b[1], Q[k].m.b[2], Q[k].m.b[3], Q[k].m.nb);
    fprintf(stderr, "\n");
  }
  Sup16 S;
  sup_identity(S);
  for (size_t k = maps.size(); k-- > 0;) sup_lmul(S, maps[k]);   // earliest op first
  const double e = fin.flip;
  cplx K[32 + 128];
  for (int m = 0; m < 2; ++m) {
    cplx* T = K + 32 + 64 * m;
    for (int o = 0; o < 4; ++o) {
      const int rA = o >> 1, cA = o & 1;
      for (int i = 0; i < 16; ++i) {
        cplx acc = 0.0;
        for (int k = 0; k < 2; ++k) {
          const double w = k == m ? 1.0 - e : e;
          acc += w * S.m[(rA * 8 + k * 4 + cA * 2 + k) * 16 + i];
        }
        T[o * 16 + i] = acc;
      }
    }
    for (int i = 0; i < 16; ++i) K[16 * m + i] = T[0 * 16 + i] + T[3 * 16 + i];
  }
  bool real_tab = true;   // packed doubles when every entry is real, as in the injecting form
  for (int k = 0; k < 32 + 128; ++k) real_tab = real_tab && K[k].imag() == 0.0;
  int nk = 32 + 128;
  if (real_tab) {
    for (int k = 0; k < 64; ++k) K[32 + k] = cplx(K[32 + 2 * k].real(), K[32 + 2 * k + 1].real());
    nk = 32 + 64;
  }
  HostOp f = pr;
  f.m.nb = 4;
  f.m.b[7] = real_tab ? 1 : 0;
  f.m.b[0] = (uint8_t)(A + n); f.m.b[1] = (uint8_t)(x + n); f.m.b[2] = (uint8_t)A; f.m.b[3] = (uint8_t)x;
  f.m.b[4] = (uint8_t)A; f.m.b[5] = (uint8_t)x; f.m.b[6] = 1;   // b[6] = 1: contraction form
  f.need = Ab | xb;
  f.live_after = col.live_after;
  f.m.c_off = add_consts(h, K, nk); f.nconst = nk;
  HostOp g = fin;
  g.flip = 0.0;
  g.live_after = col.live_after;
  std::vector<HostOp> out;
  out.reserve(Q.size());
  for (size_t j = 0; j < Q.size(); ++j) {
    if (j == jp + 2) continue;
    if (std::find(absorbed.begin(), absorbed.end(), j) != absorbed.end()) continue;
    if (j == jp) { out.push_back(f); continue; }
    if (j == jp + 1) { out.push_back(g); continue; }
    out.push_back(Q[j]);
  }
  Q.swap(out);
  h->stats[7] += (int64_t)absorbed.size() + 1;
  h->cat_fused++;
  return true;
}

// Third form (catfuse = 3): the WHOLE remote gate as two ops on the data pair (A, B) -- the second comm qubit y never
// goes live either.  After the two rewrites above a cat-comm remote gate reads
//     E = fused injecting measurement on (A, y)   ...   light ops on y   ...   C = fused contracting measurement on (B, y)
// with, in between, only axis-local ops on y: constant ones, ONE predicate bit (the X correction of
// dqc_control.py:347-348) and up to TWO per-shot idle-time channels D(q) = D0 + q D1 (depolarising / dephasing / Z flip;
// two of them make K bilinear: K00 + q1 K10 + q2 K01 + q1 q2 K11, b[7] = 2).
// Nothing in between may touch A or measure anything (the map T1_m1 on A is applied late, at C).  Then
//     rho'[(rA cA rB cB)] = sum  T2_m2[(rB cB), (rB' ry cB' cy)]  K[(rA cA ry cy), (rA' cA')]  rho[(rA' cA' rB' cB')],
//     K = Mid(pred, q) o T1_m1 = K0 + q K1   (per m1 and predicate value),
// and both outcome distributions are linear functionals of the SAME 5-qubit state: p(m1) = t1_m1 (x) delta(rB, cB)
// (op E', form 3: probabilities + renormalisation only), p(m2 | m1) = u_m2 built per CTA from t2_m2 and K (op W, form 2,
// which then applies the 16 x 16 matrix C_m2 = T2_m2 K to every group).  Real tables only (complex ones keep the
// two-op path).  Constants of W (double2 units): K[m1][pred][order][64 doubles] = 256, T2[m2][64 doubles] = 64,
// t2[m2][16 doubles] = 16, then 160 of scratch (u: 32, C: 128) that the kernel fills.
constexpr int kWholeK = 256, kWholeT2 = 64, kWholet2 = 16, kWholeScratch = 160;   // kWholeK per pair of q-orders (one dynamic channel)
bool try_fuse_whole_at(dqe_t* h, size_t ie) {
  std::vector<HostOp>& Q = h->queue;
  const int n = h->nmax;
  if (ie + 1 >= Q.size()) return false;
  const HostOp& E = Q[ie];
  const HostOp& fin1 = Q[ie + 1];
  if (fin1.kind != 1 || fin1.creg_bit < 0 || !E.m.b[7]) return false;
  const int A = E.m.b[4], y = E.m.b[5];
  const uint64_t Ab = (1ull << A) | (1ull << (A + n)), yb = (1ull << y) | (1ull << (y + n));
  struct Mid { cplx M[16]; int pred; int dyn_reg; int dyn_kind; };
  std::vector<Mid> mids;
  std::vector<size_t> mid_idx;
  size_t jc = Q.size();
  int pred_bit = -1, dyn_reg = -1, dyn_reg2 = -1, ndyn = 0;
  for (size_t j = ie + 2; j < Q.size(); ++j) {
    const HostOp& e = Q[j];
    if (e.kind != 0) continue;
    if (e.m.type == OP_PROBS) {
      if (e.m.nb == 4 && e.m.b[6] == 1 && e.m.b[5] == y && e.m.b[7]) { jc = j; break; }
      return false;   // another measurement in between would see the state without T1_m1 applied
    }
    const uint64_t tb = touched_bits(e);
    if (tb & Ab) return false;
    if (!(tb & yb)) continue;
    if (tb & ~yb) return false;
    Mid md;
    md.pred = e.m.pred_bit; md.dyn_reg = -1; md.dyn_kind = 0;
    if (e.m.type == OP_DYN1 && e.m.b[1] == y) {
      if (e.m.pred_bit >= 0 || ndyn >= 2 || e.m.flags == DYN_AMPDAMP) return false;
      md.dyn_reg = ndyn;   // (index of the dynamic channel, 0 or 1)
      if (ndyn == 0) dyn_reg = e.m.aux; else dyn_reg2 = e.m.aux;
      ++ndyn;
      md.dyn_kind = e.m.flags;
    } else if (axis_local_matrix(h, e, y, md.M)) {
      if (e.m.pred_bit >= 0) {
        if (pred_bit >= 0 && pred_bit != e.m.pred_bit) return false;
        pred_bit = e.m.pred_bit;
      }
    } else {
      return false;
    }
    mids.push_back(md);
    mid_idx.push_back(j);
  }
  if (jc + 1 >= Q.size()) return false;
  const HostOp& C = Q[jc];
  const int B = C.m.b[4];
  if (B == A || B == y) return false;
  const uint64_t Bb = (1ull << B) | (1ull << (B + n));
  if ((E.live_after & Bb) != Bb) return false;   // B must hold a qubit already when the first half runs
  // tables of the two fused ops (real, packed two doubles to an entry)
  auto dbl = [&](const HostOp& o, int first, int k) {   // k-th double of the packed block starting at constant `first`
    const double2 v = h->cpool[o.m.c_off + first + k / 2];
    return (k & 1) ? v.y : v.x;
  };
  double T1[2][64], T2[2][64], t1[2][4], t2[2][16];
  for (int m = 0; m < 2; ++m) {
    for (int k = 0; k < 64; ++k) { T1[m][k] = dbl(E, 8, 64 * m + k); T2[m][k] = dbl(C, 32, 64 * m + k); }
    for (int k = 0; k < 4; ++k) t1[m][k] = h->cpool[E.m.c_off + 4 * m + k].x;
    for (int k = 0; k < 16; ++k) t2[m][k] = h->cpool[C.m.c_off + 16 * m + k].x;
  }
  // Mid(pred, order): product of the mid ops in program order, the predicated ones present or not, the dynamic one
  // replaced by D0 = identity (order 0) or D1 = dD/dq (order 1)
  const int nord = ndyn == 2 ? 4 : 2;   // monomials 1, q1 (, q2, q1 q2)
  std::vector<double> K((size_t)kWholeK * nord, 0.0);
  for (int pv = 0; pv < 2; ++pv)
    for (int order = 0; order < nord; ++order) {
      double Mid[16];
      for (int i = 0; i < 16; ++i) Mid[i] = (i / 4 == i % 4) ? 1.0 : 0.0;
      bool zero = order >= 1 && ndyn == 0;   // no dynamic op: the q-linear part vanishes
      for (const auto& md : mids) {
        double L[16];
        if (md.dyn_reg >= 0) {
          for (int i = 0; i < 16; ++i) L[i] = 0.0;
          if (((order >> md.dyn_reg) & 1) == 0) { for (int i = 0; i < 4; ++i) L[i * 5] = 1.0; }
          else if (md.dyn_kind == DYN_DEPOL) {      // D(q) = (1 - q) Id + q R,  R: e00, e11 -> (e00 + e11) / 2, e01 = e10 = 0
            for (int i = 0; i < 4; ++i) L[i * 5] = -1.0;
            L[0 * 4 + 0] += 0.5; L[0 * 4 + 3] += 0.5; L[3 * 4 + 0] += 0.5; L[3 * 4 + 3] += 0.5;
          } else {                                   // dephasing: e01, e10 *= (1 - q); Z flip with probability q: *= (1 - 2 q)
            const double f = md.dyn_kind == DYN_DEPHASE ? -1.0 : -2.0;
            L[1 * 4 + 1] = f; L[2 * 4 + 2] = f;
          }
        } else {
          bool real_m = true;
          for (int i = 0; i < 16; ++i) real_m = real_m && md.M[i].imag() == 0.0;
          if (!real_m) return false;
          const bool on = md.pred < 0 || pv == 1;
          for (int i = 0; i < 16; ++i) L[i] = on ? md.M[i].real() : ((i / 4 == i % 4) ? 1.0 : 0.0);
        }
        double R[16];
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += L[i * 4 + k] * Mid[k * 4 + j];
            R[i * 4 + j] = acc;
          }
        for (int i = 0; i < 16; ++i) Mid[i] = R[i];
      }
      for (int m1 = 0; m1 < 2; ++m1)
        for (int rA = 0; rA < 2; ++rA)
          for (int cA = 0; cA < 2; ++cA)
            for (int yy = 0; yy < 4; ++yy)        // (ry cy) of the output
              for (int jn = 0; jn < 4; ++jn) {    // (rA' cA')
                double acc = 0.0;
                if (!zero)
                  for (int y2 = 0; y2 < 4; ++y2) {  // (ry' cy')
                    const int e1 = rA * 8 + (y2 >> 1) * 4 + cA * 2 + (y2 & 1);
                    acc += Mid[yy * 4 + y2] * T1[m1][e1 * 4 + jn];
                  }
                K[(size_t)(((m1 * 2 + pv) * nord + order) * 64) + ((rA * 2 + cA) * 4 + yy) * 4 + jn] = acc;
              }
    }
  // ---- E': probabilities of m1 as a functional on the (A, B) group + renormalisation (form 3)
  std::vector<cplx> KE(32 + 64, cplx(0.0));
  for (int m = 0; m < 2; ++m)
    for (int i = 0; i < 16; ++i) {
      const int rA = (i >> 3) & 1, rB = (i >> 2) & 1, cA = (i >> 1) & 1, cB = i & 1;
      KE[16 * m + i] = rB == cB ? t1[m][rA * 2 + cA] : 0.0;
    }
  HostOp e2 = E;
  e2.m.b[0] = (uint8_t)(A + n); e2.m.b[1] = (uint8_t)(B + n); e2.m.b[2] = (uint8_t)A; e2.m.b[3] = (uint8_t)B;
  e2.m.b[4] = (uint8_t)A; e2.m.b[5] = (uint8_t)B; e2.m.b[6] = 3; e2.m.b[7] = 1;
  e2.need = Ab | Bb;
  e2.live_after = E.live_after & ~yb;
  e2.m.c_off = add_consts(h, KE.data(), 32); e2.nconst = 32;
  HostOp f1 = fin1;
  f1.live_after &= ~yb;
  // ---- W: the second measurement and the whole map (form 2)
  // The predicate of the mid ops is nearly always the first outcome itself (the X correction of dqc_control.py:347-348)
  // or absent: then only K[m1][pred = m1] (or pred = 0) can ever be selected, and the table ships without the other
  // half (2 KiB less per gate, 4 KiB with two idle channels: shared memory per CTA and L2 traffic per pass)
  const bool compact = pred_bit < 0 || pred_bit == fin1.creg_bit;
  const int nK = (compact ? kWholeK / 2 : kWholeK) * nord / 2;   // double2 entries of K
  std::vector<cplx> KW((size_t)(nK + kWholeT2 + kWholet2 + kWholeScratch), cplx(0.0));
  if (compact) {
    for (int m1 = 0; m1 < 2; ++m1) {
      const int pv = pred_bit < 0 ? 0 : m1;
      const size_t src = (size_t)((m1 * 2 + pv) * nord) * 64, dst = (size_t)(m1 * nord) * 64;
      for (int k = 0; k < nord * 32; ++k) KW[dst / 2 + k] = cplx(K[src + 2 * k], K[src + 2 * k + 1]);
    }
  } else {
    for (int k = 0; k < nK; ++k) KW[k] = cplx(K[2 * k], K[2 * k + 1]);
  }
  for (int m = 0; m < 2; ++m) {
    for (int k = 0; k < 32; ++k) KW[nK + 32 * m + k] = cplx(T2[m][2 * k], T2[m][2 * k + 1]);
    for (int k = 0; k < 8; ++k) KW[nK + kWholeT2 + 8 * m + k] = cplx(t2[m][2 * k], t2[m][2 * k + 1]);
  }
  HostOp w = C;
  w.m.b[0] = (uint8_t)(A + n); w.m.b[1] = (uint8_t)(B + n); w.m.b[2] = (uint8_t)A; w.m.b[3] = (uint8_t)B;
  w.m.b[4] = (uint8_t)A; w.m.b[5] = (uint8_t)B; w.m.b[6] = 2; w.m.b[7] = (uint8_t)((nord == 4 ? 2 : 1) | (compact ? 4 : 0));
  w.need = Ab | Bb;
  w.m.aux = (int32_t)((uint32_t)fin1.creg_bit | ((uint32_t)(pred_bit + 1) << 8) | ((uint32_t)(dyn_reg + 1) << 16) |
                      ((uint32_t)(dyn_reg2 + 1) << 24));
  w.m.c_off = add_consts(h, KW.data(), (int)KW.size()); w.nconst = (int)KW.size();
  std::vector<HostOp> out;
  out.reserve(Q.size());
  size_t nm = 0;
  for (size_t j = 0; j < Q.size(); ++j) {
    if (nm < mid_idx.size() && mid_idx[nm] == j) { ++nm; continue; }
    if (j == ie) { out.push_back(e2); continue; }
    if (j == ie + 1) { out.push_back(f1); continue; }
    if (j == jc) { out.push_back(w); continue; }
    out.push_back(Q[j]);
    if (j > ie && j < jc) out.back().live_after &= ~yb;
  }
  Q.swap(out);
  h->stats[7] += (int64_t)mid_idx.size();
  h->cat_fused++;
  return true;
}

void fuse_cat_measure(dqe_t* h) {
  if (h->formalism != DQE_DM || !h->catfuse || !cat_cluster_mode(h)) return;
  for (size_t i = 0; i < h->queue.size(); ++i) {
    const HostOp& o = h->queue[i];
    if (o.kind != 0 || o.m.type != OP_INJECT || o.m.nb != 4 || (o.m.flags & 2) || o.m.pred_bit >= 0) continue;
    const int a0 = o.m.b[2], a1 = o.m.b[3];
    if (try_fuse_cat_at(h, i, a0, a1, true) || try_fuse_cat_at(h, i, a1, a0, false)) --i;   // INJECT erased: rescan slot i
  }
  if (h->catfuse < 2) return;
  for (size_t j = 0; j < h->queue.size(); ++j) {
    const HostOp& o = h->queue[j];
    if (o.kind != 0 || o.m.type != OP_PROBS || o.m.nb == 4) continue;
    const size_t before = h->queue.size();
    if (try_fuse_out_at(h, j)) j -= (before - h->queue.size()) - 1;   // the fused PROBS moved down by the absorbed ops; skip its finish
  }
  if (h->catfuse < 3) return;
  for (size_t j = 0; j < h->queue.size(); ++j) {
    const HostOp& o = h->queue[j];
    if (o.kind != 0 || o.m.type != OP_PROBS || o.m.nb != 4 || o.m.b[6] != 0) continue;
    try_fuse_whole_at(h, j);   // (E stays at index j, only ops behind it disappear)
  }
}

int build_segs(uint64_t mask, Seg* segs, int cap) {
  int n = 0, src = 0;
  for (int b = 0; b < 64;) {
    if ((mask >> b) & 1ull) {
      int len = 0;
      while (b + len < 64 && ((mask >> (b + len)) & 1ull)) ++len;
      if (n >= cap) return -1;
      segs[n].src = (uint8_t)src; segs[n].len = (uint8_t)len; segs[n].dst = (uint8_t)b; segs[n].pad = 0;
      ++n; src += len; b += len;
    } else ++b;
  }
  return n;
}

// tile = lowest c live bits + the needed bits above them; returns the tile mask or 0 if the
// needed bits do not fit with a contiguous low run of at least min_run bits.
// Tile size of a pass.  Small density-matrix registers (tile_auto): passes over <= 12 live bits take 2^11 tiles on 128
// threads (two CTAs per shot: fewer per-CTA prologues, cluster barriers and constant copies -- Grover-5 with fused
// remote gates: 65k vs 53k shots/s), wider ones 2^10 tiles on 64 threads (16 small CTAs per shot overlap better:
// 21.0k vs 18.0k shots/s on the unfused 7-qubit passes).  Measured on B200, profiles/r02t_summary.md.
int tile_bits_for(const dqe_t* h, uint64_t live) {
  if (!h->tile_auto) return h->tile_bits;
  return __builtin_popcountll(live) <= 12 ? 11 : 10;
}
int threads_for(const dqe_t* h, int t) {
  if (!h->tile_auto) return h->threads;
  return t >= 11 ? 128 : 64;
}
bool choose_tile(const dqe_t* h, uint64_t need, uint64_t live, bool force, uint64_t* tile_out, int* t_out) {
  const int L = __builtin_popcountll(live);
  const int t = std::min(tile_bits_for(h, live), L);
  int pos[64];
  int n = 0;
  for (int b = 0; b < 64; ++b) if ((live >> b) & 1ull) pos[n++] = b;
  for (int c = t; c >= 0; --c) {
    int hi = 0;
    for (int r = c; r < L; ++r) if ((need >> pos[r]) & 1ull) ++hi;
    if (c + hi <= t) {
      if (!force && c < std::min(h->min_run_bits, L) && c + hi < L) return false;
      uint64_t tile = 0;
      for (int r = 0; r < c; ++r) tile |= 1ull << pos[r];
      tile |= need;
      // when fewer than t bits are used (c == t impossible here), extend with next live bits
      int used = c + hi;
      for (int r = c; r < L && used < t; ++r)
        if (!((tile >> pos[r]) & 1ull)) { tile |= 1ull << pos[r]; ++used; }
      *t_out = used;
      *tile_out = tile;
      return true;
    }
  }
  return false;
}

int local_of(uint64_t tile, int phys) { return __builtin_popcountll(tile & ((1ull << phys) - 1ull)); }

struct Planner {
  dqe_t* h;
  std::deque<HostOp> fused;          // OP_MEASURE ops synthesised from PROBS + finish pairs
  std::vector<PlannedPass> passes;
  std::vector<MicroOp> dev_ops;
  std::vector<double2> dev_consts;
  std::string err;

  // ---- helpers ------------------------------------------------------------------------------
  int copy_consts(const HostOp* ho, size_t pass_const_begin, const int* perm4 = nullptr) {
    const int off = (int)(dev_consts.size() - pass_const_begin);
    if (perm4) {  // 4x4 table re-indexed for swapped qubit order
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) dev_consts.push_back(h->cpool[ho->m.c_off + perm4[i] * 4 + perm4[j]]);
    } else {
      for (int i = 0; i < ho->nconst; ++i) dev_consts.push_back(h->cpool[ho->m.c_off + i]);
    }
    return off;
  }

  // localise one host op as a stand-alone micro-op of the pass
  MicroOp localise(const HostOp* ho, uint64_t tile, size_t pass_const_begin) {
    MicroOp m = ho->m;
    if (m.type == OP_LADDER) {
      // centre like a diagonal bit; leaves: tile-local ones first (ascending, so that the ones on thread-id bits
      // lead), then the outer ones; the phases follow their leaves
      const uint8_t* lf = reinterpret_cast<const uint8_t*>(&ho->m) + kLadderLeafOff;
      uint8_t* out = reinterpret_cast<uint8_t*>(&m) + kLadderLeafOff;
      const int ph = ho->m.b[0], nl = ho->m.aux;
      m.b[0] = ((tile >> ph) & 1ull) ? (uint8_t)local_of(tile, ph) : (uint8_t)(kOuter + ph);
      std::vector<std::pair<int, int>> loc, outer;   // (position, leaf index)
      for (int j = 0; j < nl; ++j) {
        if ((tile >> lf[j]) & 1ull) loc.push_back({local_of(tile, lf[j]), j});
        else outer.push_back({kOuter + lf[j], j});
      }
      std::sort(loc.begin(), loc.end());
      m.c_off = (int)(dev_consts.size() - pass_const_begin);
      int k = 0;
      for (auto& pr : loc) { out[k++] = (uint8_t)pr.first; dev_consts.push_back(h->cpool[ho->m.c_off + pr.second]); }
      for (auto& pr : outer) { out[k++] = (uint8_t)pr.first; dev_consts.push_back(h->cpool[ho->m.c_off + pr.second]); }
      m.flags = (int)loc.size() | ((int)outer.size() << 8);
      return m;
    }
    if (m.type == OP_MEASURE && ho->m.nb == 4) {
      // fused inject -> interact -> measure (fuse_cat_measure): b[0..3] = tile positions of (row A, row y, col A, col y),
      // b[4], b[5] = the axes A and y themselves (every OTHER axis must be diagonal for an element to count)
      for (int i = 0; i < 4; ++i) m.b[i] = (uint8_t)local_of(tile, ho->m.b[i]);
      m.c_off = copy_consts(ho, pass_const_begin);
      // which groups count towards the probabilities: those diagonal in every other axis.  Resolved here into
      // (column, row) tile-position pairs (both bits in the tile), (tile position, base bit) pairs (one bit outside)
      // and a mask of axes that live in the tile base only -- the kernel tests a few shifts per group
      const int n = h->nmax, axA = ho->m.b[4], axY = ho->m.b[5];
      uint8_t pairs[16], mixed[8];
      int np = 0, nm = 0;
      bool fallback = false;
      uint64_t eq = 0;
      for (int a = 0; a < n; ++a) {
        if (a == axA || a == axY) continue;
        const bool in_c = (tile >> a) & 1ull, in_r = (tile >> (a + n)) & 1ull;
        if (in_c && in_r) {
          if (np == 8) { fallback = true; break; }
          pairs[2 * np] = (uint8_t)local_of(tile, a); pairs[2 * np + 1] = (uint8_t)local_of(tile, a + n); ++np;
        } else if (in_c || in_r) {
          if (nm == 4) { fallback = true; break; }
          mixed[2 * nm] = (uint8_t)local_of(tile, in_c ? a : a + n); mixed[2 * nm + 1] = (uint8_t)(in_c ? a + n : a); ++nm;
        } else {
          eq |= 1ull << a;
        }
      }
      memset(&m.p[1], 0, 24);
      if (fallback) m.lctrl2 = 0x80000000u;
      else {
        m.lctrl2 = (uint32_t)np | ((uint32_t)nm << 8);
        memcpy(&m.p[1], pairs, 2 * np);
        memcpy(&m.p[3], mixed, 2 * nm);
        m.octrl2 = eq;
      }
      return m;
    }
    if (m.type == OP_DIAG) {
      for (uint32_t i = 0; i < m.nb; ++i) {
        const int ph = ho->m.b[i];
        m.b[i] = ((tile >> ph) & 1ull) ? (uint8_t)local_of(tile, ph) : (uint8_t)(kOuter + ph);
      }
    } else if (m.type != OP_PROBS && m.type != OP_MEASURE) {
      for (uint32_t i = 0; i < m.nb; ++i) m.b[i] = (uint8_t)local_of(tile, ho->m.b[i]);
    } else {
      const int ph = m.aux;   // measured axis: column / ket bit
      m.b[0] = ((tile >> ph) & 1ull) ? (uint8_t)local_of(tile, ph) : (uint8_t)(kOuter + ph);
      // dm: index of the axis among the free variables of the diagonal enumeration (column bits in the
      // tile, ascending), 0xff = fixed by the tile base
      m.b[1] = ((tile >> ph) & 1ull) ? (uint8_t)__builtin_popcountll(tile & ((1ull << ph) - 1ull)) : (uint8_t)0xff;
    }
    if (m.type == OP_MEASURE) return m;   // lctrl / octrl / flags carry forced outcome, draw index, creg bit
    m.lctrl = 0; m.octrl = 0; m.lctrl2 = 0; m.octrl2 = 0;
    for (int b = 0; b < 64; ++b) {
      if ((ho->pctrl >> b) & 1ull) {
        if ((tile >> b) & 1ull) m.lctrl |= 1u << local_of(tile, b); else m.octrl |= 1ull << b;
      }
      if ((ho->pctrl2 >> b) & 1ull) {
        if ((tile >> b) & 1ull) m.lctrl2 |= 1u << local_of(tile, b); else m.octrl2 |= 1ull << b;
      }
    }
    if (m.type == OP_MAT2 && h->formalism == DQE_KET && is_swap(ho)) {   // a permutation: no constants, no arithmetic
      m.type = OP_SWAP; m.c_off = 0;
      return m;
    }
    m.c_off = ho->nconst ? copy_consts(ho, pass_const_begin) : 0;
    return m;
  }
  bool is_swap(const HostOp* ho) const {
    static const int one[4] = {0, 6, 9, 15};
    const double2* u = &h->cpool[ho->m.c_off];
    for (int i = 0; i < 16; ++i) {
      const bool is1 = i == one[0] || i == one[1] || i == one[2] || i == one[3];
      if (u[i].y != 0.0 || u[i].x != (is1 ? 1.0 : 0.0)) return false;
    }
    return true;
  }

  // dm: axes whose (row, col) bits an op mixes or uses as in-tile controls; false = not groupable
  bool group_axes(const HostOp* ho, uint64_t tile, int* ax, int* nax) {
    const int n = h->nmax;
    *nax = 0;
    auto both_in_tile = [&](int a) { return ((tile >> a) & 1ull) && ((tile >> (a + n)) & 1ull); };
    auto add = [&](int a) {
      for (int i = 0; i < *nax; ++i) if (ax[i] == a) return true;
      if (*nax >= 2 || !both_in_tile(a)) return false;
      ax[(*nax)++] = a;
      return true;
    };
    const MicroOp& m = ho->m;
    switch (m.type) {
      case OP_PAULI1: case OP_COLLAPSE: case OP_TRACE: case OP_XPAT: case OP_DYN1:
        return add(m.b[1]);
      case OP_MAT2:
        // a dense 4x4 block inside the register-resident group interpreter costs every other sub-op its
        // registers (see run_group): an axis-local superoperator joins a group as its light factors
        // (Pauli channels, U rho U^dagger); anything else stays a stand-alone sweep
        if (ho->pctrl) return false;
        if (m.b[0] == m.b[1] + n && ho->nf > 0 && h->factor) return add(m.b[1]);
        return false;
      case OP_DEPOL2: case OP_INJECT:
        return add(m.b[2]) && add(m.b[3]);
      case OP_MAT1X2: {
        if (!add(m.b[1])) return false;
        if (__builtin_popcountll(ho->pctrl) > 1 || __builtin_popcountll(ho->pctrl2) > 1) return false;
        const int rc = ho->pctrl ? __builtin_ctzll(ho->pctrl) - n : -1;    // control axis seen by the row side
        const int cc = ho->pctrl2 ? __builtin_ctzll(ho->pctrl2) : -1;     // ... by the column side
        if (rc >= 0 && cc >= 0 && rc != cc) return false;
        const int c = rc >= 0 ? rc : cc;
        if (c < 0) return true;
        const bool r_local = rc >= 0 && ((tile >> (rc + n)) & 1ull), c_local = cc >= 0 && ((tile >> cc) & 1ull);
        if (!r_local && !c_local) return true;   // control entirely outside the tile: uniform per CTA
        return add(c);                           // needs both bits of the control axis in the tile
      }
      case OP_DIAG:
        // D rho D^dagger on one axis rides as an X-pattern sub-op; wider tables stay stand-alone
        if (m.nb == 2 && m.b[0] == m.b[1] + n && h->factor) return add(m.b[1]);
        return false;
      default:
        return false;
    }
  }

  bool is_x(const HostOp* ho) const {
    if (ho->m.type != OP_MAT1X2) return false;
    const double2* u = &h->cpool[ho->m.c_off];
    return u[0].x == 0 && u[0].y == 0 && u[3].x == 0 && u[3].y == 0 && u[1].x == 1 && u[1].y == 0 && u[2].x == 1 &&
           u[2].y == 0;
  }
  // X / CX whose controls (if any) are the other axis of the group or uniform per tile, unpredicated:
  // can be applied by the group's load / store addresses
  bool hoistable_x(const HostOp* ho) const { return is_x(ho) && ho->m.pred_bit < 0; }
  bool x_ctrl_in_group(const HostOp* ho, uint64_t tile) const {
    if (!ho->pctrl || !ho->pctrl2) return false;
    return ((tile >> __builtin_ctzll(ho->pctrl)) & 1ull) && ((tile >> __builtin_ctzll(ho->pctrl2)) & 1ull);
  }

  MicroOp make_sub(const HostOp* ho, int A, int B, uint64_t tile, size_t pcb) {
    static const int kSwap[4] = {0, 2, 1, 3};
    MicroOp m;
    memset(&m, 0, sizeof(m));
    const MicroOp& s = ho->m;
    m.pred_bit = s.pred_bit;
    auto role = [&](int axis) { return axis == A ? 0 : 1; };
    switch (s.type) {
      case OP_PAULI1:
        m.type = G_PAULI1; m.flags = role(s.b[1]);
        for (int i = 0; i < 4; ++i) m.p[i] = s.p[i];
        break;
      case OP_XPAT:
        m.type = G_XPAT; m.flags = role(s.b[1]); m.c_off = copy_consts(ho, pcb);
        break;
      case OP_DYN1:
        m.type = G_DYN1; m.flags = role(s.b[1]) | (s.flags << 1); m.aux = s.aux;
        break;
      case OP_MAT1X2: {
        m.type = G_MAT1X2; m.flags = role(s.b[1]);
        if (is_x(ho)) m.flags |= 16;   // Pauli-X: a permutation, done in the addresses (GPerm)
        if (ho->pctrl) {
          const int b = __builtin_ctzll(ho->pctrl);
          if ((tile >> b) & 1ull) m.flags |= 2; else m.octrl = ho->pctrl;
        }
        if (ho->pctrl2) {
          const int b = __builtin_ctzll(ho->pctrl2);
          if ((tile >> b) & 1ull) m.flags |= 4; else m.octrl2 = ho->pctrl2;
        }
        m.c_off = copy_consts(ho, pcb);
        break;
      }
      case OP_DEPOL2: m.type = G_DEPOL2; m.p[0] = s.p[0]; break;
      case OP_COLLAPSE: m.type = G_COLLAPSE; m.flags = role(s.b[1]) | ((s.flags & 1) ? 8 : 0); m.p[0] = s.p[0]; break;
      case OP_TRACE: m.type = G_TRACE; m.flags = role(s.b[1]); break;
      case OP_INJECT:
        m.type = G_INJECT;
        m.c_off = copy_consts(ho, pcb, s.b[2] == A ? nullptr : kSwap);
        break;
      default: break;
    }
    return m;
  }

  // sub-ops of one host op inside the group on axes (A, B)
  void emit_sub(const HostOp* ho, int A, int B, uint64_t tile, size_t pcb) {
    const MicroOp& s = ho->m;
    const int n = h->nmax;
    if (s.type == OP_MAT2 && s.b[0] == s.b[1] + n && ho->nf > 0) {
      for (int i = 0; i < ho->nf; ++i) {
        const Factor& f = ho->f[i];
        MicroOp m;
        memset(&m, 0, sizeof(m));
        m.pred_bit = s.pred_bit;
        m.flags = s.b[1] == A ? 0 : 1;
        if (f.kind == 1) {
          m.type = G_PAULI1;
          for (int k = 0; k < 4; ++k) m.p[k] = f.d[k];
        } else {
          m.type = G_MAT1X2;
          if (f.d[0] == 0 && f.d[1] == 0 && f.d[6] == 0 && f.d[7] == 0 && f.d[2] == 1 && f.d[3] == 0 && f.d[4] == 1 &&
              f.d[5] == 0)
            m.flags |= 16;
          m.c_off = (int)(dev_consts.size() - pcb);
          for (int k = 0; k < 4; ++k) dev_consts.push_back(make_double2(f.d[2 * k], f.d[2 * k + 1]));
        }
        dev_ops.push_back(m);
      }
      return;
    }
    if (s.type == OP_DIAG && s.nb == 2 && s.b[0] == s.b[1] + n) {
      // table t[r][c] (index r * 2 + c) as the X pattern e00' = a e00, e11' = d e11, e01' = f e01, e10' = k e10
      MicroOp m;
      memset(&m, 0, sizeof(m));
      m.pred_bit = s.pred_bit;
      m.type = G_XPAT; m.flags = s.b[1] == A ? 0 : 1;
      m.c_off = (int)(dev_consts.size() - pcb);
      const double2* t = &h->cpool[s.c_off];
      const double2 z = make_double2(0.0, 0.0);
      const double2 c8[8] = {t[0], z, z, t[3], t[1], z, z, t[2]};
      for (int k = 0; k < 8; ++k) dev_consts.push_back(c8[k]);
      dev_ops.push_back(m);
      return;
    }
    dev_ops.push_back(make_sub(ho, A, B, tile, pcb));
  }

  struct Item {
    bool group;
    bool classical;     // classical instructions of one span, hoisted to the span's start
    std::vector<const HostOp*> ops;
    int ax[2], nax;
    uint64_t touched;
  };

  // classical instructions of one span -> OP_CLASSICAL micro-ops: the clock arithmetic in program order (it is a
  // dependency chain: lane 0 of warp 0 runs it), then the QEXP evaluations (independent of each other, one exp
  // each: lane i runs the i-th).  MicroOp.flags = instruction count, c_off = 1 for the QEXP micro-ops.
  // Immediates go to the pass constant pool as packed doubles (equal values share a slot).
  void emit_classical(const std::vector<const HostOp*>& ops, size_t pcb) {
    std::vector<double> vals;
    std::vector<int> idx(ops.size()), idx2(ops.size(), 0);
    auto slot_of = [&](double v) {
      for (size_t j = 0; j < vals.size(); ++j)
        if (memcmp(&vals[j], &v, sizeof(double)) == 0) return (int)j;
      vals.push_back(v);
      return (int)vals.size() - 1;
    };
    for (size_t i = 0; i < ops.size(); ++i) {
      idx[i] = slot_of(ops[i]->imm);
      if (ops[i]->ci.op == CL_QEXP2 || ops[i]->ci.op == CL_GEOM) idx2[i] = slot_of(ops[i]->imm2);
    }
    const int base = (int)(dev_consts.size() - pcb) * 2;
    for (size_t j = 0; j < vals.size(); j += 2)
      dev_consts.push_back(make_double2(vals[j], j + 1 < vals.size() ? vals[j + 1] : 0.0));
    // a block right behind another block (two spans separated by a fence only): the threads must meet in
    // between, or a fast warp's writes of this block race with a slow warp's writes of the previous one
    bool need_barrier = !dev_ops.empty() && dev_ops.back().type == OP_CLASSICAL;
    {
      // one block in program order (every thread runs every instruction, so there is nothing to gain from
      // separating the clock chain from the exp() evaluations -- and a micro-op less to dispatch per span)
      MicroOp m;
      int n = 0, nserial = 0;
      auto open_op = [&]() {
        memset(&m, 0, sizeof(m));
        m.type = OP_CLASSICAL; m.pred_bit = -1;
        n = 0; nserial = 0;
      };
      auto close_op = [&]() {
        if (n == 0) return;
        // the payload overlays the fields behind the 16-byte dispatch header; the marker rides in c_off
        m.flags = n | (nserial << 8); m.c_off = need_barrier ? 2 : 0;
        need_barrier = false;
        dev_ops.push_back(m);
      };
      open_op();
      // the clock chain first, in program order; the exp() evaluations behind it (they feed channels only --
      // dqe_classical_op contract -- and are independent of each other: the kernel spreads them over lanes)
      std::vector<size_t> order;
      for (size_t i = 0; i < ops.size(); ++i) if (!(ops[i]->ci.op == CL_QEXP || ops[i]->ci.op == CL_QEXP2)) order.push_back(i);
      for (size_t i = 0; i < ops.size(); ++i) if (ops[i]->ci.op == CL_QEXP || ops[i]->ci.op == CL_QEXP2) order.push_back(i);
      for (size_t oi = 0; oi < order.size(); ++oi) {
        const size_t i = order[oi];
        const HostOp* ho = ops[i];
        if (!(ho->ci.op == CL_QEXP || ho->ci.op == CL_QEXP2)) nserial = n + 1;
        ClInstr ci = ho->ci;
        ci.imm = (uint16_t)(base + idx[i]);
        if (ci.op == CL_QEXP2 || ci.op == CL_GEOM) { const int i2 = base + idx2[i]; ci.cbit = (int8_t)(i2 & 0xff); ci.pad = (uint8_t)(i2 >> 8); }
        memcpy(reinterpret_cast<char*>(&m) + 16 + 8 * n, &ci, sizeof(ci));
        if (++n == kClPerOp) { close_op(); open_op(); }
      }
      close_op();
    }
  }

  // ---- ket: dense k-qubit blocks (OP_KBLOCK, FP64 tensor cores) -----------------------------------------
  // bits an op reads or writes inside the tile (targets, controls, diagonal bits); false = not fusable
  bool kb_bits(const HostOp* ho, uint64_t tile, uint64_t* bits) const {
    if (ho->kind != 0 || ho->m.pred_bit >= 0 || ho->pctrl2) return false;
    uint64_t b = 0;
    const MicroOp& m = ho->m;
    switch (m.type) {
      case OP_MAT1: b = (1ull << m.b[0]) | ho->pctrl; break;
      case OP_MAT2: if (ho->pctrl) return false; b = (1ull << m.b[0]) | (1ull << m.b[1]); break;
      case OP_DIAG: case OP_KBLOCK: for (uint32_t i = 0; i < m.nb; ++i) b |= 1ull << m.b[i]; break;
      default: return false;
    }
    if (b & ~tile) return false;   // a control / diagonal bit outside the tile depends on the tile base
    *bits = b;
    return true;
  }
  // apply a fusable ket op to every column of M (dimension 2^s, row-major), S = block bits, S[0] = MSB
  void kb_apply(std::vector<cplx>& M, const std::vector<int>& S, const HostOp* ho) const {
    const int s = (int)S.size(), D = 1 << s;
    auto pos = [&](int bit) { for (int i = 0; i < s; ++i) if (S[i] == bit) return s - 1 - i; return -1; };   // weight
    auto cst = [&](int i) { return cplx(h->cpool[ho->m.c_off + i].x, h->cpool[ho->m.c_off + i].y); };
    const MicroOp& m = ho->m;
    std::vector<cplx> out(M.size());
    for (int col = 0; col < D; ++col) {
      std::vector<cplx> v(D), w(D);
      for (int r = 0; r < D; ++r) v[r] = M[r * D + col];
      if (m.type == OP_MAT1) {
        const int t = pos(m.b[0]);
        uint32_t cm = 0;
        for (int b = 0; b < 64; ++b) if ((ho->pctrl >> b) & 1ull) cm |= 1u << pos(b);
        w = v;
        for (int r = 0; r < D; ++r) {
          if ((r >> t) & 1) continue;
          if ((r & cm) != cm) continue;
          const int r1 = r | (1 << t);
          w[r] = cst(0) * v[r] + cst(1) * v[r1];
          w[r1] = cst(2) * v[r] + cst(3) * v[r1];
        }
      } else if (m.type == OP_MAT2) {
        const int th = pos(m.b[0]), tl = pos(m.b[1]);
        w = v;
        for (int r = 0; r < D; ++r) {
          if (((r >> th) & 1) || ((r >> tl) & 1)) continue;
          const int idx[4] = {r, r | (1 << tl), r | (1 << th), r | (1 << th) | (1 << tl)};
          for (int i = 0; i < 4; ++i) {
            cplx acc = 0.0;
            for (int j = 0; j < 4; ++j) acc += cst(i * 4 + j) * v[idx[j]];
            w[idx[i]] = acc;
          }
        }
      } else if (m.type == OP_DIAG) {
        for (int r = 0; r < D; ++r) {
          int t = 0;
          for (uint32_t i = 0; i < m.nb; ++i) t |= ((r >> pos(m.b[i])) & 1) << (m.nb - 1 - i);
          w[r] = cst(t) * v[r];
        }
      } else {   // OP_KBLOCK on a subset of S
        const int k = (int)m.nb, K = 1 << k;
        w = v;
        for (int r = 0; r < D; ++r) {
          bool base = true;
          for (int i = 0; i < k; ++i) if ((r >> pos(m.b[i])) & 1) base = false;
          if (!base) continue;
          std::vector<int> idx(K);
          for (int e = 0; e < K; ++e) {
            int x = r;
            for (int i = 0; i < k; ++i) if ((e >> (k - 1 - i)) & 1) x |= 1 << pos(m.b[i]);
            idx[e] = x;
          }
          for (int i = 0; i < K; ++i) {
            cplx acc = 0.0;
            for (int j = 0; j < K; ++j) acc += cst(i * K + j) * v[idx[j]];
            w[idx[i]] = acc;
          }
        }
      }
      for (int r = 0; r < D; ++r) out[r * D + col] = w[r];
    }
    M.swap(out);
  }
  std::deque<HostOp> kb_ops;       // synthesised OP_KBLOCK host ops (constants appended to the handle's pool)

  // greedy: consecutive fusable ops whose bits stay inside <= 4 tile bits become one dense block when that saves
  // enough shared-memory sweeps; ops on disjoint bits pass in front of the pending block (they commute with it)
  void fuse_kblocks(std::vector<const HostOp*>& cur, uint64_t tile, int t) {
    if (!h->kblock || h->formalism != DQE_KET || t < 6) return;
    const int kmax = std::min(4, t - 3);
    std::vector<const HostOp*> out, blk;
    uint64_t S = 0;
    double weight = 0.0;
    auto flush_blk = [&]() {
      const int k = __builtin_popcountll(S);
      bool has_block = false;
      for (const HostOp* ho : blk) has_block |= ho->m.type == OP_KBLOCK;
      // cost model (r02h microbench, ket-28): a k = 4 block is 32 DMMA per 128 amplitudes ~ 0.9 ms of FP64 pipe,
      // a plain in-tile 1q sweep ~ 0.24 ms: the block pays from ~7 one-qubit-gate equivalents on (k = 3: ~4)
      if (blk.size() >= 2 && k >= 3 && (weight >= (k == 3 ? 4.0 : 7.0) || has_block)) {
        std::vector<int> bits;
        for (int b = 63; b >= 0; --b) if ((S >> b) & 1ull) bits.push_back(b);
        const int D = 1 << k;
        std::vector<cplx> M((size_t)D * D, cplx(0.0));
        for (int i = 0; i < D; ++i) M[(size_t)i * D + i] = 1.0;
        for (const HostOp* ho : blk) kb_apply(M, bits, ho);
        HostOp o = *blk.back();
        memset(&o.m, 0, sizeof(o.m));
        o.nf = 0; o.pctrl = o.pctrl2 = 0;
        o.m.type = OP_KBLOCK; o.m.pred_bit = -1; o.m.nb = (uint32_t)k;
        for (int i = 0; i < k; ++i) o.m.b[i] = (uint8_t)bits[i];
        o.need = S;
        o.m.c_off = add_consts(h, M.data(), D * D); o.nconst = D * D;
        kb_ops.push_back(o);
        out.push_back(&kb_ops.back());
        h->kblocks_fused++;
      } else {
        for (const HostOp* ho : blk) out.push_back(ho);
      }
      blk.clear(); S = 0; weight = 0.0;
    };
    for (const HostOp* ho : cur) {
      uint64_t b = 0;
      if (kb_bits(ho, tile, &b)) {
        if (__builtin_popcountll(S | b) > kmax) flush_blk();
        if (__builtin_popcountll(b) <= kmax) {
          blk.push_back(ho); S |= b;
          weight += ho->m.type == OP_MAT1 ? (ho->pctrl ? 0.6 : 1.0) : ho->m.type == OP_MAT2 ? 2.0 :
                    ho->m.type == OP_DIAG ? 0.25 : 4.0;
          continue;
        }
      }
      if (ho->kind >= 2) { out.push_back(ho); continue; }               // classical: touches no amplitude
      uint64_t tb = touched_bits(*ho);
      if (ho->m.type == OP_PROBS || ho->m.type == OP_MEASURE || ho->m.type == OP_COLLAPSE) tb = ~0ull;
      if (tb & S) flush_blk();
      out.push_back(ho);
    }
    flush_blk();
    cur.swap(out);
  }

  // ---- ket: block sweeps (OP_KGROUP) ---------------------------------------------------------------------
  // A run of ops whose matrices act on <= 4 tile bits (all at tile positions >= 3), with any diagonal tables and
  // stars in between, is one shared-memory round trip: the thread holds the 16 amplitudes those bits span.
  void emit_kgroup(const std::vector<const HostOp*>& blk, uint64_t S, uint64_t tile, int t, size_t pcb) {
    int s4[4], n4 = 0;
    for (int b = 0; b < 64; ++b) if ((S >> b) & 1ull) s4[n4++] = local_of(tile, b);
    for (int pos = t - 1; pos >= 3 && n4 < 4; --pos) {
      bool seen = false;
      for (int i = 0; i < n4; ++i) seen |= s4[i] == pos;
      if (!seen) s4[n4++] = pos;
    }
    std::sort(s4, s4 + 4);
    auto rank_of = [&](int pos) { for (int i = 0; i < 4; ++i) if (s4[i] == pos) return i; return -1; };
    MicroOp g;
    memset(&g, 0, sizeof(g));
    g.type = OP_KGROUP; g.pred_bit = -1; g.nb = 4;
    for (int i = 0; i < 4; ++i) g.b[i] = (uint8_t)s4[i];
    const size_t hdr = dev_ops.size();
    dev_ops.push_back(g);
    for (const HostOp* ho : blk) {
      if (ho->m.type == OP_LADDER) {
        MicroOp m = ho->m;
        const uint8_t* lf = reinterpret_cast<const uint8_t*>(&ho->m) + kLadderLeafOff;
        uint8_t* out = reinterpret_cast<uint8_t*>(&m) + kLadderLeafOff;
        const int ph = ho->m.b[0], nl = ho->m.aux;
        m.b[1] = 0;
        if ((tile >> ph) & 1ull) {
          m.b[0] = (uint8_t)local_of(tile, ph);
          const int r = rank_of(m.b[0]);
          if (r >= 0) m.b[1] = (uint8_t)(r + 1);
        } else m.b[0] = (uint8_t)(kOuter + ph);
        std::vector<std::pair<int, int>> thr, el, outer;
        for (int j = 0; j < nl; ++j) {
          if ((tile >> lf[j]) & 1ull) {
            const int pos = local_of(tile, lf[j]), r = rank_of(pos);
            if (r >= 0) el.push_back({r, j}); else thr.push_back({pos, j});
          } else outer.push_back({kOuter + lf[j], j});
        }
        m.c_off = (int)(dev_consts.size() - pcb);
        int k = 0;
        for (auto* lst : {&thr, &el, &outer})
          for (auto& pr : *lst) { out[k++] = (uint8_t)pr.first; dev_consts.push_back(h->cpool[ho->m.c_off + pr.second]); }
        m.flags = (int)thr.size() | ((int)el.size() << 8) | ((int)outer.size() << 16);
        dev_ops.push_back(m);
        continue;
      }
      MicroOp m = localise(ho, tile, pcb);
      if (m.type == OP_MAT1) {
        m.b[4] = (uint8_t)rank_of(m.b[0]);
        uint32_t thr = 0, el = 0;
        for (int b = 0; b < 32; ++b)
          if ((m.lctrl >> b) & 1u) { const int r = rank_of(b); if (r >= 0) el |= 1u << r; else thr |= 1u << b; }
        m.lctrl = thr; m.lctrl2 = el;
      } else {   // OP_DIAG
        for (uint32_t i = 0; i < m.nb; ++i) {
          m.b[4 + i] = 0;
          if (m.b[i] < kOuter) { const int r = rank_of(m.b[i]); if (r >= 0) m.b[4 + i] = (uint8_t)(r + 1); }
        }
      }
      dev_ops.push_back(m);
    }
    dev_ops[hdr].aux = (int)(dev_ops.size() - hdr - 1);
    dev_ops[hdr].c_off = dev_ops[hdr].aux;
    h->kgroups_built++;
  }

  void emit_ket_seq(const std::vector<const HostOp*>& ops, uint64_t tile, int t, PassParams& pp, size_t pcb, bool& writes) {
    auto emit_plain = [&](const HostOp* ho) {
      if (ho->m.type != OP_PROBS && ho->m.type != OP_MEASURE) writes = true;
      if (ho->m.type == OP_MEASURE) { pp.has_measure = 1; pp.writes_creg = 1; }
      dev_ops.push_back(localise(ho, tile, pcb));
    };
    if (!h->kgroup || t < 7) { for (const HostOp* ho : ops) emit_plain(ho); return; }
    std::vector<const HostOp*> blk;
    uint64_t S = 0;
    int nmat = 0;
    auto flush = [&]() {
      if (nmat >= 1 && blk.size() >= 2) { emit_kgroup(blk, S, tile, t, pcb); writes = true; }
      else for (const HostOp* ho : blk) emit_plain(ho);
      blk.clear(); S = 0; nmat = 0;
    };
    for (const HostOp* ho : ops) {
      bool blockable = false;
      uint64_t tg = 0;
      if (ho->kind == 0) {
        if (ho->m.type == OP_DIAG || ho->m.type == OP_LADDER) blockable = true;
        else if (ho->m.type == OP_MAT1 && !ho->pctrl2 && local_of(tile, ho->m.b[0]) >= 3) { blockable = true; tg = 1ull << ho->m.b[0]; }
      }
      if (blockable && __builtin_popcountll(S | tg) <= 4 && (int)blk.size() < kMaxDiagRun) {
        blk.push_back(ho); S |= tg; nmat += tg != 0;
        continue;
      }
      flush();
      if (blockable) { blk.push_back(ho); S = tg; nmat = tg != 0; }
      else emit_plain(ho);
    }
    flush();
  }

  // Close `cur` as one pass; when the emitted micro-program overflows the per-pass budget (the planner's
  // running estimate is optimistic) the list is cut in two and each half becomes its own pass.
  // live0 = live mask in front of the first op of `cur`.
  size_t cur_first = 0;    // queue index of the first op of `cur`
  bool close(std::vector<const HostOp*>& cur, uint64_t live0) {
    const size_t mark = passes.size();
    const bool ok = close_tagged(cur, live0);
    for (size_t i = mark; i < passes.size(); ++i) { passes[i].first_op = cur_first; passes[i].live_before = live0; }
    return ok;
  }
  bool close_tagged(std::vector<const HostOp*>& cur, uint64_t live0) {
    if (cur.empty()) return true;
    uint64_t need = 0, live = live0;
    for (const HostOp* ho : cur) { need |= ho->need; live |= ho->live_after | ho->need; }
    const size_t ops_mark = dev_ops.size(), consts_mark = dev_consts.size(), passes_mark = passes.size();
    std::vector<const HostOp*> keep = cur;
    overflow = false;
    if (!close_one(cur, need, live)) {
      if (!overflow || keep.size() < 2) return false;
      dev_ops.resize(ops_mark); dev_consts.resize(consts_mark); passes.resize(passes_mark);
      err.clear();
      const size_t mid = keep.size() / 2;
      std::vector<const HostOp*> a(keep.begin(), keep.begin() + mid), b(keep.begin() + mid, keep.end());
      uint64_t live_mid = live0;
      for (const HostOp* ho : a) if (ho->kind <= 1 || true) live_mid = ho->live_after;
      if (!close_tagged(a, live0)) return false;
      if (!close_tagged(b, live_mid)) return false;
      cur.clear();
    }
    return true;
  }
  bool overflow = false;

  bool close_one(std::vector<const HostOp*>& cur, uint64_t need, uint64_t live) {
    if (cur.empty()) return true;
    live |= phys_live(h, h->pin_axes);
    int t = 0;
    uint64_t tile = 0;
    if (!choose_tile(h, need, live, true, &tile, &t)) {
      err = "planner: op needs more target bits than the tile holds";
      return false;
    }
    fuse_kblocks(cur, tile, t);
    PlannedPass P;
    memset(&P, 0, sizeof(P));
    P.is_finish = false;
    PassParams& pp = P.pp;
    pp.t = t; pp.P = h->P; pp.nmax = h->nmax; pp.formalism = h->formalism;
    const uint64_t outer = live & ~tile;
    pp.outer_bits = __builtin_popcountll(outer);
    pp.n_tile_segs = build_segs(tile, pp.tile_segs, 16);
    pp.n_outer_segs = build_segs(outer, pp.outer_segs, 24);
    if (pp.n_tile_segs < 0 || pp.n_outer_segs < 0) { err = "planner: too many bit segments"; return false; }
    if (pp.n_tile_segs == 0) { pp.n_tile_segs = 1; pp.tile_segs[0] = Seg{0, 0, 0, 0}; }
    pp.seed = h->seed; pp.shot_offset = h->shot_offset; pp.nthreads = threads_for(h, t);
    for (int a = 0; a < 24; ++a) { pp.ax_col[a] = -1; pp.ax_row[a] = -1; }
    if (h->formalism == DQE_DM)
      for (int a = 0; a < h->nmax; ++a) {
        if ((tile >> a) & 1ull) pp.ax_col[a] = (int8_t)local_of(tile, a);
        if ((tile >> (a + h->nmax)) & 1ull) pp.ax_row[a] = (int8_t)local_of(tile, a + h->nmax);
      }
    for (int a = 0; a < h->nmax; ++a) pp.nfree_cols += pp.ax_col[a] >= 0;
    // diagonal enumeration tables for the probability ops (dm_diag_partials)
    for (int a = 0; a < 24; ++a) pp.ax_ord[a] = -1;
    for (int k = 0; k < 16; ++k) { pp.dg_mask[k] = 0; pp.dg_req[k] = -1; }
    if (h->formalism == DQE_DM) {
      int k = 0;
      for (int a = 0; a < h->nmax; ++a) {
        const int cl = pp.ax_col[a], rl = pp.ax_row[a];
        if (cl >= 0) {
          if (k >= 16) { err = "planner: too many column bits in a tile"; return false; }
          pp.ax_ord[a] = (int8_t)k;
          pp.dg_mask[k] = (1u << cl) | (rl >= 0 ? (1u << rl) : 0u);
          pp.dg_req[k] = rl >= 0 ? (int8_t)0 : (int8_t)(a + h->nmax);
          if (rl < 0) pp.dg_req_mask |= 1u << k;
          ++k;
        } else if (rl >= 0) {
          if (pp.dg_nfix >= 16) { err = "planner: too many row-only axes in a tile"; return false; }
          pp.dg_fix_src[pp.dg_nfix] = (int8_t)a; pp.dg_fix_dst[pp.dg_nfix] = (int8_t)rl;
          ++pp.dg_nfix;
        } else {
          pp.dg_eq_mask |= 1u << a;
        }
      }
    }
    {  // TMA bulk path: low tile segment contiguous from physical bit 0, runs of >= 128 B, <= 128 runs
      const int c0 = pp.tile_segs[0].len;
      pp.use_tma = h->tma && pp.tile_segs[0].dst == 0 && c0 >= 3 && (t - c0) <= 7;
      if (pp.use_tma) {
        const uint32_t nruns = 1u << (t - c0);
        for (uint32_t r = 0; r < nruns; ++r) {
          uint64_t off = 0;
          for (int sgi = 1; sgi < pp.n_tile_segs; ++sgi) {
            const Seg& sg = pp.tile_segs[sgi];
            off |= ((((uint64_t)r << c0) >> sg.src) & ((1ull << sg.len) - 1ull)) << sg.dst;
          }
          pp.run_hi[r] = (uint32_t)(off >> c0);
        }
      }
    }
    P.op_begin = dev_ops.size();
    P.const_begin = dev_consts.size();

    // ---- group ops by qubit pair (dm): an op joins the earliest group at or after its last
    //      dependency (an item sharing any bit with it) that still spans <= 2 axes
    std::vector<Item> items;
    const bool grouping = h->formalism == DQE_DM && h->group && t >= 4;
    int cl_item = -1;        // classical item of the current span
    int span_begin = 0;      // first item index of the current span (just behind the last measurement / fence)
    for (const HostOp* ho : cur) {
      if (ho->kind == 3) {   // classical fence: later classical instructions stay behind everything queued so far
        cl_item = -1;
        span_begin = (int)items.size();
        for (Item& it : items) it.touched = ~0ull;
        continue;
      }
      if (ho->kind == 2) {
        // Classical instructions read creg bits of EARLIER measurements and registers written by earlier
        // classical instructions only, and the registers / bits they write are not read by anything queued
        // before them in this span (dqe_classical_op contract): they run at the start of the span, so that
        // the quantum ops around them still group freely.
        if (cl_item < 0) {
          Item it;
          it.group = false; it.classical = true; it.nax = 0; it.ax[0] = it.ax[1] = -1; it.touched = 0;
          items.insert(items.begin() + span_begin, it);
          cl_item = span_begin;
        }
        items[cl_item].ops.push_back(ho);
        pp.uses_treg = 1;
        if (ho->ci.op >= CL_BAND) pp.writes_creg = 1; else pp.writes_treg = 1;
        continue;
      }
      if (ho->m.type == OP_DYN1 || (ho->m.type == OP_PAULI_SAMPLED && ho->m.flags >= 3 && ho->m.flags <= 5)) pp.uses_treg = 1;
      uint64_t tb = touched_bits(*ho);
      if (ho->m.type == OP_PROBS || ho->m.type == OP_MEASURE) tb = ~0ull;
      int ax[2], nax = 0;
      const bool can = grouping && group_axes(ho, tile, ax, &nax);
      int dep = -1;
      for (int k = (int)items.size() - 1; k >= 0; --k)
        if (items[k].touched & tb) { dep = k; break; }
      bool placed = false;
      if (can) {
        for (int k = std::max(dep, 0); k < (int)items.size() && !placed; ++k) {
          Item& it = items[k];
          if (!it.group) continue;
          int u[4], nu = 0;
          for (int i = 0; i < it.nax; ++i) u[nu++] = it.ax[i];
          for (int i = 0; i < nax; ++i) {
            bool seen = false;
            for (int j = 0; j < nu; ++j) seen |= u[j] == ax[i];
            if (!seen) u[nu++] = ax[i];
          }
          if (nu > 2) continue;
          it.nax = nu; it.ax[0] = u[0]; if (nu > 1) it.ax[1] = u[1];
          it.ops.push_back(ho); it.touched |= tb;
          placed = true;
        }
      }
      if (!placed) {
        Item it;
        it.group = can; it.classical = false; it.ops.push_back(ho); it.nax = nax; it.ax[0] = nax > 0 ? ax[0] : -1;
        it.ax[1] = nax > 1 ? ax[1] : -1; it.touched = tb;
        items.push_back(it);
      }
      if (ho->m.type == OP_PROBS || ho->m.type == OP_MEASURE) {   // a measurement opens the next span
        cl_item = -1;
        span_begin = (int)items.size();
      }
    }

    // ---- emit
    bool writes = false;
    const int n = h->nmax;
    std::vector<const HostOp*> kseq;   // ket: ops between classical blocks, emitted with block sweeps
    for (Item& it : items) {
      if (h->formalism == DQE_KET) {
        if (!it.classical) { for (const HostOp* ho : it.ops) kseq.push_back(ho); continue; }
        emit_ket_seq(kseq, tile, t, pp, P.const_begin, writes);
        kseq.clear();
      }
      if (it.classical) {
        emit_classical(it.ops, P.const_begin);
        continue;
      }
      bool as_group = it.group && (it.ops.size() >= 2 ||
                                   (h->cxperm && it.nax == 2 && hoistable_x(it.ops[0]) && x_ctrl_in_group(it.ops[0], tile)));
      int A = it.ax[0], B = it.nax > 1 ? it.ax[1] : -1;
      if (as_group && B < 0) {  // single-axis group: borrow any other axis fully inside the tile
        for (int a = 0; a < n && B < 0; ++a)
          if (a != A && ((tile >> a) & 1ull) && ((tile >> (a + n)) & 1ull)) B = a;
        if (B < 0) as_group = false;
      }
      if (as_group) {
        // for two-qubit matrices the first listed axis of the op is the MSB; keep A = that axis if any
        MicroOp g;
        memset(&g, 0, sizeof(g));
        g.type = OP_GROUP; g.pred_bit = -1; g.nb = 4;
        g.b[0] = (uint8_t)local_of(tile, A + n); g.b[1] = (uint8_t)local_of(tile, B + n);
        g.b[2] = (uint8_t)local_of(tile, A); g.b[3] = (uint8_t)local_of(tile, B);
        for (int i = 0; i < 4; ++i) g.b[4 + i] = g.b[i];
        std::sort(g.b + 4, g.b + 8);
        // X / CX is a permutation of the pair's 16 elements: when it opens or closes the group it rides in
        // the load / store addresses of the sweep (GPerm).  The two-qubit depolarising channel on the same
        // pair commutes with it (unitary covariance), so "CX, DEPOL2" is reordered to put the CX last.
        std::vector<const HostOp*> ops = it.ops;
        if (h->cxperm) {
          for (size_t k = 0; k + 1 < ops.size(); ++k)
            if (hoistable_x(ops[k]) && x_ctrl_in_group(ops[k], tile) && ops[k + 1]->m.type == OP_DEPOL2 &&
                ops[k + 1]->m.pred_bit < 0)
              std::swap(ops[k], ops[k + 1]);
        }
        const HostOp* head = nullptr;
        const HostOp* tail = nullptr;
        if (h->cxperm && !ops.empty() && hoistable_x(ops.front())) { head = ops.front(); ops.erase(ops.begin()); }
        if (h->cxperm && !ops.empty() && hoistable_x(ops.back())) { tail = ops.back(); ops.pop_back(); }
        if (head) {
          const MicroOp x = make_sub(head, A, B, tile, P.const_begin);   // flags: role, in_r (2), in_c (4), 16
          g.flags |= 1 | ((x.flags & 7) << 1);
          g.octrl = x.octrl; g.octrl2 = x.octrl2;
        }
        if (tail) {
          const MicroOp x = make_sub(tail, A, B, tile, P.const_begin);
          g.flags |= 16 | ((x.flags & 7) << 5);
          uint64_t* tc = reinterpret_cast<uint64_t*>(&g.p[0]);
          tc[0] = x.octrl; tc[1] = x.octrl2;
        }
        const size_t hdr = dev_ops.size();
        dev_ops.push_back(g);
        for (const HostOp* ho : ops) emit_sub(ho, A, B, tile, P.const_begin);
        dev_ops[hdr].aux = (int)(dev_ops.size() - hdr - 1);
        dev_ops[hdr].c_off = dev_ops[hdr].aux;   // the interpreter skips the sub-ops from the dispatch header alone
        writes = true;
      } else {
        for (const HostOp* ho : it.ops) {
          if ((ho->m.type != OP_PROBS && ho->m.type != OP_MEASURE) || ho->m.nb == 4) writes = true;   // (fused measurements write)
          if (ho->m.type == OP_MEASURE) { pp.has_measure = 1; pp.writes_creg = 1; }
          dev_ops.push_back(localise(ho, tile, P.const_begin));
        }
      }
    }
    if (h->formalism == DQE_KET) emit_ket_seq(kseq, tile, t, pp, P.const_begin, writes);
    // a classical block right behind a fused measurement runs inside it (op_measure): no dispatch, no barrier of its own
    for (size_t i = P.op_begin; i + 1 < dev_ops.size(); ++i)
      if (dev_ops[i].type == OP_MEASURE && dev_ops[i + 1].type == OP_CLASSICAL) dev_ops[i].lctrl2 |= kMeasRunsClassical;
    pp.nops = (int)(dev_ops.size() - P.op_begin);
    pp.nconsts = (int)(dev_consts.size() - P.const_begin);
    if (pp.nops > kMaxOpsPerPass || pp.nconsts > max_consts(h->formalism)) {
      err = "planner: pass exceeds the micro-op / constant budget";
      overflow = true;
      return false;
    }
    P.classical_only = true;
    for (const HostOp* ho : cur) if (ho->kind != 2 && ho->kind != 3) P.classical_only = false;
    pp.store_back = writes ? 1 : 0;
    pp.ntiles = (uint64_t)h->batch << pp.outer_bits;
    if (h->persistent) {
      // persistent grid + double buffering: as many CTAs as fit the machine (<= 512 threads per SM)
      pp.nbuf = pp.use_tma ? 2 : 1;
      P.smem = (sizeof(double2) << t) * pp.nbuf + sizeof(double2) * (size_t)pp.nconsts + sizeof(MicroOp) * (size_t)pp.nops;
      const size_t per_cta = P.smem + 8 * 1024;
      int per_sm = (int)std::min<size_t>((size_t)(227 * 1024) / per_cta, (size_t)(512 / h->threads));
      per_sm = std::max(1, std::min(per_sm, h->ctas_per_sm > 0 ? h->ctas_per_sm : 16));
      P.grid = std::min<uint64_t>(pp.ntiles, (uint64_t)h->num_sms * per_sm);
    } else {
      // one tile per CTA; overlap of load / compute / store comes from 4+ resident CTAs per SM
      pp.nbuf = 1;
      P.smem = (sizeof(double2) << t) + sizeof(double2) * (size_t)pp.nconsts +   // tile + the pass's constants
               sizeof(MicroOp) * (size_t)pp.nops;                                  // + its micro-ops
      P.grid = pp.ntiles;
      if (P.grid > 0x7fffffffull) { err = "grid too large: enable the persistent option"; return false; }
    }
    P.amps = 1ull << __builtin_popcountll(live);
    passes.push_back(P);
    cur.clear();
    return true;
  }

  // every pass of this handle has at most 8 tiles per shot (cluster size limit) when the whole
  // register is at most 3 bits wider than a tile
  bool cluster_mode() const { return cat_cluster_mode(h); }

  bool plan() {
    std::vector<const HostOp*> cur;
    uint64_t live_now = h->planned_live, need = 0, live = live_now, live0 = live_now;
    int nconst = 0, nslots = 0, ncl = 0;
    for (const HostOp& op : h->queue) {
      if (cur.empty()) cur_first = (size_t)(&op - h->queue.data());
      if (op.kind == 1 && cluster_mode() && !cur.empty() && cur.back()->m.type == OP_PROBS) {
        // small registers: all tiles of a shot fit one thread-block cluster, so the measurement is a
        // micro-op (DSMEM sum + on-chip draw) and the pass goes on
        HostOp mo = *cur.back();
        mo.m.type = OP_MEASURE;
        mo.m.flags = op.creg_bit + 1;
        mo.m.lctrl = (uint32_t)(op.forced + 1);
        mo.m.p[0] = op.flip;
        mo.m.octrl = op.draw;
        mo.live_after = op.live_after;
        fused.push_back(mo);
        cur.back() = &fused.back();
        live_now = op.live_after;
        continue;
      }
      if (op.kind == 1) {
        // the PROBS op is the tail of `cur`
        if (!close(cur, live0)) return false;
        const PlannedPass& last = passes.back();
        PlannedPass F;
        memset(&F, 0, sizeof(F));
        F.is_finish = true;
        F.fp.seed = h->seed; F.fp.shot_offset = h->shot_offset; F.fp.draw = op.draw;
        F.fp.batch = h->batch;
        F.fp.tiles_per_shot = 1ull << last.pp.outer_bits;
        F.fp.flip = op.flip; F.fp.forced = op.forced; F.fp.creg_bit = op.creg_bit;
        F.fp.is_dm = h->formalism == DQE_DM;
        passes.push_back(F);
        need = 0; live = live_now; live0 = live_now; nconst = 0; nslots = 0; ncl = 0;
        continue;
      }
      uint64_t need2 = need | op.need;
      uint64_t live2 = live | op.live_after | op.need;
      int t = 0;
      uint64_t tile = 0;
      // Running estimate of the pass's micro-program (close() emits the real one and splits the pass if the
      // estimate was too optimistic): an op runs as max(1, nf) sub-ops, roughly every third op opens a group
      // (one header), classical instructions pack ten to a micro-op with up to two part-filled micro-ops
      // (serial + independent) per span.
      int op_slots = std::max(1, op.nf), op_consts = op.nconst;
      if (op.kind >= 2) {
        op_consts = 0;
        op_slots = 0;
        if (op.kind == 2) {
          ++ncl;
          if (ncl % 2 == 0 || op.ci.op == CL_QEXP2 || op.ci.op == CL_GEOM) op_consts = 1;
          if (ncl % 8 == 1) op_slots = 1;
        }
      } else {
        if (op.m.type == OP_MEASURE || op.m.type == OP_PROBS) op_slots += 2;
        if (op.m.type == OP_DIAG && h->formalism == DQE_DM) op_consts = std::max(op_consts, 8);
        if (op.nf > 0) op_consts = std::max(op_consts, 4 * op.nf);
        if (++nq3 % 3 == 0) op_slots += 1;
      }
      if (cur.empty()) { nslots = 3; ncl = 0; nconst = 0; }
      // Small density matrices (one 64-thread CTA per <= 2^10 tile): a pass may stage only as many constants as leave
      // room for EIGHT resident CTAs per SM (registers allow exactly eight) -- tile + constants + ~20 micro-ops + the
      // classical registers in use + 768 B static + 1 KiB reserved <= 228 KiB / 8.  One whole remote gate fits (416
      // constants, 544 with two idle channels); a second one would cost every CTA of the pass its eighth neighbour.
      auto const_cap = [&](int tt) {
        int cap = max_consts(h->formalism);
        if (h->formalism == DQE_DM && h->tile_auto && tt <= 10) {
          const int regs = std::min(kTimeRegs, (h->max_reg + 2) & ~1);
          const int room = (228 * 1024) / 8 - 1024 - 256 - (16 << tt) - 20 * (int)sizeof(MicroOp) - 8 * regs;
          cap = std::max(64, std::min(std::min(cap, room / 16), 560));   // (560: one gate per pass on smaller tiles too --
                                                                         //  two per pass measured 2 % slower on QFT-5)
        }
        return cap;
      };
      const bool fits = cur.empty() || (h->fuse && nslots + op_slots <= h->pass_ops &&
                                        choose_tile(h, need2, live2 | phys_live(h, h->pin_axes), false, &tile, &t) &&
                                        nconst + op_consts <= const_cap(t));
      if (!fits) {
        if (!close(cur, live0)) return false;
        cur_first = (size_t)(&op - h->queue.data());
        need2 = op.need;
        live0 = live_now;
        live2 = live_now | op.live_after | op.need;
        nconst = 0; nslots = 3; ncl = op.kind == 2 ? 1 : 0;
        if (op.kind == 2) op_slots = 1;
      }
      nconst += op_consts;
      nslots += op_slots;
      cur.push_back(&op);
      need = need2; live = live2;
      live_now = op.live_after;
    }
    if (!close(cur, live0)) return false;
    h->planned_live = live_now;
    return true;
  }
  int nq3 = 0;
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Tensor map of one pass: the tile as a single TMA box (see PassParams::tmap).  Leaves tm_rank = 0 (bulk-run
// path) when the tile is one or two runs anyway, when the shape does not fit 5 dimensions, or when the
// driver refuses the descriptor.
void build_tensor_map(dqe_t* h, PassParams& pp) {
  pp.tm_rank = 0;
  if (!h->tmap || !pp.use_tma || h->persistent) return;
  const int c0 = pp.tile_segs[0].len;
  if (c0 > 7 || (pp.t - c0) < 2) return;
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return;
  struct Dim { uint64_t stride, size; uint32_t box; bool base; };
  std::vector<Dim> dims;
  for (int i = 1; i < pp.n_tile_segs; ++i) {
    const Seg& sg = pp.tile_segs[i];
    if (sg.len > 8) return;
    dims.push_back(Dim{16ull << sg.dst, 1ull << sg.len, 1u << sg.len, false});
  }
  const int lb = pp.n_outer_segs > 0 && pp.outer_bits > 0 ? pp.outer_segs[0].dst : pp.P;
  const uint64_t total = (uint64_t)h->batch << pp.P;
  const uint64_t nbase = total >> lb;
  if (nbase == 0 || nbase > (1ull << 31)) return;
  dims.push_back(Dim{16ull << lb, nbase, 1u, true});
  std::sort(dims.begin(), dims.end(), [](const Dim& a, const Dim& b) { return a.stride < b.stride; });
  const int rank = 1 + (int)dims.size();
  if (rank < 3 || rank > 5) return;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  gdim[0] = 2ull << c0; box[0] = 2u << c0;   // complex128 = 2 x FLOAT64
  int base_dim = 0;
  for (int i = 0; i < (int)dims.size(); ++i) {
    if (dims[i].stride >= (1ull << 40)) return;
    gdim[i + 1] = dims[i].size; box[i + 1] = dims[i].box; gstr[i] = dims[i].stride;
    if (dims[i].base) base_dim = i + 1;
  }
  const CUresult rc = enc(&pp.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, (void*)pp.state, gdim, gstr, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { h->tmap_refused++; return; }
  pp.tm_rank = rank; pp.tm_base_dim = base_dim; pp.tm_shift = lb;
  // L2 prefetch distance: one wave of resident CTAs ahead (-1 = automatic: SMs x CTAs per SM)
  pp.prefetch_dist = h->prefetch >= 0 ? h->prefetch : h->num_sms * (pp.nthreads <= 64 ? 8 : 4);
}

struct Planner;
int launch_passes(dqe_t* h, Planner& pl);
int do_flush(dqe_t* h, bool partial);
inline int do_flush(dqe_t* h) { return do_flush(h, false); }

int ensure_partials(dqe_t* h, size_t n) {
  if (n <= h->partials_cap) return 0;
  if (h->partials) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->partials)); h->partials = nullptr; }
  CK(cudaMalloc(&h->partials, n * sizeof(double)));
  h->partials_cap = n;
  return 0;
}

// partial = streaming flush: everything but the last (still open) pass is launched, its ops stay queued, so that
// the device works on the head of a long program while the host is still enqueueing its tail
int do_flush(dqe_t* h, bool partial) {
  if (int rc = check_ready(h)) return rc;
  if (h->queue.empty()) return 0;
  ladderize(h);
  fuse_cat_measure(h);
  if (getenv("DQE_DUMP_QUEUE")) {   // the queue the planner sees, with constants (debugging / host-side tests)
    for (const HostOp& o : h->queue) {
      if (o.kind != 0) { fprintf(stderr, "Q kind=%d\n", o.kind); continue; }
      fprintf(stderr, "Q type=%d pred=%d nb=%u aux=%d pctrl=%llu b=", o.m.type, o.m.pred_bit, o.m.nb, o.m.aux,
              (unsigned long long)o.pctrl);
      for (uint32_t i = 0; i < o.m.nb; ++i) fprintf(stderr, "%s%d", i ? "," : "", (int)o.m.b[i]);
      if (o.m.type == OP_LADDER) {
        fprintf(stderr, " leaves=");
        for (int j = 0; j < o.m.aux; ++j) fprintf(stderr, "%s%d", j ? "," : "", (int)(reinterpret_cast<const uint8_t*>(&o.m) + kLadderLeafOff)[j]);
      }
      fprintf(stderr, " c=");
      for (int i = 0; i < o.nconst; ++i) fprintf(stderr, "%s%.17g,%.17g", i ? ";" : "", h->cpool[o.m.c_off + i].x, h->cpool[o.m.c_off + i].y);
      fprintf(stderr, "\n");
    }
  }
  const uint64_t live_head = h->planned_live;
  Planner pl;
  pl.h = h;
  if (!pl.plan()) {
    h->queue.clear(); h->cpool.clear();
    return fail(h, DQE_EINVAL, pl.err);
  }
  if (getenv("DQE_DUMP_PLAN")) {
    static const char* names[] = {"?", "MAT1", "MAT2", "DIAG", "PAULI1", "DEPOL2", "PROBS", "COLLAPSE", "TRACE",
                                  "INJECT", "PSAMPLED", "MAT1X2", "GROUP", "MEASURE", "XPAT", "DYN1", "CL", "KBLOCK", "LADDER", "KGROUP", "SWAP"};
    static const char* gnames[] = {"?", "gPAULI1", "-", "gMAT1X2", "gDEPOL2", "-", "-", "gCOLLAPSE",
                                   "gTRACE", "gINJECT", "gDIAG", "gXPAT", "-", "-", "-", "gDYN1"};
    for (PlannedPass& P : pl.passes) {
      if (P.is_finish) { fprintf(stderr, "  finish\n"); continue; }
      fprintf(stderr, "pass t=%d outer=%d nops=%d:", P.pp.t, P.pp.outer_bits, P.pp.nops);
      for (int i = 0; i < P.pp.nops; ++i) {
        const MicroOp& m = pl.dev_ops[P.op_begin + i];
        if (m.type > 100) { fprintf(stderr, " %s%s.%d", gnames[m.type - 100], m.pred_bit >= 0 ? "?" : "", m.flags & 1); continue; }
        if (m.type == OP_GROUP && (m.flags & 17))
          fprintf(stderr, " [%s%s]", (m.flags & 1) ? "x>" : "", (m.flags & 16) ? ">x" : "");
        if (m.type == OP_CLASSICAL) { fprintf(stderr, " CL%s%d", m.c_off ? "p" : "", m.flags & 0xff); continue; }
        if (m.type == OP_LADDER) {
          fprintf(stderr, " LADDER%s(%d;%d+%d)", m.pred_bit >= 0 ? "?" : "", (int)m.b[0], m.flags & 0xff, (m.flags >> 8) & 0xff);
          continue;
        }
        fprintf(stderr, " %s%s(", names[m.type], m.pred_bit >= 0 ? "?" : "");
        for (uint32_t b = 0; b < m.nb; ++b) fprintf(stderr, "%s%d", b ? "," : "", (int)m.b[b]);
        fprintf(stderr, ")");
      }
      fprintf(stderr, "\n");
    }
  }
  if (getenv("DQE_DUMP_SMEM"))   // shared-memory budget of every pass (debugging / occupancy planning)
    for (const PlannedPass& P : pl.passes)
      if (!P.is_finish) fprintf(stderr, "smem t=%d nops=%d nconsts=%d dyn=%zu max_reg=%d\n", P.pp.t, P.pp.nops, P.pp.nconsts, P.smem, h->max_reg);
  size_t keep_from = h->queue.size();
  uint64_t keep_live = 0;
  if (partial) {
    h->planned_live = live_head;
    if (pl.passes.size() < 2 || pl.passes.back().is_finish) return 0;   // nothing closed yet (or nothing open)
    keep_from = pl.passes.back().first_op;
    keep_live = pl.passes.back().live_before;
    while (!pl.passes.empty() && !pl.passes.back().is_finish && pl.passes.back().first_op == keep_from) pl.passes.pop_back();
    if (pl.passes.empty() || keep_from == 0) return 0;
  }
  auto drop_done = [&]() {
    if (keep_from >= h->queue.size()) {
      h->queue.clear(); h->cpool.clear();
      h->stream_next = (size_t)h->stream_ops;   // empty queue: the next streaming point counts from here
      h->planned_live = phys_live(h, h->live_axes);
    } else {   // the open tail stays (its constants too: c_off indexes the pool)
      h->queue.erase(h->queue.begin(), h->queue.begin() + keep_from);
      h->planned_live = keep_live;
    }
  };
  if (h->device < 0) {  // plan-only: account and drop
    for (PlannedPass& P : pl.passes) {
      if (P.is_finish) { h->stats[2]++; h->stats[5]++; continue; }
      h->stats[0]++; h->stats[2]++; h->stats[1] += P.pp.nops;
      h->stats[3] += (int64_t)(P.amps * (uint64_t)h->batch);
    }
    h->stats[6] += (int64_t)(pl.dev_ops.size() * sizeof(MicroOp) + pl.dev_consts.size() * sizeof(double2));
    drop_done();
    return 0;
  }
  CK(cudaSetDevice(h->device));
  // upload micro-ops + constants
  if (pl.dev_ops.size() > h->d_ops_cap) {
    if (h->d_ops) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->d_ops)); }
    h->d_ops_cap = std::max<size_t>(pl.dev_ops.size() * 2, 1024);
    CK(cudaMalloc(&h->d_ops, h->d_ops_cap * sizeof(MicroOp)));
  }
  if (pl.dev_consts.size() > h->d_consts_cap) {
    if (h->d_consts) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->d_consts)); }
    h->d_consts_cap = std::max<size_t>(pl.dev_consts.size() * 2, 4096);
    CK(cudaMalloc(&h->d_consts, h->d_consts_cap * sizeof(double2)));
  }
  h->stats[6] += (int64_t)(pl.dev_ops.size() * sizeof(MicroOp) + pl.dev_consts.size() * sizeof(double2));
  if (!pl.dev_ops.empty())
    CK(cudaMemcpyAsync(h->d_ops, pl.dev_ops.data(), pl.dev_ops.size() * sizeof(MicroOp),
                       cudaMemcpyHostToDevice, h->stream));
  if (!pl.dev_consts.empty())
    CK(cudaMemcpyAsync(h->d_consts, pl.dev_consts.data(), pl.dev_consts.size() * sizeof(double2),
                       cudaMemcpyHostToDevice, h->stream));
  // partials: sized for the widest grid of a pass that ends in PROBS
  size_t need_part = 0;
  for (const PlannedPass& P : pl.passes)
    if (!P.is_finish) need_part = std::max<size_t>(need_part, (size_t)P.pp.ntiles * 2);
  if (int rc = ensure_partials(h, need_part)) return rc;
  // validate every pass before the first launch: a failure must not leave half of the queue applied
  for (const PlannedPass& P : pl.passes) {
    if (P.is_finish) continue;
    if (P.pp.ntiles > (1ull << 40)) { h->queue.clear(); h->cpool.clear(); return fail(h, DQE_EUNSUPPORTED, "too many tiles"); }
    if (P.pp.uses_treg && !h->treg_buf[0]) {
      h->queue.clear(); h->cpool.clear();
      return fail(h, DQE_EINVAL, "classical registers used before any classical instruction");
    }
  }
  const bool whole_program = h->fresh;   // first launch since the register was reinitialised
  h->fresh = false;
  const int rc_launch = launch_passes(h, pl);
  // whatever happened, the queued ops are gone: either they ran, or a CUDA error made the handle unusable
  // (sticky); re-applying ops that already ran on the next flush would be worse than either
  if (rc_launch) keep_from = h->queue.size();
  drop_done();
  if (rc_launch) return rc_launch;
  if (h->plan_cache && !partial && whole_program) {   // a whole program from a fresh register: replayable
    if (!h->cached) h->cached = new CachedPlan();
    h->cached->passes = pl.passes;
    h->cached->dev_ops = pl.dev_ops;
    h->cached->dev_consts = pl.dev_consts;
    h->cached->live_axes = h->live_axes; h->cached->planned_live = h->planned_live; h->cached->draw = h->draw;
    h->cached->api_ops = h->stats[4] - h->api_mark;
  } else if (h->cached && !h->plan_cache) {
    delete h->cached; h->cached = nullptr;
  }
  if (getenv("DQE_DEBUG"))
    fprintf(stderr, "[dqe] flush: %zu passes (%lld chained behind another so far), tensor maps used %lld refused %lld (cumulative)\n",
            pl.passes.size(), (long long)h->chained_links, (long long)h->tmap_used, (long long)h->tmap_refused);
  return 0;
}

// Two passes may share a launch when the second finds the tile exactly where the first leaves it: one tile per
// shot (no cluster, no other CTA of the shot to wait for), the same tile bits in the same order, the same CTA shape.
bool same_tile_layout(const PassParams& a, const PassParams& b) {
  return a.outer_bits == 0 && b.outer_bits == 0 && a.t == b.t && a.P == b.P && a.nthreads == b.nthreads &&
         a.use_tma == b.use_tma && a.nbuf == b.nbuf && a.ntiles == b.ntiles && a.n_tile_segs == b.n_tile_segs &&
         a.n_outer_segs == b.n_outer_segs && a.nfree_cols == b.nfree_cols && a.dg_nfix == b.dg_nfix &&
         a.dg_req_mask == b.dg_req_mask && a.dg_eq_mask == b.dg_eq_mask &&
         !memcmp(a.tile_segs, b.tile_segs, sizeof(a.tile_segs)) && !memcmp(a.outer_segs, b.outer_segs, sizeof(a.outer_segs)) &&
         !memcmp(a.ax_col, b.ax_col, sizeof(a.ax_col)) && !memcmp(a.ax_row, b.ax_row, sizeof(a.ax_row)) &&
         !memcmp(a.ax_ord, b.ax_ord, sizeof(a.ax_ord)) && !memcmp(a.dg_mask, b.dg_mask, sizeof(a.dg_mask)) &&
         !memcmp(a.dg_req, b.dg_req, sizeof(a.dg_req)) && !memcmp(a.dg_fix_src, b.dg_fix_src, sizeof(a.dg_fix_src)) &&
         !memcmp(a.dg_fix_dst, b.dg_fix_dst, sizeof(a.dg_fix_dst)) && !memcmp(a.run_hi, b.run_hi, sizeof(a.run_hi));
}

// resident CTAs per SM as far as shared memory decides (228 KiB per SM, 1 KiB reserved per CTA, ~6.3 KiB static)
int ctas_per_sm_by_smem(size_t dyn) { return (int)((size_t)(228 * 1024) / (dyn + 512 + 1024)); }

int launch_passes(dqe_t* h, Planner& pl) {
  // pass chains: chain_len[i] = number of passes launched together with pass i as their head (0: rides in a chain)
  const size_t np = pl.passes.size();
  std::vector<int> chain_len(np, 1);
  h->links_host.clear();
  std::vector<size_t> link_at(np, 0);
  if (h->chain && !h->persistent && h->device >= 0) {
    for (size_t i = 0; i < np;) {
      const PlannedPass& A = pl.passes[i];
      size_t j = i + 1;
      if (!A.is_finish && !A.classical_only && A.pp.store_back)
        while (j < np && !pl.passes[j].is_finish && !pl.passes[j].classical_only && pl.passes[j].pp.store_back &&
               (h->chain_max <= 0 || (int)(j - i) < h->chain_max) && same_tile_layout(A.pp, pl.passes[j].pp) &&
               ctas_per_sm_by_smem(pl.passes[j].smem) == ctas_per_sm_by_smem(A.smem)) ++j;   // (one fat link must not cost the whole chain a resident CTA)
      if (j - i > 1) {
        chain_len[i] = (int)(j - i);
        link_at[i] = h->links_host.size();
        for (size_t k = i; k < j; ++k) {
          chain_len[k] = k == i ? (int)(j - i) : 0;
          const PlannedPass& Q = pl.passes[k];
          h->links_host.push_back(make_uint4((unsigned)Q.op_begin, (unsigned)Q.pp.nops, (unsigned)Q.const_begin, (unsigned)Q.pp.nconsts));
        }
      }
      i = j;
    }
    if (!h->links_host.empty()) {
      if (h->links_host.size() > h->d_links_cap) {
        if (h->d_links) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->d_links)); h->d_links = nullptr; }
        h->d_links_cap = std::max<size_t>(h->links_host.size() * 2, 256);
        CK(cudaMalloc(&h->d_links, h->d_links_cap * sizeof(uint4)));
      }
      CK(cudaMemcpyAsync(h->d_links, h->links_host.data(), h->links_host.size() * sizeof(uint4), cudaMemcpyHostToDevice, h->stream));
      h->stats[6] += (int64_t)(h->links_host.size() * sizeof(uint4));
    }
  }
  for (size_t pi = 0; pi < np; ++pi) {
    PlannedPass& P = pl.passes[pi];
    if (chain_len[pi] == 0) continue;   // ran inside the launch of its chain's head
    if (P.is_finish) {
      P.fp.partials = h->partials; P.fp.rec = cur_rec(h); P.fp.creg = cur_creg(h);
      P.fp.olog = (h->olog && (int64_t)P.fp.draw < h->olog_draws) ? h->olog : nullptr;
      const unsigned grid = (unsigned)((h->batch + 3) / 4);
      k_finish_measure<<<grid, 128, 0, h->stream>>>(P.fp);
      h->stats[2]++; h->stats[5]++;
      continue;
    }
    if (P.classical_only) {
      if (P.pp.nops == 0) continue;
      ClassicalParams cp;
      memset(&cp, 0, sizeof(cp));
      cp.ops = h->d_ops + P.op_begin;
      cp.imm = reinterpret_cast<const double*>(h->d_consts + P.const_begin);
      cp.creg = cur_creg(h); cp.treg = h->treg_buf[h->tpar]; cp.batch = h->batch; cp.nops = P.pp.nops;
      cp.seed = h->seed; cp.shot_offset = h->shot_offset;
      k_classical<<<(unsigned)((h->batch + 127) / 128), 128, 0, h->stream>>>(cp);
      h->stats[2]++;
      continue;
    }
    P.pp.state = h->state; P.pp.partials = h->partials;
    P.pp.creg = cur_creg(h); P.pp.rec = cur_rec(h);
    P.pp.batch = h->batch;
    P.pp.olog = h->olog;   // draws beyond the log's capacity are rejected when they are queued
    P.pp.ops = h->d_ops + P.op_begin; P.pp.consts = h->d_consts + P.const_begin;
    P.pp.nlinks = 1; P.pp.links = nullptr; P.pp.ops_all = h->d_ops; P.pp.consts_all = h->d_consts;
    size_t smem = P.smem + (size_t)h->smem_pad;
    P.pp.cap_consts = P.pp.nconsts;
    int cap_ops = P.pp.nops;
    P.pp.nregs = 0;
    uint64_t amps = P.amps;
    int64_t nops_all = P.pp.nops;
    const int saved_flags[5] = {P.pp.has_measure, P.pp.writes_creg, P.pp.uses_treg, P.pp.writes_treg, 0};
    if (chain_len[pi] > 1) {
      // the head's launch carries the union of what its links do to the classical state, and the largest
      // constant area among them
      for (int k = 1; k < chain_len[pi]; ++k) {
        const PlannedPass& Q = pl.passes[pi + k];
        P.pp.has_measure |= Q.pp.has_measure; P.pp.writes_creg |= Q.pp.writes_creg;
        P.pp.uses_treg |= Q.pp.uses_treg; P.pp.writes_treg |= Q.pp.writes_treg;
        P.pp.cap_consts = std::max(P.pp.cap_consts, Q.pp.nconsts);
        cap_ops = std::max(cap_ops, Q.pp.nops);
        amps += Q.amps; nops_all += Q.pp.nops;
      }
      smem = (sizeof(double2) << P.pp.t) + sizeof(double2) * (size_t)P.pp.cap_consts + sizeof(MicroOp) * (size_t)cap_ops +
             (size_t)h->smem_pad;
      P.pp.nlinks = chain_len[pi];
      P.pp.links = h->d_links + link_at[pi];
      h->chained_links += chain_len[pi] - 1;
    }
    if (P.pp.writes_creg) {
      h->cpar ^= 1;
      P.pp.creg_out = cur_creg(h); P.pp.rec_out = cur_rec(h);
    }
    if (P.pp.uses_treg) {
      P.pp.nregs = std::min(kTimeRegs, (h->max_reg + 2) & ~1);
      P.pp.treg = h->treg_buf[h->tpar];
      if (P.pp.writes_treg) { h->tpar ^= 1; P.pp.treg_out = h->treg_buf[h->tpar]; }
    }
    P.pp.cap_ops = cap_ops;
    smem += sizeof(double) * (size_t)P.pp.nregs;   // the classical registers in use ride behind the micro-ops
    if (P.pp.has_measure && P.pp.outer_bits) smem += 32u << P.pp.outer_bits;   // ... and the measurement partials of the
                                                                               // shot's CTAs (a cluster of <= 16) behind them
    build_tensor_map(h, P.pp);
    if (P.pp.tm_rank) h->tmap_used++;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) {
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0, h->stream));
    }
    // one kernel per formalism (their micro-op sets are disjoint enough that sharing one
    // interpreter costs registers), 128 registers per thread (a 161- and a 187-register build of the same
    // code were measured slower: fewer resident CTAs).
    // Passes that carry a fused measurement launch as thread-block clusters: the 2^outer_bits tiles of
    // a shot form one cluster and exchange their probability partials through DSMEM.
    {
      const void* fn;
      int threads = P.pp.nthreads;
      const bool pers = h->persistent != 0;
#define DQE_KFN(M, B, F) (pers ? (const void*)k_tile_pass<M, B, F, true> : (const void*)k_tile_pass<M, B, F, false>)
      fn = h->formalism == DQE_KET ? DQE_KFN(256, 2, 0) : DQE_KFN(256, 2, 1);
#undef DQE_KFN
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)P.grid);
      cfg.blockDim = dim3((unsigned)threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = h->stream;
      cudaLaunchAttribute attr[1];
      if (P.pp.has_measure) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1u << P.pp.outer_bits;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
      }
      void* args[1] = {(void*)&P.pp};
      P.pp.tile0 = 0;
      if (P.pp.nlinks > 1 && h->chain_waves > 0) {
        // a chain keeps its CTAs alive for the whole program: the batch goes through it as chunks of whole waves, so
        // that the CTAs sharing an SM start together and stay in step (measured: their instruction fetch is shared)
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem));
        const uint64_t chunk = (uint64_t)std::max(1, occ) * (uint64_t)h->num_sms * (uint64_t)h->chain_waves;
        for (uint64_t t0 = 0; t0 < P.pp.ntiles; t0 += chunk) {
          P.pp.tile0 = t0;
          cfg.gridDim = dim3((unsigned)std::min<uint64_t>(chunk, P.pp.ntiles - t0));
          CK(cudaLaunchKernelExC(&cfg, fn, args));
          if (t0) h->stats[2]++;
        }
        P.pp.tile0 = 0;
      } else
      CK(cudaLaunchKernelExC(&cfg, fn, args));
    }
    if (h->timing) {
      CK(cudaEventRecord(e1, h->stream));
      h->ev_pool.push_back(std::make_pair(e0, e1));
      // a chain streams the state through HBM once, whatever the number of links (a fused run counts once)
      h->ev_amps.push_back(P.amps * (uint64_t)h->batch);
    }
    h->stats[0] += chain_len[pi]; h->stats[2]++;
    h->stats[1] += nops_all;
    h->stats[3] += (int64_t)(amps * (uint64_t)h->batch);
    // the plan may be launched again (dqe_replay): the head gets its own flags back
    P.pp.has_measure = saved_flags[0]; P.pp.writes_creg = saved_flags[1]; P.pp.uses_treg = saved_flags[2]; P.pp.writes_treg = saved_flags[3];
    P.pp.creg_out = nullptr; P.pp.rec_out = nullptr; P.pp.treg_out = nullptr;
  }
  CK(cudaGetLastError());
  return 0;
}

// read-out staging buffer kept on the handle (no cudaMalloc / cudaFree per read, nothing to leak on an error path)
int ensure_scratch(dqe_t* h, size_t bytes) {
  if (bytes <= h->scratch_cap) return 0;
  if (h->scratch) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->scratch)); h->scratch = nullptr; h->scratch_cap = 0; }
  const size_t cap = std::max<size_t>(bytes, 1 << 16);
  CK(cudaMalloc(&h->scratch, cap));
  h->scratch_cap = cap;
  return 0;
}

int collect_timing(dqe_t* h) {
  for (size_t i = 0; i < h->ev_pool.size(); ++i) {
    float ms = 0.f;
    CK(cudaEventSynchronize(h->ev_pool[i].second));
    CK(cudaEventElapsedTime(&ms, h->ev_pool[i].first, h->ev_pool[i].second));
    h->t_ms += ms; h->t_passes++; h->t_amps += (int64_t)h->ev_amps[i];
    if (getenv("DQE_PRINT_PASS_MS"))
      fprintf(stderr, "pass %zu: %.4f ms, %.1f GB/s\n", i, ms, 32.0 * (double)h->ev_amps[i] / (ms * 1e-3) / 1e9);
    cudaEventDestroy(h->ev_pool[i].first); cudaEventDestroy(h->ev_pool[i].second);
  }
  h->ev_pool.clear(); h->ev_amps.clear();
  return 0;
}

int build_readout(dqe_t* h, int k, const int* axes, ReadoutParams* rp) {
  if (k < 1 || k > 5) return fail(h, DQE_EUNSUPPORTED, "read-out supports 1..5 axes");
  memset(rp, 0, sizeof(*rp));
  uint64_t sel = 0;
  for (int i = 0; i < k; ++i) {
    if (int rc = check_live(h, axes[i], "read-out")) return rc;
    if ((sel >> axes[i]) & 1ull) return fail(h, DQE_EINVAL, "read-out: repeated axis");
    sel |= 1ull << axes[i];
    rp->axes[i] = axes[i];
  }
  const uint64_t rest = h->live_axes & ~sel;
  rp->state = h->state; rp->batch = h->batch; rp->P = h->P; rp->nmax = h->nmax;
  rp->formalism = h->formalism; rp->k = k;
  rp->rest_bits = __builtin_popcountll(rest);
  rp->n_rest_segs = build_segs(rest, rp->rest_segs, 24);
  if (rp->n_rest_segs < 0) return fail(h, DQE_EUNSUPPORTED, "read-out: too many segments");
  return 0;
}

}  // namespace

// =============================================================================================
extern "C" {

const char* dqe_version(void) { return "dqc-b200-engine 0.1 (sm_100a)"; }

const char* dqe_last_error(const dqe_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int dqe_create(dqe_t** out, int formalism, int max_qubits, int64_t batch, int device, void* ext_state,
               void* stream) {
  if (!out) return DQE_EINVAL;
  *out = nullptr;
  if (formalism != DQE_KET && formalism != DQE_DM) return fail(nullptr, DQE_EINVAL, "formalism must be 0 (ket) or 1 (dm)");
  const int P = formalism == DQE_KET ? max_qubits : 2 * max_qubits;
  if (max_qubits < 1 || P > 40 || (formalism == DQE_DM && max_qubits > 16))
    return fail(nullptr, DQE_EINVAL, "max_qubits out of range (ket <= 40, dm <= 16: 4^16 complex128 = 64 GiB)");
  if (batch < 1) return fail(nullptr, DQE_EINVAL, "batch must be >= 1");
  dqe_t* h = new dqe_engine();
  h->formalism = formalism; h->nmax = max_qubits; h->P = P; h->device = device; h->batch = batch;
  h->own_stream = false; h->own_state = false; h->state = nullptr;
  h->creg_buf[0] = h->creg_buf[1] = nullptr; h->rec_buf[0] = h->rec_buf[1] = nullptr;
  h->treg_buf[0] = h->treg_buf[1] = nullptr; h->cpar = 0; h->tpar = 0; h->max_reg = 0; h->olog = nullptr; h->olog_draws = 0;
  h->state_dirty = false;
  h->partials = nullptr; h->partials_cap = 0; h->d_ops = nullptr; h->d_ops_cap = 0;
  h->scratch = nullptr; h->scratch_cap = 0;
  h->d_consts = nullptr; h->d_consts_cap = 0; h->live_axes = 0; h->planned_live = 0; h->pin_axes = 0;
  h->seed = 0; h->shot_offset = 0; h->draw = 0; h->pred = -1; h->cuda_failed = 0; h->shot_stride = 1;
  h->tile_bits = P <= 14 ? 10 : 11; h->min_run_bits = 5; h->fuse = 1; h->merge = 1; h->group = 1; h->tma = 1; h->threads = P <= 14 ? 64 : 128; h->ctas_per_sm = 0; h->num_sms = 148; h->persistent = 0; h->chain = getenv("DQE_CHAIN") ? atoi(getenv("DQE_CHAIN")) : 0; h->chained_links = 0; h->chain_max = 0; h->chain_waves = 1; h->smem_pad = 0; h->d_links = nullptr; h->d_links_cap = 0; h->cluster = 1; h->cxperm = 1; h->factor = 1; h->tmap = 1; h->prefetch = -1; h->kblock = 1; h->kblocks_fused = 0; memset(h->peer_epoch, 0, sizeof(h->peer_epoch)); h->xev_ok = false; h->xev_pending = false; h->x_ms = 0; h->x_count = 0; h->x_bytes = 0;
  h->plan_cache = 0; h->cached = nullptr; h->api_mark = 0; h->fresh = true; h->pass_ops = kMaxOpsPerPass; h->tile_auto = (formalism == DQE_DM && P <= 14) ? 1 : 0; h->catfuse = getenv("DQE_CATFUSE") ? atoi(getenv("DQE_CATFUSE")) : 3; h->cat_fused = 0; h->ladder = 1; h->phase_tables = getenv("DQE_PHASE_TABLES") ? atoi(getenv("DQE_PHASE_TABLES")) : 0; h->phases_merged = 0; h->ladders_built = 0; h->kgroup = 1; h->kgroups_built = 0; h->stream_ops = 96; h->stream_next = 96; h->stream_min_bits = 22; h->stream_flushes = 0; h->tmap_used = 0; h->tmap_refused = 0; h->timing = false;
  h->t_ms = 0; h->t_passes = 0; h->t_amps = 0;
  memset(h->stats, 0, sizeof(h->stats));
  if (device < 0) {  // plan-only handle: the queue and the planner run, nothing is launched
    h->stream = nullptr;
    *out = h;
    return 0;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    std::string msg = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    delete h;
    return fail(nullptr, DQE_ECUDA, msg);
  }
  auto bail = [&](const char* what, cudaError_t err) {
    std::string msg = std::string(what) + ": " + cudaGetErrorString(err);
    dqe_destroy(h);
    return fail(nullptr, err == cudaErrorMemoryAllocation ? DQE_ENOMEM : DQE_ECUDA, msg);
  };
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) h->num_sms = sms;
  }
  if (stream) h->stream = (cudaStream_t)stream;
  else {
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return bail("cudaStreamCreate", e);
    h->own_stream = true;
  }
  const size_t elems = (size_t)batch << P;
  if (ext_state) h->state = (double2*)ext_state;
  else {
    // engine-owned buffers carry the peer flag block behind the state (same IPC mapping, see k_peer_exchange)
    e = cudaMalloc(&h->state, elems * sizeof(double2) + kPeerFlagWords * sizeof(uint64_t));
    if (e != cudaSuccess) return bail("cudaMalloc(state)", e);
    h->own_state = true;
    if ((e = cudaMemsetAsync(h->state + elems, 0, kPeerFlagWords * sizeof(uint64_t), h->stream)) != cudaSuccess)
      return bail("memset(flags)", e);
  }
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaMalloc(&h->creg_buf[i], batch * sizeof(uint64_t))) != cudaSuccess) return bail("cudaMalloc(creg)", e);
    if ((e = cudaMalloc(&h->rec_buf[i], batch * sizeof(MeasRec))) != cudaSuccess) return bail("cudaMalloc(rec)", e);
  }
  {
#define DQE_KV(M, B, F) (const void*)k_tile_pass<M, B, F, false>, (const void*)k_tile_pass<M, B, F, true>
    const void* variants[4] = {DQE_KV(256, 2, 0), DQE_KV(256, 2, 1)};
#undef DQE_KV
    for (const void* fn : variants) {
      if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess)
        return bail("cudaFuncSetAttribute", e);
      if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared)) != cudaSuccess)
        return bail("cudaFuncSetAttribute(carveout)", e);
      if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess)
        return bail("cudaFuncSetAttribute(cluster16)", e);
    }
  }
  if ((e = cudaMemsetAsync(h->state, 0, elems * sizeof(double2), h->stream)) != cudaSuccess) return bail("memset", e);
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaMemsetAsync(h->creg_buf[i], 0, batch * sizeof(uint64_t), h->stream)) != cudaSuccess) return bail("memset", e);
    if ((e = cudaMemsetAsync(h->rec_buf[i], 0, batch * sizeof(MeasRec), h->stream)) != cudaSuccess) return bail("memset", e);
  }
  k_init_state<<<(unsigned)((batch + 255) / 256), 256, 0, h->stream>>>(h->state, batch, P);
  if ((e = cudaGetLastError()) != cudaSuccess) return bail("k_init_state", e);
  *out = h;
  return 0;
}

int dqe_destroy(dqe_t* h) {
  if (h && h->cached) { delete h->cached; h->cached = nullptr; }
  if (!h) return 0;
  if (h->device < 0) { delete h; return 0; }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& p : h->ev_pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  if (h->xev_ok) { cudaEventDestroy(h->xev0); cudaEventDestroy(h->xev1); }
  if (h->own_state && h->state) cudaFree(h->state);
  for (int i = 0; i < 2; ++i) {
    if (h->creg_buf[i]) cudaFree(h->creg_buf[i]);
    if (h->rec_buf[i]) cudaFree(h->rec_buf[i]);
    if (h->treg_buf[i]) cudaFree(h->treg_buf[i]);
  }
  if (h->olog) cudaFree(h->olog);
  if (h->partials) cudaFree(h->partials);
  if (h->scratch) cudaFree(h->scratch);
  if (h->d_ops) cudaFree(h->d_ops);
  if (h->d_links) cudaFree(h->d_links);
  if (h->d_consts) cudaFree(h->d_consts);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int dqe_reinit(dqe_t* h) {
  if (int rc = check_ready(h)) return rc;
  h->queue.clear(); h->cpool.clear();
  // support of the state on the device: live mask at the head of the (dropped) queue, plus whatever is live now
  const uint64_t was_live = h->planned_live | phys_live(h, h->live_axes | h->pin_axes);
  h->live_axes = 0; h->planned_live = 0; h->draw = 0; h->pred = -1; h->api_mark = h->stats[4]; h->fresh = true;
  h->stream_next = (size_t)h->stream_ops;   // the same program streams at the same points on every batch
  // the next batch draws from the next block of global shot ids: reinit alone must not replay the stream
  h->shot_offset += (uint64_t)h->batch * (uint64_t)h->shot_stride;
  if (h->device < 0) return 0;
  CK(cudaSetDevice(h->device));
  if (h->state_dirty) {
    const size_t elems = (size_t)h->batch << h->P;
    CK(cudaMemsetAsync(h->state, 0, elems * sizeof(double2), h->stream));
    k_init_state<<<(unsigned)((h->batch + 255) / 256), 256, 0, h->stream>>>(h->state, h->batch, h->P);
    h->state_dirty = false;
  } else {
    // free axes are |0> by invariant: only the block the live axes span holds anything but zeros
    ResetParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.state = h->state; rp.batch = h->batch; rp.P = h->P;
    rp.live_bits = __builtin_popcountll(was_live);
    rp.n_segs = build_segs(was_live, rp.segs, 24);
    if (rp.n_segs < 0) return fail(h, DQE_EUNSUPPORTED, "reinit: too many bit segments");
    const uint64_t total = (uint64_t)h->batch << rp.live_bits;
    const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)h->num_sms * 32);
    k_reset_live<<<grid, 256, 0, h->stream>>>(rp);
  }
  CK(cudaMemsetAsync(cur_creg(h), 0, h->batch * sizeof(uint64_t), h->stream));
  if (h->olog) CK(cudaMemsetAsync(h->olog, 0xff, (size_t)h->olog_draws * (size_t)h->batch, h->stream));
  h->stats[2]++;
  CK(cudaGetLastError());
  return 0;
}

int dqe_seed(dqe_t* h, uint64_t seed, uint64_t shot_offset) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = do_flush(h)) return rc;
  h->seed = seed; h->shot_offset = shot_offset; h->draw = 0;
  return 0;
}

int dqe_alloc_axis(dqe_t* h, int axis) {
  if (int rc = check_ready(h)) return rc;
  if (axis < 0 || axis >= h->nmax) return fail(h, DQE_EINVAL, "alloc_axis: axis out of range");
  if ((h->live_axes >> axis) & 1ull) return fail(h, DQE_EINVAL, "alloc_axis: axis already live");
  h->live_axes |= 1ull << axis;
  // a free axis is |0> by invariant: nothing to do on the device.  The queue still has to learn
  // that the bits are live, which the next op's live_after carries.
  if (h->queue.empty()) h->planned_live = phys_live(h, h->live_axes);
  h->stats[4]++;
  return 0;
}

int dqe_free_axis(dqe_t* h, int axis) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = check_live(h, axis, "free_axis")) return rc;
  h->stats[4]++;
  if (h->formalism == DQE_KET) return lower_measure(h, axis, -1, -1, 1, 0.0);
  const int pred = h->pred;
  h->pred = -1;
  h->live_axes &= ~(1ull << axis);
  HostOp o = blank(h, OP_TRACE);
  o.m.b[0] = (uint8_t)rowbit(h, axis); o.m.b[1] = (uint8_t)axis; o.m.nb = 2;
  o.need = (1ull << rowbit(h, axis)) | (1ull << axis);
  push(h, o);
  h->pred = pred;
  return 0;
}

int dqe_reset(dqe_t* h, int axis) {
  if (int rc = check_ready(h)) return rc;
  if (axis < 0 || axis >= h->nmax) return fail(h, DQE_EINVAL, "reset: axis out of range");
  if ((h->live_axes >> axis) & 1ull)
    if (int rc = dqe_free_axis(h, axis)) return rc;
  return dqe_alloc_axis(h, axis);
}

int dqe_live_mask(dqe_t* h, uint64_t* mask) {
  if (!h || !mask) return DQE_EINVAL;
  *mask = h->live_axes;
  return 0;
}

int dqe_force_live_mask(dqe_t* h, uint64_t mask) {
  if (int rc = do_flush(h)) return rc;
  if (h->nmax < 64 && (mask >> h->nmax)) return fail(h, DQE_EINVAL, "force_live_mask: axis out of range");
  h->live_axes = mask;
  h->planned_live = phys_live(h, mask);
  return 0;
}

static void to_cplx(const double* in, cplx* out, int n) {
  for (int i = 0; i < n; ++i) out[i] = cplx(in[2 * i], in[2 * i + 1]);
}

// Big single-shot kets (a pass costs milliseconds): every `stream_ops` queued ops the passes that are already
// closed are launched, so that the device runs the head of the program while the host still walks its tail.
static int maybe_stream(dqe_t* h) {
  if (!h->stream_ops || h->formalism != DQE_KET || h->P < h->stream_min_bits || !h->fuse) return 0;
  if (h->queue.size() < h->stream_next) return 0;
  const int rc = do_flush(h, true);
  h->stream_flushes++;
  h->stream_next = h->queue.size() + (size_t)h->stream_ops;
  return rc;
}

int dqe_enqueue_1q(dqe_t* h, int axis, const double* m) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = maybe_stream(h)) return rc;
  if (int rc = check_live(h, axis, "1q gate")) return rc;
  cplx u[4];
  to_cplx(m, u, 4);
  h->stats[4]++;
  return lower_1q(h, axis, u);
}

int dqe_enqueue_2q(dqe_t* h, int a, int b, const double* m) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = maybe_stream(h)) return rc;
  if (int rc = check_live(h, a, "2q gate")) return rc;
  if (int rc = check_live(h, b, "2q gate")) return rc;
  if (a == b) return fail(h, DQE_EINVAL, "2q gate: axes must differ");
  cplx u[16];
  to_cplx(m, u, 16);
  h->stats[4]++;
  return lower_2q(h, a, b, u);
}

int dqe_enqueue_ctrl_1q(dqe_t* h, int c, int t, const double* m) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = maybe_stream(h)) return rc;
  if (int rc = check_live(h, c, "ctrl gate")) return rc;
  if (int rc = check_live(h, t, "ctrl gate")) return rc;
  if (c == t) return fail(h, DQE_EINVAL, "ctrl gate: axes must differ");
  cplx u[4];
  to_cplx(m, u, 4);
  h->stats[4]++;
  return lower_ctrl_1q(h, c, t, u);
}

int dqe_enqueue_diag(dqe_t* h, int k, const int* axes, const double* d) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = maybe_stream(h)) return rc;
  if (k < 1 || k > 4) return fail(h, DQE_EUNSUPPORTED, "diag: 1 <= k <= 4");
  for (int i = 0; i < k; ++i) {
    if (int rc = check_live(h, axes[i], "diag")) return rc;
    for (int j = 0; j < i; ++j) if (axes[i] == axes[j]) return fail(h, DQE_EINVAL, "diag: repeated axis");
  }
  cplx t[16];
  to_cplx(d, t, 1 << k);
  h->stats[4]++;
  return lower_diag(h, k, axes, t);
}

int dqe_enqueue_kq(dqe_t* h, int k, const int* axes, const double* m) {
  if (int rc = check_ready(h)) return rc;
  if (k < 1 || k > 4 || !axes || !m) return fail(h, DQE_EUNSUPPORTED, "kq gate: 1 <= k <= 4");
  for (int i = 0; i < k; ++i) {
    if (int rc = check_live(h, axes[i], "kq gate")) return rc;
    for (int j = 0; j < i; ++j) if (axes[i] == axes[j]) return fail(h, DQE_EINVAL, "kq gate: repeated axis");
  }
  if (k == 1) return dqe_enqueue_1q(h, axes[0], m);
  if (k == 2) return dqe_enqueue_2q(h, axes[0], axes[1], m);
  const int D = 1 << k;
  std::vector<cplx> u((size_t)D * D);
  to_cplx(m, u.data(), D * D);
  h->stats[4]++;
  for (int side = 0; side < (h->formalism == DQE_DM ? 2 : 1); ++side) {
    HostOp o = blank(h, OP_KBLOCK);
    o.m.nb = (uint32_t)k;
    for (int i = 0; i < k; ++i) {
      const int bit = (h->formalism == DQE_DM && side == 0) ? rowbit(h, axes[i]) : axes[i];
      o.m.b[i] = (uint8_t)bit; o.need |= 1ull << bit;
    }
    if (side == 1) for (auto& z : u) z = std::conj(z);   // column side of U rho U^dagger
    o.m.c_off = add_consts(h, u.data(), D * D); o.nconst = D * D;
    push(h, o);
  }
  return 0;
}

int dqe_set_predicate(dqe_t* h, int creg_bit) {
  if (int rc = check_ready(h)) return rc;
  if (creg_bit >= 64) return fail(h, DQE_EINVAL, "creg has 64 bits");
  h->pred = creg_bit < 0 ? -1 : creg_bit;
  return 0;
}

int dqe_enqueue_cond_1q(dqe_t* h, int axis, const double* m, int creg_bit) {
  if (int rc = check_ready(h)) return rc;
  if (creg_bit < 0 || creg_bit >= 64) return fail(h, DQE_EINVAL, "cond_1q: creg bit out of range");
  const int pred = h->pred;
  h->pred = creg_bit;
  const int rc = dqe_enqueue_1q(h, axis, m);
  h->pred = pred;
  return rc;
}

int dqe_enqueue_pauli_channel(dqe_t* h, int axis, const double* p) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = check_live(h, axis, "pauli_channel")) return rc;
  h->stats[4]++;
  if (h->formalism == DQE_DM) {
    q_pauli1(h, axis, p[0], p[1], p[2], p[3]);
  } else {
    const uint64_t draw = h->draw++;
    q_sampled(h, 1, &axis, 0, draw, p[0], p[0] + p[1], p[0] + p[1] + p[2]);
  }
  return 0;
}

// k-qubit Pauli channel sum_s p[s] P_s rho P_s (s = 4 * (Pauli on axes[0]) + (Pauli on axes[1]), order I X Y Z).
// dm: the channel's 16 x 16 superoperator on the bits (row a, row b, col a, col b) as ONE dense block (OP_KBLOCK, FP64
// tensor cores) -- any correlated two-qubit Pauli noise costs one op; ket: one Pauli pair drawn per shot.
int dqe_enqueue_pauli_channel_k(dqe_t* h, int k, const int* axes, const double* p) {
  if (int rc = check_ready(h)) return rc;
  if (!axes || !p) return fail(h, DQE_EINVAL, "pauli_channel: null");
  if (k == 1) return dqe_enqueue_pauli_channel(h, axes[0], p);
  if (k != 2) return fail(h, DQE_EUNSUPPORTED, "pauli_channel: k = 1 or 2");
  for (int i = 0; i < 2; ++i) if (int rc = check_live(h, axes[i], "pauli_channel")) return rc;
  if (axes[0] == axes[1]) return fail(h, DQE_EINVAL, "pauli_channel: axes must differ");
  double tot = 0.0;
  for (int i = 0; i < 16; ++i) { if (!(p[i] >= 0.0)) return fail(h, DQE_EINVAL, "pauli_channel: negative weight"); tot += p[i]; }
  if (std::abs(tot - 1.0) > 1e-9) return fail(h, DQE_EINVAL, "pauli_channel: weights must sum to 1");
  h->stats[4]++;
  if (h->formalism == DQE_DM) {
    static const cplx I1(0.0, 1.0);
    const cplx P[4][4] = {{1.0, 0.0, 0.0, 1.0}, {0.0, 1.0, 1.0, 0.0}, {0.0, -I1, I1, 0.0}, {1.0, 0.0, 0.0, -1.0}};
    std::vector<cplx> S(256, cplx(0.0, 0.0));
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        const double w = p[4 * a + b];
        if (w == 0.0) continue;
        for (int o = 0; o < 16; ++o)
          for (int i = 0; i < 16; ++i) {
            const int rA = (o >> 3) & 1, rB = (o >> 2) & 1, cA = (o >> 1) & 1, cB = o & 1;
            const int rA2 = (i >> 3) & 1, rB2 = (i >> 2) & 1, cA2 = (i >> 1) & 1, cB2 = i & 1;
            S[o * 16 + i] += w * P[a][rA * 2 + rA2] * P[b][rB * 2 + rB2] * std::conj(P[a][cA * 2 + cA2]) *
                             std::conj(P[b][cB * 2 + cB2]);
          }
      }
    HostOp o = blank(h, OP_KBLOCK);
    o.m.nb = 4;
    const int bits[4] = {rowbit(h, axes[0]), rowbit(h, axes[1]), axes[0], axes[1]};
    for (int i = 0; i < 4; ++i) { o.m.b[i] = (uint8_t)bits[i]; o.need |= 1ull << bits[i]; }
    o.m.c_off = add_consts(h, S.data(), 256); o.nconst = 256;
    push(h, o);
  } else {
    HostOp o = blank(h, OP_PAULI_SAMPLED);
    o.m.nb = 2;
    for (int i = 0; i < 2; ++i) { o.m.b[i] = (uint8_t)axes[i]; o.need |= 1ull << axes[i]; }
    o.m.flags = 6;
    o.m.aux = (int32_t)h->draw++;
    cplx cum[8];
    double run = 0.0, c[16];
    for (int i = 0; i < 16; ++i) { run += p[i]; c[i] = run; }
    for (int i = 0; i < 8; ++i) cum[i] = cplx(c[2 * i], c[2 * i + 1]);
    o.m.c_off = add_consts(h, cum, 8); o.nconst = 8;
    push(h, o);
  }
  return 0;
}

int dqe_enqueue_depolarize(dqe_t* h, int k, const int* axes, double p) {
  if (int rc = check_ready(h)) return rc;
  if (k != 1 && k != 2) return fail(h, DQE_EUNSUPPORTED, "depolarize: k = 1 or 2");
  for (int i = 0; i < k; ++i) if (int rc = check_live(h, axes[i], "depolarize")) return rc;
  if (k == 2 && axes[0] == axes[1]) return fail(h, DQE_EINVAL, "depolarize: axes must differ");
  h->stats[4]++;
  if (h->formalism == DQE_DM) {
    if (k == 1) q_pauli1(h, axes[0], 1.0 - 0.75 * p, 0.25 * p, 0.25 * p, 0.25 * p);
    else {
      HostOp o = blank(h, OP_DEPOL2);
      o.m.b[0] = (uint8_t)rowbit(h, axes[0]); o.m.b[1] = (uint8_t)rowbit(h, axes[1]);
      o.m.b[2] = (uint8_t)axes[0]; o.m.b[3] = (uint8_t)axes[1]; o.m.nb = 4;
      for (int i = 0; i < 4; ++i) o.need |= 1ull << o.m.b[i];
      o.m.p[0] = p;
      push(h, o);
    }
  } else {
    const uint64_t draw = h->draw++;
    const double n = k == 1 ? 4.0 : 16.0;
    const double w = p / n, w0 = 1.0 - (n - 1.0) * w;
    q_sampled(h, k, axes, 1, draw, w0, w, 0.0);
  }
  return 0;
}

int dqe_enqueue_kraus_1q(dqe_t* h, int axis, int nk, const double* K) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = check_live(h, axis, "kraus_1q")) return rc;
  if (h->formalism != DQE_DM) return fail(h, DQE_EUNSUPPORTED, "generic Kraus channels are dm-only");
  if (nk < 1 || nk > 16) return fail(h, DQE_EINVAL, "kraus_1q: 1 <= nk <= 16");
  cplx S[16];
  for (int i = 0; i < 16; ++i) S[i] = 0.0;
  for (int q = 0; q < nk; ++q) {
    cplx m[4];
    to_cplx(K + 8 * q, m, 4);
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 2; ++c)
        for (int r2 = 0; r2 < 2; ++r2)
          for (int c2 = 0; c2 < 2; ++c2)
            S[(r * 2 + c) * 4 + (r2 * 2 + c2)] += m[r * 2 + r2] * std::conj(m[c * 2 + c2]);
  }
  h->stats[4]++;
  q_mat2(h, rowbit(h, axis), axis, S, 0);
  return 0;
}

static int inject_common(dqe_t* h, int a, int b) {
  if (a < 0 || a >= h->nmax || b < 0 || b >= h->nmax || a == b) return fail(h, DQE_EINVAL, "inject_2q: bad axes");
  if (((h->live_axes >> a) & 1ull) || ((h->live_axes >> b) & 1ull))
    return fail(h, DQE_EINVAL, "inject_2q: axis is occupied");
  h->live_axes |= (1ull << a) | (1ull << b);
  h->stats[4]++;
  return 0;
}

int dqe_inject_2q(dqe_t* h, int a, int b, const double* state, int is_dm) {
  if (int rc = check_ready(h)) return rc;
  if (h->formalism == DQE_KET && is_dm)
    return fail(h, DQE_EUNSUPPORTED, "ket formalism: pass a decomposition to dqe_inject_2q_mixture");
  if (int rc = inject_common(h, a, b)) return rc;
  const int pred = h->pred;
  h->pred = -1;
  HostOp o = blank(h, OP_INJECT);
  if (h->formalism == DQE_KET) {
    cplx v[4];
    to_cplx(state, v, 4);
    o.m.b[0] = (uint8_t)a; o.m.b[1] = (uint8_t)b; o.m.nb = 2;
    o.need = (1ull << a) | (1ull << b);
    o.m.c_off = add_consts(h, v, 4); o.nconst = 4;
  } else {
    cplx sig[16];
    if (is_dm) to_cplx(state, sig, 16);
    else {
      cplx v[4];
      to_cplx(state, v, 4);
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) sig[r * 4 + c] = v[r] * std::conj(v[c]);
    }
    o.m.b[0] = (uint8_t)rowbit(h, a); o.m.b[1] = (uint8_t)rowbit(h, b);
    o.m.b[2] = (uint8_t)a; o.m.b[3] = (uint8_t)b; o.m.nb = 4;
    for (int i = 0; i < 4; ++i) o.need |= 1ull << o.m.b[i];
    o.m.c_off = add_consts(h, sig, 16); o.nconst = 16;
  }
  push(h, o);
  h->pred = pred;
  return 0;
}

int dqe_inject_2q_mixture(dqe_t* h, int a, int b, int ncomp, const double* probs, const double* kets) {
  if (int rc = check_ready(h)) return rc;
  if (h->formalism != DQE_KET) return fail(h, DQE_EUNSUPPORTED, "inject_2q_mixture is ket-only (dm takes the matrix)");
  if (ncomp < 1 || ncomp > 4) return fail(h, DQE_EINVAL, "inject_2q_mixture: 1 <= ncomp <= 4");
  if (int rc = inject_common(h, a, b)) return rc;
  const int pred = h->pred;
  h->pred = -1;
  HostOp o = blank(h, OP_INJECT);
  cplx v[16];
  for (int i = 0; i < 16; ++i) v[i] = 0.0;
  to_cplx(kets, v, 4 * ncomp);
  double cum[4] = {2.0, 2.0, 2.0, 2.0}, acc = 0.0;   // thresholds beyond the last component never fire
  for (int i = 0; i < ncomp; ++i) { acc += probs[i]; cum[i] = acc; }
  if (ncomp < 4) for (int i = ncomp - 1; i < 4; ++i) cum[i] = 2.0;
  o.m.b[0] = (uint8_t)a; o.m.b[1] = (uint8_t)b; o.m.nb = 2;
  o.need = (1ull << a) | (1ull << b);
  o.m.flags = 2;
  o.m.aux = (int32_t)h->draw++;
  o.m.p[0] = cum[0]; o.m.p[1] = cum[1]; o.m.p[2] = cum[2];
  o.m.c_off = add_consts(h, v, 16); o.nconst = 16;
  push(h, o);
  h->pred = pred;
  return 0;
}

int dqe_measure(dqe_t* h, int axis, int creg_bit, int forced, int discard, double flip_prob) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = check_live(h, axis, "measure")) return rc;
  if (creg_bit >= 64) return fail(h, DQE_EINVAL, "creg has 64 bits");
  if (forced > 1) return fail(h, DQE_EINVAL, "forced outcome must be -1, 0 or 1");
  h->stats[4]++;
  const int pred = h->pred;
  h->pred = -1;
  const int rc = lower_measure(h, axis, creg_bit, forced, discard, flip_prob);
  h->pred = pred;
  return rc;
}

int dqe_n_time_regs(void) { return kTimeRegs; }

static int ensure_tregs(dqe_t* h) {
  if (h->device < 0 || h->treg_buf[0]) return 0;
  CK(cudaSetDevice(h->device));
  const size_t bytes = (size_t)h->batch * kTimeRegs * sizeof(double);
  for (int i = 0; i < 2; ++i) {
    CK(cudaMalloc(&h->treg_buf[i], bytes));
    CK(cudaMemsetAsync(h->treg_buf[i], 0, bytes, h->stream));
  }
  return 0;
}

int dqe_classical_op(dqe_t* h, int opcode, int dst, int a, int b, int cbit, double imm) {
  if (int rc = check_ready(h)) return rc;
  const bool bitop = opcode >= CL_BAND && opcode <= CL_BCOPY;
  const bool regop = opcode >= CL_CONST && opcode <= CL_ADDIF;
  if (!bitop && !regop) return fail(h, DQE_EINVAL, "classical_op: unknown opcode");
  const int lim = bitop ? 64 : kTimeRegs;
  if (dst < 0 || dst >= lim || a < 0 || a >= lim || b < 0 || b >= lim)
    return fail(h, DQE_EINVAL, bitop ? "classical_op: creg bit out of range" : "classical_op: register out of range");
  const bool uses_bit = opcode == CL_SEL || opcode == CL_ADDIF;
  if (uses_bit && (cbit < 0 || cbit >= 64)) return fail(h, DQE_EINVAL, "classical_op: SEL / ADDIF need a creg bit");
  if (opcode == CL_CEILQ && !(imm > 0.0)) return fail(h, DQE_EINVAL, "classical_op: CEILQ needs a positive divisor");
  if (int rc = ensure_tregs(h)) return rc;
  HostOp o = blank(h, 0);
  o.kind = 2;
  o.m.pred_bit = -1;
  o.ci.op = (uint8_t)opcode; o.ci.dst = (uint8_t)dst; o.ci.a = (uint8_t)a; o.ci.b = (uint8_t)b;
  o.ci.cbit = (int8_t)(uses_bit ? cbit : 0); o.ci.pad = 0; o.ci.imm = 0;
  o.imm = imm;
  if (regop) h->max_reg = std::max(h->max_reg, dst);
  h->queue.push_back(o);
  h->stats[4]++;
  return 0;
}

int dqe_classical_qexp2(dqe_t* h, int dst, int a, int b, double offset, double rate) {
  if (int rc = check_ready(h)) return rc;
  auto bad = [](int r) { return r != 255 && (r < 0 || r >= kTimeRegs); };
  if (dst < 0 || dst >= kTimeRegs || bad(a) || bad(b)) return fail(h, DQE_EINVAL, "classical_qexp2: register out of range");
  if (int rc = ensure_tregs(h)) return rc;
  HostOp o = blank(h, 0);
  o.kind = 2;
  o.m.pred_bit = -1;
  o.ci.op = (uint8_t)CL_QEXP2; o.ci.dst = (uint8_t)dst; o.ci.a = (uint8_t)a; o.ci.b = (uint8_t)b;
  o.ci.cbit = 0; o.ci.pad = 0; o.ci.imm = 0;
  o.imm = offset; o.imm2 = rate;
  h->max_reg = std::max(h->max_reg, dst);
  h->queue.push_back(o);
  h->stats[4]++;
  return 0;
}

// Per-shot random numbers for the classical machine (the heralded entanglement link, physical_layer.py:685-753):
// kind 0: creg bit dst <- (u < p);  kind 1: register dst <- min(kmax, 1 + floor(ln u / ln(1 - p))), the number of
// attempts until the first success of probability p.  u = the shot's Philox uniform of the next draw index.
int dqe_classical_rand(dqe_t* h, int kind, int dst, double p, double kmax) {
  if (int rc = check_ready(h)) return rc;
  if (kind != 0 && kind != 1) return fail(h, DQE_EINVAL, "classical_rand: kind 0 (bit) or 1 (geometric)");
  if (!(p >= 0.0 && p <= 1.0)) return fail(h, DQE_EINVAL, "classical_rand: probability in [0, 1]");
  if (kind == 0 ? (dst < 0 || dst >= 64) : (dst < 0 || dst >= kTimeRegs)) return fail(h, DQE_EINVAL, "classical_rand: dst out of range");
  if (h->draw >= 65536) return fail(h, DQE_EUNSUPPORTED, "classical_rand: draw index beyond 16 bits");
  if (h->olog_draws > 0 && (int64_t)h->draw >= h->olog_draws) return fail(h, DQE_EINVAL, "outcome log is full: raise the outcome_log option");
  if (int rc = ensure_tregs(h)) return rc;
  const uint64_t draw = h->draw++;
  HostOp o = blank(h, 0);
  o.kind = 2;
  o.m.pred_bit = -1;
  o.ci.op = (uint8_t)(kind == 0 ? CL_RANDBIT : CL_GEOM); o.ci.dst = (uint8_t)dst;
  o.ci.a = (uint8_t)(draw & 0xff); o.ci.b = (uint8_t)(draw >> 8);
  o.ci.cbit = 0; o.ci.pad = 0; o.ci.imm = 0;
  if (kind == 0) o.imm = p;
  else { o.imm = p >= 1.0 ? -1e300 : (p <= 0.0 ? -1e-300 : std::log1p(-p)); o.imm2 = kmax; }
  if (kind == 1) h->max_reg = std::max(h->max_reg, dst);
  h->queue.push_back(o);
  h->stats[4]++;
  return 0;
}

int dqe_classical_fence(dqe_t* h) {
  if (int rc = check_ready(h)) return rc;
  HostOp o = blank(h, 0);
  o.kind = 3;
  o.m.pred_bit = -1;
  h->queue.push_back(o);
  return 0;
}

int dqe_enqueue_dyn_channel(dqe_t* h, int axis, int kind, int reg) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = check_live(h, axis, "dyn_channel")) return rc;
  if (reg < 0 || reg >= kTimeRegs) return fail(h, DQE_EINVAL, "dyn_channel: register out of range");
  if (kind < DYN_DEPOL || kind > DYN_ZFLIP) return fail(h, DQE_EINVAL, "dyn_channel: kind in {0 depol, 1 dephase, 2 ampdamp, 3 zflip}");
  if (int rc = ensure_tregs(h)) return rc;
  h->stats[4]++;
  if (h->formalism == DQE_DM) {
    HostOp o = blank(h, OP_DYN1);
    o.m.b[0] = (uint8_t)rowbit(h, axis); o.m.b[1] = (uint8_t)axis; o.m.nb = 2;
    o.need = (1ull << rowbit(h, axis)) | (1ull << axis);
    o.m.flags = kind; o.m.aux = reg;
    push(h, o);
    return 0;
  }
  if (kind == DYN_AMPDAMP) return fail(h, DQE_EUNSUPPORTED, "amplitude damping is dm-only");
  HostOp o = blank(h, OP_PAULI_SAMPLED);
  o.m.nb = 1; o.m.b[0] = (uint8_t)axis; o.need = 1ull << axis;
  o.m.flags = kind == DYN_DEPOL ? 3 : kind == DYN_DEPHASE ? 4 : 5;
  o.m.aux = (int32_t)h->draw++;
  o.m.p[3] = (double)reg;   // (lctrl is rebuilt from the control masks when the op is localised)
  push(h, o);
  return 0;
}

int dqe_read_time_reg(dqe_t* h, int reg, double* out) {
  if (int rc = do_flush(h)) return rc;
  if (h && h->device < 0) return fail(h, DQE_EUNSUPPORTED, "plan-only handle (device < 0) has no state");
  if (reg < 0 || reg >= kTimeRegs || !out) return fail(h, DQE_EINVAL, "read_time_reg: bad argument");
  if (!h->treg_buf[0]) return fail(h, DQE_EINVAL, "read_time_reg: no classical instruction has run");
  CK(cudaMemcpy2DAsync(out, sizeof(double), h->treg_buf[h->tpar] + reg, kTimeRegs * sizeof(double), sizeof(double),
                       (size_t)h->batch, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dqe_read_outcomes(dqe_t* h, int64_t first, int64_t count, int8_t* out) {
  if (int rc = do_flush(h)) return rc;
  if (h && h->device < 0) return fail(h, DQE_EUNSUPPORTED, "plan-only handle (device < 0) has no state");
  if (!h->olog) return fail(h, DQE_EINVAL, "read_outcomes: enable the outcome_log option first");
  if (first < 0 || count < 0 || first + count > h->olog_draws || !out) return fail(h, DQE_EINVAL, "read_outcomes: range");
  CK(cudaMemcpyAsync(out, h->olog + (size_t)first * (size_t)h->batch, (size_t)count * (size_t)h->batch,
                     cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int64_t dqe_draw_count(dqe_t* h) { return h ? (int64_t)h->draw : -1; }

int dqe_flush(dqe_t* h) { return do_flush(h); }

int dqe_replay(dqe_t* h) {
  if (int rc = check_ready(h)) return rc;
  if (h->device < 0) return fail(h, DQE_EUNSUPPORTED, "plan-only handle (device < 0) has no state");
  if (!h->cached) return fail(h, DQE_EINVAL, "replay: no cached plan (set the plan_cache option and flush a whole program first)");
  if (!h->queue.empty() || !h->fresh || h->live_axes != 0 || h->draw != 0)
    return fail(h, DQE_EINVAL, "replay: the register must be fresh (dqe_reinit) and the queue empty");
  CachedPlan& c = *h->cached;
  CK(cudaSetDevice(h->device));
  if (c.dev_ops.size() > h->d_ops_cap || c.dev_consts.size() > h->d_consts_cap)
    return fail(h, DQE_EINVAL, "replay: device program buffers were released");
  for (PlannedPass& P : c.passes) {   // the one thing that changes from batch to batch: the shots' Philox counters
    if (P.is_finish) { P.fp.seed = h->seed; P.fp.shot_offset = h->shot_offset; }
    else { P.pp.seed = h->seed; P.pp.shot_offset = h->shot_offset; P.pp.creg_out = nullptr; P.pp.rec_out = nullptr; P.pp.treg_out = nullptr; }
  }
  h->stats[6] += (int64_t)(c.dev_ops.size() * sizeof(MicroOp) + c.dev_consts.size() * sizeof(double2));
  if (!c.dev_ops.empty())
    CK(cudaMemcpyAsync(h->d_ops, c.dev_ops.data(), c.dev_ops.size() * sizeof(MicroOp), cudaMemcpyHostToDevice, h->stream));
  if (!c.dev_consts.empty())
    CK(cudaMemcpyAsync(h->d_consts, c.dev_consts.data(), c.dev_consts.size() * sizeof(double2), cudaMemcpyHostToDevice,
                       h->stream));
  Planner pl;
  pl.h = h;
  pl.passes = c.passes;
  h->fresh = false;
  if (int rc = launch_passes(h, pl)) return rc;
  h->live_axes = c.live_axes; h->planned_live = c.planned_live; h->draw = c.draw;
  h->stats[4] += c.api_ops;
  return 0;
}

#define NEED_DEVICE(h) do { if ((h) && (h)->device < 0) return fail((h), DQE_EUNSUPPORTED, "plan-only handle (device < 0) has no state"); } while (0)

int dqe_sync(dqe_t* h) {
  if (int rc = do_flush(h)) return rc;
  if (h->device < 0) return 0;
  CK(cudaStreamSynchronize(h->stream));
  if (h->timing) return collect_timing(h);
  return 0;
}

int dqe_read_creg(dqe_t* h, int bit, int8_t* out) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (bit < 0 || bit >= 64 || !out) return fail(h, DQE_EINVAL, "read_creg: bad argument");
  if (int rc = ensure_scratch(h, (size_t)h->batch)) return rc;
  int8_t* d = (int8_t*)h->scratch;
  k_extract_bit<<<(unsigned)((h->batch + 255) / 256), 256, 0, h->stream>>>(cur_creg(h), bit, d, h->batch);
  h->stats[2]++;
  CK(cudaMemcpyAsync(out, d, h->batch, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dqe_read_creg_words(dqe_t* h, uint64_t* out) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (!out) return fail(h, DQE_EINVAL, "read_creg_words: null");
  CK(cudaMemcpyAsync(out, cur_creg(h), h->batch * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dqe_histogram(dqe_t* h, int k, const int* cbits, int64_t* out) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (k < 1 || k > 20 || !cbits || !out) return fail(h, DQE_EINVAL, "histogram: 1 <= k <= 20");
  HistParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.creg = cur_creg(h); hp.batch = h->batch; hp.k = k;
  for (int i = 0; i < k; ++i) {
    if (cbits[i] < 0 || cbits[i] >= 64) return fail(h, DQE_EINVAL, "histogram: creg bit out of range");
    hp.bits[i] = cbits[i];
  }
  const size_t nb = (size_t)1 << k;
  if (int rc = ensure_scratch(h, nb * sizeof(unsigned long long))) return rc;
  hp.bins = (unsigned long long*)h->scratch;
  CK(cudaMemsetAsync(hp.bins, 0, nb * sizeof(unsigned long long), h->stream));
  k_histogram<<<(unsigned)((h->batch + 255) / 256), 256, 0, h->stream>>>(hp);
  h->stats[2]++;
  CK(cudaMemcpyAsync(out, hp.bins, nb * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dqe_reduced_dm(dqe_t* h, int k, const int* axes, int64_t shot, double* out) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (shot < 0 || shot >= h->batch || !out) return fail(h, DQE_EINVAL, "reduced_dm: bad shot");
  ReadoutParams rp;
  if (int rc = build_readout(h, k, axes, &rp)) return rc;
  const size_t n = (size_t)1 << (2 * k);
  if (int rc = ensure_scratch(h, n * sizeof(double2))) return rc;
  rp.out_rdm = (double2*)h->scratch;
  k_reduced_dm<<<(unsigned)n, 256, 0, h->stream>>>(rp, shot);
  h->stats[2]++;
  CK(cudaMemcpyAsync(out, rp.out_rdm, n * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dqe_fidelity_ket(dqe_t* h, int k, const int* axes, const double* ket, double* out) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (!ket || !out) return fail(h, DQE_EINVAL, "fidelity_ket: null");
  ReadoutParams rp;
  if (int rc = build_readout(h, k, axes, &rp)) return rc;
  const size_t n = (size_t)1 << k;
  if (int rc = ensure_scratch(h, n * sizeof(double2) + (size_t)h->batch * sizeof(double))) return rc;
  double2* dk = (double2*)h->scratch;
  rp.out_fid = (double*)(dk + n);
  CK(cudaMemcpyAsync(dk, ket, n * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  rp.ket = dk;
  k_fidelity_ket<<<(unsigned)h->batch, 128, 0, h->stream>>>(rp);
  h->stats[2]++;
  CK(cudaMemcpyAsync(out, rp.out_fid, h->batch * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dqe_get_state(dqe_t* h, int64_t shot, double* out) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (shot < 0 || shot >= h->batch || !out) return fail(h, DQE_EINVAL, "get_state: bad shot");
  const size_t n = (size_t)1 << h->P;
  CK(cudaMemcpyAsync(out, h->state + (size_t)shot * n, n * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dqe_set_state(dqe_t* h, int64_t shot, const double* in, uint64_t live_mask) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (shot < 0 || shot >= h->batch || !in) return fail(h, DQE_EINVAL, "set_state: bad shot");
  if (h->nmax < 64 && (live_mask >> h->nmax)) return fail(h, DQE_EINVAL, "set_state: live mask out of range");
  const size_t n = (size_t)1 << h->P;
  CK(cudaMemcpyAsync(h->state + (size_t)shot * n, in, n * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->live_axes = live_mask;
  h->planned_live = phys_live(h, live_mask);
  h->state_dirty = true;
  return 0;
}

int dqe_state_ptr(dqe_t* h, void** dev, int64_t* nbytes) {
  if (!h || !dev || !nbytes) return DQE_EINVAL;
  *dev = h->state;
  *nbytes = (int64_t)(((size_t)h->batch << h->P) * sizeof(double2));
  return 0;
}

int dqe_ipc_export(dqe_t* h, void* handle64) {
  if (int rc = check_ready(h)) return rc;
  NEED_DEVICE(h);
  if (!handle64) return fail(h, DQE_EINVAL, "ipc_export: null");
  if (!h->own_state) return fail(h, DQE_EUNSUPPORTED, "ipc_export: only engine-owned (cudaMalloc) state buffers");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CK(cudaSetDevice(h->device));
  cudaIpcMemHandle_t hd;
  CK(cudaIpcGetMemHandle(&hd, h->state));
  memcpy(handle64, &hd, 64);
  return 0;
}

int dqe_ipc_open(dqe_t* h, const void* handle64, void** peer_state) {
  if (int rc = check_ready(h)) return rc;
  NEED_DEVICE(h);
  if (!handle64 || !peer_state) return fail(h, DQE_EINVAL, "ipc_open: null");
  CK(cudaSetDevice(h->device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  CK(cudaIpcOpenMemHandle(peer_state, hd, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int dqe_ipc_close(dqe_t* h, void* peer_state) {
  if (int rc = check_ready(h)) return rc;
  NEED_DEVICE(h);
  CK(cudaSetDevice(h->device));
  CK(cudaIpcCloseMemHandle(peer_state));
  return 0;
}

int dqe_peer_1q(dqe_t* h, void* peer_state, int my_bit, int half, const double* m, int ctrl_axis) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (h->formalism != DQE_KET || h->batch != 1) return fail(h, DQE_EUNSUPPORTED, "peer_1q: single-shot ket only");
  if (!peer_state || !m || (my_bit | 1) != 1 || (half | 1) != 1) return fail(h, DQE_EINVAL, "peer_1q: bad argument");
  if (h->live_axes != (h->nmax >= 64 ? ~0ull : ((1ull << h->nmax) - 1ull)))
    return fail(h, DQE_EINVAL, "peer_1q: every local axis must be live");
  if (ctrl_axis >= h->nmax) return fail(h, DQE_EINVAL, "peer_1q: control axis out of range");
  const uint64_t n = 1ull << h->P, mid = n >> 1;
  double2* lo = my_bit == 0 ? h->state : (double2*)peer_state;
  double2* hi = my_bit == 0 ? (double2*)peer_state : h->state;
  const uint64_t begin = half ? mid : 0, end = half ? n : mid;
  const uint64_t cm = ctrl_axis >= 0 ? (1ull << ctrl_axis) : 0ull;
  CK(cudaSetDevice(h->device));
  k_peer_1q<<<h->num_sms * 8, 256, 0, h->stream>>>(lo, hi, begin, end, cm, make_double2(m[0], m[1]),
                                                    make_double2(m[2], m[3]), make_double2(m[4], m[5]),
                                                    make_double2(m[6], m[7]));
  h->stats[2]++;
  CK(cudaGetLastError());
  return 0;
}

int dqe_peer_swap(dqe_t* h, void* peer_state, int my_bit, int half, int local_axis) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (h->formalism != DQE_KET || h->batch != 1) return fail(h, DQE_EUNSUPPORTED, "peer_swap: single-shot ket only");
  if (!peer_state || (my_bit | 1) != 1 || (half | 1) != 1 || local_axis < 0 || local_axis >= h->nmax)
    return fail(h, DQE_EINVAL, "peer_swap: bad argument");
  if (h->live_axes != (h->nmax >= 64 ? ~0ull : ((1ull << h->nmax) - 1ull)))
    return fail(h, DQE_EINVAL, "peer_swap: every local axis must be live");
  const uint64_t npairs = 1ull << (h->P - 1), mid = npairs >> 1;
  double2* lo = my_bit == 0 ? h->state : (double2*)peer_state;
  double2* hi = my_bit == 0 ? (double2*)peer_state : h->state;
  CK(cudaSetDevice(h->device));
  k_peer_swap<<<h->num_sms * 8, 256, 0, h->stream>>>(lo, hi, half ? mid : 0, half ? npairs : mid, local_axis);
  h->stats[2]++;
  CK(cudaGetLastError());
  return 0;
}

static int collect_exchange_time(dqe_t* h) {
  if (!h->xev_pending) return 0;
  float ms = 0.f;
  CK(cudaEventSynchronize(h->xev1));
  CK(cudaEventElapsedTime(&ms, h->xev0, h->xev1));
  h->x_ms += ms;
  h->xev_pending = false;
  return 0;
}

int dqe_peer_exchange(dqe_t* h, void* const* peer_states, int rank, int world, int s, const int* rank_bits,
                      const int* local_axes) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (h->formalism != DQE_KET || h->batch != 1) return fail(h, DQE_EUNSUPPORTED, "peer_exchange: single-shot ket only");
  if (!h->own_state) return fail(h, DQE_EUNSUPPORTED, "peer_exchange: engine-owned state buffers only");
  if (!peer_states || !rank_bits || !local_axes || s < 1 || s > 6 || world < 2 || world > 64 || rank < 0 || rank >= world)
    return fail(h, DQE_EINVAL, "peer_exchange: bad argument");
  if (h->P - s < 1) return fail(h, DQE_EINVAL, "peer_exchange: too many axes for this shard");
  ExchangeParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mine = h->state; ep.flag_off = 1ull << h->P; ep.rank = rank; ep.s = s; ep.m = h->P; ep.nparts = (1 << s) - 1;
  uint64_t lseen = 0;
  for (int k = 0; k < s; ++k) {
    if (rank_bits[k] < 0 || (1 << rank_bits[k]) >= world || local_axes[k] < 0 || local_axes[k] >= h->nmax ||
        ((lseen >> local_axes[k]) & 1ull) || (k > 0 && rank_bits[k] <= rank_bits[k - 1]))
      return fail(h, DQE_EINVAL, "peer_exchange: rank bits ascending and inside the world, local axes distinct");
    lseen |= 1ull << local_axes[k];
    ep.jbits[k] = rank_bits[k]; ep.lbits[k] = local_axes[k]; ep.lsorted[k] = local_axes[k];
  }
  std::sort(ep.lsorted, ep.lsorted + s);
  for (int p = 0; p < world; ++p) ep.peer[p] = p == rank ? h->state : (double2*)peer_states[p];
  for (int d = 1; d < (1 << s); ++d) {
    int p = rank;
    for (int k = 0; k < s; ++k) if ((d >> k) & 1) p ^= 1 << rank_bits[k];
    if (!ep.peer[p]) return fail(h, DQE_EINVAL, "peer_exchange: partner shard not mapped");
    ep.epoch[p] = ++h->peer_epoch[p];
  }
  CK(cudaSetDevice(h->device));
  if (int rc = collect_exchange_time(h)) return rc;
  if (!h->xev_ok) { CK(cudaEventCreate(&h->xev0)); CK(cudaEventCreate(&h->xev1)); h->xev_ok = true; }
  CK(cudaEventRecord(h->xev0, h->stream));
  // every CTA must be resident (the last one to finish waits for the partners): <= 8 CTAs of 256 threads per SM
  k_peer_exchange<<<h->num_sms * 4, 256, 0, h->stream>>>(ep);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->xev1, h->stream));
  h->xev_pending = true;
  h->x_count++;
  h->x_bytes += (int64_t)((1ull << (h->P - s)) * (uint64_t)((1 << s) - 1) * sizeof(double2));
  h->stats[2]++;
  return 0;
}

int dqe_peer_stats(dqe_t* h, double* ms, int64_t* count, int64_t* bytes_sent, int64_t* error) {
  if (int rc = check_ready(h)) return rc;
  NEED_DEVICE(h);
  CK(cudaSetDevice(h->device));
  if (int rc = collect_exchange_time(h)) return rc;
  if (ms) *ms = h->x_ms;
  if (count) *count = h->x_count;
  if (bytes_sent) *bytes_sent = h->x_bytes;
  if (error) {
    uint64_t e = 0;
    if (h->own_state) {
      CK(cudaStreamSynchronize(h->stream));
      CK(cudaMemcpy(&e, reinterpret_cast<uint64_t*>(h->state + ((size_t)h->batch << h->P)) + kPeerError, sizeof(e), cudaMemcpyDeviceToHost));
    }
    *error = (int64_t)e;
  }
  return 0;
}

int dqe_probs(dqe_t* h, int axis, double* out2) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (h->formalism != DQE_KET || h->batch != 1) return fail(h, DQE_EUNSUPPORTED, "probs: single-shot ket only");
  if (!out2) return fail(h, DQE_EINVAL, "probs: null");
  if (int rc = check_live(h, axis, "probs")) return rc;
  const int nblk = h->num_sms * 8;
  if (int rc = ensure_scratch(h, (size_t)nblk * 2 * sizeof(double))) return rc;
  CK(cudaSetDevice(h->device));
  k_probs_bit<<<nblk, 256, 0, h->stream>>>(h->state, 1ull << h->P, axis, (double*)h->scratch);
  CK(cudaGetLastError());
  std::vector<double> part((size_t)nblk * 2);
  CK(cudaMemcpyAsync(part.data(), h->scratch, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  out2[0] = out2[1] = 0.0;
  for (int b = 0; b < nblk; ++b) { out2[0] += part[2 * b]; out2[1] += part[2 * b + 1]; }
  return 0;
}

int dqe_probs_diag(dqe_t* h, int npairs, const int* a, const int* b, const int* fixed, int target, double* out2) {
  if (int rc = do_flush(h)) return rc;
  NEED_DEVICE(h);
  if (h->formalism != DQE_KET || h->batch != 1) return fail(h, DQE_EUNSUPPORTED, "probs_diag: single-shot ket view only");
  if (!out2 || npairs < 0 || npairs > 32 || (npairs && (!a || !b || !fixed)) || target >= h->nmax)
    return fail(h, DQE_EINVAL, "probs_diag: bad argument");
  DiagProbParams dp;
  memset(&dp, 0, sizeof(dp));
  dp.state = h->state; dp.n = 1ull << h->P; dp.npairs = npairs; dp.target = target;
  for (int k = 0; k < npairs; ++k) {
    if (a[k] < 0 || a[k] >= h->nmax || (fixed[k] < 0 && (b[k] < 0 || b[k] >= h->nmax)) || fixed[k] > 1)
      return fail(h, DQE_EINVAL, "probs_diag: pair out of range");
    dp.a[k] = (int8_t)a[k]; dp.b[k] = (int8_t)(fixed[k] < 0 ? b[k] : 0); dp.fixed[k] = (int8_t)fixed[k];
  }
  const int nblk = h->num_sms * 8;
  if (int rc = ensure_scratch(h, (size_t)nblk * 2 * sizeof(double))) return rc;
  dp.out = (double*)h->scratch;
  CK(cudaSetDevice(h->device));
  k_probs_diag<<<nblk, 256, 0, h->stream>>>(dp);
  CK(cudaGetLastError());
  std::vector<double> part((size_t)nblk * 2);
  CK(cudaMemcpyAsync(part.data(), h->scratch, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  out2[0] = out2[1] = 0.0;
  for (int blk = 0; blk < nblk; ++blk) { out2[0] += part[2 * blk]; out2[1] += part[2 * blk + 1]; }
  return 0;
}

int dqe_set_option(dqe_t* h, const char* key, int64_t value) {
  if (!h || !key) return DQE_EINVAL;
  if (int rc = do_flush(h)) return rc;
  if (h->cached) { delete h->cached; h->cached = nullptr; }   // planned under the old options
  if (!strcmp(key, "tile_bits")) {
    if (value < 4 || value > 13) return fail(h, DQE_EINVAL, "tile_bits in [4, 13]");
    h->tile_bits = (int)value; h->tile_auto = 0;
  } else if (!strcmp(key, "min_run_bits")) {
    if (value < 0 || value > 12) return fail(h, DQE_EINVAL, "min_run_bits in [0, 12]");
    h->min_run_bits = (int)value;
  } else if (!strcmp(key, "fuse")) {
    h->fuse = value ? 1 : 0;
  } else if (!strcmp(key, "merge")) {
    h->merge = value ? 1 : 0;
  } else if (!strcmp(key, "group")) {
    h->group = value ? 1 : 0;
  } else if (!strcmp(key, "tma")) {
    h->tma = value ? 1 : 0;
  } else if (!strcmp(key, "pin_mask")) {
    if (h->nmax < 64 && ((uint64_t)value >> h->nmax)) return fail(h, DQE_EINVAL, "pin_mask out of range");
    h->pin_axes = (uint64_t)value;
  } else if (!strcmp(key, "kblock")) {
    h->kblock = value ? 1 : 0;
  } else if (!strcmp(key, "plan_cache")) {
    h->plan_cache = value ? 1 : 0;
  } else if (!strcmp(key, "pass_ops")) {
    if (value < 4 || value > kMaxOpsPerPass) return fail(h, DQE_EINVAL, "pass_ops in [4, 48]");
    h->pass_ops = (int)value;
  } else if (!strcmp(key, "catfuse")) {
    if (value < 0 || value > 3) return fail(h, DQE_EINVAL, "catfuse in {0 off, 1 inject runs, 2 + measure-out runs, 3 + whole remote gates on the data pair}");
    h->catfuse = (int)value;
  } else if (!strcmp(key, "ladder")) {
    h->ladder = value ? 1 : 0;
  } else if (!strcmp(key, "phase_tables")) {
    h->phase_tables = value ? 1 : 0;
  } else if (!strcmp(key, "kgroup")) {
    h->kgroup = value ? 1 : 0;
  } else if (!strcmp(key, "stream_ops")) {
    if (value < 0) return fail(h, DQE_EINVAL, "stream_ops >= 0");
    h->stream_ops = (int)value; h->stream_next = (size_t)value;
  } else if (!strcmp(key, "stream_min_bits")) {
    h->stream_min_bits = (int)value;
  } else if (!strcmp(key, "shot_stride")) {
    if (value < 1) return fail(h, DQE_EINVAL, "shot_stride >= 1");
    h->shot_stride = value;
  } else if (!strcmp(key, "outcome_log")) {
    if (value < 0 || value > (1 << 20)) return fail(h, DQE_EINVAL, "outcome_log: number of draws in [0, 2^20]");
    if (h->device < 0) { h->olog_draws = value; return 0; }
    CK(cudaSetDevice(h->device));
    if (h->olog) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->olog)); h->olog = nullptr; }
    h->olog_draws = value;
    if (value > 0) {
      CK(cudaMalloc(&h->olog, (size_t)value * (size_t)h->batch));
      CK(cudaMemsetAsync(h->olog, 0xff, (size_t)value * (size_t)h->batch, h->stream));
    }
  } else if (!strcmp(key, "chain")) {
    h->chain = value ? 1 : 0;
  } else if (!strcmp(key, "chain_waves")) {
    if (value < 0) return fail(h, DQE_EINVAL, "chain_waves >= 0");
    h->chain_waves = (int)value;
  } else if (!strcmp(key, "smem_pad")) {
    if (value < 0 || value > 160 * 1024) return fail(h, DQE_EINVAL, "smem_pad in [0, 160 KiB]");
    h->smem_pad = (int)value;
  } else if (!strcmp(key, "chain_max")) {
    if (value < 0) return fail(h, DQE_EINVAL, "chain_max >= 0");
    h->chain_max = (int)value;
  } else if (!strcmp(key, "prefetch")) {
    if (value < -1 || value > (1 << 20)) return fail(h, DQE_EINVAL, "prefetch distance in [-1, 2^20] tiles");
    h->prefetch = (int)value;
  } else if (!strcmp(key, "tmap")) {
    h->tmap = value ? 1 : 0;
  } else if (!strcmp(key, "factor")) {
    h->factor = value ? 1 : 0;
  } else if (!strcmp(key, "cxperm")) {
    h->cxperm = value ? 1 : 0;
  } else if (!strcmp(key, "cluster")) {
    h->cluster = value ? 1 : 0;
  } else if (!strcmp(key, "persistent")) {
    h->persistent = value ? 1 : 0;
  } else if (!strcmp(key, "ctas_per_sm")) {
    h->ctas_per_sm = (int)value;
  } else if (!strcmp(key, "threads")) {
    if (value != 64 && value != 128 && value != 256) return fail(h, DQE_EINVAL, "threads in {64, 128, 256}");
    h->threads = (int)value; h->tile_auto = 0;
  } else return fail(h, DQE_EINVAL, std::string("unknown option ") + key);
  return 0;
}

int dqe_get_stats(dqe_t* h, int64_t* out) {
  if (!h || !out) return DQE_EINVAL;
  memcpy(out, h->stats, sizeof(h->stats));
  return 0;
}
int dqe_reset_stats(dqe_t* h) {
  if (!h) return DQE_EINVAL;
  memset(h->stats, 0, sizeof(h->stats));
  return 0;
}

int dqe_time_passes(dqe_t* h, int enable) {
  if (int rc = dqe_sync(h)) return rc;
  NEED_DEVICE(h);
  h->timing = enable != 0;
  h->t_ms = 0; h->t_passes = 0; h->t_amps = 0;
  return 0;
}
int dqe_pass_time_ms(dqe_t* h, double* total_ms, int64_t* n_passes, int64_t* amp_updates) {
  if (int rc = dqe_sync(h)) return rc;
  if (total_ms) *total_ms = h->t_ms;
  if (n_passes) *n_passes = h->t_passes;
  if (amp_updates) *amp_updates = h->t_amps;
  return 0;
}

}  // extern "C"
