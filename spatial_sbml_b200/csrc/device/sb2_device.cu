// sb2_device.cu — C-ABI device layer (see include/spatial_b200.h): device memory, the RPN → device
// bytecode compiler, and the per-step kernel schedule that replaces performStep/updateValues
// (reference spatialsimulator.cpp:1350-1544).
#include "sb2_internal.h"
#include "sb2_aux.h"
#include "sb2_mem.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <set>
#include <sstream>

namespace {
thread_local std::string g_create_error;
}

struct Sb2Species {
  int geo;
  double* u[2];
  int cur;
  double* k[4];
  int dkind[3], dsrc[3];
  int akind[3], asrc[3];
  int afacesrc[3];   // field holding a field-valued velocity at the plus faces along the same axis, -1: none
  int bckind[6], bcskind[6], bcsrc[6];
  bool has_diff, has_adv, has_bc;
  double* faces[3];  // CIP face values (advected species only)
  double* faces_alt[3];  // 2-D: the other half of their double buffer (a fused direction pass writes the new faces of its axis there)
  int face_par;          // bit a: the halves of axis a have been swapped an odd number of times
  // models with membrane transport: what the transports add to this species in a stage.  Zero everywhere except at the cells next
  // to a membrane, which the transport kernels rewrite in every stage - so it never needs clearing, and the stage kernel starts
  // its accumulators from it (Sb2SpeciesArgs::kcarry) instead of from a k_m that had to be zero-filled first
  double* tcarry;
};

struct Sb2Stmt {
  int geo, kind, n_reactants;
  std::vector<sb2_token> toks;
  std::vector<sb2_ref> refs;
};

struct sb2_sim {
  sb2_grid g;
  int64_t P, PL, first, nalloc;
  cudaStream_t stream;
  bool own_stream;
  std::vector<uint8_t*> d_flags;
  std::vector<std::vector<uint8_t> > h_flags;  // compact host copies (boundary lists, reductions)
  std::vector<Sb2Species> species;
  std::vector<double*> d_fields;
  double consts[SB2_MAX_CONSTS];
  int n_user_consts, n_internal_consts;
  std::vector<Sb2Stmt> stmts;
  std::vector<uint32_t> program;
  std::vector<Sb2MemTransport> transports;
  // species on membranes (sb2_mem.h) and the order in which transports and membrane reactions were declared: the
  // reference walks its reaction list once per stage, so contributions to a membrane species add up in that order
  std::vector<Sb2PointSet> point_sets;
  std::vector<Sb2MemSpecies> mem_species;
  std::vector<Sb2MemReaction> mem_reactions;
  struct MemOp { int type, index; };  // type 0: transports[index], 1: mem_reactions[index]
  std::vector<MemOp> mem_ops;
  std::vector<Sb2CellRule> cell_rules;  // assignment rules re-evaluated after every step, in dependency order
  Sb2BoundaryLists blists[3];  // every owned row / the edge rows of a slab / its other rows
  bool model_dirichlet;        // some species declares a Dirichlet side (whether or not this slab touches it)
  int edge_w, end_w;           // slab: owned rows per side computed first and exchanged per stage / rows exchanged at the end of a step
  // sharded stepping (sb2_sharded_begin / sb2_sharded_next)
  cudaStream_t sh_edge, sh_comm;
  cudaEvent_t ev_edge, ev_interior, ev_comm, ev_ready, ev_pulled;
  int sh_phase, sh_passes;
  int host_nslabs;       // sb2_step_host_begin / _finish: sub-slabs of the step in flight
  bool host_exchange;    // ... and whether the caller was handed an exchange of the edge rows
  bool halo_stale;       // slab: the halo rows of u were not refreshed after the last step (sb2_step_host_finish)
  bool sh_apron;         // slab: sharded steps exchange four rows of u once and recompute the apron rows (sb2_sharded_set_apron)
  // small grids: the CIP-CSLR passes of different species run side by side on streams of their own (created on first use)
  std::vector<cudaStream_t> adv_streams;
  std::vector<cudaEvent_t> adv_join;
  cudaEvent_t adv_fork;
  bool sh_awaiting, sh_comm_valid, sh_first_stage;
  bool finalized, any_adv;
  bool fuse_combine;    // this step: RK combine fused into stage 4
  bool carry_k0;        // this step: a non-zero Neumann flux is carried in k0 (spatialsimulator.cpp:1540-1542)
  bool k0_carry_valid;  // k0 currently holds the flux parked by the previous step's post-update pass
  double* d_tmp_face;   // CIP scratch
  Sb2StageCore base;
  uint32_t base_tok[SB2_MAX_TOKENS + 2];  // first-generation program (kernel parameter)
  Sb2RdTok* d_rdtok;                      // rd_stage token records on the device: two buffers, used alternately
  int rd_buf;                             // half of d_rdtok the current records live in
  bool multi_stream;                      // stages have been launched on more than one stream (sb2_stage_on)
  std::vector<Sb2RdTok> rd_uploaded;      // what d_rdtok currently holds
  std::string err;
  int launches;
  // split-phase step state
  double cur_dt;
  bool in_step;
  double* d_scratch;  // reduction scratch
  // rd_stage kernel (2-D): variant chosen at finalize and its strip-class map
  Sb2RdVariant rd;
  int rd_id;          // -1: first-generation stage kernel
  uint8_t* d_cls;
  int nstrips;
  // sb2_step_host: copy streams and events of the slab pipeline (created on first use)
  cudaStream_t h2d, d2h;
  std::vector<cudaEvent_t> ev_up, ev_done;
  cudaEvent_t ev_start;
  // sb2_step: CUDA graph of two consecutive steps (one per buffer parity) for small, launch-latency-bound grids
#define SB2_GRAPH_SLOTS 8
  cudaGraphExec_t gexec[SB2_GRAPH_SLOTS];
  double g_dt[SB2_GRAPH_SLOTS];
  uint64_t g_version[SB2_GRAPH_SLOTS];
  int g_launches[SB2_GRAPH_SLOTS];
  int64_t g_key[SB2_GRAPH_SLOTS];  // the buffer state (par_key) a slot's graph was captured in
  uint64_t version;  // bumped by everything that can change a kernel argument (constants, uploads that allocate)
  std::vector<Sb2RdTok> rdprog;  // lowered token records (constants patched per launch from rdslot)
  std::vector<int> rdslot;       // constants-table slot of each record's inline constant, -1 = none
  Sb2RdJit* jit;      // model-specialised build of the rd_stage kernel (rd_jit.cu), NULL: the interpreter kernel runs
  Sb2RdJitJob* jit_job;  // build running in the background (picked up at a step boundary)
  int jit_mode;          // 0 off, 1 build at finalize, 2 build in the background once the simulator has stepped for a while
  int64_t steps_seen;
  bool capturing;        // inside the stream capture of sb2_step
  std::string jit_info;
  uint32_t* d_dbg;    // SB2_DEBUG=1: device debug words, checked after every stage launch
  bool trace;         // SB2_TRACE=1: entry points report to stderr
};

namespace {

int fail(sb2_sim* s, int code, const std::string& msg) {
  if (s) s->err = msg; else g_create_error = msg;
  return code;
}
int cuda_fail(sb2_sim* s, cudaError_t e, const char* what) {
  std::ostringstream os;
  os << what << ": " << cudaGetErrorName(e) << " (" << cudaGetErrorString(e) << ")";
  return fail(s, SB2_ERR_CUDA, os.str());
}
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(sim, e__, #call); } while (0)

int64_t ncells(const sb2_sim* s) { return (int64_t)s->g.nx * s->g.ny * s->g.nz; }

// which halves of the double buffers are current (a captured step graph is valid for exactly one such state)
uint64_t par_key(const sb2_sim* s) {
  uint64_t k = 1;
  for (size_t i = 0; i < s->species.size(); ++i) k = k * 16 + (uint64_t)(s->species[i].cur * 8 + s->species[i].face_par);
  return k;
}

Sb2MemView mem_view(const sb2_sim* s) {
  Sb2MemView v;
  memset(&v, 0, sizeof(v));
  for (size_t i = 0; i < s->mem_species.size(); ++i) {
    v.u[i] = s->mem_species[i].u;
    for (int m = 0; m < 4; ++m) v.k[i][m] = s->mem_species[i].k[m];
  }
  return v;
}

template <typename T>
int upload_list(sb2_sim* sim, T** dptr, const T* h, size_t n) {
  *dptr = NULL;
  if (n == 0) return SB2_OK;
  CK(cudaMalloc((void**)dptr, sizeof(T) * n));
  CK(cudaMemcpyAsync(*dptr, h, sizeof(T) * n, cudaMemcpyHostToDevice, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return SB2_OK;
}

// compact host array → padded device layout (guard rows / planes, even pitch)
template <typename T>
int upload_padded(sb2_sim* sim, T* dst, const T* src) {
  const sb2_grid& g = sim->g;
  CK(cudaMemsetAsync(dst, 0, sizeof(T) * sim->nalloc, sim->stream));
  for (int z = 0; z < g.nz; ++z) {
    CK(cudaMemcpy2DAsync(dst + sim->first + (int64_t)z * sim->PL, sizeof(T) * sim->P,
                         src + (int64_t)z * g.nx * g.ny, sizeof(T) * g.nx, sizeof(T) * g.nx, g.ny,
                         cudaMemcpyHostToDevice, sim->stream));
  }
  return SB2_OK;
}
template <typename T>
int upload_interior(sb2_sim* sim, T* dst, const T* src) {
  const sb2_grid& g = sim->g;
  for (int z = 0; z < g.nz; ++z) {
    CK(cudaMemcpy2DAsync(dst + sim->first + (int64_t)z * sim->PL, sizeof(T) * sim->P,
                         src + (int64_t)z * g.nx * g.ny, sizeof(T) * g.nx, sizeof(T) * g.nx, g.ny,
                         cudaMemcpyHostToDevice, sim->stream));
  }
  return SB2_OK;
}
template <typename T>
int download_padded(sb2_sim* sim, T* dst, const T* src) {
  const sb2_grid& g = sim->g;
  for (int z = 0; z < g.nz; ++z) {
    CK(cudaMemcpy2DAsync(dst + (int64_t)z * g.nx * g.ny, sizeof(T) * g.nx,
                         src + sim->first + (int64_t)z * sim->PL, sizeof(T) * sim->P, sizeof(T) * g.nx, g.ny,
                         cudaMemcpyDeviceToHost, sim->stream));
  }
  CK(cudaStreamSynchronize(sim->stream));
  return SB2_OK;
}

int alloc_field(sb2_sim* sim, double** out) {
  CK(cudaMalloc((void**)out, sizeof(double) * sim->nalloc));
  CK(cudaMemsetAsync(*out, 0, sizeof(double) * sim->nalloc, sim->stream));
  return SB2_OK;
}

int internal_const(sb2_sim* sim, double v) {
  // internal constants (stoichiometries) grow down from the top of the table
  for (int i = 0; i < sim->n_internal_consts; ++i) {
    const int slot = SB2_MAX_CONSTS - 1 - i;
    if (memcmp(&sim->consts[slot], &v, sizeof(double)) == 0) return slot;
  }
  const int slot = SB2_MAX_CONSTS - 1 - sim->n_internal_consts;
  if (slot < sim->n_user_consts) return -1;
  sim->consts[slot] = v;
  sim->n_internal_consts++;
  return slot;
}

bool is_unary(int op) { return op >= SB2_OP_ABS && op <= SB2_OP_NOT; }
bool is_binary(int op) { return op >= SB2_OP_PLUS && op <= SB2_OP_NEQ; }

uint32_t enc(int opc, int d, int arg) { return SB2_CODE(opc, d) | ((uint32_t)arg << 16); }

// RPN (stack program) → register-addressed device bytecode.  Follows the reference interpreter's
// stack discipline token by token (calcPDE.cpp:244-449): operands push; an operator first pops
// (if the stack is non-empty), binary operators then combine into the new top only if something is
// left below, unary operators rewrite the popped slot and push it back, listed-but-empty operators
// and unresolved names only pop.  The rate is the top of the stack.
//
// Constants are pushed lazily: a constant operand never occupies a stack register, it is folded into
// the instruction that consumes it (r op c, or the commuted / reversed forms c + r, c * r, c - r,
// c / r, c < r ...; IEEE addition and multiplication commute bit for bit).  Registers are positional:
// the register of a materialised entry is the number of materialised entries below it.
int compile_stmt(sb2_sim* sim, const Sb2Stmt& st, std::vector<uint32_t>& prog) {
  // accumulation targets first: a law that changes nothing is not evaluated at all
  struct Target { int species; bool subtract; double stoich; };
  std::vector<Target> targets;
  if (st.kind == 0) {
    for (int k = 0; k < (int)st.refs.size(); ++k) {
      const sb2_ref& r = st.refs[k];
      if (!r.is_variable || r.species < 0) continue;
      if (r.species >= (int)sim->species.size()) return fail(sim, SB2_ERR_ARG, "reaction references unknown species");
      Target t = {r.species, k < st.n_reactants, r.stoichiometry};
      targets.push_back(t);
    }
  } else if (!st.refs.empty()) {
    const sb2_ref& r = st.refs[0];
    if (r.is_variable && r.species >= 0) {
      if (r.species >= (int)sim->species.size()) return fail(sim, SB2_ERR_ARG, "rate rule references unknown species");
      Target t = {r.species, false, r.stoichiometry};
      targets.push_back(t);
    }
  }
  if (targets.empty()) return SB2_OK;
  // The law acts on the cells of its own domain (spatialsimulator.cpp:1381-1390).  When every target
  // species lives in that same domain the kernel's output select already discards the other cells,
  // so no mask is needed.
  bool need_mask = false;
  for (size_t k = 0; k < targets.size(); ++k) need_mask = need_mask || sim->species[targets[k].species].geo != st.geo;

  std::vector<uint32_t> out;
  if (need_mask) out.push_back(enc(OPC_MASK, 0, st.geo));
  struct Ent { bool lazy; int cslot; };
  std::vector<Ent> stk;
  int nmat = 0;  // materialised entries on the stack
  bool too_deep = false;
  // What the reference's rpStack[0] holds: a law whose stream runs the stack empty (any law with a listed-but-empty operator
  // does: such an operator pops without pushing) ends with st_index == 0 and takes rpStack[0], the value last stored at the
  // bottom of the stack, as its rate (calcPDE.cpp:430-431).  valid: something was stored there during this evaluation.
  struct { bool valid, lazy; int cslot; } slot0 = {false, false, 0};
  auto emit = [&](int opc, int d, int arg) {
    if (d < 0 || d >= SB2_MAX_DEPTH || (opc == OPC_MOVUP && d + 1 >= SB2_MAX_DEPTH)) too_deep = true;
    out.push_back(enc(opc, d < 0 ? 0 : d, arg));
  };
  auto push_mat = [&]() { Ent e = {false, 0}; stk.push_back(e); ++nmat; };
  auto pop = [&]() { if (!stk.back().lazy) --nmat; stk.pop_back(); };
  const int zero_slot = internal_const(sim, 0.0);
  if (zero_slot < 0) return fail(sim, SB2_ERR_LIMIT, "constants table full");

  const int n = (int)st.toks.size();
  for (int i = 0; i < n; ++i) {
    const sb2_token& t = st.toks[i];
    if (t.kind == SB2_TOK_VAR || t.kind == SB2_TOK_FIELD || t.kind == SB2_TOK_CONST) {
      if (t.kind == SB2_TOK_VAR && t.slot >= sim->species.size()) return fail(sim, SB2_ERR_ARG, "token references unknown species");
      if (t.kind == SB2_TOK_FIELD && t.slot >= sim->d_fields.size()) return fail(sim, SB2_ERR_ARG, "token references unknown field");
      if (t.kind == SB2_TOK_CONST && t.slot >= SB2_MAX_CONSTS) return fail(sim, SB2_ERR_ARG, "token references constant slot out of range");
      if (t.kind == SB2_TOK_CONST) {
        if (stk.empty()) { slot0.valid = true; slot0.lazy = true; slot0.cslot = (int)t.slot; }
        Ent e = {true, (int)t.slot}; stk.push_back(e); continue;
      }
      // species operand immediately consumed by + - * / on a materialised top → fused form acting on that register
      if (t.kind == SB2_TOK_VAR && !stk.empty() && !stk.back().lazy && i + 1 < n && st.toks[i + 1].kind == SB2_TOK_OP) {
        int opc = -1;
        switch (st.toks[i + 1].op) {
          case SB2_OP_PLUS: opc = OPC_ADDV; break;
          case SB2_OP_MINUS: opc = OPC_SUBV; break;
          case SB2_OP_TIMES: opc = OPC_MULV; break;
          case SB2_OP_DIVIDE: opc = OPC_DIVV; break;
          default: break;
        }
        if (opc >= 0) { emit(opc, nmat - 1, t.slot); ++i; continue; }
      }
      if (stk.empty()) { slot0.valid = true; slot0.lazy = false; }
      emit(t.kind == SB2_TOK_VAR ? OPC_PUSHV : OPC_PUSHF, nmat, t.slot);
      push_mat();
      continue;
    }
    // operator (or unresolved name): pops the top first, whatever follows
    const bool had_top = !stk.empty();
    Ent b = {false, 0};
    if (had_top) { b = stk.back(); pop(); }
    if (t.kind != SB2_TOK_OP) continue;  // unresolved name: pop only
    const int op = t.op;
    if (is_binary(op)) {
      if (!had_top || stk.empty()) continue;  // nothing below: the popped value is lost, the stack stays as it is
      Ent a = stk.back();
      int opc_rr = -1, opc_rc = -1, opc_cr = -1;  // reg op reg, reg op const, const op reg
      switch (op) {
        case SB2_OP_PLUS: opc_rr = OPC_ADD; opc_rc = OPC_ADDC; opc_cr = OPC_ADDC; break;
        case SB2_OP_MINUS: opc_rr = OPC_SUB; opc_rc = OPC_SUBC; opc_cr = OPC_RSUBC; break;
        case SB2_OP_TIMES: opc_rr = OPC_MUL; opc_rc = OPC_MULC; opc_cr = OPC_MULC; break;
        case SB2_OP_DIVIDE: opc_rr = OPC_DIV; opc_rc = OPC_DIVC; opc_cr = OPC_RDIVC; break;
        case SB2_OP_GEQ: opc_rr = OPC_GEQ; opc_rc = OPC_GEQC; opc_cr = OPC_LEQC; break;  // c >= r  <=>  r <= c
        case SB2_OP_GT: opc_rr = OPC_GT; opc_rc = OPC_GTC; opc_cr = OPC_LTC; break;
        case SB2_OP_LEQ: opc_rr = OPC_LEQ; opc_rc = OPC_LEQC; opc_cr = OPC_GEQC; break;
        case SB2_OP_LT: opc_rr = OPC_LT; opc_rc = OPC_LTC; opc_cr = OPC_GTC; break;
        case SB2_OP_AND: opc_rr = OPC_AND; break;
        case SB2_OP_OR: opc_rr = OPC_OR; break;
        default: break;
      }
      // nmat counts the entries still on the stack (b has been popped)
      if (a.lazy && b.lazy) {  // materialise a, then r op c
        emit(OPC_PUSHC, nmat, a.cslot);
        stk.back().lazy = false; ++nmat;
        a = stk.back();
      }
      if (!a.lazy && !b.lazy) {  // a in register nmat-1, b in register nmat
        if (opc_rr >= 0) emit(opc_rr, nmat, 0); else emit(OPC_BINR, nmat, op);
      } else if (!a.lazy && b.lazy) {
        if (opc_rc >= 0) emit(opc_rc, nmat - 1, b.cslot);
        else { emit(OPC_PUSHC, nmat, b.cslot); if (opc_rr >= 0) emit(opc_rr, nmat, 0); else emit(OPC_BINR, nmat, op); }
      } else {  // a lazy, b in register nmat (it sits where a would have been)
        if (opc_cr >= 0) emit(opc_cr, nmat, a.cslot);
        else { emit(OPC_MOVUP, nmat, 0); emit(OPC_PUSHC, nmat, a.cslot); if (opc_rr >= 0) emit(opc_rr, nmat + 1, 0); else emit(OPC_BINR, nmat + 1, op); }
        stk.back().lazy = false; ++nmat;
      }
      if (stk.size() == 1) { slot0.valid = true; slot0.lazy = false; }  // the result sits at the bottom, in register 0
    } else if (is_unary(op)) {
      // st--; s[st] = f(s[st]); st++  (an empty stack leaves a stale slot 0 in the reference; 0.0 here)
      if (!had_top) emit(OPC_PUSHC, nmat, zero_slot);
      else if (b.lazy) emit(OPC_PUSHC, nmat, b.cslot);
      emit(OPC_UNARY, nmat, op);
      if (stk.empty()) { slot0.valid = true; slot0.lazy = false; }
      push_mat();
    }
    // SB2_OP_POP_ONLY and anything else: pop only
  }
  // rate = top of the stack
  bool const_rate = false;
  int rate_cslot = zero_slot;
  if (stk.empty()) {
    // stale bottom slot: a constant, or the value a materialised entry left in register 0 (nothing has overwritten it: every
    // later store to the bottom of the stack updated slot0, and higher entries live in higher registers)
    if (slot0.valid && slot0.lazy) { const_rate = true; rate_cslot = slot0.cslot; }
    else if (!slot0.valid) const_rate = true;
  }
  else if (stk.back().lazy) { const_rate = true; rate_cslot = stk.back().cslot; }
  else if (nmat - 1 != 0) emit(OPC_MOV0, nmat - 1, 0);
  if (too_deep) return fail(sim, SB2_ERR_LIMIT, "expression stack deeper than SB2_MAX_DEPTH");

  bool all_unit = true;
  for (size_t k = 0; k < targets.size(); ++k) all_unit = all_unit && targets[k].stoich == 1.0;
  if (const_rate && all_unit && !need_mask) {
    // a law that is a single constant with unit stoichiometries needs no stack at all
    for (size_t k = 0; k < targets.size(); ++k) out.push_back(enc(targets[k].subtract ? OPC_ACCSUBC1 : OPC_ACCADDC1, targets[k].species, rate_cslot));
  } else {
    if (const_rate) out.push_back(enc(OPC_PUSHC, 0, rate_cslot));
    for (size_t k = 0; k < targets.size(); ++k) {
      const Target& t = targets[k];
      if (!need_mask && t.stoich == 1.0) { out.push_back(enc(t.subtract ? OPC_ACCSUB1 : OPC_ACCADD1, t.species, 0)); continue; }
      const int slot = internal_const(sim, t.stoich);
      if (slot < 0) return fail(sim, SB2_ERR_LIMIT, "constants table full");
      if (need_mask) out.push_back(enc(t.subtract ? OPC_ACCSUBM : OPC_ACCADDM, t.species, slot));
      else out.push_back(enc(t.subtract ? OPC_ACCSUB : OPC_ACCADD, t.species, slot));
    }
  }
  prog.insert(prog.end(), out.begin(), out.end());
  return SB2_OK;
}

// Lowering of the register bytecode for the rd_stage kernel: 16-byte records whose constants travel inline, and
// three peephole fusions that save a dispatch each —
//   PUSHV d v ; <op>C d c   →  <op>VC d v c      (r[d] = v op c, or c op v for the reversed forms)
//   PUSHV d v1 ; <op>V d v2 →  <op>VV d v1 v2
//   ACCSUB1 s1 ; ACCADD1 s2 →  XFER s1 s2        (reactant and product of a unit-stoichiometry law)
void rd_lower(const std::vector<uint32_t>& prog, std::vector<Sb2RdTok>& out, std::vector<int>& slots) {
  out.clear(); slots.clear();
  auto opc_of = [](uint32_t t) { return (int)((t & 0xffffu) / SB2_MAX_DEPTH); };
  auto d_of = [](uint32_t t) { return (int)((t & 0xffffu) % SB2_MAX_DEPTH); };
  auto emit = [&](int opc, int d, uint32_t arg, int slot) {
    Sb2RdTok r; r.code = SB2_CODE(opc, d); r.arg = arg; r.c = 0.0;
    out.push_back(r); slots.push_back(slot);
  };
  auto has_const = [](int opc) {
    return opc == OPC_PUSHC || (opc >= OPC_ADDC && opc <= OPC_LTC) || opc == OPC_ACCSUB || opc == OPC_ACCADD ||
           opc == OPC_ACCSUBC1 || opc == OPC_ACCADDC1 || opc == OPC_ACCSUBM || opc == OPC_ACCADDM;
  };
  for (size_t i = 0; i < prog.size(); ++i) {
    const uint32_t t = prog[i];
    const int opc = opc_of(t), d = d_of(t);
    const uint32_t arg = t >> 16;
    if (i + 1 < prog.size()) {
      const uint32_t n = prog[i + 1];
      const int nopc = opc_of(n), nd = d_of(n);
      const uint32_t narg = n >> 16;
      if (opc == OPC_PUSHV && nd == d) {
        int f = -1;
        switch (nopc) {
          case OPC_ADDC: f = OPC_ADDVC; break;
          case OPC_SUBC: f = OPC_SUBVC; break;
          case OPC_RSUBC: f = OPC_RSUBVC; break;
          case OPC_MULC: f = OPC_MULVC; break;
          case OPC_DIVC: f = OPC_DIVVC; break;
          case OPC_RDIVC: f = OPC_RDIVVC; break;
          default: break;
        }
        if (f >= 0) { emit(f, d, arg, (int)narg); ++i; continue; }
        switch (nopc) {
          case OPC_ADDV: f = OPC_ADDVV; break;
          case OPC_SUBV: f = OPC_SUBVV; break;
          case OPC_MULV: f = OPC_MULVV; break;
          case OPC_DIVV: f = OPC_DIVVV; break;
          default: break;
        }
        if (f >= 0) { emit(f, d, arg | (narg << 8), -1); ++i; continue; }
      }
      if (opc == OPC_ACCSUB1 && nopc == OPC_ACCADD1) { emit(OPC_XFER, d, (uint32_t)nd, -1); ++i; continue; }
    }
    if (has_const(opc)) emit(opc, d, 0, (int)arg);
    else emit(opc, d, arg, -1);
  }
}

const char* kOpcNames[OPC_COUNT] = {"END", "PUSHC", "PUSHV", "PUSHF", "ADD", "SUB", "MUL", "DIV", "GEQ", "GT",
  "LEQ", "LT", "AND", "OR", "ADDC", "SUBC", "MULC", "DIVC", "RSUBC", "RDIVC", "GEQC", "GTC", "LEQC", "LTC", "ADDV", "SUBV", "MULV",
  "DIVV", "UNARY", "BINR", "MOV0", "MOVUP", "MASK", "ACCSUB", "ACCADD", "ACCSUB1", "ACCADD1", "ACCSUBC1", "ACCADDC1", "ACCSUBM", "ACCADDM"};

// Membrane transport of stage m goes to the species' transport buffers (no zero-filling of k_m) unless k_0 carries a Neumann flux
// from the previous step, which the transports then add to in place
bool transport_in_buffer(const sb2_sim* sim, int m) { return !sim->transports.empty() && !(sim->carry_k0 && m == 0); }

void fill_stage_args(sb2_sim* sim, Sb2StageCore& a, int m, double dt, int row_begin, int row_end) {
  static const double rk[4] = {0.0, 0.5, 0.5, 1.0};
  a = sim->base;
  a.m = m;
  a.dt = dt;
  a.cdt = rk[m] * dt;
  a.final = (m == 3 && sim->fuse_combine) ? 1 : 0;
  // (cells next to a membrane are boundary cells of their geometry: the transport buffer is zero in every strip row that is all
  // interior or all outside, so those rows need not read it)
  a.carry = (sim->carry_k0 && m == 0) ? 1 : transport_in_buffer(sim, m) ? 2 : 0;
  a.row_begin = row_begin;
  a.row_end = row_end;
  const int S = (int)sim->species.size();
  for (int s = 0; s < S; ++s) {
    Sb2Species& sp = sim->species[s];
    Sb2SpeciesArgs& o = a.sp[s];
    for (int ax = 0; ax < 3; ++ax) o.dconst[ax] = sp.dkind[ax] == SB2_SRC_CONST ? sim->consts[sp.dsrc[ax]] : 0.0;
    o.u = sp.u[sp.cur];
    o.u_out = sp.u[sp.cur ^ 1];
    o.kprev = m > 0 ? sp.k[m - 1] : sp.k[0];
    o.kout = sp.k[m];
    o.kcarry = transport_in_buffer(sim, m) ? sp.tcarry : sp.k[m];
    o.k0 = sp.k[0];
    o.k1 = sp.k[1];
  }
  // rows per CTA: enough CTAs to fill the machine several times over, long enough strips to
  // amortise the two extra rows each CTA loads
  if (sim->rd_id >= 0) {
    // a CTA walks its rows in batches of W.  Measured at 8192^2 (ms per step): with the per-batch CTA barrier 3, 4, 6, 8,
    // 16, 32, 64 batches per tile → 3.53, 3.43, 3.33, 3.33, 3.40, 3.75, 4.31; with row-ready barriers 4, 6, 8, 12, 16 →
    // 3.13, 3.00, 2.97, 2.94, 2.99: short tiles keep the rows in flight close together in memory and the waves even,
    // the prologue batch they pay for is cheap
    const int W = sim->rd.w;
    a.np = sim->rd.np;
    if (sim->g.dim == 3) {
      // 3-D: a CTA owns W rows of a strip and marches over rows_per_block planes; each CTA stages two planes it does
      // not compute, so long marches are cheap, short ones fill the machine on small grids
      static const int zpb_default = getenv("SB2_RD_ZPB") ? atoi(getenv("SB2_RD_ZPB")) : 32;
      const int planes = row_end - row_begin;
      const int64_t tiles = (int64_t)((sim->g.nx + 64 * sim->rd.np - 1) / (64 * sim->rd.np)) * ((sim->g.ny + W - 1) / W);
      int zpb = zpb_default < 1 ? 1 : zpb_default;
      while (zpb > 1 && tiles * ((planes + zpb - 1) / zpb) < 148 * 6) zpb = (zpb + 1) / 2;
      a.rows_per_block = zpb;
      return;
    }
    const int rows = row_end - row_begin;
    static const int rpb_batches = getenv("SB2_RD_BATCHES") ? atoi(getenv("SB2_RD_BATCHES")) : 12;
    int rpb = rpb_batches * W;
    while (rpb > W && (int64_t)sim->nstrips * ((rows + rpb - 1) / rpb) < 148 * 6) rpb -= W;
    // Grids of a few waves of CTAs (a slab of a strongly scaled grid: 8192 x 1024 cells are 2.4 waves of 12-batch tiles) lose
    // up to a third to the last, partly filled wave.  Where another tile height within 4 .. 12 batches makes waves x (rows of a
    // tile + its prologue batch) more than 3 % smaller, take it; large grids (many waves) keep 12 batches.
    static const bool wavefit = getenv("SB2_RD_NO_WAVEFIT") == NULL;
    if (wavefit && (int64_t)sim->nstrips * rows >= (int64_t)148 * 2 * W * 4) {
      const int per_sm = (m == 0 && sim->jit && ncells(sim) >= (1 << 20) && sim->rd.w == 8 && sim->base.S <= 2) ? 3 : 2;
      const int64_t slots = 148 * per_sm;
      auto cost = [&](int r) {
        const int64_t tiles = (int64_t)sim->nstrips * ((rows + r - 1) / r);
        return ((tiles + slots - 1) / slots) * (int64_t)(r + W);
      };
      int best = rpb;
      for (int b = rpb_batches; b >= 4; --b)
        if (cost(b * W) * 100 < cost(best) * 97) best = b * W;
      rpb = best;
    }
    a.rows_per_block = rpb;
    return;
  }
  a.np = sb2_stage_pairs(sim->g.nx, S);
  const int strips = (sim->g.nx + 256 * a.np - 1) / (256 * a.np);
  const int rows = (sim->g.dim == 3) ? sim->g.ny : (row_end - row_begin);
  const int planes = (sim->g.dim == 3) ? (row_end - row_begin) : 1;
  int rpb = 64;
  while (rpb > 8 && (int64_t)strips * planes * ((rows + rpb - 1) / rpb) < 148 * 8) rpb /= 2;
  a.rows_per_block = rpb;
}

// Model-specialised stage kernel (rd_jit.cu).  background = false: build now (about a second).  background = true:
// start the build on another thread, or - when one is running - install its result if it is ready (wait = true: block
// until it is).  Installing bumps `version`: captured step graphs are rebuilt with the new kernel.
Sb2RdJitSpec jit_spec(const sb2_sim* sim) {
  Sb2RdJitSpec spec;
  spec.ns = sim->base.S; spec.np = sim->rd.np; spec.w = sim->rd.w; spec.depth = sim->base.depth; spec.d3 = sim->g.dim == 3;
  spec.minb = 2; spec.np_final = 0; spec.minb_final = 0; spec.final_ldg = 0;
  // large 2-D grids: the first stage (no k rows) at three CTAs per SM; small grids are bound by latency and lose a microsecond
  // per step to the tighter register budget
  spec.minb_first = (!spec.d3 && spec.ns <= 2 && spec.w == 8 && ncells(sim) >= (1 << 20)) ? 3 : 0;
  if (spec.d3 && spec.w == 8 && spec.np == 1) {
    // 3-D (the class map has 64-cell granularity, so the strip width is free per stage).  Last stage: 2 cells per
    // thread, 74-80 registers, three CTAs per SM (deep expression stacks keep the room of two).  Other stages, up to
    // two species: 4 cells per thread, two CTAs per SM (106 KB of shared memory each).
    spec.np_final = 1; spec.minb_final = spec.depth <= 4 ? 3 : 2;
    if (spec.ns <= 2 && spec.depth <= 4 && !getenv("SB2_JIT_NP1")) { spec.np = 2; spec.minb = 2; }
    else spec.minb = spec.minb_final;
    // experiments: the last stage with four cells per thread and k0 / k1 read by plain loads (two CTAs of 105 KB per SM)
    if (const char* e = getenv("SB2_JIT_FINAL_LDG")) {
      spec.final_ldg = 1;
      if (atoi(e) >= 2 && spec.np == 2) { spec.np_final = 2; spec.minb_final = 2; }
    }
  }
  if (const char* e = getenv("SB2_JIT_MINB")) spec.minb = atoi(e) > 0 ? atoi(e) : 2;  // experiments
  spec.fold = getenv("SB2_JIT_NOFOLD") ? 0 : 1;
  spec.G = sim->base.G;
  spec.hasdiff_mask = spec.fastpath_mask = spec.geo_bits = 0;
  for (int s = 0; s < sim->base.S; ++s) {
    if (sim->base.sp[s].has_diff) spec.hasdiff_mask |= 1u << s;
    if (sim->base.sp[s].fastpath) spec.fastpath_mask |= 1u << s;
    spec.geo_bits |= (unsigned)(sim->base.sp[s].geo & 3) << (2 * s);
  }
  return spec;
}
int specialise(sb2_sim* sim, bool background, bool wait) {
  if (sim->jit || sim->rd_id < 0) return SB2_OK;
  std::ostringstream os;
  std::string jerr;
  if (!background && !sim->jit_job) {
    const auto t0 = std::chrono::steady_clock::now();
    sim->jit = sb2_rd_jit_build(jit_spec(sim), sim->rdprog, &jerr);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (sim->jit) os << "specialised kernel: " << sim->rdprog.size() << " records compiled in " << secs << " s with " << sb2_rd_jit_library();
    else { os << "interpreter kernel: specialised build failed: " << jerr; sim->jit_mode = 0; }
  } else {
    if (!sim->jit_job) {
      sim->jit_job = sb2_rd_jit_start(jit_spec(sim), sim->rdprog);
      sim->jit_info = "interpreter kernel: specialised build running in the background";
      if (!wait) return SB2_OK;
    }
    if (!wait && sb2_rd_jit_poll(sim->jit_job) == 0) return SB2_OK;
    double secs = 0;
    sim->jit = sb2_rd_jit_finish(sim->jit_job, &jerr, &secs);
    sim->jit_job = NULL;
    if (sim->jit) os << "specialised kernel: " << sim->rdprog.size() << " records compiled in " << secs << " s in the background with "
                     << sb2_rd_jit_library() << ", in use from step " << sim->steps_seen;
    else { os << "interpreter kernel: specialised build failed: " << jerr; sim->jit_mode = 0; }
  }
  sim->jit_info = os.str();
  if (sim->jit) ++sim->version;
  if (sim->trace) fprintf(stderr, "[sb2] %s\n", sim->jit_info.c_str());
  return SB2_OK;
}

// One stage on rows [row_begin, row_end) of the slab axis with whichever kernel family the model runs on.
int launch_stage_rows(sb2_sim* sim, int m, double dt, int row_begin, int row_end) {
  if (row_end <= row_begin) return SB2_OK;
  cudaError_t e;
  if (sim->rd_id >= 0) {
    Sb2RdArgs a;
    fill_stage_args(sim, a, m, dt, row_begin, row_end);
    if (sim->jit) {
      // specialised kernel: the constants of the records travel in the kernel parameters
      double kc[SB2_RD_MAX_TOKENS];
      const int n = (int)sim->rdprog.size();
      for (int i = 0; i < n; ++i) kc[i] = sim->rdslot[i] >= 0 ? sim->consts[sim->rdslot[i]] : 0.0;
      a.rdtok = NULL;
      if (sim->trace) fprintf(stderr, "[sb2] stage m=%d rows [%d,%d) rpb %d final %d carry %d (specialised)\n", m, row_begin, row_end, a.rows_per_block, a.final, a.carry);
      ++sim->launches;
      e = sb2_rd_jit_launch(sim->jit, a, kc, n, sim->stream);
      if (e != cudaSuccess) return cuda_fail(sim, e, "specialised stage kernel launch");
      goto launched;
    }
    {
    // inline constants of the token records = current values of the constants table; upload only when they changed
    std::vector<Sb2RdTok> cur(sim->rdprog);
    for (size_t i = 0; i < sim->rdslot.size(); ++i)
      if (sim->rdslot[i] >= 0) cur[i].c = sim->consts[sim->rdslot[i]];
    Sb2RdTok end; end.code = 0; end.arg = 0; end.c = 0.0;
    cur.push_back(end); cur.push_back(end);
    if (cur.size() != sim->rd_uploaded.size() || memcmp(cur.data(), sim->rd_uploaded.data(), sizeof(Sb2RdTok) * cur.size()) != 0) {
      // new records go to the other half: kernels already queued keep reading the half they were launched with
      sim->rd_buf ^= 1;
      e = cudaMemcpyAsync(sim->d_rdtok + sim->rd_buf * (SB2_RD_MAX_TOKENS + 2), cur.data(), sizeof(Sb2RdTok) * cur.size(),
                          cudaMemcpyHostToDevice, sim->stream);  // pageable source: staged before the call returns
      if (e != cudaSuccess) return cuda_fail(sim, e, "token upload");
      // kernels of this stage may run on another stream: they must not start before the records have landed
      if (sim->multi_stream) CK(cudaStreamSynchronize(sim->stream));
      sim->rd_uploaded.swap(cur);
    }
    a.rdtok = sim->d_rdtok + sim->rd_buf * (SB2_RD_MAX_TOKENS + 2);
    if (sim->trace) fprintf(stderr, "[sb2] stage m=%d rows [%d,%d) rpb %d final %d carry %d\n", m, row_begin, row_end, a.rows_per_block, a.final, a.carry);
    ++sim->launches;
    e = sb2_rd_launch(a, sim->stream);
    }
  } else {
    Sb2StageArgs a;
    fill_stage_args(sim, a, m, dt, row_begin, row_end);
    memcpy(a.tok, sim->base_tok, sizeof(a.tok));
    memcpy(a.consts, sim->consts, sizeof(a.consts));
    e = sb2_launch_stage(a, sim->stream, &sim->launches);
  }
  if (e != cudaSuccess) return cuda_fail(sim, e, "stage kernel launch");
launched:
  if (sim->d_dbg) {
    uint32_t h[8];
    CK(cudaStreamSynchronize(sim->stream));
    CK(cudaMemcpy(h, sim->d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
    if (h[0]) {
      std::ostringstream os;
      os << "rd_stage: " << h[0] << " bulk-copy waits timed out; first: strip " << h[1] << " tile " << h[2] << " warp " << h[3]
         << " row " << (int)h[4] << " parity " << h[5] << " tile rows [" << (int)h[6] << "," << (int)h[7] << ")";
      return fail(sim, SB2_ERR_CUDA, os.str());
    }
  }
  return SB2_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* sb2_last_error(const sb2_sim* sim) { return sim ? sim->err.c_str() : g_create_error.c_str(); }

int sb2_create(const sb2_grid* grid, sb2_sim** out) {
  if (!grid || !out) return fail(NULL, SB2_ERR_ARG, "null argument");
  *out = NULL;
  if (grid->nx < 1 || grid->ny < 1 || grid->nz < 1) return fail(NULL, SB2_ERR_ARG, "grid dimensions must be positive");
  if (grid->dim != 2 && grid->dim != 3) return fail(NULL, SB2_ERR_ARG, "only 2-D and 3-D geometries are supported on the device");
  if (grid->dim == 2 && grid->nz != 1) return fail(NULL, SB2_ERR_ARG, "nz must be 1 for a 2-D grid");
  if (grid->own_hi > grid->own_lo) {
    const int n_held = grid->dim == 3 ? grid->nz : grid->ny;
    if (grid->own_lo < 0 || grid->own_hi > n_held || grid->row0 < 0) return fail(NULL, SB2_ERR_ARG, "owned row range outside the held rows");
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(NULL, SB2_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (grid->device < 0 || grid->device >= ndev) return fail(NULL, SB2_ERR_ARG, "device ordinal out of range");
  e = cudaSetDevice(grid->device);
  if (e != cudaSuccess) return fail(NULL, SB2_ERR_CUDA, cudaGetErrorString(e));
  sb2_sim* sim = new sb2_sim();
  sim->g = *grid;
  // Rows start on 32-byte boundaries (pitch a multiple of 4 doubles), one zero guard row before and after
  // every plane, one guard plane either side in 3-D.  The stage kernel's bulk row copies start two cells
  // left of a strip and may run past the end of the last strip, so a front pad and a tail pad keep every
  // copy inside the allocation.
  sim->P = (grid->nx + 3) / 4 * 4;
  sim->PL = sim->P * (grid->ny + 2);
  sim->first = (grid->dim == 2 ? sim->P : sim->PL + sim->P) + SB2_FRONT_PAD;
  sim->nalloc = SB2_FRONT_PAD + (grid->dim == 2 ? sim->PL : sim->PL * (grid->nz + 2)) + SB2_TAIL_PAD;
  sim->stream = 0; sim->own_stream = false;
  e = cudaStreamCreateWithFlags(&sim->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete sim; return fail(NULL, SB2_ERR_CUDA, cudaGetErrorString(e)); }
  sim->own_stream = true;
  memset(sim->consts, 0, sizeof(sim->consts));
  sim->n_user_consts = 0; sim->n_internal_consts = 0;
  sim->finalized = false; sim->fuse_combine = true; sim->any_adv = false; sim->carry_k0 = false;
  sim->k0_carry_valid = false; sim->d_tmp_face = NULL;
  sim->launches = 0; sim->in_step = false; sim->cur_dt = 0; sim->d_scratch = NULL;
  sim->rd_id = -1; sim->d_cls = NULL; sim->nstrips = 0; sim->d_dbg = NULL;
  sim->h2d = NULL; sim->d2h = NULL; sim->ev_start = NULL; sim->d_rdtok = NULL; sim->rd_buf = 0; sim->multi_stream = false;
  for (int p = 0; p < SB2_GRAPH_SLOTS; ++p) { sim->gexec[p] = NULL; sim->g_key[p] = -1; }
  sim->version = 1; sim->jit = NULL;
  sim->jit_job = NULL; sim->jit_mode = 0; sim->steps_seen = 0; sim->capturing = false;
  sim->trace = getenv("SB2_TRACE") != NULL;
  memset(&sim->base, 0, sizeof(sim->base));
  memset(sim->blists, 0, sizeof(sim->blists));
  sim->edge_w = 1; sim->end_w = 1;
  sim->sh_edge = NULL; sim->sh_comm = NULL;
  sim->ev_edge = sim->ev_interior = sim->ev_comm = sim->ev_ready = sim->ev_pulled = NULL;
  sim->host_nslabs = 0; sim->host_exchange = false; sim->halo_stale = false; sim->sh_apron = false;
  sim->adv_fork = NULL;
  sim->sh_phase = -1; sim->sh_passes = 0; sim->sh_awaiting = false; sim->sh_comm_valid = false; sim->sh_first_stage = true;
  *out = sim;
  return SB2_OK;
}

void sb2_destroy(sb2_sim* sim) {
  if (!sim) return;
  cudaSetDevice(sim->g.device);
  cudaStreamSynchronize(sim->stream);
  for (size_t i = 0; i < sim->d_flags.size(); ++i) cudaFree(sim->d_flags[i]);
  for (size_t i = 0; i < sim->d_fields.size(); ++i) cudaFree(sim->d_fields[i]);
  for (size_t i = 0; i < sim->species.size(); ++i) {
    Sb2Species& s = sim->species[i];
    cudaFree(s.u[0]); cudaFree(s.u[1]);
    for (int m = 0; m < 4; ++m) cudaFree(s.k[m]);
    for (int a = 0; a < 3; ++a) { if (s.faces[a]) cudaFree(s.faces[a]); if (s.faces_alt[a]) cudaFree(s.faces_alt[a]); }
    if (s.tcarry) cudaFree(s.tcarry);
  }
  for (size_t i = 0; i < sim->transports.size(); ++i) sb2_free_transport(sim->transports[i]);
  for (int part = 0; part < 3; ++part) sb2_free_boundary_lists(sim->blists[part]);
  for (size_t i = 0; i < sim->adv_streams.size(); ++i) cudaStreamDestroy(sim->adv_streams[i]);
  for (size_t i = 0; i < sim->adv_join.size(); ++i) cudaEventDestroy(sim->adv_join[i]);
  if (sim->adv_fork) cudaEventDestroy(sim->adv_fork);
  if (sim->sh_edge) cudaStreamDestroy(sim->sh_edge);
  if (sim->sh_comm) cudaStreamDestroy(sim->sh_comm);
  if (sim->ev_edge) cudaEventDestroy(sim->ev_edge);
  if (sim->ev_interior) cudaEventDestroy(sim->ev_interior);
  if (sim->ev_comm) cudaEventDestroy(sim->ev_comm);
  if (sim->ev_ready) cudaEventDestroy(sim->ev_ready);
  if (sim->ev_pulled) cudaEventDestroy(sim->ev_pulled);
  for (size_t i = 0; i < sim->mem_reactions.size(); ++i) sb2_free_mem_reaction(sim->mem_reactions[i]);
  for (size_t i = 0; i < sim->cell_rules.size(); ++i) sb2_free_cell_rule(sim->cell_rules[i]);
  for (size_t i = 0; i < sim->point_sets.size(); ++i) {
    cudaFree(sim->point_sets[i].d_fill_target); cudaFree(sim->point_sets[i].d_fill_a); cudaFree(sim->point_sets[i].d_fill_b);
  }
  for (size_t i = 0; i < sim->mem_species.size(); ++i) {
    Sb2MemSpecies& ms = sim->mem_species[i];
    cudaFree(ms.u);
    for (int m = 0; m < 4; ++m) cudaFree(ms.k[m]);
    cudaFree(ms.d_diff_pts); cudaFree(ms.d_term_start); cudaFree(ms.d_adj); cudaFree(ms.d_s); cudaFree(ms.d_d); cudaFree(ms.d_area);
    cudaFree(ms.d_dvalues); cudaFree(ms.d_update_pts);
  }
  if (sim->d_scratch) cudaFree(sim->d_scratch);
  if (sim->d_cls) cudaFree(sim->d_cls);
  if (sim->d_rdtok) cudaFree(sim->d_rdtok);
  sb2_rd_jit_free(sim->jit);
  if (sim->jit_job) sb2_rd_jit_abandon(sim->jit_job);
  for (int p = 0; p < SB2_GRAPH_SLOTS; ++p) if (sim->gexec[p]) cudaGraphExecDestroy(sim->gexec[p]);
  if (sim->h2d) cudaStreamDestroy(sim->h2d);
  if (sim->d2h) cudaStreamDestroy(sim->d2h);
  if (sim->ev_start) cudaEventDestroy(sim->ev_start);
  for (size_t i = 0; i < sim->ev_up.size(); ++i) cudaEventDestroy(sim->ev_up[i]);
  for (size_t i = 0; i < sim->ev_done.size(); ++i) cudaEventDestroy(sim->ev_done[i]);
  if (sim->d_dbg) cudaFree(sim->d_dbg);
  if (sim->d_tmp_face) cudaFree(sim->d_tmp_face);
  if (sim->own_stream) cudaStreamDestroy(sim->stream);
  delete sim;
}

int sb2_set_stream(sb2_sim* sim, void* cuda_stream) {
  if (!sim) return SB2_ERR_ARG;
  ++sim->version;
  cudaStreamSynchronize(sim->stream);
  if (sim->own_stream) { cudaStreamDestroy(sim->stream); sim->own_stream = false; }
  sim->stream = (cudaStream_t)cuda_stream;
  return SB2_OK;
}

int sb2_add_geometry(sb2_sim* sim, const uint8_t* cell_flags) {
  if (!sim || !cell_flags) return fail(sim, SB2_ERR_ARG, "null argument");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (sim->d_flags.size() >= SB2_MAX_GEOS) return fail(sim, SB2_ERR_LIMIT, "more than SB2_MAX_GEOS geometries host spatial species");
  cudaSetDevice(sim->g.device);
  uint8_t* d = NULL;
  CK(cudaMalloc((void**)&d, sim->nalloc));
  sim->d_flags.push_back(d);
  sim->h_flags.push_back(std::vector<uint8_t>(cell_flags, cell_flags + ncells(sim)));
  int rc = upload_padded<uint8_t>(sim, d, cell_flags);
  if (rc) return rc;
  CK(cudaStreamSynchronize(sim->stream));
  return (int)sim->d_flags.size() - 1;
}

int sb2_add_species(sb2_sim* sim, int geo, const double* init) {
  if (!sim || !init) return fail(sim, SB2_ERR_ARG, "null argument");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (geo < 0 || geo >= (int)sim->d_flags.size()) return fail(sim, SB2_ERR_ARG, "unknown geometry id");
  if (sim->species.size() >= SB2_MAX_SPECIES) return fail(sim, SB2_ERR_LIMIT, "more than SB2_MAX_SPECIES staged species");
  cudaSetDevice(sim->g.device);
  Sb2Species s;
  memset(&s, 0, sizeof(s));
  s.geo = geo;
  int rc;
  for (int b = 0; b < 2; ++b) if ((rc = alloc_field(sim, &s.u[b]))) return rc;
  for (int m = 0; m < 4; ++m) if ((rc = alloc_field(sim, &s.k[m]))) return rc;
  if ((rc = upload_interior<double>(sim, s.u[0], init))) return rc;
  CK(cudaStreamSynchronize(sim->stream));
  sim->species.push_back(s);
  return (int)sim->species.size() - 1;
}

int sb2_add_field(sb2_sim* sim, const double* values) {
  if (!sim || !values) return fail(sim, SB2_ERR_ARG, "null argument");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (sim->d_fields.size() >= SB2_MAX_FIELDS) return fail(sim, SB2_ERR_LIMIT, "more than SB2_MAX_FIELDS read-only fields");
  cudaSetDevice(sim->g.device);
  double* d = NULL;
  int rc = alloc_field(sim, &d);
  if (rc) return rc;
  if ((rc = upload_interior<double>(sim, d, values))) return rc;
  CK(cudaStreamSynchronize(sim->stream));
  sim->d_fields.push_back(d);
  return (int)sim->d_fields.size() - 1;
}

int sb2_set_constants(sb2_sim* sim, const double* table, int n) {
  if (!sim || (!table && n > 0) || n < 0) return fail(sim, SB2_ERR_ARG, "bad constants table");
  ++sim->version;
  if (n > SB2_MAX_CONSTS - sim->n_internal_consts) return fail(sim, SB2_ERR_LIMIT, "constants table full");
  memcpy(sim->consts, table, sizeof(double) * n);
  if (n > sim->n_user_consts) sim->n_user_consts = n;
  return SB2_OK;
}

int sb2_set_constant(sb2_sim* sim, int slot, double value) {
  if (!sim) return SB2_ERR_ARG;
  ++sim->version;
  if (slot < 0 || slot >= SB2_MAX_CONSTS - sim->n_internal_consts) return fail(sim, SB2_ERR_ARG, "constant slot out of range");
  sim->consts[slot] = value;
  if (slot >= sim->n_user_consts) sim->n_user_consts = slot + 1;
  return SB2_OK;
}

int sb2_add_reaction(sb2_sim* sim, int geo, int kind, const sb2_token* toks, int ntoks, const sb2_ref* refs,
                     int nrefs, int n_reactants) {
  if (!sim || (!toks && ntoks > 0) || (!refs && nrefs > 0)) return fail(sim, SB2_ERR_ARG, "null argument");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (geo < 0 || geo >= (int)sim->d_flags.size()) return fail(sim, SB2_ERR_ARG, "unknown geometry id");
  Sb2Stmt st;
  st.geo = geo; st.kind = kind; st.n_reactants = n_reactants;
  st.toks.assign(toks, toks + ntoks);
  st.refs.assign(refs, refs + nrefs);
  sim->stmts.push_back(st);
  return SB2_OK;
}

int sb2_set_diffusion(sb2_sim* sim, int species, int axis, int src_kind, int src) {
  if (!sim || species < 0 || species >= (int)sim->species.size() || axis < 0 || axis > 2)
    return fail(sim, SB2_ERR_ARG, "bad species / axis");
  Sb2Species& s = sim->species[species];
  s.dkind[axis] = src_kind; s.dsrc[axis] = src;
  s.has_diff = s.dkind[0] || s.dkind[1] || s.dkind[2];
  return SB2_OK;
}

int sb2_set_advection(sb2_sim* sim, int species, int axis, int src_kind, int src) {
  if (!sim || species < 0 || species >= (int)sim->species.size() || axis < 0 || axis > 2)
    return fail(sim, SB2_ERR_ARG, "bad species / axis");
  Sb2Species& s = sim->species[species];
  s.akind[axis] = src_kind; s.asrc[axis] = src;
  s.has_adv = true;  // adCInfo array exists as soon as one axis is declared (setInfoFunction.cpp:240-244)
  return SB2_OK;
}

int sb2_set_advection_faces(sb2_sim* sim, int species, int axis, int field) {
  if (!sim || species < 0 || species >= (int)sim->species.size() || axis < 0 || axis > 2)
    return fail(sim, SB2_ERR_ARG, "bad species / axis");
  if (field < 0 || field >= (int)sim->d_fields.size()) return fail(sim, SB2_ERR_ARG, "unknown field");
  sim->species[species].afacesrc[axis] = field + 1;  // stored + 1: the species record is zero-initialised
  return SB2_OK;
}

int sb2_set_boundary(sb2_sim* sim, int species, int side, int bc_kind, int src_kind, int src) {
  if (!sim || species < 0 || species >= (int)sim->species.size() || side < 0 || side > 5)
    return fail(sim, SB2_ERR_ARG, "bad species / side");
  Sb2Species& s = sim->species[species];
  s.bckind[side] = bc_kind; s.bcskind[side] = src_kind; s.bcsrc[side] = src;
  s.has_bc = true;
  return SB2_OK;
}

int sb2_add_mem_transport(sb2_sim* sim, const sb2_token* toks, int ntoks, const sb2_ref* refs, int nrefs,
                          int n_reactants, const sb2_mem_point* pts, int npts, const int32_t* operand_cells,
                          const double* operand_values, const int32_t* target_cells) {
  if (!sim) return SB2_ERR_ARG;
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  cudaSetDevice(sim->g.device);
  Sb2MemTransport t;
  std::string err;
  int rc = sb2_build_transport(sim->g, sim->P, sim->PL, sim->first, (int)sim->species.size(), toks, ntoks, refs,
                               nrefs, n_reactants, pts, npts, operand_cells, operand_values, target_cells, sim->stream, t, err);
  if (rc) return fail(sim, rc, err);
  for (int i = 0; i < ntoks; ++i)
    if (toks[i].kind == SB2_TOK_MEMVAR && toks[i].slot >= sim->mem_species.size()) { sb2_free_transport(t); return fail(sim, SB2_ERR_ARG, "membrane transport: token references unknown membrane species"); }
  for (int r = 0; r < nrefs; ++r)
    if (refs[r].species >= SB2_MEM_SPECIES && refs[r].species - SB2_MEM_SPECIES >= (int)sim->mem_species.size()) { sb2_free_transport(t); return fail(sim, SB2_ERR_ARG, "membrane transport: reference to unknown membrane species"); }
  sim->transports.push_back(t);
  sb2_sim::MemOp op = {0, (int)sim->transports.size() - 1};
  sim->mem_ops.push_back(op);
  return SB2_OK;
}

int sb2_add_point_set(sb2_sim* sim, int npts) {
  if (!sim || npts < 0) return fail(sim, SB2_ERR_ARG, "bad point set");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (sim->g.own_hi > sim->g.own_lo) return fail(sim, SB2_ERR_STATE, "membrane species are not supported on a slab of a sharded grid yet");
  Sb2PointSet ps;
  memset(&ps, 0, sizeof(ps));
  ps.npts = npts;
  sim->point_sets.push_back(ps);
  return (int)sim->point_sets.size() - 1;
}

int sb2_add_mem_species(sb2_sim* sim, int set, const double* init) {
  if (!sim || set < 0 || set >= (int)sim->point_sets.size()) return fail(sim, SB2_ERR_ARG, "unknown point set");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (sim->mem_species.size() >= SB2_MAX_MEM_SPECIES) return fail(sim, SB2_ERR_LIMIT, "more than SB2_MAX_MEM_SPECIES membrane species");
  cudaSetDevice(sim->g.device);
  const int n = sim->point_sets[set].npts;
  if (n > 0 && !init) return fail(sim, SB2_ERR_ARG, "null argument");
  Sb2MemSpecies ms;
  memset(&ms, 0, sizeof(ms));
  ms.set = set;
  const size_t bytes = sizeof(double) * (size_t)(n > 0 ? n : 1);
  CK(cudaMalloc((void**)&ms.u, bytes));
  CK(cudaMemsetAsync(ms.u, 0, bytes, sim->stream));
  for (int m = 0; m < 4; ++m) {
    CK(cudaMalloc((void**)&ms.k[m], bytes));
    CK(cudaMemsetAsync(ms.k[m], 0, bytes, sim->stream));
  }
  if (n > 0) CK(cudaMemcpyAsync(ms.u, init, sizeof(double) * n, cudaMemcpyHostToDevice, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  sim->mem_species.push_back(ms);
  return (int)sim->mem_species.size() - 1;
}

int sb2_add_mem_reaction(sb2_sim* sim, int set, int kind, const sb2_token* toks, int ntoks, const sb2_ref* refs, int nrefs,
                         int n_reactants, const int32_t* points, int npts, const double* operand_values) {
  if (!sim || set < 0 || set >= (int)sim->point_sets.size()) return fail(sim, SB2_ERR_ARG, "unknown point set");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  cudaSetDevice(sim->g.device);
  Sb2MemReaction r;
  std::string err;
  int rc = sb2_build_mem_reaction(set, kind, toks, ntoks, refs, nrefs, n_reactants, points, npts, operand_values,
                                  (int)sim->mem_species.size(), sim->point_sets[set].npts, sim->stream, r, err);
  if (rc) return fail(sim, rc, err);
  for (int i = 0; i < r.nrefs; ++i)
    if (sim->mem_species[r.ref_species[i]].set != set) { sb2_free_mem_reaction(r); return fail(sim, SB2_ERR_ARG, "membrane reaction: species of another point set"); }
  sim->mem_reactions.push_back(r);
  sb2_sim::MemOp op = {1, (int)sim->mem_reactions.size() - 1};
  sim->mem_ops.push_back(op);
  return SB2_OK;
}

int sb2_set_mem_diffusion(sb2_sim* sim, int mem_species, int src_kind, int src, const double* d_values, const int32_t* points, int npts,
                          const int32_t* term_start, const int32_t* adj, const double* s_ij, const double* d_ij, const double* area) {
  if (!sim || mem_species < 0 || mem_species >= (int)sim->mem_species.size()) return fail(sim, SB2_ERR_ARG, "unknown membrane species");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (npts < 0 || (npts > 0 && (!points || !term_start || !area))) return fail(sim, SB2_ERR_ARG, "null argument");
  if (src_kind != SB2_SRC_CONST && src_kind != SB2_SRC_FIELD) return fail(sim, SB2_ERR_ARG, "membrane diffusion coefficient: bad source kind");
  if (src_kind == SB2_SRC_CONST && (src < 0 || src >= SB2_MAX_CONSTS)) return fail(sim, SB2_ERR_ARG, "membrane diffusion coefficient: constant slot out of range");
  if (src_kind == SB2_SRC_FIELD && npts > 0 && !d_values) return fail(sim, SB2_ERR_ARG, "membrane diffusion coefficient: values missing");
  cudaSetDevice(sim->g.device);
  Sb2MemSpecies& ms = sim->mem_species[mem_species];
  if (ms.has_diff) return fail(sim, SB2_ERR_STATE, "membrane diffusion already set for this species");
  const int set_n = sim->point_sets[ms.set].npts;
  const int nterms = npts > 0 ? term_start[npts] : 0;
  for (int p = 0; p < npts; ++p)
    if (points[p] < 0 || points[p] >= set_n || term_start[p + 1] < term_start[p]) return fail(sim, SB2_ERR_ARG, "membrane diffusion: bad point list");
  for (int t = 0; t < nterms; ++t)
    if (adj[t] < 0 || adj[t] >= set_n) return fail(sim, SB2_ERR_ARG, "membrane diffusion: neighbour outside the point set");
  ms.has_diff = true; ms.dkind = src_kind; ms.dsrc = src; ms.n_diff_pts = npts; ms.n_terms = nterms;
  int rc;
  if ((rc = upload_list(sim, &ms.d_diff_pts, points, (size_t)npts)) || (rc = upload_list(sim, &ms.d_term_start, term_start, (size_t)npts + 1)) ||
      (rc = upload_list(sim, &ms.d_adj, adj, (size_t)nterms)) || (rc = upload_list(sim, &ms.d_s, s_ij, (size_t)nterms)) ||
      (rc = upload_list(sim, &ms.d_d, d_ij, (size_t)nterms)) || (rc = upload_list(sim, &ms.d_area, area, (size_t)npts)))
    return rc;
  if (src_kind == SB2_SRC_FIELD && (rc = upload_list(sim, &ms.d_dvalues, d_values, (size_t)npts))) return rc;
  return SB2_OK;
}

int sb2_set_mem_update_points(sb2_sim* sim, int mem_species, const int32_t* points, int npts) {
  if (!sim || mem_species < 0 || mem_species >= (int)sim->mem_species.size()) return fail(sim, SB2_ERR_ARG, "unknown membrane species");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (npts < 0 || (npts > 0 && !points)) return fail(sim, SB2_ERR_ARG, "null argument");
  cudaSetDevice(sim->g.device);
  Sb2MemSpecies& ms = sim->mem_species[mem_species];
  const int set_n = sim->point_sets[ms.set].npts;
  for (int p = 0; p < npts; ++p)
    if (points[p] < 0 || points[p] >= set_n) return fail(sim, SB2_ERR_ARG, "membrane update list: point outside the point set");
  if (ms.d_update_pts) { cudaFree(ms.d_update_pts); ms.d_update_pts = NULL; }
  ms.n_update = npts;
  return upload_list(sim, &ms.d_update_pts, points, (size_t)npts);
}

int sb2_set_mem_fill(sb2_sim* sim, int set, const int32_t* target, const int32_t* a, const int32_t* b, int n) {
  if (!sim || set < 0 || set >= (int)sim->point_sets.size()) return fail(sim, SB2_ERR_ARG, "unknown point set");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (n < 0 || (n > 0 && (!target || !a || !b))) return fail(sim, SB2_ERR_ARG, "null argument");
  cudaSetDevice(sim->g.device);
  Sb2PointSet& ps = sim->point_sets[set];
  for (int i = 0; i < n; ++i)
    if (target[i] < 0 || target[i] >= ps.npts || a[i] < 0 || a[i] >= ps.npts || b[i] < 0 || b[i] >= ps.npts)
      return fail(sim, SB2_ERR_ARG, "pseudo-membrane fill: point outside the point set");
  ps.n_fill = n;
  int rc;
  if ((rc = upload_list(sim, &ps.d_fill_target, target, (size_t)n)) || (rc = upload_list(sim, &ps.d_fill_a, a, (size_t)n)) ||
      (rc = upload_list(sim, &ps.d_fill_b, b, (size_t)n)))
    return rc;
  return SB2_OK;
}

int sb2_add_assignment_rule(sb2_sim* sim, int target_kind, int target, int geo, const sb2_token* toks, int ntoks) {
  if (!sim || (!toks && ntoks > 0) || ntoks < 0) return fail(sim, SB2_ERR_ARG, "null argument");
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  if (target_kind == 0 && (target < 0 || target >= (int)sim->d_fields.size())) return fail(sim, SB2_ERR_ARG, "assignment rule: unknown field");
  if (target_kind == 1 && (target < 0 || target >= (int)sim->species.size())) return fail(sim, SB2_ERR_ARG, "assignment rule: unknown species");
  if (target_kind != 0 && target_kind != 1) return fail(sim, SB2_ERR_ARG, "assignment rule: bad target kind");
  if (geo >= (int)sim->d_flags.size()) return fail(sim, SB2_ERR_ARG, "assignment rule: unknown geometry");
  for (int i = 0; i < ntoks; ++i) {
    if (toks[i].kind == SB2_TOK_VAR && toks[i].slot >= sim->species.size()) return fail(sim, SB2_ERR_ARG, "assignment rule: token references unknown species");
    if (toks[i].kind == SB2_TOK_FIELD && toks[i].slot >= sim->d_fields.size()) return fail(sim, SB2_ERR_ARG, "assignment rule: token references unknown field");
    if (toks[i].kind == SB2_TOK_CONST && toks[i].slot >= SB2_MAX_CONSTS) return fail(sim, SB2_ERR_ARG, "assignment rule: constant slot out of range");
    if (toks[i].kind == SB2_TOK_MEMVAR) return fail(sim, SB2_ERR_ARG, "assignment rule: membrane species operands are not supported");
  }
  cudaSetDevice(sim->g.device);
  Sb2CellRule r;
  std::string err;
  int rc = sb2_build_cell_rule(toks, ntoks, target_kind, target, geo, sim->stream, r, err);
  if (rc) return fail(sim, rc, err);
  sim->cell_rules.push_back(r);
  return SB2_OK;
}

int sb2_download_mem_species(sb2_sim* sim, int mem_species, double* out) {
  if (!sim || !out || mem_species < 0 || mem_species >= (int)sim->mem_species.size()) return fail(sim, SB2_ERR_ARG, "bad membrane species");
  cudaSetDevice(sim->g.device);
  const Sb2MemSpecies& ms = sim->mem_species[mem_species];
  const int n = sim->point_sets[ms.set].npts;
  if (n > 0) CK(cudaMemcpyAsync(out, ms.u, sizeof(double) * n, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return SB2_OK;
}

int sb2_upload_mem_species(sb2_sim* sim, int mem_species, const double* in) {
  if (!sim || !in || mem_species < 0 || mem_species >= (int)sim->mem_species.size()) return fail(sim, SB2_ERR_ARG, "bad membrane species");
  cudaSetDevice(sim->g.device);
  const Sb2MemSpecies& ms = sim->mem_species[mem_species];
  const int n = sim->point_sets[ms.set].npts;
  if (n > 0) CK(cudaMemcpyAsync(ms.u, in, sizeof(double) * n, cudaMemcpyHostToDevice, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return SB2_OK;
}

int sb2_download_mem_stage(sb2_sim* sim, int mem_species, int m, double* out) {
  if (!sim || !out || mem_species < 0 || mem_species >= (int)sim->mem_species.size() || m < 0 || m > 3) return fail(sim, SB2_ERR_ARG, "bad membrane species / stage");
  cudaSetDevice(sim->g.device);
  const Sb2MemSpecies& ms = sim->mem_species[mem_species];
  const int n = sim->point_sets[ms.set].npts;
  if (n > 0) CK(cudaMemcpyAsync(out, ms.k[m], sizeof(double) * n, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return SB2_OK;
}

int sb2_finalize(sb2_sim* sim) {
  if (!sim) return SB2_ERR_ARG;
  if (sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator already finalized");
  cudaSetDevice(sim->g.device);
  sim->program.clear();
  for (size_t i = 0; i < sim->stmts.size(); ++i) {
    int rc = compile_stmt(sim, sim->stmts[i], sim->program);
    if (rc) return rc;
  }
  if (sim->program.size() > SB2_MAX_TOKENS) return fail(sim, SB2_ERR_LIMIT, "device program longer than SB2_MAX_TOKENS");
  sim->any_adv = false;
  for (size_t s = 0; s < sim->species.size(); ++s) if (sim->species[s].has_adv) sim->any_adv = true;

  Sb2StageCore& b = sim->base;
  memset(&b, 0, sizeof(b));
  b.nx = sim->g.nx; b.ny = sim->g.ny; b.nz = sim->g.nz; b.dim = sim->g.dim;
  b.P = sim->P; b.PL = sim->PL; b.first = sim->first;
  b.S = (int)sim->species.size();
  b.G = (int)sim->d_flags.size();
  const double del[3] = {sim->g.dx, sim->g.dy, sim->g.dz};
  for (int a = 0; a < 3; ++a) {
    b.dx2[a] = pow(del[a], 2);  // the reference divides by pow(delta, 2)
    b.divide[a] = (b.dx2[a] != 1.0) ? 1 : 0;
  }
  for (int s = 0; s < b.S; ++s) {
    Sb2Species& sp = sim->species[s];
    b.sp[s].geo = sp.geo;
    b.sp[s].has_diff = sp.has_diff ? 1 : 0;
    b.sp[s].fast2d = (sp.has_diff && sim->g.dim == 2 && sp.dkind[0] == SB2_SRC_CONST && sp.dkind[1] == SB2_SRC_CONST &&
                      !b.divide[0] && !b.divide[1]) ? 1 : 0;
    b.sp[s].fastpath = sp.has_diff ? 1 : 0;
    for (int a = 0; a < sim->g.dim; ++a)
      if (sp.dkind[a] != SB2_SRC_CONST || b.divide[a]) b.sp[s].fastpath = 0;
    for (int a = 0; a < 3; ++a) {
      b.sp[s].dkind[a] = sp.dkind[a]; b.sp[s].dsrc[a] = sp.dsrc[a];
      if (sp.dkind[a] == SB2_SRC_FIELD && (sp.dsrc[a] < 0 || sp.dsrc[a] >= (int)sim->d_fields.size()))
        return fail(sim, SB2_ERR_ARG, "diffusion coefficient references unknown field");
      if (sp.dkind[a] == SB2_SRC_CONST && (sp.dsrc[a] < 0 || sp.dsrc[a] >= SB2_MAX_CONSTS))
        return fail(sim, SB2_ERR_ARG, "diffusion coefficient references constant slot out of range");
    }
  }
  for (int g = 0; g < b.G; ++g) b.flags[g] = sim->d_flags[g];
  for (size_t f = 0; f < sim->d_fields.size(); ++f) b.fields[f] = sim->d_fields[f];
  b.ntok = (int)sim->program.size();
  b.depth = 1;
  for (size_t i = 0; i < sim->program.size(); ++i) {
    const uint32_t code = sim->program[i] & 0xffffu;
    const int opc = (int)(code / SB2_MAX_DEPTH), d = (int)(code % SB2_MAX_DEPTH);
    if (opc < OPC_MASK && d + 1 + (opc == OPC_MOVUP ? 1 : 0) > b.depth) b.depth = d + 1 + (opc == OPC_MOVUP ? 1 : 0);
  }
  memset(sim->base_tok, 0, sizeof(sim->base_tok));
  memcpy(sim->base_tok, sim->program.data(), sizeof(uint32_t) * sim->program.size());
  sim->base_tok[sim->program.size()] = enc(OPC_END, 0, 0);
  sim->base_tok[sim->program.size() + 1] = enc(OPC_END, 0, 0);

  // 2-D models with up to four staged species run the rd_stage kernel; SB2_STAGE_KERNEL=gen1 forces the
  // first-generation kernel (kept for 3-D grids and more species, and as a cross-check in the tests)
  const char* force = getenv("SB2_STAGE_KERNEL");
  sim->rd_id = (force && !strcmp(force, "gen1")) ? -1 : sb2_rd_pick_variant(sim->g.dim, sim->g.nx, b.S, b.depth, &sim->rd, ncells(sim));
  if (sim->rd_id >= 0) {
    if (const char* ge = getenv("SB2_RD_GEOM")) {  // experiments with the specialised kernel only: "np,w" (no prebuilt kernel has this shape)
      int np = 0, w = 0;
      if (sscanf(ge, "%d,%d", &np, &w) == 2 && np >= 1 && np <= 4 && w >= 2 && w <= 16) { sim->rd.np = np; sim->rd.w = w; }
    }
    rd_lower(sim->program, sim->rdprog, sim->rdslot);
    if (sim->rdprog.size() > SB2_RD_MAX_TOKENS) sim->rd_id = -1;  // very long programs stay on the first-generation kernel
  }
  b.rd_variant = sim->rd_id;
  if (sim->rd_id >= 0) {
    b.nrd = (int)sim->rdprog.size();
    CK(cudaMalloc((void**)&sim->d_rdtok, sizeof(Sb2RdTok) * (SB2_RD_MAX_TOKENS + 2) * 2));
    sim->rd_uploaded.clear();
    // class map: one byte per strip row in 2-D, one per 64 cells of a row in 3-D (the z-marching kernel combines them)
    const int sw = sim->g.dim == 3 ? 64 : 64 * sim->rd.np;
    if (64 * sim->rd.np + 4 > SB2_TAIL_PAD) return fail(sim, SB2_ERR_STATE, "strip wider than the tail pad of the field layout");
    sim->nstrips = (sim->g.nx + sw - 1) / sw;
    const int nrows = sim->g.ny * sim->g.nz;  // every row of every held plane
    CK(cudaMalloc((void**)&sim->d_cls, (size_t)sim->nstrips * nrows));
    const uint8_t* fl[SB2_MAX_GEOS] = {0, 0, 0, 0};
    for (int g = 0; g < b.G; ++g) fl[g] = sim->d_flags[g];
    cudaError_t ce = sb2_rd_classify(fl, b.G, sim->g.nx, sim->g.ny, sim->g.nz, sim->P, sim->PL, sim->first, sw, sim->d_cls, sim->nstrips, sim->stream);
    if (ce != cudaSuccess) return cuda_fail(sim, ce, "strip classification kernel");
    if (getenv("SB2_RD_ALL_GENERAL")) CK(cudaMemsetAsync(sim->d_cls, 0, (size_t)sim->nstrips * nrows, sim->stream));  // experiments: per-cell flag path everywhere
    b.cls = sim->d_cls;
    b.nstrips = sim->nstrips;
    if (getenv("SB2_DEBUG")) {
      CK(cudaMalloc((void**)&sim->d_dbg, 64 * sizeof(uint32_t)));
      CK(cudaMemsetAsync(sim->d_dbg, 0, 64 * sizeof(uint32_t), sim->stream));
      b.dbg = sim->d_dbg;
    }
    // model-specialised build (include/spatial_b200.h, sb2_specialized)
    {
      const char* jenv = getenv("SB2_JIT");
      const bool require = jenv && !strcmp(jenv, "require");
      if (getenv("SB2_DEBUG")) { sim->jit_mode = 0; sim->jit_info = "interpreter kernel: SB2_DEBUG is set (the specialised build carries no debug words)"; }
      else if (jenv && !strcmp(jenv, "0")) { sim->jit_mode = 0; sim->jit_info = "interpreter kernel: SB2_JIT=0"; }
      else if (jenv && !strcmp(jenv, "background")) sim->jit_mode = 2;
      else if (jenv) sim->jit_mode = 1;
      else sim->jit_mode = ncells(sim) >= (1 << 20) ? 1 : 2;
      if (sim->jit_mode == 1) {
        specialise(sim, false, true);
        if (!sim->jit && require) return fail(sim, SB2_ERR_STATE, sim->jit_info);
      } else if (sim->jit_mode == 2) {
        if (jenv) specialise(sim, true, false);  // SB2_JIT=background: start right away
        else sim->jit_info = "interpreter kernel: the specialised build starts in the background after 256 steps";
      }
      if (sim->trace) fprintf(stderr, "[sb2] %s\n", sim->jit_info.c_str());
    }
    if (sim->trace) fprintf(stderr, "[sb2] rd_stage variant %d (np %d, w %d, depth %d), %d strips, program depth %d, %d tokens\n",
                            sim->rd_id, sim->rd.np, sim->rd.w, sim->rd.depth, sim->nstrips, b.depth, b.ntok);
  }

  // boundary-condition lists (Neumann flux / Dirichlet overwrite), built from the host flag copies
  std::string err;
  std::vector<Sb2BcSpecies> bcs(sim->species.size());
  for (size_t s = 0; s < sim->species.size(); ++s) {
    Sb2Species& sp = sim->species[s];
    bcs[s].geo = sp.geo; bcs[s].has_bc = sp.has_bc;
    for (int k = 0; k < 6; ++k) { bcs[s].kind[k] = sp.bckind[k]; bcs[s].skind[k] = sp.bcskind[k]; bcs[s].src[k] = sp.bcsrc[k]; }
  }
  // slab: a model with membrane transport reads cells two rows away from a membrane point, one with advection needs two rows
  // of u and of the face values around every owned row
  sim->model_dirichlet = false;
  for (size_t s = 0; s < sim->species.size(); ++s)
    for (int k = 0; k < 6; ++k) if (sim->species[s].has_bc && sim->species[s].bckind[k] == SB2_BC_DIRICHLET) sim->model_dirichlet = true;
  sim->edge_w = sim->transports.empty() ? 1 : 2;
  sim->end_w = (sim->any_adv || !sim->transports.empty()) ? 2 : 1;
  if (!sim->transports.empty()) {
    // transport buffers of the volume species, and per transport which of its (species, cell) groups no earlier transport feeds
    for (size_t s = 0; s < sim->species.size(); ++s)
      if (!sim->species[s].tcarry) { int rc0 = alloc_field(sim, &sim->species[s].tcarry); if (rc0) return rc0; }
    std::set<std::pair<int, int64_t> > seen;
    for (size_t i = 0; i < sim->mem_ops.size(); ++i) {
      if (sim->mem_ops[i].type != 0) continue;
      Sb2MemTransport& t = sim->transports[sim->mem_ops[i].index];
      std::vector<uint8_t> first(t.h_group_species.size(), 0);
      for (size_t g = 0; g < first.size(); ++g) {
        if (t.h_group_species[g] >= SB2_MAX_SPECIES) continue;  // membrane species accumulate in their own k arrays
        first[g] = seen.insert(std::make_pair((int)t.h_group_species[g], t.h_group_cell[g])).second ? 1 : 0;
      }
      cudaError_t e = sb2_transport_set_first(t, first, sim->stream);
      if (e != cudaSuccess) return cuda_fail(sim, e, "membrane transport lists");
    }
  }
  int rc = sb2_build_boundary_lists(sim->g, sim->P, sim->PL, sim->first, sim->h_flags, bcs, sim->consts, sim->stream,
                                    sim->edge_w, sim->blists, err);
  if (rc) return fail(sim, rc, err);

  // advected species carry CIP face values (allocated by sb2_upload_faces or here, zero-filled)
  for (size_t s = 0; s < sim->species.size(); ++s) {
    Sb2Species& sp = sim->species[s];
    if (!sp.has_adv) continue;
    for (int a = 0; a < sim->g.dim; ++a) {
      if (sp.akind[a] == SB2_SRC_FIELD && (sp.asrc[a] < 0 || sp.asrc[a] >= (int)sim->d_fields.size() || sp.afacesrc[a] <= 0))
        return fail(sim, SB2_ERR_ARG, "field-valued advection velocity: cell field and face field (sb2_set_advection_faces) are both needed");
      if (!sp.faces[a]) { rc = alloc_field(sim, &sp.faces[a]); if (rc) return rc; }
      if (sim->g.dim == 2 && !sp.faces_alt[a] && !getenv("SB2_CIP_UNFUSED")) { rc = alloc_field(sim, &sp.faces_alt[a]); if (rc) return rc; }
    }
  }
  if (sim->any_adv) { rc = alloc_field(sim, &sim->d_tmp_face); if (rc) return rc; }
  CK(cudaMalloc((void**)&sim->d_scratch, sizeof(double) * 4096));
  CK(cudaStreamSynchronize(sim->stream));
  sim->finalized = true;
  return SB2_OK;
}

int sb2_begin_step(sb2_sim* sim, double t_arg, double dt, int time_slot) {
  if (!sim || !sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator not finalized");
  cudaSetDevice(sim->g.device);
  if (time_slot >= 0) {
    if (time_slot >= SB2_MAX_CONSTS) return fail(sim, SB2_ERR_ARG, "time slot out of range");
    sim->consts[time_slot] = t_arg * dt;  // sim_time = t * dt, spatialsimulator.cpp:1358
  }
  sim->cur_dt = dt;
  // tiered specialisation: a simulator that keeps stepping gets its own stage kernel, built in the background and
  // picked up here, at a step boundary (never inside a stream capture)
  ++sim->steps_seen;
  if (sim->jit_mode == 2 && !sim->jit && !sim->capturing && (sim->jit_job || sim->steps_seen > 256)) specialise(sim, true, false);
  // A Neumann flux that is exactly zero adds +0.0 everywhere (all shipped models): the boundary
  // passes are skipped and the combine can be fused.  A non-zero flux takes the un-fused schedule,
  // which reproduces the reference's order of additions including the flux carried in k0.
  sim->carry_k0 = sb2_neumann_active(sim->blists[0], (int)sim->species.size(), sim->consts, sim->g.own_hi <= sim->g.own_lo);
  sim->fuse_combine = !sim->any_adv && !sim->carry_k0;
  // on a slab a Dirichlet value imposed after the update has to reach the neighbours' halo rows: the new u is exchanged
  // after sb2_end_step instead of right after the last stage
  if (sim->g.own_hi > sim->g.own_lo && sim->model_dirichlet) sim->fuse_combine = false;
  if (sim->carry_k0 && !sim->k0_carry_valid) {
    for (size_t s = 0; s < sim->species.size(); ++s)
      CK(cudaMemsetAsync(sim->species[s].k[0], 0, sizeof(double) * sim->nalloc, sim->stream));
  }
  sim->in_step = true;
  return SB2_OK;
}

namespace {
// rows / planes of the slab axis: what is held, what is owned, and how many owned rows on either side face a neighbour
struct SlabRows { int n_held, own_lo, own_hi, lo_edge, hi_edge; };
SlabRows slab_rows(const sb2_sim* sim) {
  SlabRows r;
  r.n_held = sim->g.dim == 3 ? sim->g.nz : sim->g.ny;
  const bool whole = sim->g.own_hi <= sim->g.own_lo;
  r.own_lo = whole ? 0 : sim->g.own_lo;
  r.own_hi = whole ? r.n_held : sim->g.own_hi;
  // owned rows next to a neighbouring slab (their results are what the neighbours' halos need): edge_w on either side
  r.lo_edge = r.own_lo > 0 ? sim->edge_w : 0;
  r.hi_edge = r.own_hi < r.n_held ? sim->edge_w : 0;
  if (r.lo_edge + r.hi_edge > r.own_hi - r.own_lo) {  // a thin slab is all edge
    r.lo_edge = r.own_lo > 0 ? r.own_hi - r.own_lo : 0;
    r.hi_edge = r.lo_edge ? 0 : r.own_hi - r.own_lo;
  }
  return r;
}

// membrane transport and the work on membrane species of one stage (everything is O(perimeter))
int launch_membrane_ops(sb2_sim* sim, int m) {
  static const double rk[4] = {0.0, 0.5, 0.5, 1.0};
  const double dt = sim->cur_dt;
  // membrane transport first: its flux is the carry-in of the stage accumulators of the volume species
  const bool in_buffer = transport_in_buffer(sim, m);  // else: k0 already holds the carried Neumann flux, the transports add to it
  // transports and membrane reactions in the order of the reference's reaction list (what a membrane species receives
  // adds up in that order), then surface diffusion (spatialsimulator.cpp:1369-1479)
  if (sim->mem_ops.empty() && sim->mem_species.empty()) return SB2_OK;
  const Sb2MemView mv = mem_view(sim);
  const double* mu[SB2_MAX_MEM_SPECIES]; const double* mkp[SB2_MAX_MEM_SPECIES]; double* mko[SB2_MAX_MEM_SPECIES];
  for (int i = 0; i < SB2_MAX_MEM_SPECIES; ++i) { mu[i] = mv.u[i]; mkp[i] = m > 0 ? mv.k[i][m - 1] : NULL; mko[i] = mv.k[i][m]; }
  std::vector<const double*> u(sim->species.size()), kp(sim->species.size());
  std::vector<double*> ko(sim->species.size());
  for (size_t s = 0; s < sim->species.size(); ++s) {
    u[s] = sim->species[s].u[sim->species[s].cur];
    kp[s] = m > 0 ? sim->species[s].k[m - 1] : NULL;
    ko[s] = in_buffer ? sim->species[s].tcarry : sim->species[s].k[m];
  }
  for (size_t i = 0; i < sim->mem_ops.size(); ++i) {
    cudaError_t e;
    if (sim->mem_ops[i].type == 0)
      e = sb2_launch_transport(sim->transports[sim->mem_ops[i].index], sim->g, u.data(), kp.data(), ko.data(), mu, mkp, mko, sim->consts, m,
                               rk[m] * dt, in_buffer, sim->stream, &sim->launches);
    else
      e = sb2_launch_mem_reaction(sim->mem_reactions[sim->mem_ops[i].index], mv, sim->consts, m, rk[m] * dt, sim->stream, &sim->launches);
    if (e != cudaSuccess) return cuda_fail(sim, e, "membrane transport / reaction kernels");
  }
  for (size_t i = 0; i < sim->mem_species.size(); ++i) {
    cudaError_t e = sb2_launch_mem_diffusion(sim->mem_species[i], (int)i, mv, sim->consts, m, rk[m] * dt, sim->g.dim, sim->stream, &sim->launches);
    if (e != cudaSuccess) return cuda_fail(sim, e, "membrane diffusion kernel");
  }
  return SB2_OK;
}

// calcBoundary of stage m on the rows of one phase (0 all owned rows, 1 edge rows, 2 the others); it follows the diffusion of
// the same stage and the same rows (spatialsimulator.cpp:1476-1478)
int launch_boundary(sb2_sim* sim, int m, int phase) {
  if (m == 3 && sim->fuse_combine) return SB2_OK;
  std::vector<double*> ko(sim->species.size()), uc(sim->species.size());
  for (size_t s = 0; s < sim->species.size(); ++s) { ko[s] = sim->species[s].k[m]; uc[s] = sim->species[s].u[sim->species[s].cur]; }
  // A Dirichlet value rewrites u itself (calcPDE.cpp:995), after every cell has read u for this stage.  The edge rows of a slab
  // run beside its other rows, which read them as neighbours: their Dirichlet values are written with the second phase, once
  // both kernels of the stage are done, and travel to the neighbouring slabs after that.
  cudaError_t e = sb2_apply_boundary(sim->blists[phase], ko.data(), uc.data(), sim->consts, sim->d_fields.data(), sim->carry_k0,
                                     phase != 1, sim->stream, &sim->launches);
  if (e == cudaSuccess && phase == 2)
    e = sb2_apply_boundary(sim->blists[1], ko.data(), uc.data(), sim->consts, sim->d_fields.data(), false, true, sim->stream, &sim->launches);
  if (e != cudaSuccess) return cuda_fail(sim, e, "boundary kernels");
  return SB2_OK;
}
}  // namespace

int sb2_stage(sb2_sim* sim, int m, int phase) {
  if (!sim || !sim->in_step) return fail(sim, SB2_ERR_STATE, "sb2_stage outside sb2_begin_step / sb2_end_step");
  if (m < 0 || m > 3) return fail(sim, SB2_ERR_ARG, "stage index out of range");
  if (phase < 0 || phase > 2) return fail(sim, SB2_ERR_ARG, "phase out of range");
  cudaSetDevice(sim->g.device);
  const SlabRows r = slab_rows(sim);
  const double dt = sim->cur_dt;
  int rc = SB2_OK;
  if (phase == 0 || phase == 1) {
    if ((rc = launch_membrane_ops(sim, m))) return rc;
  }
  auto launch = [&](int rb, int re) -> int { return launch_stage_rows(sim, m, dt, rb, re); };
  if (phase == 0) rc = launch(r.own_lo, r.own_hi);
  else if (phase == 1) {
    if (r.lo_edge) rc = launch(r.own_lo, r.own_lo + r.lo_edge);
    if (!rc && r.hi_edge) rc = launch(r.own_hi - r.hi_edge, r.own_hi);
  } else {
    const int b = r.own_lo + r.lo_edge, e = r.own_hi - r.hi_edge;
    if (e > b) rc = launch(b, e);  // (a thin slab is fully covered by phase 1)
  }
  if (rc) return rc;
  return launch_boundary(sim, m, phase);
}

// sb2_stage on another stream (cudaStream_t as void*): lets the caller run the edge rows and the interior rows of a
// stage concurrently.  Ordering between the streams is the caller's business (events).
int sb2_stage_on(sb2_sim* sim, int m, int phase, void* cuda_stream) {
  if (!sim) return SB2_ERR_ARG;
  cudaStream_t keep = sim->stream;
  sim->multi_stream = true;
  sim->stream = (cudaStream_t)cuda_stream;
  const int rc = sb2_stage(sim, m, phase);
  sim->stream = keep;
  return rc;
}

namespace {
// the CIP-CSLR passes of one step, flattened: (species, axis) for every advected species and every axis it has a velocity on
struct AdvPass { int species, axis; };
std::vector<AdvPass> advection_passes(const sb2_sim* sim) {
  std::vector<AdvPass> out;
  for (size_t s = 0; s < sim->species.size(); ++s) {
    if (!sim->species[s].has_adv) continue;
    for (int a = 0; a < sim->g.dim; ++a)
      if (sim->species[s].akind[a] != SB2_SRC_NONE) { AdvPass p = {(int)s, a}; out.push_back(p); }
  }
  return out;
}

// one direction pass of cipCSLR for one species (calcPDE.cpp:581-673) on the owned rows
int launch_advection_pass(sb2_sim* sim, const AdvPass& p) {
  Sb2Species& sp = sim->species[p.species];
  const SlabRows r = slab_rows(sim);
  Sb2CipArgs c;
  memset(&c, 0, sizeof(c));
  c.nx = sim->g.nx; c.ny = sim->g.ny; c.nz = sim->g.nz; c.dim = sim->g.dim;
  c.P = sim->P; c.PL = sim->PL; c.first = sim->first; c.nalloc = sim->nalloc;
  c.delta[0] = sim->g.dx; c.delta[1] = sim->g.dy; c.delta[2] = sim->g.dz;
  c.dt = sim->cur_dt;
  c.u = sp.u[sp.cur];
  for (int a = 0; a < 3; ++a) {
    c.faces[a] = sp.faces[a];
    c.akind[a] = sp.akind[a];
    c.aconst[a] = sp.akind[a] == SB2_SRC_CONST ? sim->consts[sp.asrc[a]] : 0.0;
    c.afield[a] = sp.akind[a] == SB2_SRC_FIELD ? sim->d_fields[sp.asrc[a]] : NULL;
    c.afield_face[a] = (sp.akind[a] == SB2_SRC_FIELD && sp.afacesrc[a] > 0) ? sim->d_fields[sp.afacesrc[a] - 1] : NULL;
  }
  c.flags = sim->d_flags[sp.geo];
  c.tmp_cell = sp.u[sp.cur ^ 1];  // the idle half of the double buffer is scratch here
  c.tmp_face = sim->d_tmp_face;
  c.row_begin = r.own_lo; c.row_end = r.own_hi;
  // the update sweep moves the faces of the OTHER axes by the mean increment of the two cells they separate: the face between
  // the last owned row and the neighbouring slab needs the increment of that slab's first row, which is computed here as well
  c.flux_row_end = (p.axis != (sim->g.dim == 3 ? 2 : 1) && r.own_hi < r.n_held) ? r.own_hi + 1 : r.own_hi;
  const bool fused = sim->g.dim == 2 && sp.faces_alt[p.axis] != NULL;
  if (fused) { c.u_out = sp.u[sp.cur ^ 1]; c.face_out = sp.faces_alt[p.axis]; }
  cudaError_t e = sb2_launch_cip_pass(c, p.axis, sim->stream, &sim->launches);
  if (e != cudaSuccess) return cuda_fail(sim, e, "CIP-CSLR advection kernels");
  if (fused) {
    // the pass wrote the new cell averages and the new faces of its axis out of place: they are the current ones now.  Rows this
    // handle does not own (halo rows of a slab) are stale in the new halves until the exchange that follows every pass.
    sp.cur ^= 1;
    std::swap(sp.faces[p.axis], sp.faces_alt[p.axis]);
    sp.face_par ^= 1 << p.axis;
  }
  return SB2_OK;
}

// what follows the stages and the advection: updateSpecies for every species (unless fused into the last stage), the
// post-update boundary pass, membrane species, assignment rules, pseudo-membrane fill
int finish_step(sb2_sim* sim) {
  const double dt = sim->cur_dt;
  const size_t S = sim->species.size();
  if (sim->fuse_combine) {
    for (size_t s = 0; s < S; ++s) sim->species[s].cur ^= 1;
  } else {
    Sb2CombineArgs c;
    memset(&c, 0, sizeof(c));
    c.nx = sim->g.nx; c.ny = sim->g.ny; c.nz = sim->g.nz; c.dim = sim->g.dim; c.S = (int)S;
    c.P = sim->P; c.PL = sim->PL; c.first = sim->first; c.dt = dt;
    for (size_t s = 0; s < S; ++s) {
      c.u[s] = sim->species[s].u[sim->species[s].cur];
      for (int m = 0; m < 4; ++m) c.k[s][m] = sim->species[s].k[m];
      c.flags[s] = sim->d_flags[sim->species[s].geo];
    }
    cudaError_t e = sb2_launch_combine(c, sim->stream, &sim->launches);
    if (e != cudaSuccess) return cuda_fail(sim, e, "combine kernel launch");
  }
  // post-update calcBoundary(m = 0) (spatialsimulator.cpp:1540-1542): Dirichlet values are
  // re-imposed, Neumann flux is parked in the zeroed k0 and carried into the next step
  if (sim->blists[0].n_dirichlet > 0 || sim->carry_k0) {  // (no entries: nothing to write, on a slab as well)
    std::vector<double*> k0(S), uc(S);
    for (size_t s = 0; s < S; ++s) { k0[s] = sim->species[s].k[0]; uc[s] = sim->species[s].u[sim->species[s].cur]; }
    if (sim->carry_k0)
      for (size_t s = 0; s < S; ++s) CK(cudaMemsetAsync(k0[s], 0, sizeof(double) * sim->nalloc, sim->stream));
    cudaError_t e = sb2_apply_boundary(sim->blists[0], k0.data(), uc.data(), sim->consts, sim->d_fields.data(), sim->carry_k0, true,
                                       sim->stream, &sim->launches);
    if (e != cudaSuccess) return cuda_fail(sim, e, "boundary kernels");
  }
  // species on membranes: RK4 update on the points of their geometry (updateSpecies, spatialsimulator.cpp:1526-1537)
  if (!sim->mem_species.empty()) {
    const Sb2MemView mv = mem_view(sim);
    for (size_t i = 0; i < sim->mem_species.size(); ++i) {
      cudaError_t e = sb2_launch_mem_update(sim->mem_species[i], (int)i, mv, dt, sim->stream, &sim->launches);
      if (e != cudaSuccess) return cuda_fail(sim, e, "membrane species update kernel");
    }
  }
  // assignment rules in dependency order (updateAssignmentRules, spatialsimulator.cpp:1566-1581), on the new values
  for (size_t i = 0; i < sim->cell_rules.size(); ++i) {
    const Sb2CellRule& r = sim->cell_rules[i];
    Sb2CellRuleArgs a;
    memset(&a, 0, sizeof(a));
    a.nx = sim->g.nx; a.ny = sim->g.ny; a.nz = sim->g.nz; a.P = sim->P; a.PL = sim->PL; a.first = sim->first;
    a.ntoks = r.ntoks; a.tok = r.d_tok;
    a.out = r.target_kind == 1 ? sim->species[r.target].u[sim->species[r.target].cur] : sim->d_fields[r.target];
    a.mask = r.geo >= 0 ? sim->d_flags[r.geo] : NULL;
    for (size_t q = 0; q < S; ++q) a.species[q] = sim->species[q].u[sim->species[q].cur];
    for (size_t f = 0; f < sim->d_fields.size(); ++f) a.fields[f] = sim->d_fields[f];
    memcpy(a.consts, sim->consts, sizeof(a.consts));
    cudaError_t e = sb2_launch_cell_rule(a, sim->stream, &sim->launches);
    if (e != cudaSuccess) return cuda_fail(sim, e, "assignment rule kernel");
  }
  // pseudo-membrane points take the smaller of their two membrane neighbours (spatialsimulator.cpp:1614-1663)
  for (size_t i = 0; i < sim->mem_species.size(); ++i) {
    cudaError_t e = sb2_launch_mem_fill(sim->point_sets[sim->mem_species[i].set], sim->mem_species[i].u, sim->stream, &sim->launches);
    if (e != cudaSuccess) return cuda_fail(sim, e, "pseudo-membrane fill kernel");
  }
  sim->k0_carry_valid = sim->carry_k0;
  sim->in_step = false;
  return SB2_OK;
}
}  // namespace

int sb2_end_step(sb2_sim* sim) {
  if (!sim || !sim->in_step) return fail(sim, SB2_ERR_STATE, "sb2_end_step without sb2_begin_step");
  cudaSetDevice(sim->g.device);
  if (!sim->fuse_combine) {
    // advection between the last stage and the update (spatialsimulator.cpp:1484-1491)
    const std::vector<AdvPass> passes = advection_passes(sim);
    if (!passes.empty() && sim->g.own_hi > sim->g.own_lo)
      return fail(sim, SB2_ERR_STATE, "advection on a slab needs a halo exchange after every pass: step it with sb2_sharded_begin / sb2_sharded_next");
    // The passes of one species follow each other (X, then Y, then Z); different species share nothing.  Grids small enough to be
    // bound by the latency of a launch run them side by side: species 2, 3, ... on streams of their own, forked from and joined to
    // the step's stream (inside a captured step graph these become parallel branches).
    static const bool serial_adv = getenv("SB2_ADV_SERIAL") != NULL;
    int nadv = 0;
    for (size_t s = 0; s < sim->species.size(); ++s) nadv += sim->species[s].has_adv ? 1 : 0;
    // (only the fused 2-D passes: the two-kernel pass of 3-D grids shares one scratch array between the species)
    bool all_fused = sim->g.dim == 2;
    for (size_t i = 0; i < passes.size(); ++i) all_fused = all_fused && sim->species[passes[i].species].faces_alt[passes[i].axis] != NULL;
    const bool side_by_side = !serial_adv && nadv > 1 && all_fused && ncells(sim) <= (1 << 18);
    if (side_by_side) {
      if (!sim->adv_fork) CK(cudaEventCreateWithFlags(&sim->adv_fork, cudaEventDisableTiming));
      while ((int)sim->adv_streams.size() < nadv - 1) {
        cudaStream_t st; cudaEvent_t ev;
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        sim->adv_streams.push_back(st); sim->adv_join.push_back(ev);
      }
      CK(cudaEventRecord(sim->adv_fork, sim->stream));
    }
    const cudaStream_t main_stream = sim->stream;
    int branch = -1, last_species = -1;
    for (size_t i = 0; i < passes.size(); ++i) {
      if (side_by_side && passes[i].species != last_species) {
        if (branch >= 1) CK(cudaEventRecord(sim->adv_join[branch - 1], sim->stream));  // the previous species' branch is complete
        ++branch;
        last_species = passes[i].species;
        sim->stream = branch == 0 ? main_stream : sim->adv_streams[branch - 1];
        if (branch >= 1) CK(cudaStreamWaitEvent(sim->stream, sim->adv_fork, 0));
      }
      int rc = launch_advection_pass(sim, passes[i]);
      if (rc) { sim->stream = main_stream; return rc; }
    }
    if (side_by_side) {
      if (branch >= 1) CK(cudaEventRecord(sim->adv_join[branch - 1], sim->stream));
      sim->stream = main_stream;
      for (int b = 1; b <= branch; ++b) CK(cudaStreamWaitEvent(main_stream, sim->adv_join[b - 1], 0));
    }
  }
  return finish_step(sim);
}

// ---------------------------------------------------------------------------------------------
// Sharded stepping, orchestrated here: which rows run on which stream, what is exchanged with the neighbouring slabs and when
// (SURVEY.md §8e).  The caller only moves the rows sb2_sharded_next hands it, on the stream it names.
// ---------------------------------------------------------------------------------------------
namespace {
int ensure_shard_streams(sb2_sim* sim) {
  if (sim->sh_edge) return SB2_OK;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // (lowest, highest): the edge rows go first
  CK(cudaStreamCreateWithPriority(&sim->sh_edge, cudaStreamNonBlocking, hi));
  CK(cudaStreamCreateWithFlags(&sim->sh_comm, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&sim->ev_edge, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&sim->ev_interior, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&sim->ev_comm, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&sim->ev_ready, cudaEventDisableTiming));
  return SB2_OK;
}

void fill_exchange(sb2_sim* sim, sb2_exchange* x, const std::vector<double*>& bufs, int width) {
  const SlabRows r = slab_rows(sim);
  memset(x, 0, sizeof(*x));
  x->n = (int)bufs.size();
  for (size_t i = 0; i < bufs.size() && i < SB2_MAX_XBUF; ++i) x->ptr[i] = bufs[i];
  x->width = width;
  if (sim->g.dim == 3) { x->unit = sim->PL; x->base = sim->first - sim->P; }
  else { x->unit = sim->P; x->base = sim->first; }
  x->own_lo = r.own_lo; x->own_hi = r.own_hi;
  x->has_lo = r.own_lo > 0 ? 1 : 0;
  x->has_hi = r.own_hi < r.n_held ? 1 : 0;
  x->stream = (void*)sim->sh_comm;
  x->ready_event = (void*)sim->ev_ready;
  x->device = sim->g.device;
}
}  // namespace

// The schedule on which a step is four stage kernels and nothing else (the RK combine fused into the last one): no advection,
// membrane transport, membrane species, Dirichlet or non-zero flux conditions, per-step rules.  Known after sb2_begin_step.
static bool step_is_fused_only(const sb2_sim* sim) {
  return sim->fuse_combine && sim->transports.empty() && sim->blists[0].n_dirichlet == 0 && !sim->model_dirichlet && sim->mem_species.empty() &&
         sim->cell_rules.empty();
}

// Sharded steps of a slab on the fused schedule can trade the exchange per stage for redundant work: the four edge rows of u
// travel ONCE per step (behind the part of stage 1 that needs no halo), stage m then runs on rows [own_lo - 3 + m, own_hi + 3 - m)
// - the apron rows reach into the four halo rows and are recomputed on both sides.  One exchange and six launches per step
// instead of four exchanges and twelve launches, which is what strong scaling (short slabs) and callers with an expensive
// exchange need; the result is the same step bit for bit.  Every slab of a grid must be set alike (the caller knows all of them:
// sharded.py, SpatialSimulator::setDevices); models off the fused schedule keep the per-stage exchanges whatever is set here.
int sb2_sharded_set_apron(sb2_sim* sim, int on) {
  if (!sim) return SB2_ERR_ARG;
  sim->sh_apron = on != 0;
  return SB2_OK;
}

int sb2_sharded_begin(sb2_sim* sim, double t_arg, double dt, int time_slot) {
  if (!sim || !sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator not finalized");
  if (sim->in_step) return fail(sim, SB2_ERR_STATE, "sb2_sharded_begin inside a step");
  cudaSetDevice(sim->g.device);
  int rc = ensure_shard_streams(sim);
  if (rc) return rc;
  if ((rc = sb2_begin_step(sim, t_arg, dt, time_slot))) return rc;
  sim->sh_phase = 0;
  sim->sh_passes = (int)advection_passes(sim).size();
  sim->sh_first_stage = true;
  return SB2_OK;
}

// Runs the step up to its next halo exchange.  Returns 1 and fills *x when the caller has to exchange rows now (then call
// again), 0 when the step is complete, negative on error.
//   phases 0..7: stage m = phase / 2: even = edge rows (then exchange k_m, or the new u after a fused last stage), odd = the
//                other owned rows;   8 .. 8+P-1: advection pass p (then exchange u and the faces of its species);
//   8+P: update, boundary pass, membrane species, rules (then exchange u unless it travelled after the last stage)
int sb2_sharded_next(sb2_sim* sim, sb2_exchange* x) {
  if (!sim || !x) return SB2_ERR_ARG;
  if (!sim->in_step && sim->sh_phase != 1000) return fail(sim, SB2_ERR_STATE, "sb2_sharded_next without sb2_sharded_begin");
  cudaSetDevice(sim->g.device);
  const size_t S = sim->species.size();
  const cudaStream_t compute = sim->stream;
  if (sim->sh_awaiting) {
    // the exchange handed out by the previous call has been queued on the communication stream by now
    CK(cudaEventRecord(sim->ev_comm, sim->sh_comm));
    sim->sh_comm_valid = true;
    sim->sh_awaiting = false;
  }
  int rc;
  if (sim->sh_apron && step_is_fused_only(sim) && sim->sh_phase >= 0 && sim->sh_phase <= 1) {
    const SlabRows r = slab_rows(sim);
    const bool has_lo = r.own_lo > 0, has_hi = r.own_hi < r.n_held;
    if ((has_lo && r.own_lo < 4) || (has_hi && r.n_held - r.own_hi < 4) || r.own_hi - r.own_lo < 8)
      return fail(sim, SB2_ERR_STATE, "the apron schedule needs four halo rows on either side and at least eight owned rows");
    if (sim->sh_phase == 0) {
      // the four edge rows of the current u go to the neighbours (halo rows a host step or an upload left stale are covered too)
      std::vector<double*> bufs;
      for (size_t s = 0; s < S; ++s) bufs.push_back(sim->species[s].u[sim->species[s].cur]);
      CK(cudaEventRecord(sim->ev_ready, compute));
      CK(cudaStreamWaitEvent(sim->sh_comm, sim->ev_ready, 0));
      fill_exchange(sim, x, bufs, 4);
      sim->halo_stale = false;
      sim->sh_awaiting = true;
      sim->sh_phase = 1;
      if (has_lo || has_hi) return 1;
      sim->sh_awaiting = false;  // a slab without neighbours: nothing travels
    }
    const double dt = sim->cur_dt;
    // stage 1: the rows that read no halo row run beside the exchange, the rows around the edges once it has landed
    const int bi = has_lo ? r.own_lo + 1 : r.own_lo, ei = has_hi ? r.own_hi - 1 : r.own_hi;
    if ((rc = launch_stage_rows(sim, 0, dt, bi, ei))) return rc;
    if (sim->sh_comm_valid) CK(cudaStreamWaitEvent(compute, sim->ev_comm, 0));
    if (has_lo && (rc = launch_stage_rows(sim, 0, dt, r.own_lo - 3, bi))) return rc;
    if (has_hi && (rc = launch_stage_rows(sim, 0, dt, ei, r.own_hi + 3))) return rc;
    for (int m = 1; m < 4; ++m)
      if ((rc = launch_stage_rows(sim, m, dt, has_lo ? r.own_lo - (3 - m) : r.own_lo, has_hi ? r.own_hi + (3 - m) : r.own_hi))) return rc;
    for (size_t s = 0; s < S; ++s) sim->species[s].cur ^= 1;
    sim->k0_carry_valid = false;
    sim->in_step = false;
    sim->sh_comm_valid = false;  // (the halo rows of the new u are exchanged by the next step, whichever kind it is)
    sim->halo_stale = true;
    sim->sh_phase = -1;
    return 0;
  }
  if (sim->halo_stale && sim->sh_phase == 0) {
    // the last step came through sb2_step_host_finish, which leaves the halo rows of the new u unexchanged: refresh them first
    std::vector<double*> bufs;
    for (size_t s = 0; s < S; ++s) bufs.push_back(sim->species[s].u[sim->species[s].cur]);
    CK(cudaEventRecord(sim->ev_ready, compute));
    CK(cudaStreamWaitEvent(sim->sh_comm, sim->ev_ready, 0));
    fill_exchange(sim, x, bufs, sim->edge_w > sim->end_w ? sim->edge_w : sim->end_w);
    sim->halo_stale = false;
    sim->sh_awaiting = true;
    return 1;
  }
  for (;;) {
    const int ph = sim->sh_phase;
    if (ph == 1000) {  // the exchange that ended the step is in flight: everything later waits for it
      CK(cudaStreamWaitEvent(compute, sim->ev_comm, 0));
      sim->sh_phase = -1;
      return 0;
    }
    if (ph < 8 && (ph & 1) == 0) {
      const int m = ph / 2;
      // edge rows of stage m: they need the neighbours' rows of the previous exchange and this slab's rows of stage m-1
      if (sim->sh_comm_valid) CK(cudaStreamWaitEvent(sim->sh_edge, sim->ev_comm, 0));
      if (sim->sh_first_stage) {
        CK(cudaEventRecord(sim->ev_interior, compute));  // whatever the compute stream still has queued (uploads, earlier steps)
        sim->sh_first_stage = false;
      }
      CK(cudaStreamWaitEvent(sim->sh_edge, sim->ev_interior, 0));
      if (m > 0) CK(cudaStreamWaitEvent(compute, sim->ev_edge, 0));  // the interior of stage m reads the edge rows of stage m-1
      sim->multi_stream = true;
      sim->stream = sim->sh_edge;
      rc = sb2_stage(sim, m, 1);
      sim->stream = compute;
      if (rc) return rc;
      CK(cudaEventRecord(sim->ev_edge, sim->sh_edge));
      if (!sim->transports.empty() || !sim->mem_species.empty()) CK(cudaStreamWaitEvent(compute, sim->ev_edge, 0));  // the membrane work ran there
      sim->sh_phase = ph + 1;
      std::vector<double*> bufs;
      const bool last = m == 3;
      if (!last) for (size_t s = 0; s < S; ++s) bufs.push_back(sim->species[s].k[m]);
      else if (sim->fuse_combine) for (size_t s = 0; s < S; ++s) bufs.push_back(sim->species[s].u[sim->species[s].cur ^ 1]);
      if (bufs.empty()) continue;
      CK(cudaEventRecord(sim->ev_ready, sim->sh_edge));
      CK(cudaStreamWaitEvent(sim->sh_comm, sim->ev_ready, 0));
      fill_exchange(sim, x, bufs, sim->edge_w);
      sim->sh_awaiting = true;
      return 1;
    }
    if (ph < 8) {
      const int m = ph / 2;
      if ((rc = sb2_stage(sim, m, 2))) return rc;
      CK(cudaEventRecord(sim->ev_interior, compute));
      sim->sh_phase = ph + 1;
      if (sim->model_dirichlet && m < 3) {
        // the boundary pass of this stage may have rewritten u on the edge rows (Dirichlet): the neighbours need it for stage m+1
        std::vector<double*> bufs;
        for (size_t s = 0; s < S; ++s) bufs.push_back(sim->species[s].u[sim->species[s].cur]);
        CK(cudaEventRecord(sim->ev_ready, compute));
        CK(cudaStreamWaitEvent(sim->sh_comm, sim->ev_ready, 0));
        fill_exchange(sim, x, bufs, sim->edge_w);
        sim->sh_awaiting = true;
        return 1;
      }
      continue;
    }
    if (ph == 8) {
      // everything after the stages runs on the compute stream: it needs the edge rows and the halos of the last exchange
      CK(cudaStreamWaitEvent(compute, sim->ev_edge, 0));
      if (sim->sh_comm_valid) CK(cudaStreamWaitEvent(compute, sim->ev_comm, 0));
    }
    const int npass = sim->fuse_combine ? 0 : sim->sh_passes;
    if (ph - 8 < npass) {
      if (ph > 8 && sim->sh_comm_valid) CK(cudaStreamWaitEvent(compute, sim->ev_comm, 0));  // halos of the previous pass
      const std::vector<AdvPass> passes = advection_passes(sim);
      const AdvPass& p = passes[ph - 8];
      if ((rc = launch_advection_pass(sim, p))) return rc;
      sim->sh_phase = ph + 1;
      std::vector<double*> bufs;
      Sb2Species& sp = sim->species[p.species];
      bufs.push_back(sp.u[sp.cur]);
      for (int a = 0; a < sim->g.dim; ++a) if (sp.faces[a]) bufs.push_back(sp.faces[a]);
      CK(cudaEventRecord(sim->ev_ready, compute));
      CK(cudaStreamWaitEvent(sim->sh_comm, sim->ev_ready, 0));
      fill_exchange(sim, x, bufs, 2);
      sim->sh_awaiting = true;
      return 1;
    }
    // last phase
    if (npass > 0 && sim->sh_comm_valid) CK(cudaStreamWaitEvent(compute, sim->ev_comm, 0));
    const bool fused = sim->fuse_combine;
    if ((rc = finish_step(sim))) return rc;
    if (fused) {  // the new u travelled after the last stage; later work must see it
      if (sim->sh_comm_valid) CK(cudaStreamWaitEvent(compute, sim->ev_comm, 0));
      sim->sh_phase = -1;
      return 0;
    }
    std::vector<double*> bufs;
    for (size_t s = 0; s < S; ++s) bufs.push_back(sim->species[s].u[sim->species[s].cur]);
    CK(cudaEventRecord(sim->ev_ready, compute));
    CK(cudaStreamWaitEvent(sim->sh_comm, sim->ev_ready, 0));
    fill_exchange(sim, x, bufs, sim->end_w);
    sim->sh_awaiting = true;
    sim->sh_phase = 1000;
    return 1;
  }
}

namespace {
// Every slab of a group pulls the rows of one exchange from its neighbours (peer copies on the handles' communication streams,
// after the neighbour's producer event).  sims[i] owns the rows below sims[i+1]'s.
int group_pull(sb2_sim* const* sims, const std::vector<sb2_exchange>& xs, int n) {
  for (int i = 0; i < n; ++i) {
    sb2_sim* sim = sims[i];
    const sb2_exchange& x = xs[i];
    cudaSetDevice(sim->g.device);
    cudaStream_t st = (cudaStream_t)x.stream;
    for (int side = 0; side < 2; ++side) {
      const int j = side == 0 ? i - 1 : i + 1;
      if (j < 0 || j >= n || !(side == 0 ? x.has_lo : x.has_hi)) continue;
      const sb2_exchange& y = xs[j];
      if (y.n != x.n || y.width != x.width || y.unit != x.unit) return fail(sim, SB2_ERR_STATE, "neighbouring slabs disagree on an exchange");
      CK(cudaStreamWaitEvent(st, (cudaEvent_t)y.ready_event, 0));
      const size_t bytes = sizeof(double) * (size_t)x.unit * x.width;
      for (int b = 0; b < x.n; ++b) {
        // my halo rows below own_lo <- the neighbour's last owned rows; my halo rows from own_hi <- its first owned rows
        double* dst = (double*)x.ptr[b] + x.base + (int64_t)(side == 0 ? x.own_lo - x.width : x.own_hi) * x.unit;
        const double* src = (const double*)y.ptr[b] + y.base + (int64_t)(side == 0 ? y.own_hi - y.width : y.own_lo) * y.unit;
        CK(cudaMemcpyPeerAsync(dst, x.device, src, y.device, bytes, st));
      }
    }
  }
  // a slab may not overwrite rows a neighbour is still pulling: its communication stream waits for the neighbours' pulls
  for (int i = 0; i < n; ++i) {
    sb2_sim* sim = sims[i];
    cudaSetDevice(sim->g.device);
    if (!sim->ev_pulled) CK(cudaEventCreateWithFlags(&sim->ev_pulled, cudaEventDisableTiming));
    CK(cudaEventRecord(sim->ev_pulled, (cudaStream_t)xs[i].stream));
  }
  for (int i = 0; i < n; ++i) {
    sb2_sim* sim = sims[i];
    cudaSetDevice(sim->g.device);
    if (i > 0) CK(cudaStreamWaitEvent((cudaStream_t)xs[i].stream, sims[i - 1]->ev_pulled, 0));
    if (i + 1 < n) CK(cudaStreamWaitEvent((cudaStream_t)xs[i].stream, sims[i + 1]->ev_pulled, 0));
  }
  return SB2_OK;
}
}  // namespace

// The same step for several slabs held by ONE process (one handle per slab, on the same or on different devices): the rows
// move with peer copies on the handles' communication streams.  sims[i] owns the rows below sims[i+1]'s.
int sb2_group_step(sb2_sim* const* sims, int n, double t_arg, double dt, int time_slot) {
  if (!sims || n < 1) return SB2_ERR_ARG;
  int rc;
  for (int i = 0; i < n; ++i)
    if ((rc = sb2_sharded_begin(sims[i], t_arg, dt, time_slot))) return rc;
  std::vector<sb2_exchange> xs(n);
  for (;;) {
    int more = 0, done = 0;
    for (int i = 0; i < n; ++i) {
      rc = sb2_sharded_next(sims[i], &xs[i]);
      if (rc < 0) return rc;
      if (rc == 1) ++more; else ++done;
    }
    if (more && done) return fail(sims[0], SB2_ERR_STATE, "sb2_group_step: the slabs disagree on the schedule of the step");
    if (!more) return SB2_OK;
    if ((rc = group_pull(sims, xs, n))) return rc;
  }
}

namespace {
int one_plain_step(sb2_sim* sim, double t_arg, double dt, int time_slot) {
  int rc = sb2_begin_step(sim, t_arg, dt, time_slot);
  if (rc) return rc;
  for (int m = 0; m < 4; ++m) if ((rc = sb2_stage(sim, m, 0))) return rc;
  return sb2_end_step(sim);
}
}  // namespace

int sb2_step(sb2_sim* sim, double t_arg, double t_incr, double dt, int nsteps, int time_slot) {
  if (!sim || !sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator not finalized");
  int i = 0, rc;
  // Small grids are bound by launch latency (a step is 4 - 13 tiny kernels): run two ordinary steps (token records
  // uploaded, carried-flux state settled), capture the next two - one per parity of the u double buffer - into a CUDA
  // graph and replay it.  Only for models whose steps are all alike: no `time` symbol, nothing reading debug words.
  static const bool no_graph = getenv("SB2_NO_GRAPH") != NULL;
  const bool graphable = !no_graph && sim->own_stream && !sim->d_dbg && !sim->trace && time_slot < 0 &&
                         ncells(sim) <= (1 << 18) && nsteps >= 12;
  if (graphable) {
    for (; i < 2; ++i) if ((rc = one_plain_step(sim, t_arg + i * t_incr, dt, time_slot))) return rc;
    // steps per graph: an even number (the u double buffer is back in the captured state after two steps); longer graphs spread
    // the cost of a graph launch over more steps
    static const int gsteps_env = getenv("SB2_GRAPH_STEPS") ? atoi(getenv("SB2_GRAPH_STEPS")) : 8;
    const int gsteps = gsteps_env >= 2 ? (gsteps_env & ~1) : 2;
    const int pairs = (nsteps - i) / gsteps;
    // which halves of the double buffers are current: u of every species and, for advected species, the faces
    int par = 0;
    for (size_t s = 0; s < sim->species.size(); ++s) par = (par * 5 + sim->species[s].cur * 8 + sim->species[s].face_par) & 0xffff;
    par = (int)((unsigned)par % SB2_GRAPH_SLOTS);
    if (sim->g_key[par] != (int64_t)par_key(sim)) {  // another buffer state hashed to this slot: rebuild
      if (sim->gexec[par]) { cudaGraphExecDestroy(sim->gexec[par]); sim->gexec[par] = NULL; }
      sim->g_key[par] = (int64_t)par_key(sim);
    }
    bool fresh = false;
    if (!sim->gexec[par] || sim->g_dt[par] != dt || sim->g_version[par] != sim->version) {
      if (sim->gexec[par]) { cudaGraphExecDestroy(sim->gexec[par]); sim->gexec[par] = NULL; }
      // host-side state the two captured steps will advance
      std::vector<int> cur0(sim->species.size());
      for (size_t s = 0; s < sim->species.size(); ++s) cur0[s] = sim->species[s].cur;
      const std::vector<Sb2Species> species0(sim->species);
      const bool carry_valid0 = sim->k0_carry_valid;
      const int launches0 = sim->launches;
      cudaGraph_t graph = NULL;
      cudaError_t e = cudaStreamBeginCapture(sim->stream, cudaStreamCaptureModeThreadLocal);
      rc = SB2_OK;
      if (e == cudaSuccess) {
        sim->capturing = true;
        for (int k = 0; k < gsteps && !rc; ++k) rc = one_plain_step(sim, t_arg + (i + k) * t_incr, dt, time_slot);
        sim->capturing = false;
        e = cudaStreamEndCapture(sim->stream, &graph);
      }
      if (e == cudaSuccess && !rc && graph) e = cudaGraphInstantiate(&sim->gexec[par], graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (e != cudaSuccess || rc || !sim->gexec[par]) {
        // not capturable after all: put the host state back and take the ordinary path
        cudaGetLastError();
        sim->gexec[par] = NULL;
        sim->species = species0;
        for (size_t s = 0; s < sim->species.size(); ++s) sim->species[s].cur = cur0[s];
        sim->k0_carry_valid = carry_valid0;
        sim->launches = launches0;
        sim->in_step = false;
      } else {
        sim->g_dt[par] = dt; sim->g_version[par] = sim->version; sim->g_launches[par] = sim->launches - launches0;
        fresh = true;  // the host state already stands two steps further
      }
    }
    if (sim->gexec[par] && pairs > 0) {
      for (int r = 0; r < pairs; ++r) CK(cudaGraphLaunch(sim->gexec[par], sim->stream));
      sim->launches += sim->g_launches[par] * (pairs - (fresh ? 1 : 0));
      sim->steps_seen += gsteps * (pairs - (fresh ? 1 : 0));
      i += gsteps * pairs;
    } else if (fresh) {
      // captured but nothing to replay (cannot happen with nsteps >= 12): execute the captured pair once
      CK(cudaGraphLaunch(sim->gexec[par], sim->stream));
      i += gsteps;
    }
  }
  for (; i < nsteps; ++i)
    if ((rc = one_plain_step(sim, t_arg + i * t_incr, dt, time_slot))) return rc;
  return SB2_OK;
}

// One step with HOST inputs and outputs, pipelined over sub-slabs of the slowest axis: while sub-slab j is computed (all four
// stages, on a row range that shrinks by one row per stage so that every stage finds its neighbours' values computed:
// rows [r0-3+m, r1+3-m) in stage m), sub-slab j+1 is uploaded and sub-slab j-1 is downloaded, each on its own stream.  The
// redundant apron rows cost 12 row-stages per sub-slab.  The result is the same step, bit for bit.
//
// A handle that is itself a slab of a sharded grid (and holds four halo rows on either side) steps the same way: the rows of
// its two edge sub-slabs are uploaded first, the four edge rows of u travel to the neighbouring slabs ONCE (the exchange
// sb2_step_host_begin hands out), and the apron rows that reach into the halo are recomputed from them - no exchange per
// stage, so upload, compute and download of a rank overlap exactly as on a single GPU.
int sb2_step_host_begin(sb2_sim* sim, double t_arg, double dt, int time_slot, const double* const* in, int nslabs, sb2_exchange* x) {
  if (!sim || !sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator not finalized");
  if (!in) return fail(sim, SB2_ERR_ARG, "null argument");
  cudaSetDevice(sim->g.device);
  const int S = (int)sim->species.size();
  const bool d3 = sim->g.dim == 3;
  const SlabRows r = slab_rows(sim);
  const bool slab = sim->g.own_hi > sim->g.own_lo;
  const bool has_lo = r.own_lo > 0, has_hi = r.own_hi < r.n_held;
  if (slab && ((has_lo && r.own_lo < 4) || (has_hi && r.n_held - r.own_hi < 4)))
    return fail(sim, SB2_ERR_STATE, "sb2_step_host on a slab needs four halo rows on either side");
  int rc = sb2_begin_step(sim, t_arg, dt, time_slot);
  if (rc) return rc;
  if (!sim->fuse_combine || !sim->transports.empty() || sim->blists[0].n_dirichlet > 0 || sim->model_dirichlet || !sim->mem_species.empty() ||
      !sim->cell_rules.empty()) {
    sim->in_step = false;
    return fail(sim, SB2_ERR_UNFUSED, "sb2_step_host: the model needs the unfused schedule (advection, membrane transport, non-zero flux or Dirichlet)");
  }
  const int n = r.own_hi - r.own_lo;
  if (nslabs < 1) nslabs = 16;
  if (nslabs > n / 8) nslabs = n / 8 > 0 ? n / 8 : 1;
  if (!sim->h2d) {
    CK(cudaStreamCreateWithFlags(&sim->h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sim->d2h, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&sim->ev_start, cudaEventDisableTiming));
  }
  while ((int)sim->ev_up.size() < nslabs) {
    cudaEvent_t a, b;
    CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    sim->ev_up.push_back(a); sim->ev_done.push_back(b);
  }
  sim->host_nslabs = nslabs;
  const sb2_grid& g = sim->g;
  const int64_t plane_cells = (int64_t)g.nx * g.ny;
  auto upload_rows = [&](double* dev, const double* host, int r0, int r1) -> cudaError_t {
    // rows [r0, r1) of the slab axis from a compact host array (the rows this handle holds) into the padded device array
    // (rows without padding are one contiguous block on both sides: a plain copy)
    if (!d3 && sim->P == g.nx) return cudaMemcpyAsync(dev + sim->first + (int64_t)r0 * sim->P, host + (int64_t)r0 * g.nx,
                                                      sizeof(double) * (size_t)g.nx * (r1 - r0), cudaMemcpyHostToDevice, sim->h2d);
    if (!d3) return cudaMemcpy2DAsync(dev + sim->first + (int64_t)r0 * sim->P, sizeof(double) * sim->P, host + (int64_t)r0 * g.nx,
                                      sizeof(double) * g.nx, sizeof(double) * g.nx, r1 - r0, cudaMemcpyHostToDevice, sim->h2d);
    for (int z = r0; z < r1; ++z) {
      cudaError_t e = cudaMemcpy2DAsync(dev + sim->first + (int64_t)z * sim->PL, sizeof(double) * sim->P, host + (int64_t)z * plane_cells,
                                        sizeof(double) * g.nx, sizeof(double) * g.nx, g.ny, cudaMemcpyHostToDevice, sim->h2d);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  // the copy streams start after whatever the compute stream still has queued (previous steps, uploads)
  CK(cudaEventRecord(sim->ev_start, sim->stream));
  CK(cudaStreamWaitEvent(sim->h2d, sim->ev_start, 0));
  CK(cudaStreamWaitEvent(sim->d2h, sim->ev_start, 0));
  auto lo_of = [&](int j) { return r.own_lo + (int)((int64_t)j * n / nslabs); };
  // uploads are queued up front; on a slab the two edge sub-slabs go first (their edge rows travel to the neighbours)
  std::vector<int> order;
  if (slab) { order.push_back(0); if (nslabs > 1) order.push_back(nslabs - 1); for (int j = 1; j + 1 < nslabs; ++j) order.push_back(j); }
  else for (int j = 0; j < nslabs; ++j) order.push_back(j);
  for (size_t q = 0; q < order.size(); ++q) {
    const int j = order[q];
    for (int s = 0; s < S; ++s)
      if (in[s]) CK(upload_rows(sim->species[s].u[sim->species[s].cur], in[s], lo_of(j), lo_of(j + 1)));
    CK(cudaEventRecord(sim->ev_up[j], sim->h2d));
  }
  sim->host_exchange = false;
  if (slab && (has_lo || has_hi) && !x) {
    sim->in_step = false;
    return fail(sim, SB2_ERR_ARG, "sb2_step_host_begin on a slab with neighbours needs an exchange to fill");
  }
  if (slab && (has_lo || has_hi)) {
    if ((rc = ensure_shard_streams(sim))) return rc;
    CK(cudaStreamWaitEvent(sim->sh_comm, sim->ev_up[0], 0));
    CK(cudaStreamWaitEvent(sim->sh_comm, sim->ev_up[nslabs - 1], 0));
    CK(cudaEventRecord(sim->ev_ready, sim->sh_comm));
    std::vector<double*> bufs;
    for (int s = 0; s < S; ++s) bufs.push_back(sim->species[s].u[sim->species[s].cur]);
    fill_exchange(sim, x, bufs, 4);
    sim->host_exchange = true;
    return 1;
  }
  return 0;
}

namespace {
// second half of a host step: everything is queued, nothing waited for
int step_host_queue(sb2_sim* sim, double* const* out) {
  if (!sim || !sim->in_step) return fail(sim, SB2_ERR_STATE, "sb2_step_host_finish without sb2_step_host_begin");
  if (!out) return fail(sim, SB2_ERR_ARG, "null argument");
  cudaSetDevice(sim->g.device);
  const int S = (int)sim->species.size();
  const bool d3 = sim->g.dim == 3;
  const SlabRows r = slab_rows(sim);
  const int n = r.own_hi - r.own_lo, nslabs = sim->host_nslabs;
  const double dt = sim->cur_dt;
  const sb2_grid& g = sim->g;
  const int64_t plane_cells = (int64_t)g.nx * g.ny;
  auto download_rows = [&](const double* dev, double* host, int r0, int r1) -> cudaError_t {
    if (!d3 && sim->P == g.nx) return cudaMemcpyAsync(host + (int64_t)r0 * g.nx, dev + sim->first + (int64_t)r0 * sim->P,
                                                      sizeof(double) * (size_t)g.nx * (r1 - r0), cudaMemcpyDeviceToHost, sim->d2h);
    if (!d3) return cudaMemcpy2DAsync(host + (int64_t)r0 * g.nx, sizeof(double) * g.nx, dev + sim->first + (int64_t)r0 * sim->P,
                                      sizeof(double) * sim->P, sizeof(double) * g.nx, r1 - r0, cudaMemcpyDeviceToHost, sim->d2h);
    for (int z = r0; z < r1; ++z) {
      cudaError_t e = cudaMemcpy2DAsync(host + (int64_t)z * plane_cells, sizeof(double) * g.nx, dev + sim->first + (int64_t)z * sim->PL,
                                        sizeof(double) * sim->P, sizeof(double) * g.nx, g.ny, cudaMemcpyDeviceToHost, sim->d2h);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  if (sim->host_exchange) {
    // the caller queued the exchange of the edge rows on the communication stream: everything computed waits for it
    CK(cudaEventRecord(sim->ev_comm, sim->sh_comm));
    CK(cudaStreamWaitEvent(sim->stream, sim->ev_comm, 0));
  }
  auto lo_of = [&](int j) { return r.own_lo + (int)((int64_t)j * n / nslabs); };
  int rc;
  for (int j = 0; j < nslabs; ++j) {
    const int r0 = lo_of(j), r1 = lo_of(j + 1);
    CK(cudaStreamWaitEvent(sim->stream, sim->ev_up[j], 0));
    CK(cudaStreamWaitEvent(sim->stream, sim->ev_up[j + 1 < nslabs ? j + 1 : j], 0));  // the apron rows lie in the next sub-slab
    for (int m = 0; m < 4; ++m) {
      int b = r0 - (3 - m), e = r1 + (3 - m);
      if (b < 0) b = 0;
      if (e > r.n_held) e = r.n_held;
      if ((rc = launch_stage_rows(sim, m, dt, b, e))) return rc;
    }
    CK(cudaEventRecord(sim->ev_done[j], sim->stream));
    CK(cudaStreamWaitEvent(sim->d2h, sim->ev_done[j], 0));
    for (int s = 0; s < S; ++s)
      if (out[s]) CK(download_rows(sim->species[s].u[sim->species[s].cur ^ 1], out[s], r0, r1));
  }
  for (int s = 0; s < S; ++s) sim->species[s].cur ^= 1;
  sim->k0_carry_valid = false;
  sim->in_step = false;
  // a slab's halo rows of the new u are not valid (they were never exchanged): a later sharded step must not rely on them
  sim->sh_comm_valid = false;
  sim->halo_stale = sim->g.own_hi > sim->g.own_lo;
  return SB2_OK;
}
int step_host_wait(sb2_sim* sim) {
  cudaSetDevice(sim->g.device);
  CK(cudaStreamSynchronize(sim->d2h));
  CK(cudaStreamSynchronize(sim->stream));
  return SB2_OK;
}
}  // namespace

int sb2_step_host_finish(sb2_sim* sim, double* const* out) {
  int rc = step_host_queue(sim, out);
  if (rc) return rc;
  return step_host_wait(sim);
}

// The same for the slabs of a group held by one process: in[i] / out[i] are slab i's arrays of per-species host pointers (each
// to the first row that slab HOLDS); the edge rows move with peer copies, then every slab's pipeline runs beside the others'.
int sb2_group_step_host(sb2_sim* const* sims, int n, double t_arg, double dt, int time_slot, const double* const* const* in,
                        double* const* const* out, int nslabs) {
  if (!sims || n < 1 || !in || !out) return SB2_ERR_ARG;
  std::vector<sb2_exchange> xs(n);
  int rc, handed = 0;
  for (int i = 0; i < n; ++i) {
    rc = sb2_step_host_begin(sims[i], t_arg, dt, time_slot, in[i], nslabs, &xs[i]);
    if (rc < 0) {
      // nothing of the step has been computed: the slabs that had begun leave it again (their uploads are harmless)
      for (int j = 0; j < i; ++j) sims[j]->in_step = false;
      return rc;
    }
    handed += rc;
  }
  if (handed != (n > 1 ? n : 0)) return fail(sims[0], SB2_ERR_STATE, "sb2_group_step_host: the slabs disagree on the exchange");
  if (n > 1 && (rc = group_pull(sims, xs, n))) return rc;
  for (int i = 0; i < n; ++i) if ((rc = step_host_queue(sims[i], out[i]))) return rc;
  for (int i = 0; i < n; ++i) if ((rc = step_host_wait(sims[i]))) return rc;
  return SB2_OK;
}

int sb2_step_host(sb2_sim* sim, double t_arg, double dt, int time_slot, const double* const* in, double* const* out, int nslabs) {
  if (!sim || !in || !out) return fail(sim, SB2_ERR_ARG, "null argument");
  if (sim->g.own_hi > sim->g.own_lo) return fail(sim, SB2_ERR_STATE, "sb2_step_host needs a handle that owns its whole grid (a slab steps with sb2_step_host_begin / _finish)");
  int rc = sb2_step_host_begin(sim, t_arg, dt, time_slot, in, nslabs, NULL);
  if (rc < 0) return rc;
  return sb2_step_host_finish(sim, out);
}

int sb2_download_species(sb2_sim* sim, int species, double* out) {
  if (!sim || !out || species < 0 || species >= (int)sim->species.size()) return fail(sim, SB2_ERR_ARG, "bad species");
  cudaSetDevice(sim->g.device);
  return download_padded<double>(sim, out, sim->species[species].u[sim->species[species].cur]);
}

int sb2_upload_species(sb2_sim* sim, int species, const double* in) {
  if (!sim || !in || species < 0 || species >= (int)sim->species.size()) return fail(sim, SB2_ERR_ARG, "bad species");
  cudaSetDevice(sim->g.device);
  int rc = upload_interior<double>(sim, sim->species[species].u[sim->species[species].cur], in);
  if (rc) return rc;
  sim->halo_stale = false;  // (the caller uploads every held row, halo rows included)
  CK(cudaStreamSynchronize(sim->stream));
  return SB2_OK;
}

int sb2_download_stage(sb2_sim* sim, int species, int m, double* out) {
  if (!sim || !out || species < 0 || species >= (int)sim->species.size() || m < 0 || m > 3) return fail(sim, SB2_ERR_ARG, "bad species / stage");
  cudaSetDevice(sim->g.device);
  return download_padded<double>(sim, out, sim->species[species].k[m]);
}

int sb2_download_field(sb2_sim* sim, int field, double* out) {
  if (!sim || !out || field < 0 || field >= (int)sim->d_fields.size()) return fail(sim, SB2_ERR_ARG, "bad field");
  cudaSetDevice(sim->g.device);
  return download_padded<double>(sim, out, sim->d_fields[field]);
}

int sb2_upload_field(sb2_sim* sim, int field, const double* in) {
  if (!sim || !in || field < 0 || field >= (int)sim->d_fields.size()) return fail(sim, SB2_ERR_ARG, "bad field");
  cudaSetDevice(sim->g.device);
  int rc = upload_interior<double>(sim, sim->d_fields[field], in);
  if (rc) return rc;
  CK(cudaStreamSynchronize(sim->stream));
  return SB2_OK;
}

int sb2_upload_faces(sb2_sim* sim, int species, int axis, const double* in) {
  if (!sim || !in || species < 0 || species >= (int)sim->species.size() || axis < 0 || axis > 2) return fail(sim, SB2_ERR_ARG, "bad species / axis");
  cudaSetDevice(sim->g.device);
  Sb2Species& sp = sim->species[species];
  int rc;
  ++sim->version;
  if (!sp.faces[axis] && (rc = alloc_field(sim, &sp.faces[axis]))) return rc;
  if ((rc = upload_interior<double>(sim, sp.faces[axis], in))) return rc;
  CK(cudaStreamSynchronize(sim->stream));
  return SB2_OK;
}

int sb2_download_faces(sb2_sim* sim, int species, int axis, double* out) {
  if (!sim || !out || species < 0 || species >= (int)sim->species.size() || axis < 0 || axis > 2) return fail(sim, SB2_ERR_ARG, "bad species / axis");
  if (!sim->species[species].faces[axis]) return fail(sim, SB2_ERR_STATE, "species has no face values along this axis");
  cudaSetDevice(sim->g.device);
  return download_padded<double>(sim, out, sim->species[species].faces[axis]);
}

int sb2_min_dt(sb2_sim* sim, double dt, double* out) {
  if (!sim || !out || !sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator not finalized");
  cudaSetDevice(sim->g.device);
  double best = dt;
  for (size_t s = 0; s < sim->species.size(); ++s) {
    Sb2Species& sp = sim->species[s];
    Sb2MinDtArgs a;
    memset(&a, 0, sizeof(a));
    a.nx = sim->g.nx; a.ny = sim->g.ny; a.nz = sim->g.nz; a.dim = sim->g.dim;
    a.P = sim->P; a.PL = sim->PL; a.first = sim->first;
    a.delta[0] = sim->g.dx; a.delta[1] = sim->g.dy; a.delta[2] = sim->g.dz;
    a.dt = dt;
    a.flags = sim->d_flags[sp.geo];
    a.has_diff = sp.has_diff; a.has_adv = sp.has_adv;
    for (int ax = 0; ax < 3; ++ax) {
      a.dkind[ax] = sp.dkind[ax];
      a.dconst[ax] = sp.dkind[ax] == SB2_SRC_CONST ? sim->consts[sp.dsrc[ax]] : 0.0;
      a.dfield[ax] = sp.dkind[ax] == SB2_SRC_FIELD ? sim->d_fields[sp.dsrc[ax]] : NULL;
      a.akind[ax] = sp.akind[ax];
      a.aconst[ax] = sp.akind[ax] == SB2_SRC_CONST ? sim->consts[sp.asrc[ax]] : 0.0;
      a.afield[ax] = sp.akind[ax] == SB2_SRC_FIELD ? sim->d_fields[sp.asrc[ax]] : NULL;
    }
    if (!a.has_diff && !a.has_adv) continue;
    double r = dt;
    cudaError_t e = sb2_launch_min_dt(a, sim->d_scratch, sim->stream, &sim->launches, &r);
    if (e != cudaSuccess) return cuda_fail(sim, e, "min-dt reduction");
    if (r < best) best = r;
  }
  for (size_t i = 0; i < sim->mem_species.size(); ++i) {  // checkMemDiffusionStab
    double r = dt;
    cudaError_t e = sb2_launch_mem_min_dt(sim->mem_species[i], sim->consts, dt, sim->d_scratch, sim->stream, &sim->launches, &r);
    if (e != cudaSuccess) return cuda_fail(sim, e, "membrane min-dt reduction");
    if (r < best) best = r;
  }
  *out = best;
  return SB2_OK;
}

// Measurement aid: `reps` times the CIP-CSLR direction passes of one step, timed with CUDA events on the handle's stream.
int sb2_time_advection(sb2_sim* sim, double dt, int reps, double* ms_per_pass, int* passes_per_step) {
  if (!sim || !sim->finalized || !ms_per_pass || reps < 1) return fail(sim, SB2_ERR_ARG, "bad argument");
  if (sim->g.own_hi > sim->g.own_lo) return fail(sim, SB2_ERR_STATE, "sb2_time_advection needs a handle that owns its whole grid");
  cudaSetDevice(sim->g.device);
  sim->cur_dt = dt;
  const std::vector<AdvPass> passes = advection_passes(sim);
  if (passes_per_step) *passes_per_step = (int)passes.size();
  *ms_per_pass = 0.0;
  if (passes.empty()) return SB2_OK;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int rc = SB2_OK;
  for (size_t i = 0; i < passes.size() && !rc; ++i) rc = launch_advection_pass(sim, passes[i]);  // warm-up
  CK(cudaEventRecord(e0, sim->stream));
  for (int r = 0; r < reps && !rc; ++r)
    for (size_t i = 0; i < passes.size() && !rc; ++i) rc = launch_advection_pass(sim, passes[i]);
  CK(cudaEventRecord(e1, sim->stream));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc) return rc;
  *ms_per_pass = (double)ms / ((double)reps * (double)passes.size());
  ++sim->version;
  return SB2_OK;
}

int sb2_species_view(sb2_sim* sim, int species, int which, sb2_field_view* out) {
  if (!sim || !out || species < 0 || species >= (int)sim->species.size() || which < 0 || which > 5)
    return fail(sim, SB2_ERR_ARG, "bad species / selector");
  Sb2Species& s = sim->species[species];
  out->ptr = which == 0 ? (void*)s.u[s.cur] : which == 5 ? (void*)s.u[s.cur ^ 1] : (void*)s.k[which - 1];
  out->row_pitch = sim->P; out->plane_pitch = sim->PL; out->first_cell = sim->first; out->n_alloc = sim->nalloc;
  return SB2_OK;
}

int sb2_combine_is_fused(const sb2_sim* sim) { return sim && sim->fuse_combine ? 1 : 0; }

int64_t sb2_launch_count(const sb2_sim* sim) { return sim ? sim->launches : 0; }

int sb2_specialized(const sb2_sim* sim) { return sim && sim->jit ? 1 : 0; }

int sb2_specialize_now(sb2_sim* sim) {
  if (!sim || !sim->finalized) return fail(sim, SB2_ERR_STATE, "simulator not finalized");
  if (sim->in_step) return fail(sim, SB2_ERR_STATE, "sb2_specialize_now inside a step");
  cudaSetDevice(sim->g.device);
  if (sim->rd_id < 0 || (sim->jit_mode == 0 && !sim->jit)) return 0;  // SB2_JIT=0, or an earlier build failed
  specialise(sim, sim->jit_job != NULL, true);
  return sim->jit ? 1 : 0;
}

int sb2_specialize_info(const sb2_sim* sim, char* buf, int buflen) {
  if (!sim || !buf || buflen <= 0) return SB2_ERR_ARG;
  const std::string s = sim->jit_info.empty() ? std::string("interpreter kernel: first-generation stage kernel") : sim->jit_info;
  snprintf(buf, buflen, "%s", s.c_str());
  return (int)s.size();
}

int sb2_specialize_dry_run(const uint32_t* code, const uint32_t* arg, int n, int ns, int np, int w, int depth, int minb, int d3,
                           const char* cubin_path, char* log, int loglen) {
  if (n < 0 || (n > 0 && (!code || !arg)) || n > SB2_RD_MAX_TOKENS) return SB2_ERR_ARG;
  std::vector<Sb2RdTok> prog(n);
  for (int i = 0; i < n; ++i) { prog[i].code = code[i]; prog[i].arg = arg[i]; prog[i].c = 0.0; }
  Sb2RdJitSpec spec;
  spec.ns = ns; spec.np = np; spec.w = w; spec.depth = depth; spec.minb = minb; spec.d3 = d3;
  spec.np_final = 0; spec.minb_final = 0; spec.final_ldg = 0; spec.minb_first = 0;
  // a typical model: one geometry, every species diffuses with uniform coefficients
  spec.fold = 1; spec.G = 1; spec.hasdiff_mask = spec.fastpath_mask = (1u << ns) - 1u; spec.geo_bits = 0;
  std::vector<char> cubin;
  std::string l;
  const int rc = sb2_rd_jit_compile(sb2_rd_jit_source(spec, prog), &cubin, &l);
  if (log && loglen > 0) snprintf(log, loglen, "%s", l.c_str());
  if (rc != 0) return SB2_ERR_STATE;
  if (cubin_path) {
    FILE* f = fopen(cubin_path, "wb");
    if (!f) return SB2_ERR_ARG;
    fwrite(cubin.data(), 1, cubin.size(), f);
    fclose(f);
  }
  return SB2_OK;
}

int sb2_program_text(sb2_sim* sim, char* buf, int buflen) {
  if (!sim || !buf || buflen <= 0) return SB2_ERR_ARG;
  std::ostringstream os;
  for (size_t i = 0; i < sim->program.size(); ++i) {
    const uint32_t t = sim->program[i];
    const uint32_t code = t & 0xffffu;
    os << kOpcNames[code / SB2_MAX_DEPTH] << " " << (code % SB2_MAX_DEPTH) << " " << (t >> 16) << "\n";
  }
  const std::string s = os.str();
  snprintf(buf, buflen, "%s", s.c_str());
  return (int)s.size();
}

}  // extern "C"
