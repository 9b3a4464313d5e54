// sb2_internal.h — shared definitions between the C-ABI layer (sb2_device.cu) and the kernels.
#ifndef SB2_INTERNAL_H_
#define SB2_INTERNAL_H_

#ifndef __CUDACC_RTC__  // the run-time compiler (rd_jit.cpp) gets the fixed-width types and the public header from its own prelude
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#endif
#include "../../../include/spatial_b200.h"

// ---------------------------------------------------------------------------------------------
// Device bytecode.  The host-facing RPN stream (sb2_token) is a stack program whose stack pointer
// is known statically at every token, so the library resolves it once into register-addressed
// instructions: code = opcode * SB2_MAX_DEPTH + d, where d is the stack register the instruction
// acts on.  Every interpreter case then touches r[] with compile-time indices and the expression
// stack lives in registers.  "operand, binary-op" pairs are fused into one instruction.
// ---------------------------------------------------------------------------------------------
enum Sb2Opc {
  OPC_END = 0,
  OPC_PUSHC,  // r[d] = consts[arg]
  OPC_PUSHV,  // r[d] = staged species arg (u + c*dt*k_prev)
  OPC_PUSHF,  // r[d] = field arg
  OPC_ADD, OPC_SUB, OPC_MUL, OPC_DIV,         // r[d-1] = r[d-1] op r[d]
  OPC_GEQ, OPC_GT, OPC_LEQ, OPC_LT, OPC_AND, OPC_OR,
  OPC_ADDC, OPC_SUBC, OPC_MULC, OPC_DIVC,     // r[d] = r[d] op consts[arg]
  OPC_RSUBC, OPC_RDIVC,                       // r[d] = consts[arg] op r[d]
  OPC_GEQC, OPC_GTC, OPC_LEQC, OPC_LTC,
  OPC_ADDV, OPC_SUBV, OPC_MULV, OPC_DIVV,     // r[d] = r[d] op species arg
  OPC_UNARY,  // r[d] = f_arg(r[d])
  OPC_BINR,   // r[d-1] = g_arg(r[d-1], r[d])     (rare binaries: pow, log, xor, eq, neq)
  OPC_MOV0,   // r[0] = r[d]
  OPC_MOVUP,  // r[d+1] = r[d]
  OPC_MASK,   // statement mask = cell in domain of geo arg
  OPC_ACCSUB, // acc[d] -= consts[arg] * r[0]   (d = species index here)
  OPC_ACCADD, // acc[d] += consts[arg] * r[0]
  OPC_ACCSUB1, OPC_ACCADD1,    // stoichiometry exactly 1.0: acc[d] -/+= r[0]   (1.0 * x == x bit for bit)
  OPC_ACCSUBC1, OPC_ACCADDC1,  // constant law, stoichiometry 1.0: acc[d] -/+= consts[arg]
  OPC_ACCSUBM, OPC_ACCADDM,    // like ACCSUB / ACCADD, but only on the cells selected by the last MASK
  OPC_COUNT,
  // ---- rd_stage only: fused forms produced by the lowering pass (sb2_rd_lower); arg = species, c = inline constant
  OPC_ADDVC = OPC_COUNT, OPC_SUBVC, OPC_RSUBVC, OPC_MULVC, OPC_DIVVC, OPC_RDIVVC,  // r[d] = v op c   (R: c op v)
  OPC_ADDVV, OPC_SUBVV, OPC_MULVV, OPC_DIVVV,                                      // r[d] = v1 op v2  (arg = v1 | v2 << 8)
  OPC_XFER,                                                                        // acc[d] -= r[0]; acc[arg] += r[0]
  OPC_RD_COUNT
};

// field layout pads (in elements) around the guard rows, see sb2_create
#define SB2_FRONT_PAD 4
#define SB2_TAIL_PAD 264

#define SB2_CODE(opc, d) ((uint32_t)(opc) * SB2_MAX_DEPTH + (uint32_t)(d))

// rd_stage token record: 16 bytes, fetched with one shared-memory load; constants travel inline
struct Sb2RdTok { uint32_t code; uint32_t arg; double c; };
#define SB2_RD_MAX_TOKENS 384

struct Sb2SpeciesArgs {
  const double* u;      // current value
  double* u_out;        // fused combine target (other half of the double buffer)
  const double* kprev;  // k_{m-1}
  double* kout;         // k_m
  const double* kcarry; // carry != 0: what the accumulators start from (kout itself, or the membrane-transport buffer of the species)
  const double* k0;     // fused combine inputs
  const double* k1;
  int32_t geo;
  int32_t has_diff;
  int32_t fast2d;    // gen1 kernel: 2-D, uniform coefficients on both axes, dx^2 == dy^2 == 1: straight-line interior stencil
  int32_t fastpath;  // rd_stage: uniform coefficients on every axis of the grid and no division by delta^2
  int32_t dkind[3];  // SB2_SRC_*
  int32_t dsrc[3];
  double dconst[3];  // value of a uniform coefficient at launch time
};

// what both stage-kernel families need; small enough to keep kernel launches cheap
struct Sb2StageCore {
  int32_t nx, ny, nz, dim;
  int64_t P, PL, first;  // row pitch, plane pitch, offset of cell (0,0,0), all in elements
  int32_t row_begin, row_end;      // rows (2-D: y) or planes (3-D: z) of the slab axis to compute
  int32_t rows_per_block;
  int32_t np;      // cell pairs per thread of the kernel variant (strip width = 256*np cells)
  int32_t m;       // RK stage 0..3
  int32_t S, G;
  int32_t final;   // fused RK combine: write u_out instead of k3
  int32_t carry;   // 1: the accumulators start from kcarry (= kout: a carried flux); 2: from kcarry (the transport buffer), which is
                   // zero outside mixed strip rows - those keep their fast paths
  int32_t ntok;
  int32_t depth;   // deepest expression-stack slot the program touches + 1
  double cdt;      // rk[m]*dt
  double dt;
  double dx2[3];
  int32_t divide[3];  // 0 when dx2 == 1.0 exactly (x/1.0 == x, the divide is skipped bit-exactly)
  Sb2SpeciesArgs sp[SB2_MAX_SPECIES];
  const uint8_t* flags[SB2_MAX_GEOS];
  const double* fields[SB2_MAX_FIELDS];
  const uint8_t* cls;  // rd_stage: class byte per (row, strip): 0 mixed, 1 all interior, 2 all outside
  int32_t nstrips;     // strips per row of the class map
  uint32_t* dbg;       // rd_stage: debug words (SB2_DEBUG=1), else NULL; [0] = mbarrier waits that timed out
  int32_t rd_variant;  // rd_stage kernel variant chosen at finalize (-1: the first-generation kernel runs)
  int32_t nrd;         // rd_stage: token records (without the trailing OPC_END pair)
};

// first-generation kernel: program and constants table travel in the kernel parameters (constant memory)
struct Sb2StageArgs : Sb2StageCore {
  uint32_t tok[SB2_MAX_TOKENS + 2];  // + trailing OPC_END (the interpreter fetches one token ahead)
  double consts[SB2_MAX_CONSTS];
};

// rd_stage kernel: the token records live in a small device buffer (re-uploaded only when a constant changed) and are
// copied to shared memory by every CTA; 1 KB of parameters instead of 15 KB keeps the launch cheap on small grids
struct Sb2RdArgs : Sb2StageCore {
  const Sb2RdTok* rdtok;
};

#ifndef __CUDACC_RTC__
// launches (implemented in stage_kernel.cu)
cudaError_t sb2_launch_stage(const Sb2StageArgs& args, cudaStream_t stream, int* launches);
int sb2_stage_pairs(int nx, int nspecies);

// Grid and dynamic shared memory of an rd_stage launch (rd_stage.cuh describes the layouts): np cell pairs per thread,
// w warps per CTA, mode 0 / 1 / 2 = first / middle / last stage, tokb = bytes of token records held in shared memory.
// 2-D: a CTA owns a strip of 64*np cells and rows_per_block rows.  3-D: a strip, w rows and rows_per_block planes.
static inline void sb2_rd_launch_shape(const Sb2RdArgs& a, int np, int w, int mode, bool d3, int tokb, dim3* grid, size_t* smem,
                                       bool final_ldg = false) {
  const int sw = 64 * np, swp = sw + 4, narr = mode == 0 ? 1 : 2;
  const size_t ring = d3 ? (size_t)a.S * 3 * (w + 2) * swp : (size_t)a.S * (2 * w + 2) * swp;
  const size_t raw = (size_t)(d3 ? w + 2 : w) * a.S * narr * swp;
  const size_t cs = (mode == 2 && !(d3 && final_ldg)) ? (size_t)w * a.S * 2 * sw : 0;
  const size_t hdr = d3 ? (size_t)(8 * w * 8 + 511) / 512 * 512 : 512;  // mbarriers
  *smem = hdr + tokb + sizeof(double) * (ring + raw + cs);
  grid->x = (a.nx + sw - 1) / sw;
  if (d3) {
    grid->y = (a.ny + w - 1) / w;
    grid->z = (a.row_end - a.row_begin + a.rows_per_block - 1) / a.rows_per_block;
  } else {
    grid->y = (a.row_end - a.row_begin + a.rows_per_block - 1) / a.rows_per_block;
    grid->z = 1;
  }
}

// Launch configuration of a stage kernel: programmatic stream serialisation lets the CTAs of a stage kernel be scheduled
// while the previous kernel of the stream drains (the kernels call griddepcontrol.wait before they read anything a
// predecessor wrote); SB2_NO_PDL=1 turns it off.
static inline void sb2_rd_launch_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, dim3 grid, int threads, size_t smem,
                                        cudaStream_t stream) {
  static const bool pdl = getenv("SB2_NO_PDL") == NULL;
  memset(cfg, 0, sizeof(*cfg));
  cfg->gridDim = grid; cfg->blockDim = dim3(threads); cfg->dynamicSmemBytes = smem; cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg->attrs = attr;
  cfg->numAttrs = pdl ? 1 : 0;
}

// rd_stage kernel family (rd_stage.cuh, instantiated in rd_stage_ns*.cu)
struct Sb2RdVariant { int id, ns, np, w, depth; };
// picks the variant for a 2-D model (returns id or -1 when only the first-generation kernel applies)
int sb2_rd_pick_variant(int dim, int nx, int nspecies, int depth, Sb2RdVariant* out, int64_t ncells = 0);
cudaError_t sb2_rd_launch(const Sb2RdArgs& a, cudaStream_t stream);
// class map: cls[(z * ny + y) * nstrips + strip] for strips of sw cells, every held row of every held plane
cudaError_t sb2_rd_classify(const uint8_t* const* flags, int G, int nx, int ny, int nz, int64_t P, int64_t PL, int64_t first,
                            int sw, uint8_t* cls, int nstrips, cudaStream_t stream);

// model-specialised rd_stage kernels built at run time (rd_jit.cu)
struct Sb2RdJitSpec {
  int ns, np, w, depth, minb, d3;
  int minb_first;            // CTAs per SM the first-stage kernel is built for (0: like the others)
  int np_final, minb_final;  // the last stage may use a narrower strip (3-D: it also holds the k0 / k1 rows); 0: like the others
  int final_ldg;             // 3-D last stage: k0 / k1 are read with plain loads instead of bulk copies into shared-memory slots
  // model constants folded into the build when fold != 0: geometries, bit s of the masks = species s diffuses / takes the
  // uniform-coefficient stencil, 2 bits per species = its geometry
  int fold, G;
  unsigned hasdiff_mask, fastpath_mask, geo_bits;
};
struct Sb2RdJit;
std::string sb2_rd_jit_source(const Sb2RdJitSpec& spec, const std::vector<Sb2RdTok>& prog);
int sb2_rd_jit_compile(const std::string& src, std::vector<char>* cubin, std::string* log);  // needs no GPU
Sb2RdJit* sb2_rd_jit_build(const Sb2RdJitSpec& spec, const std::vector<Sb2RdTok>& prog, std::string* err);
cudaError_t sb2_rd_jit_launch(Sb2RdJit* j, const Sb2RdArgs& a, const double* kc, int nkc, cudaStream_t stream);
void sb2_rd_jit_free(Sb2RdJit* j);
// the same build on a background thread: start, poll (0 running, 1 compiled, -1 failed), finish (waits, loads, frees the
// job) or abandon (the thread finishes on its own)
struct Sb2RdJitJob;
Sb2RdJitJob* sb2_rd_jit_start(const Sb2RdJitSpec& spec, const std::vector<Sb2RdTok>& prog);
int sb2_rd_jit_poll(Sb2RdJitJob* job);
Sb2RdJit* sb2_rd_jit_finish(Sb2RdJitJob* job, std::string* err, double* secs);
void sb2_rd_jit_abandon(Sb2RdJitJob* job);
const char* sb2_rd_jit_library();  // path of the NVRTC library in use ("" when none was found)

struct Sb2CombineArgs {
  int32_t nx, ny, nz, dim, S;
  int64_t P, PL, first;
  double dt;
  double* u[SB2_MAX_SPECIES];
  const double* k[SB2_MAX_SPECIES][4];
  const uint8_t* flags[SB2_MAX_SPECIES];
};
cudaError_t sb2_launch_combine(const Sb2CombineArgs& args, cudaStream_t stream, int* launches);

#endif  // __CUDACC_RTC__

#endif  // SB2_INTERNAL_H_
