/* spatial_b200.h — C ABI of the B200 device layer behind SpatialSimulator::oneStep().
 *
 * This is the drop-in boundary for the hot path ("lower boundary").  The host side
 * (SBML parsing, geometry / membrane classification, boundary-type setup, AST → reverse-polish
 * streams) hands flattened arrays through these entry points; everything behind them is
 * hand-written sm_100a CUDA.  No C++ / torch / libSBML types cross this boundary.
 *
 * All grids are COMPACT: nx*ny*nz volume cells, x fastest; compact cell (i,j,k) is the reference's
 * doubled-grid point (2i,2j,2k) (reference SpatialSBML/spatialsimulator.cpp:354-358).
 *
 * Every function returns 0 on success and a negative sb2 error code otherwise;
 * sb2_last_error() gives the message.  A handle is used from one host thread at a time.
 * There is no CPU fallback: sb2_create fails if no CUDA device is usable.
 */
#ifndef SPATIAL_B200_H_
#define SPATIAL_B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define SB2_API __attribute__((visibility("default")))
#else
#define SB2_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SB2_OK 0
#define SB2_ERR_ARG (-1)
#define SB2_ERR_CUDA (-2)
#define SB2_ERR_STATE (-3)
#define SB2_ERR_LIMIT (-4) /* model exceeds a documented device limit (species / depth / tokens) */
#define SB2_ERR_UNFUSED (-5) /* sb2_step_host: the model needs the unfused schedule; nothing was computed, step with sb2_step instead */

#define SB2_MAX_SPECIES 8   /* staged (RK) volume species per simulator */
#define SB2_MAX_GEOS 4      /* volume geometries that host staged species */
#define SB2_MAX_FIELDS 16   /* read-only fields (coordinates, spatial parameters) */
#define SB2_MAX_CONSTS 512
#define SB2_MAX_TOKENS 1024
#define SB2_MAX_DEPTH 8     /* expression stack depth after operand fusion */

typedef struct sb2_sim sb2_sim;

/* Grid description.  Replaces the Xdiv/Ydiv/Zdiv/deltaX.. members the reference computes in
 * initFromModel (spatialsimulator.cpp:281-358) and setParameterInfo (setInfoFunction.cpp:296-342). */
typedef struct {
  int32_t nx, ny, nz; /* compact cells per axis (nz = 1 in 2-D) */
  int32_t dim;        /* 2 or 3 */
  double dx, dy, dz;  /* deltaX/Y/Z */
  int32_t device;     /* CUDA device ordinal */
  /* Slab sharding.  The slab axis is y in 2-D and z in 3-D.  The handle HOLDS
   * ny (2-D) / nz (3-D) rows of a larger grid: the rows it owns, [own_lo, own_hi) in local
   * indices, plus halo rows of the neighbouring slabs on either side, which the caller keeps up to
   * date (sb2_species_view + its own exchange).  Only owned rows are computed.  row0 = index of
   * local row 0 in the whole grid.  own_lo == own_hi == 0: the handle owns every row it holds. */
  int32_t own_lo, own_hi, row0;
} sb2_grid;

/* Per-cell geometry flag byte (replaces GeometryInfo::isDomain + bType, mystruct.h:28-61):
 * bit0 = isDomain==1, bits1..6 = isBofXp, isBofXm, isBofYp, isBofYm, isBofZp, isBofZm. */
#define SB2_F_DOMAIN 0x01
#define SB2_F_XP 0x02
#define SB2_F_XM 0x04
#define SB2_F_YP 0x08
#define SB2_F_YM 0x10
#define SB2_F_ZP 0x20
#define SB2_F_ZM 0x40

/* Reverse-polish token (replaces one slot of reversePolishInfo {varList, deltaList, constList,
 * opfuncList}, mystruct.h:20-26, as filled by parseAST, astFunction.cpp:18-107). */
typedef struct {
  uint8_t kind;  /* SB2_TOK_* */
  uint8_t op;    /* SB2_OP_* when kind == SB2_TOK_OP */
  uint16_t slot; /* VAR: species id; FIELD: field id; CONST: constants-table slot */
} sb2_token;

#define SB2_TOK_OP 0
#define SB2_TOK_VAR 1    /* staged species: u + rk[m]*dt*k_{m-1} (calcPDE.cpp:245-248) */
#define SB2_TOK_FIELD 2  /* raw field value (calcPDE.cpp:249-251) */
#define SB2_TOK_CONST 3  /* *constList[i] (calcPDE.cpp:253-255) */
#define SB2_TOK_NOP 4    /* unresolved name: an all-zero slot, acts as a pop (astFunction.cpp:71-88) */
#define SB2_TOK_MEMVAR 5 /* membrane species (slot = id from sb2_add_mem_species): value at the point + rk[m]*dt*k_{m-1} */

/* Operator codes: our own enumeration of the cases of the interpreter switch at
 * calcPDE.cpp:259-426 (never libSBML's ASTNodeType_t values). */
enum {
  SB2_OP_PLUS = 1, SB2_OP_MINUS, SB2_OP_TIMES, SB2_OP_DIVIDE, SB2_OP_POWER,
  SB2_OP_LOG, /* log(x)/log(base) */
  SB2_OP_AND, SB2_OP_OR, SB2_OP_XOR,
  SB2_OP_EQ, SB2_OP_GEQ, SB2_OP_GT, SB2_OP_LEQ, SB2_OP_LT, SB2_OP_NEQ,
  /* unary */
  SB2_OP_ABS = 32, SB2_OP_ARCCOS, SB2_OP_ARCCOSH, SB2_OP_ARCCSC, SB2_OP_ARCSIN, SB2_OP_ARCTAN,
  SB2_OP_COS, SB2_OP_COSH, SB2_OP_CSC, SB2_OP_EXP, SB2_OP_LN, SB2_OP_ROOT, SB2_OP_SIN,
  SB2_OP_SINH, SB2_OP_TAN, SB2_OP_TANH, SB2_OP_NOT,
  /* listed-but-empty cases: the reference only pops (calcPDE.cpp:297-302 etc.) */
  SB2_OP_POP_ONLY = 64
};

/* Species reference of a reaction (replaces reactionInfo::spRefList / srStoichiometry /
 * isVariable, mystruct.h:83-92).  species < 0: reference to a uniform species (skipped). */
typedef struct {
  int32_t species;
  int32_t is_variable;
  double stoichiometry;
} sb2_ref;

/* sb2_ref.species >= SB2_MEM_SPECIES names the membrane species (species - SB2_MEM_SPECIES) */
#define SB2_MEM_SPECIES 0x1000
#define SB2_MAX_MEM_SPECIES 8

#define SB2_SRC_NONE 0
#define SB2_SRC_CONST 1 /* uniform: constants-table slot */
#define SB2_SRC_FIELD 2 /* field id */

#define SB2_BC_NONE 0
#define SB2_BC_NEUMANN 1
#define SB2_BC_DIRICHLET 2

/* sides, in the reference's boundaryIndex order (mystruct.h:12-14) */
enum { SB2_XMAX = 0, SB2_XMIN, SB2_YMAX, SB2_YMIN, SB2_ZMAX, SB2_ZMIN };

SB2_API int sb2_create(const sb2_grid* grid, sb2_sim** out);
SB2_API void sb2_destroy(sb2_sim* sim);
SB2_API const char* sb2_last_error(const sb2_sim* sim); /* sim may be NULL: last create error */

/* Run all work of this handle on an externally owned CUDA stream (cudaStream_t as void*), e.g.
 * the caller's current torch stream so that its CUDA events bracket the kernels. */
SB2_API int sb2_set_stream(sb2_sim* sim, void* cuda_stream);

/* Volume geometry that hosts staged species: nx*ny*nz flag bytes.  Returns geo id >= 0.
 * Replaces GeometryInfo::isDomain/bType/domainIndex for one compartment
 * (spatialsimulator.cpp:411-717, boundaryFunction.cpp:13-217). */
SB2_API int sb2_add_geometry(sb2_sim* sim, const uint8_t* cell_flags);

/* Staged species field (value[] + delta[4N], setInfoFunction.cpp:121-126).  Returns species id. */
SB2_API int sb2_add_species(sb2_sim* sim, int geo, const double* init);
/* Read-only field (coordinate x/y/z, spatial parameter, constant spatial species). Returns field id. */
SB2_API int sb2_add_field(sb2_sim* sim, const double* values);

SB2_API int sb2_set_constants(sb2_sim* sim, const double* table, int n);
SB2_API int sb2_set_constant(sb2_sim* sim, int slot, double value); /* setParameter path (spatialsimulator.cpp:167-223) */

/* kind: 0 = reaction (reactants n_reactants first, then products; calcPDE.cpp:432-446),
 *       1 = rate rule (refs[0] += stoich*rate; calcPDE.cpp:447-449).
 * geo: the domain whose cells the law is evaluated on (spatialsimulator.cpp:1381-1390). */
SB2_API int sb2_add_reaction(sb2_sim* sim, int geo, int kind, const sb2_token* toks, int ntoks,
                     const sb2_ref* refs, int nrefs, int n_reactants);

/* Diffusion coefficient of a species along axis 0..2 (variableInfo::diffCInfo; calcPDE.cpp:486-511). */
SB2_API int sb2_set_diffusion(sb2_sim* sim, int species, int axis, int src_kind, int src);
/* Advection velocity along axis (variableInfo::adCInfo; calcPDE.cpp:564-869). */
SB2_API int sb2_set_advection(sb2_sim* sim, int species, int axis, int src_kind, int src);
/* A field-valued velocity is also read BETWEEN the cells (the divergence correction of the face value, calcPDE.cpp:623-625,
 * takes it at the odd doubled-grid points): field[cell] = the velocity at the face between cell and cell + e_axis. */
SB2_API int sb2_set_advection_faces(sb2_sim* sim, int species, int axis, int field);
/* Boundary condition of a species on one side (variableInfo::boundaryInfo; calcPDE.cpp:871-1034). */
SB2_API int sb2_set_boundary(sb2_sim* sim, int species, int side, int bc_kind, int src_kind, int src);

/* Membrane transport reaction (calcMemTransport, calcPDE.cpp:1036-1482).  The host passes the
 * membrane points that take part (isDomain==1, off the frame), their axis (0/1/2 from the
 * membrane bType), fabs(n_axis), and per species reference the compact index of the adjacent
 * cells on either side (-1 when not in that species' domain) plus the cell two further out used
 * for the 1.5/-0.5 extrapolation of volume operands (-1 when unavailable). */
typedef struct {
  int32_t axis;      /* 0 x, 1 y, 2 z */
  double abs_normal; /* fabs(nuVec[index].n{x,y,z}) */
} sb2_mem_point;
/* Binding (one side of the reaction is a membrane species, spatialsimulator.cpp:1440-1450 and the "binding" branches at
 * calcPDE.cpp:1405-1407 etc.): a SB2_TOK_MEMVAR operand is read at the membrane point itself — operand_cells[k][p][0] holds
 * the point's STORAGE index in the species' point set, [1] is -1 — and a reference to a membrane species (SB2_MEM_SPECIES + id)
 * receives fabs(n) * stoich * rate / (delta / 2.0) at target_cells[r][p][0] = the point's storage index ([1] = -1). */
SB2_API int sb2_add_mem_transport(sb2_sim* sim, const sb2_token* toks, int ntoks, const sb2_ref* refs, int nrefs,
                          int n_reactants, const sb2_mem_point* pts, int npts,
                          const int32_t* operand_cells /* [n VAR tokens][npts][2]: near, far compact cell or -1 */,
                          const double* operand_values /* [n FIELD tokens][npts]: value at the membrane point */,
                          const int32_t* target_cells /* [nrefs][npts][2]: plus side, minus side compact cell or -1 */);

/* ---- species that live ON a membrane ----------------------------------------------------------------------------
 * The reference keeps a membrane species in a full doubled-grid array and touches it only at the membrane's points
 * (GeometryInfo::domainIndex, isDomain == 1) and pseudo-membrane points (isDomain == 2, spatialsimulator.cpp:719-1038).
 * Here such a species is an array over a POINT SET: the host numbers the points it needs storage for (membrane points,
 * pseudo-membrane points) 0 .. npts-1 and passes every list below in those storage indices. */
SB2_API int sb2_add_point_set(sb2_sim* sim, int npts);                       /* returns set id */
SB2_API int sb2_add_mem_species(sb2_sim* sim, int set, const double* init);  /* value[] + delta[4N] on the set; returns id */
/* Reaction (kind 0) or rate rule (kind 1) between membrane species, evaluated at the listed points of the set in list
 * order (reversePolishRK over the membrane's domainIndex, calcPDE.cpp:224-451).  Tokens: SB2_TOK_MEMVAR operands are read
 * at the point; SB2_TOK_FIELD / SB2_TOK_VAR operands (coordinates, parameter fields, volume species, whose values at a
 * membrane point never change) come from operand_values[k][p], k counting those tokens in stream order.  refs name
 * membrane species (SB2_MEM_SPECIES + id); others are skipped. */
SB2_API int sb2_add_mem_reaction(sb2_sim* sim, int set, int kind, const sb2_token* toks, int ntoks, const sb2_ref* refs, int nrefs,
                                 int n_reactants, const int32_t* points, int npts, const double* operand_values);
/* Surface diffusion of a membrane species (calcMemDiffusion, calcPDE.cpp:1484-1581) over the Voronoi neighbours that
 * setVoronoiInfo (setInfoFunction.cpp:1134-1546) finds: point points[p] has the terms term_start[p] .. term_start[p+1]-1, each
 * with a neighbour (storage index), the facet length s_ij and the distance d_ij, in the reference's order (xy, yz, xz planes, two
 * neighbours each); area[p] = sum(d_ij * s_ij) / 2 (2-D) or / 4 (3-D).  The coefficient is a constants-table slot
 * (SB2_SRC_CONST) or one value per listed point (SB2_SRC_FIELD, d_values). */
SB2_API int sb2_set_mem_diffusion(sb2_sim* sim, int mem_species, int src_kind, int src, const double* d_values, const int32_t* points,
                                  int npts, const int32_t* term_start, const int32_t* adj, const double* s_ij, const double* d_ij,
                                  const double* area);
/* Points a membrane species is advanced on by updateSpecies (its geometry's domainIndex, spatialsimulator.cpp:1526-1537). */
SB2_API int sb2_set_mem_update_points(sb2_sim* sim, int mem_species, const int32_t* points, int npts);
/* Pseudo-membrane fill after every step (updateAssignmentRules, spatialsimulator.cpp:1614-1663): for every membrane species
 * of the set, value[target[i]] = min(value[a[i]], value[b[i]]), in list order. */
SB2_API int sb2_set_mem_fill(sb2_sim* sim, int set, const int32_t* target, const int32_t* a, const int32_t* b, int n);
SB2_API int sb2_download_mem_species(sb2_sim* sim, int mem_species, double* out);   /* npts of its set */
SB2_API int sb2_upload_mem_species(sb2_sim* sim, int mem_species, const double* in);
SB2_API int sb2_download_mem_stage(sb2_sim* sim, int mem_species, int m, double* out);

/* Assignment rule re-evaluated after every step, in the order given (updateAssignmentRules, spatialsimulator.cpp:1566-1581;
 * reversePolishInitial, calcPDE.cpp:21-222): the stream is evaluated with the raw current values (no RK offset) and written to
 * a read-only field (target_kind 0, every cell: a rule on a parameter) or to a staged species (target_kind 1, only the cells
 * of geometry `geo`: a rule on a species). */
SB2_API int sb2_add_assignment_rule(sb2_sim* sim, int target_kind, int target, int geo, const sb2_token* toks, int ntoks);

SB2_API int sb2_finalize(sb2_sim* sim);

/* nsteps × oneStep (performStep + updateValues, spatialsimulator.cpp:1236-1544).  The `time`
 * constant slot is set to (t_arg + i*t_incr)*dt for step i, reproducing sim_time = t*dt
 * (spatialsimulator.cpp:1358).  time_slot < 0: model does not use time. */
SB2_API int sb2_step(sb2_sim* sim, double t_arg, double t_incr, double dt, int nsteps, int time_slot);

/* One oneStep whose inputs and outputs are HOST arrays (compact nx*ny*nz per staged species, ideally pinned; in[s] /
 * out[s] may be NULL to skip that species' upload / download): the grid is cut into nslabs slabs of the slowest axis
 * (<= 0: 16) and upload, the four stages and download of successive slabs overlap on three streams.  Replaces the
 * sequence "write through getVariable pointers, oneStep, read them back" (spatialsimulator.h:88, :209) for callers
 * that exchange the whole state with the host every step.  Only for models on the fused schedule (no advection,
 * membrane transport, Dirichlet or non-zero flux conditions): SB2_ERR_UNFUSED otherwise (nothing computed), and the caller falls back. */
SB2_API int sb2_step_host(sb2_sim* sim, double t_arg, double dt, int time_slot, const double* const* in, double* const* out, int nslabs);
/* The same in two halves, which is how a SLAB of a sharded grid (holding four halo rows on either side; host arrays cover the rows
 * the handle holds) steps from host memory: _begin queues the uploads, the rows of the two edge sub-slabs first, and - when it
 * returns 1 - hands out ONE exchange of the four edge rows of u with the neighbouring slabs (sb2_exchange below; queue it on
 * x->stream); _finish recomputes the apron rows that reach into the halo, so no exchange per stage is needed, runs the stages
 * and the downloads and waits for them.  Upload, compute and download of every rank overlap as on a single GPU. */
struct sb2_exchange_s;
SB2_API int sb2_step_host_begin(sb2_sim* sim, double t_arg, double dt, int time_slot, const double* const* in, int nslabs, struct sb2_exchange_s* x);
SB2_API int sb2_step_host_finish(sb2_sim* sim, double* const* out);

SB2_API int sb2_download_species(sb2_sim* sim, int species, double* out);      /* compact nx*ny*nz */
SB2_API int sb2_upload_species(sb2_sim* sim, int species, const double* in);
SB2_API int sb2_download_stage(sb2_sim* sim, int species, int m, double* out); /* k_m, for tests */
SB2_API int sb2_upload_field(sb2_sim* sim, int field, const double* in);
SB2_API int sb2_download_field(sb2_sim* sim, int field, double* out);  /* fields change when an assignment rule targets them */
/* CIP-CSLR face point values of an advected species (the odd doubled-grid points between two cells
 * along `axis`, calcPDE.cpp:588-603): faces[cell] = value at the face between cell and cell + e_axis. */
SB2_API int sb2_upload_faces(sb2_sim* sim, int species, int axis, const double* in);
SB2_API int sb2_download_faces(sb2_sim* sim, int species, int axis, double* out);

/* Measurement aid (bench.py): runs the CIP-CSLR direction passes of one step `reps` times and reports the mean duration of a
 * pass (CUDA events on the handle's stream) and how many passes a step has.  The state is advected along. */
SB2_API int sb2_time_advection(sb2_sim* sim, double dt, int reps, double* ms_per_pass, int* passes_per_step);

/* checkStability.cpp:10-52 / 119-154: largest admissible dt (min-reduction over the domain). */
SB2_API int sb2_min_dt(sb2_sim* sim, double dt, double* out);

/* Device pointers for callers that exchange slab halos themselves (multi-GPU).
 * which: 0 = u (current), 1..4 = k0..k3, 5 = the other half of u's double buffer (where a fused
 * last stage writes the new u before sb2_end_step makes it current).  row_pitch / plane_pitch in
 * doubles; first_cell = offset of compact cell (0,0,0) inside the allocation (one guard row / plane
 * precedes it). */
typedef struct {
  void* ptr;
  int64_t row_pitch, plane_pitch, first_cell, n_alloc;
} sb2_field_view;
SB2_API int sb2_species_view(sb2_sim* sim, int species, int which, sb2_field_view* out);

/* Split-phase stepping for slab halo exchange.  phase: 0 = every owned row, 1 = only the owned
 * rows / planes next to a neighbouring slab (what the neighbours need next), 2 = the other owned rows. */
SB2_API int sb2_begin_step(sb2_sim* sim, double t_arg, double dt, int time_slot);
SB2_API int sb2_stage(sb2_sim* sim, int m, int phase);
SB2_API int sb2_stage_on(sb2_sim* sim, int m, int phase, void* cuda_stream); /* the same on another stream (ordering: caller's events) */
SB2_API int sb2_end_step(sb2_sim* sim);
/* ---- sharded stepping, orchestrated by the library -------------------------------------------------------
 * A handle that owns rows [own_lo, own_hi) of a larger grid steps with sb2_sharded_begin + a loop over sb2_sharded_next.  The
 * library decides which rows run when and on which of its streams (edge rows first, on a high-priority stream, so that their
 * transfer overlaps the interior kernel) and what has to travel: whenever sb2_sharded_next returns 1 the caller exchanges the
 * rows described by *x with the neighbouring slabs — for every buffer, send the first `width` owned rows to the slab below
 * and the last `width` owned rows to the slab above, receive the neighbours' rows into the `width` halo rows on either side —
 * queued on x->stream (a cudaStream_t that already waits for the producing kernels), with NCCL send / recv, peer copies, or
 * anything else that is stream-ordered, and calls sb2_sharded_next again.  0 = the step is complete.  Row r of a buffer is
 * ptr + base + r * unit (doubles).  What travels per step: k_m after each of the first three stages (and u where a Dirichlet
 * condition rewrites it), the new u after the last one; for models with advection u and the face values after every
 * CIP-CSLR pass (two rows); the new u after the update when it is not fused into the last stage. */
#define SB2_MAX_XBUF 40
typedef struct sb2_exchange_s {
  int32_t n;                 /* buffers */
  void* ptr[SB2_MAX_XBUF];   /* device arrays of doubles */
  int32_t width;             /* rows (2-D) / planes (3-D) per side */
  int64_t unit, base;        /* doubles per row / plane; offset of local row 0 */
  int32_t own_lo, own_hi;    /* owned local rows */
  int32_t has_lo, has_hi;    /* a neighbouring slab exists below / above */
  void* stream;              /* cudaStream_t to queue the exchange on */
  void* ready_event;         /* cudaEvent_t recorded when the rows to send are final (for callers that PULL from a neighbour) */
  int32_t device;
} sb2_exchange;
/* Apron schedule (before the first sharded step; every slab of a grid alike): models on the fused schedule exchange the four
 * edge rows of u ONCE per step and recompute the apron rows that reach into the halo, instead of exchanging after every stage:
 * one exchange and six launches per step instead of four and twelve.  Needs four halo rows on the sides with a neighbour and
 * at least eight owned rows per slab; other models keep the per-stage exchanges.  Bit-identical results either way. */
SB2_API int sb2_sharded_set_apron(sb2_sim* sim, int on);
SB2_API int sb2_sharded_begin(sb2_sim* sim, double t_arg, double dt, int time_slot);
SB2_API int sb2_sharded_next(sb2_sim* sim, sb2_exchange* x);
/* One step of n slabs held by ONE process (handles on the same or on different devices; sims[i] lies below sims[i+1]): the
 * loop above with peer copies between the handles.  A C++ caller of SpatialSimulator gets its multi-GPU step from this. */
SB2_API int sb2_group_step(sb2_sim* const* sims, int n, double t_arg, double dt, int time_slot);
/* sb2_step_host for the slabs of such a group: in[i] / out[i] are slab i's arrays of per-species host pointers, each to the
 * first row slab i HOLDS.  One peer copy of four edge rows of u per neighbour and step; every slab's upload / stages / download
 * pipeline then runs beside the others'.  SB2_ERR_UNFUSED (nothing computed) when the model needs the unfused schedule. */
SB2_API int sb2_group_step_host(sb2_sim* const* sims, int n, double t_arg, double dt, int time_slot, const double* const* const* in,
                                double* const* const* out, int nslabs);

/* After sb2_begin_step: 1 when the RK combine of this step is fused into stage 3 (the new u is
 * written to view 5 by sb2_stage(3, ...)), 0 when sb2_end_step computes it in place. */
SB2_API int sb2_combine_is_fused(const sb2_sim* sim);

/* Counters: kernels launched by this handle since creation. */
SB2_API int64_t sb2_launch_count(const sb2_sim* sim);
/* Human readable listing of the compiled device program (for tests / debugging). */
SB2_API int sb2_program_text(sb2_sim* sim, char* buf, int buflen);


/* Model-specialised stage kernel.  The reference interprets a reaction's reverse-polish stream at every grid
 * point (reversePolishRK, calcPDE.cpp:224-451).  Large grids (>= 2^20 cells, or SB2_JIT=1; SB2_JIT=0 turns it off,
 * SB2_JIT=require makes a failed build an error) get, at sb2_finalize, a build of the stage kernel in which the
 * model's token records are expanded into straight-line code, compiled for sm_100a with the CUDA run-time compiler
 * (libnvrtc, loaded with dlopen).  Results are bit-identical to the interpreter kernel, which keeps running when
 * the run-time compiler is not available.
 * sb2_specialized: 1 when the handle runs the specialised kernel.  sb2_specialize_info: text describing what
 * happened (library used, build time, or why the interpreter runs); returns the full length. */
SB2_API int sb2_specialized(const sb2_sim* sim);
/* Smaller grids get the same build from a background thread once the simulator has taken 256 steps (a short-lived
 * simulator never pays for it); the kernel is swapped in at a step boundary, results unchanged.  SB2_JIT=background
 * starts that build at sb2_finalize.  sb2_specialize_now: build (or wait for the running build) and install the kernel
 * before returning; 1 when the specialised kernel is in place, 0 when the interpreter kernel stays (see
 * sb2_specialize_info), negative on misuse. */
SB2_API int sb2_specialize_now(sb2_sim* sim);
SB2_API int sb2_specialize_info(const sb2_sim* sim, char* buf, int buflen);
/* Builds the specialised kernel for a list of n token records (code = opcode * SB2_MAX_DEPTH + slot, arg) without
 * touching a GPU and writes the sm_100a cubin to cubin_path (may be NULL): build check and tests.  ns / np / w /
 * depth / minb / d3: species, cell pairs per thread, warps per CTA, stack depth, CTAs per SM, 3-D grid.
 * The compiler log (or the reason of a failure) is copied into log.  Returns 0 on success. */
SB2_API int sb2_specialize_dry_run(const uint32_t* code, const uint32_t* arg, int n, int ns, int np, int w, int depth,
                                   int minb, int d3, const char* cubin_path, char* log, int loglen);

#ifdef __cplusplus
}
#endif
#endif /* SPATIAL_B200_H_ */
