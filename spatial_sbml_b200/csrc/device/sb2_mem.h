// sb2_mem.h — species that live ON a membrane (sparse point sets) and the generic per-point / per-cell expression
// evaluator (implemented in mem_kernels.cu): membrane reactions (reversePolishRK over a membrane's domainIndex), surface
// diffusion (calcMemDiffusion), the RK4 update of membrane species, the pseudo-membrane fill, checkMemDiffusionStab, and
// assignment rules re-evaluated on the device after every step (reversePolishInitial).
#ifndef SB2_MEM_H_
#define SB2_MEM_H_

#include "sb2_internal.h"

struct Sb2PointSet {
  int npts;
  // pseudo-membrane fill: value[target] = min(value[a], value[b])
  int n_fill;
  int32_t *d_fill_target, *d_fill_a, *d_fill_b;
};

struct Sb2MemSpecies {
  int set;
  double* u;
  double* k[4];
  // surface diffusion
  bool has_diff;
  int dkind, dsrc;
  int n_diff_pts, n_terms;
  int32_t *d_diff_pts, *d_term_start, *d_adj;
  double *d_s, *d_d, *d_area, *d_dvalues;
  // update list (domainIndex of the species' geometry)
  int n_update;
  int32_t* d_update_pts;
};

// one membrane reaction / rate rule
struct Sb2MemReaction {
  int set, kind, n_reactants, npts, ntoks, nrefs;
  std::vector<uint32_t> tok;        // kind | op << 8 | slot << 16
  uint32_t* d_tok;
  int32_t* d_operand;  // per token: row of d_opvalues (FIELD / VAR tokens), else -1
  int32_t* d_points;   // [npts] storage index
  double* d_opvalues;  // [rows][npts]
  int ref_species[8];  // membrane species id
  int ref_sign[8];     // -1 reactant, +1 product / rate-rule target
  double ref_stoich[8];
};

struct Sb2MemView {  // what the kernels see of the membrane species of a simulator
  double* u[SB2_MAX_MEM_SPECIES];
  double* k[SB2_MAX_MEM_SPECIES][4];
};

int sb2_build_mem_reaction(int set, int kind, const sb2_token* toks, int ntoks, const sb2_ref* refs, int nrefs, int n_reactants,
                           const int32_t* points, int npts, const double* operand_values, int n_mem_species, int set_npts,
                           cudaStream_t stream, Sb2MemReaction& out, std::string& err);
void sb2_free_mem_reaction(Sb2MemReaction& r);
cudaError_t sb2_launch_mem_reaction(const Sb2MemReaction& r, const Sb2MemView& v, const double* consts, int m, double cdt,
                                    cudaStream_t stream, int* launches);
cudaError_t sb2_launch_mem_diffusion(const Sb2MemSpecies& s, int id, const Sb2MemView& v, const double* consts, int m, double cdt,
                                     int dim, cudaStream_t stream, int* launches);
cudaError_t sb2_launch_mem_update(const Sb2MemSpecies& s, int id, const Sb2MemView& v, double dt, cudaStream_t stream, int* launches);
cudaError_t sb2_launch_mem_fill(const Sb2PointSet& set, double* u, cudaStream_t stream, int* launches);
// checkMemDiffusionStab (checkStability.cpp:54-117): smallest d_ij^2 / (2 D) that dt reaches, else dt
cudaError_t sb2_launch_mem_min_dt(const Sb2MemSpecies& s, const double* consts, double dt, double* d_scratch, cudaStream_t stream,
                                  int* launches, double* result);

// ---- expression evaluated on every cell of the compact grid (assignment rules, spatialsimulator.cpp:1566-1581) ----
struct Sb2CellRule {
  int ntoks;
  std::vector<uint32_t> tok;
  uint32_t* d_tok;
  int target_kind;   // 0: read-only field id, 1: staged species id
  int target;
  int geo;           // >= 0: only the cells of this geometry (species targets), -1: every cell
};
struct Sb2CellRuleArgs {
  int32_t nx, ny, nz;
  int64_t P, PL, first;
  int32_t ntoks;
  const uint32_t* tok;
  double* out;
  const uint8_t* mask;  // flag bytes of the geometry, or NULL
  const double* species[SB2_MAX_SPECIES];
  const double* fields[SB2_MAX_FIELDS];
  double consts[SB2_MAX_CONSTS];
};
int sb2_build_cell_rule(const sb2_token* toks, int ntoks, int target_kind, int target, int geo, cudaStream_t stream, Sb2CellRule& out,
                        std::string& err);
void sb2_free_cell_rule(Sb2CellRule& r);
cudaError_t sb2_launch_cell_rule(const Sb2CellRuleArgs& a, cudaStream_t stream, int* launches);

#endif  // SB2_MEM_H_
