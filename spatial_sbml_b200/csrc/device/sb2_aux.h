// sb2_aux.h — the kernels around the fused stage kernel: boundary-condition lists, CIP-CSLR
// advection, membrane transport, min-dt reduction (implemented in aux_kernels.cu).
#ifndef SB2_AUX_H_
#define SB2_AUX_H_

#include "sb2_internal.h"

// ---------------- boundary conditions (calcBoundary, calcPDE.cpp:871-1034) ----------------
struct Sb2BcSpecies {
  int geo;
  bool has_bc;
  int kind[6], skind[6], src[6];
};

struct Sb2BoundaryLists {
  int n_groups;      // distinct (species, cell) targets
  int n_entries;
  int n_neumann, n_dirichlet;
  bool any_neumann_nonzero_possible;
  // device arrays
  int32_t* d_group_start;   // [n_groups + 1]
  int32_t* d_group_species; // [n_groups]
  int64_t* d_entry_cell;    // [n_entries] padded index written (k or u)
  int64_t* d_entry_valcell; // [n_entries] padded index a field-valued BC is read at
  int32_t* d_entry_side;    // [n_entries] SB2_XMAX..
  // per (species, side) descriptors, host side
  int kind[SB2_MAX_SPECIES][6], skind[SB2_MAX_SPECIES][6], src[SB2_MAX_SPECIES][6];
  double delta[3];
};

// out3[0]: every owned row; on a slab also out3[1]: the edge_w owned rows next to each neighbouring slab, out3[2]: the others
int sb2_build_boundary_lists(const sb2_grid& g, int64_t P, int64_t PL, int64_t first,
                             const std::vector<std::vector<uint8_t> >& h_flags, const std::vector<Sb2BcSpecies>& bcs,
                             const double* consts, cudaStream_t stream, int edge_w, Sb2BoundaryLists* out3, std::string& err);
void sb2_free_boundary_lists(Sb2BoundaryLists& l);
// true when some Neumann side currently carries a value that is not exactly zero.  whole_grid: a handle that holds the whole
// grid may skip the pass when no cell receives a flux; a slab may not (the slabs of a grid must agree on the schedule).
bool sb2_neumann_active(const Sb2BoundaryLists& l, int nspecies, const double* consts, bool whole_grid);
cudaError_t sb2_apply_boundary(const Sb2BoundaryLists& l, double* const* k, double* const* u, const double* consts,
                               double* const* fields, bool do_neumann, bool do_dirichlet, cudaStream_t stream, int* launches);

// ---------------- CIP-CSLR advection (cipCSLR, calcPDE.cpp:564-869) ----------------
struct Sb2CipArgs {
  int32_t nx, ny, nz, dim;
  int64_t P, PL, first, nalloc;
  double delta[3];
  double dt;
  double* u;          // cell averages (in place)
  double* faces[3];   // face point values: faces[a][cell] = face between cell and cell + e_a
  int32_t akind[3];
  double aconst[3];
  const double* afield[3];       // field-valued velocity at the cells
  const double* afield_face[3];  // ... and at the plus faces along the same axis (the odd doubled-grid points between two cells)
  const uint8_t* flags;
  double* tmp_cell;   // scratch: cell increments
  double* tmp_face;   // scratch: face increments of the pass direction
  // 2-D fused pass (NULL: the two-kernel pass): the new cell averages and the new faces of the pass axis are written here,
  // the caller swaps them with u / faces[axis] afterwards
  double* u_out;
  double* face_out;
  // rows (2-D) / planes (3-D) of the slab axis: the update sweep runs on [row_begin, row_end), the flux loop on
  // [row_begin, flux_row_end) (one more row where a neighbouring slab's first row feeds a face of the last owned row)
  int32_t row_begin, row_end, flux_row_end;
};
// one direction pass (flux loop + update sweep) along `axis`
cudaError_t sb2_launch_cip_pass(const Sb2CipArgs& a, int axis, cudaStream_t stream, int* launches);

// ---------------- membrane transport (calcMemTransport, calcPDE.cpp:1036-1482) ----------------
struct Sb2MemTransport {
  int npts, ntoks, n_groups, n_entries, nspecies;
  // program (host copy is passed by value to the kernel)
  std::vector<uint32_t> tok;        // kind | op<<8 | slot<<16
  std::vector<int32_t> tok_operand; // per token: operand-table row (VAR / FIELD tokens), else -1
  uint32_t* d_tok;    // device copy of tok
  int32_t* d_operand; // device copy of tok_operand
  int32_t* d_near;    // [n_operand_rows][npts] padded cell index or -1
  int32_t* d_far;     // [n_operand_rows][npts]
  double* d_fieldval; // [n_operand_rows][npts] value of FIELD operands at the membrane point
  double* d_rate;     // [npts]
  // gather lists: contributions grouped by (species, target cell), in the reference's order
  int32_t* d_group_start;   // [n_groups + 1]
  int32_t* d_group_species; // [n_groups]
  int64_t* d_group_cell;    // [n_groups]
  int32_t* d_entry_pt;      // [n_entries]
  double* d_entry_coef;     // [n_entries] fabs(n)*stoich with the sign of reactant(-)/product(+) folded in separately
  double* d_entry_delta;    // [n_entries] deltaX/Y/Z of the point's axis
  int32_t* d_entry_sign;    // [n_entries] -1 reactant, +1 product
  // host copies of the group keys and, per group, whether no transport declared earlier feeds the same (species, cell): such a
  // group starts from zero instead of from what the buffer holds (sb2_finalize fills it in; see launch_membrane_ops)
  std::vector<int32_t> h_group_species;
  std::vector<int64_t> h_group_cell;
  uint8_t* d_group_first;   // [n_groups] or NULL
};
int sb2_build_transport(const sb2_grid& g, int64_t P, int64_t PL, int64_t first, int nspecies, const sb2_token* toks,
                        int ntoks, const sb2_ref* refs, int nrefs, int n_reactants, const sb2_mem_point* pts, int npts,
                        const int32_t* operand_cells, const double* operand_values, const int32_t* target_cells,
                        cudaStream_t stream, Sb2MemTransport& out, std::string& err);
void sb2_free_transport(Sb2MemTransport& t);
// mu / mkprev / mkout: the membrane species (binding reactions), indexed by membrane species id; may be NULL
// first_resets: groups of volume species flagged in d_group_first start from zero (kout is then the persistent transport
// buffer of the species, which needs no clearing between stages)
cudaError_t sb2_launch_transport(const Sb2MemTransport& t, const sb2_grid& g, const double* const* u,
                                 const double* const* kprev, double* const* kout, const double* const* mu,
                                 const double* const* mkprev, double* const* mkout, const double* consts, int m,
                                 double cdt, bool first_resets, cudaStream_t stream, int* launches);
cudaError_t sb2_transport_set_first(Sb2MemTransport& t, const std::vector<uint8_t>& first, cudaStream_t stream);

// ---------------- stability (checkStability.cpp:10-52, 119-154) ----------------
struct Sb2MinDtArgs {
  int32_t nx, ny, nz, dim;
  int64_t P, PL, first;
  double delta[3];
  double dt;
  const uint8_t* flags;
  int32_t has_diff, has_adv;
  int32_t dkind[3], akind[3];
  double dconst[3], aconst[3];
  const double* dfield[3];
  const double* afield[3];
};
cudaError_t sb2_launch_min_dt(const Sb2MinDtArgs& a, double* d_scratch, cudaStream_t stream, int* launches, double* result);

#endif  // SB2_AUX_H_
