// aux_kernels.cu — kernels around the fused stage kernel (sm_100a): boundary-condition lists,
// CIP-CSLR advection, membrane transport, min-dt reduction.  This translation unit is compiled
// with -fmad=false so that plain C expressions keep the reference's operation order and rounding.
#include "sb2_aux.h"
#include <algorithm>
#include <cstdlib>
#include <thread>
#include "sb2_ops.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace {

template <typename T>
cudaError_t to_device(T** dptr, const std::vector<T>& h, cudaStream_t stream) {
  *dptr = NULL;
  if (h.empty()) return cudaSuccess;
  cudaError_t e = cudaMalloc((void**)dptr, sizeof(T) * h.size());
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(stream);
}

}  // namespace

// =============================================================================================
// boundary conditions
// =============================================================================================
namespace {

struct BcEntry {
  int species;
  int64_t cell, valcell;
  int side;
  int seq;
  int slabc;  // row (2-D) / plane (3-D) of the written cell along the slab axis
};

struct BcKernelArgs {
  int n_groups;
  const int32_t* group_start;
  const int32_t* group_species;
  const int64_t* entry_cell;
  const int64_t* entry_valcell;
  const int32_t* entry_side;
  int kind[SB2_MAX_SPECIES][6];
  int skind[SB2_MAX_SPECIES][6];
  int src[SB2_MAX_SPECIES][6];
  double cval[SB2_MAX_SPECIES][6];
  double delta[3];
  double* k[SB2_MAX_SPECIES];
  double* u[SB2_MAX_SPECIES];
  const double* fields[SB2_MAX_FIELDS];
  int do_neumann, do_dirichlet;
};

__global__ void __launch_bounds__(128) bc_kernel(const __grid_constant__ BcKernelArgs A) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= A.n_groups) return;
  const int s = A.group_species[g];
  for (int e = A.group_start[g]; e < A.group_start[g + 1]; ++e) {
    const int side = A.entry_side[e];
    const int kind = A.kind[s][side];
    const double v = (A.skind[s][side] == SB2_SRC_FIELD) ? A.fields[A.src[s][side]][A.entry_valcell[e]] : A.cval[s][side];
    const int64_t c = A.entry_cell[e];
    if (kind == SB2_BC_NEUMANN) {
      // delta += -2.0 * (-bc) / deltaX   (calcPDE.cpp:994)
      if (A.do_neumann) A.k[s][c] = A.k[s][c] + (-2.0 * (-v)) / A.delta[side >> 1];
    } else if (kind == SB2_BC_DIRICHLET) {
      if (A.do_dirichlet) A.u[s][c] = v;  // calcPDE.cpp:995
    }
  }
}

}  // namespace

// out[0]: every owned row; for a slab also out[1]: the edge_w owned rows next to each neighbouring slab, out[2]: the other
// owned rows (the boundary pass of a sharded step is split like its stage kernels)
int sb2_build_boundary_lists(const sb2_grid& g, int64_t P, int64_t PL, int64_t first,
                             const std::vector<std::vector<uint8_t> >& h_flags, const std::vector<Sb2BcSpecies>& bcs,
                             const double* consts, cudaStream_t stream, int edge_w, Sb2BoundaryLists* out3, std::string& err) {
  (void)consts;
  Sb2BoundaryLists& out = out3[0];
  for (int part = 0; part < 3; ++part) {
    memset(&out3[part], 0, sizeof(out3[part]));
    out3[part].delta[0] = g.dx; out3[part].delta[1] = g.dy; out3[part].delta[2] = g.dz;
  }
  std::vector<BcEntry> entries;
  const int nx = g.nx, ny = g.ny, nz = g.nz;
  auto pad = [&](int i, int j, int k) { return first + (int64_t)k * PL + (int64_t)j * P + i; };
  // slab: only owned rows / planes receive entries; the reference's row-0 quirks refer to the whole grid
  const bool whole = g.own_hi <= g.own_lo;
  const int n_held = g.dim == 3 ? nz : ny;
  const int own_lo = whole ? 0 : g.own_lo, own_hi = whole ? n_held : g.own_hi;
  const int gj0 = g.dim == 2 ? g.row0 : 0, gk0 = g.dim == 3 ? g.row0 : 0;
  for (size_t s = 0; s < bcs.size(); ++s) {
    for (int part = 0; part < 3; ++part)
      for (int k = 0; k < 6; ++k) { out3[part].kind[s][k] = bcs[s].kind[k]; out3[part].skind[s][k] = bcs[s].skind[k]; out3[part].src[s][k] = bcs[s].src[k]; }
    if (!bcs[s].has_bc) continue;
    const std::vector<uint8_t>& fl = h_flags[bcs[s].geo];
    auto dom = [&](int i, int j, int k) -> bool {
      if (i < 0 || j < 0 || k < 0 || i >= nx || j >= ny || k >= nz) return false;
      return fl[((int64_t)k * ny + j) * nx + i] & SB2_F_DOMAIN;
    };
    // The loop of calcPDE.cpp:958-1033 over every even point c, restated on compact coordinates.
    // Flat-index wrap-around of the reference (X±2, X±4 leaving the row) always lands on an odd
    // row / plane, which no volume domain contains, so it reads as "not in domain" (SURVEY §8 a8').
    // Rows are scanned by several host threads on large grids; the pieces are concatenated in row order, so the
    // entries (and their sequence numbers) come out exactly as from one serial scan.
    auto scan = [&](int64_t row_begin, int64_t row_end, std::vector<BcEntry>& local) {
      auto add = [&](int side, int ai, int aj, int ak, int vi, int vj, int vk) {
        if (bcs[s].kind[side] == SB2_BC_NONE) return;
        const int slabc = g.dim == 3 ? ak : aj;
        if (slabc < own_lo || slabc >= own_hi) return;
        BcEntry e;
        e.species = (int)s; e.cell = pad(ai, aj, ak); e.valcell = pad(vi, vj, vk); e.side = side; e.seq = 0; e.slabc = slabc;
        local.push_back(e);
      };
      for (int64_t row = row_begin; row < row_end; ++row) {
        const int k = (int)(row / ny), j = (int)(row % ny);
        for (int i = 0; i < nx; ++i) {
          // X: p = c+1 gets the Xmax condition when its outer neighbour is outside
          if (dom(i + 1, j, k) && !dom(i + 2, j, k)) add(SB2_XMAX, i + 1, j, k, i + 1, j, k);
          // Xm is gated by desc.Y.mOutside (calcPDE.cpp:997): skipped on row Y=0 of plane Z=0
          if (!(k + gk0 == 0 && j + gj0 == 0) && dom(i - 1, j, k) && !dom(i - 2, j, k)) add(SB2_XMIN, i - 1, j, k, i - 1, j, k);
          if (g.dim >= 2) {
            if (dom(i, j + 1, k) && !dom(i, j + 2, k)) add(SB2_YMAX, i, j + 1, k, i, j + 1, k);
            if (dom(i, j - 1, k) && !dom(i, j - 2, k)) add(SB2_YMIN, i, j - 1, k, i, j - 1, k);
          }
          if (g.dim >= 3) {
            // coordinateDesc::set stores Z.p = xp and Z.m = xm (mystruct.h:186-187): the Z tests look
            // at the X neighbour of c but at the Z±4 point for "outside"
            if (dom(i + 1, j, k) && !dom(i, j, k + 2)) add(SB2_ZMAX, i + 1, j, k, i + 1, j, k);
            if (!(i == 0 && j == 0 && k + gk0 == 0) && dom(i - 1, j, k) && !dom(i, j, k - 2)) {
              if (bcs[s].kind[SB2_ZMIN] == SB2_BC_DIRICHLET) {
                // value[Z.mm] = bc (calcPDE.cpp:1028); out of the array when k < 2 → skipped
                if (k - 2 >= 0) add(SB2_ZMIN, i, j, k - 2, i - 1, j, k);
              } else {
                add(SB2_ZMIN, i - 1, j, k, i - 1, j, k);
              }
            }
          }
        }
      }
    };
    const int64_t rows = (int64_t)nz * ny;
    unsigned nthreads = std::thread::hardware_concurrency();
    if (const char* e = getenv("SSB_HOST_THREADS")) nthreads = (unsigned)atoi(e);
    if (nthreads > 32) nthreads = 32;
    const int64_t par_min = getenv("SSB_HOST_PAR_MIN") ? atoll(getenv("SSB_HOST_PAR_MIN")) : ((int64_t)1 << 18);  // tests lower it
    if (nthreads < 1 || rows * nx < par_min) nthreads = 1;
    std::vector<std::vector<BcEntry> > pieces(nthreads);
    if (nthreads == 1) scan(0, rows, pieces[0]);
    else {
      std::vector<std::thread> pool;
      const int64_t chunk = (rows + nthreads - 1) / nthreads;
      for (unsigned t = 0; t < nthreads; ++t) {
        const int64_t b = (int64_t)t * chunk, e = std::min<int64_t>(rows, b + chunk);
        if (b >= e) break;
        pool.emplace_back([&scan, &pieces, t, b, e] { scan(b, e, pieces[t]); });
      }
      for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
    }
    int seq = 0;
    for (unsigned t = 0; t < nthreads; ++t)
      for (size_t q = 0; q < pieces[t].size(); ++q) {
        pieces[t][q].seq = seq++;
        entries.push_back(pieces[t][q]);
      }
  }
  std::stable_sort(entries.begin(), entries.end(), [](const BcEntry& a, const BcEntry& b) {
    if (a.species != b.species) return a.species < b.species;
    return a.cell < b.cell;
  });
  const bool lo_nb = !whole && own_lo > 0, hi_nb = !whole && own_hi < n_held;
  for (int part = 0; part < (whole ? 1 : 3); ++part) {
    Sb2BoundaryLists& o = out3[part];
    std::vector<int32_t> gstart, gspecies, eside;
    std::vector<int64_t> ecell, evalcell;
    int last = -1;
    for (size_t i = 0; i < entries.size(); ++i) {
      const bool edge = (lo_nb && entries[i].slabc < own_lo + edge_w) || (hi_nb && entries[i].slabc >= own_hi - edge_w);
      if ((part == 1 && !edge) || (part == 2 && edge)) continue;
      if (last < 0 || entries[i].species != entries[last].species || entries[i].cell != entries[last].cell) {
        gstart.push_back((int32_t)ecell.size());
        gspecies.push_back(entries[i].species);
      }
      last = (int)i;
      ecell.push_back(entries[i].cell);
      evalcell.push_back(entries[i].valcell);
      eside.push_back(entries[i].side);
      const int kind = bcs[entries[i].species].kind[entries[i].side];
      if (kind == SB2_BC_NEUMANN) o.n_neumann++;
      if (kind == SB2_BC_DIRICHLET) o.n_dirichlet++;
    }
    gstart.push_back((int32_t)ecell.size());
    o.n_groups = (int)gspecies.size();
    o.n_entries = (int)ecell.size();
    cudaError_t e;
    if ((e = to_device(&o.d_group_start, gstart, stream)) != cudaSuccess ||
        (e = to_device(&o.d_group_species, gspecies, stream)) != cudaSuccess ||
        (e = to_device(&o.d_entry_cell, ecell, stream)) != cudaSuccess ||
        (e = to_device(&o.d_entry_valcell, evalcell, stream)) != cudaSuccess ||
        (e = to_device(&o.d_entry_side, eside, stream)) != cudaSuccess) {
      err = std::string("boundary lists: ") + cudaGetErrorString(e);
      return SB2_ERR_CUDA;
    }
  }
  (void)out;
  return SB2_OK;
}

void sb2_free_boundary_lists(Sb2BoundaryLists& l) {
  cudaFree(l.d_group_start); cudaFree(l.d_group_species); cudaFree(l.d_entry_cell);
  cudaFree(l.d_entry_valcell); cudaFree(l.d_entry_side);
  memset(&l, 0, sizeof(l));
}

bool sb2_neumann_active(const Sb2BoundaryLists& l, int nspecies, const double* consts, bool whole_grid) {
  if (whole_grid && l.n_neumann == 0) return false;
  for (int s = 0; s < nspecies; ++s)
    for (int k = 0; k < 6; ++k) {
      if (l.kind[s][k] != SB2_BC_NEUMANN) continue;
      if (l.skind[s][k] == SB2_SRC_FIELD) return true;
      if (l.skind[s][k] == SB2_SRC_CONST && consts[l.src[s][k]] != 0.0) return true;
    }
  return false;
}

cudaError_t sb2_apply_boundary(const Sb2BoundaryLists& l, double* const* k, double* const* u, const double* consts,
                               double* const* fields, bool do_neumann, bool do_dirichlet, cudaStream_t stream, int* launches) {
  if (l.n_groups == 0) return cudaSuccess;
  if (!(do_neumann && l.n_neumann > 0) && !(do_dirichlet && l.n_dirichlet > 0)) return cudaSuccess;
  BcKernelArgs a;
  memset(&a, 0, sizeof(a));
  a.n_groups = l.n_groups;
  a.group_start = l.d_group_start; a.group_species = l.d_group_species;
  a.entry_cell = l.d_entry_cell; a.entry_valcell = l.d_entry_valcell; a.entry_side = l.d_entry_side;
  for (int s = 0; s < SB2_MAX_SPECIES; ++s) {
    for (int j = 0; j < 6; ++j) {
      a.kind[s][j] = l.kind[s][j]; a.skind[s][j] = l.skind[s][j]; a.src[s][j] = l.src[s][j];
      a.cval[s][j] = (l.skind[s][j] == SB2_SRC_CONST) ? consts[l.src[s][j]] : 0.0;
    }
  }
  for (int d = 0; d < 3; ++d) a.delta[d] = l.delta[d];
  // only species that own groups are dereferenced
  for (int s = 0; s < SB2_MAX_SPECIES; ++s) { a.k[s] = NULL; a.u[s] = NULL; }
  // callers pass arrays sized to the species count; copy what the kinds table says exists
  for (int s = 0; s < SB2_MAX_SPECIES; ++s) {
    bool used = false;
    for (int j = 0; j < 6; ++j) if (l.kind[s][j] != SB2_BC_NONE) used = true;
    if (used) { a.k[s] = k[s]; a.u[s] = u[s]; }
  }
  for (int f = 0; f < SB2_MAX_FIELDS; ++f) a.fields[f] = NULL;
  for (int s = 0; s < SB2_MAX_SPECIES; ++s)
    for (int j = 0; j < 6; ++j)
      if (l.kind[s][j] != SB2_BC_NONE && l.skind[s][j] == SB2_SRC_FIELD) a.fields[l.src[s][j]] = fields[l.src[s][j]];
  a.do_neumann = do_neumann ? 1 : 0;
  a.do_dirichlet = do_dirichlet ? 1 : 0;
  if (launches) ++*launches;
  bc_kernel<<<(l.n_groups + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

// =============================================================================================
// CIP-CSLR advection
// =============================================================================================
namespace {

__device__ __forceinline__ bool cip_dom(const Sb2CipArgs& A, int64_t idx, int c, int n, int off, int64_t st) {
  const int cc = c + off;
  if (cc < 0 || cc >= n) return false;
  return A.flags[idx + off * st] & SB2_F_DOMAIN;
}

// One cell of the flux loop of a direction pass (calcPDE.cpp:581-651 for X; the Y and Z passes are identical up to the axis):
// the increment of the cell average dc = (g_in - g_out) / delta and, when the cell has a plus face inside the domain, the
// increment of that face's point value.
__device__ __forceinline__ void cip_cell(const Sb2CipArgs& A, const int axis, const int64_t idx, const int c, const int n, const int64_t st,
                                         const uint32_t f, double& dc, double& newface, bool& has_face) {
  const double* val = A.u;
  const double* face = A.faces[axis];
  const double delta = A.delta[axis];
  const double dt = A.dt;
  const bool d_p1 = cip_dom(A, idx, c, n, 1, st), d_p2 = cip_dom(A, idx, c, n, 2, st);
  const bool d_m1 = cip_dom(A, idx, c, n, -1, st), d_m2 = cip_dom(A, idx, c, n, -2, st);
  const double v0 = val[idx];
  // index collapsing of calcPDE.cpp:596-611
  const double vp3 = d_p2 ? face[idx + st] : (d_p1 ? val[idx + st] : v0);
  const double vm3 = d_m2 ? face[idx - 2 * st] : (d_m1 ? val[idx - st] : v0);
  const double vp1 = d_p1 ? face[idx] : v0;
  const double vp2 = d_p1 ? val[idx + st] : v0;
  const double vm1 = d_m1 ? face[idx - st] : v0;
  const double vm2 = d_m1 ? val[idx - st] : v0;

  // uniform velocity, or the mean of the cell's value and its (collapsed) plus neighbour's (calcPDE.cpp:612-613)
  const bool ufield = A.akind[axis] == SB2_SRC_FIELD;
  const double ux = ufield ? (A.afield[axis][idx] + A.afield[axis][d_p1 ? idx + st : idx]) / 2.0 : A.aconst[axis];
  const double Dx = -copysign(1.0, ux) * delta;
  const double xi = ux * dt;
  const bool bofp = f & (SB2_F_XP << (2 * axis));
  const bool bofm = f & (SB2_F_XM << (2 * axis));
  double g_in = 0.0, g_out = 0.0;
  has_face = false;
  newface = 0.0;
  if (!bofp) {
    double a_i, b_i, beta_i;
    if (ux >= 0) {
      a_i = vp1;
      beta_i = ((fabs(vp1 - v0) + DBL_EPSILON) / (fabs(v0 - vm1) + DBL_EPSILON) - 1.0) / Dx;
      b_i = ((1.0 + beta_i * Dx) * v0 - vp1) / Dx;
    } else {
      a_i = vp1;
      beta_i = ((fabs(vp1 - vp2) + DBL_EPSILON) / (fabs(vp2 - vp3) + DBL_EPSILON) - 1.0) / Dx;
      b_i = ((1.0 + beta_i * Dx) * vp2 - vp1) / Dx;
    }
    const double mxi = -xi;
    const double den = 1.0 + beta_i * mxi;
    double nf = (a_i + 2 * b_i * mxi + beta_i * b_i * (mxi * mxi)) / (den * den) - vp1;
    if (ufield) {
      // correction due to velocity divergence (calcPDE.cpp:623-625): the velocity at Xplus3 and Xminus1, i.e. at the face of the
      // next cell and of the previous cell, each collapsed like the value indices
      const double u3 = d_p2 ? A.afield_face[axis][idx + st] : (d_p1 ? A.afield[axis][idx + st] : A.afield[axis][idx]);
      const double um1 = d_m1 ? A.afield_face[axis][idx - st] : A.afield[axis][idx];
      nf = nf + ((dt / (2 * delta)) * (vp1 + nf)) * (u3 - um1);
    }
    if (d_p1) { has_face = true; newface = nf; }
    const double G = (-a_i * xi + b_i * (xi * xi)) / (-1.0 + beta_i * xi);
    if (ux >= 0) g_out = G; else g_in = -G;
  }
  if (!bofm) {
    double a_i2, b_i2, beta_i2;
    if (ux >= 0) {
      a_i2 = vm1;
      beta_i2 = ((fabs(vm1 - vm2) + DBL_EPSILON) / (fabs(vm2 - vm3) + DBL_EPSILON) - 1.0) / Dx;
      b_i2 = ((1.0 + beta_i2 * Dx) * vm2 - vm1) / Dx;
      g_in = (-a_i2 * xi + b_i2 * (xi * xi)) / (-1.0 + beta_i2 * xi);
    } else {
      a_i2 = vm1;
      beta_i2 = ((fabs(vm1 - v0) + DBL_EPSILON) / (fabs(v0 - vp1) + DBL_EPSILON) - 1.0) / Dx;
      b_i2 = ((1.0 + beta_i2 * Dx) * v0 - vm1) / Dx;
      g_out = -(-a_i2 * xi + b_i2 * (xi * xi)) / (-1.0 + beta_i2 * xi);
    }
  }
  dc = (g_in - g_out) / delta;
}

// The plus-face half of cip_cell for a UNIFORM velocity: the flux through the face between the cell and its plus neighbour,
// F = (-a xi + b xi^2) / (-1 + beta xi), and the increment of the face's point value.  With a uniform velocity the minus-face
// flux of a cell is, operand for operand, the plus-face flux of its minus neighbour (g_in(i) = F(i-1) for u >= 0,
// g_out(i) = -F(i-1) for u < 0, whenever the cell is not a minus-boundary cell), so each face is evaluated once.
__device__ __forceinline__ void cip_plus_face(const Sb2CipArgs& A, const int axis, const int64_t idx, const int c, const int n, const int64_t st,
                                              const uint32_t f, double& F, double& newface, bool& has_face) {
  F = 0.0; newface = 0.0; has_face = false;
  if (f & (SB2_F_XP << (2 * axis))) return;
  const double* val = A.u;
  const double* face = A.faces[axis];
  const double delta = A.delta[axis];
  const bool d_p1 = cip_dom(A, idx, c, n, 1, st), d_p2 = cip_dom(A, idx, c, n, 2, st);
  const bool d_m1 = cip_dom(A, idx, c, n, -1, st);
  const double v0 = val[idx];
  const double vp1 = d_p1 ? face[idx] : v0;
  const double ux = A.aconst[axis];
  const double Dx = -copysign(1.0, ux) * delta;
  const double xi = ux * A.dt;
  double a_i, b_i, beta_i;
  if (ux >= 0) {
    const double vm1 = d_m1 ? face[idx - st] : v0;
    a_i = vp1;
    beta_i = ((fabs(vp1 - v0) + DBL_EPSILON) / (fabs(v0 - vm1) + DBL_EPSILON) - 1.0) / Dx;
    b_i = ((1.0 + beta_i * Dx) * v0 - vp1) / Dx;
  } else {
    const double vp3 = d_p2 ? face[idx + st] : (d_p1 ? val[idx + st] : v0);
    const double vp2 = d_p1 ? val[idx + st] : v0;
    a_i = vp1;
    beta_i = ((fabs(vp1 - vp2) + DBL_EPSILON) / (fabs(vp2 - vp3) + DBL_EPSILON) - 1.0) / Dx;
    b_i = ((1.0 + beta_i * Dx) * vp2 - vp1) / Dx;
  }
  const double mxi = -xi;
  const double den = 1.0 + beta_i * mxi;
  const double nf = (a_i + 2 * b_i * mxi + beta_i * b_i * (mxi * mxi)) / (den * den) - vp1;
  if (d_p1) { has_face = true; newface = nf; }
  F = (-a_i * xi + b_i * (xi * xi)) / (-1.0 + beta_i * xi);
}
// increment of the cell average from the fluxes through its two faces (Fm: the plus-face flux of the minus neighbour)
__device__ __forceinline__ double cip_dc(const Sb2CipArgs& A, const int axis, const uint32_t f, const double Fp, const double Fm) {
  const double ux = A.aconst[axis];
  const bool bofp = f & (SB2_F_XP << (2 * axis)), bofm = f & (SB2_F_XM << (2 * axis));
  double g_in = 0.0, g_out = 0.0;
  if (ux >= 0) { if (!bofp) g_out = Fp; if (!bofm) g_in = Fm; }
  else { if (!bofp) g_in = -Fp; if (!bofm) g_out = -Fm; }
  return (g_in - g_out) / A.delta[axis];
}

// flux loop of one direction pass, one cell per thread (3-D grids; 2-D grids run the fused pass below)
__global__ void __launch_bounds__(256) cip_flux_kernel(const __grid_constant__ Sb2CipArgs A, int axis) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y + (A.dim == 2 ? A.row_begin : 0), k = blockIdx.z + (A.dim == 3 ? A.row_begin : 0);
  if (i >= A.nx) return;
  const int64_t idx = A.first + (int64_t)k * A.PL + (int64_t)j * A.P + i;
  const uint32_t f = A.flags[idx];
  if (!(f & SB2_F_DOMAIN)) return;
  const int c = axis == 0 ? i : axis == 1 ? j : k;
  const int n = axis == 0 ? A.nx : axis == 1 ? A.ny : A.nz;
  const int64_t st = axis == 0 ? 1 : axis == 1 ? A.P : A.PL;
  double dc, nf;
  bool has_face;
  cip_cell(A, axis, idx, c, n, st, f, dc, nf, has_face);
  if (has_face) A.tmp_face[idx] = nf;
  A.tmp_cell[idx] = dc;
}

// ---------------------------------------------------------------------------------------------
// 2-D: one kernel per direction pass.  A thread owns a column of a strip and marches over the rows of its block: flux loop,
// update sweep and the advance of the other axis' faces in one go, out of place for u and for the faces of the pass axis
// (the flux loop of a neighbouring cell must still see the old values; the caller swaps the buffers), in place for the
// faces of the other axis (only their own cell touches them).  Per cell and pass: read u, the axis faces and a flag byte,
// write u and the axis faces, read-modify-write the other faces — 49 bytes, against two full-array memsets, two kernels
// and ~140 bytes before.  The increment of the cell across the other axis, which the face between them needs, comes from
// the previous row of the march (X pass) or from the neighbouring thread through shared memory (Y pass).
constexpr int kCipThreads = 128;
constexpr int kCipRows = 32;

__global__ void __launch_bounds__(kCipThreads) cip2d_x_kernel(const __grid_constant__ Sb2CipArgs A) {
  // strips overlap by one column on the left: thread 0 only supplies the face flux its right neighbour needs
  const int i = (int)blockIdx.x * (kCipThreads - 1) + (int)threadIdx.x - 1;
  const int jb = A.row_begin + (int)blockIdx.y * kCipRows;
  const int je = min(jb + kCipRows, A.row_end);
  // one more row (flux only) where the face above the block's last row needs the increment of the row beyond
  const int jl = je < A.flux_row_end ? je + 1 : je;
  const bool inside = i >= 0 && i < A.nx;
  const bool owner = inside && threadIdx.x > 0;
  const bool shared_flux = A.akind[0] == SB2_SRC_CONST;  // uniform velocity: every face flux is evaluated once
  __shared__ double sF[2][kCipThreads];
  double dc_prev = 0.0;
  uint32_t f_prev = 0;
  int buf = 0;
  for (int j = jb; j < jl; ++j, buf ^= 1) {
    const int64_t idx = A.first + (int64_t)j * A.P + i;
    uint32_t f = 0;
    double dc = 0.0, nf = 0.0, Fp = 0.0;
    bool has_face = false;
    if (inside) f = A.flags[idx];
    if (shared_flux) {
      if (f & SB2_F_DOMAIN) cip_plus_face(A, 0, idx, i, A.nx, 1, f, Fp, nf, has_face);
      sF[buf][threadIdx.x] = Fp;
      __syncthreads();
      if (owner && (f & SB2_F_DOMAIN)) dc = cip_dc(A, 0, f, Fp, sF[buf][threadIdx.x - 1]);
    } else if (owner && (f & SB2_F_DOMAIN)) {
      cip_cell(A, 0, idx, i, A.nx, 1, f, dc, nf, has_face);
    }
    if (!owner) continue;
    if (j < je) {
      // val += val_delta at every point of the row (calcPDE.cpp:657): cells and the faces of this axis
      A.u_out[idx] = A.u[idx] + dc;
      A.face_out[idx] = A.faces[0][idx] + (has_face ? nf : 0.0);
    }
    if (j > jb && (f_prev & SB2_F_DOMAIN) && !(f_prev & SB2_F_YP) && j < A.ny) {
      // val(x, y+1/2) += (val_delta(x, y+1) + val_delta(x, y)) / 2 for the row below (calcPDE.cpp:659-663)
      const int64_t ip = idx - A.P;
      A.faces[1][ip] = A.faces[1][ip] + (dc + dc_prev) / 2.0;
    }
    dc_prev = dc;
    f_prev = f;
  }
}

__global__ void __launch_bounds__(kCipThreads) cip2d_y_kernel(const __grid_constant__ Sb2CipArgs A) {
  // strips overlap by one column: the last thread of a block only supplies the increment its left neighbour needs
  const int i = (int)blockIdx.x * (kCipThreads - 1) + threadIdx.x;
  const int jb = A.row_begin + (int)blockIdx.y * kCipRows;
  const int je = min(jb + kCipRows, A.row_end);
  __shared__ double sdc[2][kCipThreads];
  const bool inside = i < A.nx;
  const bool owner = inside && threadIdx.x < kCipThreads - 1;
  const bool shared_flux = A.akind[1] == SB2_SRC_CONST;
  // uniform velocity: the flux through the face below a cell is the plus-face flux of the row below, carried along the march
  double F_prev = 0.0;
  if (shared_flux && inside && jb > 0) {
    const int64_t idx = A.first + (int64_t)(jb - 1) * A.P + i;
    const uint32_t f = A.flags[idx];
    double nf; bool hf;
    if (f & SB2_F_DOMAIN) cip_plus_face(A, 1, idx, jb - 1, A.ny, A.P, f, F_prev, nf, hf);
  }
  int buf = 0;
  for (int j = jb; j < je; ++j, buf ^= 1) {
    const int64_t idx = A.first + (int64_t)j * A.P + i;
    uint32_t f = 0;
    double dc = 0.0, nf = 0.0;
    bool has_face = false;
    if (inside) {
      f = A.flags[idx];
      if (shared_flux) {
        double Fp = 0.0;
        if (f & SB2_F_DOMAIN) { cip_plus_face(A, 1, idx, j, A.ny, A.P, f, Fp, nf, has_face); dc = cip_dc(A, 1, f, Fp, F_prev); }
        F_prev = Fp;
      } else if (f & SB2_F_DOMAIN) {
        cip_cell(A, 1, idx, j, A.ny, A.P, f, dc, nf, has_face);
      }
    }
    sdc[buf][threadIdx.x] = dc;
    __syncthreads();
    if (owner) {
      A.u_out[idx] = A.u[idx] + dc;
      A.face_out[idx] = A.faces[1][idx] + (has_face ? nf : 0.0);
      if ((f & SB2_F_DOMAIN) && !(f & SB2_F_XP) && i + 1 < A.nx)
        A.faces[0][idx] = A.faces[0][idx] + (sdc[buf][threadIdx.x + 1] + dc) / 2.0;  // calcPDE.cpp:747-751
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 2-D, UNIFORM velocity along the pass axis (the common case: an advectionCoefficient that is a plain parameter): the same
// fused pass with warps that share nothing.  A warp owns a strip of a row block and marches over its rows; a lane holds TWO
// adjacent cells (16-byte loads and stores); what a cell needs from its neighbours across x comes through warp shuffles
// (the strips overlap by the one or two lanes that only supply values), what it needs along y is carried in registers along
// the march; the rows of the NEXT iteration are requested before the current one is computed, so the long fp64 dependency
// chain of a face (three divisions in sequence) runs under the loads.  No shared memory, no block barrier.  Divisions by
// +-delta are multiplications by the reciprocal when delta is a power of two (exact, hence bit-identical).
// The arithmetic is cip_plus_face / cip_dc operand for operand.
struct CipRow {
  double2 v, fa, fo;  // cell averages, faces of the pass axis, faces of the other axis
  uint32_t fl;        // flag bytes of the two cells (byte 0: first cell); zero for cells outside the grid
};

__device__ __forceinline__ CipRow cip_load_row(const Sb2CipArgs& A, const int axis, const int j, const int xa, const bool in_a, const bool in_b) {
  CipRow r;
  r.v = make_double2(0.0, 0.0); r.fa = r.v; r.fo = r.v; r.fl = 0;
  if (j < 0 || j >= A.ny || !(in_a || in_b)) return r;
  const int64_t idx = A.first + (int64_t)j * A.P + xa;
  // (a pair is 16-byte aligned: first and P are multiples of 4, xa is even; its second cell may lie in the row padding)
  r.v = *reinterpret_cast<const double2*>(A.u + idx);
  r.fa = *reinterpret_cast<const double2*>(A.faces[axis] + idx);
  r.fo = *reinterpret_cast<const double2*>(A.faces[axis ^ 1] + idx);
  const uint32_t f2 = *reinterpret_cast<const uint16_t*>(A.flags + idx);
  r.fl = (in_a ? (f2 & 0xffu) : 0u) | (in_b ? (f2 & 0xff00u) : 0u);
  return r;
}

template <bool P2>
__device__ __forceinline__ double cip_div_delta(const double x, const double d, const double rd) { return P2 ? x * rd : x / d; }

// F and the face increment of one cell (cip_plus_face with its operands in registers).  f: the cell's flag byte; d_*: whether
// the neighbours along the axis are cells of the domain; POS: ux >= 0 (uses v0, vp1 = face of the cell, fm1 = face of the
// minus neighbour), else (v0, vp1, vn = value of the plus neighbour, fn = its face)
template <int AXIS, bool POS, bool P2>
__device__ __forceinline__ void cip_face_regs(const uint32_t f, const bool d_p1, const bool d_p2, const bool d_m1, const double v0,
                                              const double face0, const double fm1, const double vn, const double fn, const double xi,
                                              const double Dx, const double rDx, double& F, double& nf, bool& has_face) {
  F = 0.0; nf = 0.0; has_face = false;
  if (!(f & SB2_F_DOMAIN) || (f & (SB2_F_XP << (2 * AXIS)))) return;
  const double vp1 = d_p1 ? face0 : v0;
  double a_i, b_i, beta_i;
  if (POS) {
    const double vm1 = d_m1 ? fm1 : v0;
    a_i = vp1;
    beta_i = cip_div_delta<P2>((fabs(vp1 - v0) + DBL_EPSILON) / (fabs(v0 - vm1) + DBL_EPSILON) - 1.0, Dx, rDx);
    b_i = cip_div_delta<P2>((1.0 + beta_i * Dx) * v0 - vp1, Dx, rDx);
  } else {
    const double vp3 = d_p2 ? fn : (d_p1 ? vn : v0);
    const double vp2 = d_p1 ? vn : v0;
    a_i = vp1;
    beta_i = cip_div_delta<P2>((fabs(vp1 - vp2) + DBL_EPSILON) / (fabs(vp2 - vp3) + DBL_EPSILON) - 1.0, Dx, rDx);
    b_i = cip_div_delta<P2>((1.0 + beta_i * Dx) * vp2 - vp1, Dx, rDx);
  }
  const double mxi = -xi;
  const double den = 1.0 + beta_i * mxi;
  const double n = (a_i + 2 * b_i * mxi + beta_i * b_i * (mxi * mxi)) / (den * den) - vp1;
  if (d_p1) { has_face = true; nf = n; }
  F = (-a_i * xi + b_i * (xi * xi)) / (-1.0 + beta_i * xi);
}

template <int AXIS, bool POS, bool P2>
__device__ __forceinline__ double cip_dc_regs(const uint32_t f, const double Fp, const double Fm, const double delta, const double rdelta) {
  if (!(f & SB2_F_DOMAIN)) return 0.0;
  const bool bofp = f & (SB2_F_XP << (2 * AXIS)), bofm = f & (SB2_F_XM << (2 * AXIS));
  double g_in = 0.0, g_out = 0.0;
  if (POS) { if (!bofp) g_out = Fp; if (!bofm) g_in = Fm; }
  else { if (!bofp) g_in = -Fp; if (!bofm) g_out = -Fm; }
  return cip_div_delta<P2>(g_in - g_out, delta, rdelta);
}

__device__ __forceinline__ double shfl_up_d(double x) { return __shfl_up_sync(0xffffffffu, x, 1); }
__device__ __forceinline__ double shfl_down_d(double x) { return __shfl_down_sync(0xffffffffu, x, 1); }

constexpr int kCipUWarps = 4;

// X pass: lanes 1..30 own their pair, lanes 0 and 31 only supply neighbour values: 60 owned cells per warp and row
template <bool POS, bool P2>
__global__ void __launch_bounds__(kCipUWarps * 32) cip2d_ux_kernel(const __grid_constant__ Sb2CipArgs A, const int rows_per_block) {
  const int lane = threadIdx.x & 31;
  const int strip = (int)blockIdx.x * kCipUWarps + (threadIdx.x >> 5);
  const int xa = strip * 60 - 2 + 2 * lane;
  if (strip * 60 >= A.nx) return;
  const bool in_a = xa >= 0 && xa < A.nx, in_b = xa + 1 >= 0 && xa + 1 < A.nx;
  const bool owner = lane >= 1 && lane <= 30 && in_a;
  const int jb = A.row_begin + (int)blockIdx.y * rows_per_block;
  const int je = min(jb + rows_per_block, A.row_end);
  const int jl = je < A.flux_row_end ? je + 1 : je;  // one more row (flux only): the face above the block's last row needs its increment
  const double ux = A.aconst[0], delta = A.delta[0];
  const double Dx = -copysign(1.0, ux) * delta, rDx = 1.0 / Dx, rdelta = 1.0 / delta, xi = ux * A.dt;
  double2 dc_prev = make_double2(0.0, 0.0), fo_prev = dc_prev;
  uint32_t fl_prev = 0;
  CipRow cur = cip_load_row(A, 0, jb, xa, in_a, in_b);
  for (int j = jb; j < jl; ++j) {
    const CipRow nxt = cip_load_row(A, 0, j + 1 < jl ? j + 1 : -1, xa, in_a, in_b);
    const uint32_t fa_ = cur.fl & 0xffu, fb_ = cur.fl >> 8;
    const uint32_t fl_l = __shfl_up_sync(0xffffffffu, cur.fl, 1), fl_r = __shfl_down_sync(0xffffffffu, cur.fl, 1);
    // neighbours across x: cell a has (lane-1).b on its left, b on its right, (lane+1).a two to the right
    const bool dom_b = fb_ & SB2_F_DOMAIN, dom_a = fa_ & SB2_F_DOMAIN;
    const bool dom_lb = (fl_l >> 8) & SB2_F_DOMAIN, dom_ra = fl_r & SB2_F_DOMAIN, dom_rb = (fl_r >> 8) & SB2_F_DOMAIN;
    double Fa, Fb, nfa, nfb;
    bool hfa, hfb;
    if (POS) {
      const double fa_l = shfl_up_d(cur.fa.y);
      cip_face_regs<0, true, P2>(fa_, dom_b, false, dom_lb, cur.v.x, cur.fa.x, fa_l, 0.0, 0.0, xi, Dx, rDx, Fa, nfa, hfa);
      cip_face_regs<0, true, P2>(fb_, dom_ra, false, dom_a, cur.v.y, cur.fa.y, cur.fa.x, 0.0, 0.0, xi, Dx, rDx, Fb, nfb, hfb);
    } else {
      const double v_r = shfl_down_d(cur.v.x), fa_r = shfl_down_d(cur.fa.x);
      cip_face_regs<0, false, P2>(fa_, dom_b, dom_ra, false, cur.v.x, cur.fa.x, 0.0, cur.v.y, cur.fa.y, xi, Dx, rDx, Fa, nfa, hfa);
      cip_face_regs<0, false, P2>(fb_, dom_ra, dom_rb, false, cur.v.y, cur.fa.y, 0.0, v_r, fa_r, xi, Dx, rDx, Fb, nfb, hfb);
    }
    const double Fl = shfl_up_d(Fb);  // the plus-face flux of the cell left of a
    double2 dc;
    dc.x = cip_dc_regs<0, POS, P2>(fa_, Fa, Fl, delta, rdelta);
    dc.y = cip_dc_regs<0, POS, P2>(fb_, Fb, Fa, delta, rdelta);
    if (owner) {
      const int64_t idx = A.first + (int64_t)j * A.P + xa;
      if (j < je) {
        // val += val_delta at every point of the row (calcPDE.cpp:657): cells and the faces of this axis
        const double2 un = make_double2(cur.v.x + dc.x, cur.v.y + dc.y);
        const double2 fn = make_double2(cur.fa.x + (hfa ? nfa : 0.0), cur.fa.y + (hfb ? nfb : 0.0));
        if (in_b) {
          *reinterpret_cast<double2*>(A.u_out + idx) = un;
          *reinterpret_cast<double2*>(A.face_out + idx) = fn;
        } else {
          A.u_out[idx] = un.x;
          A.face_out[idx] = fn.x;
        }
      }
      if (j > jb && j < A.ny) {
        // val(x, y+1/2) += (val_delta(x, y+1) + val_delta(x, y)) / 2 for the row below (calcPDE.cpp:659-663)
        const uint32_t pa = fl_prev & 0xffu, pb = fl_prev >> 8;
        const bool ua = (pa & SB2_F_DOMAIN) && !(pa & SB2_F_YP), ub = (pb & SB2_F_DOMAIN) && !(pb & SB2_F_YP);
        if (ua || ub) {
          double2 fo = fo_prev;
          if (ua) fo.x = fo.x + (dc.x + dc_prev.x) / 2.0;
          if (ub) fo.y = fo.y + (dc.y + dc_prev.y) / 2.0;
          double* dst = A.faces[1] + (idx - A.P);
          if (ub) *reinterpret_cast<double2*>(dst) = fo; else dst[0] = fo.x;
        }
      }
    }
    dc_prev = dc; fo_prev = cur.fo; fl_prev = cur.fl;
    cur = nxt;
  }
}

// Y pass: the march runs along the pass axis, so the minus-face flux of a row is the plus-face flux of the previous one
// (carried); across x only the increment of the right neighbour is needed (other-axis faces): lane 31 only supplies it,
// 62 owned cells per warp and row
template <bool POS, bool P2>
__global__ void __launch_bounds__(kCipUWarps * 32) cip2d_uy_kernel(const __grid_constant__ Sb2CipArgs A, const int rows_per_block) {
  const int lane = threadIdx.x & 31;
  const int strip = (int)blockIdx.x * kCipUWarps + (threadIdx.x >> 5);
  const int xa = strip * 62 + 2 * lane;
  if (strip * 62 >= A.nx) return;
  const bool in_a = xa < A.nx, in_b = xa + 1 < A.nx;
  const bool owner = lane <= 30 && in_a;
  const int jb = A.row_begin + (int)blockIdx.y * rows_per_block;
  const int je = min(jb + rows_per_block, A.row_end);
  const double ux = A.aconst[1], delta = A.delta[1];
  const double Dx = -copysign(1.0, ux) * delta, rDx = 1.0 / Dx, rdelta = 1.0 / delta, xi = ux * A.dt;
  // the march starts one row early (flux only) to have the flux through the lower face of the block's first row
  const int js = jb > 0 ? jb - 1 : jb;
  CipRow prev = cip_load_row(A, 1, js - 1, xa, in_a, in_b);
  CipRow cur = cip_load_row(A, 1, js, xa, in_a, in_b);
  CipRow nxt = cip_load_row(A, 1, js + 1, xa, in_a, in_b);
  double2 F_prev = make_double2(0.0, 0.0);
  for (int j = js; j < je; ++j) {
    // two rows ahead: ux < 0 needs the flags of row j + 2, and the loads get a longer start
    const CipRow nn = cip_load_row(A, 1, j + 2, xa, in_a, in_b);
    const uint32_t fa_ = cur.fl & 0xffu, fb_ = cur.fl >> 8;
    double2 F, nf;
    bool hfa, hfb;
    cip_face_regs<1, POS, P2>(fa_, nxt.fl & SB2_F_DOMAIN, nn.fl & SB2_F_DOMAIN, prev.fl & SB2_F_DOMAIN, cur.v.x, cur.fa.x, prev.fa.x, nxt.v.x,
                              nxt.fa.x, xi, Dx, rDx, F.x, nf.x, hfa);
    cip_face_regs<1, POS, P2>(fb_, (nxt.fl >> 8) & SB2_F_DOMAIN, (nn.fl >> 8) & SB2_F_DOMAIN, (prev.fl >> 8) & SB2_F_DOMAIN, cur.v.y, cur.fa.y,
                              prev.fa.y, nxt.v.y, nxt.fa.y, xi, Dx, rDx, F.y, nf.y, hfb);
    if (j >= jb) {
      double2 dc;
      dc.x = cip_dc_regs<1, POS, P2>(fa_, F.x, F_prev.x, delta, rdelta);
      dc.y = cip_dc_regs<1, POS, P2>(fb_, F.y, F_prev.y, delta, rdelta);
      const double dc_r = shfl_down_d(dc.x);  // the increment of the cell right of b
      if (owner) {
        const int64_t idx = A.first + (int64_t)j * A.P + xa;
        const double2 un = make_double2(cur.v.x + dc.x, cur.v.y + dc.y);
        const double2 fn = make_double2(cur.fa.x + (hfa ? nf.x : 0.0), cur.fa.y + (hfb ? nf.y : 0.0));
        if (in_b) {
          *reinterpret_cast<double2*>(A.u_out + idx) = un;
          *reinterpret_cast<double2*>(A.face_out + idx) = fn;
        } else {
          A.u_out[idx] = un.x;
          A.face_out[idx] = fn.x;
        }
        // val(x+1/2, y) += (val_delta(x+1, y) + val_delta(x, y)) / 2 (calcPDE.cpp:747-751)
        const bool ua = (fa_ & SB2_F_DOMAIN) && !(fa_ & SB2_F_XP) && xa + 1 < A.nx;
        const bool ub = (fb_ & SB2_F_DOMAIN) && !(fb_ & SB2_F_XP) && xa + 2 < A.nx;
        if (ua || ub) {
          double2 fo = cur.fo;
          if (ua) fo.x = fo.x + (dc.y + dc.x) / 2.0;
          if (ub) fo.y = fo.y + (dc_r + dc.y) / 2.0;
          double* dst = A.faces[0] + idx;
          if (ub) *reinterpret_cast<double2*>(dst) = fo; else dst[0] = fo.x;
        }
      }
    }
    F_prev = F;
    prev = cur; cur = nxt; nxt = nn;
  }
}

template <bool POS, bool P2>
cudaError_t cip2d_uniform_launch(const Sb2CipArgs& a, int axis, cudaStream_t stream) {
  const int nupd = a.row_end - a.row_begin;
  const int per = axis == 0 ? 60 : 62;
  const int nstrips = (a.nx + per - 1) / per;
  // rows per block: long marches amortise the prologue rows, short ones fill the machine
  // (small grids are bound by the latency of a march, not by throughput: down to two rows per warp, four along the pass axis
  // where every march pays a flux-only row)
  int rpb = 32;  // (measured at 4096^2 / 8192^2: 8, 16, 32, 64, 128 rows -> 0.178, 0.177, 0.175, 0.211, 0.200 / -, 0.661, 0.649, 0.661, 0.669 ms per pass)
  while (rpb > (axis == 0 ? 2 : 4) && (int64_t)nstrips * ((nupd + rpb - 1) / rpb) < 148 * 48) rpb /= 2;
  if (const char* e = getenv("SB2_CIP_RPB")) rpb = atoi(e) > 0 ? atoi(e) : rpb;  // experiments
  const dim3 grid((nstrips + kCipUWarps - 1) / kCipUWarps, (nupd + rpb - 1) / rpb);
  if (axis == 0) cip2d_ux_kernel<POS, P2><<<grid, kCipUWarps * 32, 0, stream>>>(a, rpb);
  else cip2d_uy_kernel<POS, P2><<<grid, kCipUWarps * 32, 0, stream>>>(a, rpb);
  return cudaGetLastError();
}

// update sweep (calcPDE.cpp:654-673): add the increments and advance the faces of the other axes
__global__ void __launch_bounds__(256) cip_update_kernel(const __grid_constant__ Sb2CipArgs A, int axis) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y + (A.dim == 2 ? A.row_begin : 0), k = blockIdx.z + (A.dim == 3 ? A.row_begin : 0);
  if (i >= A.nx) return;
  const int64_t idx = A.first + (int64_t)k * A.PL + (int64_t)j * A.P + i;
  const double dc = A.tmp_cell[idx];
  A.u[idx] = A.u[idx] + dc;
  A.faces[axis][idx] = A.faces[axis][idx] + A.tmp_face[idx];
  const uint32_t f = A.flags[idx];
  if (!(f & SB2_F_DOMAIN)) return;
  for (int b = 0; b < A.dim; ++b) {
    if (b == axis) continue;
    if (f & (SB2_F_XP << (2 * b))) continue;
    const int c = b == 0 ? i : b == 1 ? j : k;
    const int n = b == 0 ? A.nx : b == 1 ? A.ny : A.nz;
    if (c + 1 >= n) continue;
    const int64_t st = b == 0 ? 1 : b == 1 ? A.P : A.PL;
    A.faces[b][idx] = A.faces[b][idx] + (A.tmp_cell[idx + st] + dc) / 2.0;
  }
}

}  // namespace

cudaError_t sb2_launch_cip_pass(const Sb2CipArgs& a, int axis, cudaStream_t stream, int* launches) {
  // the reference runs the update sweep of a pass even without a velocity along it; with zero
  // increments that sweep changes nothing, so the caller skips such passes
  if (a.akind[axis] == SB2_SRC_NONE) return cudaSuccess;
  const int nflux = a.flux_row_end - a.row_begin, nupd = a.row_end - a.row_begin;
  if (nupd <= 0) return cudaSuccess;
  if (a.dim == 2 && a.u_out && a.face_out) {
    // fused pass: the caller swaps u / u_out and faces[axis] / face_out afterwards
    if (launches) *launches += 1;
    static const bool old_pass = getenv("SB2_CIP_OLD") != NULL;  // experiments: the block-cooperative pass for uniform velocities too
    if (a.akind[axis] == SB2_SRC_CONST && !old_pass) {
      int e2 = 0;
      const bool p2 = frexp(a.delta[axis], &e2) == 0.5;  // a power of two: x / delta == x * (1 / delta) bit for bit
      const bool pos = a.aconst[axis] >= 0;
      return pos ? (p2 ? cip2d_uniform_launch<true, true>(a, axis, stream) : cip2d_uniform_launch<true, false>(a, axis, stream))
                 : (p2 ? cip2d_uniform_launch<false, true>(a, axis, stream) : cip2d_uniform_launch<false, false>(a, axis, stream));
    }
    if (axis == 0) cip2d_x_kernel<<<dim3((a.nx + kCipThreads - 2) / (kCipThreads - 1), (nupd + kCipRows - 1) / kCipRows), kCipThreads, 0, stream>>>(a);
    else cip2d_y_kernel<<<dim3((a.nx + kCipThreads - 2) / (kCipThreads - 1), (nupd + kCipRows - 1) / kCipRows), kCipThreads, 0, stream>>>(a);
    return cudaGetLastError();
  }
  cudaError_t e = cudaMemsetAsync(a.tmp_cell, 0, sizeof(double) * a.nalloc, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(a.tmp_face, 0, sizeof(double) * a.nalloc, stream);
  if (e != cudaSuccess) return e;
  if (launches) *launches += 2;
  dim3 block(256);
  dim3 gflux((a.nx + 255) / 256, a.dim == 2 ? nflux : a.ny, a.dim == 3 ? nflux : 1);
  dim3 gupd((a.nx + 255) / 256, a.dim == 2 ? nupd : a.ny, a.dim == 3 ? nupd : 1);
  cip_flux_kernel<<<gflux, block, 0, stream>>>(a, axis);
  cip_update_kernel<<<gupd, block, 0, stream>>>(a, axis);
  return cudaGetLastError();
}

// =============================================================================================
// membrane transport
// =============================================================================================
namespace {

struct MtEvalArgs {
  int npts, ntoks, m;
  double cdt;
  const uint32_t* tok;         // kind | op<<8 | slot<<16
  const int32_t* tok_operand;  // operand-table row or -1
  const int32_t* near_cell;
  const int32_t* far_cell;
  const double* fieldval;
  double* rate;
  const double* u[SB2_MAX_SPECIES];
  const double* kprev[SB2_MAX_SPECIES];
  const double* mu[SB2_MAX_MEM_SPECIES];      // membrane species (binding): value at the point's storage index
  const double* mkprev[SB2_MAX_MEM_SPECIES];
  double consts[SB2_MAX_CONSTS];
};

// interpreter of calcMemTransport (calcPDE.cpp:1084-1393): the rate at membrane point p
__device__ __noinline__ double mt_rate(const MtEvalArgs& A, const int p) {
  double st[50];
  // (the stack never grows beyond the number of tokens: only that part has to read as zero)
  const int ninit = A.ntoks + 2 < 50 ? A.ntoks + 2 : 50;
  for (int i = 0; i < ninit; ++i) st[i] = 0.0;
  int sp = 0;
  for (int i = 0; i < A.ntoks; ++i) {
    const uint32_t t = A.tok[i];
    const int kind = t & 0xff, op = (t >> 8) & 0xff, slot = t >> 16;
    if (kind == SB2_TOK_VAR) {
      const int row = A.tok_operand[i];
      const int nc = A.near_cell[(int64_t)row * A.npts + p];
      const int fc = A.far_cell[(int64_t)row * A.npts + p];
      if (nc >= 0) {
        // 1.5*value(boundary) - 0.5*value(boundary_next) when the second cell exists (calcPDE.cpp:1107-1213)
        double wn = A.u[slot][nc];
        if (A.m > 0) wn = wn + A.cdt * A.kprev[slot][nc];
        if (fc >= 0) {
          double wf = A.u[slot][fc];
          if (A.m > 0) wf = wf + A.cdt * A.kprev[slot][fc];
          st[sp] = 1.5 * wn - 0.5 * wf;
        } else {
          st[sp] = wn;
        }
      }
      if (sp < 49) ++sp;
    } else if (kind == SB2_TOK_MEMVAR) {
      // symbol in the membrane: its value at the point itself (calcPDE.cpp:1214-1217)
      const int nc = A.near_cell[(int64_t)A.tok_operand[i] * A.npts + p];
      if (nc >= 0) {
        double w = A.mu[slot][nc];
        if (A.m > 0) w = w + A.cdt * A.mkprev[slot][nc];
        st[sp] = w;
      }
      if (sp < 49) ++sp;
    } else if (kind == SB2_TOK_FIELD) {
      st[sp] = A.fieldval[(int64_t)A.tok_operand[i] * A.npts + p];
      if (sp < 49) ++sp;
    } else if (kind == SB2_TOK_CONST) {
      st[sp] = A.consts[slot];
      if (sp < 49) ++sp;
    } else {
      if (sp > 0) --sp;  // the reference decrements unconditionally; guard only keeps us in the array
      if (kind != SB2_TOK_OP) continue;
      if (op >= SB2_OP_PLUS && op <= SB2_OP_NEQ) {
        if (sp > 0) st[sp - 1] = sb2_binary_any(op, st[sp - 1], st[sp]);
      } else if (op >= SB2_OP_ABS && op <= SB2_OP_NOT) {
        st[sp] = sb2_unary(op, st[sp]);
        if (sp < 49) ++sp;
      }
    }
  }
  // "if (st_index > 1) st_index--;" then the rate is rpStack[st_index] (calcPDE.cpp:1392-1393): with the
  // normal final depth 1 this reads slot 1 — the right operand of the last binary operator
  if (sp > 1) --sp;
  return st[sp];
}

struct MtApplyArgs {
  int n_groups;
  const int32_t* group_start;
  const int32_t* group_species;
  const int64_t* group_cell;
  const int32_t* entry_pt;
  const double* entry_coef;
  const double* entry_delta;
  const int32_t* entry_sign;
  const uint8_t* group_first;  // NULL: every group accumulates on top of what its cell holds
  double* kout[SB2_MAX_SPECIES + SB2_MAX_MEM_SPECIES];  // volume species, then membrane species
};

// gather formulation of the scatter at calcPDE.cpp:1394-1479: one thread per (species, cell) sums its
// contributions in the reference's order
// The rate of every membrane point a cell is fed by is evaluated by that cell's thread (a point feeds two to four cells; the
// evaluation is a handful of operations), so one launch does what took a rate kernel and a gather kernel.
__global__ void __launch_bounds__(128) mt_fused_kernel(const __grid_constant__ MtEvalArgs E, const __grid_constant__ MtApplyArgs A) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= A.n_groups) return;
  double* kp = A.kout[A.group_species[g]] + A.group_cell[g];
  double acc = (A.group_first && A.group_first[g]) ? 0.0 : *kp;
  for (int e = A.group_start[g]; e < A.group_start[g + 1]; ++e) {
    // fabs(n) * stoich * rate / delta
    const double term = A.entry_coef[e] * mt_rate(E, A.entry_pt[e]) / A.entry_delta[e];
    acc = A.entry_sign[e] < 0 ? acc - term : acc + term;
  }
  *kp = acc;
}

struct MtEntry {
  int species;
  int64_t cell;
  int pt;
  double coef, delta;
  int sign;
};

}  // namespace

int sb2_build_transport(const sb2_grid& g, int64_t P, int64_t PL, int64_t first, int nspecies, const sb2_token* toks,
                        int ntoks, const sb2_ref* refs, int nrefs, int n_reactants, const sb2_mem_point* pts, int npts,
                        const int32_t* operand_cells, const double* operand_values, const int32_t* target_cells,
                        cudaStream_t stream, Sb2MemTransport& out, std::string& err) {
  out = Sb2MemTransport();
  out.npts = npts; out.ntoks = ntoks; out.nspecies = nspecies;
  if (npts < 0 || ntoks < 0 || (npts > 0 && (!pts || !target_cells))) { err = "membrane transport: null argument"; return SB2_ERR_ARG; }
  const int64_t nxy = (int64_t)g.nx * g.ny;
  auto pad = [&](int32_t c) -> int32_t {
    if (c < 0) return -1;
    const int k = (int)(c / nxy), j = (int)((c - k * nxy) / g.nx), i = (int)(c - k * nxy - (int64_t)j * g.nx);
    return (int32_t)(first + (int64_t)k * PL + (int64_t)j * P + i);
  };
  std::vector<int32_t> near_c, far_c;
  std::vector<double> fval;
  int nvar = 0, nfield = 0, rows = 0;
  for (int i = 0; i < ntoks; ++i) {
    out.tok.push_back((uint32_t)toks[i].kind | ((uint32_t)toks[i].op << 8) | ((uint32_t)toks[i].slot << 16));
    if (toks[i].kind == SB2_TOK_VAR || toks[i].kind == SB2_TOK_MEMVAR) {
      const bool mem = toks[i].kind == SB2_TOK_MEMVAR;
      if (!mem && toks[i].slot >= nspecies) { err = "membrane transport: token references unknown species"; return SB2_ERR_ARG; }
      if (mem && toks[i].slot >= SB2_MAX_MEM_SPECIES) { err = "membrane transport: token references unknown membrane species"; return SB2_ERR_ARG; }
      if (!operand_cells && npts > 0) { err = "membrane transport: operand cells missing"; return SB2_ERR_ARG; }
      out.tok_operand.push_back(rows++);
      for (int p = 0; p < npts; ++p) {
        const int32_t c0 = operand_cells[((int64_t)nvar * npts + p) * 2 + 0], c1 = operand_cells[((int64_t)nvar * npts + p) * 2 + 1];
        near_c.push_back(mem ? c0 : pad(c0));  // membrane species: storage index of the point, as it is
        far_c.push_back(mem ? -1 : pad(c1));
        fval.push_back(0.0);
      }
      ++nvar;
    } else if (toks[i].kind == SB2_TOK_FIELD) {
      if (!operand_values && npts > 0) { err = "membrane transport: field operand values missing"; return SB2_ERR_ARG; }
      out.tok_operand.push_back(rows++);
      for (int p = 0; p < npts; ++p) {
        near_c.push_back(-1); far_c.push_back(-1);
        fval.push_back(operand_values[(int64_t)nfield * npts + p]);
      }
      ++nfield;
    } else {
      out.tok_operand.push_back(-1);
    }
  }
  // contributions in the reference's order: points ascending; reactants (plus side, minus side), then products
  const double delta[3] = {g.dx, g.dy, g.dz};
  std::vector<MtEntry> entries;
  for (int p = 0; p < npts; ++p) {
    for (int r = 0; r < nrefs; ++r) {
      if (!refs[r].is_variable || refs[r].species < 0) continue;
      const bool mem = refs[r].species >= SB2_MEM_SPECIES;
      if (!mem && refs[r].species >= nspecies) { err = "membrane transport: reference to unknown species"; return SB2_ERR_ARG; }
      if (mem && refs[r].species - SB2_MEM_SPECIES >= SB2_MAX_MEM_SPECIES) { err = "membrane transport: reference to unknown membrane species"; return SB2_ERR_ARG; }
      for (int side = 0; side < 2; ++side) {
        const int32_t c = target_cells[((int64_t)r * npts + p) * 2 + side];
        if (c < 0) continue;
        MtEntry e;
        e.pt = p;
        e.coef = pts[p].abs_normal * refs[r].stoichiometry;
        if (mem) {  // binding: the point itself, divided by (delta / 2.0) (calcPDE.cpp:1405-1407)
          e.species = SB2_MAX_SPECIES + (refs[r].species - SB2_MEM_SPECIES); e.cell = c;
          e.delta = delta[pts[p].axis] / 2.0;
        } else {
          e.species = refs[r].species; e.cell = pad(c);
          e.delta = delta[pts[p].axis];
        }
        e.sign = r < n_reactants ? -1 : +1;
        entries.push_back(e);
      }
    }
  }
  std::stable_sort(entries.begin(), entries.end(), [](const MtEntry& a, const MtEntry& b) {
    if (a.species != b.species) return a.species < b.species;
    return a.cell < b.cell;
  });
  std::vector<int32_t> gstart, gspecies, ept, esign;
  std::vector<int64_t> gcell;
  std::vector<double> ecoef, edelta;
  for (size_t i = 0; i < entries.size(); ++i) {
    if (i == 0 || entries[i].species != entries[i - 1].species || entries[i].cell != entries[i - 1].cell) {
      gstart.push_back((int32_t)i); gspecies.push_back(entries[i].species); gcell.push_back(entries[i].cell);
    }
    ept.push_back(entries[i].pt); ecoef.push_back(entries[i].coef); edelta.push_back(entries[i].delta);
    esign.push_back(entries[i].sign);
  }
  gstart.push_back((int32_t)entries.size());
  out.n_groups = (int)gspecies.size(); out.n_entries = (int)entries.size();
  out.h_group_species = gspecies; out.h_group_cell = gcell; out.d_group_first = NULL;
  cudaError_t e;
  std::vector<double> rate0(std::max(npts, 1), 0.0);
  if ((e = to_device(&out.d_near, near_c, stream)) != cudaSuccess || (e = to_device(&out.d_far, far_c, stream)) != cudaSuccess ||
      (e = to_device(&out.d_fieldval, fval, stream)) != cudaSuccess || (e = to_device(&out.d_rate, rate0, stream)) != cudaSuccess ||
      (e = to_device(&out.d_group_start, gstart, stream)) != cudaSuccess ||
      (e = to_device(&out.d_group_species, gspecies, stream)) != cudaSuccess ||
      (e = to_device(&out.d_group_cell, gcell, stream)) != cudaSuccess || (e = to_device(&out.d_entry_pt, ept, stream)) != cudaSuccess ||
      (e = to_device(&out.d_entry_coef, ecoef, stream)) != cudaSuccess ||
      (e = to_device(&out.d_entry_delta, edelta, stream)) != cudaSuccess ||
      (e = to_device(&out.d_entry_sign, esign, stream)) != cudaSuccess) {
    err = std::string("membrane transport: ") + cudaGetErrorString(e);
    return SB2_ERR_CUDA;
  }
  // the program itself: uploaded once (it never changes; the constants it reads travel with every launch)
  std::vector<uint32_t> tokv(out.tok);
  std::vector<int32_t> opv(out.tok_operand);
  if (tokv.empty()) { tokv.push_back(0); opv.push_back(-1); }
  if ((e = to_device(&out.d_tok, tokv, stream)) != cudaSuccess || (e = to_device(&out.d_operand, opv, stream)) != cudaSuccess) {
    err = std::string("membrane transport: ") + cudaGetErrorString(e);
    return SB2_ERR_CUDA;
  }
  return SB2_OK;
}

void sb2_free_transport(Sb2MemTransport& t) {
  cudaFree(t.d_near); cudaFree(t.d_far); cudaFree(t.d_fieldval); cudaFree(t.d_rate);
  cudaFree(t.d_group_start); cudaFree(t.d_group_species); cudaFree(t.d_group_cell);
  cudaFree(t.d_entry_pt); cudaFree(t.d_entry_coef); cudaFree(t.d_entry_delta); cudaFree(t.d_entry_sign);
  cudaFree(t.d_tok); cudaFree(t.d_operand); cudaFree(t.d_group_first);
}

cudaError_t sb2_launch_transport(const Sb2MemTransport& t, const sb2_grid& g, const double* const* u,
                                 const double* const* kprev, double* const* kout, const double* const* mu,
                                 const double* const* mkprev, double* const* mkout, const double* consts, int m,
                                 double cdt, bool first_resets, cudaStream_t stream, int* launches) {
  (void)g;
  if (t.npts == 0 || t.n_groups == 0) return cudaSuccess;
  MtEvalArgs a;
  a.npts = t.npts; a.ntoks = t.ntoks; a.m = m; a.cdt = cdt;
  a.tok = t.d_tok; a.tok_operand = t.d_operand;
  a.near_cell = t.d_near; a.far_cell = t.d_far; a.fieldval = t.d_fieldval; a.rate = t.d_rate;
  for (int s = 0; s < SB2_MAX_SPECIES; ++s) { a.u[s] = NULL; a.kprev[s] = NULL; }
  for (int s = 0; s < SB2_MAX_MEM_SPECIES; ++s) { a.mu[s] = NULL; a.mkprev[s] = NULL; }
  for (size_t i = 0; i < t.tok.size(); ++i) {
    if ((t.tok[i] & 0xff) == SB2_TOK_VAR) { const int s = t.tok[i] >> 16; a.u[s] = u[s]; a.kprev[s] = kprev[s]; }
    if ((t.tok[i] & 0xff) == SB2_TOK_MEMVAR && mu) { const int s = t.tok[i] >> 16; a.mu[s] = mu[s]; a.mkprev[s] = mkprev[s]; }
  }
  memcpy(a.consts, consts, sizeof(a.consts));
  MtApplyArgs b;
  b.n_groups = t.n_groups;
  b.group_start = t.d_group_start; b.group_species = t.d_group_species; b.group_cell = t.d_group_cell;
  b.entry_pt = t.d_entry_pt; b.entry_coef = t.d_entry_coef; b.entry_delta = t.d_entry_delta; b.entry_sign = t.d_entry_sign;
  b.group_first = first_resets ? t.d_group_first : NULL;
  for (int s = 0; s < SB2_MAX_SPECIES; ++s) b.kout[s] = s < t.nspecies ? kout[s] : NULL;
  for (int s = 0; s < SB2_MAX_MEM_SPECIES; ++s) b.kout[SB2_MAX_SPECIES + s] = mkout ? mkout[s] : NULL;
  if (launches) ++*launches;
  mt_fused_kernel<<<(t.n_groups + 127) / 128, 128, 0, stream>>>(a, b);
  return cudaGetLastError();
}

cudaError_t sb2_transport_set_first(Sb2MemTransport& t, const std::vector<uint8_t>& first, cudaStream_t stream) {
  if (t.d_group_first) { cudaFree(t.d_group_first); t.d_group_first = NULL; }
  return to_device(&t.d_group_first, first, stream);
}

// =============================================================================================
// stability: largest admissible dt (checkStability.cpp:10-52 and 119-154)
// =============================================================================================
namespace {

__global__ void __launch_bounds__(256) min_dt_kernel(const __grid_constant__ Sb2MinDtArgs A, double* block_min) {
  const int64_t ncell = (int64_t)A.nx * A.ny * A.nz;
  double best = A.dt;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(c / ((int64_t)A.nx * A.ny));
    const int j = (int)((c - (int64_t)k * A.nx * A.ny) / A.nx);
    const int i = (int)(c - (int64_t)k * A.nx * A.ny - (int64_t)j * A.nx);
    const int64_t idx = A.first + (int64_t)k * A.PL + (int64_t)j * A.P + i;
    if (!(A.flags[idx] & SB2_F_DOMAIN)) continue;
    if (A.has_diff && A.dkind[0] != SB2_SRC_NONE) {
      const double d0 = A.dkind[0] == SB2_SRC_CONST ? A.dconst[0] : A.dfield[0][idx];
      // x; the y / z tests compare against their own coefficient but take the minimum with the
      // x coefficient (checkStability.cpp:39-40, 45-46)
      if (A.dt >= A.delta[0] * A.delta[0] / (2.0 * d0)) best = fmin(best, A.delta[0] * A.delta[0] / (2.0 * d0));
      for (int a = 1; a < 3; ++a) {
        if (A.dkind[a] == SB2_SRC_NONE) continue;
        const double da = A.dkind[a] == SB2_SRC_CONST ? A.dconst[a] : A.dfield[a][idx];
        if (A.dt >= A.delta[a] * A.delta[a] / (2.0 * da)) best = fmin(best, A.delta[a] * A.delta[a] / (2.0 * d0));
      }
    }
    if (A.has_adv) {
      for (int a = 0; a < A.dim; ++a) {
        if (A.akind[a] == SB2_SRC_NONE) continue;
        const double ua = A.akind[a] == SB2_SRC_CONST ? A.aconst[a] : A.afield[a][idx];
        if (ua * (A.dt / A.delta[a]) >= 1) best = fmin(best, A.delta[a] / ua);
      }
    }
  }
  // warp-shuffle min, then one value per warp through shared memory
  for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
  __shared__ double wmin[8];
  if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? wmin[threadIdx.x] : A.dt;
    for (int o = 4; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) block_min[blockIdx.x] = v;
  }
}

}  // namespace

cudaError_t sb2_launch_min_dt(const Sb2MinDtArgs& a, double* d_scratch, cudaStream_t stream, int* launches, double* result) {
  const int64_t ncell = (int64_t)a.nx * a.ny * a.nz;
  int blocks = (int)std::min<int64_t>((ncell + 255) / 256, 148 * 8);
  if (blocks < 1) blocks = 1;
  if (launches) ++*launches;
  min_dt_kernel<<<blocks, 256, 0, stream>>>(a, d_scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  std::vector<double> h(blocks);
  e = cudaMemcpyAsync(h.data(), d_scratch, sizeof(double) * blocks, cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return e;
  double best = a.dt;
  for (int i = 0; i < blocks; ++i) best = std::min(best, h[i]);
  *result = best;
  return cudaSuccess;
}
