// mem_kernels.cu — membrane species on sparse point sets and the generic expression evaluator (sm_100a).  Compiled with
// -fmad=false; every arithmetic step that has to round like the reference uses the explicit round-to-nearest helpers of
// sb2_ops.cuh in the reference's order of operations.
//
//   mem_react_kernel      reversePolishRK over a membrane's domainIndex            calcPDE.cpp:224-451
//   mem_diffusion_kernel  calcMemDiffusion                                         calcPDE.cpp:1484-1581
//   mem_update_kernel     updateSpecies for a membrane species                     spatialsimulator.cpp:1517-1544
//   mem_fill_kernel       pseudo-membrane fill of updateAssignmentRules            spatialsimulator.cpp:1614-1663
//   mem_min_dt_kernel     checkMemDiffusionStab                                    checkStability.cpp:54-117
//   cell_rule_kernel      reversePolishInitial for per-step assignment rules       calcPDE.cpp:21-222, spatialsimulator.cpp:1566-1581
#include "sb2_mem.h"

#include <algorithm>
#include <cstring>

#include "sb2_ops.cuh"

namespace {

template <typename T>
cudaError_t to_device(T** dptr, const T* h, size_t n, cudaStream_t stream) {
  *dptr = NULL;
  if (n == 0) return cudaSuccess;
  cudaError_t e = cudaMalloc((void**)dptr, sizeof(T) * n);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(*dptr, h, sizeof(T) * n, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(stream);
}

// The reference interpreter's stack discipline (calcPDE.cpp:244-430): operands push; every operator first pops, binary
// operators then combine into the new top if something is left below, unary operators rewrite the popped slot and push it
// back, listed-but-empty operators and unresolved names only pop.  `load(i)` yields the operand of token i.
template <class Load>
__device__ __forceinline__ int eval_stream(const uint32_t* tok, int ntoks, const double* consts, double* st, Load load) {
  int sp = 0;
  for (int i = 0; i < ntoks; ++i) {
    const uint32_t t = tok[i];
    const int kind = t & 0xff, op = (t >> 8) & 0xff, slot = t >> 16;
    if (kind == SB2_TOK_VAR || kind == SB2_TOK_FIELD || kind == SB2_TOK_MEMVAR) {
      st[sp] = load(i, kind, slot);
      if (sp < 49) ++sp;
    } else if (kind == SB2_TOK_CONST) {
      st[sp] = consts[slot];
      if (sp < 49) ++sp;
    } else {
      if (sp > 0) --sp;
      if (kind != SB2_TOK_OP) continue;
      if (op >= SB2_OP_PLUS && op <= SB2_OP_NEQ) {
        if (sp > 0) st[sp - 1] = sb2_binary_any(op, st[sp - 1], st[sp]);
      } else if (op >= SB2_OP_ABS && op <= SB2_OP_NOT) {
        st[sp] = sb2_unary(op, st[sp]);
        if (sp < 49) ++sp;
      }
    }
  }
  return sp;
}

// ---------------------------------------------------------------------------------------------
struct MemReactArgs {
  int npts, ntoks, m, kind, nrefs, set_npts;
  double cdt;
  const uint32_t* tok;
  const int32_t* operand;
  const int32_t* points;
  const double* opvalues;
  int ref_species[8], ref_sign[8];
  double ref_stoich[8];
  Sb2MemView v;
  double consts[SB2_MAX_CONSTS];
};

__global__ void __launch_bounds__(128) mem_react_kernel(const __grid_constant__ MemReactArgs A) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.npts) return;
  const int idx = A.points[p];
  double st[50];
  for (int i = 0; i < 50; ++i) st[i] = 0.0;
  int sp = eval_stream(A.tok, A.ntoks, A.consts, st, [&](int i, int kind, int slot) -> double {
    if (kind == SB2_TOK_MEMVAR) {
      double w = A.v.u[slot][idx];
      if (A.m > 0) w = dadd(w, dmul(A.cdt, A.v.k[slot][A.m - 1][idx]));  // value + rk[m]*dt*delta[m-1] (calcPDE.cpp:247)
      return w;
    }
    return A.opvalues[(int64_t)A.operand[i] * A.npts + p];
  });
  if (sp > 0) --sp;  // "if (st_index > 0) st_index--;" then the rate is rpStack[st_index] (calcPDE.cpp:430-431)
  const double rate = st[sp];
  if (A.kind == 0) {
    for (int r = 0; r < A.nrefs; ++r) {
      double* kp = A.v.k[A.ref_species[r]][A.m] + idx;
      const double term = dmul(A.ref_stoich[r], rate);
      *kp = A.ref_sign[r] < 0 ? dsub(*kp, term) : dadd(*kp, term);
    }
  } else if (A.nrefs > 0) {
    double* kp = A.v.k[A.ref_species[0]][A.m] + idx;
    *kp = dadd(*kp, dmul(A.ref_stoich[0], rate));
  }
}

// ---------------------------------------------------------------------------------------------
struct MemDiffArgs {
  int npts, m, id, dkind;
  double cdt, dconst;
  const int32_t* pts;
  const int32_t* term_start;
  const int32_t* adj;
  const double *s, *d, *area, *dvalues;
  Sb2MemView v;
};

__global__ void __launch_bounds__(128) mem_diffusion_kernel(const __grid_constant__ MemDiffArgs A) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.npts) return;
  const int i = A.pts[p];
  const double* u = A.v.u[A.id];
  const double* kprev = A.m > 0 ? A.v.k[A.id][A.m - 1] : NULL;
  const double D = A.dkind == SB2_SRC_FIELD ? A.dvalues[p] : A.dconst;
  const double area = A.area[p];
  double acc = A.v.k[A.id][A.m][i];
  if (A.m == 0) {
    const double ui = u[i];
    for (int t = A.term_start[p]; t < A.term_start[p + 1]; ++t) {
      // ((D * (val[adj] - val[i]) * s_ij) / d_ij) / area   (calcPDE.cpp:1532-1533)
      const double diff = dsub(u[A.adj[t]], ui);
      acc = dadd(acc, ddiv(ddiv(dmul(dmul(D, diff), A.s[t]), A.d[t]), area));
    }
  } else {
    const double wi = dadd(u[i], dmul(A.cdt, kprev[i]));
    for (int t = A.term_start[p]; t < A.term_start[p + 1]; ++t) {
      // D * (((w[adj] - w[i]) * s_ij) / d_ij) / area   (calcPDE.cpp:1554-1557)
      const int j = A.adj[t];
      const double wj = dadd(u[j], dmul(A.cdt, kprev[j]));
      acc = dadd(acc, ddiv(dmul(D, ddiv(dmul(dsub(wj, wi), A.s[t]), A.d[t])), area));
    }
  }
  A.v.k[A.id][A.m][i] = acc;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mem_update_kernel(int n, const int32_t* pts, double* u, double* k0, double* k1, double* k2,
                                                         double* k3, double dt) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int i = pts[p];
  // value += dt * (k0 + 2.0*k1 + 2.0*k2 + k3) / 6.0, then the four slopes are zeroed (spatialsimulator.cpp:1535-1536)
  u[i] = dadd(u[i], ddiv(dmul(dt, dadd(dadd(dadd(k0[i], dmul(2.0, k1[i])), dmul(2.0, k2[i])), k3[i])), 6.0));
  k0[i] = 0.0; k1[i] = 0.0; k2[i] = 0.0; k3[i] = 0.0;
}

// sources are membrane points (isDomain == 1), targets pseudo-membrane points (isDomain == 2): entries are independent
__global__ void __launch_bounds__(128) mem_fill_kernel(int n, const int32_t* target, const int32_t* a, const int32_t* b, double* u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = u[a[i]], y = u[b[i]];
  u[target[i]] = (y < x) ? y : x;  // std::min(x, y)
}

__global__ void __launch_bounds__(256) mem_min_dt_kernel(int npts, const int32_t* term_start, const double* d, const double* dvalues,
                                                         int dkind, double dconst, double dt, double* block_min) {
  double best = dt;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += gridDim.x * blockDim.x) {
    const double D = dkind == SB2_SRC_FIELD ? dvalues[p] : dconst;
    for (int t = term_start[p]; t < term_start[p + 1]; ++t) {
      const double lim = ddiv(dmul(d[t], d[t]), dmul(2.0, D));  // pow(d_ij, 2) / (2.0 * D)
      if (dt >= lim) best = fmin(best, lim);
    }
  }
  for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
  __shared__ double wmin[8];
  if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? wmin[threadIdx.x] : dt;
    for (int o = 4; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) block_min[blockIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cell_rule_kernel(const __grid_constant__ Sb2CellRuleArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, k = blockIdx.z;
  if (i >= A.nx) return;
  const int64_t idx = A.first + (int64_t)k * A.PL + (int64_t)j * A.P + i;
  if (A.mask && !(A.mask[idx] & SB2_F_DOMAIN)) return;
  double st[50];
  for (int q = 0; q < 50; ++q) st[q] = 0.0;
  int sp = eval_stream(A.tok, A.ntoks, A.consts, st, [&](int, int kind, int slot) -> double {
    return kind == SB2_TOK_VAR ? A.species[slot][idx] : A.fields[slot][idx];  // raw values, no stage offset (calcPDE.cpp:39-40)
  });
  if (sp > 0) A.out[idx] = st[sp - 1];  // "if (st_index > 0) value[index] = rpStack[--st_index]" (calcPDE.cpp:218-219)
}

}  // namespace

// =============================================================================================
int sb2_build_mem_reaction(int set, int kind, const sb2_token* toks, int ntoks, const sb2_ref* refs, int nrefs, int n_reactants,
                           const int32_t* points, int npts, const double* operand_values, int n_mem_species, int set_npts,
                           cudaStream_t stream, Sb2MemReaction& out, std::string& err) {
  out = Sb2MemReaction();
  out.set = set; out.kind = kind; out.n_reactants = n_reactants; out.npts = npts; out.ntoks = ntoks; out.nrefs = 0;
  out.d_tok = NULL; out.d_operand = NULL; out.d_points = NULL; out.d_opvalues = NULL;
  if (npts < 0 || ntoks < 0 || (npts > 0 && !points) || (ntoks > 0 && !toks)) { err = "membrane reaction: null argument"; return SB2_ERR_ARG; }
  for (int p = 0; p < npts; ++p)
    if (points[p] < 0 || points[p] >= set_npts) { err = "membrane reaction: point outside its point set"; return SB2_ERR_ARG; }
  std::vector<int32_t> operand(std::max(ntoks, 1), -1);
  int rows = 0;
  for (int i = 0; i < ntoks; ++i) {
    out.tok.push_back((uint32_t)toks[i].kind | ((uint32_t)toks[i].op << 8) | ((uint32_t)toks[i].slot << 16));
    if (toks[i].kind == SB2_TOK_MEMVAR) {
      if (toks[i].slot >= n_mem_species) { err = "membrane reaction: token references unknown membrane species"; return SB2_ERR_ARG; }
    } else if (toks[i].kind == SB2_TOK_VAR || toks[i].kind == SB2_TOK_FIELD) {
      if (!operand_values) { err = "membrane reaction: operand values missing"; return SB2_ERR_ARG; }
      operand[i] = rows++;
    } else if (toks[i].kind == SB2_TOK_CONST && toks[i].slot >= SB2_MAX_CONSTS) {
      err = "membrane reaction: constant slot out of range"; return SB2_ERR_ARG;
    }
  }
  // reactants first, then products (calcPDE.cpp:432-446); a rate rule adds to its first reference only
  for (int r = 0; r < nrefs; ++r) {
    if (!refs[r].is_variable || refs[r].species < SB2_MEM_SPECIES) continue;
    const int id = refs[r].species - SB2_MEM_SPECIES;
    if (id >= n_mem_species) { err = "membrane reaction: reference to unknown membrane species"; return SB2_ERR_ARG; }
    if (kind == 1 && r != 0) continue;
    if (out.nrefs >= 8) { err = "membrane reaction: more than 8 variable species references"; return SB2_ERR_LIMIT; }
    out.ref_species[out.nrefs] = id;
    out.ref_sign[out.nrefs] = (kind == 0 && r < n_reactants) ? -1 : +1;
    out.ref_stoich[out.nrefs] = refs[r].stoichiometry;
    ++out.nrefs;
  }
  std::vector<uint32_t> tokv(out.tok);
  if (tokv.empty()) tokv.push_back(SB2_TOK_NOP);
  cudaError_t e;
  if ((e = to_device(&out.d_tok, tokv.data(), tokv.size(), stream)) != cudaSuccess ||
      (e = to_device(&out.d_operand, operand.data(), operand.size(), stream)) != cudaSuccess ||
      (e = to_device(&out.d_points, points, (size_t)npts, stream)) != cudaSuccess ||
      (e = to_device(&out.d_opvalues, operand_values, (size_t)rows * npts, stream)) != cudaSuccess) {
    err = std::string("membrane reaction: ") + cudaGetErrorString(e);
    return SB2_ERR_CUDA;
  }
  return SB2_OK;
}

void sb2_free_mem_reaction(Sb2MemReaction& r) {
  cudaFree(r.d_tok); cudaFree(r.d_operand); cudaFree(r.d_points); cudaFree(r.d_opvalues);
}

cudaError_t sb2_launch_mem_reaction(const Sb2MemReaction& r, const Sb2MemView& v, const double* consts, int m, double cdt,
                                    cudaStream_t stream, int* launches) {
  if (r.npts == 0 || r.nrefs == 0) return cudaSuccess;  // a law that changes nothing is not evaluated
  MemReactArgs a;
  a.npts = r.npts; a.ntoks = r.ntoks; a.m = m; a.kind = r.kind; a.nrefs = r.nrefs; a.set_npts = 0; a.cdt = cdt;
  a.tok = r.d_tok; a.operand = r.d_operand; a.points = r.d_points; a.opvalues = r.d_opvalues;
  for (int i = 0; i < 8; ++i) { a.ref_species[i] = r.ref_species[i]; a.ref_sign[i] = r.ref_sign[i]; a.ref_stoich[i] = r.ref_stoich[i]; }
  a.v = v;
  memcpy(a.consts, consts, sizeof(a.consts));
  if (launches) ++*launches;
  mem_react_kernel<<<(r.npts + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t sb2_launch_mem_diffusion(const Sb2MemSpecies& s, int id, const Sb2MemView& v, const double* consts, int m, double cdt,
                                     int dim, cudaStream_t stream, int* launches) {
  (void)dim;
  if (!s.has_diff || s.n_diff_pts == 0) return cudaSuccess;
  MemDiffArgs a;
  a.npts = s.n_diff_pts; a.m = m; a.id = id; a.dkind = s.dkind; a.cdt = cdt;
  a.dconst = s.dkind == SB2_SRC_CONST ? consts[s.dsrc] : 0.0;
  a.pts = s.d_diff_pts; a.term_start = s.d_term_start; a.adj = s.d_adj; a.s = s.d_s; a.d = s.d_d; a.area = s.d_area;
  a.dvalues = s.d_dvalues;
  a.v = v;
  if (launches) ++*launches;
  mem_diffusion_kernel<<<(a.npts + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t sb2_launch_mem_update(const Sb2MemSpecies& s, int id, const Sb2MemView& v, double dt, cudaStream_t stream, int* launches) {
  if (s.n_update == 0) return cudaSuccess;
  if (launches) ++*launches;
  mem_update_kernel<<<(s.n_update + 127) / 128, 128, 0, stream>>>(s.n_update, s.d_update_pts, v.u[id], v.k[id][0], v.k[id][1], v.k[id][2],
                                                                 v.k[id][3], dt);
  return cudaGetLastError();
}

cudaError_t sb2_launch_mem_fill(const Sb2PointSet& set, double* u, cudaStream_t stream, int* launches) {
  if (set.n_fill == 0) return cudaSuccess;
  if (launches) ++*launches;
  mem_fill_kernel<<<(set.n_fill + 127) / 128, 128, 0, stream>>>(set.n_fill, set.d_fill_target, set.d_fill_a, set.d_fill_b, u);
  return cudaGetLastError();
}

cudaError_t sb2_launch_mem_min_dt(const Sb2MemSpecies& s, const double* consts, double dt, double* d_scratch, cudaStream_t stream,
                                  int* launches, double* result) {
  *result = dt;
  if (!s.has_diff || s.n_diff_pts == 0) return cudaSuccess;
  const int blocks = std::max(1, std::min((s.n_diff_pts + 255) / 256, 1024));
  if (launches) ++*launches;
  mem_min_dt_kernel<<<blocks, 256, 0, stream>>>(s.n_diff_pts, s.d_term_start, s.d_d, s.d_dvalues, s.dkind,
                                                s.dkind == SB2_SRC_CONST ? consts[s.dsrc] : 0.0, dt, d_scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  std::vector<double> h(blocks);
  if ((e = cudaMemcpyAsync(h.data(), d_scratch, sizeof(double) * blocks, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
  for (int i = 0; i < blocks; ++i) *result = std::min(*result, h[i]);
  return cudaSuccess;
}

int sb2_build_cell_rule(const sb2_token* toks, int ntoks, int target_kind, int target, int geo, cudaStream_t stream, Sb2CellRule& out,
                        std::string& err) {
  out = Sb2CellRule();
  out.ntoks = ntoks; out.target_kind = target_kind; out.target = target; out.geo = geo; out.d_tok = NULL;
  for (int i = 0; i < ntoks; ++i) out.tok.push_back((uint32_t)toks[i].kind | ((uint32_t)toks[i].op << 8) | ((uint32_t)toks[i].slot << 16));
  std::vector<uint32_t> tokv(out.tok);
  if (tokv.empty()) tokv.push_back(SB2_TOK_NOP);
  cudaError_t e = to_device(&out.d_tok, tokv.data(), tokv.size(), stream);
  if (e != cudaSuccess) { err = std::string("assignment rule: ") + cudaGetErrorString(e); return SB2_ERR_CUDA; }
  return SB2_OK;
}

void sb2_free_cell_rule(Sb2CellRule& r) { cudaFree(r.d_tok); }

cudaError_t sb2_launch_cell_rule(const Sb2CellRuleArgs& a, cudaStream_t stream, int* launches) {
  dim3 block(256), grid((a.nx + 255) / 256, a.ny, a.nz);
  if (launches) ++*launches;
  cell_rule_kernel<<<grid, block, 0, stream>>>(a);
  return cudaGetLastError();
}
