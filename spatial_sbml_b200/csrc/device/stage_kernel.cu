// stage_kernel.cu — the fused RK-stage kernel (sm_100a).
//
// One launch computes, for every volume cell of the slab and every staged species,
//     k_m = sum over reactions / rate rules (bytecode interpreter)  +  diffusion stencil
// and, on the last stage, the RK4 combine u_new = u + dt*(k0 + 2k1 + 2k2 + k3)/6 — the work the
// reference spreads over reversePolishRK (calcPDE.cpp:224-451), calcDiffusion (calcPDE.cpp:453-562)
// and updateSpecies (spatialsimulator.cpp:1518-1544), one full-grid pass per reaction and per
// species.  Here each species field is read once per stage.
//
// Layout and schedule
//   * A CTA owns a strip of TX = 256 consecutive x cells and marches along y.  Each thread owns two
//     adjacent cells (one 16-byte double2 per array per row → fully coalesced 4 KB row segments).
//   * The stage input w = u + rk[m]*dt*k_{m-1} is computed ONCE per cell when its row is loaded and
//     kept in a 4-deep shared-memory ring of rows (+ two halo columns); the 5-point stencil and the
//     interpreter's species operands read it from there.  Ring depth 4 needs one barrier per row.
//   * z neighbours (3-D) are read straight from global memory (guard planes make them unconditional).
//   * The token stream and the constants table arrive as a __grid_constant__ kernel parameter, i.e.
//     in constant memory; the program counter is warp-uniform, so every token fetch is a uniform
//     constant load and every dispatch a uniform indexed branch.
//   * The expression stack lives in registers: tokens carry their (statically known) stack slot, so
//     each interpreter case uses compile-time register indices (see sb2_internal.h).
//
// Arithmetic follows the reference's evaluation order with explicit round-to-nearest intrinsics
// (never contracted into FMAs), so results are bit-identical to the CPU code for + - * / and
// comparisons.
#include "sb2_internal.h"
#include "sb2_ops.cuh"

namespace {

constexpr int NT = 128;       // threads per CTA
constexpr int TX = 2 * NT;    // cells per strip row
constexpr int TXP = TX + 2;   // + left halo (index TX) and right halo (index TX+1)
constexpr int RING = 4;
constexpr int D = SB2_MAX_DEPTH;

#define REP_D0(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)
#define REP_D1(M) M(1) M(2) M(3) M(4) M(5) M(6) M(7)

template <int NS>
__global__ void __launch_bounds__(NT) stage_kernel(const __grid_constant__ Sb2StageArgs A) {
  extern __shared__ __align__(16) double wsm[];  // [NS][RING][TXP]

  const int tx = threadIdx.x;
  const int x0 = blockIdx.x * TX;
  const int x = x0 + 2 * tx;
  const bool active = x < A.nx;
  const int S = A.S;
  const int m = A.m;

  int z = 0, yb, ye;
  if (A.dim == 3) {
    z = A.row_begin + blockIdx.z;
    yb = blockIdx.y * A.rows_per_block;
    ye = min(yb + A.rows_per_block, A.ny);
  } else {
    yb = A.row_begin + blockIdx.y * A.rows_per_block;
    ye = min(yb + A.rows_per_block, A.row_end);
  }
  const int64_t plane0 = A.first + (int64_t)z * A.PL;
  const int idxL = (tx == 0) ? TX : 2 * tx - 1;            // smem index of the cell left of cell 0
  const int idxR = (tx == NT - 1) ? TX + 1 : 2 * tx + 2;   // smem index of the cell right of cell 1

  auto stage_value = [&](double u, double k) -> double {
    // u + ((rk[m]*dt) * k), calcPDE.cpp:247-248
    return (m == 0) ? u : dadd(u, dmul(A.cdt, k));
  };

  // load one row of every species, form the stage input and park it in the ring
  auto load_row = [&](int yy) {
    const int slot = yy & (RING - 1);
    const int64_t base = plane0 + (int64_t)yy * A.P;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < S) {
        double* row = wsm + (s * RING + slot) * TXP;
        double2 w = make_double2(0.0, 0.0);
        if (active) {
          const double2 u = *reinterpret_cast<const double2*>(A.sp[s].u + base + x);
          double2 k = make_double2(0.0, 0.0);
          if (m > 0) k = *reinterpret_cast<const double2*>(A.sp[s].kprev + base + x);
          w.x = stage_value(u.x, k.x);
          w.y = stage_value(u.y, k.y);
        }
        *reinterpret_cast<double2*>(row + 2 * tx) = w;
        if (tx == 0) {
          double wl = 0.0;
          if (x0 > 0) {
            const double u = A.sp[s].u[base + x0 - 1];
            const double k = (m > 0) ? A.sp[s].kprev[base + x0 - 1] : 0.0;
            wl = stage_value(u, k);
          }
          row[TX] = wl;
        }
        if (tx == NT - 1) {
          double wr = 0.0;
          if (x0 + TX < A.P) {
            const double u = A.sp[s].u[base + x0 + TX];
            const double k = (m > 0) ? A.sp[s].kprev[base + x0 + TX] : 0.0;
            wr = stage_value(u, k);
          }
          row[TX + 1] = wr;
        }
      }
    }
  };

  // packed per-cell flag bytes of all geometries (byte g = geometry g)
  auto load_flags = [&](int yy, uint32_t& f0, uint32_t& f1) {
    f0 = 0; f1 = 0;
    if (!active) return;
    const int64_t base = plane0 + (int64_t)yy * A.P + x;
#pragma unroll
    for (int g = 0; g < SB2_MAX_GEOS; ++g) {
      if (g < A.G) {
        const uchar2 f = *reinterpret_cast<const uchar2*>(A.flags[g] + base);
        f0 |= (uint32_t)f.x << (8 * g);
        f1 |= (uint32_t)f.y << (8 * g);
      }
    }
  };

  load_row(yb - 1);
  load_row(yb);
  uint32_t fl0, fl1;
  load_flags(yb, fl0, fl1);

  for (int y = yb; y < ye; ++y) {
    load_row(y + 1);
    uint32_t fn0, fn1;
    load_flags(y + 1, fn0, fn1);
    __syncthreads();

    if (active && ((fl0 | fl1) & 0x01010101u)) {
      const int slot = y & (RING - 1);
      const int64_t idx = plane0 + (int64_t)y * A.P + x;

      double acc[NS][2];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        acc[s][0] = 0.0; acc[s][1] = 0.0;
        if (A.carry && s < S) {
          const double2 c = *reinterpret_cast<const double2*>(A.sp[s].kout + idx);
          acc[s][0] = c.x; acc[s][1] = c.y;
        }
      }

      // ---------------- reaction / rate-rule bytecode ----------------
      double r[D][2];
#pragma unroll
      for (int d = 0; d < D; ++d) { r[d][0] = 0.0; r[d][1] = 0.0; }
      bool mk0 = false, mk1 = false;
      const double* wrow = wsm + slot * TXP + 2 * tx;  // + s*RING*TXP selects the species

      for (int pc = 0; pc < A.ntok; ++pc) {
        const uint32_t t = A.tok[pc];
        const uint32_t arg = t >> 16;
        switch (t & 0xffffu) {
#define C_PUSHC(d) case SB2_CODE(OPC_PUSHC, d): { const double c = A.consts[arg]; r[d][0] = c; r[d][1] = c; } break;
          REP_D0(C_PUSHC)
#define C_PUSHV(d) case SB2_CODE(OPC_PUSHV, d): { const double2 v = *reinterpret_cast<const double2*>(wrow + arg * (RING * TXP)); r[d][0] = v.x; r[d][1] = v.y; } break;
          REP_D0(C_PUSHV)
#define C_PUSHF(d) case SB2_CODE(OPC_PUSHF, d): { const double2 v = *reinterpret_cast<const double2*>(A.fields[arg] + idx); r[d][0] = v.x; r[d][1] = v.y; } break;
          REP_D0(C_PUSHF)
#define C_BIN(opc, d, EXPR) case SB2_CODE(opc, d): { { const double a = r[d - 1][0], b = r[d][0]; r[d - 1][0] = (EXPR); } { const double a = r[d - 1][1], b = r[d][1]; r[d - 1][1] = (EXPR); } } break;
#define C_ADD(d) C_BIN(OPC_ADD, d, dadd(a, b))
#define C_SUB(d) C_BIN(OPC_SUB, d, dsub(a, b))
#define C_MUL(d) C_BIN(OPC_MUL, d, dmul(a, b))
#define C_DIV(d) C_BIN(OPC_DIV, d, ddiv(a, b))
#define C_GEQ(d) C_BIN(OPC_GEQ, d, (a >= b) ? 1.0 : 0.0)
#define C_GT(d) C_BIN(OPC_GT, d, (a > b) ? 1.0 : 0.0)
#define C_LEQ(d) C_BIN(OPC_LEQ, d, (a <= b) ? 1.0 : 0.0)
#define C_LT(d) C_BIN(OPC_LT, d, (a < b) ? 1.0 : 0.0)
#define C_AND(d) C_BIN(OPC_AND, d, logic_and(a, b))
#define C_OR(d) C_BIN(OPC_OR, d, logic_or(a, b))
          REP_D1(C_ADD) REP_D1(C_SUB) REP_D1(C_MUL) REP_D1(C_DIV)
          REP_D1(C_GEQ) REP_D1(C_GT) REP_D1(C_LEQ) REP_D1(C_LT) REP_D1(C_AND) REP_D1(C_OR)
#define C_BINC(opc, d, EXPR) case SB2_CODE(opc, d): { const double b = A.consts[arg]; { const double a = r[d][0]; r[d][0] = (EXPR); } { const double a = r[d][1]; r[d][1] = (EXPR); } } break;
#define C_ADDC(d) C_BINC(OPC_ADDC, d, dadd(a, b))
#define C_SUBC(d) C_BINC(OPC_SUBC, d, dsub(a, b))
#define C_MULC(d) C_BINC(OPC_MULC, d, dmul(a, b))
#define C_DIVC(d) C_BINC(OPC_DIVC, d, ddiv(a, b))
#define C_GEQC(d) C_BINC(OPC_GEQC, d, (a >= b) ? 1.0 : 0.0)
#define C_GTC(d) C_BINC(OPC_GTC, d, (a > b) ? 1.0 : 0.0)
#define C_LEQC(d) C_BINC(OPC_LEQC, d, (a <= b) ? 1.0 : 0.0)
#define C_LTC(d) C_BINC(OPC_LTC, d, (a < b) ? 1.0 : 0.0)
          REP_D0(C_ADDC) REP_D0(C_SUBC) REP_D0(C_MULC) REP_D0(C_DIVC)
          REP_D0(C_GEQC) REP_D0(C_GTC) REP_D0(C_LEQC) REP_D0(C_LTC)
#define C_BINV(opc, d, EXPR) case SB2_CODE(opc, d): { const double2 v = *reinterpret_cast<const double2*>(wrow + arg * (RING * TXP)); { const double a = r[d][0], b = v.x; r[d][0] = (EXPR); } { const double a = r[d][1], b = v.y; r[d][1] = (EXPR); } } break;
#define C_ADDV(d) C_BINV(OPC_ADDV, d, dadd(a, b))
#define C_SUBV(d) C_BINV(OPC_SUBV, d, dsub(a, b))
#define C_MULV(d) C_BINV(OPC_MULV, d, dmul(a, b))
#define C_DIVV(d) C_BINV(OPC_DIVV, d, ddiv(a, b))
          REP_D0(C_ADDV) REP_D0(C_SUBV) REP_D0(C_MULV) REP_D0(C_DIVV)
#define C_UNARY(d) case SB2_CODE(OPC_UNARY, d): { r[d][0] = sb2_unary((int)arg, r[d][0]); r[d][1] = sb2_unary((int)arg, r[d][1]); } break;
          REP_D0(C_UNARY)
#define C_BINR(d) case SB2_CODE(OPC_BINR, d): { r[d - 1][0] = sb2_binary((int)arg, r[d - 1][0], r[d][0]); r[d - 1][1] = sb2_binary((int)arg, r[d - 1][1], r[d][1]); } break;
          REP_D1(C_BINR)
#define C_MOV0(d) case SB2_CODE(OPC_MOV0, d): { r[0][0] = r[d][0]; r[0][1] = r[d][1]; } break;
          REP_D1(C_MOV0)
          case SB2_CODE(OPC_MASK, 0): {
            mk0 = (fl0 >> (8 * arg)) & 1u;
            mk1 = (fl1 >> (8 * arg)) & 1u;
          } break;
          // delta -= stoich*rate (reactants) / += (products, rate rules); calcPDE.cpp:432-449
#define C_ACC(s) \
  case SB2_CODE(OPC_ACCSUB, s): if (s < NS) { const double st = A.consts[arg]; \
      if (mk0) acc[s < NS ? s : 0][0] = dsub(acc[s < NS ? s : 0][0], dmul(st, r[0][0])); \
      if (mk1) acc[s < NS ? s : 0][1] = dsub(acc[s < NS ? s : 0][1], dmul(st, r[0][1])); } break; \
  case SB2_CODE(OPC_ACCADD, s): if (s < NS) { const double st = A.consts[arg]; \
      if (mk0) acc[s < NS ? s : 0][0] = dadd(acc[s < NS ? s : 0][0], dmul(st, r[0][0])); \
      if (mk1) acc[s < NS ? s : 0][1] = dadd(acc[s < NS ? s : 0][1], dmul(st, r[0][1])); } break;
          REP_D0(C_ACC)
          default: break;
        }
      }

      // ---------------- diffusion (calcPDE.cpp:484-559) + output ----------------
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (s < S) {
          const Sb2SpeciesArgs& sp = A.sp[s];
          const uint32_t f0 = (fl0 >> (8 * sp.geo)) & 0xffu;
          const uint32_t f1 = (fl1 >> (8 * sp.geo)) & 0xffu;
          const double* base = wsm + s * RING * TXP;
          if (sp.has_diff) {
            const double* cur = base + slot * TXP;
            const double2 wc = *reinterpret_cast<const double2*>(cur + 2 * tx);
            if (sp.dkind[0] != SB2_SRC_NONE) {
              double d0, d1;
              if (sp.dkind[0] == SB2_SRC_CONST) { d0 = d1 = A.consts[sp.dsrc[0]]; }
              else { const double2 dv = *reinterpret_cast<const double2*>(A.fields[sp.dsrc[0]] + idx); d0 = dv.x; d1 = dv.y; }
              const double wl = cur[idxL], wr = cur[idxR];
              auto term = [&](double dc, double wn, double w0) {
                double t = dmul(dc, dsub(wn, w0));
                if (A.divide[0]) t = ddiv(t, A.dx2[0]);
                return t;
              };
              if (f0 & SB2_F_DOMAIN) {
                if (!(f0 & SB2_F_XP)) acc[s][0] = dadd(acc[s][0], term(d0, wc.y, wc.x));
                if (!(f0 & SB2_F_XM)) acc[s][0] = dadd(acc[s][0], term(d0, wl, wc.x));
              }
              if (f1 & SB2_F_DOMAIN) {
                if (!(f1 & SB2_F_XP)) acc[s][1] = dadd(acc[s][1], term(d1, wr, wc.y));
                if (!(f1 & SB2_F_XM)) acc[s][1] = dadd(acc[s][1], term(d1, wc.x, wc.y));
              }
            }
            if (sp.dkind[1] != SB2_SRC_NONE) {
              double d0, d1;
              if (sp.dkind[1] == SB2_SRC_CONST) { d0 = d1 = A.consts[sp.dsrc[1]]; }
              else { const double2 dv = *reinterpret_cast<const double2*>(A.fields[sp.dsrc[1]] + idx); d0 = dv.x; d1 = dv.y; }
              const double2 wu = *reinterpret_cast<const double2*>(base + ((y + 1) & (RING - 1)) * TXP + 2 * tx);
              const double2 wd = *reinterpret_cast<const double2*>(base + ((y - 1) & (RING - 1)) * TXP + 2 * tx);
              auto term = [&](double dc, double wn, double w0) {
                double t = dmul(dc, dsub(wn, w0));
                if (A.divide[1]) t = ddiv(t, A.dx2[1]);
                return t;
              };
              if (f0 & SB2_F_DOMAIN) {
                if (!(f0 & SB2_F_YP)) acc[s][0] = dadd(acc[s][0], term(d0, wu.x, wc.x));
                if (!(f0 & SB2_F_YM)) acc[s][0] = dadd(acc[s][0], term(d0, wd.x, wc.x));
              }
              if (f1 & SB2_F_DOMAIN) {
                if (!(f1 & SB2_F_YP)) acc[s][1] = dadd(acc[s][1], term(d1, wu.y, wc.y));
                if (!(f1 & SB2_F_YM)) acc[s][1] = dadd(acc[s][1], term(d1, wd.y, wc.y));
              }
            }
            if (A.dim == 3 && sp.dkind[2] != SB2_SRC_NONE) {
              double d0, d1;
              if (sp.dkind[2] == SB2_SRC_CONST) { d0 = d1 = A.consts[sp.dsrc[2]]; }
              else { const double2 dv = *reinterpret_cast<const double2*>(A.fields[sp.dsrc[2]] + idx); d0 = dv.x; d1 = dv.y; }
              const double2 uu = *reinterpret_cast<const double2*>(sp.u + idx + A.PL);
              const double2 ud = *reinterpret_cast<const double2*>(sp.u + idx - A.PL);
              double2 ku = make_double2(0.0, 0.0), kd = ku;
              if (m > 0) {
                ku = *reinterpret_cast<const double2*>(sp.kprev + idx + A.PL);
                kd = *reinterpret_cast<const double2*>(sp.kprev + idx - A.PL);
              }
              auto term = [&](double dc, double wn, double w0) {
                double t = dmul(dc, dsub(wn, w0));
                if (A.divide[2]) t = ddiv(t, A.dx2[2]);
                return t;
              };
              if (f0 & SB2_F_DOMAIN) {
                if (!(f0 & SB2_F_ZP)) acc[s][0] = dadd(acc[s][0], term(d0, stage_value(uu.x, ku.x), wc.x));
                if (!(f0 & SB2_F_ZM)) acc[s][0] = dadd(acc[s][0], term(d0, stage_value(ud.x, kd.x), wc.x));
              }
              if (f1 & SB2_F_DOMAIN) {
                if (!(f1 & SB2_F_ZP)) acc[s][1] = dadd(acc[s][1], term(d1, stage_value(uu.y, ku.y), wc.y));
                if (!(f1 & SB2_F_ZM)) acc[s][1] = dadd(acc[s][1], term(d1, stage_value(ud.y, kd.y), wc.y));
              }
            }
          }

          const bool in0 = f0 & SB2_F_DOMAIN, in1 = f1 & SB2_F_DOMAIN;
          if (!A.final) {
            double2 o;
            o.x = in0 ? acc[s][0] : 0.0;
            o.y = in1 ? acc[s][1] : 0.0;
            *reinterpret_cast<double2*>(sp.kout + idx) = o;
          } else {
            // value += dt*(k0 + 2.0*k1 + 2.0*k2 + k3)/6.0, spatialsimulator.cpp:1535
            const double2 u = *reinterpret_cast<const double2*>(sp.u + idx);
            const double2 k0 = *reinterpret_cast<const double2*>(sp.k0 + idx);
            const double2 k1 = *reinterpret_cast<const double2*>(sp.k1 + idx);
            const double2 k2 = *reinterpret_cast<const double2*>(sp.kprev + idx);
            double2 o = u;
            if (in0) o.x = dadd(u.x, ddiv(dmul(A.dt, dadd(dadd(dadd(k0.x, dmul(2.0, k1.x)), dmul(2.0, k2.x)), acc[s][0])), 6.0));
            if (in1) o.y = dadd(u.y, ddiv(dmul(A.dt, dadd(dadd(dadd(k0.y, dmul(2.0, k1.y)), dmul(2.0, k2.y)), acc[s][1])), 6.0));
            *reinterpret_cast<double2*>(sp.u_out + idx) = o;
          }
        }
      }
    } else if (active) {
      // whole pair outside every domain: k_m = 0 / u unchanged
      const int64_t idx = plane0 + (int64_t)y * A.P + x;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (s < S) {
          if (!A.final) {
            if (!A.carry) *reinterpret_cast<double2*>(A.sp[s].kout + idx) = make_double2(0.0, 0.0);
          } else {
            *reinterpret_cast<double2*>(A.sp[s].u_out + idx) = *reinterpret_cast<const double2*>(A.sp[s].u + idx);
          }
        }
      }
    }
    fl0 = fn0; fl1 = fn1;
  }
}

template <int NS>
cudaError_t launch_ns(const Sb2StageArgs& a, cudaStream_t stream) {
  static bool configured = false;
  const size_t smem = (size_t)NS * RING * TXP * sizeof(double);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(stage_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid;
  grid.x = (a.nx + TX - 1) / TX;
  if (a.dim == 3) {
    grid.y = (a.ny + a.rows_per_block - 1) / a.rows_per_block;
    grid.z = a.row_end - a.row_begin;
  } else {
    grid.y = (a.row_end - a.row_begin + a.rows_per_block - 1) / a.rows_per_block;
    grid.z = 1;
  }
  if (grid.y == 0 || grid.z == 0) return cudaSuccess;
  stage_kernel<NS><<<grid, NT, smem, stream>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// stand-alone RK combine (used when the combine cannot be fused into stage 4, i.e. when advection
// runs between the last stage and the update; spatialsimulator.cpp:1484-1491, 1535-1536)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) combine_kernel(const __grid_constant__ Sb2CombineArgs A) {
  const int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y;
  const int z = blockIdx.z;
  if (x >= A.nx) return;
  const int64_t idx = A.first + (int64_t)z * A.PL + (int64_t)y * A.P + x;
  for (int s = 0; s < A.S; ++s) {
    const uchar2 f = *reinterpret_cast<const uchar2*>(A.flags[s] + idx);
    if (!((f.x | f.y) & SB2_F_DOMAIN)) continue;
    double2 u = *reinterpret_cast<const double2*>(A.u[s] + idx);
    const double2 k0 = *reinterpret_cast<const double2*>(A.k[s][0] + idx);
    const double2 k1 = *reinterpret_cast<const double2*>(A.k[s][1] + idx);
    const double2 k2 = *reinterpret_cast<const double2*>(A.k[s][2] + idx);
    const double2 k3 = *reinterpret_cast<const double2*>(A.k[s][3] + idx);
    if (f.x & SB2_F_DOMAIN) u.x = dadd(u.x, ddiv(dmul(A.dt, dadd(dadd(dadd(k0.x, dmul(2.0, k1.x)), dmul(2.0, k2.x)), k3.x)), 6.0));
    if (f.y & SB2_F_DOMAIN) u.y = dadd(u.y, ddiv(dmul(A.dt, dadd(dadd(dadd(k0.y, dmul(2.0, k1.y)), dmul(2.0, k2.y)), k3.y)), 6.0));
    *reinterpret_cast<double2*>(A.u[s] + idx) = u;
  }
}

}  // namespace

cudaError_t sb2_launch_stage(const Sb2StageArgs& a, cudaStream_t stream, int* launches) {
  if (a.row_end <= a.row_begin) return cudaSuccess;
  if (launches) ++*launches;
  if (a.S <= 2) return launch_ns<2>(a, stream);
  if (a.S <= 4) return launch_ns<4>(a, stream);
  return launch_ns<8>(a, stream);
}

cudaError_t sb2_launch_combine(const Sb2CombineArgs& a, cudaStream_t stream, int* launches) {
  dim3 block(256), grid(((a.nx + 1) / 2 + 255) / 256, a.ny, a.nz);
  if (launches) ++*launches;
  combine_kernel<<<grid, block, 0, stream>>>(a);
  return cudaGetLastError();
}
