// stage_kernel.cuh — the fused RK-stage kernel (sm_100a), included by one translation unit per
// species-count variant (stage_kernel_ns*.cu) so that the variants compile in parallel.
//
// One launch computes, for every volume cell of the slab and every staged species,
//     k_m = sum over reactions / rate rules (bytecode interpreter)  +  diffusion stencil
// and, on the last stage, the RK4 combine u_new = u + dt*(k0 + 2k1 + 2k2 + k3)/6 — the work the
// reference spreads over reversePolishRK (calcPDE.cpp:224-451), calcDiffusion (calcPDE.cpp:453-562)
// and updateSpecies (spatialsimulator.cpp:1518-1544), one full-grid pass per reaction and per
// species.  Here each species field is read once per stage.
//
// Layout and schedule
//   * A CTA (128 threads) owns a strip of TX = 2*NP*128 consecutive x cells and marches along y.
//     Each thread owns NP pairs of adjacent cells; pair p of thread t sits at x0 + p*256 + 2t, so
//     every 16-byte load / store instruction of a warp covers 512 contiguous bytes.
//   * The stage input w = u + rk[m]*dt*k_{m-1} is computed ONCE per cell and parked in a 4-deep
//     shared-memory ring of rows (+ one halo column either side); the 5-point stencil and the
//     interpreter's species operands read it from there.  The raw values of row y+2 are fetched into
//     registers before row y is computed and parked afterwards, so the HBM latency of a row hides
//     behind the arithmetic of another; the ring depth makes one barrier per row sufficient.
//   * z neighbours (3-D) are read straight from global memory (guard planes make them unconditional).
//   * The token stream and the constants table arrive as a __grid_constant__ kernel parameter, i.e.
//     in constant memory.  The program counter is warp-uniform: a token fetch is a uniform constant
//     load and the dispatch one indexed branch through a dense jump table.
//   * The expression stack lives in registers: tokens carry their (statically known) stack slot, so
//     each interpreter case uses compile-time register indices (see sb2_internal.h), and acts on all
//     2*NP cells of the thread at once, which is what amortises the dispatch.
//   * Stage position (first / middle / last+combine) is a template parameter.
//
// Arithmetic follows the reference's evaluation order with explicit round-to-nearest intrinsics
// (never contracted into FMAs), so results are bit-identical to the CPU code for + - * / and
// comparisons.
#ifndef SB2_STAGE_KERNEL_CUH_
#define SB2_STAGE_KERNEL_CUH_
#include "sb2_internal.h"
#include "sb2_ops.cuh"

namespace {

constexpr int NT = 128;  // threads per CTA
constexpr int RING = 4;

enum { MODE_FIRST = 0, MODE_MIDDLE = 1, MODE_FINAL = 2 };

#define REP_D0(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)
#define REP_D1(M) M(1) M(2) M(3) M(4) M(5) M(6) M(7)

template <int NS, int NP, int MODE, int DEPTH>
__global__ void __launch_bounds__(NT, (NS * NP <= 2) ? 5 : (NS * NP <= 4) ? 3 : 2) stage_kernel(const __grid_constant__ Sb2StageArgs A) {
  constexpr int TX = 2 * NP * NT;  // cells per strip row
  constexpr int TXP = TX + 4;      // [pad][left halo][TX cells][right halo][pad]; cell i at index i + 2 (16-byte aligned pairs)
  constexpr int NC = 2 * NP;       // cells per thread
  extern __shared__ __align__(16) double wsm[];  // [NS][RING][TXP]

  const int tx = threadIdx.x;
  const int x0 = blockIdx.x * TX;
  const int S = A.S;

  int z = 0, yb, ye;
  if (A.dim == 3) {
    z = A.row_begin + blockIdx.z;
    yb = blockIdx.y * A.rows_per_block;
    ye = min(yb + A.rows_per_block, A.ny);
  } else {
    yb = A.row_begin + blockIdx.y * A.rows_per_block;
    ye = min(yb + A.rows_per_block, A.row_end);
  }
  if (yb >= ye) return;
  const int64_t plane0 = A.first + (int64_t)z * A.PL;

  bool active[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) active[p] = x0 + p * 2 * NT + 2 * tx < A.nx;
  // threads 0 / 1 also fetch the halo column left / right of the strip
  const int hx = (tx == 0) ? x0 - 1 : x0 + TX;
  const bool hactive = (tx == 0) ? (x0 > 0) : (tx == 1 && x0 + TX < A.P);

  auto stage_value = [&](double u, double k) -> double {
    // u + ((rk[m]*dt) * k), calcPDE.cpp:247-248
    return (MODE == MODE_FIRST) ? u : dadd(u, dmul(A.cdt, k));
  };

  // ---- row staging ------------------------------------------------------------------------------
  double2 pu[NS][NP], pk[NS][NP];
  double hu[NS], hk[NS];
  auto fetch_row = [&](int yy) {
    const int64_t base = plane0 + (int64_t)yy * A.P;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < S) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          pu[s][p] = make_double2(0.0, 0.0);
          pk[s][p] = make_double2(0.0, 0.0);
          if (active[p]) {
            const int64_t i = base + x0 + p * 2 * NT + 2 * tx;
            pu[s][p] = *reinterpret_cast<const double2*>(A.sp[s].u + i);
            if (MODE != MODE_FIRST) pk[s][p] = *reinterpret_cast<const double2*>(A.sp[s].kprev + i);
          }
        }
        hu[s] = 0.0; hk[s] = 0.0;
        if (hactive) {
          hu[s] = A.sp[s].u[base + hx];
          if (MODE != MODE_FIRST) hk[s] = A.sp[s].kprev[base + hx];
        }
      }
    }
  };
  auto park_row = [&](int yy) {
    const int slot = yy & (RING - 1);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < S) {
        double* row = wsm + (s * RING + slot) * TXP;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          double2 w;
          w.x = stage_value(pu[s][p].x, pk[s][p].x);
          w.y = stage_value(pu[s][p].y, pk[s][p].y);
          *reinterpret_cast<double2*>(row + 2 + p * 2 * NT + 2 * tx) = w;
        }
        if (tx < 2) row[tx == 0 ? 1 : TX + 2] = stage_value(hu[s], hk[s]);
      }
    }
  };
  // packed per-cell flag bytes of all geometries (byte g = geometry g)
  uint32_t fl[NC], fn[NC];
  auto fetch_flags = [&](int yy, uint32_t* f) {
    const int64_t base = plane0 + (int64_t)yy * A.P + x0 + 2 * tx;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      f[2 * p] = 0; f[2 * p + 1] = 0;
      if (active[p]) {
#pragma unroll
        for (int g = 0; g < SB2_MAX_GEOS; ++g) {
          if (g < A.G) {
            const uchar2 v = *reinterpret_cast<const uchar2*>(A.flags[g] + base + p * 2 * NT);
            f[2 * p] |= (uint32_t)v.x << (8 * g);
            f[2 * p + 1] |= (uint32_t)v.y << (8 * g);
          }
        }
      }
    }
  };

  fetch_row(yb - 1); park_row(yb - 1);
  fetch_row(yb); park_row(yb);
  fetch_row(yb + 1); park_row(yb + 1);
  fetch_flags(yb, fl);
  __syncthreads();

  for (int y = yb; y < ye; ++y) {
    const bool more = y + 1 < ye;
    if (more) {
      fetch_row(y + 2);  // in flight while row y is computed; parked at the bottom of the loop
      fetch_flags(y + 1, fn);
    }

    uint32_t anyf = 0, allf = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < NC; ++c) { anyf |= fl[c]; allf &= fl[c]; }
    const bool mine = (anyf & 0x01010101u) != 0;
    // every cell of this thread is an interior cell of every geometry: flag byte == SB2_F_DOMAIN, no
    // boundary bit anywhere → the stencil and the output run without a single per-cell test
    const bool interior = (anyf == allf) && (anyf == (0x01010101u >> (8 * (4 - A.G))));
    // rows / warps entirely outside every domain skip the arithmetic (masked geometries)
    if (__any_sync(0xffffffffu, mine)) {
      const int slot = y & (RING - 1);
      const int64_t idx0 = plane0 + (int64_t)y * A.P + x0 + 2 * tx;  // + p*2*NT for pair p

      double acc[NS][NC];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          acc[s][2 * p] = 0.0; acc[s][2 * p + 1] = 0.0;
          if (A.carry && s < S && active[p]) {
            const double2 c = *reinterpret_cast<const double2*>(A.sp[s].kcarry + idx0 + p * 2 * NT);
            acc[s][2 * p] = c.x; acc[s][2 * p + 1] = c.y;
          }
        }
      }
      // ---------------- reaction / rate-rule bytecode ----------------
      double r[DEPTH][NC];  // RD(d): slot d of the expression stack; slots >= DEPTH do not exist in this variant
#define RD(d) r[(d) < DEPTH ? (d) : 0]
#define LIVE(d) if ((d) < DEPTH)
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int c = 0; c < NC; ++c) r[d][c] = 0.0;
      bool mk[NC];  // set by MASK, read by the ACC*M forms only
#pragma unroll
      for (int c = 0; c < NC; ++c) mk[c] = false;
      const double* wrow = wsm + slot * TXP + 2 + 2 * tx;  // + s*RING*TXP selects the species, + p*2*NT the pair

      uint32_t tnext = A.tok[0];
#pragma unroll 1
      for (int pc = 0; pc < A.ntok; ++pc) {
        const uint32_t t = tnext;
        tnext = A.tok[pc + 1];  // the stream ends with OPC_END, so this is always a valid fetch
        const uint32_t arg = t >> 16;
        switch (t & 0xffffu) {
#define LOADV(v, p) const double2 v = *reinterpret_cast<const double2*>(wrow + arg * (RING * TXP) + (p) * 2 * NT)
#define C_PUSHC(d) case SB2_CODE(OPC_PUSHC, d): LIVE(d) { const double c = A.consts[arg]; _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) RD(d)[c_] = c; } break;
          REP_D0(C_PUSHC)
#define C_PUSHV(d) case SB2_CODE(OPC_PUSHV, d): LIVE(d) { _Pragma("unroll") for (int p = 0; p < NP; ++p) { LOADV(v, p); RD(d)[2 * p] = v.x; RD(d)[2 * p + 1] = v.y; } } break;
          REP_D0(C_PUSHV)
#define C_PUSHF(d) case SB2_CODE(OPC_PUSHF, d): LIVE(d) { _Pragma("unroll") for (int p = 0; p < NP; ++p) { double2 v = make_double2(0.0, 0.0); if (active[p]) v = *reinterpret_cast<const double2*>(A.fields[arg] + idx0 + p * 2 * NT); RD(d)[2 * p] = v.x; RD(d)[2 * p + 1] = v.y; } } break;
          REP_D0(C_PUSHF)
#define C_BIN(opc, d, EXPR) case SB2_CODE(opc, d): LIVE(d) { _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) { const double a = RD(d - 1)[c_], b = RD(d)[c_]; RD(d - 1)[c_] = (EXPR); } } break;
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
#define C_BINC(opc, d, EXPR) case SB2_CODE(opc, d): LIVE(d) { const double b = A.consts[arg]; _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) { const double a = RD(d)[c_]; RD(d)[c_] = (EXPR); } } break;
#define C_ADDC(d) C_BINC(OPC_ADDC, d, dadd(a, b))
#define C_SUBC(d) C_BINC(OPC_SUBC, d, dsub(a, b))
#define C_MULC(d) C_BINC(OPC_MULC, d, dmul(a, b))
#define C_DIVC(d) C_BINC(OPC_DIVC, d, ddiv(a, b))
#define C_GEQC(d) C_BINC(OPC_GEQC, d, (a >= b) ? 1.0 : 0.0)
#define C_GTC(d) C_BINC(OPC_GTC, d, (a > b) ? 1.0 : 0.0)
#define C_LEQC(d) C_BINC(OPC_LEQC, d, (a <= b) ? 1.0 : 0.0)
#define C_LTC(d) C_BINC(OPC_LTC, d, (a < b) ? 1.0 : 0.0)
#define C_RSUBC(d) C_BINC(OPC_RSUBC, d, dsub(b, a))
#define C_RDIVC(d) C_BINC(OPC_RDIVC, d, ddiv(b, a))
          REP_D0(C_ADDC) REP_D0(C_SUBC) REP_D0(C_MULC) REP_D0(C_DIVC) REP_D0(C_RSUBC) REP_D0(C_RDIVC)
          REP_D0(C_GEQC) REP_D0(C_GTC) REP_D0(C_LEQC) REP_D0(C_LTC)
#define C_BINV(opc, d, EXPR) case SB2_CODE(opc, d): LIVE(d) { _Pragma("unroll") for (int p = 0; p < NP; ++p) { LOADV(v, p); { const double a = RD(d)[2 * p], b = v.x; RD(d)[2 * p] = (EXPR); } { const double a = RD(d)[2 * p + 1], b = v.y; RD(d)[2 * p + 1] = (EXPR); } } } break;
#define C_ADDV(d) C_BINV(OPC_ADDV, d, dadd(a, b))
#define C_SUBV(d) C_BINV(OPC_SUBV, d, dsub(a, b))
#define C_MULV(d) C_BINV(OPC_MULV, d, dmul(a, b))
#define C_DIVV(d) C_BINV(OPC_DIVV, d, ddiv(a, b))
          REP_D0(C_ADDV) REP_D0(C_SUBV) REP_D0(C_MULV) REP_D0(C_DIVV)
#define C_UNARY(d) case SB2_CODE(OPC_UNARY, d): LIVE(d) { _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) RD(d)[c_] = sb2_unary((int)arg, RD(d)[c_]); } break;
          REP_D0(C_UNARY)
#define C_BINR(d) case SB2_CODE(OPC_BINR, d): LIVE(d) { _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) RD(d - 1)[c_] = sb2_binary((int)arg, RD(d - 1)[c_], RD(d)[c_]); } break;
          REP_D1(C_BINR)
#define C_MOV0(d) case SB2_CODE(OPC_MOV0, d): LIVE(d) { _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) RD(0)[c_] = RD(d)[c_]; } break;
          REP_D1(C_MOV0)
#define C_MOVUP(d) case SB2_CODE(OPC_MOVUP, d): LIVE(d + 1) { _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) RD(d + 1)[c_] = RD(d)[c_]; } break;
          REP_D0(C_MOVUP)
          // statement mask: the law only acts on cells of its own domain (spatialsimulator.cpp:1381-1390)
          case SB2_CODE(OPC_MASK, 0): {
#pragma unroll
            for (int c = 0; c < NC; ++c) mk[c] = (fl[c] >> (8 * arg)) & 1u;
          } break;
          // delta -= stoich*rate (reactants) / += (products, rate rules); calcPDE.cpp:432-449
#define SI(s) ((s) < NS ? (s) : 0)
#define C_ACC(s) \
  case SB2_CODE(OPC_ACCSUB, s): if (s < NS) { const double st = A.consts[arg]; \
      _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) acc[SI(s)][c_] = dsub(acc[SI(s)][c_], dmul(st, RD(0)[c_])); } break; \
  case SB2_CODE(OPC_ACCADD, s): if (s < NS) { const double st = A.consts[arg]; \
      _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) acc[SI(s)][c_] = dadd(acc[SI(s)][c_], dmul(st, RD(0)[c_])); } break; \
  case SB2_CODE(OPC_ACCSUB1, s): if (s < NS) { \
      _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) acc[SI(s)][c_] = dsub(acc[SI(s)][c_], RD(0)[c_]); } break; \
  case SB2_CODE(OPC_ACCADD1, s): if (s < NS) { \
      _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) acc[SI(s)][c_] = dadd(acc[SI(s)][c_], RD(0)[c_]); } break; \
  case SB2_CODE(OPC_ACCSUBC1, s): if (s < NS) { const double cv = A.consts[arg]; \
      _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) acc[SI(s)][c_] = dsub(acc[SI(s)][c_], cv); } break; \
  case SB2_CODE(OPC_ACCADDC1, s): if (s < NS) { const double cv = A.consts[arg]; \
      _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) acc[SI(s)][c_] = dadd(acc[SI(s)][c_], cv); } break; \
  case SB2_CODE(OPC_ACCSUBM, s): if (s < NS) { const double st = A.consts[arg]; \
      _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) if (mk[c_]) acc[SI(s)][c_] = dsub(acc[SI(s)][c_], dmul(st, RD(0)[c_])); } break; \
  case SB2_CODE(OPC_ACCADDM, s): if (s < NS) { const double st = A.consts[arg]; \
      _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_) if (mk[c_]) acc[SI(s)][c_] = dadd(acc[SI(s)][c_], dmul(st, RD(0)[c_])); } break;
          REP_D0(C_ACC)
          // every code of the dense range gets a case, so that the switch compiles to one jump table
#define C_NONE0(opc) case SB2_CODE(opc, 0): break;
          C_NONE0(OPC_ADD) C_NONE0(OPC_SUB) C_NONE0(OPC_MUL) C_NONE0(OPC_DIV) C_NONE0(OPC_GEQ) C_NONE0(OPC_GT)
          C_NONE0(OPC_LEQ) C_NONE0(OPC_LT) C_NONE0(OPC_AND) C_NONE0(OPC_OR) C_NONE0(OPC_BINR) C_NONE0(OPC_MOV0)
#define C_NONE_END(d) case SB2_CODE(OPC_END, d): break;
          REP_D0(C_NONE_END)
#define C_NONE_MASK(d) case SB2_CODE(OPC_MASK, d): break;
          REP_D1(C_NONE_MASK)
          default: break;
        }
      }

      // fused combine inputs: issued once the expression stack is dead, in flight during the stencil
      double2 cu[NS][NP], ck0[NS][NP], ck1[NS][NP], ck2[NS][NP];
      if (MODE == MODE_FINAL) {
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            cu[s][p] = ck0[s][p] = ck1[s][p] = ck2[s][p] = make_double2(0.0, 0.0);
            if (s < S && active[p]) {
              const int64_t i = idx0 + p * 2 * NT;
              cu[s][p] = *reinterpret_cast<const double2*>(A.sp[s].u + i);
              ck0[s][p] = *reinterpret_cast<const double2*>(A.sp[s].k0 + i);
              ck1[s][p] = *reinterpret_cast<const double2*>(A.sp[s].k1 + i);
              ck2[s][p] = *reinterpret_cast<const double2*>(A.sp[s].kprev + i);
            }
          }
      }

      // ---------------- diffusion (calcPDE.cpp:484-559) + output ----------------
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (s < S) {
          const Sb2SpeciesArgs& sp = A.sp[s];
          const double* base = wsm + s * RING * TXP + 2 + 2 * tx;
          const double* cur = base + slot * TXP;
          const double* up = base + ((y + 1) & (RING - 1)) * TXP;
          const double* dn = base + ((y - 1) & (RING - 1)) * TXP;
          if (interior && (sp.fast2d || !sp.has_diff) && MODE != MODE_FINAL) {
            // straight-line 5-point stencil, additions in the reference's order Xp, Xm, Yp, Ym
            const double dcx = sp.has_diff ? A.consts[sp.dsrc[0]] : 0.0, dcy = sp.has_diff ? A.consts[sp.dsrc[1]] : 0.0;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              double a0 = acc[s][2 * p], a1 = acc[s][2 * p + 1];
              if (sp.has_diff) {
                const double2 wc = *reinterpret_cast<const double2*>(cur + p * 2 * NT);
                const double wl = cur[p * 2 * NT - 1], wr = cur[p * 2 * NT + 2];
                const double2 wu = *reinterpret_cast<const double2*>(up + p * 2 * NT);
                const double2 wd = *reinterpret_cast<const double2*>(dn + p * 2 * NT);
                a0 = dadd(a0, dmul(dcx, dsub(wc.y, wc.x)));
                a0 = dadd(a0, dmul(dcx, dsub(wl, wc.x)));
                a0 = dadd(a0, dmul(dcy, dsub(wu.x, wc.x)));
                a0 = dadd(a0, dmul(dcy, dsub(wd.x, wc.x)));
                a1 = dadd(a1, dmul(dcx, dsub(wr, wc.y)));
                a1 = dadd(a1, dmul(dcx, dsub(wc.x, wc.y)));
                a1 = dadd(a1, dmul(dcy, dsub(wu.y, wc.y)));
                a1 = dadd(a1, dmul(dcy, dsub(wd.y, wc.y)));
              }
              *reinterpret_cast<double2*>(sp.kout + idx0 + p * 2 * NT) = make_double2(a0, a1);
            }
            continue;
          }
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const int64_t idx = idx0 + p * 2 * NT;
            const uint32_t f0 = (fl[2 * p] >> (8 * sp.geo)) & 0xffu;
            const uint32_t f1 = (fl[2 * p + 1] >> (8 * sp.geo)) & 0xffu;
            double a0 = acc[s][2 * p], a1 = acc[s][2 * p + 1];
            if (sp.has_diff && ((f0 | f1) & SB2_F_DOMAIN)) {
              const double2 wc = *reinterpret_cast<const double2*>(cur + p * 2 * NT);
              // axis 0: x
              if (sp.dkind[0] != SB2_SRC_NONE) {
                double d0, d1;
                if (sp.dkind[0] == SB2_SRC_CONST) { d0 = d1 = A.consts[sp.dsrc[0]]; }
                else { const double2 dv = *reinterpret_cast<const double2*>(A.fields[sp.dsrc[0]] + idx); d0 = dv.x; d1 = dv.y; }
                const double wl = cur[p * 2 * NT - 1], wr = cur[p * 2 * NT + 2];
                auto term = [&](double dc, double wn, double w0) {
                  double t = dmul(dc, dsub(wn, w0));
                  if (A.divide[0]) t = ddiv_pos(t, A.dx2[0]);
                  return t;
                };
                if (f0 & SB2_F_DOMAIN) {
                  if (!(f0 & SB2_F_XP)) a0 = dadd(a0, term(d0, wc.y, wc.x));
                  if (!(f0 & SB2_F_XM)) a0 = dadd(a0, term(d0, wl, wc.x));
                }
                if (f1 & SB2_F_DOMAIN) {
                  if (!(f1 & SB2_F_XP)) a1 = dadd(a1, term(d1, wr, wc.y));
                  if (!(f1 & SB2_F_XM)) a1 = dadd(a1, term(d1, wc.x, wc.y));
                }
              }
              // axis 1: y
              if (sp.dkind[1] != SB2_SRC_NONE) {
                double d0, d1;
                if (sp.dkind[1] == SB2_SRC_CONST) { d0 = d1 = A.consts[sp.dsrc[1]]; }
                else { const double2 dv = *reinterpret_cast<const double2*>(A.fields[sp.dsrc[1]] + idx); d0 = dv.x; d1 = dv.y; }
                const double2 wu = *reinterpret_cast<const double2*>(up + p * 2 * NT);
                const double2 wd = *reinterpret_cast<const double2*>(dn + p * 2 * NT);
                auto term = [&](double dc, double wn, double w0) {
                  double t = dmul(dc, dsub(wn, w0));
                  if (A.divide[1]) t = ddiv_pos(t, A.dx2[1]);
                  return t;
                };
                if (f0 & SB2_F_DOMAIN) {
                  if (!(f0 & SB2_F_YP)) a0 = dadd(a0, term(d0, wu.x, wc.x));
                  if (!(f0 & SB2_F_YM)) a0 = dadd(a0, term(d0, wd.x, wc.x));
                }
                if (f1 & SB2_F_DOMAIN) {
                  if (!(f1 & SB2_F_YP)) a1 = dadd(a1, term(d1, wu.y, wc.y));
                  if (!(f1 & SB2_F_YM)) a1 = dadd(a1, term(d1, wd.y, wc.y));
                }
              }
              // axis 2: z
              if (A.dim == 3 && sp.dkind[2] != SB2_SRC_NONE) {
                double d0, d1;
                if (sp.dkind[2] == SB2_SRC_CONST) { d0 = d1 = A.consts[sp.dsrc[2]]; }
                else { const double2 dv = *reinterpret_cast<const double2*>(A.fields[sp.dsrc[2]] + idx); d0 = dv.x; d1 = dv.y; }
                const double2 uu = *reinterpret_cast<const double2*>(sp.u + idx + A.PL);
                const double2 ud = *reinterpret_cast<const double2*>(sp.u + idx - A.PL);
                double2 ku = make_double2(0.0, 0.0), kd = ku;
                if (MODE != MODE_FIRST) {
                  ku = *reinterpret_cast<const double2*>(sp.kprev + idx + A.PL);
                  kd = *reinterpret_cast<const double2*>(sp.kprev + idx - A.PL);
                }
                auto term = [&](double dc, double wn, double w0) {
                  double t = dmul(dc, dsub(wn, w0));
                  if (A.divide[2]) t = ddiv_pos(t, A.dx2[2]);
                  return t;
                };
                if (f0 & SB2_F_DOMAIN) {
                  if (!(f0 & SB2_F_ZP)) a0 = dadd(a0, term(d0, stage_value(uu.x, ku.x), wc.x));
                  if (!(f0 & SB2_F_ZM)) a0 = dadd(a0, term(d0, stage_value(ud.x, kd.x), wc.x));
                }
                if (f1 & SB2_F_DOMAIN) {
                  if (!(f1 & SB2_F_ZP)) a1 = dadd(a1, term(d1, stage_value(uu.y, ku.y), wc.y));
                  if (!(f1 & SB2_F_ZM)) a1 = dadd(a1, term(d1, stage_value(ud.y, kd.y), wc.y));
                }
              }
            }

            if (active[p]) {
              const bool in0 = f0 & SB2_F_DOMAIN, in1 = f1 & SB2_F_DOMAIN;
              if (MODE != MODE_FINAL) {
                double2 o;
                o.x = in0 ? a0 : 0.0;
                o.y = in1 ? a1 : 0.0;
                *reinterpret_cast<double2*>(sp.kout + idx) = o;
              } else {
                // value += dt*(k0 + 2.0*k1 + 2.0*k2 + k3)/6.0, spatialsimulator.cpp:1535
                const double2 u = cu[s][p], k0 = ck0[s][p], k1 = ck1[s][p], k2 = ck2[s][p];
                double2 o = u;
                if (in0) o.x = dadd(u.x, ddiv_pos(dmul(A.dt, dadd(dadd(dadd(k0.x, dmul(2.0, k1.x)), dmul(2.0, k2.x)), a0)), 6.0));
                if (in1) o.y = dadd(u.y, ddiv_pos(dmul(A.dt, dadd(dadd(dadd(k0.y, dmul(2.0, k1.y)), dmul(2.0, k2.y)), a1)), 6.0));
                *reinterpret_cast<double2*>(sp.u_out + idx) = o;
              }
            }
          }
        }
      }
    } else {
      // nothing of this warp's cells lies in a domain: k_m = 0 / u unchanged
      const int64_t idx0 = plane0 + (int64_t)y * A.P + x0 + 2 * tx;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (s < S) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            if (!active[p]) continue;
            const int64_t idx = idx0 + p * 2 * NT;
            if (MODE != MODE_FINAL) {
              *reinterpret_cast<double2*>(A.sp[s].kout + idx) =
                  A.carry ? *reinterpret_cast<const double2*>(A.sp[s].kcarry + idx) : make_double2(0.0, 0.0);
            } else {
              *reinterpret_cast<double2*>(A.sp[s].u_out + idx) = *reinterpret_cast<const double2*>(A.sp[s].u + idx);
            }
          }
        }
      }
    }

    if (more) {
      park_row(y + 2);
#pragma unroll
      for (int c = 0; c < NC; ++c) fl[c] = fn[c];
    }
    __syncthreads();
  }
}

template <int NS, int NP, int MODE, int DEPTH>
cudaError_t launch_variant(const Sb2StageArgs& a, cudaStream_t stream) {
  constexpr int TX = 2 * NP * NT;
  static bool configured = false;
  const size_t smem = (size_t)NS * RING * (TX + 4) * sizeof(double);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(stage_kernel<NS, NP, MODE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  stage_kernel<NS, NP, MODE, DEPTH><<<grid, NT, smem, stream>>>(a);
  return cudaGetLastError();
}

template <int NS, int NP, int DEPTH>
cudaError_t launch_mode(const Sb2StageArgs& a, cudaStream_t stream) {
  if (a.m == 0) return launch_variant<NS, NP, MODE_FIRST, DEPTH>(a, stream);
  if (a.final) return launch_variant<NS, NP, MODE_FINAL, DEPTH>(a, stream);
  return launch_variant<NS, NP, MODE_MIDDLE, DEPTH>(a, stream);
}

// shallow expression stacks (every shipped model but fish) get the variant with half the stack registers
template <int NS, int NP>
cudaError_t launch_depth(const Sb2StageArgs& a, cudaStream_t stream) {
  if (a.depth <= 4) return launch_mode<NS, NP, 4>(a, stream);
  return launch_mode<NS, NP, SB2_MAX_DEPTH>(a, stream);
}

}  // namespace
#endif  // SB2_STAGE_KERNEL_CUH_
