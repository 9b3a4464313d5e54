// rd_stage.cuh — the fused reaction–diffusion RK-stage kernel (sm_100a), second design; 2-D and 3-D grids.
//
// One launch computes, for every volume cell of the slab and every staged species,
//     k_m = sum over reactions / rate rules (bytecode interpreter)  +  diffusion stencil
// and, on the last stage, the RK4 combine u_new = u + dt*(k0 + 2k1 + 2k2 + k3)/6 — the work the
// reference spreads over reversePolishRK (calcPDE.cpp:224-451), calcDiffusion (calcPDE.cpp:453-562)
// and updateSpecies (spatialsimulator.cpp:1518-1544), one full-grid pass per reaction and per
// species.  Each species field is read once per stage from HBM.
//
// Data movement and schedule
//   * A CTA of W warps owns a strip of SW = 64*NP consecutive x cells and a range of rows; it walks the
//     rows in batches of W.  In batch b warp j
//       (1) waits for the raw row  y+1  (y = Y_b + j) that the TMA unit copied into the warp's private
//           slot of shared memory (cp.async.bulk global→shared, completion on the warp's mbarrier),
//       (2) turns it into the stage input  w = u + rk[m]*dt*k_{m-1}  (calcPDE.cpp:247-248), ONCE per
//           cell, and parks it in a ring of w rows shared by the CTA,
//       (3) re-arms its mbarrier and issues the bulk copy of the row it needs in the next batch, which
//           is in flight during everything that follows,
//       (4) waits, on two row-ready mbarriers, for exactly the two ring rows other warps stage for it (rows y-1 and y);
//           there is no CTA-wide barrier in the loop, so staging, token dispatch and stencil of different warps overlap,
//       (5) evaluates the reaction bytecode and the 5-point stencil of row y from the w ring (rows
//           y-1, y, y+1) and stores k_m (or the new u) straight from registers with 16-byte stores.
//     The ring holds 2W+2 rows: the warp that re-stages a slot has, through the chain of step-(4) waits, seen the three
//     rows around the old content computed (see the comment at the wait).
//   * Each thread owns NP pairs of adjacent cells; pair p of lane l sits at x0 + 64p + 2l, so every
//     16-byte shared or global access of a warp covers 512 contiguous bytes.
//   * A per-(row, strip) class byte, computed once at finalize, says whether every cell of the strip row
//     is an interior cell of every geometry (class 1: straight-line stencil, no per-cell test), lies
//     outside every domain (class 2: nothing to compute) or needs the per-cell flag logic (class 0: every
//     stencil term is computed and added under a predicate, no data-dependent branch).
//   * The token stream (4-byte code words and 8-byte inline constants, two parallel arrays) is copied from a small
//     device buffer to shared memory once per CTA and fetched one token ahead.  The program counter is
//     warp-uniform: the dispatch is one indexed branch through a dense jump table without a range check; each
//     case acts on all 2*NP cells of the thread with compile-time register indices (tokens carry their stack slot).
//   * The last stage fuses the RK4 combine: the k0 / k1 rows of a warp's row arrive through a second per-warp TMA slot
//     requested one batch ahead; u and k2 of its own cells are re-read with plain 16-byte loads that hit L2, because the
//     bulk copies that staged them one batch earlier carry an L2::evict_last hint.
//   * 3-D grids march along z instead, with a ring of three plane tiles (rd_stage3_body, described there).
//   * A model-specialised build (rd_jit.cu, SB2_RD_PROGRAM) replaces the interpreter loop by the straight-line expansion
//     of the model's token records; everything else is the same code.
//
// Arithmetic follows the reference's evaluation order with explicit round-to-nearest intrinsics
// (never contracted into FMAs), so results are bit-identical to the CPU code for + - * / and
// comparisons.
#ifndef SB2_RD_STAGE_CUH_
#define SB2_RD_STAGE_CUH_
#include "sb2_internal.h"
#include "sb2_ops.cuh"

namespace {
namespace rd {

enum { MODE_FIRST = 0, MODE_MIDDLE = 1, MODE_FINAL = 2 };

// ---------------------------------------------------------------------------------------------
// PTX: mbarrier + bulk asynchronous copy (TMA, non-tensor form)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of the (converged) warp, chosen by the hardware: code under it may use the uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Programmatic dependent launch: consecutive stage kernels of a step are launched with programmatic stream serialisation
// (sb2_rd_launch_config), so a kernel's CTAs are scheduled and set up their barriers while the previous kernel drains;
// pdl_wait() returns when the previous kernel has completed and its writes are visible (a no-op without the attribute).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copies below read what the previous kernel wrote with plain stores
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// L2 policy for rows that are read a second time one batch later (the last stage re-reads u and k2 for the combine):
// without it ncu showed the re-reads served from DRAM (6.5 GB instead of 4.4 GB per launch at 8192^2).
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// the same with a suspend-time hint (ns): the warp may sleep in the barrier unit until the phase completes instead of
// coming back to spin (ncu: a tenth of the executed instructions of the 3-D kernel were spins of the loop below)
__device__ __forceinline__ uint32_t mbar_try_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return done;
}
// A wait that does not complete within ~0.2 s is a bug (a copy that never lands, a broken dependency chain).  With
// debug words attached (SB2_DEBUG=1) the caller records where and carries on; otherwise the kernel traps: the launch
// fails loudly at the next synchronisation instead of handing back fields computed from stale rows.
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, bool may_trap) {
  const long long t0 = clock64();
  while (!mbar_try_wait_sleep(bar, parity, 100000u))
    if (clock64() - t0 > 400000000LL) {
      if (may_trap) __trap();
      return false;
    }
  return true;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, bool may_trap) {
  if (mbar_try_wait(bar, parity)) return true;  // the copy was issued a whole batch ago: normally it has landed
  return mbar_wait_slow(bar, parity, may_trap);
}

__host__ __device__ inline int tok_bytes(int nrd) { return ((nrd + 2) * 16 + 127) / 128 * 128; }

// Shared-memory header: W bulk-copy barriers, W combine barriers, RW row-ready barriers (8 bytes each)
constexpr int HDR = 512;
#ifndef SB2_RD_P2P
#define SB2_RD_P2P 1  // 1: warps wait for the two ring rows they need (row-ready mbarriers); 0: one __syncthreads() per batch
#endif

template <bool B>
struct BoolTag { static constexpr bool value = B; };

// What never changes for a model - species and geometry counts, which species diffuse, which take the uniform-coefficient
// stencil, the geometry a species lives in - is read from the kernel parameters by the interpreter build and is a literal
// in a model-specialised build (rd_jit.cu defines the SB2_RD_* macros), where the tests on it fold away.
// Debug words (SB2_DEBUG=1: waits that time out are recorded instead of trapping).  A model-specialised build never carries
// them (the library runs the interpreter kernel when debugging), so its waits need no bookkeeping at all.
#ifdef SB2_RD_PROGRAM
#define RD_DBG(A) ((uint32_t*)nullptr)
#else
#define RD_DBG(A) (A).dbg
#endif

#ifdef SB2_RD_MODEL_BITS
#define MODEL_S(A) SB2_RD_NS
#define MODEL_G(A) SB2_RD_G
#define SP_HAS_DIFF(sp, s) (((SB2_RD_HASDIFF_MASK) >> (s)) & 1)
#define SP_FASTPATH(sp, s) (((SB2_RD_FASTPATH_MASK) >> (s)) & 1)
#define SP_GEO(sp, s) (((SB2_RD_GEO_BITS) >> (2 * (s))) & 3)
#else
#define MODEL_S(A) (A).S
#define MODEL_G(A) (A).G
#define SP_HAS_DIFF(sp, s) (sp).has_diff
#define SP_FASTPATH(sp, s) (sp).fastpath
#define SP_GEO(sp, s) (sp).geo
#endif

template <int NP, int W>
struct Geo {
  static constexpr int NC = 2 * NP;     // cells per thread
  static constexpr int SW = 64 * NP;    // cells per strip row
  static constexpr int SWP = SW + 4;    // [pad][left halo][SW cells][right halo][pad]; cell i at index i + 2
  static constexpr int RW = 2 * W + 2;  // rows of the w ring
};

// Division by 6.0 for the RK combine: q = RN(x * RN(1/6)), r = x - 6q (exact in an FMA), q' = RN(q + r*RN(1/6))
// is the correctly rounded quotient whenever no intermediate leaves the normal range (Markstein 1990; the
// significand of 6.0 is not all ones).  N numerators at once: the three-instruction chains overlap, one range test
// covers them all, and only if some numerator lies outside a generous exponent window (or is a NaN / infinity) the
// true division runs for every one of them.  A zero numerator returns itself (x / 6 == x bit for bit, sign included).
template <int N>
__device__ __forceinline__ void div6_batch(double (&x)[N]) {
  const double y = 0.16666666666666666;  // RN(1/6)
  double q[N];
  bool ok = true;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const uint32_t e = (uint32_t)__double2hiint(x[i]) & 0x7ff00000u;
    const bool zero = x[i] == 0.0;
    ok = ok && (zero || (e - 0x08100000u) < 0x6fe00000u);
    const double q0 = __dmul_rn(x[i], y);
    const double r = __fma_rn(-q0, 6.0, x[i]);
    q[i] = zero ? x[i] : __fma_rn(r, y, q0);
  }
  if (!ok) {
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = ddiv_pos(x[i], 6.0);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = q[i];
}
__device__ __forceinline__ double div6(double x) {
  double a[1] = {x};
  div6_batch<1>(a);
  return a[0];
}

#define RD_REP_D0(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)
#define RD_REP_D1(M) M(1) M(2) M(3) M(4) M(5) M(6) M(7)

// Diffusion stencil (calcPDE.cpp:484-559) and output of one strip row: k_m, or on the last stage the new u.
// acc holds the reaction terms of the thread's cells; cur0 / up0 / dn0 (and, in 3-D, zp0 / zm0) point at this lane's
// first cell in the w rows y, y+1, y-1 (z+1, z-1) of species 0, species s sits sstride doubles further.  segmask: in a
// mixed strip row (class 0), bit p says whether the 64-cell segment of cell pair p needs the per-cell predicates.
template <int NS, int NP, int MODE, bool D3>
__device__ __forceinline__ void rd_emit_row(const Sb2RdArgs& A, const int S, const int cls, const int lane, const bool (&active)[NP],
                                            const uint32_t (&fl)[2 * NP], double (&acc)[NS][2 * NP], const int64_t idx0,
                                            const double* mycs, const double* cur0, const double* up0, const double* dn0,
                                            const double* zp0, const double* zm0, const int sstride, const uint32_t segmask) {
  constexpr int SW = 64 * NP;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (s < S) {
      const Sb2SpeciesArgs& sp = A.sp[s];
      const double* cur = cur0 + s * sstride;
      const double* up = up0 + s * sstride;
      const double* dn = dn0 + s * sstride;
      const double* zp = D3 ? zp0 + s * sstride : nullptr;
      const double* zm = D3 ? zm0 + s * sstride : nullptr;
      // Uniform coefficients, no division by dx^2: straight-line stencil, additions in the reference's order Xp, Xm, Yp,
      // Ym (, Zp, Zm).  MASKED = false: every cell of the strip row is an interior cell, no per-cell test at all.
      // MASKED = true (mixed strip rows): the same arithmetic, every addition under a predicate built from the cell's
      // flag byte, outputs selected by the domain bit (neighbour values outside a domain are finite, never used).
      auto straight = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        const double dcx = sp.dconst[0], dcy = sp.dconst[1];
        double2 cu[NP], ck0[NP], ck1[NP], ck2[NP];
        if (MODE == MODE_FINAL) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const int64_t i = idx0 + p * 64;
            cu[p] = ck2[p] = make_double2(0.0, 0.0);
            if (!MASKED || active[p]) {
              cu[p] = *reinterpret_cast<const double2*>(sp.u + i);        // L2 hits: staged through L2 one batch ago
              ck2[p] = *reinterpret_cast<const double2*>(sp.kprev + i);
            }
#ifdef SB2_RD_FINAL_LDG
            // k0 / k1 are read once and never shared: plain loads (no shared-memory slot, which buys the last stage of a 3-D
            // grid a second CTA per SM at four cells per thread)
            ck0[p] = ck1[p] = make_double2(0.0, 0.0);
            if (!MASKED || active[p]) {
              ck0[p] = *reinterpret_cast<const double2*>(sp.k0 + i);
              ck1[p] = *reinterpret_cast<const double2*>(sp.k1 + i);
            }
#else
            ck0[p] = *reinterpret_cast<const double2*>(mycs + (s * 2) * SW + p * 64 + 2 * lane);
            ck1[p] = *reinterpret_cast<const double2*>(mycs + (s * 2 + 1) * SW + p * 64 + 2 * lane);
#endif
          }
        }
        // one pair of cells; PM: this pair's 64-cell segment needs the per-cell predicates (segmask, warp-uniform)
        auto pair = [&](const int p, auto pair_tag) {
          constexpr bool PM = decltype(pair_tag)::value;
          const uint32_t f0 = PM ? (fl[2 * p] >> (8 * SP_GEO(sp, s))) & 0xffu : (uint32_t)SB2_F_DOMAIN;
          const uint32_t f1 = PM ? (fl[2 * p + 1] >> (8 * SP_GEO(sp, s))) & 0xffu : (uint32_t)SB2_F_DOMAIN;
          // a term is skipped when the cell carries the boundary bit of that side; cells outside the domain accumulate
          // whatever their neighbours hold and are dropped at the output select (in0 / in1), so the domain bit need not
          // be part of every term's predicate
#define RD_ADD(acc_, f_, bit_, term_) if (!PM || !((f_) & (bit_))) acc_ = dadd(acc_, term_)
          double a0 = acc[s][2 * p], a1 = acc[s][2 * p + 1];
          if (SP_HAS_DIFF(sp, s)) {
            const double2 wc = *reinterpret_cast<const double2*>(cur + p * 64);
            const double wl = cur[p * 64 - 1], wr = cur[p * 64 + 2];
            const double2 wu = *reinterpret_cast<const double2*>(up + p * 64);
            const double2 wd = *reinterpret_cast<const double2*>(dn + p * 64);
            RD_ADD(a0, f0, SB2_F_XP, dmul(dcx, dsub(wc.y, wc.x)));
            RD_ADD(a0, f0, SB2_F_XM, dmul(dcx, dsub(wl, wc.x)));
            RD_ADD(a0, f0, SB2_F_YP, dmul(dcy, dsub(wu.x, wc.x)));
            RD_ADD(a0, f0, SB2_F_YM, dmul(dcy, dsub(wd.x, wc.x)));
            RD_ADD(a1, f1, SB2_F_XP, dmul(dcx, dsub(wr, wc.y)));
            RD_ADD(a1, f1, SB2_F_XM, dmul(dcx, dsub(wc.x, wc.y)));
            RD_ADD(a1, f1, SB2_F_YP, dmul(dcy, dsub(wu.y, wc.y)));
            RD_ADD(a1, f1, SB2_F_YM, dmul(dcy, dsub(wd.y, wc.y)));
            if (D3) {  // Zp, Zm
              const double dcz = sp.dconst[2];
              const double2 wzp = *reinterpret_cast<const double2*>(zp + p * 64);
              const double2 wzm = *reinterpret_cast<const double2*>(zm + p * 64);
              RD_ADD(a0, f0, SB2_F_ZP, dmul(dcz, dsub(wzp.x, wc.x)));
              RD_ADD(a0, f0, SB2_F_ZM, dmul(dcz, dsub(wzm.x, wc.x)));
              RD_ADD(a1, f1, SB2_F_ZP, dmul(dcz, dsub(wzp.y, wc.y)));
              RD_ADD(a1, f1, SB2_F_ZM, dmul(dcz, dsub(wzm.y, wc.y)));
            }
          }
#undef RD_ADD
          const bool in0 = !PM || (f0 & SB2_F_DOMAIN), in1 = !PM || (f1 & SB2_F_DOMAIN);
          if (PM && !active[p]) return;
          if (MODE != MODE_FINAL) {
            *reinterpret_cast<double2*>(sp.kout + idx0 + p * 64) = make_double2(in0 ? a0 : 0.0, in1 ? a1 : 0.0);
          } else {
            // value += dt*(k0 + 2.0*k1 + 2.0*k2 + k3)/6.0, spatialsimulator.cpp:1535
            double inc[2];
            inc[0] = dmul(A.dt, dadd(dadd(dadd(ck0[p].x, dmul(2.0, ck1[p].x)), dmul(2.0, ck2[p].x)), a0));
            inc[1] = dmul(A.dt, dadd(dadd(dadd(ck0[p].y, dmul(2.0, ck1[p].y)), dmul(2.0, ck2[p].y)), a1));
            div6_batch<2>(inc);
            *reinterpret_cast<double2*>(sp.u_out + idx0 + p * 64) =
                make_double2(in0 ? dadd(cu[p].x, inc[0]) : cu[p].x, in1 ? dadd(cu[p].y, inc[1]) : cu[p].y);
          }
        };
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          if (MASKED && ((segmask >> p) & 1u)) pair(p, BoolTag<true>());
          else pair(p, BoolTag<false>());
        }
      };
      if (SP_FASTPATH(sp, s) || !SP_HAS_DIFF(sp, s)) {
        if (cls == 1) straight(BoolTag<false>());
        else straight(BoolTag<true>());
        continue;
      }
      // general path: per-cell domain / boundary flags, field-valued coefficients, dx^2 != 1.  Every stencil term
      // is computed unconditionally (neighbour values outside the domain are finite) and added under a predicate,
      // so the only branches are the warp-uniform ones on the coefficient kind and on dx^2 == 1.
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int64_t idx = idx0 + p * 64;
        const uint32_t f0 = (fl[2 * p] >> (8 * SP_GEO(sp, s))) & 0xffu;
        const uint32_t f1 = (fl[2 * p + 1] >> (8 * SP_GEO(sp, s))) & 0xffu;
        double a0 = acc[s][2 * p], a1 = acc[s][2 * p + 1];
        if (SP_HAS_DIFF(sp, s)) {
          const double2 wc = *reinterpret_cast<const double2*>(cur + p * 64);
          // one axis: the plus / minus neighbours of both cells, the flag bits that switch each term off
          auto axis = [&](int ax, double n0p, double n0m, double n1p, double n1m, uint32_t bitp, uint32_t bitm) {
            if (sp.dkind[ax] == SB2_SRC_NONE) return;
            double d0, d1;
            if (sp.dkind[ax] == SB2_SRC_CONST) { d0 = d1 = sp.dconst[ax]; }
            else {
              double2 dv = make_double2(0.0, 0.0);
              if (active[p]) dv = *reinterpret_cast<const double2*>(A.fields[sp.dsrc[ax]] + idx);
              d0 = dv.x; d1 = dv.y;
            }
            double t0p = dmul(d0, dsub(n0p, wc.x)), t0m = dmul(d0, dsub(n0m, wc.x));
            double t1p = dmul(d1, dsub(n1p, wc.y)), t1m = dmul(d1, dsub(n1m, wc.y));
            if (A.divide[ax]) {
              t0p = ddiv_pos(t0p, A.dx2[ax]); t0m = ddiv_pos(t0m, A.dx2[ax]);
              t1p = ddiv_pos(t1p, A.dx2[ax]); t1m = ddiv_pos(t1m, A.dx2[ax]);
            }
            if ((f0 & (SB2_F_DOMAIN | bitp)) == SB2_F_DOMAIN) a0 = dadd(a0, t0p);
            if ((f0 & (SB2_F_DOMAIN | bitm)) == SB2_F_DOMAIN) a0 = dadd(a0, t0m);
            if ((f1 & (SB2_F_DOMAIN | bitp)) == SB2_F_DOMAIN) a1 = dadd(a1, t1p);
            if ((f1 & (SB2_F_DOMAIN | bitm)) == SB2_F_DOMAIN) a1 = dadd(a1, t1m);
          };
          {
            const double wl = cur[p * 64 - 1], wr = cur[p * 64 + 2];
            axis(0, wc.y, wl, wr, wc.x, SB2_F_XP, SB2_F_XM);
          }
          {
            const double2 wu = *reinterpret_cast<const double2*>(up + p * 64);
            const double2 wd = *reinterpret_cast<const double2*>(dn + p * 64);
            axis(1, wu.x, wd.x, wu.y, wd.y, SB2_F_YP, SB2_F_YM);
          }
          if (D3) {
            const double2 wzp = *reinterpret_cast<const double2*>(zp + p * 64);
            const double2 wzm = *reinterpret_cast<const double2*>(zm + p * 64);
            axis(2, wzp.x, wzm.x, wzp.y, wzm.y, SB2_F_ZP, SB2_F_ZM);
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
            const double2 u = *reinterpret_cast<const double2*>(sp.u + idx);
            double2 o = u;
            if (in0 || in1) {
#ifdef SB2_RD_FINAL_LDG
              const double2 k0 = *reinterpret_cast<const double2*>(sp.k0 + idx);
              const double2 k1 = *reinterpret_cast<const double2*>(sp.k1 + idx);
#else
              const double2 k0 = *reinterpret_cast<const double2*>(mycs + (s * 2) * SW + p * 64 + 2 * lane);
              const double2 k1 = *reinterpret_cast<const double2*>(mycs + (s * 2 + 1) * SW + p * 64 + 2 * lane);
#endif
              const double2 k2 = *reinterpret_cast<const double2*>(sp.kprev + idx);
              if (in0) o.x = dadd(u.x, div6(dmul(A.dt, dadd(dadd(dadd(k0.x, dmul(2.0, k1.x)), dmul(2.0, k2.x)), a0))));
              if (in1) o.y = dadd(u.y, div6(dmul(A.dt, dadd(dadd(dadd(k0.y, dmul(2.0, k1.y)), dmul(2.0, k2.y)), a1))));
            }
            *reinterpret_cast<double2*>(sp.u_out + idx) = o;
          }
        }
      }
    }
  }
}

// Everything a warp does for one strip row once its w rows are in shared memory: reaction program, stencil, output.
// cur0 / up0 / dn0 (/ zp0 / zm0): this lane's first cell in the w rows y, y+1, y-1 (z+1, z-1) of species 0; species s
// sits SSTRIDE doubles further.  tokbase: the token records in shared memory (interpreter build).  Returns false when
// the row needed no arithmetic (the caller still has to wait for the k0 / k1 rows it requested), true otherwise.
template <int NS, int NP, int MODE, int DEPTH, bool D3, int SSTRIDE>
__device__ __forceinline__ bool rd_compute_row(const Sb2RdArgs& A, const double* __restrict__ kc, const unsigned char* tokbase,
                                               const int S, const int cls, const int lane, const bool (&active)[NP],
                                               const uint32_t (&fl)[2 * NP], const int64_t idx0, const double* cur0,
                                               const double* up0, const double* dn0, const double* zp0, const double* zm0,
                                               const double* mycs, const uint32_t mycbar, const uint32_t cparity,
                                               const uint32_t segmask) {
  constexpr int NC = 2 * NP;
  if (cls == 2) {
    // nothing of this strip row lies in a domain: k_m = 0 / u unchanged
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < S) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          if (!active[p]) continue;
          const int64_t idx = idx0 + p * 64;
          if (MODE != MODE_FINAL) *reinterpret_cast<double2*>(A.sp[s].kout + idx) = make_double2(0.0, 0.0);
          else *reinterpret_cast<double2*>(A.sp[s].u_out + idx) = *reinterpret_cast<const double2*>(A.sp[s].u + idx);
        }
      }
    }
    return false;
  }
  if (cls == 0) {
    uint32_t anyf = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) anyf |= fl[c];
    if (!__any_sync(0xffffffffu, (anyf & 0x01010101u) != 0) && A.carry != 1) {
      // this warp's cells are all outside (a class-0 strip row is mixed only as a whole)
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (s < S) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            if (!active[p]) continue;
            const int64_t idx = idx0 + p * 64;
            if (MODE != MODE_FINAL) *reinterpret_cast<double2*>(A.sp[s].kout + idx) = make_double2(0.0, 0.0);
            else *reinterpret_cast<double2*>(A.sp[s].u_out + idx) = *reinterpret_cast<const double2*>(A.sp[s].u + idx);
          }
        }
      }
      return false;
    }
  }

  double acc[NS][NC];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      acc[s][2 * p] = 0.0; acc[s][2 * p + 1] = 0.0;
      if (A.carry && (A.carry == 1 || cls == 0) && s < S && active[p]) {
        const double2 c = *reinterpret_cast<const double2*>(A.sp[s].kcarry + idx0 + p * 64);
        acc[s][2 * p] = c.x; acc[s][2 * p + 1] = c.y;
      }
    }
  }

  // ---------------- reaction / rate-rule bytecode ----------------
  {
    double r[DEPTH][NC];  // RD(d): slot d of the expression stack; slots >= DEPTH do not exist in this variant
#define RD(d) r[(d) < DEPTH ? (d) : 0]
#define LIVE(d) if ((d) < DEPTH)
    uint32_t mask_bit = 0;  // set by MASK (bit of the law's geometry in the flag words), read by the ACC*M forms only
    const double* wop = cur0;  // + s*SSTRIDE selects the species, + 64p the pair

    // What each instruction does to the 2*NP cells of the thread: B_<op>(d, arg, cv) with d the stack slot (compile
    // time), arg the operand field and cv the inline constant.  The interpreter passes the fields of the fetched
    // record, a model-specialised build passes literals.
#define LOADVS(v, sidx, p) const double2 v = *reinterpret_cast<const double2*>(wop + (sidx) * SSTRIDE + (p) * 64)
#define FORC _Pragma("unroll") for (int c_ = 0; c_ < NC; ++c_)
#define FORP _Pragma("unroll") for (int p = 0; p < NP; ++p)
#define B_PUSHC(d, arg, cv) { const double b_ = (cv); FORC RD(d)[c_] = b_; }
#define B_PUSHV(d, arg, cv) { FORP { LOADVS(v, arg, p); RD(d)[2 * p] = v.x; RD(d)[2 * p + 1] = v.y; } }
#define B_PUSHF(d, arg, cv) { FORP { double2 v = make_double2(0.0, 0.0); if (active[p]) v = *reinterpret_cast<const double2*>(A.fields[arg] + idx0 + p * 64); RD(d)[2 * p] = v.x; RD(d)[2 * p + 1] = v.y; } }
    // r[d-1] = r[d-1] op r[d]
#define B_BIN(d, EXPR) { FORC { const double a = RD(d - 1)[c_], b = RD(d)[c_]; RD(d - 1)[c_] = (EXPR); } }
#define B_ADD(d, arg, cv) B_BIN(d, dadd(a, b))
#define B_SUB(d, arg, cv) B_BIN(d, dsub(a, b))
#define B_MUL(d, arg, cv) B_BIN(d, dmul(a, b))
#define B_DIV(d, arg, cv) { ddiv_batch<NC>(RD(d - 1), RD(d), RD(d - 1)); }
#define B_GEQ(d, arg, cv) B_BIN(d, (a >= b) ? 1.0 : 0.0)
#define B_GT(d, arg, cv) B_BIN(d, (a > b) ? 1.0 : 0.0)
#define B_LEQ(d, arg, cv) B_BIN(d, (a <= b) ? 1.0 : 0.0)
#define B_LT(d, arg, cv) B_BIN(d, (a < b) ? 1.0 : 0.0)
#define B_AND(d, arg, cv) B_BIN(d, logic_and(a, b))
#define B_OR(d, arg, cv) B_BIN(d, logic_or(a, b))
    // r[d] = r[d] op c   (R forms: c op r[d])
#define B_BINC(d, cv, EXPR) { const double b = (cv); FORC { const double a = RD(d)[c_]; RD(d)[c_] = (EXPR); } }
#define B_ADDC(d, arg, cv) B_BINC(d, cv, dadd(a, b))
#define B_SUBC(d, arg, cv) B_BINC(d, cv, dsub(a, b))
#define B_MULC(d, arg, cv) B_BINC(d, cv, dmul(a, b))
#define B_RSUBC(d, arg, cv) B_BINC(d, cv, dsub(b, a))
#define B_DIVC(d, arg, cv) { double b[NC]; FORC b[c_] = (cv); ddiv_batch<NC>(RD(d), b, RD(d)); }
#define B_RDIVC(d, arg, cv) { double a[NC]; FORC a[c_] = (cv); ddiv_batch<NC>(a, RD(d), RD(d)); }
#define B_GEQC(d, arg, cv) B_BINC(d, cv, (a >= b) ? 1.0 : 0.0)
#define B_GTC(d, arg, cv) B_BINC(d, cv, (a > b) ? 1.0 : 0.0)
#define B_LEQC(d, arg, cv) B_BINC(d, cv, (a <= b) ? 1.0 : 0.0)
#define B_LTC(d, arg, cv) B_BINC(d, cv, (a < b) ? 1.0 : 0.0)
    // r[d] = r[d] op species arg
#define B_BINV(d, arg, EXPR) { FORP { LOADVS(v, arg, p); { const double a = RD(d)[2 * p], b = v.x; RD(d)[2 * p] = (EXPR); } { const double a = RD(d)[2 * p + 1], b = v.y; RD(d)[2 * p + 1] = (EXPR); } } }
#define B_ADDV(d, arg, cv) B_BINV(d, arg, dadd(a, b))
#define B_SUBV(d, arg, cv) B_BINV(d, arg, dsub(a, b))
#define B_MULV(d, arg, cv) B_BINV(d, arg, dmul(a, b))
#define B_DIVV(d, arg, cv) { double b[NC]; FORP { LOADVS(v, arg, p); b[2 * p] = v.x; b[2 * p + 1] = v.y; } ddiv_batch<NC>(RD(d), b, RD(d)); }
    // fused forms of the lowering pass: r[d] = species op constant, r[d] = species op species
#define B_BINVC(d, arg, cv, EXPR) { const double b = (cv); FORP { LOADVS(v, arg, p); { const double a = v.x; RD(d)[2 * p] = (EXPR); } { const double a = v.y; RD(d)[2 * p + 1] = (EXPR); } } }
#define B_ADDVC(d, arg, cv) B_BINVC(d, arg, cv, dadd(a, b))
#define B_SUBVC(d, arg, cv) B_BINVC(d, arg, cv, dsub(a, b))
#define B_RSUBVC(d, arg, cv) B_BINVC(d, arg, cv, dsub(b, a))
#define B_MULVC(d, arg, cv) B_BINVC(d, arg, cv, dmul(a, b))
#define B_DIVVC(d, arg, cv) { double a[NC], b[NC]; FORP { LOADVS(v, arg, p); a[2 * p] = v.x; a[2 * p + 1] = v.y; b[2 * p] = (cv); b[2 * p + 1] = (cv); } ddiv_batch<NC>(a, b, RD(d)); }
#define B_RDIVVC(d, arg, cv) { double a[NC], b[NC]; FORP { LOADVS(v, arg, p); b[2 * p] = v.x; b[2 * p + 1] = v.y; a[2 * p] = (cv); a[2 * p + 1] = (cv); } ddiv_batch<NC>(a, b, RD(d)); }
#define B_BINVV(d, arg, EXPR) { const uint32_t s1_ = (arg) & 0xffu, s2_ = (arg) >> 8; FORP { LOADVS(v1, s1_, p); LOADVS(v2, s2_, p); { const double a = v1.x, b = v2.x; RD(d)[2 * p] = (EXPR); } { const double a = v1.y, b = v2.y; RD(d)[2 * p + 1] = (EXPR); } } }
#define B_ADDVV(d, arg, cv) B_BINVV(d, arg, dadd(a, b))
#define B_SUBVV(d, arg, cv) B_BINVV(d, arg, dsub(a, b))
#define B_MULVV(d, arg, cv) B_BINVV(d, arg, dmul(a, b))
#define B_DIVVV(d, arg, cv) { const uint32_t s1_ = (arg) & 0xffu, s2_ = (arg) >> 8; double a[NC], b[NC]; FORP { LOADVS(v1, s1_, p); LOADVS(v2, s2_, p); a[2 * p] = v1.x; a[2 * p + 1] = v1.y; b[2 * p] = v2.x; b[2 * p + 1] = v2.y; } ddiv_batch<NC>(a, b, RD(d)); }
#define B_UNARY(d, arg, cv) { FORC RD(d)[c_] = sb2_unary((int)(arg), RD(d)[c_]); }
#define B_BINR(d, arg, cv) { FORC RD(d - 1)[c_] = sb2_binary((int)(arg), RD(d - 1)[c_], RD(d)[c_]); }
#define B_MOV0(d, arg, cv) { FORC RD(0)[c_] = RD(d)[c_]; }
#define B_MOVUP(d, arg, cv) { FORC RD(d + 1)[c_] = RD(d)[c_]; }
    // statement mask: the law only acts on cells of its own domain (spatialsimulator.cpp:1381-1390)
#define B_MASK(d, arg, cv) { mask_bit = 1u << (8 * (arg)); }
    // delta -= stoich*rate (reactants) / += (products, rate rules); calcPDE.cpp:432-449.  Here d is the species.
#define SI(s) ((s) < NS ? (s) : 0)
#define B_ACCSUB(s, arg, cv) { const double st = (cv); FORC acc[SI(s)][c_] = dsub(acc[SI(s)][c_], dmul(st, RD(0)[c_])); }
#define B_ACCADD(s, arg, cv) { const double st = (cv); FORC acc[SI(s)][c_] = dadd(acc[SI(s)][c_], dmul(st, RD(0)[c_])); }
#define B_ACCSUB1(s, arg, cv) { FORC acc[SI(s)][c_] = dsub(acc[SI(s)][c_], RD(0)[c_]); }
#define B_ACCADD1(s, arg, cv) { FORC acc[SI(s)][c_] = dadd(acc[SI(s)][c_], RD(0)[c_]); }
#define B_ACCSUBC1(s, arg, cv) { const double st = (cv); FORC acc[SI(s)][c_] = dsub(acc[SI(s)][c_], st); }
#define B_ACCADDC1(s, arg, cv) { const double st = (cv); FORC acc[SI(s)][c_] = dadd(acc[SI(s)][c_], st); }
#define B_ACCSUBM(s, arg, cv) { const double st = (cv); FORC if (fl[c_] & mask_bit) acc[SI(s)][c_] = dsub(acc[SI(s)][c_], dmul(st, RD(0)[c_])); }
#define B_ACCADDM(s, arg, cv) { const double st = (cv); FORC if (fl[c_] & mask_bit) acc[SI(s)][c_] = dadd(acc[SI(s)][c_], dmul(st, RD(0)[c_])); }
    // reactant and product of a unit-stoichiometry law in one record: acc[s] -= r0; acc[arg] += r0
#define B_XFER(s, arg, cv) { FORC acc[SI(s)][c_] = dsub(acc[SI(s)][c_], RD(0)[c_]); \
    _Pragma("unroll") for (int s2_ = 0; s2_ < NS; ++s2_) if ((arg) == (uint32_t)s2_) { FORC acc[s2_][c_] = dadd(acc[s2_][c_], RD(0)[c_]); } }

#ifdef SB2_RD_PROGRAM
#define KC(i) kc[i]
    SB2_RD_PROGRAM
#undef KC
#else
    // token stream, two parallel arrays: code | arg << 16 (4 bytes each) and the inline constants (8 bytes each)
    const double* tokc = reinterpret_cast<const double*>(tokbase);
    const uint32_t* toks = reinterpret_cast<const uint32_t*>(tokbase + (A.nrd + 2) * 8);
    uint32_t tnext = toks[0];
    int pc = 0;
#pragma unroll 1
    for (;;) {
      const uint32_t t = tnext;
      const double* cvp = tokc + pc;  // inline constant of the C forms, loaded by the cases that use it
      tnext = toks[++pc];  // one token ahead; the stream ends with two OPC_END tokens
      const uint32_t arg = t >> 16;
      switch (t & 0xffffu) {
        // slot d must exist in this variant (DL: the deepest slot the instruction touches)
#define CASE(OP, d, DL) case SB2_CODE(OPC_##OP, d): LIVE(DL) B_##OP(d, arg, *cvp) break;
#define CASE_D(OP) CASE(OP, 0, 0) CASE(OP, 1, 1) CASE(OP, 2, 2) CASE(OP, 3, 3) CASE(OP, 4, 4) CASE(OP, 5, 5) CASE(OP, 6, 6) CASE(OP, 7, 7)
#define CASE_D1(OP) CASE(OP, 1, 1) CASE(OP, 2, 2) CASE(OP, 3, 3) CASE(OP, 4, 4) CASE(OP, 5, 5) CASE(OP, 6, 6) CASE(OP, 7, 7)
#define CASE_UP(OP) CASE(OP, 0, 1) CASE(OP, 1, 2) CASE(OP, 2, 3) CASE(OP, 3, 4) CASE(OP, 4, 5) CASE(OP, 5, 6) CASE(OP, 6, 7) CASE(OP, 7, 8)
#define CASE_S(OP) CASE(OP, 0, 0) CASE(OP, 1, (1 < NS ? 0 : DEPTH)) CASE(OP, 2, (2 < NS ? 0 : DEPTH)) CASE(OP, 3, (3 < NS ? 0 : DEPTH)) \
CASE(OP, 4, (4 < NS ? 0 : DEPTH)) CASE(OP, 5, (5 < NS ? 0 : DEPTH)) CASE(OP, 6, (6 < NS ? 0 : DEPTH)) CASE(OP, 7, (7 < NS ? 0 : DEPTH))
        CASE_D(PUSHC) CASE_D(PUSHV) CASE_D(PUSHF)
        CASE_D1(ADD) CASE_D1(SUB) CASE_D1(MUL) CASE_D1(DIV)
        CASE_D1(GEQ) CASE_D1(GT) CASE_D1(LEQ) CASE_D1(LT) CASE_D1(AND) CASE_D1(OR)
        CASE_D(ADDC) CASE_D(SUBC) CASE_D(MULC) CASE_D(DIVC) CASE_D(RSUBC) CASE_D(RDIVC)
        CASE_D(GEQC) CASE_D(GTC) CASE_D(LEQC) CASE_D(LTC)
        CASE_D(ADDV) CASE_D(SUBV) CASE_D(MULV) CASE_D(DIVV)
        CASE_D(ADDVC) CASE_D(SUBVC) CASE_D(RSUBVC) CASE_D(MULVC) CASE_D(DIVVC) CASE_D(RDIVVC)
        CASE_D(ADDVV) CASE_D(SUBVV) CASE_D(MULVV) CASE_D(DIVVV)
        CASE_D(UNARY) CASE_D1(BINR) CASE_D1(MOV0) CASE_UP(MOVUP)
        CASE(MASK, 0, 0)
        CASE_S(ACCSUB) CASE_S(ACCADD) CASE_S(ACCSUB1) CASE_S(ACCADD1) CASE_S(ACCSUBC1) CASE_S(ACCADDC1)
        CASE_S(ACCSUBM) CASE_S(ACCADDM) CASE_S(XFER)
        case SB2_CODE(OPC_END, 0): goto program_done;
        // the library only emits codes this variant implements (sb2_finalize checks depth and species count),
        // so the dispatch needs no range check
        default: __builtin_unreachable();
      }
    }
  program_done:;
#endif
#undef RD
#undef LIVE
  }

  // ---------------- diffusion (calcPDE.cpp:484-559) + output ----------------
#ifdef SB2_RD_FINAL_LDG
  constexpr bool kSlots = !D3;  // (the 2-D kernel keeps its k0 / k1 slots)
#else
  constexpr bool kSlots = true;
#endif
  if (MODE == MODE_FINAL && kSlots) mbar_wait(mycbar, cparity, RD_DBG(A) == nullptr);  // k0 / k1 of this row (requested one batch ago)
  rd_emit_row<NS, NP, MODE, D3>(A, S, cls, lane, active, fl, acc, idx0, mycs, cur0, up0, dn0, zp0, zm0, SSTRIDE, segmask);
  return true;
}

// SB2_RD_PROGRAM (model-specialised build, rd_jit.cpp): the bytecode interpreter is replaced by the straight-line
// expansion of one model's token records; KC(i) is the inline constant of record i, read from the kernel parameters.
template <int NS, int NP, int W, int MODE, int DEPTH>
__device__ __forceinline__ void rd_stage_body(const Sb2RdArgs& A, const double* __restrict__ kc) {
  using G = Geo<NP, W>;
  constexpr int NC = G::NC, SW = G::SW, SWP = G::SWP, RW = G::RW;
  constexpr int NARR = (MODE == MODE_FIRST) ? 1 : 2;  // arrays per species in a raw slot: u (, k_{m-1})
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [2W mbarriers, 128 B] [token records] [w ring: S x RW x SWP] [raw slots: W x S x NARR x SWP]
  // [last stage: k0 / k1 slots: W x S x 2 x SW]
  const int S = MODEL_S(A);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
#ifdef SB2_RD_PROGRAM
  double* wring = reinterpret_cast<double*>(smem_raw + HDR);  // no token records in a model-specialised build
#else
  double* wring = reinterpret_cast<double*>(smem_raw + HDR + tok_bytes(A.nrd));
#endif
  double* rawbase = wring + (size_t)S * RW * SWP;
  double* cbase = rawbase + (size_t)W * S * NARR * SWP;

  const int lane = threadIdx.x & 31;
  // broadcast from lane 0: the compiler then knows the warp index (and every row / slot address derived from it)
  // is warp-uniform and issues the bulk copies from uniform registers instead of a per-lane waterfall loop
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int x0 = blockIdx.x * SW;
  // the slab axis is y, a tile is a range of the slab's rows (3-D grids: rd_stage3_body below)
  const int yb = A.row_begin + (int)blockIdx.y * A.rows_per_block;
  const int ye = min(yb + A.rows_per_block, A.row_end);
  pdl_launch_dependents();
  if (yb >= ye) return;
  const int64_t plane0 = A.first;

  double* myraw = rawbase + (size_t)warp * S * NARR * SWP;
  const uint32_t mybar = smem_addr(&bars[warp]);
  const uint32_t myraw_s = smem_addr(myraw);
  const uint32_t row_bytes = SWP * sizeof(double);

  double* mycs = cbase + (size_t)warp * S * 2 * SW;  // last stage: k0 / k1 of the row this warp computes next
  const uint32_t mycbar = smem_addr(&bars[W + warp]);
  const uint32_t mycs_s = smem_addr(mycs);
  const uint32_t rowbar0 = smem_addr(&bars[2 * W]);  // row-ready barrier of ring slot q at rowbar0 + 8q
  if (lane == 0) {
    mbar_init(mybar, 1);
    if (MODE == MODE_FINAL) mbar_init(mycbar, 1);
    if (SB2_RD_P2P && warp == 0)
      for (int q = 0; q < RW; ++q) mbar_init(rowbar0 + 8 * q, 1);
    mbar_fence_init();
  }
  pdl_wait();
#ifndef SB2_RD_PROGRAM
  // the token records move from their device buffer to shared memory once (first read after the first barrier)
  for (int i = threadIdx.x; i < A.nrd + 2; i += W * 32) {
    const Sb2RdTok t = A.rdtok[i];
    reinterpret_cast<double*>(smem_raw + HDR)[i] = t.c;
    reinterpret_cast<uint32_t*>(smem_raw + HDR + (A.nrd + 2) * 8)[i] = t.code | (t.arg << 16);
  }
  __syncwarp();
#endif
  if (SB2_RD_P2P) __syncthreads();  // once: row barriers and token stream are in place

  // The last stage reads u and k2 of its own cells a second time for the combine, one batch after the bulk copy
  // staged them: those copies (and the k0 / k1 prefetches) ask L2 to keep the lines.
  const uint64_t keep = (MODE == MODE_FINAL) ? l2_policy_keep() : 0;
  // lane 0: bulk copy of raw row r (u and k_{m-1} of every species, halo columns included) into the warp's slot
  auto issue_row = [&](int r) {
    mbar_expect_tx(mybar, row_bytes * NARR * S);
    const int64_t g = plane0 + (int64_t)r * A.P + x0 - 2;
#pragma unroll
    for (int s = 0; s < NS; ++s) {  // compile-time s: the array pointers come from fixed parameter offsets
      if (s < S) {
        if (MODE == MODE_FINAL) {
          bulk_g2s_hint(myraw_s + (s * NARR) * row_bytes, A.sp[s].u + g, row_bytes, mybar, keep);
          bulk_g2s_hint(myraw_s + (s * NARR + 1) * row_bytes, A.sp[s].kprev + g, row_bytes, mybar, keep);
        } else {
          bulk_g2s(myraw_s + (s * NARR) * row_bytes, A.sp[s].u + g, row_bytes, mybar);
          if (MODE != MODE_FIRST) bulk_g2s(myraw_s + (s * NARR + 1) * row_bytes, A.sp[s].kprev + g, row_bytes, mybar);
        }
      }
    }
  };
  // last stage, lane 0: bulk copy of the k0 / k1 rows (own cells) the combine of row r needs, one batch ahead
  auto issue_combine = [&](int r) {
    mbar_expect_tx(mycbar, SW * sizeof(double) * 2 * S);
    const int64_t g = plane0 + (int64_t)r * A.P + x0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < S) {
        bulk_g2s(mycs_s + (s * 2) * SW * sizeof(double), A.sp[s].k0 + g, SW * sizeof(double), mycbar);
        bulk_g2s(mycs_s + (s * 2 + 1) * SW * sizeof(double), A.sp[s].k1 + g, SW * sizeof(double), mycbar);
      }
    }
  };
  uint32_t cparity = 0;

  bool active[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) active[p] = x0 + p * 64 + 2 * lane < A.nx;

  // batch b = -1 is the prologue: warps W-2 and W-1 stage rows yb-1 and yb, nobody computes
  int wrow = yb - W + warp + 1;  // row this warp turns into w in the current batch
  if (wrow >= yb - 1 && wrow <= ye && elect_one()) issue_row(wrow);
  uint32_t parity = 0;
  int slot_w = warp + 2 - W;  // (wrow - (yb - 1)) mod RW, carried from batch to batch
  if (slot_w < 0) slot_w += RW;

  for (int Y = yb - W; Y < ye; Y += W, wrow += W, slot_w = slot_w + W >= RW ? slot_w + W - RW : slot_w + W) {
    const int srow = Y + warp;  // row this warp computes in the current batch
    const bool compute = (Y >= yb) && (srow < ye);
    int cls = 2;
    if (compute) cls = A.carry == 1 ? 0 : (int)A.cls[(uint32_t)srow * (uint32_t)A.nstrips + blockIdx.x];

    // (1) + (2): raw row → w ring
    if (wrow >= yb - 1 && wrow <= ye) {
      if (!mbar_wait(mybar, parity, RD_DBG(A) == nullptr) && RD_DBG(A) && lane == 0) {
        if (atomicAdd(&RD_DBG(A)[0], 1u) == 0) { RD_DBG(A)[1] = blockIdx.x; RD_DBG(A)[2] = blockIdx.y; RD_DBG(A)[3] = warp; RD_DBG(A)[4] = (uint32_t)wrow; RD_DBG(A)[5] = parity; RD_DBG(A)[6] = (uint32_t)yb; RD_DBG(A)[7] = (uint32_t)ye; }
      }
      parity ^= 1;
      const int wslot = slot_w;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (s < S) {
          const double* ru = myraw + (s * NARR) * SWP;
          const double* rk = ru + SWP;
          double* wr = wring + ((size_t)s * RW + wslot) * SWP;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const int i = 2 + p * 64 + 2 * lane;
            double2 w = *reinterpret_cast<const double2*>(ru + i);
            if (MODE != MODE_FIRST) {
              const double2 k = *reinterpret_cast<const double2*>(rk + i);
              w.x = dadd(w.x, dmul(A.cdt, k.x));  // u + ((rk[m]*dt) * k), calcPDE.cpp:247-248
              w.y = dadd(w.y, dmul(A.cdt, k.y));
            }
            *reinterpret_cast<double2*>(wr + i) = w;
          }
          if (lane < 2) {  // halo columns x0-1 and x0+SW
            const int i = lane == 0 ? 1 : SW + 2;
            double w = ru[i];
            if (MODE != MODE_FIRST) w = dadd(w, dmul(A.cdt, rk[i]));
            wr[i] = w;
          }
        }
      }
      __syncwarp();
      // the row is in the ring: the warps that compute the rows around it may go ahead
      if (SB2_RD_P2P && lane == 0) mbar_arrive(rowbar0 + 8 * wslot);
    }
    // (3) the slot is free again: fetch the row of the next batch (in the prologue this is the first fetch of
    // the warps that had nothing to stage yet)
    if (elect_one()) {
      const int nrow = wrow + W;
      if (nrow >= yb - 1 && nrow <= ye) issue_row(nrow);
      // the first k0 / k1 rows; later ones are requested as soon as the combine has read the previous ones
      if (MODE == MODE_FINAL && Y < yb && srow + W < ye) issue_combine(srow + W);
    }
    // row srow's k0 / k1 have been requested one batch ago: whoever leaves this iteration without the combine
    // still has to wait for them and to request the next row
    auto combine_done = [&](bool waited) {
      if (MODE != MODE_FINAL) return;
      if (!waited) { mbar_wait(mycbar, cparity, RD_DBG(A) == nullptr); }
      cparity ^= 1;
      __syncwarp();
      if (srow + W < ye && elect_one()) issue_combine(srow + W);
    };

    // per-cell flag bytes of all geometries (byte g = geometry g); only class-0 rows look at them
    uint32_t fl[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) fl[c] = 0x01010101u;
    if (cls == 0) {
      const int64_t base = plane0 + (int64_t)srow * A.P + x0 + 2 * lane;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        fl[2 * p] = 0; fl[2 * p + 1] = 0;
        if (active[p]) {
#pragma unroll
          for (int g = 0; g < SB2_MAX_GEOS; ++g) {
            if (g < MODEL_G(A)) {
              const uchar2 v = *reinterpret_cast<const uchar2*>(A.flags[g] + base + p * 64);
              fl[2 * p] |= (uint32_t)v.x << (8 * g);
              fl[2 * p + 1] |= (uint32_t)v.y << (8 * g);
            }
          }
        }
      }
    }

    if (!SB2_RD_P2P) __syncthreads();  // (4)
    if (!compute) continue;
    if (SB2_RD_P2P) {
      // (4) rows srow - 1 and srow have been staged by other warps (srow + 1 by this one): wait for exactly those.  The
      // k-th use of a ring slot completes phase k of its barrier.  Rows that turn out to need no arithmetic wait too:
      // a warp may only re-stage a ring slot after the rows around the old content have been computed, and that
      // ordering is carried from warp to warp by these waits.
      const uint32_t q = (uint32_t)(srow - (yb - 1));  // >= 1
      const uint32_t qc = q % RW, qd = (q - 1) % RW;
      bool ok = mbar_wait(rowbar0 + 8 * qd, ((q - 1) / RW) & 1u, RD_DBG(A) == nullptr);
      ok = mbar_wait(rowbar0 + 8 * qc, (q / RW) & 1u, RD_DBG(A) == nullptr) && ok;
      if (!ok && RD_DBG(A) && lane == 0) {
        if (atomicAdd(&RD_DBG(A)[0], 1u) == 0) { RD_DBG(A)[1] = blockIdx.x; RD_DBG(A)[2] = blockIdx.y; RD_DBG(A)[3] = warp; RD_DBG(A)[4] = (uint32_t)srow; RD_DBG(A)[5] = 99; RD_DBG(A)[6] = (uint32_t)yb; RD_DBG(A)[7] = (uint32_t)ye; }
      }
    }

    const int64_t idx0 = plane0 + (int64_t)srow * A.P + x0 + 2 * lane;  // + 64p for pair p
    // ring slots of rows srow, srow + 1, srow - 1 (row srow + 1 is the one this warp just staged)
    const int slot_u = slot_w, slot_c = slot_w == 0 ? RW - 1 : slot_w - 1, slot_d = slot_c == 0 ? RW - 1 : slot_c - 1;
    const double* base = wring + 2 + 2 * lane;
    const bool waited = rd_compute_row<NS, NP, MODE, DEPTH, false, RW * SWP>(A, kc, smem_raw + HDR, S, cls, lane, active, fl, idx0, base + slot_c * SWP,
                                                                                 base + slot_u * SWP, base + slot_d * SWP, nullptr, nullptr,
                                                                                 mycs, mycbar, cparity, 0xffffffffu);
    combine_done(waited);
  }
}

// ---------------------------------------------------------------------------------------------
// 3-D grids: the same stage, marching along z.
//
// A CTA of W warps owns a strip of SW = 64*NP cells in x, W consecutive rows in y and a range of planes in z; warp j
// owns row y0 + j of every plane.  The w values live in a ring of NSLOT = 3 plane tiles of W + 2 rows (the tile rows plus
// the row below and the row above), so all six neighbours of a cell are read from shared memory and every u / k_{m-1}
// value is fetched from global memory once per CTA (the reference and a row-marching kernel read the z neighbours from
// two planes away in memory).  In iteration z warp j
//   (1) waits on the row-computed barriers of rows j-1 and j+1 of plane z-2, whose ring slot it is about to overwrite,
//   (2) waits for the bulk copy of its raw row of plane z+1 (issued one iteration earlier), turns it into w and parks
//       it in ring slot (z+1) mod 3; warps 0 and W-1 do the same for the halo rows y0-1 and y0+W, which only they read,
//   (3) requests the raw rows of plane z+2,
//   (4) waits on the row-ready barriers of rows j-1 and j+1 of plane z (its own rows of planes z-1, z, z+1 it staged
//       itself),
//   (5) evaluates reactions and the 7-point stencil of row j of plane z, stores k_m (or the new u) and signals the
//       row-computed barrier of (z, j).
// Every wait refers to work of an earlier iteration of a neighbouring warp, so the warps of a CTA run about one plane
// apart and the waits rarely block.  Row (p, j) of the ring is read by warps j-1, j, j+1 while they compute plane p and
// by warp j itself for planes p-1 and p+1; (1) plus program order cover all of them.
// The class map of a 3-D grid has one byte per 64 cells of a row whatever NP is (a strip combines NP of them), so that
// the stages of a step may use different strip widths: two CTAs of 4 cells per thread per SM for the stages that fit,
// three CTAs of 2 cells per thread for the last one, which also holds the k0 / k1 rows.
// ---------------------------------------------------------------------------------------------
template <int NP, int W>
struct Geo3 {
  static constexpr int NC = 2 * NP;
  static constexpr int SW = 64 * NP;
  static constexpr int SWP = SW + 4;
  static constexpr int NR = W + 2;   // ring rows per plane tile: y0-1, y0 .. y0+W-1, y0+W
  static constexpr int NSLOT = 3;    // plane tiles in the ring
};

template <int NS, int NP, int W, int MODE, int DEPTH>
__device__ __forceinline__ void rd_stage3_body(const Sb2RdArgs& A, const double* __restrict__ kc) {
  using G = Geo3<NP, W>;
  constexpr int NC = G::NC, SW = G::SW, SWP = G::SWP, NR = G::NR, NSLOT = G::NSLOT;
  constexpr int NARR = (MODE == MODE_FIRST) ? 1 : 2;  // arrays per species in a raw slot: u (, k_{m-1})
  constexpr int SSTRIDE = NSLOT * NR * SWP;           // doubles between the w tiles of consecutive species
  constexpr int HDRZ = ((2 + 2 * NSLOT) * W * 8 + 511) / 512 * 512;  // barriers (sb2_rd_launch_shape computes the same)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [W bulk-copy barriers, W combine barriers, NSLOT*W row-ready, NSLOT*W row-computed barriers] [token records]
  // [w ring: S x NSLOT x NR x SWP]
  // [raw slots: (W + 2) x S x NARR x SWP] [last stage: k0 / k1 slots: W x S x 2 x SW]
  const int S = MODEL_S(A);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
#ifdef SB2_RD_PROGRAM
  double* wring = reinterpret_cast<double*>(smem_raw + HDRZ);
#else
  double* wring = reinterpret_cast<double*>(smem_raw + HDRZ + tok_bytes(A.nrd));
#endif
  double* rawbase = wring + (size_t)S * SSTRIDE;
  double* cbase = rawbase + (size_t)(W + 2) * S * NARR * SWP;

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler (uniform datapath)
  const int x0 = blockIdx.x * SW;
  const int y0 = (int)blockIdx.y * W;
  const int y = y0 + warp;
  const int zb = A.row_begin + (int)blockIdx.z * A.rows_per_block;
  const int ze = min(zb + A.rows_per_block, A.row_end);
  pdl_launch_dependents();
  if (zb >= ze) return;

  const uint32_t mybar = smem_addr(&bars[warp]);
  const uint32_t mycbar = smem_addr(&bars[W + warp]);
  const uint32_t rowbar0 = smem_addr(&bars[2 * W]);  // row-ready barrier of (slot q, warp j) at rowbar0 + 8 * (q * W + j)
  const uint32_t donebar0 = smem_addr(&bars[(2 + NSLOT) * W]);  // row-computed barrier, same indexing
  if (lane == 0) {
    mbar_init(mybar, 1);
    if (MODE == MODE_FINAL) mbar_init(mycbar, 1);
    if (warp == 0)
      for (int q = 0; q < 2 * NSLOT * W; ++q) mbar_init(rowbar0 + 8 * q, 1);
    mbar_fence_init();
  }
  pdl_wait();
#ifndef SB2_RD_PROGRAM
  for (int i = threadIdx.x; i < A.nrd + 2; i += W * 32) {
    const Sb2RdTok t = A.rdtok[i];
    reinterpret_cast<double*>(smem_raw + HDRZ)[i] = t.c;
    reinterpret_cast<uint32_t*>(smem_raw + HDRZ + (A.nrd + 2) * 8)[i] = t.code | (t.arg << 16);
  }
#endif
  __syncthreads();  // once: barriers and token records are in place
  if (y > A.ny) return;  // rows past the guard row above the plane: nothing to stage, nothing to compute

  // the halo row this warp stages besides its own: y0-1 (warp 0), y0+W (warp W-1, when that row exists)
  const bool halo_lo = warp == 0, halo_hi = warp == W - 1 && y0 + W <= A.ny;
  const int nrows = 1 + (halo_lo ? 1 : 0) + (halo_hi ? 1 : 0);
  double* myraw = rawbase + (size_t)warp * S * NARR * SWP;
  double* hraw = rawbase + (size_t)(halo_lo ? W : W + 1) * S * NARR * SWP;  // W >= 2: a warp has at most one halo row
  const uint32_t myraw_s = smem_addr(myraw), hraw_s = smem_addr(hraw);
  const uint32_t row_bytes = SWP * sizeof(double);
  double* mycs = cbase + (size_t)warp * S * 2 * SW;
  const uint32_t mycs_s = smem_addr(mycs);
  const bool computes = y < A.ny;
  const bool wait_lo = warp > 0, wait_hi = warp < W - 1 && y + 1 <= A.ny;  // neighbours that stage rows this warp reads
  const bool done_hi = warp < W - 1 && y + 1 < A.ny;  // the neighbour above computes (the one below always does)

  const uint64_t keep = (MODE == MODE_FINAL) ? l2_policy_keep() : 0;
  // one lane: bulk copies of the raw rows (u and k_{m-1} of every species, halo columns included) of plane pz
  // (element offsets of this warp's rows in plane 0; the issue adds pz * PL: the address arithmetic of the bulk copies was a
  // ninth of all instructions issued at 512^3 when every term was recomputed per copy)
  // Running element offsets instead of products per use: grow / ghrow = this warp's own / halo row in the plane the next bulk copies
  // fetch (planes are requested in order, one per iteration); the address arithmetic of the copies - 64-bit products of plane
  // and row pitch for every array - was a ninth of all instructions issued at 512^3.
  int64_t grow = A.first + (int64_t)(zb - 1) * A.PL + (int64_t)y * A.P + x0 - 2;
  int64_t ghrow = A.first + (int64_t)(zb - 1) * A.PL + (int64_t)(halo_lo ? y0 - 1 : y0 + W) * A.P + x0 - 2;
  auto issue_rows = [&]() {
    mbar_expect_tx(mybar, row_bytes * NARR * S * nrows);
    const int64_t g = grow;
    const int64_t gh = ghrow;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < S) {
        if (MODE == MODE_FINAL) {
          bulk_g2s_hint(myraw_s + (s * NARR) * row_bytes, A.sp[s].u + g, row_bytes, mybar, keep);
          bulk_g2s_hint(myraw_s + (s * NARR + 1) * row_bytes, A.sp[s].kprev + g, row_bytes, mybar, keep);
        } else {
          bulk_g2s(myraw_s + (s * NARR) * row_bytes, A.sp[s].u + g, row_bytes, mybar);
          if (MODE != MODE_FIRST) bulk_g2s(myraw_s + (s * NARR + 1) * row_bytes, A.sp[s].kprev + g, row_bytes, mybar);
        }
        if (nrows > 1) {
          bulk_g2s(hraw_s + (s * NARR) * row_bytes, A.sp[s].u + gh, row_bytes, mybar);
          if (MODE != MODE_FIRST) bulk_g2s(hraw_s + (s * NARR + 1) * row_bytes, A.sp[s].kprev + gh, row_bytes, mybar);
        }
      }
    }
  };
  auto issue_combine = [&](int64_t g) {  // g: element offset of the strip row's first cell in the plane wanted
    mbar_expect_tx(mycbar, SW * sizeof(double) * 2 * S);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < S) {
        bulk_g2s(mycs_s + (s * 2) * SW * sizeof(double), A.sp[s].k0 + g, SW * sizeof(double), mycbar);
        bulk_g2s(mycs_s + (s * 2 + 1) * SW * sizeof(double), A.sp[s].k1 + g, SW * sizeof(double), mycbar);
      }
    }
  };
  // raw row → w row of the ring:  w = u + (rk[m]*dt) * k_{m-1}, calcPDE.cpp:247-248
  auto stage = [&](const double* raw, double* wrow) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < S) {
        const double* ru = raw + (s * NARR) * SWP;
        const double* rk = ru + SWP;
        double* wr = wrow + (size_t)s * SSTRIDE;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int i = 2 + p * 64 + 2 * lane;
          double2 w = *reinterpret_cast<const double2*>(ru + i);
          if (MODE != MODE_FIRST) {
            const double2 k = *reinterpret_cast<const double2*>(rk + i);
            w.x = dadd(w.x, dmul(A.cdt, k.x));
            w.y = dadd(w.y, dmul(A.cdt, k.y));
          }
          *reinterpret_cast<double2*>(wr + i) = w;
        }
        if (lane < 2) {  // halo columns x0-1 and x0+SW
          const int i = lane == 0 ? 1 : SW + 2;
          double w = ru[i];
          if (MODE != MODE_FIRST) w = dadd(w, dmul(A.cdt, rk[i]));
          wr[i] = w;
        }
      }
    }
  };

  bool active[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) active[p] = x0 + p * 64 + 2 * lane < A.nx;

#ifdef SB2_RD_FINAL_LDG
  constexpr bool kSlots = false;
#else
  constexpr bool kSlots = true;
#endif
  if (elect_one()) {
    issue_rows();
    if (MODE == MODE_FINAL && kSlots && computes) issue_combine(A.first + (int64_t)zb * A.PL + (int64_t)y * A.P + x0);
  }
  grow += A.PL; ghrow += A.PL;
  uint32_t parity = 0, cparity = 0;
  // zc: the plane computed in this iteration; plane zc + 1 is staged first.  Planes zb-1 .. ze are staged (ring slot and
  // use count of plane pz follow from pz - (zb - 1)), planes zb .. ze-1 are computed.
  // Class bytes and per-cell flag bytes come from global memory; both are requested a whole iteration before they are decoded,
  // so that their latency is covered by the staging and the arithmetic of the plane in between (decoding them right after the
  // load cost a fifth of all issue-stall samples at 512^3).
  //   cls_cur / seg_cur: class of this warp's strip row in the plane computed in this iteration, and which 64-cell segments need
  //   the per-cell predicates;  cb_next: the raw class bytes of the next plane;  fraw: raw flag bytes of the plane computed in
  //   this iteration (requested in the previous one, when its class was decoded)
  int cls_cur = 2;
  uint32_t seg_cur = 0xffffffffu;  // bit p: the 64-cell segment of cell pair p is not all-interior
  uint32_t cb_next[NP];  // as loaded: nothing may consume them before decode_class(), an iteration later
  bool cb_valid[NP];     // the 64-cell segment of pair p exists (the last strip of a row may be narrower)
  uint32_t fraw[NP][SB2_MAX_GEOS];  // the two flag bytes of a cell pair as loaded (16 bits), split only where they are used
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    cb_next[p] = 2u;
    cb_valid[p] = (int)blockIdx.x * NP + p < A.nstrips;
#pragma unroll
    for (int g = 0; g < SB2_MAX_GEOS; ++g) fraw[p][g] = 0u;
  }
  auto load_class = [&](int plane) {  // one byte per 64 cells: 1 all interior, 2 all outside, 0 mixed
    const uint8_t* cb = A.cls + ((uint32_t)plane * (uint32_t)A.ny + (uint32_t)y) * (uint32_t)A.nstrips + blockIdx.x * NP;
#pragma unroll
    for (int p = 0; p < NP; ++p) cb_next[p] = cb[cb_valid[p] ? p : 0];  // (the select is on the address: the byte itself stays untouched)
  };
  auto decode_class = [&]() {
    if (A.carry == 1) { cls_cur = 0; seg_cur = 0xffffffffu; return; }
    int c = cb_valid[0] ? (int)cb_next[0] : -1;
    uint32_t m = c == 1 ? 0u : 1u;
#pragma unroll
    for (int p = 1; p < NP; ++p) {
      const int cp = cb_valid[p] ? (int)cb_next[p] : -1;
      if (cp >= 0) {
        if (cp != c) c = 0;
        if (cp != 1) m |= 1u << p;
      } else {
        // past the grid: the pair is inactive and must take the predicated path, so an interior row becomes mixed
        m |= 1u << p;
        if (c == 1) c = 0;
      }
    }
    cls_cur = c;
    seg_cur = m;
  };
  auto load_flags = [&](int64_t i0, uint32_t segs) {  // raw flag bytes of the segments that need them; i0: the lane's first cell

#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (!((segs >> p) & 1u) || !active[p]) continue;
#pragma unroll
      for (int g = 0; g < SB2_MAX_GEOS; ++g)
        if (g < MODEL_G(A)) fraw[p][g] = (uint32_t)*reinterpret_cast<const unsigned short*>(A.flags[g] + i0 + p * 64);
    }
  };
  if (computes && A.carry != 1) load_class(zb);
  int64_t idx_run = A.first + (int64_t)(zb - 2) * A.PL + (int64_t)y * A.P + x0 + 2 * lane;  // this lane's first cell in plane zc
  // ring slot and barrier phase of the plane staged in this iteration (qs, phs: iteration count modulo / over NSLOT) and of the
  // plane computed in it (qc, phc: one iteration older; qm: two), kept as running values instead of divisions by NSLOT
  int qs = 0, qc = 0, qm = 0;
  uint32_t phs = 0, phc = 0;
  for (int zc = zb - 2; zc < ze; ++zc, idx_run += A.PL, qm = qc, qc = qs, phc = phs, qs = (qs + 1 == NSLOT ? 0 : qs + 1), phs ^= (qs == 0 ? 1u : 0u)) {
    const int pz = zc + 1;
    const bool compute = computes && zc >= zb;
    const int cls = compute ? cls_cur : 2;
#ifdef SB2_RD_ROWMASK  // experiments: per-cell predicates on every segment of a mixed strip row
    const uint32_t segmask = 0xffffffffu;
#else
    const uint32_t segmask = seg_cur;
#endif
    const int64_t idx0 = idx_run;
    // flag words of the plane computed now, from the bytes requested one iteration ago
    uint32_t fl[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) fl[c] = 0x01010101u;
    if (cls == 0) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (!((segmask >> p) & 1u)) continue;  // an all-interior segment of a mixed strip row keeps the interior flags
        fl[2 * p] = 0; fl[2 * p + 1] = 0;
        if (active[p]) {
#pragma unroll
          for (int g = 0; g < SB2_MAX_GEOS; ++g) {
            if (g < MODEL_G(A)) {
              fl[2 * p] |= (fraw[p][g] & 0xffu) << (8 * g);
              fl[2 * p + 1] |= (fraw[p][g] >> 8) << (8 * g);
            }
          }
        }
      }
    }
    // the class of plane pz (computed in the next iteration) was requested an iteration ago: decode it, request the flag bytes
    // that plane needs and the class bytes of the plane after it
    if (computes && pz >= zb && pz < ze) {
      decode_class();
      if (cls_cur == 0) {
#ifdef SB2_RD_ROWMASK
        load_flags(idx_run + A.PL, 0xffffffffu);
#else
        load_flags(idx_run + A.PL, seg_cur);
#endif
      }
      if (A.carry != 1 && pz + 1 < ze) load_class(pz + 1);
    }
    {
      const int q = qs;
      // (1) the slot still holds plane pz - NSLOT: the neighbours must have computed it
      if (pz - (zb - 1) >= NSLOT) {
        const uint32_t ph = phs ^ 1u;
        bool ok = true;
        if (wait_lo) ok = mbar_wait(donebar0 + 8 * (q * W + warp - 1), ph, RD_DBG(A) == nullptr);
        if (done_hi) ok = mbar_wait(donebar0 + 8 * (q * W + warp + 1), ph, RD_DBG(A) == nullptr) && ok;
        if (!ok && RD_DBG(A) && lane == 0) {
          if (atomicAdd(&RD_DBG(A)[0], 1u) == 0) { RD_DBG(A)[1] = blockIdx.x; RD_DBG(A)[2] = blockIdx.y; RD_DBG(A)[3] = warp; RD_DBG(A)[4] = (uint32_t)pz; RD_DBG(A)[5] = 98; RD_DBG(A)[6] = (uint32_t)zb; RD_DBG(A)[7] = (uint32_t)ze; }
        }
      }
      if (!mbar_wait(mybar, parity, RD_DBG(A) == nullptr) && RD_DBG(A) && lane == 0) {
        if (atomicAdd(&RD_DBG(A)[0], 1u) == 0) { RD_DBG(A)[1] = blockIdx.x; RD_DBG(A)[2] = blockIdx.y; RD_DBG(A)[3] = warp; RD_DBG(A)[4] = (uint32_t)pz; RD_DBG(A)[5] = parity; RD_DBG(A)[6] = (uint32_t)zb; RD_DBG(A)[7] = (uint32_t)ze; }
      }
      parity ^= 1;
      double* tile = wring + (size_t)q * NR * SWP;
      stage(myraw, tile + (warp + 1) * SWP);
      if (halo_lo) stage(hraw, tile);
      if (halo_hi) stage(hraw, tile + (W + 1) * SWP);
      __syncwarp();
      if (lane == 0) mbar_arrive(rowbar0 + 8 * (q * W + warp));
      // the raw slots are free again: fetch the rows of the next plane
      if (pz + 1 <= ze && elect_one()) issue_rows();
      grow += A.PL; ghrow += A.PL;
    }
    if (zc < zb - 1) continue;
    // (3) rows j-1 and j+1 of plane zc; every warp that stages waits, whether it computes or not: the order in which
    // ring slots may be overwritten is carried from warp to warp by these waits
    {
      const uint32_t q = (uint32_t)qc, ph = phc;
      bool ok = true;
      if (wait_lo) ok = mbar_wait(rowbar0 + 8 * (q * W + warp - 1), ph, RD_DBG(A) == nullptr);
      if (wait_hi) ok = mbar_wait(rowbar0 + 8 * (q * W + warp + 1), ph, RD_DBG(A) == nullptr) && ok;
      if (!ok && RD_DBG(A) && lane == 0) {
        if (atomicAdd(&RD_DBG(A)[0], 1u) == 0) { RD_DBG(A)[1] = blockIdx.x; RD_DBG(A)[2] = blockIdx.y; RD_DBG(A)[3] = warp; RD_DBG(A)[4] = (uint32_t)zc; RD_DBG(A)[5] = 99; RD_DBG(A)[6] = (uint32_t)zb; RD_DBG(A)[7] = (uint32_t)ze; }
      }
    }
    const int qp = qs;  // plane zc + 1 was staged in this iteration
    if (!compute) {
      // plane zb-1 is never computed: its row-computed phase completes here, so that the use count of a slot is the
      // same for both kinds of barrier
      if (computes && lane == 0) mbar_arrive(donebar0 + 8 * (qc * W + warp));
      continue;
    }
    const double* base = wring + (warp + 1) * SWP + 2 + 2 * lane;  // this lane's first cell in row j of tile 0
    const double* cur = base + (size_t)qc * NR * SWP;
    const bool waited = rd_compute_row<NS, NP, MODE, DEPTH, true, SSTRIDE>(
        A, kc, smem_raw + HDRZ, S, cls, lane, active, fl, idx0, cur, cur + SWP, cur - SWP, base + (size_t)qp * NR * SWP,
        base + (size_t)qm * NR * SWP, mycs, mycbar, cparity, segmask);
    if (MODE == MODE_FINAL && kSlots) {
      if (!waited) mbar_wait(mycbar, cparity, RD_DBG(A) == nullptr);
      cparity ^= 1;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(donebar0 + 8 * (qc * W + warp));  // (5) row (zc, j) will not be read by this warp's stencil again
    if (MODE == MODE_FINAL && kSlots && zc + 1 < ze && elect_one()) issue_combine(idx_run + A.PL - 2 * lane);
  }
}

// 2-D grids march along y, 3-D grids along z
template <int NS, int NP, int W, int MODE, int DEPTH, bool D3>
__device__ __forceinline__ void rd_stage_any(const Sb2RdArgs& A, const double* __restrict__ kc) {
  if constexpr (D3) rd_stage3_body<NS, NP, W, MODE, DEPTH>(A, kc);
  else rd_stage_body<NS, NP, W, MODE, DEPTH>(A, kc);
}

#ifdef SB2_RD_PROGRAM
// model-specialised build: entry points with C names (expanded at global scope by the generated translation unit),
// parameters = the stage arguments followed by the constants
struct Sb2RdJitParams { Sb2RdArgs a; double kc[SB2_RD_NKC]; };
#define RD_JIT_ENTRY(name, MODE, NP_, MINB_) \
  extern "C" __global__ void __launch_bounds__(SB2_RD_W * 32, MINB_) name(const __grid_constant__ rd::Sb2RdJitParams P) { \
    rd::rd_stage_any<SB2_RD_NS, NP_, SB2_RD_W, MODE, SB2_RD_DEPTH, SB2_RD_D3>(P.a, P.kc); }
#else
template <int NS, int NP, int W, int MODE, int DEPTH, int MINB, bool D3>
__global__ void __launch_bounds__(W * 32, MINB) rd_stage_kernel(const __grid_constant__ Sb2RdArgs A) {
  rd_stage_any<NS, NP, W, MODE, DEPTH, D3>(A, nullptr);
}

template <int NS, int NP, int W, int MODE, int DEPTH, int MINB, bool D3>
cudaError_t launch_variant(const Sb2RdArgs& a, cudaStream_t stream) {
  dim3 grid;
  size_t smem;
  sb2_rd_launch_shape(a, NP, W, MODE, D3, tok_bytes(a.nrd), &grid, &smem);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(rd_stage_kernel<NS, NP, W, MODE, DEPTH, MINB, D3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // one shared-memory carve-out for all stage kernels: no SM reconfiguration between consecutive launches (see rd_jit.cu)
    cudaFuncSetAttribute(rd_stage_kernel<NS, NP, W, MODE, DEPTH, MINB, D3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = smem;
  }
  if (grid.y == 0 || grid.z == 0) return cudaSuccess;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  sb2_rd_launch_config(&cfg, attr, grid, W * 32, smem, stream);
  return cudaLaunchKernelEx(&cfg, rd_stage_kernel<NS, NP, W, MODE, DEPTH, MINB, D3>, a);
}

template <int NS, int NP, int W, int DEPTH, int MINB, bool D3>
cudaError_t launch_dim(const Sb2RdArgs& a, cudaStream_t stream) {
  if (a.m == 0) return launch_variant<NS, NP, W, MODE_FIRST, DEPTH, MINB, D3>(a, stream);
  if (a.final) return launch_variant<NS, NP, W, MODE_FINAL, DEPTH, MINB, D3>(a, stream);
  return launch_variant<NS, NP, W, MODE_MIDDLE, DEPTH, MINB, D3>(a, stream);
}

template <int NS, int NP, int W, int DEPTH, int MINB>
cudaError_t launch_mode(const Sb2RdArgs& a, cudaStream_t stream) {
  if (a.dim == 3) return launch_dim<NS, NP, W, DEPTH, MINB, true>(a, stream);
  return launch_dim<NS, NP, W, DEPTH, MINB, false>(a, stream);
}

#endif  // SB2_RD_PROGRAM

}  // namespace rd
}  // namespace
#endif  // SB2_RD_STAGE_CUH_
