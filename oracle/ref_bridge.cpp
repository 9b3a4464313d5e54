// ORACLE / TEST INFRASTRUCTURE — never linked into or called by the product library.
//
// The binding INTEGRATION.md §A describes, for real: the reference's OWN host code (its unmodified SpatialSBML sources, compiled
// in place by oracle/Makefile) sets a model up — variableInfo / GeometryInfo / reactionInfo, masks, boundary types, reverse-polish
// streams — and this file flattens those structs onto the device library's C ABI (include/spatial_b200.h) and steps on the GPU.
// Beside it the reference's own oneStep() runs on the CPU on a second, identical simulator, and the two are compared after
// every step at the volume cells.  What it proves: the sb2_* boundary is sufficient for a maintainer of the reference to drop
// the device step under the reference's host set-up, and it gives the reference's results.
//
//   ref_bridge <model.xml> <xdim> <ydim> <zdim> <dt> <steps>
//
// Covered: species in volumes, reactions and rate rules, uniform and field-valued diffusion coefficients, Neumann / Dirichlet
// boundary conditions, CIP-CSLR advection with uniform and field-valued velocities, membrane transport between two volumes,
// species ON membranes (point sets from domainIndex / pseudoMemIndex, surface diffusion from the reference's voronoiInfo table,
// reactions on the membrane, binding / unbinding, the pseudo-membrane fill), assignment rules re-evaluated every step.
// Refused (exit 3): 1-D models; assignment rules on membrane species.
#define private public  // (test-only access to varInfoList / geoInfoList / rInfoList, as in ref_driver.cpp)
#include "spatialsimulator.h"
#undef private
#include "searchFunction.h"

#include <sbml/SBMLTypes.h>
#include <sbml/packages/spatial/common/SpatialExtensionTypes.h>

#include <execinfo.h>
#include <malloc.h>
#include <signal.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../include/spatial_b200.h"

#define DEV(call)                                                                                         \
  do {                                                                                                    \
    int rc_ = (call);                                                                                     \
    if (rc_ < 0) { fprintf(stderr, "ref_bridge: %s failed: %s\n", #call, sb2_last_error(dev)); return 2; } \
  } while (0)

// AST type (libSBML's ASTNodeType_t, what opfuncList holds) -> operator code of the device library; the cases of the
// reference's interpreter switch (calcPDE.cpp:259-426)
static int opOfAst(int t) {
  switch (t) {
    case AST_PLUS: return SB2_OP_PLUS;
    case AST_MINUS: return SB2_OP_MINUS;
    case AST_TIMES: return SB2_OP_TIMES;
    case AST_DIVIDE: return SB2_OP_DIVIDE;
    case AST_POWER: case AST_FUNCTION_POWER: return SB2_OP_POWER;
    case AST_FUNCTION_LOG: return SB2_OP_LOG;
    case AST_LOGICAL_AND: return SB2_OP_AND;
    case AST_LOGICAL_OR: return SB2_OP_OR;
    case AST_LOGICAL_XOR: return SB2_OP_XOR;
    case AST_RELATIONAL_EQ: return SB2_OP_EQ;
    case AST_RELATIONAL_GEQ: return SB2_OP_GEQ;
    case AST_RELATIONAL_GT: return SB2_OP_GT;
    case AST_RELATIONAL_LEQ: return SB2_OP_LEQ;
    case AST_RELATIONAL_LT: return SB2_OP_LT;
    case AST_RELATIONAL_NEQ: return SB2_OP_NEQ;
    case AST_FUNCTION_ABS: return SB2_OP_ABS;
    case AST_FUNCTION_ARCCOS: return SB2_OP_ARCCOS;
    case AST_FUNCTION_ARCCOSH: return SB2_OP_ARCCOSH;
    case AST_FUNCTION_ARCCSC: return SB2_OP_ARCCSC;
    case AST_FUNCTION_ARCSIN: return SB2_OP_ARCSIN;
    case AST_FUNCTION_ARCTAN: return SB2_OP_ARCTAN;
    case AST_FUNCTION_COS: return SB2_OP_COS;
    case AST_FUNCTION_COSH: return SB2_OP_COSH;
    case AST_FUNCTION_CSC: return SB2_OP_CSC;
    case AST_FUNCTION_EXP: return SB2_OP_EXP;
    case AST_FUNCTION_LN: return SB2_OP_LN;
    case AST_FUNCTION_ROOT: return SB2_OP_ROOT;
    case AST_FUNCTION_SIN: return SB2_OP_SIN;
    case AST_FUNCTION_SINH: return SB2_OP_SINH;
    case AST_FUNCTION_TAN: return SB2_OP_TAN;
    case AST_FUNCTION_TANH: return SB2_OP_TANH;
    case AST_LOGICAL_NOT: return SB2_OP_NOT;
    default: return SB2_OP_POP_ONLY;  // listed-but-empty cases of the switch: they only pop
  }
}

struct Bridge {
  SpatialSimulator& ref;
  sb2_sim* dev;
  int nx, ny, nz;
  std::map<GeometryInfo*, int> geoId;
  std::map<variableInfo*, int> speciesId, fieldId;
  std::map<const double*, int> constSlot;  // address of a uniform double the streams alias -> slot (astFunction.cpp:76-81)
  std::vector<const double*> constAddr;
  int timeSlot;
  // a membrane that hosts species: its storage points (membrane + pseudo-membrane points, ascending doubled-grid index)
  struct PointSet { int id; std::vector<int64_t> storage; std::map<int64_t, int> at; };
  std::map<GeometryInfo*, PointSet> pointSets;
  std::map<variableInfo*, int> memSpeciesId;
  explicit Bridge(SpatialSimulator& r) : ref(r), dev(NULL), timeSlot(-1) {}

  int storageOf(GeometryInfo* g, int64_t idx) const {
    std::map<GeometryInfo*, PointSet>::const_iterator it = pointSets.find(g);
    if (it == pointSets.end()) return -1;
    std::map<int64_t, int>::const_iterator q = it->second.at.find(idx);
    return q == it->second.at.end() ? -1 : q->second;
  }

  int64_t even(int i, int j, int k) const { return ((int64_t)(2 * k) * ref.Yindex + 2 * j) * ref.Xindex + 2 * i; }
  std::vector<double> compact(const double* full) const {
    std::vector<double> out((size_t)nx * ny * nz);
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) out[((size_t)k * ny + j) * nx + i] = full[even(i, j, k)];
    return out;
  }
  // the value at the odd doubled-grid point on the plus side of every cell along axis a (zero past the last cell)
  std::vector<double> oddPoints(const double* full, int a) const {
    std::vector<double> out((size_t)nx * ny * nz, 0.0);
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
          const int X = 2 * i + (a == 0), Y = 2 * j + (a == 1), Z = 2 * k + (a == 2);
          if (X < ref.Xindex && Y < ref.Yindex && Z < ref.Zindex) out[((size_t)k * ny + j) * nx + i] = full[((int64_t)Z * ref.Yindex + Y) * ref.Xindex + X];
        }
    return out;
  }
  variableInfo* ownerOfValue(const double* p) const {
    for (size_t i = 0; i < ref.varInfoList.size(); ++i)
      if (ref.varInfoList[i]->value == p) return ref.varInfoList[i];
    return NULL;
  }
  int slotOf(const double* addr) {
    std::map<const double*, int>::iterator it = constSlot.find(addr);
    if (it != constSlot.end()) return it->second;
    const int s = (int)constAddr.size();
    constSlot[addr] = s;
    constAddr.push_back(addr);
    return s;
  }
  int fieldOf(variableInfo* v) {
    std::map<variableInfo*, int>::iterator it = fieldId.find(v);
    if (it != fieldId.end()) return it->second;
    const std::vector<double> c = compact(v->value);
    const int id = sb2_add_field(dev, c.data());
    fieldId[v] = id;
    return id;
  }
  // a coefficient / boundary value: uniform double -> constant slot, spatial array -> field
  bool source(variableInfo* v, int& kind, int& src) {
    if (!v) { kind = SB2_SRC_NONE; src = 0; return true; }
    if (v->isUniform) { kind = SB2_SRC_CONST; src = slotOf(v->value); return true; }
    kind = SB2_SRC_FIELD; src = fieldOf(v);
    return src >= 0;
  }

  variableInfo* infoById(const std::string& id) const {
    for (size_t i = 0; i < ref.varInfoList.size(); ++i)
      if (id == ref.varInfoList[i]->id) return ref.varInfoList[i];
    return NULL;
  }
  // the geometry of a species, or of another species of its compartment (spatialsimulator.cpp:1394-1426)
  GeometryInfo* geoOfSpecies(const std::string& sid) const {
    variableInfo* v = infoById(sid);
    if (v && v->geoi) return v->geoi;
    const Species* s = ref.model->getSpecies(sid);
    if (!s) return NULL;
    for (unsigned int i = 0; i < ref.model->getNumSpecies(); ++i) {
      const Species* cur = ref.model->getSpecies(i);
      if (cur->getId() != sid && cur->getCompartment() == s->getCompartment()) {
        variableInfo* cv = infoById(cur->getId());
        if (cv && cv->geoi) return cv->geoi;
      }
    }
    return NULL;
  }
  // compact cell of a doubled-grid point, -1 unless it is a volume cell of `geo`
  int32_t cellIn(GeometryInfo* geo, int X, int Y, int Z) const {
    if (!geo || !geo->isVol || !geo->isDomain) return -1;
    if (X < 0 || Y < 0 || Z < 0 || X >= ref.Xindex || Y >= ref.Yindex || Z >= ref.Zindex) return -1;
    if ((X | Y | Z) & 1) return -1;
    return geo->isDomain[((int64_t)Z * ref.Yindex + Y) * ref.Xindex + X] == 1 ? (int32_t)(((int64_t)(Z / 2) * ny + Y / 2) * nx + X / 2) : -1;
  }
  // membrane transport between two volumes (calcMemTransport, calcPDE.cpp:1036-1482): the membrane's points from the reference's
  // domainIndex / isDomain / bType / nuVec, the operand and target cells from the species' isDomain masks
  int addTransport(reactionInfo* ri) {
    Reaction* r = ri->reaction;
    if (r->getNumReactants() == 0 || r->getNumProducts() == 0) return 0;
    GeometryInfo* rg = geoOfSpecies(r->getReactant(0)->getSpecies());
    GeometryInfo* pg = geoOfSpecies(r->getProduct(0)->getSpecies());
    // every membrane the reference calls calcMemTransport on for this reaction, in call order (spatialsimulator.cpp:1428-1450):
    // the membrane between the two compartments, then — binding / unbinding — the membrane one side lives on
    std::vector<GeometryInfo*> mems;
    for (size_t j = 0; j < ref.geoInfoList.size(); ++j) {
      GeometryInfo* g = ref.geoInfoList[j];
      if (!g->isVol && ((g->adjacentGeo1 == rg && g->adjacentGeo2 == pg) || (g->adjacentGeo1 == pg && g->adjacentGeo2 == rg))) { mems.push_back(g); break; }
    }
    const bool haveRV = rg && rg->isVol, havePV = pg && pg->isVol;
    if (haveRV != havePV) {
      if (!haveRV && rg) mems.push_back(rg);
      if (!havePV && pg) mems.push_back(pg);
    }
    for (size_t q = 0; q < mems.size(); ++q) {
      const int rc = addTransportOn(ri, mems[q]);
      if (rc) return rc;
    }
    return 0;  // (no membrane between the two compartments: nothing happens)
  }
  int addTransportOn(reactionInfo* ri, GeometryInfo* mem) {
    Reaction* r = ri->reaction;
    const int Xi = ref.Xindex, Yi = ref.Yindex, Zi = ref.Zindex;
    const int64_t N = (int64_t)Xi * Yi * Zi;
    struct Pt { int X, Y, Z; int64_t index; };
    std::vector<Pt> pts;
    for (size_t k = 0; k < mem->domainIndex.size(); ++k) {
      const int64_t idx = mem->domainIndex[k];
      if (idx < 0 || idx >= N) continue;
      Pt p;
      p.index = idx; p.Z = (int)(idx / ((int64_t)Xi * Yi)); p.Y = (int)((idx - (int64_t)p.Z * Xi * Yi) / Xi); p.X = (int)(idx - (int64_t)p.Z * Xi * Yi - (int64_t)p.Y * Xi);
      if (ref.dimension == 3 && (p.X == 0 || p.Y == 0 || p.Z == 0 || p.X == Xi - 1 || p.Y == Yi - 1 || p.Z == Zi - 1)) continue;
      if (ref.dimension == 2 && (p.X == 0 || p.Y == 0 || p.X == Xi - 1 || p.Y == Yi - 1)) continue;
      if (mem->isDomain[idx] != 1) continue;
      pts.push_back(p);
    }
    const int npts = (int)pts.size();
    reversePolishInfo* rp = ri->rpInfo;
    std::vector<sb2_token> toks;
    std::vector<int32_t> operandCells;
    std::vector<double> operandValues;
    variableInfo* symbol = NULL;  // persists across tokens like the reference's symbolInfo
    for (int i = 0; i < rp->listNum; ++i) {
      sb2_token d;
      d.kind = SB2_TOK_NOP; d.op = 0; d.slot = 0;
      if (rp->varList && rp->varList[i]) {
        variableInfo* v = ownerOfValue(rp->varList[i]);
        if (v && rp->deltaList && rp->deltaList[i] && memSpeciesId.count(v)) {
          // a symbol that lives on a membrane is read at the point itself (calcPDE.cpp:1214-1217)
          for (size_t j = 0; j < ri->spRefList.size(); ++j)
            if (ri->spRefList[j] == v) { symbol = ri->spRefList[j]; break; }
          d.kind = SB2_TOK_MEMVAR; d.slot = (uint16_t)memSpeciesId[v];
          for (int p = 0; p < npts; ++p) {
            operandCells.push_back(storageOf(v->geoi, pts[p].index));
            operandCells.push_back(-1);
          }
        } else if (v && rp->deltaList && rp->deltaList[i] && speciesId.count(v)) {
          for (size_t j = 0; j < ri->spRefList.size(); ++j)
            if (ri->spRefList[j] == v) { symbol = ri->spRefList[j]; break; }
          if (!symbol || !symbol->geoi) { fprintf(stderr, "ref_bridge: transport operand without geometry\n"); return 2; }
          d.kind = SB2_TOK_VAR; d.slot = (uint16_t)speciesId[v];
          for (int p = 0; p < npts; ++p) {
            int32_t nearC = -1, farC = -1;
            const int dX[3] = {1, 0, 0}, dY[3] = {0, 1, 0}, dZ[3] = {0, 0, 1};
            for (int a = 0; a < (int)ref.dimension; ++a) {  // later axes overwrite earlier ones (calcPDE.cpp:1107-1213)
              int32_t c = cellIn(symbol->geoi, pts[p].X + dX[a], pts[p].Y + dY[a], pts[p].Z + dZ[a]);
              int sgn = 1;
              if (c < 0) { c = cellIn(symbol->geoi, pts[p].X - dX[a], pts[p].Y - dY[a], pts[p].Z - dZ[a]); sgn = -1; }
              if (c < 0) continue;
              nearC = c;
              const int64_t flat3 = pts[p].index + sgn * 3 * (a == 0 ? 1 : a == 1 ? (int64_t)Xi : (int64_t)Xi * Yi);
              farC = -1;
              if (flat3 < N && flat3 >= 0) {
                const int Z3 = (int)(flat3 / ((int64_t)Xi * Yi)), Y3 = (int)((flat3 - (int64_t)Z3 * Xi * Yi) / Xi), X3 = (int)(flat3 - (int64_t)Z3 * Xi * Yi - (int64_t)Y3 * Xi);
                farC = cellIn(symbol->geoi, X3, Y3, Z3);
              }
            }
            operandCells.push_back(nearC);
            operandCells.push_back(farC);
          }
        } else if (v) {  // no delta array: the raw value at the membrane point itself (calcPDE.cpp:1218-1221)
          d.kind = SB2_TOK_FIELD; d.slot = 0;
          for (int p = 0; p < npts; ++p) operandValues.push_back(v->value[pts[p].index]);
        }
      } else if (rp->constList && rp->constList[i]) {
        d.kind = SB2_TOK_CONST; d.slot = (uint16_t)slotOf(rp->constList[i]);
        if (rp->constList[i] == &ref.sim_time) timeSlot = d.slot;
      } else if (rp->opfuncList && rp->opfuncList[i] != 0) {
        d.kind = SB2_TOK_OP; d.op = (uint8_t)opOfAst(rp->opfuncList[i]);
      }
      toks.push_back(d);
    }
    std::vector<sb2_mem_point> dpts((size_t)npts);
    for (int p = 0; p < npts; ++p) {
      const boundaryType& b = mem->bType[pts[p].index];
      const int axis = (b.isBofXp || b.isBofXm) ? 0 : (b.isBofYp || b.isBofYm) ? 1 : 2;
      const normalUnitVector& n = ref.nuVec[pts[p].index];
      dpts[p].axis = axis;
      dpts[p].abs_normal = fabs(axis == 0 ? n.nx : axis == 1 ? n.ny : n.nz);
    }
    std::vector<sb2_ref> refs;
    std::vector<int32_t> targets;
    for (size_t k = 0; k < ri->spRefList.size(); ++k) {
      variableInfo* v = ri->spRefList[k];
      const bool ok = v && !v->isUniform && speciesId.count(v) && v->geoi && v->geoi->isVol;
      const bool onMem = v && !v->isUniform && memSpeciesId.count(v) && v->geoi && !v->geoi->isVol;
      sb2_ref rr;
      rr.species = ok ? speciesId[v] : onMem ? SB2_MEM_SPECIES + memSpeciesId[v] : -1;
      rr.is_variable = ri->isVariable[k] ? 1 : 0;
      rr.stoichiometry = ri->srStoichiometry[k];
      refs.push_back(rr);
      for (int p = 0; p < npts; ++p) {
        const int a = dpts[p].axis, dx = a == 0, dy = a == 1, dz = a == 2;
        if (onMem) {  // binding: the species' own membrane holds this point (isDomain[index] == 1, calcPDE.cpp:1405-1407)
          targets.push_back(v->geoi->isDomain[pts[p].index] == 1 ? storageOf(v->geoi, pts[p].index) : -1);
          targets.push_back(-1);
          continue;
        }
        targets.push_back(ok ? cellIn(v->geoi, pts[p].X + dx, pts[p].Y + dy, pts[p].Z + dz) : -1);
        targets.push_back(ok ? cellIn(v->geoi, pts[p].X - dx, pts[p].Y - dy, pts[p].Z - dz) : -1);
      }
    }
    DEV(sb2_add_mem_transport(dev, toks.data(), (int)toks.size(), refs.data(), (int)refs.size(), (int)r->getNumReactants(), dpts.data(), npts,
                              operandCells.empty() ? NULL : operandCells.data(), operandValues.empty() ? NULL : operandValues.data(),
                              targets.empty() ? NULL : targets.data()));
    return 0;
  }

  // Point set of a membrane that hosts species, with the pseudo-membrane fill pairs (spatialsimulator.cpp:1609-1663)
  int ensurePointSet(GeometryInfo* g) {
    if (pointSets.count(g)) return 0;
    PointSet ps;
    for (size_t k = 0; k < g->domainIndex.size(); ++k) ps.storage.push_back(g->domainIndex[k]);
    for (size_t k = 0; k < g->pseudoMemIndex.size(); ++k) ps.storage.push_back(g->pseudoMemIndex[k]);
    std::sort(ps.storage.begin(), ps.storage.end());
    ps.storage.erase(std::unique(ps.storage.begin(), ps.storage.end()), ps.storage.end());
    for (size_t k = 0; k < ps.storage.size(); ++k) ps.at[ps.storage[k]] = (int)k;
    ps.id = sb2_add_point_set(dev, (int)ps.storage.size());
    if (ps.id < 0) { fprintf(stderr, "ref_bridge: sb2_add_point_set: %s\n", sb2_last_error(dev)); return 2; }
    pointSets[g] = ps;
    const int64_t Xi = ref.Xindex, Yi = ref.Yindex, total = Xi * Yi * ref.Zindex;
    std::vector<int32_t> t, a, b;
    for (size_t q = 0; q < g->pseudoMemIndex.size(); ++q) {
      const int64_t idx = g->pseudoMemIndex[q];
      const int64_t n[6] = {idx + 1, idx - 1, idx + Xi, idx - Xi, idx + Xi * Yi, idx - Xi * Yi};  // Xp Xm Yp Ym Zp Zm
      bool is1[6];
      for (int k = 0; k < 6; ++k) is1[k] = n[k] >= 0 && n[k] < total && g->isDomain[n[k]] == 1;
      if (ref.dimension != 3) is1[4] = is1[5] = false;
      static const int order[15][2] = {{0, 1}, {2, 3}, {4, 5}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {2, 4}, {2, 5}, {3, 4}, {3, 5}};
      for (int c = 0; c < 15; ++c) {
        if (!(is1[order[c][0]] && is1[order[c][1]])) continue;
        const int st = storageOf(g, idx), sa = storageOf(g, n[order[c][0]]), sb = storageOf(g, n[order[c][1]]);
        if (st >= 0 && sa >= 0 && sb >= 0) { t.push_back(st); a.push_back(sa); b.push_back(sb); }
        break;
      }
    }
    DEV(sb2_set_mem_fill(dev, ps.id, t.data(), a.data(), b.data(), (int)t.size()));
    return 0;
  }
  // A species on a membrane: values at the storage points, the points updateSpecies advances, calcMemDiffusion's terms from the
  // reference's voronoiInfo table (calcPDE.cpp:1484-1581)
  int addMemSpecies(variableInfo* v) {
    GeometryInfo* g = v->geoi;
    int rc = ensurePointSet(g);
    if (rc) return rc;
    const PointSet& ps = pointSets[g];
    std::vector<double> init(ps.storage.size() ? ps.storage.size() : 1);
    for (size_t k = 0; k < ps.storage.size(); ++k) init[k] = v->value[ps.storage[k]];
    const int id = sb2_add_mem_species(dev, ps.id, init.data());
    if (id < 0) { fprintf(stderr, "ref_bridge: sb2_add_mem_species: %s\n", sb2_last_error(dev)); return 2; }
    memSpeciesId[v] = id;
    std::vector<int32_t> upd;
    for (size_t q = 0; q < g->domainIndex.size(); ++q) upd.push_back(storageOf(g, g->domainIndex[q]));
    DEV(sb2_set_mem_update_points(dev, id, upd.data(), (int)upd.size()));
    if (!v->diffCInfo || !v->diffCInfo[0] || !ref.vorI) return 0;  // calcMemDiffusion reads diffCInfo[0] only
    variableInfo* dc = v->diffCInfo[0];
    std::vector<int32_t> start(1, 0), adj;
    std::vector<double> sij, dij, area, dvals;
    for (size_t q = 0; q < g->domainIndex.size(); ++q) {
      const int64_t idx = g->domainIndex[q];
      const voronoiInfo& vor = ref.vorI[idx];
      double ar = 0.0;
      for (int j = 0; j < 2; ++j) {
        ar += vor.diXY[j] * vor.siXY[j];
        ar += vor.diYZ[j] * vor.siYZ[j];
        ar += vor.diXZ[j] * vor.siXZ[j];
      }
      if (ref.dimension == 2) ar /= 2.0;
      else if (ref.dimension == 3) ar /= 4.0;
      area.push_back(ar);
      if (!dc->isUniform) dvals.push_back(dc->value[idx]);
      const boundaryType& bt = g->bType[idx];
      const bool bx = bt.isBofXp && bt.isBofXm, by = bt.isBofYp && bt.isBofYm, bz = bt.isBofZp && bt.isBofZm;
      const bool plane[3] = {bx || by, ref.dimension == 3 && (by || bz), ref.dimension == 3 && (bx || bz)};
      const int* adjOf[3] = {vor.adjacentIndexXY, vor.adjacentIndexYZ, vor.adjacentIndexXZ};
      const double* sOf[3] = {vor.siXY, vor.siYZ, vor.siXZ};
      const double* dOf[3] = {vor.diXY, vor.diYZ, vor.diXZ};
      for (int pl = 0; pl < 3; ++pl) {
        if (!plane[pl]) continue;
        for (int j = 0; j < 2; ++j) {
          const int a = adjOf[pl][j] >= 0 ? storageOf(g, adjOf[pl][j]) : -1;
          if (a < 0) continue;  // no neighbour found on this side (the reference would read outside its array)
          adj.push_back(a); sij.push_back(sOf[pl][j]); dij.push_back(dOf[pl][j]);
        }
      }
      start.push_back((int32_t)adj.size());
    }
    const int kind = dc->isUniform ? SB2_SRC_CONST : SB2_SRC_FIELD;
    const int src = dc->isUniform ? slotOf(dc->value) : 0;
    DEV(sb2_set_mem_diffusion(dev, id, kind, src, dvals.empty() ? NULL : dvals.data(), upd.data(), (int)upd.size(), start.data(), adj.data(),
                              sij.data(), dij.data(), area.data()));
    return 0;
  }
  // A reaction (kind 0) or rate rule (kind 1) evaluated over a membrane's domainIndex (reversePolishRK, calcPDE.cpp:224-451)
  int addMemReaction(reactionInfo* ri, GeometryInfo* g, int kind) {
    if (!pointSets.count(g)) return 0;  // no variable species lives on this membrane: nothing can change
    reversePolishInfo* rp = ri->rpInfo;
    std::vector<int32_t> pts;
    for (size_t q = 0; q < g->domainIndex.size(); ++q) pts.push_back(storageOf(g, g->domainIndex[q]));
    std::vector<sb2_token> toks;
    std::vector<double> values;
    for (int i = 0; i < rp->listNum; ++i) {
      sb2_token d;
      d.kind = SB2_TOK_NOP; d.op = 0; d.slot = 0;
      if (rp->varList && rp->varList[i]) {
        variableInfo* v = ownerOfValue(rp->varList[i]);
        if (v && rp->deltaList && rp->deltaList[i] && memSpeciesId.count(v) && v->geoi == g) { d.kind = SB2_TOK_MEMVAR; d.slot = (uint16_t)memSpeciesId[v]; }
        else if (v) {  // anything else keeps, at a membrane point, the value it was set up with
          d.kind = SB2_TOK_FIELD;
          for (size_t q = 0; q < g->domainIndex.size(); ++q) values.push_back(v->value[g->domainIndex[q]]);
        }
      } else if (rp->constList && rp->constList[i]) {
        d.kind = SB2_TOK_CONST; d.slot = (uint16_t)slotOf(rp->constList[i]);
        if (rp->constList[i] == &ref.sim_time) timeSlot = d.slot;
      } else if (rp->opfuncList && rp->opfuncList[i] != 0) {
        d.kind = SB2_TOK_OP; d.op = (uint8_t)opOfAst(rp->opfuncList[i]);
      }
      toks.push_back(d);
    }
    std::vector<sb2_ref> refs;
    for (size_t k = 0; k < ri->spRefList.size(); ++k) {
      variableInfo* v = ri->spRefList[k];
      sb2_ref rr;
      rr.species = (v && !v->isUniform && memSpeciesId.count(v) && v->geoi == g) ? SB2_MEM_SPECIES + memSpeciesId[v] : -1;
      rr.is_variable = ri->isVariable[k] ? 1 : 0;
      rr.stoichiometry = ri->srStoichiometry[k];
      refs.push_back(rr);
    }
    DEV(sb2_add_mem_reaction(dev, pointSets[g].id, kind, toks.data(), (int)toks.size(), refs.data(), (int)refs.size(),
                             kind == 0 ? (int)ri->reaction->getNumReactants() : 1, pts.data(), (int)pts.size(), values.empty() ? NULL : values.data()));
    return 0;
  }
  // Assignment rules, re-evaluated after every step in the reference's order (updateAssignmentRules, spatialsimulator.cpp:1566-1581):
  // a rule on a species writes the cells of its compartment's volume, a rule on a parameter every cell
  int addAssignmentRules() {
    for (size_t i = 0; i < ref.orderedARule.size(); ++i) {
      variableInfo* info = ref.orderedARule[i];
      reversePolishInfo* rp = info->rpInfo;
      if (!rp) continue;
      int geo = -1;
      if (info->sp) {
        GeometryInfo* g = searchAvolInfoByCompartment(ref.geoInfoList, info->sp->getCompartment().c_str());
        if (!g || !g->isVol) { fprintf(stderr, "ref_bridge: assignment rule on a membrane species is outside the bridge's scope\n"); return 3; }
        if (!geoId.count(g)) {
          const int rc = addGeometry(g);
          if (rc) return rc;
        }
        geo = geoId[g];
      }
      std::vector<sb2_token> toks((size_t)rp->listNum);
      for (int k = 0; k < rp->listNum; ++k) {
        sb2_token t;
        t.kind = SB2_TOK_NOP; t.op = 0; t.slot = 0;
        if (rp->varList && rp->varList[k]) {
          variableInfo* v = ownerOfValue(rp->varList[k]);
          if (v && memSpeciesId.count(v)) { fprintf(stderr, "ref_bridge: membrane species in an assignment rule\n"); return 3; }
          if (v && speciesId.count(v)) { t.kind = SB2_TOK_VAR; t.slot = (uint16_t)speciesId[v]; }
          else if (v) { const int f = fieldOf(v); if (f < 0) return 2; t.kind = SB2_TOK_FIELD; t.slot = (uint16_t)f; }
        } else if (rp->constList && rp->constList[k]) {
          t.kind = SB2_TOK_CONST; t.slot = (uint16_t)slotOf(rp->constList[k]);
          if (rp->constList[k] == &ref.sim_time) timeSlot = t.slot;
        } else if (rp->opfuncList && rp->opfuncList[k] != 0) {
          t.kind = SB2_TOK_OP; t.op = (uint8_t)opOfAst(rp->opfuncList[k]);
        }
        toks[k] = t;
      }
      int kind, target;
      if (speciesId.count(info)) { kind = 1; target = speciesId[info]; }
      else { kind = 0; target = fieldOf(info); if (target < 0) return 2; ruleFields.push_back(info); }
      DEV(sb2_add_assignment_rule(dev, kind, target, geo, toks.data(), (int)toks.size()));
    }
    return 0;
  }
  std::vector<variableInfo*> ruleFields;  // variables without a time derivative that a rule rewrites every step (compared too)

  int addGeometry(GeometryInfo* geo) {
    std::vector<uint8_t> f((size_t)nx * ny * nz);
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
          const int64_t e = even(i, j, k);
          const boundaryType& b = geo->bType[e];
          f[((size_t)k * ny + j) * nx + i] = (uint8_t)((geo->isDomain[e] == 1 ? SB2_F_DOMAIN : 0) | (b.isBofXp ? SB2_F_XP : 0) | (b.isBofXm ? SB2_F_XM : 0) |
                                                      (b.isBofYp ? SB2_F_YP : 0) | (b.isBofYm ? SB2_F_YM : 0) | (b.isBofZp ? SB2_F_ZP : 0) |
                                                      (b.isBofZm ? SB2_F_ZM : 0));
        }
    const int id = sb2_add_geometry(dev, f.data());
    if (id < 0) { fprintf(stderr, "ref_bridge: sb2_add_geometry: %s\n", sb2_last_error(dev)); return 2; }
    geoId[geo] = id;
    return 0;
  }

  int build(int device) {
    nx = ref.Xdiv; ny = ref.Ydiv; nz = ref.Zdiv;
    if (ref.dimension < 2) { fprintf(stderr, "ref_bridge: 2-D and 3-D models only\n"); return 3; }
    sb2_grid g;
    memset(&g, 0, sizeof(g));
    g.nx = nx; g.ny = ny; g.nz = nz; g.dim = (int)ref.dimension;
    g.dx = ref.deltaX; g.dy = ref.deltaY; g.dz = ref.deltaZ; g.device = device;
    int rc = sb2_create(&g, &dev);
    if (rc < 0) { fprintf(stderr, "ref_bridge: sb2_create: %s\n", sb2_last_error(NULL)); return 2; }
    // geometries: one flag byte per volume cell from isDomain + bType (mystruct.h:28-61)
    for (size_t gi = 0; gi < ref.geoInfoList.size(); ++gi) {
      GeometryInfo* geo = ref.geoInfoList[gi];
      if (!geo->isVol || !geo->isDomain || !geo->bType) continue;
      bool hosts = false;
      for (size_t i = 0; i < ref.varInfoList.size(); ++i)
        hosts = hosts || (ref.varInfoList[i]->geoi == geo && ref.varInfoList[i]->sp && !ref.varInfoList[i]->isUniform && ref.varInfoList[i]->delta);
      if (!hosts) continue;
      const int rcg = addGeometry(geo);
      if (rcg) return rcg;
    }
    // species with a time derivative: value[] + delta[4N] (setInfoFunction.cpp:121-126)
    for (size_t i = 0; i < ref.varInfoList.size(); ++i) {
      variableInfo* v = ref.varInfoList[i];
      if (!v->sp || v->isUniform || !v->delta) continue;
      if (v->geoi && !v->geoi->isVol) {  // a species on a membrane
        const int rcm = addMemSpecies(v);
        if (rcm) return rcm;
        continue;
      }
      if (!v->geoi || !geoId.count(v->geoi)) { fprintf(stderr, "ref_bridge: species %s has no volume geometry\n", v->id); return 2; }
      const std::vector<double> c = compact(v->value);
      const int id = sb2_add_species(dev, geoId[v->geoi], c.data());
      if (id < 0) { fprintf(stderr, "ref_bridge: sb2_add_species: %s\n", sb2_last_error(dev)); return 2; }
      speciesId[v] = id;
    }
    // reactions, then rate rules: one token per slot of the reverse-polish arrays (mystruct.h:20-26)
    for (size_t r = 0; r < ref.rInfoList.size(); ++r) {
      reactionInfo* ri = ref.rInfoList[r];
      reversePolishInfo* rp = ri->rpInfo;
      if (ri->reaction && ri->isMemTransport) {  // (the flag is only set for reactions)
        const int rc2 = addTransport(ri);
        if (rc2) return rc2;
        continue;
      }
      std::vector<sb2_token> toks((size_t)rp->listNum);
      for (int i = 0; i < rp->listNum; ++i) {
        sb2_token t;
        t.kind = SB2_TOK_NOP; t.op = 0; t.slot = 0;
        if (rp->varList && rp->varList[i]) {
          variableInfo* v = ownerOfValue(rp->varList[i]);
          if (v && speciesId.count(v) && rp->deltaList && rp->deltaList[i]) { t.kind = SB2_TOK_VAR; t.slot = (uint16_t)speciesId[v]; }
          else if (v) { const int f = fieldOf(v); if (f < 0) return 2; t.kind = SB2_TOK_FIELD; t.slot = (uint16_t)f; }
        } else if (rp->constList && rp->constList[i]) {
          t.kind = SB2_TOK_CONST; t.slot = (uint16_t)slotOf(rp->constList[i]);
          if (rp->constList[i] == &ref.sim_time) timeSlot = t.slot;
        } else if (rp->opfuncList && rp->opfuncList[i] != 0) {
          t.kind = SB2_TOK_OP; t.op = (uint8_t)opOfAst(rp->opfuncList[i]);
        }
        toks[i] = t;
      }
      std::vector<sb2_ref> refs;
      GeometryInfo* geo = NULL;
      for (size_t k = 0; k < ri->spRefList.size(); ++k) {
        variableInfo* v = ri->spRefList[k];
        sb2_ref rr;
        rr.species = (v && speciesId.count(v)) ? speciesId[v] : -1;
        rr.is_variable = ri->isVariable[k] ? 1 : 0;
        rr.stoichiometry = ri->srStoichiometry[k];
        refs.push_back(rr);
        if (!geo && v && v->geoi) geo = v->geoi;  // "first non-null geometry", spatialsimulator.cpp:1381-1390
      }
      const bool isRule = ri->reaction == NULL;
      if (geo && !geo->isVol) {  // a reaction between species of one membrane
        const int rcm = addMemReaction(ri, geo, isRule ? 1 : 0);
        if (rcm) return rcm;
        continue;
      }
      if (!geo || !geoId.count(geo)) continue;  // every participant uniform: the reference changes nothing
      const int nReact = isRule ? 0 : (int)ri->reaction->getNumReactants();
      DEV(sb2_add_reaction(dev, geoId[geo], isRule ? 1 : 0, toks.data(), (int)toks.size(), refs.data(), (int)refs.size(), nReact));
    }
    // diffusion coefficients and boundary conditions (setInfoFunction.cpp:157-342, calcPDE.cpp:903-953)
    for (std::map<variableInfo*, int>::iterator it = speciesId.begin(); it != speciesId.end(); ++it) {
      variableInfo* v = it->first;
      for (int a = 0; a < (int)ref.dimension; ++a) {
        if (!v->diffCInfo || !v->diffCInfo[a]) continue;
        int kind, src;
        if (!source(v->diffCInfo[a], kind, src)) return 2;
        DEV(sb2_set_diffusion(dev, it->second, a, kind, src));
      }
      // advection velocities (variableInfo::adCInfo) and the CIP face point values, which the reference keeps at the odd points of
      // the species' own array (calcPDE.cpp:588-603)
      bool advected = false;
      for (int a = 0; a < (int)ref.dimension; ++a) {
        if (!v->adCInfo || !v->adCInfo[a]) continue;
        advected = true;
        int kind, src;
        if (!source(v->adCInfo[a], kind, src)) return 2;
        DEV(sb2_set_advection(dev, it->second, a, kind, src));
        if (kind == SB2_SRC_FIELD) {  // ... and the velocity between the cells (divergence correction, calcPDE.cpp:623-625)
          const std::vector<double> fv = oddPoints(v->adCInfo[a]->value, a);
          const int ff = sb2_add_field(dev, fv.data());
          if (ff < 0) return 2;
          DEV(sb2_set_advection_faces(dev, it->second, a, ff));
        }
      }
      if (advected)
        for (int a = 0; a < (int)ref.dimension; ++a) {
          const std::vector<double> faces = oddPoints(v->value, a);
          DEV(sb2_upload_faces(dev, it->second, a, faces.data()));
        }
      for (int side = 0; side < 2 * (int)ref.dimension; ++side) {
        if (!v->boundaryInfo || !v->boundaryInfo[side]) continue;
        variableInfo* b = v->boundaryInfo[side];
        BoundaryCondition* bc = static_cast<SpatialParameterPlugin*>(b->para->getPlugin("spatial"))->getBoundaryCondition();
        const int bk = bc->getType() == SPATIAL_BOUNDARYKIND_DIRICHLET ? SB2_BC_DIRICHLET : bc->getType() == SPATIAL_BOUNDARYKIND_NEUMANN ? SB2_BC_NEUMANN : SB2_BC_NONE;
        int kind, src;
        if (!source(b, kind, src)) return 2;
        DEV(sb2_set_boundary(dev, it->second, side, bk, kind, src));
      }
    }
    const int rcr = addAssignmentRules();
    if (rcr) return rcr;
    std::vector<double> table(constAddr.size() ? constAddr.size() : 1, 0.0);
    for (size_t i = 0; i < constAddr.size(); ++i) table[i] = *constAddr[i];
    DEV(sb2_set_constants(dev, table.data(), (int)table.size()));
    DEV(sb2_finalize(dev));
    return 0;
  }

  // what a bridged SpatialSimulator::oneStep does: constants the caller may have changed (setParameter), then the device step
  int step(double t, double dt) {
    for (size_t i = 0; i < constAddr.size(); ++i)
      if (constAddr[i] != &ref.sim_time) DEV(sb2_set_constant(dev, (int)i, *constAddr[i]));
    DEV(sb2_step(dev, t, 0.0, dt, 1, timeSlot));  // sim_time = t * dt inside, like performStep (:1358)
    return 0;
  }
};

static void onCrash(int sig) {  // REF_BRIDGE_VERBOSE: where a crash happened, then the default action
  void* frames[48];
  const int n = backtrace(frames, 48);
  const char msg[] = "ref_bridge: fatal signal, backtrace:\n";
  if (write(2, msg, sizeof(msg) - 1) < 0) {}
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}

int main(int argc, char** argv) {
  if (getenv("REF_BRIDGE_VERBOSE")) { signal(SIGSEGV, onCrash); signal(SIGABRT, onCrash); }
  // The reference's set-up reads isDomain[] of a membrane's neighbours without bounds tests (one row / plane past either end of the
  // array at points on the frame, spatialsimulator.cpp:719-1038).  With large arrays in mappings of their own those reads can fault;
  // kept on the ordinary heap they land in neighbouring allocations, as they do in the reference's own executables.
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TOP_PAD, 64 << 20);
  if (argc < 7) { fprintf(stderr, "usage: ref_bridge <model.xml> <xdim> <ydim> <zdim> <dt> <steps>\n"); return 64; }
  const int xd = atoi(argv[2]), yd = atoi(argv[3]), zd = atoi(argv[4]);
  const double dt = atof(argv[5]);
  const int steps = atoi(argv[6]);
  // two identical simulators of the reference: `host` is set up by the reference and stepped on the GPU through the bridge,
  // `cpu` is the reference stepping itself
  // (static: zero-initialised before their constructors run.  The reference's constructor leaves sim_time unset and initFromFile
  // evaluates the assignment rules with it - `time` in a rule reads whatever the object's storage held; on the stack that differed
  // between the two objects and with the process environment)
  static SpatialSimulator host, cpu;
  host.initFromFile(argv[1], xd, yd, zd);
  cpu.initFromFile(argv[1], xd, yd, zd);
  Bridge b(host);
  const bool verbose = getenv("REF_BRIDGE_VERBOSE") != NULL;  // per step and variable: the error, to stderr
  if (verbose) fprintf(stderr, "reference set-up done (twice)\n");
  int rc = b.build(0);
  if (rc) return rc;
  if (verbose) fprintf(stderr, "device model built\n");
  double worst = 0.0;
  bool identical = true;
  std::string worstVar = "-";
  int firstBad = -1;
  std::vector<double> got((size_t)b.nx * b.ny * b.nz);
  for (int t = 0; t < steps; ++t) {
    cpu.oneStep(t, dt);
    if (verbose) fprintf(stderr, "step %d: reference stepped\n", t);
    if ((rc = b.step(t, dt))) return rc;
    if (verbose) fprintf(stderr, "step %d: device stepped\n", t);
    for (std::map<variableInfo*, int>::iterator it = b.speciesId.begin(); it != b.speciesId.end(); ++it) {
      if (sb2_download_species(b.dev, it->second, got.data()) < 0) { fprintf(stderr, "ref_bridge: download failed\n"); return 2; }
      int len = 0;
      const double* want = cpu.getVariable(it->first->id, len);
      double scale = 1e-300, err = 0.0;
      for (int k = 0; k < b.nz; ++k)
        for (int j = 0; j < b.ny; ++j)
          for (int i = 0; i < b.nx; ++i) {
            const double w = want[b.even(i, j, k)], g = got[((size_t)k * b.ny + j) * b.nx + i];
            if (fabs(w) > scale) scale = fabs(w);
            if (fabs(g - w) > err) err = fabs(g - w);
            if (memcmp(&w, &g, sizeof(double)) != 0 && !(w == 0.0 && g == 0.0)) identical = false;
          }
      if (err / scale > worst) { worst = err / scale; worstVar = it->first->id; }
      if (err / scale > 1e-10 && firstBad < 0) firstBad = t;
      if (verbose) fprintf(stderr, "step %d %s rel %.3e\n", t, it->first->id, err / scale);
    }
    // species on membranes: every storage point (membrane and pseudo-membrane points)
    for (std::map<variableInfo*, int>::iterator it = b.memSpeciesId.begin(); it != b.memSpeciesId.end(); ++it) {
      const Bridge::PointSet& ps = b.pointSets[it->first->geoi];
      std::vector<double> mg(ps.storage.size() ? ps.storage.size() : 1);
      if (sb2_download_mem_species(b.dev, it->second, mg.data()) < 0) { fprintf(stderr, "ref_bridge: download failed\n"); return 2; }
      int len = 0;
      const double* want = cpu.getVariable(it->first->id, len);
      double scale = 1e-300, err = 0.0;
      for (size_t k = 0; k < ps.storage.size(); ++k) {
        const double w = want[ps.storage[k]], g = mg[k];
        if (fabs(w) > scale) scale = fabs(w);
        if (fabs(g - w) > err) err = fabs(g - w);
        if (memcmp(&w, &g, sizeof(double)) != 0 && !(w == 0.0 && g == 0.0)) identical = false;
      }
      if (err / scale > worst) { worst = err / scale; worstVar = it->first->id; }
      if (err / scale > 1e-10 && firstBad < 0) firstBad = t;
      if (verbose) fprintf(stderr, "step %d %s rel %.3e\n", t, it->first->id, err / scale);
    }
    // variables an assignment rule rewrites every step, at the volume cells
    for (size_t r = 0; r < b.ruleFields.size(); ++r) {
      variableInfo* v = b.ruleFields[r];
      if (sb2_download_field(b.dev, b.fieldId[v], got.data()) < 0) { fprintf(stderr, "ref_bridge: download failed\n"); return 2; }
      int len = 0;
      const double* want = cpu.getVariable(v->id, len);
      if (!want) continue;
      GeometryInfo* g = v->sp ? searchAvolInfoByCompartment(host.geoInfoList, v->sp->getCompartment().c_str()) : NULL;
      double scale = 1e-300, err = 0.0;
      for (int k = 0; k < b.nz; ++k)
        for (int j = 0; j < b.ny; ++j)
          for (int i = 0; i < b.nx; ++i) {
            if (g && g->isDomain[b.even(i, j, k)] != 1) continue;
            const double w = want[b.even(i, j, k)], gg = got[((size_t)k * b.ny + j) * b.nx + i];
            if (fabs(w) > scale) scale = fabs(w);
            if (fabs(gg - w) > err) err = fabs(gg - w);
            if (memcmp(&w, &gg, sizeof(double)) != 0 && !(w == 0.0 && gg == 0.0)) identical = false;
          }
      if (err / scale > worst) { worst = err / scale; worstVar = v->id; }
      if (err / scale > 1e-10 && firstBad < 0) firstBad = t;
      if (verbose) fprintf(stderr, "step %d %s rel %.3e\n", t, v->id, err / scale);
    }
  }
  printf("{\"bridge\": \"reference host set-up -> sb2_* C ABI -> GPU step\", \"model\": \"%s\", \"dims\": [%d, %d, %d], \"steps\": %d, "
         "\"species\": %d, \"membrane_species\": %d, \"rule_fields\": %d, \"constants\": %d, \"launches\": %lld, \"rel_linf_vs_reference_cpu\": %.3e, \"worst_variable\": \"%s\", \"first_step_over_1e-10\": %d, \"bit_identical\": %s}\n",
         argv[1], b.nx, b.ny, b.nz, steps, (int)b.speciesId.size(), (int)b.memSpeciesId.size(), (int)b.ruleFields.size(), (int)b.constAddr.size(), (long long)sb2_launch_count(b.dev), worst, worstVar.c_str(), firstBad,
         identical ? "true" : "false");
  sb2_destroy(b.dev);
  return worst <= 1e-10 ? 0 : 1;
}
