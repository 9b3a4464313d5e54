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
// boundary conditions, CIP-CSLR advection with uniform and field-valued velocities, membrane transport between two volumes.
// Refused (exit 3): species on membranes (and binding to them), assignment rules — the
// library has them (see csrc/host/spatialsimulator.cpp for the same flattening written against our own host model), the
// bridge keeps to what the reference's own example models (BASELINE configs 1-4) need.
#define private public  // (test-only access to varInfoList / geoInfoList / rInfoList, as in ref_driver.cpp)
#include "spatialsimulator.h"
#undef private

#include <sbml/SBMLTypes.h>
#include <sbml/packages/spatial/common/SpatialExtensionTypes.h>

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
  explicit Bridge(SpatialSimulator& r) : ref(r), dev(NULL), timeSlot(-1) {}

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
    if (!rg || !pg || !rg->isVol || !pg->isVol) { fprintf(stderr, "ref_bridge: binding to a membrane species is outside the bridge's scope\n"); return 3; }
    GeometryInfo* mem = NULL;
    for (size_t j = 0; j < ref.geoInfoList.size() && !mem; ++j) {
      GeometryInfo* g = ref.geoInfoList[j];
      if (!g->isVol && ((g->adjacentGeo1 == rg && g->adjacentGeo2 == pg) || (g->adjacentGeo1 == pg && g->adjacentGeo2 == rg))) mem = g;
    }
    if (!mem) return 0;  // no membrane between the two compartments: nothing happens
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
        if (v && rp->deltaList && rp->deltaList[i] && speciesId.count(v)) {
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
      sb2_ref rr;
      rr.species = ok ? speciesId[v] : -1;
      rr.is_variable = ri->isVariable[k] ? 1 : 0;
      rr.stoichiometry = ri->srStoichiometry[k];
      refs.push_back(rr);
      for (int p = 0; p < npts; ++p) {
        const int a = dpts[p].axis, dx = a == 0, dy = a == 1, dz = a == 2;
        targets.push_back(ok ? cellIn(v->geoi, pts[p].X + dx, pts[p].Y + dy, pts[p].Z + dz) : -1);
        targets.push_back(ok ? cellIn(v->geoi, pts[p].X - dx, pts[p].Y - dy, pts[p].Z - dz) : -1);
      }
    }
    DEV(sb2_add_mem_transport(dev, toks.data(), (int)toks.size(), refs.data(), (int)refs.size(), (int)r->getNumReactants(), dpts.data(), npts,
                              operandCells.empty() ? NULL : operandCells.data(), operandValues.empty() ? NULL : operandValues.data(),
                              targets.empty() ? NULL : targets.data()));
    return 0;
  }

  int build(int device) {
    nx = ref.Xdiv; ny = ref.Ydiv; nz = ref.Zdiv;
    if (ref.dimension < 2) { fprintf(stderr, "ref_bridge: 2-D and 3-D models only\n"); return 3; }
    if (!ref.orderedARule.empty()) { fprintf(stderr, "ref_bridge: assignment rules are outside the bridge's scope\n"); return 3; }
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
    }
    // species with a time derivative: value[] + delta[4N] (setInfoFunction.cpp:121-126)
    for (size_t i = 0; i < ref.varInfoList.size(); ++i) {
      variableInfo* v = ref.varInfoList[i];
      if (!v->sp || v->isUniform || !v->delta) continue;
      if (!v->inVol) { fprintf(stderr, "ref_bridge: species on membranes are outside the bridge's scope\n"); return 3; }
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
      if (!geo || !geoId.count(geo)) continue;  // every participant uniform: the reference changes nothing
      const bool isRule = ri->reaction == NULL;
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

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: ref_bridge <model.xml> <xdim> <ydim> <zdim> <dt> <steps>\n"); return 64; }
  const int xd = atoi(argv[2]), yd = atoi(argv[3]), zd = atoi(argv[4]);
  const double dt = atof(argv[5]);
  const int steps = atoi(argv[6]);
  // two identical simulators of the reference: `host` is set up by the reference and stepped on the GPU through the bridge,
  // `cpu` is the reference stepping itself
  SpatialSimulator host, cpu;
  host.initFromFile(argv[1], xd, yd, zd);
  cpu.initFromFile(argv[1], xd, yd, zd);
  Bridge b(host);
  int rc = b.build(0);
  if (rc) return rc;
  double worst = 0.0;
  bool identical = true;
  std::vector<double> got((size_t)b.nx * b.ny * b.nz);
  for (int t = 0; t < steps; ++t) {
    cpu.oneStep(t, dt);
    if ((rc = b.step(t, dt))) return rc;
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
      if (err / scale > worst) worst = err / scale;
    }
  }
  printf("{\"bridge\": \"reference host set-up -> sb2_* C ABI -> GPU step\", \"model\": \"%s\", \"dims\": [%d, %d, %d], \"steps\": %d, "
         "\"species\": %d, \"constants\": %d, \"launches\": %lld, \"rel_linf_vs_reference_cpu\": %.3e, \"bit_identical\": %s}\n",
         argv[1], b.nx, b.ny, b.nz, steps, (int)b.speciesId.size(), (int)b.constAddr.size(), (long long)sb2_launch_count(b.dev), worst,
         identical ? "true" : "false");
  sb2_destroy(b.dev);
  return worst <= 1e-10 ? 0 : 1;
}
