// host_model.cpp — model set-up on compact arrays (see host_model.h).
//
// Each section cites the reference code whose observable result it has to reproduce.  The data
// layout is different by design: volume masks are one byte per volume cell, membranes are sparse
// point lists keyed by doubled-grid index, and nothing of size (2n-1)^d is ever allocated here.
#include "host_model.h"
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>

#include "../../../include/spatial_b200.h"

namespace sb2host {

namespace {
// f(begin, end) on disjoint ranges of [0, n), on several host threads when the range is large (set-up of 10^7 .. 10^8
// cells); the loops it is used for write only to the cells of their own range
template <class F>
void parallelRanges(int64_t n, int64_t grain, F f) {
  unsigned nthreads = std::thread::hardware_concurrency();
  if (const char* e = getenv("SSB_HOST_THREADS")) nthreads = (unsigned)atoi(e);
  if (nthreads > 32) nthreads = 32;
  if (nthreads <= 1 || n < grain * 2) { f((int64_t)0, n); return; }
  const int64_t chunk = (n + nthreads - 1) / nthreads;
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nthreads; ++t) {
    const int64_t b = (int64_t)t * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    pool.emplace_back([=] { f(b, e); });
  }
  for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
}
}  // namespace


Var::Var()
    : sp(NULL), com(NULL), para(NULL), hasValue(false), isUniform(false), isResolved(false), inVol(true),
      hasDelta(false), hasAssignmentRule(false), uniformValue(0.0), offGridInit(0.0), geo(NULL), hasDiff(false),
      hasAdv(false), hasBC(false), hasRpn(false), onMembrane(false), constSlot(-1), fieldId(-1), speciesId(-1), memSpeciesId(-1) {
  for (int i = 0; i < 3; ++i) { diffC[i] = NULL; advC[i] = NULL; }
  for (int i = 0; i < 6; ++i) { bc[i] = NULL; bcKind[i] = SB2_BC_NONE; }
}

Model::Model()
    : doc(NULL), model(NULL), nx(1), ny(1), nz(1), Xindex(1), Yindex(1), Zindex(1), dimension(0), volDimension(0),
      memDimension(0), deltaX(0), deltaY(0), deltaZ(0), Xsize(0), Ysize(0), Zsize(0), simTime(0), xVar(NULL),
      yVar(NULL), zVar(NULL), tVar(NULL), slab(false) {}

Model::~Model() {
  for (size_t i = 0; i < vars.size(); ++i) delete vars[i];
  for (size_t i = 0; i < geos.size(); ++i) delete geos[i];
  for (size_t i = 0; i < rxns.size(); ++i) delete rxns[i];
  for (size_t i = 0; i < fastRxns.size(); ++i) delete fastRxns[i];
}

// searchInfoById (searchFunction.cpp:9-19): first match in list order wins
Var* Model::findVar(const std::string& id) const {
  for (size_t i = 0; i < vars.size(); ++i) if (vars[i]->id == id) return vars[i];
  return NULL;
}
Geo* Model::geoByCompartment(const std::string& id) const {
  for (size_t i = 0; i < geos.size(); ++i) if (geos[i]->compartmentId == id) return geos[i];
  return NULL;
}
Geo* Model::geoByDomainType(const std::string& id) const {
  for (size_t i = 0; i < geos.size(); ++i) if (geos[i]->domainTypeId == id) return geos[i];
  return NULL;
}

int Geo::storageOf(int64_t idx) const {
  std::vector<int64_t>::const_iterator it = std::lower_bound(storage.begin(), storage.end(), idx);
  return (it != storage.end() && *it == idx) ? (int)(it - storage.begin()) : -1;
}

// What the reference's full doubled-grid array of a variable holds at one point when the model is set up: coordinates are
// analytic, a parameter with an initial assignment / assignment rule was evaluated at EVERY point (isAllArea, calcPDE.cpp:28),
// a volume species was flooded with its initial value and later assigned / zeroed at volume cells only, a membrane species
// lives on the points of its membrane.
double Model::valueAtDoubled(const Var* v, int X, int Y, int Z, int depth) const {
  if (!v || !v->hasValue) return 0.0;
  if (v->isUniform) return v->uniformValue;
  if (v == xVar) return coordinateAt(v, X);
  if (v == yVar) return coordinateAt(v, Y);
  if (v == zVar) return coordinateAt(v, Z);
  const int64_t idx = ((int64_t)Z * Yindex + Y) * Xindex + X;
  if (v->onMembrane) {
    const int k = v->geo ? v->geo->storageOf(idx) : -1;
    return k >= 0 && (size_t)k < v->field.size() ? v->field[k] : 0.0;
  }
  if (!((X | Y | Z) & 1)) {
    const int i = X / 2, j = Y / 2 - off[1], k = Z / 2 - off[2];
    if (i >= 0 && i < nx && j >= 0 && j < ny && k >= 0 && k < nz) return v->field[((int64_t)k * ny + j) * nx + i];
    return v->offGridInit;
  }
  if (!v->sp && v->hasRpn && depth < 8) {
    double st[64] = {0}, out = v->offGridInit;
    evalRpnAtPoint(v->rpn, X, Y, Z, st, &out, depth + 1);
    return out;
  }
  return v->offGridInit;
}

bool Model::evalRpnAtPoint(const Rpn& rpn, int X, int Y, int Z, double* st, double* out, int depth) const {
  int sp = 0;
  for (size_t i = 0; i < rpn.size(); ++i) {
    const Tok& t = rpn[i];
    if (t.kind == Tok::VAR) { st[sp++] = valueAtDoubled(t.var, X, Y, Z, depth); if (sp > 60) sp = 60; continue; }
    if (t.kind == Tok::CONSTVAR) { st[sp++] = t.var->uniformValue; if (sp > 60) sp = 60; continue; }
    if (t.kind == Tok::LITERAL) { st[sp++] = t.literal; if (sp > 60) sp = 60; continue; }
    if (sp > 0) --sp;
    if (t.kind != Tok::OP) continue;
    rpnApplyOp(t.astType, st, sp);
    if (sp > 60) sp = 60;
  }
  if (sp > 0) { *out = st[--sp]; return true; }
  return false;
}

namespace {

SpatialParameterPlugin* spatialOf(Parameter* p) { return static_cast<SpatialParameterPlugin*>(p->getPlugin("spatial")); }

// setInfoFunction.cpp:15-28
int axisIndex(CoordinateKind_t kind) {
  switch (kind) {
    case SPATIAL_COORDINATEKIND_CARTESIAN_Y: return 1;
    case SPATIAL_COORDINATEKIND_CARTESIAN_Z: return 2;
    default: return 0;
  }
}

// setInfoFunction.cpp:33-70
bool containsSpatialMath(const ASTNode* node, ::Model* model) {
  for (unsigned int i = 0; i < node->getNumChildren(); ++i)
    if (containsSpatialMath(node->getChild(i), model)) return true;
  if (!node->isName()) return false;
  SBase* el = model->getElementBySId(node->getName());
  if (!el || el->getTypeCode() != SBML_SPECIES) return false;
  SpatialSpeciesPlugin* plug = dynamic_cast<SpatialSpeciesPlugin*>(el->getPlugin("spatial"));
  return plug && plug->getIsSpatial();
}
bool dependencesAreSpatial(Species* s, ::Model* model) {
  if (!model || !s || !s->isSetId()) return false;
  Rule* rule = model->getRule(s->getId());
  if (!rule || rule->getTypeCode() != SBML_ASSIGNMENT_RULE || !rule->isSetMath()) return false;
  return containsSpatialMath(rule->getMath(), model);
}

bool speciesIsVariable(Species* s) { return !s->isSetConstant() || !s->getConstant(); }

// spatialsimulator.cpp:101-117
std::string firstCompartmentId(::Model* model, const ASTNode* ast) {
  for (unsigned int i = 0; i < ast->getNumChildren(); ++i) {
    std::string id = firstCompartmentId(model, ast->getChild(i));
    if (!id.empty()) return id;
  }
  if (!ast->isName()) return "";
  if (model->getCompartment(ast->getName()) != NULL) return ast->getName();
  return "";
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// variables
// ---------------------------------------------------------------------------------------------

// setCompartmentInfo, setInfoFunction.cpp:72-91
void Model::setCompartments() {
  for (unsigned int i = 0; i < model->getNumCompartments(); ++i) {
    Compartment* c = model->getCompartment(i);
    Var* v = new Var;
    vars.push_back(v);
    v->com = c;
    v->id = c->getId();
    if (model->getRule(v->id) == NULL) {
      v->isResolved = true;
      v->isUniform = true;
      v->hasValue = true;
      v->uniformValue = c->isSetSize() ? c->getSize() : 1.0;
    }
  }
}

// setSpeciesInfo, setInfoFunction.cpp:93-155
void Model::setSpecies() {
  for (unsigned int i = 0; i < model->getNumSpecies(); ++i) {
    Species* s = model->getSpecies(i);
    SpatialSpeciesPlugin* plug = static_cast<SpatialSpeciesPlugin*>(s->getPlugin("spatial"));
    Var* v = new Var;
    vars.push_back(v);
    v->sp = s;
    v->com = model->getCompartment(s->getCompartment());
    v->id = s->getId();
    if (v->com) {
      if (v->com->getSpatialDimensions() == volDimension) v->inVol = true;
      else if (v->com->getSpatialDimensions() == memDimension) v->inVol = false;
    }
    if ((plug && plug->getIsSpatial()) || dependencesAreSpatial(s, model)) {
      if (s->isSetInitialAmount() || s->isSetInitialConcentration()) {
        const double init = s->isSetInitialAmount() ? s->getInitialAmount() : s->getInitialConcentration();
        v->hasValue = true;
        v->onMembrane = !v->inVol && v->com && memDimension != 0 && v->com->getSpatialDimensions() == memDimension;
        if (!v->onMembrane) v->field.assign(ncells(), init);  // the reference floods every doubled-grid point
        v->offGridInit = init;                                // (membrane species: sized by allocateMembraneFields)
        v->hasDelta = speciesIsVariable(s);
        v->isResolved = true;
      }
    } else {
      v->isResolved = true;
      v->isUniform = true;
      v->hasValue = true;
      v->uniformValue = s->isSetInitialConcentration() ? s->getInitialConcentration() : s->getInitialAmount();
    }
  }
}

// setParameterInfo, setInfoFunction.cpp:157-350
bool Model::setParameters() {
  SpatialModelPlugin* mp = static_cast<SpatialModelPlugin*>(model->getPlugin("spatial"));
  Geometry* geometry = mp->getGeometry();
  std::string sideId[6];  // Xmax, Xmin, Ymax, Ymin, Zmax, Zmin (mystruct.h:12-14)
  for (unsigned int i = 0; i < geometry->getNumCoordinateComponents(); ++i) {
    CoordinateComponent* cc = geometry->getCoordinateComponent(i);
    const int a = cc->getType() == SPATIAL_COORDINATEKIND_CARTESIAN_X   ? 0
                  : cc->getType() == SPATIAL_COORDINATEKIND_CARTESIAN_Y ? 1
                  : cc->getType() == SPATIAL_COORDINATEKIND_CARTESIAN_Z ? 2
                                                                         : -1;
    if (a < 0) continue;
    sideId[2 * a] = cc->getBoundaryMax()->getId();
    sideId[2 * a + 1] = cc->getBoundaryMin()->getId();
  }
  for (unsigned int i = 0; i < model->getNumParameters(); ++i) {
    Parameter* p = model->getParameter(i);
    SpatialParameterPlugin* pp = spatialOf(p);
    Var* v = new Var;
    vars.push_back(v);
    v->para = p;
    v->id = p->getId();
    const bool plainValue = model->getRule(v->id) == NULL && p->isSetValue();
    if (!pp || !pp->isSpatialParameter()) {
      if (plainValue) { v->isResolved = true; v->isUniform = true; v->hasValue = true; v->uniformValue = p->getValue(); }
      continue;
    }
    switch (pp->getType()) {
      case SBML_SPATIAL_DIFFUSIONCOEFFICIENT: {
        DiffusionCoefficient* dc = pp->getDiffusionCoefficient();
        Var* s = findVar(dc->getVariable());
        if (!s) { error = "diffusion coefficient " + v->id + " refers to unknown variable " + dc->getVariable(); return false; }
        s->hasDiff = true;
        if (!dc->isSetCoordinateReference1()) {
          for (int a = 0; a < 3; ++a) s->diffC[a] = v;  // no axis given: all three axes
        } else {
          s->diffC[axisIndex(dc->getCoordinateReference1())] = v;
        }
        if (plainValue) { v->isResolved = true; v->isUniform = true; v->hasValue = true; v->uniformValue = p->getValue(); }
      } break;
      case SBML_SPATIAL_ADVECTIONCOEFFICIENT: {
        AdvectionCoefficient* ac = pp->getAdvectionCoefficient();
        Var* s = findVar(ac->getVariable());
        if (!s) { error = "advection coefficient " + v->id + " refers to unknown variable " + ac->getVariable(); return false; }
        s->hasAdv = true;
        s->advC[axisIndex(ac->getCoordinate())] = v;
        if (plainValue) { v->isResolved = true; v->isUniform = true; v->hasValue = true; v->uniformValue = p->getValue(); }
      } break;
      case SBML_SPATIAL_BOUNDARYCONDITION: {
        BoundaryCondition* bcon = pp->getBoundaryCondition();
        Var* s = findVar(bcon->getVariable());
        if (!s) { error = "boundary condition " + v->id + " refers to unknown variable " + bcon->getVariable(); return false; }
        s->hasBC = true;
        int side = -1;
        for (int k = 0; k < 6; ++k)
          if (!sideId[k].empty() && bcon->getCoordinateBoundary() == sideId[k]) side = k;  // last match wins
        if (side >= 0) {
          s->bc[side] = v;
          s->bcKind[side] = bcon->getType() == SPATIAL_BOUNDARYKIND_NEUMANN     ? SB2_BC_NEUMANN
                            : bcon->getType() == SPATIAL_BOUNDARYKIND_DIRICHLET ? SB2_BC_DIRICHLET
                                                                                : SB2_BC_NONE;
          if (plainValue) { v->isResolved = true; v->isUniform = true; v->hasValue = true; v->uniformValue = p->getValue(); }
        }
      } break;
      case SBML_SPATIAL_SPATIALSYMBOLREFERENCE: {
        CoordinateComponent* cc = geometry->getCoordinateComponent(p->getId());
        if (!cc) {
          SpatialSymbolReference* ref = pp->getSpatialSymbolReference();
          cc = ref ? geometry->getCoordinateComponent(ref->getSpatialRef()) : NULL;
        }
        if (!cc) break;
        const double mn = cc->getBoundaryMin()->getValue();
        const double mx = cc->getBoundaryMax()->getValue();
        int a = -1;
        if (cc->getType() == SPATIAL_COORDINATEKIND_CARTESIAN_X) a = 0;
        else if (cc->getType() == SPATIAL_COORDINATEKIND_CARTESIAN_Y) a = 1;
        else if (cc->getType() == SPATIAL_COORDINATEKIND_CARTESIAN_Z) a = 2;
        if (a < 0) break;
        const double delta = (mx - mn) / (gn[a] - 1);
        if (a == 0) { xVar = v; Xsize = mx - mn; deltaX = delta; }
        if (a == 1) { yVar = v; Ysize = mx - mn; deltaY = delta; }
        if (a == 2) { zVar = v; Zsize = mx - mn; deltaZ = delta; }
        v->hasValue = true;
        v->isResolved = true;
        v->field.resize(ncells());
        // value at doubled-grid coordinate C is min + C*delta/2.0 (setInfoFunction.cpp:308); cell c sits at C = 2c
        for (int k = 0; k < nz; ++k)
          for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
              const int c = (a == 0 ? i : a == 1 ? j : k) + off[a];
              v->field[((int64_t)k * ny + j) * nx + i] = mn + (double)(2 * c) * delta / 2.0;
            }
        v->offGridInit = mn;  // only meaningful through coordinateAt()
      } break;
      default: break;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// volume geometry: masks, layering, boundary cells (spatialsimulator.cpp:411-717)
// ---------------------------------------------------------------------------------------------
bool Model::setVolumeGeometries() {
  SpatialModelPlugin* mp = static_cast<SpatialModelPlugin*>(model->getPlugin("spatial"));
  Geometry* geometry = mp->getGeometry();
  const int64_t nc = ncells();
  std::vector<double> tmp;
  const bool timing = getenv("SSB_TIMING") != NULL;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[ssb]   geo: %-18s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  };
  for (unsigned int gi = 0; gi < geometry->getNumGeometryDefinitions(); ++gi) {
    GeometryDefinition* gd = geometry->getGeometryDefinition(gi);
    if (gd->isAnalyticGeometry()) {
      AnalyticGeometry* ag = static_cast<AnalyticGeometry*>(gd);
      for (unsigned int j = 0; j < model->getNumCompartments(); ++j) {
        Compartment* c = model->getCompartment(j);
        SpatialCompartmentPlugin* cp = static_cast<SpatialCompartmentPlugin*>(c->getPlugin("spatial"));
        if (!cp || !cp->getCompartmentMapping()) continue;
        const std::string& dt = cp->getCompartmentMapping()->getDomainType();
        AnalyticVolume* av = NULL;
        for (unsigned int k = 0; k < ag->getNumAnalyticVolumes(); ++k)
          if (ag->getAnalyticVolume(k)->getDomainType() == dt) { av = ag->getAnalyticVolume(k); break; }
        if (!av || !av->getMath()) continue;
        Geo* g = new Geo;
        g->compartmentId = c->getId();
        g->domainTypeId = dt;
        g->isVol = true;
        ASTNode* ast = av->getMath()->deepCopy();
        rearrangeAst(ast);
        emitRpn(ast, *this, g->rpn);
        delete ast;
        lap("rpn");
        tmp.assign(nc, 0.0);
        lap("tmp.assign");
        evalRpnCompact(g->rpn, NULL, nc, tmp.data());
        lap("eval");
        g->isDomain.resize(nc);
        {
          const double* tp = tmp.data();
          auto* dp = g->isDomain.data();
          parallelRanges(nc, 1 << 18, [tp, dp](int64_t b, int64_t e) {
            for (int64_t k = b; k < e; ++k) {
              const int d = (int)tp[k];  // isDomain[k] = (int)tmp_isDomain[k], spatialsimulator.cpp:469
              dp[k] = d == 1 ? 1 : d == 0 ? 0 : 3;
            }
          });
        }
        g->hasMask = true;
        geos.push_back(g);
        lap("convert");
      }
    } else if (gd->isSampledFieldGeometry()) {
      // spatialsimulator.cpp:483-604: image rows are stored top-down (y flipped)
      SampledFieldGeometry* sg = static_cast<SampledFieldGeometry*>(gd);
      SampledField* sf = geometry->getSampledField(sg->getSampledField());
      if (!sf) continue;
      std::vector<double> samples(sf->getUncompressedLength());
      sf->getUncompressed(samples.data());
      for (unsigned int j = 0; j < model->getNumCompartments(); ++j) {
        Compartment* c = model->getCompartment(j);
        if (c->getSpatialDimensions() != volDimension) continue;
        SpatialCompartmentPlugin* cp = static_cast<SpatialCompartmentPlugin*>(c->getPlugin("spatial"));
        if (!cp || !cp->getCompartmentMapping()) continue;
        const std::string& dt = cp->getCompartmentMapping()->getDomainType();
        SampledVolume* sv = NULL;
        for (unsigned int k = 0; k < sg->getNumSampledVolumes(); ++k)
          if (sg->getSampledVolume(k)->getDomainType() == dt) sv = sg->getSampledVolume(k);
        if (!sv) continue;
        Geo* g = new Geo;
        g->compartmentId = c->getId();
        g->domainTypeId = dt;
        g->isVol = true;
        g->isDomain.assign(nc, 0);
        g->isBoundary.assign(nc, 0);
        g->flags.assign(nc, 0);
        g->hasMask = true;
        const int n1 = sf->getNumSamples1(), n2 = sf->getNumSamples2();
        for (int k = 0; k < nz; ++k)
          for (int jj = 0; jj < ny; ++jj)
            for (int i = 0; i < nx; ++i) {
              const int64_t s = (int64_t)(k + off[2]) * n2 * n1 + (int64_t)(n2 - (jj + off[1]) - 1) * n1 + i;
              if (s < 0 || s >= (int64_t)samples.size()) continue;
              if ((unsigned int)samples[s] == (unsigned int)sv->getSampledValue()) g->isDomain[((int64_t)k * ny + jj) * nx + i] = 1;
            }
        // boundary cells and boundary types straight away (spatialsimulator.cpp:551-600)
        for (int k = 0; k < nz; ++k)
          for (int jj = 0; jj < ny; ++jj)
            for (int i = 0; i < nx; ++i) {
              const int64_t c0 = ((int64_t)k * ny + jj) * nx + i;
              if (g->isDomain[c0] != 1) continue;
              g->nDomain++;
              const int gj = jj + off[1], gk = k + off[2];
              const bool frame = (dimension == 2 && (i == 0 || i == nx - 1 || gj == 0 || gj == gn[1] - 1)) ||
                                 (dimension == 3 && (i == 0 || i == nx - 1 || gj == 0 || gj == gn[1] - 1 || gk == 0 || gk == gn[2] - 1));
              uint8_t f = 0;
              if (frame) {
                g->isBoundary[c0] = 1;
                if (dimension >= 2) {
                  if (i == 0) f |= SB2_F_XM;
                  if (i == nx - 1) f |= SB2_F_XP;
                  if (gj == 0) f |= SB2_F_YM;
                  if (gj == gn[1] - 1) f |= SB2_F_YP;
                }
                if (dimension == 3) {
                  if (gk == 0) f |= SB2_F_ZM;
                  if (gk == gn[2] - 1) f |= SB2_F_ZP;
                }
              } else if (cutEdge(i, jj, k)) {
                f = 0;  // halo row / plane of a slab: owned by the neighbouring slab
              } else {
                if (dimension >= 2) {
                  if (g->isDomain[c0 + 1] == 0) f |= SB2_F_XP;
                  if (g->isDomain[c0 - 1] == 0) f |= SB2_F_XM;
                  if (g->isDomain[c0 + nx] == 0) f |= SB2_F_YP;
                  if (g->isDomain[c0 - nx] == 0) f |= SB2_F_YM;
                }
                if (dimension == 3) {
                  if (g->isDomain[c0 + (int64_t)nx * ny] == 0) f |= SB2_F_ZP;
                  if (g->isDomain[c0 - (int64_t)nx * ny] == 0) f |= SB2_F_ZM;
                }
                if (f) g->isBoundary[c0] = 1;
              }
              g->flags[c0] = f;
            }
        geos.push_back(g);
      }
    }
  }

  // layering by ordinal + domain / boundary cells (spatialsimulator.cpp:613-717)
  for (unsigned int gi = 0; gi < geometry->getNumGeometryDefinitions(); ++gi) {
    GeometryDefinition* gd = geometry->getGeometryDefinition(gi);
    if (!gd->isAnalyticGeometry()) continue;
    AnalyticGeometry* ag = static_cast<AnalyticGeometry*>(gd);
    for (unsigned int j = 0; j < ag->getNumAnalyticVolumes(); ++j) {
      AnalyticVolume* ex = ag->getAnalyticVolume(j);
      Geo* gex = geoByDomainType(ex->getDomainType());
      if (!gex || !gex->hasMask) continue;
      for (unsigned int k = 0; k < ag->getNumAnalyticVolumes(); ++k) {
        AnalyticVolume* in = ag->getAnalyticVolume(k);
        if (!(ex->getOrdinal() < in->getOrdinal())) continue;
        Geo* gin = geoByDomainType(in->getDomainType());
        if (!gin || !gin->hasMask) continue;
        {
          auto* ex_d = gex->isDomain.data();
          const auto* in_d = gin->isDomain.data();
          parallelRanges(nc, 1 << 18, [ex_d, in_d](int64_t b, int64_t e) {
            for (int64_t c = b; c < e; ++c)
              if (ex_d[c] == 1 && in_d[c] == 1) ex_d[c] = 0;
          });
        }
      }
      // NB the reference appends to domainIndex here; a domain type listed twice would be appended twice
      gex->nDomain = 0;
      gex->isBoundary.assign(nc, 0);
      const int64_t sy = nx, sz = (int64_t)nx * ny;
      std::atomic<int64_t> domainCount(0);
      // rows (k, jj) are independent: a row only reads isDomain and writes its own isBoundary entries
      parallelRanges((int64_t)nz * ny, 256, [&](int64_t rb, int64_t re) {
      int64_t localCount = 0;
      for (int64_t row = rb; row < re; ++row) {
          const int k = (int)(row / ny), jj = (int)(row % ny);
          for (int i = 0; i < nx; ++i) {
            const int64_t c = (int64_t)k * sz + (int64_t)jj * sy + i;
            if (gex->isDomain[c] != 1) continue;
            ++localCount;
            bool b = false;
            const int gj = jj + off[1], gk = k + off[2];
            if (dimension == 1) {
              b = (i == 0 || i == nx - 1) || gex->isDomain[c + 1] == 0 || gex->isDomain[c - 1] == 0;
            } else if (dimension == 2) {
              if (i == 0 || i == nx - 1 || gj == 0 || gj == gn[1] - 1) b = true;
              else if (cutEdge(i, jj, k)) b = false;  // halo row of a slab (owned elsewhere)
              else b = gex->isDomain[c + 1] == 0 || gex->isDomain[c - 1] == 0 || gex->isDomain[c + sy] == 0 || gex->isDomain[c - sy] == 0;
            } else if (dimension == 3) {
              if (i == 0 || i == nx - 1 || gj == 0 || gj == gn[1] - 1 || gk == 0 || gk == gn[2] - 1) b = true;
              else if (cutEdge(i, jj, k)) b = false;
              else b = gex->isDomain[c + 1] == 0 || gex->isDomain[c - 1] == 0 || gex->isDomain[c + sy] == 0 ||
                       gex->isDomain[c - sy] == 0 || gex->isDomain[c + sz] == 0 || gex->isDomain[c - sz] == 0;
            }
            gex->isBoundary[c] = b ? 1 : 0;
          }
      }
      domainCount += localCount;
      });
      gex->nDomain = domainCount.load();
      lap("layer+boundary");
    }
  }
  for (size_t g = 0; g < geos.size(); ++g)
    if (geos[g]->isVol && geos[g]->hasMask) {
      if (geos[g]->flags.empty()) geos[g]->flags.assign(nc, 0);
      if (geos[g]->isBoundary.empty()) geos[g]->isBoundary.assign(nc, 0);
    }
  return true;
}

// ---------------------------------------------------------------------------------------------
// membranes: compartments, adjacency, membrane / pseudo-membrane points (spatialsimulator.cpp:719-1038)
// ---------------------------------------------------------------------------------------------
void Model::setMembraneGeometries() {
  SpatialModelPlugin* mp = static_cast<SpatialModelPlugin*>(model->getPlugin("spatial"));
  Geometry* geometry = mp->getGeometry();
  for (unsigned int i = 0; i < model->getNumCompartments(); ++i) {
    Compartment* c = model->getCompartment(i);
    SpatialCompartmentPlugin* cp = static_cast<SpatialCompartmentPlugin*>(c->getPlugin("spatial"));
    if (!cp || !cp->getCompartmentMapping()) continue;
    if (geoByCompartment(c->getId())) continue;
    DomainType* dType = geometry->getDomainType(cp->getCompartmentMapping()->getDomainType());
    Geo* g = new Geo;
    g->compartmentId = c->getId();
    g->domainTypeId = cp->getCompartmentMapping()->getDomainType();
    if (!dType || dType->getSpatialDimensions() == volDimension) g->isVol = true;
    else if (dType->getSpatialDimensions() == memDimension) g->isVol = false;
    geos.push_back(g);
  }
  for (unsigned int i = 0; i < geometry->getNumAdjacentDomains(); ++i) {
    AdjacentDomains* ad = geometry->getAdjacentDomains(i);
    Domain* d1 = geometry->getDomain(ad->getDomain1());
    Domain* d2 = geometry->getDomain(ad->getDomain2());
    if (!d1 || !d2) continue;
    Geo* g1 = geoByDomainType(d1->getDomainType());
    Geo* g2 = geoByDomainType(d2->getDomainType());
    if (!g1 || !g2) continue;
    if (!g1->adjacent1) g1->adjacent1 = g2;
    else if (!g1->adjacent2 && g1->adjacent1 != g2) g1->adjacent2 = g2;
    if (!g2->adjacent1) g2->adjacent1 = g1;
    else if (!g2->adjacent2 && g2->adjacent1 != g1) g2->adjacent2 = g1;
  }

  const int64_t Xi = Xindex, Yi = Yindex;
  const int64_t sy = nx, sz = (int64_t)nx * ny;
  for (size_t gi = 0; gi < geos.size(); ++gi) {
    Geo* g = geos[gi];
    if (g->isVol) continue;
    g->hasMask = true;
    if (isSlab()) continue;  // membrane points of a sharded grid are not classified (no consumer is supported there)
    Geo* a1 = g->adjacent1;
    Geo* a2 = g->adjacent2;
    if (!a1 || !a2 || !a1->isVol || !a2->isVol || !a1->hasMask || !a2->hasMask) continue;
    const uint8_t* m1 = a1->isDomain.data();
    const uint8_t* m2 = a2->isDomain.data();
    // A candidate (odd flat index) lies between two axis-neighbouring volume cells; it is a membrane
    // point when one of them belongs to adjacentGeo1 and the other to adjacentGeo2.  The tests of
    // the other axes in the reference look at points that are never volume cells, so they cannot fire.
    auto between = [&](int64_t ca, int64_t cb) {
      return (m1[cb] == 1 && m2[ca] == 1) || (m1[ca] == 1 && m2[cb] == 1);
    };
    auto addPoint = [&](int X, int Y, int Z, uint8_t bof) {
      MemPoint p;
      p.X = X; p.Y = Y; p.Z = Z;
      p.index = (int64_t)Z * Yi * Xi + (int64_t)Y * Xi + X;
      p.bof = bof;
      p.nx = p.ny = p.nz = 0.0;
      g->memClass[p.index] = 1;
      g->memPoints.push_back(p);
      // the reference pushes X instead of the flat index on the Y-frame rows in 2-D (spatialsimulator.cpp:846)
      int64_t pushed = p.index;
      if (dimension == 2 && !(X != 0 && X != Xindex - 1 && Y != 0 && Y != Yindex - 1) && !(X == 0 || X == Xindex - 1) &&
          (Y == 0 || Y == Yindex - 1))
        pushed = X;
      g->memDomainIndex.push_back(pushed);
    };
    if (dimension == 1) {
      for (int X = 1; X < Xindex - 1; X += 2)
        if (between((X - 1) / 2, (X + 1) / 2)) addPoint(X, 0, 0, 0);
    } else {
      for (int Z = 0; Z < Zindex; ++Z)
        for (int Y = 0; Y < Yindex; ++Y) {
          const int oddZY = (Z & 1) + (Y & 1);
          if (oddZY == 0) {
            // X-odd points of an even row: between cells (i, i+1)
            const int64_t row = (int64_t)(Z / 2) * sz + (int64_t)(Y / 2) * sy;
            for (int X = 1; X < Xindex - 1; X += 2)
              if (between(row + (X - 1) / 2, row + (X + 1) / 2)) addPoint(X, Y, Z, SB2_F_XP | SB2_F_XM);
          } else if (oddZY == 1 && (Y & 1)) {
            const int64_t row = (int64_t)(Z / 2) * sz + (int64_t)((Y - 1) / 2) * sy;
            for (int X = 0; X < Xindex; X += 2)
              if (between(row + X / 2, row + sy + X / 2)) addPoint(X, Y, Z, SB2_F_YP | SB2_F_YM);
          } else if (oddZY == 1) {
            const int64_t row = (int64_t)((Z - 1) / 2) * sz + (int64_t)(Y / 2) * sy;
            for (int X = 0; X < Xindex; X += 2)
              if (between(row + X / 2, row + sz + X / 2)) addPoint(X, Y, Z, SB2_F_ZP | SB2_F_ZM);
          }
        }
    }
    if (dimension < 2) continue;

    // pseudo membrane: points with >= 2 odd coordinates next to membrane points
    std::set<int64_t> cand;
    const int64_t stride[3] = {1, Xi, Xi * Yi};
    const int lim[3] = {Xindex, Yindex, Zindex};
    for (size_t k = 0; k < g->memPoints.size(); ++k) {
      const MemPoint& p = g->memPoints[k];
      const int co[3] = {p.X, p.Y, p.Z};
      for (int a = 0; a < dimension; ++a)
        for (int s = -1; s <= 1; s += 2) {
          int q[3] = {co[0], co[1], co[2]};
          q[a] += s;
          if (q[a] < 0 || q[a] >= lim[a]) continue;
          const int odd = (q[0] & 1) + (q[1] & 1) + (q[2] & 1);
          if (odd >= 2) cand.insert(p.index + s * stride[a]);
        }
    }
    for (std::set<int64_t>::const_iterator it = cand.begin(); it != cand.end(); ++it) {
      const int64_t idx = *it;
      const int Z = (int)(idx / (Xi * Yi));
      const int Y = (int)((idx - (int64_t)Z * Xi * Yi) / Xi);
      const int X = (int)(idx - (int64_t)Z * Xi * Yi - (int64_t)Y * Xi);
      auto is1 = [&](int dx, int dy, int dz) {
        const int x = X + dx, y = Y + dy, z = Z + dz;
        if (x < 0 || y < 0 || z < 0 || x >= Xindex || y >= Yindex || z >= Zindex) return false;
        return g->memIsDomain(idx + dx + dy * Xi + dz * Xi * Yi) == 1;
      };
      const bool xp = is1(1, 0, 0), xm = is1(-1, 0, 0), yp = is1(0, 1, 0), ym = is1(0, -1, 0);
      if (dimension == 2) {
        if (!(X & 1) || !(Y & 1)) continue;
        // X odd and Y odd can never sit on the frame (Xindex, Yindex are odd)
        if ((xp && xm) || (yp && ym)) { g->memClass[idx] = 2; g->pseudoMemIndex.push_back(idx); }
        if ((xp && yp) || (xp && ym) || (yp && xm) || (ym && xm)) { g->memClass[idx] = 2; g->pseudoMemIndex.push_back(idx); }
      } else {
        const bool zp = is1(0, 0, 1), zm = is1(0, 0, -1);
        if (X != 0 && X != Xindex - 1 && Y != 0 && Y != Yindex - 1 && Z != 0 && Z != Zindex - 1) {
          if ((xp && xm) || (yp && ym) || (zp && zm)) { g->memClass[idx] = 2; g->pseudoMemIndex.push_back(idx); }
          if ((xp && yp) || (xp && ym) || (xp && zp) || (xp && zm) || (xm && yp) || (xm && ym) || (xm && zp) || (xm && zm) ||
              (yp && zp) || (yp && zm) || (ym && zp) || (ym && zm)) {
            g->memClass[idx] = 2;
            g->pseudoMemIndex.push_back(idx);
          }
        } else if (X == 0 || X == Xindex - 1) {
          if ((yp && ym) || (zp && zm) || (yp && zp) || (yp && zm) || (ym && zp) || (ym && zm)) {
            g->memClass[idx] = 2;
            g->pseudoMemIndex.push_back(idx);
          }
        } else if (Y == 0 || Y == Yindex - 1) {
          // the reference marks the point but does not list it (spatialsimulator.cpp:1013-1021)
          if ((xp && xm) || (zp && zm) || (yp && zp) || (yp && zm) || (ym && zp) || (ym && zm)) g->memClass[idx] = 2;
        } else if (Z == 0 || Z == Zindex - 1) {
          if ((xp && xm) || (yp && ym)) { g->memClass[idx] = 2; g->pseudoMemIndex.push_back(idx); }
        }
      }
    }
  }
}

// Storage of the species that live on a membrane: one value per membrane point, pseudo-membrane point and any other point the
// membrane's domainIndex names (its 2-D frame quirk), in ascending doubled-grid index.
void Model::allocateMembraneFields() {
  for (size_t gi = 0; gi < geos.size(); ++gi) {
    Geo* g = geos[gi];
    if (g->isVol) continue;
    std::set<int64_t> pts;
    for (size_t k = 0; k < g->memPoints.size(); ++k) pts.insert(g->memPoints[k].index);
    for (size_t k = 0; k < g->pseudoMemIndex.size(); ++k) pts.insert(g->pseudoMemIndex[k]);
    for (size_t k = 0; k < g->memDomainIndex.size(); ++k) pts.insert(g->memDomainIndex[k]);
    for (std::unordered_map<int64_t, uint8_t>::const_iterator it = g->memClass.begin(); it != g->memClass.end(); ++it) pts.insert(it->first);
    g->storage.assign(pts.begin(), pts.end());
  }
  for (size_t i = 0; i < vars.size(); ++i) {
    Var* v = vars[i];
    if (!v->onMembrane || !v->sp) continue;
    v->geo = geoByCompartment(v->sp->getCompartment());
    v->field.assign(v->geo ? v->geo->storage.size() : 0, v->offGridInit);
  }
}

// ---------------------------------------------------------------------------------------------
// initial assignments / assignment rules in dependency order (spatialsimulator.cpp:1051-1153)
// ---------------------------------------------------------------------------------------------
bool Model::resolveAssignments() {
  std::vector<Var*> pending;
  std::vector<const ASTNode*> pendingMath;
  const int64_t nc = ncells();
  for (size_t i = 0; i < vars.size(); ++i) {
    Var* v = vars[i];
    const ASTNode* math = NULL;
    InitialAssignment* ia = model->getInitialAssignment(v->id);
    Rule* rule = model->getRule(v->id);
    if (ia) {
      math = ia->getMath();
    } else if (rule && rule->isAssignment()) {
      v->hasAssignmentRule = true;
      math = rule->getMath();
    } else {
      continue;
    }
    if (!v->hasValue) {
      v->hasValue = true;
      v->onMembrane = v->sp && !v->inVol && v->com && memDimension != 0 && v->com->getSpatialDimensions() == memDimension;
      if (v->onMembrane) {
        v->geo = geoByCompartment(v->sp->getCompartment());
        v->field.assign(v->geo ? v->geo->storage.size() : 0, 0.0);
      } else {
        v->field.assign(nc, 0.0);
      }
      v->offGridInit = 0.0;
      if (v->sp && speciesIsVariable(v->sp)) v->hasDelta = true;
    }
    if (!math) continue;
    if (v->isUniform) {
      // a uniform variable is written through value[index] over the whole grid in the reference
      // (out of bounds); there is no sane behaviour to mirror
      error = "initial assignment / assignment rule on the non-spatial variable '" + v->id + "' is not supported";
      return false;
    }
    v->isResolved = false;
    collectDependence(math, *this, v->dependence);
    pending.push_back(v);
    pendingMath.push_back(math);
  }
  size_t guard = 0;
  for (size_t i = 0; i < pending.size(); ++i) {
    Var* v = pending[i];
    bool ready = !v->isResolved;
    for (size_t d = 0; ready && d < v->dependence.size(); ++d) ready = v->dependence[d]->isResolved;
    if (ready) {
      ASTNode* ast = pendingMath[i]->deepCopy();
      rearrangeAst(ast);
      v->rpn.clear();
      emitRpn(ast, *this, v->rpn);
      v->hasRpn = true;
      if (v->sp) v->geo = geoByCompartment(v->sp->getCompartment());
      else if (v->com) v->geo = geoByCompartment(v->com->getId());
      else {
        const std::string id = firstCompartmentId(model, ast);
        if (!id.empty()) v->geo = geoByCompartment(id);
      }
      delete ast;
      const bool allArea = v->sp == NULL;
      if (allArea) {
        evalRpnCompact(v->rpn, NULL, nc, v->field.data());
      } else if (v->geo && v->geo->isVol && v->geo->hasMask) {
        // species: only the cells of its compartment (isDomain == 1) are written
        std::vector<uint8_t> mask(nc);
        for (int64_t c = 0; c < nc; ++c) mask[c] = v->geo->isDomain[c] == 1;
        evalRpnCompact(v->rpn, mask.data(), nc, v->field.data());
      } else if (v->geo && !v->geo->isVol && v->onMembrane) {
        // membrane species: the points its membrane's domainIndex names, in that order, with one stack for the whole sweep
        double st[64] = {0};
        const int64_t Xi = Xindex, Yi = Yindex;
        for (size_t q = 0; q < v->geo->memDomainIndex.size(); ++q) {
          const int64_t idx = v->geo->memDomainIndex[q];
          const int k = v->geo->storageOf(idx);
          if (k < 0) continue;
          const int Z = (int)(idx / (Xi * Yi)), Y = (int)((idx - (int64_t)Z * Xi * Yi) / Xi), X = (int)(idx - (int64_t)Z * Xi * Yi - (int64_t)Y * Xi);
          evalRpnAtPoint(v->rpn, X, Y, Z, st, &v->field[k]);
        }
      }
      v->isResolved = true;
      if (v->hasAssignmentRule) orderedAssignmentRules.push_back(v);
    }
    if (i + 1 == pending.size()) {
      size_t resolved = 0;
      for (size_t j = 0; j < pending.size(); ++j) resolved += pending[j]->isResolved ? 1 : 0;
      if (resolved == pending.size()) break;
      if (++guard > pending.size() + 1) {
        error = "circular or unresolvable dependencies between initial assignments / assignment rules";
        return false;
      }
      i = (size_t)-1;  // restart the sweep
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// membrane normals by contour walking (setNormalAngle / stepSearch / oneStepSearch,
// setInfoFunction.cpp:495-1124)
// ---------------------------------------------------------------------------------------------
namespace {

enum Dir { DN = 0, DNE, DE, DSE, DS, DSW, DW, DNW, DNONE = -1 };
const int kDh[8] = {0, 1, 2, 1, 0, -1, -2, -1};
const int kDv[8] = {2, 1, 0, -1, -2, -1, 0, 1};
const int kOpp[8] = {DS, DSW, DW, DNW, DN, DNE, DE, DSE};

struct PlaneWalker {
  const Geo* g;
  int64_t total, sh, sv;  // doubled-grid size, strides of the horizontal / vertical axis of the plane
  int cls(int64_t idx) const { return (idx < 0 || idx >= total) ? 0 : g->memIsDomain(idx); }
  // can the contour continue from flat index `at` in direction d?
  bool open(int64_t at, int d) const {
    const int64_t to = at + kDh[d] * sh + kDv[d] * sv;
    if (cls(to) != 1) return false;
    if (d & 1) return true;                             // diagonal: the next membrane point itself
    return cls(at + (kDh[d] / 2) * sh + (kDv[d] / 2) * sv) == 2;  // straight: across a pseudo-membrane point
  }
  // stepSearch: keep walking until step_k points have been passed
  void walk(int64_t at, int pre, int count, int stepK, int& h, int& v) const {
    while (count != stepK) {
      int d = 0;
      for (; d < 8; ++d)
        if (pre != kOpp[d] && open(at, d)) break;
      if (d == 8) return;
      h += kDh[d]; v += kDv[d];
      at += kDh[d] * sh + kDv[d] * sv;
      pre = d;
      ++count;
    }
  }
  // oneStepSearch: the two contour points step_k steps away on either side
  void ends(int64_t at, int h0, int v0, int stepK, int hor[2], int ver[2]) const {
    int i = 0;
    while (i < 2) {
      bool any = false;
      for (int d = 0; d < 8; ++d) {
        if (!open(at, d)) continue;
        any = true;
        int h = h0 + kDh[d], v = v0 + kDv[d];
        walk(at + kDh[d] * sh + kDv[d] * sv, d, 1, stepK, h, v);
        hor[i] = h; ver[i] = v;
        if (i == 0) ++i; else return;
      }
      if (!any) return;  // isolated point: the reference would spin forever here
    }
  }
};

}  // namespace

void Model::setNormals() {
  if (dimension < 2) return;
  const double hX = Xsize / (nx - 1), hY = Ysize / (ny - 1);
  double hZ = 1.0, hXY = 1.0, hYZ = 1.0, hXZ = 1.0;
  hXY = (hX + hY) / 2.0;
  if (dimension == 3) {
    hZ = Zsize / (nz - 1);
    hYZ = (hY + hZ) / 2.0;
    hXZ = (hX + hZ) / 2.0;
  }
  const int64_t Xi = Xindex, Yi = Yindex;
  const int64_t total = ndoubled();
  for (size_t gi = 0; gi < geos.size(); ++gi) {
    Geo* g = geos[gi];
    if (g->isVol || g->memPoints.empty()) continue;
    // the reference iterates domainIndex, which (with its 2-D frame quirk) may name a point that is
    // not a membrane point; those entries only receive a garbage normal nobody reads
    std::map<int64_t, size_t> pointOf;
    for (size_t k = 0; k < g->memPoints.size(); ++k) pointOf[g->memPoints[k].index] = k;
    int planeH[3][2] = {{0, 0}, {0, 0}, {0, 0}}, planeV[3][2] = {{0, 0}, {0, 0}, {0, 0}};  // xy, yz, xz (persist across points)
    const size_t M = g->memDomainIndex.size();
    for (size_t j = 0; j < M; ++j) {
      const int64_t index = g->memDomainIndex[j];
      std::map<int64_t, size_t>::const_iterator pit = pointOf.find(index);
      if (pit == pointOf.end()) continue;
      MemPoint& p = g->memPoints[pit->second];
      const bool bx = (p.bof & SB2_F_XP) && (p.bof & SB2_F_XM);
      const bool by = (p.bof & SB2_F_YP) && (p.bof & SB2_F_YM);
      const bool bz = (p.bof & SB2_F_ZP) && (p.bof & SB2_F_ZM);
      const int co[3] = {p.X, p.Y, p.Z};
      // planes: 0 = xy (hor X, ver Y), 1 = yz (hor Y, ver Z), 2 = xz (hor X, ver Z)
      const int hAxis[3] = {0, 1, 0}, vAxis[3] = {1, 2, 2};
      const double hh[3] = {hX, hY, hX}, hv[3] = {hY, hZ, hZ}, hP[3] = {hXY, hYZ, hXZ};
      const bool use[3] = {bx || by, dimension == 3 && (by || bz), dimension == 3 && (bx || bz)};
      const int64_t stride[3] = {1, Xi, Xi * Yi};
      const int lim[3] = {Xindex, Yindex, Zindex};
      for (int pl = 0; pl < 3; ++pl) {
        if (!use[pl]) continue;
        PlaneWalker w;
        w.g = g; w.total = total; w.sh = stride[hAxis[pl]]; w.sv = stride[vAxis[pl]];
        // largest half-distance to any point of the closed contour through this point (:559-603)
        double maxRadius = 0.0;
        {
          int h = co[hAxis[pl]], v = co[vAxis[pl]];
          const int h0 = h, v0 = v;
          int64_t at = index;
          int pre = DNONE;
          for (size_t k = 0; k < M; ++k) {
            int d = 0;
            for (; d < 8; ++d) {
              if (pre == kOpp[d]) continue;
              // bounds the reference applies on the vertical coordinate only
              if (d == DN && !(v + 2 < lim[vAxis[pl]])) continue;
              if (d == DSE && pl == 0 && !(v > 0)) continue;
              if (d == DS && !(v - 2 >= 0)) continue;
              if (d == DSW && pl == 0 && !(v - 1 >= 0)) continue;
              if (w.open(at, d)) break;
            }
            if (d < 8) {
              h += kDh[d]; v += kDv[d];
              at += kDh[d] * w.sh + kDv[d] * w.sv;
              pre = d;
            }
            maxRadius = std::max(maxRadius, sqrt(pow(((h - h0) / 2) * hh[pl], 2) + pow(((v - v0) / 2) * hv[pl], 2)) / 2.0);
            if (h == h0 && v == v0) break;
          }
        }
        // local radius of curvature → number of contour steps used for the tangent (:704-735)
        int stepK = (int)(pow(hP[pl], -(2.0 / 3.0)) + 0.5);
        if (stepK == 0) stepK = 1;
        w.ends(index, co[hAxis[pl]], co[vAxis[pl]], stepK, planeH[pl], planeV[pl]);
        const double a = sqrt(pow(((co[hAxis[pl]] - planeH[pl][0]) * hh[pl]) / 2.0, 2) + pow(((co[vAxis[pl]] - planeV[pl][0]) * hv[pl]) / 2.0, 2));
        const double b = sqrt(pow(((co[hAxis[pl]] - planeH[pl][1]) * hh[pl]) / 2.0, 2) + pow(((co[vAxis[pl]] - planeV[pl][1]) * hv[pl]) / 2.0, 2));
        const double c = sqrt(pow(((planeH[pl][0] - planeH[pl][1]) * hh[pl]) / 2.0, 2) + pow(((planeV[pl][0] - planeV[pl][1]) * hv[pl]) / 2.0, 2));
        const bool straight = pl == 0 ? ((a + b + c) * (-a + b + c) * (a - b + c) * (a + b - c) == 0)
                                      : ((a + b + c) == 0 || (-a + b + c) == 0 || (a - b + c) == 0 || (a + b - c) == 0);
        if (straight) {
          stepK = 1;
        } else {
          const double rho = std::min(maxRadius, (a * b * c) / sqrt((a + b + c) * (-a + b + c) * (a - b + c) * (a + b - c)));
          stepK = (int)(pow((rho / hP[pl]), 2.0 / 3.0) + 0.5);
        }
        w.ends(index, co[hAxis[pl]], co[vAxis[pl]], stepK, planeH[pl], planeV[pl]);
      }
      // normal = tangent1 x tangent2 (:777-827)
      double X1 = 0, Y1 = 0, Z1 = 0, X2 = 0, Y2 = 0, Z2 = 0;
      if (dimension == 2) {
        X1 = planeH[0][0] - planeH[0][1];
        Y1 = planeV[0][0] - planeV[0][1];
        Z2 = 1.0;
      } else {
        if ((p.bof & SB2_F_XP) || (p.bof & SB2_F_XM)) {  // xy x xz
          X1 = planeH[0][0] - planeH[0][1]; Y1 = planeV[0][0] - planeV[0][1]; Z1 = 0.0;
          X2 = planeH[2][0] - planeH[2][1]; Y2 = 0.0; Z2 = planeV[2][0] - planeV[2][1];
        }
        if ((p.bof & SB2_F_YP) || (p.bof & SB2_F_YM)) {  // xy x yz
          X1 = planeH[0][0] - planeH[0][1]; Y1 = planeV[0][0] - planeV[0][1]; Z1 = 0.0;
          X2 = 0.0; Y2 = planeH[1][0] - planeH[1][1]; Z2 = planeV[1][0] - planeV[1][1];
        }
        if ((p.bof & SB2_F_ZP) || (p.bof & SB2_F_ZM)) {  // xz x yz
          X1 = planeH[2][0] - planeH[2][1]; Y1 = 0.0; Z1 = planeV[2][0] - planeV[2][1];
          X2 = 0.0; Y2 = planeH[1][0] - planeH[1][1]; Z2 = planeV[1][0] - planeV[1][1];
        }
      }
      p.nx = Y1 * Z2 - Z1 * Y2;
      p.ny = Z1 * X2 - X1 * Z2;
      p.nz = X1 * Y2 - Y1 * X2;
      double len = sqrt(pow(p.nx, 2) + pow(p.ny, 2) + pow(p.nz, 2));
      if (len == 0) len = 1;
      p.nx /= len;
      p.ny /= len;
      p.nz /= len;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Voronoi neighbours of the membrane points: contour neighbours one step away in each plane the point's membrane crosses,
// their distance d_ij measured in the tangent plane and (3-D) the facet length s_ij, then the symmetrising average
// (setVoronoiInfo, setInfoFunction.cpp:1134-1546)
// ---------------------------------------------------------------------------------------------
void Model::setVoronoi() {
  if (dimension < 2 || !xVar || !yVar || (dimension == 3 && !zVar)) return;
  const int64_t Xi = Xindex, Yi = Yindex;
  const int64_t total = ndoubled();
  const int64_t stride[3] = {1, Xi, Xi * Yi};
  struct P3 { double x, y, z; };
  for (size_t gi = 0; gi < geos.size(); ++gi) {
    Geo* g = geos[gi];
    if (g->isVol || g->memPoints.empty()) continue;
    std::map<int64_t, size_t> pointOf;
    for (size_t k = 0; k < g->memPoints.size(); ++k) pointOf[g->memPoints[k].index] = k;
    std::unordered_map<int64_t, std::vector<P3> > contour;  // planeAdjacent: projected neighbours, [XY0, XY1, YZ0, YZ1, XZ0, XZ1]
    int planeH[3][2] = {{0, 0}, {0, 0}, {0, 0}}, planeV[3][2] = {{0, 0}, {0, 0}, {0, 0}};  // persist across points like the reference's locals
    const size_t M = g->memDomainIndex.size();
    auto crosses = [](const MemPoint& p, int pl) {
      const bool bx = (p.bof & SB2_F_XP) && (p.bof & SB2_F_XM), by = (p.bof & SB2_F_YP) && (p.bof & SB2_F_YM),
                 bz = (p.bof & SB2_F_ZP) && (p.bof & SB2_F_ZM);
      return pl == 0 ? (bx || by) : pl == 1 ? (by || bz) : (bx || bz);
    };
    auto X_of = [&](int64_t idx) { return (int)(idx % Xi); };
    auto Y_of = [&](int64_t idx) { return (int)((idx / Xi) % Yi); };
    auto Z_of = [&](int64_t idx) { return (int)(idx / (Xi * Yi)); };
    auto xAt = [&](int64_t idx) { return coordinateAt(xVar, X_of(idx)); };
    auto yAt = [&](int64_t idx) { return coordinateAt(yVar, Y_of(idx)); };
    auto zAt = [&](int64_t idx) { return zVar ? coordinateAt(zVar, Z_of(idx)) : 0.0; };
    // planes: 0 = xy (hor X, ver Y), 1 = yz (hor Y, ver Z), 2 = xz (hor X, ver Z)
    const int hAxis[3] = {0, 1, 0}, vAxis[3] = {1, 2, 2};
    // --- neighbours and d_ij (:1160-1232)
    for (size_t j = 0; j < M; ++j) {
      const int64_t index = g->memDomainIndex[j];
      std::map<int64_t, size_t>::const_iterator pit = pointOf.find(index);
      if (pit == pointOf.end()) continue;  // an entry that is no membrane point has an all-false bType: no plane
      const MemPoint& p = g->memPoints[pit->second];
      const int co[3] = {p.X, p.Y, p.Z};
      Geo::Voronoi& v = g->voronoi[index];
      std::vector<P3>& ct = contour[index];
      if (ct.empty()) ct.assign(6, P3{0.0, 0.0, 0.0});
      for (int pl = 0; pl < (dimension == 3 ? 3 : 1); ++pl) {
        if (!crosses(p, pl)) continue;
        PlaneWalker w;
        w.g = g; w.total = total; w.sh = stride[hAxis[pl]]; w.sv = stride[vAxis[pl]];
        w.ends(index, co[hAxis[pl]], co[vAxis[pl]], 1, planeH[pl], planeV[pl]);
        for (int l = 0; l < 2; ++l) {
          int q[3] = {co[0], co[1], co[2]};
          q[hAxis[pl]] = planeH[pl][l];
          q[vAxis[pl]] = planeV[pl][l];
          const int64_t aidx = (int64_t)q[2] * Xi * Yi + (int64_t)q[1] * Xi + q[0];
          v.adj[2 * pl + l] = aidx;
          double inner;
          if (dimension == 2) inner = p.nx * (xAt(aidx) - xAt(index)) + p.ny * (yAt(aidx) - yAt(index));
          else inner = p.nx * (xAt(aidx) - xAt(index)) + p.ny * (yAt(aidx) - yAt(index)) + p.nz * (zAt(aidx) - zAt(index));
          P3& c = ct[2 * pl + l];
          c.x = xAt(aidx) - p.nx * inner;
          c.y = yAt(aidx) - p.ny * inner;
          if (dimension == 2) {
            v.di[2 * pl + l] = sqrt(pow(xAt(index) - c.x, 2) + pow(yAt(index) - c.y, 2));
          } else {
            c.z = zAt(aidx) - p.nz * inner;
            v.di[2 * pl + l] = sqrt(pow(xAt(index) - c.x, 2) + pow(yAt(index) - c.y, 2) + pow(zAt(index) - c.z, 2));
          }
        }
      }
    }
    // --- s_ij: the tangent plane is rotated into z = const and the perpendicular bisectors of neighbouring planes are
    // intersected (3-D only, :1234-1430)
    if (dimension == 3) {
      for (size_t j = 0; j < M; ++j) {
        const int64_t index = g->memDomainIndex[j];
        std::map<int64_t, size_t>::const_iterator pit = pointOf.find(index);
        if (pit == pointOf.end()) continue;
        const MemPoint& p = g->memPoints[pit->second];
        Geo::Voronoi& v = g->voronoi[index];
        const std::vector<P3>& ct = contour[index];
        const double phi = atan2(p.ny, p.nx), theta = acos(p.nz);
        const double rix = cos(theta) * (xAt(index) * cos(phi) + yAt(index) * sin(phi)) - zAt(index) * sin(theta);
        const double riy = -xAt(index) * sin(phi) + yAt(index) * cos(phi);
        double rx[6], ry[6];
        for (int q = 0; q < 6; ++q) {
          rx[q] = cos(theta) * (ct[q].x * cos(phi) + ct[q].y * sin(phi)) - ct[q].z * sin(theta);
          ry[q] = -ct[q].x * sin(phi) + ct[q].y * cos(phi);
        }
        // facet lengths of plane A against the two neighbours of plane B
        auto facets = [&](int A, int B) {
          for (int l = 0; l < 2; ++l) {
            double Px[4], Py[4], cpx[2], cpy[2];
            Px[0] = (rx[2 * A + l] + rix) / 2.0;
            Py[0] = (ry[2 * A + l] + riy) / 2.0;
            Px[2] = -(ry[2 * A + l] - riy) + Px[0];
            Py[2] = (rx[2 * A + l] - rix) + Py[0];
            for (int k = 0; k < 2; ++k) {
              Px[1] = (rx[2 * B + k] + rix) / 2.0;
              Py[1] = (ry[2 * B + k] + riy) / 2.0;
              Px[3] = -(ry[2 * B + k] - riy) + Px[1];
              Py[3] = (rx[2 * B + k] - rix) + Py[1];
              const double a013 = ((Px[3] - Px[1]) * (Py[0] - Py[1]) - (Py[3] - Py[1]) * (Px[0] - Px[1])) / 2.0;
              const double a123 = ((Px[3] - Px[1]) * (Py[1] - Py[2]) - (Py[3] - Py[1]) * (Px[1] - Px[2])) / 2.0;
              cpx[k] = Px[0] + ((Px[2] - Px[0]) * a013) / (a013 + a123);
              cpy[k] = Py[0] + ((Py[2] - Py[0]) * a013) / (a013 + a123);
            }
            v.si[2 * A + l] = sqrt(pow(cpx[0] - cpx[1], 2) + pow(cpy[0] - cpy[1], 2));
          }
        };
        const bool pxy = crosses(p, 0), pyz = crosses(p, 1), pxz = crosses(p, 2);
        if (pxy && pyz) { facets(0, 1); facets(1, 0); }
        if (pxy && pxz) { facets(0, 2); facets(2, 0); }
        if (pyz && pxz) { facets(1, 2); facets(2, 1); }
      }
    }
    // --- d_ij and s_ij are averaged with the neighbour's view of the same edge (:1476-1540).  When the neighbour does not list
    // this point (k runs to 2) the reference reuses the d_ji / s_ji of the previous edge and writes one slot past the
    // neighbour's pair, which in voronoiInfo's layout is the first slot of the next plane; both are mirrored here for the
    // xy and yz planes (the xz overflow leaves the struct).
    double d_ji = 0.0, s_ji = 0.0;
    for (size_t j = 0; j < M; ++j) {
      const int64_t index = g->memDomainIndex[j];
      std::map<int64_t, size_t>::const_iterator pit = pointOf.find(index);
      if (pit == pointOf.end()) continue;
      const MemPoint& p = g->memPoints[pit->second];
      for (int pl = 0; pl < (dimension == 3 ? 3 : 1); ++pl) {
        if (!crosses(p, pl)) continue;
        for (int l = 0; l < 2; ++l) {
          Geo::Voronoi& v = g->voronoi[index];
          if (v.isAve[2 * pl + l]) continue;
          const int64_t aidx = v.adj[2 * pl + l];
          if (aidx < 0) continue;
          const double d_ij = v.di[2 * pl + l], s_ij = v.si[2 * pl + l];
          Geo::Voronoi& o = g->voronoi[aidx];
          int k = 0;
          for (; k < 2; ++k)
            if (index == o.adj[2 * pl + k]) { d_ji = o.di[2 * pl + k]; s_ji = o.si[2 * pl + k]; break; }
          Geo::Voronoi& v2 = g->voronoi[index];  // (operator[] above may have rehashed)
          v2.di[2 * pl + l] = (d_ij + d_ji) / 2.0;
          v2.si[2 * pl + l] = (s_ij + s_ji) / 2.0;
          v2.isAve[2 * pl + l] = true;
          const int slot = 2 * pl + k;
          if (slot < 6) {
            Geo::Voronoi& o2 = g->voronoi[aidx];
            o2.di[slot] = v2.di[2 * pl + l];
            o2.si[slot] = v2.si[2 * pl + l];
            o2.isAve[slot] = true;
          }
        }
      }
    }
  }
}

// Pseudo-membrane fill: which two membrane points a pseudo-membrane point takes the minimum of.  The same cascade of tests
// appears in setBoundaryType (boundaryFunction.cpp:98-111, 173-203: every point classified 2) and in updateAssignmentRules
// (spatialsimulator.cpp:1614-1663: the points of pseudoMemIndex, in list order).  Lists are in storage indices of the geometry.
void pseudoFillPairs(const Model& m, const Geo* g, bool everyClassifiedPoint, std::vector<int32_t>& target, std::vector<int32_t>& a,
                     std::vector<int32_t>& b) {
  target.clear(); a.clear(); b.clear();
  std::vector<int64_t> pts;
  if (everyClassifiedPoint) {
    for (std::unordered_map<int64_t, uint8_t>::const_iterator it = g->memClass.begin(); it != g->memClass.end(); ++it)
      if (it->second == 2) pts.push_back(it->first);
    std::sort(pts.begin(), pts.end());
  } else {
    pts = g->pseudoMemIndex;
  }
  const int64_t Xi = m.Xindex, Yi = m.Yindex, total = m.ndoubled();
  for (size_t q = 0; q < pts.size(); ++q) {
    const int64_t idx = pts[q];
    // flat-index arithmetic as in the reference (no bounds tests); points outside the array read as "not a membrane point"
    const int64_t n[6] = {idx + 1, idx - 1, idx + Xi, idx - Xi, idx + Xi * Yi, idx - Xi * Yi};  // Xp Xm Yp Ym Zp Zm
    bool is1[6];
    for (int k = 0; k < 6; ++k) is1[k] = n[k] >= 0 && n[k] < total && g->memIsDomain(n[k]) == 1;
    if (m.dimension != 3) is1[4] = is1[5] = false;
    static const int order[15][2] = {{0, 1}, {2, 3}, {4, 5}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {2, 4}, {2, 5}, {3, 4}, {3, 5}};
    for (int c = 0; c < 15; ++c) {
      if (!(is1[order[c][0]] && is1[order[c][1]])) continue;
      const int t = g->storageOf(idx), sa = g->storageOf(n[order[c][0]]), sb = g->storageOf(n[order[c][1]]);
      if (t >= 0 && sa >= 0 && sb >= 0) { target.push_back(t); a.push_back(sa); b.push_back(sb); }
      break;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// boundary types of the volumes that host spatial species (setBoundaryType, boundaryFunction.cpp:13-217)
// ---------------------------------------------------------------------------------------------
void Model::setBoundaryTypes() {
  const int64_t nc = ncells();
  const int64_t sy = nx, sz = (int64_t)nx * ny;
  for (unsigned int i = 0; i < model->getNumSpecies(); ++i) {
    Species* s = model->getSpecies(i);
    Var* v = findVar(s->getId());
    Geo* g = geoByCompartment(s->getCompartment());
    if (!v || v->isUniform || !g) continue;
    v->geo = g;
    if (!g->hasMask) continue;
    if (!g->isVol) {
      // membrane species (boundaryFunction.cpp:89-118, 170-210): values off the membrane are zero (they have no storage
      // here), pseudo-membrane points take the smaller of two neighbouring membrane points
      if (!v->onMembrane || !v->hasValue) continue;
      std::vector<int32_t> t, a, b;
      pseudoFillPairs(*this, g, true, t, a, b);
      for (size_t q = 0; q < t.size(); ++q) v->field[t[q]] = std::min(v->field[a[q]], v->field[b[q]]);
      continue;
    }
    if (v->hasValue) {
      for (int64_t c = 0; c < nc; ++c)
        if (g->isDomain[c] == 0) v->field[c] = 0.0;
    }
    if (dimension < 1 || dimension > 3) continue;
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j)
        for (int ii = 0; ii < nx; ++ii) {
          const int64_t c = (int64_t)k * sz + (int64_t)j * sy + ii;
          if (g->isDomain[c] != 1 || g->isBoundary[c] != 1) continue;
          uint8_t f = g->flags[c];
          const int gj = j + off[1], gk = k + off[2];
          const bool frame = ii == 0 || ii == nx - 1 || (dimension >= 2 && (gj == 0 || gj == gn[1] - 1)) ||
                             (dimension == 3 && (gk == 0 || gk == gn[2] - 1));
          if (frame) {
            if (ii == 0) f |= SB2_F_XM;
            if (ii == nx - 1) f |= SB2_F_XP;
            if (dimension >= 2) {
              if (gj == 0) f |= SB2_F_YM;
              if (gj == gn[1] - 1) f |= SB2_F_YP;
            }
            if (dimension == 3) {
              if (gk == 0) f |= SB2_F_ZM;
              if (gk == gn[2] - 1) f |= SB2_F_ZP;
            }
          } else if (cutEdge(ii, j, k)) {
            // halo row / plane of a slab: its flags are computed by the slab that owns it
          } else {
            if (g->isDomain[c - 1] == 0) f |= SB2_F_XM;
            if (g->isDomain[c + 1] == 0) f |= SB2_F_XP;
            if (dimension >= 2) {
              if (g->isDomain[c - sy] == 0) f |= SB2_F_YM;
              if (g->isDomain[c + sy] == 0) f |= SB2_F_YP;
            }
            if (dimension == 3) {
              if (g->isDomain[c - sz] == 0) f |= SB2_F_ZM;
              if (g->isDomain[c + sz] == 0) f |= SB2_F_ZP;
            }
          }
          g->flags[c] = f;
        }
  }
  for (size_t gi = 0; gi < geos.size(); ++gi) {
    Geo* g = geos[gi];
    if (!g->isVol || !g->hasMask) continue;
    for (int64_t c = 0; c < nc; ++c) {
      g->flags[c] &= ~SB2_F_DOMAIN;
      if (g->isDomain[c] == 1) g->flags[c] |= SB2_F_DOMAIN;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// reactions and rate rules (setReactionInfo / setRateRuleInfo, setInfoFunction.cpp:352-494)
// ---------------------------------------------------------------------------------------------
void Model::setReactions() {
  for (unsigned int i = 0; i < model->getNumReactions(); ++i) {
    Reaction* r = model->getReaction(i);
    KineticLaw* kl = r->getKineticLaw();
    if (!kl) continue;
    // local parameters join the global symbol table; an id that already exists is shadowed by the
    // earlier entry (first match wins in findVar)
    for (unsigned int j = 0; j < kl->getNumLocalParameters(); ++j) {
      const LocalParameter* lp = kl->getLocalParameter(j);
      Var* v = new Var;
      v->id = lp->getId();
      if (model->getRule(v->id) == NULL && lp->isSetValue()) {
        v->isResolved = true;
        v->isUniform = true;
        v->hasValue = true;
        v->uniformValue = lp->getValue();
      }
      vars.push_back(v);
    }
    Rxn* x = new Rxn;
    x->reaction = r;
    x->id = r->getId();
    x->isMemTransport = false;
    for (unsigned int j = 0; j < r->getNumReactants() && !x->isMemTransport; ++j) {
      Species* re = model->getSpecies(r->getReactant(j)->getSpecies());
      for (unsigned int k = 0; k < r->getNumProducts(); ++k) {
        Species* pr = model->getSpecies(r->getProduct(k)->getSpecies());
        if (re && pr && re->getCompartment() != pr->getCompartment()) { x->isMemTransport = true; break; }
      }
    }
    x->nReactants = (int)r->getNumReactants();
    for (int side = 0; side < 2; ++side) {
      const unsigned int n = side == 0 ? r->getNumReactants() : r->getNumProducts();
      for (unsigned int j = 0; j < n; ++j) {
        SpeciesReference* sr = side == 0 ? r->getReactant(j) : r->getProduct(j);
        x->spRefs.push_back(findVar(sr->getSpecies()));
        x->stoichiometry.push_back(sr->getStoichiometry());
        Species* s = model->getSpecies(sr->getSpecies());
        x->isVariable.push_back(s ? (!s->isSetConstant() || !s->getConstant() || !s->getBoundaryCondition()) : false);
      }
    }
    if (kl->getMath()) {
      ASTNode* ast = kl->getMath()->deepCopy();
      rearrangeAst(ast);
      emitRpn(ast, *this, x->rpn);
      delete ast;
    }
    if (!r->getFast()) rxns.push_back(x);
    else fastRxns.push_back(x);
  }
}

void Model::setRateRules() {
  for (unsigned int i = 0; i < model->getNumRules(); ++i) {
    Rule* rule = model->getRule(i);
    if (!rule->isRate()) continue;
    Rxn* x = new Rxn;
    x->reaction = NULL;
    x->id = rule->getVariable();
    x->isMemTransport = false;
    x->nReactants = 1;
    if (rule->getMath()) {
      ASTNode* ast = rule->getMath()->deepCopy();
      rearrangeAst(ast);
      emitRpn(ast, *this, x->rpn);
      delete ast;
    }
    x->spRefs.push_back(findVar(rule->getVariable()));
    x->stoichiometry.push_back(1.0);
    Species* s = model->getSpecies(rule->getVariable());
    x->isVariable.push_back(s ? (!s->isSetConstant() || !s->getConstant() || !s->getBoundaryCondition()) : false);
    rxns.push_back(x);
  }
}

// ---------------------------------------------------------------------------------------------
// initFromModel, spatialsimulator.cpp:275-1223
// ---------------------------------------------------------------------------------------------
bool Model::build(SBMLDocument* d, int xdim, int ydim, int zdim, int slabLo, int slabHi) {
  doc = d;
  if (!d || !d->getModel()) { error = "no SBML model"; return false; }
  model = d->getModel();
  SpatialModelPlugin* mp = static_cast<SpatialModelPlugin*>(model->getPlugin("spatial"));
  if (!mp || !mp->getGeometry()) { error = "the model has no spatial geometry"; return false; }
  dimension = (int)mp->getGeometry()->getNumCoordinateComponents();
  nx = xdim; ny = ydim; nz = zdim;
  if (dimension <= 1) { ny = 1; nz = 1; }
  if (dimension <= 2) nz = 1;
  if (nx < 1 || ny < 1 || nz < 1) { error = "grid dimensions must be positive"; return false; }
  gn[0] = nx; gn[1] = ny; gn[2] = nz;
  off[0] = off[1] = off[2] = 0;
  Xindex = 2 * nx - 1; Yindex = 2 * ny - 1; Zindex = 2 * nz - 1;  // doubled grid of the WHOLE grid
  slab = slabHi > slabLo;
  if (slab) {
    const int axis = dimension == 3 ? 2 : 1;
    if (dimension < 2 || slabLo < 0 || slabHi > gn[axis]) { error = "slab range outside the grid"; return false; }
    off[axis] = slabLo;
    if (axis == 1) ny = slabHi - slabLo; else nz = slabHi - slabLo;
  }
  volDimension = 0; memDimension = 0;
  for (unsigned int i = 0; i < model->getNumCompartments(); ++i) {
    Compartment* c = model->getCompartment(i);
    if (volDimension == 0) volDimension = c->getSpatialDimensions();
    else if (c->getSpatialDimensions() >= volDimension) volDimension = c->getSpatialDimensions();
    else memDimension = c->getSpatialDimensions();
  }
  const bool timing = getenv("SSB_TIMING") != NULL;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[ssb] %-22s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  };
  setCompartments();
  setSpecies();
  if (!setParameters()) return false;
  lap("parameters");
  tVar = new Var;
  tVar->id = "t";
  tVar->isResolved = true; tVar->isUniform = true; tVar->hasValue = true; tVar->uniformValue = 0.0;
  vars.push_back(tVar);
  if (!setVolumeGeometries()) return false;
  lap("volume geometries");
  setMembraneGeometries();
  allocateMembraneFields();
  lap("membrane geometries");
  if (!resolveAssignments()) return false;
  lap("assignments");
  setBoundaryTypes();
  lap("boundary types");
  setReactions();
  setRateRules();
  lap("reactions");
  // Membrane normals feed only membrane transport and membrane diffusion; computing them walks the
  // whole contour per point, so skip the pass when nothing consumes it.
  bool needNormals = false;
  for (size_t i = 0; i < rxns.size(); ++i) needNormals = needNormals || rxns[i]->isMemTransport;
  for (size_t i = 0; i < vars.size(); ++i)
    needNormals = needNormals || (vars[i]->sp && !vars[i]->isUniform && vars[i]->geo && !vars[i]->geo->isVol);
  if (needNormals) {
    setNormals();
    setVoronoi();
  }
  return true;
}

}  // namespace sb2host
