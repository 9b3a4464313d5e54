// spatialsimulator.cpp — SpatialSimulator over the B200 device layer.
//
// Public behaviour follows the reference class (SpatialSBML/spatialsimulator.cpp:152-260 and
// :1236-1783); set-up lives in host_model.cpp; the time step is sb2_step() (include/spatial_b200.h).
#include "spatialsimulator.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

#include "../../../include/spatial_b200.h"
#include "host_model.h"

using sb2host::Geo;
using sb2host::MemPoint;
using sb2host::Rpn;
using sb2host::Rxn;
using sb2host::Tok;
using sb2host::Var;

namespace {

// rows / planes of the neighbouring slabs held locally: two are exchanged by models with advection / membrane transport, four
// let a slab step from host memory with a single exchange of edge rows (sb2_step_host_begin)
const int kSlabHalo = 4;

struct VarMirror {
  std::vector<double> data;  // doubled grid
  bool handedOut;
  VarMirror() : handedOut(false) {}
};

int opOfAst(int t) {
  switch (t) {
    case AST_PLUS: return SB2_OP_PLUS;
    case AST_MINUS: return SB2_OP_MINUS;
    case AST_TIMES: return SB2_OP_TIMES;
    case AST_DIVIDE: return SB2_OP_DIVIDE;
    case AST_POWER:
    case AST_FUNCTION_POWER: return SB2_OP_POWER;
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
    default: return SB2_OP_POP_ONLY;  // listed-but-empty cases of the reference interpreter
  }
}

}  // namespace

struct SimImpl {
  sb2host::Model m;
  SBMLDocument* ownedDoc;
  sb2_sim* dev;
  std::vector<double> consts;
  std::map<uint64_t, int> literalSlot;
  int timeSlot;
  std::vector<Var*> staged;     // device species id → variable
  std::vector<Var*> fieldVars;  // device field id → variable
  std::vector<char> stale;      // per staged species: device state is newer than Var::field
  std::vector<SpatialSimulator*> slabs;  // group mode (setDevices): one simulator per slab; this object keeps the whole grid's host model
  std::vector<std::pair<int, int> > slabRows;  // owned rows [lo, hi) of each slab
  std::set<const Var*> groupFresh;       // group mode: variables whose host copy is as new as the slabs' device state
  std::vector<Var*> ruleTargets;  // variables an assignment rule rewrites on the device after every step
  bool rulesStale;                // ... whose host copies are older than the device's
  std::vector<Var*> memStaged;  // device membrane species id → variable
  std::vector<char> memStale;
  std::map<std::string, VarMirror> varMirrors;
  std::map<std::string, std::vector<int> > geoMirrors, boundaryMirrors;
  std::map<std::string, std::vector<boundaryType> > btMirrors;
  int ownLo, ownHi, heldLo;  // slab rows (global) owned / first row held locally
  long domainCells;
  // slab with membrane transport: host model of the WHOLE grid (membrane points, their normals and the frame tests of the
  // reference refer to the whole doubled grid); built on first use, never stepped
  sb2host::Model* whole;
  SimImpl() : ownedDoc(NULL), dev(NULL), timeSlot(-1), rulesStale(false), ownLo(0), ownHi(0), heldLo(0), domainCells(0), whole(NULL) {}
  sb2host::Model* wholeModel(std::string& err) {
    if (!m.isSlab()) return &m;
    if (!whole) {
      whole = new sb2host::Model;
      if (!whole->build(m.doc, m.gn[0], m.gn[1], m.gn[2], 0, 0)) { err = whole->error; delete whole; whole = NULL; }
    }
    return whole;
  }
  ~SimImpl() {
    delete whole;
    for (size_t i = 0; i < slabs.size(); ++i) delete slabs[i];
    if (dev) sb2_destroy(dev);
    delete ownedDoc;
  }

  int constSlotOf(Var* v) {
    if (v->constSlot < 0) {
      v->constSlot = (int)consts.size();
      consts.push_back(v->uniformValue);
    }
    return v->constSlot;
  }
  int literalSlotOf(double x) {
    uint64_t bits;
    memcpy(&bits, &x, sizeof(bits));
    std::map<uint64_t, int>::const_iterator it = literalSlot.find(bits);
    if (it != literalSlot.end()) return it->second;
    const int s = (int)consts.size();
    consts.push_back(x);
    literalSlot[bits] = s;
    return s;
  }
  // doubled-grid coordinate value (setInfoFunction.cpp:308): min + C*delta/2.0
  double coordinateAt(const Var* v, int C) const {
    const double delta = v == m.xVar ? m.deltaX : v == m.yVar ? m.deltaY : m.deltaZ;
    return v->offGridInit + (double)C * delta / 2.0;
  }
  bool isCoordinate(const Var* v) const { return v && (v == m.xVar || v == m.yVar || v == m.zVar); }
  void markStale() {
    for (size_t s = 0; s < stale.size(); ++s) stale[s] = 1;
    for (size_t s = 0; s < memStale.size(); ++s) memStale[s] = 1;
    rulesStale = !ruleTargets.empty();
  }
};

// =============================================================================================
// construction / initialisation
// =============================================================================================
SpatialSimulator::SpatialSimulator()
    : Xdiv(100), Ydiv(100), Zdiv(1), impl(NULL), deviceOrdinal(0), externalStream(NULL), hostOnly(false), slabLo(0), slabHi(0) {}

SpatialSimulator::SpatialSimulator(SBMLDocument* doc, int xdim, int ydim, int zdim)
    : Xdiv(xdim), Ydiv(ydim), Zdiv(zdim), impl(NULL), deviceOrdinal(0), externalStream(NULL), hostOnly(false), slabLo(0), slabHi(0) {
  initFromModel(doc, xdim, ydim, zdim);
}

SpatialSimulator::~SpatialSimulator() { clear(); }

void SpatialSimulator::clear() {
  delete impl;
  impl = NULL;
}

void SpatialSimulator::initFromFile(const std::string& fileName, int xdim, int ydim, int zdim) {
  SBMLDocument* doc = readSBMLFromFile(fileName.c_str());
  if (!doc || !doc->getModel()) {
    clear();
    lastError = "cannot read SBML file '" + fileName + "'";
    delete doc;
    return;
  }
  initFromModel(doc, xdim, ydim, zdim);
  if (impl) impl->ownedDoc = doc; else delete doc;
}

void SpatialSimulator::initFromString(const std::string& sbml, int xdim, int ydim, int zdim) {
  SBMLDocument* doc = readSBMLFromString(sbml.c_str());
  if (!doc || !doc->getModel()) {
    clear();
    lastError = "cannot parse the SBML string";
    delete doc;
    return;
  }
  initFromModel(doc, xdim, ydim, zdim);
  if (impl) impl->ownedDoc = doc; else delete doc;
}

void SpatialSimulator::initFromModel(SBMLDocument* doc, int xdim, int ydim, int zdim) {
  SBMLDocument* keep = NULL;
  if (impl && impl->ownedDoc == doc) { keep = impl->ownedDoc; impl->ownedDoc = NULL; }  // re-init on our own document
  clear();
  lastError.clear();
  impl = new SimImpl;
  impl->ownedDoc = keep;
  int heldLo = 0, heldHi = 0;
  if (slabHi > slabLo) {
    // which rows are held: the owned ones plus kSlabHalo rows of each neighbouring slab
    SpatialModelPlugin* mp = doc && doc->getModel() ? static_cast<SpatialModelPlugin*>(doc->getModel()->getPlugin("spatial")) : NULL;
    const int dim = mp && mp->getGeometry() ? (int)mp->getGeometry()->getNumCoordinateComponents() : 0;
    const int n = dim == 3 ? zdim : ydim;
    heldLo = slabLo - kSlabHalo < 0 ? 0 : slabLo - kSlabHalo;
    heldHi = slabHi + kSlabHalo > n ? n : slabHi + kSlabHalo;
    if (slabLo == 0 && slabHi == n) { heldLo = 0; heldHi = 0; }
  }
  if (!impl->m.build(doc, xdim, ydim, zdim, heldLo, heldHi)) {
    lastError = impl->m.error;
    SBMLDocument* d = impl->ownedDoc;
    impl->ownedDoc = NULL;
    clear();
    delete d;
    return;
  }
  impl->heldLo = impl->m.isSlab() ? heldLo : 0;
  impl->ownLo = impl->m.isSlab() ? slabLo : 0;
  impl->ownHi = impl->m.isSlab() ? slabHi : 0;
  Xdiv = impl->m.gn[0];
  Ydiv = impl->m.gn[1];
  Zdiv = impl->m.gn[2];
  if (!hostOnly && groupDevices.size() >= 2 && !impl->m.isSlab()) {
    if (!buildSlabs(doc, xdim, ydim, zdim)) {
      for (size_t i = 0; i < impl->slabs.size(); ++i) delete impl->slabs[i];
      impl->slabs.clear();
    }
    return;
  }
  if (!hostOnly && !buildDevice()) {
    // keep the host model for inspection, but there is no device → oneStep() will throw
    if (impl->dev) { sb2_destroy(impl->dev); impl->dev = NULL; }
  }
}

// group mode: one simulator per slab, each on its device; this object keeps the host model of the whole grid
bool SpatialSimulator::buildSlabs(SBMLDocument* doc, int xdim, int ydim, int zdim) {
  sb2host::Model& m = impl->m;
  const int n = m.dimension == 3 ? m.gn[2] : m.gn[1];
  const int G = (int)groupDevices.size();
  if (m.dimension != 2 && m.dimension != 3) { lastError = "slabs need a 2-D or 3-D grid"; return false; }
  if (G > n) { lastError = "more devices than rows along the slab axis"; return false; }
  for (int i = 0; i < G; ++i) {
    const int lo = (int)((int64_t)i * n / G), hi = (int)((int64_t)(i + 1) * n / G);
    SpatialSimulator* s = new SpatialSimulator();
    impl->slabs.push_back(s);
    impl->slabRows.push_back(std::make_pair(lo, hi));
    s->setDevice(groupDevices[i]);
    s->setSlab(lo, hi);
    s->initFromModel(doc, xdim, ydim, zdim);
    if (!s->isInitialized()) {
      lastError = "slab " + std::to_string(i) + " (rows " + std::to_string(lo) + ".." + std::to_string(hi) + ", device " +
                  std::to_string(groupDevices[i]) + "): " + s->getLastError();
      return false;
    }
  }
  impl->domainCells = 0;
  for (int i = 0; i < G; ++i) impl->domainCells += impl->slabs[i]->getNumDomainCells();
  // slabs of eight rows or more: one exchange of four rows of u per step, apron rows recomputed (sb2_sharded_set_apron; models off
  // the fused schedule keep the exchange per stage)
  int minRows = n;
  for (int i = 0; i < G; ++i) minRows = std::min(minRows, impl->slabRows[i].second - impl->slabRows[i].first);
  // (2-D grids only: in 3-D the apron is twelve whole planes per slab and step, measured slower than the exchanges it saves)
  if (minRows >= 8 && !getenv("SB2_NO_APRON") && (impl->m.dimension == 2 || getenv("SB2_APRON")))
    for (int i = 0; i < G; ++i) sb2_sharded_set_apron(impl->slabs[i]->impl->dev, 1);
  return true;
}

// group mode: the owned rows of every slab, put together in a compact array of the whole grid
bool SpatialSimulator::gatherCompact(const std::string& id, double* out) {
  sb2host::Model& m = impl->m;
  const int64_t rowCells = m.dimension == 3 ? (int64_t)m.nx * m.ny : (int64_t)m.nx;
  std::vector<double> tmp;
  for (size_t i = 0; i < impl->slabs.size(); ++i) {
    SpatialSimulator* s = impl->slabs[i];
    const int lo = impl->slabRows[i].first, hi = impl->slabRows[i].second;
    const int heldLo = s->impl->heldLo;
    const size_t nloc = (size_t)s->impl->m.ncells();
    tmp.resize(nloc);
    if (!s->getVariableCompact(id, tmp.data(), nloc)) return false;
    memcpy(out + (int64_t)lo * rowCells, tmp.data() + (int64_t)(lo - heldLo) * rowCells, sizeof(double) * (size_t)((hi - lo) * rowCells));
  }
  return true;
}

bool SpatialSimulator::isInitialized() const { return impl != NULL && (hostOnly || impl->dev != NULL || !impl->slabs.empty()); }
void SpatialSimulator::setDevices(const int* ordinals, int n) {
  groupDevices.clear();
  for (int i = 0; i < n && ordinals; ++i) groupDevices.push_back(ordinals[i]);
  if (n == 1) deviceOrdinal = ordinals[0];
}
int SpatialSimulator::getNumSlabs() const { return impl ? (int)impl->slabs.size() : 0; }
void SpatialSimulator::getSlabRows(int& heldLo, int& ownLo, int& ownHi) const {
  heldLo = impl ? impl->heldLo : 0; ownLo = impl ? impl->ownLo : 0; ownHi = impl ? impl->ownHi : 0;
}
const std::string& SpatialSimulator::getLastError() const { return lastError; }
void SpatialSimulator::setDevice(int ordinal) { deviceOrdinal = ordinal; }
void SpatialSimulator::setHostOnly(bool h) { hostOnly = h; }
void SpatialSimulator::setSlab(int lo, int hi) { slabLo = lo; slabHi = hi; }
void SpatialSimulator::setCudaStream(void* stream) {
  externalStream = stream;
  if (impl && impl->dev) sb2_set_stream(impl->dev, stream);
}
void* SpatialSimulator::getDeviceHandle() { return impl ? impl->dev : NULL; }
const Model* SpatialSimulator::getModel() const { return impl ? impl->m.model : NULL; }
int SpatialSimulator::runOldMain(int, const char*[]) { return -1; }
void SpatialSimulator::setGnuplotExecutable(const std::string& location) { gnuplotExecutable = location; }
const std::string& SpatialSimulator::getGnuplotExecutable() const { return gnuplotExecutable; }

// =============================================================================================
// host model → device program (C ABI calls only)
// =============================================================================================
#define DEV(call)                                                                                     \
  do {                                                                                                \
    int rc__ = (call);                                                                                \
    if (rc__ < 0) {                                                                                   \
      lastError = std::string(#call) + ": " + sb2_last_error(impl->dev);                              \
      return false;                                                                                   \
    }                                                                                                 \
  } while (0)

bool SpatialSimulator::buildDevice() {
  sb2host::Model& m = impl->m;
  if (m.dimension != 2 && m.dimension != 3) {
    lastError = "the device step supports 2-D and 3-D geometries (this model has " + std::to_string(m.dimension) + " coordinate components)";
    return false;
  }
  sb2_grid g;
  memset(&g, 0, sizeof(g));
  g.nx = m.nx; g.ny = m.ny; g.nz = m.nz; g.dim = m.dimension;
  g.dx = m.deltaX; g.dy = m.deltaY; g.dz = m.deltaZ;
  g.device = deviceOrdinal;
  if (m.isSlab()) {
    g.own_lo = impl->ownLo - impl->heldLo;
    g.own_hi = impl->ownHi - impl->heldLo;
    g.row0 = impl->heldLo;
  }
  int rc = sb2_create(&g, &impl->dev);
  if (rc < 0) {
    lastError = std::string("sb2_create: ") + sb2_last_error(NULL);
    impl->dev = NULL;
    return false;
  }
  if (externalStream) sb2_set_stream(impl->dev, externalStream);

  // --- staged species and the geometries that host them -----------------------------------------
  auto deviceGeoOf = [&](Geo* geo) -> int {
    if (geo->deviceGeo >= 0) return geo->deviceGeo;
    geo->deviceGeo = sb2_add_geometry(impl->dev, geo->flags.data());
    return geo->deviceGeo;
  };
  for (unsigned int i = 0; i < m.model->getNumSpecies(); ++i) {
    Var* v = m.findVar(m.model->getSpecies(i)->getId());
    if (!v || v->sp != m.model->getSpecies(i) || v->isUniform || !v->hasValue) continue;
    if (!v->geo) continue;  // no geometry for its compartment: the reference never touches it
    if (!v->geo->isVol) continue;  // species on a membrane: staged below (addMembraneSpecies)
    if (!v->hasDelta) continue;  // constant spatial species: a read-only field
    if (!v->geo->hasMask) continue;
    const int dg = deviceGeoOf(v->geo);
    if (dg < 0) { lastError = std::string("sb2_add_geometry: ") + sb2_last_error(impl->dev); return false; }
    const int id = sb2_add_species(impl->dev, dg, v->field.data());
    if (id < 0) { lastError = std::string("sb2_add_species: ") + sb2_last_error(impl->dev); return false; }
    v->speciesId = id;
    impl->staged.push_back(v);
    impl->stale.push_back(0);
  }
  impl->domainCells = 0;
  {
    std::vector<char> counted(m.geos.size(), 0);
    for (size_t s = 0; s < impl->staged.size(); ++s)
      for (size_t gi = 0; gi < m.geos.size(); ++gi)
        if (m.geos[gi] == impl->staged[s]->geo && !counted[gi]) {
          counted[gi] = 1;
          if (!m.isSlab()) impl->domainCells += (long)m.geos[gi]->nDomain;
          else {
            // owned rows only
            const int axis = m.dimension == 3 ? 2 : 1;
            const int64_t rowCells = axis == 1 ? m.nx : (int64_t)m.nx * m.ny;
            const int lo = impl->ownLo - impl->heldLo, hi = impl->ownHi - impl->heldLo;
            for (int64_t c = (int64_t)lo * rowCells; c < (int64_t)hi * rowCells; ++c) impl->domainCells += m.geos[gi]->isDomain[c] == 1;
          }
        }
  }

  if (!addMembraneSpecies()) return false;

  auto fieldIdOf = [&](Var* v) -> int {
    if (v->fieldId >= 0) return v->fieldId;
    v->fieldId = sb2_add_field(impl->dev, v->field.data());
    if (v->fieldId >= 0) impl->fieldVars.push_back(v);
    return v->fieldId;
  };
  // source descriptor of a coefficient variable (uniform → constants table, field → device field)
  auto sourceOf = [&](Var* c, int& kind, int& src) -> bool {
    kind = SB2_SRC_NONE; src = 0;
    if (!c || !c->hasValue) return true;
    if (c->isUniform) { kind = SB2_SRC_CONST; src = impl->constSlotOf(c); return true; }
    if (c->speciesId >= 0) { lastError = "coefficient '" + c->id + "' is an integrated species (unsupported)"; return false; }
    kind = SB2_SRC_FIELD;
    src = fieldIdOf(c);
    if (src < 0) { lastError = std::string("sb2_add_field: ") + sb2_last_error(impl->dev); return false; }
    return true;
  };

  // --- transport terms (performStep, spatialsimulator.cpp:1464-1491) -----------------------------
  for (size_t s = 0; s < impl->staged.size(); ++s) {
    Var* v = impl->staged[s];
    if (!v->isResolved) continue;
    int kind, src;
    if (v->hasDiff)
      for (int a = 0; a < 3; ++a) {
        if (!v->diffC[a]) continue;
        if (!sourceOf(v->diffC[a], kind, src)) return false;
        if (kind != SB2_SRC_NONE) DEV(sb2_set_diffusion(impl->dev, v->speciesId, a, kind, src));
      }
    if (v->hasBC)
      for (int k = 0; k < 6; ++k) {
        if (!v->bc[k] || v->bcKind[k] == SB2_BC_NONE) continue;
        if (!sourceOf(v->bc[k], kind, src)) return false;
        if (kind != SB2_SRC_NONE) DEV(sb2_set_boundary(impl->dev, v->speciesId, k, v->bcKind[k], kind, src));
      }
    if (v->hasAdv) {
      for (int a = 0; a < 3; ++a) {
        if (!v->advC[a]) continue;
        if (!sourceOf(v->advC[a], kind, src)) return false;
        if (kind != SB2_SRC_NONE) DEV(sb2_set_advection(impl->dev, v->speciesId, a, kind, src));
        if (kind == SB2_SRC_FIELD) {
          // the velocity between the cells: its value at the odd doubled-grid point on the plus side of every cell
          Var* c = v->advC[a];
          std::vector<double> fv((size_t)m.ncells(), 0.0);
          for (int kk = 0; kk < m.nz; ++kk)
            for (int jj = 0; jj < m.ny; ++jj)
              for (int ii = 0; ii < m.nx; ++ii) {
                const int X = 2 * (ii + m.off[0]) + (a == 0), Y = 2 * (jj + m.off[1]) + (a == 1), Z = 2 * (kk + m.off[2]) + (a == 2);
                if (X < m.Xindex && Y < m.Yindex && Z < m.Zindex) fv[((size_t)kk * m.ny + jj) * m.nx + ii] = m.valueAtDoubled(c, X, Y, Z);
              }
          const int ff = sb2_add_field(impl->dev, fv.data());
          if (ff < 0) { lastError = std::string("sb2_add_field: ") + sb2_last_error(impl->dev); return false; }
          impl->fieldVars.push_back(NULL);  // device field without a model variable of its own
          DEV(sb2_set_advection_faces(impl->dev, v->speciesId, a, ff));
        }
      }
      // CIP face point values = the odd doubled-grid points of the species array (calcPDE.cpp:588-603)
      std::vector<double> faces((size_t)m.ncells(), v->offGridInit);
      for (int a = 0; a < m.dimension; ++a) DEV(sb2_upload_faces(impl->dev, v->speciesId, a, faces.data()));
    }
  }

  // --- kinetic laws → token streams --------------------------------------------------------------
  auto toDeviceTokens = [&](const Rpn& rpn, std::vector<sb2_token>& out) -> bool {
    out.clear();
    for (size_t i = 0; i < rpn.size(); ++i) {
      const Tok& t = rpn[i];
      sb2_token d;
      d.kind = SB2_TOK_NOP; d.op = 0; d.slot = 0;
      switch (t.kind) {
        case Tok::OP: d.kind = SB2_TOK_OP; d.op = (uint8_t)opOfAst(t.astType); break;
        case Tok::LITERAL: d.kind = SB2_TOK_CONST; d.slot = (uint16_t)impl->literalSlotOf(t.literal); break;
        case Tok::CONSTVAR: d.kind = SB2_TOK_CONST; d.slot = (uint16_t)impl->constSlotOf(t.var); break;
        case Tok::VAR:
          if (t.var->speciesId >= 0) { d.kind = SB2_TOK_VAR; d.slot = (uint16_t)t.var->speciesId; }
          else {
            const int f = fieldIdOf(t.var);
            if (f < 0) { lastError = std::string("sb2_add_field: ") + sb2_last_error(impl->dev); return false; }
            d.kind = SB2_TOK_FIELD; d.slot = (uint16_t)f;
          }
          break;
        default: break;
      }
      out.push_back(d);
    }
    return true;
  };
  auto toDeviceRefs = [&](const Rxn* x, std::vector<sb2_ref>& out) {
    out.clear();
    for (size_t k = 0; k < x->spRefs.size(); ++k) {
      sb2_ref r;
      Var* v = x->spRefs[k];
      r.species = (v && !v->isUniform && v->speciesId >= 0) ? v->speciesId : -1;
      r.is_variable = x->isVariable[k] ? 1 : 0;
      r.stoichiometry = x->stoichiometry[k];
      out.push_back(r);
    }
  };

  std::vector<sb2_token> toks;
  std::vector<sb2_ref> refs;
  for (size_t i = 0; i < m.rxns.size(); ++i) {
    Rxn* x = m.rxns[i];
    if (!x->reaction) continue;  // rate-rule entries are dispatched below (spatialsimulator.cpp:1376-1377)
    if (!x->isMemTransport) {
      Geo* geo = NULL;
      for (size_t k = 0; k < x->spRefs.size(); ++k)
        if (x->spRefs[k] && x->spRefs[k]->geo) { geo = x->spRefs[k]->geo; break; }
      if (!geo) continue;  // every participant is uniform: the reference changes nothing
      if (!geo->isVol) {
        if (!addMembraneReaction(x, geo, 0)) return false;
        continue;
      }
      if (!geo->hasMask) { lastError = "reaction '" + x->id + "': its compartment has no geometry"; return false; }
      const int dg = deviceGeoOf(geo);
      if (dg < 0) { lastError = std::string("sb2_add_geometry: ") + sb2_last_error(impl->dev); return false; }
      if (!toDeviceTokens(x->rpn, toks)) return false;
      toDeviceRefs(x, refs);
      DEV(sb2_add_reaction(impl->dev, dg, 0, toks.data(), (int)toks.size(), refs.data(), (int)refs.size(), x->nReactants));
    } else {
      if (!addMemTransport(x)) return false;
    }
  }
  // rate rules: the reference evaluates rInfoList[<rule index>] on the rule variable's domain
  // (spatialsimulator.cpp:1455-1461) — with no reactions in the model that is the rule itself
  for (unsigned int i = 0; i < m.model->getNumRules(); ++i) {
    Rule* rule = m.model->getRule(i);
    if (!rule->isRate()) continue;
    Var* target = m.findVar(rule->getVariable());
    if (i >= m.rxns.size()) { lastError = "rate rule '" + rule->getVariable() + "': the reference indexes its reaction list out of range here"; return false; }
    if (!target || !target->geo) continue;
    Rxn* x = m.rxns[i];
    if (!target->geo->isVol) {
      if (!addMembraneReaction(x, target->geo, 1)) return false;
      continue;
    }
    if (!target->geo->hasMask) { lastError = "rate rule '" + rule->getVariable() + "': its compartment has no geometry"; return false; }
    const int dg = deviceGeoOf(target->geo);
    if (dg < 0) { lastError = std::string("sb2_add_geometry: ") + sb2_last_error(impl->dev); return false; }
    if (!toDeviceTokens(x->rpn, toks)) return false;
    toDeviceRefs(x, refs);
    DEV(sb2_add_reaction(impl->dev, dg, 1, toks.data(), (int)toks.size(), refs.data(), (int)refs.size(), 1));
  }

  // assignment rules are re-evaluated after every step, in dependency order (updateAssignmentRules, spatialsimulator.cpp:1566-1581)
  for (size_t i = 0; i < m.orderedAssignmentRules.size(); ++i) {
    Var* v = m.orderedAssignmentRules[i];
    if (v->onMembrane || (v->geo && !v->geo->isVol && v->sp)) {
      lastError = "assignment rule on the membrane species '" + v->id + "' is not supported on the device yet";
      return false;
    }
    int kind, target, geo = -1;
    if (v->sp) {
      // a species: only the cells of its compartment are written (isAllArea == false)
      if (!v->geo || !v->geo->hasMask) continue;  // the reference would dereference a NULL geometry here
      geo = deviceGeoOf(v->geo);
      if (geo < 0) { lastError = std::string("sb2_add_geometry: ") + sb2_last_error(impl->dev); return false; }
    }
    if (v->speciesId >= 0) { kind = 1; target = v->speciesId; }
    else {
      kind = 0; target = fieldIdOf(v);
      if (target < 0) { lastError = std::string("sb2_add_field: ") + sb2_last_error(impl->dev); return false; }
    }
    if (!toDeviceTokens(v->rpn, toks)) return false;
    DEV(sb2_add_assignment_rule(impl->dev, kind, target, geo, toks.data(), (int)toks.size()));
    impl->ruleTargets.push_back(v);
  }

  if (m.tVar->constSlot >= 0) impl->timeSlot = m.tVar->constSlot;
  if (!impl->consts.empty()) DEV(sb2_set_constants(impl->dev, impl->consts.data(), (int)impl->consts.size()));
  DEV(sb2_finalize(impl->dev));
  return true;
}

namespace {
void pointCoords(const sb2host::Model& m, int64_t idx, int& X, int& Y, int& Z) {
  const int64_t Xi = m.Xindex, Yi = m.Yindex;
  Z = (int)(idx / (Xi * Yi));
  Y = (int)((idx - (int64_t)Z * Xi * Yi) / Xi);
  X = (int)(idx - (int64_t)Z * Xi * Yi - (int64_t)Y * Xi);
}
}  // namespace

// Species that live on a membrane: storage on a point set, update list, surface diffusion over the Voronoi neighbours
// (calcMemDiffusion, calcPDE.cpp:1484-1581), pseudo-membrane fill (spatialsimulator.cpp:1614-1663)
bool SpatialSimulator::addMembraneSpecies() {
  sb2host::Model& m = impl->m;
  for (unsigned int i = 0; i < m.model->getNumSpecies(); ++i) {
    Var* v = m.findVar(m.model->getSpecies(i)->getId());
    if (!v || v->sp != m.model->getSpecies(i) || v->isUniform || !v->hasValue || !v->geo || v->geo->isVol) continue;
    if (!v->onMembrane || !v->hasDelta) continue;  // a constant membrane species is only ever read (per-point operand values)
    if (m.isSlab()) { lastError = "species '" + v->id + "' lives on a membrane; membrane species are not supported on a sharded grid yet"; return false; }
    Geo* g = v->geo;
    if (g->storage.size() > (size_t)0x7fffffff) { lastError = "membrane '" + g->compartmentId + "' has too many points"; return false; }
    if (g->devicePointSet < 0) {
      g->devicePointSet = sb2_add_point_set(impl->dev, (int)g->storage.size());
      if (g->devicePointSet < 0) { lastError = std::string("sb2_add_point_set: ") + sb2_last_error(impl->dev); return false; }
      std::vector<int32_t> t, a, b;
      sb2host::pseudoFillPairs(m, g, false, t, a, b);
      DEV(sb2_set_mem_fill(impl->dev, g->devicePointSet, t.data(), a.data(), b.data(), (int)t.size()));
    }
    const int id = sb2_add_mem_species(impl->dev, g->devicePointSet, v->field.data());
    if (id < 0) { lastError = std::string("sb2_add_mem_species: ") + sb2_last_error(impl->dev); return false; }
    v->memSpeciesId = id;
    impl->memStaged.push_back(v);
    impl->memStale.push_back(0);
    std::vector<int32_t> upd;
    for (size_t q = 0; q < g->memDomainIndex.size(); ++q) {
      const int k = g->storageOf(g->memDomainIndex[q]);
      if (k >= 0) upd.push_back(k);
    }
    DEV(sb2_set_mem_update_points(impl->dev, id, upd.data(), (int)upd.size()));
    if (!v->hasDiff || !v->diffC[0] || !v->diffC[0]->hasValue) continue;  // calcMemDiffusion reads diffCInfo[0] only
    Var* dc = v->diffC[0];
    std::map<int64_t, const MemPoint*> pointOf;
    for (size_t k = 0; k < g->memPoints.size(); ++k) pointOf[g->memPoints[k].index] = &g->memPoints[k];
    std::vector<int32_t> pts, start(1, 0), adj;
    std::vector<double> sij, dij, area, dvals;
    for (size_t q = 0; q < g->memDomainIndex.size(); ++q) {
      const int64_t idx = g->memDomainIndex[q];
      const int k = g->storageOf(idx);
      if (k < 0) continue;
      std::unordered_map<int64_t, Geo::Voronoi>::const_iterator vit = g->voronoi.find(idx);
      Geo::Voronoi vor;
      if (vit != g->voronoi.end()) vor = vit->second;
      // area = sum_j (d_xy s_xy + d_yz s_yz + d_xz s_xz), halved in 2-D, quartered in 3-D (calcPDE.cpp:1520-1527)
      double ar = 0.0;
      for (int j = 0; j < 2; ++j) {
        ar += vor.di[0 + j] * vor.si[0 + j];
        ar += vor.di[2 + j] * vor.si[2 + j];
        ar += vor.di[4 + j] * vor.si[4 + j];
      }
      if (m.dimension == 2) ar /= 2.0;
      else if (m.dimension == 3) ar /= 4.0;
      pts.push_back(k);
      area.push_back(ar);
      if (!dc->isUniform) {
        int X, Y, Z;
        pointCoords(m, idx, X, Y, Z);
        dvals.push_back(m.valueAtDoubled(dc, X, Y, Z));
      }
      std::map<int64_t, const MemPoint*>::const_iterator pit = pointOf.find(idx);
      if (pit != pointOf.end()) {
        const uint8_t bof = pit->second->bof;
        const bool bx = (bof & SB2_F_XP) && (bof & SB2_F_XM), by = (bof & SB2_F_YP) && (bof & SB2_F_YM), bz = (bof & SB2_F_ZP) && (bof & SB2_F_ZM);
        const bool plane[3] = {bx || by, m.dimension == 3 && (by || bz), m.dimension == 3 && (bx || bz)};
        for (int pl = 0; pl < 3; ++pl) {
          if (!plane[pl]) continue;
          for (int j = 0; j < 2; ++j) {
            const int a = vor.adj[2 * pl + j] >= 0 ? g->storageOf(vor.adj[2 * pl + j]) : -1;
            if (a < 0) continue;  // no neighbour found on this side (the reference would read outside its array)
            adj.push_back(a); sij.push_back(vor.si[2 * pl + j]); dij.push_back(vor.di[2 * pl + j]);
          }
        }
      }
      start.push_back((int32_t)adj.size());
    }
    const int kind = dc->isUniform ? SB2_SRC_CONST : SB2_SRC_FIELD;
    const int src = dc->isUniform ? impl->constSlotOf(dc) : 0;
    DEV(sb2_set_mem_diffusion(impl->dev, id, kind, src, dvals.empty() ? NULL : dvals.data(), pts.data(), (int)pts.size(), start.data(),
                              adj.data(), sij.data(), dij.data(), area.data()));
  }
  return true;
}

// A reaction between species of one membrane (kind 0) or a rate rule on a membrane variable (kind 1): the reference runs
// reversePolishRK over every entry of the membrane's domainIndex (spatialsimulator.cpp:1381-1390, calcPDE.cpp:224-451)
bool SpatialSimulator::addMembraneReaction(void* rx, void* geo, int kind) {
  Rxn* x = static_cast<Rxn*>(rx);
  Geo* g = static_cast<Geo*>(geo);
  sb2host::Model& m = impl->m;
  if (m.isSlab()) { lastError = "reaction '" + x->id + "' takes place on a membrane; not supported on a sharded grid yet"; return false; }
  if (g->devicePointSet < 0) return true;  // no variable species lives on this membrane: nothing can change
  std::vector<int32_t> pts;
  std::vector<int64_t> idxs;
  for (size_t q = 0; q < g->memDomainIndex.size(); ++q) {
    const int k = g->storageOf(g->memDomainIndex[q]);
    if (k >= 0) { pts.push_back(k); idxs.push_back(g->memDomainIndex[q]); }
  }
  std::vector<sb2_token> toks;
  std::vector<double> values;
  for (size_t i = 0; i < x->rpn.size(); ++i) {
    const Tok& t = x->rpn[i];
    sb2_token d;
    d.kind = SB2_TOK_NOP; d.op = 0; d.slot = 0;
    if (t.kind == Tok::OP) { d.kind = SB2_TOK_OP; d.op = (uint8_t)opOfAst(t.astType); }
    else if (t.kind == Tok::LITERAL) { d.kind = SB2_TOK_CONST; d.slot = (uint16_t)impl->literalSlotOf(t.literal); }
    else if (t.kind == Tok::CONSTVAR) { d.kind = SB2_TOK_CONST; d.slot = (uint16_t)impl->constSlotOf(t.var); }
    else if (t.kind == Tok::VAR && t.var->memSpeciesId >= 0 && t.var->geo == g) { d.kind = SB2_TOK_MEMVAR; d.slot = (uint16_t)t.var->memSpeciesId; }
    else if (t.kind == Tok::VAR) {
      // anything else keeps, at a membrane point, the value it was set up with (volume species are only ever advanced at
      // volume cells, their slopes at other points stay zero)
      d.kind = SB2_TOK_FIELD;
      for (size_t q = 0; q < idxs.size(); ++q) {
        int X, Y, Z;
        pointCoords(m, idxs[q], X, Y, Z);
        values.push_back(m.valueAtDoubled(t.var, X, Y, Z));
      }
    }
    toks.push_back(d);
  }
  std::vector<sb2_ref> refs;
  for (size_t k = 0; k < x->spRefs.size(); ++k) {
    sb2_ref r;
    Var* v = x->spRefs[k];
    r.species = (v && !v->isUniform && v->memSpeciesId >= 0 && v->geo == g) ? SB2_MEM_SPECIES + v->memSpeciesId : -1;
    r.is_variable = x->isVariable[k] ? 1 : 0;
    r.stoichiometry = x->stoichiometry[k];
    refs.push_back(r);
  }
  DEV(sb2_add_mem_reaction(impl->dev, g->devicePointSet, kind, toks.data(), (int)toks.size(), refs.data(), (int)refs.size(),
                           kind == 0 ? x->nReactants : 1, pts.data(), (int)pts.size(), values.empty() ? NULL : values.data()));
  return true;
}

// calcMemTransport call site (spatialsimulator.cpp:1393-1451) + per-point tables (calcPDE.cpp:1056-1479)
bool SpatialSimulator::addMemTransport(void* rx) {
  Rxn* x = static_cast<Rxn*>(rx);
  sb2host::Model& m = impl->m;
  Reaction* r = x->reaction;
  if (r->getNumReactants() == 0 || r->getNumProducts() == 0) return true;
  auto geoOfSpecies = [&](const std::string& sid) -> Geo* {
    Var* v = m.findVar(sid);
    if (v && v->geo) return v->geo;
    Species* s = m.model->getSpecies(sid);
    if (!s) return NULL;
    for (unsigned int i = 0; i < m.model->getNumSpecies(); ++i) {
      Species* cur = m.model->getSpecies(i);
      if (cur->getId() != sid && cur->getCompartment() == s->getCompartment()) {
        Var* cv = m.findVar(cur->getId());
        if (cv && cv->geo) return cv->geo;
      }
    }
    return NULL;
  };
  Geo* reactantGeo = geoOfSpecies(r->getReactant(0)->getSpecies());
  Geo* productGeo = geoOfSpecies(r->getProduct(0)->getSpecies());
  const bool haveRV = reactantGeo && reactantGeo->isVol, havePV = productGeo && productGeo->isVol;
  std::vector<Geo*> mems;  // every membrane calcMemTransport is called on for this reaction, in call order
  for (size_t j = 0; j < m.geos.size(); ++j) {
    Geo* g = m.geos[j];
    if (g->isVol) continue;
    if ((g->adjacent1 == reactantGeo && g->adjacent2 == productGeo) || (g->adjacent1 == productGeo && g->adjacent2 == reactantGeo)) { mems.push_back(g); break; }
  }
  // binding: one side of the reaction lives on a membrane (spatialsimulator.cpp:1440-1450)
  if (haveRV != havePV) {
    if (!haveRV && reactantGeo) mems.push_back(reactantGeo);
    if (!havePV && productGeo) mems.push_back(productGeo);
  }
  for (size_t q = 0; q < mems.size(); ++q)
    if (!addMemTransportOn(x, mems[q])) return false;
  return true;  // (no membrane between the two compartments: nothing happens)
}

bool SpatialSimulator::addMemTransportOn(void* rx, void* memGeo) {
  Rxn* x = static_cast<Rxn*>(rx);
  Geo* mem = static_cast<Geo*>(memGeo);
  sb2host::Model& m = impl->m;
  // On a slab the membrane comes from the host model of the whole grid (gm): its points, their normals and the volume masks
  // around them are looked up in whole-grid coordinates, the cells the device is handed are local ones.  A slab evaluates the
  // points next to the rows it owns and accumulates into owned cells only (the neighbouring slab does the same for its rows;
  // the two halo rows the operands may reach into are exchanged per stage, sb2_device.cu edge_w).
  const bool slab = m.isSlab();
  sb2host::Model* gmp = impl->wholeModel(lastError);
  if (!gmp) return false;
  sb2host::Model& gm = *gmp;
  if (slab) {
    mem = gm.geoByCompartment(mem->compartmentId);
    if (!mem) { lastError = "membrane transport '" + x->id + "': the membrane is missing from the whole-grid model"; return false; }
    for (size_t k = 0; k < x->spRefs.size(); ++k)
      if (x->spRefs[k] && x->spRefs[k]->onMembrane) { lastError = "reaction '" + x->id + "' binds to a membrane species; not supported on a sharded grid yet"; return false; }
  }
  const int slabAxis = m.dimension == 3 ? 2 : 1;
  const int rowOff = slab ? m.off[slabAxis] : 0;                      // first held row, in rows of the whole grid
  const int ownLoL = slab ? impl->ownLo - impl->heldLo : 0;           // owned rows, local
  const int ownHiL = slab ? impl->ownHi - impl->heldLo : (m.dimension == 3 ? m.nz : m.ny);

  const int Xi = gm.Xindex, Yi = gm.Yindex, Zi = gm.Zindex;
  const int64_t N = gm.ndoubled();
  // participating points: listed in domainIndex, off the frame, isDomain == 1
  std::vector<const MemPoint*> pts;
  std::map<int64_t, const MemPoint*> pointOf;
  for (size_t k = 0; k < mem->memPoints.size(); ++k) pointOf[mem->memPoints[k].index] = &mem->memPoints[k];
  for (size_t k = 0; k < mem->memDomainIndex.size(); ++k) {
    const int64_t idx = mem->memDomainIndex[k];
    const int Z = (int)(idx / ((int64_t)Xi * Yi)), Y = (int)((idx - (int64_t)Z * Xi * Yi) / Xi), X = (int)(idx - (int64_t)Z * Xi * Yi - (int64_t)Y * Xi);
    if (m.dimension == 3 && (X == 0 || Y == 0 || Z == 0 || X == Xi - 1 || Y == Yi - 1 || Z == Zi - 1)) continue;
    if (m.dimension == 2 && (X == 0 || Y == 0 || X == Xi - 1 || Y == Yi - 1)) continue;
    std::map<int64_t, const MemPoint*>::const_iterator it = pointOf.find(idx);
    if (it == pointOf.end()) continue;  // isDomain[index] != 1
    if (slab) {
      // the cells a point feeds lie at +-1 along its axis: doubled rows 2 * (ownLo + heldLo) - 1 ... 2 * (ownHi + heldLo - 1) + 1
      const int R = slabAxis == 2 ? Z : Y;
      if (R < 2 * (ownLoL + rowOff) - 1 || R > 2 * (ownHiL + rowOff - 1) + 1) continue;
    }
    pts.push_back(it->second);
  }
  const int npts = (int)pts.size();
  // compact LOCAL cell of a doubled-grid point of the whole grid, -1 unless it is a volume cell of `geo` that is held locally
  auto cellIn = [&](Geo* geo, int X, int Y, int Z) -> int32_t {
    if (!geo || !geo->isVol || !geo->hasMask) return -1;
    if (X < 0 || Y < 0 || Z < 0 || X >= Xi || Y >= Yi || Z >= Zi) return -1;
    if ((X | Y | Z) & 1) return -1;
    if (!slab) {
      const int64_t c = ((int64_t)(Z / 2) * m.ny + Y / 2) * m.nx + X / 2;
      return geo->isDomain[c] == 1 ? (int32_t)c : -1;
    }
    Geo* wg = gm.geoByCompartment(geo->compartmentId);
    if (!wg || !wg->hasMask) return -1;
    const int64_t cw = ((int64_t)(Z / 2) * gm.ny + Y / 2) * gm.nx + X / 2;
    if (wg->isDomain[cw] != 1) return -1;
    const int lj = Y / 2 - (slabAxis == 1 ? rowOff : 0), lk = Z / 2 - (slabAxis == 2 ? rowOff : 0);
    if (lj < 0 || lj >= m.ny || lk < 0 || lk >= m.nz) return -1;
    return (int32_t)(((int64_t)lk * m.ny + lj) * m.nx + X / 2);
  };
  // a target cell must be OWNED by this slab (halo cells receive their increments from the slab that owns them)
  auto ownedCell = [&](int32_t c) -> int32_t {
    if (!slab || c < 0) return c;
    const int row = slabAxis == 2 ? (int)(c / ((int64_t)m.nx * m.ny)) : (int)((c / m.nx) % m.ny);
    return (row >= ownLoL && row < ownHiL) ? c : -1;
  };

  std::vector<sb2_token> toks;
  std::vector<int32_t> operandCells;
  std::vector<double> operandValues;
  Var* symbol = NULL;  // persists across tokens like the reference's symbolInfo
  for (size_t i = 0; i < x->rpn.size(); ++i) {
    const Tok& t = x->rpn[i];
    sb2_token d;
    d.kind = SB2_TOK_NOP; d.op = 0; d.slot = 0;
    if (t.kind == Tok::OP) { d.kind = SB2_TOK_OP; d.op = (uint8_t)opOfAst(t.astType); }
    else if (t.kind == Tok::LITERAL) { d.kind = SB2_TOK_CONST; d.slot = (uint16_t)impl->literalSlotOf(t.literal); }
    else if (t.kind == Tok::CONSTVAR) { d.kind = SB2_TOK_CONST; d.slot = (uint16_t)impl->constSlotOf(t.var); }
    else if (t.kind == Tok::VAR && t.var->hasDelta && t.var->memSpeciesId >= 0) {
      // a symbol that lives on a membrane is read at the point itself (calcPDE.cpp:1214-1217); the reference decides
      // "volume or membrane" by the species reference it matched last
      for (size_t j = 0; j < x->spRefs.size(); ++j)
        if (x->spRefs[j] == t.var) { symbol = x->spRefs[j]; break; }
      d.kind = SB2_TOK_MEMVAR; d.slot = (uint16_t)t.var->memSpeciesId;
      for (int p = 0; p < npts; ++p) {
        operandCells.push_back(t.var->geo ? t.var->geo->storageOf(pts[p]->index) : -1);
        operandCells.push_back(-1);
      }
    }
    else if (t.kind == Tok::VAR && t.var->hasDelta && t.var->speciesId >= 0) {
      for (size_t j = 0; j < x->spRefs.size(); ++j)
        if (x->spRefs[j] == t.var) { symbol = x->spRefs[j]; break; }
      if (!symbol || !symbol->geo) { lastError = "membrane transport '" + x->id + "': operand '" + t.var->id + "' has no geometry"; return false; }
      d.kind = SB2_TOK_VAR; d.slot = (uint16_t)t.var->speciesId;
      for (int p = 0; p < npts; ++p) {
        const MemPoint& mp = *pts[p];
        int32_t nearC = -1, farC = -1;
        const int dX[3] = {1, 0, 0}, dY[3] = {0, 1, 0}, dZ[3] = {0, 0, 1};
        for (int a = 0; a < m.dimension; ++a) {  // later axes overwrite earlier ones
          int32_t c = cellIn(symbol->geo, mp.X + dX[a], mp.Y + dY[a], mp.Z + dZ[a]);
          int sgn = 1;
          if (c < 0) { c = cellIn(symbol->geo, mp.X - dX[a], mp.Y - dY[a], mp.Z - dZ[a]); sgn = -1; }
          if (c < 0) continue;
          nearC = c;
          const int64_t flat3 = mp.index + sgn * 3 * (a == 0 ? 1 : a == 1 ? (int64_t)Xi : (int64_t)Xi * Yi);
          farC = -1;
          if (flat3 < N && flat3 >= 0) {
            // flat-index arithmetic as in the reference: a wrapped index lands on a non-cell point
            const int Z3 = (int)(flat3 / ((int64_t)Xi * Yi)), Y3 = (int)((flat3 - (int64_t)Z3 * Xi * Yi) / Xi), X3 = (int)(flat3 - (int64_t)Z3 * Xi * Yi - (int64_t)Y3 * Xi);
            farC = cellIn(symbol->geo, X3, Y3, Z3);
          }
        }
        operandCells.push_back(nearC);
        operandCells.push_back(farC);
      }
    } else if (t.kind == Tok::VAR) {
      // no delta array: the raw value at the membrane point itself (calcPDE.cpp:1218-1221)
      d.kind = SB2_TOK_FIELD; d.slot = 0;
      for (int p = 0; p < npts; ++p) {
        const MemPoint& mp = *pts[p];
        operandValues.push_back(gm.valueAtDoubled(slab ? gm.findVar(t.var->id) : t.var, mp.X, mp.Y, mp.Z));
      }
    }
    toks.push_back(d);
  }
  // operand_cells is laid out [VAR token][point][2]; the loop above appended in exactly that order
  std::vector<sb2_ref> refs;
  std::vector<int32_t> targets;
  std::vector<sb2_mem_point> dpts(npts);
  for (int p = 0; p < npts; ++p) {
    const MemPoint& mp = *pts[p];
    const int axis = (mp.bof & (SB2_F_XP | SB2_F_XM)) ? 0 : (mp.bof & (SB2_F_YP | SB2_F_YM)) ? 1 : 2;
    dpts[p].axis = axis;
    dpts[p].abs_normal = fabs(axis == 0 ? mp.nx : axis == 1 ? mp.ny : mp.nz);
  }
  for (size_t k = 0; k < x->spRefs.size(); ++k) {
    sb2_ref rr;
    Var* v = x->spRefs[k];
    const bool ok = v && !v->isUniform && v->speciesId >= 0 && v->geo && v->geo->isVol;
    const bool onMem = v && !v->isUniform && v->memSpeciesId >= 0 && v->geo && !v->geo->isVol;
    rr.species = ok ? v->speciesId : onMem ? SB2_MEM_SPECIES + v->memSpeciesId : -1;
    rr.is_variable = x->isVariable[k] ? 1 : 0;
    rr.stoichiometry = x->stoichiometry[k];
    refs.push_back(rr);
    for (int p = 0; p < npts; ++p) {
      const MemPoint& mp = *pts[p];
      const int a = dpts[p].axis;
      const int dx = a == 0, dy = a == 1, dz = a == 2;
      if (onMem) {
        // binding: the species' own membrane holds this point (isDomain[index] == 1, calcPDE.cpp:1405)
        targets.push_back(v->geo->memIsDomain(mp.index) == 1 ? v->geo->storageOf(mp.index) : -1);
        targets.push_back(-1);
        continue;
      }
      targets.push_back(ok ? ownedCell(cellIn(v->geo, mp.X + dx, mp.Y + dy, mp.Z + dz)) : -1);
      targets.push_back(ok ? ownedCell(cellIn(v->geo, mp.X - dx, mp.Y - dy, mp.Z - dz)) : -1);
    }
  }
  int rc = sb2_add_mem_transport(impl->dev, toks.data(), (int)toks.size(), refs.data(), (int)refs.size(), x->nReactants,
                                 dpts.data(), npts, operandCells.empty() ? NULL : operandCells.data(),
                                 operandValues.empty() ? NULL : operandValues.data(), targets.empty() ? NULL : targets.data());
  if (rc < 0) { lastError = std::string("sb2_add_mem_transport: ") + sb2_last_error(impl->dev); return false; }
  return true;
}

// =============================================================================================
// stepping
// =============================================================================================
void SpatialSimulator::syncBeforeStep() {
  sb2host::Model& m = impl->m;
  if (!impl->slabs.empty()) {
    // group: values written through a getVariable() pointer go to the slabs that hold the rows (the mirror is current: handed-out
    // mirrors are refreshed after every step); then every slab settles its own constants
    for (std::map<std::string, VarMirror>::iterator it = impl->varMirrors.begin(); it != impl->varMirrors.end(); ++it) {
      if (!it->second.handedOut) continue;
      Var* v = m.findVar(it->first);
      if (!v || v->isUniform || v->onMembrane || impl->isCoordinate(v) || !impl->groupFresh.count(v)) continue;
      bool changed = false;
      const double* d = it->second.data.data();
      for (int k = 0; k < m.nz; ++k)
        for (int j = 0; j < m.ny; ++j) {
          const double* row = d + (int64_t)(2 * k) * m.Yindex * m.Xindex + (int64_t)(2 * j) * m.Xindex;
          double* f = v->field.data() + ((int64_t)k * m.ny + j) * m.nx;
          for (int i = 0; i < m.nx; ++i)
            if (memcmp(&row[2 * i], &f[i], sizeof(double)) != 0) { f[i] = row[2 * i]; changed = true; }
        }
      if (changed) {
        const std::vector<double> copy(v->field);
        setVariableCompact(it->first, copy.data(), copy.size());
      }
    }
    for (size_t i = 0; i < impl->slabs.size(); ++i) impl->slabs[i]->syncBeforeStep();
    return;
  }
  // values written through a pointer handed out by getVariable()
  for (std::map<std::string, VarMirror>::iterator it = impl->varMirrors.begin(); it != impl->varMirrors.end(); ++it) {
    if (!it->second.handedOut) continue;
    Var* v = m.findVar(it->first);
    if (!v || v->isUniform || m.isSlab()) continue;
    if (v->onMembrane) {
      if (v->memSpeciesId < 0 || impl->memStale[v->memSpeciesId] || !v->geo) continue;
      bool changed = false;
      for (size_t k = 0; k < v->geo->storage.size(); ++k) {
        const double x = it->second.data[(size_t)v->geo->storage[k]];
        if (memcmp(&x, &v->field[k], sizeof(double)) != 0) { v->field[k] = x; changed = true; }
      }
      if (changed) sb2_upload_mem_species(impl->dev, v->memSpeciesId, v->field.data());
      continue;
    }
    if (v->speciesId >= 0 && impl->stale[v->speciesId]) continue;  // mirror was not refreshed since the last step
    bool changed = false;
    const double* d = it->second.data.data();
    for (int k = 0; k < m.nz; ++k)
      for (int j = 0; j < m.ny; ++j) {
        const double* row = d + (int64_t)(2 * k) * m.Yindex * m.Xindex + (int64_t)(2 * j) * m.Xindex;
        double* f = v->field.data() + ((int64_t)k * m.ny + j) * m.nx;
        for (int i = 0; i < m.nx; ++i)
          if (memcmp(&row[2 * i], &f[i], sizeof(double)) != 0) { f[i] = row[2 * i]; changed = true; }
      }
    if (changed) {
      if (v->speciesId >= 0) sb2_upload_species(impl->dev, v->speciesId, v->field.data());
      else if (v->fieldId >= 0) sb2_upload_field(impl->dev, v->fieldId, v->field.data());
    }
  }
  // uniform variables alias the constants table (astFunction.cpp:76-81)
  for (size_t i = 0; i < m.vars.size(); ++i) {
    Var* v = m.vars[i];
    if (v->constSlot < 0 || v == m.tVar) continue;
    if (memcmp(&impl->consts[v->constSlot], &v->uniformValue, sizeof(double)) != 0) {
      impl->consts[v->constSlot] = v->uniformValue;
      sb2_set_constant(impl->dev, v->constSlot, v->uniformValue);
    }
  }
}

double SpatialSimulator::oneStep(double initialTime, double step) {
  if (impl && !impl->slabs.empty()) {
    std::vector<sb2_sim*> handles;
    syncBeforeStep();
    for (size_t i = 0; i < impl->slabs.size(); ++i) handles.push_back(impl->slabs[i]->impl->dev);
    impl->m.simTime = initialTime * step;
    impl->m.tVar->uniformValue = impl->m.simTime;
    const int rc = sb2_group_step(handles.data(), (int)handles.size(), initialTime, step, impl->slabs[0]->impl->timeSlot);
    if (rc < 0) {
      lastError = "sb2_group_step: ";
      for (size_t i = 0; i < handles.size(); ++i) if (*sb2_last_error(handles[i])) { lastError += sb2_last_error(handles[i]); break; }
      throw std::runtime_error(lastError);
    }
    for (size_t i = 0; i < impl->slabs.size(); ++i) impl->slabs[i]->deviceStateChanged();
    impl->groupFresh.clear();
    refreshHandedOut();
    return initialTime + step;
  }
  if (!impl || !impl->dev)
    throw std::runtime_error("SpatialSimulator::oneStep: no device simulator (" + (lastError.empty() ? std::string("not initialised") : lastError) +
                             "); there is no CPU fallback");
  syncBeforeStep();
  impl->m.simTime = initialTime * step;
  impl->m.tVar->uniformValue = impl->m.simTime;
  const int rc = sb2_step(impl->dev, initialTime, 0.0, step, 1, impl->timeSlot);
  if (rc < 0) {
    lastError = std::string("sb2_step: ") + sb2_last_error(impl->dev);
    throw std::runtime_error(lastError);
  }
  impl->markStale();
  refreshHandedOut();
  return initialTime + step;
}

double SpatialSimulator::oneStepHost(double initialTime, double step, int n, const char* const* ids, const double* const* in,
                                     double* const* out, int nslabs) {
  if (impl && !impl->slabs.empty()) {
    // group: every slab pipelines its own rows (sb2_group_step_host); models on the unfused schedule take the plain sequence
    const size_t G = impl->slabs.size();
    const int64_t rowCells = impl->m.dimension == 3 ? (int64_t)impl->m.nx * impl->m.ny : (int64_t)impl->m.nx;
    const size_t S = (size_t)impl->slabs[0]->getNumStagedSpecies();
    std::vector<std::vector<const double*> > din(G, std::vector<const double*>(S, (const double*)NULL));
    std::vector<std::vector<double*> > dout(G, std::vector<double*>(S, (double*)NULL));
    bool staged_only = in != NULL && out != NULL;
    for (int i = 0; i < n && staged_only; ++i) {
      const Var* v = impl->slabs[0]->impl->m.findVar(ids[i]);
      if (!v || v->speciesId < 0) { staged_only = false; break; }
      for (size_t g = 0; g < G; ++g) {
        const int64_t off = (int64_t)impl->slabs[g]->impl->heldLo * rowCells;
        din[g][v->speciesId] = in[i] ? in[i] + off : NULL;
        dout[g][v->speciesId] = out[i] ? out[i] + off : NULL;
      }
    }
    int rc = SB2_ERR_UNFUSED;
    if (staged_only) {
      std::vector<sb2_sim*> handles;
      std::vector<const double* const*> pin;
      std::vector<double* const*> pout;
      syncBeforeStep();
      for (size_t g = 0; g < G; ++g) {
        impl->slabs[g]->impl->m.simTime = initialTime * step;
        impl->slabs[g]->impl->m.tVar->uniformValue = initialTime * step;
        handles.push_back(impl->slabs[g]->impl->dev);
        pin.push_back(din[g].data());
        pout.push_back(dout[g].data());
      }
      rc = sb2_group_step_host(handles.data(), (int)G, initialTime, step, impl->slabs[0]->impl->timeSlot, pin.data(), pout.data(), nslabs);
      if (rc < 0 && rc != SB2_ERR_UNFUSED) {
        lastError = std::string("sb2_group_step_host: ") + sb2_last_error(handles[0]);
        throw std::runtime_error(lastError);
      }
    }
    if (rc == SB2_ERR_UNFUSED) {
      for (int i = 0; i < n; ++i)
        if (in && in[i] && !setVariableCompact(ids[i], in[i], (size_t)impl->m.ncells())) throw std::runtime_error(std::string("oneStepHost: cannot set ") + ids[i]);
      oneStep(initialTime, step);
      for (int i = 0; i < n; ++i)
        if (out && out[i] && !getVariableCompact(ids[i], out[i], (size_t)impl->m.ncells())) throw std::runtime_error(std::string("oneStepHost: cannot read ") + ids[i]);
      return initialTime + step;
    }
    impl->m.simTime = initialTime * step;
    impl->groupFresh.clear();
    for (size_t g = 0; g < G; ++g) impl->slabs[g]->deviceStateChanged();
    refreshHandedOut();
    return initialTime + step;
  }
  if (!impl || !impl->dev)
    throw std::runtime_error("SpatialSimulator::oneStepHost: no device simulator (" + (lastError.empty() ? std::string("not initialised") : lastError) +
                             "); there is no CPU fallback");
  syncBeforeStep();
  impl->m.simTime = initialTime * step;
  impl->m.tVar->uniformValue = impl->m.simTime;
  const size_t S = impl->stale.size();
  std::vector<const double*> din(S, (const double*)NULL);
  std::vector<double*> dout(S, (double*)NULL);
  std::vector<Var*> vars(n, (Var*)NULL);
  bool staged_only = true;
  for (int i = 0; i < n; ++i) {
    Var* v = impl->m.findVar(ids[i]);
    vars[i] = v;
    if (!v || v->speciesId < 0) { staged_only = false; continue; }
    din[v->speciesId] = in ? in[i] : NULL;
    dout[v->speciesId] = out ? out[i] : NULL;
  }
  int rc = staged_only ? sb2_step_host(impl->dev, initialTime, step, impl->timeSlot, din.data(), dout.data(), nslabs) : SB2_ERR_UNFUSED;
  if (rc == SB2_ERR_UNFUSED) {
    // unfused schedule (or a variable that is not a staged species): the plain sequence
    for (int i = 0; i < n; ++i)
      if (in && in[i] && !setVariableCompact(ids[i], in[i], (size_t)impl->m.ncells())) throw std::runtime_error(std::string("oneStepHost: cannot set ") + ids[i]);
    oneStep(initialTime, step);
    for (int i = 0; i < n; ++i)
      if (out && out[i] && !getVariableCompact(ids[i], out[i], (size_t)impl->m.ncells())) throw std::runtime_error(std::string("oneStepHost: cannot read ") + ids[i]);
    return initialTime + step;
  }
  if (rc < 0) {
    lastError = std::string("sb2_step_host: ") + sb2_last_error(impl->dev);
    throw std::runtime_error(lastError);
  }
  impl->markStale();
  refreshHandedOut();
  return initialTime + step;
}

void SpatialSimulator::runSteps(double firstStepIndex, double step, int nsteps) {
  if (impl && !impl->slabs.empty()) {
    for (int i = 0; i < nsteps; ++i) oneStep(firstStepIndex + i, step);
    return;
  }
  if (!impl || !impl->dev)
    throw std::runtime_error("SpatialSimulator::runSteps: no device simulator (" + (lastError.empty() ? std::string("not initialised") : lastError) +
                             "); there is no CPU fallback");
  if (nsteps <= 0) return;
  syncBeforeStep();
  const int rc = sb2_step(impl->dev, firstStepIndex, 1.0, step, nsteps, impl->timeSlot);
  if (rc < 0) {
    lastError = std::string("sb2_step: ") + sb2_last_error(impl->dev);
    throw std::runtime_error(lastError);
  }
  impl->m.simTime = (firstStepIndex + nsteps - 1) * step;
  impl->m.tVar->uniformValue = impl->m.simTime;
  impl->markStale();
  refreshHandedOut();
}

// The reference's getVariable() pointer aliases the live state (SpatialUI reads concentrations and applies doses through it,
// simulationthread.cpp:136-185).  Here it points to a host mirror: after every step the mirrors that were handed out are
// filled again from the device, so reads through the pointer see the new values and writes (compared with the host copy
// before the next step, syncBeforeStep) reach the device.  releaseVariable() ends that for callers that are done with a pointer.
void SpatialSimulator::refreshHandedOut() {
  if (!impl || impl->m.isSlab()) return;
  std::vector<std::string> ids;
  for (std::map<std::string, VarMirror>::iterator it = impl->varMirrors.begin(); it != impl->varMirrors.end(); ++it)
    if (it->second.handedOut) ids.push_back(it->first);
  int len = 0;
  for (size_t i = 0; i < ids.size(); ++i) getVariable(ids[i], len);
}

void SpatialSimulator::releaseVariable(const std::string& id) {
  if (!impl) return;
  std::map<std::string, VarMirror>::iterator it = impl->varMirrors.find(id);
  if (it != impl->varMirrors.end()) it->second.handedOut = false;
}

void SpatialSimulator::deviceStateChanged() {
  if (!impl) return;
  impl->groupFresh.clear();
  impl->markStale();
}

// spatialsimulator.cpp:1667-1703 without the gnuplot pipe
void SpatialSimulator::run(double endTime, double step) {
  const int last = static_cast<int>(endTime / step);
  int percent = 0;
  for (int t = 0; t <= last; ++t) {
    oneStep(t, step);
    if (t == (last / 10) * percent) {
      std::cout << percent * 10 << "% finished" << std::endl;
      ++percent;
    }
  }
}

double SpatialSimulator::getStableTimeStep(double dt) {
  if (impl && !impl->slabs.empty()) {
    double best = dt;  // the reduction of checkStability.cpp over the whole grid: the minimum over the slabs
    for (size_t i = 0; i < impl->slabs.size(); ++i) best = std::min(best, impl->slabs[i]->getStableTimeStep(dt));
    return best;
  }
  if (!impl || !impl->dev) throw std::runtime_error("SpatialSimulator::getStableTimeStep: no device simulator; there is no CPU fallback");
  syncBeforeStep();
  double out = dt;
  if (sb2_min_dt(impl->dev, dt, &out) < 0) {
    lastError = std::string("sb2_min_dt: ") + sb2_last_error(impl->dev);
    throw std::runtime_error(lastError);
  }
  return out;
}

// =============================================================================================
// parameters
// =============================================================================================
void SpatialSimulator::setParameterUniformly(const std::string& id, double value) {
  if (!impl) return;
  sb2host::Model& m = impl->m;
  Var* v = m.findVar(id);
  if (!v || !v->hasValue) return;
  for (size_t i = 0; i < impl->slabs.size(); ++i) impl->slabs[i]->setParameterUniformly(id, value);
  impl->groupFresh.erase(v);
  if (v->isUniform) {
    v->uniformValue = value;  // reaches the device constants table in syncBeforeStep()
    return;
  }
  // flood the field (every doubled-grid point in the reference)
  if (v->speciesId >= 0 && impl->dev && impl->stale[v->speciesId]) impl->stale[v->speciesId] = 0;
  std::fill(v->field.begin(), v->field.end(), value);
  v->offGridInit = value;
  if (impl->dev) {
    if (v->speciesId >= 0) sb2_upload_species(impl->dev, v->speciesId, v->field.data());
    else if (v->fieldId >= 0) sb2_upload_field(impl->dev, v->fieldId, v->field.data());
  }
  std::map<std::string, VarMirror>::iterator it = impl->varMirrors.find(id);
  if (it != impl->varMirrors.end()) std::fill(it->second.data.begin(), it->second.data.end(), value);
}

// spatialsimulator.cpp:167-202
void SpatialSimulator::setParameter(const std::string& id, double value) {
  if (!impl || !impl->m.model) return;
  sb2host::Model& m = impl->m;
  Parameter* param = m.model->getParameter(id);
  if (!param) return;
  param->setValue(value);
  SpatialParameterPlugin* pp = static_cast<SpatialParameterPlugin*>(param->getPlugin("spatial"));
  if (pp && pp->getDiffusionCoefficient() && !pp->getDiffusionCoefficient()->getVariable().empty()) {
    DiffusionCoefficient* dc = pp->getDiffusionCoefficient();
    Var* s = m.findVar(dc->getVariable());
    if (s && s->hasDiff) {
      const int a = (int)dc->getCoordinateReference1();  // the enum value is the array index (:180)
      Var* c = (a >= 0 && a < 3) ? s->diffC[a] : NULL;
      if (c && c->hasValue) setParameterUniformly(c->id, value);
    }
  }
  setParameterUniformly(id, value);
}

// spatialsimulator.cpp:1729-1767
void SpatialSimulator::flipVolumeOrder() {
  if (!impl || !impl->m.model) return;
  SpatialModelPlugin* plugin = dynamic_cast<SpatialModelPlugin*>(impl->m.model->getPlugin("spatial"));
  if (!plugin || !plugin->getGeometry()) return;
  Geometry* geometry = plugin->getGeometry();
  for (unsigned int i = 0; i < geometry->getNumGeometryDefinitions(); ++i) {
    GeometryDefinition* gd = geometry->getGeometryDefinition(i);
    if (!gd->isAnalyticGeometry()) continue;
    AnalyticGeometry* ag = static_cast<AnalyticGeometry*>(gd);
    const int n = (int)ag->getNumAnalyticVolumes();
    std::vector<int> ordinals(n);
    for (int k = 0; k < n; ++k) ordinals[k] = ag->getAnalyticVolume(k)->getOrdinal();
    for (int k = 0; k < n; ++k) ag->getAnalyticVolume(k)->setOrdinal((n - 1) - ordinals[k]);
  }
}

// =============================================================================================
// read-back
// =============================================================================================
int SpatialSimulator::getVariableLength() const {
  if (!impl) return 0;
  const sb2host::Model& m = impl->m;
  return (int)((int64_t)(m.Zindex - 1) * m.Yindex * m.Xindex + (int64_t)(m.Yindex - 1) * m.Xindex + m.Xindex - 1);
}
int SpatialSimulator::getGeometryLength() const { return impl ? (int)impl->m.ndoubled() : 0; }
long SpatialSimulator::getNumDomainCells() const { return impl ? impl->domainCells : 0; }
int SpatialSimulator::getNumStagedSpecies() const {
  if (impl && !impl->slabs.empty()) return impl->slabs[0]->getNumStagedSpecies();
  return impl ? (int)impl->staged.size() : 0;
}
int SpatialSimulator::getTimeSlot() const {
  if (impl && !impl->slabs.empty()) return impl->slabs[0]->getTimeSlot();
  return impl ? impl->timeSlot : -1;
}
long SpatialSimulator::getDeviceLaunchCount() const {
  if (impl && !impl->slabs.empty()) {
    long n = 0;
    for (size_t i = 0; i < impl->slabs.size(); ++i) n += impl->slabs[i]->getDeviceLaunchCount();
    return n;
  }
  return impl && impl->dev ? (long)sb2_launch_count(impl->dev) : 0;
}

namespace {
// bring Var::field up to date with the device
bool refreshCompact(SimImpl* impl, Var* v) {
  if (impl->rulesStale && impl->dev) {  // fields rewritten by assignment rules
    for (size_t i = 0; i < impl->ruleTargets.size(); ++i) {
      Var* r = impl->ruleTargets[i];
      if (r->fieldId >= 0 && sb2_download_field(impl->dev, r->fieldId, r->field.data()) < 0) return false;
    }
    impl->rulesStale = false;
  }
  if (v->memSpeciesId >= 0 && impl->dev && impl->memStale[v->memSpeciesId]) {
    if (sb2_download_mem_species(impl->dev, v->memSpeciesId, v->field.data()) < 0) return false;
    impl->memStale[v->memSpeciesId] = 0;
    return true;
  }
  if (v->speciesId < 0 || !impl->dev || !impl->stale[v->speciesId]) return true;
  if (sb2_download_species(impl->dev, v->speciesId, v->field.data()) < 0) return false;
  impl->stale[v->speciesId] = 0;
  return true;
}
}  // namespace

bool SpatialSimulator::getVariableCompact(const std::string& id, double* out, size_t n) {
  if (!impl || !out) return false;
  Var* v = impl->m.findVar(id);
  if (!v || !v->hasValue) return false;
  if (v->isUniform) {
    for (size_t i = 0; i < n; ++i) out[i] = v->uniformValue;
    return true;
  }
  if (!impl->slabs.empty()) {
    if (v->onMembrane || n != (size_t)impl->m.ncells()) return false;
    if (impl->groupFresh.count(v)) { memcpy(out, v->field.data(), n * sizeof(double)); return true; }
    return gatherCompact(id, out);
  }
  if (v->onMembrane) {  // one value per storage point of its membrane (getMembranePoints)
    if (n != v->field.size() || !refreshCompact(impl, v)) return false;
    memcpy(out, v->field.data(), n * sizeof(double));
    return true;
  }
  if (n != (size_t)impl->m.ncells()) return false;
  // device state newer than the host copy: copy straight into the caller's buffer (which may be
  // pinned) instead of going through the host copy
  if (v->speciesId >= 0 && impl->dev && impl->stale[v->speciesId]) return sb2_download_species(impl->dev, v->speciesId, out) >= 0;
  memcpy(out, v->field.data(), n * sizeof(double));
  return true;
}

bool SpatialSimulator::setVariableCompact(const std::string& id, const double* in, size_t n) {
  if (!impl || !in) return false;
  Var* v = impl->m.findVar(id);
  if (!v || !v->hasValue || v->isUniform) return false;
  if (!impl->slabs.empty()) {
    // every slab takes the rows it holds (its own and the halo rows)
    if (v->onMembrane || n != (size_t)impl->m.ncells()) return false;
    const int64_t rowCells = impl->m.dimension == 3 ? (int64_t)impl->m.nx * impl->m.ny : (int64_t)impl->m.nx;
    for (size_t i = 0; i < impl->slabs.size(); ++i) {
      SpatialSimulator* sl = impl->slabs[i];
      if (!sl->setVariableCompact(id, in + (int64_t)sl->impl->heldLo * rowCells, (size_t)sl->impl->m.ncells())) return false;
    }
    memcpy(v->field.data(), in, n * sizeof(double));
    impl->groupFresh.insert(v);
    return true;
  }
  if (v->onMembrane) {
    if (n != v->field.size()) return false;
    memcpy(v->field.data(), in, n * sizeof(double));
    if (impl->dev && v->memSpeciesId >= 0) {
      impl->memStale[v->memSpeciesId] = 0;
      return sb2_upload_mem_species(impl->dev, v->memSpeciesId, v->field.data()) >= 0;
    }
    return true;
  }
  if (n != (size_t)impl->m.ncells()) return false;
  if (impl->dev && v->speciesId >= 0) {
    // upload straight from the caller's buffer; the host copy is refreshed on demand
    if (sb2_upload_species(impl->dev, v->speciesId, in) < 0) return false;
    impl->stale[v->speciesId] = 1;
    return true;
  }
  memcpy(v->field.data(), in, n * sizeof(double));
  if (impl->dev && v->fieldId >= 0) return sb2_upload_field(impl->dev, v->fieldId, v->field.data()) >= 0;
  return true;
}

double* SpatialSimulator::getVariable(const std::string& id, int& length) {
  length = 0;
  if (!impl) return NULL;
  sb2host::Model& m = impl->m;
  Var* v = m.findVar(id);
  if (!v || !v->hasValue) return NULL;
  length = getVariableLength();
  if (v->isUniform) return &v->uniformValue;  // the reference hands out the single double as well
  if (m.isSlab()) { length = 0; lastError = "doubled-grid mirrors are not available on a sharded grid; use getVariableCompact"; return NULL; }
  bool wasStale = (v->speciesId >= 0 && impl->stale[v->speciesId]) || (v->memSpeciesId >= 0 && impl->memStale[v->memSpeciesId]);
  if (!impl->slabs.empty() && !v->onMembrane && !impl->isCoordinate(v) && !impl->groupFresh.count(v)) {
    if (!gatherCompact(id, v->field.data())) { length = 0; return NULL; }
    impl->groupFresh.insert(v);
    wasStale = true;
  }
  if (!refreshCompact(impl, v)) { length = 0; return NULL; }
  VarMirror& mir = impl->varMirrors[id];
  const bool fresh = mir.data.empty();
  const int Xi = m.Xindex, Yi = m.Yindex, Zi = m.Zindex;
  if (v->onMembrane) {
    // zero off the membrane (boundaryFunction.cpp:97, 169), the species' values at its membrane / pseudo-membrane points
    if (fresh) mir.data.assign((size_t)m.ndoubled(), 0.0);
    if ((fresh || wasStale || !mir.handedOut) && v->geo)
      for (size_t k = 0; k < v->geo->storage.size() && k < v->field.size(); ++k) mir.data[(size_t)v->geo->storage[k]] = v->field[k];
    mir.handedOut = true;
    return mir.data.data();
  }
  if (fresh) {
    mir.data.assign((size_t)m.ndoubled(), v->offGridInit);
    if (impl->isCoordinate(v)) {
      for (int Z = 0; Z < Zi; ++Z)
        for (int Y = 0; Y < Yi; ++Y)
          for (int X = 0; X < Xi; ++X)
            mir.data[((int64_t)Z * Yi + Y) * Xi + X] = impl->coordinateAt(v, v == m.xVar ? X : v == m.yVar ? Y : Z);
    } else if (!v->sp && v->hasRpn) {
      // a parameter set by an initial assignment / assignment rule is evaluated on EVERY point of the doubled grid by the
      // reference (isAllArea, spatialsimulator.cpp:1126-1139): the points that are not volume cells are filled here
      for (int Z = 0; Z < Zi; ++Z)
        for (int Y = 0; Y < Yi; ++Y)
          for (int X = 0; X < Xi; ++X)
            if ((X | Y | Z) & 1) mir.data[((int64_t)Z * Yi + Y) * Xi + X] = m.valueAtDoubled(v, X, Y, Z);
    }
  }
  if (fresh || wasStale || !mir.handedOut) {
    for (int k = 0; k < m.nz; ++k)
      for (int j = 0; j < m.ny; ++j) {
        double* row = mir.data.data() + (int64_t)(2 * k) * Yi * Xi + (int64_t)(2 * j) * Xi;
        const double* f = v->field.data() + ((int64_t)k * m.ny + j) * m.nx;
        for (int i = 0; i < m.nx; ++i) row[2 * i] = f[i];
      }
    // CIP face values of an advected species live at the odd points of its array
    if (v->speciesId >= 0 && v->hasAdv && impl->dev) {
      std::vector<double> faces((size_t)m.ncells());
      const int64_t stride[3] = {1, Xi, (int64_t)Xi * Yi};
      for (int a = 0; a < m.dimension; ++a) {
        if (sb2_download_faces(impl->dev, v->speciesId, a, faces.data()) < 0) continue;
        const int lim[3] = {m.nx - (a == 0), m.ny - (a == 1), m.nz - (a == 2)};
        for (int k = 0; k < lim[2]; ++k)
          for (int j = 0; j < lim[1]; ++j)
            for (int i = 0; i < lim[0]; ++i)
              mir.data[(int64_t)(2 * k) * Yi * Xi + (int64_t)(2 * j) * Xi + 2 * i + stride[a]] = faces[((int64_t)k * m.ny + j) * m.nx + i];
      }
    }
  }
  mir.handedOut = true;
  return mir.data.data();
}

int* SpatialSimulator::getGeometry(const std::string& id, int& length) {
  length = 0;
  if (!impl) return NULL;
  sb2host::Model& m = impl->m;
  Geo* g = m.geoByCompartment(id);
  if (!g || m.isSlab()) return NULL;
  length = (int)m.ndoubled();
  if (!g->hasMask) return NULL;  // the reference returns its NULL isDomain pointer
  std::vector<int>& out = impl->geoMirrors[id];
  if (out.empty()) {
    out.assign((size_t)m.ndoubled(), 0);
    if (g->isVol) {
      for (int k = 0; k < m.nz; ++k)
        for (int j = 0; j < m.ny; ++j)
          for (int i = 0; i < m.nx; ++i)
            out[(int64_t)(2 * k) * m.Yindex * m.Xindex + (int64_t)(2 * j) * m.Xindex + 2 * i] = g->isDomain[((int64_t)k * m.ny + j) * m.nx + i];
    } else {
      for (std::unordered_map<int64_t, uint8_t>::const_iterator it = g->memClass.begin(); it != g->memClass.end(); ++it) out[it->first] = it->second;
    }
  }
  return out.data();
}

boundaryType* SpatialSimulator::getBoundaryType(const std::string& id, int& length) {
  length = 0;
  if (!impl) return NULL;
  sb2host::Model& m = impl->m;
  Geo* g = m.geoByCompartment(id);
  if (!g || m.isSlab()) return NULL;
  length = (int)m.ndoubled();
  if (!g->hasMask) return NULL;
  std::vector<boundaryType>& out = impl->btMirrors[id];
  if (out.empty()) {
    boundaryType none = {false, false, false, false, false, false};
    out.assign((size_t)m.ndoubled(), none);
    auto set = [](boundaryType& b, uint8_t f) {
      b.isBofXp = f & SB2_F_XP; b.isBofXm = f & SB2_F_XM; b.isBofYp = f & SB2_F_YP;
      b.isBofYm = f & SB2_F_YM; b.isBofZp = f & SB2_F_ZP; b.isBofZm = f & SB2_F_ZM;
    };
    if (g->isVol) {
      for (int k = 0; k < m.nz; ++k)
        for (int j = 0; j < m.ny; ++j)
          for (int i = 0; i < m.nx; ++i)
            set(out[(int64_t)(2 * k) * m.Yindex * m.Xindex + (int64_t)(2 * j) * m.Xindex + 2 * i], g->flags[((int64_t)k * m.ny + j) * m.nx + i]);
    } else {
      for (size_t k = 0; k < g->memPoints.size(); ++k) set(out[g->memPoints[k].index], g->memPoints[k].bof);
    }
  }
  return out.data();
}

int* SpatialSimulator::getBoundary(const std::string& id, int& length) {
  length = 0;
  if (!impl) return NULL;
  sb2host::Model& m = impl->m;
  Geo* g = m.geoByCompartment(id);
  if (!g || m.isSlab()) return NULL;
  length = (int)m.ndoubled();
  if (!g->hasMask) return NULL;
  std::vector<int>& out = impl->boundaryMirrors[id];
  if (out.empty()) {
    // only volume cells are ever written by the reference; the other entries of its array are
    // uninitialised heap and read as 0 here
    out.assign((size_t)m.ndoubled(), 0);
    if (g->isVol)
      for (int k = 0; k < m.nz; ++k)
        for (int j = 0; j < m.ny; ++j)
          for (int i = 0; i < m.nx; ++i)
            out[(int64_t)(2 * k) * m.Yindex * m.Xindex + (int64_t)(2 * j) * m.Xindex + 2 * i] = g->isBoundary[((int64_t)k * m.ny + j) * m.nx + i];
  }
  return out.data();
}

double* SpatialSimulator::getX(int& length) {
  length = 0;
  if (!impl || !impl->m.xVar) return NULL;
  return getVariable(impl->m.xVar->id, length);
}
double* SpatialSimulator::getY(int& length) {
  length = 0;
  if (!impl || !impl->m.yVar) return NULL;
  return getVariable(impl->m.yVar->id, length);
}
double* SpatialSimulator::getZ(int& length) {
  length = 0;
  if (!impl || !impl->m.zVar) return NULL;
  return getVariable(impl->m.zVar->id, length);
}

// spatialsimulator.cpp:1769-1783: coordinates are matched with ==, the z plane is always 0
int SpatialSimulator::getIndexForPosition(double x, double y) {
  if (!impl || !impl->m.xVar || !impl->m.yVar || impl->m.isSlab()) return -1;
  const sb2host::Model& m = impl->m;
  int ci = -1, cj = -1;
  for (int i = 0; i < m.nx && ci < 0; ++i) if (impl->coordinateAt(m.xVar, 2 * i) == x) ci = i;
  for (int j = 0; j < m.ny && cj < 0; ++j) if (impl->coordinateAt(m.yVar, 2 * j) == y) cj = j;
  if (ci < 0 || cj < 0) return -1;
  return (2 * cj) * m.Xindex + 2 * ci;
}

// spatialsimulator.cpp:1316-1332
double SpatialSimulator::getVariableAt(const std::string& id, int x, int y, int z) {
  (void)z;
  if (!impl) return 0.0;
  sb2host::Model& m = impl->m;
  Var* v = m.findVar(id);
  if (!v) return 0.0;
  const int index = getIndexForPosition(x, y);
  if (index == -1) return 0.0;
  if (!v->geo || !v->geo->hasMask || v->isUniform || !v->hasValue) return 0.0;
  const int Y = index / m.Xindex, X = index - Y * m.Xindex;
  const int64_t c = (int64_t)(Y / 2) * m.nx + X / 2;
  if (!v->geo->isVol) return 0.0;
  if (v->geo->isDomain[c] == 0) return 0.0;
  if (!impl->slabs.empty() && !impl->groupFresh.count(v)) {
    if (!gatherCompact(id, v->field.data())) return 0.0;
    impl->groupFresh.insert(v);
  }
  if (!refreshCompact(impl, v)) return 0.0;
  return v->field[c];
}

// =============================================================================================
// inspection
// =============================================================================================
bool SpatialSimulator::specializeNow() {
  if (impl && !impl->slabs.empty()) {
    bool all = true;
    for (size_t i = 0; i < impl->slabs.size(); ++i) all = impl->slabs[i]->specializeNow() && all;
    return all;
  }
  return impl && impl->dev && sb2_specialize_now(impl->dev) == 1;
}

bool SpatialSimulator::isSpecialized() {
  if (impl && !impl->slabs.empty()) {
    for (size_t i = 0; i < impl->slabs.size(); ++i) if (!impl->slabs[i]->isSpecialized()) return false;
    return true;
  }
  return impl && impl->dev && sb2_specialized(impl->dev) == 1;
}

std::string SpatialSimulator::getSpecializeInfo() {
  if (impl && !impl->slabs.empty()) return impl->slabs[0]->getSpecializeInfo();
  if (!impl || !impl->dev) return "";
  std::vector<char> buf(1 << 14);
  sb2_specialize_info(impl->dev, buf.data(), (int)buf.size());
  return std::string(buf.data());
}

std::string SpatialSimulator::getDeviceProgramText() {
  if (!impl || !impl->dev) return "";
  std::vector<char> buf(1 << 16);
  sb2_program_text(impl->dev, buf.data(), (int)buf.size());
  return std::string(buf.data());
}

namespace {
// same text layout as the oracle's dump (oracle/ref_driver.cpp: rpnText)
std::string rpnText(const Rpn& rpn) {
  std::ostringstream os;
  os.precision(17);
  for (size_t i = 0; i < rpn.size(); ++i) {
    const Tok& t = rpn[i];
    if (t.kind == Tok::VAR) os << "V " << t.var->id << " " << (t.var->hasDelta ? 1 : 0) << "\n";
    else if (t.kind == Tok::CONSTVAR) os << "C " << t.var->id << " " << t.var->uniformValue << "\n";
    else if (t.kind == Tok::LITERAL) os << "C # " << t.literal << "\n";
    else os << "O " << t.astType << "\n";
  }
  return os.str();
}
}  // namespace

std::string SpatialSimulator::getSetupText(const std::string& what) {
  if (!impl) return "";
  sb2host::Model& m = impl->m;
  std::ostringstream os;
  os.precision(17);
  if (what == "summary") {
    os << "dims " << m.nx << " " << m.ny << " " << m.nz << " dimension " << m.dimension << " delta " << m.deltaX << " " << m.deltaY << " " << m.deltaZ << "\n";
    for (size_t i = 0; i < m.geos.size(); ++i) {
      Geo* g = m.geos[i];
      os << "geo " << g->compartmentId << " domainType " << g->domainTypeId << " isVol " << g->isVol << " cells " << g->nDomain
         << " memPoints " << g->memPoints.size() << " pseudo " << g->pseudoMemIndex.size() << " adjacent "
         << (g->adjacent1 ? g->adjacent1->compartmentId : "-") << " " << (g->adjacent2 ? g->adjacent2->compartmentId : "-") << "\n";
    }
    for (size_t i = 0; i < m.vars.size(); ++i) {
      Var* v = m.vars[i];
      os << "var " << v->id << " uniform " << v->isUniform << " hasValue " << v->hasValue << " hasDelta " << v->hasDelta << " geo "
         << (v->geo ? v->geo->compartmentId : "-") << " species " << v->speciesId << " field " << v->fieldId << " const " << v->constSlot;
      if (v->isUniform && v->hasValue) os << " value " << v->uniformValue;
      os << "\n";
    }
    for (size_t i = 0; i < m.rxns.size(); ++i)
      os << "rxn " << i << " " << m.rxns[i]->id << " transport " << m.rxns[i]->isMemTransport << " tokens " << m.rxns[i]->rpn.size() << "\n";
    return os.str();
  }
  if (what.compare(0, 4, "rxn/") == 0) {
    const size_t i = (size_t)atoi(what.c_str() + 4);
    if (i < m.rxns.size()) return rpnText(m.rxns[i]->rpn);
    return "";
  }
  if (what.compare(0, 6, "arule/") == 0) {
    // ordered assignment rules (orderedARule of the reference): the target's id, "species" / "parameter", then its stream
    const size_t i = (size_t)atoi(what.c_str() + 6);
    if (i >= m.orderedAssignmentRules.size()) return "";
    const Var* v = m.orderedAssignmentRules[i];
    return v->id + "\n" + (v->sp ? "species" : "parameter") + "\n" + rpnText(v->rpn);
  }
  if (what.compare(0, 5, "refs/") == 0) {
    const size_t i = (size_t)atoi(what.c_str() + 5);
    if (i >= m.rxns.size()) return "";
    for (size_t k = 0; k < m.rxns[i]->spRefs.size(); ++k)
      os << (m.rxns[i]->spRefs[k] ? m.rxns[i]->spRefs[k]->id : "?") << " " << m.rxns[i]->stoichiometry[k] << " " << (m.rxns[i]->isVariable[k] ? 1 : 0) << "\n";
    return os.str();
  }
  if (what.compare(0, 4, "geo/") == 0) {
    Geo* g = m.geoByCompartment(what.substr(4));
    return g ? rpnText(g->rpn) : "";
  }
  if (what.compare(0, 8, "normals/") == 0) {
    Geo* g = m.geoByCompartment(what.substr(8));
    if (!g) return "";
    for (size_t k = 0; k < g->memPoints.size(); ++k)
      os << g->memPoints[k].index << " " << g->memPoints[k].nx << " " << g->memPoints[k].ny << " " << g->memPoints[k].nz << "\n";
    return os.str();
  }
  if (what.compare(0, 10, "memindex/") == 0 || what.compare(0, 9, "memindex/") == 0) {
    Geo* g = m.geoByCompartment(what.substr(9));
    if (!g) return "";
    for (size_t k = 0; k < g->memDomainIndex.size(); ++k) os << g->memDomainIndex[k] << "\n";
    return os.str();
  }
  if (what.compare(0, 8, "voronoi/") == 0) {
    // per domainIndex entry: the six neighbours [XY0 YZ0 XZ0 XY1 YZ1 XZ1], then s_ij and d_ij in the same slot order
    // (the layout oracle/ref_driver.cpp dumps vorI in)
    Geo* g = m.geoByCompartment(what.substr(8));
    if (!g) return "";
    static const int slot[6] = {0, 2, 4, 1, 3, 5};
    for (size_t k = 0; k < g->memDomainIndex.size(); ++k) {
      std::unordered_map<int64_t, Geo::Voronoi>::const_iterator it = g->voronoi.find(g->memDomainIndex[k]);
      Geo::Voronoi v;
      if (it != g->voronoi.end()) v = it->second;
      for (int q = 0; q < 6; ++q) os << v.adj[slot[q]] << " ";
      for (int q = 0; q < 6; ++q) os << v.si[slot[q]] << " ";
      for (int q = 0; q < 6; ++q) os << v.di[slot[q]] << (q == 5 ? "\n" : " ");
    }
    return os.str();
  }
  if (what.compare(0, 8, "storage/") == 0) {
    Geo* g = m.geoByCompartment(what.substr(8));
    if (!g) return "";
    for (size_t k = 0; k < g->storage.size(); ++k) os << g->storage[k] << "\n";
    return os.str();
  }
  if (what.compare(0, 7, "pseudo/") == 0) {
    Geo* g = m.geoByCompartment(what.substr(7));
    if (!g) return "";
    for (size_t k = 0; k < g->pseudoMemIndex.size(); ++k) os << g->pseudoMemIndex[k] << "\n";
    return os.str();
  }
  return "";
}
