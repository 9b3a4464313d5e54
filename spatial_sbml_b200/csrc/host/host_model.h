// host_model.h — host-side model set-up on COMPACT arrays.
//
// Restates what the reference does in initFromModel (spatialsimulator.cpp:275-1223),
// setInfoFunction.cpp:72-494, boundaryFunction.cpp:13-217 and astFunction.cpp:18-235, but stores one
// entry per VOLUME CELL (nx*ny*nz, x fastest; compact cell (i,j,k) = doubled-grid point (2i,2j,2k))
// and keeps membrane points as sparse lists, instead of ~15 arrays over the doubled grid.
// Everything the device needs leaves this layer through the C ABI in include/spatial_b200.h.
#ifndef SB2_HOST_MODEL_H_
#define SB2_HOST_MODEL_H_

#include <stdint.h>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../sbml/sbml_lite.h"

namespace sb2host {

struct Geo;
struct Var;

// One slot of a reverse-polish stream (reference reversePolishInfo, mystruct.h:20-26).
struct Tok {
  enum Kind { OP, VAR, CONSTVAR, LITERAL, NOP };
  Kind kind;
  int astType;     // OP: the AST node type
  Var* var;        // VAR (field valued) / CONSTVAR (uniform: aliases the variable's live value)
  double literal;  // LITERAL
};
typedef std::vector<Tok> Rpn;

// variableInfo (mystruct.h:63-81) on compact storage.
struct Var {
  std::string id;
  Species* sp;
  Compartment* com;
  Parameter* para;
  bool hasValue;    // reference: value != NULL
  bool isUniform;
  bool isResolved;
  bool inVol;
  bool hasDelta;    // reference: delta != NULL (a species the integrator advances)
  bool hasAssignmentRule;
  double uniformValue;
  std::vector<double> field;  // compact values (non-uniform variables)
  double offGridInit;         // initial value of the doubled-grid points that are not volume cells
  Geo* geo;
  Var* diffC[3];
  Var* advC[3];
  Var* bc[6];
  bool hasDiff, hasAdv, hasBC;
  int bcKind[6];  // SB2_BC_*
  Rpn rpn;        // initial assignment / assignment rule
  bool hasRpn;
  std::vector<Var*> dependence;
  // A species whose compartment is a membrane: `field` holds one value per storage point of its geometry (Geo::storage)
  // instead of one per volume cell.
  bool onMembrane;
  // device binding
  int constSlot, fieldId, speciesId, memSpeciesId;
  Var();
};

struct MemPoint {
  int64_t index;  // doubled-grid flat index
  int X, Y, Z;    // doubled-grid coordinates
  uint8_t bof;    // SB2_F_XP|XM / YP|YM / ZP|ZM pair (membrane bType)
  double nx, ny, nz;  // normal unit vector (setNormalAngle)
};

// GeometryInfo (mystruct.h:43-61).
struct Geo {
  std::string compartmentId, domainTypeId;
  bool isVol;
  bool hasMask;                     // reference: isDomain != NULL
  std::vector<uint8_t> isDomain;    // compact, volumes: 0/1
  std::vector<uint8_t> isBoundary;  // compact, volumes
  std::vector<uint8_t> flags;       // compact, volumes: SB2_F_* (bit0 isDomain, bits1-6 bType)
  int64_t nDomain;
  Geo* adjacent1;
  Geo* adjacent2;
  // membranes: sparse doubled-grid classification (isDomain 1 = membrane point, 2 = pseudo membrane)
  std::unordered_map<int64_t, uint8_t> memClass;
  std::vector<MemPoint> memPoints;         // in scan order (reference domainIndex order)
  std::vector<int64_t> memDomainIndex;     // what the reference pushes into domainIndex (incl. its quirk)
  std::vector<int64_t> pseudoMemIndex;
  Rpn rpn;
  int deviceGeo;
  // membranes that host species: the doubled-grid points such a species needs storage for (membrane points, pseudo-membrane
  // points and whatever else domainIndex names), ascending; Var::field of a membrane species is aligned with it
  std::vector<int64_t> storage;
  int storageOf(int64_t idx) const;  // -1: no storage
  // voronoiInfo (mystruct.h:116-129) of setVoronoiInfo (setInfoFunction.cpp:1134-1546), keyed by doubled-grid index; slots are
  // [XY0, XY1, YZ0, YZ1, XZ0, XZ1]
  struct Voronoi {
    int64_t adj[6];
    double si[6], di[6];
    bool isAve[6];
    Voronoi() { for (int i = 0; i < 6; ++i) { adj[i] = -1; si[i] = 1.0; di[i] = 0.0; isAve[i] = false; } }
  };
  std::unordered_map<int64_t, Voronoi> voronoi;
  int devicePointSet;
  Geo() : isVol(true), hasMask(false), nDomain(0), adjacent1(NULL), adjacent2(NULL), deviceGeo(-1), devicePointSet(-1) {}
  int memIsDomain(int64_t idx) const {
    std::unordered_map<int64_t, uint8_t>::const_iterator it = memClass.find(idx);
    return it == memClass.end() ? 0 : it->second;
  }
};

// reactionInfo (mystruct.h:83-92).
struct Rxn {
  std::string id;
  Reaction* reaction;  // NULL for rate rules
  bool isMemTransport;
  Rpn rpn;
  std::vector<Var*> spRefs;
  std::vector<bool> isVariable;
  std::vector<double> stoichiometry;
  int nReactants;
};

struct Model {
  SBMLDocument* doc;
  ::Model* model;
  int nx, ny, nz;  // local cells per axis (the whole grid, or a slab plus its halo rows / planes)
  int gn[3];       // Xdiv, Ydiv, Zdiv of the whole grid
  int off[3];      // global index of local cell 0 per axis (non-zero only along the slab axis)
  int Xindex, Yindex, Zindex;
  int dimension;
  unsigned volDimension, memDimension;
  double deltaX, deltaY, deltaZ, Xsize, Ysize, Zsize;
  double simTime;
  std::vector<Var*> vars;  // order = the reference's varInfoList (lookup by id finds the first match)
  std::vector<Geo*> geos;
  std::vector<Rxn*> rxns;      // rInfoList (slow reactions, then rate rules)
  std::vector<Rxn*> fastRxns;  // parsed and never evaluated (calcFastReaction.cxx:1-4)
  std::vector<Var*> orderedAssignmentRules;
  Var *xVar, *yVar, *zVar, *tVar;
  std::string error;

  Model();
  ~Model();
  int64_t ncells() const { return (int64_t)nx * ny * nz; }
  int64_t ndoubled() const { return (int64_t)Xindex * Yindex * Zindex; }
  Var* findVar(const std::string& id) const;
  Geo* geoByCompartment(const std::string& id) const;
  Geo* geoByDomainType(const std::string& id) const;
  // slabLo/slabHi: rows (2-D: y, 3-D: z) [slabLo, slabHi) of the whole grid are held locally
  // (owned rows plus halo); slabLo == slabHi == 0 means the whole grid.
  bool build(SBMLDocument* d, int xdim, int ydim, int zdim, int slabLo = 0, int slabHi = 0);
  bool slab;  // built for a slab of a sharded grid (the rows held may still be the whole grid when the slabs are thin)
  bool isSlab() const { return slab; }
  // value of a coordinate variable at doubled-grid coordinate C of its axis (setInfoFunction.cpp:308): min + C*delta/2.0
  double coordinateAt(const Var* v, int C) const {
    const double delta = v == xVar ? deltaX : v == yVar ? deltaY : deltaZ;
    return v->offGridInit + (double)C * delta / 2.0;
  }
  // what the reference's value[] array of `v` holds at doubled-grid point (X, Y, Z) at set-up time
  double valueAtDoubled(const Var* v, int X, int Y, int Z, int depth = 0) const;
  // reversePolishInitial at a single doubled-grid point; returns false when the stream leaves nothing on the stack
  bool evalRpnAtPoint(const Rpn& rpn, int X, int Y, int Z, double* st, double* out, int depth = 0) const;
  // local cell on a local array edge that is not an edge of the whole grid (its neighbour lives in another slab)
  bool cutEdge(int i, int j, int k) const {
    return (j == 0 && off[1] > 0) || (j == ny - 1 && off[1] + ny < gn[1]) || (k == 0 && off[2] > 0) ||
           (k == nz - 1 && off[2] + nz < gn[2]);
  }

 private:
  void setCompartments();
  void setSpecies();
  bool setParameters();
  bool setVolumeGeometries();
  void setMembraneGeometries();
  bool resolveAssignments();
  void setNormals();
  void setVoronoi();
  void allocateMembraneFields();
  void setBoundaryTypes();
  void setReactions();
  void setRateRules();
};

// which two membrane points every pseudo-membrane point takes the minimum of (storage indices of g)
void pseudoFillPairs(const Model& m, const Geo* g, bool everyClassifiedPoint, std::vector<int32_t>& target, std::vector<int32_t>& a,
                     std::vector<int32_t>& b);

// AST normalisation + post-order emission (astFunction.cpp:18-235)
void rearrangeAst(ASTNode* ast);
void emitRpn(const ASTNode* ast, const Model& m, Rpn& out);
void collectDependence(const ASTNode* ast, const Model& m, std::vector<Var*>& dep);

// Host interpreter used at set-up time only (geometry masks, initial assignments): the
// reference's reversePolishInitial (calcPDE.cpp:21-222) over compact cells.  `cells` == NULL
// evaluates every cell; `mask` (optional) restricts to cells with a non-zero byte.
void evalRpnCompact(const Rpn& rpn, const uint8_t* mask, int64_t ncells, double* out);
// one operator token of the reference interpreter acting on the stack st[0..sp] after the pop (calcPDE.cpp:47-215)
void rpnApplyOp(int astType, double* st, int& sp);

}  // namespace sb2host

#endif  // SB2_HOST_MODEL_H_
