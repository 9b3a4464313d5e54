// sbml_lite.h — a small SBML Level 3 (core + spatial 0.9x) object model and reader.
//
// libSBML is not available where this library is built, so the host side of the
// simulator reads SBML through this from-scratch reader over expat.  Class and method names follow
// libSBML's public API for the subset the reference simulator calls (SURVEY.md Appendix D lists the
// call sites: reference SpatialSBML/spatialsimulator.cpp:275-1223, setInfoFunction.cpp:72-494,
// astFunction.cpp:18-235), so that host code reads like code written against libSBML.
//
// MathML → ASTNode mapping mirrors libSBML's reader: <cn> → AST_REAL, <cn type="integer"> →
// AST_INTEGER, n-ary <plus/>/<times/>/<and/>/<or/> kept n-ary, unary <minus/> → one-child
// AST_MINUS, <piecewise> children [v0,c0,v1,c1,...,otherwise], getLeftChild = first child,
// getRightChild = last child, reduceToBinary = left-nested.
#ifndef SBML_LITE_H_
#define SBML_LITE_H_

#include <cfloat>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include <iostream>

#define LIBSBML_VERSION 51800  // >= 50903: the reduceToBinary branch of rearrangeAST is live

// ---------------------------------------------------------------------------------------------
// AST
// ---------------------------------------------------------------------------------------------
typedef enum {
  AST_PLUS = '+', AST_MINUS = '-', AST_TIMES = '*', AST_DIVIDE = '/', AST_POWER = '^',
  AST_INTEGER = 256, AST_REAL, AST_REAL_E, AST_RATIONAL,
  AST_NAME, AST_NAME_AVOGADRO, AST_NAME_TIME,
  AST_CONSTANT_E, AST_CONSTANT_FALSE, AST_CONSTANT_PI, AST_CONSTANT_TRUE,
  AST_LAMBDA,
  AST_FUNCTION,
  AST_FUNCTION_ABS, AST_FUNCTION_ARCCOS, AST_FUNCTION_ARCCOSH, AST_FUNCTION_ARCCOT,
  AST_FUNCTION_ARCCOTH, AST_FUNCTION_ARCCSC, AST_FUNCTION_ARCCSCH, AST_FUNCTION_ARCSEC,
  AST_FUNCTION_ARCSECH, AST_FUNCTION_ARCSIN, AST_FUNCTION_ARCSINH, AST_FUNCTION_ARCTAN,
  AST_FUNCTION_ARCTANH, AST_FUNCTION_CEILING, AST_FUNCTION_COS, AST_FUNCTION_COSH,
  AST_FUNCTION_COT, AST_FUNCTION_COTH, AST_FUNCTION_CSC, AST_FUNCTION_CSCH,
  AST_FUNCTION_DELAY, AST_FUNCTION_EXP, AST_FUNCTION_FACTORIAL, AST_FUNCTION_FLOOR,
  AST_FUNCTION_LN, AST_FUNCTION_LOG, AST_FUNCTION_PIECEWISE, AST_FUNCTION_POWER,
  AST_FUNCTION_ROOT, AST_FUNCTION_SEC, AST_FUNCTION_SECH, AST_FUNCTION_SIN,
  AST_FUNCTION_SINH, AST_FUNCTION_TAN, AST_FUNCTION_TANH,
  AST_LOGICAL_AND, AST_LOGICAL_NOT, AST_LOGICAL_OR, AST_LOGICAL_XOR,
  AST_RELATIONAL_EQ, AST_RELATIONAL_GEQ, AST_RELATIONAL_GT, AST_RELATIONAL_LEQ,
  AST_RELATIONAL_LT, AST_RELATIONAL_NEQ,
  AST_UNKNOWN
} ASTNodeType_t;

class ASTNode {
public:
  explicit ASTNode(ASTNodeType_t type = AST_UNKNOWN);
  explicit ASTNode(int type);
  ~ASTNode();

  ASTNodeType_t getType() const { return mType; }
  void setType(ASTNodeType_t t) { mType = t; }
  void setType(int t) { mType = (ASTNodeType_t)t; }
  void setValue(double v) { mType = AST_REAL; mReal = v; mExponent = 0; }
  void setValue(int v) { mType = AST_INTEGER; mInteger = v; }
  void setValue(long v) { mType = AST_INTEGER; mInteger = v; }
  void setValue(double mantissa, long exponent) { mType = AST_REAL_E; mReal = mantissa; mExponent = exponent; }
  void setValue(long num, long den) { mType = AST_RATIONAL; mInteger = num; mDenominator = den; }
  void setName(const char* n) { mName = n ? n : ""; }
  const char* getName() const { return mName.c_str(); }
  double getReal() const;
  long getInteger() const { return mInteger; }

  unsigned int getNumChildren() const { return (unsigned int)mChildren.size(); }
  ASTNode* getChild(unsigned int n) const { return n < mChildren.size() ? mChildren[n] : NULL; }
  ASTNode* getLeftChild() const { return mChildren.empty() ? NULL : mChildren.front(); }
  ASTNode* getRightChild() const { return mChildren.empty() ? NULL : mChildren.back(); }
  int addChild(ASTNode* c) { mChildren.push_back(c); return 0; }
  int removeChild(unsigned int n);
  void swapChildren(ASTNode* that) { mChildren.swap(that->mChildren); }
  ASTNode* deepCopy() const;
  void reduceToBinary();

  bool isOperator() const {
    return mType == AST_PLUS || mType == AST_MINUS || mType == AST_TIMES || mType == AST_DIVIDE ||
           mType == AST_POWER;
  }
  bool isFunction() const { return mType >= AST_FUNCTION && mType <= AST_FUNCTION_TANH; }
  bool isLogical() const { return mType >= AST_LOGICAL_AND && mType <= AST_LOGICAL_XOR; }
  bool isRelational() const { return mType >= AST_RELATIONAL_EQ && mType <= AST_RELATIONAL_NEQ; }
  bool isReal() const { return mType == AST_REAL || mType == AST_REAL_E || mType == AST_RATIONAL; }
  bool isInteger() const { return mType == AST_INTEGER; }
  bool isConstant() const { return mType >= AST_CONSTANT_E && mType <= AST_CONSTANT_TRUE; }
  bool isName() const { return mType >= AST_NAME && mType <= AST_NAME_TIME; }
  bool isLambda() const { return mType == AST_LAMBDA; }

private:
  ASTNodeType_t mType;
  double mReal;
  long mExponent;
  long mInteger;
  long mDenominator;
  std::string mName;
  std::vector<ASTNode*> mChildren;
  ASTNode(const ASTNode&);
  ASTNode& operator=(const ASTNode&);
};

char* SBML_formulaToString(const ASTNode* ast);  // caller owns (malloc); diagnostic only

// ---------------------------------------------------------------------------------------------
// type codes and spatial enums
// ---------------------------------------------------------------------------------------------
enum {
  SBML_UNKNOWN = 0, SBML_COMPARTMENT = 1, SBML_DOCUMENT = 5, SBML_KINETIC_LAW = 9, SBML_MODEL = 11,
  SBML_PARAMETER = 12, SBML_REACTION = 13, SBML_SPECIES = 15, SBML_SPECIES_REFERENCE = 16,
  SBML_INITIAL_ASSIGNMENT = 21, SBML_ASSIGNMENT_RULE = 22, SBML_RATE_RULE = 23,
  SBML_ALGEBRAIC_RULE = 24, SBML_LOCAL_PARAMETER = 29,
  SBML_SPATIAL_SPATIALSYMBOLREFERENCE = 313, SBML_SPATIAL_DIFFUSIONCOEFFICIENT = 314,
  SBML_SPATIAL_ADVECTIONCOEFFICIENT = 315, SBML_SPATIAL_BOUNDARYCONDITION = 316
};

typedef enum {
  SPATIAL_COORDINATEKIND_CARTESIAN_X = 0,  // numeric values matter: used as array index
  SPATIAL_COORDINATEKIND_CARTESIAN_Y = 1,  // (reference spatialsimulator.cpp:180)
  SPATIAL_COORDINATEKIND_CARTESIAN_Z = 2,
  SPATIAL_COORDINATEKIND_INVALID
} CoordinateKind_t;

typedef enum {
  SPATIAL_BOUNDARYKIND_ROBIN_VALUE_COEFFICIENT,
  SPATIAL_BOUNDARYKIND_ROBIN_INWARD_NORMAL_GRADIENT_COEFFICIENT,
  SPATIAL_BOUNDARYKIND_ROBIN_SUM,
  SPATIAL_BOUNDARYKIND_NEUMANN,
  SPATIAL_BOUNDARYKIND_DIRICHLET,
  SPATIAL_BOUNDARYKIND_INVALID
} BoundaryKind_t;

// ---------------------------------------------------------------------------------------------
// core
// ---------------------------------------------------------------------------------------------
class SBasePlugin {
public:
  virtual ~SBasePlugin() {}
};

class SBase {
public:
  SBase() : mPlugin(NULL) {}
  virtual ~SBase() { delete mPlugin; }
  virtual int getTypeCode() const { return SBML_UNKNOWN; }
  const std::string& getId() const { return mId; }
  bool isSetId() const { return !mId.empty(); }
  void setId(const std::string& id) { mId = id; }
  // Any package name / prefix / namespace URI returns the (single) spatial plugin.
  SBasePlugin* getPlugin(const std::string&) { return mPlugin; }
  const SBasePlugin* getPlugin(const std::string&) const { return mPlugin; }
  void setPlugin(SBasePlugin* p) { delete mPlugin; mPlugin = p; }
protected:
  std::string mId;
  SBasePlugin* mPlugin;
};

template <class T> class ListOfT {
public:
  ~ListOfT() { for (size_t i = 0; i < mItems.size(); ++i) delete mItems[i]; }
  T* get(unsigned int i) const { return i < mItems.size() ? mItems[i] : NULL; }
  T* get(const std::string& id) const {
    for (size_t i = 0; i < mItems.size(); ++i) if (mItems[i]->getId() == id) return mItems[i];
    return NULL;
  }
  unsigned int size() const { return (unsigned int)mItems.size(); }
  void append(T* t) { mItems.push_back(t); }
private:
  std::vector<T*> mItems;
};

class Compartment : public SBase {
public:
  Compartment() : mDims(3), mDimsSet(false), mSize(0), mSizeSet(false), mConstant(true) {}
  int getTypeCode() const { return SBML_COMPARTMENT; }
  unsigned int getSpatialDimensions() const { return mDims; }
  bool isSetSize() const { return mSizeSet; }
  double getSize() const { return mSize; }
  unsigned int mDims; bool mDimsSet; double mSize; bool mSizeSet; bool mConstant;
};

class Species : public SBase {
public:
  Species() : mInitialAmount(0), mInitialConcentration(0), mAmountSet(false), mConcSet(false),
              mConstant(false), mConstantSet(false), mBoundaryCondition(false) {}
  int getTypeCode() const { return SBML_SPECIES; }
  const std::string& getCompartment() const { return mCompartment; }
  bool isSetInitialAmount() const { return mAmountSet; }
  bool isSetInitialConcentration() const { return mConcSet; }
  double getInitialAmount() const { return mInitialAmount; }
  double getInitialConcentration() const { return mInitialConcentration; }
  bool isSetConstant() const { return mConstantSet; }
  bool getConstant() const { return mConstant; }
  bool getBoundaryCondition() const { return mBoundaryCondition; }
  std::string mCompartment;
  double mInitialAmount, mInitialConcentration;
  bool mAmountSet, mConcSet, mConstant, mConstantSet, mBoundaryCondition;
};

class Parameter : public SBase {
public:
  Parameter() : mValue(0), mValueSet(false), mConstant(true) {}
  int getTypeCode() const { return SBML_PARAMETER; }
  bool isSetValue() const { return mValueSet; }
  double getValue() const { return mValue; }
  int setValue(double v) { mValue = v; mValueSet = true; return 0; }
  double mValue; bool mValueSet; bool mConstant;
};

class LocalParameter : public Parameter {
public:
  int getTypeCode() const { return SBML_LOCAL_PARAMETER; }
};

class SpeciesReference : public SBase {
public:
  SpeciesReference() : mStoichiometry(1.0), mStoichiometrySet(false) {}
  int getTypeCode() const { return SBML_SPECIES_REFERENCE; }
  const std::string& getSpecies() const { return mSpecies; }
  double getStoichiometry() const { return mStoichiometry; }
  std::string mSpecies; double mStoichiometry; bool mStoichiometrySet;
};

class KineticLaw : public SBase {
public:
  KineticLaw() : mMath(NULL) {}
  ~KineticLaw() { delete mMath; }
  int getTypeCode() const { return SBML_KINETIC_LAW; }
  const ASTNode* getMath() const { return mMath; }
  bool isSetMath() const { return mMath != NULL; }
  unsigned int getNumLocalParameters() const { return mLocalParameters.size(); }
  const LocalParameter* getLocalParameter(unsigned int i) const { return mLocalParameters.get(i); }
  unsigned int getNumParameters() const { return mLocalParameters.size(); }
  ASTNode* mMath;
  ListOfT<LocalParameter> mLocalParameters;
};

class Reaction : public SBase {
public:
  Reaction() : mKineticLaw(NULL), mFast(false), mReversible(true) {}
  ~Reaction() { delete mKineticLaw; }
  int getTypeCode() const { return SBML_REACTION; }
  const KineticLaw* getKineticLaw() const { return mKineticLaw; }
  KineticLaw* getKineticLaw() { return mKineticLaw; }
  unsigned int getNumReactants() const { return mReactants.size(); }
  unsigned int getNumProducts() const { return mProducts.size(); }
  unsigned int getNumModifiers() const { return mModifiers.size(); }
  SpeciesReference* getReactant(unsigned int i) const { return mReactants.get(i); }
  SpeciesReference* getProduct(unsigned int i) const { return mProducts.get(i); }
  SpeciesReference* getModifier(unsigned int i) const { return mModifiers.get(i); }
  bool getFast() const { return mFast; }
  KineticLaw* mKineticLaw; bool mFast; bool mReversible;
  ListOfT<SpeciesReference> mReactants, mProducts, mModifiers;
};

class Rule : public SBase {
public:
  explicit Rule(int tc) : mMath(NULL), mTypeCode(tc) {}
  ~Rule() { delete mMath; }
  int getTypeCode() const { return mTypeCode; }
  bool isRate() const { return mTypeCode == SBML_RATE_RULE; }
  bool isAssignment() const { return mTypeCode == SBML_ASSIGNMENT_RULE; }
  bool isAlgebraic() const { return mTypeCode == SBML_ALGEBRAIC_RULE; }
  const std::string& getVariable() const { return mVariable; }
  const ASTNode* getMath() const { return mMath; }
  bool isSetMath() const { return mMath != NULL; }
  std::string mVariable; ASTNode* mMath; int mTypeCode;
};
class RateRule : public Rule { public: RateRule() : Rule(SBML_RATE_RULE) {} };
class AssignmentRule : public Rule { public: AssignmentRule() : Rule(SBML_ASSIGNMENT_RULE) {} };
class AlgebraicRule : public Rule { public: AlgebraicRule() : Rule(SBML_ALGEBRAIC_RULE) {} };

class InitialAssignment : public SBase {
public:
  InitialAssignment() : mMath(NULL) {}
  ~InitialAssignment() { delete mMath; }
  int getTypeCode() const { return SBML_INITIAL_ASSIGNMENT; }
  const std::string& getSymbol() const { return mSymbol; }
  const ASTNode* getMath() const { return mMath; }
  std::string mSymbol; ASTNode* mMath;
};

class ListOfSpecies : public ListOfT<Species> {};
class ListOfCompartments : public ListOfT<Compartment> {};
class ListOfReactions : public ListOfT<Reaction> {};
class ListOfParameters : public ListOfT<Parameter> {};
class ListOfRules : public ListOfT<Rule> {
public:
  // rules are looked up by the variable they set, not by id
  Rule* getByVariable(const std::string& v) const {
    for (unsigned int i = 0; i < size(); ++i) if (get(i)->getVariable() == v) return get(i);
    return NULL;
  }
};

class Model : public SBase {
public:
  int getTypeCode() const { return SBML_MODEL; }
  ListOfSpecies* getListOfSpecies() { return &mSpecies; }
  ListOfCompartments* getListOfCompartments() { return &mCompartments; }
  ListOfReactions* getListOfReactions() { return &mReactions; }
  ListOfParameters* getListOfParameters() { return &mParameters; }
  ListOfRules* getListOfRules() { return &mRules; }
  unsigned int getNumSpecies() const { return mSpecies.size(); }
  unsigned int getNumCompartments() const { return mCompartments.size(); }
  unsigned int getNumReactions() const { return mReactions.size(); }
  unsigned int getNumParameters() const { return mParameters.size(); }
  unsigned int getNumRules() const { return mRules.size(); }
  unsigned int getNumInitialAssignments() const { return mInitialAssignments.size(); }
  Species* getSpecies(unsigned int i) const { return mSpecies.get(i); }
  Species* getSpecies(const std::string& id) const { return mSpecies.get(id); }
  Compartment* getCompartment(unsigned int i) const { return mCompartments.get(i); }
  Compartment* getCompartment(const std::string& id) const { return mCompartments.get(id); }
  Parameter* getParameter(unsigned int i) const { return mParameters.get(i); }
  Parameter* getParameter(const std::string& id) const { return mParameters.get(id); }
  Reaction* getReaction(unsigned int i) const { return mReactions.get(i); }
  Reaction* getReaction(const std::string& id) const { return mReactions.get(id); }
  Rule* getRule(unsigned int i) const { return mRules.get(i); }
  Rule* getRule(const std::string& variable) const { return mRules.getByVariable(variable); }
  InitialAssignment* getInitialAssignment(unsigned int i) const { return mInitialAssignments.get(i); }
  InitialAssignment* getInitialAssignment(const std::string& symbol) const {
    for (unsigned int i = 0; i < mInitialAssignments.size(); ++i)
      if (mInitialAssignments.get(i)->getSymbol() == symbol) return mInitialAssignments.get(i);
    return NULL;
  }
  SBase* getElementBySId(const std::string& id);
  ListOfSpecies mSpecies; ListOfCompartments mCompartments; ListOfReactions mReactions;
  ListOfParameters mParameters; ListOfRules mRules; ListOfT<InitialAssignment> mInitialAssignments;
};

class XMLNamespaces {
public:
  // prefix bound to the given namespace URI on the <sbml> element ("" if none)
  std::string getPrefix(const std::string& uri) const {
    std::map<std::string, std::string>::const_iterator it = mUriToPrefix.find(uri);
    return it == mUriToPrefix.end() ? std::string() : it->second;
  }
  std::map<std::string, std::string> mUriToPrefix;
};

class SBMLDocument : public SBase {
public:
  SBMLDocument() : mModel(NULL), mLevel(3), mVersion(1), mSpatialRequired(false), mHasSpatial(false) {}
  ~SBMLDocument() { delete mModel; }
  int getTypeCode() const { return SBML_DOCUMENT; }
  Model* getModel() { return mModel; }
  const Model* getModel() const { return mModel; }
  XMLNamespaces* getNamespaces() { return &mNamespaces; }
  unsigned int getNumErrors() const { return (unsigned int)mErrors.size(); }
  void printErrors(std::ostream& os = std::cerr) const {
    for (size_t i = 0; i < mErrors.size(); ++i) os << mErrors[i] << std::endl;
  }
  bool getPkgRequired(const std::string&) const { return mSpatialRequired; }
  bool isPackageEnabled(const std::string&) const { return mHasSpatial; }
  Model* mModel; XMLNamespaces mNamespaces; unsigned int mLevel, mVersion;
  bool mSpatialRequired, mHasSpatial; std::vector<std::string> mErrors;
};

SBMLDocument* readSBMLFromFile(const char* filename);
SBMLDocument* readSBMLFromString(const char* xml);
inline SBMLDocument* readSBML(const char* filename) { return readSBMLFromFile(filename); }

// ---------------------------------------------------------------------------------------------
// spatial package (spec ~0.90-0.93 as written by the reference's example models)
// ---------------------------------------------------------------------------------------------
class Boundary : public SBase {
public:
  Boundary() : mValue(0) {}
  double getValue() const { return mValue; }
  double mValue;
};

class CoordinateComponent : public SBase {
public:
  CoordinateComponent() : mType(SPATIAL_COORDINATEKIND_INVALID) {}
  CoordinateKind_t getType() const { return mType; }
  Boundary* getBoundaryMin() { return &mMin; }
  Boundary* getBoundaryMax() { return &mMax; }
  const Boundary* getBoundaryMin() const { return &mMin; }
  const Boundary* getBoundaryMax() const { return &mMax; }
  CoordinateKind_t mType; Boundary mMin, mMax;
};
class ListOfCoordinateComponents : public ListOfT<CoordinateComponent> {};

class DomainType : public SBase {
public:
  DomainType() : mDims(3) {}
  unsigned int getSpatialDimensions() const { return mDims; }
  unsigned int mDims;
};

class Domain : public SBase {
public:
  const std::string& getDomainType() const { return mDomainType; }
  std::string mDomainType;
};

class AdjacentDomains : public SBase {
public:
  const std::string& getDomain1() const { return mDomain1; }
  const std::string& getDomain2() const { return mDomain2; }
  std::string mDomain1, mDomain2;
};

class GeometryDefinition : public SBase {
public:
  enum Kind { ANALYTIC, SAMPLED_FIELD, CSG, PARAMETRIC };
  explicit GeometryDefinition(Kind k) : mKind(k) {}
  bool isAnalyticGeometry() const { return mKind == ANALYTIC; }
  bool isSampledFieldGeometry() const { return mKind == SAMPLED_FIELD; }
  bool isCSGeometry() const { return mKind == CSG; }
  bool isParametricGeometry() const { return mKind == PARAMETRIC; }
  Kind mKind;
};

class AnalyticVolume : public SBase {
public:
  AnalyticVolume() : mOrdinal(0), mMath(NULL) {}
  ~AnalyticVolume() { delete mMath; }
  const std::string& getDomainType() const { return mDomainType; }
  int getOrdinal() const { return mOrdinal; }
  int setOrdinal(int o) { mOrdinal = o; return 0; }
  const ASTNode* getMath() const { return mMath; }
  std::string mDomainType; int mOrdinal; ASTNode* mMath;
};

class AnalyticGeometry : public GeometryDefinition {
public:
  AnalyticGeometry() : GeometryDefinition(ANALYTIC) {}
  unsigned int getNumAnalyticVolumes() const { return mVolumes.size(); }
  AnalyticVolume* getAnalyticVolume(unsigned int i) const { return mVolumes.get(i); }
  ListOfT<AnalyticVolume> mVolumes;
};

class SampledVolume : public SBase {
public:
  SampledVolume() : mSampledValue(0) {}
  const std::string& getDomainType() const { return mDomainType; }
  double getSampledValue() const { return mSampledValue; }
  std::string mDomainType; double mSampledValue;
};

class SampledField : public SBase {
public:
  SampledField() : mN1(0), mN2(0), mN3(0) {}
  int getNumSamples1() const { return mN1; }
  int getNumSamples2() const { return mN2; }
  int getNumSamples3() const { return mN3; }
  unsigned int getUncompressedLength() const { return (unsigned int)mSamples.size(); }
  void getUncompressed(double* out) const { for (size_t i = 0; i < mSamples.size(); ++i) out[i] = mSamples[i]; }
  int mN1, mN2, mN3; std::vector<double> mSamples;  // uncompressed samples only
};

class SampledFieldGeometry : public GeometryDefinition {
public:
  SampledFieldGeometry() : GeometryDefinition(SAMPLED_FIELD) {}
  const std::string& getSampledField() const { return mSampledField; }
  unsigned int getNumSampledVolumes() const { return mVolumes.size(); }
  SampledVolume* getSampledVolume(unsigned int i) const { return mVolumes.get(i); }
  std::string mSampledField; ListOfT<SampledVolume> mVolumes;
};

class Geometry : public SBase {
public:
  unsigned int getNumCoordinateComponents() const { return mCoordinateComponents.size(); }
  CoordinateComponent* getCoordinateComponent(unsigned int i) const { return mCoordinateComponents.get(i); }
  CoordinateComponent* getCoordinateComponent(const std::string& id) const { return mCoordinateComponents.get(id); }
  ListOfCoordinateComponents* getListOfCoordinateComponents() { return &mCoordinateComponents; }
  unsigned int getNumGeometryDefinitions() const { return mGeometryDefinitions.size(); }
  GeometryDefinition* getGeometryDefinition(unsigned int i) const { return mGeometryDefinitions.get(i); }
  unsigned int getNumDomainTypes() const { return mDomainTypes.size(); }
  DomainType* getDomainType(unsigned int i) const { return mDomainTypes.get(i); }
  DomainType* getDomainType(const std::string& id) const { return mDomainTypes.get(id); }
  unsigned int getNumDomains() const { return mDomains.size(); }
  Domain* getDomain(unsigned int i) const { return mDomains.get(i); }
  Domain* getDomain(const std::string& id) const { return mDomains.get(id); }
  unsigned int getNumAdjacentDomains() const { return mAdjacentDomains.size(); }
  AdjacentDomains* getAdjacentDomains(unsigned int i) const { return mAdjacentDomains.get(i); }
  unsigned int getNumSampledFields() const { return mSampledFields.size(); }
  SampledField* getSampledField(unsigned int i) const { return mSampledFields.get(i); }
  SampledField* getSampledField(const std::string& id) const { return mSampledFields.get(id); }
  ListOfCoordinateComponents mCoordinateComponents;
  ListOfT<GeometryDefinition> mGeometryDefinitions;
  ListOfT<DomainType> mDomainTypes;
  ListOfT<Domain> mDomains;
  ListOfT<AdjacentDomains> mAdjacentDomains;
  ListOfT<SampledField> mSampledFields;
};

class SpatialModelPlugin : public SBasePlugin {
public:
  SpatialModelPlugin() : mGeometry(NULL) {}
  ~SpatialModelPlugin() { delete mGeometry; }
  Geometry* getGeometry() { return mGeometry; }
  const Geometry* getGeometry() const { return mGeometry; }
  Geometry* mGeometry;
};

class CompartmentMapping : public SBase {
public:
  CompartmentMapping() : mUnitSize(1.0) {}
  const std::string& getDomainType() const { return mDomainType; }
  double getUnitSize() const { return mUnitSize; }
  std::string mDomainType; double mUnitSize;
};

class SpatialCompartmentPlugin : public SBasePlugin {
public:
  SpatialCompartmentPlugin() : mMapping(NULL) {}
  ~SpatialCompartmentPlugin() { delete mMapping; }
  CompartmentMapping* getCompartmentMapping() { return mMapping; }
  const CompartmentMapping* getCompartmentMapping() const { return mMapping; }
  bool isSetCompartmentMapping() const { return mMapping != NULL; }
  CompartmentMapping* mMapping;
};

class SpatialSpeciesPlugin : public SBasePlugin {
public:
  SpatialSpeciesPlugin() : mIsSpatial(false) {}
  bool getIsSpatial() const { return mIsSpatial; }
  bool mIsSpatial;
};

class SpatialSymbolReference : public SBase {
public:
  const std::string& getSpatialRef() const { return mSpatialRef; }
  std::string mSpatialRef;
};

class DiffusionCoefficient : public SBase {
public:
  DiffusionCoefficient() : mCoord1(SPATIAL_COORDINATEKIND_INVALID), mCoord1Set(false) {}
  const std::string& getVariable() const { return mVariable; }
  bool isSetCoordinateReference1() const { return mCoord1Set; }
  CoordinateKind_t getCoordinateReference1() const { return mCoord1; }
  std::string mVariable; CoordinateKind_t mCoord1; bool mCoord1Set;
};

class AdvectionCoefficient : public SBase {
public:
  AdvectionCoefficient() : mCoord(SPATIAL_COORDINATEKIND_INVALID) {}
  const std::string& getVariable() const { return mVariable; }
  CoordinateKind_t getCoordinate() const { return mCoord; }
  std::string mVariable; CoordinateKind_t mCoord;
};

class BoundaryCondition : public SBase {
public:
  BoundaryCondition() : mType(SPATIAL_BOUNDARYKIND_INVALID) {}
  const std::string& getVariable() const { return mVariable; }
  const std::string& getCoordinateBoundary() const { return mCoordinateBoundary; }
  BoundaryKind_t getType() const { return mType; }
  std::string mVariable, mCoordinateBoundary; BoundaryKind_t mType;
};

class SpatialParameterPlugin : public SBasePlugin {
public:
  SpatialParameterPlugin() : mSymbolRef(NULL), mDiff(NULL), mAdv(NULL), mBC(NULL) {}
  ~SpatialParameterPlugin() { delete mSymbolRef; delete mDiff; delete mAdv; delete mBC; }
  SpatialSymbolReference* getSpatialSymbolReference() const { return mSymbolRef; }
  DiffusionCoefficient* getDiffusionCoefficient() const { return mDiff; }
  AdvectionCoefficient* getAdvectionCoefficient() const { return mAdv; }
  BoundaryCondition* getBoundaryCondition() const { return mBC; }
  int getType() const {
    if (mSymbolRef) return SBML_SPATIAL_SPATIALSYMBOLREFERENCE;
    if (mDiff) return SBML_SPATIAL_DIFFUSIONCOEFFICIENT;
    if (mAdv) return SBML_SPATIAL_ADVECTIONCOEFFICIENT;
    if (mBC) return SBML_SPATIAL_BOUNDARYCONDITION;
    return -1;
  }
  bool isSpatialParameter() const { return getType() != -1; }
  SpatialSymbolReference* mSymbolRef; DiffusionCoefficient* mDiff; AdvectionCoefficient* mAdv;
  BoundaryCondition* mBC;
};

#endif  // SBML_LITE_H_
