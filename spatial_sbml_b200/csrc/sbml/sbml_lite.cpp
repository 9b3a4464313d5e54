// sbml_lite.cpp — reader for the SBML subset described in sbml_lite.h (expat based).
#include "sbml_lite.h"

#include <expat.h>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

// ---------------------------------------------------------------------------------------------
// ASTNode
// ---------------------------------------------------------------------------------------------
ASTNode::ASTNode(ASTNodeType_t type)
    : mType(type), mReal(0), mExponent(0), mInteger(0), mDenominator(1) {}
ASTNode::ASTNode(int type)
    : mType((ASTNodeType_t)type), mReal(0), mExponent(0), mInteger(0), mDenominator(1) {}

ASTNode::~ASTNode() {
  for (size_t i = 0; i < mChildren.size(); ++i) delete mChildren[i];
}

double ASTNode::getReal() const {
  if (mType == AST_REAL_E) return mReal * pow(10.0, (double)mExponent);
  if (mType == AST_RATIONAL) return (double)mInteger / (double)mDenominator;
  return mReal;
}

int ASTNode::removeChild(unsigned int n) {
  if (n >= mChildren.size()) return -1;
  mChildren.erase(mChildren.begin() + n);  // not deleted: ownership passes to the caller
  return 0;
}

ASTNode* ASTNode::deepCopy() const {
  ASTNode* c = new ASTNode(mType);
  c->mReal = mReal; c->mExponent = mExponent; c->mInteger = mInteger;
  c->mDenominator = mDenominator; c->mName = mName;
  for (size_t i = 0; i < mChildren.size(); ++i) c->mChildren.push_back(mChildren[i]->deepCopy());
  return c;
}

// n-ary node → left-nested binary tree: op(a,b,c,d) → op(op(op(a,b),c),d)
void ASTNode::reduceToBinary() {
  unsigned int n = getNumChildren();
  if (n < 3) return;
  ASTNode* op = new ASTNode(mType);
  ASTNode* op2 = new ASTNode(mType);
  op->addChild(getChild(0));
  op->addChild(getChild(1));
  op2->addChild(op);
  for (unsigned int i = 2; i < n; ++i) op2->addChild(getChild(i));
  swapChildren(op2);
  op2->mChildren.clear();  // its former children (ours) now live under this / op
  delete op2;
  reduceToBinary();
}

static void formulaRec(const ASTNode* n, std::ostringstream& os) {
  if (!n) return;
  if (n->isInteger()) { os << n->getInteger(); return; }
  if (n->isReal()) { os << n->getReal(); return; }
  if (n->isName()) { os << (n->getType() == AST_NAME_TIME && !*n->getName() ? "time" : n->getName()); return; }
  if (n->isConstant()) {
    os << (n->getType() == AST_CONSTANT_E ? "exponentiale" : n->getType() == AST_CONSTANT_PI ? "pi"
           : n->getType() == AST_CONSTANT_TRUE ? "true" : "false");
    return;
  }
  if (n->isOperator() && n->getNumChildren() >= 2) {
    os << "(";
    for (unsigned int i = 0; i < n->getNumChildren(); ++i) {
      if (i) os << " " << (char)n->getType() << " ";
      formulaRec(n->getChild(i), os);
    }
    os << ")";
    return;
  }
  if (n->isOperator()) os << (char)n->getType(); else os << "f" << (int)n->getType();
  os << "(";
  for (unsigned int i = 0; i < n->getNumChildren(); ++i) { if (i) os << ", "; formulaRec(n->getChild(i), os); }
  os << ")";
}

char* SBML_formulaToString(const ASTNode* ast) {
  std::ostringstream os;
  formulaRec(ast, os);
  return strdup(os.str().c_str());
}

SBase* Model::getElementBySId(const std::string& id) {
  if (SBase* s = mSpecies.get(id)) return s;
  if (SBase* s = mCompartments.get(id)) return s;
  if (SBase* s = mParameters.get(id)) return s;
  if (SBase* s = mReactions.get(id)) return s;
  return NULL;
}

// ---------------------------------------------------------------------------------------------
// tiny DOM over expat (namespace aware: element/attribute names are "uri|local")
// ---------------------------------------------------------------------------------------------
namespace {

const char* kSpatialNsPrefix = "http://www.sbml.org/sbml/level3/version1/spatial/";
const char* kMathNs = "http://www.w3.org/1998/Math/MathML";

struct XmlNode {
  std::string uri, name, text;
  std::vector<std::pair<std::string, std::string> > attrs;  // (local or "uri|local", value)
  std::vector<XmlNode*> children;
  XmlNode* parent;
  XmlNode() : parent(NULL) {}
  ~XmlNode() { for (size_t i = 0; i < children.size(); ++i) delete children[i]; }

  // attribute by local name; a namespaced attribute (spatial:foo) wins over a bare one
  const std::string* attr(const char* local) const {
    const std::string* bare = NULL;
    for (size_t i = 0; i < attrs.size(); ++i) {
      const std::string& k = attrs[i].first;
      size_t bar = k.rfind('|');
      if (bar == std::string::npos) { if (k == local) bare = &attrs[i].second; }
      else if (k.compare(bar + 1, std::string::npos, local) == 0) return &attrs[i].second;
    }
    return bare;
  }
  // bare (un-namespaced) attribute only
  const std::string* coreAttr(const char* local) const {
    for (size_t i = 0; i < attrs.size(); ++i) if (attrs[i].first == local) return &attrs[i].second;
    return NULL;
  }
  std::string attrStr(const char* local, const char* dflt = "") const {
    const std::string* a = attr(local); return a ? *a : std::string(dflt);
  }
  XmlNode* child(const char* local) const {
    for (size_t i = 0; i < children.size(); ++i) if (children[i]->name == local) return children[i];
    return NULL;
  }
};

struct ParseState {
  XmlNode* root; XmlNode* cur;
  std::map<std::string, std::string> rootNs;  // uri -> prefix, from the outermost element
  int depth;
  ParseState() : root(NULL), cur(NULL), depth(0) {}
};

void splitName(const char* qn, std::string& uri, std::string& local) {
  const char* bar = strrchr(qn, '|');
  if (bar) { uri.assign(qn, bar - qn); local.assign(bar + 1); } else { uri.clear(); local.assign(qn); }
}

void XMLCALL onStart(void* ud, const XML_Char* name, const XML_Char** atts) {
  ParseState* st = (ParseState*)ud;
  XmlNode* n = new XmlNode;
  splitName(name, n->uri, n->name);
  if (n->name == "sep" && st->cur) st->cur->text += ' ';  // <cn>1<sep/>3</cn> → "1 3"
  for (int i = 0; atts[i]; i += 2) n->attrs.push_back(std::make_pair(std::string(atts[i]), std::string(atts[i + 1])));
  n->parent = st->cur;
  if (st->cur) st->cur->children.push_back(n); else st->root = n;
  st->cur = n;
  st->depth++;
}
void XMLCALL onEnd(void* ud, const XML_Char*) {
  ParseState* st = (ParseState*)ud;
  if (st->cur) st->cur = st->cur->parent;
  st->depth--;
}
void XMLCALL onText(void* ud, const XML_Char* s, int len) {
  ParseState* st = (ParseState*)ud;
  if (st->cur) st->cur->text.append(s, len);
}
void XMLCALL onNsStart(void* ud, const XML_Char* prefix, const XML_Char* uri) {
  ParseState* st = (ParseState*)ud;
  if (st->depth == 0 && uri) st->rootNs[uri] = prefix ? prefix : "";
}

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
bool toBool(const std::string& s) { return s == "true" || s == "1"; }
double toDouble(const std::string& s) {
  std::string t = trim(s);
  if (t == "INF") return INFINITY;
  if (t == "-INF") return -INFINITY;
  if (t == "NaN") return NAN;
  return strtod(t.c_str(), NULL);
}

// ---------------------------------------------------------------------------------------------
// MathML → ASTNode
// ---------------------------------------------------------------------------------------------
struct MathName { const char* name; ASTNodeType_t type; };
const MathName kMathOps[] = {
  {"plus", AST_PLUS}, {"minus", AST_MINUS}, {"times", AST_TIMES}, {"divide", AST_DIVIDE},
  {"power", AST_POWER}, {"abs", AST_FUNCTION_ABS}, {"arccos", AST_FUNCTION_ARCCOS},
  {"arccosh", AST_FUNCTION_ARCCOSH}, {"arccot", AST_FUNCTION_ARCCOT}, {"arccoth", AST_FUNCTION_ARCCOTH},
  {"arccsc", AST_FUNCTION_ARCCSC}, {"arccsch", AST_FUNCTION_ARCCSCH}, {"arcsec", AST_FUNCTION_ARCSEC},
  {"arcsech", AST_FUNCTION_ARCSECH}, {"arcsin", AST_FUNCTION_ARCSIN}, {"arcsinh", AST_FUNCTION_ARCSINH},
  {"arctan", AST_FUNCTION_ARCTAN}, {"arctanh", AST_FUNCTION_ARCTANH}, {"ceiling", AST_FUNCTION_CEILING},
  {"cos", AST_FUNCTION_COS}, {"cosh", AST_FUNCTION_COSH}, {"cot", AST_FUNCTION_COT},
  {"coth", AST_FUNCTION_COTH}, {"csc", AST_FUNCTION_CSC}, {"csch", AST_FUNCTION_CSCH},
  {"exp", AST_FUNCTION_EXP}, {"factorial", AST_FUNCTION_FACTORIAL}, {"floor", AST_FUNCTION_FLOOR},
  {"ln", AST_FUNCTION_LN}, {"log", AST_FUNCTION_LOG}, {"root", AST_FUNCTION_ROOT},
  {"sec", AST_FUNCTION_SEC}, {"sech", AST_FUNCTION_SECH}, {"sin", AST_FUNCTION_SIN},
  {"sinh", AST_FUNCTION_SINH}, {"tan", AST_FUNCTION_TAN}, {"tanh", AST_FUNCTION_TANH},
  {"and", AST_LOGICAL_AND}, {"not", AST_LOGICAL_NOT}, {"or", AST_LOGICAL_OR}, {"xor", AST_LOGICAL_XOR},
  {"eq", AST_RELATIONAL_EQ}, {"geq", AST_RELATIONAL_GEQ}, {"gt", AST_RELATIONAL_GT},
  {"leq", AST_RELATIONAL_LEQ}, {"lt", AST_RELATIONAL_LT}, {"neq", AST_RELATIONAL_NEQ},
  {NULL, AST_UNKNOWN}
};

ASTNode* mathToAst(const XmlNode* n);

ASTNode* cnToAst(const XmlNode* n) {
  ASTNode* a = new ASTNode;
  std::string type = n->attrStr("type", "real");
  if (type == "integer") {
    a->setValue((long)strtol(trim(n->text).c_str(), NULL, 10));
  } else if (type == "e-notation" || type == "rational") {
    // text is split around a <sep/> child; expat gave us the concatenation, so re-split on the
    // recorded position: we stored no position, so fall back to whitespace splitting.
    std::istringstream is(n->text);
    std::string p1, p2; is >> p1 >> p2;
    if (type == "e-notation") a->setValue(toDouble(p1), (long)strtol(p2.c_str(), NULL, 10));
    else a->setValue((long)strtol(p1.c_str(), NULL, 10), (long)strtol(p2.c_str(), NULL, 10));
  } else {
    a->setValue(toDouble(n->text));
  }
  return a;
}

ASTNode* applyToAst(const XmlNode* n) {
  if (n->children.empty()) return new ASTNode;
  const XmlNode* head = n->children[0];
  ASTNode* a = new ASTNode;
  size_t firstArg = 1;
  if (head->name == "ci") {  // user function call
    a->setType(AST_FUNCTION);
    a->setName(trim(head->text).c_str());
  } else if (head->name == "csymbol") {
    std::string url = head->attrStr("definitionURL");
    if (url.find("delay") != std::string::npos) { a->setType(AST_FUNCTION_DELAY); a->setName(trim(head->text).c_str()); }
    else { a->setType(AST_FUNCTION); a->setName(trim(head->text).c_str()); }
  } else {
    ASTNodeType_t t = AST_UNKNOWN;
    for (const MathName* m = kMathOps; m->name; ++m) if (head->name == m->name) { t = m->type; break; }
    a->setType(t);
  }
  // qualifiers of root / log become the first child, defaulting to 2 / 10
  if (a->getType() == AST_FUNCTION_ROOT || a->getType() == AST_FUNCTION_LOG) {
    const char* qual = a->getType() == AST_FUNCTION_ROOT ? "degree" : "logbase";
    ASTNode* q = NULL;
    if (n->children.size() > 1 && n->children[1]->name == qual) {
      const XmlNode* qn = n->children[1];
      if (!qn->children.empty()) q = mathToAst(qn->children[0]);
      firstArg = 2;
    }
    if (!q) { q = new ASTNode; q->setValue(a->getType() == AST_FUNCTION_ROOT ? 2 : 10); }
    a->addChild(q);
  }
  for (size_t i = firstArg; i < n->children.size(); ++i) a->addChild(mathToAst(n->children[i]));
  return a;
}

ASTNode* mathToAst(const XmlNode* n) {
  const std::string& e = n->name;
  if (e == "math" || e == "semantics") {
    for (size_t i = 0; i < n->children.size(); ++i) {
      const std::string& c = n->children[i]->name;
      if (c != "annotation" && c != "annotation-xml") return mathToAst(n->children[i]);
    }
    return NULL;
  }
  if (e == "cn") return cnToAst(n);
  if (e == "ci") { ASTNode* a = new ASTNode(AST_NAME); a->setName(trim(n->text).c_str()); return a; }
  if (e == "csymbol") {
    std::string url = n->attrStr("definitionURL");
    ASTNode* a = new ASTNode(url.find("avogadro") != std::string::npos ? AST_NAME_AVOGADRO
                             : url.find("time") != std::string::npos ? AST_NAME_TIME : AST_NAME);
    a->setName(trim(n->text).c_str());
    return a;
  }
  if (e == "exponentiale") return new ASTNode(AST_CONSTANT_E);
  if (e == "pi") return new ASTNode(AST_CONSTANT_PI);
  if (e == "true") return new ASTNode(AST_CONSTANT_TRUE);
  if (e == "false") return new ASTNode(AST_CONSTANT_FALSE);
  if (e == "infinity") { ASTNode* a = new ASTNode; a->setValue(INFINITY); return a; }
  if (e == "notanumber") { ASTNode* a = new ASTNode; a->setValue(NAN); return a; }
  if (e == "apply") return applyToAst(n);
  if (e == "piecewise") {
    ASTNode* a = new ASTNode(AST_FUNCTION_PIECEWISE);
    for (size_t i = 0; i < n->children.size(); ++i) {
      const XmlNode* c = n->children[i];
      for (size_t j = 0; j < c->children.size(); ++j) a->addChild(mathToAst(c->children[j]));
    }
    return a;
  }
  if (e == "lambda") {
    ASTNode* a = new ASTNode(AST_LAMBDA);
    for (size_t i = 0; i < n->children.size(); ++i) {
      const XmlNode* c = n->children[i];
      if (c->name == "bvar") { if (!c->children.empty()) a->addChild(mathToAst(c->children[0])); }
      else a->addChild(mathToAst(c));
    }
    return a;
  }
  return new ASTNode;
}

ASTNode* readMathChild(const XmlNode* parent) {
  const XmlNode* m = parent->child("math");
  return m ? mathToAst(m) : NULL;
}

// ---------------------------------------------------------------------------------------------
// DOM → object model
// ---------------------------------------------------------------------------------------------
CoordinateKind_t parseCoordinateKind(const std::string& s) {
  if (s == "cartesianX") return SPATIAL_COORDINATEKIND_CARTESIAN_X;
  if (s == "cartesianY") return SPATIAL_COORDINATEKIND_CARTESIAN_Y;
  if (s == "cartesianZ") return SPATIAL_COORDINATEKIND_CARTESIAN_Z;
  return SPATIAL_COORDINATEKIND_INVALID;
}
BoundaryKind_t parseBoundaryKind(const std::string& s) {
  if (s == "Neumann") return SPATIAL_BOUNDARYKIND_NEUMANN;
  if (s == "Dirichlet") return SPATIAL_BOUNDARYKIND_DIRICHLET;
  if (s == "Robin_valueCoefficient") return SPATIAL_BOUNDARYKIND_ROBIN_VALUE_COEFFICIENT;
  if (s == "Robin_inwardNormalGradientCoefficient") return SPATIAL_BOUNDARYKIND_ROBIN_INWARD_NORMAL_GRADIENT_COEFFICIENT;
  if (s == "Robin_sum") return SPATIAL_BOUNDARYKIND_ROBIN_SUM;
  return SPATIAL_BOUNDARYKIND_INVALID;
}

void setCoreId(SBase* sb, const XmlNode* n) {
  if (const std::string* id = n->coreAttr("id")) sb->setId(*id);
}
void setSpatialId(SBase* sb, const XmlNode* n) {
  if (const std::string* id = n->attr("id")) sb->setId(*id);
}

Geometry* buildGeometry(const XmlNode* g) {
  Geometry* geo = new Geometry;
  setSpatialId(geo, g);
  if (const XmlNode* l = g->child("listOfCoordinateComponents"))
    for (size_t i = 0; i < l->children.size(); ++i) {
      const XmlNode* c = l->children[i];
      if (c->name != "coordinateComponent") continue;
      CoordinateComponent* cc = new CoordinateComponent;
      setSpatialId(cc, c);
      cc->mType = parseCoordinateKind(c->attrStr("type"));
      if (const XmlNode* b = c->child("boundaryMin")) { setSpatialId(&cc->mMin, b); cc->mMin.mValue = toDouble(b->attrStr("value", "0")); }
      if (const XmlNode* b = c->child("boundaryMax")) { setSpatialId(&cc->mMax, b); cc->mMax.mValue = toDouble(b->attrStr("value", "0")); }
      geo->mCoordinateComponents.append(cc);
    }
  if (const XmlNode* l = g->child("listOfDomainTypes"))
    for (size_t i = 0; i < l->children.size(); ++i) {
      const XmlNode* c = l->children[i];
      if (c->name != "domainType") continue;
      DomainType* dt = new DomainType;
      setSpatialId(dt, c);
      dt->mDims = (unsigned int)toDouble(c->attrStr("spatialDimensions", "3"));
      geo->mDomainTypes.append(dt);
    }
  if (const XmlNode* l = g->child("listOfDomains"))
    for (size_t i = 0; i < l->children.size(); ++i) {
      const XmlNode* c = l->children[i];
      if (c->name != "domain") continue;
      Domain* d = new Domain;
      setSpatialId(d, c);
      d->mDomainType = c->attrStr("domainType");
      geo->mDomains.append(d);
    }
  if (const XmlNode* l = g->child("listOfAdjacentDomains"))
    for (size_t i = 0; i < l->children.size(); ++i) {
      const XmlNode* c = l->children[i];
      if (c->name != "adjacentDomains") continue;
      AdjacentDomains* ad = new AdjacentDomains;
      setSpatialId(ad, c);
      ad->mDomain1 = c->attrStr("domain1");
      ad->mDomain2 = c->attrStr("domain2");
      geo->mAdjacentDomains.append(ad);
    }
  if (const XmlNode* l = g->child("listOfSampledFields"))
    for (size_t i = 0; i < l->children.size(); ++i) {
      const XmlNode* c = l->children[i];
      if (c->name != "sampledField") continue;
      SampledField* sf = new SampledField;
      setSpatialId(sf, c);
      sf->mN1 = (int)toDouble(c->attrStr("numSamples1", "0"));
      sf->mN2 = (int)toDouble(c->attrStr("numSamples2", "0"));
      sf->mN3 = (int)toDouble(c->attrStr("numSamples3", "0"));
      if (c->attrStr("compression", "uncompressed") == "uncompressed") {
        std::istringstream is(c->text); double v;
        while (is >> v) sf->mSamples.push_back(v);
      }
      geo->mSampledFields.append(sf);
    }
  if (const XmlNode* l = g->child("listOfGeometryDefinitions"))
    for (size_t i = 0; i < l->children.size(); ++i) {
      const XmlNode* c = l->children[i];
      if (c->name == "analyticGeometry") {
        AnalyticGeometry* ag = new AnalyticGeometry;
        setSpatialId(ag, c);
        if (const XmlNode* lv = c->child("listOfAnalyticVolumes"))
          for (size_t j = 0; j < lv->children.size(); ++j) {
            const XmlNode* v = lv->children[j];
            if (v->name != "analyticVolume") continue;
            AnalyticVolume* av = new AnalyticVolume;
            setSpatialId(av, v);
            av->mDomainType = v->attrStr("domainType");
            av->mOrdinal = (int)toDouble(v->attrStr("ordinal", "0"));
            av->mMath = readMathChild(v);
            ag->mVolumes.append(av);
          }
        geo->mGeometryDefinitions.append(ag);
      } else if (c->name == "sampledFieldGeometry") {
        SampledFieldGeometry* sg = new SampledFieldGeometry;
        setSpatialId(sg, c);
        sg->mSampledField = c->attrStr("sampledField");
        if (const XmlNode* lv = c->child("listOfSampledVolumes"))
          for (size_t j = 0; j < lv->children.size(); ++j) {
            const XmlNode* v = lv->children[j];
            if (v->name != "sampledVolume") continue;
            SampledVolume* sv = new SampledVolume;
            setSpatialId(sv, v);
            sv->mDomainType = v->attrStr("domainType");
            sv->mSampledValue = toDouble(v->attrStr("sampledValue", "0"));
            sg->mVolumes.append(sv);
          }
        geo->mGeometryDefinitions.append(sg);
      } else if (c->name == "csGeometry") {
        GeometryDefinition* gd = new GeometryDefinition(GeometryDefinition::CSG);
        setSpatialId(gd, c);
        geo->mGeometryDefinitions.append(gd);
      } else if (c->name == "parametricGeometry") {
        GeometryDefinition* gd = new GeometryDefinition(GeometryDefinition::PARAMETRIC);
        setSpatialId(gd, c);
        geo->mGeometryDefinitions.append(gd);
      }
    }
  return geo;
}

void readSpeciesReferences(const XmlNode* list, ListOfT<SpeciesReference>& out) {
  if (!list) return;
  for (size_t i = 0; i < list->children.size(); ++i) {
    const XmlNode* c = list->children[i];
    if (c->name != "speciesReference" && c->name != "modifierSpeciesReference") continue;
    SpeciesReference* sr = new SpeciesReference;
    setCoreId(sr, c);
    sr->mSpecies = c->attrStr("species");
    if (const std::string* st = c->coreAttr("stoichiometry")) { sr->mStoichiometry = toDouble(*st); sr->mStoichiometrySet = true; }
    out.append(sr);
  }
}

Model* buildModel(const XmlNode* m) {
  Model* model = new Model;
  setCoreId(model, m);
  SpatialModelPlugin* mp = new SpatialModelPlugin;
  model->setPlugin(mp);

  for (size_t k = 0; k < m->children.size(); ++k) {
    const XmlNode* l = m->children[k];
    if (l->name == "listOfCompartments") {
      for (size_t i = 0; i < l->children.size(); ++i) {
        const XmlNode* c = l->children[i];
        if (c->name != "compartment") continue;
        Compartment* comp = new Compartment;
        setCoreId(comp, c);
        if (const std::string* a = c->coreAttr("spatialDimensions")) { comp->mDims = (unsigned int)toDouble(*a); comp->mDimsSet = true; }
        if (const std::string* a = c->coreAttr("size")) { comp->mSize = toDouble(*a); comp->mSizeSet = true; }
        if (const std::string* a = c->coreAttr("constant")) comp->mConstant = toBool(*a);
        SpatialCompartmentPlugin* cp = new SpatialCompartmentPlugin;
        if (const XmlNode* cm = c->child("compartmentMapping")) {
          cp->mMapping = new CompartmentMapping;
          setSpatialId(cp->mMapping, cm);
          cp->mMapping->mDomainType = cm->attrStr("domainType");
          cp->mMapping->mUnitSize = toDouble(cm->attrStr("unitSize", "1"));
        }
        comp->setPlugin(cp);
        model->mCompartments.append(comp);
      }
    } else if (l->name == "listOfSpecies") {
      for (size_t i = 0; i < l->children.size(); ++i) {
        const XmlNode* c = l->children[i];
        if (c->name != "species") continue;
        Species* s = new Species;
        setCoreId(s, c);
        s->mCompartment = c->attrStr("compartment");
        if (const std::string* a = c->coreAttr("initialAmount")) { s->mInitialAmount = toDouble(*a); s->mAmountSet = true; }
        if (const std::string* a = c->coreAttr("initialConcentration")) { s->mInitialConcentration = toDouble(*a); s->mConcSet = true; }
        if (const std::string* a = c->coreAttr("constant")) { s->mConstant = toBool(*a); s->mConstantSet = true; }
        if (const std::string* a = c->coreAttr("boundaryCondition")) s->mBoundaryCondition = toBool(*a);
        SpatialSpeciesPlugin* sp = new SpatialSpeciesPlugin;
        if (const std::string* a = c->attr("isSpatial")) sp->mIsSpatial = toBool(*a);
        s->setPlugin(sp);
        model->mSpecies.append(s);
      }
    } else if (l->name == "listOfParameters") {
      for (size_t i = 0; i < l->children.size(); ++i) {
        const XmlNode* c = l->children[i];
        if (c->name != "parameter") continue;
        Parameter* p = new Parameter;
        setCoreId(p, c);
        if (const std::string* a = c->coreAttr("value")) { p->mValue = toDouble(*a); p->mValueSet = true; }
        if (const std::string* a = c->coreAttr("constant")) p->mConstant = toBool(*a);
        SpatialParameterPlugin* pp = new SpatialParameterPlugin;
        if (const XmlNode* x = c->child("spatialSymbolReference")) {
          pp->mSymbolRef = new SpatialSymbolReference;
          setSpatialId(pp->mSymbolRef, x);
          pp->mSymbolRef->mSpatialRef = x->attrStr("spatialRef");
        }
        if (const XmlNode* x = c->child("diffusionCoefficient")) {
          pp->mDiff = new DiffusionCoefficient;
          pp->mDiff->mVariable = x->attrStr("variable");
          if (const std::string* a = x->attr("coordinateReference1")) { pp->mDiff->mCoord1 = parseCoordinateKind(*a); pp->mDiff->mCoord1Set = true; }
        }
        if (const XmlNode* x = c->child("advectionCoefficient")) {
          pp->mAdv = new AdvectionCoefficient;
          pp->mAdv->mVariable = x->attrStr("variable");
          pp->mAdv->mCoord = parseCoordinateKind(x->attrStr("coordinate"));
        }
        if (const XmlNode* x = c->child("boundaryCondition")) {
          pp->mBC = new BoundaryCondition;
          pp->mBC->mVariable = x->attrStr("variable");
          pp->mBC->mCoordinateBoundary = x->attrStr("coordinateBoundary");
          pp->mBC->mType = parseBoundaryKind(x->attrStr("type"));
        }
        p->setPlugin(pp);
        model->mParameters.append(p);
      }
    } else if (l->name == "listOfInitialAssignments") {
      for (size_t i = 0; i < l->children.size(); ++i) {
        const XmlNode* c = l->children[i];
        if (c->name != "initialAssignment") continue;
        InitialAssignment* ia = new InitialAssignment;
        setCoreId(ia, c);
        ia->mSymbol = c->attrStr("symbol");
        ia->mMath = readMathChild(c);
        model->mInitialAssignments.append(ia);
      }
    } else if (l->name == "listOfRules") {
      for (size_t i = 0; i < l->children.size(); ++i) {
        const XmlNode* c = l->children[i];
        Rule* r = NULL;
        if (c->name == "assignmentRule") r = new AssignmentRule;
        else if (c->name == "rateRule") r = new RateRule;
        else if (c->name == "algebraicRule") r = new AlgebraicRule;
        else continue;
        setCoreId(r, c);
        r->mVariable = c->attrStr("variable");
        r->mMath = readMathChild(c);
        model->mRules.append(r);
      }
    } else if (l->name == "listOfReactions") {
      for (size_t i = 0; i < l->children.size(); ++i) {
        const XmlNode* c = l->children[i];
        if (c->name != "reaction") continue;
        Reaction* r = new Reaction;
        setCoreId(r, c);
        if (const std::string* a = c->coreAttr("fast")) r->mFast = toBool(*a);
        if (const std::string* a = c->coreAttr("reversible")) r->mReversible = toBool(*a);
        readSpeciesReferences(c->child("listOfReactants"), r->mReactants);
        readSpeciesReferences(c->child("listOfProducts"), r->mProducts);
        readSpeciesReferences(c->child("listOfModifiers"), r->mModifiers);
        if (const XmlNode* kl = c->child("kineticLaw")) {
          r->mKineticLaw = new KineticLaw;
          r->mKineticLaw->mMath = readMathChild(kl);
          const XmlNode* lp = kl->child("listOfLocalParameters");
          if (!lp) lp = kl->child("listOfParameters");
          if (lp)
            for (size_t j = 0; j < lp->children.size(); ++j) {
              const XmlNode* pn = lp->children[j];
              if (pn->name != "localParameter" && pn->name != "parameter") continue;
              LocalParameter* p = new LocalParameter;
              setCoreId(p, pn);
              if (const std::string* a = pn->coreAttr("value")) { p->mValue = toDouble(*a); p->mValueSet = true; }
              r->mKineticLaw->mLocalParameters.append(p);
            }
        }
        model->mReactions.append(r);
      }
    } else if (l->name == "geometry" && l->uri.compare(0, strlen(kSpatialNsPrefix), kSpatialNsPrefix) == 0) {
      delete mp->mGeometry;
      mp->mGeometry = buildGeometry(l);
    }
  }
  return model;
}

SBMLDocument* parseDocument(const char* data, size_t len) {
  SBMLDocument* doc = new SBMLDocument;
  ParseState st;
  XML_Parser p = XML_ParserCreateNS(NULL, '|');
  XML_SetUserData(p, &st);
  XML_SetElementHandler(p, onStart, onEnd);
  XML_SetCharacterDataHandler(p, onText);
  XML_SetStartNamespaceDeclHandler(p, onNsStart);
  if (XML_Parse(p, data, (int)len, 1) == XML_STATUS_ERROR) {
    std::ostringstream os;
    os << "XML error line " << XML_GetCurrentLineNumber(p) << ": " << XML_ErrorString(XML_GetErrorCode(p));
    doc->mErrors.push_back(os.str());
    XML_ParserFree(p);
    delete st.root;
    return doc;
  }
  XML_ParserFree(p);
  if (!st.root || st.root->name != "sbml") {
    doc->mErrors.push_back("not an SBML document");
    delete st.root;
    return doc;
  }
  for (std::map<std::string, std::string>::iterator it = st.rootNs.begin(); it != st.rootNs.end(); ++it) {
    doc->mNamespaces.mUriToPrefix[it->first] = it->second;
    if (it->first.compare(0, strlen(kSpatialNsPrefix), kSpatialNsPrefix) == 0) doc->mHasSpatial = true;
  }
  // the reference asks for the prefix of the spatial *version1* URI; answer that for whatever
  // spatial version is declared so that model->getPlugin(prefix) resolves
  if (doc->mHasSpatial) {
    std::string pfx;
    for (std::map<std::string, std::string>::iterator it = st.rootNs.begin(); it != st.rootNs.end(); ++it)
      if (it->first.compare(0, strlen(kSpatialNsPrefix), kSpatialNsPrefix) == 0) pfx = it->second;
    doc->mNamespaces.mUriToPrefix[std::string(kSpatialNsPrefix) + "version1"] = pfx;
  }
  doc->mLevel = (unsigned int)toDouble(st.root->attrStr("level", "3"));
  doc->mVersion = (unsigned int)toDouble(st.root->attrStr("version", "1"));
  if (const std::string* r = st.root->attr("required")) doc->mSpatialRequired = toBool(*r);
  if (const XmlNode* m = st.root->child("model")) doc->mModel = buildModel(m);
  else doc->mErrors.push_back("document has no <model>");
  delete st.root;
  (void)kMathNs;
  return doc;
}

}  // namespace

SBMLDocument* readSBMLFromString(const char* xml) {
  if (!xml) { SBMLDocument* d = new SBMLDocument; d->mErrors.push_back("null input"); return d; }
  return parseDocument(xml, strlen(xml));
}

SBMLDocument* readSBMLFromFile(const char* filename) {
  std::ifstream in(filename ? filename : "", std::ios::binary);
  if (!in) {
    SBMLDocument* d = new SBMLDocument;
    d->mErrors.push_back(std::string("cannot open file: ") + (filename ? filename : "(null)"));
    return d;
  }
  std::ostringstream ss; ss << in.rdbuf();
  std::string data = ss.str();
  return parseDocument(data.data(), data.size());
}
