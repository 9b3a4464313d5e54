// rpn.cpp — AST normalisation, reverse-polish emission and the set-up-time host interpreter.
//
// Behavioural source: reference SpatialSBML/astFunction.cpp:18-235 (parseAST, rearrangeAST,
// parseDependence) and calcPDE.cpp:21-222 (reversePolishInitial).  The interpreter here runs at
// set-up time only (geometry masks, initial assignments); the per-step evaluation of kinetic laws
// is the device bytecode interpreter in csrc/device/stage_kernel.cu.
#include "host_model.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include <cmath>

#ifndef M_E
#define M_E 2.71828182845904523536
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace sb2host {

namespace {

ASTNode* realNode(double v) {
  ASTNode* n = new ASTNode(AST_REAL);
  n->setValue(v);
  return n;
}
ASTNode* intNode(int v) {
  ASTNode* n = new ASTNode(AST_INTEGER);
  n->setValue(v);
  return n;
}

bool isZeroLiteral(const ASTNode* n) {
  if (!n) return false;
  if (n->isReal()) return fabs(n->getReal()) == 0;
  if (n->isInteger()) return n->getInteger() == 0;
  return false;
}

// piecewise(v0,c0,v1,c1,...[,otherwise]) → otherwiseTerm + (v0*c0 + (v1*c1 + ...))
// with otherwiseTerm = and(not c0, not c1, ...) * otherwise, or the integer 0 when there is no
// otherwise branch (astFunction.cpp:138-183).
void lowerPiecewise(ASTNode* pw) {
  const unsigned int noc = pw->getNumChildren();
  std::vector<ASTNode*> pieces, conditions;
  for (unsigned int j = 0; j < noc / 2; ++j) {
    ASTNode* value = pw->getChild(0);
    ASTNode* cond = pw->getChild(1);
    conditions.push_back(cond->deepCopy());
    ASTNode* product = new ASTNode(AST_TIMES);
    product->addChild(value);
    product->addChild(cond);
    pieces.push_back(product);
    pw->removeChild(0);
    pw->removeChild(0);
  }
  ASTNode* other = new ASTNode;
  if (noc % 2 != 0) {
    ASTNode* expr = pw->getChild(0);
    pw->removeChild(0);
    other->setType(AST_TIMES);
    ASTNode* none = new ASTNode(AST_LOGICAL_AND);
    for (size_t k = 0; k < conditions.size(); ++k) {
      ASTNode* neg = new ASTNode(AST_LOGICAL_NOT);
      neg->addChild(conditions[k]);
      none->addChild(neg);
    }
    other->addChild(none);
    other->addChild(expr);
  } else {
    other->setType(AST_INTEGER);
    other->setValue(0);
    for (size_t k = 0; k < conditions.size(); ++k) delete conditions[k];
  }
  pw->setType(AST_PLUS);
  pw->addChild(other);
  ASTNode* tail = pw;
  for (size_t j = 0; j + 1 < pieces.size(); ++j) {
    tail->addChild(new ASTNode(AST_PLUS));
    tail = tail->getChild(1);
    tail->addChild(pieces[j]);
  }
  if (!pieces.empty()) tail->addChild(pieces.back());
}

}  // namespace

void rearrangeAst(ASTNode* ast) {
  if (!ast) return;
  const int type = (int)ast->getType();  // the type BEFORE rewriting decides the n-ary reduction below
  if (type == AST_MINUS && ast->getNumChildren() == 1) {
    ast->setType(AST_TIMES);  // -a → a * -1.0
    ast->addChild(realNode(-1.0));
  } else if (type == AST_PLUS && ast->getNumChildren() == 1) {
    ast->setType(AST_TIMES);  // +a → a * 1.0
    ast->addChild(realNode(1.0));
  } else if (type == AST_FUNCTION_PIECEWISE) {
    lowerPiecewise(ast);
  } else if (ast->isLogical()) {
    if (type != AST_LOGICAL_NOT) {
      if (ast->getNumChildren() == 1) ast->addChild(intNode(1));
      else ast->reduceToBinary();
    }
  } else if (type == AST_TIMES) {
    if (ast->getNumChildren() > 0 && (isZeroLiteral(ast->getLeftChild()) || isZeroLiteral(ast->getRightChild()))) {
      // 0 * x → 0.0; only the first two children are dropped, as in the reference
      ASTNode* a = ast->getChild(0);
      ast->removeChild(0);
      ASTNode* b = ast->getChild(0);
      ast->removeChild(0);
      delete a;
      delete b;
      ast->setType(AST_REAL);
      ast->setValue(0.0);
    }
  }
  // LIBSBML_VERSION >= 50903 branch (astFunction.cpp:219-224)
  if (type == AST_TIMES || type == AST_PLUS) ast->reduceToBinary();
  for (unsigned int i = 0; i < ast->getNumChildren(); ++i) rearrangeAst(ast->getChild(i));
}

void emitRpn(const ASTNode* ast, const Model& m, Rpn& out) {
  for (unsigned int i = 0; i < ast->getNumChildren(); ++i) emitRpn(ast->getChild(i), m, out);
  Tok t;
  t.kind = Tok::NOP;  // an all-zero slot: the interpreter treats it as an operator that only pops
  t.astType = 0;
  t.var = NULL;
  t.literal = 0.0;
  if (ast->isFunction() || ast->isOperator() || ast->isRelational() || ast->isLogical()) {
    t.kind = Tok::OP;
    t.astType = (int)ast->getType();
  } else if (ast->isReal()) {
    t.kind = Tok::LITERAL;
    t.literal = ast->getReal();
  } else if (ast->isInteger()) {
    t.kind = Tok::LITERAL;
    t.literal = (double)ast->getInteger();
  } else if (ast->isConstant()) {
    t.kind = Tok::LITERAL;
    switch ((int)ast->getType()) {
      case AST_CONSTANT_E: t.literal = M_E; break;
      case AST_CONSTANT_PI: t.literal = M_PI; break;
      case AST_CONSTANT_FALSE: t.literal = 0.0; break;
      case AST_CONSTANT_TRUE: t.literal = 1.0; break;
      default: t.kind = Tok::NOP; break;
    }
  } else if (ast->isName()) {
    const int type = (int)ast->getType();
    if (type == AST_NAME) {
      Var* v = m.findVar(ast->getName());
      if (v && v->hasValue) {
        t.kind = v->isUniform ? Tok::CONSTVAR : Tok::VAR;
        t.var = v;
      }
    } else if (type == AST_NAME_AVOGADRO) {
      t.kind = Tok::LITERAL;
      t.literal = 6.0221367e+23;
    } else if (type == AST_NAME_TIME) {
      if (m.tVar) { t.kind = Tok::CONSTVAR; t.var = m.tVar; }
    }
  }
  out.push_back(t);
}

void collectDependence(const ASTNode* ast, const Model& m, std::vector<Var*>& dep) {
  for (unsigned int i = 0; i < ast->getNumChildren(); ++i) collectDependence(ast->getChild(i), m, dep);
  if (!ast->isName()) return;
  const std::string name = ast->getName();
  for (size_t i = 0; i < dep.size(); ++i) if (dep[i]->id == name) return;
  Var* v = m.findVar(name);
  if (v) dep.push_back(v);
}

// one operator of the reference interpreter (calcPDE.cpp:47-215) acting on the stack after the pop
void rpnApplyOp(int astType, double* st, int& sp) {
  double& a = st[sp > 0 ? sp - 1 : 0];
  const double b = st[sp];
  switch (astType) {
    case AST_PLUS: if (sp > 0) a += b; break;
    case AST_MINUS: if (sp > 0) a -= b; break;
    case AST_TIMES: if (sp > 0) a *= b; break;
    case AST_DIVIDE: if (sp > 0) a /= b; break;
    case AST_POWER:
    case AST_FUNCTION_POWER: if (sp > 0) a = pow(a, b); break;
    case AST_FUNCTION_LOG: if (sp > 0) a = log(b) / log(a); break;
    case AST_LOGICAL_AND: if (sp > 0) a = ((int)a == 1 && (int)(b == 1)) ? 1.0 : 0.0; break;
    case AST_LOGICAL_OR: if (sp > 0) a = ((int)a == 1 || (int)(b == 1)) ? 1.0 : 0.0; break;
    case AST_LOGICAL_XOR:
      if (sp > 0) a = (((int)a == 1 && (int)(b == 0)) || ((int)a == 0 && (int)(b == 1))) ? 1.0 : 0.0;
      break;
    case AST_RELATIONAL_EQ: if (sp > 0) a = (a == b) ? 1.0 : 0.0; break;
    case AST_RELATIONAL_GEQ: if (sp > 0) a = (a >= b) ? 1.0 : 0.0; break;
    case AST_RELATIONAL_GT: if (sp > 0) a = (a > b) ? 1.0 : 0.0; break;
    case AST_RELATIONAL_LEQ: if (sp > 0) a = (a <= b) ? 1.0 : 0.0; break;
    case AST_RELATIONAL_LT: if (sp > 0) a = (a < b) ? 1.0 : 0.0; break;
    case AST_RELATIONAL_NEQ: if (sp > 0) a = (a != b) ? 1.0 : 0.0; break;
    // unary functions rewrite the popped slot and push it back
    case AST_FUNCTION_ABS: st[sp] = fabs(b); ++sp; break;
    case AST_FUNCTION_ARCCOS: st[sp] = acos(b); ++sp; break;
    case AST_FUNCTION_ARCCOSH: st[sp] = acosh(b); ++sp; break;
    case AST_FUNCTION_ARCCSC: st[sp] = asin(1.0 / b); ++sp; break;
    case AST_FUNCTION_ARCSIN: st[sp] = asin(b); ++sp; break;
    case AST_FUNCTION_ARCTAN: st[sp] = atan(b); ++sp; break;
    case AST_FUNCTION_COS: st[sp] = cos(b); ++sp; break;
    case AST_FUNCTION_COSH: st[sp] = cosh(b); ++sp; break;
    case AST_FUNCTION_CSC: st[sp] = 1.0 / sin(b); ++sp; break;
    case AST_FUNCTION_EXP: st[sp] = exp(b); ++sp; break;
    case AST_FUNCTION_LN: st[sp] = log(b); ++sp; break;
    case AST_FUNCTION_ROOT: st[sp] = sqrt(b); ++sp; break;  // sqrt whatever the degree
    case AST_FUNCTION_SIN: st[sp] = sin(b); ++sp; break;
    case AST_FUNCTION_SINH: st[sp] = sinh(b); ++sp; break;
    case AST_FUNCTION_TAN: st[sp] = tan(b); ++sp; break;
    case AST_FUNCTION_TANH: st[sp] = tanh(b); ++sp; break;
    case AST_LOGICAL_NOT: st[sp] = ((int)b == 0) ? 1.0 : 0.0; ++sp; break;
    default: break;  // listed-but-empty operators only pop
  }
}

namespace {
void evalRange(const Rpn& rpn, const uint8_t* mask, int64_t c0, int64_t c1, double* out);

// A stream in which every operator finds its operands on the stack never reads a slot left over from the previous
// cell, so its cells are independent of each other.
bool cellsAreIndependent(const Rpn& rpn) {
  int sp = 0;
  for (size_t i = 0; i < rpn.size(); ++i) {
    const Tok& t = rpn[i];
    if (t.kind == Tok::VAR || t.kind == Tok::CONSTVAR || t.kind == Tok::LITERAL) { if (++sp > 60) return false; continue; }
    if (t.kind != Tok::OP) return false;  // unresolved names: keep the reference's exact stale-slot behaviour
    bool unary = false;
    switch (t.astType) {
      case AST_FUNCTION_ABS: case AST_FUNCTION_ARCCOS: case AST_FUNCTION_ARCCOSH: case AST_FUNCTION_ARCCSC:
      case AST_FUNCTION_ARCSIN: case AST_FUNCTION_ARCTAN: case AST_FUNCTION_COS: case AST_FUNCTION_COSH:
      case AST_FUNCTION_CSC: case AST_FUNCTION_EXP: case AST_FUNCTION_LN: case AST_FUNCTION_ROOT: case AST_FUNCTION_SIN:
      case AST_FUNCTION_SINH: case AST_FUNCTION_TAN: case AST_FUNCTION_TANH: case AST_LOGICAL_NOT: unary = true; break;
      case AST_PLUS: case AST_MINUS: case AST_TIMES: case AST_DIVIDE: case AST_POWER: case AST_FUNCTION_POWER:
      case AST_FUNCTION_LOG: case AST_LOGICAL_AND: case AST_LOGICAL_OR: case AST_LOGICAL_XOR: case AST_RELATIONAL_EQ:
      case AST_RELATIONAL_GEQ: case AST_RELATIONAL_GT: case AST_RELATIONAL_LEQ: case AST_RELATIONAL_LT:
      case AST_RELATIONAL_NEQ: break;
      default: return false;  // listed-but-empty operators
    }
    if (unary) { if (sp < 1) return false; }
    else { if (sp < 2) return false; --sp; }
  }
  return sp >= 1;
}
}  // namespace

// calcPDE.cpp:21-222.  Large grids with a well-formed stream are evaluated by several host threads (set-up of an
// 8192^2 grid spends most of its time here); anything else takes the serial loop, whose stack outlives the cell loop
// exactly as in the reference.
void evalRpnCompact(const Rpn& rpn, const uint8_t* mask, int64_t ncells, double* out) {
  unsigned nthreads = std::thread::hardware_concurrency();
  if (const char* e = getenv("SSB_HOST_THREADS")) nthreads = (unsigned)atoi(e);
  if (nthreads > 32) nthreads = 32;
  if (nthreads > 1 && ncells >= (int64_t)1 << 18 && cellsAreIndependent(rpn)) {
    std::vector<std::thread> pool;
    const int64_t chunk = (ncells + nthreads - 1) / nthreads;
    for (unsigned t = 0; t < nthreads; ++t) {
      const int64_t c0 = (int64_t)t * chunk, c1 = std::min<int64_t>(ncells, c0 + chunk);
      if (c0 >= c1) break;
      pool.emplace_back([&rpn, mask, c0, c1, out] { evalRange(rpn, mask, c0, c1, out); });
    }
    for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
    return;
  }
  evalRange(rpn, mask, 0, ncells, out);
}

namespace {
// The stack outlives the cell loop (stale slots are visible to malformed
// streams exactly as in the reference); a cell is written only when something is left on the stack.
void evalRange(const Rpn& rpn, const uint8_t* mask, int64_t cbegin, int64_t cend, double* out) {
  double st[64] = {0};
  const int n = (int)rpn.size();
  for (int64_t c = cbegin; c < cend; ++c) {
    if (mask && !mask[c]) continue;
    int sp = 0;
    for (int i = 0; i < n; ++i) {
      const Tok& t = rpn[i];
      if (t.kind == Tok::VAR) { st[sp++] = t.var->field[c]; continue; }
      if (t.kind == Tok::CONSTVAR) { st[sp++] = t.var->uniformValue; continue; }
      if (t.kind == Tok::LITERAL) { st[sp++] = t.literal; continue; }
      if (sp > 0) --sp;
      if (t.kind != Tok::OP) continue;
      rpnApplyOp(t.astType, st, sp);
      if (sp > 60) sp = 60;
    }
    if (sp > 0) out[c] = st[--sp];
  }
}
}  // namespace

}  // namespace sb2host
