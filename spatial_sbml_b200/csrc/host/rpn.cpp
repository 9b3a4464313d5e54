// rpn.cpp — AST normalisation, reverse-polish emission and the set-up-time host interpreter.
//
// Behavioural source: reference SpatialSBML/astFunction.cpp:18-235 (parseAST, rearrangeAST,
// parseDependence) and calcPDE.cpp:21-222 (reversePolishInitial).  The interpreter here runs at
// set-up time only (geometry masks, initial assignments); the per-step evaluation of kinetic laws
// is the device bytecode interpreter in csrc/device/stage_kernel.cu.
#include "host_model.h"

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

// calcPDE.cpp:21-222.  The stack outlives the cell loop (stale slots are visible to malformed
// streams exactly as in the reference); a cell is written only when something is left on the stack.
void evalRpnCompact(const Rpn& rpn, const uint8_t* mask, int64_t ncells, double* out) {
  double st[64] = {0};
  const int n = (int)rpn.size();
  for (int64_t c = 0; c < ncells; ++c) {
    if (mask && !mask[c]) continue;
    int sp = 0;
    for (int i = 0; i < n; ++i) {
      const Tok& t = rpn[i];
      if (t.kind == Tok::VAR) { st[sp++] = t.var->field[c]; continue; }
      if (t.kind == Tok::CONSTVAR) { st[sp++] = t.var->uniformValue; continue; }
      if (t.kind == Tok::LITERAL) { st[sp++] = t.literal; continue; }
      if (sp > 0) --sp;
      if (t.kind != Tok::OP) continue;
      double& a = st[sp > 0 ? sp - 1 : 0];
      const double b = st[sp];
      switch (t.astType) {
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
      if (sp > 60) sp = 60;
    }
    if (sp > 0) out[c] = st[--sp];
  }
}

}  // namespace sb2host
