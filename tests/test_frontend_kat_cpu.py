"""Known-answer tests of the SBML / MathML front end (csrc/sbml/sbml_lite.cpp) and of the AST normalisation + reverse-polish
emission above it (csrc/host/rpn.cpp) that do NOT go through oracle/_ref: the MathML below is written by hand and the expected
token streams are derived by hand from the reference's rules (astFunction.cpp:118-235: unary minus -> a * -1.0, unary plus ->
a * 1.0, piecewise -> PLUS(AND(NOT c...) * otherwise | integer 0, PLUS(v0 * c0, ...)), one-child logical gets an integer 1
sibling, n-ary logical / plus / times -> left-nested binary, 0 * x -> 0.0; parseAST :18-107 emits post-order) and from libSBML's
MathML mapping (SURVEY.md "Shim-fidelity risks").  The goldens of tests/golden were produced through the same reader the product
uses, so a misparse there would be invisible; here it is not.

Token text: "C # v" literal, "C id v" uniform parameter, "V id 1" field, "O n" operator with libSBML's ASTNodeType_t value n."""
import numpy as np
import pytest

import spatial_sbml_b200 as sb
from spatial_sbml_b200 import synthetic

PLUS, MINUS, TIMES, DIVIDE, POWER = 43, 45, 42, 47, 94
EXP, LN, AND, NOT, OR, XOR, EQ, GEQ, GT, LEQ, LT, NEQ = 290, 293, 304, 305, 306, 307, 308, 309, 310, 311, 312, 313
M = '<math xmlns="http://www.w3.org/1998/Math/MathML">%s</math>'
S1, S2 = "<ci> species_1 </ci>", "<ci> species_2 </ci>"


def with_law(law_mathml, extra_params=""):
    """The synthetic 8 x 6 Turing document with the kinetic law of reaction_2 replaced"""
    doc = synthetic.turing(8, 6)
    at = doc.index('id="reaction_2"')
    old = doc[doc.index("<kineticLaw>", at):doc.index("</kineticLaw>", at)]
    doc = doc.replace(old, "<kineticLaw>" + M % law_mathml)
    return doc.replace("<listOfParameters>", "<listOfParameters>" + extra_params)


def stream(doc):
    sim = sb.SpatialSimulator(host_only=True)
    sim.initFromString(doc, 8, 6, 1)
    out = sim.getSetupText("rxn/1").strip().split("\n")
    sim.close()
    return out


def C(v, name="#"):
    return "C %s %s" % (name, v)


def V(name):
    return "V %s 1" % name


def O(n):
    return "O %d" % n


COND = ('<apply><and/><apply><gt/>%s<cn type="e-notation"> 1 <sep/> -3 </cn></apply><apply><geq/>%s<cn> 3 </cn></apply>'
        '<apply><lt/>%s<cn> 50 </cn></apply></apply>') % (S1, S2, S1)
COND_RPN = [V("species_1"), C("0.001"), O(GT), V("species_2"), C("3"), O(GEQ), O(AND), V("species_1"), C("50"), O(LT), O(AND)]


def test_three_term_sum_with_unary_minus_piecewise_and_integer_division():
    law = ('<apply><plus/><apply><times/><ci> kk </ci><apply><minus/>%s</apply></apply>'
           '<piecewise><piece><cn type="integer"> 2 </cn>%s</piece><otherwise><cn> 0.25 </cn></otherwise></piecewise>'
           '<apply><divide/>%s<cn type="integer"> 4 </cn></apply></apply>') % (S1, COND, S2)
    want = ([C("0.75", "kk"), V("species_1"), C("-1"), O(TIMES), O(TIMES)]                        # kk * (species_1 * -1.0)
            + COND_RPN + [O(NOT), C("1"), O(AND), C("0.25"), O(TIMES)]                            # AND(NOT cond, 1) * otherwise
            + [C("2")] + COND_RPN + [O(TIMES), O(PLUS)]                                            # + 2 * cond
            + [O(PLUS), V("species_2"), C("4"), O(DIVIDE), O(PLUS)])                               # ((A + B) + C), left-nested
    assert stream(with_law(law, '<parameter id="kk" value="0.75" constant="true"/>')) == want


def test_zero_times_x_collapses_to_a_real_zero():
    assert stream(with_law("<apply><times/><cn> 0 </cn>%s</apply>" % S1)) == [C("0")]
    assert stream(with_law('<apply><times/>%s<cn type="integer"> 0 </cn></apply>' % S1)) == [C("0")]


def test_piecewise_without_otherwise_gets_an_integer_zero():
    law = ("<piecewise><piece>%s<apply><lt/>%s<cn> 1 </cn></apply></piece><piece>%s<apply><gt/>%s<cn> 2 </cn></apply></piece></piecewise>"
           % (S1, S1, S2, S2))
    want = [C("0"), V("species_1"), V("species_1"), C("1"), O(LT), O(TIMES), V("species_2"), V("species_2"), C("2"), O(GT), O(TIMES),
            O(PLUS), O(PLUS)]
    assert stream(with_law(law)) == want


def test_power_functions_unary_plus_and_binary_minus():
    assert stream(with_law('<apply><power/>%s<cn type="integer"> 2 </cn></apply>' % S1)) == [V("species_1"), C("2"), O(POWER)]
    assert stream(with_law("<apply><exp/><apply><minus/>%s%s</apply></apply>" % (S1, S2))) == [V("species_1"), V("species_2"), O(MINUS), O(EXP)]
    assert stream(with_law("<apply><ln/><apply><plus/>%s</apply></apply>" % S1)) == [V("species_1"), C("1"), O(TIMES), O(LN)]


def test_one_child_logical_gets_an_integer_one_sibling():
    law = "<apply><times/>%s<apply><or/><apply><neq/>%s<cn> 1 </cn></apply></apply></apply>" % (S2, S1)
    assert stream(with_law(law)) == [V("species_2"), V("species_1"), C("1"), O(NEQ), C("1"), O(OR), O(TIMES)]


def test_n_ary_times_is_left_nested_and_numbers_parse():
    law = ('<apply><times/><cn type="rational"> 1 <sep/> 4 </cn>%s<cn type="e-notation"> 2.5 <sep/> 2 </cn>%s<cn> 1e-2 </cn></apply>' % (S1, S2))
    want = [C("0.25"), V("species_1"), O(TIMES), C("250"), O(TIMES), V("species_2"), O(TIMES), C("0.01"), O(TIMES)]
    assert stream(with_law(law)) == want


def test_relational_and_xor_codes():
    law = ("<apply><xor/><apply><eq/>%s%s</apply><apply><leq/>%s<cn> 2 </cn></apply></apply>" % (S1, S2, S1))
    assert stream(with_law(law)) == [V("species_1"), V("species_2"), O(EQ), V("species_1"), C("2"), O(LEQ), O(XOR)]


def test_initial_assignment_is_evaluated_like_the_formula_says():
    """A hand-written initial assignment, evaluated by the host (reversePolishInitial) at the cell centres x = i * dx, y = j * dy:
    piecewise(2.5 + x * 1e-1, and(x >= 2, y < 4), -(y) + 7) inside the box 1 <= x <= 6, 1 <= y <= 4 of the synthetic document."""
    doc = synthetic.turing(8, 6)
    at = doc.index('<initialAssignment symbol="species_1">')
    old = doc[at:doc.index("</initialAssignment>", at)]
    ia = ('<piecewise><piece><apply><plus/><cn> 2.5 </cn><apply><times/><ci> x </ci><cn type="e-notation"> 1 <sep/> -1 </cn></apply></apply>'
          '<apply><and/><apply><geq/><ci> x </ci><cn type="integer"> 2 </cn></apply><apply><lt/><ci> y </ci><cn> 4 </cn></apply></apply></piece>'
          '<otherwise><apply><plus/><apply><minus/><ci> y </ci></apply><cn> 7 </cn></apply></otherwise></piecewise>')
    doc = doc.replace(old, '<initialAssignment symbol="species_1">' + M % ia)
    sim = sb.SpatialSimulator(host_only=True)
    sim.initFromString(doc, 8, 6, 1)
    got = sim.getVariableCompact("species_1")[0]
    sim.close()
    y, x = np.meshgrid(np.arange(6.0), np.arange(8.0), indexing="ij")
    inside = (x >= 1) & (x <= 6) & (y >= 1) & (y <= 4)
    want = np.where((x >= 2) & (y < 4), 2.5 + x * 0.1, 7.0 - y) * inside
    assert np.array_equal(got, want)
