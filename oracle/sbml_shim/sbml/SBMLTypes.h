// ORACLE / TEST INFRASTRUCTURE — not product code.
// Forwarding header: lets the reference's SpatialSBML/*.cpp (which #include <sbml/SBMLTypes.h>)
// compile UNMODIFIED against the repo's small SBML object model, because libSBML itself is not
// installed in this image (oracle plan A).
#ifndef ORACLE_SHIM_SBMLTYPES_H_
#define ORACLE_SHIM_SBMLTYPES_H_
#include <sbml/common/common.h>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#include <iostream>
#include "../../../spatial_sbml_b200/csrc/sbml/sbml_lite.h"
#endif
