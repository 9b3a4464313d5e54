// ORACLE / TEST INFRASTRUCTURE — forwarding header (see sbml/SBMLTypes.h in this shim).
#include <sbml/SBMLTypes.h>
