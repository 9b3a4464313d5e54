// ORACLE / TEST INFRASTRUCTURE — forwarding header (see ../SBMLTypes.h).
#include <sbml/common/common.h>
