// see Compadre_GMLS.hpp in this directory (declarations-only stand-in, oracle/_ref only)
#include "Compadre_GMLS.hpp"
