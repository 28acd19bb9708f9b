/* stand-in for R.h (test infrastructure, see Rinternals.h in this directory) */
#include "Rinternals.h"
