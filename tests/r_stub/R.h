/* stand-in for <R.h> (see Rinternals.h in this directory) */
#ifndef CCGP_STUB_R_H
#define CCGP_STUB_R_H
#include <stdlib.h>
#include <stdio.h>
#endif
