/* orc_margin.c -- CPU ORACLE: the marginalization restatement lives in orc_solver.c (orc_marginalize_x0) because it
 * shares the solver's problem container; this translation unit is kept for the build recipe. */
#include "orc_oracle.h"
