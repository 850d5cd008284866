// cupss.h -- umbrella header of the B200-native cuPSS drop-in.
// Same include path and the same public classes as the reference's inc/cupss.h
// (/root/reference/inc/cupss.h:1-6), so example solvers compile unchanged.
#ifndef CUPSS_B200_UMBRELLA_H
#define CUPSS_B200_UMBRELLA_H
#include "cupss/defines.h"
#include "cupss/cu_utils.h"
#include "cupss/evolver.h"
#include "cupss/field.h"
#include "cupss/term.h"
#include "cupss/parser.h"
#endif
