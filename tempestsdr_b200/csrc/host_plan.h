/* host_plan.h -- data-independent host-side planning (plain C, compiled by gcc with -ffp-contract=off).
 * The functions are part of the public C-ABI and are declared in include/tsdrgpu.h:
 *   tsdrgpu_plan_resample, tsdrgpu_geometry, tsdrgpu_gauss_taps. */
#ifndef TSDRGPU_HOST_PLAN_H_
#define TSDRGPU_HOST_PLAN_H_
#include "../../include/tsdrgpu.h"
#endif
