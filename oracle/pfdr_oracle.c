/* TEST INFRASTRUCTURE ONLY (see pfdr_oracle_impl.h).  Instantiates the type-generic oracle
 * body for float (suffix _f32) and double (suffix _f64).  Build: make -C oracle  */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define REAL float
#define REAL_EPS FLT_EPSILON
#define REAL_HUGE HUGE_VALF
#define FN(name) name##_f32
#include "pfdr_oracle_impl.h"
#include "cp_reduce_oracle_impl.h"
#undef REAL
#undef REAL_EPS
#undef REAL_HUGE
#undef FN

#define REAL double
#define REAL_EPS DBL_EPSILON
#define REAL_HUGE HUGE_VAL
#define FN(name) name##_f64
#include "pfdr_oracle_impl.h"
#include "cp_reduce_oracle_impl.h"
