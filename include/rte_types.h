/* rte_types.h - C types matching the reference's working precision and logical kind.
 *
 * replaces rte/kernels/api/rte_types.h.in:20-26 (configured by CMake from RTE_ENABLE_SP) and
 * mirrors rte/kernels/mo_rte_kind.F90:24-41:  wp = c_double (default) or c_float (-DRTE_USE_SP,
 * the reference's own preprocessor name), wl = c_bool (1 byte).
 */
#ifndef RTE_TYPES_H
#define RTE_TYPES_H

#include <stdbool.h>

typedef bool Bool;

#ifdef RTE_USE_SP
typedef float Float;
#else
typedef double Float;
#endif

#endif /* RTE_TYPES_H */
