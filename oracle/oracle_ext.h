/* ORACLE - TEST INFRASTRUCTURE ONLY.  Oracle-private switches (not part of the product ABI). */
#ifndef ORACLE_EXT_H
#define ORACLE_EXT_H
#include "rte_types.h"
#include "rrtmgp_b200_ext.h"
#ifdef __cplusplus
extern "C" {
#endif
/* 0 (default) = reference default-kernel quirk; 1 = per-g-point level source (accel behaviour) */
void oracle_set_lw_2stream_lev_source_per_gpt(int on);
#ifdef __cplusplus
}
#endif
#endif
