// M2 stage FD (finite-difference mode), 8 lanes per problem: the forward-difference sweeps as a kernel of their own,
// with the evaluators inlined value-only (see tg_kernels_solve.inc)
#define TG_GS 8
#define TG_INLINE_ALL
#define TG_FD_ONLY
#define TG_SFX _g8
#include "tg_kernels_solve.inc"
