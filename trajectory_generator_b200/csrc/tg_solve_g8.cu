// M2 solve kernels, 8 lanes per problem (see tg_kernels_solve.inc)
#define TG_GS 8
#define TG_SFX _g8
#include "tg_kernels_solve.inc"
