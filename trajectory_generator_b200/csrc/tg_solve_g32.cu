// M2 solve kernels, 32 lanes per problem (see tg_kernels_solve.inc)
#define TG_GS 32
#define TG_SFX _g32
#include "tg_kernels_solve.inc"
