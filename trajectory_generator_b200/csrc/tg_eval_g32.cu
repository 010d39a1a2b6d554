// M1 evaluation kernels, 32 lanes per problem (see tg_kernels_eval.inc)
#define TG_GS 32
#define TG_SFX _g32
#define TG_INLINE_ALL
#define TG_INLINE_LEAVES 1
#include "tg_kernels_eval.inc"
