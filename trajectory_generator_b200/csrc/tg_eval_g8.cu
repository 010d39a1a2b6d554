// M1 evaluation kernels, 8 lanes per problem (see tg_kernels_eval.inc)
#define TG_GS 8
#define TG_SFX _g8
#define TG_INLINE_ALL
#define TG_INLINE_LEAVES 1
#include "tg_kernels_eval.inc"
