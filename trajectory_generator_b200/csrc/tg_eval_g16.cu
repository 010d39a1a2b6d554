// M1 evaluation kernels, 16 lanes per problem (see tg_kernels_eval.inc)
#define TG_GS 16
#define TG_SFX _g16
#define TG_INLINE_ALL
#define TG_INLINE_LEAVES 1
// fixed shapes (tg_shape.h) with instantiations in this translation unit
#define TG_EVAL_FIXED TG_FIXED_CASE(TG_FIX_C3)
#include "tg_kernels_eval.inc"
