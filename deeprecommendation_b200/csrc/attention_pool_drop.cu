// Training-mode build of K2 with AttentionNet's inner Dropout (attention_ncf.py:112-117) fused into the score computation: the same source as
// csrc/attention_pool.cu compiled with B200REC_ATT_DROPOUT (own namespace, ONE entry point: b200rec_attention_pool_dropout), so that the
// scoring kernels of the default build are untouched.  Mask = Philox2x32-10 keyed by (seed; rated item, candidate row, hidden-unit group), see
// common.cuh; csrc/attention_pool_bwd.cu regenerates it.
#define B200REC_ATT_DROPOUT 1
#include "attention_pool.cu"
