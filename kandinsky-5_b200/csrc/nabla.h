#pragma once
#include "common.h"

namespace k5 {

// NABLA block selection (nablaT_v2, kandinsky/models/utils.py:136-163) and the STA mask
// (fast_sta_nabla, models/utils.py:108-133).
size_t nabla_workspace_floats(int S, int heads);
int nabla_select_launches();
// q: Sq query tokens (a rank's slab on a temporal shard), k: all Sk key tokens; sta rows [sta_row0, sta_row0 + Sq/64).
int nabla_select(const bf16* q, int ldq, int Sq, const bf16* k, int ldk, int Sk, int heads, float P, const uint8_t* sta,
                 int sta_row0, int32_t* kv_count, int32_t* kv_index, float* workspace, float* density_acc, cudaStream_t st);
int sta_mask(int T, int Hb, int Wb, int wT, int wH, int wW, uint8_t* out, cudaStream_t st);

}  // namespace k5
