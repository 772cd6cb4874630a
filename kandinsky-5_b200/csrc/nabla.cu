// NABLA block selection and STA mask — see nabla.h.  (Implementation lands after the dense path.)
#include "nabla.h"

namespace k5 {

size_t nabla_workspace_floats(int S, int heads) {
    const size_t nb = S / 64;
    return static_cast<size_t>(heads) * nb * nb + 2 * nb * static_cast<size_t>(heads) * 64;
}
int nabla_select_launches() { return 0; }
int nabla_select(const bf16*, int, const bf16*, int, int, int, float, const uint8_t*, int32_t*, int32_t*, float*, float*,
                 cudaStream_t) {
    set_last_error("NABLA block selection is not implemented yet");
    return K5_ERR_UNSUPPORTED;
}
int sta_mask(int, int, int, int, int, int, uint8_t*, cudaStream_t) {
    set_last_error("STA mask is not implemented yet");
    return K5_ERR_UNSUPPORTED;
}

}  // namespace k5
