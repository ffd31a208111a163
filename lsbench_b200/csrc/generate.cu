// generate.cu -- synthetic operators on the device.  STUB for milestone 1.
#include "common.cuh"
extern "C" int b200_mat_generate(b200_ctx *c, int kind, uint64_t size,
                                 uint64_t seed, uint32_t flags, b200_mat **M) {
  (void)c, (void)kind, (void)size, (void)seed, (void)flags, (void)M;
  B_FAIL(B200_EINVAL, "b200_mat_generate: not built yet");
}
