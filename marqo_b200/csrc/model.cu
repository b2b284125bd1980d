// TEMPORARY: encoder entry points are implemented in the next milestone.
#include "common.cuh"
extern "C" {
#define STUB(...) { return mb::guarded([&] { mb::fail(B200_ERR_UNSUPPORTED, "encoder not built yet"); }); }
int b200_model_create(int, const b200_model_desc*, b200_model**) STUB()
int b200_model_destroy(b200_model*) STUB()
int b200_model_load_tensor(b200_model*, const char*, const float*, int64_t) STUB()
int b200_model_finalize(b200_model*) STUB()
int b200_model_encode_images_u8(b200_model*, const uint8_t*, int, int, int, int, float*) STUB()
int b200_model_encode_images_f32(b200_model*, const float*, int, int, float*) STUB()
int b200_model_encode_tokens(b200_model*, const int32_t*, const int32_t*, int, int, int, float*) STUB()
int b200_model_encode_images_u8_device(b200_model*, const uint8_t*, int, int, int, int, float*, int) STUB()
int b200_model_encode_tokens_device(b200_model*, const int32_t*, const int32_t*, int, int, int, float*, int) STUB()
int b200_model_last_timing(b200_model*, float*, int*) STUB()
}
