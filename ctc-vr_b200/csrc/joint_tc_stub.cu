// Temporary: tcgen05 path placeholder (replaced by joint_tc.cu).
#include "common.cuh"
namespace ctcvr {
size_t joint_fwd_tc_ws_bytes(int, int, int, int, int) { return 256; }
int joint_fwd_tc(const float*, const float*, const float*, const float*, const int32_t*, const int32_t*,
                 const int32_t*, float*, float*, float*, int, int, int, int, int, int, void*, size_t, cudaStream_t) { set_error("bf16 path not built"); return 3; }
size_t joint_bwd_tc_ws_bytes(int, int, int, int, int) { return 256; }
int joint_bwd_tc(const float*, const float*, const float*, const float*, const int32_t*, const int32_t*,
                 const int32_t*, const float*, const float*, const float*, const float*, const float*, float, float*,
                 float*, float*, float*, int, int, int, int, int, int, void*, size_t, cudaStream_t) { set_error("bf16 path not built"); return 3; }
}
