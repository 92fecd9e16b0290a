// Temporary: beam decoders not implemented yet (replaced by beam_*.cu).
#include "common.cuh"
namespace ctcvr {
size_t rnnt_beam_state_bytes(const ctcvr_decoder_weights&, int, int, int) { return 0; }
int rnnt_beam_reset(void*, const ctcvr_decoder_weights&, int, int, int, cudaStream_t) { set_error("rnnt_beam: not implemented"); return 3; }
int rnnt_beam_chunk(const ctcvr_decoder_weights&, const float*, int, void*, int, int, int, int, int32_t*, int32_t*,
                    int32_t*, double*, float*, float*, cudaStream_t) { set_error("rnnt_beam: not implemented"); return 3; }
size_t rnnt_prefix_beam_ws_bytes(const ctcvr_decoder_weights&, int, int) { return 0; }
int rnnt_prefix_beam(const ctcvr_decoder_weights&, const float*, const float*, int, int, int, float, float, int32_t*,
                     int32_t*, int32_t*, double*, void*, size_t, cudaStream_t) { set_error("rnnt_prefix_beam: not implemented"); return 3; }
size_t ctc_prefix_beam_ws_bytes(int, int, int, int) { return 0; }
int ctc_prefix_beam(const float*, const int32_t*, int, int, int, int, int, int32_t*, int32_t*, int32_t*, double*,
                    int32_t*, void*, size_t, cudaStream_t) { set_error("ctc_prefix_beam: not implemented"); return 3; }
}
