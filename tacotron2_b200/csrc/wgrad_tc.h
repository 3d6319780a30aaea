// Tensor-core engine for the time-batched LSTM weight gradients of the decoder backward (wgrad_tc.cu).
#pragma once
#include "decoder.h"

namespace t2 {

constexpr int kWgSeg = 100;      // decoder steps (K chunks of 64 batch rows) per K split
inline int wgrad_seg(int T) { int seg = kWgSeg; while ((T + seg - 1) / seg > 15) seg += 50; return seg; }
size_t wgrad_tc_ws_bytes(int B, int T);
int wgrad_tc_run(T2Model* m, int B, int T, const float* dga, const float* dgd, const float* x2, const DecoderStash& st,
                 float* const* G, void* ws, size_t ws_bytes, cudaStream_t s);

}  // namespace t2
