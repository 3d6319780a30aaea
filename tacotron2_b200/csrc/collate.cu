// Device-side batch collation (data_utils.py:73-111): a ragged batch that already lives in
// HBM -- token ids and (n_mel, L_i) mel spectrograms (e.g. straight out of t2_mel_spectrogram), concatenated, with
// prefix offsets -- becomes the padded, length-sorted 5-tuple Tacotron2.parse_batch expects, without a host round trip:
//     order          rows sorted by decreasing text length (ties: original order)         data_utils.py:80-82
//     text_padded    (B, T_max) int64, zero padded                                         :85-89
//     mel_padded     (B, n_mel, L_pad) fp32, zero padded; gate_padded (B, L_pad) = 1 from the last real frame on   :97-107
//     input_lengths / output_lengths (B) int64                                             :80, :108
#include "common.cuh"

namespace t2 {
namespace {

__global__ void collate_rank_kernel(const int64_t* __restrict__ text_off, int B, int32_t* __restrict__ order) {
  // rank by counting (B <= 1024): position of row i = #rows that are longer, or as long and earlier
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const int64_t li = text_off[i + 1] - text_off[i];
    int r = 0;
    for (int j = 0; j < B; ++j) {
      const int64_t lj = text_off[j + 1] - text_off[j];
      r += (lj > li || (lj == li && j < i)) ? 1 : 0;
    }
    order[r] = i;
  }
}

__global__ void collate_fill_kernel(const int64_t* __restrict__ text_flat, const int64_t* __restrict__ text_off,
                                    const float* __restrict__ mel_flat, const int64_t* __restrict__ mel_off,
                                    const int32_t* __restrict__ order, int B, int n_mel, int T_max, int L_pad,
                                    int64_t* __restrict__ text_padded, int64_t* __restrict__ in_len, float* __restrict__ mel_padded,
                                    float* __restrict__ gate_padded, int64_t* __restrict__ out_len) {
  const int b = blockIdx.y, src = order[b];
  const int64_t t0 = text_off[src], tl = text_off[src + 1] - t0;
  const int64_t m0 = mel_off[src], ml = mel_off[src + 1] - m0;       // frames
  if (blockIdx.x == 0 && threadIdx.x == 0) { in_len[b] = tl; out_len[b] = ml; }
  const long n_text = T_max, n_gate = L_pad, n_mel_el = (long)n_mel * L_pad, total = n_text + n_gate + n_mel_el;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    if (i < n_text) {
      text_padded[(long)b * T_max + i] = i < tl ? text_flat[t0 + i] : 0;
    } else if (i < n_text + n_gate) {
      const long f = i - n_text;
      gate_padded[(long)b * L_pad + f] = f >= ml - 1 ? 1.f : 0.f;                      // data_utils.py:107
    } else {
      const long e = i - n_text - n_gate, c = e / L_pad, f = e - c * L_pad;
      mel_padded[((long)b * n_mel + c) * L_pad + f] = f < ml ? mel_flat[m0 * n_mel + c * ml + f] : 0.f;
    }
  }
}

}  // namespace
}  // namespace t2

extern "C" {

int t2_collate(const T2CollateArgs* a, void* stream) {
  using namespace t2;
  if (!a || !a->text_flat || !a->text_offsets || !a->mel_flat || !a->mel_offsets || !a->order || !a->text_padded ||
      !a->input_lengths || !a->mel_padded || !a->gate_padded || !a->output_lengths)
    return fail(T2_ERR_INVALID, "collate: null argument");
  if (a->B <= 0 || a->B > 1024 || a->n_mel <= 0 || a->T_max <= 0 || a->L_pad <= 0)
    return fail(T2_ERR_INVALID, "collate: 1 <= B <= 1024 and positive sizes (B=%d)", a->B);
  cudaStream_t s = (cudaStream_t)stream;
  collate_rank_kernel<<<1, 256, 0, s>>>(a->text_offsets, a->B, a->order);
  T2_LAUNCH_CHECK();
  const long total = (long)a->T_max + a->L_pad + (long)a->n_mel * a->L_pad;
  const unsigned gx = (unsigned)((total + 255) / 256 > 64 ? 64 : (total + 255) / 256);
  collate_fill_kernel<<<dim3(gx, a->B), 256, 0, s>>>(a->text_flat, a->text_offsets, a->mel_flat, a->mel_offsets, a->order, a->B,
                                                      a->n_mel, a->T_max, a->L_pad, a->text_padded, a->input_lengths, a->mel_padded,
                                                      a->gate_padded, a->output_lengths);
  T2_LAUNCH_CHECK();
  return T2_OK;
}

}  // extern "C"
