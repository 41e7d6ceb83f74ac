// Audio data feed: training batches gathered on the device from a resident cache of ground-truth log-magnitude STFT
// columns.  The reference builds every sample on the host -- NeRAF_dataset.py:89-132
// (RAF: load a wav, STFT the whole file, keep ONE column) and :272-296 (SoundSpaces: np.load a (C, F, T) magnitude
// file, keep one column) -- 2048 times per step through DataLoader workers.  Here the columns are laid out once as
// cache[(rir * T + t), C*F] (time-major per RIR, padding columns already filled the reference's way), so the dataset
// index idx = rir * T + t (get_id_tmp, NeRAF_dataset.py:86-87 / :268-269) IS the cache row, and a batch is one row
// gather plus the pose look-ups: 2 x 4 B per element of (B, C, F), HBM-bound, no host work.
#include "common.cuh"
#include "kernels.h"

namespace neraf {

struct GatherArgs {
  const float* cache; long long n_rows; int T; int CF;
  const double* mic_table; const double* src_table; const double* rot_table;
  const long long* idx; long long B;
  float* data; long long* time_query; long long* audio_idx;
  double* mic; double* src; double* rot;
  int* status;
};

// one block per sample: 128 threads stream the 2 KB column, the first 9 threads copy the poses
__global__ void __launch_bounds__(128) gather_batch_kernel(GatherArgs a) {
  const long long b = blockIdx.x;
  long long row = a.idx[b];
  const bool bad = row < 0 || row >= a.n_rows;
  if (bad) {                                     // reported once, the row reads as dataset index 0
    if (threadIdx.x == 0 && a.status) atomicExch(a.status, 1);
    row = 0;
  }
  const long long rir = row / a.T;
  const float* in = a.cache + row * (long long)a.CF;
  float* out = a.data + b * (long long)a.CF;
  for (int i = threadIdx.x; i < a.CF; i += blockDim.x) out[i] = __ldg(in + i);
  const int t = threadIdx.x;
  if (t < 3) {
    a.mic[b * 3 + t] = a.mic_table[rir * 3 + t];
  } else if (t < 6) {
    a.src[b * 3 + t - 3] = a.src_table[rir * 3 + t - 3];
  } else if (t < 9) {
    a.rot[b * 3 + t - 6] = a.rot_table[rir * 3 + t - 6];
  } else if (t == 9) {
    a.time_query[b] = row - rir * a.T;
    if (a.audio_idx) a.audio_idx[b] = rir;
  }
}

}  // namespace neraf

using namespace neraf;

extern "C" int neraf_gather_batch(const float* cache, int64_t n_rirs, int32_t n_frames, int32_t column_floats,
                                  const double* mic_table, const double* source_table, const double* rot_table,
                                  const int64_t* sample_idx, int64_t batch, float* data, int64_t* time_query,
                                  int64_t* audio_idx, double* mic_pose, double* source_pose, double* rot,
                                  int32_t* status, neraf_stream_t stream) {
  NERAF_REQUIRE(n_rirs >= 0 && n_frames >= 1 && column_floats >= 1 && batch >= 0, "gather_batch: bad sizes");
  if (batch == 0) return NERAF_OK;
  NERAF_REQUIRE(n_rirs > 0, "gather_batch: the cache is empty");
  NERAF_REQUIRE(cache && mic_table && source_table && rot_table && sample_idx && data && time_query && mic_pose &&
                source_pose && rot, "gather_batch: null pointer");
  NERAF_REQUIRE(batch <= 0x7fffffffLL, "gather_batch: batch too large for one launch (%lld)", (long long)batch);
  GatherArgs a;
  a.cache = cache; a.n_rows = (long long)n_rirs * n_frames; a.T = n_frames; a.CF = column_floats;
  a.mic_table = mic_table; a.src_table = source_table; a.rot_table = rot_table;
  a.idx = reinterpret_cast<const long long*>(sample_idx); a.B = batch;
  a.data = data; a.time_query = reinterpret_cast<long long*>(time_query);
  a.audio_idx = reinterpret_cast<long long*>(audio_idx);
  a.mic = mic_pose; a.src = source_pose; a.rot = rot; a.status = status;
  gather_batch_kernel<<<(unsigned)batch, 128, 0, (cudaStream_t)stream>>>(a);
  NERAF_CHECK_LAUNCH("gather_batch_kernel");
  return NERAF_OK;
}
