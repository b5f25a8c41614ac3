// Deterministic segmented sum of rows by integer key -- the engine behind the
// k-means M-step (calculate_prototypes_from_labels per image,
// hsg/utils/segsort/common.py:11-41), prototype pooling and segment_mean
// (hsg/utils/general/common.py:123-147).
//
// The reference uses scatter_add_ (fp32 atomics, order not reproducible on a
// GPU).  Here rows are first ordered by key with a stable counting sort that
// exploits the segment structure (keys of segment s lie in
// [s*kmax, (s+1)*kmax)), then summed run by run in that fixed order:
//
//   tiles   : segment-aligned blocks of `tile` rows             build_tiles
//   hist    : per tile, count of each local key                 tile_hist
//   scan    : per (segment,key) exclusive scan over tiles       -> positions
//   scatter : stable placement of row ids into `perm`           (1 warp / tile)
//   gather  : each warp sums a run of 64 consecutive perm entries, emitting
//             one partial row ("piece") per (run, key) pair; piece id =
//             run + key is unique and increasing because perm is key-sorted
//   combine : per output row, add its pieces in run order, then finish
//             (normalise / mean / plain sum)
//
// Bytes per row: the row itself once (gather) + ~16 B of keys/perm traffic;
// pieces add dim*4*(N/64 + bins) bytes in total (1.6 % at dim 258).
#pragma once

#include "common.cuh"

namespace hsg {

constexpr int SR_RUN = 64;          // perm entries per gather warp
constexpr int SR_MAX_KEYS = 49152;   // keys per segment (shared-memory histogram / cursors)
constexpr int SR_TILE_MIN = 2048;   // rows per tile (doubles until <= 1024 tiles/segment)

struct Tiles {
  int64_t tile;          // rows per tile
  int64_t bound;         // host upper bound on the number of tiles
  int32_t* seg;          // [bound] segment of each tile
  int64_t* begin;        // [bound]
  int64_t* end;          // [bound]
  int32_t* seg_first;    // [S+1] first tile of each segment
  int32_t* count;        // [1] number of tiles
};

struct SegReducePlan {
  int64_t N;
  int dim;
  int S;
  int kmax;
  int64_t bins;          // S * kmax
  bool exact;            // pieces are int64 fixed-point sums (order / partition independent)
  Tiles tiles;
  int32_t* keys;         // [N] global key = s*kmax + local key
  uint32_t* perm;        // [N] row id relative to its segment start
  uint32_t* tile_hist;   // [bound * kmax]
  int64_t* bin_start;    // [bins] first perm position of each key
  int32_t* bin_count;    // [bins]
  float* pieces;         // [(N/SR_RUN + bins + 2) * dim]
  int32_t* piece_cnt;    // [N/SR_RUN + bins + 2]
};

// Device-side switch: a gated kernel returns at once unless *flag == want (flag NULL = open).
// Lets a stream carry both variants of a step and pick one from a value computed on the
// device, with no host synchronisation.
struct Gate {
  const int32_t* flag;
  int want;
#ifdef __CUDACC__
  __device__ __forceinline__ bool closed() const { return flag && *flag != want; }
#endif
};

int64_t sr_tile_size(int64_t max_seg_len);
int64_t sr_tiles_bound(int64_t N, int S, int64_t tile);
size_t sr_workspace_bytes(int64_t N, int dim, int S, int kmax, int64_t max_seg_len, bool exact = false);
// carve the plan's buffers out of `c`
void sr_carve(Carver& c, SegReducePlan& p, int64_t N, int dim, int S, int kmax, int64_t max_seg_len, bool exact = false);

int sr_build_tiles(const SegReducePlan& p, const int64_t* seg_offsets, cudaStream_t st);
// keys[i] = s*kmax + clamp(labels[i] - base_s, 0, kmax-1); seg_base may be NULL (base 0)
int sr_labels_to_keys(const SegReducePlan& p, const int64_t* labels, const int64_t* seg_base,
                      cudaStream_t st);
// labels[i] = keys[i] - s*kmax (+ base_s)
int sr_keys_to_labels(const SegReducePlan& p, const int32_t* keys, int64_t* labels, cudaStream_t st);
// hist + scan + scatter + gather: after this, pieces/bin_start/bin_count describe the sums
int sr_sort_and_sum(const SegReducePlan& p, const float* x, const int64_t* seg_offsets,
                    cudaStream_t st);
// the same behind a gate; with eoff/erow (kmeans.cuh: DeltaList) the sorted objects are signed
// delta entries (p.keys = entry keys, p.N = entry capacity, p.tiles built from eoff) and the
// pieces are float64
int sr_sort_and_sum_gated(const SegReducePlan& p, const float* x, const int64_t* seg_offsets, Gate gate,
                          const int64_t* eoff, const uint32_t* erow, cudaStream_t st);
// k-means finish on running float64 sums: *delta_flag == 1 adds the (float64) pieces of a delta
// pass, otherwise the sums are set from the float pieces of a full pass; out = normalised rows
int sr_combine64(const SegReducePlan& p, const float* pieces_full, const double* pieces_delta,
                 const int32_t* delta_flag, double* sums, int32_t* members, float* out, cudaStream_t st,
                 Gate gate = Gate{nullptr, 0});
// rows of p whose key differs from keys_prev -> signed entries (erow/ekey, per-segment offsets
// eoff[S+1]); flag[0] = 1 when they fit `cap` entries (else 0: take the full pass), flag[1] =
// changed rows; keys_prev is brought up to date either way
int sr_delta_build(const SegReducePlan& p, const int64_t* seg_offsets, int32_t* keys_prev,
                   int32_t* tile_entries, int64_t* eoff, int32_t* flag, int64_t cap, uint32_t* erow,
                   int32_t* ekey, cudaStream_t st);
// exact mode: out[r,:] = int64 fixed-point (2^-36) sums of the rows of bin r
int sr_combine_exact(const SegReducePlan& p, int64_t P, const int64_t* seg_base, long long* out, cudaStream_t st);
// out[r,:] for r in [0,P): row r = key (seg_base == NULL) or the label whose
// key it is; mode as in hsg_b200.h.  sums_out / counts_out optional.
int sr_combine(const SegReducePlan& p, int64_t P, const int64_t* seg_base, int mode, float* out,
               float* sums_out, float* counts_out, cudaStream_t st);

}  // namespace hsg
