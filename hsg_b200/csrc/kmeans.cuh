// Internal interfaces of the spherical k-means path (K1).
#pragma once

#include "segreduce.cuh"

namespace hsg {

// list of pixels whose arg-max has to be re-decided in float64
struct FixList {
  int32_t* count;      // [1]
  int32_t* pixels;     // [capacity]
  int32_t* segs;       // [capacity] segment of each listed pixel
  uint16_t* cand;      // [capacity * FIX_MAX_CAND] candidate local ids, 0xFFFF-terminated;
                       // first entry 0xFFFF = "all clusters"; NULL = always all
  int64_t capacity;
};
constexpr int FIX_MAX_CAND = 8;

struct EStepArgs {
  const float* x;            // [N,dim]
  int64_t N;
  int dim;
  const float* centroids;    // [S,kmax,dim]
  const int64_t* seg_offsets;
  int S;
  const int32_t* seg_k;      // may be NULL
  int kmax;
  Tiles tiles;
  int32_t* keys_out;         // [N] global keys
  FixList fix;
};

// fp32 CUDA-core E-step (any shape)
int estep_simt(const EStepArgs& a, cudaStream_t st);
// float64 re-decision of the listed pixels
int estep_fixup(const EStepArgs& a, cudaStream_t st);

// tensor-core (tcgen05) E-step -- tc_estep.cu
struct TcState {
  bool enabled;
  const __half* xh;          // [N,d16+HSG_XH_TAIL]
  const float* xerr;         // [N]
  int d16;
  int kpad;                  // centroids per pass: kmax rounded up to 16, or the tile size when K needs several passes
  int kpad_total, n_pass;    // rows per segment in `ch` = n_pass * kpad
  float* st_val;             // [N,3] / [N,4]: running top-3 between passes (n_pass > 1)
  uint8_t* st_tile;
  __half* ch;                // [S*kpad, d16+HSG_XH_TAIL] fp16 centroids, same row layout as xh
  float* cerr;               // [S*kmax] ||c - fp16(c)|| over the first d16 dims
  float* cerr_max;           // [S]
  unsigned char tmap_x[128]; // CUtensorMap images (host encoded): pixel main slabs / tail slab,
  unsigned char tmap_xt[128];//   centroid main slabs / tail slab
  unsigned char tmap_c[128];
  unsigned char tmap_ct[128];
  unsigned char tmap_c2[128];  // centroid boxes of kpad/2 rows (CTA-pair kernel)
  unsigned char tmap_ct2[128];
};
bool tc_shape_supported(int dim, int d16, int kmax);
void tc_carve(Carver& c, TcState& t, int S, int kmax, int d16, int64_t N);
int tc_prepare(TcState& t, int64_t N, int S);                 // encode tensor maps
int tc_convert_centroids(const EStepArgs& a, const TcState& t, cudaStream_t st);
int estep_tc(const EStepArgs& a, const TcState& t, cudaStream_t st);

}  // namespace hsg
