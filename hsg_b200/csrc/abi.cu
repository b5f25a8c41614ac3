// Error plumbing and device queries behind include/hsg_b200.h.
#include "common.cuh"

#include <atomic>
#include <mutex>

namespace hsg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cache[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

// ---------------------------------------------------------------- launch counter + profiler
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

constexpr int PROF_MAX = 8192;
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static bool g_prof_coarse = false;   // hsg_profile_enable(2): one range per library call, not per kernel group
static int g_prof_n = 0;
static int g_prof_phase[PROF_MAX];
static cudaEvent_t g_prof_ev[PROF_MAX][2];
static bool g_prof_created[PROF_MAX];

static thread_local int g_prof_suppress = 0;
ProfSuppress::ProfSuppress() { ++g_prof_suppress; }
ProfSuppress::~ProfSuppress() { --g_prof_suppress; }

ProfRange::ProfRange(int phase, cudaStream_t stream) : slot(-1), st(stream) {
  if (!g_prof_on || g_prof_suppress) return;
  if (g_prof_coarse && ((phase >= PROF_MSTEP_SORT && phase <= PROF_ESTEP_FIXUP) || phase == PROF_CONVERT)) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_prof_on || g_prof_n >= PROF_MAX) return;
  slot = g_prof_n++;
  g_prof_phase[slot] = phase;
  if (!g_prof_created[slot]) {
    cudaEventCreate(&g_prof_ev[slot][0]);
    cudaEventCreate(&g_prof_ev[slot][1]);
    g_prof_created[slot] = true;
  }
  cudaEventRecord(g_prof_ev[slot][0], st);
}

ProfRange::~ProfRange() {
  if (slot >= 0) cudaEventRecord(g_prof_ev[slot][1], st);
}

}  // namespace hsg

extern "C" {

long long hsg_launch_count(void) { return hsg::g_launches.load(); }

int hsg_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(hsg::g_prof_mu);
  hsg::g_prof_on = on != 0;
  hsg::g_prof_coarse = on == 2;
  if (on) hsg::g_prof_n = 0;
  return HSG_OK;
}

// sums the recorded ranges per phase (call after synchronising the stream)
int hsg_profile_collect(double* total_ms, long long* counts, int n_phases) {
  std::lock_guard<std::mutex> lock(hsg::g_prof_mu);
  for (int i = 0; i < n_phases; ++i) { total_ms[i] = 0.0; counts[i] = 0; }
  for (int i = 0; i < hsg::g_prof_n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, hsg::g_prof_ev[i][0], hsg::g_prof_ev[i][1]) != cudaSuccess) continue;
    const int ph = hsg::g_prof_phase[i];
    if (ph < n_phases) { total_ms[ph] += ms; counts[ph] += 1; }
  }
  hsg::g_prof_n = 0;
  return HSG_OK;
}

const char* hsg_last_error(void) { return hsg::g_err; }

int hsg_version(void) { return 100; }

int hsg_device_sms(void) {
  int dev = 0, n = 0;
  HSG_CUDA(cudaGetDevice(&dev));
  HSG_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

}  // extern "C"
