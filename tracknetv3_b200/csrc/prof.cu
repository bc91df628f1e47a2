#include "prof.cuh"
#include "common.cuh"
#include <vector>

namespace tnb {
namespace {
struct Rec { cudaEvent_t a, b; int kind, n, h, w, cin, cout; };
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
bool g_on = false;
cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool prof_enabled() { return g_on; }

ProfScope::ProfScope(int kind, cudaStream_t s, int n, int h, int w, int cin, int cout) : idx(-1), st(s) {
  if (!g_on) return;
  Rec r{get_event(), get_event(), kind, n, h, w, cin, cout};
  cudaEventRecord(r.a, st);
  idx = (int)g_recs.size();
  g_recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(g_recs[idx].b, st);
}
}  // namespace tnb

extern "C" {
// Turn per-launch timing on (clearing old records) or off.
int tnb_profile_enable(int on) {
  using namespace tnb;
  if (on) {
    for (auto& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
    g_recs.clear();
  }
  g_on = on != 0;
  return 0;
}
// Synchronise and read back up to `max` records: desc[i*6..] = kind,n,h,w,cin,cout ; ms[i] = duration.
int tnb_profile_collect(int max, int* desc, float* ms) {
  using namespace tnb;
  int cnt = 0;
  for (auto& r : g_recs) {
    if (cnt >= max) break;
    if (cudaEventSynchronize(r.b) != cudaSuccess) return -1;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return -1;
    desc[cnt * 6 + 0] = r.kind; desc[cnt * 6 + 1] = r.n; desc[cnt * 6 + 2] = r.h; desc[cnt * 6 + 3] = r.w;
    desc[cnt * 6 + 4] = r.cin; desc[cnt * 6 + 5] = r.cout;
    ms[cnt] = t;
    ++cnt;
  }
  return cnt;
}
}
