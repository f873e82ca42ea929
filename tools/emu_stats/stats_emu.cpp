// Development tool (not product, not a test): the CPU-fibre emulation of the one-warp ADMM kernel body (tests/emu/warp_emu.cpp)
// with the adaptation statistics hook of admm_warp.cuh enabled.  Built and driven by tools/adapt_stats.py.
#include <cstdio>
#include <vector>
static std::vector<int> g_stats;
static inline void qpc_adapt_stats(int iter, int big, int cnt) {
  g_stats.push_back(iter);
  g_stats.push_back(big);
  g_stats.push_back(cnt);
}
#define QPC_WARP_ADAPT_STATS 1
#include "../../tests/emu/warp_emu.cpp"
extern "C" int emu_adapt_stats(int* out, int cap) {
  int n = (int)g_stats.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = g_stats[i];
  g_stats.clear();
  return n;
}
