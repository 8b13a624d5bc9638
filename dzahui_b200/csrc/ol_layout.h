// ol_layout.h — where the CTAs of the one-launch kernel sit along the bar, from where the block scheduler put them.
// Pure host code (no CUDA): used by onelaunch_kernels.cuh's placement probe and compiled on its own by tests/test_layout_cpu.py.
#pragma once
#include <cstddef>
#include <vector>

namespace dzb {

// smid[b] = the SM CTA b ran on in the probe launch (b = blockIdx.x, 0 <= b < grid).
// On success perm[b] = position of CTA b along the bar, a permutation of 0 .. grid-1 such that
//   * the CTAs of one SM hold consecutive positions (so an SM's share is `per` consecutive ranges of the even split:
//     with 16.5 tiles per CTA and two CTAs per SM every SM gets 33 tiles, not 32 or 34);
//   * within an SM the CTA with the LOWEST blockIdx — the one placed first, which wins the issue slots and runs ahead of
//     its neighbour — holds the LAST of the SM's positions (the longer range when ranges alternate between floor and ceil);
//   * SMs are ordered by their first CTA.
// Returns false (perm untouched or partly written: do not use it) unless every SM holds the same number (>= 2) of CTAs —
// the layout by blockIdx is kept then.  Correctness of the solve never depends on the placement being reproduced by later
// launches: any permutation is a valid layout, a different placement only costs the balance.
inline bool onelaunch_layout(const int* smid, int grid, int* perm) {
  if (!smid || !perm || grid <= 0) return false;
  std::vector<int> order;
  std::vector<std::vector<int>> on_sm;
  for (int b = 0; b < grid; ++b) {
    if (smid[b] < 0 || smid[b] > (1 << 20)) return false;
    if ((size_t)smid[b] >= on_sm.size()) on_sm.resize((size_t)smid[b] + 1);
    if (on_sm[smid[b]].empty()) order.push_back(smid[b]);
    on_sm[smid[b]].push_back(b);
  }
  const size_t per = on_sm[order[0]].size();
  if (per < 2 || per * order.size() != (size_t)grid) return false;
  for (int sm : order)
    if (on_sm[sm].size() != per) return false;
  int v = 0;
  for (int sm : order) {
    for (size_t q = 0; q < per; ++q) perm[on_sm[sm][q]] = v + (int)(per - 1 - q);
    v += (int)per;
  }
  return true;
}

}  // namespace dzb
