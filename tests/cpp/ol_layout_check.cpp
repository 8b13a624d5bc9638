// Reads "grid smid[0] ... smid[grid-1]" from stdin, prints "ok perm[0] ... perm[grid-1]" or "none" (dzb::onelaunch_layout).
#include <cstdio>
#include <vector>

#include "../../dzahui_b200/csrc/ol_layout.h"

int main() {
  int grid = 0;
  if (std::scanf("%d", &grid) != 1 || grid < 0 || grid > (1 << 20)) return 2;
  std::vector<int> smid(grid), perm(grid, -1);
  for (int b = 0; b < grid; ++b)
    if (std::scanf("%d", &smid[b]) != 1) return 2;
  if (!dzb::onelaunch_layout(grid ? smid.data() : nullptr, grid, perm.data())) {
    std::puts("none");
    return 0;
  }
  std::printf("ok");
  for (int b = 0; b < grid; ++b) std::printf(" %d", perm[b]);
  std::puts("");
  return 0;
}
