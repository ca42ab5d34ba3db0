// Stand-in for <cooperative_groups.h> when the kernels are compiled by the HOST compiler (tests/host_emul/emul.hpp).
// Only what kernels.cuh names; the device-resident loop that uses it is parsed, never run, under emulation.
#pragma once
namespace cooperative_groups {
struct grid_group {
  void sync() const {}
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
