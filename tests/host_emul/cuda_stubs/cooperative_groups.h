// Stand-in for <cooperative_groups.h> when the kernels are compiled by the HOST compiler (tests/host_emul/emul.hpp).
// Only what kernels.cuh names. grid.sync() is forwarded to the harness: a no-op in the serial emulation (the
// device-resident loop is then only parsed), the block barrier in the threaded one (which runs it as a one-block grid).
#pragma once
void emul_grid_sync();
namespace cooperative_groups {
struct grid_group {
  void sync() const { ::emul_grid_sync(); }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
