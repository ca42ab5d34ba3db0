// quad_group.hpp — host-side grouping of hermiteInterpolate samples by the knot interval that completes them, for
// simpson_fused_kernel (quad_kernels.cuh). Pure C++ (shared by quadrature.cu and the host-emulation tests).
#pragma once
#include <vector>

#include "quad_kernels.cuh"

namespace b200rk {

// plan: the samples in the order the reference returns them (HermiteOut::j = knot interval; kind 1 = copy of the LAST
// knot's integral, which the last of the `steps` intervals completes). Output: emit_begin[steps + 1] offsets, the
// samples regrouped by interval (stable within an interval), and for each regrouped sample its position in `plan`.
inline void group_samples_by_interval(const std::vector<HermiteOut>& plan, int steps, std::vector<int>* emit_begin,
                                      std::vector<HermiteOut>* grouped, std::vector<int>* slot) {
  std::vector<int> count(steps + 1, 0);
  auto key = [&](const HermiteOut& p) { return p.kind == 1 ? steps - 1 : p.j; };
  for (const HermiteOut& p : plan) count[key(p) + 1] += 1;
  emit_begin->assign(steps + 1, 0);
  for (int j = 0; j < steps; ++j) (*emit_begin)[j + 1] = (*emit_begin)[j] + count[j + 1];
  grouped->assign(plan.size(), HermiteOut{});
  slot->assign(plan.size(), 0);
  std::vector<int> next(emit_begin->begin(), emit_begin->end() - 1);
  for (size_t o = 0; o < plan.size(); ++o) {
    const int q = next[key(plan[o])]++;
    (*grouped)[q] = plan[o];
    (*slot)[q] = (int)o;
  }
}

}  // namespace b200rk
