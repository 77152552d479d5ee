#pragma once
#include <cstdint>
#include <memory>

#include "model.h"

namespace kmodel {
// BASELINE.json configs 1..5. n_nodes only used by config 5. Caller owns the result.
Problem* synth_problem(int config, int64_t n_pods, int64_t n_types, uint64_t seed, int64_t n_nodes);
}  // namespace kmodel
