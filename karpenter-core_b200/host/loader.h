#pragma once
#include <string>

#include "model.h"

namespace kmodel {
// resource.MustParse -> int64 milli-units (throws std::runtime_error on sub-milli / unknown suffix)
int64_t parse_quantity_milli(const std::string& s);
// Build a Problem from the JSON test format (see tests/README in DESIGN.md). Caller owns the result.
Problem* problem_from_json(const char* text);
}  // namespace kmodel
