// oracle/parallel.h — TEST INFRASTRUCTURE (CPU oracle): pieces shared by compaction.cc and parallel.cc
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../include/b200c.h"
namespace oracle {
struct Unsupported { std::string what; };
struct Corrupt { int input; int kind; uint64_t chunk; uint64_t offset; std::string what; };
struct RangeOut { std::vector<uint8_t> ustream, index; uint64_t partitions = 0, rows = 0; };
int compact_impl(const b200c_manifest* m, b200c_result* res, RangeOut* ro);
}
