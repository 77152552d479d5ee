// Launch choice for a finished new node: the step right after Scheduler.Solve.
//   instance type: the surviving option whose cheapest compatible AVAILABLE offering is cheapest — the order the fake
//     cloud provider sorts by before it takes element 0 (cloudprovider/fake/cloudprovider.go:74-84); ties keep the
//     provider's input order (canonical rule R4: sort.Slice treated as a stable sort);
//   offering: Offerings.Available().Requirements(reqs).Cheapest() (cloudprovider/types.go:128-145); lo.MinBy keeps the
//     first minimum, i.e. Offerings list order.
// Shared by the device kernel and (host instantiation) the CPU tests of the encoder's offering table.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KS_LHD __host__ __device__ __forceinline__
#else
#define KS_LHD inline
#endif

namespace ksched {

constexpr uint64_t kNoOffering = ~0ull;

// min over the offering slots the node's zone / capacity-type requirements admit; keys64 = the type's 64 table entries
KS_LHD uint64_t offering_min_key(const uint64_t* keys64, uint32_t zmask, uint32_t cmask) {
  uint64_t best = kNoOffering;
  for (int ct = 0; ct < 4; ++ct) {
    if (!((cmask >> ct) & 1)) continue;
    for (int z = 0; z < 16; ++z) {
      if (!((zmask >> z) & 1)) continue;
      const uint64_t k = keys64[ct * 16 + z];
      best = k < best ? k : best;
    }
  }
  return best;
}
// ordering key of an option: (price rank of its cheapest compatible offering, provider input order)
KS_LHD uint64_t option_key(uint64_t offering_key, uint32_t input_index) { return ((offering_key >> 16) << 32) | input_index; }

}  // namespace ksched
