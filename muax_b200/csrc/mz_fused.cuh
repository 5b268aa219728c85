// mz_fused.cuh — fused persistent engine (placeholder until the kernel lands; the stepwise engine is complete).
#pragma once
#include <string>

#include "mz_device.cuh"

namespace mz {

struct FusedState {
  bool available = false;
};

inline int fused_init(FusedState&, const Net&, int, int, int, std::string*) { return 0; }
inline void fused_destroy(FusedState&) {}
inline bool fused_supported(const FusedState& st, const Net&, const SearchParams&) { return st.available; }
inline int fused_launch(FusedState&, const Net&, const float*, const Tree&, const SearchParams&, const float*,
                        const uint8_t*, const float*, int32_t*, float*, float*, cudaStream_t, std::string* err) {
  *err = "fused engine not built";
  return 1;
}

}  // namespace mz
