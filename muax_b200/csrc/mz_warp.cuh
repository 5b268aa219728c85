// mz_warp.cuh — interface of the warp-autonomous, compile-time specialised engine for the stock muax MLP family
// (implementation: mz_warp.cu, its own translation unit).  One launch per act, trees in shared memory, a warp owns
// its trees for the whole act; covers the shapes BASELINE.json's metric is quoted on (see warp_variants()).
#pragma once
#include <string>
#include <vector>

#include "mz_device.cuh"

namespace mz {

struct WarpState {
  bool available = false;  // a compiled shape variant matches the handle's networks
  void* impl = nullptr;    // WarpImpl: packed-layer descriptors + the variant's kernels
  float* packed = nullptr; // row-padded weight blob (device)
  float* noise_table = nullptr;   // producers == 0 only: [B][NS][32] tie-break noise from the pre-pass kernel
  uint32_t* cont_keys = nullptr;  // [B][NS][2]
  size_t noise_capacity = 0;
  int max_smem = 0, num_sms = 0;
  int force_warps = 0;  // MZ_WARP_WARPS: search warps per CTA
  int lanes = 16;       // MZ_WARP_LANES: lanes per tree (8 or 16)
  int producers = 2;    // MZ_WARP_PRODUCERS: noise-producer warps per CTA (0 = pre-pass kernel + table in HBM)
  // sharded act with peer stores: completion flags (mz_set_peer_flags) and the launch's finished-CTA counter
  // (done_counter[0]; done_counter[1] = a peer wait timed out)
  int32_t* flag = nullptr;
  int32_t flag_rank = 0, flag_step = 0;
  uint32_t* done_counter = nullptr;
};

int warp_init(WarpState& st, const Net& net, int device, std::string* err);
void warp_destroy(WarpState& st);
int warp_pack(WarpState& st, const float* raw_weights, cudaStream_t stream, int64_t* launches);
bool warp_supported(const WarpState& st, const SearchParams& p, int B);
// inline_keys != null: the simulate keys travel in the kernel parameters (p.sim_keys is ignored).
// peers: byte offsets from this rank's output buffers to the peers' (mz_set_peer_outputs).
// A one-CTA kernel on `stream` that returns once flags[q] >= step for every q < world.
int warp_peer_wait(WarpState& st, const int32_t* flags, int world, int step, cudaStream_t stream, int64_t* launches);
int warp_launch(WarpState& st, const Tree& out, const SearchParams& p, const SimKeys* inline_keys, const float* obs,
                const uint8_t* invalid, const float* noise, int32_t* action_out, float* weights_out,
                float* root_value_out, const std::vector<int64_t>& peers, bool dump_tree, cudaStream_t stream,
                int64_t* launches, std::string* err);

}  // namespace mz
