// mz_treewarp.cuh — interface of the tree-warp engine (implementation: mz_treewarp.cu, its own translation unit).
//
// ONE launch per act for any network whose fp32 weights fit shared memory (LunarLander-sized nets: BASELINE configs
// 3 and 4), both policies, both qtransforms, any root mode.  A warp owns 32 / LG trees for the WHOLE act — select,
// recurrent_fn, expand + backup — and never meets a CTA barrier inside the simulation loop (the CTA-resident engine
// spent 68 % of its stall samples on one: profiles/r01_resident_lunar_lines.txt).  Trees live as packed 16-byte
// records in global memory (mz_records.cuh; they are touched by one warp only, so they sit in that SM's L1 / in L2),
// the MLP of a tree runs on its LG lanes out of shared-memory weights.
#pragma once
#include <string>

#include "mz_device.cuh"
#include "mz_resident.cuh"

namespace mz {

struct TreeWarpState {
  bool available = false;
  int max_smem = 0, num_sms = 0, G = 0;
  int lanes = 0;          // MZ_TREEWARP_LANES: lanes per tree (8 / 16 / 32); 0 = choose per launch
  int warps = 0;          // MZ_TREEWARP_WARPS: warps per CTA; 0 = choose per launch
  int noise_levels = 32;  // MZ_TREEWARP_K: tie-break noise levels produced ahead of the search
  int prefetch = 0;       // MZ_TREEWARP_PREFETCH: prefetch the children's records while a level is scored
  void* batched = nullptr;  // state of the per-simulation kernels (throughput mode)
  void* scores = nullptr;   // MZ_TW_CACHED builds: [B][NS + 1][A] float2 selection-score cache of the fused kernel
  size_t scores_bytes = 0;
};

int treewarp_init(TreeWarpState& st, const Net& net, int device, std::string* err);
bool treewarp_supported(const TreeWarpState& st, const Net& net, int B, int num_simulations, int max_depth);
// Trees are kept in `rs`'s record arrays (shared with the CTA-resident engine, so mz_get_tree unpacks them the same
// way).  obs [B,obs_dim], or obs == null and root_emb [B,E] (+ optionally root_logits [B,A] and root_value [B]).
int treewarp_launch(TreeWarpState& st, ResidentState& rs, const Net& net, const float* weights, const Tree& tree,
                    const SearchParams& p, const float* obs, const float* root_emb, const float* root_logits,
                    const float* root_value, const uint8_t* invalid, const float* noise, int32_t* action_out,
                    float* weights_out, float* root_value_out, cudaStream_t stream, int64_t* launches,
                    std::string* err);

// Throughput mode (precision = bf16): the same walks as separate launches per simulation around the tcgen05 recurrent
// kernel.  begin (policy prologue + tree init; `sel5` = 5 x B ints of scratch: parent, action, next, depth, fresh — the
// first two feed the recurrent kernel), then per simulation select -> [recurrent] -> backup, then finish.
int treewarp_batched_begin(TreeWarpState& st, ResidentState& rs, const Tree& tree, const SearchParams& p,
                           const float* root_logits, const float* root_value, const float* root_emb,
                           const uint8_t* invalid, const float* noise, int32_t* sel5, cudaStream_t stream,
                           int64_t* launches, std::string* err);
int treewarp_batched_select(TreeWarpState& st, int sim, cudaStream_t stream, int64_t* launches, std::string* err);
int treewarp_batched_backup(TreeWarpState& st, const float* reward, const float* value, const float* logits,
                            const float* next_emb, cudaStream_t stream, int64_t* launches, std::string* err);
// backup of simulation sim - 1 + selection of simulation sim in one launch (same lanes, records hot in L1)
int treewarp_batched_backup_select(TreeWarpState& st, int sim, const float* reward, const float* value,
                                   const float* logits, const float* next_emb, cudaStream_t stream, int64_t* launches,
                                   std::string* err);
int treewarp_batched_finish(TreeWarpState& st, int32_t* action_out, float* weights_out, cudaStream_t stream,
                            int64_t* launches, std::string* err);
void treewarp_destroy(TreeWarpState& st);

}  // namespace mz
