// mz_resident.cuh — interface of the CTA-resident engine (implementation: mz_resident.cu, its own translation unit).
//
// ONE launch per act for shapes whose trees do not fit shared memory.  The stepwise engine pays three launches
// per simulation and the shared-memory engines need the whole tree of every tree they own on chip (C3: 78 KB per
// tree, C5: 75 KB).  Trees are independent, so nothing forces a grid-wide step: here a CTA owns T trees for the
// whole act, the trees stay in the handle's SoA arrays in HBM (a CTA's rows are touched by that CTA only, so they
// live in its SM's L1 / in L2 between visits), the weights are staged once into shared memory by a TMA bulk copy
// when they fit (else read through the read-only path, L2 resident), and the loop
//   select -> recurrent_fn -> expand + backup
// runs num_simulations times with CTA barriers only.  All B trees are in flight at once.
#pragma once
#include <string>

#include "mz_device.cuh"

namespace mz {

struct ResidentState {
  bool available = false;
  int max_smem = 0, num_sms = 0, G = 0;
  int threads = 256;
  int trees_per_cta = 0;         // 0 = choose per launch (MZ_RESIDENT_TREES)
  int force_global_weights = 0;  // MZ_RESIDENT_GLOBAL_WEIGHTS
  int cluster_size = 1;          // CTAs per cluster sharing the streamed weights by TMA multicast (MZ_RESIDENT_CLUSTER);
                                 // measured on B200 (C5 shapes): 1 -> 7.6 ms, 4 -> 9.1 ms, 8 -> 17.5 ms per act (lockstep
                                 // of the cluster costs more than the L2 traffic it saves), so clusters are opt-in
  int no_tma_ring = 0;           // MZ_RESIDENT_NO_TMA: read streamed weights with plain loads (debug / A-B)
  int noise_levels = 8;          // tie-break noise levels produced ahead of the search (MZ_RESIDENT_K); measured
                                 // on B200: 0 / 8 / 32 levels give the same search time, the pre-pass costs 0.85 ms at 32
  float* noise_table = nullptr;  // [B][NS][K][A]
  uint32_t* cont_keys = nullptr; // [B][NS][2] carried key after K levels
  size_t noise_capacity = 0, cont_capacity = 0;
  void* rec_nodes = nullptr;     // [B][NS + 1] 16-byte node records of the last search
  void* rec_childs = nullptr;    // [B][NS + 1][A] 16-byte child records
  float* rec_logits = nullptr;   // [B][NS + 1][A] prior logits
  size_t rec_capacity = 0;       // in nodes
  uint32_t* path = nullptr;      // [B][path slots] selected edges of the simulation in flight
  size_t path_capacity = 0;
  bool dirty = false;            // the SoA tree view is stale: resident_unpack() refreshes it
  cudaStream_t last_stream = nullptr;
  int last_num_sims = 0;
};

int resident_init(ResidentState& st, const Net& net, int device, std::string* err);
void resident_destroy(ResidentState& st);
bool resident_supported(const ResidentState& st, const Net& net, int B, int num_simulations);
// obs [B,obs_dim], or obs == null and root_emb [B,E] (+ optionally root_logits [B,A] and root_value [B]).
int resident_launch(ResidentState& st, const Net& net, const float* weights, const Tree& tree, const SearchParams& p,
                    const float* obs, const float* root_emb, const float* root_logits, const float* root_value,
                    const uint8_t* invalid, const float* noise, int32_t* action_out, float* weights_out,
                    float* root_value_out, cudaStream_t stream, int64_t* launches, std::string* err);

// Shared with the tree-warp engine (mz_treewarp.cu), which keeps its trees in the same record arrays:
int net_weight_bytes(const Net& net);  // bytes of the raw fp32 blob the stacks address, rounded up to 16
int records_reserve(ResidentState& st, int B, int NS, int A, int PL, std::string* err);
int records_noise_prepass(ResidentState& st, const SearchParams& p, int B, int A, int levels, int PL,
                          cudaStream_t stream, int64_t* launches, int* K_out, std::string* err);
int records_noise_reserve(ResidentState& st, const SearchParams& p, int B, int A, int levels, int PL, int* K_out,
                          std::string* err);
void records_noise_range(ResidentState& st, const SearchParams& p, int B, int A, int K, int sim0, int sim1,
                         cudaStream_t stream, int64_t* launches);

// Records of the last search -> the handle's SoA arrays (no-op unless a resident search ran since the last call).
int resident_unpack(ResidentState& st, const Tree& tree, float gamma, std::string* err);

}  // namespace mz
