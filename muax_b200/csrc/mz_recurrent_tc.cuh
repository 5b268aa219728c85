// mz_recurrent_tc.cuh — interface of the tcgen05 recurrent_fn kernel (implementation: mz_recurrent_tc.cu).
//
// Throughput mode of the search (mz_search_args.precision = MZ_PRECISION_BF16): muax's `_recurrent_inference`
// (muax/model.py:265-282: Dynamic MLP -> min-max -> Prediction MLP -> 2x support_to_scalar(softmax)) for the whole
// batch of trees awaiting expansion as ONE kernel whose dense layers run on the 5th-generation tensor cores:
// 128-row tiles, bf16 operands (activations rounded per layer, weights pre-packed once per mz_set_weights), fp32
// accumulation in tensor memory, weights staged into shared memory by TMA bulk copies.
#pragma once
#include <string>

#include "mz_device.cuh"

namespace mz {

struct RecurrentTcState {
  bool available = false;   // the network shapes fit the kernel (see recurrent_tc_init)
  std::string why;          // ... and if not, why
  void* impl = nullptr;
};

int recurrent_tc_init(RecurrentTcState& st, const Net& net, int batch, int device, std::string* err);
void recurrent_tc_destroy(RecurrentTcState& st);
// Re-packs the raw fp32 blob (device) into the bf16 UMMA operand images; call after every mz_set_weights.
int recurrent_tc_pack(RecurrentTcState& st, const Net& net, const float* raw_weights_dev, cudaStream_t stream,
                      int64_t* launches);
// recurrent_fn for all B rows: embeddings[b, parent[b]] and action[b] -> reward[b], value[b], prior logits [b, A],
// next embedding [b, E] (the I/O of the fp32 recurrent_kernel of mzsearch.cu).
int recurrent_tc_launch(RecurrentTcState& st, const Net& net, const Tree& t, const int32_t* parent,
                        const int32_t* action, float* reward, float* value, float* logits, float* next_emb,
                        cudaStream_t stream, int64_t* launches, std::string* err);

// The search of the throughput mode keeps its tree embeddings in bf16 (the precision the tensor core reads them in):
// rows [B][N][E rounded up to 8], so that gathering a parent row and storing the next state are plain 16-byte copies
// to / from the operand buffer.  begin: (re)allocate, node 0 = bf16(root_emb); launch: recurrent_fn from
// [b][parent[b]] + action[b], next state stored at [b][next[b]]; export: the fp32 [B][N][E] array of the tree view.
int recurrent_tc_tree_begin(RecurrentTcState& st, const Net& net, int B, int N, const float* root_emb, bool clear,
                            cudaStream_t stream, int64_t* launches, std::string* err);
int recurrent_tc_tree_launch(RecurrentTcState& st, const Net& net, int B, int N, const int32_t* parent,
                             const int32_t* action, const int32_t* next, float* reward, float* value, float* logits,
                             cudaStream_t stream, int64_t* launches, std::string* err);
int recurrent_tc_tree_export(RecurrentTcState& st, const Net& net, int B, int N, float* embeddings, cudaStream_t stream,
                             std::string* err);

// `_root_inference` (muax/model.py:251-263) on the same kernel: from observations (Representation -> emb_out, value,
// prior logits) or from a caller-made embedding (value, prior logits).  Only when recurrent_tc_has_root says so.
bool recurrent_tc_has_root(const RecurrentTcState& st, bool from_obs);
int recurrent_tc_root(RecurrentTcState& st, const Net& net, int B, const float* obs, const float* emb_in, float* value,
                      float* logits, float* emb_out, cudaStream_t stream, int64_t* launches, std::string* err);

}  // namespace mz
