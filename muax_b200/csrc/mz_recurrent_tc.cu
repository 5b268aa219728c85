#include "mz_recurrent_tc.cuh"
namespace mz {
int recurrent_tc_init(RecurrentTcState& st, const Net&, int, int, std::string*) { st.available = false; st.why = "not built yet"; return 0; }
void recurrent_tc_destroy(RecurrentTcState&) {}
int recurrent_tc_pack(RecurrentTcState&, const Net&, const float*, cudaStream_t, int64_t*) { return 0; }
int recurrent_tc_launch(RecurrentTcState&, const Net&, const Tree&, const int32_t*, const int32_t*, float*, float*, float*, float*, cudaStream_t, int64_t*, std::string* err) { *err = "not built yet"; return 1; }
}
