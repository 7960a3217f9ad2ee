// Sharded state vectors over NCCL (placeholder until the distributed path lands).
#include "qsv_internal.h"

namespace qsv {
void dist_free(State &) {}
}  // namespace qsv

extern "C" {
int qsv_dist_unique_id(void *) { qsv::set_last_error("distributed path not built"); return 1; }
int qsv_dist_init(qsv_state *, const void *, int, int) { qsv::set_last_error("distributed path not built"); return 1; }
int qsv_dist_finalize(qsv_state *) { return 0; }
int qsv_dist_swap_bits(qsv_state *, int, int, size_t) { qsv::set_last_error("distributed path not built"); return 1; }
int qsv_dist_allreduce_f64(qsv_state *, double *, int) { qsv::set_last_error("distributed path not built"); return 1; }
int qsv_dist_last_swap_stats(const qsv_state *, uint64_t *, float *) { qsv::set_last_error("distributed path not built"); return 1; }
}
