// Fused circuit execution (placeholder: gate-by-gate until the tile executor lands).
#include "qsv_internal.h"

namespace qsv {

void apply_ops_fused(State &sv, const std::vector<LoweredGate> &gates) {
    for (const auto &g : gates) launch_gate(sv, g);
}

}  // namespace qsv
