// One tile shape of the chain kernels (see tnb_launch.h): compiled once per (TNB_INST_TILE, TNB_INST_WPL).
#include "tnb_launch.h"

namespace tnb {
template bool launch_tw<TNB_INST_TILE, TNB_INST_WPL>(Rt&, const Params&, bool, bool, bool);
template bool launch_treegen_t<TNB_INST_TILE, TNB_INST_WPL>(Rt&, const Params&);
}  // namespace tnb
