// sb_inst.cu -- instantiates every kernel of one registry functor.
// Compiled once per functor with -DSB_FN=<functor struct> -DSB_ENTRY=<entry function> -DSB_NAME="<name>".
#include "sb_registry.h"

namespace sb {
FunctorEntry SB_ENTRY() { return make_entry<SB_FN>(SB_NAME); }
}  // namespace sb
