// kernels_exact.cu -- BITEXACT flavour.  MUST be compiled with -fmad=false (build.py does): the
// reference's arithmetic never fuses a multiply with an add (JS doubles / separately rounded f32).
#include "table.cuh"

namespace tsim {
const KernelTable *exact_kernels() { return Launchers<true>::table(); }
}  // namespace tsim
