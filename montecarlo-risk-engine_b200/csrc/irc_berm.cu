// Template instantiations of the fused interest-rate / credit main kernel for books that
// contain Bermudan exercise units (own translation unit: compiles in parallel with irc.cu).
#include "irc_main.cuh"

int irc_dispatch_main_berm(mcre_irc_plan *p, const mcre::RngDev &r, const mcre::ShardDev &sh, double *d_partial,
                           double *d_spill, double *d_shift, cudaStream_t st) {
  return irc_dispatch_main<true>(p, r, sh, d_partial, d_spill, d_shift, st);
}
