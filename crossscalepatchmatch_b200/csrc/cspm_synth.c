/* libcspm_synth.so — host-only build of include/cspm_synth.h for the Python
 * side (synth.py, bench.py, tests).  No CUDA, no dependency on libcspm_b200.so:
 * the benchmark's reference arm generates its inputs without loading the product. */
#include "../../include/cspm_synth.h"

int cspm_synth_make_pair(int w, int h, int max_dis, uint64_t seed, uint8_t* left, uint8_t* right, double* gt) {
  return cspm_synth_pair(w, h, max_dis, seed, left, right, gt);
}
int cspm_synth_make_pair_variant(int w, int h, int max_dis, uint64_t seed, uint64_t variant, uint8_t* left, uint8_t* right, double* gt) {
  return cspm_synth_pair_variant(w, h, max_dis, seed, variant, left, right, gt);
}
uint64_t cspm_synth_seed_base(void) { return CSPM_SYNTH_SEED_BASE; }
