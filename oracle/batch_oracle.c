/* batch_oracle.c -- batched CPU pass of the hot path (TEST INFRASTRUCTURE ONLY). Filled in below. */
#include "trgt_oracle.h"
int tro_process_loci(const tro_batch *b, uint32_t lo, uint32_t hi, tro_batch_out *out) {
  (void)b; (void)lo; (void)hi; (void)out;
  return -1;
}
