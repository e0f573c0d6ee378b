/* orc_margin.c -- CPU ORACLE: marginalization. Placeholder translation unit;
 * the restatement of Marginalizer::marginalize lands with that row. */
#include "orc_oracle.h"
int orc_marginalize(orc_handle *o, int32_t n_remove, const int64_t *remove_frame_ids, int32_t *m_out,
                    int32_t max_m, double *A_out, double *b_out, int32_t *nblk_out, int32_t max_blk,
                    d2ba_blockref *refs_out) {
  (void)o; (void)n_remove; (void)remove_frame_ids; (void)m_out; (void)max_m; (void)A_out; (void)b_out;
  (void)nblk_out; (void)max_blk; (void)refs_out;
  return 100;
}
