/* tfrs_b200_probe.h -- hardware probes used while developing libtfrs_b200 (tools/libtfrs_b200_probe.so).
 * NOT part of the product ABI (include/tfrs_b200.h) and never loaded by recommenders_b200/. */
#ifndef TFRS_B200_PROBE_H_
#define TFRS_B200_PROBE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* one 128x128x64 UMMA tile with a caller-chosen instruction descriptor, raw TMEM dump (tools/probe_f16acc.py) */
int tfrs_debug_umma_probe(const void* a_img, const void* b_img, uint32_t idesc, int n_cols, uint32_t* out, void* stream);
/* memory-system probe, 128-byte rows (tools/hbm_probe.py): mode 0 copy, 1 random-row read, 2 write only,
 * 3 random read + sequential write, 4 random read + strided write */
int tfrs_debug_hbm_probe(int mode, const void* src, int64_t src_rows, void* dst, int64_t n, int64_t n_rows_out,
                         int64_t ld_floats, float* sink, void* stream);
/* tensor-pipe / TMEM timing probe (tools/tc_rate_probe.py): clock64 counts of R rounds of mode 0 tcgen05.ld x64,
 * 1 the same with .pack::16b, 2 SS MMAs, 3 TS MMAs, 4 both interleaved, 5 tcgen05.st */
int tfrs_debug_tc_rate_probe(int mode, int rounds, int n_ctas, long long* out_cycles, uint32_t* sink, void* stream);
#ifdef __cplusplus
}
#endif
#endif
