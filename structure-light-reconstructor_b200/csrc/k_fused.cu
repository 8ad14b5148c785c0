// k_fused.cu — fused pipelines (decode + match + triangulate in one kernel).
#include "slr_device.cuh"

slr_status slr_unfused_mf(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S, int black_thr, int mode,
                          float *d_xyz, uint8_t *d_valid, int32_t *d_match_k, unsigned long long *d_n_points);
slr_status slr_unfused_ge(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col, int black_thr,
                          int white_thr, int scan_w, int have_color, float *d_xyz, uint8_t *d_valid,
                          int32_t *d_match_k, uint8_t *d_color, unsigned long long *d_n_points);

slr_status slr_launch_fused_mf(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S, int black_thr,
                               int mode, float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                               unsigned long long *d_n_points)
{
    return slr_unfused_mf(e, d_stack, batch, F, S, black_thr, mode, d_xyz, d_valid, d_match_k, d_n_points);
}

slr_status slr_launch_fused_ge(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col, int black_thr,
                               int white_thr, int scan_w, int have_color, float *d_xyz, uint8_t *d_valid,
                               int32_t *d_match_k, uint8_t *d_color, unsigned long long *d_n_points)
{
    return slr_unfused_ge(e, d_stack, batch, nbits_col, black_thr, white_thr, scan_w, have_color, d_xyz, d_valid,
                          d_match_k, d_color, d_n_points);
}
