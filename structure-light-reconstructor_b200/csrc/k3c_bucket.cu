// k3c_bucket.cu — K3c: Gray-only projector-cell bucket triangulation (placeholder launcher).
#include "slr_device.cuh"

slr_status slr_launch_bucket_triangulate(slr_engine *e, const int32_t *d_col, const int32_t *d_row,
                                         const uint8_t *d_mask, int batch, int scan_w, int scan_h,
                                         float *d_sum, uint8_t *d_cnt, unsigned long long *d_n_cells)
{
    slr_set_error("slr_bucket_triangulate: not built yet");
    return SLR_ERR_INVALID;
}
