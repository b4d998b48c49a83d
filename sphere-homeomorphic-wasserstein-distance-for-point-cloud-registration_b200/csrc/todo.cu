// Temporary: entry points not implemented yet return SHWD_ERR_UNSUPPORTED (removed as each lands).
#include "common.cuh"
extern "C" {
int shwd_project_circle(const float*, const float*, int, int, int, float*, void*) { return SHWD_ERR_UNSUPPORTED; }
int shwd_project_circle_bwd(const float*, const float*, int, int, int, const float*, float*, void*) { return SHWD_ERR_UNSUPPORTED; }
int shwd_project_line(const float*, const float*, int, int, int, float*, void*) { return SHWD_ERR_UNSUPPORTED; }
int shwd_project_line_bwd(const float*, int, int, int, const float*, float*, void*) { return SHWD_ERR_UNSUPPORTED; }
size_t shwd_segmented_sort_workspace_bytes(int, int) { return 0; }
int shwd_segmented_sort(const float*, int, int, float*, int64_t*, void*, size_t, void*) { return SHWD_ERR_UNSUPPORTED; }
size_t shwd_circular_w1_workspace_bytes(int, int, int) { return 0; }
int shwd_circular_w1(const float*, const float*, int, int, int, float*, float*, float*, void*, size_t, void*) { return SHWD_ERR_UNSUPPORTED; }
int shwd_euclid_sw(const float*, const float*, int, int, float, float*, float*, float*, void*) { return SHWD_ERR_UNSUPPORTED; }
int shwd_unsort(const float*, const int64_t*, int, int, float*, void*) { return SHWD_ERR_UNSUPPORTED; }
int shwd_peak_fp32(float*, int, double*, void*) { return SHWD_ERR_UNSUPPORTED; }
int shwd_peak_mufu(float*, int, double*, void*) { return SHWD_ERR_UNSUPPORTED; }
}
