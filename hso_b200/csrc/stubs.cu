// TEMPORARY: entry points whose kernels land in align.cu / pose.cu (next milestone). They fail loudly.
#include "hso_internal.h"
extern "C" {
int hso_align_batch(hso_ctx*, hso_frame_id, int, const hso_align_job*, const hso_frame_id*, int, hso_align_result*) { return HSO_ERR_INVALID; }
int hso_pose_optimize(hso_ctx*, double, int, int, int, const double*, const double*, const int32_t*, int, const double*, const double*,
                      const int8_t*, const int8_t*, const int8_t*, const double*, uint8_t*, hso_pose_result*) { return HSO_ERR_INVALID; }
int hso_pose_optimize_batch(hso_ctx*, double, int, int, const int32_t*, const int32_t*, const double*, const double*, const int32_t*,
                            const int32_t*, const double*, const double*, const int8_t*, const int8_t*, const int8_t*, const double*,
                            uint8_t*, hso_pose_result*) { return HSO_ERR_INVALID; }
}
