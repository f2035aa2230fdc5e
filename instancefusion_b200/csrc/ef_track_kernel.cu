// placeholder: EF_SOLVE_DEVICE path (persistent tracker kernel) -- filled in next
#include "ef_tracker.h"
namespace ef
{
int device_track_init(ef_tracker *) { return EF_OK; }
void device_track_destroy(ef_tracker *) {}
int device_track_launch(ef_tracker * t, const float *, const float *, int, float, int, int, int)
{
    t->err = "EF_SOLVE_DEVICE not built";
    return EF_ERR_UNSUPPORTED;
}
int device_track_finish(ef_tracker *, float *, float *) { return EF_ERR_UNSUPPORTED; }
} // namespace ef
