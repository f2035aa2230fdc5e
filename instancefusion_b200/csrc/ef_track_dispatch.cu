// ef_track_dispatch.cu -- picks the CTA shape of the persistent tracker kernel per handle.
// ef_track_kernel.cu is compiled twice (EF_TRACK_THREADS = 256 and 384, see its header); small images run the 8-warp
// variant, large ones the 12-warp variant.  EF_TRACK_VARIANT=256|384 in the environment overrides (sweeps).
#include <stdlib.h>

#include "ef_tracker.h"

namespace ef
{

#define EF_DECLARE_VARIANT(T)                                                                                                              \
    int device_track_init_t##T(ef_tracker * t);                                                                                            \
    int device_track_configure_t##T(ef_tracker * t, int grid_ctas);                                                                        \
    bool device_track_supported_t##T(const ef_tracker * t);                                                                                \
    int device_track_trace_t##T(ef_tracker * t, double * out32, long long * calls);                                                                                \
    void device_track_destroy_t##T(ef_tracker * t);                                                                                        \
    int device_track_launch_t##T(ef_tracker * t, const float * trans, const float * rot, int rgb_only, float icp_weight, int pyramid,     \
                                 int fast_odom, int so3);                                                                                  \
    int device_track_finish_t##T(ef_tracker * t, float * trans, float * rot);
EF_DECLARE_VARIANT(256)
EF_DECLARE_VARIANT(384)

// the batched builds (csrc/Makefile): EF_TRACK_ALT -- up to 4 sequences per launch, the worker CTAs time-sliced between them
// (the default) -- and EF_TRACK_GROUPS = 2 -- two sequences, two thread groups per CTA (EF_BATCH_MODE=groups)
int device_track_launch_batch_g2(ef_tracker * const * ts, const float * const * trans, const float * const * rot, int rgb_only, float icp_weight,
                                     int pyramid, int fast_odom, int so3, cudaStream_t stream);
int device_track_launch_batch_alt(ef_tracker * const * ts, int n, const float * const * trans, const float * const * rot, int rgb_only, float icp_weight,
                                  int pyramid, int fast_odom, int so3, int grid_ctas, cudaStream_t stream);
static bool batch_groups_mode()
{
    const char * env = getenv("EF_BATCH_MODE");
    return env && env[0] == 'g';
}
int device_track_batch_width() { return batch_groups_mode() ? 2 : 4; }
int device_track_launch_batch(ef_tracker * const * ts, int n, const float * const * trans, const float * const * rot, int rgb_only, float icp_weight,
                              int pyramid, int fast_odom, int so3, cudaStream_t stream)
{
    if(batch_groups_mode())
    {
        if(n != 2) return EF_ERR_INVALID_ARGUMENT;
        return device_track_launch_batch_g2(ts, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3, stream);
    }
    if(n < 2 || n > 4) return EF_ERR_INVALID_ARGUMENT;
    return device_track_launch_batch_alt(ts, n, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3, ts[0]->grid_ctas, stream);
}

static int pick_variant(const ef_tracker * t)
{
    const char * env = getenv("EF_TRACK_VARIANT");
    if(env && atoi(env) == 256) return 256;
    if(env && atoi(env) == 384) return 384;
    // measured crossover between 640x480 (256 threads 3 % faster) and 1280x720 (384 threads 10 % faster)
    return ((size_t)t->width * t->height >= (size_t)600000) ? 384 : 256;
}

int device_track_init(ef_tracker * t)
{
    t->track_variant = pick_variant(t);
    return t->track_variant == 384 ? device_track_init_t384(t) : device_track_init_t256(t);
}
int device_track_configure(ef_tracker * t, int grid_ctas)
{
    return t->track_variant == 384 ? device_track_configure_t384(t, grid_ctas) : device_track_configure_t256(t, grid_ctas);
}
int device_track_trace(ef_tracker * t, double * out32, long long * calls)
{
    return t->track_variant == 384 ? device_track_trace_t384(t, out32, calls) : device_track_trace_t256(t, out32, calls);
}
bool device_track_supported(const ef_tracker * t)
{
    return t->track_variant == 384 ? device_track_supported_t384(t) : device_track_supported_t256(t);
}
void device_track_destroy(ef_tracker * t)
{
    if(t->track_variant == 384) device_track_destroy_t384(t);
    else device_track_destroy_t256(t);
}
int device_track_launch(ef_tracker * t, const float * trans, const float * rot, int rgb_only, float icp_weight, int pyramid, int fast_odom,
                        int so3)
{
    return t->track_variant == 384 ? device_track_launch_t384(t, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3)
                                   : device_track_launch_t256(t, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
}
int device_track_finish(ef_tracker * t, float * trans, float * rot)
{
    return t->track_variant == 384 ? device_track_finish_t384(t, trans, rot) : device_track_finish_t256(t, trans, rot);
}

} // namespace ef
