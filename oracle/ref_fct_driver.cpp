// ref_fct_driver.cpp -- C entry points around three pieces of the REFERENCE's corner tracker ("FCT":
// event-cam-tracking/event-cam-fast-corner-tracker/metavision_time_surface_periodic_group_track.cpp),
// compiled where they lie (TEST INFRASTRUCTURE ONLY).  The file as a whole needs the Metavision SDK
// and OpenCV; the Makefile (target ref_fct) extracts three line ranges of it into oracle/_ref/ at
// build time -- never into the repository -- and this driver compiles them against fct_shim.h:
//   _ref/fct_globals.inc   FCT:44-45    the circle tables circle3_ / circle4_
//   _ref/fct_filter.inc    FCT:61-153   struct Corner, class CornerFilter (the box-NMS)
//   _ref/fct_tracker.inc   FCT:163-537  DirectionVector, TrackedCorner, CornerGroup, CornerTracker
//   _ref/fct_callback.inc  FCT:884-1070 the event callback lambda: stamp the time surface, test
//                                       every event of the range for a corner
#include "fct_shim.h"

#include "_ref/fct_globals.inc"
const int ARRAY_SIZE = 16384;  // FCT:48
#include "_ref/fct_filter.inc"
#include "_ref/fct_tracker.inc"

extern "C" {

// One callback range through the reference's lambda.  surface: [720][1280] timestamps, updated in
// place (the lambda hard-codes the 1280 x 720 border test, FCT:948).  flag = time_surface_flag (the
// reference sets it to 1 after the first slice; 0 only stamps).  Returns the number of corners;
// their (x, y) pairs go to out_xy.
long ref_fct_callback(const void* ev_begin, const void* ev_end, long long* surface, int flag,
                      int* out_xy, long cap) {
    static int data[ARRAY_SIZE];
    int data_index = 0;
    std::mutex frame_mutex;
    Metavision::MostRecentTimestampBuffer time_surface(720, 1280, surface);
    Metavision::timestamp last_time = 0;
    int time_surface_flag = flag;
    cv::Mat corner_img;
    std::vector<Corner> corners;
#include "_ref/fct_callback.inc"
    aggregate_events_fct(static_cast<const Metavision::EventCD*>(ev_begin),
                         static_cast<const Metavision::EventCD*>(ev_end));
    (void)last_time;
    for (size_t i = 0; i < corners.size() && (long)i < cap; i++) {
        out_xy[2 * i] = corners[i].x;
        out_xy[2 * i + 1] = corners[i].y;
    }
    return (long)corners.size();
}

// CornerFilter::filterCorners (FCT:81-151; call site FCT:832 with box size 15, threshold 0.5)
long ref_fct_filter(const int* xy, long n, int width, int height, int box_size, int* out_xy_label,
                    long cap) {
    std::vector<Corner> in((size_t)n);
    for (long i = 0; i < n; i++) in[(size_t)i] = Corner{xy[2 * i], xy[2 * i + 1], 0};
    const std::vector<Corner> out = CornerFilter::filterCorners(in, width, height, box_size, 0.5f);
    for (size_t i = 0; i < out.size() && (long)i < cap; i++) {
        out_xy_label[3 * i] = out[i].x;
        out_xy_label[3 * i + 1] = out[i].y;
        out_xy_label[3 * i + 2] = out[i].label;
    }
    return (long)out.size();
}

// CornerTracker with the reference app's parameters (FCT:805-813); one object per handle
void* ref_fct_tracker_new() { return new CornerTracker(30.0f, 30, 10, 5, 0.8f, 0.3f, 100.0f); }
void ref_fct_tracker_delete(void* t) { delete static_cast<CornerTracker*>(t); }
// one slice: the filtered corners in, the active tracks out, 8 floats per track:
// x, y, label, frame_count, frames_since_last_detection, velocity.x, velocity.y, group_id
long ref_fct_tracker_update(void* tp, const int* xy, long n, float* out, long cap) {
    CornerTracker* t = static_cast<CornerTracker*>(tp);
    std::vector<Corner> in((size_t)n);
    for (long i = 0; i < n; i++) in[(size_t)i] = Corner{xy[2 * i], xy[2 * i + 1], (int)i};
    const std::vector<TrackedCorner> tr = t->updateTrackedCorners(in);
    for (size_t i = 0; i < tr.size() && (long)i < cap; i++) {
        float* o = out + 8 * i;
        o[0] = (float)tr[i].x;
        o[1] = (float)tr[i].y;
        o[2] = (float)tr[i].label;
        o[3] = (float)tr[i].frame_count;
        o[4] = (float)tr[i].frames_since_last_detection;
        o[5] = tr[i].velocity.x;
        o[6] = tr[i].velocity.y;
        o[7] = (float)tr[i].group_id;
    }
    return (long)tr.size();
}
}
