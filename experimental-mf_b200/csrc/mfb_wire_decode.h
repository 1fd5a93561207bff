// Device-side blocks.proto decoder (mfb_wire_decode.cu), used by the out-of-core epoch (mfb_file_epoch.cu).
#ifndef MFB_WIRE_DECODE_H
#define MFB_WIRE_DECODE_H

#include <stdint.h>

#include "mfb_internal.h"

namespace mfb {

enum { WIRE_OK = 0, WIRE_E_FORMAT = 1, WIRE_E_UID = 2, WIRE_E_VID = 3, WIRE_E_CAPACITY = 4 };

// what the decode of one chunk reports back to the host (one 24-byte D2H copy)
struct WireResult {
  int32_t err;        // first error seen (WIRE_E_*), 0 = none
  int32_t err_run;    // ... in this user-run of the chunk
  int32_t top_count;  // records of the most rated item of the chunk
  int32_t pad;
  long long nratings;  // records in the chunk
};

// Decodes `nruns` serialized mf.User messages: d_span[2r], d_span[2r+1] = byte range of user r inside d_raw (8-byte
// aligned, readable 16 bytes past the last range).  Writes run_uid[nruns], run_off[nruns+1], vid/rating[nratings]
// (at most cap_ratings), per-item counts into d_hist[nv] (may be NULL) and *d_res.  d_count[nruns] is scratch.
int launch_wire_decode(Context* c, cudaStream_t stream, const void* d_raw, const int32_t* d_span, int nruns,
                       int64_t cap_ratings, int32_t* d_run_uid, int32_t* d_run_off, int32_t* d_count, int32_t* d_vid,
                       float* d_rating, int32_t* d_hist, WireResult* d_res);

}  // namespace mfb
#endif
