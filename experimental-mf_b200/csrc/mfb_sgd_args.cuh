// Kernel arguments of the plain-SGD epoch kernels (mfb_sgd.cu, mfb_sgd_stream.cu).
#ifndef MFB_SGD_ARGS_CUH
#define MFB_SGD_ARGS_CUH

#include <stdint.h>

namespace mfb {

struct SgdArgs {
  float* theta;
  float* phi;
  float* bu;
  float* bv;
  const int32_t* run_uid;
  const int32_t* run_off;
  const int32_t* vid;
  const float* rating;
  int* counter;
  int run_begin, nruns, nvec;  // runs [run_begin, nruns) are processed
  // stream/burst kernels: the first big_spans spans are full (LPR resp. 32 consecutive runs), the runs
  // after them are handed out one by one, so that the kernel's tail is one run long, not one span
  int big_spans;
  int span_runs;  // burst kernel: runs per claim (<= 32)
  // stream kernel: float4 index of (item v, this lane's vector i) = v*phi_row4 + i*phi_line4 + lane.
  // Rows as they are: (nvec, LPR).  Plane layout: (LPR, planes of nv*LPR float4) - the 128-byte lines
  // of one row then lie nv*128 bytes apart and hash to different L2 slices.
  // phi_pair4 != 0: lanes 2m, 2m+1 (one 32-byte sector) of a vector lie phi_pair4 float4 after those of
  // lanes 2m-2, 2m-1 - the sector layout: plane j holds sector j of every item (32 bytes per item), so a
  // row spreads over 16 chunks of 256 bytes and a chunk holds one sector of 8 consecutive items
  int64_t phi_row4, phi_line4, phi_pair4;
  float eta, lameta, lm1, gb;
  int ld_flavour, st_flavour, bias_flavour;  // see mfb_group.cuh; bias: 0 red.add, 1 skip, 2 st.cg
  int throttle;  // streaming kernel: wait for the previous record's bias atomic before the next reductions
  // staleness probe (debug option "probe"): version[v] counts the updates of item v that the L2 has
  // performed; probe_out[0..3] += {sum of (updates performed between my read and my update), updates,
  // the same two restricted to item probe_item}
  int* version;
  unsigned long long* probe_out;
  int probe_item;
};

struct Context;
struct Dataset;
// mfb_sgd_stream.cu: the sub-warp streaming kernel (production schedule); returns MFB_OK or an
// error; `handled` is false when the row shape has no streaming instantiation
int launch_sgd_stream(Context* c, const Dataset* d, const SgdArgs& a, int mode, bool* handled);
// mfb_sgd_burst.cu: warp per run, B records per step, everything requested one batch ahead (the
// kernel for launches that the concurrency bounds keep narrow)
int launch_sgd_burst(Context* c, const Dataset* d, const SgdArgs& a, int mode, bool* handled);

}  // namespace mfb
#endif
