// Host-side rating files behind the mfb_blocks_* entry points (no GPU involved).
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "mfb_internal.h"
#include "proto_wire.h"

struct mfb_blocks {
  mfb::Dataset d;  // host staging arrays only
};

using namespace mfb;

namespace mfb {
const Dataset* blocks_data(const mfb_blocks* b) { return &b->d; }
mfb_blocks* blocks_new() {
  mfb_blocks* b = new mfb_blocks();
  b->d.h_run_off.push_back(0);
  b->d.h_block_off.push_back(0);
  return b;
}
Dataset* blocks_mut(mfb_blocks* b) { return &b->d; }
}  // namespace mfb

extern "C" {

int mfb_blocks_read(const char* path, mfb_blocks** out) {
  MFB_REQUIRE(path && out, "NULL argument");
  mfb_blocks* b = blocks_new();
  int rc = load_blocks_file(path, &b->d);
  if (rc) {
    delete b;
    *out = nullptr;
    return rc;
  }
  *out = b;
  return MFB_OK;
}

int mfb_blocks_from_arrays(int64_t nblocks, const int64_t* block_off, int64_t nruns,
                           const int32_t* run_uid, const int32_t* run_off, const int32_t* vid,
                           const float* rating, mfb_blocks** out) {
  MFB_REQUIRE(out && block_off && run_off && nblocks >= 0 && nruns >= 0, "bad argument");
  MFB_REQUIRE(block_off[0] == 0 && block_off[nblocks] == nruns, "block_off must span [0,nruns]");
  mfb_blocks* b = new mfb_blocks();
  const int64_t n = nruns ? run_off[nruns] : 0;
  b->d.h_block_off.assign(block_off, block_off + nblocks + 1);
  b->d.h_run_uid.assign(run_uid, run_uid + nruns);
  b->d.h_run_off.assign(run_off, run_off + nruns + 1);
  b->d.h_vid.assign(vid, vid + n);
  b->d.h_rating.assign(rating, rating + n);
  *out = b;
  return MFB_OK;
}

int mfb_blocks_write(const mfb_blocks* b, const char* path) {
  MFB_REQUIRE(b && path, "NULL argument");
  FILE* f = fopen(path, "wb");
  if (!f) {
    set_error("cannot create %s", path);
    return MFB_E_IO;
  }
  const Dataset& d = b->d;
  std::string buf;
  std::vector<int32_t> rec_off;
  const int64_t nblocks = (int64_t)d.h_block_off.size() - 1;
  for (int64_t k = 0; k < nblocks; k++) {
    const int64_t r0 = d.h_block_off[k], r1 = d.h_block_off[k + 1];
    const int32_t base = d.h_run_off[r0];
    rec_off.resize(r1 - r0 + 1);
    for (int64_t r = r0; r <= r1; r++) rec_off[r - r0] = d.h_run_off[r] - base;
    encode_block((int32_t)(r1 - r0), d.h_run_uid.data() + r0, rec_off.data(), d.h_vid.data() + base,
                 d.h_rating.data() + base, &buf);
    const uint32_t sz = (uint32_t)buf.size();  // getdata.cc:100-103
    if (fwrite(&sz, 1, 4, f) != 4 || fwrite(buf.data(), 1, sz, f) != sz) {
      fclose(f);
      set_error("short write to %s", path);
      return MFB_E_IO;
    }
  }
  if (fclose(f) != 0) {
    set_error("close failed on %s", path);
    return MFB_E_IO;
  }
  return MFB_OK;
}

// DSGD strata: part j keeps the records whose item lies in [bounds[j], bounds[j+1]), with the
// user-runs (and Block boundaries) of the source file; runs left without records are dropped.
int mfb_blocks_split_by_item(const mfb_blocks* b, int nparts, const int32_t* bounds, mfb_blocks** out) {
  MFB_REQUIRE(b && bounds && out && nparts >= 1, "bad argument");
  for (int j = 0; j < nparts; j++) MFB_REQUIRE(bounds[j] <= bounds[j + 1], "bounds must be non-decreasing");
  const Dataset& d = b->d;
  std::vector<mfb_blocks*> parts(nparts);
  for (auto& p : parts) p = blocks_new();
  const int64_t nblocks = (int64_t)d.h_block_off.size() - 1;
  std::vector<int32_t> start(nparts);
  for (int64_t k = 0; k < nblocks; k++) {
    for (int64_t r = d.h_block_off[k]; r < d.h_block_off[k + 1]; r++) {
      for (int j = 0; j < nparts; j++) start[j] = (int32_t)parts[j]->d.h_vid.size();
      for (int32_t t = d.h_run_off[r]; t < d.h_run_off[r + 1]; t++) {
        const int32_t v = d.h_vid[t];
        const int j = (int)(std::upper_bound(bounds, bounds + nparts + 1, v) - bounds) - 1;
        if (j < 0 || j >= nparts) continue;  // item outside every part
        parts[j]->d.h_vid.push_back(v);
        parts[j]->d.h_rating.push_back(d.h_rating[t]);
      }
      for (int j = 0; j < nparts; j++) {
        Dataset& o = parts[j]->d;
        if ((int32_t)o.h_vid.size() == start[j]) continue;
        o.h_run_uid.push_back(d.h_run_uid[r]);
        o.h_run_off.push_back((int32_t)o.h_vid.size());
      }
    }
    for (int j = 0; j < nparts; j++) parts[j]->d.h_block_off.push_back((int64_t)parts[j]->d.h_run_uid.size());
  }
  for (int j = 0; j < nparts; j++) out[j] = parts[j];
  return MFB_OK;
}

// Regrouping of a file's runs (the order of updates is the caller's business: see DESIGN.md 5 - after the first
// epoch it does not matter for the result, so the steady-state DSGD cells are regrouped for speed).
//   merge_users: one run per user - the user's runs concatenated in file order, users in the order of their first
//     run.  (A DSGD cell holds, for each user, the pieces of `split` runs of the source file; merged, the factor row
//     is read and written once per cell instead of once per piece.)
//   longest_first: 1 = runs in descending order of length (stable) - longest-processing-time-first scheduling: a launch
//     ends when its longest run still in flight ends, and a run is a sequential chain; N > 1 = only the runs of N
//     records or more move to the front, the others keep their order.
// `users_per_block` runs per Block.
int mfb_blocks_regroup(const mfb_blocks* b, int merge_users, int longest_first, int users_per_block, mfb_blocks** out) {
  MFB_REQUIRE(b && out && users_per_block >= 1, "bad argument");
  const Dataset& d = b->d;
  const int64_t nruns = (int64_t)d.h_run_uid.size();
  // groups: the output runs, each a list of source runs in file order
  std::vector<int64_t> group_of((size_t)nruns, -1);
  std::vector<int32_t> g_uid;
  std::vector<int64_t> g_len;
  if (merge_users) {
    int32_t maxu = -1;
    for (int32_t u : d.h_run_uid) maxu = std::max(maxu, u);
    std::vector<int64_t> g_of_user((size_t)maxu + 2, -1);
    for (int64_t r = 0; r < nruns; r++) {
      const int64_t len = d.h_run_off[r + 1] - d.h_run_off[r];
      if (len == 0) continue;
      const int32_t u = d.h_run_uid[r];
      if (g_of_user[u] < 0) {
        g_of_user[u] = (int64_t)g_uid.size();
        g_uid.push_back(u);
        g_len.push_back(0);
      }
      group_of[r] = g_of_user[u];
      g_len[g_of_user[u]] += len;
    }
  } else {
    for (int64_t r = 0; r < nruns; r++) {
      const int64_t len = d.h_run_off[r + 1] - d.h_run_off[r];
      if (len == 0) continue;
      group_of[r] = (int64_t)g_uid.size();
      g_uid.push_back(d.h_run_uid[r]);
      g_len.push_back(len);
    }
  }
  const int64_t ng = (int64_t)g_uid.size();
  std::vector<int64_t> order((size_t)ng);
  for (int64_t i = 0; i < ng; i++) order[i] = i;
  if (longest_first == 1) {
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return g_len[x] > g_len[y]; });
  } else if (longest_first > 1) {
    // only the runs of `longest_first` records or more move to the front (longest first); the others keep their order
    const int64_t thr = longest_first;
    std::stable_partition(order.begin(), order.end(), [&](int64_t x) { return g_len[x] >= thr; });
    int64_t nlong = 0;
    while (nlong < ng && g_len[order[nlong]] >= thr) nlong++;
    std::stable_sort(order.begin(), order.begin() + nlong, [&](int64_t x, int64_t y) { return g_len[x] > g_len[y]; });
  }
  std::vector<int64_t> slot((size_t)ng), fill((size_t)ng, 0);
  mfb_blocks* o = blocks_new();  // run_off = block_off = {0}
  Dataset& m = o->d;
  int64_t total = 0;
  for (int64_t i = 0; i < ng; i++) {
    const int64_t g = order[i];
    slot[g] = total;
    total += g_len[g];
    MFB_REQUIRE(total < (int64_t)INT32_MAX, "file too large for int32 offsets");
    m.h_run_uid.push_back(g_uid[g]);
    m.h_run_off.push_back((int32_t)total);
    if ((i + 1) % users_per_block == 0) m.h_block_off.push_back((int64_t)m.h_run_uid.size());
  }
  if (m.h_block_off.back() != (int64_t)m.h_run_uid.size()) m.h_block_off.push_back((int64_t)m.h_run_uid.size());
  m.h_vid.resize(total);
  m.h_rating.resize(total);
  for (int64_t r = 0; r < nruns; r++) {
    const int64_t g = group_of[r];
    if (g < 0) continue;
    const int64_t n = d.h_run_off[r + 1] - d.h_run_off[r], at = slot[g] + fill[g];
    memcpy(m.h_vid.data() + at, d.h_vid.data() + d.h_run_off[r], (size_t)n * sizeof(int32_t));
    memcpy(m.h_rating.data() + at, d.h_rating.data() + d.h_run_off[r], (size_t)n * sizeof(float));
    fill[g] += n;
  }
  *out = o;
  return MFB_OK;
}

// one run per user, users in order of first appearance
int mfb_blocks_merge_runs(const mfb_blocks* b, int users_per_block, mfb_blocks** out) {
  return mfb_blocks_regroup(b, 1, 0, users_per_block, out);
}

void mfb_blocks_free(mfb_blocks* b) {
  if (b && b->d.pinned) mfb_blocks_unpin(b);  // also releases the page-locked compact copy
  delete b;
}
int64_t mfb_blocks_num_blocks(const mfb_blocks* b) { return b ? (int64_t)b->d.h_block_off.size() - 1 : -1; }
int64_t mfb_blocks_num_runs(const mfb_blocks* b) { return b ? (int64_t)b->d.h_run_uid.size() : -1; }
int64_t mfb_blocks_num_ratings(const mfb_blocks* b) { return b ? (int64_t)b->d.h_vid.size() : -1; }
const int64_t* mfb_blocks_block_off(const mfb_blocks* b) { return b->d.h_block_off.data(); }
const int32_t* mfb_blocks_run_uid(const mfb_blocks* b) { return b->d.h_run_uid.data(); }
const int32_t* mfb_blocks_run_off(const mfb_blocks* b) { return b->d.h_run_off.data(); }
const int32_t* mfb_blocks_vid(const mfb_blocks* b) { return b->d.h_vid.data(); }
const float* mfb_blocks_rating(const mfb_blocks* b) { return b->d.h_rating.data(); }

}  // extern "C"
