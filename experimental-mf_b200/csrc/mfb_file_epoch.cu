// Out-of-core epoch: one SGD epoch straight from the [u32 size][mf.Block] training FILE, the way
// the reference runs every epoch (read -> parse -> update as a pipeline, main.cc:45-50; mf.h:24-45 reads a frame,
// mf.h:57-69 parses it, mf.h:76 updates) - nothing of the file stays resident, so data larger than HBM can be
// trained on.  B200 shape of that pipeline:
//
//   frames of the next chunk  --decode (all host cores, proto_wire.cc)-->  flat SoA arrays in PINNED host memory
//      --cudaMemcpyAsync on the copy stream-->  one of TWO device tile buffers  --epoch kernel on the compute stream
//
// The decode of chunk i+1 runs while chunk i is copied and updated; the device holds two tile buffers of
// `tile_ratings` records each, whatever the size of the file.  Chunks are whole Blocks in file order, so the
// update order is the resident epoch's (bit-exact in the ORDERED schedule).
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <future>
#include <thread>
#include <vector>

#include "mfb_internal.h"
#include "proto_wire.h"

namespace mfb {
namespace {

struct Frame {
  const uint8_t* data;
  size_t size;
};

// pinned staging + device tiles of one pipeline slot
struct Slot {
  int32_t *h_run_uid = nullptr, *h_run_off = nullptr, *h_vid = nullptr;
  float* h_rating = nullptr;
  int32_t *d_run_uid = nullptr, *d_run_off = nullptr, *d_vid = nullptr;
  float* d_rating = nullptr;
  cudaEvent_t copied = nullptr, computed = nullptr;
  bool busy = false;  // a kernel that reads the device tiles may still be queued
};

struct Decoded {  // one chunk, decoded frame by frame
  std::vector<BlockSink> sinks;
  std::vector<int64_t> block_runs;  // runs per block
  int64_t nruns = 0, nratings = 0;
  bool ok = true;
  size_t bad_frame = 0;
};

void decode_chunk(const std::vector<Frame>& frames, size_t f0, size_t f1, Decoded* out) {
  const size_t nf = f1 - f0;
  out->sinks.assign(nf, BlockSink());
  std::vector<char> ok(nf, 1);
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nthreads = std::min<size_t>(hw, std::max<size_t>(1, nf / 2));
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nf) break;
      out->sinks[i].rec_off.assign(1, 0);
      ok[i] = decode_block(frames[f0 + i].data, frames[f0 + i].size, &out->sinks[i]) ? 1 : 0;
    }
  };
  std::vector<std::thread> pool;
  for (size_t t = 1; t < nthreads; t++) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  out->nruns = out->nratings = 0;
  out->ok = true;
  for (size_t i = 0; i < nf; i++) {
    if (!ok[i] && out->ok) {
      out->ok = false;
      out->bad_frame = f0 + i;
    }
    out->nruns += (int64_t)out->sinks[i].uid.size();
    out->nratings += (int64_t)out->sinks[i].vid.size();
  }
}

// flat arrays of the chunk into the pinned buffers of a slot (offsets relative to the chunk); returns the share
// of the most rated item among the chunk's records (for the hot-row budget)
double stitch_chunk(const Decoded& dc, Slot* s, int nv, int nu, bool* ids_ok) {
  const size_t nf = dc.sinks.size();
  std::vector<int64_t> rec0(nf + 1, 0), run0(nf + 1, 0);
  for (size_t i = 0; i < nf; i++) {
    rec0[i + 1] = rec0[i] + (int64_t)dc.sinks[i].vid.size();
    run0[i + 1] = run0[i] + (int64_t)dc.sinks[i].uid.size();
  }
  s->h_run_off[0] = 0;
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nthreads = std::min<size_t>(hw, std::max<size_t>(1, nf / 2));
  std::vector<std::vector<int32_t>> hist(nthreads);
  std::atomic<size_t> next(0);
  std::atomic<int> bad(0);
  auto worker = [&](size_t t) {
    std::vector<int32_t>& cnt = hist[t];
    cnt.assign((size_t)nv, 0);
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nf) break;
      const BlockSink& b = dc.sinks[i];
      if (!b.vid.empty()) {
        memcpy(s->h_vid + rec0[i], b.vid.data(), b.vid.size() * sizeof(int32_t));
        memcpy(s->h_rating + rec0[i], b.rating.data(), b.rating.size() * sizeof(float));
        for (int32_t v : b.vid) {
          if ((uint32_t)v >= (uint32_t)nv) bad = 1;
          else cnt[v]++;
        }
      }
      for (size_t k = 0; k < b.uid.size(); k++) {
        if ((uint32_t)b.uid[k] >= (uint32_t)nu) bad = 1;
        s->h_run_uid[run0[i] + k] = b.uid[k];
        s->h_run_off[run0[i] + k + 1] = (int32_t)(rec0[i] + b.rec_off[k + 1]);
      }
    }
  };
  std::vector<std::thread> pool;
  for (size_t t = 1; t < nthreads; t++) pool.emplace_back(worker, t);
  worker(0);
  for (auto& t : pool) t.join();
  *ids_ok = bad.load() == 0;
  int64_t top = 0;
  for (int v = 0; v < nv; v++) {
    int64_t sum = 0;
    for (auto& c : hist) sum += c[v];
    top = std::max(top, sum);
  }
  return dc.nratings ? (double)top / (double)dc.nratings : 0.0;
}

void free_slot(Slot* s) {
  cudaFreeHost(s->h_run_uid); cudaFreeHost(s->h_run_off); cudaFreeHost(s->h_vid); cudaFreeHost(s->h_rating);
  cudaFree(s->d_run_uid); cudaFree(s->d_run_off); cudaFree(s->d_vid); cudaFree(s->d_rating);
  if (s->copied) cudaEventDestroy(s->copied);
  if (s->computed) cudaEventDestroy(s->computed);
  *s = Slot();
}

int alloc_slot(Slot* s, int64_t cap_ratings, int64_t cap_runs) {
  MFB_CUDA(cudaMallocHost(&s->h_run_uid, cap_runs * sizeof(int32_t)));
  MFB_CUDA(cudaMallocHost(&s->h_run_off, (cap_runs + 1) * sizeof(int32_t)));
  MFB_CUDA(cudaMallocHost(&s->h_vid, cap_ratings * sizeof(int32_t)));
  MFB_CUDA(cudaMallocHost(&s->h_rating, cap_ratings * sizeof(float)));
  MFB_CUDA(cudaMalloc(&s->d_run_uid, cap_runs * sizeof(int32_t)));
  MFB_CUDA(cudaMalloc(&s->d_run_off, (cap_runs + 1) * sizeof(int32_t)));
  MFB_CUDA(cudaMalloc(&s->d_vid, cap_ratings * sizeof(int32_t)));
  MFB_CUDA(cudaMalloc(&s->d_rating, cap_ratings * sizeof(float)));
  MFB_CUDA(cudaEventCreateWithFlags(&s->copied, cudaEventDisableTiming));
  MFB_CUDA(cudaEventCreateWithFlags(&s->computed, cudaEventDisableTiming));
  return MFB_OK;
}

}  // namespace
}  // namespace mfb

using namespace mfb;

extern "C" int mfb_sgd_epoch_from_file(mfb_ctx* h, const char* path, float eta, float lambda, float gb, int mode,
                                       int64_t tile_ratings, int64_t* ratings_out) {
  MFB_REQUIRE(h && path, "NULL argument");
  MFB_REQUIRE(mode == MFB_MODE_HOGWILD || mode == MFB_MODE_ORDERED || mode == MFB_MODE_ATOMIC, "bad mode %d", mode);
  Context* c = &h->c;
  MFB_CUDA(cudaSetDevice(c->device));
  if (tile_ratings <= 0) tile_ratings = (int64_t)8 << 20;
  tile_ratings = std::max<int64_t>(tile_ratings, 1024);
  if (ratings_out) *ratings_out = 0;

  const int fd = open(path, O_RDONLY);
  if (fd < 0) {
    set_error("cannot open %s", path);
    return MFB_E_IO;
  }
  struct stat st;
  if (fstat(fd, &st) != 0) {
    close(fd);
    set_error("cannot stat %s", path);
    return MFB_E_IO;
  }
  const size_t size = (size_t)st.st_size;
  if (size == 0) {
    close(fd);
    return MFB_OK;
  }
  void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) {
    set_error("cannot mmap %s", path);
    return MFB_E_IO;
  }
  madvise(map, size, MADV_SEQUENTIAL);
  // frame boundaries: one jump per frame
  std::vector<Frame> frames;
  {
    const uint8_t* p = (const uint8_t*)map;
    const uint8_t* end = p + size;
    while (end - p >= 4) {  // util.h:81
      uint32_t isize;
      memcpy(&isize, p, 4);
      p += 4;
      if ((size_t)(end - p) < isize) {
        set_error("%s: truncated frame (%u bytes wanted, %zu left)", path, isize, (size_t)(end - p));
        munmap(map, size);
        return MFB_E_IO;
      }
      frames.push_back(Frame{p, isize});
      p += isize;
    }
  }
  // A record written by protobuf takes 9 bytes or more on the wire (tag, length, vid tag + varint, rating tag + 4
  // bytes; both fields are `required`, blocks.proto:4-5) and a user 4 more: frames are grouped so that bytes / 9 <=
  // tile_ratings, which bounds the decoded size of a chunk; a single larger frame gets a chunk of its own and sizes
  // the buffers.  (A file with shorter records is still decoded; a chunk that outgrows its buffers is an error.)
  std::vector<size_t> chunk_first{0};
  size_t acc = 0, biggest = 0;
  for (size_t i = 0; i < frames.size(); i++) {
    if (acc > 0 && (acc + frames[i].size) / 9 + 1 > (size_t)tile_ratings) {
      chunk_first.push_back(i);
      biggest = std::max(biggest, acc);
      acc = 0;
    }
    acc += frames[i].size;
  }
  biggest = std::max(biggest, acc);
  chunk_first.push_back(frames.size());
  const size_t nchunks = chunk_first.size() - 1;
  const int64_t cap_ratings = (int64_t)(biggest / 9 + 16);
  const int64_t cap_runs = (int64_t)(biggest / 4 + 16);
  MFB_REQUIRE(cap_ratings < (int64_t)INT32_MAX, "tile too large for int32 offsets");

  Slot slots[2];
  int rc = MFB_OK;
  for (int b = 0; b < 2 && rc == MFB_OK; b++) rc = alloc_slot(&slots[b], cap_ratings, cap_runs);
  if (!c->copy_stream && rc == MFB_OK) {
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cannot create the copy stream");
      rc = MFB_E_CUDA;
    }
  }
  cudaEvent_t start_ev = nullptr;
  if (rc == MFB_OK && cudaEventCreateWithFlags(&start_ev, cudaEventDisableTiming) != cudaSuccess) rc = MFB_E_CUDA;
  if (rc == MFB_OK) {
    cudaEventRecord(c->ev0, c->stream);
    cudaEventRecord(start_ev, c->stream);
    cudaStreamWaitEvent(c->copy_stream, start_ev, 0);  // the copies must not run ahead of work queued earlier
  }

  // the decode of the next chunk runs on host threads while this one is copied and updated
  Decoded cur, nxt;
  std::future<void> pending;
  if (rc == MFB_OK && nchunks > 0) decode_chunk(frames, chunk_first[0], chunk_first[1], &cur);
  int64_t total_ratings = 0, runs_seen = 0;
  size_t bytes_seen = 0;
  for (size_t k = 0; k < nchunks && rc == MFB_OK; k++) {
    if (k + 1 < nchunks)
      pending = std::async(std::launch::async, decode_chunk, std::cref(frames), chunk_first[k + 1], chunk_first[k + 2], &nxt);
    Slot* s = &slots[k & 1];
    do {
      if (!cur.ok) {
        set_error("%s: malformed mf.Block in frame %zu", path, cur.bad_frame);
        rc = MFB_E_IO;
        break;
      }
      if (cur.nratings > cap_ratings || cur.nruns > cap_runs) {
        set_error("%s: chunk %zu decodes to %lld records in %lld runs: records shorter than protobuf writes them",
                  path, k, (long long)cur.nratings, (long long)cur.nruns);
        rc = MFB_E_IO;
        break;
      }
      // the pinned buffers of this slot are free once the copy that last read them is done; the device tiles once
      // the kernel that last read them is done (the copy stream waits for that)
      if (s->busy) {
        if (cudaEventSynchronize(s->copied) != cudaSuccess) { rc = MFB_E_CUDA; set_error("event sync failed"); break; }
      }
      bool ids_ok = true;
      const double share = stitch_chunk(cur, s, c->nv, c->nu, &ids_ok);
      if (!ids_ok) {
        set_error("%s: uid or vid outside [0,%d) / [0,%d)", path, c->nu, c->nv);
        rc = MFB_E_ARG;
        break;
      }
      if (cur.nratings == 0) break;
      if (s->busy) cudaStreamWaitEvent(c->copy_stream, s->computed, 0);
      cudaMemcpyAsync(s->d_run_uid, s->h_run_uid, cur.nruns * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream);
      cudaMemcpyAsync(s->d_run_off, s->h_run_off, (cur.nruns + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream);
      cudaMemcpyAsync(s->d_vid, s->h_vid, cur.nratings * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream);
      cudaMemcpyAsync(s->d_rating, s->h_rating, cur.nratings * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream);
      c->h2d_bytes += cur.nratings * 8 + cur.nruns * 8 + 4;
      cudaEventRecord(s->copied, c->copy_stream);
      cudaStreamWaitEvent(c->stream, s->copied, 0);
      // a view of the chunk as a dataset; the bounds on concurrency refer to the whole file, whose run count is
      // extrapolated from what has been decoded so far
      bytes_seen += (size_t)(frames[chunk_first[k + 1] - 1].data + frames[chunk_first[k + 1] - 1].size - frames[chunk_first[k]].data);
      runs_seen += cur.nruns;
      const int64_t est_total_runs = (int64_t)((double)runs_seen * (double)size / (double)std::max<size_t>(bytes_seen, 1));
      Dataset view;  // (owns nothing: the device pointers are the slot's)
      view.used = view.finalized = true;
      view.nruns = std::max<int64_t>(est_total_runs, cur.nruns);  // read for the run bound only; the range is explicit
      view.nratings = cur.nratings;
      view.d_run_uid = s->d_run_uid;
      view.d_run_off = s->d_run_off;
      view.d_vid = s->d_vid;
      view.d_rating = s->d_rating;
      view.max_item_share = share;
      rc = launch_sgd(c, &view, eta, lambda, gb, mode, 0, cur.nruns);
      cudaEventRecord(s->computed, c->stream);
      s->busy = true;
      total_ratings += cur.nratings;
    } while (0);
    if (k + 1 < nchunks) {
      pending.get();
      std::swap(cur, nxt);
    }
  }
  if (pending.valid()) pending.get();
  cudaEventRecord(c->ev1, c->stream);
  c->timed = true;
  // the buffers die here: everything queued on them must have run
  cudaStreamSynchronize(c->copy_stream);
  cudaStreamSynchronize(c->stream);
  if (rc == MFB_OK && cudaGetLastError() != cudaSuccess) {
    set_error("CUDA error during the streamed epoch");
    rc = MFB_E_CUDA;
  }
  for (int b = 0; b < 2; b++) free_slot(&slots[b]);
  if (start_ev) cudaEventDestroy(start_ev);
  munmap(map, size);
  if (rc == MFB_OK) {
    c->model_age++;
    if (ratings_out) *ratings_out = total_ratings;
  }
  return rc;
}
