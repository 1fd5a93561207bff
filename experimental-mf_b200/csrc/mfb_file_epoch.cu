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
#include <chrono>
#include <future>
#include <thread>
#include <vector>

#include "mfb_internal.h"
#include "mfb_wire_decode.h"
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

static int epoch_from_file_device(Context* c, const char* path, float eta, float lambda, float gb, int mode,
                                  int64_t tile_ratings, int64_t* ratings_out, Dataset* keep = nullptr, bool do_epoch = true);

// the chunks decoded by the host cores (option file_decode = 0; the round-2 path, kept for comparison)
static int epoch_from_file_host(Context* c, const char* path, float eta, float lambda, float gb, int mode,
                                int64_t tile_ratings, int64_t* ratings_out) {

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

// ---- the same epoch with the records decoded on the GPU (option file_decode = 1, the default) ---------------------
//
//   file (page cache) --pread by the host cores, frame by frame--> PINNED raw bytes of the chunk
//        + the byte range of every serialized mf.User (one jump per user: the host only walks the top-level fields)
//   --cudaMemcpyAsync (copy stream)--> device raw buffer --wire_{count,scan,decode}_kernel (copy stream)--> SoA tiles
//   --epoch kernel (compute streams; consecutive chunks alternate over two of them at half the width each, so the
//     tail of one launch overlaps the body of the next).
// Three slots (pinned raw, device raw, device tiles) are kept in the context between epochs.  The host touches each
// byte of the file once (the copy out of the page cache); the 100 M varints of a Netflix-sized file are read by the GPU.
namespace mfb {
namespace {

struct FSlot {
  uint8_t* h_raw = nullptr;
  int32_t* h_span = nullptr;  // [2 * cap_runs] begin, end of every user inside the raw buffer
  WireResult* h_res = nullptr;
  uint8_t* d_raw = nullptr;
  int32_t *d_span = nullptr, *d_run_uid = nullptr, *d_run_off = nullptr, *d_count = nullptr, *d_vid = nullptr;
  float* d_rating = nullptr;
  int32_t* d_hist = nullptr;
  WireResult* d_res = nullptr;
  size_t cap_bytes = 0;
  int64_t cap_runs = 0, dev_cap_runs = 0, cap_ratings = 0;  // (span table on the host / run arrays on the device)
  cudaEvent_t copied = nullptr, decode_begin = nullptr, decoded = nullptr, computed = nullptr;
  bool copy_pending = false, decode_pending = false, compute_pending = false;
  // the chunk the host stage left in the pinned buffers
  std::vector<int64_t> frame_runs;  // serialized users per frame of the chunk (the Blocks of the file)
  int64_t nruns = 0;
  size_t nbytes = 0;
  bool ok = true;
  size_t bad_frame = 0;
};

struct FilePipe {
  FSlot slot[3];
  // frame table of the file the pipe was last used on
  std::string path;
  int64_t size = -1, mtime_ns = -1;
  std::vector<int64_t> frame_off;  // offset of every frame's payload; frame_off[i] - 4 is its header
  std::vector<uint32_t> frame_size;
  cudaEvent_t start = nullptr;
  cudaStream_t decode_stream = nullptr;  // the decode kernels of chunk k run while the raw bytes of chunk k+1 are copied
};

void free_fslot(FSlot* s) {
  cudaFreeHost(s->h_raw); cudaFreeHost(s->h_span); cudaFreeHost(s->h_res);
  cudaFree(s->d_raw); cudaFree(s->d_span); cudaFree(s->d_run_uid); cudaFree(s->d_run_off); cudaFree(s->d_count);
  cudaFree(s->d_vid); cudaFree(s->d_rating); cudaFree(s->d_hist); cudaFree(s->d_res);
  if (s->copied) cudaEventDestroy(s->copied);
  if (s->decoded) cudaEventDestroy(s->decoded);
  if (s->decode_begin) cudaEventDestroy(s->decode_begin);
  if (s->computed) cudaEventDestroy(s->computed);
  *s = FSlot();
}

int ensure_bytes(FSlot* s, size_t bytes, int64_t ratings, int nv) {
  if (!s->copied) {
    MFB_CUDA(cudaEventCreateWithFlags(&s->copied, cudaEventDisableTiming));
    MFB_CUDA(cudaEventCreate(&s->decoded));  // (timed: MFB_FILE_TIMING reports the decode kernels)
    MFB_CUDA(cudaEventCreate(&s->decode_begin));
    MFB_CUDA(cudaEventCreateWithFlags(&s->computed, cudaEventDisableTiming));
    MFB_CUDA(cudaMallocHost(&s->h_res, sizeof(WireResult)));
    MFB_CUDA(cudaMalloc(&s->d_res, sizeof(WireResult)));
    MFB_CUDA(cudaMalloc(&s->d_hist, (size_t)std::max(nv, 1) * sizeof(int32_t)));
  }
  if (bytes > s->cap_bytes) {
    if (s->decode_pending) MFB_CUDA(cudaEventSynchronize(s->decoded));
    cudaFreeHost(s->h_raw); cudaFree(s->d_raw);
    s->h_raw = nullptr; s->d_raw = nullptr; s->cap_bytes = 0;
    MFB_CUDA(cudaMallocHost(&s->h_raw, bytes + 64));
    MFB_CUDA(cudaMalloc(&s->d_raw, bytes + 64));  // the decoder reads whole 8-byte words, up to 16 bytes past the data
    MFB_CUDA(cudaMemset(s->d_raw + bytes, 0, 64));
    s->cap_bytes = bytes;
  }
  if (ratings > s->cap_ratings) {
    if (s->compute_pending) MFB_CUDA(cudaEventSynchronize(s->computed));
    cudaFree(s->d_vid); cudaFree(s->d_rating);
    s->d_vid = nullptr; s->d_rating = nullptr; s->cap_ratings = 0;
    MFB_CUDA(cudaMalloc(&s->d_vid, ratings * sizeof(int32_t)));
    MFB_CUDA(cudaMalloc(&s->d_rating, ratings * sizeof(float)));
    s->cap_ratings = ratings;
  }
  return MFB_OK;
}

// host side of the span table only (called from the worker thread of stage A: no device work pending on it there)
int ensure_runs_host(FSlot* s, int64_t runs) {
  if (runs <= s->cap_runs) return MFB_OK;
  const int64_t cap = std::max<int64_t>(runs + runs / 4, 1 << 16);
  cudaFreeHost(s->h_span);
  s->h_span = nullptr;
  s->cap_runs = 0;
  MFB_CUDA(cudaMallocHost(&s->h_span, 2 * cap * sizeof(int32_t)));
  // the device arrays follow in stage B (ensure_runs_device), on the thread that owns the streams
  s->cap_runs = cap;
  return MFB_OK;
}
int ensure_runs_device(FSlot* s) {
  if (s->dev_cap_runs >= s->cap_runs) return MFB_OK;
  if (s->compute_pending) MFB_CUDA(cudaEventSynchronize(s->computed));
  cudaFree(s->d_span); cudaFree(s->d_run_uid); cudaFree(s->d_run_off); cudaFree(s->d_count);
  s->d_span = s->d_run_uid = s->d_run_off = s->d_count = nullptr;
  s->dev_cap_runs = 0;
  MFB_CUDA(cudaMalloc(&s->d_span, 2 * s->cap_runs * sizeof(int32_t)));
  MFB_CUDA(cudaMalloc(&s->d_run_uid, s->cap_runs * sizeof(int32_t)));
  MFB_CUDA(cudaMalloc(&s->d_run_off, (s->cap_runs + 1) * sizeof(int32_t)));
  MFB_CUDA(cudaMalloc(&s->d_count, s->cap_runs * sizeof(int32_t)));
  s->dev_cap_runs = s->cap_runs;
  return MFB_OK;
}

// top-level fields of one serialized mf.Block: the byte range of every User (field 1, length-delimited;
// blocks.proto:14-16), anything else skipped by wire type.  Offsets are relative to `base`.
bool walk_block(const uint8_t* p, size_t size, int64_t base, std::vector<int32_t>* spans) {
  const uint8_t* const begin = p;
  const uint8_t* const end = p + size;
  auto varint = [&](uint64_t* v) -> bool {
    *v = 0;
    for (int i = 0; i < 10 && p < end; i++) {
      const uint8_t b = *p++;
      *v |= (uint64_t)(b & 0x7f) << (7 * i);
      if (!(b & 0x80)) return true;
    }
    return false;
  };
  while (p < end) {
    uint64_t tag, v;
    if (!varint(&tag)) return false;
    if (tag == 0x0A) {
      if (!varint(&v) || v > (uint64_t)(end - p)) return false;
      spans->push_back((int32_t)(base + (p - begin)));
      spans->push_back((int32_t)(base + (p - begin) + (int64_t)v));
      p += v;
      continue;
    }
    switch (tag & 7) {
      case 0: if (!varint(&v)) return false; break;
      case 1: if (end - p < 8) return false; p += 8; break;
      case 2: if (!varint(&v) || v > (uint64_t)(end - p)) return false; p += v; break;
      case 5: if (end - p < 4) return false; p += 4; break;
      default: return false;
    }
  }
  return true;
}

// Stage A of one chunk (frames [f0, f1) of the file): the workers take frames one by one - pread into the pinned raw
// buffer, walk the users while the bytes are still in the core's cache - then the spans are laid out in file order.
void stage_host(FilePipe* fp, int device, int fd, size_t f0, size_t f1, FSlot* s) {
  cudaSetDevice(device);  // (a fresh thread: the event wait and the pinned allocation below need the context's device)
  s->ok = true;
  s->nruns = 0;
  const int64_t byte0 = fp->frame_off[f0] - 4;
  s->nbytes = (size_t)(fp->frame_off[f1 - 1] + fp->frame_size[f1 - 1] - byte0);
  if (s->copy_pending) {  // the pinned buffers are free once the copy that last read them is done
    cudaEventSynchronize(s->copied);
    s->copy_pending = false;
  }
  const size_t nf = f1 - f0;
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nthreads = std::min<size_t>(hw, nf);
  std::vector<std::vector<int32_t>> spans(nf);
  std::vector<char> ok(nf, 1);
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nf) break;
      const size_t f = f0 + i;
      const int64_t off = fp->frame_off[f] - byte0;  // of the payload inside the raw buffer
      size_t got = 0;
      const size_t want = fp->frame_size[f];
      while (got < want) {
        const ssize_t n = pread(fd, s->h_raw + off + got, want - got, fp->frame_off[f] + (int64_t)got);
        if (n <= 0) break;
        got += (size_t)n;
      }
      spans[i].reserve(2048);
      if (got != want || !walk_block(s->h_raw + off, want, off, &spans[i])) ok[i] = 0;
    }
  };
  std::vector<std::thread> pool;
  for (size_t t = 1; t < nthreads; t++) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  int64_t runs = 0;
  s->frame_runs.assign(nf, 0);
  for (size_t i = 0; i < nf; i++) {
    if (!ok[i] && s->ok) {
      s->ok = false;
      s->bad_frame = f0 + i;
    }
    s->frame_runs[i] = (int64_t)spans[i].size() / 2;
    runs += s->frame_runs[i];
  }
  if (!s->ok) return;
  if (ensure_runs_host(s, runs) != MFB_OK) {
    s->ok = false;
    s->bad_frame = (size_t)-1;
    return;
  }
  int64_t at = 0;
  for (size_t i = 0; i < nf; i++) {
    if (!spans[i].empty()) memcpy(s->h_span + at, spans[i].data(), spans[i].size() * sizeof(int32_t));
    at += (int64_t)spans[i].size();
  }
  s->nruns = runs;
}

}  // namespace

void free_file_pipe(Context* c) {
  FilePipe* fp = (FilePipe*)c->file_pipe;
  if (!fp) return;
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->stream2) cudaStreamSynchronize(c->stream2);
  if (fp->decode_stream) cudaStreamSynchronize(fp->decode_stream);
  for (auto& s : fp->slot) free_fslot(&s);
  if (fp->start) cudaEventDestroy(fp->start);
  if (fp->decode_stream) {
    cudaStreamSynchronize(fp->decode_stream);
    cudaStreamDestroy(fp->decode_stream);
  }
  delete fp;
  c->file_pipe = nullptr;
}

}  // namespace mfb

// streaming ingest: the run offsets of a chunk, shifted by the records already kept; per-item counts accumulated
__global__ void keep_offsets_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, int n, int32_t base) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i] + base;
}
__global__ void keep_hist_kernel(int32_t* __restrict__ total, const int32_t* __restrict__ chunk, int n) {
  // (consecutive chunks run on two streams: their additions may meet)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (chunk[i]) atomicAdd(total + i, chunk[i]);
}
__global__ void keep_max_kernel(const int32_t* __restrict__ hist, int n, int32_t* out) {
  int m = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = max(m, hist[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// keep != NULL: STREAMING INGEST: what the chunks decode to is also appended to the resident tiles of
// `keep` (device to device, behind the chunk's update kernel), so that after this one pass over the file the dataset is
// finalized in HBM - the one-time parse + upload stall of mfb_dataset_load_file / finalize disappears into the first
// epoch.  do_epoch = false: ingest only (no update kernels).
static int epoch_from_file_device(Context* c, const char* path, float eta, float lambda, float gb, int mode,
                                  int64_t tile_ratings, int64_t* ratings_out, Dataset* keep, bool do_epoch) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) {
    set_error("cannot open %s", path);
    return MFB_E_IO;
  }
  struct Closer {
    int fd;
    ~Closer() { close(fd); }
  } closer{fd};
  struct stat st;
  if (fstat(fd, &st) != 0) {
    set_error("cannot stat %s", path);
    return MFB_E_IO;
  }
  const int64_t size = (int64_t)st.st_size;
  if (size == 0) return MFB_OK;
  if (!c->file_pipe) c->file_pipe = new FilePipe();
  FilePipe* fp = (FilePipe*)c->file_pipe;
  const int64_t mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000LL + st.st_mtim.tv_nsec;
  if (fp->path != path || fp->size != size || fp->mtime_ns != mtime_ns) {
    // frame boundaries: one 4-byte read per frame (util.h:81)
    fp->frame_off.clear();
    fp->frame_size.clear();
    fp->size = -1;
    int64_t p = 0;
    while (size - p >= 4) {
      uint32_t isize;
      if (pread(fd, &isize, 4, p) != 4) {
        set_error("%s: read error at offset %lld", path, (long long)p);
        return MFB_E_IO;
      }
      p += 4;
      if ((int64_t)isize > size - p) {
        set_error("%s: truncated frame (%u bytes wanted, %lld left)", path, isize, (long long)(size - p));
        return MFB_E_IO;
      }
      fp->frame_off.push_back(p);
      fp->frame_size.push_back(isize);
      p += isize;
    }
    fp->path = path;
    fp->size = size;
    fp->mtime_ns = mtime_ns;
  }
  const size_t nframes = fp->frame_off.size();
  // chunks = runs of whole frames whose bytes / 9 stay within tile_ratings (see the host path for the bound)
  std::vector<size_t> chunk_first{0};
  size_t acc = 0, biggest = 0;
  for (size_t i = 0; i < nframes; i++) {
    const size_t fb = (size_t)fp->frame_size[i] + 4;
    if (acc > 0 && (acc + fb) / 9 + 1 > (size_t)tile_ratings) {
      chunk_first.push_back(i);
      biggest = std::max(biggest, acc);
      acc = 0;
    }
    acc += fb;
  }
  biggest = std::max(biggest, acc);
  chunk_first.push_back(nframes);
  const size_t nchunks = nframes ? chunk_first.size() - 1 : 0;
  MFB_REQUIRE(biggest < (size_t)INT32_MAX - 64, "tile too large for int32 offsets");
  const int64_t cap_ratings = (int64_t)(biggest / 9 + 16);
  for (auto& s : fp->slot)
    if (int rc = ensure_bytes(&s, biggest, cap_ratings, c->nv)) return rc;
  if (!c->copy_stream) MFB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  if (!c->stream2) {
    MFB_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    MFB_CUDA(cudaEventCreateWithFlags(&c->ev_s2, cudaEventDisableTiming));
  }
  if (!fp->start) MFB_CUDA(cudaEventCreateWithFlags(&fp->start, cudaEventDisableTiming));
  if (!fp->decode_stream) MFB_CUDA(cudaStreamCreateWithFlags(&fp->decode_stream, cudaStreamNonBlocking));
  cudaEventRecord(c->ev0, c->stream);
  // neither the copy stream nor the second compute stream may run ahead of work queued earlier
  MFB_CUDA(cudaEventRecord(fp->start, c->stream));
  MFB_CUDA(cudaStreamWaitEvent(c->copy_stream, fp->start, 0));
  MFB_CUDA(cudaStreamWaitEvent(c->stream2, fp->start, 0));
  MFB_CUDA(cudaStreamWaitEvent(fp->decode_stream, fp->start, 0));
  cudaStream_t const main_stream = c->stream;
  const bool two = c->opt_two_streams != 0 && nchunks > 1 && mode != MFB_MODE_ORDERED;
  // resident tiles of the dataset being ingested: records bounded by bytes / 9 (grown if the file has shorter ones)
  int64_t keep_runs = 0, keep_recs = 0, cap_keep_runs = 0, cap_keep_recs = 0;
  int32_t* d_ghist = nullptr;  // [nv + 1]: records per item over the whole file, then their maximum
  auto grow = [&](void** ptr, size_t elem, int64_t used, int64_t cap, int64_t want) -> int {
    void* q = nullptr;
    MFB_CUDA(cudaDeviceSynchronize());  // (rare: kernels of earlier chunks may still read the old array)
    MFB_CUDA(cudaMalloc(&q, (size_t)want * elem));
    if (*ptr && used > 0) MFB_CUDA(cudaMemcpy(q, *ptr, (size_t)used * elem, cudaMemcpyDeviceToDevice));
    (void)cap;
    cudaFree(*ptr);
    *ptr = q;
    return MFB_OK;
  };
  if (keep) {
    cap_keep_recs = size / 9 + 16;
    cap_keep_runs = std::max<int64_t>(size / 96, 1 << 16);
    MFB_CUDA(cudaMalloc(&keep->d_vid, cap_keep_recs * sizeof(int32_t)));
    MFB_CUDA(cudaMalloc(&keep->d_rating, cap_keep_recs * sizeof(float)));
    MFB_CUDA(cudaMalloc(&keep->d_run_uid, cap_keep_runs * sizeof(int32_t)));
    MFB_CUDA(cudaMalloc(&keep->d_run_off, (cap_keep_runs + 1) * sizeof(int32_t)));
    MFB_CUDA(cudaMemsetAsync(keep->d_run_off, 0, sizeof(int32_t), c->stream));
    MFB_CUDA(cudaMalloc(&d_ghist, ((size_t)c->nv + 1) * sizeof(int32_t)));
    MFB_CUDA(cudaMemsetAsync(d_ghist, 0, ((size_t)c->nv + 1) * sizeof(int32_t), c->stream));
    keep->h_block_off.assign(1, 0);
  }
  std::future<void> staged[3];
  auto stage = [&](size_t k) {
    staged[k % 3] = std::async(std::launch::async, stage_host, fp, c->device, fd, chunk_first[k], chunk_first[k + 1], &fp->slot[k % 3]);
  };
  int rc = MFB_OK;
  // stage B: H2D of the raw bytes and spans (copy stream), then the decode kernels and the result back (decode stream)
  auto issue = [&](size_t k) -> int {
    FSlot* s = &fp->slot[k % 3];
    staged[k % 3].get();
    if (!s->ok) {
      if (s->bad_frame == (size_t)-1) return MFB_E_CUDA;
      set_error("%s: malformed mf.Block in frame %zu", path, s->bad_frame);
      return MFB_E_IO;
    }
    if (int e = ensure_runs_device(s)) return e;  // as large as the host span table
    // the device raw buffer of this slot is free once the decode that last read it is done
    if (s->decode_pending) MFB_CUDA(cudaStreamWaitEvent(c->copy_stream, s->decoded, 0));
    MFB_CUDA(cudaMemcpyAsync(s->d_raw, s->h_raw, s->nbytes, cudaMemcpyHostToDevice, c->copy_stream));
    if (s->nruns)
      MFB_CUDA(cudaMemcpyAsync(s->d_span, s->h_span, 2 * s->nruns * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream));
    MFB_CUDA(cudaEventRecord(s->copied, c->copy_stream));
    s->copy_pending = true;
    c->h2d_bytes += (int64_t)s->nbytes + 2 * s->nruns * (int64_t)sizeof(int32_t);
    // ... the tiles once the kernel that last read them is done
    cudaStream_t ds = fp->decode_stream;
    MFB_CUDA(cudaStreamWaitEvent(ds, s->copied, 0));
    if (s->compute_pending) MFB_CUDA(cudaStreamWaitEvent(ds, s->computed, 0));
    MFB_CUDA(cudaEventRecord(s->decode_begin, ds));
    if (int e = launch_wire_decode(c, ds, s->d_raw, s->d_span, (int)s->nruns, s->cap_ratings, s->d_run_uid, s->d_run_off,
                                   s->d_count, s->d_vid, s->d_rating, s->d_hist, s->d_res))
      return e;
    MFB_CUDA(cudaMemcpyAsync(s->h_res, s->d_res, 4 * sizeof(int32_t) + sizeof(long long), cudaMemcpyDeviceToHost, ds));
    MFB_CUDA(cudaEventRecord(s->decoded, ds));
    s->decode_pending = true;
    return MFB_OK;
  };

  int64_t total_ratings = 0, runs_seen = 0, bytes_seen = 0;
  // MFB_FILE_TIMING=1: where the host thread of this call waits (stderr, one line per epoch)
  const bool clock_on = getenv("MFB_FILE_TIMING") != nullptr;
  double t_stage = 0, t_decode = 0, dev_decode_ms = 0;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  if (nchunks > 0) stage(0);
  if (nchunks > 1) stage(1);
  for (size_t k = 0; k < nchunks && rc == MFB_OK; k++) {
    double t0 = now();
    if (k == 0) rc = issue(0);
    if (rc == MFB_OK && k + 1 < nchunks) rc = issue(k + 1);  // queued before this thread blocks on chunk k's decode
    if (rc == MFB_OK && k + 2 < nchunks) stage(k + 2);
    if (rc != MFB_OK) break;
    t_stage += now() - t0;
    t0 = now();
    FSlot* s = &fp->slot[k % 3];
    const cudaError_t de = cudaEventSynchronize(s->decoded);
    t_decode += now() - t0;
    if (clock_on && de == cudaSuccess) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, s->decode_begin, s->decoded) == cudaSuccess) dev_decode_ms += ms;
    }
    if (de != cudaSuccess) {
      set_error("%s: decode of chunk %zu failed: %s", path, k, cudaGetErrorString(cudaGetLastError()));
      rc = MFB_E_CUDA;
      break;
    }
    const WireResult res = *s->h_res;
    if (res.err) {
      const long long frame_lo = (long long)chunk_first[k], frame_hi = (long long)chunk_first[k + 1];
      if (res.err == WIRE_E_UID || res.err == WIRE_E_VID) {
        set_error("%s: uid or vid outside [0,%d) / [0,%d) (frames %lld..%lld, user-run %d of the chunk)", path, c->nu, c->nv,
                  frame_lo, frame_hi - 1, res.err_run);
        rc = MFB_E_ARG;
      } else if (res.err == WIRE_E_CAPACITY) {
        set_error("%s: chunk %zu decodes to %lld records: records shorter than protobuf writes them", path, k, res.nratings);
        rc = MFB_E_IO;
      } else {
        set_error("%s: malformed mf.User in frames %lld..%lld (user-run %d of the chunk)", path, frame_lo, frame_hi - 1, res.err_run);
        rc = MFB_E_IO;
      }
      break;
    }
    const bool second = two && (k & 1) == 1;
    cudaStream_t const kstream = second ? c->stream2 : main_stream;  // where this chunk's kernels go
    if (keep) {
      // append what the chunk decoded to: user ids, shifted run offsets, records, per-item counts
      if (keep_runs + s->nruns > cap_keep_runs) {
        const int64_t want = std::max(keep_runs + s->nruns, 2 * cap_keep_runs);
        if ((rc = grow((void**)&keep->d_run_uid, sizeof(int32_t), keep_runs, cap_keep_runs, want)) ||
            (rc = grow((void**)&keep->d_run_off, sizeof(int32_t), keep_runs + 1, cap_keep_runs + 1, want + 1)))
          break;
        cap_keep_runs = want;
      }
      if (keep_recs + res.nratings > cap_keep_recs) {
        const int64_t want = std::max<int64_t>(keep_recs + res.nratings, 2 * cap_keep_recs);
        if ((rc = grow((void**)&keep->d_vid, sizeof(int32_t), keep_recs, cap_keep_recs, want)) ||
            (rc = grow((void**)&keep->d_rating, sizeof(float), keep_recs, cap_keep_recs, want)))
          break;
        cap_keep_recs = want;
      }
      if (keep_recs + res.nratings >= (long long)INT32_MAX) {
        set_error("%s: more than 2^31 records", path);
        rc = MFB_E_IO;
        break;
      }
      cudaStreamWaitEvent(kstream, s->decoded, 0);
      if (s->nruns) {
        cudaMemcpyAsync(keep->d_run_uid + keep_runs, s->d_run_uid, s->nruns * sizeof(int32_t), cudaMemcpyDeviceToDevice, kstream);
        keep_offsets_kernel<<<(int)std::min<int64_t>((s->nruns + 255) / 256, 1024), 256, 0, kstream>>>(
            keep->d_run_off + keep_runs + 1, s->d_run_off + 1, (int)s->nruns, (int32_t)keep_recs);
      }
      if (res.nratings) {
        cudaMemcpyAsync(keep->d_vid + keep_recs, s->d_vid, res.nratings * sizeof(int32_t), cudaMemcpyDeviceToDevice, kstream);
        cudaMemcpyAsync(keep->d_rating + keep_recs, s->d_rating, res.nratings * sizeof(float), cudaMemcpyDeviceToDevice, kstream);
        keep_hist_kernel<<<(c->nv + 255) / 256, 256, 0, kstream>>>(d_ghist, s->d_hist, c->nv);
      }
      c->launches += 2;
      for (int64_t fr : s->frame_runs) keep->h_block_off.push_back(keep->h_block_off.back() + fr);
      keep_runs += s->nruns;
      keep_recs += res.nratings;
      if (res.nratings == 0 || !do_epoch) {  // no update kernel follows: the slot is free once the copies are done
        cudaEventRecord(s->computed, kstream);
        s->compute_pending = true;
      }
    }
    if (res.nratings == 0) continue;
    total_ratings += do_epoch ? 0 : res.nratings;
    if (!do_epoch) continue;
    bytes_seen += (int64_t)s->nbytes;
    runs_seen += s->nruns;
    const int64_t est_total_runs = (int64_t)((double)runs_seen * (double)size / (double)std::max<int64_t>(bytes_seen, 1));
    Dataset view;  // (owns nothing: the device pointers are the slot's)
    view.used = view.finalized = true;
    view.nruns = std::max<int64_t>(est_total_runs, s->nruns);  // read for the run bound only; the range is explicit
    view.nratings = res.nratings;
    view.d_run_uid = s->d_run_uid;
    view.d_run_off = s->d_run_off;
    view.d_vid = s->d_vid;
    view.d_rating = s->d_rating;
    view.max_item_share = (double)res.top_count / (double)res.nratings;
    c->stream = kstream;
    c->counter_slot = second ? 2 : 0;
    c->width_div = two ? 2 : 1;
    cudaError_t e = cudaStreamWaitEvent(c->stream, s->decoded, 0);
    if (e == cudaSuccess) rc = launch_sgd(c, &view, eta, lambda, gb, mode, 0, s->nruns);
    if (e == cudaSuccess && rc == MFB_OK) e = cudaEventRecord(s->computed, c->stream);
    s->compute_pending = true;
    c->stream = main_stream;
    c->counter_slot = 0;
    c->width_div = 1;
    if (e != cudaSuccess) {
      set_error("CUDA error during the streamed epoch: %s", cudaGetErrorString(e));
      rc = MFB_E_CUDA;
    }
    total_ratings += res.nratings;
  }
  for (auto& f : staged)
    if (f.valid()) f.get();
  cudaEventRecord(c->ev_s2, c->stream2);
  cudaStreamWaitEvent(c->stream, c->ev_s2, 0);
  cudaEventRecord(c->ev1, c->stream);
  c->timed = true;
  if (rc == MFB_OK && keep) {
    // the dataset is resident: bring the small run tables to the host (epoch slices, DSGD splits and the streamed
    // epochs read them there) and the share of the most rated item
    keep_max_kernel<<<1, 1024, 0, c->stream>>>(d_ghist, c->nv, d_ghist + c->nv);
    int32_t top = 0;
    keep->h_run_uid.resize((size_t)keep_runs);
    keep->h_run_off.resize((size_t)keep_runs + 1);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(&top, d_ghist + c->nv, sizeof top, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && keep_runs)
      e = cudaMemcpy(keep->h_run_uid.data(), keep->d_run_uid, keep_runs * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
      e = cudaMemcpy(keep->h_run_off.data(), keep->d_run_off, (keep_runs + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
      set_error("streaming ingest: %s", cudaGetErrorString(e));
      rc = MFB_E_CUDA;
    } else {
      keep->nruns = keep_runs;
      keep->nratings = keep_recs;
      keep->nblocks = (int64_t)keep->h_block_off.size() - 1;
      keep->max_item_share = keep_recs ? (double)top / (double)keep_recs : 0.0;
      keep->finalized = true;
    }
  }
  if (rc != MFB_OK) {  // leave nothing queued on buffers whose content is undefined
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(fp->decode_stream);
    cudaStreamSynchronize(c->stream2);
    cudaStreamSynchronize(c->stream);
    cudaFree(d_ghist);
    if (keep) {  // back to an empty, unfinalized dataset
      cudaFree(keep->d_vid); cudaFree(keep->d_rating); cudaFree(keep->d_run_uid); cudaFree(keep->d_run_off);
      keep->d_vid = keep->d_run_uid = keep->d_run_off = nullptr;
      keep->d_rating = nullptr;
      keep->h_run_uid.clear();
      keep->h_run_off.clear();
      keep->h_block_off.clear();
      keep->finalized = false;
    }
    return rc;
  }
  cudaFree(d_ghist);
  if (do_epoch) c->model_age++;
  if (ratings_out) *ratings_out = keep ? keep_recs : total_ratings;
  if (clock_on)
    fprintf(stderr, "mfb_sgd_epoch_from_file: %zu chunks, %.1f MB: host thread %.1f ms = %.1f waiting for the pread/walk stage "
            "+ %.1f waiting for copy+decode + %.1f launching; decode kernels %.1f ms on the device\n", nchunks, size / 1e6,
            1e3 * (now() - t_begin), 1e3 * t_stage, 1e3 * t_decode, 1e3 * (now() - t_begin - t_stage - t_decode), dev_decode_ms);
  return MFB_OK;
}

// The host half of the device-decode path on its own, no GPU involved: walks the frames of a [u32 size][mf.Block]
// file and the top-level fields of every Block the way stage_host does, and reports what it found - frames, serialized
// mf.User messages and the bytes inside them (what the GPU would be handed).  MFB_E_IO on a truncated frame or a
// malformed Block.  Used by the CPU tests; also a cheap "is this a rating file" check for a caller.
extern "C" int mfb_wire_index_file(const char* path, int64_t* nframes, int64_t* nusers, int64_t* user_bytes) {
  MFB_REQUIRE(path && nframes && nusers && user_bytes, "NULL argument");
  *nframes = *nusers = *user_bytes = 0;
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
  const int64_t size = (int64_t)st.st_size;
  std::vector<uint8_t> buf;
  std::vector<int32_t> spans;
  int rc = MFB_OK;
  int64_t p = 0;
  while (size - p >= 4 && rc == MFB_OK) {
    uint32_t isize;
    if (pread(fd, &isize, 4, p) != 4) {
      set_error("%s: read error at offset %lld", path, (long long)p);
      rc = MFB_E_IO;
      break;
    }
    p += 4;
    if ((int64_t)isize > size - p) {
      set_error("%s: truncated frame (%u bytes wanted, %lld left)", path, isize, (long long)(size - p));
      rc = MFB_E_IO;
      break;
    }
    buf.resize(isize);
    if (isize && pread(fd, buf.data(), isize, p) != (ssize_t)isize) {
      set_error("%s: read error at offset %lld", path, (long long)p);
      rc = MFB_E_IO;
      break;
    }
    spans.clear();
    if (!walk_block(buf.data(), isize, 0, &spans)) {
      set_error("%s: malformed mf.Block in frame %lld", path, (long long)*nframes);
      rc = MFB_E_IO;
      break;
    }
    for (size_t i = 0; i + 1 < spans.size(); i += 2) *user_bytes += spans[i + 1] - spans[i];
    *nusers += (int64_t)spans.size() / 2;
    ++*nframes;
    p += isize;
  }
  close(fd);
  return rc;
}

extern "C" int mfb_sgd_epoch_from_file(mfb_ctx* h, const char* path, float eta, float lambda, float gb, int mode,
                                       int64_t tile_ratings, int64_t* ratings_out) {
  MFB_REQUIRE(h && path, "NULL argument");
  MFB_REQUIRE(mode == MFB_MODE_HOGWILD || mode == MFB_MODE_ORDERED || mode == MFB_MODE_ATOMIC, "bad mode %d", mode);
  Context* c = &h->c;
  MFB_CUDA(cudaSetDevice(c->device));
  if (tile_ratings <= 0) tile_ratings = (int64_t)8 << 20;
  tile_ratings = std::max<int64_t>(tile_ratings, 1024);
  if (ratings_out) *ratings_out = 0;
  return c->opt_file_decode ? epoch_from_file_device(c, path, eta, lambda, gb, mode, tile_ratings, ratings_out)
                            : epoch_from_file_host(c, path, eta, lambda, gb, mode, tile_ratings, ratings_out);
}

// SURVEY 8f-2, streaming ingest: one pass over the training file that leaves dataset `ds` finalized in HBM (like
// mfb_dataset_load_file + mfb_dataset_finalize) with the records decoded on the GPU, and - with_epoch != 0 - runs the
// SGD epoch on every chunk as it lands, so the first epoch costs what an out-of-core epoch costs and the parse + upload
// stall is gone.  The dataset must be new (nothing appended).  Not for contexts with the dpmf arrays enabled (the
// static logical clock is built from the host copy of the records: use mfb_dataset_load_file there).
extern "C" int mfb_dataset_ingest_file(mfb_ctx* h, int ds, const char* path, int with_epoch, float eta, float lambda,
                                       float gb, int mode, int64_t tile_ratings, int64_t* ratings_out) {
  MFB_REQUIRE(h && path, "NULL argument");
  MFB_REQUIRE(mode == MFB_MODE_HOGWILD || mode == MFB_MODE_ORDERED || mode == MFB_MODE_ATOMIC, "bad mode %d", mode);
  Context* c = &h->c;
  MFB_REQUIRE(ds >= 0 && ds < (int)c->datasets.size() && c->datasets[ds].used, "bad dataset id %d", ds);
  Dataset* d = &c->datasets[ds];
  MFB_REQUIRE(!d->finalized && d->h_run_uid.empty() && d->h_vid.empty(), "dataset %d is not empty", ds);
  MFB_REQUIRE(!c->arr[MFB_UR], "dpmf is enabled on this context: its logical clock needs mfb_dataset_load_file");
  MFB_CUDA(cudaSetDevice(c->device));
  if (tile_ratings <= 0) tile_ratings = (int64_t)8 << 20;
  tile_ratings = std::max<int64_t>(tile_ratings, 1024);
  if (ratings_out) *ratings_out = 0;
  struct stat st;
  if (stat(path, &st) != 0) {
    set_error("cannot open %s", path);
    return MFB_E_IO;
  }
  if (st.st_size == 0) {  // an empty file is an empty dataset
    d->h_run_off.assign(1, 0);
    d->h_block_off.assign(1, 0);
    MFB_CUDA(cudaMalloc(&d->d_run_off, sizeof(int32_t)));
    MFB_CUDA(cudaMemset(d->d_run_off, 0, sizeof(int32_t)));
    d->finalized = true;
    return MFB_OK;
  }
  return epoch_from_file_device(c, path, eta, lambda, gb, mode, tile_ratings, ratings_out, d, with_epoch != 0);
}
