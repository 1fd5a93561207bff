// blocks.proto wire codec (host).  Replaces the protoc-2.6 generated blocks.pb.{h,cc} + libprotobuf
// on the ingest path: ParseFilter's Block::ParseFromArray (mf.h:57-69), plain_read (util.h:76-88)
// and getdata's writer (getdata.cc:82-126).  The file is a sequence of [u32 size][mf.Block]
// frames; messages (blocks.proto:1-18, tags blocks.pb.cc:267,281,526,540,786):
//   Block{ repeated User user = 1 (0x0A) }
//   User { int32 uid = 1 (0x08);  repeated Record record = 2 (0x12) }
//   Record{ int32 vid = 1 (0x08); float rating = 2 (0x15) }
// Decoding goes straight into the flat SoA arrays the kernels consume: no per-record objects.
#include "proto_wire.h"

#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include <chrono>
#define TIMER_T0 auto _t0 = std::chrono::steady_clock::now()
#define TIMER_LAP(name) do { auto _t1 = std::chrono::steady_clock::now(); if (getenv("MFB_TIMING")) fprintf(stderr, "load_blocks_file %s %.3f s\n", name, std::chrono::duration<double>(_t1 - _t0).count()); _t0 = _t1; } while (0)
namespace mfb {
namespace {

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool more() const { return ok && p < end; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (p < end) {
      const uint8_t b = *p++;
      if (shift < 64) v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
      if (shift >= 70) break;  // a varint is at most 10 bytes
    }
    ok = false;
    return 0;
  }
  Cursor sub() {  // length-delimited payload
    const uint64_t len = varint();
    if (!ok || (uint64_t)(end - p) < len) {
      ok = false;
      return Cursor{p, p, false};
    }
    Cursor c{p, p + len, true};
    p += len;
    return c;
  }
  void skip(uint32_t wire_type) {
    switch (wire_type) {
      case 0: varint(); break;
      case 1: if (end - p < 8) ok = false; else p += 8; break;
      case 2: sub(); break;
      case 5: if (end - p < 4) ok = false; else p += 4; break;
      default: ok = false;
    }
  }
};

}  // namespace

bool decode_block(const void* data, size_t size, BlockSink* out) {
  Cursor blk{(const uint8_t*)data, (const uint8_t*)data + size, true};
  out->vid.reserve(out->vid.size() + size / 9);  // a record takes at least 9 bytes on the wire
  out->rating.reserve(out->rating.size() + size / 9);
  while (blk.more()) {
    const uint64_t tag = blk.varint();
    if (tag != 0x0A) {
      blk.skip((uint32_t)(tag & 7));
      continue;
    }
    Cursor usr = blk.sub();
    if (!blk.ok) return false;
    int32_t uid = 0;
    out->uid.push_back(0);
    const size_t slot = out->uid.size() - 1;
    while (usr.more()) {
      // Fast path for the canonical record protobuf's encoder writes (blocks.pb.cc:267,281):
      // 12 LL 08 <vid varint, 1..3 bytes> 15 <4 bytes>; anything else goes the general way below.
      if (usr.end - usr.p >= 11 && usr.p[0] == 0x12 && usr.p[2] == 0x08) {
        const uint8_t* q = usr.p;
        const uint32_t len = q[1];
        uint32_t v = q[3];
        uint32_t vl = 1;
        if (v & 0x80) {
          v = (v & 0x7F) | ((uint32_t)q[4] << 7);
          vl = 2;
          if (v & (0x80u << 7)) {
            v = (v & 0x3FFF) | ((uint32_t)q[5] << 14);
            vl = 3;
          }
        }
        if (len == vl + 6 && !(v & (0x80u << (7 * (vl - 1)))) && q[3 + vl] == 0x15) {
          float rating;
          memcpy(&rating, q + 4 + vl, 4);
          out->vid.push_back((int32_t)v);
          out->rating.push_back(rating);
          usr.p = q + 2 + len;
          continue;
        }
      }
      const uint64_t t = usr.varint();
      if (t == 0x08) {
        uid = (int32_t)(uint32_t)usr.varint();
      } else if (t == 0x12) {
        Cursor rec = usr.sub();
        if (!usr.ok) return false;
        int32_t vid = 0;
        float rating = 0.f;
        while (rec.more()) {
          const uint64_t rt = rec.varint();
          if (rt == 0x08) {
            vid = (int32_t)(uint32_t)rec.varint();
          } else if (rt == 0x15) {
            if (rec.end - rec.p < 4) return false;
            memcpy(&rating, rec.p, 4);
            rec.p += 4;
          } else {
            rec.skip((uint32_t)(rt & 7));
          }
        }
        if (!rec.ok) return false;
        out->vid.push_back(vid);
        out->rating.push_back(rating);
      } else {
        usr.skip((uint32_t)(t & 7));
      }
    }
    if (!usr.ok) return false;
    out->uid[slot] = uid;
    out->rec_off.push_back((int32_t)out->vid.size());
  }
  return blk.ok;
}

static inline size_t varint_len(uint64_t v) {
  size_t n = 1;
  while (v >= 0x80) { v >>= 7; n++; }
  return n;
}
static inline void put_varint(std::string* s, uint64_t v) {
  while (v >= 0x80) { s->push_back((char)((v & 0x7F) | 0x80)); v >>= 7; }
  s->push_back((char)v);
}

void encode_block(int32_t nusers, const int32_t* uid, const int32_t* rec_off, const int32_t* vid,
                  const float* rating, std::string* out) {
  out->clear();
  for (int32_t i = 0; i < nusers; i++) {
    size_t body = 1 + varint_len((uint64_t)(int64_t)uid[i]);
    for (int32_t k = rec_off[i]; k < rec_off[i + 1]; k++)
      body += 2 + 1 + varint_len((uint64_t)(int64_t)vid[k]) + 5;
    out->push_back((char)0x0A);
    put_varint(out, body);
    out->push_back((char)0x08);
    put_varint(out, (uint64_t)(int64_t)uid[i]);
    for (int32_t k = rec_off[i]; k < rec_off[i + 1]; k++) {
      out->push_back((char)0x12);
      out->push_back((char)(1 + varint_len((uint64_t)(int64_t)vid[k]) + 5));
      out->push_back((char)0x08);
      put_varint(out, (uint64_t)(int64_t)vid[k]);
      out->push_back((char)0x15);
      char b[4];
      memcpy(b, &rating[k], 4);
      out->append(b, 4);
    }
  }
}

int for_each_frame(const char* path, const std::function<bool(const void*, size_t)>& fn) {
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
  const uint8_t* p = (const uint8_t*)map;
  const uint8_t* end = p + size;
  int rc = MFB_OK;
  while (end - p >= 4) {  // util.h:81 `while(fread(&isize, 1, sizeof(isize), fr))`
    uint32_t isize;
    memcpy(&isize, p, 4);
    p += 4;
    if ((size_t)(end - p) < isize) {
      set_error("%s: truncated frame (%u bytes wanted, %zu left)", path, isize, (size_t)(end - p));
      rc = MFB_E_IO;
      break;
    }
    if (!fn(p, isize)) {
      set_error("%s: malformed mf.Block at offset %zu", path, (size_t)(p - (const uint8_t*)map));
      rc = MFB_E_IO;
      break;
    }
    p += isize;
  }
  munmap(map, size);
  return rc;
}

// Parse a whole file: the frame boundaries come from one sequential walk over the length prefixes
// (a jump per frame), the frames are decoded by worker threads (the reference decodes one Block
// per ParseFilter call on a TBB worker, mf.h:57-69, and does so again every epoch; here it happens
// once), and the per-frame results are stitched together in file order.
int load_blocks_file(const char* path, Dataset* d) {
  if (d->h_run_off.empty()) d->h_run_off.push_back(0);
  if (d->h_block_off.empty()) d->h_block_off.push_back(0);
  struct Frame {
    const void* data;
    size_t size;
  };
  std::vector<Frame> frames;
  // for_each_frame unmaps the file when it returns, so the decoding happens inside the callback of a
  // second walk; the first walk only counts (cheap: one 4-byte read per frame)
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
  const uint8_t* p = (const uint8_t*)map;
  const uint8_t* end = p + size;
  while (end - p >= 4) {  // util.h:81 `while(fread(&isize, 1, sizeof(isize), fr))`
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
  const size_t nf = frames.size();
  TIMER_T0;
  std::vector<BlockSink> sinks(nf);
  std::vector<char> ok(nf, 1);
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nthreads = std::min<size_t>(hw, std::max<size_t>(1, nf / 4));
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nf) break;
      sinks[i].rec_off.assign(1, 0);
      ok[i] = decode_block(frames[i].data, frames[i].size, &sinks[i]) ? 1 : 0;
    }
  };
  {
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nthreads; t++) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
  }
  TIMER_LAP("decode");
  int rc = MFB_OK;
  int64_t total = (int64_t)d->h_vid.size(), runs = (int64_t)d->h_run_uid.size();
  for (size_t i = 0; i < nf && rc == MFB_OK; i++) {
    if (!ok[i]) {
      set_error("%s: malformed mf.Block at offset %zu", path, (size_t)((const uint8_t*)frames[i].data - (const uint8_t*)map));
      rc = MFB_E_IO;
    }
    total += (int64_t)sinks[i].vid.size();
    runs += (int64_t)sinks[i].uid.size();
    if (total >= (int64_t)INT32_MAX) {
      set_error("%s: more than 2^31 records", path);
      rc = MFB_E_IO;
    }
  }
  munmap(map, size);
  if (rc) return rc;
  // stitch: offsets by prefix sum, payload copied by the same workers
  std::vector<int64_t> rec0(nf + 1), run0(nf + 1);
  rec0[0] = (int64_t)d->h_vid.size();
  run0[0] = (int64_t)d->h_run_uid.size();
  for (size_t i = 0; i < nf; i++) {
    rec0[i + 1] = rec0[i] + (int64_t)sinks[i].vid.size();
    run0[i + 1] = run0[i] + (int64_t)sinks[i].uid.size();
  }
  TIMER_LAP("check+munmap");
  // first touch of the two big arrays by all threads (a serial resize() spends more time in page
  // faults than the whole decode): reserve, pre-fault the reserved pages in parallel, then resize
  d->h_vid.reserve(rec0[nf]);
  d->h_rating.reserve(rec0[nf]);
  {
    auto prefault = [&](void* base, size_t bytes, size_t t) {
      const uintptr_t lo = ((uintptr_t)base + 4095) & ~(uintptr_t)4095, hi = ((uintptr_t)base + bytes) & ~(uintptr_t)4095;
      if (hi <= lo) return;
      const size_t pages = (hi - lo) >> 12, p0 = pages * t / nthreads, p1 = pages * (t + 1) / nthreads;
#ifdef MADV_POPULATE_WRITE
      if (p1 > p0) madvise((void*)(lo + (p0 << 12)), (p1 - p0) << 12, MADV_POPULATE_WRITE);  // best effort
#endif
    };
    auto job = [&](size_t t) {
      prefault(d->h_vid.data(), (size_t)rec0[nf] * sizeof(int32_t), t);
      prefault(d->h_rating.data(), (size_t)rec0[nf] * sizeof(float), t);
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nthreads; t++) pool.emplace_back(job, t);
    job(0);
    for (auto& t : pool) t.join();
  }
  TIMER_LAP("prefault");
  d->h_vid.resize(rec0[nf]);
  d->h_rating.resize(rec0[nf]);
  d->h_run_uid.resize(run0[nf]);
  d->h_run_off.resize(run0[nf] + 1);
  const size_t blocks0 = d->h_block_off.size();
  d->h_block_off.resize(blocks0 + nf);
  TIMER_LAP("resize");
  next = 0;
  auto stitch = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nf) break;
      const BlockSink& b = sinks[i];
      if (!b.vid.empty()) {
        memcpy(d->h_vid.data() + rec0[i], b.vid.data(), b.vid.size() * sizeof(int32_t));
        memcpy(d->h_rating.data() + rec0[i], b.rating.data(), b.rating.size() * sizeof(float));
      }
      for (size_t k = 0; k < b.uid.size(); k++) {
        d->h_run_uid[run0[i] + k] = b.uid[k];
        d->h_run_off[run0[i] + k + 1] = (int32_t)(rec0[i] + b.rec_off[k + 1]);
      }
      d->h_block_off[blocks0 + i] = run0[i + 1];
      std::vector<int32_t>().swap(sinks[i].vid);
      std::vector<float>().swap(sinks[i].rating);
    }
  };
  {
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nthreads; t++) pool.emplace_back(stitch);
    stitch();
    for (auto& t : pool) t.join();
  }
  TIMER_LAP("stitch");
  return MFB_OK;
}

}  // namespace mfb
