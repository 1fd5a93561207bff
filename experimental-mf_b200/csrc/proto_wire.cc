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
      if (shift > 63 + 7) break;
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

int load_blocks_file(const char* path, Dataset* d) {
  if (d->h_run_off.empty()) d->h_run_off.push_back(0);
  if (d->h_block_off.empty()) d->h_block_off.push_back(0);
  BlockSink sink;
  return for_each_frame(path, [&](const void* data, size_t size) {
    sink.uid.clear();
    sink.vid.clear();
    sink.rating.clear();
    sink.rec_off.assign(1, 0);
    if (!decode_block(data, size, &sink)) return false;
    const int64_t base = (int64_t)d->h_vid.size();
    if (base + (int64_t)sink.vid.size() >= (int64_t)INT32_MAX) return false;
    for (size_t i = 0; i < sink.uid.size(); i++) {
      d->h_run_uid.push_back(sink.uid[i]);
      d->h_run_off.push_back((int32_t)(base + sink.rec_off[i + 1]));
    }
    d->h_vid.insert(d->h_vid.end(), sink.vid.begin(), sink.vid.end());
    d->h_rating.insert(d->h_rating.end(), sink.rating.begin(), sink.rating.end());
    d->h_block_off.push_back((int64_t)d->h_run_uid.size());
    return true;
  });
}

}  // namespace mfb
