// TEST INFRASTRUCTURE (oracle/): stand-in for the protoc-2.6 generated blocks.pb.h + libprotobuf
// (neither protoc nor libprotobuf exists in this image).  The reference's util.h:16 does
// `#include "blocks.pb.h"`, which resolves next to util.h; oracle/Makefile therefore pre-defines
// that file's include guard (PROTOBUF_blocks_2eproto__INCLUDED) so its body is skipped, and
// force-includes this header instead.  Not product code.
//
// Provides mf::User_Record / mf::User / mf::Block / mf::Blocks with the accessor names the
// reference uses (blocks.pb.h:409-546) and a hand-written decoder/encoder for the wire format of
// blocks.proto:1-18:
//   Block  : repeated User   user   = 1  -> tag 0x0A, length-delimited
//   User   : required int32  uid    = 1  -> tag 0x08, varint
//            repeated Record record = 2  -> tag 0x12, length-delimited
//   Record : required int32  vid    = 1  -> tag 0x08, varint
//            required float  rating = 2  -> tag 0x15, fixed32 little-endian
// (tags as in blocks.pb.cc:267,281,526,540,786).  Unknown fields are skipped by wire type, as
// protobuf does.  Negative int32 are 10-byte sign-extended varints (proto2 int32 rule).
#ifndef ORACLE_SHIM_BLOCKS_H
#define ORACLE_SHIM_BLOCKS_H

#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

namespace mf {

namespace wire {
inline bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t& out) {
  uint64_t v = 0;
  for (int shift = 0; shift < 70 && p < end; shift += 7) {
    uint8_t b = *p++;
    if (shift < 64) v |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      out = v;
      return true;
    }
  }
  return false;
}
inline bool skip_field(uint32_t wt, const uint8_t*& p, const uint8_t* end) {
  uint64_t tmp;
  switch (wt) {
    case 0: return get_varint(p, end, tmp);
    case 1: if (end - p < 8) return false; p += 8; return true;
    case 2: if (!get_varint(p, end, tmp) || (uint64_t)(end - p) < tmp) return false; p += tmp; return true;
    case 5: if (end - p < 4) return false; p += 4; return true;
    default: return false;
  }
}
inline void put_varint(std::string& s, uint64_t v) {
  while (v >= 0x80) {
    s.push_back((char)((v & 0x7F) | 0x80));
    v >>= 7;
  }
  s.push_back((char)v);
}
inline int varint_size(uint64_t v) {
  int n = 1;
  while (v >= 0x80) { v >>= 7; n++; }
  return n;
}
}  // namespace wire

class User_Record {
 public:
  User_Record() : vid_(0), rating_(0.f) {}
  int vid() const { return vid_; }
  float rating() const { return rating_; }
  void set_vid(int v) { vid_ = v; }
  void set_rating(float r) { rating_ = r; }
  bool parse(const uint8_t* p, const uint8_t* end) {
    while (p < end) {
      uint64_t tag;
      if (!wire::get_varint(p, end, tag)) return false;
      if (tag == 0x08) {
        uint64_t v;
        if (!wire::get_varint(p, end, v)) return false;
        vid_ = (int32_t)(uint32_t)v;
      } else if (tag == 0x15) {
        if (end - p < 4) return false;
        memcpy(&rating_, p, 4);
        p += 4;
      } else if (!wire::skip_field((uint32_t)(tag & 7), p, end)) {
        return false;
      }
    }
    return true;
  }
  int byte_size() const { return 1 + wire::varint_size((uint64_t)(int64_t)vid_) + 1 + 4; }
  void append_to(std::string& s) const {
    s.push_back((char)0x08);
    wire::put_varint(s, (uint64_t)(int64_t)vid_);
    s.push_back((char)0x15);
    char b[4];
    memcpy(b, &rating_, 4);
    s.append(b, 4);
  }

 private:
  int vid_;
  float rating_;
};

class User {
 public:
  typedef User_Record Record;
  User() : uid_(0) {}
  int uid() const { return uid_; }
  void set_uid(int u) { uid_ = u; }
  int record_size() const { return (int)rec_.size(); }
  const User_Record& record(int i) const { return rec_[i]; }
  User_Record* add_record() {
    rec_.emplace_back();
    return &rec_.back();
  }
  void Clear() {
    uid_ = 0;
    rec_.clear();
  }
  bool parse(const uint8_t* p, const uint8_t* end) {
    while (p < end) {
      uint64_t tag;
      if (!wire::get_varint(p, end, tag)) return false;
      if (tag == 0x08) {
        uint64_t v;
        if (!wire::get_varint(p, end, v)) return false;
        uid_ = (int32_t)(uint32_t)v;
      } else if (tag == 0x12) {
        uint64_t len;
        if (!wire::get_varint(p, end, len) || (uint64_t)(end - p) < len) return false;
        rec_.emplace_back();
        if (!rec_.back().parse(p, p + len)) return false;
        p += len;
      } else if (!wire::skip_field((uint32_t)(tag & 7), p, end)) {
        return false;
      }
    }
    return true;
  }
  int byte_size() const {
    int n = 1 + wire::varint_size((uint64_t)(int64_t)uid_);
    for (const auto& r : rec_) {
      int b = r.byte_size();
      n += 1 + wire::varint_size((uint64_t)b) + b;
    }
    return n;
  }
  void append_to(std::string& s) const {
    s.push_back((char)0x08);
    wire::put_varint(s, (uint64_t)(int64_t)uid_);
    for (const auto& r : rec_) {
      s.push_back((char)0x12);
      wire::put_varint(s, (uint64_t)r.byte_size());
      r.append_to(s);
    }
  }

 private:
  int uid_;
  std::vector<User_Record> rec_;
};

class Block {
 public:
  int user_size() const { return (int)used_; }
  const User& user(int i) const { return users_[i]; }
  User* add_user() {
    if (used_ == users_.size()) users_.emplace_back();
    users_[used_].Clear();
    return &users_[used_++];
  }
  void Clear() { used_ = 0; }
  // protobuf semantics: ParseFromArray clears the message first, then merges.
  bool ParseFromArray(const void* data, int size) {
    Clear();
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* end = p + size;
    while (p < end) {
      uint64_t tag;
      if (!wire::get_varint(p, end, tag)) return false;
      if (tag == 0x0A) {
        uint64_t len;
        if (!wire::get_varint(p, end, len) || (uint64_t)(end - p) < len) return false;
        if (!add_user()->parse(p, p + len)) return false;
        p += len;
      } else if (!wire::skip_field((uint32_t)(tag & 7), p, end)) {
        return false;
      }
    }
    return true;
  }
  bool SerializeToString(std::string* out) const {
    out->clear();
    for (size_t i = 0; i < used_; i++) {
      out->push_back((char)0x0A);
      wire::put_varint(*out, (uint64_t)users_[i].byte_size());
      users_[i].append_to(*out);
    }
    return true;
  }

 private:
  std::vector<User> users_;  // capacity is retained across Clear(), like protobuf's arenas
  size_t used_ = 0;
};

class Blocks {
 public:
  int block_size() const { return (int)blocks_.size(); }
  const Block& block(int i) const { return blocks_[i]; }
  Block* mutable_block(int i) { return &blocks_[i]; }
  Block* add_block() {
    blocks_.emplace_back();
    return &blocks_.back();
  }
  void Clear() { blocks_.clear(); }

 private:
  std::deque<Block> blocks_;  // add_block() must not move earlier blocks (model.cc:274 keeps pbk)
};

}  // namespace mf

#endif
