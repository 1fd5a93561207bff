// blocks.proto wire codec (see proto_wire.cc).
#ifndef MFB_PROTO_WIRE_H
#define MFB_PROTO_WIRE_H

#include <stddef.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

#include "mfb_internal.h"

namespace mfb {

// One decoded mf.Block as flat arrays: users in order, rec_off[nusers+1] into vid/rating.
struct BlockSink {
  std::vector<int32_t> uid;
  std::vector<int32_t> rec_off{0};
  std::vector<int32_t> vid;
  std::vector<float> rating;
};

// Appends the users of one serialized mf.Block to `out` (out->rec_off must start as {0}).
bool decode_block(const void* data, size_t size, BlockSink* out);
// Serializes one mf.Block, byte-identical to protobuf's encoder for the same content.
void encode_block(int32_t nusers, const int32_t* uid, const int32_t* rec_off, const int32_t* vid,
                  const float* rating, std::string* out);
// Walks the [u32 size][bytes] frames of a file; fn returns false to abort (malformed block).
int for_each_frame(const char* path, const std::function<bool(const void*, size_t)>& fn);

}  // namespace mfb
#endif
