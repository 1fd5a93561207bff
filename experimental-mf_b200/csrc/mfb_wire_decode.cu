// Device-side decoder of the blocks.proto wire format: the reference parses every
// mf.Block again in every epoch on the host (ParseFromArray, mf.h:57-69; generated code blocks.pb.cc:253-300,
// 512-560); here the raw bytes of a chunk of the training file are copied to the GPU as they are and the
// records are decoded there.
//
// What the host keeps: the walk over the frames ([u32 size], util.h:81) and over the top-level fields of each
// Block - one jump per user - which yields the byte range of every serialized mf.User in the chunk.  What the
// device does, one thread per user-run:
//   wire_count_kernel   walks the fields of the User (uid = field 1 varint, record = field 2 length-delimited,
//                       anything else skipped by wire type, blocks.proto:8-12), counts the records, checks the
//                       framing and the uid range;
//   wire_scan_kernel    exclusive prefix sum of the counts -> run_off[nruns+1];
//   wire_decode_kernel  walks the User again and writes vid / rating of record j at run_off[run] + j (field 1
//                       varint and field 2 fixed32 of the Record, any order, unknown fields skipped,
//                       blocks.proto:3-6), checks the item range, counts the records per item;
//   wire_finish_kernel  records of the most rated item (the hot-row budget of the launch needs its share).
// Bytes are read through aligned 64-bit loads (the L1 keeps the line for the following records of the run); the
// canonical record protobuf's encoder writes - 12 LL 08 <vid varint> 15 <4 bytes> - is recognised from one
// 16-byte window without a loop.  Semantics equal proto_wire.cc's host decoder (last value of a repeated scalar
// wins, missing fields read as zero, truncation is an error).
#include "mfb_internal.h"
#include "mfb_wire_decode.h"

namespace mfb {
namespace {

// 8 bytes at byte offset o of an 8-byte aligned buffer (which is padded: reading up to 16 bytes past the end of
// the data is harmless)
__device__ __forceinline__ uint64_t ld8(const uint64_t* __restrict__ raw, int64_t o) {
  const uint64_t* a = raw + (o >> 3);
  const unsigned sh = ((unsigned)o & 7u) * 8u;
  const uint64_t lo = __ldg(a);
  if (sh == 0) return lo;
  return (lo >> sh) | (__ldg(a + 1) << (64u - sh));
}
// 16 bytes at byte offset o as two words
__device__ __forceinline__ void ld16(const uint64_t* __restrict__ raw, int64_t o, uint64_t* w0, uint64_t* w1) {
  const uint64_t* a = raw + (o >> 3);
  const unsigned sh = ((unsigned)o & 7u) * 8u;
  const uint64_t x0 = __ldg(a), x1 = __ldg(a + 1);
  if (sh == 0) {
    *w0 = x0;
    *w1 = x1;
    return;
  }
  const uint64_t x2 = __ldg(a + 2);
  *w0 = (x0 >> sh) | (x1 << (64u - sh));
  *w1 = (x1 >> sh) | (x2 << (64u - sh));
}

// varint at *o (at most 10 bytes, must end before `end`); false = malformed / truncated
__device__ __forceinline__ bool varint(const uint64_t* __restrict__ raw, int64_t* o, int64_t end, uint64_t* out) {
  uint64_t v = 0;
  int64_t p = *o;
  for (int i = 0; i < 10; i++) {
    if (p >= end) return false;
    const uint64_t b = ld8(raw, p) & 0xffu;
    p++;
    v |= (b & 0x7fu) << (7 * i);
    if (!(b & 0x80u)) {
      *o = p;
      *out = v;
      return true;
    }
  }
  return false;
}
// skip a field of the given wire type (proto_wire.cc Cursor::skip)
__device__ __forceinline__ bool skip_field(const uint64_t* __restrict__ raw, int64_t* o, int64_t end, unsigned wt) {
  uint64_t v;
  switch (wt) {
    case 0: return varint(raw, o, end, &v);
    case 1: *o += 8; return *o <= end;
    case 2:
      if (!varint(raw, o, end, &v)) return false;
      if (v > (uint64_t)(end - *o)) return false;
      *o += (int64_t)v;
      return true;
    case 5: *o += 4; return *o <= end;
    default: return false;  // groups / reserved wire types
  }
}

__device__ __forceinline__ void report(WireResult* res, int code, int run) {
  if (atomicCAS(&res->err, 0, code) == 0) res->err_run = run;
}

__global__ void __launch_bounds__(256) wire_count_kernel(const uint64_t* __restrict__ raw, const int32_t* __restrict__ span,
                                                         int nruns, int nu, int32_t* __restrict__ run_uid,
                                                         int32_t* __restrict__ count, WireResult* res) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nruns) return;
  int64_t o = span[2 * r];
  const int64_t end = span[2 * r + 1];
  int n = 0;
  uint32_t uid = 0;
  bool ok = true;
  while (o < end) {
    const uint64_t w = ld8(raw, o);
    if ((w & 0x80ffu) == 0x0012u) {  // record with a one-byte length: the common case
      o += 2 + (int64_t)((w >> 8) & 0x7fu);
      n++;
      continue;
    }
    uint64_t tag, v;
    if (!varint(raw, &o, end, &tag)) { ok = false; break; }
    if (tag == 0x08) {
      if (!varint(raw, &o, end, &v)) { ok = false; break; }
      uid = (uint32_t)v;
    } else if (tag == 0x12) {
      if (!varint(raw, &o, end, &v) || v > (uint64_t)(end - o)) { ok = false; break; }
      o += (int64_t)v;
      n++;
    } else if (!skip_field(raw, &o, end, (unsigned)(tag & 7))) {
      ok = false;
      break;
    }
  }
  if (!ok || o != end) {
    report(res, WIRE_E_FORMAT, r);
    n = 0;
  } else if (uid >= (uint32_t)nu) {
    report(res, WIRE_E_UID, r);
  }
  run_uid[r] = (int32_t)uid;
  count[r] = n;
}

// run_off[0] = 0, run_off[r+1] = count[0] + ... + count[r]; one block (a chunk holds ~1e5 runs): every warp owns a
// contiguous slice and walks it 32 counts at a time (coalesced), first to sum it, then - once the sums of the slices
// before it are known - to write the running totals
__global__ void __launch_bounds__(1024) wire_scan_kernel(const int32_t* __restrict__ count, int nruns,
                                                         int32_t* __restrict__ run_off, WireResult* res) {
  __shared__ long long part[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per = ((nruns + 31) / 32 + 31) & ~31;  // counts per warp, a multiple of 32
  const int lo = min(w * per, nruns), hi = min(lo + per, nruns);
  long long s = 0;
  for (int i = lo + lane; i < hi; i += 32) s += count[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) part[w] = s;
  __syncthreads();
  if (w == 0) {  // exclusive scan of the 32 slice sums
    long long v = part[lane], x = v;
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    part[lane] = x - v;
    if (lane == 31) {
      res->nratings = x;
      if (x > 0x7fffffffLL) report(res, WIRE_E_FORMAT, -1);
    }
  }
  __syncthreads();
  long long run = part[w];
  if (threadIdx.x == 0) run_off[0] = 0;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const long long c = i < hi ? count[i] : 0;
    long long x = c;
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (i < hi) run_off[i + 1] = (int32_t)(run + x);
    run += __shfl_sync(0xffffffffu, x, 31);
  }
}

__global__ void __launch_bounds__(256) wire_decode_kernel(const uint64_t* __restrict__ raw, const int32_t* __restrict__ span,
                                                          const int32_t* __restrict__ run_off, int nruns, int nv,
                                                          int64_t cap_ratings, int32_t* __restrict__ vid,
                                                          float* __restrict__ rating, int32_t* __restrict__ hist,
                                                          WireResult* res) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nruns) return;
  if (res->err) return;  // a malformed chunk is not decoded (its offsets mean nothing)
  int64_t o = span[2 * r];
  const int64_t end = span[2 * r + 1];
  int64_t j = run_off[r];
  const int64_t jend = run_off[r + 1];
  if (jend > cap_ratings) {
    report(res, WIRE_E_CAPACITY, r);
    return;
  }
  bool ok = true, in_range = true;
  while (o < end) {
    uint64_t w0, w1;
    ld16(raw, o, &w0, &w1);
    uint32_t v = 0;
    uint32_t bits = 0;
    if ((w0 & 0xff80ffu) == 0x080012u) {
      // 12 LL 08 vv [vv [vv]] 15 r0 r1 r2 r3 with LL = 6 + length of the vid varint
      const unsigned len = (unsigned)(w0 >> 8) & 0x7fu;
      v = (unsigned)(w0 >> 24) & 0x7fu;
      unsigned vl = 1;
      if (w0 & (0x80ull << 24)) {
        v |= ((unsigned)(w0 >> 32) & 0x7fu) << 7;
        vl = 2;
        if (w0 & (0x80ull << 32)) {
          v |= ((unsigned)(w0 >> 40) & 0x7fu) << 14;
          vl = (w0 & (0x80ull << 40)) ? 0 : 3;  // longer varints go the general way
        }
      }
      const unsigned q = 3 + vl;  // offset of the rating tag: 4, 5 or 6
      if (vl && len == vl + 6 && ((w0 >> (8 * q)) & 0xffu) == 0x15u) {
        const unsigned sh = 8 * (q + 1);  // 40, 48 or 56
        bits = (uint32_t)((w0 >> sh) | (w1 << (64u - sh)));
        o += 2 + len;
        goto emit;
      }
    }
    {  // general path (proto_wire.cc decode_block)
      uint64_t tag, x;
      if (!varint(raw, &o, end, &tag)) { ok = false; break; }
      if (tag == 0x12) {
        if (!varint(raw, &o, end, &x) || x > (uint64_t)(end - o)) { ok = false; break; }
        const int64_t rend = o + (int64_t)x;
        while (o < rend) {
          uint64_t rt;
          if (!varint(raw, &o, rend, &rt)) { ok = false; break; }
          if (rt == 0x08) {
            if (!varint(raw, &o, rend, &x)) { ok = false; break; }
            v = (uint32_t)x;
          } else if (rt == 0x15) {
            if (rend - o < 4) { ok = false; break; }
            bits = (uint32_t)ld8(raw, o);
            o += 4;
          } else if (!skip_field(raw, &o, rend, (unsigned)(rt & 7))) {
            ok = false;
            break;
          }
        }
        if (!ok || o != rend) { ok = false; break; }
      } else {
        if (tag == 0x08) {
          if (!varint(raw, &o, end, &x)) { ok = false; break; }
        } else if (!skip_field(raw, &o, end, (unsigned)(tag & 7))) {
          ok = false;
          break;
        }
        continue;  // not a record
      }
    }
  emit:
    if (j >= jend) { ok = false; break; }  // more records than the count pass saw: cannot happen on the same bytes
    if (v >= (uint32_t)nv) {
      in_range = false;
      v = 0;
    } else if (hist) {
      atomicAdd(hist + v, 1);
    }
    vid[j] = (int32_t)v;
    rating[j] = __uint_as_float(bits);
    j++;
  }
  if (!ok || j != jend) report(res, WIRE_E_FORMAT, r);
  else if (!in_range) report(res, WIRE_E_VID, r);
}

__global__ void __launch_bounds__(1024) wire_finish_kernel(const int32_t* __restrict__ hist, int nv, WireResult* res) {
  __shared__ int best[32];
  int m = 0;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) m = max(m, hist[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = best[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) res->top_count = m;
  }
}

}  // namespace

int launch_wire_decode(Context* c, cudaStream_t stream, const void* d_raw, const int32_t* d_span, int nruns,
                       int64_t cap_ratings, int32_t* d_run_uid, int32_t* d_run_off, int32_t* d_count, int32_t* d_vid,
                       float* d_rating, int32_t* d_hist, WireResult* d_res) {
  MFB_CUDA(cudaMemsetAsync(d_res, 0, sizeof(WireResult), stream));
  if (d_hist) MFB_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)c->nv * sizeof(int32_t), stream));
  const uint64_t* raw = reinterpret_cast<const uint64_t*>(d_raw);
  const int grid = (nruns + 255) / 256;
  if (nruns > 0) wire_count_kernel<<<grid, 256, 0, stream>>>(raw, d_span, nruns, c->nu, d_run_uid, d_count, d_res);
  wire_scan_kernel<<<1, 1024, 0, stream>>>(d_count, nruns, d_run_off, d_res);
  if (nruns > 0)
    wire_decode_kernel<<<grid, 256, 0, stream>>>(raw, d_span, d_run_off, nruns, c->nv, cap_ratings, d_vid, d_rating,
                                                 d_hist, d_res);
  if (d_hist) wire_finish_kernel<<<1, 1024, 0, stream>>>(d_hist, c->nv, d_res);
  MFB_CUDA(cudaGetLastError());
  c->launches += (nruns > 0 ? 2 : 0) + 1 + (d_hist ? 1 : 0);
  return MFB_OK;
}

}  // namespace mfb
