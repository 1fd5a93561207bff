// Synthetic rating files of a named shape (mfb_generate, include/mf_b200.h; recipe: SURVEY.md 8d).
// The reference ships no data; its own converter is data/getdata.cc, whose output layout
// (--method userwise --split S, then --method protobuf --size B) this generator mirrors.
//
// Everything random is a pure function of (seed, stream, index) through Philox4x32-10, so the
// data set is identical for any thread count and any user sharding.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <numeric>
#include <thread>
#include <vector>

#include "mfb_internal.h"

struct mfb_blocks;
namespace mfb {
mfb_blocks* blocks_new();
Dataset* blocks_mut(mfb_blocks* b);
}  // namespace mfb

namespace {

struct U4 {
  uint32_t x, y, z, w;
};

inline U4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

inline float u01_open(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }  // (0,1]
inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }                // [0,1)

inline void normal2(uint32_t a, uint32_t b, float* z0, float* z1) {
  const float r = sqrtf(-2.0f * logf(u01_open(a)));
  const float ang = 6.28318530717958647692f * u01(b);
  *z0 = r * cosf(ang);
  *z1 = r * sinf(ang);
}

enum Stream : uint32_t { S_DEGREE = 1, S_UFACTOR = 2, S_VFACTOR = 3, S_ITEMPERM = 4, S_DRAW = 5, S_USERPERM = 6 };

// Walker alias table over Zipf(s) ranks
struct Alias {
  std::vector<float> prob;
  std::vector<int32_t> alias;
  void build(int n, double s) {
    std::vector<double> w(n);
    double sum = 0;
    for (int k = 0; k < n; k++) sum += (w[k] = pow((double)(k + 1), -s));
    prob.assign(n, 0.f);
    alias.assign(n, 0);
    std::vector<double> scaled(n);
    std::vector<int32_t> small, large;
    for (int k = 0; k < n; k++) {
      scaled[k] = w[k] / sum * n;
      (scaled[k] < 1.0 ? small : large).push_back(k);
    }
    while (!small.empty() && !large.empty()) {
      const int32_t s_ = small.back(), l = large.back();
      small.pop_back();
      prob[s_] = (float)scaled[s_];
      alias[s_] = l;
      scaled[l] = scaled[l] + scaled[s_] - 1.0;
      if (scaled[l] < 1.0) {
        large.pop_back();
        small.push_back(l);
      }
    }
    for (int32_t k : large) prob[k] = 1.f, alias[k] = k;
    for (int32_t k : small) prob[k] = 1.f, alias[k] = k;
  }
  inline int32_t sample(uint32_t a, uint32_t b) const {
    const int32_t col = (int32_t)(((uint64_t)a * prob.size()) >> 32);
    return u01(b) < prob[col] ? col : alias[col];
  }
};

// ids 0..n-1 ordered by a Philox hash: a deterministic pseudo-random permutation
std::vector<int32_t> hashed_order(int32_t n, uint32_t stream, uint32_t salt, uint64_t seed) {
  std::vector<uint64_t> key(n);
  for (int32_t i = 0; i < n; i++) {
    const U4 h = philox((uint32_t)i, salt, 0, stream, seed);
    key[i] = ((uint64_t)h.x << 32) | (uint32_t)i;  // ties broken by id
  }
  std::sort(key.begin(), key.end());
  std::vector<int32_t> out(n);
  for (int32_t i = 0; i < n; i++) out[i] = (int32_t)(uint32_t)key[i];
  return out;
}

struct Staged {  // one user's ratings, dealt into destinations
  std::vector<int32_t> vid;
  std::vector<float> rating;
  std::vector<uint8_t> dest;  // 0..split-1 train chunk, split = test, split+1 = valid
};

}  // namespace

using namespace mfb;

extern "C" {

void mfb_gen_defaults(mfb_gen_params* p, int32_t nu, int32_t nv, int64_t nnz) {
  memset(p, 0, sizeof *p);
  p->nu = nu;
  p->nv = nv;
  p->nnz = nnz;
  p->rank = 16;
  p->gb = 2.76f;  // main.cc:100
  p->noise_sd = 0.5f;
  p->degree_sigma = 1.0f;
  p->zipf_s = 1.0f;
  p->test_frac = 0.01f;
  p->valid_frac = 0.0f;
  p->split = 4;              // run.py:4 "..._train_4by500"
  p->users_per_block = 500;
  p->seed = 0x4D46B200ull;
  p->user_begin = 0;
  p->user_end = nu;
  p->threads = 0;
}

int mfb_generate(const mfb_gen_params* pp, mfb_blocks** train_out, mfb_blocks** test_out,
                 mfb_blocks** valid_out) {
  MFB_REQUIRE(pp && train_out, "NULL argument");
  const mfb_gen_params p = *pp;
  MFB_REQUIRE(p.nu > 0 && p.nv > 1 && p.nnz > 0 && p.rank > 0 && p.rank <= 64 && p.split >= 1 &&
                  p.split <= 200 && p.users_per_block >= 1,
              "bad generator parameters");
  MFB_REQUIRE(0 <= p.user_begin && p.user_begin <= p.user_end && p.user_end <= p.nu, "bad user range");
  MFB_REQUIRE(p.test_frac >= 0 && p.valid_frac >= 0 && p.test_frac + p.valid_frac < 0.9f, "bad split fractions");
  const uint64_t seed = p.seed;
  const int nthreads = p.threads > 0 ? p.threads : std::max(1u, std::thread::hardware_concurrency());

  // 1. degrees: lognormal weights over ALL users (so a shard sees the same scale), scaled to nnz
  std::vector<float> w(p.nu);
  double wsum = 0;
  for (int32_t u = 0; u < p.nu; u++) {
    const U4 h = philox((uint32_t)u, 0, 0, S_DEGREE, seed);
    float z0, z1;
    normal2(h.x, h.y, &z0, &z1);
    w[u] = expf(p.degree_sigma * z0);
    wsum += w[u];
  }
  const int32_t max_deg = std::max(1, p.nv / 2);
  const double scale = (double)p.nnz / wsum;
  auto degree = [&](int32_t u) {
    const long d = lround((double)w[u] * scale);
    return (int32_t)std::min<long>(std::max<long>(d, 1), max_deg);
  };

  // 2. item popularity: Zipf over ranks, rank k -> item id item_of_rank[k]
  Alias zipf;
  zipf.build(p.nv, p.zipf_s);
  const std::vector<int32_t> item_of_rank = hashed_order(p.nv, S_ITEMPERM, 0, seed);

  // 3. planted item factors
  const int rank = p.rank;
  std::vector<float> vstar((size_t)p.nv * rank);
  for (int32_t i = 0; i < p.nv; i++)
    for (int j = 0; j < rank; j += 2) {
      const U4 h = philox((uint32_t)i, (uint32_t)j, 0, S_VFACTOR, seed);
      float z0, z1;
      normal2(h.x, h.y, &z0, &z1);
      vstar[(size_t)i * rank + j] = 0.5f * z0;
      if (j + 1 < rank) vstar[(size_t)i * rank + j + 1] = 0.5f * z1;
    }

  // 4. per-user draws (parallel over users of the shard)
  const int32_t ub = p.user_begin, ue = p.user_end, nlocal = ue - ub;
  std::vector<Staged> staged(nlocal);
  const int ndest = p.split + 2;
  const uint32_t test_thr = (uint32_t)(p.test_frac * 65536.0f);
  const uint32_t valid_thr = test_thr + (uint32_t)(p.valid_frac * 65536.0f);
  std::atomic<int32_t> next(0);
  auto worker = [&]() {
    std::vector<uint32_t> seen((size_t)(p.nv + 31) / 32);
    std::vector<float> ustar(rank);
    for (;;) {
      const int32_t lo = next.fetch_add(64);
      if (lo >= nlocal) break;
      const int32_t hi = std::min(nlocal, lo + 64);
      for (int32_t li = lo; li < hi; li++) {
        const int32_t u = ub + li;
        const int32_t deg = degree(u);
        for (int j = 0; j < rank; j += 2) {
          const U4 h = philox((uint32_t)u, (uint32_t)j, 0, S_UFACTOR, seed);
          float z0, z1;
          normal2(h.x, h.y, &z0, &z1);
          ustar[j] = 0.5f * z0;
          if (j + 1 < rank) ustar[j + 1] = 0.5f * z1;
        }
        Staged& st = staged[li];
        st.vid.reserve(deg);
        st.rating.reserve(deg);
        st.dest.reserve(deg);
        uint32_t attempt = 0;
        while ((int32_t)st.vid.size() < deg) {
          const U4 h = philox((uint32_t)u, attempt++, 0, S_DRAW, seed);
          const int32_t item = item_of_rank[zipf.sample(h.x, h.y)];
          uint32_t& word = seen[item >> 5];
          const uint32_t bit = 1u << (item & 31);
          if (word & bit) continue;  // no duplicate (u,i)
          word |= bit;
          const float* vs = &vstar[(size_t)item * rank];
          float dot = 0.f;
          for (int j = 0; j < rank; j++) dot += ustar[j] * vs[j];
          float z0, z1;
          normal2(h.z, h.w, &z0, &z1);
          float r = roundf(p.gb + dot + p.noise_sd * z0);
          r = std::min(5.f, std::max(1.f, r));
          // the low bytes of h.z / h.w are not used by the 24-bit uniforms above
          const uint32_t pick = ((h.z & 0xFF) << 8) | (h.w & 0xFF);
          uint8_t dest;
          if (pick < test_thr) dest = (uint8_t)p.split;
          else if (pick < valid_thr) dest = (uint8_t)(p.split + 1);
          else dest = (uint8_t)(((uint64_t)(pick - valid_thr) * p.split) / (65536u - valid_thr));
          st.vid.push_back(item);
          st.rating.push_back(r);
          st.dest.push_back(dest);
        }
        for (int32_t it : st.vid) seen[it >> 5] = 0;  // words touched by this user
      }
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
  }

  // 5. layout: per destination, users in a hashed order, B users per block
  mfb_blocks* outs[3] = {blocks_new(), blocks_new(), blocks_new()};
  std::vector<int32_t> cnt(nlocal);
  for (int dest = 0; dest < ndest; dest++) {
    Dataset* d = blocks_mut(outs[dest < p.split ? 0 : dest - p.split + 1]);
    const std::vector<int32_t> order = hashed_order(p.nu, S_USERPERM, (uint32_t)dest, seed);
    int32_t in_block = 0;
    for (int32_t u : order) {
      if (u < ub || u >= ue) continue;
      const Staged& st = staged[u - ub];
      int32_t c = 0;
      for (uint8_t x : st.dest) c += (x == dest);
      if (c == 0) continue;
      if ((int64_t)d->h_vid.size() + c >= (int64_t)INT32_MAX) {
        for (auto* o : outs) mfb_blocks_free(o);
        set_error("generated file exceeds int32 offsets; shard the users");
        return MFB_E_ARG;
      }
      d->h_run_uid.push_back(u);
      for (size_t k = 0; k < st.dest.size(); k++)
        if (st.dest[k] == dest) {
          d->h_vid.push_back(st.vid[k]);
          d->h_rating.push_back(st.rating[k]);
        }
      d->h_run_off.push_back((int32_t)d->h_vid.size());
      if (++in_block == p.users_per_block) {
        d->h_block_off.push_back((int64_t)d->h_run_uid.size());
        in_block = 0;
      }
    }
    if (in_block) d->h_block_off.push_back((int64_t)d->h_run_uid.size());  // chunk ends its block
  }
  *train_out = outs[0];
  if (test_out) *test_out = outs[1]; else mfb_blocks_free(outs[1]);
  if (valid_out) *valid_out = outs[2]; else mfb_blocks_free(outs[2]);
  return MFB_OK;
}

}  // extern "C"
