// getdata - the reference's rating-file converter (data/getdata.cc) on top of libmf_b200's host-side
// block writer (no GPU needed), plus a generator of the synthetic shapes used by bench.py.
//
//   getdata -r raw.txt   -w user.txt  --method userwise --split S      (getdata.cc:21-80,157-164)
//   getdata -r user.txt  -w train.bin --method protobuf --size B       (getdata.cc:82-126)
//   getdata              -w prefix    --method synth --nu U --nv V --nnz N [--split S --size B
//                                      --test F --valid F --seed X --head M]  (new: prefix.train/.test/.valid;
//                                      --head M keeps the first Blocks of the training file, M ratings or more)
//
// raw.txt: first line the record count, then "user,item,rating,timestamp" lines (getdata.cc:21-32).
// user.txt: "uid:" lines each followed by "vid,rating" lines (getdata.cc:39-50).
// Output of --method protobuf: [u32 size][mf.Block] frames with B users per Block, byte-compatible
// with the reference's writer (hand-written wire encoder, csrc/proto_wire.cc).
//
// Differences from the reference, both in the userwise method only: the four std::random_shuffle
// passes (libstdc++ rand()-driven, getdata.cc:33-36) are one Fisher-Yates pass driven by a fixed
// xorshift stream (--seed), and users inside a chunk are written in order of first appearance
// instead of unordered_map iteration order - both orders are arbitrary in the reference.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mf_b200.h"

namespace {

struct Rec {
  int32_t u, v;
  float r;
};

void hint() {  // getdata.cc:128-134
  printf("-r         [input_file_name]\n");
  printf("-w         [output_file_name]\n");
  printf("--method   [userwise/protobuf/synth]\n");
  printf("--split    [number_of_splits_for_rating_matrix]\thints: 1~10 splits are recommended\n");
  printf("--size     [number_of_users_in_each_block]\thints: 1 fread reads 1 block each time\n");
}

int die(const char* what) {
  fprintf(stderr, "getdata: %s: %s\n", what, mfb_last_error());
  return 3;
}

int userwise(const char* in, const char* out, int split, uint64_t seed) {
  FILE* fp = fopen(in, "r");
  if (!fp) {
    fprintf(stderr, "getdata: cannot open %s\n", in);
    return 2;
  }
  int nn = 0;
  if (fscanf(fp, "%d", &nn) != 1 || nn < 0) {
    fprintf(stderr, "getdata: %s: missing record count\n", in);
    fclose(fp);
    return 2;
  }
  std::vector<Rec> data;
  data.reserve(nn);
  for (int i = 0; i < nn; i++) {
    Rec x;
    int t;
    if (fscanf(fp, "%d,%d,%f,%d", &x.u, &x.v, &x.r, &t) != 4) break;  // getdata.cc:30
    data.push_back(x);
  }
  fclose(fp);
  uint64_t s = seed ? seed : 0x9E3779B97F4A7C15ull;
  for (size_t i = data.size(); i > 1; i--) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    std::swap(data[i - 1], data[s % i]);
  }
  FILE* fo = fopen(out, "w");
  if (!fo) {
    fprintf(stderr, "getdata: cannot create %s\n", out);
    return 2;
  }
  const size_t n = data.size(), nb = split > 0 ? n / split : n;
  for (int c = 0; c < split; c++) {
    const size_t lo = c * nb, hi = (c == split - 1) ? n : lo + nb;  // the last chunk takes the remainder
    std::unordered_map<int32_t, std::vector<std::pair<int32_t, float>>> du;
    std::vector<int32_t> order;
    for (size_t j = lo; j < hi; j++) {
      auto it = du.find(data[j].u);
      if (it == du.end()) {
        order.push_back(data[j].u);
        it = du.emplace(data[j].u, std::vector<std::pair<int32_t, float>>()).first;
      }
      it->second.push_back({data[j].v, data[j].r});
    }
    for (int32_t u : order) {
      fprintf(fo, "%d:\n", u);
      for (auto& e : du[u]) fprintf(fo, "%d,%f\n", e.first, e.second);  // getdata.cc:47
    }
  }
  fclose(fo);
  return 0;
}

int protobuf(const char* in, const char* out, int block_size) {
  FILE* fp = fopen(in, "r");
  if (!fp) {
    fprintf(stderr, "getdata: cannot open %s\n", in);
    return 2;
  }
  std::vector<int64_t> block_off{0};
  std::vector<int32_t> run_uid, run_off{0}, vid;
  std::vector<float> rating;
  char line[256];
  while (fgets(line, sizeof line, fp)) {
    size_t len = strlen(line);
    while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
    if (!len) continue;
    if (line[len - 1] == ':') {  // a new user (getdata.cc:97); a Block closes every block_size users
      if (!run_uid.empty()) run_off.push_back((int32_t)vid.size());
      if (!run_uid.empty() && run_uid.size() % (size_t)block_size == 0) block_off.push_back((int64_t)run_uid.size());
      run_uid.push_back(atoi(line));
      continue;
    }
    int v;
    float r;
    if (sscanf(line, "%d,%f", &v, &r) == 2 && !run_uid.empty()) {
      vid.push_back(v);
      rating.push_back(r);
    }
  }
  fclose(fp);
  if (!run_uid.empty()) run_off.push_back((int32_t)vid.size());
  block_off.push_back((int64_t)run_uid.size());
  mfb_blocks* b = nullptr;
  if (mfb_blocks_from_arrays((int64_t)block_off.size() - 1, block_off.data(), (int64_t)run_uid.size(), run_uid.data(),
                             run_off.data(), vid.data(), rating.data(), &b))
    return die("blocks");
  const int rc = mfb_blocks_write(b, out);
  mfb_blocks_free(b);
  return rc ? die("write") : 0;
}

}  // namespace

int main(int argc, char** argv) {
  const char *read = nullptr, *write = nullptr, *method = nullptr;
  int bk = 1, block_size = 1000;  // getdata.cc:19,138
  long long nu = 0, nv = 0, nnz = 0;
  double test = 0.01, valid = 0.0;
  uint64_t seed = 0;
  long long head = 0;
  for (int i = 1; i < argc; i++) {
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (!strcmp(argv[i], "-r")) read = next();
    else if (!strcmp(argv[i], "-w")) write = next();
    else if (!strcmp(argv[i], "--method")) method = next();
    else if (!strcmp(argv[i], "--split")) bk = atoi(next());
    else if (!strcmp(argv[i], "--size")) block_size = atoi(next());
    else if (!strcmp(argv[i], "--nu")) nu = atoll(next());
    else if (!strcmp(argv[i], "--nv")) nv = atoll(next());
    else if (!strcmp(argv[i], "--nnz")) nnz = atoll(next());
    else if (!strcmp(argv[i], "--test")) test = atof(next());
    else if (!strcmp(argv[i], "--valid")) valid = atof(next());
    else if (!strcmp(argv[i], "--seed")) seed = strtoull(next(), nullptr, 0);
    else if (!strcmp(argv[i], "--head")) head = atoll(next());
    else {
      printf("unknown parameters.\n\n");
      hint();
      return 1;
    }
  }
  if (!write || !method || (!read && strcmp(method, "synth"))) {
    printf("Please at least indicate the input, output and method.\n\n");
    hint();
    return 1;
  }
  if (bk < 1 || block_size < 1) {
    fprintf(stderr, "getdata: --split and --size must be positive\n");
    return 1;
  }
  if (!strcmp(method, "userwise")) return userwise(read, write, bk, seed);
  if (!strcmp(method, "protobuf")) return protobuf(read, write, block_size);
  if (!strcmp(method, "synth")) {
    if (nu <= 0 || nv <= 1 || nnz <= 0) {
      fprintf(stderr, "getdata: --method synth needs --nu --nv --nnz\n");
      return 1;
    }
    mfb_gen_params p;
    mfb_gen_defaults(&p, (int32_t)nu, (int32_t)nv, nnz);
    p.split = bk > 1 ? bk : p.split;
    p.users_per_block = block_size != 1000 ? block_size : p.users_per_block;
    p.test_frac = (float)test;
    p.valid_frac = (float)valid;
    if (seed) p.seed = seed;
    mfb_blocks *tr = nullptr, *te = nullptr, *va = nullptr;
    if (head > 0 && head < nnz) {  // the sample needs only the users of its Blocks (the generator shards by user;
      // a user's ratings are dealt over `split` chunks, so a prefix of M ratings needs ~split*M/nnz of the users)
      const double frac = std::min(1.0, 1.3 * p.split * (double)head / (double)nnz + 0.01);
      p.user_end = std::max<int32_t>(1, (int32_t)(nu * frac));
    }
    if (mfb_generate(&p, &tr, &te, &va)) return die("generate");
    if (head > 0 && head < mfb_blocks_num_ratings(tr)) {
      const int64_t* bo = mfb_blocks_block_off(tr);
      const int32_t* ro = mfb_blocks_run_off(tr);
      int64_t nb = 0;
      while (nb < mfb_blocks_num_blocks(tr) && ro[bo[nb]] < head) nb++;
      mfb_blocks* cut = nullptr;
      if (mfb_blocks_from_arrays(nb, bo, bo[nb], mfb_blocks_run_uid(tr), ro, mfb_blocks_vid(tr), mfb_blocks_rating(tr), &cut))
        return die("head");
      mfb_blocks_free(tr);
      tr = cut;
    }
    const std::string base(write);
    int rc = mfb_blocks_write(tr, (base + ".train").c_str());
    if (!rc) rc = mfb_blocks_write(te, (base + ".test").c_str());
    if (!rc && valid > 0) rc = mfb_blocks_write(va, (base + ".valid").c_str());
    printf("train %lld ratings in %lld blocks, test %lld, valid %lld\n", (long long)mfb_blocks_num_ratings(tr),
           (long long)mfb_blocks_num_blocks(tr), (long long)mfb_blocks_num_ratings(te), (long long)mfb_blocks_num_ratings(va));
    mfb_blocks_free(tr);
    mfb_blocks_free(te);
    mfb_blocks_free(va);
    return rc ? die("write") : 0;
  }
  return 1;  // getdata.cc:169
}
