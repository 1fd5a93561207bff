// Host-side mirror of the reference's model.cc / main.cc run() drivers on top of the C ABI.
// Every method cites the reference code whose behaviour it keeps (file:line under the
// reference's src/).  All numerics of the hot path happen in libmf_b200.so on the GPU.
#include "model.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

typedef std::chrono::high_resolution_clock Time;
static std::chrono::time_point<Time> s, e;  // model.cc:4 / util.h:48

static void check(int rc, const char* what) {
  if (rc != MFB_OK) {
    fprintf(stderr, "mf_b200: %s failed (%d): %s\n", what, rc, mfb_last_error());
    exit(3);
  }
}

static uint64_t default_seed() {
  const char* env = getenv("MF_SEED");
  if (env) return strtoull(env, nullptr, 0);
  // the reference seeds its engine from the system clock (model.cc:3)
  return (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();
}

// ---------------------------------------------------------------------------------- mf::Blocks
namespace mf {
Blocks::~Blocks() {
  if (owner_ && ds_ >= 0) mfb_dataset_free(owner_, ds_);
  if (blocks_) mfb_blocks_free(blocks_);
}
int64_t Blocks::ratings() const { return blocks_ ? mfb_blocks_num_ratings(blocks_) : 0; }
}  // namespace mf

void plain_read(const char* data, mf::Blocks& blocks) {  // util.h:76-88
  check(mfb_blocks_read(data, &blocks.blocks_), "plain_read");
}

// ------------------------------------------------------------------------------------------ MF
MF::MF(char* train_data, char* test_data, char* result, char* model, int dim, int iter, float eta,
       float gam, float lambda, float gb, int nu, int nv, int fly, int stride)
    : theta_(nullptr), phi_(nullptr), bu_(nullptr), bv_(nullptr), train_data_(train_data),
      test_data_(test_data), result_(result), model_(model), gb_(gb), dim_(dim), iter_(iter),
      eta_(eta), gam_(gam), lambda_(lambda), eta0_(eta), nu_(nu), nv_(nv), data_in_fly_(fly),
      prefetch_stride_(stride), ctx_(nullptr), start_round_(0), train_ds_(-1), device_(0) {
  const char* dev = getenv("MF_DEVICE");
  if (dev) device_ = atoi(dev);
}

MF::~MF() {
  if (ctx_) mfb_destroy(ctx_);
}

// host mirrors with the reference's layout: bu_|bv_ contiguous (model.cc:12-13), one table of
// nu+nv row pointers (model.cc:17-18), row stride padding(dim) (model.cc:15, util.h:163-165)
void MF::alloc_host(int extra_floats) {
  const int pad = mfb_padding(dim_);
  bias_store_.assign((size_t)nu_ + nv_ + extra_floats, 0.f);
  bu_ = bias_store_.data();
  bv_ = bu_ + nu_;
  theta_store_.assign((size_t)nu_ * pad, 0.f);
  phi_store_.assign((size_t)nv_ * pad, 0.f);
  row_ptrs_.resize((size_t)nu_ + nv_);
  for (int i = 0; i < nu_; i++) row_ptrs_[i] = theta_store_.data() + (size_t)i * pad;
  for (int i = 0; i < nv_; i++) row_ptrs_[nu_ + i] = phi_store_.data() + (size_t)i * pad;
  theta_ = row_ptrs_.data();
  phi_ = theta_ + nu_;
}

// `--fly N > 1` = the production schedule at the library's own width (model.h); an explicit hot-row budget
// comes from the environment
void MF::apply_options() {
  const char* rc = getenv("MF_ROW_CONCURRENCY");
  if (rc) check(mfb_set_option(ctx_, "row_concurrency", atoi(rc)), "row_concurrency");
  if (start_round_ > 0) check(mfb_set_option(ctx_, "model_age", start_round_), "model_age");
}

void MF::init() {  // model.cc:10-34
  check(mfb_create(&ctx_, device_, nu_, nv_, dim_), "mfb_create");
  apply_options();
  alloc_host(0);
  check(mfb_init_normal(ctx_, default_seed(), 1e-2f), "mfb_init_normal");  // N(0,1)*1e-2
  pull();
}

void MF::pull() {
  const int pad = mfb_padding(dim_);
  check(mfb_download(ctx_, MFB_THETA, theta_store_.data(), 0, nu_, pad), "download theta");
  check(mfb_download(ctx_, MFB_PHI, phi_store_.data(), 0, nv_, pad), "download phi");
  check(mfb_download(ctx_, MFB_BU, bu_, 0, nu_, 1), "download bu");
  check(mfb_download(ctx_, MFB_BV, bv_, 0, nv_, 1), "download bv");
}

void MF::push() {
  const int pad = mfb_padding(dim_);
  check(mfb_upload(ctx_, MFB_THETA, theta_store_.data(), 0, nu_, pad), "upload theta");
  check(mfb_upload(ctx_, MFB_PHI, phi_store_.data(), 0, nv_, pad), "upload phi");
  check(mfb_upload(ctx_, MFB_BU, bu_, 0, nu_, 1), "upload bu");
  check(mfb_upload(ctx_, MFB_BV, bv_, 0, nv_, 1), "upload bv");
}

void MF::seteta(int round) { eta_ = mfb_seteta(eta0_, round, gam_); }  // model.cc:36-38

int MF::schedule() const { return data_in_fly_ <= 1 ? MFB_MODE_ORDERED : MFB_MODE_ATOMIC; }

// MF_STREAM_INGEST=0: parse the training file on the host cores and upload it (mfb_dataset_load_file + finalize);
// default: one streaming pass with the records decoded on the GPU (mfb_dataset_ingest_file), which run(MF&) fuses
// with the first epoch - the reference overlaps read, parse and update the same way (main.cc:45-50)
static bool stream_ingest() {
  const char* e = getenv("MF_STREAM_INGEST");
  return !e || atoi(e) != 0;
}

void MF::load_train() {
  if (train_ds_ >= 0) return;
  check(mfb_dataset_create(ctx_, &train_ds_), "mfb_dataset_create");
  if (stream_ingest() && !needs_host_records()) {
    check(mfb_dataset_ingest_file(ctx_, train_ds_, train_data_, 0, 0.f, 0.f, gb_, schedule(), 0, nullptr), "ingest training file");
    return;
  }
  check(mfb_dataset_load_file(ctx_, train_ds_, train_data_), "load training file");
  check(mfb_dataset_finalize(ctx_, train_ds_), "mfb_dataset_finalize");
}

int MF::dataset_of(const mf::Blocks& blocks) {
  if (blocks.ds_ < 0 || blocks.owner_ != ctx_) {
    int ds;
    check(mfb_dataset_create(ctx_, &ds), "mfb_dataset_create");
    check(mfb_dataset_append_blocks(ctx_, ds, blocks.blocks_), "mfb_dataset_append_blocks");
    check(mfb_dataset_finalize(ctx_, ds), "mfb_dataset_finalize");
    blocks.ds_ = ds;
    blocks.owner_ = ctx_;
  }
  return blocks.ds_;
}

float MF::calc_mse(const mf::Blocks& blocks, int& ndata) {  // model.cc:41-73: returns the SUM
  double sse = 0.0;
  int64_t n = 0;
  check(mfb_sse(ctx_, dataset_of(blocks), gb_, &sse, &n), "mfb_sse");
  ndata = (int)n;
  return (float)sse;
}

// MF_TILE_RATINGS=n: out of core, as the reference runs (it re-reads and re-parses the training file every
// epoch and never holds it, mf.h:24-69): the epoch goes straight from the file through two device tile buffers
// of n ratings each (mfb_sgd_epoch_from_file); unset: the file is parsed once and stays resident in HBM
static int64_t tile_ratings() {
  const char* t = getenv("MF_TILE_RATINGS");
  return t ? atoll(t) : 0;
}

void MF::sgd_epoch() {
  if (tile_ratings() > 0) {
    check(mfb_sgd_epoch_from_file(ctx_, train_data_, eta_, lambda_, gb_, schedule(), tile_ratings(), nullptr),
          "mfb_sgd_epoch_from_file");
    return;
  }
  if (train_ds_ < 0 && stream_ingest() && !needs_host_records()) {
    // the first epoch of the run: ingest and update in one pass over the file
    check(mfb_dataset_create(ctx_, &train_ds_), "mfb_dataset_create");
    check(mfb_dataset_ingest_file(ctx_, train_ds_, train_data_, 1, eta_, lambda_, gb_, schedule(), 0, nullptr),
          "ingest training file + first epoch");
    return;
  }
  load_train();
  check(mfb_sgd_epoch(ctx_, train_ds_, eta_, lambda_, gb_, schedule()), "mfb_sgd_epoch");
}

// checkpoint layout of model.cc:75-122: nv nu dim (int32) | lambda (f32) | bv | phi | bu | theta
static void die_io(const char* what, const char* path) {
  fprintf(stderr, "mf_b200: cannot %s %s\n", what, path);
  exit(3);
}
static void must_read(void* p, size_t n, FILE* f, const char* path) {
  if (fread(p, 1, n, f) != n) die_io("read", path);
}

static void read_factors(FILE* fp, const char* path, MF* m) {
  must_read(m->bv_, sizeof(float) * m->nv_, fp, path);
  for (int i = 0; i < m->nv_; i++) must_read(m->phi_[i], sizeof(float) * m->dim_, fp, path);
  must_read(m->bu_, sizeof(float) * m->nu_, fp, path);
  for (int i = 0; i < m->nu_; i++) must_read(m->theta_[i], sizeof(float) * m->dim_, fp, path);
}
static void write_factors(FILE* fp, const MF* m) {
  fwrite(m->bv_, sizeof(float), m->nv_, fp);
  for (int i = 0; i < m->nv_; i++) fwrite(m->phi_[i], sizeof(float), m->dim_, fp);
  fwrite(m->bu_, sizeof(float), m->nu_, fp);
  for (int i = 0; i < m->nu_; i++) fwrite(m->theta_[i], sizeof(float), m->dim_, fp);
}
static void read_header(FILE* fp, const char* path, MF* m) {
  int nv, nu, dim;
  must_read(&nv, 4, fp, path);
  must_read(&nu, 4, fp, path);
  must_read(&dim, 4, fp, path);
  if (nv != m->nv_ || nu != m->nu_ || dim != m->dim_) {  // the reference silently adopts them
    fprintf(stderr, "mf_b200: %s holds nv=%d nu=%d dim=%d, expected %d %d %d\n", path, nv, nu, dim,
            m->nv_, m->nu_, m->dim_);
    exit(3);
  }
}

// the reference's checkpoint carries neither the round nor the step size, so a reloaded model
// starts over at eta0.  A small text file next to the checkpoint keeps them; absent (a checkpoint written by the
// reference), the run starts at round 1 as the reference would.
void MF::write_state(const char* file, int round) const {
  char path[600];
  snprintf(path, sizeof path, "%s.state", file);
  FILE* fp = fopen(path, "w");
  if (!fp) die_io("create", path);
  fprintf(fp, "mf_b200_state 1\nround %d\neta0 %.9g\ngam %.9g\neta %.9g\n", round, eta0_, gam_, eta_);
  fclose(fp);
}
bool MF::read_state(const char* file) {
  char path[600];
  snprintf(path, sizeof path, "%s.state", file);
  FILE* fp = fopen(path, "r");
  if (!fp) return false;
  int version = 0, round = 0;
  float eta0 = 0.f, gam = 0.f, eta = 0.f;
  const int got = fscanf(fp, "mf_b200_state %d round %d eta0 %g gam %g eta %g", &version, &round, &eta0, &gam, &eta);
  fclose(fp);
  if (got != 5 || version != 1 || round < 0) {
    fprintf(stderr, "mf_b200: %s is not a state file of this program\n", path);
    exit(3);
  }
  start_round_ = round;  // eta0 and gam stay the command line's: the next round uses eta0 / (round+1)^gam
  if (ctx_) check(mfb_set_option(ctx_, "model_age", start_round_), "model_age");
  return true;
}

void MF::read_model() {  // model.cc:75-97
  FILE* fp = fopen(model_, "rb");
  if (!fp) die_io("open", model_);
  read_header(fp, model_, this);
  must_read(&lambda_, 4, fp, model_);
  read_factors(fp, model_, this);
  fclose(fp);
  push();
  read_state(model_);  // (after push: an upload marks the model as new)
}

void MF::save_model(int round) {  // model.cc:98-122
  pull();
  char file[512];
  snprintf(file, sizeof file, "%s_%d", result_, round);
  FILE* fp = fopen(file, "wb");
  if (!fp) die_io("create", file);
  fwrite(&nv_, 4, 1, fp);
  fwrite(&nu_, 4, 1, fp);
  fwrite(&dim_, 4, 1, fp);
  fwrite(&lambda_, 4, 1, fp);
  write_factors(fp, this);
  fclose(fp);
  write_state(file, round);
}

// ---------------------------------------------------------------------------------------- DPMF
DPMF::DPMF(char* train_data, char* test_data, char* result, char* model, int dim, int iter,
           float eta, float gam, float lambda, float gb, int nu, int nv, int fly, int stride,
           float hypera, float hyperb, float epsilon, int tau, int noise_size, float temp,
           float mineta)
    : MF(train_data, test_data, result, model, dim, iter, eta, gam, lambda, gb, nu, nv, fly, stride),
      ur_(nullptr), vr_(nullptr), lambda_u_(nullptr), lambda_v_(nullptr), hyper_a_(hypera),
      hyper_b_(hyperb), temp_(temp), mineta_(mineta), noise_size_(noise_size), tau_(tau),
      epsilon_(epsilon), bound_(1.f), lambda_r_(1e0f), lambda_ub_(1e2f), lambda_vb_(1e2f),
      ntrain_(0), ntest_(0), seed_(0), round_(1) {}

DPMF::~DPMF() {}

void DPMF::init() {  // model.cc:197-245
  check(mfb_create(&ctx_, device_, nu_, nv_, dim_), "mfb_create");
  apply_options();
  check(mfb_enable(ctx_, 2), "mfb_enable(dpmf)");
  alloc_host(nu_ + nv_ + 2 * dim_);  // bu|bv|ur|vr|lambda_u|lambda_v, model.cc:199-204
  ur_ = bv_ + nv_;
  vr_ = ur_ + nu_;
  lambda_u_ = vr_ + nv_;
  lambda_v_ = lambda_u_ + dim_;
  seed_ = default_seed();
  check(mfb_init_normal(ctx_, seed_, 1e-2f), "mfb_init_normal");
  pull();
  for (int i = 0; i < 2 * dim_; i++) lambda_u_[i] = 1e2f;  // model.cc:226
  // model.cc:229-231 fills an 8 GB table with N(0,1) draws: not needed, the kernels evaluate a
  // counter-based Philox stream instead (noise_size_ is accepted for command-line compatibility)
  sample_train_and_precompute_weight();
  if (tau_ <= 0) tau_ = nv_;                                // model.cc:239
  bound_ = mfb_dp_bound(epsilon_, tau_, nv_);               // model.cc:240-242
}

void DPMF::sample_train_and_precompute_weight() {  // model.cc:263-297
  load_train();  // with dpmf enabled, finalize also builds the static logical clock
  int32_t n = 0;
  check(mfb_dp_weights(ctx_, train_ds_, &n), "mfb_dp_weights");
  ntrain_ = n;
  check(mfb_download(ctx_, MFB_UR, ur_, 0, nu_, 1), "download ur");
  check(mfb_download(ctx_, MFB_VR, vr_, 0, nv_, 1), "download vr");
}

void DPMF::seteta_cutoff(int round) { eta_ = mfb_seteta_cutoff(eta0_, round, gam_, mineta_); }  // model.cc:350-352

mfb_sgld_params DPMF::params() const {
  mfb_sgld_params p;
  memset(&p, 0, sizeof p);
  p.eta = eta_;
  p.temp = temp_;
  p.bound = bound_;
  p.ntrain = ntrain_;
  p.lambda_r = lambda_r_;
  p.lambda_ub = lambda_ub_;
  p.lambda_vb = lambda_vb_;
  p.seed = seed_;
  p.round = (uint32_t)round_;
  return p;
}

void DPMF::sgld_epoch() {
  check(mfb_upload(ctx_, MFB_LAMBDA_U, lambda_u_, 0, dim_, 1), "upload lambda_u");
  check(mfb_upload(ctx_, MFB_LAMBDA_V, lambda_v_, 0, dim_, 1), "upload lambda_v");
  const mfb_sgld_params p = params();
  check(mfb_sgld_epoch(ctx_, train_ds_, &p, gb_, data_in_fly_ <= 1 ? MFB_MODE_ORDERED : MFB_MODE_HOGWILD),
        "mfb_sgld_epoch");
}

void DPMF::finish_noise() {  // model.cc:312-332
  const mfb_sgld_params p = params();
  check(mfb_sgld_flush_noise(ctx_, train_ds_, &p), "mfb_sgld_flush_noise");
}

// util.h:103-154: the reference's samplers, driven by glibc rand() like the original so that a
// run is reproducible in the same way (rand() is never seeded there: srand(1) sequence).
static float uniform_open0() { return (float)((double)(float)rand() / ((double)(float)RAND_MAX + 1.0)); }
static float uniform_open01() { return (float)(((double)(float)rand() + 1.0) / ((double)(float)RAND_MAX + 2.0)); }
static float polar_normal() {
  for (;;) {
    const float x = (float)((double)(2 * uniform_open01()) - 1.0);
    const float y = (float)((double)(2 * uniform_open01()) - 1.0);
    const float r2 = x * x + y * y;
    if (r2 < 1.0 && r2 != 0.0) return (float)((double)x * sqrt(-2.0 * (double)logf(r2) / (double)r2));
  }
}
static float gamma_draw(float shape, float rate) {  // Marsaglia-Tsang, util.h:126-148
  if (shape < 1.0) {
    float u;
    do u = uniform_open0(); while (u == 0.0);
    return (float)((double)gamma_draw((float)((double)shape + 1.0), rate) * pow((double)u, 1.0 / (double)shape));
  }
  const float d = (float)((double)shape - 1.0 / 3.0);
  const float c = (float)(1.0 / sqrt(9.0 * (double)d));
  for (;;) {
    float x, v;
    do {
      x = polar_normal();
      v = (float)(1.0 + (double)(c * x));
    } while (v <= 0.0);
    v = v * v * v;
    const float u = uniform_open0();
    const bool squeeze = (double)u < 1.0 - 0.0331 * (double)(x * x) * (double)(x * x);
    if (squeeze || (double)logf(u) < 0.5 * (double)x * (double)x + (double)d * (1.0 - (double)v + (double)logf(v)))
      return d * v / rate;
  }
}
static void gamma_posterior(float& lambda, float prior_a, float prior_b, float sum_sqr, float count) {  // util.h:150-154
  lambda = gamma_draw((float)((double)prior_a + 0.5 * (double)count), (float)((double)prior_b + 0.5 * (double)sum_sqr));
}

void DPMF::sample_hyper(float mse) {  // model.cc:335-348; the reductions run on the GPU (K7)
  std::vector<double> normu(dim_), normv(dim_);
  double bu2 = 0, bv2 = 0;
  check(mfb_col_sqnorms(ctx_, normu.data(), normv.data(), &bu2, &bv2), "mfb_col_sqnorms");
  gamma_posterior(lambda_r_, hyper_a_, hyper_b_, mse, (float)ntrain_);
  gamma_posterior(lambda_ub_, hyper_a_, hyper_b_, (float)bu2, (float)nu_);
  gamma_posterior(lambda_vb_, hyper_a_, hyper_b_, (float)bv2, (float)nv_);
  for (int i = 0; i < dim_; i++) {
    gamma_posterior(lambda_u_[i], hyper_a_, hyper_b_, (float)normu[i], (float)nu_);
    gamma_posterior(lambda_v_[i], hyper_a_, hyper_b_, (float)normv[i], (float)nv_);
  }
}

void DPMF::finish_round(mf::Blocks& blocks_test, int round) {  // model.cc:299-310
  finish_noise();
  double sse = 0.0;
  int64_t ntr = 0;
  check(mfb_sse(ctx_, train_ds_, gb_, &sse, &ntr), "mfb_sse(train)");  // train_sample_ is the whole file
  const float mse = (float)sse;
  int nt;
  const float tmse = calc_mse(blocks_test, nt);
  printf("round #%d\tRMSE=%f\ttRMSE=%f\t", round, sqrt(mse * 1.0 / ntr), sqrt(tmse * 1.0 / nt));
  sample_hyper(mse);
  seteta_cutoff(round + 1);
  round_ = round + 1;
  e = Time::now();
  printf("%f\n", std::chrono::duration<float>(e - s).count());
  if (round >= 100 && round % 20 == 0) save_model(round);
}

// model.cc:123-195: the DPMF checkpoint inserts lambda_r, lambda_ub, lambda_vb, lambda_u[dim],
// lambda_v[dim] where MF has the single lambda
void DPMF::save_model(int round) {
  pull();
  char file[512];
  snprintf(file, sizeof file, "%s_%d", result_, round);
  FILE* fp = fopen(file, "wb");
  if (!fp) die_io("create", file);
  fwrite(&nv_, 4, 1, fp);
  fwrite(&nu_, 4, 1, fp);
  fwrite(&dim_, 4, 1, fp);
  fwrite(&lambda_r_, 4, 1, fp);
  fwrite(&lambda_ub_, 4, 1, fp);
  fwrite(&lambda_vb_, 4, 1, fp);
  fwrite(lambda_u_, sizeof(float), dim_, fp);
  fwrite(lambda_v_, sizeof(float), dim_, fp);
  write_factors(fp, this);
  fclose(fp);
  write_state(file, round);
}

void DPMF::read_hyper() {  // model.cc:153-167
  FILE* fp = fopen(model_, "rb");
  if (!fp) die_io("open", model_);
  read_header(fp, model_, this);
  must_read(&lambda_r_, 4, fp, model_);
  must_read(&lambda_ub_, 4, fp, model_);
  must_read(&lambda_vb_, 4, fp, model_);
  must_read(lambda_u_, sizeof(float) * dim_, fp, model_);
  must_read(lambda_v_, sizeof(float) * dim_, fp, model_);
  fclose(fp);
}

void DPMF::read_model() {  // model.cc:169-195
  FILE* fp = fopen(model_, "rb");
  if (!fp) die_io("open", model_);
  read_header(fp, model_, this);
  must_read(&lambda_r_, 4, fp, model_);
  must_read(&lambda_ub_, 4, fp, model_);
  must_read(&lambda_vb_, 4, fp, model_);
  must_read(lambda_u_, sizeof(float) * dim_, fp, model_);
  must_read(lambda_v_, sizeof(float) * dim_, fp, model_);
  read_factors(fp, model_, this);
  fclose(fp);
  push();
  if (read_state(model_)) round_ = start_round_ + 1;
}

// ---------------------------------------------------------------------------------- AdaptRegMF
AdaptRegMF::AdaptRegMF(char* train_data, char* test_data, char* valid_data, char* result,
                       char* model, int dim, int iter, float eta, float gam, float lambda, float gb,
                       int nu, int nv, int fly, int stride, int loss, int measure, float eta_reg)
    : MF(train_data, test_data, result, model, dim, iter, eta, gam, lambda, gb, nu, nv, fly, stride),
      valid_data_(valid_data), eta_reg_(eta_reg), eta0_reg_(eta_reg), loss_(loss), measure_(measure),
      lam_u_(lambda), lam_v_(lambda), lam_bu_(lambda), lam_bv_(lambda) {}  // model.h:82

AdaptRegMF::~AdaptRegMF() {}

void AdaptRegMF::init1() {  // model.cc:355-383
  init();
  check(mfb_enable(ctx_, 1), "mfb_enable(admf)");
  check(mfb_snapshot_old(ctx_), "mfb_snapshot_old");
  const float lams[4] = {lam_u_, lam_v_, lam_bu_, lam_bv_};
  check(mfb_admf_set_lams(ctx_, lams), "mfb_admf_set_lams");
}

void AdaptRegMF::set_etareg(int round) { eta_reg_ = mfb_seteta(eta0_reg_, round, gam_); }  // model.cc:386-388

void AdaptRegMF::plain_read_valid(const char* valid) {  // model.cc:390-415
  mfb_blocks* b = nullptr;
  check(mfb_blocks_read(valid, &b), "plain_read_valid");
  const int64_t nruns = mfb_blocks_num_runs(b);
  const int32_t *uid = mfb_blocks_run_uid(b), *off = mfb_blocks_run_off(b), *vid = mfb_blocks_vid(b);
  const float* rating = mfb_blocks_rating(b);
  for (int64_t r = 0; r < nruns; r++)
    for (int32_t k = off[r]; k < off[r + 1]; k++) recsv_.push_back(Record{uid[r], vid[k], rating[k]});
  mfb_blocks_free(b);
  // std::random_shuffle (model.cc:413) as libstdc++ runs it: swap(a[i], a[rand() % (i+1)])
  for (size_t i = 1; i < recsv_.size(); i++) {
    const size_t j = (size_t)rand() % (i + 1);
    if (i != j) std::swap(recsv_[i], recsv_[j]);
  }
  if (recsv_.empty()) {
    fprintf(stderr, "mf_b200: validation file %s holds no records\n", valid);
    exit(3);
  }
  std::vector<int32_t> u(recsv_.size()), v(recsv_.size());
  std::vector<float> r(recsv_.size());
  for (size_t i = 0; i < recsv_.size(); i++) {
    u[i] = recsv_[i].u_;
    v[i] = recsv_[i].v_;
    r[i] = recsv_[i].r_;
  }
  check(mfb_admf_set_validation(ctx_, (int64_t)recsv_.size(), u.data(), v.data(), r.data()), "mfb_admf_set_validation");
}

float AdaptRegMF::calc_measure(const mf::Blocks& blocks, int& ndata) {
  double sse = 0.0;
  int64_t n = 0;
  check(mfb_sse_link(ctx_, dataset_of(blocks), gb_, (measure_ == 1 && loss_ == 1) ? 1 : 0, &sse, &n), "mfb_sse_link");
  ndata = (int)n;
  return (float)sse;
}

void AdaptRegMF::admf_epoch() {
  const int64_t nruns = mfb_dataset_num_runs(ctx_, train_ds_);
  std::vector<int32_t> draws((size_t)nruns);
  for (int64_t r = 0; r < nruns; r++) draws[r] = (int32_t)((size_t)rand() % recsv_.size());  // admf.h:82
  check(mfb_admf_set_draws(ctx_, nruns, draws.data()), "mfb_admf_set_draws");
  check(mfb_admf_epoch(ctx_, train_ds_, eta_, eta_reg_, loss_, gb_, schedule()), "mfb_admf_epoch");
  float lams[4];
  check(mfb_admf_get_lams(ctx_, lams), "mfb_admf_get_lams");
  lam_u_ = lams[0];
  lam_v_ = lams[1];
  lam_bu_ = lams[2];
  lam_bv_ = lams[3];
}

// ------------------------------------------------------------------------------------ drivers
// main.cc:36-52 + SgdReadFilter's end-of-file branch (mf.h:32-45).  Epochs are clean barriers
// (the reference changes eta while earlier blocks may still be in flight; with --fly 1 the two
// coincide).  The printed clock is cumulative since the first epoch started and, as in the
// reference, includes the evaluation passes of earlier epochs.
// MF_TIMING=1: where the start-up goes (stderr)
static void lap(const char* what) {
  static Time::time_point t0 = Time::now();
  const Time::time_point t1 = Time::now();
  if (getenv("MF_TIMING")) fprintf(stderr, "mf_b200: %-28s %.3f s\n", what, std::chrono::duration<float>(t1 - t0).count());
  t0 = t1;
}

void run(MF& mf) {
  lap("process start -> run()");
  mf.init();
  lap("init (context, fill, mirrors)");
  if (mf.model_ != NULL) mf.read_model();
  mf::Blocks blocks_test;
  plain_read(mf.test_data_, blocks_test);
  lap("test file");
  if (tile_ratings() <= 0 && !stream_ingest()) mf.load_train();  // (default: ingested by the first epoch itself)
  lap("training file (parse, ingest)");
  s = Time::now();
  for (int iter = mf.start_round_ + 1; iter <= mf.iter_; iter++) {  // (a model loaded with its .state continues)
    if (iter > 1) mf.seteta(iter);  // mf.h:38
    mf.sgd_epoch();
    check(mfb_sync(mf.ctx_), "mfb_sync");
    e = Time::now();
    int nn;
    const float sse = mf.calc_mse(blocks_test, nn);
    printf("iter#%d\t%f\ttRMSE=%f\n", iter, std::chrono::duration<float>(e - s).count(), sqrt(sse * 1.0 / nn));  // mf.h:35
    fflush(stdout);
    // the reference never calls MF::save_model on this path; opt-in checkpoints
    if (mf.result_ != NULL && getenv("MF_SAVE_EVERY") && iter % atoi(getenv("MF_SAVE_EVERY")) == 0) mf.save_model(iter);
  }
}

void run(DPMF& dpmf) {  // main.cc:55-74
  dpmf.init();
  if (dpmf.model_ != NULL) {
    // main.cc:57 loads the hyper-parameters only; a checkpoint that has its .state sidecar (written by
    // DPMF::save_model here) is a resume point: factors, hyper-parameters, round and step size continue
    char st[600];
    snprintf(st, sizeof st, "%s.state", dpmf.model_);
    FILE* f = fopen(st, "r");
    if (f) {
      fclose(f);
      dpmf.read_model();
      dpmf.seteta_cutoff(dpmf.start_round_ + 1);
    } else {
      dpmf.read_hyper();
    }
  }
  mf::Blocks blocks_test;
  plain_read(dpmf.test_data_, blocks_test);
  s = Time::now();
  for (int i = dpmf.start_round_ + 1; i <= dpmf.iter_; i++) {
    dpmf.sgld_epoch();
    dpmf.finish_round(blocks_test, i);
    fflush(stdout);
    if (dpmf.result_ != NULL && getenv("MF_SAVE_EVERY") && i % atoi(getenv("MF_SAVE_EVERY")) == 0 &&
        !(i >= 100 && i % 20 == 0))  // (finish_round saved that one already, model.cc:309)
      dpmf.save_model(i);
  }
}

void run(AdaptRegMF& admf) {  // main.cc:77-93 + AdRegReadFilter (admf.h:29-37)
  admf.init1();
  mf::Blocks blocks_test;
  plain_read(admf.test_data_, blocks_test);
  admf.plain_read_valid(admf.valid_data_);
  admf.load_train();
  s = Time::now();
  for (int iter = 1; iter <= admf.iter_; iter++) {
    if (iter > 1) {
      admf.seteta(iter);      // admf.h:35
      admf.set_etareg(iter);  // admf.h:36
    }
    admf.admf_epoch();
    e = Time::now();
    int nn;
    const float sse = admf.calc_measure(blocks_test, nn);  // --measure 0: MF::calc_mse, as admf.h:32 calls it
    printf("iter#%d\t%f\ttRMSE=%f\n", iter, std::chrono::duration<float>(e - s).count(), sqrt(sse * 1.0 / nn));  // admf.h:32
    if (getenv("MF_PRINT_LAMBDA"))  // the reference never prints them; opt-in extra line
      printf("lambda#%d\t%g\t%g\t%g\t%g\n", iter, admf.lam_u_, admf.lam_v_, admf.lam_bu_, admf.lam_bv_);
    fflush(stdout);
  }
}
