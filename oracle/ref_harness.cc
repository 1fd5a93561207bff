// TEST INFRASTRUCTURE (oracle/): C-ABI harness around the UNMODIFIED reference sources.
//
// This file is compiled together with /root/reference/src/model.cc, and #includes
// /root/reference/src/{model.h,mf.h,dpmf.h,admf.h} where they lie (-I/root/reference/src), into
// oracle/_ref/libmf_ref.so (see oracle/Makefile).  Nothing from the reference is copied into the
// repo.  TBB / MKL / protobuf are replaced by the shims in oracle/shim/ (this image has none of
// them); everything else - the filter bodies, model init, calc_mse, finish_noise, sample_hyper,
// updateReg - is the reference's own code.
//
// What the harness does: drives the reference's hot-path operators
//   SgdFilter::operator()  (mf.h:76),  SgldFilter::operator() (dpmf.h:41),
//   AdRegFilter::operator() (admf.h:52),  MF::calc_mse (model.cc:41)
// one block at a time in file order (== `--fly 1`, the single-thread update order), one epoch
// per call, and lets a test overwrite / read back the public model arrays (model.h:23) so that
// the reference's clock-seeded racy init (model.cc:3-6,22-33) does not enter the comparison.
// It is used (a) to pin oracle/mf_oracle.c, (b) to generate tests/golden/*.npz.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
#include "model.h"
#include "mf.h"
#include "dpmf.h"
#include "admf.h"

#include <omp.h>

namespace {

enum Kind { K_MF = 0, K_DPMF = 1, K_ADMF = 2 };

struct Handle {
  int kind;
  MF* mf;          // also the DPMF / AdaptRegMF object (both derive from MF)
  DPMF* dp;
  AdaptRegMF* ad;
  mf::Blocks train, test;
  char *train_path, *test_path, *valid_path;
};

char* dup_or_null(const char* s) { return s ? strdup(s) : NULL; }
// result_/model_ are `const char* const` members fixed by the constructors (model.h:24): the next ref_create_*
// call passes these (ref_set_io_paths), so that save_model / read_model / read_hyper can be exercised
char *g_result = NULL, *g_model = NULL;

}  // namespace

extern "C" {

void ref_srand(unsigned seed) { srand(seed); }
void ref_seed_generator(unsigned seed) { generator.seed(seed); }

void* ref_create_mf(const char* train, const char* test, int dim, float eta, float gam,
                    float lambda, float gb, int nu, int nv) {
  omp_set_num_threads(1);
  Handle* h = new Handle();
  h->kind = K_MF;
  h->train_path = dup_or_null(train);
  h->test_path = dup_or_null(test);
  h->valid_path = NULL;
  h->mf = new MF(h->train_path, h->test_path, g_result, g_model, dim, 1, eta, gam, lambda, gb, nu, nv,
                 /*fly*/ 1, /*stride*/ 2);
  h->dp = NULL;
  h->ad = NULL;
  h->mf->init();
  plain_read(h->train_path, h->train);
  if (h->test_path) plain_read(h->test_path, h->test);
  return h;
}

void* ref_create_dpmf(const char* train, const char* test, int dim, float eta, float gam,
                      float lambda, float gb, int nu, int nv, float hypera, float hyperb,
                      float epsilon, int tau, int noise_size, float temp, float mineta) {
  omp_set_num_threads(1);
  Handle* h = new Handle();
  h->kind = K_DPMF;
  h->train_path = dup_or_null(train);
  h->test_path = dup_or_null(test);
  h->valid_path = NULL;
  h->dp = new DPMF(h->train_path, h->test_path, g_result, g_model, dim, 1, eta, gam, lambda, gb, nu, nv,
                   1, 2, hypera, hyperb, epsilon, tau, noise_size, temp, mineta);
  h->mf = h->dp;
  h->ad = NULL;
  h->dp->init();  // also loads the whole training file into train_sample_ (model.cc:263-297)
  // model.cc:235 `new std::atomic<uint64>[nv_]` leaves the item clocks uninitialised (C++11
  // atomics have a trivial default constructor).  At realistic nv the array comes from a fresh
  // mmap and is zero; at test sizes it is recycled heap and sqrt(negative) poisons the model.
  // Zero it, which is what finish_noise (model.cc:326) establishes from epoch 2 on anyway.
  for (int i = 0; i < nv; i++) h->dp->gcountv[i] = 0;
  plain_read(h->train_path, h->train);
  if (h->test_path) plain_read(h->test_path, h->test);
  return h;
}

void* ref_create_admf(const char* train, const char* test, const char* valid, int dim, float eta,
                      float gam, float lambda, float gb, int nu, int nv, int loss,
                      float eta_reg) {
  omp_set_num_threads(1);
  Handle* h = new Handle();
  h->kind = K_ADMF;
  h->train_path = dup_or_null(train);
  h->test_path = dup_or_null(test);
  h->valid_path = dup_or_null(valid);
  h->ad = new AdaptRegMF(h->train_path, h->test_path, h->valid_path, NULL, NULL, dim, 1, eta, gam,
                         lambda, gb, nu, nv, 1, 2, loss, 0, eta_reg);
  h->mf = h->ad;
  h->dp = NULL;
  h->ad->init1();
  plain_read(h->train_path, h->train);
  if (h->test_path) plain_read(h->test_path, h->test);
  h->ad->plain_read_valid(h->valid_path);  // consumes rand() (std::random_shuffle, model.cc:413)
  return h;
}

int ref_num_valid(void* hv) { return (int)((Handle*)hv)->ad->recsv_.size(); }
void ref_get_valid(void* hv, int* u, int* v, float* r) {
  Handle* h = (Handle*)hv;
  for (size_t i = 0; i < h->ad->recsv_.size(); i++) {
    u[i] = h->ad->recsv_[i].u_;
    v[i] = h->ad->recsv_[i].v_;
    r[i] = h->ad->recsv_[i].r_;
  }
}

// dense row-major [n][dim] <-> the reference's row-pointer tables
void ref_set_factors(void* hv, const float* theta, const float* phi, const float* bu,
                     const float* bv) {
  Handle* h = (Handle*)hv;
  MF* m = h->mf;
  const int d = m->dim_;
  for (int i = 0; i < m->nu_; i++) memcpy(m->theta_[i], theta + (size_t)i * d, sizeof(float) * d);
  for (int i = 0; i < m->nv_; i++) memcpy(m->phi_[i], phi + (size_t)i * d, sizeof(float) * d);
  memcpy(m->bu_, bu, sizeof(float) * m->nu_);
  memcpy(m->bv_, bv, sizeof(float) * m->nv_);
  if (h->ad) {  // what init1 (model.cc:369-382) would have copied had init() produced these values
    for (int i = 0; i < m->nu_; i++) memcpy(h->ad->theta_old_[i], m->theta_[i], sizeof(float) * d);
    for (int i = 0; i < m->nv_; i++) memcpy(h->ad->phi_old_[i], m->phi_[i], sizeof(float) * d);
    memcpy(h->ad->bu_old_, m->bu_, sizeof(float) * (m->nu_ + m->nv_));
  }
}

void ref_get_factors(void* hv, float* theta, float* phi, float* bu, float* bv) {
  MF* m = ((Handle*)hv)->mf;
  const int d = m->dim_;
  for (int i = 0; i < m->nu_; i++) memcpy(theta + (size_t)i * d, m->theta_[i], sizeof(float) * d);
  for (int i = 0; i < m->nv_; i++) memcpy(phi + (size_t)i * d, m->phi_[i], sizeof(float) * d);
  memcpy(bu, m->bu_, sizeof(float) * m->nu_);
  memcpy(bv, m->bv_, sizeof(float) * m->nv_);
}

void ref_get_old(void* hv, float* theta_old, float* phi_old, float* bu_old, float* bv_old) {
  Handle* h = (Handle*)hv;
  MF* m = h->mf;
  const int d = m->dim_;
  for (int i = 0; i < m->nu_; i++)
    memcpy(theta_old + (size_t)i * d, h->ad->theta_old_[i], sizeof(float) * d);
  for (int i = 0; i < m->nv_; i++)
    memcpy(phi_old + (size_t)i * d, h->ad->phi_old_[i], sizeof(float) * d);
  memcpy(bu_old, h->ad->bu_old_, sizeof(float) * m->nu_);
  memcpy(bv_old, h->ad->bv_old_, sizeof(float) * m->nv_);
}

// replicates what the read filters do at an epoch boundary (mf.h:38, admf.h:35-36,
// model.cc:350-352) for epoch number `round` (1-based)
void ref_seteta(void* hv, int round) {
  Handle* h = (Handle*)hv;
  if (h->kind == K_DPMF) {
    h->dp->seteta_cutoff(round);
  } else {
    h->mf->seteta(round);
    if (h->ad) h->ad->set_etareg(round);
  }
}
float ref_get_eta(void* hv) { return ((Handle*)hv)->mf->eta_; }
float ref_get_etareg(void* hv) { return ((Handle*)hv)->ad->eta_reg_; }

// one pass over the training file in file order through the reference's update operator
void ref_epoch(void* hv) {
  Handle* h = (Handle*)hv;
  const int nb = h->train.block_size();
  if (h->kind == K_MF) {
    SgdFilter f(*h->mf);
    for (int b = 0; b < nb; b++) f((void*)h->train.mutable_block(b));
  } else if (h->kind == K_DPMF) {
    SgldFilter f(*h->dp);
    for (int b = 0; b < nb; b++) f((void*)h->train.mutable_block(b));
  } else {
    AdRegFilter f(*h->ad);
    for (int b = 0; b < nb; b++) f((void*)h->train.mutable_block(b));
  }
}

// MF::calc_mse: returns the SUM of squared errors (model.cc:41-73); which: 0 train, 1 test
float ref_calc_mse(void* hv, int which, int* ndata) {
  Handle* h = (Handle*)hv;
  int n = 0;
  float s = h->mf->calc_mse(which == 0 ? h->train : h->test, n);
  *ndata = n;
  return s;
}

// ---- DPMF only ---------------------------------------------------------------------------
void ref_dpmf_set_noise(void* hv, const float* table, int n) {
  DPMF* d = ((Handle*)hv)->dp;
  assert(n <= d->noise_size_);
  memcpy(d->noise_, table, sizeof(float) * (size_t)n);
}
// make every table offset drawn by dpmf.h:53-54 / model.cc:316,324 equal to c
void ref_dpmf_set_offset(void* hv, int c) {
  ((Handle*)hv)->dp->uniform_int_ = std::uniform_int_distribution<>(c, c);
}
void ref_dpmf_finish_noise(void* hv) { ((Handle*)hv)->dp->finish_noise(); }
void ref_dpmf_sample_hyper(void* hv, float mse) { ((Handle*)hv)->dp->sample_hyper(mse); }
// out: lambda_r, lambda_ub, lambda_vb, lambda_u[dim], lambda_v[dim]
void ref_dpmf_get_hyper(void* hv, float* out) {
  DPMF* d = ((Handle*)hv)->dp;
  out[0] = d->lambda_r_;
  out[1] = d->lambda_ub_;
  out[2] = d->lambda_vb_;
  memcpy(out + 3, d->lambda_u_, sizeof(float) * d->dim_);
  memcpy(out + 3 + d->dim_, d->lambda_v_, sizeof(float) * d->dim_);
}
void ref_dpmf_set_hyper(void* hv, const float* in) {
  DPMF* d = ((Handle*)hv)->dp;
  d->lambda_r_ = in[0];
  d->lambda_ub_ = in[1];
  d->lambda_vb_ = in[2];
  memcpy(d->lambda_u_, in + 3, sizeof(float) * d->dim_);
  memcpy(d->lambda_v_, in + 3 + d->dim_, sizeof(float) * d->dim_);
}
void ref_dpmf_get_weights(void* hv, float* ur, float* vr) {
  DPMF* d = ((Handle*)hv)->dp;
  memcpy(ur, d->ur_, sizeof(float) * d->nu_);
  memcpy(vr, d->vr_, sizeof(float) * d->nv_);
}
void ref_dpmf_info(void* hv, int* ntrain, float* bound, int* tau) {
  DPMF* d = ((Handle*)hv)->dp;
  *ntrain = d->ntrain_;
  *bound = d->bound_;
  *tau = d->tau_;
}

// ---- AdaptRegMF only ------------------------------------------------------------------------
void ref_admf_get_lams(void* hv, float* out4) {
  AdaptRegMF* a = ((Handle*)hv)->ad;
  out4[0] = a->lam_u_;
  out4[1] = a->lam_v_;
  out4[2] = a->lam_bu_;
  out4[3] = a->lam_bv_;
}

// ---- checkpoints (model.cc:75-195) -------------------------------------------------------------
void ref_set_io_paths(const char* result, const char* model) {
  g_result = dup_or_null(result);
  g_model = dup_or_null(model);
}
// MF::save_model / DPMF::save_model: writes "<result>_<round>"
void ref_save_model(void* hv, int round) {
  Handle* h = (Handle*)hv;
  if (h->kind == K_DPMF) h->dp->save_model(round);
  else h->mf->save_model(round);
}
// MF::read_model / DPMF::read_model from model_
void ref_read_model(void* hv) {
  Handle* h = (Handle*)hv;
  if (h->kind == K_DPMF) h->dp->read_model();
  else h->mf->read_model();
}
void ref_dpmf_read_hyper(void* hv) { ((Handle*)hv)->dp->read_hyper(); }
float ref_get_lambda(void* hv) { return ((Handle*)hv)->mf->lambda_; }

int ref_padding(int dim) { return padding(dim); }

// Objects are leaked on purpose: ~DPMF followed by ~MF frees theta_[0] twice (model.h:16,46).
void ref_destroy(void* hv) { (void)hv; }

}  // extern "C"
