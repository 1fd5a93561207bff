// Host-side mirror of the reference's model API (reference src/model.h:6-118) on top of the C ABI
// in include/mf_b200.h.  Same class names, constructor signatures, public members and method
// names, so code written against the reference's MF / DPMF / AdaptRegMF reads the same here; the
// factor matrices live in B200 HBM and the public host arrays are MIRRORS of them.
//
// Differences a caller can observe (all forced by the device boundary):
//   * theta_/phi_/bu_/bv_ are host copies: pull() refreshes them from the device, push() uploads
//     edits.  read_model() pushes, save_model() pulls, so checkpoint code needs no change.
//   * the training file is parsed ONCE (ingest) instead of once per epoch (mf.h:38-45);
//   * `--fly 1` selects the ordered schedule (the reference's single-thread update order, bit-exact with
//     the CPU oracle); any `--fly N > 1` - the reference's default is 8 - selects the parallel production
//     schedule at the width the library's own bounds allow (what bench.py measures).  The reference's N
//     CPU threads have no counterpart on a GPU; MF_ROW_CONCURRENCY overrides the hot-row budget;
//   * save_model() also writes `<file>.state` (round, eta0, gam): read_model() picks it up, and the run()
//     drivers then continue with the NEXT round's step size instead of starting over;
//   * errors are reported (message + exit code 3) instead of being ignored.
#ifndef MFB_MODEL_H
#define MFB_MODEL_H

#include <stdint.h>

#include <vector>

#include "../../include/mf_b200.h"

#ifndef CACHE_LINE_SIZE
#define CACHE_LINE_SIZE 64
#endif

typedef struct {
  int u_, v_;
  float r_;
} Record;  // reference util.h:43-46

namespace mf {
// Stand-in for the protobuf type the reference passes to calc_mse / finish_round
// (blocks.proto:15-17): a parsed rating file.  Device residency is cached per model.
class Blocks {
 public:
  Blocks() : blocks_(nullptr), ds_(-1), owner_(nullptr) {}
  ~Blocks();
  Blocks(const Blocks&) = delete;
  Blocks& operator=(const Blocks&) = delete;
  int64_t ratings() const;
  mfb_blocks* blocks_;
  mutable int ds_;          // dataset id inside owner_'s context once uploaded
  mutable mfb_ctx* owner_;
};
}  // namespace mf

// util.h:76-88
void plain_read(const char* data, mf::Blocks& blocks);

class MF {
 public:
  MF(char* train_data, char* test_data, char* result, char* model, int dim, int iter, float eta,
     float gam, float lambda, float gb, int nu, int nv, int fly, int stride);
  virtual ~MF();
  void init();
  float calc_mse(const mf::Blocks& blocks, int& ndata);
  void read_model();
  void save_model(int round);
  void seteta(int round);
  float **theta_, **phi_, *bu_, *bv_;
  const char *const train_data_, *const test_data_, *const result_, *const model_;
  float gb_;
  int dim_, iter_;
  float eta_, gam_;
  float lambda_, eta0_;
  int nu_, nv_, data_in_fly_, prefetch_stride_;

  // ---- B200 side ----
  mfb_ctx* ctx_;
  int start_round_;       // rounds already applied to a model loaded with its .state sidecar (0: fresh)
  void write_state(const char* file, int round) const;
  bool read_state(const char* file);
  void apply_options();   // MF_ROW_CONCURRENCY, model age of a resumed run
  int train_ds_;          // the training file as SoA tiles in HBM
  int device_;
  void pull();            // HBM -> theta_/phi_/bu_/bv_
  void push();            // theta_/phi_/bu_/bv_ -> HBM
  void load_train();      // one-time ingest of train_data_
  virtual bool needs_host_records() const { return false; }  // DPMF: its logical clock is built from the host copy
  int schedule() const;   // MFB_MODE_ORDERED for --fly 1, else the parallel schedule
  void sgd_epoch();       // SgdFilter over the whole file with the current eta_
  int dataset_of(const mf::Blocks& blocks);

 protected:
  void alloc_host(int extra_floats);
  std::vector<float> theta_store_, phi_store_, bias_store_;
  std::vector<float*> row_ptrs_;
};

class DPMF : public MF {
 public:
  DPMF(char* train_data, char* test_data, char* result, char* model, int dim, int iter, float eta,
       float gam, float lambda, float gb, int nu, int nv, int fly, int stride, float hypera,
       float hyperb, float epsilon, int tau, int noise_size, float temp, float mineta);
  ~DPMF();
  void init();
  void sample_train_and_precompute_weight();
  void seteta_cutoff(int round);
  void read_model();
  void read_hyper();
  void save_model(int round);
  void finish_noise();
  void finish_round(mf::Blocks& blocks_test, int round);
  void sample_hyper(float mse);
  float *ur_, *vr_, *lambda_u_, *lambda_v_;
  const float hyper_a_, hyper_b_;
  float temp_, mineta_;
  int noise_size_, tau_;  // noise_size_ is accepted and ignored: the table is replaced by Philox
  float epsilon_, bound_;
  float lambda_r_, lambda_ub_, lambda_vb_;
  int ntrain_, ntest_;
  // ---- B200 side ----
  uint64_t seed_;
  int round_;
  void sgld_epoch();      // SgldFilter over the whole file
  mfb_sgld_params params() const;
  bool needs_host_records() const override { return true; }
};

class AdaptRegMF : public MF {
 public:
  AdaptRegMF(char* train_data, char* test_data, char* valid_data, char* result, char* model,
             int dim, int iter, float eta, float gam, float lambda, float gb, int nu, int nv,
             int fly, int stride, int loss, int measure, float eta_reg);
  ~AdaptRegMF();
  void init1();
  void set_etareg(int round);
  void plain_read_valid(const char* valid);
  std::vector<Record> recsv_;
  const char* valid_data_;
  float eta_reg_, eta0_reg_;
  int loss_, measure_;
  float lam_u_, lam_v_, lam_bu_, lam_bv_;  // refreshed from the device after every epoch
  // ---- B200 side ----
  // the evaluation `--measure` selects: 0 = MF::calc_mse as the reference runs it (the link of --loss ignored,
  // model.cc:62), 1 = the prediction goes through active() first (util.h:90-95) - what --loss 1 trains for
  float calc_measure(const mf::Blocks& blocks, int& ndata);
  void admf_epoch();      // AdRegFilter over the whole file + one updateReg per user
};

void run(MF& mf);
void run(DPMF& dpmf);
void run(AdaptRegMF& admf);

#endif
