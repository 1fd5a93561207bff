/* TEST INFRASTRUCTURE - CPU oracle for the blocked SGD matrix-factorization epoch.
 *
 * A plain-C restatement of the reference's algorithm for the hot path (cjolivier01/experimental-mf,
 * citations are file:line under /root/reference/src).  It is the checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (experimental-mf_b200/) never includes, links or calls anything here.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below bit-for-bit
 * against the reference's own sources compiled in place (oracle/_ref/libmf_ref.so, built by
 * oracle/Makefile; TBB/MKL/protobuf replaced by oracle/shim/), and tests/golden/ holds outputs
 * of that reference build (generator: tests/golden/make_golden.py) for boxes without
 * /root/reference.  The reference itself ships no tests or golden vectors.
 *
 * Execution order is the reference's single-thread order: blocks in file order, users in block
 * order, records in user order (== `./mf --fly 1`).
 */
#ifndef MF_ORACLE_H
#define MF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Dense model.  Row i of theta is theta + i*stride, stride >= dim (the reference uses
 * stride = padding(dim), util.h:163-165). */
typedef struct {
  int32_t nu, nv, dim, stride;
  float* theta; /* [nu][stride] */
  float* phi;   /* [nv][stride] */
  float* bu;    /* [nu] */
  float* bv;    /* [nv] */
} mfo_model;

/* A rating file in file order: blocks -> user-runs -> records (blocks.proto:1-18). */
typedef struct {
  int64_t nblocks;
  const int64_t* block_off; /* [nblocks+1] first run of each block */
  int64_t nruns;
  const int32_t* run_uid;   /* [nruns] */
  const int64_t* run_off;   /* [nruns+1] first record of each run */
  const int32_t* vid;       /* [nratings] */
  const float* rating;      /* [nratings] */
} mfo_data;

/* util.h:163-165 */
int mfo_padding(int dim);
/* model.cc:36-38 ; model.cc:350-352 ; model.cc:386-388 (same formula on eta0_reg) */
float mfo_seteta(float eta0, int round, float gam);
float mfo_seteta_cutoff(float eta0, int round, float gam, float mineta);

/* ---- blocks.proto reader / writer (wire format: blocks.pb.cc:267,281,526,540,786; framing
 * [u32 size][Block bytes]: getdata.cc:100-103, util.h:76-88).  Returned data is malloc'ed. */
typedef struct {
  mfo_data d;
  int64_t nratings;
} mfo_file;
mfo_file* mfo_read_blocks(const char* path);
void mfo_free_file(mfo_file* f);
int mfo_write_blocks(const char* path, const mfo_data* d);

/* ---- plain SGD: SgdFilter::operator(), mf.h:76-132 ---- */
void mfo_sgd_epoch(mfo_model* m, const mfo_data* d, float eta, float lambda, float gb);

/* ---- evaluation: MF::calc_mse, model.cc:41-73.  Returns the SUM of squared errors in fp32
 * (per-block fp32 partials added in block order == the reference with one OpenMP thread). */
float mfo_sse(const mfo_model* m, const mfo_data* d, float gb, int64_t* ndata);
/* calc_mse with the link of --loss applied to the prediction first (util.h:90-95, admf.h:69, model.h:87) */
float mfo_sse_link(const mfo_model* m, const mfo_data* d, float gb, int loss, int64_t* ndata);

/* ---- SGLD / DP: SgldFilter::operator(), dpmf.h:41-91, + model.cc:197-352 ---- */
typedef struct {
  float eta, temp, bound;
  int32_t ntrain;
  float lambda_r, lambda_ub, lambda_vb;
  float* lambda_u; /* [dim] */
  float* lambda_v; /* [dim] */
  float* ur;       /* [nu] */
  float* vr;       /* [nv] */
  uint64_t gcount;
  uint64_t* gcountu; /* [nu] */
  uint64_t* gcountv; /* [nv] */
} mfo_dp_state;

/* Noise source.  kind 0: the `dim`+1 values applied to user row `row`, kind 1: to item row.
 * `t` is the logical clock gc of the rating (dpmf.h:62), `j` the record's index inside its
 * user-run, or -1 when called from finish_noise.  Must fill out[0..dim] (dim+1 values). */
typedef void (*mfo_noise_fn)(void* ctx, int kind, int32_t row, int64_t t, int32_t j, int32_t dim,
                             float* out);

/* The reference's table source (dpmf.h:53-54,67-70,87; model.cc:316-330) with every drawn
 * offset equal to `offset` - what ref_dpmf_set_offset() arranges in the reference build. */
typedef struct {
  const float* table;
  int64_t size;
  int32_t offset;
} mfo_noise_table;
void mfo_noise_from_table(void* ctx, int kind, int32_t row, int64_t t, int32_t j, int32_t dim,
                          float* out);

/* Counter-based source: Philox4x32-10 + Box-Muller, the layout the CUDA kernels use
 * (DESIGN.md "SGLD noise stream").  ctx = mfo_noise_philox*. */
typedef struct {
  uint64_t seed;
  uint32_t round;
} mfo_noise_philox;
void mfo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void mfo_noise_from_philox(void* ctx, int kind, int32_t row, int64_t t, int32_t j, int32_t dim,
                           float* out);
/* the 4 normals of one Philox block (exposed for the known-answer / moment tests) */
void mfo_philox_normal4(uint64_t seed, uint32_t round, int kind, int32_t row, int64_t t,
                        uint32_t chunk, float out[4]);
/* the bias normal of (kind,row,t): Box-Muller on the low bytes of chunks 0 and 1 (mf_oracle.c) */
float mfo_philox_bias_normal(uint64_t seed, uint32_t round, int kind, int32_t row, int64_t t);

/* model.cc:240-242 */
float mfo_dp_bound(float epsilon, int tau);
/* model.cc:263-297: ur[u] = ntrain/count(u), vr[v] = ntrain/count(v); returns ntrain */
int32_t mfo_dp_weights(const mfo_data* d, int nu, int nv, float* ur, float* vr);
void mfo_sgld_epoch(mfo_model* m, const mfo_data* d, mfo_dp_state* st, float gb,
                    mfo_noise_fn noise, void* noise_ctx);
/* DPMF::finish_noise, model.cc:312-332 */
void mfo_finish_noise(mfo_model* m, mfo_dp_state* st, mfo_noise_fn noise, void* noise_ctx);
/* DPMF::sample_hyper, model.cc:335-348 (+ util.h:103-161); consumes glibc rand() */
void mfo_sample_hyper(const mfo_model* m, mfo_dp_state* st, float hyper_a, float hyper_b,
                      float train_sse);
/* util.h:126-154 on its own (for the moment tests) */
float mfo_sample_gamma(float alpha, float beta);

/* ---- adaptive regulariser: AdRegFilter::operator(), admf.h:52-86, + model.h:86-102 ---- */
typedef struct {
  float eta, eta_reg;
  int32_t loss;
  float lam_u, lam_v, lam_bu, lam_bv;
  float* theta_old; /* [nu][stride] */
  float* phi_old;   /* [nv][stride] */
  float* bu_old;    /* [nu] */
  float* bv_old;    /* [nv] */
  int64_t nvalid;
  const int32_t* val_u;
  const int32_t* val_v;
  const float* val_r;
  /* if non-NULL, record i of the epoch's per-user draws is taken from draws[draw_pos++] instead
   * of rand() % nvalid (admf.h:82) */
  const int32_t* draws;
  int64_t draw_pos;
} mfo_ad_state;

/* AdaptRegMF::plain_read_valid, model.cc:390-415: flatten + std::random_shuffle (libstdc++:
 * for i in 1..n-1: swap(a[i], a[rand() % (i+1)])).  Arrays are shuffled in place. */
void mfo_shuffle_valid(int64_t n, int32_t* u, int32_t* v, float* r);
void mfo_admf_epoch(mfo_model* m, const mfo_data* d, mfo_ad_state* st, float gb);

void mfo_srand(unsigned seed);
/* n times rand() % nvalid (admf.h:82) */
void mfo_rand_draws(int64_t n, int64_t nvalid, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif
